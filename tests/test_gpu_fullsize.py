"""BASELINE.json's FULL sizes, checked through size-independent properties (the oracle cannot
run whole sweeps at 27.75 M ratings in test time):

* K4 index: the groupings are stable sorts (sorted keys, ascending positions inside a group,
  a permutation, pointer array == histogram prefix sums);
* algorithm 4: an exact half-sweep satisfies the normal equations of sampled rows
  (|A^T (A x - b)| / |A^T b| < 1e-9), is idempotent, the SSE the factorisation reports equals a
  NumPy evaluation over all ratings, and the SSE decreases sweep after sweep;
* config 2 at full size: the reference-order solver is bit-identical to the oracle (which is
  pinned bit-exact to the real library) -- same 97 iterations the real reference takes;
* config 4 at full size: sampled queries bit-exact against the CPU definition; symmetric pairs
  carry bit-equal scores.
One data set is generated once per module (~30 s)."""
import numpy as np
import pytest

from conftest import bits_equal
from movie_recommender_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def c3():
    c = synth.CONFIGS["C3"]
    u, i = synth.rating_pairs(c["num_users"], c["num_items"], c["num_ratings"], c["k"] + 1, c["k"])
    raw = synth.planted_ratings(u, i, c["num_users"], c["num_items"], subtract_median=False)
    r = raw - synth.movie_medians(i, raw, c["num_items"])[i]
    uf0, if0 = synth.initial_factors(c["num_users"], c["num_items"], c["k"])
    return dict(c, user_ids=u, item_ids=i, ratings=r, raw=raw, uf0=uf0, if0=if0)


def test_c3_index_and_exact_sweeps(require_gpu, cpp_ls, c3):
    nu, ni, k, nnz = c3["num_users"], c3["num_items"], c3["k"], len(c3["ratings"])
    assert nnz == 27753444
    u, i, r = c3["user_ids"], c3["item_ids"], c3["ratings"]
    with cpp_ls.AlsProblem(u, i, r, k, nu, ni) as prob:
        u_ptr, u_idx, i_ptr, i_idx = prob.get_index()
        for key, ptr, idx, groups in ((u, u_ptr, u_idx, nu), (i, i_ptr, i_idx, ni)):
            ks = key[idx]
            assert np.all(np.diff(ks) >= 0)                                  # sorted by key
            assert np.all(np.diff(idx)[np.diff(ks) == 0] > 0)                # stable
            assert np.array_equal(ptr, np.concatenate([[0], np.cumsum(np.bincount(key, minlength=groups))]))
            seen = np.zeros(nnz, dtype=bool)
            seen[idx] = True
            assert seen.all()                                                # a permutation
        prob.set_factors(c3["uf0"], c3["if0"])
        sse = []
        for _ in range(3):
            sse.append(prob.run(4, -1e300, 1).last_rr)
        assert sse[0] > sse[1] > sse[2] > 0
        uf, itf = prob.get_factors()
        U, V = uf.reshape(nu, k + 1), itf.reshape(ni, k)
        # SSE reported by the factorisation == NumPy over all ratings
        tot = 0.0
        for s in range(0, nnz, 1 << 21):
            e = slice(s, s + (1 << 21))
            pred = np.einsum("ij,ij->i", U[u[e], :k], V[i[e]]) + U[u[e], k]
            tot += float(np.sum((pred - r[e]) ** 2))
        assert abs(tot - sse[2]) <= 1e-8 * tot
        # the movie half-sweep just done solved its normal equations exactly (sampled rows)
        rng = np.random.default_rng(0)
        for m in rng.choice(ni, size=40, replace=False):
            rows = i_idx[i_ptr[m]:i_ptr[m + 1]]
            A = U[u[rows], :k]
            b = r[rows] - U[u[rows], k]
            g = A.T @ b
            assert np.linalg.norm(A.T @ (A @ V[m]) - g) <= 1e-9 * np.linalg.norm(g)
        # idempotent: a second movie half-sweep from the solution changes nothing measurable
        prob.half_sweep(False, 0)
        _, itf2 = prob.get_factors()
        prob.shard_sse(0)
        _, itf2 = prob.get_factors()
        assert np.linalg.norm(itf2 - itf) <= 1e-9 * np.linalg.norm(itf)


def test_c2_full_size_bitexact(require_gpu, cpp_ls, oracle, c3):
    nu, ni = c3["num_users"], c3["num_items"]
    rowptr, col, vals, cols, b, x0 = synth.bias_model_system(c3["user_ids"], c3["item_ids"], c3["raw"], nu, ni)
    cpp_ls.set_thread_count(16)
    x, it, rr = cpp_ls.cg_least_squares(rowptr, col, vals, cols, b, x0=x0)
    xo, ito, rro = oracle.cg_least_squares(rowptr, col, vals, cols, b, x0, thread_count=16)
    assert it == ito and bits_equal(x, xo) and bits_equal([rr], [rro])


def test_c4_full_size_sampled_queries(require_gpu, oracle):
    from movie_recommender_b200 import similarity
    rng = np.random.default_rng(20181001)
    M = rng.standard_normal((53889, 50))
    ids, scores, info = similarity.factor_cosine_topk(M, topk=50)
    assert ids.shape == (53889, 50) and np.all(np.diff(scores, axis=1) <= 0)
    for q in rng.choice(53889, size=24, replace=False):
        oi, os_ = oracle.cosine_topk(M, 50, int(q), int(q) + 1)
        assert np.array_equal(ids[q], oi[0]) and bits_equal(scores[q], os_[0])
    # symmetry: whenever i lists j and j lists i, the two scores are the same bits
    checked = 0
    for q in range(0, 53889, 977):
        for pos, j in enumerate(ids[q]):
            back = np.flatnonzero(ids[j] == q)
            if len(back):
                assert scores[q, pos].view(np.uint64) == scores[j, back[0]].view(np.uint64)
                checked += 1
    assert checked > 0
