"""K5 factor-cosine top-k (config 4) against the CPU definition (oracle_cosine_topk; PARITY
UNPINNED -- the reference has no factor-based similarity).  ids AND scores must be
bit-exact: candidate selection runs on the tensor cores (default: tcgen05.mma kind::tf32 with TMA
operands and TMEM accumulators, csrc/similarity_tc.cu; ``MRB_SIM_KERNEL=dmma``: the fp64 mma.sync
kernel it replaced), the final scores are recomputed in the oracle's summation order, and a
certificate -- widened by the candidate GEMM's error bound -- triggers an exhaustive recomputation
when near-ties at the boundary could not be excluded.  Every test runs on both kernels."""
import os

import numpy as np
import pytest

from conftest import bits_equal

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True, params=["tcgen05", "dmma"])
def candidate_kernel(request):
    old = os.environ.get("MRB_SIM_KERNEL")
    if request.param != "tcgen05":
        os.environ["MRB_SIM_KERNEL"] = request.param
    else:
        os.environ.pop("MRB_SIM_KERNEL", None)
    yield request.param
    if old is None:
        os.environ.pop("MRB_SIM_KERNEL", None)
    else:
        os.environ["MRB_SIM_KERNEL"] = old


def sim():
    from movie_recommender_b200 import similarity
    return similarity


@pytest.mark.parametrize("n,k,topk", [(3000, 50, 50), (1000, 10, 20), (700, 33, 50), (257, 64, 56),
                                      (64, 3, 50), (40, 50, 50), (6000, 50, 50), (129, 8, 5)])
def test_topk_ids_and_scores_bitexact(require_gpu, oracle, n, k, topk):
    rng = np.random.default_rng(n + k)
    M = rng.standard_normal((n, k))
    ids, scores, info = sim().factor_cosine_topk(M, topk=topk)
    oi, os_ = oracle.cosine_topk(M, topk)
    assert np.array_equal(ids, oi)
    assert bits_equal(scores, os_)


@pytest.mark.parametrize("mode", ["two-pass", "tc1"])
def test_large_catalogue_two_pass_and_one_pass(require_gpu, oracle, mode, monkeypatch):
    """Catalogues of >= 128 tiles take the two-pass tcgen05 path (tile maxima -> fixed per-row
    threshold -> collect); MRB_SIM_KERNEL=tc1 forces the one-pass online top-64 kernel.  Whole
    result against properties, sampled query blocks against the oracle, incl. duplicates."""
    if os.environ.get("MRB_SIM_KERNEL") == "dmma":
        pytest.skip("the fp64 kernel is covered by the other tests")
    if mode == "tc1":
        monkeypatch.setenv("MRB_SIM_KERNEL", "tc1")
    n, k, topk = 17000, 50, 50
    rng = np.random.default_rng(17)
    M = rng.standard_normal((n, k))
    M[5000:5040] = M[5000]                      # 40 identical rows: ties at the top
    M[123] = 0.0
    ids, scores, info = sim().factor_cosine_topk(M, topk=topk)
    assert ids.shape == (n, topk) and np.all(ids != np.arange(n)[:, None])
    assert np.all(np.diff(scores, axis=1) <= 0)
    for lo, hi in [(0, 48), (4990, 5050), (11111, 11143), (n - 20, n)]:
        oi, os_ = oracle.cosine_topk(M, topk, lo, hi)
        assert np.array_equal(ids[lo:hi], oi) and bits_equal(scores[lo:hi], os_)
    part = sim().factor_cosine_topk(M, topk=topk, q_lo=4000, q_hi=4300)
    assert np.array_equal(part[0], ids[4000:4300]) and bits_equal(part[1], scores[4000:4300])


def test_duplicates_and_ties_force_the_exact_fallback(require_gpu, oracle):
    """30 identical movies + exact duplicates elsewhere: massive ties at the top-k boundary."""
    rng = np.random.default_rng(7)
    M = rng.standard_normal((500, 12))
    M[100:130] = M[100]          # 30 identical rows
    M[200:260] = 2.5 * M[200]    # 60 parallel rows (same direction, cosine exactly equal)
    M[300] = 0.0                 # a zero row (norm 0 -> all scores 0)
    ids, scores, info = sim().factor_cosine_topk(M, topk=50)
    oi, os_ = oracle.cosine_topk(M, 50)
    assert np.array_equal(ids, oi) and bits_equal(scores, os_)


def test_query_block_split_equals_whole(require_gpu):
    """The multi-GPU split is by query block: blocks computed separately == one call."""
    rng = np.random.default_rng(3)
    M = rng.standard_normal((1500, 50))
    ids, scores, _ = sim().factor_cosine_topk(M, topk=50)
    parts = [sim().factor_cosine_topk(M, topk=50, q_lo=lo, q_hi=hi)
             for lo, hi in [(0, 187), (187, 750), (750, 1500)]]
    assert np.array_equal(ids, np.concatenate([p[0] for p in parts]))
    assert bits_equal(scores, np.concatenate([p[1] for p in parts]))


def test_properties_on_als_item_factors(require_gpu, cpp_ls, oracle):
    """On real ALS output: sorted scores, self excluded, and sampled queries equal the oracle."""
    from movie_recommender_b200 import synth
    nu, ni, k = 3000, 2500, 20
    p = synth.als_problem(nu, ni, 200000, k, seed=9)
    _, itf, _ = cpp_ls.als(p["user_ids"], p["item_ids"], p["ratings"], k, nu, ni, -1e300, 2, 4,
                           user_factors=p["user_factors0"], item_factors=p["item_factors0"])
    ids, scores, info = sim().factor_cosine_topk(itf, num_factors=k, topk=50)
    assert ids.shape == (ni, 50) and np.all(ids != np.arange(ni)[:, None])
    assert np.all(np.diff(scores, axis=1) <= 0) and np.all(np.abs(scores) <= 1 + 1e-12)
    for lo in (0, 1234, ni - 40):
        oi, os_ = oracle.cosine_topk(itf.reshape(ni, k), 50, lo, lo + 40)
        assert np.array_equal(ids[lo:lo + 40], oi) and bits_equal(scores[lo:lo + 40], os_)
