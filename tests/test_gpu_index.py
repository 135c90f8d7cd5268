"""K4 index build on the GPU: bit-exact against a stable sort (numpy.argsort(kind="stable"))
and against the oracle's restatement of the reference transpose (matrix.cpp:617-692)."""
import numpy as np
import pytest

from conftest import bits_equal
from movie_recommender_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,groups", [(0, 5), (1, 1), (31, 3), (4096, 256), (4097, 257),
                                      (100000, 1), (100000, 70000), (1 << 20, 300),
                                      (3_000_000, 283228)])
def test_group_by_matches_stable_argsort(require_gpu, cpp_ls, n, groups):
    rng = np.random.default_rng(n + groups)
    # skewed keys (power law) so some groups are huge and many are empty
    key = np.minimum((rng.random(n) ** 3 * groups).astype(np.int32), groups - 1)
    ptr, idx = cpp_ls.group_by(key, groups)
    assert np.array_equal(ptr, np.concatenate([[0], np.cumsum(np.bincount(key, minlength=groups))]))
    assert np.array_equal(idx, np.argsort(key, kind="stable"))


def test_group_by_already_sorted_and_reversed(require_gpu, cpp_ls):
    key = np.repeat(np.arange(1000, dtype=np.int32), 37)
    _, idx = cpp_ls.group_by(key, 1000)
    assert np.array_equal(idx, np.arange(len(key)))
    _, idx = cpp_ls.group_by(key[::-1].copy(), 1000)
    assert np.array_equal(idx, np.argsort(key[::-1], kind="stable"))


@pytest.mark.parametrize("rows,cols,per", [(500, 60, 6), (20000, 3000, 11), (1000, 50, 50)])
def test_transpose_matches_oracle(require_gpu, cpp_ls, oracle, rows, cols, per):
    rowptr, col, vals, cols, *_ = synth.random_sparse_system(rows, cols, per, seed=rows)
    a = cpp_ls.csr_transpose(rows, cols, rowptr, col, vals)
    b = oracle.transpose(rows, cols, rowptr, col, vals)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and bits_equal(a[2], b[2])


def test_transpose_ragged_and_empty_rows(require_gpu, cpp_ls, oracle):
    rng = np.random.default_rng(5)
    deg = rng.integers(0, 9, size=700)
    deg[::7] = 0
    rowptr = np.concatenate([[0], np.cumsum(deg)]).astype(np.int32)
    col = rng.integers(0, 40, size=int(rowptr[-1])).astype(np.int32)  # duplicates allowed
    vals = rng.standard_normal(len(col))
    a = cpp_ls.csr_transpose(700, 40, rowptr, col, vals)
    b = oracle.transpose(700, 40, rowptr, col, vals)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and bits_equal(a[2], b[2])


def test_als_problem_index_is_the_oracle_grouping(require_gpu, cpp_ls, oracle):
    p = synth.als_problem(610, 9724, 100836, 10, min_degrees=False, shuffle=True)
    with cpp_ls.AlsProblem(p["user_ids"], p["item_ids"], p["ratings"], 10, 610, 9724) as prob:
        u_ptr, u_idx, i_ptr, i_idx = prob.get_index()
    ou = oracle.group_by(p["user_ids"], 610)
    oi = oracle.group_by(p["item_ids"], 9724)
    assert np.array_equal(u_ptr, ou[0]) and np.array_equal(u_idx, ou[1])
    assert np.array_equal(i_ptr, oi[0]) and np.array_equal(i_idx, oi[1])
