"""The oracle (oracle/ls_oracle.c) against the REAL reference: bit-exact at several thread
counts when oracle/_ref/cpp_ls_lib.so is present, and against the committed golden vectors
(tests/golden/*.npz, produced by the real reference via tests/golden/make_golden.py)."""
import numpy as np
import pytest

from conftest import bits_equal, load_golden
from movie_recommender_b200 import synth

THREADS = (1, 4)


@pytest.mark.parametrize("name", ["ls_200x50", "ls_sparse"])
def test_oracle_ls_matches_golden(oracle, name):
    g = load_golden(name)
    for T in THREADS:
        for alg in (1, 2):
            x, it, rr = oracle.cg_least_squares(g["rowptr"], g["colidx"], g["vals"], int(g["cols"]),
                                                g["b"], g["x0"], algorithm=alg, thread_count=T)
            assert it == int(g["it_T%d_a%d" % (T, alg)])
            assert bits_equal(x, g["x_T%d_a%d" % (T, alg)])
            assert bits_equal([rr], [g["rr_T%d_a%d" % (T, alg)]])


def test_oracle_als_matches_golden(oracle):
    g = load_golden("als_planted")
    for T in THREADS:
        for alg in (1, 2):
            uf, itf, it = oracle.als(g["user_ids"], g["item_ids"], g["ratings"], int(g["k"]),
                                     g["uf0"], g["if0"], algorithm=alg, thread_count=T)
            assert it == int(g["it_T%d_a%d" % (T, alg)])
            assert bits_equal(uf, g["uf_T%d_a%d" % (T, alg)])
            assert bits_equal(itf, g["if_T%d_a%d" % (T, alg)])


def test_golden_als_quality():
    """The reference's own pass criterion on this case (cpp_ls_test.py:121-147): the model
    recovered from 80 % of the planted ratings reproduces them to a mean abs error < 0.15."""
    g = load_golden("als_planted")
    k = int(g["k"])
    uf = g["uf_T1_a1"].reshape(-1, k + 1)
    itf = g["if_T1_a1"].reshape(-1, k)
    pred = (uf[g["user_ids"], :k] * itf[g["item_ids"]]).sum(axis=1) + uf[g["user_ids"], k]
    assert np.mean(np.abs(pred - g["ratings"])) < 0.15


@pytest.mark.parametrize("T", [1, 3, 8])
@pytest.mark.parametrize("alg", [1, 2])
def test_oracle_als_bitexact_vs_reference(oracle, T, alg):
    if not oracle.has_ref():
        pytest.skip("oracle/_ref not built (reference sources absent)")
    p = synth.als_problem(120, 400, 6000, 6, seed=11, min_degrees=False, shuffle=(T == 3))
    a = oracle.ref_als(p["user_ids"], p["item_ids"], p["ratings"], 6, p["user_factors0"],
                       p["item_factors0"], -1e300, 3, alg, T)
    b = oracle.als(p["user_ids"], p["item_ids"], p["ratings"], 6, p["user_factors0"],
                   p["item_factors0"], -1e300, 3, alg, T)
    assert a[2] == b[2] == 3
    assert bits_equal(a[0], b[0]) and bits_equal(a[1], b[1])


def test_oracle_als_default_termination_vs_reference(oracle):
    if not oracle.has_ref():
        pytest.skip("oracle/_ref not built")
    p = synth.als_problem(80, 150, 4000, 4, seed=5)
    a = oracle.ref_als(p["user_ids"], p["item_ids"], p["ratings"], 4, p["user_factors0"],
                       p["item_factors0"], thread_count=4)
    b = oracle.als(p["user_ids"], p["item_ids"], p["ratings"], 4, p["user_factors0"],
                   p["item_factors0"], thread_count=4)
    assert a[2] == b[2]
    assert bits_equal(a[0], b[0]) and bits_equal(a[1], b[1])


@pytest.mark.parametrize("T", [1, 4, 7])
@pytest.mark.parametrize("alg", [1, 2])
def test_oracle_ls_bitexact_vs_reference(oracle, T, alg):
    if not oracle.has_ref():
        pytest.skip("oracle/_ref not built")
    rowptr, col, vals, cols, b, x0, _ = synth.random_sparse_system(3000, 120, 7, seed=3)
    xa, ia, ra = oracle.ref_cg_least_squares(rowptr, col, vals, cols, b, x0, algorithm=alg,
                                             thread_count=T)
    xb, ib, rb = oracle.cg_least_squares(rowptr, col, vals, cols, b, x0, algorithm=alg,
                                         thread_count=T)
    assert ia == ib and bits_equal(xa, xb) and bits_equal([ra], [rb])


def test_oracle_bias_model_thread_independent(oracle):
    """the bias-model LS terminates via rr < 1e-6 and is thread-count
    independent to round-off -- the property that makes config-2 parity well posed."""
    u, i = synth.rating_pairs(300, 200, 9000, 3, 3, seed=9)
    raw = synth.planted_ratings(u, i, 300, 200, seed=9, subtract_median=False)
    rowptr, col, vals, cols, b, x0 = synth.bias_model_system(u, i, raw, 300, 200, seed=9)
    x1, it1, rr1 = oracle.cg_least_squares(rowptr, col, vals, cols, b, x0, thread_count=1)
    x8, it8, rr8 = oracle.cg_least_squares(rowptr, col, vals, cols, b, x0, thread_count=8)
    assert it1 == it8 and rr1 < 1e-6
    pred1 = x1[u] + x1[300 + i]
    pred8 = x8[u] + x8[300 + i]
    # converged to the rr < 1e-6 stopping tolerance, far inside the 1e-4 parity budget
    assert np.max(np.abs(pred1 - pred8)) < 1e-5


def test_chunk_table_is_float32_formula(oracle):
    # matrix.cpp:12 -- the float32 rounding matters for large lengths
    for T, n in [(3, 10), (8, 27753444), (7, 14444628), (18, 100836), (5, 3)]:
        bd = oracle.chunk_table(T, n)
        exp = [0] + [int(np.float32(np.float32(i) / np.float32(T)) * np.float32(n)) for i in range(1, T)] + [n]
        assert list(bd) == exp


def test_group_by_and_transpose_are_stable(oracle):
    rng = np.random.default_rng(0)
    key = rng.integers(0, 37, size=5000).astype(np.int32)
    ptr, idx = oracle.group_by(key, 37)
    assert np.array_equal(idx, np.argsort(key, kind="stable"))
    assert np.array_equal(ptr, np.concatenate([[0], np.cumsum(np.bincount(key, minlength=37))]))
    rowptr, col, vals, cols, *_ = synth.random_sparse_system(500, 60, 6, seed=1)
    t_ptr, t_row, t_val = oracle.transpose(500, cols, rowptr, col, vals)
    order = np.argsort(col, kind="stable")
    rows_of = np.repeat(np.arange(500), np.diff(rowptr))
    assert np.array_equal(t_row, rows_of[order]) and bits_equal(t_val, vals[order])
