"""The generic sparse least-squares solver (cpp_ls.cg_least_squares -> C ABI -> CUDA) against
the oracle and the reference's golden vectors.  Bit-exact: the CUDA path reproduces the
reference's floating-point association order at the selected thread_count."""
import numpy as np
import pytest

from conftest import bits_equal, load_golden
from movie_recommender_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["ls_200x50", "ls_sparse"])
def test_ls_matches_reference_golden_bitexact(require_gpu, cpp_ls, name):
    g = load_golden(name)
    for T in (1, 4):
        cpp_ls.set_thread_count(T)
        for alg in (1, 2):
            x, it, rr = cpp_ls.cg_least_squares(g["rowptr"], g["colidx"], g["vals"], int(g["cols"]),
                                                g["b"], algorithm=alg, x0=g["x0"])
            assert x.shape == (int(g["cols"]), 1)
            assert it == int(g["it_T%d_a%d" % (T, alg)])
            assert bits_equal(x, g["x_T%d_a%d" % (T, alg)])
            assert bits_equal([rr], [g["rr_T%d_a%d" % (T, alg)]])


@pytest.mark.parametrize("T", [1, 3, 8, 64])
@pytest.mark.parametrize("alg", [1, 2])
def test_ls_random_sparse_bitexact_vs_oracle(require_gpu, cpp_ls, oracle, T, alg):
    rowptr, col, vals, cols, b, x0, x_real = synth.random_sparse_system(20000, 500, 9, seed=T)
    cpp_ls.set_thread_count(T)
    x, it, rr = cpp_ls.cg_least_squares(rowptr, col, vals, cols, b, algorithm=alg, x0=x0)
    xo, ito, rro = oracle.cg_least_squares(rowptr, col, vals, cols, b, x0, algorithm=alg,
                                           thread_count=T)
    assert it == ito and bits_equal(x, xo) and bits_equal([rr], [rro])
    # the reference's own acceptance criterion (cpp/ls/main.cpp:356-464): planted x recovered
    assert np.sum(np.abs(x.reshape(-1) - x_real)) <= 0.01 * cols * 5


def test_ls_termination_rules_bitexact(require_gpu, cpp_ls, oracle):
    rowptr, col, vals, cols, b, x0, _ = synth.random_sparse_system(4000, 200, 5, seed=77)
    cpp_ls.set_thread_count(4)
    for mrd, maxit in [(0.01, 200), (0.01, 3), (0.5, 200), (-1e300, 40), (0.01, 0)]:
        x, it, rr = cpp_ls.cg_least_squares(rowptr, col, vals, cols, b, mrd, maxit, x0=x0)
        xo, ito, rro = oracle.cg_least_squares(rowptr, col, vals, cols, b, x0, mrd, maxit,
                                               thread_count=4)
        assert it == ito and bits_equal(x, xo) and bits_equal([rr], [rro]), (mrd, maxit)


def test_ls_ragged_rows_and_empty_columns(require_gpu, cpp_ls, oracle):
    rng = np.random.default_rng(8)
    rows, cols = 3000, 90
    deg = rng.integers(0, 12, size=rows)
    rowptr = np.concatenate([[0], np.cumsum(deg)]).astype(np.int32)
    col = rng.integers(0, cols - 10, size=int(rowptr[-1])).astype(np.int32)  # last 10 cols empty
    vals = rng.standard_normal(len(col))
    b = rng.standard_normal(rows)
    x0 = rng.uniform(-1, 1, cols)
    cpp_ls.set_thread_count(5)
    for alg in (1, 2):
        x, it, rr = cpp_ls.cg_least_squares(rowptr, col, vals, cols, b, algorithm=alg, x0=x0)
        xo, ito, rro = oracle.cg_least_squares(rowptr, col, vals, cols, b, x0, algorithm=alg,
                                               thread_count=5)
        assert it == ito and bits_equal(x, xo) and bits_equal([rr], [rro])
        assert bits_equal(x.reshape(-1)[-10:], x0[-10:])  # untouched unknowns keep x0


def test_bias_model_bitexact_and_solves(require_gpu, cpp_ls, oracle):
    """Config 2's model at a size the oracle finishes in seconds."""
    nu, ni, nnz = 3000, 1200, 200000
    u, i = synth.rating_pairs(nu, ni, nnz, 3, 3, seed=21)
    raw = synth.planted_ratings(u, i, nu, ni, seed=21, subtract_median=False)
    rowptr, col, vals, cols, b, x0 = synth.bias_model_system(u, i, raw, nu, ni, seed=21)
    cpp_ls.set_thread_count(8)
    x, it, rr = cpp_ls.cg_least_squares(rowptr, col, vals, cols, b, x0=x0)
    xo, ito, rro = oracle.cg_least_squares(rowptr, col, vals, cols, b, x0, thread_count=8)
    assert it == ito and bits_equal(x, xo) and rr < 1e-6


def test_dimension_mismatch_is_an_error_not_a_crash(require_gpu, cpp_ls):
    rowptr, col, vals, cols, b, x0, _ = synth.random_sparse_system(100, 20, 3, seed=1)
    with pytest.raises(cpp_ls.CppLsError):
        cpp_ls.cg_least_squares(rowptr, col, vals, cols, b[:-1], x0=x0)


# ------------------------------------------------------------------------------------------------
# algorithm 3: same CG + stopping rule, GPU-native summation (K3, the config-2 fast path).
# Tolerance: relative solution error <= 1e-6 against the reference-order result (north star 1e-4).
# ------------------------------------------------------------------------------------------------
def _rel(a, b):
    a, b = np.asarray(a).reshape(-1), np.asarray(b).reshape(-1)
    return np.linalg.norm(a - b) / np.linalg.norm(b)


@pytest.mark.parametrize("rows,cols,per", [(20000, 500, 9), (5000, 300, 40), (300, 40, 3)])
def test_ls_native_matches_reference_order(require_gpu, cpp_ls, oracle, rows, cols, per):
    rowptr, col, vals, cols, b, x0, _ = synth.random_sparse_system(rows, cols, per, seed=rows)
    x, it, rr = cpp_ls.cg_least_squares(rowptr, col, vals, cols, b, algorithm=3, x0=x0)
    xo, ito, rro = oracle.cg_least_squares(rowptr, col, vals, cols, b, x0, thread_count=1)
    assert it == ito
    assert _rel(x, xo) < 1e-6 and abs(rr - rro) <= 1e-6 * max(rro, 1e-12) + 1e-12


def test_ls_native_bias_model(require_gpu, cpp_ls, oracle):
    nu, ni, nnz = 3000, 1200, 200000
    u, i = synth.rating_pairs(nu, ni, nnz, 3, 3, seed=21)
    raw = synth.planted_ratings(u, i, nu, ni, seed=21, subtract_median=False)
    rowptr, col, vals, cols, b, x0 = synth.bias_model_system(u, i, raw, nu, ni, seed=21)
    x, it, rr = cpp_ls.cg_least_squares(rowptr, col, vals, cols, b, algorithm=3, x0=x0)
    xo, ito, rro = oracle.cg_least_squares(rowptr, col, vals, cols, b, x0, thread_count=8)
    assert rr < 1e-6 and abs(it - ito) <= 1
    x = x.reshape(-1)
    pred, predo = x[u] + x[nu + i], xo[u] + xo[nu + i]
    assert np.max(np.abs(pred - predo)) < 1e-5      # both converged to the rr < 1e-6 rule
    info = cpp_ls.cg_least_squares.last_info
    assert info.iterations == it and info.solve_ms > 0


def test_ls_native_indicator_matrix_skips_values_without_changing_bits(require_gpu, cpp_ls, monkeypatch):
    """All stored values 1.0 (the bias model): the kernels that never read the value streams give
    the bits of the general kernels; one value off 1.0 switches the general kernels back on."""
    nu, ni, nnz = 2500, 900, 150000
    u, i = synth.rating_pairs(nu, ni, nnz, 3, 3, seed=4)
    raw = synth.planted_ratings(u, i, nu, ni, seed=4, subtract_median=False)
    rowptr, col, vals, cols, b, x0 = synth.bias_model_system(u, i, raw, nu, ni, seed=4)
    assert np.all(vals == 1.0)
    x, it, rr = cpp_ls.cg_least_squares(rowptr, col, vals, cols, b, algorithm=3, x0=x0)
    monkeypatch.setenv("MRB_LS_NO_UNIT", "1")
    xg, itg, rrg = cpp_ls.cg_least_squares(rowptr, col, vals, cols, b, algorithm=3, x0=x0)
    monkeypatch.delenv("MRB_LS_NO_UNIT")
    assert it == itg and rr == rrg and np.array_equal(x, xg)
    vals2 = vals.copy()
    vals2[len(vals2) // 2] = 0.5
    x2, it2, _ = cpp_ls.cg_least_squares(rowptr, col, vals2, cols, b, algorithm=3, x0=x0)
    monkeypatch.setenv("MRB_LS_NO_UNIT", "1")
    x2g, it2g, _ = cpp_ls.cg_least_squares(rowptr, col, vals2, cols, b, algorithm=3, x0=x0)
    assert it2 == it2g and np.array_equal(x2, x2g) and not np.array_equal(x2, x)


def test_ls_native_ragged_empty_and_termination(require_gpu, cpp_ls, oracle):
    rng = np.random.default_rng(8)
    rows, cols = 3000, 90
    deg = rng.integers(0, 12, size=rows)
    rowptr = np.concatenate([[0], np.cumsum(deg)]).astype(np.int32)
    col = rng.integers(0, cols - 10, size=int(rowptr[-1])).astype(np.int32)
    vals = rng.standard_normal(len(col))
    b = rng.standard_normal(rows)
    x0 = rng.uniform(-1, 1, cols)
    for mrd, maxit in [(0.01, 200), (0.01, 3), (-1e300, 25), (0.01, 0)]:
        x, it, rr = cpp_ls.cg_least_squares(rowptr, col, vals, cols, b, mrd, maxit, algorithm=3, x0=x0)
        xo, ito, rro = oracle.cg_least_squares(rowptr, col, vals, cols, b, x0, mrd, maxit, thread_count=1)
        assert it == ito and _rel(x, xo) < 1e-6, (mrd, maxit)
        assert bits_equal(x.reshape(-1)[-10:], x0[-10:])      # empty columns keep x0


def test_malformed_csr_is_refused(require_gpu, cpp_ls):
    """A column index outside [0, columns) or decreasing row pointers raise CppLsError instead of
    becoming an out-of-bounds device access; the library keeps working afterwards."""
    from movie_recommender_b200._lib import CppLsError
    rowptr = np.array([0, 2, 4], dtype=np.int32)
    col = np.array([0, 1, 1, 7], dtype=np.int32)           # 7 >= 3 columns
    vals = np.ones(4)
    b = np.ones(2)
    for alg in (1, 2, 3):
        with pytest.raises(CppLsError):
            cpp_ls.cg_least_squares(rowptr, col, vals, 3, b, algorithm=alg, x0=np.zeros(3))
    bad_ptr = np.array([0, 3, 2], dtype=np.int32)
    with pytest.raises(CppLsError):
        cpp_ls.cg_least_squares(bad_ptr, col[:2], vals[:2], 3, b, algorithm=3, x0=np.zeros(3))
    col[3] = 2
    x, it, rr = cpp_ls.cg_least_squares(rowptr, col, vals, 3, b, algorithm=1, x0=np.zeros(3))
    assert np.all(np.isfinite(x))
