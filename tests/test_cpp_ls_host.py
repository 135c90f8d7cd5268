"""The Python boundary (movie_recommender_b200/cpp_ls.py) on CPU.  The mirror binds the same five
symbols as the reference's cpp_ls.py, so pointing it at the UNMODIFIED reference library
(oracle/_ref/cpp_ls_lib.so) must reproduce the golden vectors the real library produced: that
checks the marshalling (dtypes, lengths, in/out buffers, c_double wrapping, the algorithm-2 symbol
name) and the order in which the initial vectors are drawn from the global NumPy RNG
(python/full_data/cpp_ls.py:92, :150-151) without a GPU."""
import numpy as np
import pytest

from conftest import bits_equal, load_golden
from oracle import oracle


@pytest.mark.parametrize("T", [1, 4])
@pytest.mark.parametrize("alg", [1, 2])
def test_als_wrapper_reproduces_the_golden_vectors(cpp_ls_on_reference, T, alg):
    cpp_ls = cpp_ls_on_reference
    g = load_golden("als_planted")
    k, nu, ni = int(g["k"]), int(g["num_users"]), int(g["num_items"])
    assert cpp_ls.has_dll_loaded()
    cpp_ls.set_thread_count(T)
    assert cpp_ls.get_thread_count() == T
    uf0, if0 = g["uf0"].copy(), g["if0"].copy()
    uf, itf, it = cpp_ls.als(g["user_ids"], g["item_ids"], g["ratings"], k, nu, ni, algorithm=alg,
                             user_factors=uf0, item_factors=if0)
    assert it == int(g["it_T%d_a%d" % (T, alg)])
    assert bits_equal(uf, g["uf_T%d_a%d" % (T, alg)]) and bits_equal(itf, g["if_T%d_a%d" % (T, alg)])
    assert bits_equal(uf0, g["uf0"]) and bits_equal(if0, g["if0"])      # the caller's arrays are copied
    # in-place variant: the very same buffers come back updated
    w_u, w_i = g["uf0"].copy(), g["if0"].copy()
    uf2, itf2, _ = cpp_ls.als(g["user_ids"].astype(np.int64), g["item_ids"].tolist(), g["ratings"], k,
                              nu, ni, algorithm=alg, user_factors=cpp_ls.inplace_factors(w_u),
                              item_factors=cpp_ls.inplace_factors(w_i))
    assert uf2 is w_u and itf2 is w_i and bits_equal(w_u, uf)
    with pytest.raises(ValueError):
        cpp_ls.als(g["user_ids"], g["item_ids"], g["ratings"], k, nu, ni, user_factors=uf0[:-1],
                   item_factors=if0)


def test_als_draws_its_initial_factors_like_the_reference(cpp_ls_on_reference):
    cpp_ls = cpp_ls_on_reference
    g = load_golden("als_planted")
    k, nu, ni = int(g["k"]), int(g["num_users"]), int(g["num_items"])
    cpp_ls.set_thread_count(4)
    np.random.seed(20181001)
    uf, itf, it = cpp_ls.als(g["user_ids"], g["item_ids"], g["ratings"], k, nu, ni)
    np.random.seed(20181001)                       # cpp_ls.py:150-151: users first, then items
    uf0 = np.random.uniform(-1, 1, nu * (k + 1))
    if0 = np.random.uniform(-1, 1, ni * k)
    uo, io, ito = oracle.ref_als(g["user_ids"], g["item_ids"], g["ratings"], k, uf0, if0,
                                 thread_count=4)
    assert it == ito and bits_equal(uf, uo) and bits_equal(itf, io)
    assert uf.shape == (nu * (k + 1),) and itf.shape == (ni * k,)


@pytest.mark.parametrize("name", ["ls_200x50", "ls_sparse"])
@pytest.mark.parametrize("alg", [1, 2])
def test_cg_least_squares_wrapper(cpp_ls_on_reference, name, alg):
    cpp_ls = cpp_ls_on_reference
    g = load_golden(name)
    cols = int(g["cols"])
    cpp_ls.set_thread_count(4)
    x, it, rr = cpp_ls.cg_least_squares(g["rowptr"], g["colidx"], g["vals"], cols, g["b"],
                                        algorithm=alg, x0=g["x0"])
    assert x.shape == (cols, 1) and x.dtype == np.float64              # cpp_ls.py:92: a column vector
    assert it == int(g["it_T4_a%d" % alg]) and rr == float(g["rr_T4_a%d" % alg])
    assert bits_equal(x, g["x_T4_a%d" % alg])
    np.random.seed(5)
    x1, it1, _ = cpp_ls.cg_least_squares(g["rowptr"], g["colidx"], g["vals"], cols, g["b"], algorithm=alg)
    np.random.seed(5)
    x0 = np.random.uniform(-1, 1, (cols, 1))
    xo, ito, _ = oracle.ref_cg_least_squares(g["rowptr"], g["colidx"], g["vals"], cols, g["b"], x0,
                                             algorithm=alg, thread_count=4)
    assert it1 == ito and bits_equal(x1, xo)
