"""ALS through the drop-in boundary (cpp_ls.als -> als_from_python -> CUDA) against the oracle
and the reference's golden vectors.  algorithm 1 and 2 are bit-exact at the selected
thread_count -- for every sweep, so there is no termination-flip drift."""
import numpy as np
import pytest

from conftest import bits_equal, load_golden
from movie_recommender_b200 import synth

pytestmark = pytest.mark.gpu


def test_als_matches_reference_golden_bitexact(require_gpu, cpp_ls):
    g = load_golden("als_planted")
    k = int(g["k"])
    for T in (1, 4):
        cpp_ls.set_thread_count(T)
        for alg in (1, 2):
            uf, itf, it = cpp_ls.als(g["user_ids"], g["item_ids"], g["ratings"], k,
                                     int(g["num_users"]), int(g["num_items"]), algorithm=alg,
                                     user_factors=g["uf0"], item_factors=g["if0"])
            assert it == int(g["it_T%d_a%d" % (T, alg)])
            assert bits_equal(uf, g["uf_T%d_a%d" % (T, alg)])
            assert bits_equal(itf, g["if_T%d_a%d" % (T, alg)])


@pytest.mark.parametrize("T,alg,shuffle", [(1, 1, False), (4, 1, True), (8, 2, False), (3, 2, True)])
def test_als_config1_bitexact_vs_oracle(require_gpu, cpp_ls, oracle, T, alg, shuffle):
    """Config 1 (ml-latest-small shape: 610 x 9724, 100 836 ratings, k = 10), fixed sweeps."""
    c = synth.CONFIGS["C1"]
    p = synth.als_problem(c["num_users"], c["num_items"], c["num_ratings"], c["k"],
                          min_degrees=False, shuffle=shuffle)
    sweeps = 4
    cpp_ls.set_thread_count(T)
    uf, itf, it = cpp_ls.als(p["user_ids"], p["item_ids"], p["ratings"], c["k"], c["num_users"],
                             c["num_items"], -1e300, sweeps, alg,
                             user_factors=p["user_factors0"], item_factors=p["item_factors0"])
    uo, io, ito = oracle.als(p["user_ids"], p["item_ids"], p["ratings"], c["k"],
                             p["user_factors0"], p["item_factors0"], -1e300, sweeps, alg, T)
    assert it == ito == sweeps
    assert bits_equal(uf, uo) and bits_equal(itf, io)
    rmse = oracle.rmse(p["user_ids"], p["item_ids"], p["ratings"], c["k"], uf, itf)
    assert rmse < 0.75  # the planted model is being fitted


def test_als_default_termination_bitexact(require_gpu, cpp_ls, oracle):
    p = synth.als_problem(400, 900, 40000, 7, seed=31)
    cpp_ls.set_thread_count(6)
    uf, itf, it = cpp_ls.als(p["user_ids"], p["item_ids"], p["ratings"], 7, 400, 900,
                             user_factors=p["user_factors0"], item_factors=p["item_factors0"])
    uo, io, ito = oracle.als(p["user_ids"], p["item_ids"], p["ratings"], 7, p["user_factors0"],
                             p["item_factors0"], thread_count=6)
    assert it == ito and bits_equal(uf, uo) and bits_equal(itf, io)


def test_als_rank50_bitexact(require_gpu, cpp_ls, oracle):
    """Headline rank (k = 50: two unknowns per lane in the transposed product)."""
    p = synth.als_problem(300, 500, 40000, 50, seed=41)
    cpp_ls.set_thread_count(4)
    uf, itf, it = cpp_ls.als(p["user_ids"], p["item_ids"], p["ratings"], 50, 300, 500, -1e300, 2,
                             user_factors=p["user_factors0"], item_factors=p["item_factors0"])
    uo, io, ito = oracle.als(p["user_ids"], p["item_ids"], p["ratings"], 50, p["user_factors0"],
                             p["item_factors0"], -1e300, 2, 1, 4)
    assert it == ito and bits_equal(uf, uo) and bits_equal(itf, io)


@pytest.mark.parametrize("k", [1, 31, 32, 33, 70, 128])
def test_als_rank_edges_bitexact(require_gpu, cpp_ls, oracle, k):
    p = synth.als_problem(90, 160, 9000, k, seed=k, min_degrees=False)
    cpp_ls.set_thread_count(3)
    uf, itf, it = cpp_ls.als(p["user_ids"], p["item_ids"], p["ratings"], k, 90, 160, -1e300, 2,
                             user_factors=p["user_factors0"], item_factors=p["item_factors0"])
    uo, io, ito = oracle.als(p["user_ids"], p["item_ids"], p["ratings"], k, p["user_factors0"],
                             p["item_factors0"], -1e300, 2, 1, 3)
    assert it == ito and bits_equal(uf, uo) and bits_equal(itf, io)


def test_als_users_and_items_without_ratings_keep_their_factors(require_gpu, cpp_ls, oracle):
    p = synth.als_problem(50, 60, 1500, 4, seed=3, min_degrees=False)
    nu, ni = 57, 66  # 7 users and 6 items never appear
    rng = np.random.default_rng(1)
    uf0 = rng.uniform(-1, 1, nu * 5)
    if0 = rng.uniform(-1, 1, ni * 4)
    cpp_ls.set_thread_count(2)
    uf, itf, it = cpp_ls.als(p["user_ids"], p["item_ids"], p["ratings"], 4, nu, ni, -1e300, 3,
                             user_factors=uf0, item_factors=if0)
    uo, io, ito = oracle.als(p["user_ids"], p["item_ids"], p["ratings"], 4, uf0, if0, -1e300, 3,
                             1, 2)
    assert bits_equal(uf, uo) and bits_equal(itf, io)
    assert bits_equal(uf[50 * 5:], uf0[50 * 5:]) and bits_equal(itf[60 * 4:], if0[60 * 4:])


def test_als_resume_equals_one_run(require_gpu, cpp_ls):
    """N calls with max_iteration=1 == one call with max_iteration=N."""
    p = synth.als_problem(200, 300, 12000, 6, seed=13)
    cpp_ls.set_thread_count(4)
    a = cpp_ls.als(p["user_ids"], p["item_ids"], p["ratings"], 6, 200, 300, -1e300, 3,
                   user_factors=p["user_factors0"], item_factors=p["item_factors0"])
    uf, itf = p["user_factors0"], p["item_factors0"]
    for _ in range(3):
        uf, itf, _ = cpp_ls.als(p["user_ids"], p["item_ids"], p["ratings"], 6, 200, 300, -1e300, 1,
                                user_factors=uf, item_factors=itf)
    assert bits_equal(a[0], uf) and bits_equal(a[1], itf)


def test_als_empty_and_bad_ids(require_gpu, cpp_ls):
    e = np.zeros(0, np.int32)
    uf, itf, it = cpp_ls.als(e, e, np.zeros(0), 3, 2, 2, max_iterations=2,
                             user_factors=np.ones(8), item_factors=np.ones(6))
    assert np.all(uf == 1) and np.all(itf == 1)
    with pytest.raises(cpp_ls.CppLsError):
        cpp_ls.als(np.array([0, 5], np.int32), np.array([0, 1], np.int32), np.zeros(2), 3, 2, 2,
                   max_iterations=1)
