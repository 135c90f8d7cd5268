"""als_train (python/full_data/movie_lens_data.py:684-713): same files in, same files out, and
with a seeded global NumPy RNG the factors equal the oracle's bit for bit (the wrapper draws the
initial factors with the reference's own two numpy.random.uniform calls, cpp_ls.py:150-151)."""
import os
import pickle

import numpy as np
import pytest

from conftest import bits_equal
from movie_recommender_b200 import synth

pytestmark = pytest.mark.gpu


def test_als_train_round_trip(require_gpu, cpp_ls, oracle, tmp_path):
    from movie_recommender_b200 import movie_lens_data
    d = str(tmp_path) + os.sep
    problems = {}
    for k in (3, 5):
        p = synth.als_problem(150, 120, 6000, k, seed=k)
        problems[k] = p
        with open(d + "als%d_movie_ids.bin" % k, "wb") as f:
            pickle.dump({m + 1000: m for m in range(120)}, f)
        with open(d + "als%d_user_ids.bin" % k, "wb") as f:
            pickle.dump({u + 1: u for u in range(150)}, f)
        with open(d + "als%d_user_ratings_train.bin" % k, "wb") as f:
            pickle.dump([p["user_ids"], p["item_ids"], p["ratings"]], f)
    np.random.seed(1234)
    its = movie_lens_data.als_train([3, 5], thread_count=4, algorithm=1, directory=d, verbose=False)
    np.random.seed(1234)
    for k in (3, 5):
        p = problems[k]
        uf0 = np.random.uniform(-1, 1, 150 * (k + 1))     # cpp_ls.py:150
        if0 = np.random.uniform(-1, 1, 120 * k)           # cpp_ls.py:151
        uo, io, ito = oracle.als(p["user_ids"], p["item_ids"], p["ratings"], k, uf0, if0,
                                 thread_count=4)
        uf = pickle.load(open(d + "als%d_user_factors.bin" % k, "rb"))
        itf = pickle.load(open(d + "als%d_item_factors.bin" % k, "rb"))
        assert uf.dtype == np.float64 and uf.shape == (150 * (k + 1),) and itf.shape == (120 * k,)
        assert its[k] == ito and bits_equal(uf, uo) and bits_equal(itf, io)
