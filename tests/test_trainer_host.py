"""als_train (python/full_data/movie_lens_data.py:684-713) on CPU: the trainer mirror with its
cpp_ls bound to the UNMODIFIED reference library must read the reference's pickles, draw the
initial factors in the reference's order and write the flat float64 factor files the app loads
(python/app_local/recommend.py:152-172) -- host logic only; the GPU run of the same function is
tests/test_gpu_trainer.py."""
import os
import pickle

import numpy as np
import pytest

from conftest import bits_equal
from movie_recommender_b200 import synth
from oracle import oracle


def _write_inputs(d, k, nu, ni, nnz):
    p = synth.als_problem(nu, ni, nnz, k, seed=10 + k)
    with open(d + "als%d_movie_ids.bin" % k, "wb") as f:
        pickle.dump({m + 1000: m for m in range(ni)}, f)            # only len() is used (:692)
    with open(d + "als%d_user_ids.bin" % k, "wb") as f:
        pickle.dump({u + 1: u for u in range(nu)}, f)
    with open(d + "als%d_user_ratings_train.bin" % k, "wb") as f:
        pickle.dump([p["user_ids"], p["item_ids"], p["ratings"]], f)
    return p


def test_als_train_files_in_files_out(cpp_ls_on_reference, tmp_path, capsys):
    from movie_recommender_b200 import movie_lens_data
    d = str(tmp_path) + os.sep
    nu, ni = 120, 90
    problems = {k: _write_inputs(d, k, nu, ni, 4000) for k in (2, 4)}
    np.random.seed(99)
    its = movie_lens_data.als_train([2, 4], thread_count=3, algorithm=1, directory=d)
    assert cpp_ls_on_reference.get_thread_count() == 3              # movie_lens_data.py:687-688
    np.random.seed(99)
    for k in (2, 4):                                                # the factors list in order
        p = problems[k]
        uf0 = np.random.uniform(-1, 1, nu * (k + 1))                # cpp_ls.py:150
        if0 = np.random.uniform(-1, 1, ni * k)                      # cpp_ls.py:151
        uo, io, ito = oracle.ref_als(p["user_ids"], p["item_ids"], p["ratings"], k, uf0, if0,
                                     thread_count=3)
        uf = pickle.load(open(d + "als%d_user_factors.bin" % k, "rb"))
        itf = pickle.load(open(d + "als%d_item_factors.bin" % k, "rb"))
        assert isinstance(uf, np.ndarray) and uf.dtype == np.float64 and uf.ndim == 1
        assert uf.shape == (nu * (k + 1),) and itf.shape == (ni * k,)
        assert its[k] == ito and bits_equal(uf, uo) and bits_equal(itf, io)
    out = capsys.readouterr().out                                   # the reference's progress lines
    assert out.count("Building ALS factor") == 2 and out.count("ALS took") == 2


def test_als_train_keeps_the_thread_count_when_none_is_given(cpp_ls_on_reference, tmp_path):
    from movie_recommender_b200 import movie_lens_data
    d = str(tmp_path) + os.sep
    _write_inputs(d, 3, 60, 50, 1500)
    cpp_ls_on_reference.set_thread_count(5)
    movie_lens_data.als_train([3], directory=d, verbose=False)
    assert cpp_ls_on_reference.get_thread_count() == 5
    with pytest.raises(FileNotFoundError):                          # a factor that was not prepared
        movie_lens_data.als_train([7], directory=d, verbose=False)
