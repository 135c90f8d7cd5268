"""Host logic of the data-preparation mirror (movie_recommender_b200/movie_lens_data.py:
load_training_sets, compute_movie_medians, als_data_set_shrink_mp) on CPU: flattening of the
reference's lists, the id tables in both numberings, what stays in memory for the next factor,
the test-set bookkeeping and the files written.  The two GPU entry points are replaced by the
oracle's NumPy restatement here (the real ones are exercised by tests/test_gpu_prep.py); the
expected outputs are the golden files produced by the REAL reference functions."""
import os

import numpy as np
import pytest

from conftest import bits_equal, load_golden
from oracle import prep_oracle as po


@pytest.fixture()
def mld(monkeypatch):
    from movie_recommender_b200 import movie_lens_data as m
    from movie_recommender_b200 import prep

    def fake_medians(movie_ids, ratings, num_movie_slots=None):
        slots = int(np.max(movie_ids)) + 1 if num_movie_slots is None else num_movie_slots
        med, cnt = po.medians_coo(np.asarray(movie_ids), np.asarray(ratings), slots)
        return med, cnt, 0.0

    def fake_shrink(user_slot_ids, movie_ids, ratings, num_user_slots, num_movie_slots, medians,
                    min_user_ratings, min_movie_ratings):
        o = po.shrink_coo(np.asarray(user_slot_ids), np.asarray(movie_ids), np.asarray(ratings),
                          num_user_slots, num_movie_slots, np.asarray(medians), min_user_ratings,
                          min_movie_ratings)
        r = prep.ShrinkResult()
        r.user_ids, r.movie_ids, r.ratings, r.keep_pos = (o["user_ids"], o["movie_ids"], o["ratings"],
                                                          o["keep_pos"])
        r.user_new_id, r.movie_new_id = o["user_new_id"], o["movie_new_id"]
        r.num_users = int((o["user_new_id"] >= 0).sum())
        r.num_movies = int((o["movie_new_id"] >= 0).sum())
        r.rounds, r.kernel_ms = o["rounds"], 0.0
        return r

    monkeypatch.setattr(prep, "movie_medians", fake_medians)
    monkeypatch.setattr(prep, "als_shrink", fake_shrink)
    return m


def rebuild_lists(g):
    out = [(int(u), []) for u in g["user_raw"]]
    for p, m, r in zip(g["user_pos"], g["movie_raw"], g["ratings"]):
        out[int(p)][1].append((int(m), float(r)))
    return out


@pytest.mark.parametrize("name,id_order", [("prep_small", "sorted"), ("prep_small", "reference"),
                                           ("prep_test_set", "sorted"), ("prep_descending", "reference")])
def test_mirror_host_logic(mld, tmp_path, name, id_order):
    g = load_golden(name)
    train = rebuild_lists(g)
    with_test = bool(g["with_test"])
    test = [(u, e[:2]) for u, e in train] if with_test else None
    d = str(tmp_path) + os.sep
    mld.load_training_sets(train, test)
    user_raw, slot, movie, rating, _ = mld.training_sets_in_memory()
    assert user_raw == g["user_raw"].tolist() and np.array_equal(slot, g["user_pos"])
    assert np.array_equal(movie, g["movie_raw"]) and bits_equal(rating, g["ratings"])
    medians = mld.compute_movie_medians(directory=d, save=True)
    assert os.path.exists(d + "movie_medians_train.bin")
    assert sorted(medians) == g["median_ids"].tolist()
    assert bits_equal([medians[int(m)] for m in g["median_ids"]], g["median_values"])
    cov = mld.als_data_set_shrink_mp(medians, g["factors"].tolist(), no_test_set=not with_test,
                                     directory=d, id_order=id_order, verbose=False)
    shrunk = train
    for k in g["factors"].tolist():
        users = mld.get_als_obj("als%d_user_ids" % k, d)
        movies = mld.get_als_obj("als%d_movie_ids" % k, d)
        u, m, r = mld.get_als_obj("als%d_user_ratings_train" % k, d)
        assert (len(users), len(movies), len(r)) == cov[k]
        assert len(users) == int(g["k%d_num_users" % k]) and len(movies) == int(g["k%d_num_movies" % k])
        assert sorted(users.values()) == list(range(len(users)))
        assert sorted(movies.values()) == list(range(len(movies)))
        inv_u = np.zeros(len(users), dtype=np.int64)
        inv_u[list(users.values())] = list(users.keys())
        inv_m = np.zeros(len(movies), dtype=np.int64)
        inv_m[list(movies.values())] = list(movies.keys())
        assert np.array_equal(inv_u[u], g["k%d_users_raw" % k])
        assert np.array_equal(inv_m[m], g["k%d_movies_raw" % k])
        assert bits_equal(r, g["k%d_ratings" % k])
        shrunk, _, _ = po.shrink_lists(shrunk, k)
        if id_order == "reference":       # exactly the tables of a single-process reference run
            ref_users, ref_movies = po.reference_set_order(shrunk)
            assert list(users.items()) == list(ref_users.items())
            assert list(movies.items()) == list(ref_movies.items())
        else:                             # ascending standard ids
            assert list(movies) == sorted(movies) and list(users.values()) == list(range(len(users)))
        if with_test:
            t = mld.get_als_obj("als%d_user_ratings_test" % k, d)
            assert [e[0] for e in t] == g["k%d_test_users" % k].tolist()
            assert mld.get_als_obj("als%d_user_ratings_test_length" % k, d) == len(t)
    # what stays in memory is the shrunk, NOT median-subtracted data of the last factor
    user_raw, slot, movie, rating, _ = mld.training_sets_in_memory()
    up2, raw2, mr2, r2 = po.flatten(shrunk)
    assert user_raw == [u for u, e in shrunk if e] and np.array_equal(movie, mr2) and bits_equal(rating, r2)


def test_stale_test_files_are_removed_and_errors(mld, tmp_path):
    g = load_golden("prep_small")
    d = str(tmp_path) + os.sep
    for name in ("als3_user_ratings_test.bin", "als3_user_ratings_test_length.bin"):
        open(d + name, "wb").close()
    mld.load_training_sets(rebuild_lists(g))
    medians = mld.compute_movie_medians()
    mld.als_data_set_shrink_mp(medians, [3], no_test_set=True, directory=d, verbose=False)
    assert not os.path.exists(d + "als3_user_ratings_test.bin")
    assert not os.path.exists(d + "als3_user_ratings_test_length.bin")
    with pytest.raises(ValueError):
        mld.als_data_set_shrink_mp(medians, [3], directory=d, id_order="random", verbose=False)
    del medians[next(iter(medians))]
    mld.load_training_sets(rebuild_lists(g))
    with pytest.raises(KeyError):
        mld.als_data_set_shrink_mp(medians, [3], no_test_set=True, directory=d, verbose=False)
    with pytest.raises(ValueError):
        mld.load_training_coo([1], [0], [-5], [1.0])
    with pytest.raises(ValueError):
        mld.load_training_sets([(1, [(2, 3.0)])], user_ratings_test=[])
