"""K7 -- ALS data preparation on the GPU against

* the golden outputs of the REAL reference functions run in multi-process mode
  (tests/golden/prep_*.npz; compared in raw id space, bit-exact),
* the oracle restatement (oracle/prep_oracle.py, pinned to the same golden files) on seeded
  problems the reference would take minutes for,
* size-independent properties at the full ML-27M shape (every survivor meets the minimum
  counts, a second pass changes nothing, positions kept in order), plus the oracle itself there
  (NumPy bincounts finish in seconds).
Everything is integer work and one rounded subtraction: the bar is bit-exact."""
import os
import pickle

import numpy as np
import pytest

from conftest import bits_equal, load_golden
from oracle import prep_oracle as po

pytestmark = pytest.mark.gpu

CASES = ["prep_small", "prep_test_set", "prep_descending"]


@pytest.fixture(scope="module")
def prep(require_gpu):
    from movie_recommender_b200 import prep as m
    return m


def rebuild_lists(g):
    out = [(int(u), []) for u in g["user_raw"]]
    for p, m, r in zip(g["user_pos"], g["movie_raw"], g["ratings"]):
        out[int(p)][1].append((int(m), float(r)))
    return out


@pytest.mark.parametrize("name", CASES)
def test_medians_match_the_reference(prep, name):
    g = load_golden(name)
    med, cnt, _ = prep.movie_medians(g["movie_raw"], g["ratings"])
    assert np.array_equal(np.nonzero(cnt)[0], g["median_ids"])
    assert bits_equal(med[g["median_ids"]], g["median_values"])
    assert np.isnan(med[cnt == 0]).all()
    assert np.array_equal(cnt, np.bincount(g["movie_raw"], minlength=len(cnt)))


@pytest.mark.parametrize("name", CASES)
def test_shrink_matches_the_reference(prep, name):
    g = load_golden(name)
    up, mr, r, raw_u = g["user_pos"], g["movie_raw"], g["ratings"], g["user_raw"]
    slots_m = int(mr.max()) + 1
    med, _, _ = prep.movie_medians(mr, r, slots_m)
    for k in g["factors"].tolist():          # the reference keeps shrinking the same data
        s = prep.als_shrink(up, mr, r, len(raw_u), slots_m, med, k + 1, k)
        assert np.array_equal(raw_u[up[s.keep_pos]], g["k%d_users_raw" % k])
        assert np.array_equal(mr[s.keep_pos], g["k%d_movies_raw" % k])
        assert bits_equal(s.ratings, g["k%d_ratings" % k])
        assert s.num_users == int(g["k%d_num_users" % k])
        assert s.num_movies == int(g["k%d_num_movies" % k])
        o = po.shrink_coo(up, mr, r, len(raw_u), slots_m, med, k + 1, k)
        assert s.rounds == o["rounds"]
        for key in ("user_ids", "movie_ids", "keep_pos", "user_new_id", "movie_new_id"):
            assert np.array_equal(getattr(s, key), o[key]), key
        # next factor: what the reference leaves in memory
        raw_u = raw_u[np.nonzero(s.user_new_id >= 0)[0]]
        up, mr, r = s.user_ids, mr[s.keep_pos], r[s.keep_pos]


def test_medians_of_arbitrary_doubles(prep):
    """Every radix pass runs: full-precision values of both signs, ties, odd and even counts."""
    rng = np.random.default_rng(7)
    n, slots = 300000, 5000
    movie = rng.integers(0, slots - 50, size=n).astype(np.int32)       # the last 50 slots stay empty
    r = rng.standard_normal(n) * 10.0 ** rng.integers(-3, 4, size=n)
    r[rng.integers(0, n, size=n // 10)] = 2.5                         # ties
    r[rng.integers(0, n, size=100)] = 0.0
    med, cnt, _ = prep.movie_medians(movie, r, slots)
    med_o, cnt_o = po.medians_coo(movie, r, slots)
    assert np.array_equal(cnt, cnt_o)
    assert bits_equal(med[cnt > 0], med_o[cnt_o > 0]) and np.isnan(med[cnt == 0]).all()
    # spot-check against numpy.median itself
    for m in rng.integers(0, slots - 50, size=25):
        assert bits_equal([med[m]], [np.median(r[movie == m])])


def test_shrink_against_the_oracle_on_a_long_tail(prep):
    train = po.synthetic_user_ratings(6000, 2500, 22, seed=21)
    up, raw_u, mr, r = po.flatten(train)
    slots_m = int(mr.max()) + 1
    med, _, _ = prep.movie_medians(mr, r, slots_m)
    assert bits_equal(med[np.bincount(mr, minlength=slots_m) > 0],
                      po.medians_coo(mr, r, slots_m)[0][np.bincount(mr, minlength=slots_m) > 0])
    for k in (2, 10, 30):
        s = prep.als_shrink(up, mr, r, len(raw_u), slots_m, med, k + 1, k)
        o = po.shrink_coo(up, mr, r, len(raw_u), slots_m, med, k + 1, k)
        assert s.rounds == o["rounds"] and len(s.ratings) == len(o["ratings"]) > 0
        for key in ("user_ids", "movie_ids", "keep_pos", "user_new_id", "movie_new_id"):
            assert np.array_equal(getattr(s, key), o[key]), key
        assert bits_equal(s.ratings, o["ratings"])
    assert po.shrink_coo(up, mr, r, len(raw_u), slots_m, med, 11, 10)["rounds"] >= 3


def test_edge_cases(prep):
    from movie_recommender_b200 import _lib, cpp_ls
    e_i, e_d = np.zeros(0, np.int32), np.zeros(0)
    # empty input; listed users without ratings
    med, cnt, _ = prep.movie_medians(e_i, e_d, 4)
    assert np.isnan(med).all() and not cnt.any()
    s = prep.als_shrink(e_i, e_i, e_d, 3, 4, np.full(4, np.nan), 2, 1)
    assert len(s.ratings) == 0 and s.num_users == 0 and s.num_movies == 0
    assert (s.user_new_id == -1).all() and (s.movie_new_id == -1).all() and s.rounds == 2
    # nothing to drop: one round, identity
    u = np.repeat(np.arange(4, dtype=np.int32), 3)
    m = np.tile(np.arange(3, dtype=np.int32), 4)
    r = np.arange(12, dtype=np.float64) / 2
    med, _, _ = prep.movie_medians(m, r, 3)
    s = prep.als_shrink(u, m, r, 4, 3, med, 3, 3)
    assert s.rounds == 1 and np.array_equal(s.keep_pos, np.arange(12))
    assert np.array_equal(s.user_ids, u) and np.array_equal(s.movie_ids, m)
    assert bits_equal(s.ratings, r - med[m])
    # everything collapses
    s = prep.als_shrink(u, m, r, 4, 3, med, 4, 3)
    assert len(s.ratings) == 0 and s.num_users == 0 and s.num_movies == 0
    # a cascade: dropping the light user starves a movie, which starves the next user
    u = np.array([0, 0, 1, 1, 2, 2, 2, 3, 3, 3], dtype=np.int32)
    m = np.array([0, 1, 1, 2, 2, 3, 4, 2, 3, 4], dtype=np.int32)
    r = np.linspace(0.5, 5.0, 10)
    med, _, _ = prep.movie_medians(m, r, 5)
    s = prep.als_shrink(u, m, r, 4, 5, med, 2, 2)
    o = po.shrink_coo(u, m, r, 4, 5, med, 2, 2)
    assert np.array_equal(s.keep_pos, o["keep_pos"]) and s.rounds == o["rounds"] >= 2
    # ids out of range are refused, not dereferenced
    with pytest.raises(cpp_ls.CppLsError) as e:
        prep.als_shrink(np.array([5], np.int32), np.array([0], np.int32), np.ones(1), 2, 1,
                        np.zeros(1), 1, 1)
    assert e.value.code == _lib.ERR_ARGUMENT
    with pytest.raises(cpp_ls.CppLsError):
        prep.movie_medians(np.array([-1], np.int32), np.ones(1), 3)


@pytest.mark.parametrize("name,id_order", [("prep_small", "sorted"), ("prep_test_set", "sorted"),
                                           ("prep_descending", "reference")])
def test_trainer_mirror_writes_the_reference_files(prep, tmp_path, name, id_order):
    """movie_lens_data.als_data_set_shrink_mp: same call, same files as the reference's
    (movie_lens_data.py:547-680); the files then feed als_train unchanged."""
    from movie_recommender_b200 import movie_lens_data as mld
    g = load_golden(name)
    train = rebuild_lists(g)
    with_test = bool(g["with_test"])
    test = [(u, e[:2]) for u, e in train] if with_test else None
    d = str(tmp_path) + os.sep
    mld.load_training_sets(train, test)
    medians = mld.compute_movie_medians()
    assert sorted(medians) == g["median_ids"].tolist()
    assert bits_equal([medians[int(m)] for m in g["median_ids"]], g["median_values"])
    factors = g["factors"].tolist()
    cov = mld.als_data_set_shrink_mp(medians, factors, no_test_set=not with_test, directory=d,
                                     id_order=id_order, verbose=False)
    shrunk = train
    for k in factors:
        users = mld.get_als_obj("als%d_user_ids" % k, d)
        movies = mld.get_als_obj("als%d_movie_ids" % k, d)
        u, m, r = mld.get_als_obj("als%d_user_ratings_train" % k, d)
        assert u.dtype == np.int32 and m.dtype == np.int32 and r.dtype == np.float64
        assert cov[k] == (len(users), len(movies), len(r))
        inv_u = np.zeros(len(users), dtype=np.int64)
        inv_u[list(users.values())] = list(users.keys())
        inv_m = np.zeros(len(movies), dtype=np.int64)
        inv_m[list(movies.values())] = list(movies.keys())
        assert np.array_equal(inv_u[u], g["k%d_users_raw" % k])
        assert np.array_equal(inv_m[m], g["k%d_movies_raw" % k])
        assert bits_equal(r, g["k%d_ratings" % k])
        if id_order == "reference":      # the labels of a single-process reference run
            shrunk, _, _ = po.shrink_lists(shrunk, k)
            ref_users, ref_movies = po.reference_set_order(shrunk)
            assert users == ref_users and movies == ref_movies
            assert list(users) == list(ref_users) and list(movies) == list(ref_movies)
        if with_test:
            t = mld.get_als_obj("als%d_user_ratings_test" % k, d)
            assert [e[0] for e in t] == g["k%d_test_users" % k].tolist()
            assert mld.get_als_obj("als%d_user_ratings_test_length" % k, d) == len(t)
        else:
            assert not os.path.exists(d + "als%d_user_ratings_test.bin" % k)
    # the prepared files are exactly what the trainer entry reads
    k = factors[-1]
    its = mld.als_train([k], thread_count=4, algorithm=1, directory=d, verbose=False)
    with open(d + "als%d_item_factors.bin" % k, "rb") as f:
        itf = pickle.load(f)
    assert its[k] >= 0 and itf.shape == (cov[k][1] * k,) and np.isfinite(itf).all()


def test_full_size_ml27m_shape(prep):
    """283 228 users x 53 889 movies, 27.75 M ratings: bit-exact against the NumPy oracle, and
    the size-independent properties of the result."""
    from movie_recommender_b200 import synth
    c = synth.CONFIGS["C3"]
    nu, ni = c["num_users"], c["num_items"]
    u, i = synth.rating_pairs(nu, ni, c["num_ratings"], c["k"] + 1, c["k"])
    raw = synth.planted_ratings(u, i, nu, ni, subtract_median=False)
    med, cnt, med_ms = prep.movie_medians(i, raw, ni)
    assert np.array_equal(cnt, np.bincount(i, minlength=ni))
    assert bits_equal(med, synth.movie_medians(i, raw, ni))
    # the generator enforces the k = 50 minimum degrees: nothing may be dropped
    s = prep.als_shrink(u, i, raw, nu, ni, med, c["k"] + 1, c["k"])
    assert s.rounds == 1 and len(s.ratings) == len(raw) and s.num_users == nu and s.num_movies == ni
    assert np.array_equal(s.user_ids, u) and np.array_equal(s.movie_ids, i)
    assert bits_equal(s.ratings, raw - med[i])
    # a much stricter rule: several rounds, ~half of the users go
    k = 120
    s = prep.als_shrink(u, i, raw, nu, ni, med, k + 1, k)
    o = po.shrink_coo(u, i, raw, nu, ni, med, k + 1, k)
    assert s.rounds == o["rounds"] >= 2 and 0 < len(s.ratings) < len(raw)
    for key in ("user_ids", "movie_ids", "keep_pos", "user_new_id", "movie_new_id"):
        assert np.array_equal(getattr(s, key), o[key]), key
    assert bits_equal(s.ratings, o["ratings"])
    assert np.all(np.diff(s.keep_pos) > 0)                                   # stable compaction
    assert np.bincount(s.user_ids, minlength=s.num_users).min() >= k + 1      # every survivor
    assert np.bincount(s.movie_ids, minlength=s.num_movies).min() >= k        # meets the rule
    kept_u, kept_m = s.user_new_id >= 0, s.movie_new_id >= 0
    dropped = np.ones(len(raw), dtype=bool)
    dropped[s.keep_pos] = False
    assert not np.any(kept_u[u[dropped]] & kept_m[i[dropped]])               # nothing kept was cut
    again = prep.als_shrink(s.user_ids, s.movie_ids, raw[s.keep_pos], s.num_users, s.num_movies,
                            med[np.nonzero(kept_m)[0]], k + 1, k)
    assert again.rounds == 1 and np.array_equal(again.keep_pos, np.arange(len(s.ratings)))   # idempotent
    print("K7 at C3: medians %.2f ms, shrink(k=50) %.2f ms, shrink(k=120, %d rounds) %.2f ms"
          % (med_ms, prep.als_shrink(u, i, raw, nu, ni, med, 51, 50).kernel_ms, s.rounds, s.kernel_ms))
