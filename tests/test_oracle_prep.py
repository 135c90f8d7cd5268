"""Pins the data-preparation oracle (oracle/prep_oracle.py) against outputs of the REAL reference
functions run in multi-process mode (tests/golden/prep_*.npz, tests/golden/make_golden_prep.py).
CPU only."""
import numpy as np
import pytest

from conftest import bits_equal, load_golden
from oracle import prep_oracle as po

CASES = ["prep_small", "prep_test_set", "prep_descending"]


def rebuild_lists(g):
    """The [(user id, [(movie id, rating)])] list the golden inputs were flattened from."""
    out = [(int(u), []) for u in g["user_raw"]]
    for p, m, r in zip(g["user_pos"], g["movie_raw"], g["ratings"]):
        out[int(p)][1].append((int(m), float(r)))
    return out


@pytest.mark.parametrize("name", CASES)
def test_medians_match_the_reference(name):
    g = load_golden(name)
    med = po.medians_lists(rebuild_lists(g))
    assert sorted(med) == g["median_ids"].tolist()
    assert bits_equal([med[int(m)] for m in g["median_ids"]], g["median_values"])
    slots = int(g["movie_raw"].max()) + 1
    med_coo, cnt = po.medians_coo(g["movie_raw"], g["ratings"], slots)
    assert bits_equal(med_coo[g["median_ids"]], g["median_values"])
    assert np.isnan(med_coo[cnt == 0]).all() and (cnt > 0).sum() == len(g["median_ids"])


@pytest.mark.parametrize("name", CASES)
def test_shrink_lists_match_the_reference(name):
    g = load_golden(name)
    train = rebuild_lists(g)
    med = po.medians_lists(train)
    test = [(u, e[:2]) for u, e in train] if bool(g["with_test"]) else None
    for k in g["factors"].tolist():      # the reference keeps shrinking the SAME in-memory data
        train, test, _ = po.shrink_lists(train, k, test)
        users, movies = po.sorted_order(train)
        u, m, r = po.convert_lists(train, med, users, movies)
        inv_u = np.array(sorted(users), dtype=np.int64)
        inv_m = np.array(sorted(movies), dtype=np.int64)
        assert np.array_equal(inv_u[u], g["k%d_users_raw" % k])
        assert np.array_equal(inv_m[m], g["k%d_movies_raw" % k])
        assert bits_equal(r, g["k%d_ratings" % k])
        assert len(users) == int(g["k%d_num_users" % k]) and len(movies) == int(g["k%d_num_movies" % k])
        if test is not None:
            assert [e[0] for e in test] == g["k%d_test_users" % k].tolist()


@pytest.mark.parametrize("name", CASES)
def test_shrink_coo_matches_the_reference(name):
    g = load_golden(name)
    slots_u, slots_m = len(g["user_raw"]), int(g["movie_raw"].max()) + 1
    med, _ = po.medians_coo(g["movie_raw"], g["ratings"], slots_m)
    up, mr, r = g["user_pos"], g["movie_raw"], g["ratings"]
    raw_u = g["user_raw"]
    for k in g["factors"].tolist():
        s = po.shrink_coo(up, mr, r, slots_u, slots_m, med, k + 1, k)
        keep = s["keep_pos"]
        assert np.array_equal(raw_u[up[keep]], g["k%d_users_raw" % k])
        assert np.array_equal(mr[keep], g["k%d_movies_raw" % k])
        assert bits_equal(s["ratings"], g["k%d_ratings" % k])
        assert s["user_ids"].max() + 1 == int(g["k%d_num_users" % k])
        assert s["movie_ids"].max() + 1 == int(g["k%d_num_movies" % k])
        # ascending relabelling: label order == raw id order
        assert np.array_equal(np.argsort(s["user_ids"], kind="stable"),
                              np.argsort(raw_u[up[keep]], kind="stable"))
        # the reference shrinks the already shrunk data for the next factor
        up, mr, r = up[keep], mr[keep], r[keep]


def test_rounds_agree_between_the_two_forms():
    train = po.synthetic_user_ratings(400, 300, 10, seed=5)
    up, raw_u, mr, r = po.flatten(train)
    med, _ = po.medians_coo(mr, r, int(mr.max()) + 1)
    for k in (2, 6):
        shrunk, _, rounds = po.shrink_lists(train, k)
        s = po.shrink_coo(up, mr, r, len(raw_u), int(mr.max()) + 1, med, k + 1, k)
        assert rounds == s["rounds"] and rounds >= 2
        assert sum(len(e) for _, e in shrunk) == len(s["keep_pos"])


def test_single_process_set_order_is_a_relabelling():
    g = load_golden("prep_small")
    train, _, _ = po.shrink_lists(rebuild_lists(g), 3)
    users, movies = po.reference_set_order(train)
    assert sorted(users.values()) == list(range(len(users)))
    assert len(movies) == int(g["k3_num_movies"])


def test_the_two_forms_agree_on_random_inputs():
    """Property test (hypothesis): list form == COO form on arbitrary small rating sets,
    including users without ratings, repeated movies inside a user and rules that empty the set."""
    from hypothesis import given, settings, strategies as st

    entry = st.tuples(st.integers(0, 11), st.sampled_from([0.5, 1.0, 2.5, 3.0, 4.5, 5.0]))
    users = st.lists(st.lists(entry, max_size=9), min_size=1, max_size=14)

    @settings(max_examples=150, deadline=None)
    @given(users, st.integers(1, 4))
    def check(user_lists, k):
        train = [(10 + 3 * u, e) for u, e in enumerate(user_lists)]
        up, raw_u, mr, r = po.flatten(train)
        slots_m = 12
        med, cnt = po.medians_coo(mr, r, slots_m) if len(r) else (np.full(slots_m, np.nan), np.zeros(slots_m, np.int32))
        med_l = po.medians_lists(train)
        assert sorted(med_l) == np.nonzero(cnt)[0].tolist()
        assert bits_equal([med_l[m] for m in sorted(med_l)], med[cnt > 0])
        shrunk, _, rounds = po.shrink_lists(train, k)
        s = po.shrink_coo(up, mr, r, len(raw_u), slots_m, med, k + 1, k)
        assert rounds == s["rounds"]
        users_l, movies_l = po.sorted_order(shrunk)
        u, m, rr = po.convert_lists(shrunk, med_l, users_l, movies_l)
        assert np.array_equal(u, s["user_ids"]) and np.array_equal(m, s["movie_ids"])
        assert bits_equal(rr, s["ratings"])

    check()
