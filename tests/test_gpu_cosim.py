"""K6: the reference's co-rating similarity (SimilarMovieFinder) on the GPU, through the Python
mirror of the reference class.  ids equal and scores BIT-equal to the REAL reference class
(golden results in tests/golden/similar_*.json) and to the oracle on further catalogues,
including the reliability cut, movies without genres, and tune()."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN
from oracle.similar_oracle import SimilarOracle, synthetic_catalogue

pytestmark = pytest.mark.gpu


def finder_cls():
    from movie_recommender_b200.build_similar_movies_db import SimilarMovieFinder
    return SimilarMovieFinder


def load(name):
    with open(os.path.join(GOLDEN, "similar_%s.json" % name)) as fh:
        return json.load(fh)


@pytest.mark.parametrize("name", ["small", "dense", "cut"])
def test_matches_the_real_reference_class(require_gpu, name):
    g = load(name)
    genres, ratings = synthetic_catalogue(**g["params"])
    f = finder_cls()(genres, ratings)
    for key, nres in (("results", 20), ("results_top3", 3)):
        for i_str, (ids, hexscores) in g[key].items():
            gi, gs = f.find_similar_movie(int(i_str), num_results=nres)
            assert list(gi) == ids, (key, i_str)
            assert [float(s).hex() for s in gs] == hexscores, (key, i_str)
    if "tune" in g:
        t = g["tune"]
        f.tune(t["movie_id1"], t["movie_id2"], 2, 20)
        assert f.buff_point == t["buff_point"] and float(f.buff_limit).hex() == t["buff_limit"]
        gi, gs = f.find_similar_movie(f.find_movie_index(t["movie_id1"]))
        assert list(gi) == t["after"][0] and [float(s).hex() for s in gs] == t["after"][1]
    f.close()


def test_build_equals_per_movie_queries_and_oracle(require_gpu):
    genres, ratings = synthetic_catalogue(num_movies=300, num_users=400, density=0.5, seed=11)
    f = finder_cls()(genres, ratings, buff_limit=0.2, buff_point=40)
    o = SimilarOracle(genres, ratings, buff_limit=0.2, buff_point=40)
    db = f.build()
    parts = {}
    parts.update(f.build(start=0, length=100))          # the multi-GPU split is by query range
    parts.update(f.build(start=100, length=200))
    assert parts == db
    for i, (mid, _) in enumerate(ratings):
        oi, os_ = o.find_similar_movie(i)
        assert tuple(db.get(mid, ())) == tuple(oi)
        gi, gs = f.find_similar_movie(i)
        assert [float(s).hex() for s in gs] == [float(s).hex() for s in os_]
    f.close()


def test_edge_cases(require_gpu):
    F = finder_cls()
    # nobody shares three raters: no results at all
    ratings = [(1, {10: 4.0, 11: 3.5}), (2, {10: 2.0, 12: 5.0}), (3, {13: 1.0})]
    genres = {1: {0}, 2: {0}, 3: {0}}
    f = F(genres, ratings)
    assert f.find_similar_movie(0) == ([], []) and f.build() == {}
    assert f.find_movie_index(3) == 2 and f.find_movie_index(99) == -1
    f.close()
    # off-grid ratings are refused (the kernel's exact integer accumulation needs the 0.5 grid)
    with pytest.raises(ValueError):
        F({1: {0}}, [(1, {10: 3.3})])


def test_pair_score_is_computed_on_the_device(require_gpu):
    """tune()'s pair score (build_similar_movies_db.py:72-119): common raters and cosine from
    mrb_cosim_pair, bit-equal to the NumPy expression the reference evaluates, for every pair of
    a catalogue whose rater lists are NOT sorted by user id."""
    import math
    genres, ratings = synthetic_catalogue(num_movies=40, num_users=90, density=0.35, seed=5)
    rng = np.random.default_rng(0)
    shuffled = []
    for mid, d in ratings:
        users = list(d)
        rng.shuffle(users)
        shuffled.append((mid, {u: d[u] for u in users}))
    f = finder_cls()(genres, shuffled, buff_limit=0.3, buff_point=12)
    seen_boost = False
    for a in range(len(shuffled)):
        for b in range(len(shuffled)):
            ra, rb = shuffled[a][1], shuffled[b][1]
            if len(ra) > len(rb):
                ra, rb = rb, ra
            common = [u for u in ra if u in rb]
            score, n, sim = f._scaled_dot_product(a, b)
            assert n == len(common)
            if n < 3:
                assert (score, sim) == (0.0, 0.0)
                continue
            r1 = np.array([ra[u] for u in common])
            r2 = np.array([rb[u] for u in common])
            want = r1.dot(r2) / (np.linalg.norm(r1) * np.linalg.norm(r2))
            assert float(sim).hex() == float(want).hex(), (a, b)
            x = 3 + (3 * math.exp(0.3) - 3) * (n - 3) / (12 - 3)
            buff = min(max(math.log(x) - math.log(3), 0), 0.3)
            assert float(score).hex() == float(want * (1.0 + buff)).hex()
            seen_boost |= buff > 0
    assert seen_boost
    with pytest.raises(Exception):
        f._pair(0, len(shuffled))
    f.close()


def test_catalogue_wider_than_one_shared_memory_part(require_gpu):
    """The accumulators of a query live in shared memory, 13 824 movies per part: a catalogue of
    15 000 movies is walked in two parts (movies on both sides of the cut, heavy and light
    queries), ids and scores bit-equal to the oracle."""
    genres, ratings = synthetic_catalogue(num_movies=15000, num_users=1200, density=0.01, seed=3)
    f = finder_cls()(genres, ratings, buff_limit=0.2, buff_point=30)
    o = SimilarOracle(genres, ratings, buff_limit=0.2, buff_point=30)
    deg = np.array([len(d) for _, d in ratings])
    queries = list(np.argsort(-deg)[:6]) + [0, 1, 6911, 6912, 7499, 7500, 13823, 13824, 14999]
    seen = 0
    for q in queries:
        gi, gs = f.find_similar_movie(int(q))
        oi, os_ = o.find_similar_movie(int(q))
        assert list(gi) == list(oi), q
        assert [float(s).hex() for s in gs] == [float(s).hex() for s in os_], q
        seen += len(oi)
    assert seen > 50
    f.close()
