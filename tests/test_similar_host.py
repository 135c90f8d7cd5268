"""Host logic of the SimilarMovieFinder mirror (movie_recommender_b200/build_similar_movies_db.py)
on CPU: marshalling of the reference's list-of-dicts input, the buff table, find_similar_movie /
build / tune around the device calls.  The device calls themselves (`mrb_cosim_create`, `_query`,
`_pair`) are replaced by the oracle restatement here (the real one is exercised by
tests/test_gpu_cosim.py); the expectation is the golden output of the REAL reference class."""
import json
import math
import os

import numpy as np
import pytest

from conftest import GOLDEN
from oracle.similar_oracle import SimilarOracle, synthetic_catalogue


@pytest.fixture()
def Finder(monkeypatch):
    from movie_recommender_b200 import build_similar_movies_db as mod
    captured = {}

    def fake_create(n_movies, n_users, m_ptr, m_user, m_rq, u_ptr, u_movie, u_rq, mask, cnt, out):
        captured["n_movies"], captured["n_users"] = n_movies, n_users
        return 0

    def fake_query(self, q_lo, q_hi, num_results):
        o = SimilarOracle(self.movie_genres, self.movie_ratings, self.buff_limit, self.buff_point)
        # the table the device would receive must reproduce the reference's buff() bit for bit
        table = self._buff_table()
        for n in (3, 4, 7, len(table) - 1):
            if 3 <= n < len(table):
                assert table[n] == o.buff(n)
        nq = q_hi - q_lo
        idx = np.full((nq, num_results), -1, dtype=np.int32)
        score = np.zeros((nq, num_results))
        count = np.zeros(nq, dtype=np.int32)
        for q in range(nq):
            ids, scores = o.find_similar_movie(q_lo + q, num_results)
            count[q] = len(ids)
            for j, (mid, s) in enumerate(zip(ids, scores)):
                idx[q, j] = self.find_movie_index(mid)
                score[q, j] = s
        return idx, score, count

    def fake_pair(self, a, b):
        # build_similar_movies_db.py:72-107 restated: cosine over the common raters
        ra, rb = self.movie_ratings[a][1], self.movie_ratings[b][1]
        common = [u for u in ra if u in rb]
        if len(common) < 3:
            return len(common), 0.0
        r1, r2 = np.array([ra[u] for u in common]), np.array([rb[u] for u in common])
        return len(common), float(r1.dot(r2) / (np.linalg.norm(r1) * np.linalg.norm(r2)))

    monkeypatch.setattr(mod._dll, "mrb_cosim_create", fake_create)
    monkeypatch.setattr(mod.SimilarMovieFinder, "_pair", fake_pair)
    monkeypatch.setattr(mod.SimilarMovieFinder, "_query", fake_query)
    mod.SimilarMovieFinder._captured = captured
    return mod.SimilarMovieFinder


def load(name):
    with open(os.path.join(GOLDEN, "similar_%s.json" % name)) as fh:
        return json.load(fh)


def test_find_build_and_tune_against_the_reference(Finder):
    g = load("small")
    genres, ratings = synthetic_catalogue(**g["params"])
    f = Finder(genres, ratings)
    assert Finder._captured["n_movies"] == len(ratings)
    assert Finder._captured["n_users"] == len({u for _, d in ratings for u in d})
    for i_str, (ids, hexscores) in list(g["results"].items())[:40]:
        gi, gs = f.find_similar_movie(int(i_str))
        assert list(gi) == ids and [float(s).hex() for s in gs] == hexscores
    db = f.build(start=0, length=30)
    for q in range(30):
        ids = g["results"][str(q)][0]
        assert list(db.get(ratings[q][0], ())) == ids
    t = g["tune"]
    f.tune(t["movie_id1"], t["movie_id2"], 2, 20)
    assert f.buff_point == t["buff_point"] and float(f.buff_limit).hex() == t["buff_limit"]
    gi, gs = f.find_similar_movie(f.find_movie_index(t["movie_id1"]))
    assert list(gi) == t["after"][0] and [float(s).hex() for s in gs] == t["after"][1]


def test_marshalling_and_scaled_dot_product(Finder):
    ratings = [(7, {1: 4.0, 2: 3.5, 3: 1.0, 9: 5.0}), (5, {1: 2.0, 2: 5.0, 3: 4.5}), (9, {4: 1.0})]
    genres = {7: {0, 3}, 5: {3}, 9: set()}
    f = Finder(genres, ratings, buff_limit=0.1, buff_point=10)
    assert f.find_movie_index(5) == 1 and f.find_movie_index(42) == -1
    assert f._n_movies == 3 and f._max_deg == 4 and list(f._m_ptr) == [0, 4, 7, 8]
    score, n, sim = f._scaled_dot_product(0, 1)
    r1, r2 = np.array([2.0, 5.0, 4.5]), np.array([4.0, 3.5, 1.0])
    want = r1.dot(r2) / (np.linalg.norm(r1) * np.linalg.norm(r2))
    assert n == 3 and sim == want and score == want * 1.0      # n = 3: no buff yet
    assert f._scaled_dot_product(0, 2) == (0.0, 0, 0.0)
    table = f._buff_table()
    x = 3 + (3 * math.exp(0.1) - 3) * (4 - 3) / (10 - 3)
    assert table[3] == 0.0 and table[4] == math.log(x) - math.log(3)
    with pytest.raises(ValueError):
        Finder({1: [0, 0, 2]}, [(1, {10: 3.0})])                 # genres with duplicates are no set
    with pytest.raises(ValueError):
        Finder({1: {0}}, [(1, {10: 3.3})])                       # off the 0.5 grid
    with pytest.raises(ValueError):
        Finder.from_arrays({}, [1, 2], [1, 0], [5, 6], [1.0, 2.0])   # not grouped by movie
