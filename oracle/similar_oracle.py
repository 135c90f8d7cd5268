"""oracle/similar_oracle.py -- TEST INFRASTRUCTURE ONLY.

An array-based restatement of the reference's SimilarMovieFinder
(python/full_data/build_similar_movies_db.py:21-221).  The reference intersects Python dicts pair
by pair; this restatement scores one query against all movies with dense per-user vectors, which
is a different program with the same arithmetic: the co-rating sums are exact (ratings are
multiples of 0.5), so dot / (norm * norm), the buff and the two stable sorts give bit-identical
results.  Pinned: tests/test_oracle_similar.py compares it with the REAL class imported from
/root/reference (when present) and with tests/golden/similar_*.npz produced by the real class.
"""
import math

import numpy as np


class SimilarOracle:
    def __init__(self, movie_genres, movie_ratings, buff_limit=0.05, buff_point=100):
        self.movie_genres = movie_genres
        self.movie_ratings = movie_ratings
        self.buff_limit, self.buff_point = buff_limit, buff_point
        users = sorted({u for _, d in movie_ratings for u in d})
        self.upos = {u: i for i, u in enumerate(users)}
        n, m = len(movie_ratings), len(users)
        self.R = np.zeros((n, m))          # ratings, 0 = not rated
        self.M = np.zeros((n, m))          # rated mask
        for i, (_, d) in enumerate(movie_ratings):
            for u, r in d.items():
                self.R[i, self.upos[u]] = r
                self.M[i, self.upos[u]] = 1.0

    def genres_similar(self, id1, id2):                      # :44-69
        g = self.movie_genres
        if id1 not in g or id2 not in g:
            return False
        a, b = g[id1], g[id2]
        short = min(len(a), len(b))
        if short == 0:
            return False                                     # the reference divides by zero here
        return len(a & b) / short >= 0.5

    def buff(self, n):                                       # :109-119
        x_limit = 3 * math.exp(self.buff_limit)
        x = 3 + (x_limit - 3) * (n - 3) / (self.buff_point - 3)
        b = math.log(x) - math.log(3)
        if b > self.buff_limit: b = self.buff_limit
        if b < 0: b = 0
        return b

    def find_similar_movie(self, qi, num_results=20):        # :151-180
        common = self.M * self.M[qi]                         # common raters of (qi, b) per b
        n = common.sum(axis=1).astype(np.int64)
        dot = (self.R * self.R[qi]).sum(axis=1)              # exact: multiples of 0.25
        s_q = ((self.R[qi] ** 2) * common).sum(axis=1)       # |r_q|^2 over the common raters
        s_b = ((self.R ** 2) * self.M[qi]).sum(axis=1)
        qid = self.movie_ratings[qi][0]
        cand = []
        for b in range(len(self.movie_ratings)):
            if b == qi or n[b] < 3:
                continue
            if not self.genres_similar(qid, self.movie_ratings[b][0]):
                continue
            sim = dot[b] / (math.sqrt(s_q[b]) * math.sqrt(s_b[b])) if s_q[b] > 0 and s_b[b] > 0 else float("nan")
            score = sim * (1.0 + self.buff(int(n[b])))
            if score > 0.3:
                cand.append((self.movie_ratings[b][0], score, int(n[b])))
        if len(cand) > num_results * 20:                     # :166-168 (stable)
            cand.sort(key=lambda e: e[2], reverse=True)
            cand = cand[:num_results * 20]
        cand.sort(key=lambda e: e[1], reverse=True)          # :171 (stable)
        ids = tuple(c[0] for c in cand[:num_results])
        scores = tuple(c[1] for c in cand[:num_results])
        return (ids, scores) if ids else ([], [])


def synthetic_catalogue(num_movies, num_users, density, num_genres=12, seed=0, min_common_bias=True):
    """A small MovieLens-like catalogue in the reference's own data structures:
    movie_genres {movie_id: set}, movie_ratings [(movie_id, {user_id: rating})] (shuffled)."""
    rng = np.random.default_rng(seed)
    pop = rng.random(num_movies) ** 2 * density * 3 + 0.01
    taste = rng.standard_normal((num_users, 3))
    style = rng.standard_normal((num_movies, 3))
    movie_ids = rng.permutation(np.arange(10, 10 + 3 * num_movies))[:num_movies]
    user_ids = rng.permutation(np.arange(1000, 1000 + 2 * num_users))[:num_users]
    movie_ratings = []
    for i in range(num_movies):
        who = np.flatnonzero(rng.random(num_users) < pop[i])
        r = np.clip(np.round((3.2 + taste[who] @ style[i] * 0.8 + rng.standard_normal(len(who)) * 0.5) * 2) / 2,
                    0.5, 5.0)
        movie_ratings.append((int(movie_ids[i]), {int(user_ids[u]): float(x) for u, x in zip(who, r)}))
    movie_genres = {}
    for i in range(num_movies):
        if rng.random() < 0.05:
            continue                                        # a movie without a genre entry
        g = set(int(x) for x in rng.choice(num_genres, size=int(rng.integers(1, min(5, num_genres + 1))), replace=False))
        movie_genres[int(movie_ids[i])] = g
    order = rng.permutation(num_movies)
    return movie_genres, [movie_ratings[i] for i in order]
