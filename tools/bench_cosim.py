"""The reference's co-rating similarity (SimilarMovieFinder, a8) at catalogue scale on 1 B200:
ML-27M-shaped ratings (53 889 movies x 283 228 users, 27.75 M ratings on the 0.5 grid), 20 random
genres.  similarities/s = N (N-1) / kernel time (CUDA events).  The work is integer atomics:
2 x 8-byte reductions per (query, rater, co-rated movie) triple, sum_u deg(u)^2 triples in total.
The reference publishes 25.9 movies/s (~1.5 M pairs/s) on 16 vCPU for this job."""
import argparse, json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from movie_recommender_b200 import synth
from movie_recommender_b200.build_similar_movies_db import SimilarMovieFinder

ap = argparse.ArgumentParser()
ap.add_argument("--users", type=int, default=283228)
ap.add_argument("--items", type=int, default=53889)
ap.add_argument("--ratings", type=int, default=27753444)
ap.add_argument("--queries", type=int, default=None)
a = ap.parse_args()
u, i = synth.rating_pairs(a.users, a.items, a.ratings, 51, 50)
raw = synth.planted_ratings(u, i, a.users, a.items, subtract_median=False)
order = np.argsort(i, kind="stable")
rng = np.random.default_rng(5)
movie_ids = np.arange(1, a.items + 1)
genres = {int(m): set(int(g) for g in rng.choice(20, size=int(rng.integers(1, 4)), replace=False))
          for m in movie_ids}
t0 = time.time()
f = SimilarMovieFinder.from_arrays(genres, movie_ids, i[order], u[order], raw[order])
setup_s = time.time() - t0
nq = a.items if a.queries is None else a.queries
f.build(length=min(256, nq))                           # warm-up
t0 = time.time()
db = f.build(length=nq)
wall = time.time() - t0
ms = f.last_kernel_ms
deg = np.bincount(u, minlength=a.users).astype(np.float64)
triples = float((deg ** 2).sum()) * nq / a.items
out = {"metric": "similarities_per_sec", "value": nq * (a.items - 1) / (ms * 1e-3), "unit": "pairs/s",
       "config": {"workload": "a8: co-rating similarity, %d movies x %d users, %d ratings, %d queries" %
                  (a.items, a.users, len(u), nq)},
       "kernel_ms": ms, "e2e_s": wall, "setup_s": setup_s, "movies_with_results": len(db),
       "movies_per_sec": nq / (ms * 1e-3), "co_rating_triples": triples,
       "atomic_bytes_per_sec_GB": triples * 16 / (ms * 1e-3) / 1e9,
       "reference_published": {"movies_per_sec": 25.9, "hardware": "m5.4xlarge 16 vCPU (BASELINE.md)"}}
print(json.dumps(out))
