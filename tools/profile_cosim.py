"""Where K6's time goes at the a8 workload: single-query times of the heaviest movies (one CTA
each: the tail of the longest-first schedule) against the whole catalogue.
usage: python tools/profile_cosim.py [--small]"""
import sys, time
sys.path.insert(0, ".")
import numpy as np
from movie_recommender_b200 import synth
from movie_recommender_b200.build_similar_movies_db import SimilarMovieFinder

small = "--small" in sys.argv
nu, ni, nr = (283228, 53889, 27753444) if not small else (17700, 3368, 1734590)
u, i = synth.rating_pairs(nu, ni, nr, 51, 50, seed=20181001)
raw = synth.planted_ratings(u, i, nu, ni, seed=20181001, subtract_median=False)
order = np.argsort(i, kind="stable")
rng = np.random.default_rng(5)
movie_ids = np.arange(1, ni + 1)
genres = {int(m): set(int(g) for g in rng.choice(20, size=int(rng.integers(1, 4)), replace=False))
          for m in movie_ids}
f = SimilarMovieFinder.from_arrays(genres, movie_ids, i[order], u[order], raw[order])
udeg = np.bincount(u, minlength=nu).astype(np.int64)
mdeg = np.bincount(i, minlength=ni)
work = np.bincount(i, weights=udeg[u].astype(np.float64), minlength=ni)   # triples per query
print("total triples %.3e; per-SM share %.3e; heaviest query %.3e (%d raters)" % (
    work.sum(), work.sum() / 148, work.max(), mdeg[np.argmax(work)]))
f._query(0, min(256, ni), 20)
heavy = np.argsort(-work)[:4]
for a in list(heavy) + [int(np.argsort(-work)[147]), int(np.argsort(-work)[1000])]:
    f._query(int(a), int(a) + 1, 20)
    print("query %6d: %8d raters, %.3e triples, %.3f ms -> %.3f triples/cycle" % (
        a, mdeg[a], work[a], f.last_kernel_ms, work[a] / (f.last_kernel_ms * 1e-3 * 1.965e9)))
srt = np.argsort(work)
for a in (int(srt[0]), int(srt[ni // 4]), int(srt[ni // 2]), int(srt[3 * ni // 4])):
    f._query(a, a + 1, 20)
    print("light query %6d: %6d raters, %.3e triples, %.4f ms" % (a, mdeg[a], work[a], f.last_kernel_ms))
for lo, hi in ((0, ni), (0, ni // 8)):
    f._query(lo, hi, 20)
    w = work[lo:hi].sum()
    print("queries [%d, %d): %.3e triples, %.3f ms -> %.3f triples/cycle/SM" % (
        lo, hi, w, f.last_kernel_ms, w / (f.last_kernel_ms * 1e-3 * 1.965e9 * 148)))
light = np.argsort(work)[: ni // 2]
print("lighter half of the queries holds %.2f %% of the triples" % (100 * work[light].sum() / work.sum()))
