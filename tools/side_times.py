"""Per-side timing of the algorithm-4 half-sweeps (development aid): user / movie launch times from
the library's own CUDA events, with and without the solve epilogue (MRB_DEBUG_SKIP_SOLVE=1 times the
accumulation alone; its results are invalid, so it runs last).
usage: python tools/side_times.py [NU NI NNZ K [REPS]]"""
import os
import sys
import time

sys.path.insert(0, ".")
from movie_recommender_b200 import cpp_ls, synth

a = [int(x) for x in sys.argv[1:]]
nu, ni, nnz, k = a[:4] if len(a) >= 4 else (283228, 53889, 27753444, 50)
reps = a[4] if len(a) > 4 else 5
t = time.time()
import numpy as np
cache = "/tmp/side_times_%d_%d_%d_%d.npz" % (nu, ni, nnz, k)     # A/B runs share one generation
if os.path.exists(cache):
    z = np.load(cache)
    p = {key: z[key] for key in z.files}
else:
    p = synth.als_problem(nu, ni, nnz, k)
    np.savez(cache, **{key: v for key, v in p.items() if isinstance(v, np.ndarray)})
prob = cpp_ls.AlsProblem(p["user_ids"], p["item_ids"], p["ratings"], k, nu, ni)
prob.set_factors(p["user_factors0"], p["item_factors0"])
prob.run(4, -1e300, 2)          # warm-up, builds the work lists
prob.collect_gram_ms()


def side(user_side):
    best = 1e9
    for _ in range(reps):
        prob.half_sweep(user_side, 0)
        prob.shard_sse(0)       # synchronises stream 0
        best = min(best, prob.collect_gram_ms())
    return best


out = {}
out["user_ms"], out["item_ms"] = side(True), side(False)
info = prob.run(4, -1e300, 10)
out["sweep_ms"] = info.device_ms / 10
os.environ["MRB_DEBUG_SKIP_SOLVE"] = "1"
out["user_accumulate_only_ms"], out["item_accumulate_only_ms"] = side(True), side(False)
os.environ["MRB_DEBUG_SKIP_SOLVE"] = "4"
out["user_solve_only_ms"], out["item_solve_only_ms"] = side(True), side(False)
print(" ".join("%s=%.3f" % kv for kv in out.items()), "setup %.1fs" % (time.time() - t))
