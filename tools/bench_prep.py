"""K7 at the full ML-27M shape: movie medians + ALS shrink on the GPU (CUDA-event kernel time and
wall time through the C ABI from host buffers), with the CPU port (oracle/prep_oracle.py, the
reference's list-of-tuples loops, one process) timed on a bounded sample beside it.
usage: python tools/bench_prep.py [factor] [cpu_sample_users]"""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from movie_recommender_b200 import prep, synth  # noqa: E402

factor = int(sys.argv[1]) if len(sys.argv) > 1 else 120
cpu_users = int(sys.argv[2]) if len(sys.argv) > 2 else 8000
c = synth.CONFIGS["C3"]
nu, ni = c["num_users"], c["num_items"]
u, i = synth.rating_pairs(nu, ni, c["num_ratings"], c["k"] + 1, c["k"])
raw = synth.planted_ratings(u, i, nu, ni, subtract_median=False)
n = len(raw)

prep.movie_medians(i, raw, ni)                     # warm-up (arena, module load)
t0 = time.time()
med, cnt, med_ms = prep.movie_medians(i, raw, ni)
med_wall = (time.time() - t0) * 1e3
prep.als_shrink(u, i, raw, nu, ni, med, factor + 1, factor)
t0 = time.time()
s = prep.als_shrink(u, i, raw, nu, ni, med, factor + 1, factor)
shr_wall = (time.time() - t0) * 1e3
s50 = prep.als_shrink(u, i, raw, nu, ni, med, c["k"] + 1, c["k"])

# algorithmic bytes: medians = per executed radix pass 8 B read + 8 B write per rating is the
# sort; the lower bound used here is ONE read of (movie id, rating) per rating.  Shrink: per
# round two streams of (user id, movie id), then flags + scan + compaction.
rounds = s.rounds
shrink_bytes = n * 8 * 2 * rounds + n * (8 + 4 + 4 + 4) + n * 8 + len(s.ratings) * 20
out = {
    "workload": "ML-27M shape, %d ratings" % n,
    "medians": {"kernel_ms": med_ms, "wall_ms_from_host_buffers": med_wall,
                "ratings_per_s": n / (med_ms * 1e-3)},
    "shrink": {"factor": factor, "rounds": rounds, "ratings_out": int(len(s.ratings)),
               "users_out": s.num_users, "movies_out": s.num_movies, "kernel_ms": s.kernel_ms,
               "wall_ms_from_host_buffers": shr_wall, "algorithmic_GBps": shrink_bytes / (s.kernel_ms * 1e-3) / 1e9},
    "shrink_k50_noop": {"rounds": s50.rounds, "kernel_ms": s50.kernel_ms},
}

# CPU port on a sample: the first cpu_users users in the reference's list form
from oracle import prep_oracle as po  # noqa: E402  (bench tools may time the oracle as the CPU baseline)
m = int(np.searchsorted(u, cpu_users))
lists = [(uu, []) for uu in range(cpu_users)]
for uu, mm, rr in zip(u[:m].tolist(), i[:m].tolist(), raw[:m].tolist()):
    lists[uu][1].append((mm, rr))
t0 = time.time()
po.medians_lists(lists)
t_med = time.time() - t0
t0 = time.time()
shrunk, _, r_cpu = po.shrink_lists(lists, factor)
users, movies = po.sorted_order(shrunk)
t_shr = time.time() - t0
out["cpu_port"] = {"kind": "port", "cores": 1, "sample": "first %d users (%d ratings), list form" % (cpu_users, m),
                   "medians_ratings_per_s": m / t_med, "shrink_ratings_per_s": m / t_shr, "shrink_rounds": r_cpu}
out["shrink"]["ratings_per_s"] = n / (s.kernel_ms * 1e-3)
print(json.dumps(out))
