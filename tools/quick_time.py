"""Ad-hoc timing of the ALS modes on a mid-size problem (development aid, not the bench)."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from movie_recommender_b200 import cpp_ls, synth

nu, ni, nnz, k = [int(x) for x in (sys.argv[1:5] if len(sys.argv) > 4 else (20000, 8000, 2000000, 50))]
algs = [int(a) for a in (sys.argv[5].split(",") if len(sys.argv) > 5 else ["1"])]
T = int(sys.argv[6]) if len(sys.argv) > 6 else 16
t = time.time()
p = synth.als_problem(nu, ni, nnz, k)
print("gen %.1fs nnz=%d" % (time.time() - t, len(p["ratings"])), flush=True)
cpp_ls.set_thread_count(T)
t = time.time()
prob = cpp_ls.AlsProblem(p["user_ids"], p["item_ids"], p["ratings"], k, nu, ni)
print("create %.2fs" % (time.time() - t), flush=True)
for alg in algs:
    prob.set_factors(p["user_factors0"], p["item_factors0"])
    for sweeps in (1, 2):
        info = prob.run(alg, -1e300, sweeps)
        print("alg %d sweeps %d: %.2f ms, cg_it %d, index %.2f ms, rr %.6g -> %.2f Mratings/s/sweep" % (
            alg, sweeps, info.device_ms, info.cg_iterations, info.index_build_ms, info.last_rr,
            len(p["ratings"]) * sweeps / info.device_ms / 1e3), flush=True)
