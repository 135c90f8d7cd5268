"""Small driver for ncu: builds one ALS problem and runs `sweeps` sweeps of one algorithm.
usage: python tools/profile_als.py NU NI NNZ K ALG SWEEPS"""
import sys, time
sys.path.insert(0, ".")
from movie_recommender_b200 import cpp_ls, synth
nu, ni, nnz, k, alg, sweeps = [int(x) for x in sys.argv[1:7]]
t = time.time()
p = synth.als_problem(nu, ni, nnz, k)
prob = cpp_ls.AlsProblem(p["user_ids"], p["item_ids"], p["ratings"], k, nu, ni)
prob.set_factors(p["user_factors0"], p["item_factors0"])
cpp_ls.set_thread_count(16)
prob.run(alg, -1e300, 1)   # warm-up (also builds the per-side work lists)
info = prob.run(alg, -1e300, sweeps)
print("alg %d: %d sweeps in %.2f ms (%.2f Mratings/s/sweep), index build %.1f ms, setup %.1fs" % (
    alg, sweeps, info.device_ms, len(p["ratings"]) * sweeps / info.device_ms / 1e3,
    info.index_build_ms, time.time() - t))
