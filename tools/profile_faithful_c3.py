import sys, time
sys.path.insert(0, ".")
from movie_recommender_b200 import cpp_ls, synth
p = synth.als_problem(283228, 53889, 27753444, 50)
prob = cpp_ls.AlsProblem(p["user_ids"], p["item_ids"], p["ratings"], 50, 283228, 53889)
prob.set_factors(p["user_factors0"], p["item_factors0"])
cpp_ls.set_thread_count(16)
info = prob.run(1, -1e300, 1)
print("alg 1: 1 sweep %.1f ms, cg iterations %d, launches %d" % (info.device_ms, info.cg_iterations, info.kernel_launches))
