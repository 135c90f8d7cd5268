"""Cold first sweep of a fresh AlsProblem against warm sweeps (full C3): where does run() go?"""
import sys, time
sys.path.insert(0, ".")
import numpy as np
from movie_recommender_b200 import cpp_ls, synth
p = synth.als_problem(283228, 53889, 27753444, 50)
for rep in range(3):
    with cpp_ls.AlsProblem(p["user_ids"], p["item_ids"], p["ratings"], 50, 283228, 53889) as prob:
        prob.set_factors(p["user_factors0"], p["item_factors0"])
        for sweep in range(3):
            t = time.time()
            info = prob.run(4, -1e300, 1)
            print("rep %d sweep %d: wall %.2f ms, gram_ms %.2f, device_ms %.2f" % (
                rep, sweep, (time.time() - t) * 1e3, info.gram_ms, info.device_ms))
