"""Phase breakdown of one als_from_python call on the full C3 workload (MRB_TIMING=1)."""
import os, sys, time
os.environ["MRB_TIMING"] = "1"
sys.path.insert(0, ".")
import numpy as np, torch
from movie_recommender_b200 import cpp_ls, synth
p = synth.als_problem(283228, 53889, 27753444, 50)
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
u, i, r, uf, itf = pin(p["user_ids"]), pin(p["item_ids"]), pin(p["ratings"]), pin(p["user_factors0"]), pin(p["item_factors0"])
for rep in range(3):
    print("---- call", rep, file=sys.stderr)
    t = time.time()
    cpp_ls.als(u, i, r, 50, 283228, 53889, -1e300, 1, 4, user_factors=cpp_ls.inplace_factors(uf), item_factors=cpp_ls.inplace_factors(itf))
    print("python wall %.1f ms" % ((time.time() - t) * 1e3), file=sys.stderr)
