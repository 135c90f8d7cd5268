import os, sys, time
os.environ["MRB_TIMING"] = "1"
sys.path.insert(0, ".")
import numpy as np
from movie_recommender_b200 import cpp_ls, synth
u, i = synth.rating_pairs(283228, 53889, 27753444, 51, 50)
raw = synth.planted_ratings(u, i, 283228, 53889, subtract_median=False)
rowptr, col, vals, cols, b, x0 = synth.bias_model_system(u, i, raw, 283228, 53889)
cpp_ls.set_thread_count(16)
for rep in range(3):
    print("--- call", rep, file=sys.stderr)
    t = time.time()
    x, it, rr = cpp_ls.cg_least_squares(rowptr, col, vals, cols, b, algorithm=1, x0=x0)
    print("python wall %.3f s, %d iterations" % (time.time() - t, it), file=sys.stderr)
