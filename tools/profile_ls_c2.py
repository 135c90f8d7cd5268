"""Driver for ncu: the bias-model least-squares solve at full C2 size, one algorithm."""
import sys, time
sys.path.insert(0, ".")
import numpy as np
from movie_recommender_b200 import cpp_ls, synth
alg = int(sys.argv[1]) if len(sys.argv) > 1 else 1
u, i = synth.rating_pairs(283228, 53889, 27753444, 51, 50)
raw = synth.planted_ratings(u, i, 283228, 53889, subtract_median=False)
rowptr, col, vals, cols, b, x0 = synth.bias_model_system(u, i, raw, 283228, 53889)
cpp_ls.set_thread_count(16)
t = time.time()
x, it, rr = cpp_ls.cg_least_squares(rowptr, col, vals, cols, b, algorithm=alg, x0=x0)
print("alg %d: %d iterations, rr %.6g, %.3f s" % (alg, it, rr, time.time() - t))
