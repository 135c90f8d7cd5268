"""End-to-end time of the drop-in call cpp_ls.als(...) with ordinary (pageable) NumPy arrays, as the
reference's callers pass them, and with page-locked ones.  MRB_STAGE_THREADS is read once per
process: run once per setting.   usage: python tools/e2e_pageable.py [reps]"""
import os, sys, time
sys.path.insert(0, ".")
import numpy as np
from movie_recommender_b200 import cpp_ls, synth
nu, ni, nnz, k = 283228, 53889, 27753444, 50
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 4
cache = "/tmp/side_times_%d_%d_%d_%d.npz" % (nu, ni, nnz, k)
if os.path.exists(cache):
    z = np.load(cache); p = {key: z[key] for key in z.files}
else:
    p = synth.als_problem(nu, ni, nnz, k)
    np.savez(cache, **{key: v for key, v in p.items() if isinstance(v, np.ndarray)})
def run(pin):
    import torch
    conv = (lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()) if pin else np.ascontiguousarray
    u, i, r = conv(p["user_ids"]), conv(p["item_ids"]), conv(p["ratings"])
    best = 1e9
    for _ in range(reps):
        uf, itf = conv(p["user_factors0"].copy()), conv(p["item_factors0"].copy())
        t = time.time()
        cpp_ls.als(u, i, r, k, nu, ni, -1e300, 1, 4, user_factors=cpp_ls.inplace_factors(uf),
                   item_factors=cpp_ls.inplace_factors(itf))
        best = min(best, time.time() - t)
    return best * 1e3
print("MRB_STAGE_THREADS=%s: pageable %.2f ms, page-locked %.2f ms per call" % (
    os.environ.get("MRB_STAGE_THREADS", "default"), run(False), run(True)))
