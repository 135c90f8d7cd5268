"""Config 2: the linear (user + movie bias) least-squares model on ML-27M-shaped ratings through
the generic sparse solver, 1 B200.  Reports iterations, time to termination, rows*iterations/s and
the HBM roofline fraction of the CG loop (SURVEY.md 8d bytes per iteration), with the unmodified
reference library timed on the host beside it (bounded by --cpu-rows)."""
import argparse, json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from movie_recommender_b200 import cpp_ls, synth

ap = argparse.ArgumentParser()
ap.add_argument("--users", type=int, default=283228)
ap.add_argument("--items", type=int, default=53889)
ap.add_argument("--ratings", type=int, default=27753444)
ap.add_argument("--cpu-rows", type=int, default=4000000)
ap.add_argument("--faithful", action="store_true", help="also time the bit-faithful algorithm 1")
ap.add_argument("--no-warmup", action="store_true", help="skip the small warm-up solve (ncu captures)")
a = ap.parse_args()

u, i = synth.rating_pairs(a.users, a.items, a.ratings, 51, 50)
raw = synth.planted_ratings(u, i, a.users, a.items, subtract_median=False)
rowptr, col, vals, cols, b, x0 = synth.bias_model_system(u, i, raw, a.users, a.items)
rows, nnz = len(b), len(vals)
if not a.no_warmup:
    cpp_ls.cg_least_squares(rowptr[:1001], col[:2000], vals[:2000], cols, b[:1000], algorithm=3, x0=x0)  # warm-up
# the first full-size call pays cudaMalloc for ~2 GB of buffers inside the timed region (tens of ms
# on this platform); the library's arena hands the same blocks out again on the second call
first_info = None
if not a.no_warmup:
    cpp_ls.cg_least_squares(rowptr, col, vals, cols, b, algorithm=3, x0=x0)
    first_info = cpp_ls.cg_least_squares.last_info
t0 = time.time()
x, it, rr = cpp_ls.cg_least_squares(rowptr, col, vals, cols, b, algorithm=3, x0=x0)
wall = time.time() - t0
info = cpp_ls.cg_least_squares.last_info
bytes_per_it = 2 * nnz * 12 + (rows + cols + 2) * 4 + rows * 8 + nnz * 8 + 7 * cols * 8
hbm = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"] if os.path.exists("MEASURED_PEAKS.json") else 6650.0
# the solve also does one A^T b and one A^T A x before the loop: it + 1.5 operator applications
ach = bytes_per_it * (it + 1.5) / (info.solve_ms * 1e-3) / 1e9
out = {"metric": "ls_rows_iterations_per_sec", "value": rows * it / (info.solve_ms * 1e-3),
       "unit": "rows*iterations/s", "config": {"workload": "C2: bias model, %d rows x %d cols, 2 nnz/row" % (rows, cols),
       "algorithm": 3}, "iterations": it, "final_rr": rr, "solve_ms": info.solve_ms,
       "transpose_ms": info.transpose_ms, "e2e_s": wall, "ms_per_iteration": info.solve_ms / (it + 1.5),
       "first_call_solve_ms": first_info.solve_ms if first_info is not None else None,
       "roofline": {"bound": "hbm", "achieved": ach, "peak": hbm, "unit": "GB/s", "frac": ach / hbm,
                    "algorithmic_bytes_per_iteration": bytes_per_it}}
if a.faithful:
    cpp_ls.set_thread_count(os.cpu_count())
    cpp_ls.cg_least_squares(rowptr, col, vals, cols, b, 0.01, 2, algorithm=1, x0=x0)   # warm-up
    t0 = time.time()
    xf, itf, rrf = cpp_ls.cg_least_squares(rowptr, col, vals, cols, b, algorithm=1, x0=x0)
    out["faithful"] = {"iterations": itf, "final_rr": rrf, "e2e_s": time.time() - t0,
                       "thread_count": os.cpu_count(),
                       "max_abs_pred_diff_vs_native": float(np.max(np.abs((xf - x).reshape(-1))))}
from oracle import oracle
if oracle.has_ref():
    m = min(a.cpu_rows, rows)
    t0 = time.time()
    xr, itr, rrr = oracle.ref_cg_least_squares(rowptr[:m + 1], col[:2 * m], vals[:2 * m], cols, b[:m], x0,
                                               thread_count=os.cpu_count())
    dt = time.time() - t0
    out["cpu_baseline"] = {"value": m * itr / dt, "unit": "rows*iterations/s", "cores": os.cpu_count(),
                           "kind": "reference", "sample": "first %d rows, %d iterations, %.1f s" % (m, itr, dt)}
    if m == rows and a.faithful:
        out["faithful"]["bitexact_vs_reference"] = bool(
            itf == itr and np.array_equal(xf.reshape(-1).view(np.uint64), xr.reshape(-1).view(np.uint64)))
    if m == rows:
        xs, xr = x.reshape(-1), xr.reshape(-1)
        out["parity"] = {"iterations_ref": itr, "max_abs_pred_diff": float(np.max(np.abs(
            (xs[u] + xs[a.users + i]) - (xr[u] + xr[a.users + i]))))}
print(json.dumps(out))
