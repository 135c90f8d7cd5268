"""The configurations of BASELINE.json that are not the headline, behind ``bench.py --config``:

  C2    linear (user + movie bias) least squares on ML-27M-shaped ratings, generic sparse solver
  C4    movie-movie factor-cosine similarity, top-50 over 53 889 rank-50 movie factors
  a8    the reference's co-rating similarity (SimilarMovieFinder) at catalogue scale
  prep  the ALS data preparation (movie medians + shrink) at the ML-27M shape
  C5    ALS rank 128 on a power-law synthetic (10 M users x 500 k movies, 1e9 ratings x --scale)

Each prints ONE JSON line with the bench contract's keys (metric, value, unit, n_gpus, steps,
warmup, ms_per_step, e2e, roofline, cpu_baseline, clocks, gpu_launches).  ``value`` is measured
with CUDA events inside the library (inputs resident in HBM), ``e2e`` is the wall clock of the
reference-facing call with HOST buffers.  A step is one complete pass of the path (one solve, one
catalogue, one preparation).  The CPU legs execute oracle/ or oracle/_ref as the thing timed on
the host, never as part of the product path."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return float(p["hbm_gbs"]), float(p["bf16_tflops"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, 1600.0, "fallback (B200_PROFILING.md)"


def _finish(out, args, world, sampler, launches, t_begin=None):
    out.setdefault("n_gpus", world)
    out.setdefault("steps", args.steps)
    out.setdefault("warmup", args.warmup)
    out.setdefault("higher_is_better", True)
    out.setdefault("scaling", "strong")
    out.setdefault("vs_baseline", None)
    out.setdefault("data", "synthetic (seeded %d)" % args.seed)
    out["clocks"] = sampler.stop() if sampler is not None else None
    out["gpu_launches"] = int(launches)
    print(json.dumps(out))
    return 0


def _timed(fn, warmup, steps, sampler):
    """warmup untimed calls, then `steps` timed ones; returns the list of (result, wall_s)."""
    for _ in range(warmup):
        fn()
    sampler.mark_begin()
    res = []
    for _ in range(steps):
        t0 = time.time()
        r = fn()
        res.append((r, time.time() - t0))
    sampler.mark_end()
    if sampler.needs_continuation():
        c0 = time.time()
        while time.time() - c0 < 0.6:
            fn()
        sampler.mark_continuation(c0, time.time())
    return res


# ------------------------------------------------------------------------------------------ C2
def bench_c2(args, sampler):
    from movie_recommender_b200 import cpp_ls, synth
    nu, ni, nr = 283228, 53889, 27753444
    if args.small:
        nu, ni, nr = 17700, 3368, 1734590
    u, i = synth.rating_pairs(nu, ni, nr, 51, 50, seed=args.seed)
    raw = synth.planted_ratings(u, i, nu, ni, seed=args.seed, subtract_median=False)
    rowptr, col, vals, cols, b, x0 = synth.bias_model_system(u, i, raw, nu, ni, seed=args.seed)
    rows, nnz = len(b), len(vals)
    cpp_ls.set_thread_count(os.cpu_count() or 1)
    cpp_ls.cg_least_squares(rowptr[:1001], col[:2000], vals[:2000], cols, b[:1000], algorithm=3, x0=x0)
    sampler.start()
    l0 = cpp_ls.kernel_launches()

    def solve():
        x, it, rr = cpp_ls.cg_least_squares(rowptr, col, vals, cols, b, algorithm=3, x0=x0)
        return x, it, rr, cpp_ls.cg_least_squares.last_info.solve_ms
    res = _timed(solve, args.warmup, args.steps, sampler)
    launches = (cpp_ls.kernel_launches() - l0) * args.steps // (args.steps + args.warmup)
    x, it, rr, _ = res[-1][0]
    solve_ms = float(np.mean([r[0][3] for r in res]))
    wall = float(np.mean([r[1] for r in res]))
    hbm, _, src = _peaks()
    bytes_per_it = 2 * nnz * 12 + (rows + cols + 2) * 4 + rows * 8 + nnz * 8 + 7 * cols * 8
    ach = bytes_per_it * (it + 1.5) / (solve_ms * 1e-3) / 1e9   # + A^T b and A^T A x0 before the loop
    out = {"metric": "ls_rows_iterations_per_sec", "value": rows * it / (solve_ms * 1e-3),
           "unit": "rows*iterations/s", "ms_per_step": solve_ms, "dtype": "f64",
           "config": {"workload": "C2: bias model, %d rows x %d cols, 2 nnz/row" % (rows, cols),
                      "algorithm": "3 (reference CG and stopping rule, GPU-native summation order)",
                      "iterations": it, "final_rr": rr, "ms_per_iteration": solve_ms / (it + 1.5),
                      "l2": "operator streams (0.78 GB) larger than L2; no flush"},
           "e2e": {"value": rows * it / wall, "unit": "rows*iterations/s", "ms_per_step": wall * 1e3,
                   "h2d_bytes_per_step": nnz * 12 + (rows + 1) * 4 + rows * 8 + cols * 8,
                   "d2h_bytes_per_step": cols * 8,
                   "call": "cpp_ls.cg_least_squares(..., algorithm=3), pageable NumPy arrays"},
           "roofline": {"bound": "hbm", "kernel": "CG iteration (k_csr_mul_thread + k_csc_flat + fold + updates)",
                        "achieved": ach, "peak": hbm, "unit": "GB/s", "frac": ach / hbm, "traffic": None,
                        "algorithmic_bytes_per_iteration": bytes_per_it, "peak_source": src,
                        "note": "every stored value of the bias model is 1.0: found at upload, the two value "
                                "streams (2 x 8 x nnz = %d of the algorithmic bytes) are then not read -- "
                                "bit-identical results (MRB_LS_NO_UNIT=1 keeps the general kernels); history at C2: "
                                "0.555 ms per iteration, 0.512 without the value streams, 0.439 with four rows "
                                "per thread in A p (profiles/ncu_k_ls_native_C2_r02.txt)" % (16 * nnz)}}
    # parity: algorithm 1 is bit-identical to the reference (same iteration count, same bits)
    t0 = time.time()
    xf, itf, rrf = cpp_ls.cg_least_squares(rowptr, col, vals, cols, b, algorithm=1, x0=x0)
    t_f = time.time() - t0
    xs, xfl = x.reshape(-1), xf.reshape(-1)
    out["config"]["parity"] = {
        "algorithm1_iterations": itf, "algorithm1_final_rr": rrf, "algorithm1_e2e_s": t_f,
        "max_abs_prediction_diff_alg3_vs_alg1": float(np.max(np.abs(
            (xs[u] + xs[nu + i]) - (xfl[u] + xfl[nu + i])))),
        "rmse_alg3": float(np.sqrt(np.mean((xs[u] + xs[nu + i] - raw) ** 2))),
        "rmse_alg1": float(np.sqrt(np.mean((xfl[u] + xfl[nu + i] - raw) ** 2)))}
    from oracle import oracle
    if oracle.has_ref():
        m = min(4000000 if not args.small else rows, rows)
        t0 = time.time()
        xr, itr, rrr = oracle.ref_cg_least_squares(rowptr[:m + 1], col[:2 * m], vals[:2 * m], cols, b[:m],
                                                   x0, thread_count=os.cpu_count() or 1)
        dt = time.time() - t0
        out["cpu_baseline"] = {"value": m * itr / dt, "unit": "rows*iterations/s", "cores": os.cpu_count(),
                               "kind": "reference",
                               "sample": "first %d rows, %d iterations, %.1f s" % (m, itr, dt)}
        if m == rows:
            out["config"]["parity"]["alg1_bits_equal_reference"] = bool(
                itf == itr and np.array_equal(xfl.view(np.uint64), xr.reshape(-1).view(np.uint64)))
    return out, launches


# ------------------------------------------------------------------------------------------ C4
def bench_c4(args, sampler, rank, world):
    from movie_recommender_b200 import similarity, cpp_ls
    n, k, topk = (53889, 50, 50) if not args.small else (8192, 50, 50)
    rng = np.random.default_rng(args.seed)
    M = rng.standard_normal((n, k))
    similarity.factor_cosine_topk(M[:512], topk=topk)            # module load
    q_lo, q_hi = n * rank // world, n * (rank + 1) // world
    sampler.start()
    l0 = cpp_ls.kernel_launches()

    def run():
        return similarity.factor_cosine_topk(M, topk=topk, q_lo=q_lo, q_hi=q_hi)
    if world > 1:
        import torch
        import torch.distributed as dist
        dist.barrier()
        torch.cuda.synchronize()
    res = _timed(run, args.warmup, args.steps, sampler)
    launches = (cpp_ls.kernel_launches() - l0) * args.steps // (args.steps + args.warmup)
    ids, scores, info = res[-1][0]
    total_ms = float(np.mean([r[0][2].total_ms for r in res]))
    cand_ms = float(np.mean([r[0][2].candidates_ms for r in res]))
    wall = float(np.mean([r[1] for r in res]))
    fallback = int(info.fallback_rows)
    if world > 1:
        import torch
        import torch.distributed as dist
        from movie_recommender_b200 import sharded
        t = torch.tensor([total_ms, cand_ms, wall], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)            # max over ranks
        total_ms, cand_ms, wall = (float(v) for v in t)
        fb = torch.tensor([fallback], device="cuda")
        dist.all_reduce(fb)
        fallback = int(fb[0])
        bounds = sharded.query_blocks(n, world)
        ids = sharded._gather_rows(np.ascontiguousarray(ids, dtype=np.int32), bounds, rank, world, -1)
        scores = sharded._gather_rows(np.ascontiguousarray(scores, dtype=np.float64), bounds, rank, world, 0.0)
        launches *= world
        if rank != 0:
            return None, 0
    pairs = n * (n - 1)
    _, bf16_peak, src = _peaks()
    kind = similarity.kernel_kind() if hasattr(similarity, "kernel_kind") else "fp64 DMMA (mma.sync m8n8k4)"
    kp = similarity.padded_k(k) if hasattr(similarity, "padded_k") else 52
    flops = 2.0 * n * n * kp
    tf = flops / (cand_ms * 1e-3) / 1e12
    out = {"metric": "similarities_per_sec", "value": pairs / (total_ms * 1e-3), "unit": "pairs/s",
           "ms_per_step": total_ms, "dtype": "tf32 candidates + f64 exact re-score" if "tcgen05" in kind else "f64",
           "config": {"workload": "C4: %d movies x %d factors, top-%d" % (n, k, topk),
                      "candidates_ms": cand_ms, "fallback_rows": fallback, "kernel": kind,
                      "parallelism": "query blocks over %d GPU(s), catalogue replicated" % world,
                      "l2": "catalogue (%.1f MB) fits L2 by design; the N x N score matrix never exists" %
                            (n * k * 8 / 1e6)},
           "e2e": {"value": pairs / wall, "unit": "pairs/s", "ms_per_step": wall * 1e3,
                   "h2d_bytes_per_step": world * n * k * 8, "d2h_bytes_per_step": n * topk * 12,
                   "call": "similarity.factor_cosine_topk(M, topk) with a host array, results on the host"},
           "roofline": {"bound": "tensor", "kernel": "candidate GEMM + fused selection", "achieved": tf,
                        "peak": bf16_peak * world * (0.5 if "tf32" in kind else 1.0), "unit": "TFLOP/s",
                        "frac": tf / (bf16_peak * world * (0.5 if "tf32" in kind else 1.0)), "traffic": None,
                        "flops": "2 N^2 K with K padded to %d" % kp,
                        "peak_source": src + (" (tf32 = half the bf16 rate)" if "tf32" in kind else "")}}
    from oracle import oracle
    sample = list(range(0, n, max(1, n // 64)))[:64]
    ok = True
    for q in sample:
        oi, os_ = oracle.cosine_topk(M, topk, q, q + 1)
        ok = ok and np.array_equal(ids[q], oi[0]) and np.array_equal(scores[q].view(np.uint64),
                                                                      os_[0].view(np.uint64))
    out["config"]["parity"] = {"sampled_queries": len(sample), "ids_and_scores_bitexact": bool(ok),
                               "note": "parity unpinned: the reference has no factor-based similarity; "
                                       "the definition is oracle/ls_oracle.c:oracle_cosine_topk"}
    H = M / np.linalg.norm(M, axis=1, keepdims=True)
    m = min(2048, n)
    t0 = time.time()
    S = H[:m] @ H.T
    S[np.arange(m), np.arange(m)] = -9
    np.argpartition(-S, topk, axis=1)[:, :topk]
    dt = time.time() - t0
    out["cpu_baseline"] = {"value": m * (n - 1) / dt, "unit": "pairs/s", "cores": os.cpu_count(), "kind": "port",
                           "sample": "NumPy fp64 GEMM + argpartition, %d queries, %.2f s" % (m, dt)}
    return out, launches


# ------------------------------------------------------------------------------------------ a8
def bench_a8(args, sampler):
    from movie_recommender_b200 import synth, cpp_ls
    from movie_recommender_b200.build_similar_movies_db import SimilarMovieFinder
    nu, ni, nr = (283228, 53889, 27753444) if not args.small else (17700, 3368, 1734590)
    u, i = synth.rating_pairs(nu, ni, nr, 51, 50, seed=args.seed)
    raw = synth.planted_ratings(u, i, nu, ni, seed=args.seed, subtract_median=False)
    order = np.argsort(i, kind="stable")
    rng = np.random.default_rng(5)
    movie_ids = np.arange(1, ni + 1)
    genres = {int(m): set(int(g) for g in rng.choice(20, size=int(rng.integers(1, 4)), replace=False))
              for m in movie_ids}
    t0 = time.time()
    f = SimilarMovieFinder.from_arrays(genres, movie_ids, i[order], u[order], raw[order])
    setup_s = time.time() - t0
    f.build(length=min(256, ni))
    sampler.start()
    l0 = cpp_ls.kernel_launches()

    def run():
        db = f.build(length=ni)
        return db, f.last_kernel_ms
    res = _timed(run, args.warmup, args.steps, sampler)
    launches = (cpp_ls.kernel_launches() - l0) * args.steps // (args.steps + args.warmup)
    ms = float(np.mean([r[0][1] for r in res]))
    wall = float(np.mean([r[1] for r in res]))
    db = res[-1][0][0]
    deg = np.bincount(u, minlength=nu).astype(np.float64)
    triples = float((deg ** 2).sum())
    hbm, _, src = _peaks()
    # algorithmic bytes: 16 bytes of accumulator update per (query, rater, co-rated movie) triple
    # (SURVEY 8d's figure; since round 2 the accumulators live in shared memory, so this is an
    # equivalent rate, not DRAM traffic -- DRAM sees the 4-byte packed list entries only)
    ach = triples * 16 / (ms * 1e-3) / 1e9
    out = {"metric": "similarities_per_sec", "value": ni * (ni - 1) / (ms * 1e-3), "unit": "pairs/s",
           "ms_per_step": ms, "dtype": "int64 accumulation + f64 scores",
           "config": {"workload": "a8: co-rating similarity (SimilarMovieFinder), %d movies x %d users, %d ratings"
                                  % (ni, nu, len(u)), "movies_with_results": len(db),
                      "movies_per_sec": ni / (ms * 1e-3), "co_rating_triples": triples, "setup_s": setup_s,
                      "parity": "bit-exact ids and scores against the real class: tests/test_gpu_cosim.py"},
           "e2e": {"value": ni * (ni - 1) / wall, "unit": "pairs/s", "ms_per_step": wall * 1e3,
                   "h2d_bytes_per_step": 0, "d2h_bytes_per_step": ni * 20 * 12,
                   "call": "SimilarMovieFinder.build() -> {movie id: similar ids} on the host (ratings uploaded "
                           "once at construction, setup_s)"},
           "roofline": {"bound": "hbm", "kernel": "k_cosim (shared-memory integer atomics, 4 x u32 per movie, catalogue in 216 KB parts)", "achieved": ach,
                        "peak": hbm, "unit": "GB/s", "frac": ach / hbm, "traffic": None, "peak_source": src,
                        "triples_per_cycle_per_sm": triples / (ms * 1e-3) / 1.965e9 / 148,
                        "note": "16 algorithmic bytes per co-rating triple; accumulators in shared memory "
                                "(native ATOMS.ADD), so the bound is shared-memory atomic issue and the latency "
                                "of the list gathers, not HBM: a heavy query alone runs at 1.2-1.3 triples per "
                                "cycle and SM, the light two thirds of the queries are latency bound "
                                "(profiles/cosim_query_times_r02.txt)"},
           "cpu_baseline": {"value": 25.9 * (ni - 1), "unit": "pairs/s", "cores": 16, "kind": "reference",
                            "sample": "published by the reference (BASELINE.md): 25.9 movies/s on m5.4xlarge, "
                                      "not re-timed here (pure-Python O(N^2 deg) loop)"}}
    return out, launches


# ---------------------------------------------------------------------------------------- prep
def bench_prep(args, sampler):
    from movie_recommender_b200 import prep, synth, cpp_ls
    c = synth.CONFIGS["C3"]
    nu, ni, nr = (c["num_users"], c["num_items"], c["num_ratings"]) if not args.small else (17700, 3368, 1734590)
    factor = 120
    u, i = synth.rating_pairs(nu, ni, nr, c["k"] + 1, c["k"], seed=args.seed)
    raw = synth.planted_ratings(u, i, nu, ni, seed=args.seed, subtract_median=False)
    n = len(raw)
    prep.movie_medians(i, raw, ni)
    sampler.start()
    l0 = cpp_ls.kernel_launches()

    def run():
        med, cnt, med_ms = prep.movie_medians(i, raw, ni)
        s = prep.als_shrink(u, i, raw, nu, ni, med, factor + 1, factor)
        return med_ms, s
    res = _timed(run, args.warmup, args.steps, sampler)
    launches = (cpp_ls.kernel_launches() - l0) * args.steps // (args.steps + args.warmup)
    med_ms = float(np.mean([r[0][0] for r in res]))
    shr_ms = float(np.mean([r[0][1].kernel_ms for r in res]))
    wall = float(np.mean([r[1] for r in res]))
    s = res[-1][0][1]
    hbm, _, src = _peaks()
    shrink_bytes = n * 8 * 2 * s.rounds + n * (8 + 4 + 4 + 4) + n * 8 + len(s.ratings) * 20
    median_bytes = n * 12
    ach = (shrink_bytes + median_bytes) / ((med_ms + shr_ms) * 1e-3) / 1e9
    out = {"metric": "prep_ratings_per_sec", "value": n / ((med_ms + shr_ms) * 1e-3), "unit": "ratings/s",
           "ms_per_step": med_ms + shr_ms, "dtype": "int32 / f64 (bit-exact)",
           "config": {"workload": "prep: movie medians + ALS shrink (factor %d), ML-27M shape, %d ratings" % (factor, n),
                      "medians_ms": med_ms, "shrink_ms": shr_ms, "shrink_rounds": s.rounds,
                      "ratings_out": int(len(s.ratings)), "users_out": s.num_users, "movies_out": s.num_movies,
                      "parity": "bit-exact against the real reference functions: tests/test_gpu_prep.py"},
           "e2e": {"value": n / wall, "unit": "ratings/s", "ms_per_step": wall * 1e3,
                   "h2d_bytes_per_step": 2 * n * 16, "d2h_bytes_per_step": int(len(s.ratings)) * 20 + ni * 12,
                   "call": "prep.movie_medians + prep.als_shrink with pageable NumPy arrays"},
           "roofline": {"bound": "hbm", "kernel": "radix passes + k_count_alive rounds + compaction", "achieved": ach,
                        "peak": hbm, "unit": "GB/s", "frac": ach / hbm, "traffic": None, "peak_source": src,
                        "algorithmic_bytes": shrink_bytes + median_bytes}}
    from oracle import prep_oracle as po
    cpu_users = 8000 if not args.small else 2000
    m = int(np.searchsorted(u, cpu_users))
    lists = [(uu, []) for uu in range(cpu_users)]
    for uu, mm, rr in zip(u[:m].tolist(), i[:m].tolist(), raw[:m].tolist()):
        lists[uu][1].append((mm, rr))
    t0 = time.time()
    po.medians_lists(lists)
    shrunk, _, r_cpu = po.shrink_lists(lists, factor)
    po.sorted_order(shrunk)
    dt = time.time() - t0
    out["cpu_baseline"] = {"value": m / dt, "unit": "ratings/s", "cores": 1, "kind": "port",
                           "sample": "first %d users (%d ratings), the reference's list loops, one process" % (cpu_users, m)}
    return out, launches


def run(args, rank, world, local_rank, ClockSampler):
    import torch
    from movie_recommender_b200 import _lib
    if args.impl == "reference":
        if rank == 0:
            print(json.dumps({"impl": "reference", "unavailable":
                              "the reference arm is defined for --config C3 (the headline); the other "
                              "configurations report their CPU leg as cpu_baseline"}))
        return 0
    if _lib.dll.mrb_device_count() < 1:
        raise SystemExit("bench.py: no CUDA device visible; there is no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    sampler = ClockSampler(local_rank)
    if args.config == "C5":
        import bench_c5
        out, launches = bench_c5.bench(args, sampler, rank, world)
    elif args.config == "C4":
        out, launches = bench_c4(args, sampler, rank, world)
    elif world > 1:
        raise SystemExit("--config %s is a one-GPU configuration" % args.config)
    elif args.config == "C2":
        out, launches = bench_c2(args, sampler)
    elif args.config == "a8":
        out, launches = bench_a8(args, sampler)
    else:
        out, launches = bench_prep(args, sampler)
    rc = 0
    if out is not None:
        rc = _finish(out, args, world, sampler, launches)
    else:
        sampler.stop()
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()
    return rc
