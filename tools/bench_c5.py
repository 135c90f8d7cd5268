"""Config 5 of BASELINE.json: ALS rank 128 on a power-law synthetic, 10 M users x 500 k movies,
1e9 ratings (x --scale), users / movies sharded over the GPUs of one box with the factor exchange
fused into the solve kernel (NVLink peer stores).  ``bench.py --config C5 --gpus N [--scale S]``.

The reference cannot run this configuration at all: it materialises ``nnz * (k+1)`` matrix values
and computes that product in ``int`` (cpp/ls_lib/matrix.cpp:757-759), which overflows at 1.3e11.

Data.  Every rank generates ONLY its own 1/N of the users (a contiguous user range, seeded by the
rank): power-law user activity (exponent 0.6), Zipf-like movie popularity (exponent 0.9),
ratings from the planted rank-8 model of synth.py + user bias + N(0, 0.4) on the 0.5 grid minus
3.5.  Unlike the C1-C3 generator the (user, movie) pairs are drawn with replacement (a de-duplicating
sort of 1e9 keys per box is not what this line measures) and no minimum degrees are enforced:
10 M users x 129 unknowns exceed 1e9 ratings, so most user systems are under-determined and the
pivot-skipping Cholesky keeps the undetermined unknowns at their previous value (the reference's
warm-started CG would do the same).  The slices travel through the sliced-upload path of
sharded.ShardedAls: each rank uploads its own ratings, pushes them to the peers over NVLink.
"""
import os
import time

import numpy as np

FULL = dict(num_users=10_000_000, num_items=500_000, num_ratings=1_000_000_000, k=128)


def _weights(n, exponent, head, seed):
    rng = np.random.default_rng(seed)
    w = (np.arange(1, n + 1, dtype=np.float64) + head) ** (-exponent)
    rng.shuffle(w)
    return w / w.sum()


def generate_slice(w, rank, world, seed):
    """This rank's users [u_lo, u_hi) and their ratings, grouped by user."""
    nu, ni, nnz = w["num_users"], w["num_items"], w["num_ratings"]
    u_lo, u_hi = nu * rank // world, nu * (rank + 1) // world
    wu = _weights(nu, 0.6, 20.0, seed)[u_lo:u_hi]                 # the same table on every rank
    deg = np.maximum(1, np.round(wu * nnz)).astype(np.int64)
    users = np.repeat(np.arange(u_lo, u_hi, dtype=np.int32), deg)
    m = len(users)
    rng = np.random.default_rng(seed + 1000 + rank)
    cw = np.cumsum(_weights(ni, 0.9, 30.0, seed + 1))
    items = np.minimum(np.searchsorted(cw, rng.random(m)), ni - 1).astype(np.int32)
    mrng = np.random.default_rng(seed + 2)                        # the planted model, same on every rank
    rank_p = 8
    pv = mrng.standard_normal((ni, rank_p))
    item_off = mrng.standard_normal(ni) * 0.5
    urng = np.random.default_rng(seed + 3000 + rank)
    pu = urng.standard_normal((u_hi - u_lo, rank_p)) * (0.9 / np.sqrt(rank_p))
    bias = urng.standard_normal(u_hi - u_lo) * 0.4
    raw = np.empty(m, dtype=np.float64)
    step = 1 << 22
    for s in range(0, m, step):
        uu = users[s:s + step] - u_lo
        ii = items[s:s + step]
        raw[s:s + step] = 3.4 + bias[uu] + item_off[ii] + np.einsum("ij,ij->i", pu[uu], pv[ii])
    raw += rng.standard_normal(m) * 0.4
    ratings = np.clip(np.round(raw * 2.0) / 2.0, 0.5, 5.0) - 3.5
    return users, items, ratings, (u_lo, u_hi)


def bench(args, sampler, rank, world):
    import torch
    import torch.distributed as dist
    from movie_recommender_b200 import cpp_ls, sharded
    w = dict(FULL)
    scale = float(args.scale)
    if args.small:
        scale = min(scale, 0.002)
    if scale != 1.0:
        w.update(num_users=int(FULL["num_users"] * scale), num_items=max(256, int(FULL["num_items"] * scale)),
                 num_ratings=int(FULL["num_ratings"] * scale))
    k, nu, ni = w["k"], w["num_users"], w["num_items"]
    t0 = time.time()
    users, items, ratings, (u_lo, u_hi) = generate_slice(w, rank, world, args.seed)
    gen_s = time.time() - t0
    device = torch.device("cuda", torch.cuda.current_device())
    # slice lengths -> slice offsets (every rank needs the same total)
    lens = torch.zeros(world, dtype=torch.int64, device=device)
    lens[rank] = len(ratings)
    if world > 1:
        dist.all_reduce(lens)
    lens = [int(v) for v in lens.cpu()]
    begin, total = sum(lens[:rank]), sum(lens)
    if total >= 2 ** 31 - 1:
        raise SystemExit("C5: %d ratings exceed the int32 index range of the API" % total)
    # initial factors: only this rank's I/O rows are ever read on the host
    i_rows = sharded.io_slice(ni, rank, world)
    frng = np.random.default_rng(args.seed + 5000 + rank)
    uf_rows = frng.uniform(-1, 1, (u_hi - u_lo) * (k + 1))
    itf_rows = frng.uniform(-1, 1, (i_rows[1] - i_rows[0]) * k)
    sampler.start()
    t0 = time.time()
    problem = dict(user_ids=users, item_ids=items, ratings=ratings, slice_begin=begin, num_ratings_total=total,
                   user_factors_rows=uf_rows, item_factors_rows=itf_rows, user_rows=(u_lo, u_hi), item_rows=i_rows)
    runner = sharded.ShardedAls(problem, k, nu, ni, rank, world, sliced_arrays=True)
    torch.cuda.synchronize()
    setup_s = time.time() - t0
    for _ in range(args.warmup):
        runner.sweep()
    runner.prob.collect_gram_ms()
    l0 = cpp_ls.kernel_launches()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler.mark_begin()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        runner.sweep()
    e1.record()
    torch.cuda.synchronize()
    sampler.mark_end()
    gram = runner.prob.collect_gram_ms()
    ms = torch.tensor([e0.elapsed_time(e1), gram, setup_s * 1e3, gen_s * 1e3], dtype=torch.float64, device=device)
    sse = torch.tensor([runner.prob.shard_sse(torch.cuda.current_stream().cuda_stream)], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(sse)
    launches = (cpp_ls.kernel_launches() - l0) * world
    runner.download_own_rows_into(uf_rows, itf_rows)
    timed_out = cpp_ls._dll.mrb_peer_barrier_timed_out()
    runner.close()
    # ---- end to end from the host slices, with the device arena warm (the first construction of
    # a process pays cudaMalloc for ~80 GB of buffers): construction [slice upload, NVLink push,
    # index build, work lists] + one sweep + download of this rank's factor rows
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.time()
    runner = sharded.ShardedAls(problem, k, nu, ni, rank, world, sliced_arrays=True)
    runner.sweep()
    runner.download_own_rows_into(uf_rows, itf_rows)
    e2e_t = torch.tensor([time.time() - t0], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    runner.close()
    cpp_ls._dll.mrb_trim_memory()

    # ---- parity on a down-scaled replica: the same generator at 1/2000 of
    # the shape, the N-GPU factors against the one-GPU factors (bits) and against NumPy lstsq
    parity = None
    small = dict(FULL)
    s2 = 0.0005
    small.update(num_users=int(FULL["num_users"] * s2), num_items=max(256, int(FULL["num_items"] * s2)),
                 num_ratings=int(FULL["num_ratings"] * s2))
    parts = [generate_slice(small, r, world, args.seed) for r in range(world)]
    su = np.concatenate([p[0] for p in parts])
    si = np.concatenate([p[1] for p in parts])
    sr = np.concatenate([p[2] for p in parts])
    snu, sni = small["num_users"], small["num_items"]
    prng = np.random.default_rng(args.seed + 9)
    suf0, sitf0 = prng.uniform(-1, 1, snu * (k + 1)), prng.uniform(-1, 1, sni * k)
    sp = dict(user_ids=su, item_ids=si, ratings=sr, user_factors0=suf0.copy(), item_factors0=sitf0.copy())
    sh = sharded.ShardedAls(sp, k, snu, sni, rank, world)
    for _ in range(2):
        sh.sweep()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    uf_n, itf_n = sh.prob.get_factors()
    sh.close()
    if rank == 0:
        with cpp_ls.AlsProblem(su, si, sr, k, snu, sni) as one:
            one.set_factors(suf0, sitf0)
            info = one.run(4, -1e300, 2)
            uf_1, itf_1 = one.get_factors()
        # one movie half-sweep of NumPy lstsq from the one-GPU user factors, on a few movies with
        # enough ratings to be well determined
        U = uf_1.reshape(snu, k + 1)
        cnt = np.bincount(si, minlength=sni)
        worst = 0.0
        checked = 0
        for mv in np.argsort(-cnt)[:8]:
            rows = np.flatnonzero(si == mv)
            if len(rows) < 2 * k:
                continue
            A = U[su[rows], :k]
            b = sr[rows] - U[su[rows], k]
            x = np.linalg.lstsq(A, b, rcond=None)[0]
            worst = max(worst, float(np.max(np.abs(A @ x - A @ itf_1.reshape(sni, k)[mv]))))
            checked += 1
        parity = {"replica": "%d users x %d movies, %d ratings, k = %d, 2 sweeps" % (snu, sni, len(sr), k),
                  "n_gpu_bits_equal_one_gpu": bool(np.array_equal(uf_n.view(np.uint64), uf_1.view(np.uint64)) and
                                                   np.array_equal(itf_n.view(np.uint64), itf_1.view(np.uint64))),
                  "train_rmse_after_2_sweeps": float(np.sqrt(info.last_rr / len(sr))),
                  "movies_checked_against_lstsq": checked,
                  "max_abs_prediction_diff_vs_lstsq": worst}
    if rank != 0:
        return None, 0
    dev_ms, gram_ms = float(ms[0]), float(ms[1])
    step_ms = dev_ms / args.steps
    from bench import fp64_peak_live, load_peaks
    peak, peak_src = fp64_peak_live()
    n_u, n_i = k + 1, k
    flops = total * ((n_u) * (n_u + 1) + 2 * n_u) + total * (n_i * (n_i + 1) + 2 * n_i) + \
        nu * n_u ** 3 / 3.0 + ni * n_i ** 3 / 3.0
    tf = flops / (gram_ms / args.steps * 1e-3) / 1e12
    hbm, hbm_src = load_peaks()
    alg_bytes = total * (4 + 8 + 8 * k) + total * (4 + 8 + 8 * (k + 1)) + (nu * n_u + ni * n_i) * 16
    e2e_s = float(e2e_t[0])
    out = {"metric": "als_ratings_per_sec_per_sweep", "value": total / (step_ms * 1e-3), "unit": "ratings/s",
           "ms_per_step": step_ms, "dtype": "f64", "scaling": "strong",
           "data": "synthetic (seeded %d): power-law users x Zipf movies, pairs drawn with replacement, "
                   "no minimum degrees (see tools/bench_c5.py)" % args.seed,
           "config": {"workload": "C5: ALS rank %d, %d users x %d movies, %d ratings%s" %
                                  (k, nu, ni, total, "" if scale == 1.0 else " (scale %g of the 1e9 shape)" % scale),
                      "algorithm": 4, "kernel": "k_gram_wide<17> (fused: 4-warp Gram bands, shared-memory "
                                                "tensor-core Cholesky)",
                      "parallelism": "rows dealt over %d GPU(s); solved rows stored into every replica over "
                                     "NVLink; device-side barrier" % world,
                      "train_rmse": float(np.sqrt(float(sse[0]) / total)), "data_generation_s": float(ms[3]) * 1e-3,
                      "setup_s": float(ms[2]) * 1e-3, "peer_barrier_timed_out": bool(timed_out),
                      "l2": "ratings (%.1f GB) and factors (%.1f GB) far larger than L2; no flush" %
                            (total * 16 / 1e9, (nu * n_u + ni * n_i) * 8 / 1e9),
                      "parity": parity},
           "e2e": {"value": total / e2e_s, "unit": "ratings/s", "ms_per_step": e2e_s * 1e3,
                   "h2d_bytes_per_step": total * 16 + (nu * n_u + ni * n_i) * 8,
                   "d2h_bytes_per_step": (nu * n_u + ni * n_i) * 8,
                   "call": "sharded.ShardedAls(host slices) [upload of 1/N per rank, NVLink push, index build, work "
                           "lists] + one sweep + download of the rank's factor rows (pageable host memory)"},
           "roofline": {"bound": "tensor", "kernel": "k_gram_wide (two launches per sweep, max over ranks)",
                        "achieved": tf, "peak": peak * world, "unit": "TFLOP/s", "frac": tf / (peak * world),
                        "flops": "algorithmic (SURVEY 8d)", "traffic": None, "peak_source": peak_src,
                        "hbm_fraction_of_algorithmic_bytes": alg_bytes / (gram_ms / args.steps * 1e-3) / 1e9 / (hbm * world)},
           "cpu_baseline": {"value": None, "unit": "ratings/s", "cores": os.cpu_count(), "kind": "reference",
                            "sample": "not runnable: the reference computes nnz*(k+1) in int "
                                      "(cpp/ls_lib/matrix.cpp:757-759), which overflows for this configuration"}}
    return out, launches
