#!/usr/bin/env python
"""bench.py -- the headline measurement: ALS ratings/s per sweep on ML-27M-shaped synthetic
ratings (283 228 users x 53 889 movies, 27 753 444 ratings, rank 50), BASELINE.json's metric.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--algorithm A]

One "step" = one ALS sweep (user half-sweep + item half-sweep) over the whole workload.

* ``value``       : device-resident throughput.  The COO ratings, both groupings and the factors
                    are in HBM before the timed region; K sweeps are timed with CUDA events on the
                    library's stream (the lib records them around its sweep loop) and
                    cross-checked by a host clock bracketed by stream synchronisation.
* ``e2e``         : the same metric through the reference-facing call ``cpp_ls.als`` (ctypes ->
                    ``als_from_python`` of the C ABI) with HOST buffers in pinned memory, one call
                    per step (max_iterations=1: the reference's own resume mechanism, SURVEY.md
                    section 5), so every step pays the host->device copy of ratings + factors, the
                    index build, the sweep and the device->host copy of the factors.
* ``roofline``    : the gather-Gram + Cholesky kernel (k_gram), algorithmic bytes of SURVEY.md
                    section 8d (B_u + B_i per sweep) / CUDA-event time of the two launches per sweep,
                    against the measured HBM copy bandwidth (MEASURED_PEAKS.json).  The kernel is
                    fp64-tensor bound, not HBM bound, so ``roofline_fp64`` gives the same launches
                    against the measured DMMA peak (tools/fp64_peak.cu).
* ``cpu_baseline``: the UNMODIFIED reference library (oracle/_ref/cpp_ls_lib.so) on the host
                    cores, one sweep over a bounded user-subsample of the same workload.
* ``--impl reference`` times only that reference arm (rank 0) and prints its own line.

Synthetic data, seeded (movie_recommender_b200/synth.py); inputs are far larger than L2 (ratings
444 MB, user factors 116 MB, item factors 22 MB vs 126 MB of L2), so no explicit L2 flush.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# NCCL writes its version banner to stdout at N > 1; stdout carries exactly one JSON line
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")

METRIC = "als_ratings_per_sec_per_sweep"
UNIT = "ratings/s"
WORKLOAD = dict(name="C3: ALS rank 50, ML-27M shape", num_users=283228, num_items=53889,
                num_ratings=27753444, k=50)
FP64_DMMA_PEAK_TFLOPS = 37.09   # measured on this pool's B200 with tools/fp64_peak.cu (profiles/)
# dram__bytes_read.sum + dram__bytes_write.sum of the two k_gram launches of one sweep at C3 from
# the ncu --set full capture profiles/ncu_k_gram_C3_r01_v4.txt (user side 0.67 GB, movie side
# 4.17 GB): far BELOW the algorithmic gather bytes because the factor rows are served by L2.
NCU_DRAM_BYTES_PER_LAUNCH_C3 = (0.558801e9 + 0.115597e9 + 4.056275e9 + 0.114527e9) / 2.0


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.rows, self.proc, self.gpu = [], None, gpu_index
        self.t_begin = self.t_end = None

    def mark_begin(self):
        self.t_begin = time.time()

    def mark_end(self):
        self.t_end = time.time()

    def needs_continuation(self):
        """True when the timed region was too short for nvidia-smi to sample it (its period is
        tens of ms): the caller then keeps running the identical step untimed for ~0.6 s."""
        if self.proc is None or self.t_begin is None or self.t_end is None:
            return False
        inside = [r for r in self.rows if self.t_begin <= r[0] <= self.t_end + 0.05]
        return len(inside) < 3

    def mark_continuation(self, t0, t1):
        self.c_begin, self.c_end = t0, t1

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "20"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = self.rows
        window = "all"
        if self.t_begin is not None and self.t_end is not None:
            # a sample is printed up to one period after it was taken
            inside = [r for r in rows if self.t_begin <= r[0] <= self.t_end + 0.05]
            cont = [r for r in rows if getattr(self, "c_begin", None) is not None
                    and self.c_begin + 0.1 <= r[0] <= self.c_end + 0.05]
            if len(inside) >= 3:
                rows, window = inside, "timed region"
            elif cont:
                rows, window = inside + cont, "timed region + identical untimed continuation"
            elif inside:
                rows, window = inside, "timed region"
        for _, r in rows:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1]))
                smax.append(float(r[2]))
                power.append(float(r[3]))
            except ValueError:
                continue
            for name, val in zip(names, r[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(power) if power else None,
                "samples": len(sm), "window": window, "reasons": sorted(reasons)}


def algorithmic_bytes_per_sweep(w):
    """B_u + B_i, gather counted once per rating (no-cache model)."""
    nnz, k, nu, ni = w["num_ratings"], w["k"], w["num_users"], w["num_items"]
    b_u = nnz * (4 + 8 + 8 * k) + nu * (k + 1) * 8 * 2 + (nu + 1) * 4
    b_i = nnz * (4 + 8 + 8 * (k + 1)) + ni * k * 8 * 2 + (ni + 1) * 4
    return b_u + b_i


def executed_flops_per_sweep(w):
    """fp64 FLOPs the k_gram launches execute per sweep: 28 lower 8x8 tiles x 8x8x4 FMAs per 4
    ratings and side (augmented order 52 / 51 padded to 56), plus the tensor-core Cholesky."""
    nnz, nu, ni = w["num_ratings"], w["num_users"], w["num_items"]
    m8 = (w["k"] + 2 + 7) // 8
    tiles = m8 * (m8 + 1) // 2
    gram = 2 * (nnz / 4.0) * tiles * 256 * 2
    chol_dmma = sum((m8 - t - 1) * (m8 - t) // 2 * 2 for t in range(m8)) * 256 * 2
    return gram + (nu + ni) * chol_dmma


def make_problem(w, seed):
    from movie_recommender_b200 import synth
    t0 = time.time()
    p = synth.als_problem(w["num_users"], w["num_items"], w["num_ratings"], w["k"], seed=seed)
    return p, time.time() - t0


def pinned_copy(a):
    """The same array in page-locked host memory (so H2D/D2H run at full PCIe speed)."""
    import torch
    t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    return t.numpy(), t


def user_subsample(p, w, target_ratings):
    """First users of the workload (all their ratings) up to ~target_ratings; item ids kept."""
    u = p["user_ids"]
    cut = int(np.searchsorted(np.cumsum(np.bincount(u, minlength=w["num_users"])), target_ratings)) + 1
    cut = min(cut, w["num_users"])
    m = int(np.searchsorted(u, cut))       # user_ids are sorted (grouped by user)
    k = w["k"]
    return dict(user_ids=u[:m].copy(), item_ids=p["item_ids"][:m].copy(),
                ratings=p["ratings"][:m].copy(), num_users=cut, num_items=w["num_items"], k=k,
                user_factors0=p["user_factors0"][:cut * (k + 1)].copy(),
                item_factors0=p["item_factors0"].copy())


def time_reference(sample, steps, warmup, threads, budget_s=150.0):
    """The unmodified reference library, one sweep per step (als_from_python, max_iteration=1,
    factors carried over -- bit-identical to one multi-sweep call).  A sweep's
    cost follows its CG iteration count, so sweeps differ; the run stops early (and says so) if
    the wall-clock budget is exhausted."""
    from oracle import oracle
    if not oracle.has_ref():
        return None
    uf, itf = sample["user_factors0"], sample["item_factors0"]
    times = []
    t_start = time.time()
    for s in range(warmup + steps):
        t0 = time.time()
        uf, itf, _ = oracle.ref_als(sample["user_ids"], sample["item_ids"], sample["ratings"],
                                    sample["k"], uf, itf, -1e300, 1, 1, threads)
        if s >= warmup:
            times.append(time.time() - t0)
        if times and time.time() - t_start > budget_s:
            break
    rmse = oracle.rmse(sample["user_ids"], sample["item_ids"], sample["ratings"], sample["k"], uf, itf)
    return dict(sec_per_sweep=float(np.mean(times)), rmse=rmse, sweeps=warmup + len(times),
                steps_timed=len(times))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--algorithm", type=int, default=4,
                    help="4 = gathered Gram + Cholesky (north-star path, default); 1 = the "
                         "reference's CG, bit-faithful; 3 = the same CG on Gram blocks")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--cpu-sample-ratings", type=int, default=600000)
    ap.add_argument("--seed", type=int, default=20181001)
    ap.add_argument("--small", action="store_true", help="1/16-size workload (development)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    w = dict(WORKLOAD)
    if args.small:
        w.update(name="C3/16 (development)", num_users=17700, num_items=3368, num_ratings=1734590)
    host_threads = os.cpu_count() or 1

    # ------------------------------------------------------------------ reference arm
    if args.impl == "reference":
        if rank != 0:
            return 0
        p, gen_s = make_problem(w, args.seed)
        sample = user_subsample(p, w, args.cpu_sample_ratings)
        r = time_reference(sample, args.steps, max(args.warmup, 1), host_threads)
        if r is None:
            print(json.dumps({"impl": "reference", "unavailable":
                              "oracle/_ref/cpp_ls_lib.so was not built (reference sources absent)"}))
            return 0
        n_s = len(sample["ratings"])
        val = n_s / r["sec_per_sweep"]
        sample_desc = ("first %d users of the workload (%d ratings, all %d movies), 1 sweep per "
                       "step, als_from_python(max_iteration=1) carried over" %
                       (sample["num_users"], n_s, w["num_items"]))
        print(json.dumps({
            "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
            "steps": r["steps_timed"], "steps_requested": args.steps, "warmup": max(args.warmup, 1),
            "ms_per_step": r["sec_per_sweep"] * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic (seeded, ML-27M shape)",
            "config": {"workload": w["name"], "algorithm": "reference als(), algorithm=1",
                       "sample": sample_desc, "train_rmse_after": r["rmse"]},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": host_threads, "kind": "reference",
                             "sample": sample_desc},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        }))
        return 0

    # ------------------------------------------------------------------ our arm
    import torch
    from movie_recommender_b200 import _lib, cpp_ls
    if _lib.dll.mrb_device_count() < 1:
        raise SystemExit("bench.py: no CUDA device visible; there is no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    cpp_ls.set_thread_count(host_threads)

    p, gen_s = make_problem(w, args.seed)
    nnz = len(p["ratings"])
    k, nu, ni = w["k"], w["num_users"], w["num_items"]

    sampler = ClockSampler(local_rank)
    if world > 1:
        from movie_recommender_b200 import sharded
        sampler.start()                       # nvidia-smi needs ~0.3 s before its first sample
        runner = sharded.ShardedAls(p, k, nu, ni, rank, world)
    else:
        runner = None

    prob = None
    if world == 1:
        sampler.start()                                         # nvidia-smi needs ~0.3 s to start
        prob = cpp_ls.AlsProblem(p["user_ids"], p["item_ids"], p["ratings"], k, nu, ni)
        prob.set_factors(p["user_factors0"], p["item_factors0"])
        for _ in range(args.warmup):
            prob.run(args.algorithm, -1e300, 1)
        torch.cuda.synchronize()
        sampler.mark_begin()
        t0 = time.time()
        info = prob.run(args.algorithm, -1e300, args.steps)     # syncs its stream before returning
        torch.cuda.synchronize()
        wall_ms = (time.time() - t0) * 1e3
        sampler.mark_end()
        uf, itf = prob.get_factors()
        if sampler.needs_continuation():
            c0 = time.time()
            while time.time() - c0 < 0.6:
                prob.run(args.algorithm, -1e300, 2)
            sampler.mark_continuation(c0, time.time())
        clocks = sampler.stop()
        dev_ms = float(info.device_ms)
        gram_ms = float(getattr(info, "gram_ms", 0.0))
        launches = int(getattr(info, "kernel_launches", 0))
        step_ms = dev_ms / args.steps
    else:
        res = runner.bench(args.algorithm, args.warmup, args.steps, sampler)
        dev_ms, wall_ms, clocks, gram_ms, launches = (res["device_ms"], res["wall_ms"], res["clocks"],
                                                      res["gram_ms"], res["launches"])
        uf, itf = res["user_factors"], res["item_factors"]
        step_ms = dev_ms / args.steps
        runner.prob.close()
        # e2e at N GPUs: one sharded sweep per step from page-locked HOST buffers (all ranks)
        from movie_recommender_b200 import sharded as _sh
        pinned = dict(p)
        keep_alive = []
        for key in ("user_ids", "item_ids", "ratings", "user_factors0", "item_factors0"):
            pinned[key], t = pinned_copy(p[key])
            keep_alive.append(t)
        e2e_s_multi, _, _ = _sh.e2e_steps(pinned, k, nu, ni, rank, world, max(1, min(args.e2e_steps, 3)),
                                           timing=bool(os.environ.get("MRB_E2E_TIMING")))
    value = nnz / (step_ms * 1e-3)

    if rank != 0:
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
            dist.destroy_process_group()
        return 0

    # ---------------- parity / quality beside the number
    from oracle import oracle
    rmse = oracle.rmse(p["user_ids"][:2000000], p["item_ids"][:2000000], p["ratings"][:2000000], k, uf, itf)

    # ---------------- e2e through the drop-in call with pinned host buffers (N=1 path)
    e2e = None
    if world > 1:
        h2d = world * (nnz * (4 + 4 + 8) + (nu * (k + 1) + ni * k) * 8)
        d2h = world * (nu * (k + 1) + ni * k) * 8
        e2e = {"value": nnz / e2e_s_multi, "unit": UNIT, "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": d2h, "ms_per_step": e2e_s_multi * 1e3,
               "steps": max(1, min(args.e2e_steps, 3)),
               "call": "sharded.ShardedAls(host buffers) + one sweep + get_factors per step on every "
                       "rank (upload, index build, work lists, peer mapping exchange included)"}
    if world == 1:
        if prob is not None:
            prob.close()
        u_pin, _k1 = pinned_copy(p["user_ids"])
        i_pin, _k2 = pinned_copy(p["item_ids"])
        r_pin, _k3 = pinned_copy(p["ratings"])
        uf_h, _k4 = pinned_copy(p["user_factors0"])
        if_h, _k5 = pinned_copy(p["item_factors0"])
        wrap = cpp_ls.inplace_factors       # keep the factor buffers page-locked across steps
        cpp_ls.als(u_pin, i_pin, r_pin, k, nu, ni, -1e300, 1, args.algorithm,
                   user_factors=wrap(uf_h), item_factors=wrap(if_h))          # warm-up call
        t0 = time.time()
        for _ in range(args.e2e_steps):
            cpp_ls.als(u_pin, i_pin, r_pin, k, nu, ni, -1e300, 1, args.algorithm,
                       user_factors=wrap(uf_h), item_factors=wrap(if_h))
        e2e_s = (time.time() - t0) / args.e2e_steps
        h2d = nnz * (4 + 4 + 8) + (nu * (k + 1) + ni * k) * 8
        d2h = (nu * (k + 1) + ni * k) * 8
        e2e = {"value": nnz / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": d2h, "ms_per_step": e2e_s * 1e3, "steps": args.e2e_steps,
               "call": "cpp_ls.als(..., max_iterations=1, algorithm=%d) per step, pinned host "
                       "buffers, includes index build" % args.algorithm}

    # ---------------- roofline of the dominant kernel (k_gram, two launches per sweep)
    hbm_peak, peak_src = load_peaks()
    alg_bytes = algorithmic_bytes_per_sweep(w)
    roofline = roofline_fp64 = None
    if args.algorithm == 4 and gram_ms > 0:
        per_launch_ms = gram_ms / (2.0 * args.steps) * (1 if world == 1 else 1)
        achieved = (alg_bytes / 2.0 / world) / (per_launch_ms * 1e-3) / 1e9
        roofline = {"bound": "hbm", "kernel": "k_gram<7,USER|ITEM,SOLVE>", "achieved": achieved,
                    "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                    "traffic": (NCU_DRAM_BYTES_PER_LAUNCH_C3 / world if not args.small else None),
                    "traffic_source": "profiles/ncu_k_gram_C3_r01_v4.txt (1 GPU capture)",
                    "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": alg_bytes / 2.0 / world,
                    "avg_launch_ms": per_launch_ms,
                    "note": "gather bytes counted once per rating (SURVEY 8d); the kernel is "
                            "fp64-tensor bound, see roofline_fp64"}
        tf = executed_flops_per_sweep(w) / 2.0 / world / (per_launch_ms * 1e-3) / 1e12
        roofline_fp64 = {"bound": "tensor", "pipe": "fp64 DMMA (mma.sync m8n8k4)", "achieved": tf,
                         "peak": FP64_DMMA_PEAK_TFLOPS, "unit": "TFLOP/s",
                         "frac": tf / FP64_DMMA_PEAK_TFLOPS,
                         "peak_source": "tools/fp64_peak.cu on this pool (profiles/fp64_peak_r01.txt)"}

    # ---------------- CPU baseline: the reference library on a bounded sample (N=1 only)
    cpu_baseline = None
    if world == 1:
        sample = user_subsample(p, w, args.cpu_sample_ratings)
        r = time_reference(sample, 1, 1, host_threads)
        if r is not None:
            cpu_baseline = {"value": len(sample["ratings"]) / r["sec_per_sweep"], "unit": UNIT,
                            "cores": host_threads, "kind": "reference",
                            "sample": "first %d users of the workload (%d ratings), 1 warm-up + 1 "
                                      "timed sweep of the unmodified reference als()" %
                                      (sample["num_users"], len(sample["ratings"]))}
        else:
            cpu_baseline = {"value": None, "unit": UNIT, "cores": host_threads, "kind": "reference",
                            "sample": "oracle/_ref/cpp_ls_lib.so not present"}

    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": step_ms, "higher_is_better": True, "scaling": "strong"
        if world > 1 else "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic (seeded %d, ML-27M shape: %d users x %d movies, %d ratings)" %
                (args.seed, nu, ni, nnz),
        "config": {"workload": w["name"], "algorithm": args.algorithm,
                   "algorithm_name": {1: "reference CG, bit-faithful", 2: "reference CG (transpose "
                                      "variant), bit-faithful", 3: "reference CG on Gram blocks",
                                      4: "gathered Gram + Cholesky (exact half-sweeps)"}.get(
                                          args.algorithm, "?"),
                   "rank": k, "sweeps_timed": args.steps,
                   "l2": "inputs (444 MB ratings + 138 MB factors) larger than the 126 MB L2; no flush",
                   "parallelism": "users then movies row-partitioned over %d GPU(s)" % world,
                   "host_wall_ms_per_step": wall_ms / args.steps,
                   "train_rmse_after_%d_sweeps_first_2M_ratings" % (args.warmup + args.steps): rmse,
                   "data_generation_s": gen_s},
        "clocks": clocks, "gpu_launches": launches,
    }
    if e2e is not None:
        out["e2e"] = e2e
    if roofline is not None:
        out["roofline"] = roofline
        out["roofline_fp64"] = roofline_fp64
    if cpu_baseline is not None:
        out["cpu_baseline"] = cpu_baseline
    print(json.dumps(out))
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
