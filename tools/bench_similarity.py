"""Config 4: movie-movie cosine similarity with top-50 over 53 889 rank-50 item factors, 1 B200.
similarities/s = N (N-1) / device time (CUDA events, normalisation + GEMM + selection + exact
re-score); tensor-pipe roofline = 2 N^2 K / measured fp64 DMMA peak; parity: sampled query rows
against the CPU definition; CPU baseline: NumPy fp64 GEMM + argpartition on a query subsample.
Under torchrun (WORLD_SIZE > 1) every rank takes one query block
(sharded.sharded_factor_cosine_topk), the device time is the max over ranks, rank 0 prints the
whole-job line.  (The N > 1 mode was written after this round's GPU budget was spent: its host
logic is covered by tests/test_sharding_cpu.py on gloo, it has not run on GPUs yet.)"""
import argparse, json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # keep stdout to the one JSON line
from movie_recommender_b200 import similarity

ap = argparse.ArgumentParser()
ap.add_argument("--items", type=int, default=53889)
ap.add_argument("--factors", type=int, default=50)
ap.add_argument("--topk", type=int, default=50)
ap.add_argument("--q-lo", type=int, default=0)
ap.add_argument("--q-hi", type=int, default=None)
a = ap.parse_args()
rng = np.random.default_rng(20181001)
M = rng.standard_normal((a.items, a.factors))
world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
similarity.factor_cosine_topk(M[:512], topk=a.topk)     # warm-up (module load)
if world > 1:
    import torch
    import torch.distributed as dist
    from movie_recommender_b200 import sharded
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    infos = []

    def compute(q_lo, q_hi):
        i_, s_, info_ = similarity.factor_cosine_topk(M, topk=a.topk, q_lo=q_lo, q_hi=q_hi)
        infos.append(info_)
        return i_, s_
    dist.barrier()
    torch.cuda.synchronize()
    t0 = time.time()
    ids, scores = sharded.sharded_factor_cosine_topk(M, a.topk, rank, world, compute=compute)
    torch.cuda.synchronize()
    dist.barrier()
    wall = time.time() - t0
    ms = torch.tensor([infos[0].total_ms, infos[0].candidates_ms], dtype=torch.float64, device="cuda")
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)          # device time = max over ranks
    fb = torch.tensor([infos[0].fallback_rows], device="cuda")
    dist.all_reduce(fb)

    class _Info:
        total_ms, candidates_ms, fallback_rows = float(ms[0]), float(ms[1]), int(fb[0])
    info = _Info()
    if rank != 0:
        dist.destroy_process_group()
        sys.exit(0)
else:
    t0 = time.time()
    ids, scores, info = similarity.factor_cosine_topk(M, topk=a.topk, q_lo=a.q_lo, q_hi=a.q_hi)
    wall = time.time() - t0
nq = ids.shape[0]
pairs = nq * (a.items - 1)
kp = 52 if a.factors > 32 else (32 if a.factors > 16 else 16)
flops = 2.0 * nq * a.items * kp
peak = 37.09 * world
out = {"metric": "similarities_per_sec", "value": pairs / (info.total_ms * 1e-3), "unit": "pairs/s",
       "n_gpus": world,
       "config": {"workload": "C4: %d items x %d factors, top-%d, queries %d" % (a.items, a.factors, a.topk, nq)},
       "total_ms": info.total_ms, "candidates_ms": info.candidates_ms, "fallback_rows": info.fallback_rows,
       "e2e_s": wall,
       "roofline": {"bound": "tensor", "pipe": "fp64 DMMA", "achieved": flops / (info.candidates_ms * 1e-3) / 1e12,
                    "peak": peak, "unit": "TFLOP/s", "frac": flops / (info.candidates_ms * 1e-3) / 1e12 / peak}}
from oracle import oracle
sample = list(range(0, nq, max(1, nq // 64)))[:64]
ok = True
for q in sample:
    oi, os_ = oracle.cosine_topk(M, a.topk, a.q_lo + q, a.q_lo + q + 1)
    ok = ok and np.array_equal(ids[q], oi[0]) and np.array_equal(scores[q].view(np.uint64), os_[0].view(np.uint64))
out["parity"] = {"sampled_queries": len(sample), "ids_and_scores_bitexact": bool(ok)}
# CPU baseline: NumPy fp64 GEMM + argpartition on 2048 queries
H = M / np.linalg.norm(M, axis=1, keepdims=True)
m = min(2048, nq)
t0 = time.time()
S = H[a.q_lo:a.q_lo + m] @ H.T
S[np.arange(m), a.q_lo + np.arange(m)] = -9
part = np.argpartition(-S, a.topk, axis=1)[:, :a.topk]
dt = time.time() - t0
out["cpu_baseline"] = {"value": m * (a.items - 1) / dt, "unit": "pairs/s", "cores": os.cpu_count(),
                       "kind": "port", "sample": "NumPy fp64 GEMM + argpartition, %d queries, %.2f s" % (m, dt)}
print(json.dumps(out))
if world > 1:
    dist.destroy_process_group()
