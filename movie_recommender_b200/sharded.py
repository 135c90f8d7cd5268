"""Multi-GPU ALS: one process per GPU, ``torch.distributed`` for the
plumbing (rendezvous, the 384-byte handle exchange, max-over-ranks timing), hand-written CUDA for
everything on the data path.

Partitioning: the degree-sorted users, then movies, are dealt over the ``world`` ranks in snake
order (or cut into contiguous cost-balanced ranges for the NCCL baseline); every rank holds the
COO and full replicas of both factor matrices, but uploads only its 1/world slice (the slices are
pushed to the peers over NVLink) and groups only the rows it owns.  Exchange: the one place the
path shards is the all-gather of the freshly solved factor rows after each half-sweep.  Two
implementations:

* ``exchange="p2p"`` (default, the product): the solve kernel stores every solved row into all
  replicas through NVLink peer pointers (CUDA IPC mappings of the other ranks' buffers), i.e. the
  all-gather is fused into the producing kernel and overlaps the math row by row; between
  half-sweeps only a device-side flag barrier over the same mappings remains (``csrc/peer.cu``),
  enqueued on the stream -- no host round trip, no library collective.
* ``exchange="nccl"`` (the baseline it is compared with): every rank uploads everything, rows are
  written locally and the contiguous ranges are exchanged with ``torch.distributed.all_gather``
  on tensors aliasing the library's buffers.
"""
import numpy as np

from . import cpp_ls


class _DevArray:
    """Wraps a raw device pointer for torch.as_tensor via __cuda_array_interface__."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (ptr, False),
                                         "version": 3, "strides": None}


def io_slice(n, rank, world):
    """Contiguous 1/world share of n items that rank moves across its host link."""
    return n * rank // world, n * (rank + 1) // world


class ShardedAls:
    """One ALS problem over ``world`` GPUs (this process drives GPU ``rank``).

    ``exchange="p2p"`` (the product).  Host traffic: every rank uploads only its 1/world slice
    of the COO and of the initial factors and pushes it to the peers over NVLink (copy engines,
    CUDA IPC mappings), so every byte crosses the box's host memory system once.  Work: the
    degree-sorted rows are dealt over the ranks (``partition=1``; rows are stored individually
    into every replica, so ownership need not be contiguous).  Exchange: fused into the solve
    kernel (peer stores).  Synchronisation: a device-side flag barrier over the same mappings
    (``peer.cu``), enqueued on the stream -- no host round trip, no library collective on the
    data path.  ``exchange="nccl"`` is the baseline: full upload per rank, contiguous ranges,
    ``all_gather`` of the ranges after each half-sweep."""

    def __init__(self, problem, k, num_users, num_items, rank, world, exchange="p2p",
                 async_upload=False, partition=None, sliced_arrays=False):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.rank, self.world, self.k = rank, world, k
        self.nu, self.ni = num_users, num_items
        self.exchange = exchange
        # sliced_arrays: problem["user_ids" / "item_ids" / "ratings"] are THIS RANK'S slice of a
        # problem of problem["num_ratings_total"] ratings starting at problem["slice_begin"], and
        # problem["user_factors_rows" / "item_factors_rows"] hold just the factor rows of
        # problem["user_rows" / "item_rows"] -- no process holds the whole problem on the host
        # (config 5: 16 GB of ratings, 10 GB of user factors)
        self.sliced_arrays = sliced_arrays
        self.nnz = int(problem["num_ratings_total"]) if sliced_arrays else len(problem["ratings"])
        self.partition = (1 if exchange == "p2p" else 0) if partition is None else partition
        import os
        import time
        marks = [("start", time.time())]
        self.device = torch.device("cuda", torch.cuda.current_device())
        self.u_rows = io_slice(num_users, rank, world)
        self.i_rows = io_slice(num_items, rank, world)
        if sliced_arrays:
            if exchange != "p2p":
                raise ValueError("sliced host arrays need the peer-to-peer path")
            self.u_rows, self.i_rows = tuple(problem["user_rows"]), tuple(problem["item_rows"])
        if exchange == "p2p":
            # the device-side barriers below spin (with a 10 s timeout) until every rank has
            # arrived: enter the construction together, however long each rank's host-side
            # preparation took
            if world > 1:
                dist.barrier()
            if sliced_arrays:
                b = int(problem["slice_begin"])
                self.prob = cpp_ls.AlsProblem(problem["user_ids"], problem["item_ids"], problem["ratings"],
                                              k, num_users, num_items,
                                              coo_slice=(b, b + len(problem["ratings"])), total_ratings=self.nnz)
            else:
                self.prob = cpp_ls.AlsProblem(problem["user_ids"], problem["item_ids"], problem["ratings"],
                                              k, num_users, num_items,
                                              coo_slice=io_slice(self.nnz, rank, world))
            marks.append(("create (slice upload enqueued)", time.time()))
            if world > 1:
                mine = torch.frombuffer(bytearray(self.prob.ipc_handles_all()), dtype=torch.uint8).to(self.device)
                allh = torch.empty(384 * world, dtype=torch.uint8, device=self.device)
                dist.all_gather_into_tensor(allh, mine)
                allh = bytes(allh.cpu().numpy())
            else:
                allh = self.prob.ipc_handles_all()
            marks.append(("handle exchange", time.time()))
            self.prob.open_peer_group([allh[384 * r:384 * (r + 1)] for r in range(world)], rank,
                                      self.partition)
            marks.append(("open peers", time.time()))
            self.prob.push_coo()            # own slice -> every peer (NVLink)
            self.prob.peer_barrier()        # problem stream: every rank's slice is everywhere
            if sliced_arrays:
                self.prob.upload_factor_rows(problem["user_factors_rows"], problem["item_factors_rows"],
                                             *self.u_rows, *self.i_rows, rows_only=True)
            else:
                uf0 = problem["user_factors0"].reshape(-1)
                itf0 = problem["item_factors0"].reshape(-1)
                self.prob.upload_factor_rows(uf0, itf0, *self.u_rows, *self.i_rows)
            self.prob.build_index()         # id check, both groupings, this rank's work lists
            marks.append(("index + work lists", time.time()))
            # second barrier, on the stream the sweeps use and behind this rank's factor pushes:
            # past it every replica holds the complete initial factors
            self.prob.peer_barrier(torch.cuda.current_stream().cuda_stream)
            self.ranges = None
        else:
            self.prob = cpp_ls.AlsProblem(problem["user_ids"], problem["item_ids"], problem["ratings"],
                                          k, num_users, num_items)
            marks.append(("problem", time.time()))
            # async_upload: the factors ride the copy stream behind the ratings (the caller keeps
            # problem["*_factors0"] alive and untouched until the first get_factors)
            self.prob.set_factors(problem["user_factors0"], problem["item_factors0"],
                                  wait=not async_upload)
            self.prob.set_shard_partition(rank, world, 0)
            self.ranges = self.prob.shard_ranges()
            pu, pi = self.prob.device_factors()
            self.uf = torch.as_tensor(_DevArray(pu, num_users * (k + 1)), device=self.device)
            self.itf = torch.as_tensor(_DevArray(pi, num_items * k), device=self.device)
            all_ranges = [None] * world
            dist.all_gather_object(all_ranges, self.ranges)
            self.u_views = [self.uf[r[0] * (k + 1):r[1] * (k + 1)] for r in all_ranges]
            self.i_views = [self.itf[r[2] * k:r[3] * k] for r in all_ranges]
            self.prob.finish_uploads()
            dist.barrier()
        marks.append(("ready", time.time()))
        if rank == 0 and os.environ.get("MRB_E2E_TIMING"):
            import sys
            print("[ShardedAls] " + ", ".join("%s %.1f ms" % (marks[i][0], (marks[i][1] - marks[i - 1][1]) * 1e3)
                                              for i in range(1, len(marks))), file=sys.stderr)

    def close(self):
        self.prob.close()

    def _exchange(self, user_side, stream):
        if self.exchange == "p2p":
            # rows are already in every replica; wait until every rank's kernel has finished
            self.prob.peer_barrier(stream)
        else:
            views = self.u_views if user_side else self.i_views
            self.dist.all_gather(views, views[self.rank])

    def sweep(self):
        stream = self.torch.cuda.current_stream().cuda_stream
        self.prob.half_sweep(True, stream)
        self._exchange(True, stream)
        self.prob.half_sweep(False, stream)
        self._exchange(False, stream)

    def sse(self):
        """Training sum of squared errors after the last sweep, summed over ranks."""
        stream = self.torch.cuda.current_stream().cuda_stream
        t = self.torch.tensor([self.prob.shard_sse(stream)], dtype=self.torch.float64,
                              device=self.device)
        self.dist.all_reduce(t)
        return float(t.item())

    def download_own_rows(self, user_factors, item_factors):
        """This rank's I/O rows of the (replicated, complete) factor matrices into full-size
        host arrays; the union over ranks is the whole result."""
        stream = self.torch.cuda.current_stream().cuda_stream
        self.prob.download_factor_rows(user_factors, item_factors, *self.u_rows, *self.i_rows, stream)

    def download_own_rows_into(self, user_rows_array, item_rows_array):
        """The same with arrays that hold just this rank's I/O rows (``sliced_arrays`` problems)."""
        stream = self.torch.cuda.current_stream().cuda_stream
        self.prob.download_factor_rows(user_rows_array, item_rows_array, *self.u_rows, *self.i_rows,
                                       stream, rows_only=True)

    def bench(self, algorithm, warmup, steps, sampler=None):
        torch, dist = self.torch, self.dist
        if algorithm != 4:
            raise ValueError("the sharded path implements algorithm 4 (exact half-sweeps)")
        for _ in range(warmup):
            self.sweep()
        self.prob.collect_gram_ms()
        launches0 = cpp_ls.kernel_launches()
        dist.barrier()
        torch.cuda.synchronize()
        if sampler is not None:
            sampler.mark_begin()
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        import time
        t0 = time.time()
        e0.record()
        for _ in range(steps):
            self.sweep()
        e1.record()
        torch.cuda.synchronize()
        dist.barrier()
        wall_ms = (time.time() - t0) * 1e3
        if sampler is not None:
            sampler.mark_end()
        my_gram = self.prob.collect_gram_ms()
        ms = torch.tensor([e0.elapsed_time(e1), my_gram], dtype=torch.float64, device=self.device)
        per_rank = torch.zeros(self.world, dtype=torch.float64, device=self.device)
        per_rank[self.rank] = my_gram / (2.0 * steps)
        dist.all_reduce(per_rank)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)     # device time = max over ranks
        launches = cpp_ls.kernel_launches() - launches0
        sse = self.sse()
        uf, itf = self.prob.get_factors()
        # the timed region is a few tens of ms at N > 1: too short for nvidia-smi's sampling
        # period, so the identical sweep keeps running untimed for ~0.6 s to be sampled
        need = torch.tensor([1.0 if (sampler is not None and sampler.needs_continuation()) else 0.0],
                            device=self.device)
        dist.all_reduce(need, op=dist.ReduceOp.MAX)
        if float(need.item()) > 0:
            c0 = time.time()
            # the count must be the SAME on every rank (every sweep holds two device barriers):
            # it is derived from the all-reduced device time, never from this rank's own clock
            for _ in range(max(8, int(0.6 / max(float(ms[0]) / steps * 1e-3, 1e-4)))):
                self.sweep()
            torch.cuda.synchronize()
            if sampler is not None:
                sampler.mark_continuation(c0, time.time())
            self.prob.collect_gram_ms()
        clocks = sampler.stop() if sampler is not None else None
        return dict(device_ms=float(ms[0]), gram_ms=float(ms[1]), wall_ms=wall_ms, clocks=clocks,
                    launches=int(launches) * self.world, user_factors=uf, item_factors=itf,
                    sse=sse, exchange=self.exchange, ranges=self.ranges,
                    per_rank_gram_ms=[round(float(v), 4) for v in per_rank.cpu()])


def e2e_steps(problem, k, num_users, num_items, rank, world, steps, exchange="p2p", timing=False):
    """End-to-end time of one sharded sweep from HOST buffers, per step: every rank uploads its
    1/world slice of the ratings and of the factors (pushed on to the peers over NVLink), builds
    the indices and its work lists, runs one sweep and copies its 1/world share of the factor
    rows back INTO `problem["user_factors0"/"item_factors0"]` (pass page-locked arrays).  After a
    step the union of the ranks' host rows is the complete result, and it is what the next step
    uploads.  Returns a dict: seconds per step (max over ranks), bytes moved per step over all
    ranks, a description of the call."""
    import time

    import torch
    import torch.distributed as dist
    times = []
    uf_host = problem["user_factors0"].reshape(-1)
    itf_host = problem["item_factors0"].reshape(-1)
    nnz = len(problem["ratings"])
    for step in range(steps + 1):              # step 0 is a warm-up
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.time()
        s = ShardedAls(problem, k, num_users, num_items, rank, world, exchange=exchange,
                       async_upload=True)
        t1 = time.time()
        s.sweep()
        if exchange == "p2p":
            s.download_own_rows(uf_host, itf_host)      # synchronises the stream
        else:
            torch.cuda.synchronize()
            s.prob.get_factors(uf_host, itf_host)
        t2 = time.time()
        dt = torch.tensor([t2 - t0], dtype=torch.float64, device=s.device)
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        s.close()
        if timing and rank == 0:
            import sys
            print("[e2e N=%d step %d] setup %.1f ms, sweep + download %.1f ms" % (
                world, step, (t1 - t0) * 1e3, (t2 - t1) * 1e3), file=sys.stderr)
        if step > 0:
            times.append(float(dt.item()))
    fbytes = (num_users * (k + 1) + num_items * k) * 8
    mult = 1 if exchange == "p2p" else world
    return dict(sec_per_step=sum(times) / len(times), steps=len(times),
                h2d_bytes_per_step=mult * (nnz * 16 + fbytes), d2h_bytes_per_step=mult * fbytes,
                call="sharded.ShardedAls(host buffers) + one sweep + factor download per step on "
                     "every rank; each rank moves its 1/%d slice of the ratings and factors across "
                     "its host link, slices are exchanged over NVLink (peer copies), index build, "
                     "work lists and peer mapping exchange included" % world
                if exchange == "p2p" else
                "ShardedAls(exchange='nccl'): every rank uploads everything")


# ----------------------------------------------------------------------------------------------
# Similarity over several GPUs: query movies are independent, so the split
# is by contiguous query block with the whole catalogue replicated on every rank; the only
# exchange is the final gather of the per-block results.  The reference splits the same job the
# same way over its worker processes (movie_lens_data_proc.py:127-150 split_range_and_send,
# :657-700 _find_similar_movies, :264-279 update_var_into_dict).
# ----------------------------------------------------------------------------------------------
def query_blocks(num_queries, world):
    """bounds[0..world]: contiguous query blocks whose sizes differ by at most one."""
    return [num_queries * r // world for r in range(world + 1)]


def _gather_rows(local, bounds, rank, world, fill):
    """All-gather of per-rank row blocks of different heights (padded to the tallest block):
    every rank returns the stacked (bounds[-1], cols) array."""
    import torch
    import torch.distributed as dist
    tallest = max(bounds[r + 1] - bounds[r] for r in range(world))
    device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" \
        else torch.device("cpu")
    padded = np.full((tallest,) + local.shape[1:], fill, dtype=local.dtype)
    padded[:local.shape[0]] = local
    mine = torch.from_numpy(padded).to(device)
    parts = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine)
    return np.concatenate([parts[r][:bounds[r + 1] - bounds[r]].cpu().numpy() for r in range(world)])


def sharded_factor_cosine_topk(item_factors, topk, rank, world, num_factors=None, compute=None):
    """Config 4 on ``world`` GPUs: this rank computes the top-``topk`` of its query block
    (``similarity.factor_cosine_topk(..., q_lo, q_hi)``), then the ``(N / world) x topk`` id and
    score blocks are all-gathered.  Returns ``(ids int32[N, topk], scores f64[N, topk])`` on every
    rank; ``compute`` replaces the GPU call (host-logic tests)."""
    M = np.ascontiguousarray(item_factors, dtype=np.float64)
    if M.ndim == 1:
        M = M.reshape(-1, num_factors)
    bounds = query_blocks(M.shape[0], world)
    if compute is None:
        from . import similarity

        def compute(q_lo, q_hi):
            ids, scores, _ = similarity.factor_cosine_topk(M, topk=topk, q_lo=q_lo, q_hi=q_hi)
            return ids, scores
    ids, scores = compute(bounds[rank], bounds[rank + 1])
    if world == 1:
        return ids, scores
    return (_gather_rows(np.ascontiguousarray(ids, dtype=np.int32), bounds, rank, world, -1),
            _gather_rows(np.ascontiguousarray(scores, dtype=np.float64), bounds, rank, world, 0.0))


def sharded_build_similar_movies(finder, rank, world, num_results=20):
    """The reference's similar_movies.bin job on ``world`` GPUs: every rank holds the whole
    ``SimilarMovieFinder`` (anything with ``num_movies`` and ``build(num_results, start, length)``)
    and builds ``{movie_id: (similar movie ids)}`` for its block of the movie list; the
    dictionaries are merged on every rank in rank order, as the reference merges its workers'
    (movie_lens_data_proc.py:264-279)."""
    import torch.distributed as dist
    bounds = query_blocks(finder.num_movies, world)
    mine = finder.build(num_results=num_results, start=bounds[rank],
                        length=bounds[rank + 1] - bounds[rank])
    if world == 1:
        return mine
    parts = [None] * world
    dist.all_gather_object(parts, mine)
    merged = {}
    for part in parts:
        merged.update(part)
    return merged
