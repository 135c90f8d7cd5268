// Wide ranks (k = 57 ... 159; config 5 has k = 128): the lower triangle of a k = 128 Gram matrix is
// 153 tiles of 8x8 -- 306 accumulator registers per lane, more than one warp has.  One CTA of four
// warps therefore owns an owner (a user / a movie):
//
//   accumulate  the tile ROWS are cut into four bands of (nearly) equal tile count, one per warp
//               (M8 = 17: rows [0,8) [8,12) [12,15) [15,17) = 36 / 42 / 42 / 33 tiles).  The four
//               warps walk the owner's ratings together: per k-step (4 ratings) warp w copies
//               rating w's factor row into a shared-memory ring (cp.async, 3 k-steps in flight, one
//               barrier per k-step), every warp reads its band's mma fragments from there
//               (conflict free: the padded row length is 8 mod 16 doubles) and issues its DMMAs
//               (mma.sync m8n8k4 f64).  The rows cross L2 -> SM once per CTA, not once per warp or
//               once per 4x7 block as in the block-wise path this replaces.
//   hand-over   the tiles go to shared memory in fragment order (lane-contiguous 16 bytes: conflict
//               free both ways).  Owners with more than GRAM_SEG ratings are cut into segments
//               handled by different CTAs; their partial tiles meet in HBM and the CTA that
//               arrives last sums them in segment order (deterministic).
//   factorise   blocked LEFT-looking Cholesky on the shared-memory tiles, the same mathematics as
//               gram_solve (correction form, pivots below 1e-12 of the original diagonal skipped):
//               per tile column every warp brings its tiles up to date in registers against the
//               finished columns (T -= L L^T on the tensor cores; the operands are the stored C
//               fragments -- the even/odd column split makes them valid A / B fragments as they
//               are), warp 0 factors the 8x8 diagonal tile with an identity tile riding along
//               (-> W = L_d^-T) meanwhile, then the panel tiles become X W.  Every tile is written
//               once; two barriers per tile column.
//   solve       warp 0: delta_t = W_t (y_t - sum L^T delta), one 8x8 product per tile row; the
//               solved row goes to this GPU's replica and to every peer replica.
#pragma once
#include "gram_solve.cuh"

namespace mrb {

namespace {

constexpr int WIDE_WARPS = 4;

// first tile row of band w (w = 0..4) for M8 tile rows: the cut that minimises the largest band
__host__ __device__ constexpr int wide_band(int m8, int w) {
    int best0 = 1, best1 = 2, best2 = 3, best = 1 << 30;
    for (int a = 1; a < m8; a++)
        for (int b = a + 1; b < m8; b++)
            for (int c = b + 1; c < m8; c++) {
                const int t0 = a * (a + 1) / 2, t1 = b * (b + 1) / 2 - t0 - 0,
                          t2 = c * (c + 1) / 2 - b * (b + 1) / 2, t3 = m8 * (m8 + 1) / 2 - c * (c + 1) / 2;
                int mx = t0 > t1 ? t0 : t1;
                mx = mx > t2 ? mx : t2;
                mx = mx > t3 ? mx : t3;
                if (mx < best) { best = mx; best0 = a; best1 = b; best2 = c; }
            }
    return w == 0 ? 0 : w == 1 ? best0 : w == 2 ? best1 : w == 3 ? best2 : m8;
}

template <int M8>
struct WideSmem {
    static constexpr int ST = M8 * (M8 + 1) / 2;
    double S[ST * 64];        // lower-triangular tiles, fragment order: [tile][lane][2]
    double Wc[M8 * 64];       // W = L_d^-T per tile column, C-fragment order
    double Wt[M8 * 64];       // W^T per tile column, C-fragment order (= B operand of X W)
    double x0[M8 * 8];        // the owner's current factors (0 beyond n)
    double thr[M8 * 8];       // pivot thresholds: 1e-12 of the original diagonal
    double gd[M8 * 8];        // terms of x0.(g + g')
    double corner;            // sum b^2, later the residual corner
    int ticket, last;         // scheduler broadcast
};

// element (i, j), i >= j, of the tile array
__device__ __forceinline__ int wide_at(int i, int j) {
    return TI(i >> 3, j >> 3) * 64 + (((i & 7) << 2) + ((j & 7) >> 1)) * 2 + (j & 1);
}

// The 8-pivot factorisation of one diagonal tile with an identity tile riding along (the pivot
// loop of gram_solve restricted to the tiles D and W): on return D = L_d (lower part), W = L_d^-T.
// npiv < 8 only in the last tile column (index n, the right-hand side, is not a pivot); *corner
// receives element (npiv, npiv) before the scaling -- the residual corner -- on every lane.
// (Measured and rejected: every lane holding the whole 8x8 tile and factoring it redundantly in
// registers, no shuffles, results through a 1 KB scratch -- 58.3 vs 56.5 ms on the user side of
// a 27 M-rating k = 128 problem: the scalar chain is no shorter than the shuffle round trips.)
__device__ __forceinline__ void wide_factor_diag(double& d0, double& d1, double& w0, double& w1,
                                                 double thr, int npiv, int lane, double* corner) {
    const int p = lane >> 2, q = lane & 3;
    double invd = 0;
    // fully unrolled: one copy of this code per kernel (the tile-column loop around it is rolled),
    // and with c, cp static every lane predicate and shuffle source folds to a constant (79.8 vs
    // 82.0 ms per k = 128 sweep against the rolled-in-pairs form, profiles/ab_unroll_r02.log)
#pragma unroll
    for (int cp = 0; cp < 4; cp++) {
#pragma unroll
        for (int j = 0; j < 2; j++) {
            const int c = 2 * cp + j;
            if (c < npiv) {
                const double dv = j ? d1 : d0;
                const bool ok = dv > thr && thr > 1e-290;
                const double r = shfl_double(ok ? fast_rsqrt(dv) : 0.0, c * 4 + cp);
                if (p == c) invd = r;
                if (c < 7) {
                    const double inv_d = r * r;
                    const double m0 = shfl_double(dv, (2 * q) * 4 + cp);
                    const double m1 = shfl_double(dv, (2 * q + 1) * 4 + cp);
                    const double f0 = q > cp ? m0 * inv_d : 0.0;
                    const double f1 = (j == 0 ? q >= cp : q > cp) ? m1 * inv_d : 0.0;
                    const double xrc = shfl_double(dv, p * 4 + cp);
                    d0 = fma(-xrc, f0, d0);
                    d1 = fma(-xrc, f1, d1);
                    const double wrc = shfl_double(j ? w1 : w0, p * 4 + cp);
                    w0 = fma(-wrc, f0, w0);
                    w1 = fma(-wrc, f1, w1);
                }
            }
        }
    }
    if (corner != nullptr) {
        const int pr = npiv;   // 0..7
        *corner = shfl_double((pr & 1) ? d1 : d0, pr * 4 + (pr >> 1));
    }
    const double r0 = shfl_double(invd, (2 * q) * 4), r1 = shfl_double(invd, (2 * q + 1) * 4);
    d0 *= r0;
    d1 *= r1;
    w0 *= r0;
    w1 *= r1;
}

constexpr int WIDE_STAGES = 3;     // k-steps of gathered rows in flight in shared memory

template <int M8>
struct WideStage {
    double row[WIDE_STAGES][4][M8 * 8];   // 4 ratings per k-step; row length 8 (mod 16) doubles: conflict free
};

__device__ __forceinline__ void cp_async8(void* smem, const void* gmem) {
    const unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem));
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit_group() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait_group() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// One warp's band of the accumulation: tile rows [R0, R1), all columns up to the diagonal.  The
// four warps of the CTA walk the owner's ratings TOGETHER: warp w copies the factor row of rating
// 4 st + w of k-step st into shared memory (cp.async, WIDE_STAGES k-steps ahead, one barrier per
// k-step), and every warp reads the fragments of its band from there -- the rows cross L2 -> SM
// once per CTA instead of once per warp.
template <int M8, bool USER, int R0, int R1>
__device__ __forceinline__ void wide_accumulate(const GramArgs& A, const WorkItem& wi, int warp, int lane,
                                                WideStage<M8>& stg, double* __restrict__ out_tiles,
                                                const double* zero_row) {
    constexpr int T0 = R0 * (R0 + 1) / 2;
    constexpr int NT = R1 * (R1 + 1) / 2 - T0;
    const int q = lane & 3, p = lane >> 2;
    const int k = A.k;
    const int row_len = USER ? k : k + 1;          // doubles per gathered factor row
    double acc[NT][2];
#pragma unroll
    for (int t = 0; t < NT; t++) { acc[t][0] = 0; acc[t][1] = 0; }
    const int cnt = wi.end - wi.beg;
    const int nsteps = (cnt + 3) >> 2;
    // ids / ratings of 32 grouped positions per coalesced batch, two batches in registers
    int ids_cur = 0, ids_nxt = 0;
    double rts_cur = 0, rts_nxt = 0;
    if (lane < cnt) { ids_cur = A.other_g[wi.beg + lane]; rts_cur = A.rating_g[wi.beg + lane]; }
    if (32 + lane < cnt) { ids_nxt = A.other_g[wi.beg + 32 + lane]; rts_nxt = A.rating_g[wi.beg + 32 + lane]; }
    int batch_of_cur = 0;                           // ids_cur holds batch batch_of_cur, ids_nxt the next one
    // request k-step st: this warp copies rating 4 st + warp's row
    auto request = [&](int st) {
        if (st < nsteps) {
            const bool from_next = (st >> 3) != batch_of_cur;
            const int src = ((st & 7) << 2) + warp;
            const int id = __shfl_sync(0xffffffffu, from_next ? ids_nxt : ids_cur, src);
            const bool valid = (st << 2) + warp < cnt;
            const double* row = valid ? A.other_f + static_cast<size_t>(id) * A.other_stride : zero_row;
            double* dstrow = stg.row[st % WIDE_STAGES][warp];
            for (int e = lane; e < row_len; e += 32) cp_async8(dstrow + e, row + e);
        }
        cp_async_commit_group();                    // (possibly empty: keeps the group count in step)
    };
    // zero the tail of every staged row once (elements row_len .. 8 M8 - 1 are never copied)
    for (int sidx = 0; sidx < WIDE_STAGES; sidx++)
        for (int e = row_len + lane; e < M8 * 8; e += 32) stg.row[sidx][warp][e] = 0.0;
#pragma unroll
    for (int st = 0; st < WIDE_STAGES - 1; st++) request(st);
    for (int step = 0; step < nsteps; step++) {
        cp_async_wait_group<WIDE_STAGES - 2>();     // this thread's copies of k-step `step` have landed
        __syncthreads();                            // ... everyone's have, and k-step step-1 is consumed
        // the ratings of this k-step (for the special elements) before the batch registers move on
        const bool fn = (step >> 3) != batch_of_cur;
        const double rt = shfl_double(fn ? rts_nxt : rts_cur, ((step & 7) << 2) + q);
        const bool valid = (step << 2) + q < cnt;
        // refill the buffer k-step step-1 used; entering a new batch of 32 ratings first
        const int nxt = step + WIDE_STAGES - 1;
        if ((nxt >> 3) > batch_of_cur + 1) {
            ids_cur = ids_nxt;
            rts_cur = rts_nxt;
            batch_of_cur++;
            const int e = ((batch_of_cur + 1) << 5) + lane;
            ids_nxt = 0;
            rts_nxt = 0;
            if (e < cnt) { ids_nxt = A.other_g[wi.beg + e]; rts_nxt = A.rating_g[wi.beg + e]; }
        }
        request(nxt);
        const double* srow = stg.row[step % WIDE_STAGES][q];
        double f[R1];
#pragma unroll
        for (int t = 0; t < R1; t++) {
            const int j = 8 * t + p;
            double v = srow[j];                                   // raw factor (0 beyond the row)
            // only the last two (users) / the last (movies) tiles can hold special elements: the
            // order k + 2 (k + 1) fills the last tile, so every tile before them is plain -- a
            // compile-time fact (a run-time test on k cost a select per element of EVERY tile:
            // 30 % of the accumulation's instructions, profiles/ncu_k_gram_wide_k128_r02_v2.txt)
            if (t >= (USER ? M8 - 2 : M8 - 1)) {
                if (USER) v = j < k ? v : (!valid ? 0.0 : (j == k ? 1.0 : (j == k + 1 ? rt : 0.0)));
                else v = j < k ? v : ((valid && j == k) ? rt - v : 0.0);   // rating - user bias (matrix.cpp:1029)
            }
            f[t] = v;
        }
#pragma unroll
        for (int ti = R0; ti < R1; ti++)
#pragma unroll
            for (int tj = 0; tj <= ti; tj++)
                dmma884(acc[TI(ti, tj) - T0][0], acc[TI(ti, tj) - T0][1], f[ti], f[tj]);
    }
    cp_async_wait_group<0>();
#pragma unroll
    for (int t = 0; t < NT; t++)
        *reinterpret_cast<double2*>(out_tiles + (static_cast<size_t>(T0 + t) * 32 + lane) * 2) =
            make_double2(acc[t][0], acc[t][1]);
}

// Sum of the segments' partial tiles of one band, in segment order, into shared memory.
template <int M8, int R0, int R1>
__device__ __forceinline__ void wide_reduce_band(const double* __restrict__ partials, int nseg, int lane,
                                                 double* __restrict__ S) {
    constexpr int ST = M8 * (M8 + 1) / 2;
    constexpr int T0 = R0 * (R0 + 1) / 2, T1 = R1 * (R1 + 1) / 2;
    for (int t = T0; t < T1; t++) {
        double2 a = make_double2(0.0, 0.0);
        for (int s = 0; s < nseg; s++) {
            const double2 v = __ldcg(reinterpret_cast<const double2*>(
                partials + (static_cast<size_t>(s) * ST + t) * 64 + lane * 2));
            a.x += v.x;
            a.y += v.y;
        }
        *reinterpret_cast<double2*>(S + (static_cast<size_t>(t) * 32 + lane) * 2) = a;
    }
}

template <int M8, bool USER>
__global__ void __launch_bounds__(WIDE_WARPS * 32, 2)
k_gram_wide(const GramArgs A, const double* __restrict__ zero_row) {
    constexpr int ST = M8 * (M8 + 1) / 2;
    constexpr int TN = M8 - 1;
    constexpr int NS = (M8 + WIDE_WARPS - 1) / WIDE_WARPS;      // tiles of one tile column per warp
    constexpr int B0 = wide_band(M8, 0), B1 = wide_band(M8, 1), B2 = wide_band(M8, 2),
                  B3 = wide_band(M8, 3), B4 = wide_band(M8, 4);
    extern __shared__ __align__(16) unsigned char wide_smem_raw[];
    WideSmem<M8>& sm = *reinterpret_cast<WideSmem<M8>*>(wide_smem_raw);
    WideStage<M8>& stg = *reinterpret_cast<WideStage<M8>*>(wide_smem_raw + sizeof(WideSmem<M8>));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int p = lane >> 2, q = lane & 3;
    const int n = A.n;
    const int pr = n & 7;          // fragment row / column of index n (the right-hand side) in tile TN
    auto tile = [&](int ti, int tj) { return sm.S + (static_cast<size_t>(TI(ti, tj)) * 32 + lane) * 2; };

    for (;;) {
        __syncthreads();           // the previous owner's shared memory is no longer needed
        if (threadIdx.x == 0) sm.ticket = atomicAdd(A.work_counter, 1);
        __syncthreads();
        const int w = sm.ticket;
        if (w >= A.n_work) break;
        const WorkItem wi = A.work[w];
        const bool multi = wi.nseg > 1;
        double* dst = multi ? A.partials + (static_cast<size_t>(wi.slot) + wi.seg) * (ST * 64) : sm.S;
        switch (warp) {
            case 0: wide_accumulate<M8, USER, B0, B1>(A, wi, warp, lane, stg, dst, zero_row); break;
            case 1: wide_accumulate<M8, USER, B1, B2>(A, wi, warp, lane, stg, dst, zero_row); break;
            case 2: wide_accumulate<M8, USER, B2, B3>(A, wi, warp, lane, stg, dst, zero_row); break;
            default: wide_accumulate<M8, USER, B3, B4>(A, wi, warp, lane, stg, dst, zero_row); break;
        }
        if (multi) {
            // ordered reduction of the segments by the CTA that arrives last
            __threadfence();
            __syncthreads();
            if (threadIdx.x == 0) sm.last = atomicAdd(A.seg_done + wi.multi, 1) == wi.nseg - 1;
            __syncthreads();
            if (!sm.last) continue;
            __threadfence();
            const double* src = A.partials + static_cast<size_t>(wi.slot) * (ST * 64);
            switch (warp) {
                case 0: wide_reduce_band<M8, B0, B1>(src, wi.nseg, lane, sm.S); break;
                case 1: wide_reduce_band<M8, B1, B2>(src, wi.nseg, lane, sm.S); break;
                case 2: wide_reduce_band<M8, B2, B3>(src, wi.nseg, lane, sm.S); break;
                default: wide_reduce_band<M8, B3, B4>(src, wi.nseg, lane, sm.S); break;
            }
        }
        if (A.debug_skip_solve & 1) continue;     // measurement only: time the accumulation alone
        double* xo = A.x + static_cast<size_t>(wi.owner) * n;
        for (int c = threadIdx.x; c < M8 * 8; c += blockDim.x) sm.x0[c] = c < n ? xo[c] : 0.0;
        __syncthreads();

        // ---- correction form: v = G x0 tile row by tile row (rows dealt over the warps; the row
        // part from the tiles left of the diagonal, the column part from the tiles below it, as in
        // gram_solve), then rhs' = g - v on row n, thresholds and the terms of x0.(g + g')
        for (int ti = warp; ti < M8; ti += WIDE_WARPS) {
            double rowsum = 0, col0 = 0, col1 = 0;
            for (int tj = 0; tj <= ti; tj++) {
                const double2 t = *reinterpret_cast<const double2*>(tile(ti, tj));
                if (tj < ti) rowsum += t.x * sm.x0[8 * tj + 2 * q] + t.y * sm.x0[8 * tj + 2 * q + 1];
                else { col0 += t.x * sm.x0[8 * ti + p]; col1 += t.y * sm.x0[8 * ti + p]; }   // diagonal tile: once
            }
            for (int tk = ti + 1; tk < M8; tk++) {
                const double2 t = *reinterpret_cast<const double2*>(tile(tk, ti));
                const double xr = sm.x0[8 * tk + p];
                col0 += t.x * xr;
                col1 += t.y * xr;
            }
            col0 = xor_sum_p(col0);
            col1 = xor_sum_p(col1);
            rowsum = xor_sum_q(rowsum);                       // indexed by p, replicated over q
            const double v0 = col0 + shfl_double(rowsum, (2 * q) * 4);
            const double v1 = col1 + shfl_double(rowsum, (2 * q + 1) * 4);
            if (p == 0) {                                     // column layout: lanes (0, q) own columns 2q, 2q+1
                sm.gd[8 * ti + 2 * q] = v0;                   // (gd doubles as the scratch for v)
                sm.gd[8 * ti + 2 * q + 1] = v1;
            }
        }
        __syncthreads();
        for (int c = threadIdx.x; c < M8 * 8; c += blockDim.x) {
            if (c < n) {
                const double g0 = sm.S[wide_at(n, c)];
                const double r = g0 - sm.gd[c];
                sm.thr[c] = 1e-12 * sm.S[wide_at(c, c)];
                sm.gd[c] = sm.x0[c] * (g0 + r);
                sm.S[wide_at(n, c)] = r;
            } else {
                sm.thr[c] = 0.0;
                sm.gd[c] = 0.0;
            }
        }
        __syncthreads();

        // ---- blocked LEFT-LOOKING Cholesky over the tile columns: every tile is read as it was
        // accumulated, brought up to date in registers against the finished columns to its left
        // (T -= L(ti,j) L(tk,j)^T, operands straight from the stored C fragments: even columns in
        // one mma, odd ones in the other), and written once, as L.  Warp 0 takes the diagonal tile
        // and factors it while the other warps update their panel tiles.
        for (int tk = 0; tk < M8; tk++) {
            const int npiv = tk < TN ? 8 : pr;
            double2 acc[NS];
            // slot s of warp w: tile row ti = tk + w + 4 s (slot 0 of warp 0 is the diagonal tile)
#pragma unroll
            for (int sl = 0; sl < NS; sl++) {
                const int ti = tk + warp + WIDE_WARPS * sl;
                acc[sl] = ti < M8 ? *reinterpret_cast<const double2*>(tile(ti, tk)) : make_double2(0.0, 0.0);
            }
            for (int j = 0; j < tk; j++) {
                const double2 lk = *reinterpret_cast<const double2*>(tile(tk, j));
#pragma unroll
                for (int sl = 0; sl < NS; sl++) {
                    const int ti = tk + warp + WIDE_WARPS * sl;
                    if (ti < M8) {                            // warp-uniform
                        const double2 li = *reinterpret_cast<const double2*>(tile(ti, j));
                        dmma884(acc[sl].x, acc[sl].y, -li.x, lk.x);
                        dmma884(acc[sl].x, acc[sl].y, -li.y, lk.y);
                    }
                }
            }
            if (warp == 0) {
                double d0 = acc[0].x, d1 = acc[0].y;
                double w0 = p == 2 * q ? 1.0 : 0.0, w1 = p == 2 * q + 1 ? 1.0 : 0.0;
                double corner = 0;
                wide_factor_diag(d0, d1, w0, w1, sm.thr[8 * tk + p], npiv, lane, tk == TN ? &corner : nullptr);
                *reinterpret_cast<double2*>(tile(tk, tk)) = make_double2(d0, d1);
                sm.Wc[(tk * 32 + lane) * 2] = w0;
                sm.Wc[(tk * 32 + lane) * 2 + 1] = w1;
                // W^T as a C fragment: slot s of lane (p, q) = W[2q + s][p]
#pragma unroll
                for (int sl = 0; sl < 2; sl++) {
                    const int src = (2 * q + sl) * 4 + (p >> 1);
                    const double v0 = shfl_double(w0, src), v1 = shfl_double(w1, src);
                    sm.Wt[(tk * 32 + lane) * 2 + sl] = (p & 1) ? v1 : v0;
                }
                if (tk == TN && lane == 0) sm.corner = corner;
            }
            __syncthreads();                                  // W^T of this column is published
            if (tk == TN) break;
            {
                const double b0 = sm.Wt[(tk * 32 + lane) * 2], b1 = sm.Wt[(tk * 32 + lane) * 2 + 1];
#pragma unroll
                for (int sl = 0; sl < NS; sl++) {
                    const int ti = tk + warp + WIDE_WARPS * sl;
                    if (ti > tk && ti < M8) {                 // panel: L(ti,tk) = X W
                        double c0 = 0.0, c1 = 0.0;
                        dmma884(c0, c1, acc[sl].x, b0);
                        dmma884(c0, c1, acc[sl].y, b1);
                        *reinterpret_cast<double2*>(tile(ti, tk)) = make_double2(c0, c1);
                    }
                }
            }
            __syncthreads();                                  // column tk is final
        }

        // ---- residual and back substitution (warp 0); row n of the factor is y = L^-1 g'
        if (warp == 0) {
            if (A.sse_out != nullptr && lane == 0) {
                double gdot = 0;
                for (int c = 0; c < n; c++) gdot += sm.gd[c];
                A.sse_out[wi.owner] = sm.corner - gdot;
            }
            double part[M8][2];
#pragma unroll
            for (int t = 0; t < M8; t++) { part[t][0] = 0; part[t][1] = 0; }
#pragma unroll
            for (int tj = M8 - 1; tj >= 0; tj--) {
                // y[8 tj + 2q + s]: row pr of tile (TN, tj)
                const double2 y = *reinterpret_cast<const double2*>(
                    sm.S + (static_cast<size_t>(TI(TN, tj)) * 32 + pr * 4 + q) * 2);
                double yv0 = y.x, yv1 = y.y;
                if (tj < M8 - 1) {
                    yv0 -= xor_sum_p(part[tj][0]);
                    yv1 -= xor_sum_p(part[tj][1]);
                }
                const double wc0 = sm.Wc[(tj * 32 + lane) * 2], wc1 = sm.Wc[(tj * 32 + lane) * 2 + 1];
                const double dp = xor_sum_q(fma(wc0, yv0, wc1 * yv1));     // delta[8 tj + p]
#pragma unroll
                for (int t2 = 0; t2 < tj; t2++) {
                    const double2 l = *reinterpret_cast<const double2*>(tile(tj, t2));
                    part[t2][0] = fma(l.x, dp, part[t2][0]);
                    part[t2][1] = fma(l.y, dp, part[t2][1]);
                }
                // x = x0 + delta for this tile row: delta[8 tj + p] is replicated over q
                if (q == 0 && 8 * tj + p < n) sm.x0[8 * tj + p] += dp;
            }
            // the solved row, lane-linear: full 32-byte sectors into this GPU's replica and, fused
            // all-gather, into every peer replica over NVLink
            __syncwarp();
            for (int c = lane; c < n; c += 32) {
                const double v = sm.x0[c];
                xo[c] = v;
                for (int j = 0; j < A.n_peers; j++) A.x_peers[j][static_cast<size_t>(wi.owner) * n + c] = v;
            }
        }
    }
}

}  // namespace

}  // namespace mrb
