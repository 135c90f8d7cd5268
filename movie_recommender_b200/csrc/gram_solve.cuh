// K2b -- the per-owner solve of algorithm 4 on mma accumulator fragments (gram_solve) together
// with the argument block of the gather-Gram kernel and the small warp helpers it needs.
// Split out of als_gram.cu so that the SAME source also compiles for the host, where the 32 lanes
// of a warp are emulated by 32 threads and every shuffle / mma is a rendezvous
// (tests/emu/warp_emu.h, MRB_HOST_EMU): the Cholesky epilogue gets CPU coverage
// (tests/test_emu_gram_solve.py).  Device code generation is unchanged by the split (the SASS of
// als_gram.o is byte-identical before and after).
#pragma once
#ifdef MRB_HOST_EMU
#include "warp_emu.h"        // tests/emu: __shfl_sync, __shfl_xor_sync, shfl_double, mma, rsqrt
#define MRB_DEVICE_INLINE inline
#define MRB_HOST_DEVICE
#else
#include "common.cuh"
#define MRB_DEVICE_INLINE __device__ __forceinline__
#define MRB_HOST_DEVICE __host__ __device__
#endif

namespace mrb {

namespace {

struct WorkItem {
    int owner;   // row of the factor matrix being solved
    int beg;     // first grouped rating position
    int end;     // one past the last
    int seg;     // segment index within the owner
    int nseg;    // number of segments of the owner
    int slot;    // multi-segment owners: index of the owner's first partial buffer / counter id
    int multi;   // multi-segment owners: dense id (counter index), else -1
    int pad;
};

MRB_DEVICE_INLINE void dmma884(double& c0, double& c1, double a, double b) {
#ifdef MRB_HOST_EMU
    warp_emu::mma884(c0, c1, a, b);
#else
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
#endif
}

struct GramArgs {
    const WorkItem* work;
    int n_work;
    int* work_counter;        // dynamic scheduler
    const int* other_g;       // opposite-side id of each grouped rating position
    const double* rating_g;   // rating of each grouped rating position
    const double* other_f;    // opposite-side factors
    int other_stride;
    int k;                    // gathered factor count
    int n;                    // unknowns per owner (k+1 users, k items)
    double* x;                // owner factors, in/out, row stride n
    double* partials;         // [slot][ST*64] partial augmented Gram tiles (multi-segment owners)
    int* seg_done;            // [multi] arrival counters
    double* G_out;            // EPI_STORE: [owner][n*n]
    double* g_out;            // EPI_STORE: [owner][n]
    double* sse_out;          // optional [owner]: sum of squared residuals after the solve
    double* x_peers[8];       // other replicas of the owner factor matrix (NVLink peer memory)
    int n_peers;              // number of entries of x_peers (0 on a single GPU)
    int debug_skip_solve;     // MRB_DEBUG_SKIP_SOLVE=1: time the accumulation alone (results invalid)
    int order_mode;           // how scheduler tickets map to the degree-sorted work list (work_index)
};

// index of lower-triangular tile (ti, tj), tj <= ti
MRB_HOST_DEVICE constexpr int TI(int ti, int tj) { return ti * (ti + 1) / 2 + tj; }

// 1/sqrt(d) for a normal, positive d: the hardware approximation (about 22 bits) and one
// third-order correction  y (1 + e/2 + 3 e^2/8),  e = 1 - d y^2  -- no special-case path.
MRB_DEVICE_INLINE double fast_rsqrt(double d) {
    double y;
#ifdef MRB_HOST_EMU
    y = warp_emu::rsqrt_approx(d);
#else
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
#endif
    const double e = fma(-d, y * y, 1.0);
    return fma(fma(e, 0.375, 0.5), y * e, y);
}

MRB_DEVICE_INLINE double xor_sum_p(double v) {   // sum over the 8 lanes sharing q
    v += __shfl_xor_sync(0xffffffffu, v, 4);
    v += __shfl_xor_sync(0xffffffffu, v, 8);
    v += __shfl_xor_sync(0xffffffffu, v, 16);
    return v;
}
MRB_DEVICE_INLINE double xor_sum_q(double v) {   // sum over the 4 lanes sharing p
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    return v;
}

// ------------------------------------------------------------------------------------------
// K2b: register-resident blocked Cholesky on the mma accumulator fragments.
//
// The augmented matrix [G g; g^T s] (order n+1 <= 8*M8) sits in the lower-triangular 8x8 tiles
// acc[TI(ti,tj)]; lane (p = lane>>2, q = lane&3) holds elements (row p, cols 2q, 2q+1) of every
// tile.  Right-looking by tile column tk:
//   1. the 8 pivot columns of tile column tk are eliminated one by one with warp shuffles
//      (diagonal tile and the panel tiles below it together);
//   2. the trailing tiles get  T(ti,tj) -= L(ti,tk) L(tj,tk)^T  on the tensor cores (2 DMMA per
//      tile; the C-fragment -> A/B-fragment conversion is two shuffles per 8x4 chunk).
// Row n of the factor is y = L^-1 g', so only the back substitution L^T delta = y remains; it is
// done on the fragments as well.  No shared memory is used.
// The system solved is the CORRECTION form  G delta = g - G x0,  x = x0 + delta: unknowns whose
// pivot falls below 1e-12 of the original diagonal get delta = 0 (they keep their previous value,
// as they do under the reference's warm-started CG), the others are solved consistently.
// ------------------------------------------------------------------------------------------
template <int M8, int VARIANT = 0>
MRB_DEVICE_INLINE void gram_solve(double (&acc)[M8 * (M8 + 1) / 2][2], int n,
                                           double* __restrict__ xo, double* __restrict__ sse_slot,
                                           int lane, const GramArgs& A, size_t row_offset) {
    constexpr int TN = M8 - 1;            // tile row/column holding index n (the rhs)
    const int p = lane >> 2, q = lane & 3;
    const int pr = n & 7;

    // ---- x0 in row-indexed (xp) and column-indexed (xq) layouts; zero beyond n
    double xp[M8], xq[M8][2];
#pragma unroll
    for (int t = 0; t < M8; t++) {
        const int r = 8 * t + p;
        xp[t] = r < n ? xo[r] : 0.0;
#pragma unroll
        for (int s = 0; s < 2; s++) {
            const int c = 8 * t + 2 * q + s;
            xq[t][s] = c < n ? xo[c] : 0.0;
        }
    }
    // ---- v = G x0 in column layout (replicated over p)
    double vy[M8][2];
    {
        double colsum[M8][2], rowsum[M8];
#pragma unroll
        for (int t = 0; t < M8; t++) { colsum[t][0] = 0; colsum[t][1] = 0; rowsum[t] = 0; }
#pragma unroll
        for (int ti = 0; ti < M8; ti++)
#pragma unroll
            for (int tj = 0; tj <= ti; tj++) {
                colsum[tj][0] += acc[TI(ti, tj)][0] * xp[ti];
                colsum[tj][1] += acc[TI(ti, tj)][1] * xp[ti];
                if (ti != tj) rowsum[ti] += acc[TI(ti, tj)][0] * xq[tj][0] + acc[TI(ti, tj)][1] * xq[tj][1];
            }
#pragma unroll
        for (int t = 0; t < M8; t++) {
            colsum[t][0] = xor_sum_p(colsum[t][0]);
            colsum[t][1] = xor_sum_p(colsum[t][1]);
            rowsum[t] = xor_sum_q(rowsum[t]);   // indexed by p, replicated over q
        }
#pragma unroll
        for (int t = 0; t < M8; t++)
#pragma unroll
            for (int s = 0; s < 2; s++)
                vy[t][s] = colsum[t][s] + shfl_double(rowsum[t], (2 * q + s) * 4);
    }
    // ---- rhs' = g - G x0 on the augmented row; x0.(g + g') for the residual bookkeeping
    double gdot = 0;
#pragma unroll
    for (int t = 0; t < M8; t++)
#pragma unroll
        for (int s = 0; s < 2; s++) {
            const int c = 8 * t + 2 * q + s;
            if (p == pr && c < n) {
                const double g0 = acc[TI(TN, t)][s];
                const double g1 = g0 - vy[t][s];
                acc[TI(TN, t)][s] = g1;
                gdot += xq[t][s] * (g0 + g1);
            }
        }
    gdot = xor_sum_q(gdot);   // valid on lanes with p == pr

    // ---- pivot thresholds from the original diagonal: lane (p, p>>1) holds column 8t+p's
    double thr[M8];
#pragma unroll
    for (int t = 0; t < M8; t++) thr[t] = 1e-12 * ((p & 1) ? acc[TI(t, t)][1] : acc[TI(t, t)][0]);

    double invd[M8];   // lane (p, *) : 1/L[j][j] for j = 8t+p (0 for a skipped pivot)
#pragma unroll
    for (int t = 0; t < M8; t++) invd[t] = 0;
    double corner = 0;

    if constexpr (VARIANT == 1) {
        // ---- CANDIDATE (not the default; developed on the emulated warp, tests/emu): pivots in
        // blocks of FOUR = the k of mma.m8n8k4.  Per block: the 4x4 diagonal block S and its four
        // thresholds are broadcast to every lane; every lane factors S = L4 L4^T and forms
        // W = L4^-T in registers (no shuffle inside the dependent chain; a skipped pivot has
        // r = 0, which zeroes its column of L4 and of W); the block's columns of every tile of
        // the tile column become P W with one DMMA (B operand picked from the lane's own W,
        // placed in columns 4h .. 4h+3); every trailing update, the right half of the same tile
        // column included (masked B operand), is a rank-4 DMMA.
#pragma unroll
        for (int tk = 0; tk < M8; tk++) {
            const int D = TI(tk, tk);
#pragma unroll
            for (int h = 0; h < 2; h++) {
                // warp-uniform: in the last tile column only the blocks up to the one holding
                // index n (fragment row/column pr) exist
                if (tk == M8 - 1 && 4 * h > pr) continue;
                const int l0 = (4 * h) * 4 + 2 * h;          // lane (4h, 2h): row 4h, columns 4h, 4h+1
                const double S00 = shfl_double(acc[D][0], l0);
                const double S10 = shfl_double(acc[D][0], l0 + 4), S11 = shfl_double(acc[D][1], l0 + 4);
                const double S20 = shfl_double(acc[D][0], l0 + 8), S21 = shfl_double(acc[D][1], l0 + 8);
                const double S22 = shfl_double(acc[D][0], l0 + 9);
                const double S30 = shfl_double(acc[D][0], l0 + 12), S31 = shfl_double(acc[D][1], l0 + 12);
                const double S32 = shfl_double(acc[D][0], l0 + 13), S33 = shfl_double(acc[D][1], l0 + 13);
                // threshold of column c = 4h + j sits on lane (c, c >> 1)
                const double t0 = shfl_double(thr[tk], (4 * h + 0) * 4 + 2 * h);
                const double t1 = shfl_double(thr[tk], (4 * h + 1) * 4 + 2 * h);
                const double t2 = shfl_double(thr[tk], (4 * h + 2) * 4 + 2 * h + 1);
                const double t3 = shfl_double(thr[tk], (4 * h + 3) * 4 + 2 * h + 1);
                const bool last = tk == M8 - 1;
                // local factorisation, identical on every lane
                const double d0 = S00;
                const bool ok0 = (!last || 4 * h + 0 < pr) && d0 > t0 && t0 > 1e-290;
                const double r0 = ok0 ? fast_rsqrt(d0) : 0.0;
                const double L10 = S10 * r0, L20 = S20 * r0, L30 = S30 * r0;
                const double d1 = fma(-L10, L10, S11);
                const bool ok1 = (!last || 4 * h + 1 < pr) && d1 > t1 && t1 > 1e-290;
                const double r1 = ok1 ? fast_rsqrt(d1) : 0.0;
                const double L21 = fma(-L20, L10, S21) * r1, L31 = fma(-L30, L10, S31) * r1;
                const double d2 = fma(-L21, L21, fma(-L20, L20, S22));
                const bool ok2 = (!last || 4 * h + 2 < pr) && d2 > t2 && t2 > 1e-290;
                const double r2 = ok2 ? fast_rsqrt(d2) : 0.0;
                const double L32 = fma(-L31, L21, fma(-L30, L20, S32)) * r2;
                const double d3 = fma(-L32, L32, fma(-L31, L31, fma(-L30, L30, S33)));
                const bool ok3 = (!last || 4 * h + 3 < pr) && d3 > t3 && t3 > 1e-290;
                const double r3 = ok3 ? fast_rsqrt(d3) : 0.0;
                if (last && (pr >> 2) == h) {
                    const int j = pr & 3;
                    corner = j == 0 ? d0 : (j == 1 ? d1 : (j == 2 ? d2 : d3));
                }
                // W = L4^-T, upper triangular: W[:,j] = (e_j - sum_{k<j} W[:,k] L4[j][k]) r_j
                const double W00 = r0;
                const double W01 = -(W00 * L10) * r1, W11 = r1;
                const double W02 = -fma(W01, L21, W00 * L20) * r2, W12 = -(W11 * L21) * r2, W22 = r2;
                const double W03 = -fma(W02, L32, fma(W01, L31, W00 * L30)) * r3;
                const double W13 = -fma(W12, L32, W11 * L31) * r3, W23 = -(W22 * L32) * r3, W33 = r3;
                // B operand of P W: lane (p, q) holds B[q][p] = W[q][p - 4h] inside the block
                const int jj = p & 3;
                const double w_q0 = jj == 0 ? W00 : (jj == 1 ? W01 : (jj == 2 ? W02 : W03));
                const double w_q1 = jj == 0 ? 0.0 : (jj == 1 ? W11 : (jj == 2 ? W12 : W13));
                const double w_q2 = jj < 2 ? 0.0 : (jj == 2 ? W22 : W23);
                const double w_q3 = jj == 3 ? W33 : 0.0;
                double bw = q == 0 ? w_q0 : (q == 1 ? w_q1 : (q == 2 ? w_q2 : w_q3));
                const bool in_rows = (p >> 2) == h;
                bw = in_rows ? bw : 0.0;
                if (in_rows) invd[tk] = jj == 0 ? r0 : (jj == 1 ? r1 : (jj == 2 ? r2 : r3));
                // the block's columns of the tile column: P <- P W; A fragments of the result
                const int src = p * 4 + 2 * h + (q >> 1);       // lane holding X[p][4h + q]
                const bool in_cols = (q >> 1) == h;
                double ax[M8];
#pragma unroll
                for (int ti = tk; ti < M8; ti++) {
                    const int X = TI(ti, tk);
                    const double v0 = shfl_double(acc[X][0], src);
                    const double v1 = shfl_double(acc[X][1], src);
                    const double a_old = (q & 1) ? v1 : v0;
                    double c0 = in_cols ? 0.0 : acc[X][0], c1 = in_cols ? 0.0 : acc[X][1];
                    dmma884(c0, c1, a_old, bw);
                    acc[X][0] = c0;
                    acc[X][1] = c1;
                    const double n0 = shfl_double(c0, src);
                    const double n1 = shfl_double(c1, src);
                    ax[ti] = (q & 1) ? n1 : n0;                  // L(ti,tk)[p][4h + q]
                }
                if (h == 0) {
                    // right half of the same tile column: X[:, 4..7] -= L(ti)[:, 0..3] Ld[4..7, 0..3]^T
                    const double bmask = (p >> 2) == 1 ? ax[tk] : 0.0;
#pragma unroll
                    for (int ti = tk; ti < M8; ti++)
                        dmma884(acc[TI(ti, tk)][0], acc[TI(ti, tk)][1], -ax[ti], bmask);
                }
#pragma unroll
                for (int ti = tk + 1; ti < M8; ti++)
#pragma unroll
                    for (int tj = tk + 1; tj <= ti; tj++)
                        dmma884(acc[TI(ti, tj)][0], acc[TI(ti, tj)][1], -ax[ti], ax[tj]);
            }
        }
    } else {
        // The whole factorisation is branch-free per lane (selects on multipliers, never divergent
        // control flow around a shuffle) and keeps the pivot columns UNSCALED inside a tile column:
        //   X[p][c2] -= X[p][c] * (D[c2][c] / d_c)      for the 8 pivots c of the tile column,
        // then one scaling of the finished panel by 1/sqrt(d_c) per column.  Per pivot and tile that
        // is one shuffle and two DFMAs.
    #pragma unroll
        for (int tk = 0; tk < M8; tk++) {
            const int D = TI(tk, tk);
            // the 8 pivot columns, rolled in pairs (cp = c >> 1 at run time, slot j = c & 1 static):
            // 4x less code than a full unroll -- the epilogue was instruction-cache bound
    #pragma unroll 1
            for (int cp = 0; cp < 4; cp++) {
    #pragma unroll
                for (int j = 0; j < 2; j++) {
                    const int c = 2 * cp + j;
                    // warp-uniform: only the last tile column has non-pivot columns
                    if (tk < M8 - 1 || c < pr) {
                        // the owner lane (p = c, q = cp) holds d = D[c][c] in slot j and its threshold
                        // in thr[tk]; every lane runs the reciprocal square root on its own value
                        const double dv = acc[D][j];
                        const bool ok = dv > thr[tk] && thr[tk] > 1e-290;   // false for NaN
                        const double r = shfl_double(ok ? fast_rsqrt(dv) : 0.0, c * 4 + cp);
                        if (p == c) invd[tk] = r;
                        if (c < 7) {
                            const double inv_d = r * r;
                            // D[c2][c] / d for this lane's two columns c2 = 2q, 2q+1 (0 for c2 <= c)
                            const double m0 = shfl_double(dv, (2 * q) * 4 + cp);
                            const double m1 = shfl_double(dv, (2 * q + 1) * 4 + cp);
                            const double f0 = q > cp ? m0 * inv_d : 0.0;
                            const double f1 = (j == 0 ? q >= cp : q > cp) ? m1 * inv_d : 0.0;
    #pragma unroll
                            for (int ti = tk; ti < M8; ti++) {
                                const int X = TI(ti, tk);
                                const double xrc = shfl_double(acc[X][j], p * 4 + cp);   // X[p][c]
                                acc[X][0] = fma(-xrc, f0, acc[X][0]);
                                acc[X][1] = fma(-xrc, f1, acc[X][1]);
                            }
                        }
                    }
                }
            }
            if (tk == M8 - 1) corner = (pr & 1) ? acc[D][1] : acc[D][0];   // valid on lane (pr, pr>>1)
            {
                // L = X diag(1/sqrt(d)); skipped pivots and the non-pivot columns become 0
                const double r0 = shfl_double(invd[tk], (2 * q) * 4);
                const double r1 = shfl_double(invd[tk], (2 * q + 1) * 4);
    #pragma unroll
                for (int ti = tk; ti < M8; ti++) {
                    acc[TI(ti, tk)][0] *= r0;
                    acc[TI(ti, tk)][1] *= r1;
                }
            }
            if (tk < M8 - 1) {
                // trailing update on the tensor cores
                double ax[M8][2];
    #pragma unroll
                for (int ti = tk + 1; ti < M8; ti++)
    #pragma unroll
                    for (int h = 0; h < 2; h++) {
                        const int src = p * 4 + 2 * h + (q >> 1);
                        const double v0 = shfl_double(acc[TI(ti, tk)][0], src);
                        const double v1 = shfl_double(acc[TI(ti, tk)][1], src);
                        ax[ti][h] = (q & 1) ? v1 : v0;   // L(ti,tk)[p][4h+q]
                    }
    #pragma unroll
                for (int ti = tk + 1; ti < M8; ti++)
    #pragma unroll
                    for (int tj = tk + 1; tj <= ti; tj++) {
                        dmma884(acc[TI(ti, tj)][0], acc[TI(ti, tj)][1], -ax[ti][0], ax[tj][0]);
                        dmma884(acc[TI(ti, tj)][0], acc[TI(ti, tj)][1], -ax[ti][1], ax[tj][1]);
                    }
            }
        }

    }

    // ---- residual: corner - x0.(g + g')  ==  sum (b - a.x)^2 at the solution
    if (sse_slot != nullptr && p == pr && q == (pr >> 1)) *sse_slot = corner - gdot;

    // ---- back substitution L^T delta = y, tile rows from the bottom.  y sits in row n of the
    // factor (tile row TN, fragment row pr); the products L(tj,t2)^T delta_tj are accumulated
    // per lane in part[] and reduced over the fragment rows once, when tile t2 is solved.
    double part[M8][2];
#pragma unroll
    for (int t = 0; t < M8; t++) { part[t][0] = 0; part[t][1] = 0; }
#pragma unroll
    for (int tj = M8 - 1; tj >= 0; tj--) {
        const int D = TI(tj, tj);
        double yv[2], rq[2], dlt[2] = {0.0, 0.0};
#pragma unroll
        for (int s = 0; s < 2; s++) {
            yv[s] = shfl_double(acc[TI(TN, tj)][s], pr * 4 + q);
            if (tj < M8 - 1) yv[s] -= xor_sum_p(part[tj][s]);
            rq[s] = shfl_double(invd[tj], (2 * q + s) * 4);
        }
#pragma unroll 1
        for (int cp = 3; cp >= 0; cp--) {
#pragma unroll
            for (int j = 1; j >= 0; j--) {
                const int c = 2 * cp + j;
                if (tj < M8 - 1 || c < pr) {
                    const double dc = shfl_double(yv[j] * rq[j], cp);   // delta[8 tj + c]
                    if (q == cp) dlt[j] = dc;
                    if (c > 0) {
                        const double l0 = shfl_double(acc[D][0], c * 4 + q);   // L[c][2q]
                        const double l1 = shfl_double(acc[D][1], c * 4 + q);   // L[c][2q+1]
                        const double d0 = (j == 0 ? q < cp : q <= cp) ? dc : 0.0;
                        const double d1 = q < cp ? dc : 0.0;
                        yv[0] = fma(-l0, d0, yv[0]);
                        yv[1] = fma(-l1, d1, yv[1]);
                    }
                }
            }
        }
        if (tj > 0) {
            const double v0 = shfl_double(dlt[0], p >> 1);
            const double v1 = shfl_double(dlt[1], p >> 1);
            const double dp = (p & 1) ? v1 : v0;   // delta[8 tj + p] (0 for rows >= n)
#pragma unroll
            for (int t2 = 0; t2 < tj; t2++) {
                part[t2][0] = fma(acc[TI(tj, t2)][0], dp, part[t2][0]);
                part[t2][1] = fma(acc[TI(tj, t2)][1], dp, part[t2][1]);
            }
        }
        xq[tj][0] += dlt[0];
        xq[tj][1] += dlt[1];
    }
    if (p == 0) {
#pragma unroll
        for (int t = 0; t < M8; t++)
#pragma unroll
            for (int s = 0; s < 2; s++) {
                const int c = 8 * t + 2 * q + s;
                if (c < n) {
                    const double v = xq[t][s];
                    xo[c] = v;
                    // fused all-gather: the solved row goes into every peer replica as well
                    for (int j = 0; j < A.n_peers; j++) A.x_peers[j][row_offset + c] = v;
                }
            }
    }
}

}  // namespace

}  // namespace mrb
