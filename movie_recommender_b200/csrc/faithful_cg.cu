// Reference-order conjugate gradient (see faithful_cg.cuh for the contract).
#include "faithful_cg.cuh"

#include <algorithm>
#include <climits>

#include "index_build.cuh"

namespace mrb {

using State = FaithfulCG::State;

// ------------------------------------------------------------------------------------------
// Dots: one CTA per reference chunk.  Warps 1..3 stream a[i]*b[i] (separately rounded) into a
// double-buffered shared tile; lane 0 of warp 0 adds the tile in index order.  The chain of
// dependent adds IS the reference's semantics (matrix.cpp:97-99) -- it cannot be split.
// ------------------------------------------------------------------------------------------
namespace {
constexpr int DOT_THREADS = 128;
constexpr int DOT_TILE = 1536;

__global__ void __launch_bounds__(DOT_THREADS)
k_dot_partials(const double* __restrict__ a, const double* __restrict__ b,
               const int* __restrict__ bounds, double* __restrict__ partials,
               const State* __restrict__ guard) {
    if (guard && guard->done) return;
    __shared__ double buf[2][DOT_TILE];
    const int beg = bounds[blockIdx.x], end = bounds[blockIdx.x + 1];
    const int len = end - beg;
    const int ntiles = (len + DOT_TILE - 1) / DOT_TILE;
    const int tid = threadIdx.x;
    auto fill = [&](int tile, int which) {
        if (tid < 32) return;
        const int off = beg + tile * DOT_TILE;
        for (int i = tid - 32; i < DOT_TILE; i += DOT_THREADS - 32) {
            const int idx = off + i;
            if (idx < end) buf[which][i] = xmul(a[idx], b[idx]);
        }
    };
    double s = 0;
    if (ntiles > 0) fill(0, 0);
    __syncthreads();
    for (int t = 0; t < ntiles; t++) {
        if (t + 1 < ntiles) fill(t + 1, (t + 1) & 1);
        if (tid == 0) {
            // the dependent-add chain; the next 8 addends are fetched from shared memory while
            // the current 8 are being added, so only the add latency is on the critical path
            const int cnt = min(DOT_TILE, len - t * DOT_TILE);
            const double* src = buf[t & 1];
            int i = 0;
            if (cnt >= 8) {
                double v0 = src[0], v1 = src[1], v2 = src[2], v3 = src[3];
                double v4 = src[4], v5 = src[5], v6 = src[6], v7 = src[7];
                for (i = 8; i + 8 <= cnt; i += 8) {
                    const double w0 = src[i], w1 = src[i + 1], w2 = src[i + 2], w3 = src[i + 3];
                    const double w4 = src[i + 4], w5 = src[i + 5], w6 = src[i + 6], w7 = src[i + 7];
                    s = xadd(s, v0); s = xadd(s, v1); s = xadd(s, v2); s = xadd(s, v3);
                    s = xadd(s, v4); s = xadd(s, v5); s = xadd(s, v6); s = xadd(s, v7);
                    v0 = w0; v1 = w1; v2 = w2; v3 = w3; v4 = w4; v5 = w5; v6 = w6; v7 = w7;
                }
                s = xadd(s, v0); s = xadd(s, v1); s = xadd(s, v2); s = xadd(s, v3);
                s = xadd(s, v4); s = xadd(s, v5); s = xadd(s, v6); s = xadd(s, v7);
            }
            for (; i < cnt; i++) s = xadd(s, src[i]);
        }
        __syncthreads();
    }
    if (tid == 0) partials[blockIdx.x] = s;
}

__device__ __forceinline__ double merge_partials(const double* partials, int T) {
    double total = 0;  // matrix.cpp:388-391
    for (int t = 0; t < T; t++) total = xadd(total, partials[t]);
    return total;
}

__global__ void k_cg_init(State* st, const double* partials, int T, double min_r_decrease,
                          int max_it) {
    const double rr = merge_partials(partials, T);  // matrix.cpp:485
    st->rr = rr;
    st->final_rr = rr;
    st->alpha = 0;
    st->beta = 0;
    st->one_minus_mrd = 1 - min_r_decrease;
    st->it = 0;
    st->slow = 0;
    st->max_it = max_it;
    st->done = (max_it <= 0 || rr < 1e-6) ? 2 : 0;  // matrix.cpp:488, :490
}

__global__ void k_cg_alpha(State* st, const double* partials, int T) {
    if (st->done) return;
    const double pAp = merge_partials(partials, T);  // matrix.cpp:497
    st->alpha = st->rr / pAp;                        // :498
}

__global__ void k_cg_beta(State* st, const double* partials, int T) {
    if (st->done) return;
    const double rr2 = merge_partials(partials, T);  // matrix.cpp:507
    st->final_rr = rr2;
    const double beta = rr2 / st->rr;                // :510
    st->beta = beta;
    if (beta > st->one_minus_mrd) st->slow++; else st->slow = 0;  // :513-516
    if (st->slow >= 2) { st->done = 1; return; }     // :518, iteration not incremented
    st->rr = rr2;
    st->it++;
    // loop head of the next iteration (:488, :490); the skipped p update is not observable
    if (st->it >= st->max_it || rr2 < 1e-6) st->done = 2;
}

__global__ void k_residual_init(const double* __restrict__ Ap, const double* __restrict__ b2,
                                double* __restrict__ r, double* __restrict__ p, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double ri = xadd(xmul(1.0, Ap[i]), xmul(-1.0, b2[i]));  // matrix.cpp:472
    r[i] = ri;
    p[i] = xmul(ri, -1.0);                                        // :476
}

__global__ void k_update_xr(const State* __restrict__ st, double* __restrict__ x,
                            double* __restrict__ r, const double* __restrict__ p,
                            const double* __restrict__ Ap, int n) {
    if (st->done) return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double alpha = st->alpha;
    x[i] = xadd(xmul(1.0, x[i]), xmul(alpha, p[i]));   // matrix.cpp:501
    r[i] = xadd(xmul(1.0, r[i]), xmul(alpha, Ap[i]));  // :504
}

__global__ void k_update_p(const State* __restrict__ st, const double* __restrict__ r,
                           double* __restrict__ p, int n) {
    if (st->done) return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    p[i] = xadd(xmul(-1.0, r[i]), xmul(st->beta, p[i]));  // matrix.cpp:521
}
}  // namespace

FaithfulCG::FaithfulCG(int rows, int cols, int thread_count, cudaStream_t s)
    : rows_(rows), cols_(cols), T_(thread_count < 1 ? 1 : thread_count), s_(s),
      b2_(cols), r_(cols), Ap_(cols), p_(cols), tmp_(rows), partials_(T_),
      row_bounds_(T_ + 1), col_bounds_(T_ + 1), one_chunk_(2), state_(1), host_state_(1) {
    std::vector<int> bd(T_ + 1);
    chunk_table(T_, rows_, bd.data());
    MRB_CUDA(cudaMemcpyAsync(row_bounds_.p, bd.data(), sizeof(int) * (T_ + 1),
                             cudaMemcpyHostToDevice, s_));
    MRB_CUDA(cudaStreamSynchronize(s_));
    chunk_table(T_, cols_, bd.data());
    MRB_CUDA(cudaMemcpyAsync(col_bounds_.p, bd.data(), sizeof(int) * (T_ + 1),
                             cudaMemcpyHostToDevice, s_));
    const int one[2] = {0, rows_};
    MRB_CUDA(cudaMemcpyAsync(one_chunk_.p, one, sizeof(one), cudaMemcpyHostToDevice, s_));
    MRB_CUDA(cudaStreamSynchronize(s_));
}

CgResult FaithfulCG::solve(FaithfulOp& A, const double* d_b, double* d_x, double min_r_decrease,
                           int max_iteration, int variant) {
    MRB_REQUIRE(A.rows == rows_ && A.cols == cols_, "FaithfulCG: operator shape mismatch");
    const int* bounds = variant == 1 ? row_bounds_.p : one_chunk_.p;
    const int nchunks = variant == 1 ? T_ : 1;
    const int vb = ceil_div(cols_ > 0 ? cols_ : 1, 256);
    State* st = state_.p;

    A.tmul(d_b, b2_.p, bounds, nchunks, s_);       // b2 = A^T b      (matrix.cpp:465 / :549)
    A.mul(d_x, tmp_.p, s_);                        // tmp = A x       (:469)
    A.tmul(tmp_.p, Ap_.p, bounds, nchunks, s_);    // Ap = A^T tmp    (:470)
    k_residual_init<<<vb, 256, 0, s_>>>(Ap_.p, b2_.p, r_.p, p_.p, cols_); MRB_LAUNCHED(1);
    k_dot_partials<<<T_, DOT_THREADS, 0, s_>>>(r_.p, r_.p, col_bounds_.p, partials_.p, nullptr); MRB_LAUNCHED(1);
    k_cg_init<<<1, 1, 0, s_>>>(st, partials_.p, T_, min_r_decrease, max_iteration); MRB_LAUNCHED(1);
    MRB_CUDA(cudaGetLastError());

    // Iterations are enqueued in small batches; every kernel is a no-op once `done` is set on
    // the device, so the host only needs to look at the state once per batch.
    const int batch = 4;
    A.guard = &st->done;
    for (;;) {
        MRB_CUDA(cudaMemcpyAsync(host_state_.p, st, sizeof(State), cudaMemcpyDeviceToHost, s_));
        MRB_CUDA(cudaStreamSynchronize(s_));
        if (host_state_.p->done) break;
        for (int i = 0; i < batch; i++) {
            A.mul(p_.p, tmp_.p, s_);                        // :493
            A.tmul(tmp_.p, Ap_.p, bounds, nchunks, s_);     // :494
            k_dot_partials<<<T_, DOT_THREADS, 0, s_>>>(p_.p, Ap_.p, col_bounds_.p, partials_.p, st); MRB_LAUNCHED(1);
            k_cg_alpha<<<1, 1, 0, s_>>>(st, partials_.p, T_); MRB_LAUNCHED(1);
            k_update_xr<<<vb, 256, 0, s_>>>(st, d_x, r_.p, p_.p, Ap_.p, cols_); MRB_LAUNCHED(1);
            k_dot_partials<<<T_, DOT_THREADS, 0, s_>>>(r_.p, r_.p, col_bounds_.p, partials_.p, st); MRB_LAUNCHED(1);
            k_cg_beta<<<1, 1, 0, s_>>>(st, partials_.p, T_); MRB_LAUNCHED(1);
            k_update_p<<<vb, 256, 0, s_>>>(st, r_.p, p_.p, cols_); MRB_LAUNCHED(1);
        }
        MRB_CUDA(cudaGetLastError());
    }
    A.guard = nullptr;
    CgResult res;
    res.iterations = host_state_.p->it;
    res.final_rr = host_state_.p->final_rr;
    return res;
}

// ------------------------------------------------------------------------------------------
// Generic CSR operator.
// ------------------------------------------------------------------------------------------
namespace {
__global__ void k_csr_mul(const int* __restrict__ rowptr, const int* __restrict__ colidx,
                          const double* __restrict__ vals, const double* __restrict__ x,
                          double* __restrict__ y, int rows, const int* __restrict__ guard) {
    if (guard && *guard) return;
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    double s = 0;
    const int end = rowptr[r + 1];
    for (int e = rowptr[r]; e < end; e++) s = xadd(s, xmul(vals[e], x[colidx[e]]));  // :203-209
    y[r] = s;
}

// chunk of row r: smallest c with r < bounds[c + 1]
__device__ __forceinline__ int chunk_of(const int* __restrict__ bounds, int nchunks, int r) {
    int lo = 0, hi = nchunks - 1;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (r < bounds[mid + 1]) hi = mid; else lo = mid + 1;
    }
    return lo;
}

__global__ void k_seg_flags(const int* __restrict__ pos, int n, const int* __restrict__ bounds,
                            int nchunks, int* __restrict__ flags) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e > n) return;
    if (e == n) { flags[n] = 0; return; }
    const int c = chunk_of(bounds, nchunks, pos[e]);
    const int cp = e > 0 ? chunk_of(bounds, nchunks, pos[e - 1]) : -1;
    flags[e] = c != cp ? 1 : 0;
}

__global__ void k_seg_group_flags(const int* __restrict__ ptr, int ngroups, int* __restrict__ flags) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g < ngroups && ptr[g] < ptr[g + 1]) flags[ptr[g]] = 1;
}

__global__ void k_seg_starts(const int* __restrict__ flags, const int* __restrict__ scan, int n,
                             int* __restrict__ seg_start) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e > n) return;
    if (e == n) { seg_start[scan[n]] = n; return; }
    if (flags[e]) seg_start[scan[e]] = e;
}

__global__ void k_grp_seg_ptr(const int* __restrict__ ptr, const int* __restrict__ scan,
                              int ngroups, int* __restrict__ grp_seg_ptr) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g <= ngroups) grp_seg_ptr[g] = scan[ptr[g]];
}

// One warp per (column, chunk) segment: 32 entries are fetched and multiplied in parallel (the
// next 32 are requested before the current ones are consumed), then added in entry order.
__global__ void __launch_bounds__(256)
k_csc_seg_partial(const int* __restrict__ seg_start, int nseg, const int* __restrict__ t_row,
                  const double* __restrict__ t_val, const double* __restrict__ t,
                  double* __restrict__ partial, const int* __restrict__ guard) {
    if (guard && *guard) return;
    const int sg = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (sg >= nseg) return;
    const int lane = threadIdx.x & 31;
    const int beg = seg_start[sg], end = seg_start[sg + 1];
    double acc = 0;
    double prod_next = 0;
    if (beg + lane < end) prod_next = xmul(t[t_row[beg + lane]], t_val[beg + lane]);   // :244
    for (int e0 = beg; e0 < end; e0 += 32) {
        const double prod_l = prod_next;
        const int en = e0 + 32 + lane;
        prod_next = 0;
        if (en < end) prod_next = xmul(t[t_row[en]], t_val[en]);
        const int cnt = min(32, end - e0);
        if (cnt == 32) {
#pragma unroll
            for (int i = 0; i < 32; i++) acc = xadd(acc, shfl_double(prod_l, i));
        } else {
            for (int i = 0; i < cnt; i++) acc = xadd(acc, shfl_double(prod_l, i));
        }
    }
    if (lane == 0) partial[sg] = acc;
}

// Fold of the per-chunk sums in chunk order (add_merge, matrix.cpp:106-125); width values per
// group, one thread per output.
__global__ void __launch_bounds__(256)
k_seg_fold(const int* __restrict__ grp_seg_ptr, const double* __restrict__ partial,
           double* __restrict__ y, long long outputs, int width, const int* __restrict__ guard) {
    if (guard && *guard) return;
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= outputs) return;
    const int g = static_cast<int>(i / width), j = static_cast<int>(i - static_cast<long long>(g) * width);
    double tot = 0;
    const int se = grp_seg_ptr[g + 1];
    for (int sg = grp_seg_ptr[g]; sg < se; sg++)
        tot = xadd(tot, partial[static_cast<size_t>(sg) * width + j]);
    y[i] = tot;
}
}  // namespace

void build_segments(SegTable& out, const int* d_grp_ptr, const int* d_pos, int ngroups, int n,
                    const int* d_bounds, int nchunks, int width, cudaStream_t s) {
    DevBuf<int> flags(static_cast<size_t>(n) + 1), scan(static_cast<size_t>(n) + 1);
    k_seg_flags<<<ceil_div(n + 1ll, 256), 256, 0, s>>>(d_pos, n, d_bounds, nchunks, flags.p);
    k_seg_group_flags<<<ceil_div(ngroups > 0 ? ngroups : 1, 256), 256, 0, s>>>(d_grp_ptr, ngroups, flags.p);
    MRB_LAUNCHED(2);
    MRB_CUDA(cudaGetLastError());
    exclusive_scan_i32(flags.p, scan.p, static_cast<long long>(n) + 1, s);
    int nseg = 0;
    MRB_CUDA(cudaMemcpyAsync(&nseg, scan.p + n, sizeof(int), cudaMemcpyDeviceToHost, s));
    MRB_CUDA(cudaStreamSynchronize(s));
    out.nseg = nseg;
    out.seg_start.alloc(static_cast<size_t>(nseg) + 1);
    out.grp_seg_ptr.alloc(static_cast<size_t>(ngroups) + 1);
    out.partial.alloc(std::max<size_t>(static_cast<size_t>(nseg) * width, 1));
    k_seg_starts<<<ceil_div(n + 1ll, 256), 256, 0, s>>>(flags.p, scan.p, n, out.seg_start.p);
    k_grp_seg_ptr<<<ceil_div(ngroups + 1ll, 256), 256, 0, s>>>(d_grp_ptr, scan.p, ngroups, out.grp_seg_ptr.p);
    MRB_LAUNCHED(2);
    MRB_CUDA(cudaGetLastError());
    MRB_CUDA(cudaStreamSynchronize(s));   // flags / scan die here
    out.bounds_key = d_bounds;
    out.nchunks_key = nchunks;
}

CsrFaithfulOp::CsrFaithfulOp(int rows_in, int cols_in, int nnz, const int* d_rowptr,
                             const int* d_colidx, const double* d_vals, cudaStream_t s)
    : nnz_(nnz), rowptr_(d_rowptr), colidx_(d_colidx), vals_(d_vals),
      t_ptr_(static_cast<size_t>(cols_in) + 1), t_row_(nnz), t_val_(nnz) {
    rows = rows_in;
    cols = cols_in;
    csr_transpose(rows, cols, nnz, d_rowptr, d_colidx, d_vals, t_ptr_.p, t_row_.p, t_val_.p, s);
}

void CsrFaithfulOp::mul(const double* d_x, double* d_y, cudaStream_t s) {
    if (rows == 0) return;
    k_csr_mul<<<ceil_div(rows, 256), 256, 0, s>>>(rowptr_, colidx_, vals_, d_x, d_y, rows, guard); MRB_LAUNCHED(1);
}

void CsrFaithfulOp::tmul(const double* d_t, double* d_y, const int* d_row_bounds, int nchunks,
                         cudaStream_t s) {
    if (cols == 0) return;
    SegTable& st = seg_[nchunks == 1 ? 1 : 0];
    if (st.bounds_key != d_row_bounds || st.nchunks_key != nchunks)
        build_segments(st, t_ptr_.p, t_row_.p, cols, nnz_, d_row_bounds, nchunks, 1, s);
    if (st.nseg > 0)
        k_csc_seg_partial<<<ceil_div(static_cast<long long>(st.nseg) * 32, 256), 256, 0, s>>>(
            st.seg_start.p, st.nseg, t_row_.p, t_val_.p, d_t, st.partial.p, guard);
    k_seg_fold<<<ceil_div(cols, 256), 256, 0, s>>>(st.grp_seg_ptr.p, st.partial.p, d_y, cols, 1, guard);
    MRB_LAUNCHED(2);
}

// ------------------------------------------------------------------------------------------
// Implicit ALS operator.
// ------------------------------------------------------------------------------------------
namespace {
constexpr int AM_WARPS = 8;   // warps per CTA in k_als_mul
constexpr int AM_JCH = 32;    // unknowns staged per pass
constexpr int AM_PS = AM_JCH + 1;  // odd row stride (doubles): conflict-free column walks

// y[r] = sum_j A[r,j] * x[owner_r*width + j], j ascending, sequential from 0.
// A warp takes 32 consecutive ratings: all lanes cooperate to fetch each rating's two rows
// (coalesced) and store the separately rounded products in shared memory; lane i then adds
// rating i's products in order.
__global__ void __launch_bounds__(AM_WARPS * 32)
k_als_mul(const int* __restrict__ owner, const int* __restrict__ other,
          const double* __restrict__ other_f, const double* __restrict__ x,
          double* __restrict__ y, int rows, int width, int other_stride, int k,
          const int* __restrict__ guard) {
    if (guard && *guard) return;
    extern __shared__ double sm[];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    double* P = sm + static_cast<size_t>(w) * 32 * AM_PS;
    const long long base = (static_cast<long long>(blockIdx.x) * AM_WARPS + w) * 32;
    if (base >= rows) return;
    const long long my_r = base + lane;
    const bool valid = my_r < rows;
    const int own_l = valid ? owner[my_r] : 0;
    const int oth_l = valid ? other[my_r] : 0;
    const int cnt = static_cast<int>(min(32LL, rows - base));
    double s = 0;
    for (int j0 = 0; j0 < width; j0 += AM_JCH) {
        const int jn = min(AM_JCH, width - j0);
        const int j = j0 + lane;
#pragma unroll 4
        for (int i = 0; i < cnt; i++) {
            const int own = __shfl_sync(0xffffffffu, own_l, i);
            const int oth = __shfl_sync(0xffffffffu, oth_l, i);
            if (lane < jn) {
                const double a = j < k ? other_f[static_cast<size_t>(oth) * other_stride + j] : 1.0;
                P[i * AM_PS + lane] = xmul(a, x[static_cast<size_t>(own) * width + j]);  // :208
            }
        }
        __syncwarp();
        if (valid) {
            const double* mine = P + lane * AM_PS;
            for (int jj = 0; jj < jn; jj++) s = xadd(s, mine[jj]);
        }
        __syncwarp();
    }
    if (valid) y[my_r] = s;
}

// partial[sg*width + j] = sequential sum over the ratings of one (owner, chunk) segment, in
// input order, of t[r] * A[r,j].  One warp per segment; lane l owns unknowns j = l, l+32, ...
template <int JPL>
__global__ void __launch_bounds__(256)
k_als_seg_partial(const int* __restrict__ seg_start, int nseg, const int* __restrict__ grp_idx,
                  const int* __restrict__ other, const double* __restrict__ other_f,
                  const double* __restrict__ t, double* __restrict__ partial, int width,
                  int other_stride, int k, const int* __restrict__ guard) {
    if (guard && *guard) return;
    const int sg = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (sg >= nseg) return;
    const int lane = threadIdx.x & 31;
    const int beg = seg_start[sg], end = seg_start[sg + 1];
    double acc[JPL];
#pragma unroll
    for (int m = 0; m < JPL; m++) acc[m] = 0;
    for (int e0 = beg; e0 < end; e0 += 32) {
        const int e = e0 + lane;
        int oth_l = 0;
        double t_l = 0;
        if (e < end) {
            const int r = grp_idx[e];
            oth_l = other[r];
            t_l = t[r];
        }
        const int cnt = min(32, end - e0);
        for (int i0 = 0; i0 < cnt; i0 += 4) {
            double a[4][JPL];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int oth = __shfl_sync(0xffffffffu, oth_l, (i0 + u) & 31);
#pragma unroll
                for (int m = 0; m < JPL; m++) {
                    const int j = lane + 32 * m;
                    a[u][m] = (i0 + u < cnt && j < k)
                                  ? other_f[static_cast<size_t>(oth) * other_stride + j]
                                  : 1.0;
                }
            }
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const double tv = shfl_double(t_l, (i0 + u) & 31);
                if (i0 + u < cnt) {
#pragma unroll
                    for (int m = 0; m < JPL; m++) acc[m] = xadd(acc[m], xmul(tv, a[u][m]));  // :244
                }
            }
        }
    }
#pragma unroll
    for (int m = 0; m < JPL; m++) {
        const int j = lane + 32 * m;
        if (j < width) partial[static_cast<size_t>(sg) * width + j] = acc[m];
    }
}
}  // namespace

AlsFaithfulOp::AlsFaithfulOp(int num_ratings, int num_owners, const int* d_owner,
                             const int* d_other, const int* d_grp_ptr, const int* d_grp_idx,
                             const double* d_other_f, int width, int other_stride, int k,
                             bool has_one)
    : owners_(num_owners), owner_(d_owner), other_(d_other), grp_ptr_(d_grp_ptr),
      grp_idx_(d_grp_idx), other_f_(d_other_f), width_(width), other_stride_(other_stride),
      k_(k), has_one_(has_one) {
    rows = num_ratings;
    cols = num_owners * width;
    MRB_REQUIRE(width == k + (has_one ? 1 : 0), "AlsFaithfulOp: width/k mismatch");
    MRB_REQUIRE(width <= 256, "ALS rank above 255 is not supported by the reference-order path");
}

void AlsFaithfulOp::mul(const double* d_x, double* d_y, cudaStream_t s) {
    if (rows == 0) return;
    static bool attr_set = false;
    const size_t smem = sizeof(double) * AM_WARPS * 32 * AM_PS;
    if (!attr_set) {
        MRB_CUDA(cudaFuncSetAttribute(k_als_mul, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      static_cast<int>(smem)));
        attr_set = true;
    }
    k_als_mul<<<ceil_div(rows, AM_WARPS * 32), AM_WARPS * 32, smem, s>>>(
        owner_, other_, other_f_, d_x, d_y, rows, width_, other_stride_, k_, guard); MRB_LAUNCHED(1);
}

void AlsFaithfulOp::tmul(const double* d_t, double* d_y, const int* d_row_bounds, int nchunks,
                         cudaStream_t s) {
    if (owners_ == 0) return;
    SegTable& st = seg_[nchunks == 1 ? 1 : 0];
    if (st.bounds_key != d_row_bounds || st.nchunks_key != nchunks)
        build_segments(st, grp_ptr_, grp_idx_, owners_, rows, d_row_bounds, nchunks, width_, s);
    const int grid = ceil_div(static_cast<long long>(st.nseg > 0 ? st.nseg : 1) * 32, 256);
    const int jpl = (width_ + 31) / 32;
#define MRB_TMUL(J)                                                                          \
    k_als_seg_partial<J><<<grid, 256, 0, s>>>(st.seg_start.p, st.nseg, grp_idx_, other_, other_f_, \
                                              d_t, st.partial.p, width_, other_stride_, k_, guard)
    if (st.nseg > 0) {
        switch (jpl) {
            case 1: MRB_TMUL(1); break;
            case 2: MRB_TMUL(2); break;
            case 3: MRB_TMUL(3); break;
            case 4: MRB_TMUL(4); break;
            case 5: MRB_TMUL(5); break;
            case 6: MRB_TMUL(6); break;
            case 7: MRB_TMUL(7); break;
            default: MRB_TMUL(8); break;
        }
    }
#undef MRB_TMUL
    const long long outputs = static_cast<long long>(owners_) * width_;
    k_seg_fold<<<ceil_div(outputs, 256), 256, 0, s>>>(st.grp_seg_ptr.p, st.partial.p, d_y, outputs,
                                                     width_, guard);
    MRB_LAUNCHED(2);
}

}  // namespace mrb
