// K3 -- the generic sparse least-squares solver (cpp_ls.cg_least_squares,
// python/full_data/cpp_ls.py:47-111 -> cg_least_squares, cpp/ls_lib/matrix.cpp:456-529) with
// GPU-native summation: the "linear model" path of config 2 (user + movie bias model, 2 non-zeros
// per row).  Same algorithm and stopping rule as the reference; per CG iteration
//   t  = A p      CSR, one thread per short row / one warp per long row
//   Ap = A^T t    through the stable transpose built once by K4 (CSC): one lane per entry in
//                 32-entry windows (k_csc_flat), then an ordered fold per column fused with
//                 the per-column partial of p . Ap
//   x, r, p updates fused with the partial sums of r . r
// All HBM streaming; algorithmic bytes per iteration (nnz_A = non-zeros):
//   2*nnz_A*(8+4) + (rows+cols+2)*4 + rows*8 + nnz_A*8 + 7*cols*8.
#include "ls_native.cuh"
#include "prep.cuh"

#include <cstdlib>
#include <string>
#include <vector>

#include "faithful_cg.cuh"
#include "index_build.cuh"
#include "native_cg.cuh"

namespace mrb {

namespace {

// UNIT: every stored value is exactly 1.0 (the indicator matrices of the bias model, config 2):
// found once at upload (k_all_ones), the value streams are then not read at all -- 1.0 * x is x
// bit for bit, so the results do not change.
// One thread per FOUR short rows (r, r + T, r + 2T, r + 3T with T = threads of the grid, so every
// load instruction stays coalesced): the rowptr -> column index -> gather chain of one row is
// three dependent memory round trips; with a single row per thread the kernel ran at the latency
// of that chain (ncu: 2.6 TB/s of DRAM, 83 % of the stall samples on the long scoreboard,
// profiles/ncu_k_ls_native_C2_r02.txt).  The order of the additions inside a row is unchanged.
constexpr int CSR_ROWS_PER_THREAD = 4;
template <bool UNIT>
__global__ void __launch_bounds__(256)
k_csr_mul_thread(const int* __restrict__ rowptr, const int* __restrict__ colidx,
                 const double* __restrict__ vals, const double* __restrict__ x,
                 double* __restrict__ y, int rows, const CgState* __restrict__ guard) {
    if (guard && guard->done) return;
    const int T = gridDim.x * blockDim.x;
    const int r0 = blockIdx.x * blockDim.x + threadIdx.x;
    int beg[CSR_ROWS_PER_THREAD], end[CSR_ROWS_PER_THREAD], maxlen = 0;
    double s[CSR_ROWS_PER_THREAD];
#pragma unroll
    for (int j = 0; j < CSR_ROWS_PER_THREAD; j++) {
        const int r = r0 + j * T;
        beg[j] = r < rows ? rowptr[r] : 0;
        end[j] = r < rows ? rowptr[r + 1] : 0;
        s[j] = 0;
    }
#pragma unroll
    for (int j = 0; j < CSR_ROWS_PER_THREAD; j++) maxlen = max(maxlen, end[j] - beg[j]);
    for (int o = 0; o < maxlen; o++) {
        int c[CSR_ROWS_PER_THREAD];
        double v[CSR_ROWS_PER_THREAD];
#pragma unroll
        for (int j = 0; j < CSR_ROWS_PER_THREAD; j++) {
            const bool in = beg[j] + o < end[j];
            c[j] = in ? colidx[beg[j] + o] : -1;
            v[j] = in && !UNIT ? vals[beg[j] + o] : 1.0;
        }
#pragma unroll
        for (int j = 0; j < CSR_ROWS_PER_THREAD; j++)
            if (c[j] >= 0) s[j] += UNIT ? x[c[j]] : v[j] * x[c[j]];
    }
#pragma unroll
    for (int j = 0; j < CSR_ROWS_PER_THREAD; j++) {
        const int r = r0 + j * T;
        if (r < rows) y[r] = s[j];
    }
}

__global__ void k_all_ones(const double* __restrict__ vals, int n, int* __restrict__ not_one) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && vals[i] != 1.0) *not_one = 1;
}

template <bool UNIT>
__global__ void __launch_bounds__(256)
k_csr_mul_warp(const int* __restrict__ rowptr, const int* __restrict__ colidx,
               const double* __restrict__ vals, const double* __restrict__ x,
               double* __restrict__ y, int rows, const CgState* __restrict__ guard) {
    if (guard && guard->done) return;
    const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (r >= rows) return;
    const int lane = threadIdx.x & 31;
    double s = 0;
    const int end = rowptr[r + 1];
    for (int e = rowptr[r] + lane; e < end; e += 32) s += UNIT ? x[colidx[e]] : vals[e] * x[colidx[e]];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    if (lane == 0) y[r] = s;
}

// G lanes per (column, row-block) segment (G = 8, 16 or 32, chosen from the mean segment length):
// the lanes stride over the segment, fixed-order xor reduction inside the group.  Splitting the
// rows into 64 equal blocks bounds the longest segment, so that a movie column with 55 000
// entries is spread over 64 groups instead of serialising one.  With a full warp per segment the
// kernel ran at the latency of its dependent chain (bounds -> row/value -> gather of t) with 15
// useful elements per warp at config 2 (ncu: 21 % of the HBM peak at 80 % occupancy, issue slots
// 30 % busy); 8-lane groups put four segments in flight per warp.
template <int G>
__global__ void __launch_bounds__(256)
k_csc_seg_native(const int* __restrict__ seg_start, int nseg, const int* __restrict__ t_row,
                 const double* __restrict__ t_val, const double* __restrict__ t,
                 double* __restrict__ partial, const CgState* __restrict__ guard) {
    if (guard && guard->done) return;
    const long long gid = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) / G;
    const int sub = threadIdx.x & (G - 1);
    // every lane of the warp stays in the shuffles below; groups past the end carry empty ranges
    const bool live = gid < nseg;
    const int beg = live ? seg_start[gid] : 0, end = live ? seg_start[gid + 1] : 0;
    double s0 = 0, s1 = 0;
    int e = beg + sub;
    for (; e + G < end; e += 2 * G) {
        s0 += t_val[e] * t[t_row[e]];
        s1 += t_val[e + G] * t[t_row[e + G]];
    }
    if (e < end) s0 += t_val[e] * t[t_row[e]];
    double s = s0 + s1;
#pragma unroll
    for (int off = G / 2; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    if (sub == 0 && live) partial[gid] = s;
}

// ---- flat A^T t: the transposed entries are cut into WINDOWS of 32 consecutive entries; a piece
// is (column intersected with a window).  heads[w] has bit b set where entry 32 w + b starts a
// piece (bit 0 always), win_first[w] = number of pieces before window w.  One lane per entry:
// perfectly coalesced index / value streams, no per-segment dependent chain, the same number of
// entries per lane whatever the column lengths; the pieces of a column are then added in order
// by k_csc_fold_native.  Deterministic (fixed shuffle tree inside a window, ordered fold).
__global__ void k_flat_init_heads(unsigned* __restrict__ heads, int nwords) {
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w <= nwords) heads[w] = w < nwords ? 1u : 0u;
}

__global__ void k_flat_mark_heads(const int* __restrict__ t_ptr, int cols, unsigned* __restrict__ heads) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= cols) return;
    const int e = t_ptr[c];
    if (e < t_ptr[c + 1]) atomicOr(&heads[e >> 5], 1u << (e & 31));
}

__global__ void k_flat_popc(const unsigned* __restrict__ heads, int nwords, int* __restrict__ cnt) {
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w <= nwords) cnt[w] = __popc(heads[w]);
}

// col_piece_ptr[c] = number of pieces that start before entry t_ptr[c]  (c = 0 .. cols)
__global__ void k_flat_col_ptr(const int* __restrict__ t_ptr, int cols,
                               const unsigned* __restrict__ heads, const int* __restrict__ win_first,
                               int* __restrict__ col_piece_ptr) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c > cols) return;
    const int e = t_ptr[c], w = e >> 5, b = e & 31;
    col_piece_ptr[c] = win_first[w] + __popc(heads[w] & ((1u << b) - 1u));
}

template <int WPW, bool UNIT>   // windows per warp: all index/value loads, then all gathers, then the scans
__global__ void __launch_bounds__(256)
k_csc_flat(const unsigned* __restrict__ heads, const int* __restrict__ win_first,
           const int* __restrict__ t_row, const double* __restrict__ t_val,
           const double* __restrict__ t, double* __restrict__ partial, int nnz, int nwords,
           const CgState* __restrict__ guard) {
    if (guard && guard->done) return;
    const int lane = threadIdx.x & 31;
    const long long w0 = ((static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5) * WPW;
    int row[WPW], first[WPW];
    double val[WPW];
    unsigned hd[WPW];
#pragma unroll
    for (int j = 0; j < WPW; j++) {
        const long long w = w0 + j, e = w * 32 + lane;
        const bool win = w < nwords, ok = win && e < nnz;
        row[j] = ok ? t_row[e] : -1;
        val[j] = ok && !UNIT ? t_val[e] : 0.0;
        hd[j] = win ? heads[w] : 0u;
        first[j] = win ? win_first[w] : 0;
    }
    double prod[WPW];
#pragma unroll
    for (int j = 0; j < WPW; j++) prod[j] = row[j] >= 0 ? (UNIT ? t[row[j]] : val[j] * t[row[j]]) : 0.0;
#pragma unroll
    for (int j = 0; j < WPW; j++) {
        const unsigned below = hd[j] & (0xffffffffu >> (31 - lane));   // heads at or before this lane
        const int head = 31 - __clz(below);                             // -1 for a window past the end
        double v = prod[j];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const double o = __shfl_up_sync(0xffffffffu, v, d);
            if (lane - d >= head) v += o;
        }
        const bool last = lane == 31 || ((hd[j] >> (lane + 1)) & 1u);
        if (hd[j] != 0u && last) partial[first[j] + __popc(below) - 1] = v;
    }
}

// out[c] = sum of the column's segment sums in block order; dots[c] = v[c] * out[c]
__global__ void __launch_bounds__(256)
k_csc_fold_native(const int* __restrict__ grp_seg_ptr, const double* __restrict__ partial,
                  const double* __restrict__ v, double* __restrict__ out, double* __restrict__ dots,
                  int cols, const CgState* __restrict__ guard) {
    if (guard && guard->done) return;
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= cols) return;
    double s = 0;
    const int se = grp_seg_ptr[c + 1];
    for (int sg = grp_seg_ptr[c]; sg < se; sg++) s += partial[sg];
    out[c] = s;
    if (dots) dots[c] = v ? v[c] * s : 0.0;
}

// The same fold with G lanes per column (flat path: a popular movie column has up to 1 700 pieces,
// a user column about four): lanes stride over the column's pieces, fixed-order xor reduction.
template <int G>
__global__ void __launch_bounds__(256)
k_csc_fold_group(const int* __restrict__ col_piece_ptr, const double* __restrict__ piece_sum,
                 const double* __restrict__ v, double* __restrict__ out, double* __restrict__ dots,
                 int cols, const CgState* __restrict__ guard) {
    if (guard && guard->done) return;
    const long long c = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) / G;
    const int sub = threadIdx.x & (G - 1);
    const bool live = c < cols;
    const int beg = live ? col_piece_ptr[c] : 0, end = live ? col_piece_ptr[c + 1] : 0;
    double s = 0;
    for (int i = beg + sub; i < end; i += G) s += piece_sum[i];
#pragma unroll
    for (int off = G / 2; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    if (live && sub == 0) {
        out[c] = s;
        if (dots) dots[c] = v ? v[c] * s : 0.0;
    }
}

}  // namespace

LsNativeResult solve_ls_native(int rows, int cols, const int* rowptr, const int* colidx,
                               const double* vals, const double* b, double* x,
                               double min_r_decrease, int max_iteration) {
    MRB_REQUIRE(rowptr[0] == 0 && rowptr[rows] >= 0, "cg_least_squares: bad row pointers");
    const int nnz = rowptr[rows];
    cudaStream_t s;
    MRB_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    struct StreamGuard { cudaStream_t s; ~StreamGuard() { cudaStreamDestroy(s); } } sg{s};
    DevBuf<int> d_rowptr(static_cast<size_t>(rows) + 1), d_col(nnz);
    DevBuf<double> d_vals(nnz), d_b(rows), d_x(cols);
    d_rowptr.upload(rowptr, static_cast<size_t>(rows) + 1, s);
    d_col.upload(colidx, nnz, s);
    d_vals.upload(vals, nnz, s);
    d_b.upload(b, rows, s);
    d_x.upload(x, cols, s);
    check_csr(d_rowptr.p, rows, d_col.p, nnz, cols, "cg_least_squares", s);

    cudaEvent_t e0, e1, e2;
    MRB_CUDA(cudaEventCreate(&e0));
    MRB_CUDA(cudaEventCreate(&e1));
    MRB_CUDA(cudaEventCreate(&e2));
    MRB_CUDA(cudaEventRecord(e0, s));
    DevBuf<int> t_ptr(static_cast<size_t>(cols) + 1), t_row(nnz);
    DevBuf<double> t_val(nnz);
    csr_transpose(rows, cols, nnz, d_rowptr.p, d_col.p, d_vals.p, t_ptr.p, t_row.p, t_val.p, s);
    MRB_CUDA(cudaEventRecord(e1, s));

    const size_t len = std::max(cols, 1);
    DevBuf<double> g(len), r(len), p(len), Ap(len), dots(len), tmp(std::max(rows, 1));
    DevBuf<double> partials(static_cast<size_t>(ceil_div(static_cast<long long>(len), 256)) + 1);
    DevBuf<CgState> state(1);
    const bool long_rows = rows > 0 && nnz / rows > 8;
    // indicator matrix?  (MRB_LS_NO_UNIT=1 keeps the general kernels for A/B runs)
    bool unit = false;
    if (nnz > 0 && std::getenv("MRB_LS_NO_UNIT") == nullptr) {
        DevBuf<int> not_one(1);
        MRB_CUDA(cudaMemsetAsync(not_one.p, 0, sizeof(int), s));
        k_all_ones<<<ceil_div(nnz, 256), 256, 0, s>>>(d_vals.p, nnz, not_one.p); MRB_LAUNCHED(1);
        int h = 1;
        MRB_CUDA(cudaMemcpyAsync(&h, not_one.p, sizeof(int), cudaMemcpyDeviceToHost, s));
        MRB_CUDA(cudaStreamSynchronize(s));
        unit = h == 0;
    }
    auto mul = [&](const double* v, double* out, const CgState* guard) {
        if (rows == 0) return;
        const int gw = ceil_div(static_cast<long long>(rows) * 32, 256);
        const int gt = ceil_div(ceil_div(rows, CSR_ROWS_PER_THREAD), 256);
        if (long_rows && unit)
            k_csr_mul_warp<true><<<gw, 256, 0, s>>>(d_rowptr.p, d_col.p, d_vals.p, v, out, rows, guard);
        else if (long_rows)
            k_csr_mul_warp<false><<<gw, 256, 0, s>>>(d_rowptr.p, d_col.p, d_vals.p, v, out, rows, guard);
        else if (unit)
            k_csr_mul_thread<true><<<gt, 256, 0, s>>>(d_rowptr.p, d_col.p, d_vals.p, v, out, rows, guard);
        else
            k_csr_mul_thread<false><<<gt, 256, 0, s>>>(d_rowptr.p, d_col.p, d_vals.p, v, out, rows, guard);
        MRB_LAUNCHED(1);
    };
    // row blocks for the segment split (any fixed table gives a deterministic summation order)
    constexpr int kRowBlocks = 64;
    std::vector<int> h_bounds(kRowBlocks + 1);
    for (int i = 0; i <= kRowBlocks; i++) h_bounds[i] = static_cast<int>(static_cast<long long>(rows) * i / kRowBlocks);
    DevBuf<int> d_bounds(kRowBlocks + 1);
    d_bounds.upload(h_bounds.data(), kRowBlocks + 1, s);
    MRB_CUDA(cudaStreamSynchronize(s));
    SegTable seg;
    // flat window table (default); MRB_LS_TMUL=seg selects the per-segment groups instead
    const bool use_flat = !(std::getenv("MRB_LS_TMUL") && std::string(std::getenv("MRB_LS_TMUL")) == "seg");
    const int nwords = ceil_div(nnz, 32);
    DevBuf<unsigned> heads(static_cast<size_t>(nwords) + 1);
    DevBuf<int> win_first(static_cast<size_t>(nwords) + 1), col_piece_ptr(static_cast<size_t>(cols) + 1);
    DevBuf<double> piece_sum;
    int npieces = 0;
    if (use_flat) {
        k_flat_init_heads<<<ceil_div(nwords + 1, 256), 256, 0, s>>>(heads.p, nwords);
        if (cols > 0) k_flat_mark_heads<<<ceil_div(cols, 256), 256, 0, s>>>(t_ptr.p, cols, heads.p);
        k_flat_popc<<<ceil_div(nwords + 1, 256), 256, 0, s>>>(heads.p, nwords, win_first.p);
        MRB_LAUNCHED(3);
        MRB_CUDA(cudaGetLastError());
        exclusive_scan_i32(win_first.p, win_first.p, static_cast<long long>(nwords) + 1, s);
        k_flat_col_ptr<<<ceil_div(cols + 1, 256), 256, 0, s>>>(t_ptr.p, cols, heads.p, win_first.p,
                                                            col_piece_ptr.p);
        MRB_LAUNCHED(1);
        MRB_CUDA(cudaGetLastError());
        MRB_CUDA(cudaMemcpyAsync(&npieces, win_first.p + nwords, sizeof(int), cudaMemcpyDeviceToHost, s));
        MRB_CUDA(cudaStreamSynchronize(s));
        piece_sum.alloc(static_cast<size_t>(std::max(npieces, 1)));
    }
    if (!use_flat) build_segments(seg, t_ptr.p, t_row.p, cols, nnz, d_bounds.p, kRowBlocks, 1, s);
    constexpr int kFlatWpw = 4;
    auto tmul = [&](const double* t, const double* v, double* out, double* dd, const CgState* guard) {
        if (cols == 0) return;
        if (use_flat) {
            const int gf = ceil_div(static_cast<long long>(ceil_div(nwords, kFlatWpw)) * 32, 256);
            if (nwords > 0 && unit)
                k_csc_flat<kFlatWpw, true><<<gf, 256, 0, s>>>(heads.p, win_first.p, t_row.p, t_val.p, t,
                                                             piece_sum.p, nnz, nwords, guard);
            else if (nwords > 0)
                k_csc_flat<kFlatWpw, false><<<gf, 256, 0, s>>>(heads.p, win_first.p, t_row.p, t_val.p, t,
                                                              piece_sum.p, nnz, nwords, guard);
            k_csc_fold_group<8><<<ceil_div(static_cast<long long>(cols) * 8, 256), 256, 0, s>>>(
                col_piece_ptr.p, piece_sum.p, v, out, dd, cols, guard);
            MRB_LAUNCHED(2);
            return;
        }
        if (seg.nseg > 0) {
            const long long mean_len = nnz / seg.nseg;
            if (mean_len < 24)
                k_csc_seg_native<8><<<ceil_div(static_cast<long long>(seg.nseg) * 8, 256), 256, 0, s>>>(
                    seg.seg_start.p, seg.nseg, t_row.p, t_val.p, t, seg.partial.p, guard);
            else if (mean_len < 48)
                k_csc_seg_native<16><<<ceil_div(static_cast<long long>(seg.nseg) * 16, 256), 256, 0, s>>>(
                    seg.seg_start.p, seg.nseg, t_row.p, t_val.p, t, seg.partial.p, guard);
            else
                k_csc_seg_native<32><<<ceil_div(static_cast<long long>(seg.nseg) * 32, 256), 256, 0, s>>>(
                    seg.seg_start.p, seg.nseg, t_row.p, t_val.p, t, seg.partial.p, guard);
        }
        k_csc_fold_native<<<ceil_div(cols, 256), 256, 0, s>>>(seg.grp_seg_ptr.p, seg.partial.p, v, out,
                                                            dd, cols, guard);
        MRB_LAUNCHED(2);
    };
    tmul(d_b.p, nullptr, g.p, nullptr, nullptr);                    // g = A^T b   (matrix.cpp:465)
    auto apply = [&](const double* v, double* out, const CgState* guard) {
        mul(v, tmp.p, guard);
        tmul(tmp.p, v, out, dots.p, guard);
    };
    NativeCgWorkspace ws{r.p, p.p, Ap.p, dots.p, partials.p, state.p};
    CgResult cr = native_cg_solve(apply, g.p, d_x.p, cols, cols, min_r_decrease, max_iteration, ws, s);
    MRB_CUDA(cudaEventRecord(e2, s));
    d_x.download(x, cols, s);
    MRB_CUDA(cudaStreamSynchronize(s));
    LsNativeResult res;
    res.iterations = cr.iterations;
    res.final_rr = cr.final_rr;
    MRB_CUDA(cudaEventElapsedTime(&res.transpose_ms, e0, e1));
    MRB_CUDA(cudaEventElapsedTime(&res.solve_ms, e1, e2));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaEventDestroy(e2);
    return res;
}

}  // namespace mrb
