// K4 -- stable index build (see index_build.cuh).  Integer-only; deterministic by construction
// (the only atomics are integer counters whose final values do not depend on ordering).
#include "index_build.cuh"

namespace mrb {

// ------------------------------------------------------------------------------------------
// Exclusive scan: 4096-element tiles, one 512-thread CTA per tile, recursive over tile sums.
// ------------------------------------------------------------------------------------------
namespace {
constexpr int SCAN_THREADS = 512;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__global__ void __launch_bounds__(SCAN_THREADS)
k_scan_tile(const int* __restrict__ in, int* __restrict__ out, int* __restrict__ tile_sums,
            long long n) {
    __shared__ int warp_sums[SCAN_THREADS / 32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const long long base = static_cast<long long>(blockIdx.x) * SCAN_TILE +
                           static_cast<long long>(threadIdx.x) * SCAN_ITEMS;
    int v[SCAN_ITEMS];
    int run = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++) {
        const long long idx = base + i;
        const int t = idx < n ? in[idx] : 0;
        v[i] = run;
        run += t;
    }
    int inc = run;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, inc, off);
        if (lane >= off) inc += t;
    }
    if (lane == 31) warp_sums[w] = inc;
    __syncthreads();
    if (w == 0) {
        const int orig = lane < SCAN_THREADS / 32 ? warp_sums[lane] : 0;
        int s = orig;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, s, off);
            if (lane >= off) s += t;
        }
        if (lane < SCAN_THREADS / 32) warp_sums[lane] = s - orig;
        if (lane == 31 && tile_sums) tile_sums[blockIdx.x] = s;
    }
    __syncthreads();
    const int thread_off = warp_sums[w] + inc - run;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++) {
        const long long idx = base + i;
        if (idx < n) out[idx] = v[i] + thread_off;
    }
}

__global__ void k_add_tile_offsets(int* __restrict__ out, const int* __restrict__ tile_off,
                                   long long n) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) out[i] += tile_off[i / SCAN_TILE];
}
}  // namespace

void exclusive_scan_i32(const int* d_in, int* d_out, long long n, cudaStream_t s) {
    if (n <= 0) return;
    const int tiles = ceil_div(n, SCAN_TILE);
    if (tiles == 1) {
        k_scan_tile<<<1, SCAN_THREADS, 0, s>>>(d_in, d_out, nullptr, n); MRB_LAUNCHED(1);
        MRB_CUDA(cudaGetLastError());
        return;
    }
    DevBuf<int> sums(tiles);
    k_scan_tile<<<tiles, SCAN_THREADS, 0, s>>>(d_in, d_out, sums.p, n); MRB_LAUNCHED(1);
    MRB_CUDA(cudaGetLastError());
    exclusive_scan_i32(sums.p, sums.p, tiles, s);
    k_add_tile_offsets<<<ceil_div(n, 256), 256, 0, s>>>(d_out, sums.p, n); MRB_LAUNCHED(1);
    MRB_CUDA(cudaGetLastError());
    MRB_CUDA(cudaStreamSynchronize(s));  // `sums` is freed on return
}

// ------------------------------------------------------------------------------------------
// Stable LSD radix sort of (key, position), 8-bit digits.
// Tile = 4096 items per CTA; warp w owns the contiguous 512-item sub-tile, walked in 16 groups
// of 32 (coalesced).  Stability: destination = base[digit][cta] + (items with that digit in
// lower warps of the CTA) + (in earlier groups of the warp) + (in lower lanes of the group).
// ------------------------------------------------------------------------------------------
namespace {
constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_GROUPS = 16;
constexpr int RS_TILE = RS_THREADS * RS_GROUPS;
constexpr unsigned RS_INVALID = 0xFFFFu;

__device__ __forceinline__ long long rs_item(int block, int w, int g, int lane) {
    return static_cast<long long>(block) * RS_TILE + w * (32 * RS_GROUPS) + g * 32 + lane;
}

__global__ void __launch_bounds__(RS_THREADS)
k_radix_hist(const int* __restrict__ keys, int n, int shift, int* __restrict__ block_hist,
             int num_blocks) {
    __shared__ int hist[256];
    hist[threadIdx.x] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int g = 0; g < RS_GROUPS; g++) {
        const long long i = rs_item(blockIdx.x, w, g, lane);
        if (i < n) atomicAdd(&hist[(static_cast<unsigned>(keys[i]) >> shift) & 0xFFu], 1);
    }
    __syncthreads();
    block_hist[static_cast<size_t>(threadIdx.x) * num_blocks + blockIdx.x] = hist[threadIdx.x];
}

// vals_in == nullptr  =>  the value of item i is i (first pass).
__global__ void __launch_bounds__(RS_THREADS)
k_radix_scatter(const int* __restrict__ keys_in, const int* __restrict__ vals_in, int n, int shift,
                const int* __restrict__ block_base, int num_blocks, int* __restrict__ keys_out,
                int* __restrict__ vals_out) {
    __shared__ int warp_cnt[RS_WARPS][256];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const unsigned lt_mask = (1u << lane) - 1u;
    for (int i = threadIdx.x; i < RS_WARPS * 256; i += RS_THREADS) (&warp_cnt[0][0])[i] = 0;
    __syncthreads();

    int key[RS_GROUPS], val[RS_GROUPS];
    unsigned dig[RS_GROUPS];
#pragma unroll
    for (int g = 0; g < RS_GROUPS; g++) {
        const long long i = rs_item(blockIdx.x, w, g, lane);
        if (i < n) {
            key[g] = keys_in[i];
            val[g] = vals_in ? vals_in[i] : static_cast<int>(i);
            dig[g] = (static_cast<unsigned>(key[g]) >> shift) & 0xFFu;
        } else {
            key[g] = 0;
            val[g] = 0;
            dig[g] = RS_INVALID;
        }
    }
    // pass 1: per-warp digit counts
#pragma unroll
    for (int g = 0; g < RS_GROUPS; g++) {
        const unsigned peers = __match_any_sync(0xffffffffu, dig[g]);
        if (dig[g] != RS_INVALID && lane == __ffs(peers) - 1) warp_cnt[w][dig[g]] += __popc(peers);
        __syncwarp();
    }
    __syncthreads();
    // per digit: global base of this CTA + exclusive scan over the CTA's warps
    {
        const int d = threadIdx.x;
        int base = block_base[static_cast<size_t>(d) * num_blocks + blockIdx.x];
#pragma unroll
        for (int ww = 0; ww < RS_WARPS; ww++) {
            const int c = warp_cnt[ww][d];
            warp_cnt[ww][d] = base;
            base += c;
        }
    }
    __syncthreads();
    // pass 2: stable destinations
#pragma unroll
    for (int g = 0; g < RS_GROUPS; g++) {
        const unsigned peers = __match_any_sync(0xffffffffu, dig[g]);
        int pos = 0;
        if (dig[g] != RS_INVALID) pos = warp_cnt[w][dig[g]] + __popc(peers & lt_mask);
        __syncwarp();
        if (dig[g] != RS_INVALID && lane == __ffs(peers) - 1) warp_cnt[w][dig[g]] += __popc(peers);
        __syncwarp();
        if (dig[g] != RS_INVALID) {
            keys_out[pos] = key[g];
            vals_out[pos] = val[g];
        }
    }
}

__global__ void k_count_keys(const int* __restrict__ keys, int n, int* __restrict__ counts) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) atomicAdd(&counts[keys[i]], 1);
}

__global__ void k_iota(int* __restrict__ out, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = i;
}
}  // namespace

void group_pointers(const int* d_key, int n, int num_groups, int* d_ptr, cudaStream_t s) {
    MRB_REQUIRE(n >= 0 && num_groups >= 0, "group_pointers: negative size");
    MRB_CUDA(cudaMemsetAsync(d_ptr, 0, sizeof(int) * (static_cast<size_t>(num_groups) + 1), s));
    if (n == 0) return;
    MRB_REQUIRE(num_groups > 0, "group_pointers: items but no groups");
    k_count_keys<<<ceil_div(n, 256), 256, 0, s>>>(d_key, n, d_ptr); MRB_LAUNCHED(1);
    MRB_CUDA(cudaGetLastError());
    exclusive_scan_i32(d_ptr, d_ptr, static_cast<long long>(num_groups) + 1, s);
}

void stable_group_by(const int* d_key, int n, int num_groups, int* d_ptr, int* d_idx,
                     cudaStream_t s) {
    MRB_REQUIRE(n >= 0 && num_groups >= 0, "stable_group_by: negative size");
    // group pointers: integer histogram + exclusive scan
    MRB_CUDA(cudaMemsetAsync(d_ptr, 0, sizeof(int) * (static_cast<size_t>(num_groups) + 1), s));
    if (n == 0) return;
    MRB_REQUIRE(num_groups > 0, "stable_group_by: items but no groups");
    k_count_keys<<<ceil_div(n, 256), 256, 0, s>>>(d_key, n, d_ptr); MRB_LAUNCHED(1);
    MRB_CUDA(cudaGetLastError());
    exclusive_scan_i32(d_ptr, d_ptr, static_cast<long long>(num_groups) + 1, s);

    int bits = 0;
    while (bits < 31 && (1LL << bits) < num_groups) bits++;
    const int passes = bits == 0 ? 0 : (bits + 7) / 8;
    if (passes == 0) {  // a single group: identity order
        k_iota<<<ceil_div(n, 256), 256, 0, s>>>(d_idx, n); MRB_LAUNCHED(1);
        MRB_CUDA(cudaGetLastError());
        return;
    }
    const int num_blocks = ceil_div(n, RS_TILE);
    DevBuf<int> hist(static_cast<size_t>(256) * num_blocks);
    DevBuf<int> key_a(n), key_b(n), val_b(passes > 1 ? n : 0);
    const int* kin = d_key;
    const int* vin = nullptr;
    // ping-pong so that the LAST pass writes its values into d_idx
    int* kbuf[2] = {key_a.p, key_b.p};
    // values alternate between d_idx and val_b; choose so pass (passes-1) lands in d_idx
    for (int p = 0; p < passes; p++) {
        int* kout = kbuf[p & 1];
        int* vout = ((passes - 1 - p) & 1) ? val_b.p : d_idx;
        k_radix_hist<<<num_blocks, RS_THREADS, 0, s>>>(kin, n, 8 * p, hist.p, num_blocks); MRB_LAUNCHED(1);
        MRB_CUDA(cudaGetLastError());
        exclusive_scan_i32(hist.p, hist.p, static_cast<long long>(256) * num_blocks, s);
        k_radix_scatter<<<num_blocks, RS_THREADS, 0, s>>>(kin, vin, n, 8 * p, hist.p, num_blocks,
                                                           kout, vout); MRB_LAUNCHED(1);
        MRB_CUDA(cudaGetLastError());
        kin = kout;
        vin = vout;
    }
    MRB_CUDA(cudaStreamSynchronize(s));  // temporaries are freed on return
}

void stable_sort_pairs(const int* d_key, const int* d_val_in, int n, unsigned digit_mask,
                       int* d_key_out, int* d_val_out, cudaStream_t s) {
    MRB_REQUIRE(n >= 0, "stable_sort_pairs: negative size");
    if (n == 0) return;
    MRB_REQUIRE(d_key != d_key_out && d_val_in != d_val_out, "stable_sort_pairs: aliased buffers");
    digit_mask &= 0xFu;
    const int passes = __builtin_popcount(digit_mask);
    if (passes == 0) {
        MRB_CUDA(cudaMemcpyAsync(d_key_out, d_key, sizeof(int) * static_cast<size_t>(n),
                                 cudaMemcpyDeviceToDevice, s));
        if (d_val_in) {
            MRB_CUDA(cudaMemcpyAsync(d_val_out, d_val_in, sizeof(int) * static_cast<size_t>(n),
                                     cudaMemcpyDeviceToDevice, s));
        } else {
            k_iota<<<ceil_div(n, 256), 256, 0, s>>>(d_val_out, n); MRB_LAUNCHED(1);
            MRB_CUDA(cudaGetLastError());
        }
        return;
    }
    const int num_blocks = ceil_div(n, RS_TILE);
    DevBuf<int> hist(static_cast<size_t>(256) * num_blocks);
    DevBuf<int> key_t(passes > 1 ? n : 0), val_t(passes > 1 ? n : 0);
    const int* kin = d_key;
    const int* vin = d_val_in;
    int done = 0;
    for (int d = 0; d < 4; d++) {
        if (!((digit_mask >> d) & 1u)) continue;
        // alternate between the temporaries and the outputs so that the LAST pass lands in the outputs
        const bool to_out = ((passes - 1 - done) & 1) == 0;
        int* kout = to_out ? d_key_out : key_t.p;
        int* vout = to_out ? d_val_out : val_t.p;
        k_radix_hist<<<num_blocks, RS_THREADS, 0, s>>>(kin, n, 8 * d, hist.p, num_blocks); MRB_LAUNCHED(1);
        MRB_CUDA(cudaGetLastError());
        exclusive_scan_i32(hist.p, hist.p, static_cast<long long>(256) * num_blocks, s);
        k_radix_scatter<<<num_blocks, RS_THREADS, 0, s>>>(kin, vin, n, 8 * d, hist.p, num_blocks,
                                                           kout, vout); MRB_LAUNCHED(1);
        MRB_CUDA(cudaGetLastError());
        kin = kout;
        vin = vout;
        done++;
    }
    MRB_CUDA(cudaStreamSynchronize(s));  // temporaries are freed on return
}

// ------------------------------------------------------------------------------------------
// Explicit transpose.
// ------------------------------------------------------------------------------------------
namespace {
__global__ void k_expand_rows(const int* __restrict__ rowptr, int rows, int* __restrict__ row_of) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    for (int e = rowptr[r]; e < rowptr[r + 1]; e++) row_of[e] = r;
}

__global__ void k_gather_transposed(const int* __restrict__ idx, const int* __restrict__ row_of,
                                    const double* __restrict__ vals, int nnz,
                                    int* __restrict__ t_row, double* __restrict__ t_val) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nnz) return;
    const int e = idx[i];
    t_row[i] = row_of[e];
    t_val[i] = vals[e];
}
}  // namespace

void csr_transpose(int rows, int cols, int nnz, const int* d_rowptr, const int* d_colidx,
                   const double* d_vals, int* d_t_ptr, int* d_t_row, double* d_t_val,
                   cudaStream_t s) {
    if (nnz == 0) {
        MRB_CUDA(cudaMemsetAsync(d_t_ptr, 0, sizeof(int) * (static_cast<size_t>(cols) + 1), s));
        return;
    }
    DevBuf<int> idx(nnz), row_of(nnz);
    stable_group_by(d_colidx, nnz, cols, d_t_ptr, idx.p, s);
    k_expand_rows<<<ceil_div(rows, 256), 256, 0, s>>>(d_rowptr, rows, row_of.p); MRB_LAUNCHED(1);
    MRB_CUDA(cudaGetLastError());
    k_gather_transposed<<<ceil_div(nnz, 256), 256, 0, s>>>(idx.p, row_of.p, d_vals, nnz, d_t_row,
                                                          d_t_val); MRB_LAUNCHED(1);
    MRB_CUDA(cudaGetLastError());
    MRB_CUDA(cudaStreamSynchronize(s));
}

}  // namespace mrb
