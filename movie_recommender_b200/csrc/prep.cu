// K7 -- ALS data preparation (see prep.cuh).  HBM-bound integer work: every pass streams the
// COO arrays once, coalesced; the only atomics are integer counters (order-independent), so the
// results are deterministic and bit-exact against the reference's Python.
#include "prep.cuh"

#include <algorithm>

#include "index_build.cuh"

namespace mrb {
namespace {

constexpr int PT = 256;  // threads per CTA of the streaming kernels

// ---------------------------------------------------------------------------------- id check
__global__ void __launch_bounds__(PT)
k_check_range(const int* __restrict__ id, int n, int slots, int* __restrict__ bad, int lowest = 0) {
    const int i = blockIdx.x * PT + threadIdx.x;
    if (i < n) {
        const int v = id[i];
        if (v < lowest || v >= slots) *bad = 1;
    }
}

// ---------------------------------------------------------------------------------- medians
// Order-preserving map double -> uint64 (negative values: all bits flipped; others: sign set).
__device__ __forceinline__ unsigned long long orderable(double v) {
    const unsigned long long u = static_cast<unsigned long long>(__double_as_longlong(v));
    return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}

// key_lo / key_hi = the two halves of the orderable key; orand[0..3] = OR(lo), AND(lo), OR(hi),
// AND(hi) over all ratings (which 8-bit digits vary at all decides which radix passes run:
// ratings on a 0.5 grid have constant low words, so 4-6 of the 8 passes disappear).
__global__ void __launch_bounds__(PT)
k_rating_keys(const double* __restrict__ rating, int n, int* __restrict__ key_lo,
              int* __restrict__ key_hi, unsigned* __restrict__ orand) {
    // persistent grid-stride CTAs (a multiple of the SM count): the OR/AND words are combined in
    // registers and reach the four global words once per warp of the whole grid.  (One atomic
    // -- or even one look -- per 32 ratings made 867 k warps queue on a single L2 line: 2.4 ms
    // for a kernel that streams 444 MB.)
    unsigned lo_or = 0u, lo_and = 0xFFFFFFFFu, hi_or = 0u, hi_and = 0xFFFFFFFFu;
    const long long stride = static_cast<long long>(gridDim.x) * PT;
    for (long long i = static_cast<long long>(blockIdx.x) * PT + threadIdx.x; i < n; i += stride) {
        const unsigned long long k = orderable(rating[i]);
        const unsigned lo = static_cast<unsigned>(k), hi = static_cast<unsigned>(k >> 32);
        key_lo[i] = static_cast<int>(lo);
        key_hi[i] = static_cast<int>(hi);
        lo_or |= lo;
        lo_and &= lo;
        hi_or |= hi;
        hi_and &= hi;
    }
    lo_or = __reduce_or_sync(0xffffffffu, lo_or);
    lo_and = __reduce_and_sync(0xffffffffu, lo_and);
    hi_or = __reduce_or_sync(0xffffffffu, hi_or);
    hi_and = __reduce_and_sync(0xffffffffu, hi_and);
    if ((threadIdx.x & 31) == 0) {
        atomicOr(&orand[0], lo_or);
        atomicAnd(&orand[1], lo_and);
        atomicOr(&orand[2], hi_or);
        atomicAnd(&orand[3], hi_and);
    }
}

__global__ void __launch_bounds__(PT)
k_gather_i32(const int* __restrict__ src, const int* __restrict__ idx, int n, int* __restrict__ out) {
    const int i = blockIdx.x * PT + threadIdx.x;
    if (i < n) out[i] = src[idx[i]];
}

__global__ void __launch_bounds__(PT)
k_count(const int* __restrict__ key, int n, int* __restrict__ cnt) {
    const int i = blockIdx.x * PT + threadIdx.x;
    const unsigned act = __ballot_sync(0xffffffffu, i < n);
    if (i >= n) return;
    const int k = key[i];
    const unsigned peers = __match_any_sync(act, k);
    if ((threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(&cnt[k], __popc(peers));
}

// order[] lists the rating positions sorted by (movie, rating, position); ptr = segment starts.
// numpy.median: the middle element for an odd count, else mean() of the two middle elements,
// i.e. (a + b) / 2 with both operations rounded (movie_lens_data_proc.py:455-471).
__global__ void __launch_bounds__(PT)
k_median(const int* __restrict__ ptr, const int* __restrict__ order,
         const double* __restrict__ rating, int movie_slots, double* __restrict__ median,
         int* __restrict__ count) {
    const int m = blockIdx.x * PT + threadIdx.x;
    if (m >= movie_slots) return;
    const int b = ptr[m], c = ptr[m + 1] - b;
    count[m] = c;
    if (c == 0) {
        median[m] = __longlong_as_double(0x7ff8000000000000ll);  // NaN: movie without ratings
        return;
    }
    const double hi = rating[order[b + c / 2]];
    if (c & 1) {
        median[m] = hi;
    } else {
        const double lo = rating[order[b + c / 2 - 1]];
        median[m] = __ddiv_rn(__dadd_rn(lo, hi), 2.0);
    }
}

unsigned digits_that_vary(unsigned or_bits, unsigned and_bits) {
    const unsigned varying = or_bits ^ and_bits;
    unsigned mask = 0;
    for (int d = 0; d < 4; d++)
        if ((varying >> (8 * d)) & 0xFFu) mask |= 1u << d;
    return mask;
}

// ---------------------------------------------------------------------------------- shrink
// cnt[key] += 1 for every rating whose user AND movie are still in (warp-aggregated: ratings
// are grouped by user, so a warp usually carries one or two users).  Each thread carries CA_ILP
// ratings (PT apart, so every load stays coalesced) and issues all its id loads, then all its
// flag gathers, before the first vote: with one rating per thread the kernel ran at the latency
// of the dependent chain id -> flag -> atomic (1.98 TB/s of the id streams at 80 % occupancy).
constexpr int CA_ILP = 4;
__global__ void __launch_bounds__(PT)
k_count_alive(const int* __restrict__ user, const int* __restrict__ movie, int n,
              const int* __restrict__ user_ok, const int* __restrict__ movie_ok, int by_user,
              int* __restrict__ cnt) {
    const long long base = static_cast<long long>(blockIdx.x) * (PT * CA_ILP) + threadIdx.x;
    int u[CA_ILP], m[CA_ILP];
    bool alive[CA_ILP];
#pragma unroll
    for (int j = 0; j < CA_ILP; j++) {
        const long long i = base + j * PT;
        u[j] = i < n ? user[i] : -1;
        m[j] = i < n ? movie[i] : -1;
    }
#pragma unroll
    for (int j = 0; j < CA_ILP; j++)
        alive[j] = u[j] >= 0 && user_ok[u[j]] != 0 && movie_ok[m[j]] != 0;
#pragma unroll
    for (int j = 0; j < CA_ILP; j++) {
        const int key = by_user ? u[j] : m[j];
        const unsigned act = __ballot_sync(0xffffffffu, alive[j]);
        if (alive[j]) {
            const unsigned peers = __match_any_sync(act, key);
            if ((threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(&cnt[key], __popc(peers));
        }
    }
}

// ok[s] &= cnt[s] >= min_count.  *changed is raised the way the reference raises has_changed:
// a still-listed user that falls short (empty lists included, _drop_users :494-535); a movie
// that was COUNTED (cnt > 0) and falls short (_count_movies / uncommon_movies,
// movie_lens_data.py:574-588).
__global__ void __launch_bounds__(PT)
k_update_ok(const int* __restrict__ cnt, int* __restrict__ ok, int slots, int min_count,
            int need_positive, int* __restrict__ changed) {
    const int s = blockIdx.x * PT + threadIdx.x;
    if (s >= slots) return;
    if (!ok[s]) return;
    const int c = cnt[s];
    if (c < min_count) {
        ok[s] = 0;
        if (!need_positive || c > 0) *changed = 1;
    }
}

__global__ void __launch_bounds__(PT)
k_fill_i32(int* __restrict__ p, int n, int v) {
    const int i = blockIdx.x * PT + threadIdx.x;
    if (i < n) p[i] = v;
}

// new_id[s] = ok[s] ? (number of ok slots below s) : -1
__global__ void __launch_bounds__(PT)
k_new_ids(const int* __restrict__ ok, const int* __restrict__ scan, int slots,
          int* __restrict__ new_id) {
    const int s = blockIdx.x * PT + threadIdx.x;
    if (s < slots) new_id[s] = ok[s] ? scan[s] : -1;
}

__global__ void __launch_bounds__(PT)
k_alive_flags(const int* __restrict__ user, const int* __restrict__ movie, int n,
              const int* __restrict__ user_ok, const int* __restrict__ movie_ok,
              int* __restrict__ flag /* n + 1 */) {
    const int i = blockIdx.x * PT + threadIdx.x;
    if (i < n) flag[i] = (user_ok[user[i]] != 0 && movie_ok[movie[i]] != 0) ? 1 : 0;
    else if (i == n) flag[i] = 0;
}

// Stable compaction (_convert_training_data_to_numpy, movie_lens_data_proc.py:611-654):
// surviving ratings keep their order; rating - median[movie] is one rounded subtraction.
__global__ void __launch_bounds__(PT)
k_compact(const int* __restrict__ user, const int* __restrict__ movie,
          const double* __restrict__ rating, int n, const int* __restrict__ flag,
          const int* __restrict__ pos, const int* __restrict__ user_new,
          const int* __restrict__ movie_new, const double* __restrict__ median,
          int* __restrict__ out_user, int* __restrict__ out_movie, double* __restrict__ out_rating,
          int* __restrict__ keep_pos) {
    const int i = blockIdx.x * PT + threadIdx.x;
    if (i >= n || !flag[i]) return;
    const int p = pos[i], u = user[i], m = movie[i];
    out_user[p] = user_new[u];
    out_movie[p] = movie_new[m];
    out_rating[p] = __dsub_rn(rating[i], median[m]);
    keep_pos[p] = i;
}

int read_int(const int* d_p, cudaStream_t s) {
    int v = 0;
    MRB_CUDA(cudaMemcpyAsync(&v, d_p, sizeof(int), cudaMemcpyDeviceToHost, s));
    MRB_CUDA(cudaStreamSynchronize(s));
    return v;
}

}  // namespace

void check_id_range(const int* d_id, int n, int slots, const char* what, cudaStream_t s) {
    if (n <= 0) return;
    DevBuf<int> bad(1);
    MRB_CUDA(cudaMemsetAsync(bad.p, 0, sizeof(int), s));
    k_check_range<<<ceil_div(n, PT), PT, 0, s>>>(d_id, n, slots, bad.p); MRB_LAUNCHED(1);
    MRB_CUDA(cudaGetLastError());
    if (read_int(bad.p, s) != 0) throw Error(kErrArgument, std::string(what) + ": id out of range");
}

namespace {
__global__ void k_check_rowptr(const int* __restrict__ ptr, int rows, int* __restrict__ bad) {
    const int i = blockIdx.x * PT + threadIdx.x;
    if (i < rows && ptr[i + 1] < ptr[i]) *bad = 1;
    if (i == 0 && ptr[0] != 0) *bad = 1;
}
}  // namespace

// A CSR operator handed over by a caller: rowptr[0] == 0, rowptr non-decreasing, every column
// index inside [0, cols).  One bad index would otherwise become an out-of-bounds device access
// and a sticky context error for every later call of the process.
void check_csr(const int* d_rowptr, int rows, const int* d_colidx, int nnz, int cols, const char* what,
               cudaStream_t s) {
    DevBuf<int> bad(1);
    MRB_CUDA(cudaMemsetAsync(bad.p, 0, sizeof(int), s));
    if (rows > 0) { k_check_rowptr<<<ceil_div(rows, PT), PT, 0, s>>>(d_rowptr, rows, bad.p); MRB_LAUNCHED(1); }
    if (nnz > 0) { k_check_range<<<ceil_div(nnz, PT), PT, 0, s>>>(d_colidx, nnz, cols, bad.p); MRB_LAUNCHED(1); }
    MRB_CUDA(cudaGetLastError());
    if (read_int(bad.p, s) != 0)
        throw Error(kErrArgument, std::string(what) + ": row pointers must start at 0 and not decrease, "
                                                      "column indices must lie in [0, columns)");
}

void check_id_range_allow_minus1(const int* d_id, int n, int limit, const char* what, cudaStream_t s) {
    if (n <= 0) return;
    DevBuf<int> bad(1);
    MRB_CUDA(cudaMemsetAsync(bad.p, 0, sizeof(int), s));
    k_check_range<<<ceil_div(n, PT), PT, 0, s>>>(d_id, n, limit, bad.p, -1); MRB_LAUNCHED(1);
    MRB_CUDA(cudaGetLastError());
    if (read_int(bad.p, s) != 0) throw Error(kErrArgument, std::string(what) + ": id out of range");
}

void movie_medians(const int* d_movie, const double* d_rating, int n, int movie_slots,
                   double* d_median, int* d_count, cudaStream_t s) {
    MRB_REQUIRE(n >= 0 && movie_slots >= 0, "movie_medians: negative size");
    if (movie_slots == 0) {
        MRB_REQUIRE(n == 0, "movie_medians: ratings but no movie slots");
        return;
    }
    DevBuf<int> ptr(static_cast<size_t>(movie_slots) + 1);
    MRB_CUDA(cudaMemsetAsync(ptr.p, 0, sizeof(int) * (static_cast<size_t>(movie_slots) + 1), s));
    DevBuf<int> order(n);
    if (n > 0) {
        const int blocks = ceil_div(n, PT);
        DevBuf<int> key_lo(n), key_hi(n), key_a(n), key_b(n), perm_a(n), perm_b(n);
        DevBuf<unsigned> orand(4);
        const unsigned init[4] = {0u, 0xFFFFFFFFu, 0u, 0xFFFFFFFFu};
        MRB_CUDA(cudaMemcpyAsync(orand.p, init, sizeof(init), cudaMemcpyHostToDevice, s));
        k_rating_keys<<<std::min(blocks, 148 * 8), PT, 0, s>>>(d_rating, n, key_lo.p, key_hi.p, orand.p); MRB_LAUNCHED(1);
        MRB_CUDA(cudaGetLastError());
        unsigned h[4];
        MRB_CUDA(cudaMemcpyAsync(h, orand.p, sizeof(h), cudaMemcpyDeviceToHost, s));
        MRB_CUDA(cudaStreamSynchronize(s));
        // (1) by the low word of the rating, (2) stably by the high word, (3) stably by movie
        stable_sort_pairs(key_lo.p, nullptr, n, digits_that_vary(h[0], h[1]), key_a.p, perm_a.p, s);
        k_gather_i32<<<blocks, PT, 0, s>>>(key_hi.p, perm_a.p, n, key_b.p); MRB_LAUNCHED(1);
        MRB_CUDA(cudaGetLastError());
        stable_sort_pairs(key_b.p, perm_a.p, n, digits_that_vary(h[2], h[3]), key_a.p, perm_b.p, s);
        k_gather_i32<<<blocks, PT, 0, s>>>(d_movie, perm_b.p, n, key_b.p); MRB_LAUNCHED(1);
        MRB_CUDA(cudaGetLastError());
        unsigned movie_digits = 0;
        for (int d = 0; d < 4; d++)
            if ((static_cast<unsigned>(movie_slots - 1) >> (8 * d)) != 0u) movie_digits |= 1u << d;
        stable_sort_pairs(key_b.p, perm_b.p, n, movie_digits, key_a.p, order.p, s);
        k_count<<<blocks, PT, 0, s>>>(d_movie, n, ptr.p); MRB_LAUNCHED(1);
        MRB_CUDA(cudaGetLastError());
        exclusive_scan_i32(ptr.p, ptr.p, static_cast<long long>(movie_slots) + 1, s);
    }
    k_median<<<ceil_div(movie_slots, PT), PT, 0, s>>>(ptr.p, order.p, d_rating, movie_slots,
                                                      d_median, d_count); MRB_LAUNCHED(1);
    MRB_CUDA(cudaGetLastError());
    MRB_CUDA(cudaStreamSynchronize(s));  // temporaries are freed on return
}

ShrinkCounts als_shrink(const int* d_user, const int* d_movie, const double* d_rating, int n,
                        int user_slots, int movie_slots, const double* d_median, int min_user,
                        int min_movie, int* d_out_user, int* d_out_movie, double* d_out_rating,
                        int* d_keep_pos, int* d_user_new, int* d_movie_new, cudaStream_t s) {
    MRB_REQUIRE(n >= 0 && user_slots >= 0 && movie_slots >= 0, "als_shrink: negative size");
    MRB_REQUIRE(min_user >= 1 && min_movie >= 1, "als_shrink: minimum counts must be >= 1");
    MRB_REQUIRE(n == 0 || (user_slots > 0 && movie_slots > 0), "als_shrink: ratings but no slots");
    ShrinkCounts out;
    const int ub = ceil_div(user_slots + 1, PT), mb = ceil_div(movie_slots + 1, PT);
    const int nb = ceil_div(static_cast<long long>(n) + 1, PT);
    // ok arrays carry one trailing 0 so that the exclusive scan's last entry is the total
    DevBuf<int> user_ok(static_cast<size_t>(user_slots) + 1), movie_ok(static_cast<size_t>(movie_slots) + 1);
    DevBuf<int> user_cnt(static_cast<size_t>(user_slots) + 1), movie_cnt(static_cast<size_t>(movie_slots) + 1);
    DevBuf<int> changed(1);
    k_fill_i32<<<ub, PT, 0, s>>>(user_ok.p, user_slots, 1); MRB_LAUNCHED(1);
    k_fill_i32<<<mb, PT, 0, s>>>(movie_ok.p, movie_slots, 1); MRB_LAUNCHED(1);
    MRB_CUDA(cudaGetLastError());
    MRB_CUDA(cudaMemsetAsync(user_ok.p + user_slots, 0, sizeof(int), s));
    MRB_CUDA(cudaMemsetAsync(movie_ok.p + movie_slots, 0, sizeof(int), s));

    // the reference's while has_changed loop (movie_lens_data.py:569-588): one host read per round
    for (;;) {
        out.rounds++;
        MRB_CUDA(cudaMemsetAsync(changed.p, 0, sizeof(int), s));
        MRB_CUDA(cudaMemsetAsync(user_cnt.p, 0, sizeof(int) * (static_cast<size_t>(user_slots) + 1), s));
        MRB_CUDA(cudaMemsetAsync(movie_cnt.p, 0, sizeof(int) * (static_cast<size_t>(movie_slots) + 1), s));
        if (n > 0) {
            k_count_alive<<<ceil_div(n, PT * CA_ILP), PT, 0, s>>>(d_user, d_movie, n, user_ok.p, movie_ok.p, 1,
                                                        user_cnt.p); MRB_LAUNCHED(1);
        }
        if (user_slots > 0) {
            k_update_ok<<<ceil_div(user_slots, PT), PT, 0, s>>>(user_cnt.p, user_ok.p, user_slots,
                                                                min_user, 0, changed.p); MRB_LAUNCHED(1);
        }
        if (n > 0) {
            k_count_alive<<<ceil_div(n, PT * CA_ILP), PT, 0, s>>>(d_user, d_movie, n, user_ok.p, movie_ok.p, 0,
                                                        movie_cnt.p); MRB_LAUNCHED(1);
        }
        if (movie_slots > 0) {
            k_update_ok<<<ceil_div(movie_slots, PT), PT, 0, s>>>(movie_cnt.p, movie_ok.p, movie_slots,
                                                                 min_movie, 1, changed.p); MRB_LAUNCHED(1);
        }
        MRB_CUDA(cudaGetLastError());
        if (read_int(changed.p, s) == 0) break;
        MRB_REQUIRE(out.rounds < (1 << 30), "als_shrink: no fixpoint");
    }

    // ascending renumbering of the surviving slots (every surviving id still has >= 1 rating
    // because the minimum counts are >= 1, which is what _collect_ids sees, :589-608)
    DevBuf<int> user_scan(static_cast<size_t>(user_slots) + 1), movie_scan(static_cast<size_t>(movie_slots) + 1);
    exclusive_scan_i32(user_ok.p, user_scan.p, static_cast<long long>(user_slots) + 1, s);
    exclusive_scan_i32(movie_ok.p, movie_scan.p, static_cast<long long>(movie_slots) + 1, s);
    if (user_slots > 0) {
        k_new_ids<<<ceil_div(user_slots, PT), PT, 0, s>>>(user_ok.p, user_scan.p, user_slots, d_user_new); MRB_LAUNCHED(1);
    }
    if (movie_slots > 0) {
        k_new_ids<<<ceil_div(movie_slots, PT), PT, 0, s>>>(movie_ok.p, movie_scan.p, movie_slots, d_movie_new); MRB_LAUNCHED(1);
    }
    MRB_CUDA(cudaGetLastError());
    out.users_out = read_int(user_scan.p + user_slots, s);
    out.movies_out = read_int(movie_scan.p + movie_slots, s);

    if (n > 0) {
        DevBuf<int> flag(static_cast<size_t>(n) + 1), pos(static_cast<size_t>(n) + 1);
        k_alive_flags<<<nb, PT, 0, s>>>(d_user, d_movie, n, user_ok.p, movie_ok.p, flag.p); MRB_LAUNCHED(1);
        MRB_CUDA(cudaGetLastError());
        exclusive_scan_i32(flag.p, pos.p, static_cast<long long>(n) + 1, s);
        k_compact<<<ceil_div(n, PT), PT, 0, s>>>(d_user, d_movie, d_rating, n, flag.p, pos.p, d_user_new,
                                                d_movie_new, d_median, d_out_user, d_out_movie,
                                                d_out_rating, d_keep_pos); MRB_LAUNCHED(1);
        MRB_CUDA(cudaGetLastError());
        out.ratings_out = read_int(pos.p + n, s);
    }
    MRB_CUDA(cudaStreamSynchronize(s));
    return out;
}

}  // namespace mrb
