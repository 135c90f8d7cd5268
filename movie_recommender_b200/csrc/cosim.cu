// K6 -- the reference's REAL movie-movie similarity (SimilarMovieFinder,
// python/full_data/build_similar_movies_db.py:21-221) on the GPU, bit-exact.
//
// For query movie a and every other movie b (find_similar_movie, :151-180):
//   gate   : both have genres and |g_a & g_b| / min(|g_a|, |g_b|) >= 0.5          (:44-69)
//   common : raters of both; n = #common >= 3                                      (:86-99)
//   sim    : (r_a . r_b) / (|r_a| |r_b|) over the common raters only                (:101-107)
//   score  : sim * (1 + buff(n)), buff = clamp(ln x - ln 3, 0, L),
//            x = 3 + (3 e^L - 3)(n - 3)/(P - 3)                                      (:109-119)
//   keep score > 0.3; if more than 20*num_results survive keep the 20*num_results with the most
//   common raters (stable); stable sort by score descending; first num_results.   (:161-180)
//
// The reference does this with Python dict intersections, O(N^2 deg).  Here the work per query is
// sum over a's raters u of deg(u): for every rater u of a, every movie b that u rated receives
// (n += 1, D += r_a r_b, S_a += r_a^2, S_b += r_b^2).  Ratings live on the 0.5 grid, so with
// q = 2 r the four sums are small exact integers -- exact and order independent, hence
// deterministic.  Reshaping this into dense N x N x U GEMMs would cost 1e5 x more operations.
//
// The accumulators live in SHARED memory: one CTA per SM owns a query and walks the catalogue in
// `parts` movie ranges of `part_movies` movies, 4 x u32 per movie (216 KB per part; 4 parts for
// the 53 889 movies of ML-27M), updated with native 32-bit shared-memory atomics (ATOMS.ADD; a
// 64-bit shared add compiles to a compare-and-swap loop).  Every user's movie list is ascending,
// so the slice of it that falls into a part is a contiguous range whose bounds are precomputed
// once (`split`).  Round 1 kept two packed 64-bit words per movie and CTA in global memory
// (255 MB of scratch, L2-atomic bound, 62.7 GB of DRAM traffic per 8 192 queries); now DRAM
// sees the rating lists only.  The scores are then formed in fp64 with the reference's operation
// order, sqrt / multiply / divide being correctly rounded on both sides; buff(n) comes from a
// host table computed with the same libm calls the reference makes.
#include "cosim.cuh"

#include <algorithm>
#include <numeric>
#include <vector>

namespace mrb {

namespace {

constexpr int CS_THREADS = 1024;
constexpr int CS_KEEP_MAX = 1120;   // 20 * num_results, num_results <= 56
constexpr int CS_KEEP_BYTES = CS_KEEP_MAX * 16;          // kb, kn, ks of the final selection
constexpr int CS_HIST_BYTES = 4096 * 4;                  // select_threshold's histogram, behind them
constexpr int CS_SEL_BYTES = CS_KEEP_BYTES + CS_HIST_BYTES;
constexpr int CS_ACC_BYTES = 216 * 1024;                 // accumulators of one part
constexpr int CS_ID_BITS = 27;                           // packed word: id | rq << 27 (rq <= 20)
constexpr unsigned CS_ID_MASK = (1u << CS_ID_BITS) - 1;
constexpr unsigned CS_NONE = 0xffffffffu;                // rq = 31: never a real entry

struct CosimArgs {
    int num_movies;
    const int* m_ptr;            // CSR by movie: raters
    const int* m_user;
    const unsigned char* m_rq;   // 2 * rating
    const int* u_ptr;            // CSR by user: movies (ascending within a user)
    const int* u_movie;
    const unsigned char* u_rq;
    const unsigned long long* genre_mask;
    const int* genre_cnt;        // 0 = movie has no genre entry
    const double* buff;          // buff[n], n = 0 .. buff_len-1
    int buff_len;
    int num_results;
    int keep;                    // 20 * num_results
    int q_lo, q_hi;
    int* work_counter;
    const unsigned* m_pack;        // by movie: rater | rq << 27 (one load per rater)
    const unsigned* u_pack;        // by user:  movie | rq << 27 (one load per co-rating triple)
    const int* order;              // [n_tickets] queries, most raters first (longest first)
    int n_tickets;
    const int* split;              // [num_users][parts + 1]: where each part starts in u's list
    int parts;
    int part_movies;               // multiple of 32
    int* cand_b;                   // [ctas][num_movies]
    int* cand_n;
    double* cand_s;
    int* out_idx;                  // [(q_hi-q_lo)][num_results], -1 padded
    double* out_score;
    int* out_count;
};

__device__ __forceinline__ int block_sum(int v, int* red) {
    // sum over the CTA; red has CS_THREADS/32 + 1 ints
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    __syncthreads();
    if (lane == 0) red[w] = v;
    __syncthreads();
    int t = 0;
    for (int i = 0; i < CS_THREADS / 32; i++) t += red[i];
    return t;
}

// strict order "a before b" for the reliability cut: more common raters first, then list index
__device__ __forceinline__ bool rel_before(int na, int ba, int nb, int bb) {
    return na > nb || (na == nb && ba < bb);
}

// Largest threshold t with  #{i < C : key(i) >= t} >= k  (the caller guarantees that at least k
// candidates have a key; key(i) < 0 = not a candidate), and how many keys lie strictly above it.
// Two 12-bit digits, one 4096-bin shared-memory histogram each, scanned from the top by the
// whole CTA (thread t owns bins 4t .. 4t+3).  `hist` must be zero on entry and is zero on exit.
// Replaces a 30-step binary search with a block reduction per step.
template <class KeyFn>
__device__ __forceinline__ void select_threshold(int C, int k, KeyFn key, unsigned* hist, int* wtot,
                                                 int* sel, int& thr, int& above) {
    static_assert(CS_THREADS == 1024, "thread t owns histogram bins 4t .. 4t+3");
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int prefix = 0, kk = k, abv = 0;
#pragma unroll 1
    for (int level = 0; level < 2; level++) {
        const int shift = level == 0 ? 12 : 0;
        for (int i = threadIdx.x; i < C; i += CS_THREADS) {
            const int v = key(i);
            if (v >= 0 && (level == 0 || (v >> 12) == prefix)) atomicAdd(&hist[(v >> shift) & 4095], 1u);
        }
        __syncthreads();
        int c[4], tot = 0;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            c[j] = static_cast<int>(hist[4 * threadIdx.x + j]);
            hist[4 * threadIdx.x + j] = 0;
            tot += c[j];
        }
        int v = tot;   // sum over the lanes >= lane of this warp
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int o = __shfl_down_sync(0xffffffffu, v, d);
            if (lane + d < 32) v += o;
        }
        if (lane == 0) wtot[warp] = v;
        __syncthreads();
        int higher = 0;
        for (int w = warp + 1; w < CS_THREADS / 32; w++) higher += wtot[w];
        const int incl = v + higher, excl = incl - tot;   // keys in bins >= 4t, > 4t+3
        if (excl < kk && kk <= incl) {
            int run = excl;
#pragma unroll
            for (int j = 3; j >= 0; j--) {
                if (run >= 0 && run + c[j] >= kk) { sel[0] = 4 * threadIdx.x + j; sel[1] = run; run = -(1 << 30); }
                else if (run >= 0) run += c[j];
            }
        }
        __syncthreads();
        const int digit = sel[0], over = sel[1];
        __syncthreads();   // sel is rewritten by the next level
        abv += over;
        kk -= over;
        prefix = level == 0 ? digit : (prefix << 12) | digit;
    }
    thr = prefix;
    above = abv;
}

__global__ void __launch_bounds__(CS_THREADS)
k_cosim(const CosimArgs A) {
    extern __shared__ __align__(16) unsigned char cs_smem[];
    __shared__ int red[CS_THREADS / 32 + 1];
    __shared__ int s_query, s_count, s_next;
    // accumulators of the current part (structure of arrays: consecutive movies in consecutive
    // banks); the selection arrays alias them once the last part has been scored
    const int PM = A.part_movies;
    unsigned* Sn = reinterpret_cast<unsigned*>(cs_smem);
    unsigned* Sd = Sn + PM;
    unsigned* Sa = Sd + PM;
    unsigned* Sb = Sa + PM;
    int* kb = reinterpret_cast<int*>(cs_smem);
    int* kn = kb + CS_KEEP_MAX;
    double* ks = reinterpret_cast<double*>(kn + CS_KEEP_MAX);
    unsigned* hist = reinterpret_cast<unsigned*>(cs_smem + CS_KEEP_BYTES);
    __shared__ int wtot[CS_THREADS / 32], sel[2];
    const int N = A.num_movies;
    int* cb = A.cand_b + static_cast<size_t>(blockIdx.x) * N;
    int* cn = A.cand_n + static_cast<size_t>(blockIdx.x) * N;
    double* cs = A.cand_s + static_cast<size_t>(blockIdx.x) * N;
    const int lane = threadIdx.x & 31;
    const int zero_words = max(4 * PM, CS_SEL_BYTES / 4);
    bool first = true;

    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) {
            const int t = atomicAdd(A.work_counter, 1);
            s_query = t < A.n_tickets ? A.order[t] : -1;
            s_count = 0;
            s_next = s_query >= 0 ? A.m_ptr[s_query] : 0;
        }
        // all four accumulator arrays are zero between parts (phase 2 clears what it reads); the
        // selection of the previous query left its arrays in the first bytes (its histogram
        // behind them is zero again)
        for (int i = threadIdx.x; i < (first ? zero_words : CS_KEEP_BYTES / 4); i += CS_THREADS)
            reinterpret_cast<unsigned*>(cs_smem)[i] = 0;
        first = false;
        __syncthreads();
        const int a = s_query;
        if (a < 0) break;

        const int ab = A.m_ptr[a], ae = A.m_ptr[a + 1];
        // raters per warp batch: up to 1024 raters every warp gets ONE batch (one pass through the
        // dependent rater -> list bounds -> list address chain per part), beyond that 32 each
        const int batch = min(32, (ae - ab + CS_THREADS / 32 - 1) / (CS_THREADS / 32));
        const unsigned long long ga = A.genre_mask[a];
        const int gca = __popcll(ga);
        for (int part = 0; part < A.parts; part++) {
            const int base = part * PM;
            // ---- phase 1: accumulate over the raters of a.  A warp takes a batch of up to 32
            // raters at a time (dynamic: list lengths vary by orders of magnitude); lane l fetches
            // rater l's id, rating and the bounds of its list inside this part -- the dependent
            // part of the address chain, paid once per batch.  The sub-lists are then walked in
            // chunks of 64 consecutive elements (two coalesced loads per lane), software
            // pipelined: the loads of the next chunk -- of the same rater or the next one -- are
            // issued before the atomics of the current one.
            for (;;) {
                int eb = 0;
                if (lane == 0) eb = atomicAdd(&s_next, batch);
                eb = __shfl_sync(0xffffffffu, eb, 0);
                if (eb >= ae) break;
                const int R = min(batch, ae - eb);
                int fbL = 0, lenL = 0;
                unsigned raL = 0;
                if (lane < R) {
                    const unsigned w = A.m_pack[eb + lane];
                    raL = w >> CS_ID_BITS;
                    const int* sp = A.split + static_cast<size_t>(w & CS_ID_MASK) * (A.parts + 1) + part;
                    fbL = sp[0];
                    lenL = sp[1] - fbL;
                }
                int r = -1, pos = 0, cfb = 0, clen = 0;   // warp-uniform cursor
                unsigned cra = 0;
                // next chunk: its two packed words per lane (CS_NONE past the end of the list)
                // and the rater's rating; false when the batch is exhausted
                auto next_chunk = [&](unsigned& w0, unsigned& w1, unsigned& ra) {
                    while (pos >= clen) {
                        if (++r >= R) return false;
                        pos = 0;
                        cfb = __shfl_sync(0xffffffffu, fbL, r);
                        clen = __shfl_sync(0xffffffffu, lenL, r);
                        cra = __shfl_sync(0xffffffffu, raL, r);
                    }
                    w0 = pos + lane < clen ? A.u_pack[cfb + pos + lane] : CS_NONE;
                    w1 = pos + 32 + lane < clen ? A.u_pack[cfb + pos + 32 + lane] : CS_NONE;
                    ra = cra;
                    pos += 64;
                    return true;
                };
                auto update = [&](unsigned w, unsigned ra) {
                    if (w != CS_NONE) {
                        const int b = static_cast<int>(w & CS_ID_MASK) - base;
                        const unsigned rb = w >> CS_ID_BITS;
                        atomicAdd(&Sn[b], 1u);
                        atomicAdd(&Sd[b], ra * rb);
                        atomicAdd(&Sa[b], ra * ra);
                        atomicAdd(&Sb[b], rb * rb);
                    }
                };
                // three chunks in flight
                unsigned a0 = CS_NONE, a1 = CS_NONE, ra_a = 0, b0 = CS_NONE, b1 = CS_NONE, ra_b = 0;
                bool more_a = next_chunk(a0, a1, ra_a);
                bool more_b = more_a && next_chunk(b0, b1, ra_b);
                while (more_a) {
                    unsigned c0 = CS_NONE, c1 = CS_NONE, ra_c = 0;
                    const bool more_c = more_b && next_chunk(c0, c1, ra_c);
                    update(a0, ra_a);
                    update(a1, ra_a);
                    a0 = b0; a1 = b1; ra_a = ra_b; more_a = more_b;
                    b0 = c0; b1 = c1; ra_b = ra_c; more_b = more_c;
                }
            }
            __syncthreads();
            if (threadIdx.x == 0) s_next = ab;   // for the next part (phase 2 ends with a barrier)

            // ---- phase 2: scores, candidates (score > 0.3), accumulator reset
            const int pm = min(PM, N - base);
            // the genre mask of the NEXT entry is requested before the current one is scored
            // (coalesced, unconditional): the loop is otherwise a chain of L2 round trips
            unsigned long long gm_next = threadIdx.x < pm ? A.genre_mask[base + threadIdx.x] : 0ull;
            for (int bl = threadIdx.x; bl < pm; bl += CS_THREADS) {
                const unsigned long long gb = gm_next;
                if (bl + CS_THREADS < pm) gm_next = A.genre_mask[base + bl + CS_THREADS];
                const int n = static_cast<int>(Sn[bl]);
                if (n == 0) continue;
                const unsigned wd = Sd[bl], wa = Sa[bl], wb = Sb[bl];
                Sn[bl] = 0; Sd[bl] = 0; Sa[bl] = 0; Sb[bl] = 0;
                const int b = base + bl;
                if (b == a) continue;
                const int gcb = __popcll(gb);          // = the movie's genre count (checked at creation)
                if (n < 3 || gca == 0 || gcb == 0) continue;
                const int common = __popcll(ga & gb);
                if (2 * common < min(gca, gcb)) continue;                     // matches / length >= 0.5
                const double dot = static_cast<double>(wd) * 0.25;            // exact
                const double na = sqrt(static_cast<double>(wa) * 0.25);       // |r_a| on the common raters
                const double nb = sqrt(static_cast<double>(wb) * 0.25);
                const double sim = __ddiv_rn(dot, __dmul_rn(na, nb));                       // :107
                const double bf = A.buff[n < A.buff_len ? n : A.buff_len - 1];
                const double score = __dmul_rn(sim, __dadd_rn(1.0, bf));                   // :119
                if (score > 0.3) {
                    const int pos = atomicAdd(&s_count, 1);
                    cb[pos] = b;
                    cn[pos] = n;
                    cs[pos] = score;
                }
            }
            __syncthreads();
        }
        const int C = s_count;
        const int keep = A.keep;
        int kept = C;
        bool truncated = false;

        // ---- reliability cut: more than `keep` candidates -> the `keep` with the most common
        // raters, ties by list index (the reference's stable sort by n descending, :166-168)
        if (C > keep) {
            truncated = true;
            // largest n* with count(n >= n*) >= keep, then -- `need` of the n == n* candidates
            // survive, those with the smallest list index -- the largest index that survives
            int nstar, above, kstar, dummy;
            select_threshold(C, keep, [&](int i) { return cn[i]; }, hist, wtot, sel, nstar, above);
            const int need = keep - above;   // >= 1 by the maximality of n*
            constexpr int KMAX = (1 << 24) - 1;
            select_threshold(C, need, [&](int i) { return cn[i] == nstar ? KMAX - cb[i] : -1; }, hist, wtot,
                             sel, kstar, dummy);
            const int bstar = KMAX - kstar + 1;   // keep b < bstar
            __syncthreads();
            if (threadIdx.x == 0) s_count = 0;
            __syncthreads();
            for (int i = threadIdx.x; i < C; i += CS_THREADS) {
                const bool in = cn[i] > nstar || (cn[i] == nstar && cb[i] < bstar);
                if (in) {
                    const int pos = atomicAdd(&s_count, 1);
                    if (pos < CS_KEEP_MAX) { kb[pos] = cb[i]; kn[pos] = cn[i]; ks[pos] = cs[i]; }
                }
            }
            __syncthreads();
            kept = min(s_count, CS_KEEP_MAX);
        } else {
            for (int i = threadIdx.x; i < C; i += CS_THREADS) { kb[i] = cb[i]; kn[i] = cn[i]; ks[i] = cs[i]; }
            __syncthreads();
        }

        // ---- final order: stable sort by score descending (:171) of a list that is in list-index
        // order (no cut) or in (n desc, index asc) order (after the cut); first num_results
        const int out_row = a - A.q_lo;
        for (int i = threadIdx.x; i < kept; i += CS_THREADS) {
            int rank = 0;
            const double si = ks[i];
            const int bi = kb[i], ni = kn[i];
            for (int j = 0; j < kept; j++) {
                const double sj = ks[j];
                bool before = sj > si;
                if (sj == si && j != i)
                    before = truncated ? rel_before(kn[j], kb[j], ni, bi) : kb[j] < bi;
                rank += before ? 1 : 0;
            }
            if (rank < A.num_results) {
                A.out_idx[static_cast<size_t>(out_row) * A.num_results + rank] = bi;
                A.out_score[static_cast<size_t>(out_row) * A.num_results + rank] = si;
            }
        }
        for (int r = kept + threadIdx.x; r < A.num_results; r += CS_THREADS) {
            A.out_idx[static_cast<size_t>(out_row) * A.num_results + r] = -1;
            A.out_score[static_cast<size_t>(out_row) * A.num_results + r] = 0.0;
        }
        if (threadIdx.x == 0) A.out_count[out_row] = min(kept, A.num_results);
    }
}

// One warp per user: the movie list must be strictly ascending and inside the catalogue (the
// per-part ranges below and the shared-memory indexing rely on it).
__global__ void k_cosim_check_lists(const int* __restrict__ u_ptr, const int* __restrict__ u_movie,
                                    int num_users, int num_movies, int* __restrict__ bad) {
    const int u = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (u >= num_users) return;
    for (int f = u_ptr[u] + lane; f < u_ptr[u + 1]; f += 32) {
        const int b = u_movie[f];
        if (b < 0 || b >= num_movies || (f > u_ptr[u] && u_movie[f - 1] >= b)) *bad = 1;
    }
}

__global__ void k_cosim_pack(const int* __restrict__ id, const unsigned char* __restrict__ rq, size_t n,
                             unsigned* __restrict__ out) {
    const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) out[i] = static_cast<unsigned>(id[i]) | (static_cast<unsigned>(rq[i]) << CS_ID_BITS);
}

// split[u][j] = first position of u's list whose movie is >= j * part_movies (binary search).
__global__ void k_cosim_split(const int* __restrict__ u_ptr, const int* __restrict__ u_movie,
                              int num_users, int parts, int part_movies, int* __restrict__ split) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= num_users * (parts + 1)) return;
    const int u = t / (parts + 1), j = t - u * (parts + 1);
    int lo = u_ptr[u], hi = u_ptr[u + 1];
    if (j == parts) { split[t] = hi; return; }
    const int key = j * part_movies;
    while (lo < hi) {
        const int mid = lo + (hi - lo) / 2;
        if (u_movie[mid] < key) lo = mid + 1; else hi = mid;
    }
    split[t] = lo;
}

}  // namespace

Cosim::Cosim(int num_movies, int num_users, const int* m_ptr, const int* m_user,
             const unsigned char* m_rq, const int* u_ptr, const int* u_movie,
             const unsigned char* u_rq, const unsigned long long* genre_mask,
             const int* genre_cnt)
    : N_(num_movies), U_(num_users) {
    MRB_REQUIRE(num_movies >= 0 && num_users >= 0, "cosim: negative size");
    MRB_CUDA(cudaStreamCreateWithFlags(&s_, cudaStreamNonBlocking));
    const size_t nnz = static_cast<size_t>(m_ptr[num_movies]);
    MRB_REQUIRE(nnz == static_cast<size_t>(u_ptr[num_users]), "cosim: the two CSR views disagree");
    m_ptr_.alloc(num_movies + 1ull); m_user_.alloc(nnz); m_rq_.alloc(nnz);
    u_ptr_.alloc(num_users + 1ull); u_movie_.alloc(nnz); u_rq_.alloc(nnz);
    gmask_.alloc(std::max(num_movies, 1)); gcnt_.alloc(std::max(num_movies, 1));
    m_ptr_.upload(m_ptr, num_movies + 1ull, s_); m_user_.upload(m_user, nnz, s_); m_rq_.upload(m_rq, nnz, s_);
    u_ptr_.upload(u_ptr, num_users + 1ull, s_); u_movie_.upload(u_movie, nnz, s_); u_rq_.upload(u_rq, nnz, s_);
    gmask_.upload(genre_mask, num_movies, s_); gcnt_.upload(genre_cnt, num_movies, s_);
    int dev = 0, sms = 148;
    MRB_CUDA(cudaGetDevice(&dev));
    MRB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    ctas_ = sms;   // one CTA per SM: the accumulators of a part fill its shared memory
    const size_t n1 = std::max<size_t>(static_cast<size_t>(num_movies), 1);
    // parts: as few as fit CS_ACC_BYTES of accumulators (16 B per movie) each
    parts_ = static_cast<int>((n1 * 16 + CS_ACC_BYTES - 1) / CS_ACC_BYTES);
    part_movies_ = static_cast<int>((n1 + parts_ - 1) / parts_ + 31) / 32 * 32;
    MRB_REQUIRE(static_cast<size_t>(part_movies_) * 16 <= CS_ACC_BYTES, "cosim: part does not fit shared memory");
    smem_bytes_ = std::max(part_movies_ * 16, CS_SEL_BYTES);
    MRB_CUDA(cudaFuncSetAttribute(k_cosim, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes_));
    // queries are handed out longest first: movies by number of raters, descending (stable)
    order_all_.resize(num_movies);
    std::iota(order_all_.begin(), order_all_.end(), 0);
    std::stable_sort(order_all_.begin(), order_all_.end(), [&](int x, int y) {
        return m_ptr[x + 1] - m_ptr[x] > m_ptr[y + 1] - m_ptr[y];
    });
    int max_deg = 0;
    deg_.resize(num_movies);
    for (int i = 0; i < num_movies; i++) {
        deg_[i] = m_ptr[i + 1] - m_ptr[i];
        max_deg = std::max(max_deg, deg_[i]);
    }
    // 32-bit sums: 400 (= 20 * 20) per common rater at most
    MRB_REQUIRE(static_cast<long long>(max_deg) * 400 < (1LL << 32), "cosim: movie with too many raters");
    MRB_REQUIRE(num_movies < (1 << 24) && num_users < (1 << CS_ID_BITS), "cosim: more than 2^24 movies or 2^27 users");
    for (int i = 0; i < num_movies; i++)
        MRB_REQUIRE(genre_cnt[i] == __builtin_popcountll(genre_mask[i]),
                    "cosim: genre_cnt must be the number of bits set in genre_mask");
    for (size_t e = 0; e < nnz; e++) MRB_REQUIRE(m_rq[e] <= 20 && u_rq[e] <= 20, "cosim: rating outside [0, 10]");
    split_.alloc(static_cast<size_t>(std::max(num_users, 1)) * (parts_ + 1));
    m_pack_.alloc(nnz);
    u_pack_.alloc(nnz);
    cand_b_.alloc(static_cast<size_t>(ctas_) * n1);
    cand_n_.alloc(static_cast<size_t>(ctas_) * n1);
    cand_s_.alloc(static_cast<size_t>(ctas_) * n1);
    counter_.alloc(1);
    DevBuf<int> bad(1);
    MRB_CUDA(cudaMemsetAsync(bad.p, 0, sizeof(int), s_));
    if (nnz > 0) {
        k_cosim_pack<<<ceil_div(nnz, 256), 256, 0, s_>>>(m_user_.p, m_rq_.p, nnz, m_pack_.p); MRB_LAUNCHED(1);
        k_cosim_pack<<<ceil_div(nnz, 256), 256, 0, s_>>>(u_movie_.p, u_rq_.p, nnz, u_pack_.p); MRB_LAUNCHED(1);
    }
    if (num_users > 0) {
        k_cosim_check_lists<<<ceil_div(num_users * 32ll, 256), 256, 0, s_>>>(u_ptr_.p, u_movie_.p, num_users,
                                                                          num_movies, bad.p);
        MRB_LAUNCHED(1);
        k_cosim_split<<<ceil_div(num_users * (parts_ + 1ll), 256), 256, 0, s_>>>(u_ptr_.p, u_movie_.p, num_users,
                                                                                 parts_, part_movies_, split_.p);
        MRB_LAUNCHED(1);
        MRB_CUDA(cudaGetLastError());
    }
    int h_bad = 0;
    bad.download(&h_bad, 1, s_);
    MRB_CUDA(cudaStreamSynchronize(s_));
    MRB_REQUIRE(h_bad == 0, "cosim: every user's movie list must be strictly ascending and inside the catalogue");
}

Cosim::~Cosim() {
    if (s_) cudaStreamDestroy(s_);
}

namespace {
// One pair (a, b): common raters n, D = sum q_a q_b, S_a = sum q_a^2, S_b = sum q_b^2 over them
// (q = 2 * rating: small exact integers), then the cosine in the reference's operation order
// (build_similar_movies_db.py:72-107).  For every rater u of a, u's movie list is scanned for b.
__global__ void __launch_bounds__(256)
k_cosim_pair(const int* __restrict__ m_ptr, const int* __restrict__ m_user, const unsigned char* __restrict__ m_rq,
             const int* __restrict__ u_ptr, const int* __restrict__ u_movie, const unsigned char* __restrict__ u_rq,
             int a, int b, int* __restrict__ n_out, double* __restrict__ sim_out) {
    __shared__ unsigned long long acc[4];
    if (threadIdx.x < 4) acc[threadIdx.x] = 0;
    __syncthreads();
    unsigned long long n = 0, d = 0, sa = 0, sb = 0;
    for (int e = m_ptr[a] + threadIdx.x; e < m_ptr[a + 1]; e += blockDim.x) {
        const int u = m_user[e];
        const unsigned long long ra = m_rq[e];
        for (int f = u_ptr[u]; f < u_ptr[u + 1]; f++)
            if (u_movie[f] == b) {
                const unsigned long long rb = u_rq[f];
                n++;
                d += ra * rb;
                sa += ra * ra;
                sb += rb * rb;
                break;
            }
    }
    atomicAdd(&acc[0], n);
    atomicAdd(&acc[1], d);
    atomicAdd(&acc[2], sa);
    atomicAdd(&acc[3], sb);
    __syncthreads();
    if (threadIdx.x == 0) {
        *n_out = static_cast<int>(acc[0]);
        double sim = 0.0;
        if (acc[0] >= 3) {
            const double dot = static_cast<double>(acc[1]) * 0.25;
            const double na = sqrt(static_cast<double>(acc[2]) * 0.25);
            const double nb = sqrt(static_cast<double>(acc[3]) * 0.25);
            sim = __ddiv_rn(dot, __dmul_rn(na, nb));                       // :107
        }
        *sim_out = sim;
    }
}
}  // namespace

void Cosim::pair(int a, int b, int* n_out, double* sim_out) {
    MRB_REQUIRE(a >= 0 && a < N_ && b >= 0 && b < N_, "cosim: movie index out of range");
    DevBuf<int> d_n(1);
    DevBuf<double> d_sim(1);
    k_cosim_pair<<<1, 256, 0, s_>>>(m_ptr_.p, m_user_.p, m_rq_.p, u_ptr_.p, u_movie_.p, u_rq_.p, a, b, d_n.p, d_sim.p);
    MRB_LAUNCHED(1);
    MRB_CUDA(cudaGetLastError());
    d_n.download(n_out, 1, s_);
    d_sim.download(sim_out, 1, s_);
    MRB_CUDA(cudaStreamSynchronize(s_));
}

float Cosim::query(int q_lo, int q_hi, const double* buff, int buff_len, int num_results,
                   int* out_idx, double* out_score, int* out_count) {
    MRB_REQUIRE(q_lo >= 0 && q_lo <= q_hi && q_hi <= N_, "cosim: bad query range");
    MRB_REQUIRE(num_results >= 1 && num_results * 20 <= CS_KEEP_MAX, "cosim: num_results must be in 1..56");
    MRB_REQUIRE(buff_len >= 1, "cosim: empty buff table");
    const int nq = q_hi - q_lo;
    if (nq == 0) return 0.f;
    DevBuf<double> d_buff(buff_len), d_score(static_cast<size_t>(nq) * num_results);
    DevBuf<int> d_idx(static_cast<size_t>(nq) * num_results), d_cnt(nq);
    d_buff.upload(buff, buff_len, s_);
    MRB_CUDA(cudaMemsetAsync(counter_.p, 0, sizeof(int), s_));
    std::vector<int> order;
    order.reserve(nq);
    for (int m : order_all_)
        if (m >= q_lo && m < q_hi) order.push_back(m);
    DevBuf<int> d_order(order.size());
    d_order.upload(order.data(), order.size(), s_);
    CosimArgs a{};
    a.num_movies = N_;
    a.m_ptr = m_ptr_.p; a.m_user = m_user_.p; a.m_rq = m_rq_.p;
    a.u_ptr = u_ptr_.p; a.u_movie = u_movie_.p; a.u_rq = u_rq_.p;
    a.genre_mask = gmask_.p; a.genre_cnt = gcnt_.p;
    a.buff = d_buff.p; a.buff_len = buff_len;
    a.num_results = num_results; a.keep = 20 * num_results;
    a.q_lo = q_lo; a.q_hi = q_hi;
    a.work_counter = counter_.p;
    a.m_pack = m_pack_.p; a.u_pack = u_pack_.p;
    a.order = d_order.p; a.n_tickets = static_cast<int>(order.size()); a.split = split_.p; a.parts = parts_; a.part_movies = part_movies_;
    a.cand_b = cand_b_.p; a.cand_n = cand_n_.p; a.cand_s = cand_s_.p;
    a.out_idx = d_idx.p; a.out_score = d_score.p; a.out_count = d_cnt.p;
    cudaEvent_t e0, e1;
    MRB_CUDA(cudaEventCreate(&e0));
    MRB_CUDA(cudaEventCreate(&e1));
    MRB_CUDA(cudaEventRecord(e0, s_));
    k_cosim<<<std::min(ctas_, a.n_tickets), CS_THREADS, smem_bytes_, s_>>>(a);
    MRB_LAUNCHED(1);
    MRB_CUDA(cudaGetLastError());
    MRB_CUDA(cudaEventRecord(e1, s_));
    d_idx.download(out_idx, static_cast<size_t>(nq) * num_results, s_);
    d_score.download(out_score, static_cast<size_t>(nq) * num_results, s_);
    d_cnt.download(out_count, nq, s_);
    MRB_CUDA(cudaStreamSynchronize(s_));
    float ms = 0;
    MRB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return ms;
}

}  // namespace mrb
