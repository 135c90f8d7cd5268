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
// q = 2 r the four sums are small exact integers, accumulated with two packed 64-bit integer
// atomics per (a, u, b) -- exact and order independent, hence deterministic.  This is HBM/L2-bound
// integer work (2 x 8 B of atomic traffic per update); reshaping it into dense N x N x U GEMMs
// would cost 1e5 x more operations.  The scores are then formed in fp64 with the reference's
// operation order, sqrt / multiply / divide being correctly rounded on both sides; buff(n) comes
// from a host table computed with the same libm calls the reference makes.
#include "cosim.cuh"

#include <algorithm>
#include <vector>

namespace mrb {

namespace {

constexpr int CS_THREADS = 256;
constexpr int CS_KEEP_MAX = 1120;   // 20 * num_results, num_results <= 56

struct CosimArgs {
    int num_movies;
    const int* m_ptr;            // CSR by movie: raters
    const int* m_user;
    const unsigned char* m_rq;   // 2 * rating
    const int* u_ptr;            // CSR by user: movies
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
    unsigned long long* scratch;   // [ctas][num_movies][2] packed accumulators, zero on entry/exit
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

__global__ void __launch_bounds__(CS_THREADS)
k_cosim(const CosimArgs A) {
    __shared__ int red[CS_THREADS / 32 + 1];
    __shared__ int s_query, s_count;
    __shared__ int kb[CS_KEEP_MAX], kn[CS_KEEP_MAX];
    __shared__ double ks[CS_KEEP_MAX];
    const int N = A.num_movies;
    unsigned long long* S = A.scratch + static_cast<size_t>(blockIdx.x) * N * 2;
    int* cb = A.cand_b + static_cast<size_t>(blockIdx.x) * N;
    int* cn = A.cand_n + static_cast<size_t>(blockIdx.x) * N;
    double* cs = A.cand_s + static_cast<size_t>(blockIdx.x) * N;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) { s_query = A.q_lo + atomicAdd(A.work_counter, 1); s_count = 0; }
        __syncthreads();
        const int a = s_query;
        if (a >= A.q_hi) break;

        // ---- phase 1: accumulate over the raters of a
        const int ab = A.m_ptr[a], ae = A.m_ptr[a + 1];
        for (int e = ab + warp; e < ae; e += CS_THREADS / 32) {
            const int u = A.m_user[e];
            const unsigned long long ra = A.m_rq[e];
            const int fb = A.u_ptr[u], fe = A.u_ptr[u + 1];
            for (int f = fb + lane; f < fe; f += 32) {
                const int b = A.u_movie[f];
                const unsigned long long rb = A.u_rq[f];
                atomicAdd(&S[2 * b], (1ull << 40) | (ra * rb));
                atomicAdd(&S[2 * b + 1], ((ra * ra) << 32) | (rb * rb));
            }
        }
        __syncthreads();

        // ---- phase 2: scores, candidates (score > 0.3), scratch reset
        const unsigned long long ga = A.genre_mask[a];
        const int gca = A.genre_cnt[a];
        for (int b = threadIdx.x; b < N; b += CS_THREADS) {
            const unsigned long long w1 = S[2 * b], w2 = S[2 * b + 1];
            if (w1 == 0) continue;
            S[2 * b] = 0;
            S[2 * b + 1] = 0;
            if (b == a) continue;
            const int n = static_cast<int>(w1 >> 40);
            const int gcb = A.genre_cnt[b];
            if (n < 3 || gca == 0 || gcb == 0) continue;
            const int common = __popcll(ga & A.genre_mask[b]);
            if (2 * common < min(gca, gcb)) continue;                     // matches / length >= 0.5
            const double dot = static_cast<double>(w1 & ((1ull << 40) - 1)) * 0.25;   // exact
            const double na = sqrt(static_cast<double>(w2 >> 32) * 0.25);             // |r_a| on the common raters
            const double nb = sqrt(static_cast<double>(w2 & 0xffffffffull) * 0.25);
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
        const int C = s_count;
        const int keep = A.keep;
        int kept = C;
        bool truncated = false;

        // ---- reliability cut: more than `keep` candidates -> the `keep` with the most common
        // raters, ties by list index (the reference's stable sort by n descending, :166-168)
        if (C > keep) {
            truncated = true;
            // largest n* with count(n >= n*) >= keep
            int lo = 0, hi = 1 << 30;   // invariant: count(n >= lo) >= keep, count(n >= hi) < keep
            while (hi - lo > 1) {
                const int mid = lo + (hi - lo) / 2;
                int c = 0;
                for (int i = threadIdx.x; i < C; i += CS_THREADS) c += cn[i] >= mid ? 1 : 0;
                if (block_sum(c, red) >= keep) lo = mid; else hi = mid;
            }
            const int nstar = lo;
            int c = 0;
            for (int i = threadIdx.x; i < C; i += CS_THREADS) c += cn[i] > nstar ? 1 : 0;
            const int above = block_sum(c, red);
            const int need = keep - above;   // how many of the n == n* candidates survive (smallest index)
            // largest index bound bstar with count(n == n*, b < bstar) <= need  -> keep b < bstar
            int blo = 0, bhi = N + 1;        // invariant: count(b < blo) <= need, count(b < bhi) > need (or bhi = N+1)
            while (bhi - blo > 1) {
                const int mid = blo + (bhi - blo) / 2;
                int cc = 0;
                for (int i = threadIdx.x; i < C; i += CS_THREADS) cc += (cn[i] == nstar && cb[i] < mid) ? 1 : 0;
                if (block_sum(cc, red) <= need) blo = mid; else bhi = mid;
            }
            const int bstar = blo;
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
    ctas_ = sms * 2;
    const size_t n1 = std::max<size_t>(static_cast<size_t>(num_movies), 1);
    scratch_.alloc(static_cast<size_t>(ctas_) * n1 * 2);
    cand_b_.alloc(static_cast<size_t>(ctas_) * n1);
    cand_n_.alloc(static_cast<size_t>(ctas_) * n1);
    cand_s_.alloc(static_cast<size_t>(ctas_) * n1);
    counter_.alloc(1);
    MRB_CUDA(cudaMemsetAsync(scratch_.p, 0, sizeof(unsigned long long) * scratch_.n, s_));
    MRB_CUDA(cudaStreamSynchronize(s_));
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
    CosimArgs a{};
    a.num_movies = N_;
    a.m_ptr = m_ptr_.p; a.m_user = m_user_.p; a.m_rq = m_rq_.p;
    a.u_ptr = u_ptr_.p; a.u_movie = u_movie_.p; a.u_rq = u_rq_.p;
    a.genre_mask = gmask_.p; a.genre_cnt = gcnt_.p;
    a.buff = d_buff.p; a.buff_len = buff_len;
    a.num_results = num_results; a.keep = 20 * num_results;
    a.q_lo = q_lo; a.q_hi = q_hi;
    a.work_counter = counter_.p;
    a.scratch = scratch_.p; a.cand_b = cand_b_.p; a.cand_n = cand_n_.p; a.cand_s = cand_s_.p;
    a.out_idx = d_idx.p; a.out_score = d_score.p; a.out_count = d_cnt.p;
    cudaEvent_t e0, e1;
    MRB_CUDA(cudaEventCreate(&e0));
    MRB_CUDA(cudaEventCreate(&e1));
    MRB_CUDA(cudaEventRecord(e0, s_));
    k_cosim<<<std::min(ctas_, nq), CS_THREADS, 0, s_>>>(a);
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
