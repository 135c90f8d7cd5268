// Per-user evaluation of a trained ALS model: predictions for every
// test rating in the reference's arithmetic (python/full_data/als_predictor.py:35-60: a
// sequential sum of separately rounded products, then + user bias, then + movie median) and the
// reference's "rank agreement" (python/full_data/my_util.py:101-145): over all pairs of a user's
// test movies whose ACTUAL ratings differ, the fraction whose PREDICTED ratings are strictly in
// the same order.  Integer pair counts, so the result is exact; the reference loops over Python
// lists per user (worker_process.py:262-306), here one CTA per user counts the pairs.
#include "evaluate.cuh"

namespace mrb {

namespace {

__global__ void k_eval_predict(const int* __restrict__ entry_user_row, const int* __restrict__ entry_movie_row,
                               const double* __restrict__ median, const double* __restrict__ uf,
                               const double* __restrict__ itf, int k, int n, double* __restrict__ pred) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    const int m = entry_movie_row[e], u = entry_user_row[e];
    if (m < 0 || u < 0) return;                       // no prediction (als_predictor.py:41-43)
    const double* a = uf + static_cast<size_t>(u) * (k + 1);
    const double* b = itf + static_cast<size_t>(m) * k;
    double r = 0.0;
    for (int f = 0; f < k; f++) r = __dadd_rn(r, __dmul_rn(a[f], b[f]));   // als_predictor.py:54-55
    r = __dadd_rn(r, a[k]);                                                   // :57
    pred[e] = __dadd_rn(r, median[e]);                                        // :58
}

__global__ void __launch_bounds__(128)
k_eval_pairs(const int* __restrict__ ptr, const int* __restrict__ entry_movie_row,
             const int* __restrict__ entry_user_row, const double* __restrict__ actual,
             const double* __restrict__ pred, long long* __restrict__ agree,
             long long* __restrict__ disagree, int* __restrict__ n_pred) {
    const int u = blockIdx.x;
    const int beg = ptr[u], end = ptr[u + 1];
    long long a = 0, d = 0;
    int np = 0;
    for (int i = beg + threadIdx.x; i < end; i += blockDim.x) {
        if (entry_movie_row[i] < 0 || entry_user_row[i] < 0) continue;
        np++;
        const double ai = actual[i], pi = pred[i];
        for (int j = i + 1; j < end; j++) {
            if (entry_movie_row[j] < 0 || entry_user_row[j] < 0) continue;
            const double aj = actual[j], pj = pred[j];
            if (ai > aj) { if (pi > pj) a++; else d++; }            // my_util.py:136-141
            else if (aj > ai) { if (pj > pi) a++; else d++; }
        }
    }
    __shared__ long long sa[128], sd[128];
    __shared__ int sn[128];
    sa[threadIdx.x] = a;
    sd[threadIdx.x] = d;
    sn[threadIdx.x] = np;
    __syncthreads();
    for (int off = 64; off > 0; off >>= 1) {
        if (threadIdx.x < off) {
            sa[threadIdx.x] += sa[threadIdx.x + off];
            sd[threadIdx.x] += sd[threadIdx.x + off];
            sn[threadIdx.x] += sn[threadIdx.x + off];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) { agree[u] = sa[0]; disagree[u] = sd[0]; n_pred[u] = sn[0]; }
}

}  // namespace

float als_rank_agreement(const int* ptr, int num_users, const int* entry_user_row,
                         const int* entry_movie_row, const double* actual, const double* median,
                         const double* user_factors, int num_user_rows, const double* item_factors,
                         int num_items, int k, long long* agree, long long* disagree, int* n_pred) {
    MRB_REQUIRE(num_users >= 0 && k >= 1 && num_user_rows >= 0 && num_items >= 0, "rank agreement: bad sizes");
    const int n = num_users > 0 ? ptr[num_users] : 0;
    MRB_REQUIRE(n >= 0, "rank agreement: negative entry count");
    cudaStream_t s;
    MRB_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    struct Guard {
        cudaStream_t s;
        cudaEvent_t e0 = nullptr, e1 = nullptr;
        ~Guard() {
            if (e0) cudaEventDestroy(e0);
            if (e1) cudaEventDestroy(e1);
            cudaStreamDestroy(s);
        }
    } g{s};
    MRB_CUDA(cudaEventCreate(&g.e0));
    MRB_CUDA(cudaEventCreate(&g.e1));
    DevBuf<int> d_ptr(static_cast<size_t>(num_users) + 1), d_urow(n), d_mrow(n), d_np(num_users);
    DevBuf<double> d_actual(n), d_median(n), d_pred(n), d_uf(static_cast<size_t>(num_user_rows) * (k + 1)),
        d_itf(static_cast<size_t>(num_items) * k);
    DevBuf<long long> d_agree(num_users), d_dis(num_users);
    d_ptr.upload(ptr, static_cast<size_t>(num_users) + 1, s);
    d_urow.upload(entry_user_row, n, s);
    d_mrow.upload(entry_movie_row, n, s);
    d_actual.upload(actual, n, s);
    d_median.upload(median, n, s);
    d_uf.upload(user_factors, d_uf.n, s);
    d_itf.upload(item_factors, d_itf.n, s);
    // ids index the factor arrays: refuse anything outside (-1 = "no prediction" is allowed)
    check_id_range_allow_minus1(d_urow.p, n, num_user_rows, "rank agreement: user rows", s);
    check_id_range_allow_minus1(d_mrow.p, n, num_items, "rank agreement: movie rows", s);
    MRB_CUDA(cudaEventRecord(g.e0, s));
    if (n > 0) {
        k_eval_predict<<<ceil_div(n, 256), 256, 0, s>>>(d_urow.p, d_mrow.p, d_median.p, d_uf.p, d_itf.p, k, n, d_pred.p);
        MRB_LAUNCHED(1);
    }
    if (num_users > 0) {
        k_eval_pairs<<<num_users, 128, 0, s>>>(d_ptr.p, d_mrow.p, d_urow.p, d_actual.p, d_pred.p, d_agree.p,
                                               d_dis.p, d_np.p);
        MRB_LAUNCHED(1);
    }
    MRB_CUDA(cudaGetLastError());
    MRB_CUDA(cudaEventRecord(g.e1, s));
    d_agree.download(agree, num_users, s);
    d_dis.download(disagree, num_users, s);
    d_np.download(n_pred, num_users, s);
    MRB_CUDA(cudaStreamSynchronize(s));
    float ms = 0;
    MRB_CUDA(cudaEventElapsedTime(&ms, g.e0, g.e1));
    return ms;
}

}  // namespace mrb
