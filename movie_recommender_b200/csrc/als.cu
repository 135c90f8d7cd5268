// ALS problem set-up and the reference-order sweep loop (see als.cuh).
#include "als.cuh"

#include "index_build.cuh"

namespace mrb {

namespace {
// b[r] = ratings[r] - user_factors[(user_r + 1)*(k+1) - 1]      (matrix.cpp:1012-1031)
__global__ void k_ratings_minus_bias(const double* __restrict__ ratings,
                                     const int* __restrict__ user_ids,
                                     const double* __restrict__ uf, int n_user_factors, int nnz,
                                     double* __restrict__ out) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nnz) return;
    const double bias = uf[static_cast<size_t>(user_ids[r] + 1) * n_user_factors - 1];
    out[r] = __dadd_rn(ratings[r], -bias);
}

__global__ void k_check_ids(const int* __restrict__ ids, int n, int limit, int* __restrict__ bad) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && (ids[i] < 0 || ids[i] >= limit)) atomicExch(bad, 1);
}
}  // namespace

AlsProblem::AlsProblem(const int* user_ids, const int* item_ids, int nnz, const double* ratings,
                       int k, int num_users, int num_items)
    : nnz_(nnz), k_(k), nu_(num_users), ni_(num_items) {
    MRB_REQUIRE(nnz >= 0 && k >= 1 && num_users >= 0 && num_items >= 0, "als: bad sizes");
    MRB_CUDA(cudaStreamCreateWithFlags(&s_, cudaStreamNonBlocking));
    MRB_CUDA(cudaStreamCreateWithFlags(&s_copy_, cudaStreamNonBlocking));
    MRB_CUDA(cudaEventCreateWithFlags(&ev_ratings_, cudaEventDisableTiming));
    MRB_CUDA(cudaEventCreateWithFlags(&ev_factors_, cudaEventDisableTiming));
    MRB_CUDA(cudaEventCreateWithFlags(&ev_user_done_, cudaEventDisableTiming));
    MRB_CUDA(cudaEventCreateWithFlags(&ev_uf_copied_, cudaEventDisableTiming));
    MRB_CUDA(cudaEventCreateWithFlags(&ev_prepared_, cudaEventDisableTiming));
    PhaseTimer t_all("AlsProblem ctor total");
    user_ids_.alloc(nnz);
    item_ids_.alloc(nnz);
    ratings_.alloc(nnz);
    rmb_.alloc(nnz);
    u_ptr_.alloc(static_cast<size_t>(nu_) + 1);
    i_ptr_.alloc(static_cast<size_t>(ni_) + 1);
    u_idx_.alloc(nnz);
    i_idx_.alloc(nnz);
    uf_.alloc(static_cast<size_t>(nu_) * (k + 1));
    itf_.alloc(static_cast<size_t>(ni_) * k);
    // The ids go first on the compute stream; the ratings (half of the bytes, not needed by
    // the grouping) follow on the copy stream while the ids are being checked and grouped.
    user_ids_.upload(user_ids, nnz, s_);
    item_ids_.upload(item_ids, nnz, s_);
    ratings_.upload(ratings, nnz, s_copy_);
    MRB_CUDA(cudaEventRecord(ev_ratings_, s_copy_));
    ratings_pending_ = true;

    // ids must be zero based and inside the factor arrays (python/full_data/cpp_ls.py:120-123);
    // the reference would read out of bounds, we refuse.
    DevBuf<int> bad(1);
    MRB_CUDA(cudaMemsetAsync(bad.p, 0, sizeof(int), s_));
    if (nnz > 0) {
        k_check_ids<<<ceil_div(nnz, 256), 256, 0, s_>>>(user_ids_.p, nnz, nu_, bad.p); MRB_LAUNCHED(1);
        k_check_ids<<<ceil_div(nnz, 256), 256, 0, s_>>>(item_ids_.p, nnz, ni_, bad.p); MRB_LAUNCHED(1);
    }
    int h_bad = 0;
    MRB_CUDA(cudaMemcpyAsync(&h_bad, bad.p, sizeof(int), cudaMemcpyDeviceToHost, s_));
    MRB_CUDA(cudaStreamSynchronize(s_));
    if (h_bad != 0) {
        cudaStreamSynchronize(s_copy_);
        MRB_REQUIRE(false, "als: user/item id outside [0, num_users/num_items)");
    }

    PhaseTimer t_idx("  index build (2 group_by)");
    cudaEvent_t e0, e1;
    MRB_CUDA(cudaEventCreate(&e0));
    MRB_CUDA(cudaEventCreate(&e1));
    MRB_CUDA(cudaEventRecord(e0, s_));
    stable_group_by(user_ids_.p, nnz, nu_, u_ptr_.p, u_idx_.p, s_);
    stable_group_by(item_ids_.p, nnz, ni_, i_ptr_.p, i_idx_.p, s_);
    MRB_CUDA(cudaEventRecord(e1, s_));
    MRB_CUDA(cudaEventSynchronize(e1));
    MRB_CUDA(cudaEventElapsedTime(&index_ms_, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
}

AlsProblem::~AlsProblem() {
    if (s_copy_) cudaStreamSynchronize(s_copy_);
    if (s_) cudaStreamSynchronize(s_);
    for (cudaEvent_t e : gram_events_) cudaEventDestroy(e);
    gram_events_.clear();
    gram_.reset();
    for (cudaEvent_t e : {ev_ratings_, ev_factors_, ev_user_done_, ev_uf_copied_, ev_prepared_})
        if (e) cudaEventDestroy(e);
    if (s_copy_) cudaStreamDestroy(s_copy_);
    if (s_) cudaStreamDestroy(s_);
}

void AlsProblem::wait_ratings() {
    if (!ratings_pending_) return;
    MRB_CUDA(cudaStreamWaitEvent(s_, ev_ratings_, 0));
    ratings_pending_ = false;
}

void AlsProblem::wait_factors() {
    if (!factors_pending_) return;
    MRB_CUDA(cudaStreamWaitEvent(s_, ev_factors_, 0));
    factors_pending_ = false;
}

void AlsProblem::order_after_inputs(cudaStream_t stream) {
    if (stream == s_) {
        wait_factors();
        return;   // s_ is ordered after its own work; the ratings were awaited by ensure_gram
    }
    MRB_CUDA(cudaStreamWaitEvent(stream, ev_ratings_, 0));
    if (factors_recorded_) MRB_CUDA(cudaStreamWaitEvent(stream, ev_factors_, 0));
    if (prepared_recorded_) MRB_CUDA(cudaStreamWaitEvent(stream, ev_prepared_, 0));
}

void AlsProblem::set_factors(const double* user_factors, const double* item_factors) {
    wait_factors();
    uf_.upload(user_factors, uf_.n, s_);
    itf_.upload(item_factors, itf_.n, s_);
    MRB_CUDA(cudaStreamSynchronize(s_));
}

void AlsProblem::set_factors_async(const double* user_factors, const double* item_factors) {
    // behind the ratings on the copy stream; the compute stream picks the event up when the
    // first kernel that reads the factors is about to be enqueued (wait_factors)
    itf_.upload(item_factors, itf_.n, s_copy_);
    uf_.upload(user_factors, uf_.n, s_copy_);
    MRB_CUDA(cudaEventRecord(ev_factors_, s_copy_));
    factors_pending_ = true;
    factors_recorded_ = true;
}

void AlsProblem::set_host_outputs(double* user_factors, double* item_factors) {
    out_uf_ = user_factors;
    out_itf_ = item_factors;
    outputs_written_ = false;
}

void AlsProblem::get_factors(double* user_factors, double* item_factors) {
    wait_factors();
    uf_.download(user_factors, uf_.n, s_);
    itf_.download(item_factors, itf_.n, s_);
    MRB_CUDA(cudaStreamSynchronize(s_));
}

AlsRunInfo AlsProblem::run(int algorithm, double min_r_decrease, int max_iteration,
                           int thread_count) {
    cudaEvent_t e0, e1;
    MRB_CUDA(cudaEventCreate(&e0));
    MRB_CUDA(cudaEventCreate(&e1));
    MRB_CUDA(cudaEventRecord(e0, s_));
    const long long launches0 = g_kernel_launches.load();
    AlsRunInfo info;
    if (algorithm == ALS_GRAM_CG || algorithm == ALS_GRAM_CHOLESKY)
        info = run_gram(algorithm, min_r_decrease, max_iteration);
    else
        info = run_faithful(algorithm, min_r_decrease, max_iteration, thread_count);
    MRB_CUDA(cudaEventRecord(e1, s_));
    MRB_CUDA(cudaEventSynchronize(e1));
    MRB_CUDA(cudaEventElapsedTime(&info.device_ms, e0, e1));
    info.kernel_launches = static_cast<int>(g_kernel_launches.load() - launches0);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return info;
}

// The sweep loop of matrix.cpp:814-890 with the reference's floating-point order.
AlsRunInfo AlsProblem::run_faithful(int algorithm, double min_r_decrease, int max_iteration,
                                    int T) {
    const int n = k_ + 1;
    wait_ratings();
    wait_factors();
    AlsFaithfulOp user_op(nnz_, nu_, user_ids_.p, item_ids_.p, u_ptr_.p, u_idx_.p, itf_.p, n, k_,
                          k_, true);
    AlsFaithfulOp item_op(nnz_, ni_, item_ids_.p, user_ids_.p, i_ptr_.p, i_idx_.p, uf_.p, k_, n,
                          k_, false);
    FaithfulCG user_cg(nnz_, nu_ * n, T, s_);
    FaithfulCG item_cg(nnz_, ni_ * k_, T, s_);

    AlsRunInfo info;
    int sweep = 0;
    double old_rr = 0;
    while (sweep < max_iteration) {
        // algorithm != 1: explicit-transpose CG on the very first sweep only (:824-827, :861-866)
        const int variant = (algorithm == ALS_REF_CG || sweep >= 1) ? 1 : 2;
        CgResult ur = user_cg.solve(user_op, ratings_.p, uf_.p, 0.01, 200, variant);   // :818
        if (nnz_ > 0) {
            k_ratings_minus_bias<<<ceil_div(nnz_, 256), 256, 0, s_>>>(ratings_.p, user_ids_.p,
                                                                      uf_.p, n, nnz_, rmb_.p);
            MRB_LAUNCHED(1);
        }
        CgResult ir = item_cg.solve(item_op, rmb_.p, itf_.p, 0.01, 200, variant);      // :854
        info.cg_iterations += ur.iterations + ir.iterations;
        info.sweeps_run++;
        info.last_rr = ir.final_rr;
        if (sweep >= 3) {                                                              // :871-875
            const double decrease = (old_rr - ir.final_rr) / old_rr;
            if (decrease < min_r_decrease) break;
        }
        old_rr = ir.final_rr;
        sweep++;
    }
    info.sweeps_returned = sweep;
    return info;
}

}  // namespace mrb
