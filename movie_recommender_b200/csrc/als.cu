// ALS problem set-up and the reference-order sweep loop (see als.cuh).
#include <exception>
#include <thread>

#include "als.cuh"

#include "index_build.cuh"

namespace mrb {

namespace {
struct EventPair {   // RAII: released when an exception unwinds between record and read
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    EventPair() {
        MRB_CUDA(cudaEventCreate(&e0));
        MRB_CUDA(cudaEventCreate(&e1));
    }
    ~EventPair() {
        if (e0) cudaEventDestroy(e0);
        if (e1) cudaEventDestroy(e1);
    }
};

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
                       int k, int num_users, int num_items, int slice_begin, int slice_len)
    : nnz_(nnz), k_(k), nu_(num_users), ni_(num_items), slice_begin_(slice_begin),
      slice_len_(slice_len) {
    MRB_REQUIRE(nnz >= 0 && k >= 1 && num_users >= 0 && num_items >= 0, "als: bad sizes");
    const bool whole = slice_len < 0;
    if (whole) { slice_begin_ = 0; slice_len_ = nnz; }
    MRB_REQUIRE(slice_begin_ >= 0 && slice_len_ >= 0 && slice_begin_ + slice_len_ <= nnz,
                "als: COO slice outside the ratings");
    // RAII for the raw handles: a throwing constructor does not run the destructor
    struct Cleanup {
        AlsProblem* p;
        bool armed = true;
        ~Cleanup() { if (armed) p->destroy_handles(); }
    } cleanup{this};
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
    u_ptr_.alloc(static_cast<size_t>(nu_) + 1);
    i_ptr_.alloc(static_cast<size_t>(ni_) + 1);
    // (u_idx_ / i_idx_ are allocated by build_index: a dealt multi-GPU rank never needs them)
    uf_.alloc(static_cast<size_t>(nu_) * (k + 1));
    itf_.alloc(static_cast<size_t>(ni_) * k);
    // The ids go first on the compute stream; the ratings (half of the bytes, not needed by
    // the grouping) follow on the copy stream while the ids are being checked and grouped.
    // Pageable ratings are staged by host threads and the call blocks: with the whole COO in hand
    // that copy runs on a helper thread while this one checks and groups the ids.
    std::thread helper;
    std::exception_ptr helper_error;
    if (slice_len_ > 0) {
        copy_h2d(user_ids_.p + slice_begin_, user_ids, sizeof(int) * slice_len_, s_);
        copy_h2d(item_ids_.p + slice_begin_, item_ids, sizeof(int) * slice_len_, s_);
        if (whole && copy_is_staged(ratings, sizeof(double) * slice_len_)) {
            int device = 0;
            MRB_CUDA(cudaGetDevice(&device));
            helper = std::thread([&, device] {
                try {
                    MRB_CUDA(cudaSetDevice(device));
                    copy_h2d(ratings_.p + slice_begin_, ratings, sizeof(double) * slice_len_, s_copy_);
                } catch (...) {
                    helper_error = std::current_exception();
                }
            });
        } else {
            copy_h2d(ratings_.p + slice_begin_, ratings, sizeof(double) * slice_len_, s_copy_);
        }
    }
    struct Joiner {
        std::thread& t;
        ~Joiner() { if (t.joinable()) t.join(); }
    } joiner{helper};
    if (whole) build_index();
    if (helper.joinable()) helper.join();
    if (helper_error) std::rethrow_exception(helper_error);
    MRB_CUDA(cudaEventRecord(ev_ratings_, s_copy_));
    ratings_pending_ = true;
    cleanup.armed = false;
}

void AlsProblem::set_coo_peers(const std::vector<int*>& user_ids, const std::vector<int*>& item_ids,
                               const std::vector<double*>& ratings) {
    uid_peers_ = user_ids;
    iid_peers_ = item_ids;
    rating_peers_ = ratings;
}

// The own slice into every peer's copy of the COO: ids behind their upload on the compute
// stream, ratings behind theirs on the copy stream (device-to-peer copies run on the copy
// engines over NVLink).  The caller follows up with a peer barrier on the compute stream.
void AlsProblem::push_coo_slice() {
    if (slice_len_ > 0) {
        for (size_t r = 0; r < uid_peers_.size(); r++) {
            if (static_cast<int>(r) == rank_ || uid_peers_[r] == nullptr) continue;
            MRB_CUDA(cudaMemcpyAsync(uid_peers_[r] + slice_begin_, user_ids_.p + slice_begin_,
                                     sizeof(int) * slice_len_, cudaMemcpyDefault, s_));
            MRB_CUDA(cudaMemcpyAsync(iid_peers_[r] + slice_begin_, item_ids_.p + slice_begin_,
                                     sizeof(int) * slice_len_, cudaMemcpyDefault, s_));
            MRB_CUDA(cudaMemcpyAsync(rating_peers_[r] + slice_begin_, ratings_.p + slice_begin_,
                                     sizeof(double) * slice_len_, cudaMemcpyDefault, s_copy_));
        }
    }
    MRB_CUDA(cudaEventRecord(ev_ratings_, s_copy_));
    // the barrier that follows on s_ must cover the rating pushes as well
    MRB_CUDA(cudaStreamWaitEvent(s_, ev_ratings_, 0));
}

// ids must be zero based and inside the factor arrays (python/full_data/cpp_ls.py:120-123); the
// reference would read out of bounds, we refuse.
void AlsProblem::check_ids() {
    if (ids_checked_) return;
    const int nnz = nnz_;
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
    ids_checked_ = true;
}

void AlsProblem::build_index() {
    if (index_built_) return;
    const int nnz = nnz_;
    check_ids();
    if (u_idx_.n != static_cast<size_t>(nnz)) u_idx_.alloc(nnz);
    if (i_idx_.n != static_cast<size_t>(nnz)) i_idx_.alloc(nnz);
    PhaseTimer t_idx("  index build (2 group_by)");
    EventPair ev;
    MRB_CUDA(cudaEventRecord(ev.e0, s_));
    stable_group_by(user_ids_.p, nnz, nu_, u_ptr_.p, u_idx_.p, s_);
    stable_group_by(item_ids_.p, nnz, ni_, i_ptr_.p, i_idx_.p, s_);
    MRB_CUDA(cudaEventRecord(ev.e1, s_));
    MRB_CUDA(cudaEventSynchronize(ev.e1));
    MRB_CUDA(cudaEventElapsedTime(&index_ms_, ev.e0, ev.e1));
    index_built_ = true;
    pointers_built_ = true;
}

// The row pointers alone (degree histograms + scans): all a rank of a dealt multi-GPU run needs
// of the other ranks' rows -- it groups only the ratings of the rows it owns (als_gram.cu).
void AlsProblem::build_pointers() {
    if (pointers_built_) return;
    check_ids();
    PhaseTimer t_idx("  row pointers (2 histograms)");
    group_pointers(user_ids_.p, nnz_, nu_, u_ptr_.p, s_);
    group_pointers(item_ids_.p, nnz_, ni_, i_ptr_.p, s_);
    pointers_built_ = true;
}

void AlsProblem::destroy_handles() {
    for (cudaEvent_t e : gram_events_) cudaEventDestroy(e);
    gram_events_.clear();
    for (cudaEvent_t* e : {&ev_ratings_, &ev_factors_, &ev_user_done_, &ev_uf_copied_, &ev_prepared_}) {
        if (*e) cudaEventDestroy(*e);
        *e = nullptr;
    }
    if (s_copy_) cudaStreamDestroy(s_copy_);
    if (s_) cudaStreamDestroy(s_);
    s_copy_ = s_ = nullptr;
}

AlsProblem::~AlsProblem() {
    if (s_copy_) cudaStreamSynchronize(s_copy_);
    if (s_) cudaStreamSynchronize(s_);
    gram_.reset();
    destroy_handles();
}

// Rows of the factor matrices from full-size host arrays: into the own replica, then on into
// every peer replica over NVLink (each byte crosses the host link once per box).  Asynchronous
// on the copy stream; the compute stream picks the event up like after set_factors_async.
void AlsProblem::upload_factor_rows(const double* user_factors, const double* item_factors,
                                    int u_lo, int u_hi, int i_lo, int i_hi) {
    MRB_REQUIRE(0 <= u_lo && u_lo <= u_hi && u_hi <= nu_ && 0 <= i_lo && i_lo <= i_hi && i_hi <= ni_,
                "als: factor row range outside the matrices");
    const size_t n = k_ + 1, uo = static_cast<size_t>(u_lo) * n, ub = static_cast<size_t>(u_hi - u_lo) * n;
    const size_t io = static_cast<size_t>(i_lo) * k_, ib = static_cast<size_t>(i_hi - i_lo) * k_;
    if (ib) copy_h2d(itf_.p + io, item_factors + io, sizeof(double) * ib, s_copy_);
    if (ub) copy_h2d(uf_.p + uo, user_factors + uo, sizeof(double) * ub, s_copy_);
    for (size_t r = 0; r < uf_peers_.size(); r++) {
        if (static_cast<int>(r) == rank_ || uf_peers_[r] == nullptr) continue;
        if (ib) MRB_CUDA(cudaMemcpyAsync(itf_peers_[r] + io, itf_.p + io, sizeof(double) * ib, cudaMemcpyDefault, s_copy_));
        if (ub) MRB_CUDA(cudaMemcpyAsync(uf_peers_[r] + uo, uf_.p + uo, sizeof(double) * ub, cudaMemcpyDefault, s_copy_));
    }
    MRB_CUDA(cudaEventRecord(ev_factors_, s_copy_));
    factors_pending_ = true;
    factors_recorded_ = true;
}

void AlsProblem::download_factor_rows(double* user_factors, double* item_factors, int u_lo,
                                      int u_hi, int i_lo, int i_hi, cudaStream_t after) {
    MRB_REQUIRE(0 <= u_lo && u_lo <= u_hi && u_hi <= nu_ && 0 <= i_lo && i_lo <= i_hi && i_hi <= ni_,
                "als: factor row range outside the matrices");
    const size_t n = k_ + 1, uo = static_cast<size_t>(u_lo) * n, ub = static_cast<size_t>(u_hi - u_lo) * n;
    const size_t io = static_cast<size_t>(i_lo) * k_, ib = static_cast<size_t>(i_hi - i_lo) * k_;
    if (ub) copy_d2h(user_factors + uo, uf_.p + uo, sizeof(double) * ub, after);
    if (ib) copy_d2h(item_factors + io, itf_.p + io, sizeof(double) * ib, after);
    MRB_CUDA(cudaStreamSynchronize(after));
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
    EventPair ev;
    cudaEvent_t e0 = ev.e0, e1 = ev.e1;
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
    return info;
}

// The sweep loop of matrix.cpp:814-890 with the reference's floating-point order.
AlsRunInfo AlsProblem::run_faithful(int algorithm, double min_r_decrease, int max_iteration,
                                    int T) {
    const int n = k_ + 1;
    build_index();
    if (rmb_.n != static_cast<size_t>(nnz_)) rmb_.alloc(nnz_);   // rating - user bias, reference-order modes only
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
