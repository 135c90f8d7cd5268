// Reference-order ("faithful") conjugate gradient on the normal equations.
//
// Same algorithm AND same floating-point association order as the reference's
// cg_least_squares / cg_least_squares2 (cpp/ls_lib/matrix.cpp:456-613) run with `thread_count`
// CPU threads, so the results are bit-identical to the reference at that thread count:
//   * A x      : every row a sequential sum from 0 (matrix.cpp:188-213);
//   * A^T t    : per output column, one sequential partial sum per row chunk of the reference's
//                thread table, partials folded in chunk order (matrix.cpp:217-247, 418-449,
//                106-125); or one sequential sum per column for the explicit-transpose variant
//                (matrix.cpp:536-613);
//   * dots     : one sequential partial per vector chunk, merged serially (matrix.cpp:92-102,
//                375-396);
//   * updates  : c1*a + c2*b with separately rounded multiply and add (matrix.cpp:69-77).
// Parallelism comes from rows / columns / chunks, never from re-associating a sum.
#pragma once
#include <vector>

#include "common.cuh"

namespace mrb {

// A linear operator A (rows x cols) living on the device.
struct FaithfulOp {
    int rows = 0, cols = 0;
    // Optional device flag: when non-null and non-zero the products are skipped (the CG driver
    // enqueues iterations in batches and turns the tail of a batch into no-ops once done).
    const int* guard = nullptr;
    virtual ~FaithfulOp() = default;
    // y[rows] = A x
    virtual void mul(const double* d_x, double* d_y, cudaStream_t s) = 0;
    // y[cols] = A^T t with the row range cut at d_row_bounds[0..nchunks] (nchunks = 1 gives the
    // explicit-transpose order).
    virtual void tmul(const double* d_t, double* d_y, const int* d_row_bounds, int nchunks,
                      cudaStream_t s) = 0;
};

// Segment table of a grouped entry list against a chunk table: segment = maximal run of one
// group's entries that fall into one reference thread chunk.  The reference's A^T t is, per
// output, a fold over chunks of per-chunk sequential sums (matrix.cpp:418-449); the per-chunk
// sums are independent chains, so (group, chunk) segments are the unit of parallel work and a
// heavy group (a movie with 50 000 ratings) is spread over up to T warps instead of one.
struct SegTable {
    DevBuf<int> seg_start;     // [nseg + 1] first entry of each segment
    DevBuf<int> grp_seg_ptr;   // [ngroups + 1] first segment of each group
    DevBuf<double> partial;    // [nseg * width] per-segment sequential sums
    int nseg = 0;
    const int* bounds_key = nullptr;   // cache key: the chunk table it was built for
    int nchunks_key = -1;
};
// d_pos: the row (or rating position) of every grouped entry, ascending inside a group.
void build_segments(SegTable& out, const int* d_grp_ptr, const int* d_pos, int ngroups, int n,
                    const int* d_bounds, int nchunks, int width, cudaStream_t s);

struct CgResult {
    int iterations = 0;
    double final_rr = 0;
};

// Workspace + driver.  One instance can be reused for any number of solves of the same shape.
class FaithfulCG {
public:
    FaithfulCG(int rows, int cols, int thread_count, cudaStream_t s);
    // x (device, length cols) in/out; b (device, length rows).  variant 1 = chunked A^T
    // (cg_least_squares), 2 = explicit-transpose order (cg_least_squares2).
    CgResult solve(FaithfulOp& A, const double* d_b, double* d_x, double min_r_decrease,
                   int max_iteration, int variant);

    struct State {  // lives in device memory; mirrored to pinned host memory per batch
        double rr, alpha, beta, final_rr, one_minus_mrd;
        int it, slow, done, max_it;
    };

private:
    int rows_, cols_, T_;
    cudaStream_t s_;
    DevBuf<double> b2_, r_, Ap_, p_, tmp_, partials_;
    DevBuf<int> row_bounds_, col_bounds_, one_chunk_;
    DevBuf<State> state_;
    PinnedBuf<State> host_state_;
};

// Generic CSR operator with its stable transpose (K3: the solver behind
// cpp_ls.cg_least_squares, python/full_data/cpp_ls.py:47-111).
class CsrFaithfulOp : public FaithfulOp {
public:
    CsrFaithfulOp(int rows, int cols, int nnz, const int* d_rowptr, const int* d_colidx,
                  const double* d_vals, cudaStream_t s);
    void mul(const double* d_x, double* d_y, cudaStream_t s) override;
    void tmul(const double* d_t, double* d_y, const int* d_row_bounds, int nchunks,
              cudaStream_t s) override;

private:
    int nnz_;
    const int *rowptr_, *colidx_;
    const double* vals_;
    DevBuf<int> t_ptr_, t_row_;
    DevBuf<double> t_val_;
    SegTable seg_[2];   // [0] reference thread chunks, [1] a single chunk (explicit transpose)
};

// Implicit ALS operator (one row per rating, values gathered from the opposite side's
// factors; replaces the materialised user_A / item_A of matrix.cpp:754-807, 898-1007).
class AlsFaithfulOp : public FaithfulOp {
public:
    // owner/other: device id arrays in INPUT order; grp_ptr/grp_idx: stable grouping of the
    // ratings by owner; other_f: the opposite side's factors (row stride other_stride);
    // width: unknowns per owner; k: gathered values per row; has_one: trailing 1 (bias column).
    AlsFaithfulOp(int num_ratings, int num_owners, const int* d_owner, const int* d_other,
                  const int* d_grp_ptr, const int* d_grp_idx, const double* d_other_f, int width,
                  int other_stride, int k, bool has_one);
    void mul(const double* d_x, double* d_y, cudaStream_t s) override;
    void tmul(const double* d_t, double* d_y, const int* d_row_bounds, int nchunks,
              cudaStream_t s) override;

private:
    int owners_;
    const int *owner_, *other_, *grp_ptr_, *grp_idx_;
    const double* other_f_;
    int width_, other_stride_, k_;
    bool has_one_;
    SegTable seg_[2];
};

}  // namespace mrb
