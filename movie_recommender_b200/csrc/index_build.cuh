// K4 -- stable index build on the GPU (replaces the reference's threaded counting sort,
// cpp/ls_lib/matrix.cpp:617-738).  Everything here is integer work and must be BIT-EXACT
// against a stable sort (numpy.argsort(kind="stable") / oracle_group_by / oracle_transpose).
#pragma once
#include "common.cuh"

namespace mrb {

// Exclusive prefix sum of int32 (in place allowed).  n may be 0.
void exclusive_scan_i32(const int* d_in, int* d_out, long long n, cudaStream_t s);

// Stable grouping of positions 0..n-1 by key in [0, num_groups):
//   d_ptr[g] .. d_ptr[g+1] delimit, in d_idx, the positions whose key is g, ascending.
// d_ptr has num_groups+1 entries, d_idx has n.
void stable_group_by(const int* d_key, int n, int num_groups, int* d_ptr, int* d_idx,
                     cudaStream_t s);

// The pointer array of stable_group_by alone: d_ptr[g] = number of keys below g.
void group_pointers(const int* d_key, int n, int num_groups, int* d_ptr, cudaStream_t s);

// Stable LSD radix sort of (key, value) pairs on selected 8-bit digits of the 32-bit key: bit d
// of digit_mask selects bits 8d..8d+7 (unselected digits are ignored, e.g. because every key
// carries the same value there).  d_val_in == nullptr means value i = i.  Outputs must not
// alias the inputs.  With digit_mask == 0 the pairs are copied through unchanged.
void stable_sort_pairs(const int* d_key, const int* d_val_in, int n, unsigned digit_mask,
                       int* d_key_out, int* d_val_out, cudaStream_t s);

// Stable CSR -> CSC (explicit transpose), matrix.cpp:617-692: for every column the entries in
// ascending (row, position-in-row) order.  Outputs: t_ptr[cols+1], t_row[nnz], t_val[nnz].
void csr_transpose(int rows, int cols, int nnz, const int* d_rowptr, const int* d_colidx,
                   const double* d_vals, int* d_t_ptr, int* d_t_row, double* d_t_val,
                   cudaStream_t s);

}  // namespace mrb
