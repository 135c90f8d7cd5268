// Per-user evaluation of a trained ALS model: predictions + rank agreement (see evaluate.cu).
#pragma once
#include "common.cuh"

namespace mrb {

// Host arrays in, host arrays out; returns the kernel time in ms (CUDA events).
//   ptr[num_users + 1]      CSR pointers of the users' test ratings (entries)
//   entry_user_row[n]       row of the entry's user in user_factors, or -1
//   entry_movie_row[n]      row of the entry's movie in item_factors, or -1 (no prediction)
//   actual[n], median[n]    the test rating and the training median of the entry's movie
// agree / disagree / n_pred per user: pair counts and the number of entries with a prediction.
float als_rank_agreement(const int* ptr, int num_users, const int* entry_user_row,
                         const int* entry_movie_row, const double* actual, const double* median,
                         const double* user_factors, int num_user_rows, const double* item_factors,
                         int num_items, int k, long long* agree, long long* disagree, int* n_pred);

// ids must lie in [-1, limit): throws kErrArgument otherwise (prep.cu)
void check_id_range_allow_minus1(const int* d_ids, int n, int limit, const char* what, cudaStream_t s);

}  // namespace mrb
