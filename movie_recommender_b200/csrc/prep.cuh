// K7 -- ALS data preparation on the GPU: per-movie medians of the
// training ratings and the ALS data-set "shrink".  Replaces the multi-process Python of
// python/full_data/movie_lens_data.py:453-464, 547-680 and its workers
// python/full_data/movie_lens_data_proc.py:393-431, 455-471, 494-654.  Integer / byte work plus
// one exactly rounded subtraction per rating: results must be BIT-EXACT against the reference.
#pragma once
#include "common.cuh"

namespace mrb {

// medians[m] = numpy.median of the ratings whose movie id is m (NaN when there are none),
// counts[m] = how many there are.  d_median / d_count have movie_slots entries.
void movie_medians(const int* d_movie, const double* d_rating, int n, int movie_slots,
                   double* d_median, int* d_count, cudaStream_t s);

struct ShrinkCounts {
    int ratings_out = 0, users_out = 0, movies_out = 0, rounds = 0;
};

// Fixpoint degree filter (users need >= min_user ratings, movies >= min_movie, alternately until
// nothing changes), then a stable compaction of the surviving ratings with ids renumbered in
// ascending order of their slot and the movie's median subtracted.
//   d_user[i] in [0, user_slots), d_movie[i] in [0, movie_slots)
//   outputs (capacity n): d_out_user, d_out_movie, d_out_rating, d_keep_pos (original position)
//   d_user_new[user_slots], d_movie_new[movie_slots]: new id or -1
ShrinkCounts als_shrink(const int* d_user, const int* d_movie, const double* d_rating, int n,
                        int user_slots, int movie_slots, const double* d_median, int min_user,
                        int min_movie, int* d_out_user, int* d_out_movie, double* d_out_rating,
                        int* d_keep_pos, int* d_user_new, int* d_movie_new, cudaStream_t s);

// Throws kErrArgument if any id[i] is outside [0, slots).
void check_id_range(const int* d_id, int n, int slots, const char* what, cudaStream_t s);
// CSR sanity (rowptr[0] == 0, non-decreasing; colidx in [0, cols)): throws kErrArgument
void check_csr(const int* d_rowptr, int rows, const int* d_colidx, int nnz, int cols, const char* what,
               cudaStream_t s);

}  // namespace mrb
