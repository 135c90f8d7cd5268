// K6: the reference's co-rating similarity (SimilarMovieFinder) on the GPU (see cosim.cu).
#pragma once
#include <vector>

#include "common.cuh"

namespace mrb {

class Cosim {
public:
    // Host CSR views of the same ratings: by movie (raters) and by user (movies, strictly
    // ascending within a user -- checked); rq = 2*rating
    // (ratings on the 0.5 grid, 0 <= rq <= 20); genre bit masks and genre counts per movie
    // (count 0 = the movie has no genre entry).
    Cosim(int num_movies, int num_users, const int* m_ptr, const int* m_user,
          const unsigned char* m_rq, const int* u_ptr, const int* u_movie,
          const unsigned char* u_rq, const unsigned long long* genre_mask, const int* genre_cnt);
    ~Cosim();
    Cosim(const Cosim&) = delete;
    Cosim& operator=(const Cosim&) = delete;

    // Queries q_lo..q_hi-1 (movie list indices).  buff[n] = score boost for n common raters.
    // Outputs (q_hi-q_lo) x num_results list indices (-1 padded) and scores, and the number of
    // results per query.  Returns the CUDA-event time of the kernel in ms.
    float query(int q_lo, int q_hi, const double* buff, int buff_len, int num_results,
                int* out_idx, double* out_score, int* out_count);

    // One pair of movies (list indices): the number of common raters and the cosine of their
    // ratings over them (0 when fewer than 3), as _scaled_dot_product computes them (:72-107).
    void pair(int a, int b, int* n_out, double* sim_out);

private:
    int N_, U_, ctas_ = 0, parts_ = 1, part_movies_ = 32, smem_bytes_ = 0;
    cudaStream_t s_ = nullptr;
    std::vector<int> order_all_;   // movies by number of raters, descending
    std::vector<int> deg_;         // raters per movie
    DevBuf<int> m_ptr_, m_user_, u_ptr_, u_movie_, gcnt_, cand_b_, cand_n_, counter_, split_;
    DevBuf<unsigned char> m_rq_, u_rq_;
    DevBuf<unsigned long long> gmask_;
    DevBuf<unsigned> m_pack_, u_pack_;   // id | rq << 27: what the query kernel streams
    DevBuf<double> cand_s_;
};

}  // namespace mrb
