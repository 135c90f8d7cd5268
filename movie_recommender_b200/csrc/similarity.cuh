// K5: factor-cosine similarity with per-row top-k (see similarity.cu).
#pragma once
#include "common.cuh"

namespace mrb {

struct SimResult {
    float candidates_ms = 0;   // normalisation + DMMA GEMM fused with candidate selection
    float total_ms = 0;        // + exact re-score, final order, fallbacks (CUDA events)
    int fallback_rows = 0;     // queries recomputed exhaustively (certificate failed)
};

// M: host, n x k row-major factors.  Queries q_lo..q_hi-1; outputs (q_hi-q_lo) x topk, ids -1
// padded when fewer than topk other movies exist.  topk <= 56, k <= 64.
SimResult cosine_topk(const double* M, int n, int k, int topk, int q_lo, int q_hi, int* ids_out,
                      double* scores_out);

}  // namespace mrb
