// K3: generic sparse least squares with GPU-native summation (see ls_native.cu).
#pragma once
#include "common.cuh"

namespace mrb {

struct LsNativeResult {
    int iterations = 0;
    double final_rr = 0;
    float transpose_ms = 0;   // K4 stable CSR -> CSC
    float solve_ms = 0;       // A^T b + the CG loop, CUDA events
};

// Host CSR in, x in/out (host).  Same algorithm and stopping rule as matrix.cpp:456-529.
LsNativeResult solve_ls_native(int rows, int cols, const int* rowptr, const int* colidx,
                               const double* vals, const double* b, double* x,
                               double min_r_decrease, int max_iteration);

}  // namespace mrb
