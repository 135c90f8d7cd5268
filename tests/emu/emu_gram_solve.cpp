// Runs the device function mrb::gram_solve<M8> (csrc/gram_solve.cuh) on an emulated warp.
// TEST INFRASTRUCTURE ONLY.  Build: tests/emu/Makefile.
#define MRB_HOST_EMU 1
#include "gram_solve.cuh"

#include <thread>
#include <vector>

namespace {
// peers / row_offset: the fused all-gather of the multi-GPU path -- the solved row is also stored
// at peers[j] + row_offset of every other replica (GramArgs::x_peers, NVLink peer memory on the GPU)
template <int M8, int VARIANT>
void run(int n, const double* aug, int ld, double* x, double* sse, int n_peers = 0,
         double* const* peers = nullptr, size_t row_offset = 0) {
    constexpr int ST = M8 * (M8 + 1) / 2;
    warp_emu::Warp warp;
    mrb::GramArgs args{};
    args.n_peers = n_peers;
    for (int j = 0; j < n_peers; j++) args.x_peers[j] = peers[j];
    std::vector<std::thread> lanes;
    for (int lane = 0; lane < 32; lane++)
        lanes.emplace_back([&, lane] {
            warp_emu::t_warp = &warp;
            warp_emu::t_lane = lane;
            const int p = lane >> 2, q = lane & 3;
            double acc[ST][2];
            // the accumulator fragment layout of k_gram: lane (p, q) holds rows 8 ti + p,
            // columns 8 tj + 2q, + 1 of every lower-triangular tile
            for (int ti = 0; ti < M8; ti++)
                for (int tj = 0; tj <= ti; tj++)
                    for (int h = 0; h < 2; h++)
                        acc[mrb::TI(ti, tj)][h] = aug[(8 * ti + p) * ld + 8 * tj + 2 * q + h];
            mrb::gram_solve<M8, VARIANT>(acc, n, x, sse, lane, args, row_offset);
        });
    for (auto& t : lanes) t.join();
}
}  // namespace

// aug: row-major (8 m8) x ld matrix holding the augmented symmetric matrix [G g; g^T s] of order
// n + 1 (lower triangle used, zero beyond); x: n values, previous factors in, solution out;
// sse: receives s - (residual bookkeeping) = sum of squared residuals at the solution.
namespace {
template <int VARIANT>
int dispatch(int m8, int n, const double* aug, int ld, double* x, double* sse, int n_peers = 0,
             double* const* peers = nullptr, size_t row_offset = 0) {
    if (n + 1 > 8 * m8 || n + 1 <= 8 * (m8 - 1)) return -2;   // the rhs must sit in the last tile row
    if (n_peers < 0 || n_peers > 8) return -4;
    switch (m8) {
        case 1: run<1, VARIANT>(n, aug, ld, x, sse, n_peers, peers, row_offset); break;
        case 2: run<2, VARIANT>(n, aug, ld, x, sse, n_peers, peers, row_offset); break;
        case 3: run<3, VARIANT>(n, aug, ld, x, sse, n_peers, peers, row_offset); break;
        case 4: run<4, VARIANT>(n, aug, ld, x, sse, n_peers, peers, row_offset); break;
        case 5: run<5, VARIANT>(n, aug, ld, x, sse, n_peers, peers, row_offset); break;
        case 6: run<6, VARIANT>(n, aug, ld, x, sse, n_peers, peers, row_offset); break;
        case 7: run<7, VARIANT>(n, aug, ld, x, sse, n_peers, peers, row_offset); break;
        default: return -2;
    }
    return 0;
}
}  // namespace

// variants of gram_solve (csrc/gram_solve.cuh): 0 scalar back substitution, 2 W = L_d^-T per
// tile column, 3 panel formed on the tensor cores as well
extern "C" int emu_gram_solve(int variant, int m8, int n, const double* aug, int ld, double* x,
                              double* sse) {
    if (variant == 0) return dispatch<0>(m8, n, aug, ld, x, sse);
    if (variant == 2) return dispatch<2>(m8, n, aug, ld, x, sse);
    if (variant == 3) return dispatch<3>(m8, n, aug, ld, x, sse);
    return -3;
}

// The same with the fused peer stores of the N-GPU path: x is this rank's row (xo = replica +
// row_offset on the device), peers[j] the BASE of replica j; the solved row must land at
// peers[j] + row_offset .. + n in every replica and nowhere else.
extern "C" int emu_gram_solve_peers(int variant, int m8, int n, const double* aug, int ld, double* x,
                                    double* sse, int n_peers, double* const* peers,
                                    unsigned long long row_offset) {
    if (variant == 0) return dispatch<0>(m8, n, aug, ld, x, sse, n_peers, peers, row_offset);
    if (variant == 2) return dispatch<2>(m8, n, aug, ld, x, sse, n_peers, peers, row_offset);
    if (variant == 3) return dispatch<3>(m8, n, aug, ld, x, sse, n_peers, peers, row_offset);
    return -3;
}
