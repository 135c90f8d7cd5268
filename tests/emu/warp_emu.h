// Host emulation of one warp for the device functions of csrc/gram_solve.cuh -- TEST
// INFRASTRUCTURE ONLY (tests/test_emu_gram_solve.py).  The 32 lanes are 32 threads; every
// shuffle and every mma is a rendezvous on a barrier, so the data flow between lanes is exactly
// the device's (warp-synchronous) one.  Arithmetic differs from the GPU only where the hardware
// is not specified bit for bit: the summation order inside one mma and the ~22-bit reciprocal
// square root approximation.
#pragma once
#include <barrier>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstring>

namespace warp_emu {

struct Warp {
    std::barrier<> bar{32};
    unsigned char slot[32][8];
    double a[32], b[32];
};
inline thread_local Warp* t_warp = nullptr;
inline thread_local int t_lane = 0;

template <typename T>
inline T exchange(T v, int src) {
    static_assert(sizeof(T) <= 8, "shuffles carry at most 64 bits");
    Warp& w = *t_warp;
    std::memcpy(w.slot[t_lane], &v, sizeof(T));
    w.bar.arrive_and_wait();
    T r;
    std::memcpy(&r, w.slot[src & 31], sizeof(T));
    w.bar.arrive_and_wait();
    return r;
}

// mma.sync.aligned.m8n8k4.row.col.f64: lane (p = lane >> 2, q = lane & 3) holds A[p][q], B[q][p]
// and C[p][2q], C[p][2q + 1].
inline void mma884(double& c0, double& c1, double a, double b) {
    Warp& w = *t_warp;
    w.a[t_lane] = a;
    w.b[t_lane] = b;
    w.bar.arrive_and_wait();
    const int p = t_lane >> 2, q = t_lane & 3;
    double s0 = c0, s1 = c1;
    for (int k = 0; k < 4; k++) {
        s0 = std::fma(w.a[p * 4 + k], w.b[(2 * q) * 4 + k], s0);
        s1 = std::fma(w.a[p * 4 + k], w.b[(2 * q + 1) * 4 + k], s1);
    }
    w.bar.arrive_and_wait();
    c0 = s0;
    c1 = s1;
}

// rsqrt.approx.ftz.f64: about 22 good bits
inline double rsqrt_approx(double d) { return static_cast<double>(static_cast<float>(1.0 / std::sqrt(d))); }

}  // namespace warp_emu

template <typename T>
inline T __shfl_sync(unsigned, T v, int src) { return warp_emu::exchange(v, src); }
template <typename T>
inline T __shfl_xor_sync(unsigned, T v, int mask) { return warp_emu::exchange(v, warp_emu::t_lane ^ mask); }
template <typename T>
inline T __shfl_up_sync(unsigned, T v, int delta) {      // lanes below delta keep their own value
    return warp_emu::exchange(v, warp_emu::t_lane >= delta ? warp_emu::t_lane - delta : warp_emu::t_lane);
}
inline int __clz(unsigned x) { return x ? __builtin_clz(x) : 32; }
inline int __popc(unsigned x) { return __builtin_popcount(x); }

namespace mrb {
inline double shfl_double(double v, int src) { return warp_emu::exchange(v, src); }
}  // namespace mrb
