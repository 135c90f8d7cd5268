// Common plumbing for the B200 (sm_100a) least-squares / ALS library: error handling, device
// buffers, exact-rounding arithmetic helpers.  No torch types anywhere in this library.
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <chrono>
#include <cstdlib>
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>

namespace mrb {

// Error codes returned through the C ABI (never on a success path; the reference's functions
// return iteration counts >= 0, cpp/ls_lib/ls_linux_dll.cpp:28-103).
// (values mirror the MRB_ERR_* macros of include/cpp_ls_b200.h)
enum : int {
    kErrCuda = -1,       // a CUDA runtime call failed (no device, OOM, launch failure ...)
    kErrArgument = -2,   // dimension mismatch / unsupported size (the reference throws)
    kErrInternal = -3,
};

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

void set_last_error(const std::string& msg);

#define MRB_CUDA(expr)                                                                     \
    do {                                                                                   \
        cudaError_t e__ = (expr);                                                          \
        if (e__ != cudaSuccess)                                                            \
            throw ::mrb::Error(::mrb::kErrCuda, std::string(#expr) + " failed at " +   \
                                                        __FILE__ + ":" +                   \
                                                        std::to_string(__LINE__) + ": " +  \
                                                        cudaGetErrorString(e__));          \
    } while (0)

#define MRB_REQUIRE(cond, msg)                                            \
    do {                                                                  \
        if (!(cond)) throw ::mrb::Error(::mrb::kErrArgument, (msg));  \
    } while (0)

// Device memory arena.  cudaMalloc / cudaFree of the multi-hundred-MB buffers of one als() call
// cost tens to hundreds of milliseconds on this platform (and cudaFree synchronises the device),
// which dominated the end-to-end time of the drop-in call.  Freed blocks are therefore kept in
// per-size free lists and handed out again; a repeated call with the same shapes allocates
// nothing.  Safe because every library entry point synchronises its stream before its buffers
// die (and a buffer released during exception unwinding synchronises the device first).
// MRB_NO_CACHE=1 disables the cache; mrb_trim_memory() returns everything to the driver.
void* arena_alloc(size_t bytes);
void arena_free(void* p);
void arena_trim();

// Host <-> device copies that do not collapse on PAGEABLE host memory (hostcopy.cu): large
// pageable buffers are staged through page-locked chunks by a few host threads; page-locked
// buffers and small copies are plain cudaMemcpyAsync on `s`.  A staged upload returns when the
// source may be reused and is ordered after the work already enqueued on `s`; a staged download
// returns when the data is in the destination.
void copy_h2d(void* dev, const void* host, size_t bytes, cudaStream_t s);
// true when copy_h2d / copy_d2h would stage this host buffer (pageable and large): the call then
// BLOCKS the calling thread until the data has moved
bool copy_is_staged(const void* host, size_t bytes);
void copy_d2h(void* host, const void* dev, size_t bytes, cudaStream_t s);

// RAII device allocation.
template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    DevBuf() = default;
    explicit DevBuf(size_t count) { alloc(count); }
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    DevBuf(DevBuf&& o) noexcept : p(o.p), n(o.n) { o.p = nullptr; o.n = 0; }
    DevBuf& operator=(DevBuf&& o) noexcept {
        if (this != &o) { release(); p = o.p; n = o.n; o.p = nullptr; o.n = 0; }
        return *this;
    }
    ~DevBuf() { release(); }
    void alloc(size_t count) {
        release();
        n = count;
        if (count) p = static_cast<T*>(arena_alloc(count * sizeof(T)));
    }
    void release() {
        if (p) arena_free(p);
        p = nullptr;
        n = 0;
    }
    void upload(const T* host, size_t count, cudaStream_t s) {
        if (count) copy_h2d(p, host, count * sizeof(T), s);
    }
    void download(T* host, size_t count, cudaStream_t s) const {
        if (count) copy_d2h(host, p, count * sizeof(T), s);
    }
};

// RAII pinned host allocation (small status words read back every CG batch).
template <typename T>
struct PinnedBuf {
    T* p = nullptr;
    explicit PinnedBuf(size_t count) { MRB_CUDA(cudaMallocHost(reinterpret_cast<void**>(&p), count * sizeof(T))); }
    PinnedBuf(const PinnedBuf&) = delete;
    PinnedBuf& operator=(const PinnedBuf&) = delete;
    ~PinnedBuf() { if (p) cudaFreeHost(p); }
};

// Number of kernels this library has launched (what bench.py reports as gpu_launches).
inline std::atomic<long long> g_kernel_launches{0};
#define MRB_LAUNCHED(n) (::mrb::g_kernel_launches.fetch_add((n), std::memory_order_relaxed))

// Host-side phase timer, printed to stderr when MRB_TIMING is set (development aid).
struct PhaseTimer {
    const char* what;
    std::chrono::steady_clock::time_point t0;
    bool on;
    explicit PhaseTimer(const char* w) : what(w), t0(std::chrono::steady_clock::now()),
                                         on(std::getenv("MRB_TIMING") != nullptr) {}
    ~PhaseTimer() {
        if (on) {
            cudaDeviceSynchronize();
            const double ms = std::chrono::duration<double, std::milli>(
                                  std::chrono::steady_clock::now() - t0).count();
            std::fprintf(stderr, "[mrb timing] %-28s %8.2f ms\n", what, ms);
        }
    }
};

inline int ceil_div(long long a, long long b) { return static_cast<int>((a + b - 1) / b); }

// The reference's chunk boundary formula, evaluated in FLOAT32 exactly as
// cpp/ls_lib/matrix.cpp:12 / :180 / :639 / :768 do: (int)(((float)i) / T * len).
inline void chunk_table(int T, int len, int* bounds /* T+1 */) {
    bounds[0] = 0;
    for (int i = 1; i < T; i++) {
        volatile float q = static_cast<float>(i) / static_cast<float>(T);
        volatile float f = q * static_cast<float>(len);
        bounds[i] = static_cast<int>(f);
    }
    bounds[T] = len;
}

#ifdef __CUDACC__
// Separately rounded multiply / add: the reference is compiled for baseline x86-64 (no FMA), so
// bit-faithful kernels must never contract a*b+c.
__device__ __forceinline__ double xmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double xadd(double a, double b) { return __dadd_rn(a, b); }

__device__ __forceinline__ double shfl_double(double v, int src) {
    return __shfl_sync(0xffffffffu, v, src);
}
#endif

}  // namespace mrb
