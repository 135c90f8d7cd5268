// Host <-> device copies for PAGEABLE host memory.  The reference's callers hand over ordinary
// NumPy arrays (python/full_data/cpp_ls.py:150-151); cudaMemcpyAsync from such memory is staged by
// the driver through one internal buffer on the calling thread (~8 GB/s here: 75 ms for the 581 MB
// of config 3 against 11 ms from page-locked memory).  Large pageable copies are therefore staged
// by this library: a few host threads copy interleaved 4 MB chunks into their own page-locked
// double buffers and enqueue them on their own streams, so the host-side memcpy of one chunk
// overlaps the DMA of the others.  Page-locked sources / destinations and small copies take the
// plain cudaMemcpyAsync path.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

#include "common.cuh"

namespace mrb {

namespace {

constexpr size_t STAGE_CHUNK = size_t(4) << 20;
constexpr size_t STAGE_MIN = size_t(16) << 20;      // below this the plain path is as good
constexpr int STAGE_THREADS_MAX = 16;
// host threads that copy chunks into page-locked buffers; one memcpy thread moves ~10 GB/s, the
// host link takes ~55 (MRB_STAGE_THREADS overrides)
int stage_threads() {
    static const int n = [] {
        const char* e = std::getenv("MRB_STAGE_THREADS");
        const int v = e ? std::atoi(e) : 8;
        return std::max(1, std::min(STAGE_THREADS_MAX, v));
    }();
    return n;
}

struct Lane {
    cudaStream_t stream = nullptr;
    void* buf[2] = {nullptr, nullptr};
    cudaEvent_t ev[2] = {nullptr, nullptr};
};

struct Stager {
    std::mutex mu;               // one staged copy at a time per process
    int device = -1;
    Lane lanes[STAGE_THREADS_MAX];
    void ensure(int dev) {
        if (device == dev) return;
        release();
        for (int t = 0; t < stage_threads(); t++) {
            Lane& l = lanes[t];
            MRB_CUDA(cudaStreamCreateWithFlags(&l.stream, cudaStreamNonBlocking));
            for (int b = 0; b < 2; b++) {
                MRB_CUDA(cudaMallocHost(&l.buf[b], STAGE_CHUNK));
                MRB_CUDA(cudaEventCreateWithFlags(&l.ev[b], cudaEventDisableTiming));
            }
        }
        device = dev;
    }
    void release() {
        for (Lane& l : lanes) {
            for (int b = 0; b < 2; b++) {
                if (l.ev[b]) cudaEventDestroy(l.ev[b]);
                if (l.buf[b]) cudaFreeHost(l.buf[b]);
                l.ev[b] = nullptr;
                l.buf[b] = nullptr;
            }
            if (l.stream) cudaStreamDestroy(l.stream);
            l.stream = nullptr;
        }
        device = -1;
    }
};
Stager& stager() {
    static Stager* s = new Stager();   // intentionally leaked (process lifetime)
    return *s;
}

bool is_pageable(const void* p) {
    cudaPointerAttributes a{};
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return true;
    }
    return a.type == cudaMemoryTypeUnregistered;
}

// direction: true = host -> device
void staged_copy(char* dev, char* host, size_t bytes, bool h2d, cudaStream_t s) {
    Stager& st = stager();
    std::lock_guard<std::mutex> lock(st.mu);
    int device = 0;
    MRB_CUDA(cudaGetDevice(&device));
    st.ensure(device);
    // the staged copies are ordered after everything already enqueued on s ...
    cudaEvent_t before = nullptr;
    MRB_CUDA(cudaEventCreateWithFlags(&before, cudaEventDisableTiming));
    MRB_CUDA(cudaEventRecord(before, s));
    const size_t chunks = (bytes + STAGE_CHUNK - 1) / STAGE_CHUNK;
    const int STAGE_THREADS = stage_threads();
    cudaError_t errs[STAGE_THREADS_MAX];
    std::vector<std::thread> pool;
    for (int t = 0; t < STAGE_THREADS; t++) {
        errs[t] = cudaSuccess;
        pool.emplace_back([&, t] {
            Lane& l = st.lanes[t];
            auto ok = [&](cudaError_t e) { if (e != cudaSuccess && errs[t] == cudaSuccess) errs[t] = e; return e == cudaSuccess; };
            if (!ok(cudaSetDevice(device))) return;
            if (!ok(cudaStreamWaitEvent(l.stream, before, 0))) return;
            int b = 0;
            size_t pending_c[2] = {0, 0};
            bool pending[2] = {false, false};
            for (size_t c = t; c < chunks; c += STAGE_THREADS, b ^= 1) {
                const size_t off = c * STAGE_CHUNK, len = std::min(STAGE_CHUNK, bytes - off);
                if (pending[b]) {                       // the buffer's previous DMA must be over
                    if (!ok(cudaEventSynchronize(l.ev[b]))) return;
                    if (!h2d) {
                        const size_t o2 = pending_c[b] * STAGE_CHUNK;
                        std::memcpy(host + o2, l.buf[b], std::min(STAGE_CHUNK, bytes - o2));
                    }
                }
                if (h2d) {
                    std::memcpy(l.buf[b], host + off, len);
                    if (!ok(cudaMemcpyAsync(dev + off, l.buf[b], len, cudaMemcpyHostToDevice, l.stream))) return;
                } else {
                    if (!ok(cudaMemcpyAsync(l.buf[b], dev + off, len, cudaMemcpyDeviceToHost, l.stream))) return;
                }
                if (!ok(cudaEventRecord(l.ev[b], l.stream))) return;
                pending[b] = true;
                pending_c[b] = c;
            }
            for (int bb = 0; bb < 2; bb++)
                if (pending[bb]) {
                    if (!ok(cudaEventSynchronize(l.ev[bb]))) return;
                    if (!h2d) {
                        const size_t o2 = pending_c[bb] * STAGE_CHUNK;
                        std::memcpy(host + o2, l.buf[bb], std::min(STAGE_CHUNK, bytes - o2));
                    }
                }
        });
    }
    for (std::thread& th : pool) th.join();
    cudaEventDestroy(before);
    for (int t = 0; t < STAGE_THREADS; t++) MRB_CUDA(errs[t]);
    // ... and everything enqueued on s afterwards is ordered after them: every lane's copies
    // have completed (the threads synchronised on their events), so nothing more is needed.
}

}  // namespace

bool copy_is_staged(const void* host, size_t bytes) {
    return bytes >= STAGE_MIN && std::getenv("MRB_NO_STAGING") == nullptr && is_pageable(host);
}

void copy_h2d(void* dev, const void* host, size_t bytes, cudaStream_t s) {
    if (bytes == 0) return;
    if (bytes >= STAGE_MIN && std::getenv("MRB_NO_STAGING") == nullptr && is_pageable(host)) {
        staged_copy(static_cast<char*>(dev), const_cast<char*>(static_cast<const char*>(host)), bytes, true, s);
        return;
    }
    MRB_CUDA(cudaMemcpyAsync(dev, host, bytes, cudaMemcpyHostToDevice, s));
}

void copy_d2h(void* host, const void* dev, size_t bytes, cudaStream_t s) {
    if (bytes == 0) return;
    if (bytes >= STAGE_MIN && std::getenv("MRB_NO_STAGING") == nullptr && is_pageable(host)) {
        staged_copy(const_cast<char*>(static_cast<const char*>(dev)), static_cast<char*>(host), bytes, false, s);
        return;
    }
    MRB_CUDA(cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, s));
}

}  // namespace mrb
