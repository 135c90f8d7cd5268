// Device memory arena (see common.cuh).
#include <exception>
#include <map>
#include <mutex>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "common.cuh"

namespace mrb {

namespace {
struct Arena {
    std::mutex mu;
    // (device, rounded size) -> free blocks
    std::map<std::pair<int, size_t>, std::vector<void*>> free_lists;
    std::unordered_map<void*, std::pair<int, size_t>> live;   // block -> (device, rounded size)
    // blocks whose CUDA IPC handle went to another process: a peer may hold a cached mapping, so
    // they are recycled through the free lists for the life of the process and never cudaFree'd
    std::unordered_set<void*> exported;
    bool enabled = std::getenv("MRB_NO_CACHE") == nullptr;
};
Arena& arena() {
    static Arena* a = new Arena();   // intentionally leaked: must outlive every static DevBuf
    return *a;
}
size_t round_size(size_t bytes) {
    if (bytes >= (1u << 20)) return (bytes + (1u << 20) - 1) & ~static_cast<size_t>((1u << 20) - 1);
    size_t r = 256;
    while (r < bytes) r <<= 1;
    return r;
}
}  // namespace

void arena_trim() {
    Arena& a = arena();
    std::lock_guard<std::mutex> lock(a.mu);
    for (auto& kv : a.free_lists) {
        std::vector<void*> keep;
        for (void* p : kv.second) {
            if (a.exported.count(p)) keep.push_back(p);
            else cudaFree(p);
        }
        kv.second.swap(keep);
    }
}

void arena_pin_exported(void* p) {
    Arena& a = arena();
    std::lock_guard<std::mutex> lock(a.mu);
    a.exported.insert(p);
}

void* arena_alloc(size_t bytes) {
    Arena& a = arena();
    int dev = 0;
    MRB_CUDA(cudaGetDevice(&dev));
    const size_t sz = round_size(bytes);
    {
        std::lock_guard<std::mutex> lock(a.mu);
        auto it = a.free_lists.find({dev, sz});
        if (it != a.free_lists.end() && !it->second.empty()) {
            void* p = it->second.back();
            it->second.pop_back();
            a.live[p] = {dev, sz};
            return p;
        }
    }
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, sz);
    if (e == cudaErrorMemoryAllocation) {   // give cached blocks back and retry once
        cudaGetLastError();
        arena_trim();
        e = cudaMalloc(&p, sz);
    }
    if (e != cudaSuccess)
        throw Error(kErrCuda, std::string("cudaMalloc of ") + std::to_string(sz) +
                                  " bytes failed: " + cudaGetErrorString(e));
    std::lock_guard<std::mutex> lock(a.mu);
    a.live[p] = {dev, sz};
    return p;
}

void arena_free(void* p) {
    if (!p) return;
    Arena& a = arena();
    // a buffer dying while an exception unwinds may still be in use by queued kernels
    if (std::uncaught_exceptions() > 0) cudaDeviceSynchronize();
    std::lock_guard<std::mutex> lock(a.mu);
    auto it = a.live.find(p);
    if (it == a.live.end()) { cudaFree(p); return; }
    const auto key = it->second;
    a.live.erase(it);
    if (a.enabled || a.exported.count(p)) a.free_lists[key].push_back(p);
    else cudaFree(p);
}

}  // namespace mrb
