// See peer.cuh.
#include "peer.cuh"

#include <atomic>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <mutex>

namespace mrb {

void arena_pin_exported(void* p);   // arena.cu

namespace {

std::mutex g_mu;
int* g_words = nullptr;             // PEER_MAX arrival slots + [PEER_MAX] timeout marker
std::atomic<int> g_epoch{0};
std::map<std::string, void*> g_ipc_cache;

struct PeerWords {
    int* w[PEER_MAX];
};

__device__ __forceinline__ void st_release_sys(int* p, int v) {
    asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ int ld_acquire_sys(const int* p) {
    int v;
    asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__global__ void k_peer_barrier(PeerWords peers, int* mine, int rank, int world, int epoch) {
    const int t = threadIdx.x;
    __threadfence_system();
    if (t < world && t != rank) st_release_sys(peers.w[t] + rank, epoch);
    if (t < world && t != rank) {
        long long t0 = 0;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        // epochs only grow: a peer that is already one barrier ahead has passed this one
        while (ld_acquire_sys(mine + t) - epoch < 0) {
            long long t1;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            if (t1 - t0 > 10000000000ll) {   // 10 s in ns
                mine[PEER_MAX] = epoch;                    // which barrier of this process
                atomicOr(mine + PEER_MAX + 1, 1 << t);     // which peers were missing
                mine[PEER_MAX + 2 + (t & 3)] = ld_acquire_sys(mine + t);   // a missing peer's last epoch
                break;
            }
            __nanosleep(200);
        }
    }
    __threadfence_system();
}

}  // namespace

int* peer_barrier_words() {
    std::lock_guard<std::mutex> lock(g_mu);
    if (g_words == nullptr) {
        MRB_CUDA(cudaMalloc(reinterpret_cast<void**>(&g_words), sizeof(int) * (PEER_MAX + 8)));
        MRB_CUDA(cudaMemset(g_words, 0, sizeof(int) * (PEER_MAX + 8)));
        MRB_CUDA(cudaDeviceSynchronize());
    }
    return g_words;
}

void enqueue_peer_barrier(int* const* peer_words, int rank, int world, cudaStream_t s) {
    MRB_REQUIRE(world >= 1 && world <= PEER_MAX && rank >= 0 && rank < world, "peer barrier: bad group");
    int* mine = peer_barrier_words();
    PeerWords pw{};
    for (int r = 0; r < world; r++) pw.w[r] = r == rank ? mine : peer_words[r];
    const int epoch = g_epoch.fetch_add(1) + 1;
    k_peer_barrier<<<1, 32, 0, s>>>(pw, mine, rank, world, epoch);
    MRB_LAUNCHED(1);
    MRB_CUDA(cudaGetLastError());
}

bool peer_barrier_timed_out() {
    if (g_words == nullptr) return false;
    int v = 0;
    MRB_CUDA(cudaMemcpy(&v, g_words + PEER_MAX, sizeof(int), cudaMemcpyDeviceToHost));
    return v != 0;
}

std::string peer_barrier_timeout_report() {
    if (g_words == nullptr) return "no barrier words";
    int v[PEER_MAX + 8];
    MRB_CUDA(cudaMemcpy(v, g_words, sizeof(v), cudaMemcpyDeviceToHost));
    char buf[256];
    std::snprintf(buf, sizeof(buf),
                  "barrier %d of this process timed out (enqueued so far: %d), missing ranks mask 0x%x, "
                  "arrival words %d %d %d %d %d %d %d %d",
                  v[PEER_MAX], g_epoch.load(), v[PEER_MAX + 1], v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7]);
    return buf;
}

void ipc_export(void* d_ptr, unsigned char* handle64) {
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    cudaIpcMemHandle_t h;
    MRB_CUDA(cudaIpcGetMemHandle(&h, d_ptr));
    std::memcpy(handle64, &h, 64);
    arena_pin_exported(d_ptr);   // peers may keep it mapped: never give it back to the driver
}

void* ipc_open_cached(const unsigned char* handle64) {
    std::lock_guard<std::mutex> lock(g_mu);
    const std::string key(reinterpret_cast<const char*>(handle64), 64);
    auto it = g_ipc_cache.find(key);
    if (it != g_ipc_cache.end()) return it->second;
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle64, 64);
    void* q = nullptr;
    MRB_CUDA(cudaIpcOpenMemHandle(&q, h, cudaIpcMemLazyEnablePeerAccess));
    g_ipc_cache.emplace(key, q);
    return q;
}

void ipc_close_all() {
    std::lock_guard<std::mutex> lock(g_mu);
    for (auto& kv : g_ipc_cache) cudaIpcCloseMemHandle(kv.second);
    g_ipc_cache.clear();
}

}  // namespace mrb
