// Cross-GPU plumbing for the sharded ALS path (one process per GPU):
// CUDA IPC mappings of the peers' buffers and a device-side barrier over them, so that the two
// half-sweeps of a sweep chain on the stream without a host round trip or a library collective.
#pragma once
#include <string>
#include "common.cuh"

namespace mrb {

// This process's barrier words: PEER_MAX arrival slots (slot r is written by rank r) plus a
// timeout marker.  Allocated once per process, zeroed, never freed or recycled, so the cached
// IPC mappings the peers hold stay valid for the life of the process.
constexpr int PEER_MAX = 8;
int* peer_barrier_words();

// Enqueues one barrier on `s`: every rank of the group must enqueue the same sequence of
// barriers (a process-wide epoch counter numbers them).  The kernel publishes this rank's
// arrival to every peer (system-scope release store after a system fence, i.e. after everything
// this stream did before, peer stores and peer copies included) and spins until every peer's
// arrival for this epoch is visible locally (acquire loads).  A peer that never arrives trips a
// ~10 s timeout that is reported by peer_barrier_timed_out() instead of hanging the GPU.
void enqueue_peer_barrier(int* const* peer_words /* [world], own entry ignored */, int rank,
                          int world, cudaStream_t s);
bool peer_barrier_timed_out();
std::string peer_barrier_timeout_report();   // which barrier, which ranks were missing   // after a stream synchronisation

// Peer mappings are cached by handle for the life of the process: the arena hands a re-created
// problem the same device blocks, so a training loop that builds one problem per step opens
// every peer buffer once (cudaIpcOpenMemHandle costs milliseconds).  To keep a cached mapping
// from ever pointing at freed memory, a block whose handle has been exported is pinned in the
// arena: arena_trim / MRB_NO_CACHE recycle it instead of returning it to the driver.
void ipc_export(void* d_ptr, unsigned char* handle64);
void* ipc_open_cached(const unsigned char* handle64);
void ipc_close_all();

}  // namespace mrb
