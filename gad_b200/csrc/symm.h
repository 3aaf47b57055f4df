// symm.h — one symmetric memory region per rank, visible to every rank of a communicator (see symm.cu).
#pragma once
#include <cstddef>
#include <cstdint>
#include <functional>
#include <string>

namespace gadcu {

constexpr int kMaxRanks = 8;

struct SymmRegion {
  size_t bytes = 0;
  void* local = nullptr;                // this rank's memory
  void* peer[kMaxRanks] = {nullptr};    // rank r's memory as mapped into THIS process (peer[rank] == local): loads / stores travel over NVLink
  void* mc = nullptr;                   // multicast mapping over all ranks' memories (NVLS: multimem.ld_reduce / multimem.st), or null
  bool vmm = false;                     // cuMem* allocation (else cudaMalloc + CUDA IPC: no multicast)
  int nranks = 1, rank = 0;
  // bookkeeping for release
  unsigned long long mem_handle[kMaxRanks] = {0};
  unsigned long long mc_handle = 0;
  size_t mapped_bytes = 0;
};

// Collectives over the communicator's ranks (every rank calls them at the same point). `allreduce_min` takes this rank's
// int and returns the minimum over ranks (used as "did everybody succeed" and as a barrier); `allgather` gathers `bytes`
// bytes per rank (host buffers). Both are provided by comm.cu on top of NCCL.
struct SymmCollectives {
  std::function<int(int)> allreduce_min;
  std::function<void(const void* mine, void* all, size_t bytes)> allgather;
};

// Allocates `bytes` on every rank and maps all of them into every process. With `want_multicast` (and hardware support)
// the regions are also bound to one multicast object. `key` makes the rendezvous names unique (the communicator's id).
// Throws GadError on failure AFTER all ranks have agreed to fail (so every rank takes the same branch).
void symm_alloc(SymmRegion& r, int device, int nranks, int rank, size_t bytes, bool want_multicast, const std::string& key,
                const SymmCollectives& coll, std::string* how);
void symm_free(SymmRegion& r);

}  // namespace gadcu
