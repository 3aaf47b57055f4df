// comm.cu — data parallelism over the batch dimension (north-star subsystem 5): one process per GPU,
// NCCL all-reduce (sum) of the weight gradients over NVLink/NVSwitch. The reference has no
// collective of any kind; the seam is DiffNet::apply_gradient_step, which sums per-example
// gradients before update_weights (gad/src/net_ext.rs:62-70) — sharding the examples across ranks
// and summing the per-rank gradient sums reproduces it up to summation order.
//
// NCCL is bound at run time (dlopen) so the library loads on hosts without it and shares the NCCL
// already mapped by the host process (e.g. torch's) instead of loading a second copy.
#include <dlfcn.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <vector>

#include "algebra.h"
#include "batched.h"
#include "symm.h"

namespace {

typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclFloat32 = 7, ncclFloat64 = 8 };
enum { ncclSum = 0, ncclMin = 3 };

struct Nccl {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

Nccl& nccl() {
  static Nccl n;
  static std::once_flag once;
  std::call_once(once, [] {
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (auto name : names) {
      n.lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
      if (n.lib) break;
    }
    if (!n.lib) return;
    n.GetUniqueId = reinterpret_cast<decltype(n.GetUniqueId)>(dlsym(n.lib, "ncclGetUniqueId"));
    n.CommInitRank = reinterpret_cast<decltype(n.CommInitRank)>(dlsym(n.lib, "ncclCommInitRank"));
    n.CommDestroy = reinterpret_cast<decltype(n.CommDestroy)>(dlsym(n.lib, "ncclCommDestroy"));
    n.AllReduce = reinterpret_cast<decltype(n.AllReduce)>(dlsym(n.lib, "ncclAllReduce"));
    n.AllGather = reinterpret_cast<decltype(n.AllGather)>(dlsym(n.lib, "ncclAllGather"));
    n.GroupStart = reinterpret_cast<decltype(n.GroupStart)>(dlsym(n.lib, "ncclGroupStart"));
    n.GroupEnd = reinterpret_cast<decltype(n.GroupEnd)>(dlsym(n.lib, "ncclGroupEnd"));
    n.GetErrorString = reinterpret_cast<decltype(n.GetErrorString)>(dlsym(n.lib, "ncclGetErrorString"));
  });
  if (!n.lib || !n.AllReduce) gadcu::fail(GADCU_ERR_NCCL, "libnccl.so.2 could not be loaded");
  return n;
}

void nccl_check(ncclResult_t r, const char* what) {
  if (r != 0) gadcu::fail(GADCU_ERR_NCCL, std::string(what) + ": " + (nccl().GetErrorString ? nccl().GetErrorString(r) : "error"));
}

}  // namespace

// ---- symmetric memory for the gradient exchange --------------------------------------------------------------------
// One region per rank (symm.cu), mapped into every process and, where the fabric has it, bound to a multicast object:
//   [ flags: 3 sets | small: 2 x 8 x kSmallBytes | data ]
// flag sets 0 / 1: cross-rank barriers issued on the communicator's / the compute stream (32 words each); set 2: the
// per-block barriers of the NVLS exchange kernel (kNvlsMaxBlocks x 8 words).
// small[parity][r] receives rank r's copy of a batch of small arrays (biases, the loss) to be summed locally.
// data, NVLS mode   : kSlotsNv gradient-sized slots. A dW GEMM writes its partial product straight into a slot; one kernel
//                     then sums every element over the ranks INSIDE the switch (multimem.ld_reduce) and stores the sum to
//                     all ranks (multimem.st), each rank doing 1/N of the elements; the slot IS the gradient array the
//                     store hands out (no staging, no copy-out).
// data, legacy mode : staging[kSlots] (partial tiles pushed by every rank's GEMM epilogue) + result[kSlots], as in round 1.
constexpr int kSlots = 3;
constexpr int kSlotsNv = 6;  // two steps' worth of weight gradients: a step's results may be held while the next one runs
constexpr int kNvlsMaxBlocks = 128;
constexpr size_t kSmallBytes = 256 * 1024;
constexpr size_t kFlagBytes = 2 * 32 * 4 + size_t(kNvlsMaxBlocks) * 8 * 4;  // 4352
constexpr size_t kHeapHeader = 8192 + 2 * 8 * kSmallBytes;
struct SymmHeap {
  size_t slot_bytes = 0;   // legacy: 2 x gradient bytes (k-split pushes two partial blocks); NVLS: capacity of one slot
  size_t grad_bytes = 0;   // the gradient size the heap was made for
  void* local = nullptr;
  void* peer[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};  // peer[rank] == local
  void* mc = nullptr;      // multicast mapping of the whole region (NVLS), or null
  gadcu::SymmRegion region;
  uint32_t* flags(int r, int set) const { return static_cast<uint32_t*>(peer[r]) + (set < 2 ? 32 * set : 64); }
  float* small(int r, int parity, int src) const {
    return reinterpret_cast<float*>(static_cast<char*>(peer[r]) + 8192 + (size_t(parity) * 8 + size_t(src)) * kSmallBytes);
  }
  float* staging(int r, int slot) const { return reinterpret_cast<float*>(static_cast<char*>(peer[r]) + kHeapHeader + size_t(slot) * slot_bytes); }
  float* result(int r, int slot) const { return reinterpret_cast<float*>(static_cast<char*>(peer[r]) + kHeapHeader + size_t(kSlots + slot) * slot_bytes); }
  size_t nv_offset(int slot) const { return kHeapHeader + size_t(slot) * grad_bytes; }  // byte offset of NVLS slot `slot` in every mapping
};

struct gadcu_comm {
  gadcu_ctx* ctx = nullptr;
  ncclComm_t comm = nullptr;
  int nranks = 1, rank = 0;
  cudaStream_t stream = nullptr;  // collectives that overlap the compute stream
  uint64_t pending = 0;           // async collectives issued since the last join
  SymmHeap heap;
  uint64_t exchanges = 0;         // fused exchanges so far (slot = exchanges % kSlots)
  uint32_t barriers[2] = {0, 0};  // cross-rank barriers so far on the communicator's / the compute stream
  uint64_t small_calls = 0;       // small all-reduces so far (parity of the buffer)
  cudaEvent_t slot_done[kSlots] = {nullptr, nullptr, nullptr};
  std::vector<gadcu::ArrayRef> deferred;  // small gradients: one grouped NCCL all-reduce at the join
  std::vector<std::pair<const gadcu_array*, gadcu::ArrayRef>> replaced;  // this sweep: gradient array -> the private copy reduced in its place
  bool no_peer_memory = false;    // CUDA IPC / peer access unavailable: everything goes through NCCL
  std::vector<std::pair<const char*, cudaEvent_t>> trace;  // GADCU_DP_TRACE=1: timeline of one step
  // NVLS exchange
  bool nvls = false;              // the heap has a multicast mapping and GADCU_DP_NVLS != 0
  uint64_t nv_exchanges = 0;      // slot = nv_exchanges % kSlotsNv
  uint32_t nv_seq = 0;            // barrier sequence numbers of the exchange kernel (two per launch)
  gadcu::Ref<gadcu::Buffer> nv_slot_buf[kSlotsNv];  // the gradient array living in each slot (copied out if still held when the slot comes round)
  cudaEvent_t nv_slot_done[kSlotsNv] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  std::string key;                // rendezvous key of this communicator (hash of the NCCL id)
  std::string heap_kind;          // how the symmetric memory was obtained (symm.cu)
  uint64_t heaps_made = 0;
};

using namespace gadcu;

extern "C" {

gadcu_status gadcu_comm_unique_id(void* out_128_bytes) {
  return guard([&] {
    ncclUniqueId id;
    nccl_check(nccl().GetUniqueId(&id), "ncclGetUniqueId");
    std::memcpy(out_128_bytes, &id, sizeof(id));
  });
}

gadcu_status gadcu_comm_init(gadcu_ctx* ctx, int nranks, int rank, const void* id_128_bytes, gadcu_comm** out) {
  return guard([&] {
    auto c = std::make_unique<gadcu_comm>();
    c->ctx = ctx;
    c->nranks = nranks;
    c->rank = rank;
    if (nranks > 1) {
      GADCU_CUDA(cudaSetDevice(ctx->device));
      ncclUniqueId id;
      std::memcpy(&id, id_128_bytes, sizeof(id));
      nccl_check(nccl().CommInitRank(&c->comm, nranks, id, rank), "ncclCommInitRank");
      uint64_t h = 1469598103934665603ull;
      for (unsigned char b : id.internal) { h ^= b; h *= 1099511628211ull; }
      char hex[32];
      snprintf(hex, sizeof(hex), "%016llx", (unsigned long long)h);
      c->key = hex;
      int lo = 0, hi = 0;
      GADCU_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
      GADCU_CUDA(cudaStreamCreateWithPriority(&c->stream, cudaStreamNonBlocking, hi));  // collectives first when SMs free up
    }
    *out = c.release();
  });
}

namespace {
bool small_allreduce(gadcu_comm* comm, gadcu_array* const* arrays, size_t n);
}

// In-place sum over ranks of every array, issued as one NCCL group on the context's stream (so it
// is ordered after the kernels that produced the gradients and before their consumers).
gadcu_status gadcu_comm_allreduce_sum(gadcu_comm* comm, gadcu_array* const* arrays, size_t n) {
  return guard([&] {
    if (comm->nranks <= 1) return;
    comm->ctx->segment_flush();
    for (size_t i = 0; i < n; i++)
      if (arrays[i]->symbolic()) materialize(ArrayRef(arrays[i], true));
    if (small_allreduce(comm, arrays, n)) return;  // once the peer-memory exchange is up, small arrays use it too
    Nccl& nc = nccl();
    nccl_check(nc.GroupStart(), "ncclGroupStart");
    for (size_t i = 0; i < n; i++) {
      gadcu_array* a = arrays[i];
      if (a->symbolic()) materialize(ArrayRef(a, true));  // a deferred elementwise value: compute it (the handle becomes dense)
      if (a->lazy() || a->symbolic()) fail(GADCU_ERR_INVALID, "allreduce needs dense arrays");
      lazy_before_write(comm->ctx, a->ptr());
      a->buf->mutated();
      nccl_check(nc.AllReduce(a->ptr(), a->ptr(), a->elements(), a->dtype == GADCU_F32 ? ncclFloat32 : ncclFloat64, ncclSum,
                              comm->comm, comm->ctx->stream),
                 "ncclAllReduce");
    }
    nccl_check(nc.GroupEnd(), "ncclGroupEnd");
    comm->ctx->count_launch(n);
  });
}

namespace {

// ---- kernels of the fused exchange ----------------------------------------------------------------------
// Cross-GPU barrier number `seq` (1, 2, ...): thread r tells rank r "I am here" and waits for rank r to say the same.
// Everything this GPU wrote to peer memory before the barrier is visible to the peers after it.
struct PeerFlags { uint32_t* p[8]; };
__global__ void dp_barrier_kernel(PeerFlags flags, int nranks, int rank, uint32_t seq) {
  const int r = threadIdx.x;
  if (r < nranks) {
    __threadfence_system();
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flags.p[r] + rank), "r"(seq) : "memory");
    uint32_t seen;
    do {
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(seen) : "l"(flags.p[rank] + r) : "memory");
    } while (seen < seq);
    __threadfence_system();
  }
}

// Owner side of the reduce-scatter + the all-gather: for every tile this rank owns, add the nranks partial tiles in
// rank order (deterministic) and store the sum into every rank's result buffer at the tile's place in the
// column-major [M, N] gradient. 128-bit accesses; the peer stores travel over NVLink.
struct PeerResults { float* p[8]; };
__global__ void __launch_bounds__(256) dp_reduce_publish_kernel(const float4* __restrict__ staging, PeerResults results, uint32_t M,
                                                                uint32_t num_mt, uint32_t tiles, uint32_t nranks, uint32_t rank,
                                                                uint32_t parts /* partial blocks per tile: nranks x k-splits */) {
  const uint32_t owned = tiles > rank ? (tiles - rank + nranks - 1) / nranks : 0;
  const uint64_t total = uint64_t(owned) * 16384u;  // float4s
  for (uint64_t i = uint64_t(blockIdx.x) * 256 + threadIdx.x; i < total; i += uint64_t(gridDim.x) * 256) {
    const uint32_t k = uint32_t(i >> 14), v = uint32_t(i & 16383u);
    const float4* part = staging + (uint64_t(k) * parts) * 16384u + v;
    float4 acc = part[0];
    for (uint32_t r = 1; r < parts; r++) {
      const float4 x = part[uint64_t(r) * 16384u];
      acc.x += x.x; acc.y += x.y; acc.z += x.z; acc.w += x.w;
    }
    const uint32_t t = k * nranks + rank, mt = t % num_mt, nt = t / num_mt;
    const uint32_t row = (v * 4u) & 255u, col = (v * 4u) >> 8;
    const uint64_t off = uint64_t(nt * 256u + col) * M + mt * 256u + row;
    for (uint32_t d = 0; d < nranks; d++) *reinterpret_cast<float4*>(results.p[d] + off) = acc;
  }
}

// Alternative to the tile-pushing GEMM epilogue (GADCU_DP_PUSH=kernel): the GEMM writes its partial product locally at
// full speed and this kernel, on the communicator's stream, scatters the tiles to their owners' staging buffers.
struct PeerStaging { float* p[8]; };
__global__ void __launch_bounds__(256) dp_scatter_kernel(const float* __restrict__ partial, PeerStaging staging, uint32_t M, uint32_t num_mt,
                                                         uint32_t tiles, uint32_t nranks, uint32_t rank) {
  const uint64_t total = uint64_t(tiles) * 16384u;  // float4s
  for (uint64_t i = uint64_t(blockIdx.x) * 256 + threadIdx.x; i < total; i += uint64_t(gridDim.x) * 256) {
    const uint32_t t = uint32_t(i >> 14), v = uint32_t(i & 16383u);
    const uint32_t mt = t % num_mt, nt = t / num_mt;
    const uint32_t row = (v * 4u) & 255u, col = (v * 4u) >> 8;
    const float4 x = __ldcs(reinterpret_cast<const float4*>(partial + uint64_t(nt * 256u + col) * M + mt * 256u + row));
    float* dst = staging.p[t % nranks] + (uint64_t(t / nranks) * nranks + rank) * 65536u + v * 4u;
    *reinterpret_cast<float4*>(dst) = x;
  }
}

// host-side collectives for the rendezvous of symm.cu, on top of NCCL (device scratch, context stream, synchronous)
gadcu::SymmCollectives nccl_collectives(gadcu_comm* comm) {
  gadcu::SymmCollectives c;
  c.allreduce_min = [comm](int v) {
    gadcu_ctx* ctx = comm->ctx;
    int* dev = nullptr;
    GADCU_CUDA(cudaMalloc(&dev, sizeof(int)));
    GADCU_CUDA(cudaMemcpyAsync(dev, &v, sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    nccl_check(nccl().AllReduce(dev, dev, 1, /*ncclInt32*/ 2, ncclMin, comm->comm, ctx->stream), "ncclAllReduce(min)");
    int out = 0;
    GADCU_CUDA(cudaMemcpyAsync(&out, dev, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    GADCU_CUDA(cudaStreamSynchronize(ctx->stream));
    GADCU_CUDA(cudaFree(dev));
    return out;
  };
  c.allgather = [comm](const void* mine, void* all, size_t bytes) {
    gadcu_ctx* ctx = comm->ctx;
    if (!nccl().AllGather) fail(GADCU_ERR_NCCL, "ncclAllGather is not available");
    char* dev = nullptr;
    GADCU_CUDA(cudaMalloc(&dev, bytes * size_t(comm->nranks + 1)));
    GADCU_CUDA(cudaMemcpyAsync(dev, mine, bytes, cudaMemcpyHostToDevice, ctx->stream));
    nccl_check(nccl().AllGather(dev, dev + bytes, bytes, /*ncclUint8*/ 1, comm->comm, ctx->stream), "ncclAllGather");
    GADCU_CUDA(cudaMemcpyAsync(all, dev + bytes, bytes * size_t(comm->nranks), cudaMemcpyDeviceToHost, ctx->stream));
    GADCU_CUDA(cudaStreamSynchronize(ctx->stream));
    GADCU_CUDA(cudaFree(dev));
  };
  return c;
}

void nvls_selfbench(gadcu_comm* comm);
// Every rank calls this at the same point of the same step, with the same `bytes` (the size of the gradient coming through).
void ensure_heap(gadcu_comm* comm, size_t bytes) {
  if (comm->heap.grad_bytes >= bytes) return;
  if (comm->heap.local) {  // sized by the first gradient that came through; a larger one later goes through NCCL
    throw GadError(GADCU_ERR_UNSUPPORTED, "gradient larger than the symmetric heap");
  }
  if (comm->nranks > 8) fail(GADCU_ERR_UNSUPPORTED, "data-parallel exchange: at most 8 ranks");
  gadcu_ctx* ctx = comm->ctx;
  SymmHeap& h = comm->heap;
  // GADCU_DP_NVLS = auto (default) | 1 | 0. The switch-side reduction moves (1 + 1/N) S bytes per direction and GPU where the
  // tile-pushing exchange moves 2 (N-1)/N S: fewer from 4 ranks on, more at 2 — and it leaves the GEMMs alone. Measured
  // (steps/s, C4, one call each): 2 GPUs 206.6 vs 210.1 pushing, 4 GPUs 370.4 vs 361.6, 8 GPUs 635.1 vs 585.2.
  const bool want_nvls = [&] {
    const char* e = getenv("GADCU_DP_NVLS");
    if (!e || std::string(e) == "auto") return comm->nranks >= 4;
    return atoi(e) != 0;
  }();
  const size_t grad = (bytes + 255) / 256 * 256;
  const size_t legacy_slot = 2 * grad;  // (a k-split GEMM pushes two partial blocks per tile and rank)
  const size_t data = std::max(size_t(2 * kSlots) * legacy_slot, size_t(kSlotsNv) * grad);
  const size_t total = kHeapHeader + data;
  gadcu::symm_alloc(h.region, ctx->device, comm->nranks, comm->rank, total, want_nvls, comm->key + "-" + std::to_string(comm->heaps_made++),
                    nccl_collectives(comm), &comm->heap_kind);
  h.local = h.region.local;
  for (int r = 0; r < comm->nranks; r++) h.peer[r] = h.region.peer[r];
  h.mc = h.region.mc;
  h.slot_bytes = legacy_slot;
  h.grad_bytes = grad;
  comm->nvls = want_nvls && h.mc != nullptr;
  // the flags start at zero on every rank before anybody signals
  GADCU_CUDA(cudaMemsetAsync(h.local, 0, 8192, ctx->stream));
  GADCU_CUDA(cudaStreamSynchronize(ctx->stream));
  nccl_collectives(comm).allreduce_min(1);
  for (int s = 0; s < kSlots; s++) GADCU_CUDA(cudaEventCreateWithFlags(&comm->slot_done[s], cudaEventDisableTiming));
  for (int s = 0; s < kSlotsNv; s++) GADCU_CUDA(cudaEventCreateWithFlags(&comm->nv_slot_done[s], cudaEventDisableTiming));
  if (comm->nvls && getenv("GADCU_NVLS_BENCH")) nvls_selfbench(comm);
  if (getenv("GADCU_DP_TRACE") && comm->rank == 0)
    fprintf(stderr, "[dp] symmetric heap: %zu MiB per rank, %s; exchange = %s\n", total >> 20, comm->heap_kind.c_str(),
            comm->nvls ? "NVLS (multimem.ld_reduce + multimem.st, gradients live in the heap)" : "tile-pushing GEMM epilogue + owner reduce/publish");
}

// ---- the NVLS exchange kernel -----------------------------------------------------------------------------------
// One launch per gradient (or per column chunk of the sweep's last one), every rank at once, on the communicator's stream:
//   1. block b of every rank meets block b of every other rank (flags in peer memory): all partial products are complete
//      — each rank's GEMM finished before its kernel started (stream order), and the release / acquire pair at system
//      scope makes those stores visible to the switch's reads;
//   2. this rank's 1/N share of the float4s: multimem.ld_reduce.add.v4.f32 — the NVSwitch reads the address in all N
//      memories and returns the sum (the order in which the switch adds the N terms is fixed by the fabric, not by us:
//      the result is reproducible run to run on one topology and differs from a rank-ordered sum by f32 rounding of an
//      N-term sum) — then multimem.st.v4.f32: one store, the switch writes all N memories;
//   3. the blocks meet again: every rank's memory now holds every share.
// The slot the GEMM wrote its partial product into is overwritten element by element with the sum: it IS the gradient.
struct NvFlags { uint32_t* p[8]; };
__device__ __forceinline__ void nv_block_sync(const NvFlags& f, int nranks, int rank, uint32_t seq) {
  __syncthreads();
  const int t = threadIdx.x;
  if (t < nranks) {
    __threadfence_system();
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(f.p[t] + blockIdx.x * 8 + rank), "r"(seq) : "memory");
    uint32_t seen;
    unsigned long long t0 = 0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    do {
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(seen) : "l"(f.p[rank] + blockIdx.x * 8 + t) : "memory");
      if (seen >= seq) break;
      unsigned long long now;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
      if (now - t0 > 20000000000ull) __trap();  // a lost rank must not hang the GPU for good
    } while (true);
    __threadfence_system();
  }
  __syncthreads();
}
extern "C++" {
template <int U>
__global__ void __launch_bounds__(512) dp_nvls_allreduce_kernel(NvFlags flags, float4* __restrict__ mc, uint64_t v0, uint64_t v1, int nranks, int rank,
                                                                uint32_t seq) {
  nv_block_sync(flags, nranks, rank, seq);
  const uint64_t stride = uint64_t(gridDim.x) * blockDim.x;
  uint64_t i = v0 + uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  // U independent switch reductions in flight per thread (the round trip through the NVSwitch is several microseconds:
  // bandwidth = bytes in flight / latency)
  for (; i + (U - 1) * stride < v1; i += U * stride) {
    float4 a[U];
#pragma unroll
    for (int u = 0; u < U; u++)
      asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
                   : "=f"(a[u].x), "=f"(a[u].y), "=f"(a[u].z), "=f"(a[u].w)
                   : "l"(mc + i + u * stride)
                   : "memory");
#pragma unroll
    for (int u = 0; u < U; u++)
      asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc + i + u * stride), "f"(a[u].x), "f"(a[u].y), "f"(a[u].z),
                   "f"(a[u].w)
                   : "memory");
  }
  for (; i < v1; i += stride) {
    float4 a;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w) : "l"(mc + i) : "memory");
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc + i), "f"(a.x), "f"(a.y), "f"(a.z), "f"(a.w) : "memory");
  }
  nv_block_sync(flags, nranks, rank, seq + 1);
}
}  // extern "C++"

// Sum [byte_offset, byte_offset + bytes) of the heap over the ranks in place (both multiples of 16), on the communicator's stream.
void nvls_allreduce_cfg(gadcu_comm* comm, size_t byte_offset, size_t bytes, unsigned blocks, int unroll) {
  const uint64_t base = byte_offset / 16, V = bytes / 16;
  const uint64_t v0 = base + V * uint64_t(comm->rank) / uint64_t(comm->nranks), v1 = base + V * uint64_t(comm->rank + 1) / uint64_t(comm->nranks);
  NvFlags f;
  for (int r = 0; r < 8; r++) f.p[r] = r < comm->nranks ? comm->heap.flags(r, 2) : nullptr;
  comm->nv_seq += 2;
  float4* mc = static_cast<float4*>(comm->heap.mc);
  if (unroll >= 16) dp_nvls_allreduce_kernel<16><<<blocks, 512, 0, comm->stream>>>(f, mc, v0, v1, comm->nranks, comm->rank, comm->nv_seq - 1);
  else if (unroll >= 8) dp_nvls_allreduce_kernel<8><<<blocks, 512, 0, comm->stream>>>(f, mc, v0, v1, comm->nranks, comm->rank, comm->nv_seq - 1);
  else dp_nvls_allreduce_kernel<4><<<blocks, 512, 0, comm->stream>>>(f, mc, v0, v1, comm->nranks, comm->rank, comm->nv_seq - 1);
  comm->ctx->count_launch();
}
void nvls_allreduce(gadcu_comm* comm, size_t byte_offset, size_t bytes) {
  static const unsigned blocks = [] {
    const char* e = getenv("GADCU_NVLS_BLOCKS");
    // 8 blocks already run the fabric at its rate (measured alone, 8 ranks, 64 MiB: 0.166 ms with 8 blocks, 0.18 with 64)
    // and leave the SMs to the GEMMs the exchange overlaps
    return unsigned(std::max(1, std::min(kNvlsMaxBlocks, e ? atoi(e) : 8)));
  }();
  // (4 in flight: 32 registers x 512 threads fit beside the GEMM's CTA on an SM; 8 would not)
  static const int unroll = [] { const char* e = getenv("GADCU_NVLS_UNROLL"); return e ? atoi(e) : 4; }();
  nvls_allreduce_cfg(comm, byte_offset, bytes, blocks, unroll);
}
// GADCU_NVLS_BENCH=1: right after the heap is up, time the exchange kernel alone over one slot for a grid of launch shapes
void nvls_selfbench(gadcu_comm* comm) {
  const size_t bytes = comm->heap.grad_bytes;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  GADCU_CUDA(cudaMemsetAsync(static_cast<char*>(comm->heap.local) + comm->heap.nv_offset(0), 0, bytes, comm->stream));
  for (unsigned blocks : {8u, 16u, 32u, 64u, 128u})
    for (int unroll : {4, 8, 16}) {
      for (int w = 0; w < 3; w++) nvls_allreduce_cfg(comm, comm->heap.nv_offset(0), bytes, blocks, unroll);
      cudaEventRecord(e0, comm->stream);
      const int reps = 20;
      for (int i = 0; i < reps; i++) nvls_allreduce_cfg(comm, comm->heap.nv_offset(0), bytes, blocks, unroll);
      cudaEventRecord(e1, comm->stream);
      cudaEventSynchronize(e1);
      float ms = 0;
      cudaEventElapsedTime(&ms, e0, e1);
      if (comm->rank == 0)
        fprintf(stderr, "[nvls bench] %d ranks, %zu MiB, blocks %3u x 512 threads, %2d in flight: %.4f ms per all-reduce = %.1f GB/s (bytes / time)\n", comm->nranks,
                bytes >> 20, blocks, unroll, ms / reps, double(bytes) / (ms / reps * 1e-3) / 1e9);
    }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
}

bool tracing() {
  static const bool on = getenv("GADCU_DP_TRACE") != nullptr;
  return on;
}
void mark(gadcu_comm* comm, const char* what, cudaStream_t stream) {
  if (!tracing()) return;
  cudaEvent_t e;
  cudaEventCreate(&e);
  cudaEventRecord(e, stream);
  comm->trace.emplace_back(what, e);
}
void dump_trace(gadcu_comm* comm) {
  if (!tracing() || comm->trace.empty()) return;
  cudaStreamSynchronize(comm->stream);
  cudaStreamSynchronize(comm->ctx->stream);
  static int dumps = 0;
  const bool print = ++dumps == 6;
  for (auto& t : comm->trace) {
    float ms = 0;
    cudaEventElapsedTime(&ms, comm->trace.front().second, t.second);
    if (print) fprintf(stderr, "[dp trace r%d] %-32s %8.3f ms\n", comm->rank, t.first, ms);
  }
  for (auto& t : comm->trace) cudaEventDestroy(t.second);
  comm->trace.clear();
}

void barrier_on(gadcu_comm* comm, cudaStream_t stream) {
  const int set = stream == comm->stream ? 0 : 1;
  PeerFlags f;
  for (int r = 0; r < 8; r++) f.p[r] = r < comm->nranks ? comm->heap.flags(r, set) : nullptr;
  dp_barrier_kernel<<<1, 32, 0, stream>>>(f, comm->nranks, comm->rank, ++comm->barriers[set]);
  comm->ctx->count_launch();
}

// ---- small arrays (bias gradients, the loss): push a packed copy to every rank, barrier, sum in rank order ----------
struct SmallList {
  float* ptr[16];
  uint32_t n[16];
  uint32_t count, total;
};
struct PeerSmall { float* p[8]; };
__global__ void __launch_bounds__(256) dp_small_push_kernel(SmallList list, PeerSmall dst, uint32_t nranks) {
  uint32_t base = 0;
  for (uint32_t a = 0; a < list.count; a++) {
    for (uint32_t i = blockIdx.x * 256 + threadIdx.x; i < list.n[a]; i += gridDim.x * 256) {
      const float v = list.ptr[a][i];
      for (uint32_t d = 0; d < nranks; d++) dst.p[d][base + i] = v;
    }
    base += list.n[a];
  }
}
__global__ void __launch_bounds__(256) dp_small_sum_kernel(SmallList list, const float* __restrict__ parts, uint32_t nranks) {
  uint32_t base = 0;
  for (uint32_t a = 0; a < list.count; a++) {
    for (uint32_t i = blockIdx.x * 256 + threadIdx.x; i < list.n[a]; i += gridDim.x * 256) {
      float acc = parts[base + i];
      for (uint32_t r = 1; r < nranks; r++) acc += parts[size_t(r) * (kSmallBytes / 4) + base + i];
      list.ptr[a][i] = acc;
    }
    base += list.n[a];
  }
}
// true if it handled them (f32, dense, few and small, and the symmetric heap exists)
bool small_allreduce(gadcu_comm* comm, gadcu_array* const* arrays, size_t n) {
  if (tracing() && comm->rank == 0) {
    fprintf(stderr, "[dp trace] small_allreduce: heap %p n %zu:", comm->heap.local, n);
    for (size_t i = 0; i < n; i++)
      fprintf(stderr, " [%s dt%d lazy%d sym%d]", arrays[i]->dims.str().c_str(), int(arrays[i]->dtype), int(arrays[i]->lazy()), int(arrays[i]->symbolic()));
    fprintf(stderr, "\n");
  }
  if (!comm->heap.local || n == 0 || n > 16) return false;
  SmallList list{};
  for (size_t i = 0; i < n; i++) {
    gadcu_array* a = arrays[i];
    if (a->dtype != GADCU_F32 || a->lazy() || a->symbolic()) return false;
    list.ptr[i] = static_cast<float*>(a->ptr());
    list.n[i] = uint32_t(a->elements());
    list.total += uint32_t(a->elements());
  }
  if (size_t(list.total) * 4 > kSmallBytes) return false;
  list.count = uint32_t(n);
  gadcu_ctx* ctx = comm->ctx;
  for (size_t i = 0; i < n; i++) {
    lazy_before_write(ctx, arrays[i]->ptr());
    arrays[i]->buf->mutated();
  }
  const int parity = int(comm->small_calls++ & 1);
  PeerSmall dst;
  for (int r = 0; r < 8; r++) dst.p[r] = r < comm->nranks ? comm->heap.small(r, parity, comm->rank) : nullptr;
  const unsigned grid = unsigned(std::min<uint32_t>((list.total + 255) / 256, 64));
  mark(comm, "small: start", ctx->stream);
  dp_small_push_kernel<<<grid, 256, 0, ctx->stream>>>(list, dst, uint32_t(comm->nranks));
  mark(comm, "small: pushed", ctx->stream);
  barrier_on(comm, ctx->stream);
  mark(comm, "small: barrier done", ctx->stream);
  dp_small_sum_kernel<<<grid, 256, 0, ctx->stream>>>(list, comm->heap.small(comm->rank, parity, 0), uint32_t(comm->nranks));
  mark(comm, "small: summed", ctx->stream);
  ctx->count_launch(2);
  return true;
}

// NVLS form of the exchange: the GEMM is the ordinary one (no pushing epilogue, full speed, last-wave split and all) and
// writes its partial product into a slot of the symmetric heap; ONE kernel on the communicator's stream sums the slot
// over the ranks in place, overlapping whatever the sweep does next; the slot is handed out as the gradient array.
// A slot comes round again kSlotsNv exchanges later: if somebody still holds its array then, the contents move to
// ordinary memory first (the handle is rebound; holders notice nothing).
ArrayRef dp_gemm_nvls(gadcu_comm* comm, const k::GemmArgs& g, const Dim4& rd, bool last) {
  gadcu_ctx* ctx = comm->ctx;
  const size_t bytes = size_t(g.M) * g.N * 4;
  const int slot = int(comm->nv_exchanges % kSlotsNv);
  const bool used_before = comm->nv_exchanges >= uint64_t(kSlotsNv);
  comm->nv_exchanges++;
  if (Buffer* old = comm->nv_slot_buf[slot].get()) {
    if (old->refs.load() > 1) {  // still held by a store / the caller: its data leaves the heap before the slot is written again
      std::shared_ptr<Pool> pool = ctx->active_pool();
      void* moved = pool->take(old->bytes);
      GADCU_CUDA(cudaMemcpyAsync(moved, old->ptr, old->bytes, cudaMemcpyDeviceToDevice, ctx->stream));
      old->ptr = moved;
      old->pool = pool;
    }
    comm->nv_slot_buf[slot].reset();
  }
  // peers finished reading / writing this slot for its previous exchange when that exchange's kernel finished here
  if (used_before) GADCU_CUDA(cudaStreamWaitEvent(ctx->stream, comm->nv_slot_done[slot], 0));
  auto* buf = new Buffer;
  buf->ctx = ctx;
  buf->bytes = bytes;
  buf->ptr = static_cast<char*>(comm->heap.local) + comm->heap.nv_offset(slot);  // pool == null: the heap owns the memory
  comm->nv_slot_buf[slot] = Ref<Buffer>(buf, false);
  auto* arr = new gadcu_array;
  arr->buf = comm->nv_slot_buf[slot];
  arr->dims = rd;
  arr->phys = rd;
  arr->dtype = GADCU_F32;
  arr->ctx = ctx;
  ArrayRef out(arr, false);
  // the sweep's LAST gradient has nothing left to hide behind: cut into column chunks, the exchange of one chunk overlaps
  // the GEMM of the next and only the last chunk's exchange is exposed
  // (measured, 8 GPUs: 2 chunks 595 steps/s vs 635 whole — the half-width GEMMs and the exchange running beside the second
  // one cost more than the overlap buys; GADCU_DP_TAIL_CHUNKS=2|4 to try)
  static const uint64_t max_chunks = [] { const char* e = getenv("GADCU_DP_TAIL_CHUNKS"); return e ? uint64_t(std::max(1, atoi(e))) : uint64_t(1); }();
  uint64_t chunks = 1;
  if (last && !g.tb)
    while (chunks * 2 <= max_chunks && g.N % (chunks * 2 * 256) == 0 && (bytes / (chunks * 2)) % (16 * 8) == 0) chunks *= 2;
  const uint64_t Nc = g.N / chunks;
  const float* Bfull = static_cast<const float*>(g.B);
  for (uint64_t c = 0; c < chunks; c++) {
    k::GemmArgs gc = g;
    gc.push = nullptr;
    gc.N = Nc;
    gc.B = Bfull + c * Nc * g.ldb;
    if (chunks > 1) gc.b_buf = nullptr;
    gc.C = static_cast<float*>(out->ptr()) + c * Nc * g.M;
    mark(comm, "compute: gemm start", ctx->stream);
    k::launch_gemm(ctx, gc);
    mark(comm, "compute: gemm end", ctx->stream);
    cudaEvent_t ev;
    GADCU_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    GADCU_CUDA(cudaEventRecord(ev, ctx->stream));
    GADCU_CUDA(cudaStreamWaitEvent(comm->stream, ev, 0));
    GADCU_CUDA(cudaEventDestroy(ev));
    mark(comm, "comm: after gemm", comm->stream);
    nvls_allreduce(comm, comm->heap.nv_offset(slot) + size_t(c) * Nc * g.M * 4, size_t(Nc) * g.M * 4);
    mark(comm, "comm: nvls all-reduce done", comm->stream);
  }
  GADCU_CUDA(cudaEventRecord(comm->nv_slot_done[slot], comm->stream));
  comm->pending++;
  return out;
}

// The matmul VJP's product IS a variable's gradient: compute it with the tile-pushing epilogue and run the rest of
// the exchange on the communicator's stream. Returns the array that holds the rank-summed gradient after the join.
ArrayRef dp_gemm(gadcu_comm* comm, const ArrayRef& a_in, const ArrayRef& b_in, MatProp p1, MatProp p2, bool last) {
  gadcu_ctx* ctx = comm->ctx;
  if (a_in->dtype != GADCU_F32 || b_in->dtype != GADCU_F32) return ArrayRef();
  const Dim4 rd = check::matmul(a_in->dims, b_in->dims, p1, p2);
  if (rd[2] * rd[3] != 1) return ArrayRef();
  ArrayRef a = materialize(a_in), b = materialize(b_in);
  k::GemmArgs g;
  g.dt = GADCU_F32;
  g.A = a->ptr();
  g.B = b->ptr();
  g.M = rd[0];
  g.N = rd[1];
  g.K = p1.transposed ? a->dims[0] : a->dims[1];
  g.ta = p1.transposed;
  g.tb = p2.transposed;
  g.lda = a->dims[0];
  g.ldb = b->dims[0];
  g.strideC = rd[0] * rd[1];
  if (a->ptr() == a->buf->ptr) g.a_buf = a->buf.get();
  if (b->ptr() == b->buf->ptr) g.b_buf = b->buf.get();
  if (!k::gemm_push_supported(g) || comm->no_peer_memory) return ArrayRef();
  const size_t bytes = size_t(g.M) * g.N * 4;
  try {
    ensure_heap(comm, bytes);
  } catch (const GadError& e) {
    // no peer memory on this system (IPC disabled, no P2P): decline — the gradient then travels through the grouped
    // NCCL all-reduce at the end of the sweep like the small ones. Systemic, so every rank takes the same branch.
    fprintf(stderr, "gadcu: peer-memory gradient exchange unavailable (%s); using NCCL\n", e.msg.c_str());
    cudaGetLastError();
    comm->no_peer_memory = true;
    return ArrayRef();
  }
  if (comm->nvls && bytes % 16 == 0 && bytes <= comm->heap.grad_bytes) return dp_gemm_nvls(comm, g, rd, last);
  // GADCU_DP_PUSH = epilogue (default): the GEMM's epilogue stores each tile into its owner's staging buffer;
  //               = kernel: plain GEMM, then a scatter kernel on the communicator's stream
  static const bool push_from_epilogue = [] {
    const char* e = getenv("GADCU_DP_PUSH");
    return !(e && std::string(e) == "kernel");
  }();
  // The exchange of the sweep's LAST gradient has nothing left to hide behind: it is cut into column chunks so that
  // the gather of one chunk overlaps the GEMM of the next and only the last chunk's gather is exposed.
  static const uint64_t max_chunks = [] { const char* e = getenv("GADCU_DP_TAIL_CHUNKS"); return e ? uint64_t(std::max(1, atoi(e))) : uint64_t(1); }();  // measured on 8 GPUs: 2 chunks = no gain, 4 = slower (barrier latency dominates)
  uint64_t chunks = 1;
  if (last && !g.tb)
    while (chunks * 2 <= max_chunks && g.N % (chunks * 2 * 256) == 0) chunks *= 2;
  const uint64_t Nc = g.N / chunks;
  ArrayRef out = new_array(ctx, GADCU_F32, rd);
  const float* Bfull = static_cast<const float*>(g.B);
  for (uint64_t c = 0; c < chunks; c++) {
    k::GemmArgs gc = g;
    gc.N = Nc;
    gc.B = Bfull + c * Nc * g.ldb;                          // columns [c Nc, (c+1) Nc) of the stored [K, N] operand
    if (chunks > 1) gc.b_buf = nullptr;                     // (a sub-view: its TF32 split is its own)
    float* out_c = static_cast<float*>(out->ptr()) + c * Nc * g.M;
    gc.C = out_c;
    const size_t bytes_c = size_t(g.M) * Nc * 4;
    const int slot = int(comm->exchanges % kSlots);
    // the slot's previous exchange (kSlots exchanges ago) must have been copied out here before anyone may push into it
    // again: peers push only after passing a barrier that this rank enters after this wait
    if (comm->exchanges >= uint64_t(kSlots)) GADCU_CUDA(cudaStreamWaitEvent(ctx->stream, comm->slot_done[slot], 0));
    comm->exchanges++;
    k::GemmPush push;
    push.nranks = uint32_t(comm->nranks);
    push.rank = uint32_t(comm->rank);
    for (int r = 0; r < comm->nranks; r++) push.peer[r] = comm->heap.staging(r, slot);
    // a k-split GEMM pushes two partial blocks per tile: a gain up to 4 ranks (4 GPUs: 342 -> 350 steps/s), but at 8 the
    // pushes are what the NVLink-bound exchange waits for (517 -> 480), so there the GEMM keeps whole tiles
    push.splits = push_from_epilogue && comm->nranks <= 4 ? k::gemm_push_splits(ctx, gc) : 1;
    if (push_from_epilogue) gc.push = &push;
    mark(comm, "compute: gemm start", ctx->stream);
    k::launch_gemm(ctx, gc);  // compute stream; with the pushing epilogue the partial tiles land in their owners' staging buffers
    mark(comm, "compute: gemm end", ctx->stream);
    cudaEvent_t ev;
    GADCU_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    GADCU_CUDA(cudaEventRecord(ev, ctx->stream));
    GADCU_CUDA(cudaStreamWaitEvent(comm->stream, ev, 0));
    GADCU_CUDA(cudaEventDestroy(ev));
    // communicator stream, overlapping the rest of the sweep: all partials have arrived -> reduce + publish -> all
    // results have arrived -> copy out
    mark(comm, "comm: after gemm", comm->stream);
    const uint32_t num_mt = uint32_t(g.M / 256), tiles = num_mt * uint32_t(Nc / 256);
    if (!push_from_epilogue) {
      PeerStaging st;
      for (int r = 0; r < 8; r++) st.p[r] = push.peer[r];
      const unsigned grid = unsigned(std::min<uint64_t>(uint64_t(tiles) * 64, uint64_t(ctx->sm_count) * 4));
      dp_scatter_kernel<<<grid, 256, 0, comm->stream>>>(out_c, st, uint32_t(g.M), num_mt, tiles, uint32_t(comm->nranks), uint32_t(comm->rank));
      ctx->count_launch();
      mark(comm, "comm: scattered", comm->stream);
    }
    barrier_on(comm, comm->stream);
    mark(comm, "comm: barrier 1 done", comm->stream);
    PeerResults res;
    for (int r = 0; r < 8; r++) res.p[r] = r < comm->nranks ? comm->heap.result(r, slot) : nullptr;
    const uint32_t owned = tiles > uint32_t(comm->rank) ? (tiles - comm->rank + comm->nranks - 1) / comm->nranks : 0;
    if (owned > 0) {
      const unsigned grid = unsigned(std::min<uint64_t>(uint64_t(owned) * 64, uint64_t(ctx->sm_count) * 4));
      dp_reduce_publish_kernel<<<grid, 256, 0, comm->stream>>>(reinterpret_cast<const float4*>(comm->heap.staging(comm->rank, slot)), res,
                                                               uint32_t(g.M), num_mt, tiles, uint32_t(comm->nranks), uint32_t(comm->rank),
                                                               uint32_t(comm->nranks) * push.splits);
      ctx->count_launch();
    }
    mark(comm, "comm: reduce+publish done", comm->stream);
    barrier_on(comm, comm->stream);
    mark(comm, "comm: barrier 2 done", comm->stream);
    GADCU_CUDA(cudaMemcpyAsync(out_c, comm->heap.result(comm->rank, slot), bytes_c, cudaMemcpyDeviceToDevice, comm->stream));
    mark(comm, "comm: copy out done", comm->stream);
    GADCU_CUDA(cudaEventRecord(comm->slot_done[slot], comm->stream));
  }
  comm->pending++;
  return out;
}

// One array's in-place sum over ranks on the communicator's own stream, ordered after everything already issued on
// the compute stream (the kernels that produced the array) but not before what is issued afterwards.
void allreduce_async(gadcu_comm* comm, const ArrayRef& a_in) {
  ArrayRef a = a_in;
  if (a->symbolic()) materialize(a);
  if (a->lazy() || a->symbolic()) fail(GADCU_ERR_INVALID, "allreduce needs dense arrays");
  gadcu_ctx* ctx = comm->ctx;
  lazy_before_write(ctx, a->ptr());
  a->buf->mutated();
  cudaEvent_t ev;
  GADCU_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
  GADCU_CUDA(cudaEventRecord(ev, ctx->stream));
  GADCU_CUDA(cudaStreamWaitEvent(comm->stream, ev, 0));
  GADCU_CUDA(cudaEventDestroy(ev));
  nccl_check(nccl().AllReduce(a->ptr(), a->ptr(), a->elements(), a->dtype == GADCU_F32 ? ncclFloat32 : ncclFloat64, ncclSum, comm->comm,
                              comm->stream),
             "ncclAllReduce");
  comm->pending++;
  ctx->comm_inflight.store(true, std::memory_order_relaxed);
  ctx->count_launch();
}

// The compute stream waits (on the device) for every async collective issued so far.
void join(gadcu_comm* comm) {
  if (comm->pending == 0) return;
  gadcu_ctx* ctx = comm->ctx;
  cudaEvent_t ev;
  GADCU_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
  GADCU_CUDA(cudaEventRecord(ev, comm->stream));
  GADCU_CUDA(cudaStreamWaitEvent(ctx->stream, ev, 0));
  GADCU_CUDA(cudaEventDestroy(ev));
  comm->pending = 0;
  ctx->comm_inflight.store(false, std::memory_order_relaxed);
}

}  // namespace

gadcu_status gadcu_comm_allreduce_sum_async(gadcu_comm* comm, gadcu_array* const* arrays, size_t n) {
  return guard([&] {
    if (comm->nranks <= 1) return;
    comm->ctx->segment_flush();
    for (size_t i = 0; i < n; i++) allreduce_async(comm, ArrayRef(arrays[i], true));
  });
}
gadcu_status gadcu_comm_join(gadcu_comm* comm) {
  return guard([&] {
    if (comm->nranks > 1) join(comm);
  });
}

// Graph1::evaluate_gradients[_once] for a batch shard: each variable's gradient is all-reduced (sum over ranks) as
// soon as the sweep has completed it — the last layer's weight gradient travels over NVLink while the layers below
// are still being differentiated — and the compute stream waits for all of them before the call returns.
gadcu_status gadcu_evaluate_gradients_dp(gadcu_algebra* g, uint32_t arena, uint32_t index, gadcu_array* seed, int once, gadcu_comm* comm,
                                         gadcu_store** out) {
  return gadcu_evaluate_gradients_dp_with(g, arena, index, seed, once, comm, nullptr, 0, out);
}
gadcu_status gadcu_evaluate_gradients_dp_with(gadcu_algebra* g, uint32_t arena, uint32_t index, gadcu_array* seed, int once, gadcu_comm* comm,
                                              gadcu_array* const* also, size_t n_also, gadcu_store** out) {
  return guard([&] {
    if (index == 0) fail(GADCU_ERR_MISSING_ID, "Trying to obtain `id` of a constant value.");
    auto* graph = dynamic_cast<GraphAlgebra*>(g);
    if (!graph || graph->kind() != GADCU_GRAPH1) fail(GADCU_ERR_INVALID, "evaluate_gradients_dp needs a Graph1 algebra");
    const bool dp = comm && comm->nranks > 1;
    const bool was_auto = g->ctx->auto_replay;
    if (dp) {  // the exchange interleaves streams and cross-rank barriers with the sweep: it is issued eagerly
      g->ctx->segment_flush();
      g->ctx->auto_replay = false;
    }
    struct Restore { gadcu_ctx* c; bool v; ~Restore() { c->auto_replay = v; } } restore{g->ctx, was_auto};
    // GADCU_DP = fused (default): GEMM-produced gradients are reduce-scattered by the GEMM epilogue over peer memory
    // and gathered on the communicator's stream while the sweep continues; everything else (biases) goes through
    // one grouped NCCL all-reduce at the end. GADCU_DP = nccl-overlap: every gradient through NCCL as soon as final.
    static const int mode = [] {
      const char* e = getenv("GADCU_DP");
      return e && std::string(e) == "nccl-overlap" ? 1 : 0;
    }();
    if (dp) {
      // Every variable's gradient is reduced IN PLACE exactly once. ADD / SUB / ADDC / ADD_ALL hand the same array to
      // several inputs (core.rs:227-232): the second id shares the first one's reduction instead of being summed over
      // the ranks twice. A gradient whose buffer the caller can see (the seed itself, when the root is a variable; a
      // lazy broadcast of it) is reduced in a private copy that replaces the store entry.
      const Buffer* seed_buf = seed && seed->buf ? seed->buf.get() : nullptr;
      graph->set_final_hook([comm, seed_buf](Id, const ArrayRef& grad_in) -> ArrayRef {
        ArrayRef grad = grad_in;
        if (grad->symbolic()) materialize(grad);
        for (auto& r : comm->replaced)
          if (r.first == grad.get()) return r.second;  // the same array under another id: shares the copy's reduction
        for (auto& q : comm->deferred) {
          if (q.get() == grad.get()) return ArrayRef();                       // same array under another id: queued already
          if (q->buf.get() == grad->buf.get() && q->dims == grad->dims && !q->lazy() && !grad->lazy()) return q;
        }
        ArrayRef target = grad;
        bool shared_buffer = grad->lazy() || grad->buf.get() == seed_buf;
        for (auto& q : comm->deferred) shared_buffer = shared_buffer || q->buf.get() == grad->buf.get();  // (a view with other dims)
        if (shared_buffer) {
          ArrayRef dense = materialize(grad);
          target = new_array(comm->ctx, dense->dtype, dense->dims, dense->is_scalar);
          GADCU_CUDA(cudaMemcpyAsync(target->ptr(), dense->ptr(), dense->elements() * dense->elem_size(), cudaMemcpyDeviceToDevice, comm->ctx->stream));
        }
        comm->deferred.push_back(target);
        if (target.get() != grad.get()) comm->replaced.emplace_back(grad.get(), target);
        if (mode == 1) allreduce_async(comm, target);
        return target.get() == grad_in.get() ? ArrayRef() : target;
      });
      if (mode != 1)
        graph->set_gemm_sink([comm](Id, const ArrayRef& lhs, const ArrayRef& rhs, MatProp pl, MatProp pr, bool last) { return dp_gemm(comm, lhs, rhs, pl, pr, last); });
    }
    auto finish = [&] {
      graph->set_final_hook(nullptr);
      graph->set_gemm_sink(nullptr);
      if (!dp) return;
      comm->replaced.clear();
      if (mode == 1) comm->deferred.clear();  // (already on their way, one all-reduce each)
      for (size_t i = 0; i < n_also; i++) comm->deferred.emplace_back(also[i], true);
      if (!comm->deferred.empty()) {
        std::vector<gadcu_array*> raw;
        for (auto& a : comm->deferred) raw.push_back(a.get());
        mark(comm, "compute: before nccl(deferred)", comm->ctx->stream);
        gadcu_status st = gadcu_comm_allreduce_sum(comm, raw.data(), raw.size());
        mark(comm, "compute: after nccl(deferred)", comm->ctx->stream);
        comm->deferred.clear();
        if (st != GADCU_OK) fail(st, last_error());
      }
      join(comm);
    };
    std::unique_ptr<Store> store;
    if (dp) mark(comm, "compute: sweep start", comm->ctx->stream);
    try {
      store = graph->evaluate_gradients(Id{arena, index}, ArrayRef(seed, true), once != 0);
    } catch (...) {
      if (comm) comm->deferred.clear();  // (comm may be NULL: the one-process form of the call)
      try { finish(); } catch (...) {}
      throw;
    }
    if (dp) mark(comm, "compute: sweep issued", comm->ctx->stream);
    finish();
    if (dp) mark(comm, "compute: joined", comm->ctx->stream);
    if (dp) dump_trace(comm);
    *out = store.release();
  });
}

// GraphN::compute_gradients (graph.rs:307-318) of one batch shard. The sweep records the backward pass on this rank's
// tape as usual; every VARIABLE's gradient is, as soon as the sweep has completed it, copied and summed over the ranks on
// the communicator's stream — overlapping the rest of the sweep — and the store hands out that sum as a CONSTANT (the
// rank-summed array is not a node of this rank's tape; the recorded local gradient stays what later differentiation
// sees). Joined before returning. comm == NULL or one rank: identical to gadcu_compute_gradients.
gadcu_status gadcu_compute_gradients_dp(gadcu_algebra* g, uint32_t arena, uint32_t index, const gadcu_value* seed, gadcu_comm* comm,
                                        gadcu_store** out) {
  return guard([&] {
    if (index == 0) fail(GADCU_ERR_MISSING_ID, "Trying to obtain `id` of a constant value.");
    auto* graph = dynamic_cast<GraphAlgebra*>(g);
    if (!graph || graph->kind() != GADCU_GRAPHN) fail(GADCU_ERR_INVALID, "compute_gradients_dp needs a GraphN algebra");
    if (!seed) fail(GADCU_ERR_INVALID, "null seed");
    const Value sv(ArrayRef(seed->data, true), Id{seed->arena, seed->index});
    const bool dp = comm && comm->nranks > 1;
    const bool was_auto = g->ctx->auto_replay;
    if (dp) {
      g->ctx->segment_flush();
      g->ctx->auto_replay = false;
    }
    struct Restore { gadcu_ctx* c; bool v; ~Restore() { c->auto_replay = v; } } restore{g->ctx, was_auto};
    if (dp)
      graph->set_final_hook([comm](Id, const ArrayRef& grad_in) -> ArrayRef {
        ArrayRef dense = materialize(grad_in);
        ArrayRef sum = new_array(comm->ctx, dense->dtype, dense->dims, dense->is_scalar);
        GADCU_CUDA(cudaMemcpyAsync(sum->ptr(), dense->ptr(), dense->elements() * dense->elem_size(), cudaMemcpyDeviceToDevice, comm->ctx->stream));
        allreduce_async(comm, sum);
        return sum;
      });
    std::unique_ptr<Store> store;
    try {
      store = graph->compute_gradients(Id{arena, index}, sv);
    } catch (...) {
      graph->set_final_hook(nullptr);
      if (dp) { try { join(comm); } catch (...) {} }
      throw;
    }
    graph->set_final_hook(nullptr);
    if (dp) join(comm);
    *out = store.release();
  });
}

gadcu_status gadcu_comm_destroy(gadcu_comm* comm) {
  return guard([&] {
    if (!comm) return;
    if (comm->stream) { cudaStreamSynchronize(comm->stream); cudaStreamDestroy(comm->stream); }
    cudaStreamSynchronize(comm->ctx->stream);
    for (auto& b : comm->nv_slot_buf) {  // arrays that outlive the communicator take their data along
      if (b && b->refs.load() > 1) {
        std::shared_ptr<Pool> pool = comm->ctx->pool;
        void* moved = pool->take(b->bytes);
        cudaMemcpy(moved, b->ptr, b->bytes, cudaMemcpyDeviceToDevice);
        b->ptr = moved;
        b->pool = pool;
      }
      b.reset();
    }
    if (comm->heap.local) gadcu::symm_free(comm->heap.region);
    for (auto& e : comm->slot_done)
      if (e) cudaEventDestroy(e);
    for (auto& e : comm->nv_slot_done)
      if (e) cudaEventDestroy(e);
    if (comm->comm) nccl().CommDestroy(comm->comm);
    delete comm;
  });
}

}  // extern "C"
