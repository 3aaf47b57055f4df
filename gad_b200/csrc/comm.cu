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

namespace {

typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclFloat32 = 7, ncclFloat64 = 8 };
enum { ncclSum = 0 };

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

// ---- symmetric peer memory for the fused GEMM -> reduce-scatter -> all-gather exchange ---------------------
// One cudaMalloc per rank, opened by every other rank through CUDA IPC (handles travel through an NCCL all-gather):
//   [ flags: 2 x 32 x u32 | small: 2 x 8 x kSmallBytes | staging: kSlots x slot_bytes | result: kSlots x slot_bytes ]
// staging[slot] receives, for every tile this rank owns, the nranks partial tiles (pushed by the GEMM epilogues of all
// ranks); result[slot] receives the reduced gradient in its final column-major layout (pushed by every owner);
// small[parity][r] receives rank r's copy of a batch of small arrays (biases, the loss) to be summed locally.
// Barriers issued on the communicator's stream and on the compute stream count separately (flag set 0 / 1).
constexpr int kSlots = 3;
constexpr size_t kSmallBytes = 256 * 1024;
constexpr size_t kHeapHeader = 256 + 2 * 8 * kSmallBytes;
struct SymmHeap {
  size_t slot_bytes = 0;
  void* local = nullptr;
  void* peer[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};  // peer[rank] == local
  uint32_t* flags(int r, int set) const { return static_cast<uint32_t*>(peer[r]) + 32 * set; }
  float* small(int r, int parity, int src) const {
    return reinterpret_cast<float*>(static_cast<char*>(peer[r]) + 256 + (size_t(parity) * 8 + size_t(src)) * kSmallBytes);
  }
  float* staging(int r, int slot) const { return reinterpret_cast<float*>(static_cast<char*>(peer[r]) + kHeapHeader + size_t(slot) * slot_bytes); }
  float* result(int r, int slot) const { return reinterpret_cast<float*>(static_cast<char*>(peer[r]) + kHeapHeader + size_t(kSlots + slot) * slot_bytes); }
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

void ensure_heap(gadcu_comm* comm, size_t bytes) {
  if (comm->heap.slot_bytes >= bytes) return;
  if (comm->heap.local) {  // sized by the first gradient that came through; a larger one later goes through NCCL
    throw GadError(GADCU_ERR_UNSUPPORTED, "gradient larger than the symmetric heap");
  }
  if (comm->nranks > 8) fail(GADCU_ERR_UNSUPPORTED, "data-parallel exchange: at most 8 ranks");
  gadcu_ctx* ctx = comm->ctx;
  Nccl& nc = nccl();
  if (!nc.AllGather) fail(GADCU_ERR_NCCL, "ncclAllGather is not available");
  SymmHeap& h = comm->heap;
  h.slot_bytes = (bytes + 255) / 256 * 256;
  const size_t total = kHeapHeader + size_t(2 * kSlots) * h.slot_bytes;
  GADCU_CUDA(cudaMalloc(&h.local, total));
  GADCU_CUDA(cudaMemsetAsync(h.local, 0, 256, ctx->stream));
  // exchange the IPC handles through NCCL (every rank calls this at the same point of the same step)
  cudaIpcMemHandle_t mine;
  GADCU_CUDA(cudaIpcGetMemHandle(&mine, h.local));
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  char* dev = nullptr;
  GADCU_CUDA(cudaMalloc(&dev, 64 * size_t(comm->nranks + 1)));
  GADCU_CUDA(cudaMemcpyAsync(dev, &mine, 64, cudaMemcpyHostToDevice, ctx->stream));
  nccl_check(nc.AllGather(dev, dev + 64, 64, /*ncclUint8*/ 1, comm->comm, ctx->stream), "ncclAllGather");
  std::vector<cudaIpcMemHandle_t> all(comm->nranks);
  GADCU_CUDA(cudaMemcpyAsync(all.data(), dev + 64, 64 * size_t(comm->nranks), cudaMemcpyDeviceToHost, ctx->stream));
  GADCU_CUDA(cudaStreamSynchronize(ctx->stream));
  GADCU_CUDA(cudaFree(dev));
  for (int r = 0; r < comm->nranks; r++) {
    if (r == comm->rank) { h.peer[r] = h.local; continue; }
    GADCU_CUDA(cudaIpcOpenMemHandle(&h.peer[r], all[r], cudaIpcMemLazyEnablePeerAccess));
  }
  for (int s = 0; s < kSlots; s++) GADCU_CUDA(cudaEventCreateWithFlags(&comm->slot_done[s], cudaEventDisableTiming));
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
    ensure_heap(comm, 2 * bytes);  // (a k-split GEMM pushes two partial blocks per tile and rank)
  } catch (const GadError& e) {
    // no peer memory on this system (IPC disabled, no P2P): decline — the gradient then travels through the grouped
    // NCCL all-reduce at the end of the sweep like the small ones. Systemic, so every rank takes the same branch.
    fprintf(stderr, "gadcu: peer-memory gradient exchange unavailable (%s); using NCCL\n", e.msg.c_str());
    cudaGetLastError();
    comm->no_peer_memory = true;
    return ArrayRef();
  }
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
      comm->deferred.clear();
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
    for (int r = 0; r < comm->nranks; r++)
      if (comm->heap.peer[r] && r != comm->rank) cudaIpcCloseMemHandle(comm->heap.peer[r]);
    if (comm->heap.local) cudaFree(comm->heap.local);
    for (auto& e : comm->slot_done)
      if (e) cudaEventDestroy(e);
    if (comm->comm) nccl().CommDestroy(comm->comm);
    delete comm;
  });
}

}  // extern "C"
