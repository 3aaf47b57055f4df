// comm.cu — data parallelism over the batch dimension (north-star subsystem 5): one process per GPU,
// NCCL all-reduce (sum) of the weight gradients over NVLink/NVSwitch. The reference has no
// collective of any kind; the seam is DiffNet::apply_gradient_step, which sums per-example
// gradients before update_weights (gad/src/net_ext.rs:62-70) — sharding the examples across ranks
// and summing the per-rank gradient sums reproduces it up to summation order.
//
// NCCL is bound at run time (dlopen) so the library loads on hosts without it and shares the NCCL
// already mapped by the host process (e.g. torch's) instead of loading a second copy.
#include <dlfcn.h>

#include <cstring>

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

struct gadcu_comm {
  gadcu_ctx* ctx = nullptr;
  ncclComm_t comm = nullptr;
  int nranks = 1, rank = 0;
  cudaStream_t stream = nullptr;  // collectives that overlap the compute stream
  uint64_t pending = 0;           // async collectives issued since the last join
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

// In-place sum over ranks of every array, issued as one NCCL group on the context's stream (so it
// is ordered after the kernels that produced the gradients and before their consumers).
gadcu_status gadcu_comm_allreduce_sum(gadcu_comm* comm, gadcu_array* const* arrays, size_t n) {
  return guard([&] {
    if (comm->nranks <= 1) return;
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
  return guard([&] {
    if (index == 0) fail(GADCU_ERR_MISSING_ID, "Trying to obtain `id` of a constant value.");
    auto* graph = dynamic_cast<GraphAlgebra*>(g);
    if (!graph || graph->kind() != GADCU_GRAPH1) fail(GADCU_ERR_INVALID, "evaluate_gradients_dp needs a Graph1 algebra");
    const bool dp = comm && comm->nranks > 1;
    if (dp) graph->set_final_hook([comm](Id, const ArrayRef& grad) { allreduce_async(comm, grad); });
    std::unique_ptr<Store> store;
    try {
      store = graph->evaluate_gradients(Id{arena, index}, ArrayRef(seed, true), once != 0);
    } catch (...) {
      graph->set_final_hook(nullptr);
      if (dp) join(comm);
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
    if (comm->comm) nccl().CommDestroy(comm->comm);
    delete comm;
  });
}

}  // extern "C"
