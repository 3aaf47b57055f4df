// runtime.cu — context, caching allocator, arrays, CUDA-graph capture.
#include <cstring>

#include "algebra.h"
#include "batched.h"
#include "kernels.h"

namespace gadcu {

static thread_local std::string g_last_error;
void set_last_error(const std::string& msg) { g_last_error = msg; }
const std::string& last_error() { return g_last_error; }
void fail(int status, const std::string& msg) { throw GadError(status, msg); }

std::string Dim4::str() const {
  return "[" + std::to_string(d[0]) + " " + std::to_string(d[1]) + " " + std::to_string(d[2]) + " " + std::to_string(d[3]) + "]";
}

// ---- pool ---------------------------------------------------------------------------------------
Pool::~Pool() {
  for (auto& kv : free_blocks) cudaFree(kv.second);
}
void* Pool::take(size_t bytes) {
  bytes = round_up(bytes);
  {
    std::lock_guard<std::mutex> lk(mu);
    auto it = free_blocks.find(bytes);
    if (it != free_blocks.end()) {
      void* p = it->second;
      free_blocks.erase(it);
      in_use += bytes;
      return p;
    }
  }
  void* p = nullptr;
  cudaError_t err = cudaMalloc(&p, bytes);
  if (err != cudaSuccess) {
    // release cached blocks and retry once
    {
      std::lock_guard<std::mutex> lk(mu);
      for (auto& kv : free_blocks) { cudaFree(kv.second); reserved -= kv.first; }
      free_blocks.clear();
    }
    cudaGetLastError();
    err = cudaMalloc(&p, bytes);
    if (err != cudaSuccess) fail(GADCU_ERR_CUDA, std::string("cudaMalloc(") + std::to_string(bytes) + "): " + cudaGetErrorString(err));
  }
  std::lock_guard<std::mutex> lk(mu);
  reserved += bytes;
  in_use += bytes;
  return p;
}
void Pool::put(void* p, size_t bytes) {
  bytes = round_up(bytes);
  std::lock_guard<std::mutex> lk(mu);
  free_blocks.emplace(bytes, p);
  in_use -= bytes;
}

}  // namespace gadcu

void gadcu_ctx::segment_begin() {
  if (!auto_replay || capturing || profiling || segment_open) return;
  if (cudaStreamBeginCapture(stream, cudaStreamCaptureModeRelaxed) != cudaSuccess) {
    cudaGetLastError();
    return;  // (issue eagerly)
  }
  segment_open = true;
  segment_launches = 0;
}

void gadcu_ctx::segment_flush() {
  if (!segment_open) return;
  segment_open = false;
  cudaGraph_t graph = nullptr;
  const cudaError_t err = cudaStreamEndCapture(stream, &graph);
  if (err != cudaSuccess || !graph) {
    cudaGetLastError();
    gadcu::fail(GADCU_ERR_CUDA, std::string("auto replay: the recorded segment was invalidated (") + cudaGetErrorString(err) +
                                    "); the work issued since the last flush is lost");
  }
  size_t nodes = 0;
  cudaGraphGetNodes(graph, nullptr, &nodes);
  if (nodes == 0) {
    cudaGraphDestroy(graph);
    return;
  }
  replay_segments++;
  cudaGraphExec_t exec = nullptr;
  auto range = replay_execs.equal_range(nodes);
  for (auto it = range.first; it != range.second; ++it) {
    cudaGraphExecUpdateResultInfo info;
    if (cudaGraphExecUpdate(it->second, graph, &info) == cudaSuccess) {  // same kernels in the same order: only parameters change
      exec = it->second;
      replay_hits++;
      break;
    }
    cudaGetLastError();
  }
  if (!exec) {
    if (replay_execs.size() >= 64) {  // a program that keeps inventing step shapes: start over
      for (auto& kv : replay_execs) cudaGraphExecDestroy(kv.second);
      replay_execs.clear();
    }
    const cudaError_t ie = cudaGraphInstantiate(&exec, graph, 0);
    if (ie != cudaSuccess) {
      cudaGraphDestroy(graph);
      gadcu::fail(GADCU_ERR_CUDA, std::string("auto replay: cudaGraphInstantiate: ") + cudaGetErrorString(ie));
    }
    replay_execs.emplace(nodes, exec);
    replay_misses++;
  }
  const cudaError_t le = cudaGraphLaunch(exec, stream);
  cudaGraphDestroy(graph);
  if (le != cudaSuccess) gadcu::fail(GADCU_ERR_CUDA, std::string("auto replay: cudaGraphLaunch: ") + cudaGetErrorString(le));
}

namespace gadcu {

std::recursive_mutex& api_mutex() {
  static std::recursive_mutex m;
  return m;
}

void Buffer::release() {
  if (refs.fetch_sub(1, std::memory_order_acq_rel) == 1) {
    if (split) split->release();
    if (pool && ptr) pool->put(ptr, bytes);
    delete this;
  }
}
void Buffer::mutated() {
  if (split) {
    split->release();
    split = nullptr;
  }
}

ArrayRef new_array(gadcu_ctx* ctx, gadcu_dtype dt, const Dim4& dims, bool is_scalar) {
  auto* buf = new Buffer;
  Ref<Buffer> bref(buf, false);
  buf->ctx = ctx;
  buf->bytes = dims.elements() * dtype_size(dt);
  buf->pool = ctx->active_pool();
  buf->ptr = buf->pool->take(buf->bytes);
  auto* a = new gadcu_array;
  a->buf = bref;
  a->dims = dims;
  a->phys = dims;
  a->dtype = dt;
  a->is_scalar = is_scalar;
  a->ctx = ctx;
  return ArrayRef(a, false);
}

ArrayRef view_array(const ArrayRef& src_in, const Dim4& dims, const Dim4& phys, bool is_scalar) {
  ArrayRef src = src_in->symbolic() ? materialize(src_in) : src_in;  // a view needs bytes to point at
  auto* a = new gadcu_array;
  a->buf = src->buf;
  a->dims = dims;
  a->phys = phys;
  a->dtype = src->dtype;
  a->is_scalar = is_scalar;
  a->ctx = src->ctx;
  return ArrayRef(a, false);
}

ArrayRef clone_handle(const ArrayRef& src) {
  auto* a = new gadcu_array;
  a->buf = src->buf;
  a->dims = src->dims;
  a->phys = src->phys;
  a->dtype = src->dtype;
  a->is_scalar = src->is_scalar;
  a->ctx = src->ctx;
  a->sym = src->sym;
  a->sym_reg = src->sym_reg;
  a->sym_direct = src->sym_direct;
  return ArrayRef(a, false);
}

ArrayRef materialize(const ArrayRef& a) {
  if (a->symbolic()) return force_symbolic(a);
  if (!a->lazy()) return a;
  ArrayRef out = new_array(a->ctx, a->dtype, a->dims, a->is_scalar);
  if (a->bcast1()) {
    k::EwArgs e;
    e.dt = a->dtype; e.nin = 1; e.in[0] = a->ptr(); e.bcast[0] = true; e.out = out->ptr(); e.n = a->elements();
    k::launch_ew(a->ctx, k::EwOp::COPY, e);
  } else {
    k::launch_tile(a->ctx, a->dtype, a->ptr(), a->phys, out->ptr(), a->dims);
  }
  return out;
}

}  // namespace gadcu

using namespace gadcu;

struct gadcu_capture {
  gadcu_ctx* ctx = nullptr;
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t exec = nullptr;
  std::shared_ptr<Pool> pool;
  uint64_t launches = 0;
};

extern "C" {

const char* gadcu_last_error(void) { return gadcu::last_error().c_str(); }
const char* gadcu_version(void) { return "gadcu 0.1 (sm_100a)"; }
size_t gadcu_debug_gemm_schedule(int sm_count, uint64_t M, uint64_t N, uint64_t K, int push_splits, uint32_t* rows6, size_t cap, uint32_t* info8) {
  return gadcu::k::gemm_schedule_debug(sm_count, M, N, K, push_splits, rows6, cap, info8);
}

gadcu_status gadcu_ctx_create(int device, gadcu_ctx** out) {
  return guard([&] {
    GADCU_CUDA(cudaSetDevice(device));
    auto* ctx = new gadcu_ctx;
    ctx->device = device;
    GADCU_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    GADCU_CUDA(cudaStreamCreateWithFlags(&ctx->side_stream, cudaStreamNonBlocking));
    cudaDeviceProp prop;
    GADCU_CUDA(cudaGetDeviceProperties(&prop, device));
    ctx->sm_count = prop.multiProcessorCount;
    ctx->pool = std::make_shared<Pool>();
    if (const char* e = getenv("GADCU_EPILOGUE")) ctx->gemm_epilogue = atoi(e) < 0 ? 0 : (atoi(e) > 2 ? 2 : atoi(e));  // 0 never, 1 where it pays, 2 always
    GADCU_CUDA(cudaMallocHost(&ctx->pinned_scratch, 4096));
    GADCU_CUDA(cudaMalloc(&ctx->gemm_flags, 4096));  // (fewer split tiles than SM pairs: 8 x 74 words)
    GADCU_CUDA(cudaMemset(ctx->gemm_flags, 0, 4096));
    *out = ctx;
  });
}

gadcu_status gadcu_ctx_destroy(gadcu_ctx* ctx) {
  return guard([&] {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    try { ctx->segment_flush(); } catch (...) {}
    for (auto& kv : ctx->replay_execs) cudaGraphExecDestroy(kv.second);
    ctx->replay_execs.clear();
    cudaStreamSynchronize(ctx->stream);
    cudaStreamSynchronize(ctx->side_stream);
    for (auto& d : ctx->downloads) cudaEventDestroy(d.second);
    ctx->downloads.clear();
    if (ctx->pinned_scratch) cudaFreeHost(ctx->pinned_scratch);
    if (ctx->dev_scratch) cudaFree(ctx->dev_scratch);
    if (ctx->gemm_flags) cudaFree(ctx->gemm_flags);
    cudaStreamDestroy(ctx->stream);
    cudaStreamDestroy(ctx->side_stream);
    delete ctx;
  });
}

gadcu_status gadcu_ctx_synchronize(gadcu_ctx* ctx) {
  return guard([&] {
    ctx->segment_flush();
    GADCU_CUDA(cudaStreamSynchronize(ctx->stream));
  });
}
void* gadcu_ctx_stream(gadcu_ctx* ctx) {
  guard([&] { ctx->segment_flush(); });  // whoever takes the raw stream sees everything issued so far on it
  return ctx->stream;
}
gadcu_status gadcu_ctx_set_auto_replay(gadcu_ctx* ctx, int enable) {
  return guard([&] {
    ctx->segment_flush();
    ctx->auto_replay = enable != 0;
  });
}
gadcu_status gadcu_ctx_flush(gadcu_ctx* ctx) {
  return guard([&] { ctx->segment_flush(); });
}
gadcu_status gadcu_ctx_replay_stats(gadcu_ctx* ctx, uint64_t* segments, uint64_t* reused, uint64_t* instantiated) {
  if (segments) *segments = ctx->replay_segments;
  if (reused) *reused = ctx->replay_hits;
  if (instantiated) *instantiated = ctx->replay_misses;
  return GADCU_OK;
}
gadcu_status gadcu_ctx_set_fusion(gadcu_ctx* ctx, int fuse) {
  ctx->fuse = fuse != 0;
  return GADCU_OK;
}
gadcu_status gadcu_ctx_set_lazy(gadcu_ctx* ctx, int enable, uint64_t min_elements) {
  ctx->lazy = enable != 0;
  ctx->lazy_threshold = min_elements < 2 ? 2 : min_elements;  // single elements (gad's scalars) always run eagerly
  return GADCU_OK;
}
gadcu_status gadcu_ctx_set_gemm_epilogue(gadcu_ctx* ctx, int enable) {
  ctx->gemm_epilogue = enable < 0 ? 0 : (enable > 2 ? 2 : enable);
  return GADCU_OK;
}
gadcu_status gadcu_ctx_set_strict_reference(gadcu_ctx* ctx, int strict) {
  ctx->strict_reference = strict != 0;
  return GADCU_OK;
}
gadcu_status gadcu_ctx_memory_stats(gadcu_ctx* ctx, uint64_t* in_use, uint64_t* reserved, uint64_t* launches) {
  std::lock_guard<std::mutex> lk(ctx->pool->mu);
  if (in_use) *in_use = ctx->pool->in_use;
  if (reserved) *reserved = ctx->pool->reserved;
  if (launches) *launches = ctx->launches.load();
  return GADCU_OK;
}

gadcu_status gadcu_ctx_profile_begin(gadcu_ctx* ctx) {
  return guard([&] {
    ctx->segment_flush();  // (per-kernel event timing and recorded segments exclude each other: profiled stretches run eagerly)
    for (auto& v : ctx->prof_events) {
      for (auto& e : v) { cudaEventDestroy(e.first); cudaEventDestroy(e.second); }
      v.clear();
    }
    ctx->profiling = true;
  });
}
gadcu_status gadcu_ctx_profile_end(gadcu_ctx* ctx, double* ms4, uint64_t* count4) {
  return guard([&] {
    ctx->profiling = false;
    GADCU_CUDA(cudaStreamSynchronize(ctx->stream));
    for (int c = 0; c < 4; c++) {
      double total = 0.0;
      for (auto& e : ctx->prof_events[c]) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e.first, e.second);
        total += ms;
        cudaEventDestroy(e.first);
        cudaEventDestroy(e.second);
      }
      if (ms4) ms4[c] = total;
      if (count4) count4[c] = ctx->prof_events[c].size();
      ctx->prof_events[c].clear();
    }
  });
}

gadcu_status gadcu_array_alloc(gadcu_ctx* ctx, gadcu_dtype dt, const gadcu_dim4* dims, gadcu_array** out) {
  return guard([&] { *out = new_array(ctx, dt, Dim4(*dims)).detach(); });
}
gadcu_status gadcu_array_from_host(gadcu_ctx* ctx, gadcu_dtype dt, const gadcu_dim4* dims, const void* host, gadcu_array** out) {
  return guard([&] {
    ArrayRef a = new_array(ctx, dt, Dim4(*dims));
    ctx->segment_flush();
    GADCU_CUDA(cudaMemcpyAsync(a->ptr(), host, a->buf->bytes, cudaMemcpyHostToDevice, ctx->stream));
    GADCU_CUDA(cudaStreamSynchronize(ctx->stream));  // `host` may be pageable and reused by the caller
    *out = a.detach();
  });
}
gadcu_status gadcu_array_upload(gadcu_array* a, const void* pinned_host, void* stream_or_null) {
  return guard([&] {
    if (a->lazy()) fail(GADCU_ERR_INVALID, "upload into a lazy view");
    cudaStream_t s = stream_or_null ? static_cast<cudaStream_t>(stream_or_null) : a->ctx->stream;
    lazy_before_write(a->ctx, a->ptr());
    a->ctx->segment_flush();
    a->buf->mutated();
    GADCU_CUDA(cudaMemcpyAsync(a->ptr(), pinned_host, a->elements() * a->elem_size(), cudaMemcpyHostToDevice, s));
  });
}
gadcu_status gadcu_array_upload_async(gadcu_array* a, const void* pinned_host) {
  return guard([&] {
    if (a->lazy()) fail(GADCU_ERR_INVALID, "upload into a lazy view");
    gadcu_ctx* ctx = a->ctx;
    lazy_before_write(ctx, a->ptr());
    ctx->segment_flush();
    cudaEvent_t ev;
    GADCU_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    GADCU_CUDA(cudaEventRecord(ev, ctx->stream));
    GADCU_CUDA(cudaStreamWaitEvent(ctx->side_stream, ev, 0));
    GADCU_CUDA(cudaEventDestroy(ev));
    a->buf->mutated();
    GADCU_CUDA(cudaMemcpyAsync(a->ptr(), pinned_host, a->elements() * a->elem_size(), cudaMemcpyHostToDevice, ctx->side_stream));
  });
}
gadcu_status gadcu_ctx_wait_uploads(gadcu_ctx* ctx) {
  return guard([&] {
    ctx->segment_flush();
    cudaEvent_t ev;
    GADCU_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    GADCU_CUDA(cudaEventRecord(ev, ctx->side_stream));
    GADCU_CUDA(cudaStreamWaitEvent(ctx->stream, ev, 0));
    GADCU_CUDA(cudaEventDestroy(ev));
  });
}
gadcu_status gadcu_array_download_async(const gadcu_array* a, void* pinned_host, uint64_t* ticket) {
  return guard([&] {
    if (!a || !pinned_host || !ticket) fail(GADCU_ERR_INVALID, "download_async: null argument");
    ArrayRef dense = materialize(ArrayRef(const_cast<gadcu_array*>(a), true));
    if (dense->elements() != a->elements()) fail(GADCU_ERR_INVALID, "download_async on a lane value of a batched tape: read its column through gadcu_batched_force");
    gadcu_ctx* ctx = a->ctx;
    ctx->segment_flush();
    GADCU_CUDA(cudaMemcpyAsync(pinned_host, dense->ptr(), dense->elements() * dense->elem_size(), cudaMemcpyDeviceToHost, ctx->stream));
    cudaEvent_t ev;
    GADCU_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    GADCU_CUDA(cudaEventRecord(ev, ctx->stream));
    ctx->downloads.emplace_back(++ctx->download_ticket, ev);
    *ticket = ctx->download_ticket;
  });
}
gadcu_status gadcu_ctx_wait_download(gadcu_ctx* ctx, uint64_t ticket) {
  return guard([&] {
    if (ticket > ctx->download_ticket) fail(GADCU_ERR_INVALID, "wait_download: no such ticket");
    // copies complete in ticket order (one stream): the newest event not after `ticket` covers all the older ones
    cudaEvent_t last = nullptr;
    while (!ctx->downloads.empty() && ctx->downloads.front().first <= ticket) {
      if (last) cudaEventDestroy(last);
      last = ctx->downloads.front().second;
      ctx->downloads.pop_front();
    }
    if (!last) return;  // retired by an earlier wait
    const cudaError_t e = cudaEventSynchronize(last);
    cudaEventDestroy(last);
    if (e != cudaSuccess) fail(GADCU_ERR_CUDA, std::string("wait_download: ") + cudaGetErrorString(e));
  });
}
gadcu_status gadcu_array_to_host(const gadcu_array* a, void* host) {
  return guard([&] {
    ArrayRef dense = materialize(ArrayRef(const_cast<gadcu_array*>(a), true));
    // (a lane value of a batched tape reports the dims of ONE tape's scalar; its column has `lanes` elements)
    if (dense->elements() != a->elements()) fail(GADCU_ERR_INVALID, "to_host on a lane value of a batched tape: read its column through gadcu_batched_force");
    a->ctx->segment_flush();
    GADCU_CUDA(cudaMemcpyAsync(host, dense->ptr(), dense->elements() * dense->elem_size(), cudaMemcpyDeviceToHost, a->ctx->stream));
    GADCU_CUDA(cudaStreamSynchronize(a->ctx->stream));
  });
}
gadcu_status gadcu_array_retain(gadcu_array* a) { if (a) a->retain(); return GADCU_OK; }
gadcu_status gadcu_array_clone(const gadcu_array* a, gadcu_array** out) {
  return guard([&] { *out = clone_handle(ArrayRef(const_cast<gadcu_array*>(a), true)).detach(); });
}
gadcu_status gadcu_array_release(gadcu_array* a) { if (a) a->release(); return GADCU_OK; }
gadcu_status gadcu_array_dims(const gadcu_array* a, gadcu_dim4* out) { *out = a->dims.c(); return GADCU_OK; }
gadcu_dtype gadcu_array_dtype(const gadcu_array* a) { return a->dtype; }
int gadcu_array_is_scalar(const gadcu_array* a) { return a->is_scalar ? 1 : 0; }
void* gadcu_array_device_ptr(const gadcu_array* a) {
  if (a->symbolic() && !a->is_scalar) {  // a deferred elementwise array value: compute it, the handle becomes dense
    if (guard([&] { materialize(ArrayRef(const_cast<gadcu_array*>(a), true)); }) != GADCU_OK) return nullptr;
  }
  if (a->lazy() || a->symbolic()) return nullptr;
  return a->ptr();
}
gadcu_status gadcu_scalar_from_host(gadcu_ctx* ctx, gadcu_dtype dt, double v, gadcu_array** out) {
  return guard([&] {
    ArrayRef a = new_array(ctx, dt, Dim4(), true);
    k::launch_fill(ctx, dt, a->ptr(), 1, v);
    *out = a.detach();
  });
}
gadcu_status gadcu_scalar_to_host(const gadcu_array* a, double* out) {
  return guard([&] {
    if (a->symbolic() || a->phys.elements() != 1) fail(GADCU_ERR_INVALID, "scalar_to_host on a non-scalar");
    gadcu_ctx* ctx = a->ctx;
    ctx->segment_flush();
    std::lock_guard<std::mutex> lk(ctx->mu);
    GADCU_CUDA(cudaMemcpyAsync(ctx->pinned_scratch, a->ptr(), a->elem_size(), cudaMemcpyDeviceToHost, ctx->stream));
    GADCU_CUDA(cudaStreamSynchronize(ctx->stream));
    *out = a->dtype == GADCU_F32 ? double(*static_cast<float*>(ctx->pinned_scratch)) : *static_cast<double*>(ctx->pinned_scratch);
  });
}
gadcu_status gadcu_value_release(gadcu_value* v) {
  if (v && v->data) { v->data->release(); v->data = nullptr; }
  return GADCU_OK;
}
gadcu_status gadcu_array_random(gadcu_ctx* ctx, gadcu_dtype dt, const gadcu_dim4* dims, uint64_t seed, int normal, double a, double b,
                                gadcu_array** out) {
  return guard([&] {
    ArrayRef arr = new_array(ctx, dt, Dim4(*dims));
    ctx->segment_begin();
    k::launch_random(ctx, dt, arr->ptr(), arr->elements(), seed, normal != 0, a, b);
    *out = arr.detach();
  });
}
// net.rs:171-175: `*self += other` after check_equal_dimensions(other, self). On an af::Array `+=` rebinds the handle
// to a fresh array (arrays are copy-on-write values): clones taken earlier — get_weights() snapshots, the array the
// user handed to set_weights, forward values a tape still holds — keep the old contents. So the sum is written in
// place only when nobody else holds the buffer; otherwise it goes to a fresh buffer and THIS handle is rebound to it.
gadcu_status gadcu_array_add_assign(gadcu_array* self, const gadcu_array* other) {
  return guard([&] {
    if (self->dims != other->dims) fail(GADCU_ERR_DIMENSIONS, "add_assign: " + other->dims.str() + " " + self->dims.str());
    if (self->dtype != other->dtype) fail(GADCU_ERR_TYPE, "add_assign: dtype mismatch");
    if (self->lazy()) fail(GADCU_ERR_INVALID, "add_assign into a lazy view");
    if (self->symbolic()) materialize(ArrayRef(self, true));
    ArrayRef other_dense = ArrayRef(const_cast<gadcu_array*>(other), true);
    if (other_dense->symbolic() || (other_dense->lazy() && !other_dense->bcast1())) other_dense = materialize(other_dense);
    // deferred elementwise values that would read `self` later are computed first, whether the bytes are about to be
    // overwritten or the handle rebound (their programs hold this handle, not a copy of its pointer)
    lazy_before_write(self->ctx, self->ptr());
    const bool shared = self->buf->refs.load() > 1;
    ArrayRef fresh;
    if (shared) fresh = new_array(self->ctx, self->dtype, self->dims, self->is_scalar);
    // a weight matrix that feeds tensor-core GEMMs carries its TF32 [hi | lo] copy: refresh it in the same pass
    if (self->dtype == GADCU_F32 && !other_dense->bcast1() && !other_dense->lazy() &&
        k::add_assign_refresh_split(self->ctx, self->buf.get(), static_cast<const float*>(other_dense->ptr()), shared ? fresh->buf.get() : self->buf.get(),
                                    self->elements())) {
      if (shared) self->buf = fresh->buf;
      return;
    }
    if (!shared) self->buf->mutated();
    k::EwArgs e;
    e.dt = self->dtype; e.nin = 2; e.in[0] = self->ptr(); e.in[1] = other_dense->ptr(); e.bcast[1] = other_dense->bcast1();
    e.out = shared ? fresh->ptr() : self->ptr();
    e.n = self->elements();
    k::launch_ew(self->ctx, k::EwOp::ADD, e);
    if (shared) self->buf = fresh->buf;
  });
}
// net.rs:177-179: `self * lambda`
gadcu_status gadcu_array_scale(const gadcu_array* a, double lambda, gadcu_array** out) {
  return guard([&] {
    ArrayRef src = materialize(ArrayRef(const_cast<gadcu_array*>(a), true));
    ArrayRef r = new_array(a->ctx, a->dtype, a->dims, a->is_scalar);
    k::EwArgs e;
    e.dt = a->dtype; e.nin = 1; e.in[0] = src->ptr(); e.out = r->ptr(); e.n = a->elements(); e.c = lambda;
    k::launch_ew(a->ctx, k::EwOp::MULC, e);
    *out = r.detach();
  });
}

// ---- capture ------------------------------------------------------------------------------------
gadcu_status gadcu_capture_begin(gadcu_ctx* ctx) {
  return guard([&] {
    if (ctx->capturing) fail(GADCU_ERR_INVALID, "capture already active");
    ctx->segment_flush();
    ctx->capture_pool = std::make_shared<Pool>();
    ctx->capturing = true;
    cudaError_t err = cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeRelaxed);
    if (err != cudaSuccess) {
      ctx->capturing = false;
      ctx->capture_pool.reset();
      fail(GADCU_ERR_CUDA, std::string("cudaStreamBeginCapture: ") + cudaGetErrorString(err));
    }
  });
}
gadcu_status gadcu_capture_end(gadcu_ctx* ctx, gadcu_capture** out) {
  return guard([&] {
    if (!ctx->capturing) fail(GADCU_ERR_INVALID, "no capture active");
    auto cap = std::make_unique<gadcu_capture>();
    cap->ctx = ctx;
    cap->pool = ctx->capture_pool;
    ctx->capturing = false;
    ctx->capture_pool.reset();
    GADCU_CUDA(cudaStreamEndCapture(ctx->stream, &cap->graph));
    GADCU_CUDA(cudaGraphInstantiate(&cap->exec, cap->graph, 0));
    size_t nodes = 0;
    cudaGraphGetNodes(cap->graph, nullptr, &nodes);
    cap->launches = nodes;
    *out = cap.release();
  });
}
gadcu_status gadcu_capture_launch(gadcu_capture* c) {
  return guard([&] {
    GADCU_CUDA(cudaGraphLaunch(c->exec, c->ctx->stream));
    c->ctx->count_launch(c->launches);
  });
}
gadcu_status gadcu_capture_destroy(gadcu_capture* c) {
  return guard([&] {
    if (!c) return;
    cudaStreamSynchronize(c->ctx->stream);
    if (c->exec) cudaGraphExecDestroy(c->exec);
    if (c->graph) cudaGraphDestroy(c->graph);
    delete c;
  });
}

}  // extern "C"
