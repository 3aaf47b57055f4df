// core.h — internal types of the gadcu backend: context (device, stream, caching allocator),
// refcounted device arrays, error plumbing. Everything here sits behind include/gadcu.h.
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <deque>
#include <cstdint>
#include <exception>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "gadcu.h"

namespace gadcu {

// ---- errors ------------------------------------------------------------------------------------
struct GadError : std::exception {
  int status;
  std::string msg;
  GadError(int s, std::string m) : status(s), msg(std::move(m)) {}
  const char* what() const noexcept override { return msg.c_str(); }
};
[[noreturn]] void fail(int status, const std::string& msg);
void set_last_error(const std::string& msg);

#define GADCU_CUDA(expr)                                                                        \
  do {                                                                                          \
    cudaError_t err__ = (expr);                                                                 \
    if (err__ != cudaSuccess)                                                                   \
      ::gadcu::fail(GADCU_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(err__));     \
  } while (0)

// ---- dims ---------------------------------------------------------------------------------------
struct Dim4 {
  uint64_t d[4] = {1, 1, 1, 1};
  Dim4() = default;
  Dim4(uint64_t a, uint64_t b = 1, uint64_t c = 1, uint64_t e = 1) : d{a, b, c, e} {}
  explicit Dim4(const gadcu_dim4& x) : d{x.d[0], x.d[1], x.d[2], x.d[3]} {}
  uint64_t operator[](int i) const { return d[i]; }
  uint64_t& operator[](int i) { return d[i]; }
  uint64_t elements() const { return d[0] * d[1] * d[2] * d[3]; }
  bool operator==(const Dim4& o) const { return d[0] == o.d[0] && d[1] == o.d[1] && d[2] == o.d[2] && d[3] == o.d[3]; }
  bool operator!=(const Dim4& o) const { return !(*this == o); }
  gadcu_dim4 c() const { return gadcu_dim4{{d[0], d[1], d[2], d[3]}}; }
  std::string str() const;
};

struct MatProp {
  bool transposed = false, conjugated = false;
  MatProp transpose() const { return {!transposed, conjugated}; }
};

// ---- intrusive refcounting ----------------------------------------------------------------------
template <class T>
class Ref {
 public:
  Ref() = default;
  Ref(T* p, bool add_ref) : p_(p) { if (p_ && add_ref) p_->retain(); }
  Ref(const Ref& o) : p_(o.p_) { if (p_) p_->retain(); }
  Ref(Ref&& o) noexcept : p_(o.p_) { o.p_ = nullptr; }
  ~Ref() { if (p_) p_->release(); }
  Ref& operator=(Ref o) noexcept { std::swap(p_, o.p_); return *this; }
  T* get() const { return p_; }
  T* operator->() const { return p_; }
  T& operator*() const { return *p_; }
  explicit operator bool() const { return p_ != nullptr; }
  T* detach() { T* p = p_; p_ = nullptr; return p; }  // hands the reference to the caller
  void reset() { if (p_) p_->release(); p_ = nullptr; }
 private:
  T* p_ = nullptr;
};

// ---- memory pool --------------------------------------------------------------------------------
// Exact-size caching pool (sizes rounded to 512 B). A training step allocates the same sizes every
// iteration, so after the first step every allocation is a free-list hit. Stream-ordered: all work
// of a context is issued on one stream, so a block freed by the host may be reused by the next
// launch without synchronisation. A capture owns a private pool so replayed graphs never share
// memory with later allocations.
struct Pool {
  std::mutex mu;
  std::multimap<size_t, void*> free_blocks;
  uint64_t reserved = 0, in_use = 0;
  ~Pool();
  void* take(size_t bytes);
  void put(void* p, size_t bytes);
  static size_t round_up(size_t bytes) { return bytes == 0 ? 512 : (bytes + 511) / 512 * 512; }
};

}  // namespace gadcu

// The C handle types ARE these structs.
struct gadcu_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t side_stream = nullptr;  // collectives / uploads
  std::deque<std::pair<uint64_t, cudaEvent_t>> downloads;  // gadcu_array_download_async: (ticket, recorded behind the copy), oldest first
  uint64_t download_ticket = 0;
  int sm_count = 148;
  std::shared_ptr<gadcu::Pool> pool;          // general pool
  std::shared_ptr<gadcu::Pool> capture_pool;  // non-null while capturing
  bool capturing = false;
  bool fuse = true;
  bool strict_reference = false;  // reproduce the reference's matmul VJP for transposed operands as is
  int gemm_epilogue = 1;          // deferred products with their elementwise tail in the GEMM's epilogue: 0 never, 1 where it pays, 2 always (gadcu_ctx_set_gemm_epilogue)
  std::atomic<uint64_t> launches{0};
  std::atomic<bool> comm_inflight{false};  // an overlapped collective is running: GEMMs leave it a few SMs
  std::atomic<uint64_t> epoch{1};
  // deferred elementwise regions (batched.cu): open lane programs by (element count, dtype)
  bool lazy = true;
  uint64_t lazy_threshold = 32768;  // smaller arrays run eagerly (one tiny kernel beats a specialised one)
  std::map<std::pair<uint64_t, int>, std::shared_ptr<void>> lazy_programs;  // the OPEN program per (elements, dtype): new ops are appended here
  // every program that may still have live handles (open or not, this tape's or an earlier one's): what an in-place
  // write into one of their leaf buffers has to walk (lazy_before_write). Expired entries are pruned on the way.
  std::vector<std::weak_ptr<void>> lazy_all;
  void* pinned_scratch = nullptr;  // 4 KiB pinned staging for scalar reads
  void* dev_scratch = nullptr;     // reduction partials
  size_t dev_scratch_bytes = 0;
  uint32_t* gemm_flags = nullptr;  // stream-K hand-over flags (gemm_tc.cu): 8 per split tile, zero between launches
  std::mutex mu;
  // optional per-category kernel timing with CUDA events on `stream` (bench.py's roofline figures)
  bool profiling = false;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof_events[4];

  // ---- automatic CUDA-graph replay keyed by tape structure (gadcu_ctx_set_auto_replay; runtime.cu) -----------------
  // A gad program rebuilds its tape every step (net_ext.rs:52), so there is no object to attach a captured graph to — but
  // the kernels a step issues are the same sequence every time. With auto replay on, the context's stream records what
  // is issued (a capture SEGMENT) instead of launching it; a segment ends where the host needs results or order (a read, a
  // synchronise, an upload, a collective, the end of a reverse sweep) and is then run as ONE graph launch: the segment's
  // graph is laid over an executable graph of the same shape kept from an earlier step (cudaGraphExecUpdate: same kernels,
  // same order, new pointers and scalars), or instantiated if there is none. The structural key is the captured sequence
  // itself; nothing is assumed about the program.
  bool auto_replay = false;
  bool segment_open = false;
  uint32_t segment_launches = 0;
  std::multimap<size_t, cudaGraphExec_t> replay_execs;  // by node count
  uint64_t replay_segments = 0, replay_hits = 0, replay_misses = 0;
  void segment_begin();  // before anything is issued on `stream` (no-op unless auto replay applies)
  void segment_flush();  // run what has been recorded

  std::shared_ptr<gadcu::Pool> active_pool() { return capturing && capture_pool ? capture_pool : pool; }
  void* scratch(size_t bytes);
  void count_launch(uint64_t n = 1) {
    launches.fetch_add(n, std::memory_order_relaxed);
    if (segment_open && (segment_launches += uint32_t(n)) >= 512) segment_flush();  // (a long eval-only stretch: do not sit on it)
  }
};

namespace gadcu {

enum ProfCat { PROF_GEMM = 0, PROF_EW = 1, PROF_REDUCE = 2, PROF_OTHER = 3 };
// Brackets the launches issued in its scope with a pair of events when profiling is on.
struct ProfileScope {
  gadcu_ctx* ctx;
  int cat;
  cudaEvent_t start = nullptr;
  ProfileScope(gadcu_ctx* c, int category) : ctx(c), cat(category) {
    ctx->segment_begin();  // (every launch helper opens with one of these)
    if (ctx->profiling && !ctx->capturing) {
      cudaEventCreate(&start);
      cudaEventRecord(start, ctx->stream);
    }
  }
  ~ProfileScope() {
    if (start) {
      cudaEvent_t stop;
      cudaEventCreate(&stop);
      cudaEventRecord(stop, ctx->stream);
      ctx->prof_events[cat].emplace_back(start, stop);
    }
  }
};

struct Buffer {
  std::atomic<long> refs{1};
  void* ptr = nullptr;
  size_t bytes = 0;
  gadcu_ctx* ctx = nullptr;
  std::shared_ptr<Pool> pool;  // null: not owned (external memory)
  // [hi | lo] TF32 split of this buffer's f32 contents, made by the first tensor-core GEMM that reads it
  // and reused by the other GEMMs of the same tape (every GEMM operand of the MLP step is read twice:
  // forward and one backward product). Valid for one tape epoch and until the buffer is written in place.
  Buffer* split = nullptr;
  uint64_t split_epoch = 0, split_elems = 0, split_rows = 0;
  void mutated();  // call after every in-place write: drops the cached split
  void retain() { refs.fetch_add(1, std::memory_order_relaxed); }
  void release();
};

}  // namespace gadcu

struct gadcu_array {
  std::atomic<long> refs{1};
  gadcu::Ref<gadcu::Buffer> buf;
  gadcu::Dim4 dims;        // logical dims
  gadcu::Dim4 phys;        // dims of the stored data; != dims for a lazy tile (each dims[i] % phys[i] == 0)
  gadcu_dtype dtype = GADCU_F32;
  bool is_scalar = false;  // a gad `T` living on the device (dims are `()`)
  gadcu_ctx* ctx = nullptr;
  // symbolic per-lane value of a batched scalar tape (batched.cu): no buffer until forced
  std::shared_ptr<void> sym;
  int32_t sym_reg = -1;
  bool sym_direct = false;  // `sym` is the lane program itself (batched scalar tapes: nothing to count per handle); else a per-handle record of it
  bool symbolic() const { return sym_reg >= 0 && !buf; }

  void retain() { refs.fetch_add(1, std::memory_order_relaxed); }
  void release() { if (refs.fetch_sub(1, std::memory_order_acq_rel) == 1) delete this; }
  bool lazy() const { return phys != dims; }
  bool bcast1() const { return phys.elements() == 1 && dims.elements() != 1; }  // scalar broadcast
  size_t elem_size() const { return dtype == GADCU_F32 ? 4 : 8; }
  uint64_t elements() const { return dims.elements(); }
  void* ptr() const { return buf->ptr; }
  // true when nobody else can observe this buffer: safe to update in place
  bool unique() const { return buf && refs.load() == 1 && buf->refs.load() == 1 && !lazy(); }
};

namespace gadcu {

using ArrayRef = Ref<gadcu_array>;

ArrayRef new_array(gadcu_ctx* ctx, gadcu_dtype dt, const Dim4& dims, bool is_scalar = false);
ArrayRef view_array(const ArrayRef& src, const Dim4& dims, const Dim4& phys, bool is_scalar);
ArrayRef clone_handle(const ArrayRef& src);  // a NEW handle on the same buffer (af::Array::clone): O(1), and what `+=` on one of them leaves alone
ArrayRef materialize(const ArrayRef& a);  // dense copy of a lazy tile (identity for dense arrays)
ArrayRef force_symbolic(const ArrayRef& a);  // batched.cu: run the lane program for this value

inline size_t dtype_size(gadcu_dtype dt) { return dt == GADCU_F32 ? 4 : 8; }

}  // namespace gadcu
