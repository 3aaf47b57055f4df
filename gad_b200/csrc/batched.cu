// batched.cu — batched scalar tapes (north-star subsystem 4; BASELINE config C2).
//
// The reference runs one scalar Graph1 tape at a time on the CPU (gad/src/graph.rs:172-246 over
// `T: Number`). Here the SAME tape code (GraphAlgebra) is instantiated over a symbolic evaluation
// algebra whose values are "one scalar per lane": Graph<Config1<SymAlgebra>>, the derivation
// gad/tests/symbolic.rs:93 uses for its string algebra. Every op appends one SSA instruction to a
// lane program instead of launching a kernel; the reverse sweep, run with the symbolic algebra as
// gradient algebra, appends the backward instructions through the unchanged VJP bodies. Asking for
// numbers (the variables' gradients, a forward value) dead-code-eliminates the program for those
// outputs, assigns slots by liveness, and runs ONE kernel: one thread per lane interprets the
// program over SoA columns x[i][lane] (coalesced), intermediates never reach HBM.
//
// Scalar semantics follow `impl ... for Eval` over T: Number: neg = -v (arith.rs:117-119),
// log1p = ln(1+v) (analytic.rs:203-205), powc with an i16 constant = powi (const_arith.rs:100-103),
// sigmoid = 1/(1+exp(-v)) (analytic.rs:221-223), select_argmax with None = 0 (compare.rs:168-174).
// Compiled with -fmad=false: one rounding per tape op, as on the CPU.
//
// The same machinery defers the ELEMENTWISE regions of array tapes (north-star subsystem 1): with fusion on,
// EvalAlgebra::fused records large elementwise ops — forward primitives and the fused VJP bodies alike — into a
// lane program with one lane per array element instead of launching them (`lazy_elementwise`). A value is computed
// when something needs its bytes (a reduction, a GEMM, a host read, the end of a reverse sweep for the variables'
// gradients): ONE kernel evaluates the whole dependency chain of the requested outputs, reading only the arrays
// at its leaves, so chains of VJPs across nodes cost one pass (C3 at the 4-pass algorithmic traffic),
// and saved forward values that are themselves cheap elementwise expressions are recomputed instead of re-read —
// bit-identically: same ops, same order, one rounding each.
//
// Programs run through NVRTC-compiled straight-line kernels (jit.cu) — values live in registers — with the
// interpreter below as the fallback for batched tapes when NVRTC is unavailable.
#include <cmath>
#include <algorithm>
#include <atomic>
#include <cstring>
#include <sstream>
#include <tuple>
#include <unordered_map>
#include <unordered_set>

#include "batched.h"
#include "jit.h"

namespace gadcu {
namespace {

enum SymOp : uint16_t {
  S_LOAD, S_LOADU, S_CONST, S_ADD, S_SUB, S_MUL, S_DIV, S_NEG, S_ADDC, S_MULC, S_POWI, S_POWF, S_POW, S_EXP, S_LOG, S_LOG1P,
  S_SIN, S_COS, S_TANH, S_SIGMOID, S_RECIP, S_SQRT, S_SELECT, S_SELECT_R0, S_SELECT_R1, S_MIN, S_MAX, S_STORE,
  // array semantics where they differ from the scalar `impl ... for Eval` (ew.cu functors): 0 - v (arith.rs:60-62),
  // af::log1p, af::minof / af::maxof; tiled loads (a row or a column broadcast over the other dimension)
  S_NEG0, S_LOG1PF, S_FMIN, S_FMAX, S_LOADCOL, S_LOADROW,
  // array mode: the result of a matmul that has been recorded but not launched (SymProgram::pending). A leaf like a load;
  // by the time a kernel is generated it has a column (the plain GEMM ran) or nothing needs it (the GEMM ran with the
  // elementwise ops behind it fused into its epilogue and THEIR result has the column).
  S_GEMM
};

struct SymInstr {
  uint16_t op;
  int32_t a = -1, b = -1, c = -1, d = -1;  // SSA registers (or input index for loads)
  double imm = 0.0;
};

struct DevInstr {  // what the kernel interprets: slots instead of registers
  uint16_t op, dst, a, b, c, d;
  uint32_t pad;
  double imm;
};
static_assert(sizeof(DevInstr) == 24, "DevInstr layout");

struct SymProgram {
  gadcu_ctx* ctx;
  gadcu_dtype dtype;
  uint64_t lanes;
  bool array_mode = false;   // lanes are the elements of device arrays (deferred elementwise region of an array tape)
  uint64_t epoch = 0;        // tape epoch the program was opened in (array mode)
  std::vector<SymInstr> instrs;
  std::vector<ArrayRef> inputs;                      // real arrays: [lanes] columns, 1-element uniforms, row/column tiles
  std::vector<uint64_t> input_d0;                    // S_LOADCOL / S_LOADROW: dims[0] of the logical array
  std::map<std::tuple<const void*, int, uint64_t>, int32_t> input_reg;  // dedupe loads by (buffer, kind, d0)
  std::unordered_map<int32_t, ArrayRef> cache;       // register -> materialised column
  std::vector<int32_t> live;                         // handles alive per register
  std::vector<int32_t> variable_inputs;              // batched tapes: input index of every variable, in creation order
  // batched tapes: what a finished plan needs to run again for the same outputs (planned with no column cached), so a
  // re-run at new points costs a parameter block and a launch, not a re-plan of a thousand instructions
  struct Compiled {
    const jit::Kernel* kernel = nullptr;
    std::vector<int32_t> in_inputs;  // parameter slot -> index into `inputs`
    std::vector<uint64_t> tile_d0;
    std::vector<double> imms;
    int V = 1, U = 1, block = 128;
    size_t smem_bytes = 0;  // > 0: the staged kernel (persistent blocks of block + 32 threads, see JitPlan::ring)
    int stage_blocks = 0;
    uint64_t instructions = 0, values = 0;
  };
  std::map<std::vector<int32_t>, Compiled> compiled;
  std::vector<ArrayRef> launch_compiled(const Compiled& c, size_t nout);
  std::map<std::tuple<uint16_t, int32_t, int32_t, int32_t, int32_t, double>, int32_t> value_number;  // array mode: CSE
  // array mode: deferred products (register of the S_GEMM instruction -> what to launch). The operands are held.
  struct PendingGemm {
    ArrayRef a, b;
    MatProp p1, p2;
    Dim4 rd;
    bool fusable;  // false: launched plain when first needed (higher-order tapes: see lazy_matmul)
  };
  std::unordered_map<int32_t, PendingGemm> pending;
  // columns that must outlive their handles: what a deferred product left behind (its own result, or the result of the
  // tail fused into it). "Recompute from the operands when the column is gone" would mean launching the GEMM again.
  std::unordered_set<int32_t> pinned;
  void resolve_pending(const std::vector<int32_t>& todo);
  ArrayRef launch_pending(const PendingGemm& pg, const k::GemmEpilogue* epi, bool with_split);
  uint64_t last_instructions = 0, last_slots = 0;
  bool retired = false;      // one of its leaf buffers has been overwritten in place: closed for new instructions (lazy_before_write)
  std::mutex mu;

  int32_t emit(SymInstr i) {
    if (array_mode && i.op != S_LOAD && i.op != S_LOADU && i.op != S_LOADCOL && i.op != S_LOADROW) {
      // same op on the same registers gives the same bits: the VJP's recomputed forward (analytic.rs:308,388)
      // is the forward value itself
      auto key = std::make_tuple(i.op, i.a, i.b, i.c, i.d, i.imm);
      auto it = value_number.find(key);
      if (it != value_number.end()) return it->second;
      instrs.push_back(i);
      live.push_back(0);
      value_number.emplace(key, int32_t(instrs.size() - 1));
      return int32_t(instrs.size() - 1);
    }
    instrs.push_back(i);
    if (array_mode) live.push_back(0);
    return int32_t(instrs.size() - 1);
  }
  // array mode: reduce one register over contiguous segments of `reduce` lanes instead of writing it — the cut and the
  // order inside a chunk are reduce.cu's (k::reduce_contig_plan), so the bits are those of materialise-then-reduce
  struct ReduceSpec {
    k::RedOp op;
    uint64_t reduce, outer, splits, chunk;
    bool vec;    // what reduce.cu would choose for the materialised value: reduce % V == 0
    void* out;   // [outer * splits] partials (or the result itself when splits == 1)
  };
  std::vector<ArrayRef> run(const std::vector<int32_t>& outs);
  std::vector<ArrayRef> run_interpreter(const std::vector<int32_t>& todo);
  std::vector<ArrayRef> run_jit(const std::vector<int32_t>& todo, const ReduceSpec* red = nullptr, bool* red_done = nullptr);
  bool run_reduce(int32_t reg, ReduceSpec& spec);
};

// What a symbolic handle holds: keeps the program alive and counts the handles per register, so that a write
// into one of the program's input buffers can first force exactly the values somebody can still ask for.
struct SymRef {
  std::shared_ptr<SymProgram> prog;
  int32_t reg;
  SymRef(std::shared_ptr<SymProgram> p, int32_t r) : prog(std::move(p)), reg(r) {
    std::lock_guard<std::mutex> lk(prog->mu);
    prog->live[reg]++;
  }
  ~SymRef() {
    ArrayRef drop;  // released after the lock
    std::lock_guard<std::mutex> lk(prog->mu);
    if (--prog->live[reg] == 0 && prog->array_mode && !prog->pinned.count(reg)) {
      // nobody can ask for this register any more: its column (if it was computed) goes back to the pool now;
      // a later value that depends on it recomputes it from its operands
      auto it = prog->cache.find(reg);
      if (it != prog->cache.end()) { drop = std::move(it->second); prog->cache.erase(it); }
    }
  }
};
// A batched scalar tape makes thousands of handles per recorded step and needs none of the per-register bookkeeping (no
// column is ever evicted, nothing is written in place under it): its handles hold the program itself. Only the handles of
// array-mode regions go through a SymRef.
std::shared_ptr<SymProgram> program_of(const gadcu_array* a) {
  return a->sym_direct ? std::static_pointer_cast<SymProgram>(a->sym) : static_cast<SymRef*>(a->sym.get())->prog;
}
const SymProgram* program_ptr(const gadcu_array* a) {
  return a->sym_direct ? static_cast<const SymProgram*>(a->sym.get()) : static_cast<SymRef*>(a->sym.get())->prog.get();
}

ArrayRef sym_array(const std::shared_ptr<SymProgram>& prog, int32_t reg, const Dim4& dims = Dim4(), bool is_scalar = true) {
  auto* a = new gadcu_array;
  a->dims = dims;
  a->phys = dims;
  a->dtype = prog->dtype;
  a->is_scalar = is_scalar;
  a->ctx = prog->ctx;
  if (prog->array_mode) {
    a->sym = std::make_shared<SymRef>(prog, reg);
  } else {
    a->sym = prog;
    a->sym_direct = true;
  }
  a->sym_reg = reg;
  return ArrayRef(a, false);
}

// ---- device interpreter -------------------------------------------------------------------------
template <class T> __device__ __forceinline__ T d_exp(T x);
template <> __device__ __forceinline__ float d_exp(float x) { return expf(x); }
template <> __device__ __forceinline__ double d_exp(double x) { return exp(x); }
template <class T> __device__ __forceinline__ T d_log(T x);
template <> __device__ __forceinline__ float d_log(float x) { return logf(x); }
template <> __device__ __forceinline__ double d_log(double x) { return log(x); }
template <class T> __device__ __forceinline__ T d_sin(T x);
template <> __device__ __forceinline__ float d_sin(float x) { return sinf(x); }
template <> __device__ __forceinline__ double d_sin(double x) { return sin(x); }
template <class T> __device__ __forceinline__ T d_cos(T x);
template <> __device__ __forceinline__ float d_cos(float x) { return cosf(x); }
template <> __device__ __forceinline__ double d_cos(double x) { return cos(x); }
template <class T> __device__ __forceinline__ T d_tanh(T x);
template <> __device__ __forceinline__ float d_tanh(float x) { return tanhf(x); }
template <> __device__ __forceinline__ double d_tanh(double x) { return tanh(x); }
template <class T> __device__ __forceinline__ T d_sqrt(T x);
template <> __device__ __forceinline__ float d_sqrt(float x) { return sqrtf(x); }
template <> __device__ __forceinline__ double d_sqrt(double x) { return sqrt(x); }
template <class T> __device__ __forceinline__ T d_pow(T x, T y);
template <> __device__ __forceinline__ float d_pow(float x, float y) { return powf(x, y); }
template <> __device__ __forceinline__ double d_pow(double x, double y) { return pow(x, y); }
template <class T>
__device__ __forceinline__ T d_powi(T a, int b) {  // compiler-rt __powi?f2, what Rust's powi lowers to
  const bool recip = b < 0;
  T r = T(1);
  while (true) {
    if (b & 1) r *= a;
    b /= 2;
    if (b == 0) break;
    a *= a;
  }
  return recip ? T(1) / r : r;
}

constexpr int kLaneThreads = 128;
constexpr int kChunk = 256;  // instructions staged in shared memory at a time

template <class T, int SLOTS>
__global__ void __launch_bounds__(kLaneThreads) lane_interpreter(const DevInstr* __restrict__ prog, int n,
                                                                 const T* const* __restrict__ in, T* const* __restrict__ out,
                                                                 uint64_t lanes) {
  __shared__ DevInstr stage[kChunk];
  const uint64_t lane = uint64_t(blockIdx.x) * kLaneThreads + threadIdx.x;
  const bool active = lane < lanes;
  T s[SLOTS];
  for (int base = 0; base < n; base += kChunk) {
    const int cnt = n - base < kChunk ? n - base : kChunk;
    __syncthreads();
    {  // cooperative copy of the next chunk (24-byte records as 8-byte words)
      const uint64_t* src = reinterpret_cast<const uint64_t*>(prog + base);
      uint64_t* dst = reinterpret_cast<uint64_t*>(stage);
      for (int w = threadIdx.x; w < cnt * 3; w += kLaneThreads) dst[w] = src[w];
    }
    __syncthreads();
    if (!active) continue;
    for (int pc = 0; pc < cnt; pc++) {
      const DevInstr I = stage[pc];  // same address for every lane: a shared-memory broadcast
      T r;
      switch (I.op) {
        case S_LOAD: r = in[I.a][lane]; break;
        case S_LOADU: r = in[I.a][0]; break;
        case S_CONST: r = T(I.imm); break;
        case S_ADD: r = s[I.a] + s[I.b]; break;
        case S_SUB: r = s[I.a] - s[I.b]; break;
        case S_MUL: r = s[I.a] * s[I.b]; break;
        case S_DIV: r = s[I.a] / s[I.b]; break;
        case S_NEG: r = -s[I.a]; break;
        case S_ADDC: r = s[I.a] + T(I.imm); break;
        case S_MULC: r = s[I.a] * T(I.imm); break;
        case S_POWI: r = d_powi<T>(s[I.a], int(I.imm)); break;
        case S_POWF: r = d_pow<T>(s[I.a], T(I.imm)); break;
        case S_POW: r = d_pow<T>(s[I.a], s[I.b]); break;
        case S_EXP: r = d_exp<T>(s[I.a]); break;
        case S_LOG: r = d_log<T>(s[I.a]); break;
        case S_LOG1P: r = d_log<T>(T(1) + s[I.a]); break;
        case S_SIN: r = d_sin<T>(s[I.a]); break;
        case S_COS: r = d_cos<T>(s[I.a]); break;
        case S_TANH: r = d_tanh<T>(s[I.a]); break;
        case S_SIGMOID: r = T(1) / (T(1) + d_exp<T>(-s[I.a])); break;
        case S_RECIP: r = T(1) / s[I.a]; break;
        case S_SQRT: r = d_sqrt<T>(s[I.a]); break;
        case S_SELECT: r = s[I.a] >= s[I.b] ? s[I.c] : s[I.d]; break;
        case S_SELECT_R0: r = s[I.a] >= s[I.b] ? s[I.c] : T(0); break;
        case S_SELECT_R1: r = s[I.a] >= s[I.b] ? T(0) : s[I.c]; break;
        case S_MIN: r = s[I.a] <= s[I.b] ? s[I.a] : s[I.b]; break;
        case S_MAX: r = s[I.a] >= s[I.b] ? s[I.a] : s[I.b]; break;
        case S_NEG0: r = T(0) - s[I.a]; break;
        case S_FMIN: r = fmin(s[I.a], s[I.b]); break;
        case S_FMAX: r = fmax(s[I.a], s[I.b]); break;
        case S_STORE: out[I.b][lane] = s[I.a]; continue;
        default: continue;
      }
      s[I.dst] = r;
    }
  }
}

template <class T>
void launch_interpreter(gadcu_ctx* ctx, int slots, const DevInstr* prog, int n, const void* in, void* out, uint64_t lanes) {
  ProfileScope prof_scope(ctx, PROF_OTHER);
  const unsigned grid = unsigned((lanes + kLaneThreads - 1) / kLaneThreads);
  auto in_t = static_cast<const T* const*>(in);
  auto out_t = static_cast<T* const*>(out);
  if (slots <= 32) lane_interpreter<T, 32><<<grid, kLaneThreads, 0, ctx->stream>>>(prog, n, in_t, out_t, lanes);
  else if (slots <= 128) lane_interpreter<T, 128><<<grid, kLaneThreads, 0, ctx->stream>>>(prog, n, in_t, out_t, lanes);
  else if (slots <= 512) lane_interpreter<T, 512><<<grid, kLaneThreads, 0, ctx->stream>>>(prog, n, in_t, out_t, lanes);
  else if (slots <= 2048) lane_interpreter<T, 2048><<<grid, kLaneThreads, 0, ctx->stream>>>(prog, n, in_t, out_t, lanes);
  else fail(GADCU_ERR_UNSUPPORTED, "batched tape needs more than 2048 live values per lane");
  ctx->count_launch();
}

// ---- compile (DCE + liveness slot assignment) and run -------------------------------------------
// Kernel parameters may be up to 32 764 bytes on sm_70+ (CUDA 12.1+); one pointer per input / output column, one
// double per constant. Requests with more outputs than fit are computed in several kernels.
constexpr size_t kMaxParamBytes = 32000;
// persistent blocks of a staged lane program: as many as fit an SM by shared memory (227 KB), on every SM
uint64_t staged_grid(gadcu_ctx* ctx, size_t smem_bytes, int blocks_per_sm) {
  const uint64_t fit = std::max<uint64_t>(1, (227 * 1024) / (smem_bytes + 1024));
  return uint64_t(ctx->sm_count) * std::min<uint64_t>(fit, uint64_t(blocks_per_sm));
}
constexpr size_t kMaxOutputsPerKernel = 1024;

bool use_jit() {
  static const bool interp = [] {
    const char* e = getenv("GADCU_LANES");
    return e && std::string(e) == "interp";
  }();
  return !interp && jit::available();
}

std::vector<ArrayRef> SymProgram::run(const std::vector<int32_t>& outs) {
  std::lock_guard<std::mutex> lk(mu);
  std::vector<ArrayRef> result(outs.size());
  std::vector<int32_t> todo;  // registers that still need computing
  for (size_t i = 0; i < outs.size(); i++) {
    auto it = cache.find(outs[i]);
    if (it != cache.end()) result[i] = it->second;
    else todo.push_back(outs[i]);
  }
  if (todo.empty()) return result;
  std::sort(todo.begin(), todo.end());
  todo.erase(std::unique(todo.begin(), todo.end()), todo.end());
  if (array_mode && !pending.empty()) {
    // deferred products the request depends on are launched now — with the request's own elementwise chain in the
    // epilogue where it is one of the known tails — and whatever that settled drops out of the request
    resolve_pending(todo);
    std::vector<int32_t> left;
    for (int32_t r : todo)
      if (!cache.count(r)) left.push_back(r);
    todo.swap(left);
    if (todo.empty()) {
      for (size_t i = 0; i < outs.size(); i++)
        if (!result[i]) result[i] = cache[outs[i]];
      return result;
    }
  }
  std::vector<ArrayRef> cols;
  for (size_t first = 0; first < todo.size(); first += kMaxOutputsPerKernel) {
    const std::vector<int32_t> part(todo.begin() + first, todo.begin() + std::min(todo.size(), first + kMaxOutputsPerKernel));
    std::vector<ArrayRef> c = (array_mode || use_jit()) ? run_jit(part) : run_interpreter(part);
    cols.insert(cols.end(), c.begin(), c.end());
  }
  for (size_t j = 0; j < todo.size(); j++) cache[todo[j]] = cols[j];
  for (size_t i = 0; i < outs.size(); i++)
    if (!result[i]) result[i] = cache[outs[i]];
  return result;
}

std::vector<ArrayRef> SymProgram::run_interpreter(const std::vector<int32_t>& todo) {
  const int32_t n = int32_t(instrs.size());
  // 1. dead-code elimination: what do the requested outputs depend on?
  std::vector<char> needed(n, 0);
  std::vector<int32_t> stack(todo.begin(), todo.end());
  auto operands = [](const SymInstr& I, int32_t (&ops)[4]) {
    int k = 0;
    if (I.op == S_LOAD || I.op == S_LOADU || I.op == S_CONST) return 0;
    if (I.a >= 0) ops[k++] = I.a;
    if (I.b >= 0) ops[k++] = I.b;
    if (I.c >= 0) ops[k++] = I.c;
    if (I.d >= 0) ops[k++] = I.d;
    return k;
  };
  while (!stack.empty()) {
    int32_t r = stack.back();
    stack.pop_back();
    if (needed[r]) continue;
    needed[r] = 1;
    int32_t ops[4];
    int k = operands(instrs[r], ops);
    for (int j = 0; j < k; j++)
      if (!needed[ops[j]]) stack.push_back(ops[j]);
  }
  // 2. last use of every needed register (outputs live to the end)
  std::vector<int32_t> last_use(n, -1);
  for (int32_t i = 0; i < n; i++) {
    if (!needed[i]) continue;
    int32_t ops[4];
    int k = operands(instrs[i], ops);
    for (int j = 0; j < k; j++) last_use[ops[j]] = i;
  }
  for (int32_t r : todo) last_use[r] = n;
  // 3. slot assignment by liveness; inputs actually used
  std::vector<int32_t> slot(n, -1), free_slots;
  int32_t next_slot = 0;
  std::vector<DevInstr> dev;
  std::vector<int32_t> used_inputs_map(inputs.size(), -1);
  std::vector<const void*> in_ptrs;
  for (int32_t i = 0; i < n; i++) {
    if (!needed[i]) continue;
    const SymInstr& I = instrs[i];
    DevInstr D{};
    D.op = I.op;
    D.imm = I.imm;
    int32_t ops[4];
    int k = operands(I, ops);
    uint16_t* fields[4] = {&D.a, &D.b, &D.c, &D.d};
    // operands read before the destination is written, so a slot freed here may be reused as dst
    int field = 0;
    if (I.op == S_LOAD || I.op == S_LOADU) {
      int32_t idx = I.a;
      if (used_inputs_map[idx] < 0) {
        used_inputs_map[idx] = int32_t(in_ptrs.size());
        in_ptrs.push_back(inputs[idx]->ptr());
      }
      D.a = uint16_t(used_inputs_map[idx]);
    } else {
      if (I.a >= 0) { *fields[0] = uint16_t(slot[I.a]); }
      if (I.b >= 0) { *fields[1] = uint16_t(slot[I.b]); }
      if (I.c >= 0) { *fields[2] = uint16_t(slot[I.c]); }
      if (I.d >= 0) { *fields[3] = uint16_t(slot[I.d]); }
      (void)field;
      for (int j = 0; j < k; j++) {
        int32_t r = ops[j];
        if (last_use[r] == i && slot[r] >= 0) {
          bool again = false;  // the same register may appear twice (x*x)
          for (int jj = 0; jj < j; jj++) again = again || ops[jj] == r;
          if (!again) free_slots.push_back(slot[r]);
        }
      }
    }
    int32_t s;
    if (!free_slots.empty()) { s = free_slots.back(); free_slots.pop_back(); }
    else s = next_slot++;
    slot[i] = s;
    D.dst = uint16_t(s);
    dev.push_back(D);
    if (last_use[i] < 0) free_slots.push_back(s);  // needed only as... (cannot happen after DCE)
  }
  if (next_slot > 65535 || in_ptrs.size() > 65535) fail(GADCU_ERR_UNSUPPORTED, "batched tape too large");
  // 4. stores
  std::vector<ArrayRef> out_cols;
  std::vector<void*> out_ptrs;
  for (int32_t r : todo) {
    ArrayRef col = new_array(ctx, dtype, Dim4(lanes));
    DevInstr D{};
    D.op = S_STORE;
    D.a = uint16_t(slot[r]);
    D.b = uint16_t(out_ptrs.size());
    dev.push_back(D);
    out_ptrs.push_back(col->ptr());
    out_cols.push_back(col);
  }
  last_instructions = dev.size();
  last_slots = uint64_t(next_slot);

  // 5. upload program + pointer tables, launch
  const size_t prog_bytes = dev.size() * sizeof(DevInstr);
  const size_t in_bytes = (in_ptrs.size() + 1) * sizeof(void*), out_bytes = (out_ptrs.size() + 1) * sizeof(void*);
  auto align = [](size_t x) { return (x + 255) / 256 * 256; };
  ArrayRef blob = new_array(ctx, GADCU_F64, Dim4((align(prog_bytes) + align(in_bytes) + align(out_bytes)) / 8 + 1));
  std::vector<char> host(align(prog_bytes) + align(in_bytes) + align(out_bytes), 0);
  std::memcpy(host.data(), dev.data(), prog_bytes);
  std::memcpy(host.data() + align(prog_bytes), in_ptrs.data(), in_ptrs.size() * sizeof(void*));
  std::memcpy(host.data() + align(prog_bytes) + align(in_bytes), out_ptrs.data(), out_ptrs.size() * sizeof(void*));
  ctx->segment_flush();
  GADCU_CUDA(cudaMemcpyAsync(blob->ptr(), host.data(), host.size(), cudaMemcpyHostToDevice, ctx->stream));
  GADCU_CUDA(cudaStreamSynchronize(ctx->stream));  // `host` is pageable
  char* base = static_cast<char*>(blob->ptr());
  const DevInstr* dprog = reinterpret_cast<const DevInstr*>(base);
  const void* din = base + align(prog_bytes);
  void* dout = base + align(prog_bytes) + align(in_bytes);
  if (dtype == GADCU_F32) launch_interpreter<float>(ctx, next_slot, dprog, int(dev.size()), din, dout, lanes);
  else launch_interpreter<double>(ctx, next_slot, dprog, int(dev.size()), din, dout, lanes);

  return out_cols;
}

// ---- NVRTC path: the program as one straight-line kernel, values in registers ----------------------
namespace {

const char* op_expr(uint16_t op) {  // %a %b %c %d = operand names, %k = immediate; mirrors the interpreter above / ew.cu
  switch (op) {
    case S_CONST: return "%k";
    case S_ADD: return "%a + %b";
    case S_SUB: return "%a - %b";
    case S_MUL: return "%a * %b";
    case S_DIV: return "%a / %b";
    case S_NEG: return "-%a";
    case S_NEG0: return "T(0) - %a";
    case S_ADDC: return "%a + %k";
    case S_MULC: return "%a * %k";
    case S_POWI: return "g_powi(%a, int(%k))";
    case S_POWF: return "G_POW(%a, %k)";
    case S_POW: return "G_POW(%a, %b)";
    case S_EXP: return "G_EXP(%a)";
    case S_LOG: return "G_LOG(%a)";
    case S_LOG1P: return "G_LOG(T(1) + %a)";
    case S_LOG1PF: return "G_LOG1P(%a)";
    case S_SIN: return "G_SIN(%a)";
    case S_COS: return "G_COS(%a)";
    case S_TANH: return "G_TANH(%a)";
    case S_SIGMOID: return "T(1) / (T(1) + G_EXP(-%a))";
    case S_RECIP: return "T(1) / %a";
    case S_SQRT: return "G_SQRT(%a)";
    case S_SELECT: return "%a >= %b ? %c : %d";
    case S_SELECT_R0: return "%a >= %b ? %c : T(0)";
    case S_SELECT_R1: return "%a >= %b ? T(0) : %c";
    case S_MIN: return "%a <= %b ? %a : %b";
    case S_MAX: return "%a >= %b ? %a : %b";
    case S_FMIN: return "G_FMIN(%a, %b)";
    case S_FMAX: return "G_FMAX(%a, %b)";
    default: return nullptr;
  }
}

struct JitPlan {
  int V = 1, U = 1, block = 256;
  bool idx32 = true;
  std::vector<int32_t> order;                  // needed non-load instructions, ascending
  std::vector<std::pair<int32_t, int>> dense;  // (register, param slot): dense [lanes] inputs incl. cached values
  std::vector<std::pair<int32_t, int>> uniform, col, row;
  std::vector<int32_t> outs;
  std::vector<double> imms;
  std::vector<const void*> in_ptrs;
  std::vector<int32_t> in_inputs;              // per parameter slot: index into the program's inputs (-1: a cached column)
  std::vector<uint64_t> tile_d0;               // one per col/row input, in (col..., row...) order
  std::unordered_map<int32_t, int32_t> local;  // program register -> dense numbering of the needed ones, so that the
                                               // same structure at different positions of a program is the same kernel
  // long programs (a whole scalar tape, forward + reverse): one lane per thread, and the instructions — loads and
  // stores included — re-ordered to keep few values alive at a time, so the kernel fits the register file with
  // enough resident warps to stream from HBM. Any order that respects the data flow gives the same bits.
  bool flat = false;
  std::vector<int32_t> sched;
  // flat + ring > 0: the lane inputs are not loaded by the computing threads but streamed into a shared-memory ring of
  // `ring` column tiles by a producer warp (cp.async.bulk + mbarriers), which runs ahead across tiles: the bytes in
  // flight no longer cost registers (the register-prefetching kernel was pinned at 8 resident warps by its ~200
  // registers and at 8 loads in flight per thread by the same budget)
  int ring = 0, stage_blocks = 0;  // ring depth; resident blocks per SM the kernel is compiled for (its register budget)
  size_t smem_bytes = 0;
  // the one output is not stored but reduced (SymProgram::ReduceSpec): 0 no, 1 sum, 2 max; td[] carries reduce, chunk, splits
  int red = 0;
  size_t red_td = 0;  // index of `reduce` in td[] (chunk and splits follow)
};

thread_local const JitPlan* g_plan = nullptr;  // the plan being rendered by this thread
std::string reg_name(int32_t r) { return "r" + std::to_string(g_plan->local.at(r)); }

std::string gen_source(const SymProgram& P, const JitPlan& J, const std::unordered_map<int32_t, int>& imm_slot) {
  const bool f32 = P.dtype == GADCU_F32;
  std::ostringstream o;
  o << "typedef " << (f32 ? "float" : "double") << " T;\n";
  if (J.V == 4) o << "typedef float4 TV;\n";
  else if (J.V == 2) o << "typedef double2 TV;\n";
  else o << "typedef T TV;\n";
  o << "typedef unsigned long long u64;\n";
  if (f32)
    o << "#define G_EXP expf\n#define G_LOG logf\n#define G_LOG1P log1pf\n#define G_SIN sinf\n#define G_COS cosf\n#define G_TANH tanhf\n"
         "#define G_SQRT sqrtf\n#define G_POW powf\n#define G_FMIN fminf\n#define G_FMAX fmaxf\n";
  else
    o << "#define G_EXP exp\n#define G_LOG log\n#define G_LOG1P log1p\n#define G_SIN sin\n#define G_COS cos\n#define G_TANH tanh\n"
         "#define G_SQRT sqrt\n#define G_POW pow\n#define G_FMIN fmin\n#define G_FMAX fmax\n";
  o << "__device__ __forceinline__ T g_powi(T a, int b) { const bool recip = b < 0; T r = T(1); while (true) { if (b & 1) r *= a; b /= 2; "
       "if (b == 0) break; a *= a; } return recip ? T(1) / r : r; }\n";
  const size_t NI = J.in_ptrs.size(), NO = J.outs.size(), NT = J.tile_d0.size(), NM = J.imms.size();
  o << "struct P { const T* in[" << std::max<size_t>(NI, 1) << "]; T* out[" << std::max<size_t>(NO, 1) << "]; u64 n; u64 td["
    << std::max<size_t>(NT, 1) << "]; double imm[" << std::max<size_t>(NM, 1) << "]; };\n";
  if (J.flat) {
    std::unordered_map<int32_t, int> in_slot, out_slot, tile_slot;
    for (auto& d : J.dense) in_slot[d.first] = d.second;
    for (auto& d : J.uniform) in_slot[d.first] = d.second;
    for (size_t k = 0; k < J.col.size(); k++) { in_slot[J.col[k].first] = J.col[k].second; tile_slot[J.col[k].first] = int(k); }
    for (size_t k = 0; k < J.row.size(); k++) { in_slot[J.row[k].first] = J.row[k].second; tile_slot[J.row[k].first] = int(J.col.size() + k); }
    for (size_t k = 0; k < NO; k++) out_slot[J.outs[k]] = int(k);
    static const int min_blocks = [] { const char* e = getenv("GADCU_LANE_MINBLOCKS"); return e ? atoi(e) : 6; }();  // 6 x 128 threads: <= 80 registers, 24 warps per SM (measured best on C2)
    // W lanes per thread: with two, every load / store is 128 bits wide (f64) and half as many threads carry the
    // same bytes in flight
    const int W = J.V;
    const char* sfx[2] = {W == 1 ? "" : "_0", "_1"};
    const char* comp[2] = {".x", ".y"};
    o << "typedef " << (f32 ? "float2" : "double2") << " TV2;\n";
    std::unordered_map<int32_t, int> ring_index;  // lane load -> its position among the tile's loads, in schedule order
    if (J.ring) {
      const int stage_blocks = J.stage_blocks;
      std::vector<int> table;
      for (int32_t r : J.sched)
        if (P.cache.count(r) || P.instrs[r].op == S_LOAD) {
          ring_index[r] = int(table.size());
          table.push_back(in_slot.at(r));
        }
      const int D = J.ring, NL = int(table.size());
      int logd = 0;
      while ((1 << logd) < D) logd++;
      o << "#define TILE_BYTES " << J.block * W * (f32 ? 4 : 8) << "u\n";
      o << "__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {\n  unsigned done;\n  do {\n"
           "    asm volatile(\"{\\n.reg .pred p;\\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\\nselp.u32 %0, 1, 0, p;\\n}\" : \"=r\"(done) : \"r\"(bar), \"r\"(parity) : \"memory\");\n"
           "  } while (!done);\n}\n";
      o << "__constant__ unsigned short ring_slot_of[" << NL << "] = {";
      for (int j = 0; j < NL; j++) o << table[size_t(j)] << (j + 1 < NL ? "," : "");
      o << "};\n";
      o << "extern \"C\" __global__ void __launch_bounds__(" << J.block + 32 << ", " << stage_blocks << ") lane_program(const P p) {\n";
      o << "  extern __shared__ __align__(128) unsigned char smem[];\n";
      o << "  const unsigned sbase = (unsigned)__cvta_generic_to_shared(smem), bfull = sbase + " << D << "u * TILE_BYTES, bempty = bfull + " << 8 * D << "u;\n";
      o << "  if (threadIdx.x == 0) {\n    for (unsigned s = 0; s < " << D << "u; s++) {\n"
           "      asm volatile(\"mbarrier.init.shared::cta.b64 [%0], %1;\" ::\"r\"(bfull + 8 * s), \"r\"(1u));\n"
           "      asm volatile(\"mbarrier.init.shared::cta.b64 [%0], %1;\" ::\"r\"(bempty + 8 * s), \"r\"(" << J.block / 32 << "u));\n    }\n"
           "    asm volatile(\"fence.mbarrier_init.release.cluster;\" ::: \"memory\");\n"
           "    asm volatile(\"fence.proxy.async.shared::cta;\" ::: \"memory\");\n  }\n  __syncthreads();\n";
      o << "  const u64 units = p.n / " << W << ";\n  const u64 tiles = (units + " << J.block - 1 << "ull) / " << J.block << "ull;\n";
      o << "  if (threadIdx.x >= " << J.block << ") {\n    if (threadIdx.x == " << J.block << ") {\n      unsigned q = 0;\n"
           "      for (u64 t = blockIdx.x; t < tiles; t += gridDim.x) {\n"
           "        const u64 first = t * " << J.block << "ull;\n        const u64 left = units - first;\n"
           "        const unsigned bytes = (unsigned)(left < " << J.block << "ull ? left : " << J.block << "ull) * " << W * (f32 ? 4 : 8) << "u;\n"
           "#pragma unroll 1\n        for (int j = 0; j < " << NL << "; j++, q++) {\n"
           "          const unsigned s = q & " << D - 1 << "u, ph = (q >> " << logd << ") & 1u;\n"
           "          mbar_wait(bempty + 8 * s, ph ^ 1u);\n"
           "          asm volatile(\"mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\" ::\"r\"(bfull + 8 * s), \"r\"(bytes) : \"memory\");\n"
           "          asm volatile(\"cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\" ::\"r\"(sbase + s * TILE_BYTES), "
           "\"l\"(p.in[ring_slot_of[j]] + first * " << W << "ull), \"r\"(bytes), \"r\"(bfull + 8 * s) : \"memory\");\n"
           "        }\n      }\n    }\n    return;\n  }\n";
      o << "  unsigned q0 = 0;\n  for (u64 t = blockIdx.x; t < tiles; t += gridDim.x, q0 += " << NL << "u) {\n"
           "    const u64 i = t * " << J.block << "ull + threadIdx.x;\n    const bool live = i < units;\n";
    } else {
    o << "extern \"C\" __global__ void __launch_bounds__(" << J.block;
    if (min_blocks > 0) o << ", " << (W == 1 ? min_blocks : std::max(1, min_blocks / 3));  // two lanes: 2 x 128 threads, ~200 registers, no spills
    o << ") lane_program(const P p) {\n";
    o << "  const u64 tid = (u64)blockIdx.x * " << J.block << "ull + threadIdx.x;\n  const u64 stride = (u64)gridDim.x * " << J.block << "ull;\n";
    o << "  for (u64 i = tid; i < p.n / " << W << "; i += stride) {\n";
    }
    for (int32_t r : J.sched) {
      const SymInstr& I = P.instrs[r];
      const bool cached = P.cache.count(r) != 0;
      if ((cached || I.op == S_LOAD) && J.ring) {
        int logd = 0;
        while ((1 << logd) < J.ring) logd++;
        const int j = ring_index.at(r);
        o << "    TV2 v" << reg_name(r) << ";\n    {\n      const unsigned q = q0 + " << j << "u, s = q & " << J.ring - 1 << "u;\n"
             "      mbar_wait(bfull + 8 * s, (q >> " << logd << ") & 1u);\n"
             "      v" << reg_name(r) << " = *reinterpret_cast<const TV2*>(smem + s * TILE_BYTES + threadIdx.x * sizeof(TV2));\n"
             "      __syncwarp();\n"
             "      if ((threadIdx.x & 31) == 0) asm volatile(\"mbarrier.arrive.shared::cta.b64 _, [%0];\" ::\"r\"(bempty + 8 * s) : \"memory\");\n    }\n";
        for (int l = 0; l < W; l++) o << "    const T " << reg_name(r) << sfx[l] << " = v" << reg_name(r) << comp[l] << ";\n";
      } else if (cached || I.op == S_LOAD) {
        if (W == 1) {
          o << "    const T " << reg_name(r) << " = __ldcs(p.in[" << in_slot.at(r) << "] + i);\n";
        } else {
          o << "    const TV2 v" << reg_name(r) << " = __ldcs(reinterpret_cast<const TV2*>(p.in[" << in_slot.at(r) << "]) + i);\n";
          for (int l = 0; l < W; l++) o << "    const T " << reg_name(r) << sfx[l] << " = v" << reg_name(r) << comp[l] << ";\n";
        }
      } else if (I.op == S_LOADU) {
        for (int l = 0; l < W; l++) o << "    const T " << reg_name(r) << sfx[l] << " = p.in[" << in_slot.at(r) << "][0];\n";
      } else if (I.op == S_LOADCOL) {
        o << "    const T " << reg_name(r) << " = p.in[" << in_slot.at(r) << "][" << (J.idx32 ? "(unsigned)i / (unsigned)p.td[" : "i / p.td[") << tile_slot.at(r) << "]];\n";
      } else if (I.op == S_LOADROW) {
        o << "    const T " << reg_name(r) << " = p.in[" << in_slot.at(r) << "][" << (J.idx32 ? "(unsigned)i % (unsigned)p.td[" : "i % p.td[") << tile_slot.at(r) << "]];\n";
      } else {
        const char* e = op_expr(I.op);
        for (int l = 0; l < W; l++) {
          std::string expr;
          for (const char* c = e; *c; c++) {
            if (*c == '%') {
              c++;
              switch (*c) {
                case 'a': expr += reg_name(I.a) + sfx[l]; break;
                case 'b': expr += reg_name(I.b) + sfx[l]; break;
                case 'c': expr += reg_name(I.c) + sfx[l]; break;
                case 'd': expr += reg_name(I.d) + sfx[l]; break;
                case 'k': expr += "T(p.imm[" + std::to_string(imm_slot.at(r)) + "])"; break;
              }
            } else {
              expr += *c;
            }
          }
          o << "    const T " << reg_name(r) << sfx[l] << " = " << expr << ";\n";
        }
      }
      auto os = out_slot.find(r);
      if (os != out_slot.end()) {
        if (W == 1) {
          o << "    p.out[" << os->second << "][i] = " << reg_name(r) << ";\n";
        } else {
          o << "    " << (J.ring ? "if (live) " : "") << "{ TV2 w; w.x = " << reg_name(r) << sfx[0] << "; w.y = " << reg_name(r) << sfx[1]
            << "; reinterpret_cast<TV2*>(p.out[" << os->second << "])[i] = w; }\n";
        }
      }
    }
    o << "  }\n}\n";
    return o.str();
  }
  // ---- body: one lane
  o << "__device__ __forceinline__ void body(const P& p, u64 i";
  for (auto& d : J.dense) o << ", const T " << reg_name(d.first);
  for (auto& d : J.uniform) o << ", const T " << reg_name(d.first);
  for (auto& d : J.col) o << ", const T " << reg_name(d.first);
  for (size_t k = 0; k < NO; k++) o << ", T& o" << k;
  o << ") {\n";
  for (size_t k = 0; k < J.row.size(); k++) {
    const size_t t = J.col.size() + k;
    o << "  const T " << reg_name(J.row[k].first) << " = p.in[" << J.row[k].second << "]["
      << (J.idx32 ? "(unsigned)i % (unsigned)p.td[" : "i % p.td[") << t << "]];\n";
  }
  for (int32_t r : J.order) {
    const SymInstr& I = P.instrs[r];
    const char* e = op_expr(I.op);
    std::string expr;
    for (const char* c = e; *c; c++) {
      if (*c == '%') {
        c++;
        switch (*c) {
          case 'a': expr += reg_name(I.a); break;
          case 'b': expr += reg_name(I.b); break;
          case 'c': expr += reg_name(I.c); break;
          case 'd': expr += reg_name(I.d); break;
          case 'k': expr += "T(p.imm[" + std::to_string(imm_slot.at(r)) + "])"; break;
        }
      } else {
        expr += *c;
      }
    }
    o << "  const T " << reg_name(r) << " = " << expr << ";\n";
  }
  for (size_t k = 0; k < NO; k++) o << "  o" << k << " = " << reg_name(J.outs[k]) << ";\n";
  o << "}\n";
  auto col_index = [&](size_t k, const std::string& i) {
    return J.idx32 ? "(unsigned)(" + i + ") / (unsigned)p.td[" + std::to_string(k) + "]" : "(" + i + ") / p.td[" + std::to_string(k) + "]";
  };
  const char* lanes4[] = {".x", ".y", ".z", ".w"};
  if (J.red) {
    // ---- reducing skeleton: reduce.cu's reduce_contig_kernel with the elementwise chain in place of the load. One block
    // per (segment, chunk); 256 threads walk the chunk with a stride of 256 vectors keeping the V vector lanes apart,
    // lanes combined in order, butterfly over the warp, the 8 warps in order. Same cut, same order => same bits.
    const std::string td = "p.td[" + std::to_string(J.red_td) + "]", tdc = "p.td[" + std::to_string(J.red_td + 1) + "]",
                      tds = "p.td[" + std::to_string(J.red_td + 2) + "]";
    o << (J.red == 1 ? "#define COMB(a, b) ((a) + (b))\n#define IDENT T(0)\n" : "#define COMB(a, b) ((b) > (a) ? (b) : (a))\n") << (J.red == 1 ? "" : f32 ? "#define IDENT (-__int_as_float(0x7f800000))\n" : "#define IDENT (-__longlong_as_double(0x7ff0000000000000ll))\n");
    if (J.ring) {
      // ring-fed variant of the same walk: a producer warp streams the chunk's tiles of 256 vectors per dense input through a
      // shared-memory ring (cp.async.bulk + full / empty mbarriers, as in the ring-fed region skeleton below) and runs
      // ahead; computing thread t takes vector t of every tile — the very vectors, in the very order, of the register-fed
      // walk (vlo + t, + 256, ...), so the accumulators see the same additions. Measured on the C3 forward (64 Mi f32, one
      // call, same box): register-fed 0.101 ms, ring of 4 / 8 / 16 tiles 0.121 / 0.118 / 0.116 ms — with 256-thread blocks
      // the tiles are 4 KB and every block pays its own barrier set-up for ~56 tiles; OFF unless GADCU_REDUCE_RING says so.
      const int D = J.ring, ND = int(J.dense.size());
      int logd = 0;
      while ((1 << logd) < D) logd++;
      o << "#define TILE_BYTES 4096u\n";
      o << "__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {\n  unsigned done;\n  do {\n"
           "    asm volatile(\"{\\n.reg .pred p;\\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\\nselp.u32 %0, 1, 0, p;\\n}\" : \"=r\"(done) : \"r\"(bar), \"r\"(parity) : \"memory\");\n"
           "  } while (!done);\n}\n";
      o << "extern \"C\" __global__ void __launch_bounds__(288" << (J.stage_blocks ? ", " + std::to_string(J.stage_blocks) : std::string()) << ") lane_program(const P p) {\n";
      o << "  extern __shared__ __align__(128) unsigned char smem[];\n  __shared__ T warp_part[8];\n";
      o << "  const unsigned sbase = (unsigned)__cvta_generic_to_shared(smem), bfull = sbase + " << D << "u * TILE_BYTES, bempty = bfull + " << 8 * D << "u;\n";
      o << "  if (threadIdx.x == 0) {\n    for (unsigned s = 0; s < " << D << "u; s++) {\n"
           "      asm volatile(\"mbarrier.init.shared::cta.b64 [%0], %1;\" ::\"r\"(bfull + 8 * s), \"r\"(1u));\n"
           "      asm volatile(\"mbarrier.init.shared::cta.b64 [%0], %1;\" ::\"r\"(bempty + 8 * s), \"r\"(8u));\n    }\n"
           "    asm volatile(\"fence.mbarrier_init.release.cluster;\" ::: \"memory\");\n"
           "    asm volatile(\"fence.proxy.async.shared::cta;\" ::: \"memory\");\n  }\n  __syncthreads();\n";
      o << "  const u64 reduce = " << td << ", chunk = " << tdc << ", splits = " << tds << ";\n";
      o << "  const u64 seg = blockIdx.x / splits, split = blockIdx.x % splits;\n";
      o << "  const u64 lo = split * chunk;\n  u64 hi = lo + chunk;\n  if (hi > reduce) hi = reduce;\n  if (hi < lo) hi = lo;\n";
      o << "  const u64 vbase = seg * reduce / " << J.V << ", vlo = lo / " << J.V << ", vhi = hi / " << J.V << ";\n";
      o << "  const u64 tiles = (vhi - vlo + 255ull) / 256ull;\n";
      o << "  if (threadIdx.x >= 256) {\n    if (threadIdx.x == 256) {\n      unsigned q = 0;\n      for (u64 t = 0; t < tiles; t++) {\n"
           "        const u64 first = vlo + t * 256ull;\n        const u64 left = vhi - first;\n"
           "        const unsigned bytes = (unsigned)(left < 256ull ? left : 256ull) * 16u;\n";
      for (int j = 0; j < ND; j++) {
        o << "        {\n          const unsigned s = q & " << D - 1 << "u, ph = (q >> " << logd << ") & 1u;\n"
             "          mbar_wait(bempty + 8 * s, ph ^ 1u);\n"
             "          asm volatile(\"mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\" ::\"r\"(bfull + 8 * s), \"r\"(bytes) : \"memory\");\n"
             "          asm volatile(\"cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\" ::\"r\"(sbase + s * TILE_BYTES), "
             "\"l\"(reinterpret_cast<const TV*>(p.in[" << J.dense[size_t(j)].second << "]) + vbase + first), \"r\"(bytes), \"r\"(bfull + 8 * s) : \"memory\");\n"
             "          q++;\n        }\n";
      }
      o << "      }\n    }\n    return;\n  }\n";
      for (auto& d : J.uniform) o << "  const T " << reg_name(d.first) << " = p.in[" << d.second << "][0];\n";
      for (int l = 0; l < J.V; l++) o << "  T acc" << l << " = IDENT;\n";
      o << "  unsigned q = 0;\n  for (u64 t = 0; t < tiles; t++) {\n    const u64 vi = vlo + t * 256ull + threadIdx.x;\n";
      for (int j = 0; j < ND; j++) {
        o << "    TV v" << J.local.at(J.dense[size_t(j)].first) << ";\n    {\n      const unsigned s = q & " << D - 1 << "u;\n"
             "      mbar_wait(bfull + 8 * s, (q >> " << logd << ") & 1u);\n"
             "      v" << J.local.at(J.dense[size_t(j)].first) << " = *reinterpret_cast<const TV*>(smem + s * TILE_BYTES + threadIdx.x * 16u);\n"
             "      __syncwarp();\n"
             "      if ((threadIdx.x & 31) == 0) asm volatile(\"mbarrier.arrive.shared::cta.b64 _, [%0];\" ::\"r\"(bempty + 8 * s) : \"memory\");\n"
             "      q++;\n    }\n";
      }
      o << "    if (vi < vhi) {\n      const u64 gi = vbase + vi;\n";
      for (size_t k = 0; k < J.col.size(); k++)
        o << "      const T " << reg_name(J.col[k].first) << " = p.in[" << J.col[k].second << "][" << col_index(k, "gi * " + std::to_string(J.V)) << "];\n";
      for (int l = 0; l < J.V; l++) {
        o << "      T y" << l << ";\n      body(p, gi * " << J.V << " + " << l;
        for (auto& d : J.dense) o << ", v" << J.local.at(d.first) << lanes4[l];
        for (auto& d : J.uniform) o << ", " << reg_name(d.first);
        for (auto& d : J.col) o << ", " << reg_name(d.first);
        o << ", y" << l << ");\n";
      }
      for (int l = 0; l < J.V; l++) o << "      acc" << l << " = COMB(acc" << l << ", y" << l << ");\n";
      o << "    }\n  }\n  T v = acc0;\n";
      for (int l = 1; l < J.V; l++) o << "  v = COMB(v, acc" << l << ");\n";
      o << "#pragma unroll\n  for (int off = 16; off > 0; off >>= 1) { const T w = __shfl_xor_sync(0xffffffffu, v, off); v = COMB(v, w); }\n";
      o << "  if ((threadIdx.x & 31) == 0) warp_part[threadIdx.x >> 5] = v;\n  asm volatile(\"bar.sync 1, 256;\" ::: \"memory\");\n";
      o << "  if (threadIdx.x == 0) {\n    T r = warp_part[0];\n#pragma unroll\n    for (int w = 1; w < 8; w++) r = COMB(r, warp_part[w]);\n"
           "    p.out[0][seg * splits + split] = r;\n  }\n}\n";
      return o.str();
    }
    o << "extern \"C\" __global__ void __launch_bounds__(256" << (J.stage_blocks ? ", " + std::to_string(J.stage_blocks) : std::string()) << ") lane_program(const P p) {\n";
    o << "  __shared__ T warp_part[8];\n";
    o << "  const u64 reduce = " << td << ", chunk = " << tdc << ", splits = " << tds << ";\n";
    o << "  const u64 seg = blockIdx.x / splits, split = blockIdx.x % splits;\n";
    o << "  const u64 lo = split * chunk;\n  u64 hi = lo + chunk;\n  if (hi > reduce) hi = reduce;\n  const u64 base = seg * reduce;\n";
    for (auto& d : J.uniform) o << "  const T " << reg_name(d.first) << " = p.in[" << d.second << "][0];\n";
    o << "  T v = IDENT;\n  if (lo < hi) {\n";
    if (J.V > 1) {
      o << "    const u64 vbase = base / " << J.V << ", vlo = lo / " << J.V << ", vhi = hi / " << J.V << ";\n";
      for (int l = 0; l < J.V; l++) o << "    T acc" << l << " = IDENT;\n";
      // RU steps of a thread's walk at a time: all their loads first, then the chain and the accumulation step by step in
      // walk order (the order of the additions is what fixes the bits, not when the loads were issued)
      const int RU = J.U;
      o << "    for (u64 walk = vlo + threadIdx.x; walk < vhi; walk += " << 256 * RU << ") {\n";
      for (auto& d : J.dense) o << "      TV v" << J.local.at(d.first) << "[" << RU << "];\n";
      o << "#pragma unroll\n      for (int u = 0; u < " << RU << "; u++) {\n        const u64 vi = walk + u * 256;\n        if (vi < vhi) {\n";
      for (auto& d : J.dense) o << "          v" << J.local.at(d.first) << "[u] = __ldcs(reinterpret_cast<const TV*>(p.in[" << d.second << "]) + vbase + vi);\n";
      o << "        }\n      }\n";
      o << "#pragma unroll\n      for (int u = 0; u < " << RU << "; u++) {\n        const u64 vi = walk + u * 256;\n        if (vi < vhi) {\n          const u64 gi = vbase + vi;\n";
      for (size_t k = 0; k < J.col.size(); k++)
        o << "          const T " << reg_name(J.col[k].first) << " = p.in[" << J.col[k].second << "][" << col_index(k, "gi * " + std::to_string(J.V)) << "];\n";
      for (int l = 0; l < J.V; l++) {
        o << "          T y" << l << ";\n          body(p, gi * " << J.V << " + " << l;
        for (auto& d : J.dense) o << ", v" << J.local.at(d.first) << "[u]" << lanes4[l];
        for (auto& d : J.uniform) o << ", " << reg_name(d.first);
        for (auto& d : J.col) o << ", " << reg_name(d.first);
        o << ", y" << l << ");\n";
      }
      for (int l = 0; l < J.V; l++) o << "          acc" << l << " = COMB(acc" << l << ", y" << l << ");\n";
      o << "        }\n      }\n    }\n    v = acc0;\n";
      for (int l = 1; l < J.V; l++) o << "    v = COMB(v, acc" << l << ");\n";
    } else {
      o << "    for (u64 i = lo + threadIdx.x; i < hi; i += 256) {\n      const u64 gi = base + i;\n";
      for (auto& d : J.dense) o << "      const T s" << J.local.at(d.first) << " = p.in[" << d.second << "][gi];\n";
      for (size_t k = 0; k < J.col.size(); k++)
        o << "      const T " << reg_name(J.col[k].first) << " = p.in[" << J.col[k].second << "][" << col_index(k, "gi") << "];\n";
      o << "      T y;\n      body(p, gi";
      for (auto& d : J.dense) o << ", s" << J.local.at(d.first);
      for (auto& d : J.uniform) o << ", " << reg_name(d.first);
      for (auto& d : J.col) o << ", " << reg_name(d.first);
      o << ", y);\n      v = COMB(v, y);\n    }\n";
    }
    o << "  }\n";
    o << "#pragma unroll\n  for (int off = 16; off > 0; off >>= 1) { const T w = __shfl_xor_sync(0xffffffffu, v, off); v = COMB(v, w); }\n";
    o << "  if ((threadIdx.x & 31) == 0) warp_part[threadIdx.x >> 5] = v;\n  __syncthreads();\n";
    o << "  if (threadIdx.x == 0) {\n    T r = warp_part[0];\n#pragma unroll\n    for (int w = 1; w < 8; w++) r = COMB(r, warp_part[w]);\n"
         "    p.out[0][seg * splits + split] = r;\n  }\n}\n";
    return o.str();
  }
  if (J.ring) {
    // ---- ring-fed skeleton (array mode, large regions): a producer warp streams the dense inputs, tile by tile
    // (BLOCK vectors of 16 bytes per input = 8 KB bulk copies), through a shared-memory ring with cp.async.bulk +
    // full / empty mbarriers and runs ahead across tiles; the BLOCK computing threads take one vector per input and tile
    // from the ring. Bytes in flight then cost no registers and do not depend on what the computing warps are doing
    // (the register-fed shape alternates between waiting for its loads and ~1000 instructions of exp / log / div:
    // measured on the C3 sweep, 64 Mi f32: 0.191 ms register-fed, 0.177 ms ring-fed, same bits; smaller tiles or
    // deeper rings are slower — experiments/c3_sweep_ring.cu).
    const int D = J.ring, B = J.block, ND = int(J.dense.size());
    int logd = 0;
    while ((1 << logd) < D) logd++;
    o << "#define TILE_BYTES " << B * 16 << "u\n";
    o << "__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {\n  unsigned done;\n  do {\n"
         "    asm volatile(\"{\\n.reg .pred p;\\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\\nselp.u32 %0, 1, 0, p;\\n}\" : \"=r\"(done) : \"r\"(bar), \"r\"(parity) : \"memory\");\n"
         "  } while (!done);\n}\n";
    o << "extern \"C\" __global__ void __launch_bounds__(" << B + 32 << ", " << J.stage_blocks << ") lane_program(const P p) {\n";
    o << "  extern __shared__ __align__(128) unsigned char smem[];\n";
    o << "  const unsigned sbase = (unsigned)__cvta_generic_to_shared(smem), bfull = sbase + " << D << "u * TILE_BYTES, bempty = bfull + " << 8 * D << "u;\n";
    o << "  if (threadIdx.x == 0) {\n    for (unsigned s = 0; s < " << D << "u; s++) {\n"
         "      asm volatile(\"mbarrier.init.shared::cta.b64 [%0], %1;\" ::\"r\"(bfull + 8 * s), \"r\"(1u));\n"
         "      asm volatile(\"mbarrier.init.shared::cta.b64 [%0], %1;\" ::\"r\"(bempty + 8 * s), \"r\"(" << B / 32 << "u));\n    }\n"
         "    asm volatile(\"fence.mbarrier_init.release.cluster;\" ::: \"memory\");\n"
         "    asm volatile(\"fence.proxy.async.shared::cta;\" ::: \"memory\");\n  }\n  __syncthreads();\n";
    o << "  const u64 nvec = p.n / " << J.V << ";\n  const u64 tiles = (nvec + " << B - 1 << "ull) / " << B << "ull;\n";
    // producer
    o << "  if (threadIdx.x >= " << B << ") {\n    if (threadIdx.x == " << B << ") {\n      unsigned q = 0;\n"
         "      for (u64 t = blockIdx.x; t < tiles; t += gridDim.x) {\n"
         "        const u64 first = t * " << B << "ull;\n        const u64 left = nvec - first;\n"
         "        const unsigned bytes = (unsigned)(left < " << B << "ull ? left : " << B << "ull) * 16u;\n";
    for (int j = 0; j < ND; j++) {
      o << "        {\n          const unsigned s = q & " << D - 1 << "u, ph = (q >> " << logd << ") & 1u;\n"
           "          mbar_wait(bempty + 8 * s, ph ^ 1u);\n"
           "          asm volatile(\"mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\" ::\"r\"(bfull + 8 * s), \"r\"(bytes) : \"memory\");\n"
           "          asm volatile(\"cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\" ::\"r\"(sbase + s * TILE_BYTES), "
           "\"l\"(reinterpret_cast<const TV*>(p.in[" << J.dense[size_t(j)].second << "]) + first), \"r\"(bytes), \"r\"(bfull + 8 * s) : \"memory\");\n"
           "          q++;\n        }\n";
    }
    o << "      }\n    }\n    return;\n  }\n";
    // consumers
    for (auto& d : J.uniform) o << "  const T " << reg_name(d.first) << " = p.in[" << d.second << "][0];\n";
    o << "  unsigned q = 0;\n  for (u64 t = blockIdx.x; t < tiles; t += gridDim.x) {\n    const u64 vi = t * " << B << "ull + threadIdx.x;\n";
    for (int j = 0; j < ND; j++) {
      o << "    TV v" << J.local.at(J.dense[size_t(j)].first) << ";\n    {\n      const unsigned s = q & " << D - 1 << "u;\n"
           "      mbar_wait(bfull + 8 * s, (q >> " << logd << ") & 1u);\n"
           "      v" << J.local.at(J.dense[size_t(j)].first) << " = *reinterpret_cast<const TV*>(smem + s * TILE_BYTES + threadIdx.x * 16u);\n"
           "      __syncwarp();\n"
           "      if ((threadIdx.x & 31) == 0) asm volatile(\"mbarrier.arrive.shared::cta.b64 _, [%0];\" ::\"r\"(bempty + 8 * s) : \"memory\");\n"
           "      q++;\n    }\n";
    }
    o << "    if (vi < nvec) {\n";
    for (size_t k = 0; k < J.col.size(); k++)
      o << "      const T " << reg_name(J.col[k].first) << " = p.in[" << J.col[k].second << "][" << col_index(k, "vi * " + std::to_string(J.V)) << "];\n";
    for (size_t k = 0; k < NO; k++) o << "      TV y" << k << ";\n";
    for (int l = 0; l < J.V; l++) {
      o << "      body(p, vi * " << J.V << " + " << l;
      for (auto& d : J.dense) o << ", v" << J.local.at(d.first) << lanes4[l];
      for (auto& d : J.uniform) o << ", " << reg_name(d.first);
      for (auto& d : J.col) o << ", " << reg_name(d.first);
      for (size_t k = 0; k < NO; k++) o << ", y" << k << lanes4[l];
      o << ");\n";
    }
    for (size_t k = 0; k < NO; k++) o << "      reinterpret_cast<TV*>(p.out[" << k << "])[vi] = y" << k << ";\n";
    o << "    }\n  }\n";
    // scalar tail (n not a multiple of the vector width): the computing threads of all blocks share it
    o << "  for (u64 i = nvec * " << J.V << " + (u64)blockIdx.x * " << B << "ull + threadIdx.x; i < p.n; i += (u64)gridDim.x * " << B << "ull) {\n";
    for (auto& d : J.dense) o << "    const T s" << J.local.at(d.first) << " = __ldcs(p.in[" << d.second << "] + i);\n";
    for (size_t k = 0; k < J.col.size(); k++)
      o << "    const T " << reg_name(J.col[k].first) << " = p.in[" << J.col[k].second << "][" << col_index(k, "i") << "];\n";
    for (size_t k = 0; k < NO; k++) o << "    T y" << k << ";\n";
    o << "    body(p, i";
    for (auto& d : J.dense) o << ", s" << J.local.at(d.first);
    for (auto& d : J.uniform) o << ", " << reg_name(d.first);
    for (auto& d : J.col) o << ", " << reg_name(d.first);
    for (size_t k = 0; k < NO; k++) o << ", y" << k;
    o << ");\n";
    for (size_t k = 0; k < NO; k++) o << "    p.out[" << k << "][i] = y" << k << ";\n";
    o << "  }\n}\n";
    return o.str();
  }
  // ---- kernel skeleton: all loads of U vectors first, then compute + store
  o << "extern \"C\" __global__ void __launch_bounds__(" << J.block << ") lane_program(const P p) {\n";
  o << "  const u64 tid = (u64)blockIdx.x * " << J.block << "ull + threadIdx.x;\n  const u64 stride = (u64)gridDim.x * " << J.block << "ull;\n";
  for (auto& d : J.uniform) o << "  const T " << reg_name(d.first) << " = p.in[" << d.second << "][0];\n";
  if (J.V > 1) {
    o << "  const u64 nvec = p.n / " << J.V << ";\n";
    o << "  for (u64 base = tid; base < nvec; base += stride * " << J.U << ") {\n";
    for (auto& d : J.dense) o << "    TV v" << J.local.at(d.first) << "[" << J.U << "];\n";
    o << "#pragma unroll\n    for (int u = 0; u < " << J.U << "; u++) {\n      const u64 vi = base + u * stride;\n      if (vi < nvec) {\n";
    for (auto& d : J.dense) o << "        v" << J.local.at(d.first) << "[u] = __ldcs(reinterpret_cast<const TV*>(p.in[" << d.second << "]) + vi);\n";
    o << "      }\n    }\n";
    o << "#pragma unroll\n    for (int u = 0; u < " << J.U << "; u++) {\n      const u64 vi = base + u * stride;\n      if (vi < nvec) {\n";
    for (size_t k = 0; k < J.col.size(); k++)  // dims[0] is a multiple of V: the V lanes of a vector share the column
      o << "        const T " << reg_name(J.col[k].first) << " = p.in[" << J.col[k].second << "][" << col_index(k, "vi * " + std::to_string(J.V)) << "];\n";
    for (size_t k = 0; k < NO; k++) o << "        TV y" << k << ";\n";
    for (int l = 0; l < J.V; l++) {
      o << "        body(p, vi * " << J.V << " + " << l;
      for (auto& d : J.dense) o << ", v" << J.local.at(d.first) << "[u]" << lanes4[l];
      for (auto& d : J.uniform) o << ", " << reg_name(d.first);
      for (auto& d : J.col) o << ", " << reg_name(d.first);
      for (size_t k = 0; k < NO; k++) o << ", y" << k << lanes4[l];
      o << ");\n";
    }
    for (size_t k = 0; k < NO; k++) o << "        reinterpret_cast<TV*>(p.out[" << k << "])[vi] = y" << k << ";\n";
    o << "      }\n    }\n  }\n";
    o << "  for (u64 i = nvec * " << J.V << " + tid; i < p.n; i += stride) {\n";
  } else {
    o << "  for (u64 i = tid; i < p.n; i += stride) {\n";
  }
  // scalar loop (the whole program for V == 1, the tail otherwise)
  for (auto& d : J.dense) o << "    const T s" << J.local.at(d.first) << " = __ldcs(p.in[" << d.second << "] + i);\n";
  for (size_t k = 0; k < J.col.size(); k++)
    o << "    const T " << reg_name(J.col[k].first) << " = p.in[" << J.col[k].second << "][" << col_index(k, "i") << "];\n";
  for (size_t k = 0; k < NO; k++) o << "    T y" << k << ";\n";
  o << "    body(p, i";
  for (auto& d : J.dense) o << ", s" << J.local.at(d.first);
  for (auto& d : J.uniform) o << ", " << reg_name(d.first);
  for (auto& d : J.col) o << ", " << reg_name(d.first);
  for (size_t k = 0; k < NO; k++) o << ", y" << k;
  o << ");\n";
  for (size_t k = 0; k < NO; k++) o << "    p.out[" << k << "][i] = y" << k << ";\n";
  o << "  }\n}\n";
  return o.str();
}

}  // namespace

std::vector<ArrayRef> SymProgram::launch_compiled(const Compiled& c, size_t nout) {
  // outputs + parameter block (layout of `struct P`: every field is 8 bytes wide)
  std::vector<ArrayRef> out_cols;
  for (size_t j = 0; j < nout; j++) out_cols.push_back(new_array(ctx, dtype, Dim4(lanes)));
  const size_t NI = std::max<size_t>(c.in_inputs.size(), 1), NO = std::max<size_t>(nout, 1), NT = std::max<size_t>(c.tile_d0.size(), 1),
               NM = std::max<size_t>(c.imms.size(), 1);
  std::vector<uint64_t> params(NI + NO + 1 + NT + NM, 0);
  for (size_t i = 0; i < c.in_inputs.size(); i++) params[i] = reinterpret_cast<uint64_t>(inputs[size_t(c.in_inputs[i])]->ptr());
  for (size_t i = 0; i < out_cols.size(); i++) params[NI + i] = reinterpret_cast<uint64_t>(out_cols[i]->ptr());
  params[NI + NO] = lanes;
  for (size_t i = 0; i < c.tile_d0.size(); i++) params[NI + NO + 1 + i] = c.tile_d0[i];
  for (size_t i = 0; i < c.imms.size(); i++) std::memcpy(&params[NI + NO + 1 + NT + i], &c.imms[i], 8);
  const uint64_t per_block = uint64_t(c.block) * uint64_t(c.V) * uint64_t(c.U);
  uint64_t grid = (lanes + per_block - 1) / per_block;
  const uint64_t cap = uint64_t(ctx->sm_count) * (c.block == 256 ? 8 : 16);
  if (grid > cap && c.V > 1) grid = cap;
  if (c.smem_bytes) grid = std::min<uint64_t>(grid, staged_grid(ctx, c.smem_bytes, c.stage_blocks));
  if (grid > 0x7fffffffull) grid = 0x7fffffffull;
  if (grid == 0) grid = 1;
  ProfileScope prof_scope(ctx, array_mode ? PROF_EW : PROF_OTHER);
  void* args[] = {params.data()};
  jit::launch(ctx, c.kernel, unsigned(grid), unsigned(c.block + (c.smem_bytes ? 32 : 0)), args, c.smem_bytes);
  last_instructions = c.instructions;
  last_slots = c.values;
  return out_cols;
}

// ---- plans by tape structure ------------------------------------------------------------------------------------
// A gad program re-records its tape every step (net_ext.rs:52), so every step brings a NEW SymProgram with the same
// instruction stream as the last one. Dead-code elimination, list scheduling and building the kernel's structural key
// for a thousand-instruction tape cost far more than recording it; they depend on nothing but the recorded stream, the
// requested outputs and a few shape facts — so finished plans are kept process-wide under exactly that, and a later tape
// with the same structure goes from "sweep recorded" to "kernel launched" with one hash of its stream.
struct PlanKey {
  std::vector<uint64_t> words;
  uint64_t hash = 0;
  bool operator==(const PlanKey& o) const { return hash == o.hash && words == o.words; }
};
struct PlanKeyHash { size_t operator()(const PlanKey& k) const { return size_t(k.hash); } };
std::mutex g_plans_mu;
std::unordered_map<PlanKey, SymProgram::Compiled, PlanKeyHash> g_plans;

PlanKey plan_key(const SymProgram& P, const std::vector<int32_t>& todo) {
  PlanKey k;
  k.words.reserve(P.instrs.size() * 4 + todo.size() + 4);
  bool aligned = true;
  for (const ArrayRef& in : P.inputs) aligned = aligned && (reinterpret_cast<uintptr_t>(in->ptr()) & 15) == 0;
  k.words.push_back(uint64_t(P.dtype) | uint64_t(P.lanes % 4) << 8 | uint64_t(P.lanes < (1ull << 32)) << 16 | uint64_t(aligned) << 17 |
                    uint64_t(P.instrs.size()) << 32);
  for (const SymInstr& I : P.instrs) {
    k.words.push_back(uint64_t(I.op) | uint64_t(uint32_t(I.a)) << 16);
    k.words.push_back(uint64_t(uint32_t(I.b)) | uint64_t(uint32_t(I.c)) << 32);
    uint64_t bits;
    std::memcpy(&bits, &I.imm, 8);
    k.words.push_back(uint64_t(uint32_t(I.d)) ^ (bits * 0x9E3779B97F4A7C15ull));
    k.words.push_back(bits);
  }
  for (int32_t r : todo) k.words.push_back(uint64_t(uint32_t(r)) | 1ull << 40);
  uint64_t h = 1469598103934665603ull;
  for (uint64_t w : k.words) { h ^= w; h *= 1099511628211ull; }
  k.hash = h;
  return k;
}

std::vector<ArrayRef> SymProgram::run_jit(const std::vector<int32_t>& todo, const ReduceSpec* red, bool* red_done) {
  const int32_t n = int32_t(instrs.size());
  const bool reusable = !array_mode && cache.empty() && !red;
  PlanKey pkey;
  if (reusable) {
    auto it = compiled.find(todo);
    if (it != compiled.end()) return launch_compiled(it->second, todo.size());
    static const bool plans_on = [] { const char* e = getenv("GADCU_PLAN_CACHE"); return !e || atoi(e) != 0; }();
    if (plans_on) {
      pkey = plan_key(*this, todo);
      bool hit = false;
      {
        std::lock_guard<std::mutex> lk(g_plans_mu);
        auto g = g_plans.find(pkey);
        if (g != g_plans.end()) { compiled[todo] = g->second; hit = true; }
      }
      if (hit) return launch_compiled(compiled[todo], todo.size());
    }
  }
  auto is_load = [](uint16_t op) { return op == S_LOAD || op == S_LOADU || op == S_LOADCOL || op == S_LOADROW || op == S_GEMM; };
  // 1. what the requested outputs depend on; a register that already has a materialised column is a leaf
  std::vector<char> needed(n, 0);
  std::vector<int32_t> stack(todo.begin(), todo.end());
  while (!stack.empty()) {
    int32_t r = stack.back();
    stack.pop_back();
    if (needed[r]) continue;
    needed[r] = 1;
    const SymInstr& I = instrs[r];
    if (is_load(I.op) || I.op == S_CONST || cache.count(r)) continue;
    for (int32_t x : {I.a, I.b, I.c, I.d})
      if (x >= 0 && !needed[x]) stack.push_back(x);
  }
  JitPlan J;
  J.outs = todo;
  std::unordered_map<int32_t, int> imm_slot;
  for (int32_t r = 0; r < n; r++) {
    if (!needed[r]) continue;
    J.local.emplace(r, int32_t(J.local.size()));
    const SymInstr& I = instrs[r];
    auto cached = cache.find(r);
    if (cached != cache.end()) {
      J.dense.push_back({r, int(J.in_ptrs.size())});
      J.in_ptrs.push_back(cached->second->ptr());
      J.in_inputs.push_back(-1);
    } else if (I.op == S_LOAD) {
      J.dense.push_back({r, int(J.in_ptrs.size())});
      J.in_ptrs.push_back(inputs[I.a]->ptr());
      J.in_inputs.push_back(I.a);
    } else if (I.op == S_LOADU) {
      J.uniform.push_back({r, int(J.in_ptrs.size())});
      J.in_ptrs.push_back(inputs[I.a]->ptr());
      J.in_inputs.push_back(I.a);
    } else if (I.op == S_LOADCOL) {
      J.col.push_back({r, int(J.in_ptrs.size())});
      J.in_ptrs.push_back(inputs[I.a]->ptr());
      J.in_inputs.push_back(I.a);
    } else if (I.op == S_LOADROW) {
      J.row.push_back({r, int(J.in_ptrs.size())});
      J.in_ptrs.push_back(inputs[I.a]->ptr());
      J.in_inputs.push_back(I.a);
    } else {
      J.order.push_back(r);
      if (I.op == S_CONST || I.op == S_ADDC || I.op == S_MULC || I.op == S_POWI || I.op == S_POWF) {
        imm_slot[r] = int(J.imms.size());
        J.imms.push_back(I.imm);
      }
    }
  }
  for (auto& c : J.col) J.tile_d0.push_back(input_d0[instrs[c.first].a]);
  for (auto& r : J.row) J.tile_d0.push_back(input_d0[instrs[r.first].a]);
  // 2. shape of the kernel: 128-bit vectors and several vectors in flight for small programs, one lane per thread
  //    for long ones (register pressure)
  const int vmax = dtype == GADCU_F32 ? 4 : 2;
  J.V = J.order.size() <= 256 ? vmax : 1;
  for (uint64_t d0 : std::vector<uint64_t>(J.tile_d0.begin(), J.tile_d0.begin() + J.col.size()))
    if (d0 % uint64_t(J.V) != 0) J.V = 1;
  const size_t streams = J.dense.size() + J.outs.size();
  J.U = J.V == 1 ? 1 : (streams <= 4 && J.order.size() <= 64 ? 4 : (streams <= 8 && J.order.size() <= 128 ? 2 : 1));
  J.block = J.order.size() > 256 ? 128 : 256;
  J.idx32 = lanes < (1ull << 32);
  if (red) {
    // the reducing shape exists for regions short enough to sit in one `body`, with the vector width reduce.cu would
    // use on the materialised value (anything else: the caller materialises and reduces as before)
    bool aligned = true;
    for (auto& d : J.dense) aligned = aligned && (reinterpret_cast<uintptr_t>(J.in_ptrs[size_t(d.second)]) & 15) == 0;
    if (J.order.size() > 256 || (red->vec && (J.V != vmax || !aligned || red->reduce % uint64_t(vmax) != 0))) return {};
    if (!red->vec) J.V = 1;
    static const int red_unroll = [] { const char* e = getenv("GADCU_REDUCE_UNROLL"); return e ? std::max(1, atoi(e)) : 4; }();
    J.U = J.V == 1 ? 1 : (J.dense.size() <= 4 ? red_unroll : 1);  // loads in flight per thread (see gen_source)
    J.block = 256;
    J.red = red->op == k::RedOp::SUM ? 1 : 2;
    static const int red_ring = [] { const char* e = getenv("GADCU_REDUCE_RING"); return e ? atoi(e) : 0; }();
    static const uint64_t red_ring_min = [] { const char* e = getenv("GADCU_REDUCE_RING_MIN"); return e ? uint64_t(atoll(e)) : (1ull << 22); }();
    if (red_ring > 0 && J.V == vmax && !J.dense.empty() && J.dense.size() <= 4 && lanes >= red_ring_min) {
      J.ring = red_ring >= 2 ? red_ring : 8;  // slots of one 4 KB tile, shared by the inputs in turn
      J.smem_bytes = size_t(J.ring) * 4096 + size_t(J.ring) * 16;
      J.U = 1;
    }
    J.red_td = J.tile_d0.size();
    J.tile_d0.push_back(red->reduce);
    J.tile_d0.push_back(red->chunk);
    J.tile_d0.push_back(red->splits);
  }
  if (!red && array_mode && J.V == vmax && J.order.size() <= 256 && !J.dense.empty() && J.dense.size() <= 4 && lanes >= (1ull << 22)) {
    // large region: feed the dense inputs through a shared-memory ring (see gen_source). 512-thread tiles, ring of 4
    // (8 with more than two inputs: two tiles' worth in flight), two blocks per SM if the registers allow.
    static const int region_ring = [] { const char* e = getenv("GADCU_REGION_RING"); return e ? atoi(e) : 1; }();
    bool aligned = true;
    for (auto& d : J.dense) aligned = aligned && (reinterpret_cast<uintptr_t>(J.in_ptrs[size_t(d.second)]) & 15) == 0;
    if (region_ring && aligned) {
      J.ring = J.dense.size() <= 2 ? 4 : 8;
      J.block = 512;
      J.U = 1;
      J.smem_bytes = size_t(J.ring) * 512 * 16 + size_t(J.ring) * 16;
    }
  }
  if (!red && J.V == 1 && J.order.size() > 256) {
    // greedy list scheduling for few live values: among the ready instructions take the one that ends the most
    // live ranges (net of the one it starts); prefer computing over loading, then tape order
    J.flat = true;
    static const int lane_width = [] { const char* e = getenv("GADCU_LANE_WIDTH"); return e ? atoi(e) : 2; }();  // measured on C2: 0.249 ms with one lane per thread, 0.200 ms with two
    if (lane_width == 2 && lanes % 2 == 0 && J.col.empty() && J.row.empty()) J.V = 2;
    std::vector<int32_t> nodes;  // needed registers, ascending
    for (int32_t r = 0; r < n; r++)
      if (needed[r]) nodes.push_back(r);
    auto operands_of = [&](int32_t r, int32_t (&ops)[4]) {
      const SymInstr& I = instrs[r];
      if (is_load(I.op) || I.op == S_CONST || cache.count(r)) return 0;
      int k = 0;
      for (int32_t x : {I.a, I.b, I.c, I.d})
        if (x >= 0) {
          bool dup = false;
          for (int j = 0; j < k; j++) dup = dup || ops[j] == x;
          if (!dup) ops[k++] = x;
        }
      return k;
    };
    std::unordered_map<int32_t, int> uses, pending;        // remaining consumers; unscheduled operands
    std::unordered_map<int32_t, std::vector<int32_t>> consumers;
    for (int32_t r : nodes) {
      int32_t ops[4];
      const int k = operands_of(r, ops);
      pending[r] = k;
      for (int j = 0; j < k; j++) { uses[ops[j]]++; consumers[ops[j]].push_back(r); }
    }
    std::vector<int32_t> ready;
    for (int32_t r : nodes)
      if (pending[r] == 0) ready.push_back(r);
    J.sched.reserve(nodes.size());
    std::unordered_map<int32_t, int> when;  // position in the schedule
    auto place = [&](int32_t r) {
      when[r] = int(J.sched.size());
      J.sched.push_back(r);
      int32_t ops[4];
      const int k = operands_of(r, ops);
      for (int j = 0; j < k; j++) uses[ops[j]]--;
      for (int32_t cns : consumers[r])
        if (--pending[cns] == 0) ready.push_back(cns);
    };
    {  // lane-invariant values (the seed, constants) first: they are alive throughout anyway, and the reverse
       // instructions that need them become ready as soon as their forward values exist
      std::vector<int32_t> invariant;
      for (size_t c = 0; c < ready.size();) {
        const SymInstr& I = instrs[ready[c]];
        if ((I.op == S_LOADU || I.op == S_CONST) && !cache.count(ready[c])) {
          invariant.push_back(ready[c]);
          ready[c] = ready.back();
          ready.pop_back();
        } else {
          c++;
        }
      }
      std::sort(invariant.begin(), invariant.end());
      for (int32_t r : invariant) place(r);
    }
    while (!ready.empty()) {
      // key, larger is better: (live ranges ended - started, computes rather than loads, continues from the most
      // recently produced value [depth first: finish a chain before opening another], a load that completes a waiting
      // instruction, tape order)
      size_t best = 0;
      std::tuple<int, int, int, int, int32_t> best_key{-1000, 0, 0, 0, 0};
      for (size_t c = 0; c < ready.size(); c++) {
        const int32_t r = ready[c];
        int32_t ops[4];
        const int k = operands_of(r, ops);
        int freed = 0, recent = -1;
        for (int j = 0; j < k; j++) {
          if (uses[ops[j]] == 1) freed++;
          recent = std::max(recent, when[ops[j]]);
        }
        int completes = 0;
        if (k == 0)
          for (int32_t cns : consumers[r]) completes = std::max(completes, pending[cns] == 1 ? 1 : 0);
        const std::tuple<int, int, int, int, int32_t> key{freed - (uses[r] > 0 ? 1 : 0), k == 0 ? 0 : 1, recent, completes, -r};
        if (c == 0 || key > best_key) {
          best = c;
          best_key = key;
        }
      }
      const int32_t r = ready[best];
      ready[best] = ready.back();
      ready.pop_back();
      place(r);
    }
    {  // shared-memory staging of the lane inputs (see JitPlan::ring): two lanes per thread, whole 16-byte units
      static const int stage = [] { const char* e = getenv("GADCU_LANE_STAGE"); return e ? atoi(e) : 8; }();  // C2: 8 tiles x 5 blocks 0.184 ms, 32 x 3 0.187, 32 x 2 0.199, registers only 0.196
      size_t lane_loads = 0;
      bool aligned = true;
      for (int32_t r : J.sched)
        if (cache.count(r) || instrs[r].op == S_LOAD) lane_loads++;
      for (auto& d : J.dense) aligned = aligned && (reinterpret_cast<uintptr_t>(J.in_ptrs[size_t(d.second)]) & 15) == 0;
      if (stage > 0 && J.V == 2 && lane_loads >= 8 && aligned && lanes % 4 == 0) {
        int D = 1;
        while (D * 2 <= stage && size_t(D * 2) <= lane_loads) D *= 2;
        J.ring = D;
        J.smem_bytes = size_t(D) * size_t(J.block) * 2 * (dtype == GADCU_F32 ? 4 : 8) + size_t(D) * 16;
      }
    }
    if (!J.ring) {  // software pipelining of the lane loads: the load that was about to be used is replaced, in place, by the load
       // needed kPrefetch loads later, and the first kPrefetch are issued up front — every warp keeps kPrefetch
       // independent 256-byte requests in flight (one in flight per warp cannot cover HBM latency at any occupancy)
      static const size_t kPrefetch = [] { const char* e = getenv("GADCU_LANE_PREFETCH"); return e ? size_t(atoi(e)) : size_t(8); }();
      std::vector<int32_t> lane_loads;
      for (int32_t r : J.sched)
        if (cache.count(r) || instrs[r].op == S_LOAD || instrs[r].op == S_LOADCOL || instrs[r].op == S_LOADROW) lane_loads.push_back(r);
      if (lane_loads.size() > kPrefetch) {
        std::unordered_map<int32_t, size_t> load_index;
        for (size_t j = 0; j < lane_loads.size(); j++) load_index[lane_loads[j]] = j;
        std::vector<int32_t> out;
        out.reserve(J.sched.size());
        bool primed = false;
        for (int32_t r : J.sched) {
          auto it = load_index.find(r);
          if (it == load_index.end()) {
            out.push_back(r);
            continue;
          }
          if (!primed) {
            for (size_t j = 0; j < kPrefetch; j++) out.push_back(lane_loads[j]);
            primed = true;
          }
          if (it->second + kPrefetch < lane_loads.size()) out.push_back(lane_loads[it->second + kPrefetch]);
        }
        J.sched.swap(out);
      }
    }
    // register names follow the schedule, so equal structures still give equal source text
    J.local.clear();
    for (int32_t r : J.sched) J.local.emplace(r, int32_t(J.local.size()));
  }
  last_instructions = J.order.size() + J.dense.size() + J.uniform.size() + J.col.size() + J.row.size() + J.outs.size();
  last_slots = J.local.size();  // SSA values of the kernel (they live in registers)
  // 3. structural key: identical tapes (every training step) reuse the compiled kernel
  std::string key;
  key.reserve(J.order.size() * 24 + 64);
  key += dtype == GADCU_F32 ? "f32" : "f64";
  key += "|V" + std::to_string(J.V) + "U" + std::to_string(J.U) + "B" + std::to_string(J.block) + (J.idx32 ? "i32" : "i64") + (J.flat ? "flat" : "") + (J.ring ? "ring" + std::to_string(J.ring) : "") + (J.red ? "red" + std::to_string(J.red) : "");
  auto L = [&](int32_t r) { return r < 0 ? std::string("-") : std::to_string(J.local.at(r)); };
  auto list = [&](const char* tag, const std::vector<std::pair<int32_t, int>>& v) {
    key += tag;
    for (auto& d : v) key += L(d.first) + ",";
  };
  list("|d", J.dense);
  list("|u", J.uniform);
  list("|c", J.col);
  list("|r", J.row);
  key += "|o";
  for (int32_t r : J.outs) key += L(r) + ",";
  key += "|";
  for (int32_t r : J.order) {
    const SymInstr& I = instrs[r];
    key += L(r) + ":" + std::to_string(I.op) + "(" + L(I.a) + "," + L(I.b) + "," + L(I.c) + "," + L(I.d) + ");";
  }
  auto build = [&](const std::string& k) {
    const jit::Kernel* kernel = jit::find(k);
    if (!kernel) {
      ctx->segment_flush();  // (compiling and loading a module is not something to do inside a recorded segment)
      g_plan = &J;
      const std::string src = gen_source(*this, J, imm_slot);
      g_plan = nullptr;
      if (const char* dump = getenv("GADCU_JIT_DUMP")) {
        FILE* f = fopen(dump, "a");
        if (f) { fprintf(f, "// ---- key %zu bytes, %zu instructions\n%s\n", k.size(), J.order.size(), src.c_str()); fclose(f); }
      }
      kernel = jit::compile(k, src, "lane_program");
    }
    return kernel;
  };
  const jit::Kernel* kernel = nullptr;
  if (J.red) {
    // reduce.cu cuts a long segment into 8 chunks per SM: compiled for 8 resident blocks (32 registers) they all run at
    // once; if the chain does not fit that budget, whatever the compiler needs (the chunks then take two rounds)
    static const int want = [] { const char* e = getenv("GADCU_REDUCE_MINBLOCKS"); return e ? atoi(e) : 8; }();
    const std::vector<int> candidates = J.ring ? std::vector<int>{4, 3, 0} : want > 0 ? std::vector<int>{want, 0} : std::vector<int>{0};
    for (size_t c = 0; c < candidates.size(); c++) {
      J.stage_blocks = candidates[c];
      kernel = build(key + "|b" + std::to_string(J.stage_blocks));
      if (kernel->local_bytes == 0) break;
    }
  } else if (J.ring) {
    // the staged kernel is compiled for as many resident blocks per SM as its registers allow without spilling:
    // 5 (<= 80 registers), else 3 (<= 136), else 2; the choice is remembered per structure
    static std::mutex chosen_mu;
    static std::unordered_map<std::string, int> chosen;
    static const int forced = [] { const char* e = getenv("GADCU_LANE_STAGE_BLOCKS"); return e ? atoi(e) : 0; }();
    int known = 0;
    {
      std::lock_guard<std::mutex> lk(chosen_mu);
      auto it = chosen.find(key);
      if (it != chosen.end()) known = it->second;
    }
    const std::vector<int> candidates = known ? std::vector<int>{known} : (forced > 0 && J.flat) ? std::vector<int>{forced}
                                              : J.flat ? std::vector<int>{5, 3, 2} : std::vector<int>{2, 1};  // (array mode: 544-thread blocks, <= 56 registers for two per SM)
    for (size_t c = 0; c < candidates.size(); c++) {
      J.stage_blocks = candidates[c];
      kernel = build(key + "|b" + std::to_string(J.stage_blocks));
      if (kernel->local_bytes == 0 || c + 1 == candidates.size()) break;
    }
    std::lock_guard<std::mutex> lk(chosen_mu);
    chosen[key] = J.stage_blocks;
  } else {
    kernel = build(key);
  }
  if (reusable && J.imms.size() * 8 + (J.in_ptrs.size() + J.outs.size() + J.tile_d0.size() + 1) * 8 <= kMaxParamBytes) {
    Compiled c;
    c.kernel = kernel;
    c.in_inputs = J.in_inputs;
    c.tile_d0 = J.tile_d0;
    c.imms = J.imms;
    c.V = J.V; c.U = J.U; c.block = J.block;
    c.smem_bytes = J.smem_bytes;
    c.stage_blocks = J.stage_blocks;
    c.instructions = last_instructions;
    c.values = last_slots;
    if (!pkey.words.empty()) {
      std::lock_guard<std::mutex> lk(g_plans_mu);
      if (g_plans.size() >= 256) g_plans.clear();  // (a process that keeps inventing tape structures: start over)
      g_plans[pkey] = c;
    }
    compiled[todo] = std::move(c);
  }
  // 4. outputs + parameter block (layout of `struct P`: every field is 8 bytes wide)
  std::vector<ArrayRef> out_cols;
  if (red) {
    const size_t NI = std::max<size_t>(J.in_ptrs.size(), 1), NT = J.tile_d0.size(), NM = std::max<size_t>(J.imms.size(), 1);
    std::vector<uint64_t> params(NI + 1 + 1 + NT + NM, 0);
    for (size_t i = 0; i < J.in_ptrs.size(); i++) params[i] = reinterpret_cast<uint64_t>(J.in_ptrs[i]);
    params[NI] = reinterpret_cast<uint64_t>(red->out);
    params[NI + 1] = lanes;
    for (size_t i = 0; i < NT; i++) params[NI + 2 + i] = J.tile_d0[i];
    for (size_t i = 0; i < J.imms.size(); i++) std::memcpy(&params[NI + 2 + NT + i], &J.imms[i], 8);
    if (params.size() * 8 > kMaxParamBytes) return {};
    ProfileScope prof_scope(ctx, PROF_REDUCE);
    void* args[] = {params.data()};
    jit::launch(ctx, kernel, unsigned(red->splits * red->outer), J.ring ? 288u : 256u, args, J.smem_bytes);
    if (red_done) *red_done = true;
    return {};
  }
  for (size_t j = 0; j < todo.size(); j++) out_cols.push_back(new_array(ctx, dtype, Dim4(lanes)));
  const size_t NI = std::max<size_t>(J.in_ptrs.size(), 1), NO = std::max<size_t>(J.outs.size(), 1), NT = std::max<size_t>(J.tile_d0.size(), 1),
               NM = std::max<size_t>(J.imms.size(), 1);
  std::vector<uint64_t> params(NI + NO + 1 + NT + NM, 0);
  for (size_t i = 0; i < J.in_ptrs.size(); i++) params[i] = reinterpret_cast<uint64_t>(J.in_ptrs[i]);
  for (size_t i = 0; i < out_cols.size(); i++) params[NI + i] = reinterpret_cast<uint64_t>(out_cols[i]->ptr());
  params[NI + NO] = lanes;
  for (size_t i = 0; i < J.tile_d0.size(); i++) params[NI + NO + 1 + i] = J.tile_d0[i];
  for (size_t i = 0; i < J.imms.size(); i++) std::memcpy(&params[NI + NO + 1 + NT + i], &J.imms[i], 8);
  if (params.size() * 8 > kMaxParamBytes) fail(GADCU_ERR_UNSUPPORTED, "lane program has too many inputs / outputs / constants for one kernel");
  const uint64_t per_block = uint64_t(J.block) * uint64_t(J.V) * uint64_t(J.U);
  uint64_t grid = (lanes + per_block - 1) / per_block;
  const uint64_t cap = uint64_t(ctx->sm_count) * (J.block == 256 ? 8 : 16);
  if (grid > cap && J.V > 1) grid = cap;          // grid-stride for the streaming kernels
  if (J.ring) grid = std::min<uint64_t>(grid, staged_grid(ctx, J.smem_bytes, J.stage_blocks));
  if (grid > 0x7fffffffull) grid = 0x7fffffffull;
  if (grid == 0) grid = 1;
  {
    ProfileScope prof_scope(ctx, array_mode ? PROF_EW : PROF_OTHER);
    void* args[] = {params.data()};
    jit::launch(ctx, kernel, unsigned(grid), unsigned(J.block + (J.ring ? 32 : 0)), args, J.smem_bytes);
  }
  return out_cols;
}

// ---- deferred products and their fused tails (array mode) ------------------------------------------------------
ArrayRef SymProgram::launch_pending(const PendingGemm& pg, const k::GemmEpilogue* epi_in, bool with_split) {
  const ArrayRef &a = pg.a, &b = pg.b;
  k::GemmArgs g;
  g.dt = GADCU_F32;
  g.A = a->ptr();
  g.B = b->ptr();
  g.M = pg.rd[0];
  g.N = pg.rd[1];
  g.K = pg.p1.transposed ? a->dims[0] : a->dims[1];
  g.ta = pg.p1.transposed;
  g.tb = pg.p2.transposed;
  g.lda = a->dims[0];
  g.ldb = b->dims[0];
  g.strideC = g.M * g.N;
  if (a->ptr() == a->buf->ptr) g.a_buf = a->buf.get();
  if (b->ptr() == b->buf->ptr) g.b_buf = b->buf.get();
  ArrayRef out = new_array(ctx, GADCU_F32, pg.rd);
  g.C = out->ptr();
  k::GemmEpilogue epi;
  ArrayRef split;
  if (epi_in) {
    epi = *epi_in;
    if (with_split && g.M % 32 == 0) {
      // the result is the operand of the next tensor-core products: its [hi | lo] copies are written by the same epilogue
      // and left on the buffer where split_operand looks for them
      const uint64_t elems = g.M * g.N;
      split = new_array(ctx, GADCU_F32, Dim4(2 * elems));
      epi.hi = static_cast<float*>(split->ptr());
      epi.lo = epi.hi + elems;
      Buffer* owner = out->buf.get();
      owner->split = split->buf.get();
      owner->split->retain();
      owner->split_epoch = ctx->epoch.load(std::memory_order_relaxed);
      owner->split_elems = elems;
      owner->split_rows = g.M;
    }
    g.epi = &epi;
  }
  k::launch_gemm(ctx, g);
  return out;
}

void SymProgram::resolve_pending(const std::vector<int32_t>& todo) {
  const int32_t n = int32_t(instrs.size());
  auto leaf = [&](int32_t r) {
    const uint16_t op = instrs[r].op;
    return op == S_LOAD || op == S_LOADU || op == S_LOADCOL || op == S_LOADROW || op == S_CONST || op == S_GEMM || cache.count(r) != 0;
  };
  // what the request depends on
  std::vector<char> needed(n, 0);
  std::vector<int32_t> stack(todo.begin(), todo.end());
  while (!stack.empty()) {
    const int32_t r = stack.back();
    stack.pop_back();
    if (needed[r]) continue;
    needed[r] = 1;
    if (leaf(r)) continue;
    const SymInstr& I = instrs[r];
    for (int32_t x : {I.a, I.b, I.c, I.d})
      if (x >= 0 && !needed[x]) stack.push_back(x);
  }
  std::vector<int32_t> gemms;
  for (auto& kv : pending)
    if (needed[kv.first] && !cache.count(kv.first)) gemms.push_back(kv.first);
  if (gemms.empty()) return;
  std::sort(gemms.begin(), gemms.end());
  static const bool fuse_on = [] { const char* e = getenv("GADCU_EPILOGUE"); return !e || atoi(e) != 0; }();
  static const bool split_on = [] { const char* e = getenv("GADCU_EPILOGUE_SPLIT"); return !e || atoi(e) != 0; }();
  // a dense [lanes] operand the epilogue can read by pointer: a computed column or a dense load
  auto dense_ptr = [&](int32_t r) -> const float* {
    auto c = cache.find(r);
    if (c != cache.end()) return static_cast<const float*>(c->second->ptr());
    if (instrs[r].op == S_LOAD) return static_cast<const float*>(inputs[instrs[r].a]->ptr());
    return nullptr;
  };
  // the bias of an add(matmul, tile_as(b, [M, N])): a [1, N] array read as column tile with d0 == M
  auto bias_ptr = [&](int32_t r, const PendingGemm& pg) -> const float* {
    const SymInstr& I = instrs[r];
    if (I.op != S_LOADCOL || cache.count(r)) return nullptr;
    if (input_d0[I.a] != pg.rd[0] || inputs[I.a]->phys.elements() != pg.rd[1]) return nullptr;
    return static_cast<const float*>(inputs[I.a]->ptr());
  };
  for (int32_t g : gemms) {
    const PendingGemm& pg = pending.at(g);
    // every needed instruction that reads g, and what reads those: the product may be fused only if all of that is the
    // inside of ONE known tail ending in a requested register
    std::vector<int32_t> readers;
    for (int32_t r = 0; r < n; r++) {
      if (!needed[r] || leaf(r)) continue;
      const SymInstr& I = instrs[r];
      if (I.a == g || I.b == g || I.c == g || I.d == g) readers.push_back(r);
    }
    k::GemmEpilogue epi;
    int32_t result_reg = -1;
    std::vector<int32_t> inside;  // registers of the tail other than its result
    bool split = false;
    if (fuse_on && pg.fusable && readers.size() == 1) {
      const int32_t r1 = readers[0];
      const SymInstr& I1 = instrs[r1];
      auto other = [&](const SymInstr& I) { return I.a == g ? I.b : I.a; };
      auto sole_reader = [&](int32_t x) {  // the one needed instruction that reads x, or -1
        int32_t found = -1;
        for (int32_t r = 0; r < n; r++) {
          if (!needed[r] || leaf(r)) continue;
          const SymInstr& I = instrs[r];
          if (I.a == x || I.b == x || I.c == x || I.d == x) {
            if (found >= 0) return int32_t(-1);
            found = r;
          }
        }
        return found;
      };
      const bool r1_requested = std::binary_search(todo.begin(), todo.end(), r1);
      if (I1.op == S_ADD && !r1_requested) {
        const float* bias = bias_ptr(other(I1), pg);
        const int32_t r2 = bias ? sole_reader(r1) : -1;
        if (r2 >= 0 && std::binary_search(todo.begin(), todo.end(), r2)) {
          const SymInstr& I2 = instrs[r2];
          if (I2.op == S_TANH && I2.a == r1) {  // tanh(add(matmul, tile_as(b)))
            epi.kind = k::GemmEpilogue::BIAS_TANH;
            epi.bias = bias;
            result_reg = r2;
            inside = {r1};
            split = true;
          } else if (I2.op == S_SUB && I2.b == r1 && I2.a != r1 && dense_ptr(I2.a)) {  // sub(target, add(matmul, tile_as(b)))
            epi.kind = k::GemmEpilogue::BIAS_RESIDUAL;
            epi.bias = bias;
            epi.aux = dense_ptr(I2.a);
            result_reg = r2;
            inside = {r1};
          }
        }
      } else if (I1.op == S_MUL && r1_requested && I1.a != I1.b) {
        // mul(g, addc(neg0(mul(t, t)), 1)): the tanh VJP (analytic.rs:385-395) on an incoming product
        const int32_t e = other(I1);
        if (e >= 0 && !leaf(e) && instrs[e].op == S_ADDC && instrs[e].imm == 1.0) {
          const int32_t ng = instrs[e].a;
          if (ng >= 0 && !leaf(ng) && instrs[ng].op == S_NEG0) {
            const int32_t sq = instrs[ng].a;
            if (sq >= 0 && !leaf(sq) && instrs[sq].op == S_MUL && instrs[sq].a == instrs[sq].b && dense_ptr(instrs[sq].a)) {
              epi.kind = k::GemmEpilogue::TANH_GRAD;
              epi.aux = dense_ptr(instrs[sq].a);
              result_reg = r1;
              split = true;
            }
          }
        }
      }
    }
    if (result_reg >= 0) {
      k::GemmArgs probe;  // (shape check only)
      probe.dt = GADCU_F32;
      probe.M = pg.rd[0]; probe.N = pg.rd[1];
      probe.K = pg.p1.transposed ? pg.a->dims[0] : pg.a->dims[1];
      probe.ta = pg.p1.transposed; probe.tb = pg.p2.transposed;
      probe.lda = pg.a->dims[0]; probe.ldb = pg.b->dims[0];
      if (!k::gemm_epilogue_supported(probe)) result_reg = -1;
    }
    if (result_reg >= 0) {
      cache[result_reg] = launch_pending(pg, &epi, split && split_on);
      pinned.insert(result_reg);
    } else {
      cache[g] = launch_pending(pg, nullptr, false);
      pinned.insert(g);
    }
  }
}

// matmul recorded instead of launched (EvalAlgebra::matmul when deferred regions are on): the product joins the region of
// its [M, N] result as a leaf, so that the elementwise ops a layer puts behind it can ride in the GEMM's epilogue.
ArrayRef lazy_matmul(gadcu_ctx* ctx, const ArrayRef& a, const ArrayRef& b, MatProp p1, MatProp p2, const Dim4& rd, bool fusable);

// ---- the symbolic per-lane eval algebra ---------------------------------------------------------
class SymAlgebra final : public Algebra {
 public:
  SymAlgebra(gadcu_ctx* c, std::shared_ptr<SymProgram> p) : Algebra(c), prog_(std::move(p)) {}
  gadcu_algebra_kind kind() const override { return GADCU_EVAL; }
  gadcu_algebra* clone() const override { return new SymAlgebra(ctx, prog_); }
  Value link(const Value& v) override { return Value(v.data); }
  const std::shared_ptr<SymProgram>& program() const { return prog_; }

  // register of a value: symbolic values carry it; real arrays become (deduplicated) loads
  int32_t reg(const Value& v) {
    if (!v.data) fail(GADCU_ERR_INVALID, "batched tape: value has no data");
    const ArrayRef& a = v.data;
    if (a->sym_reg >= 0 && a->sym) {
      if (program_ptr(a.get()) != prog_.get()) fail(GADCU_ERR_INVALID, "batched tape: value belongs to another tape");
      return a->sym_reg;
    }
    if (a->dtype != prog_->dtype) fail(GADCU_ERR_TYPE, "batched tape: dtype mismatch");
    ArrayRef dense = a->bcast1() ? a : materialize(a);
    const bool uniform = dense->phys.elements() == 1;
    if (!uniform && dense->elements() != prog_->lanes)
      fail(GADCU_ERR_DIMENSIONS, "batched tape: expected a [lanes] column or a scalar, got " + dense->dims.str());
    const auto key = std::make_tuple(static_cast<const void*>(dense->ptr()), uniform ? 1 : 0, uint64_t(0));
    auto it = prog_->input_reg.find(key);
    if (it != prog_->input_reg.end()) return it->second;
    prog_->inputs.push_back(dense);
    prog_->input_d0.push_back(0);
    SymInstr I{uint16_t(uniform ? S_LOADU : S_LOAD)};
    I.a = int32_t(prog_->inputs.size() - 1);
    int32_t r = prog_->emit(I);
    prog_->input_reg[key] = r;
    return r;
  }
  Value load(const ArrayRef& a) { return Value(sym_array(prog_, reg(Value(a)))); }
  Value op(SymOp o, int32_t a = -1, int32_t b = -1, int32_t c = -1, int32_t d = -1, double imm = 0.0) {
    SymInstr I{uint16_t(o)};
    I.a = a; I.b = b; I.c = c; I.d = d; I.imm = imm;
    return Value(sym_array(prog_, prog_->emit(I)));
  }
  static double cval(double c, bool is_int) { return is_int ? double(int16_t(c)) : c; }

  // core.rs:95-112
  Value variable(ArrayRef data) override { return load(data); }
  Value constant(ArrayRef data) override { return load(data); }
  Value add(const Value& a, const Value& b) override { return op(S_ADD, reg(a), reg(b)); }
  // arith.rs:105-130
  Value sub(const Value& a, const Value& b) override { return op(S_SUB, reg(a), reg(b)); }
  Value mul(const Value& a, const Value& b) override { return op(S_MUL, reg(a), reg(b)); }
  Value zeros(const Value&) override { return op(S_CONST, -1, -1, -1, -1, 0.0); }
  Value ones(const Value&) override { return op(S_CONST, -1, -1, -1, -1, 1.0); }
  Value neg(const Value& v) override { return op(S_NEG, reg(v)); }
  // const_arith.rs:81-104
  Value setc(const Value&, double c, bool is_int) override { return op(S_CONST, -1, -1, -1, -1, cval(c, is_int)); }
  Value addc(const Value& v, double c, bool is_int) override { return op(S_ADDC, reg(v), -1, -1, -1, cval(c, is_int)); }
  Value mulc(const Value& v, double c, bool is_int) override { return op(S_MULC, reg(v), -1, -1, -1, cval(c, is_int)); }
  Value powc(const Value& v, double c, bool is_int) override {
    return is_int ? op(S_POWI, reg(v), -1, -1, -1, double(int16_t(c))) : op(S_POWF, reg(v), -1, -1, -1, c);
  }
  // analytic.rs:188-246
  Value exp(const Value& v) override { return op(S_EXP, reg(v)); }
  Value log(const Value& v) override { return op(S_LOG, reg(v)); }
  Value log1p(const Value& v) override { return op(S_LOG1P, reg(v)); }
  Value sin(const Value& v) override { return op(S_SIN, reg(v)); }
  Value cos(const Value& v) override { return op(S_COS, reg(v)); }
  Value tanh(const Value& v) override { return op(S_TANH, reg(v)); }
  Value sigmoid(const Value& v) override { return op(S_SIGMOID, reg(v)); }
  Value reciprocal(const Value& v) override { return op(S_RECIP, reg(v)); }
  Value sqrt(const Value& v) override { return op(S_SQRT, reg(v)); }
  Value div(const Value& a, const Value& b) override { return op(S_DIV, reg(a), reg(b)); }
  Value pow(const Value& a, const Value& b) override { return op(S_POW, reg(a), reg(b)); }
  // compare.rs:148-175
  Value min(const Value& a, const Value& b) override { return op(S_MIN, reg(a), reg(b)); }
  Value max(const Value& a, const Value& b) override { return op(S_MAX, reg(a), reg(b)); }
  Value select_argmax(const Value& v0, const Value& v1, const Value* r0, const Value* r1) override {
    if (r0 && r1) return op(S_SELECT, reg(v0), reg(v1), reg(*r0), reg(*r1));
    if (r0) return op(S_SELECT_R0, reg(v0), reg(v1), reg(*r0));
    if (r1) return op(S_SELECT_R1, reg(v0), reg(v1), reg(*r1));
    return op(S_CONST, -1, -1, -1, -1, 0.0);
  }
  // array / matrix operators have no meaning on a scalar tape
#define GADCU_SYM_UNSUPPORTED(SIG) \
  Value SIG override { fail(GADCU_ERR_UNSUPPORTED, "array operator on a batched scalar tape"); }
  GADCU_SYM_UNSUPPORTED(flat(const Value&))
  GADCU_SYM_UNSUPPORTED(moddims(const Value&, const Dim4&))
  GADCU_SYM_UNSUPPORTED(tile_as(const Value&, const Dim4&))
  GADCU_SYM_UNSUPPORTED(sum_as(const Value&, const Dim4&))
  GADCU_SYM_UNSUPPORTED(constant_as(const Value&, const Dim4&))
  GADCU_SYM_UNSUPPORTED(as_scalar(const Value&))
  GADCU_SYM_UNSUPPORTED(scale(const Value&, const Value&))
  GADCU_SYM_UNSUPPORTED(dot(const Value&, const Value&))
  GADCU_SYM_UNSUPPORTED(matmul(const Value&, const Value&, MatProp, MatProp))
  GADCU_SYM_UNSUPPORTED(transpose(const Value&, bool))
  GADCU_SYM_UNSUPPORTED(max_as(const Value&, const Dim4&))
  GADCU_SYM_UNSUPPORTED(argmax_as(const Value&, const Dim4&))
  GADCU_SYM_UNSUPPORTED(softmax_as(const Value&, const Dim4&))
#undef GADCU_SYM_UNSUPPORTED

 private:
  std::shared_ptr<SymProgram> prog_;
};

}  // namespace

bool SymProgram::run_reduce(int32_t reg, ReduceSpec& spec) {
  std::lock_guard<std::mutex> lk(mu);
  if (!array_mode || cache.count(reg)) return false;
  if (!pending.empty()) {
    resolve_pending({reg});  // products the value depends on are launched now; if that settled the value itself, it exists
    if (cache.count(reg)) return false;
  }
  bool done = false;
  run_jit({reg}, &spec, &done);
  return done;
}

// sum_as / max_as over leading axes of a value that is still a deferred elementwise chain: one kernel computes the chain
// and reduces it, the value itself is never written (C3 forward: d = exp(x) log(y) + x only feeds the sum; an MLP loss:
// the squared residual only feeds the sum). Returns null when the value exists already or the shape is not covered.
ArrayRef lazy_reduce(gadcu_ctx* ctx, k::RedOp op, const ArrayRef& v, uint64_t reduce, uint64_t outer, const Dim4& out_dims) {
  static const bool on = [] { const char* e = getenv("GADCU_FUSE_REDUCE"); return !e || atoi(e) != 0; }();
  if (!on || !v->symbolic() || !v->sym || reduce == 0 || outer == 0 || outer > 65535) return ArrayRef();
  std::shared_ptr<SymProgram> prog = program_of(v.get());
  if (!prog || !prog->array_mode || prog->lanes != reduce * outer || !lazy_enabled(ctx, prog->lanes, false)) return ArrayRef();
  SymProgram::ReduceSpec spec;
  spec.op = op;
  spec.reduce = reduce;
  spec.outer = outer;
  k::reduce_contig_plan(ctx, prog->dtype, reduce, outer, &spec.splits, &spec.chunk);
  if (spec.splits * outer > 0x7fffffffull) return ArrayRef();
  spec.vec = reduce % uint64_t(prog->dtype == GADCU_F32 ? 4 : 2) == 0;
  ArrayRef out = new_array(ctx, prog->dtype, out_dims);
  ArrayRef partials = spec.splits > 1 ? new_array(ctx, prog->dtype, Dim4(spec.splits * outer)) : out;
  spec.out = partials->ptr();
  if (!prog->run_reduce(v->sym_reg, spec)) return ArrayRef();
  if (spec.splits > 1) k::launch_reduce_partials(ctx, prog->dtype, op, partials->ptr(), out->ptr(), spec.splits, outer);
  return out;
}

ArrayRef force_symbolic(const ArrayRef& a) {
  if (!a->symbolic()) return a;
  std::shared_ptr<SymProgram> prog = program_of(a.get());
  ArrayRef col = prog->run({a->sym_reg})[0];
  if (!prog->array_mode) return col;
  // an array value: from now on the handle IS the dense array (same bytes; later uses read it, and the program
  // loads it instead of recomputing)
  a->buf = col->buf;  // (the handle keeps its register: later ops of the same region reuse the column)
  return a;
}

// =================================================================================================
// deferred elementwise regions of array tapes
// =================================================================================================
namespace {

std::shared_ptr<SymProgram> lazy_program(gadcu_ctx* ctx, gadcu_dtype dt, uint64_t n) {
  const uint64_t epoch = ctx->epoch.load(std::memory_order_relaxed);
  std::lock_guard<std::mutex> lk(ctx->mu);
  auto key = std::make_pair(n, int(dt));
  auto it = ctx->lazy_programs.find(key);
  if (it != ctx->lazy_programs.end()) {
    auto prog = std::static_pointer_cast<SymProgram>(it->second);
    // a new tape starts a new program (the old one dies with the last handle into it and releases its inputs);
    // so does a program that has grown long outside any tape
    if (prog->epoch == epoch && prog->instrs.size() < 8192) return prog;
  }
  // drop programs of finished tapes from the table (their handles keep them alive as long as needed)
  for (auto i = ctx->lazy_programs.begin(); i != ctx->lazy_programs.end();) {
    if (std::static_pointer_cast<SymProgram>(i->second)->epoch != epoch) i = ctx->lazy_programs.erase(i);
    else ++i;
  }
  auto prog = std::make_shared<SymProgram>();
  prog->ctx = ctx;
  prog->dtype = dt;
  prog->lanes = n;
  prog->array_mode = true;
  prog->epoch = epoch;
  ctx->lazy_programs[key] = prog;
  if (ctx->lazy_all.size() >= 64) {  // (drop the records of programs that have died since; lazy_before_write prunes too)
    size_t keep = 0;
    for (auto& w : ctx->lazy_all)
      if (!w.expired()) ctx->lazy_all[keep++] = w;
    ctx->lazy_all.resize(keep);
  }
  ctx->lazy_all.emplace_back(prog);
  return prog;
}

// register of an operand of an n-element elementwise op
int32_t lazy_operand(const std::shared_ptr<SymProgram>& prog, const ArrayRef& a_in) {
  ArrayRef a = a_in;
  if (a->symbolic()) {
    if (program_ptr(a.get()) == prog.get()) return a->sym_reg;
    a = materialize(a);  // a value of an earlier tape / another size class: read its bytes
  }
  if (a->sym_reg >= 0 && a->sym && program_ptr(a.get()) == prog.get()) return a->sym_reg;  // forced earlier: the program loads the column
  int kind = 0;  // 0 dense, 1 uniform, 2 column tile ([1,D] over [B,D]), 3 row tile ([B,1] over [B,D])
  uint64_t d0 = 0;
  if (a->bcast1() || (a->elements() == 1 && prog->lanes != 1)) {
    kind = 1;
  } else if (a->lazy()) {
    const Dim4 &d = a->dims, &ph = a->phys;
    if (ph[0] == 1 && ph[1] == d[1] && d[2] == 1 && d[3] == 1 && ph[2] == 1 && ph[3] == 1) { kind = 2; d0 = d[0]; }
    else if (ph[0] == d[0] && ph[1] == 1 && ph[2] == 1 && ph[3] == 1) { kind = 3; d0 = d[0]; }
    else a = materialize(a);
  }
  if (kind != 1 && a->elements() != prog->lanes) fail(GADCU_ERR_DIMENSIONS, "elementwise operand of " + a->dims.str() + " in a region of " + std::to_string(prog->lanes) + " elements");
  const auto key = std::make_tuple(static_cast<const void*>(a->ptr()), kind, d0);
  auto it = prog->input_reg.find(key);
  if (it != prog->input_reg.end()) return it->second;
  prog->inputs.push_back(a);
  prog->input_d0.push_back(d0);
  static const SymOp load_op[] = {S_LOAD, S_LOADU, S_LOADCOL, S_LOADROW};
  SymInstr I{uint16_t(load_op[kind])};
  I.a = int32_t(prog->inputs.size() - 1);
  int32_t r = prog->emit(I);
  prog->input_reg[key] = r;
  return r;
}

}  // namespace

bool lazy_enabled(gadcu_ctx* ctx, uint64_t n, bool is_scalar) {
  return ctx->fuse && ctx->lazy && !is_scalar && n >= ctx->lazy_threshold && jit::available();
}

// out = [acc +] op(ins...), recorded instead of launched. The instruction sequence of every op is the one its
// ew.cu functor evaluates (one rounding per step), so the value is bit-identical to the eager kernel's.
ArrayRef lazy_elementwise(gadcu_ctx* ctx, k::EwOp op, const ArrayRef& acc, const std::vector<const ArrayRef*>& ins, double c, double c2,
                          gadcu_dtype dt, const Dim4& dims, bool is_scalar) {
  std::shared_ptr<SymProgram> prog = lazy_program(ctx, dt, dims.elements());
  int32_t r = -1;
  {
  std::lock_guard<std::mutex> lk(prog->mu);
  int32_t x[4] = {-1, -1, -1, -1};
  for (size_t i = 0; i < ins.size() && i < 4; i++) x[i] = lazy_operand(prog, *ins[i]);
  const int32_t racc = acc ? lazy_operand(prog, acc) : -1;
  auto E = [&](SymOp o, int32_t a = -1, int32_t b = -1, int32_t cc = -1, int32_t d = -1, double imm = 0.0) {
    SymInstr I{uint16_t(o)};
    I.a = a; I.b = b; I.c = cc; I.d = d; I.imm = imm;
    return prog->emit(I);
  };
  // the functors take their constants as T(c)
  const double k = dt == GADCU_F32 ? double(float(c)) : c, k2 = dt == GADCU_F32 ? double(float(c2)) : c2;
  const int32_t a = x[0], b = x[1], cc = x[2], d = x[3];
  using O = k::EwOp;
  switch (op) {
    case O::ADD: r = E(S_ADD, a, b); break;
    case O::SUB: r = E(S_SUB, a, b); break;
    case O::MUL: case O::SCALE: case O::VJP_MUL: r = E(S_MUL, a, b); break;
    case O::DIV: case O::VJP_LOG: r = E(S_DIV, a, b); break;
    case O::POW: r = E(S_POW, a, b); break;
    case O::MIN: r = E(S_FMIN, a, b); break;
    case O::MAX: r = E(S_FMAX, a, b); break;
    case O::NEG: case O::VJP_NEG: r = E(S_NEG0, a); break;
    case O::EXP: r = E(S_EXP, a); break;
    case O::LOG: r = E(S_LOG, a); break;
    case O::LOG1P: r = E(S_LOG1PF, a); break;
    case O::SIN: r = E(S_SIN, a); break;
    case O::COS: r = E(S_COS, a); break;
    case O::TANH: r = E(S_TANH, a); break;
    case O::SIGMOID: r = E(S_SIGMOID, a); break;
    case O::RECIPROCAL: r = E(S_RECIP, a); break;
    case O::SQRT: r = E(S_SQRT, a); break;
    case O::COPY: r = a; break;
    case O::ADDC: r = E(S_ADDC, a, -1, -1, -1, k); break;
    case O::MULC: case O::VJP_MULC: r = E(S_MULC, a, -1, -1, -1, k); break;
    case O::POWC: r = E(S_POWF, a, -1, -1, -1, k); break;
    case O::SELECT_BOTH: r = E(S_SELECT, a, b, cc, d); break;
    case O::SELECT_R0: r = E(S_SELECT_R0, a, b, cc); break;
    case O::SELECT_R1: r = E(S_SELECT_R1, a, b, cc); break;
    case O::VJP_POWC: { int32_t e = E(S_POWF, b, -1, -1, -1, k2); int32_t f = E(S_MULC, e, -1, -1, -1, k); r = E(S_MUL, f, a); break; }
    case O::VJP_EXP: { int32_t e = E(S_EXP, b); r = E(S_MUL, a, e); break; }
    case O::VJP_LOG1P: { int32_t p1 = E(S_ADDC, b, -1, -1, -1, 1.0); r = E(S_DIV, a, p1); break; }
    case O::VJP_SIN: { int32_t e = E(S_COS, b); r = E(S_MUL, a, e); break; }
    case O::VJP_COS: { int32_t s = E(S_SIN, b); int32_t e = E(S_NEG0, s); r = E(S_MUL, a, e); break; }
    case O::VJP_TANH: { int32_t t = E(S_TANH, b); int32_t s = E(S_MUL, t, t); int32_t n = E(S_NEG0, s); int32_t e = E(S_ADDC, n, -1, -1, -1, 1.0); r = E(S_MUL, a, e); break; }
    case O::VJP_SIGMOID: { int32_t s = E(S_SIGMOID, b); int32_t n = E(S_NEG0, s); int32_t p1 = E(S_ADDC, n, -1, -1, -1, 1.0); int32_t e = E(S_MUL, s, p1); r = E(S_MUL, a, e); break; }
    case O::VJP_RECIPROCAL: { int32_t s = E(S_MUL, b, b); int32_t n = E(S_NEG0, s); int32_t e = E(S_RECIP, n); r = E(S_MUL, a, e); break; }
    case O::VJP_SQRT: { int32_t s = E(S_SQRT, b); int32_t s2 = E(S_MULC, s, -1, -1, -1, 2.0); int32_t e = E(S_RECIP, s2); r = E(S_MUL, a, e); break; }
    case O::VJP_DIV0: { int32_t rr = E(S_RECIP, b); r = E(S_MUL, a, rr); break; }
    case O::VJP_DIV1: { int32_t rr = E(S_RECIP, b); int32_t g0 = E(S_MUL, a, rr); int32_t e = E(S_MUL, g0, rr); int32_t e2 = E(S_MUL, e, cc); r = E(S_NEG0, e2); break; }
    case O::VJP_SELECT0: r = E(S_SELECT_R0, b, cc, a); break;
    case O::VJP_SELECT1: r = E(S_SELECT_R1, b, cc, a); break;
    default: fail(GADCU_ERR_INVALID, "unknown elementwise op");
  }
  if (racc >= 0) r = E(S_ADD, racc, r);  // store.rs:52: current + new
  }
  return sym_array(prog, r, dims, is_scalar);
}

// `fusable` = false on higher-order tapes (GraphN): there the values inside a tail (the product itself, product + bias) are
// captured by recorded backward nodes and ARE read later, by the next sweep; a product that ran fused has not kept them
// and would have to be launched a second time. Such products are still deferred (one nothing depends on is never
// launched) but run plain.
ArrayRef lazy_matmul(gadcu_ctx* ctx, const ArrayRef& a_in, const ArrayRef& b_in, MatProp p1, MatProp p2, const Dim4& rd, bool fusable) {
  // the program keeps the operands' BUFFERS, through handles of its own: a computed deferred value still carries the
  // handle of the program that computed it, and holding that here would tie the program to itself for good
  const ArrayRef a = view_array(a_in, a_in->dims, a_in->phys, a_in->is_scalar), b = view_array(b_in, b_in->dims, b_in->phys, b_in->is_scalar);
  std::shared_ptr<SymProgram> prog = lazy_program(ctx, GADCU_F32, rd.elements());
  int32_t r;
  {
    std::lock_guard<std::mutex> lk(prog->mu);
    SymInstr I{uint16_t(S_GEMM)};
    prog->instrs.push_back(I);  // (never value-numbered: two products are two launches)
    prog->live.push_back(0);
    r = int32_t(prog->instrs.size() - 1);
    prog->pending.emplace(r, SymProgram::PendingGemm{a, b, p1, p2, rd, fusable});
  }
  return sym_array(prog, r, rd, false);
}

// Before `buffer` is overwritten in place (WeightOps::add_assign, an upload, an all-reduce): compute every value
// that still has a handle and would read the old contents later — in EVERY program that can still be asked for
// values (the open ones and those of earlier tapes that live on through their handles). A program that loads the
// buffer is then RETIRED: it leaves the table of open programs, so nothing recorded after the write is appended to it.
// That is what keeps later ops honest: a new `exp(W)` opens a fresh program with a fresh load of W (no stale
// register, no value-numbering hit on the pre-write `exp`), and a value forced here is only ever read back as the
// bytes of its column (by another program), never recomputed from the overwritten leaf once its handle is gone.
void lazy_before_write(gadcu_ctx* ctx, const void* buffer) {
  std::vector<std::shared_ptr<SymProgram>> progs;
  {
    std::lock_guard<std::mutex> lk(ctx->mu);
    size_t keep = 0;
    for (auto& w : ctx->lazy_all) {
      std::shared_ptr<void> p = w.lock();
      if (!p) continue;
      progs.push_back(std::static_pointer_cast<SymProgram>(p));
      ctx->lazy_all[keep++] = w;
    }
    ctx->lazy_all.resize(keep);
  }
  for (auto& prog : progs) {
    std::vector<int32_t> todo;
    {
      std::lock_guard<std::mutex> lk(prog->mu);
      const int32_t n = int32_t(prog->instrs.size());
      std::vector<char> dep(n, 0);
      bool any = false;
      for (int32_t r = 0; r < n; r++) {
        const SymInstr& I = prog->instrs[r];
        if (I.op == S_LOAD || I.op == S_LOADU || I.op == S_LOADCOL || I.op == S_LOADROW) {
          dep[r] = prog->inputs[I.a]->ptr() == buffer;
        } else if (I.op == S_GEMM) {
          auto pg = prog->pending.find(r);  // a product not launched yet reads its operands when it is
          dep[r] = pg != prog->pending.end() && !prog->cache.count(r) && (pg->second.a->ptr() == buffer || pg->second.b->ptr() == buffer);
        } else if (I.op != S_CONST) {
          // (true data dependence, also THROUGH registers that already have a column: such a column goes back to the
          // pool with its last handle, and a live value above it would then be recomputed from the overwritten leaf)
          for (int32_t x : {I.a, I.b, I.c, I.d})
            if (x >= 0 && dep[x]) dep[r] = 1;
        }
        any = any || dep[r];
        if (dep[r] && prog->live[r] > 0 && !prog->cache.count(r) && !(I.op == S_LOAD || I.op == S_LOADU || I.op == S_LOADCOL || I.op == S_LOADROW))
          todo.push_back(r);
      }
      if (!any) continue;
    }
    if (!todo.empty()) prog->run(todo);
    {
      // (a forced column may still go back to the pool with its last handle: nothing can be appended to a retired
      // program, so no later instruction can ask for it to be recomputed from the overwritten leaf)
      std::lock_guard<std::mutex> lk(prog->mu);
      prog->retired = true;
    }
    std::lock_guard<std::mutex> lk(ctx->mu);
    for (auto i = ctx->lazy_programs.begin(); i != ctx->lazy_programs.end();) {
      if (i->second.get() == prog.get()) i = ctx->lazy_programs.erase(i);
      else ++i;
    }
  }
}

// End of a Graph1 sweep: the gradients of the tape's variables are what the caller reads — compute them, all
// those of one region in one kernel. Other store entries stay deferred until somebody reads them.
void lazy_flush(const std::vector<ArrayRef>& values) {
  std::map<SymProgram*, std::pair<std::shared_ptr<SymProgram>, std::vector<size_t>>> groups;
  for (size_t i = 0; i < values.size(); i++) {
    const ArrayRef& a = values[i];
    if (!a || !a->symbolic()) continue;
    std::shared_ptr<SymProgram> prog = program_of(a.get());
    if (!prog->array_mode) continue;
    auto& g = groups[prog.get()];
    g.first = prog;
    g.second.push_back(i);
  }
  for (auto& kv : groups) {
    std::vector<int32_t> regs;
    for (size_t i : kv.second.second) regs.push_back(values[i]->sym_reg);
    std::vector<ArrayRef> cols = kv.second.first->run(regs);
    for (size_t j = 0; j < regs.size(); j++) {
      values[kv.second.second[j]]->buf = cols[j]->buf;
    }
  }
}

std::atomic<uint32_t> g_instrs_hint{0};

std::unique_ptr<GraphAlgebra> make_batched(gadcu_ctx* ctx, gadcu_dtype dt, uint64_t lanes) {
  auto prog = std::make_shared<SymProgram>();
  prog->ctx = ctx;
  prog->dtype = dt;
  prog->lanes = lanes;
  prog->instrs.reserve(g_instrs_hint.load(std::memory_order_relaxed));  // (the last batched program's length: same tape every step)
  auto sym = std::make_shared<SymAlgebra>(ctx, prog);
  auto g = std::make_unique<GraphAlgebra>(ctx, false, sym);
  g->set_kind(GADCU_BATCHED1);
  return g;
}

// ---- sweep memo ---------------------------------------------------------------------------------------------------
// A gad program records the same tape every step (net_ext.rs:52), and the reverse sweep over a batched tape does nothing but
// append instructions to the lane program and fill the store with (id -> register): a pure function of the recorded forward
// instructions, the tape's nodes, the root and the seed's register. So that function is tabulated: the first sweep over a
// structure runs (GraphAlgebra::sweep through the symbolic algebra, as always) and leaves behind the instruction tail it
// appended, the store's (id, register) pairs and the nodes it visited; a later sweep over the same structure appends the
// tail, rebuilds the store from the pairs — every visited id present, as values of the lane program, exactly as the sweep
// leaves them — and clears the visited nodes if it is a consuming one. What it skips is the walk itself: heap, VJP
// dispatch, accumulation, a thousand handle objects (0.3 ms of the 0.41 ms a recorded Rosenbrock-64 call took).
struct SweepMemo {
  std::vector<SymInstr> tail;
  std::vector<std::pair<uint32_t, int32_t>> store_regs;  // ascending node index -> register of its gradient
};
std::mutex g_sweeps_mu;
std::unordered_map<PlanKey, SweepMemo, PlanKeyHash> g_sweeps;
std::atomic<uint64_t> g_sweeps_replayed{0}, g_sweeps_walked{0};

bool sweep_key(const SymProgram& P, const GraphAlgebra& g, Id id, const gadcu_array& seed, bool once, PlanKey& k) {
  const int32_t seed_reg = seed.sym_reg;
  uint64_t nodes[2];
  // a captured real array (a variable's column) is known to the program once the forward pass has loaded it; one the sweep
  // would be the first to load makes the sweep more than a function of the structure, and is not tabulated
  auto register_of = [&](const gadcu_array& a) -> int64_t {
    if (!a.buf || a.dtype != P.dtype) return -1;
    auto it = P.input_reg.find(std::make_tuple(static_cast<const void*>(a.ptr()), a.phys.elements() == 1 ? 1 : 0, uint64_t(0)));
    return it == P.input_reg.end() ? -1 : it->second;
  };
  if (P.array_mode || !g.structure_hash(nodes, register_of)) return false;
  k.words.clear();
  k.words.reserve(P.instrs.size() * 4 + 8);
  k.words.push_back(uint64_t(P.dtype) | uint64_t(once) << 8 | uint64_t(id.index) << 16);
  k.words.push_back(uint64_t(uint32_t(seed_reg)) | uint64_t(P.instrs.size()) << 32);
  k.words.push_back(nodes[0]);
  k.words.push_back(nodes[1]);
  k.words.push_back(uint64_t(seed.dtype) | uint64_t(seed.is_scalar) << 8);  // (what the sweep checks the root's gradient for)
  for (int i = 0; i < 4; i++) k.words.push_back(seed.dims[i]);
  for (const SymInstr& I : P.instrs) {
    uint64_t bits;
    std::memcpy(&bits, &I.imm, 8);
    k.words.push_back(uint64_t(I.op) | uint64_t(uint32_t(I.a)) << 16);
    k.words.push_back(uint64_t(uint32_t(I.b)) | uint64_t(uint32_t(I.c)) << 32);
    k.words.push_back(uint64_t(uint32_t(I.d)));
    k.words.push_back(bits);
  }
  uint64_t h = 1469598103934665603ull;
  for (uint64_t w : k.words) { h ^= w; h *= 1099511628211ull; }
  k.hash = h;
  return true;
}

std::unique_ptr<Store> batched_evaluate_gradients(GraphAlgebra& g, Id id, ArrayRef seed, bool once) {
  auto* sym = dynamic_cast<SymAlgebra*>(&g.eval());
  if (!sym) fail(GADCU_ERR_INVALID, "not a batched tape");
  if (id.index == 0 || id.index > g.num_nodes()) fail(GADCU_ERR_MISSING_NODE, "Trying to obtain a node from an incorrect `id`.");
  Value sseed = seed->sym_reg >= 0 ? Value(seed) : sym->load(seed);
  // which ids are variables must be asked before a consuming sweep clears the nodes
  std::vector<Id> var_ids;
  for (uint32_t i = 1; i <= g.num_nodes(); i++) {
    Id vid{id.arena, i};
    if (g.is_variable(vid)) var_ids.push_back(vid);
  }
  static const bool memo_on = [] { const char* e = getenv("GADCU_SWEEP_MEMO"); return !e || atoi(e) != 0; }();
  const std::shared_ptr<SymProgram>& prog0 = sym->program();
  PlanKey skey;
  const bool keyed = memo_on && sseed.data->sym_reg >= 0 && sweep_key(*prog0, g, id, *sseed.data, once, skey);
  std::unique_ptr<Store> store;
  if (keyed) {
    const SweepMemo* hit = nullptr;
    SweepMemo copy;
    {
      std::lock_guard<std::mutex> lk(g_sweeps_mu);
      auto it = g_sweeps.find(skey);
      if (it != g_sweeps.end()) { copy = it->second; hit = &copy; }
    }
    if (hit) {
      {
        std::lock_guard<std::mutex> lk(prog0->mu);
        prog0->instrs.insert(prog0->instrs.end(), hit->tail.begin(), hit->tail.end());
      }
      store = std::make_unique<Store>();
      std::vector<uint32_t> visited;
      visited.reserve(hit->store_regs.size());
      for (auto& pr : hit->store_regs) visited.push_back(pr.first);
      // the entries stay (index, register) pairs until somebody looks at them (gadcu_store::realize)
      store->deferred = std::move(copy.store_regs);
      store->deferred_arena = id.arena;
      store->deferred_handle = [](const std::shared_ptr<void>& program, int32_t reg) {
        return sym_array(std::static_pointer_cast<SymProgram>(program), reg);
      };
      store->lane_program = prog0;
      if (once) g.consume(visited);
      g_sweeps_replayed++;
    }
  }
  if (!store) {
    const size_t n_before = prog0->instrs.size();
    store = g.evaluate_gradients(id, sseed.data, once);
    g_sweeps_walked++;
    if (keyed) {
      SweepMemo m;
      bool ok = true;
      for (auto& kv : store->values) {
        const ArrayRef& d = kv.second.data;
        if (kv.first.arena != id.arena || !d || !(d->sym_reg >= 0 && d->sym) || program_ptr(d.get()) != prog0.get() || d->buf) { ok = false; break; }
        m.store_regs.emplace_back(kv.first.index, d->sym_reg);
      }
      if (ok) {
        std::lock_guard<std::mutex> lk(prog0->mu);
        m.tail.assign(prog0->instrs.begin() + long(n_before), prog0->instrs.end());
        std::lock_guard<std::mutex> lk2(g_sweeps_mu);
        if (g_sweeps.size() >= 64) g_sweeps.clear();
        g_sweeps.emplace(std::move(skey), std::move(m));
      }
    }
  }
  // one kernel for the gradients of all variables
  std::vector<int32_t> regs;
  std::vector<Id> ids;
  for (Id vid : var_ids) {
    if (const int32_t* reg = store->deferred_slot(vid)) {  // (a replayed sweep: no handle yet, and none needed)
      regs.push_back(*reg);
      ids.push_back(vid);
      continue;
    }
    auto it = store->values.find(vid);
    if (it != store->values.end() && it->second.data->symbolic()) {
      regs.push_back(it->second.data->sym_reg);
      ids.push_back(vid);
    }
  }
  if (!regs.empty()) {
    std::vector<ArrayRef> cols = sym->program()->run(regs);
    for (size_t i = 0; i < ids.size(); i++) {
      if (int32_t* reg = store->deferred_slot(ids[i])) {
        *reg = -1;
        store->values.emplace(ids[i], Value(cols[i]));
      } else {
        store->values[ids[i]].data = cols[i];
      }
    }
  }
  store->program_instructions = sym->program()->last_instructions;
  store->program_slots = sym->program()->last_slots;
  // what a re-run needs: the program and, per variable (creation order), the register of its gradient
  store->lane_program = sym->program();
  g_instrs_hint.store(uint32_t(std::min<size_t>(sym->program()->instrs.size(), 1u << 16)), std::memory_order_relaxed);
  {
    // input slot of every variable (creation order): variables are loaded by buffer on first use
    auto prog = sym->program();
    std::lock_guard<std::mutex> lk(prog->mu);
    prog->variable_inputs.clear();
    for (const ArrayRef& data : g.variable_data()) {
      int32_t slot = -1;
      if (data && data->buf) {
        for (int kind = 0; kind < 2 && slot < 0; kind++) {
          auto it = prog->input_reg.find(std::make_tuple(static_cast<const void*>(data->ptr()), kind, uint64_t(0)));
          if (it != prog->input_reg.end()) slot = prog->instrs[it->second].a;  // (a load's operand is its input index)
        }
      }
      prog->variable_inputs.push_back(slot);  // -1: the tape never read this variable
    }
  }
  for (Id vid : var_ids) {
    auto it = std::find(ids.begin(), ids.end(), vid);
    store->lane_gradient_regs.push_back(it == ids.end() ? -1 : regs[size_t(it - ids.begin())]);
  }
  return store;
}

// The same tapes at new points: swap the variables' columns, run the compiled program again (one kernel launch, no
// re-recording). Returns the gradient column of every variable (null where the sweep left none).
std::vector<ArrayRef> batched_rerun(Store& store, const std::vector<ArrayRef>& columns) {
  auto prog = std::static_pointer_cast<SymProgram>(store.lane_program);
  if (!prog) fail(GADCU_ERR_INVALID, "rerun: the store does not come from a batched sweep");
  if (columns.size() != prog->variable_inputs.size())
    fail(GADCU_ERR_LENGTHS, "rerun: expected one column per variable (" + std::to_string(prog->variable_inputs.size()) + "), got " +
                                std::to_string(columns.size()));
  std::vector<int32_t> regs;
  {
    std::lock_guard<std::mutex> lk(prog->mu);
    for (size_t i = 0; i < columns.size(); i++) {
      ArrayRef c = materialize(columns[i]);
      if (c->dtype != prog->dtype) fail(GADCU_ERR_TYPE, "rerun: dtype mismatch");
      if (c->elements() != prog->lanes) fail(GADCU_ERR_DIMENSIONS, "rerun: expected a [lanes] column, got " + c->dims.str());
      if (prog->variable_inputs[i] >= 0) prog->inputs[size_t(prog->variable_inputs[i])] = c;
    }
    prog->cache.clear();  // every computed column belonged to the old points
    for (int32_t r : store.lane_gradient_regs)
      if (r >= 0) regs.push_back(r);
  }
  std::vector<ArrayRef> cols = prog->run(regs);
  std::vector<ArrayRef> out(store.lane_gradient_regs.size());
  size_t j = 0;
  for (size_t i = 0; i < out.size(); i++)
    if (store.lane_gradient_regs[i] >= 0) out[i] = cols[j++];
  store.program_instructions = prog->last_instructions;
  return out;
}

}  // namespace gadcu

using namespace gadcu;

extern "C" {

gadcu_status gadcu_batched_create(gadcu_ctx* ctx, gadcu_dtype dt, uint64_t lanes, gadcu_algebra** out) {
  return guard([&] {
    if (lanes == 0) fail(GADCU_ERR_EMPTY, "batched tape with zero lanes");
    *out = make_batched(ctx, dt, lanes).release();
  });
}
gadcu_status gadcu_batched_force(gadcu_algebra* g, const gadcu_value* v, gadcu_array** out) {
  return guard([&] {
    (void)g;
    ArrayRef a(v->data, true);
    *out = materialize(a).detach();
  });
}
gadcu_status gadcu_batched_rerun(gadcu_store* s, gadcu_array* const* columns, size_t n, gadcu_array** gradients_out) {
  return guard([&] {
    std::vector<ArrayRef> cols;
    for (size_t i = 0; i < n; i++) cols.emplace_back(columns[i], true);
    std::vector<ArrayRef> grads = batched_rerun(*s, cols);
    for (size_t i = 0; i < grads.size(); i++) gradients_out[i] = grads[i] ? grads[i].detach() : nullptr;
  });
}
gadcu_status gadcu_batched_sweep_memo_stats(uint64_t* replayed, uint64_t* walked, uint64_t* structures) {
  if (replayed) *replayed = g_sweeps_replayed.load();
  if (walked) *walked = g_sweeps_walked.load();
  if (structures) {
    std::lock_guard<std::mutex> lk(g_sweeps_mu);
    *structures = g_sweeps.size();
  }
  return GADCU_OK;
}
gadcu_status gadcu_batched_program_stats(const gadcu_store* s, uint64_t* instructions, uint64_t* slots) {
  if (instructions) *instructions = s->program_instructions;
  if (slots) *slots = s->program_slots;
  return GADCU_OK;
}

}  // extern "C"
