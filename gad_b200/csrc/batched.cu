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
#include <cmath>
#include <algorithm>
#include <cstring>
#include <unordered_map>

#include "batched.h"

namespace gadcu {
namespace {

enum SymOp : uint16_t {
  S_LOAD, S_LOADU, S_CONST, S_ADD, S_SUB, S_MUL, S_DIV, S_NEG, S_ADDC, S_MULC, S_POWI, S_POWF, S_POW, S_EXP, S_LOG, S_LOG1P,
  S_SIN, S_COS, S_TANH, S_SIGMOID, S_RECIP, S_SQRT, S_SELECT, S_SELECT_R0, S_SELECT_R1, S_MIN, S_MAX, S_STORE
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
  std::vector<SymInstr> instrs;
  std::vector<ArrayRef> inputs;                      // real arrays: [lanes] columns or 1-element uniforms
  std::unordered_map<const void*, int32_t> input_reg;  // dedupe loads by buffer
  std::unordered_map<int32_t, ArrayRef> cache;       // register -> materialised column
  uint64_t last_instructions = 0, last_slots = 0;
  std::mutex mu;

  int32_t emit(SymInstr i) {
    instrs.push_back(i);
    return int32_t(instrs.size() - 1);
  }
  std::vector<ArrayRef> run(const std::vector<int32_t>& outs);
};

ArrayRef sym_array(const std::shared_ptr<SymProgram>& prog, int32_t reg) {
  auto* a = new gadcu_array;
  a->dims = Dim4();
  a->phys = Dim4();
  a->dtype = prog->dtype;
  a->is_scalar = true;
  a->ctx = prog->ctx;
  a->sym = prog;
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
  GADCU_CUDA(cudaMemcpyAsync(blob->ptr(), host.data(), host.size(), cudaMemcpyHostToDevice, ctx->stream));
  GADCU_CUDA(cudaStreamSynchronize(ctx->stream));  // `host` is pageable
  char* base = static_cast<char*>(blob->ptr());
  const DevInstr* dprog = reinterpret_cast<const DevInstr*>(base);
  const void* din = base + align(prog_bytes);
  void* dout = base + align(prog_bytes) + align(in_bytes);
  if (dtype == GADCU_F32) launch_interpreter<float>(ctx, next_slot, dprog, int(dev.size()), din, dout, lanes);
  else launch_interpreter<double>(ctx, next_slot, dprog, int(dev.size()), din, dout, lanes);

  for (size_t j = 0; j < todo.size(); j++) cache[todo[j]] = out_cols[j];
  for (size_t i = 0; i < outs.size(); i++)
    if (!result[i]) result[i] = cache[outs[i]];
  return result;
}

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
    if (a->sym_reg >= 0) {
      if (a->sym.get() != prog_.get()) fail(GADCU_ERR_INVALID, "batched tape: value belongs to another tape");
      return a->sym_reg;
    }
    if (a->dtype != prog_->dtype) fail(GADCU_ERR_TYPE, "batched tape: dtype mismatch");
    ArrayRef dense = a->bcast1() ? a : materialize(a);
    const bool uniform = dense->phys.elements() == 1;
    if (!uniform && dense->elements() != prog_->lanes)
      fail(GADCU_ERR_DIMENSIONS, "batched tape: expected a [lanes] column or a scalar, got " + dense->dims.str());
    auto it = prog_->input_reg.find(dense->ptr());
    if (it != prog_->input_reg.end()) return it->second;
    prog_->inputs.push_back(dense);
    SymInstr I{uint16_t(uniform ? S_LOADU : S_LOAD)};
    I.a = int32_t(prog_->inputs.size() - 1);
    int32_t r = prog_->emit(I);
    prog_->input_reg[dense->ptr()] = r;
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

ArrayRef force_symbolic(const ArrayRef& a) {
  if (!a->symbolic()) return a;
  auto prog = std::static_pointer_cast<SymProgram>(a->sym);
  return prog->run({a->sym_reg})[0];
}

std::unique_ptr<GraphAlgebra> make_batched(gadcu_ctx* ctx, gadcu_dtype dt, uint64_t lanes) {
  auto prog = std::make_shared<SymProgram>();
  prog->ctx = ctx;
  prog->dtype = dt;
  prog->lanes = lanes;
  auto sym = std::make_shared<SymAlgebra>(ctx, prog);
  auto g = std::make_unique<GraphAlgebra>(ctx, false, sym);
  g->set_kind(GADCU_BATCHED1);
  return g;
}

std::unique_ptr<Store> batched_evaluate_gradients(GraphAlgebra& g, Id id, ArrayRef seed, bool once) {
  auto* sym = dynamic_cast<SymAlgebra*>(&g.eval());
  if (!sym) fail(GADCU_ERR_INVALID, "not a batched tape");
  Value sseed = seed->sym_reg >= 0 ? Value(seed) : sym->load(seed);
  // which ids are variables must be asked before a consuming sweep clears the nodes
  std::vector<Id> var_ids;
  for (uint32_t i = 1; i <= g.num_nodes(); i++) {
    Id vid{id.arena, i};
    if (g.is_variable(vid)) var_ids.push_back(vid);
  }
  std::unique_ptr<Store> store = g.evaluate_gradients(id, sseed.data, once);
  // one kernel for the gradients of all variables
  std::vector<int32_t> regs;
  std::vector<Id> ids;
  for (Id vid : var_ids) {
    auto it = store->values.find(vid);
    if (it != store->values.end() && it->second.data->symbolic()) {
      regs.push_back(it->second.data->sym_reg);
      ids.push_back(vid);
    }
  }
  if (!regs.empty()) {
    std::vector<ArrayRef> cols = sym->program()->run(regs);
    for (size_t i = 0; i < ids.size(); i++) store->values[ids[i]].data = cols[i];
  }
  store->program_instructions = sym->program()->last_instructions;
  store->program_slots = sym->program()->last_slots;
  return store;
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
gadcu_status gadcu_batched_program_stats(const gadcu_store* s, uint64_t* instructions, uint64_t* slots) {
  if (instructions) *instructions = s->program_instructions;
  if (slots) *slots = s->program_slots;
  return GADCU_OK;
}

}  // extern "C"
