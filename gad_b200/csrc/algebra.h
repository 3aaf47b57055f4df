// algebra.h — gad's algebras over device arrays: Check (dims), Eval (kernels), Graph1/GraphN (tape).
// One abstract interface carries every operator of the AfAlgebra bundle
// (gad/src/arrayfire.rs:9-37); VJP bodies are written once against it, so the same code runs the
// Graph1 sweep (gradient algebra = Eval, gad/src/graph.rs:250-256) and records the GraphN sweep
// (gradient algebra = the graph itself, graph.rs:294-300).
#pragma once
#include <algorithm>
#include <climits>
#include <cstdint>
#include <initializer_list>
#include <new>
#include <functional>
#include <map>
#include <queue>

#include "core.h"
#include "kernels.h"

namespace gadcu {

const std::string& last_error();

// Every entry point of the C ABI runs under one process-wide recursive lock: handles may be shared between host
// threads (gad's types are Send + Sync, and evaluate_gradients(&self) may be called from several threads on one tape,
// lib.rs:69-72), and a context issues all its work on one stream with shared scratch buffers, so calls are serialised
// rather than made lock-free. Recursive because a user-defined VJP (gadcu_make_node) calls back into the API.
std::recursive_mutex& api_mutex();

template <class F>
gadcu_status guard(F f) {
  std::lock_guard<std::recursive_mutex> api_lock(api_mutex());
  try {
    f();
    return GADCU_OK;
  } catch (const GadError& e) {
    set_last_error(e.msg);
    return e.status;
  } catch (const std::exception& e) {
    set_last_error(e.what());
    return GADCU_ERR_INVALID;
  }
}

// store.rs:12-16 — ordered by (arena, index); index 0 encodes `None`.
struct Id {
  uint32_t arena = 0, index = 0;
  bool some() const { return index != 0; }
  bool operator<(const Id& o) const { return arena != o.arena ? arena < o.arena : index < o.index; }
  bool operator==(const Id& o) const { return arena == o.arena && index == o.index; }
  Id next_id() const { return Id{arena, index + 1}; }
};

// graph.rs:35-43
struct Value {
  ArrayRef data;
  Id id;
  Value() = default;
  Value(ArrayRef d, Id i = Id{}) : data(std::move(d)), id(i) {}
};

// ---- Check: dimension rules. Throw GadError with gad's error variant. -----------
namespace check {
Dim4 equal(const char* name, std::initializer_list<const Dim4*> dims);
Dim4 equal(const char* name, const std::vector<const Dim4*>& dims);
Dim4 select_argmax(const Dim4& v0, const Dim4& v1, const Dim4* r0, const Dim4* r1);
Dim4 flat(const Dim4& v);
Dim4 moddims(const Dim4& v, const Dim4& d);
Dim4 tile_as(const Dim4& v, const Dim4& r);
Dim4 sum_as(const Dim4& v, const Dim4& r);
void as_scalar(const Dim4& v);
void dot(const Dim4& a, const Dim4& b);
Dim4 reduced(const char* name, const Dim4& v, const Dim4& r);
Dim4 transpose(const Dim4& v);
Dim4 matmul(const Dim4& a, const Dim4& b, MatProp p1, MatProp p2);
}  // namespace check

}  // namespace gadcu

struct gadcu_store;

// The algebra interface (the C handle type is this class).
struct gadcu_algebra {
  gadcu_ctx* ctx = nullptr;
  explicit gadcu_algebra(gadcu_ctx* c) : ctx(c) {}
  virtual ~gadcu_algebra() = default;
  using Value = gadcu::Value;
  using Dim4 = gadcu::Dim4;
  using MatProp = gadcu::MatProp;

  virtual gadcu_algebra_kind kind() const = 0;
  virtual gadcu_algebra* clone() const = 0;
  virtual size_t num_nodes() const { return 0; }
  virtual bool is_eval() const { return false; }
  // LinkedAlgebra::link (linked.rs): data view for eval algebras, identity for graphs.
  virtual Value link(const Value& v) { return v; }

  // CoreAlgebra
  virtual Value variable(gadcu::ArrayRef data) = 0;
  virtual Value constant(gadcu::ArrayRef data) = 0;
  virtual Value add(const Value& a, const Value& b) = 0;
  virtual Value add_all(const std::vector<const Value*>& vs);  // core.rs:27-37 left fold
  // ArithAlgebra
  virtual Value sub(const Value& a, const Value& b) = 0;
  virtual Value mul(const Value& a, const Value& b) = 0;
  virtual Value zeros(const Value& v) = 0;
  virtual Value ones(const Value& v) = 0;
  virtual Value neg(const Value& v) = 0;
  // ConstArithAlgebra
  virtual Value setc(const Value& v, double c, bool is_int) = 0;
  virtual Value addc(const Value& v, double c, bool is_int) = 0;
  virtual Value mulc(const Value& v, double c, bool is_int) = 0;
  virtual Value powc(const Value& v, double c, bool is_int) = 0;
  // AnalyticAlgebra
  virtual Value exp(const Value& v) = 0;
  virtual Value log(const Value& v) = 0;
  virtual Value log1p(const Value& v) = 0;
  virtual Value sin(const Value& v) = 0;
  virtual Value cos(const Value& v) = 0;
  virtual Value tanh(const Value& v) = 0;
  virtual Value sigmoid(const Value& v) = 0;
  virtual Value reciprocal(const Value& v) = 0;
  virtual Value sqrt(const Value& v) = 0;
  virtual Value div(const Value& a, const Value& b) = 0;
  virtual Value pow(const Value& v, const Value& p);  // analytic.rs:48-55 default: exp(p * log v)
  // CompareAlgebra (defaults: compare.rs:17-55)
  virtual Value select_argmax(const Value& v0, const Value& v1, const Value* r0, const Value* r1) = 0;
  virtual Value min(const Value& a, const Value& b) { return select_argmax(a, b, &b, &a); }
  virtual Value max(const Value& a, const Value& b) { return select_argmax(a, b, &a, &b); }
  virtual Value abs(const Value& v);
  virtual Value sign(const Value& v);
  virtual Value relu(const Value& v);
  // ArrayAlgebra
  virtual Value flat(const Value& v) = 0;
  virtual Value moddims(const Value& v, const Dim4& d) = 0;
  virtual Value tile_as(const Value& v, const Dim4& d) = 0;
  virtual Value sum_as(const Value& v, const Dim4& d) = 0;
  virtual Value constant_as(const Value& scalar, const Dim4& d) = 0;
  virtual Value as_scalar(const Value& v) = 0;
  virtual Value scale(const Value& lambda, const Value& v) = 0;
  virtual Value dot(const Value& a, const Value& b) = 0;
  virtual Value norm2(const Value& v) { return dot(v, v); }  // array.rs:42-44
  // MatrixAlgebra
  virtual Value matmul(const Value& a, const Value& b, MatProp p1, MatProp p2) = 0;
  virtual Value matmul_nn(const Value& a, const Value& b) { return matmul(a, b, MatProp{}, MatProp{}); }
  virtual Value transpose(const Value& v, bool conjugate) = 0;
  // ArrayCompareAlgebra
  virtual Value max_as(const Value& v, const Dim4& d) = 0;
  virtual Value argmax_as(const Value& v, const Dim4& d) = 0;
  virtual Value softmax_as(const Value& v, const Dim4& d) = 0;

  // GradientStore::add_gradient's `current + new` when `current` may be consumed (store.rs:52).
  virtual Value accumulate(Value&& current, const Value& value) { return add(current, value); }
};

namespace gadcu {

using Algebra = ::gadcu_algebra;

// ---- Eval over device arrays: dims check on the host, then one kernel ---------------------------
class EvalAlgebra final : public Algebra {
 public:
  explicit EvalAlgebra(gadcu_ctx* c) : Algebra(c) {}
  bool fuse_tails = true;  // deferred products may run with their elementwise tail in the epilogue (off under a GraphN: batched.cu lazy_matmul)
  gadcu_algebra_kind kind() const override { return GADCU_EVAL; }
  gadcu_algebra* clone() const override { return new EvalAlgebra(ctx); }
  bool is_eval() const override { return true; }
  Value link(const Value& v) override { return Value(v.data); }

  Value variable(ArrayRef data) override { return Value(std::move(data)); }
  Value constant(ArrayRef data) override { return Value(std::move(data)); }
  Value add(const Value& a, const Value& b) override;
  Value sub(const Value& a, const Value& b) override;
  Value mul(const Value& a, const Value& b) override;
  Value zeros(const Value& v) override;
  Value ones(const Value& v) override;
  Value neg(const Value& v) override;
  Value setc(const Value& v, double c, bool is_int) override;
  Value addc(const Value& v, double c, bool is_int) override;
  Value mulc(const Value& v, double c, bool is_int) override;
  Value powc(const Value& v, double c, bool is_int) override;
  Value exp(const Value& v) override;
  Value log(const Value& v) override;
  Value log1p(const Value& v) override;
  Value sin(const Value& v) override;
  Value cos(const Value& v) override;
  Value tanh(const Value& v) override;
  Value sigmoid(const Value& v) override;
  Value reciprocal(const Value& v) override;
  Value sqrt(const Value& v) override;
  Value div(const Value& a, const Value& b) override;
  Value pow(const Value& v, const Value& p) override;  // af::pow (analytic.rs:124-127)
  Value select_argmax(const Value& v0, const Value& v1, const Value* r0, const Value* r1) override;
  Value min(const Value& a, const Value& b) override;  // af::minof (compare.rs:82-86)
  Value max(const Value& a, const Value& b) override;  // af::maxof (compare.rs:88-92)
  Value flat(const Value& v) override;
  Value moddims(const Value& v, const Dim4& d) override;
  Value tile_as(const Value& v, const Dim4& d) override;
  Value sum_as(const Value& v, const Dim4& d) override;
  Value constant_as(const Value& scalar, const Dim4& d) override;
  Value as_scalar(const Value& v) override;
  Value scale(const Value& lambda, const Value& v) override;
  Value dot(const Value& a, const Value& b) override;
  Value matmul(const Value& a, const Value& b, MatProp p1, MatProp p2) override;
  Value transpose(const Value& v, bool conjugate) override;
  Value max_as(const Value& v, const Dim4& d) override;
  Value argmax_as(const Value& v, const Dim4& d) override;
  Value softmax_as(const Value& v, const Dim4& d) override;
  Value accumulate(Value&& current, const Value& value) override;

  // One kernel: out = [acc +] op(ins...). In place when `acc` is uniquely owned.
  ArrayRef fused(k::EwOp op, ArrayRef acc, std::initializer_list<const ArrayRef*> ins, double c = 0.0, double c2 = 0.0);
  // One GEMM: out = [acc +] op(a)*op(b)
  // `defer`: the product may be recorded as a leaf of a deferred region instead of launched (never for a variable's
  // gradient, which the caller wants computed where the sweep reaches it)
  ArrayRef matmul_acc(ArrayRef acc, const ArrayRef& a, const ArrayRef& b, MatProp p1, MatProp p2, bool defer = false);
  // sum_as with the accumulation folded into the last reduction pass
  ArrayRef sum_as_acc(ArrayRef acc, const ArrayRef& v, const Dim4& d);

 private:
  Value unary(k::EwOp op, const Value& v, double c = 0.0);
  Value binary(k::EwOp op, const char* name, const Value& a, const Value& b);
  ArrayRef reduce_as(k::RedOp op, const ArrayRef& v, const Dim4& rd, ArrayRef acc);
};

// ---- the tape ------------------------------------------------------------------------------------
enum class Op : uint8_t {
  VARIABLE, ADD, ADD_ALL, NEG, SUB, MUL, ADDC, MULC, POWC, EXP, LOG, LOG1P, SIN, COS, TANH, SIGMOID, RECIPROCAL, SQRT, DIV,
  SELECT_ARGMAX, MODDIMS, TILE_AS, SUM_AS, CONSTANT_AS, AS_SCALAR, SCALE, DOT, MATMUL, TRANSPOSE, MAX_AS, SOFTMAX_AS, CUSTOM
};

struct CustomVjp {
  gadcu_vjp_fn fn = nullptr;
  void* user = nullptr;
  void (*drop)(void*) = nullptr;
  ~CustomVjp() { if (drop && user) drop(user); }
};

// A vector whose first N elements live inside the object: a node has one or two inputs and captures at most two values
// in all but a few ops, and a tape is recorded again every step (net_ext.rs:52) — two heap blocks per node were a third
// of the cost of recording one.
template <class T, unsigned N>
class SmallVec {
 public:
  SmallVec() = default;
  SmallVec(std::initializer_list<T> l) { assign(l.begin(), l.end()); }
  SmallVec(const SmallVec& o) { assign(o.begin(), o.end()); }
  SmallVec(SmallVec&& o) noexcept { take(std::move(o)); }
  SmallVec(const std::vector<T>& v) { assign(v.begin(), v.end()); }
  ~SmallVec() { destroy(); }
  SmallVec& operator=(const SmallVec& o) { if (this != &o) { clear(); assign(o.begin(), o.end()); } return *this; }
  SmallVec& operator=(SmallVec&& o) noexcept { if (this != &o) { destroy(); take(std::move(o)); } return *this; }
  SmallVec& operator=(std::initializer_list<T> l) { clear(); assign(l.begin(), l.end()); return *this; }
  SmallVec& operator=(const std::vector<T>& v) { clear(); assign(v.begin(), v.end()); return *this; }
  void push_back(const T& v) { grow(size_ + 1); new (data() + size_) T(v); size_++; }
  void push_back(T&& v) { grow(size_ + 1); new (data() + size_) T(std::move(v)); size_++; }
  void clear() { for (uint32_t i = 0; i < size_; i++) data()[i].~T(); size_ = 0; }
  size_t size() const { return size_; }
  bool empty() const { return size_ == 0; }
  T* data() { return heap_ ? heap_ : reinterpret_cast<T*>(inline_); }
  const T* data() const { return heap_ ? heap_ : reinterpret_cast<const T*>(inline_); }
  T& operator[](size_t i) { return data()[i]; }
  const T& operator[](size_t i) const { return data()[i]; }
  T* begin() { return data(); }
  T* end() { return data() + size_; }
  const T* begin() const { return data(); }
  const T* end() const { return data() + size_; }

 private:
  template <class It> void assign(It first, It last) { for (; first != last; ++first) push_back(*first); }
  void grow(uint32_t want) {
    if (want <= cap_) return;
    uint32_t cap = cap_ * 2 > want ? cap_ * 2 : want;
    T* fresh = static_cast<T*>(::operator new(sizeof(T) * cap));
    for (uint32_t i = 0; i < size_; i++) { new (fresh + i) T(std::move(data()[i])); data()[i].~T(); }
    if (heap_) ::operator delete(heap_);
    heap_ = fresh;
    cap_ = cap;
  }
  void destroy() { clear(); if (heap_) { ::operator delete(heap_); heap_ = nullptr; cap_ = N; } }
  void take(SmallVec&& o) {  // *this holds nothing
    if (o.heap_) { heap_ = o.heap_; cap_ = o.cap_; size_ = o.size_; o.heap_ = nullptr; o.cap_ = N; o.size_ = 0; return; }
    for (uint32_t i = 0; i < o.size_; i++) new (reinterpret_cast<T*>(inline_) + i) T(std::move(o.data()[i]));
    size_ = o.size_;
    o.clear();
  }
  T* heap_ = nullptr;
  uint32_t size_ = 0, cap_ = N;
  alignas(T) unsigned char inline_[sizeof(T) * N];
};

// graph.rs:46-62: inputs + update function. The update function is an op code plus the captured
// forward values instead of a boxed closure, so the sweep can choose fused kernels per node.
struct Node {
  Op op = Op::VARIABLE;
  bool has_update = false;
  SmallVec<Id, 2> inputs;    // pushed on the heap after the update (index 0 = None)
  SmallVec<Value, 2> saved;  // captured forward values (clones are O(1) handle copies)
  double c = 0.0;
  bool c_is_int = false;
  Dim4 d;                    // op-specific dims (vdims or rdims)
  MatProp p1, p2;
  bool flag = false;
  // forward dims captured by make_generic_node for the wrapper's check (graph.rs:149,156)
  Dim4 fwd_dims;
  bool fwd_scalar = false;
  gadcu_dtype fwd_dtype = GADCU_F32;
  std::shared_ptr<CustomVjp> custom;
  void clear() { inputs.clear(); saved.clear(); custom.reset(); has_update = false; }
};

}  // namespace gadcu

// store.rs:60-140 — one map serves Graph1 (values carry no id) and GraphN (values carry ids).
struct gadcu_store {
  mutable std::map<gadcu::Id, gadcu::Value> values;
  // Entries a replayed batched sweep (batched.cu: sweep memo) has not turned into handles yet: (node index, register of the
  // lane program), ascending, register -1 once the entry has moved into `values`. A 569-node tape leaves 569 entries of which
  // its caller reads the 64 variables'; the handle of an entry is made when somebody first looks at it.
  mutable std::vector<std::pair<uint32_t, int32_t>> deferred;
  uint32_t deferred_arena = 0;
  gadcu::ArrayRef (*deferred_handle)(const std::shared_ptr<void>& lane_program, int32_t reg) = nullptr;
  int32_t* deferred_slot(gadcu::Id id) const {  // the live deferred entry of `id`, or null
    if (deferred.empty() || id.arena != deferred_arena) return nullptr;
    auto it = std::lower_bound(deferred.begin(), deferred.end(), std::make_pair(id.index, INT32_MIN));
    return it != deferred.end() && it->first == id.index && it->second >= 0 ? &it->second : nullptr;
  }
  void realize(gadcu::Id id) const {
    if (int32_t* reg = deferred_slot(id)) {
      values.emplace(id, gadcu::Value(deferred_handle(lane_program, *reg)));
      *reg = -1;
    }
  }
  const gadcu::Value* get(gadcu::Id id) const {
    realize(id);
    auto it = values.find(id);
    return it == values.end() ? nullptr : &it->second;
  }
  std::vector<gadcu::Id> ids() const {  // ascending
    std::vector<gadcu::Id> out;
    out.reserve(values.size() + deferred.size());
    for (auto& kv : values) out.push_back(kv.first);
    const size_t mid = out.size();
    for (auto& d : deferred)
      if (d.second >= 0) out.push_back(gadcu::Id{deferred_arena, d.first});
    std::inplace_merge(out.begin(), out.begin() + long(mid), out.end());
    return out;
  }
  void insert(gadcu::Id id, gadcu::Value v) { realize(id); values[id] = std::move(v); }
  // store.rs:44-55
  void add_gradient(gadcu::Algebra& graph, gadcu::Id id, const gadcu::Value& value);
  // fused: contribution = op(ins...) computed and accumulated by one kernel (Graph1 + fusion only)
  void add_fused(gadcu::EvalAlgebra& ev, gadcu::Id id, gadcu::k::EwOp op, std::initializer_list<const gadcu::ArrayRef*> ins,
                 double c = 0.0, double c2 = 0.0);
  // batched-tape stats, and what gadcu_batched_rerun needs (batched.cu)
  uint64_t program_instructions = 0, program_slots = 0;
  std::shared_ptr<void> lane_program;
  std::vector<int32_t> lane_gradient_regs;
};

namespace gadcu {

using Store = ::gadcu_store;

class GraphAlgebra final : public Algebra {
 public:
  // `eval` computes forward values (graph.rs:19-22 `eval: C::EvalAlgebra`): EvalAlgebra for device
  // arrays, or the symbolic per-lane algebra of batched.cu — Graph<Config<E>> for any E.
  GraphAlgebra(gadcu_ctx* c, bool higher_order, std::shared_ptr<Algebra> eval = nullptr);
  gadcu_algebra_kind kind() const override { return kind_override_ ? kind_override_ : (higher_order_ ? GADCU_GRAPHN : GADCU_GRAPH1); }
  gadcu_algebra* clone() const override { return new GraphAlgebra(*this); }  // keeps arena id
  size_t num_nodes() const override { return nodes_.size(); }
  Algebra& eval() { return *eval_; }
  void set_kind(gadcu_algebra_kind k) { kind_override_ = k; }
  const std::vector<ArrayRef>& variable_data() const { return variable_data_; }
  bool is_variable(Id id) const { return id.arena == arena_id_ && id.index >= 1 && id.index <= nodes_.size() && nodes_[id.index - 1].op == Op::VARIABLE && !nodes_[id.index - 1].has_update; }

  Value variable(ArrayRef data) override;
  Value constant(ArrayRef data) override { return Value(std::move(data)); }
  Value add(const Value& a, const Value& b) override;
  Value add_all(const std::vector<const Value*>& vs) override;
  Value sub(const Value& a, const Value& b) override;
  Value mul(const Value& a, const Value& b) override;
  Value zeros(const Value& v) override { return Value(eval_->zeros(v).data); }
  Value ones(const Value& v) override { return Value(eval_->ones(v).data); }
  Value neg(const Value& v) override;
  Value setc(const Value& v, double c, bool is_int) override { return Value(eval_->addc(v, c, is_int).data); }  // sic, const_arith.rs:134-137
  Value addc(const Value& v, double c, bool is_int) override;
  Value mulc(const Value& v, double c, bool is_int) override;
  Value powc(const Value& v, double c, bool is_int) override;
  Value exp(const Value& v) override;
  Value log(const Value& v) override;
  Value log1p(const Value& v) override;
  Value sin(const Value& v) override;
  Value cos(const Value& v) override;
  Value tanh(const Value& v) override;
  Value sigmoid(const Value& v) override;
  Value reciprocal(const Value& v) override;
  Value sqrt(const Value& v) override;
  Value div(const Value& a, const Value& b) override;
  Value select_argmax(const Value& v0, const Value& v1, const Value* r0, const Value* r1) override;
  Value flat(const Value& v) override;
  Value moddims(const Value& v, const Dim4& d) override;
  Value tile_as(const Value& v, const Dim4& d) override;
  Value sum_as(const Value& v, const Dim4& d) override;
  Value constant_as(const Value& scalar, const Dim4& d) override;
  Value as_scalar(const Value& v) override;
  Value scale(const Value& lambda, const Value& v) override;
  Value dot(const Value& a, const Value& b) override;
  Value matmul(const Value& a, const Value& b, MatProp p1, MatProp p2) override;
  Value transpose(const Value& v, bool conjugate) override;
  Value max_as(const Value& v, const Dim4& d) override;
  Value argmax_as(const Value& v, const Dim4& d) override { return Value(eval_->argmax_as(v, d).data); }
  Value softmax_as(const Value& v, const Dim4& d) override;

  // graph.rs:106-165
  Value make_node(ArrayRef data, Node node);
  Value make_custom(ArrayRef data, std::vector<Id> inputs, std::shared_ptr<CustomVjp> vjp);

  // graph.rs:263-290 / 307-318
  // Called during a Graph1 sweep with each VARIABLE's gradient as soon as no unprocessed node can contribute to it
  // any more (and, at the end, for the variables the sweep never completed): the data-parallel sweep starts the
  // all-reduce of a layer's weight gradient while the layers below are still being differentiated.
  // The hook may return a replacement array for the store entry (a private copy it will write into, when the gradient's
  // buffer is visible outside the store — the caller's seed, an array stored under several ids); null keeps the entry.
  using FinalHook = std::function<ArrayRef(Id, const ArrayRef&)>;
  void set_final_hook(FinalHook h) { final_hook_ = std::move(h); }
  // Asked by the matmul VJP when the product it is about to compute is the ONLY contribution to a variable's
  // gradient: may compute it itself (the data-parallel GEMM whose epilogue reduce-scatters over NVLink) and return
  // the array that will hold the rank-summed gradient, or return null to decline. The final hook is not called for
  // a variable served this way.
  using GemmSink = std::function<ArrayRef(Id, const ArrayRef& lhs, const ArrayRef& rhs, MatProp pl, MatProp pr, bool last)>;  // last: no other large variable is still waiting for its gradient
  void set_gemm_sink(GemmSink h) { gemm_sink_ = std::move(h); }
  std::unique_ptr<Store> evaluate_gradients(Id id, ArrayRef seed, bool once);
  std::unique_ptr<Store> compute_gradients(Id id, const Value& seed);
  // batched.cu, sweep memo: everything a reverse sweep reads from the tape, folded into two words (false: the tape holds
  // something the fold cannot describe — a user-defined VJP, a captured array `register_of` has no register for); and the effect of
  // a consuming sweep on the nodes it visited (graph.rs:242), for a sweep that was replayed rather than run
  bool structure_hash(uint64_t out[2], const std::function<int64_t(const gadcu_array&)>& register_of) const;
  void consume(const std::vector<uint32_t>& node_indices);

 private:
  FinalHook final_hook_;
  GemmSink gemm_sink_;
  std::vector<uint32_t> waiting_;  // during a hooked sweep: nodes still to run before a variable's gradient is complete
  std::vector<char> served_;       // variables whose gradient a GemmSink produced
  std::unique_ptr<Store> sweep(Algebra& ga, Id gid, Value gradient, bool once);
  void vjp(const Node& n, Algebra& ga, Store& store, const Value& g);
  ArrayRef try_gemm_sink(Id id, const ArrayRef& lhs, const ArrayRef& rhs, MatProp pl, MatProp pr);

  bool higher_order_;
  gadcu_algebra_kind kind_override_ = gadcu_algebra_kind(0);
  uint32_t arena_id_;
  std::shared_ptr<Algebra> eval_;
  std::vector<Node> nodes_;
  std::vector<ArrayRef> variable_data_;  // batched tapes: the column of every variable, in creation order
};

}  // namespace gadcu
