// gadcu.hpp — header-only C++17 host mirror of gad's algebra interface above the C ABI of include/gadcu.h.
//
// The reference's host language is Rust (no toolchain in this image; the Rust binding is rust/src/lib.rs, source
// only). This header is the compiled-language mirror that IS built and run here: the same names, argument order and
// error behaviour as the reference's traits, so a test written against it reads like the reference's own tests
// (examples/reference_tests.cc replays gad/tests/{graph,extension}.rs and parts of the array tests through it).
//
//   reference (gad/src)                                   here
//   --------------------------------------------------   -----------------------------------------------------
//   af::Array<T>, HasDims::dims   core.rs:41-73,120-136   gadcu::Array<T>::{from_host, host, dims}; copy = O(1) gadcu_array_clone
//   Value<D> {data, id}, gid()    graph.rs:35-43,321-342  gadcu::Value<T>::{data, id, gid}
//   Result<_> / Error variants    error.rs:9-40           gadcu::Error (thrown) with .kind in the reference's order
//   Check                         lib.rs:519-520          gadcu::Check (static functions on Dim4)
//   Eval, Graph1, GraphN          lib.rs:523-532          gadcu::Eval, gadcu::Graph1, gadcu::GraphN (: gadcu::Algebra)
//   CoreAlgebra .. MatrixAlgebra  core.rs:12-38 etc.      member functions of gadcu::Algebra, one per trait method
//   Graph::make_node, add_gradient graph.rs:106-165       Algebra::make_node(data, inputs, vjp), StoreRef::add_gradient
//   evaluate_gradients[_once]     graph.rs:263-290        Graph1::evaluate_gradients[_once](id, seed)
//   compute_gradients             graph.rs:307-318        GraphN::compute_gradients(id, seed)
//   GradientReader::read          store.rs:27-29          Store<T>::get(id) -> std::optional (absent is not zero)
//   WeightOps                     net.rs:161-180          Array<T>::{add_assign, scale}
//   Graph<Config1<_>> over per-lane scalars               gadcu::BatchedGraph1<T>: one scalar tape per lane (tests/symbolic.rs:93 is the
//     (the example loop of net_ext.rs:50-66)                precedent for deriving a tape from another algebra), Store::rerun
//   (none: the reference is one process)                  gadcu::Comm + Graph1::evaluate_gradients_dp / GraphN::compute_gradients_dp:
//                                                         the batch split over the GPUs of a box, gradients summed over the ranks
//
// `Err(Error::X)` in the reference is `throw gadcu::Error{X}` here; nothing else changes meaning. There is no CPU
// fallback: constructing a Context without a B200 throws Error{Cuda}.
#pragma once
#include <cstdint>
#include <functional>
#include <initializer_list>
#include <memory>
#include <optional>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

#include "gadcu.h"

namespace gadcu {

// error.rs:9-40 (1..7, same order) + the backend's own codes
enum class ErrorKind {
  Dimensions = 1, ReducedDimensions, Lengths, Empty, MissingId, MissingGradient, MissingNode, Cuda, Nccl, Unsupported, Invalid, Type
};
class Error : public std::runtime_error {
 public:
  Error(ErrorKind k, const std::string& what) : std::runtime_error(what), kind(k) {}
  ErrorKind kind;
};
inline void check(gadcu_status st) {
  if (st != GADCU_OK) throw Error(static_cast<ErrorKind>(st), gadcu_last_error());
}

// af::Dim4: four extents, dim 0 fastest, trailing ones
struct Dim4 {
  uint64_t d[4] = {1, 1, 1, 1};
  Dim4() = default;
  Dim4(std::initializer_list<uint64_t> l) {
    size_t i = 0;
    for (uint64_t v : l)
      if (i < 4) d[i++] = v;
  }
  explicit Dim4(const gadcu_dim4& c) { for (int i = 0; i < 4; i++) d[i] = c.d[i]; }
  gadcu_dim4 c() const { return gadcu_dim4{{d[0], d[1], d[2], d[3]}}; }
  uint64_t elements() const { return d[0] * d[1] * d[2] * d[3]; }
  uint64_t operator[](int i) const { return d[i]; }
  bool operator==(const Dim4& o) const { return d[0] == o.d[0] && d[1] == o.d[1] && d[2] == o.d[2] && d[3] == o.d[3]; }
  bool operator!=(const Dim4& o) const { return !(*this == o); }
};

// matrix.rs:13-17
struct MatProp {
  bool transposed = false, conjugated = false;
  gadcu_matprop c() const { return gadcu_matprop{uint8_t(transposed), uint8_t(conjugated)}; }
};

template <class T> struct dtype_of;
template <> struct dtype_of<float> { static constexpr gadcu_dtype value = GADCU_F32; };
template <> struct dtype_of<double> { static constexpr gadcu_dtype value = GADCU_F64; };

// device, stream, allocator
class Context {
 public:
  explicit Context(int device = 0) { check(gadcu_ctx_create(device, &h_)); }
  ~Context() { if (h_) gadcu_ctx_destroy(h_); }
  Context(const Context&) = delete;
  Context& operator=(const Context&) = delete;
  gadcu_ctx* handle() const { return h_; }
  void synchronize() { check(gadcu_ctx_synchronize(h_)); }
  void set_fusion(bool fuse) { check(gadcu_ctx_set_fusion(h_, fuse)); }
  void set_lazy(bool enable, uint64_t min_elements = 32768) { check(gadcu_ctx_set_lazy(h_, enable, min_elements)); }
  void set_gemm_epilogue(bool enable) { check(gadcu_ctx_set_gemm_epilogue(h_, enable)); }
  void set_auto_replay(bool enable) { check(gadcu_ctx_set_auto_replay(h_, enable)); }  // graphs keyed by the recorded kernel sequence
  void flush() { check(gadcu_ctx_flush(h_)); }
  void wait_download(uint64_t ticket) { check(gadcu_ctx_wait_download(h_, ticket)); }  // see Array::download_async
  uint64_t launches() {
    uint64_t a, b, c;
    check(gadcu_ctx_memory_stats(h_, &a, &b, &c));
    return c;
  }

 private:
  gadcu_ctx* h_ = nullptr;
};

// af::Array<T>: immutable, refcounted, column-major; copying is O(1) (gad clones values into every VJP closure)
template <class T>
class Array {
 public:
  Array() = default;
  static Array adopt(gadcu_array* h) { Array a; a.h_ = h; return a; }  // takes over one reference
  static Array from_host(Context& ctx, const std::vector<T>& values, Dim4 dims) {  // af::Array::new (column-major)
    if (values.size() != dims.elements()) throw Error(ErrorKind::Lengths, "Array::from_host: values vs dims");
    gadcu_array* h = nullptr;
    const gadcu_dim4 d = dims.c();
    check(gadcu_array_from_host(ctx.handle(), dtype_of<T>::value, &d, values.data(), &h));
    return adopt(h);
  }
  static Array scalar(Context& ctx, T v) {
    gadcu_array* h = nullptr;
    check(gadcu_scalar_from_host(ctx.handle(), dtype_of<T>::value, double(v), &h));
    return adopt(h);
  }
  static Array random(Context& ctx, Dim4 dims, uint64_t seed, bool normal = false, double a = 0.0, double b = 1.0) {
    gadcu_array* h = nullptr;
    const gadcu_dim4 d = dims.c();
    check(gadcu_array_random(ctx.handle(), dtype_of<T>::value, &d, seed, normal, a, b, &h));
    return adopt(h);
  }
  // af::Array::clone: a new handle on the same buffer — `+=` on one copy leaves the others alone (copy-on-write)
  Array(const Array& o) { if (o.h_) check(gadcu_array_clone(o.h_, &h_)); }
  Array(Array&& o) noexcept : h_(o.h_) { o.h_ = nullptr; }
  Array& operator=(Array o) noexcept { std::swap(h_, o.h_); return *this; }
  ~Array() { if (h_) gadcu_array_release(h_); }

  gadcu_array* handle() const { return h_; }
  explicit operator bool() const { return h_ != nullptr; }
  Dim4 dims() const {  // HasDims::dims
    gadcu_dim4 d;
    check(gadcu_array_dims(h_, &d));
    return Dim4(d);
  }
  std::vector<T> host() const {  // af::Array::host (synchronises)
    std::vector<T> out(dims().elements());
    check(gadcu_array_to_host(h_, out.data()));
    return out;
  }
  // Queue the device -> host copy of this array into caller-owned (pinned) memory of dims().elements() values behind the
  // work that produces it; returns a ticket for Context::wait_download. The host is not held up (host() / item() are).
  uint64_t download_async(T* pinned) const {
    uint64_t ticket = 0;
    check(gadcu_array_download_async(h_, pinned, &ticket));
    return ticket;
  }
  T item() const {  // a one-element array or a device scalar (`dot`, `as_scalar`: array.rs:106-126); synchronises
    double v = 0;
    check(gadcu_scalar_to_host(h_, &v));
    return T(v);
  }
  // WeightOps (net.rs:161-180)
  void add_assign(const Array& other) { check(gadcu_array_add_assign(h_, other.h_)); }
  Array scale(T lambda) const {
    gadcu_array* out = nullptr;
    check(gadcu_array_scale(h_, double(lambda), &out));
    return adopt(out);
  }

 private:
  gadcu_array* h_ = nullptr;
};

// store.rs:12-23: (arena, 1-based index), ordered lexicographically
struct Id {
  uint32_t arena = 0, index = 0;
  bool operator==(const Id& o) const { return arena == o.arena && index == o.index; }
  bool operator<(const Id& o) const { return arena != o.arena ? arena < o.arena : index < o.index; }
};

// graph.rs:35-43: forward data + optional id (constants have none)
template <class T>
class Value {
 public:
  Value() = default;
  explicit Value(gadcu_value v) : data_(Array<T>::adopt(v.data)), id_{v.arena, v.index} {}
  Value(Array<T> data, Id id) : data_(std::move(data)), id_(id) {}
  const Array<T>& data() const { return data_; }                                             // graph.rs:329-331
  std::optional<Id> id() const { return id_.index ? std::optional<Id>(id_) : std::nullopt; }  // graph.rs:334-336
  Id gid() const {                                                                           // net.rs HasGradientId for Value<D>
    if (!id_.index) throw Error(ErrorKind::MissingId, "Trying to obtain `id` of a constant value.");
    return id_;
  }
  gadcu_value c() const { return gadcu_value{data_.handle(), id_.arena, id_.index}; }  // borrowed view for a call

 private:
  Array<T> data_;
  Id id_;
};

// Check (lib.rs:519-520): the dimension algebra, pure host functions
struct Check {
  static Dim4 equal(const char*, std::initializer_list<Dim4> dims) {  // add sub mul div pow min max add_all: error.rs:139-153
    std::vector<gadcu_dim4> c;
    for (const Dim4& d : dims) c.push_back(d.c());
    std::vector<const gadcu_dim4*> p;
    for (auto& d : c) p.push_back(&d);
    gadcu_dim4 out;
    check(gadcu_check_equal(p.data(), p.size(), &out));
    return Dim4(out);
  }
  static Dim4 add(Dim4 a, Dim4 b) { return equal("add", {a, b}); }
  static Dim4 sub(Dim4 a, Dim4 b) { return equal("sub", {a, b}); }
  static Dim4 mul(Dim4 a, Dim4 b) { return equal("mul", {a, b}); }
  static Dim4 div(Dim4 a, Dim4 b) { return equal("div", {a, b}); }
  static Dim4 pow(Dim4 a, Dim4 b) { return equal("pow", {a, b}); }
  static Dim4 select_argmax(Dim4 v0, Dim4 v1, const Dim4* r0, const Dim4* r1) {  // compare.rs:128-144
    const gadcu_dim4 a = v0.c(), b = v1.c();
    gadcu_dim4 c0, c1, out;
    if (r0) c0 = r0->c();
    if (r1) c1 = r1->c();
    check(gadcu_check_select_argmax(&a, &b, r0 ? &c0 : nullptr, r1 ? &c1 : nullptr, &out));
    return Dim4(out);
  }
#define GADCU_CHECK2(name, fn)                         \
  static Dim4 name(Dim4 v, Dim4 r) {                   \
    const gadcu_dim4 a = v.c(), b = r.c();             \
    gadcu_dim4 out;                                    \
    check(fn(&a, &b, &out));                           \
    return Dim4(out);                                  \
  }
  GADCU_CHECK2(moddims, gadcu_check_moddims)  // array.rs:138-145
  GADCU_CHECK2(tile_as, gadcu_check_tile_as)  // array.rs:147-157
  GADCU_CHECK2(sum_as, gadcu_check_sum_as)    // array.rs:159-170
#undef GADCU_CHECK2
  static Dim4 flat(Dim4 v) {
    const gadcu_dim4 a = v.c();
    gadcu_dim4 out;
    check(gadcu_check_flat(&a, &out));
    return Dim4(out);
  }
  static void as_scalar(Dim4 v) { const gadcu_dim4 a = v.c(); check(gadcu_check_as_scalar(&a)); }
  static void dot(Dim4 x, Dim4 y) { const gadcu_dim4 a = x.c(), b = y.c(); check(gadcu_check_dot(&a, &b)); }
  static Dim4 reduced(Dim4 v, Dim4 r, bool keeps_shape) {  // array_compare.rs:84-101
    const gadcu_dim4 a = v.c(), b = r.c();
    gadcu_dim4 out;
    check(gadcu_check_reduced(&a, &b, keeps_shape, &out));
    return Dim4(out);
  }
  static Dim4 max_as(Dim4 v, Dim4 r) { return reduced(v, r, false); }
  static Dim4 argmax_as(Dim4 v, Dim4 r) { return reduced(v, r, true); }
  static Dim4 softmax_as(Dim4 v, Dim4 r) { return reduced(v, r, true); }
  static Dim4 transpose(Dim4 v) {  // matrix.rs:118-125
    const gadcu_dim4 a = v.c();
    gadcu_dim4 out;
    check(gadcu_check_transpose(&a, &out));
    return Dim4(out);
  }
  static Dim4 matmul(Dim4 x, Dim4 y, MatProp p1 = {}, MatProp p2 = {}) {  // matrix.rs:88-116
    const gadcu_dim4 a = x.c(), b = y.c();
    gadcu_dim4 out;
    check(gadcu_check_matmul(&a, &b, p1.c(), p2.c(), &out));
    return Dim4(out);
  }
};

class Algebra;

// the store as a VJP callback sees it (borrowed): GradientStore::add_gradient, store.rs:44-55
class StoreRef {
 public:
  StoreRef(gadcu_store* s) : s_(s) {}
  template <class T> void add_gradient(Algebra& graph, Id id, const Value<T>& value);

 private:
  gadcu_store* s_;
};

// The gradients of a sweep. get() is GradientReader::read (store.rs:27-29): an id the sweep never reached is absent.
template <class T>
class Store {
 public:
  explicit Store(gadcu_store* s) : s_(s, [](gadcu_store* p) { gadcu_store_destroy(p); }) {}
  std::optional<Value<T>> get(Id id) const {
    gadcu_value out{nullptr, 0, 0};
    int found = 0;
    check(gadcu_store_get(s_.get(), id.arena, id.index, &out, &found));
    if (!found) return std::nullopt;
    return Value<T>(out);
  }
  gadcu_store* handle() const { return s_.get(); }
  // Batched tapes only (BatchedGraph1): the same tapes at new points — `columns[i]` replaces the data of the i-th variable
  // (creation order) and the compiled lane program runs again, one launch, nothing re-recorded. Returns the variables'
  // gradient columns (an empty Array where the sweep produced none).
  std::vector<Array<T>> rerun(const std::vector<Array<T>>& columns) const {
    std::vector<gadcu_array*> in, out(columns.size(), nullptr);
    for (const Array<T>& c : columns) in.push_back(c.handle());
    check(gadcu_batched_rerun(s_.get(), in.data(), in.size(), out.data()));
    std::vector<Array<T>> gradients;
    for (gadcu_array* h : out) gradients.push_back(Array<T>::adopt(h));
    return gradients;
  }
  std::pair<uint64_t, uint64_t> program_stats() const {  // (instructions, value slots) of the lane program behind a batched sweep
    uint64_t instructions = 0, slots = 0;
    check(gadcu_batched_program_stats(s_.get(), &instructions, &slots));
    return {instructions, slots};
  }

 private:
  std::shared_ptr<gadcu_store> s_;
};

// Data parallelism over the batch dimension (gadcu.h "data parallelism"; new — the reference is single-process): one process
// per GPU, every rank builds a Comm from the same 128-byte id (rank 0 makes it with Comm::unique_id() and hands it to the
// others by whatever channel the host program has). apply_gradient_step sums per-example gradients (net_ext.rs:62-65), so
// sharding the batch and summing the weight gradients over the ranks reproduces it up to the order of the sum.
class Comm {
 public:
  static std::vector<uint8_t> unique_id() {
    std::vector<uint8_t> id(128);
    check(gadcu_comm_unique_id(id.data()));
    return id;
  }
  Comm(Context& ctx, int nranks, int rank, const std::vector<uint8_t>& id) : nranks_(nranks), rank_(rank) {
    if (id.size() != 128) throw Error(ErrorKind::Lengths, "Comm: the unique id is 128 bytes");
    check(gadcu_comm_init(ctx.handle(), nranks, rank, id.data(), &h_));
  }
  ~Comm() { if (h_) gadcu_comm_destroy(h_); }
  Comm(const Comm&) = delete;
  Comm& operator=(const Comm&) = delete;
  gadcu_comm* handle() const { return h_; }
  int nranks() const { return nranks_; }
  int rank() const { return rank_; }
  // in-place sum over the ranks; the async form runs on the communicator's stream behind what the compute stream has
  // been given so far, and join() makes the compute stream wait for it (the host is never held up)
  template <class T> void allreduce_sum(const std::vector<Array<T>*>& arrays) { check(gadcu_comm_allreduce_sum(h_, handles(arrays).data(), arrays.size())); }
  template <class T> void allreduce_sum_async(const std::vector<Array<T>*>& arrays) {
    check(gadcu_comm_allreduce_sum_async(h_, handles(arrays).data(), arrays.size()));
  }
  void join() { check(gadcu_comm_join(h_)); }

 private:
  template <class T> static std::vector<gadcu_array*> handles(const std::vector<Array<T>*>& arrays) {
    std::vector<gadcu_array*> h;
    for (const Array<T>* a : arrays) h.push_back(a->handle());
    return h;
  }
  gadcu_comm* h_ = nullptr;
  int nranks_, rank_;
};

// One of gad's default algebras over device arrays; the SAME member functions serve Eval, Graph1 and GraphN, as one
// generic gad program is instantiated with different algebras (arrayfire.rs:39-61).
class Algebra {
 public:
  Algebra(Context& ctx, gadcu_algebra_kind kind) : ctx_(&ctx), owned_(true) { check(gadcu_algebra_create(ctx.handle(), kind, &h_)); }
  Algebra(Context& ctx, gadcu_algebra* borrowed) : ctx_(&ctx), h_(borrowed), owned_(false) {}  // the gradient algebra inside a VJP
  struct Adopt {};
  Algebra(Context& ctx, gadcu_algebra* owned, Adopt) : ctx_(&ctx), h_(owned), owned_(true) {}  // a handle made by another entry point
  Algebra(const Algebra& o) : ctx_(o.ctx_), owned_(true) { check(gadcu_algebra_clone(o.h_, &h_)); }  // graph.rs:353-360
  Algebra& operator=(const Algebra&) = delete;
  virtual ~Algebra() { if (owned_ && h_) gadcu_algebra_destroy(h_); }
  gadcu_algebra* handle() const { return h_; }
  Context& ctx() const { return *ctx_; }
  size_t num_nodes() const { return gadcu_algebra_num_nodes(h_); }
  gadcu_algebra_kind kind() const { return gadcu_algebra_get_kind(h_); }
  // LinkedAlgebra::link (linked.rs:9-30), for VJP bodies: how the gradient algebra sees a captured forward value —
  // plain data when gradients are data (Graph1: the algebra is Eval), the tracked value itself on a higher-order tape
  template <class T> Value<T> link(const Value<T>& v) { return kind() == GADCU_GRAPHN ? v : constant(v.data()); }

  // CoreAlgebra (core.rs:12-38)
  template <class T> Value<T> variable(const Array<T>& data) { gadcu_value o; check(gadcu_variable(h_, data.handle(), &o)); return Value<T>(o); }
  template <class T> Value<T> constant(const Array<T>& data) { gadcu_value o; check(gadcu_constant(h_, data.handle(), &o)); return Value<T>(o); }
  template <class T> Value<T> add(const Value<T>& a, const Value<T>& b) { return op2(gadcu_add, a, b); }
  template <class T> Value<T> add_all(const std::vector<const Value<T>*>& vs) {
    std::vector<gadcu_value> c;
    for (auto* v : vs) c.push_back(v->c());
    std::vector<const gadcu_value*> p;
    for (auto& v : c) p.push_back(&v);
    gadcu_value o;
    check(gadcu_add_all(h_, p.data(), p.size(), &o));
    return Value<T>(o);
  }
  // ArithAlgebra (arith.rs:14-32)
  template <class T> Value<T> sub(const Value<T>& a, const Value<T>& b) { return op2(gadcu_sub, a, b); }
  template <class T> Value<T> mul(const Value<T>& a, const Value<T>& b) { return op2(gadcu_mul, a, b); }
  template <class T> Value<T> zeros(const Value<T>& v) { return op1(gadcu_zeros, v); }
  template <class T> Value<T> ones(const Value<T>& v) { return op1(gadcu_ones, v); }
  template <class T> Value<T> neg(const Value<T>& v) { return op1(gadcu_neg, v); }
  // ConstArithAlgebra<_, T> and <_, i16> (const_arith.rs:14-26)
  template <class T> Value<T> setc(const Value<T>& v, T c) { return opc(gadcu_setc, v, double(c), 0); }
  template <class T> Value<T> addc(const Value<T>& v, T c) { return opc(gadcu_addc, v, double(c), 0); }
  template <class T> Value<T> mulc(const Value<T>& v, T c) { return opc(gadcu_mulc, v, double(c), 0); }
  template <class T> Value<T> powc(const Value<T>& v, T c) { return opc(gadcu_powc, v, double(c), 0); }
  template <class T> Value<T> setc(const Value<T>& v, int16_t c) { return opc(gadcu_setc, v, double(c), 1); }
  template <class T> Value<T> addc(const Value<T>& v, int16_t c) { return opc(gadcu_addc, v, double(c), 1); }
  template <class T> Value<T> mulc(const Value<T>& v, int16_t c) { return opc(gadcu_mulc, v, double(c), 1); }
  template <class T> Value<T> powc(const Value<T>& v, int16_t c) { return opc(gadcu_powc, v, double(c), 1); }
  // AnalyticAlgebra (analytic.rs:16-56)
  template <class T> Value<T> exp(const Value<T>& v) { return op1(gadcu_exp, v); }
  template <class T> Value<T> log(const Value<T>& v) { return op1(gadcu_log, v); }
  template <class T> Value<T> log1p(const Value<T>& v) { return op1(gadcu_log1p, v); }
  template <class T> Value<T> sin(const Value<T>& v) { return op1(gadcu_sin, v); }
  template <class T> Value<T> cos(const Value<T>& v) { return op1(gadcu_cos, v); }
  template <class T> Value<T> tanh(const Value<T>& v) { return op1(gadcu_tanh, v); }
  template <class T> Value<T> sigmoid(const Value<T>& v) { return op1(gadcu_sigmoid, v); }
  template <class T> Value<T> reciprocal(const Value<T>& v) { return op1(gadcu_reciprocal, v); }
  template <class T> Value<T> sqrt(const Value<T>& v) { return op1(gadcu_sqrt, v); }
  template <class T> Value<T> div(const Value<T>& a, const Value<T>& b) { return op2(gadcu_div, a, b); }
  template <class T> Value<T> pow(const Value<T>& v, const Value<T>& p) { return op2(gadcu_pow, v, p); }
  // CompareAlgebra (compare.rs:15-66): `if v0 >= v1 then r0 else r1`, absent results are zeros
  template <class T> Value<T> select_argmax(const Value<T>& v0, const Value<T>& v1, const Value<T>* r0, const Value<T>* r1) {
    const gadcu_value a = v0.c(), b = v1.c();
    gadcu_value c0, c1, o;
    if (r0) c0 = r0->c();
    if (r1) c1 = r1->c();
    check(gadcu_select_argmax(h_, &a, &b, r0 ? &c0 : nullptr, r1 ? &c1 : nullptr, &o));
    return Value<T>(o);
  }
  template <class T> Value<T> min(const Value<T>& a, const Value<T>& b) { return op2(gadcu_min, a, b); }
  template <class T> Value<T> max(const Value<T>& a, const Value<T>& b) { return op2(gadcu_max, a, b); }
  template <class T> Value<T> abs(const Value<T>& v) { return op1(gadcu_abs, v); }
  template <class T> Value<T> sign(const Value<T>& v) { return op1(gadcu_sign, v); }
  template <class T> Value<T> relu(const Value<T>& v) { return op1(gadcu_relu, v); }
  // ArrayAlgebra (array.rs:13-45); scalars (`as_scalar`, `dot`, `norm2`) are one-element device values: .data().item()
  template <class T> Value<T> flat(const Value<T>& v) { return op1(gadcu_flat, v); }
  template <class T> Value<T> moddims(const Value<T>& v, Dim4 d) { return opd(gadcu_moddims, v, d); }
  template <class T> Value<T> tile_as(const Value<T>& v, Dim4 d) { return opd(gadcu_tile_as, v, d); }
  template <class T> Value<T> sum_as(const Value<T>& v, Dim4 d) { return opd(gadcu_sum_as, v, d); }
  template <class T> Value<T> constant_as(const Value<T>& scalar, Dim4 d) { return opd(gadcu_constant_as, scalar, d); }
  template <class T> Value<T> as_scalar(const Value<T>& v) { return op1(gadcu_as_scalar, v); }
  template <class T> Value<T> scale(const Value<T>& lambda, const Value<T>& v) { return op2(gadcu_scale, lambda, v); }
  template <class T> Value<T> dot(const Value<T>& a, const Value<T>& b) { return op2(gadcu_dot, a, b); }
  template <class T> Value<T> norm2(const Value<T>& v) { return op1(gadcu_norm2, v); }
  // MatrixAlgebra (matrix.rs:20-32)
  template <class T> Value<T> matmul(const Value<T>& a, const Value<T>& b, MatProp p1 = {}, MatProp p2 = {}) {
    const gadcu_value x = a.c(), y = b.c();
    gadcu_value o;
    check(gadcu_matmul(h_, &x, &y, p1.c(), p2.c(), &o));
    return Value<T>(o);
  }
  template <class T> Value<T> matmul_nn(const Value<T>& a, const Value<T>& b) { return op2(gadcu_matmul_nn, a, b); }
  template <class T> Value<T> transpose(const Value<T>& v, bool conjugate = false) {
    const gadcu_value x = v.c();
    gadcu_value o;
    check(gadcu_transpose(h_, &x, conjugate, &o));
    return Value<T>(o);
  }
  // ArrayCompareAlgebra (array_compare.rs:16-22)
  template <class T> Value<T> max_as(const Value<T>& v, Dim4 d) { return opd(gadcu_max_as, v, d); }
  template <class T> Value<T> argmax_as(const Value<T>& v, Dim4 d) { return opd(gadcu_argmax_as, v, d); }
  template <class T> Value<T> softmax_as(const Value<T>& v, Dim4 d) { return opd(gadcu_softmax_as, v, d); }

  // Graph::make_node (graph.rs:106-124): a user-defined operator, as in gad/tests/extension.rs:61-96. `vjp` receives the
  // gradient algebra (Eval for Graph1, the graph itself for GraphN), the store, and the node's own gradient.
  template <class T>
  Value<T> make_node(const Array<T>& data, const std::vector<const Value<T>*>& inputs,
                     std::function<void(Algebra&, StoreRef&, const Value<T>&)> vjp) {
    struct User {
      Context* ctx;
      std::function<void(Algebra&, StoreRef&, const Value<T>&)> fn;
    };
    auto* user = new User{ctx_, std::move(vjp)};
    std::vector<gadcu_value> c;
    for (auto* v : inputs) c.push_back(v->c());
    std::vector<const gadcu_value*> p;
    for (auto& v : c) p.push_back(&v);
    gadcu_value o;
    const gadcu_status st = gadcu_make_node(
        h_, data.handle(), p.data(), p.size(),
        [](void* u, gadcu_algebra* ga, gadcu_store* store, const gadcu_value* gradient) -> gadcu_status {
          auto* self = static_cast<User*>(u);
          try {
            Algebra view(*self->ctx, ga);
            StoreRef sref(store);
            gadcu_array_retain(gradient->data);  // the Value below owns one reference
            Value<T> g(*gradient);
            self->fn(view, sref, g);
            return GADCU_OK;
          } catch (const Error& e) {
            return static_cast<gadcu_status>(e.kind);
          }
        },
        user, [](void* u) { delete static_cast<User*>(u); }, &o);
    check(st);
    return Value<T>(o);
  }

 protected:
  using Fn1 = gadcu_status (*)(gadcu_algebra*, const gadcu_value*, gadcu_value*);
  using Fn2 = gadcu_status (*)(gadcu_algebra*, const gadcu_value*, const gadcu_value*, gadcu_value*);
  using FnD = gadcu_status (*)(gadcu_algebra*, const gadcu_value*, const gadcu_dim4*, gadcu_value*);
  using FnC = gadcu_status (*)(gadcu_algebra*, const gadcu_value*, double, int, gadcu_value*);
  template <class T> Value<T> op1(Fn1 f, const Value<T>& v) { const gadcu_value a = v.c(); gadcu_value o; check(f(h_, &a, &o)); return Value<T>(o); }
  template <class T> Value<T> op2(Fn2 f, const Value<T>& x, const Value<T>& y) {
    const gadcu_value a = x.c(), b = y.c();
    gadcu_value o;
    check(f(h_, &a, &b, &o));
    return Value<T>(o);
  }
  template <class T> Value<T> opd(FnD f, const Value<T>& v, Dim4 d) {
    const gadcu_value a = v.c();
    const gadcu_dim4 c = d.c();
    gadcu_value o;
    check(f(h_, &a, &c, &o));
    return Value<T>(o);
  }
  template <class T> Value<T> opc(FnC f, const Value<T>& v, double c, int is_int) {
    const gadcu_value a = v.c();
    gadcu_value o;
    check(f(h_, &a, c, is_int, &o));
    return Value<T>(o);
  }
  Context* ctx_;
  gadcu_algebra* h_ = nullptr;
  bool owned_;
};

template <class T>
void StoreRef::add_gradient(Algebra& graph, Id id, const Value<T>& value) {
  const gadcu_value v = value.c();
  check(gadcu_store_add_gradient(s_, graph.handle(), id.arena, id.index, &v));
}

// lib.rs:523-524
class Eval : public Algebra {
 public:
  explicit Eval(Context& ctx) : Algebra(ctx, GADCU_EVAL) {}
};

// lib.rs:527-528: first-order tape, gradients are data
class Graph1 : public Algebra {
 public:
  explicit Graph1(Context& ctx) : Algebra(ctx, GADCU_GRAPH1) {}
  Graph1(const Graph1&) = default;  // clone(): graph.rs:353-360
  Graph1(Context& ctx, gadcu_algebra* owned, Adopt a) : Algebra(ctx, owned, a) {}
  template <class T> Store<T> evaluate_gradients(Id id, const Array<T>& seed) const { return sweep(id, seed, 0); }       // graph.rs:263-274
  template <class T> Store<T> evaluate_gradients_once(Id id, const Array<T>& seed) { return sweep(id, seed, 1); }        // graph.rs:277-290
  template <class T, class = std::enable_if_t<std::is_arithmetic<T>::value>>
  Store<T> evaluate_gradients_once(Id id, T seed) { return sweep(id, Array<T>::scalar(*ctx_, seed), 1); }  // scalar roots (`as_scalar`, `dot`)
  // The sweep of one batch shard: every variable's gradient is summed over the ranks as soon as the sweep has completed it,
  // overlapping the rest of the sweep (gadcu_evaluate_gradients_dp); `also` — e.g. the step's loss — joins the same exchange.
  // With one rank it IS evaluate_gradients[_once].
  template <class T> Store<T> evaluate_gradients_dp(Id id, const Array<T>& seed, Comm& comm, bool once = true, const std::vector<Array<T>*>& also = {}) {
    std::vector<gadcu_array*> extra;
    for (const Array<T>* a : also) extra.push_back(a->handle());
    gadcu_store* s = nullptr;
    check(gadcu_evaluate_gradients_dp_with(h_, id.arena, id.index, seed.handle(), once ? 1 : 0, comm.handle(), extra.data(), extra.size(), &s));
    return Store<T>(s);
  }
  template <class T, class = std::enable_if_t<std::is_arithmetic<T>::value>>
  Store<T> evaluate_gradients_dp(Id id, T seed, Comm& comm, bool once = true, const std::vector<Array<T>*>& also = {}) {
    return evaluate_gradients_dp(id, Array<T>::scalar(*ctx_, seed), comm, once, also);
  }

 private:
  template <class T> Store<T> sweep(Id id, const Array<T>& seed, int once) const {
    gadcu_store* s = nullptr;
    check(gadcu_evaluate_gradients(h_, id.arena, id.index, seed.handle(), once, &s));
    return Store<T>(s);
  }
};

// One scalar Graph1 tape per LANE (gadcu.h "batched scalar tapes"): what the reference does by looping a scalar tape over the
// examples (net_ext.rs:50-66) runs as one kernel, one lane per thread. Variables and constants are [lanes] columns (structure
// of arrays); every op only records; evaluate_gradients[_once] with a scalar seed compiles forward + reverse sweep into one
// per-lane program and launches it. The same generic gad program runs on it and on Graph1 — values handed out here are lazy:
// `data()` is empty, `force(v)` computes a forward value's column; gradients come out of the store as [lanes] columns.
template <class T>
class BatchedGraph1 : public Graph1 {
 public:
  BatchedGraph1(Context& ctx, uint64_t lanes) : Graph1(ctx, create(ctx, lanes), Adopt{}), lanes_(lanes) {}
  uint64_t lanes() const { return lanes_; }
  Array<T> force(const Value<T>& v) {
    const gadcu_value c = v.c();
    gadcu_array* out = nullptr;
    check(gadcu_batched_force(h_, &c, &out));
    return Array<T>::adopt(out);
  }

 private:
  static gadcu_algebra* create(Context& ctx, uint64_t lanes) {
    gadcu_algebra* h = nullptr;
    check(gadcu_batched_create(ctx.handle(), dtype_of<T>::value, lanes, &h));
    return h;
  }
  uint64_t lanes_;
};

// lib.rs:531-532: higher-order tape, the backward pass is recorded
class GraphN : public Algebra {
 public:
  explicit GraphN(Context& ctx) : Algebra(ctx, GADCU_GRAPHN) {}
  GraphN(const GraphN&) = default;
  template <class T> Store<T> compute_gradients(Id id, const Value<T>& seed) {  // graph.rs:307-318
    const gadcu_value v = seed.c();
    gadcu_store* s = nullptr;
    check(gadcu_compute_gradients(h_, id.arena, id.index, &v, &s));
    return Store<T>(s);
  }
  // of one batch shard: each gradient is summed over the ranks when the sweep completes it and returned as a constant
  template <class T> Store<T> compute_gradients_dp(Id id, const Value<T>& seed, Comm& comm) {
    const gadcu_value v = seed.c();
    gadcu_store* s = nullptr;
    check(gadcu_compute_gradients_dp(h_, id.arena, id.index, &v, comm.handle(), &s));
    return Store<T>(s);
  }
};

}  // namespace gadcu
