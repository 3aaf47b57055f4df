// gadcu_net.hpp — header-only C++17 mirror of gad's network layer (gad/src/net.rs, gad/src/net_ext.rs) above gadcu.hpp.
//
// "`Net` combinators + `DiffNet::apply_gradient_step` + `SquareLoss` over the new array type (C++
// mirror; Rust source)". On the Rust side nothing has to be written — gad's own `net.rs` is generic and works as soon as
// `CuArray<T>: WeightOps<T>` (rust/src/lib.rs). This is the compiled-language mirror that is built here: the same
// combinators, the same method names (`using` and `and` are C++ keywords: `using_`, `and_`), the same error variants.
//
//   reference (gad/src)                                           here (namespace gadcu::net)
//   ----------------------------------------------------------   ------------------------------------------------------
//   trait Net<Algebra> {Input, Output, Weights, GradientInfo}     any class derived from NetBase<Self> with the five methods
//     eval_with_gradient_info / get_weights / set_weights /        below; `Output` and `GradientInfo` follow from the algebra a
//     update_weights / read_weight_gradients   net.rs:33-96        net is evaluated with (member templates, deduced)
//   eval, map, using, then, and                 net.rs:58-92      NetBase::{eval, map, using_, then, and_}
//   EvalNet::evaluate, CheckNet::check          net.rs:118-136    NetBase::{evaluate(ctx, input), check(input)}
//   WeightOps {add_assign, scale}               net.rs:100-104    weights_add_assign / weights_scale over the weight structure
//     for af::Array (dims check, then +=)       net.rs:161-180      leaf: anything with add_assign / scale / dims (gadcu::Array<T>)
//     for (), Then/Using pairs, tuples, Vec     net.rs:478-704      Unit, std::pair, std::tuple, std::vector (Lengths on mismatch)
//   InputData / ConstantData / WeightData       net.rs:193-380    the same names
//   Map / Then / Using / tuples / Vec<N>        net.rs:382-704    Map / Then / Using / TupleNet / VecNet
//   SingleOutputNet::add_square_loss, SquareLoss net_ext.rs:17-140 NetBase::add_square_loss, SquareLoss
//   DiffNet::apply_gradient_step                net_ext.rs:39-73  apply_gradient_step(net, lambda, batch, new_graph)
//   (none: the reference is one process)                          apply_gradient_step(net, ctx, lambda, batch, comm): the examples
//                                                                 split over the ranks, update and loss summed before update_weights
//   Check as an algebra (Value = Dim4)          lib.rs:519-520    CheckAlgebra (values are Dim4; delegates to gadcu::Check)
//   bincode::serialize(&net.get_weights())      tests/net.rs:136  serialize_weights_bincode / deserialize_weights_bincode
//
// Nothing in this header computes: nets only call the algebra they are given (`graph.variable`, `graph.constant`, whatever
// the user's `map` closures call) and the WeightOps of their leaves. It is therefore generic over the algebra exactly as the
// reference is — gadcu::Eval / Graph1 / GraphN / CheckAlgebra on the device side, and any other type with the same member
// surface (tests/cpp/net_test.cc instantiates it over the CPU oracle's tape to check the combinator logic without a GPU).
#pragma once
#include <array>
#include <cstring>
#include <optional>
#include <tuple>
#include <utility>
#include <vector>

#include "gadcu.hpp"

namespace gadcu {
namespace net {

// `()`: the weights / gradient info / input of a net that has none (net.rs:258-262, 297-301)
struct Unit {
  bool operator==(const Unit&) const { return true; }
  bool operator!=(const Unit&) const { return false; }
};

// ---- dims and ids of whatever an algebra hands out ---------------------------------------------------------------
// HasDims (core.rs:41-73): arrays have dims(); values have the dims of their data; Check's values ARE dims.
inline Dim4 dims_of(const Dim4& d) { return d; }
template <class X> auto dims_of(const X& x) -> decltype(x.dims()) { return x.dims(); }
template <class X> auto dims_of(const X& x) -> decltype(x.data().dims()) { return x.data().dims(); }
// HasGradientId (net.rs:19-24, 143-159, 183-190): a tracked value has an id; plain data and dims have `()`. A value of
// an algebra that does not track (Eval) carries no id: that is `()` too, and reading a gradient through it finds nothing.
inline Unit gid_of(const Dim4&) { return {}; }
template <class V> auto gid_of(const V& v) -> decltype(v.id()) { return v.id(); }

inline std::string debug(const Dim4& d) {
  return "[" + std::to_string(d[0]) + " " + std::to_string(d[1]) + " " + std::to_string(d[2]) + " " + std::to_string(d[3]) + "]";
}
template <class D> std::string debug(const D&) { return "<dims>"; }
// error.rs:139-153 / 156-167
template <class D> void check_equal_dimensions(const char* name, const D& a, const D& b) {
  if (a != b) throw Error(ErrorKind::Dimensions, std::string("Incompatible dimensions for ") + name + ": " + debug(a) + " " + debug(b));
}
inline void check_equal_lengths(const char* name, size_t a, size_t b) {
  if (a != b) throw Error(ErrorKind::Lengths, std::string("Incompatible lengths for ") + name + ": " + std::to_string(a) + " " + std::to_string(b));
}

// ---- WeightOps over weight structures (net.rs:100-104) ------------------------------------------------------------
// (all overloads are declared before any is defined: they recurse into each other through nested structures)
inline void weights_add_assign(Unit&, const Unit&) {}
template <class L> auto weights_add_assign(L& self, const L& other) -> decltype(self.add_assign(other), void());
template <class A, class B> void weights_add_assign(std::pair<A, B>& self, const std::pair<A, B>& other);
template <class... W> void weights_add_assign(std::tuple<W...>& self, const std::tuple<W...>& other);
template <class W> void weights_add_assign(std::vector<W>& self, const std::vector<W>& other);
template <class S> Unit weights_scale(const Unit&, S) { return {}; }
template <class L, class S> auto weights_scale(const L& w, S lambda) -> decltype(w.scale(lambda));
template <class A, class B, class S> std::pair<A, B> weights_scale(const std::pair<A, B>& w, S lambda);
template <class... W, class S> std::tuple<W...> weights_scale(const std::tuple<W...>& w, S lambda);
template <class W, class S> std::vector<W> weights_scale(const std::vector<W>& w, S lambda);

// a leaf, net.rs:171-179: `check_equal_dimensions(other, self)` then `*self += other`; `self * lambda`
template <class L> auto weights_add_assign(L& self, const L& other) -> decltype(self.add_assign(other), void()) {
  check_equal_dimensions("add_assign", dims_of(other), dims_of(self));
  self.add_assign(other);
}
template <class L, class S> auto weights_scale(const L& w, S lambda) -> decltype(w.scale(lambda)) { return w.scale(lambda); }
// the weights of Then / Using (net.rs:478-493, 545-560)
template <class A, class B> void weights_add_assign(std::pair<A, B>& self, const std::pair<A, B>& other) {
  weights_add_assign(self.first, other.first);
  weights_add_assign(self.second, other.second);
}
template <class A, class B, class S> std::pair<A, B> weights_scale(const std::pair<A, B>& w, S lambda) {
  return std::pair<A, B>(weights_scale(w.first, lambda), weights_scale(w.second, lambda));
}
// tuples (net.rs:603-617)
template <class... W> void weights_add_assign(std::tuple<W...>& self, const std::tuple<W...>& other) {
  std::apply([&](auto&... s) { std::apply([&](const auto&... o) { (weights_add_assign(s, o), ...); }, other); }, self);
}
template <class... W, class S> std::tuple<W...> weights_scale(const std::tuple<W...>& w, S lambda) {
  return std::apply([&](const auto&... x) { return std::tuple<W...>(weights_scale(x, lambda)...); }, w);
}
// Vec (net.rs:689-704): `Lengths` when the two differ
template <class W> void weights_add_assign(std::vector<W>& self, const std::vector<W>& other) {
  check_equal_lengths("add_assign", self.size(), other.size());
  for (size_t i = 0; i < self.size(); i++) weights_add_assign(self[i], other[i]);
}
template <class W, class S> std::vector<W> weights_scale(const std::vector<W>& w, S lambda) {
  std::vector<W> out;
  out.reserve(w.size());
  for (const W& x : w) out.push_back(weights_scale(x, lambda));
  return out;
}

// the leaves of a weight structure, in order (what a data-parallel step hands to the all-reduce)
template <class L> void weights_leaves(Unit&, std::vector<L*>&) {}
template <class L> void weights_leaves(L& leaf, std::vector<L*>& out) { out.push_back(&leaf); }
template <class A, class B, class L> void weights_leaves(std::pair<A, B>& w, std::vector<L*>& out);
template <class... W, class L> void weights_leaves(std::tuple<W...>& w, std::vector<L*>& out);
template <class W, class L> void weights_leaves(std::vector<W>& w, std::vector<L*>& out);
template <class A, class B, class L> void weights_leaves(std::pair<A, B>& w, std::vector<L*>& out) {
  weights_leaves(w.first, out);
  weights_leaves(w.second, out);
}
template <class... W, class L> void weights_leaves(std::tuple<W...>& w, std::vector<L*>& out) {
  std::apply([&](auto&... x) { (weights_leaves(x, out), ...); }, w);
}
template <class W, class L> void weights_leaves(std::vector<W>& w, std::vector<L*>& out) {
  for (W& x : w) weights_leaves(x, out);
}

// ---- Check as an algebra: values are dims (lib.rs:519-520; the per-family `impl ... for Check` blocks) ----
// What `CheckNet::check` evaluates a net with. Scalars (`dot`, `as_scalar`, `norm2`) have the all-ones shape here, as the
// backend's scalars are one-element device values.
class CheckAlgebra {
 public:
  template <class D> Dim4 variable(const D& data) { return dims_of(data); }  // core.rs:141-149
  template <class D> Dim4 constant(const D& data) { return dims_of(data); }
  Dim4 add(Dim4 a, Dim4 b) { return Check::add(a, b); }
  Dim4 add_all(const std::vector<const Dim4*>& vs) {
    if (vs.empty()) throw Error(ErrorKind::Empty, "Unexpected empty input for add_all");
    for (const Dim4* v : vs) Check::add(*vs[0], *v);
    return *vs[0];
  }
  Dim4 sub(Dim4 a, Dim4 b) { return Check::sub(a, b); }
  Dim4 mul(Dim4 a, Dim4 b) { return Check::mul(a, b); }
  Dim4 div(Dim4 a, Dim4 b) { return Check::div(a, b); }
  Dim4 pow(Dim4 a, Dim4 b) { return Check::pow(a, b); }
  Dim4 min(Dim4 a, Dim4 b) { return Check::equal("min", {a, b}); }
  Dim4 max(Dim4 a, Dim4 b) { return Check::equal("max", {a, b}); }
  Dim4 select_argmax(Dim4 v0, Dim4 v1, const Dim4* r0, const Dim4* r1) { return Check::select_argmax(v0, v1, r0, r1); }
#define GADCU_NET_IDENTITY(name) \
  Dim4 name(Dim4 v) { return v; }
  GADCU_NET_IDENTITY(zeros) GADCU_NET_IDENTITY(ones) GADCU_NET_IDENTITY(neg) GADCU_NET_IDENTITY(exp) GADCU_NET_IDENTITY(log)
  GADCU_NET_IDENTITY(log1p) GADCU_NET_IDENTITY(sin) GADCU_NET_IDENTITY(cos) GADCU_NET_IDENTITY(tanh) GADCU_NET_IDENTITY(sigmoid)
  GADCU_NET_IDENTITY(reciprocal) GADCU_NET_IDENTITY(sqrt) GADCU_NET_IDENTITY(abs) GADCU_NET_IDENTITY(sign) GADCU_NET_IDENTITY(relu)
#undef GADCU_NET_IDENTITY
  template <class C> Dim4 setc(Dim4 v, C) { return v; }
  template <class C> Dim4 addc(Dim4 v, C) { return v; }
  template <class C> Dim4 mulc(Dim4 v, C) { return v; }
  template <class C> Dim4 powc(Dim4 v, C) { return v; }
  Dim4 flat(Dim4 v) { return Check::flat(v); }
  Dim4 moddims(Dim4 v, Dim4 d) { return Check::moddims(v, d); }
  Dim4 tile_as(Dim4 v, Dim4 d) { return Check::tile_as(v, d); }
  Dim4 sum_as(Dim4 v, Dim4 d) { return Check::sum_as(v, d); }
  Dim4 constant_as(Dim4, Dim4 d) { return d; }
  Dim4 as_scalar(Dim4 v) { Check::as_scalar(v); return Dim4{}; }
  Dim4 scale(Dim4, Dim4 v) { return v; }
  Dim4 dot(Dim4 a, Dim4 b) { Check::dot(a, b); return Dim4{}; }
  Dim4 norm2(Dim4 v) { return dot(v, v); }
  Dim4 matmul(Dim4 a, Dim4 b, MatProp p1 = {}, MatProp p2 = {}) { return Check::matmul(a, b, p1, p2); }
  Dim4 matmul_nn(Dim4 a, Dim4 b) { return Check::matmul(a, b); }
  Dim4 transpose(Dim4 v, bool = false) { return Check::transpose(v); }
  Dim4 max_as(Dim4 v, Dim4 d) { return Check::max_as(v, d); }
  Dim4 argmax_as(Dim4 v, Dim4 d) { return Check::argmax_as(v, d); }
  Dim4 softmax_as(Dim4 v, Dim4 d) { return Check::softmax_as(v, d); }
};

template <class N, class F> class Map;
template <class N1, class N2> class Then;
template <class N1, class N2> class Using;
template <class... N> class TupleNet;
template <class N, class Data> class SquareLoss;

// ---- Net: the provided methods (net.rs:58-92, 118-136; net_ext.rs:17-27) -------------------------------------------
// A net derives from NetBase<Self>, names its `Input` and `Weights`, and implements
//   template <class A> std::pair<Output, GradientInfo> eval_with_gradient_info(A& graph, const Input& input) const;
//   Weights get_weights() const;   void set_weights(const Weights&);   void update_weights(const Weights& delta);
//   template <class Info, class Reader> Weights read_weight_gradients(const Info&, const Reader& store) const;
template <class Self>
class NetBase {
 public:
  template <class A, class I> auto eval(A& graph, const I& input) const { return self().eval_with_gradient_info(graph, input).first; }
  template <class F> Map<Self, F> map(F f) const { return Map<Self, F>(self(), std::move(f)); }
  template <class N2> Using<Self, N2> using_(N2 net) const { return Using<Self, N2>(self(), std::move(net)); }
  template <class N2> Then<Self, N2> then(N2 net) const { return Then<Self, N2>(self(), std::move(net)); }
  template <class N2> TupleNet<Self, N2> and_(N2 net) const { return TupleNet<Self, N2>(self(), std::move(net)); }
  // the L2 distance (squared, net_ext.rs:116-117) between the output and a target handed in beside the input
  template <class Data> SquareLoss<Self, Data> add_square_loss() const { return SquareLoss<Self, Data>(self()); }
  template <class I> auto evaluate(Context& ctx, const I& input) const {  // EvalNet::evaluate: a forward pass on Eval
    Eval e(ctx);
    return eval(e, input);
  }
  template <class I> auto check(const I& input) const {  // CheckNet::check: dimensions only, nothing is launched
    CheckAlgebra c;
    return eval(c, input);
  }

 private:
  const Self& self() const { return static_cast<const Self&>(*this); }
};

// ---- leaves (net.rs:193-380) ----------------------------------------------------------------------------------------
// The input of the network: checked against the declared dims, recorded as a constant (net.rs:254-291).
template <class Data>
class InputData : public NetBase<InputData<Data>> {
 public:
  using Input = Data;
  using Weights = Unit;
  using Dims = decltype(dims_of(std::declval<const Data&>()));
  explicit InputData(Dims dims) : dims_(dims) {}
  template <class A> auto eval_with_gradient_info(A& graph, const Input& input) const {
    check_equal_dimensions("eval_with_gradient_info", dims_of(input), dims_);
    return std::make_pair(graph.constant(input), Unit{});
  }
  Weights get_weights() const { return {}; }
  void set_weights(const Weights&) {}
  void update_weights(const Weights&) {}
  template <class Info, class Reader> Weights read_weight_gradients(const Info&, const Reader&) const { return {}; }

 private:
  Dims dims_;
};

// A constant the network carries (net.rs:293-328). Input is `()`.
template <class Data>
class ConstantData : public NetBase<ConstantData<Data>> {
 public:
  using Input = Unit;
  using Weights = Unit;
  explicit ConstantData(Data data) : data_(std::move(data)) {}
  const Data& get() const { return data_; }
  template <class A> auto eval_with_gradient_info(A& graph, const Input& = {}) const { return std::make_pair(graph.constant(data_), Unit{}); }
  Weights get_weights() const { return {}; }
  void set_weights(const Weights&) {}
  void update_weights(const Weights&) {}
  template <class Info, class Reader> Weights read_weight_gradients(const Info&, const Reader&) const { return {}; }

 private:
  Data data_;
};

// A trainable array, recorded as a variable; its gradient info is the variable's id (net.rs:330-380). Input is `()`.
template <class Data>
class WeightData : public NetBase<WeightData<Data>> {
 public:
  using Input = Unit;
  using Weights = Data;
  explicit WeightData(Data data) : data_(std::move(data)) {}
  const Data& get() const { return data_; }
  template <class A> auto eval_with_gradient_info(A& graph, const Input& = {}) const {
    auto value = graph.variable(data_);
    auto id = gid_of(value);
    return std::make_pair(std::move(value), std::move(id));
  }
  Weights get_weights() const { return data_; }  // a clone: an O(1) handle copy that a later `+=` leaves alone (copy-on-write)
  void set_weights(const Weights& weights) {
    check_equal_dimensions("set_weights", dims_of(weights), dims_of(data_));
    data_ = weights;
  }
  void update_weights(const Weights& delta) {
    check_equal_dimensions("update_weights", dims_of(delta), dims_of(data_));
    data_.add_assign(delta);  // `self.data += delta`
  }
  // `reader.read(info)`, else MissingGradient — also when the algebra tracked nothing (Eval, Check: EmptyGradientMap)
  template <class Reader> Weights read_weight_gradients(const Unit&, const Reader&) const { throw missing(); }
  template <class I, class Reader> Weights read_weight_gradients(const std::optional<I>& info, const Reader& store) const {
    if (!info) throw missing();
    auto gradient = store.get(*info);
    if (!gradient) throw missing();
    return gradient->data();
  }

 private:
  static Error missing() { return Error(ErrorKind::MissingGradient, "No gradient of the expected type could be found in gradient store."); }
  Data data_;
};

// ---- combinators (net.rs:382-704) -----------------------------------------------------------------------------------
// `f(graph, output)` applied to the output of a net; weights are the inner net's (net.rs:384-426).
template <class N, class F>
class Map : public NetBase<Map<N, F>> {
 public:
  using Input = typename N::Input;
  using Weights = typename N::Weights;
  Map(N net, F f) : net_(std::move(net)), f_(std::move(f)) {}
  template <class A> auto eval_with_gradient_info(A& graph, const Input& input) const {
    auto r = net_.eval_with_gradient_info(graph, input);
    return std::make_pair(f_(graph, r.first), std::move(r.second));
  }
  Weights get_weights() const { return net_.get_weights(); }
  void set_weights(const Weights& w) { net_.set_weights(w); }
  void update_weights(const Weights& d) { net_.update_weights(d); }
  template <class Info, class Reader> Weights read_weight_gradients(const Info& info, const Reader& store) const {
    return net_.read_weight_gradients(info, store);
  }

 private:
  N net_;
  F f_;
};

namespace detail {
template <class N1, class N2>
class PairNet {
 public:
  using Weights = std::pair<typename N1::Weights, typename N2::Weights>;
  PairNet(N1 first, N2 second) : first_(std::move(first)), second_(std::move(second)) {}
  Weights get_weights() const { return Weights(first_.get_weights(), second_.get_weights()); }
  void set_weights(const Weights& w) {
    first_.set_weights(w.first);
    second_.set_weights(w.second);
  }
  void update_weights(const Weights& d) {
    first_.update_weights(d.first);
    second_.update_weights(d.second);
  }
  template <class I1, class I2, class Reader> Weights read_weight_gradients(const std::pair<I1, I2>& info, const Reader& store) const {
    return Weights(first_.read_weight_gradients(info.first, store), second_.read_weight_gradients(info.second, store));
  }

 protected:
  N1 first_;
  N2 second_;
};
}  // namespace detail

// The first net's output is the second net's input (net.rs:429-493).
template <class N1, class N2>
class Then : public detail::PairNet<N1, N2>, public NetBase<Then<N1, N2>> {
 public:
  using Input = typename N1::Input;
  using detail::PairNet<N1, N2>::PairNet;
  template <class A> auto eval_with_gradient_info(A& graph, const Input& input) const {
    auto r0 = this->first_.eval_with_gradient_info(graph, input);
    auto r1 = this->second_.eval_with_gradient_info(graph, r0.first);
    return std::make_pair(std::move(r1.first), std::make_pair(std::move(r0.second), std::move(r1.second)));
  }
};

// The first net's output paired with the output of a net that takes no input — weights, constants (net.rs:496-560).
template <class N1, class N2>
class Using : public detail::PairNet<N1, N2>, public NetBase<Using<N1, N2>> {
 public:
  using Input = typename N1::Input;
  using detail::PairNet<N1, N2>::PairNet;
  template <class A> auto eval_with_gradient_info(A& graph, const Input& input) const {
    auto r0 = this->first_.eval_with_gradient_info(graph, input);
    auto r1 = this->second_.eval_with_gradient_info(graph, Unit{});
    return std::make_pair(std::make_pair(std::move(r0.first), std::move(r1.first)), std::make_pair(std::move(r0.second), std::move(r1.second)));
  }
};

// A tuple of nets maps a tuple of inputs to a tuple of outputs (net.rs:563-629).
template <class... N>
class TupleNet : public NetBase<TupleNet<N...>> {
 public:
  using Input = std::tuple<typename N::Input...>;
  using Weights = std::tuple<typename N::Weights...>;
  explicit TupleNet(N... nets) : nets_(std::move(nets)...) {}
  template <class A> auto eval_with_gradient_info(A& graph, const Input& input) const {
    return eval_impl(graph, input, std::index_sequence_for<N...>{});
  }
  Weights get_weights() const {
    return std::apply([](const auto&... n) { return Weights(n.get_weights()...); }, nets_);
  }
  void set_weights(const Weights& w) { each(w, [](auto& n, const auto& x) { n.set_weights(x); }, std::index_sequence_for<N...>{}); }
  void update_weights(const Weights& d) { each(d, [](auto& n, const auto& x) { n.update_weights(x); }, std::index_sequence_for<N...>{}); }
  template <class... I, class Reader> Weights read_weight_gradients(const std::tuple<I...>& info, const Reader& store) const {
    return read_impl(info, store, std::index_sequence_for<N...>{});
  }

 private:
  template <class A, size_t... K> auto eval_impl(A& graph, const Input& input, std::index_sequence<K...>) const {
    // (braced initialisation: the nets are evaluated left to right, so node ids follow the tuple's order as in the reference)
    std::tuple<decltype(std::get<K>(nets_).eval_with_gradient_info(graph, std::get<K>(input)))...> results{
        std::get<K>(nets_).eval_with_gradient_info(graph, std::get<K>(input))...};
    return std::make_pair(std::make_tuple(std::move(std::get<K>(results).first)...), std::make_tuple(std::move(std::get<K>(results).second)...));
  }
  template <class W, class F, size_t... K> void each(const W& w, F f, std::index_sequence<K...>) { (f(std::get<K>(nets_), std::get<K>(w)), ...); }
  template <class Info, class Reader, size_t... K> Weights read_impl(const Info& info, const Reader& store, std::index_sequence<K...>) const {
    return Weights{std::get<K>(nets_).read_weight_gradients(std::get<K>(info), store)...};
  }
  std::tuple<N...> nets_;
};

// `Vec<N>`: inputs, outputs, weights and gradient infos are vectors; `Lengths` when they do not match (net.rs:631-704).
template <class N>
class VecNet : public NetBase<VecNet<N>> {
 public:
  using Input = std::vector<typename N::Input>;
  using Weights = std::vector<typename N::Weights>;
  explicit VecNet(std::vector<N> nets) : nets_(std::move(nets)) {}
  size_t size() const { return nets_.size(); }
  template <class A> auto eval_with_gradient_info(A& graph, const Input& input) const {
    check_equal_lengths("eval_with_gradient_info", nets_.size(), input.size());
    using R = decltype(std::declval<const N&>().eval_with_gradient_info(graph, input[0]));
    std::vector<typename R::first_type> outputs;
    std::vector<typename R::second_type> infos;
    for (size_t i = 0; i < nets_.size(); i++) {
      R r = nets_[i].eval_with_gradient_info(graph, input[i]);
      outputs.push_back(std::move(r.first));
      infos.push_back(std::move(r.second));
    }
    return std::make_pair(std::move(outputs), std::move(infos));
  }
  Weights get_weights() const {
    Weights w;
    for (const N& n : nets_) w.push_back(n.get_weights());
    return w;
  }
  void set_weights(const Weights& w) {
    check_equal_lengths("set_weights", nets_.size(), w.size());
    for (size_t i = 0; i < nets_.size(); i++) nets_[i].set_weights(w[i]);
  }
  void update_weights(const Weights& d) {
    check_equal_lengths("update_weights", nets_.size(), d.size());
    for (size_t i = 0; i < nets_.size(); i++) nets_[i].update_weights(d[i]);
  }
  template <class I, class Reader> Weights read_weight_gradients(const std::vector<I>& info, const Reader& store) const {
    check_equal_lengths("read_weight_gradients", nets_.size(), info.size());
    Weights w;
    for (size_t i = 0; i < nets_.size(); i++) w.push_back(nets_[i].read_weight_gradients(info[i], store));
    return w;
  }

 private:
  std::vector<N> nets_;
};

// ---- SquareLoss (net_ext.rs:85-140) ---------------------------------------------------------------------------------
// Input = (the net's input, a target); output = norm2(target − net(input)): the SQUARED L2 norm, no ½ and no mean
// (array.rs:42-44). Output dims are checked against the target's before anything is recorded.
template <class N, class Data>
class SquareLoss : public NetBase<SquareLoss<N, Data>> {
 public:
  using Input = std::pair<typename N::Input, Data>;
  using Weights = typename N::Weights;
  explicit SquareLoss(N net) : net_(std::move(net)) {}
  template <class A> auto eval_with_gradient_info(A& graph, const Input& input) const {
    auto r = net_.eval_with_gradient_info(graph, input.first);
    check_equal_dimensions("eval_with_gradient_info", dims_of(r.first), dims_of(input.second));
    auto target = graph.constant(input.second);
    auto delta = graph.sub(target, r.first);
    return std::make_pair(graph.norm2(delta), std::move(r.second));
  }
  Weights get_weights() const { return net_.get_weights(); }
  void set_weights(const Weights& w) { net_.set_weights(w); }
  void update_weights(const Weights& d) { net_.update_weights(d); }
  template <class Info, class Reader> Weights read_weight_gradients(const Info& info, const Reader& store) const {
    return net_.read_weight_gradients(info, store);
  }

 private:
  N net_;
};

// ---- DiffNet::apply_gradient_step (net_ext.rs:39-73) ----------------------------------------------------------------
// One "mini-batch" gradient step. Per example: a fresh first-order tape from `new_graph()`, the forward pass, the reverse
// sweep from the (scalar) output with seed 1, `delta += lambda * gradients`; then ONE `update_weights(delta)`. Returns the
// cumulated output; `Empty` for an empty batch. `lambda` is negative for loss minimisation.
namespace detail {
// the per-example body of net_ext.rs:51-67 over batch[first, last)
template <class N, class T, class NewGraph>
void accumulate_examples(N& net, T lambda, const std::vector<typename N::Input>& batch, size_t first, size_t last, NewGraph& new_graph,
                         std::optional<typename N::Weights>& delta, std::optional<T>& cumulated_output) {
  for (size_t e = first; e < last; e++) {
    auto g = new_graph();  // Graph1::new()
    auto r = net.eval_with_gradient_info(g, batch[e]);
    const T value = T(r.first.data().item());
    cumulated_output = cumulated_output ? T(*cumulated_output + value) : value;
    auto store = g.evaluate_gradients_once(r.first.gid(), T(1));
    auto gradients = net.read_weight_gradients(r.second, store);
    if (!delta) delta = weights_scale(gradients, lambda);
    else weights_add_assign(*delta, weights_scale(gradients, lambda));
  }
}
}  // namespace detail

template <class N, class T, class NewGraph>
T apply_gradient_step(N& net, T lambda, const std::vector<typename N::Input>& batch, NewGraph&& new_graph) {
  std::optional<typename N::Weights> delta;
  std::optional<T> cumulated_output;
  detail::accumulate_examples(net, lambda, batch, 0, batch.size(), new_graph, delta, cumulated_output);
  if (delta) net.update_weights(*delta);
  if (!cumulated_output) throw Error(ErrorKind::Empty, "Unexpected empty input for apply_gradient_step");
  return *cumulated_output;
}
// the device's first-order tape (lib.rs:527-528)
template <class N, class T>
T apply_gradient_step(N& net, Context& ctx, T lambda, const std::vector<typename N::Input>& batch) {
  return apply_gradient_step(net, lambda, batch, [&ctx] { return Graph1(ctx); });
}
// [first, last) of `rank`'s contiguous share of `total` units: the first `total % world` ranks take one more
// (the same split as gad_b200/dp.py: shard_bounds)
inline std::pair<size_t, size_t> shard_bounds(size_t total, size_t world, size_t rank) {
  if (world < 1 || rank >= world) throw Error(ErrorKind::Invalid, "shard_bounds: rank " + std::to_string(rank) + " of " + std::to_string(world));
  const size_t base = total / world, extra = total % world;
  const size_t first = rank * base + (rank < extra ? rank : extra);
  return {first, first + base + (rank < extra ? 1 : 0)};
}
// Over the GPUs of a box (one process each; every rank passes the SAME batch): a rank works on its contiguous share of the
// examples — the first `batch.size() % nranks` ranks take one more — and the weight update and the loss are summed over the
// ranks before update_weights, so every rank ends with the same weights. The reference sums per-example gradients
// (net_ext.rs:62-65): the result differs from the one-process step only by the association of that sum. Returns the loss
// of the WHOLE batch on every rank. Leaves of N::Weights are Array<T>.
template <class N, class T>
T apply_gradient_step(N& net, Context& ctx, T lambda, const std::vector<typename N::Input>& batch, Comm& comm) {
  if (comm.nranks() <= 1) return apply_gradient_step(net, ctx, lambda, batch);
  if (batch.empty()) throw Error(ErrorKind::Empty, "Unexpected empty input for apply_gradient_step");
  const std::pair<size_t, size_t> share = shard_bounds(batch.size(), size_t(comm.nranks()), size_t(comm.rank()));
  const size_t first = share.first, last = share.second;
  std::optional<typename N::Weights> delta;
  std::optional<T> cumulated_output;
  auto new_graph = [&ctx] { return Graph1(ctx); };
  detail::accumulate_examples(net, lambda, batch, first, last, new_graph, delta, cumulated_output);
  if (!delta) delta = weights_scale(net.get_weights(), T(0));  // a rank without examples contributes zeros
  Array<T> loss = Array<T>::from_host(ctx, {cumulated_output ? *cumulated_output : T(0)}, Dim4{1});
  std::vector<Array<T>*> leaves;
  weights_leaves(*delta, leaves);
  leaves.push_back(&loss);
  comm.allreduce_sum(leaves);
  net.update_weights(*delta);
  return loss.host()[0];
}

// ---- checkpoints: the reference's own bytes (tests/net.rs:136-149) ---------------------------------------------------
// `bincode::serialize(&net.get_weights())` with the arrayfire crate's serde impl for `Array<T>` (crate 3.8.0, `afserde`;
// not under /root/reference — restated from the published crate, so the layout is unpinned against a real blob; it is the
// one gad_b200/net.py writes, and tests/cpp/net_test.cc holds the two against each other): per array a `u32` DType index
// (f32 = 0, f64 = 2), the Dim4 as 4 × `u64`, the element count as `u64`, then the elements, all little endian (bincode 1.x
// fixed-width integers); `()` is empty, tuple structs and tuples are their fields one after the other, a `Vec` is a `u64`
// length and its elements.
// How a leaf is read from / rebuilt on a device; specialise for other leaf types (the CPU tests do).
template <class L> struct LeafIO {};
template <class T> struct LeafIO<Array<T>> {
  using Scalar = T;
  using Maker = Context;
  static std::vector<T> host(const Array<T>& a) { return a.host(); }
  static std::array<uint64_t, 4> dims(const Array<T>& a) {
    const Dim4 d = a.dims();
    return {d[0], d[1], d[2], d[3]};
  }
  static Array<T> make(Context& ctx, const std::vector<T>& values, const std::array<uint64_t, 4>& d) {
    return Array<T>::from_host(ctx, values, Dim4{d[0], d[1], d[2], d[3]});
  }
};
namespace detail {
template <class I> void put_le(std::vector<uint8_t>& out, I v) {
  for (size_t i = 0; i < sizeof(I); i++) out.push_back(uint8_t(uint64_t(v) >> (8 * i)));
}
template <class I> I get_le(const std::vector<uint8_t>& in, size_t& pos) {
  if (pos + sizeof(I) > in.size()) throw Error(ErrorKind::Invalid, "truncated bincode weight blob");
  uint64_t v = 0;
  for (size_t i = 0; i < sizeof(I); i++) v |= uint64_t(in[pos + i]) << (8 * i);
  pos += sizeof(I);
  return I(v);
}
template <class T> constexpr uint32_t af_dtype() {
  static_assert(std::is_same<T, float>::value || std::is_same<T, double>::value, "weights are f32 or f64 arrays");
  return std::is_same<T, float>::value ? 0u : 2u;  // af::DType::F32 = 0, F64 = 2
}
inline void put(std::vector<uint8_t>&, const Unit&) {}
template <class L> auto put(std::vector<uint8_t>& out, const L& leaf) -> decltype(LeafIO<L>::host(leaf), void());
template <class A, class B> void put(std::vector<uint8_t>& out, const std::pair<A, B>& w);
template <class... W> void put(std::vector<uint8_t>& out, const std::tuple<W...>& w);
template <class W> void put(std::vector<uint8_t>& out, const std::vector<W>& w);
template <class L> auto put(std::vector<uint8_t>& out, const L& leaf) -> decltype(LeafIO<L>::host(leaf), void()) {
  using T = typename LeafIO<L>::Scalar;
  const std::vector<T> values = LeafIO<L>::host(leaf);
  put_le<uint32_t>(out, af_dtype<T>());
  for (uint64_t d : LeafIO<L>::dims(leaf)) put_le<uint64_t>(out, d);
  put_le<uint64_t>(out, values.size());
  for (const T& v : values) {
    typename std::conditional<sizeof(T) == 4, uint32_t, uint64_t>::type bits;
    std::memcpy(&bits, &v, sizeof(T));
    put_le(out, bits);
  }
}
template <class A, class B> void put(std::vector<uint8_t>& out, const std::pair<A, B>& w) {
  put(out, w.first);
  put(out, w.second);
}
template <class... W> void put(std::vector<uint8_t>& out, const std::tuple<W...>& w) {
  std::apply([&](const auto&... x) { (put(out, x), ...); }, w);
}
template <class W> void put(std::vector<uint8_t>& out, const std::vector<W>& w) {
  put_le<uint64_t>(out, w.size());
  for (const W& x : w) put(out, x);
}

template <class W> struct Getter;  // the weight TYPE drives the decoder, as the Rust type drives bincode's
template <> struct Getter<Unit> {
  template <class M> static Unit get(M&, const std::vector<uint8_t>&, size_t&) { return {}; }
};
template <class L> struct Getter {
  template <class M> static L get(M& maker, const std::vector<uint8_t>& in, size_t& pos) {
    using T = typename LeafIO<L>::Scalar;
    const uint32_t dt = get_le<uint32_t>(in, pos);
    if (dt != af_dtype<T>()) throw Error(ErrorKind::Type, "af::DType index " + std::to_string(dt) + " in the weight blob does not match the weight type");
    std::array<uint64_t, 4> d;
    for (uint64_t& x : d) x = get_le<uint64_t>(in, pos);
    const uint64_t count = get_le<uint64_t>(in, pos);
    // (written so that absurd dims in a corrupt blob cannot overflow the product)
    uint64_t elements = 1;
    for (uint64_t x : d) {
      if (x != 0 && elements > (uint64_t(1) << 40) / x) throw Error(ErrorKind::Invalid, "dims in the weight blob are not credible");
      elements *= x;
    }
    if (count != elements) throw Error(ErrorKind::Invalid, "element count does not match dims in the weight blob");
    if (count > (in.size() - pos) / sizeof(T)) throw Error(ErrorKind::Invalid, "truncated bincode weight blob");
    std::vector<T> values(count);
    for (T& v : values) {
      const auto bits = get_le<typename std::conditional<sizeof(T) == 4, uint32_t, uint64_t>::type>(in, pos);
      std::memcpy(&v, &bits, sizeof(T));
    }
    return LeafIO<L>::make(maker, values, d);
  }
};
template <class A, class B> struct Getter<std::pair<A, B>> {
  template <class M> static std::pair<A, B> get(M& maker, const std::vector<uint8_t>& in, size_t& pos) {
    A a = Getter<A>::get(maker, in, pos);  // (sequenced: first, then second)
    B b = Getter<B>::get(maker, in, pos);
    return std::pair<A, B>(std::move(a), std::move(b));
  }
};
template <class... W> struct Getter<std::tuple<W...>> {
  template <class M> static std::tuple<W...> get(M& maker, const std::vector<uint8_t>& in, size_t& pos) {
    return std::tuple<W...>{Getter<W>::get(maker, in, pos)...};  // (braced: left to right)
  }
};
template <class W> struct Getter<std::vector<W>> {
  template <class M> static std::vector<W> get(M& maker, const std::vector<uint8_t>& in, size_t& pos) {
    const uint64_t n = get_le<uint64_t>(in, pos);
    if (n > in.size()) throw Error(ErrorKind::Invalid, "vector length in the weight blob is not credible");
    std::vector<W> out;
    for (uint64_t i = 0; i < n; i++) out.push_back(Getter<W>::get(maker, in, pos));
    return out;
  }
};
}  // namespace detail

template <class W> std::vector<uint8_t> serialize_weights_bincode(const W& weights) {
  std::vector<uint8_t> out;
  detail::put(out, weights);
  return out;
}
// `maker` is what a leaf is rebuilt with: the gadcu::Context the arrays are to live on
template <class W, class M> W deserialize_weights_bincode(M& maker, const std::vector<uint8_t>& bytes) {
  size_t pos = 0;
  W w = detail::Getter<W>::get(maker, bytes, pos);
  if (pos != bytes.size()) throw Error(ErrorKind::Invalid, "trailing bytes in the bincode weight blob");
  return w;
}

}  // namespace net
}  // namespace gadcu
