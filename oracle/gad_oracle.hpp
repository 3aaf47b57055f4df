// ORACLE — TEST INFRASTRUCTURE ONLY (see host_array.hpp header). A CPU restatement, in C++20, of
// gad's tape-based reverse-mode autodiff:
//   gad/src/graph.rs   (Graph, Node, Value, make_node, do_compute_gradients[_once], Config1/N)
//   gad/src/store.rs   (Id ordering, GradientStore::add_gradient, GenericGradientMap1/N)
//   gad/src/linked.rs  (LinkedAlgebra::link)
//   gad/src/{core,arith,const_arith,analytic,compare,array,array_compare,matrix}.rs
//                      (Check / Eval / Graph impls; VJP bodies transcribed op for op)
// Each function cites the reference lines it follows. The reference cannot be compiled here (no
// Rust toolchain, no arrayfire), so this restatement is pinned instead against every known-answer
// test the reference holds for this path — see oracle/test_oracle.cc, which replays
// gad/tests/{graph,symbolic,core,arith,const_arith,analytic,compare,extension}.rs and the lib.rs
// doctests with their exact expected values.
#pragma once
#include <any>
#include <atomic>
#include <functional>
#include <map>
#include <optional>
#include <queue>
#include <type_traits>

#include "host_array.hpp"

namespace gad_oracle {

// ---------------------------------------------------------------------------------------------
// Id — store.rs:12-16 (derive(PartialOrd, Ord) over (arena_id, index)), :155-178 (1-based index),
// :204-212 (next_id).
// ---------------------------------------------------------------------------------------------
struct Id {
  uint32_t arena = 0;
  uint32_t index = 0;  // NonZeroU32 in the reference
  auto operator<=>(const Id&) const = default;
  Id next_id() const { return Id{arena, index + 1}; }
};

inline std::atomic<uint32_t>& arena_counter() {
  static std::atomic<uint32_t> c{0};
  return c;
}

// Value<D> — graph.rs:35-43, 321-342.
template <class D>
struct Value {
  D data{};
  std::optional<Id> id;
  static Value constant(D d) { return Value{std::move(d), std::nullopt}; }
  std::optional<Id> input() const { return id; }
  // net.rs:183-190 (HasGradientId for Value): MissingId for constants.
  Id gid() const {
    if (!id) throw Error(ErrKind::MissingId, "gid");
    return *id;
  }
};

// HasDims — core.rs:41-73, 120-136.
template <class T>
  requires std::is_arithmetic_v<T>
Unit dims(const T&) { return {}; }
template <class T>
Dim4 dims(const HostArray<T>& a) { return a.dims(); }
inline Dim4 dims(const Dim4& d) { return d; }
inline Unit dims(const Unit&) { return {}; }
template <class D>
auto dims(const Value<D>& v) { return dims(v.data); }

// ---------------------------------------------------------------------------------------------
// Check — lib.rs:519-520; rules per family.
// ---------------------------------------------------------------------------------------------
struct Check {
  // core.rs:75-93 (scalars), :138-160 (arrays)
  template <class T>
    requires std::is_arithmetic_v<T>
  Unit variable(const T&) { return {}; }
  template <class T>
    requires std::is_arithmetic_v<T>
  Unit constant(const T&) { return {}; }
  template <class T>
  Dim4 variable(const HostArray<T>& a) { return a.dims(); }
  template <class T>
  Dim4 constant(const HostArray<T>& a) { return a.dims(); }
  Dim4 variable(const Dim4& d) { return d; }
  Dim4 constant(const Dim4& d) { return d; }

  Unit add(Unit, Unit) { return {}; }
  Dim4 add(const Dim4& a, const Dim4& b) { return check_equal_dimensions<Dim4>("add", {&a, &b}); }
  Unit add_all(const std::vector<const Unit*>&) { return {}; }  // core.rs:89-92: always Ok(())
  Dim4 add_all(const std::vector<const Dim4*>& v) { return check_equal_dimensions<Dim4>("add_all", v); }

  // arith.rs:78-101, 132-151
  Unit zeros(Unit) { return {}; }
  Unit ones(Unit) { return {}; }
  Unit neg(Unit) { return {}; }
  Unit sub(Unit, Unit) { return {}; }
  Unit mul(Unit, Unit) { return {}; }
  Dim4 zeros(const Dim4& v) { return v; }
  Dim4 ones(const Dim4& v) { return v; }
  Dim4 neg(const Dim4& v) { return v; }
  Dim4 sub(const Dim4& a, const Dim4& b) { return check_equal_dimensions<Dim4>("sub", {&a, &b}); }
  Dim4 mul(const Dim4& a, const Dim4& b) { return check_equal_dimensions<Dim4>("mul", {&a, &b}); }

  // const_arith.rs:58-78, 106-118
  template <class C> Unit setc(Unit, C) { return {}; }
  template <class C> Unit addc(Unit, C) { return {}; }
  template <class C> Unit mulc(Unit, C) { return {}; }
  template <class C> Unit powc(Unit, C) { return {}; }
  template <class C> Dim4 setc(const Dim4& v, C) { return v; }
  template <class C> Dim4 addc(const Dim4& v, C) { return v; }
  template <class C> Dim4 mulc(const Dim4& v, C) { return v; }
  template <class C> Dim4 powc(const Dim4& v, C) { return v; }

  // analytic.rs:131-184, 248-285
#define GAD_ORACLE_CHECK_UNARY(op) \
  Unit op(Unit) { return {}; }     \
  Dim4 op(const Dim4& v) { return v; }
  GAD_ORACLE_CHECK_UNARY(exp)
  GAD_ORACLE_CHECK_UNARY(log)
  GAD_ORACLE_CHECK_UNARY(log1p)
  GAD_ORACLE_CHECK_UNARY(sin)
  GAD_ORACLE_CHECK_UNARY(cos)
  GAD_ORACLE_CHECK_UNARY(tanh)
  GAD_ORACLE_CHECK_UNARY(sigmoid)
  GAD_ORACLE_CHECK_UNARY(reciprocal)
  GAD_ORACLE_CHECK_UNARY(sqrt)
#undef GAD_ORACLE_CHECK_UNARY
  Unit div(Unit, Unit) { return {}; }
  Unit pow(Unit, Unit) { return {}; }
  Dim4 div(const Dim4& a, const Dim4& b) { return check_equal_dimensions<Dim4>("div", {&a, &b}); }
  Dim4 pow(const Dim4& a, const Dim4& b) { return check_equal_dimensions<Dim4>("pow", {&a, &b}); }

  // compare.rs:117-144, 177-188
  Unit select_argmax(Unit, Unit, const Unit*, const Unit*) { return {}; }
  Unit min(Unit, Unit) { return {}; }
  Unit max(Unit, Unit) { return {}; }
  Dim4 min(const Dim4& a, const Dim4& b) { return check_equal_dimensions<Dim4>("min", {&a, &b}); }
  Dim4 max(const Dim4& a, const Dim4& b) { return check_equal_dimensions<Dim4>("max", {&a, &b}); }
  Dim4 select_argmax(const Dim4& v0, const Dim4& v1, const Dim4* r0, const Dim4* r1) {
    check_equal_dimensions<Dim4>("select_argmax", {&v0, &v1});
    if (r0) check_equal_dimensions<Dim4>("select_argmax", {&v0, r0});
    if (r1) check_equal_dimensions<Dim4>("select_argmax", {&v1, r1});
    return v0;
  }

  // array.rs:129-193
  Dim4 flat(const Dim4& v) { return Dim4(v.elements()); }
  Dim4 moddims(const Dim4& v, const Dim4& d) {
    if (v.elements() != d.elements()) throw Error(ErrKind::Dimensions, "moddims", debug(v) + " " + debug(d));
    return d;
  }
  Dim4 tile_as(const Dim4& v, const Dim4& r) {
    for (int i = 0; i < 4; i++)
      if (r[i] % v[i] != 0) throw Error(ErrKind::Dimensions, "tile_as", debug(v) + " " + debug(r));
    return r;
  }
  Dim4 sum_as(const Dim4& v, const Dim4& r) {
    for (int i = 0; i < 4; i++) {
      if (r[i] == v[i]) continue;
      if (r[i] != 1) throw Error(ErrKind::Dimensions, "sum_as", debug(v) + " " + debug(r));
    }
    return r;
  }
  Dim4 constant_as(Unit, const Dim4& d) { return d; }
  Unit as_scalar(const Dim4& v) {
    Dim4 one;
    check_equal_dimensions<Dim4>("as_scalar", {&v, &one});
    return {};
  }
  Dim4 scale(Unit, const Dim4& v) { return v; }
  Unit dot(const Dim4& a, const Dim4& b) {
    check_equal_dimensions<Dim4>("dot", {&a, &b});
    return {};
  }
  Unit norm2(const Dim4& v) { return dot(v, v); }

  // array_compare.rs:84-101
  Dim4 max_as(const Dim4& v, const Dim4& r) { return check_reduced_dimensions("max_as", v, r); }
  Dim4 argmax_as(const Dim4& v, const Dim4& r) {
    check_reduced_dimensions("argmax_as", v, r);
    return v;
  }
  Dim4 softmax_as(const Dim4& v, const Dim4& r) {
    check_reduced_dimensions("softmax_as", v, r);
    return v;
  }

  // matrix.rs:86-126
  Dim4 transpose(const Dim4& v, bool /*conjugate*/) {
    if (!(v[2] == 1 && v[3] == 1)) throw Error(ErrKind::Dimensions, "transpose", debug(v));
    return Dim4(v[1], v[0], 1, 1);
  }
  Dim4 matmul(const Dim4& v1, const Dim4& v2, MatProp p1, MatProp p2) {
    Dim4 tv1 = p1.transposed ? transpose(v1, false) : v1;
    Dim4 tv2 = p2.transposed ? transpose(v2, false) : v2;
    if (tv1[1] != tv2[0]) throw Error(ErrKind::Dimensions, "matmul", debug(v1) + " " + debug(v2));
    uint64_t a, b;
    if (tv1[2] == 1 && tv1[3] == 1) {
      a = tv2[2]; b = tv2[3];
    } else if (tv2[2] == 1 && tv2[3] == 1) {
      a = tv1[2]; b = tv1[3];
    } else if (tv1[2] == tv2[2] && tv1[3] == tv2[3]) {
      a = tv1[2]; b = tv1[3];
    } else {
      throw Error(ErrKind::Dimensions, "matmul", debug(v1) + " " + debug(v2));
    }
    return Dim4(tv1[0], tv2[1], a, b);
  }
  Dim4 matmul_nn(const Dim4& a, const Dim4& b) { return matmul(a, b, {}, {}); }

  template <class V> const V& link(const V& v) { return v; }
};

// ---------------------------------------------------------------------------------------------
// Eval — lib.rs:523-526. Scalars: `impl ... for Eval` over T: Number; arrays: the
// `Eval for af::Array<T>` impls restated over HostArray<T> (Check first, then the primitive).
// ---------------------------------------------------------------------------------------------
namespace detail {
// compiler-rt __powi{s,d}f2 (what Rust's f32::powi / num Pow<i16> lowers to).
template <class T>
T powi(T a, int b) {
  const bool recip = b < 0;
  T r = 1;
  while (true) {
    if (b & 1) r *= a;
    b /= 2;
    if (b == 0) break;
    a *= a;
  }
  return recip ? T(1) / r : r;
}
template <class T>
T ipow(T base, unsigned exp) {  // core::num::int::pow
  T acc = 1;
  while (exp > 0) {
    if (exp & 1) acc *= base;
    exp >>= 1;
    if (exp) base *= base;
  }
  return acc;
}
}  // namespace detail

template <class T> concept Scalar = std::is_arithmetic_v<T>;
template <class T> concept FloatScalar = std::is_floating_point_v<T>;

struct Eval {
  Check check_;
  Check& check() { return check_; }

  // LinkedAlgebra for eval algebras: linked.rs:14-22.
  template <class D> const D& link(const Value<D>& v) { return v.data; }

  // ----- scalars: core.rs:95-112, arith.rs:105-130, const_arith.rs:81-104,
  //                analytic.rs:188-246, compare.rs:148-175
  template <Scalar T> T variable(T v) { return v; }
  template <Scalar T> T constant(T v) { return v; }
  template <Scalar T> T add(const T& a, const T& b) { return a + b; }
  template <class D>
  D add_all(const std::vector<const D*>& vs) {  // core.rs:27-37 default: left fold
    if (vs.empty()) throw Error(ErrKind::Empty, "add_all");
    D r = *vs[0];
    for (size_t i = 1; i < vs.size(); i++) r = add(r, *vs[i]);
    return r;
  }
  template <Scalar T> T zeros(const T&) { return T(0); }
  template <Scalar T> T ones(const T&) { return T(1); }
  template <Scalar T> T neg(const T& v) { return -v; }
  template <Scalar T> T sub(const T& a, const T& b) { return a - b; }
  template <Scalar T> T mul(const T& a, const T& b) { return a * b; }
  template <Scalar T, class C> T setc(const T&, C c) { return T(c); }
  template <Scalar T, class C> T addc(const T& v, C c) { return v + T(c); }
  template <Scalar T, class C> T mulc(const T& v, C c) { return v * T(c); }
  template <Scalar T, class C>
  T powc(const T& v, C c) {
    if constexpr (std::is_integral_v<T>) return detail::ipow<T>(v, unsigned(c));
    else if constexpr (std::is_integral_v<C>) return detail::powi<T>(v, int(c));
    else return std::pow(v, T(c));
  }
  template <FloatScalar T> T exp(const T& v) { return std::exp(v); }
  template <FloatScalar T> T log(const T& v) { return std::log(v); }
  template <FloatScalar T> T log1p(const T& v) { return std::log(T(1) + v); }  // analytic.rs:204
  template <FloatScalar T> T sin(const T& v) { return std::sin(v); }
  template <FloatScalar T> T cos(const T& v) { return std::cos(v); }
  template <FloatScalar T> T tanh(const T& v) { return std::tanh(v); }
  template <FloatScalar T> T sigmoid(const T& v) { return T(1) / (T(1) + std::exp(-v)); }
  template <FloatScalar T> T reciprocal(const T& v) { return T(1) / v; }
  template <FloatScalar T> T sqrt(const T& v) { return std::sqrt(v); }
  template <FloatScalar T> T div(const T& a, const T& b) { return a / b; }
  template <FloatScalar T> T pow(const T& a, const T& b) { return std::pow(a, b); }
  template <Scalar T> T min(const T& a, const T& b) { return a <= b ? a : b; }
  template <Scalar T> T max(const T& a, const T& b) { return a >= b ? a : b; }
  template <Scalar T>
  T select_argmax(const T& v0, const T& v1, const T* r0, const T* r1) {
    if (v0 >= v1) return r0 ? *r0 : T(0);
    return r1 ? *r1 : T(0);
  }
  // derived ops (compare.rs:17-55 defaults); Eval for scalars overrides min/max only.
  template <class V> V abs(const V& v) { V n = neg(v); return select_argmax(v, n, &v, &n); }
  template <class V> V sign(const V& v) {
    V n = neg(v); V one = ones(v); V mone = neg(one);
    return select_argmax(v, n, &one, &mone);
  }
  template <class V> V relu(const V& v) { V z = zeros(v); return max(z, v); }

  // ----- arrays (column-major HostArray<T>)
  template <class T> HostArray<T> variable(HostArray<T> a) { return a; }
  template <class T> HostArray<T> constant(HostArray<T> a) { return a; }
  template <class T> HostArray<T> add(const HostArray<T>& a, const HostArray<T>& b) {  // core.rs:179-182
    check_.add(a.dims(), b.dims());
    return ha::map2(a, b, [](T x, T y) { return x + y; });
  }
  // arith.rs:49-75
  template <class T> HostArray<T> zeros(const HostArray<T>& v) { return HostArray<T>(v.dims(), T(0)); }
  template <class T> HostArray<T> ones(const HostArray<T>& v) { return HostArray<T>(v.dims(), T(1)); }
  template <class T> HostArray<T> neg(const HostArray<T>& v) {  // constant(0) - v  (arith.rs:60-62)
    return ha::map1(v, [](T x) { return T(0) - x; });
  }
  template <class T> HostArray<T> sub(const HostArray<T>& a, const HostArray<T>& b) {
    check_.sub(a.dims(), b.dims());
    return ha::map2(a, b, [](T x, T y) { return x - y; });
  }
  template <class T> HostArray<T> mul(const HostArray<T>& a, const HostArray<T>& b) {
    check_.mul(a.dims(), b.dims());
    return ha::map2(a, b, [](T x, T y) { return x * y; });
  }
  // const_arith.rs:41-55: constant array then binary op.
  template <class T, class C> HostArray<T> setc(const HostArray<T>& v, C c) { return HostArray<T>(v.dims(), T(c)); }
  template <class T, class C> HostArray<T> addc(const HostArray<T>& v, C c) {
    T k = T(c); return ha::map1(v, [k](T x) { return x + k; });
  }
  template <class T, class C> HostArray<T> mulc(const HostArray<T>& v, C c) {
    T k = T(c); return ha::map1(v, [k](T x) { return x * k; });
  }
  template <class T, class C> HostArray<T> powc(const HostArray<T>& v, C c) {
    T k = T(c); return ha::map1(v, [k](T x) { return std::pow(x, k); });
  }
  // analytic.rs:74-127
  template <class T> HostArray<T> exp(const HostArray<T>& v) { return ha::map1(v, [](T x) { return std::exp(x); }); }
  template <class T> HostArray<T> log(const HostArray<T>& v) { return ha::map1(v, [](T x) { return std::log(x); }); }
  template <class T> HostArray<T> log1p(const HostArray<T>& v) { return ha::map1(v, [](T x) { return std::log1p(x); }); }
  template <class T> HostArray<T> sin(const HostArray<T>& v) { return ha::map1(v, [](T x) { return std::sin(x); }); }
  template <class T> HostArray<T> cos(const HostArray<T>& v) { return ha::map1(v, [](T x) { return std::cos(x); }); }
  template <class T> HostArray<T> tanh(const HostArray<T>& v) { return ha::map1(v, [](T x) { return std::tanh(x); }); }
  template <class T> HostArray<T> sigmoid(const HostArray<T>& v) {
    return ha::map1(v, [](T x) { return T(1) / (T(1) + std::exp(-x)); });
  }
  template <class T> HostArray<T> reciprocal(const HostArray<T>& v) { return ha::map1(v, [](T x) { return T(1) / x; }); }
  template <class T> HostArray<T> sqrt(const HostArray<T>& v) { return ha::map1(v, [](T x) { return std::sqrt(x); }); }
  template <class T> HostArray<T> div(const HostArray<T>& a, const HostArray<T>& b) {
    check_.div(a.dims(), b.dims());
    return ha::map2(a, b, [](T x, T y) { return x / y; });
  }
  template <class T> HostArray<T> pow(const HostArray<T>& a, const HostArray<T>& b) {
    check_.pow(a.dims(), b.dims());
    return ha::map2(a, b, [](T x, T y) { return std::pow(x, y); });
  }
  // compare.rs:82-115
  template <class T> HostArray<T> min(const HostArray<T>& a, const HostArray<T>& b) {
    check_.min(a.dims(), b.dims());
    return ha::map2(a, b, [](T x, T y) { return std::fmin(x, y); });
  }
  template <class T> HostArray<T> max(const HostArray<T>& a, const HostArray<T>& b) {
    check_.max(a.dims(), b.dims());
    return ha::map2(a, b, [](T x, T y) { return std::fmax(x, y); });
  }
  template <class T>
  HostArray<T> select_argmax(const HostArray<T>& v0, const HostArray<T>& v1, const HostArray<T>* r0,
                             const HostArray<T>* r1) {
    Dim4 d0 = r0 ? r0->dims() : Dim4(), d1 = r1 ? r1->dims() : Dim4();
    check_.select_argmax(v0.dims(), v1.dims(), r0 ? &d0 : nullptr, r1 ? &d1 : nullptr);
    HostArray<T> r(v0.dims());
    for (size_t i = 0; i < r.size(); i++) {
      bool ge = v0[i] >= v1[i];
      r[i] = ge ? (r0 ? (*r0)[i] : T(0)) : (r1 ? (*r1)[i] : T(0));
    }
    return r;
  }
  // array.rs:57-127
  template <class T> HostArray<T> flat(const HostArray<T>& v) {
    HostArray<T> r = v; r.dims_ = Dim4(v.dims().elements()); return r;
  }
  template <class T> HostArray<T> moddims(const HostArray<T>& v, const Dim4& d) {
    check_.moddims(v.dims(), d);
    HostArray<T> r = v; r.dims_ = d; return r;
  }
  template <class T> HostArray<T> tile_as(const HostArray<T>& v, const Dim4& rd) {
    check_.tile_as(v.dims(), rd);
    return ha::tile(v, rd);
  }
  template <class T> HostArray<T> sum_as(const HostArray<T>& v, const Dim4& rd) {
    check_.sum_as(v.dims(), rd);
    HostArray<T> r = v;
    for (int i = 0; i < 4; i++) {
      if (rd[i] == v.dims()[i]) continue;
      r = ha::sum_axis<T>(r, i);
    }
    return r;
  }
  template <FloatScalar T> HostArray<T> constant_as(const T& v, const Dim4& d) { return HostArray<T>(d, v); }
  template <class T> T as_scalar(const HostArray<T>& v) {
    check_.as_scalar(v.dims());
    return v[0];
  }
  template <FloatScalar T> HostArray<T> scale(const T& lambda, const HostArray<T>& v) {
    return ha::map1(v, [lambda](T x) { return x * lambda; });
  }
  template <class T> T dot(const HostArray<T>& a, const HostArray<T>& b) {
    check_.dot(a.dims(), b.dims());
    return ha::dot<T>(a, b);
  }
  template <class T> T norm2(const HostArray<T>& v) { return dot(v, v); }  // array.rs:42-44
  // array_compare.rs:38-82
  template <class T> HostArray<T> max_as(const HostArray<T>& v, const Dim4& rd) {
    check_.max_as(v.dims(), rd);
    HostArray<T> r = v;
    for (int i = 0; i < 4; i++) {
      if (rd[i] == v.dims()[i]) continue;
      r = ha::max_axis<T>(r, i);
    }
    return r;
  }
  template <class T> HostArray<T> argmax_as(const HostArray<T>& v, const Dim4& rd) {
    Dim4 d = v.dims();
    HostArray<T> mx = tile_as(max_as(v, rd), d);
    HostArray<T> one(d, T(1));
    HostArray<T> arg = select_argmax(v, mx, &one, (const HostArray<T>*)nullptr);
    HostArray<T> sum = tile_as(sum_as(arg, rd), d);
    return div(arg, sum);
  }
  template <class T> HostArray<T> softmax_as(const HostArray<T>& v, const Dim4& rd) {
    Dim4 d = v.dims();
    HostArray<T> mx = tile_as(max_as(v, rd), d);
    HostArray<T> e = exp(sub(v, mx));
    HostArray<T> sum = tile_as(sum_as(e, rd), d);
    return div(e, sum);
  }
  // matrix.rs:44-61
  template <class T> HostArray<T> matmul(const HostArray<T>& a, const HostArray<T>& b, MatProp p1, MatProp p2) {
    Dim4 rd = check_.matmul(a.dims(), b.dims(), p1, p2);
    return ha::matmul<T>(a, b, p1.transposed, p2.transposed, rd);
  }
  template <class T> HostArray<T> transpose(const HostArray<T>& v, bool conj) {
    check_.transpose(v.dims(), conj);
    return ha::transpose(v);
  }
  template <class T> HostArray<T> matmul_nn(const HostArray<T>& a, const HostArray<T>& b) { return matmul(a, b, {}, {}); }
};

// ---------------------------------------------------------------------------------------------
// Gradient stores — store.rs:27-140.
// ---------------------------------------------------------------------------------------------
template <bool kHoldsValues>
struct GenericGradientMap {
  std::map<Id, std::any> values;

  template <class G>
  void insert(Id id, G g) { values[id] = std::any(std::move(g)); }

  // Typed access; a type mismatch is the reference's `expect("indices should have a unique type")`.
  template <class G>
  G* get_mut(Id id) {
    auto it = values.find(id);
    if (it == values.end()) return nullptr;
    G* p = std::any_cast<G>(&it->second);
    if (!p) throw std::logic_error("indices should have a unique type");
    return p;
  }
  template <class G>
  const G* get(Id id) const { return const_cast<GenericGradientMap*>(this)->template get_mut<G>(id); }

  // store.rs:44-55
  template <class G, class Alg>
  void add_gradient(Alg& graph, Id id, const G& value) {
    if (G* cur = get_mut<G>(id)) {
      *cur = graph.add(*cur, value);
    } else {
      insert<G>(id, value);
    }
  }
  bool contains(Id id) const { return values.count(id) != 0; }
};
using GenericGradientMap1 = GenericGradientMap<false>;
using GenericGradientMapN = GenericGradientMap<true>;

// ---------------------------------------------------------------------------------------------
// Graph — graph.rs.
// ---------------------------------------------------------------------------------------------
// see Graph::matmul: false (default) = the reference's matmul VJP as written, quirk included
inline bool& corrected_matmul_adjoint() {
  static bool on = false;
  return on;
}

template <class E>
struct Config1 {  // graph.rs:250-256
  using EvalAlgebra = E;
  template <class Self> using GradientAlgebra = E;
  using Store = GenericGradientMap1;
  template <class D> using GradValue = D;
  static constexpr bool kHigherOrder = false;
};
template <class E>
struct ConfigN {  // graph.rs:294-300
  using EvalAlgebra = E;
  template <class Self> using GradientAlgebra = Self;
  using Store = GenericGradientMapN;
  template <class D> using GradValue = Value<D>;
  static constexpr bool kHigherOrder = true;
};

template <class Cfg>
class Graph {
 public:
  using E = typename Cfg::EvalAlgebra;
  using GA = typename Cfg::template GradientAlgebra<Graph>;
  using Store = typename Cfg::Store;
  template <class D> using G = typename Cfg::template GradValue<D>;
  using UpdateFn = std::function<void(GA&, Store&, Id)>;

  struct Node {  // graph.rs:46-62
    std::vector<std::optional<Id>> inputs;
    std::shared_ptr<UpdateFn> update;
    void clear() { inputs.clear(); update.reset(); }
  };

  Graph() : arena_id_(arena_counter().fetch_add(1)) {}
  // Clone keeps arena_id so ids stay valid (graph.rs:353-360; relied on at :316).
  Graph(const Graph&) = default;
  Graph& operator=(const Graph&) = default;

  E& eval() { return eval_; }
  size_t num_nodes() const { return nodes_.size(); }

  // graph.rs:94-101
  template <class D>
  Value<D> make_variable(D data) {
    nodes_.push_back(Node{});
    return Value<D>{std::move(data), Id{arena_id_, uint32_t(nodes_.size())}};
  }

  // graph.rs:106-124
  template <class D, class F>
  Value<D> make_node(D data, std::vector<std::optional<Id>> inputs, F f) {
    return make_generic_node<D, G<D>, F>(std::move(data), std::move(inputs), std::move(f));
  }

  // graph.rs:127-165
  template <class D, class GD, class F>
  Value<D> make_generic_node(D data, std::vector<std::optional<Id>> inputs, F f) {
    bool all_none = true;
    for (auto& i : inputs) all_none = all_none && !i.has_value();
    if (all_none) return Value<D>::constant(std::move(data));
    auto d = dims(data);
    auto update = std::make_shared<UpdateFn>([d, f = std::move(f)](GA& algebra, Store& store, Id index) {
      const GD* p = store.template get<GD>(index);
      if (!p) throw Error(ErrKind::MissingGradient, "make_generic_node");
      GD value = *p;  // clone out of the store (graph.rs:152-155)
      auto vd = dims(value);
      check_equal_dimensions<decltype(d)>("make_generic_node", {&vd, &d});
      f(algebra, store, value);
    });
    nodes_.push_back(Node{std::move(inputs), std::move(update)});
    return Value<D>{std::move(data), Id{arena_id_, uint32_t(nodes_.size())}};
  }

  // ----- reverse sweep: graph.rs:172-246
  template <class GD>
  Store do_compute_gradients(GA& graph, Id gid, GD gradient, bool once) {
    Store store;
    store.template insert<GD>(gid, std::move(gradient));
    std::priority_queue<Id> heap;
    heap.push(gid);
    Id guard = gid.next_id();
    while (!heap.empty()) {
      Id id = heap.top();
      heap.pop();
      if (id < guard) {
        guard = id;
        if (id.arena != arena_id_ || id.index == 0 || id.index > nodes_.size())
          throw Error(ErrKind::MissingNode, "do_compute_gradients");
        Node& node = nodes_[id.index - 1];
        if (node.update) {
          auto keep = node.update;  // Arc clone: the closure may outlive node.clear()
          (*keep)(graph, store, id);
        }
        for (auto& in : node.inputs)
          if (in) heap.push(*in);
        if (once) node.clear();
      }
    }
    return store;
  }

  // Config1: graph.rs:263-290. Gradients are pure data; the gradient algebra is a clone of Eval.
  template <class T>
    requires(!Cfg::kHigherOrder)
  Store evaluate_gradients(Id id, T gradient) {
    E eval = eval_;
    return do_compute_gradients<T>(eval, id, std::move(gradient), false);
  }
  template <class T>
    requires(!Cfg::kHigherOrder)
  Store evaluate_gradients_once(Id id, T gradient) {
    E eval = eval_;
    return do_compute_gradients<T>(eval, id, std::move(gradient), true);
  }
  // ConfigN: graph.rs:307-318. Walk (and consume) a clone; record VJP ops into *this.
  template <class D>
    requires(Cfg::kHigherOrder)
  Store compute_gradients(Id id, Value<D> gradient) {
    Graph current = *this;
    return current.template do_compute_gradients<Value<D>>(*this, id, std::move(gradient), true);
  }

  // LinkedAlgebra for graphs: identity (linked.rs:25-30).
  template <class V> const V& link(const V& v) { return v; }

  // ----- CoreAlgebra: core.rs:203-262
  template <class D> Value<D> variable(D data) { return make_variable<D>(std::move(data)); }
  template <class D> Value<D> constant(D data) { return Value<D>::constant(std::move(data)); }

  template <class D>
  Value<D> add(const Value<D>& v1, const Value<D>& v2) {
    D result = eval_.add(v1.data, v2.data);
    auto id1 = v1.id, id2 = v2.id;
    return make_node<D>(std::move(result), {v1.input(), v2.input()},
                        [id1, id2](GA& graph, Store& store, const G<D>& gradient) {
                          if (id1) store.template add_gradient<G<D>>(graph, *id1, gradient);
                          if (id2) store.template add_gradient<G<D>>(graph, *id2, gradient);
                        });
  }
  template <class D>
  Value<D> add_all(const std::vector<const Value<D>*>& values) {
    std::vector<const D*> datas;
    std::vector<std::optional<Id>> inputs;
    for (auto v : values) { datas.push_back(&v->data); inputs.push_back(v->input()); }
    D result = eval_.template add_all<D>(datas);
    auto ids = inputs;
    return make_node<D>(std::move(result), std::move(inputs),
                        [ids](GA& graph, Store& store, const G<D>& gradient) {
                          for (auto& id : ids)
                            if (id) store.template add_gradient<G<D>>(graph, *id, gradient);
                        });
  }

  // ----- ArithAlgebra: arith.rs:153-234
  template <class D> Value<D> zeros(const Value<D>& v) { return constant<D>(eval_.zeros(v.data)); }
  template <class D> Value<D> ones(const Value<D>& v) { return constant<D>(eval_.ones(v.data)); }
  template <class D>
  Value<D> neg(const Value<D>& v) {
    D result = eval_.neg(v.data);
    auto id = v.id;
    return make_node<D>(std::move(result), {v.input()}, [id](GA& graph, Store& store, const G<D>& gradient) {
      if (id) {
        auto n = graph.neg(gradient);
        store.template add_gradient<G<D>>(graph, *id, n);
      }
    });
  }
  template <class D>
  Value<D> sub(const Value<D>& v0, const Value<D>& v1) {
    D result = eval_.sub(v0.data, v1.data);
    auto id0 = v0.id, id1 = v1.id;
    return make_node<D>(std::move(result), {v0.input(), v1.input()},
                        [id0, id1](GA& graph, Store& store, const G<D>& gradient) {
                          if (id0) store.template add_gradient<G<D>>(graph, *id0, gradient);
                          if (id1) {
                            auto n = graph.neg(gradient);
                            store.template add_gradient<G<D>>(graph, *id1, n);
                          }
                        });
  }
  template <class D>
  Value<D> mul(const Value<D>& v0, const Value<D>& v1) {
    D result = eval_.mul(v0.data, v1.data);
    return make_node<D>(std::move(result), {v0.input(), v1.input()},
                        [v0, v1](GA& graph, Store& store, const G<D>& gradient) {
                          if (v0.id) {
                            const auto& c1 = graph.link(v1);
                            auto grad = graph.mul(gradient, c1);
                            store.template add_gradient<G<D>>(graph, *v0.id, grad);
                          }
                          if (v1.id) {
                            const auto& c0 = graph.link(v0);
                            auto grad = graph.mul(c0, gradient);
                            store.template add_gradient<G<D>>(graph, *v1.id, grad);
                          }
                        });
  }

  // ----- ConstArithAlgebra: const_arith.rs:120-187
  template <class D, class C>
  Value<D> setc(const Value<D>& v, C c) {  // sic: evaluates addc (const_arith.rs:134-137)
    return constant<D>(eval_.addc(v.data, c));
  }
  template <class D, class C>
  Value<D> addc(const Value<D>& v, C c) {
    D result = eval_.addc(v.data, c);
    auto id = v.id;
    return make_node<D>(std::move(result), {v.input()}, [id](GA& graph, Store& store, const G<D>& gradient) {
      if (id) store.template add_gradient<G<D>>(graph, *id, gradient);
    });
  }
  template <class D, class C>
  Value<D> mulc(const Value<D>& v, C c) {
    D result = eval_.mulc(v.data, c);
    auto id = v.id;
    return make_node<D>(std::move(result), {v.input()}, [id, c](GA& graph, Store& store, const G<D>& gradient) {
      if (id) {
        auto grad = graph.mulc(gradient, c);
        store.template add_gradient<G<D>>(graph, *id, grad);
      }
    });
  }
  template <class D, class C>
  Value<D> powc(const Value<D>& v, C c) {
    D result = eval_.powc(v.data, c);
    return make_node<D>(std::move(result), {v.input()}, [v, c](GA& graph, Store& store, const G<D>& gradient) {
      if (v.id) {
        const auto& lv = graph.link(v);
        auto e = graph.powc(lv, C(c - C(1)));
        auto f = graph.mulc(e, c);
        auto grad = graph.mul(f, gradient);
        store.template add_gradient<G<D>>(graph, *v.id, grad);
      }
    });
  }

  // ----- AnalyticAlgebra: analytic.rs:287-484
  template <class D>
  Value<D> exp(const Value<D>& v) {
    D result = eval_.exp(v.data);
    return make_node<D>(std::move(result), {v.input()}, [v](GA& graph, Store& store, const G<D>& gradient) {
      if (v.id) {
        const auto& lv = graph.link(v);
        auto k = graph.exp(lv);
        auto grad = graph.mul(gradient, k);
        store.template add_gradient<G<D>>(graph, *v.id, grad);
      }
    });
  }
  template <class D>
  Value<D> log(const Value<D>& v) {
    D result = eval_.log(v.data);
    return make_node<D>(std::move(result), {v.input()}, [v](GA& graph, Store& store, const G<D>& gradient) {
      if (v.id) {
        const auto& lv = graph.link(v);
        auto grad = graph.div(gradient, lv);
        store.template add_gradient<G<D>>(graph, *v.id, grad);
      }
    });
  }
  template <class D>
  Value<D> log1p(const Value<D>& v) {
    D result = eval_.log1p(v.data);
    return make_node<D>(std::move(result), {v.input()}, [v](GA& graph, Store& store, const G<D>& gradient) {
      if (v.id) {
        const auto& lv = graph.link(v);
        auto v1p = graph.addc(lv, int16_t(1));
        auto grad = graph.div(gradient, v1p);
        store.template add_gradient<G<D>>(graph, *v.id, grad);
      }
    });
  }
  template <class D>
  Value<D> sin(const Value<D>& v) {
    D result = eval_.sin(v.data);
    return make_node<D>(std::move(result), {v.input()}, [v](GA& graph, Store& store, const G<D>& gradient) {
      if (v.id) {
        const auto& lv = graph.link(v);
        auto k = graph.cos(lv);
        auto grad = graph.mul(gradient, k);
        store.template add_gradient<G<D>>(graph, *v.id, grad);
      }
    });
  }
  template <class D>
  Value<D> cos(const Value<D>& v) {
    D result = eval_.cos(v.data);
    return make_node<D>(std::move(result), {v.input()}, [v](GA& graph, Store& store, const G<D>& gradient) {
      if (v.id) {
        const auto& lv = graph.link(v);
        auto c = graph.sin(lv);
        auto k = graph.neg(c);
        auto grad = graph.mul(gradient, k);
        store.template add_gradient<G<D>>(graph, *v.id, grad);
      }
    });
  }
  template <class D>
  Value<D> tanh(const Value<D>& v) {
    D result = eval_.tanh(v.data);
    return make_node<D>(std::move(result), {v.input()}, [v](GA& graph, Store& store, const G<D>& gradient) {
      if (v.id) {
        const auto& lv = graph.link(v);
        auto t = graph.tanh(lv);
        auto c = graph.mul(t, t);
        auto c2 = graph.neg(c);
        auto k = graph.addc(c2, int16_t(1));
        auto grad = graph.mul(gradient, k);
        store.template add_gradient<G<D>>(graph, *v.id, grad);
      }
    });
  }
  template <class D>
  Value<D> sigmoid(const Value<D>& v) {
    D result = eval_.sigmoid(v.data);
    return make_node<D>(std::move(result), {v.input()}, [v](GA& graph, Store& store, const G<D>& gradient) {
      if (v.id) {
        const auto& lv = graph.link(v);
        auto c = graph.sigmoid(lv);
        auto d = graph.neg(c);
        auto d2 = graph.addc(d, int16_t(1));
        auto k = graph.mul(c, d2);
        auto grad = graph.mul(gradient, k);
        store.template add_gradient<G<D>>(graph, *v.id, grad);
      }
    });
  }
  template <class D>
  Value<D> reciprocal(const Value<D>& v) {
    D result = eval_.reciprocal(v.data);
    return make_node<D>(std::move(result), {v.input()}, [v](GA& graph, Store& store, const G<D>& gradient) {
      if (v.id) {
        const auto& lv = graph.link(v);
        auto c = graph.mul(lv, lv);
        auto c2 = graph.neg(c);
        auto k = graph.reciprocal(c2);
        auto grad = graph.mul(gradient, k);
        store.template add_gradient<G<D>>(graph, *v.id, grad);
      }
    });
  }
  template <class D>
  Value<D> sqrt(const Value<D>& v) {
    D result = eval_.sqrt(v.data);
    return make_node<D>(std::move(result), {v.input()}, [v](GA& graph, Store& store, const G<D>& gradient) {
      if (v.id) {
        const auto& lv = graph.link(v);
        auto c = graph.sqrt(lv);
        auto c2 = graph.mulc(c, int16_t(2));
        auto k = graph.reciprocal(c2);
        auto grad = graph.mul(gradient, k);
        store.template add_gradient<G<D>>(graph, *v.id, grad);
      }
    });
  }
  template <class D>
  Value<D> div(const Value<D>& v0, const Value<D>& v1) {
    D result = eval_.div(v0.data, v1.data);
    return make_node<D>(std::move(result), {v0.input(), v1.input()},
                        [v0, v1](GA& graph, Store& store, const G<D>& gradient) {
                          const auto& c1 = graph.link(v1);
                          auto r1 = graph.reciprocal(c1);
                          auto g0 = graph.mul(gradient, r1);
                          if (v0.id) store.template add_gradient<G<D>>(graph, *v0.id, g0);
                          if (v1.id) {
                            const auto& c0 = graph.link(v0);
                            auto c = graph.mul(g0, r1);
                            auto c2 = graph.mul(c, c0);
                            auto g1 = graph.neg(c2);
                            store.template add_gradient<G<D>>(graph, *v1.id, g1);
                          }
                        });
  }
  // analytic.rs:48-55: Graph does not override pow → exp(p * log(v)) as three nodes.
  template <class D>
  Value<D> pow(const Value<D>& v, const Value<D>& p) {
    auto l = log(v);
    auto e = mul(p, l);
    return exp(e);
  }

  // ----- CompareAlgebra: compare.rs:190-251 (+ defaults :17-55)
  template <class D>
  Value<D> select_argmax(const Value<D>& v0, const Value<D>& v1, const Value<D>* r0, const Value<D>* r1) {
    D result = eval_.select_argmax(v0.data, v1.data, r0 ? &r0->data : nullptr, r1 ? &r1->data : nullptr);
    std::vector<std::optional<Id>> inputs;
    if (r0) inputs.push_back(r0->input());
    if (r1) inputs.push_back(r1->input());
    std::optional<Id> id0 = r0 ? r0->id : std::nullopt, id1 = r1 ? r1->id : std::nullopt;
    return make_node<D>(std::move(result), std::move(inputs),
                        [v0, v1, id0, id1](GA& graph, Store& store, const G<D>& gradient) {
                          const auto& c0 = graph.link(v0);
                          const auto& c1 = graph.link(v1);
                          using GV = std::remove_cvref_t<decltype(c0)>;
                          if (id0) {
                            auto grad = graph.select_argmax(c0, c1, (const GV*)&gradient, (const GV*)nullptr);
                            store.template add_gradient<G<D>>(graph, *id0, grad);
                          }
                          if (id1) {
                            auto grad = graph.select_argmax(c0, c1, (const GV*)nullptr, (const GV*)&gradient);
                            store.template add_gradient<G<D>>(graph, *id1, grad);
                          }
                        });
  }
  template <class D> Value<D> min(const Value<D>& v0, const Value<D>& v1) { return select_argmax(v0, v1, &v1, &v0); }
  template <class D> Value<D> max(const Value<D>& v0, const Value<D>& v1) { return select_argmax(v0, v1, &v0, &v1); }
  template <class D> Value<D> abs(const Value<D>& v) { auto n = neg(v); return select_argmax(v, n, &v, &n); }
  template <class D> Value<D> sign(const Value<D>& v) {
    auto n = neg(v); auto one = ones(v); auto mone = neg(one);
    return select_argmax(v, n, &one, &mone);
  }
  template <class D> Value<D> relu(const Value<D>& v) { auto z = zeros(v); return max(z, v); }

  // ----- ArrayAlgebra: array.rs:196-357 (D = HostArray<T>, Scalar = Value<T>, Dims = Dim4)
  template <class D>
  Value<D> flat(const Value<D>& v) {
    D result = eval_.flat(v.data);
    Dim4 vdims = dims(v.data);
    auto id = v.id;
    return make_node<D>(std::move(result), {v.input()}, [id, vdims](GA& graph, Store& store, const G<D>& gradient) {
      if (id) {
        auto x = graph.moddims(gradient, vdims);
        store.template add_gradient<G<D>>(graph, *id, x);
      }
    });
  }
  template <class D>
  Value<D> moddims(const Value<D>& v, const Dim4& rdims) {
    D result = eval_.moddims(v.data, rdims);
    Dim4 vdims = dims(v.data);
    auto id = v.id;
    return make_node<D>(std::move(result), {v.input()}, [id, vdims](GA& graph, Store& store, const G<D>& gradient) {
      if (id) {
        auto x = graph.moddims(gradient, vdims);
        store.template add_gradient<G<D>>(graph, *id, x);
      }
    });
  }
  template <class D>
  Value<D> tile_as(const Value<D>& v, const Dim4& rdims) {
    D result = eval_.tile_as(v.data, rdims);
    Dim4 vdims = dims(v.data);
    auto id = v.id;
    return make_node<D>(std::move(result), {v.input()}, [id, vdims](GA& graph, Store& store, const G<D>& gradient) {
      if (id) {
        auto x = graph.sum_as(gradient, vdims);
        store.template add_gradient<G<D>>(graph, *id, x);
      }
    });
  }
  template <class D>
  Value<D> sum_as(const Value<D>& v, const Dim4& rdims) {
    D result = eval_.sum_as(v.data, rdims);
    Dim4 vdims = dims(v.data);
    auto id = v.id;
    return make_node<D>(std::move(result), {v.input()}, [id, vdims](GA& graph, Store& store, const G<D>& gradient) {
      if (id) {
        auto x = graph.tile_as(gradient, vdims);
        store.template add_gradient<G<D>>(graph, *id, x);
      }
    });
  }
  template <FloatScalar T>
  Value<HostArray<T>> constant_as(const Value<T>& v, const Dim4& d) {
    using D = HostArray<T>;
    D result = eval_.constant_as(v.data, d);
    auto id = v.id;
    return make_generic_node<D, G<D>>(std::move(result), {v.input()},
                                      [id](GA& graph, Store& store, const G<D>& gradient) {
                                        if (id) {
                                          auto x = graph.sum_as(gradient, Dim4());
                                          auto y = graph.as_scalar(x);
                                          store.template add_gradient<G<T>>(graph, *id, y);
                                        }
                                      });
  }
  template <class T>
  Value<T> as_scalar(const Value<HostArray<T>>& v) {
    using D = HostArray<T>;
    T result = eval_.as_scalar(v.data);
    Dim4 vdims = dims(v.data);
    auto id = v.id;
    return make_generic_node<T, G<T>>(std::move(result), {v.input()},
                                      [id, vdims](GA& graph, Store& store, const G<T>& gradient) {
                                        if (id) {
                                          auto x = graph.constant_as(gradient, vdims);
                                          store.template add_gradient<G<D>>(graph, *id, x);
                                        }
                                      });
  }
  template <FloatScalar T>
  Value<HostArray<T>> scale(const Value<T>& v1, const Value<HostArray<T>>& v2) {
    using D = HostArray<T>;
    D result = eval_.scale(v1.data, v2.data);
    return make_node<D>(std::move(result), {v1.input(), v2.input()},
                        [v1, v2](GA& graph, Store& store, const G<D>& gradient) {
                          if (v1.id) {
                            const auto& c2 = graph.link(v2);
                            auto grad = graph.dot(gradient, c2);
                            store.template add_gradient<G<T>>(graph, *v1.id, grad);
                          }
                          if (v2.id) {
                            const auto& c1 = graph.link(v1);
                            auto grad = graph.scale(c1, gradient);
                            store.template add_gradient<G<D>>(graph, *v2.id, grad);
                          }
                        });
  }
  template <class T>
  Value<T> dot(const Value<HostArray<T>>& v1, const Value<HostArray<T>>& v2) {
    using D = HostArray<T>;
    T result = eval_.dot(v1.data, v2.data);
    return make_generic_node<T, G<T>>(std::move(result), {v1.input(), v2.input()},
                                      [v1, v2](GA& graph, Store& store, const G<T>& gradient) {
                                        if (v1.id) {
                                          const auto& c2 = graph.link(v2);
                                          auto grad = graph.scale(gradient, c2);
                                          store.template add_gradient<G<D>>(graph, *v1.id, grad);
                                        }
                                        if (v2.id) {
                                          const auto& c1 = graph.link(v1);
                                          auto grad = graph.scale(gradient, c1);
                                          store.template add_gradient<G<D>>(graph, *v2.id, grad);
                                        }
                                      });
  }
  template <class T> Value<T> norm2(const Value<HostArray<T>>& v) { return dot(v, v); }

  // ----- MatrixAlgebra: matrix.rs:141-199
  // The reference's VJP (matrix.rs:164-176) is `dA = matmul(G, B, p1, p2^T)`, `dB = matmul(A, G, p1^T, p2)`: the true
  // adjoint only when the differentiated operand is not itself transposed; for a transposed operand it
  // is a Dimensions error (non-square) or the transposed adjoint (square). That is what this oracle reproduces BY DEFAULT
  // and what the golden fixtures pin. corrected_matmul_adjoint() switches to the true adjoint for transposed operands
  // (dA = op2(B) G^T, dB = G^T op1(A)) — NOT the reference's behaviour; it exists so the device library's default mode
  // (gadcu_ctx_set_strict_reference(ctx, 0), which GraphN needs to differentiate through recorded backward GEMMs: C5)
  // has something to be checked against on the CPU. Identical wherever the reference formula is valid.
  template <class D>
  Value<D> matmul(const Value<D>& v1, const Value<D>& v2, MatProp p1, MatProp p2) {
    D result = eval_.matmul(v1.data, v2.data, p1, p2);
    const bool fix1 = p1.transposed && corrected_matmul_adjoint(), fix2 = p2.transposed && corrected_matmul_adjoint();
    return make_node<D>(std::move(result), {v1.input(), v2.input()},
                        [v1, v2, p1, p2, fix1, fix2](GA& graph, Store& store, const G<D>& gradient) {
                          const MatProp T{true, false};
                          if (v1.id) {
                            const auto& c2 = graph.link(v2);
                            auto grad = fix1 ? graph.matmul(c2, gradient, p2, T) : graph.matmul(gradient, c2, p1, p2.transpose());
                            store.template add_gradient<G<D>>(graph, *v1.id, grad);
                          }
                          if (v2.id) {
                            const auto& c1 = graph.link(v1);
                            auto grad = fix2 ? graph.matmul(gradient, c1, T, p1) : graph.matmul(c1, gradient, p1.transpose(), p2);
                            store.template add_gradient<G<D>>(graph, *v2.id, grad);
                          }
                        });
  }
  template <class D> Value<D> matmul_nn(const Value<D>& a, const Value<D>& b) { return matmul(a, b, {}, {}); }
  template <class D>
  Value<D> transpose(const Value<D>& v, bool conjugate) {
    D result = eval_.transpose(v.data, conjugate);
    auto id = v.id;
    return make_node<D>(std::move(result), {v.input()}, [id, conjugate](GA& graph, Store& store, const G<D>& gradient) {
      if (id) {
        auto grad = graph.transpose(gradient, conjugate);
        store.template add_gradient<G<D>>(graph, *id, grad);
      }
    });
  }

  // ----- ArrayCompareAlgebra: array_compare.rs:104-174
  template <class D>
  Value<D> max_as(const Value<D>& v, const Dim4& rdims) {
    D result = eval_.max_as(v.data, rdims);
    return make_node<D>(std::move(result), {v.input()}, [v, rdims](GA& graph, Store& store, const G<D>& gradient) {
      if (v.id) {
        const auto& lv = graph.link(v);
        auto mask = graph.argmax_as(lv, rdims);
        auto tiled = graph.tile_as(gradient, dims(lv));
        auto grad = graph.mul(tiled, mask);
        store.template add_gradient<G<D>>(graph, *v.id, grad);
      }
    });
  }
  template <class D>
  Value<D> argmax_as(const Value<D>& v, const Dim4& rdims) {
    return constant<D>(eval_.argmax_as(v.data, rdims));
  }
  template <class D>
  Value<D> softmax_as(const Value<D>& v, const Dim4& rdims) {
    D result = eval_.softmax_as(v.data, rdims);
    Dim4 d = dims(v.data);
    return make_node<D>(std::move(result), {v.input()}, [v, rdims, d](GA& graph, Store& store, const G<D>& gradient) {
      if (v.id) {
        const auto& lv = graph.link(v);
        auto res = graph.softmax_as(lv, rdims);
        auto g1 = graph.mul(gradient, res);
        auto rg = graph.sum_as(g1, rdims);
        auto gt = graph.tile_as(rg, d);
        auto g2 = graph.mul(gt, res);
        auto grad = graph.sub(g1, g2);
        store.template add_gradient<G<D>>(graph, *v.id, grad);
      }
    });
  }

 private:
  std::vector<Node> nodes_;
  uint32_t arena_id_;
  E eval_;
};

using Graph1 = Graph<Config1<Eval>>;  // lib.rs:529
using GraphN = Graph<ConfigN<Eval>>;  // lib.rs:532

// Typed reads of a store, mirroring GradientReader (store.rs:27-29, 71-79, 108-127).
template <class T>
const T* read1(const GenericGradientMap1& s, Id id) { return s.get<T>(id); }
template <class T>
const Value<T>* readN(const GenericGradientMapN& s, Id id) { return s.get<Value<T>>(id); }

}  // namespace gad_oracle
