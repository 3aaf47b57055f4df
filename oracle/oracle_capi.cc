// ORACLE — TEST INFRASTRUCTURE ONLY (see host_array.hpp). A type-erased extern "C" surface over
// gad_oracle.hpp so tests/ (Python, ctypes) can run the SAME generic program on the oracle and on
// the CUDA backend and compare. Supports data types f32/f64, scalars and column-major arrays, and
// the four algebras Check / Eval / Graph1 / GraphN (gad/src/lib.rs:519-532).
#include <cstdio>
#include <variant>

#include "gad_oracle.hpp"
#include "ops.h"

using namespace gad_oracle;

namespace {

thread_local std::string g_last_error;

template <class T> using S = Value<T>;
template <class T> using A = Value<HostArray<T>>;

struct OVal {
  bool is_check = false;
  std::variant<S<float>, S<double>, A<float>, A<double>> v;
  std::variant<Unit, Dim4> c;
};

template <class X>
OVal* wrap(X x) {
  auto* o = new OVal;
  o->v = std::move(x);
  return o;
}
OVal* wrap_check(Unit u) { auto* o = new OVal; o->is_check = true; o->c = u; return o; }
OVal* wrap_check(Dim4 d) { auto* o = new OVal; o->is_check = true; o->c = d; return o; }

// Eval presented with the Value<D>-based signatures of the graph algebras (ids are always None).
struct EvalW {
  Eval e;
  template <class D> Value<D> variable(D d) { return Value<D>::constant(std::move(d)); }
  template <class D> Value<D> constant(D d) { return Value<D>::constant(std::move(d)); }
#define EW_U(op) \
  template <class D> Value<D> op(const Value<D>& v) { return Value<D>::constant(e.op(v.data)); }
#define EW_B(op) \
  template <class D> Value<D> op(const Value<D>& a, const Value<D>& b) { return Value<D>::constant(e.op(a.data, b.data)); }
#define EW_C(op) \
  template <class D, class C> Value<D> op(const Value<D>& v, C c) { return Value<D>::constant(e.op(v.data, c)); }
#define EW_D(op) \
  template <class D> Value<D> op(const Value<D>& v, const Dim4& d) { return Value<D>::constant(e.op(v.data, d)); }
  EW_B(add) EW_B(sub) EW_B(mul) EW_U(zeros) EW_U(ones) EW_U(neg)
  EW_C(setc) EW_C(addc) EW_C(mulc) EW_C(powc)
  EW_U(exp) EW_U(log) EW_U(log1p) EW_U(sin) EW_U(cos) EW_U(tanh) EW_U(sigmoid) EW_U(reciprocal) EW_U(sqrt)
  EW_B(div) EW_B(pow) EW_B(min) EW_B(max) EW_U(abs) EW_U(sign) EW_U(relu)
  EW_U(flat) EW_D(moddims) EW_D(tile_as) EW_D(sum_as) EW_D(max_as) EW_D(argmax_as) EW_D(softmax_as)
  EW_B(matmul_nn)
  template <class D> Value<D> add_all(const std::vector<const Value<D>*>& vs) {
    std::vector<const D*> ds;
    for (auto v : vs) ds.push_back(&v->data);
    return Value<D>::constant(e.template add_all<D>(ds));
  }
  template <class D>
  Value<D> select_argmax(const Value<D>& v0, const Value<D>& v1, const Value<D>* r0, const Value<D>* r1) {
    return Value<D>::constant(e.select_argmax(v0.data, v1.data, r0 ? &r0->data : nullptr, r1 ? &r1->data : nullptr));
  }
  template <class T> A<T> constant_as(const S<T>& v, const Dim4& d) { return A<T>::constant(e.constant_as(v.data, d)); }
  template <class T> S<T> as_scalar(const A<T>& v) { return S<T>::constant(e.as_scalar(v.data)); }
  template <class T> A<T> scale(const S<T>& l, const A<T>& v) { return A<T>::constant(e.scale(l.data, v.data)); }
  template <class T> S<T> dot(const A<T>& a, const A<T>& b) { return S<T>::constant(e.dot(a.data, b.data)); }
  template <class T> S<T> norm2(const A<T>& a) { return S<T>::constant(e.norm2(a.data)); }
  template <class D> Value<D> matmul(const Value<D>& a, const Value<D>& b, MatProp p1, MatProp p2) {
    return Value<D>::constant(e.matmul(a.data, b.data, p1, p2));
  }
  template <class D> Value<D> transpose(const Value<D>& a, bool c) { return Value<D>::constant(e.transpose(a.data, c)); }
};

struct OAlg {
  int kind;
  Check check;
  EvalW eval;
  Graph1 g1;
  GraphN gN;
};

struct OStore {
  int kind;  // OALG_GRAPH1 or OALG_GRAPHN
  GenericGradientMap1 s1;
  GenericGradientMapN sN;
};

// ----- typed op dispatch for Eval / Graph1 / GraphN
template <class T, class Alg>
OVal* op_typed(Alg& g, int op, OVal* const* in, int n, const double* sp, const uint64_t* dp) {
  auto is_arr = [&](int i) { return std::holds_alternative<A<T>>(in[i]->v); };
  auto arr = [&](int i) -> const A<T>& { return std::get<A<T>>(in[i]->v); };
  auto sc = [&](int i) -> const S<T>& { return std::get<S<T>>(in[i]->v); };
  Dim4 d = dp ? Dim4(dp[0], dp[1], dp[2], dp[3]) : Dim4();
#define BOTH1(name) return is_arr(0) ? wrap(g.name(arr(0))) : wrap(g.name(sc(0)))
#define BOTH2(name) return is_arr(0) ? wrap(g.name(arr(0), arr(1))) : wrap(g.name(sc(0), sc(1)))
#define BOTHC(name)                                                                        \
  do {                                                                                     \
    if (sp[1] != 0.0) {                                                                    \
      int16_t c = int16_t(sp[0]);                                                          \
      return is_arr(0) ? wrap(g.name(arr(0), c)) : wrap(g.name(sc(0), c));                 \
    } else {                                                                               \
      T c = T(sp[0]);                                                                      \
      return is_arr(0) ? wrap(g.name(arr(0), c)) : wrap(g.name(sc(0), c));                 \
    }                                                                                      \
  } while (0)
  switch (op) {
    case OOP_VARIABLE: return is_arr(0) ? wrap(g.variable(arr(0).data)) : wrap(g.variable(sc(0).data));
    case OOP_CONSTANT: return is_arr(0) ? wrap(g.constant(arr(0).data)) : wrap(g.constant(sc(0).data));
    case OOP_ADD: BOTH2(add);
    case OOP_ADD_ALL: {
      if (n == 0) throw Error(ErrKind::Empty, "add_all");
      if (is_arr(0)) {
        std::vector<const A<T>*> vs;
        for (int i = 0; i < n; i++) vs.push_back(&arr(i));
        return wrap(g.template add_all<HostArray<T>>(vs));
      }
      std::vector<const S<T>*> vs;
      for (int i = 0; i < n; i++) vs.push_back(&sc(i));
      return wrap(g.template add_all<T>(vs));
    }
    case OOP_SUB: BOTH2(sub);
    case OOP_MUL: BOTH2(mul);
    case OOP_ZEROS: BOTH1(zeros);
    case OOP_ONES: BOTH1(ones);
    case OOP_NEG: BOTH1(neg);
    case OOP_SETC: BOTHC(setc);
    case OOP_ADDC: BOTHC(addc);
    case OOP_MULC: BOTHC(mulc);
    case OOP_POWC: BOTHC(powc);
    case OOP_EXP: BOTH1(exp);
    case OOP_LOG: BOTH1(log);
    case OOP_LOG1P: BOTH1(log1p);
    case OOP_SIN: BOTH1(sin);
    case OOP_COS: BOTH1(cos);
    case OOP_TANH: BOTH1(tanh);
    case OOP_SIGMOID: BOTH1(sigmoid);
    case OOP_RECIPROCAL: BOTH1(reciprocal);
    case OOP_SQRT: BOTH1(sqrt);
    case OOP_DIV: BOTH2(div);
    case OOP_POW: BOTH2(pow);
    case OOP_SELECT_ARGMAX:
      if (is_arr(0))
        return wrap(g.select_argmax(arr(0), arr(1), in[2] ? &arr(2) : nullptr, in[3] ? &arr(3) : nullptr));
      return wrap(g.select_argmax(sc(0), sc(1), in[2] ? &sc(2) : nullptr, in[3] ? &sc(3) : nullptr));
    case OOP_MIN: BOTH2(min);
    case OOP_MAX: BOTH2(max);
    case OOP_ABS: BOTH1(abs);
    case OOP_SIGN: BOTH1(sign);
    case OOP_RELU: BOTH1(relu);
    case OOP_FLAT: return wrap(g.flat(arr(0)));
    case OOP_MODDIMS: return wrap(g.moddims(arr(0), d));
    case OOP_TILE_AS: return wrap(g.tile_as(arr(0), d));
    case OOP_SUM_AS: return wrap(g.sum_as(arr(0), d));
    case OOP_CONSTANT_AS: return wrap(g.constant_as(sc(0), d));
    case OOP_AS_SCALAR: return wrap(g.as_scalar(arr(0)));
    case OOP_SCALE: return wrap(g.scale(sc(0), arr(1)));
    case OOP_DOT: return wrap(g.dot(arr(0), arr(1)));
    case OOP_NORM2: return wrap(g.norm2(arr(0)));
    case OOP_MATMUL:
      return wrap(g.matmul(arr(0), arr(1), MatProp{sp[0] != 0, sp[1] != 0}, MatProp{sp[2] != 0, sp[3] != 0}));
    case OOP_TRANSPOSE: return wrap(g.transpose(arr(0), sp[0] != 0));
    case OOP_MATMUL_NN: return wrap(g.matmul_nn(arr(0), arr(1)));
    case OOP_MAX_AS: return wrap(g.max_as(arr(0), d));
    case OOP_ARGMAX_AS: return wrap(g.argmax_as(arr(0), d));
    case OOP_SOFTMAX_AS: return wrap(g.softmax_as(arr(0), d));
  }
  throw std::invalid_argument("unknown op");
}

// ----- Check algebra on Unit / Dim4
OVal* op_check(Check& c, int op, OVal* const* in, int n, const double* sp, const uint64_t* dp) {
  auto is_arr = [&](int i) { return std::holds_alternative<Dim4>(in[i]->c); };
  auto arr = [&](int i) -> const Dim4& { return std::get<Dim4>(in[i]->c); };
  Dim4 d = dp ? Dim4(dp[0], dp[1], dp[2], dp[3]) : Dim4();
#define CB1(name) return is_arr(0) ? wrap_check(c.name(arr(0))) : wrap_check(c.name(Unit{}))
#define CB2(name) return is_arr(0) ? wrap_check(c.name(arr(0), arr(1))) : wrap_check(c.name(Unit{}, Unit{}))
#define CBC(name) return is_arr(0) ? wrap_check(c.name(arr(0), 0)) : wrap_check(c.name(Unit{}, 0))
  switch (op) {
    case OOP_VARIABLE: case OOP_CONSTANT: return is_arr(0) ? wrap_check(arr(0)) : wrap_check(Unit{});
    case OOP_ADD: CB2(add);
    case OOP_ADD_ALL: {
      if (n == 0) throw Error(ErrKind::Empty, "add_all");
      if (!is_arr(0)) return wrap_check(Unit{});
      std::vector<const Dim4*> vs;
      for (int i = 0; i < n; i++) vs.push_back(&arr(i));
      return wrap_check(c.add_all(vs));
    }
    case OOP_SUB: CB2(sub);
    case OOP_MUL: CB2(mul);
    case OOP_ZEROS: CB1(zeros);
    case OOP_ONES: CB1(ones);
    case OOP_NEG: CB1(neg);
    case OOP_SETC: CBC(setc);
    case OOP_ADDC: CBC(addc);
    case OOP_MULC: CBC(mulc);
    case OOP_POWC: CBC(powc);
    case OOP_EXP: CB1(exp);
    case OOP_LOG: CB1(log);
    case OOP_LOG1P: CB1(log1p);
    case OOP_SIN: CB1(sin);
    case OOP_COS: CB1(cos);
    case OOP_TANH: CB1(tanh);
    case OOP_SIGMOID: CB1(sigmoid);
    case OOP_RECIPROCAL: CB1(reciprocal);
    case OOP_SQRT: CB1(sqrt);
    case OOP_DIV: CB2(div);
    case OOP_POW: CB2(pow);
    case OOP_SELECT_ARGMAX:
      if (!is_arr(0)) return wrap_check(Unit{});
      return wrap_check(c.select_argmax(arr(0), arr(1), in[2] ? &arr(2) : nullptr, in[3] ? &arr(3) : nullptr));
    case OOP_MIN: CB2(min);
    case OOP_MAX: CB2(max);
    // abs / sign / relu are derived from select_argmax on the operand's own dims (compare.rs:27-55)
    case OOP_ABS: case OOP_SIGN: case OOP_RELU:
      if (!is_arr(0)) return wrap_check(Unit{});
      return wrap_check(c.select_argmax(arr(0), arr(0), &arr(0), &arr(0)));
    case OOP_FLAT: return wrap_check(c.flat(arr(0)));
    case OOP_MODDIMS: return wrap_check(c.moddims(arr(0), d));
    case OOP_TILE_AS: return wrap_check(c.tile_as(arr(0), d));
    case OOP_SUM_AS: return wrap_check(c.sum_as(arr(0), d));
    case OOP_CONSTANT_AS: return wrap_check(c.constant_as(Unit{}, d));
    case OOP_AS_SCALAR: return wrap_check(c.as_scalar(arr(0)));
    case OOP_SCALE: return wrap_check(c.scale(Unit{}, arr(1)));
    case OOP_DOT: return wrap_check(c.dot(arr(0), arr(1)));
    case OOP_NORM2: return wrap_check(c.norm2(arr(0)));
    case OOP_MATMUL:
      return wrap_check(c.matmul(arr(0), arr(1), MatProp{sp[0] != 0, sp[1] != 0}, MatProp{sp[2] != 0, sp[3] != 0}));
    case OOP_TRANSPOSE: return wrap_check(c.transpose(arr(0), sp[0] != 0));
    case OOP_MATMUL_NN: return wrap_check(c.matmul_nn(arr(0), arr(1)));
    case OOP_MAX_AS: return wrap_check(c.max_as(arr(0), d));
    case OOP_ARGMAX_AS: return wrap_check(c.argmax_as(arr(0), d));
    case OOP_SOFTMAX_AS: return wrap_check(c.softmax_as(arr(0), d));
  }
  throw std::invalid_argument("unknown op");
}

template <class F>
int guarded(F f) {
  try {
    f();
    return 0;
  } catch (const Error& e) {
    g_last_error = e.what();
    return int(e.kind);
  } catch (const std::exception& e) {
    g_last_error = e.what();
    return 99;
  }
}

bool is_f32(const OVal* v) { return v->v.index() == 0 || v->v.index() == 2; }

}  // namespace

extern "C" {

const char* oracle_last_error() { return g_last_error.c_str(); }

void* oracle_alg_new(int kind) {
  auto* a = new OAlg;
  a->kind = kind;
  return a;
}
void* oracle_alg_clone(void* p) { return new OAlg(*static_cast<OAlg*>(p)); }
void oracle_alg_free(void* p) { delete static_cast<OAlg*>(p); }
size_t oracle_alg_num_nodes(void* p) {
  auto* a = static_cast<OAlg*>(p);
  return a->kind == OALG_GRAPH1 ? a->g1.num_nodes() : a->kind == OALG_GRAPHN ? a->gN.num_nodes() : 0;
}

void* oracle_scalar(int dtype, double v) {
  return dtype == 0 ? wrap(S<float>::constant(float(v))) : wrap(S<double>::constant(v));
}
void* oracle_array(int dtype, const uint64_t* d, const void* data) {
  Dim4 dims(d[0], d[1], d[2], d[3]);
  if (dtype == 0) return wrap(A<float>::constant(HostArray<float>(dims, static_cast<const float*>(data))));
  return wrap(A<double>::constant(HostArray<double>(dims, static_cast<const double*>(data))));
}
void* oracle_check_dims(int is_scalar, const uint64_t* d) {
  return is_scalar ? wrap_check(Unit{}) : wrap_check(Dim4(d[0], d[1], d[2], d[3]));
}
void oracle_val_free(void* p) { delete static_cast<OVal*>(p); }

// dtype: 0 f32, 1 f64; index 0 = constant (no id).
void oracle_val_info(void* p, int* dtype, int* is_scalar, uint64_t* dims, uint32_t* arena, uint32_t* index) {
  auto* v = static_cast<OVal*>(p);
  *arena = 0; *index = 0;
  for (int i = 0; i < 4; i++) dims[i] = 1;
  if (v->is_check) {
    *dtype = -1;
    *is_scalar = std::holds_alternative<Unit>(v->c);
    if (!*is_scalar) for (int i = 0; i < 4; i++) dims[i] = std::get<Dim4>(v->c)[i];
    return;
  }
  *dtype = is_f32(v) ? 0 : 1;
  *is_scalar = v->v.index() < 2;
  std::visit([&](auto& val) {
    if (val.id) { *arena = val.id->arena; *index = val.id->index; }
    if constexpr (!std::is_arithmetic_v<std::remove_cvref_t<decltype(val.data)>>)
      for (int i = 0; i < 4; i++) dims[i] = val.data.dims()[i];
  }, v->v);
}
void oracle_val_read(void* p, void* out) {
  auto* v = static_cast<OVal*>(p);
  std::visit([&](auto& val) {
    using D = std::remove_cvref_t<decltype(val.data)>;
    if constexpr (std::is_arithmetic_v<D>) std::memcpy(out, &val.data, sizeof(D));
    else std::memcpy(out, val.data.data(), val.data.size() * sizeof(val.data[0]));
  }, v->v);
}

int oracle_op(void* alg, int op, void* const* ins, int n_in, const double* sp, const uint64_t* dp, void** out) {
  auto* a = static_cast<OAlg*>(alg);
  auto* in = reinterpret_cast<OVal* const*>(ins);
  *out = nullptr;
  return guarded([&] {
    if (a->kind == OALG_CHECK) { *out = op_check(a->check, op, in, n_in, sp, dp); return; }
    if (n_in == 0) throw Error(ErrKind::Empty, "op");
    bool f32 = is_f32(in[0]);
    switch (a->kind) {
      case OALG_EVAL: *out = f32 ? op_typed<float>(a->eval, op, in, n_in, sp, dp) : op_typed<double>(a->eval, op, in, n_in, sp, dp); break;
      case OALG_GRAPH1: *out = f32 ? op_typed<float>(a->g1, op, in, n_in, sp, dp) : op_typed<double>(a->g1, op, in, n_in, sp, dp); break;
      case OALG_GRAPHN: *out = f32 ? op_typed<float>(a->gN, op, in, n_in, sp, dp) : op_typed<double>(a->gN, op, in, n_in, sp, dp); break;
      default: throw std::invalid_argument("bad algebra kind");
    }
  });
}

// mode: 0 = evaluate_gradients (Graph1), 1 = evaluate_gradients_once (Graph1),
//       2 = compute_gradients (GraphN; the seed may carry an id).
int oracle_gradients(void* alg, uint32_t arena, uint32_t index, void* seed, int mode, void** store) {
  auto* a = static_cast<OAlg*>(alg);
  auto* s = static_cast<OVal*>(seed);
  *store = nullptr;
  return guarded([&] {
    if (index == 0) throw Error(ErrKind::MissingId, "gradients");
    Id id{arena, index};
    auto st = std::make_unique<OStore>();
    st->kind = a->kind;
    std::visit([&](auto& val) {
      using D = std::remove_cvref_t<decltype(val.data)>;
      if (a->kind == OALG_GRAPH1) {
        st->s1 = mode == 1 ? a->g1.template evaluate_gradients_once<D>(id, val.data)
                           : a->g1.template evaluate_gradients<D>(id, val.data);
      } else if (a->kind == OALG_GRAPHN) {
        st->sN = a->gN.template compute_gradients<D>(id, val);
      } else {
        throw std::invalid_argument("gradients need a graph algebra");
      }
    }, s->v);
    *store = st.release();
  });
}

void oracle_store_free(void* p) { delete static_cast<OStore*>(p); }

// 0 (default): the reference's matmul VJP as written (matrix.rs:164-176, quirk for transposed operands included);
// 1: the true adjoint for transposed operands — the device library's default mode. Returns the previous setting.
int oracle_set_corrected_adjoint(int on) {
  const int before = corrected_matmul_adjoint() ? 1 : 0;
  corrected_matmul_adjoint() = on != 0;
  return before;
}

// Returns NULL when the id has no gradient in the store (absent ≠ zero: tests/compare.rs:90-91).
void* oracle_store_get(void* store, uint32_t arena, uint32_t index, int dtype, int is_scalar) {
  auto* st = static_cast<OStore*>(store);
  Id id{arena, index};
  OVal* out = nullptr;
  int rc = guarded([&] {
    auto fetch = [&](auto tag) {
      using D = decltype(tag);
      if (st->kind == OALG_GRAPH1) {
        if (const D* p = st->s1.get<D>(id)) out = wrap(Value<D>::constant(*p));
      } else {
        if (const Value<D>* p = st->sN.get<Value<D>>(id)) out = wrap(*p);
      }
    };
    if (dtype == 0 && is_scalar) fetch(float{});
    else if (dtype == 1 && is_scalar) fetch(double{});
    else if (dtype == 0) fetch(HostArray<float>{});
    else fetch(HostArray<double>{});
  });
  return rc == 0 ? out : nullptr;
}

// Number of entries, and their indices in ascending order.
size_t oracle_store_ids(void* store, uint32_t* out, size_t cap) {
  auto* st = static_cast<OStore*>(store);
  auto& m = st->kind == OALG_GRAPH1 ? st->s1.values : st->sN.values;
  size_t k = 0;
  for (auto& kv : m) {
    if (k < cap) out[k] = kv.first.index;
    k++;
  }
  return k;
}

}  // extern "C"
