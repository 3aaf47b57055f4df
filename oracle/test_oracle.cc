// ORACLE ACCEPTANCE SUITE — pins oracle/gad_oracle.hpp against every feature-free known-answer
// test the reference holds for the tape path. Each block cites the reference test
// it replays; expected values are the reference's own.
//   build+run:  make -C oracle test
#include <cstdio>
#include <cstdlib>
#include <string>

#include "gad_oracle.hpp"

using namespace gad_oracle;

static int g_failures = 0;
static int g_checks = 0;
#define CHECK(cond)                                                         \
  do {                                                                      \
    g_checks++;                                                             \
    if (!(cond)) {                                                          \
      g_failures++;                                                         \
      std::printf("FAIL %s:%d  %s\n", __FILE__, __LINE__, #cond);           \
    }                                                                       \
  } while (0)
#define CHECK_NEAR(x, y) CHECK(std::fabs(double(x) - double(y)) < 0.001)  // tests/analytic.rs:6-9

// ---------------------------------------------------------------------------------------------
// tests/symbolic.rs:9-108 — a string-building evaluation algebra (a fake backend).
// ---------------------------------------------------------------------------------------------
namespace gad_oracle {
struct Exp;
using Sym = std::shared_ptr<const Exp>;
struct Exp {
  enum Kind { Zero, One, Num, Neg, Add, Mul } kind;
  std::string num;
  Sym a, b;
};
inline Sym sym_num(std::string s) { return std::make_shared<const Exp>(Exp{Exp::Num, std::move(s), nullptr, nullptr}); }
inline Unit dims(const Sym&) { return {}; }
inline std::string show(const Sym& e) {  // Display impl, tests/symbolic.rs:76-88
  switch (e->kind) {
    case Exp::Zero: return "0";
    case Exp::One: return "1";
    case Exp::Num: return e->num;
    case Exp::Neg: return "(-" + show(e->a) + ")";
    case Exp::Add: return "(" + show(e->a) + "+" + show(e->b) + ")";
    case Exp::Mul: return show(e->a) + show(e->b);
  }
  return "?";
}
struct SymEval {
  template <class D> const D& link(const Value<D>& v) { return v.data; }
  Sym variable(Sym d) { return d; }
  Sym constant(Sym d) { return d; }
  Sym add(const Sym& a, const Sym& b) { return std::make_shared<const Exp>(Exp{Exp::Add, "", a, b}); }
  Sym zeros(const Sym&) { return std::make_shared<const Exp>(Exp{Exp::Zero, "", nullptr, nullptr}); }
  Sym ones(const Sym&) { return std::make_shared<const Exp>(Exp{Exp::One, "", nullptr, nullptr}); }
  Sym neg(const Sym& v) { return std::make_shared<const Exp>(Exp{Exp::Neg, "", v, nullptr}); }
  Sym sub(const Sym& a, const Sym& b) { return add(a, neg(b)); }
  Sym mul(const Sym& a, const Sym& b) { return std::make_shared<const Exp>(Exp{Exp::Mul, "", a, b}); }
};
using SymGraph1 = Graph<Config1<SymEval>>;
}  // namespace gad_oracle

static void test_symgraph1() {  // tests/symbolic.rs:95-108 and the doctest lib.rs:341-438
  SymGraph1 g;
  auto a = g.variable(sym_num("a"));
  auto b = g.variable(sym_num("b"));
  auto c = g.mul(a, b);
  auto d = g.mul(a, c);
  CHECK(show(d.data) == "aab");
  auto gradients = g.evaluate_gradients_once(d.gid(), sym_num("1"));
  CHECK(show(*gradients.get<Sym>(a.gid())) == "(1ab+a1b)");
  CHECK(show(*gradients.get<Sym>(b.gid())) == "aa1");
}

// ---------------------------------------------------------------------------------------------
// tests/graph.rs
// ---------------------------------------------------------------------------------------------
static void test_gradient_simple() {  // tests/graph.rs:8-26
  Graph1 g;
  auto a = g.variable<int32_t>(3);
  auto x = g.mul(a, a);
  auto c = g.add(a, x);
  CHECK(c.data == 3 * 3 + 3);
  Graph1 g_clone = g;
  auto g1 = g_clone.evaluate_gradients_once<int32_t>(c.gid(), 1);
  auto g2 = g.evaluate_gradients_once<int32_t>(c.gid(), 2);
  CHECK(*g1.get<int32_t>(a.gid()) == 7);
  CHECK(*g2.get<int32_t>(a.gid()) == 14);
}

template <class MakeZ>
static void hessian_and_more(MakeZ make_z) {  // tests/graph.rs:28-69 and :107-153
  GraphN g;
  auto x = g.variable<float>(1.0f);
  auto y = g.variable<float>(0.4f);
  Value<float> z = make_z(g, x, y);
  Id xi = x.gid(), yi = y.gid(), zi = z.gid();

  auto dz = g.constant<float>(1.0f);
  auto dz_d = g.compute_gradients(zi, dz);
  const Value<float>* dz_dx = dz_d.get<Value<float>>(xi);
  const Value<float>* dz_dy = dz_d.get<Value<float>>(yi);
  CHECK(dz_dx && dz_dy);

  auto ddz = g.constant<float>(1.0f);
  auto ddz_dxd = g.compute_gradients(dz_dx->gid(), ddz);
  auto ddz_dyd = g.compute_gradients(dz_dy->gid(), ddz);

  CHECK(ddz_dxd.get<Value<float>>(xi) == nullptr);  // None
  CHECK(ddz_dxd.get<Value<float>>(yi) && ddz_dxd.get<Value<float>>(yi)->data == 0.8f);  // 2y
  CHECK(ddz_dyd.get<Value<float>>(xi) && ddz_dyd.get<Value<float>>(xi)->data == 0.8f);
  CHECK(ddz_dyd.get<Value<float>>(yi) && ddz_dyd.get<Value<float>>(yi)->data == 2.0f);

  auto dddz = g.constant<float>(1.0f);
  auto dddz_dxdyd = g.compute_gradients(ddz_dxd.get<Value<float>>(yi)->gid(), dddz);
  CHECK(dddz_dxdyd.get<Value<float>>(xi) == nullptr);
  CHECK(dddz_dxdyd.get<Value<float>>(yi) && dddz_dxdyd.get<Value<float>>(yi)->data == 2.0f);
}

static void test_hessian_and_more() {
  hessian_and_more([](GraphN& g, const Value<float>& x, const Value<float>& y) {
    auto h = g.mul(x, y);
    return g.mul(h, y);
  });
}
static void test_hessian_and_more_arrays() {  // tests/graph.rs:107-153 ("using matrix operations for fun")
  hessian_and_more([](GraphN& g, const Value<float>& x, const Value<float>& y) {
    Dim4 d(1);
    auto xa = g.constant_as(x, d);
    auto ya = g.constant_as(y, d);
    auto h = g.matmul_nn(xa, ya);
    auto i = g.matmul_nn(h, ya);
    return g.as_scalar(i);
  });
}

// ---------------------------------------------------------------------------------------------
// tests/core.rs:6-31
// ---------------------------------------------------------------------------------------------
static void test_core() {
  {
    Graph1 g;
    auto a = g.variable<int32_t>(3), b = g.variable<int32_t>(2);
    auto c = g.add(a, b);
    auto gr = g.evaluate_gradients_once<int32_t>(c.gid(), 5);
    CHECK(*gr.get<int32_t>(a.gid()) == 5);
    CHECK(*gr.get<int32_t>(b.gid()) == 5);
  }
  {
    Graph1 g;
    auto a = g.variable<int32_t>(1), b = g.variable<int32_t>(2);
    auto c = g.constant<int32_t>(3);
    auto d = g.add_all<int32_t>({&a, &b, &c});
    CHECK(d.data == 6);
    auto gr = g.evaluate_gradients_once<int32_t>(d.gid(), 1);
    CHECK(*gr.get<int32_t>(a.gid()) == 1);
    CHECK(*gr.get<int32_t>(b.gid()) == 1);
  }
}

// ---------------------------------------------------------------------------------------------
// tests/arith.rs:8-63 (+ doctests lib.rs:74-89, 248-261)
// ---------------------------------------------------------------------------------------------
static void test_arith() {
  {
    Graph1 g;
    auto a = g.variable<int32_t>(3);
    auto b = g.zeros(a);
    CHECK(b.data == 0 && !b.id.has_value());
    auto c = g.ones(a);
    CHECK(c.data == 1 && !c.id.has_value());
  }
  {
    Graph1 g;
    auto a = g.variable<int32_t>(3);
    auto b = g.neg(a);
    CHECK(b.data == -3);
    auto gr = g.evaluate_gradients_once<int32_t>(b.gid(), 1);
    CHECK(*gr.get<int32_t>(a.gid()) == -1);
  }
  {
    Graph1 g;
    auto a = g.variable<int32_t>(1), b = g.variable<int32_t>(2);
    auto c = g.sub(a, b);
    CHECK(c.data == -1);
    auto gr = g.evaluate_gradients_once<int32_t>(c.gid(), 1);
    CHECK(*gr.get<int32_t>(a.gid()) == 1);
    CHECK(*gr.get<int32_t>(b.gid()) == -1);
  }
  {
    Graph1 g;
    auto a = g.variable<int32_t>(1), b = g.variable<int32_t>(2);
    auto c = g.mul(a, b);
    CHECK(c.data == 2);
    auto gr = g.evaluate_gradients_once<int32_t>(c.gid(), 1);
    CHECK(*gr.get<int32_t>(a.gid()) == 2);
    CHECK(*gr.get<int32_t>(b.gid()) == 1);
  }
  {  // doctest lib.rs:74-89: evaluate_gradients (non-consuming), f32
    Graph1 g;
    auto a = g.variable<float>(1.0f), b = g.variable<float>(2.0f);
    auto c = g.mul(a, b);
    auto gr = g.evaluate_gradients<float>(c.gid(), 1.0f);
    CHECK(*gr.get<float>(a.gid()) == 2.0f);
    auto gr2 = g.evaluate_gradients<float>(c.gid(), 1.0f);  // tape is reusable
    CHECK(*gr2.get<float>(b.gid()) == 1.0f);
  }
}

// ---------------------------------------------------------------------------------------------
// tests/const_arith.rs:6-37 (i32 with u8 constants)
// ---------------------------------------------------------------------------------------------
static void test_const_arith() {
  {
    Graph1 g;
    auto a = g.variable<int32_t>(2);
    auto b = g.addc(a, uint8_t(1));
    CHECK(b.data == 3);
    auto gr = g.evaluate_gradients_once<int32_t>(b.gid(), 1);
    CHECK(*gr.get<int32_t>(a.gid()) == 1);
  }
  {
    Graph1 g;
    auto a = g.variable<int32_t>(2);
    auto b = g.mulc(a, uint8_t(3));
    CHECK(b.data == 6);
    auto gr = g.evaluate_gradients_once<int32_t>(b.gid(), 1);
    CHECK(*gr.get<int32_t>(a.gid()) == 3);
  }
  {
    Graph1 g;
    auto a = g.variable<int32_t>(2);
    auto b = g.powc(a, uint8_t(3));
    CHECK(b.data == 8);
    auto gr = g.evaluate_gradients_once<int32_t>(b.gid(), 1);
    CHECK(*gr.get<int32_t>(a.gid()) == 12);
  }
}

// ---------------------------------------------------------------------------------------------
// tests/analytic.rs:11-169 (scalar f32 closed forms, direction 1.7, |err| < 1e-3)
// ---------------------------------------------------------------------------------------------
static void test_analytic() {
  const float dir = 1.7f;
  {
    Graph1 g; auto a = g.variable<float>(1.0f);
    auto b = g.exp(g.mulc(a, int16_t(2)));
    CHECK_NEAR(b.data, std::exp(2.0f));
    auto gr = g.evaluate_gradients_once<float>(b.gid(), dir);
    CHECK_NEAR(*gr.get<float>(a.gid()), b.data * 2.0f * dir);
  }
  {
    Graph1 g; auto a = g.variable<float>(1.0f);
    auto b = g.log(g.mulc(a, int16_t(2)));
    CHECK_NEAR(b.data, std::log(2.0f));
    auto gr = g.evaluate_gradients_once<float>(b.gid(), dir);
    CHECK_NEAR(*gr.get<float>(a.gid()), 1.0f * dir);
  }
  {
    Graph1 g; auto a = g.variable<float>(1.0f);
    auto b = g.log1p(a);
    CHECK_NEAR(b.data, std::log(2.0f));
    auto gr = g.evaluate_gradients_once<float>(b.gid(), dir);
    CHECK_NEAR(*gr.get<float>(a.gid()), 0.5f * dir);
  }
  {
    Graph1 g; auto a = g.variable<float>(1.0f);
    auto b = g.sin(a);
    CHECK_NEAR(b.data, std::sin(1.0f));
    auto gr = g.evaluate_gradients_once<float>(b.gid(), dir);
    CHECK_NEAR(*gr.get<float>(a.gid()), std::cos(1.0f) * dir);
  }
  {
    Graph1 g; auto a = g.variable<float>(1.0f);
    auto b = g.cos(a);
    CHECK_NEAR(b.data, std::cos(1.0f));
    auto gr = g.evaluate_gradients_once<float>(b.gid(), dir);
    CHECK_NEAR(*gr.get<float>(a.gid()), -std::sin(1.0f) * dir);
  }
  {
    Graph1 g; auto a = g.variable<float>(1.0f);
    auto b = g.tanh(a);
    CHECK_NEAR(b.data, std::tanh(1.0f));
    auto gr = g.evaluate_gradients_once<float>(b.gid(), dir);
    CHECK_NEAR(*gr.get<float>(a.gid()), (1.0f - b.data * b.data) * dir);
  }
  {
    Graph1 g; auto a = g.variable<float>(1.0f);
    auto b = g.sigmoid(a);
    CHECK_NEAR(b.data, 1.0f / (1.0f + std::exp(-1.0f)));
    auto gr = g.evaluate_gradients_once<float>(b.gid(), dir);
    CHECK_NEAR(*gr.get<float>(a.gid()), (1.0f - b.data) * b.data * dir);
  }
  {
    Graph1 g; auto a = g.variable<float>(2.0f);
    auto b = g.reciprocal(a);
    CHECK_NEAR(b.data, 0.5f);
    auto gr = g.evaluate_gradients_once<float>(b.gid(), dir);
    CHECK_NEAR(*gr.get<float>(a.gid()), -0.25f * dir);
  }
  {
    Graph1 g; auto a = g.variable<float>(2.0f);
    auto b = g.sqrt(a);
    CHECK_NEAR(b.data, std::sqrt(2.0f));
    auto gr = g.evaluate_gradients_once<float>(b.gid(), dir);
    CHECK_NEAR(*gr.get<float>(a.gid()), 0.5f / std::sqrt(2.0f) * dir);
  }
  {
    Graph1 g; auto a = g.variable<float>(2.0f), b = g.variable<float>(3.0f);
    auto c = g.div(a, b);
    CHECK_NEAR(c.data, 2.0f / 3.0f);
    auto gr = g.evaluate_gradients_once<float>(c.gid(), dir);
    CHECK_NEAR(*gr.get<float>(a.gid()), 1.0f / 3.0f * dir);
    CHECK_NEAR(*gr.get<float>(b.gid()), -2.0f / 9.0f * dir);
  }
  {
    Graph1 g; auto a = g.variable<float>(2.0f), b = g.variable<float>(3.0f);
    auto c = g.pow(a, b);
    CHECK_NEAR(c.data, 8.0f);
    auto gr = g.evaluate_gradients_once<float>(c.gid(), dir);
    CHECK_NEAR(*gr.get<float>(a.gid()), 12.0f * dir);
    CHECK_NEAR(*gr.get<float>(b.gid()), std::log(2.0f) * c.data * dir);
  }
}

// ---------------------------------------------------------------------------------------------
// tests/compare.rs:9-95
// ---------------------------------------------------------------------------------------------
static void test_compare() {
  {
    Graph1 g; auto a = g.variable<int32_t>(1), b = g.variable<int32_t>(2);
    auto c = g.min(a, b);
    CHECK(c.data == 1);
    auto gr = g.evaluate_gradients_once<int32_t>(c.gid(), 1);
    CHECK(*gr.get<int32_t>(a.gid()) == 1);
    CHECK(*gr.get<int32_t>(b.gid()) == 0);
  }
  {
    Graph1 g; auto a = g.variable<int32_t>(1), b = g.variable<int32_t>(2);
    auto c = g.max(a, b);
    CHECK(c.data == 2);
    auto gr = g.evaluate_gradients_once<int32_t>(c.gid(), 1);
    CHECK(*gr.get<int32_t>(a.gid()) == 0);
    CHECK(*gr.get<int32_t>(b.gid()) == 1);
  }
  {
    Graph1 g; auto a1 = g.variable<int32_t>(-1), a2 = g.variable<int32_t>(2);
    auto b1 = g.abs(a1), b2 = g.abs(a2);
    CHECK(b1.data == 1 && b2.data == 2);
    auto gr = g.evaluate_gradients<int32_t>(b1.gid(), 3);
    CHECK(*gr.get<int32_t>(a1.gid()) == -3);
    gr = g.evaluate_gradients<int32_t>(b2.gid(), 3);
    CHECK(*gr.get<int32_t>(a2.gid()) == 3);
  }
  {
    Graph1 g; auto a1 = g.variable<int32_t>(-1), a2 = g.variable<int32_t>(2);
    auto b1 = g.relu(a1), b2 = g.relu(a2);
    CHECK(b1.data == 0 && b2.data == 2);
    auto gr = g.evaluate_gradients<int32_t>(b1.gid(), 3);
    CHECK(*gr.get<int32_t>(a1.gid()) == 0);
    gr = g.evaluate_gradients<int32_t>(b2.gid(), 3);
    CHECK(*gr.get<int32_t>(a2.gid()) == 3);
  }
  {
    Graph1 g; auto a1 = g.variable<int32_t>(-1), a2 = g.variable<int32_t>(2);
    auto b1 = g.sign(a1), b2 = g.sign(a2);
    CHECK(b1.data == -1);
    CHECK(!b1.id.has_value() && !b2.id.has_value());
  }
  {
    Graph1 g;
    auto a = g.variable<int32_t>(1), b = g.variable<int32_t>(2), c = g.variable<int32_t>(3), d = g.variable<int32_t>(4);
    auto e = g.select_argmax(a, b, &c, &d);
    CHECK(e.data == 4);
    auto gr = g.evaluate_gradients_once<int32_t>(e.gid(), 1);
    CHECK(gr.get<int32_t>(a.gid()) == nullptr);
    CHECK(gr.get<int32_t>(b.gid()) == nullptr);
    CHECK(*gr.get<int32_t>(c.gid()) == 0);
    CHECK(*gr.get<int32_t>(d.gid()) == 1);
  }
}

// ---------------------------------------------------------------------------------------------
// tests/extension.rs:61-107 and doctest lib.rs:282-330 — a user-defined op through make_node.
// ---------------------------------------------------------------------------------------------
template <class Cfg, class D>
static Value<D> user_square(Graph<Cfg>& g, const Value<D>& v, bool doc_variant) {
  using Gr = Graph<Cfg>;
  using GD = typename Gr::template G<D>;
  D result = v.data * v.data;
  return g.template make_node<D>(result, {v.input()},
                                 [v, doc_variant](typename Gr::GA& graph, typename Gr::Store& store, const GD& gradient) {
                                   if (v.id) {
                                     const auto& c = graph.link(v);
                                     auto grad1 = graph.mul(gradient, c);
                                     auto grad2 = graph.mul(c, gradient);
                                     if (doc_variant) {
                                       auto grad = graph.add(grad1, grad2);
                                       store.template add_gradient<GD>(graph, *v.id, grad);
                                     } else {
                                       store.template add_gradient<GD>(graph, *v.id, grad1);
                                       store.template add_gradient<GD>(graph, *v.id, grad2);
                                     }
                                   }
                                 });
}
static void test_extension() {
  for (bool doc : {false, true}) {
    Graph1 g;
    auto a = g.variable<int32_t>(3);
    auto b = user_square(g, a, doc);
    CHECK(b.data == 9);
    auto gr = g.evaluate_gradients_once<int32_t>(b.gid(), 1);
    CHECK(*gr.get<int32_t>(a.gid()) == 6);
  }
}

// ---------------------------------------------------------------------------------------------
// Error behaviour pinned by the reference's semantics (graph.rs:150-158, 185-188; net.rs:187-189).
// ---------------------------------------------------------------------------------------------
static void test_errors() {
  {
    Graph1 g;
    auto c = g.constant<int32_t>(3);
    bool threw = false;
    try { c.gid(); } catch (const Error& e) { threw = e.kind == ErrKind::MissingId; }
    CHECK(threw);
  }
  {
    Graph1 g1, g2;
    auto a = g1.variable<int32_t>(3);
    (void)g2.variable<int32_t>(1);
    bool threw = false;
    // An id of another arena does not name a node here (id_arena::get → None → MissingNode).
    try { g2.evaluate_gradients<int32_t>(a.gid(), 1); } catch (const Error& e) { threw = e.kind == ErrKind::MissingNode; }
    CHECK(threw);
  }
  {  // gradient dims must equal forward dims (graph.rs:156)
    Graph1 g;
    HostArray<float> x(Dim4(4, 3), 1.0f);
    auto a = g.variable(x);
    auto b = g.neg(a);
    bool threw = false;
    try { g.evaluate_gradients_once(b.gid(), HostArray<float>(Dim4(3, 4), 1.0f)); }
    catch (const Error& e) { threw = e.kind == ErrKind::Dimensions; }
    CHECK(threw);
  }
  {  // Check rules (App. B) spot checks incl. error variants
    Check c;
    CHECK(c.matmul(Dim4(4, 3), Dim4(3, 5), {}, {}) == Dim4(4, 5));
    CHECK(c.matmul(Dim4(4, 3, 2, 2), Dim4(3, 5), {}, {}) == Dim4(4, 5, 2, 2));
    CHECK(c.matmul(Dim4(4, 3), Dim4(3, 5, 7, 1), {}, {}) == Dim4(4, 5, 7, 1));
    bool threw = false;
    try { c.matmul(Dim4(4, 3, 2, 1), Dim4(3, 5, 3, 1), {}, {}); } catch (const Error& e) { threw = e.kind == ErrKind::Dimensions; }
    CHECK(threw);
    threw = false;
    try { c.max_as(Dim4(4, 3), Dim4(2, 3)); } catch (const Error& e) { threw = e.kind == ErrKind::ReducedDimensions; }
    CHECK(threw);
    threw = false;
    try { c.sum_as(Dim4(4, 3), Dim4(2, 3)); } catch (const Error& e) { threw = e.kind == ErrKind::Dimensions; }
    CHECK(threw);
    CHECK(c.tile_as(Dim4(1, 3), Dim4(4, 3)) == Dim4(4, 3));
    CHECK(c.flat(Dim4(4, 3)) == Dim4(12));
    CHECK(c.transpose(Dim4(4, 3), false) == Dim4(3, 4));
  }
}

// ---------------------------------------------------------------------------------------------
// Array tape vs central finite differences, as the reference's af tests do
// (gad/src/arrayfire.rs:120-163; tests/{arith,analytic,array,matrix,array_compare}.rs af modules),
// on seeded inputs. L∞ < 1e-3 (2e-3 for matmul).
// ---------------------------------------------------------------------------------------------
static HostArray<float> randu(Dim4 d, uint64_t seed) {
  HostArray<float> r(d);
  uint64_t s = seed * 0x9E3779B97F4A7C15ull + 1;
  for (size_t i = 0; i < r.size(); i++) {
    s = s * 6364136223846793005ull + 1442695040888963407ull;
    r[i] = float((s >> 40) & 0xFFFFFF) / float(1 << 24);
  }
  return r;
}
template <class F>
static HostArray<float> estimate_gradient(const HostArray<float>& input, const HostArray<float>& direction, float eps, F f) {
  Eval ev;
  HostArray<float> grad(input.dims());
  std::vector<float> v(input.data(), input.data() + input.size());
  for (size_t i = 0; i < v.size(); i++) {
    float x = v[i];
    v[i] = x + eps;
    float y2 = ev.dot(f(HostArray<float>(input.dims(), v.data())), direction);
    v[i] = x - eps;
    float y1 = ev.dot(f(HostArray<float>(input.dims(), v.data())), direction);
    grad[i] = (y2 - y1) / (eps + eps);
    v[i] = x;
  }
  return grad;
}
static bool almost_all_equal(const HostArray<float>& a, const HostArray<float>& b, float prec) {
  if (!(a.dims() == b.dims())) return false;
  float d = 0;
  for (size_t i = 0; i < a.size(); i++) d = std::fmax(d, std::fabs(a[i] - b[i]));
  return d < prec;
}
static void test_array_fd() {
  Eval ev;
  Dim4 d43(4, 3);
  {  // tests/analytic.rs af test_tanh-style, tests/arith.rs af test_mul
    Graph1 g;
    auto a = g.variable(randu(d43, 1)), b = g.variable(randu(d43, 2));
    auto c = g.mul(g.tanh(a), b);
    HostArray<float> dir(d43, 1.0f);
    auto gr = g.evaluate_gradients_once(c.gid(), dir);
    auto est = estimate_gradient(a.data, dir, 0.001f, [&](const HostArray<float>& x) { return ev.mul(ev.tanh(x), b.data); });
    CHECK(almost_all_equal(*gr.get<HostArray<float>>(a.gid()), est, 0.001f));
  }
  {  // tests/matrix.rs:25-49
    Graph1 g;
    auto a = g.variable(randu(Dim4(4, 3), 3)), b = g.variable(randu(Dim4(3, 5), 4));
    auto c = g.matmul_nn(a, b);
    HostArray<float> dir(Dim4(4, 5), 1.0f);
    auto gr = g.evaluate_gradients_once(c.gid(), dir);
    auto est_a = estimate_gradient(a.data, dir, 0.001f, [&](const HostArray<float>& x) { return ev.matmul_nn(x, b.data); });
    auto est_b = estimate_gradient(b.data, dir, 0.001f, [&](const HostArray<float>& x) { return ev.matmul_nn(a.data, x); });
    CHECK(almost_all_equal(*gr.get<HostArray<float>>(a.gid()), est_a, 0.002f));
    CHECK(almost_all_equal(*gr.get<HostArray<float>>(b.gid()), est_b, 0.002f));
  }
  {  // tests/array.rs:40-52 sum_as, :54-68 tile_as, :70-86 norm2
    Graph1 g;
    auto a = g.variable(randu(d43, 5));
    auto b = g.sum_as(a, Dim4(1, 3));
    HostArray<float> dir(Dim4(1, 3), 1.0f);
    auto gr = g.evaluate_gradients_once(b.gid(), dir);
    auto est = estimate_gradient(a.data, dir, 0.001f, [&](const HostArray<float>& x) { return ev.sum_as(x, Dim4(1, 3)); });
    CHECK(almost_all_equal(*gr.get<HostArray<float>>(a.gid()), est, 0.001f));
  }
  {
    Graph1 g;
    auto a = g.variable(randu(Dim4(1, 3), 6));
    auto b = g.tile_as(a, d43);
    HostArray<float> dir(d43, 1.0f);
    auto gr = g.evaluate_gradients_once(b.gid(), dir);
    auto est = estimate_gradient(a.data, dir, 0.001f, [&](const HostArray<float>& x) { return ev.tile_as(x, d43); });
    CHECK(almost_all_equal(*gr.get<HostArray<float>>(a.gid()), est, 0.001f));
  }
  {
    Graph1 g;
    auto a = g.variable(randu(d43, 7));
    auto c = g.norm2(a);
    auto b = g.constant_as(c, Dim4(1));
    HostArray<float> dir(Dim4(1), 1.0f);
    auto gr = g.evaluate_gradients_once(b.gid(), dir);
    auto est = estimate_gradient(a.data, dir, 0.001f, [&](const HostArray<float>& x) { return ev.constant_as(ev.norm2(x), Dim4(1)); });
    CHECK(almost_all_equal(*gr.get<HostArray<float>>(a.gid()), est, 0.001f));
  }
  {  // tests/array_compare.rs:24-45 (argmax masses), :47-76 (softmax)
    Graph1 g;
    auto a = g.variable(randu(d43, 8));
    for (auto [rd, mass] : {std::pair{Dim4(1, 3), 3.0f}, std::pair{Dim4(4, 1), 4.0f}, std::pair{Dim4(1), 1.0f}}) {
      auto b = g.argmax_as(a, rd);
      CHECK(!b.id.has_value());
      auto s = g.as_scalar(g.sum_as(b, Dim4(1)));
      CHECK(std::fabs(s.data - mass) < 1.2e-7f);
      auto sm = g.softmax_as(a, rd);
      auto s2 = g.as_scalar(g.sum_as(sm, Dim4(1)));
      CHECK(std::fabs(s2.data - mass) < 1e-6f);
    }
    auto b = g.softmax_as(a, Dim4(1, 3));
    HostArray<float> dir(d43, 2.0f);
    auto gr = g.evaluate_gradients(b.gid(), dir);
    auto est = estimate_gradient(a.data, dir, 0.001f, [&](const HostArray<float>& x) { return ev.softmax_as(x, Dim4(1, 3)); });
    CHECK(almost_all_equal(*gr.get<HostArray<float>>(a.gid()), est, 0.001f));
    auto m = g.max_as(a, Dim4(1, 3));
    HostArray<float> dir2(Dim4(1, 3), 1.0f);
    auto gr2 = g.evaluate_gradients(m.gid(), dir2);
    auto est2 = estimate_gradient(a.data, dir2, 0.001f, [&](const HostArray<float>& x) { return ev.max_as(x, Dim4(1, 3)); });
    CHECK(almost_all_equal(*gr2.get<HostArray<float>>(a.gid()), est2, 0.001f));
  }
}

// ---------------------------------------------------------------------------------------------
// Rosenbrock-N built from gad ops; closed-form gradient.
// ---------------------------------------------------------------------------------------------
static void test_rosenbrock() {
  {
    Graph1 g;
    std::vector<Value<double>> x = {g.variable<double>(-1.2), g.variable<double>(1.0)};
    auto sq = g.mul(x[0], x[0]);
    auto d = g.sub(x[1], sq);
    auto d2 = g.mul(d, d);
    auto t1 = g.mulc(d2, int16_t(100));
    auto nx = g.neg(x[0]);
    auto e = g.addc(nx, int16_t(1));
    auto e2 = g.mul(e, e);
    auto term = g.add(t1, e2);
    auto f = g.add_all<double>({&term});
    CHECK(std::fabs(f.data - 24.2) < 1e-12);
    CHECK(g.num_nodes() == 2 + 8 * 1 + 1);
    auto gr = g.evaluate_gradients_once<double>(f.gid(), 1.0);
    CHECK(std::fabs(*gr.get<double>(x[0].gid()) - (-215.6)) < 1e-10);
    CHECK(std::fabs(*gr.get<double>(x[1].gid()) - (-88.0)) < 1e-10);
  }
}

int main() {
  test_symgraph1();
  test_gradient_simple();
  test_hessian_and_more();
  test_hessian_and_more_arrays();
  test_core();
  test_arith();
  test_const_arith();
  test_analytic();
  test_compare();
  test_extension();
  test_errors();
  test_array_fd();
  test_rosenbrock();
  std::printf("oracle acceptance: %d checks, %d failures\n", g_checks, g_failures);
  return g_failures == 0 ? 0 : 1;
}
