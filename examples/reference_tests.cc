// The reference's own tests, replayed through the C++ host mirror (include/gadcu.hpp) on the device backend.
// gad/tests/graph.rs:8-69, gad/tests/extension.rs:61-107, plus the array-flavoured checks of gad/tests/{core,array,
// matrix}.rs (dims, error variants, absent-vs-present store entries). Scalars of the reference's tests are one-element
// f32 arrays here (the backend's data type is the array). Exit code 0 = all passed.
//   g++ -std=c++17 -Iinclude examples/reference_tests.cc -Lgad_b200/lib -lgadcu -Wl,-rpath,$PWD/gad_b200/lib -o examples/reference_tests
#include <cmath>
#include <cstdio>

#include "gadcu.hpp"

using namespace gadcu;
using A = Array<float>;
using V = Value<float>;

static int failures = 0;
#define EXPECT(cond)                                                        \
  do {                                                                      \
    if (!(cond)) {                                                          \
      std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond);         \
      failures++;                                                           \
    }                                                                       \
  } while (0)
template <class F> static bool throws(ErrorKind kind, F f) {
  try {
    f();
  } catch (const Error& e) {
    return e.kind == kind;
  }
  return false;
}
static A one(Context& ctx, float v) { return A::from_host(ctx, {v}, Dim4{1}); }

// gad/tests/graph.rs:8-26
static void test_gradient_simple(Context& ctx) {
  Graph1 g(ctx);
  V a = g.variable(one(ctx, 3.0f));
  V x = g.mul(a, a);
  V c = g.add(a, x);
  EXPECT(c.data().item() == 3 * 3 + 3);
  const Id ia = a.gid(), ic = c.gid();
  Store<float> g1 = Graph1(g).evaluate_gradients_once(ic, one(ctx, 1.0f));  // g.clone()
  Store<float> g2 = g.evaluate_gradients_once(ic, one(ctx, 2.0f));
  EXPECT(g1.get(ia)->data().item() == 7);
  EXPECT(g2.get(ia)->data().item() == 14);
}

// gad/tests/graph.rs:28-69
static void test_hessian_and_more(Context& ctx) {
  GraphN g(ctx);
  V x = g.variable(one(ctx, 1.0f)), y = g.variable(one(ctx, 0.4f));
  V h = g.mul(x, y);
  V z = g.mul(h, y);  // z = x * y^2
  const Id ix = x.gid(), iy = y.gid(), iz = z.gid();
  V dz = g.constant(one(ctx, 1.0f));
  Store<float> dz_d = g.compute_gradients(iz, dz);
  V dz_dx = *dz_d.get(ix), dz_dy = *dz_d.get(iy);
  V ddz = g.constant(one(ctx, 1.0f));
  Store<float> ddz_dxd = g.compute_gradients(dz_dx.gid(), ddz);
  Store<float> ddz_dyd = g.compute_gradients(dz_dy.gid(), ddz);
  EXPECT(!ddz_dxd.get(ix).has_value());                       // None, not zero
  EXPECT(ddz_dxd.get(iy)->data().item() == 0.8f);             // 2y
  EXPECT(ddz_dyd.get(ix)->data().item() == 0.8f);
  EXPECT(ddz_dyd.get(iy)->data().item() == 2.0f);
  V dddz = g.constant(one(ctx, 1.0f));
  Store<float> dddz_dxdyd = g.compute_gradients(ddz_dxd.get(iy)->gid(), dddz);
  EXPECT(!dddz_dxdyd.get(ix).has_value());
  EXPECT(dddz_dxdyd.get(iy)->data().item() == 2.0f);
}

// gad/tests/extension.rs:61-107: a user-defined operator with its own VJP
template <class G> static V square(G& g, Eval& eval, const V& v) {
  V result = eval.mul(eval.constant(v.data()), eval.constant(v.data()));
  return g.template make_node<float>(result.data(), {&v}, [v](Algebra& graph, StoreRef& store, const V& gradient) {
    if (auto id = v.id()) {
      V c = graph.link(v);
      V grad1 = graph.mul(gradient, c);
      V grad2 = graph.mul(c, gradient);
      store.add_gradient(graph, *id, grad1);
      store.add_gradient(graph, *id, grad2);
    }
  });
}
static void test_square(Context& ctx) {
  Eval eval(ctx);
  Graph1 g(ctx);
  V a = g.variable(one(ctx, 3.0f));
  V b = square(g, eval, a);
  EXPECT(b.data().item() == 9);
  Store<float> gradients = g.evaluate_gradients_once(b.gid(), one(ctx, 1.0f));
  EXPECT(gradients.get(a.gid())->data().item() == 6);
  // the same operator on the higher-order tape: d/da a^2 = 2a, d2/da2 = 2
  GraphN gn(ctx);
  V an = gn.variable(one(ctx, 3.0f));
  V bn = square(gn, eval, an);
  Store<float> d1 = gn.compute_gradients(bn.gid(), gn.constant(one(ctx, 1.0f)));
  V da = *d1.get(an.gid());
  EXPECT(da.data().item() == 6);
  Store<float> d2 = gn.compute_gradients(da.gid(), gn.constant(one(ctx, 1.0f)));
  EXPECT(d2.get(an.gid())->data().item() == 2);
}

// Check + Error variants: dims are decided on the host before any launch (arith.rs:66, matrix.rs:52, array.rs:71-120)
static void test_check_and_errors(Context& ctx) {
  EXPECT((Check::add(Dim4{2, 3}, Dim4{2, 3}) == Dim4{2, 3}));
  EXPECT(throws(ErrorKind::Dimensions, [] { Check::add(Dim4{2, 3}, Dim4{3, 2}); }));
  EXPECT((Check::matmul(Dim4{4, 3}, Dim4{4, 5}, MatProp{true, false}, MatProp{}) == Dim4{3, 5}));
  EXPECT(throws(ErrorKind::Dimensions, [] { Check::matmul(Dim4{4, 3}, Dim4{4, 5}); }));
  EXPECT((Check::transpose(Dim4{2, 7}) == Dim4{7, 2}));
  EXPECT(throws(ErrorKind::Dimensions, [] { Check::transpose(Dim4{2, 7, 2}); }));
  EXPECT((Check::tile_as(Dim4{1, 5}, Dim4{4, 5}) == Dim4{4, 5}));
  EXPECT(throws(ErrorKind::Dimensions, [] { Check::sum_as(Dim4{4, 5}, Dim4{2, 5}); }));
  EXPECT((Check::flat(Dim4{2, 3, 4}) == Dim4{24}));
  Graph1 g(ctx);
  V a = g.variable(A::from_host(ctx, {1, 2, 3, 4, 5, 6}, Dim4{2, 3}));
  V b = g.constant(A::from_host(ctx, {1, 2, 3, 4, 5, 6}, Dim4{3, 2}));
  const uint64_t launches = ctx.launches();
  EXPECT(throws(ErrorKind::Dimensions, [&] { g.add(a, b); }));
  EXPECT(throws(ErrorKind::Dimensions, [&] { g.as_scalar(a); }));
  EXPECT(throws(ErrorKind::MissingId, [&] { b.gid(); }));
  EXPECT(ctx.launches() == launches);  // a failing call launches nothing
  EXPECT(!b.id().has_value() && a.id().has_value());
  EXPECT((g.zeros(a).id() == std::nullopt));  // zeros / ones are constants (arith.rs:165-173)
}

// matmul forward + VJP on small integers (exact), column-major: gad/src/matrix.rs:141-199
static void test_matmul_gradients(Context& ctx) {
  Graph1 g(ctx);
  // A [2,3] = [[1,3,5],[2,4,6]], B [3,2] = [[1,4],[2,5],[3,6]] in column-major order
  V a = g.variable(A::from_host(ctx, {1, 2, 3, 4, 5, 6}, Dim4{2, 3}));
  V b = g.variable(A::from_host(ctx, {1, 2, 3, 4, 5, 6}, Dim4{3, 2}));
  V c = g.matmul_nn(a, b);
  EXPECT((c.data().dims() == Dim4{2, 2}));
  EXPECT((c.data().host() == std::vector<float>{22, 28, 49, 64}));
  V s = g.as_scalar(g.sum_as(c, Dim4{1}));
  EXPECT(s.data().item() == 22 + 28 + 49 + 64);
  Store<float> grads = g.evaluate_gradients_once(s.gid(), 1.0f);
  // dA = 1·Bᵀ: every row of dA is the row sums of B; dB = Aᵀ·1: every column of dB is the column sums of A
  EXPECT((grads.get(a.gid())->data().host() == std::vector<float>{5, 5, 7, 7, 9, 9}));
  EXPECT((grads.get(b.gid())->data().host() == std::vector<float>{3, 7, 11, 3, 7, 11}));
}

// WeightOps (net.rs:161-180) under an apply_gradient_step-style loop (net_ext.rs:47-73): w += -lambda * dL/dw for a tanh
// layer with square loss; every step's loss against the same recurrence computed on the host in double
static void test_weight_ops_and_descent(Context& ctx) {
  const std::vector<float> w0 = {0.5f, -0.25f, 0.125f, 1.0f}, xs = {1, 2, 3, 4, 5, 6}, ts = {0.1f, 0.2f, 0.3f, 0.4f, 0.5f, 0.6f};
  A w = A::from_host(ctx, w0, Dim4{2, 2});
  const A x = A::from_host(ctx, xs, Dim4{3, 2}), t = A::from_host(ctx, ts, Dim4{3, 2});
  double hw[4] = {w0[0], w0[1], w0[2], w0[3]};  // column-major [2,2]
  const float lambda = -0.005f;
  float first = 0, last = 0;
  for (int step = 0; step < 10; step++) {
    Graph1 g(ctx);
    V vw = g.variable(w);
    V out = g.tanh(g.matmul_nn(g.constant(x), vw));
    V loss = g.norm2(g.sub(g.constant(t), out));
    Store<float> grads = g.evaluate_gradients_once(loss.gid(), 1.0f);
    w.add_assign(grads.get(vw.gid())->data().scale(lambda));
    // host: out[i,j] = tanh(sum_k x[i,k] w[k,j]); L = sum (t - out)^2; dL/dw[k,j] = sum_i x[i,k] * (-2 (t-out) (1-out^2))[i,j]
    double L = 0, gw[4] = {0, 0, 0, 0};
    for (int j = 0; j < 2; j++)
      for (int i = 0; i < 3; i++) {
        const double o = std::tanh(xs[i] * hw[2 * j] + xs[3 + i] * hw[2 * j + 1]), d = ts[3 * j + i] - o;
        L += d * d;
        for (int k = 0; k < 2; k++) gw[2 * j + k] += xs[3 * k + i] * (-2.0 * d * (1.0 - o * o));
      }
    for (int k = 0; k < 4; k++) hw[k] += double(lambda) * gw[k];
    const float got = loss.data().item();
    EXPECT(std::fabs(got - L) <= 1e-4 * L);
    (step == 0 ? first : last) = got;
  }
  EXPECT(last < first);
  const std::vector<float> wf = w.host();
  for (int k = 0; k < 4; k++) EXPECT(std::fabs(wf[k] - hw[k]) <= 1e-4);
}

int main() {
  try {
    Context ctx(0);
    test_gradient_simple(ctx);
    test_hessian_and_more(ctx);
    test_square(ctx);
    test_check_and_errors(ctx);
    test_matmul_gradients(ctx);
    test_weight_ops_and_descent(ctx);
  } catch (const Error& e) {
    std::printf("FAILED with gadcu::Error kind %d: %s\n", int(e.kind), e.what());
    return 2;
  }
  std::printf(failures ? "%d expectation(s) failed\n" : "all reference tests passed (%d failures)\n", failures);
  return failures ? 1 : 0;
}
