// One generic gad program, three algebras: Rosenbrock-N as SURVEY App. D writes it in gad ops, instantiated through the C++
// mirror (include/gadcu.hpp) on Eval (forward only), on Graph1 with one-element arrays (the reference's scalar tape, one
// example at a time: net_ext.rs:50-66) and on BatchedGraph1 (one scalar tape per lane, forward + reverse sweep in ONE kernel).
// Checked: N = 2 at (−1.2, 1): f = 24.2, ∇f = (−215.6, −88); at (1, …, 1): f = 0, ∇f = 0; N = 64 over 4096 lanes against
// the closed form ∂f/∂x_i = −400 x_i (x_{i+1} − x_i²) − 2 (1 − x_i) [i < N−1] + 200 (x_i − x_{i−1}²) [i > 0] in f64, and
// BIT FOR BIT against the scalar tape of sampled lanes (the program only adds, subtracts and multiplies: every step is
// correctly rounded on both paths); the finished sweep re-run at new points (Store::rerun) equals a freshly recorded one.
//   g++ -std=c++17 -Iinclude examples/rosenbrock_cpp.cc -Lgad_b200/lib -lgadcu -Wl,-rpath,$PWD/gad_b200/lib -o examples/rosenbrock_cpp
#include <cmath>
#include <cstdio>
#include <cstring>

#include "gadcu.hpp"

using namespace gadcu;
using A = Array<double>;
using V = Value<double>;

static int failures = 0;
#define EXPECT(cond)                                                \
  do {                                                              \
    if (!(cond)) {                                                  \
      std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); \
      failures++;                                                   \
    }                                                               \
  } while (0)

// f(x) = Σ_i 100 (x_{i+1} − x_i²)² + (1 − x_i)², op for op
template <class G>
static V rosenbrock(G& g, const std::vector<V>& x) {
  std::vector<V> terms;
  for (size_t i = 0; i + 1 < x.size(); i++) {
    V sq = g.mul(x[i], x[i]);
    V d = g.sub(x[i + 1], sq);
    V d2 = g.mul(d, d);
    V t1 = g.mulc(d2, int16_t(100));
    V nx = g.neg(x[i]);
    V e = g.addc(nx, int16_t(1));
    V e2 = g.mul(e, e);
    terms.push_back(g.add(t1, e2));
  }
  std::vector<const V*> ptrs;
  for (const V& t : terms) ptrs.push_back(&t);
  return g.add_all(ptrs);
}

static std::vector<double> closed_form(const std::vector<double>& x) {
  const size_t n = x.size();
  std::vector<double> g(n, 0.0);
  for (size_t i = 0; i < n; i++) {
    if (i + 1 < n) g[i] += -400.0 * x[i] * (x[i + 1] - x[i] * x[i]) - 2.0 * (1.0 - x[i]);
    if (i > 0) g[i] += 200.0 * (x[i] - x[i - 1] * x[i - 1]);
  }
  return g;
}

// the scalar tape: one example, every variable a one-element array
static std::pair<double, std::vector<double>> scalar_tape(Context& ctx, const std::vector<double>& point) {
  Graph1 g(ctx);
  std::vector<V> x;
  for (double v : point) x.push_back(g.variable(A::from_host(ctx, {v}, Dim4{1})));
  V f = rosenbrock(g, x);
  Store<double> store = g.evaluate_gradients_once(f.gid(), A::from_host(ctx, {1.0}, Dim4{1}));
  std::vector<double> grad;
  for (const V& v : x) grad.push_back(store.get(v.gid())->data().item());
  return {f.data().item(), grad};
}

// the batched tapes: columns[i][lane] = x_i of that lane's tape
struct Batched {
  std::unique_ptr<BatchedGraph1<double>> tape;
  std::vector<double> f;                  // [lanes]
  std::vector<std::vector<double>> grad;  // [N][lanes]
  Store<double> store;
};
static Batched batched_tapes(Context& ctx, const std::vector<A>& columns, uint64_t lanes) {
  auto g = std::make_unique<BatchedGraph1<double>>(ctx, lanes);
  std::vector<V> x;
  for (const A& c : columns) x.push_back(g->variable(c));
  V f = rosenbrock(*g, x);
  Store<double> store = g->evaluate_gradients_once(f.gid(), 1.0);
  std::vector<std::vector<double>> grad;
  for (const V& v : x) grad.push_back(store.get(v.gid())->data().host());
  std::vector<double> fv = g->force(f).host();  // a forward value's column
  return Batched{std::move(g), std::move(fv), std::move(grad), store};
}

static std::vector<A> random_columns(Context& ctx, size_t n, uint64_t lanes, uint64_t seed) {
  std::vector<A> columns;
  for (size_t i = 0; i < n; i++) columns.push_back(A::random(ctx, Dim4{lanes}, seed + i, false, -1.2, 1.2));  // x_i ~ U(−1.2, 1.2), App. D
  return columns;
}
static bool same_bits(double a, double b) { return std::memcmp(&a, &b, 8) == 0; }

int main() {
  try {
    Context ctx(0);
    // known answers
    auto r2 = scalar_tape(ctx, {-1.2, 1.0});
    EXPECT(std::fabs(r2.first - 24.2) <= 1e-12 && std::fabs(r2.second[0] + 215.6) <= 1e-11 && std::fabs(r2.second[1] + 88.0) <= 1e-12);
    auto ones = scalar_tape(ctx, std::vector<double>(8, 1.0));
    EXPECT(ones.first == 0.0);
    for (double v : ones.second) EXPECT(v == 0.0);
    {  // forward only, on Eval: the same program, no tape
      Eval e(ctx);
      std::vector<V> x{e.constant(A::from_host(ctx, {-1.2}, Dim4{1})), e.constant(A::from_host(ctx, {1.0}, Dim4{1}))};
      EXPECT(std::fabs(rosenbrock(e, x).data().item() - 24.2) <= 1e-12);
    }
    // N = 64 over 4096 lanes
    const size_t N = 64;
    const uint64_t lanes = 4096;
    const std::vector<A> columns = random_columns(ctx, N, lanes, 200);
    std::vector<std::vector<double>> xs;
    for (const A& c : columns) xs.push_back(c.host());
    Batched b = batched_tapes(ctx, columns, lanes);
    EXPECT(b.f.size() == lanes && b.grad.size() == N && b.grad[0].size() == lanes);
    const auto stats = b.store.program_stats();
    EXPECT(stats.first > 0 && stats.second > 0);
    double worst = 0;
    for (uint64_t lane = 0; lane < lanes; lane++) {
      std::vector<double> point(N);
      for (size_t i = 0; i < N; i++) point[i] = xs[i][lane];
      const std::vector<double> expect = closed_form(point);
      for (size_t i = 0; i < N; i++) {
        // the tape's own rounding: a dozen f64 operations on terms of magnitude up to 400 |x| |x_{i+1} − x_i²|
        const double scale = 400.0 * std::fabs(point[i]) * (2.7 + point[i] * point[i]) + 200.0 * 2.7 + 4.4;
        worst = std::fmax(worst, std::fabs(b.grad[i][lane] - expect[i]) / scale);
      }
    }
    EXPECT(worst <= 16 * 2.220446049250313e-16);
    for (uint64_t lane : {uint64_t(0), uint64_t(1), uint64_t(77), lanes - 1}) {  // bit for bit against the scalar tape
      std::vector<double> point(N);
      for (size_t i = 0; i < N; i++) point[i] = xs[i][lane];
      auto one = scalar_tape(ctx, point);
      EXPECT(same_bits(one.first, b.f[lane]));
      for (size_t i = 0; i < N; i++) EXPECT(same_bits(one.second[i], b.grad[i][lane]));
    }
    // the same tapes at new points through the compiled program == recording them again
    const std::vector<A> columns2 = random_columns(ctx, N, lanes, 900);
    const std::vector<A> again = b.store.rerun(columns2);
    Batched fresh = batched_tapes(ctx, columns2, lanes);
    EXPECT(again.size() == N);
    for (size_t i = 0; i < N && i < again.size(); i++) {
      EXPECT(bool(again[i]));
      if (!again[i]) continue;
      const std::vector<double> h = again[i].host();
      EXPECT(h.size() == lanes && std::memcmp(h.data(), fresh.grad[i].data(), lanes * 8) == 0);
    }
  } catch (const Error& e) {
    std::printf("FAILED with gadcu::Error kind %d: %s\n", int(e.kind), e.what());
    return 2;
  }
  std::printf(failures ? "%d expectation(s) failed\n" : "rosenbrock: scalar tape, batched tapes and closed form agree (%d failures)\n", failures);
  return failures ? 1 : 0;
}
