// CPU test of include/gadcu_net.hpp (the C++ mirror of gad/src/net.rs + net_ext.rs): the combinators, WeightOps over
// weight structures, apply_gradient_step and the bincode checkpoint layout, instantiated over the CPU ORACLE's tape
// instead of the device algebra — the header is generic over the algebra like the reference, so the logic that is
// checked here is the logic that runs above libgadcu.so (examples/net_cpp.cc is the device instantiation).
// Replays gad/tests/net.rs:8-154 (a hand-written Net, make_net = input.using(weight).map(matmul_nn), training a 3×3
// map to the identity, weight round trips through bytes, check() dims) and adds the error paths the reference's
// combinators have (Lengths, Dimensions, Empty, MissingGradient) plus a closed-form check of one gradient step.
//   g++ -std=c++20 -Iinclude -Ioracle tests/cpp/net_test.cc -Lgad_b200/lib -lgadcu -Wl,-rpath,$PWD/gad_b200/lib
// (libgadcu.so is linked for gadcu::Check's host-side dimension rules only; no device is touched.)
// `net_test hex` prints the bincode bytes of a fixed weight structure: tests/test_net_cpp.py compares them with
// gad_b200/net.py's writer.
#include <cmath>
#include <cstdio>
#include <cstring>

#include "gad_oracle.hpp"
#include "gadcu_net.hpp"

namespace go = gad_oracle;
using gadcu::Dim4;
using gadcu::Error;
using gadcu::ErrorKind;
using namespace gadcu::net;

// ---- the oracle behind the member surface gadcu.hpp gives the device algebra -------------------------------------------
namespace host {

template <class T>
struct Leaf {  // plays gadcu::Array<T>: a value with O(1) copies, dims(), item() and the WeightOps of net.rs:161-180
  go::HostArray<T> a;
  Leaf() = default;
  explicit Leaf(go::HostArray<T> x) : a(std::move(x)) {}
  Leaf(std::initializer_list<T> values, Dim4 d) : a(go::Dim4(d[0], d[1], d[2], d[3]), std::data(values)) {}
  Dim4 dims() const { return Dim4{a.dims()[0], a.dims()[1], a.dims()[2], a.dims()[3]}; }
  void add_assign(const Leaf& o) { a = go::Eval().add(a, o.a); }  // `+=` rebinds to a fresh array, like af::Array
  Leaf scale(T lambda) const { return Leaf(go::Eval().scale(lambda, a)); }
  T item() const { return a[0]; }
};
template <class T> struct Scalar { T x; T item() const { return x; } };

template <class T>
struct Val {
  go::Value<go::HostArray<T>> v;
  Leaf<T> data() const { return Leaf<T>(v.data); }
  std::optional<go::Id> id() const { return v.id; }
  go::Id gid() const {
    if (!v.id) throw Error(ErrorKind::MissingId, "Trying to obtain `id` of a constant value.");
    return *v.id;
  }
};
template <class T>
struct SVal {
  go::Value<T> v;
  Scalar<T> data() const { return {v.data}; }
  std::optional<go::Id> id() const { return v.id; }
  go::Id gid() const {
    if (!v.id) throw Error(ErrorKind::MissingId, "Trying to obtain `id` of a constant value.");
    return *v.id;
  }
};
template <class T>
struct Store {
  go::GenericGradientMap1 m;
  std::optional<Val<T>> get(go::Id id) const {
    const go::HostArray<T>* p = m.get<go::HostArray<T>>(id);
    if (!p) return std::nullopt;
    return Val<T>{go::Value<go::HostArray<T>>::constant(*p)};
  }
};
// translate the oracle's errors into the mirror's (same variants, error.rs:9-40)
template <class F> auto translated(F f) {
  try {
    return f();
  } catch (const go::Error& e) {
    throw Error(ErrorKind(int(e.kind)), e.what());
  }
}
template <class G>
struct AlgebraOver {  // G = go::Graph1 (values carry ids) or go::Eval wrapped below
  G g;
  template <class T> Val<T> variable(const Leaf<T>& d) { return {g.variable(d.a)}; }
  template <class T> Val<T> constant(const Leaf<T>& d) { return {g.constant(d.a)}; }
  template <class T> Val<T> add(const Val<T>& a, const Val<T>& b) { return translated([&] { return Val<T>{g.add(a.v, b.v)}; }); }
  template <class T> Val<T> sub(const Val<T>& a, const Val<T>& b) { return translated([&] { return Val<T>{g.sub(a.v, b.v)}; }); }
  template <class T> Val<T> mul(const Val<T>& a, const Val<T>& b) { return translated([&] { return Val<T>{g.mul(a.v, b.v)}; }); }
  template <class T> Val<T> tanh(const Val<T>& a) { return {g.tanh(a.v)}; }
  template <class T> Val<T> matmul_nn(const Val<T>& a, const Val<T>& b) { return translated([&] { return Val<T>{g.matmul(a.v, b.v, {}, {})}; }); }
  template <class T> SVal<T> norm2(const Val<T>& a) { return {g.norm2(a.v)}; }
  template <class T> Store<T> evaluate_gradients_once(go::Id id, T seed) { return translated([&] { return Store<T>{g.evaluate_gradients_once(id, seed)}; }); }
};
using Graph1 = AlgebraOver<go::Graph1>;
// Eval: values are plain data and carry no id (lib.rs:523-524)
struct EvalAlgebra {
  go::Eval e;
  template <class T> Val<T> variable(const Leaf<T>& d) { return {go::Value<go::HostArray<T>>::constant(d.a)}; }
  template <class T> Val<T> constant(const Leaf<T>& d) { return variable(d); }
  template <class T> Val<T> matmul_nn(const Val<T>& a, const Val<T>& b) {
    return translated([&] { return Val<T>{go::Value<go::HostArray<T>>::constant(e.matmul(a.v.data, b.v.data, {}, {}))}; });
  }
  template <class T> Val<T> sub(const Val<T>& a, const Val<T>& b) { return translated([&] { return Val<T>{go::Value<go::HostArray<T>>::constant(e.sub(a.v.data, b.v.data))}; }); }
  template <class T> Val<T> tanh(const Val<T>& a) { return {go::Value<go::HostArray<T>>::constant(e.tanh(a.v.data))}; }
  template <class T> SVal<T> norm2(const Val<T>& a) { return {go::Value<T>::constant(e.dot(a.v.data, a.v.data))}; }
};

}  // namespace host

// how a host leaf is written to / rebuilt from a checkpoint (gadcu_net.hpp: LeafIO)
namespace gadcu { namespace net {
struct HostMaker {};
template <class T> struct LeafIO<host::Leaf<T>> {
  using Scalar = T;
  static std::vector<T> host(const ::host::Leaf<T>& l) { return std::vector<T>(l.a.data(), l.a.data() + l.a.size()); }
  static std::array<uint64_t, 4> dims(const ::host::Leaf<T>& l) { return {l.a.dims()[0], l.a.dims()[1], l.a.dims()[2], l.a.dims()[3]}; }
  static ::host::Leaf<T> make(HostMaker&, const std::vector<T>& values, const std::array<uint64_t, 4>& d) {
    return ::host::Leaf<T>(go::HostArray<T>(go::Dim4(d[0], d[1], d[2], d[3]), values.data()));
  }
};
}}  // namespace gadcu::net

using L = host::Leaf<float>;
static int failures = 0;
#define EXPECT(cond)                                                \
  do {                                                              \
    if (!(cond)) {                                                  \
      std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); \
      failures++;                                                   \
    }                                                               \
  } while (0)
template <class F> static bool throws(ErrorKind kind, F f) {
  try {
    f();
  } catch (const Error& e) {
    return e.kind == kind;
  }
  return false;
}
static bool close(const L& a, const L& b, float tol) {
  if (a.dims() != b.dims()) return false;
  for (size_t i = 0; i < a.a.size(); i++)
    if (!(std::fabs(a.a[i] - b.a[i]) <= tol)) return false;
  return true;
}

// gad/tests/net.rs:8-78: a net written by hand against the five-method interface
class TestNet : public NetBase<TestNet> {
 public:
  using Input = L;
  using Weights = L;
  TestNet(Dim4 dims, L weights) : dims_(dims), weights_(std::move(weights)) {}
  template <class A> auto eval_with_gradient_info(A& g, const Input& input) const {
    if (input.dims() != dims_) throw Error(ErrorKind::Dimensions, "TestNet input");
    auto in = g.constant(input);
    auto w = g.variable(weights_);
    auto out = g.matmul_nn(in, w);
    return std::make_pair(std::move(out), gid_of(w));
  }
  Weights get_weights() const { return weights_; }
  void update_weights(const Weights& delta) {
    check_equal_dimensions("update_weights", delta.dims(), weights_.dims());
    weights_.add_assign(delta);
  }
  void set_weights(const Weights& w) {
    check_equal_dimensions("set_weights", w.dims(), weights_.dims());
    weights_ = w;
  }
  template <class I, class Reader> Weights read_weight_gradients(const std::optional<I>& info, const Reader& store) const {
    auto g = info ? store.get(*info) : std::nullopt;
    if (!g) throw Error(ErrorKind::MissingGradient, "TestNet");
    return g->data();
  }

 private:
  Dim4 dims_;
  L weights_;
};

// gad/tests/net.rs:80-89
static auto make_net(const L& w0) {
  return InputData<L>(Dim4{3, 3}).using_(WeightData<L>(w0)).map([](auto& g, const auto& iw) { return g.matmul_nn(iw.first, iw.second); });
}

static const L kA({1.0f, 2.0f, 1.0f, 1.0f, 0.0f, 1.0f, 0.0f, -2.0f, -1.0f}, Dim4{3, 3});  // tests/net.rs:95-98 (column-major)
static const L kI({1, 0, 0, 0, 1, 0, 0, 0, 1}, Dim4{3, 3});
static const L kW0({0.3f, -0.2f, 0.5f, 0.1f, 0.4f, -0.6f, -0.7f, 0.2f, 0.9f}, Dim4{3, 3});  // (the reference draws randn; any start converges)
static auto new_graph = [] { return host::Graph1{}; };

template <class Train, class Plain>
static void train_to_identity(Train train, Plain fresh) {
  std::vector<typename Train::Input> samples{{kA, kI}};
  float loss = 0, first = -1;
  int steps = 0;
  do {
    loss = apply_gradient_step(train, -0.01f, samples, new_graph);
    EXPECT(std::isfinite(loss));
    if (first < 0) first = loss;
  } while (loss >= 1e-6f && ++steps < 200000);
  EXPECT(loss < 1e-6f && loss < first);
  fresh.set_weights(train.get_weights());
  host::EvalAlgebra eval;
  EXPECT(close(fresh.eval(eval, kA).data(), kI, 0.01f));  // tests/net.rs:108-111: net.evaluate(a) ≈ identity
}

static void test_testnet() { train_to_identity(TestNet(Dim4{3, 3}, kW0).add_square_loss<L>(), TestNet(Dim4{3, 3}, kI)); }

static void test_make_net() {  // tests/net.rs:116-154
  train_to_identity(make_net(kW0).add_square_loss<L>(), make_net(kI));
  auto train = make_net(kW0).add_square_loss<L>();
  apply_gradient_step(train, -0.01f, std::vector<decltype(train)::Input>{{kA, kI}}, new_graph);
  using W = decltype(train)::Weights;  // pair<Unit, Leaf>: Using(InputData weights, WeightData weights)
  static_assert(std::is_same<W, std::pair<Unit, L>>::value, "weights follow the net's structure");
  const std::vector<uint8_t> bytes = serialize_weights_bincode(train.get_weights());
  EXPECT(bytes.size() == 4 + 32 + 8 + 9 * 4);  // u32 dtype, 4 x u64 dims, u64 count, 9 f32 — `()` writes nothing
  HostMaker maker;
  auto net = make_net(kI);
  net.set_weights(deserialize_weights_bincode<W>(maker, bytes));
  EXPECT(close(net.get_weights().second, train.get_weights().second, 0.0f));
  EXPECT(net.check(kA) == kI.dims());  // tests/net.rs:147-151: check() computes dims only
  EXPECT(throws(ErrorKind::Dimensions, [&] { net.check(L({1, 2}, Dim4{2, 1})); }));  // InputData's declared dims (net.rs:270)
  EXPECT(throws(ErrorKind::Dimensions, [&] { net.set_weights(W(Unit{}, L({1, 2}, Dim4{2, 1}))); }));
  std::vector<uint8_t> bad = bytes;
  bad.pop_back();
  EXPECT(throws(ErrorKind::Invalid, [&] { deserialize_weights_bincode<W>(maker, bad); }));
  bad = bytes;
  bad[0] = 2;  // an f64 blob read as f32 weights
  EXPECT(throws(ErrorKind::Type, [&] { deserialize_weights_bincode<W>(maker, bad); }));
  bad = bytes;
  bad.push_back(0);
  EXPECT(throws(ErrorKind::Invalid, [&] { deserialize_weights_bincode<W>(maker, bad); }));
}

// one step against the closed form: out = A·W, loss = ‖T − A·W‖², dL/dW = −2 Aᵀ(T − A·W); W += λ · Σ_examples dL/dW
static void test_gradient_step_closed_form() {
  auto train = make_net(kW0).add_square_loss<L>();
  const L t2({0.5f, 0, 1, -1, 2, 0, 0, 1, 1}, Dim4{3, 3});
  std::vector<decltype(train)::Input> batch{{kA, kI}, {kI, t2}};  // two examples: gradients are summed, ONE update (net_ext.rs:62-70)
  const float lambda = -0.05f;
  const float loss = apply_gradient_step(train, lambda, batch, new_graph);
  double expect_loss = 0, w[9], dw[9] = {0};
  for (int i = 0; i < 9; i++) w[i] = kW0.a[i];
  const L* xs[2] = {&kA, &kI};
  const L* ts[2] = {&kI, &t2};
  for (int e = 0; e < 2; e++) {
    double r[9];  // residual T − X·W, column-major
    for (int j = 0; j < 3; j++)
      for (int i = 0; i < 3; i++) {
        double o = 0;
        for (int k = 0; k < 3; k++) o += double(xs[e]->a[i + 3 * k]) * w[k + 3 * j];
        r[i + 3 * j] = double(ts[e]->a[i + 3 * j]) - o;
        expect_loss += r[i + 3 * j] * r[i + 3 * j];
      }
    for (int j = 0; j < 3; j++)
      for (int k = 0; k < 3; k++)
        for (int i = 0; i < 3; i++) dw[k + 3 * j] += -2.0 * double(xs[e]->a[i + 3 * k]) * r[i + 3 * j];
  }
  EXPECT(std::fabs(loss - expect_loss) <= 1e-5 * expect_loss);
  const L got = train.get_weights().second;
  for (int i = 0; i < 9; i++) EXPECT(std::fabs(got.a[i] - (w[i] + double(lambda) * dw[i])) <= 1e-5);
  EXPECT(throws(ErrorKind::Empty, [&] { apply_gradient_step(train, lambda, std::vector<decltype(train)::Input>{}, new_graph); }));  // net_ext.rs:72
}

// snapshots are values (net.rs:357-359 clones; `+=` rebinds): get_weights() taken before a step keeps the old contents
static void test_weight_snapshot_is_a_value() {
  auto train = make_net(kW0).add_square_loss<L>();
  const auto before = train.get_weights();
  apply_gradient_step(train, -0.01f, std::vector<decltype(train)::Input>{{kA, kI}}, new_graph);
  EXPECT(close(before.second, kW0, 0.0f));
  EXPECT(!close(train.get_weights().second, kW0, 0.0f));
}

// a net whose input is a VALUE (the previous layer's output): what `then` feeds its second net (net.rs:429-452)
struct Identity : NetBase<Identity> {
  using Input = host::Val<float>;
  using Weights = Unit;
  template <class A> auto eval_with_gradient_info(A&, const Input& v) const { return std::make_pair(v, Unit{}); }
  Weights get_weights() const { return {}; }
  void set_weights(const Weights&) {}
  void update_weights(const Weights&) {}
  template <class I, class R> Weights read_weight_gradients(const I&, const R&) const { return {}; }
};

static void test_combinators_and_weight_structures() {
  // then: a two-layer net, second layer = (previous output).using(weight).map(tanh ∘ matmul)
  auto layer1 = make_net(kW0);
  auto layer2 = Identity().using_(WeightData<L>(kI)).map([](auto& g, const auto& hw) { return g.tanh(g.matmul_nn(hw.first, hw.second)); });
  auto two = layer1.then(layer2);
  using W2 = decltype(two)::Weights;
  static_assert(std::is_same<W2, std::pair<std::pair<Unit, L>, std::pair<Unit, L>>>::value, "Then pairs the two nets' weights");
  host::Graph1 g;
  auto r = two.eval_with_gradient_info(g, kA);
  EXPECT(r.first.id().has_value());
  EXPECT(r.second.first.second.has_value() && r.second.second.second.has_value());           // both weights are variables of this tape
  EXPECT(r.second.first.second->index < r.second.second.second->index);                      // recorded in evaluation order
  auto loss = g.norm2(r.first);
  auto store = g.evaluate_gradients_once(loss.gid(), 1.0f);
  W2 grads = two.read_weight_gradients(r.second, store);
  EXPECT(grads.first.second.dims() == kW0.dims() && grads.second.second.dims() == kI.dims());
  // WeightOps recurse through the structure (net.rs:478-493)
  W2 scaled = weights_scale(grads, 2.0f);
  EXPECT(std::fabs(scaled.first.second.a[4] - 2.0f * grads.first.second.a[4]) <= 1e-6f);
  W2 sum = grads;
  weights_add_assign(sum, grads);
  EXPECT(close(sum.second.second, scaled.second.second, 1e-6f));
  EXPECT(close(grads.second.second, weights_scale(scaled, 0.5f).second.second, 1e-6f));      // `sum` was a copy: `grads` is untouched
  two.update_weights(weights_scale(grads, -0.01f));
  EXPECT(!close(two.get_weights().first.second, kW0, 0.0f));
  EXPECT(throws(ErrorKind::Dimensions, [&] { W2 bad = grads; bad.first.second = L({1, 2}, Dim4{2, 1}); weights_add_assign(sum, bad); }));
  // evaluated on Eval nothing is tracked: reading a gradient is MissingGradient (EmptyGradientMap, net.rs:110-116)
  host::EvalAlgebra eval;
  auto re = two.eval_with_gradient_info(eval, kA);
  EXPECT(!re.second.first.second.has_value());
  EXPECT(throws(ErrorKind::MissingGradient, [&] { two.read_weight_gradients(re.second, store); }));
  // and a variable the sweep never reached is MissingGradient too, not zeros
  host::Graph1 g2;
  auto r2 = two.eval_with_gradient_info(g2, kA);
  auto unrelated = g2.variable(kI);
  auto s2 = g2.evaluate_gradients_once(g2.norm2(unrelated).gid(), 1.0f);
  EXPECT(throws(ErrorKind::MissingGradient, [&] { two.read_weight_gradients(r2.second, s2); }));

  // and: a tuple of nets over a tuple of inputs (net.rs:563-629)
  auto both = make_net(kW0).and_(ConstantData<L>(kI));
  static_assert(std::is_same<decltype(both)::Input, std::tuple<L, Unit>>::value, "tuple input");
  host::Graph1 g3;
  auto r3 = both.eval_with_gradient_info(g3, std::make_tuple(kA, Unit{}));
  EXPECT(std::get<0>(r3.first).id().has_value() && !std::get<1>(r3.first).id().has_value());  // the constant stays a constant
  EXPECT(close(std::get<1>(r3.first).data(), kI, 0.0f));
  auto wt = both.get_weights();
  static_assert(std::is_same<decltype(wt), std::tuple<std::pair<Unit, L>, Unit>>::value, "tuple weights");
  both.update_weights(weights_scale(wt, 1.0f));  // W += W
  EXPECT(std::fabs(std::get<0>(both.get_weights()).second.a[0] - 2.0f * kW0.a[0]) <= 1e-6f);

  // Vec<N>: Lengths on every mismatch (net.rs:646, 661, 669, 681, 695)
  VecNet<WeightData<L>> vec({WeightData<L>(kW0), WeightData<L>(kI)});
  host::Graph1 g4;
  auto r4 = vec.eval_with_gradient_info(g4, {Unit{}, Unit{}});
  EXPECT(r4.first.size() == 2 && r4.second.size() == 2 && r4.second[0]->index == 1 && r4.second[1]->index == 2);
  EXPECT(throws(ErrorKind::Lengths, [&] { vec.eval_with_gradient_info(g4, {Unit{}}); }));
  EXPECT(throws(ErrorKind::Lengths, [&] { vec.set_weights({kW0}); }));
  EXPECT(throws(ErrorKind::Lengths, [&] { vec.update_weights({kW0, kI, kI}); }));
  auto prod = g4.mul(r4.first[0], r4.first[1]);
  auto s4 = g4.evaluate_gradients_once(g4.norm2(prod).gid(), 1.0f);
  EXPECT(throws(ErrorKind::Lengths, [&] { vec.read_weight_gradients(std::vector<std::optional<go::Id>>{r4.second[0]}, s4); }));
  std::vector<L> gv = vec.read_weight_gradients(r4.second, s4);
  EXPECT(gv.size() == 2);
  std::vector<L> shorter{gv[0]};
  EXPECT(throws(ErrorKind::Lengths, [&] { weights_add_assign(gv, shorter); }));
  const std::vector<uint8_t> blob = serialize_weights_bincode(vec.get_weights());
  EXPECT(blob.size() == 8 + 2 * (4 + 32 + 8 + 36));  // a Vec is length-prefixed
  HostMaker maker;
  auto back = deserialize_weights_bincode<std::vector<L>>(maker, blob);
  EXPECT(back.size() == 2 && close(back[1], kI, 0.0f));
}

// a corrupt checkpoint is an Error, never a crash or a wild allocation (the build runs under ASan + UBSan)
static void test_corrupt_checkpoints_are_errors() {
  using W = std::tuple<std::vector<L>, std::pair<Unit, host::Leaf<double>>>;
  const W w{std::vector<L>{kW0, kI, L({1, 2}, Dim4{2, 1})}, {Unit{}, host::Leaf<double>({0.25, -3.0, 8.0, 1e300}, Dim4{2, 2})}};
  const std::vector<uint8_t> good = serialize_weights_bincode(w);
  HostMaker maker;
  const W back = deserialize_weights_bincode<W>(maker, good);
  EXPECT(serialize_weights_bincode(back) == good);
  uint64_t state = 0x9E3779B97F4A7C15ull;
  auto next = [&] { state ^= state << 13; state ^= state >> 7; state ^= state << 17; return state; };
  int accepted = 0, rejected = 0;
  for (int trial = 0; trial < 4000; trial++) {
    std::vector<uint8_t> bad = good;
    switch (trial % 4) {
      case 0: bad.resize(next() % good.size()); break;                                             // truncated anywhere
      case 1: bad[next() % bad.size()] ^= uint8_t(1u << (next() % 8)); break;                      // one flipped bit
      case 2: for (int k = 0; k < 8; k++) bad[next() % bad.size()] = uint8_t(next()); break;       // a few random bytes
      default: bad.insert(bad.begin() + long(next() % bad.size()), size_t(1 + next() % 16), uint8_t(next())); break;  // inserted junk
    }
    try {
      const W got = deserialize_weights_bincode<W>(maker, bad);
      accepted++;  // (a flipped data bit is still a well-formed blob)
      EXPECT(serialize_weights_bincode(got) == bad);
    } catch (const Error& e) {
      rejected++;
      EXPECT(e.kind == ErrorKind::Invalid || e.kind == ErrorKind::Type);
    }
  }
  EXPECT(accepted > 0 && rejected > 1000);
}

int main(int argc, char** argv) {
  if (argc > 1 && !std::strcmp(argv[1], "shards")) {  // `total world` per line on stdin -> every rank's [first, last)
    unsigned long total, world;
    while (std::scanf("%lu %lu", &total, &world) == 2) {
      for (size_t r = 0; r < world; r++) {
        const auto b = shard_bounds(total, world, r);
        std::printf("%zu %zu ", b.first, b.second);
      }
      std::printf("\n");
    }
    return 0;
  }
  if (argc > 1 && !std::strcmp(argv[1], "hex")) {  // the checkpoint bytes of ((), W0) followed by those of [W0 (f32), 2x1 f64]
    for (uint8_t b : serialize_weights_bincode(std::make_pair(Unit{}, kW0))) std::printf("%02x", b);
    std::printf("\n");
    const host::Leaf<double> d({0.25, -3.0}, Dim4{2, 1});
    for (uint8_t b : serialize_weights_bincode(std::make_tuple(std::vector<L>{kW0, kI}, d))) std::printf("%02x", b);
    std::printf("\n");
    return 0;
  }
  try {
    test_testnet();
    test_make_net();
    test_gradient_step_closed_form();
    test_weight_snapshot_is_a_value();
    test_combinators_and_weight_structures();
    test_corrupt_checkpoints_are_errors();
  } catch (const Error& e) {
    std::printf("FAILED with gadcu::Error kind %d: %s\n", int(e.kind), e.what());
    return 2;
  } catch (const go::Error& e) {
    std::printf("FAILED with oracle error: %s\n", e.what());
    return 2;
  }
  std::printf(failures ? "%d expectation(s) failed\n" : "ok (%d failures)\n", failures);
  return failures ? 1 : 0;
}
