// gad/tests/net.rs replayed through the C++ mirror of gad's network layer (include/gadcu_net.hpp) on the device backend:
// a hand-written Net and make_net = input.using(weight).map(matmul_nn) trained to map a 3×3 matrix to the identity with
// DiffNet::apply_gradient_step (tests/net.rs:91-134), the trained weights handed to a fresh net through the reference's
// checkpoint bytes, evaluate() ≈ identity within 0.01, check() = the dims (tests/net.rs:136-151) — plus one gradient step
// over a two-example batch against its closed form, and a weight snapshot that must survive the update (copy-on-write).
// The same header is checked over the CPU oracle's tape without a GPU in tests/cpp/net_test.cc. Exit code 0 = all passed.
//   g++ -std=c++17 -Iinclude examples/net_cpp.cc -Lgad_b200/lib -lgadcu -Wl,-rpath,$PWD/gad_b200/lib -o examples/net_cpp
// `net_cpp dp <rank> <nranks> <id file>` is one rank of the data-parallel form of the step (one process per GPU, device =
// rank; rank 0 writes the communicator's id to the file, the others wait for it): the examples of a batch split over the
// ranks, update and loss summed — against the same steps taken by this process alone on the whole batch.
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <thread>

#include "gadcu_net.hpp"

using namespace gadcu;
using namespace gadcu::net;
using A = Array<float>;

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

static const std::vector<float> kA{1.0f, 2.0f, 1.0f, 1.0f, 0.0f, 1.0f, 0.0f, -2.0f, -1.0f};  // tests/net.rs:95-98, column-major
static const std::vector<float> kI{1, 0, 0, 0, 1, 0, 0, 0, 1};
static const std::vector<float> kW0{0.3f, -0.2f, 0.5f, 0.1f, 0.4f, -0.6f, -0.7f, 0.2f, 0.9f};
static A mat(Context& ctx, const std::vector<float>& v) { return A::from_host(ctx, v, Dim4{3, 3}); }
static bool close(const A& a, const std::vector<float>& b, float tol) {
  const std::vector<float> h = a.host();
  if (h.size() != b.size()) return false;
  for (size_t i = 0; i < h.size(); i++)
    if (!(std::fabs(h[i] - b[i]) <= tol)) return false;
  return true;
}

// tests/net.rs:8-78: a net written by hand against the five-method interface
class TestNet : public NetBase<TestNet> {
 public:
  using Input = A;
  using Weights = A;
  TestNet(Dim4 dims, A weights) : dims_(dims), weights_(std::move(weights)) {}
  template <class G> auto eval_with_gradient_info(G& g, const Input& input) const {
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
    if (!info) throw Error(ErrorKind::MissingGradient, "TestNet");
    auto g = store.get(*info);
    if (!g) throw Error(ErrorKind::MissingGradient, "TestNet");
    return g->data();
  }

 private:
  Dim4 dims_;
  A weights_;
};

// tests/net.rs:80-89
static auto make_net(const A& w0) {
  return InputData<A>(Dim4{3, 3}).using_(WeightData<A>(w0)).map([](auto& g, const auto& iw) { return g.matmul_nn(iw.first, iw.second); });
}

template <class Train, class Plain>
static void train_to_identity(Context& ctx, Train train, Plain fresh) {
  std::vector<typename Train::Input> samples{{mat(ctx, kA), mat(ctx, kI)}};
  float loss = 0, first = -1;
  int steps = 0;
  do {
    loss = apply_gradient_step(train, ctx, -0.01f, samples);
    EXPECT(std::isfinite(loss));
    if (first < 0) first = loss;
  } while (loss >= 1e-6f && ++steps < 20000);
  EXPECT(loss < 1e-6f && loss < first);
  // the weights travel as the reference's checkpoint bytes (tests/net.rs:136-143)
  const std::vector<uint8_t> bytes = serialize_weights_bincode(train.get_weights());
  fresh.set_weights(deserialize_weights_bincode<typename Plain::Weights>(ctx, bytes));
  EXPECT(close(fresh.evaluate(ctx, mat(ctx, kA)).data(), kI, 0.01f));
  EXPECT(fresh.check(mat(ctx, kA)) == (Dim4{3, 3}));  // tests/net.rs:147-151
}

static void test_gradient_step_closed_form(Context& ctx) {
  auto train = make_net(mat(ctx, kW0)).add_square_loss<A>();
  const std::vector<float> t2{0.5f, 0, 1, -1, 2, 0, 0, 1, 1};
  std::vector<decltype(train)::Input> batch{{mat(ctx, kA), mat(ctx, kI)}, {mat(ctx, kI), mat(ctx, t2)}};
  const auto snapshot = train.get_weights();
  const float lambda = -0.05f;
  const float loss = apply_gradient_step(train, ctx, lambda, batch);
  // out = X·W, loss = Σ ‖T − X·W‖², dL/dW = −2 Xᵀ(T − X·W), summed over the examples; ONE update (net_ext.rs:62-70)
  double expect_loss = 0, dw[9] = {0};
  const std::vector<float>* xs[2] = {&kA, &kI};
  const std::vector<float>* ts[2] = {&kI, &t2};
  for (int e = 0; e < 2; e++) {
    double r[9];
    for (int j = 0; j < 3; j++)
      for (int i = 0; i < 3; i++) {
        double o = 0;
        for (int k = 0; k < 3; k++) o += double((*xs[e])[i + 3 * k]) * double(kW0[k + 3 * j]);
        r[i + 3 * j] = double((*ts[e])[i + 3 * j]) - o;
        expect_loss += r[i + 3 * j] * r[i + 3 * j];
      }
    for (int j = 0; j < 3; j++)
      for (int k = 0; k < 3; k++)
        for (int i = 0; i < 3; i++) dw[k + 3 * j] += -2.0 * double((*xs[e])[i + 3 * k]) * r[i + 3 * j];
  }
  EXPECT(std::fabs(loss - expect_loss) <= 1e-5 * expect_loss);
  std::vector<float> expect_w(9);
  for (int i = 0; i < 9; i++) expect_w[i] = float(double(kW0[i]) + double(lambda) * dw[i]);
  EXPECT(close(train.get_weights().second, expect_w, 1e-5f));
  EXPECT(close(snapshot.second, kW0, 0.0f));  // a snapshot taken before the step is a value (net.rs:357-359; `+=` is copy-on-write)
  EXPECT(throws(ErrorKind::Empty, [&] { apply_gradient_step(train, ctx, lambda, std::vector<decltype(train)::Input>{}); }));
  EXPECT(throws(ErrorKind::Dimensions, [&] { train.check(std::make_pair(A::from_host(ctx, {1, 2}, Dim4{2, 1}), mat(ctx, kI))); }));
  // evaluated on Eval nothing is tracked: there is no gradient to read (EmptyGradientMap, net.rs:110-116)
  Eval eval(ctx);
  auto r = train.eval_with_gradient_info(eval, batch[0]);
  Graph1 g(ctx);
  auto v = g.variable(mat(ctx, kI));
  auto store = g.evaluate_gradients_once(g.norm2(v).gid(), 1.0f);
  EXPECT(throws(ErrorKind::MissingGradient, [&] { train.read_weight_gradients(r.second, store); }));
}

// gadcu::Comm + the data-parallel apply_gradient_step: uneven shares (5 examples), a rank without examples (1 example)
static int run_dp_rank(int rank, int nranks, const std::string& id_file) {
  Context ctx(rank);
  std::vector<uint8_t> id(128);
  if (rank == 0) {
    id = Comm::unique_id();
    std::ofstream(id_file + ".tmp", std::ios::binary).write(reinterpret_cast<const char*>(id.data()), 128);
    std::rename((id_file + ".tmp").c_str(), id_file.c_str());
  } else {
    for (int tries = 0;; tries++) {
      std::ifstream in(id_file, std::ios::binary);
      if (in && in.read(reinterpret_cast<char*>(id.data()), 128)) break;
      if (tries > 600) { std::printf("rank %d: no communicator id after 60 s\n", rank); return 3; }
      std::this_thread::sleep_for(std::chrono::milliseconds(100));
    }
  }
  Comm comm(ctx, nranks, rank, id);
  const std::vector<float> t2{0.5f, 0, 1, -1, 2, 0, 0, 1, 1}, x3{0, 1, 0, 2, 0, -1, 1, 1, 0.5f};
  auto example = [&](const std::vector<float>& x, const std::vector<float>& t) { return std::make_pair(mat(ctx, x), mat(ctx, t)); };
  for (size_t batch_size : {size_t(5), size_t(1)}) {
    auto together = make_net(mat(ctx, kW0)).add_square_loss<A>();  // stepped over the ranks
    auto alone = make_net(mat(ctx, kW0)).add_square_loss<A>();     // stepped by this process on the whole batch
    std::vector<decltype(together)::Input> batch{example(kA, kI), example(kI, t2), example(x3, kI), example(kA, t2), example(x3, t2)};
    batch.resize(batch_size, example(kA, kI));
    for (int step = 0; step < 5; step++) {
      const float l_dp = apply_gradient_step(together, ctx, -0.01f, batch, comm);
      const float l_one = apply_gradient_step(alone, ctx, -0.01f, batch);
      EXPECT(std::fabs(l_dp - l_one) <= 1e-5f * std::fabs(l_one));
    }
    EXPECT(close(together.get_weights().second, alone.get_weights().second.host(), 1e-5f));
    EXPECT(!close(together.get_weights().second, kW0, 1e-3f));  // (it did move)
  }
  // a plain in-place sum: every rank contributes rank + 1
  A v = A::from_host(ctx, std::vector<float>(9, float(rank + 1)), Dim4{3, 3});
  comm.allreduce_sum<float>({&v});
  EXPECT(close(v, std::vector<float>(9, float(nranks * (nranks + 1) / 2)), 0.0f));
  ctx.synchronize();
  std::printf(failures ? "rank %d: %d expectation(s) failed\n" : "rank %d: data-parallel net steps ok (%d failures)\n", rank, failures);
  return failures ? 1 : 0;
}

int main(int argc, char** argv) {
  try {
    if (argc == 5 && !std::strcmp(argv[1], "dp")) return run_dp_rank(std::atoi(argv[2]), std::atoi(argv[3]), argv[4]);
    Context ctx(0);
    train_to_identity(ctx, TestNet(Dim4{3, 3}, mat(ctx, kW0)).add_square_loss<A>(), TestNet(Dim4{3, 3}, mat(ctx, kI)));
    train_to_identity(ctx, make_net(mat(ctx, kW0)).add_square_loss<A>(), make_net(mat(ctx, kI)));
    test_gradient_step_closed_form(ctx);
  } catch (const Error& e) {
    std::printf("FAILED with gadcu::Error kind %d: %s\n", int(e.kind), e.what());
    return 2;
  }
  std::printf(failures ? "%d expectation(s) failed\n" : "all net tests passed (%d failures)\n", failures);
  return failures ? 1 : 0;
}
