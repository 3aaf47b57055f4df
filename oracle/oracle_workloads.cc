// ORACLE — TEST INFRASTRUCTURE ONLY (see host_array.hpp). The BASELINE.json workloads expressed as
// gad programs over the oracle's Graph1 tape, used (a) as the parity reference for the CUDA path
// on the same seeded inputs and (b) as the reported CPU baseline ("port") in bench.py.
//   C1/C2  Rosenbrock-N scalar tape, one tape per lane, f64
//   C3     sum(exp(x)*log(y) + x), f32 arrays, reverse sweep
//   C4     3-layer tanh MLP with square loss, f32, column-major
#include <dlfcn.h>

#include <chrono>
#include <thread>

#include "gad_oracle.hpp"

using namespace gad_oracle;

namespace {

// variables first (ids 1..N), then 8 nodes per term, then add_all.
template <class G>
Value<double> rosenbrock(G& g, const std::vector<Value<double>>& x) {
  std::vector<Value<double>> terms;
  size_t n = x.size();
  terms.reserve(n - 1);
  for (size_t i = 0; i + 1 < n; i++) {
    auto sq = g.mul(x[i], x[i]);
    auto d = g.sub(x[i + 1], sq);
    auto d2 = g.mul(d, d);
    auto t1 = g.mulc(d2, int16_t(100));
    auto nx = g.neg(x[i]);
    auto e = g.addc(nx, int16_t(1));
    auto e2 = g.mul(e, e);
    terms.push_back(g.add(t1, e2));
  }
  std::vector<const Value<double>*> ptrs;
  for (auto& t : terms) ptrs.push_back(&t);
  return g.template add_all<double>(ptrs);
}

void rosenbrock_lanes(int n, size_t lanes, size_t lo, size_t hi, const double* x, double* f, double* grad) {
  std::vector<Value<double>> xs(n);
  for (size_t lane = lo; lane < hi; lane++) {
    Graph1 g;  // a fresh tape per lane, as apply_gradient_step does per example (net_ext.rs:52)
    for (int i = 0; i < n; i++) xs[i] = g.variable<double>(x[size_t(i) * lanes + lane]);
    auto fv = rosenbrock(g, xs);
    if (f) f[lane] = fv.data;
    auto store = g.evaluate_gradients_once<double>(fv.gid(), 1.0);
    for (int i = 0; i < n; i++) grad[size_t(i) * lanes + lane] = *store.get<double>(xs[i].gid());
  }
}

// A cache-friendlier column-major GEMM for the CPU baseline: for each output column, axpy over k.
// Every C[m,n] still accumulates its products in increasing k, in T, starting from 0 — the same
// rounding sequence as ha::matmul's inner loop — so results are bit-identical to the plain loop.
template <class T>
void gemm_cols(const T* A, const T* B, T* C, uint64_t M, uint64_t N, uint64_t K, bool ta, bool tb, uint64_t lda,
               uint64_t ldb, uint64_t n0, uint64_t n1) {
  for (uint64_t n = n0; n < n1; n++) {
    T* c = C + M * n;
    for (uint64_t m = 0; m < M; m++) c[m] = T(0);
    if (!ta) {
      for (uint64_t k = 0; k < K; k++) {
        T b = tb ? B[n + ldb * k] : B[k + ldb * n];
        const T* a = A + lda * k;
        for (uint64_t m = 0; m < M; m++) c[m] += a[m] * b;
      }
    } else {
      for (uint64_t m = 0; m < M; m++) {
        const T* a = A + lda * m;
        T acc = T(0);
        if (!tb) {
          const T* b = B + ldb * n;
          for (uint64_t k = 0; k < K; k++) acc += a[k] * b[k];
        } else {
          for (uint64_t k = 0; k < K; k++) acc += a[k] * B[n + ldb * k];
        }
        c[m] = acc;
      }
    }
  }
}

int g_threads = 1;

// Optional BLAS for the CPU BASELINE only (never for parity): the reference's arrays are ArrayFire's, whose CPU backend
// hands af::matmul (gad/src/matrix.rs:53) to a BLAS sgemm, so a baseline timed with the triple loop above understates
// the reference's CPU path. oracle_set_blas(path) dlopens a CBLAS (the OpenBLAS numpy ships: ILP64 symbols with the
// scipy_ prefix; or a plain LP64 libopenblas / libcblas) and FastEval::matmul then calls sgemm. Results differ from the
// loop's in summation order, which is why parity checks never turn this on.
using sgemm64_fn = void (*)(int, int, int, int64_t, int64_t, int64_t, float, const float*, int64_t, const float*, int64_t, float, float*, int64_t);
using sgemm32_fn = void (*)(int, int, int, int, int, int, float, const float*, int, const float*, int, float, float*, int);
sgemm64_fn g_sgemm64 = nullptr;
sgemm32_fn g_sgemm32 = nullptr;
void (*g_blas_set_threads)(int) = nullptr;

bool blas_sgemm(const float* A, const float* B, float* C, uint64_t M, uint64_t N, uint64_t K, bool ta, bool tb, uint64_t lda, uint64_t ldb) {
  constexpr int kColMajor = 102, kNoTrans = 111, kTrans = 112;
  if (g_sgemm64) {
    g_sgemm64(kColMajor, ta ? kTrans : kNoTrans, tb ? kTrans : kNoTrans, int64_t(M), int64_t(N), int64_t(K), 1.0f, A, int64_t(lda), B, int64_t(ldb), 0.0f, C, int64_t(M));
    return true;
  }
  if (g_sgemm32) {
    g_sgemm32(kColMajor, ta ? kTrans : kNoTrans, tb ? kTrans : kNoTrans, int(M), int(N), int(K), 1.0f, A, int(lda), B, int(ldb), 0.0f, C, int(M));
    return true;
  }
  return false;
}

// Eval with a threaded matmul, everything else as the oracle's Eval. Same arithmetic, same order.
struct FastEval : Eval {
  template <class D> const D& link(const Value<D>& v) { return v.data; }
  using Eval::matmul;
  HostArray<float> matmul(const HostArray<float>& a, const HostArray<float>& b, MatProp p1, MatProp p2) {
    Dim4 rd = check_.matmul(a.dims(), b.dims(), p1, p2);
    if (rd[2] != 1 || rd[3] != 1) return ha::matmul<float>(a, b, p1.transposed, p2.transposed, rd);
    uint64_t M = rd[0], N = rd[1], K = p1.transposed ? a.dims()[0] : a.dims()[1];
    HostArray<float> r(rd);
    if (blas_sgemm(a.data(), b.data(), r.data(), M, N, K, p1.transposed, p2.transposed, a.dims()[0], b.dims()[0])) return r;
    int nt = std::max(1, std::min<int>(g_threads, int(N)));
    std::vector<std::thread> th;
    for (int t = 0; t < nt; t++) {
      uint64_t n0 = N * t / nt, n1 = N * (t + 1) / nt;
      th.emplace_back([&, n0, n1] {
        gemm_cols<float>(a.data(), b.data(), r.data(), M, N, K, p1.transposed, p2.transposed, a.dims()[0], b.dims()[0], n0, n1);
      });
    }
    for (auto& t : th) t.join();
    return r;
  }
  HostArray<float> matmul_nn(const HostArray<float>& a, const HostArray<float>& b) { return matmul(a, b, {}, {}); }
};
using FastGraph1 = Graph<Config1<FastEval>>;
using FastGraphN = Graph<ConfigN<FastEval>>;

// the C4 forward on any graph over HostArray<float>
template <class G>
Value<float> mlp_forward(G& g, uint64_t B, uint64_t D, int L, const float* X, const float* const* W, const float* const* b,
                                    const float* target, std::vector<Value<HostArray<float>>>& Ws, std::vector<Value<HostArray<float>>>& bs) {
  auto h = g.constant(HostArray<float>(Dim4(B, D), X));
  for (int l = 0; l < L; l++) {
    Ws.push_back(g.variable(HostArray<float>(Dim4(D, D), W[l])));
    bs.push_back(g.variable(HostArray<float>(Dim4(1, D), b[l])));
  }
  for (int l = 0; l < L; l++) {
    auto z = g.matmul_nn(h, Ws[l]);
    auto bt = g.tile_as(bs[l], Dim4(B, D));
    auto zb = g.add(z, bt);
    h = (l + 1 < L) ? g.tanh(zb) : zb;
  }
  auto t = g.constant(HostArray<float>(Dim4(B, D), target));
  auto delta = g.sub(t, h);
  return g.norm2(delta);
}

}  // namespace

extern "C" {

// x, grad: SoA [n][lanes] doubles; f (optional): [lanes]. Returns seconds spent.
double oracle_rosenbrock_batch(int n, size_t lanes, const double* x, double* f, double* grad, int nthreads) {
  auto t0 = std::chrono::steady_clock::now();
  nthreads = std::max(1, nthreads);
  std::vector<std::thread> th;
  for (int t = 0; t < nthreads; t++) {
    size_t lo = lanes * t / nthreads, hi = lanes * (t + 1) / nthreads;
    th.emplace_back([=] { rosenbrock_lanes(n, lanes, lo, hi, x, f, grad); });
  }
  for (auto& t : th) t.join();
  return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

// C3: a=exp(x); b=log(y); c=mul(a,b); d=add(c,x); s=sum_as(d,[1]); z=as_scalar(s); sweep from z
// with seed 1.0. Outputs z, dz/dx, dz/dy. Returns seconds spent in forward+sweep.
double oracle_c3(size_t n, const float* x, const float* y, float* z, float* gx, float* gy, double* sweep_seconds) {
  auto t0 = std::chrono::steady_clock::now();
  Graph1 g;
  Dim4 d(n);
  auto vx = g.variable(HostArray<float>(d, x));
  auto vy = g.variable(HostArray<float>(d, y));
  auto a = g.exp(vx);
  auto b = g.log(vy);
  auto c = g.mul(a, b);
  auto dd = g.add(c, vx);
  auto s = g.sum_as(dd, Dim4(1));
  auto zz = g.as_scalar(s);
  *z = zz.data;
  auto t1 = std::chrono::steady_clock::now();
  auto store = g.evaluate_gradients_once<float>(zz.gid(), 1.0f);
  auto t2 = std::chrono::steady_clock::now();
  std::memcpy(gx, store.get<HostArray<float>>(vx.gid())->data(), n * sizeof(float));
  std::memcpy(gy, store.get<HostArray<float>>(vy.gid())->data(), n * sizeof(float));
  if (sweep_seconds) *sweep_seconds = std::chrono::duration<double>(t2 - t1).count();
  return std::chrono::duration<double>(t2 - t0).count();
}

// C4: L-layer MLP, h = tanh(add(matmul_nn(h, W_l), tile_as(b_l,[B,D]))) for hidden layers, last
// layer linear; loss = norm2(sub(target, out)); sweep seed 1.0.
//   X, target: [B, D] column-major; W: L arrays [D, D]; b: L arrays [1, D].
//   gW, gb: outputs (same shapes). Returns seconds (forward + sweep).
double oracle_mlp_step(uint64_t B, uint64_t D, int L, const float* X, const float* const* W, const float* const* b,
                       const float* target, float* loss, float* const* gW, float* const* gb, int nthreads) {
  g_threads = std::max(1, nthreads);
  auto t0 = std::chrono::steady_clock::now();
  FastGraph1 g;
  std::vector<Value<HostArray<float>>> Ws, bs;
  auto lv = mlp_forward(g, B, D, L, X, W, b, target, Ws, bs);
  *loss = lv.data;
  auto store = g.evaluate_gradients_once<float>(lv.gid(), 1.0f);
  for (int l = 0; l < L; l++) {
    std::memcpy(gW[l], store.get<HostArray<float>>(Ws[l].gid())->data(), D * D * sizeof(float));
    std::memcpy(gb[l], store.get<HostArray<float>>(bs[l].gid())->data(), D * sizeof(float));
  }
  return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

// C5: Hessian-vector product of the C4 loss by reverse-over-reverse on GraphN (graph.rs:307-318):
//   g = compute_gradients(L, 1); s = add_all_l dot(g_{W_l}, v_l); hv = compute_gradients(s, 1).
// The second sweep differentiates through the recorded backward GEMMs, whose operands are transposed: with the
// reference's matmul VJP as written that is a Dimensions error for B != D, so this workload runs with
// the corrected adjoint switched on for its duration (and says so in its name). hv: L arrays [D, D]. Returns seconds.
double oracle_mlp_hvp_corrected(uint64_t B, uint64_t D, int L, const float* X, const float* const* W, const float* const* b,
                                const float* target, const float* const* v, float* loss, float* const* hv, int nthreads,
                                uint64_t* tape_nodes) {
  g_threads = std::max(1, nthreads);
  const bool before = corrected_matmul_adjoint();
  corrected_matmul_adjoint() = true;
  auto t0 = std::chrono::steady_clock::now();
  try {
    FastGraphN g;
    std::vector<Value<HostArray<float>>> Ws, bs;
    auto lv = mlp_forward(g, B, D, L, X, W, b, target, Ws, bs);
    *loss = lv.data;
    auto one = Value<float>::constant(1.0f);
    auto first = g.compute_gradients<float>(lv.gid(), one);
    std::vector<Value<float>> terms;
    for (int l = 0; l < L; l++) {
      const auto* gw = first.get<Value<HostArray<float>>>(Ws[l].gid());
      auto vl = g.constant(HostArray<float>(Dim4(D, D), v[l]));
      terms.push_back(g.dot(*gw, vl));
    }
    std::vector<const Value<float>*> ptrs;
    for (auto& t : terms) ptrs.push_back(&t);
    auto s = g.template add_all<float>(ptrs);
    auto second = g.compute_gradients<float>(s.gid(), one);
    for (int l = 0; l < L; l++)
      std::memcpy(hv[l], second.get<Value<HostArray<float>>>(Ws[l].gid())->data.data(), D * D * sizeof(float));
    if (tape_nodes) *tape_nodes = g.num_nodes();
  } catch (...) {
    corrected_matmul_adjoint() = before;
    throw;
  }
  corrected_matmul_adjoint() = before;
  return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

// CPU baseline only: route FastEval's matmul through a CBLAS sgemm (see blas_sgemm). path == NULL switches back to the
// loop. Returns 1 if a usable sgemm was found, 0 otherwise (the loop stays).
int oracle_set_blas(const char* path, int threads) {
  g_sgemm64 = nullptr;
  g_sgemm32 = nullptr;
  g_blas_set_threads = nullptr;
  if (!path) return 0;
  void* lib = dlopen(path, RTLD_NOW | RTLD_LOCAL);
  if (!lib) return 0;
  if (void* f = dlsym(lib, "scipy_cblas_sgemm64_")) {
    g_sgemm64 = reinterpret_cast<sgemm64_fn>(f);
    g_blas_set_threads = reinterpret_cast<void (*)(int)>(dlsym(lib, "scipy_openblas_set_num_threads64_"));
  } else if (void* f64 = dlsym(lib, "cblas_sgemm64_")) {
    g_sgemm64 = reinterpret_cast<sgemm64_fn>(f64);
    g_blas_set_threads = reinterpret_cast<void (*)(int)>(dlsym(lib, "openblas_set_num_threads64_"));
  } else if (void* f32 = dlsym(lib, "cblas_sgemm")) {
    g_sgemm32 = reinterpret_cast<sgemm32_fn>(f32);
    g_blas_set_threads = reinterpret_cast<void (*)(int)>(dlsym(lib, "openblas_set_num_threads"));
  }
  if (g_blas_set_threads && threads > 0) g_blas_set_threads(threads);
  return (g_sgemm64 || g_sgemm32) ? 1 : 0;
}

}  // extern "C"
