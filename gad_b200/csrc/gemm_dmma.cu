// gemm_dmma.cu — f64 GEMM on the tensor cores (north-star subsystem 3: "DMMA for f64"; tcgen05 has no f64 kind).
//
// Warp-level mma.sync.m8n8k4.f64 (DMMA): each instruction multiplies an 8x4 by a 4x8 f64 fragment into an 8x8 f64
// accumulator held in registers (2 per thread). Block tile 128 x 128 x 16, 8 warps as 2 (M) x 4 (N), a warp owns
// 64 x 32 = 8 x 4 DMMA tiles (64 accumulator doubles per thread). Operand tiles are staged k-major in shared memory
// ([k][m] and [k][n], padded) whatever the storage order, so column-major operands with any MatProp (forward, and the
// transposed products of the VJPs, matrix.rs:167,172) take the same inner loop; the next tile is fetched into registers
// while the current one is multiplied (register double buffering). Any M, N, K (edges are zero-filled / predicated),
// batch with broadcast strides, `C + acc` epilogue.
// Per output element the k-sum runs in increasing k, 4 at a time, in f64 (each DMMA rounds its 4-term dot product
// into the accumulator) — documented for the f64 tolerance in tests/test_gpu_ops.py.
#include "kernels.h"

namespace gadcu {
namespace k {
namespace {

constexpr int BM = 128, BN = 128, BK = 16, PAD = 8, kThreads = 256;
constexpr int LD = BM + PAD;  // 136 doubles: fragment loads of a warp fall in two conflict-free wavefronts

__device__ __forceinline__ void dmma_m8n8k4(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};" : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}

// one operand tile (128 x 16) from global memory into registers: 8 doubles per thread, laid out for store_tile
template <bool K_CONTIGUOUS>
__device__ __forceinline__ void fetch_tile(const double* __restrict__ X, uint64_t ld, uint64_t mn0, uint64_t k0, uint64_t MN, uint64_t K,
                                           double (&r)[8]) {
  const int tid = threadIdx.x;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    const int e = tid + i * kThreads;  // 0 .. 2047
    int mm, kk;
    if (K_CONTIGUOUS) { kk = e % BK; mm = e / BK; } else { mm = e % BM; kk = e / BM; }
    const uint64_t mn = mn0 + mm, k = k0 + kk;
    double v = 0.0;
    if (mn < MN && k < K) v = K_CONTIGUOUS ? X[k + ld * mn] : X[mn + ld * k];
    r[i] = v;
  }
}
template <bool K_CONTIGUOUS>
__device__ __forceinline__ void store_tile(double* __restrict__ S, const double (&r)[8]) {
  const int tid = threadIdx.x;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    const int e = tid + i * kThreads;
    int mm, kk;
    if (K_CONTIGUOUS) { kk = e % BK; mm = e / BK; } else { mm = e % BM; kk = e / BM; }
    S[kk * LD + mm] = r[i];
  }
}

// TA: op(A) = A^T, i.e. A is stored [K, M] (k contiguous). TB: op(B) = B^T, B stored [N, K] (n contiguous).
template <bool TA, bool TB>
__global__ void __launch_bounds__(kThreads) gemm_dmma_kernel(const double* __restrict__ A, const double* __restrict__ B, double* C, uint64_t M,
                                                             uint64_t N, uint64_t K, uint64_t lda, uint64_t ldb, uint64_t strideA,
                                                             uint64_t strideB, uint64_t strideC, bool accumulate) {
  extern __shared__ double smem[];
  double* As = smem;                 // [2][BK][LD]
  double* Bs = smem + 2 * BK * LD;   // [2][BK][LD]
  const uint64_t batch = blockIdx.z;
  A += batch * strideA;
  B += batch * strideB;
  C += batch * strideC;
  const uint64_t m0 = uint64_t(blockIdx.x) * BM, n0 = uint64_t(blockIdx.y) * BN;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wm = (warp & 1) * 64, wn = (warp >> 1) * 32;  // this warp's 64 x 32 corner of the block tile
  const int g = lane >> 2, t = lane & 3;                    // DMMA fragment coordinates

  double acc[8][4][2];
#pragma unroll
  for (int i = 0; i < 8; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) acc[i][j][0] = acc[i][j][1] = 0.0;

  double ra[8], rb[8];
  fetch_tile<TA>(A, lda, m0, 0, M, K, ra);
  fetch_tile<!TB>(B, ldb, n0, 0, N, K, rb);
  store_tile<TA>(As, ra);
  store_tile<!TB>(Bs, rb);
  __syncthreads();
  const uint64_t ktiles = (K + BK - 1) / BK;
  for (uint64_t kt = 0; kt < ktiles; kt++) {
    const int cur = int(kt & 1);
    const bool more = kt + 1 < ktiles;
    if (more) {
      fetch_tile<TA>(A, lda, m0, (kt + 1) * BK, M, K, ra);
      fetch_tile<!TB>(B, ldb, n0, (kt + 1) * BK, N, K, rb);
    }
    const double* as = As + cur * BK * LD;
    const double* bs = Bs + cur * BK * LD;
#pragma unroll
    for (int kk = 0; kk < BK; kk += 4) {
      double a[8], b[4];
#pragma unroll
      for (int i = 0; i < 8; i++) a[i] = as[(kk + t) * LD + wm + i * 8 + g];
#pragma unroll
      for (int j = 0; j < 4; j++) b[j] = bs[(kk + t) * LD + wn + j * 8 + g];
#pragma unroll
      for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) dmma_m8n8k4(acc[i][j], a[i], b[j]);
    }
    if (more) {
      store_tile<TA>(As + (cur ^ 1) * BK * LD, ra);
      store_tile<!TB>(Bs + (cur ^ 1) * BK * LD, rb);
    }
    __syncthreads();
  }
  // accumulator fragment: rows g, columns 2t and 2t+1 of each 8x8 tile
#pragma unroll
  for (int i = 0; i < 8; i++) {
    const uint64_t m = m0 + wm + i * 8 + g;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; j++) {
#pragma unroll
      for (int c = 0; c < 2; c++) {
        const uint64_t n = n0 + wn + j * 8 + 2 * t + c;
        if (n >= N) continue;
        double v = acc[i][j][c];
        if (accumulate) v = C[m + M * n] + v;  // store.rs:52 `current + new`
        C[m + M * n] = v;
      }
    }
  }
}

constexpr size_t kSmem = size_t(4) * BK * LD * sizeof(double);  // 69 632 B

template <bool TA, bool TB>
void launch(gadcu_ctx* ctx, const GemmArgs& g, dim3 grid) {
  static std::once_flag once;
  std::call_once(once, [] { cudaFuncSetAttribute(gemm_dmma_kernel<TA, TB>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kSmem)); });
  gemm_dmma_kernel<TA, TB><<<grid, kThreads, kSmem, ctx->stream>>>(static_cast<const double*>(g.A), static_cast<const double*>(g.B),
                                                                  static_cast<double*>(g.C), g.M, g.N, g.K, g.lda, g.ldb, g.strideA,
                                                                  g.strideB, g.strideC, g.accumulate);
}

}  // namespace

bool gemm_dmma_supported(const GemmArgs& g) {
  static const bool off = [] {
    const char* e = getenv("GADCU_GEMM");
    return e && std::string(e) == "simt";
  }();
  // tiny products (the reference's 4x3 * 3x5 tests) stay on the CUDA cores: a 128 x 128 tile would be almost empty
  return !off && g.dt == GADCU_F64 && g.M * g.N >= 64 * 64 && g.K >= 4;
}

void launch_gemm_dmma(gadcu_ctx* ctx, const GemmArgs& g) {
  dim3 grid(unsigned((g.M + BM - 1) / BM), unsigned((g.N + BN - 1) / BN), unsigned(g.batch));
  if (grid.y > 65535 || grid.z > 65535) fail(GADCU_ERR_UNSUPPORTED, "gemm: grid too large");
  if (g.ta && g.tb) launch<true, true>(ctx, g, grid);
  else if (g.ta) launch<true, false>(ctx, g, grid);
  else if (g.tb) launch<false, true>(ctx, g, grid);
  else launch<false, false>(ctx, g, grid);
  ctx->count_launch();
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) fail(GADCU_ERR_CUDA, std::string("gemm_dmma launch: ") + cudaGetErrorString(err));
}

}  // namespace k
}  // namespace gadcu
