// gemm_simt.cu — general column-major GEMM on the CUDA cores: any shape, any transposition, batch
// broadcast, f32 and f64. It is the path for shapes the tensor-core kernel does not take (ragged
// sizes, f64, tiny matrices such as the reference's 4x3 * 3x5 tests); the MLP-sized f32 products go
// to gemm_tc.cu (tcgen05, 3xTF32). Per output element the k-sum runs in increasing k in type T.
#include "kernels.h"

namespace gadcu {
namespace k {

void launch_gemm_tc(gadcu_ctx* ctx, const GemmArgs& g);    // gemm_tc.cu
void launch_gemm_dmma(gadcu_ctx* ctx, const GemmArgs& g);  // gemm_dmma.cu
bool gemm_dmma_supported(const GemmArgs& g);

namespace {

constexpr int BM = 64, BN = 64, BK = 16, TM = 4, TN = 4;

template <class T>
__global__ void __launch_bounds__(256) gemm_simt_kernel(const T* __restrict__ A, const T* __restrict__ B, T* C, uint64_t M,
                                                        uint64_t N, uint64_t K, bool ta, bool tb, uint64_t lda, uint64_t ldb,
                                                        uint64_t strideA, uint64_t strideB, uint64_t strideC, bool accumulate) {
  __shared__ T As[BK][BM + 4];
  __shared__ T Bs[BK][BN + 4];
  const uint64_t batch = blockIdx.z;
  A += batch * strideA;
  B += batch * strideB;
  C += batch * strideC;
  const uint64_t m0 = uint64_t(blockIdx.x) * BM, n0 = uint64_t(blockIdx.y) * BN;
  const int tid = threadIdx.x;
  const int tx = tid % 16, ty = tid / 16;
  T acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; i++)
#pragma unroll
    for (int j = 0; j < TN; j++) acc[i][j] = T(0);

  for (uint64_t k0 = 0; k0 < K; k0 += BK) {
    // A tile: BM x BK
    for (int e = tid; e < BM * BK; e += 256) {
      int mm, kk;
      if (!ta) { mm = e % BM; kk = e / BM; } else { kk = e % BK; mm = e / BK; }
      uint64_t m = m0 + mm, k = k0 + kk;
      T v = T(0);
      if (m < M && k < K) v = ta ? A[k + lda * m] : A[m + lda * k];
      As[kk][mm] = v;
    }
    for (int e = tid; e < BN * BK; e += 256) {
      int nn, kk;
      if (!tb) { kk = e % BK; nn = e / BK; } else { nn = e % BN; kk = e / BN; }
      uint64_t n = n0 + nn, k = k0 + kk;
      T v = T(0);
      if (n < N && k < K) v = tb ? B[n + ldb * k] : B[k + ldb * n];
      Bs[kk][nn] = v;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; kk++) {
      T a[TM], b[TN];
#pragma unroll
      for (int i = 0; i < TM; i++) a[i] = As[kk][tx * TM + i];
#pragma unroll
      for (int j = 0; j < TN; j++) b[j] = Bs[kk][ty * TN + j];
#pragma unroll
      for (int i = 0; i < TM; i++)
#pragma unroll
        for (int j = 0; j < TN; j++) acc[i][j] += a[i] * b[j];
    }
    __syncthreads();
  }
#pragma unroll
  for (int j = 0; j < TN; j++) {
    uint64_t n = n0 + ty * TN + j;
    if (n >= N) continue;
#pragma unroll
    for (int i = 0; i < TM; i++) {
      uint64_t m = m0 + tx * TM + i;
      if (m >= M) continue;
      T v = acc[i][j];
      if (accumulate) v = C[m + M * n] + v;
      C[m + M * n] = v;
    }
  }
}

}  // namespace

void launch_gemm(gadcu_ctx* ctx, const GemmArgs& g) {
  ProfileScope prof_scope(ctx, PROF_GEMM);
  if (g.M * g.N * g.batch == 0) return;
  if (gemm_tc_supported(g)) {
    launch_gemm_tc(ctx, g);
    return;
  }
  // af::matmul over batch dims 2, 3 with broadcast (gad/src/matrix.rs:108-114): when one matrix of the batch is a product
  // the tensor-core kernel takes, the batch is that kernel once per matrix (a broadcast operand — stride 0 — keeps its
  // TF32 split across the launches through the buffer's cache)
  if (g.batch > 1 && !g.push && !g.epi && g.batch <= 4096) {
    GemmArgs one = g;
    one.batch = 1;
    if (gemm_tc_supported(one)) {
      const size_t es = 4;
      for (uint64_t b = 0; b < g.batch; b++) {
        one.A = static_cast<const char*>(g.A) + b * g.strideA * es;
        one.B = static_cast<const char*>(g.B) + b * g.strideB * es;
        one.C = static_cast<char*>(g.C) + b * g.strideC * es;
        one.a_buf = g.strideA == 0 ? g.a_buf : nullptr;  // (a slice of a batched operand is not a whole buffer)
        one.b_buf = g.strideB == 0 ? g.b_buf : nullptr;
        launch_gemm_tc(ctx, one);
      }
      return;
    }
  }
  if (g.push) fail(GADCU_ERR_UNSUPPORTED, "gemm: the tile-pushing epilogue exists on the f32 tensor-core path only");
  if (gemm_dmma_supported(g)) {
    launch_gemm_dmma(ctx, g);
    return;
  }
  dim3 grid(unsigned((g.M + BM - 1) / BM), unsigned((g.N + BN - 1) / BN), unsigned(g.batch));
  if (grid.y > 65535 || grid.z > 65535) fail(GADCU_ERR_UNSUPPORTED, "gemm: grid too large");
  if (g.dt == GADCU_F32)
    gemm_simt_kernel<float><<<grid, 256, 0, ctx->stream>>>((const float*)g.A, (const float*)g.B, (float*)g.C, g.M, g.N, g.K, g.ta,
                                                           g.tb, g.lda, g.ldb, g.strideA, g.strideB, g.strideC, g.accumulate);
  else
    gemm_simt_kernel<double><<<grid, 256, 0, ctx->stream>>>((const double*)g.A, (const double*)g.B, (double*)g.C, g.M, g.N, g.K,
                                                            g.ta, g.tb, g.lda, g.ldb, g.strideA, g.strideB, g.strideC, g.accumulate);
  ctx->count_launch();
}

}  // namespace k
}  // namespace gadcu
