// kernels.h — host-side launchers of the hand-written sm_100a kernels (definitions in *.cu).
#pragma once
#include "core.h"

namespace gadcu {
namespace k {

// Elementwise functors. The first group are gad's Eval primitives (one per `Eval for af::Array`
// method); the VJP_* group are the fused reverse-sweep bodies of SURVEY App. A: each
// performs, per element and with a rounding after every step, exactly the primitive sequence the
// reference's VJP closure issues, then (optionally) the `current + new` accumulation of
// GradientStore::add_gradient (gad/src/store.rs:52) — one pass over HBM instead of up to six.
enum class EwOp : int {
  // binary
  ADD, SUB, MUL, DIV, POW, MIN, MAX,
  // unary
  NEG, EXP, LOG, LOG1P, SIN, COS, TANH, SIGMOID, RECIPROCAL, SQRT, COPY,
  // with a host constant (args.c)
  ADDC, MULC, POWC,
  // in0 * in1 where in1 is a broadcast device scalar; also plain a*lambda via MULC
  SCALE,
  // select by v0 >= v1 (in0=v0, in1=v1, then r0 and/or r1)
  SELECT_BOTH, SELECT_R0, SELECT_R1,
  // fused VJPs: in0 = incoming gradient g, then the saved forward operands
  VJP_NEG,         // 0 - g
  VJP_MUL,         // g * other            (arith.rs:214-223)
  VJP_MULC,        // g * c
  VJP_POWC,        // (pow(v, c-1) * c) * g (const_arith.rs:172-176); args.c2 = c-1 as computed in C
  VJP_EXP,         // g * exp(v)
  VJP_LOG,         // g / v
  VJP_LOG1P,       // g / (v + 1)
  VJP_SIN,         // g * cos(v)
  VJP_COS,         // g * (0 - sin(v))
  VJP_TANH,        // g * ((0 - t*t) + 1)
  VJP_SIGMOID,     // g * (c * ((0 - c) + 1))
  VJP_RECIPROCAL,  // g * (1 / (0 - v*v))
  VJP_SQRT,        // g * (1 / (sqrt(v) * 2))
  VJP_DIV0,        // g * (1/v1)                                   in: g, v1
  VJP_DIV1,        // 0 - (((g * (1/v1)) * (1/v1)) * v0)           in: g, v1, v0
  VJP_SELECT0,     // v0 >= v1 ? g : 0                             in: g, v0, v1
  VJP_SELECT1,     // v0 >= v1 ? 0 : g
  COUNT_
};

struct EwArgs {
  gadcu_dtype dt = GADCU_F32;
  int nin = 0;
  const void* in[4] = {nullptr, nullptr, nullptr, nullptr};
  bool bcast[4] = {false, false, false, false};  // operand is a single element broadcast to n
  const void* acc = nullptr;                     // if set: out = acc + f(in...)   (may alias out)
  bool acc_bcast = false;
  void* out = nullptr;
  uint64_t n = 0;
  double c = 0.0, c2 = 0.0;
};
void launch_ew(gadcu_ctx* ctx, EwOp op, const EwArgs& a);
void launch_fill(gadcu_ctx* ctx, gadcu_dtype dt, void* out, uint64_t n, double value);

// out[idx] = in[idx mod phys] over 4-D column-major index space
void launch_tile(gadcu_ctx* ctx, gadcu_dtype dt, const void* in, const Dim4& phys, void* out, const Dim4& dims);

enum class RedOp : int { SUM, MAX };
// in viewed as [inner, reduce, outer] (column-major), out as [inner, outer]. If acc != null the result
// is acc + reduction (fused tile_as VJP + accumulation).
void launch_reduce(gadcu_ctx* ctx, gadcu_dtype dt, RedOp op, const void* in, void* out, uint64_t inner, uint64_t reduce,
                   uint64_t outer, const void* acc);
// The cut of an inner == 1 reduction (splits per segment, elements per chunk) and its second stage, for the lane-program
// kernel that reduces a deferred value in the same order without materialising it (batched.cu: lazy_reduce).
void reduce_contig_plan(gadcu_ctx* ctx, gadcu_dtype dt, uint64_t reduce, uint64_t outer, uint64_t* splits, uint64_t* chunk);
void launch_reduce_partials(gadcu_ctx* ctx, gadcu_dtype dt, RedOp op, const void* partials, void* out, uint64_t splits, uint64_t outer);
// out[0] = sum_i a[i]*b[i]; b_bcast: b is one element.  (af::dot(flat, flat), array.rs:119-126)
void launch_dot(gadcu_ctx* ctx, gadcu_dtype dt, const void* a, const void* b, bool a_bcast, bool b_bcast, void* out,
                uint64_t n);
void launch_transpose(gadcu_ctx* ctx, gadcu_dtype dt, const void* in, void* out, uint64_t rows, uint64_t cols);
void launch_random(gadcu_ctx* ctx, gadcu_dtype dt, void* out, uint64_t n, uint64_t seed, bool normal, double a, double b);

// Fused GEMM -> reduce-scatter (comm.cu): instead of writing C, the epilogue of output tile t = nt * (M/256) + mt
// stores it, as a dense 256 x 256 column-major block, into the staging buffer of the rank that owns the tile
// (t mod nranks) over NVLink peer memory: block index (t / nranks) * nranks + rank. The owner sums the nranks
// partials in rank order afterwards.
struct GemmPush {
  float* peer[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};  // staging base of every rank
  uint32_t nranks = 0, rank = 0;
  uint32_t splits = 1;  // partial blocks per (tile, rank): 2 when the GEMM splits its k range (gemm_push_splits)
};

// An elementwise tail evaluated in the GEMM's epilogue, on the accumulator while it is still in registers, instead of by
// a separate pass over C (SURVEY §2.1 "epilogue fusions"): the chains an MLP layer puts right behind its products. Each
// is the exact op sequence the tape would run as separate kernels — one rounding per op, no contraction — so the
// fused and the unfused sweep agree bit for bit (tests/test_gpu_epilogue.py).
//   BIAS_TANH     out = tanh(acc + bias[col])               add(matmul, tile_as(b)) -> tanh   (core.rs:179-182, analytic.rs:96-98)
//   TANH_GRAD     out = acc * ((0 - t*t) + 1), t = aux      the tanh VJP on an incoming dA    (analytic.rs:385-395)
//   BIAS_RESIDUAL out = aux - (acc + bias[col])             SquareLoss: sub(target, add(matmul, tile_as(b)))  (net_ext.rs:110-118)
// With hi / lo set the TF32 split of `out` (the operand copies the next tensor-core GEMM reads) is written in the same pass.
struct GemmEpilogue {
  enum Kind : int { NONE = 0, BIAS_TANH = 1, TANH_GRAD = 2, BIAS_RESIDUAL = 3 };
  int kind = NONE;
  const float* bias = nullptr;  // [N]
  const float* aux = nullptr;   // [M, N] column-major, leading dimension M
  float* hi = nullptr;          // [M, N], leading dimension M (needs M % 32 == 0), or null
  float* lo = nullptr;
};

// C[M,N] (+)= op(A) * op(B), column-major; batch over `batch` matrices with element strides (0 = broadcast).
struct GemmArgs {
  gadcu_dtype dt = GADCU_F32;
  const void* A = nullptr;
  const void* B = nullptr;
  void* C = nullptr;
  uint64_t M = 0, N = 0, K = 0;
  bool ta = false, tb = false;
  uint64_t lda = 0, ldb = 0;  // leading dims of the stored matrices
  uint64_t batch = 1, strideA = 0, strideB = 0, strideC = 0;
  bool accumulate = false;    // C = C + A*B (gradient accumulation fused into the epilogue)
  Buffer* a_buf = nullptr;    // owning buffers when an operand is a whole dense buffer (TF32 split cache)
  Buffer* b_buf = nullptr;
  const GemmPush* push = nullptr;  // tensor-core pair kernel only (M, N multiples of 256)
  const GemmEpilogue* epi = nullptr;  // tensor-core pair kernel only (gemm_epilogue_supported)
};
void launch_gemm(gadcu_ctx* ctx, const GemmArgs& g);
bool gemm_tc_supported(const GemmArgs& g);  // tcgen05 3xTF32 path applies
bool gemm_epilogue_supported(const GemmArgs& g);  // ... and its CTA-pair kernel would run it, so an epilogue can ride along
bool gemm_push_supported(const GemmArgs& g);  // ... and its CTA-pair kernel, which can push tiles to peers
size_t gemm_schedule_debug(int sm_count, uint64_t M, uint64_t N, uint64_t K, int push_splits, uint32_t* rows, size_t cap, uint32_t* info);  // gemm_tc.cu
// dst = src + d with the cached TF32 split of `src` refreshed for the sum in the same pass (gemm_tc.cu); false: not applicable
bool add_assign_refresh_split(gadcu_ctx* ctx, Buffer* src, const float* d, Buffer* dst, uint64_t n);
uint32_t gemm_push_splits(gadcu_ctx* ctx, const GemmArgs& g);  // 1 or 2: how the pair kernel would split k for this shape

}  // namespace k
}  // namespace gadcu
