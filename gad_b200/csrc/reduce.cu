// reduce.cu — sum/max reductions (`sum_as`, `max_as`, `dot`), broadcast (`tile_as`), 2-D transpose
// and the seeded synthetic-data generator.
//
// Reduction order (documented for the parity tolerance): the input is viewed as
// [inner, reduce, outer]. For inner == 1 every contiguous segment is cut into `splits` equal chunks;
// inside a chunk each of the 256 threads accumulates a strided subsequence (vector lanes kept
// separate), lanes are combined, then a butterfly warp shuffle, then the 8 warp results in warp
// order; the chunk partials are finally added in chunk order. For inner > 1 each output element is
// accumulated sequentially along the reduced axis within a split, splits added in order. Both are
// deterministic for a given shape and SM count. All accumulation is in the array's own type T.
#include "kernels.h"

namespace gadcu {
namespace k {
namespace {

constexpr int kThreads = 256;

template <class T> struct VecT;
template <> struct VecT<float> { using type = float4; static constexpr int N = 4; };
template <> struct VecT<double> { using type = double2; static constexpr int N = 2; };

template <class T, RedOp OP>
__device__ __forceinline__ T combine(T a, T b) {
  if constexpr (OP == RedOp::SUM) return a + b;
  else return b > a ? b : a;
}
template <class T, RedOp OP>
__device__ __forceinline__ T identity() {
  if constexpr (OP == RedOp::SUM) return T(0);
  else return -INFINITY;
}

template <class T, RedOp OP>
__device__ __forceinline__ T block_reduce(T v) {
  __shared__ T warp_part[kThreads / 32];
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v = combine<T, OP>(v, __shfl_xor_sync(0xffffffffu, v, off));
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) warp_part[warp] = v;
  __syncthreads();
  T r = identity<T, OP>();
  if (threadIdx.x == 0) {
    r = warp_part[0];
#pragma unroll
    for (int w = 1; w < kThreads / 32; w++) r = combine<T, OP>(r, warp_part[w]);
  }
  __syncthreads();
  return r;  // valid in thread 0
}

// inner == 1: segment `seg` = in[seg*reduce, (seg+1)*reduce); blockIdx.x = split, blockIdx.y = seg.
// DOT: elements are a[i]*b[i]. Writes out[seg*splits + split] (+ acc[seg] when ACC and splits == 1).
template <class T, RedOp OP, bool DOT, bool VEC>
__global__ void __launch_bounds__(kThreads) reduce_contig_kernel(const T* __restrict__ a, const T* __restrict__ b, T* out,
                                                                 const T* acc, uint64_t reduce, uint64_t chunk) {
  const uint64_t seg = blockIdx.y, split = blockIdx.x, splits = gridDim.x;
  const uint64_t lo = split * chunk;
  uint64_t hi = lo + chunk;
  if (hi > reduce) hi = reduce;
  const T* pa = a + seg * reduce;
  const T* pb = DOT ? b + seg * reduce : nullptr;
  T v = identity<T, OP>();
  if (lo < hi) {
    if constexpr (VEC) {
      constexpr int V = VecT<T>::N;
      using VT = typename VecT<T>::type;
      // chunk and reduce are multiples of V and the segment base is 16-byte aligned (checked on host)
      const uint64_t vlo = lo / V, vhi = hi / V;
      T lane_acc[V];
#pragma unroll
      for (int j = 0; j < V; j++) lane_acc[j] = identity<T, OP>();
      for (uint64_t vi = vlo + threadIdx.x; vi < vhi; vi += kThreads) {
        VT x = __ldcs(reinterpret_cast<const VT*>(pa) + vi);
        T xs[V];
        if constexpr (V == 4) { xs[0] = x.x; xs[1] = x.y; xs[2] = x.z; xs[3] = x.w; } else { xs[0] = x.x; xs[1] = x.y; }
        if constexpr (DOT) {
          VT y = __ldcs(reinterpret_cast<const VT*>(pb) + vi);
          if constexpr (V == 4) { xs[0] *= y.x; xs[1] *= y.y; xs[2] *= y.z; xs[3] *= y.w; } else { xs[0] *= y.x; xs[1] *= y.y; }
        }
#pragma unroll
        for (int j = 0; j < V; j++) lane_acc[j] = combine<T, OP>(lane_acc[j], xs[j]);
      }
      v = lane_acc[0];
#pragma unroll
      for (int j = 1; j < V; j++) v = combine<T, OP>(v, lane_acc[j]);
    } else {
      for (uint64_t i = lo + threadIdx.x; i < hi; i += kThreads) {
        T x = pa[i];
        if constexpr (DOT) x *= pb[i];
        v = combine<T, OP>(v, x);
      }
    }
  }
  T r = block_reduce<T, OP>(v);
  if (threadIdx.x == 0) {
    if (splits == 1 && acc) r = acc[seg] + r;
    out[seg * splits + split] = r;
  }
}

// inner > 1: thread per (i, outer) with i fastest; blockIdx.y = split over the reduced axis,
// blockIdx.z = outer. out[(o*splits + split)*inner + i].
template <class T, RedOp OP>
__global__ void __launch_bounds__(kThreads) reduce_strided_kernel(const T* __restrict__ in, T* out, const T* acc,
                                                                  uint64_t inner, uint64_t reduce, uint64_t chunk) {
  const uint64_t i = uint64_t(blockIdx.x) * kThreads + threadIdx.x;
  if (i >= inner) return;
  const uint64_t split = blockIdx.y, splits = gridDim.y, o = blockIdx.z;
  const uint64_t lo = split * chunk;
  uint64_t hi = lo + chunk;
  if (hi > reduce) hi = reduce;
  const T* p = in + i + inner * reduce * o;
  T v = identity<T, OP>();
  uint64_t r = lo;
  for (; r + 4 <= hi; r += 4) {  // 4 loads in flight; combined in index order
    T x0 = p[inner * r], x1 = p[inner * (r + 1)], x2 = p[inner * (r + 2)], x3 = p[inner * (r + 3)];
    v = combine<T, OP>(v, x0);
    v = combine<T, OP>(v, x1);
    v = combine<T, OP>(v, x2);
    v = combine<T, OP>(v, x3);
  }
  for (; r < hi; r++) v = combine<T, OP>(v, p[inner * r]);
  if (splits == 1 && acc) v = acc[i + inner * o] + v;
  out[(o * splits + split) * inner + i] = v;
}

template <class T>
__global__ void __launch_bounds__(kThreads) tile_generic_kernel(const T* __restrict__ in, T* __restrict__ out, Dim4 phys,
                                                                Dim4 dims, uint64_t n) {
  const uint64_t stride = uint64_t(gridDim.x) * kThreads;
  for (uint64_t idx = uint64_t(blockIdx.x) * kThreads + threadIdx.x; idx < n; idx += stride) {
    uint64_t r = idx;
    const uint64_t i0 = r % dims.d[0]; r /= dims.d[0];
    const uint64_t i1 = r % dims.d[1]; r /= dims.d[1];
    const uint64_t i2 = r % dims.d[2]; r /= dims.d[2];
    const uint64_t i3 = r;
    const uint64_t src = (i0 % phys.d[0]) + phys.d[0] * ((i1 % phys.d[1]) + phys.d[1] * ((i2 % phys.d[2]) + phys.d[2] * (i3 % phys.d[3])));
    out[idx] = in[src];
  }
}

// dims[0] is either broadcast from 1 (ROW: every element of an output column segment equals one
// input element) or copied whole (phys[0] == dims[0]). blockIdx.y walks the other three dims.
template <class T, bool ROW>
__global__ void __launch_bounds__(kThreads) tile_dim0_kernel(const T* __restrict__ in, T* __restrict__ out, Dim4 phys,
                                                             Dim4 dims) {
  uint64_t rest = blockIdx.y;
  const uint64_t i1 = rest % dims.d[1]; rest /= dims.d[1];
  const uint64_t i2 = rest % dims.d[2]; rest /= dims.d[2];
  const uint64_t i3 = rest;
  const uint64_t src_rest = (i1 % phys.d[1]) + phys.d[1] * ((i2 % phys.d[2]) + phys.d[2] * (i3 % phys.d[3]));
  const uint64_t n0 = dims.d[0];
  T* o = out + uint64_t(blockIdx.y) * n0;
  const uint64_t stride = uint64_t(gridDim.x) * kThreads;
  if constexpr (ROW) {
    const T v = in[src_rest];
    for (uint64_t i = uint64_t(blockIdx.x) * kThreads + threadIdx.x; i < n0; i += stride) o[i] = v;
  } else {
    const T* s = in + src_rest * n0;
    for (uint64_t i = uint64_t(blockIdx.x) * kThreads + threadIdx.x; i < n0; i += stride) o[i] = s[i];
  }
}

template <class T>
__global__ void transpose_kernel(const T* __restrict__ in, T* __restrict__ out, uint64_t rows, uint64_t cols) {
  // in: [rows, cols] column-major (rows fastest); out: [cols, rows]
  __shared__ T tile[32][33];
  uint64_t r = uint64_t(blockIdx.x) * 32 + threadIdx.x;
  uint64_t c0 = uint64_t(blockIdx.y) * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    uint64_t c = c0 + j;
    if (r < rows && c < cols) tile[j][threadIdx.x] = in[r + rows * c];
  }
  __syncthreads();
  uint64_t oc = c0 + threadIdx.x;  // fastest index of out = original column
  uint64_t r0 = uint64_t(blockIdx.x) * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    uint64_t orow = r0 + j;
    if (oc < cols && orow < rows) out[oc + cols * orow] = tile[threadIdx.x][j];
  }
}

__device__ __forceinline__ uint64_t mix64(uint64_t z) {  // splitmix64 finaliser
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
template <class T>
__global__ void __launch_bounds__(kThreads) random_kernel(T* out, uint64_t n, uint64_t seed, bool normal, double a, double b) {
  const uint64_t stride = uint64_t(gridDim.x) * kThreads;
  for (uint64_t i = uint64_t(blockIdx.x) * kThreads + threadIdx.x; i < n; i += stride) {
    uint64_t h1 = mix64(seed * 0x9E3779B97F4A7C15ull + 2 * i + 1);
    double u1 = double(h1 >> 11) * (1.0 / 9007199254740992.0);
    double v;
    if (normal) {
      uint64_t h2 = mix64(seed * 0x9E3779B97F4A7C15ull + 2 * i + 2);
      double u2 = double(h2 >> 11) * (1.0 / 9007199254740992.0);
      double rad = sqrt(-2.0 * log(1.0 - u1));
      v = a + b * rad * cospi(2.0 * u2);
    } else {
      v = a + (b - a) * u1;
    }
    out[i] = T(v);
  }
}

// How a contiguous segment of `reduce` elements is cut (the documented order above depends on nothing else): also used
// by the lane-program kernel that reduces a deferred elementwise value without writing it (batched.cu), which has to
// produce the same bits as this file's kernel run over the materialised value.
void contig_plan(int sm_count, int V, uint64_t reduce, uint64_t outer, uint64_t& splits, uint64_t& chunk) {
  const uint64_t target_blocks = uint64_t(sm_count) * 8;
  splits = 1;
  if (outer < target_blocks) {
    splits = target_blocks / outer;
    const uint64_t max_splits = (reduce + 8191) / 8192;  // at least 8192 elements per chunk
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    if (splits > 65535) splits = 65535;
  }
  chunk = (reduce + splits - 1) / splits;
  chunk = (chunk + V - 1) / V * V;
  splits = reduce == 0 ? 1 : (reduce + chunk - 1) / chunk;
}

template <class T, RedOp OP>
void reduce_typed(gadcu_ctx* ctx, const T* in, const T* in2, bool dot, T* out, uint64_t inner, uint64_t reduce, uint64_t outer,
                  const T* acc) {
  if (inner * outer == 0) return;
  const uint64_t target_blocks = uint64_t(ctx->sm_count) * 8;
  if (inner == 1) {
    constexpr int V = VecT<T>::N;
    const bool vec = (reduce % V == 0) && (reinterpret_cast<uintptr_t>(in) % 16 == 0) &&
                     (!dot || reinterpret_cast<uintptr_t>(in2) % 16 == 0);
    uint64_t splits, chunk;
    contig_plan(ctx->sm_count, V, reduce, outer, splits, chunk);
    if (outer > 65535) {
      // fold: a segment per blockIdx.y is limited to 65535; process in slabs
      for (uint64_t o0 = 0; o0 < outer; o0 += 65535) {
        uint64_t cnt = outer - o0 < 65535 ? outer - o0 : 65535;
        reduce_typed<T, OP>(ctx, in + o0 * reduce, dot ? in2 + o0 * reduce : nullptr, dot, out + o0, 1, reduce, cnt,
                            acc ? acc + o0 : nullptr);
      }
      return;
    }
    ArrayRef tmp;
    T* stage_out = out;
    if (splits > 1) {
      tmp = new_array(ctx, sizeof(T) == 4 ? GADCU_F32 : GADCU_F64, Dim4(splits * outer));
      stage_out = static_cast<T*>(tmp->ptr());
    }
    dim3 grid((unsigned)splits, (unsigned)outer);
    const T* acc1 = splits == 1 ? acc : nullptr;
#define LAUNCH_CONTIG(DOT_, VEC_) \
  reduce_contig_kernel<T, OP, DOT_, VEC_><<<grid, kThreads, 0, ctx->stream>>>(in, in2, stage_out, acc1, reduce, chunk)
    if (dot) { if (vec) LAUNCH_CONTIG(true, true); else LAUNCH_CONTIG(true, false); }
    else { if (vec) LAUNCH_CONTIG(false, true); else LAUNCH_CONTIG(false, false); }
#undef LAUNCH_CONTIG
    ctx->count_launch();
    if (splits > 1) {
      // second stage: each segment's `splits` partials, in chunk order
      dim3 grid2(1, (unsigned)outer);
      reduce_contig_kernel<T, OP, false, false><<<grid2, kThreads, 0, ctx->stream>>>(stage_out, nullptr, out, acc, splits, splits);
      ctx->count_launch();
    }
    return;
  }
  // strided
  const uint64_t bx = (inner + kThreads - 1) / kThreads;
  uint64_t splits = 1;
  if (bx * outer < target_blocks) {
    splits = target_blocks / (bx * outer);
    const uint64_t max_splits = (reduce + 63) / 64;
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    if (splits > 65535) splits = 65535;
  }
  const uint64_t chunk = (reduce + splits - 1) / splits;
  splits = reduce == 0 ? 1 : (reduce + chunk - 1) / chunk;
  if (outer > 65535) fail(GADCU_ERR_UNSUPPORTED, "strided reduction with more than 65535 outer slices");
  ArrayRef tmp;
  T* stage_out = out;
  if (splits > 1) {
    tmp = new_array(ctx, sizeof(T) == 4 ? GADCU_F32 : GADCU_F64, Dim4(inner * splits * outer));
    stage_out = static_cast<T*>(tmp->ptr());
  }
  dim3 grid((unsigned)bx, (unsigned)splits, (unsigned)outer);
  reduce_strided_kernel<T, OP><<<grid, kThreads, 0, ctx->stream>>>(in, stage_out, splits == 1 ? acc : nullptr, inner, reduce, chunk);
  ctx->count_launch();
  if (splits > 1) {
    dim3 grid2((unsigned)bx, 1, (unsigned)outer);
    reduce_strided_kernel<T, OP><<<grid2, kThreads, 0, ctx->stream>>>(stage_out, out, acc, inner, splits, splits);
    ctx->count_launch();
  }
}

}  // namespace

void launch_reduce(gadcu_ctx* ctx, gadcu_dtype dt, RedOp op, const void* in, void* out, uint64_t inner, uint64_t reduce,
                   uint64_t outer, const void* acc) {
  ProfileScope prof_scope(ctx, PROF_REDUCE);
  if (dt == GADCU_F32) {
    if (op == RedOp::SUM) reduce_typed<float, RedOp::SUM>(ctx, (const float*)in, nullptr, false, (float*)out, inner, reduce, outer, (const float*)acc);
    else reduce_typed<float, RedOp::MAX>(ctx, (const float*)in, nullptr, false, (float*)out, inner, reduce, outer, (const float*)acc);
  } else {
    if (op == RedOp::SUM) reduce_typed<double, RedOp::SUM>(ctx, (const double*)in, nullptr, false, (double*)out, inner, reduce, outer, (const double*)acc);
    else reduce_typed<double, RedOp::MAX>(ctx, (const double*)in, nullptr, false, (double*)out, inner, reduce, outer, (const double*)acc);
  }
}

void reduce_contig_plan(gadcu_ctx* ctx, gadcu_dtype dt, uint64_t reduce, uint64_t outer, uint64_t* splits, uint64_t* chunk) {
  contig_plan(ctx->sm_count, dt == GADCU_F32 ? 4 : 2, reduce, outer, *splits, *chunk);
}

// second stage of a split contiguous reduction: out[seg] = the segment's `splits` partials combined in chunk order
void launch_reduce_partials(gadcu_ctx* ctx, gadcu_dtype dt, RedOp op, const void* partials, void* out, uint64_t splits, uint64_t outer) {
  ProfileScope prof_scope(ctx, PROF_REDUCE);
  dim3 grid2(1, (unsigned)outer);
#define LAUNCH_STAGE2(T, OP) \
  reduce_contig_kernel<T, OP, false, false><<<grid2, kThreads, 0, ctx->stream>>>((const T*)partials, nullptr, (T*)out, nullptr, splits, splits)
  if (dt == GADCU_F32) { if (op == RedOp::SUM) LAUNCH_STAGE2(float, RedOp::SUM); else LAUNCH_STAGE2(float, RedOp::MAX); }
  else { if (op == RedOp::SUM) LAUNCH_STAGE2(double, RedOp::SUM); else LAUNCH_STAGE2(double, RedOp::MAX); }
#undef LAUNCH_STAGE2
  ctx->count_launch();
}

void launch_dot(gadcu_ctx* ctx, gadcu_dtype dt, const void* a, const void* b, bool a_bcast, bool b_bcast, void* out, uint64_t n) {
  ProfileScope prof_scope(ctx, PROF_REDUCE);
  if (a_bcast || b_bcast) fail(GADCU_ERR_INVALID, "dot operands must be materialised");
  if (dt == GADCU_F32) reduce_typed<float, RedOp::SUM>(ctx, (const float*)a, (const float*)b, true, (float*)out, 1, n, 1, nullptr);
  else reduce_typed<double, RedOp::SUM>(ctx, (const double*)a, (const double*)b, true, (double*)out, 1, n, 1, nullptr);
}

void launch_tile(gadcu_ctx* ctx, gadcu_dtype dt, const void* in, const Dim4& phys, void* out, const Dim4& dims) {
  ProfileScope prof_scope(ctx, PROF_OTHER);
  const uint64_t n = dims.elements();
  if (n == 0) return;
  const uint64_t rest = dims[1] * dims[2] * dims[3];
  const bool row = phys[0] == 1, whole = phys[0] == dims[0];
  if ((row || whole) && rest <= 65535 && dims[0] >= 64) {
    uint64_t bx = (dims[0] + kThreads * 4 - 1) / (kThreads * 4);
    if (bx > 1024) bx = 1024;
    dim3 grid((unsigned)bx, (unsigned)rest);
#define LAUNCH_TILE0(T) \
  if (row) tile_dim0_kernel<T, true><<<grid, kThreads, 0, ctx->stream>>>((const T*)in, (T*)out, phys, dims); \
  else tile_dim0_kernel<T, false><<<grid, kThreads, 0, ctx->stream>>>((const T*)in, (T*)out, phys, dims)
    if (dt == GADCU_F32) { LAUNCH_TILE0(float); } else { LAUNCH_TILE0(double); }
#undef LAUNCH_TILE0
    ctx->count_launch();
    return;
  }
  uint64_t blocks = (n + kThreads - 1) / kThreads;
  const uint64_t cap = uint64_t(ctx->sm_count) * 8;
  if (blocks > cap) blocks = cap;
  if (dt == GADCU_F32) tile_generic_kernel<float><<<unsigned(blocks), kThreads, 0, ctx->stream>>>((const float*)in, (float*)out, phys, dims, n);
  else tile_generic_kernel<double><<<unsigned(blocks), kThreads, 0, ctx->stream>>>((const double*)in, (double*)out, phys, dims, n);
  ctx->count_launch();
}

void launch_transpose(gadcu_ctx* ctx, gadcu_dtype dt, const void* in, void* out, uint64_t rows, uint64_t cols) {
  ProfileScope prof_scope(ctx, PROF_OTHER);
  if (rows * cols == 0) return;
  dim3 block(32, 8);
  dim3 grid(unsigned((rows + 31) / 32), unsigned((cols + 31) / 32));
  if (grid.y > 65535) fail(GADCU_ERR_UNSUPPORTED, "transpose: too many columns");
  if (dt == GADCU_F32) transpose_kernel<float><<<grid, block, 0, ctx->stream>>>((const float*)in, (float*)out, rows, cols);
  else transpose_kernel<double><<<grid, block, 0, ctx->stream>>>((const double*)in, (double*)out, rows, cols);
  ctx->count_launch();
}

void launch_random(gadcu_ctx* ctx, gadcu_dtype dt, void* out, uint64_t n, uint64_t seed, bool normal, double a, double b) {
  if (n == 0) return;
  uint64_t blocks = (n + kThreads - 1) / kThreads;
  const uint64_t cap = uint64_t(ctx->sm_count) * 8;
  if (blocks > cap) blocks = cap;
  if (dt == GADCU_F32) random_kernel<float><<<unsigned(blocks), kThreads, 0, ctx->stream>>>((float*)out, n, seed, normal, a, b);
  else random_kernel<double><<<unsigned(blocks), kThreads, 0, ctx->stream>>>((double*)out, n, seed, normal, a, b);
  ctx->count_launch();
}

}  // namespace k
}  // namespace gadcu
