// Experiment for round 2 (NOT part of the library; nothing builds or imports it): does the shared-memory ring that took the
// batched lane program from 0.196 to 0.184 ms (batched.cu, JitPlan::ring) also help the ARRAY-mode region kernel?
// The C3 reverse sweep (read x, y; write dz/dx, dz/dy; n = 64 Mi f32) as the sweep compiler generates it today
// (kernel A: 128-bit __ldcs loads, 4 vectors in flight per thread, grid-stride) against the same arithmetic fed by a
// cp.async.bulk producer warp through a ring of column tiles (kernel B). ncu of kernel A: 24 resident warps per SM,
// issue slots 64 % busy, long-scoreboard stalls dominate, DRAM 70 % of nominal — every warp alternates between
// waiting for its 8 loads and ~1000 instructions of exp/log/div, so the bytes in flight average ~40 % of what the
// registers could hold. Kernel B keeps the ring full regardless of what the computing warps do.
//
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -fmad=false -lineinfo experiments/c3_sweep_ring.cu -o experiments/c3_sweep_ring
//   gpurun -- ./experiments/c3_sweep_ring            # prints ms and algorithmic GB/s of A and of B for several ring shapes
#include <cstdio>
#include <cstdlib>
#include <vector>

#include <cuda_runtime.h>

#define CHECK(x)                                                                              \
  do {                                                                                        \
    cudaError_t e_ = (x);                                                                     \
    if (e_ != cudaSuccess) {                                                                  \
      std::printf("%s failed: %s\n", #x, cudaGetErrorString(e_));                            \
      std::exit(1);                                                                           \
    }                                                                                         \
  } while (0)

// the region's arithmetic for one lane, in the tape's order (add VJP, mul VJP, log VJP, exp VJP + accumulation)
__device__ __forceinline__ void body(float g, float x, float y, float& gx, float& gy) {
  const float a = expf(x), b = logf(y);
  const float ga = g * b, gb = a * g;
  gy = gb / y;
  gx = g + ga * a;
}

// ---- A: what gen_source emits today for this region (V = 4, U = 4, block 256) ---------------------------------
__global__ void __launch_bounds__(256) sweep_a(const float* __restrict__ xs, const float* __restrict__ ys, const float* __restrict__ seed,
                                               float* __restrict__ gxs, float* __restrict__ gys, unsigned long long n) {
  const unsigned long long tid = (unsigned long long)blockIdx.x * 256ull + threadIdx.x, stride = (unsigned long long)gridDim.x * 256ull;
  const float g = seed[0];
  const unsigned long long nvec = n / 4;
  for (unsigned long long base = tid; base < nvec; base += stride * 4) {
    float4 vx[4], vy[4];
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const unsigned long long vi = base + u * stride;
      if (vi < nvec) {
        vx[u] = __ldcs(reinterpret_cast<const float4*>(xs) + vi);
        vy[u] = __ldcs(reinterpret_cast<const float4*>(ys) + vi);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const unsigned long long vi = base + u * stride;
      if (vi < nvec) {
        float4 ox, oy;
        body(g, vx[u].x, vy[u].x, ox.x, oy.x);
        body(g, vx[u].y, vy[u].y, ox.y, oy.y);
        body(g, vx[u].z, vy[u].z, ox.z, oy.z);
        body(g, vx[u].w, vy[u].w, ox.w, oy.w);
        reinterpret_cast<float4*>(gxs)[vi] = ox;
        reinterpret_cast<float4*>(gys)[vi] = oy;
      }
    }
  }
  for (unsigned long long i = nvec * 4 + tid; i < n; i += stride) body(g, xs[i], ys[i], gxs[i], gys[i]);
}

// ---- B: a producer warp streams the two input columns through a ring of D tiles of BLOCK float4s ---------------
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
  unsigned done;
  do {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  } while (!done);
}

template <int BLOCK, int D, int VPT = 1>  // VPT float4s per thread and tile: tile = BLOCK * VPT * 16 bytes per input column
__global__ void __launch_bounds__(BLOCK + 32) sweep_b(const float* __restrict__ xs, const float* __restrict__ ys, const float* __restrict__ seed,
                                                      float* __restrict__ gxs, float* __restrict__ gys, unsigned long long n) {
  constexpr unsigned kTileBytes = BLOCK * VPT * 16u;
  constexpr unsigned long long kTileVecs = (unsigned long long)BLOCK * VPT;
  extern __shared__ __align__(128) unsigned char smem[];
  const unsigned sbase = (unsigned)__cvta_generic_to_shared(smem), bfull = sbase + D * kTileBytes, bempty = bfull + 8u * D;
  if (threadIdx.x == 0) {
    for (unsigned s = 0; s < D; s++) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bfull + 8 * s), "r"(1u));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bempty + 8 * s), "r"(unsigned(BLOCK / 32)));
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  const unsigned long long nvec = n / 4, tiles = (nvec + kTileVecs - 1) / kTileVecs;
  if (threadIdx.x >= BLOCK) {
    if (threadIdx.x == BLOCK) {
      unsigned q = 0;
      for (unsigned long long t = blockIdx.x; t < tiles; t += gridDim.x) {
        const unsigned long long first = t * kTileVecs, left = nvec - first;
        const unsigned bytes = unsigned(left < kTileVecs ? left : kTileVecs) * 16u;
#pragma unroll 1
        for (int j = 0; j < 2; j++, q++) {
          const unsigned s = q % D, ph = (q / D) & 1u;
          mbar_wait(bempty + 8 * s, ph ^ 1u);
          asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bfull + 8 * s), "r"(bytes) : "memory");
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(sbase + s * kTileBytes),
                       "l"((j == 0 ? xs : ys) + first * 4), "r"(bytes), "r"(bfull + 8 * s)
                       : "memory");
        }
      }
    }
    return;
  }
  const float g = seed[0];
  unsigned q = 0;
  for (unsigned long long t = blockIdx.x; t < tiles; t += gridDim.x) {
    float4 v[2][VPT];
#pragma unroll
    for (int j = 0; j < 2; j++, q++) {
      const unsigned s = q % D;
      mbar_wait(bfull + 8 * s, (q / D) & 1u);
#pragma unroll
      for (int k = 0; k < VPT; k++) v[j][k] = *reinterpret_cast<const float4*>(smem + s * kTileBytes + (k * BLOCK + threadIdx.x) * 16u);
      __syncwarp();
      if ((threadIdx.x & 31) == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bempty + 8 * s) : "memory");
    }
#pragma unroll
    for (int k = 0; k < VPT; k++) {
      const unsigned long long vi = t * kTileVecs + k * BLOCK + threadIdx.x;
      if (vi < nvec) {
        float4 ox, oy;
        body(g, v[0][k].x, v[1][k].x, ox.x, oy.x);
        body(g, v[0][k].y, v[1][k].y, ox.y, oy.y);
        body(g, v[0][k].z, v[1][k].z, ox.z, oy.z);
        body(g, v[0][k].w, v[1][k].w, ox.w, oy.w);
        reinterpret_cast<float4*>(gxs)[vi] = ox;
        reinterpret_cast<float4*>(gys)[vi] = oy;
      }
    }
  }
  // (n is a multiple of 4 in this experiment: no scalar tail)
}

__global__ void fill(float* p, unsigned long long n, float lo, float hi, unsigned seed) {
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (unsigned long long)gridDim.x * blockDim.x) {
    unsigned h = unsigned(i) * 2654435761u ^ seed;
    h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
    p[i] = lo + (hi - lo) * float(h >> 8) * (1.0f / 16777216.0f);
  }
}

template <class F> static float time_ms(F launch, int reps) {
  cudaEvent_t e0, e1;
  CHECK(cudaEventCreate(&e0));
  CHECK(cudaEventCreate(&e1));
  for (int i = 0; i < 3; i++) launch();
  CHECK(cudaDeviceSynchronize());
  CHECK(cudaEventRecord(e0));
  for (int i = 0; i < reps; i++) launch();
  CHECK(cudaEventRecord(e1));
  CHECK(cudaEventSynchronize(e1));
  float ms = 0;
  CHECK(cudaEventElapsedTime(&ms, e0, e1));
  CHECK(cudaGetLastError());
  return ms / reps;
}

template <int BLOCK, int D, int VPT = 1> static void run_b(int blocks_per_sm, int sms, const float* x, const float* y, const float* seed, float* gx, float* gy,
                                               unsigned long long n, const float* ref_gx, const float* ref_gy) {
  const size_t smem = size_t(D) * BLOCK * VPT * 16 + 16 * D;
  CHECK(cudaFuncSetAttribute(sweep_b<BLOCK, D, VPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
  const int grid = sms * blocks_per_sm;
  const float ms = time_ms([&] { sweep_b<BLOCK, D, VPT><<<grid, BLOCK + 32, smem>>>(x, y, seed, gx, gy, n); }, 20);
  std::vector<float> a(4096), b(4096), c(4096), d(4096);
  CHECK(cudaMemcpy(a.data(), gx + n / 2, 4096 * 4, cudaMemcpyDeviceToHost));
  CHECK(cudaMemcpy(b.data(), ref_gx + n / 2, 4096 * 4, cudaMemcpyDeviceToHost));
  CHECK(cudaMemcpy(c.data(), gy + n - 4096, 4096 * 4, cudaMemcpyDeviceToHost));
  CHECK(cudaMemcpy(d.data(), ref_gy + n - 4096, 4096 * 4, cudaMemcpyDeviceToHost));
  std::printf("B block %3d x %d vec ring %2d x %d blocks/SM (%3zu KB smem/SM): %.4f ms  %.0f GB/s  %s\n", BLOCK, VPT, D, blocks_per_sm, smem * blocks_per_sm / 1024, ms,
              16.0 * n / (ms * 1e-3) / 1e9, a == b && c == d ? "same bits as A" : "DIFFERS from A");
}

int main() {
  const unsigned long long n = 64ull << 20;
  cudaDeviceProp prop;
  CHECK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  float *x, *y, *gx, *gy, *rx, *ry, *seed;
  for (float** p : {&x, &y, &gx, &gy, &rx, &ry}) CHECK(cudaMalloc(p, n * 4));
  CHECK(cudaMalloc(&seed, 4));
  const float one = 1.0f;
  CHECK(cudaMemcpy(seed, &one, 4, cudaMemcpyHostToDevice));
  fill<<<sms * 8, 256>>>(x, n, -1.0f, 1.0f, 3);
  fill<<<sms * 8, 256>>>(y, n, 0.5f, 2.0f, 4);
  const int grid_a = sms * 8;
  const float ms_a = time_ms([&] { sweep_a<<<grid_a, 256>>>(x, y, seed, rx, ry, n); }, 20);
  std::printf("A (today's generated kernel shape): %.4f ms  %.0f GB/s algorithmic\n", ms_a, 16.0 * n / (ms_a * 1e-3) / 1e9);
  run_b<256, 8>(3, sms, x, y, seed, gx, gy, n, rx, ry);
  run_b<256, 8>(5, sms, x, y, seed, gx, gy, n, rx, ry);
  run_b<256, 16>(3, sms, x, y, seed, gx, gy, n, rx, ry);
  run_b<128, 8>(6, sms, x, y, seed, gx, gy, n, rx, ry);
  run_b<128, 16>(6, sms, x, y, seed, gx, gy, n, rx, ry);
  run_b<128, 8>(10, sms, x, y, seed, gx, gy, n, rx, ry);
  run_b<512, 4>(2, sms, x, y, seed, gx, gy, n, rx, ry);   // round 1: the winner, 0.1774 ms
  // round 2: larger copies still? (never run)
  run_b<512, 2>(2, sms, x, y, seed, gx, gy, n, rx, ry);
  run_b<512, 4>(3, sms, x, y, seed, gx, gy, n, rx, ry);
  run_b<512, 4, 2>(2, sms, x, y, seed, gx, gy, n, rx, ry);  // 16 KB per copy
  run_b<512, 2, 2>(3, sms, x, y, seed, gx, gy, n, rx, ry);
  run_b<256, 4, 4>(3, sms, x, y, seed, gx, gy, n, rx, ry);  // 16 KB per copy, fewer threads
  run_b<960, 4>(1, sms, x, y, seed, gx, gy, n, rx, ry);     // 15 KB per copy, one block per SM
  run_b<960, 2, 2>(1, sms, x, y, seed, gx, gy, n, rx, ry);  // 30 KB per copy
  return 0;
}
