// gemm_tc.cu — f32 GEMM on the 5th-generation tensor cores: tcgen05.mma (kind::tf32) fed by TMA,
// accumulators in TMEM, 3xTF32 operand splitting for f32 accuracy. This is the only tensor-core path
// of the backend: gad's `matmul` forward and its two VJP products dA = G.B^T, dB = A^T.G
// (gad/src/matrix.rs:153-179), which are >95 % of the MLP step.
//
// 3xTF32: every f32 operand x is split (split_tf32_kernel, one HBM pass) into hi = rna_tf32(x) and
// lo = rna_tf32(x - hi); the product is accumulated in f32 (TMEM) as
//      A.B  ~=  Ahi.Bhi + Ahi.Blo + Alo.Bhi          (the dropped Alo.Blo term is <= 2^-22 |a||b|)
// i.e. three tensor-core MMAs per logical product.
//
// Kernel shape (v1, one CTA per 128 x BN output tile):
//   warp 4      TMA producer : per k-block of 32, four bulk tensor copies (Ahi, Alo, Bhi, Blo) into a
//                              kStages-deep shared-memory ring, completion on an mbarrier (expect_tx)
//   warp 5      MMA issuer   : one elected thread issues 4 k-steps x 3 tcgen05.mma per k-block into the
//                              TMEM accumulator, tcgen05.commit releases the ring slot
//   warps 0-3   epilogue     : tcgen05.ld (32 lanes x 32 columns per warp), optional `C + acc`
//                              (gradient accumulation fused into the epilogue), coalesced column-major stores
// Operands are column-major with optional transposition, so each is either K-major or MN-major in
// UMMA terms; both are fed directly (no transpose pass): the TMA tensor map lays the tile out in the
// canonical 128-byte-swizzled layout of its major-ness and the smem/instruction descriptors say which.
#include <cuda.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <string>

#include "kernels.h"

namespace gadcu {
namespace k {
namespace {

constexpr int BM = 128;      // UMMA M (cta_group::1)
constexpr int BK = 32;       // k-block: 32 tf32 = 128 bytes = one swizzle row
constexpr int UMMA_K = 8;    // tf32: 32 bytes per instruction
constexpr int kEpiWarps = 4;
constexpr int kThreads = 192;

template <int BN> struct Cfg {
  static constexpr int kStages = BN == 256 ? 2 : 3;
  static constexpr uint32_t kTileA = BM * BK * 4;       // 16 KiB
  static constexpr uint32_t kTileB = BN * BK * 4;
  static constexpr uint32_t kStageBytes = 2 * kTileA + 2 * kTileB;
  static constexpr uint32_t kSmem = kStages * kStageBytes + 1024 /*align*/ + 256 /*barriers*/;
};

// ---- PTX wrappers -------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra WAIT_DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t"
      "}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
               "l"(map), "r"(bar), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
               "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t smem_slot, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_slot), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[smem] . B[smem]   (kind::tf32, f32 accumulate)
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start>>4 [0,14), LBO>>4 [16,30),
// SBO>>4 [32,46), version=1 [46,48), layout SWIZZLE_128B=2 [61,64).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
  uint64_t d = 0;
  d |= uint64_t((addr >> 4) & 0x3FFF);
  d |= uint64_t((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= uint64_t((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= uint64_t(1) << 46;
  d |= uint64_t(layout) << 61;
  return d;
}

// Layout of one operand tile in shared memory (rows of 128 bytes, swizzled within 128-byte spans):
//   K-major  : [rows = MN][32 k], 16-byte chunks swizzled over 8-row groups (SWIZZLE_128B, layout 2),
//              8-row groups 1024 B apart -> SBO = 1024, k-step (8 k) +32 B inside the row
//   MN-major : [MN/32 slabs][32 k][32 mn]; for 32-bit operands the only MN-major layout tcgen05 takes is
//              SWIZZLE_128B_BASE32B (layout 1): 32-byte chunks swizzled over 4-row groups, i.e. an atom
//              of 32 mn x 4 k = 512 B -> SBO = 512 (next 4 k), LBO = slab = BK*128 B (next 32 mn),
//              k-step (8 k-rows) +1024 B. The TMA map uses CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B to match.
struct OperandLayout {
  uint32_t lbo, sbo, kstep, layout;
};

template <int BN>
__global__ void __launch_bounds__(kThreads, 1)
gemm_tf32x3_kernel(const __grid_constant__ CUtensorMap map_ah, const __grid_constant__ CUtensorMap map_al,
                   const __grid_constant__ CUtensorMap map_bh, const __grid_constant__ CUtensorMap map_bl, float* __restrict__ C,
                   uint32_t M, uint32_t N, uint32_t K, uint32_t a_mn_major, uint32_t b_mn_major, uint32_t accumulate,
                   OperandLayout la, OperandLayout lb) {
  using cfg = Cfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B tiles need 1024-byte alignment
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t bars = smem_base + cfg::kStages * cfg::kStageBytes;  // full[S], empty[S], tmem_full, tmem slot
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (cfg::kStages + s); };
  const uint32_t tmem_full_bar = bars + 8u * (2 * cfg::kStages);
  const uint32_t tmem_slot = bars + 8u * (2 * cfg::kStages + 1);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem + cfg::kStages * cfg::kStageBytes + 8 * (2 * cfg::kStages + 1));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const uint32_t num_kb = K / BK;

  if (threadIdx.x == 0) {
    for (int s = 0; s < cfg::kStages; s++) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&map_ah);
    tma_prefetch_desc(&map_al);
    tma_prefetch_desc(&map_bh);
    tma_prefetch_desc(&map_bl);
  }
  if (warp == 5) tmem_alloc(tmem_slot, BN);  // BN f32 accumulator columns (power of two >= 32)
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 4) {
    // ===== TMA producer =====
    if (lane == 0) {
      for (uint32_t kb = 0; kb < num_kb; kb++) {
        const int s = kb % cfg::kStages;
        const uint32_t phase = (kb / cfg::kStages) & 1;
        mbar_wait(empty_bar(s), phase ^ 1);
        const uint32_t stage = smem_base + s * cfg::kStageBytes;
        const uint32_t sa_h = stage, sa_l = stage + cfg::kTileA, sb_h = stage + 2 * cfg::kTileA, sb_l = sb_h + cfg::kTileB;
        mbar_expect_tx(full_bar(s), cfg::kStageBytes);
        const int k0 = int(kb * BK);
        if (a_mn_major) {
          tma_load_3d(sa_h, &map_ah, full_bar(s), 0, k0, int(m0 / 32));
          tma_load_3d(sa_l, &map_al, full_bar(s), 0, k0, int(m0 / 32));
        } else {
          tma_load_2d(sa_h, &map_ah, full_bar(s), k0, int(m0));
          tma_load_2d(sa_l, &map_al, full_bar(s), k0, int(m0));
        }
        if (b_mn_major) {
          tma_load_3d(sb_h, &map_bh, full_bar(s), 0, k0, int(n0 / 32));
          tma_load_3d(sb_l, &map_bl, full_bar(s), 0, k0, int(n0 / 32));
        } else {
          tma_load_2d(sb_h, &map_bh, full_bar(s), k0, int(n0));
          tma_load_2d(sb_l, &map_bl, full_bar(s), k0, int(n0));
        }
      }
    }
  } else if (warp == 5) {
    // ===== MMA issuer (single thread) =====
    if (lane == 0) {
      // instruction descriptor (cute::UMMA::InstrDescriptor): c=F32 [4,6), a=b=TF32 [7,10)/[10,13),
      // a_major [15], b_major [16], N>>3 [17,23), M>>4 [24,29)
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((a_mn_major ? 1u : 0u) << 15) | ((b_mn_major ? 1u : 0u) << 16) |
                             (uint32_t(BN >> 3) << 17) | (uint32_t(BM >> 4) << 24);
      for (uint32_t kb = 0; kb < num_kb; kb++) {
        const int s = kb % cfg::kStages;
        const uint32_t phase = (kb / cfg::kStages) & 1;
        mbar_wait(full_bar(s), phase);
        tc_fence_after();
        const uint32_t stage = smem_base + s * cfg::kStageBytes;
        const uint32_t sa_h = stage, sa_l = stage + cfg::kTileA, sb_h = stage + 2 * cfg::kTileA, sb_l = sb_h + cfg::kTileB;
#pragma unroll
        for (int k = 0; k < BK / UMMA_K; k++) {
          const uint64_t dah = make_smem_desc(sa_h + k * la.kstep, la.lbo, la.sbo, la.layout);
          const uint64_t dal = make_smem_desc(sa_l + k * la.kstep, la.lbo, la.sbo, la.layout);
          const uint64_t dbh = make_smem_desc(sb_h + k * lb.kstep, lb.lbo, lb.sbo, lb.layout);
          const uint64_t dbl = make_smem_desc(sb_l + k * lb.kstep, lb.lbo, lb.sbo, lb.layout);
          // small terms first, then the leading product
          umma_tf32(tmem_base, dal, dbh, idesc, (kb | uint32_t(k)) != 0);
          umma_tf32(tmem_base, dah, dbl, idesc, 1);
          umma_tf32(tmem_base, dah, dbh, idesc, 1);
        }
        umma_commit(empty_bar(s));  // frees the ring slot once these MMAs have read it
      }
      umma_commit(tmem_full_bar);   // accumulator complete
    }
  } else {
    // ===== epilogue: TMEM -> registers -> global (column-major: lanes = consecutive rows) =====
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    __syncwarp();
    const uint32_t row = m0 + warp * 32 + lane;
    float* crow = C + row;
#pragma unroll 1
    for (int c = 0; c < BN; c += 32) {
      uint32_t r[32];
      tmem_ld32(tmem_base + (uint32_t(warp * 32) << 16) + uint32_t(c), r);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; j++) {
        float* p = crow + size_t(n0 + c + j) * M;
        float v = __uint_as_float(r[j]);
        if (accumulate) v = *p + v;  // store.rs:52 `current + new`, fused into the epilogue
        *p = v;
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    tmem_dealloc(tmem_base, BN);
  }
}


// =================================================================================================
// v2: persistent CTA-pair kernel (cta_group::2). One cluster of two CTAs owns a 256 x 256 output tile:
// CTA r holds rows [128r, 128r+128) of the A tile and columns [128r, 128r+128) of the B tile, so each
// SM loads half of what two independent 128 x 256 CTAs would (42.7 B/clk/SM of L2->SM traffic at full
// tensor rate instead of 62.5 — the v1 kernel was pinned at the ~6.3 KB/clk chip-wide L2 output cap).
// The accumulator (128 lanes x 256 f32 columns per CTA) is double-buffered over all 512 TMEM columns:
// the epilogue of tile i drains one buffer while the MMAs of tile i+1 fill the other.
//   warp 0      TMA producer (both CTAs; completion bytes land on the LEADER's full barrier)
//   warp 1      MMA issuer   (leader CTA only; tcgen05.commit multicasts to both CTAs' barriers)
//   warp 2      TMEM allocator
//   warps 4-7   epilogue     (both CTAs, each drains its own 128 accumulator lanes)
// Tiles are handed out statically, cluster c takes tiles c, c+G, c+2G, ... in an order that walks
// `band_m` tile-rows at a time (a wave of 74 clusters covers ~8 x 9 tiles: operand slices are shared
// through L2 while they are hot).
constexpr int kThreads2 = 256;
constexpr uint32_t kAccCols = 256;

template <int BKT, int S> struct Cfg2 {
  static constexpr uint32_t kTile = 128 * BKT * 4;
  static constexpr uint32_t kStageBytes = 4 * kTile;  // A_hi, A_lo, B_hi, B_lo (this CTA's halves)
  static constexpr uint32_t kSmem = S * kStageBytes + 1024 /*align*/ + 256 /*barriers*/;
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// address of the same shared-memory offset in CTA `rank` of the cluster (shared::cluster window)
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// L2 eviction policies (the encodings createpolicy.fractional.L2::evict_{first,last} produces for fraction 1.0): the
// operand whose slices a band of tiles keeps re-reading stays, the one that streams past once per band goes first.
constexpr uint64_t kEvictFirst = 0x12F0000000000000ull, kEvictLast = 0x14F0000000000000ull;
__device__ __forceinline__ void tma2_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar_cluster, int c0, int c1, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(
          dst),
      "l"(map), "r"(bar_cluster), "r"(c0), "r"(c1), "l"(policy)
      : "memory");
}
__device__ __forceinline__ void tma2_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar_cluster, int c0, int c1, int c2, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4, %5}], [%2], %6;" ::"r"(
          dst),
      "l"(map), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2), "l"(policy)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t smem_slot, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_slot), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma2_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on the barrier at this shared-memory offset in both CTAs of the pair once the MMAs issued so far complete
__device__ __forceinline__ void umma2_commit_mc(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"(uint16_t(3))
               : "memory");
}

// Tile order: bands of `band` tile-rows (or, with along_n, tile-columns) are walked one after the other; inside a band
// the banded index runs fastest. The operand slices of the band stay in L2 while the other operand streams past once
// per band, so the band goes along the dimension whose own operand is the SMALLER one to keep resident.
struct TileSched {
  uint32_t num_mt, num_nt, total, band, along_n;
  __host__ __device__ __forceinline__ void decode(uint32_t t, uint32_t& mt, uint32_t& nt) const {
    const uint32_t banded = along_n ? num_nt : num_mt, other = along_n ? num_mt : num_nt;
    const uint32_t band_size = band * other;
    const uint32_t b = t / band_size, r = t - b * band_size;
    const uint32_t rows = band < banded - b * band ? band : banded - b * band;
    const uint32_t o = r / rows, i = b * band + (r - o * rows);
    mt = along_n ? o : i;
    nt = along_n ? i : o;
  }
};

// What a cluster works on. First the data-parallel units u = c, c+G, ... (one output tile each or, with splits == 2,
// one k half of a tile: tile = u % total, half = u / total). Then the ragged last wave of `rem` < G tiles, stream-K
// fashion but k-ALIGNED: cluster c < rem takes k-blocks [0, KB - r) of tile c (all of them sweep k in lockstep, as in
// a full wave, so the operand slices they share are fetched from HBM once), and the G - rem clusters that would have
// idled take the tails [KB - r, KB) of those tiles, `ceil(rem / (G - rem))` each (a window of k small enough to stay in
// L2). r is chosen so both kinds finish together: 64 tiles on 74 clusters cost 0.875 of a tile time instead of 1,
// 256 tiles 3.5 instead of 4. (Cutting the tiles' k-blocks into contiguous equal ranges, the textbook stream-K,
// measured SLOWER than idling: clusters at different k offsets share nothing in L2 and the GEMM went DRAM-bound.)
struct Work {
  uint32_t tile, half, kb0, kb1, part;  // k-blocks [kb0, kb1) of `tile`; part: 0 whole unit, 1 tail (stores first), 2 head (adds)
};
struct StreamK {
  uint32_t rem, tail_kb;
  uint32_t* flags;  // 8 per tile of the last wave (one per epilogue warp of the pair), all zero between launches
  unsigned long long* trace;  // debugging (GADCU_SK_TRACE): per cluster and work item, 4 timestamps of epilogue warp 0
};
struct WorkIter {
  uint32_t u, dp_units, stride, total, unit_kb;
  uint32_t next_tile, end_tile, tile_step, kb0, kb1, part;
  __host__ __device__ __forceinline__ WorkIter(uint32_t cluster, uint32_t clusters, uint32_t total_tiles, uint32_t splits, uint32_t kb_per_tile,
                                      const StreamK& sk)
      : u(cluster), dp_units((total_tiles - sk.rem) * splits), stride(clusters), total(total_tiles - sk.rem),
        unit_kb(kb_per_tile / splits), end_tile(total_tiles) {
    if (sk.rem == 0) {
      next_tile = end_tile;
    } else if (cluster < sk.rem) {
      next_tile = total + cluster, tile_step = total_tiles;  // (one item)
      kb0 = 0, kb1 = kb_per_tile - sk.tail_kb, part = 2;
    } else {
      next_tile = total + (cluster - sk.rem), tile_step = clusters - sk.rem;
      kb0 = kb_per_tile - sk.tail_kb, kb1 = kb_per_tile, part = 1;
    }
  }
  __host__ __device__ __forceinline__ bool next(Work& w) {
    if (u < dp_units) {
      const uint32_t half = u / total;
      w = Work{u - half * total, half, half * unit_kb, (half + 1) * unit_kb, 0u};
      u += stride;
      return true;
    }
    if (next_tile >= end_tile) return false;
    w = Work{next_tile, 0u, kb0, kb1, part};
    next_tile += tile_step;
    return true;
  }
};
__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu(uint32_t* p, uint32_t v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint64_t global_timer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

template <int BKT, int S>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads2, 1)
gemm_tf32x3_pair_kernel(const __grid_constant__ CUtensorMap map_ah, const __grid_constant__ CUtensorMap map_al,
                        const __grid_constant__ CUtensorMap map_bh, const __grid_constant__ CUtensorMap map_bl, float* __restrict__ C,
                        uint32_t M, uint32_t N, uint32_t K, uint32_t a_mn_major, uint32_t b_mn_major, uint32_t accumulate,
                        OperandLayout la, OperandLayout lb, uint32_t band_m, uint32_t splits, const StreamK sk, const GemmPush push,
                        const GemmEpilogue epi) {
  using cfg = Cfg2<BKT, S>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t bars = smem_base + S * cfg::kStageBytes;
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (S + s); };
  auto tmem_full_bar = [&](int a) { return bars + 8u * (2 * S + a); };
  auto tmem_empty_bar = [&](int a) { return bars + 8u * (2 * S + 2 + a); };
  const uint32_t tmem_slot = bars + 8u * (2 * S + 4);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem + S * cfg::kStageBytes + 8 * (2 * S + 4));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const uint32_t cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;
  // A work unit is one output tile or, with splits == 2, one half of its k range (units t and total + t): when the tile
  // count fills the last wave badly (256 tiles on 74 clusters: 3.46 waves cost 4), twice as many half-length units fill
  // it well (6.92 cost 7). The two halves of a tile meet in C through red.add on a zeroed C (a + b in either order is
  // the same float) or, in push mode, as two separate partial blocks that the owner adds in fixed order.
  // (splits == 2 is used by the tile-pushing mode, and without it only when stream-K is switched off)
  // M, N, K need not be multiples of the tile: the tensor maps carry the true extents (TMA fills what lies outside with
  // zeros, which add nothing to the product) and the epilogue stores only rows < M and columns < N.
  const uint32_t num_kb = (K + BKT - 1) / BKT;
  const uint32_t num_mt = (M + 255) / 256, num_nt = (N + 255) / 256;
  TileSched sched{num_mt, num_nt, num_mt * num_nt, band_m & 0xffffu, (band_m >> 16) & 1u};

  if (threadIdx.x == 0) {
    for (int s = 0; s < S; s++) {
      mbar_init(full_bar(s), 1);   // the leader's producer arms it with the byte count of both CTAs
      mbar_init(empty_bar(s), 1);  // one multicast tcgen05.commit
    }
    for (int a = 0; a < 2; a++) {
      mbar_init(tmem_full_bar(a), 1);
      mbar_init(tmem_empty_bar(a), 2 * kEpiWarps);  // every epilogue warp of both CTAs
    }
    fence_barrier_init();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_ah);
    tma_prefetch_desc(&map_al);
    tma_prefetch_desc(&map_bh);
    tma_prefetch_desc(&map_bl);
  }
  if (warp == 2) tmem_alloc2(tmem_slot, 512);
  tc_fence_before();
  cluster_sync_all();  // barrier inits visible to the peer before any remote arrive / complete_tx
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ===== TMA producer: this CTA's halves of the A and B tiles =====
    if (lane == 0) {
      uint32_t it = 0;  // k-blocks issued so far (ring position)
      // bands along N keep B's slices hot and stream A past them; bands along M the other way round
      // (measured: evict_first on the streaming operand more than doubles the DRAM reads — the tiles of a wave share
      // its slices through L2 — so the streaming side keeps the normal policy; hints 0 = none, 1 = resident evict_last)
      const uint32_t hints = band_m >> 17;
      constexpr uint64_t kNormal = 0x1000000000000000ull;
      const uint64_t pol_a = hints == 0 ? kNormal : (sched.along_n ? (hints == 2 ? kEvictFirst : kNormal) : kEvictLast);
      const uint64_t pol_b = hints == 0 ? kNormal : (sched.along_n ? kEvictLast : (hints == 2 ? kEvictFirst : kNormal));
      WorkIter work(cluster_id, num_clusters, sched.total, splits, num_kb, sk);
      Work w;
      while (work.next(w)) {
        uint32_t mt, nt;
        sched.decode(w.tile, mt, nt);  // (all first halves, then all second halves: a wave's operand slices are half as long)
        const uint32_t m0 = mt * 256 + rank * 128, n0 = nt * 256 + rank * 128;
        for (uint32_t kb = w.kb0; kb < w.kb1; kb++, it++) {
          const int s = it % S;
          const uint32_t phase = (it / S) & 1;
          mbar_wait(empty_bar(s), phase ^ 1);
          const uint32_t stage = smem_base + s * cfg::kStageBytes;
          const uint32_t sa_h = stage, sa_l = stage + cfg::kTile, sb_h = stage + 2 * cfg::kTile, sb_l = stage + 3 * cfg::kTile;
          if (rank == 0) mbar_expect_tx(full_bar(s), 2 * cfg::kStageBytes);
          const uint32_t fb = mapa(full_bar(s), 0);
          const int k0 = int(kb * BKT);
          if (a_mn_major) {
            tma2_load_3d(sa_h, &map_ah, fb, 0, k0, int(m0 / 32), pol_a);
            tma2_load_3d(sa_l, &map_al, fb, 0, k0, int(m0 / 32), pol_a);
          } else {
            tma2_load_2d(sa_h, &map_ah, fb, k0, int(m0), pol_a);
            tma2_load_2d(sa_l, &map_al, fb, k0, int(m0), pol_a);
          }
          if (b_mn_major) {
            tma2_load_3d(sb_h, &map_bh, fb, 0, k0, int(n0 / 32), pol_b);
            tma2_load_3d(sb_l, &map_bl, fb, 0, k0, int(n0 / 32), pol_b);
          } else {
            tma2_load_2d(sb_h, &map_bh, fb, k0, int(n0), pol_b);
            tma2_load_2d(sb_l, &map_bl, fb, k0, int(n0), pol_b);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: one thread of the leader CTA drives both tensor cores =====
    if (rank == 0 && lane == 0) {
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((a_mn_major ? 1u : 0u) << 15) | ((b_mn_major ? 1u : 0u) << 16) |
                             (uint32_t(256 >> 3) << 17) | (uint32_t(256 >> 4) << 24);
      uint32_t it = 0, ti = 0;
      WorkIter work(cluster_id, num_clusters, sched.total, splits, num_kb, sk);
      Work w;
      for (; work.next(w); ti++) {
        const uint32_t as = ti & 1, use = ti >> 1;
        mbar_wait(tmem_empty_bar(as), (use & 1) ^ 1);  // both epilogues have drained this accumulator
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + as * kAccCols;
        for (uint32_t kb = w.kb0; kb < w.kb1; kb++, it++) {
          const int s = it % S;
          const uint32_t phase = (it / S) & 1;
          mbar_wait(full_bar(s), phase);
          tc_fence_after();
          const uint32_t stage = smem_base + s * cfg::kStageBytes;
          const uint32_t sa_h = stage, sa_l = stage + cfg::kTile, sb_h = stage + 2 * cfg::kTile, sb_l = stage + 3 * cfg::kTile;
#pragma unroll
          for (int k = 0; k < BKT / UMMA_K; k++) {
            const uint64_t dah = make_smem_desc(sa_h + k * la.kstep, la.lbo, la.sbo, la.layout);
            const uint64_t dal = make_smem_desc(sa_l + k * la.kstep, la.lbo, la.sbo, la.layout);
            const uint64_t dbh = make_smem_desc(sb_h + k * lb.kstep, lb.lbo, lb.sbo, lb.layout);
            const uint64_t dbl = make_smem_desc(sb_l + k * lb.kstep, lb.lbo, lb.sbo, lb.layout);
            umma2_tf32(tmem_d, dal, dbh, idesc, ((kb - w.kb0) | uint32_t(k)) != 0);  // small terms first
            umma2_tf32(tmem_d, dah, dbl, idesc, 1);
            umma2_tf32(tmem_d, dah, dbh, idesc, 1);
          }
          umma2_commit_mc(empty_bar(s));  // both CTAs' ring slots are free once these MMAs have read them
        }
        umma2_commit_mc(tmem_full_bar(as));
      }
    }
  } else if (warp >= 4) {
    // ===== epilogue: this CTA's 128 accumulator lanes -> global (column-major: lanes = consecutive rows) =====
    const int q = warp - 4;  // TMEM lane quarter this warp may read
    const uint32_t te = mapa(tmem_empty_bar(0), 0);
    uint32_t ti = 0;
    WorkIter work(cluster_id, num_clusters, sched.total, splits, num_kb, sk);
    Work w;
    for (; work.next(w); ti++) {
      uint32_t mt, nt;
      sched.decode(w.tile, mt, nt);
      const uint32_t half = w.half;
      const uint32_t as = ti & 1, use = ti >> 1;
      const bool tracing = sk.trace && rank == 0 && q == 0 && lane == 0 && ti < 16;
      unsigned long long* tr = sk.trace + (size_t(cluster_id) * 16 + ti) * 4;
      if (tracing) tr[0] = global_timer_ns();
      mbar_wait(tmem_full_bar(as), use & 1);
      tc_fence_after();
      if (tracing) tr[1] = global_timer_ns();
      const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16) + as * kAccCols;
      if (push.nranks) {
        // fused reduce-scatter: this rank's partial tile goes to the staging buffer of the tile's owner, as a dense
        // 256 x 256 block (a warp store = 32 consecutive rows = 128 contiguous bytes over NVLink)
        const uint32_t tl = nt * sched.num_mt + mt;
        float* dst = push.peer[tl % push.nranks] + ((size_t(tl / push.nranks) * push.nranks + push.rank) * splits + half) * 65536u + rank * 128 +
                     q * 32 + lane;
#pragma unroll 1
        for (int c = 0; c < 256; c += 32) {
          uint32_t r[32];
          tmem_ld32(taddr + uint32_t(c), r);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; j++) dst[size_t(c + j) * 256] = __uint_as_float(r[j]);
        }
      } else {
        // How this unit's accumulator meets C: stored, added to what is there (`accumulate`: store.rs:52 `current + new`
        // fused into the epilogue; or the second part of a last-wave tile), or red.add (split-in-two on a zeroed C).
        // The two parts of a last-wave tile: the tail stores and raises the flag of ITS 32 rows; the head waits for it
        // and stores C + acc. Two terms, so the sum does not depend on who came first; the tail of a tile is among the
        // first things its cluster does, so the wait is short and cannot form a cycle.
        const bool head = w.part == 2;
        uint32_t* flag = sk.flags + ((w.tile - (sched.total - sk.rem)) * 8 + rank * 4 + q);
        if (head) {
          if (lane == 0) {
            const uint64_t t0 = global_timer_ns();
            while (ld_acquire_gpu(flag) != 1u)
              if (global_timer_ns() - t0 > 4000000000ull) __trap();  // a lost partner must not hang the GPU
          }
          __syncwarp();
        }
        if (tracing) tr[2] = global_timer_ns();
        const uint32_t row = mt * 256 + rank * 128 + q * 32 + lane;
        const bool row_ok = row < M;
        const uint32_t ncols = min(256u, N - nt * 256);
        float* crow = C + row + size_t(nt) * 256 * M;
        const bool add = head || accumulate, red = splits == 2;
        // the fused elementwise tail runs where the product is complete: on whole units and on the HEAD of a split tile
        // (the tail of a split tile stores its raw partial sum, which the head adds before transforming)
        const int ek = w.part == 1 ? int(GemmEpilogue::NONE) : epi.kind;
        const size_t elem0 = size_t(row) + size_t(nt) * 256 * M;  // (row, first column of the tile) in any [M, N] companion array
#pragma unroll 1
        for (uint32_t c = 0; c < ncols; c += 32) {
          uint32_t r[32];
          float cur[32];
          const uint32_t nj = ncols - c;  // (columns left: all 32 of the chunk unless this is N's ragged edge)
          if (add && row_ok) {
#pragma unroll
            for (int j = 0; j < 32; j++)
              if (uint32_t(j) < nj) cur[j] = __ldcg(crow + size_t(c + j) * M);  // (all in flight at once; L2, never a stale L1 line)
          }
          if (ek == GemmEpilogue::NONE) {
            tmem_ld32(taddr + c, r);
            tmem_ld_wait();
            if (row_ok) {
              if (red) {
#pragma unroll
                for (int j = 0; j < 32; j++)
                  if (uint32_t(j) < nj) atomicAdd(crow + size_t(c + j) * M, __uint_as_float(r[j]));  // (result unused: RED)
              } else {
#pragma unroll
                for (int j = 0; j < 32; j++)
                  if (uint32_t(j) < nj) crow[size_t(c + j) * M] = add ? cur[j] + __uint_as_float(r[j]) : __uint_as_float(r[j]);
              }
            }
          } else {
            float aux[32];
            if (ek != GemmEpilogue::BIAS_TANH && row_ok) {
#pragma unroll
              for (int j = 0; j < 32; j++)
                if (uint32_t(j) < nj) aux[j] = __ldcs(epi.aux + elem0 + size_t(c + j) * M);
            }
            // this chunk's 32 bias values: lane j holds column c + j's, handed round by shuffle
            float bias_mine = 0.0f;
            if (ek != GemmEpilogue::TANH_GRAD && uint32_t(lane) < nj) bias_mine = __ldg(epi.bias + nt * 256 + c + lane);
            tmem_ld32(taddr + c, r);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; j++) {
              const float b = __shfl_sync(0xffffffffu, bias_mine, j);
              if (row_ok && uint32_t(j) < nj) {
                float v = add ? cur[j] + __uint_as_float(r[j]) : __uint_as_float(r[j]);
                // (this translation unit is compiled with -fmad=false: every operation below rounds on its own, as the
                // tape's separate kernels do)
                if (ek == GemmEpilogue::BIAS_TANH) {
                  v = tanhf(v + b);
                } else if (ek == GemmEpilogue::TANH_GRAD) {
                  const float t = aux[j], s = t * t, n = 0.0f - s, e = n + 1.0f;
                  v = v * e;
                } else {
                  const float zb = v + b;
                  v = aux[j] - zb;
                }
                const size_t at = size_t(c + j) * M;
                crow[at] = v;
                if (epi.hi) {
                  uint32_t t32;
                  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t32) : "f"(v));
                  const float h = __uint_as_float(t32);
                  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t32) : "f"(v - h));
                  epi.hi[elem0 + at] = h;
                  epi.lo[elem0 + at] = __uint_as_float(t32);
                }
              }
            }
          }
        }
        if (w.part == 1) {
          __threadfence();
          __syncwarp();
          if (lane == 0) st_release_gpu(flag, 1u);
        } else if (head && lane == 0) {
          *flag = 0;  // ready for the next launch
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(te + 8u * as);
      if (tracing) tr[3] = global_timer_ns();
    }
  }
  tc_fence_before();
  cluster_sync_all();  // nobody leaves while the peer can still read this CTA's smem / signal its barriers
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc2(tmem_base, 512);
  }
}

// hi = rna_tf32(x), lo = rna_tf32(x - hi); one pass, 128-bit accesses
__global__ void __launch_bounds__(256) split_tf32_kernel(const float4* __restrict__ x, float4* __restrict__ hi, float4* __restrict__ lo,
                                                         uint64_t nvec) {
  const uint64_t stride = uint64_t(gridDim.x) * 256;
  for (uint64_t i = uint64_t(blockIdx.x) * 256 + threadIdx.x; i < nvec; i += stride) {
    const float4 v = __ldcs(x + i);
    float4 h, l;
    uint32_t t;
#define GADCU_SPLIT(c)                                                      \
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(v.c));                     \
  h.c = __uint_as_float(t);                                                 \
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(v.c - h.c));               \
  l.c = __uint_as_float(t);
    GADCU_SPLIT(x) GADCU_SPLIT(y) GADCU_SPLIT(z) GADCU_SPLIT(w)
#undef GADCU_SPLIT
    hi[i] = h;
    lo[i] = l;
  }
}

// the same for a column-major [R, C] operand whose rows do not make whole 128-byte lines (or whose base is not 16-byte
// aligned): the copies get a leading dimension of Rp = R rounded up to 32, the padding zero, so TMA can address them
__global__ void __launch_bounds__(256) split_tf32_pad_kernel(const float* __restrict__ x, float* __restrict__ hi, float* __restrict__ lo,
                                                             uint64_t R, uint64_t Rp, uint64_t n_out) {
  const uint64_t stride = uint64_t(gridDim.x) * 256;
  for (uint64_t i = uint64_t(blockIdx.x) * 256 + threadIdx.x; i < n_out; i += stride) {
    const uint64_t c = i / Rp, r = i - c * Rp;
    const float v = r < R ? __ldcs(x + c * R + r) : 0.0f;
    uint32_t t;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(v));
    const float h = __uint_as_float(t);
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(v - h));
    hi[i] = h;
    lo[i] = __uint_as_float(t);
  }
}

// out = w + d (WeightOps::add_assign, gad/src/net.rs:171-175) with the TF32 split of the sum written in the same pass:
// the weight matrix is a GEMM operand of the next step, which then finds its [hi | lo] copy ready instead of spending a
// 12 B/element pass on it (8 B/element extra here: the sum is already in registers)
__global__ void __launch_bounds__(256) add_split_tf32_kernel(const float4* __restrict__ w, const float4* __restrict__ d, float4* __restrict__ out,
                                                             float4* __restrict__ hi, float4* __restrict__ lo, uint64_t nvec) {
  const uint64_t stride = uint64_t(gridDim.x) * 256;
  for (uint64_t i = uint64_t(blockIdx.x) * 256 + threadIdx.x; i < nvec; i += stride) {
    const float4 a = w[i], b = __ldcs(d + i);
    float4 v, h, l;
    v.x = a.x + b.x; v.y = a.y + b.y; v.z = a.z + b.z; v.w = a.w + b.w;
    uint32_t t;
#define GADCU_SPLIT(c)                                                      \
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(v.c));                     \
  h.c = __uint_as_float(t);                                                 \
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(v.c - h.c));               \
  l.c = __uint_as_float(t);
    GADCU_SPLIT(x) GADCU_SPLIT(y) GADCU_SPLIT(z) GADCU_SPLIT(w)
#undef GADCU_SPLIT
    out[i] = v;
    hi[i] = h;
    lo[i] = l;
  }
}

// ---- host side ----------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  if (!fn) fail(GADCU_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
  return fn;
}

// Tensor map for one operand whose stored matrix is column-major with leading dimension `ld`.
//   mn_major: element (mn, k) at mn + ld*k  -> 3-D view [32, K, MN/32], box {32, BK, rows/32}
//   k_major : element (mn, k) at k + ld*mn  -> 2-D view [K, MN],        box {32, rows}
CUtensorMap make_map(const float* base, bool mn_major, uint64_t mn, uint64_t kdim, uint64_t ld, uint32_t rows, uint32_t bk = BK,
                     CUtensorMapSwizzle k_swizzle = CU_TENSOR_MAP_SWIZZLE_128B) {
  const CUtensorMapSwizzle mn_swizzle = CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B;
  CUtensorMap map;
  CUresult rc;
  if (mn_major) {
    cuuint64_t dims[3] = {32, kdim, mn / 32};
    cuuint64_t strides[2] = {ld * 4, 128};
    cuuint32_t box[3] = {32, bk, rows / 32};
    cuuint32_t estr[3] = {1, 1, 1};
    rc = encode_fn()(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, mn_swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  } else {
    cuuint64_t dims[2] = {kdim, mn};
    cuuint64_t strides[1] = {ld * 4};
    cuuint32_t box[2] = {bk, rows};
    cuuint32_t estr[2] = {1, 1};
    rc = encode_fn()(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, k_swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  }
  if (rc != CUDA_SUCCESS) fail(GADCU_ERR_CUDA, "cuTensorMapEncodeTiled failed with code " + std::to_string(int(rc)));
  return map;
}

void split(gadcu_ctx* ctx, const float* x, float* hi, float* lo, uint64_t R, uint64_t C, uint64_t Rp) {
  const uint64_t cap = uint64_t(ctx->sm_count) * 8;
  if (R == Rp && (reinterpret_cast<uintptr_t>(x) & 15) == 0) {
    const uint64_t nvec = R * C / 4;
    const uint64_t blocks = std::min((nvec + 255) / 256, cap);
    split_tf32_kernel<<<unsigned(blocks), 256, 0, ctx->stream>>>(reinterpret_cast<const float4*>(x), reinterpret_cast<float4*>(hi),
                                                                reinterpret_cast<float4*>(lo), nvec);
  } else {
    const uint64_t n_out = Rp * C;
    const uint64_t blocks = std::min((n_out + 255) / 256, cap * 2);
    split_tf32_pad_kernel<<<unsigned(blocks), 256, 0, ctx->stream>>>(x, hi, lo, R, Rp, n_out);
  }
  ctx->count_launch();
}

int pair_variant() {  // GADCU_GEMM_PAIR = off | 32x3 | 16x6 (default 32x3)
  static int v = [] {
    const char* e = getenv("GADCU_GEMM_PAIR");
    if (!e) return 1;
    const std::string s(e);
    return s == "off" ? 0 : s == "16x6" ? 2 : s == "16x4" ? 3 : 1;
  }();
  return v;
}

int tc_mode() {  // GADCU_GEMM=simt disables the tensor-core path (A/B testing)
  static int mode = [] {
    const char* e = getenv("GADCU_GEMM");
    return (e && std::string(e) == "simt") ? 0 : 1;
  }();
  return mode;
}

}  // namespace

bool gemm_tc_supported(const GemmArgs& g) {
  if (!tc_mode()) return false;
  if (g.dt != GADCU_F32 || g.batch != 1 || g.K == 0) return false;
  // the stored matrices must be whole (leading dimension = rows)
  const uint64_t a_rows = g.ta ? g.K : g.M, b_rows = g.tb ? g.N : g.K;
  if (g.lda != a_rows || g.ldb != b_rows) return false;
  if (g.M >= (1ull << 31) || g.N >= (1ull << 31) || g.K >= (1ull << 31)) return false;
  if (g.M % 128 == 0 && g.N % 128 == 0 && g.K % BK == 0) return true;
  // ragged shapes go through the pair kernel (edge tiles zero-filled by TMA, predicated stores) when the product is big
  // enough for 256 x 256 tiles to pay; small or skinny ones stay on the CUDA-core kernel
  return pair_variant() != 0 && g.M >= 64 && g.N >= 64 && g.K >= 32 && g.M * g.N * g.K >= (1ull << 24);
}

namespace {

// split the k range in two when that fills the last wave markedly better (see the kernel)
uint32_t choose_splits_for(int sm_count, uint64_t M, uint64_t N, uint64_t K, bool accumulate, int bk) {
  static const int split_mode = [] { const char* e = getenv("GADCU_GEMM_SPLITK"); return e ? atoi(e) : -1; }();  // -1 auto, 1 off, 2 force
  const uint64_t tiles = (M / 256) * (N / 256), kb = K / uint64_t(bk);
  if (M % 256 || N % 256 || K % uint64_t(bk)) return 1;
  if (accumulate || kb % 2 != 0 || kb / 2 < 8 || split_mode == 1) return 1;
  if (split_mode == 2) return 2;
  const uint64_t c = uint64_t(sm_count / 2);
  auto wave_eff = [&](uint64_t units) { return double(units) / double((units + c - 1) / c * c); };
  return wave_eff(2 * tiles) > wave_eff(tiles) + 0.05 ? 2 : 1;
}
uint32_t choose_splits(gadcu_ctx* ctx, const GemmArgs& g, int bk) { return choose_splits_for(ctx->sm_count, g.M, g.N, g.K, g.accumulate, bk); }

// Everything the pair kernel's work distribution depends on, decided on the host from the shape alone (so it can be
// enumerated and checked without a GPU: gadcu_debug_gemm_schedule, tests/test_gemm_schedule.py).
struct PairPlan {
  uint64_t clusters;            // 2-CTA clusters launched
  uint32_t splits;              // 1, or 2: every tile as two k halves (tile-pushing mode; stream-K switched off)
  uint32_t rem, tail_kb;        // last-wave split (WorkIter): tiles in the ragged last wave, k-blocks of their tails; 0 = none
  uint32_t band_arg;            // band | along_n << 16 | l2 hints << 17 (TileSched)
  uint32_t num_kb;              // k-blocks per tile
};
PairPlan plan_pair(int sm_count, bool comm_inflight, uint64_t M, uint64_t N, uint64_t K, int bkt, bool accumulate,
                   int push_splits /* 0: not the tile-pushing mode */, bool have_flags, bool no_split = false /* an epilogue rides along: no red.add halves */) {
  PairPlan p{};
  const uint64_t tiles = ((M + 255) / 256) * ((N + 255) / 256);
  // while an overlapped all-reduce is in flight a few SM pairs are left to its CTAs (a persistent CTA per SM would
  // otherwise keep them out until the GEMM ends)
  static const uint64_t reserve = [] { const char* e = getenv("GADCU_DP_RESERVE_CLUSTERS"); return e ? uint64_t(atoi(e)) : uint64_t(0); }();
  const uint64_t usable = uint64_t(sm_count / 2) - (comm_inflight ? reserve : 0);
  // the ragged last wave is shared out stream-K fashion (see WorkIter; GADCU_GEMM_STREAMK=0 for the older split-in-two)
  static const bool streamk_on = [] { const char* e = getenv("GADCU_GEMM_STREAMK"); return !e || atoi(e) != 0; }();
  static const uint64_t min_tail = [] { const char* e = getenv("GADCU_GEMM_SK_MIN_TAIL"); return e ? uint64_t(atoi(e)) : uint64_t(8); }();
  const bool streamk = streamk_on && !push_splits && have_flags;
  const uint64_t kb_per_tile = (K + uint64_t(bkt) - 1) / uint64_t(bkt);
  p.num_kb = uint32_t(kb_per_tile);
  p.clusters = std::min<uint64_t>(tiles * 2, usable);  // (a cluster without work exits at once)
  if (streamk && tiles % p.clusters != 0) {
    const uint64_t rem = tiles % p.clusters, helpers = p.clusters - rem, per_helper = (rem + helpers - 1) / helpers;
    const uint64_t tail = kb_per_tile / (per_helper + 1);  // head (KB - tail) = per_helper tails
    // worth it when it shortens the launch by 3 % or more: 512 tiles (6.92 waves: 1 % to gain) measured 5 % SLOWER split
    // than whole — twelve 9-block tails per helper are epilogue-bound and the heads' read-add-store epilogue is exposed
    const double gain = double(tail) / double((tiles / p.clusters + 1) * kb_per_tile);
    // ... and not for a single partial wave that already covers most of the chip (64 tiles on 74 clusters, the 1024-row
    // shards of an 8-GPU step): no gain alone (0.132 ms either way: the idle SMs' power goes to the busy ones) and 2.5 %
    // slower inside the 8-GPU step (516 vs 503 steps/s), where the idle SMs are what the exchange kernels run on
    static const bool lone_ok = [] { const char* e = getenv("GADCU_GEMM_SK_LONE"); return e && atoi(e) != 0; }();  // (A/B switch)
    const bool lone_wide_wave = !lone_ok && tiles < p.clusters && tiles * 2 > p.clusters;
    if (tail >= min_tail && gain >= 0.03 && !lone_wide_wave) p.rem = uint32_t(rem), p.tail_kb = uint32_t(tail);
  }
  p.splits = push_splits ? uint32_t(push_splits) : ((streamk || no_split) ? 1u : choose_splits_for(sm_count, M, N, K, accumulate, bkt));  // (push: the exchange laid its staging out for this)
  // band of tiles whose operand slices (hi + lo, 8 bytes per element) should fit in about half of L2 (126 MB), along
  // the dimension with the smaller operand: that operand is then read from HBM once, the other once per band
  static const uint32_t band_override = [] { const char* e = getenv("GADCU_GEMM_BAND"); return e ? uint32_t(atoi(e)) : 0u; }();
  const uint64_t slice_bytes = 256ull * (K / p.splits) * 8;  // one tile-row of A or tile-column of B (of one k half)
  uint32_t band = uint32_t(std::max<uint64_t>(1, (64ull << 20) / std::max<uint64_t>(slice_bytes, 1)));
  const uint32_t along_n = N < M ? 1u : 0u;   // fewer tile-columns than tile-rows: B is the smaller operand
  band = std::min<uint32_t>(band, uint32_t(((along_n ? N : M) + 255) / 256));
  if (band_override) band = band_override & 0xffffu;
  static const uint32_t l2_hints = [] { const char* e = getenv("GADCU_GEMM_L2_HINTS"); return e ? uint32_t(atoi(e)) & 3u : 0u; }();
  p.band_arg = band | (((band_override ? (band_override >> 16) : along_n) & 1u) << 16) | (l2_hints << 17);
  return p;
}

template <int BKT, int S>
void launch_pair(gadcu_ctx* ctx, const GemmArgs& g, const float* ah, const float* al, const float* bh, const float* bl, bool a_mn,
                 bool b_mn, uint64_t lda, uint64_t ldb /* leading dimensions of the split copies (multiples of 32) */) {
  // K-major tiles have rows of BKT*4 bytes: 128 B -> SWIZZLE_128B, 64 B -> SWIZZLE_64B (8-row groups either way)
  const OperandLayout k_layout = BKT == 32 ? OperandLayout{16u, 1024u, UMMA_K * 4u, 2u} : OperandLayout{16u, 512u, UMMA_K * 4u, 4u};
  const CUtensorMapSwizzle k_swz = BKT == 32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  const OperandLayout mn_layout{BKT * 128u, 512u, 1024u, 1u};
  const OperandLayout la = a_mn ? mn_layout : k_layout, lb = b_mn ? mn_layout : k_layout;
  // extents as stored: the contiguous dimension is the padded one (zeros), the strided one is exact (TMA zero-fills past it)
  const uint64_t a_mn_ext = a_mn ? lda : g.M, a_k_ext = a_mn ? g.K : lda, b_mn_ext = b_mn ? ldb : g.N, b_k_ext = b_mn ? g.K : ldb;
  CUtensorMap mah = make_map(ah, a_mn, a_mn_ext, a_k_ext, lda, 128, BKT, k_swz), mal = make_map(al, a_mn, a_mn_ext, a_k_ext, lda, 128, BKT, k_swz);
  CUtensorMap mbh = make_map(bh, b_mn, b_mn_ext, b_k_ext, ldb, 128, BKT, k_swz), mbl = make_map(bl, b_mn, b_mn_ext, b_k_ext, ldb, 128, BKT, k_swz);
  const PairPlan plan = plan_pair(ctx->sm_count, ctx->comm_inflight.load(std::memory_order_relaxed), g.M, g.N, g.K, BKT, g.accumulate,
                                  g.push ? int(g.push->splits) : 0, ctx->gemm_flags != nullptr, g.epi != nullptr);
  const uint64_t clusters = plan.clusters;
  StreamK sk{plan.rem, plan.tail_kb, ctx->gemm_flags, nullptr};
  static const int trace_at = [] { const char* e = getenv("GADCU_SK_TRACE"); return e ? atoi(e) : 0; }();  // trace the n-th launch
  static int trace_count = 0;
  const bool trace_on = trace_at && ++trace_count == trace_at;
  if (trace_on) {
    GADCU_CUDA(cudaMallocManaged(&sk.trace, 74 * 16 * 4 * 8));
    memset(sk.trace, 0, 74 * 16 * 4 * 8);
  }
  const uint32_t splits = plan.splits, band_arg = plan.band_arg;
  if (splits == 2 && !g.push) GADCU_CUDA(cudaMemsetAsync(g.C, 0, size_t(g.M) * g.N * 4, ctx->stream));
  static std::once_flag once;
  std::call_once(once, [] {
    cudaFuncSetAttribute(gemm_tf32x3_pair_kernel<BKT, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg2<BKT, S>::kSmem);
  });
  gemm_tf32x3_pair_kernel<BKT, S><<<unsigned(2 * clusters), kThreads2, Cfg2<BKT, S>::kSmem, ctx->stream>>>(
      mah, mal, mbh, mbl, static_cast<float*>(g.C), uint32_t(g.M), uint32_t(g.N), uint32_t(g.K), a_mn, b_mn, g.accumulate, la, lb, band_arg,
      splits, sk, g.push ? *g.push : GemmPush{}, g.epi ? *g.epi : GemmEpilogue{});
  if (trace_on) {
    cudaStreamSynchronize(ctx->stream);
    unsigned long long t0 = ~0ull;
    for (int i = 0; i < 74 * 16 * 4; i++) if (sk.trace[i] && sk.trace[i] < t0) t0 = sk.trace[i];
    fprintf(stderr, "sk trace M=%llu N=%llu K=%llu clusters=%llu rem=%u tail_kb=%u\n", (unsigned long long)g.M, (unsigned long long)g.N,
            (unsigned long long)g.K, (unsigned long long)clusters, sk.rem, sk.tail_kb);
    for (int c = 0; c < 74; c++) {
      fprintf(stderr, " c%02d:", c);
      for (int w = 0; w < 16 && sk.trace[(c * 16 + w) * 4]; w++) {
        unsigned long long* t = sk.trace + (c * 16 + w) * 4;
        fprintf(stderr, " [%.1f full %.1f waited %.1f done %.1f]", (t[0] - t0) / 1e3, (t[1] - t0) / 1e3, t[2] ? (t[2] - t0) / 1e3 : 0.0, (t[3] - t0) / 1e3);
      }
      fprintf(stderr, "\n");
    }
    cudaFree(sk.trace);
  }
}

}  // namespace

// [hi | lo] of one operand (stored column-major [R, C]): from the owning buffer's cache when it has a current one, else
// a fresh split that is left on the buffer for the other GEMMs of this tape. The copies have leading dimension
// `ld` = R rounded up to 32 (see split_tf32_pad_kernel). `keep` holds the split alive until the launch is issued.
static const float* split_operand(gadcu_ctx* ctx, const float* x, uint64_t R, uint64_t C, Buffer* owner, ArrayRef& keep, uint64_t& ld) {
  ld = (R + 31) / 32 * 32;
  const uint64_t elems = ld * C;
  // The split is a function of the buffer's contents alone: it stays valid until the buffer is written in place
  // (Buffer::mutated drops it) or dies. Activations and gradients die with their tape, so their copies go with them; a
  // weight matrix keeps its copy across steps for as long as nothing updates it, and WeightOps::add_assign — the one
  // thing that does — refreshes the copy in the same pass (add_assign_refresh_split below). GADCU_SPLIT_CACHE=tape
  // restores round 1's rule (a copy is trusted only within the tape that made it).
  static const bool per_tape = [] { const char* e = getenv("GADCU_SPLIT_CACHE"); return e && std::string(e) == "tape"; }();
  const uint64_t epoch = ctx->epoch.load(std::memory_order_relaxed);
  // (not while a CUDA graph is being captured: the replayed graph must redo the split of whatever the buffer holds then)
  if (!ctx->capturing && owner && owner->split && (!per_tape || owner->split_epoch == epoch) && owner->split_elems == elems && owner->split_rows == R)
    return static_cast<const float*>(owner->split->ptr);
  keep = new_array(ctx, GADCU_F32, Dim4(2 * elems));
  float* hi = static_cast<float*>(keep->ptr());
  split(ctx, x, hi, hi + elems, R, C, ld);
  if (owner && !ctx->capturing) {
    owner->mutated();
    owner->split = keep->buf.get();
    owner->split->retain();
    owner->split_epoch = epoch;
    owner->split_elems = elems;
    owner->split_rows = R;
  }
  return hi;
}

// The work list of every cluster for an f32 product of this shape, computed on the host with the kernel's own iterator
// (no device needed). Rows of 6: cluster, tile row, tile column, first k-block, end k-block, part (0 whole unit, 1 tail,
// 2 head). info[8] = clusters, splits, rem, tail_kb, k-blocks per tile, band, along_n, tile count. Returns the number of
// rows the schedule has (which may exceed `cap`; only `cap` are written).
size_t gemm_schedule_debug(int sm_count, uint64_t M, uint64_t N, uint64_t K, int push_splits, uint32_t* rows, size_t cap, uint32_t* info) {
  const int bkt = pair_variant() == 1 || pair_variant() == 0 ? 32 : 16;
  const PairPlan plan = plan_pair(sm_count, false, M, N, K, bkt, false, push_splits, true);
  const uint32_t num_mt = uint32_t((M + 255) / 256), num_nt = uint32_t((N + 255) / 256);
  const TileSched sched{num_mt, num_nt, num_mt * num_nt, plan.band_arg & 0xffffu, (plan.band_arg >> 16) & 1u};
  if (info) {
    const uint32_t v[8] = {uint32_t(plan.clusters), plan.splits, plan.rem, plan.tail_kb, plan.num_kb, sched.band, sched.along_n, sched.total};
    for (int i = 0; i < 8; i++) info[i] = v[i];
  }
  const StreamK sk{plan.rem, plan.tail_kb, nullptr, nullptr};
  size_t n = 0;
  for (uint32_t c = 0; c < plan.clusters; c++) {
    WorkIter work(c, uint32_t(plan.clusters), sched.total, plan.splits, plan.num_kb, sk);
    Work w;
    while (work.next(w)) {
      if (n < cap) {
        uint32_t mt, nt;
        sched.decode(w.tile, mt, nt);
        const uint32_t row[6] = {c, mt, nt, w.kb0, w.kb1, w.part};
        for (int i = 0; i < 6; i++) rows[n * 6 + i] = row[i];
      }
      n++;
    }
  }
  return n;
}

// `dst` (the buffer the sum goes to: `src` itself for an in-place update, a fresh one for a copy-on-write one) receives
// src + d and, if `src` carried a current, unpadded TF32 split, a refreshed split of the sum. Returns false (nothing
// launched) when that does not apply and the caller should do the plain add.
bool add_assign_refresh_split(gadcu_ctx* ctx, Buffer* src, const float* d, Buffer* dst, uint64_t n) {
  if (!src->split || src->split_elems != n || src->split_rows % 32 != 0 || n % 4 != 0) return false;
  if ((reinterpret_cast<uintptr_t>(src->ptr) | reinterpret_cast<uintptr_t>(d) | reinterpret_cast<uintptr_t>(dst->ptr)) & 15) return false;
  Buffer* split = src->split;  // [hi | lo], 2 n floats
  ArrayRef fresh;
  if (dst != src || split->refs.load() > 1) {  // (a GEMM still in flight may hold the old copy: stream order covers reads, but a new buffer is simplest)
    fresh = new_array(ctx, GADCU_F32, Dim4(2 * n));
    split = fresh->buf.get();
  }
  float* hi = static_cast<float*>(split->ptr);
  const uint64_t nvec = n / 4;
  const uint64_t blocks = std::min<uint64_t>((nvec + 255) / 256, uint64_t(ctx->sm_count) * 8);
  ctx->segment_begin();
  add_split_tf32_kernel<<<unsigned(blocks), 256, 0, ctx->stream>>>(static_cast<const float4*>(src->ptr), reinterpret_cast<const float4*>(d),
                                                                  static_cast<float4*>(dst->ptr), reinterpret_cast<float4*>(hi),
                                                                  reinterpret_cast<float4*>(hi + n), nvec);
  ctx->count_launch();
  const uint64_t rows = src->split_rows;
  if (fresh) {
    dst->mutated();
    dst->split = split;
    split->retain();
  }
  dst->split_elems = n;
  dst->split_rows = rows;
  dst->split_epoch = ctx->epoch.load(std::memory_order_relaxed);
  return true;
}

uint32_t gemm_push_splits(gadcu_ctx* ctx, const GemmArgs& g) { return choose_splits(ctx, g, pair_variant() == 1 ? 32 : 16); }

bool gemm_epilogue_supported(const GemmArgs& g) {
  if (!gemm_tc_supported(g) || pair_variant() == 0 || g.accumulate || g.push) return false;
  const bool ragged = g.M % 128 != 0 || g.N % 128 != 0 || g.K % BK != 0;
  return ragged || (g.M % 256 == 0 && g.N % 256 == 0);  // the shapes launch_gemm_tc gives to the pair kernel
}

bool gemm_push_supported(const GemmArgs& g) {
  return gemm_tc_supported(g) && pair_variant() != 0 && g.M % 256 == 0 && g.N % 256 == 0 && !g.accumulate;
}

void launch_gemm_tc(gadcu_ctx* ctx, const GemmArgs& g) {
  const bool a_mn = !g.ta;  // stored [M,K] col-major: M contiguous
  const bool b_mn = g.tb;   // stored [N,K] col-major: N contiguous
  ArrayRef keep_a, keep_b;
  uint64_t lda = 0, ldb = 0;
  const float* ah = split_operand(ctx, static_cast<const float*>(g.A), a_mn ? g.M : g.K, a_mn ? g.K : g.M, g.a_buf, keep_a, lda);
  const float* bh = split_operand(ctx, static_cast<const float*>(g.B), b_mn ? g.N : g.K, b_mn ? g.K : g.N, g.b_buf, keep_b, ldb);
  const float *al = ah + lda * (a_mn ? g.K : g.M), *bl = bh + ldb * (b_mn ? g.K : g.N);

  const bool ragged = g.M % 128 != 0 || g.N % 128 != 0 || g.K % BK != 0;
  if (pair_variant() && (ragged || (g.M % 256 == 0 && g.N % 256 == 0))) {
    switch (pair_variant()) {
      case 2: launch_pair<16, 6>(ctx, g, ah, al, bh, bl, a_mn, b_mn, lda, ldb); break;
      case 3: launch_pair<16, 4>(ctx, g, ah, al, bh, bl, a_mn, b_mn, lda, ldb); break;
      default: launch_pair<32, 3>(ctx, g, ah, al, bh, bl, a_mn, b_mn, lda, ldb); break;
    }
  } else {
    if (g.push) fail(GADCU_ERR_UNSUPPORTED, "gemm: the tile-pushing epilogue needs M and N to be multiples of 256");
    if (g.epi) fail(GADCU_ERR_UNSUPPORTED, "gemm: a fused elementwise epilogue needs the pair kernel (gemm_epilogue_supported)");
    const OperandLayout mn_layout{BK * 128u, 512u, 1024u, 1u};
    const OperandLayout k_layout{16u, 1024u, UMMA_K * 4u, 2u};
    const OperandLayout la = a_mn ? mn_layout : k_layout, lb = b_mn ? mn_layout : k_layout;
    const int BN = (g.N % 256 == 0 && (g.M / BM) * (g.N / 256) >= uint64_t(ctx->sm_count)) ? 256 : 128;
    CUtensorMap mah = make_map(ah, a_mn, g.M, g.K, g.lda, BM), mal = make_map(al, a_mn, g.M, g.K, g.lda, BM);
    CUtensorMap mbh = make_map(bh, b_mn, g.N, g.K, g.ldb, BN), mbl = make_map(bl, b_mn, g.N, g.K, g.ldb, BN);
    dim3 grid(unsigned(g.M / BM), unsigned(g.N / BN));
    if (BN == 256) {
      static std::once_flag once;
      std::call_once(once, [] { cudaFuncSetAttribute(gemm_tf32x3_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<256>::kSmem); });
      gemm_tf32x3_kernel<256><<<grid, kThreads, Cfg<256>::kSmem, ctx->stream>>>(mah, mal, mbh, mbl, static_cast<float*>(g.C), uint32_t(g.M),
                                                                               uint32_t(g.N), uint32_t(g.K), a_mn, b_mn, g.accumulate, la, lb);
    } else {
      static std::once_flag once;
      std::call_once(once, [] { cudaFuncSetAttribute(gemm_tf32x3_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<128>::kSmem); });
      gemm_tf32x3_kernel<128><<<grid, kThreads, Cfg<128>::kSmem, ctx->stream>>>(mah, mal, mbh, mbl, static_cast<float*>(g.C), uint32_t(g.M),
                                                                               uint32_t(g.N), uint32_t(g.K), a_mn, b_mn, g.accumulate, la, lb);
    }
  }
  ctx->count_launch();
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) fail(GADCU_ERR_CUDA, std::string("gemm_tf32x3 launch: ") + cudaGetErrorString(err));
}

}  // namespace k
}  // namespace gadcu
