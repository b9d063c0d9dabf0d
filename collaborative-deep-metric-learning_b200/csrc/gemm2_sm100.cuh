// CTA-pair (cta_group::2) variant of the tcgen05 GEMM core.
//
// Two CTAs of a 2-CTA cluster (same TPC) cooperate on one 256 x 256 output tile: each CTA stages its own 128 rows of A
// and its own 128-row half of B (16 KB + 16 KB per k-block instead of 16 KB + 32 KB), the leader CTA issues
// tcgen05.mma.cta_group::2 (M = 256, N = 256) which reads both halves of B from the two shared memories, and each CTA
// ends up with its 128 x 256 slice of the accumulator in its own TMEM.  Per flop this moves 1.5x fewer operand bytes
// from L2 into shared memory and leaves room for 6 pipeline stages instead of 4 -- the long-K tower GEMMs (forward
// layer 1, weight gradient 1) are L2-bandwidth bound with the single-CTA kernel.
//
// Barrier protocol (all mbarriers live at the same offsets in both CTAs):
//   full[s]   used in the LEADER only, count 2: one producer arrival per CTA; TMA bytes of both CTAs complete on it
//   empty[s]  per CTA, count 1: tcgen05.commit multicast to both CTAs
//   tfull[a]  per CTA, count 1: tcgen05.commit multicast to both CTAs
//   tempty[a] LEADER only, count 16: the 8 epilogue warps of each CTA arrive (the peer's arrive remotely)
#pragma once
#include "gemm_sm100.cuh"

namespace cdml {

constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;  // clears the CTA-rank bit of a shared::cluster address -> leader CTA
constexpr int kBN2 = 256;                       // tile columns of the pair
constexpr int kStages2 = 6;

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  // non-.aligned forms: the single-lane role loops leave warps 0/1 formally divergent at the kernel's tail
  asm volatile("barrier.cluster.arrive.release;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish2() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_commit2(uint32_t bar) {  // arrive on `bar` in BOTH CTAs of the pair
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"(static_cast<uint16_t>(3))
               : "memory");
}
__device__ __forceinline__ void umma2_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// TMA load whose completion bytes are credited to the LEADER CTA's mbarrier
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar & kPeerBitMask), "r"(c0), "r"(c1)
      : "memory");
}

struct Gemm2Smem {
  static constexpr uint32_t kABytes = kBM * kBK * 2;          // 16 KB: this CTA's 128 rows of A
  static constexpr uint32_t kBBytes = (kBN2 / 2) * kBK * 2;   // 16 KB: this CTA's 128-row half of B
  static constexpr uint32_t kEpiOff = kStages2 * (kABytes + kBBytes);
  static constexpr uint32_t kBarOff = kEpiOff + kEpiStageBytes;
  static constexpr uint32_t kNumBars = 2 * kStages2 + 4;
  static constexpr uint32_t kTotal = kBarOff + kNumBars * 8 + 16 + 1024;
};

// GemmShape semantics: m_tiles counts 256-row PAIR tiles; everything else as in the single-CTA kernel.
template <int kAMajorMN, int kBMajorMN, class Epi>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm2_tcgen05_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                     const GemmShape s, const Epi epi) {
  using L = Gemm2Smem;
  constexpr int kStages = kStages2;
  constexpr int BN = kBN2;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_u32 = smem_u32(smem_raw);
  const uint32_t base = (raw_u32 + 1023u) & ~1023u;
  const uint32_t bar0 = base + L::kBarOff;
  auto a_smem = [&](int st) { return base + st * L::kABytes; };
  auto b_smem = [&](int st) { return base + kStages * L::kABytes + st * L::kBBytes; };
  auto full_bar = [&](int st) { return bar0 + 8u * st; };
  auto empty_bar = [&](int st) { return bar0 + 8u * (kStages + st); };
  auto tfull_bar = [&](int i) { return bar0 + 8u * (2 * kStages + i); };
  auto tempty_bar = [&](int i) { return bar0 + 8u * (2 * kStages + 2 + i); };
  const uint32_t tmem_slot = bar0 + 8u * L::kNumBars;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - raw_u32));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int cluster_id = blockIdx.x >> 1;
  const int num_clusters = gridDim.x >> 1;
  constexpr uint32_t kTmemCols = 512;

  cluster_sync_all();  // both CTAs resident before the pair allocates tensor memory
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_a);
    tma_prefetch_desc(&tma_b);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(full_bar(i), 2);
      mbar_init(empty_bar(i), 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(tfull_bar(i), 1);
      mbar_init(tempty_bar(i), 16);
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc2(tmem_slot, kTmemCols);
    tmem_relinquish2();
  }
  if (threadIdx.x == 96) epi.block_begin(base + L::kEpiOff);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  const int tiles = s.m_tiles * s.n_tiles;
  const int units = tiles * s.num_splits;

  if (warp == 0 && lane == 0) {
    // ===================== TMA producer (both CTAs) =====================
    int stage = 0;
    uint32_t phase = 0;
    for (int u = cluster_id; u < units; u += num_clusters) {
      const int split = u / tiles, t = u - split * tiles;
      const int m0 = (s.m_fastest ? t % s.m_tiles : t / s.n_tiles) * (2 * kBM) + static_cast<int>(rank) * kBM;
      const int n0 = (s.m_fastest ? t / s.m_tiles : t % s.n_tiles) * BN + static_cast<int>(rank) * (BN / 2);
      const int kb0 = split * s.kb_per_split;
      const int kb1 = min(s.num_kb, kb0 + s.kb_per_split);
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(empty_bar(stage), phase ^ 1, 100 + stage);
        const int k0 = kb * kBK;
        if constexpr (kAMajorMN == 0) {
          tma_load_2d_pair(a_smem(stage), &tma_a, full_bar(stage), k0, m0);
        } else {
#pragma unroll
          for (int i = 0; i < kBM / 64; ++i)
            tma_load_2d_pair(a_smem(stage) + i * (kBK * 128), &tma_a, full_bar(stage), m0 + i * 64, k0);
        }
        if constexpr (kBMajorMN == 0) {
          tma_load_2d_pair(b_smem(stage), &tma_b, full_bar(stage), k0, n0);
        } else {
#pragma unroll
          for (int i = 0; i < BN / 128; ++i)
            tma_load_2d_pair(b_smem(stage) + i * (kBK * 128), &tma_b, full_bar(stage), n0 + i * 64, k0);
        }
        if (leader) mbar_arrive_expect_tx(full_bar(stage), 2 * (L::kABytes + L::kBBytes));
        else mbar_arrive_cluster(full_bar(stage) & kPeerBitMask);
        if (++stage == kStages) stage = 0, phase ^= 1;
      }
    }
  } else if (warp == 1 && lane == 0) {
    if (leader) {
      // ===================== MMA issuer (leader CTA only) =====================
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      constexpr uint32_t kAStep = kAMajorMN ? (kUK * 128) : (kUK * 2);
      constexpr uint32_t kBStep = kBMajorMN ? (kUK * 128) : (kUK * 2);
      constexpr uint32_t kALbo = kAMajorMN ? (kBK * 128) : 16;
      constexpr uint32_t kBLbo = kBMajorMN ? (kBK * 128) : 16;
      for (int u = cluster_id; u < units; u += num_clusters, ++it) {
        const int split = u / tiles;
        const int kb0 = split * s.kb_per_split;
        const int kb1 = min(s.num_kb, kb0 + s.kb_per_split);
        const int as = it & 1;
        const uint32_t ap = (it >> 1) & 1;
        mbar_wait(tempty_bar(as), ap ^ 1, 200 + as);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + as * BN;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(full_bar(stage), phase, 300 + stage);
          tc_fence_after();
#pragma unroll
          for (int k = 0; k < kBK / kUK; ++k) {
            const uint64_t ad = make_smem_desc(a_smem(stage) + k * kAStep, kALbo, 1024);
            const uint64_t bd = make_smem_desc(b_smem(stage) + k * kBStep, kBLbo, 1024);
            umma2_f16(tmem_d, ad, bd, s.idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          tc_commit2(empty_bar(stage));
          if (kb == kb1 - 1) tc_commit2(tfull_bar(as));
          if (++stage == kStages) stage = 0, phase ^= 1;
        }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue (both CTAs, own 128 rows) =====================
    const int q = warp & 3;
    const int half = (warp - 4) >> 2;
    const uint32_t stg = base + L::kEpiOff + (warp - 4) * (Epi::kSplitColumns ? 2048u : 4096u);
    int it = 0;
    for (int u = cluster_id; u < units; u += num_clusters, ++it) {
      const int split = u / tiles, t = u - split * tiles;
      const int m0 = (s.m_fastest ? t % s.m_tiles : t / s.n_tiles) * (2 * kBM) + static_cast<int>(rank) * kBM;
      const int n0 = (s.m_fastest ? t / s.m_tiles : t % s.n_tiles) * BN;
      const int as = it & 1;
      const uint32_t ap = (it >> 1) & 1;
      const int ec0 = Epi::kSplitColumns ? half * (BN / 64) : 0, ec1 = Epi::kSplitColumns ? (half + 1) * (BN / 64) : BN / 32;
      typename Epi::State est;
      if (Epi::kSplitColumns || half == 0) {
        epi.cols(n0, s, ec0, ec1, stg);
        epi.pre(est, m0 + q * 32 + lane, n0, s, ec0, ec1, stg);
      }
      mbar_wait(tfull_bar(as), ap, 400 + as);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * BN;
      if (Epi::kSplitColumns || half == 0) epi.run(taddr, m0 + q * 32 + lane, n0, split, s, ec0, ec1, stg, est);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(tempty_bar(as) & kPeerBitMask);  // the leader's barrier collects both CTAs
    }
  }

  // Nobody may leave while the peer can still touch this CTA's shared memory / barriers.
  tc_fence_before();
  cluster_sync_all();
  if (threadIdx.x == 96) epi.block_end(base + L::kEpiOff);
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc2(tmem_base, kTmemCols);
  }
}

}  // namespace cdml
