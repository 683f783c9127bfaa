// 2-CTA (cta_group::2) variant of the persistent bf16 GEMM: a CTA PAIR (two SMs of one TPC, launched as a cluster of 2)
// computes one 256 x 256 tile.  CTA r of the pair stages A rows [128 r, 128 r + 128) and only HALF of the W tile (N rows
// [128 r, 128 r + 128)); the leader's single thread issues tcgen05.mma.cta_group::2 (UMMA 256 x 256 x 16) which reads both
// CTAs' shared memory and accumulates into both CTAs' TMEM (128 lanes x 256 columns each).
//
// Why: at cta_group::1 the 128 x 256 x 16 UMMA reads 12 KB of shared memory per instruction and is bound by the
// shared-memory bandwidth of ONE SM (ncu: tensor pipe "active" 82% of the time at 62% of the nominal rate,
// profiles/r1d_gateup_gemm_ncu.txt); with the pair each SM feeds 8 KB per instruction and fetches a third fewer bytes
// from L2 per FLOP.
//
// Pipeline (same roles as gemm_tcgen05.cuh): both CTAs' producer threads issue their own TMA loads, all of which
// complete_tx on the LEADER's full barrier; the leader's MMA thread commits with .multicast::cluster so that both CTAs'
// empty / accumulator-full barriers advance; the epilogue warps of both CTAs release the accumulator on the leader's
// barrier (remote mbarrier.arrive).  Epilogues and rounding points are shared with the 1-CTA kernel.
#pragma once
#include "gemm_tcgen05.cuh"

namespace cvb {

__device__ __forceinline__ uint32_t cluster_rank_2sm() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_2sm() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t map_to_cta(uint32_t smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish_2sm() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// TMA load whose bytes are accounted on the barrier at cluster address `bar_cluster_addr` (the leader CTA's)
__device__ __forceinline__ void tma_load_2d_2sm(void* smem_dst, const void* desc, uint32_t bar_cluster_addr, int32_t crd0,
                                                int32_t crd1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(desc)), "r"(bar_cluster_addr), "r"(crd0), "r"(crd1)
      : "memory");
}
__device__ __forceinline__ void umma_bf16_2sm(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                              uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive (once all previously issued MMAs have completed) on the barrier at the same offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(static_cast<uint16_t>(3))
      : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t bar_cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster_addr) : "memory");
}

constexpr int GEMM2_BN = 256;
constexpr int GEMM2_STAGES = 6;
constexpr uint32_t GEMM2_A_BYTES = 128 * GEMM_BK * 2;   // this CTA's 128 rows of A
constexpr uint32_t GEMM2_B_BYTES = 128 * GEMM_BK * 2;   // this CTA's half (128 rows) of the 256-row W tile
constexpr uint32_t GEMM2_STAGE_BYTES = GEMM2_A_BYTES + GEMM2_B_BYTES;
constexpr uint32_t GEMM2_SMEM = GEMM2_STAGES * GEMM2_STAGE_BYTES + 256 + 1024;

template <int EPI>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_tcgen05_2sm(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmArgs g) {
  constexpr int STAGES = GEMM2_STAGES;
  constexpr uint32_t IDESC = make_idesc(/*bf16*/ 1, 256, GEMM2_BN);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * GEMM2_STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_rank_2sm();
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const int m_tiles = (g.M + 255) / 256;
  const int n_tiles = (g.N + GEMM2_BN - 1) / GEMM2_BN;
  const int k_blocks = (g.K + GEMM_BK - 1) / GEMM_BK;
  const int num_tiles = m_tiles * n_tiles;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < STAGES; ++s) {
        mbar_init(&full_bar[s], 1);
        mbar_init(&empty_bar[s], 1);
      }
      for (int a = 0; a < 2; ++a) {
        mbar_init(&tfull_bar[a], 1);
        mbar_init(&tempty_bar[a], 256);  // 4 epilogue warps of BOTH CTAs (only the leader's copy is used)
      }
      fence_barrier_init();
      fence_proxy_async();
    }
    __syncwarp();
    tmem_alloc_2sm(tmem_slot, 512);
    tmem_relinquish_2sm();
  }
  tc_fence_before();
  cluster_sync_2sm();  // barriers of both CTAs are initialised before any remote arrive / multicast commit
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch();

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer (both CTAs)
    if (lane == 0) {
      pdl_wait();
      uint32_t stage = 0, phase = 0;
      for (int tile = pair; tile < num_tiles; tile += npairs) {
        const int mt = tile % m_tiles, nt = tile / m_tiles;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * GEMM2_STAGE_BYTES;
          uint8_t* sb = sa + GEMM2_A_BYTES;
          if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * GEMM2_STAGE_BYTES);  // both CTAs' bytes
          const uint32_t full_leader = map_to_cta(smem_u32(&full_bar[stage]), 0);
          tma_load_2d_2sm(sa, &tmA, full_leader, kb * GEMM_BK, mt * 256 + static_cast<int>(rank) * 128);
          tma_load_2d_2sm(sb, &tmB, full_leader, kb * GEMM_BK, nt * GEMM2_BN + static_cast<int>(rank) * 128);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (one thread of the leader CTA)
    if (lane == 0 && rank == 0) {
      uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
      for (int tile = pair; tile < num_tiles; tile += npairs) {
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * GEMM2_BN;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + stage * GEMM2_STAGE_BYTES);
          const uint64_t adesc = make_desc_kmajor_sw128(a_addr);
          const uint64_t bdesc = make_desc_kmajor_sw128(a_addr + GEMM2_A_BYTES);
#pragma unroll
          for (int k = 0; k < GEMM_BK / 16; ++k) umma_bf16_2sm(d_tmem, adesc + 2 * k, bdesc + 2 * k, IDESC, (kb | k) != 0 ? 1u : 0u);
          umma_commit_2sm(&empty_bar[stage]);  // frees the slot in BOTH CTAs
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit_2sm(&tfull_bar[acc]);  // accumulator complete -> both CTAs' epilogues
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue (4 warps, this CTA's 128 rows)
    pdl_wait();
    const int q = warp & 3;
    const int row_in_tile = q * 32 + lane;
    uint32_t acc = 0, acc_phase = 0;
    for (int tile = pair; tile < num_tiles; tile += npairs) {
      const int mt = tile % m_tiles, nt = tile / m_tiles;
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * GEMM2_BN;
      const int m = mt * 256 + static_cast<int>(rank) * 128 + row_in_tile;
      gemm_epilogue_rows<GEMM2_BN, EPI>(g, taddr, m, m < g.M, nt);
      tc_fence_before();
      mbar_arrive_remote(map_to_cta(smem_u32(&tempty_bar[acc]), 0));
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  }

  tc_fence_before();
  cluster_sync_2sm();  // nobody leaves (or frees TMEM) while the peer may still address this CTA
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, 512);
  }
}

}  // namespace cvb
