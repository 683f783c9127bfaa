// Persistent, warp-specialised bf16 GEMM for sm_100a:  C[M,N] = A[M,K] * W[N,K]^T  (+ fused epilogue)
//
//   * operands arrive by TMA (cp.async.bulk.tensor, 128-byte swizzle) into a STAGES-deep smem ring,
//   * one elected thread issues tcgen05.mma (UMMA 128 x BN x 16, cta_group::1) into TMEM,
//   * the fp32 accumulator is double-buffered in TMEM (2 x BN columns) so the epilogue of tile i
//     overlaps the main loop of tile i+1,
//   * four epilogue warps read TMEM with tcgen05.ld and apply the fused epilogue with exactly the
//     bf16 rounding points of the reference graph.
//
// Both operands are K-major ([rows, K] row-major), which is exactly how activations and PyTorch
// nn.Linear weights ([out, in]) lie in HBM - no transposes anywhere.
//
// Replaces (reference): every nn.Linear on the pi0 path, e.g. q/k/v/o_proj and the Gemma MLP in
// lerobot_custom/lerobot/common/policies/pi0/paligemma_with_expert.py:273-276,327-336 and the
// SigLIP / ViT-L trunk linears reached through embed_image (:229-230).
#pragma once
#include <cuda.h>

#include "ptx.cuh"

namespace cvb {

enum EpiKind : int {
  EPI_STORE = 0,  // C = bf16(acc + bias)
  EPI_GELU = 1,   // C = bf16(gelu_tanh(bf16(acc + bias)))
  EPI_RESID = 2,  // C = bf16(bf16(acc + bias) + R)            (R may alias C)
  EPI_GEGLU = 3,  // W rows packed [128 gate | 128 up] per 256-row block:
                  // C[:, f] = bf16(bf16(gelu_tanh(bf16(g))) * bf16(u))
  EPI_F32 = 4,    // C(fp32) = acc + bias
};

struct GemmArgs {
  void* C;
  long ldc;
  const void* bias;  // [N] or nullptr
  int bias_is_f32;
  const void* resid;  // bf16 (or fp32 when resid_is_f32) [M, ldr]
  int resid_is_f32;
  long ldr;
  int M, N, K;       // N = number of W rows (for EPI_GEGLU: packed rows, 256 per 128 features)
  int n_out;         // number of valid output columns (EPI_GEGLU: the intermediate size)
  const int* m_dev;  // optional: device-side row count (rows >= *m_dev are skipped), graph-safe varlen
};

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;
constexpr int GEMM_THREADS = 192;

template <int BN, int STAGES>
struct GemmSmem {
  static constexpr uint32_t A_BYTES = GEMM_BM * GEMM_BK * 2;
  static constexpr uint32_t B_BYTES = BN * GEMM_BK * 2;
  static constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr uint32_t BAR_OFFSET = STAGES * STAGE_BYTES;
  static constexpr uint32_t TOTAL = BAR_OFFSET + 256 + 1024;  // + barriers + alignment slack
};

__device__ __forceinline__ float load_bias(const void* bias, int is_f32, int n) {
  if (bias == nullptr) return 0.f;
  return is_f32 ? reinterpret_cast<const float*>(bias)[n]
                : __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(bias)[n]);
}

// Fused epilogue of one accumulator row: this thread owns row m of the tile (TMEM lane), BN fp32 columns starting at
// taddr; nt = N-tile index.  Shared by the 1-CTA and the 2-CTA (cta_group::2) kernels.
template <int BN, int EPI>
__device__ __forceinline__ void gemm_epilogue_rows(const GemmArgs& g, uint32_t taddr, int m, bool row_ok, int nt) {
  if constexpr (EPI == EPI_GEGLU) {
    // packing: [BN/2 gate rows | BN/2 up rows] per BN-row block (BN = 256: prefix; BN = 128: expert, many small tiles)
    static_assert(EPI != EPI_GEGLU || BN == 256 || BN == 128, "GEGLU tiles hold BN/2 gate + BN/2 up rows");
    __nv_bfloat16* crow = reinterpret_cast<__nv_bfloat16*>(g.C) + static_cast<long>(m) * g.ldc;
#pragma unroll 1
    for (int c = 0; c < BN / 2; c += 32) {
      uint32_t rg[32], ru[32];
      tmem_ld_x32(taddr + c, rg);
      tmem_ld_x32(taddr + BN / 2 + c, ru);
      tmem_wait_ld();
      const int f0 = nt * (BN / 2) + c;
      if (row_ok) {
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          const int f = f0 + v * 8;
          if (f < g.n_out) {
            uint32_t o[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              float r2[2];
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                const int i = v * 8 + e * 2 + h;
                const float gt = bf16_round(__uint_as_float(rg[i]));
                const float up = bf16_round(__uint_as_float(ru[i]));
                const float act = bf16_round(gelu_tanh_f(gt));
                r2[h] = act * up;
              }
              o[e] = pack_bf16x2(r2[0], r2[1]);
            }
            *reinterpret_cast<uint4*>(crow + f) = make_uint4(o[0], o[1], o[2], o[3]);
          }
        }
      }
    }
  } else {
#pragma unroll 1
    for (int c = 0; c < BN; c += 32) {
      uint32_t r[32];
      tmem_ld_x32(taddr + c, r);
      tmem_wait_ld();
      const int n0 = nt * BN + c;
      if (row_ok) {
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          const int n = n0 + v * 8;
          if (n < g.N) {
            float x[8];
#pragma unroll
            for (int e = 0; e < 8; ++e)
              x[e] = __uint_as_float(r[v * 8 + e]) + load_bias(g.bias, g.bias_is_f32, n + e);
            if constexpr (EPI == EPI_F32) {
              float* crow = reinterpret_cast<float*>(g.C) + static_cast<long>(m) * g.ldc + n;
              *reinterpret_cast<float4*>(crow) = make_float4(x[0], x[1], x[2], x[3]);
              *reinterpret_cast<float4*>(crow + 4) = make_float4(x[4], x[5], x[6], x[7]);
            } else {
              if constexpr (EPI == EPI_GELU) {
#pragma unroll
                for (int e = 0; e < 8; ++e) x[e] = gelu_tanh_f(bf16_round(x[e]));
              }
              if constexpr (EPI == EPI_RESID) {
                if (g.resid_is_f32) {
                  const float* rp = reinterpret_cast<const float*>(g.resid) +
                                    static_cast<long>(m) * g.ldr + n;
                  const float4 r0 = *reinterpret_cast<const float4*>(rp);
                  const float4 r1 = *reinterpret_cast<const float4*>(rp + 4);
                  const float rr[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
                  for (int e = 0; e < 8; ++e) x[e] = bf16_round(x[e]) + rr[e];
                } else {
                  const uint4 rv = *reinterpret_cast<const uint4*>(
                      reinterpret_cast<const __nv_bfloat16*>(g.resid) +
                      static_cast<long>(m) * g.ldr + n);
                  const uint32_t rr[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
                  for (int e = 0; e < 4; ++e) {
                    const float2 f2 = unpack_bf16x2(rr[e]);
                    x[2 * e] = bf16_round(x[2 * e]) + f2.x;
                    x[2 * e + 1] = bf16_round(x[2 * e + 1]) + f2.y;
                  }
                }
              }
              __nv_bfloat16* crow =
                  reinterpret_cast<__nv_bfloat16*>(g.C) + static_cast<long>(m) * g.ldc + n;
              *reinterpret_cast<uint4*>(crow) =
                  make_uint4(pack_bf16x2(x[0], x[1]), pack_bf16x2(x[2], x[3]),
                             pack_bf16x2(x[4], x[5]), pack_bf16x2(x[6], x[7]));
            }
          }
        }
      }
    }
  }
}

template <int BN, int STAGES, int EPI>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_tcgen05(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const GemmArgs g) {
  using S = GemmSmem<BN, STAGES>;
  constexpr uint32_t TMEM_COLS = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128
                                 : (2 * BN <= 256) ? 256 : 512;
  static_assert(2 * BN <= 512, "accumulator double buffer must fit TMEM");
  constexpr uint32_t IDESC = make_idesc(/*bf16*/ 1, GEMM_BM, BN);

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S::BAR_OFFSET);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  int M = g.M;
  if (g.m_dev != nullptr) {
    pdl_wait();  // the device-side row count may come from the preceding kernel
    M = min(M, *g.m_dev);
  }
  const int m_tiles = (M + GEMM_BM - 1) / GEMM_BM;
  const int n_tiles = (g.N + BN - 1) / BN;
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
        mbar_init(&tempty_bar[a], 128);
      }
      fence_barrier_init();
      fence_proxy_async();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch();  // the next kernel may start its prologue (and its own weight prefetch) now

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      // Weights (operand B) never depend on the preceding kernel: put the first ring-full of weight tiles in flight
      // BEFORE waiting for it (programmatic dependent launch), then add the activation tiles.
      int pre = 0;
      if (static_cast<int>(blockIdx.x) < num_tiles) {
        const int nt0 = blockIdx.x / m_tiles;
        pre = min(STAGES, k_blocks);
        for (int kb = 0; kb < pre; ++kb) {
          mbar_arrive_expect_tx(&full_bar[kb], S::STAGE_BYTES);
          tma_load_2d(smem + kb * S::STAGE_BYTES + S::A_BYTES, &tmB, &full_bar[kb], kb * GEMM_BK, nt0 * BN);
        }
      }
      pdl_wait();
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int mt = tile % m_tiles, nt = tile / m_tiles;
        for (int kb = 0; kb < k_blocks; ++kb) {
          uint8_t* sa = smem + stage * S::STAGE_BYTES;
          uint8_t* sb = sa + S::A_BYTES;
          if (tile == static_cast<int>(blockIdx.x) && kb < pre) {
            tma_load_2d(sa, &tmA, &full_bar[stage], kb * GEMM_BK, mt * GEMM_BM);
            if (++stage == STAGES) {
              stage = 0;
              phase ^= 1;
            }
            continue;
          }
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full_bar[stage], S::STAGE_BYTES);
          tma_load_2d(sa, &tmA, &full_bar[stage], kb * GEMM_BK, mt * GEMM_BM);
          tma_load_2d(sb, &tmB, &full_bar[stage], kb * GEMM_BK, nt * BN);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (one thread)
    if (lane == 0) {
      uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + stage * S::STAGE_BYTES);
          const uint64_t adesc = make_desc_kmajor_sw128(a_addr);
          const uint64_t bdesc = make_desc_kmajor_sw128(a_addr + S::A_BYTES);
#pragma unroll
          for (int k = 0; k < GEMM_BK / 16; ++k) {
            // +32 bytes per UMMA_K step inside the 128-byte swizzle row (address field is >>4)
            umma_bf16(d_tmem, adesc + 2 * k, bdesc + 2 * k, IDESC, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);  // frees this smem slot when the MMAs retire
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(&tfull_bar[acc]);  // accumulator complete -> epilogue
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue (4 warps, 128 rows)
    pdl_wait();  // residual reads / C writes must not overtake the preceding kernel
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    const int row_in_tile = q * 32 + lane;
    uint32_t acc = 0, acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int mt = tile % m_tiles, nt = tile / m_tiles;
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN;
      const int m = mt * GEMM_BM + row_in_tile;
      const bool row_ok = m < M;

      gemm_epilogue_rows<BN, EPI>(g, taddr, m, row_ok, nt);
      tc_fence_before();
      mbar_arrive(&tempty_bar[acc]);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

}  // namespace cvb
