// Persistent "expert chain" kernel for the denoise loop (sm_100a): ONE launch walks a host-built list of phases -
// split-K weight-streaming GEMMs, the GeGLU gate/up GEMM and the residual + RMSNorm reductions of the action expert's
// layers - separated by device-wide barriers instead of kernel boundaries.
//
//   * one CTA per SM, all co-resident (cooperative launch); every CTA owns a fixed slice of every phase's work;
//   * warp 0 = WEIGHT producer: weights never depend on activations, so it runs ahead of the barriers and keeps an
//     8 x 16 KB TMA ring full with the tiles of the NEXT phases while the current phase drains (the 148 rings hold
//     19 MB - more than half a layer of expert weights - so the HBM stream does not stop at op boundaries);
//   * warp 1 = ACTIVATION producer: waits for the device-wide barrier of the previous phase, then TMA-loads the
//     phase's activation k-blocks (they come out of L2) into a second ring;
//   * warp 2 = MMA issuer (tcgen05.mma, fp32 accumulators double-buffered in TMEM);
//   * warps 3..18 = 16 worker warps: TMEM epilogues (fp32 split-K partials / GeGLU), the reduce + residual + RMSNorm
//     phases, the barrier arrival.
//
// Rounding ledger = the separate kernels it replaces (gemm_splitk_partial_tcgen05 + rmsnorm_reduce_kernel,
// gemm_bf16_tcgen05<EPI_GEGLU>): partial sums are added in split order (deterministic), h = bf16(bf16(sum) + resid),
// y = bf16(h * rsqrt(mean(h^2) + eps) * (1 + w)), act = bf16(bf16(gelu_tanh(bf16(g))) * bf16(u)).
//
// Replaces (reference): the per-layer body of PaliGemmaWithExpertModel.forward for the suffix tokens,
// paligemma_with_expert.py:258-349, inside PI0FlowMatching.denoise_step (modeling_pi0.py:717-752).
#pragma once
#include <cuda.h>

#include "expert_mega.h"
#include "gemm_skinny.cuh"
#include "ptx.cuh"

namespace cvb {

constexpr int MK_WORK_WARPS = 16;
constexpr int MK_WORKERS = MK_WORK_WARPS * 32;      // 512
constexpr int MK_THREADS = (3 + MK_WORK_WARPS) * 32;  // 608
constexpr int MK_NSW = 8;                           // weight ring stages (16 KB each)
constexpr int MK_NSA = 3;                           // activation ring stages (32 KB each)
constexpr uint32_t MK_W_BYTES = 128 * 64 * 2;
constexpr uint32_t MK_A_BYTES = 256 * 64 * 2;
constexpr uint32_t MK_RING_BYTES = MK_NSW * MK_W_BYTES + MK_NSA * MK_A_BYTES;  // 229376
constexpr uint32_t MK_MISC_BYTES = 1024;
constexpr uint32_t MK_SMEM = MK_RING_BYTES + MK_MISC_BYTES + 1024;  // + alignment slack

__device__ __forceinline__ unsigned ld_volatile_shared(const unsigned* p) {
  unsigned v;
  asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory");
  return v;
}
__device__ __forceinline__ void st_volatile_shared(unsigned* p, unsigned v) {
  asm volatile("st.volatile.shared.u32 [%0], %1;" ::"r"(smem_u32(p)), "r"(v) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

// every spin loop of the kernel is bounded: a protocol bug must end in a trapped launch, never in a hung GPU
struct MegaWatch {
  unsigned long long t0;
  long long limit;
  unsigned* err;
  __device__ __forceinline__ void start() { t0 = gtimer(); }
  __device__ __forceinline__ void check(unsigned code) {
    if (static_cast<long long>(gtimer() - t0) > limit) {
      if (err != nullptr) atomicExch(err, code);
      __threadfence_system();
      __trap();
    }
  }
};

__device__ __forceinline__ void mk_mbar_wait(uint64_t* bar, uint32_t parity, MegaWatch& w, unsigned code) {
  if (mbar_try_wait(bar, parity)) return;
  w.start();
  unsigned n = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++n & 0x3FFu) == 0) w.check(code);
  }
}

// unit -> (feature tile, split / row tile, k-block range); identical in every role
struct MegaUnit {
  int ftile, sub, kb0, kb1;
};
__device__ __forceinline__ MegaUnit mk_unit(const MegaPhase& ph, int u) {
  MegaUnit r;
  r.ftile = u / ph.splits;
  r.sub = u % ph.splits;
  if (ph.kind == MK_GEMM_PARTIAL) {
    r.kb0 = static_cast<int>(static_cast<long>(r.sub) * ph.kb_total / ph.splits);
    r.kb1 = static_cast<int>(static_cast<long>(r.sub + 1) * ph.kb_total / ph.splits);
  } else {
    r.kb0 = 0;
    r.kb1 = ph.kb_total;
  }
  return r;
}
__device__ __forceinline__ int mk_first_unit(const MegaPhase& ph) {
  const int g = static_cast<int>(gridDim.x);
  return (static_cast<int>(blockIdx.x) - (ph.rot % g) + g) % g;
}

__global__ void __launch_bounds__(MK_THREADS, 1)
expert_mega_kernel(const __grid_constant__ MegaMaps maps, const MegaArgs g) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* ringW = smem;
  uint8_t* ringA = smem + MK_NSW * MK_W_BYTES;
  uint8_t* misc = smem + MK_RING_BYTES;
  uint64_t* wfull = reinterpret_cast<uint64_t*>(misc);
  uint64_t* wempty = wfull + MK_NSW;
  uint64_t* afull = wempty + MK_NSW;
  uint64_t* aempty = afull + MK_NSA;
  uint64_t* tfull = aempty + MK_NSA;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  unsigned* go_count = tmem_slot + 1;            // number of device-wide barriers this CTA has seen complete
  float* red = reinterpret_cast<float*>(misc + 512);  // [2 row slots][8 warps]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int G = static_cast<int>(gridDim.x);
  MegaWatch watch;
  watch.limit = g.spin_limit_ns;
  watch.err = g.err;

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < 8; ++i) tma_prefetch_desc(&maps.m[i]);
  }
  if (warp == 2) {
    if (lane == 0) {
      for (int s = 0; s < MK_NSW; ++s) {
        mbar_init(&wfull[s], 1);
        mbar_init(&wempty[s], 1);
      }
      for (int s = 0; s < MK_NSA; ++s) {
        mbar_init(&afull[s], 1);
        mbar_init(&aempty[s], 1);
      }
      for (int a = 0; a < 2; ++a) {
        mbar_init(&tfull[a], 1);
        mbar_init(&tempty[a], MK_WORK_WARPS);
      }
      *go_count = 0u;
      fence_barrier_init();
      fence_proxy_async();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------------------ weight producer (runs ahead)
    if (lane == 0) {
      uint32_t stage = 0, par = 0;
      for (int p = 0; p < g.n_phases; ++p) {
        const MegaPhase ph = g.prog[p];  // by value: fields stay in registers (a reference would be re-read after every store)
        if (ph.kind == MK_NORM) continue;
        const CUtensorMap* tm = &maps.m[ph.wmap];
        const int units = ph.ftiles * ph.splits;
        for (int u = mk_first_unit(ph); u < units; u += G) {
          const MegaUnit un = mk_unit(ph, u);
          for (int kb = un.kb0; kb < un.kb1; ++kb) {
            mk_mbar_wait(&wempty[stage], par ^ 1, watch, 0x100u + p);
            mbar_arrive_expect_tx(&wfull[stage], MK_W_BYTES);
            tma_load_2d_hint(ringW + stage * MK_W_BYTES, tm, &wfull[stage], kb * 64, ph.w_row0 + un.ftile * 128, kEvictFirst);
            if (++stage == MK_NSW) stage = 0, par ^= 1;
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------------------------ activation producer
    if (lane == 0) {
      pdl_wait();  // activations of phase 0 come from the preceding kernel
      uint32_t stage = 0, par = 0;
      for (int p = 0; p < g.n_phases; ++p) {
        const MegaPhase ph = g.prog[p];  // by value: fields stay in registers (a reference would be re-read after every store)
        if (ph.kind == MK_NORM) continue;
        if (p > 0) {
          if (ld_volatile_shared(go_count) < static_cast<unsigned>(p)) {
            watch.start();
            unsigned n = 0;
            while (ld_volatile_shared(go_count) < static_cast<unsigned>(p)) {
              if ((++n & 0xFFFu) == 0) watch.check(0x200u + p);
            }
          }
          __threadfence();
          fence_proxy_async_all();
        }
        const CUtensorMap* tm = &maps.m[ph.amap];
        const int units = ph.ftiles * ph.splits;
        const uint32_t bytes = ph.kind == MK_GEMM_PARTIAL ? static_cast<uint32_t>(ph.rows_pad) * 128u : MK_W_BYTES;
        for (int u = mk_first_unit(ph); u < units; u += G) {
          const MegaUnit un = mk_unit(ph, u);
          const int row0 = ph.kind == MK_GEMM_PARTIAL ? 0 : un.sub * 128;
          for (int kb = un.kb0; kb < un.kb1; ++kb) {
            mk_mbar_wait(&aempty[stage], par ^ 1, watch, 0x300u + p);
            mbar_arrive_expect_tx(&afull[stage], bytes);
            tma_load_2d_hint(ringA + stage * MK_A_BYTES, tm, &afull[stage], kb * 64, row0, kEvictLast);
            if (++stage == MK_NSA) stage = 0, par ^= 1;
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 2) {
    // ------------------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      uint32_t ws = 0, wpar = 0, as = 0, apar = 0, acc = 0, accpar = 0;
      for (int p = 0; p < g.n_phases; ++p) {
        const MegaPhase ph = g.prog[p];  // by value: fields stay in registers (a reference would be re-read after every store)
        if (ph.kind == MK_NORM) continue;
        const bool swapped = ph.kind == MK_GEMM_PARTIAL;
        const uint32_t idesc = swapped ? make_idesc_rt(1, 128, ph.rows_pad) : make_idesc_rt(1, 128, 128);
        const int units = ph.ftiles * ph.splits;
        for (int u = mk_first_unit(ph); u < units; u += G) {
          const MegaUnit un = mk_unit(ph, u);
          mk_mbar_wait(&tempty[acc], accpar ^ 1, watch, 0x400u + p);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + acc * 256u;
          for (int kb = un.kb0; kb < un.kb1; ++kb) {
            mk_mbar_wait(&wfull[ws], wpar, watch, 0x500u + p);
            mk_mbar_wait(&afull[as], apar, watch, 0x600u + p);
            tc_fence_after();
            const uint64_t wdesc = make_desc_kmajor_sw128(smem_u32(ringW + ws * MK_W_BYTES));
            const uint64_t adesc = make_desc_kmajor_sw128(smem_u32(ringA + as * MK_A_BYTES));
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint32_t accum = (kb != un.kb0 || k != 0) ? 1u : 0u;
              if (swapped)
                umma_bf16(d_tmem, wdesc + 2 * k, adesc + 2 * k, idesc, accum);  // D[feature][row]
              else
                umma_bf16(d_tmem, adesc + 2 * k, wdesc + 2 * k, idesc, accum);  // D[row][packed feature]
            }
            umma_commit(&wempty[ws]);
            umma_commit(&aempty[as]);
            if (++ws == MK_NSW) ws = 0, wpar ^= 1;
            if (++as == MK_NSA) as = 0, apar ^= 1;
          }
          umma_commit(&tfull[acc]);
          acc ^= 1;
          if (acc == 0) accpar ^= 1;
        }
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------------------ workers
    const int wtid = threadIdx.x - 96;       // 0..511
    const int q = warp & 3;                  // TMEM lane quarter this warp may access
    const int cg = (warp - 3) >> 2;          // column group 0..3
    uint32_t acc = 0, accpar = 0;
    pdl_wait();  // no global write / read of this kernel may overtake the preceding kernel
    unsigned long long* ts = g.ts != nullptr && wtid == 0 ? g.ts + static_cast<long>(blockIdx.x) * 64 : nullptr;
    if (ts != nullptr) ts[0] = gtimer();
    for (int p = 0; p < g.n_phases; ++p) {
      const MegaPhase ph = g.prog[p];  // by value: fields stay in registers (a reference would be re-read after every store)
      if (ph.kind == MK_NORM) {
        // ---- rows of this CTA, two at a time (256 threads per row)
        const int slot = wtid >> 8, t = wtid & 255;
        const int first = mk_first_unit(ph);
        for (int it = 0;; ++it) {
          const int row = first + (2 * it + slot) * G;
          if (first + 2 * it * G >= ph.rows) break;   // uniform over both slots
          const bool on = row < ph.rows;
          float hv[4][4];
          float ss = 0.f;
          if (on) {
#pragma unroll
            for (int v = 0; v < 4; ++v) {
              const int i = (t + 256 * v) * 4;
              if (i < ph.width) {
                float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
                const float* pr = ph.nP + static_cast<long>(row) * ph.n_ldp + i;
                int s = 0;
                for (; s + 4 <= ph.nS; s += 4) {  // four loads in flight, summed in split order
                  const float4 a0 = __ldcg(reinterpret_cast<const float4*>(pr + (s + 0) * ph.n_split_stride));
                  const float4 a1 = __ldcg(reinterpret_cast<const float4*>(pr + (s + 1) * ph.n_split_stride));
                  const float4 a2 = __ldcg(reinterpret_cast<const float4*>(pr + (s + 2) * ph.n_split_stride));
                  const float4 a3 = __ldcg(reinterpret_cast<const float4*>(pr + (s + 3) * ph.n_split_stride));
                  a.x = (((a.x + a0.x) + a1.x) + a2.x) + a3.x;
                  a.y = (((a.y + a0.y) + a1.y) + a2.y) + a3.y;
                  a.z = (((a.z + a0.z) + a1.z) + a2.z) + a3.z;
                  a.w = (((a.w + a0.w) + a1.w) + a2.w) + a3.w;
                }
                for (; s < ph.nS; ++s) {
                  const float4 b = __ldcg(reinterpret_cast<const float4*>(pr + s * ph.n_split_stride));
                  a.x += b.x, a.y += b.y, a.z += b.z, a.w += b.w;
                }
                float rr[4];
                if (ph.resid_f32) {
                  const float4 r4 = __ldcg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(ph.resid) +
                                                                           static_cast<long>(row) * ph.ldr + i));
                  rr[0] = r4.x, rr[1] = r4.y, rr[2] = r4.z, rr[3] = r4.w;
                } else {
                  const uint2 r2 = __ldcg(reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(ph.resid) +
                                                                         static_cast<long>(row) * ph.ldr + i));
                  const float2 f0 = unpack_bf16x2(r2.x), f1 = unpack_bf16x2(r2.y);
                  rr[0] = f0.x, rr[1] = f0.y, rr[2] = f1.x, rr[3] = f1.y;
                }
                if (ph.nS > 0) {
                  hv[v][0] = bf16_round(bf16_round(a.x) + rr[0]);
                  hv[v][1] = bf16_round(bf16_round(a.y) + rr[1]);
                  hv[v][2] = bf16_round(bf16_round(a.z) + rr[2]);
                  hv[v][3] = bf16_round(bf16_round(a.w) + rr[3]);
                } else {
                  hv[v][0] = rr[0], hv[v][1] = rr[1], hv[v][2] = rr[2], hv[v][3] = rr[3];
                }
                if (ph.h_out != nullptr)
                  *reinterpret_cast<uint2*>(ph.h_out + static_cast<long>(row) * ph.ldh + i) =
                      make_uint2(pack_bf16x2(hv[v][0], hv[v][1]), pack_bf16x2(hv[v][2], hv[v][3]));
                ss += hv[v][0] * hv[v][0] + hv[v][1] * hv[v][1] + hv[v][2] * hv[v][2] + hv[v][3] * hv[v][3];
              }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
            if (lane == 0) red[slot * 8 + (t >> 5)] = ss;
          }
          named_bar_sync(3 + slot, 256);
          if (on) {
            float tot = 0.f;
#pragma unroll
            for (int k = 0; k < 8; ++k) tot += red[slot * 8 + k];
            const float r = 1.0f / sqrtf(tot / static_cast<float>(ph.width) + ph.eps);
#pragma unroll
            for (int v = 0; v < 4; ++v) {
              const int i = (t + 256 * v) * 4;
              if (i < ph.width) {
                float wv[4];
                if (ph.nw_f32) {
                  const float4 w4 = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(ph.nw) + i);
                  wv[0] = w4.x, wv[1] = w4.y, wv[2] = w4.z, wv[3] = w4.w;
                } else {
                  const uint2 w2 = *reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(ph.nw) + i);
                  const float2 f0 = unpack_bf16x2(w2.x), f1 = unpack_bf16x2(w2.y);
                  wv[0] = f0.x, wv[1] = f0.y, wv[2] = f1.x, wv[3] = f1.y;
                }
                *reinterpret_cast<uint2*>(ph.y + static_cast<long>(row) * ph.ldy + i) =
                    make_uint2(pack_bf16x2((hv[v][0] * r) * (1.0f + wv[0]), (hv[v][1] * r) * (1.0f + wv[1])),
                               pack_bf16x2((hv[v][2] * r) * (1.0f + wv[2]), (hv[v][3] * r) * (1.0f + wv[3])));
              }
            }
          }
          named_bar_sync(3 + slot, 256);  // red[] is reused by the next pair of rows
        }
      } else {
        const int units = ph.ftiles * ph.splits;
        for (int u = mk_first_unit(ph); u < units; u += G) {
          const MegaUnit un = mk_unit(ph, u);
          mk_mbar_wait(&tfull[acc], accpar, watch, 0x700u + p);
          tc_fence_after();
          if (ts != nullptr && p < 21) ts[1 + 3 * p] = gtimer();
          const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * 256u;
          if (ph.kind == MK_GEMM_PARTIAL) {
            // TMEM lane = feature, column = activation row: P[split][m][n], 32 lanes = 128 contiguous bytes per row
            const int n = un.ftile * 128 + q * 32 + lane;
            float* pp = ph.P + static_cast<long>(un.sub) * ph.split_stride + n;
            for (int c0 = cg * 16; c0 < ph.rows_pad; c0 += 64) {
              uint32_t r[16];
              tmem_ld_x16(taddr + c0, r);
              tmem_wait_ld();
              if (n < ph.n_feat) {
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                  const int m = c0 + i;
                  if (m < ph.rows) pp[static_cast<long>(m) * ph.ldp] = __uint_as_float(r[i]);
                }
              }
            }
          } else {
            // TMEM lane = activation row, columns = [64 gate | 64 up] of feature block `ftile`
            const int m = un.sub * 128 + q * 32 + lane;
            uint32_t rg[16], ru[16];
            tmem_ld_x16(taddr + cg * 16, rg);
            tmem_ld_x16(taddr + 64 + cg * 16, ru);
            tmem_wait_ld();
            const int f0 = un.ftile * 64 + cg * 16;
            if (m < ph.rows && f0 < ph.n_feat) {
              uint32_t o[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                float r2[2];
#pragma unroll
                for (int h2 = 0; h2 < 2; ++h2) {
                  const float gt = bf16_round(__uint_as_float(rg[2 * e + h2]));
                  const float up = bf16_round(__uint_as_float(ru[2 * e + h2]));
                  r2[h2] = bf16_round(gelu_tanh_f(gt)) * up;
                }
                o[e] = pack_bf16x2(r2[0], r2[1]);
              }
              __nv_bfloat16* cp = ph.C + static_cast<long>(m) * ph.ldc + f0;
              *reinterpret_cast<uint4*>(cp) = make_uint4(o[0], o[1], o[2], o[3]);
              *reinterpret_cast<uint4*>(cp + 8) = make_uint4(o[4], o[5], o[6], o[7]);
            }
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tempty[acc]);
          acc ^= 1;
          if (acc == 0) accpar ^= 1;
        }
      }

      // ---- device-wide barrier after every phase but the last
      if (ts != nullptr && p < 21) ts[2 + 3 * p] = gtimer();
      if (p + 1 < g.n_phases) {
        named_bar_sync(1, MK_WORKERS);  // every worker's global writes of this phase are issued
        if (wtid == 0) {
          const unsigned nb = blockIdx.x == 0 ? 0x80000000u - static_cast<unsigned>(G - 1) : 1u;
          __threadfence();
          const unsigned old = atomicAdd(g.bar, nb);
          if (((old ^ ld_acquire_gpu(g.bar)) & 0x80000000u) == 0u) {
            watch.start();
            unsigned n = 0;
            while (((old ^ ld_acquire_gpu(g.bar)) & 0x80000000u) == 0u) {
              if ((++n & 0xFFu) == 0) watch.check(0x800u + p);
            }
          }
          __threadfence();
          st_volatile_shared(go_count, static_cast<unsigned>(p + 1));
          if (ts != nullptr && p < 21) ts[3 + 3 * p] = gtimer();
        }
        if (g.prog[p + 1].kind == MK_NORM) named_bar_sync(2, MK_WORKERS);  // the next phase reads other CTAs' results
      }
    }
  }

  pdl_launch();
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace cvb
