// Skinny ("weight-streaming") bf16 GEMM for sm_100a:  C[M,N] = A[M,K] * W[N,K]^T with M <= 256.
//
// The expert's denoise step (M = 5 rows x 40 candidates = 200), the SigLIP tower (M = 256) and the verifier text
// tower (M = 64) are weight-streaming problems: every weight byte is used once and the work per layer is a few
// microseconds of HBM time.  The general persistent GEMM gives such shapes N/64 x 2 half-empty tiles that walk K
// serially.  Here the operands are SWAPPED so one UMMA covers all activation rows:
//
//     D[128 features x Mp rows] (TMEM, fp32) += W_tile[128 x 64] (UMMA "A") * A_tile[Mp x 64]^T (UMMA "B"),  Mp = M -> 16
//
//   * grid = (N / 128 weight tiles) x S, launched as clusters of S CTAs: every CTA of a cluster owns one K-slice of
//     the same weight tile, so >= 100 SMs stream disjoint weight slabs (TMA, 128B swizzle, mbarrier ring) at once;
//   * the S partial accumulators are reduced through DISTRIBUTED SHARED MEMORY: CTA j receives the fp32 slice
//     "activation rows [j*slice, (j+1)*slice)" of every peer (st.shared::cluster), sums them in rank order (no atomics:
//     deterministic, graph replays are bit-identical) and applies the fused epilogue with the reference's bf16
//     rounding points;
//   * no HBM round trip for partial sums, one launch per linear layer.
//
// Replaces (reference): the same nn.Linear calls as gemm_tcgen05.cuh (paligemma_with_expert.py:273-276, 327-341) in
// the regime of PI0FlowMatching.denoise_step (modeling_pi0.py:717-752) and embed_image (:229-230).
#pragma once
#include <cuda.h>

#include "gemm_tcgen05.cuh"
#include "ptx.cuh"

namespace cvb {

// extra epilogue kind (skinny only): GeGLU on weights packed [64 gate | 64 up] per 128-row tile
constexpr int EPI_GEGLU64 = 5;

struct SkinnyArgs {
  void* C;
  long ldc;
  const void* bias;
  int bias_is_f32;
  const void* resid;
  int resid_is_f32;
  long ldr;
  int M, Mp, N, K;
  int n_out;    // EPI_GEGLU64: intermediate size
  int S;        // K-splits = cluster size
  int slice;    // activation rows finalised per CTA (multiple of 4, S * slice >= Mp)
  int kbs;      // 64-wide k-blocks per split
  int stages;   // smem ring depth
  int tmem_cols;
  unsigned long long* ts;  // optional per-CTA phase timestamps [grid][8] (diagnostics), or nullptr
};

constexpr int SK_THREADS = 320;  // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue (two per TMEM lane quarter)
constexpr int SK_W_BYTES = 128 * 64 * 2;

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
// barrier without memory ordering: "every CTA of the cluster got here" (used before remote writes start)
__device__ __forceinline__ void cluster_sync_relaxed() {
  asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.aligned;" ::: "memory");
}
__device__ __forceinline__ float4 lds_f4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_shared(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_cluster_f32(uint32_t addr, float v) {
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ void st_cluster_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void tmem_ld_x8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
// same as make_idesc but with run-time N
__device__ __forceinline__ uint32_t make_idesc_rt(int fmt, int m, int n) {
  return (1u << 4) | (static_cast<uint32_t>(fmt) << 7) | (static_cast<uint32_t>(fmt) << 10) |
         (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(m >> 4) << 24);
}

__device__ __forceinline__ unsigned long long gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define SK_TS(i)                                                                  \
  do {                                                                            \
    if (g.ts != nullptr && threadIdx.x == 64 + 128) g.ts[blockIdx.x * 8 + (i)] = gtimer(); \
  } while (0)

__device__ __forceinline__ void sts_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
// shared::cta -> shared::cluster bulk copy (copy engine, asynchronous); completion = complete_tx on the DESTINATION's mbarrier
__device__ __forceinline__ void bulk_copy_to_peer(uint32_t dst_cluster_addr, uint32_t src_cta_addr, uint32_t bytes,
                                                  uint32_t mbar_cluster_addr) {
  asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   dst_cluster_addr),
               "r"(src_cta_addr), "r"(bytes), "r"(mbar_cluster_addr)
               : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

// Shared-memory map (after 1024-byte alignment):
//   [0, body)            TMA ring (stages x (16 KB weights + Mp x 128 B activations)); once this CTA's MMAs have retired
//                        the same bytes hold `stage_out`: its fp32 partial tile as [Mp/4 column quads][128 features][4]
//   [body, body + recv)  partial slices received from the S-1 peers: [slot][slice/4][128][4]
//   barriers
template <int EPI>
__global__ void __launch_bounds__(SK_THREADS, 1)
gemm_skinny_tcgen05(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmA,
                    const SkinnyArgs g) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  const uint32_t a_bytes = static_cast<uint32_t>(g.Mp) * 128u;
  const uint32_t stage_bytes = SK_W_BYTES + a_bytes;
  const uint32_t ring_bytes = static_cast<uint32_t>(g.stages) * stage_bytes;
  const uint32_t body = (max(ring_bytes, static_cast<uint32_t>(g.Mp) * 512u) + 1023u) & ~1023u;
  const uint32_t recv_bytes = static_cast<uint32_t>(g.S - 1) * g.slice * 512u;
  uint8_t* recv = smem + body;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(recv + recv_bytes);
  uint64_t* empty_bar = full_bar + g.stages;
  uint64_t* tfull_bar = empty_bar + g.stages;
  uint64_t* recv_bar = tfull_bar + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(recv_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int S = g.S;
  const int rank = static_cast<int>(cluster_ctarank());
  const int tile = blockIdx.x / S;
  const int kb_total = (g.K + 63) / 64;
  const int kb0 = rank * g.kbs;
  const int nkb = max(0, min(kb0 + g.kbs, kb_total) - kb0);
  const int s_act = (kb_total + g.kbs - 1) / g.kbs;  // splits that own k-blocks (a prefix of the ranks)
  const int my_cols = max(0, min(g.slice, g.Mp - rank * g.slice));
  SK_TS(0);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmW);
    tma_prefetch_desc(&tmA);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < g.stages; ++s) {
        mbar_init(&full_bar[s], 1);
        mbar_init(&empty_bar[s], 1);
      }
      mbar_init(tfull_bar, 1);
      mbar_init(recv_bar, 1);
      fence_barrier_init();
      fence_proxy_async();
      // bytes this CTA will receive: its column slice from every OTHER split that owns k-blocks
      const int senders = s_act - (rank < s_act ? 1 : 0);
      mbar_arrive_expect_tx(recv_bar, static_cast<uint32_t>(senders) * my_cols * 512u);
    }
    __syncwarp();
    tmem_alloc(tmem_slot, g.tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // every CTA's recv_bar is initialised before any peer may complete_tx on it (waited on just before the sends)
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  SK_TS(1);

  pdl_launch();
  if (warp == 0) {
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      // weight tiles first (independent of the preceding kernel), activations after the dependency resolves
      const int pre = min(nkb, g.stages);
      for (int i = 0; i < pre; ++i) {
        mbar_arrive_expect_tx(&full_bar[i], stage_bytes);
        tma_load_2d_hint(smem + i * stage_bytes, &tmW, &full_bar[i], (kb0 + i) * 64, tile * 128, kEvictFirst);
      }
      pdl_wait();
      for (int i = 0; i < nkb; ++i) {
        uint8_t* sw = smem + stage * stage_bytes;
        if (i < pre) {
          tma_load_2d_hint(sw + SK_W_BYTES, &tmA, &full_bar[stage], (kb0 + i) * 64, 0, kEvictLast);
        } else {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full_bar[stage], stage_bytes);
          tma_load_2d_hint(sw, &tmW, &full_bar[stage], (kb0 + i) * 64, tile * 128, kEvictFirst);
          tma_load_2d_hint(sw + SK_W_BYTES, &tmA, &full_bar[stage], (kb0 + i) * 64, 0, kEvictLast);
        }
        if (++stage == static_cast<uint32_t>(g.stages)) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0 && nkb > 0) {
      const uint32_t idesc = make_idesc_rt(1, 128, g.Mp);
      uint32_t stage = 0, phase = 0;
      for (int i = 0; i < nkb; ++i) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t w_addr = smem_u32(smem + stage * stage_bytes);
        const uint64_t wdesc = make_desc_kmajor_sw128(w_addr);
        const uint64_t adesc = make_desc_kmajor_sw128(w_addr + SK_W_BYTES);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16(tmem_base, wdesc + 2 * k, adesc + 2 * k, idesc, (i | k) != 0 ? 1u : 0u);
        umma_commit(&empty_bar[stage]);
        if (++stage == static_cast<uint32_t>(g.stages)) {
          stage = 0;
          phase ^= 1;
        }
      }
      umma_commit(tfull_bar);
    }
    __syncwarp();
  } else {
    pdl_wait();  // residual reads / C writes must not overtake the preceding kernel
    if (nkb > 0) {
      mbar_wait(tfull_bar, 0);  // this CTA's MMAs have retired: the ring is dead and becomes stage_out
      tc_fence_after();
    }
    SK_TS(2);
  }
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  SK_TS(3);

  const int half = (warp - 2) >> 2;  // which of the two epilogue warps of this lane quarter
  const uint32_t so_base = smem_u32(smem);
  const uint32_t rv_base = smem_u32(recv);
  const int squads = g.slice >> 2;
  if (warp >= 2) {
    if (nkb > 0) {
      // ---- TMEM -> registers -> stage_out[quad][feature][4]
      const int q = warp & 3;
      const int L = q * 32 + lane;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
      for (int c0 = half * 32; c0 < g.Mp; c0 += 64) {
        uint32_t r[32];
        const bool wide = c0 + 32 <= g.Mp;  // Mp is a multiple of 16: the last chunk may be 16 columns
        if (wide) {
          tmem_ld_x32(taddr + c0, r);
        } else {
          uint32_t r16[16];
          tmem_ld_x16(taddr + c0, r16);
#pragma unroll
          for (int i = 0; i < 16; ++i) r[i] = r16[i];
        }
        tmem_wait_ld();
        const uint32_t dst = so_base + (static_cast<uint32_t>(c0 >> 2) * 128u + L) * 16u;
#pragma unroll
        for (int v = 0; v < 8; ++v)
          if (v < 4 || wide) sts_v4(dst + v * 2048u, r[4 * v], r[4 * v + 1], r[4 * v + 2], r[4 * v + 3]);
      }
      tc_fence_before();
      fence_proxy_async();  // generic-proxy smem writes -> visible to the copy engine
    }
    named_bar_sync(1, SK_THREADS - 64);
    if (nkb > 0 && warp == 2 && lane == 0) {
      // ---- ship every peer its column slice of my partial tile (copy engine; lands with complete_tx on the peer)
      const int slot = rank;  // my slot in peer j: sources ordered by rank, the peer itself skipped
      for (int j = 0; j < S; ++j) {
        const int cols = max(0, min(g.slice, g.Mp - j * g.slice));
        if (j == rank || cols == 0) continue;
        const uint32_t dst = mapa_shared(rv_base + static_cast<uint32_t>((slot > j ? slot - 1 : slot) * squads) * 2048u, j);
        bulk_copy_to_peer(dst, so_base + static_cast<uint32_t>(j * squads) * 2048u, static_cast<uint32_t>(cols) * 512u,
                          mapa_shared(smem_u32(recv_bar), j));
      }
    }
    SK_TS(4);
    mbar_wait(recv_bar, 0);
    SK_TS(5);

    // ---- final: sum the partial slices in rank order, fused epilogue, store
    const int e = (warp & 3) * 32 + lane;  // feature inside the tile, 0..127
    const int c_begin = rank * g.slice;
    const int c_end = min(min(c_begin + g.slice, g.Mp), g.M);
    // address of source s's copy of (my quad `quad`, feature f)
    auto src_addr = [&](int s, int quad, int f) -> uint32_t {
      if (s == rank) return so_base + static_cast<uint32_t>((rank * squads + quad) * 128 + f) * 16u;
      const int slot = s > rank ? s - 1 : s;
      return rv_base + static_cast<uint32_t>((slot * squads + quad) * 128 + f) * 16u;
    };
    if constexpr (EPI == EPI_GEGLU64) {
      // lanes 0..63 of the tile hold gate features, 64..127 the matching up features
      const int fl = e & 63;
      const int f = tile * 64 + fl;
      const int sub = (e >> 6) + 2 * half;  // 4 workers per feature
      if (f < g.n_out) {
        for (int m0 = c_begin + 4 * sub; m0 < c_end; m0 += 16) {
          const int quad = (m0 - c_begin) >> 2;
          float4 gt = make_float4(0.f, 0.f, 0.f, 0.f), up = gt;
          for (int s = 0; s < s_act; ++s) {
            const float4 a = lds_f4(src_addr(s, quad, fl)), b = lds_f4(src_addr(s, quad, fl + 64));
            gt.x += a.x, gt.y += a.y, gt.z += a.z, gt.w += a.w;
            up.x += b.x, up.y += b.y, up.z += b.z, up.w += b.w;
          }
          const float gv[4] = {gt.x, gt.y, gt.z, gt.w}, uv[4] = {up.x, up.y, up.z, up.w};
          __nv_bfloat16* cp = reinterpret_cast<__nv_bfloat16*>(g.C) + static_cast<long>(m0) * g.ldc + f;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            if (m0 + i < c_end) {
              const float act = bf16_round(gelu_tanh_f(bf16_round(gv[i])));
              cp[i * g.ldc] = __float2bfloat16_rn(act * bf16_round(uv[i]));
            }
          }
        }
      }
    } else {
      const int n = tile * 128 + e;
      if (n < g.N) {
        const float b = load_bias(g.bias, g.bias_is_f32, n);
        for (int m0 = c_begin + 4 * half; m0 < c_end; m0 += 8) {
          const int quad = (m0 - c_begin) >> 2;
          float rr[4] = {0.f, 0.f, 0.f, 0.f};
          if constexpr (EPI == EPI_RESID) {  // issue the residual loads before the smem reduction
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              if (m0 + i < c_end) {
                const long ri = static_cast<long>(m0 + i) * g.ldr + n;
                rr[i] = g.resid_is_f32 ? reinterpret_cast<const float*>(g.resid)[ri]
                                       : __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(g.resid)[ri]);
              }
            }
          }
          float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
          for (int s = 0; s < s_act; ++s) {
            const float4 a = lds_f4(src_addr(s, quad, e));
            acc.x += a.x, acc.y += a.y, acc.z += a.z, acc.w += a.w;
          }
          const float av[4] = {acc.x, acc.y, acc.z, acc.w};
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int m = m0 + i;
            if (m < c_end) {
              float x = av[i] + b;
              if constexpr (EPI == EPI_F32) {
                reinterpret_cast<float*>(g.C)[static_cast<long>(m) * g.ldc + n] = x;
              } else {
                if constexpr (EPI == EPI_GELU) x = gelu_tanh_f(bf16_round(x));
                if constexpr (EPI == EPI_RESID) x = bf16_round(x) + rr[i];
                reinterpret_cast<__nv_bfloat16*>(g.C)[static_cast<long>(m) * g.ldc + n] = __float2bfloat16_rn(x);
              }
            }
          }
        }
      }
    }
  }

  SK_TS(6);
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, g.tmem_cols);
  }
  // nobody may exit while a peer's copy engine could still be reading its stage_out: every CTA got here only after all
  // bytes addressed to it (ours included) have landed
  cluster_sync_relaxed();
}

// ------------------------------------------------------------------------------------------------------------------
// Split-K variant WITHOUT a cluster: CTA (tile, split) streams its K-slice of one 128-row weight tile and stores its
// fp32 partial accumulator to P[split][m][n] (coalesced 128-byte rows).  The consumer (rmsnorm_reduce_kernel,
// ops_misc.cu) sums the S partials in split order (deterministic), applies the residual and the norm that follows the
// linear layer anyway - so the reduction costs no extra launch and no SM-to-SM traffic (measured: the DSMEM
// reduce-scatter above takes 4-5 us of a 10-13 us launch at the expert's o_proj / down_proj shapes; an L2 round trip
// of the same partials takes ~1 us).
constexpr int EPI_PARTIAL = 6;

struct SplitKArgs {
  float* P;
  long ldp;           // row stride of one partial (elements)
  long split_stride;  // elements between partials
  int M, Mp, N, K;
  int S;       // K-splits (grid = n_tiles * S)
  int stages;  // smem ring depth
  int tmem_cols;
};

__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

template <int UNUSED>  // (template only so that the header may be included by several translation units)
__global__ void __launch_bounds__(SK_THREADS, 1)
gemm_splitk_partial_tcgen05(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmA,
                            const SplitKArgs g) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  const uint32_t a_bytes = static_cast<uint32_t>(g.Mp) * 128u;
  const uint32_t stage_bytes = SK_W_BYTES + a_bytes;
  // the ring is reused as the transposed fp32 output tile [Mp][128]: the barriers sit behind the larger of the two
  const uint32_t body = max(static_cast<uint32_t>(g.stages) * stage_bytes, static_cast<uint32_t>(g.Mp) * 512u);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + body);
  uint64_t* empty_bar = full_bar + g.stages;
  uint64_t* tfull_bar = empty_bar + g.stages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int tile = blockIdx.x / g.S;
  const int split = blockIdx.x % g.S;
  const int kb_total = (g.K + 63) / 64;
  // balanced k-block ranges: every split owns >= 1 block (host guarantees S <= kb_total)
  const int kb0 = static_cast<int>(static_cast<long>(split) * kb_total / g.S);
  const int kb1 = static_cast<int>(static_cast<long>(split + 1) * kb_total / g.S);
  const int nkb = kb1 - kb0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmW);
    tma_prefetch_desc(&tmA);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < g.stages; ++s) {
        mbar_init(&full_bar[s], 1);
        mbar_init(&empty_bar[s], 1);
      }
      mbar_init(tfull_bar, 1);
      fence_barrier_init();
      fence_proxy_async();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, g.tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  pdl_launch();
  if (warp == 0) {
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      // weight tiles first (independent of the preceding kernel), activations after the dependency resolves
      const int pre = min(nkb, g.stages);
      for (int i = 0; i < pre; ++i) {
        mbar_arrive_expect_tx(&full_bar[i], stage_bytes);
        tma_load_2d_hint(smem + i * stage_bytes, &tmW, &full_bar[i], (kb0 + i) * 64, tile * 128, kEvictFirst);
      }
      pdl_wait();
      for (int i = 0; i < nkb; ++i) {
        uint8_t* sw = smem + stage * stage_bytes;
        if (i < pre) {
          tma_load_2d_hint(sw + SK_W_BYTES, &tmA, &full_bar[stage], (kb0 + i) * 64, 0, kEvictLast);
        } else {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full_bar[stage], stage_bytes);
          tma_load_2d_hint(sw, &tmW, &full_bar[stage], (kb0 + i) * 64, tile * 128, kEvictFirst);
          tma_load_2d_hint(sw + SK_W_BYTES, &tmA, &full_bar[stage], (kb0 + i) * 64, 0, kEvictLast);
        }
        if (++stage == static_cast<uint32_t>(g.stages)) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc_rt(1, 128, g.Mp);
      uint32_t stage = 0, phase = 0;
      for (int i = 0; i < nkb; ++i) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t w_addr = smem_u32(smem + stage * stage_bytes);
        const uint64_t wdesc = make_desc_kmajor_sw128(w_addr);
        const uint64_t adesc = make_desc_kmajor_sw128(w_addr + SK_W_BYTES);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16(tmem_base, wdesc + 2 * k, adesc + 2 * k, idesc, (i | k) != 0 ? 1u : 0u);
        umma_commit(&empty_bar[stage]);
        if (++stage == static_cast<uint32_t>(g.stages)) {
          stage = 0;
          phase ^= 1;
        }
      }
      umma_commit(tfull_bar);
    }
    __syncwarp();
  } else {
    pdl_wait();  // P may still be read by the kernel that consumed the previous partials
    mbar_wait(tfull_bar, 0);
    tc_fence_after();
    // TMEM lane = feature, column = activation row.  The TMA ring is dead (every load consumed, every MMA retired), so
    // the tile is transposed through it: each warp parks its 32 features x 32 rows as stage[row][feature] (conflict-free:
    // a warp writes 32 consecutive floats), then the 8 epilogue warps store whole rows - one 512-byte float4 store per
    // row instead of four scattered 128-byte ones (SM -> L2 writes were the slow part of this kernel).
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    float* stage = reinterpret_cast<float*>(smem);
    for (int c0 = half * 32; c0 < g.Mp; c0 += 64) {
      uint32_t r[32];
      const bool wide = c0 + 32 <= g.Mp;  // Mp is a multiple of 16: the last chunk may be 16 columns
      if (wide) {
        tmem_ld_x32(taddr + c0, r);
      } else {
        uint32_t r16[16];
        tmem_ld_x16(taddr + c0, r16);
#pragma unroll
        for (int i = 0; i < 16; ++i) r[i] = r16[i];
      }
      tmem_wait_ld();
      float* sp = stage + static_cast<long>(c0) * 128 + q * 32 + lane;
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (i < 16 || wide) sp[i * 128] = __uint_as_float(r[i]);
    }
    named_bar_sync(1, SK_THREADS - 64);
    const int n0 = tile * 128 + lane * 4;
    float* pp = g.P + static_cast<long>(split) * g.split_stride + n0;
    if (n0 < g.N) {  // N is a multiple of 4 (checked on the host)
      for (int m = warp - 2; m < g.M; m += 8)
        *reinterpret_cast<float4*>(pp + static_cast<long>(m) * g.ldp) = *reinterpret_cast<const float4*>(stage + m * 128 + lane * 4);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, g.tmem_cols);
  }
}

}  // namespace cvb
