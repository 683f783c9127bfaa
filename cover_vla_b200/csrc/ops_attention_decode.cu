// Denoise-step attention (PI0FlowMatching.denoise_step, modeling_pi0.py:717-752 -> eager_attention_forward,
// paligemma_with_expert.py:376-434): every candidate's 8 heads x 5 suffix tokens = 40 query rows attend the prefix KV
// cache of the candidate's rephrase (<= 328 keys) plus the candidate's own 5 suffix keys.
//
// One CTA per candidate walks 6 K tiles and 6 V tiles serially and is pure latency.  Here a CLUSTER of C = ceil(keys/64)
// CTAs serves one candidate, one 64-key tile each:
//   1. Q (all heads of the KV group folded into rows) and the tile's K / V are staged once (cp.async); RoPE
//      (apply_rope, paligemma_with_expert.py:34-57) is applied to Q and to the suffix keys WHILE staging, from a
//      (cos, sin) table computed once per sample - the separate RoPE launch of the denoise layer disappears;
//   2. S = Q K^T on mma.sync (fp32), scale, mask (suffix mask [1,1,0,0,0], modeling_pi0.py:590,619);
//   3. per-row (max, sum-exp) of every tile are exchanged through distributed shared memory, so the softmax uses the
//      EXACT global row max / row sum and P is rounded to bf16 after normalisation (reference ledger);
//   4. partial P.V tiles are reduce-scattered by query row over DSMEM (16-byte remote stores) and summed in rank order
//      (deterministic), rounded to bf16 and stored.
#include "attn_common.cuh"
#include "gemm_skinny.cuh"
#include "host_common.h"
#include "ops.h"

namespace cvb {

extern unsigned long long* g_skinny_ts;

namespace {

struct DecodeParams {
  const bf16* q;
  long q_bs, q_rs;
  const bf16* k0;
  const bf16* v0;
  long kv0_bs, kv0_rs;
  const int* kv0_len_dev;
  int kv0_len;
  int q_per_kv_batch;
  const bf16* k1;
  const bf16* v1;
  long kv1_bs, kv1_rs;
  int kv1_len;
  int suffix_mask;
  bf16* out;
  long o_bs, o_rs;
  int heads, kv_heads, tq, head_dim;
  float scale;
  int kv0_static;      // segment 0 and its length are not produced by the preceding kernel (PDL prefetch allowed)
  const float2* rope;  // [kv batches][tq][head_dim/2] (cos, sin) of position kv0_len + t, or nullptr (no RoPE)
  int C;               // cluster size = key tiles
  int rows_per;        // query rows finalised per CTA
  unsigned long long* ts;  // diagnostics (cvb_debug_set_timestamps) or nullptr
};

#define AD_TS(i)                                                                                   \
  do {                                                                                             \
    if (p.ts != nullptr && threadIdx.x == 0)                                                       \
      p.ts[((blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * 8 + (i)] = gtimer(); \
  } while (0)

__device__ __forceinline__ void st_cluster_v2f(uint32_t addr, float a, float b) {
  asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};" ::"r"(addr), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ void st_cluster_v4f(uint32_t addr, float a, float b, float c, float d) {
  asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d)
               : "memory");
}

// rotate 8 (x1, x2) pairs: x1' = x1 c - x2 s ; x2' = x2 c + x1 s   (separate mul / add as torch does)
__device__ __forceinline__ void rope8(uint4& v1, uint4& v2, const float2* cs) {
  uint32_t u1[4] = {v1.x, v1.y, v1.z, v1.w}, u2[4] = {v2.x, v2.y, v2.z, v2.w};
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const float2 a = unpack_bf16x2(u1[e]), b = unpack_bf16x2(u2[e]);
    const float2 t0 = cs[2 * e], t1 = cs[2 * e + 1];
    u1[e] = pack_bf16x2(__fsub_rn(__fmul_rn(a.x, t0.x), __fmul_rn(b.x, t0.y)),
                        __fsub_rn(__fmul_rn(a.y, t1.x), __fmul_rn(b.y, t1.y)));
    u2[e] = pack_bf16x2(__fadd_rn(__fmul_rn(b.x, t0.x), __fmul_rn(a.x, t0.y)),
                        __fadd_rn(__fmul_rn(b.y, t1.x), __fmul_rn(a.y, t1.y)));
  }
  v1 = make_uint4(u1[0], u1[1], u1[2], u1[3]);
  v2 = make_uint4(u2[0], u2[1], u2[2], u2[3]);
}

template <int HD>
__global__ void __launch_bounds__(ATT_THREADS) attn_decode_kernel(const DecodeParams p) {
  constexpr int LDS = HD + 8;
  constexpr int HALF = HD / 2;
  extern __shared__ __align__(16) uint8_t smem_raw[];
  bf16* Qs = reinterpret_cast<bf16*>(smem_raw);                      // [64][LDS]
  bf16* Ks = Qs + BQ * LDS;                                          // [64][LDS]
  bf16* Vs = Ks + BKV * LDS;                                         // [64][LDS]
  float2* stats = reinterpret_cast<float2*>(Vs + BKV * LDS);         // [C][64] (row max, row sum-exp) per source tile
  float* orecv = reinterpret_cast<float*>(smem_raw);                 // [C][rows_per][HD] fp32, aliases Q + K

  const int C = p.C;
  const int rank = static_cast<int>(cluster_ctarank());
  const int b = blockIdx.z, kvh = blockIdx.y;
  const int G = p.heads / p.kv_heads;
  const int rows_total = G * p.tq;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kvb = b / p.q_per_kv_batch;
  const int tile = rank;
  const float2* rope = p.rope != nullptr ? p.rope + static_cast<long>(kvb) * p.tq * HALF : nullptr;
  AD_TS(0);

  // ---- stage K / V prefix rows by cp.async.  With kv0_static the prefix cache (and its length) does not depend on the
  // preceding kernel, so these loads are put in flight BEFORE the programmatic-dependency wait.
  pdl_launch();
  if (!p.kv0_static) pdl_wait();
  const int n0 = p.kv0_len_dev != nullptr ? p.kv0_len_dev[kvb] : p.kv0_len;
  const int nk = n0 + p.kv1_len;
  {
    constexpr int CH = HD / 8;
    for (int idx = threadIdx.x; idx < BKV * CH; idx += ATT_THREADS) {
      const int r = idx / CH, c = (idx % CH) * 8;
      const int j = tile * BKV + r;
      const bool pre = j < n0;
      const bf16* src = pre ? p.k0 + kvb * p.kv0_bs + j * p.kv0_rs + kvh * HD + c : p.q;
      cp_async16(smem_u32(Ks + r * LDS + c), src, pre);  // rows >= n0 are zero-filled here, suffix rows overwritten below
    }
    cp_async_commit();
    for (int idx = threadIdx.x; idx < BKV * CH; idx += ATT_THREADS) {
      const int r = idx / CH, c = (idx % CH) * 8;
      const int j = tile * BKV + r;
      const bool pre = j < n0;
      const bf16* src = pre ? p.v0 + kvb * p.kv0_bs + j * p.kv0_rs + kvh * HD + c : p.q;
      cp_async16(smem_u32(Vs + r * LDS + c), src, pre);
    }
    cp_async_commit();
    if (p.kv0_static) pdl_wait();
    // Q rows (+ RoPE): work item = (row, 8-wide chunk of the first half).  All global loads of a batch of items are
    // issued before any is consumed (the items are independent: one exposed L2 round trip per batch, not per item).
    constexpr int CH2 = HALF / 8;
    constexpr int QB = 4;
    const int q_items = rows_total * CH2;
    for (int base = 0; base < q_items; base += QB * ATT_THREADS) {
      uint4 v1[QB], v2[QB];
      float4 cs[QB][4];
#pragma unroll
      for (int u = 0; u < QB; ++u) {
        const int idx = base + u * ATT_THREADS + threadIdx.x;
        if (idx < q_items) {
          const int r = idx / CH2, c = (idx % CH2) * 8;
          const int hl = r / p.tq, t = r % p.tq;
          const bf16* qp = p.q + b * p.q_bs + t * p.q_rs + (kvh * G + hl) * HD + c;
          v1[u] = *reinterpret_cast<const uint4*>(qp);
          v2[u] = *reinterpret_cast<const uint4*>(qp + HALF);
          if (rope != nullptr) {
            const float4* tp = reinterpret_cast<const float4*>(rope + t * HALF + c);
#pragma unroll
            for (int e = 0; e < 4; ++e) cs[u][e] = tp[e];
          }
        }
      }
#pragma unroll
      for (int u = 0; u < QB; ++u) {
        const int idx = base + u * ATT_THREADS + threadIdx.x;
        if (idx < q_items) {
          const int r = idx / CH2, c = (idx % CH2) * 8;
          if (rope != nullptr) rope8(v1[u], v2[u], reinterpret_cast<const float2*>(cs[u]));
          *reinterpret_cast<uint4*>(Qs + r * LDS + c) = v1[u];
          *reinterpret_cast<uint4*>(Qs + r * LDS + HALF + c) = v2[u];
        }
      }
    }
    // padding rows of the 64-row tile
    for (int idx = q_items + threadIdx.x; idx < BQ * CH2; idx += ATT_THREADS) {
      const int r = idx / CH2, c = (idx % CH2) * 8;
      *reinterpret_cast<uint4*>(Qs + r * LDS + c) = make_uint4(0, 0, 0, 0);
      *reinterpret_cast<uint4*>(Qs + r * LDS + HALF + c) = make_uint4(0, 0, 0, 0);
    }
    cp_async_wait<1>();  // this thread's K chunks have landed; the suffix rows it overwrites are its own or ordered below
    __syncthreads();
    // suffix keys that fall into this tile
    const int j_lo = max(n0, tile * BKV), j_hi = min(nk, tile * BKV + BKV);
    for (int idx = threadIdx.x; idx < max(0, j_hi - j_lo) * CH2; idx += ATT_THREADS) {
      const int j = j_lo + idx / CH2, c = (idx % CH2) * 8;
      const int t = j - n0;
      const bf16* kp = p.k1 + b * p.kv1_bs + t * p.kv1_rs + kvh * HD + c;
      uint4 v1 = *reinterpret_cast<const uint4*>(kp);
      uint4 v2 = *reinterpret_cast<const uint4*>(kp + HALF);
      if (rope != nullptr) rope8(v1, v2, rope + t * HALF + c);
      const int r = j - tile * BKV;
      *reinterpret_cast<uint4*>(Ks + r * LDS + c) = v1;
      *reinterpret_cast<uint4*>(Ks + r * LDS + HALF + c) = v2;
    }
    cp_async_wait<0>();  // V zero-fill of the suffix rows has landed before they are overwritten
    __syncthreads();
    for (int idx = threadIdx.x; idx < max(0, j_hi - j_lo) * (HD / 8); idx += ATT_THREADS) {
      const int j = j_lo + idx / (HD / 8), c = (idx % (HD / 8)) * 8;
      const bf16* vp = p.v1 + b * p.kv1_bs + (j - n0) * p.kv1_rs + kvh * HD + c;
      *reinterpret_cast<uint4*>(Vs + (j - tile * BKV) * LDS + c) = *reinterpret_cast<const uint4*>(vp);
    }
    __syncthreads();
  }

  AD_TS(1);
  const int r_lo = warp * 16 + (lane >> 2);
  const int r_hi = r_lo + 8;
  const int t_lo = r_lo % p.tq, t_hi = r_hi % p.tq;
  const bool warp_active = warp * 16 < rows_total;
  auto key_ok = [&](int j, int t) -> bool {
    if (j >= nk) return false;
    if (p.suffix_mask && j >= n0 && t == 0) return (j - n0) == 0;
    return true;
  };

  // ---- S = Q K^T, local row statistics
  float s[8][4];
  float m_lo = -INFINITY, m_hi = -INFINITY, l_lo = 0.f, l_hi = 0.f;
  if (warp_active) {
    qk_tile<HD>(Qs, Ks, warp, lane, s);
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int j = tile * BKV + nt * 8 + (lane & 3) * 2 + e;
        s[nt][e] = key_ok(j, t_lo) ? s[nt][e] * p.scale : -INFINITY;
        s[nt][2 + e] = key_ok(j, t_hi) ? s[nt][2 + e] * p.scale : -INFINITY;
        m_lo = fmaxf(m_lo, s[nt][e]);
        m_hi = fmaxf(m_hi, s[nt][2 + e]);
      }
    }
    m_lo = fmaxf(m_lo, __shfl_xor_sync(0xffffffffu, m_lo, 1));
    m_lo = fmaxf(m_lo, __shfl_xor_sync(0xffffffffu, m_lo, 2));
    m_hi = fmaxf(m_hi, __shfl_xor_sync(0xffffffffu, m_hi, 1));
    m_hi = fmaxf(m_hi, __shfl_xor_sync(0xffffffffu, m_hi, 2));
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        l_lo += (s[nt][e] == -INFINITY) ? 0.f : expf(s[nt][e] - m_lo);
        l_hi += (s[nt][2 + e] == -INFINITY) ? 0.f : expf(s[nt][2 + e] - m_hi);
      }
    }
    l_lo += __shfl_xor_sync(0xffffffffu, l_lo, 1);
    l_lo += __shfl_xor_sync(0xffffffffu, l_lo, 2);
    l_hi += __shfl_xor_sync(0xffffffffu, l_hi, 1);
    l_hi += __shfl_xor_sync(0xffffffffu, l_hi, 2);
    // broadcast (max, sum) of my rows to every CTA of the cluster; the 4 lanes of a quad share the peers
    const uint32_t sbase = smem_u32(stats + rank * BQ);
    for (int peer = lane & 3; peer < C; peer += 4) {
      const uint32_t dst = mapa_shared(sbase, peer);
      st_cluster_v2f(dst + r_lo * 8, m_lo, l_lo);
      st_cluster_v2f(dst + r_hi * 8, m_hi, l_hi);
    }
  }
  AD_TS(2);
  cluster_sync_all();  // statistics visible everywhere; every CTA is past Q.K^T, so Q / K may be overwritten (orecv)

  // ---- exact softmax with the global statistics, P -> bf16, O_partial = P V
  AD_TS(3);
  float o[HD / 8][4];
  cp_async_wait<0>();
  __syncthreads();  // every thread's V chunks are in shared memory
  if (warp_active) {
    float M_lo = -INFINITY, M_hi = -INFINITY;
    for (int c = 0; c < C; ++c) {
      M_lo = fmaxf(M_lo, stats[c * BQ + r_lo].x);
      M_hi = fmaxf(M_hi, stats[c * BQ + r_hi].x);
    }
    float L_lo = 0.f, L_hi = 0.f;
    for (int c = 0; c < C; ++c) {
      const float2 a = stats[c * BQ + r_lo], bq = stats[c * BQ + r_hi];
      L_lo += a.x == -INFINITY ? 0.f : a.y * expf(a.x - M_lo);
      L_hi += bq.x == -INFINITY ? 0.f : bq.y * expf(bq.x - M_hi);
    }
    const float inv_lo = L_lo > 0.f ? 1.0f / L_lo : 0.f;
    const float inv_hi = L_hi > 0.f ? 1.0f / L_hi : 0.f;
    uint32_t pa[4][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      float pv[4];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        pv[e] = s[nt][e] == -INFINITY ? 0.f : expf(s[nt][e] - M_lo) * inv_lo;
        pv[2 + e] = s[nt][2 + e] == -INFINITY ? 0.f : expf(s[nt][2 + e] - M_hi) * inv_hi;
      }
      pa[nt >> 1][(nt & 1) * 2 + 0] = pack_bf16x2(pv[0], pv[1]);
      pa[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16x2(pv[2], pv[3]);
    }
#pragma unroll
    for (int dt = 0; dt < HD / 8; ++dt) o[dt][0] = o[dt][1] = o[dt][2] = o[dt][3] = 0.f;
    const int mi = lane >> 3, ri = lane & 7;
    const uint32_t v_base = smem_u32(Vs + ((mi & 1) * 8 + ri) * LDS + (mi >> 1) * 8);
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
      for (int dp = 0; dp < HD / 16; ++dp) {
        uint32_t b0, b1, b2, b3;
        ldsm_x4_t(v_base + (kk * 16 * LDS) * 2 + dp * 32, b0, b1, b2, b3);
        mma_bf16(o[2 * dp], pa[kk][0], pa[kk][1], pa[kk][2], pa[kk][3], b0, b1);
        mma_bf16(o[2 * dp + 1], pa[kk][0], pa[kk][1], pa[kk][2], pa[kk][3], b2, b3);
      }
    }
    // ---- reduce-scatter by query row: even lanes of a pair ship 4 columns of row_lo, odd lanes of row_hi
    const bool odd = lane & 1;
    const int my_row = odd ? r_hi : r_lo;
    const bool row_ok = my_row < rows_total;
    const int owner = row_ok ? my_row / p.rows_per : 0;
    const int row_local = my_row - owner * p.rows_per;
    const uint32_t obase = mapa_shared(smem_u32(orecv), owner) +
                           static_cast<uint32_t>((rank * p.rows_per + row_local) * HD) * 4u;
#pragma unroll
    for (int dt = 0; dt < HD / 8; ++dt) {
      // partner = lane ^ 1 holds the neighbouring 2 columns of the same rows
      const float send0 = odd ? o[dt][0] : o[dt][2];
      const float send1 = odd ? o[dt][1] : o[dt][3];
      const float got0 = __shfl_xor_sync(0xffffffffu, send0, 1);
      const float got1 = __shfl_xor_sync(0xffffffffu, send1, 1);
      // even lane: row_lo cols [4q, 4q+4) = {own o0,o1, partner o0,o1}; odd lane: row_hi = {partner o2,o3, own o2,o3}
      const int col = dt * 8 + (lane & 2) * 2;
      if (row_ok) {
        if (!odd)
          st_cluster_v4f(obase + col * 4, o[dt][0], o[dt][1], got0, got1);
        else
          st_cluster_v4f(obase + col * 4, got0, got1, o[dt][2], o[dt][3]);
      }
    }
  }
  AD_TS(4);
  cluster_sync_all();
  AD_TS(5);

  // ---- final: sum the C partial rows I own, round to bf16, store
  {
    const int row0 = rank * p.rows_per;
    const int nrows = max(0, min(p.rows_per, rows_total - row0));
    for (int idx = threadIdx.x; idx < nrows * (HD / 4); idx += ATT_THREADS) {
      const int rl = idx / (HD / 4), d = (idx % (HD / 4)) * 4;
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int c = 0; c < C; ++c) {
        const float4 a = *reinterpret_cast<const float4*>(orecv + (c * p.rows_per + rl) * HD + d);
        acc.x += a.x, acc.y += a.y, acc.z += a.z, acc.w += a.w;
      }
      const int r = row0 + rl;
      const int hl = r / p.tq, t = r % p.tq;
      bf16* op = p.out + b * p.o_bs + t * p.o_rs + (kvh * G + hl) * HD + d;
      *reinterpret_cast<uint2*>(op) = make_uint2(pack_bf16x2(acc.x, acc.y), pack_bf16x2(acc.z, acc.w));
    }
  }
  AD_TS(6);
}

template <int HD>
int launch_decode(cudaStream_t st, const DecodeParams& p, dim3 grid) {
  const int smem = 3 * BQ * (HD + 8) * 2 + p.C * BQ * 8;
  auto kern = attn_decode_kernel<HD>;
  CVB_TRY(ensure_dyn_smem(kern, smem));
  CVB_TRY(launch_pdl(kern, grid, dim3(ATT_THREADS), smem, st, p.C, p));
  CVB_LAUNCHED();
  return 0;
}

}  // namespace

// Eligibility: all query rows of a KV group fit one 64-row tile, <= 8 key tiles, head_dim 64 / 128 / 256.
bool attention_decode_eligible(const AttnCall& c) {
  const int G = c.heads / c.kv_heads;
  const int max_keys = (c.kv0_len_dev != nullptr ? c.kv0_max : c.kv0_len) + (c.k1 != nullptr ? c.kv1_len : 0);
  const int C = (max_keys + BKV - 1) / BKV;
  return G * c.tq <= BQ && C >= 1 && C <= 8 && (c.head_dim == 64 || c.head_dim == 128 || c.head_dim == 256) &&
         !c.force_two_pass;
}

int attention_decode(cudaStream_t st, const AttnCall& c, const float2* rope) {
  CVB_REQUIRE(attention_decode_eligible(c), "shape not eligible for the cluster decode attention");
  DecodeParams p;
  p.q = c.q, p.q_bs = c.q_batch_stride, p.q_rs = c.q_row_stride;
  p.k0 = c.k0, p.v0 = c.v0, p.kv0_bs = c.kv0_batch_stride, p.kv0_rs = c.kv0_row_stride;
  p.kv0_len_dev = c.kv0_len_dev, p.kv0_len = c.kv0_len, p.q_per_kv_batch = c.q_per_kv_batch;
  p.k1 = c.k1, p.v1 = c.v1, p.kv1_bs = c.kv1_batch_stride, p.kv1_rs = c.kv1_row_stride;
  p.kv1_len = c.k1 != nullptr ? c.kv1_len : 0;
  p.suffix_mask = c.suffix_mask;
  p.out = c.out, p.o_bs = c.o_batch_stride, p.o_rs = c.o_row_stride;
  p.heads = c.heads, p.kv_heads = c.kv_heads, p.tq = c.tq, p.head_dim = c.head_dim, p.scale = c.scale;
  p.rope = rope;
  p.kv0_static = c.kv0_static;
  p.ts = g_skinny_ts;
  const int G = c.heads / c.kv_heads;
  const int max_keys = (c.kv0_len_dev != nullptr ? c.kv0_max : c.kv0_len) + p.kv1_len;
  p.C = (max_keys + BKV - 1) / BKV;
  p.rows_per = (G * c.tq + p.C - 1) / p.C;
  CVB_REQUIRE((long)p.C * p.rows_per * c.head_dim * 4 <= 2L * BQ * (c.head_dim + 8) * 2,
              "partial-output receive buffer must fit the Q + K staging area");
  dim3 grid(p.C, c.kv_heads, c.batches);
  if (c.head_dim == 64) return launch_decode<64>(st, p, grid);
  if (c.head_dim == 128) return launch_decode<128>(st, p, grid);
  return launch_decode<256>(st, p, grid);
}

}  // namespace cvb
