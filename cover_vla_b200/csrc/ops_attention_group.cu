// Denoise-step attention, "rephrase-grouped" schedule (eager_attention_forward, paligemma_with_expert.py:376-434, as
// called from PI0FlowMatching.denoise_step, modeling_pi0.py:717-752).
//
// The K candidates sampled from one rephrase share that rephrase's prefix KV cache.  All their query rows
// (K candidates x G heads x tq suffix tokens = 200 for 5 x 8 x 5) are folded into one row space per rephrase and cut
// into 16-row tiles: grid = (row tiles, kv heads, rephrases) ~ 104 CTAs, no cluster, no cross-CTA reduction.  The
// candidates' own suffix keys are appended to the key axis (K x tq extra keys) under a block-diagonal mask, so one
// K/V stream serves every candidate of the rephrase.
//
// Per CTA (4 warps): K tiles then V tiles stream once through a cp.async double buffer;
//   * Q.K^T: warp w multiplies the 16 query rows with its 16-key quarter of the tile (mma.sync m16n8k16, fp32) and
//     parks the scaled, masked logits in shared memory [16][keys] - the exact row max / row sum are known before P is
//     rounded to bf16 (reference ledger);
//   * P.V: warp w owns head_dim/4 output columns, rebuilds the bf16 P fragments from the parked logits and
//     accumulates its [16 x head_dim/4] slice in registers;
//   * RoPE (apply_rope, :34-57) of Q and of the suffix keys is applied while staging, from the per-sample (cos, sin)
//     table; with kv0_static the first prefix tile is in flight before the programmatic-dependency wait.
#include "attn_common.cuh"
#include "host_common.h"
#include "ops.h"

namespace cvb {

namespace {

struct GroupParams {
  const bf16* q;
  long q_bs, q_rs;
  const bf16* k0;
  const bf16* v0;
  long kv0_bs, kv0_rs;
  const int* kv0_len_dev;
  int kv0_len;
  int cands;  // candidates per kv batch (q_per_kv_batch)
  const bf16* k1;
  const bf16* v1;
  long kv1_bs, kv1_rs;
  int kv1_len;
  int suffix_mask;
  bf16* out;
  long o_bs, o_rs;
  int heads, kv_heads, tq;
  float scale;
  const float2* rope;
  int kv0_static;
  int s_ld;  // floats per parked-logit row (>= max keys rounded to 64, + 4 pad)
};

constexpr int GR = 16;  // query rows per CTA
constexpr int NS = 4;   // K/V tiles in the cp.async ring (3 in flight: the stream is latency-, not bandwidth-bound)

__device__ __forceinline__ void rope8g(uint4& v1, uint4& v2, const float2* cs) {
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
__global__ void __launch_bounds__(ATT_THREADS) attn_group_kernel(const GroupParams p) {
  constexpr int LDS = HD + 8;
  constexpr int HALF = HD / 2;
  constexpr int CH = HD / 8;
  constexpr int CH2 = HALF / 8;
  constexpr int DW = HD / 4;  // output columns per warp
  extern __shared__ __align__(16) uint8_t smem_raw[];
  bf16* Qs = reinterpret_cast<bf16*>(smem_raw);             // [16][LDS]
  bf16* KV = Qs + GR * LDS;                                  // NS x [64][LDS]
  float* Sb = reinterpret_cast<float*>(KV + NS * BKV * LDS);  // [16][s_ld] parked logits
  float* Mx = Sb + GR * p.s_ld;                             // [16] row max
  float* Li = Mx + GR;                                      // [16] 1 / row sum

  const int kvb = blockIdx.z, kvh = blockIdx.y, rt = blockIdx.x;
  const int G = p.heads / p.kv_heads;
  const int rows_total = p.cands * G * p.tq;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float2* rope = p.rope != nullptr ? p.rope + static_cast<long>(kvb) * p.tq * HALF : nullptr;

  pdl_launch();
  if (!p.kv0_static) pdl_wait();
  const int n0 = p.kv0_len_dev != nullptr ? p.kv0_len_dev[kvb] : p.kv0_len;
  const int nk = n0 + p.cands * p.kv1_len;
  const int n_tiles = (nk + BKV - 1) / BKV;

  // prefix rows of a K or V tile by cp.async (suffix rows are zero-filled here and written through registers later)
  auto issue = [&](int it) {
    const bool is_v = it >= n_tiles;
    const int tile = is_v ? it - n_tiles : it;
    bf16* dst = KV + (it % NS) * BKV * LDS;
    const bf16* base = (is_v ? p.v0 : p.k0) + kvb * p.kv0_bs + kvh * HD;
    for (int idx = threadIdx.x; idx < BKV * CH; idx += ATT_THREADS) {
      const int r = idx / CH, c = (idx % CH) * 8;
      const int j = tile * BKV + r;
      const bool pre = j < n0;
      cp_async16(smem_u32(dst + r * LDS + c), pre ? base + j * p.kv0_rs + c : p.q, pre);
    }
  };
  // suffix keys / values of the candidates that fall into this tile (K gets RoPE)
  auto fill_suffix = [&](int it) {
    const bool is_v = it >= n_tiles;
    const int tile = is_v ? it - n_tiles : it;
    bf16* dst = KV + (it % NS) * BKV * LDS;
    const int j_lo = max(n0, tile * BKV), j_hi = min(nk, tile * BKV + BKV);
    if (j_hi <= j_lo) return;
    if (!is_v) {
      for (int idx = threadIdx.x; idx < (j_hi - j_lo) * CH2; idx += ATT_THREADS) {
        const int j = j_lo + idx / CH2, c = (idx % CH2) * 8;
        const int cand = (j - n0) / p.kv1_len, t = (j - n0) % p.kv1_len;
        const bf16* kp = p.k1 + (static_cast<long>(kvb) * p.cands + cand) * p.kv1_bs + t * p.kv1_rs + kvh * HD + c;
        uint4 v1 = *reinterpret_cast<const uint4*>(kp);
        uint4 v2 = *reinterpret_cast<const uint4*>(kp + HALF);
        if (rope != nullptr) rope8g(v1, v2, rope + t * HALF + c);
        const int r = j - tile * BKV;
        *reinterpret_cast<uint4*>(dst + r * LDS + c) = v1;
        *reinterpret_cast<uint4*>(dst + r * LDS + HALF + c) = v2;
      }
    } else {
      for (int idx = threadIdx.x; idx < (j_hi - j_lo) * CH; idx += ATT_THREADS) {
        const int j = j_lo + idx / CH, c = (idx % CH) * 8;
        const int cand = (j - n0) / p.kv1_len, t = (j - n0) % p.kv1_len;
        const bf16* vp = p.v1 + (static_cast<long>(kvb) * p.cands + cand) * p.kv1_bs + t * p.kv1_rs + kvh * HD + c;
        *reinterpret_cast<uint4*>(dst + (j - tile * BKV) * LDS + c) = *reinterpret_cast<const uint4*>(vp);
      }
    }
  };

  for (int i = 0; i < NS - 1; ++i) {
    if (i < 2 * n_tiles) issue(i);
    cp_async_commit();
  }
  if (p.kv0_static) pdl_wait();

  // ---- Q rows (+ RoPE).  row -> (candidate, head, token): r = (cand * G + h) * tq + t
  for (int idx = threadIdx.x; idx < GR * CH2; idx += ATT_THREADS) {
    const int rl = idx / CH2, c = (idx % CH2) * 8;
    const int r = rt * GR + rl;
    uint4 v1 = make_uint4(0, 0, 0, 0), v2 = v1;
    if (r < rows_total) {
      const int t = r % p.tq, ch = r / p.tq, hl = ch % G, cand = ch / G;
      const bf16* qp = p.q + (static_cast<long>(kvb) * p.cands + cand) * p.q_bs + t * p.q_rs + (kvh * G + hl) * HD + c;
      v1 = *reinterpret_cast<const uint4*>(qp);
      v2 = *reinterpret_cast<const uint4*>(qp + HALF);
      if (rope != nullptr) rope8g(v1, v2, rope + t * HALF + c);
    }
    *reinterpret_cast<uint4*>(Qs + rl * LDS + c) = v1;
    *reinterpret_cast<uint4*>(Qs + rl * LDS + HALF + c) = v2;
  }

  // the two query rows of this thread (mma accumulator layout) and what they may see
  const int r_lo = rt * GR + (lane >> 2), r_hi = r_lo + 8;
  const int t_lo = r_lo % p.tq, t_hi = r_hi % p.tq;
  const int c_lo = r_lo / (p.tq * G), c_hi = r_hi / (p.tq * G);
  auto key_ok = [&](int j, int t, int cand) -> bool {
    if (j >= nk) return false;
    if (j < n0) return true;
    const int kc = (j - n0) / p.kv1_len, kt = (j - n0) % p.kv1_len;
    if (kc != cand) return false;
    if (p.suffix_mask && t == 0) return kt == 0;
    return true;
  };

  float o[DW / 8][4];
#pragma unroll
  for (int i = 0; i < DW / 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
  float m_lo = 0.f, m_hi = 0.f, i_lo = 0.f, i_hi = 0.f;

  const int mi = lane >> 3, ri = lane & 7;
  for (int it = 0; it < 2 * n_tiles; ++it) {
    if (it + NS - 1 < 2 * n_tiles) issue(it + NS - 1);
    cp_async_commit();
    cp_async_wait<NS - 1>();
    __syncthreads();
    fill_suffix(it);
    if (it == n_tiles) {
      // ---- exact softmax statistics over the parked logits: warp w owns rows 4w .. 4w+3
      for (int rr = 0; rr < 4; ++rr) {
        const int row = warp * 4 + rr;
        const float* sr = Sb + row * p.s_ld;
        float mx = -INFINITY;
        for (int j = lane; j < n_tiles * BKV; j += 32) mx = fmaxf(mx, sr[j]);
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, s));
        float sum = 0.f;
        for (int j = lane; j < n_tiles * BKV; j += 32) sum += sr[j] == -INFINITY ? 0.f : expf(sr[j] - mx);
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, s);
        if (lane == 0) {
          Mx[row] = mx;
          Li[row] = sum > 0.f ? 1.0f / sum : 0.f;
        }
      }
    }
    __syncthreads();
    const bf16* buf = KV + (it % NS) * BKV * LDS;
    if (it < n_tiles) {
      // ---- S[16 x 16] = Q[16 x HD] . K[16 keys of this warp]^T
      float s[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
      const uint32_t q_base = smem_u32(Qs + (lane & 15) * LDS + (lane >> 4) * 8);
      const uint32_t k_base = smem_u32(buf + (warp * 16 + (mi >> 1) * 8 + ri) * LDS + (mi & 1) * 8);
#pragma unroll 4
      for (int ks = 0; ks < HD / 16; ++ks) {
        uint32_t a0, a1, a2, a3, b0, b1, b2, b3;
        ldsm_x4(q_base + ks * 32, a0, a1, a2, a3);
        ldsm_x4(k_base + ks * 32, b0, b1, b2, b3);
        mma_bf16(s[0], a0, a1, a2, a3, b0, b1);
        mma_bf16(s[1], a0, a1, a2, a3, b2, b3);
      }
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) {
        const int j = it * BKV + warp * 16 + nt * 8 + (lane & 3) * 2;
        float2 lo, hi;
        lo.x = key_ok(j, t_lo, c_lo) ? s[nt][0] * p.scale : -INFINITY;
        lo.y = key_ok(j + 1, t_lo, c_lo) ? s[nt][1] * p.scale : -INFINITY;
        hi.x = key_ok(j, t_hi, c_hi) ? s[nt][2] * p.scale : -INFINITY;
        hi.y = key_ok(j + 1, t_hi, c_hi) ? s[nt][3] * p.scale : -INFINITY;
        *reinterpret_cast<float2*>(Sb + (lane >> 2) * p.s_ld + j) = lo;
        *reinterpret_cast<float2*>(Sb + ((lane >> 2) + 8) * p.s_ld + j) = hi;
      }
    } else {
      // ---- O[16 x DW] += bf16(P[16 x 64]) . V[64 x DW]
      const int tile = it - n_tiles;
      if (tile == 0) {
        m_lo = Mx[lane >> 2], m_hi = Mx[(lane >> 2) + 8];
        i_lo = Li[lane >> 2], i_hi = Li[(lane >> 2) + 8];
      }
      const float* s_lo = Sb + (lane >> 2) * p.s_ld + tile * BKV + (lane & 3) * 2;
      const float* s_hi = s_lo + 8 * p.s_ld;
      const uint32_t v_base = smem_u32(buf + ((mi & 1) * 8 + ri) * LDS + warp * DW + (mi >> 1) * 8);
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const float2 a = *reinterpret_cast<const float2*>(s_lo + kk * 16);
        const float2 b = *reinterpret_cast<const float2*>(s_hi + kk * 16);
        const float2 c = *reinterpret_cast<const float2*>(s_lo + kk * 16 + 8);
        const float2 d = *reinterpret_cast<const float2*>(s_hi + kk * 16 + 8);
        auto pr = [](float x, float m, float inv) { return x == -INFINITY ? 0.f : expf(x - m) * inv; };
        const uint32_t pa0 = pack_bf16x2(pr(a.x, m_lo, i_lo), pr(a.y, m_lo, i_lo));
        const uint32_t pa1 = pack_bf16x2(pr(b.x, m_hi, i_hi), pr(b.y, m_hi, i_hi));
        const uint32_t pa2 = pack_bf16x2(pr(c.x, m_lo, i_lo), pr(c.y, m_lo, i_lo));
        const uint32_t pa3 = pack_bf16x2(pr(d.x, m_hi, i_hi), pr(d.y, m_hi, i_hi));
#pragma unroll
        for (int dp = 0; dp < DW / 16; ++dp) {
          uint32_t b0, b1, b2, b3;
          ldsm_x4_t(v_base + (kk * 16 * LDS) * 2 + dp * 32, b0, b1, b2, b3);
          mma_bf16(o[2 * dp], pa0, pa1, pa2, pa3, b0, b1);
          mma_bf16(o[2 * dp + 1], pa0, pa1, pa2, pa3, b2, b3);
        }
      }
    }
    __syncthreads();  // the buffer just consumed is refilled by the next iteration's prefetch
  }

  // ---- store: row -> (candidate, head, token)
  auto out_ptr = [&](int r) -> bf16* {
    const int t = r % p.tq, ch = r / p.tq, hl = ch % G, cand = ch / G;
    return p.out + (static_cast<long>(kvb) * p.cands + cand) * p.o_bs + t * p.o_rs + (kvh * G + hl) * HD + warp * DW;
  };
  if (r_lo < rows_total) {
    bf16* op = out_ptr(r_lo);
#pragma unroll
    for (int i = 0; i < DW / 8; ++i)
      *reinterpret_cast<uint32_t*>(op + i * 8 + (lane & 3) * 2) = pack_bf16x2(o[i][0], o[i][1]);
  }
  if (r_hi < rows_total) {
    bf16* op = out_ptr(r_hi);
#pragma unroll
    for (int i = 0; i < DW / 8; ++i)
      *reinterpret_cast<uint32_t*>(op + i * 8 + (lane & 3) * 2) = pack_bf16x2(o[i][2], o[i][3]);
  }
}

template <int HD>
int launch_group(cudaStream_t st, const GroupParams& p, dim3 grid) {
  const int smem = (GR + NS * BKV) * (HD + 8) * 2 + GR * p.s_ld * 4 + 2 * GR * 4;
  auto kern = attn_group_kernel<HD>;
  CVB_TRY(ensure_dyn_smem(kern, smem));
  CVB_TRY(launch_pdl(kern, grid, dim3(ATT_THREADS), smem, st, 1, p));
  CVB_LAUNCHED();
  return 0;
}

}  // namespace

// Eligibility: two key segments with batches = kv batches x q_per_kv_batch, head_dim 64 / 128 / 256, logits fit smem.
bool attention_group_eligible(const AttnCall& c) {
  if (c.k1 == nullptr || c.force_two_pass) return false;
  if (!(c.head_dim == 64 || c.head_dim == 128 || c.head_dim == 256)) return false;
  if (c.q_per_kv_batch < 1 || c.batches % c.q_per_kv_batch != 0) return false;
  const int max_keys = (c.kv0_len_dev != nullptr ? c.kv0_max : c.kv0_len) + c.q_per_kv_batch * c.kv1_len;
  const int s_ld = (max_keys + BKV - 1) / BKV * BKV + 4;
  const long smem = (long)(GR + NS * BKV) * (c.head_dim + 8) * 2 + (long)GR * s_ld * 4 + 2 * GR * 4;
  return smem <= 220 * 1024;
}

int attention_group(cudaStream_t st, const AttnCall& c) {
  CVB_REQUIRE(attention_group_eligible(c), "shape not eligible for the rephrase-grouped attention");
  GroupParams p;
  p.q = c.q, p.q_bs = c.q_batch_stride, p.q_rs = c.q_row_stride;
  p.k0 = c.k0, p.v0 = c.v0, p.kv0_bs = c.kv0_batch_stride, p.kv0_rs = c.kv0_row_stride;
  p.kv0_len_dev = c.kv0_len_dev, p.kv0_len = c.kv0_len, p.cands = c.q_per_kv_batch;
  p.k1 = c.k1, p.v1 = c.v1, p.kv1_bs = c.kv1_batch_stride, p.kv1_rs = c.kv1_row_stride, p.kv1_len = c.kv1_len;
  p.suffix_mask = c.suffix_mask;
  p.out = c.out, p.o_bs = c.o_batch_stride, p.o_rs = c.o_row_stride;
  p.heads = c.heads, p.kv_heads = c.kv_heads, p.tq = c.tq, p.scale = c.scale;
  p.rope = c.rope, p.kv0_static = c.kv0_static;
  const int max_keys = (c.kv0_len_dev != nullptr ? c.kv0_max : c.kv0_len) + c.q_per_kv_batch * c.kv1_len;
  p.s_ld = (max_keys + BKV - 1) / BKV * BKV + 4;
  const int G = c.heads / c.kv_heads;
  const int rows_total = c.q_per_kv_batch * G * c.tq;
  dim3 grid((rows_total + GR - 1) / GR, c.kv_heads, c.batches / c.q_per_kv_batch);
  if (c.head_dim == 64) return launch_group<64>(st, p, grid);
  if (c.head_dim == 128) return launch_group<128>(st, p, grid);
  return launch_group<256>(st, p, grid);
}

}  // namespace cvb
