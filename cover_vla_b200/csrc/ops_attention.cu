// Exact-softmax (two-pass) attention on bf16 tensor cores (mma.sync m16n8k16, fp32 accumulate).
//
// Rounding ledger = the reference eager attention (paligemma_with_expert.py:376-434): logits in fp32
// from bf16 Q/K, * head_dim^-0.5, masked keys dropped, softmax in fp32 with the true row max and row
// sum (pass 1), NORMALISED probabilities rounded to bf16, P@V accumulated in fp32 (pass 2), output
// rounded to bf16.  MQA/GQA query heads that share a KV head are folded into the row dimension of
// the tile, so K/V are staged in shared memory once per CTA for all of them.
//
// Two key segments are supported so the denoise step (modeling_pi0.py:717-752) needs no concat:
// segment 0 = the prefix KV cache of the candidate's rephrase (length read from device memory),
// segment 1 = the candidate's own suffix keys, with the [1,1,0,0,0] suffix mask of
// modeling_pi0.py:590,619 (token 0 sees only suffix key 0).
//
// This is the round-1 attention: attention is ~3 % of the path's FLOPs, the GEMMs are tcgen05.
#include "host_common.h"
#include "ops.h"
#include "ptx.cuh"
#include "attn_common.cuh"

namespace cvb {

namespace {

struct AttnParams {
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
};


template <int HDP>
__device__ __forceinline__ void load_kv_tile(const AttnParams& p, bf16* Ks, bf16* Vs, int tile,
                                             int n0, int nk, int b, int kvb, int kvh, bool want_v) {
  constexpr int LDS = HDP + 8;
  constexpr int CH = HDP / 8;
  for (int idx = threadIdx.x; idx < BKV * CH; idx += ATT_THREADS) {
    const int r = idx / CH, c = (idx % CH) * 8;
    const int j = tile * BKV + r;
    uint4 kv = make_uint4(0, 0, 0, 0), vv = make_uint4(0, 0, 0, 0);
    if (j < nk && c < p.head_dim) {
      const bf16 *kp, *vp;
      if (j < n0) {
        const long off = kvb * p.kv0_bs + j * p.kv0_rs + kvh * p.head_dim + c;
        kp = p.k0 + off, vp = p.v0 + off;
      } else {
        const long off = b * p.kv1_bs + (j - n0) * p.kv1_rs + kvh * p.head_dim + c;
        kp = p.k1 + off, vp = p.v1 + off;
      }
      kv = *reinterpret_cast<const uint4*>(kp);
      if (want_v) vv = *reinterpret_cast<const uint4*>(vp);
    }
    *reinterpret_cast<uint4*>(Ks + r * LDS + c) = kv;
    if (want_v) *reinterpret_cast<uint4*>(Vs + r * LDS + c) = vv;
  }
}

template <int HDP>
__global__ void __launch_bounds__(ATT_THREADS) attn_kernel(const AttnParams p) {
  pdl_wait();
  pdl_launch();
  constexpr int LDS = HDP + 8;
  extern __shared__ __align__(16) uint8_t smem_raw[];
  bf16* Qs = reinterpret_cast<bf16*>(smem_raw);
  bf16* Ks = Qs + BQ * LDS;
  bf16* Vs = Ks + BKV * LDS;

  const int b = blockIdx.z, kvh = blockIdx.y, qt = blockIdx.x;
  const int G = p.heads / p.kv_heads;
  const int rows_total = G * p.tq;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kvb = b / p.q_per_kv_batch;
  const int n0 = p.kv0_len_dev != nullptr ? p.kv0_len_dev[kvb] : p.kv0_len;
  const int nk = n0 + p.kv1_len;
  const int n_tiles = (nk + BKV - 1) / BKV;

  {
    constexpr int CH = HDP / 8;
    for (int idx = threadIdx.x; idx < BQ * CH; idx += ATT_THREADS) {
      const int r = idx / CH, c = (idx % CH) * 8;
      const int i = qt * BQ + r;
      uint4 v = make_uint4(0, 0, 0, 0);
      if (i < rows_total && c < p.head_dim) {
        const int hl = i / p.tq, t = i % p.tq;
        v = *reinterpret_cast<const uint4*>(p.q + b * p.q_bs + t * p.q_rs +
                                            (kvh * G + hl) * p.head_dim + c);
      }
      *reinterpret_cast<uint4*>(Qs + r * LDS + c) = v;
    }
  }

  // the two query rows this thread owns inside the warp's 16-row slab
  const int r_lo = qt * BQ + warp * 16 + (lane >> 2);
  const int r_hi = r_lo + 8;
  const int t_lo = r_lo % p.tq, t_hi = r_hi % p.tq;
  auto key_ok = [&](int j, int t) -> bool {
    if (j >= nk) return false;
    if (p.suffix_mask && j >= n0 && t == 0) return (j - n0) == 0;
    return true;
  };

  // ---------------- pass 1: exact row max and row sum
  float m_lo = -INFINITY, m_hi = -INFINITY, l_lo = 0.f, l_hi = 0.f;
  for (int tile = 0; tile < n_tiles; ++tile) {
    __syncthreads();
    load_kv_tile<HDP>(p, Ks, Vs, tile, n0, nk, b, kvb, kvh, false);
    __syncthreads();
    float s[8][4];
    qk_tile<HDP>(Qs, Ks, warp, lane, s);
    float tm_lo = -INFINITY, tm_hi = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int j = tile * BKV + nt * 8 + (lane & 3) * 2 + e;
        s[nt][e] = key_ok(j, t_lo) ? s[nt][e] * p.scale : -INFINITY;
        s[nt][2 + e] = key_ok(j, t_hi) ? s[nt][2 + e] * p.scale : -INFINITY;
        tm_lo = fmaxf(tm_lo, s[nt][e]);
        tm_hi = fmaxf(tm_hi, s[nt][2 + e]);
      }
    }
    tm_lo = fmaxf(tm_lo, __shfl_xor_sync(0xffffffffu, tm_lo, 1));
    tm_lo = fmaxf(tm_lo, __shfl_xor_sync(0xffffffffu, tm_lo, 2));
    tm_hi = fmaxf(tm_hi, __shfl_xor_sync(0xffffffffu, tm_hi, 1));
    tm_hi = fmaxf(tm_hi, __shfl_xor_sync(0xffffffffu, tm_hi, 2));
    const float nm_lo = fmaxf(m_lo, tm_lo), nm_hi = fmaxf(m_hi, tm_hi);
    float ts_lo = 0.f, ts_hi = 0.f;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        ts_lo += (s[nt][e] == -INFINITY) ? 0.f : expf(s[nt][e] - nm_lo);
        ts_hi += (s[nt][2 + e] == -INFINITY) ? 0.f : expf(s[nt][2 + e] - nm_hi);
      }
    }
    ts_lo += __shfl_xor_sync(0xffffffffu, ts_lo, 1);
    ts_lo += __shfl_xor_sync(0xffffffffu, ts_lo, 2);
    ts_hi += __shfl_xor_sync(0xffffffffu, ts_hi, 1);
    ts_hi += __shfl_xor_sync(0xffffffffu, ts_hi, 2);
    l_lo = (m_lo == -INFINITY ? 0.f : l_lo * expf(m_lo - nm_lo)) + ts_lo;
    l_hi = (m_hi == -INFINITY ? 0.f : l_hi * expf(m_hi - nm_hi)) + ts_hi;
    m_lo = nm_lo, m_hi = nm_hi;
  }
  const float inv_lo = l_lo > 0.f ? 1.0f / l_lo : 0.f;
  const float inv_hi = l_hi > 0.f ? 1.0f / l_hi : 0.f;

  // ---------------- pass 2: O = bf16(P) @ V
  float o[HDP / 8][4];
#pragma unroll
  for (int dt = 0; dt < HDP / 8; ++dt) o[dt][0] = o[dt][1] = o[dt][2] = o[dt][3] = 0.f;
  const int mi = lane >> 3, ri = lane & 7;
  for (int tile = 0; tile < n_tiles; ++tile) {
    __syncthreads();
    load_kv_tile<HDP>(p, Ks, Vs, tile, n0, nk, b, kvb, kvh, true);
    __syncthreads();
    float s[8][4];
    qk_tile<HDP>(Qs, Ks, warp, lane, s);
    uint32_t pa[4][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      float pv[4];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int j = tile * BKV + nt * 8 + (lane & 3) * 2 + e;
        pv[e] = key_ok(j, t_lo) ? expf(s[nt][e] * p.scale - m_lo) * inv_lo : 0.f;
        pv[2 + e] = key_ok(j, t_hi) ? expf(s[nt][2 + e] * p.scale - m_hi) * inv_hi : 0.f;
      }
      pa[nt >> 1][(nt & 1) * 2 + 0] = pack_bf16x2(pv[0], pv[1]);
      pa[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16x2(pv[2], pv[3]);
    }
    const uint32_t v_base = smem_u32(Vs + ((mi & 1) * 8 + ri) * LDS + (mi >> 1) * 8);
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
      for (int dp = 0; dp < HDP / 16; ++dp) {
        uint32_t b0, b1, b2, b3;
        ldsm_x4_t(v_base + (kk * 16 * LDS) * 2 + dp * 32, b0, b1, b2, b3);
        mma_bf16(o[2 * dp], pa[kk][0], pa[kk][1], pa[kk][2], pa[kk][3], b0, b1);
        mma_bf16(o[2 * dp + 1], pa[kk][0], pa[kk][1], pa[kk][2], pa[kk][3], b2, b3);
      }
    }
  }

  // ---------------- store
  const bool ok_lo = r_lo < rows_total, ok_hi = r_hi < rows_total;
  bf16* o_lo = p.out + b * p.o_bs + t_lo * p.o_rs + (kvh * G + r_lo / p.tq) * p.head_dim;
  bf16* o_hi = p.out + b * p.o_bs + t_hi * p.o_rs + (kvh * G + r_hi / p.tq) * p.head_dim;
#pragma unroll
  for (int dt = 0; dt < HDP / 8; ++dt) {
    const int d = dt * 8 + (lane & 3) * 2;
    if (d < p.head_dim) {
      if (ok_lo) *reinterpret_cast<uint32_t*>(o_lo + d) = pack_bf16x2(o[dt][0], o[dt][1]);
      if (ok_hi) *reinterpret_cast<uint32_t*>(o_hi + d) = pack_bf16x2(o[dt][2], o[dt][3]);
    }
  }
}


// ---------------------------------------------------------------------------------------------------
// Fast path: K and V each stream through shared memory ONCE (cp.async, double buffered); the scaled and
// masked logits of the whole row are parked in shared memory in each thread's own mma-fragment order, so
// the exact softmax statistics are known before P is rounded to bf16 and no Q.K^T is recomputed.
// Used whenever the logits fit (keys <= ~384 at head_dim 256, ~900 at head_dim 64).
template <int HDP>
__device__ __forceinline__ void issue_tile(const AttnParams& p, bf16* dst, int tile, bool is_v, int n0, int nk,
                                           int b, int kvb, int kvh) {
  constexpr int LDS = HDP + 8;
  constexpr int CH = HDP / 8;
  for (int idx = threadIdx.x; idx < BKV * CH; idx += ATT_THREADS) {
    const int r = idx / CH, c = (idx % CH) * 8;
    const int j = tile * BKV + r;
    const bool valid = j < nk && c < p.head_dim;
    const bf16* src = p.q;
    if (valid) {
      if (j < n0)
        src = (is_v ? p.v0 : p.k0) + kvb * p.kv0_bs + j * p.kv0_rs + kvh * p.head_dim + c;
      else
        src = (is_v ? p.v1 : p.k1) + b * p.kv1_bs + (j - n0) * p.kv1_rs + kvh * p.head_dim + c;
    }
    cp_async16(smem_u32(dst + r * LDS + c), src, valid);
  }
}

template <int HDP>
__global__ void __launch_bounds__(ATT_THREADS) attn_smem_kernel(const AttnParams p) {
  pdl_wait();
  pdl_launch();
  constexpr int LDS = HDP + 8;
  extern __shared__ __align__(16) uint8_t smem_raw[];
  bf16* Qs = reinterpret_cast<bf16*>(smem_raw);
  bf16* KV = Qs + BQ * LDS;                                   // 2 stages of [64][LDS]
  float* Sp = reinterpret_cast<float*>(KV + 2 * BKV * LDS);   // [tile][32][128] thread-private logits

  const int b = blockIdx.z, kvh = blockIdx.y, qt = blockIdx.x;
  const int G = p.heads / p.kv_heads;
  const int rows_total = G * p.tq;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kvb = b / p.q_per_kv_batch;
  const int n0 = p.kv0_len_dev != nullptr ? p.kv0_len_dev[kvb] : p.kv0_len;
  const int nk = n0 + p.kv1_len;
  const int n_tiles = (nk + BKV - 1) / BKV;

  {  // Q tile + first K tile
    constexpr int CH = HDP / 8;
    for (int idx = threadIdx.x; idx < BQ * CH; idx += ATT_THREADS) {
      const int r = idx / CH, c = (idx % CH) * 8;
      const int i = qt * BQ + r;
      const bool valid = i < rows_total && c < p.head_dim;
      const bf16* src = p.q;
      if (valid) src = p.q + b * p.q_bs + (i % p.tq) * p.q_rs + (kvh * G + i / p.tq) * p.head_dim + c;
      cp_async16(smem_u32(Qs + r * LDS + c), src, valid);
    }
    issue_tile<HDP>(p, KV, 0, false, n0, nk, b, kvb, kvh);
    cp_async_commit();
  }

  const int r_lo = qt * BQ + warp * 16 + (lane >> 2);
  const int r_hi = r_lo + 8;
  const int t_lo = r_lo % p.tq, t_hi = r_hi % p.tq;
  auto key_ok = [&](int j, int t) -> bool {
    if (j >= nk) return false;
    if (p.suffix_mask && j >= n0 && t == 0) return (j - n0) == 0;
    return true;
  };

  float m_lo = -INFINITY, m_hi = -INFINITY, l_lo = 0.f, l_hi = 0.f;
  float o[HDP / 8][4];
#pragma unroll
  for (int dt = 0; dt < HDP / 8; ++dt) o[dt][0] = o[dt][1] = o[dt][2] = o[dt][3] = 0.f;
  float inv_lo = 0.f, inv_hi = 0.f;
  const int mi = lane >> 3, ri = lane & 7;

  // items 0..n_tiles-1 = K tiles (logits), n_tiles..2*n_tiles-1 = V tiles (P @ V)
  for (int it = 0; it < 2 * n_tiles; ++it) {
    const int nxt = it + 1;
    if (nxt < 2 * n_tiles)
      issue_tile<HDP>(p, KV + (nxt & 1) * BKV * LDS, nxt % n_tiles, nxt >= n_tiles, n0, nk, b, kvb, kvh);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    const bf16* buf = KV + (it & 1) * BKV * LDS;
    if (it < n_tiles) {
      const int tile = it;
      float s[8][4];
      qk_tile<HDP>(Qs, buf, warp, lane, s);
      float tm_lo = -INFINITY, tm_hi = -INFINITY;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int j = tile * BKV + nt * 8 + (lane & 3) * 2 + e;
          s[nt][e] = key_ok(j, t_lo) ? s[nt][e] * p.scale : -INFINITY;
          s[nt][2 + e] = key_ok(j, t_hi) ? s[nt][2 + e] * p.scale : -INFINITY;
          tm_lo = fmaxf(tm_lo, s[nt][e]);
          tm_hi = fmaxf(tm_hi, s[nt][2 + e]);
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) Sp[((tile * 8 + nt) * 4 + e) * ATT_THREADS + threadIdx.x] = s[nt][e];
      }
      tm_lo = fmaxf(tm_lo, __shfl_xor_sync(0xffffffffu, tm_lo, 1));
      tm_lo = fmaxf(tm_lo, __shfl_xor_sync(0xffffffffu, tm_lo, 2));
      tm_hi = fmaxf(tm_hi, __shfl_xor_sync(0xffffffffu, tm_hi, 1));
      tm_hi = fmaxf(tm_hi, __shfl_xor_sync(0xffffffffu, tm_hi, 2));
      const float nm_lo = fmaxf(m_lo, tm_lo), nm_hi = fmaxf(m_hi, tm_hi);
      float ts_lo = 0.f, ts_hi = 0.f;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          ts_lo += (s[nt][e] == -INFINITY) ? 0.f : expf(s[nt][e] - nm_lo);
          ts_hi += (s[nt][2 + e] == -INFINITY) ? 0.f : expf(s[nt][2 + e] - nm_hi);
        }
      }
      ts_lo += __shfl_xor_sync(0xffffffffu, ts_lo, 1);
      ts_lo += __shfl_xor_sync(0xffffffffu, ts_lo, 2);
      ts_hi += __shfl_xor_sync(0xffffffffu, ts_hi, 1);
      ts_hi += __shfl_xor_sync(0xffffffffu, ts_hi, 2);
      l_lo = (m_lo == -INFINITY ? 0.f : l_lo * expf(m_lo - nm_lo)) + ts_lo;
      l_hi = (m_hi == -INFINITY ? 0.f : l_hi * expf(m_hi - nm_hi)) + ts_hi;
      m_lo = nm_lo, m_hi = nm_hi;
      if (it == n_tiles - 1) {
        inv_lo = l_lo > 0.f ? 1.0f / l_lo : 0.f;
        inv_hi = l_hi > 0.f ? 1.0f / l_hi : 0.f;
      }
    } else {
      const int tile = it - n_tiles;
      uint32_t pa[4][4];
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        float sv[4], pv[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) sv[e] = Sp[((tile * 8 + nt) * 4 + e) * ATT_THREADS + threadIdx.x];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          pv[e] = sv[e] == -INFINITY ? 0.f : expf(sv[e] - m_lo) * inv_lo;
          pv[2 + e] = sv[2 + e] == -INFINITY ? 0.f : expf(sv[2 + e] - m_hi) * inv_hi;
        }
        pa[nt >> 1][(nt & 1) * 2 + 0] = pack_bf16x2(pv[0], pv[1]);
        pa[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16x2(pv[2], pv[3]);
      }
      const uint32_t v_base = smem_u32(buf + ((mi & 1) * 8 + ri) * LDS + (mi >> 1) * 8);
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
        for (int dp = 0; dp < HDP / 16; ++dp) {
          uint32_t b0, b1, b2, b3;
          ldsm_x4_t(v_base + (kk * 16 * LDS) * 2 + dp * 32, b0, b1, b2, b3);
          mma_bf16(o[2 * dp], pa[kk][0], pa[kk][1], pa[kk][2], pa[kk][3], b0, b1);
          mma_bf16(o[2 * dp + 1], pa[kk][0], pa[kk][1], pa[kk][2], pa[kk][3], b2, b3);
        }
      }
    }
    __syncthreads();  // the buffer just consumed is refilled by the next iteration's prefetch
  }

  const bool ok_lo = r_lo < rows_total, ok_hi = r_hi < rows_total;
  bf16* o_lo = p.out + b * p.o_bs + t_lo * p.o_rs + (kvh * G + r_lo / p.tq) * p.head_dim;
  bf16* o_hi = p.out + b * p.o_bs + t_hi * p.o_rs + (kvh * G + r_hi / p.tq) * p.head_dim;
#pragma unroll
  for (int dt = 0; dt < HDP / 8; ++dt) {
    const int d = dt * 8 + (lane & 3) * 2;
    if (d < p.head_dim) {
      if (ok_lo) *reinterpret_cast<uint32_t*>(o_lo + d) = pack_bf16x2(o[dt][0], o[dt][1]);
      if (ok_hi) *reinterpret_cast<uint32_t*>(o_hi + d) = pack_bf16x2(o[dt][2], o[dt][3]);
    }
  }
}

template <int HDP>
int launch_attn_smem(cudaStream_t st, const AttnParams& p, dim3 grid, int max_tiles) {
  const int smem = (BQ + 2 * BKV) * (HDP + 8) * 2 + max_tiles * 32 * ATT_THREADS * 4;
  auto kern = attn_smem_kernel<HDP>;
  CVB_TRY(ensure_dyn_smem(kern, smem));
  CVB_TRY(launch_pdl(kern, dim3(grid), dim3(ATT_THREADS), smem, st, 1, p));
  CVB_LAUNCHED();
  return 0;
}

template <int HDP>
int launch_attn(cudaStream_t st, const AttnParams& p, dim3 grid) {
  constexpr int SMEM = (BQ + 2 * BKV) * (HDP + 8) * 2;
  auto kern = attn_kernel<HDP>;
  CVB_TRY(ensure_dyn_smem(kern, SMEM));
  CVB_TRY(launch_pdl(kern, dim3(grid), dim3(ATT_THREADS), SMEM, st, 1, p));
  CVB_LAUNCHED();
  return 0;
}

}  // namespace

bool attention_decode_umma_eligible(const AttnCall& c);
int attention_decode_umma(cudaStream_t st, const AttnCall& c);
bool attention_mha_umma_eligible(const AttnCall& c);
int attention_mha_umma(cudaStream_t st, const AttnCall& c);
bool attention_mha_long_umma_eligible(const AttnCall& c);
int attention_mha_long_umma(cudaStream_t st, const AttnCall& c);

int attention(cudaStream_t st, const AttnCall& c) {
  CVB_REQUIRE(c.head_dim % 8 == 0 && c.head_dim <= 256, "head_dim must be a multiple of 8, <= 256");
  CVB_REQUIRE(c.heads % c.kv_heads == 0, "heads must be a multiple of kv_heads");
  // tcgen05 kernels for every shape of the full-size models (denoise step, SigLIP tower, verifier trunk); the exact
  // mma.sync kernels below serve the remaining shapes (other head dims, > 768 keys) - one fallback, no algorithm menu
  if ((c.algo == 0 || c.algo == 3) && attention_decode_umma_eligible(c)) return attention_decode_umma(st, c);
  if ((c.algo == 0 || c.algo == 3) && c.k1 == nullptr && attention_mha_umma_eligible(c)) return attention_mha_umma(st, c);
  if ((c.algo == 0 || c.algo == 3) && c.k1 == nullptr && attention_mha_long_umma_eligible(c)) return attention_mha_long_umma(st, c);
  CVB_REQUIRE(c.algo != 3 && c.q_part == nullptr && c.kv1_cached_k == nullptr && c.kv1_cache_out_k == nullptr,
              "shape not eligible for a tcgen05 attention kernel");
  CVB_REQUIRE(c.rope == nullptr, "fused RoPE needs the cluster decode attention (shape not eligible)");
  CVB_REQUIRE(c.heads % c.kv_heads == 0, "heads must be a multiple of kv_heads");
  CVB_REQUIRE(c.batches > 0 && c.tq > 0, "empty attention");
  AttnParams p;
  p.q = c.q, p.q_bs = c.q_batch_stride, p.q_rs = c.q_row_stride;
  p.k0 = c.k0, p.v0 = c.v0, p.kv0_bs = c.kv0_batch_stride, p.kv0_rs = c.kv0_row_stride;
  p.kv0_len_dev = c.kv0_len_dev, p.kv0_len = c.kv0_len, p.q_per_kv_batch = c.q_per_kv_batch;
  p.k1 = c.k1, p.v1 = c.v1, p.kv1_bs = c.kv1_batch_stride, p.kv1_rs = c.kv1_row_stride;
  p.kv1_len = c.k1 != nullptr ? c.kv1_len : 0;
  p.suffix_mask = c.suffix_mask;
  p.out = c.out, p.o_bs = c.o_batch_stride, p.o_rs = c.o_row_stride;
  p.heads = c.heads, p.kv_heads = c.kv_heads, p.tq = c.tq, p.head_dim = c.head_dim;
  p.scale = c.scale;
  const int G = c.heads / c.kv_heads;
  dim3 grid((G * c.tq + BQ - 1) / BQ, c.kv_heads, c.batches);
  // logits-in-smem fast path when the whole key range fits next to the Q / K / V tiles
  const int hdp = c.head_dim <= 64 ? 64 : c.head_dim <= 80 ? 80 : c.head_dim <= 128 ? 128 : 256;
  const int max_keys = (c.kv0_len_dev != nullptr ? c.kv0_max : c.kv0_len) + p.kv1_len;
  const int max_tiles = (max_keys + BKV - 1) / BKV;
  const long smem_fast = (long)(BQ + 2 * BKV) * (hdp + 8) * 2 + (long)max_tiles * 32 * ATT_THREADS * 4;
  if (c.kv0_len_dev != nullptr) CVB_REQUIRE(c.kv0_max > 0, "kv0_max (upper bound of the device-side length) is required");
  if (smem_fast <= 220 * 1024 && !c.force_two_pass) {
    if (hdp == 64) return launch_attn_smem<64>(st, p, grid, max_tiles);
    if (hdp == 80) return launch_attn_smem<80>(st, p, grid, max_tiles);
    if (hdp == 128) return launch_attn_smem<128>(st, p, grid, max_tiles);
    return launch_attn_smem<256>(st, p, grid, max_tiles);
  }
  if (hdp == 64) return launch_attn<64>(st, p, grid);
  if (hdp == 80) return launch_attn<80>(st, p, grid);
  if (hdp == 128) return launch_attn<128>(st, p, grid);
  return launch_attn<256>(st, p, grid);
}

}  // namespace cvb
