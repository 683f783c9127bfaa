// Prefix self-attention on the 5th-generation tensor cores (tcgen05 / TMEM), for the PaliGemma prefix pass of
// PI0FlowMatching.sample_actions (modeling_pi0.py:672-695 -> PaliGemmaWithExpertModel.forward,
// paligemma_with_expert.py:236-360 -> eager_attention_forward :376-434): multi-query attention, 8 query heads sharing
// one 256-wide KV head, every prefix token attends every valid prefix token (right-padded language tokens masked).
//
// One CTA = 16 query tokens x 8 heads = 128 UMMA rows of one prompt (the heads are folded into rows by a 3-D TMA box,
// so K / V are read once for all heads):
//   1. TMA: Q tile (4 hd-blocks of [128 rows x 64]) and ALL keys of the prompt (4 hd-blocks of [tk_pad x 64]), 128-byte
//      swizzle, one mbarrier per hd-block so the first MMAs start while the rest is in flight;
//   2. S[128 x tk_pad] = Q K^T by tcgen05.mma (128 x N x 16, fp32 accumulator in TMEM, N = 256 + remainder);
//   3. softmax straight out of TMEM, one row per thread (exact max / sum over all keys, fp32, masked keys dropped);
//      the NORMALISED probabilities are rounded to bf16 (the reference's ledger) and written into shared memory in the
//      canonical K-major 128-byte-swizzled layout, i.e. as the "A" operand of the second GEMM, on top of the dead Q / K;
//   4. O[128 x 256] = P V by tcgen05.mma with V^T tiles ([256 hd x 64 keys], written transposed by the RoPE / KV-cache
//      kernel so that both operands stay K-major) streamed by TMA through a 4-slot ring while the softmax runs;
//   5. O: TMEM -> bf16 -> global (one 512-byte row per thread).
// Replaces the mma.sync kernel of ops_attention.cu for this shape (measured 140 us per layer, bound by the legacy
// mma.sync issue rate; profiles/r1b_launches_by_kernel.txt).
#include <mutex>
#include <unordered_map>

#include "host_common.h"
#include "ops.h"
#include "ptx.cuh"

namespace cvb {

int get_tmap_cached(const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows, CUtensorMap* out);
int make_tmap_3d_heads(CUtensorMap* out, const void* ptr, uint64_t rows, uint64_t heads, uint64_t hd, uint64_t ld,
                       uint32_t box_rows, uint32_t box_heads);

extern unsigned long long* g_skinny_ts;

namespace {

constexpr int UA_THREADS = 320;  // warp 0: TMA producer, warp 1: MMA issuer + TMEM owner, warps 2..9: softmax / epilogue
constexpr int UA_SOFT = 256;     // two warps per TMEM lane quarter, even / odd 16-column chunks
constexpr int UA_HD = 256;
constexpr int UA_TOK = 16;     // query tokens per CTA
constexpr int UA_HEADS = 8;    // query heads folded into rows: 16 x 8 = 128 UMMA rows
constexpr int UA_QBLK = 128 * 128;      // one hd-block of the Q tile: 128 rows x 64 bf16
constexpr int UA_VBLK = UA_HD * 128;    // one key-block of V^T: 256 rows x 64 keys
constexpr int UA_PBLK = 128 * 128;      // one key-block of P: 128 rows x 64 keys
constexpr int UA_VSLOTS = 4;
constexpr int UA_MAX_KEYS = 384;
constexpr int UA_BODY_MAX = 222 * 1024;  // operand bytes per CTA (barriers and the alignment slack come on top)

struct UmmaAttnParams {
  const int* klen_dev;
  int klen;
  int tq;
  int tk_pad;
  long q_rows_per_batch;
  long k_rows_per_batch;
  bf16* out;
  long o_bs, o_rs;
  float scale;
};

__device__ __forceinline__ uint32_t make_idesc_n(int n) {  // bf16 x bf16 -> fp32, M = 128, run-time N
  return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(128 >> 4) << 24);
}

__device__ __forceinline__ void sts_u4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void named_bar(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

// Softmax building blocks on one 16-column chunk of a TMEM row whose first nv (1..16) columns are valid keys.
// exp(x * scale - m * scale) is evaluated as 2^(x * c2 - m * c2), c2 = scale * log2(e): one FFMA + one ex2 per element.
// (1) online update of the running (raw max, sum of exponentials relative to it)
__device__ __forceinline__ void online_chunk(const uint32_t (&rr)[16], int nv, float c2, float& m, float& l) {
  float cm = -INFINITY;
  if (nv == 16) {
#pragma unroll
    for (int i = 0; i < 16; ++i) cm = fmaxf(cm, __uint_as_float(rr[i]));
  } else {
#pragma unroll
    for (int i = 0; i < 16; ++i)
      if (i < nv) cm = fmaxf(cm, __uint_as_float(rr[i]));
  }
  const float mn = fmaxf(m, cm);
  const float mc = mn * c2;
  float acc = 0.f;
  if (nv == 16) {
#pragma unroll
    for (int i = 0; i < 16; ++i) acc += ex2_approx(fmaf(__uint_as_float(rr[i]), c2, -mc));
  } else {
#pragma unroll
    for (int i = 0; i < 16; ++i)
      if (i < nv) acc += ex2_approx(fmaf(__uint_as_float(rr[i]), c2, -mc));
  }
  l = l * ex2_approx((m - mn) * c2) + acc;  // first chunk: m = -inf -> factor 0
  m = mn;
}
// (1b) same, and the exponentials (relative to the NEW running max, 0 for invalid columns) replace rr
// (valid columns: [lo, lo + nv).  The sum walks the columns in order and the invalid ones add exact zeros, so a row's
// result does not depend on `lo` - a candidate's suffix keys may sit anywhere in the 16-wide suffix tile)
__device__ __forceinline__ void online_chunk_keep(uint32_t (&rr)[16], int nv, float c2, float& m, float& l, int lo = 0) {
  float cm = -INFINITY;
#pragma unroll
  for (int i = 0; i < 16; ++i)
    if (i >= lo && i < lo + nv) cm = fmaxf(cm, __uint_as_float(rr[i]));
  const float mn = fmaxf(m, cm);
  const float mc = mn * c2;
  float acc = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const float e = (i >= lo && i < lo + nv) ? ex2_approx(fmaf(__uint_as_float(rr[i]), c2, -mc)) : 0.f;
    acc += e;
    rr[i] = __float_as_uint(e);
  }
  l = l * ex2_approx((m - mn) * c2) + acc;
  m = mn;
}
// (2) normalised probabilities, rounded to bf16 and packed in key order (invalid columns -> 0)
__device__ __forceinline__ void prob_chunk(const uint32_t (&rr)[16], int nv, float c2, float Mc, float inv, uint32_t (&pk)[8]) {
  if (nv == 16) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      pk[i] = pack_bf16x2(ex2_approx(fmaf(__uint_as_float(rr[2 * i]), c2, -Mc)) * inv,
                          ex2_approx(fmaf(__uint_as_float(rr[2 * i + 1]), c2, -Mc)) * inv);
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float a = 2 * i < nv ? ex2_approx(fmaf(__uint_as_float(rr[2 * i]), c2, -Mc)) * inv : 0.f;
      const float bb = 2 * i + 1 < nv ? ex2_approx(fmaf(__uint_as_float(rr[2 * i + 1]), c2, -Mc)) * inv : 0.f;
      pk[i] = pack_bf16x2(a, bb);
    }
  }
}

__global__ void __launch_bounds__(UA_THREADS, 1)
attn_prefix_umma_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                        const __grid_constant__ CUtensorMap tmVT, const UmmaAttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  const int tk_pad = p.tk_pad;
  const uint32_t kblk = static_cast<uint32_t>(tk_pad) * 128u;  // one hd-block of K: tk_pad rows x 64 bf16
  const int nkb = (tk_pad + 63) / 64;                          // key blocks of P / V^T
  const int nslots = min(min(nkb, UA_VSLOTS), (UA_BODY_MAX - nkb * UA_PBLK) / UA_VBLK);
  // all four hd-blocks of K stay resident when they fit; otherwise the last one reuses the slot of the first once its
  // MMAs have retired (only for > 320 keys)
  const int kslots = 4u * UA_QBLK + 4u * kblk <= static_cast<uint32_t>(UA_BODY_MAX) ? 4 : 3;
  const uint32_t qk_bytes = 4u * UA_QBLK + static_cast<uint32_t>(kslots) * kblk;
  const uint32_t pv_bytes = static_cast<uint32_t>(nkb) * UA_PBLK + static_cast<uint32_t>(nslots) * UA_VBLK;
  const uint32_t body = max(qk_bytes, pv_bytes);
  uint8_t* sQ = smem;
  uint8_t* sK = smem + 4 * UA_QBLK;
  uint8_t* sP = smem;                                     // overlays Q / K once S is complete
  uint8_t* sV = smem + static_cast<uint32_t>(nkb) * UA_PBLK;  // overlays K once S is complete
  uint64_t* qk_full = reinterpret_cast<uint64_t*>(smem + body);
  uint64_t* s_full = qk_full + 4;
  uint64_t* p_ready = s_full + 1;
  uint64_t* v_full = p_ready + 1;
  uint64_t* v_empty = v_full + UA_VSLOTS;
  uint64_t* o_full = v_empty + UA_VSLOTS;
  uint64_t* k0_free = o_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(k0_free + 1);
  float2* red2 = reinterpret_cast<float2*>(k0_free + 3);  // [2][128] (max, sum) of the two column halves

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = blockIdx.x, b = blockIdx.y;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmVT);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int c = 0; c < 4; ++c) mbar_init(&qk_full[c], 1);
      mbar_init(s_full, 1);
      mbar_init(p_ready, UA_SOFT);
      for (int s = 0; s < UA_VSLOTS; ++s) {
        mbar_init(&v_full[s], 1);
        mbar_init(&v_empty[s], 1);
      }
      mbar_init(o_full, 1);
      mbar_init(k0_free, 1);
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
  pdl_launch();

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      pdl_wait();  // Q, K and V^T are all written by the preceding kernels
      const int q_row0 = static_cast<int>(b * p.q_rows_per_batch) + tile * UA_TOK;
      const int k_row0 = static_cast<int>(b * p.k_rows_per_batch);
      const int half_rows = tk_pad / 2;
      for (int c = 0; c < 4; ++c) {
        uint8_t* kdst = sK + (c % kslots) * kblk;
        if (c >= kslots) mbar_wait(k0_free, 0);
        mbar_arrive_expect_tx(&qk_full[c], UA_QBLK + kblk);
        tma_load_3d(sQ + c * UA_QBLK, &tmQ, &qk_full[c], c * 64, 0, q_row0);
        tma_load_2d(kdst, &tmK, &qk_full[c], c * 64, k_row0);
        tma_load_2d(kdst + static_cast<uint32_t>(half_rows) * 128u, &tmK, &qk_full[c], c * 64, k_row0 + half_rows);
      }
      mbar_wait(s_full, 0);  // every Q.K^T MMA has retired: Q / K are dead, V^T may land on top of K
      for (int kb = 0; kb < nkb; ++kb) {
        const int slot = kb % nslots;
        if (kb >= nslots) mbar_wait(&v_empty[slot], ((kb / nslots) - 1) & 1);
        mbar_arrive_expect_tx(&v_full[slot], UA_VBLK);
        tma_load_2d(sV + slot * UA_VBLK, &tmVT, &v_full[slot], kb * 64, b * UA_HD);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      const int n1 = min(tk_pad, 256), n2 = tk_pad - n1;
      const uint32_t idesc1 = make_idesc_n(n1), idesc2 = make_idesc_n(n2);
      for (int c = 0; c < 4; ++c) {
        mbar_wait(&qk_full[c], 0);
        tc_fence_after();
        const uint64_t qd = make_desc_kmajor_sw128(smem_u32(sQ + c * UA_QBLK));
        const uint32_t k_addr = smem_u32(sK + (c % kslots) * kblk);
        const uint64_t kd = make_desc_kmajor_sw128(k_addr);
        const uint64_t kd2 = make_desc_kmajor_sw128(k_addr + 256u * 128u);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          umma_bf16(tmem_base, qd + 2 * k, kd + 2 * k, idesc1, (c | k) != 0 ? 1u : 0u);
          if (n2 > 0) umma_bf16(tmem_base + 256, qd + 2 * k, kd2 + 2 * k, idesc2, (c | k) != 0 ? 1u : 0u);
        }
        if (c == 0 && kslots < 4) umma_commit(k0_free);
      }
      umma_commit(s_full);
      mbar_wait(p_ready, 0);  // P is in shared memory (and S has been consumed: O may overwrite its columns)
      tc_fence_after();
      const uint32_t idesc_o = make_idesc_n(UA_HD);
      for (int kb = 0; kb < nkb; ++kb) {
        const int slot = kb % nslots;
        mbar_wait(&v_full[slot], (kb / nslots) & 1);
        tc_fence_after();
        const uint64_t pd = make_desc_kmajor_sw128(smem_u32(sP + kb * UA_PBLK));
        const uint64_t vd = make_desc_kmajor_sw128(smem_u32(sV + slot * UA_VBLK));
        const int ksteps = min(4, (tk_pad - kb * 64) / 16);
        for (int k = 0; k < ksteps; ++k) umma_bf16(tmem_base, pd + 2 * k, vd + 2 * k, idesc_o, (kb | k) != 0 ? 1u : 0u);
        umma_commit(&v_empty[slot]);
      }
      umma_commit(o_full);
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ softmax + epilogue
    // Row = TMEM lane; the two warps of a lane quarter take the even / odd 16-column chunks (a fixed assignment, so the
    // result does not depend on how far the key range is padded).
    pdl_wait();
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    const int tl = row >> 3, h = row & 7;
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    int n_keys = p.klen_dev != nullptr ? p.klen_dev[b] : p.klen;
    n_keys = max(1, min(n_keys, tk_pad));
    const float c2 = p.scale * 1.4426950408889634f;
    mbar_wait(s_full, 0);
    tc_fence_after();
    float m = -INFINITY, l = 0.f;
    for (int ch = half; ch * 16 < n_keys; ch += 2) {
      uint32_t rr[16];
      tmem_ld_x16(taddr + ch * 16, rr);
      tmem_wait_ld();
      online_chunk(rr, min(16, n_keys - ch * 16), c2, m, l);
    }
    red2[half * 128 + row] = make_float2(m, l);
    named_bar(1, UA_SOFT);
    const float2 s0 = red2[row], s1 = red2[128 + row];
    const float M = fmaxf(s0.x, s1.x);
    const float Mc = M * c2;
    const float Ls = s0.y * ex2_approx((s0.x - M) * c2) + s1.y * ex2_approx((s1.x - M) * c2);  // an empty half has (-inf, 0)
    const float inv = 1.0f / Ls;
    // P[row][key] -> canonical K-major SWIZZLE_128B tile: 8-row groups of 1024 B, 16-byte chunk index XOR (row % 8)
    const uint32_t p_row = smem_u32(sP) + static_cast<uint32_t>(row >> 3) * 1024u + static_cast<uint32_t>(row & 7) * 128u;
    for (int ch = half; ch * 16 < tk_pad; ch += 2) {
      const int nv = max(0, min(16, n_keys - ch * 16));
      uint32_t pk[8];
      if (nv > 0) {  // warp-uniform
        uint32_t rr[16];
        tmem_ld_x16(taddr + ch * 16, rr);
        tmem_wait_ld();
        prob_chunk(rr, nv, c2, Mc, inv, pk);
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) pk[i] = 0u;
      }
      const int c0 = ch * 16;
      const uint32_t blk = p_row + static_cast<uint32_t>(c0 >> 6) * UA_PBLK;
      const int chunk = (c0 & 63) >> 3;
      sts_u4(blk + (static_cast<uint32_t>(chunk ^ (row & 7)) << 4), pk[0], pk[1], pk[2], pk[3]);
      sts_u4(blk + (static_cast<uint32_t>((chunk + 1) ^ (row & 7)) << 4), pk[4], pk[5], pk[6], pk[7]);
    }
    fence_proxy_async();  // generic-proxy writes of P -> visible to the tensor core's async-proxy reads
    tc_fence_before();
    mbar_arrive(p_ready);

    mbar_wait(o_full, 0);
    tc_fence_after();
    const int t = tile * UA_TOK + tl;
    bf16* op = p.out + b * p.o_bs + static_cast<long>(t) * p.o_rs + h * UA_HD + half * 128;
    for (int c0 = 0; c0 < 128; c0 += 32) {
      uint32_t r[32];
      tmem_ld_x32(taddr + half * 128 + c0, r);
      tmem_wait_ld();
      if (t < p.tq) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint32_t o[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) o[e] = pack_bf16x2(__uint_as_float(r[8 * j + 2 * e]), __uint_as_float(r[8 * j + 2 * e + 1]));
          *reinterpret_cast<uint4*>(op + c0 + 8 * j) = make_uint4(o[0], o[1], o[2], o[3]);
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Denoise-step attention on tcgen05 (PI0FlowMatching.denoise_step, modeling_pi0.py:717-752 -> eager_attention_forward,
// paligemma_with_expert.py:376-434): one CTA per candidate, its heads x suffix-token query rows (8 x 5 = 40) attend the
// rephrase's prefix KV cache (<= 336 keys, static during the loop: K tiles are prefetched by TMA BEFORE the
// programmatic-dependency wait) plus the candidate's own suffix keys (pi0 suffix mask).
//   * Q rows and the suffix keys are loaded from the fused qkv buffer, rotated (RoPE table, same arithmetic as
//     rope_kernel) and written straight into K-major 128-byte-swizzled UMMA tiles; query row r sits in TMEM lane
//     (r % 4) * 32 + r / 4 so that all four lane quarters - and therefore 16 softmax warps - share the rows;
//   * S = Q K^T: tcgen05.mma N = 256 + remainder (prefix keys) and N = 16 (suffix keys) into one TMEM accumulator;
//   * exact softmax (global max / sum over all keys, probabilities normalised THEN rounded to bf16) by 4 warps per
//     lane quarter, 1/4 of the columns each, combined through shared memory in fixed order (deterministic);
//   * O = P V for the prefix keys on tcgen05 with V^T tiles (transposed cache written by the prefix RoPE kernel); the
//     <= 8 suffix keys are added in fp32 by the epilogue threads (5 x 256 MACs per row - not worth a tile).
// Replaces the cluster split-KV mma.sync kernel (ops_attention_decode.cu: 20 us, bound by the mma.sync issue rate and
// the DSMEM reduce-scatter; tools/decode_ts.py).
constexpr int UD_THREADS = 576;   // warp 0 TMA, warp 1 MMA, warps 2..17 staging / softmax / epilogue
constexpr int UD_SOFT = 512;
constexpr int UD_VSLOTS = 8;         // V^T slots (mbarrier pairs); every key-block is resident when the tiles are halved
constexpr int UD_KSBLK = 16 * 128;  // suffix-key tile of one hd-block: 16 rows x 64 bf16
constexpr int UD_MISC = 14 * 1024;  // red2[4][128] (max, sum) | psuf[128][8] bf16 | v1[15][256] bf16 (<= 3 candidates x 5 keys)
constexpr int UD_OS_LD = 260;       // fp32 row stride of the output staging tile (bank spread, 16-byte aligned)
__host__ __device__ inline int ud_misc_off(int nkb, int vslots, int rows_total, int hdw) {
  const int pv = nkb * UA_PBLK + vslots * hdw * 128;
  const int os = (rows_total * UD_OS_LD * 4 + 1023) / 1024 * 1024;
  return pv > os ? pv : os;
}

struct UmmaDecodeParams {
  const bf16* q;
  long q_bs, q_rs;
  const int* kv0_len_dev;
  int kv0_len;
  int q_per_kv_batch;
  long k_rows_per_batch;
  const bf16* k1;
  const bf16* v1;
  long kv1_bs, kv1_rs;
  int kv1_len;
  int suffix_mask;
  bf16* out;
  long o_bs, o_rs;
  int heads, tq, tk_pad;
  int hdw;  // output columns per CTA: 256, or 128 with gridDim.y = 2 (each CTA repeats Q.K^T / softmax, halves P.V)
  float scale;
  int kv0_static;
  const float2* rope;  // [kv batches][tq][128] (cos, sin) or nullptr
  unsigned long long* ts;  // diagnostics (cvb_debug_set_timestamps) or nullptr
  // optional: the fused qkv projection as fp32 split-K partials [part_S][rows][same strides as q / k1 / v1] (written by
  // the persistent expert kernel, expert_mega.cuh): summed in split order and rounded to bf16 while staging
  const float* qf;
  const float* k1f;
  const float* v1f;
  int part_S;
  long part_ss;
  // F7 hoist: the first `kc` suffix keys / values (the state token) are not recomputed every denoise step -
  // they come from kc_k / kc_v ([batches][kc][256] bf16, K already rotated); k1 / v1 then hold only kv1_len - kc rows.
  // kout_k / kout_v (step 0): the rotated key / the value of suffix key 0 are also written there.
  const bf16* kc_k;
  const bf16* kc_v;
  int kc;
  bf16* kout_k;
  bf16* kout_v;
  int rope_rows, rope_off;  // rows of the rope table per kv batch, and the table row of query token 0 / new key 0
  // Candidate grouping (more candidates than SMs, i.e. several observations per call): one CTA takes `group` consecutive
  // candidates of the SAME kv batch (rephrase) - their query rows share the 128 UMMA rows and ONE copy of the prefix K /
  // V^T, their suffix keys share the 16-wide suffix tile behind a block-diagonal mask.  group <= 1: one candidate per CTA.
  int group;
  int rows_max;  // group * heads * tq: sizes the shared-memory layout (the last group of a kv batch may be smaller)
};

// 8 consecutive bf16 of the qkv projection at element offset `off`: plain load, or sum of the fp32 partials -> bf16
__device__ __forceinline__ uint4 ud_load8(const bf16* base, const float* part, long off, int S, long ss) {
  if (part == nullptr) return *reinterpret_cast<const uint4*>(base + off);
  float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
  int s = 0;
  for (; s + 4 <= S; s += 4) {  // eight loads in flight, summed in split order
    float4 x[4], y[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      x[j] = __ldcg(reinterpret_cast<const float4*>(part + (s + j) * ss + off));
      y[j] = __ldcg(reinterpret_cast<const float4*>(part + (s + j) * ss + off + 4));
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      a.x += x[j].x, a.y += x[j].y, a.z += x[j].z, a.w += x[j].w;
      b.x += y[j].x, b.y += y[j].y, b.z += y[j].z, b.w += y[j].w;
    }
  }
  for (; s < S; ++s) {
    const float4 x = __ldcg(reinterpret_cast<const float4*>(part + s * ss + off));
    const float4 y = __ldcg(reinterpret_cast<const float4*>(part + s * ss + off + 4));
    a.x += x.x, a.y += x.y, a.z += x.z, a.w += x.w;
    b.x += y.x, b.y += y.y, b.z += y.z, b.w += y.w;
  }
  return make_uint4(pack_bf16x2(a.x, a.y), pack_bf16x2(a.z, a.w), pack_bf16x2(b.x, b.y), pack_bf16x2(b.z, b.w));
}

__device__ __forceinline__ unsigned long long ud_timer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define UD_TS(i)                                                                            \
  do {                                                                                      \
    if (p.ts != nullptr && threadIdx.x == 64) p.ts[(blockIdx.y * gridDim.x + blockIdx.x) * 8 + (i)] = ud_timer();       \
  } while (0)

// rotate 8 (x1, x2) pairs exactly like rope_kernel / the cluster decode kernel (separate mul / add, no contraction)
__device__ __forceinline__ void rope8_umma(uint4& v1, uint4& v2, const float2* cs) {
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

template <int UD_IPT>  // output items (8 columns of one row) per thread: ceil(rows * 32 / 512)
__global__ void __launch_bounds__(UD_THREADS, 1)
attn_decode_umma_kernel(const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmVT,
                        const UmmaDecodeParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  const int tk_pad = p.tk_pad;
  // S columns: tk_pad prefix keys, then the 16-wide suffix-key tile
  const uint32_t kblk = static_cast<uint32_t>(tk_pad) * 128u;
  const int nkb = (tk_pad + 63) / 64;
  const int kslots = 4u * UA_QBLK + 4u * UD_KSBLK + 4u * kblk <= static_cast<uint32_t>(UA_BODY_MAX) ? 4 : 3;
  const int hdw = p.hdw;                                   // output columns of this CTA (256, or 128 when two CTAs share a candidate)
  const int hy = blockIdx.y;
  const uint32_t vblk = static_cast<uint32_t>(hdw) * 128u;  // one key-block of V^T: hdw rows x 64 keys
  const int vslots = min(min(nkb, UD_VSLOTS), (UA_BODY_MAX - UD_MISC - nkb * UA_PBLK) / static_cast<int>(vblk));
  const uint32_t qk_bytes = 4u * UA_QBLK + 4u * UD_KSBLK + static_cast<uint32_t>(kslots) * kblk;
  // which candidates: b .. b + G - 1 (all of kv batch kvb)
  int b, kvb, G = 1;
  if (p.group > 1) {
    const int ngr = (p.q_per_kv_batch + p.group - 1) / p.group;
    kvb = blockIdx.x / ngr;
    const int c0 = (blockIdx.x - kvb * ngr) * p.group;
    G = min(p.group, p.q_per_kv_batch - c0);
    b = kvb * p.q_per_kv_batch + c0;
  } else {
    b = blockIdx.x;
    kvb = b / p.q_per_kv_batch;
  }
  const int rows_cand = p.heads * p.tq;
  const int rows_total = G * rows_cand;
  const uint32_t misc_off = static_cast<uint32_t>(ud_misc_off(nkb, vslots, p.rows_max, hdw));
  const uint32_t body = max(qk_bytes, misc_off + UD_MISC);
  uint8_t* sQ = smem;
  uint8_t* sKs = smem + 4 * UA_QBLK;
  uint8_t* sK = sKs + 4 * UD_KSBLK;
  uint8_t* sP = smem;
  uint8_t* sV = smem + static_cast<uint32_t>(nkb) * UA_PBLK;
  uint8_t* misc = smem + misc_off;
  float2* red2 = reinterpret_cast<float2*>(misc);              // [4][128] (max, sum) per column group
  bf16* psuf = reinterpret_cast<bf16*>(misc + 4096);           // [128][8]
  bf16* sV1 = reinterpret_cast<bf16*>(misc + 6144);            // [8][256]
  uint64_t* k_full = reinterpret_cast<uint64_t*>(smem + body);
  uint64_t* q_ready = k_full + 4;
  uint64_t* s_full = q_ready + 1;
  uint64_t* p_ready = s_full + 1;
  uint64_t* v_full = p_ready + 1;
  uint64_t* v_empty = v_full + UD_VSLOTS;
  uint64_t* o_full = v_empty + UD_VSLOTS;
  uint64_t* k0_free = o_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(k0_free + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmVT);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int c = 0; c < 4; ++c) mbar_init(&k_full[c], 1);
      mbar_init(q_ready, UD_SOFT);
      mbar_init(s_full, 1);
      mbar_init(p_ready, UD_SOFT);
      for (int s = 0; s < UD_VSLOTS; ++s) {
        mbar_init(&v_full[s], 1);
        mbar_init(&v_empty[s], 1);
      }
      mbar_init(o_full, 1);
      mbar_init(k0_free, 1);
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
  pdl_launch();

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      if (!p.kv0_static) pdl_wait();
      const int k_row0 = static_cast<int>(kvb * p.k_rows_per_batch);
      const int half_rows = tk_pad / 2;
      for (int c = 0; c < 4; ++c) {
        uint8_t* kdst = sK + (c % kslots) * kblk;
        if (c >= kslots) mbar_wait(k0_free, 0);
        mbar_arrive_expect_tx(&k_full[c], kblk);
        tma_load_2d_hint(kdst, &tmK, &k_full[c], c * 64, k_row0, kEvictLast);
        tma_load_2d_hint(kdst + static_cast<uint32_t>(half_rows) * 128u, &tmK, &k_full[c], c * 64, k_row0 + half_rows, kEvictLast);
      }
      mbar_wait(s_full, 0);
      for (int kb = 0; kb < nkb; ++kb) {
        const int slot = kb % vslots;
        if (kb >= vslots) mbar_wait(&v_empty[slot], ((kb / vslots) - 1) & 1);
        mbar_arrive_expect_tx(&v_full[slot], vblk);
        tma_load_2d_hint(sV + slot * vblk, &tmVT, &v_full[slot], kb * 64, kvb * UA_HD + hy * hdw, kEvictLast);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      const int n1 = min(tk_pad, 256), n2 = tk_pad - n1;
      const uint32_t idesc1 = make_idesc_n(n1), idesc2 = make_idesc_n(n2), idesc3 = make_idesc_n(16);
      mbar_wait(q_ready, 0);
      tc_fence_after();
      for (int c = 0; c < 4; ++c) {
        mbar_wait(&k_full[c], 0);
        tc_fence_after();
        const uint64_t qd = make_desc_kmajor_sw128(smem_u32(sQ + c * UA_QBLK));
        const uint32_t k_addr = smem_u32(sK + (c % kslots) * kblk);
        const uint64_t kd = make_desc_kmajor_sw128(k_addr);
        const uint64_t kd2 = make_desc_kmajor_sw128(k_addr + 256u * 128u);
        const uint64_t ksd = make_desc_kmajor_sw128(smem_u32(sKs + c * UD_KSBLK));
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint32_t acc = (c | k) != 0 ? 1u : 0u;
          umma_bf16(tmem_base, qd + 2 * k, kd + 2 * k, idesc1, acc);
          if (n2 > 0) umma_bf16(tmem_base + 256, qd + 2 * k, kd2 + 2 * k, idesc2, acc);
          umma_bf16(tmem_base + tk_pad, qd + 2 * k, ksd + 2 * k, idesc3, acc);
        }
        if (c == 0 && kslots < 4) umma_commit(k0_free);
      }
      umma_commit(s_full);
      mbar_wait(p_ready, 0);
      tc_fence_after();
      const uint32_t idesc_o = make_idesc_n(hdw);
      for (int kb = 0; kb < nkb; ++kb) {
        const int slot = kb % vslots;
        mbar_wait(&v_full[slot], (kb / vslots) & 1);
        tc_fence_after();
        const uint64_t pd = make_desc_kmajor_sw128(smem_u32(sP + kb * UA_PBLK));
        const uint64_t vd = make_desc_kmajor_sw128(smem_u32(sV + slot * vblk));
        const int ksteps = min(4, (tk_pad - kb * 64) / 16);
        for (int k = 0; k < ksteps; ++k) umma_bf16(tmem_base, pd + 2 * k, vd + 2 * k, idesc_o, (kb | k) != 0 ? 1u : 0u);
        umma_commit(&v_empty[slot]);
      }
      umma_commit(o_full);
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ staging, softmax, epilogue (512 threads)
    const int sid = threadIdx.x - 64;
    const int q = warp & 3;          // TMEM lane quarter this warp may access
    const int g = (warp - 2) >> 2;   // column group 0..3
    UD_TS(0);
    pdl_wait();
    UD_TS(1);
    const int n0 = min(p.kv0_len_dev != nullptr ? p.kv0_len_dev[kvb] : p.kv0_len, tk_pad);
    const float2* rope = p.rope != nullptr ? p.rope + (static_cast<long>(kvb) * p.rope_rows + p.rope_off) * 128 : nullptr;
    // ---- Q rows and suffix keys: global -> registers -> RoPE -> swizzled UMMA tiles
    const int nkeys = G * p.kv1_len;  // suffix keys of the CTA's candidates: candidate gi owns tile rows [gi kv1_len, ..)
    const int n_items = (rows_total + nkeys) * 16;
    for (int item = sid; item < n_items; item += UD_SOFT) {
      const bool isq = item < rows_total * 16;
      const int r = isq ? item >> 4 : (item - rows_total * 16) >> 4;  // query row / suffix-key row of the CTA
      const int c = item & 15;  // 8-wide chunk of the first half of the head
      int gi, t, rk = 0;        // candidate within the group, token within the candidate, key within the candidate
      if (isq) {
        const int tt = r / p.heads;
        gi = tt / p.tq, t = tt - gi * p.tq;
      } else {
        gi = r / p.kv1_len, rk = r - gi * p.kv1_len, t = rk - p.kc;
      }
      const bool cached = !isq && rk < p.kc;  // a hoisted suffix key: already rotated, read from the cache
      const long bb = b + gi;
      const long soff = isq ? bb * p.q_bs + t * p.q_rs + (r % p.heads) * UA_HD + c * 8 : bb * p.kv1_bs + t * p.kv1_rs + c * 8;
      uint4 x1, x2;
      if (cached) {
        const bf16* kp = p.kc_k + (bb * p.kc + rk) * UA_HD + c * 8;
        x1 = *reinterpret_cast<const uint4*>(kp);
        x2 = *reinterpret_cast<const uint4*>(kp + 128);
      } else {
        x1 = ud_load8(isq ? p.q : p.k1, isq ? p.qf : p.k1f, soff, p.part_S, p.part_ss);
        x2 = ud_load8(isq ? p.q : p.k1, isq ? p.qf : p.k1f, soff + 128, p.part_S, p.part_ss);
      }
      if (rope != nullptr && !cached) {
        float4 cs[4];
        const float4* tp = reinterpret_cast<const float4*>(rope + t * 128 + c * 8);
#pragma unroll
        for (int e = 0; e < 4; ++e) cs[e] = tp[e];
        rope8_umma(x1, x2, reinterpret_cast<const float2*>(cs));
      }
      if (p.kout_k != nullptr && !isq && rk == 0 && hy == 0) {  // step 0: keep the state token's rotated key
        bf16* kp = p.kout_k + bb * UA_HD + c * 8;
        *reinterpret_cast<uint4*>(kp) = x1;
        *reinterpret_cast<uint4*>(kp + 128) = x2;
      }
      const int row = isq ? (r & 3) * 32 + (r >> 2) : r;
      const uint32_t tile = isq ? smem_u32(sQ) + static_cast<uint32_t>(c >> 3) * UA_QBLK
                                : smem_u32(sKs) + static_cast<uint32_t>(c >> 3) * UD_KSBLK;
      const uint32_t tile2 = tile + 2u * (isq ? UA_QBLK : UD_KSBLK);  // columns + 128 live two hd-blocks further
      const uint32_t off = static_cast<uint32_t>(row >> 3) * 1024u + static_cast<uint32_t>(row & 7) * 128u +
                           (static_cast<uint32_t>((c & 7) ^ (row & 7)) << 4);
      sts_u4(tile + off, x1.x, x1.y, x1.z, x1.w);
      sts_u4(tile2 + off, x2.x, x2.y, x2.z, x2.w);
    }
    uint4 v1reg = make_uint4(0, 0, 0, 0);
    if (sid < nkeys * 32) {
      const int gi = (sid >> 5) / p.kv1_len, j = (sid >> 5) - gi * p.kv1_len;
      const long bb = b + gi;
      if (j < p.kc)
        v1reg = *reinterpret_cast<const uint4*>(p.kc_v + (bb * p.kc + j) * UA_HD + (sid & 31) * 8);
      else
        v1reg = ud_load8(p.v1, p.v1f, bb * p.kv1_bs + (j - p.kc) * p.kv1_rs + (sid & 31) * 8, p.part_S, p.part_ss);
      if (p.kout_v != nullptr && j == 0 && hy == 0)
        *reinterpret_cast<uint4*>(p.kout_v + bb * UA_HD + (sid & 31) * 8) = v1reg;
    }
    fence_proxy_async();
    mbar_arrive(q_ready);
    UD_TS(2);

    mbar_wait(s_full, 0);
    tc_fence_after();
    UD_TS(3);
    if (sid < nkeys * 32) *reinterpret_cast<uint4*>(sV1 + (sid >> 5) * UA_HD + (sid & 31) * 8) = v1reg;

    // ---- softmax: row r of the candidate sits in lane L = (r % 4) * 32 + r / 4  ->  this thread owns r = lane * 4 + q.
    // A warp can only read its own TMEM lane quarter and the 4 warps of a quarter share one scheduler, so the cost is
    // (columns x passes x instructions per element) per scheduler whatever the number of live rows: the loops below are
    // kept to ~7 instructions per element (raw max, one FFMA + ex2 per exponential, no per-element mask in full chunks).
    const int L = q * 32 + lane;
    const int r = lane * 4 + q;
    const bool active = r < rows_total;
    const int gi = active ? (r / p.heads) / p.tq : 0;  // this row's candidate within the group ...
    const int t = r / p.heads - gi * p.tq;             // ... its token ...
    const int lo = gi * p.kv1_len;                     // ... and where that candidate's keys start in the suffix tile
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const bool first_only = p.suffix_mask && t == 0;  // the state token sees only itself among the suffix keys
    // Column group g takes the prefix chunks g, g + 4, g + 8, ... and group 3 ends with the suffix chunk: a fixed
    // assignment, so the sums do not depend on how far the prefix range is padded.  Valid columns of a chunk are
    // always a prefix of it.
    const int npc = tk_pad / 16;  // prefix chunks
    const int nv_suffix = first_only ? 1 : p.kv1_len;
    const float c2 = p.scale * 1.4426950408889634f;
    // Pass A (online max / sum): the exponentials are written back over S in TMEM together with the running max they
    // refer to, so pass B needs one ex2 per CHUNK instead of one per element (the SFU handles 4 lanes per clock and
    // scheduler: the exponentials are the longest chain of the kernel).
    constexpr int UD_CPT = 6;  // prefix chunks per thread: ceil(352 / 16 / 4)
    float m = -INFINITY, l = 0.f;
    float mref[UD_CPT + 1];
#pragma unroll
    for (int k = 0; k < UD_CPT; ++k) {
      const int ch = g + 4 * k;
      mref[k] = -INFINITY;
      if (ch * 16 < n0) {  // warp-uniform
        uint32_t rr[16];
        tmem_ld_x16(taddr + ch * 16, rr);
        tmem_wait_ld();
        online_chunk_keep(rr, min(16, n0 - ch * 16), c2, m, l);
        tmem_st_x16(taddr + ch * 16, rr);
        mref[k] = m;
      }
    }
    mref[UD_CPT] = -INFINITY;
    if (g == 3) {
      uint32_t rr[16];
      tmem_ld_x16(taddr + tk_pad, rr);
      tmem_wait_ld();
      online_chunk_keep(rr, nv_suffix, c2, m, l, lo);
      tmem_st_x16(taddr + tk_pad, rr);
      mref[UD_CPT] = m;
    }
    tmem_wait_st();
    red2[g * 128 + L] = make_float2(m, l);
    named_bar(1, UD_SOFT);
    UD_TS(4);
    float M = -INFINITY;
#pragma unroll
    for (int gg = 0; gg < 4; ++gg) M = fmaxf(M, red2[gg * 128 + L].x);
    float Ls = 0.f;
#pragma unroll
    for (int gg = 0; gg < 4; ++gg) {
      const float2 sg = red2[gg * 128 + L];
      Ls += sg.y * ex2_approx((sg.x - M) * c2);  // a group without valid columns holds (-inf, 0)
    }
    const float inv = active ? 1.0f / Ls : 0.f;  // rows past the candidate's are written as zeros
    // Pass B: p = e * 2^((m_ref - M) c2) / L  ->  bf16  ->  swizzled A tile of the P.V GEMM
    const uint32_t p_row = smem_u32(sP) + static_cast<uint32_t>(L >> 3) * 1024u + static_cast<uint32_t>(L & 7) * 128u;
#pragma unroll
    for (int k = 0; k < UD_CPT; ++k) {
      const int ch = g + 4 * k;
      if (ch < npc) {
        uint32_t pk[8];
        if (ch * 16 < n0) {
          uint32_t rr[16];
          tmem_ld_x16(taddr + ch * 16, rr);
          tmem_wait_ld();
          const float f = active ? ex2_approx((mref[k] - M) * c2) * inv : 0.f;
#pragma unroll
          for (int i = 0; i < 8; ++i) pk[i] = pack_bf16x2(__uint_as_float(rr[2 * i]) * f, __uint_as_float(rr[2 * i + 1]) * f);
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i) pk[i] = 0u;
        }
        const int c0 = ch * 16;
        const uint32_t blk = p_row + static_cast<uint32_t>(c0 >> 6) * UA_PBLK;
        const int chunk = (c0 & 63) >> 3;
        sts_u4(blk + (static_cast<uint32_t>(chunk ^ (L & 7)) << 4), pk[0], pk[1], pk[2], pk[3]);
        sts_u4(blk + (static_cast<uint32_t>((chunk + 1) ^ (L & 7)) << 4), pk[4], pk[5], pk[6], pk[7]);
      }
    }
    if (g == 3) {
      uint32_t rr[16];
      tmem_ld_x16(taddr + tk_pad, rr);
      tmem_wait_ld();
      const float f = active ? ex2_approx((mref[UD_CPT] - M) * c2) * inv : 0.f;
      if (p.group <= 1) {
        *reinterpret_cast<uint4*>(psuf + L * 8) =  // suffix keys 0..7
            make_uint4(pack_bf16x2(__uint_as_float(rr[0]) * f, __uint_as_float(rr[1]) * f),
                       pack_bf16x2(__uint_as_float(rr[2]) * f, __uint_as_float(rr[3]) * f),
                       pack_bf16x2(__uint_as_float(rr[4]) * f, __uint_as_float(rr[5]) * f),
                       pack_bf16x2(__uint_as_float(rr[6]) * f, __uint_as_float(rr[7]) * f));
      } else {  // the row's OWN candidate's keys, shifted to 0.. (same values, same rounding as the packed stores)
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int j = i - lo;
          if (j >= 0 && j < 8) psuf[L * 8 + j] = __float2bfloat16_rn(__uint_as_float(rr[i]) * f);
        }
      }
    }
    fence_proxy_async();
    tc_fence_before();
    mbar_arrive(p_ready);
    UD_TS(5);

    // ---- epilogue.  While the P.V MMAs run: the suffix keys' contribution (<= 8 keys x 8 columns per item, fp32) for
    // the output items this thread will store (item = 8 consecutive columns of one row; 512 B per warp).
    named_bar(2, UD_SOFT);  // psuf / sV1 (written by other warps) are visible
    const int cgs = hdw >> 3;  // 8-column groups per output row of this CTA
    float sc[UD_IPT][8];
#pragma unroll
    for (int it = 0; it < UD_IPT; ++it) {
#pragma unroll
      for (int e = 0; e < 8; ++e) sc[it][e] = 0.f;
      const int item = sid + it * UD_SOFT;
      if (item < rows_total * cgs) {
        const int ro = item / cgs, cg = item % cgs;
        const int Lr = (ro & 3) * 32 + (ro >> 2);
        const int vlo = ((ro / p.heads) / p.tq) * p.kv1_len;  // first suffix-tile row of this row's candidate
        const uint4 pu = *reinterpret_cast<const uint4*>(psuf + Lr * 8);
        const uint32_t pw[4] = {pu.x, pu.y, pu.z, pu.w};
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (j < p.kv1_len) {
            const float2 pp = unpack_bf16x2(pw[j >> 1]);
            const float pj = (j & 1) ? pp.y : pp.x;
            const uint4 vv = *reinterpret_cast<const uint4*>(sV1 + (vlo + j) * UA_HD + hy * hdw + cg * 8);
            const uint32_t vw[4] = {vv.x, vv.y, vv.z, vv.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float2 f = unpack_bf16x2(vw[e]);
              sc[it][2 * e] = fmaf(pj, f.x, sc[it][2 * e]);
              sc[it][2 * e + 1] = fmaf(pj, f.y, sc[it][2 * e + 1]);
            }
          }
        }
      }
    }
    // (a) TMEM -> fp32 rows in shared memory (P / V^T are dead once o_full fires); warp (q, g) moves
    // O[rows of quarter q][64 g .. 64 g + 64)
    mbar_wait(o_full, 0);
    tc_fence_after();
    UD_TS(6);
    float* Os = reinterpret_cast<float*>(smem);
    const int wcols = hdw >> 2;  // columns moved by one warp: 64 or 32
    for (int cc = 0; cc < (wcols >> 4); ++cc) {
      uint32_t rr[16];
      tmem_ld_x16(taddr + g * wcols + cc * 16, rr);  // warp-collective (.sync.aligned): never under a divergent branch
      tmem_wait_ld();
      if (active) {
        float* dst = Os + r * UD_OS_LD + g * wcols + cc * 16;
#pragma unroll
        for (int v = 0; v < 4; ++v)
          *reinterpret_cast<uint4*>(dst + 4 * v) = make_uint4(rr[4 * v], rr[4 * v + 1], rr[4 * v + 2], rr[4 * v + 3]);
      }
    }
    named_bar(1, UD_SOFT);
    // (b) prefix part + suffix part -> bf16 -> global
#pragma unroll
    for (int it = 0; it < UD_IPT; ++it) {
      const int item = sid + it * UD_SOFT;
      if (item < rows_total * cgs) {
        const int ro = item / cgs, cg = item % cgs;
        const float4 o0 = *reinterpret_cast<const float4*>(Os + ro * UD_OS_LD + cg * 8);
        const float4 o1 = *reinterpret_cast<const float4*>(Os + ro * UD_OS_LD + cg * 8 + 4);
        const int og = (ro / p.heads) / p.tq;  // candidate within the group
        bf16* op = p.out + static_cast<long>(b + og) * p.o_bs + static_cast<long>(ro / p.heads - og * p.tq) * p.o_rs +
                   (ro % p.heads) * UA_HD + hy * hdw + cg * 8;
        *reinterpret_cast<uint4*>(op) =
            make_uint4(pack_bf16x2(o0.x + sc[it][0], o0.y + sc[it][1]), pack_bf16x2(o0.z + sc[it][2], o0.w + sc[it][3]),
                       pack_bf16x2(o1.x + sc[it][4], o1.y + sc[it][5]), pack_bf16x2(o1.z + sc[it][6], o1.w + sc[it][7]));
      }
    }
    UD_TS(7);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Multi-head self-attention of the SigLIP tower on tcgen05 (SiglipAttention reached through embed_image,
// paligemma_with_expert.py:229-230; 16 heads x head_dim 72, 256 tokens): one CTA = 128 query tokens of one head.
//   * Q / K tiles by 3-D TMA straight out of the fused qkv buffer ([head_dim, heads, tokens] view: columns past
//     head_dim are zero-filled, which pads 72 to the UMMA K granularity for free);
//   * V^T (the K-major "B" operand of P.V) is built by the softmax warps from plain loads while Q.K^T runs;
//   * S = Q K^T (N = keys <= 256), online softmax out of TMEM on 8 warps, P -> bf16 -> swizzled A tile,
//     O = P V (N = head_dim rounded up to 16), O -> bf16 -> global.
// Same rounding ledger as the mma.sync kernel it replaces (probabilities normalised, then rounded to bf16).
constexpr int UM_THREADS = 320;  // warp 0 TMA, warp 1 MMA, warps 2..9 V^T staging / softmax / epilogue
constexpr int UM_SOFT = 256;

struct UmmaMhaParams {
  const bf16* v;
  long v_bs, v_rs;
  const int* klen_dev;
  int klen;
  int tq, tk_pad, hd, hdp;
  long q_rows_per_batch, k_rows_per_batch;
  bf16* out;
  long o_bs, o_rs;
  float scale;
};

__global__ void __launch_bounds__(UM_THREADS, 1)
attn_mha_umma_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                     const UmmaMhaParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  const int tk_pad = p.tk_pad, hd = p.hd, hdp = p.hdp;
  const int nb = (hd + 63) / 64;        // hd-blocks of Q / K
  const int nkb = (tk_pad + 63) / 64;   // key-blocks of P / V^T
  const uint32_t kblk = static_cast<uint32_t>(tk_pad) * 128u;
  const uint32_t vblk = static_cast<uint32_t>(hdp) * 128u;
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + nb * UA_QBLK;
  uint8_t* sP = sK + nb * kblk;
  uint8_t* sV = sP + nkb * UA_PBLK;
  uint64_t* qk_full = reinterpret_cast<uint64_t*>(sV + nkb * vblk);  // [2]
  uint64_t* s_full = qk_full + 2;
  uint64_t* p_ready = s_full + 1;
  uint64_t* o_full = p_ready + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 1);
  float2* red2 = reinterpret_cast<float2*>(o_full + 3);  // [2][128]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = blockIdx.x, head = blockIdx.y, b = blockIdx.z;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
  }
  if (warp == 1) {
    if (lane == 0) {
      mbar_init(&qk_full[0], 1);
      mbar_init(&qk_full[1], 1);
      mbar_init(s_full, 1);
      mbar_init(p_ready, UM_SOFT);
      mbar_init(o_full, 1);
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
  pdl_launch();

  if (warp == 0) {
    if (lane == 0) {
      pdl_wait();
      const int q_row0 = static_cast<int>(b * p.q_rows_per_batch) + tile * 128;
      const int k_row0 = static_cast<int>(b * p.k_rows_per_batch);
      const int half_rows = tk_pad / 2;
      for (int c = 0; c < nb; ++c) {
        mbar_arrive_expect_tx(&qk_full[c], UA_QBLK + kblk);
        tma_load_3d(sQ + c * UA_QBLK, &tmQ, &qk_full[c], c * 64, head, q_row0);
        tma_load_3d(sK + c * kblk, &tmK, &qk_full[c], c * 64, head, k_row0);
        tma_load_3d(sK + c * kblk + static_cast<uint32_t>(half_rows) * 128u, &tmK, &qk_full[c], c * 64, head, k_row0 + half_rows);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc_s = make_idesc_n(tk_pad);
      for (int c = 0; c < nb; ++c) {
        mbar_wait(&qk_full[c], 0);
        tc_fence_after();
        const uint64_t qd = make_desc_kmajor_sw128(smem_u32(sQ + c * UA_QBLK));
        const uint64_t kd = make_desc_kmajor_sw128(smem_u32(sK + c * kblk));
        const int ksteps = min(4, (hd - c * 64 + 15) / 16);
        for (int k = 0; k < ksteps; ++k) umma_bf16(tmem_base, qd + 2 * k, kd + 2 * k, idesc_s, (c | k) != 0 ? 1u : 0u);
      }
      umma_commit(s_full);
      mbar_wait(p_ready, 0);  // P and V^T are in shared memory, S has been consumed
      tc_fence_after();
      const uint32_t idesc_o = make_idesc_n(hdp);
      for (int kb = 0; kb < nkb; ++kb) {
        const uint64_t pd = make_desc_kmajor_sw128(smem_u32(sP + kb * UA_PBLK));
        const uint64_t vd = make_desc_kmajor_sw128(smem_u32(sV + kb * vblk));
        const int ksteps = min(4, (tk_pad - kb * 64) / 16);
        for (int k = 0; k < ksteps; ++k) umma_bf16(tmem_base, pd + 2 * k, vd + 2 * k, idesc_o, (kb | k) != 0 ? 1u : 0u);
      }
      umma_commit(o_full);
    }
    __syncwarp();
  } else {
    const int sid = threadIdx.x - 64;
    const int q = warp & 3, half = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    pdl_wait();
    int n_keys = p.klen_dev != nullptr ? p.klen_dev[b] : p.klen;
    n_keys = max(1, min(n_keys, tk_pad));
    // ---- V^T[d][key] tiles (K-major, 128-byte swizzle) from V[key][d]: one key per thread, 16-byte loads, 2-byte stores
    for (int key = sid; key < tk_pad; key += UM_SOFT) {
      const bf16* vp = p.v + b * p.v_bs + static_cast<long>(key) * p.v_rs + head * hd;
      const uint32_t col = smem_u32(sV) + static_cast<uint32_t>(key >> 6) * vblk + static_cast<uint32_t>(key & 7) * 2u;
      const int chunk = (key & 63) >> 3;
      for (int d0 = 0; d0 < hdp; d0 += 8) {
        uint4 v = make_uint4(0, 0, 0, 0);
        if (key < n_keys && d0 < hd) v = *reinterpret_cast<const uint4*>(vp + d0);
        const uint32_t u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const int d = d0 + e;  // row of the tile; d & 7 == e
          const uint16_t val = static_cast<uint16_t>(e & 1 ? u[e >> 1] >> 16 : u[e >> 1] & 0xffffu);
          const uint32_t addr = col + static_cast<uint32_t>(d >> 3) * 1024u + static_cast<uint32_t>(e) * 128u +
                                (static_cast<uint32_t>(chunk ^ e) << 4);
          asm volatile("st.shared.b16 [%0], %1;" ::"r"(addr), "h"(val) : "memory");
        }
      }
    }
    const float c2 = p.scale * 1.4426950408889634f;
    mbar_wait(s_full, 0);
    tc_fence_after();
    float m = -INFINITY, l = 0.f;
    for (int ch = half; ch * 16 < n_keys; ch += 2) {
      uint32_t rr[16];
      tmem_ld_x16(taddr + ch * 16, rr);
      tmem_wait_ld();
      online_chunk(rr, min(16, n_keys - ch * 16), c2, m, l);
    }
    red2[half * 128 + row] = make_float2(m, l);
    named_bar(1, UM_SOFT);
    const float2 s0 = red2[row], s1 = red2[128 + row];
    const float M = fmaxf(s0.x, s1.x);
    const float Mc = M * c2;
    const float inv = 1.0f / (s0.y * ex2_approx((s0.x - M) * c2) + s1.y * ex2_approx((s1.x - M) * c2));
    const uint32_t p_row = smem_u32(sP) + static_cast<uint32_t>(row >> 3) * 1024u + static_cast<uint32_t>(row & 7) * 128u;
    for (int ch = half; ch * 16 < tk_pad; ch += 2) {
      const int nv = max(0, min(16, n_keys - ch * 16));
      uint32_t pk[8];
      if (nv > 0) {
        uint32_t rr[16];
        tmem_ld_x16(taddr + ch * 16, rr);
        tmem_wait_ld();
        prob_chunk(rr, nv, c2, Mc, inv, pk);
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) pk[i] = 0u;
      }
      const int c0 = ch * 16;
      const uint32_t blk = p_row + static_cast<uint32_t>(c0 >> 6) * UA_PBLK;
      const int chunk = (c0 & 63) >> 3;
      sts_u4(blk + (static_cast<uint32_t>(chunk ^ (row & 7)) << 4), pk[0], pk[1], pk[2], pk[3]);
      sts_u4(blk + (static_cast<uint32_t>((chunk + 1) ^ (row & 7)) << 4), pk[4], pk[5], pk[6], pk[7]);
    }
    fence_proxy_async();
    tc_fence_before();
    mbar_arrive(p_ready);

    mbar_wait(o_full, 0);
    tc_fence_after();
    const int t = tile * 128 + row;
    bf16* op = p.out + b * p.o_bs + static_cast<long>(t) * p.o_rs + head * hd;
    for (int ch = half; ch * 16 < hdp; ch += 2) {
      uint32_t rr[16];
      tmem_ld_x16(taddr + ch * 16, rr);
      tmem_wait_ld();
      if (t < p.tq) {
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          if (ch * 16 + j * 8 < hd)  // hd is a multiple of 8
            *reinterpret_cast<uint4*>(op + ch * 16 + j * 8) =
                make_uint4(pack_bf16x2(__uint_as_float(rr[8 * j]), __uint_as_float(rr[8 * j + 1])),
                           pack_bf16x2(__uint_as_float(rr[8 * j + 2]), __uint_as_float(rr[8 * j + 3])),
                           pack_bf16x2(__uint_as_float(rr[8 * j + 4]), __uint_as_float(rr[8 * j + 5])),
                           pack_bf16x2(__uint_as_float(rr[8 * j + 6]), __uint_as_float(rr[8 * j + 7])));
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Multi-head self-attention for LONGER sequences on tcgen05: the verifier's SigLIP2 ViT-L/16-384 trunk (576 tokens, 16
// heads x head_dim 64; timm Attention reached through VLA_SigLIP2_Bridge.extract_features,
// finetune_trajectory_bridge_ddp.py:297-355).  A 128 x 576 fp32 logit tile does not fit TMEM (512 columns), and the
// ledger normalises the probabilities BEFORE rounding them to bf16, so the kernel is exact two-pass over key chunks of
// 192:
//   pass 1:  S_c = Q K_c^T (double-buffered in TMEM) -> running (max, sum) per row;
//   pass 2:  S_c again (4 UMMAs per chunk at head_dim 64 - cheaper than keeping 576 columns), P_c = 2^((S_c - M) c2) / L
//            -> bf16 -> swizzled A tile, O += P_c V_c with O (head_dim columns) resident in TMEM.
// K for all keys stays in shared memory (one TMA pass), V^T is built by the softmax warps while pass 1 runs.
// Replaces attn_smem_kernel<64> (mma.sync, 54 us per launch at this shape) on the verifier trunk.
constexpr int UL_CH = 192;  // keys per chunk

struct UmmaLongParams {
  const bf16* v;
  long v_bs, v_rs;
  const int* klen_dev;
  int klen;
  int tq, tk_pad, hd, hdp, nc;  // nc = chunks
  long q_rows_per_batch, k_rows_per_batch;
  bf16* out;
  long o_bs, o_rs;
  float scale;
};

__host__ __device__ inline int ul_smem_bytes(int tk_pad, int hd, int hdp) {
  const int nb = (hd + 63) / 64, nc = (tk_pad + UL_CH - 1) / UL_CH, nkb = tk_pad / 64;
  return nb * UA_QBLK + nb * nc * UL_CH * 128 + 3 * UA_PBLK + nkb * hdp * 128;
}

__global__ void __launch_bounds__(UM_THREADS, 1)
attn_mha_long_umma_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                          const UmmaLongParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  const int tk_pad = p.tk_pad, hd = p.hd, hdp = p.hdp, nc = p.nc;
  const int nb = (hd + 63) / 64;   // hd-blocks of Q / K
  const int nkb = tk_pad / 64;     // key-blocks of V^T
  const uint32_t kblk = static_cast<uint32_t>(nc) * UL_CH * 128u;  // one hd-block of K: all chunks
  const uint32_t vblk = static_cast<uint32_t>(hdp) * 128u;
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + nb * UA_QBLK;
  uint8_t* sP = sK + nb * kblk;
  uint8_t* sV = sP + 3 * UA_PBLK;
  uint64_t* qk_full = reinterpret_cast<uint64_t*>(sV + nkb * vblk);  // [2]
  uint64_t* s_full = qk_full + 2;   // [2]
  uint64_t* s_free = s_full + 2;    // [2]
  uint64_t* p_ready = s_free + 2;
  uint64_t* p_free = p_ready + 1;
  uint64_t* o_full = p_free + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 1);
  float2* red2 = reinterpret_cast<float2*>(o_full + 3);  // [2][128]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = blockIdx.x, head = blockIdx.y, b = blockIdx.z;
  constexpr uint32_t O_COL = 2 * UL_CH;  // O accumulator behind the two S buffers

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int i = 0; i < 2; ++i) {
        mbar_init(&qk_full[i], 1);
        mbar_init(&s_full[i], 1);
        mbar_init(&s_free[i], UM_SOFT);
      }
      mbar_init(p_ready, UM_SOFT);
      mbar_init(p_free, 1);
      mbar_init(o_full, 1);
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
  pdl_launch();

  if (warp == 0) {
    if (lane == 0) {
      pdl_wait();
      const int q_row0 = static_cast<int>(b * p.q_rows_per_batch) + tile * 128;
      const int k_row0 = static_cast<int>(b * p.k_rows_per_batch);
      for (int c = 0; c < nb; ++c) {
        mbar_arrive_expect_tx(&qk_full[c], UA_QBLK + kblk);
        tma_load_3d(sQ + c * UA_QBLK, &tmQ, &qk_full[c], c * 64, head, q_row0);
        for (int ch = 0; ch < nc; ++ch)  // box = UL_CH rows; rows past the tensor are zero-filled
          tma_load_3d(sK + c * kblk + static_cast<uint32_t>(ch) * UL_CH * 128u, &tmK, &qk_full[c], c * 64, head, k_row0 + ch * UL_CH);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      for (int c = 0; c < nb; ++c) {
        mbar_wait(&qk_full[c], 0);
        tc_fence_after();
      }
      const uint32_t idesc_o = make_idesc_n(hdp);
      for (int j = 0; j < 2 * nc; ++j) {  // production j: chunk j % nc, pass j / nc, S buffer j & 1
        const int ch = j % nc, buf = j & 1, use = j >> 1;
        const int n_ch = min(UL_CH, tk_pad - ch * UL_CH);
        mbar_wait(&s_free[buf], (use & 1) ^ 1);
        tc_fence_after();
        const uint32_t idesc_s = make_idesc_n(n_ch);
        for (int c = 0; c < nb; ++c) {
          const uint64_t qd = make_desc_kmajor_sw128(smem_u32(sQ + c * UA_QBLK));
          const uint64_t kd = make_desc_kmajor_sw128(smem_u32(sK + c * kblk + static_cast<uint32_t>(ch) * UL_CH * 128u));
          const int ksteps = min(4, (hd - c * 64 + 15) / 16);
          for (int k = 0; k < ksteps; ++k)
            umma_bf16(tmem_base + buf * UL_CH, qd + 2 * k, kd + 2 * k, idesc_s, (c | k) != 0 ? 1u : 0u);
        }
        umma_commit(&s_full[buf]);
        if (j >= nc) {  // pass 2: O += P_ch V_ch once the softmax warps have written P_ch
          mbar_wait(p_ready, ch & 1);
          tc_fence_after();
          const int kbs = n_ch / 64;
          for (int kb = 0; kb < kbs; ++kb) {
            const uint64_t pd = make_desc_kmajor_sw128(smem_u32(sP + kb * UA_PBLK));
            const uint64_t vd = make_desc_kmajor_sw128(smem_u32(sV + (ch * 3 + kb) * vblk));
            for (int k = 0; k < 4; ++k) umma_bf16(tmem_base + O_COL, pd + 2 * k, vd + 2 * k, idesc_o, (ch | kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(p_free);
          if (ch == nc - 1) umma_commit(o_full);
        }
      }
    }
    __syncwarp();
  } else {
    const int sid = threadIdx.x - 64;
    const int q = warp & 3, half = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    pdl_wait();
    int n_keys = p.klen_dev != nullptr ? p.klen_dev[b] : p.klen;
    n_keys = max(1, min(n_keys, tk_pad));
    // ---- V^T[d][key] tiles (K-major, 128-byte swizzle) from V[key][d]: one key per thread
    for (int key = sid; key < tk_pad; key += UM_SOFT) {
      const bf16* vp = p.v + b * p.v_bs + static_cast<long>(key) * p.v_rs + head * hd;
      const uint32_t col = smem_u32(sV) + static_cast<uint32_t>(key >> 6) * vblk + static_cast<uint32_t>(key & 7) * 2u;
      const int chunk = (key & 63) >> 3;
      for (int d0 = 0; d0 < hdp; d0 += 8) {
        uint4 v = make_uint4(0, 0, 0, 0);
        if (key < n_keys && d0 < hd) v = *reinterpret_cast<const uint4*>(vp + d0);
        const uint32_t u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const int d = d0 + e;
          const uint16_t val = static_cast<uint16_t>(e & 1 ? u[e >> 1] >> 16 : u[e >> 1] & 0xffffu);
          const uint32_t addr = col + static_cast<uint32_t>(d >> 3) * 1024u + static_cast<uint32_t>(e) * 128u +
                                (static_cast<uint32_t>(chunk ^ e) << 4);
          asm volatile("st.shared.b16 [%0], %1;" ::"r"(addr), "h"(val) : "memory");
        }
      }
    }
    const float c2 = p.scale * 1.4426950408889634f;
    // ---- pass 1: exact row statistics over all keys
    float m = -INFINITY, l = 0.f;
    for (int j = 0; j < nc; ++j) {
      const int buf = j & 1, use = j >> 1;
      const int nv_c = max(0, min(UL_CH, n_keys - j * UL_CH));
      mbar_wait(&s_full[buf], use & 1);
      tc_fence_after();
      for (int ch = half; ch * 16 < nv_c; ch += 2) {
        uint32_t rr[16];
        tmem_ld_x16(taddr + buf * UL_CH + ch * 16, rr);
        tmem_wait_ld();
        online_chunk(rr, min(16, nv_c - ch * 16), c2, m, l);
      }
      tc_fence_before();
      mbar_arrive(&s_free[buf]);
    }
    red2[half * 128 + row] = make_float2(m, l);
    named_bar(1, UM_SOFT);
    const float2 s0 = red2[row], s1 = red2[128 + row];
    const float M = fmaxf(s0.x, s1.x);
    const float Mc = M * c2;
    const float inv = 1.0f / (s0.y * ex2_approx((s0.x - M) * c2) + s1.y * ex2_approx((s1.x - M) * c2));
    const uint32_t p_row = smem_u32(sP) + static_cast<uint32_t>(row >> 3) * 1024u + static_cast<uint32_t>(row & 7) * 128u;
    // ---- pass 2: probabilities of chunk c -> bf16 -> the A tile of O += P_c V_c
    for (int c = 0; c < nc; ++c) {
      const int j = nc + c, buf = j & 1, use = j >> 1;
      const int n_ch = min(UL_CH, tk_pad - c * UL_CH);
      const int nv_c = max(0, min(UL_CH, n_keys - c * UL_CH));
      mbar_wait(&s_full[buf], use & 1);
      tc_fence_after();
      if (c > 0) mbar_wait(p_free, (c - 1) & 1);  // the P.V MMAs of the previous chunk have read sP
      for (int ch = half; ch * 16 < n_ch; ch += 2) {
        const int nv = max(0, min(16, nv_c - ch * 16));
        uint32_t pk[8];
        if (nv > 0) {
          uint32_t rr[16];
          tmem_ld_x16(taddr + buf * UL_CH + ch * 16, rr);
          tmem_wait_ld();
          prob_chunk(rr, nv, c2, Mc, inv, pk);
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i) pk[i] = 0u;
        }
        const int c0 = ch * 16;
        const uint32_t blk = p_row + static_cast<uint32_t>(c0 >> 6) * UA_PBLK;
        const int chunk = (c0 & 63) >> 3;
        sts_u4(blk + (static_cast<uint32_t>(chunk ^ (row & 7)) << 4), pk[0], pk[1], pk[2], pk[3]);
        sts_u4(blk + (static_cast<uint32_t>((chunk + 1) ^ (row & 7)) << 4), pk[4], pk[5], pk[6], pk[7]);
      }
      fence_proxy_async();
      tc_fence_before();
      mbar_arrive(&s_free[buf]);
      mbar_arrive(p_ready);
    }
    // ---- epilogue
    mbar_wait(o_full, 0);
    tc_fence_after();
    const int t = tile * 128 + row;
    bf16* op = p.out + b * p.o_bs + static_cast<long>(t) * p.o_rs + head * hd;
    for (int ch = half; ch * 16 < hdp; ch += 2) {
      uint32_t rr[16];
      tmem_ld_x16(taddr + O_COL + ch * 16, rr);
      tmem_wait_ld();
      if (t < p.tq) {
#pragma unroll
        for (int jj = 0; jj < 2; ++jj) {
          if (ch * 16 + jj * 8 < hd)
            *reinterpret_cast<uint4*>(op + ch * 16 + jj * 8) =
                make_uint4(pack_bf16x2(__uint_as_float(rr[8 * jj]), __uint_as_float(rr[8 * jj + 1])),
                           pack_bf16x2(__uint_as_float(rr[8 * jj + 2]), __uint_as_float(rr[8 * jj + 3])),
                           pack_bf16x2(__uint_as_float(rr[8 * jj + 4]), __uint_as_float(rr[8 * jj + 5])),
                           pack_bf16x2(__uint_as_float(rr[8 * jj + 6]), __uint_as_float(rr[8 * jj + 7])));
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

struct Q3Key {
  const void* ptr;
  uint64_t rows, ld, heads, hd;
  uint32_t box_rows, box_heads;
  bool operator==(const Q3Key& o) const {
    return ptr == o.ptr && rows == o.rows && ld == o.ld && heads == o.heads && hd == o.hd && box_rows == o.box_rows &&
           box_heads == o.box_heads;
  }
};
struct Q3Hash {
  size_t operator()(const Q3Key& k) const {
    return reinterpret_cast<size_t>(k.ptr) ^ (k.rows * 0x9E3779B97F4A7C15ull) ^ (k.ld * 0xC2B2AE3D27D4EB4Full) ^
           (k.heads * 1315423911ull) ^ (k.hd << 20) ^ (static_cast<size_t>(k.box_rows) << 40) ^ k.box_heads;
  }
};
std::mutex g_q3_mu;
std::unordered_map<Q3Key, CUtensorMap, Q3Hash> g_q3;

int get_tmap3_cached(const void* ptr, uint64_t rows, uint64_t heads, uint64_t hd, uint64_t ld, uint32_t box_rows,
                     uint32_t box_heads, CUtensorMap* out) {
  std::lock_guard<std::mutex> lk(g_q3_mu);
  Q3Key key{ptr, rows, ld, heads, hd, box_rows, box_heads};
  auto it = g_q3.find(key);
  if (it == g_q3.end()) {
    CUtensorMap tm;
    CVB_TRY(make_tmap_3d_heads(&tm, ptr, rows, heads, hd, ld, box_rows, box_heads));
    it = g_q3.emplace(key, tm).first;
  }
  *out = it->second;
  return 0;
}

}  // namespace

bool attention_umma_eligible(const UmmaAttnCall& c) {
  const int tk_pad = (c.kmax + 15) / 16 * 16;
  return c.heads == UA_HEADS && c.head_dim == UA_HD && tk_pad >= 16 && tk_pad <= UA_MAX_KEYS && c.vt != nullptr &&
         c.q_ld % 8 == 0 && c.vt_ld % 8 == 0;
}

int attention_umma(cudaStream_t st, const UmmaAttnCall& c) {
  CVB_REQUIRE(attention_umma_eligible(c), "shape not eligible for the tcgen05 prefix attention");
  const int tk_pad = (c.kmax + 15) / 16 * 16;
  CUtensorMap tmQ, tmK, tmVT;
  CVB_TRY(get_tmap3_cached(c.q, c.q_total_rows, UA_HEADS, UA_HD, c.q_ld, UA_TOK, 0, &tmQ));
  CVB_TRY(get_tmap_cached(c.k, c.k_total_rows, UA_HD, UA_HD, tk_pad / 2, &tmK));
  CVB_TRY(get_tmap_cached(c.vt, static_cast<uint64_t>(c.batches) * UA_HD, c.vt_ld, c.vt_ld, UA_HD, &tmVT));
  UmmaAttnParams p;
  p.klen_dev = c.klen_dev, p.klen = c.klen, p.tq = c.tq, p.tk_pad = tk_pad;
  p.q_rows_per_batch = c.q_rows_per_batch, p.k_rows_per_batch = c.k_rows_per_batch;
  p.out = c.out, p.o_bs = c.o_batch_stride, p.o_rs = c.o_row_stride, p.scale = c.scale;
  const int nkb = (tk_pad + 63) / 64;
  const int nslots = std::min(std::min(nkb, UA_VSLOTS), (UA_BODY_MAX - nkb * UA_PBLK) / UA_VBLK);
  const int kslots = 4 * UA_QBLK + 4 * tk_pad * 128 <= UA_BODY_MAX ? 4 : 3;
  const int body = std::max(4 * UA_QBLK + kslots * tk_pad * 128, nkb * UA_PBLK + nslots * UA_VBLK);
  CVB_REQUIRE(body <= UA_BODY_MAX && nslots >= 1, "tcgen05 prefix attention operands do not fit shared memory");
  const int smem = 1024 + body + (4 + 1 + 1 + 2 * UA_VSLOTS + 2 + 2) * 8 + 2 * 128 * 8 + 64;
  CVB_TRY(ensure_dyn_smem(attn_prefix_umma_kernel, smem));
  dim3 grid((c.tq + UA_TOK - 1) / UA_TOK, c.batches);
  CVB_TRY(launch_pdl(attn_prefix_umma_kernel, grid, dim3(UA_THREADS), smem, st, 1, tmQ, tmK, tmVT, p));
  CVB_LAUNCHED();
  return 0;
}

bool attention_decode_umma_eligible(const AttnCall& c) {
  if ((c.k1 == nullptr && c.k1_part == nullptr) || c.vt0 == nullptr || c.kv_heads != 1 || c.head_dim != UA_HD || c.force_two_pass) return false;
  if (c.heads * c.tq > 128 || c.heads > 128 || c.kv1_len < 1 || c.kv1_len > 8) return false;
  if (c.kv0_row_stride != UA_HD || c.kv0_batch_stride % UA_HD != 0 || c.vt0_ld % 8 != 0) return false;
  const int kmax = c.kv0_len_dev != nullptr ? c.kv0_max : c.kv0_len;
  const int tk_pad = (kmax + 15) / 16 * 16;
  if (tk_pad < 16 || tk_pad > UA_MAX_KEYS - 16) return false;
  const int nkb = (tk_pad + 63) / 64;
  const int kslots = 4 * UA_QBLK + 4 * UD_KSBLK + 4 * tk_pad * 128 <= UA_BODY_MAX ? 4 : 3;
  const int vslots = std::min(nkb, (UA_BODY_MAX - UD_MISC - nkb * UA_PBLK) / UA_VBLK);
  return vslots >= 1 && 4 * UA_QBLK + 4 * UD_KSBLK + kslots * tk_pad * 128 <= UA_BODY_MAX &&
         ud_misc_off(nkb, vslots, c.heads * c.tq, UA_HD) + UD_MISC <= UA_BODY_MAX;
}

int attention_decode_umma(cudaStream_t st, const AttnCall& c) {
  CVB_REQUIRE(attention_decode_umma_eligible(c), "shape not eligible for the tcgen05 decode attention");
  const int kmax = c.kv0_len_dev != nullptr ? c.kv0_max : c.kv0_len;
  const int tk_pad = (kmax + 15) / 16 * 16;
  const int kv_batches = (c.batches + c.q_per_kv_batch - 1) / c.q_per_kv_batch;
  const long rows_per_batch = c.kv0_batch_stride / UA_HD;
  CUtensorMap tmK, tmVT;
  CVB_TRY(get_tmap_cached(c.k0, static_cast<uint64_t>(kv_batches) * rows_per_batch, UA_HD, UA_HD, tk_pad / 2, &tmK));
  // Two CTAs per candidate (each one half of the output columns) while that still fits one wave: the V^T tiles halve,
  // so all key-blocks are resident (no ring reuse on the critical path) and the P.V MMAs / epilogue take half the time
  const int split = c.batches * 2 <= device_sm_count() ? 2 : 1;
  const int hdw = UA_HD / split;
  CVB_TRY(get_tmap_cached(c.vt0, static_cast<uint64_t>(kv_batches) * UA_HD, c.vt0_ld, c.vt0_ld, hdw, &tmVT));
  UmmaDecodeParams p;
  p.hdw = hdw;
  p.q = c.q, p.q_bs = c.q_batch_stride, p.q_rs = c.q_row_stride;
  p.kv0_len_dev = c.kv0_len_dev, p.kv0_len = c.kv0_len, p.q_per_kv_batch = c.q_per_kv_batch;
  p.k_rows_per_batch = rows_per_batch;
  p.k1 = c.k1, p.v1 = c.v1, p.kv1_bs = c.kv1_batch_stride, p.kv1_rs = c.kv1_row_stride, p.kv1_len = c.kv1_len;
  p.suffix_mask = c.suffix_mask;
  p.out = c.out, p.o_bs = c.o_batch_stride, p.o_rs = c.o_row_stride;
  p.heads = c.heads, p.tq = c.tq, p.tk_pad = tk_pad, p.scale = c.scale, p.kv0_static = c.kv0_static, p.rope = c.rope;
  p.ts = g_skinny_ts;
  p.qf = c.q_part, p.k1f = c.k1_part, p.v1f = c.v1_part, p.part_S = c.part_splits, p.part_ss = c.part_split_stride;
  p.kc_k = c.kv1_cached_k, p.kc_v = c.kv1_cached_v, p.kc = c.kv1_cached_k != nullptr ? c.kv1_cached : 0;
  p.kout_k = c.kv1_cache_out_k, p.kout_v = c.kv1_cache_out_v;
  p.rope_rows = c.rope_rows > 0 ? c.rope_rows : c.tq, p.rope_off = c.rope_off;
  CVB_REQUIRE(p.kc >= 0 && p.kc < c.kv1_len, "cached suffix keys must leave at least one new key");
  CVB_REQUIRE(p.kc == 0 || !c.suffix_mask, "the suffix mask applies to the state-token query, which a hoisted call does not have");
  const int nkb = (tk_pad + 63) / 64;
  const int kslots = 4 * UA_QBLK + 4 * UD_KSBLK + 4 * tk_pad * 128 <= UA_BODY_MAX ? 4 : 3;
  const int vslots = std::min(std::min(nkb, UD_VSLOTS), (UA_BODY_MAX - UD_MISC - nkb * UA_PBLK) / (hdw * 128));
  // More candidates than SMs (several observations per call): the candidates of one rephrase share a CTA - up to 128 query
  // rows and 15 suffix keys - so the prefix K / V^T of a rephrase is staged once per group instead of once per candidate
  // and the launch fits one wave (8 observations x 40 candidates: 320 CTAs = 3 rounds -> 128 CTAs).  A row's arithmetic is
  // the same either way (bit-identical: tests/test_attention_gpu.py), so the choice may depend on the launch size.
  int group = 1, ngr = 1;
  const int rows_cand = c.heads * c.tq;
  static const int group_env = getenv("CVB_DECODE_GROUP") != nullptr ? atoi(getenv("CVB_DECODE_GROUP")) : -1;
  if ((split == 1 || group_env > 1) && c.q_per_kv_batch > 1 && c.batches % c.q_per_kv_batch == 0 &&
      (group_env > 1 || (group_env < 0 && c.batches > device_sm_count()))) {
    int gmax = std::min(std::min(128 / rows_cand, 15 / c.kv1_len), c.q_per_kv_batch);
    if (group_env > 1) gmax = std::min(gmax, group_env);
    if (gmax > 1) {
      ngr = (c.q_per_kv_batch + gmax - 1) / gmax;
      group = (c.q_per_kv_batch + ngr - 1) / ngr;  // balanced: 5 candidates -> 3 + 2
      if (ud_misc_off(nkb, vslots, group * rows_cand, hdw) + UD_MISC > UA_BODY_MAX) group = 1, ngr = 1;
    }
  }
  p.group = group, p.rows_max = group * rows_cand;
  const int body = std::max(4 * UA_QBLK + 4 * UD_KSBLK + kslots * tk_pad * 128,
                            ud_misc_off(nkb, vslots, p.rows_max, hdw) + UD_MISC);
  const int smem = 1024 + body + (4 + 1 + 1 + 1 + 2 * UD_VSLOTS + 2) * 8 + 16;
  const bool small = p.rows_max * (hdw / 8) <= 3 * UD_SOFT;  // the denoise step has 40 query rows
  auto kern = small ? attn_decode_umma_kernel<3> : attn_decode_umma_kernel<8>;
  CVB_TRY(ensure_dyn_smem(kern, smem));
  const int ctas = group > 1 ? kv_batches * ngr : c.batches;
  CVB_TRY(launch_pdl(kern, dim3(ctas, split), dim3(UD_THREADS), smem, st, 1, tmK, tmVT, p));
  CVB_LAUNCHED();
  return 0;
}

// ---- multi-head (kv_heads == heads) single-segment self-attention, <= 256 keys, head_dim <= 128 (SigLIP tower)
static int mha_smem_bytes(int tk_pad, int hd, int hdp) {
  const int nb = (hd + 63) / 64, nkb = (tk_pad + 63) / 64;
  return nb * UA_QBLK + nb * tk_pad * 128 + nkb * UA_PBLK + nkb * hdp * 128;
}

bool attention_mha_umma_eligible(const AttnCall& c) {
  if (c.k1 != nullptr || c.rope != nullptr || c.force_two_pass || c.heads != c.kv_heads) return false;
  if (c.head_dim % 8 != 0 || c.head_dim > 128 || c.head_dim < 16) return false;
  const int kmax = c.kv0_len_dev != nullptr ? c.kv0_max : c.kv0_len;
  if (kmax < 1 || kmax > 256) return false;
  if (c.q_row_stride % 8 != 0 || c.kv0_row_stride % 8 != 0 || c.o_row_stride % 8 != 0) return false;
  if ((reinterpret_cast<uintptr_t>(c.q) | reinterpret_cast<uintptr_t>(c.k0) | reinterpret_cast<uintptr_t>(c.v0) |
       reinterpret_cast<uintptr_t>(c.out)) & 15) return false;
  if (c.batches > 1 && (c.q_batch_stride % c.q_row_stride != 0 || c.kv0_batch_stride % c.kv0_row_stride != 0 ||
                        c.q_per_kv_batch != 1)) return false;
  const int tk_pad = (kmax + 15) / 16 * 16, hdp = (c.head_dim + 15) / 16 * 16;
  return mha_smem_bytes(tk_pad, c.head_dim, hdp) <= UA_BODY_MAX;
}

int attention_mha_umma(cudaStream_t st, const AttnCall& c) {
  CVB_REQUIRE(attention_mha_umma_eligible(c), "shape not eligible for the tcgen05 multi-head attention");
  const int kmax = c.kv0_len_dev != nullptr ? c.kv0_max : c.kv0_len;
  const int tk_pad = (kmax + 15) / 16 * 16, hd = c.head_dim, hdp = (hd + 15) / 16 * 16;
  const long q_rpb = c.batches > 1 ? c.q_batch_stride / c.q_row_stride : 0;
  const long k_rpb = c.batches > 1 ? c.kv0_batch_stride / c.kv0_row_stride : 0;
  CUtensorMap tmQ, tmK;
  CVB_TRY(get_tmap3_cached(c.q, static_cast<uint64_t>(q_rpb) * (c.batches - 1) + c.tq, c.heads, hd, c.q_row_stride, 128, 1, &tmQ));
  CVB_TRY(get_tmap3_cached(c.k0, static_cast<uint64_t>(k_rpb) * (c.batches - 1) + kmax, c.heads, hd, c.kv0_row_stride,
                           tk_pad / 2, 1, &tmK));
  UmmaMhaParams p;
  p.v = c.v0, p.v_bs = c.kv0_batch_stride, p.v_rs = c.kv0_row_stride;
  p.klen_dev = c.kv0_len_dev, p.klen = c.kv0_len, p.tq = c.tq, p.tk_pad = tk_pad, p.hd = hd, p.hdp = hdp;
  p.q_rows_per_batch = q_rpb, p.k_rows_per_batch = k_rpb;
  p.out = c.out, p.o_bs = c.o_batch_stride, p.o_rs = c.o_row_stride, p.scale = c.scale;
  const int smem = 1024 + mha_smem_bytes(tk_pad, hd, hdp) + 8 * 8 + 2 * 128 * 8 + 64;
  CVB_TRY(ensure_dyn_smem(attn_mha_umma_kernel, smem));
  dim3 grid((c.tq + 127) / 128, c.heads, c.batches);
  CVB_TRY(launch_pdl(attn_mha_umma_kernel, grid, dim3(UM_THREADS), smem, st, 1, tmQ, tmK, p));
  CVB_LAUNCHED();
  return 0;
}

// ---- longer multi-head self-attention (257 .. 768 keys, head_dim <= 128 as far as shared memory allows): verifier ViT-L
bool attention_mha_long_umma_eligible(const AttnCall& c) {
  if (c.k1 != nullptr || c.rope != nullptr || c.force_two_pass || c.heads != c.kv_heads) return false;
  if (c.head_dim % 8 != 0 || c.head_dim > 128 || c.head_dim < 16) return false;
  const int kmax = c.kv0_len_dev != nullptr ? c.kv0_max : c.kv0_len;
  if (kmax <= 256 || kmax > 768) return false;
  if (c.q_row_stride % 8 != 0 || c.kv0_row_stride % 8 != 0 || c.o_row_stride % 8 != 0) return false;
  if ((reinterpret_cast<uintptr_t>(c.q) | reinterpret_cast<uintptr_t>(c.k0) | reinterpret_cast<uintptr_t>(c.v0) |
       reinterpret_cast<uintptr_t>(c.out)) & 15) return false;
  if (c.batches > 1 && (c.q_batch_stride % c.q_row_stride != 0 || c.kv0_batch_stride % c.kv0_row_stride != 0 ||
                        c.q_per_kv_batch != 1)) return false;
  const int tk_pad = (kmax + 63) / 64 * 64, hdp = (c.head_dim + 15) / 16 * 16;
  return 2 * UL_CH + hdp <= 512 && ul_smem_bytes(tk_pad, c.head_dim, hdp) <= UA_BODY_MAX;
}

int attention_mha_long_umma(cudaStream_t st, const AttnCall& c) {
  CVB_REQUIRE(attention_mha_long_umma_eligible(c), "shape not eligible for the long tcgen05 multi-head attention");
  const int kmax = c.kv0_len_dev != nullptr ? c.kv0_max : c.kv0_len;
  const int tk_pad = (kmax + 63) / 64 * 64, hd = c.head_dim, hdp = (hd + 15) / 16 * 16;
  const long q_rpb = c.batches > 1 ? c.q_batch_stride / c.q_row_stride : 0;
  const long k_rpb = c.batches > 1 ? c.kv0_batch_stride / c.kv0_row_stride : 0;
  CUtensorMap tmQ, tmK;
  CVB_TRY(get_tmap3_cached(c.q, static_cast<uint64_t>(q_rpb) * (c.batches - 1) + c.tq, c.heads, hd, c.q_row_stride, 128, 1, &tmQ));
  CVB_TRY(get_tmap3_cached(c.k0, static_cast<uint64_t>(k_rpb) * (c.batches - 1) + kmax, c.heads, hd, c.kv0_row_stride, UL_CH, 1,
                           &tmK));
  UmmaLongParams p;
  p.v = c.v0, p.v_bs = c.kv0_batch_stride, p.v_rs = c.kv0_row_stride;
  p.klen_dev = c.kv0_len_dev, p.klen = c.kv0_len, p.tq = c.tq, p.tk_pad = tk_pad, p.hd = hd, p.hdp = hdp;
  p.nc = (tk_pad + UL_CH - 1) / UL_CH;
  p.q_rows_per_batch = q_rpb, p.k_rows_per_batch = k_rpb;
  p.out = c.out, p.o_bs = c.o_batch_stride, p.o_rs = c.o_row_stride, p.scale = c.scale;
  const int smem = 1024 + ul_smem_bytes(tk_pad, hd, hdp) + 12 * 8 + 2 * 128 * 8 + 64;
  CVB_TRY(ensure_dyn_smem(attn_mha_long_umma_kernel, smem));
  dim3 grid((c.tq + 127) / 128, c.heads, c.batches);
  CVB_TRY(launch_pdl(attn_mha_long_umma_kernel, grid, dim3(UM_THREADS), smem, st, 1, tmQ, tmK, p));
  CVB_LAUNCHED();
  return 0;
}

}  // namespace cvb
