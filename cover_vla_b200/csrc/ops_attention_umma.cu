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
                       uint32_t box_rows);

namespace {

constexpr int UA_THREADS = 192;  // warp 0: TMA producer, warp 1: MMA issuer + TMEM owner, warps 2..5: softmax / epilogue
constexpr int UA_HD = 256;
constexpr int UA_TOK = 16;     // query tokens per CTA
constexpr int UA_HEADS = 8;    // query heads folded into rows: 16 x 8 = 128 UMMA rows
constexpr int UA_QBLK = 128 * 128;      // one hd-block of the Q tile: 128 rows x 64 bf16
constexpr int UA_VBLK = UA_HD * 128;    // one key-block of V^T: 256 rows x 64 keys
constexpr int UA_PBLK = 128 * 128;      // one key-block of P: 128 rows x 64 keys
constexpr int UA_VSLOTS = 4;
constexpr int UA_MAX_KEYS = 384;
constexpr int UA_BODY_MAX = 224 * 1024;  // operand bytes per CTA (barriers and the alignment slack come on top)

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

// r[0..31] <- 32 (or 16) consecutive fp32 columns of this thread's TMEM lane
__device__ __forceinline__ void ld_cols(uint32_t taddr, bool wide, uint32_t (&r)[32]) {
  if (wide) {
    tmem_ld_x32(taddr, r);
  } else {
    uint32_t r16[16];
    tmem_ld_x16(taddr, r16);
#pragma unroll
    for (int i = 0; i < 16; ++i) r[i] = r16[i];
#pragma unroll
    for (int i = 16; i < 32; ++i) r[i] = 0u;
  }
  tmem_wait_ld();
}

__global__ void __launch_bounds__(UA_THREADS, 1)
attn_prefix_umma_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                        const __grid_constant__ CUtensorMap tmVT, const UmmaAttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  const int tk_pad = p.tk_pad;
  const uint32_t kblk = static_cast<uint32_t>(tk_pad) * 128u;  // one hd-block of K: tk_pad rows x 64 bf16
  const int nkb = (tk_pad + 63) / 64;                          // key blocks of P / V^T
  const int nslots = min(nkb, UA_VSLOTS);
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
      mbar_init(p_ready, 128);
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
        const int slot = kb % UA_VSLOTS;
        if (kb >= UA_VSLOTS) mbar_wait(&v_empty[slot], ((kb / UA_VSLOTS) - 1) & 1);
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
        const int slot = kb % UA_VSLOTS;
        mbar_wait(&v_full[slot], (kb / UA_VSLOTS) & 1);
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
    // ------------------------------------------------------------------ softmax + epilogue: one row per thread
    pdl_wait();
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int tl = row >> 3, h = row & 7;
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    int n_keys = p.klen_dev != nullptr ? p.klen_dev[b] : p.klen;
    n_keys = max(1, min(n_keys, tk_pad));
    mbar_wait(s_full, 0);
    tc_fence_after();
    float m = -INFINITY;
    for (int c0 = 0; c0 < n_keys; c0 += 32) {
      uint32_t r[32];
      ld_cols(taddr + c0, c0 + 32 <= tk_pad, r);
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (c0 + i < n_keys) m = fmaxf(m, __uint_as_float(r[i]) * p.scale);
    }
    float l = 0.f;
    for (int c0 = 0; c0 < n_keys; c0 += 32) {
      uint32_t r[32];
      ld_cols(taddr + c0, c0 + 32 <= tk_pad, r);
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (c0 + i < n_keys) l += __expf(__uint_as_float(r[i]) * p.scale - m);
    }
    const float inv = 1.0f / l;
    // P[row][key] -> canonical K-major SWIZZLE_128B tile: 8-row groups of 1024 B, 16-byte chunk index XOR (row % 8)
    const uint32_t p_row = smem_u32(sP) + static_cast<uint32_t>(row >> 3) * 1024u + static_cast<uint32_t>(row & 7) * 128u;
    for (int c0 = 0; c0 < tk_pad; c0 += 32) {
      const bool wide = c0 + 32 <= tk_pad;
      uint32_t r[32];
      if (c0 < n_keys) {
        ld_cols(taddr + c0, wide, r);
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) r[i] = 0u;
      }
      uint32_t pk[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float a = c0 + 2 * i < n_keys ? __expf(__uint_as_float(r[2 * i]) * p.scale - m) * inv : 0.f;
        const float bb = c0 + 2 * i + 1 < n_keys ? __expf(__uint_as_float(r[2 * i + 1]) * p.scale - m) * inv : 0.f;
        pk[i] = pack_bf16x2(a, bb);
      }
      const uint32_t blk = p_row + static_cast<uint32_t>(c0 >> 6) * UA_PBLK;
      const int ch0 = (c0 & 63) >> 3;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (j < 2 || wide)
          sts_u4(blk + (static_cast<uint32_t>((ch0 + j) ^ (row & 7)) << 4), pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
      }
    }
    fence_proxy_async();  // generic-proxy writes of P -> visible to the tensor core's async-proxy reads
    tc_fence_before();
    mbar_arrive(p_ready);

    mbar_wait(o_full, 0);
    tc_fence_after();
    const int t = tile * UA_TOK + tl;
    bf16* op = p.out + b * p.o_bs + static_cast<long>(t) * p.o_rs + h * UA_HD;
    for (int c0 = 0; c0 < UA_HD; c0 += 32) {
      uint32_t r[32];
      tmem_ld_x32(taddr + c0, r);
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

struct Q3Key {
  const void* ptr;
  uint64_t rows, ld;
  bool operator==(const Q3Key& o) const { return ptr == o.ptr && rows == o.rows && ld == o.ld; }
};
struct Q3Hash {
  size_t operator()(const Q3Key& k) const {
    return reinterpret_cast<size_t>(k.ptr) ^ (k.rows * 0x9E3779B97F4A7C15ull) ^ (k.ld * 0xC2B2AE3D27D4EB4Full);
  }
};
std::mutex g_q3_mu;
std::unordered_map<Q3Key, CUtensorMap, Q3Hash> g_q3;

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
  {
    std::lock_guard<std::mutex> lk(g_q3_mu);
    Q3Key key{c.q, static_cast<uint64_t>(c.q_total_rows), static_cast<uint64_t>(c.q_ld)};
    auto it = g_q3.find(key);
    if (it == g_q3.end()) {
      CUtensorMap tm;
      CVB_TRY(make_tmap_3d_heads(&tm, c.q, c.q_total_rows, UA_HEADS, UA_HD, c.q_ld, UA_TOK));
      it = g_q3.emplace(key, tm).first;
    }
    tmQ = it->second;
  }
  CVB_TRY(get_tmap_cached(c.k, c.k_total_rows, UA_HD, UA_HD, tk_pad / 2, &tmK));
  CVB_TRY(get_tmap_cached(c.vt, static_cast<uint64_t>(c.batches) * UA_HD, c.vt_ld, c.vt_ld, UA_HD, &tmVT));
  UmmaAttnParams p;
  p.klen_dev = c.klen_dev, p.klen = c.klen, p.tq = c.tq, p.tk_pad = tk_pad;
  p.q_rows_per_batch = c.q_rows_per_batch, p.k_rows_per_batch = c.k_rows_per_batch;
  p.out = c.out, p.o_bs = c.o_batch_stride, p.o_rs = c.o_row_stride, p.scale = c.scale;
  const int nkb = (tk_pad + 63) / 64;
  const int nslots = std::min(nkb, UA_VSLOTS);
  const int kslots = 4 * UA_QBLK + 4 * tk_pad * 128 <= UA_BODY_MAX ? 4 : 3;
  const int body = std::max(4 * UA_QBLK + kslots * tk_pad * 128, nkb * UA_PBLK + nslots * UA_VBLK);
  CVB_REQUIRE(body <= UA_BODY_MAX, "tcgen05 prefix attention operands do not fit shared memory");
  const int smem = 1024 + body + (4 + 1 + 1 + 2 * UA_VSLOTS + 2) * 8 + 16;
  static int attr_smem = 0;
  if (smem > attr_smem) {
    CVB_CUDA(cudaFuncSetAttribute(attn_prefix_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_smem = smem;
  }
  dim3 grid((c.tq + UA_TOK - 1) / UA_TOK, c.batches);
  CVB_TRY(launch_pdl(attn_prefix_umma_kernel, grid, dim3(UA_THREADS), smem, st, 1, tmQ, tmK, tmVT, p));
  CVB_LAUNCHED();
  return 0;
}

}  // namespace cvb
