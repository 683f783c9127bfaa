// Device helpers shared by the attention kernels (ops_attention.cu, ops_attention_decode.cu): ldmatrix / mma.sync
// wrappers, the per-warp Q.K^T tile and cp.async staging.
#pragma once
#include "ops.h"
#include "ptx.cuh"

namespace cvb {

constexpr int BQ = 64, BKV = 64, ATT_THREADS = 128;

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2,
                                        uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2,
                                          uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2,
                                         uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
      "{%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// S[16 x 64] (per warp) = Q_warp[16 x HDP] * K_tile[64 x HDP]^T
template <int HDP>
__device__ __forceinline__ void qk_tile(const bf16* Qs, const bf16* Ks, int warp, int lane,
                                        float (&s)[8][4]) {
  constexpr int LDS = HDP + 8;
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
  const uint32_t q_base = smem_u32(Qs + (warp * 16 + (lane & 15)) * LDS + (lane >> 4) * 8);
  const int mi = lane >> 3, ri = lane & 7;
  const uint32_t k_base = smem_u32(Ks + ((mi >> 1) * 8 + ri) * LDS + (mi & 1) * 8);
#pragma unroll 4
  for (int ks = 0; ks < HDP / 16; ++ks) {
    uint32_t a0, a1, a2, a3;
    ldsm_x4(q_base + ks * 32, a0, a1, a2, a3);
#pragma unroll
    for (int np = 0; np < 4; ++np) {
      uint32_t b0, b1, b2, b3;
      ldsm_x4(k_base + (np * 16 * LDS) * 2 + ks * 32, b0, b1, b2, b3);
      mma_bf16(s[2 * np], a0, a1, a2, a3, b0, b1);
      mma_bf16(s[2 * np + 1], a0, a1, a2, a3, b2, b3);
    }
  }
}

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}


}  // namespace cvb
