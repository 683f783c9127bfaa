// Fused elementwise / normalisation kernels of the pi0 path.  All of them are HBM- or latency-bound:
// 16-byte vectorised, coalesced accesses; one CTA per row (or per small group of rows); the rounding
// points follow SURVEY.md Appendix A.
#include <algorithm>
#include "host_common.h"
#include "ops.h"
#include "ptx.cuh"

namespace cvb {

// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float block_sum(float v, float* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads();  // `red` may still be read by a previous reduction of the same kernel
  if (lane == 0) red[warp] = v;
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  float t = (lane < nw) ? red[lane] : 0.f;
  t = warp_sum(t);
  return t;
}

// ------------------------------------------------------------------------------------------------
// Gemma RMSNorm (transformers GemmaRMSNorm; used at paligemma_with_expert.py:268,335,355)
template <bool X_F32, bool W_F32>
__global__ void __launch_bounds__(256) rmsnorm_kernel(const void* __restrict__ x, long ldx,
                                                      const void* __restrict__ w,
                                                      bf16* __restrict__ y, long ldy, int rows,
                                                      int width, float eps,
                                                      const int* __restrict__ rows_dev) {
  pdl_wait();
  pdl_launch();
  __shared__ float red[32];
  const int row = blockIdx.x;
  if (rows_dev != nullptr && row >= *rows_dev) return;
  float ss = 0.f;
  if constexpr (X_F32) {
    const float* xr = reinterpret_cast<const float*>(x) + row * ldx;
    for (int i = threadIdx.x * 4; i < width; i += blockDim.x * 4) {
      const float4 v = *reinterpret_cast<const float4*>(xr + i);
      ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    }
  } else {
    const bf16* xr = reinterpret_cast<const bf16*>(x) + row * ldx;
    for (int i = threadIdx.x * 8; i < width; i += blockDim.x * 8) {
      const uint4 v = *reinterpret_cast<const uint4*>(xr + i);
      const uint32_t u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f = unpack_bf16x2(u[e]);
        ss += f.x * f.x + f.y * f.y;
      }
    }
  }
  ss = block_sum(ss, red);
  const float r = 1.0f / sqrtf(ss / static_cast<float>(width) + eps);
  bf16* yr = y + row * ldy;
  for (int i = threadIdx.x * 8; i < width; i += blockDim.x * 8) {
    float xv[8], wv[8];
    if constexpr (X_F32) {
      const float* xr = reinterpret_cast<const float*>(x) + row * ldx + i;
      const float4 a = *reinterpret_cast<const float4*>(xr);
      const float4 b = *reinterpret_cast<const float4*>(xr + 4);
      xv[0] = a.x, xv[1] = a.y, xv[2] = a.z, xv[3] = a.w, xv[4] = b.x, xv[5] = b.y, xv[6] = b.z,
      xv[7] = b.w;
    } else {
      const uint4 v = *reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(x) + row * ldx + i);
      const uint32_t u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f = unpack_bf16x2(u[e]);
        xv[2 * e] = f.x, xv[2 * e + 1] = f.y;
      }
    }
    if constexpr (W_F32) {
      const float* wr = reinterpret_cast<const float*>(w) + i;
#pragma unroll
      for (int e = 0; e < 8; ++e) wv[e] = wr[e];
    } else {
      const uint4 v = *reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(w) + i);
      const uint32_t u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f = unpack_bf16x2(u[e]);
        wv[2 * e] = f.x, wv[2 * e + 1] = f.y;
      }
    }
    float o[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) o[e] = (xv[e] * r) * (1.0f + wv[e]);
    *reinterpret_cast<uint4*>(yr + i) = make_uint4(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]),
                                                   pack_bf16x2(o[4], o[5]), pack_bf16x2(o[6], o[7]));
  }
}

// One WARP per row (bf16 x, bf16 w, width = NV * 256 <= 2048): the whole row sits in registers (NV 16-byte loads in flight
// per lane), the norm weight is loaded before the dependency wait, the reduction is five shuffles - no block barrier, one
// pass over x.  The prefix runs this 35 x on 2240 rows x 2048: 7.0 -> 3.5 us per launch against the CTA-per-row kernel.
template <int NV>
__global__ void __launch_bounds__(256) rmsnorm_warp_kernel(const bf16* __restrict__ x, long ldx, const bf16* __restrict__ w,
                                                           bf16* __restrict__ y, long ldy, int rows, float eps,
                                                           const int* __restrict__ rows_dev) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  uint4 wv[NV];
#pragma unroll
  for (int v = 0; v < NV; ++v) wv[v] = *reinterpret_cast<const uint4*>(w + (v * 32 + lane) * 8);
  pdl_wait();
  pdl_launch();
  if (row >= rows || (rows_dev != nullptr && row >= *rows_dev)) return;
  const bf16* xr = x + row * ldx;
  uint4 xv[NV];
#pragma unroll
  for (int v = 0; v < NV; ++v) xv[v] = *reinterpret_cast<const uint4*>(xr + (v * 32 + lane) * 8);
  float ss = 0.f;
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    const uint32_t u[4] = {xv[v].x, xv[v].y, xv[v].z, xv[v].w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 f = unpack_bf16x2(u[e]);
      ss += f.x * f.x + f.y * f.y;
    }
  }
  ss = warp_sum(ss);
  const float r = 1.0f / sqrtf(ss / static_cast<float>(NV * 256) + eps);
  bf16* yr = y + row * ldy;
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    const uint32_t u[4] = {xv[v].x, xv[v].y, xv[v].z, xv[v].w}, g[4] = {wv[v].x, wv[v].y, wv[v].z, wv[v].w};
    uint32_t o[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 f = unpack_bf16x2(u[e]), q = unpack_bf16x2(g[e]);
      o[e] = pack_bf16x2((f.x * r) * (1.0f + q.x), (f.y * r) * (1.0f + q.y));
    }
    *reinterpret_cast<uint4*>(yr + (v * 32 + lane) * 8) = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

int rmsnorm(cudaStream_t st, const void* x, int x_is_f32, long ldx, const void* w, int w_is_f32,
            bf16* y, long ldy, int rows, int width, float eps, const int* rows_dev) {
  CVB_REQUIRE(width % 8 == 0, "rmsnorm width must be a multiple of 8");
  // (the choice must not depend on `rows`: the two kernels sum a row's squares in different orders, and a prompt's prefix
  // has to come out the same whether it is computed alone - one rephrase per rank of a sharded decision - or beside others)
  if (!x_is_f32 && !w_is_f32 && width % 256 == 0 && width <= 2048 && ldx % 8 == 0 && ldy % 8 == 0 &&
      ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(w) | reinterpret_cast<uintptr_t>(y)) & 15) == 0) {
    const bf16* xb = reinterpret_cast<const bf16*>(x);
    const bf16* wb = reinterpret_cast<const bf16*>(w);
    const int wpc = rows >= 512 ? 8 : 2;  // rows (warps) per CTA: few rows are spread over more SMs
    const dim3 grid((rows + wpc - 1) / wpc), block(32 * wpc);
#define CVB_RW(NV) \
  CVB_TRY(launch_pdl(rmsnorm_warp_kernel<NV>, grid, block, 0, st, 1, xb, ldx, wb, y, ldy, rows, eps, rows_dev))
    switch (width / 256) {
      case 1: CVB_RW(1); break;
      case 2: CVB_RW(2); break;
      case 3: CVB_RW(3); break;
      case 4: CVB_RW(4); break;
      case 5: CVB_RW(5); break;
      case 6: CVB_RW(6); break;
      case 7: CVB_RW(7); break;
      default: CVB_RW(8); break;
    }
#undef CVB_RW
    CVB_LAUNCHED();
    return 0;
  }
  const int threads = width >= 2048 ? 256 : (width >= 512 ? 128 : 64);
  if (x_is_f32 && w_is_f32)
    CVB_TRY(launch_pdl(rmsnorm_kernel<true, true>, dim3(rows), dim3(threads), 0, st, 1, x, ldx, w, y, ldy, rows, width, eps, rows_dev));
  else if (x_is_f32)
    CVB_TRY(launch_pdl(rmsnorm_kernel<true, false>, dim3(rows), dim3(threads), 0, st, 1, x, ldx, w, y, ldy, rows, width, eps, rows_dev));
  else if (w_is_f32)
    CVB_TRY(launch_pdl(rmsnorm_kernel<false, true>, dim3(rows), dim3(threads), 0, st, 1, x, ldx, w, y, ldy, rows, width, eps, rows_dev));
  else
    CVB_TRY(launch_pdl(rmsnorm_kernel<false, false>, dim3(rows), dim3(threads), 0, st, 1, x, ldx, w, y, ldy, rows, width, eps, rows_dev));
  CVB_LAUNCHED();
  return 0;
}

// ------------------------------------------------------------------------------------------------
// Tail of a split-K linear layer fused with the Gemma RMSNorm that follows it (o_proj -> post_attention_layernorm,
// down_proj -> next input_layernorm / final norm; paligemma_with_expert.py:327-355).  One CTA per row sums the S fp32
// partials of gemm_splitk_partial_tcgen05 in split order, applies the reference's rounding points
// (bf16(acc), + residual, bf16) and normalises the new residual-stream row.
template <bool R_F32, bool W_F32>
__global__ void __launch_bounds__(256) rmsnorm_reduce_kernel(const float* __restrict__ P, int S, long split_stride,
                                                             long ldp, const void* resid, long ldr,
                                                             const void* __restrict__ w, bf16* h_out, long ldh,
                                                             bf16* __restrict__ y, long ldy, int width, float eps) {
  // loads that do not depend on the preceding kernel (norm weight) are issued before the dependency wait
  extern __shared__ float hrow[];  // the new residual row (bf16 values held as fp32)
  __shared__ float red[32];
  const int row = blockIdx.x;
  const bool one = width <= static_cast<int>(blockDim.x) * 4;  // one float4 per thread: everything stays in registers
  const int i0 = threadIdx.x * 4;
  float wreg[4] = {0.f, 0.f, 0.f, 0.f};
  if (one && i0 < width) {
    if constexpr (W_F32) {
      const float4 v = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(w) + i0);
      wreg[0] = v.x, wreg[1] = v.y, wreg[2] = v.z, wreg[3] = v.w;
    } else {
      const uint2 v = *reinterpret_cast<const uint2*>(reinterpret_cast<const bf16*>(w) + i0);
      const float2 f0 = unpack_bf16x2(v.x), f1 = unpack_bf16x2(v.y);
      wreg[0] = f0.x, wreg[1] = f0.y, wreg[2] = f1.x, wreg[3] = f1.y;
    }
  }
  pdl_wait();
  pdl_launch();
  const float* pr = P + row * ldp;
  float ss = 0.f;
  float hreg[4] = {0.f, 0.f, 0.f, 0.f};
  for (int i = i0; i < width; i += blockDim.x * 4) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    float rr[4];
    // residual first, then ALL partials of this element group in flight at once (up to 8 per round), summed in split
    // order - the order (and therefore every bit) is the same as a one-at-a-time loop
    float4 rv4 = make_float4(0.f, 0.f, 0.f, 0.f);
    uint2 rv2 = make_uint2(0u, 0u);
    if constexpr (R_F32)
      rv4 = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(resid) + row * ldr + i);
    else
      rv2 = *reinterpret_cast<const uint2*>(reinterpret_cast<const bf16*>(resid) + row * ldr + i);
    int s = 0;
    for (; s + 8 <= S; s += 8) {
      float4 a[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) a[j] = *reinterpret_cast<const float4*>(pr + (s + j) * split_stride + i);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc.x += a[j].x, acc.y += a[j].y, acc.z += a[j].z, acc.w += a[j].w;
    }
    if (s + 4 <= S) {
      float4 a[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) a[j] = *reinterpret_cast<const float4*>(pr + (s + j) * split_stride + i);
#pragma unroll
      for (int j = 0; j < 4; ++j) acc.x += a[j].x, acc.y += a[j].y, acc.z += a[j].z, acc.w += a[j].w;
      s += 4;
    }
    for (; s < S; ++s) {
      const float4 a = *reinterpret_cast<const float4*>(pr + s * split_stride + i);
      acc.x += a.x, acc.y += a.y, acc.z += a.z, acc.w += a.w;
    }
    if constexpr (R_F32) {
      rr[0] = rv4.x, rr[1] = rv4.y, rr[2] = rv4.z, rr[3] = rv4.w;
    } else {
      const float2 f0 = unpack_bf16x2(rv2.x), f1 = unpack_bf16x2(rv2.y);
      rr[0] = f0.x, rr[1] = f0.y, rr[2] = f1.x, rr[3] = f1.y;
    }
    const float hv[4] = {bf16_round(bf16_round(acc.x) + rr[0]), bf16_round(bf16_round(acc.y) + rr[1]),
                         bf16_round(bf16_round(acc.z) + rr[2]), bf16_round(bf16_round(acc.w) + rr[3])};
    *reinterpret_cast<uint2*>(h_out + row * ldh + i) = make_uint2(pack_bf16x2(hv[0], hv[1]), pack_bf16x2(hv[2], hv[3]));
    if (one) {
      hreg[0] = hv[0], hreg[1] = hv[1], hreg[2] = hv[2], hreg[3] = hv[3];
    } else {
      *reinterpret_cast<float4*>(hrow + i) = make_float4(hv[0], hv[1], hv[2], hv[3]);
    }
    ss += hv[0] * hv[0] + hv[1] * hv[1] + hv[2] * hv[2] + hv[3] * hv[3];
  }
  ss = block_sum(ss, red);
  const float r = 1.0f / sqrtf(ss / static_cast<float>(width) + eps);
  bf16* yr = y + row * ldy;
  if (one) {
    if (i0 < width)
      *reinterpret_cast<uint2*>(yr + i0) =
          make_uint2(pack_bf16x2((hreg[0] * r) * (1.0f + wreg[0]), (hreg[1] * r) * (1.0f + wreg[1])),
                     pack_bf16x2((hreg[2] * r) * (1.0f + wreg[2]), (hreg[3] * r) * (1.0f + wreg[3])));
    return;
  }
  for (int i = threadIdx.x * 4; i < width; i += blockDim.x * 4) {  // each thread re-reads its own hrow entries
    const float4 hv = *reinterpret_cast<const float4*>(hrow + i);
    float wv[4];
    if constexpr (W_F32) {
      const float4 v = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(w) + i);
      wv[0] = v.x, wv[1] = v.y, wv[2] = v.z, wv[3] = v.w;
    } else {
      const uint2 v = *reinterpret_cast<const uint2*>(reinterpret_cast<const bf16*>(w) + i);
      const float2 f0 = unpack_bf16x2(v.x), f1 = unpack_bf16x2(v.y);
      wv[0] = f0.x, wv[1] = f0.y, wv[2] = f1.x, wv[3] = f1.y;
    }
    *reinterpret_cast<uint2*>(yr + i) =
        make_uint2(pack_bf16x2((hv.x * r) * (1.0f + wv[0]), (hv.y * r) * (1.0f + wv[1])),
                   pack_bf16x2((hv.z * r) * (1.0f + wv[2]), (hv.w * r) * (1.0f + wv[3])));
  }
}

int rmsnorm_reduce(cudaStream_t st, const float* P, int S, long split_stride, long ldp, const void* resid,
                   int resid_is_f32, long ldr, const void* w, int w_is_f32, bf16* h_out, long ldh, bf16* y, long ldy,
                   int rows, int width, float eps) {
  CVB_REQUIRE(width % 4 == 0 && ldp % 4 == 0 && split_stride % 4 == 0 && ldr % 4 == 0 && ldh % 4 == 0 && ldy % 4 == 0,
              "rmsnorm_reduce needs 4-element aligned rows");
  CVB_REQUIRE(S >= 1 && resid != nullptr, "rmsnorm_reduce needs >= 1 partial and a residual");
  const int threads = width >= 1024 ? 256 : 128;
  const size_t smem = static_cast<size_t>(width) * sizeof(float);
#define CVB_RR(RF, WF)                                                                                             \
  CVB_TRY(launch_pdl(rmsnorm_reduce_kernel<RF, WF>, dim3(rows), dim3(threads), smem, st, 1, P, S, split_stride, ldp, \
                     resid, ldr, w, h_out, ldh, y, ldy, width, eps))
  if (resid_is_f32 && w_is_f32)
    CVB_RR(true, true);
  else if (resid_is_f32)
    CVB_RR(true, false);
  else if (w_is_f32)
    CVB_RR(false, true);
  else
    CVB_RR(false, false);
#undef CVB_RR
  CVB_LAUNCHED();
  return 0;
}

// ------------------------------------------------------------------------------------------------
// LayerNorm over bf16 rows (SigLIP / ViT-L / text tower blocks).  fp32 statistics (two-pass).
__global__ void __launch_bounds__(256) layernorm_bf16_kernel(const bf16* __restrict__ x, long ldx,
                                                             const bf16* __restrict__ w,
                                                             const bf16* __restrict__ b,
                                                             bf16* __restrict__ y, long ldy,
                                                             int width, float eps) {
  pdl_wait();
  pdl_launch();
  __shared__ float red[32];
  const bf16* xr = x + blockIdx.x * ldx;
  float s = 0.f;
  for (int i = threadIdx.x * 8; i < width; i += blockDim.x * 8) {
    const uint4 v = *reinterpret_cast<const uint4*>(xr + i);
    const uint32_t u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 f = unpack_bf16x2(u[e]);
      s += f.x + f.y;
    }
  }
  const float mean = block_sum(s, red) / static_cast<float>(width);
  float vs = 0.f;
  for (int i = threadIdx.x * 8; i < width; i += blockDim.x * 8) {
    const uint4 v = *reinterpret_cast<const uint4*>(xr + i);
    const uint32_t u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 f = unpack_bf16x2(u[e]);
      vs += (f.x - mean) * (f.x - mean) + (f.y - mean) * (f.y - mean);
    }
  }
  const float var = block_sum(vs, red) / static_cast<float>(width);
  const float rstd = 1.0f / sqrtf(var + eps);
  bf16* yr = y + blockIdx.x * ldy;
  for (int i = threadIdx.x * 8; i < width; i += blockDim.x * 8) {
    const uint4 v = *reinterpret_cast<const uint4*>(xr + i);
    const uint4 wv = *reinterpret_cast<const uint4*>(w + i);
    const uint4 bv = *reinterpret_cast<const uint4*>(b + i);
    const uint32_t u[4] = {v.x, v.y, v.z, v.w}, uw[4] = {wv.x, wv.y, wv.z, wv.w},
                   ub[4] = {bv.x, bv.y, bv.z, bv.w};
    uint32_t o[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 f = unpack_bf16x2(u[e]), g = unpack_bf16x2(uw[e]), h = unpack_bf16x2(ub[e]);
      o[e] = pack_bf16x2((f.x - mean) * rstd * g.x + h.x, (f.y - mean) * rstd * g.y + h.y);
    }
    *reinterpret_cast<uint4*>(yr + i) = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

// Tail of a split-K linear layer fused with the LayerNorm that follows it (SigLIP encoder block: out_proj + residual ->
// layer_norm2, fc2 + residual -> next layer_norm1 / post_layernorm; reached through embed_image,
// paligemma_with_expert.py:229-230).  h = bf16(bf16(sum_s P[s] + bias) + resid), y = LayerNorm(h) exactly as
// layernorm_bf16_kernel computes it.
__global__ void __launch_bounds__(320) layernorm_reduce_kernel(const float* __restrict__ P, int S, long split_stride,
                                                               long ldp, const bf16* __restrict__ bias, const bf16* resid,
                                                               long ldr, const bf16* __restrict__ w,
                                                               const bf16* __restrict__ b, bf16* h_out, long ldh,
                                                               bf16* __restrict__ y, long ldy, int width, float eps) {
  // One float4 of the row per thread when the row fits (width <= 4 * blockDim: the SigLIP tower's 1152 = 288 threads):
  // everything stays in registers.  Loads that do not depend on the preceding kernel (norm weight / bias, linear bias)
  // are issued before the dependency wait; then the residual and ALL partials of the element group are in flight at
  // once (up to 8 per round) and summed in split order - the same bits as a one-at-a-time loop.
  extern __shared__ float hrow[];
  __shared__ float red[32];
  const int row = blockIdx.x;
  const bool one = width <= static_cast<int>(blockDim.x) * 4;
  const int i0 = threadIdx.x * 4;
  float wreg[4] = {0.f, 0.f, 0.f, 0.f}, breg[4] = {0.f, 0.f, 0.f, 0.f}, lbias[4] = {0.f, 0.f, 0.f, 0.f};
  if (one && i0 < width) {
    const uint2 wv = *reinterpret_cast<const uint2*>(w + i0);
    const uint2 bv = *reinterpret_cast<const uint2*>(b + i0);
    const float2 g0 = unpack_bf16x2(wv.x), g1 = unpack_bf16x2(wv.y), h0 = unpack_bf16x2(bv.x), h1 = unpack_bf16x2(bv.y);
    wreg[0] = g0.x, wreg[1] = g0.y, wreg[2] = g1.x, wreg[3] = g1.y;
    breg[0] = h0.x, breg[1] = h0.y, breg[2] = h1.x, breg[3] = h1.y;
    if (bias != nullptr) {
      const uint2 v = *reinterpret_cast<const uint2*>(bias + i0);
      const float2 f0 = unpack_bf16x2(v.x), f1 = unpack_bf16x2(v.y);
      lbias[0] = f0.x, lbias[1] = f0.y, lbias[2] = f1.x, lbias[3] = f1.y;
    }
  }
  pdl_wait();
  pdl_launch();
  const float* pr = P + row * ldp;
  float s = 0.f;
  float hreg[4] = {0.f, 0.f, 0.f, 0.f};
  for (int i = i0; i < width; i += blockDim.x * 4) {
    const uint2 rv = *reinterpret_cast<const uint2*>(resid + row * ldr + i);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    int sp = 0;
    for (; sp + 8 <= S; sp += 8) {
      float4 a[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) a[q] = *reinterpret_cast<const float4*>(pr + (sp + q) * split_stride + i);
#pragma unroll
      for (int q = 0; q < 8; ++q) acc.x += a[q].x, acc.y += a[q].y, acc.z += a[q].z, acc.w += a[q].w;
    }
    if (sp + 4 <= S) {
      float4 a[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) a[q] = *reinterpret_cast<const float4*>(pr + (sp + q) * split_stride + i);
#pragma unroll
      for (int q = 0; q < 4; ++q) acc.x += a[q].x, acc.y += a[q].y, acc.z += a[q].z, acc.w += a[q].w;
      sp += 4;
    }
    for (; sp < S; ++sp) {
      const float4 a = *reinterpret_cast<const float4*>(pr + sp * split_stride + i);
      acc.x += a.x, acc.y += a.y, acc.z += a.z, acc.w += a.w;
    }
    float bb[4] = {lbias[0], lbias[1], lbias[2], lbias[3]};
    if (!one && bias != nullptr) {
      const uint2 v = *reinterpret_cast<const uint2*>(bias + i);
      const float2 f0 = unpack_bf16x2(v.x), f1 = unpack_bf16x2(v.y);
      bb[0] = f0.x, bb[1] = f0.y, bb[2] = f1.x, bb[3] = f1.y;
    }
    const float2 r0 = unpack_bf16x2(rv.x), r1 = unpack_bf16x2(rv.y);
    const float hv[4] = {bf16_round(bf16_round(acc.x + bb[0]) + r0.x), bf16_round(bf16_round(acc.y + bb[1]) + r0.y),
                         bf16_round(bf16_round(acc.z + bb[2]) + r1.x), bf16_round(bf16_round(acc.w + bb[3]) + r1.y)};
    *reinterpret_cast<uint2*>(h_out + row * ldh + i) = make_uint2(pack_bf16x2(hv[0], hv[1]), pack_bf16x2(hv[2], hv[3]));
    if (one) {
      hreg[0] = hv[0], hreg[1] = hv[1], hreg[2] = hv[2], hreg[3] = hv[3];
    } else {
      *reinterpret_cast<float4*>(hrow + i) = make_float4(hv[0], hv[1], hv[2], hv[3]);
    }
    s += hv[0] + hv[1] + hv[2] + hv[3];
  }
  const float mean = block_sum(s, red) / static_cast<float>(width);
  float vs = 0.f;
  if (one) {
    if (i0 < width)
      vs = (hreg[0] - mean) * (hreg[0] - mean) + (hreg[1] - mean) * (hreg[1] - mean) + (hreg[2] - mean) * (hreg[2] - mean) +
           (hreg[3] - mean) * (hreg[3] - mean);
  } else {
    for (int i = i0; i < width; i += blockDim.x * 4) {
      const float4 f = *reinterpret_cast<const float4*>(hrow + i);
      vs += (f.x - mean) * (f.x - mean) + (f.y - mean) * (f.y - mean) + (f.z - mean) * (f.z - mean) +
            (f.w - mean) * (f.w - mean);
    }
  }
  const float var = block_sum(vs, red) / static_cast<float>(width);
  const float rstd = 1.0f / sqrtf(var + eps);
  bf16* yr = y + row * ldy;
  if (one) {
    if (i0 < width)
      *reinterpret_cast<uint2*>(yr + i0) =
          make_uint2(pack_bf16x2((hreg[0] - mean) * rstd * wreg[0] + breg[0], (hreg[1] - mean) * rstd * wreg[1] + breg[1]),
                     pack_bf16x2((hreg[2] - mean) * rstd * wreg[2] + breg[2], (hreg[3] - mean) * rstd * wreg[3] + breg[3]));
    return;
  }
  for (int i = i0; i < width; i += blockDim.x * 4) {
    const float4 f = *reinterpret_cast<const float4*>(hrow + i);
    const uint2 wv = *reinterpret_cast<const uint2*>(w + i);
    const uint2 bv = *reinterpret_cast<const uint2*>(b + i);
    const float2 g0 = unpack_bf16x2(wv.x), g1 = unpack_bf16x2(wv.y), h0 = unpack_bf16x2(bv.x), h1 = unpack_bf16x2(bv.y);
    *reinterpret_cast<uint2*>(yr + i) =
        make_uint2(pack_bf16x2((f.x - mean) * rstd * g0.x + h0.x, (f.y - mean) * rstd * g0.y + h0.y),
                   pack_bf16x2((f.z - mean) * rstd * g1.x + h1.x, (f.w - mean) * rstd * g1.y + h1.y));
  }
}

int layernorm_reduce(cudaStream_t st, const float* P, int S, long split_stride, long ldp, const bf16* bias,
                     const bf16* resid, long ldr, const bf16* w, const bf16* b, bf16* h_out, long ldh, bf16* y, long ldy,
                     int rows, int width, float eps) {
  CVB_REQUIRE(width % 4 == 0 && ldp % 4 == 0 && split_stride % 4 == 0 && ldr % 4 == 0 && ldh % 4 == 0 && ldy % 4 == 0,
              "layernorm_reduce needs 4-element aligned rows");
  CVB_REQUIRE(S >= 1 && resid != nullptr, "layernorm_reduce needs >= 1 partial and a residual");
  const size_t smem = static_cast<size_t>(width) * sizeof(float);
  // one float4 per thread up to 1280 columns (whole warps), 256 threads striding over wider rows
  const int threads = width <= 1280 ? std::max(32, (width / 4 + 31) / 32 * 32) : 256;
  CVB_TRY(launch_pdl(layernorm_reduce_kernel, dim3(rows), dim3(threads), smem, st, 1, P, S, split_stride, ldp, bias, resid, ldr,
                     w, b, h_out, ldh, y, ldy, width, eps));
  CVB_LAUNCHED();
  return 0;
}

int layernorm_bf16(cudaStream_t st, const bf16* x, long ldx, const bf16* w, const bf16* b, bf16* y,
                   long ldy, int rows, int width, float eps) {
  CVB_REQUIRE(width % 8 == 0, "layernorm width must be a multiple of 8");
  const int threads = width >= 1024 ? 128 : 64;
  CVB_TRY(launch_pdl(layernorm_bf16_kernel, dim3(rows), dim3(threads), 0, st, 1, x, ldx, w, b, y, ldy, width, eps));
  CVB_LAUNCHED();
  return 0;
}

// LayerNorm over fp32 rows with optional residual added first (verifier heads: post-norm layers).
__global__ void __launch_bounds__(128) layernorm_f32_kernel(const float* __restrict__ x,
                                                            const float* __restrict__ resid,
                                                            const float* __restrict__ w,
                                                            const float* __restrict__ b,
                                                            float* __restrict__ y, int width,
                                                            float eps, bf16* __restrict__ y_split, int relu) {
  pdl_wait();
  pdl_launch();
  __shared__ float red[32];
  extern __shared__ float rowbuf[];
  const long off = static_cast<long>(blockIdx.x) * width;
  float s = 0.f;
  for (int i = threadIdx.x; i < width; i += blockDim.x) {
    float v = x[off + i];
    if (resid != nullptr) v += resid[off + i];
    rowbuf[i] = v;
    s += v;
  }
  const float mean = block_sum(s, red) / static_cast<float>(width);
  float vs = 0.f;
  for (int i = threadIdx.x; i < width; i += blockDim.x) {
    const float d = rowbuf[i] - mean;
    vs += d * d;
  }
  const float var = block_sum(vs, red) / static_cast<float>(width);
  const float rstd = 1.0f / sqrtf(var + eps);
  bf16* ys = y_split != nullptr ? y_split + static_cast<long>(blockIdx.x) * 3 * width : nullptr;
  for (int i = threadIdx.x; i < width; i += blockDim.x) {
    float v = (rowbuf[i] - mean) * rstd * w[i] + b[i];
    if (relu) v = fmaxf(v, 0.f);  // nn.Sequential(Linear, LayerNorm, ReLU, ...) of the MLP action encoder
    y[off + i] = v;
    if (ys != nullptr) {  // [hi | hi | lo]: the A operand of the 3-term bf16 GEMM (split3_rows)
      const bf16 hi = __float2bfloat16_rn(v);
      const bf16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
      ys[i] = hi, ys[width + i] = hi, ys[2 * width + i] = lo;
    }
  }
}

int layernorm_f32(cudaStream_t st, const float* x, const float* resid, const float* w,
                  const float* b, float* y, int rows, int width, float eps, bf16* y_split, int relu) {
  CVB_TRY(launch_pdl(layernorm_f32_kernel, dim3(rows), dim3(128), width * sizeof(float), st, 1, x, resid, w, b, y, width, eps,
                     y_split, relu));
  CVB_LAUNCHED();
  return 0;
}

// ------------------------------------------------------------------------------------------------
// fp32-accurate GEMMs on the bf16 tensor cores (round 2): x = hi + lo + O(2^-17 |x|) with hi = bf16(x), lo = bf16(x - hi),
// so  a . w = a_hi w_hi + a_hi w_lo + a_lo w_hi + O(2^-16 |a||w|)  - three bf16 products accumulated in fp32 by ONE
// tcgen05 GEMM over a K axis of length 3K:  A' = [a_hi | a_hi | a_lo],  W' = [w_hi | w_lo | w_hi].  The dropped terms
// are below the fp32 rounding of a K = 512 .. 1024 dot product accumulated in another order, two orders of magnitude
// under the score tolerance.  Replaces the fp32 SIMT GEMMs of the verifier's trajectory encoder
// (efficient_ensemble_merged.py:229-245) and of action_time_mlp_out (modeling_pi0.py:607-609): 20-48 us each.
// mode 0: activation layout [hi | hi | lo]; mode 1: weight layout [hi | lo | hi]; act 1: ReLU applied first.
__global__ void __launch_bounds__(256) split3_rows_kernel(const float* __restrict__ x, long ldx, bf16* __restrict__ out,
                                                          int K, int mode, int act, long total) {
  pdl_wait();
  pdl_launch();
  const long i = (static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x) * 4;
  if (i >= total) return;
  const long r = i / K;
  const int k = static_cast<int>(i % K);  // K % 4 == 0: the four elements share a row
  const float4 v4 = *reinterpret_cast<const float4*>(x + r * ldx + k);
  float v[4] = {v4.x, v4.y, v4.z, v4.w};
  uint32_t hi2[2], lo2[2];
#pragma unroll
  for (int e = 0; e < 2; ++e) {
    float a = v[2 * e], b = v[2 * e + 1];
    if (act == 1) a = fmaxf(a, 0.f), b = fmaxf(b, 0.f);
    const float ah = bf16_round(a), bh = bf16_round(b);
    hi2[e] = pack_bf16x2(ah, bh);
    lo2[e] = pack_bf16x2(a - ah, b - bh);
  }
  bf16* o = out + r * 3 * K + k;
  const uint2 H = make_uint2(hi2[0], hi2[1]), L = make_uint2(lo2[0], lo2[1]);
  *reinterpret_cast<uint2*>(o) = H;
  *reinterpret_cast<uint2*>(o + K) = mode == 0 ? H : L;
  *reinterpret_cast<uint2*>(o + 2 * K) = mode == 0 ? L : H;
}

int split3_rows(cudaStream_t st, const float* x, long ldx, bf16* out, long rows, int K, int mode, int act) {
  CVB_REQUIRE(K % 4 == 0 && ldx % 4 == 0, "split3_rows needs 4-element aligned rows");
  const long total = rows * K;
  CVB_TRY(launch_pdl(split3_rows_kernel, dim3(static_cast<unsigned>((total / 4 + 255) / 256)), dim3(256), 0, st, 1, x, ldx, out, K,
                     mode, act, total));
  CVB_LAUNCHED();
  return 0;
}

// out[map(m)][n] = sum_s P[s][m][n] (split order) + bias[n]; map = the suffix layout of sgemm's out_group (0 = identity)
__global__ void __launch_bounds__(256) partial_reduce_f32_kernel(const float* __restrict__ P, int S, long split_stride, long ldp,
                                                                 const float* __restrict__ bias, float* __restrict__ out,
                                                                 long ldo, int N, int out_group, long total) {
  pdl_wait();
  pdl_launch();
  const long i = (static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x) * 4;
  if (i >= total) return;
  const long m = i / N;
  const int n = static_cast<int>(i % N);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  const float* p = P + m * ldp + n;
  int s = 0;
  for (; s + 8 <= S; s += 8) {
    float4 a[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] = *reinterpret_cast<const float4*>(p + (s + j) * split_stride);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc.x += a[j].x, acc.y += a[j].y, acc.z += a[j].z, acc.w += a[j].w;
  }
  for (; s < S; ++s) {
    const float4 a = *reinterpret_cast<const float4*>(p + s * split_stride);
    acc.x += a.x, acc.y += a.y, acc.z += a.z, acc.w += a.w;
  }
  if (bias != nullptr) {
    const float4 b = *reinterpret_cast<const float4*>(bias + n);
    acc.x += b.x, acc.y += b.y, acc.z += b.z, acc.w += b.w;
  }
  const long mo = out_group > 0 ? (m / out_group) * (out_group + 1) + 1 + m % out_group : m;
  *reinterpret_cast<float4*>(out + mo * ldo + n) = acc;
}

int partial_reduce_f32(cudaStream_t st, const float* P, int S, long split_stride, long ldp, const float* bias, float* out,
                       long ldo, int rows, int N, int out_group) {
  CVB_REQUIRE(N % 4 == 0 && ldp % 4 == 0 && ldo % 4 == 0 && split_stride % 4 == 0, "partial_reduce_f32 needs 4-element aligned rows");
  const long total = static_cast<long>(rows) * N;
  CVB_TRY(launch_pdl(partial_reduce_f32_kernel, dim3(static_cast<unsigned>((total / 4 + 255) / 256)), dim3(256), 0, st, 1, P, S,
                     split_stride, ldp, bias, out, ldo, N, out_group, total));
  CVB_LAUNCHED();
  return 0;
}

}  // namespace cvb
