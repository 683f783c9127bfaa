// Fused elementwise / normalisation kernels of the pi0 path.  All of them are HBM- or latency-bound:
// 16-byte vectorised, coalesced accesses; one CTA per row (or per small group of rows); the rounding
// points follow SURVEY.md Appendix A.
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
  __syncthreads();
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

int rmsnorm(cudaStream_t st, const void* x, int x_is_f32, long ldx, const void* w, int w_is_f32,
            bf16* y, long ldy, int rows, int width, float eps, const int* rows_dev) {
  CVB_REQUIRE(width % 8 == 0, "rmsnorm width must be a multiple of 8");
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
                                                            float eps) {
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
  for (int i = threadIdx.x; i < width; i += blockDim.x)
    y[off + i] = (rowbuf[i] - mean) * rstd * w[i] + b[i];
}

int layernorm_f32(cudaStream_t st, const float* x, const float* resid, const float* w,
                  const float* b, float* y, int rows, int width, float eps) {
  CVB_TRY(launch_pdl(layernorm_f32_kernel, dim3(rows), dim3(128), width * sizeof(float), st, 1, x, resid, w, b, y, width, eps));
  CVB_LAUNCHED();
  return 0;
}

}  // namespace cvb
