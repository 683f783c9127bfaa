// fp32 SIMT GEMM for the parts of the path the reference keeps in float32:
// the pi0 suffix embedding MLP / action_out_proj (modeling_pi0.py:577-609,751) and the verifier heads
// (bridge_verifier/ensemble_eval/efficient_ensemble_merged.py:194-247).  The score tolerance is 1e-3
// relative, so these stay true-fp32 FFMA (no TF32): 64x64 tile, 16-deep k slices, 4x4 per thread.
#include "host_common.h"
#include "ops.h"
#include "ptx.cuh"

namespace cvb {

namespace {

constexpr int TM = 64, TN = 64, TK = 16;

__device__ __forceinline__ float act_apply(float v, int act) {
  switch (act) {
    case SACT_RELU:
      return v > 0.f ? v : 0.f;
    case SACT_GELU_ERF:
      return 0.5f * v * (1.0f + erff(v * 0.70710678118654752440f));
    case SACT_SILU:
      return v / (1.0f + expf(-v));
    default:
      return v;
  }
}

__global__ void __launch_bounds__(256) sgemm_kernel(const float* __restrict__ A, long lda,
                                                    const float* __restrict__ W, long ldw, int M,
                                                    int N, int K, float* __restrict__ C, long ldc,
                                                    const float* __restrict__ bias,
                                                    const float* __restrict__ row_bias,
                                                    const float* __restrict__ resid, long ldr,
                                                    int act, int out_group, int w_kn) {
  pdl_wait();
  pdl_launch();
  __shared__ float As[TK][TM + 4];
  __shared__ float Ws[TK][TN + 4];
  const int m0 = blockIdx.y * TM, n0 = blockIdx.x * TN;
  const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
  float acc[4][4] = {};
  const bool vec_ok = (K % 4 == 0) && (lda % 4 == 0) && (ldw % 4 == 0) &&
                      ((reinterpret_cast<uintptr_t>(A) & 15) == 0) &&
                      ((reinterpret_cast<uintptr_t>(W) & 15) == 0);
  const int lr = threadIdx.x / 4;        // 0..63: tile row
  const int lc = (threadIdx.x % 4) * 4;  // 0,4,8,12: k offset
  for (int k0 = 0; k0 < K; k0 += TK) {
    float av[4] = {0.f, 0.f, 0.f, 0.f}, wv[4] = {0.f, 0.f, 0.f, 0.f};
    const int am = m0 + lr, wn = n0 + lr, kk = k0 + lc;
    if (am < M) {
      if (vec_ok && kk + 3 < K) {
        const float4 t = *reinterpret_cast<const float4*>(A + am * lda + kk);
        av[0] = t.x, av[1] = t.y, av[2] = t.z, av[3] = t.w;
      } else {
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (kk + e < K) av[e] = A[am * lda + kk + e];
      }
    }
    if (w_kn) {
      // W given as [K, N] row-major (plain A @ B): this thread loads 4 consecutive n of one k
      const int wk = k0 + threadIdx.x / 16, wn4 = n0 + (threadIdx.x % 16) * 4;
#pragma unroll
      for (int e = 0; e < 4; ++e)
        if (wk < K && wn4 + e < N) wv[e] = W[wk * ldw + wn4 + e];
    } else if (wn < N) {
      if (vec_ok && kk + 3 < K) {
        const float4 t = *reinterpret_cast<const float4*>(W + wn * ldw + kk);
        wv[0] = t.x, wv[1] = t.y, wv[2] = t.z, wv[3] = t.w;
      } else {
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (kk + e < K) wv[e] = W[wn * ldw + kk + e];
      }
    }
    __syncthreads();
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      As[lc + e][lr] = av[e];
      if (w_kn)
        Ws[threadIdx.x / 16][(threadIdx.x % 16) * 4 + e] = wv[e];
      else
        Ws[lc + e][lr] = wv[e];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < TK; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 w = *reinterpret_cast<const float4*>(&Ws[k][tx * 4]);
      const float ar[4] = {a.x, a.y, a.z, a.w}, wr[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], wr[j], acc[i][j]);
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      float v = acc[i][j];
      if (bias != nullptr) v += bias[n];
      if (row_bias != nullptr) v += row_bias[n];
      v = act_apply(v, act);
      if (resid != nullptr) v += resid[m * ldr + n];
      const long mo = out_group > 0 ? (m / out_group) * (out_group + 1) + 1 + m % out_group : m;
      C[mo * ldc + n] = v;
    }
  }
}

}  // namespace

int sgemm_f32(cudaStream_t st, const SgemmCall& c) {
  CVB_REQUIRE(c.M > 0 && c.N > 0 && c.K > 0, "empty sgemm");
  dim3 grid((c.N + TN - 1) / TN, (c.M + TM - 1) / TM);
  CVB_TRY(launch_pdl(sgemm_kernel, dim3(grid), dim3(256), 0, st, 1, c.A, c.lda, c.W, c.ldw, c.M, c.N, c.K, c.C, c.ldc, c.bias,
                                     c.row_bias, c.resid, c.ldr, c.act, c.out_group, c.w_kn));
  CVB_LAUNCHED();
  return 0;
}

}  // namespace cvb
