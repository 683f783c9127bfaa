// fp32 SIMT GEMM for the parts of the path the reference keeps in float32:
// the pi0 suffix embedding MLP / action_out_proj (modeling_pi0.py:577-609,751) and the verifier heads
// (bridge_verifier/ensemble_eval/efficient_ensemble_merged.py:194-247).  The score tolerance is 1e-3
// relative, so these stay true-fp32 FFMA (no TF32): 64x64 tile, 16-deep k slices, 4x4 per thread.
#include <cstdlib>

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

// ---- pipelined variant (the default whenever rows are 16-byte addressable) --------------------------------------
// The 64x64 kernel above pays one exposed global-load latency per 16-deep k slice and gives the denoise-loop shapes
// (M = 160, N = K = 1024) only 48 CTAs: 57 us per call.  Here: PTM x PTN tiles picked so the grid covers the SMs, a
// 4-stage cp.async ring of 32-deep k slices, weight slices prefetched before the programmatic-dependency wait.
constexpr int PK = 32, PSTAGES = 4, PLD = PK + 4;  // row stride 36 floats: 16-byte aligned, conflict-free LDS.128

__device__ __forceinline__ void cp_async16_zfill(void* smem_dst, const void* gsrc, bool valid) {
  const uint32_t d = static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst));
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit_group() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait_group_n() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

template <int PTM, int PTN>
__global__ void __launch_bounds__(128) sgemm_pipe_kernel(const float* __restrict__ A, long lda,
                                                         const float* __restrict__ W, long ldw, int M, int N, int K,
                                                         float* __restrict__ C, long ldc,
                                                         const float* __restrict__ bias,
                                                         const float* __restrict__ row_bias,
                                                         const float* __restrict__ resid, long ldr, int act,
                                                         int out_group, int w_dynamic) {
  extern __shared__ __align__(16) float sg_smem[];
  float* As = sg_smem;                                // [PSTAGES][PTM][PLD]
  float* Ws = sg_smem + PSTAGES * PTM * PLD;          // [PSTAGES][PTN][PLD]
  constexpr int RI = PTM / 16, RJ = PTN / 8;
  const int m0 = blockIdx.y * PTM, n0 = blockIdx.x * PTN;
  const int tx = threadIdx.x & 7, ty = threadIdx.x >> 3;
  const int nk = (K + PK - 1) / PK;
  auto load_rows = [&](float* dst, const float* src, long ld, int r0, int rmax, int rows, int kb) {
    // rows x 8 chunks of 16 bytes
    for (int c = threadIdx.x; c < rows * (PK / 4); c += 128) {
      const int r = c >> 3, kc = (c & 7) * 4, gk = kb * PK + kc;
      const bool ok = (r0 + r < rmax) && (gk < K);  // K % 4 == 0: a chunk is entirely inside or outside
      const float* g = ok ? src + static_cast<long>(r0 + r) * ld + gk : src;
      cp_async16_zfill(dst + r * PLD + kc, g, ok);
    }
  };
  // weights do not depend on the preceding kernel
  const int pre = min(nk, PSTAGES - 1);
  if (!w_dynamic)
    for (int s = 0; s < pre; ++s) load_rows(Ws + s * PTN * PLD, W, ldw, n0, N, PTN, s);
  pdl_wait();
  pdl_launch();
  if (w_dynamic)
    for (int s = 0; s < pre; ++s) load_rows(Ws + s * PTN * PLD, W, ldw, n0, N, PTN, s);
  for (int s = 0; s < PSTAGES - 1; ++s) {
    if (s < nk) load_rows(As + s * PTM * PLD, A, lda, m0, M, PTM, s);
    cp_async_commit_group();
  }
  float acc[RI][RJ];
#pragma unroll
  for (int i = 0; i < RI; ++i)
#pragma unroll
    for (int j = 0; j < RJ; ++j) acc[i][j] = 0.f;
  for (int kb = 0; kb < nk; ++kb) {
    cp_async_wait_group_n<PSTAGES - 2>();
    __syncthreads();
    {  // refill the slot consumed in the previous iteration
      const int nx = kb + PSTAGES - 1;
      if (nx < nk) {
        const int s = nx % PSTAGES;
        load_rows(Ws + s * PTN * PLD, W, ldw, n0, N, PTN, nx);
        load_rows(As + s * PTM * PLD, A, lda, m0, M, PTM, nx);
      }
      cp_async_commit_group();
    }
    const float* as = As + (kb % PSTAGES) * PTM * PLD;
    const float* ws = Ws + (kb % PSTAGES) * PTN * PLD;
#pragma unroll
    for (int k4 = 0; k4 < PK / 4; ++k4) {
      float4 a[RI], w[RJ];
#pragma unroll
      for (int i = 0; i < RI; ++i) a[i] = *reinterpret_cast<const float4*>(as + (ty + 16 * i) * PLD + k4 * 4);
#pragma unroll
      for (int j = 0; j < RJ; ++j) w[j] = *reinterpret_cast<const float4*>(ws + (tx + 8 * j) * PLD + k4 * 4);
#pragma unroll
      for (int i = 0; i < RI; ++i)
#pragma unroll
        for (int j = 0; j < RJ; ++j) {
          acc[i][j] = fmaf(a[i].x, w[j].x, acc[i][j]);
          acc[i][j] = fmaf(a[i].y, w[j].y, acc[i][j]);
          acc[i][j] = fmaf(a[i].z, w[j].z, acc[i][j]);
          acc[i][j] = fmaf(a[i].w, w[j].w, acc[i][j]);
        }
    }
  }
  cp_async_wait_group_n<0>();
#pragma unroll
  for (int i = 0; i < RI; ++i) {
    const int m = m0 + ty + 16 * i;
    if (m >= M) continue;
    const long mo = out_group > 0 ? (m / out_group) * (out_group + 1) + 1 + m % out_group : m;
#pragma unroll
    for (int j = 0; j < RJ; ++j) {
      const int n = n0 + tx + 8 * j;
      if (n >= N) continue;
      float v = acc[i][j];
      if (bias != nullptr) v += bias[n];
      if (row_bias != nullptr) v += row_bias[n];
      v = act_apply(v, act);
      if (resid != nullptr) v += resid[m * ldr + n];
      C[mo * ldc + n] = v;
    }
  }
}

template <int PTM, int PTN>
int launch_pipe(cudaStream_t st, const SgemmCall& c) {
  constexpr int smem = PSTAGES * (PTM + PTN) * PLD * static_cast<int>(sizeof(float));
  auto kern = sgemm_pipe_kernel<PTM, PTN>;
  if (smem > 48 * 1024) CVB_TRY(ensure_dyn_smem(kern, smem));
  dim3 grid((c.N + PTN - 1) / PTN, (c.M + PTM - 1) / PTM);
  CVB_TRY(launch_pdl(kern, grid, dim3(128), smem, st, 1, c.A, c.lda, c.W, c.ldw, c.M, c.N, c.K, c.C, c.ldc, c.bias,
                     c.row_bias, c.resid, c.ldr, c.act, c.out_group, c.w_dynamic));
  CVB_LAUNCHED();
  return 0;
}

}  // namespace

int sgemm_f32(cudaStream_t st, const SgemmCall& c) {
  CVB_REQUIRE(c.M > 0 && c.N > 0 && c.K > 0, "empty sgemm");
  const bool vec_ok = (c.K % 4 == 0) && (c.lda % 4 == 0) && (c.ldw % 4 == 0) &&
                      ((reinterpret_cast<uintptr_t>(c.A) & 15) == 0) && ((reinterpret_cast<uintptr_t>(c.W) & 15) == 0);
  if (vec_ok && !c.w_kn && c.K >= 64) {
    // largest tile that still gives every SM about two CTAs
    const long t64 = static_cast<long>((c.M + 63) / 64) * ((c.N + 63) / 64);
    const long t32x64 = static_cast<long>((c.M + 31) / 32) * ((c.N + 63) / 64);
    // (the 64 x 64 tile does 10.7 FMAs per shared-memory load, the 32 x 64 one 8, the 32 x 32 one 5.3: take the big tile
    // when every SM gets about two CTAs; measured: a lower bar (CVB_SGEMM_T64_MIN=148) changes nothing in the step)
    static const long t64_min = getenv("CVB_SGEMM_T64_MIN") != nullptr ? atol(getenv("CVB_SGEMM_T64_MIN")) : 296;
    if (t64 >= t64_min) return launch_pipe<64, 64>(st, c);
    if (t32x64 >= 296) return launch_pipe<32, 64>(st, c);
    return launch_pipe<32, 32>(st, c);
  }
  dim3 grid((c.N + TN - 1) / TN, (c.M + TM - 1) / TM);
  CVB_TRY(launch_pdl(sgemm_kernel, dim3(grid), dim3(256), 0, st, 1, c.A, c.lda, c.W, c.ldw, c.M, c.N, c.K, c.C, c.ldc, c.bias,
                                     c.row_bias, c.resid, c.ldr, c.act, c.out_group, c.w_kn));
  CVB_LAUNCHED();
  return 0;
}

}  // namespace cvb
