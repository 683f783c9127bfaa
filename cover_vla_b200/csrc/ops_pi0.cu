// pi0-specific fused kernels: patch im2col, prefix assembly (image-embedding rescale + token gather),
// RoPE + KV-cache write, action_out_proj + Euler update.  See pi0_kernels.h for the reference lines.
#include "host_common.h"
#include "pi0_kernels.h"
#include "ptx.cuh"

namespace cvb {

// ------------------------------------------------------------------------------------------------
// conv(patch, stride patch) as GEMM: patches[t, c*P*P + ky*P + kx] = bf16(img[c, py*P+ky, px*P+kx])
__global__ void im2col_kernel(const float* __restrict__ img, bf16* __restrict__ out, int C, int H,
                              int W, int P, int kpad) {
  pdl_wait();
  pdl_launch();
  const int t = blockIdx.x;
  const int gw = W / P;
  const int py = t / gw, px = t % gw;
  const int kreal = C * P * P;
  img += static_cast<long>(blockIdx.y) * C * H * W;  // image of this observation
  out += static_cast<long>(blockIdx.y) * gridDim.x * kpad;
  for (int k = threadIdx.x; k < kpad; k += blockDim.x) {
    float v = 0.f;
    if (k < kreal) {
      const int c = k / (P * P), rem = k % (P * P);
      const int ky = rem / P, kx = rem % P;
      v = img[(static_cast<long>(c) * H + py * P + ky) * W + px * P + kx];
    }
    out[static_cast<long>(t) * kpad + k] = __float2bfloat16_rn(v);
  }
}

int im2col_patches(cudaStream_t st, const float* img, bf16* out, int C, int H, int W, int P,
                   int kpad, int images) {
  const int tokens = (H / P) * (W / P);
  CVB_TRY(launch_pdl(im2col_kernel, dim3(tokens, images), dim3(128), 0, st, 1, img, out, C, H, W, P, kpad));
  CVB_LAUNCHED();
  return 0;
}

// ------------------------------------------------------------------------------------------------
// prefix[r, t, :] = t < n_img ? bf16(bf16(proj[t,:] / sqrt(D)) * bf16(sqrt(D)))       (modeling_pi0.py:533-538
//                             : bf16(embed[tok[r, t-n_img], :] * sqrt(D))               + HF get_image_features, :549-553)
__global__ void build_prefix_kernel(const bf16* __restrict__ proj, const bf16* __restrict__ embed,
                                    const int64_t* __restrict__ tok, bf16* __restrict__ prefix,
                                    int n_img, int n_lang, int tok_stride, int D, float sqrt_d, float sqrt_d_bf16,
                                    int rpo) {
  pdl_wait();
  pdl_launch();
  const int t = blockIdx.x, r = blockIdx.y;
  const int P = n_img + n_lang;
  bf16* dst = prefix + (static_cast<long>(r) * P + t) * D;
  if (t < n_img) {
    // image rows of the observation this rephrase belongs to (rpo rephrases per observation; 0 = one observation)
    const bf16* src = proj + (static_cast<long>(rpo > 0 ? r / rpo : 0) * n_img + t) * D;
    for (int i = threadIdx.x * 8; i < D; i += blockDim.x * 8) {
      const uint4 v = *reinterpret_cast<const uint4*>(src + i);
      const uint32_t u[4] = {v.x, v.y, v.z, v.w};
      uint32_t o[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f = unpack_bf16x2(u[e]);
        o[e] = pack_bf16x2(bf16_round(f.x / sqrt_d) * sqrt_d_bf16, bf16_round(f.y / sqrt_d) * sqrt_d_bf16);
      }
      *reinterpret_cast<uint4*>(dst + i) = make_uint4(o[0], o[1], o[2], o[3]);
    }
  } else {
    const int64_t id = tok[static_cast<long>(r) * tok_stride + (t - n_img)];
    const bf16* src = embed + id * D;
    for (int i = threadIdx.x * 8; i < D; i += blockDim.x * 8) {
      const uint4 v = *reinterpret_cast<const uint4*>(src + i);
      const uint32_t u[4] = {v.x, v.y, v.z, v.w};
      uint32_t o[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f = unpack_bf16x2(u[e]);
        o[e] = pack_bf16x2(f.x * sqrt_d, f.y * sqrt_d);
      }
      *reinterpret_cast<uint4*>(dst + i) = make_uint4(o[0], o[1], o[2], o[3]);
    }
  }
}

int build_prefix(cudaStream_t st, const bf16* proj, const bf16* embed, const int64_t* tok,
                 bf16* prefix, int R, int n_img, int n_lang, int tok_stride, int D, int rephrases_per_obs) {
  CVB_REQUIRE(D % 8 == 0, "lm width must be a multiple of 8");
  const float s = static_cast<float>(sqrt(static_cast<double>(D)));  // (float)(D ** 0.5)
  const float sb = __bfloat162float(__float2bfloat16_rn(s));
  dim3 grid(n_img + n_lang, R);
  CVB_TRY(launch_pdl(build_prefix_kernel, dim3(grid), dim3(128), 0, st, 1, proj, embed, tok, prefix, n_img, n_lang, tok_stride, D, s, sb,
                     rephrases_per_obs));
  CVB_LAUNCHED();
  return 0;
}

// ------------------------------------------------------------------------------------------------
// RoPE (paligemma_with_expert.py:34-57) in place on the q heads and the k head of a fused
// [rows, (heads + 2) * hd] qkv buffer; optionally scatters post-RoPE K and V into the KV cache
// (paligemma_with_expert.py:294-299).
__global__ void rope_kernel(bf16* __restrict__ qkv, long ld, const float* __restrict__ timescale,
                            int heads, int hd, int rows_per_batch, const int* __restrict__ pos_base_dev,
                            int q_per_kv_batch, bf16* __restrict__ kcache, bf16* __restrict__ vcache,
                            long cache_bs, long cache_rs, bf16* __restrict__ vt, long vt_bs, long vt_ld) {
  pdl_wait();
  pdl_launch();
  const int row = blockIdx.x;
  const int b = row / rows_per_batch, t = row % rows_per_batch;
  const int base = pos_base_dev != nullptr ? pos_base_dev[b / q_per_kv_batch] : 0;
  const float pos = static_cast<float>(base + t);
  const int half = hd / 2;
  const int chunks = half / 8;
  bf16* rp = qkv + static_cast<long>(row) * ld;
  for (int w = threadIdx.x; w < (heads + 1) * chunks; w += blockDim.x) {
    const int h = w / chunks, c = (w % chunks) * 8;
    bf16* x1p = rp + h * hd + c;
    bf16* x2p = x1p + half;
    const uint4 v1 = *reinterpret_cast<const uint4*>(x1p);
    const uint4 v2 = *reinterpret_cast<const uint4*>(x2p);
    const uint32_t u1[4] = {v1.x, v1.y, v1.z, v1.w}, u2[4] = {v2.x, v2.y, v2.z, v2.w};
    uint32_t o1[4], o2[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 a = unpack_bf16x2(u1[e]), bb = unpack_bf16x2(u2[e]);
      float s0, c0, s1, c1;
      sincosf(pos / timescale[c + 2 * e], &s0, &c0);
      sincosf(pos / timescale[c + 2 * e + 1], &s1, &c1);
      // separate mul / sub as torch does (no fma contraction)
      o1[e] = pack_bf16x2(__fsub_rn(__fmul_rn(a.x, c0), __fmul_rn(bb.x, s0)),
                          __fsub_rn(__fmul_rn(a.y, c1), __fmul_rn(bb.y, s1)));
      o2[e] = pack_bf16x2(__fadd_rn(__fmul_rn(bb.x, c0), __fmul_rn(a.x, s0)),
                          __fadd_rn(__fmul_rn(bb.y, c1), __fmul_rn(a.y, s1)));
    }
    const uint4 r1 = make_uint4(o1[0], o1[1], o1[2], o1[3]);
    const uint4 r2 = make_uint4(o2[0], o2[1], o2[2], o2[3]);
    *reinterpret_cast<uint4*>(x1p) = r1;
    *reinterpret_cast<uint4*>(x2p) = r2;
    if (kcache != nullptr && h == heads) {
      bf16* kc = kcache + b * cache_bs + t * cache_rs;
      *reinterpret_cast<uint4*>(kc + c) = r1;
      *reinterpret_cast<uint4*>(kc + half + c) = r2;
    }
  }
  if (vcache != nullptr) {
    const bf16* vp = rp + (heads + 1) * hd;
    bf16* vc = vcache + b * cache_bs + t * cache_rs;
    for (int i = threadIdx.x * 8; i < hd; i += blockDim.x * 8)
      *reinterpret_cast<uint4*>(vc + i) = *reinterpret_cast<const uint4*>(vp + i);
  }
  if (vt != nullptr) {  // V^T[b][d][t]: the K-major "B" operand of the tcgen05 P.V GEMM (ops_attention_umma.cu)
    const bf16* vp = rp + (heads + 1) * hd;
    bf16* vtp = vt + b * vt_bs + t;
    for (int d = threadIdx.x; d < hd; d += blockDim.x) vtp[d * vt_ld] = vp[d];
  }
}

int rope_qkv(cudaStream_t st, bf16* qkv, long ld, const float* timescale, int rows, int heads,
             int hd, int rows_per_batch, const int* pos_base_dev, int q_per_kv_batch, bf16* kcache,
             bf16* vcache, long cache_bs, long cache_rs, bf16* vt, long vt_bs, long vt_ld) {
  CVB_REQUIRE(hd % 16 == 0, "head_dim must be a multiple of 16 for the vectorised RoPE");
  const int work = (heads + 1) * (hd / 16);
  int threads = ((work + 31) / 32) * 32;
  if (threads > 256) threads = 256;
  if (threads < 32) threads = 32;
  CVB_TRY(launch_pdl(rope_kernel, dim3(rows), dim3(threads), 0, st, 1, qkv, ld, timescale, heads, hd, rows_per_batch, pos_base_dev,
                                        q_per_kv_batch, kcache, vcache, cache_bs, cache_rs, vt, vt_bs, vt_ld));
  CVB_LAUNCHED();
  return 0;
}

// (cos, sin) of the suffix positions: tab[b][t][i] = sincos((pos_base[b] + t) / timescale[i]) - computed once per sample
// so the denoise attention applies RoPE while staging Q / K (same sincosf as rope_kernel: bit-identical rotation)
__global__ void rope_table_kernel(const float* __restrict__ timescale, const int* __restrict__ pos_base_dev,
                                  int tq, int half, float2* __restrict__ tab) {
  pdl_wait();
  pdl_launch();
  const int b = blockIdx.x / tq, t = blockIdx.x % tq;
  const float pos = static_cast<float>(pos_base_dev[b] + t);
  for (int i = threadIdx.x; i < half; i += blockDim.x) {
    float sn, cs;
    sincosf(pos / timescale[i], &sn, &cs);
    tab[(static_cast<long>(b) * tq + t) * half + i] = make_float2(cs, sn);
  }
}

int rope_table(cudaStream_t st, const float* timescale, const int* pos_base_dev, int batches, int tq, int half,
               float2* tab) {
  CVB_TRY(launch_pdl(rope_table_kernel, dim3(batches * tq), dim3(128), 0, st, 1, timescale, pos_base_dev, tq, half, tab));
  CVB_LAUNCHED();
  return 0;
}

// ------------------------------------------------------------------------------------------------
// v = action_out_proj(float(h_norm[last chunk tokens]));  x_t += dt * v      (modeling_pi0.py:748-752,713)
__global__ void __launch_bounds__(256) action_out_euler_kernel(const bf16* __restrict__ hn, long ld,
                                                               const float* __restrict__ w,
                                                               const float* __restrict__ bias,
                                                               float* __restrict__ x_t,
                                                               float* __restrict__ v_out, int width,
                                                               int adim, int chunk, int suffix_len,
                                                               float dt) {
  pdl_wait();
  pdl_launch();
  // one CTA per (candidate, chunk step): the row is staged once in shared memory (fp32), then 8 threads per output
  // stream its weight row with independent 16-byte loads (all in flight at once: the previous one-warp-per-output
  // loop was a chain of exposed L2 round trips, 27 us per launch) and combine in a fixed shuffle order.
  extern __shared__ float hs[];
  const int n = blockIdx.x / chunk, j = blockIdx.x % chunk;
  const bf16* h = hn + (static_cast<long>(n) * suffix_len + (suffix_len - chunk) + j) * ld;
  for (int i = threadIdx.x; i < width; i += blockDim.x) hs[i] = __bfloat162float(h[i]);
  __syncthreads();
  const int s = threadIdx.x & 7;
  for (int o0 = 0; o0 < adim; o0 += 32) {
    const int o = o0 + (threadIdx.x >> 3);
    float acc = 0.f;
    if (o < adim) {
      const float* wr = w + static_cast<long>(o) * width;
#pragma unroll 8
      for (int k = s * 4; k < width; k += 32) {
        const float4 w4 = *reinterpret_cast<const float4*>(wr + k);
        acc = fmaf(hs[k], w4.x, acc);
        acc = fmaf(hs[k + 1], w4.y, acc);
        acc = fmaf(hs[k + 2], w4.z, acc);
        acc = fmaf(hs[k + 3], w4.w, acc);
      }
    }
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    acc += __shfl_xor_sync(0xffffffffu, acc, 4);
    if (s == 0 && o < adim) {
      const float v = acc + bias[o];
      const long idx = (static_cast<long>(n) * chunk + j) * adim + o;
      if (v_out != nullptr) v_out[idx] = v;
      x_t[idx] = __fadd_rn(x_t[idx], __fmul_rn(dt, v));
    }
  }
}

int action_out_euler(cudaStream_t st, const bf16* hn, long ld, const float* w, const float* bias,
                     float* x_t, float* v_out, int n_cand, int width, int adim, int chunk,
                     int suffix_len, float dt) {
  CVB_REQUIRE(width % 4 == 0 && (reinterpret_cast<uintptr_t>(w) & 15) == 0, "action_out_proj rows must be 16-byte addressable");
  CVB_TRY(launch_pdl(action_out_euler_kernel, dim3(n_cand * chunk), dim3(256), static_cast<size_t>(width) * sizeof(float), st, 1,
                     hn, ld, w, bias, x_t, v_out, width, adim, chunk, suffix_len, dt));
  CVB_LAUNCHED();
  return 0;
}

// ------------------------------------------------------------------------------------------------
// suffix[n, 0, :] = float(bf16(state_emb))  (the state row never changes across denoise steps)
__global__ void fill_state_rows_kernel(const float* __restrict__ state_emb, float* __restrict__ suffix,
                                       int width, int suffix_len, int cpo) {
  pdl_wait();
  pdl_launch();
  float* dst = suffix + static_cast<long>(blockIdx.x) * suffix_len * width;
  if (cpo > 0) state_emb += static_cast<long>(blockIdx.x / cpo) * width;  // state of this candidate's observation
  for (int i = threadIdx.x; i < width; i += blockDim.x) dst[i] = bf16_round(state_emb[i]);
}

int fill_state_rows(cudaStream_t st, const float* state_emb, float* suffix, int n_cand, int width,
                    int suffix_len, int cands_per_obs) {
  CVB_TRY(launch_pdl(fill_state_rows_kernel, dim3(n_cand), dim3(256), 0, st, 1, state_emb, suffix, width, suffix_len,
                     cands_per_obs));
  CVB_LAUNCHED();
  return 0;
}

// plen[r] = n_img + lang_len[r]; rows_valid = sum (unused for now)
__global__ void prefix_len_kernel(const int* __restrict__ lang_len, int* __restrict__ plen, int R,
                                  int n_img, int max_lang) {
  pdl_wait();
  pdl_launch();
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < R) plen[r] = n_img + min(max(lang_len[r], 0), max_lang);
}

int prefix_lengths(cudaStream_t st, const int* lang_len, int* plen, int R, int n_img, int max_lang) {
  CVB_TRY(launch_pdl(prefix_len_kernel, dim3((R + 63) / 64), dim3(64), 0, st, 1, lang_len, plen, R, n_img, max_lang));
  CVB_LAUNCHED();
  return 0;
}

__global__ void bf16_table_to_f32_kernel(const bf16* __restrict__ src, float* __restrict__ dst, long n) {
  pdl_wait();
  pdl_launch();
  const long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = __bfloat162float(src[i]);
}
int bf16_to_f32(cudaStream_t st, const bf16* src, float* dst, long n) {
  CVB_TRY(launch_pdl(bf16_table_to_f32_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, st, 1, src, dst, n));
  CVB_LAUNCHED();
  return 0;
}

}  // namespace cvb
