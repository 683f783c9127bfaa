// Verifier-specific kernels (all fp32, as the reference keeps everything after the trunk hooks in
// float32 - finetune_trajectory_bridge_ddp.py:329,352).  See verifier_kernels.h for reference lines.
#include "host_common.h"
#include "ptx.cuh"
#include "verifier_kernels.h"

namespace cvb {

namespace {
__device__ __forceinline__ float wsum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float wmax(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float bsum(float v, float* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = wsum(v);
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  float t = (lane < nw) ? red[lane] : 0.f;
  return wsum(t);
}
__device__ __forceinline__ float bmax(float v, float* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = wmax(v);
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  float t = (lane < nw) ? red[lane] : -INFINITY;
  return wmax(t);
}
}  // namespace

// ---------------------------------------------------------------------------------------------
__global__ void embed_tokens_pos_kernel(const bf16* __restrict__ table, const bf16* __restrict__ pos,
                                        const int64_t* __restrict__ tok, bf16* __restrict__ out, int width, int ctx) {
  pdl_wait();
  pdl_launch();
  const int t = blockIdx.x;
  const bf16* src = table + tok[t] * width;
  const bf16* pp = pos + static_cast<long>(t % ctx) * width;  // several texts of ctx tokens back to back
  for (int i = threadIdx.x; i < width; i += blockDim.x)
    out[static_cast<long>(t) * width + i] =
        __float2bfloat16_rn(__bfloat162float(src[i]) + __bfloat162float(pp[i]));
}
int embed_tokens_pos(cudaStream_t st, const bf16* table, const bf16* pos, const int64_t* tok, bf16* out,
                     int tokens, int width, int ctx) {
  CVB_TRY(launch_pdl(embed_tokens_pos_kernel, dim3(tokens), dim3(256), 0, st, 1, table, pos, tok, out, width,
                     ctx > 0 ? ctx : tokens));
  CVB_LAUNCHED();
  return 0;
}

// y[r,:] = float(x[r,:]) / ||float(x[r,:])||_2
__global__ void l2norm_bf16_kernel(const bf16* __restrict__ x, long ldx, float* __restrict__ y, int width) {
  pdl_wait();
  pdl_launch();
  __shared__ float red[32];
  const bf16* xr = x + blockIdx.x * ldx;
  float ss = 0.f;
  for (int i = threadIdx.x; i < width; i += blockDim.x) {
    const float v = __bfloat162float(xr[i]);
    ss += v * v;
  }
  const float nrm = sqrtf(bsum(ss, red));
  for (int i = threadIdx.x; i < width; i += blockDim.x)
    y[static_cast<long>(blockIdx.x) * width + i] = __bfloat162float(xr[i]) / nrm;
}
// y[r,:] = x[r,:] / ||x[r,:]||_2 on fp32 rows (MLP action encoder output, finetune_trajectory_bridge_ddp.py:413)
__global__ void l2norm_f32_kernel(const float* __restrict__ x, float* __restrict__ y, int width) {
  pdl_wait();
  pdl_launch();
  __shared__ float red[32];
  const float* xr = x + static_cast<long>(blockIdx.x) * width;
  float ss = 0.f;
  for (int i = threadIdx.x; i < width; i += blockDim.x) ss += xr[i] * xr[i];
  const float nrm = sqrtf(bsum(ss, red));
  for (int i = threadIdx.x; i < width; i += blockDim.x) y[static_cast<long>(blockIdx.x) * width + i] = xr[i] / nrm;
}
int l2norm_rows_f32(cudaStream_t st, const float* x, float* y, int rows, int width) {
  CVB_TRY(launch_pdl(l2norm_f32_kernel, dim3(rows), dim3(256), 0, st, 1, x, y, width));
  CVB_LAUNCHED();
  return 0;
}
int l2norm_rows_bf16_to_f32(cudaStream_t st, const bf16* x, long ldx, float* y, int rows, int width) {
  CVB_TRY(launch_pdl(l2norm_bf16_kernel, dim3(rows), dim3(256), 0, st, 1, x, ldx, y, width));
  CVB_LAUNCHED();
  return 0;
}

// in-place row softmax of x / clamp(*temp, 0, 100)
__global__ void softmax_temp_kernel(float* __restrict__ x, int cols, const float* __restrict__ temp) {
  pdl_wait();
  pdl_launch();
  __shared__ float red[32];
  float* xr = x + static_cast<long>(blockIdx.x) * cols;
  const float t = fminf(fmaxf(*temp, 0.f), 100.f);
  float m = -INFINITY;
  for (int i = threadIdx.x; i < cols; i += blockDim.x) m = fmaxf(m, xr[i] / t);
  m = bmax(m, red);
  float s = 0.f;
  for (int i = threadIdx.x; i < cols; i += blockDim.x) s += expf(xr[i] / t - m);
  s = bsum(s, red);
  for (int i = threadIdx.x; i < cols; i += blockDim.x) xr[i] = expf(xr[i] / t - m) / s;
}
int softmax_rows_temp(cudaStream_t st, float* x, int rows, int cols, const float* temp_dev) {
  CVB_TRY(launch_pdl(softmax_temp_kernel, dim3(rows), dim3(256), 0, st, 1, x, cols, temp_dev));
  CVB_LAUNCHED();
  return 0;
}

__global__ void add_f32_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ y, long n) {
  pdl_wait();
  pdl_launch();
  const long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) y[i] = a[i] + b[i];
}
int add_f32(cudaStream_t st, const float* a, const float* b, float* y, long n) {
  CVB_TRY(launch_pdl(add_f32_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, st, 1, a, b, y, n));
  CVB_LAUNCHED();
  return 0;
}

// ---------------------------------------------------------------------------------------------
// AttentionPooling with one learned query (model.py:76-112): the whole 4-block chain for one pooling
// in ONE CTA (K/V projections of all blocks are precomputed by one GEMM since they do not depend on q).
namespace {
// ---- thread-block cluster plumbing: a pooling chain is a SERIAL chain of 16 matrix-vector products (1 MB of fp32 weights
// each).  One CTA per chain (round 1: 6 CTAs on a 148-SM part, 1.0 ms) streams them at ~16 GB/s; a cluster of kPoolCluster
// CTAs per chain splits every product by output rows and broadcasts its slice of the result into every CTA's shared
// memory (st.shared::cluster), one cluster barrier per product.
constexpr int kPoolCluster = 8;
__device__ __forceinline__ uint32_t cl_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t cl_size() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cl_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// out[o] = v in the shared memory of EVERY CTA of the cluster (same offset in each)
__device__ __forceinline__ void cl_store_all(float* out_local, int o, float v, uint32_t ncta) {
  const uint32_t a = smem_u32(out_local + o);
  for (uint32_t p = 0; p < ncta; ++p) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(p));
    asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(r), "f"(v) : "memory");
  }
}

// out = W in + bias.  ncta == 1: the whole product in this CTA (then __syncthreads).  ncta > 1: this CTA computes rows
// [rank * n_out / ncta, (rank + 1) * n_out / ncta), stores them into every CTA of the cluster and the cluster syncs.
// The arithmetic of one output row is the same in both modes (bit-identical results).
__device__ void matvec(float* __restrict__ out, const float* __restrict__ W, const float* __restrict__ in,
                       const float* __restrict__ bias, int n_out, int n_in, uint32_t rank = 0, uint32_t ncta = 1) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const int o_begin = static_cast<int>(static_cast<long>(rank) * n_out / ncta);
  const int o_end = static_cast<int>(static_cast<long>(rank + 1) * n_out / ncta);
  auto put = [&](int o, float v) {
    if (ncta > 1)
      cl_store_all(out, o, v, ncta);
    else
      out[o] = v;
  };
  if ((n_in & 127) == 0) {
    // two output rows per warp iteration, float4 loads: 8+ independent 16-byte loads in flight per lane
    for (int o = o_begin + warp * 2; o < o_end; o += nw * 2) {
      const float4* w0 = reinterpret_cast<const float4*>(W + static_cast<long>(o) * n_in);
      const float4* w1 = reinterpret_cast<const float4*>(W + static_cast<long>(min(o + 1, o_end - 1)) * n_in);
      const float4* x4 = reinterpret_cast<const float4*>(in);
      float a0 = 0.f, a1 = 0.f;
#pragma unroll 4
      for (int i = lane; i < n_in / 4; i += 32) {
        const float4 a = w0[i], b = w1[i], x = x4[i];
        a0 = fmaf(a.x, x.x, fmaf(a.y, x.y, fmaf(a.z, x.z, fmaf(a.w, x.w, a0))));
        a1 = fmaf(b.x, x.x, fmaf(b.y, x.y, fmaf(b.z, x.z, fmaf(b.w, x.w, a1))));
      }
      a0 = wsum(a0);
      a1 = wsum(a1);
      if (lane == 0) {
        put(o, a0 + (bias != nullptr ? bias[o] : 0.f));
        if (o + 1 < o_end) put(o + 1, a1 + (bias != nullptr ? bias[o + 1] : 0.f));
      }
    }
  } else {
    for (int o = o_begin + warp; o < o_end; o += nw) {
      const float* wr = W + static_cast<long>(o) * n_in;
      float acc = 0.f;
      for (int i = lane; i < n_in; i += 32) acc = fmaf(wr[i], in[i], acc);
      acc = wsum(acc);
      if (lane == 0) put(o, acc + (bias != nullptr ? bias[o] : 0.f));
    }
  }
  if (ncta > 1)
    cl_sync();  // every CTA's slice has landed everywhere (release / acquire at cluster scope); also a CTA barrier
  else
    __syncthreads();
}
__device__ void layernorm_vec(float* __restrict__ out, const float* __restrict__ in, const float* __restrict__ w,
                              const float* __restrict__ b, int n, float* red) {
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += in[i];
  const float mean = bsum(s, red) / n;
  float vs = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) vs += (in[i] - mean) * (in[i] - mean);
  const float rstd = 1.0f / sqrtf(bsum(vs, red) / n + 1e-5f);
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x) out[i] = (in[i] - mean) * rstd * w[i] + b[i];
  __syncthreads();
}
// Single-query attention of ONE head over T tokens (model.py:22-56 with a 1-token query): 8 threads per token form the
// dot product from 32-byte slices of the K row (coalesced 256-byte rows), one warp does the softmax, then thread (d, group of
// tokens) accumulates P V with V rows read contiguously over d; the groups are combined in a fixed order.  The head's hd
// outputs go to out_b[h * hd ..] of every CTA of the cluster (ncta > 1) - the caller synchronises.
__device__ void attend_head(float* __restrict__ out_b, float* __restrict__ prob, float* __restrict__ part,
                            const float* __restrict__ qv, const float* __restrict__ Kp, const float* __restrict__ Vp,
                            long kv_ld, int h, int hd, int T, uint32_t ncta) {
  const int tid = threadIdx.x, nt = blockDim.x;
  float* ph = prob + h * T;
  const float* qh = qv + h * hd;
  for (int j0 = 0; j0 < T; j0 += nt / 8) {
    const int j = j0 + tid / 8, sl = tid & 7;
    float s = 0.f;
    if (j < T) {
      const float* kr = Kp + static_cast<long>(j) * kv_ld + h * hd;
      for (int d = sl * 4; d < hd; d += 32) {  // hd is a multiple of 4 (host check): float4 slices, 8 lanes = 128 bytes
        const float4 k4 = *reinterpret_cast<const float4*>(kr + d);
        s = fmaf(qh[d], k4.x, fmaf(qh[d + 1], k4.y, fmaf(qh[d + 2], k4.z, fmaf(qh[d + 3], k4.w, s))));
      }
    }
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    s += __shfl_xor_sync(0xffffffffu, s, 4);
    if (j < T && sl == 0) ph[j] = s;
  }
  __syncthreads();
  if (tid < 32) {
    float mx = -INFINITY;
    for (int j = tid; j < T; j += 32) mx = fmaxf(mx, ph[j]);
    mx = wmax(mx);
    float sum = 0.f;
    for (int j = tid; j < T; j += 32) {
      const float e = expf(ph[j] - mx);
      ph[j] = e;
      sum += e;
    }
    sum = wsum(sum);
    __syncwarp();
    for (int j = tid; j < T; j += 32) ph[j] = ph[j] / sum;
  }
  __syncthreads();
  const int groups = nt / hd;  // token groups (host check: blockDim is a multiple of hd)
  const int d = tid % hd, g = tid / hd;
  if (g < groups) {
    float o = 0.f;
    for (int j = g; j < T; j += groups) o = fmaf(ph[j], Vp[static_cast<long>(j) * kv_ld + h * hd + d], o);
    part[g * hd + d] = o;
  }
  __syncthreads();
  if (tid < hd) {
    float o = 0.f;
    for (int g2 = 0; g2 < groups; ++g2) o += part[g2 * hd + tid];
    if (ncta > 1)
      cl_store_all(out_b, h * hd + tid, o, ncta);
    else
      out_b[h * hd + tid] = o;
  }
  __syncthreads();
}
}  // namespace

__global__ void __launch_bounds__(512) pool_chain_kernel(const PoolChain* __restrict__ chains) {
  pdl_wait();
  pdl_launch();
  extern __shared__ float sm[];
  __shared__ float red[32];
  const uint32_t ncta = cl_size(), rank = cl_rank();  // cluster of kPoolCluster CTAs per chain (or 1)
  const PoolChain& c = chains[blockIdx.x / ncta];
  const int E = c.embed, H = c.heads, hd = E / H, T = c.tokens;
  float* q = sm;             // [E]
  float* a = q + E;          // [E]
  float* b = a + E;          // [E]
  float* t = b + E;          // [E] attention output projection (its own buffer: see the hazard note below)
  float* prob = t + E;       // [H * T]
  float* part = prob + H * T;  // [blockDim / hd][hd] partial P V sums of the head this CTA attends
  for (int i = threadIdx.x; i < E; i += blockDim.x) q[i] = c.query[i];
  __syncthreads();
  // Cluster mode hazard rule: a product's destination is written REMOTELY by fast peers, so it must not be a buffer a
  // slow peer may still be reading.  a <- Wq q, a <- fc1 q, b <- fc2 a are safe (their destination was last read before
  // an earlier cluster barrier); the attention reads `a` right before the o-projection, hence t <- Wo b.
  if (ncta > 1) cl_sync();
  const float qscale = sqrtf(1.0f / static_cast<float>(hd));
  for (int l = 0; l < c.layers; ++l) {
    const PoolBlockW& w = c.blk[l];
    layernorm_vec(q, q, w.qln_w, w.qln_b, E, red);         // q = q_layer_norm(q)
    matvec(a, w.wq, q, w.b_in, E, E, rank, ncta);           // a = Wq q + bq
    for (int i = threadIdx.x; i < E; i += blockDim.x) a[i] *= qscale;
    __syncthreads();
    const float* Kp = c.kv + static_cast<long>(l) * 2 * E;  // K_l at cols [l*2E, l*2E+E), V_l after it
    const float* Vp = Kp + E;
    // one head per CTA of the cluster (heads rank, rank + ncta, ...), the head's slice of b broadcast to every CTA
    for (int h = static_cast<int>(rank); h < H; h += static_cast<int>(ncta))
      attend_head(b, prob, part, a, Kp, Vp, c.kv_ld, h, hd, T, ncta);
    if (ncta > 1)
      cl_sync();
    else
      __syncthreads();
    matvec(t, w.wo, b, w.bo, E, E, rank, ncta);             // attn_out
    for (int i = threadIdx.x; i < E; i += blockDim.x) q[i] = q[i] + t[i];
    __syncthreads();
    layernorm_vec(q, q, w.ln_w, w.ln_b, E, red);            // q = layer_norm(q + attn)
    matvec(a, w.fc1_w, q, w.fc1_b, E, E, rank, ncta);
    for (int i = threadIdx.x; i < E; i += blockDim.x)
      a[i] = 0.5f * a[i] * (1.0f + erff(a[i] * 0.70710678118654752440f));
    __syncthreads();
    matvec(b, w.fc2_w, a, w.fc2_b, E, E, rank, ncta);
    for (int i = threadIdx.x; i < E; i += blockDim.x) q[i] = q[i] + b[i];
    __syncthreads();
  }
  layernorm_vec(a, q, c.fin_w, c.fin_b, E, red);
  if (rank == 0)
    for (int i = threadIdx.x; i < E; i += blockDim.x) c.out[i] = a[i];
  if (ncta > 1) cl_sync();  // no CTA may exit while a peer could still address its shared memory
}

int pool_chains(cudaStream_t st, const PoolChain* chains_dev, int n_chains, int embed, int heads, int tokens) {
  const int hd = embed / heads;
  CVB_REQUIRE(hd % 4 == 0 && 512 % hd == 0 && hd <= 512, "pooling head_dim must divide 512 and be a multiple of 4");
  const size_t smem = (4 * embed + heads * tokens + 512) * sizeof(float);
  const int cl = embed % (2 * kPoolCluster) == 0 ? kPoolCluster : 1;  // every CTA owns an even number of output rows
  CVB_TRY(launch_pdl(pool_chain_kernel, dim3(n_chains * cl), dim3(512), smem, st, cl, chains_dev));
  CVB_LAUNCHED();
  return 0;
}

// it[m] = normalize(W_ip . cat[text_tok[m], vision_tok[m]] + b_ip)      (efficient_ensemble_merged.py:220-223)
__global__ void __launch_bounds__(256) it_finalize_kernel(const ItFinal* __restrict__ items, int E) {
  pdl_wait();
  pdl_launch();
  extern __shared__ float sm[];
  __shared__ float red[32];
  const uint32_t ncta = cl_size(), rank = cl_rank();
  const ItFinal& f = items[blockIdx.x / ncta];
  float* in = sm;       // [2E]
  float* out = sm + 2 * E;
  for (int i = threadIdx.x; i < E; i += blockDim.x) {
    in[i] = f.text_tok[i];
    in[E + i] = f.vision_tok[i];
  }
  __syncthreads();
  if (ncta > 1) cl_sync();  // every CTA of the cluster is running before any remote store
  matvec(out, f.w, in, f.b, E, 2 * E, rank, ncta);
  float ss = 0.f;
  for (int i = threadIdx.x; i < E; i += blockDim.x) ss += out[i] * out[i];
  const float nrm = sqrtf(bsum(ss, red));
  if (rank == 0)
    for (int i = threadIdx.x; i < E; i += blockDim.x) f.out[i] = out[i] / nrm;
  if (ncta > 1) cl_sync();
}
int it_finalize(cudaStream_t st, const ItFinal* items_dev, int members, int embed) {
  const int cl = embed % (2 * kPoolCluster) == 0 ? kPoolCluster : 1;
  CVB_TRY(launch_pdl(it_finalize_kernel, dim3(members * cl), dim3(256), 3 * embed * sizeof(float), st, cl, items_dev, embed));
  CVB_LAUNCHED();
  return 0;
}

// ---------------------------------------------------------------------------------------------
// nn.TransformerEncoderLayer self-attention over a <=32-step action history with key-padding mask
// (efficient_ensemble_merged.py:229-235): one CTA per candidate, one warp per head, lane = query step.
__global__ void __launch_bounds__(256) traj_attention_kernel(const float* __restrict__ qkv, const float* __restrict__ traj,
                                                             float* __restrict__ out, int S, int E, int H, int adim,
                                                             float pad_value) {
  pdl_wait();
  pdl_launch();
  // one CTA per candidate: stage its [S, 3E] q/k/v rows in shared memory (coalesced), then one warp per
  // head; lane = (query step, 1/2..1/8 slice of head_dim) so all 32 lanes work and reads stay conflict-light.
  extern __shared__ float sm[];
  float* sq = sm;                       // [S][3E]
  float* sp = sm + S * 3 * E;           // [H][S][S] probabilities
  __shared__ int spad[32];
  const int n = blockIdx.x, hd = E / H;
  const float* src = qkv + static_cast<long>(n) * S * 3 * E;
  for (int i = threadIdx.x * 4; i < S * 3 * E; i += blockDim.x * 4)
    *reinterpret_cast<float4*>(sq + i) = *reinterpret_cast<const float4*>(src + i);
  if (threadIdx.x < S) spad[threadIdx.x] = traj[(static_cast<long>(n) * S + threadIdx.x) * adim] == pad_value;
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const float scale = 1.0f / sqrtf(static_cast<float>(hd));
  for (int h = warp; h < H; h += nw) {
    float* ph = sp + h * S * S;
    // scores: S*S entries spread over the 32 lanes
    for (int e = lane; e < S * S; e += 32) {
      const int i = e / S, j = e % S;
      const float* qr = sq + i * 3 * E + h * hd;
      const float* kr = sq + j * 3 * E + E + h * hd;
      float s = 0.f;
#pragma unroll 8
      for (int d = 0; d < hd; ++d) s = fmaf(qr[d], kr[d], s);
      ph[e] = spad[j] ? -INFINITY : s * scale;
    }
    __syncwarp();
    if (lane < S) {
      float mx = -INFINITY;
      for (int j = 0; j < S; ++j) mx = fmaxf(mx, ph[lane * S + j]);
      float sum = 0.f;
      for (int j = 0; j < S; ++j) {
        const float v = ph[lane * S + j];
        const float e = v == -INFINITY ? 0.f : expf(v - mx);
        ph[lane * S + j] = e;
        sum += e;
      }
      for (int j = 0; j < S; ++j) ph[lane * S + j] = ph[lane * S + j] / sum;
    }
    __syncwarp();
    // output: S*hd entries over the lanes, d fastest (coalesced stores, conflict-free V reads)
    for (int e = lane; e < S * hd; e += 32) {
      const int i = e / hd, d = e % hd;
      float o = 0.f;
      for (int j = 0; j < S; ++j) o = fmaf(ph[i * S + j], sq[j * 3 * E + 2 * E + h * hd + d], o);
      out[(static_cast<long>(n) * S + i) * E + h * hd + d] = o;
    }
  }
}
int traj_attention(cudaStream_t st, const float* qkv, const float* traj, float* out, int n_cand, int S, int E, int H,
                   int adim, float pad_value) {
  CVB_REQUIRE(S <= 32, "history length must be <= 32");
  CVB_REQUIRE((3 * E) % 4 == 0, "embed must be a multiple of 4");
  const size_t smem = (static_cast<size_t>(S) * 3 * E + static_cast<size_t>(H) * S * S) * sizeof(float);
  if (smem > 48 * 1024) CVB_TRY(ensure_dyn_smem(traj_attention_kernel, (int)smem));
  CVB_TRY(launch_pdl(traj_attention_kernel, dim3(n_cand), dim3(256), smem, st, 1, qkv, traj, out, S, E, H, adim, pad_value));
  CVB_LAUNCHED();
  return 0;
}

// masked mean over non-padded steps, then L2 normalise        (efficient_ensemble_merged.py:236-245)
__global__ void __launch_bounds__(256) masked_mean_l2_kernel(const float* __restrict__ x, const float* __restrict__ traj,
                                                             float* __restrict__ out, int S, int E, int adim,
                                                             float pad_value) {
  pdl_wait();
  pdl_launch();
  __shared__ float red[32];
  const int n = blockIdx.x;
  float cnt = 0.f;
  for (int j = 0; j < S; ++j) cnt += (traj[(static_cast<long>(n) * S + j) * adim] == pad_value) ? 0.f : 1.f;
  cnt = fmaxf(cnt, 1e-9f);
  float ss = 0.f;
  for (int e = threadIdx.x; e < E; e += blockDim.x) {
    float s = 0.f;
    for (int j = 0; j < S; ++j) {
      const float keep = (traj[(static_cast<long>(n) * S + j) * adim] == pad_value) ? 0.f : 1.f;
      s += x[(static_cast<long>(n) * S + j) * E + e] * keep;
    }
    s = s / cnt;
    out[static_cast<long>(n) * E + e] = s;
    ss += s * s;
  }
  const float nrm = sqrtf(bsum(ss, red));
  for (int e = threadIdx.x; e < E; e += blockDim.x) out[static_cast<long>(n) * E + e] /= nrm;
}
int masked_mean_l2norm(cudaStream_t st, const float* x, const float* traj, float* out, int n_cand, int S, int E,
                       int adim, float pad_value) {
  CVB_TRY(launch_pdl(masked_mean_l2_kernel, dim3(n_cand), dim3(256), 0, st, 1, x, traj, out, S, E, adim, pad_value));
  CVB_LAUNCHED();
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Fused ensemble fusion + scores (efficient_ensemble_merged.py:404-414): one warp per candidate,
// warp-shuffle reductions; then (optionally) group-mean / argmax selection in the same launch.
__device__ void select_block(const float* __restrict__ scores, int R, int K, float* __restrict__ group_mean,
                             int* __restrict__ best_idx, float* __restrict__ best_score, float* sh_mean) {
  // group means (efficient_ensemble_merged.py:427-431)
  for (int g = threadIdx.x; g < R; g += blockDim.x) {
    float s = 0.f;
    for (int k = 0; k < K; ++k) s += scores[g * K + k];
    const float m = s / static_cast<float>(K);
    sh_mean[g] = m;
    if (group_mean != nullptr) group_mean[g] = m;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int bg = 0;
    for (int g = 1; g < R; ++g)
      if (sh_mean[g] > sh_mean[bg]) bg = g;  // first maximum wins ties, like torch.max
    int bk = 0;
    for (int k = 1; k < K; ++k)
      if (scores[bg * K + k] > scores[bg * K + bk]) bk = k;
    *best_idx = bg * K + bk;
    *best_score = scores[bg * K + bk];
  }
}

// grid = (candidate blocks, observations): 8 warps per CTA, one candidate per warp, every load of a candidate in
// flight at once (the round-1 kernel ran ONE CTA whose warps walked the members in dependent L2 round trips: 52 us).
__global__ void __launch_bounds__(256) fuse_score_kernel(const float* __restrict__ it, long it_obs_stride,
                                                         const float* __restrict__ act, long act_member_stride, int M, int N,
                                                         int E, float* __restrict__ scores) {
  pdl_wait();
  pdl_launch();
  extern __shared__ float sm[];
  __shared__ float red[32];
  const int obs = blockIdx.y;
  it += obs * it_obs_stride;
  act += static_cast<long>(obs) * N * E;
  scores += static_cast<long>(obs) * N;
  float* fit = sm;  // [E]
  float ss = 0.f;
  for (int e = threadIdx.x; e < E; e += blockDim.x) {
    float s = 0.f;
    for (int m = 0; m < M; ++m) s += it[static_cast<long>(m) * E + e];
    s = s / static_cast<float>(M);
    fit[e] = s;
    ss += s * s;
  }
  const float nrm = sqrtf(bsum(ss, red));
  __syncthreads();
  for (int e = threadIdx.x; e < E; e += blockDim.x) fit[e] = fit[e] / nrm;
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = blockIdx.x * 8 + warp;
  if (n >= N) return;
  // fused action embedding of candidate n: mean over the members, then normalise (efficient_ensemble_merged.py:404-411)
  constexpr int kMaxPer = 32;  // E <= 1024
  float v[kMaxPer];
  float an = 0.f;
#pragma unroll
  for (int j = 0; j < kMaxPer; ++j) {
    const int e = lane + 32 * j;
    v[j] = 0.f;
    if (e < E) {
      float s = 0.f;
      for (int m = 0; m < M; ++m) s += act[m * act_member_stride + static_cast<long>(n) * E + e];
      v[j] = s / static_cast<float>(M);
      an += v[j] * v[j];
    }
  }
  an = sqrtf(wsum(an));
  float dot = 0.f;
#pragma unroll
  for (int j = 0; j < kMaxPer; ++j) {
    const int e = lane + 32 * j;
    if (e < E) dot = fmaf(fit[e], v[j] / an, dot);
  }
  dot = wsum(dot);
  if (lane == 0) scores[n] = dot;
}

int select_best(cudaStream_t st, const float* scores, int R, int K, float* group_mean, int* best_idx,
                float* best_score, int n_obs);
int fuse_score_select(cudaStream_t st, const float* it, const float* act, int M, int N, int E, float* scores, int R,
                      int K, float* group_mean, int* best_idx, float* best_score, int do_select, int n_obs,
                      long act_member_stride) {
  CVB_REQUIRE(!do_select || R * K == N, "R*K must equal the number of candidates");
  CVB_REQUIRE(E <= 1024, "embedding width above 1024");
  if (act_member_stride == 0) act_member_stride = static_cast<long>(n_obs) * N * E;
  CVB_TRY(launch_pdl(fuse_score_kernel, dim3((N + 7) / 8, n_obs), dim3(256), E * sizeof(float), st, 1, it,
                     static_cast<long>(M) * E, act, act_member_stride, M, N, E, scores));
  CVB_LAUNCHED();
  if (do_select) CVB_TRY(select_best(st, scores, R, K, group_mean, best_idx, best_score, n_obs));
  return 0;
}

// one CTA per observation
__global__ void __launch_bounds__(256) select_kernel(const float* __restrict__ scores, int R, int K,
                                                     float* __restrict__ group_mean, int* __restrict__ best_idx,
                                                     float* __restrict__ best_score) {
  pdl_wait();
  pdl_launch();
  extern __shared__ float sm[];
  const int obs = blockIdx.x;
  select_block(scores + static_cast<long>(obs) * R * K, R, K, group_mean != nullptr ? group_mean + obs * R : nullptr,
               best_idx + obs, best_score + obs, sm);
}
int select_best(cudaStream_t st, const float* scores, int R, int K, float* group_mean, int* best_idx,
                float* best_score, int n_obs) {
  CVB_TRY(launch_pdl(select_kernel, dim3(n_obs), dim3(256), (R + 1) * sizeof(float), st, 1, scores, R, K, group_mean, best_idx, best_score));
  CVB_LAUNCHED();
  return 0;
}


// ---------------------------------------------------------------------------------------------
// Sampler -> verifier action formatting on the device: replaces, per decision,
// process_inputs(verifier_action=True) (eval_utils.py:172-221) -> BridgeSimplerAdapter.postprocess_verifier
// (INT-ACT/src/experiments/env_adapters/simpler.py:96-121, denormalize_bound base.py:20-31, gripper
// :221-225) and the left-padding with -5 of efficient_ensemble_merged.py:379-390.  Arithmetic mirrors
// numpy's promotion: (x + 1) / 2 in float32, then * (p99 - p01) + p01 in float64, rounded to float32.
__global__ void format_traj_kernel(const float* __restrict__ actions, int chunk, int adim_stride,
                                   FormatStats st, const float* __restrict__ past, int num_past, int history,
                                   int n_future, float* __restrict__ traj, int cands_per_obs) {
  pdl_wait();
  pdl_launch();
  const int n = blockIdx.x;
  if (cands_per_obs > 0) past += static_cast<long>(n / cands_per_obs) * num_past * 7;  // this observation's history
  const int A = 7;
  const int used = num_past + n_future;
  const int lead = history - used;  // rows of -5 padding (>= 0, checked on the host)
  for (int idx = threadIdx.x; idx < history * A; idx += blockDim.x) {
    const int row = idx / A, col = idx % A;
    float v;
    if (row < lead) {
      v = -5.0f;
    } else if (row < lead + num_past) {
      v = past[(row - lead) * A + col];
    } else {
      const int j = row - lead - num_past;
      const float x = actions[(static_cast<long>(n) * chunk + j) * adim_stride + col];
      if (col < 6) {
        const float t = (x - (-1.0f)) / 2.0f;
        v = static_cast<float>(static_cast<double>(t) * (st.p99[col] - st.p01[col]) + st.p01[col]);
      } else {
        v = x < 0.5f ? 0.0f : 1.0f;
      }
    }
    traj[(static_cast<long>(n) * history + row) * A + col] = v;
  }
}
int format_trajectories(cudaStream_t stream, const float* actions, int n_cand, int chunk, int adim_stride,
                        const FormatStats& st, const float* past, int num_past, int history, int n_future,
                        float* traj, int cands_per_obs) {
  CVB_REQUIRE(n_future >= 1 && n_future <= chunk, "n_future must be in [1, chunk]");
  CVB_REQUIRE(num_past >= 0 && num_past + n_future <= history, "history too short for past + future actions");
  CVB_REQUIRE(adim_stride >= 7, "action stride must be >= 7");
  CVB_TRY(launch_pdl(format_traj_kernel, dim3(n_cand), dim3(96), 0, stream, 1, actions, chunk, adim_stride, st, past, num_past, history, n_future,
                                                traj, cands_per_obs));
  CVB_LAUNCHED();
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Execution-format action of the winner + gripper vote on the device.  Replaces, per
// decision, process_inputs(verifier_action=False) for the winning group (eval_utils.py:172-221 ->
// BridgeSimplerAdapter.postprocess, INT-ACT/src/experiments/env_adapters/simpler.py:123-166: denormalize_bound
// base.py:20-31, euler2axangle src/utils/geometry.py:261-363 + :365-436 with axes 'sxyz', postprocess_gripper
// :211-220) and the majority vote of run_simpler_eval_with_openpi.py:368-391.  All arithmetic in float64 with numpy's /
// CPython's operation order (no fused multiply-add); (x + 1) / 2 in float32 as numpy's promotion does.
__global__ void execution_action_kernel(const float* __restrict__ actions, int chunk, int adim_stride, FormatStats st,
                                        const int* __restrict__ best_idx, int K, int step, double* __restrict__ out,
                                        int* __restrict__ votes) {
  pdl_wait();
  pdl_launch();
  if (threadIdx.x != 0) return;
  const int idx = *best_idx;
  const float* a = actions + (static_cast<long>(idx) * chunk + step) * adim_stride;
  double raw[6];
  for (int c = 0; c < 6; ++c) {
    const float t = (a[c] - (-1.0f)) / 2.0f;
    raw[c] = __dadd_rn(__dmul_rn(static_cast<double>(t), __dsub_rn(st.p99[c], st.p01[c])), st.p01[c]);
  }
  // euler2quat, static xyz: q = (w, x, y, z)
  const double ai = raw[3] / 2.0, aj = raw[4] / 2.0, ak = raw[5] / 2.0;
  const double ci = cos(ai), si = sin(ai), cj = cos(aj), sj = sin(aj), ck = cos(ak), sk = sin(ak);
  const double cc = __dmul_rn(ci, ck), cs = __dmul_rn(ci, sk), sc = __dmul_rn(si, ck), ss = __dmul_rn(si, sk);
  double q[4];
  q[0] = __dadd_rn(__dmul_rn(cj, cc), __dmul_rn(sj, ss));
  q[1] = __dsub_rn(__dmul_rn(cj, sc), __dmul_rn(sj, cs));
  q[2] = __dadd_rn(__dmul_rn(cj, ss), __dmul_rn(sj, cc));
  q[3] = __dsub_rn(__dmul_rn(cj, cs), __dmul_rn(sj, sc));
  // quat2axangle
  const double eps = 2.220446049250313e-16;
  double ax[3] = {1.0, 0.0, 0.0}, theta = 0.0;
  const double Nq = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(q[0], q[0]), __dmul_rn(q[1], q[1])), __dmul_rn(q[2], q[2])),
                              __dmul_rn(q[3], q[3]));
  if (!isfinite(Nq)) {
    theta = nan("");
  } else if (!(Nq < eps * eps)) {
    if (Nq != 1.0) {
      const double sn = sqrt(Nq);
      for (int i = 0; i < 4; ++i) q[i] = q[i] / sn;
    }
    const double len2 = __dadd_rn(__dadd_rn(__dmul_rn(q[1], q[1]), __dmul_rn(q[2], q[2])), __dmul_rn(q[3], q[3]));
    const double thr = eps * 3.0;
    if (!(len2 < thr * thr)) {
      theta = 2.0 * acos(fmax(fmin(q[0], 1.0), -1.0));
      const double l = sqrt(len2);
      for (int i = 0; i < 3; ++i) ax[i] = q[1 + i] / l;
    }
  }
  // gripper of the winner and the vote of its K-sample group (>= 0 closed, < 0 open in the reference's wording)
  const int g0 = (idx / K) * K;
  int close_votes = 0, open_votes = 0;
  for (int m = g0; m < g0 + K; ++m) {
    const float gm = actions[(static_cast<long>(m) * chunk + step) * adim_stride + 6];
    if (2.0 * (gm > 0.5f ? 1.0 : 0.0) - 1.0 >= 0.0)
      ++close_votes;
    else
      ++open_votes;
  }
  double grip = 2.0 * (a[6] > 0.5f ? 1.0 : 0.0) - 1.0;
  if (close_votes > open_votes)
    grip = 1.0;
  else if (open_votes > close_votes)
    grip = -1.0;
  else
    grip = grip >= 0.0 ? 1.0 : -1.0;
  out[0] = raw[0], out[1] = raw[1], out[2] = raw[2];
  for (int i = 0; i < 3; ++i) out[3 + i] = __dmul_rn(ax[i], theta);
  out[6] = grip;
  if (votes != nullptr) votes[0] = close_votes, votes[1] = open_votes;
}

int execution_action(cudaStream_t stream, const float* actions, int n_cand, int chunk, int adim_stride, const FormatStats& st,
                     const int* best_idx, int K, int step, double* out, int* votes) {
  CVB_REQUIRE(step >= 0 && step < chunk, "step must be in [0, chunk)");
  CVB_REQUIRE(K >= 1 && n_cand % K == 0, "the candidate count must be a multiple of K");
  CVB_REQUIRE(adim_stride >= 7, "action stride must be >= 7");
  CVB_TRY(launch_pdl(execution_action_kernel, dim3(1), dim3(32), 0, stream, 1, actions, chunk, adim_stride, st, best_idx, K,
                     step, out, votes));
  CVB_LAUNCHED();
  return 0;
}

}  // namespace cvb
