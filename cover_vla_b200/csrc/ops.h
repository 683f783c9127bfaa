// Internal C++ launch API (host side).  Every function enqueues work on `st` and returns 0 or a
// negative error code with the message in cvb_last_error(); nothing here synchronises the device.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace cvb {

typedef __nv_bfloat16 bf16;

// ---- tensor-core GEMM (gemm_tcgen05.cuh) -------------------------------------------------------
struct GemmCall {
  const bf16* A = nullptr;  // [M, K], leading dim lda
  long lda = 0;
  const bf16* W = nullptr;  // [N, K], leading dim ldw   (nn.Linear weight layout)
  long ldw = 0;
  int M = 0, N = 0, K = 0;
  int epi = 0;  // EpiKind
  void* C = nullptr;
  long ldc = 0;
  const void* bias = nullptr;
  int bias_is_f32 = 0;
  const void* resid = nullptr;  // bf16, or fp32 when resid_is_f32
  int resid_is_f32 = 0;
  long ldr = 0;
  int n_out = 0;               // EPI_GEGLU: intermediate size
  const int* m_dev = nullptr;  // optional device-side row count
  int force_bn = 0;            // 0 = heuristic, else 64 / 128 / 256
  int no_skinny = 0;           // batch-capable handles: never the swap-AB split-K kernel (its fp32 summation order differs
                               // from the general kernel's, and the choice would depend on how many rows share the launch)
};
int gemm_bf16(cudaStream_t st, const GemmCall& c);
// Split-K GEMM (M <= 256) that leaves S fp32 partial products in C = float[S][M][ldc] (epi / bias / resid unused);
// splits <= 0 picks S so that n_tiles * S fills the SMs once.  The partials are consumed by rmsnorm_reduce() /
// layernorm_reduce(), which sum them in split order (deterministic) inside the norm that follows the linear layer anyway.
int gemm_splitk_partial(cudaStream_t st, const GemmCall& c, int splits, int* splits_out);

// ---- fp32 SIMT GEMM (sgemm.cuh): C = act(A[M,K] * W[N,K]^T + bias) (+ resid) -------------------
enum SgemmAct : int { SACT_NONE = 0, SACT_RELU = 1, SACT_GELU_ERF = 2, SACT_SILU = 3 };
struct SgemmCall {
  const float* A = nullptr;
  long lda = 0;
  const float* W = nullptr;
  long ldw = 0;
  int M = 0, N = 0, K = 0;
  float* C = nullptr;
  long ldc = 0;
  const float* bias = nullptr;
  const float* resid = nullptr;  // added after activation
  long ldr = 0;
  int act = 0;
  const float* row_bias = nullptr;  // optional second bias [N] (e.g. the per-step time vector)
  int w_kn = 0;       // 1: W is [K, N] row-major (C = A @ W) instead of [N, K]
  int out_group = 0;  // >0: output row m lands at (m / g) * (g + 1) + 1 + m % g (suffix layout)
  int w_dynamic = 0;  // 1: W is produced on the stream (not a weight): never read it ahead of the dependency wait
};
int sgemm_f32(cudaStream_t st, const SgemmCall& c);

// ---- normalisation ------------------------------------------------------------------------------
// Gemma RMSNorm: y = bf16( x * rsqrt(mean(x^2) + eps) * (1 + w) ), statistics in fp32.
int rmsnorm(cudaStream_t st, const void* x, int x_is_f32, long ldx, const void* w, int w_is_f32,
            bf16* y, long ldy, int rows, int width, float eps, const int* rows_dev);
// Linear-layer tail + Gemma RMSNorm in one pass over the split-K partials of gemm_splitk_partial():
//   h = bf16( bf16(sum_s P[s]) + resid )   (the EPI_RESID ledger; h_out may alias resid)
//   y = bf16( h * rsqrt(mean(h^2) + eps) * (1 + w) )
int rmsnorm_reduce(cudaStream_t st, const float* P, int S, long split_stride, long ldp, const void* resid,
                   int resid_is_f32, long ldr, const void* w, int w_is_f32, bf16* h_out, long ldh, bf16* y, long ldy,
                   int rows, int width, float eps);
// Same for a LayerNorm block (SigLIP tower): h = bf16(bf16(sum_s P[s] + bias) + resid), y = LayerNorm(h) (bf16 affine).
int layernorm_reduce(cudaStream_t st, const float* P, int S, long split_stride, long ldp, const bf16* bias,
                     const bf16* resid, long ldr, const bf16* w, const bf16* b, bf16* h_out, long ldh, bf16* y, long ldy,
                     int rows, int width, float eps);
// LayerNorm on bf16 rows (fp32 statistics), bf16 affine.
int layernorm_bf16(cudaStream_t st, const bf16* x, long ldx, const bf16* w, const bf16* b, bf16* y,
                   long ldy, int rows, int width, float eps);
// LayerNorm on fp32 rows, fp32 affine (verifier heads); optional residual added BEFORE the norm.
// y_split (optional): the result also as the [hi | hi | lo] bf16 A operand of the 3-term bf16 GEMM (split3_rows)
int layernorm_f32(cudaStream_t st, const float* x, const float* resid, const float* w,
                  const float* b, float* y, int rows, int width, float eps, bf16* y_split = nullptr,
                  int relu = 0);  // relu: max(LayerNorm(x), 0)
// fp32-accurate GEMM on the bf16 tensor cores: out[rows, 3K] bf16 = [hi | hi | lo] (mode 0, activations) or [hi | lo | hi]
// (mode 1, weights) of x[rows, K] fp32, hi = bf16(x), lo = bf16(x - hi); act 1 applies ReLU first.  A GEMM over the 3K
// axis then evaluates a_hi w_hi + a_hi w_lo + a_lo w_hi with fp32 accumulation (error ~2^-16 relative per product).
int split3_rows(cudaStream_t st, const float* x, long ldx, bf16* out, long rows, int K, int mode, int act = 0);
// out[map(m)][n] = sum_s P[s][m][n] + bias[n] (fp32; split order); out_group as in SgemmCall
int partial_reduce_f32(cudaStream_t st, const float* P, int S, long split_stride, long ldp, const float* bias, float* out,
                       long ldo, int rows, int N, int out_group);

// ---- observation pre-processing (ops_preprocess.cu) ----------------------------------------------------------
// cv2.resize(frame_u8_hwc, (dw, dh), INTER_LANCZOS4) bit-exact (+ optional x/255 -> (x - 0.5)/0.5 float32 CHW output)
int preprocess_policy_image(cudaStream_t st, const uint8_t* img_hwc, int H, int W, int dh, int dw, uint8_t* out_u8_hwc,
                            float* out_f32_chw);

// PIL Image.resize((dw, dh), BICUBIC) bit-exact (+ optional ToTensor / Normalize(0.5, 0.5) float32 CHW output)
int preprocess_verifier_image(cudaStream_t st, const uint8_t* img_hwc, int H, int W, int dh, int dw, uint8_t* out_u8_hwc,
                              float* out_f32_chw);

// tf.image.resize(frame_u8, (dh, dw), BILINEAR, antialias=True) -> uint8 (truncating cast): process_raw_image_to_jpg,
// eval_utils.py:228-286.  scratch_f32: dh * W * 3 floats (the row pass's intermediate).
int resize_bilinear_antialias_u8(cudaStream_t st, const uint8_t* img_hwc, int H, int W, int dh, int dw, float* scratch_f32,
                                 uint8_t* out_u8_hwc);

// ---- attention (attention.cuh) -----------------------------------------------------------------
struct AttnCall {
  // Q rows for batch b, head h, token t:  q + (b*q_batch_stride + t*q_row_stride + h*head_dim)
  const bf16* q = nullptr;
  long q_batch_stride = 0, q_row_stride = 0;
  // key/value segment 0 (shared per kv-batch): k0 + (kvb*kv0_batch_stride + j*kv0_row_stride + kvh*head_dim)
  const bf16* k0 = nullptr;
  const bf16* v0 = nullptr;
  long kv0_batch_stride = 0, kv0_row_stride = 0;
  const int* kv0_len_dev = nullptr;  // [kv batches] valid keys in segment 0 (nullptr -> kv0_len)
  int kv0_len = 0;
  int kv0_max = 0;  // upper bound of kv0_len_dev[] (sizes the logits buffer of the fast path)
  int force_two_pass = 0;
  int q_per_kv_batch = 1;  // kv batch index = b / q_per_kv_batch
  // optional segment 1 (per q-batch, e.g. the suffix tokens' own keys): tq1 keys
  const bf16* k1 = nullptr;
  const bf16* v1 = nullptr;
  long kv1_batch_stride = 0, kv1_row_stride = 0;
  int kv1_len = 0;
  int suffix_mask = 0;  // 1: query token 0 only sees segment-1 key 0 (pi0 suffix att mask)
  bf16* out = nullptr;  // out + (b*o_batch_stride + t*o_row_stride + h*head_dim)
  long o_batch_stride = 0, o_row_stride = 0;
  int batches = 0, heads = 0, kv_heads = 0, tq = 0, head_dim = 0;
  float scale = 1.f;
  // optional (cos, sin) table [kv batches][tq][head_dim/2] for positions kv0_len + t: RoPE is applied to q and to
  // the segment-1 keys while staging (cluster decode kernel only; requires k1)
  const float2* rope = nullptr;
  // segment 0 (and kv0_len_dev) is NOT written by the kernel launched just before this one: its tiles may be
  // prefetched before the programmatic-dependency wait (the prefix KV cache during the denoise loop)
  int kv0_static = 0;
  int algo = 0;  // 0 auto, 3 tcgen05 kernels only (error if the shape is not eligible)
  // optional TRANSPOSED copy of segment-0 values, vt0[(kv batch * head_dim + d) * vt0_ld + key] (finite past the valid
  // length): enables the tcgen05 decode kernel (ops_attention_umma.cu) for MQA / head_dim 256 shapes
  const bf16* vt0 = nullptr;
  long vt0_ld = 0;
  // optional (tcgen05 decode kernel only): q / k1 / v1 given as fp32 split-K partials of the fused qkv projection,
  // [part_splits][same element strides as q / k1 / v1]; summed in split order and rounded to bf16 while staging
  // optional (tcgen05 decode kernel only), SURVEY.md F7: the first kv1_cached of the kv1_len suffix keys / values come from
  // a cache ([batches][kv1_cached][head_dim] bf16, keys already rotated) - k1 / v1 then hold kv1_len - kv1_cached rows;
  // kv1_cache_out_*: the rotated key / value of suffix key 0 are written there (the step that fills the cache);
  // rope_rows / rope_off: rows of the rope table per kv batch (0 = tq) and the table row of query token 0 / new key 0
  const bf16* kv1_cached_k = nullptr;
  const bf16* kv1_cached_v = nullptr;
  int kv1_cached = 0;
  bf16* kv1_cache_out_k = nullptr;
  bf16* kv1_cache_out_v = nullptr;
  int rope_rows = 0, rope_off = 0;
  const float* q_part = nullptr;
  const float* k1_part = nullptr;
  const float* v1_part = nullptr;
  int part_splits = 0;
  long part_split_stride = 0;
};
int attention(cudaStream_t st, const AttnCall& c);

// ---- tcgen05 prefix attention (ops_attention_umma.cu): MQA, 8 query heads x head_dim 256, <= 384 keys ----------
struct UmmaAttnCall {
  // q rows: token (b * q_rows_per_batch + t) at q + row * q_ld, head h at column h * head_dim
  const bf16* q = nullptr;
  long q_ld = 0, q_total_rows = 0, q_rows_per_batch = 0;
  // keys: [k_total_rows, head_dim] row-major, batch b starts at row b * k_rows_per_batch
  const bf16* k = nullptr;
  long k_total_rows = 0, k_rows_per_batch = 0;
  // values TRANSPOSED: vt[(b * head_dim + d) * vt_ld + key]; columns >= the valid length must be finite
  const bf16* vt = nullptr;
  long vt_ld = 0;
  const int* klen_dev = nullptr;  // [batches] valid keys (nullptr -> klen)
  int klen = 0;
  int kmax = 0;  // upper bound of the valid key count (the kernel processes round_up(kmax, 16) keys)
  bf16* out = nullptr;  // out + b * o_batch_stride + t * o_row_stride + h * head_dim
  long o_batch_stride = 0, o_row_stride = 0;
  int batches = 0, tq = 0, heads = 0, head_dim = 0;
  float scale = 1.f;
};
bool attention_umma_eligible(const UmmaAttnCall& c);
int attention_umma(cudaStream_t st, const UmmaAttnCall& c);

}  // namespace cvb
