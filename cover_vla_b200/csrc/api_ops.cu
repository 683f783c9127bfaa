// C-ABI entry points for the individual operators (used by the parity tests and by hosts that want
// to compose their own graph).  Signatures are declared in include/coverb200.h.
#include "../../include/coverb200.h"
#include "gemm_tcgen05.cuh"
#include "host_common.h"
#include "ops.h"

namespace cvb {
const char* get_last_error();
extern unsigned long long* g_skinny_ts;
}

extern "C" {

const char* cvb_last_error(void) { return cvb::get_last_error(); }

int cvb_abi_version(void) { return CVB_ABI_VERSION; }

int64_t cvb_launch_count(void) { return cvb::launch_count(); }

void cvb_debug_set_timestamps(void* dev_u64) { cvb::g_skinny_ts = (unsigned long long*)dev_u64; }

int cvb_op_gemm_bf16(const void* A, int64_t lda, const void* W, int64_t ldw, int M, int N, int K,
                     int epilogue, void* C, int64_t ldc, const void* bias, int bias_is_f32,
                     const void* resid, int resid_is_f32, int64_t ldr, int n_out,
                     const int32_t* m_dev, int force_bn, void* stream) {
  cvb::GemmCall c;
  c.A = (const cvb::bf16*)A;
  c.lda = lda;
  c.W = (const cvb::bf16*)W;
  c.ldw = ldw;
  c.M = M;
  c.N = N;
  c.K = K;
  c.epi = epilogue;
  c.C = C;
  c.ldc = ldc;
  c.bias = bias;
  c.bias_is_f32 = bias_is_f32;
  c.resid = resid;
  c.resid_is_f32 = resid_is_f32;
  c.ldr = ldr;
  c.n_out = n_out;
  c.m_dev = m_dev;
  c.force_bn = force_bn;
  return cvb::gemm_bf16((cudaStream_t)stream, c);
}

int cvb_op_rmsnorm_reduce(const float* P, int S, int64_t split_stride, int64_t ldp, const void* resid, int resid_is_f32,
                          int64_t ldr, const void* w, int w_is_f32, void* h_out, int64_t ldh, void* y, int64_t ldy,
                          int rows, int width, float eps, void* stream) {
  return cvb::rmsnorm_reduce((cudaStream_t)stream, P, S, split_stride, ldp, resid, resid_is_f32, ldr, w, w_is_f32,
                             (cvb::bf16*)h_out, ldh, (cvb::bf16*)y, ldy, rows, width, eps);
}

int cvb_op_rmsnorm(const void* x, int x_is_f32, int64_t ldx, const void* w, int w_is_f32, void* y, int64_t ldy, int rows,
                   int width, float eps, void* stream) {
  return cvb::rmsnorm((cudaStream_t)stream, x, x_is_f32, ldx, w, w_is_f32, (cvb::bf16*)y, ldy, rows, width, eps, nullptr);
}

int cvb_op_layernorm_reduce(const float* P, int S, int64_t split_stride, int64_t ldp, const void* bias, const void* resid,
                            int64_t ldr, const void* w, const void* b, void* h_out, int64_t ldh, void* y, int64_t ldy,
                            int rows, int width, float eps, void* stream) {
  return cvb::layernorm_reduce((cudaStream_t)stream, P, S, split_stride, ldp, (const cvb::bf16*)bias,
                               (const cvb::bf16*)resid, ldr, (const cvb::bf16*)w, (const cvb::bf16*)b, (cvb::bf16*)h_out,
                               ldh, (cvb::bf16*)y, ldy, rows, width, eps);
}

int cvb_op_sgemm_f32(const float* A, int64_t lda, const float* W, int64_t ldw, int M, int N, int K, float* C,
                     int64_t ldc, const float* bias, const float* row_bias, const float* resid, int64_t ldr, int act,
                     void* stream) {
  cvb::SgemmCall c;
  c.A = A, c.lda = lda, c.W = W, c.ldw = ldw, c.M = M, c.N = N, c.K = K, c.C = C, c.ldc = ldc;
  c.bias = bias, c.row_bias = row_bias, c.resid = resid, c.ldr = ldr, c.act = act;
  c.w_dynamic = 1;  // operator level: W may have been produced by the caller's previous launch
  return cvb::sgemm_f32((cudaStream_t)stream, c);
}

int cvb_op_split3_f32(const float* x, int64_t ldx, void* out_bf16, int64_t rows, int K, int weight_layout, int relu,
                      void* stream) {
  return cvb::split3_rows((cudaStream_t)stream, x, ldx, reinterpret_cast<cvb::bf16*>(out_bf16), rows, K, weight_layout ? 1 : 0,
                          relu ? 1 : 0);
}

int cvb_op_attention_tc(const void* q, int64_t q_bs, int64_t q_rs, const void* k0, const void* v0, int64_t kv0_bs,
                        int64_t kv0_rs, const int32_t* kv0_len_dev, int kv0_len, int kv0_max, int q_per_kv_batch,
                        const void* k1, const void* v1, int64_t kv1_bs, int64_t kv1_rs, int kv1_len, int suffix_mask,
                        void* out, int64_t o_bs, int64_t o_rs, int batches, int heads, int kv_heads, int tq,
                        int head_dim, float scale, int force_two_pass, const float* rope_cos_sin, const void* vt0,
                        int64_t vt0_ld, void* stream) {
  cvb::AttnCall c;
  c.q = (const cvb::bf16*)q, c.q_batch_stride = q_bs, c.q_row_stride = q_rs;
  c.k0 = (const cvb::bf16*)k0, c.v0 = (const cvb::bf16*)v0, c.kv0_batch_stride = kv0_bs, c.kv0_row_stride = kv0_rs;
  c.kv0_len_dev = kv0_len_dev, c.kv0_len = kv0_len, c.kv0_max = kv0_max, c.q_per_kv_batch = q_per_kv_batch;
  c.k1 = (const cvb::bf16*)k1, c.v1 = (const cvb::bf16*)v1, c.kv1_batch_stride = kv1_bs, c.kv1_row_stride = kv1_rs;
  c.kv1_len = kv1_len, c.suffix_mask = suffix_mask;
  c.out = (cvb::bf16*)out, c.o_batch_stride = o_bs, c.o_row_stride = o_rs;
  c.batches = batches, c.heads = heads, c.kv_heads = kv_heads, c.tq = tq, c.head_dim = head_dim, c.scale = scale;
  c.force_two_pass = force_two_pass == 1;
  c.algo = force_two_pass == 4 ? 3 : 0;
  c.rope = reinterpret_cast<const float2*>(rope_cos_sin);
  c.vt0 = (const cvb::bf16*)vt0, c.vt0_ld = vt0_ld;
  return cvb::attention((cudaStream_t)stream, c);
}

int cvb_op_attention(const void* q, int64_t q_bs, int64_t q_rs, const void* k0, const void* v0, int64_t kv0_bs,
                     int64_t kv0_rs, const int32_t* kv0_len_dev, int kv0_len, int kv0_max, int q_per_kv_batch,
                     const void* k1, const void* v1, int64_t kv1_bs, int64_t kv1_rs, int kv1_len, int suffix_mask,
                     void* out, int64_t o_bs, int64_t o_rs, int batches, int heads, int kv_heads, int tq,
                     int head_dim, float scale, int force_two_pass, const float* rope_cos_sin, void* stream) {
  return cvb_op_attention_tc(q, q_bs, q_rs, k0, v0, kv0_bs, kv0_rs, kv0_len_dev, kv0_len, kv0_max, q_per_kv_batch, k1,
                             v1, kv1_bs, kv1_rs, kv1_len, suffix_mask, out, o_bs, o_rs, batches, heads, kv_heads, tq,
                             head_dim, scale, force_two_pass, rope_cos_sin, nullptr, 0, stream);
}

int cvb_op_attention_umma(const void* q, int64_t q_ld, int64_t q_total_rows, int64_t q_rows_per_batch, const void* k,
                          int64_t k_total_rows, int64_t k_rows_per_batch, const void* vt, int64_t vt_ld,
                          const int32_t* klen_dev, int klen, int kmax, void* out, int64_t o_bs, int64_t o_rs, int batches,
                          int tq, int heads, int head_dim, float scale, void* stream) {
  cvb::UmmaAttnCall c;
  c.q = (const cvb::bf16*)q, c.q_ld = q_ld, c.q_total_rows = q_total_rows, c.q_rows_per_batch = q_rows_per_batch;
  c.k = (const cvb::bf16*)k, c.k_total_rows = k_total_rows, c.k_rows_per_batch = k_rows_per_batch;
  c.vt = (const cvb::bf16*)vt, c.vt_ld = vt_ld, c.klen_dev = klen_dev, c.klen = klen, c.kmax = kmax;
  c.out = (cvb::bf16*)out, c.o_batch_stride = o_bs, c.o_row_stride = o_rs;
  c.batches = batches, c.tq = tq, c.heads = heads, c.head_dim = head_dim, c.scale = scale;
  return cvb::attention_umma((cudaStream_t)stream, c);
}

int cvb_preprocess_policy_image(const uint8_t* img_u8_hwc, int H, int W, int out_h, int out_w, uint8_t* out_u8_hwc,
                                float* out_f32_chw, void* stream) {
  return cvb::preprocess_policy_image((cudaStream_t)stream, img_u8_hwc, H, W, out_h, out_w, out_u8_hwc, out_f32_chw);
}

int cvb_resize_bilinear_antialias_u8(const uint8_t* img_u8_hwc, int H, int W, int out_h, int out_w, float* scratch_f32,
                                     uint8_t* out_u8_hwc, void* stream) {
  return cvb::resize_bilinear_antialias_u8((cudaStream_t)stream, img_u8_hwc, H, W, out_h, out_w, scratch_f32, out_u8_hwc);
}

int cvb_preprocess_verifier_image(const uint8_t* img_u8_hwc, int H, int W, int out_h, int out_w, uint8_t* out_u8_hwc,
                                  float* out_f32_chw, void* stream) {
  return cvb::preprocess_verifier_image((cudaStream_t)stream, img_u8_hwc, H, W, out_h, out_w, out_u8_hwc, out_f32_chw);
}

}  // extern "C"
