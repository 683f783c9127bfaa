// pi0-specific fused kernels (ops_pi0.cu).  Reference lines each one replaces:
//   im2col_patches     SiglipVisionEmbeddings patch conv reached via embed_image, paligemma_with_expert.py:229-230
//   build_prefix       modeling_pi0.py:533-538 (image embedding rescale), :549-553 (token gather * sqrt(d))
//   rope_qkv           paligemma_with_expert.py:288-299 (apply_rope :34-57, KV-cache fill)
//   action_out_euler   modeling_pi0.py:748-752 (+ Euler update :713)
//   fill_state_rows    modeling_pi0.py:577-581 (state token of the suffix)
#pragma once
#include "ops.h"

namespace cvb {

int im2col_patches(cudaStream_t st, const float* img, bf16* out, int C, int H, int W, int P, int kpad, int images = 1);
int build_prefix(cudaStream_t st, const bf16* proj, const bf16* embed, const int64_t* tok,
                 bf16* prefix, int R, int n_img, int n_lang, int tok_stride, int D, int rephrases_per_obs = 0);
int rope_qkv(cudaStream_t st, bf16* qkv, long ld, const float* timescale, int rows, int heads,
             int hd, int rows_per_batch, const int* pos_base_dev, int q_per_kv_batch, bf16* kcache,
             bf16* vcache, long cache_bs, long cache_rs, bf16* vt = nullptr, long vt_bs = 0, long vt_ld = 0);
int action_out_euler(cudaStream_t st, const bf16* hn, long ld, const float* w, const float* bias,
                     float* x_t, float* v_out, int n_cand, int width, int adim, int chunk,
                     int suffix_len, float dt);
int fill_state_rows(cudaStream_t st, const float* state_emb, float* suffix, int n_cand, int width,
                    int suffix_len, int cands_per_obs = 0);
int rope_table(cudaStream_t st, const float* timescale, const int* pos_base_dev, int batches, int tq, int half,
               float2* tab);
int prefix_lengths(cudaStream_t st, const int* lang_len, int* plen, int R, int n_img, int max_lang);
int bf16_to_f32(cudaStream_t st, const bf16* src, float* dst, long n);

}  // namespace cvb
