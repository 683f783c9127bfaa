// Verifier-specific kernels (ops_verifier.cu).  Reference lines each one replaces:
//   embed_tokens_pos          open_clip TextTransformer token + positional embedding (third-party; restated in oracle/verifier_oracle.py)
//   l2norm_rows_bf16_to_f32   finetune_trajectory_bridge_ddp.py:329-330, 352-354
//   softmax_rows_temp/add_f32 TextAwareVisualExtraction.forward, model.py:63-71
//   pool_chains               AttentionPooling.forward + CrossAttentionBlock.forward, model.py:97-112, 25-38
//   it_finalize               efficient_ensemble_merged.py:220-223
//   traj_attention            nn.TransformerEncoderLayer self-attention with key padding, efficient_ensemble_merged.py:229-235
//   masked_mean_l2norm        efficient_ensemble_merged.py:236-245
//   format_trajectories       eval_utils.py:172-221 + INT-ACT simpler.py:96-121 + efficient_ensemble_merged.py:379-390
//   fuse_score_select         efficient_ensemble_merged.py:404-447 (fuse, scores, group mean, argmax)
#pragma once
#include "ops.h"

namespace cvb {

constexpr int kMaxPoolLayers = 8;

struct PoolBlockW {
  const float *qln_w, *qln_b, *wq, *b_in, *wo, *bo, *ln_w, *ln_b, *fc1_w, *fc1_b, *fc2_w, *fc2_b;
};
struct PoolChain {
  const float* query;
  const float* kv;  // [tokens, kv_ld]: K of block l at columns [l*2E, l*2E+E), V right after
  int kv_ld, embed, heads, tokens, layers;
  const float *fin_w, *fin_b;
  float* out;
  PoolBlockW blk[kMaxPoolLayers];
};
struct ItFinal {
  const float *text_tok, *vision_tok, *w, *b;
  float* out;
};

struct FormatStats {
  double p01[6], p99[6];
};
int format_trajectories(cudaStream_t stream, const float* actions, int n_cand, int chunk, int adim_stride,
                        const FormatStats& st, const float* past, int num_past, int history, int n_future,
                        float* traj, int cands_per_obs = 0);  // cands_per_obs > 0: past is [n_obs][num_past][7]
int execution_action(cudaStream_t stream, const float* actions, int n_cand, int chunk, int adim_stride, const FormatStats& st,
                     const int* best_idx, int K, int step, double* out, int* votes);

int embed_tokens_pos(cudaStream_t st, const bf16* table, const bf16* pos, const int64_t* tok, bf16* out,
                     int tokens, int width, int ctx = 0);  // ctx > 0: position = token index mod ctx
int l2norm_rows_bf16_to_f32(cudaStream_t st, const bf16* x, long ldx, float* y, int rows, int width);
int l2norm_rows_f32(cudaStream_t st, const float* x, float* y, int rows, int width);
int softmax_rows_temp(cudaStream_t st, float* x, int rows, int cols, const float* temp_dev);
int add_f32(cudaStream_t st, const float* a, const float* b, float* y, long n);
int pool_chains(cudaStream_t st, const PoolChain* chains_dev, int n_chains, int embed, int heads, int tokens);
int it_finalize(cudaStream_t st, const ItFinal* items_dev, int members, int embed);
int traj_attention(cudaStream_t st, const float* qkv, const float* traj, float* out, int n_cand, int S, int E, int H,
                   int adim, float pad_value);
int masked_mean_l2norm(cudaStream_t st, const float* x, const float* traj, float* out, int n_cand, int S, int E,
                       int adim, float pad_value);
// n_obs observations: it [n_obs][M][E], act [M][n_obs * N][E] (member stride act_member_stride, 0 = contiguous),
// scores [n_obs * N], group_mean [n_obs * R], best_idx / best_score [n_obs]
int fuse_score_select(cudaStream_t st, const float* it, const float* act, int M, int N, int E, float* scores, int R,
                      int K, float* group_mean, int* best_idx, float* best_score, int do_select, int n_obs = 1,
                      long act_member_stride = 0);
int select_best(cudaStream_t st, const float* scores, int R, int K, float* group_mean, int* best_idx,
                float* best_score, int n_obs = 1);

}  // namespace cvb
