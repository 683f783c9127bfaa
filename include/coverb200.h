/* coverb200 - C ABI of the B200-native CoVer-VLA sample-and-verify path.
 *
 * The reference (cover-vla/cover-vla) has no FFI: its boundary for this path is two Python objects,
 *   PI0Policy.select_action / PI0FlowMatching.sample_actions
 *       lerobot_custom/lerobot/common/policies/pi0/modeling_pi0.py:263-307, :672-715
 *   EfficientEnsembleMerged.compute_max_similarity_scores_batch
 *       bridge_verifier/ensemble_eval/efficient_ensemble_merged.py:309-454
 * called from CoVer_VLA/inference/experiments/robot/simpler/run_simpler_eval_with_openpi.py:324,346-363.
 * This header is what a ctypes / cffi / pybind stub binds instead.
 *
 * Conventions: plain pointers and sizes only; every tensor is caller-owned DEVICE memory unless the
 * parameter name ends in _host; every call is asynchronous on `stream` (a cudaStream_t passed as
 * void*); return 0 = OK, negative = error with the text in cvb_last_error() (thread-local).
 * No call synchronises the device, allocates per call, or throws across the boundary.
 */
#ifndef COVERB200_H_
#define COVERB200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CVB_ABI_VERSION 3
#if defined(__GNUC__)
#define CVB_API __attribute__((visibility("default")))
#else
#define CVB_API
#endif

CVB_API const char* cvb_last_error(void);
CVB_API int cvb_abi_version(void);
/* Kernels enqueued by this library so far (launches recorded into a CUDA graph count once, at capture). */
CVB_API int64_t cvb_launch_count(void);

/* ---------------------------------------------------------------------------------------------
 * Operator level (one launch each).  GEMM epilogue kinds:
 *   0 store bf16(acc+bias) | 1 bf16(gelu_tanh(bf16(acc+bias))) | 2 bf16(bf16(acc+bias)+resid)
 *   3 GeGLU on packed [128 gate | 128 up] weight rows | 4 store fp32(acc+bias)
 *   5 GeGLU on packed [64 gate | 64 up] rows (skinny kernel, M <= 256)
 *   6 split-K partials (M <= 256): C = fp32 [S][M][ldc], S = -force_bn clamped to ceil(K/64) (0: fill the SMs once);
 *     no bias / resid - the S partials are summed by cvb_op_rmsnorm_reduce
 * Replaces nn.Linear (+ the elementwise op that follows it) wherever it appears on the path, e.g.
 * paligemma_with_expert.py:273-276 (q/k/v), :327-333 (o_proj + residual), :335-341 (MLP + residual).
 */
CVB_API int cvb_op_gemm_bf16(const void* A, int64_t lda, const void* W, int64_t ldw, int M, int N, int K,
                     int epilogue, void* C, int64_t ldc, const void* bias, int bias_is_f32,
                     const void* resid, int resid_is_f32, int64_t ldr, int n_out,
                     const int32_t* m_dev, int force_bn, void* stream);

/* Tail of a split-K linear layer fused with the Gemma RMSNorm that follows it (paligemma_with_expert.py:327-355:
 * o_proj + residual -> post_attention_layernorm, down_proj + residual -> next input_layernorm):
 *   h = bf16(bf16(sum_s P[s]) + resid)  (summed in split order; h_out may alias resid),  y = bf16(h * rsqrt(mean(h^2) + eps) * (1 + w)).
 * P = fp32 [S][rows][ldp] written by cvb_op_gemm_bf16(epilogue 6); resid bf16 or fp32; w bf16 or fp32 [width]. */
CVB_API int cvb_op_rmsnorm_reduce(const float* P, int S, int64_t split_stride, int64_t ldp, const void* resid,
                                  int resid_is_f32, int64_t ldr, const void* w, int w_is_f32, void* h_out, int64_t ldh,
                                  void* y, int64_t ldy, int rows, int width, float eps, void* stream);

/* Gemma RMSNorm alone (transformers GemmaRMSNorm, used at paligemma_with_expert.py:268,335,355):
 *   y = bf16(x * rsqrt(mean(x^2) + eps) * (1 + w)), statistics in fp32; x bf16 or fp32 [rows, ldx]; w bf16 or fp32 [width]. */
CVB_API int cvb_op_rmsnorm(const void* x, int x_is_f32, int64_t ldx, const void* w, int w_is_f32, void* y, int64_t ldy,
                           int rows, int width, float eps, void* stream);

/* LayerNorm variant (SigLIP encoder block, reached through embed_image, paligemma_with_expert.py:229-230):
 *   h = bf16(bf16(sum_s P[s] + bias) + resid),  y = LayerNorm(h) * w + b   (bias may be NULL; bias / resid / w / b bf16). */
CVB_API int cvb_op_layernorm_reduce(const float* P, int S, int64_t split_stride, int64_t ldp, const void* bias,
                                    const void* resid, int64_t ldr, const void* w, const void* b, void* h_out,
                                    int64_t ldh, void* y, int64_t ldy, int rows, int width, float eps, void* stream);

/* fp32 linear for the parts of the path the reference keeps in float32 (modeling_pi0.py:598-609 suffix MLP,
 * efficient_ensemble_merged.py:194-247 verifier heads): C = act(A[M,K] . W[N,K]^T + bias + row_bias) + resid, true-fp32
 * FFMA accumulation (no TF32).  act: 0 none | 1 relu | 2 gelu(erf) | 3 silu.  Any pointer but A / W / C may be NULL. */
CVB_API int cvb_op_sgemm_f32(const float* A, int64_t lda, const float* W, int64_t ldw, int M, int N, int K, float* C,
                             int64_t ldc, const float* bias, const float* row_bias, const float* resid, int64_t ldr,
                             int act, void* stream);

/* fp32-accurate linear layers on the bf16 tensor cores: out bf16 [rows, 3K] = [hi | hi | lo] of x fp32 [rows, K]
 * (activation layout; weight_layout = 1: [hi | lo | hi]), hi = bf16(x), lo = bf16(x - hi); relu = 1 applies ReLU first.
 * cvb_op_gemm_bf16 (epilogue 4, fp32 bias) over the 3K axis of an activation-layout A and a weight-layout W then evaluates
 * a_hi w_hi + a_hi w_lo + a_lo w_hi with fp32 accumulation: relative error ~1e-5 of |a||w| per product, far below the
 * 1e-3 score tolerance.  Used for the verifier's trajectory encoder (efficient_ensemble_merged.py:229-245) and
 * action_time_mlp_out (modeling_pi0.py:607-609), which the reference keeps in float32. */
CVB_API int cvb_op_split3_f32(const float* x, int64_t ldx, void* out_bf16, int64_t rows, int K, int weight_layout, int relu,
                              void* stream);

/* Exact-softmax attention with the reference's rounding ledger (eager_attention_forward,
 * paligemma_with_expert.py:376-434): q [batches, tq, heads*head_dim] (strides q_bs / q_rs in elements), keys in two
 * segments - segment 0 shared per kv batch (kv batch = batch / q_per_kv_batch; length from kv0_len_dev[kv batch] or
 * kv0_len), segment 1 per batch (e.g. the suffix tokens' own keys) with the pi0 suffix mask when suffix_mask = 1.
 * force_two_pass: 0 auto (tcgen05 kernel when the shape allows it, else the exact mma.sync kernel), 1 the two-pass variant
 * of the mma.sync kernel, 4 tcgen05 only (error if the shape is not eligible).
 * rope_cos_sin (optional, f32 [kv batches, tq, head_dim/2, 2]): RoPE applied to q and the segment-1 keys while staging. */
CVB_API int cvb_op_attention(const void* q, int64_t q_bs, int64_t q_rs, const void* k0, const void* v0, int64_t kv0_bs,
                             int64_t kv0_rs, const int32_t* kv0_len_dev, int kv0_len, int kv0_max,
                             int q_per_kv_batch, const void* k1, const void* v1, int64_t kv1_bs, int64_t kv1_rs,
                             int kv1_len, int suffix_mask, void* out, int64_t o_bs, int64_t o_rs, int batches,
                             int heads, int kv_heads, int tq, int head_dim, float scale, int force_two_pass,
                             const float* rope_cos_sin, void* stream);

/* cvb_op_attention plus an optional TRANSPOSED copy of the segment-0 values, vt0[(kv batch * head_dim + d) * vt0_ld + key]
 * (finite past the valid length).  With it, multi-query head_dim-256 two-segment calls (the denoise step, heads * tq <=
 * 128, <= 8 segment-1 keys) run on the tcgen05 / TMEM kernel; force_two_pass = 4 requires that kernel (error if the
 * shape is not eligible). */
CVB_API int cvb_op_attention_tc(const void* q, int64_t q_bs, int64_t q_rs, const void* k0, const void* v0,
                                int64_t kv0_bs, int64_t kv0_rs, const int32_t* kv0_len_dev, int kv0_len, int kv0_max,
                                int q_per_kv_batch, const void* k1, const void* v1, int64_t kv1_bs, int64_t kv1_rs,
                                int kv1_len, int suffix_mask, void* out, int64_t o_bs, int64_t o_rs, int batches,
                                int heads, int kv_heads, int tq, int head_dim, float scale, int force_two_pass,
                                const float* rope_cos_sin, const void* vt0, int64_t vt0_ld, void* stream);

/* tcgen05 / TMEM prefix attention (PaliGemma prefix pass, paligemma_with_expert.py:236-360, eager_attention_forward
 * :376-434): multi-query, heads = 8, head_dim = 256, <= 384 keys, every query token of a batch attends the first
 * klen_dev[b] (or klen) keys.  q: token (b * q_rows_per_batch + t) at q + row * q_ld, head h at column h * 256;
 * k: [k_total_rows, 256] row-major, batch b from row b * k_rows_per_batch; vt = V transposed, vt[(b * 256 + d) * vt_ld +
 * key] (columns past the valid length must be finite); out + b * o_bs + t * o_rs + h * 256.  Returns -1 (with
 * cvb_last_error) for shapes outside these limits. */
CVB_API int cvb_op_attention_umma(const void* q, int64_t q_ld, int64_t q_total_rows, int64_t q_rows_per_batch,
                                  const void* k, int64_t k_total_rows, int64_t k_rows_per_batch, const void* vt,
                                  int64_t vt_ld, const int32_t* klen_dev, int klen, int kmax, void* out, int64_t o_bs,
                                  int64_t o_rs, int batches, int tq, int heads, int head_dim, float scale, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Engine level.  One handle per (device, stream); a handle is not thread-safe, the library is
 * re-entrant across handles.  All sizes are configuration, nothing is hard-coded to the Bridge
 * checkpoint (chunk_size 4, 72 language tokens, 10 steps) - see SURVEY.md section 8.
 */
typedef struct cvb_config {
  int32_t struct_size; /* = sizeof(cvb_config), ABI guard */
  /* pi0: SigLIP tower (paligemma_with_expert.py:103-115) */
  int32_t vis_layers, vis_width, vis_heads, vis_mlp, vis_patch, vis_image;
  /* pi0: PaliGemma LM + action expert (paligemma_with_expert.py:91-102,127-150) */
  int32_t layers, lm_width, lm_mlp, heads, head_dim, ex_width, ex_mlp, vocab;
  /* pi0: PI0Config (configuration_pi0.py:27-80) */
  int32_t max_state_dim, max_action_dim, chunk_size, max_lang_len, num_steps;
  /* workspace sizing: calls may use any R <= max_rephrases, K <= max_samples */
  int32_t max_rephrases, max_samples;
  /* verifier trunk (SigLIP2 ViT-L/16-384 + text tower) and heads; vf_members == 0 disables it */
  int32_t vf_image, vf_patch, vf_width, vf_layers, vf_heads, vf_mlp;
  int32_t vf_text_layers, vf_text_ctx, vf_vocab;
  /* vf_traj_layers >= 1: transformer action encoder (use_transformer = True, efficient_ensemble_merged.py:135-160) with
   * feed-forward width vf_traj_ff.  vf_traj_layers == 0: the MLP `complex_action_encoder` of use_transformer = False
   * checkpoints (efficient_ensemble_merged.py:161-171) with hidden width vf_traj_ff (512 in the reference). */
  int32_t vf_members, vf_embed, vf_pool_heads, vf_pool_layers, vf_traj_layers, vf_traj_ff;
  int32_t vf_history, vf_action_dim;
  int32_t use_cuda_graph; /* capture the per-(R,K) launch sequence once and replay it */
  /* observations per batched call (cvb_pi0_sample_batch / cvb_cover_step_batch); 0 or 1 = single-observation handle.
   * Workspace (KV caches, activations) is sized for max_observations * max_rephrases prompts. */
  int32_t max_observations;
  /* cameras (image streams) per observation the handle is sized for, modeling_pi0.py:344-387; 0 or 1 = one camera.
   * With C cameras every `image` argument below is f32 [C, 3, vis_image, vis_image] per observation and a prompt holds
   * C * (vis_image / vis_patch)^2 image tokens in camera order (modeling_pi0.py:529-547). */
  int32_t num_cameras;
} cvb_config;

typedef struct cvb_handle cvb_handle;

enum { CVB_F32 = 0, CVB_BF16 = 1, CVB_I64 = 2, CVB_I32 = 3, CVB_U8 = 4 };

CVB_API int cvb_create(const cvb_config* cfg, cvb_handle** out);
CVB_API void cvb_destroy(cvb_handle* h);

/* Borrow a weight: `key` is the reference state-dict name (PI0Policy.state_dict() with or without the
 * leading "model.", transformers 4.48.3 or >=4.52 module layout - SURVEY.md Appendix C; verifier
 * tensors as "verifier.<member>.<component>.<name>" and trunk tensors as "verifier.trunk.<name>").
 * The pointer must stay valid until cvb_finalize() for tensors that get repacked (q/k/v, gate/up,
 * patch embedding) and until cvb_destroy() for all others. */
CVB_API int cvb_bind_weight(cvb_handle* h, const char* key, const void* dev_ptr, int dtype, int ndim,
                            const int64_t* shape);
/* Number of weights the configuration requires; names retrievable one by one (for loaders / tests). */
CVB_API int cvb_required_weight_count(cvb_handle* h);
CVB_API const char* cvb_required_weight_name(cvb_handle* h, int index);
/* dtype (CVB_F32 / CVB_BF16) the engine expects for required weight `index` - what the reference's
 * to_bfloat16_like_physical_intelligence (paligemma_with_expert.py:216-227) leaves it in; loaders cast to it. */
CVB_API int cvb_required_weight_dtype(cvb_handle* h, int index);
/* Validate that every required weight is bound, repack, allocate the workspace. */
CVB_API int cvb_finalize(cvb_handle* h, void* stream);

/* Sample N = R*K action chunks for ONE observation (replaces PI0FlowMatching.sample_actions,
 * modeling_pi0.py:672-715, for the batch run_simpler_eval_with_openpi.py:305-319 builds).
 *   image      f32 [3, vis_image, vis_image] in [-1, 1]      (identical for all candidates)
 *   lang_tokens i64 [R, max_lang_len], lang_len i32 [R]       (right-padded; valid prefix length)
 *   state      f32 [max_state_dim]
 *   noise      f32 [R*K, chunk_size, max_action_dim]          (rephrase-major, as the reference batches)
 *   actions    f32 [R*K, chunk_size, max_action_dim]          (output x_0)
 */
CVB_API int cvb_pi0_sample(cvb_handle* h, const float* image, const int64_t* lang_tokens,
                           const int32_t* lang_len, const float* state, const float* noise, int R,
                           int K, float* actions, void* stream);

/* B independent observations in ONE pass (the episode-batched driver around
 * run_simpler_eval_with_openpi.py:190-449; BASELINE.json configs[4]).  Same arithmetic per observation as B calls of
 * cvb_pi0_sample - each observation's rephrases attend only its own image, each candidate only its own rephrase's
 * KV cache - but every weight is streamed once for all B * R * K candidates, so the denoise GEMMs run on
 * 5 * B * R * K rows instead of 5 * R * K.  Layouts are observation-major:
 *   images f32 [B, 3, vis_image, vis_image]; lang_tokens i64 [B*R, max_lang_len]; lang_len i32 [B*R];
 *   states f32 [B, max_state_dim]; noise / actions f32 [B*R*K, chunk_size, max_action_dim].
 * Requires B <= cfg.max_observations. */
CVB_API int cvb_pi0_sample_batch(cvb_handle* h, int B, const float* images, const int64_t* lang_tokens,
                                 const int32_t* lang_len, const float* states, const float* noise, int R, int K,
                                 float* actions, void* stream);

/* Optional bound on the number of VALID language tokens per prompt for the following cvb_pi0_sample calls (0 = none:
 * max_lang_len rows are processed).  Right-padded tokens are masked as keys and never read (exact),
 * so a host that knows its tokenizer output (it produced it) lets the prefix skip them; longer prompts are truncated
 * to the hint, as the reference truncates at tokenizer_max_length (modeling_pi0.py:389-409). */
CVB_API int cvb_pi0_set_lang_len_hint(cvb_handle* h, int max_valid_tokens);

/* Number of cameras the following calls pass per observation (1 .. num_cameras; 0 = num_cameras).  The reference fills
 * missing cameras with -1 images whose mask is False (prepare_images, modeling_pi0.py:377-385): their tokens are masked
 * as keys, do not advance the position ids (cumsum of the pad mask, :684) and are never read - the host simply drops
 * them and passes the present cameras, which is exact. */
CVB_API int cvb_pi0_set_active_cameras(cvb_handle* h, int cameras);

/* Profiling hook: re-run one phase (0 vision tower, 1 prefix, 2 denoise loop) eagerly on the inputs staged by
 * the last cvb_pi0_sample call, so a host can time the phases separately with CUDA events. */
CVB_API int cvb_pi0_run_phase(cvb_handle* h, int phase, int R, int K, void* stream);
/* same for the inputs staged by the last cvb_pi0_sample_batch / cvb_cover_step_batch call with B observations */
CVB_API int cvb_pi0_run_phase_batch(cvb_handle* h, int phase, int B, int R, int K, void* stream);

/* Score N = R*K candidate action histories against ONE (image, instruction) pair and select
 * (replaces EfficientEnsembleMerged.compute_max_similarity_scores_batch, efficient_ensemble_merged.py:309-454;
 * the fast path of :330-347 - pair 0 only - is the only one whose result the reference consumes, :422-425).
 *   image       f32 [3, vf_image, vf_image]  (open_clip-preprocessed, normalised)
 *   text_tokens i64 [vf_text_ctx]
 *   traj        f32 [N, vf_history, vf_action_dim], left-padded with -5 rows (:379-390)
 *   scores f32 [N]; group_mean f32 [R] (may be NULL); best_idx i32 [1]; best_score f32 [1]
 * With R == 0 only the scores are produced (multi-GPU: gather them, then call cvb_select).
 * recompute_context = 0 reuses the image/text side (trunk + image-text heads) of the previous call. */
CVB_API int cvb_verifier_score(cvb_handle* h, const float* image, const int64_t* text_tokens, const float* traj,
                               int N, int R, int K, float* scores, float* group_mean, int32_t* best_idx,
                               float* best_score, int recompute_context, void* stream);
/* Only the image/text side of cvb_verifier_score (trunk + image-text heads of every member).  It does not depend
 * on the sampled actions, so a host can enqueue it on a second stream while cvb_pi0_sample runs, then call
 * cvb_verifier_score(..., recompute_context = 0) after joining the streams. */
CVB_API int cvb_verifier_context(cvb_handle* h, const float* image, const int64_t* text_tokens, void* stream);
/* Per-task prompt cache: the instruction of an episode is static between the swaps of
 * run_simpler_eval_with_openpi.py:409, and the text tower's output depends on nothing else.  hold = 1: the caller vouches
 * that the text tokens of the following context computations (cvb_verifier_context / _score / cvb_cover_step[_batch]) equal
 * those of the previous one - the resident text features are reused and the text tower (24 of the context's 48 transformer
 * blocks) is skipped; a call with a different observation count than the resident features still computes it.  hold = 0
 * (default): every context call runs the text tower, like the reference. */
CVB_API int cvb_verifier_hold_text(cvb_handle* h, int hold);

/* Sampler -> verifier action formatting on the device (replaces process_inputs(verifier_action=True),
 * eval_utils.py:172-221, BridgeSimplerAdapter.postprocess_verifier, INT-ACT/src/experiments/env_adapters/
 * simpler.py:96-121, and the -5 left-padding of efficient_ensemble_merged.py:379-390).
 *   actions f32 [n_cand, chunk, action_stride] (policy output, dims 0..6 used; first n_future steps)
 *   p01_host / p99_host: HOST double[6] action statistics (bridge_statistics.json "action" p01 / p99)
 *   past    f32 [num_past, 7] already in verifier format (the caller's action_history tail)
 *   traj    f32 [n_cand, history, 7] */
CVB_API int cvb_format_trajectories(const float* actions, int n_cand, int chunk, int action_stride,
                                    const double* p01_host, const double* p99_host, const float* past,
                                    int num_past, int history, int n_future, float* traj, void* stream);
/* Execution-format action of the selected candidate + gripper vote of its K-sample group, on the device (replaces
 * process_inputs(verifier_action=False) for the winning group, eval_utils.py:172-221 -> BridgeSimplerAdapter.postprocess
 * INT-ACT/src/experiments/env_adapters/simpler.py:123-166 [denormalize_bound base.py:20-31, euler2axangle
 * src/utils/geometry.py:261-436 'sxyz', postprocess_gripper :211-220], and the vote of
 * run_simpler_eval_with_openpi.py:368-391).
 *   actions f32 [n_cand, chunk, action_stride] (policy output); best_idx i32 [1] (device, e.g. from cvb_select);
 *   step = which chunk step is executed (0 in the reference); exec_action f64 [7] = (xyz, axis * angle, gripper +-1);
 *   votes i32 [2] = (close, open) or NULL.  float64 with numpy's operation order. */
CVB_API int cvb_execution_action(const float* actions, int n_cand, int chunk, int action_stride,
                                 const double* p01_host, const double* p99_host, const int32_t* best_idx, int K,
                                 int step, double* exec_action, int32_t* votes, void* stream);
/* Policy-side observation pre-processing on the device (replaces BridgeSimplerAdapter.preprocess,
 * INT-ACT/src/experiments/env_adapters/simpler.py:43-65: cv2.resize(frame, (out_w, out_h), interpolation=cv2.INTER_LANCZOS4)
 * then process_images, src/utils/pipeline.py:34-69).  img_u8_hwc: device uint8 [H, W, 3]; out_u8_hwc: device uint8
 * [out_h, out_w, 3] (bit-exact with cv2, may be NULL); out_f32_chw: device f32 [3, out_h, out_w] = (u8 / 255 - 0.5) / 0.5
 * (may be NULL).  Coefficient tables are computed once per geometry (host, synchronous) and cached. */
CVB_API int cvb_preprocess_policy_image(const uint8_t* img_u8_hwc, int H, int W, int out_h, int out_w,
                                        uint8_t* out_u8_hwc, float* out_f32_chw, void* stream);
/* Verifier-side image transform on the device (replaces the open_clip transform the reference applies at
 * bridge_verifier/ensemble_eval/efficient_ensemble_merged.py:249-254: PIL resize((out_w, out_h), BICUBIC), ToTensor,
 * Normalize(0.5, 0.5)).  Same conventions as cvb_preprocess_policy_image; out_u8_hwc is bit-exact with Pillow. */
CVB_API int cvb_preprocess_verifier_image(const uint8_t* img_u8_hwc, int H, int W, int out_h, int out_w,
                                          uint8_t* out_u8_hwc, float* out_f32_chw, void* stream);
/* The first verifier-side frame step on the device (replaces process_raw_image_to_jpg, CoVer_VLA/inference/experiments/
 * robot/simpler/eval_utils.py:228-286: tf.image.resize(frame, (256, 256), BILINEAR, antialias=True) then tf.cast(uint8)).
 * img_u8_hwc: device uint8 [H, W, 3]; scratch_f32: device float [out_h * W * 3]; out_u8_hwc: device uint8 [out_h, out_w, 3].
 * TensorFlow's scale_and_translate algorithm restated (triangle kernel, antialias); bit-exact against the numpy
 * restatement in oracle/, unpinned against TensorFlow itself (absent offline). */
CVB_API int cvb_resize_bilinear_antialias_u8(const uint8_t* img_u8_hwc, int H, int W, int out_h, int out_w,
                                             float* scratch_f32, uint8_t* out_u8_hwc, void* stream);
/* One whole CoVer decision in one call / one CUDA graph: cvb_pi0_sample -> cvb_format_trajectories ->
 * cvb_verifier_score for N = R*K candidates (the body of run_simpler_eval_with_openpi.py:322-363 on the device, no host
 * round trip in between).  The verifier's image/text side is forked onto an internal stream after the prefix and runs
 * concurrently with the denoise loop.  Arguments as in the three calls it fuses; past f32 [num_past, 7] (may be NULL
 * when num_past == 0); outputs actions [N, chunk, max_action_dim], traj [N, vf_history, 7], scores [N],
 * group_mean [R], best_idx, best_score - any output pointer may be NULL. */
CVB_API int cvb_cover_step(cvb_handle* h, const float* image, const int64_t* lang_tokens, const int32_t* lang_len,
                           const float* state, const float* noise, int R, int K, const float* vf_image,
                           const int64_t* vf_text_tokens, const double* p01_host, const double* p99_host,
                           const float* past, int num_past, int n_future, float* actions, float* traj, float* scores,
                           float* group_mean, int32_t* best_idx, float* best_score, void* stream);

/* cvb_cover_step for B independent observations in one graph (configs[4]): sampler batched as in cvb_pi0_sample_batch,
 * one verifier context per observation (vf_images f32 [B, 3, vf_image, vf_image], vf_text_tokens i64 [B, vf_text_ctx])
 * on the forked stream, the trajectory encoders over all B * N histories at once, one score / select group per
 * observation.  past f32 [B, num_past, 7]; outputs actions [B*N, chunk, max_action_dim], traj [B*N, vf_history, 7],
 * scores [B*N], group_mean [B*R], best_idx i32 [B] (index INSIDE the observation's N candidates), best_score [B]. */
CVB_API int cvb_cover_step_batch(cvb_handle* h, int B, const float* images, const int64_t* lang_tokens,
                                 const int32_t* lang_len, const float* states, const float* noise, int R, int K,
                                 const float* vf_images, const int64_t* vf_text_tokens, const double* p01_host,
                                 const double* p99_host, const float* past, int num_past, int n_future, float* actions,
                                 float* traj, float* scores, float* group_mean, int32_t* best_idx, float* best_score,
                                 void* stream);

/* Test hook: inject normalised trunk features (patch f32 [Np, W], text f32 [ctx, W]) and recompute the
 * image-text heads, so the fp32 heads can be checked in isolation from the bf16 trunk. */
CVB_API int cvb_verifier_set_features(cvb_handle* h, const float* patch, const float* text, void* stream);
/* group-mean -> argmax group -> argmax inside the group over a (gathered) score vector (:417-447). */
CVB_API int cvb_select(const float* scores, int R, int K, float* group_mean, int32_t* best_idx, float* best_score,
                       void* stream);

/* ---------------------------------------------------------------------------------------------
 * Multi-GPU: fused score all-gather + selection over NVLink peer memory (BASELINE.json
 * configs[3]: one observation, rephrases sharded over the ranks, one process per GPU).  Each rank owns a mailbox in its
 * HBM; cvb_allgather_select is ONE kernel per rank that stores its slice into every peer's mailbox (peer-mapped
 * st.global over NVLink), publishes an epoch flag, waits for the other ranks' flags and runs the selection - no NCCL
 * call, no host round trip.  Setup (once): cvb_comm_create on every rank, exchange the cvb_comm_handle_bytes()-byte
 * handles of cvb_comm_local_handle through any host channel (torch.distributed in cover_vla_b200/comm.py), then
 * cvb_comm_open_peers with all `world` handles in rank order.
 *   Shards are contiguous rephrase ranges: rank r owns R/world (+1 for r < R % world) rephrases, K candidates each.
 *   local_scores f32 [n_loc]; local_actions f32 [n_loc, act_floats] or NULL; scores f32 [R*K]; actions f32
 *   [R*K, act_floats] or NULL; group_mean f32 [R] or NULL; best_idx i32 [1]; best_score f32 [1] - identical on all ranks. */
typedef struct cvb_comm cvb_comm;
CVB_API int cvb_comm_create(int rank, int world, int max_slot_floats, cvb_comm** out);
CVB_API int cvb_comm_handle_bytes(void);
CVB_API int cvb_comm_local_handle(cvb_comm* comm, void* handle_out_host);
CVB_API int cvb_comm_open_peers(cvb_comm* comm, const void* all_handles_host);
CVB_API void cvb_comm_destroy(cvb_comm* comm);
CVB_API int cvb_allgather_select(cvb_comm* comm, const float* local_scores, const float* local_actions, int act_floats,
                                 int R, int K, float* scores, float* actions, float* group_mean, int32_t* best_idx,
                                 float* best_score, void* stream);

/* Test/diagnostic tap: copy an internal buffer ("image_emb", "prefix_k0", "prefix_vlast", "v0",
 * "time_emb", ...) to dst (device).  Returns the number of bytes copied or a negative error. */
CVB_API int64_t cvb_debug_copy(cvb_handle* h, const char* name, void* dst, int64_t max_bytes,
                               void* stream);

/* Diagnostics: when non-NULL, every CTA of the skinny GEMM writes 8 globaltimer phase stamps to dev_u64[cta * 8 + i]. */
CVB_API void cvb_debug_set_timestamps(void* dev_u64);

/* Host-only helpers (no CUDA calls): the constants of the denoise loop, for tests and hosts. */
CVB_API int cvb_denoise_times_host(int num_steps, float* times_out, int max_out, float* dt_out);
CVB_API void cvb_time_embedding_host(float t, int dim, double min_period, double max_period,
                                     uint16_t* out_bf16);

#ifdef __cplusplus
}
#endif
#endif /* COVERB200_H_ */
