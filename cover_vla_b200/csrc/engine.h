// Engine internals shared by engine_pi0.cu / engine_verifier.cu / api_engine.cu.
#pragma once
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/coverb200.h"
#include "expert_mega.h"
#include "host_common.h"
#include "ops.h"

namespace cvb {

struct Weight {
  const void* ptr = nullptr;
  int dtype = 0;
  std::vector<int64_t> shape;
  int64_t numel() const {
    int64_t n = 1;
    for (auto s : shape) n *= s;
    return n;
  }
};

struct WeightSpec {
  std::string key;
  int dtype;
  std::vector<int64_t> shape;
};

struct VisLayer {
  const bf16 *ln1_w, *ln1_b, *ln2_w, *ln2_b;
  bf16* wqkv;  // owned [3W, W]
  bf16* bqkv;  // owned [3W]
  const bf16 *wo, *bo, *w1, *b1, *w2, *b2;
};
struct GemmaLayer {
  const bf16 *in_norm, *post_norm;
  bf16* wqkv;  // owned [(H+2)*hd, D]
  const bf16* wo;
  bf16* wgu;  // owned packed gate|up blocks: [2*half*ceil(I/half), D], half = 128 (prefix) or 64 (expert)
  const bf16* wd;
};

// Per-shape CUDA-graph cache: the first call of a shape runs eagerly (sets kernel attributes, fills the
// tensor-map cache), the second is captured on a private stream (the caller's may be the legacy default
// stream, which cannot be captured), later calls replay the instantiated graph on the caller's stream.
struct GraphCache {
  std::unordered_map<long, cudaGraphExec_t> graphs;
  std::unordered_map<long, int> warm;
  cudaStream_t cap_stream = nullptr;
  template <typename F>
  int run(bool enabled, long key, cudaStream_t st, F&& body) {
    if (!enabled) return body(st);
    auto it = graphs.find(key);
    if (it != graphs.end()) {
      CVB_CUDA(cudaGraphLaunch(it->second, st));
      return 0;
    }
    if (warm[key] == 0) {
      warm[key] = 1;
      return body(st);
    }
    if (cap_stream == nullptr) CVB_CUDA(cudaStreamCreateWithFlags(&cap_stream, cudaStreamNonBlocking));
    cudaGraph_t graph = nullptr;
    CVB_CUDA(cudaStreamBeginCapture(cap_stream, cudaStreamCaptureModeThreadLocal));
    const int rc = body(cap_stream);
    const cudaError_t e = cudaStreamEndCapture(cap_stream, &graph);
    if (rc != 0) {
      if (graph) cudaGraphDestroy(graph);
      return rc;
    }
    CVB_CUDA(e);
    cudaGraphExec_t exec = nullptr;
    CVB_CUDA(cudaGraphInstantiate(&exec, graph, 0));
    cudaGraphDestroy(graph);
    graphs[key] = exec;
    CVB_CUDA(cudaGraphLaunch(exec, st));
    return 0;
  }
  void destroy() {
    for (auto& g : graphs) cudaGraphExecDestroy(g.second);
    graphs.clear();
    if (cap_stream) cudaStreamDestroy(cap_stream);
    cap_stream = nullptr;
  }
};

struct Pi0State {
  // packed / derived weights
  bf16* w_patch = nullptr;  // [Wv, kpad]
  int kpad = 0;
  std::vector<VisLayer> vis;
  std::vector<GemmaLayer> lm, ex;
  float* rope_timescale = nullptr;  // [hd/2]
  float2* rope_tab = nullptr;       // [max_rephrases][suffix_len][hd/2] (cos, sin) of the suffix positions
  float* time_vec = nullptr;        // [steps, We] = W_in[:, We:] . bf16(time_emb[s])
  float* time_emb_f32 = nullptr;    // [steps, We] (bf16-rounded values)
  float* w_ain_comb = nullptr;      // [We, max_action_dim] = W_in[:, :We] . W_action_in  (two fp32 linears with nothing between)
  float* b_ain_comb = nullptr;      // [We] = W_in[:, :We] . b_action_in + b_in
  std::vector<float> times;
  float dt = 0.f;
  // inputs (staged copies so the captured graph only touches internal memory)
  float* in_image = nullptr;
  int64_t* in_tokens = nullptr;
  int* in_lang_len = nullptr;
  float* in_state = nullptr;
  float* x_t = nullptr;
  int* plen = nullptr;
  // workspace
  bf16 *patches = nullptr, *hv = nullptr, *xv = nullptr, *qkv_v = nullptr, *attn_v = nullptr,
       *mlp_v = nullptr, *proj_out = nullptr;
  bf16* pos_tiled = nullptr;  // SigLIP position embedding repeated per observation (max_observations > 1)
  bf16 *hp = nullptr, *xp = nullptr, *qkv_p = nullptr, *attn_p = nullptr, *act_p = nullptr;
  bf16 *kcache = nullptr, *vcache = nullptr;
  bf16 *w_out3 = nullptr, *a2s = nullptr;  // action_time_mlp_out as [hi | lo | hi] bf16, its input as [hi | hi | lo]
  float *state_emb = nullptr, *a1 = nullptr, *a2 = nullptr, *suffix = nullptr, *v0 = nullptr;
  bf16 *he = nullptr, *xe = nullptr, *qkv_e = nullptr, *attn_e = nullptr, *act_e = nullptr;
  bf16 *state_k = nullptr, *state_v = nullptr;  // [layers][candidates][head_dim]: the suffix state token's rotated K / V (F7 hoist)
  bf16* vt_p = nullptr;  // V^T of the current prefix layer: [max_rephrases][head_dim][vt_ld] (tcgen05 prefix attention)
  long vt_ld = 0;
  float* part_e = nullptr;  // split-K partials of the expert's o_proj / down_proj: [kMaxSplitK][N*S][ex_width] fp32
  float* part_v = nullptr;  // split-K partials of the SigLIP tower's out_proj / fc2: [kMaxSplitK][n_img][vis_width] fp32
  int splitk_vo = 0, splitk_v2 = 0;
  int ex_gu_half = 64;  // gate/up packing of the expert: [half gate | half up] rows per block
  int splitk_o = 0, splitk_d = 0;  // K-splits of o_proj / down_proj in the denoise loop (0 = fused-epilogue GEMMs)
  int lang_hint = 0;  // caller's bound on valid language tokens per prompt (0 = max_lang_len), cvb_pi0_set_lang_len_hint
  GraphCache graphs;  // key = (lang rows << 40) | R << 16 | K
  ExpertMega mega;    // persistent expert kernel (engine_expert_mega.cu)
};

struct VerifierState;  // engine_verifier.cu

// fused decision (engine_cover.cu): one CUDA graph with a forked branch for the verifier's image/text side
struct CoverState {
  GraphCache graphs;
  cudaStream_t side = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  float* past = nullptr;  // staged copy of the caller's action-history tail [vf_history, 7]
  double stats[12] = {0};
  bool stats_set = false;
};

}  // namespace cvb

struct cvb_handle {
  cvb_config cfg;
  std::unordered_map<std::string, cvb::Weight> weights;
  std::vector<cvb::WeightSpec> required;
  std::vector<void*> owned;
  bool finalized = false;
  cvb::Pi0State pi0;
  cvb::VerifierState* vf = nullptr;
  cvb::CoverState cover;

  // cameras per observation (modeling_pi0.py:344-387, 529-547): the handle is sized for cfg.num_cameras image streams, a
  // call uses the first `active_cams` of them (cvb_pi0_set_active_cameras; empty / masked cameras are dropped by the host -
  // their tokens are masked as keys and never read, so dropping them is exact, like right-padded language tokens)
  int active_cams = 0;  // 0 = all of cfg.num_cameras
  int cams_max() const { return cfg.num_cameras > 1 ? cfg.num_cameras : 1; }
  int cams() const { return active_cams > 0 ? active_cams : cams_max(); }
  int n_img1() const { return (cfg.vis_image / cfg.vis_patch) * (cfg.vis_image / cfg.vis_patch); }  // tokens per image
  int n_img() const { return cams() * n_img1(); }          // image tokens of a prompt in THIS call
  int n_img_max() const { return cams_max() * n_img1(); }  // workspace / cache sizing
  // observations per batched call (cvb_pi0_sample_batch / cvb_cover_step_batch) and the global rephrase capacity
  int max_obs() const { return cfg.max_observations > 1 ? cfg.max_observations : 1; }
  int rm_total() const { return cfg.max_rephrases * max_obs(); }
  int prefix_len() const { return n_img_max() + cfg.max_lang_len; }  // KV-cache rows per prompt (a stride, not a length)
  int suffix_len() const { return 1 + cfg.chunk_size; }
  // language rows actually processed per prompt: the caller's hint rounded up to 8 (bounds the number of graphs)
  int lang_rows() const {
    const int hint = pi0.lang_hint > 0 ? pi0.lang_hint : cfg.max_lang_len;
    const int r = (hint + 7) / 8 * 8;
    return r < cfg.max_lang_len ? r : cfg.max_lang_len;
  }
};

namespace cvb {

std::string canonical_key(const std::string& k);
int dalloc(cvb_handle* h, void** p, size_t bytes);
template <typename T>
int dalloc_t(cvb_handle* h, T** p, size_t count) {
  return dalloc(h, reinterpret_cast<void**>(p), count * sizeof(T));
}
// fetch a bound weight, checking dtype and element count
int get_weight(cvb_handle* h, const std::string& key, int dtype, int64_t numel, const void** out);

void pi0_required_weights(const cvb_config& c, std::vector<WeightSpec>* out);
int pi0_finalize(cvb_handle* h, cudaStream_t st);
// B > 1: B observations per call; images [B,3,H,W], tokens [B*R,L], lang_len [B*R], state [B,max_state_dim], noise /
// actions [B*R*K, chunk, max_action_dim] (observation-major, then rephrase-major)
int pi0_sample(cvb_handle* h, const float* image, const int64_t* tokens, const int32_t* lang_len,
               const float* state, const float* noise, int R, int K, float* actions, cudaStream_t st, int B = 1);
int pi0_run_phase(cvb_handle* h, int phase, int R, int K, cudaStream_t st, int B = 1);
int64_t pi0_debug_copy(cvb_handle* h, const std::string& name, void* dst, int64_t max_bytes,
                       cudaStream_t st);

int expert_mega_prepare(cvb_handle* h, cudaStream_t st);
int expert_mega_program(cvb_handle* h, int rows, cudaStream_t st, const MegaProgram** out);
int expert_mega_launch(cvb_handle* h, cudaStream_t st, const MegaProgram& pg, int first, int count);

void verifier_required_weights(const cvb_config& c, std::vector<WeightSpec>* out);
int verifier_finalize(cvb_handle* h, cudaStream_t st);
void verifier_destroy(cvb_handle* h);
int verifier_score(cvb_handle* h, const float* image, const int64_t* tokens, const float* traj, int N, int R, int K,
                   float* scores, float* group_mean, int32_t* best_idx, float* best_score, int recompute_context,
                   cudaStream_t st);
int verifier_context(cvb_handle* h, const float* image, const int64_t* tokens, cudaStream_t st);
// building blocks of the fused decision (engine_cover.cu)
int pi0_stage_inputs(cvb_handle* h, const float* image, const int64_t* tokens, const int32_t* lang_len,
                     const float* state, const float* noise, int R, int K, cudaStream_t st, int B = 1);
int pi0_enqueue(cvb_handle* h, cudaStream_t st, int R, int K, int part, int B = 1);  // 0 = vision + prefix, 1 = denoise loop
float* pi0_actions_buffer(cvb_handle* h);
int verifier_stage_context_inputs(cvb_handle* h, const float* image, const int64_t* tokens, cudaStream_t st, int B = 1);
bool verifier_text_cached(cvb_handle* h, int nb);
void verifier_note_context(cvb_handle* h, int nb);
int verifier_hold_text(cvb_handle* h, int hold);
int verifier_enqueue_context(cvb_handle* h, cudaStream_t st, int obs = 0, int nb = 1);  // slots [obs, obs + nb)
int verifier_enqueue_score(cvb_handle* h, cudaStream_t st, int N, int R, int K, int B = 1);
float* verifier_traj_buffer(cvb_handle* h);
int verifier_copy_results(cvb_handle* h, int N, int R, float* scores, float* group_mean, int32_t* best_idx,
                          float* best_score, cudaStream_t st, int B = 1);
int cover_step(cvb_handle* h, const float* image, const int64_t* lang_tokens, const int32_t* lang_len,
               const float* state, const float* noise, int R, int K, const float* vf_image, const int64_t* vf_tokens,
               const double* p01_host, const double* p99_host, const float* past, int num_past, int n_future,
               float* actions, float* traj, float* scores, float* group_mean, int32_t* best_idx, float* best_score,
               cudaStream_t st, int B = 1);
void cover_destroy(cvb_handle* h);
int64_t verifier_debug_copy(cvb_handle* h, const std::string& name, void* dst, int64_t max_bytes, cudaStream_t st);
int verifier_set_features(cvb_handle* h, const float* patch, const float* text, cudaStream_t st);

}  // namespace cvb
