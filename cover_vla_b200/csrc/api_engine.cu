// C-ABI entry points of the engine (handle lifetime, weight binding, pi0 sampling, verifier scoring).
#include <cmath>
#include <cstring>

#include "engine.h"
#include "verifier_kernels.h"

namespace cvb {

static bool replace_first(std::string& s, const std::string& from, const std::string& to) {
  const size_t p = s.find(from);
  if (p == std::string::npos) return false;
  s.replace(p, from.size(), to);
  return true;
}

// Accept PI0Policy.state_dict() names ("model." prefix) and both transformers module layouts
// (SURVEY.md Appendix C): canonical = the 4.48.3 layout the published checkpoints use.
std::string canonical_key(const std::string& k_in) {
  std::string k = k_in;
  if (k.rfind("model.", 0) == 0) k = k.substr(6);
  replace_first(k, "paligemma.model.vision_tower.", "paligemma.vision_tower.");
  replace_first(k, "paligemma.model.multi_modal_projector.", "paligemma.multi_modal_projector.");
  replace_first(k, "paligemma.model.language_model.", "paligemma.language_model.model.");
  return k;
}

int dalloc(cvb_handle* h, void** p, size_t bytes) {
  if (bytes == 0) bytes = 16;
  CVB_CUDA(cudaMalloc(p, bytes));
  h->owned.push_back(*p);
  return 0;
}

int get_weight(cvb_handle* h, const std::string& key, int dtype, int64_t numel, const void** out) {
  auto it = h->weights.find(key);
  if (it == h->weights.end()) {
    set_last_error("weight not bound: " + key);
    return -3;
  }
  if (it->second.dtype != dtype) {
    set_last_error("weight has wrong dtype: " + key);
    return -3;
  }
  if (numel >= 0 && it->second.numel() != numel) {
    set_last_error("weight has wrong element count: " + key + " (got " +
                   std::to_string(it->second.numel()) + ", want " + std::to_string(numel) + ")");
    return -3;
  }
  *out = it->second.ptr;
  return 0;
}

}  // namespace cvb

extern "C" {

int cvb_create(const cvb_config* cfg, cvb_handle** out) {
  CVB_REQUIRE(cfg != nullptr && out != nullptr, "null argument");
  CVB_REQUIRE(cfg->struct_size == (int32_t)sizeof(cvb_config), "cvb_config size mismatch (ABI)");
  CVB_REQUIRE(cfg->max_rephrases >= 1 && cfg->max_samples >= 1, "max_rephrases / max_samples must be >= 1");
  cvb_handle* h = new cvb_handle();
  h->cfg = *cfg;
  if (cfg->layers > 0) cvb::pi0_required_weights(h->cfg, &h->required);
  if (cfg->vf_members > 0) cvb::verifier_required_weights(h->cfg, &h->required);
  *out = h;
  return 0;
}

void cvb_destroy(cvb_handle* h) {
  if (h == nullptr) return;
  h->pi0.graphs.destroy();
  cvb::cover_destroy(h);
  cvb::verifier_destroy(h);
  for (void* p : h->owned) cudaFree(p);
  delete h;
}

int cvb_bind_weight(cvb_handle* h, const char* key, const void* dev_ptr, int dtype, int ndim,
                    const int64_t* shape) {
  CVB_REQUIRE(h != nullptr && key != nullptr && dev_ptr != nullptr, "null argument");
  CVB_REQUIRE(!h->finalized, "handle already finalized");
  cvb::Weight w;
  w.ptr = dev_ptr;
  w.dtype = dtype;
  w.shape.assign(shape, shape + ndim);
  h->weights[cvb::canonical_key(key)] = std::move(w);
  return 0;
}

int cvb_required_weight_count(cvb_handle* h) { return h ? (int)h->required.size() : 0; }

const char* cvb_required_weight_name(cvb_handle* h, int index) {
  if (h == nullptr || index < 0 || index >= (int)h->required.size()) return nullptr;
  return h->required[index].key.c_str();
}

int cvb_required_weight_dtype(cvb_handle* h, int index) {
  if (h == nullptr || index < 0 || index >= (int)h->required.size()) return -1;
  return h->required[index].dtype;
}

int cvb_finalize(cvb_handle* h, void* stream) {
  CVB_REQUIRE(h != nullptr, "null handle");
  CVB_REQUIRE(!h->finalized, "handle already finalized");
  for (const auto& spec : h->required) {
    auto it = h->weights.find(spec.key);
    if (it == h->weights.end()) {
      cvb::set_last_error("missing weight: " + spec.key);
      return -3;
    }
    if (it->second.dtype != spec.dtype) {
      cvb::set_last_error("wrong dtype for weight: " + spec.key);
      return -3;
    }
    int64_t want = 1;
    for (auto d : spec.shape) want *= d;
    if (it->second.numel() != want) {
      cvb::set_last_error("wrong shape for weight: " + spec.key);
      return -3;
    }
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (h->cfg.layers > 0) CVB_TRY(cvb::pi0_finalize(h, st));
  if (h->cfg.vf_members > 0) CVB_TRY(cvb::verifier_finalize(h, st));
  CVB_CUDA(cudaStreamSynchronize(st));
  h->finalized = true;
  return 0;
}

int cvb_pi0_sample(cvb_handle* h, const float* image, const int64_t* lang_tokens,
                   const int32_t* lang_len, const float* state, const float* noise, int R, int K,
                   float* actions, void* stream) {
  CVB_REQUIRE(h != nullptr, "null handle");
  return cvb::pi0_sample(h, image, lang_tokens, lang_len, state, noise, R, K, actions,
                         (cudaStream_t)stream);
}

int cvb_pi0_set_lang_len_hint(cvb_handle* h, int max_valid_tokens) {
  CVB_REQUIRE(h != nullptr, "null handle");
  CVB_REQUIRE(max_valid_tokens >= 0, "hint must be >= 0 (0 = no hint)");
  h->pi0.lang_hint = max_valid_tokens;
  return 0;
}

int cvb_verifier_hold_text(cvb_handle* h, int hold) {
  CVB_REQUIRE(h != nullptr, "null handle");
  return cvb::verifier_hold_text(h, hold);
}

int cvb_pi0_set_active_cameras(cvb_handle* h, int cameras) {
  CVB_REQUIRE(h != nullptr, "null handle");
  CVB_REQUIRE(cameras >= 0 && cameras <= h->cams_max(), "cameras must be 0 (all) .. num_cameras");
  h->active_cams = cameras;
  return 0;
}

int cvb_pi0_run_phase(cvb_handle* h, int phase, int R, int K, void* stream) {
  CVB_REQUIRE(h != nullptr, "null handle");
  return cvb::pi0_run_phase(h, phase, R, K, (cudaStream_t)stream);
}

int64_t cvb_debug_copy(cvb_handle* h, const char* name, void* dst, int64_t max_bytes, void* stream) {
  if (h == nullptr || name == nullptr) return -1;
  if (std::string(name).rfind("vf_", 0) == 0)
    return cvb::verifier_debug_copy(h, name, dst, max_bytes, (cudaStream_t)stream);
  return cvb::pi0_debug_copy(h, name, dst, max_bytes, (cudaStream_t)stream);
}

int cvb_verifier_score(cvb_handle* h, const float* image, const int64_t* text_tokens, const float* traj, int N,
                       int R, int K, float* scores, float* group_mean, int32_t* best_idx, float* best_score,
                       int recompute_context, void* stream) {
  CVB_REQUIRE(h != nullptr, "null handle");
  return cvb::verifier_score(h, image, text_tokens, traj, N, R, K, scores, group_mean, best_idx, best_score,
                             recompute_context, (cudaStream_t)stream);
}

int cvb_verifier_context(cvb_handle* h, const float* image, const int64_t* text_tokens, void* stream) {
  CVB_REQUIRE(h != nullptr, "null handle");
  return cvb::verifier_context(h, image, text_tokens, (cudaStream_t)stream);
}

int cvb_verifier_set_features(cvb_handle* h, const float* patch, const float* text, void* stream) {
  CVB_REQUIRE(h != nullptr, "null handle");
  return cvb::verifier_set_features(h, patch, text, (cudaStream_t)stream);
}

int cvb_format_trajectories(const float* actions, int n_cand, int chunk, int action_stride,
                            const double* p01_host, const double* p99_host, const float* past, int num_past,
                            int history, int n_future, float* traj, void* stream) {
  CVB_REQUIRE(actions != nullptr && traj != nullptr && p01_host != nullptr && p99_host != nullptr, "null argument");
  CVB_REQUIRE(num_past == 0 || past != nullptr, "past actions pointer required when num_past > 0");
  cvb::FormatStats st;
  for (int i = 0; i < 6; ++i) st.p01[i] = p01_host[i], st.p99[i] = p99_host[i];
  return cvb::format_trajectories((cudaStream_t)stream, actions, n_cand, chunk, action_stride, st, past, num_past,
                                  history, n_future, traj);
}

int cvb_execution_action(const float* actions, int n_cand, int chunk, int action_stride, const double* p01_host,
                         const double* p99_host, const int32_t* best_idx, int K, int step, double* exec_action,
                         int32_t* votes, void* stream) {
  CVB_REQUIRE(actions != nullptr && exec_action != nullptr && best_idx != nullptr && p01_host != nullptr &&
                  p99_host != nullptr, "null argument");
  cvb::FormatStats st;
  for (int i = 0; i < 6; ++i) st.p01[i] = p01_host[i], st.p99[i] = p99_host[i];
  return cvb::execution_action((cudaStream_t)stream, actions, n_cand, chunk, action_stride, st, best_idx, K, step,
                               exec_action, votes);
}

int cvb_cover_step(cvb_handle* h, const float* image, const int64_t* lang_tokens, const int32_t* lang_len,
                   const float* state, const float* noise, int R, int K, const float* vf_image,
                   const int64_t* vf_text_tokens, const double* p01_host, const double* p99_host, const float* past,
                   int num_past, int n_future, float* actions, float* traj, float* scores, float* group_mean,
                   int32_t* best_idx, float* best_score, void* stream) {
  CVB_REQUIRE(h != nullptr, "null handle");
  return cvb::cover_step(h, image, lang_tokens, lang_len, state, noise, R, K, vf_image, vf_text_tokens, p01_host,
                         p99_host, past, num_past, n_future, actions, traj, scores, group_mean, best_idx, best_score,
                         (cudaStream_t)stream);
}

int cvb_cover_step_batch(cvb_handle* h, int B, const float* images, const int64_t* lang_tokens, const int32_t* lang_len,
                         const float* states, const float* noise, int R, int K, const float* vf_images,
                         const int64_t* vf_text_tokens, const double* p01_host, const double* p99_host, const float* past,
                         int num_past, int n_future, float* actions, float* traj, float* scores, float* group_mean,
                         int32_t* best_idx, float* best_score, void* stream) {
  CVB_REQUIRE(h != nullptr, "null handle");
  return cvb::cover_step(h, images, lang_tokens, lang_len, states, noise, R, K, vf_images, vf_text_tokens, p01_host,
                         p99_host, past, num_past, n_future, actions, traj, scores, group_mean, best_idx, best_score,
                         (cudaStream_t)stream, B);
}

int cvb_pi0_sample_batch(cvb_handle* h, int B, const float* images, const int64_t* lang_tokens, const int32_t* lang_len,
                         const float* states, const float* noise, int R, int K, float* actions, void* stream) {
  CVB_REQUIRE(h != nullptr, "null handle");
  return cvb::pi0_sample(h, images, lang_tokens, lang_len, states, noise, R, K, actions, (cudaStream_t)stream, B);
}

int cvb_pi0_run_phase_batch(cvb_handle* h, int phase, int B, int R, int K, void* stream) {
  CVB_REQUIRE(h != nullptr, "null handle");
  return cvb::pi0_run_phase(h, phase, R, K, (cudaStream_t)stream, B);
}

int cvb_select(const float* scores, int R, int K, float* group_mean, int32_t* best_idx, float* best_score,
               void* stream) {
  CVB_REQUIRE(scores != nullptr && best_idx != nullptr && best_score != nullptr, "null argument");
  CVB_REQUIRE(R >= 1 && K >= 1, "R and K must be >= 1");
  return cvb::select_best((cudaStream_t)stream, scores, R, K, group_mean, best_idx, best_score);
}

// ---- host-only constants of the denoise loop (modeling_pi0.py:697-714 and :71-89)
int cvb_denoise_times_host(int num_steps, float* times_out, int max_out, float* dt_out) {
  if (num_steps <= 0) return 0;
  const float dt = static_cast<float>(-1.0 / num_steps);
  float t = 1.0f;
  int n = 0;
  while (t >= -dt / 2) {
    if (times_out != nullptr && n < max_out) times_out[n] = t;
    ++n;
    t = t + dt;
    if (n > 100000) break;
  }
  if (dt_out != nullptr) *dt_out = dt;
  return n;
}

static uint16_t f32_to_bf16_rne(float f) {
  uint32_t x;
  memcpy(&x, &f, 4);
  if ((x & 0x7fffffffu) > 0x7f800000u) return static_cast<uint16_t>((x >> 16) | 0x40);
  const uint32_t lsb = (x >> 16) & 1u;
  x += 0x7fffu + lsb;
  return static_cast<uint16_t>(x >> 16);
}

void cvb_time_embedding_host(float t, int dim, double min_period, double max_period,
                             uint16_t* out_bf16) {
  const int half = dim / 2;
  const double step = half > 1 ? (1.0 - 0.0) / static_cast<double>(half - 1) : 0.0;
  for (int i = 0; i < half; ++i) {
    // torch.linspace(float64): start + step*i in the first half, end - step*(n-1-i) in the second
    const double fraction = (i < half / 2) ? 0.0 + step * i : 1.0 - step * (half - 1 - i);
    const double period = min_period * std::pow(max_period / min_period, fraction);
    const double scaling = 1.0 / period * 2 * M_PI;
    const double x = scaling * static_cast<double>(t);
    out_bf16[i] = f32_to_bf16_rne(static_cast<float>(std::sin(x)));
    out_bf16[half + i] = f32_to_bf16_rne(static_cast<float>(std::cos(x)));
  }
}

}  // extern "C"
