// pi0 sampling pipeline: SigLIP tower (once per observation) -> PaliGemma prefix (once per unique
// rephrase) -> 10-step action-expert denoise loop over all N = R*K candidates.  Host code only
// sequences kernels on one stream; the whole sequence is captured into a CUDA graph per (R, K).
//
// Reference: PI0FlowMatching.sample_actions, modeling_pi0.py:672-715 and everything it calls
// (embed_prefix :517-567, PaliGemmaWithExpertModel.forward paligemma_with_expert.py:236-360,
// embed_suffix :569-629, denoise_step :717-752).  De-duplication per SURVEY.md F1/F2.
#include <algorithm>
#include <cmath>
#include <cstdlib>

#include "engine.h"
#include "gemm_tcgen05.cuh"
#include "pi0_kernels.h"

namespace cvb {

bool attention_decode_umma_eligible(const AttnCall& c);

namespace {

const std::string PW = "paligemma_with_expert.";
const std::string VT = PW + "paligemma.vision_tower.vision_model.";
const std::string MM = PW + "paligemma.multi_modal_projector.linear.";
const std::string LM = PW + "paligemma.language_model.model.";
const std::string EX = PW + "gemma_expert.model.";

int round_up(int x, int m) { return (x + m - 1) / m * m; }
constexpr int kMaxSplitK = 16;

template <typename T>
int W(cvb_handle* h, const std::string& key, int dtype, int64_t numel, const T** out) {
  const void* p = nullptr;
  CVB_TRY(get_weight(h, key, dtype, numel, &p));
  *out = reinterpret_cast<const T*>(p);
  return 0;
}

// concat row-major matrices along rows into an owned buffer
int concat_rows(cvb_handle* h, cudaStream_t st, std::vector<std::pair<const bf16*, int64_t>> parts,
                int64_t cols, bf16** out) {
  int64_t rows = 0;
  for (auto& p : parts) rows += p.second;
  CVB_TRY(dalloc_t(h, out, rows * cols));
  int64_t off = 0;
  for (auto& p : parts) {
    CVB_CUDA(cudaMemcpyAsync(*out + off * cols, p.first, p.second * cols * sizeof(bf16),
                             cudaMemcpyDeviceToDevice, st));
    off += p.second;
  }
  return 0;
}

// gate/up -> [128 gate rows | 128 up rows] per 128-feature block (zero padded)
// (half = 128: prefix, 256-wide GeGLU tiles; half = 64: expert, 128-wide tiles)
int pack_gate_up(cvb_handle* h, cudaStream_t st, const bf16* wg, const bf16* wu, int I, int D,
                 bf16** out, int half = 128) {
  const int blocks = (I + half - 1) / half;
  CVB_TRY(dalloc_t(h, out, static_cast<size_t>(blocks) * 2 * half * D));
  CVB_CUDA(cudaMemsetAsync(*out, 0, static_cast<size_t>(blocks) * 2 * half * D * sizeof(bf16), st));
  for (int b = 0; b < blocks; ++b) {
    const int rows = std::min(half, I - b * half);
    CVB_CUDA(cudaMemcpyAsync(*out + (static_cast<size_t>(b) * 2 * half) * D, wg + static_cast<size_t>(b) * half * D,
                             static_cast<size_t>(rows) * D * sizeof(bf16), cudaMemcpyDeviceToDevice, st));
    CVB_CUDA(cudaMemcpyAsync(*out + (static_cast<size_t>(b) * 2 * half + half) * D,
                             wu + static_cast<size_t>(b) * half * D,
                             static_cast<size_t>(rows) * D * sizeof(bf16), cudaMemcpyDeviceToDevice, st));
  }
  return 0;
}

// set by run_vision / run_prefix / run_denoise: a handle built for several observations must give a row the same bits
// whatever the number of observations in the call, so its GEMMs never take the row-count dependent skinny kernel
thread_local bool tl_batch_handle = false;

int gemm(cudaStream_t st, const bf16* A, long lda, const bf16* Wt, long ldw, int M, int N, int K,
         int epi, void* C, long ldc, const void* bias = nullptr, const void* resid = nullptr,
         long ldr = 0, int resid_f32 = 0, int n_out = 0, int force_bn = 0) {
  GemmCall c;
  c.force_bn = force_bn;
  c.no_skinny = tl_batch_handle ? 1 : 0;
  c.A = A, c.lda = lda, c.W = Wt, c.ldw = ldw, c.M = M, c.N = N, c.K = K, c.epi = epi;
  c.C = C, c.ldc = ldc, c.bias = bias, c.bias_is_f32 = 0, c.resid = resid, c.ldr = ldr;
  c.resid_is_f32 = resid_f32, c.n_out = n_out;
  return gemm_bf16(st, c);
}

}  // namespace

// --------------------------------------------------------------------------------------------------
void pi0_required_weights(const cvb_config& c, std::vector<WeightSpec>* out) {
  auto add = [&](const std::string& k, int dt, std::vector<int64_t> shape) {
    out->push_back(WeightSpec{k, dt, std::move(shape)});
  };
  const int n_img = (c.vis_image / c.vis_patch) * (c.vis_image / c.vis_patch);
  add(VT + "embeddings.patch_embedding.weight", CVB_BF16, {c.vis_width, 3, c.vis_patch, c.vis_patch});
  add(VT + "embeddings.patch_embedding.bias", CVB_BF16, {c.vis_width});
  add(VT + "embeddings.position_embedding.weight", CVB_BF16, {n_img, c.vis_width});
  for (int l = 0; l < c.vis_layers; ++l) {
    const std::string p = VT + "encoder.layers." + std::to_string(l) + ".";
    for (const char* ln : {"layer_norm1", "layer_norm2"}) {
      add(p + ln + ".weight", CVB_BF16, {c.vis_width});
      add(p + ln + ".bias", CVB_BF16, {c.vis_width});
    }
    for (const char* nm : {"q_proj", "k_proj", "v_proj", "out_proj"}) {
      add(p + "self_attn." + nm + ".weight", CVB_BF16, {c.vis_width, c.vis_width});
      add(p + "self_attn." + nm + ".bias", CVB_BF16, {c.vis_width});
    }
    add(p + "mlp.fc1.weight", CVB_BF16, {c.vis_mlp, c.vis_width});
    add(p + "mlp.fc1.bias", CVB_BF16, {c.vis_mlp});
    add(p + "mlp.fc2.weight", CVB_BF16, {c.vis_width, c.vis_mlp});
    add(p + "mlp.fc2.bias", CVB_BF16, {c.vis_width});
  }
  add(VT + "post_layernorm.weight", CVB_BF16, {c.vis_width});
  add(VT + "post_layernorm.bias", CVB_BF16, {c.vis_width});
  add(MM + "weight", CVB_BF16, {c.lm_width, c.vis_width});
  add(MM + "bias", CVB_BF16, {c.lm_width});
  add(LM + "embed_tokens.weight", CVB_BF16, {c.vocab, c.lm_width});
  const int qd = c.heads * c.head_dim;
  for (int t = 0; t < 2; ++t) {
    const std::string& base = t == 0 ? LM : EX;
    const int D = t == 0 ? c.lm_width : c.ex_width, I = t == 0 ? c.lm_mlp : c.ex_mlp;
    for (int l = 0; l < c.layers; ++l) {
      const std::string p = base + "layers." + std::to_string(l) + ".";
      add(p + "self_attn.q_proj.weight", CVB_BF16, {qd, D});
      add(p + "self_attn.k_proj.weight", CVB_BF16, {c.head_dim, D});
      add(p + "self_attn.v_proj.weight", CVB_BF16, {c.head_dim, D});
      add(p + "self_attn.o_proj.weight", CVB_BF16, {D, qd});
      add(p + "mlp.gate_proj.weight", CVB_BF16, {I, D});
      add(p + "mlp.up_proj.weight", CVB_BF16, {I, D});
      add(p + "mlp.down_proj.weight", CVB_BF16, {D, I});
      add(p + "input_layernorm.weight", CVB_BF16, {D});
      add(p + "post_attention_layernorm.weight", CVB_BF16, {D});
    }
  }
  add(EX + "norm.weight", CVB_F32, {c.ex_width});
  add("state_proj.weight", CVB_F32, {c.ex_width, c.max_state_dim});
  add("state_proj.bias", CVB_F32, {c.ex_width});
  add("action_in_proj.weight", CVB_F32, {c.ex_width, c.max_action_dim});
  add("action_in_proj.bias", CVB_F32, {c.ex_width});
  add("action_out_proj.weight", CVB_F32, {c.max_action_dim, c.ex_width});
  add("action_out_proj.bias", CVB_F32, {c.max_action_dim});
  add("action_time_mlp_in.weight", CVB_F32, {c.ex_width, 2 * c.ex_width});
  add("action_time_mlp_in.bias", CVB_F32, {c.ex_width});
  add("action_time_mlp_out.weight", CVB_F32, {c.ex_width, c.ex_width});
  add("action_time_mlp_out.bias", CVB_F32, {c.ex_width});
}

// --------------------------------------------------------------------------------------------------
int pi0_finalize(cvb_handle* h, cudaStream_t st) {
  const cvb_config& c = h->cfg;
  Pi0State& s = h->pi0;
  CVB_REQUIRE(c.vis_width % 8 == 0 && c.vis_mlp % 8 == 0 && c.lm_width % 8 == 0 && c.ex_width % 8 == 0 &&
                  c.lm_mlp % 8 == 0 && c.ex_mlp % 8 == 0,
              "layer widths must be multiples of 8");
  CVB_REQUIRE(c.vis_width % c.vis_heads == 0 && (c.vis_width / c.vis_heads) % 8 == 0,
              "vision head_dim must be a multiple of 8");
  CVB_REQUIRE(c.head_dim % 16 == 0 && c.head_dim <= 256, "head_dim must be a multiple of 16, <= 256");
  CVB_REQUIRE(c.vis_image % c.vis_patch == 0, "image size must be a multiple of the patch size");
  const int T = h->n_img_max(), T1 = h->n_img1(), P = h->prefix_len(), S = h->suffix_len();
  const int Wv = c.vis_width, D = c.lm_width, We = c.ex_width;
  CVB_REQUIRE(c.num_cameras >= 0 && c.num_cameras <= 8, "num_cameras must be 0..8");
  const int qd = c.heads * c.head_dim, qkvw = qd + 2 * c.head_dim;
  // Bm observations per call (cvb_*_batch): every "rephrase" index below is global,
  // r = observation * R + rephrase; the image / state of rephrase r are those of observation r / R
  const int Bm = h->max_obs();
  const int Rm = h->rm_total(), Nm = Rm * c.max_samples;

  // ---- patch embedding weight, K padded to a multiple of 8 (TMA rows are 16-byte multiples)
  const int kreal = 3 * c.vis_patch * c.vis_patch;
  s.kpad = round_up(kreal, 8);
  const bf16* wp = nullptr;
  CVB_TRY(W(h, VT + "embeddings.patch_embedding.weight", CVB_BF16, (int64_t)Wv * kreal, &wp));
  CVB_TRY(dalloc_t(h, &s.w_patch, (size_t)Wv * s.kpad));
  CVB_CUDA(cudaMemsetAsync(s.w_patch, 0, (size_t)Wv * s.kpad * sizeof(bf16), st));
  CVB_CUDA(cudaMemcpy2DAsync(s.w_patch, s.kpad * sizeof(bf16), wp, kreal * sizeof(bf16),
                             kreal * sizeof(bf16), Wv, cudaMemcpyDeviceToDevice, st));

  // ---- vision layers
  s.vis.resize(c.vis_layers);
  for (int l = 0; l < c.vis_layers; ++l) {
    const std::string p = VT + "encoder.layers." + std::to_string(l) + ".";
    VisLayer& L = s.vis[l];
    CVB_TRY(W(h, p + "layer_norm1.weight", CVB_BF16, Wv, &L.ln1_w));
    CVB_TRY(W(h, p + "layer_norm1.bias", CVB_BF16, Wv, &L.ln1_b));
    CVB_TRY(W(h, p + "layer_norm2.weight", CVB_BF16, Wv, &L.ln2_w));
    CVB_TRY(W(h, p + "layer_norm2.bias", CVB_BF16, Wv, &L.ln2_b));
    const bf16 *wq, *wk, *wv, *bq, *bk, *bv;
    CVB_TRY(W(h, p + "self_attn.q_proj.weight", CVB_BF16, (int64_t)Wv * Wv, &wq));
    CVB_TRY(W(h, p + "self_attn.k_proj.weight", CVB_BF16, (int64_t)Wv * Wv, &wk));
    CVB_TRY(W(h, p + "self_attn.v_proj.weight", CVB_BF16, (int64_t)Wv * Wv, &wv));
    CVB_TRY(W(h, p + "self_attn.q_proj.bias", CVB_BF16, Wv, &bq));
    CVB_TRY(W(h, p + "self_attn.k_proj.bias", CVB_BF16, Wv, &bk));
    CVB_TRY(W(h, p + "self_attn.v_proj.bias", CVB_BF16, Wv, &bv));
    CVB_TRY(concat_rows(h, st, {{wq, Wv}, {wk, Wv}, {wv, Wv}}, Wv, &L.wqkv));
    CVB_TRY(concat_rows(h, st, {{bq, 1}, {bk, 1}, {bv, 1}}, Wv, &L.bqkv));
    CVB_TRY(W(h, p + "self_attn.out_proj.weight", CVB_BF16, (int64_t)Wv * Wv, &L.wo));
    CVB_TRY(W(h, p + "self_attn.out_proj.bias", CVB_BF16, Wv, &L.bo));
    CVB_TRY(W(h, p + "mlp.fc1.weight", CVB_BF16, (int64_t)c.vis_mlp * Wv, &L.w1));
    CVB_TRY(W(h, p + "mlp.fc1.bias", CVB_BF16, c.vis_mlp, &L.b1));
    CVB_TRY(W(h, p + "mlp.fc2.weight", CVB_BF16, (int64_t)c.vis_mlp * Wv, &L.w2));
    CVB_TRY(W(h, p + "mlp.fc2.bias", CVB_BF16, Wv, &L.b2));
  }

  // ---- Gemma towers
  for (int t = 0; t < 2; ++t) {
    const std::string& base = t == 0 ? LM : EX;
    const int Dt = t == 0 ? D : We, I = t == 0 ? c.lm_mlp : c.ex_mlp;
    std::vector<GemmaLayer>& layers = t == 0 ? s.lm : s.ex;
    layers.resize(c.layers);
    for (int l = 0; l < c.layers; ++l) {
      const std::string p = base + "layers." + std::to_string(l) + ".";
      GemmaLayer& L = layers[l];
      const bf16 *wq, *wk, *wv, *wg, *wu;
      CVB_TRY(W(h, p + "self_attn.q_proj.weight", CVB_BF16, (int64_t)qd * Dt, &wq));
      CVB_TRY(W(h, p + "self_attn.k_proj.weight", CVB_BF16, (int64_t)c.head_dim * Dt, &wk));
      CVB_TRY(W(h, p + "self_attn.v_proj.weight", CVB_BF16, (int64_t)c.head_dim * Dt, &wv));
      CVB_TRY(concat_rows(h, st, {{wq, qd}, {wk, c.head_dim}, {wv, c.head_dim}}, Dt, &L.wqkv));
      CVB_TRY(W(h, p + "self_attn.o_proj.weight", CVB_BF16, (int64_t)qd * Dt, &L.wo));
      CVB_TRY(W(h, p + "mlp.gate_proj.weight", CVB_BF16, (int64_t)I * Dt, &wg));
      CVB_TRY(W(h, p + "mlp.up_proj.weight", CVB_BF16, (int64_t)I * Dt, &wu));
      // expert: [64 gate | 64 up] blocks for 128-wide GeGLU tiles on latency handles (M = 160-200 rows: 128 CTAs instead of
      // 64).  A handle built for >= 4 observations keeps the 256-wide packing: at M >= 640 rows the 128-wide tiles make 5
      // rounds of 640 tiles re-reading the activations 64 x, the CTA-pair kernel 3 rounds of 160 (batched denoise 31.5 ->
      // 28.0 ms at 8 observations, 24.4 -> 21.0 ms at 4; tools/batch_time.py).  A handle-level choice: a row's bits do not
      // depend on how many observations share a call.  CVB_EXPERT_GU256=1 / =0 force either.
      const char* gu_env = getenv("CVB_EXPERT_GU256");
      s.ex_gu_half = (gu_env != nullptr ? atoi(gu_env) != 0 : h->max_obs() >= 4) ? 128 : 64;
      CVB_TRY(pack_gate_up(h, st, wg, wu, I, Dt, &L.wgu, t == 0 ? 128 : s.ex_gu_half));
      CVB_TRY(W(h, p + "mlp.down_proj.weight", CVB_BF16, (int64_t)I * Dt, &L.wd));
      CVB_TRY(W(h, p + "input_layernorm.weight", CVB_BF16, Dt, &L.in_norm));
      CVB_TRY(W(h, p + "post_attention_layernorm.weight", CVB_BF16, Dt, &L.post_norm));
    }
  }

  // ---- constants: RoPE timescale (paligemma_with_expert.py:43-44), denoise times, time embeddings
  {
    const int half = c.head_dim / 2;
    std::vector<float> ts(half);
    const float coef = static_cast<float>(2.0 / c.head_dim);
    for (int i = 0; i < half; ++i) ts[i] = powf(10000.0f, coef * static_cast<float>(i));
    CVB_TRY(dalloc_t(h, &s.rope_timescale, half));
    CVB_CUDA(cudaMemcpyAsync(s.rope_timescale, ts.data(), half * sizeof(float), cudaMemcpyHostToDevice, st));
    CVB_CUDA(cudaStreamSynchronize(st));  // ts is a stack-lifetime host buffer
  }
  {
    s.times.resize(256);
    int n = cvb_denoise_times_host(c.num_steps, s.times.data(), 256, &s.dt);
    CVB_REQUIRE(n > 0, "denoise time schedule is empty");
    s.times.resize(n);
    std::vector<float> temb((size_t)n * We);
    std::vector<uint16_t> row(We);
    for (int i = 0; i < n; ++i) {
      cvb_time_embedding_host(s.times[i], We, 4e-3, 4.0, row.data());
      for (int j = 0; j < We; ++j) {
        uint32_t bits = static_cast<uint32_t>(row[j]) << 16;
        float f;
        memcpy(&f, &bits, 4);
        temb[(size_t)i * We + j] = f;
      }
    }
    CVB_TRY(dalloc_t(h, &s.time_emb_f32, (size_t)n * We));
    CVB_TRY(dalloc_t(h, &s.time_vec, (size_t)n * We));
    CVB_CUDA(cudaMemcpyAsync(s.time_emb_f32, temb.data(), temb.size() * sizeof(float),
                             cudaMemcpyHostToDevice, st));
    CVB_CUDA(cudaStreamSynchronize(st));
    // time_vec[s] = W_in[:, We:2We] . time_emb[s]   (the time half of action_time_mlp_in, constant per step)
    const float* w_in = nullptr;
    CVB_TRY(W(h, "action_time_mlp_in.weight", CVB_F32, (int64_t)We * 2 * We, &w_in));
    SgemmCall g;
    g.A = s.time_emb_f32, g.lda = We, g.W = w_in + We, g.ldw = 2 * We, g.M = n, g.N = We, g.K = We;
    g.C = s.time_vec, g.ldc = We;
    CVB_TRY(sgemm_f32(st, g));
    // action_in_proj and the action half of action_time_mlp_in are two fp32 linear maps with nothing in between
    // (modeling_pi0.py:598-606): compose them once (fp32) - the per-step K = 1024 linear becomes a K = 32 one.
    // CVB_NO_FOLD_AIN=1 keeps the two-step evaluation.
    if (getenv("CVB_NO_FOLD_AIN") == nullptr) {
      const float *w_ain, *b_ain, *b_in;
      CVB_TRY(W(h, "action_in_proj.weight", CVB_F32, (int64_t)We * c.max_action_dim, &w_ain));
      CVB_TRY(W(h, "action_in_proj.bias", CVB_F32, We, &b_ain));
      CVB_TRY(W(h, "action_time_mlp_in.bias", CVB_F32, We, &b_in));
      CVB_TRY(dalloc_t(h, &s.w_ain_comb, (size_t)We * c.max_action_dim));
      CVB_TRY(dalloc_t(h, &s.b_ain_comb, We));
      SgemmCall gc;
      gc.A = w_in, gc.lda = 2 * We, gc.W = w_ain, gc.ldw = c.max_action_dim, gc.w_kn = 1;
      gc.M = We, gc.N = c.max_action_dim, gc.K = We, gc.C = s.w_ain_comb, gc.ldc = c.max_action_dim;
      CVB_TRY(sgemm_f32(st, gc));
      SgemmCall gb;
      gb.A = b_ain, gb.lda = We, gb.W = w_in, gb.ldw = 2 * We, gb.M = 1, gb.N = We, gb.K = We;
      gb.C = s.b_ain_comb, gb.ldc = We, gb.bias = b_in;
      CVB_TRY(sgemm_f32(st, gb));
    }
  }

  // ---- workspace
  CVB_TRY(dalloc_t(h, &s.in_image, (size_t)Bm * h->cams_max() * 3 * c.vis_image * c.vis_image));
  CVB_TRY(dalloc_t(h, &s.in_tokens, (size_t)Rm * c.max_lang_len));
  CVB_TRY(dalloc_t(h, &s.in_lang_len, Rm));
  CVB_TRY(dalloc_t(h, &s.plen, Rm));
  CVB_TRY(dalloc_t(h, &s.rope_tab, (size_t)Rm * S * (c.head_dim / 2)));
  CVB_TRY(dalloc_t(h, &s.in_state, (size_t)Bm * c.max_state_dim));
  CVB_TRY(dalloc_t(h, &s.x_t, (size_t)Nm * c.chunk_size * c.max_action_dim));
  CVB_TRY(dalloc_t(h, &s.v0, (size_t)Nm * c.chunk_size * c.max_action_dim));
  const size_t Tb = (size_t)Bm * T;
  CVB_TRY(dalloc_t(h, &s.patches, Tb * s.kpad));
  CVB_TRY(dalloc_t(h, &s.hv, Tb * Wv));
  CVB_TRY(dalloc_t(h, &s.xv, Tb * Wv));
  CVB_TRY(dalloc_t(h, &s.qkv_v, Tb * 3 * Wv));
  CVB_TRY(dalloc_t(h, &s.attn_v, Tb * Wv));
  CVB_TRY(dalloc_t(h, &s.mlp_v, Tb * c.vis_mlp));
  CVB_TRY(dalloc_t(h, &s.proj_out, Tb * D));
  if (Bm * h->cams_max() > 1) {  // position embedding tiled per image (the residual operand of the patch-embedding GEMM)
    const bf16* pos = nullptr;
    CVB_TRY(W(h, VT + "embeddings.position_embedding.weight", CVB_BF16, (int64_t)T1 * Wv, &pos));
    CVB_TRY(dalloc_t(h, &s.pos_tiled, Tb * Wv));
    for (int b = 0; b < Bm * h->cams_max(); ++b)
      CVB_CUDA(cudaMemcpyAsync(s.pos_tiled + (size_t)b * T1 * Wv, pos, (size_t)T1 * Wv * sizeof(bf16), cudaMemcpyDeviceToDevice, st));
  }
  const size_t Mp = (size_t)Rm * P;
  CVB_TRY(dalloc_t(h, &s.hp, Mp * D));
  CVB_TRY(dalloc_t(h, &s.xp, Mp * D));
  CVB_TRY(dalloc_t(h, &s.qkv_p, Mp * qkvw));
  CVB_TRY(dalloc_t(h, &s.attn_p, Mp * qd));
  CVB_TRY(dalloc_t(h, &s.act_p, Mp * c.lm_mlp));
  CVB_TRY(dalloc_t(h, &s.kcache, (size_t)c.layers * Mp * c.head_dim));
  CVB_TRY(dalloc_t(h, &s.vcache, (size_t)c.layers * Mp * c.head_dim));
  s.vt_ld = round_up(P, 64);
  // transposed V cache [layer][rephrase][head_dim][vt_ld]: the K-major "B" operand of the tcgen05 P.V GEMMs
  CVB_TRY(dalloc_t(h, &s.vt_p, (size_t)c.layers * Rm * c.head_dim * s.vt_ld));
  CVB_CUDA(cudaMemsetAsync(s.vt_p, 0, (size_t)c.layers * Rm * c.head_dim * s.vt_ld * sizeof(bf16), st));  // padding keys stay finite
  const size_t Me = (size_t)Nm * S, Ma = (size_t)Nm * c.chunk_size;
  CVB_TRY(dalloc_t(h, &s.state_emb, (size_t)Bm * We));
  // F7 hoist: rotated key / value of the state token of every candidate and expert layer, written by denoise step 0
  CVB_TRY(dalloc_t(h, &s.state_k, (size_t)c.layers * Nm * c.head_dim));
  CVB_TRY(dalloc_t(h, &s.state_v, (size_t)c.layers * Nm * c.head_dim));
  CVB_TRY(dalloc_t(h, &s.a1, Ma * We));
  CVB_TRY(dalloc_t(h, &s.a2, Ma * We));
  // action_time_mlp_out (fp32 in the reference, modeling_pi0.py:607-609) on the tensor cores with fp32 accuracy: 3-term
  // bf16 split (ops_misc.cu split3_rows), [hi | lo | hi] weights once, [hi | hi | lo] activations per step; the fp32
  // SIMT GEMM it replaces took 36 us of every Euler step.  CVB_SUFFIX_SIMT=1 keeps the SIMT kernel.
  if (getenv("CVB_SUFFIX_SIMT") == nullptr && We % 8 == 0) {
    const float* w_out_f32 = nullptr;
    CVB_TRY(W(h, "action_time_mlp_out.weight", CVB_F32, (int64_t)We * We, &w_out_f32));
    CVB_TRY(dalloc_t(h, &s.w_out3, (size_t)We * 3 * We));
    CVB_TRY(split3_rows(st, w_out_f32, We, s.w_out3, We, We, 1));
    CVB_TRY(dalloc_t(h, &s.a2s, std::min<size_t>(Ma, 256) * 3 * We));
  }
  CVB_TRY(dalloc_t(h, &s.suffix, Me * We));
  CVB_TRY(dalloc_t(h, &s.he, Me * We));
  CVB_TRY(dalloc_t(h, &s.xe, Me * We));
  CVB_TRY(dalloc_t(h, &s.qkv_e, Me * qkvw));
  CVB_TRY(dalloc_t(h, &s.attn_e, Me * qd));
  CVB_TRY(dalloc_t(h, &s.act_e, Me * c.ex_mlp));
  // Denoise-loop o_proj / down_proj as split-K partials reduced inside the following RMSNorm (ops_misc.cu
  // rmsnorm_reduce_kernel).  The rows of one denoise step must fit one UMMA N (<= 256); larger candidate counts keep
  // the fused-epilogue GEMMs.  CVB_SPLITK_O / CVB_SPLITK_D override the split counts (0 disables).
  if (T <= 256) {  // SigLIP tower: out_proj / fc2 the same way (LayerNorm variant); CVB_SPLITK_VO / CVB_SPLITK_V2 override
    auto pickv = [&](const char* env, int kdim, int dflt) {
      const int kb = (kdim + 63) / 64;
      int sp = std::min(dflt, std::max(1, kb / 2));
      if (const char* e = getenv(env)) sp = std::max(0, std::min(atoi(e), std::min(kMaxSplitK, kb)));
      return sp;
    };
    s.splitk_vo = pickv("CVB_SPLITK_VO", Wv, 6);
    s.splitk_v2 = pickv("CVB_SPLITK_V2", c.vis_mlp, 8);
    CVB_TRY(dalloc_t(h, &s.part_v, (size_t)kMaxSplitK * T * Wv));
  }
  {  // used whenever a call's suffix rows fit one UMMA N (<= 256); larger batches take the fused-epilogue GEMMs
    const size_t Me_sk = std::min<size_t>(Me, 256);
    auto pick = [&](const char* env, int kdim) {
      const int kb = (kdim + 63) / 64, tiles = (We + 127) / 128;
      // measured (tools/splitk_bench.py): 8 splits beat 4 / 12 / 16 at both shapes (more splits = more partial traffic)
      int sp = std::min(std::min(8, std::max(1, device_sm_count() / tiles)), std::max(1, kb / 2));
      if (const char* e = getenv(env)) sp = std::max(0, std::min(atoi(e), std::min(kMaxSplitK, kb)));
      return sp;
    };
    s.splitk_o = pick("CVB_SPLITK_O", qd);
    s.splitk_d = pick("CVB_SPLITK_D", c.ex_mlp);
    CVB_TRY(dalloc_t(h, &s.part_e, (size_t)kMaxSplitK * Me_sk * We));
  }
  CVB_TRY(expert_mega_prepare(h, st));  // persistent expert kernel (engine_expert_mega.cu), when the shape allows it
  return 0;
}

// --------------------------------------------------------------------------------------------------
static int run_vision(cvb_handle* h, cudaStream_t st, int n_obs) {
  const cvb_config& c = h->cfg;
  Pi0State& s = h->pi0;
  tl_batch_handle = h->max_obs() > 1;
  // every camera of every observation is one image of the tower's batch: rows [observation][camera][token], which is
  // also the order of the image tokens in a prompt (modeling_pi0.py:529-547)
  const int B = n_obs * h->cams();
  const int T1 = h->n_img1(), T = B * T1, Wv = c.vis_width, hd = Wv / c.vis_heads, D = c.lm_width;
  const bf16 *b_patch, *pos, *post_w, *post_b, *w_proj, *b_proj;
  CVB_TRY(W(h, VT + "embeddings.patch_embedding.bias", CVB_BF16, Wv, &b_patch));
  CVB_TRY(W(h, VT + "embeddings.position_embedding.weight", CVB_BF16, (int64_t)T1 * Wv, &pos));
  if (B > 1) pos = s.pos_tiled;
  // split-K + LayerNorm-reduce path: single-observation (latency) handles only.  A handle built for batches takes the
  // fused-epilogue GEMMs for every B, so a row's result never depends on how many observations share the call.
  const bool sk = h->max_obs() == 1 && T <= 256;  // (the split-K kernel holds all rows in one UMMA N)
  CVB_TRY(W(h, VT + "post_layernorm.weight", CVB_BF16, Wv, &post_w));
  CVB_TRY(W(h, VT + "post_layernorm.bias", CVB_BF16, Wv, &post_b));
  CVB_TRY(W(h, MM + "weight", CVB_BF16, (int64_t)D * Wv, &w_proj));
  CVB_TRY(W(h, MM + "bias", CVB_BF16, D, &b_proj));

  CVB_TRY(im2col_patches(st, s.in_image, s.patches, 3, c.vis_image, c.vis_image, c.vis_patch, s.kpad, B));
  // hv = bf16(bf16(conv + bias) + pos_emb)
  CVB_TRY(gemm(st, s.patches, s.kpad, s.w_patch, s.kpad, T, Wv, s.kpad, EPI_RESID, s.hv, Wv, b_patch, pos, Wv));
  // out_proj / fc2 leave split-K partials that the following LayerNorm kernel reduces (bias + residual + norm in one
  // pass, ops_misc.cu layernorm_reduce_kernel); pending = partials of the previous fc2 not yet folded into hv
  int pending = 0;
  const bf16* pending_bias = nullptr;
  auto gemm_part = [&](const bf16* A, long lda, const bf16* Wt, int Kd, int splits, int* used) {
    GemmCall g;
    g.A = A, g.lda = lda, g.W = Wt, g.ldw = Kd, g.M = T, g.N = Wv, g.K = Kd, g.C = s.part_v, g.ldc = Wv;
    return gemm_splitk_partial(st, g, splits, used);
  };
  for (int l = 0; l < c.vis_layers; ++l) {
    const VisLayer& L = s.vis[l];
    if (pending > 0)
      CVB_TRY(layernorm_reduce(st, s.part_v, pending, (long)T * Wv, Wv, pending_bias, s.hv, Wv, L.ln1_w, L.ln1_b, s.hv, Wv,
                               s.xv, Wv, T, Wv, 1e-6f));
    else
      CVB_TRY(layernorm_bf16(st, s.hv, Wv, L.ln1_w, L.ln1_b, s.xv, Wv, T, Wv, 1e-6f));
    pending = 0;
    CVB_TRY(gemm(st, s.xv, Wv, L.wqkv, Wv, T, 3 * Wv, Wv, EPI_STORE, s.qkv_v, 3 * Wv, L.bqkv));
    AttnCall a;
    a.q = s.qkv_v, a.q_batch_stride = (long)T1 * 3 * Wv, a.q_row_stride = 3 * Wv;
    a.k0 = s.qkv_v + Wv, a.v0 = s.qkv_v + 2 * Wv, a.kv0_batch_stride = (long)T1 * 3 * Wv, a.kv0_row_stride = 3 * Wv;
    a.kv0_len = T1, a.q_per_kv_batch = 1;
    a.out = s.attn_v, a.o_batch_stride = (long)T1 * Wv, a.o_row_stride = Wv;
    a.batches = B, a.heads = c.vis_heads, a.kv_heads = c.vis_heads, a.tq = T1, a.head_dim = hd;
    a.scale = 1.0f / sqrtf(static_cast<float>(hd));
    CVB_TRY(attention(st, a));
    if (sk && s.splitk_vo > 0) {
      int used = 0;
      CVB_TRY(gemm_part(s.attn_v, Wv, L.wo, Wv, s.splitk_vo, &used));
      CVB_TRY(layernorm_reduce(st, s.part_v, used, (long)T * Wv, Wv, L.bo, s.hv, Wv, L.ln2_w, L.ln2_b, s.hv, Wv, s.xv, Wv,
                               T, Wv, 1e-6f));
    } else {
      CVB_TRY(gemm(st, s.attn_v, Wv, L.wo, Wv, T, Wv, Wv, EPI_RESID, s.hv, Wv, L.bo, s.hv, Wv));
      CVB_TRY(layernorm_bf16(st, s.hv, Wv, L.ln2_w, L.ln2_b, s.xv, Wv, T, Wv, 1e-6f));
    }
    CVB_TRY(gemm(st, s.xv, Wv, L.w1, Wv, T, c.vis_mlp, Wv, EPI_GELU, s.mlp_v, c.vis_mlp, L.b1));
    if (sk && s.splitk_v2 > 0) {
      CVB_TRY(gemm_part(s.mlp_v, c.vis_mlp, L.w2, c.vis_mlp, s.splitk_v2, &pending));
      pending_bias = L.b2;
    } else {
      CVB_TRY(gemm(st, s.mlp_v, c.vis_mlp, L.w2, c.vis_mlp, T, Wv, c.vis_mlp, EPI_RESID, s.hv, Wv, L.b2, s.hv, Wv));
    }
  }
  if (pending > 0)
    CVB_TRY(layernorm_reduce(st, s.part_v, pending, (long)T * Wv, Wv, pending_bias, s.hv, Wv, post_w, post_b, s.hv, Wv, s.xv,
                             Wv, T, Wv, 1e-6f));
  else
    CVB_TRY(layernorm_bf16(st, s.hv, Wv, post_w, post_b, s.xv, Wv, T, Wv, 1e-6f));
  CVB_TRY(gemm(st, s.xv, Wv, w_proj, Wv, T, D, Wv, EPI_STORE, s.proj_out, D, b_proj));
  return 0;
}

static int run_prefix(cvb_handle* h, cudaStream_t st, int B, int Rpo) {
  const cvb_config& c = h->cfg;
  Pi0State& s = h->pi0;
  tl_batch_handle = h->max_obs() > 1;
  // Right-padded language tokens are masked as keys and their own rows are never read (dropping them
  // is bit-exact on the reference), so only Pe = image tokens + lang_rows() rows per prompt are processed; the KV
  // cache keeps the full-P layout the denoise attention indexes.
  const int T = h->n_img(), P = h->prefix_len(), Le = h->lang_rows(), Pe = T + Le, D = c.lm_width, hd = c.head_dim;
  const int R = B * Rpo;  // global rephrase count; rephrase r belongs to observation r / Rpo
  const int qd = c.heads * hd, qkvw = qd + 2 * hd, M = R * Pe;
  const bf16* embed;
  CVB_TRY(W(h, LM + "embed_tokens.weight", CVB_BF16, (int64_t)c.vocab * D, &embed));
  CVB_TRY(build_prefix(st, s.proj_out, embed, s.in_tokens, s.hp, R, T, Le, c.max_lang_len, D, Rpo));
  CVB_TRY(prefix_lengths(st, s.in_lang_len, s.plen, R, T, Le));
  const long layer_stride = (long)h->rm_total() * P * hd;
  for (int l = 0; l < c.layers; ++l) {
    const GemmaLayer& L = s.lm[l];
    bf16* kc = s.kcache + l * layer_stride;
    bf16* vc = s.vcache + l * layer_stride;
    CVB_TRY(rmsnorm(st, s.hp, 0, D, L.in_norm, 0, s.xp, D, M, D, 1e-6f, nullptr));
    CVB_TRY(gemm(st, s.xp, D, L.wqkv, D, M, qkvw, D, EPI_STORE, s.qkv_p, qkvw));
    // tcgen05 attention (8 heads folded into UMMA rows, V^T written by the RoPE kernel) when the shape allows it
    UmmaAttnCall u;
    u.q = s.qkv_p, u.q_ld = qkvw, u.q_total_rows = M, u.q_rows_per_batch = Pe;
    u.k = kc, u.k_total_rows = (long)h->rm_total() * P, u.k_rows_per_batch = P;
    bf16* vt = s.vt_p + (size_t)l * h->rm_total() * hd * s.vt_ld;
    u.vt = vt, u.vt_ld = s.vt_ld, u.klen_dev = s.plen, u.kmax = Pe;
    u.out = s.attn_p, u.o_batch_stride = (long)Pe * qd, u.o_row_stride = qd;
    u.batches = R, u.tq = Pe, u.heads = c.heads, u.head_dim = hd, u.scale = 1.0f / sqrtf(static_cast<float>(hd));
    static const bool umma_off = getenv("CVB_NO_UMMA_ATTN") != nullptr;
    const bool use_umma = !umma_off && attention_umma_eligible(u);
    // V^T is written for every layer (the last one included: the denoise attention reads all 18 caches)
    CVB_TRY(rope_qkv(st, s.qkv_p, qkvw, s.rope_timescale, M, c.heads, hd, Pe, nullptr, 1, kc, vc, (long)P * hd, hd,
                     umma_off ? nullptr : vt, (long)hd * s.vt_ld, s.vt_ld));
    if (l == c.layers - 1) break;  // only this layer's K/V are consumed (modeling_pi0.py:688-695)
    if (use_umma) {
      CVB_TRY(attention_umma(st, u));
    } else {
      AttnCall a;
      a.q = s.qkv_p, a.q_batch_stride = (long)Pe * qkvw, a.q_row_stride = qkvw;
      a.k0 = kc, a.v0 = vc, a.kv0_batch_stride = (long)P * hd, a.kv0_row_stride = hd;
      a.kv0_len_dev = s.plen, a.kv0_max = Pe, a.q_per_kv_batch = 1;
      a.out = s.attn_p, a.o_batch_stride = (long)Pe * qd, a.o_row_stride = qd;
      a.batches = R, a.heads = c.heads, a.kv_heads = 1, a.tq = Pe, a.head_dim = hd;
      a.scale = 1.0f / sqrtf(static_cast<float>(hd));
      CVB_TRY(attention(st, a));
    }
    CVB_TRY(gemm(st, s.attn_p, qd, L.wo, qd, M, D, qd, EPI_RESID, s.hp, D, nullptr, s.hp, D));
    CVB_TRY(rmsnorm(st, s.hp, 0, D, L.post_norm, 0, s.xp, D, M, D, 1e-6f, nullptr));
    const int packed = ((c.lm_mlp + 127) / 128) * 256;
    CVB_TRY(gemm(st, s.xp, D, L.wgu, D, M, packed, D, EPI_GEGLU, s.act_p, c.lm_mlp, nullptr, nullptr, 0, 0, c.lm_mlp));
    CVB_TRY(gemm(st, s.act_p, c.lm_mlp, L.wd, c.lm_mlp, M, D, c.lm_mlp, EPI_RESID, s.hp, D, nullptr, s.hp, D));
  }
  return 0;
}

static int run_denoise(cvb_handle* h, cudaStream_t st, int B, int Rpo, int K) {
  const cvb_config& c = h->cfg;
  Pi0State& s = h->pi0;
  tl_batch_handle = h->max_obs() > 1;
  const int P = h->prefix_len(), S = h->suffix_len(), We = c.ex_width, hd = c.head_dim;
  const int R = B * Rpo;
  const int qd = c.heads * hd, qkvw = qd + 2 * hd, N = R * K, M = N * S, Ma = N * c.chunk_size;
  const bool sk = M <= 256 && h->max_obs() == 1;  // split-K + RMSNorm-reduce path (latency handles, one UMMA N of rows)
  const float *w_state, *b_state, *w_ain, *b_ain, *w_in, *b_in, *w_out, *b_out, *w_aout, *b_aout, *w_norm;
  CVB_TRY(W(h, "state_proj.weight", CVB_F32, (int64_t)We * c.max_state_dim, &w_state));
  CVB_TRY(W(h, "state_proj.bias", CVB_F32, We, &b_state));
  CVB_TRY(W(h, "action_in_proj.weight", CVB_F32, (int64_t)We * c.max_action_dim, &w_ain));
  CVB_TRY(W(h, "action_in_proj.bias", CVB_F32, We, &b_ain));
  CVB_TRY(W(h, "action_time_mlp_in.weight", CVB_F32, (int64_t)We * 2 * We, &w_in));
  CVB_TRY(W(h, "action_time_mlp_in.bias", CVB_F32, We, &b_in));
  CVB_TRY(W(h, "action_time_mlp_out.weight", CVB_F32, (int64_t)We * We, &w_out));
  CVB_TRY(W(h, "action_time_mlp_out.bias", CVB_F32, We, &b_out));
  CVB_TRY(W(h, "action_out_proj.weight", CVB_F32, (int64_t)We * c.max_action_dim, &w_aout));
  CVB_TRY(W(h, "action_out_proj.bias", CVB_F32, c.max_action_dim, &b_aout));
  CVB_TRY(W(h, EX + "norm.weight", CVB_F32, We, &w_norm));

  // state token (identical for every candidate and every step)
  {
    SgemmCall g;
    g.A = s.in_state, g.lda = c.max_state_dim, g.W = w_state, g.ldw = c.max_state_dim;
    g.M = B, g.N = We, g.K = c.max_state_dim, g.C = s.state_emb, g.ldc = We, g.bias = b_state;
    CVB_TRY(sgemm_f32(st, g));
    CVB_TRY(fill_state_rows(st, s.state_emb, s.suffix, N, We, S, Rpo * K));
  }
  const long layer_stride = (long)h->rm_total() * P * hd;
  const int gu_half = s.ex_gu_half;
  const int packed = ((c.ex_mlp + gu_half - 1) / gu_half) * 2 * gu_half;
  // RoPE of the suffix is applied inside the cluster decode attention (from a table built once per sample) when the
  // shape is eligible; otherwise by the standalone kernel
  AttnCall probe;
  probe.k1 = s.qkv_e, probe.kv1_len = S, probe.kv0_len_dev = s.plen, probe.kv0_max = h->n_img() + h->lang_rows();
  probe.heads = c.heads, probe.kv_heads = 1, probe.tq = S, probe.head_dim = hd;
  probe.q_per_kv_batch = K, probe.batches = N;
  probe.kv0_row_stride = hd, probe.kv0_batch_stride = (long)P * hd, probe.vt0 = getenv("CVB_NO_UMMA_ATTN") == nullptr ? s.vt_p : nullptr, probe.vt0_ld = s.vt_ld;
  const bool fused_rope = attention_decode_umma_eligible(probe);
  if (fused_rope) CVB_TRY(rope_table(st, s.rope_timescale, s.plen, R, S, hd / 2, s.rope_tab));
  // Persistent expert kernel: one launch per layer covers o_proj -> norm -> gate/up -> down -> norm -> next qkv with
  // device-wide barriers instead of kernel boundaries and a weight ring that prefetches across them; only the
  // attention stays a separate launch.
  const MegaProgram* pg = nullptr;
  const bool mega = s.mega.mode != 0 && attention_decode_umma_eligible(probe) && M <= 256 && B == 1;
  if (mega) CVB_TRY(expert_mega_program(h, M, st, &pg));
  // F7 hoist: the suffix's state token attends the prefix and itself only and its input never changes, so
  // its K / V of every layer are identical in all denoise steps.  Step 0 runs all S rows per candidate and keeps them
  // (the tcgen05 attention kernel writes the rotated key / the value while staging); steps 1.. run the chunk_size action
  // rows only (M = 4 N instead of 5 N: -20 % GEMM rows, activation and partial traffic) and read the state key / value
  // from the cache.  Bit-identical to recomputing it (tests/test_pi0_gpu.py::test_state_token_hoist_is_exact).
  const bool hoist = !mega && getenv("CVB_NO_HOIST") == nullptr && attention_decode_umma_eligible(probe) && S == c.chunk_size + 1 &&
                     s.times.size() > 1;
  const long Nm_all = (long)h->rm_total() * c.max_samples;
  for (size_t step = 0; step < s.times.size(); ++step) {
    const bool hoisted = hoist && step > 0;  // this step runs without the state rows
    const int Sc = hoisted ? S - 1 : S;      // suffix rows per candidate in this step
    const int M = N * Sc;
    {  // embed_suffix (modeling_pi0.py:598-609), time half of mlp_in folded into time_vec[step]
      SgemmCall g2;
      if (s.w_ain_comb != nullptr) {
        g2.A = s.x_t, g2.lda = c.max_action_dim, g2.W = s.w_ain_comb, g2.ldw = c.max_action_dim;
        g2.M = Ma, g2.N = We, g2.K = c.max_action_dim, g2.bias = s.b_ain_comb;
      } else {
        SgemmCall g;
        g.A = s.x_t, g.lda = c.max_action_dim, g.W = w_ain, g.ldw = c.max_action_dim;
        g.M = Ma, g.N = We, g.K = c.max_action_dim, g.C = s.a1, g.ldc = We, g.bias = b_ain;
        CVB_TRY(sgemm_f32(st, g));
        g2.A = s.a1, g2.lda = We, g2.W = w_in, g2.ldw = 2 * We, g2.M = Ma, g2.N = We, g2.K = We, g2.bias = b_in;
      }
      g2.C = s.a2, g2.ldc = We, g2.row_bias = s.time_vec + step * We, g2.act = SACT_SILU;
      CVB_TRY(sgemm_f32(st, g2));
      // (latency handles only: on a batch-capable handle a row's arithmetic must not depend on the number of observations)
      if (s.w_out3 != nullptr && Ma <= 256 && h->max_obs() == 1) {  // 3-term bf16 split-K GEMM + fp32 reduce (+ bias, row layout)
        CVB_TRY(split3_rows(st, s.a2, We, s.a2s, Ma, We, 0));
        GemmCall gt;
        gt.A = s.a2s, gt.lda = 3 * We, gt.W = s.w_out3, gt.ldw = 3 * We, gt.M = Ma, gt.N = We, gt.K = 3 * We;
        gt.C = s.part_e, gt.ldc = We;
        int used = 0;
        CVB_TRY(gemm_splitk_partial(st, gt, kMaxSplitK, &used));
        CVB_TRY(partial_reduce_f32(st, s.part_e, used, (long)Ma * We, We, b_out, s.suffix, We, Ma, We,
                                   hoisted ? 0 : c.chunk_size));
      } else {
      SgemmCall g3;
      g3.A = s.a2, g3.lda = We, g3.W = w_out, g3.ldw = We, g3.M = Ma, g3.N = We, g3.K = We;
      g3.C = s.suffix, g3.ldc = We, g3.bias = b_out, g3.out_group = hoisted ? 0 : c.chunk_size;
      CVB_TRY(sgemm_f32(st, g3));
      }
    }
    if (mega) {
      const int Mmax = h->rm_total() * c.max_samples * S;
      int first = 0;
      for (int l = 0; l < c.layers; ++l) {
        CVB_TRY(expert_mega_launch(h, st, *pg, first, pg->attn_after[l] - first));
        first = pg->attn_after[l];
        AttnCall a;
        a.rope = s.rope_tab, a.kv0_static = 1;
        a.q = s.qkv_e, a.q_batch_stride = (long)S * qkvw, a.q_row_stride = qkvw;
        a.k0 = s.kcache + l * layer_stride, a.v0 = s.vcache + l * layer_stride;
        a.kv0_batch_stride = (long)P * hd, a.kv0_row_stride = hd, a.kv0_len_dev = s.plen, a.kv0_max = h->n_img() + h->lang_rows();
        a.q_per_kv_batch = K;
        a.vt0 = s.vt_p + (size_t)l * h->rm_total() * hd * s.vt_ld, a.vt0_ld = s.vt_ld;
        a.k1 = s.qkv_e + qd, a.v1 = s.qkv_e + qd + hd, a.kv1_batch_stride = (long)S * qkvw;
        a.kv1_row_stride = qkvw, a.kv1_len = S, a.suffix_mask = 1;
        a.q_part = s.mega.part_qkv, a.k1_part = s.mega.part_qkv + qd, a.v1_part = s.mega.part_qkv + qd + hd;
        a.part_splits = pg->host[pg->attn_after[l] - 1].splits, a.part_split_stride = (long)Mmax * qkvw;
        a.out = s.attn_e, a.o_batch_stride = (long)S * qd, a.o_row_stride = qd;
        a.batches = N, a.heads = c.heads, a.kv_heads = 1, a.tq = S, a.head_dim = hd;
        a.scale = 1.0f / sqrtf(static_cast<float>(hd));
        a.algo = 3;
        CVB_TRY(attention(st, a));
      }
      CVB_TRY(expert_mega_launch(h, st, *pg, first, static_cast<int>(pg->host.size()) - first));
      CVB_TRY(action_out_euler(st, s.xe, We, w_aout, b_aout, s.x_t, step == 0 ? s.v0 : nullptr, N, We,
                               c.max_action_dim, c.chunk_size, S, s.dt));
      continue;
    }
    // pending = split-K partials of the previous down_proj still to be folded into he by the next norm
    int pending = 0;
    auto gemm_part = [&](const bf16* A, long lda, const bf16* Wt, int Kd, int splits, int* used) {
      GemmCall g;
      g.A = A, g.lda = lda, g.W = Wt, g.ldw = Kd, g.M = M, g.N = We, g.K = Kd, g.C = s.part_e, g.ldc = We;
      return gemm_splitk_partial(st, g, splits, used);
    };
    for (int l = 0; l < c.layers; ++l) {
      const GemmaLayer& L = s.ex[l];
      const void* resid = l == 0 ? static_cast<const void*>(s.suffix) : static_cast<const void*>(s.he);
      const int resid_f32 = l == 0 ? 1 : 0;
      if (pending > 0)  // he = he + down_proj(l-1), xe = input_layernorm(he)
        CVB_TRY(rmsnorm_reduce(st, s.part_e, pending, (long)M * We, We, s.he, 0, We, L.in_norm, 0, s.he, We, s.xe, We, M,
                               We, 1e-6f));
      else
        CVB_TRY(rmsnorm(st, resid, resid_f32, We, L.in_norm, 0, s.xe, We, M, We, 1e-6f, nullptr));
      pending = 0;
      CVB_TRY(gemm(st, s.xe, We, L.wqkv, We, M, qkvw, We, EPI_STORE, s.qkv_e, qkvw));
      if (!fused_rope)
        CVB_TRY(rope_qkv(st, s.qkv_e, qkvw, s.rope_timescale, M, c.heads, hd, S, s.plen, K, nullptr, nullptr, 0, 0));
      AttnCall a;
      a.rope = fused_rope ? s.rope_tab : nullptr;
      a.kv0_static = 1;
      a.q = s.qkv_e, a.q_batch_stride = (long)Sc * qkvw, a.q_row_stride = qkvw;
      a.k0 = s.kcache + l * layer_stride, a.v0 = s.vcache + l * layer_stride;
      a.kv0_batch_stride = (long)P * hd, a.kv0_row_stride = hd, a.kv0_len_dev = s.plen, a.kv0_max = h->n_img() + h->lang_rows();
      a.q_per_kv_batch = K;
      if (getenv("CVB_NO_UMMA_ATTN") == nullptr) a.vt0 = s.vt_p + (size_t)l * h->rm_total() * hd * s.vt_ld, a.vt0_ld = s.vt_ld;
      a.k1 = s.qkv_e + qd, a.v1 = s.qkv_e + qd + hd, a.kv1_batch_stride = (long)Sc * qkvw;
      a.kv1_row_stride = qkvw, a.kv1_len = S, a.suffix_mask = 1;
      a.out = s.attn_e, a.o_batch_stride = (long)Sc * qd, a.o_row_stride = qd;
      a.batches = N, a.heads = c.heads, a.kv_heads = 1, a.tq = Sc, a.head_dim = hd;
      a.scale = 1.0f / sqrtf(static_cast<float>(hd));
      if (hoist) {
        bf16* ck = s.state_k + (size_t)l * Nm_all * hd;
        bf16* cv = s.state_v + (size_t)l * Nm_all * hd;
        a.rope_rows = S;
        if (hoisted) {  // suffix keys = [cached state key, the chunk_size action keys of this step]; no state query row
          a.kv1_cached_k = ck, a.kv1_cached_v = cv, a.kv1_cached = 1, a.suffix_mask = 0, a.rope_off = 1;
        } else {
          a.kv1_cache_out_k = ck, a.kv1_cache_out_v = cv;
        }
      }
      CVB_TRY(attention(st, a));
      if (sk && s.splitk_o > 0) {
        int used = 0;
        CVB_TRY(gemm_part(s.attn_e, qd, L.wo, qd, s.splitk_o, &used));
        CVB_TRY(rmsnorm_reduce(st, s.part_e, used, (long)M * We, We, resid, resid_f32, We, L.post_norm, 0, s.he, We,
                               s.xe, We, M, We, 1e-6f));
      } else {
        CVB_TRY(gemm(st, s.attn_e, qd, L.wo, qd, M, We, qd, EPI_RESID, s.he, We, nullptr, resid, We, resid_f32));
        CVB_TRY(rmsnorm(st, s.he, 0, We, L.post_norm, 0, s.xe, We, M, We, 1e-6f, nullptr));
      }
      if (gu_half == 64)  // epilogue 5 + force_bn 128: general kernel, 128 x 128 GeGLU tiles
        CVB_TRY(gemm(st, s.xe, We, L.wgu, We, M, packed, We, 5, s.act_e, c.ex_mlp, nullptr, nullptr, 0, 0, c.ex_mlp, 128));
      else
        CVB_TRY(gemm(st, s.xe, We, L.wgu, We, M, packed, We, EPI_GEGLU, s.act_e, c.ex_mlp, nullptr, nullptr, 0, 0, c.ex_mlp));
      if (sk && s.splitk_d > 0) {
        CVB_TRY(gemm_part(s.act_e, c.ex_mlp, L.wd, c.ex_mlp, s.splitk_d, &pending));
      } else {
        CVB_TRY(gemm(st, s.act_e, c.ex_mlp, L.wd, c.ex_mlp, M, We, c.ex_mlp, EPI_RESID, s.he, We, nullptr, s.he, We));
      }
    }
    if (pending > 0)
      CVB_TRY(rmsnorm_reduce(st, s.part_e, pending, (long)M * We, We, s.he, 0, We, w_norm, 1, s.he, We, s.xe, We, M, We,
                             1e-6f));
    else
      CVB_TRY(rmsnorm(st, s.he, 0, We, w_norm, 1, s.xe, We, M, We, 1e-6f, nullptr));
    CVB_TRY(action_out_euler(st, s.xe, We, w_aout, b_aout, s.x_t, step == 0 ? s.v0 : nullptr, N, We,
                             c.max_action_dim, c.chunk_size, Sc, s.dt));
  }
  return 0;
}

static int run_all(cvb_handle* h, cudaStream_t st, int B, int R, int K) {
  CVB_TRY(run_vision(h, st, B));
  CVB_TRY(run_prefix(h, st, B, R));
  CVB_TRY(run_denoise(h, st, B, R, K));
  return 0;
}

int pi0_stage_inputs(cvb_handle* h, const float* image, const int64_t* tokens, const int32_t* lang_len,
                     const float* state, const float* noise, int R, int K, cudaStream_t st, int B) {
  const cvb_config& c = h->cfg;
  Pi0State& s = h->pi0;
  CVB_REQUIRE(h->finalized, "cvb_finalize() has not been called");
  CVB_REQUIRE(c.layers > 0, "this handle was created without the pi0 model (layers == 0)");
  CVB_REQUIRE(B >= 1 && B <= h->max_obs(), "number of observations out of range (max_observations)");
  CVB_REQUIRE(R >= 1 && R <= c.max_rephrases, "R out of range (max_rephrases)");
  CVB_REQUIRE(K >= 1 && K <= c.max_samples, "K out of range (max_samples)");
  CVB_REQUIRE(image != nullptr && tokens != nullptr && lang_len != nullptr && state != nullptr && noise != nullptr,
              "null input");
  const size_t act_bytes = (size_t)B * R * K * c.chunk_size * c.max_action_dim * sizeof(float);
  CVB_CUDA(cudaMemcpyAsync(s.in_image, image, (size_t)B * h->cams() * 3 * c.vis_image * c.vis_image * sizeof(float),
                           cudaMemcpyDeviceToDevice, st));
  CVB_CUDA(cudaMemcpyAsync(s.in_tokens, tokens, (size_t)B * R * c.max_lang_len * sizeof(int64_t),
                           cudaMemcpyDeviceToDevice, st));
  CVB_CUDA(cudaMemcpyAsync(s.in_lang_len, lang_len, (size_t)B * R * sizeof(int32_t), cudaMemcpyDeviceToDevice, st));
  CVB_CUDA(cudaMemcpyAsync(s.in_state, state, (size_t)B * c.max_state_dim * sizeof(float), cudaMemcpyDeviceToDevice, st));
  CVB_CUDA(cudaMemcpyAsync(s.x_t, noise, act_bytes, cudaMemcpyDeviceToDevice, st));
  return 0;
}

int pi0_enqueue(cvb_handle* h, cudaStream_t st, int R, int K, int part, int B) {
  if (part == 0) {
    CVB_TRY(run_vision(h, st, B));
    return run_prefix(h, st, B, R);
  }
  return run_denoise(h, st, B, R, K);
}

float* pi0_actions_buffer(cvb_handle* h) { return h->pi0.x_t; }

int pi0_sample(cvb_handle* h, const float* image, const int64_t* tokens, const int32_t* lang_len,
               const float* state, const float* noise, int R, int K, float* actions, cudaStream_t st, int B) {
  const cvb_config& c = h->cfg;
  Pi0State& s = h->pi0;
  CVB_TRY(pi0_stage_inputs(h, image, tokens, lang_len, state, noise, R, K, st, B));
  const size_t act_bytes = (size_t)B * R * K * c.chunk_size * c.max_action_dim * sizeof(float);
  const long key = ((long)h->cams() << 56) | ((long)h->lang_rows() << 40) | ((long)B << 28) | ((long)R << 16) | (long)K;
  CVB_TRY(s.graphs.run(c.use_cuda_graph != 0, key, st, [&](cudaStream_t cs) { return run_all(h, cs, B, R, K); }));
  CVB_CUDA(cudaMemcpyAsync(actions, s.x_t, act_bytes, cudaMemcpyDeviceToDevice, st));
  return 0;
}

// profiling hook: run one phase eagerly on the staged inputs of the last cvb_pi0_sample call
int pi0_run_phase(cvb_handle* h, int phase, int R, int K, cudaStream_t st, int B) {
  CVB_REQUIRE(h->finalized, "cvb_finalize() has not been called");
  CVB_REQUIRE(R >= 1 && R <= h->cfg.max_rephrases && K >= 1 && K <= h->cfg.max_samples, "R/K out of range");
  CVB_REQUIRE(B >= 1 && B <= h->max_obs(), "number of observations out of range");
  switch (phase) {
    case 0:
      return run_vision(h, st, B);
    case 1:
      return run_prefix(h, st, B, R);
    case 2:
      return run_denoise(h, st, B, R, K);
    default:
      set_last_error("phase must be 0 (vision), 1 (prefix) or 2 (denoise)");
      return -1;
  }
}

int64_t pi0_debug_copy(cvb_handle* h, const std::string& name, void* dst, int64_t max_bytes,
                       cudaStream_t st) {
  const cvb_config& c = h->cfg;
  Pi0State& s = h->pi0;
  const int T = h->n_img(), P = h->prefix_len();
  const void* src = nullptr;
  int64_t bytes = 0;
  const int64_t layer_bytes = (int64_t)h->rm_total() * P * c.head_dim * sizeof(bf16);
  if (name == "image_emb") {
    src = s.proj_out, bytes = (int64_t)h->max_obs() * h->n_img_max() * c.lm_width * sizeof(bf16);
  } else if (name == "vision_hidden") {
    src = s.xv, bytes = (int64_t)T * c.vis_width * sizeof(bf16);
  } else if (name == "prefix_k0") {
    src = s.kcache, bytes = layer_bytes;
  } else if (name == "prefix_v0") {
    src = s.vcache, bytes = layer_bytes;
  } else if (name == "prefix_klast") {
    src = s.kcache + (int64_t)(c.layers - 1) * (layer_bytes / sizeof(bf16)), bytes = layer_bytes;
  } else if (name == "prefix_vlast") {
    src = s.vcache + (int64_t)(c.layers - 1) * (layer_bytes / sizeof(bf16)), bytes = layer_bytes;
  } else if (name == "v0") {
    src = s.v0, bytes = (int64_t)h->rm_total() * c.max_samples * c.chunk_size * c.max_action_dim * sizeof(float);
  } else if (name == "time_emb") {
    src = s.time_emb_f32, bytes = (int64_t)s.times.size() * c.ex_width * sizeof(float);
  } else if (name == "suffix") {
    src = s.suffix, bytes = (int64_t)h->rm_total() * c.max_samples * h->suffix_len() * c.ex_width * sizeof(float);
  } else {
    set_last_error("unknown debug buffer: " + name);
    return -1;
  }
  if (bytes > max_bytes) bytes = max_bytes;
  if (cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, st) != cudaSuccess) {
    set_last_error("debug copy failed");
    return -2;
  }
  return bytes;
}

}  // namespace cvb
