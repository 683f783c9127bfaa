// CoVer verifier pipeline: SigLIP2 trunk (bf16 tcgen05 GEMMs, up to the two hook points) -> fp32 heads
// of every ensemble member -> fused score / group-mean / argmax.
//
// Reference: EfficientEnsembleMerged.compute_max_similarity_scores_batch (efficient_ensemble_merged.py:309-454),
// get_embeddings_from_model_batch (:194-247), VLA_SigLIP2_Bridge.extract_features
// (finetune_trajectory_bridge_ddp.py:297-355).  De-duplication per SURVEY.md F3-F5: the image/text side
// is computed once per call (not N times), the last vision block stops after attn.proj, the text
// tower applies ln_final + text_projection to every token.
#include <cmath>
#include <cstdlib>

#include "engine.h"
#include "gemm_tcgen05.cuh"
#include "pi0_kernels.h"
#include "verifier_kernels.h"

namespace cvb {

struct TrunkBlock {
  const bf16 *ln1_w, *ln1_b, *wqkv, *bqkv, *wo, *bo, *ln2_w, *ln2_b, *w1, *b1, *w2, *b2;
};
struct TrajLayer {
  const float *w_in, *b_in, *wo, *bo, *w1, *b1, *w2, *b2, *n1w, *n1b, *n2w, *n2b;
  // [hi | lo | hi] bf16 copies of the four weight matrices: the B operands of the 3-term bf16 tensor-core GEMMs
  bf16 *w_in3 = nullptr, *wo3 = nullptr, *w13 = nullptr, *w23 = nullptr;
};
struct MemberW {
  const float *temp, *pos_emb, *w_ss, *b_ss;
  // vf_traj_layers == 0 (use_transformer = False checkpoints): nn.Sequential(Linear, LayerNorm, ReLU, Dropout, Linear)
  const float *mlp_w0 = nullptr, *mlp_b0 = nullptr, *mlp_lnw = nullptr, *mlp_lnb = nullptr, *mlp_w1 = nullptr,
              *mlp_b1 = nullptr;
  float* wkv_v;  // owned [L*2E, W] (vision pooling K/V projections of all blocks)
  float* bkv_v;  // owned [L*2E]
  std::vector<TrajLayer> traj;
};

struct VerifierState {
  std::vector<TrunkBlock> vis, txt;
  const bf16 *patch_b, *pos_embed, *tok_emb, *txt_pos, *lnf_w, *lnf_b, *wproj, *bproj;
  bf16* w_patch = nullptr;
  bf16* pos_tiled = nullptr;  // pos_embed repeated max_observations times (residual operand of the batched patch GEMM)
  int kpad = 0;
  std::vector<MemberW> mem;
  float* wkv_t = nullptr;  // owned [M*L*2E, W] text pooling K/V projections of all members
  float* bkv_t = nullptr;
  PoolChain* chains = nullptr;  // device [2*M]
  ItFinal* itf = nullptr;       // device [M]
  // inputs
  float* in_image = nullptr;
  int64_t* in_tokens = nullptr;
  float* in_traj = nullptr;
  // trunk workspace
  bf16 *patches = nullptr, *hv = nullptr, *xv = nullptr, *qkv = nullptr, *att = nullptr, *mlp = nullptr,
       *pfeat = nullptr, *ht = nullptr, *tfeat = nullptr;
  // head workspace
  float *Pn = nullptr, *Tn = nullptr, *sim = nullptr, *pe = nullptr, *taf = nullptr, *kv_v = nullptr,
        *kv_t = nullptr, *vtok = nullptr, *ttok = nullptr;
  float* it_obs = nullptr;  // [max_observations][members][embed]: what the score kernel reads (slot 0 for single calls)
  bf16 *txs = nullptr, *tatts = nullptr, *tffs = nullptr;  // [hi | hi | lo] copies of tx / tatt / relu(tff) per member
  bool traj_tc = false;                                     // trajectory-encoder GEMMs on the tensor cores (3-term bf16 split)
  float *tx = nullptr, *tqkv = nullptr, *tatt = nullptr, *ty = nullptr, *tff = nullptr, *act = nullptr;
  float* scores = nullptr;
  float* gmean = nullptr;
  int* bidx = nullptr;
  float* bscore = nullptr;
  bool context_valid = false;
  // per-task prompt cache (cvb_verifier_hold_text): while text_hold is set and the text features of
  // `text_nb` observations are resident (Tn), a context call skips the text tower
  bool text_hold = false;
  int text_nb = 0;
  GraphCache ctx_graph, traj_graph;
  // the ensemble members' trajectory encoders are independent until the fused score: one branch (stream) per member
  std::vector<cudaStream_t> mstream;
  std::vector<cudaEvent_t> mevent;
  cudaEvent_t ev_fork = nullptr;
};

namespace {
const std::string TRK = "verifier.trunk.";

template <typename T>
int W(cvb_handle* h, const std::string& key, int dtype, int64_t numel, const T** out) {
  const void* p = nullptr;
  CVB_TRY(get_weight(h, key, dtype, numel, &p));
  *out = reinterpret_cast<const T*>(p);
  return 0;
}
// set by run_context / run_member_trajectories: on a handle built for several observations a row's bits must not depend
// on how many observations share the call, so its GEMMs never take the row-count dependent skinny kernel
thread_local bool tl_batch_handle = false;

int gemm(cudaStream_t st, const bf16* A, long lda, const bf16* Wt, long ldw, int M, int N, int K, int epi, void* C,
         long ldc, const void* bias = nullptr, const void* resid = nullptr, long ldr = 0) {
  GemmCall c;
  c.no_skinny = tl_batch_handle ? 1 : 0;
  c.A = A, c.lda = lda, c.W = Wt, c.ldw = ldw, c.M = M, c.N = N, c.K = K, c.epi = epi;
  c.C = C, c.ldc = ldc, c.bias = bias, c.resid = resid, c.ldr = ldr;
  // One image's trunk (M = 576) runs beside the denoise loop, which is bound by L2 -> SM bytes: what the
  // step pays for this side work is the L2 traffic and the SMs it takes, not its own latency.  128-wide tiles (half the
  // activation re-reads and half the CTAs of the 64-wide tiles the stand-alone cost model picks for o_proj / fc2) make the
  // context 0.2 ms slower alone (4.06 -> 4.25 ms) and the decision 0.2-0.4 ms faster (tools/overlap_timeline.py: the
  // context costs the critical path 2.0 ms of its 4 ms).  CVB_CTX_BN = 0 restores the automatic choice, 64 / 256 force.
  static const int ctx_bn = getenv("CVB_CTX_BN") != nullptr ? atoi(getenv("CVB_CTX_BN")) : 128;
  if (M > 256 && M < 1024 && ctx_bn > 0) c.force_bn = ctx_bn;
  return gemm_bf16(st, c);
}
int sg(cudaStream_t st, const float* A, long lda, const float* Wt, long ldw, int M, int N, int K, float* C, long ldc,
       const float* bias = nullptr, int act = 0, const float* resid = nullptr, long ldr = 0, int w_kn = 0,
       int w_dynamic = 0) {
  SgemmCall c;
  c.A = A, c.lda = lda, c.W = Wt, c.ldw = ldw, c.M = M, c.N = N, c.K = K, c.C = C, c.ldc = ldc;
  c.bias = bias, c.act = act, c.resid = resid, c.ldr = ldr, c.w_kn = w_kn, c.w_dynamic = w_dynamic;
  return sgemm_f32(st, c);
}
}  // namespace

void verifier_required_weights(const cvb_config& c, std::vector<WeightSpec>* out) {
  auto add = [&](const std::string& k, int dt, std::vector<int64_t> shape) {
    out->push_back(WeightSpec{k, dt, std::move(shape)});
  };
  const int Wd = c.vf_width, E = c.vf_embed, Np = (c.vf_image / c.vf_patch) * (c.vf_image / c.vf_patch);
  const std::string v = TRK + "visual.trunk.";
  add(v + "patch_embed.proj.weight", CVB_BF16, {Wd, 3, c.vf_patch, c.vf_patch});
  add(v + "patch_embed.proj.bias", CVB_BF16, {Wd});
  add(v + "pos_embed", CVB_BF16, {1, Np, Wd});
  for (int l = 0; l < c.vf_layers; ++l) {
    const std::string p = v + "blocks." + std::to_string(l) + ".";
    add(p + "norm1.weight", CVB_BF16, {Wd});
    add(p + "norm1.bias", CVB_BF16, {Wd});
    add(p + "attn.qkv.weight", CVB_BF16, {3 * Wd, Wd});
    add(p + "attn.qkv.bias", CVB_BF16, {3 * Wd});
    add(p + "attn.proj.weight", CVB_BF16, {Wd, Wd});
    add(p + "attn.proj.bias", CVB_BF16, {Wd});
    if (l < c.vf_layers - 1) {
      add(p + "norm2.weight", CVB_BF16, {Wd});
      add(p + "norm2.bias", CVB_BF16, {Wd});
      add(p + "mlp.fc1.weight", CVB_BF16, {c.vf_mlp, Wd});
      add(p + "mlp.fc1.bias", CVB_BF16, {c.vf_mlp});
      add(p + "mlp.fc2.weight", CVB_BF16, {Wd, c.vf_mlp});
      add(p + "mlp.fc2.bias", CVB_BF16, {Wd});
    }
  }
  const std::string t = TRK + "text.";
  add(t + "token_embedding.weight", CVB_BF16, {c.vf_vocab, Wd});
  add(t + "positional_embedding", CVB_BF16, {c.vf_text_ctx, Wd});
  for (int l = 0; l < c.vf_text_layers; ++l) {
    const std::string p = t + "transformer.resblocks." + std::to_string(l) + ".";
    add(p + "ln_1.weight", CVB_BF16, {Wd});
    add(p + "ln_1.bias", CVB_BF16, {Wd});
    add(p + "attn.in_proj_weight", CVB_BF16, {3 * Wd, Wd});
    add(p + "attn.in_proj_bias", CVB_BF16, {3 * Wd});
    add(p + "attn.out_proj.weight", CVB_BF16, {Wd, Wd});
    add(p + "attn.out_proj.bias", CVB_BF16, {Wd});
    add(p + "ln_2.weight", CVB_BF16, {Wd});
    add(p + "ln_2.bias", CVB_BF16, {Wd});
    add(p + "mlp.c_fc.weight", CVB_BF16, {c.vf_mlp, Wd});
    add(p + "mlp.c_fc.bias", CVB_BF16, {c.vf_mlp});
    add(p + "mlp.c_proj.weight", CVB_BF16, {Wd, c.vf_mlp});
    add(p + "mlp.c_proj.bias", CVB_BF16, {Wd});
  }
  add(t + "ln_final.weight", CVB_BF16, {Wd});
  add(t + "ln_final.bias", CVB_BF16, {Wd});
  add(t + "text_projection.weight", CVB_BF16, {Wd, Wd});
  add(t + "text_projection.bias", CVB_BF16, {Wd});
  for (int m = 0; m < c.vf_members; ++m) {
    const std::string b = "verifier." + std::to_string(m) + ".";
    add(b + "text_aware_visual_extraction.temperature", CVB_F32, {});
    add(b + "text_aware_visual_extraction.pos_emb", CVB_F32, {Np, Wd});
    for (const char* pool : {"vision_poolings", "text_pooling"}) {
      const std::string p = b + pool + ".";
      add(p + "query", CVB_F32, {1, 1, E});
      add(p + "layer_norm.weight", CVB_F32, {E});
      add(p + "layer_norm.bias", CVB_F32, {E});
      for (int i = 0; i < c.vf_pool_layers; ++i) {
        const std::string q = p + "blocks." + std::to_string(i) + ".";
        add(q + "attention.q_proj_weight", CVB_F32, {E, E});
        add(q + "attention.k_proj_weight", CVB_F32, {E, Wd});
        add(q + "attention.v_proj_weight", CVB_F32, {E, Wd});
        add(q + "attention.in_proj_bias", CVB_F32, {3 * E});
        add(q + "attention.out_proj.weight", CVB_F32, {E, E});
        add(q + "attention.out_proj.bias", CVB_F32, {E});
        add(q + "mlp.fc1.weight", CVB_F32, {E, E});
        add(q + "mlp.fc1.bias", CVB_F32, {E});
        add(q + "mlp.fc2.weight", CVB_F32, {E, E});
        add(q + "mlp.fc2.bias", CVB_F32, {E});
        for (const char* nm : {"q_layer_norm", "layer_norm"}) {
          add(q + nm + ".weight", CVB_F32, {E});
          add(q + nm + ".bias", CVB_F32, {E});
        }
      }
    }
    add(b + "input_projection.weight", CVB_F32, {E, 2 * E});
    add(b + "input_projection.bias", CVB_F32, {E});
    if (c.vf_traj_layers == 0) {
      // MLP action encoder over the flattened history (finetune_trajectory_bridge_ddp.py:254-262, merged checkpoints with
      // use_transformer = False: efficient_ensemble_merged.py:161-171); hidden width = vf_traj_ff
      const std::string q = b + "complex_action_encoder.";
      add(q + "0.weight", CVB_F32, {c.vf_traj_ff, c.vf_history * c.vf_action_dim});
      add(q + "0.bias", CVB_F32, {c.vf_traj_ff});
      add(q + "1.weight", CVB_F32, {c.vf_traj_ff});
      add(q + "1.bias", CVB_F32, {c.vf_traj_ff});
      add(q + "4.weight", CVB_F32, {E, c.vf_traj_ff});
      add(q + "4.bias", CVB_F32, {E});
      continue;
    }
    add(b + "single_step_action_encoder.weight", CVB_F32, {E, c.vf_action_dim});
    add(b + "single_step_action_encoder.bias", CVB_F32, {E});
    for (int i = 0; i < c.vf_traj_layers; ++i) {
      const std::string q = b + "trajectory_encoder.layers." + std::to_string(i) + ".";
      add(q + "self_attn.in_proj_weight", CVB_F32, {3 * E, E});
      add(q + "self_attn.in_proj_bias", CVB_F32, {3 * E});
      add(q + "self_attn.out_proj.weight", CVB_F32, {E, E});
      add(q + "self_attn.out_proj.bias", CVB_F32, {E});
      add(q + "linear1.weight", CVB_F32, {c.vf_traj_ff, E});
      add(q + "linear1.bias", CVB_F32, {c.vf_traj_ff});
      add(q + "linear2.weight", CVB_F32, {E, c.vf_traj_ff});
      add(q + "linear2.bias", CVB_F32, {E});
      for (const char* nm : {"norm1", "norm2"}) {
        add(q + nm + ".weight", CVB_F32, {E});
        add(q + nm + ".bias", CVB_F32, {E});
      }
    }
  }
}

int verifier_finalize(cvb_handle* h, cudaStream_t st) {
  const cvb_config& c = h->cfg;
  h->vf = new VerifierState();
  VerifierState& s = *h->vf;
  const int Wd = c.vf_width, E = c.vf_embed, L = c.vf_pool_layers, M = c.vf_members;
  const int Np = (c.vf_image / c.vf_patch) * (c.vf_image / c.vf_patch), Tt = c.vf_text_ctx;
  const int Bm = h->max_obs();  // observations per batched call
  const int Nm = h->rm_total() * c.max_samples, S = c.vf_history;
  CVB_REQUIRE(Wd % 8 == 0 && c.vf_mlp % 8 == 0 && (Wd / c.vf_heads) % 8 == 0, "verifier trunk widths must be multiples of 8");
  CVB_REQUIRE(L <= kMaxPoolLayers, "too many pooling layers");
  CVB_REQUIRE(E % c.vf_pool_heads == 0, "embed must divide by pool heads");
  const std::string v = TRK + "visual.trunk.", t = TRK + "text.";
  // ---- trunk
  const int kreal = 3 * c.vf_patch * c.vf_patch;
  s.kpad = (kreal + 7) / 8 * 8;
  const bf16* wp;
  CVB_TRY(W(h, v + "patch_embed.proj.weight", CVB_BF16, (int64_t)Wd * kreal, &wp));
  CVB_TRY(dalloc_t(h, &s.w_patch, (size_t)Wd * s.kpad));
  CVB_CUDA(cudaMemsetAsync(s.w_patch, 0, (size_t)Wd * s.kpad * sizeof(bf16), st));
  CVB_CUDA(cudaMemcpy2DAsync(s.w_patch, s.kpad * sizeof(bf16), wp, kreal * sizeof(bf16), kreal * sizeof(bf16), Wd,
                             cudaMemcpyDeviceToDevice, st));
  CVB_TRY(W(h, v + "patch_embed.proj.bias", CVB_BF16, Wd, &s.patch_b));
  CVB_TRY(W(h, v + "pos_embed", CVB_BF16, (int64_t)Np * Wd, &s.pos_embed));
  s.vis.resize(c.vf_layers);
  for (int l = 0; l < c.vf_layers; ++l) {
    const std::string p = v + "blocks." + std::to_string(l) + ".";
    TrunkBlock& B = s.vis[l];
    memset(&B, 0, sizeof(B));
    CVB_TRY(W(h, p + "norm1.weight", CVB_BF16, Wd, &B.ln1_w));
    CVB_TRY(W(h, p + "norm1.bias", CVB_BF16, Wd, &B.ln1_b));
    CVB_TRY(W(h, p + "attn.qkv.weight", CVB_BF16, (int64_t)3 * Wd * Wd, &B.wqkv));
    CVB_TRY(W(h, p + "attn.qkv.bias", CVB_BF16, 3 * Wd, &B.bqkv));
    CVB_TRY(W(h, p + "attn.proj.weight", CVB_BF16, (int64_t)Wd * Wd, &B.wo));
    CVB_TRY(W(h, p + "attn.proj.bias", CVB_BF16, Wd, &B.bo));
    if (l < c.vf_layers - 1) {
      CVB_TRY(W(h, p + "norm2.weight", CVB_BF16, Wd, &B.ln2_w));
      CVB_TRY(W(h, p + "norm2.bias", CVB_BF16, Wd, &B.ln2_b));
      CVB_TRY(W(h, p + "mlp.fc1.weight", CVB_BF16, (int64_t)c.vf_mlp * Wd, &B.w1));
      CVB_TRY(W(h, p + "mlp.fc1.bias", CVB_BF16, c.vf_mlp, &B.b1));
      CVB_TRY(W(h, p + "mlp.fc2.weight", CVB_BF16, (int64_t)c.vf_mlp * Wd, &B.w2));
      CVB_TRY(W(h, p + "mlp.fc2.bias", CVB_BF16, Wd, &B.b2));
    }
  }
  CVB_TRY(W(h, t + "token_embedding.weight", CVB_BF16, (int64_t)c.vf_vocab * Wd, &s.tok_emb));
  CVB_TRY(W(h, t + "positional_embedding", CVB_BF16, (int64_t)Tt * Wd, &s.txt_pos));
  s.txt.resize(c.vf_text_layers);
  for (int l = 0; l < c.vf_text_layers; ++l) {
    const std::string p = t + "transformer.resblocks." + std::to_string(l) + ".";
    TrunkBlock& B = s.txt[l];
    CVB_TRY(W(h, p + "ln_1.weight", CVB_BF16, Wd, &B.ln1_w));
    CVB_TRY(W(h, p + "ln_1.bias", CVB_BF16, Wd, &B.ln1_b));
    CVB_TRY(W(h, p + "attn.in_proj_weight", CVB_BF16, (int64_t)3 * Wd * Wd, &B.wqkv));
    CVB_TRY(W(h, p + "attn.in_proj_bias", CVB_BF16, 3 * Wd, &B.bqkv));
    CVB_TRY(W(h, p + "attn.out_proj.weight", CVB_BF16, (int64_t)Wd * Wd, &B.wo));
    CVB_TRY(W(h, p + "attn.out_proj.bias", CVB_BF16, Wd, &B.bo));
    CVB_TRY(W(h, p + "ln_2.weight", CVB_BF16, Wd, &B.ln2_w));
    CVB_TRY(W(h, p + "ln_2.bias", CVB_BF16, Wd, &B.ln2_b));
    CVB_TRY(W(h, p + "mlp.c_fc.weight", CVB_BF16, (int64_t)c.vf_mlp * Wd, &B.w1));
    CVB_TRY(W(h, p + "mlp.c_fc.bias", CVB_BF16, c.vf_mlp, &B.b1));
    CVB_TRY(W(h, p + "mlp.c_proj.weight", CVB_BF16, (int64_t)c.vf_mlp * Wd, &B.w2));
    CVB_TRY(W(h, p + "mlp.c_proj.bias", CVB_BF16, Wd, &B.b2));
  }
  CVB_TRY(W(h, t + "ln_final.weight", CVB_BF16, Wd, &s.lnf_w));
  CVB_TRY(W(h, t + "ln_final.bias", CVB_BF16, Wd, &s.lnf_b));
  CVB_TRY(W(h, t + "text_projection.weight", CVB_BF16, (int64_t)Wd * Wd, &s.wproj));
  CVB_TRY(W(h, t + "text_projection.bias", CVB_BF16, Wd, &s.bproj));

  // ---- workspace (trunk)
  const int Tmax = std::max(Np, Tt);
  CVB_TRY(dalloc_t(h, &s.in_image, (size_t)Bm * 3 * c.vf_image * c.vf_image));
  CVB_TRY(dalloc_t(h, &s.in_tokens, (size_t)Bm * Tt));
  CVB_TRY(dalloc_t(h, &s.in_traj, (size_t)Nm * S * c.vf_action_dim));
  // the trunk runs all observations of a batched call at once (rows = observations x tokens); the heads keep one slot
  // per observation so that the pooling chains of every observation run in ONE launch
  CVB_TRY(dalloc_t(h, &s.patches, (size_t)Bm * Np * s.kpad));
  CVB_TRY(dalloc_t(h, &s.hv, (size_t)Bm * Tmax * Wd));
  CVB_TRY(dalloc_t(h, &s.xv, (size_t)Bm * Tmax * Wd));
  CVB_TRY(dalloc_t(h, &s.qkv, (size_t)Bm * Tmax * 3 * Wd));
  CVB_TRY(dalloc_t(h, &s.att, (size_t)Bm * Tmax * Wd));
  CVB_TRY(dalloc_t(h, &s.mlp, (size_t)Bm * Tmax * c.vf_mlp));
  CVB_TRY(dalloc_t(h, &s.pfeat, (size_t)Bm * Np * Wd));
  CVB_TRY(dalloc_t(h, &s.ht, (size_t)Bm * Tt * Wd));
  CVB_TRY(dalloc_t(h, &s.tfeat, (size_t)Bm * Tt * Wd));
  if (Bm > 1) {
    CVB_TRY(dalloc_t(h, &s.pos_tiled, (size_t)Bm * Np * Wd));
    for (int b = 0; b < Bm; ++b)
      CVB_CUDA(cudaMemcpyAsync(s.pos_tiled + (size_t)b * Np * Wd, s.pos_embed, (size_t)Np * Wd * sizeof(bf16),
                               cudaMemcpyDeviceToDevice, st));
  }
  // ---- workspace (heads)
  CVB_TRY(dalloc_t(h, &s.Pn, (size_t)Bm * Np * Wd));
  CVB_TRY(dalloc_t(h, &s.Tn, (size_t)Bm * Tt * Wd));
  CVB_TRY(dalloc_t(h, &s.sim, (size_t)Bm * Tt * Np));
  CVB_TRY(dalloc_t(h, &s.pe, (size_t)Bm * Np * Wd));
  CVB_TRY(dalloc_t(h, &s.taf, (size_t)M * Bm * Tt * Wd));          // [member][observation][token][W]
  CVB_TRY(dalloc_t(h, &s.kv_v, (size_t)M * Bm * Tt * L * 2 * E));  // [member][observation][token][L*2E]
  CVB_TRY(dalloc_t(h, &s.kv_t, (size_t)Bm * Tt * M * L * 2 * E));  // [observation][token][member][L*2E]
  CVB_TRY(dalloc_t(h, &s.vtok, (size_t)Bm * M * E));
  CVB_TRY(dalloc_t(h, &s.ttok, (size_t)Bm * M * E));
  CVB_TRY(dalloc_t(h, &s.it_obs, (size_t)Bm * M * E));  // image-text embeddings of every observation's context
  const size_t rows = (size_t)Nm * S;
  // CVB_TRAJ_SIMT=1 keeps the fp32 SIMT GEMMs (A/B testing); the tensor-core path needs 8-element aligned widths
  s.traj_tc = getenv("CVB_TRAJ_SIMT") == nullptr && E % 8 == 0 && c.vf_traj_ff % 8 == 0 && c.vf_traj_layers > 0;
  if (s.traj_tc) {
    CVB_TRY(dalloc_t(h, &s.txs, rows * 3 * E * M));
    CVB_TRY(dalloc_t(h, &s.tatts, rows * 3 * E * M));
    CVB_TRY(dalloc_t(h, &s.tffs, rows * 3 * c.vf_traj_ff * M));
  }
  CVB_TRY(dalloc_t(h, &s.tx, rows * E * M));
  CVB_TRY(dalloc_t(h, &s.tqkv, rows * 3 * E * M));
  CVB_TRY(dalloc_t(h, &s.tatt, rows * E * M));
  CVB_TRY(dalloc_t(h, &s.ty, rows * E * M));
  CVB_TRY(dalloc_t(h, &s.tff, rows * c.vf_traj_ff * M));
  s.mstream.assign(M, nullptr);
  s.mevent.assign(M, nullptr);
  for (int m = 1; m < M; ++m) {
    CVB_CUDA(cudaStreamCreateWithFlags(&s.mstream[m], cudaStreamNonBlocking));
    CVB_CUDA(cudaEventCreateWithFlags(&s.mevent[m], cudaEventDisableTiming));
  }
  CVB_CUDA(cudaEventCreateWithFlags(&s.ev_fork, cudaEventDisableTiming));
  CVB_TRY(dalloc_t(h, &s.act, (size_t)M * Nm * E));
  CVB_TRY(dalloc_t(h, &s.scores, Nm));
  CVB_TRY(dalloc_t(h, &s.gmean, Nm));
  CVB_TRY(dalloc_t(h, &s.bidx, Bm));
  CVB_TRY(dalloc_t(h, &s.bscore, Bm));

  // ---- heads: pack K/V projections of all pooling blocks (they only depend on the kv input)
  s.mem.resize(M);
  CVB_TRY(dalloc_t(h, &s.wkv_t, (size_t)M * L * 2 * E * Wd));
  CVB_TRY(dalloc_t(h, &s.bkv_t, (size_t)M * L * 2 * E));
  std::vector<PoolChain> chains((size_t)Bm * 2 * M);  // [observation][member][vision, text]; slot 0 filled first
  std::vector<ItFinal> itf((size_t)Bm * M);
  for (int m = 0; m < M; ++m) {
    const std::string b = "verifier." + std::to_string(m) + ".";
    MemberW& Mw = s.mem[m];
    CVB_TRY(W(h, b + "text_aware_visual_extraction.temperature", CVB_F32, 1, &Mw.temp));
    CVB_TRY(W(h, b + "text_aware_visual_extraction.pos_emb", CVB_F32, (int64_t)Np * Wd, &Mw.pos_emb));
    if (c.vf_traj_layers == 0) {
      const std::string q = b + "complex_action_encoder.";
      const int FFh = c.vf_traj_ff;
      CVB_TRY(W(h, q + "0.weight", CVB_F32, (int64_t)FFh * c.vf_history * c.vf_action_dim, &Mw.mlp_w0));
      CVB_TRY(W(h, q + "0.bias", CVB_F32, FFh, &Mw.mlp_b0));
      CVB_TRY(W(h, q + "1.weight", CVB_F32, FFh, &Mw.mlp_lnw));
      CVB_TRY(W(h, q + "1.bias", CVB_F32, FFh, &Mw.mlp_lnb));
      CVB_TRY(W(h, q + "4.weight", CVB_F32, (int64_t)E * FFh, &Mw.mlp_w1));
      CVB_TRY(W(h, q + "4.bias", CVB_F32, E, &Mw.mlp_b1));
    } else {
      CVB_TRY(W(h, b + "single_step_action_encoder.weight", CVB_F32, (int64_t)E * c.vf_action_dim, &Mw.w_ss));
      CVB_TRY(W(h, b + "single_step_action_encoder.bias", CVB_F32, E, &Mw.b_ss));
    }
    CVB_TRY(dalloc_t(h, &Mw.wkv_v, (size_t)L * 2 * E * Wd));
    CVB_TRY(dalloc_t(h, &Mw.bkv_v, (size_t)L * 2 * E));
    for (int pi = 0; pi < 2; ++pi) {
      const std::string p = b + (pi == 0 ? "vision_poolings." : "text_pooling.");
      PoolChain& ch = chains[m * 2 + pi];
      memset(&ch, 0, sizeof(ch));
      CVB_TRY(W(h, p + "query", CVB_F32, E, &ch.query));
      CVB_TRY(W(h, p + "layer_norm.weight", CVB_F32, E, &ch.fin_w));
      CVB_TRY(W(h, p + "layer_norm.bias", CVB_F32, E, &ch.fin_b));
      ch.embed = E, ch.heads = c.vf_pool_heads, ch.tokens = Tt, ch.layers = L;
      float* wdst = pi == 0 ? Mw.wkv_v : s.wkv_t + (size_t)m * L * 2 * E * Wd;
      float* bdst = pi == 0 ? Mw.bkv_v : s.bkv_t + (size_t)m * L * 2 * E;
      if (pi == 0) {
        ch.kv = s.kv_v + (size_t)m * Bm * Tt * L * 2 * E, ch.kv_ld = L * 2 * E, ch.out = s.vtok + (size_t)m * E;
      } else {
        ch.kv = s.kv_t + (size_t)m * L * 2 * E, ch.kv_ld = M * L * 2 * E, ch.out = s.ttok + (size_t)m * E;
      }
      for (int i = 0; i < L; ++i) {
        const std::string q = p + "blocks." + std::to_string(i) + ".";
        PoolBlockW& bw = ch.blk[i];
        const float *wk, *wv;
        CVB_TRY(W(h, q + "attention.q_proj_weight", CVB_F32, (int64_t)E * E, &bw.wq));
        CVB_TRY(W(h, q + "attention.k_proj_weight", CVB_F32, (int64_t)E * Wd, &wk));
        CVB_TRY(W(h, q + "attention.v_proj_weight", CVB_F32, (int64_t)E * Wd, &wv));
        CVB_TRY(W(h, q + "attention.in_proj_bias", CVB_F32, 3 * E, &bw.b_in));
        CVB_TRY(W(h, q + "attention.out_proj.weight", CVB_F32, (int64_t)E * E, &bw.wo));
        CVB_TRY(W(h, q + "attention.out_proj.bias", CVB_F32, E, &bw.bo));
        CVB_TRY(W(h, q + "mlp.fc1.weight", CVB_F32, (int64_t)E * E, &bw.fc1_w));
        CVB_TRY(W(h, q + "mlp.fc1.bias", CVB_F32, E, &bw.fc1_b));
        CVB_TRY(W(h, q + "mlp.fc2.weight", CVB_F32, (int64_t)E * E, &bw.fc2_w));
        CVB_TRY(W(h, q + "mlp.fc2.bias", CVB_F32, E, &bw.fc2_b));
        CVB_TRY(W(h, q + "q_layer_norm.weight", CVB_F32, E, &bw.qln_w));
        CVB_TRY(W(h, q + "q_layer_norm.bias", CVB_F32, E, &bw.qln_b));
        CVB_TRY(W(h, q + "layer_norm.weight", CVB_F32, E, &bw.ln_w));
        CVB_TRY(W(h, q + "layer_norm.bias", CVB_F32, E, &bw.ln_b));
        CVB_CUDA(cudaMemcpyAsync(wdst + (size_t)(i * 2) * E * Wd, wk, (size_t)E * Wd * sizeof(float),
                                 cudaMemcpyDeviceToDevice, st));
        CVB_CUDA(cudaMemcpyAsync(wdst + (size_t)(i * 2 + 1) * E * Wd, wv, (size_t)E * Wd * sizeof(float),
                                 cudaMemcpyDeviceToDevice, st));
        CVB_CUDA(cudaMemcpyAsync(bdst + (size_t)(i * 2) * E, bw.b_in + E, (size_t)2 * E * sizeof(float),
                                 cudaMemcpyDeviceToDevice, st));
      }
    }
    ItFinal& f = itf[m];
    f.text_tok = s.ttok + (size_t)m * E, f.vision_tok = s.vtok + (size_t)m * E, f.out = s.it_obs + (size_t)m * E;
    CVB_TRY(W(h, b + "input_projection.weight", CVB_F32, (int64_t)E * 2 * E, &f.w));
    CVB_TRY(W(h, b + "input_projection.bias", CVB_F32, E, &f.b));
    Mw.traj.resize(c.vf_traj_layers);
    for (int i = 0; i < c.vf_traj_layers; ++i) {
      const std::string q = b + "trajectory_encoder.layers." + std::to_string(i) + ".";
      TrajLayer& T = Mw.traj[i];
      CVB_TRY(W(h, q + "self_attn.in_proj_weight", CVB_F32, (int64_t)3 * E * E, &T.w_in));
      CVB_TRY(W(h, q + "self_attn.in_proj_bias", CVB_F32, 3 * E, &T.b_in));
      CVB_TRY(W(h, q + "self_attn.out_proj.weight", CVB_F32, (int64_t)E * E, &T.wo));
      CVB_TRY(W(h, q + "self_attn.out_proj.bias", CVB_F32, E, &T.bo));
      CVB_TRY(W(h, q + "linear1.weight", CVB_F32, (int64_t)c.vf_traj_ff * E, &T.w1));
      CVB_TRY(W(h, q + "linear1.bias", CVB_F32, c.vf_traj_ff, &T.b1));
      CVB_TRY(W(h, q + "linear2.weight", CVB_F32, (int64_t)c.vf_traj_ff * E, &T.w2));
      CVB_TRY(W(h, q + "linear2.bias", CVB_F32, E, &T.b2));
      CVB_TRY(W(h, q + "norm1.weight", CVB_F32, E, &T.n1w));
      CVB_TRY(W(h, q + "norm1.bias", CVB_F32, E, &T.n1b));
      CVB_TRY(W(h, q + "norm2.weight", CVB_F32, E, &T.n2w));
      CVB_TRY(W(h, q + "norm2.bias", CVB_F32, E, &T.n2b));
      if (s.traj_tc) {
        const int FF = c.vf_traj_ff;
        CVB_TRY(dalloc_t(h, &T.w_in3, (size_t)3 * E * 3 * E));
        CVB_TRY(dalloc_t(h, &T.wo3, (size_t)E * 3 * E));
        CVB_TRY(dalloc_t(h, &T.w13, (size_t)FF * 3 * E));
        CVB_TRY(dalloc_t(h, &T.w23, (size_t)E * 3 * FF));
        CVB_TRY(split3_rows(st, T.w_in, E, T.w_in3, 3 * E, E, 1));
        CVB_TRY(split3_rows(st, T.wo, E, T.wo3, E, E, 1));
        CVB_TRY(split3_rows(st, T.w1, E, T.w13, FF, E, 1));
        CVB_TRY(split3_rows(st, T.w2, FF, T.w23, E, FF, 1));
      }
    }
  }
  for (int o = 1; o < Bm; ++o) {  // the other observation slots: same weights, their own K/V, token and output rows
    for (int m = 0; m < M; ++m) {
      PoolChain& cv = chains[((size_t)o * M + m) * 2];
      PoolChain& ct = chains[((size_t)o * M + m) * 2 + 1];
      cv = chains[m * 2], ct = chains[m * 2 + 1];
      cv.kv = s.kv_v + ((size_t)m * Bm + o) * Tt * L * 2 * E, cv.out = s.vtok + ((size_t)o * M + m) * E;
      ct.kv = s.kv_t + (size_t)o * Tt * M * L * 2 * E + (size_t)m * L * 2 * E, ct.out = s.ttok + ((size_t)o * M + m) * E;
      ItFinal& f = itf[(size_t)o * M + m];
      f = itf[m];
      f.text_tok = s.ttok + ((size_t)o * M + m) * E, f.vision_tok = s.vtok + ((size_t)o * M + m) * E;
      f.out = s.it_obs + ((size_t)o * M + m) * E;
    }
  }
  CVB_TRY(dalloc_t(h, &s.chains, chains.size()));
  CVB_TRY(dalloc_t(h, &s.itf, itf.size()));
  CVB_CUDA(cudaMemcpyAsync(s.chains, chains.data(), chains.size() * sizeof(PoolChain), cudaMemcpyHostToDevice, st));
  CVB_CUDA(cudaMemcpyAsync(s.itf, itf.data(), itf.size() * sizeof(ItFinal), cudaMemcpyHostToDevice, st));
  CVB_CUDA(cudaStreamSynchronize(st));
  return 0;
}

void verifier_destroy(cvb_handle* h) {
  if (h->vf != nullptr) {
    h->vf->ctx_graph.destroy();
    h->vf->traj_graph.destroy();
    for (cudaStream_t st : h->vf->mstream)
      if (st) cudaStreamDestroy(st);
    for (cudaEvent_t e : h->vf->mevent)
      if (e) cudaEventDestroy(e);
    if (h->vf->ev_fork) cudaEventDestroy(h->vf->ev_fork);
  }
  delete h->vf;
  h->vf = nullptr;
}

// pre-norm transformer block stack on bf16 rows; `stop_after_attn_proj` implements the hook on
// visual.trunk.blocks[-1].attn (the last block's attention output BEFORE the residual).
static int run_blocks(cudaStream_t st, VerifierState& s, const std::vector<TrunkBlock>& blocks, bf16* hbuf, int Tseq,
                      int Wd, int heads, int mlp, bool last_is_attn_only, bf16* attn_only_out, int nb = 1) {
  const int hd = Wd / heads, T = nb * Tseq;  // T rows through the linear layers, nb attention batches of Tseq tokens
  for (size_t l = 0; l < blocks.size(); ++l) {
    const TrunkBlock& B = blocks[l];
    const bool last = last_is_attn_only && l + 1 == blocks.size();
    CVB_TRY(layernorm_bf16(st, hbuf, Wd, B.ln1_w, B.ln1_b, s.xv, Wd, T, Wd, 1e-6f));
    CVB_TRY(gemm(st, s.xv, Wd, B.wqkv, Wd, T, 3 * Wd, Wd, EPI_STORE, s.qkv, 3 * Wd, B.bqkv));
    AttnCall a;
    a.q = s.qkv, a.q_row_stride = 3 * Wd;
    a.k0 = s.qkv + Wd, a.v0 = s.qkv + 2 * Wd, a.kv0_row_stride = 3 * Wd, a.kv0_len = Tseq;
    a.out = s.att, a.o_row_stride = Wd;
    a.q_batch_stride = (long)Tseq * 3 * Wd, a.kv0_batch_stride = (long)Tseq * 3 * Wd, a.o_batch_stride = (long)Tseq * Wd;
    a.batches = nb, a.heads = heads, a.kv_heads = heads, a.tq = Tseq, a.head_dim = hd;
    a.scale = 1.0f / sqrtf(static_cast<float>(hd));
    CVB_TRY(attention(st, a));
    if (last) {
      CVB_TRY(gemm(st, s.att, Wd, B.wo, Wd, T, Wd, Wd, EPI_STORE, attn_only_out, Wd, B.bo));
      break;
    }
    CVB_TRY(gemm(st, s.att, Wd, B.wo, Wd, T, Wd, Wd, EPI_RESID, hbuf, Wd, B.bo, hbuf, Wd));
    CVB_TRY(layernorm_bf16(st, hbuf, Wd, B.ln2_w, B.ln2_b, s.xv, Wd, T, Wd, 1e-6f));
    CVB_TRY(gemm(st, s.xv, Wd, B.w1, Wd, T, mlp, Wd, EPI_GELU, s.mlp, mlp, B.b1));
    CVB_TRY(gemm(st, s.mlp, mlp, B.w2, mlp, T, Wd, mlp, EPI_RESID, hbuf, Wd, B.b2, hbuf, Wd));
  }
  return 0;
}

// image-text heads of every member (N-independent: SURVEY.md F4) from the normalised features Pn / Tn of observations
// [obs0, obs0 + nb) (rows of Pn / Tn are RELATIVE to obs0; every other buffer is indexed by the absolute slot)
static int run_heads_context(cvb_handle* h, cudaStream_t st, int obs0 = 0, int nb = 1) {
  const cvb_config& c = h->cfg;
  VerifierState& s = *h->vf;
  const int Wd = c.vf_width, E = c.vf_embed, L = c.vf_pool_layers, M = c.vf_members, Bm = h->max_obs();
  const int Np = (c.vf_image / c.vf_patch) * (c.vf_image / c.vf_patch), Tt = c.vf_text_ctx;
  const int L2E = L * 2 * E;
  // text pooling K/V of every member and observation: one GEMM over nb * Tt rows
  CVB_TRY(sg(st, s.Tn, Wd, s.wkv_t, Wd, nb * Tt, M * L2E, Wd, s.kv_t + (size_t)obs0 * Tt * M * L2E, M * L2E, s.bkv_t));
  for (int m = 0; m < M; ++m) {
    const MemberW& Mw = s.mem[m];
    float* taf = s.taf + ((size_t)m * Bm + obs0) * Tt * Wd;
    for (int o = 0; o < nb; ++o) {
      const float* Pn = s.Pn + (size_t)o * Np * Wd;
      const float* Tn = s.Tn + (size_t)o * Tt * Wd;
      float* sim = s.sim + (size_t)o * Tt * Np;
      float* pe = s.pe + (size_t)o * Np * Wd;
      CVB_TRY(sg(st, Tn, Wd, Pn, Wd, Tt, Np, Wd, sim, Np, nullptr, 0, nullptr, 0, 0, /*w_dynamic=*/1));
      CVB_TRY(softmax_rows_temp(st, sim, Tt, Np, Mw.temp));
      CVB_TRY(add_f32(st, Pn, Mw.pos_emb, pe, (long)Np * Wd));
      CVB_TRY(sg(st, sim, Np, pe, Wd, Tt, Wd, Np, taf + (size_t)o * Tt * Wd, Wd, nullptr, 0, nullptr, 0, /*w_kn=*/1));
    }
    // vision pooling K/V of this member for every observation: one GEMM over nb * Tt rows
    CVB_TRY(sg(st, taf, Wd, Mw.wkv_v, Wd, nb * Tt, L2E, Wd, s.kv_v + ((size_t)m * Bm + obs0) * Tt * L2E, L2E, Mw.bkv_v));
  }
  CVB_TRY(pool_chains(st, s.chains + (size_t)obs0 * 2 * M, nb * 2 * M, E, c.vf_pool_heads, Tt));
  CVB_TRY(it_finalize(st, s.itf + (size_t)obs0 * M, nb * M, E));  // writes it_obs[obs0 .. obs0 + nb)
  return 0;
}

// image / text towers + heads of observations [obs0, obs0 + nb) (inputs staged by verifier_stage_context_inputs)
static int run_context(cvb_handle* h, cudaStream_t st, int obs0 = 0, int nb = 1) {
  const cvb_config& c = h->cfg;
  VerifierState& s = *h->vf;
  tl_batch_handle = h->max_obs() > 1;
  const int Wd = c.vf_width;
  const int Np = (c.vf_image / c.vf_patch) * (c.vf_image / c.vf_patch), Tt = c.vf_text_ctx;
  const float* in_image = s.in_image + (size_t)obs0 * 3 * c.vf_image * c.vf_image;
  const int64_t* in_tokens = s.in_tokens + (size_t)obs0 * Tt;
  // image tower -> patch features (hook output), text tower -> per-token projected features
  CVB_TRY(im2col_patches(st, in_image, s.patches, 3, c.vf_image, c.vf_image, c.vf_patch, s.kpad, nb));
  CVB_TRY(gemm(st, s.patches, s.kpad, s.w_patch, s.kpad, nb * Np, Wd, s.kpad, EPI_RESID, s.hv, Wd, s.patch_b,
               nb > 1 ? s.pos_tiled : s.pos_embed, Wd));
  CVB_TRY(run_blocks(st, s, s.vis, s.hv, Np, Wd, c.vf_heads, c.vf_mlp, true, s.pfeat, nb));
  CVB_TRY(l2norm_rows_bf16_to_f32(st, s.pfeat, Wd, s.Pn, nb * Np, Wd));
  if (!(s.text_hold && s.text_nb == nb)) {  // (held: the caller vouches that the instructions did not change - Tn is resident)
    CVB_TRY(embed_tokens_pos(st, s.tok_emb, s.txt_pos, in_tokens, s.ht, nb * Tt, Wd, Tt));
    CVB_TRY(run_blocks(st, s, s.txt, s.ht, Tt, Wd, c.vf_heads, c.vf_mlp, false, nullptr, nb));
    CVB_TRY(layernorm_bf16(st, s.ht, Wd, s.lnf_w, s.lnf_b, s.xv, Wd, nb * Tt, Wd, 1e-6f));
    CVB_TRY(gemm(st, s.xv, Wd, s.wproj, Wd, nb * Tt, Wd, Wd, EPI_STORE, s.tfeat, Wd, s.bproj));
    CVB_TRY(l2norm_rows_bf16_to_f32(st, s.tfeat, Wd, s.Tn, nb * Tt, Wd));
  }
  CVB_TRY(run_heads_context(h, st, obs0, nb));
  return 0;
}

static int run_member_trajectories(cvb_handle* h, cudaStream_t st, int N, int m) {
  const cvb_config& c = h->cfg;
  VerifierState& s = *h->vf;
  const int E = c.vf_embed, S = c.vf_history, A = c.vf_action_dim, FF = c.vf_traj_ff;
  const int rows = N * S;
  const size_t cap = (size_t)h->rm_total() * c.max_samples * S;  // rows of workspace per member
  float* tx = s.tx + (size_t)m * cap * E;
  float* tqkv = s.tqkv + (size_t)m * cap * 3 * E;
  float* tatt = s.tatt + (size_t)m * cap * E;
  float* ty = s.ty + (size_t)m * cap * E;
  float* tff = s.tff + (size_t)m * cap * FF;
  const MemberW& Mw = s.mem[m];
  if (c.vf_traj_layers == 0) {
    // MLP action encoder: flat history [N, S*A] -> Linear -> LayerNorm -> ReLU -> Linear -> unit norm
    // (efficient_ensemble_merged.py:241-245; Dropout is the identity in eval mode)
    CVB_TRY(sg(st, s.in_traj, (long)S * A, Mw.mlp_w0, (long)S * A, N, FF, S * A, tff, FF, Mw.mlp_b0));
    CVB_TRY(layernorm_f32(st, tff, nullptr, Mw.mlp_lnw, Mw.mlp_lnb, tff, N, FF, 1e-5f, nullptr, /*relu=*/1));
    CVB_TRY(sg(st, tff, FF, Mw.mlp_w1, FF, N, E, FF, ty, E, Mw.mlp_b1));
    return l2norm_rows_f32(st, ty, s.act + (size_t)m * N * E, N, E);
  }
  CVB_TRY(sg(st, s.in_traj, A, Mw.w_ss, A, rows, E, A, tx, E, Mw.b_ss));
  if (s.traj_tc) {
    // every linear layer as ONE tcgen05 GEMM over the 3K axis of the [hi | hi | lo] x [hi | lo | hi] operands (fp32
    // accumulate, fp32 bias, fp32 store): same post-norm layer, same fp32 attention / LayerNorm kernels in between
    bf16* txs = s.txs + (size_t)m * cap * 3 * E;
    bf16* tatts = s.tatts + (size_t)m * cap * 3 * E;
    bf16* tffs = s.tffs + (size_t)m * cap * 3 * FF;
    auto tc = [&](const bf16* Ap, const bf16* Wp, int Nn, int Kk, float* Cp, const float* bias) {
      GemmCall g;
      g.A = Ap, g.lda = 3 * Kk, g.W = Wp, g.ldw = 3 * Kk, g.M = rows, g.N = Nn, g.K = 3 * Kk, g.epi = EPI_F32;
      g.C = Cp, g.ldc = Nn, g.bias = bias, g.bias_is_f32 = 1;
      g.no_skinny = h->max_obs() > 1 ? 1 : 0;
      return gemm_bf16(st, g);
    };
    CVB_TRY(split3_rows(st, tx, E, txs, rows, E, 0));
    for (const TrajLayer& T : Mw.traj) {
      CVB_TRY(tc(txs, T.w_in3, 3 * E, E, tqkv, T.b_in));
      CVB_TRY(traj_attention(st, tqkv, s.in_traj, tatt, N, S, E, c.vf_pool_heads, A, -5.0f));
      CVB_TRY(split3_rows(st, tatt, E, tatts, rows, E, 0));
      CVB_TRY(tc(tatts, T.wo3, E, E, ty, T.bo));
      CVB_TRY(layernorm_f32(st, ty, tx, T.n1w, T.n1b, tx, rows, E, 1e-5f, txs));
      CVB_TRY(tc(txs, T.w13, FF, E, tff, T.b1));
      CVB_TRY(split3_rows(st, tff, FF, tffs, rows, FF, 0, /*relu=*/1));
      CVB_TRY(tc(tffs, T.w23, E, FF, ty, T.b2));
      CVB_TRY(layernorm_f32(st, ty, tx, T.n2w, T.n2b, tx, rows, E, 1e-5f, txs));
    }
    return masked_mean_l2norm(st, tx, s.in_traj, s.act + (size_t)m * N * E, N, S, E, A, -5.0f);
  }
  for (const TrajLayer& T : Mw.traj) {
    CVB_TRY(sg(st, tx, E, T.w_in, E, rows, 3 * E, E, tqkv, 3 * E, T.b_in));
    CVB_TRY(traj_attention(st, tqkv, s.in_traj, tatt, N, S, E, c.vf_pool_heads, A, -5.0f));
    CVB_TRY(sg(st, tatt, E, T.wo, E, rows, E, E, ty, E, T.bo));
    CVB_TRY(layernorm_f32(st, ty, tx, T.n1w, T.n1b, tx, rows, E, 1e-5f));
    CVB_TRY(sg(st, tx, E, T.w1, E, rows, FF, E, tff, FF, T.b1, SACT_RELU));
    CVB_TRY(sg(st, tff, FF, T.w2, FF, rows, E, FF, ty, E, T.b2));
    CVB_TRY(layernorm_f32(st, ty, tx, T.n2w, T.n2b, tx, rows, E, 1e-5f));
  }
  return masked_mean_l2norm(st, tx, s.in_traj, s.act + (size_t)m * N * E, N, S, E, A, -5.0f);
}

// One branch per ensemble member (forked from / joined back into `st`; plain parallel graph branches under capture).
static int run_trajectories(cvb_handle* h, cudaStream_t st, int N) {
  VerifierState& s = *h->vf;
  const int M = h->cfg.vf_members;
  if (M > 1) CVB_CUDA(cudaEventRecord(s.ev_fork, st));
  int rc = 0;
  for (int m = 1; m < M; ++m) {
    CVB_CUDA(cudaStreamWaitEvent(s.mstream[m], s.ev_fork, 0));
    if (rc == 0) rc = run_member_trajectories(h, s.mstream[m], N, m);
    CVB_CUDA(cudaEventRecord(s.mevent[m], s.mstream[m]));  // always rejoin so a stream capture can end cleanly
  }
  if (rc == 0) rc = run_member_trajectories(h, st, N, 0);
  for (int m = 1; m < M; ++m) CVB_CUDA(cudaStreamWaitEvent(st, s.mevent[m], 0));
  return rc;
}

int verifier_score(cvb_handle* h, const float* image, const int64_t* tokens, const float* traj, int N, int R, int K,
                   float* scores, float* group_mean, int32_t* best_idx, float* best_score, int recompute_context,
                   cudaStream_t st) {
  const cvb_config& c = h->cfg;
  CVB_REQUIRE(h->finalized && h->vf != nullptr, "verifier not configured (vf_members == 0?) or not finalized");
  VerifierState& s = *h->vf;
  CVB_REQUIRE(N >= 1 && N <= c.max_rephrases * c.max_samples, "N out of range");
  CVB_REQUIRE(R == 0 || R * K == N, "R*K must equal N");
  if (recompute_context || !s.context_valid) {
    CVB_REQUIRE(image != nullptr && tokens != nullptr, "image / tokens required to compute the context");
    CVB_CUDA(cudaMemcpyAsync(s.in_image, image, (size_t)3 * c.vf_image * c.vf_image * sizeof(float),
                             cudaMemcpyDeviceToDevice, st));
    CVB_CUDA(cudaMemcpyAsync(s.in_tokens, tokens, c.vf_text_ctx * sizeof(int64_t), cudaMemcpyDeviceToDevice, st));
    CVB_TRY(s.ctx_graph.run(c.use_cuda_graph != 0, verifier_text_cached(h, 1) ? 1 : 0, st,
                            [&](cudaStream_t cs) { return run_context(h, cs); }));
    verifier_note_context(h, 1);
  }
  CVB_CUDA(cudaMemcpyAsync(s.in_traj, traj, (size_t)N * c.vf_history * c.vf_action_dim * sizeof(float),
                           cudaMemcpyDeviceToDevice, st));
  const long key = ((long)N << 32) | ((long)R << 16) | (long)K;
  CVB_TRY(s.traj_graph.run(c.use_cuda_graph != 0, key, st, [&](cudaStream_t cs) {
    CVB_TRY(run_trajectories(h, cs, N));
    return fuse_score_select(cs, s.it_obs, s.act, c.vf_members, N, c.vf_embed, s.scores, R, K, s.gmean, s.bidx, s.bscore,
                             R > 0 ? 1 : 0);
  }));
  CVB_CUDA(cudaMemcpyAsync(scores, s.scores, N * sizeof(float), cudaMemcpyDeviceToDevice, st));
  if (R > 0) {
    if (group_mean != nullptr)
      CVB_CUDA(cudaMemcpyAsync(group_mean, s.gmean, R * sizeof(float), cudaMemcpyDeviceToDevice, st));
    CVB_CUDA(cudaMemcpyAsync(best_idx, s.bidx, sizeof(int32_t), cudaMemcpyDeviceToDevice, st));
    CVB_CUDA(cudaMemcpyAsync(best_score, s.bscore, sizeof(float), cudaMemcpyDeviceToDevice, st));
  }
  return 0;
}

int verifier_stage_context_inputs(cvb_handle* h, const float* image, const int64_t* tokens, cudaStream_t st, int B) {
  const cvb_config& c = h->cfg;
  CVB_REQUIRE(h->finalized && h->vf != nullptr, "verifier not configured (vf_members == 0?) or not finalized");
  CVB_REQUIRE(image != nullptr && tokens != nullptr, "image / tokens required");
  CVB_REQUIRE(B >= 1 && B <= h->max_obs(), "number of observations out of range (max_observations)");
  VerifierState& s = *h->vf;
  CVB_CUDA(cudaMemcpyAsync(s.in_image, image, (size_t)B * 3 * c.vf_image * c.vf_image * sizeof(float),
                           cudaMemcpyDeviceToDevice, st));
  CVB_CUDA(cudaMemcpyAsync(s.in_tokens, tokens, (size_t)B * c.vf_text_ctx * sizeof(int64_t), cudaMemcpyDeviceToDevice, st));
  return 0;
}

// image/text side of observation `obs` (its inputs were staged by verifier_stage_context_inputs)
int verifier_enqueue_context(cvb_handle* h, cudaStream_t st, int obs, int nb) {
  CVB_REQUIRE(obs >= 0 && nb >= 1 && obs + nb <= h->max_obs(), "observation slots out of range (max_observations)");
  CVB_TRY(run_context(h, st, obs, nb));
  h->vf->context_valid = true;
  return 0;
}

// N candidates per observation, B observations: one pass of the trajectory encoders over all B * N histories, then one
// fused score / select CTA group per observation against ITS context
int verifier_enqueue_score(cvb_handle* h, cudaStream_t st, int N, int R, int K, int B) {
  const cvb_config& c = h->cfg;
  VerifierState& s = *h->vf;
  CVB_REQUIRE(N >= 1 && N <= c.max_rephrases * c.max_samples, "N out of range");
  CVB_REQUIRE(R == 0 || R * K == N, "R*K must equal N");
  CVB_REQUIRE(B >= 1 && B <= h->max_obs(), "number of observations out of range");
  CVB_TRY(run_trajectories(h, st, B * N));
  return fuse_score_select(st, s.it_obs, s.act, c.vf_members, N, c.vf_embed, s.scores, R, K, s.gmean, s.bidx, s.bscore,
                           R > 0 ? 1 : 0, B);
}

float* verifier_traj_buffer(cvb_handle* h) { return h->vf->in_traj; }

// true when a context call for nb observations will skip the text tower (part of the callers' graph keys)
bool verifier_text_cached(cvb_handle* h, int nb) { return h->vf != nullptr && h->vf->text_hold && h->vf->text_nb == nb; }
// Host-side bookkeeping of every context computation (called OUTSIDE the graph lambdas: a replayed graph does not run the
// host code of run_context): the resident text features now belong to `nb` observations.
void verifier_note_context(cvb_handle* h, int nb) {
  h->vf->context_valid = true;
  h->vf->text_nb = nb;
}
int verifier_hold_text(cvb_handle* h, int hold) {
  CVB_REQUIRE(h->vf != nullptr, "verifier not configured");
  h->vf->text_hold = hold != 0;
  return 0;
}

int verifier_copy_results(cvb_handle* h, int N, int R, float* scores, float* group_mean, int32_t* best_idx,
                          float* best_score, cudaStream_t st, int B) {
  VerifierState& s = *h->vf;
  if (scores != nullptr)
    CVB_CUDA(cudaMemcpyAsync(scores, s.scores, (size_t)B * N * sizeof(float), cudaMemcpyDeviceToDevice, st));
  if (R > 0) {
    if (group_mean != nullptr)
      CVB_CUDA(cudaMemcpyAsync(group_mean, s.gmean, (size_t)B * R * sizeof(float), cudaMemcpyDeviceToDevice, st));
    if (best_idx != nullptr)
      CVB_CUDA(cudaMemcpyAsync(best_idx, s.bidx, B * sizeof(int32_t), cudaMemcpyDeviceToDevice, st));
    if (best_score != nullptr)
      CVB_CUDA(cudaMemcpyAsync(best_score, s.bscore, B * sizeof(float), cudaMemcpyDeviceToDevice, st));
  }
  return 0;
}

int verifier_context(cvb_handle* h, const float* image, const int64_t* tokens, cudaStream_t st) {
  const cvb_config& c = h->cfg;
  CVB_REQUIRE(h->finalized && h->vf != nullptr, "verifier not configured (vf_members == 0?) or not finalized");
  CVB_REQUIRE(image != nullptr && tokens != nullptr, "image / tokens required");
  VerifierState& s = *h->vf;
  CVB_CUDA(cudaMemcpyAsync(s.in_image, image, (size_t)3 * c.vf_image * c.vf_image * sizeof(float),
                           cudaMemcpyDeviceToDevice, st));
  CVB_CUDA(cudaMemcpyAsync(s.in_tokens, tokens, c.vf_text_ctx * sizeof(int64_t), cudaMemcpyDeviceToDevice, st));
  CVB_TRY(s.ctx_graph.run(c.use_cuda_graph != 0, verifier_text_cached(h, 1) ? 1 : 0, st,
                          [&](cudaStream_t cs) { return run_context(h, cs); }));
  verifier_note_context(h, 1);
  return 0;
}

int64_t verifier_debug_copy(cvb_handle* h, const std::string& name, void* dst, int64_t max_bytes, cudaStream_t st) {
  if (h->vf == nullptr) return -1;
  const cvb_config& c = h->cfg;
  VerifierState& s = *h->vf;
  const int Np = (c.vf_image / c.vf_patch) * (c.vf_image / c.vf_patch);
  const void* src = nullptr;
  int64_t bytes = 0;
  if (name == "vf_patch_features") {
    src = s.Pn, bytes = (int64_t)Np * c.vf_width * 4;
  } else if (name == "vf_text_features") {
    src = s.Tn, bytes = (int64_t)c.vf_text_ctx * c.vf_width * 4;
  } else if (name == "vf_it_emb") {
    src = s.it_obs, bytes = (int64_t)c.vf_members * c.vf_embed * 4;
  } else if (name == "vf_act_emb") {
    src = s.act, bytes = (int64_t)c.vf_members * h->rm_total() * c.max_samples * c.vf_embed * 4;
  } else {
    return -1;
  }
  if (bytes > max_bytes) bytes = max_bytes;
  if (cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, st) != cudaSuccess) return -2;
  return bytes;
}

// test hook: overwrite the normalised trunk features (lets tests check the heads in isolation)
int verifier_set_features(cvb_handle* h, const float* patch, const float* text, cudaStream_t st) {
  CVB_REQUIRE(h->vf != nullptr, "verifier not configured");
  const cvb_config& c = h->cfg;
  VerifierState& s = *h->vf;
  const int Wd = c.vf_width;
  const int Np = (c.vf_image / c.vf_patch) * (c.vf_image / c.vf_patch), Tt = c.vf_text_ctx;
  CVB_CUDA(cudaMemcpyAsync(s.Pn, patch, (size_t)Np * Wd * 4, cudaMemcpyDeviceToDevice, st));
  CVB_CUDA(cudaMemcpyAsync(s.Tn, text, (size_t)Tt * Wd * 4, cudaMemcpyDeviceToDevice, st));
  CVB_TRY(run_heads_context(h, st, 0, 1));
  s.context_valid = true;
  return 0;
}

}  // namespace cvb
