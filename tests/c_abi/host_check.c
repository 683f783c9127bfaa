/* Plain-C host of the coverb200 C ABI (no CUDA call is made): proves that include/coverb200.h is valid C, that the
 * library binds from C without torch, and exercises the parts of the boundary that need no device - version, error
 * reporting, configuration checks, the weight manifest a loader walks, the host-side constants of
 * the denoise loop (modeling_pi0.py:697-715).  Built and run by tests/test_c_abi_from_c.py. */
#include <stdio.h>
#include <string.h>

#include "coverb200.h"

#define CHECK(cond, ...)                      \
  do {                                        \
    if (!(cond)) {                            \
      fprintf(stderr, "FAIL %s:%d: ", __FILE__, __LINE__); \
      fprintf(stderr, __VA_ARGS__);           \
      fprintf(stderr, "\n");                  \
      return 1;                               \
    }                                         \
  } while (0)

static void bridge_config(cvb_config* c) {
  /* INT-ACT/config/models/pi0_finetune_bridge.json + the CoVer-BridgeV2 verifier (3 members) */
  memset(c, 0, sizeof(*c));
  c->struct_size = (int32_t)sizeof(*c);
  c->vis_layers = 27; c->vis_width = 1152; c->vis_heads = 16; c->vis_mlp = 4304; c->vis_patch = 14; c->vis_image = 224;
  c->layers = 18; c->lm_width = 2048; c->lm_mlp = 16384; c->heads = 8; c->head_dim = 256; c->ex_width = 1024;
  c->ex_mlp = 4096; c->vocab = 257152;
  c->max_state_dim = 32; c->max_action_dim = 32; c->chunk_size = 4; c->max_lang_len = 72; c->num_steps = 10;
  c->max_rephrases = 8; c->max_samples = 5;
  c->vf_image = 384; c->vf_patch = 16; c->vf_width = 1024; c->vf_layers = 24; c->vf_heads = 16; c->vf_mlp = 4096;
  c->vf_text_layers = 24; c->vf_text_ctx = 64; c->vf_vocab = 256000;
  c->vf_members = 3; c->vf_embed = 512; c->vf_pool_heads = 8; c->vf_pool_layers = 4; c->vf_traj_layers = 4;
  c->vf_traj_ff = 1024; c->vf_history = 10; c->vf_action_dim = 7;
  c->use_cuda_graph = 1; c->max_observations = 1; c->num_cameras = 1;
}

int main(void) {
  CHECK(cvb_abi_version() == CVB_ABI_VERSION, "library ABI %d, header ABI %d", cvb_abi_version(), CVB_ABI_VERSION);

  /* errors: negative return, text in cvb_last_error(), nothing thrown across the boundary */
  cvb_handle* h = NULL;
  cvb_config cfg;
  bridge_config(&cfg);
  CHECK(cvb_create(NULL, &h) < 0 && strlen(cvb_last_error()) > 0, "null config accepted");
  cfg.struct_size -= 4;
  CHECK(cvb_create(&cfg, &h) < 0 && strstr(cvb_last_error(), "ABI") != NULL, "size guard: %s", cvb_last_error());
  bridge_config(&cfg);
  cfg.max_samples = 0;
  CHECK(cvb_create(&cfg, &h) < 0, "max_samples = 0 accepted");

  /* the weight manifest of the full-size configuration: reference state-dict names, bf16 / fp32 as the reference's
   * to_bfloat16_like_physical_intelligence leaves them (paligemma_with_expert.py:216-227) */
  bridge_config(&cfg);
  CHECK(cvb_create(&cfg, &h) == 0 && h != NULL, "cvb_create: %s", cvb_last_error());
  int n = cvb_required_weight_count(h);
  CHECK(n > 700, "only %d required weights", n);
  int n_bf16 = 0, n_f32 = 0, seen_expert_q = 0, seen_state_proj = 0, seen_trunk = 0;
  for (int i = 0; i < n; ++i) {
    const char* name = cvb_required_weight_name(h, i);
    int dt = cvb_required_weight_dtype(h, i);
    CHECK(name != NULL && name[0] != '\0', "weight %d has no name", i);
    CHECK(dt == CVB_BF16 || dt == CVB_F32, "weight %s: dtype %d", name, dt);
    if (dt == CVB_BF16) ++n_bf16; else ++n_f32;
    if (strstr(name, "gemma_expert") && strstr(name, "layers.17.self_attn.q_proj.weight")) {
      seen_expert_q = 1;
      CHECK(dt == CVB_BF16, "%s must be bf16", name);
    }
    if (strstr(name, "state_proj.weight")) {
      seen_state_proj = 1;
      CHECK(dt == CVB_F32, "%s stays fp32 in the reference", name);
    }
    if (strncmp(name, "verifier.trunk.", 15) == 0) seen_trunk = 1;
  }
  CHECK(seen_expert_q && seen_state_proj && seen_trunk, "manifest misses a family (%d %d %d)", seen_expert_q,
        seen_state_proj, seen_trunk);
  CHECK(cvb_required_weight_name(h, n) == NULL || cvb_required_weight_name(h, n)[0] == '\0' ||
            cvb_required_weight_name(h, -1) == NULL,
        "out-of-range index is not rejected");
  /* finalize without weights must fail with a message naming a missing tensor - and must not need a device to say so */
  CHECK(cvb_finalize(h, NULL) < 0 && strlen(cvb_last_error()) > 0, "finalize without weights succeeded");
  printf("missing-weight message: %s\n", cvb_last_error());
  cvb_destroy(h);
  cvb_destroy(NULL);

  /* denoise schedule: time = 1.0 stepping by dt = -1/num_steps while time >= -dt/2 (modeling_pi0.py:697-715) */
  float times[64], dt = 0.f;
  int steps = cvb_denoise_times_host(10, times, 64, &dt);
  CHECK(steps == 10 && times[0] == 1.0f && dt == -0.1f, "schedule: %d steps, t0 %g, dt %g", steps, times[0], dt);
  /* a short buffer is never overrun: the count still says how many entries the schedule has (snprintf convention) */
  times[4] = -7.0f;
  CHECK(cvb_denoise_times_host(10, times, 4, &dt) == 10 && times[4] == -7.0f && times[3] < times[2],
        "schedule overran a short buffer");

  uint16_t emb[1024];
  cvb_time_embedding_host(1.0f, 1024, 4e-3, 4.0, emb);
  /* sin(2 pi * 1.0 / 4e-3 ...) is arbitrary, but the cos half at the longest period (4.0) is cos(pi/2) ~ 0 and the
   * first sin entry is sin(2 pi / 4e-3) = sin(500 pi) ~ 0: both round to a bf16 of magnitude < 2^-7 */
  CHECK((emb[0] & 0x7fff) < 0x3c00, "time embedding[0] = 0x%04x", emb[0]);

  printf("manifest: %d tensors (%d bf16, %d fp32); ABI %d OK\n", n, n_bf16, n_f32, cvb_abi_version());
  return 0;
}
