// One CoVer decision as ONE device pass (run_simpler_eval_with_openpi.py:322-409 minus the simulator): pi0 sampling ->
// verifier-format trajectories -> ensemble scores -> group-mean / argmax, captured into a single CUDA graph.
//
// The verifier's image/text side (SigLIP2 trunk + image-text heads) does not depend on the sampled actions.  It is
// forked onto a second stream AFTER the PaliGemma prefix (the only phase whose GEMMs fill all 148 SMs) and runs
// concurrently with the denoise loop, whose kernels are latency-bound and leave most SMs idle; the branches join
// before the trajectory encoder.  Inside the captured graph the fork/join are plain graph edges.
#include <cstring>

#include "engine.h"
#include "verifier_kernels.h"

namespace cvb {

void cover_destroy(cvb_handle* h) {
  CoverState& cs = h->cover;
  cs.graphs.destroy();
  if (cs.ev_fork) cudaEventDestroy(cs.ev_fork);
  if (cs.ev_join) cudaEventDestroy(cs.ev_join);
  if (cs.side) cudaStreamDestroy(cs.side);
  cs.ev_fork = cs.ev_join = nullptr;
  cs.side = nullptr;
}

int cover_step(cvb_handle* h, const float* image, const int64_t* lang_tokens, const int32_t* lang_len,
               const float* state, const float* noise, int R, int K, const float* vf_image, const int64_t* vf_tokens,
               const double* p01_host, const double* p99_host, const float* past, int num_past, int n_future,
               float* actions, float* traj, float* scores, float* group_mean, int32_t* best_idx, float* best_score,
               cudaStream_t st, int B) {
  const cvb_config& c = h->cfg;
  CVB_REQUIRE(h->finalized && h->vf != nullptr && c.layers > 0, "cvb_cover_step needs a handle with pi0 AND the verifier");
  CVB_REQUIRE(p01_host != nullptr && p99_host != nullptr, "action statistics required");
  CVB_REQUIRE(num_past == 0 || past != nullptr, "past actions pointer required when num_past > 0");
  CVB_REQUIRE(n_future >= 1 && n_future <= c.chunk_size, "n_future must be in [1, chunk_size]");
  CVB_REQUIRE(num_past >= 0 && num_past + n_future <= c.vf_history, "history too short for past + future actions");
  CVB_REQUIRE(c.vf_action_dim == 7 && c.max_action_dim >= 7, "the Bridge formatting needs 7-d actions");
  CoverState& cs = h->cover;
  const int N = R * K;  // candidates per observation; B observations per call (cvb_cover_step_batch)
  CVB_REQUIRE(B >= 1 && B <= h->max_obs(), "number of observations out of range (max_observations)");
  if (cs.side == nullptr) {
    CVB_CUDA(cudaStreamCreateWithFlags(&cs.side, cudaStreamNonBlocking));
    CVB_CUDA(cudaEventCreateWithFlags(&cs.ev_fork, cudaEventDisableTiming));
    CVB_CUDA(cudaEventCreateWithFlags(&cs.ev_join, cudaEventDisableTiming));
    CVB_TRY(dalloc_t(h, &cs.past, (size_t)h->max_obs() * c.vf_history * 7));
  }
  double stats[12];
  for (int i = 0; i < 6; ++i) stats[i] = p01_host[i], stats[6 + i] = p99_host[i];
  if (!cs.stats_set || memcmp(stats, cs.stats, sizeof(stats)) != 0) {
    cs.graphs.destroy();  // the statistics are baked into the captured formatting kernel
    memcpy(cs.stats, stats, sizeof(stats));
    cs.stats_set = true;
  }
  FormatStats fs;
  for (int i = 0; i < 6; ++i) fs.p01[i] = stats[i], fs.p99[i] = stats[6 + i];

  CVB_TRY(pi0_stage_inputs(h, image, lang_tokens, lang_len, state, noise, R, K, st, B));
  CVB_TRY(verifier_stage_context_inputs(h, vf_image, vf_tokens, st, B));
  if (num_past > 0)
    CVB_CUDA(cudaMemcpyAsync(cs.past, past, (size_t)B * num_past * 7 * sizeof(float), cudaMemcpyDeviceToDevice, st));

  const long key = ((long)(verifier_text_cached(h, B) ? 1 : 0) << 60) | ((long)h->cams() << 56) | ((long)h->lang_rows() << 48) | ((long)num_past << 40) | ((long)n_future << 32) | ((long)B << 24) |
                   ((long)R << 12) | K;
  CVB_TRY(cs.graphs.run(c.use_cuda_graph != 0, key, st, [&](cudaStream_t s0) {
    // One observation: the verifier's image/text side forks AFTER the prefix (whose GEMMs fill every SM) and hides under
    // the latency-bound denoise loop.  A batch of observations: the denoise GEMMs are no longer latency-bound (M = 5 N B
    // rows), so the B contexts fork right away and interleave with the whole sampler.
    if (B == 1) CVB_TRY(pi0_enqueue(h, s0, R, K, 0, B));  // vision tower + prefix
    CVB_CUDA(cudaEventRecord(cs.ev_fork, s0));
    CVB_CUDA(cudaStreamWaitEvent(cs.side, cs.ev_fork, 0));
    int rc_side = 0;
    rc_side = verifier_enqueue_context(h, cs.side, 0, B);  // every observation's towers and heads in one pass
    CVB_CUDA(cudaEventRecord(cs.ev_join, cs.side));  // always rejoin, even on error, so a capture can end cleanly
    int rc = rc_side;
    if (rc == 0 && B > 1) rc = pi0_enqueue(h, s0, R, K, 0, B);
    if (rc == 0) rc = pi0_enqueue(h, s0, R, K, 1, B);  // denoise loop
    if (rc == 0)
      rc = format_trajectories(s0, pi0_actions_buffer(h), B * N, c.chunk_size, c.max_action_dim, fs, cs.past, num_past,
                               c.vf_history, n_future, verifier_traj_buffer(h), B > 1 ? N : 0);
    CVB_CUDA(cudaStreamWaitEvent(s0, cs.ev_join, 0));
    if (rc == 0) rc = verifier_enqueue_score(h, s0, N, R, K, B);
    return rc;
  }));
  verifier_note_context(h, B);

  if (actions != nullptr)
    CVB_CUDA(cudaMemcpyAsync(actions, pi0_actions_buffer(h), (size_t)B * N * c.chunk_size * c.max_action_dim * sizeof(float),
                             cudaMemcpyDeviceToDevice, st));
  if (traj != nullptr)
    CVB_CUDA(cudaMemcpyAsync(traj, verifier_traj_buffer(h), (size_t)B * N * c.vf_history * 7 * sizeof(float),
                             cudaMemcpyDeviceToDevice, st));
  return verifier_copy_results(h, N, R, scores, group_mean, best_idx, best_score, st, B);
}

}  // namespace cvb
