// Host side of the persistent expert kernel (expert_mega.cuh): stacked weight tensors, the phase program of one
// denoise step and the (cooperative) launches.  See engine_pi0.cu run_denoise for the call sites.
//
// Reference: the suffix-token layer loop of PaliGemmaWithExpertModel.forward (paligemma_with_expert.py:258-349) inside
// PI0FlowMatching.denoise_step (modeling_pi0.py:717-752).
#include <algorithm>
#include <cstdlib>

#include "engine.h"
#include "expert_mega.cuh"
#include "pi0_kernels.h"

namespace cvb {

int get_tmap_cached(const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows, CUtensorMap* out);
extern unsigned long long* g_skinny_ts;  // diagnostics: cvb_debug_set_timestamps

namespace {

int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e != nullptr ? atoi(e) : dflt;
}

int stack_rows(cvb_handle* h, cudaStream_t st, const std::vector<const bf16*>& parts, int64_t rows, int64_t cols, bf16** out) {
  CVB_TRY(dalloc_t(h, out, static_cast<size_t>(parts.size()) * rows * cols));
  for (size_t l = 0; l < parts.size(); ++l)
    CVB_CUDA(cudaMemcpyAsync(*out + l * rows * cols, parts[l], rows * cols * sizeof(bf16), cudaMemcpyDeviceToDevice, st));
  return 0;
}

}  // namespace

// Called once from pi0_finalize (after the per-layer packing): decides eligibility, stacks the expert weights so one
// 2-D tensor map per linear layer covers all layers ([layers * rows, K]; a tile never straddles real data of two
// layers because masked features are never stored), allocates the partial buffers and the barrier word.
int expert_mega_prepare(cvb_handle* h, cudaStream_t st) {
  const cvb_config& c = h->cfg;
  Pi0State& s = h->pi0;
  ExpertMega& m = s.mega;
  m.mode = 0;
  const int We = c.ex_width, hd = c.head_dim, qd = c.heads * hd, qkvw = qd + 2 * hd, I = c.ex_mlp;
  const int Mmax = h->rm_total() * c.max_samples * h->suffix_len();
  // 0 (default) = separate kernels, 1 = persistent chain kernel.  Measured on B200 (profiles/r2_mega_timeline.txt): the
  // chain kernel is correct and deterministic but not faster - the loop is bound by L2 -> SM traffic of activations and
  // split-K partials (~170 MB per layer at ~7 TB/s), not by launches; see DESIGN.md section 3.6.
  const int want = env_int("CVB_DENOISE_MEGA", 0);
  const bool shape_ok = hd == 256 && c.heads * h->suffix_len() <= 128 && We % 64 == 0 && I % 64 == 0 && qd % 64 == 0 &&
                        Mmax <= 256 && We <= 4096 && s.ex_gu_half == 64;
  if (want == 0 || !shape_ok) return 0;
  int dev = 0, coop = 0;
  CVB_CUDA(cudaGetDevice(&dev));
  CVB_CUDA(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
  if (!coop) return 0;

  const int L = c.layers;
  std::vector<const bf16*> wq(L), wo(L), wg(L), wd(L);
  for (int l = 0; l < L; ++l) wq[l] = s.ex[l].wqkv, wo[l] = s.ex[l].wo, wg[l] = s.ex[l].wgu, wd[l] = s.ex[l].wd;
  m.packed = ((I + 63) / 64) * 128;
  CVB_TRY(stack_rows(h, st, wq, qkvw, We, &m.wqkv_all));
  CVB_TRY(stack_rows(h, st, wo, We, qd, &m.wo_all));
  CVB_TRY(stack_rows(h, st, wg, m.packed, We, &m.wgu_all));
  CVB_TRY(stack_rows(h, st, wd, We, I, &m.wd_all));

  m.grid = device_sm_count();
  const int cap = env_int("CVB_MEGA_GRID", 0);
  if (cap > 0) m.grid = std::min(m.grid, cap);
  auto pick = [&](const char* env, int ftiles, int kb, int dflt) {
    int sp = std::min(std::min(dflt, std::max(1, m.grid / ftiles)), std::max(1, kb / 2));
    const int e = env_int(env, 0);
    if (e > 0) sp = std::min(e, kb);
    return std::max(1, std::min(sp, 16));
  };
  m.s_qkv = pick("CVB_MEGA_SQ", (qkvw + 127) / 128, We / 64, 7);
  m.s_o = pick("CVB_MEGA_SO", (We + 127) / 128, qd / 64, 8);
  m.s_d = pick("CVB_MEGA_SD", (We + 127) / 128, I / 64, 8);
  CVB_TRY(dalloc_t(h, &m.part_qkv, static_cast<size_t>(m.s_qkv) * Mmax * qkvw));
  CVB_TRY(dalloc_t(h, &m.part_o, static_cast<size_t>(m.s_o) * Mmax * We));
  CVB_TRY(dalloc_t(h, &m.part_d, static_cast<size_t>(m.s_d) * Mmax * We));
  CVB_TRY(dalloc_t(h, &m.bar, 4));
  CVB_CUDA(cudaMemsetAsync(m.bar, 0, 4 * sizeof(unsigned), st));
  m.err = m.bar + 2;

  CVB_CUDA(cudaFuncSetAttribute(expert_mega_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, MK_SMEM));
  int occ = 0;
  CVB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, expert_mega_kernel, MK_THREADS, MK_SMEM));
  if (occ < 1) return 0;  // does not fit: keep the separate kernels
  m.mode = want;
  return 0;
}

// Build (once per row count) the phase program of one denoise step for `rows` suffix rows.
int expert_mega_program(cvb_handle* h, int rows, cudaStream_t st, const MegaProgram** out) {
  const cvb_config& c = h->cfg;
  Pi0State& s = h->pi0;
  ExpertMega& m = s.mega;
  auto it = m.programs.find(rows);
  if (it != m.programs.end()) {
    *out = &it->second;
    return 0;
  }
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  CVB_CUDA(cudaStreamIsCapturing(st, &cap));
  CVB_REQUIRE(cap == cudaStreamCaptureStatusNone, "the expert program must be built before graph capture (warm-up call)");
  const int We = c.ex_width, hd = c.head_dim, qd = c.heads * hd, qkvw = qd + 2 * hd, I = c.ex_mlp, L = c.layers;
  const int rows_pad = (rows + 15) / 16 * 16;
  const int Mmax = h->rm_total() * c.max_samples * h->suffix_len();
  const float* w_norm = nullptr;
  {
    const void* p = nullptr;
    CVB_TRY(get_weight(h, "paligemma_with_expert.gemma_expert.model.norm.weight", CVB_F32, We, &p));
    w_norm = reinterpret_cast<const float*>(p);
  }
  MegaProgram pg;
  pg.rows = rows;
  const int G = m.grid;
  auto gemm_partial = [&](int wmap, int amap, int layer, int rows_per_layer, int K, int splits, float* P, int rot) {
    MegaPhase ph{};
    ph.kind = MK_GEMM_PARTIAL, ph.wmap = wmap, ph.amap = 4 + amap, ph.w_row0 = layer * rows_per_layer;
    ph.ftiles = (rows_per_layer + 127) / 128, ph.kb_total = K / 64, ph.splits = std::min(splits, ph.kb_total);
    ph.rows = rows, ph.rows_pad = rows_pad, ph.n_feat = rows_per_layer, ph.rot = rot;
    ph.P = P, ph.ldp = rows_per_layer, ph.split_stride = static_cast<long>(Mmax) * rows_per_layer;
    return ph;
  };
  auto norm = [&](const float* P, int S, long ldp, const void* resid, int resid_f32, const void* w, int w_f32, bf16* h_out) {
    MegaPhase ph{};
    ph.kind = MK_NORM, ph.rows = rows, ph.rot = 0;
    ph.nP = P, ph.nS = S, ph.n_ldp = ldp, ph.n_split_stride = static_cast<long>(Mmax) * ldp;
    ph.resid = resid, ph.resid_f32 = resid_f32, ph.ldr = We, ph.nw = w, ph.nw_f32 = w_f32;
    ph.h_out = h_out, ph.ldh = We, ph.y = s.xe, ph.ldy = We, ph.width = We, ph.eps = 1e-6f;
    return ph;
  };
  const int u_o = ((We + 127) / 128) * m.s_o, u_d = ((We + 127) / 128) * m.s_d;
  const int rot_o = std::max(0, G - u_o), rot_d = std::max(0, std::min(20, G - u_d));
  std::vector<MegaPhase>& v = pg.host;
  for (int l = 0; l < L; ++l) {
    const GemmaLayer& Ly = s.ex[l];
    pg.layer_first.push_back(static_cast<int>(v.size()));
    if (l == 0) v.push_back(norm(nullptr, 0, We, s.suffix, 1, Ly.in_norm, 0, nullptr));
    v.push_back(gemm_partial(0, 0, l, qkvw, We, m.s_qkv, m.part_qkv, 0));
    pg.attn_after.push_back(static_cast<int>(v.size()));  // the attention of layer l runs between these two phases
    v.push_back(gemm_partial(1, 1, l, We, qd, m.s_o, m.part_o, rot_o));
    v.push_back(norm(m.part_o, v.back().splits, We, l == 0 ? static_cast<const void*>(s.suffix) : static_cast<const void*>(s.he),
                     l == 0 ? 1 : 0, Ly.post_norm, 0, s.he));
    {
      MegaPhase ph{};
      ph.kind = MK_GEMM_GEGLU, ph.wmap = 2, ph.amap = 4 + 3, ph.w_row0 = l * m.packed;
      ph.ftiles = m.packed / 128, ph.splits = (rows + 127) / 128, ph.kb_total = We / 64;
      ph.rows = rows, ph.rows_pad = rows_pad, ph.n_feat = I, ph.rot = 0;
      ph.C = s.act_e, ph.ldc = I;
      v.push_back(ph);
    }
    v.push_back(gemm_partial(3, 2, l, We, I, m.s_d, m.part_d, rot_d));
    if (l + 1 < L)
      v.push_back(norm(m.part_d, v.back().splits, We, s.he, 0, s.ex[l + 1].in_norm, 0, s.he));
    else
      v.push_back(norm(m.part_d, v.back().splits, We, s.he, 0, w_norm, 1, s.he));
  }
  pg.layer_first.push_back(static_cast<int>(v.size()));
  CVB_TRY(dalloc_t(h, &pg.dev, v.size()));

  // tensor maps: weights [layers * rows, K] box 128; activations box rows_pad (swapped GEMMs) / 128 (gate-up)
  CVB_TRY(get_tmap_cached(m.wqkv_all, static_cast<uint64_t>(L) * qkvw, We, We, 128, &pg.maps.m[0]));
  CVB_TRY(get_tmap_cached(m.wo_all, static_cast<uint64_t>(L) * We, qd, qd, 128, &pg.maps.m[1]));
  CVB_TRY(get_tmap_cached(m.wgu_all, static_cast<uint64_t>(L) * m.packed, We, We, 128, &pg.maps.m[2]));
  CVB_TRY(get_tmap_cached(m.wd_all, static_cast<uint64_t>(L) * We, I, I, 128, &pg.maps.m[3]));
  CVB_TRY(get_tmap_cached(s.xe, rows, We, We, rows_pad, &pg.maps.m[4]));
  CVB_TRY(get_tmap_cached(s.attn_e, rows, qd, qd, rows_pad, &pg.maps.m[5]));
  CVB_TRY(get_tmap_cached(s.act_e, rows, I, I, rows_pad, &pg.maps.m[6]));
  CVB_TRY(get_tmap_cached(s.xe, rows, We, We, 128, &pg.maps.m[7]));
  CVB_CUDA(cudaMemcpyAsync(pg.dev, v.data(), v.size() * sizeof(MegaPhase), cudaMemcpyHostToDevice, st));
  CVB_CUDA(cudaStreamSynchronize(st));
  auto res = m.programs.emplace(rows, std::move(pg));
  *out = &res.first->second;
  return 0;
}

// Launch phases [first, first + count) of a program as ONE persistent kernel.
int expert_mega_launch(cvb_handle* h, cudaStream_t st, const MegaProgram& pg, int first, int count) {
  ExpertMega& m = h->pi0.mega;
  CVB_REQUIRE(first >= 0 && count >= 1 && first + count <= static_cast<int>(pg.host.size()), "phase range out of bounds");
  MegaArgs a;
  a.prog = pg.dev + first, a.n_phases = count, a.bar = m.bar, a.err = m.err;
  a.spin_limit_ns = 2000LL * 1000 * 1000;  // 2 s: a protocol bug traps instead of hanging the device
  // diagnostics (tools/mega_ts.py): the last 16 launches keep their stamps, after the region the attention kernels use
  static unsigned launch_seq = 0;
  a.ts = g_skinny_ts != nullptr ? g_skinny_ts + 8192 + static_cast<long>(launch_seq++ % 16) * m.grid * 64 : nullptr;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(m.grid), cfg.blockDim = dim3(MK_THREADS), cfg.dynamicSmemBytes = MK_SMEM, cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;
  attr[0].val.cooperative = 1;
  cfg.attrs = attr, cfg.numAttrs = 1;
  CVB_CUDA(cudaLaunchKernelEx(&cfg, expert_mega_kernel, pg.maps, a));
  CVB_LAUNCHED();
  return 0;
}

}  // namespace cvb
