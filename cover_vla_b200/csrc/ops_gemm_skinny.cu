// Host dispatch for the skinny (swap-AB, cluster split-K) GEMM: split heuristic, tensor maps, cluster launch.
#include <algorithm>
#include <mutex>
#include <unordered_map>

#include "gemm_skinny.cuh"
#include "host_common.h"
#include "ops.h"

namespace cvb {

int get_tmap_cached(const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows, CUtensorMap* out);

unsigned long long* g_skinny_ts = nullptr;  // diagnostics: cvb_debug_set_timestamps

namespace {

constexpr int kSmemBudget = 222 * 1024;  // ring (= staging) + receive buffers (barriers + alignment slack come on top)

template <int EPI>
int launch_skinny(cudaStream_t st, const GemmCall& c, const SkinnyArgs& g, int n_tiles) {
  CUtensorMap tmW, tmA;
  CVB_TRY(get_tmap_cached(c.W, c.N, c.K, c.ldw, 128, &tmW));
  CVB_TRY(get_tmap_cached(c.A, c.M, c.K, c.lda, g.Mp, &tmA));
  const uint32_t stage_bytes = SK_W_BYTES + g.Mp * 128;
  const uint32_t recv_bytes = (uint32_t)(g.S - 1) * g.slice * 512u;
  const uint32_t body = (std::max<uint32_t>(g.stages * stage_bytes, g.Mp * 512u) + 1023u) & ~1023u;
  const int smem = 1024 + body + recv_bytes + (2 * g.stages + 2) * 8 + 16;
  auto kern = gemm_skinny_tcgen05<EPI>;
  CVB_TRY(ensure_dyn_smem(kern, smem));
  if (g.S > 8) CVB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));  // per device: set on every such launch
  CVB_TRY(launch_pdl(kern, dim3(n_tiles * g.S), dim3(SK_THREADS), smem, st, g.S, tmW, tmA, g));
  CVB_LAUNCHED();
  return 0;
}

}  // namespace

// Auto policy (measured, tools/skinny_bench.py, profiles/r1_skinny_gemm.md): the split-K reduction moves
// (S-1) x M x N x 4 bytes over the SM-to-SM network, which sustains only ~2.4 TB/s in aggregate, so the cluster kernel
// wins only where the general kernel has few tiles AND a very long K walk (expert down-projection 4096 -> 1024: 18.0 ->
// 12.8 us; text-tower fc2 at M = 64: 17.7 -> 6.5 us); elsewhere the general kernel with PDL weight prefetch is as fast.
bool skinny_eligible(const GemmCall& c) {
  if (c.M > 208 || c.m_dev != nullptr || c.epi == EPI_GEGLU) return false;
  const int tiles64 = ((c.M + 127) / 128) * ((c.N + 63) / 64);
  return tiles64 <= 40 && c.K >= 4096;
}

int gemm_skinny(cudaStream_t st, const GemmCall& c, int force_split) {
  CVB_REQUIRE(c.M > 0 && c.M <= 256, "skinny GEMM handles 1..256 activation rows");
  CVB_REQUIRE(c.K % 8 == 0, "K must be a multiple of 8 (16-byte TMA rows)");
  CVB_REQUIRE(c.m_dev == nullptr, "skinny GEMM has no device-side row count");
  if (c.epi == EPI_GEGLU64) CVB_REQUIRE(c.N % 128 == 0, "EPI_GEGLU64 expects 128-row packed [64 gate | 64 up] blocks");
  SkinnyArgs g;
  g.C = c.C, g.ldc = c.ldc, g.bias = c.bias, g.bias_is_f32 = c.bias_is_f32;
  g.resid = c.resid, g.resid_is_f32 = c.resid_is_f32, g.ldr = c.ldr;
  g.M = c.M, g.Mp = (c.M + 15) / 16 * 16, g.N = c.N, g.K = c.K, g.n_out = c.n_out;
  const int n_tiles = (c.N + 127) / 128;
  const int kb_total = (c.K + 63) / 64;
  int S = force_split;
  if (S <= 0) {
    // aim at ~128 co-resident CTAs; clusters of up to 8 are portable
    S = (128 + n_tiles / 2) / n_tiles;
    if (S < 1) S = 1;
    if (S > 8) S = 8;
    if (S > 6 && g.Mp > 128) S = 6;  // measured: beyond 6 the reduction traffic outgrows the shorter K walk
  }
  if (S > 16) S = 16;
  if (S > kb_total) S = kb_total;
  const int stage_bytes = SK_W_BYTES + g.Mp * 128;
  // the staging tile (Mp x 512 B, aliased with the TMA ring) and the S-1 received slices must fit shared memory:
  // take the largest feasible split <= the requested one
  for (;; --S) {
    g.kbs = (kb_total + S - 1) / S;
    const int s_eff = (kb_total + g.kbs - 1) / g.kbs;  // drop splits that would own no k-block
    g.S = s_eff;
    g.slice = ((g.Mp + s_eff - 1) / s_eff + 3) / 4 * 4;
    const long recv_bytes = (long)(s_eff - 1) * g.slice * 512;
    if (recv_bytes + std::max(stage_bytes, g.Mp * 512) + 2048 <= kSmemBudget) {
      int stages = (int)((kSmemBudget - recv_bytes) / stage_bytes);
      if (stages > g.kbs) stages = g.kbs;
      if (stages > 8) stages = 8;
      if (stages < 1) stages = 1;
      g.stages = stages;
      break;
    }
    CVB_REQUIRE(S > 1, "skinny GEMM staging buffers do not fit shared memory");
  }
  g.ts = g_skinny_ts;
  g.tmem_cols = g.Mp <= 32 ? 32 : g.Mp <= 64 ? 64 : g.Mp <= 128 ? 128 : 256;
  switch (c.epi) {
    case EPI_STORE:
      return launch_skinny<EPI_STORE>(st, c, g, n_tiles);
    case EPI_GELU:
      return launch_skinny<EPI_GELU>(st, c, g, n_tiles);
    case EPI_RESID:
      CVB_REQUIRE(c.resid != nullptr, "EPI_RESID needs a residual pointer");
      return launch_skinny<EPI_RESID>(st, c, g, n_tiles);
    case EPI_F32:
      return launch_skinny<EPI_F32>(st, c, g, n_tiles);
    case EPI_GEGLU64:
      return launch_skinny<EPI_GEGLU64>(st, c, g, n_tiles);
    default:
      set_last_error("epilogue kind not supported by the skinny GEMM");
      return -1;
  }
}

// Split-K GEMM with fp32 partials in global memory (no cluster): C = fp32 [S][M][ldc], consumed by rmsnorm_reduce().
// Returns the number of splits actually used through *splits_out (<= requested; every split owns >= 1 k-block).
int gemm_splitk_partial(cudaStream_t st, const GemmCall& c, int splits, int* splits_out) {
  CVB_REQUIRE(c.M > 0 && c.M <= 256, "split-K partial GEMM handles 1..256 activation rows");
  CVB_REQUIRE(c.K % 8 == 0, "K must be a multiple of 8 (16-byte TMA rows)");
  CVB_REQUIRE(c.m_dev == nullptr && c.bias == nullptr && c.resid == nullptr,
              "split-K partial GEMM has no fused epilogue (bias / residual belong to the reducing kernel)");
  SplitKArgs g;
  g.P = reinterpret_cast<float*>(c.C), g.ldp = c.ldc, g.split_stride = static_cast<long>(c.M) * c.ldc;
  g.M = c.M, g.Mp = (c.M + 15) / 16 * 16, g.N = c.N, g.K = c.K;
  const int n_tiles = (c.N + 127) / 128;
  const int kb_total = (c.K + 63) / 64;
  int S = splits;
  if (S <= 0) {  // fill the SMs once: n_tiles * S <= #SMs, at least 2 k-blocks per split
    S = std::max(1, device_sm_count() / n_tiles);
    S = std::min(S, std::max(1, kb_total / 2));
  }
  S = std::max(1, std::min(S, kb_total));
  g.S = S;
  const int stage_bytes = SK_W_BYTES + g.Mp * 128;
  const int kbs = (kb_total + S - 1) / S;
  g.stages = std::max(1, std::min(std::min(kbs, 5), (kSmemBudget - 1024) / stage_bytes));
  g.tmem_cols = g.Mp <= 32 ? 32 : g.Mp <= 64 ? 64 : g.Mp <= 128 ? 128 : 256;
  CUtensorMap tmW, tmA;
  CVB_TRY(get_tmap_cached(c.W, c.N, c.K, c.ldw, 128, &tmW));
  CVB_TRY(get_tmap_cached(c.A, c.M, c.K, c.lda, g.Mp, &tmA));
  CVB_REQUIRE(c.N % 4 == 0 && c.ldc % 4 == 0, "split-K partial GEMM needs 4-element aligned output rows");
  // the ring doubles as the transposed output tile (Mp rows x 128 features, fp32) once the MMAs have retired
  const int ring = std::max(g.stages * stage_bytes, g.Mp * 512);
  const int smem = 1024 + ring + (2 * g.stages + 1) * 8 + 16;
  CVB_TRY(ensure_dyn_smem(gemm_splitk_partial_tcgen05<0>, smem));
  CVB_TRY(launch_pdl(gemm_splitk_partial_tcgen05<0>, dim3(n_tiles * S), dim3(SK_THREADS), smem, st, 1, tmW, tmA, g));
  CVB_LAUNCHED();
  if (splits_out != nullptr) *splits_out = S;
  return 0;
}

}  // namespace cvb
