// Host dispatch for the tcgen05 GEMM: tensor-map construction (cached), tile-shape heuristic, launch.
#include <algorithm>
#include <cstdlib>
#include <mutex>
#include <unordered_map>

#include "gemm_tcgen05.cuh"
#include "gemm_tcgen05_2sm.cuh"
#include "host_common.h"
#include "ops.h"

namespace cvb {

namespace {

struct TmapKey {
  const void* ptr;
  uint64_t rows, cols, ld;
  uint32_t box_rows;
  bool operator==(const TmapKey& o) const {
    return ptr == o.ptr && rows == o.rows && cols == o.cols && ld == o.ld && box_rows == o.box_rows;
  }
};
struct TmapKeyHash {
  size_t operator()(const TmapKey& k) const {
    size_t h = reinterpret_cast<size_t>(k.ptr);
    h ^= k.rows * 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2);
    h ^= k.cols * 0xC2B2AE3D27D4EB4Full + (h << 6) + (h >> 2);
    h ^= k.ld * 0x165667B19E3779F9ull + (h << 6) + (h >> 2);
    h ^= k.box_rows + (h << 6) + (h >> 2);
    return h;
  }
};

std::mutex g_tmap_mu;
std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> g_tmaps;

int get_tmap(const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows,
             CUtensorMap* out) {
  TmapKey key{ptr, rows, cols, ld, box_rows};
  std::lock_guard<std::mutex> lk(g_tmap_mu);
  auto it = g_tmaps.find(key);
  if (it != g_tmaps.end()) {
    *out = it->second;
    return 0;
  }
  CUtensorMap tm;
  CVB_TRY(make_tmap_2d(&tm, ptr, rows, cols, ld, box_rows, GEMM_BK, 2));
  g_tmaps.emplace(key, tm);
  *out = tm;
  return 0;
}

}  // namespace

int get_tmap_cached(const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows, CUtensorMap* out) {
  return get_tmap(ptr, rows, cols, ld, box_rows, out);
}
bool skinny_eligible(const GemmCall& c);
int gemm_skinny(cudaStream_t st, const GemmCall& c, int force_split);

namespace {

template <int BN, int STAGES, int EPI>
int launch(cudaStream_t st, const GemmCall& c, int grid) {
  using S = GemmSmem<BN, STAGES>;
  CUtensorMap tmA, tmB;
  CVB_TRY(get_tmap(c.A, c.M, c.K, c.lda, GEMM_BM, &tmA));
  CVB_TRY(get_tmap(c.W, c.N, c.K, c.ldw, BN, &tmB));
  GemmArgs g;
  g.C = c.C;
  g.ldc = c.ldc;
  g.bias = c.bias;
  g.bias_is_f32 = c.bias_is_f32;
  g.resid = c.resid;
  g.resid_is_f32 = c.resid_is_f32;
  g.ldr = c.ldr;
  g.M = c.M;
  g.N = c.N;
  g.K = c.K;
  g.n_out = c.n_out;
  g.m_dev = c.m_dev;
  auto kern = gemm_bf16_tcgen05<BN, STAGES, EPI>;
  CVB_TRY(ensure_dyn_smem(kern, S::TOTAL));
  CVB_TRY(launch_pdl(kern, dim3(grid), dim3(GEMM_THREADS), S::TOTAL, st, 1, tmA, tmB, g));
  CVB_LAUNCHED();
  return 0;
}

// CTA-pair kernel (cta_group::2): grid = 2 x pairs, cluster (2,1,1)
template <int EPI>
int launch_2sm(cudaStream_t st, const GemmCall& c) {
  CUtensorMap tmA, tmB;
  CVB_TRY(get_tmap(c.A, c.M, c.K, c.lda, 128, &tmA));
  CVB_TRY(get_tmap(c.W, c.N, c.K, c.ldw, 128, &tmB));
  GemmArgs g;
  g.C = c.C, g.ldc = c.ldc, g.bias = c.bias, g.bias_is_f32 = c.bias_is_f32;
  g.resid = c.resid, g.resid_is_f32 = c.resid_is_f32, g.ldr = c.ldr;
  g.M = c.M, g.N = c.N, g.K = c.K, g.n_out = c.n_out, g.m_dev = nullptr;
  auto kern = gemm_bf16_tcgen05_2sm<EPI>;
  CVB_TRY(ensure_dyn_smem(kern, GEMM2_SMEM));
  const int tiles = ((c.M + 255) / 256) * ((c.N + 255) / 256);
  const int pairs = std::min(tiles, device_sm_count() / 2);
  CVB_TRY(launch_pdl(kern, dim3(2 * pairs), dim3(GEMM_THREADS), GEMM2_SMEM, st, 2, tmA, tmB, g));
  CVB_LAUNCHED();
  return 0;
}

// The pair kernel pays off when the problem is tensor-bound: at least ~one wave of 256 x 256 tiles.  CVB_GEMM_2SM=0
// disables it; force_bn = 512 forces it (tests).
bool use_2sm(const GemmCall& c) {
  static const int env = getenv("CVB_GEMM_2SM") != nullptr ? atoi(getenv("CVB_GEMM_2SM")) : 1;
  if (c.m_dev != nullptr || c.K % 8 != 0 || c.N % 8 != 0) return false;
  if (c.ldc % 8 != 0 && c.epi != EPI_F32) return false;  // 16-byte vector stores of the shared epilogue
  if (c.force_bn == 512) return true;
  if (env == 0 || c.force_bn != 0) return false;
  const long tiles = static_cast<long>((c.M + 255) / 256) * ((c.N + 255) / 256);
  return c.M >= 1024 && tiles * 10 >= device_sm_count() / 2 * 9;
}

template <int EPI>
int launch_bn(cudaStream_t st, const GemmCall& c, int bn, int grid) {
  switch (bn) {
    case 256:
      return launch<256, 4, EPI>(st, c, grid);
    case 128:
      return launch<128, 6, EPI>(st, c, grid);
    default:
      return launch<64, 8, EPI>(st, c, grid);
  }
}

}  // namespace

int gemm_bf16(cudaStream_t st, const GemmCall& c) {
  CVB_REQUIRE(c.M > 0 && c.N > 0 && c.K > 0, "empty GEMM");
  // force_bn: 0 = auto, 64/128/256 = general kernel with that tile, -100 = skinny auto, -1..-16 = skinny with that split
  if (c.force_bn < 0 && c.epi != 6) return gemm_skinny(st, c, c.force_bn == -100 ? 0 : -c.force_bn);
  // epilogue 5 = GeGLU on [64 gate | 64 up] packed rows: skinny kernel, or (force_bn = 128) the general kernel with
  // 128 x 128 tiles - twice the CTAs of the 256-wide GeGLU tiles, a third fewer bytes per CTA (denoise gate/up)
  if (c.epi == 5 && c.force_bn == 128) {
    CVB_REQUIRE(c.N % 128 == 0 && c.K % 8 == 0, "GeGLU-64 expects 128-row packed [64 gate | 64 up] blocks");
    const int tiles128 = ((c.M + GEMM_BM - 1) / GEMM_BM) * (c.N / 128);
    return launch<128, 6, EPI_GEGLU>(st, c, std::min(tiles128, device_sm_count()));
  }
  if (c.epi == 5) return gemm_skinny(st, c, 0);
  if (c.epi == 6) return gemm_splitk_partial(st, c, c.force_bn < 0 ? -c.force_bn : 0, nullptr);  // EPI_PARTIAL
  if (c.force_bn == 0 && !c.no_skinny && skinny_eligible(c)) return gemm_skinny(st, c, 0);
  CVB_REQUIRE(c.K % 8 == 0, "K must be a multiple of 8 (16-byte TMA rows)");
  CVB_REQUIRE(c.N % 8 == 0, "N must be a multiple of 8 (16-byte vector epilogue)");
  CVB_REQUIRE(c.ldc % 8 == 0 || c.epi == EPI_F32, "ldc must be a multiple of 8");
  if (use_2sm(c)) {
    if (c.epi == EPI_GEGLU) CVB_REQUIRE(c.N % 256 == 0, "EPI_GEGLU expects 256-row packed gate|up blocks");
    switch (c.epi) {
      case EPI_STORE:
        return launch_2sm<EPI_STORE>(st, c);
      case EPI_GELU:
        return launch_2sm<EPI_GELU>(st, c);
      case EPI_RESID:
        CVB_REQUIRE(c.resid != nullptr, "EPI_RESID needs a residual pointer");
        return launch_2sm<EPI_RESID>(st, c);
      case EPI_GEGLU:
        return launch_2sm<EPI_GEGLU>(st, c);
      case EPI_F32:
        return launch_2sm<EPI_F32>(st, c);
      default:
        break;
    }
  }
  const int sms = device_sm_count();
  const int m_tiles = (c.M + GEMM_BM - 1) / GEMM_BM;
  auto tiles = [&](int bn) { return m_tiles * ((c.N + bn - 1) / bn); };
  int bn = c.force_bn;
  if (c.epi == EPI_GEGLU) {
    CVB_REQUIRE(c.N % 256 == 0, "EPI_GEGLU expects 256-row packed gate|up blocks");
    bn = 256;
  } else if (bn == 0) {
    // Measured cost model (tools/gemm_shapes_bench.py): the persistent kernel walks ceil(tiles / SMs) rounds of tiles, and a
    // k-block of a 128 x BN tile costs in proportion to the bytes it pulls from L2 (A 16 KB + W BN/64 x 8 KB: 3 : 4 : 6
    // for BN = 64 : 128 : 256) - these shapes are bound by L2 -> SM bandwidth, not by the tensor pipe.  Ties go to the
    // wider tile (fewer bytes in total).  E.g. M = 1280, N = 1024: 160 narrow tiles = 2 rounds vs 80 tiles of 128 = 1 round
    // at 4/3 the cost (36 -> 24 us at K = 4096).
    long best = 0;
    for (int cand : {256, 128, 64}) {
      const long rounds = (tiles(cand) + sms - 1) / sms;
      const long cost = rounds * (cand == 256 ? 6 : cand == 128 ? 4 : 3);
      if (bn == 0 || cost < best) bn = cand, best = cost;
    }
  }
  CVB_REQUIRE(bn == 64 || bn == 128 || bn == 256, "BN must be 64, 128 or 256");
  int grid = tiles(bn);
  if (grid > sms) grid = sms;
  switch (c.epi) {
    case EPI_STORE:
      return launch_bn<EPI_STORE>(st, c, bn, grid);
    case EPI_GELU:
      return launch_bn<EPI_GELU>(st, c, bn, grid);
    case EPI_RESID:
      CVB_REQUIRE(c.resid != nullptr, "EPI_RESID needs a residual pointer");
      return launch_bn<EPI_RESID>(st, c, bn, grid);
    case EPI_GEGLU:
      return launch<256, 4, EPI_GEGLU>(st, c, grid);
    case EPI_F32:
      return launch_bn<EPI_F32>(st, c, bn, grid);
    default:
      set_last_error("unknown GEMM epilogue kind");
      return -1;
  }
}

}  // namespace cvb
