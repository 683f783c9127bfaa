// Plain-data description of the persistent expert kernel's work (expert_mega.cuh) and its host API
// (engine_expert_mega.cu).  Included by engine.h; contains no device code.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <unordered_map>
#include <vector>

namespace cvb {

enum MegaKind : int { MK_NORM = 0, MK_GEMM_PARTIAL = 1, MK_GEMM_GEGLU = 2 };

// One phase of the program (plain data, lives in global memory; built by the host in ops_expert_mega.cu)
struct MegaPhase {
  int kind;
  // ---- GEMM phases
  int wmap, amap;   // indices into MegaMaps
  int w_row0;       // first row of this layer inside the stacked weight tensor
  int ftiles;       // 128-row weight tiles
  int splits;       // K-splits (partial) / activation-row tiles (geglu)
  int kb_total;     // 64-wide k-blocks
  int rows;         // valid activation rows M
  int rows_pad;     // UMMA N of the swapped GEMM (rows rounded up to 16)
  int n_feat;       // valid output features (weight rows of this layer)
  int rot;          // unit u runs on CTA (u + rot) % grid
  float* P;         // partial out [split][row][ldp]
  long ldp, split_stride;
  __nv_bfloat16* C;  // geglu out [row][ldc]
  long ldc;
  // ---- NORM phase: h = bf16(bf16(sum_s nP[s]) + resid) (or h = resid when nS == 0), y = RMSNorm(h)
  const float* nP;
  int nS;
  long n_split_stride, n_ldp;
  const void* resid;
  int resid_f32;
  long ldr;
  const void* nw;
  int nw_f32;
  __nv_bfloat16* h_out;  // nullptr: h is not stored
  long ldh;
  __nv_bfloat16* y;
  long ldy;
  int width;
  float eps;
};

struct MegaMaps {
  CUtensorMap m[8];
};

struct MegaArgs {
  const MegaPhase* prog;
  int n_phases;
  unsigned* bar;        // device-wide sense-reversing barrier word (never reset: the top bit flips once per barrier)
  unsigned* err;        // watchdog: set to a code != 0 before __trap()
  long long spin_limit_ns;
  unsigned long long* ts;  // diagnostics: [grid][64] globaltimer stamps (start, then per phase: mma done, work done, barrier done)
};


// host-side: the phase list of one denoise step for a given number of suffix rows
struct MegaProgram {
  int rows = 0;
  std::vector<MegaPhase> host;
  MegaPhase* dev = nullptr;
  MegaMaps maps;
  std::vector<int> layer_first;  // index of layer l's first phase (size layers + 1)
  std::vector<int> attn_after;   // index of the first phase AFTER layer l's attention (its o_proj)
};

struct ExpertMega {
  int mode = 0;  // 0 = off (separate kernels), 1 = chain kernel per layer with the attention kernel between launches
  __nv_bfloat16 *wqkv_all = nullptr, *wo_all = nullptr, *wgu_all = nullptr, *wd_all = nullptr;  // [layers * rows, K]
  int packed = 0;                     // packed gate|up rows per layer
  int s_qkv = 0, s_o = 0, s_d = 0;    // K-splits
  float *part_qkv = nullptr, *part_o = nullptr, *part_d = nullptr;
  unsigned* bar = nullptr;            // device-wide barrier word
  unsigned* err = nullptr;            // watchdog code
  int grid = 0;
  std::unordered_map<int, MegaProgram> programs;  // by suffix row count
};

}  // namespace cvb
