// Fused score all-gather + selection over NVLink peer memory (SURVEY.md section 8e, section 5 last row): the only
// exchange of the rephrase-sharded decision (BASELINE.json configs[3]).
//
// Every rank owns a "mailbox" in its own HBM: [2 parities][world slots][slot_floats] payload + [2][world] epoch flags.
// One kernel per decision and rank:
//   1. PUSH   - the rank's score / action slice is written straight into slot `rank` of EVERY peer's mailbox with plain
//               st.global on peer-mapped pointers (cudaIpc: the stores travel over NVLink / NVSwitch), followed by a
//               system-scope fence and a release store of the epoch into the peer's flag word;
//   2. WAIT   - spin (ld.acquire.sys) until all `world` flags of the LOCAL mailbox carry this epoch;
//   3. SELECT - compact the slots (shards are contiguous rephrase ranges, possibly ragged), then group-mean -> argmax
//               group -> argmax inside the group (efficient_ensemble_merged.py:417-447), identical on every rank.
// No host round trip, no NCCL call, no second launch.  Mailboxes are double-buffered by epoch parity: a peer can only
// start epoch e + 2 after this rank has published e + 1, i.e. after it finished reading e.
//
// The reference has no multi-GPU inference path; this replaces what a straight port would do
// with two dist.all_gather calls + a selection launch (cover.py gather_and_select, kept for gloo / CPU tests).
#include <cuda_runtime.h>

#include <cstring>
#include <vector>

#include "../../include/coverb200.h"
#include "host_common.h"

namespace cvb {

struct CommState {
  int rank = 0, world = 1;
  int slot_floats = 0;            // payload capacity of one slot
  float* box = nullptr;           // local mailbox payload [2][world][slot_floats]
  unsigned* flags = nullptr;      // local flags [2][world]
  unsigned* epoch = nullptr;      // device-side call counter (graph replays advance it themselves)
  float** peer_box = nullptr;     // device array [world]: every rank's payload base as mapped here
  unsigned** peer_flags = nullptr;
  std::vector<void*> opened;      // cudaIpcOpenMemHandle mappings to close
  void* base = nullptr;           // one allocation: payload | flags | epoch
  size_t flags_off = 0, epoch_off = 0, bytes = 0;
  unsigned* err = nullptr;
};

namespace {

__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long gtimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

struct GatherArgs {
  int rank, world, slot_floats;
  float* const* peer_box;
  unsigned* const* peer_flags;
  float* box;
  unsigned* flags;
  unsigned* epoch;
  unsigned* err;
  const float* scores;   // local [n_loc]
  const float* actions;  // local [n_loc * act_floats] or nullptr
  int n_loc, act_floats;
  int R, K;
  float* out_scores;   // [R*K]
  float* out_actions;  // [R*K * act_floats] or nullptr
  float* group_mean;   // [R] or nullptr
  int* best_idx;
  float* best_score;
  long long spin_limit_ns;
};

// contiguous rephrase shards, sizes differ by at most one (same rule as cover.py rephrase_shard)
__device__ __forceinline__ void shard_of(int R, int world, int r, int* start, int* count) {
  const int base = R / world, rem = R % world;
  *start = r * base + min(r, rem);
  *count = base + (r < rem ? 1 : 0);
}

__global__ void __launch_bounds__(1024) allgather_select_kernel(const GatherArgs g) {
  extern __shared__ float sh_mean[];
  const unsigned e = *g.epoch + 1u;
  const int par = static_cast<int>(e & 1u);
  const int tid = threadIdx.x, nt = blockDim.x;
  // ---- 1. push my slice into slot `rank` of every mailbox (my own included)
  const int pay = g.n_loc * (1 + g.act_floats);
  for (int p = 0; p < g.world; ++p) {
    float* dst = g.peer_box[p] + (static_cast<long>(par) * g.world + g.rank) * g.slot_floats;
    for (int i = tid; i < g.n_loc; i += nt) dst[i] = g.scores[i];
    if (g.actions != nullptr)
      for (int i = tid; i < g.n_loc * g.act_floats; i += nt) dst[g.n_loc + i] = g.actions[i];
  }
  __threadfence_system();
  __syncthreads();
  if (tid < g.world) st_release_sys(g.peer_flags[tid] + par * g.world + g.rank, e);
  // ---- 2. wait until every rank's slice of this epoch has landed here
  if (tid < g.world) {
    const unsigned long long t0 = gtimer_ns();
    while (ld_acquire_sys(g.flags + par * g.world + tid) != e) {
      if (static_cast<long long>(gtimer_ns() - t0) > g.spin_limit_ns) {  // a missing peer must not hang the device
        atomicExch(g.err, 0xC0000000u | static_cast<unsigned>(tid));
        __threadfence_system();
        __trap();
      }
    }
  }
  __syncthreads();
  // ---- 3. compact (rephrase-major order is preserved: shards are contiguous) and select
  (void)pay;
  for (int p = 0; p < g.world; ++p) {
    int r0, rc;
    shard_of(g.R, g.world, p, &r0, &rc);
    const float* src = g.box + (static_cast<long>(par) * g.world + p) * g.slot_floats;
    const int n = rc * g.K;
    for (int i = tid; i < n; i += nt) g.out_scores[r0 * g.K + i] = __ldcg(src + i);
    if (g.out_actions != nullptr)
      for (int i = tid; i < n * g.act_floats; i += nt) g.out_actions[static_cast<long>(r0) * g.K * g.act_floats + i] = __ldcg(src + n + i);
  }
  __syncthreads();
  for (int gr = tid; gr < g.R; gr += nt) {
    float s = 0.f;
    for (int k = 0; k < g.K; ++k) s += g.out_scores[gr * g.K + k];
    const float m = s / static_cast<float>(g.K);
    sh_mean[gr] = m;
    if (g.group_mean != nullptr) g.group_mean[gr] = m;
  }
  __syncthreads();
  if (tid == 0) {
    int bg = 0;
    for (int gr = 1; gr < g.R; ++gr)
      if (sh_mean[gr] > sh_mean[bg]) bg = gr;  // first maximum wins ties, like torch.max
    int bk = 0;
    for (int k = 1; k < g.K; ++k)
      if (g.out_scores[bg * g.K + k] > g.out_scores[bg * g.K + bk]) bk = k;
    *g.best_idx = bg * g.K + bk;
    *g.best_score = g.out_scores[bg * g.K + bk];
    *g.epoch = e;
  }
}

}  // namespace
}  // namespace cvb

using cvb::CommState;

extern "C" {

int cvb_comm_create(int rank, int world, int max_slot_floats, cvb_comm** out) {
  CVB_REQUIRE(out != nullptr && world >= 1 && rank >= 0 && rank < world && max_slot_floats >= 1, "bad communicator arguments");
  CVB_REQUIRE(world <= 64, "at most 64 ranks");
  CommState* c = new CommState();
  c->rank = rank, c->world = world, c->slot_floats = (max_slot_floats + 3) / 4 * 4;
  const size_t payload = static_cast<size_t>(2) * world * c->slot_floats * sizeof(float);
  c->flags_off = (payload + 255) / 256 * 256;
  c->epoch_off = c->flags_off + 2 * world * sizeof(unsigned);
  c->bytes = c->epoch_off + 4 * sizeof(unsigned);
  CVB_CUDA(cudaMalloc(&c->base, c->bytes));
  CVB_CUDA(cudaMemset(c->base, 0, c->bytes));
  c->box = reinterpret_cast<float*>(c->base);
  c->flags = reinterpret_cast<unsigned*>(reinterpret_cast<char*>(c->base) + c->flags_off);
  c->epoch = reinterpret_cast<unsigned*>(reinterpret_cast<char*>(c->base) + c->epoch_off);
  c->err = c->epoch + 1;
  CVB_CUDA(cudaMalloc(&c->peer_box, world * sizeof(float*)));
  CVB_CUDA(cudaMalloc(&c->peer_flags, world * sizeof(unsigned*)));
  *out = reinterpret_cast<cvb_comm*>(c);
  return 0;
}

int cvb_comm_handle_bytes(void) { return static_cast<int>(sizeof(cudaIpcMemHandle_t)); }

int cvb_comm_local_handle(cvb_comm* comm, void* handle_out_host) {
  CommState* c = reinterpret_cast<CommState*>(comm);
  CVB_REQUIRE(c != nullptr && handle_out_host != nullptr, "null argument");
  cudaIpcMemHandle_t hnd;
  CVB_CUDA(cudaIpcGetMemHandle(&hnd, c->base));
  memcpy(handle_out_host, &hnd, sizeof(hnd));
  return 0;
}

int cvb_comm_open_peers(cvb_comm* comm, const void* all_handles_host) {
  CommState* c = reinterpret_cast<CommState*>(comm);
  CVB_REQUIRE(c != nullptr && (all_handles_host != nullptr || c->world == 1), "null argument");
  std::vector<float*> boxes(c->world);
  std::vector<unsigned*> flags(c->world);
  for (int p = 0; p < c->world; ++p) {
    void* base = c->base;
    if (p != c->rank) {
      cudaIpcMemHandle_t hnd;
      memcpy(&hnd, reinterpret_cast<const char*>(all_handles_host) + static_cast<size_t>(p) * sizeof(hnd), sizeof(hnd));
      CVB_CUDA(cudaIpcOpenMemHandle(&base, hnd, cudaIpcMemLazyEnablePeerAccess));
      c->opened.push_back(base);
    }
    boxes[p] = reinterpret_cast<float*>(base);
    flags[p] = reinterpret_cast<unsigned*>(reinterpret_cast<char*>(base) + c->flags_off);
  }
  CVB_CUDA(cudaMemcpy(c->peer_box, boxes.data(), c->world * sizeof(float*), cudaMemcpyHostToDevice));
  CVB_CUDA(cudaMemcpy(c->peer_flags, flags.data(), c->world * sizeof(unsigned*), cudaMemcpyHostToDevice));
  return 0;
}

void cvb_comm_destroy(cvb_comm* comm) {
  CommState* c = reinterpret_cast<CommState*>(comm);
  if (c == nullptr) return;
  for (void* p : c->opened) cudaIpcCloseMemHandle(p);
  cudaFree(c->peer_box);
  cudaFree(c->peer_flags);
  cudaFree(c->base);
  delete c;
}

int cvb_allgather_select(cvb_comm* comm, const float* local_scores, const float* local_actions, int act_floats, int R, int K,
                         float* scores, float* actions, float* group_mean, int32_t* best_idx, float* best_score,
                         void* stream) {
  CommState* c = reinterpret_cast<CommState*>(comm);
  CVB_REQUIRE(c != nullptr && local_scores != nullptr && scores != nullptr && best_idx != nullptr && best_score != nullptr,
              "null argument");
  CVB_REQUIRE(R >= c->world && K >= 1, "need at least one rephrase per rank");
  CVB_REQUIRE((local_actions == nullptr) == (actions == nullptr), "actions in and out must both be given or both be NULL");
  const int base = R / c->world, rem = R % c->world;
  const int n_loc = (base + (c->rank < rem ? 1 : 0)) * K, n_max = (base + (rem ? 1 : 0)) * K;
  if (local_actions == nullptr) act_floats = 0;
  CVB_REQUIRE(n_max * (1 + act_floats) <= c->slot_floats, "slice does not fit the mailbox slot (max_slot_floats)");
  cvb::GatherArgs g;
  g.rank = c->rank, g.world = c->world, g.slot_floats = c->slot_floats;
  g.peer_box = c->peer_box, g.peer_flags = c->peer_flags, g.box = c->box, g.flags = c->flags, g.epoch = c->epoch, g.err = c->err;
  g.scores = local_scores, g.actions = local_actions, g.n_loc = n_loc, g.act_floats = act_floats, g.R = R, g.K = K;
  g.out_scores = scores, g.out_actions = actions, g.group_mean = group_mean, g.best_idx = best_idx, g.best_score = best_score;
  g.spin_limit_ns = 5000LL * 1000 * 1000;
  cvb::allgather_select_kernel<<<1, 1024, (R + 1) * sizeof(float), (cudaStream_t)stream>>>(g);
  CVB_LAUNCHED();
  return 0;
}

}  // extern "C"
