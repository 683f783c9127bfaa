// Host-side helpers shared by the C-ABI translation units: error plumbing (no exceptions cross the
// ABI), the driver entry point for TMA descriptors, and a tiny tensor-map cache.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>
#include <map>
#include <string>
#include <tuple>
#include <utility>

namespace cvb {

void set_last_error(const std::string& msg);

#define CVB_CUDA(expr)                                                                      \
  do {                                                                                      \
    cudaError_t _e = (expr);                                                                \
    if (_e != cudaSuccess) {                                                                \
      ::cvb::set_last_error(std::string(#expr) + " failed: " + cudaGetErrorString(_e) +     \
                            " at " + __FILE__ + ":" + std::to_string(__LINE__));            \
      return -2;                                                                            \
    }                                                                                       \
  } while (0)

#define CVB_REQUIRE(cond, msg)                                                              \
  do {                                                                                      \
    if (!(cond)) {                                                                          \
      ::cvb::set_last_error(std::string("requirement failed: ") + #cond + " - " + (msg) +   \
                            " at " + __FILE__ + ":" + std::to_string(__LINE__));            \
      return -1;                                                                            \
    }                                                                                       \
  } while (0)

#define CVB_TRY(expr)            \
  do {                           \
    int _rc = (expr);            \
    if (_rc != 0) return _rc;    \
  } while (0)

// 2-D bf16/fp32 row-major tensor [rows, cols] with leading dimension ld (elements); box =
// [box_rows, box_cols] with the 128-byte swizzle (box_cols * elem_size must be 128 bytes).
int make_tmap_2d(CUtensorMap* out, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld,
                 uint32_t box_rows, uint32_t box_cols, int elem_bytes);

int device_sm_count();

// cudaFuncAttributeMaxDynamicSharedMemorySize is a PER-DEVICE attribute: raise it (once per device and kernel, and again
// whenever a larger size is needed) on the CURRENT device.  Handles on several GPUs of one process each get it.
int ensure_dyn_smem_impl(const void* func, int bytes);
template <typename F>
int ensure_dyn_smem(F* func, int bytes) {
  return ensure_dyn_smem_impl(reinterpret_cast<const void*>(func), bytes);
}

// number of kernel launches enqueued by this library (bench.py reports it as gpu_launches)
void count_launch();
long launch_count();

// Launch with programmatic stream serialization (PDL) and an optional cluster dimension.  CVB_PDL=0 in the
// environment disables the attribute (kernels then serialise fully; their griddepcontrol instructions are no-ops).
bool pdl_enabled();
template <typename... P, typename... A>
int launch_pdl(void (*kern)(P...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, int cluster_x, A&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int n = 0;
  if (pdl_enabled()) {
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  if (cluster_x > 1) {
    attr[n].id = cudaLaunchAttributeClusterDimension;
    attr[n].val.clusterDim.x = cluster_x;
    attr[n].val.clusterDim.y = 1;
    attr[n].val.clusterDim.z = 1;
    ++n;
  }
  cfg.attrs = attr;
  cfg.numAttrs = n;
  CVB_CUDA(cudaLaunchKernelEx(&cfg, kern, std::forward<A>(args)...));
  return 0;
}

#define CVB_LAUNCHED()                \
  do {                                \
    ::cvb::count_launch();            \
    CVB_CUDA(cudaGetLastError());     \
  } while (0)

}  // namespace cvb
