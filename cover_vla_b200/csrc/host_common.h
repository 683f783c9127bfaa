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

// number of kernel launches enqueued by this library (bench.py reports it as gpu_launches)
void count_launch();
long launch_count();

#define CVB_LAUNCHED()                \
  do {                                \
    ::cvb::count_launch();            \
    CVB_CUDA(cudaGetLastError());     \
  } while (0)

}  // namespace cvb
