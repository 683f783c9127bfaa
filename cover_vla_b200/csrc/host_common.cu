#include "host_common.h"

#include <atomic>
#include <cstdlib>
#include <mutex>

namespace cvb {

static thread_local std::string g_last_error;

void set_last_error(const std::string& msg) { g_last_error = msg; }
const char* get_last_error() { return g_last_error.c_str(); }

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                    const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
    if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) fn = (PFN_encodeTiled)p;
  });
  return fn;
}

int make_tmap_2d(CUtensorMap* out, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld,
                 uint32_t box_rows, uint32_t box_cols, int elem_bytes) {
  PFN_encodeTiled fn = get_encode_fn();
  CVB_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled driver entry point unavailable");
  CVB_REQUIRE((reinterpret_cast<uintptr_t>(ptr) & 15) == 0, "TMA base must be 16-byte aligned");
  CVB_REQUIRE((ld * elem_bytes) % 16 == 0, "TMA row stride must be a multiple of 16 bytes");
  CVB_REQUIRE(box_cols * elem_bytes == 128, "box inner extent must be 128 bytes (SWIZZLE_128B)");
  CVB_REQUIRE(box_rows <= 256, "box rows <= 256");
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {ld * (uint64_t)elem_bytes};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUtensorMapDataType dt =
      elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  CUresult r = fn(out, dt, 2, const_cast<void*>(ptr), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_error("cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r));
    return -2;
  }
  return 0;
}

// 3-D view [rows][heads][hd] of a row-major bf16 matrix whose rows hold `heads` consecutive hd-wide heads (leading
// dimension ld): box = [box_rows rows] x [box_heads heads, 0 = all] x [64 columns], 128-byte swizzle; columns past hd
// are zero-filled (a head_dim that is not a multiple of 64 is padded for free).  One box lands in shared memory
// as (row, head)-major 128-byte lines, i.e. the heads are folded into the UMMA M dimension (ops_attention_umma.cu).
int make_tmap_3d_heads(CUtensorMap* out, const void* ptr, uint64_t rows, uint64_t heads, uint64_t hd, uint64_t ld,
                       uint32_t box_rows, uint32_t box_heads) {
  PFN_encodeTiled fn = get_encode_fn();
  CVB_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled driver entry point unavailable");
  CVB_REQUIRE((reinterpret_cast<uintptr_t>(ptr) & 15) == 0, "TMA base must be 16-byte aligned");
  CVB_REQUIRE((ld * 2) % 16 == 0 && (hd * 2) % 16 == 0, "TMA strides must be multiples of 16 bytes");
  CVB_REQUIRE(heads <= 256 && box_rows <= 256, "box extents <= 256");
  cuuint64_t gdim[3] = {hd, heads, rows};
  cuuint64_t gstride[2] = {hd * 2, ld * 2};
  cuuint32_t box[3] = {64, box_heads == 0 ? static_cast<cuuint32_t>(heads) : box_heads, box_rows};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_error("cuTensorMapEncodeTiled (3-D) failed with CUresult " + std::to_string((int)r));
    return -2;
  }
  return 0;
}

static std::atomic<long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
long launch_count() { return g_launches.load(std::memory_order_relaxed); }

bool pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("CVB_PDL");
    v = (e != nullptr && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

int ensure_dyn_smem_impl(const void* func, int bytes) {
  static std::mutex mu;
  static std::map<std::pair<int, const void*>, int> set_bytes;  // (device, kernel) -> attribute value in force
  int dev = 0;
  CVB_CUDA(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lk(mu);
  int& cur = set_bytes[{dev, func}];
  if (bytes > cur) {
    CVB_CUDA(cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    cur = bytes;
  }
  return 0;
}

int device_sm_count() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
  return sms;
}

}  // namespace cvb
