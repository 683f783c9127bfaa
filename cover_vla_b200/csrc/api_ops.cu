// C-ABI entry points for the individual operators (used by the parity tests and by hosts that want
// to compose their own graph).  Signatures are declared in include/coverb200.h.
#include "../../include/coverb200.h"
#include "gemm_tcgen05.cuh"
#include "host_common.h"
#include "ops.h"

namespace cvb {
const char* get_last_error();
}

extern "C" {

const char* cvb_last_error(void) { return cvb::get_last_error(); }

int cvb_abi_version(void) { return CVB_ABI_VERSION; }

int64_t cvb_launch_count(void) { return cvb::launch_count(); }

int cvb_op_gemm_bf16(const void* A, int64_t lda, const void* W, int64_t ldw, int M, int N, int K,
                     int epilogue, void* C, int64_t ldc, const void* bias, int bias_is_f32,
                     const void* resid, int resid_is_f32, int64_t ldr, int n_out,
                     const int32_t* m_dev, int force_bn, void* stream) {
  cvb::GemmCall c;
  c.A = (const cvb::bf16*)A;
  c.lda = lda;
  c.W = (const cvb::bf16*)W;
  c.ldw = ldw;
  c.M = M;
  c.N = N;
  c.K = K;
  c.epi = epilogue;
  c.C = C;
  c.ldc = ldc;
  c.bias = bias;
  c.bias_is_f32 = bias_is_f32;
  c.resid = resid;
  c.resid_is_f32 = resid_is_f32;
  c.ldr = ldr;
  c.n_out = n_out;
  c.m_dev = m_dev;
  c.force_bn = force_bn;
  return cvb::gemm_bf16((cudaStream_t)stream, c);
}

}  // extern "C"
