/* coverb200 - C ABI of the B200-native CoVer-VLA sample-and-verify path.
 *
 * The reference (cover-vla/cover-vla) has no FFI: its boundary for this path is two Python objects,
 *   PI0Policy.select_action / PI0FlowMatching.sample_actions
 *       lerobot_custom/lerobot/common/policies/pi0/modeling_pi0.py:263-307, :672-715
 *   EfficientEnsembleMerged.compute_max_similarity_scores_batch
 *       bridge_verifier/ensemble_eval/efficient_ensemble_merged.py:309-454
 * called from CoVer_VLA/inference/experiments/robot/simpler/run_simpler_eval_with_openpi.py:324,346-363.
 * This header is what a ctypes / cffi / pybind stub binds instead.
 *
 * Conventions: plain pointers and sizes only; every tensor is caller-owned DEVICE memory unless the
 * parameter name ends in _host; every call is asynchronous on `stream` (a cudaStream_t passed as
 * void*); return 0 = OK, negative = error with the text in cvb_last_error() (thread-local).
 * No call synchronises the device, allocates per call, or throws across the boundary.
 */
#ifndef COVERB200_H_
#define COVERB200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CVB_ABI_VERSION 1
#if defined(__GNUC__)
#define CVB_API __attribute__((visibility("default")))
#else
#define CVB_API
#endif

CVB_API const char* cvb_last_error(void);
CVB_API int cvb_abi_version(void);

/* ---------------------------------------------------------------------------------------------
 * Operator level (one launch each).  GEMM epilogue kinds:
 *   0 store bf16(acc+bias) | 1 bf16(gelu_tanh(bf16(acc+bias))) | 2 bf16(bf16(acc+bias)+resid)
 *   3 GeGLU on packed [128 gate | 128 up] weight rows | 4 store fp32(acc+bias)
 * Replaces nn.Linear (+ the elementwise op that follows it) wherever it appears on the path, e.g.
 * paligemma_with_expert.py:273-276 (q/k/v), :327-333 (o_proj + residual), :335-341 (MLP + residual).
 */
CVB_API int cvb_op_gemm_bf16(const void* A, int64_t lda, const void* W, int64_t ldw, int M, int N, int K,
                     int epilogue, void* C, int64_t ldc, const void* bias, int bias_is_f32,
                     const void* resid, int resid_is_f32, int64_t ldr, int n_out,
                     const int32_t* m_dev, int force_bn, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* COVERB200_H_ */
