/* recmv_b200_diag.h -- diagnostics of the tcgen05 engine, kept OUT of the product ABI (include/recmv_b200.h).
 *
 *   recmv_tc_microbench    lives in recmv_b200/librecmv_b200_diag.so (built from csrc/tc_microbench.cu alone);
 *   recmv_sdf_mlp_tc_debug is exported by librecmv_b200.so (it launches the product kernel with its trace / raw-accumulator
 *                          hooks enabled) but is declared only here: no reference interface maps to it.
 * Used by tools/tc_microbench.py, tools/tc_bringup.py, tools/tc_trace.py, tools/calibrate_acc_gain.py.
 */
#ifndef RECMV_B200_DIAG_H_
#define RECMV_B200_DIAG_H_
#include "recmv_b200.h"
#ifdef __cplusplus
extern "C" {
#endif

/* Diagnostics for the tcgen05 path (used by tools/tc_bringup.py and tools/tc_trace.py): same computation as
 * recmv_sdf_mlp_fwd in a TC mode (passes = 1 or 3), plus status_host[4] = {code, barrier tag, block, 0}
 * of the kernel's bounded mbarrier waits (code 0 = no wait timed out) and, when dbg_out != NULL, the raw
 * fp32 accumulator (before bias) of layer dbg_layer for the first 128 points, [128][512].              */
RECMV_API int recmv_sdf_mlp_tc_debug(const float* x, const void* packed, const float* pe_w /*host*/,
                           float* out_sdf, float* out_feat, int64_t P, int passes, int dbg_layer,
                           float* dbg_out, int* status_host /*host*/,
                           unsigned long long* trace /*device [4][2][9][16] clock stamps or NULL*/,
                           recmv_stream_t stream);

/* Diagnostics: tcgen05 issue-rate microbenchmark (cycles per M x N x 16 kind::f16 MMA with smem operands). */
RECMV_API int recmv_tc_microbench(int cta_group, int M, int N, int iters, int num_ctas, int flags, const void* gsrc,
                        unsigned long long* out, recmv_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif  /* RECMV_B200_DIAG_H_ */
