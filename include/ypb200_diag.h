/* ypb200_diag — diagnostics on top of include/ypb200.h, exported only by libypb200_diag.so.
 *
 * The product library (libypb200.so) carries the product kernels and nothing else.  Building the same translation
 * unit with -DYPB_DIAG=1 (python -m yolo_puncture_b200.build --diag) adds the debugging twins of the conv kernel
 * (ypb_set_conv_impl / ypb_conv2d_bf16 impl 1-3), the per-layer conv timer and the TMA / tensor-pipe / latency
 * micro-benchmarks declared here.  Used by tests/test_gpu_conv.py (twins) and tools/ (measurements quoted in DESIGN.md).
 */
#ifndef YPB200_DIAG_H_
#define YPB200_DIAG_H_

#include "ypb200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Diagnostics: average milliseconds of `iters` back-to-back launches of one conv (out_mode 0 bf16, 1 fp32, 2 pixel-shuffle
   bf16); dbg >= 0 overrides the experiment mask; desc receives a description of the launch the planner chose. */
int ypb_conv_bench(void* cuda_stream, const void* in_nhwc_bf16, int B, int H, int W, int in_ctot, int in_c_off, int cin,
                   const void* w_gemm_bf16, const float* bias, int cout, int k, int stride, int act,
                   const void* res_nhwc_bf16, void* out, int out_ctot, int out_c_off, int out_mode, int impl, int dbg,
                   int iters, float* ms, char* desc, int desc_len);
/* Diagnostics: ceiling of the TMA operand-fetch path for an access pattern (csrc/tma_bench.cuh). */
int ypb_mma_bench(void* buf, int rows, int n, int iters, int shifted, int tma_iters, float* ms);
int ypb_latency_probe(long long* out_dev /* 8 x int64, device */);
int ypb_debug_prof(unsigned long long* out16, int reset);  /* profiling build: device-side cycle counters */
int ypb_tma_bench(void* buf, int mode, int stages, int iters, int rows, int W, int H, int B, float* ms, double* bytes);


#ifdef __cplusplus
}
#endif
#endif /* YPB200_DIAG_H_ */
