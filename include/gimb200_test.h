/* gimb200_test.h - test / measurement hooks of libgimb200_test.so (NOT part of the product ABI).
 *
 * libgimb200_test.so is libgimb200.so plus the entry points below; it is loaded only by tests/test_umma_gpu.py and the
 * tools/ measurement scripts.  None of them has a reference counterpart. */
#ifndef GIMB200_TEST_H
#define GIMB200_TEST_H

#include "gimb200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ---- layer-level test hook
 * Runs ONE GEMM-shaped layer (NHWC conv ksize x ksize / Linear, optional channel concat `in2` for 1x1)
 * through both GEMM engines of the library on the same device buffers:
 *   out_simt        : fp32 CUDA-core implicit GEMM (conv_simt.cu)
 *   out_umma        : tcgen05 split-fp16 tensor-core GEMM, fp32 output (umma_gemm.cu)
 *   out_umma_planes : optional; the same result re-assembled from the fp16 (hi, lo) output planes
 * in [B,H,W,C1] (+ in2 [B,H,W,C2]); w [Cout, k, k, C1+C2]; outputs [B*OH*OW, Cout].                 */
int gimb_test_conv(const float* in, const float* in2, int B, int H, int W, int C1, int C2,
                   const float* w, int Cout, int ksize, int stride, const float* scale,
                   const float* bias, const float* residual, const uint8_t* row_mask, int act0,
                   int act1, int act_split, float div, float* out_umma, float* out_umma_planes,
                   float* out_simt, void* workspace, size_t workspace_bytes, void* stream);

/* Test/bench hook: average device time (ms) of `iters` back-to-back launches of the tcgen05 GEMM on one layer shape
 * (operands pre-split; flags: 1 = folded BN, 2 = residual, 4 = fp32 output, 8 = fp16-plane output). */
int gimb_bench_layer(int B, int H, int W, int C1, int C2, int Cout, int ksize, int stride, int flags, int act,
                     int iters, float* ms_out, void* stream);

/* Measurement helper (tools/probe_tma.py): aggregate L2 -> SM bulk-tensor load rate in GB/s for boxes with 64-byte
 * rows (variant 0: row mode, 2: conv patches) or 128-byte rows (1, 3). */
int gimb_probe_tma(int variant, int iters, float* gbps_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GIMB200_TEST_H */
