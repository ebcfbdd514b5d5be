// dkm_ops.cuh - launchers of the gim_dkm (DKMv3) kernels that are not GEMM-shaped layers (those run on the shared tcgen05
// engine, engine.cuh).  All activations are NHWC fp32 with an explicit channel pitch `ld` (>= C, pad channels are never
// read); flows are [B, h, w, 2] = (x, y) in normalised [-1, 1] coordinates like the reference.
#pragma once
#include "common.cuh"
#include "umma_gemm.cuh"

namespace gimb {

// F.interpolate(mode='bilinear', align_corners=False) - NCHW images (networks/dkm/models/dkm.py:670-671, 697-698)
int dkm_resize_nchw(Ctx& ctx, const float* in, int B, int C, int H, int W, float* out, int OH, int OW);
// the same for NHWC tensors (flow / certainty / DFN context between scales, dkm.py:423-428, 471-482, 517-530)
int dkm_resize_nhwc(Ctx& ctx, const float* in, int B, int H, int W, int C, int ld_in, float* out, int OH, int OW, int ld_out);
// nn.MaxPool2d(3, 2, 1) of torchvision's ResNet stem (networks/dkm/models/encoders.py:53)
int dkm_maxpool3x3s2(Ctx& ctx, const float* in, int B, int H, int W, int C, float* out, const SplitPlanes* planes);
// dst[r, c_off + c] = src[r, c]  (torch.cat along channels); optional split planes of the destination slice
int dkm_copy_channels(Ctx& ctx, const float* src, int64_t rows, int C, int ld_src, float* dst, int ld_dst, int c_off);
int dkm_fill(Ctx& ctx, float* dst, size_t n, float v);
int dkm_zero_pad_channels(Ctx& ctx, float* x, int64_t rows, int C, int ld);  // x[r, C .. ld) = 0
// x = hi + lo * 2^-8 (split planes back to fp32) ; NCHW -> NHWC with channel pitch ld (pad channels zeroed)
int planes_to_f32(Ctx& ctx, const SplitPlanes& sp, int64_t rows, int C, float* out, int ld_out);
int nchw_to_nhwc(Ctx& ctx, const float* in, int B, int C, int H, int W, float* out, int ld);
// placeholder flow = the pixel-centre grid (dkm.py:439-451)
int dkm_grid_flow(Ctx& ctx, float* flow, int B, int h, int w);

// ---- GP (dkm.py:126-144, 324-370): cosine-kernel Gram matrices, Fourier position basis, SPD solve
// K[b, i, j] = exp((<x_i, y_j> / (|x_i||y_j| + 1e-6) - 1) / T) (+ sigma on the diagonal when add_diag != 0)
int dkm_cos_gram(Ctx& ctx, const float* x, const float* y, int B, int N, int M, int C, int ld, float T, float add_diag, float* K);
// the same from a precomputed dot-product matrix D [B, N, M] (tensor-core Gram of the big scales), in place
int dkm_cos_gram_finish(Ctx& ctx, float* D, const float* x, const float* y, int B, int N, int M, int C, int ld, float T, float add_diag);
// f[b, n, d] = cos(8 pi (w[d,0] gx + w[d,1] gy + bias[d])) on the h x w pixel-centre grid
int dkm_pos_basis(Ctx& ctx, const float* w /*[D,2]*/, const float* bias, int B, int h, int w_, int D, float* f);
// solve (A) Z = F in place for SPD A [B, N, N] (destroyed: lower Cholesky factor) and F [B, N, D] (-> Z); blocked Cholesky
int dkm_chol_solve(Ctx& ctx, float* A, float* F, int B, int N, int D);
// C[b] = A[b] (N x K) * Bm[b] (K x D), row major, fp32 CUDA cores; out pitch ld_out
int dkm_matmul_nn(Ctx& ctx, const float* A, const float* Bm, int B, int N, int K, int D, float* C, int ld_out);

// ---- DFN pieces (dkm.py:147-170): channel attention
// pooled[b, c] = mean over pixels of cat(x1, x2)[b, :, c]
int dkm_cab(Ctx& ctx, const float* x1, const float* x2, int B, int HW, int C, const float* w1 /*[C,2C]*/, const float* b1,
            const float* w2 /*[C,C]*/, const float* b2, float* out /*[B,HW,C]: s * x2 + x1*/, float* scratch /*[B,3C]*/);

// ---- ConvRefiner pieces (dkm.py:75-123, utils/local_correlation.py)
// F.grid_sample(y, flow, bilinear, zeros, align_corners=False): out[b, p, c_off + c] = y sampled at flow[b, p]
int dkm_grid_sample(Ctx& ctx, const float* y, int B, int h, int w, int C, int ld_y, const float* flow, float* out, int ld_out, int c_off);
// emb[b, p, c_off + e] = w[e,0] (flow.x - gx) + w[e,1] (flow.y - gy) + bias[e]
int dkm_disp_emb(Ctx& ctx, const float* flow, int B, int h, int w, const float* wt /*[E,2]*/, const float* bias, int E, float* out,
                 int ld_out, int c_off);
// corr[b, p, c_off + k] = <x[b, p], y sampled at flow[b, p] + window offset k> / sqrt(C), (2r+1)^2 offsets of one pixel
int dkm_local_corr(Ctx& ctx, const float* x, const float* y, int B, int h, int w, int C, int ld, const float* flow, int r, float* out,
                   int ld_out, int c_off);
// depthwise 5x5 (channel multiplier `mult`) + folded BatchNorm + ReLU; fp32 and / or split planes out
int dkm_depthwise5x5(Ctx& ctx, const float* in, int B, int h, int w, int Cin, int ld_in, int mult, const float* wt /*[Cout,25]*/,
                     const float* scale, const float* bias, float* out, int ld_out, const SplitPlanes* planes);
// the same for channel multiplier 1 or 2 (C = OUTPUT channels) with transposed, padded parameters (wt_t [25][Cp], scale_p / bias_p [Cp], Cp = plane pitch):
// 4 channels x (4 x 2) pixels per thread, float4 loads
int dkm_depthwise5x5_v4(Ctx& ctx, const float* in, int B, int h, int w, int C, int ld_in, const float* wt_t, const float* scale_p,
                        const float* bias_p, int Cp, float* out, int ld_out, const SplitPlanes* planes, int mult = 1);
// flow += ins * disp / (4w, 4h); certainty (+)= delta   (dkm.py:501-510); head = [B, hw, ld_head] with (certainty, dx, dy)
int dkm_apply_delta(Ctx& ctx, float* flow, float* certainty, bool cert_accumulate, const float* head, int ld_head, int B, int hs, int ws,
                    float ins, int W, int H);
// flow / certainty from the DFN head: [certainty, x, y] (dkm.py:251-253)
int dkm_split_head(Ctx& ctx, const float* head, int ld_head, int64_t rows, float* flow, float* certainty);

// ---- match() tail (dkm.py:684-752)
struct DkmFinalArgs {
  const float* flow;       // [2, hs, ws, 2] query->support, support->query
  const float* certainty;  // [2, hs, ws]
  const float* low_cert;   // [2, hs, ws] scale-16 certainty resized to (hs, ws), or null
  const float* im1;        // [3, H1, W1] NCHW (black-pixel masks)
  const float* im2;
  int H1, W1, H2, W2, hs, ws;
  float* warp;             // [hs, 2 ws, 4]
  float* cert_out;         // [hs, 2 ws]
};
int dkm_finalize(Ctx& ctx, const DkmFinalArgs& a);
// density[i] = sum_j exp(-|x_i - x_j|^2 / (2 std^2)) for n 4-D points (the balanced sampling of dkm.py:612-619)
int dkm_kde(Ctx& ctx, const float* x /*[n,4], 16-byte aligned*/, int n, float std, float* density);

}  // namespace gimb
