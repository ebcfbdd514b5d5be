// ops.cuh - host-side launchers of every kernel in libgimb200 (declarations).
#pragma once
#include "common.cuh"
#include "umma_gemm.cuh"

namespace gimb {

// ---------------------------------------------------------------------------------------------
// GEMM-shaped op: NHWC implicit-GEMM convolution / Linear layer with fused epilogue.
//   out[p, co] = epi( sum_{kh,kw,ci} in[n, oh*s+kh-pad, ow*s+kw-pad, ci] * w[co, kh, kw, ci] )
//   epi(v)     = mask[p] * act( v * scale[co] + bias[co] + residual[p, co] )
// A Linear layer is the 1x1 case with B=1, H=rows, W=1.  `in2` concatenates a second NHWC tensor
// along channels (1x1 only): used for mlp.0(cat[x, msg]) (networks/loftr/submodules/transformer.py:54).
struct ConvGemm {
  const float* in = nullptr;
  const float* in2 = nullptr;
  int B = 1, H = 1, W = 1, C1 = 0, C2 = 0;
  int KH = 1, KW = 1, stride = 1, pad = 0;
  int OH = 1, OW = 1;
  const float* w = nullptr;  // [Cout][KH*KW*(C1+C2)]
  int Cout = 0;
  const float* scale = nullptr;     // [Cout] or null (folded BatchNorm)
  const float* bias = nullptr;      // [Cout] or null
  const float* residual = nullptr;  // [M, Cout] or null
  const uint8_t* row_mask = nullptr;  // [M] 0/1 or null
  int act0 = ACT_NONE, act1 = ACT_NONE, act_split = 1 << 30;  // columns >= act_split use act1
  float div = 1.f;                                           // ACT_DIVS divisor
  float* out = nullptr;                                      // [M, Cout]
};
int conv_gemm(Ctx& ctx, const ConvGemm& p);

// stem: 7x7 stride-2 pad-3 conv 3->64 + folded BN + ReLU; NCHW fp32 in, NHWC out
// (networks/loftr/backbone/resnet.py:158,230).
int stem_conv7x7(Ctx& ctx, const float* in_nchw, int B, int H, int W, const float* w /*[64][7][7][3]*/,
                 const float* scale, const float* bias, float* out_nhwc, const SplitPlanes* planes = nullptr);

// out[b, y, x, c] += bilinear_2x(align_corners=True)(low)[b, y, x, c]   (resnet.py:321-327)
// with `planes` the sum is written as split fp16 planes (pad channels zeroed) instead of back into `out`
int upsample2x_add(Ctx& ctx, const float* low, int B, int h, int w, int C, float* out /*[B,2h,2w,C]*/,
                   const SplitPlanes* planes = nullptr);

// tokens[b, l, c] = feat[b, l, c] + pe[l, c]   (loftr.py:74-75, position_encoding.py:43)
int add_pe(Ctx& ctx, const float* feat, const float* pe, int B, int L, int C, float* tokens,
           const SplitPlanes* planes = nullptr);

// ---------------------------------------------------------------------------------------------
// LayerNorm over the last dim (eps 1e-5) with optional residual: out = (res ? res : 0) + LN(x)
// `out` (fp32) and `planes` (split fp16, optionally with the h8 plane) are both optional sinks.
int layernorm(Ctx& ctx, const float* x, const float* gamma, const float* beta, const float* res,
              int64_t rows, int C, float* out, const SplitPlanes* planes = nullptr);

// Linear attention (networks/loftr/submodules/attentions.py:31-47), coarse flavour: D = 32 per head.
//   kv [B, S, 2C]: columns [0,C) = K' = elu(k)+1 (masked), [C,2C) = V/S (masked)
//   q  [B, L, C]  = Q' = elu(q)+1 (masked)
//   msg[b, l, h*D+v] = (sum_d Q'[l,h,d] KV[h,d,v]) / (sum_d Q'[l,h,d] Ksum[h,d] + 1e-6) * S
int linear_attention(Ctx& ctx, const float* q, const float* kv, int B, int L, int S, int C, int nhead,
                     float* msg, const SplitPlanes* planes = nullptr);

// ---------------------------------------------------------------------------------------------
// coarse matching (networks/loftr/utils/coarse_matching.py:88-259)
struct CoarseMatchArgs {
  const float* f0;  // [N, L, C] after the coarse transformer
  const float* f1;  // [N, S, C]
  int N, L, S, C;
  int h0c, w0c, h1c, w1c;
  int H0, H1;                   // input image heights (scale = H / hc)
  const uint8_t* mask0;         // [N, L] or null
  const uint8_t* mask1;         // [N, S] or null
  const float* scale0;          // [N, 2] or null
  const float* scale1;
  float thr, temperature;
  int border;
  // outputs (device)
  int64_t *b_ids, *i_ids, *j_ids;
  float *mconf, *mkpts0_c, *mkpts1_c;
  int64_t* count;        // device scalar
  int64_t capacity = 0;  // rows available in the id / mconf / mkpts arrays (matches beyond it are counted, not written)
  float* conf_matrix;    // optional [N, L, S] debug tap, written by whichever conf sweep runs
  // split-fp16 planes of f0 (hi, lo) / f1 (hi, lo, h8): when both are given the sweeps run on the tensor cores
  const SplitPlanes* planes0 = nullptr;
  const SplitPlanes* planes1 = nullptr;
  // tcgen05 engine: the fast sweeps (corr_sweep.cu) set *range_flag (device int) when a softmax sum left the range in
  // which their fixed exponent reference is exact; the caller then repeats the call with exact = true (online-max sweeps)
  int* range_flag = nullptr;
  bool exact = false;
};
int coarse_match(Ctx& ctx, const CoarseMatchArgs& a);

// ---------------------------------------------------------------------------------------------
// fine level (fine_preprocess.py:29-47, transformer.py with d=128, fine_matching.py:43-72)
// gather W x W windows around stride*cell centres: out [M, WW, C]
int fine_gather(Ctx& ctx, const float* feat_f /*[N,hf,wf,C]*/, int hf, int wf, int C, int wc /*coarse width*/,
                int stride, int Wn, const int64_t* b_ids, const int64_t* ids, int64_t m0, int64_t m,
                float* out, const SplitPlanes* planes = nullptr);
// per-match linear attention with D = 16 per head, sequence WW (<= 32)
int fine_attention(Ctx& ctx, const float* q, const float* kv, int64_t M, int WW, int C, int nhead, float* msg,
                   const SplitPlanes* planes = nullptr);
struct FineMatchArgs {
  const float* f0;  // [m, WW, C]
  const float* f1;
  int64_t m0, m;    // chunk offset / size
  int WW, C, Wn;
  float fscale;     // hw0_i[0] / hw0_f[0]
  float sim_scale;  // 1 / sqrt(C), rounded from double like the reference's python float
  const int64_t* b_ids;
  const float* scale1;  // [N,2] or null (applied iff scale0 is in data, fine_matching.py:68)
  const float* mkpts0_c;
  const float* mkpts1_c;
  float *mkpts0_f, *mkpts1_f, *expec_f;
};
int fine_match(Ctx& ctx, const FineMatchArgs& a);

}  // namespace gimb
