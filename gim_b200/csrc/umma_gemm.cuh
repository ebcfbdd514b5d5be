// umma_gemm.cuh - host-side description of the tcgen05 split-fp16 GEMM / implicit-GEMM convolution.
//
//   D[M, N] = epi( A[M, K] * B[N, K]^T )        fp32-equivalent product on the 5th-gen tensor cores
//
// Operand format ("split fp16"): every fp32 value x is carried as two fp16 planes
//   hi = fp16(x),  lo = fp16((x - hi) * 2^8)
// One fp32 TMEM accumulator receives, per chunk of k,
//   D  = sum_k (A_hi * B_lo + A_lo * B_hi)          (2 x tcgen05.mma kind::f16 per k16 step)
//   D  = sum_k  A_hi * B_hi  +  2^-8 * D            (1 x per k16 step, the first with scale-input-d = 8)
// so that D = A*B up to the dropped lo*lo term (2^-22 relative) - the precision class the survey validated
// against the fp32 reference ("3 MMAs at the f16 rate").
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>

#include "common.cuh"

namespace gimb {

constexpr float kSplitScale = 256.f;  // 2^8

struct SplitPlanes {        // an fp32 tensor carried as fp16 planes, channel (K) stride `ld` elements
  __half* hi = nullptr;
  __half* lo = nullptr;
  __half* h8 = nullptr;     // only for tensors used as the B operand
  int ld = 0;               // row pitch in elements (multiple of 8)
};

struct UmmaGemm {
  // ---- A operand
  // mode 0 (rows):  A is [M, K1] (+ optional concat [M, K2]) row-major planes
  // mode 1 (conv):  A is NHWC [B, H, W, Cin] planes; output pixel tiles are TH x TW patches
  int mode = 0;
  SplitPlanes a, a2;
  int64_t M = 0;          // rows (mode 0)
  int K1 = 0, K2 = 0;     // mode 0: channels of a / a2 ; mode 1: K1 = Cin
  int B = 1, H = 1, W = 1, KH = 1, KW = 1, stride = 1, pad = 0, OH = 1, OW = 1;
  // ---- B operand: planes [N, Kw] with Kw = KH*KW*ldk (per-tap pitch ldk >= Cin, multiple of 8)
  SplitPlanes b;
  int N = 0;
  int ldk = 0;            // per-tap K pitch of the weight planes (mode 1); mode 0: unused
  // ---- epilogue:  v = acc * 2^-8 ; v = v*scale+bias ; v += residual ; v = act(v) ; v *= row_mask
  const float* scale = nullptr;
  const float* bias = nullptr;
  const float* residual = nullptr;   // [M, N] fp32, pitch N
  SplitPlanes residual_planes;       // alternatively the residual as fp16 planes (plane-only output variant)
  const uint8_t* row_mask = nullptr;
  int act0 = ACT_NONE, act1 = ACT_NONE, act_split = 1 << 30;
  float div = 1.f;
  bool layernorm = false;            // LayerNorm(eps 1e-5) of each output row BEFORE scale (= gamma) / bias (= beta) / residual
  float* out_f32 = nullptr;          // [M, N] fp32 (optional), row pitch out_f32_ld elements
  int out_f32_ld = 0;                // 0 = N.  A larger pitch (multiple of 4) gets its pad channels [N, ld) zero-filled
  int residual_ld = 0;               // row pitch of `residual` (0 = N)
  SplitPlanes out;                   // optional fp16 planes (hi, lo[, h8]); pad channels [N, ld) are zeroed
};

int umma_gemm(Ctx& ctx, const UmmaGemm& g);

// Coarse-matching correlation sweeps on the tensor cores (networks/loftr/utils/coarse_matching.py:111-118):
// pass 0 = per-tile softmax statistics (row / column partial (max, sum exp)), pass 1 = confidence + mutual-NN maxima.
struct UmmaCorr {
  SplitPlanes f0;   // [N*L, C] (hi, lo)
  SplitPlanes f1;   // [N*S, C] (hi, lo, h8)
  int N = 0, L = 0, S = 0, C = 0;
  const uint8_t* mask0 = nullptr;
  const uint8_t* mask1 = nullptr;
  float temperature = 0.1f, thr = 0.2f;
  float2* rowpart = nullptr;   // [N*L][row_parts]
  float2* colpart = nullptr;   // [N*S][col_parts]
  const float2* rowstat = nullptr;
  const float2* colstat = nullptr;
  unsigned long long* rowbest = nullptr;
  unsigned int* colbest = nullptr;
  float* conf_matrix = nullptr;  // optional debug tap [N, L, S]: every confidence, written by the conf sweep
};
void umma_corr_parts(int L, int S, int* row_parts, int* col_parts);
int umma_corr(Ctx& ctx, const UmmaCorr& c, int pass);

// fp32 -> split planes (weights at load time; test helper)
int split_planes(Ctx& ctx, const float* src, int64_t rows, int cols, int src_ld, const SplitPlanes& dst);

// measurement helper: aggregate L2 -> SM TMA load rate (GB/s); variant 0/1 = row boxes of 64/128-byte rows, 2/3 = conv boxes
int tma_probe(Ctx& ctx, int variant, int iters, float* gbps);

}  // namespace gimb
