// conv_simt.cu - fp32 CUDA-core implicit-GEMM convolution / Linear kernel + the small backbone helpers.
//
// This is the exact-fp32 GEMM path of the library (FFMA, fp32 accumulate): every GEMM-shaped layer of
// gim_loftr can run through it.  Layout is NHWC so that the GEMM K dimension (kh, kw, ci) is
// contiguous in ci; both operands are K-major ("TN").
//
// The tile mainloop lives in simt_tile.cuh.
#include <algorithm>

#include "ops.cuh"
#include "simt_tile.cuh"
#include "split.cuh"

namespace gimb {

namespace {

using namespace simt;

struct KParams {
  TileOperands t;
  const float* scale;
  const float* bias;
  const float* residual;
  const uint8_t* row_mask;
  int act0, act1, act_split;
  float div;
  float* out;
};

template <int BN>
__global__ void __launch_bounds__(NTHREADS, (BN == 128) ? 1 : 2) conv_gemm_kernel(const KParams p) {
  constexpr int TN = BN / 16;
  extern __shared__ __align__(16) float smem[];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int m0 = blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  float acc[8][TN];
  mainloop<BN>(p.t, m0, n0, smem, acc);

#pragma unroll
  for (int j = 0; j < TN; ++j) {
    int co = n0 + tx + 16 * j;
    if (co >= p.t.N) continue;
    float sc = p.scale ? p.scale[co] : 1.f;
    float bi = p.bias ? p.bias[co] : 0.f;
    int act = co >= p.act_split ? p.act1 : p.act0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      int row = m0 + ty + 16 * i;
      if (row >= p.t.M) continue;
      float v = acc[i][j];
      if (p.scale) v = fmaf(v, sc, bi);
      if (p.residual) v += p.residual[(size_t)row * p.t.N + co];
      v = apply_act(v, act, p.div);
      if (p.row_mask) v *= (float)p.row_mask[row];
      p.out[(size_t)row * p.t.N + co] = v;
    }
  }
}

// ------------------------------------------------------------------------------------------- stem
// 7x7 s2 p3, 3 -> 64 (networks/loftr/backbone/resnet.py:276-279 conv1 + bn1 + relu).  Persistent CTAs: the 64 x 147
// weights are staged once per CTA, then it walks 8 x 64 output tiles.  Each thread owns two pixels (columns px and
// px + 32 of its row) x all 64 channels: 128 FMAs per 16 broadcast weight loads, so the loop is FMA-bound instead of
// shared-memory-bound.  The input patch is stored de-interleaved by column parity (stride-2 taps would otherwise hit
// the banks two ways).  Per output the summation order is (kh, kw, ci), one fmaf chain.
constexpr int ST_TH = 8, ST_TW = 64;
constexpr int ST_PH = ST_TH * 2 + 5;          // 21 input rows
constexpr int ST_PWH = ST_TW + 3;             // 67 columns of each parity (133 input columns)
constexpr int ST_PITCH = ST_PWH + 1;

__global__ void __launch_bounds__(256, 1) stem_kernel(const float* __restrict__ in, int B, int H, int W,
                                                      const float* __restrict__ w, const float* __restrict__ scale,
                                                      const float* __restrict__ bias, float* __restrict__ out,
                                                      const PlanesDev sp, int tiles_x, int tiles_y) {
  extern __shared__ __align__(16) float st_smem[];
  float (*wsm)[64] = reinterpret_cast<float (*)[64]>(st_smem);                       // [tap*3+ci][co]
  float (*patch)[ST_PH][2][ST_PITCH] = reinterpret_cast<float (*)[ST_PH][2][ST_PITCH]>(st_smem + 147 * 64);  // [3]
  const int OH = H / 2, OW = W / 2;
  const int tid = threadIdx.x;
  for (int i = tid; i < 147 * 64; i += 256) {
    const int co = i / 147, k = i - co * 147;  // coalesced global read; the transposing smem write happens once per CTA
    wsm[k][co] = w[i];
  }
  const int px = tid & 31;   // output columns px and px + 32 of the tile
  const int py = tid >> 5;   // output row within the tile
  const int n_tiles = B * tiles_y * tiles_x;
  for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
    const int b = t / (tiles_y * tiles_x);
    const int r0 = t - b * (tiles_y * tiles_x);
    const int ty = r0 / tiles_x, tx = r0 - ty * tiles_x;
    const int oh0 = ty * ST_TH, ow0 = tx * ST_TW;
    const int ih0 = oh0 * 2 - 3, iw0 = ow0 * 2 - 3;
    __syncthreads();  // previous tile's patch fully consumed (and, first time, the weights staged)
    for (int i = tid; i < 3 * ST_PH * (2 * ST_PWH); i += 256) {
      const int c = i / (ST_PH * 2 * ST_PWH);
      const int r = i - c * (ST_PH * 2 * ST_PWH);
      const int y = r / (2 * ST_PWH), x = r - y * (2 * ST_PWH);
      const int ih = ih0 + y, iw = iw0 + x;
      float v = 0.f;
      if (ih >= 0 && ih < H && iw >= 0 && iw < W) v = in[(((size_t)b * 3 + c) * H + ih) * W + iw];
      patch[c][y][x & 1][x >> 1] = v;
    }
    __syncthreads();
    float acc[2][64];
#pragma unroll
    for (int c = 0; c < 64; ++c) { acc[0][c] = 0.f; acc[1][c] = 0.f; }
    for (int kh = 0; kh < 7; ++kh)
#pragma unroll
      for (int kw = 0; kw < 7; ++kw) {
#pragma unroll
        for (int ci = 0; ci < 3; ++ci) {
          // input column 2 * px + kw -> parity kw & 1, index px + kw / 2
          const float v0 = patch[ci][py * 2 + kh][kw & 1][px + (kw >> 1)];
          const float v1 = patch[ci][py * 2 + kh][kw & 1][px + 32 + (kw >> 1)];
          const float4* wr = reinterpret_cast<const float4*>(&wsm[(kh * 7 + kw) * 3 + ci][0]);
#pragma unroll
          for (int c4 = 0; c4 < 16; ++c4) {
            const float4 ww = wr[c4];
            acc[0][c4 * 4 + 0] = fmaf(v0, ww.x, acc[0][c4 * 4 + 0]);
            acc[0][c4 * 4 + 1] = fmaf(v0, ww.y, acc[0][c4 * 4 + 1]);
            acc[0][c4 * 4 + 2] = fmaf(v0, ww.z, acc[0][c4 * 4 + 2]);
            acc[0][c4 * 4 + 3] = fmaf(v0, ww.w, acc[0][c4 * 4 + 3]);
            acc[1][c4 * 4 + 0] = fmaf(v1, ww.x, acc[1][c4 * 4 + 0]);
            acc[1][c4 * 4 + 1] = fmaf(v1, ww.y, acc[1][c4 * 4 + 1]);
            acc[1][c4 * 4 + 2] = fmaf(v1, ww.z, acc[1][c4 * 4 + 2]);
            acc[1][c4 * 4 + 3] = fmaf(v1, ww.w, acc[1][c4 * 4 + 3]);
          }
        }
      }
    const int oh = oh0 + py;
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int ow = ow0 + px + 32 * e;
      if (oh < OH && ow < OW) {
        const size_t pix = ((size_t)b * OH + oh) * OW + ow;
        float4* o = reinterpret_cast<float4*>(out + pix * 64);
#pragma unroll
        for (int c4 = 0; c4 < 16; ++c4) {
          const float4 sc = __ldg(reinterpret_cast<const float4*>(scale) + c4);
          const float4 bi = __ldg(reinterpret_cast<const float4*>(bias) + c4);
          float4 r;
          r.x = fmaxf(fmaf(acc[e][c4 * 4 + 0], sc.x, bi.x), 0.f);
          r.y = fmaxf(fmaf(acc[e][c4 * 4 + 1], sc.y, bi.y), 0.f);
          r.z = fmaxf(fmaf(acc[e][c4 * 4 + 2], sc.z, bi.z), 0.f);
          r.w = fmaxf(fmaf(acc[e][c4 * 4 + 3], sc.w, bi.w), 0.f);
          if (out) o[c4] = r;
          if (sp.hi) split4_store(sp, pix * sp.ld + c4 * 4, r.x, r.y, r.z, r.w);
        }
      }
    }
  }
}

// --------------------------------------------------------------------------- bilinear 2x + add
// F.interpolate(scale_factor=2, mode='bilinear', align_corners=True): src = dst * (in-1)/(out-1).
__global__ void upsample2x_add_kernel(const float* __restrict__ low, int B, int h, int w, int C4,
                                      float* __restrict__ out, float ry, float rx, const PlanesDev sp) {
  const int OH = 2 * h, OW = 2 * w;
  size_t total = (size_t)B * OH * OW * C4;
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total;
       idx += (size_t)gridDim.x * blockDim.x) {
    int c = idx % C4;
    size_t pix = idx / C4;
    int ox = pix % OW;
    size_t t = pix / OW;
    int oy = t % OH;
    int b = t / OH;
    float sy = ry * oy, sx = rx * ox;
    int y0 = (int)sy, x0 = (int)sx;
    int y1 = min(y0 + 1, h - 1), x1 = min(x0 + 1, w - 1);
    float ly = sy - y0, lx = sx - x0;
    float hy = 1.f - ly, hx = 1.f - lx;
    const float4* L = reinterpret_cast<const float4*>(low) + (size_t)b * h * w * C4;
    float4 v00 = L[((size_t)y0 * w + x0) * C4 + c], v01 = L[((size_t)y0 * w + x1) * C4 + c];
    float4 v10 = L[((size_t)y1 * w + x0) * C4 + c], v11 = L[((size_t)y1 * w + x1) * C4 + c];
    float4* o = reinterpret_cast<float4*>(out) + idx;
    float4 r = *o;
    // same operation order as ATen's upsample_bilinear2d: hy*(hx*v00 + lx*v01) + ly*(hx*v10 + lx*v11)
    r.x += hy * (hx * v00.x + lx * v01.x) + ly * (hx * v10.x + lx * v11.x);
    r.y += hy * (hx * v00.y + lx * v01.y) + ly * (hx * v10.y + lx * v11.y);
    r.z += hy * (hx * v00.z + lx * v01.z) + ly * (hx * v10.z + lx * v11.z);
    r.w += hy * (hx * v00.w + lx * v01.w) + ly * (hx * v10.w + lx * v11.w);
    if (sp.hi) {
      split4_store(sp, pix * sp.ld + c * 4, r.x, r.y, r.z, r.w);
      // zero the pad channels [C, ld) once per pixel
      if (c == C4 - 1)
        for (int z = C4 * 4; z < sp.ld; z += 4) split4_store(sp, pix * sp.ld + z, 0.f, 0.f, 0.f, 0.f);
    } else {
      *o = r;
    }
  }
}

__global__ void add_pe_kernel(const float4* __restrict__ feat, const float4* __restrict__ pe, size_t total,
                              size_t per_image, float4* __restrict__ out, const PlanesDev sp, int C4) {
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total;
       idx += (size_t)gridDim.x * blockDim.x) {
    float4 a = feat[idx], b = pe[idx % per_image];
    float4 r = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
    out[idx] = r;
    if (sp.hi) {
      size_t row = idx / C4;
      int c = (int)(idx - row * C4);
      split4_store(sp, row * sp.ld + c * 4, r.x, r.y, r.z, r.w);
    }
  }
}

}  // namespace

int conv_gemm(Ctx& ctx, const ConvGemm& c) {
  KParams p;
  TileOperands& t = p.t;
  t.in = c.in; t.in2 = c.in2;
  t.H = c.H; t.W = c.W; t.C1 = c.C1; t.C2 = c.C2; t.Cin = c.C1 + c.C2;
  t.KH = c.KH; t.KW = c.KW; t.stride = c.stride; t.pad = c.pad; t.OH = c.OH; t.OW = c.OW;
  t.w = c.w; t.N = c.Cout; t.K = c.KH * c.KW * t.Cin; t.M = c.B * c.OH * c.OW;
  p.scale = c.scale; p.bias = c.bias; p.residual = c.residual; p.row_mask = c.row_mask;
  p.act0 = c.act0; p.act1 = c.act1; p.act_split = c.act_split; p.div = c.div; p.out = c.out;
  GIMB_CHECK(t.Cin % 4 == 0 && c.C1 % 4 == 0, "conv_gemm: channel counts must be multiples of 4 (C1=%d C2=%d)", c.C1, c.C2);
  GIMB_CHECK(c.in2 == nullptr || (c.KH == 1 && c.KW == 1), "conv_gemm: channel concat only for 1x1");
  GIMB_CHECK((c.scale == nullptr) == (c.bias == nullptr), "conv_gemm: scale and bias go together");
  if (ctx.dry || t.M == 0) return 0;
  GIMB_SMEM_OPTIN(conv_gemm_kernel<128>, smem_bytes<128>());
  GIMB_SMEM_OPTIN(conv_gemm_kernel<64>, smem_bytes<64>());
  if (t.N <= 64) {
    dim3 grid(cdiv(t.M, BM), cdiv(t.N, 64));
    conv_gemm_kernel<64><<<grid, NTHREADS, smem_bytes<64>(), ctx.stream>>>(p);
  } else {
    dim3 grid(cdiv(t.M, BM), cdiv(t.N, 128));
    conv_gemm_kernel<128><<<grid, NTHREADS, smem_bytes<128>(), ctx.stream>>>(p);
  }
  ctx.launches++;
  GIMB_LAUNCH_CHECK();
  return 0;
}

int stem_conv7x7(Ctx& ctx, const float* in_nchw, int B, int H, int W, const float* w, const float* scale,
                 const float* bias, float* out_nhwc, const SplitPlanes* planes) {
  if (ctx.dry) return 0;
  const PlanesDev sp = planes ? dev(*planes) : PlanesDev{nullptr, nullptr, nullptr, 0};
  const int tiles_x = cdiv(W / 2, ST_TW), tiles_y = cdiv(H / 2, ST_TH);
  const int grid = std::min(B * tiles_x * tiles_y, ctx.sm_count);
  const int smem = (147 * 64 + 3 * ST_PH * 2 * ST_PITCH) * (int)sizeof(float);
  GIMB_SMEM_OPTIN(stem_kernel, smem);
  stem_kernel<<<grid, 256, smem, ctx.stream>>>(in_nchw, B, H, W, w, scale, bias, out_nhwc, sp, tiles_x, tiles_y);
  ctx.launches++;
  GIMB_LAUNCH_CHECK();
  return 0;
}

int upsample2x_add(Ctx& ctx, const float* low, int B, int h, int w, int C, float* out, const SplitPlanes* planes) {
  GIMB_CHECK(C % 4 == 0, "upsample2x_add: C %% 4 != 0");
  if (ctx.dry) return 0;
  const PlanesDev sp = planes ? dev(*planes) : PlanesDev{nullptr, nullptr, nullptr, 0};
  GIMB_CHECK(!planes || (planes->ld % 4 == 0 && planes->ld >= C), "upsample2x_add: bad plane pitch");
  float ry = (2 * h > 1) ? (float)(h - 1) / (float)(2 * h - 1) : 0.f;
  float rx = (2 * w > 1) ? (float)(w - 1) / (float)(2 * w - 1) : 0.f;
  size_t total = (size_t)B * 4 * h * w * (C / 4);
  int blocks = (int)std::min<size_t>((total + 255) / 256, (size_t)ctx.sm_count * 16);
  upsample2x_add_kernel<<<blocks, 256, 0, ctx.stream>>>(low, B, h, w, C / 4, out, ry, rx, sp);
  ctx.launches++;
  GIMB_LAUNCH_CHECK();
  return 0;
}

int add_pe(Ctx& ctx, const float* feat, const float* pe, int B, int L, int C, float* tokens, const SplitPlanes* planes) {
  GIMB_CHECK(C % 4 == 0, "add_pe: C %% 4 != 0");
  if (ctx.dry) return 0;
  const PlanesDev sp = planes ? dev(*planes) : PlanesDev{nullptr, nullptr, nullptr, 0};
  size_t per = (size_t)L * C / 4, total = per * B;
  int blocks = (int)std::min<size_t>((total + 255) / 256, (size_t)ctx.sm_count * 16);
  add_pe_kernel<<<blocks, 256, 0, ctx.stream>>>((const float4*)feat, (const float4*)pe, total, per, (float4*)tokens, sp, C / 4);
  ctx.launches++;
  GIMB_LAUNCH_CHECK();
  return 0;
}

}  // namespace gimb
