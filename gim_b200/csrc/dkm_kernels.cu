// dkm_kernels.cu - the gim_dkm (DKMv3) kernels that are not GEMM-shaped layers (see dkm_ops.cuh).  Reference:
// /root/reference/networks/dkm/models/dkm.py, networks/dkm/utils/local_correlation.py (file:line at each kernel).
#include <math.h>

#include <algorithm>

#include "dkm_ops.cuh"

namespace gimb {
namespace {

constexpr float kPi = 3.14159265358979323846f;

__device__ __forceinline__ float grid_coord(int i, int n) {  // linspace(-1 + 1/n, 1 - 1/n, n)[i]
  return (2.f * (float)i + 1.f) / (float)n - 1.f;
}
// source index / weight of upsample_bilinear2d(align_corners=False) (ATen: area_pixel_compute_source_index)
__device__ __forceinline__ void bil_src(int dst, float scale, int in_size, int& i0, int& i1, float& l1) {
  float s = scale * ((float)dst + 0.5f) - 0.5f;
  if (s < 0.f) s = 0.f;
  i0 = (int)s;
  if (i0 > in_size - 1) i0 = in_size - 1;
  i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
  l1 = s - (float)i0;
}

__global__ void resize_nchw_kernel(const float* __restrict__ in, int BC, int H, int W, float* __restrict__ out, int OH, int OW) {
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long total = (long long)BC * OH * OW;
  if (idx >= total) return;
  const int ox = (int)(idx % OW), oy = (int)((idx / OW) % OH);
  const long long bc = idx / ((long long)OW * OH);
  int y0, y1, x0, x1;
  float ly, lx;
  bil_src(oy, (float)H / (float)OH, H, y0, y1, ly);
  bil_src(ox, (float)W / (float)OW, W, x0, x1, lx);
  const float* p = in + bc * (long long)H * W;
  const float v00 = p[(long long)y0 * W + x0], v01 = p[(long long)y0 * W + x1];
  const float v10 = p[(long long)y1 * W + x0], v11 = p[(long long)y1 * W + x1];
  out[idx] = (1.f - ly) * ((1.f - lx) * v00 + lx * v01) + ly * ((1.f - lx) * v10 + lx * v11);
}

__global__ void resize_nhwc_kernel(const float* __restrict__ in, int B, int H, int W, int C, int ld_in, float* __restrict__ out, int OH,
                                   int OW, int ld_out) {
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long total = (long long)B * OH * OW * C;
  if (idx >= total) return;
  const int c = (int)(idx % C);
  const long long pix = idx / C;
  const int ox = (int)(pix % OW), oy = (int)((pix / OW) % OH), b = (int)(pix / ((long long)OW * OH));
  int y0, y1, x0, x1;
  float ly, lx;
  bil_src(oy, (float)H / (float)OH, H, y0, y1, ly);
  bil_src(ox, (float)W / (float)OW, W, x0, x1, lx);
  const float* p = in + (long long)b * H * W * ld_in + c;
  const float v00 = p[((long long)y0 * W + x0) * ld_in], v01 = p[((long long)y0 * W + x1) * ld_in];
  const float v10 = p[((long long)y1 * W + x0) * ld_in], v11 = p[((long long)y1 * W + x1) * ld_in];
  out[pix * ld_out + c] = (1.f - ly) * ((1.f - lx) * v00 + lx * v01) + ly * ((1.f - lx) * v10 + lx * v11);
}

__global__ void maxpool_kernel(const float* __restrict__ in, int B, int H, int W, int C, float* __restrict__ out, __half* hi, __half* lo,
                               int ldp) {
  const int OH = (H + 1) / 2, OW = (W + 1) / 2;  // floor((H + 2 - 3) / 2) + 1
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long total = (long long)B * OH * OW * C;
  if (idx >= total) return;
  const int c = (int)(idx % C);
  const long long pix = idx / C;
  const int ox = (int)(pix % OW), oy = (int)((pix / OW) % OH), b = (int)(pix / ((long long)OW * OH));
  float m = -INFINITY;
  for (int dy = -1; dy <= 1; ++dy)
    for (int dx = -1; dx <= 1; ++dx) {
      const int y = 2 * oy + dy, x = 2 * ox + dx;
      if (y >= 0 && y < H && x >= 0 && x < W) m = fmaxf(m, in[(((long long)b * H + y) * W + x) * C + c]);
    }
  if (out) out[idx] = m;
  if (hi) {
    const __half h = __float2half_rn(m);
    hi[pix * ldp + c] = h;
    lo[pix * ldp + c] = __float2half_rn((m - __half2float(h)) * kSplitScale);
  }
}

__global__ void copy_channels_kernel(const float* __restrict__ src, long long rows, int C, int ld_src, float* __restrict__ dst, int ld_dst,
                                     int c_off) {
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= rows * C) return;
  const long long r = idx / C;
  const int c = (int)(idx - r * C);
  dst[r * ld_dst + c_off + c] = src[r * ld_src + c];
}

__global__ void planes_to_f32_kernel(const __half* __restrict__ hi, const __half* __restrict__ lo, long long rows, int C, int ldp,
                                     float* __restrict__ out, int ld_out) {
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= rows * C) return;
  const long long r = idx / C;
  const int c = (int)(idx - r * C);
  out[r * ld_out + c] = fmaf(__half2float(lo[r * ldp + c]), 1.f / kSplitScale, __half2float(hi[r * ldp + c]));
}
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ in, int B, int C, int H, int W, float* __restrict__ out, int ld) {
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= (long long)B * H * W * ld) return;
  const int c = (int)(idx % ld);
  const long long pix = idx / ld;
  const long long hw = (long long)H * W;
  const long long b = pix / hw, p = pix - b * hw;
  out[idx] = c < C ? in[(b * C + c) * hw + p] : 0.f;
}

__global__ void zero_pad_channels_kernel(float* __restrict__ x, long long rows, int C, int ld) {
  const int padw = ld - C;
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= rows * padw) return;
  const long long r = idx / padw;
  x[r * ld + C + (int)(idx - r * padw)] = 0.f;
}
__global__ void fill_kernel(float* dst, size_t n, float v) {
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i < n) dst[i] = v;
}

__global__ void grid_flow_kernel(float* flow, int B, int h, int w) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * h * w) return;
  const int x = idx % w, y = (idx / w) % h;
  flow[idx * 2 + 0] = grid_coord(x, w);
  flow[idx * 2 + 1] = grid_coord(y, h);
}

// ------------------------------------------------------------------------------------------- GP
__global__ void row_norm_kernel(const float* __restrict__ x, long long rows, int C, int ld, float* __restrict__ nrm) {
  const long long r = blockIdx.x * (long long)(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= rows) return;
  const int lane = threadIdx.x & 31;
  float s = 0.f;
  for (int c = lane; c < C; c += 32) { const float v = x[r * ld + c]; s = fmaf(v, v, s); }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) nrm[r] = sqrtf(s);
}

// 32 x 32 output tile per CTA of 256 threads (each 4 outputs), K chunks of 32 through shared memory
__global__ void __launch_bounds__(256) cos_gram_kernel(const float* __restrict__ x, const float* __restrict__ y, int N, int M, int C, int ld,
                                                       const float* __restrict__ nx, const float* __restrict__ ny, float T, float add_diag,
                                                       float* __restrict__ K, const float* __restrict__ dots) {
  __shared__ float xs[32][33], ys[32][33];
  const int b = blockIdx.z, i0 = blockIdx.y * 32, j0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // ty 0..7 -> rows ty, ty+8, ty+16, ty+24
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  if (dots == nullptr) {
    const float* xb = x + (long long)b * N * ld;
    const float* yb = y + (long long)b * M * ld;
    for (int k0 = 0; k0 < C; k0 += 32) {
      for (int t = threadIdx.x; t < 1024; t += 256) {
        const int r = t >> 5, c = t & 31;
        xs[r][c] = (i0 + r < N && k0 + c < C) ? xb[(long long)(i0 + r) * ld + k0 + c] : 0.f;
        ys[r][c] = (j0 + r < M && k0 + c < C) ? yb[(long long)(j0 + r) * ld + k0 + c] : 0.f;
      }
      __syncthreads();
#pragma unroll 8
      for (int k = 0; k < 32; ++k) {
        const float yv = ys[tx][k];
#pragma unroll
        for (int r = 0; r < 4; ++r) acc[r] = fmaf(xs[ty + 8 * r][k], yv, acc[r]);
      }
      __syncthreads();
    }
  }
  const int j = j0 + tx;
  if (j >= M) return;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int i = i0 + ty + 8 * r;
    if (i >= N) continue;
    const long long o = ((long long)b * N + i) * M + j;
    const float d = dots ? dots[o] : acc[r];
    const float c = d / (nx[(long long)b * N + i] * ny[(long long)b * M + j] + 1e-6f);
    float v = expf((c - 1.f) / T);
    if (i == j) v += add_diag;
    K[o] = v;
  }
}

__global__ void pos_basis_kernel(const float* __restrict__ w, const float* __restrict__ bias, int B, int h, int w_, int D, float* __restrict__ f) {
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long total = (long long)B * h * w_ * D;
  if (idx >= total) return;
  const int d = (int)(idx % D);
  const long long pix = idx / D;
  const int x = (int)(pix % w_), y = (int)((pix / w_) % h);
  const float v = fmaf(w[d * 2], grid_coord(x, w_), fmaf(w[d * 2 + 1], grid_coord(y, h), bias[d]));
  f[idx] = cosf(8.f * kPi * v);
}

// ---- blocked Cholesky (lower), block size 32
constexpr int NB = 32;
// diagonal block: one warp, lane = row held in registers, column values exchanged by shuffles (the shared-memory version
// spent 29 us per block in a serial chain of dependent smem reads; this one ~2 us)
__global__ void __launch_bounds__(32) chol_diag_kernel(float* __restrict__ A, int N, int k0, int kb) {
  float* Ab = A + (long long)blockIdx.x * N * N;
  const int lane = threadIdx.x;
  float a[NB];
#pragma unroll
  for (int c = 0; c < NB; ++c)  // rows / columns beyond kb: identity (keeps the arithmetic finite, never written back)
    a[c] = (lane < kb && c < kb) ? Ab[(long long)(k0 + lane) * N + k0 + c] : (lane == c ? 1.f : 0.f);
#pragma unroll
  for (int j = 0; j < NB; ++j) {
    const float d = sqrtf(__shfl_sync(0xffffffffu, a[j], j));
    float lij = a[j] / d;          // lanes i > j: l[i][j]; lanes < j hold the (unused) upper triangle
    if (lane == j) lij = d;
    a[j] = lij;
#pragma unroll
    for (int k = j + 1; k < NB; ++k) a[k] -= lij * __shfl_sync(0xffffffffu, lij, k);  // a[i][k] -= l[i][j] l[k][j]
  }
#pragma unroll
  for (int c = 0; c < NB; ++c)
    if (lane < kb && c < kb) Ab[(long long)(k0 + lane) * N + k0 + c] = c <= lane ? a[c] : 0.f;
}
// rows below the diagonal block: L21 = A21 * L11^-T   (thread per row, forward substitution over the 32 columns)
__global__ void __launch_bounds__(128) chol_panel_kernel(float* __restrict__ A, int N, int k0, int kb) {
  __shared__ float l[NB][NB + 1];
  float* Ab = A + (long long)blockIdx.y * N * N;
  for (int t = threadIdx.x; t < NB * NB; t += 128) {
    const int r = t / NB, c = t % NB;
    l[r][c] = (r < kb && c < kb) ? Ab[(long long)(k0 + r) * N + k0 + c] : 0.f;
  }
  __syncthreads();
  const int i = k0 + kb + blockIdx.x * 128 + threadIdx.x;
  if (i >= N) return;
  float v[NB];
  float* row = Ab + (long long)i * N + k0;
#pragma unroll
  for (int c = 0; c < NB; ++c) v[c] = c < kb ? row[c] : 0.f;
#pragma unroll
  for (int c = 0; c < NB; ++c) {
    if (c < kb) {
      float s = v[c];
#pragma unroll
      for (int k = 0; k < NB; ++k)
        if (k < c) s -= v[k] * l[c][k];
      v[c] = s / l[c][c];
    }
  }
#pragma unroll
  for (int c = 0; c < NB; ++c)
    if (c < kb) row[c] = v[c];
}

// C[i, j] -= sum_k P[i, k] * Q[j, k]  (mode 0: Cholesky trailing update, lower tiles only, P = Q = panel)
// C[i, d] -= sum_k P[i, k] * Q[k, d]  (mode 1: forward substitution update)
// C[j, d] -= sum_k P[k, j] * Q[k, d]  (mode 2: backward substitution update)
// K <= 32.  64 x 64 tile per CTA of 256 threads (4 x 4 outputs each).
__global__ void __launch_bounds__(256) rank_update_kernel(int mode, float* __restrict__ C, long long c_batch, int ldc, int rows, int cols,
                                                          const float* __restrict__ P, long long p_batch, int ldp,
                                                          const float* __restrict__ Q, long long q_batch, int ldq, int kb) {
  __shared__ float ps[NB][65], qs[NB][65];
  const int b = blockIdx.z;
  const int i0 = blockIdx.y * 64, j0 = blockIdx.x * 64;
  if (mode == 0 && j0 > i0 + 63) return;  // strictly upper tile
  C += b * c_batch; P += b * p_batch; Q += b * q_batch;
  for (int t = threadIdx.x; t < NB * 64; t += 256) {
    const int k = t / 64, r = t % 64;
    float pv = 0.f, qv = 0.f;
    if (k < kb) {
      if (mode == 2) { if (i0 + r < rows) pv = P[(long long)k * ldp + i0 + r]; }
      else if (i0 + r < rows) pv = P[(long long)(i0 + r) * ldp + k];
      if (mode == 0) { if (j0 + r < cols) qv = Q[(long long)(j0 + r) * ldq + k]; }
      else if (j0 + r < cols) qv = Q[(long long)k * ldq + j0 + r];
    }
    ps[k][r] = pv; qs[k][r] = qv;
  }
  __syncthreads();
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float acc[4][4] = {};
#pragma unroll 8
  for (int k = 0; k < NB; ++k) {
    float pr[4], qr[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) { pr[r] = ps[k][ty + 16 * r]; qr[r] = qs[k][tx + 16 * r]; }
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[r][c] = fmaf(pr[r], qr[c], acc[r][c]);
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int i = i0 + ty + 16 * r;
    if (i >= rows) continue;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int j = j0 + tx + 16 * c;
      if (j < cols && !(mode == 0 && j > i)) C[(long long)i * ldc + j] -= acc[r][c];
    }
  }
}
// diagonal-block triangular solves for D right-hand sides (thread per column)
__global__ void __launch_bounds__(128) trsm_diag_kernel(const float* __restrict__ A, int N, int k0, int kb, float* __restrict__ F, int D,
                                                        int transposed) {
  __shared__ float l[NB][NB + 1];
  const float* Ab = A + (long long)blockIdx.y * N * N;
  float* Fb = F + (long long)blockIdx.y * N * D;
  for (int t = threadIdx.x; t < NB * NB; t += 128) {
    const int r = t / NB, c = t % NB;
    l[r][c] = (r < kb && c < kb) ? Ab[(long long)(k0 + r) * N + k0 + c] : 0.f;
  }
  __syncthreads();
  const int d = blockIdx.x * 128 + threadIdx.x;
  if (d >= D) return;
  float v[NB];
#pragma unroll
  for (int r = 0; r < NB; ++r) v[r] = r < kb ? Fb[(long long)(k0 + r) * D + d] : 0.f;
  if (!transposed) {  // L y = f
#pragma unroll
    for (int r = 0; r < NB; ++r)
      if (r < kb) {
        float s = v[r];
#pragma unroll
        for (int k = 0; k < NB; ++k)
          if (k < r) s -= l[r][k] * v[k];
        v[r] = s / l[r][r];
      }
  } else {            // L^T z = y
#pragma unroll
    for (int rr = 0; rr < NB; ++rr) {
      const int r = NB - 1 - rr;
      if (r < kb) {
        float s = v[r];
#pragma unroll
        for (int k = 0; k < NB; ++k)
          if (k > r && k < kb) s -= l[k][r] * v[k];
        v[r] = s / l[r][r];
      }
    }
  }
#pragma unroll
  for (int r = 0; r < NB; ++r)
    if (r < kb) Fb[(long long)(k0 + r) * D + d] = v[r];
}

// plain fp32 GEMM C = A (N x K) * B (K x D): 64 x 64 tiles, K chunks of 16
__global__ void __launch_bounds__(256) matmul_nn_kernel(const float* __restrict__ A, const float* __restrict__ Bm, int N, int K, int D,
                                                        float* __restrict__ C, int ldc) {
  __shared__ float as[16][65], bs[16][65];
  const int b = blockIdx.z, i0 = blockIdx.y * 64, j0 = blockIdx.x * 64;
  A += (long long)b * N * K; Bm += (long long)b * K * D; C += (long long)b * N * ldc;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < K; k0 += 16) {
    for (int t = threadIdx.x; t < 16 * 64; t += 256) {
      const int k = t & 15, r = t >> 4;
      as[k][r] = (i0 + r < N && k0 + k < K) ? A[(long long)(i0 + r) * K + k0 + k] : 0.f;
      const int kk = t >> 6, c = t & 63;
      bs[kk][c] = (j0 + c < D && k0 + kk < K) ? Bm[(long long)(k0 + kk) * D + j0 + c] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      float ar[4], br[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) { ar[r] = as[k][ty + 16 * r]; br[r] = bs[k][tx + 16 * r]; }
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[r][c] = fmaf(ar[r], br[c], acc[r][c]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int i = i0 + ty + 16 * r;
    if (i >= N) continue;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int j = j0 + tx + 16 * c;
      if (j < D) C[(long long)i * ldc + j] = acc[r][c];
    }
  }
}

// ------------------------------------------------------------------------------------------- CAB
__global__ void cab_pool_kernel(const float* __restrict__ x1, const float* __restrict__ x2, int HW, int C, float* __restrict__ pooled) {
  // pooled[b, c] for c in [0, 2C): mean over pixels; one CTA per (b, 32-channel group), 8 warps split the pixels
  __shared__ float part[8][32];
  const int b = blockIdx.y, c = blockIdx.x * 32 + (threadIdx.x & 31), wv = threadIdx.x >> 5;
  const float* src = c < C ? x1 : x2;
  const int cc = c < C ? c : c - C;
  float s = 0.f;
  for (int p = wv; p < HW; p += 8) s += src[((long long)b * HW + p) * C + cc];
  part[wv][threadIdx.x & 31] = s;
  __syncthreads();
  if (wv == 0) {
    float t = 0.f;
    for (int k = 0; k < 8; ++k) t += part[k][threadIdx.x];
    pooled[(long long)b * 2 * C + c] = t / (float)HW;
  }
}
__global__ void cab_fc_kernel(const float* __restrict__ in, int Cin, const float* __restrict__ w, const float* __restrict__ bias, int Cout,
                              int act /*1 relu, 2 sigmoid*/, float* __restrict__ out) {
  // out[b, o] = act(sum_i w[o, i] in[b, i] + bias[o]); one warp per output
  const int b = blockIdx.y, o = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (o >= Cout) return;
  float s = 0.f;
  for (int i = lane; i < Cin; i += 32) s = fmaf(w[(long long)o * Cin + i], in[(long long)b * Cin + i], s);
#pragma unroll
  for (int k = 16; k > 0; k >>= 1) s += __shfl_xor_sync(0xffffffffu, s, k);
  if (lane == 0) {
    s += bias[o];
    out[(long long)b * Cout + o] = act == 1 ? fmaxf(s, 0.f) : 1.f / (1.f + expf(-s));
  }
}
__global__ void cab_apply_kernel(const float* __restrict__ x1, const float* __restrict__ x2, const float* __restrict__ s, int HW, int C,
                                 long long total, float* __restrict__ out) {
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int c = (int)(idx % C);
  const long long b = idx / ((long long)HW * C);
  out[idx] = fmaf(s[b * C + c], x2[idx], x1[idx]);
}

// ------------------------------------------------------------------------------------------- ConvRefiner pieces
struct Bil {
  int x0, y0;
  float w00, w01, w10, w11;  // weights of (y0,x0) (y0,x0+1) (y0+1,x0) (y0+1,x0+1)
};
__device__ __forceinline__ Bil bil_setup(float gx, float gy, int w, int h) {
  // grid_sample, align_corners=False: pixel = ((g + 1) * size - 1) / 2
  const float ix = ((gx + 1.f) * (float)w - 1.f) * 0.5f, iy = ((gy + 1.f) * (float)h - 1.f) * 0.5f;
  const float fx = floorf(ix), fy = floorf(iy);
  Bil b;
  b.x0 = (int)fx; b.y0 = (int)fy;
  const float ax = ix - fx, ay = iy - fy;
  b.w00 = (1.f - ax) * (1.f - ay); b.w01 = ax * (1.f - ay); b.w10 = (1.f - ax) * ay; b.w11 = ax * ay;
  return b;
}
__global__ void grid_sample_kernel(const float* __restrict__ y, int B, int h, int w, int C, int ld_y, const float* __restrict__ flow,
                                   float* __restrict__ out, int ld_out, int c_off) {
  // one warp per pixel, lanes over channels
  const long long pix = blockIdx.x * (long long)(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (pix >= (long long)B * h * w) return;
  const int lane = threadIdx.x & 31;
  const int b = (int)(pix / ((long long)h * w));
  const float gx = flow[pix * 2], gy = flow[pix * 2 + 1];
  // out-of-range or non-finite targets sample zeros (padding_mode='zeros')
  if (!(fabsf(gx) < 1e6f) || !(fabsf(gy) < 1e6f)) {
    for (int c = lane; c < C; c += 32) out[pix * ld_out + c_off + c] = 0.f;
    return;
  }
  const Bil bl = bil_setup(gx, gy, w, h);
  const float* base = y + (long long)b * h * w * ld_y;
  const bool vx0 = bl.x0 >= 0 && bl.x0 < w, vx1 = bl.x0 + 1 >= 0 && bl.x0 + 1 < w;
  const bool vy0 = bl.y0 >= 0 && bl.y0 < h, vy1 = bl.y0 + 1 >= 0 && bl.y0 + 1 < h;
  const float* p00 = base + ((long long)bl.y0 * w + bl.x0) * ld_y;
  const float* p01 = p00 + ld_y;
  const float* p10 = p00 + (long long)w * ld_y;
  const float* p11 = p10 + ld_y;
  for (int c = lane; c < C; c += 32) {
    float v = 0.f;
    if (vy0 && vx0) v = fmaf(bl.w00, p00[c], v);
    if (vy0 && vx1) v = fmaf(bl.w01, p01[c], v);
    if (vy1 && vx0) v = fmaf(bl.w10, p10[c], v);
    if (vy1 && vx1) v = fmaf(bl.w11, p11[c], v);
    out[pix * ld_out + c_off + c] = v;
  }
}
__global__ void disp_emb_kernel(const float* __restrict__ flow, int B, int h, int w, const float* __restrict__ wt,
                                const float* __restrict__ bias, int E, float* __restrict__ out, int ld_out, int c_off) {
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= (long long)B * h * w * E) return;
  const int e = (int)(idx % E);
  const long long pix = idx / E;
  const int x = (int)(pix % w), yy = (int)((pix / w) % h);
  const float dx = flow[pix * 2] - grid_coord(x, w), dy = flow[pix * 2 + 1] - grid_coord(yy, h);
  out[pix * ld_out + c_off + e] = fmaf(wt[e * 2], dx, fmaf(wt[e * 2 + 1], dy, bias[e]));
}
// Local correlation around the flow target (networks/dkm/utils/local_correlation.py:24-39).  The (2r+1)^2 window offsets
// are exactly one pixel apart, so every window sample shares the flow target's fractional position: the warp first
// takes the dot products D[u][v] = <x[p], y[y0 + u - r, x0 + v - r]> on the (2r+2)^2 integer grid (zero outside the
// image) and then blends four neighbours per output - a quarter of the reference's bilinear samples.
__global__ void __launch_bounds__(128) local_corr_kernel(const float* __restrict__ x, const float* __restrict__ y, int B, int h, int w, int C,
                                                         int ld, const float* __restrict__ flow, int r, float* __restrict__ out, int ld_out,
                                                         int c_off) {
  extern __shared__ float dsm[];  // [4 warps][(2r+2)^2]
  const int wv = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long pix = blockIdx.x * 4ll + wv;
  const int G = 2 * r + 2, K1 = 2 * r + 1;
  float* D = dsm + wv * G * G;
  if (pix >= (long long)B * h * w) return;
  const int b = (int)(pix / ((long long)h * w));
  const float gx = flow[pix * 2], gy = flow[pix * 2 + 1];
  const float scale = rsqrtf((float)C);
  float* o = out + pix * ld_out + c_off;
  if (!(fabsf(gx) < 1e6f) || !(fabsf(gy) < 1e6f)) {
    for (int k = lane; k < K1 * K1; k += 32) o[k] = 0.f;
    return;
  }
  const Bil bl = bil_setup(gx, gy, w, h);
  const float* xp = x + pix * ld;
  const float* yb = y + (long long)b * h * w * ld;
  // channels of x[p] held in registers: up to 512 channels -> 16 per lane
  float xr[16];
#pragma unroll
  for (int t = 0; t < 16; ++t) xr[t] = (lane + 32 * t < C) ? xp[lane + 32 * t] : 0.f;
  for (int g = 0; g < G * G; ++g) {
    const int u = g / G, v = g - u * G;
    const int yy = bl.y0 + u - r, xx = bl.x0 + v - r;
    float s = 0.f;
    if (yy >= 0 && yy < h && xx >= 0 && xx < w) {
      const float* yp = yb + ((long long)yy * w + xx) * ld;
#pragma unroll
      for (int t = 0; t < 16; ++t)
        if (lane + 32 * t < C) s = fmaf(xr[t], yp[lane + 32 * t], s);
#pragma unroll
      for (int k = 16; k > 0; k >>= 1) s += __shfl_xor_sync(0xffffffffu, s, k);
    }
    if (lane == 0) D[g] = s;
  }
  __syncwarp();
  for (int k = lane; k < K1 * K1; k += 32) {
    const int u = k / K1, v = k - u * K1;  // window row (y offset u - r), column (x offset v - r)
    const float val = bl.w00 * D[u * G + v] + bl.w01 * D[u * G + v + 1] + bl.w10 * D[(u + 1) * G + v] + bl.w11 * D[(u + 1) * G + v + 1];
    o[k] = val * scale;
  }
}

__global__ void depthwise5x5_kernel(const float* __restrict__ in, int B, int h, int w, int Cin, int ld_in, int mult,
                                    const float* __restrict__ wt, const float* __restrict__ scale, const float* __restrict__ bias,
                                    float* __restrict__ out, int ld_out, __half* hi, __half* lo, int ldp) {
  // thread per (pixel, output channel); consecutive threads = consecutive channels (coalesced over NHWC)
  const int Cout = Cin * mult;
  const int cpad = hi ? ldp : Cout;  // planes: pad channels are written as zeros
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= (long long)B * h * w * cpad) return;
  const int co = (int)(idx % cpad);
  const long long pix = idx / cpad;
  if (co >= Cout) {
    hi[pix * ldp + co] = __float2half_rn(0.f);
    lo[pix * ldp + co] = __float2half_rn(0.f);
    return;
  }
  const int x = (int)(pix % w), y = (int)((pix / w) % h), b = (int)(pix / ((long long)h * w));
  const int ci = co / mult;
  const float* wp = wt + (long long)co * 25;
  float s = 0.f;
#pragma unroll
  for (int dy = 0; dy < 5; ++dy) {
    const int yy = y + dy - 2;
    if (yy < 0 || yy >= h) continue;
#pragma unroll
    for (int dx = 0; dx < 5; ++dx) {
      const int xx = x + dx - 2;
      if (xx < 0 || xx >= w) continue;
      s = fmaf(wp[dy * 5 + dx], in[(((long long)b * h + yy) * w + xx) * ld_in + ci], s);
    }
  }
  const float v = fmaxf(fmaf(s, scale[co], bias[co]), 0.f);
  if (out) out[pix * ld_out + co] = v;
  if (hi) {
    const __half hh = __float2half_rn(v);
    hi[pix * ldp + co] = hh;
    lo[pix * ldp + co] = __float2half_rn((v - __half2float(hh)) * kSplitScale);
  }
}

// Depthwise 5x5 + folded BatchNorm + ReLU, channel multiplier 1, vectorised: a thread owns 4 channels (one float4 of the
// NHWC row) and a 4 (x) by 2 (y) patch of output pixels, walks the 6 input rows of the patch once (8 float4 loads per row)
// and reuses every loaded pixel for up to 10 outputs - 6 loads per output pixel instead of 25.  Weights / scale / bias come
// transposed and padded ([25][Cp], Cp = plane pitch) so that consecutive threads read consecutive float4s.
template <bool kMult2>  // kMult2: channel multiplier 2 (output channel co reads input channel co / 2; block1 of the 1/1 refiner)
__global__ void __launch_bounds__(128) depthwise5x5_v4_kernel(const float* __restrict__ in, int B, int h, int w, int C, int ld_in,
                                                              const float* __restrict__ wt /*[25][Cp]*/, const float* __restrict__ scale,
                                                              const float* __restrict__ bias, int Cp, float* __restrict__ out, int ld_out,
                                                              __half* __restrict__ hi, __half* __restrict__ lo, int ldp) {
  // work items = (pixel tile, channel quad) pairs, channel quads fastest: every thread is busy whatever C is (C = 24 has
  // only 8 quads) and consecutive threads read consecutive float4s
  const int cq_total = Cp >> 2;
  const long long item = blockIdx.x * 128ll + threadIdx.x;
  const int tiles_x = (w + 3) >> 2, tiles_y = (h + 1) >> 1;
  if (item >= (long long)B * tiles_y * tiles_x * cq_total) return;
  const int cq = (int)(item % cq_total);
  int t = (int)(item / cq_total);
  const int tx = t % tiles_x; t /= tiles_x;
  const int ty = t % tiles_y;
  const int b = t / tiles_y;
  const int x0 = tx * 4, y0 = ty * 2;
  const int c = cq * 4;
  float4 acc[2][4];
#pragma unroll
  for (int r = 0; r < 2; ++r)
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[r][i] = make_float4(0.f, 0.f, 0.f, 0.f);
  const float* base = in + (long long)b * h * w * ld_in + (kMult2 ? c / 2 : c);
#pragma unroll
  for (int ry = 0; ry < 6; ++ry) {  // input rows y0 - 2 .. y0 + 3
    const int yy = y0 + ry - 2;
    if (yy < 0 || yy >= h) continue;
    float4 row[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int xx = x0 + i - 2;
      row[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (xx >= 0 && xx < w && c < C) {
        if (kMult2) {
          const float2 v2 = __ldg(reinterpret_cast<const float2*>(base + ((long long)yy * w + xx) * ld_in));
          row[i] = make_float4(v2.x, v2.x, v2.y, v2.y);
        } else {
          row[i] = __ldg(reinterpret_cast<const float4*>(base + ((long long)yy * w + xx) * ld_in));
        }
      }
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int dy = ry - r;  // kernel row used by output row r
      if (dy < 0 || dy > 4) continue;
#pragma unroll
      for (int dx = 0; dx < 5; ++dx) {
        const float4 wv = __ldg(reinterpret_cast<const float4*>(wt + (long long)(dy * 5 + dx) * Cp + c));
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4 v = row[i + dx];
          acc[r][i].x = fmaf(wv.x, v.x, acc[r][i].x);
          acc[r][i].y = fmaf(wv.y, v.y, acc[r][i].y);
          acc[r][i].z = fmaf(wv.z, v.z, acc[r][i].z);
          acc[r][i].w = fmaf(wv.w, v.w, acc[r][i].w);
        }
      }
    }
  }
  const float4 sc = __ldg(reinterpret_cast<const float4*>(scale + c)), bi = __ldg(reinterpret_cast<const float4*>(bias + c));
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int y = y0 + r;
    if (y >= h) continue;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int x = x0 + i;
      if (x >= w) continue;
      float4 v;  // channels >= C have zero weights / scale / bias: exactly 0 (the GEMM's K padding)
      v.x = fmaxf(fmaf(acc[r][i].x, sc.x, bi.x), 0.f);
      v.y = fmaxf(fmaf(acc[r][i].y, sc.y, bi.y), 0.f);
      v.z = fmaxf(fmaf(acc[r][i].z, sc.z, bi.z), 0.f);
      v.w = fmaxf(fmaf(acc[r][i].w, sc.w, bi.w), 0.f);
      const long long pix = ((long long)b * h + y) * w + x;
      if (out && c < C) *reinterpret_cast<float4*>(out + pix * ld_out + c) = v;
      if (hi) {
        const __half2 h01 = __floats2half2_rn(v.x, v.y), h23 = __floats2half2_rn(v.z, v.w);
        const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
        const __half2 l01 = __floats2half2_rn((v.x - f01.x) * kSplitScale, (v.y - f01.y) * kSplitScale);
        const __half2 l23 = __floats2half2_rn((v.z - f23.x) * kSplitScale, (v.w - f23.y) * kSplitScale);
        uint2 hv, lv;
        hv.x = *reinterpret_cast<const unsigned*>(&h01); hv.y = *reinterpret_cast<const unsigned*>(&h23);
        lv.x = *reinterpret_cast<const unsigned*>(&l01); lv.y = *reinterpret_cast<const unsigned*>(&l23);
        *reinterpret_cast<uint2*>(hi + pix * ldp + c) = hv;
        *reinterpret_cast<uint2*>(lo + pix * ldp + c) = lv;
      }
    }
  }
}

__global__ void apply_delta_kernel(float* __restrict__ flow, float* __restrict__ cert, int accumulate, const float* __restrict__ head,
                                   int ld_head, long long npix, float fx, float fy) {
  const long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (p >= npix) return;
  const float dc = head[p * ld_head], dx = head[p * ld_head + 1], dy = head[p * ld_head + 2];
  flow[p * 2] += dx * fx;
  flow[p * 2 + 1] += dy * fy;
  cert[p] = accumulate ? cert[p] + dc : dc;
}
__global__ void split_head_kernel(const float* __restrict__ head, int ld_head, long long rows, float* __restrict__ flow, float* __restrict__ cert) {
  const long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (p >= rows) return;
  cert[p] = head[p * ld_head];
  flow[p * 2] = head[p * ld_head + 1];
  flow[p * 2 + 1] = head[p * ld_head + 2];
}

__global__ void finalize_kernel(const DkmFinalArgs a) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= 2 * a.hs * a.ws) return;
  const int x = idx % a.ws, y = (idx / a.ws) % a.hs, half = idx / (a.hs * a.ws);
  float fx = a.flow[idx * 2], fy = a.flow[idx * 2 + 1];
  float c = a.certainty[idx];
  if (a.low_cert) {
    const float l = a.low_cert[idx];
    c -= (l < 0.f) ? 0.5f * l : 0.f;  // dkm.py:688-693, 707
  }
  c = 1.f / (1.f + expf(-c));
  if (fabsf(fx) > 1.f || fabsf(fy) > 1.f) c = 0.f;           // dkm.py:721-723
  // black pixels of the ORIGINAL image, nearest-neighbour resized (dkm.py:726-731)
  const float* im = half ? a.im2 : a.im1;
  const int H = half ? a.H2 : a.H1, W = half ? a.W2 : a.W1;
  const int sy = min((int)floorf((float)y * ((float)H / (float)a.hs)), H - 1);
  const int sx = min((int)floorf((float)x * ((float)W / (float)a.ws)), W - 1);
  const long long plane = (long long)H * W, off = (long long)sy * W + sx;
  if (im[off] < 0.03125f && im[plane + off] < 0.03125f && im[2 * plane + off] < 0.03125f) c = 0.f;
  fx = fminf(fmaxf(fx, -1.f), 1.f);
  fy = fminf(fmaxf(fy, -1.f), 1.f);
  const float qx = grid_coord(x, a.ws), qy = grid_coord(y, a.hs);
  float* wp = a.warp + ((long long)y * (2 * a.ws) + half * a.ws + x) * 4;
  if (half == 0) { wp[0] = qx; wp[1] = qy; wp[2] = fx; wp[3] = fy; }      // q_warp = (query_coords, q->s)
  else { wp[0] = fx; wp[1] = fy; wp[2] = qx; wp[3] = qy; }                 // s_warp = (s->q, support_coords)
  a.cert_out[(long long)y * (2 * a.ws) + half * a.ws + x] = c;
}

// Gaussian kernel density of n 4-D points (networks/dkm/utils/kde.py:17-26: exp(-cdist(x, x)^2 / (2 std^2)).sum(-1)) without
// the n x n distance matrix: 256 points per CTA against tiles of 256 points staged in shared memory.
__global__ void __launch_bounds__(256) kde_kernel(const float4* __restrict__ x, int n, float inv2s2, float* __restrict__ density) {
  __shared__ float4 tile[256];
  const int i = blockIdx.x * 256 + threadIdx.x;
  const float4 p = i < n ? x[i] : make_float4(0.f, 0.f, 0.f, 0.f);
  float acc = 0.f;
  for (int j0 = 0; j0 < n; j0 += 256) {
    tile[threadIdx.x] = (j0 + threadIdx.x < n) ? x[j0 + threadIdx.x] : make_float4(1e18f, 1e18f, 1e18f, 1e18f);
    __syncthreads();
#pragma unroll 8
    for (int j = 0; j < 256; ++j) {
      const float4 q = tile[j];
      const float dx = p.x - q.x, dy = p.y - q.y, dz = p.z - q.z, dw = p.w - q.w;
      acc += __expf(-(dx * dx + dy * dy + dz * dz + dw * dw) * inv2s2);
    }
    __syncthreads();
  }
  if (i < n) density[i] = acc;
}

inline unsigned blocks(long long n, int t) { return (unsigned)((n + t - 1) / t); }

}  // namespace

#define GIMB_DKM_LAUNCH_END() \
  do { ctx.launches++; GIMB_LAUNCH_CHECK(); return 0; } while (0)

int dkm_resize_nchw(Ctx& ctx, const float* in, int B, int C, int H, int W, float* out, int OH, int OW) {
  if (ctx.dry) return 0;
  const long long n = (long long)B * C * OH * OW;
  resize_nchw_kernel<<<blocks(n, 256), 256, 0, ctx.stream>>>(in, B * C, H, W, out, OH, OW);
  GIMB_DKM_LAUNCH_END();
}
int dkm_resize_nhwc(Ctx& ctx, const float* in, int B, int H, int W, int C, int ld_in, float* out, int OH, int OW, int ld_out) {
  if (ctx.dry) return 0;
  const long long n = (long long)B * OH * OW * C;
  resize_nhwc_kernel<<<blocks(n, 256), 256, 0, ctx.stream>>>(in, B, H, W, C, ld_in, out, OH, OW, ld_out);
  GIMB_DKM_LAUNCH_END();
}
int dkm_maxpool3x3s2(Ctx& ctx, const float* in, int B, int H, int W, int C, float* out, const SplitPlanes* planes) {
  if (ctx.dry) return 0;
  const long long n = (long long)B * ((H + 1) / 2) * ((W + 1) / 2) * C;
  maxpool_kernel<<<blocks(n, 256), 256, 0, ctx.stream>>>(in, B, H, W, C, out, planes ? planes->hi : nullptr, planes ? planes->lo : nullptr,
                                                         planes ? planes->ld : 0);
  GIMB_DKM_LAUNCH_END();
}
int dkm_copy_channels(Ctx& ctx, const float* src, int64_t rows, int C, int ld_src, float* dst, int ld_dst, int c_off) {
  if (ctx.dry || rows == 0) return 0;
  copy_channels_kernel<<<blocks(rows * C, 256), 256, 0, ctx.stream>>>(src, rows, C, ld_src, dst, ld_dst, c_off);
  GIMB_DKM_LAUNCH_END();
}
int planes_to_f32(Ctx& ctx, const SplitPlanes& sp, int64_t rows, int C, float* out, int ld_out) {
  if (ctx.dry || rows == 0) return 0;
  planes_to_f32_kernel<<<blocks(rows * C, 256), 256, 0, ctx.stream>>>(sp.hi, sp.lo, rows, C, sp.ld, out, ld_out);
  GIMB_DKM_LAUNCH_END();
}
int nchw_to_nhwc(Ctx& ctx, const float* in, int B, int C, int H, int W, float* out, int ld) {
  if (ctx.dry) return 0;
  nchw_to_nhwc_kernel<<<blocks((long long)B * H * W * ld, 256), 256, 0, ctx.stream>>>(in, B, C, H, W, out, ld);
  GIMB_DKM_LAUNCH_END();
}
int dkm_fill(Ctx& ctx, float* dst, size_t n, float v) {
  if (ctx.dry || n == 0) return 0;
  fill_kernel<<<blocks((long long)n, 256), 256, 0, ctx.stream>>>(dst, n, v);
  GIMB_DKM_LAUNCH_END();
}
int dkm_zero_pad_channels(Ctx& ctx, float* x, int64_t rows, int C, int ld) {
  if (ctx.dry || rows == 0 || ld == C) return 0;
  zero_pad_channels_kernel<<<blocks(rows * (ld - C), 256), 256, 0, ctx.stream>>>(x, rows, C, ld);
  GIMB_DKM_LAUNCH_END();
}
int dkm_grid_flow(Ctx& ctx, float* flow, int B, int h, int w) {
  if (ctx.dry) return 0;
  grid_flow_kernel<<<blocks((long long)B * h * w, 256), 256, 0, ctx.stream>>>(flow, B, h, w);
  GIMB_DKM_LAUNCH_END();
}

static int cos_gram_impl(Ctx& ctx, const float* x, const float* y, int B, int N, int M, int C, int ld, float T, float add_diag, float* K,
                         const float* dots) {
  size_t mark = ctx.arena.mark();
  float* nx = ctx.arena.alloc<float>((size_t)B * N);
  float* ny = ctx.arena.alloc<float>((size_t)B * M);
  if (!ctx.dry) {
    GIMB_CHECK(!ctx.arena.overflow, "dkm_cos_gram: workspace exhausted");
    row_norm_kernel<<<blocks((long long)B * N, 8), 256, 0, ctx.stream>>>(x, (long long)B * N, C, ld, nx);
    row_norm_kernel<<<blocks((long long)B * M, 8), 256, 0, ctx.stream>>>(y, (long long)B * M, C, ld, ny);
    dim3 grid(cdiv(M, 32), cdiv(N, 32), B);
    cos_gram_kernel<<<grid, 256, 0, ctx.stream>>>(x, y, N, M, C, ld, nx, ny, T, add_diag, K, dots);
    ctx.launches += 3;
    GIMB_LAUNCH_CHECK();
  }
  ctx.arena.release(mark);
  return 0;
}
int dkm_cos_gram(Ctx& ctx, const float* x, const float* y, int B, int N, int M, int C, int ld, float T, float add_diag, float* K) {
  return cos_gram_impl(ctx, x, y, B, N, M, C, ld, T, add_diag, K, nullptr);
}
int dkm_cos_gram_finish(Ctx& ctx, float* D, const float* x, const float* y, int B, int N, int M, int C, int ld, float T, float add_diag) {
  return cos_gram_impl(ctx, x, y, B, N, M, C, ld, T, add_diag, D, D);
}
int dkm_pos_basis(Ctx& ctx, const float* w, const float* bias, int B, int h, int w_, int D, float* f) {
  if (ctx.dry) return 0;
  pos_basis_kernel<<<blocks((long long)B * h * w_ * D, 256), 256, 0, ctx.stream>>>(w, bias, B, h, w_, D, f);
  GIMB_DKM_LAUNCH_END();
}

int dkm_chol_solve(Ctx& ctx, float* A, float* F, int B, int N, int D) {
  if (ctx.dry) return 0;
  const long long ab = (long long)N * N, fb = (long long)N * D;
  // ---- factor A = L L^T (lower), right-looking, block 32
  for (int k0 = 0; k0 < N; k0 += NB) {
    const int kb = std::min(NB, N - k0), rem = N - k0 - kb;
    chol_diag_kernel<<<B, 32, 0, ctx.stream>>>(A, N, k0, kb);
    if (rem > 0) {
      chol_panel_kernel<<<dim3(cdiv(rem, 128), B), 128, 0, ctx.stream>>>(A, N, k0, kb);
      const float* P = A + (long long)(k0 + kb) * N + k0;
      float* Cc = A + (long long)(k0 + kb) * N + k0 + kb;
      rank_update_kernel<<<dim3(cdiv(rem, 64), cdiv(rem, 64), B), 256, 0, ctx.stream>>>(0, Cc, ab, N, rem, rem, P, ab, N, P, ab, N, kb);
      ctx.launches += 2;
    }
    ctx.launches++;
  }
  GIMB_LAUNCH_CHECK();
  // ---- forward substitution L Y = F
  for (int k0 = 0; k0 < N; k0 += NB) {
    const int kb = std::min(NB, N - k0), rem = N - k0 - kb;
    trsm_diag_kernel<<<dim3(cdiv(D, 128), B), 128, 0, ctx.stream>>>(A, N, k0, kb, F, D, 0);
    if (rem > 0) {
      rank_update_kernel<<<dim3(cdiv(D, 64), cdiv(rem, 64), B), 256, 0, ctx.stream>>>(1, F + (long long)(k0 + kb) * D, fb, D, rem, D,
                                                                                      A + (long long)(k0 + kb) * N + k0, ab, N,
                                                                                      F + (long long)k0 * D, fb, D, kb);
      ctx.launches++;
    }
    ctx.launches++;
  }
  GIMB_LAUNCH_CHECK();
  // ---- backward substitution L^T Z = Y
  for (int k0 = (N - 1) / NB * NB; k0 >= 0; k0 -= NB) {
    const int kb = std::min(NB, N - k0);
    trsm_diag_kernel<<<dim3(cdiv(D, 128), B), 128, 0, ctx.stream>>>(A, N, k0, kb, F, D, 1);
    if (k0 > 0) {
      // Y[0:k0, :] -= L[k0:k0+kb, 0:k0]^T Z[k0:k0+kb, :]
      rank_update_kernel<<<dim3(cdiv(D, 64), cdiv(k0, 64), B), 256, 0, ctx.stream>>>(2, F, fb, D, k0, D, A + (long long)k0 * N, ab, N,
                                                                                     F + (long long)k0 * D, fb, D, kb);
      ctx.launches++;
    }
    ctx.launches++;
  }
  GIMB_LAUNCH_CHECK();
  return 0;
}
int dkm_matmul_nn(Ctx& ctx, const float* A, const float* Bm, int B, int N, int K, int D, float* C, int ld_out) {
  if (ctx.dry) return 0;
  matmul_nn_kernel<<<dim3(cdiv(D, 64), cdiv(N, 64), B), 256, 0, ctx.stream>>>(A, Bm, N, K, D, C, ld_out);
  GIMB_DKM_LAUNCH_END();
}

int dkm_cab(Ctx& ctx, const float* x1, const float* x2, int B, int HW, int C, const float* w1, const float* b1, const float* w2,
            const float* b2, float* out, float* scratch) {
  GIMB_CHECK(C % 32 == 0, "dkm_cab: C must be a multiple of 32");
  if (ctx.dry) return 0;
  float* pooled = scratch;                    // [B, 2C]
  float* hid = scratch + (size_t)B * 2 * C;   // [B, C]
  float* sig = hid + (size_t)B * C;           // [B, C]
  cab_pool_kernel<<<dim3(2 * C / 32, B), 256, 0, ctx.stream>>>(x1, x2, HW, C, pooled);
  cab_fc_kernel<<<dim3(cdiv(C, 8), B), 256, 0, ctx.stream>>>(pooled, 2 * C, w1, b1, C, 1, hid);
  cab_fc_kernel<<<dim3(cdiv(C, 8), B), 256, 0, ctx.stream>>>(hid, C, w2, b2, C, 2, sig);
  const long long total = (long long)B * HW * C;
  cab_apply_kernel<<<blocks(total, 256), 256, 0, ctx.stream>>>(x1, x2, sig, HW, C, total, out);
  ctx.launches += 4;
  GIMB_LAUNCH_CHECK();
  return 0;
}

int dkm_grid_sample(Ctx& ctx, const float* y, int B, int h, int w, int C, int ld_y, const float* flow, float* out, int ld_out, int c_off) {
  if (ctx.dry) return 0;
  grid_sample_kernel<<<blocks((long long)B * h * w, 8), 256, 0, ctx.stream>>>(y, B, h, w, C, ld_y, flow, out, ld_out, c_off);
  GIMB_DKM_LAUNCH_END();
}
int dkm_disp_emb(Ctx& ctx, const float* flow, int B, int h, int w, const float* wt, const float* bias, int E, float* out, int ld_out,
                 int c_off) {
  if (ctx.dry) return 0;
  disp_emb_kernel<<<blocks((long long)B * h * w * E, 256), 256, 0, ctx.stream>>>(flow, B, h, w, wt, bias, E, out, ld_out, c_off);
  GIMB_DKM_LAUNCH_END();
}
int dkm_local_corr(Ctx& ctx, const float* x, const float* y, int B, int h, int w, int C, int ld, const float* flow, int r, float* out,
                   int ld_out, int c_off) {
  GIMB_CHECK(C <= 512 && r >= 1 && r <= 7, "dkm_local_corr: C <= 512 and radius 1..7 expected");
  if (ctx.dry) return 0;
  const int G = 2 * r + 2;
  local_corr_kernel<<<blocks((long long)B * h * w, 4), 128, 4 * G * G * sizeof(float), ctx.stream>>>(x, y, B, h, w, C, ld, flow, r, out, ld_out,
                                                                                                      c_off);
  GIMB_DKM_LAUNCH_END();
}
int dkm_depthwise5x5(Ctx& ctx, const float* in, int B, int h, int w, int Cin, int ld_in, int mult, const float* wt, const float* scale,
                     const float* bias, float* out, int ld_out, const SplitPlanes* planes) {
  if (ctx.dry) return 0;
  const int cpad = planes ? planes->ld : Cin * mult;
  depthwise5x5_kernel<<<blocks((long long)B * h * w * cpad, 256), 256, 0, ctx.stream>>>(in, B, h, w, Cin, ld_in, mult, wt, scale, bias, out,
                                                                                        ld_out, planes ? planes->hi : nullptr,
                                                                                        planes ? planes->lo : nullptr, planes ? planes->ld : 0);
  GIMB_DKM_LAUNCH_END();
}
int dkm_depthwise5x5_v4(Ctx& ctx, const float* in, int B, int h, int w, int C, int ld_in, const float* wt_t, const float* scale_p,
                        const float* bias_p, int Cp, float* out, int ld_out, const SplitPlanes* planes, int mult) {
  GIMB_CHECK(Cp % 4 == 0 && ld_in % 4 == 0 && (mult == 2 || ld_in >= Cp - 3) && (mult == 1 || mult == 2) &&
                 (!planes || planes->ld == Cp) && (!out || ld_out % 4 == 0),
             "dkm_depthwise5x5_v4: pitches must be multiples of 4 and cover the padded channel count");
  if (ctx.dry) return 0;
  const long long items = (long long)B * ((h + 1) / 2) * ((w + 3) / 4) * (Cp / 4);
  if (mult == 2)
    depthwise5x5_v4_kernel<true><<<blocks(items, 128), 128, 0, ctx.stream>>>(in, B, h, w, C, ld_in, wt_t, scale_p, bias_p, Cp, out, ld_out,
                                                                             planes ? planes->hi : nullptr, planes ? planes->lo : nullptr,
                                                                             planes ? planes->ld : 0);
  else
    depthwise5x5_v4_kernel<false><<<blocks(items, 128), 128, 0, ctx.stream>>>(in, B, h, w, C, ld_in, wt_t, scale_p, bias_p, Cp, out, ld_out,
                                                                              planes ? planes->hi : nullptr, planes ? planes->lo : nullptr,
                                                                              planes ? planes->ld : 0);
  GIMB_DKM_LAUNCH_END();
}
int dkm_apply_delta(Ctx& ctx, float* flow, float* certainty, bool cert_accumulate, const float* head, int ld_head, int B, int hs, int ws,
                    float ins, int W, int H) {
  if (ctx.dry) return 0;
  const long long n = (long long)B * hs * ws;
  apply_delta_kernel<<<blocks(n, 256), 256, 0, ctx.stream>>>(flow, certainty, cert_accumulate ? 1 : 0, head, ld_head, n, ins / (4.f * (float)W),
                                                             ins / (4.f * (float)H));
  GIMB_DKM_LAUNCH_END();
}
int dkm_split_head(Ctx& ctx, const float* head, int ld_head, int64_t rows, float* flow, float* certainty) {
  if (ctx.dry) return 0;
  split_head_kernel<<<blocks(rows, 256), 256, 0, ctx.stream>>>(head, ld_head, rows, flow, certainty);
  GIMB_DKM_LAUNCH_END();
}
int dkm_kde(Ctx& ctx, const float* x, int n, float std, float* density) {
  if (ctx.dry || n == 0) return 0;
  kde_kernel<<<blocks(n, 256), 256, 0, ctx.stream>>>(reinterpret_cast<const float4*>(x), n, 1.f / (2.f * std * std), density);
  GIMB_DKM_LAUNCH_END();
}
int dkm_finalize(Ctx& ctx, const DkmFinalArgs& a) {
  if (ctx.dry) return 0;
  finalize_kernel<<<blocks(2ll * a.hs * a.ws, 256), 256, 0, ctx.stream>>>(a);
  GIMB_DKM_LAUNCH_END();
}

}  // namespace gimb
