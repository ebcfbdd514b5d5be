// simt_tile.cuh - the fp32 CUDA-core 128 x BN x 16 implicit-GEMM tile mainloop shared by the convolution /
// Linear kernel (conv_simt.cu) and the coarse-matching sweeps (coarse_match.cu).
//
//   256 threads, 8 x (BN/16) accumulators per thread: thread (tx, ty) owns rows ty + 16*i, cols tx + 16*j;
//   cp.async (LDGSTS, zero-fill for padding / tails) into a 4-stage shared-memory ring;
//   shared tiles are [row][16 + 4 pad] floats: float4 reads along k are bank-conflict free for the
//   strided thread->row mapping.
#pragma once
#include "common.cuh"

namespace gimb {
namespace simt {

constexpr int BM = 128;
constexpr int BK = 16;
constexpr int LDS_ = BK + 4;  // padded row stride (floats)
constexpr int STAGES = 4;
constexpr int NTHREADS = 256;
template <int BN>
constexpr size_t smem_bytes() {
  return (size_t)STAGES * (BM + BN) * LDS_ * sizeof(float);
}

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, int src_bytes) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

// operands of one implicit GEMM:  C[M, N] = A[M, K] * Bm[N, K]^T
//   A row p = output pixel (n, oh, ow); A col k = (kh, kw, ci) over the NHWC input (+ optional concat in2)
struct TileOperands {
  const float* in;
  const float* in2;
  int H, W, C1, C2, Cin;
  int KH, KW, stride, pad, OH, OW;
  const float* w;  // [N][K]
  int N, K, M;
};

template <int BN>
__device__ __forceinline__ void mainloop(const TileOperands& p, int m0, int n0, float* smem,
                                         float (&acc)[8][BN / 16]) {
  constexpr int TN = BN / 16;
  float* As = smem;                       // [STAGES][BM][LDS_]
  float* Bs = smem + STAGES * BM * LDS_;  // [STAGES][BN][LDS_]
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;

  // loader assignment: A has BM*4 float4 per stage (2 per thread), B has BN*4 (1 or 2 per thread)
  const int kq = tid & 3;
  const float* a_base[2];
  const float* a_base2[2];
  int a_ih0[2], a_iw0[2];
  bool a_ok[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    int r = (tid >> 2) + i * 64;
    int pidx = m0 + r;
    a_ok[i] = pidx < p.M;
    int pp = a_ok[i] ? pidx : 0;
    int n = pp / (p.OH * p.OW);
    int rem = pp - n * (p.OH * p.OW);
    int oh = rem / p.OW, ow = rem - oh * p.OW;
    a_ih0[i] = oh * p.stride - p.pad;
    a_iw0[i] = ow * p.stride - p.pad;
    a_base[i] = p.in + (size_t)n * p.H * p.W * p.C1;
    a_base2[i] = p.in2 ? p.in2 + (size_t)n * p.H * p.W * p.C2 : nullptr;
  }
  constexpr int NB = BN / 64;
  const float* b_ptr[NB];
  bool b_ok[NB];
#pragma unroll
  for (int i = 0; i < NB; ++i) {
    int co = n0 + (tid >> 2) + i * 64;
    b_ok[i] = co < p.N;
    b_ptr[i] = p.w + (size_t)(b_ok[i] ? co : 0) * p.K;
  }
  const int nk = (p.K + BK - 1) / BK;

  auto load_stage = [&](int kt, int stage) {
    int k0 = kt * BK + kq * 4;
    bool kin = k0 < p.K;
    int tap = 0, ci = k0;
    if (p.KH * p.KW > 1) {
      tap = k0 / p.Cin;
      ci = k0 - tap * p.Cin;
    }
    int kh = tap / p.KW, kw = tap - kh * p.KW;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      int ih = a_ih0[i] + kh, iw = a_iw0[i] + kw;
      bool ok = kin && a_ok[i] && ih >= 0 && ih < p.H && iw >= 0 && iw < p.W;
      const float* src = p.in;
      if (ok) {
        size_t pix = (size_t)ih * p.W + iw;
        src = (ci < p.C1) ? a_base[i] + pix * p.C1 + ci : a_base2[i] + pix * p.C2 + (ci - p.C1);
      }
      cp_async16(&As[(stage * BM + (tid >> 2) + i * 64) * LDS_ + kq * 4], src, ok ? 16 : 0);
    }
#pragma unroll
    for (int i = 0; i < NB; ++i) {
      bool ok = kin && b_ok[i];
      cp_async16(&Bs[(stage * BN + (tid >> 2) + i * 64) * LDS_ + kq * 4], ok ? b_ptr[i] + k0 : p.w, ok ? 16 : 0);
    }
  };

#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) {
    if (s < nk) load_stage(s, s);
    cp_async_commit();
  }
  for (int kt = 0; kt < nk; ++kt) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    {
      int nxt = kt + STAGES - 1;
      if (nxt < nk) load_stage(nxt, nxt % STAGES);
      cp_async_commit();
    }
    const float* as = As + (kt % STAGES) * BM * LDS_;
    const float* bs = Bs + (kt % STAGES) * BN * LDS_;
#pragma unroll
    for (int kk = 0; kk < BK; kk += 4) {
      float4 b4[TN];
#pragma unroll
      for (int j = 0; j < TN; ++j) b4[j] = *reinterpret_cast<const float4*>(&bs[(tx + 16 * j) * LDS_ + kk]);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float4 a4 = *reinterpret_cast<const float4*>(&as[(ty + 16 * i) * LDS_ + kk]);
#pragma unroll
        for (int j = 0; j < TN; ++j) {
          acc[i][j] = fmaf(a4.x, b4[j].x, acc[i][j]);
          acc[i][j] = fmaf(a4.y, b4[j].y, acc[i][j]);
          acc[i][j] = fmaf(a4.z, b4[j].z, acc[i][j]);
          acc[i][j] = fmaf(a4.w, b4[j].w, acc[i][j]);
        }
      }
    }
  }
  cp_async_wait<0>();
  __syncthreads();  // shared memory may be reused by the caller's epilogue
}

}  // namespace simt
}  // namespace gimb
