// transformer.cu - the memory-bound pieces of LoFTREncoderLayer (networks/loftr/submodules/transformer.py:35-58):
// LayerNorm (+ residual) and the linear-attention reductions (networks/loftr/submodules/attentions.py:31-47).
// The q/k/v/merge/MLP projections are GEMMs and go through conv_gemm().
#include <algorithm>

#include "ops.cuh"
#include "split.cuh"

namespace gimb {
namespace {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ------------------------------------------------------------------------------------------------
// LayerNorm: one warp per row, C = 32 * VPT, two-pass statistics in registers.
template <int VPT>
__global__ void __launch_bounds__(256) layernorm_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                        const float* __restrict__ beta, const float* __restrict__ res,
                                                        int64_t rows, float* __restrict__ out, const PlanesDev sp) {
  constexpr int C = 32 * VPT;
  const int lane = threadIdx.x & 31;
  int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float* xr = x + row * C;
  float v[VPT];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < VPT / 4; ++i) {
    float4 t = *reinterpret_cast<const float4*>(xr + (i * 32 + lane) * 4);
    v[i * 4 + 0] = t.x; v[i * 4 + 1] = t.y; v[i * 4 + 2] = t.z; v[i * 4 + 3] = t.w;
    s += (t.x + t.y) + (t.z + t.w);
  }
  float mean = warp_sum(s) * (1.f / C);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < VPT; ++i) {
    float d = v[i] - mean;
    q = fmaf(d, d, q);
  }
  float var = warp_sum(q) * (1.f / C);
  float rstd = 1.f / sqrtf(var + 1e-5f);
#pragma unroll
  for (int i = 0; i < VPT / 4; ++i) {
    int c = (i * 32 + lane) * 4;
    float4 g = *reinterpret_cast<const float4*>(gamma + c);
    float4 b = *reinterpret_cast<const float4*>(beta + c);
    float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
    if (res) r = *reinterpret_cast<const float4*>(res + row * C + c);
    float4 o;
    o.x = r.x + fmaf((v[i * 4 + 0] - mean) * rstd, g.x, b.x);
    o.y = r.y + fmaf((v[i * 4 + 1] - mean) * rstd, g.y, b.y);
    o.z = r.z + fmaf((v[i * 4 + 2] - mean) * rstd, g.z, b.z);
    o.w = r.w + fmaf((v[i * 4 + 3] - mean) * rstd, g.w, b.w);
    if (out) *reinterpret_cast<float4*>(out + row * C + c) = o;
    if (sp.hi) split4_store(sp, (size_t)row * sp.ld + c, o.x, o.y, o.z, o.w);
  }
}

// ------------------------------------------------------------------------------------------------
// Linear attention, coarse flavour (D = 32).
// step 1: per (image b, head h, S-split) partial KV[32][32] = sum_s K'[s,d] V[s,v] and Ksum[32].
//         deterministic: partials are written, not atomically accumulated.
constexpr int KV_CHUNK = 32;                 // rows of K/V staged per iteration
constexpr int KV_STRIDE = 32 * 32 + 32;      // floats per (b, h) result

__global__ void __launch_bounds__(256) kv_partial_kernel(const float* __restrict__ kv, int S, int C, int nhead,
                                                         int nsplit, float* __restrict__ part) {
  __shared__ float Ks[KV_CHUNK][33];
  __shared__ float Vs[KV_CHUNK][32];
  const int bh = blockIdx.x, split = blockIdx.y;
  const int b = bh / nhead, h = bh - b * nhead;
  const int tid = threadIdx.x;
  const int d = tid >> 3, v0 = (tid & 7) * 4;
  const int per = (S + nsplit - 1) / nsplit;
  const int s_begin = split * per, s_end = min(S, s_begin + per);
  const float* base = kv + (size_t)b * S * 2 * C;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  float ksum = 0.f;
  for (int s0 = s_begin; s0 < s_end; s0 += KV_CHUNK) {
    // 32 rows x (32 K + 32 V) floats = 512 float4: 2 per thread
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      int idx = tid + i * 256;       // 0..511
      int r = idx >> 4, q4 = idx & 15;  // row, float4 index (0..7 -> K, 8..15 -> V)
      int s = s0 + r;
      float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
      if (s < s_end) {
        const float* rowp = base + (size_t)s * 2 * C + (q4 < 8 ? h * 32 + q4 * 4 : C + h * 32 + (q4 - 8) * 4);
        t = *reinterpret_cast<const float4*>(rowp);
      }
      if (q4 < 8) {
        Ks[r][q4 * 4 + 0] = t.x; Ks[r][q4 * 4 + 1] = t.y; Ks[r][q4 * 4 + 2] = t.z; Ks[r][q4 * 4 + 3] = t.w;
      } else {
        *reinterpret_cast<float4*>(&Vs[r][(q4 - 8) * 4]) = t;
      }
    }
    __syncthreads();
#pragma unroll 8
    for (int r = 0; r < KV_CHUNK; ++r) {
      float k = Ks[r][d];
      float4 vv = *reinterpret_cast<const float4*>(&Vs[r][v0]);
      acc[0] = fmaf(k, vv.x, acc[0]);
      acc[1] = fmaf(k, vv.y, acc[1]);
      acc[2] = fmaf(k, vv.z, acc[2]);
      acc[3] = fmaf(k, vv.w, acc[3]);
      ksum += k;
    }
    __syncthreads();
  }
  float* o = part + ((size_t)bh * nsplit + split) * KV_STRIDE;
  *reinterpret_cast<float4*>(o + d * 32 + v0) = make_float4(acc[0], acc[1], acc[2], acc[3]);
  if ((tid & 7) == 0) o[1024 + d] = ksum;
}

// step 2: fixed-order reduction of the partials -> kvf[bh][1056]
__global__ void kv_reduce_kernel(const float* __restrict__ part, int nsplit, float* __restrict__ kvf, int total) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  int bh = idx / KV_STRIDE, e = idx - bh * KV_STRIDE;
  float s = 0.f;
  for (int i = 0; i < nsplit; ++i) s += part[((size_t)bh * nsplit + i) * KV_STRIDE + e];
  kvf[idx] = s;
}

// step 3: msg[l, h*32+v] = (sum_d Q'[l,h,d] KV[h,d,v]) * (1 / (sum_d Q'[l,h,d] Ksum[h,d] + eps)) * S
// one CTA = 8 warps, each warp walks rows; KV of all heads of image b lives in shared memory.
__global__ void __launch_bounds__(256) attn_apply_kernel(const float* __restrict__ q, const float* __restrict__ kvf,
                                                         int L, int C, int nhead, float vlen, int rows_per_cta,
                                                         float* __restrict__ msg, const PlanesDev sp) {
  extern __shared__ float sm[];  // [nhead][KV_STRIDE]
  const int b = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < nhead * KV_STRIDE; i += 256) sm[i] = kvf[(size_t)b * nhead * KV_STRIDE + i];
  __syncthreads();
  const int l0 = blockIdx.x * rows_per_cta;
  const int l1 = min(L, l0 + rows_per_cta);
  for (int l = l0 + warp; l < l1; l += 8) {
    const float* qr = q + ((size_t)b * L + l) * C;
    const size_t grow = (size_t)b * L + l;
    for (int h = 0; h < nhead; ++h) {
      const float* kvh = sm + h * KV_STRIDE;
      float qv = qr[h * 32 + lane];
      float zden = warp_sum(qv * kvh[1024 + lane]);
      float o = 0.f;
#pragma unroll
      for (int d = 0; d < 32; ++d) o = fmaf(__shfl_sync(0xffffffffu, qv, d), kvh[d * 32 + lane], o);
      float z = 1.f / (zden + 1e-6f);
      const float r = o * z * vlen;
      if (msg) msg[grow * C + h * 32 + lane] = r;
      if (sp.hi) {
        const __half hh = __float2half_rn(r);
        sp.hi[grow * sp.ld + h * 32 + lane] = hh;
        sp.lo[grow * sp.ld + h * 32 + lane] = __float2half_rn((r - __half2float(hh)) * kSplitScale);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Fine-level linear attention: sequences of WW (=25) tokens, C = 128, 8 heads x 16.  One CTA of C
// threads per match; thread c owns channel c: (head h = c / 16, value index v = c % 16).
template <int C, int D, int MAXWW>
__global__ void __launch_bounds__(C) fine_attn_kernel(const float* __restrict__ q, const float* __restrict__ kv,
                                                      int64_t M, int WW, float* __restrict__ msg, const PlanesDev sp) {
  __shared__ float Qs[MAXWW][C];
  __shared__ float Ks[MAXWW][C];
  const int c = threadIdx.x;
  const int hb = (c / D) * D;  // first channel of this head
  for (int64_t m = blockIdx.x; m < M; m += gridDim.x) {
    const float* qm = q + m * WW * C;
    const float* kvm = kv + m * WW * 2 * C;
    float vreg[MAXWW];
#pragma unroll
    for (int s = 0; s < MAXWW; ++s) {
      if (s < WW) {
        Qs[s][c] = qm[s * C + c];
        Ks[s][c] = kvm[s * 2 * C + c];
        vreg[s] = kvm[s * 2 * C + C + c];
      } else {
        vreg[s] = 0.f;
      }
    }
    __syncthreads();
    float kvacc[D];
    float ksum[D];
#pragma unroll
    for (int d = 0; d < D; ++d) { kvacc[d] = 0.f; ksum[d] = 0.f; }
#pragma unroll
    for (int s = 0; s < MAXWW; ++s) {
      if (s < WW) {
#pragma unroll
        for (int d = 0; d < D; ++d) {
          float k = Ks[s][hb + d];
          kvacc[d] = fmaf(k, vreg[s], kvacc[d]);
          ksum[d] += k;
        }
      }
    }
    for (int l = 0; l < WW; ++l) {
      float o = 0.f, zden = 0.f;
#pragma unroll
      for (int d = 0; d < D; ++d) {
        float qq = Qs[l][hb + d];
        o = fmaf(qq, kvacc[d], o);
        zden = fmaf(qq, ksum[d], zden);
      }
      const float r = o * (1.f / (zden + 1e-6f)) * (float)WW;
      if (msg) msg[(m * WW + l) * C + c] = r;
      if (sp.hi) {
        const __half hh = __float2half_rn(r);
        sp.hi[(m * WW + l) * sp.ld + c] = hh;
        sp.lo[(m * WW + l) * sp.ld + c] = __float2half_rn((r - __half2float(hh)) * kSplitScale);
      }
    }
    __syncthreads();
  }
}

}  // namespace

int layernorm(Ctx& ctx, const float* x, const float* gamma, const float* beta, const float* res, int64_t rows,
              int C, float* out, const SplitPlanes* planes) {
  const PlanesDev sp = planes ? dev(*planes) : PlanesDev{nullptr, nullptr, nullptr, 0};
  GIMB_CHECK(C == 256 || C == 128, "layernorm: C must be 128 or 256 (got %d)", C);
  if (ctx.dry || rows == 0) return 0;
  int blocks = (int)cdiv64(rows, 8);
  if (C == 256)
    layernorm_kernel<8><<<blocks, 256, 0, ctx.stream>>>(x, gamma, beta, res, rows, out, sp);
  else
    layernorm_kernel<4><<<blocks, 256, 0, ctx.stream>>>(x, gamma, beta, res, rows, out, sp);
  ctx.launches++;
  GIMB_LAUNCH_CHECK();
  return 0;
}

int linear_attention(Ctx& ctx, const float* q, const float* kv, int B, int L, int S, int C, int nhead, float* msg,
                     const SplitPlanes* planes) {
  const PlanesDev sp = planes ? dev(*planes) : PlanesDev{nullptr, nullptr, nullptr, 0};
  GIMB_CHECK(C == nhead * 32, "linear_attention: coarse flavour needs head dim 32");
  const int BH = B * nhead;
  int nsplit = std::max(1, std::min(cdiv(ctx.sm_count * 4, BH), cdiv(S, 4 * KV_CHUNK)));
  size_t mark = ctx.arena.mark();
  float* part = ctx.arena.alloc<float>((size_t)BH * nsplit * KV_STRIDE);
  float* kvf = ctx.arena.alloc<float>((size_t)BH * KV_STRIDE);
  if (!ctx.dry && BH > 0) {
    GIMB_CHECK(part && kvf, "linear_attention: workspace exhausted");
    kv_partial_kernel<<<dim3(BH, nsplit), 256, 0, ctx.stream>>>(kv, S, C, nhead, nsplit, part);
    GIMB_LAUNCH_CHECK();
    int total = BH * KV_STRIDE;
    kv_reduce_kernel<<<cdiv(total, 256), 256, 0, ctx.stream>>>(part, nsplit, kvf, total);
    GIMB_LAUNCH_CHECK();
    const int rows_per_cta = 64;
    size_t smem = (size_t)nhead * KV_STRIDE * sizeof(float);
    attn_apply_kernel<<<dim3(cdiv(L, rows_per_cta), B), 256, smem, ctx.stream>>>(q, kvf, L, C, nhead, (float)S,
                                                                                   rows_per_cta, msg, sp);
    GIMB_LAUNCH_CHECK();
    ctx.launches += 3;
  }
  ctx.arena.release(mark);
  return 0;
}

int fine_attention(Ctx& ctx, const float* q, const float* kv, int64_t M, int WW, int C, int nhead, float* msg,
                   const SplitPlanes* planes) {
  const PlanesDev sp = planes ? dev(*planes) : PlanesDev{nullptr, nullptr, nullptr, 0};
  GIMB_CHECK(C == 128 && nhead == 8 && WW <= 25, "fine_attention: built for C=128, 8 heads, WW<=25");
  if (ctx.dry || M == 0) return 0;
  int blocks = (int)std::min<int64_t>(M, (int64_t)ctx.sm_count * 32);
  fine_attn_kernel<128, 16, 25><<<blocks, 128, 0, ctx.stream>>>(q, kv, M, WW, msg, sp);
  ctx.launches++;
  GIMB_LAUNCH_CHECK();
  return 0;
}

}  // namespace gimb
