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
  // 256 threads = 4 row-subsets x (8 d-quads x 8 v-quads): each thread owns a 4 x 4 block of KV[d][v] for the rows
  // r = sub (mod 4) of every staged chunk; the four subsets are summed through shared memory in a fixed order.
  __shared__ __align__(16) float Ks[KV_CHUNK][32];
  __shared__ __align__(16) float Vs[KV_CHUNK][32];
  __shared__ float red[3][32 * 32 + 32];
  const int bh = blockIdx.x, split = blockIdx.y;
  const int b = bh / nhead, h = bh - b * nhead;
  const int tid = threadIdx.x;
  const int sub = tid >> 6, dq = (tid & 63) >> 3, vq = tid & 7;
  const int per = (S + nsplit - 1) / nsplit;
  const int s_begin = split * per, s_end = min(S, s_begin + per);
  const float* base = kv + (size_t)b * S * 2 * C;
  float acc[4][4];
  float ksum[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  // 32 rows x (32 K + 32 V) floats = 512 float4 per chunk: 2 per thread, fetched one chunk ahead into registers so the
  // global-load latency overlaps the previous chunk's FMAs (arithmetic order unchanged)
  auto fetch = [&](int s0, float4 (&t)[2]) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int idx = tid + i * 256;
      const int r = idx >> 4, q4 = idx & 15;  // row, float4 index (0..7 -> K, 8..15 -> V)
      const int s = s0 + r;
      t[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (s < s_end) {
        const float* rowp = base + (size_t)s * 2 * C + (q4 < 8 ? h * 32 + q4 * 4 : C + h * 32 + (q4 - 8) * 4);
        t[i] = *reinterpret_cast<const float4*>(rowp);
      }
    }
  };
  float4 pre[2];
  fetch(s_begin, pre);
  for (int s0 = s_begin; s0 < s_end; s0 += KV_CHUNK) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int idx = tid + i * 256;
      const int r = idx >> 4, q4 = idx & 15;
      if (q4 < 8) *reinterpret_cast<float4*>(&Ks[r][q4 * 4]) = pre[i];
      else *reinterpret_cast<float4*>(&Vs[r][(q4 - 8) * 4]) = pre[i];
    }
    __syncthreads();
    if (s0 + KV_CHUNK < s_end) fetch(s0 + KV_CHUNK, pre);
#pragma unroll
    for (int r = sub; r < KV_CHUNK; r += 4) {
      const float4 k4 = *reinterpret_cast<const float4*>(&Ks[r][dq * 4]);
      const float4 v4 = *reinterpret_cast<const float4*>(&Vs[r][vq * 4]);
      const float kk[4] = {k4.x, k4.y, k4.z, k4.w};
      const float vv[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(kk[i], vv[j], acc[i][j]);
        ksum[i] += kk[i];
      }
    }
    __syncthreads();
  }
  // fixed-order reduction over the 4 row subsets
  if (sub > 0) {
    float* rp = red[sub - 1];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
#pragma unroll
      for (int j = 0; j < 4; ++j) rp[(dq * 4 + i) * 32 + vq * 4 + j] = acc[i][j];
      if (vq == 0) rp[1024 + dq * 4 + i] = ksum[i];
    }
  }
  __syncthreads();
  if (sub == 0) {
    float* o = part + ((size_t)bh * nsplit + split) * KV_STRIDE;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float4 r4;
      float* rr = &r4.x;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int e = (dq * 4 + i) * 32 + vq * 4 + j;
        rr[j] = ((acc[i][j] + red[0][e]) + red[1][e]) + red[2][e];
      }
      *reinterpret_cast<float4*>(o + (dq * 4 + i) * 32 + vq * 4) = r4;
      if (vq == 0) {
        const int e = 1024 + dq * 4 + i;
        o[e] = ((ksum[i] + red[0][e]) + red[1][e]) + red[2][e];
      }
    }
  }
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
// 256 threads = 8 heads x 32 value channels; each thread keeps its KV column and the head's Ksum in registers
// and streams rows of Q' from shared memory (float4 broadcast reads): 64 FFMA per 8 LDS.128.
constexpr int AP_ROWS = 32;
__global__ void __launch_bounds__(256) attn_apply_kernel(const float* __restrict__ q, const float* __restrict__ kvf,
                                                         int L, int C, int nhead, float vlen, int rows_per_cta,
                                                         float* __restrict__ msg, const PlanesDev sp) {
  __shared__ __align__(16) float Qs[AP_ROWS][256];
  const int b = blockIdx.y;
  const int tid = threadIdx.x, v = tid & 31, h = tid >> 5;
  float kvr[32], ksr[32];
  {
    const float* kvh = kvf + ((size_t)b * nhead + h) * KV_STRIDE;
#pragma unroll
    for (int d = 0; d < 32; ++d) { kvr[d] = kvh[d * 32 + v]; ksr[d] = kvh[1024 + d]; }
  }
  const int l0 = blockIdx.x * rows_per_cta;
  const int l1 = min(L, l0 + rows_per_cta);
  for (int lc = l0; lc < l1; lc += AP_ROWS) {
    const int nr = min(AP_ROWS, l1 - lc);
    __syncthreads();
    for (int i = tid; i < AP_ROWS * 64; i += 256) {
      const int r = i >> 6, c4 = i & 63;
      float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r < nr) t = *reinterpret_cast<const float4*>(q + ((size_t)b * L + lc + r) * C + c4 * 4);
      *reinterpret_cast<float4*>(&Qs[r][c4 * 4]) = t;
    }
    __syncthreads();
    // four rows at a time: eight independent FMA chains per thread (one row at a time left the two 32-deep dependent
    // chains latency-bound); each row's summation order over d is unchanged
    for (int r = 0; r < nr; r += 4) {
      float o[4] = {0.f, 0.f, 0.f, 0.f}, z[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int d4 = 0; d4 < 8; ++d4) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4 q4 = *reinterpret_cast<const float4*>(&Qs[r + i][h * 32 + d4 * 4]);
          o[i] = fmaf(q4.x, kvr[d4 * 4 + 0], o[i]); z[i] = fmaf(q4.x, ksr[d4 * 4 + 0], z[i]);
          o[i] = fmaf(q4.y, kvr[d4 * 4 + 1], o[i]); z[i] = fmaf(q4.y, ksr[d4 * 4 + 1], z[i]);
          o[i] = fmaf(q4.z, kvr[d4 * 4 + 2], o[i]); z[i] = fmaf(q4.z, ksr[d4 * 4 + 2], z[i]);
          o[i] = fmaf(q4.w, kvr[d4 * 4 + 3], o[i]); z[i] = fmaf(q4.w, ksr[d4 * 4 + 3], z[i]);
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (r + i < nr) {
          const float res = o[i] * (1.f / (z[i] + 1e-6f)) * vlen;
          const size_t grow = (size_t)b * L + lc + r + i;
          if (msg) msg[grow * C + tid] = res;
          if (sp.hi) {
            const __half hh = __float2half_rn(res);
            sp.hi[grow * sp.ld + tid] = hh;
            sp.lo[grow * sp.ld + tid] = __float2half_rn((res - __half2float(hh)) * kSplitScale);
          }
        }
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
  // fixed split of the S reduction (a function of S only): results do not depend on the batch size, so a pair gives
  // bit-identical features whether it is processed alone or in a batch
  int nsplit = std::max(1, std::min(32, cdiv(S, 16 * KV_CHUNK)));
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
    const int rows_per_cta = 128;
    GIMB_CHECK(C == 256 && nhead == 8, "linear_attention: coarse flavour is built for C = 256, 8 heads");
    attn_apply_kernel<<<dim3(cdiv(L, rows_per_cta), B), 256, 0, ctx.stream>>>(q, kvf, L, C, nhead, (float)S,
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
