// coarse_match.cu - fused dual-softmax coarse matching (networks/loftr/utils/coarse_matching.py:88-259).
//
// The reference materialises sim [N,L,S], two softmaxes, their product and four more full passes
// (threshold, two max reductions, bool max).  Here the L x S matrix never exists in memory:
//
//   sweep 1  sim tile = (f0/sqrt(C)) (f1/sqrt(C))^T / T ; per-tile (max, sum exp) partials per row and per column
//   merge    partials -> row stats (softmax over dim 2) and column stats (softmax over dim 1)
//   sweep 2  recompute the tile, conf = softmax_col * softmax_row ; per row: (max conf, first argmax j) ;
//            per column: max conf  (order-independent atomics on the bit patterns of non-negative floats)
//   select   row i matches j iff conf > thr, border test, conf == colmax[j]  (mutual nearest neighbour)
//   compact  ordered by (b, i) exactly like torch.where (stable prefix sum, no host round trip)
//
// conf[i, j] is computed once, by one thread, with one formula; both maxima are taken over those
// same values, so the `conf == max` equalities of the reference hold by construction.
#include <algorithm>

#include <stdlib.h>
#include <string.h>

#include "corr_sweep.cuh"
#include "ops.cuh"
#include "simt_tile.cuh"

namespace gimb {
namespace {

using namespace simt;
constexpr int BN = 128;
constexpr int TN = BN / 16;

struct SweepArgs {
  const float* f0;
  const float* f1;
  int L, S, C;
  const uint8_t* mask0;
  const uint8_t* mask1;
  float inv_sqrt_c2;  // 1 / C  (both features divided by sqrt(C))
  float temperature;
  int tiles_m, tiles_n;
  float2* rowpart;  // [N][L][tiles_n]
  float2* colpart;  // [N][S][tiles_m]
  const float2* rowstat;  // [N][L]  (max, sum)
  const float2* colstat;  // [N][S]
  unsigned long long* rowbest;  // [N][L]  (conf bits << 32) | ~j
  unsigned int* colbest;        // [N][S]  conf bits
  float* conf_matrix;           // optional
};

__device__ __forceinline__ TileOperands make_ops(const SweepArgs& a, int b) {
  TileOperands t;
  t.in = a.f0 + (size_t)b * a.L * a.C;
  t.in2 = nullptr;
  t.H = a.L; t.W = 1; t.C1 = a.C; t.C2 = 0; t.Cin = a.C;
  t.KH = 1; t.KW = 1; t.stride = 1; t.pad = 0; t.OH = a.L; t.OW = 1;
  t.w = a.f1 + (size_t)b * a.S * a.C;
  t.N = a.S; t.K = a.C; t.M = a.L;
  return t;
}

__device__ __forceinline__ void merge_ms(float& m, float& s, float m2, float s2) {
  if (m2 == -INFINITY) return;
  if (m == -INFINITY) { m = m2; s = s2; return; }
  float nm = fmaxf(m, m2);
  s = s * expf(m - nm) + s2 * expf(m2 - nm);
  m = nm;
}

// sim value of one element, identical in both sweeps
__device__ __forceinline__ float sim_value(float dot, const SweepArgs& a, bool masked) {
  float v = __fdiv_rn(dot * a.inv_sqrt_c2, a.temperature);
  return masked ? -1e9f : v;
}

__global__ void __launch_bounds__(NTHREADS, 1) corr_stats_kernel(const SweepArgs a) {
  extern __shared__ __align__(16) float smem[];
  const int b = blockIdx.z;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float acc[8][TN];
  TileOperands t = make_ops(a, b);
  mainloop<BN>(t, m0, n0, smem, acc);

  bool rv[8], cv[TN], rmask[8], cmask[TN];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int r = m0 + ty + 16 * i;
    rv[i] = r < a.L;
    rmask[i] = a.mask0 ? (rv[i] ? a.mask0[(size_t)b * a.L + r] == 0 : true) : false;
  }
#pragma unroll
  for (int j = 0; j < TN; ++j) {
    int c = n0 + tx + 16 * j;
    cv[j] = c < a.S;
    cmask[j] = a.mask1 ? (cv[j] ? a.mask1[(size_t)b * a.S + c] == 0 : true) : false;
  }
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = sim_value(acc[i][j], a, rmask[i] || cmask[j]);

  // ---- row partials over this tile's columns
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float m = -INFINITY;
#pragma unroll
    for (int j = 0; j < TN; ++j)
      if (cv[j]) m = fmaxf(m, acc[i][j]);
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < TN; ++j)
      if (cv[j]) s += expf(acc[i][j] - m);
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (tx == 0 && rv[i]) a.rowpart[((size_t)b * a.L + m0 + ty + 16 * i) * a.tiles_n + blockIdx.y] = make_float2(m, s);
  }
  // ---- column partials over this tile's rows: per-thread, then across the 16 ty groups via smem
  float2* red = reinterpret_cast<float2*>(smem);  // [16][BN]
#pragma unroll
  for (int j = 0; j < TN; ++j) {
    float m = -INFINITY;
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (rv[i]) m = fmaxf(m, acc[i][j]);
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (rv[i]) s += expf(acc[i][j] - m);
    red[ty * BN + tx + 16 * j] = make_float2(m, s);
  }
  __syncthreads();
  if (threadIdx.x < BN) {
    int c = n0 + threadIdx.x;
    if (c < a.S) {
      float m = -INFINITY, s = 0.f;
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        float2 v = red[k * BN + threadIdx.x];
        merge_ms(m, s, v.x, v.y);
      }
      a.colpart[((size_t)b * a.S + c) * a.tiles_m + blockIdx.x] = make_float2(m, s);
    }
  }
}

__global__ void merge_stats_kernel(const float2* __restrict__ part, int ntiles, float2* __restrict__ stat, size_t n) {
  size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (idx >= n) return;
  float m = -INFINITY, s = 0.f;
  for (int t = 0; t < ntiles; ++t) {
    float2 v = part[idx * ntiles + t];
    merge_ms(m, s, v.x, v.y);
  }
  stat[idx] = make_float2(m, s);
}

__global__ void __launch_bounds__(NTHREADS, 1) corr_conf_kernel(const SweepArgs a) {
  extern __shared__ __align__(16) float smem[];
  const int b = blockIdx.z;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float acc[8][TN];
  TileOperands t = make_ops(a, b);
  mainloop<BN>(t, m0, n0, smem, acc);

  bool rv[8], cv[TN], rmask[8], cmask[TN];
  float2 rs[8], cs[TN];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int r = m0 + ty + 16 * i;
    rv[i] = r < a.L;
    rmask[i] = a.mask0 ? (rv[i] ? a.mask0[(size_t)b * a.L + r] == 0 : true) : false;
    rs[i] = rv[i] ? a.rowstat[(size_t)b * a.L + r] : make_float2(0.f, 1.f);
  }
#pragma unroll
  for (int j = 0; j < TN; ++j) {
    int c = n0 + tx + 16 * j;
    cv[j] = c < a.S;
    cmask[j] = a.mask1 ? (cv[j] ? a.mask1[(size_t)b * a.S + c] == 0 : true) : false;
    cs[j] = cv[j] ? a.colstat[(size_t)b * a.S + c] : make_float2(0.f, 1.f);
  }
  // conf = softmax(sim, dim=1) * softmax(sim, dim=2)   (coarse_matching.py:118)
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      float sim = sim_value(acc[i][j], a, rmask[i] || cmask[j]);
      float p_col = __fdiv_rn(expf(sim - cs[j].x), cs[j].y);
      float p_row = __fdiv_rn(expf(sim - rs[i].x), rs[i].y);
      acc[i][j] = (rv[i] && cv[j]) ? p_col * p_row : 0.f;
    }
  if (a.conf_matrix) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < TN; ++j)
        if (rv[i] && cv[j])
          a.conf_matrix[((size_t)b * a.L + m0 + ty + 16 * i) * a.S + n0 + tx + 16 * j] = acc[i][j];
  }
  // ---- per row: max conf and the FIRST column attaining it
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    unsigned long long best = 0ull;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      if (!cv[j]) continue;
      unsigned int col = (unsigned int)(n0 + tx + 16 * j);
      unsigned long long pk = ((unsigned long long)__float_as_uint(acc[i][j]) << 32) | (unsigned long long)(~col);
      best = pk > best ? pk : best;
    }
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) {
      unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
      best = other > best ? other : best;
    }
    if (tx == 0 && rv[i]) atomicMax(&a.rowbest[(size_t)b * a.L + m0 + ty + 16 * i], best);
  }
  // ---- per column: max conf
  unsigned int* red = reinterpret_cast<unsigned int*>(smem);  // [16][BN]
#pragma unroll
  for (int j = 0; j < TN; ++j) {
    float m = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) m = fmaxf(m, acc[i][j]);
    red[ty * BN + tx + 16 * j] = __float_as_uint(m);
  }
  __syncthreads();
  if (threadIdx.x < BN) {
    int c = n0 + threadIdx.x;
    if (c < a.S) {
      unsigned int m = 0;
#pragma unroll
      for (int k = 0; k < 16; ++k) m = max(m, red[k * BN + threadIdx.x]);
      atomicMax(&a.colbest[(size_t)b * a.S + c], m);
    }
  }
}

// ---- padded-extent helper of mask_border_with_padding (coarse_matching.py:38-39):
// hs = max over columns of the column sums, ws = max over rows of the row sums.
__global__ void mask_extent_kernel(const uint8_t* __restrict__ mask, int h, int w, int* __restrict__ ext /*[N][2]*/) {
  const int b = blockIdx.x;
  const uint8_t* m = mask + (size_t)b * h * w;
  __shared__ int best_h, best_w;
  if (threadIdx.x == 0) { best_h = 0; best_w = 0; }
  __syncthreads();
  for (int x = threadIdx.x; x < w; x += blockDim.x) {
    int s = 0;
    for (int y = 0; y < h; ++y) s += m[y * w + x] != 0;
    atomicMax(&best_h, s);
  }
  for (int y = threadIdx.x; y < h; y += blockDim.x) {
    int s = 0;
    for (int x = 0; x < w; ++x) s += m[y * w + x] != 0;
    atomicMax(&best_w, s);
  }
  __syncthreads();
  if (threadIdx.x == 0) { ext[b * 2 + 0] = best_h; ext[b * 2 + 1] = best_w; }
}

struct SelectArgs {
  const unsigned long long* rowbest;
  const unsigned int* colbest;
  int N, L, S, h0c, w0c, h1c, w1c, border;
  float thr;
  const int* ext0;  // [N][2] (hs, ws) or null
  const int* ext1;
  float cscale;     // hw0_i[0] / hw0_c[0]
  const float* scale0;
  const float* scale1;
  int* block_counts;
  int64_t *b_ids, *i_ids, *j_ids;
  float *mconf, *mkpts0_c, *mkpts1_c;
  int64_t* count;
  int64_t capacity;  // rows available in the output arrays
};

// python slice semantics of `m[start:] = v` for a possibly negative start
__device__ __forceinline__ int py_start(int start, int len) {
  if (start < 0) start += len;
  return start < 0 ? 0 : start;
}

__device__ __forceinline__ bool row_selected(const SelectArgs& a, int b, int i, int& j_out, float& conf_out) {
  unsigned long long pk = a.rowbest[(size_t)b * a.L + i];
  unsigned int bits = (unsigned int)(pk >> 32);
  int j = (int)(~(unsigned int)(pk & 0xffffffffull));
  float conf = __uint_as_float(bits);
  j_out = j;
  conf_out = conf;
  if (!(conf > a.thr)) return false;
  if (j < 0 || j >= a.S) return false;
  if (a.colbest[(size_t)b * a.S + j] != bits) return false;
  int r0 = i / a.w0c, c0 = i - r0 * a.w0c;
  int r1 = j / a.w1c, c1 = j - r1 * a.w1c;
  int bd = a.border;
  if (bd > 0) {
    if (r0 < bd || c0 < bd || r1 < bd || c1 < bd) return false;
    if (a.ext0) {
      int h0 = a.ext0[b * 2], w0 = a.ext0[b * 2 + 1], h1 = a.ext1[b * 2], w1 = a.ext1[b * 2 + 1];
      if (r0 >= py_start(h0 - bd, a.h0c) || c0 >= py_start(w0 - bd, a.w0c) ||
          r1 >= py_start(h1 - bd, a.h1c) || c1 >= py_start(w1 - bd, a.w1c))
        return false;
    } else {
      if (r0 >= py_start(a.h0c - bd, a.h0c) || c0 >= py_start(a.w0c - bd, a.w0c) ||
          r1 >= py_start(a.h1c - bd, a.h1c) || c1 >= py_start(a.w1c - bd, a.w1c))
        return false;
    }
  }
  return true;
}

__global__ void __launch_bounds__(256) select_count_kernel(const SelectArgs a) {
  int idx = blockIdx.x * 256 + threadIdx.x;
  bool f = false;
  if (idx < a.N * a.L) {
    int j; float c;
    f = row_selected(a, idx / a.L, idx % a.L, j, c);
  }
  int cnt = __syncthreads_count(f);
  if (threadIdx.x == 0) a.block_counts[blockIdx.x] = cnt;
}

// exclusive scan of the block counts (single CTA, sequential over chunks of 1024)
__global__ void __launch_bounds__(1024) scan_blocks_kernel(int* __restrict__ counts, int n, int64_t* __restrict__ total) {
  __shared__ int warp_tot[32];
  __shared__ int carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int base = 0; base < n; base += 1024) {
    int i = base + threadIdx.x;
    int v = i < n ? counts[i] : 0;
    int x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) warp_tot[warp] = x;
    __syncthreads();
    if (warp == 0) {
      int w = warp_tot[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        int y = __shfl_up_sync(0xffffffffu, w, o);
        if (lane >= o) w += y;
      }
      warp_tot[lane] = w;
    }
    __syncthreads();
    int excl = x - v + (warp ? warp_tot[warp - 1] : 0) + carry;
    if (i < n) counts[i] = excl;
    __syncthreads();
    if (threadIdx.x == 1023) carry = excl + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) *total = carry;
}

__global__ void __launch_bounds__(256) select_scatter_kernel(const SelectArgs a) {
  __shared__ int warp_tot[8];
  int idx = blockIdx.x * 256 + threadIdx.x;
  int b = 0, i = 0, j = 0;
  float conf = 0.f;
  bool f = false;
  if (idx < a.N * a.L) {
    b = idx / a.L;
    i = idx - b * a.L;
    f = row_selected(a, b, i, j, conf);
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned int ballot = __ballot_sync(0xffffffffu, f);
  int in_warp = __popc(ballot & ((1u << lane) - 1));
  if (lane == 0) warp_tot[warp] = __popc(ballot);
  __syncthreads();
  int off = a.block_counts[blockIdx.x];
  for (int w = 0; w < warp; ++w) off += warp_tot[w];
  // rows beyond the caller's capacity are counted (the host reports the overflow) but never written
  if (f && (int64_t)off + in_warp < a.capacity) {
    int64_t pos = (int64_t)off + in_warp;
    a.b_ids[pos] = b;
    a.i_ids[pos] = i;
    a.j_ids[pos] = j;
    a.mconf[pos] = conf;
    float s0x = a.cscale, s0y = a.cscale, s1x = a.cscale, s1y = a.cscale;
    if (a.scale0) {
      s0x = a.cscale * a.scale0[b * 2]; s0y = a.cscale * a.scale0[b * 2 + 1];
      s1x = a.cscale * a.scale1[b * 2]; s1y = a.cscale * a.scale1[b * 2 + 1];
    }
    a.mkpts0_c[pos * 2 + 0] = (float)(i % a.w0c) * s0x;
    a.mkpts0_c[pos * 2 + 1] = (float)(i / a.w0c) * s0y;
    a.mkpts1_c[pos * 2 + 0] = (float)(j % a.w1c) * s1x;
    a.mkpts1_c[pos * 2 + 1] = (float)(j / a.w1c) * s1y;
  }
}

}  // namespace

int coarse_match(Ctx& ctx, const CoarseMatchArgs& c) {
  GIMB_CHECK(c.C % 16 == 0, "coarse_match: C must be a multiple of 16");
  GIMB_CHECK((c.mask0 == nullptr) == (c.mask1 == nullptr), "coarse_match: mask0/mask1 go together");
  GIMB_CHECK((c.scale0 == nullptr) == (c.scale1 == nullptr), "coarse_match: scale0/scale1 go together");
  // tensor-core sweeps whenever the split planes are given; an optional conf_matrix is then written by the SAME conf
  // sweep that produces the matches (debug epilogue), so the tap and the ids come from one numerical path
  const bool tc = c.planes0 != nullptr && c.planes1 != nullptr;
  // second-generation sweeps (corr_sweep.cu) unless the caller asks for the exact online-max sweeps (fallback after a
  // range flag, or GIMB_CORR=exact)
  static int env_exact = -1;
  if (env_exact < 0) {
    const char* e = getenv("GIMB_CORR");
    env_exact = (e && strcmp(e, "exact") == 0) ? 1 : 0;
  }
  const bool fast = tc && !c.exact && !env_exact && corr_sweep_supported(c.C) && c.planes0->ld == c.C && c.planes1->ld == c.C;
  int tiles_m = cdiv(c.L, BM), tiles_n = cdiv(c.S, BN);  // partials per column / per row
  if (tc) umma_corr_parts(c.L, c.S, &tiles_n, &tiles_m);
  const size_t NL = (size_t)c.N * c.L, NS = (size_t)c.N * c.S;
  size_t mark = ctx.arena.mark();
  // workspace of the exact sweeps (also planned in dry mode: the fallback must always fit)
  float2* rowpart = ctx.arena.alloc<float2>(NL * tiles_n);
  float2* colpart = ctx.arena.alloc<float2>(NS * tiles_m);
  float2* rowstat = ctx.arena.alloc<float2>(NL);
  float2* colstat = ctx.arena.alloc<float2>(NS);
  unsigned long long* rowbest = ctx.arena.alloc<unsigned long long>(NL);
  unsigned int* colbest = ctx.arena.alloc<unsigned int>(NS);
  const int nblocks = (int)cdiv64((int64_t)NL, 256);
  int* block_counts = ctx.arena.alloc<int>(nblocks);
  int* ext = ctx.arena.alloc<int>((size_t)c.N * 4);
  // workspace of the second-generation sweeps
  CorrSweep cs;
  if (tc && corr_sweep_supported(c.C)) {
    int rp, cp, Lp, Sp;
    corr_sweep_parts(ctx, c.N, c.L, c.S, c.C, &rp, &cp, &Lp, &Sp);
    cs.normsq = ctx.arena.alloc<float>(2 * (size_t)c.N);
    cs.rowpart = ctx.arena.alloc<float>((size_t)rp * c.N * Lp);
    cs.colpart = ctx.arena.alloc<float>((size_t)cp * c.N * Sp);
    cs.rowstat = ctx.arena.alloc<float2>((size_t)c.N * Lp);
    cs.colthr = ctx.arena.alloc<float>((size_t)c.N * Sp);
    cs.colsum = ctx.arena.alloc<float>((size_t)c.N * Sp);
  }
  if (!ctx.dry && c.N > 0) {
    GIMB_CHECK(!ctx.arena.overflow, "coarse_match: workspace exhausted");
    if (c.range_flag && !fast) GIMB_CUDA(cudaMemsetAsync(c.range_flag, 0, sizeof(int), ctx.stream));
    if (fast) {
      GIMB_CHECK(c.range_flag != nullptr, "coarse_match: the fast sweeps need a range flag");
      cs.f0 = *c.planes0; cs.f1 = *c.planes1; cs.f0_f32 = c.f0; cs.f1_f32 = c.f1;
      cs.N = c.N; cs.L = c.L; cs.S = c.S; cs.C = c.C;
      cs.mask0 = c.mask0; cs.mask1 = c.mask1; cs.temperature = c.temperature; cs.thr = c.thr;
      cs.flag = c.range_flag; cs.rowbest = rowbest; cs.colbest = colbest; cs.conf_matrix = c.conf_matrix;
      GIMB_TRY(corr_sweeps(ctx, cs));
    } else {
    GIMB_SMEM_OPTIN(corr_stats_kernel, smem_bytes<BN>());
    GIMB_SMEM_OPTIN(corr_conf_kernel, smem_bytes<BN>());
    SweepArgs a;
    a.f0 = c.f0; a.f1 = c.f1; a.L = c.L; a.S = c.S; a.C = c.C;
    a.mask0 = c.mask0; a.mask1 = c.mask1;
    a.inv_sqrt_c2 = 1.f / (float)c.C;
    a.temperature = c.temperature;
    a.tiles_m = tiles_m; a.tiles_n = tiles_n;
    a.rowpart = rowpart; a.colpart = colpart; a.rowstat = rowstat; a.colstat = colstat;
    a.rowbest = rowbest; a.colbest = colbest; a.conf_matrix = c.conf_matrix;
    dim3 grid(cdiv(c.L, BM), cdiv(c.S, BN), c.N);
    UmmaCorr uc;
    if (tc) {
      uc.f0 = *c.planes0; uc.f1 = *c.planes1; uc.N = c.N; uc.L = c.L; uc.S = c.S; uc.C = c.C;
      uc.mask0 = c.mask0; uc.mask1 = c.mask1; uc.temperature = c.temperature; uc.thr = c.thr;
      uc.rowpart = rowpart; uc.colpart = colpart; uc.rowstat = rowstat; uc.colstat = colstat;
      uc.rowbest = rowbest; uc.colbest = colbest; uc.conf_matrix = c.conf_matrix;
      GIMB_TRY(umma_corr(ctx, uc, 0));
    } else {
      corr_stats_kernel<<<grid, NTHREADS, smem_bytes<BN>(), ctx.stream>>>(a);
      GIMB_LAUNCH_CHECK();
    }
    ctx.mark("corr_stats");
    merge_stats_kernel<<<(unsigned)cdiv64(NL, 256), 256, 0, ctx.stream>>>(rowpart, tiles_n, rowstat, NL);
    GIMB_LAUNCH_CHECK();
    merge_stats_kernel<<<(unsigned)cdiv64(NS, 256), 256, 0, ctx.stream>>>(colpart, tiles_m, colstat, NS);
    GIMB_LAUNCH_CHECK();
    GIMB_CUDA(cudaMemsetAsync(rowbest, 0, NL * sizeof(unsigned long long), ctx.stream));
    GIMB_CUDA(cudaMemsetAsync(colbest, 0, NS * sizeof(unsigned int), ctx.stream));
    ctx.mark("corr_merge");
    if (tc) {
      GIMB_TRY(umma_corr(ctx, uc, 1));
    } else {
      corr_conf_kernel<<<grid, NTHREADS, smem_bytes<BN>(), ctx.stream>>>(a);
      GIMB_LAUNCH_CHECK();
    }
    ctx.mark("corr_conf");
    ctx.launches += 4;
    }

    SelectArgs s;
    s.rowbest = rowbest; s.colbest = colbest;
    s.N = c.N; s.L = c.L; s.S = c.S; s.h0c = c.h0c; s.w0c = c.w0c; s.h1c = c.h1c; s.w1c = c.w1c;
    s.border = c.border; s.thr = c.thr;
    s.ext0 = s.ext1 = nullptr;
    if (c.mask0) {
      mask_extent_kernel<<<c.N, 128, 0, ctx.stream>>>(c.mask0, c.h0c, c.w0c, ext);
      GIMB_LAUNCH_CHECK();
      mask_extent_kernel<<<c.N, 128, 0, ctx.stream>>>(c.mask1, c.h1c, c.w1c, ext + 2 * c.N);
      GIMB_LAUNCH_CHECK();
      s.ext0 = ext; s.ext1 = ext + 2 * c.N;
      ctx.launches += 2;
    }
    s.cscale = (float)c.H0 / (float)c.h0c;
    s.scale0 = c.scale0; s.scale1 = c.scale1;
    s.block_counts = block_counts;
    s.b_ids = c.b_ids; s.i_ids = c.i_ids; s.j_ids = c.j_ids;
    s.mconf = c.mconf; s.mkpts0_c = c.mkpts0_c; s.mkpts1_c = c.mkpts1_c; s.count = c.count;
    s.capacity = c.capacity;
    select_count_kernel<<<nblocks, 256, 0, ctx.stream>>>(s);
    GIMB_LAUNCH_CHECK();
    scan_blocks_kernel<<<1, 1024, 0, ctx.stream>>>(block_counts, nblocks, c.count);
    GIMB_LAUNCH_CHECK();
    select_scatter_kernel<<<nblocks, 256, 0, ctx.stream>>>(s);
    GIMB_LAUNCH_CHECK();
    ctx.launches += 3;
  }
  ctx.arena.release(mark);
  return 0;
}

}  // namespace gimb
