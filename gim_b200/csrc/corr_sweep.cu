// corr_sweep.cu - dual-softmax coarse correlation on the 5th-gen tensor cores, second generation
// (networks/loftr/utils/coarse_matching.py:111-118, 174-190).  The kernel the headline metric names.
//
//   sim[i, j] = <f0[i], f1[j]> / (C * T)          conf = softmax_rows(sim) * softmax_cols(sim)
//
// Two sweeps over the L x S problem of every pair; the matrix is never written:
//
//   sweep 1 (STATS)  e[i, j] = 2^(y[i, j] - G)  with y = sim * log2(e) and ONE reference G per pair (Cauchy-Schwarz bound of
//                    y minus 64: e can neither overflow nor - for any realistic feature range - underflow).  Row sums
//                    are plain per-lane adds (lane = row), column sums a packed 31-shuffle butterfly over the warp's 32
//                    rows: one MUFU.EX2 per element, no running maxima, no rescaling, no shared memory.
//   merge            R[i] = sum_j e, Cs[j] = sum_i e; log-sum-exps Lr = G + log2 R, Lc = G + log2 Cs.  A row / column
//                    whose sum left the safe range [2^-80, 2^100] raises a flag and the caller re-runs the exact
//                    online-max sweeps of umma_gemm.cu (never observed on real features; tested by biasing G).
//   sweep 2 (CONF)   recompute the tile; log2 conf = 2 y - Lr[i] - Lc[j], so one FFMA and a running maximum per
//                    element find the rows of a block that can hold a match at all (about one block in five); only there
//                    the candidates are evaluated with the reference's formula (full-precision exp2f, IEEE division)
//                    and update rowbest / colbest exactly like the first-generation sweeps.
//
// Pipeline (one persistent CTA per SM, 384 threads; CTA pairs by default):
//   * work unit = (pair, 128-row tile of f0 [pairs: two of them], a range of 128-column tiles of f1).  The f0 tile - all
//     of K, both planes, 128 KB - stays RESIDENT in shared memory for the whole unit; only f1 streams through the ring.
//     CTA pairs run tcgen05.mma.cta_group::2 (M = 256): each CTA stages HALF of every f1 tile, so the L2 -> SM traffic
//     per 128 x 128 tile drops from 256 KB (first generation) to 64 KB - below the ~50 B/clk/SM the TMA path delivers
//     (tools/probe_tma.py) even at the full MMA rate of one tile per 3072 cycles.
//   * the split-fp16 product uses TWO accumulators instead of two passes: X += A_hi*B_lo + A_lo*B_hi and Y += A_hi*B_hi
//     over the whole K = 256 (sim = Y + 2^-8 X, one FFMA in the epilogue), so a ring stage is released the moment its 12
//     MMAs are issued - no stage is held for a second pass.  TMEM: 2 tiles x (X, Y) x 128 columns = 512.
//   * epilogue warps 4..11 = 4 TMEM lane quadrants x 2 tile parities: a warp owns every other tile, so it has two MMA
//     tile times for its 4 blocks of 32 x 32 and row sums / row best stay in registers for the whole unit.
#include <math.h>
#include <stdlib.h>

#include <algorithm>
#include <vector>

#include "corr_sweep.cuh"
#include "umma_ptx.cuh"

namespace gimb {
namespace {

constexpr int CM = 128, CN = 128, CK = 64;
constexpr int C_THREADS = 384;
constexpr int A_PLANE_BYTES = CM * CK * 2;       // one plane of one k-block of the resident f0 tile: 16 KB
constexpr int A_KB_BYTES = 2 * A_PLANE_BYTES;    // hi then lo
constexpr int MAX_KB = 4;                        // C <= 256
constexpr int C_SMEM_LIMIT = 227 * 1024;
constexpr float kRefMargin = 32.f;               // G = bound - 32 (log2 units): e <= 2^32, sums <= 2^32 * S
constexpr float kSumLo = 8.2718061e-25f;         // 2^-80: below this the flushed terms could matter -> exact fallback
constexpr float kSumHi = 1.2676506e30f;          // 2^100

enum { SWEEP_STATS = 0, SWEEP_CONF = 1 };

struct CorrMaps {
  CUtensorMap a_hi, a_lo, b_hi, b_lo;
};

struct CorrParams {
  int nb, L, S, num_kb;
  int m_tiles, n_tiles, m_groups;  // m_groups = m tiles (single CTAs) or pairs of m tiles (CTA pairs)
  int n_split, tiles_per_unit, units;
  int b_stages, b_stage_bytes;
  unsigned idesc;
  float c2;        // log2(e) / (C * T):  y = <f0, f1> * c2
  float g_bias;    // test knob: shifts the reference (GIMB_CORR_GBIAS) to exercise the fallback
  const float* normsq;  // [2][nb] max squared row norm of f0 / f1
  const uint8_t* mask0;
  const uint8_t* mask1;
  int Lp, Sp;      // row / column counts padded to whole tiles (stat arrays are padded, never out of bounds)
  float* rowpart;  // [2 * n_split][nb * Lp]
  float* colpart;  // [4 * m tiles incl. dummy][nb * Sp]
  const float2* rowstat;  // [nb * Lp]  (Lr, R);  (+inf, 1) for masked / padding rows
  const float* colthr;    // [nb * Sp]  Lc + log2(thr) - margin;  +inf for masked / padding columns
  const float* colsum;    // [nb * Sp]  Cs
  unsigned long long* rowbest;
  unsigned int* colbest;
  float* conf_out;
  long long* dbg;  // optional [grid][16] cycle counters (GIMB_CORR_DEBUG=1): where every role waits
};

__device__ __forceinline__ float pair_ref(const CorrParams& p, int img) {
  return __fsqrt_rn(__fmul_rn(p.normsq[img], p.normsq[p.nb + img])) * p.c2 - kRefMargin + p.g_bias;
}
__device__ __forceinline__ float ex2_approx(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ bool elect_one() {  // one lane of the (converged) warp
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ float max3(float a, float b, float c) { return fmaxf(fmaxf(a, b), c); }

// one step of the packed butterfly: v[0 .. N) -> v[0 .. N/2), partner lane ^ (N/2).  After the steps 32, 16, 8, 4, 2
// v[0] of lane l is the sum over the warp's 32 lanes of the original v[l].
template <int N>
__device__ __forceinline__ void bfly_step(float (&v)[32], int lane) {
  const bool up = (lane & (N / 2)) != 0;
#pragma unroll
  for (int k = 0; k < N / 2; ++k) {
    const float keep = up ? v[k + N / 2] : v[k];
    const float send = up ? v[k] : v[k + N / 2];
    v[k] = keep + __shfl_xor_sync(0xffffffffu, send, N / 2);
  }
}

struct Unit {
  int img, mg, sp, n_begin, n_end;
};
__device__ __forceinline__ Unit decode_unit(const CorrParams& p, int u) {
  Unit r;
  const int per_img = p.m_groups * p.n_split;
  r.img = u / per_img;
  const int rem = u - r.img * per_img;
  r.mg = rem / p.n_split;
  r.sp = rem - r.mg * p.n_split;
  r.n_begin = r.sp * p.tiles_per_unit;
  r.n_end = min(p.n_tiles, r.n_begin + p.tiles_per_unit);
  return r;
}

// exact confidence of one candidate (coarse_matching.py:118): softmax over dim 1 times softmax over dim 2, both from the
// same numerator 2^(y - G)
__device__ __forceinline__ float conf_exact(const CorrParams& p, float t, float G, float R, float cs) {
  const float a = exp2f(fmaf(t, p.c2, -G));
  return __fdiv_rn(a, cs) * __fdiv_rn(a, R);
}
__device__ __noinline__ unsigned long long conf_candidate(const CorrParams& p, float t, float G, float R, int img, int col) {
  const float cs = __ldg(p.colsum + (size_t)img * p.Sp + col);
  const float conf = conf_exact(p, t, G, R, cs);
  const unsigned int bits = __float_as_uint(conf);
  atomicMax(&p.colbest[(size_t)img * p.S + col], bits);
  return ((unsigned long long)bits << 32) | (unsigned long long)(~(unsigned int)col);
}

template <int SWEEP, bool kPair>
__global__ void __launch_bounds__(C_THREADS, 1) corr_sweep_kernel(const __grid_constant__ CorrMaps maps, const CorrParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - raw);
  const uint32_t a_base = base;
  const uint32_t b_base = base + (uint32_t)p.num_kb * A_KB_BYTES;
  const uint32_t bars = b_base + (uint32_t)p.b_stages * (uint32_t)p.b_stage_bytes;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gen + (bars - base) + 224);
  auto b_full = [&](int s) { return bars + 8u * s; };            // up to 8 stages
  auto b_empty = [&](int s) { return bars + 64u + 8u * s; };
  auto a_full = [&](int kb) { return bars + 128u + 8u * kb; };   // per k-block of the resident f0 tile
  auto a_empty = [&](int kb) { return bars + 160u + 8u * kb; };
  auto tfull = [&](int h) { return bars + 192u + 8u * h; };      // accumulator pair of tile parity h complete
  auto tempty = [&](int h) { return bars + 208u + 8u * h; };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = kPair ? cluster_ctarank() : 0u;
  const int first = kPair ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int step = kPair ? (int)(gridDim.x >> 1) : (int)gridDim.x;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&maps.a_hi);
    tma_prefetch_desc(&maps.a_lo);
    tma_prefetch_desc(&maps.b_hi);
    tma_prefetch_desc(&maps.b_lo);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < 8; ++s) { mbar_init(b_full(s), 1); mbar_init(b_empty(s), 1); }
      for (int k = 0; k < MAX_KB; ++k) { mbar_init(a_full(k), 1); mbar_init(a_empty(k), 1); }
      for (int h = 0; h < 2; ++h) {
        mbar_init(tfull(h), 1);
        mbar_init(tempty(h), kPair ? 8u : 4u);  // the 4 quadrant warps of the parity (of both CTAs for a pair)
      }
      fence_barrier_init();
    }
    __syncwarp();
    if constexpr (kPair) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  if (kPair) cluster_sync_all();
  tc_fence_after();
  // A 512-column allocation is the whole tensor memory: its base address is 0 by construction, and using the constant
  // keeps the MMA operands in uniform registers (the slot written by tcgen05.alloc is only read back in debug runs).
  constexpr uint32_t tmem_base = 0u;
  if (p.dbg != nullptr && *tmem_slot != 0u) __trap();
  long long dbg_acc[6] = {0, 0, 0, 0, 0, 0};  // [3]: MMA warp only - b_full waits on the first tile of a unit
  const long long dbg_t0 = clock64();
  auto timed_wait = [&](uint32_t bar, uint32_t parity, int slot) {
    if (p.dbg == nullptr) { mbar_wait(bar, parity); return; }
    const long long t0 = clock64();
    mbar_wait(bar, parity);
    dbg_acc[slot] += clock64() - t0;
  };

  if (warp == 0) {
    // =============================================================== TMA producer
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(40));
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0, ucount = 0;
      const uint32_t b_plane = (uint32_t)p.b_stage_bytes >> 1;
      for (int u = first; u < p.units; u += step, ++ucount) {
        const Unit un = decode_unit(p, u);
        const int m_tile = kPair ? un.mg * 2 + (int)rank : un.mg;
        for (int nt = un.n_begin; nt < un.n_end; ++nt) {
          for (int kb = 0; kb < p.num_kb; ++kb) {
            timed_wait(b_empty(stage), phase ^ 1u, 1);
            const uint32_t sB = b_base + (uint32_t)stage * (uint32_t)p.b_stage_bytes;
            if constexpr (kPair) {
              const uint32_t fb = mapa_cta(b_full(stage), 0u);
              if (rank == 0) mbar_expect_tx(b_full(stage), 2u * (uint32_t)p.b_stage_bytes);
              const int n0 = nt * CN + (int)rank * (CN / 2);  // this CTA's half of the f1 tile
              tma_load_3d_pair(sB, &maps.b_hi, fb, kb * CK, n0, un.img);
              tma_load_3d_pair(sB + b_plane, &maps.b_lo, fb, kb * CK, n0, un.img);
            } else {
              mbar_expect_tx(b_full(stage), (uint32_t)p.b_stage_bytes);
              tma_load_3d(sB, &maps.b_hi, b_full(stage), kb * CK, nt * CN, un.img);
              tma_load_3d(sB + b_plane, &maps.b_lo, b_full(stage), kb * CK, nt * CN, un.img);
            }
            if (++stage == p.b_stages) { stage = 0; phase ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // =============================================================== MMA issuer (pairs: the leader CTA issues for both)
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(40));
    // The whole warp runs the loop (warp-uniform control flow, every operand derived from kernel parameters and the
    // shared-memory base): ptxas then keeps descriptors and addresses in UNIFORM registers and an MMA costs one UIADD
    // plus the UTCHMMA itself.  Under `if (lane == 0)` every operand lived in a vector register and each MMA paid a
    // 15-instruction elect / R2UR "waterfall" (~100 cycles - more than the 64 cycles an N = 128 MMA takes; measured
    // with tools/probe_mma.py and the per-role cycle counters of GIMB_CORR_DEBUG).
    if (!(kPair && rank != 0)) {
      int stage = 0;
      uint32_t phase = 0, ucount = 0, tcount = 0;
      const bool leader = elect_one();
      const uint64_t dconst = make_desc_sw128(0u);
      const uint32_t b_plane = (uint32_t)p.b_stage_bytes >> 1;
      auto wait1 = [&](uint32_t bar, uint32_t parity, int slot) {  // one lane polls, the warp re-converges
        if (lane == 0) timed_wait(bar, parity, slot);
        __syncwarp();
      };
      for (int u = first; u < p.units; u += step, ++ucount) {
        const Unit un = decode_unit(p, u);
        for (int nt = un.n_begin; nt < un.n_end; ++nt, ++tcount) {
          const uint32_t h = tcount & 1u, k = tcount >> 1;
          wait1(tempty((int)h), (k & 1u) ^ 1u, 0);
          tc_fence_after();
          const uint32_t tX = h * 256u, tY = tX + 128u;  // the 512-column allocation starts at TMEM address 0 (checked below)
          for (int kb = 0; kb < p.num_kb; ++kb) {
            if (nt == un.n_begin) wait1(a_full(kb), ucount & 1u, 1);
            wait1(b_full(stage), phase, nt == un.n_begin ? 3 : 2);
            tc_fence_after();
            const uint64_t dA_hi = dconst + (((a_base + (uint32_t)kb * A_KB_BYTES) & 0x3FFFFu) >> 4);
            const uint64_t dA_lo = dA_hi + (A_PLANE_BYTES >> 4);
            const uint64_t dB_hi = dconst + (((b_base + (uint32_t)stage * (uint32_t)p.b_stage_bytes) & 0x3FFFFu) >> 4);
            const uint64_t dB_lo = dB_hi + (b_plane >> 4);
            const uint32_t acc0 = kb == 0 ? 0u : 1u;
            if (leader) {
#pragma unroll
              for (int kk = 0; kk < CK / 16; ++kk) {  // 16 fp16 = 32 bytes = 2 address units inside the swizzled row
                const uint32_t acc = kk == 0 ? acc0 : 1u;
                if constexpr (kPair) {
                  umma_f16_pair(tX, dA_hi + 2 * kk, dB_lo + 2 * kk, p.idesc, acc);
                  umma_f16_pair(tX, dA_lo + 2 * kk, dB_hi + 2 * kk, p.idesc, 1u);
                  umma_f16_pair(tY, dA_hi + 2 * kk, dB_hi + 2 * kk, p.idesc, acc);
                } else {
                  umma_f16(tX, dA_hi + 2 * kk, dB_lo + 2 * kk, p.idesc, acc);
                  umma_f16(tX, dA_lo + 2 * kk, dB_hi + 2 * kk, p.idesc, 1u);
                  umma_f16(tY, dA_hi + 2 * kk, dB_hi + 2 * kk, p.idesc, acc);
                }
              }
              // the stage is free once these MMAs have read it; last tile of the unit: so is the f0 k-block
              if constexpr (kPair) {
                umma_commit_pair(b_empty(stage));
                if (nt == un.n_end - 1) umma_commit_pair(a_empty(kb));
                if (kb == p.num_kb - 1) umma_commit_pair(tfull((int)h));
              } else {
                umma_commit(b_empty(stage));
                if (nt == un.n_end - 1) umma_commit(a_empty(kb));
                if (kb == p.num_kb - 1) umma_commit(tfull((int)h));
              }
            }
            __syncwarp();
            if (++stage == p.b_stages) { stage = 0; phase ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 2) {
    // =============================================================== TMA producer of the resident f0 tile
    // A separate thread: the f1 ring of warp 0 keeps prefetching across unit boundaries while this one waits for the
    // MMAs of the previous unit's last tile to release the k-blocks of the f0 tile.
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(40));
    if (lane == 0) {
      uint32_t ucount = 0;
      for (int u = first; u < p.units; u += step, ++ucount) {
        const Unit un = decode_unit(p, u);
        const int m_tile = kPair ? un.mg * 2 + (int)rank : un.mg;
        for (int kb = 0; kb < p.num_kb; ++kb) {
          timed_wait(a_empty(kb), (ucount & 1u) ^ 1u, 0);
          const uint32_t dst = a_base + (uint32_t)kb * A_KB_BYTES;
          if constexpr (kPair) {
            const uint32_t fb = mapa_cta(a_full(kb), 0u);  // both CTAs' bytes complete on the leader's barrier
            if (rank == 0) mbar_expect_tx(a_full(kb), 2u * A_KB_BYTES);
            tma_load_3d_pair(dst, &maps.a_hi, fb, kb * CK, m_tile * CM, un.img);
            tma_load_3d_pair(dst + A_PLANE_BYTES, &maps.a_lo, fb, kb * CK, m_tile * CM, un.img);
          } else {
            mbar_expect_tx(a_full(kb), A_KB_BYTES);
            tma_load_3d(dst, &maps.a_hi, a_full(kb), kb * CK, m_tile * CM, un.img);
            tma_load_3d(dst + A_PLANE_BYTES, &maps.a_lo, a_full(kb), kb * CK, m_tile * CM, un.img);
          }
        }
      }
    }
  } else if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(40));
  } else {
    // =============================================================== epilogue warps: quadrant q, tile parity h
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(232));
    const int q = warp & 3, h = (warp - 4) >> 2;
    const uint32_t tbase = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)h * 256u;
    uint32_t tcount = 0;
    float* thr_strip = reinterpret_cast<float*>(gen + (bars - base) + 256) + (warp - 4) * CN;  // CONF: 128 thresholds per warp
    auto release_tmem = [&]() {
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (kPair && rank != 0) mbar_arrive_cluster(mapa_cta(tempty(h), 0u));
        else mbar_arrive(tempty(h));
      }
    };
    for (int u = first; u < p.units; u += step) {
      const Unit un = decode_unit(p, u);
      const int m_tile = kPair ? un.mg * 2 + (int)rank : un.mg;
      const int r_in = m_tile * CM + q * 32 + lane;      // row inside the pair's problem
      const bool row_ok = r_in < p.L;
      const size_t grow = (size_t)un.img * p.L + r_in;
      const bool rv = row_ok && (p.mask0 == nullptr || p.mask0[grow] != 0);
      const float G = pair_ref(p, un.img);
      const float negG = -G;
      // ---- per-unit state
      float rs0 = 0.f, rs1 = 0.f, rs2 = 0.f, rs3 = 0.f;  // STATS: this row's sum over the unit's tiles of this parity
      float Lr = INFINITY, R = 1.f;                      // CONF
      unsigned long long best = 0ull;
      if (SWEEP == SWEEP_CONF && m_tile < p.m_tiles) {
        const float2 st = p.rowstat[(size_t)un.img * p.Lp + r_in];
        Lr = st.x; R = st.y;
      }
      for (int nt = un.n_begin; nt < un.n_end; ++nt, ++tcount) {
        if ((int)(tcount & 1u) != h) continue;
        const bool edge = (m_tile + 1) * CM > p.L || (nt + 1) * CN > p.S || p.mask0 != nullptr || p.mask1 != nullptr;
        const size_t cbase = (size_t)un.img * p.Sp + (size_t)nt * CN;  // padded column index of the tile's first column

        // CONF: this tile's 128 column thresholds, one coalesced load per warp (issued before the wait on the
        // accumulators), parked in the warp's shared-memory strip and read back as broadcasts
        float4 thr4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if constexpr (SWEEP == SWEEP_CONF) thr4 = __ldg(reinterpret_cast<const float4*>(p.colthr + cbase) + lane);
        timed_wait(tfull(h), (tcount >> 1) & 1u, 0);
        tc_fence_after();
        const long long dbg_t1 = p.dbg ? clock64() : 0;

        // ---- phase 1: drain.  sim (raw dot product) = Y + 2^-8 X for the tile's 128 columns of this lane's row; the
        // accumulators go back to the MMA warp before any of the statistics math starts.
        float t[4][32];
        {
          uint32_t xa[32], ya[32], xb[32], yb[32];
          auto fold = [&](const uint32_t (&xr)[32], const uint32_t (&yr)[32], float (&dst)[32]) {
#pragma unroll
            for (int j = 0; j < 32; ++j) dst[j] = fmaf(__uint_as_float(xr[j]), 1.f / kSplitScale, __uint_as_float(yr[j]));
          };
          tmem_ld32(tbase + 0, xa);
          tmem_ld32(tbase + 128 + 0, ya);
          tmem_ld_wait();
          tmem_ld32(tbase + 32, xb);
          tmem_ld32(tbase + 128 + 32, yb);
          fold(xa, ya, t[0]);
          tmem_ld_wait();
          tmem_ld32(tbase + 64, xa);
          tmem_ld32(tbase + 128 + 64, ya);
          fold(xb, yb, t[1]);
          tmem_ld_wait();
          tmem_ld32(tbase + 96, xb);
          tmem_ld32(tbase + 128 + 96, yb);
          fold(xa, ya, t[2]);
          tmem_ld_wait();
          release_tmem();
          if (p.dbg) dbg_acc[1] += clock64() - dbg_t1;
          fold(xb, yb, t[3]);
        }
        if constexpr (SWEEP == SWEEP_CONF) {
          __syncwarp();  // the previous tile's broadcasts are done
          *reinterpret_cast<float4*>(thr_strip + 4 * lane) = thr4;
          __syncwarp();
        }

        // ---- phase 2: per 32-column block, from registers
        auto block = [&](const float (&tv)[32], int b) {
          if constexpr (SWEEP == SWEEP_STATS) {
            float e[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) e[j] = ex2_approx(fmaf(tv[j], p.c2, negG));
            if (edge) {  // warp-uniform: tail tiles and masked problems only
              const int col = nt * CN + b * 32 + lane;
              const bool cv = col < p.S && (p.mask1 == nullptr || p.mask1[(size_t)un.img * p.S + col] != 0);
              const unsigned cbits = __ballot_sync(0xffffffffu, cv);
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (!rv || !((cbits >> j) & 1u)) e[j] = 0.f;
            }
#pragma unroll
            for (int j = 0; j < 32; j += 4) { rs0 += e[j]; rs1 += e[j + 1]; rs2 += e[j + 2]; rs3 += e[j + 3]; }
            bfly_step<32>(e, lane);
            bfly_step<16>(e, lane);
            bfly_step<8>(e, lane);
            bfly_step<4>(e, lane);
            bfly_step<2>(e, lane);
            p.colpart[(size_t)(m_tile * 4 + q) * ((size_t)p.nb * p.Sp) + cbase + b * 32 + lane] = e[0];
          } else {
            float z[32];
            const float c22 = 2.f * p.c2;
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4) {
              const float4 v = *reinterpret_cast<const float4*>(thr_strip + b * 32 + j4 * 4);  // broadcast
              z[j4 * 4 + 0] = fmaf(tv[j4 * 4 + 0], c22, -v.x);   // log2 conf + Lr - log2(thr) + margin
              z[j4 * 4 + 1] = fmaf(tv[j4 * 4 + 1], c22, -v.y);
              z[j4 * 4 + 2] = fmaf(tv[j4 * 4 + 2], c22, -v.z);
              z[j4 * 4 + 3] = fmaf(tv[j4 * 4 + 3], c22, -v.w);
            }
            float m0 = max3(z[0], z[1], z[2]), m1 = max3(z[3], z[4], z[5]);
#pragma unroll
            for (int j = 6; j < 30; j += 4) { m0 = max3(m0, z[j], z[j + 1]); m1 = max3(m1, z[j + 2], z[j + 3]); }
            m0 = max3(m0, z[30], z[31]);
            const bool cand = fmaxf(m0, m1) > Lr;
            if (__any_sync(0xffffffffu, cand)) {
              if (cand) {
#pragma unroll
                for (int j = 0; j < 32; ++j)
                  if (z[j] > Lr) {
                    const unsigned long long pk = conf_candidate(p, tv[j], G, R, un.img, nt * CN + b * 32 + j);
                    best = pk > best ? pk : best;
                  }
              }
            }
            if (p.conf_out != nullptr && row_ok) {  // debug tap (tests): every entry, same formula as the candidates
              float* dst = p.conf_out + grow * (size_t)p.S + (size_t)nt * CN + b * 32;
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const int col = nt * CN + b * 32 + j;
                if (col < p.S) {
                  const bool live = Lr < INFINITY && z[j] > -INFINITY;  // masked rows / columns: exactly 0
                  dst[j] = live ? conf_exact(p, tv[j], G, R, __ldg(p.colsum + cbase + b * 32 + j)) : 0.f;
                }
              }
            }
          }
        };
        block(t[0], 0);
        block(t[1], 1);
        block(t[2], 2);
        block(t[3], 3);
        if (p.dbg) dbg_acc[2] += clock64() - dbg_t1;
      }
      // ---- unit end
      if constexpr (SWEEP == SWEEP_STATS) {
        if (m_tile < p.m_tiles)
          p.rowpart[(size_t)(un.sp * 2 + h) * ((size_t)p.nb * p.Lp) + (size_t)un.img * p.Lp + r_in] = (rs0 + rs1) + (rs2 + rs3);
      } else {
        if (row_ok && best) atomicMax(&p.rowbest[grow], best);
      }
    }
  }

  if (p.dbg != nullptr && lane == 0 && (warp == 0 || warp == 1 || warp == 4 || warp == 8)) {
    // [0..3] producer: a_empty, b_empty, -, total | [4..7] MMA: tempty, a_full, b_full, total | [8..11] / [12..15]
    // epilogue warp 4 / 8: tfull wait, drain, tile total, kernel total
    const int o = warp == 0 ? 0 : (warp == 1 ? 4 : (warp == 4 ? 8 : 12));
    long long* d = p.dbg + (size_t)blockIdx.x * 16 + o;
    d[0] = dbg_acc[0]; d[1] = dbg_acc[1]; d[2] = dbg_acc[2]; d[3] = warp == 1 ? dbg_acc[3] : clock64() - dbg_t0;
  }
  // ---- teardown
  tc_fence_before();
  __syncthreads();
  if (kPair) cluster_sync_all();
  if (warp == 1) {
    if constexpr (kPair) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
  }
}

// ------------------------------------------------------------------------------------------- small kernels
// max squared row norm of f0 / f1 per pair (the Cauchy-Schwarz bound behind the reference G).  One CTA per 256 rows of
// one image, one atomicMax per CTA (non-negative floats order like their bit patterns).
__global__ void __launch_bounds__(256) corr_norm_kernel(const float* __restrict__ f0, const float* __restrict__ f1, int nb, int L,
                                                        int S, int C, int blocks0, int blocks1, unsigned int* __restrict__ normsq) {
  __shared__ float wmax[8];
  int b = blockIdx.x;
  const int which = b >= nb * blocks0;
  if (which) b -= nb * blocks0;
  const int per = which ? blocks1 : blocks0, rows = which ? S : L;
  const int img = b / per, r0 = (b - img * per) * 256;
  const float* f = (which ? f1 : f0) + (size_t)img * rows * C;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float best = 0.f;
  for (int r = r0 + warp; r < min(rows, r0 + 256); r += 8) {
    const float4* row = reinterpret_cast<const float4*>(f + (size_t)r * C);
    float s = 0.f;
    for (int c4 = lane; c4 < C / 4; c4 += 32) {
      const float4 v = __ldg(row + c4);
      s = fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, fmaf(v.w, v.w, s))));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    best = fmaxf(best, s);
  }
  if (lane == 0) wmax[warp] = best;
  __syncthreads();
  if (threadIdx.x == 0) {
    float m = wmax[0];
    for (int w = 1; w < 8; ++w) m = fmaxf(m, wmax[w]);
    atomicMax(&normsq[which * nb + img], __float_as_uint(m * 1.0001f));  // the 1e-4 covers the rounding of the sums
  }
}

// partial sums -> row statistics (Lr, R) / column statistics (threshold, Cs); raises *flag when a live sum left the
// range in which flushed terms cannot matter
__global__ void __launch_bounds__(256) corr_merge_kernel(CorrParams p, int cols, int parts, float thr_log2, float2* __restrict__ rowstat,
                                                         float* __restrict__ colthr, float* __restrict__ colsum, int* __restrict__ flag) {
  const int n = cols ? p.Sp : p.Lp, real = cols ? p.S : p.L;
  const size_t total = (size_t)p.nb * n;
  const size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int img = (int)(idx / n), i = (int)(idx - (size_t)img * n);
  const float* part = cols ? p.colpart : p.rowpart;
  const uint8_t* mask = cols ? p.mask1 : p.mask0;
  // row partials come in (unit, tile parity) pairs whose two slots may swap with the CTA's tile count: add the two
  // slots of a pair first (commutative), then the pairs in order - the result does not depend on the batch composition
  float s = 0.f;
  for (int t = 0; t < parts; t += 2) s += part[(size_t)t * total + idx] + part[(size_t)(t + 1) * total + idx];
  const bool live = i < real && (mask == nullptr || mask[(size_t)img * real + i] != 0);
  const float G = pair_ref(p, img);
  if (live && !(s >= kSumLo && s <= kSumHi)) *flag = 1;
  const float lse = G + log2f(s);
  if (cols) {
    colthr[idx] = live ? lse + thr_log2 : INFINITY;
    colsum[idx] = live ? s : 1.f;
  } else {
    rowstat[idx] = live ? make_float2(lse, s) : make_float2(INFINITY, 1.f);
  }
}

int corr_pairs_setting() {  // CTA pairs (tcgen05.mma.cta_group::2) unless GIMB_CORR_PAIR=0
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("GIMB_CORR_PAIR");
    v = (e && atoi(e) == 0) ? 0 : 1;
  }
  return v;
}

struct Plan {
  bool pair;
  int m_tiles, n_tiles, m_groups, m_tiles_padded, n_split, tiles_per_unit, units, b_stages, b_stage_bytes, smem;
};

Plan make_plan(const Ctx& ctx, int nb, int L, int S, int C) {
  Plan pl;
  pl.pair = corr_pairs_setting() != 0 && ctx.sm_count >= 2;
  pl.m_tiles = cdiv(L, CM);
  pl.n_tiles = cdiv(S, CN);
  pl.m_groups = pl.pair ? cdiv(pl.m_tiles, 2) : pl.m_tiles;
  pl.m_tiles_padded = pl.pair ? pl.m_groups * 2 : pl.m_tiles;
  const int clusters = pl.pair ? std::max(1, ctx.sm_count / 2) : ctx.sm_count;
  // split of the column range of an (image, row group) into units.  The split must NOT depend on the batch size (the
  // row sums are added unit by unit: a pair must give bit-identical results alone or inside a batch), so it is chosen
  // per problem shape: enough units per pair to occupy every cluster even for a single pair, but at least 8 tiles
  // per unit so that reloading the resident f0 tile stays a small fraction.
  {
    int s = std::max(1, cdiv(clusters, pl.m_groups));
    s = std::min(s, std::max(1, pl.n_tiles / 8));
    pl.tiles_per_unit = cdiv(pl.n_tiles, s);
    pl.n_split = cdiv(pl.n_tiles, pl.tiles_per_unit);
  }
  pl.units = nb * pl.m_groups * pl.n_split;
  pl.b_stage_bytes = (pl.pair ? CN / 2 : CN) * CK * 2 * 2;
  const int num_kb = C / CK;
  const int fixed = num_kb * A_KB_BYTES + 256 /*barriers*/ + 8 * CN * 4 /*threshold strips*/ + 1024 /*alignment*/;
  // three stages: measured faster than 4 or 5 (835 vs 868 us per 32 pairs for the stats sweep, tensor pipe 76.7 vs 70.3 %);
  // the MMA warp never waits longer for f1 with the shorter ring, and fewer bulk copies are in flight beside the operand reads
  pl.b_stages = std::min(3, (C_SMEM_LIMIT - fixed) / pl.b_stage_bytes);
  {
    static int forced = -1;  // measurement knob
    if (forced < 0) { const char* e = getenv("GIMB_CORR_STAGES"); forced = e ? atoi(e) : 0; }
    if (forced > 0) pl.b_stages = std::min((C_SMEM_LIMIT - fixed) / pl.b_stage_bytes, forced);
  }
  pl.smem = fixed + pl.b_stages * pl.b_stage_bytes;
  return pl;
}

}  // namespace

bool corr_sweep_supported(int C) { return C % CK == 0 && C / CK <= MAX_KB && C > 0; }

void corr_sweep_parts(const Ctx& ctx, int nb, int L, int S, int C, int* row_parts, int* col_parts, int* Lp, int* Sp) {
  const Plan pl = make_plan(ctx, nb, L, S, C);
  *row_parts = 2 * pl.n_split;
  *col_parts = 4 * pl.m_tiles_padded;
  *Lp = pl.m_tiles_padded * CM;
  *Sp = pl.n_tiles * CN;
}

int corr_sweeps(Ctx& ctx, const CorrSweep& c) {
  GIMB_CHECK(corr_sweep_supported(c.C), "corr_sweeps: C = %d not supported", c.C);
  GIMB_CHECK(c.f0.hi && c.f0.lo && c.f1.hi && c.f1.lo && c.f0_f32 && c.f1_f32, "corr_sweeps: operands missing");
  GIMB_CHECK(c.C % 4 == 0 && c.f0.ld == c.C && c.f1.ld == c.C, "corr_sweeps: dense [rows, C] planes expected");
  if (ctx.dry || c.N == 0) return 0;
  const Plan pl = make_plan(ctx, c.N, c.L, c.S, c.C);
  CorrParams p = {};
  CorrMaps maps;
  memset(&maps, 0, sizeof(maps));
  p.nb = c.N; p.L = c.L; p.S = c.S; p.num_kb = c.C / CK;
  p.m_tiles = pl.m_tiles; p.n_tiles = pl.n_tiles; p.m_groups = pl.m_groups;
  p.n_split = pl.n_split; p.tiles_per_unit = pl.tiles_per_unit; p.units = pl.units;
  p.b_stages = pl.b_stages; p.b_stage_bytes = pl.b_stage_bytes;
  p.idesc = (1u << 4) | ((unsigned)(CN >> 3) << 17) | ((unsigned)((pl.pair ? 2 * CM : CM) >> 4) << 24);
  p.c2 = (float)(1.4426950408889634 / ((double)c.C * (double)c.temperature));
  {
    static float gbias = NAN;
    if (gbias != gbias) {
      const char* e = getenv("GIMB_CORR_GBIAS");
      gbias = e ? (float)atof(e) : 0.f;
    }
    p.g_bias = gbias;
  }
  p.normsq = c.normsq;
  p.mask0 = c.mask0; p.mask1 = c.mask1;
  p.Lp = pl.m_tiles_padded * CM; p.Sp = pl.n_tiles * CN;
  p.rowpart = c.rowpart; p.colpart = c.colpart;
  p.rowstat = c.rowstat; p.colthr = c.colthr; p.colsum = c.colsum;
  p.rowbest = c.rowbest; p.colbest = c.colbest; p.conf_out = c.conf_matrix;
  GIMB_TRY(rows_map(&maps.a_hi, c.f0.hi, c.C, c.L, c.f0.ld, CM, c.N, CK));
  GIMB_TRY(rows_map(&maps.a_lo, c.f0.lo, c.C, c.L, c.f0.ld, CM, c.N, CK));
  GIMB_TRY(rows_map(&maps.b_hi, c.f1.hi, c.C, c.S, c.f1.ld, pl.pair ? CN / 2 : CN, c.N, CK));
  GIMB_TRY(rows_map(&maps.b_lo, c.f1.lo, c.C, c.S, c.f1.ld, pl.pair ? CN / 2 : CN, c.N, CK));

  // ---- reference G: max row norms
  GIMB_CUDA(cudaMemsetAsync(c.normsq, 0, 2 * (size_t)c.N * sizeof(float), ctx.stream));
  GIMB_CUDA(cudaMemsetAsync(c.flag, 0, sizeof(int), ctx.stream));
  const int blocks0 = cdiv(c.L, 256), blocks1 = cdiv(c.S, 256);
  corr_norm_kernel<<<c.N * (blocks0 + blocks1), 256, 0, ctx.stream>>>(c.f0_f32, c.f1_f32, c.N, c.L, c.S, c.C, blocks0, blocks1,
                                                                      reinterpret_cast<unsigned int*>(c.normsq));
  GIMB_LAUNCH_CHECK();

  const int clusters = pl.pair ? ctx.sm_count / 2 : ctx.sm_count;
  const int grid = std::min(pl.units, clusters) * (pl.pair ? 2 : 1);
  cudaLaunchConfig_t lcfg = {};
  cudaLaunchAttribute lattr[1];
  lcfg.gridDim = dim3(grid); lcfg.blockDim = dim3(C_THREADS); lcfg.dynamicSmemBytes = pl.smem; lcfg.stream = ctx.stream;
  lattr[0].id = cudaLaunchAttributeClusterDimension;
  lattr[0].val.clusterDim.x = pl.pair ? 2 : 1; lattr[0].val.clusterDim.y = 1; lattr[0].val.clusterDim.z = 1;
  lcfg.attrs = lattr; lcfg.numAttrs = 1;
  GIMB_SMEM_OPTIN((corr_sweep_kernel<SWEEP_STATS, false>), C_SMEM_LIMIT);
  GIMB_SMEM_OPTIN((corr_sweep_kernel<SWEEP_CONF, false>), C_SMEM_LIMIT);
  GIMB_SMEM_OPTIN((corr_sweep_kernel<SWEEP_STATS, true>), C_SMEM_LIMIT);
  GIMB_SMEM_OPTIN((corr_sweep_kernel<SWEEP_CONF, true>), C_SMEM_LIMIT);

  static int dbg_on = -1;
  if (dbg_on < 0) { const char* e = getenv("GIMB_CORR_DEBUG"); dbg_on = (e && atoi(e) != 0) ? 1 : 0; }
  long long* dbg = nullptr;
  if (dbg_on) {
    GIMB_CUDA(cudaMalloc(&dbg, (size_t)grid * 16 * sizeof(long long)));
    GIMB_CUDA(cudaMemsetAsync(dbg, 0, (size_t)grid * 16 * sizeof(long long), ctx.stream));
    p.dbg = dbg;
  }
  auto dump_dbg = [&](const char* what) {
    if (!dbg) return;
    std::vector<long long> h((size_t)grid * 16);
    cudaStreamSynchronize(ctx.stream);
    cudaMemcpy(h.data(), dbg, h.size() * sizeof(long long), cudaMemcpyDeviceToHost);
    double avg[16] = {0};
    for (int g = 0; g < grid; ++g) for (int k = 0; k < 16; ++k) avg[k] += (double)h[(size_t)g * 16 + k] / grid;
    fprintf(stderr, "[corr dbg %s] units %d tiles/unit %d grid %d | producer a_empty %.0f b_empty %.0f total %.0f | mma tempty %.0f a_full %.0f "
            "b_full %.0f (+ first tile of a unit %.0f) | epi4 tfull %.0f drain %.0f tiles %.0f total %.0f | epi8 tfull %.0f drain %.0f tiles %.0f total %.0f\n",
            what, pl.units, pl.tiles_per_unit, grid, avg[0], avg[1], avg[3], avg[4], avg[5], avg[6], avg[7], avg[8], avg[9], avg[10], avg[11],
            avg[12], avg[13], avg[14], avg[15]);
    cudaMemset(dbg, 0, h.size() * sizeof(long long));
  };
  // ---- sweep 1
  if (pl.pair) GIMB_CUDA(cudaLaunchKernelEx(&lcfg, corr_sweep_kernel<SWEEP_STATS, true>, maps, p));
  else GIMB_CUDA(cudaLaunchKernelEx(&lcfg, corr_sweep_kernel<SWEEP_STATS, false>, maps, p));
  GIMB_LAUNCH_CHECK();
  ctx.mark("corr_stats");
  dump_dbg("stats");
  // ---- merge
  const float thr_log2 = c.thr > 0.f ? log2f(c.thr) - 1e-3f : -INFINITY;
  const size_t nrow = (size_t)c.N * p.Lp, ncol = (size_t)c.N * p.Sp;
  corr_merge_kernel<<<(unsigned)cdiv64((int64_t)nrow, 256), 256, 0, ctx.stream>>>(p, 0, 2 * pl.n_split, thr_log2, c.rowstat, c.colthr,
                                                                                  c.colsum, c.flag);
  GIMB_LAUNCH_CHECK();
  corr_merge_kernel<<<(unsigned)cdiv64((int64_t)ncol, 256), 256, 0, ctx.stream>>>(p, 1, 4 * pl.m_tiles_padded, thr_log2, c.rowstat,
                                                                                  c.colthr, c.colsum, c.flag);
  GIMB_LAUNCH_CHECK();
  GIMB_CUDA(cudaMemsetAsync(c.rowbest, 0, (size_t)c.N * c.L * sizeof(unsigned long long), ctx.stream));
  GIMB_CUDA(cudaMemsetAsync(c.colbest, 0, (size_t)c.N * c.S * sizeof(unsigned int), ctx.stream));
  ctx.mark("corr_merge");
  // ---- sweep 2
  if (pl.pair) GIMB_CUDA(cudaLaunchKernelEx(&lcfg, corr_sweep_kernel<SWEEP_CONF, true>, maps, p));
  else GIMB_CUDA(cudaLaunchKernelEx(&lcfg, corr_sweep_kernel<SWEEP_CONF, false>, maps, p));
  GIMB_LAUNCH_CHECK();
  ctx.mark("corr_conf");
  dump_dbg("conf");
  if (dbg) cudaFree(dbg);
  ctx.launches += 5;
  return 0;
}

}  // namespace gimb
