// umma_gemm.cu - tcgen05 (5th-gen tensor core) GEMM / implicit-GEMM convolution with split-fp16 operands.
//
// One persistent CTA per SM, warp-specialised (see /opt/skills/guides/blackwell_cuda_programming.md "Anatomy"):
//   warp 0      TMA producer  : cp.async.bulk.tensor (3-D row tiles or 4-D NHWC patches, OOB zero fill = conv padding)
//                               into a ring of shared-memory stages (K = 32: SWIZZLE_64B, K = 64: SWIZZLE_128B),
//                               mbarrier complete_tx; big layers as CTA pairs (B multicast, or cta_group::2 opt-in)
//   warp 1      MMA issuer    : one elected lane issues 3 x tcgen05.mma.kind::f16 (M=128, N<=256, K=16) per k16 step
//                               [cross terms A_hi*B_lo + A_lo*B_hi, then A_hi*B_hi with scale-input-d 2^-8] into a
//                               ring of fp32 TMEM accumulator buffers (2 x 256 or 4 x 128 columns),
//                               tcgen05.commit frees smem stages / publishes the accumulator
//   warps 4..11 epilogue      : the tensor core accumulates only a chunk of K at a time (64, or the whole K <= 128); the
//                               chunks are summed in fp32 registers with round-to-nearest (tcgen05.ld, 32 lanes x 32
//                               columns).  The MMA accumulator truncates, and a long in-TMEM accumulation biases the
//                               result by ~K/16 ulp (measured: 5x the fp32 FMA error at K = 1764); short chunks + RN
//                               adds keep the result in the fp32 class.  Then, row per lane: folded-BN scale/bias,
//                               residual (bulk tensor load one block ahead), activation, row mask, and the block leaves
//                               through the warp's swizzled staging tile with ONE bulk tensor store per tensor (fp32
//                               and/or the re-split fp16 planes of the next layer).  The correlation sweeps
//                               (EPI_CORR_*) replace this by their statistics epilogues.
//
// Tile: BM = 128 output rows (mode 0: consecutive rows; mode 1: an 8 x 16 pixel patch of one image),
//       BN = whole N up to 256 (rounded to 16), stage K = 32 or 64 fp16.
// Environment knobs (measurement only): GIMB_BK=32|64, GIMB_CHUNK_KB, GIMB_CORR_CHUNK_KB, GIMB_CORR_BN=256,
//       GIMB_CLUSTER=1 (no CTA pairs), GIMB_PAIR=mma (tcgen05.mma.cta_group::2 over the pair).
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "ops.cuh"
#include "umma_gemm.cuh"
#include "umma_ptx.cuh"

namespace gimb {
namespace {

constexpr int BM = 128;
constexpr int BK = 32;
constexpr int A_TILE_BYTES = BM * BK * 2;  // 8 KB
constexpr int TH = 8, TW = 16;
constexpr int MAX_STAGES = 8;
constexpr int NUM_EPI_WARPS = 8;   // two warps per TMEM lane quadrant, each owning alternate 32-column groups
constexpr int NUM_THREADS = 128 + NUM_EPI_WARPS * 32;  // warpgroup 0: TMA, MMA, 2 idle warps; warpgroups 1-2: epilogue
constexpr int REGS_LOW = 40, REGS_HIGH = 232;           // setmaxnreg split: 4*32*40 + 8*32*232 = 64512 <= 65536
constexpr int CHUNK_KB_DEFAULT = 2;  // k-blocks (of 32) accumulated inside the tensor core before the fp32 (RN) drain
constexpr int SMEM_LIMIT = 227 * 1024;
constexpr int TBP = 36;  // pitch (floats) of the per-warp 32 x 32 transpose tile: 16-byte aligned rows, conflict-free float4 access
constexpr int SMEM_EXTRA = 1024 /*alignment slack*/ + 512 /*barriers*/ + NUM_EPI_WARPS * 32 * TBP * 4 /*transpose tiles*/ +
                           2 * NUM_EPI_WARPS * 32 * 4 /*LayerNorm row statistics*/;
constexpr int TMEM_COLS = 512;
constexpr int ACC_COLS = 256;  // accumulator buffer width for tiles wider than 128 columns (2 buffers); narrower tiles use 4 x 128

struct TMaps {
  CUtensorMap a_hi[4], a_lo[4];  // mode 1 stride 2: four phase views; otherwise index 0
  CUtensorMap a2_hi, a2_lo;      // mode 0 concat source
  CUtensorMap b_lo, b_hi;
  CUtensorMap bh_lo, bh_hi;      // cluster mode: half-height boxes (bn / 2 rows) of the B planes
  CUtensorMap r_hi, r_lo;        // residual planes: boxes of 32 channels x 32 rows (one epilogue warp's block)
  CUtensorMap r_f32;             // fp32 residual, same blocks (SWIZZLE_128B)
  CUtensorMap o_hi, o_lo, o_f32; // outputs, same blocks: written with bulk tensor stores from the warp's staging tile
};

struct KParams {
  int mode;
  long long M;
  int OH, OW, tiles_h, tiles_w, tiles_per_img;
  int KH, KW, stride, pad;
  int cb1, cb2, K1;  // k-blocks of a / a2 per tap; channels of a
  int ldk;
  int bk;            // K extent of one stage: 32 (64-byte tile rows, SWIZZLE_64B) or 64 (128-byte rows, SWIZZLE_128B)
  int N, n_tiles, bn;
  int stages, stage_bytes;
  int num_tiles, num_kb, num_chunks, chunk_kb;
  int acc_cols, nbuf_log2;  // TMEM ring: 2 x 256 or 4 x 128 columns
  int stg_off;       // TMA epilogue: byte offset (from the 1024-aligned smem base) of the per-warp staging tiles
  int cluster;       // 1, or 2: CTA pairs
  int pair_mma;      // cluster == 2 only.  1: tcgen05.mma.cta_group::2 - one M = 256 MMA over the pair, each CTA holds its
                     // 128 A rows and HALF of the B rows (operand reads and TMA writes of B per SM halve);
                     // 0: two independent M = 128 MMAs, every B tile multicast into both CTAs
  int m_tiles_real;  // cluster mode: m tiles that exist (the pair grid may carry one dummy tile)
  int n_imgs;
  unsigned idesc;
  const float* scale;
  const float* bias;
  const float* residual;
  const uint8_t* row_mask;
  int act0, act1, act_split;
  float div;
  float* out_f32;
  __half* out_hi;
  __half* out_lo;
  __half* out_h8;
  int ldp;
  const __half* res_hi;  // residual as split planes (OUT_RES_PLANES), pitch ldr
  const __half* res_lo;
  int ldr;
  // ---- coarse-matching sweeps (EPI_CORR_*): batch of nb problems, A = f0[b] [L,C], B = f1[b] [S,C]
  int nb, L, S, tiles_per_batch;
  const uint8_t* mask0;
  const uint8_t* mask1;
  float inv_c, temperature, thr_log, sim_scale;
  float2* rowpart;   // [nb*L][row_parts]
  float2* colpart;   // [nb*S][col_parts]
  int row_parts, col_parts;
  const float2* rowstat;
  const float2* colstat;
  unsigned long long* rowbest;
  unsigned int* colbest;
  float* conf_out;   // debug tap: full confidence matrix [nb, L, S] (null in the product path)
};

enum { EPI_STORE = 0, EPI_CORR_STATS = 1, EPI_CORR_CONF = 2 };


// ------------------------------------------------------------------------------------------- tile iteration
// cluster == 1: CTA b walks tiles b, b + grid, ...   cluster == 2: the pair walks "pair tiles" (two consecutive m tiles,
// same n tile); CTA rank r of the pair owns m tile 2 * m_pair + r.  Both CTAs of a pair run the same number of iterations.
__device__ __forceinline__ int tile_count(const KParams& p) { return p.cluster == 2 ? p.num_tiles / 2 : p.num_tiles; }
__device__ __forceinline__ int tile_first(const KParams& p) { return p.cluster == 2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x; }
__device__ __forceinline__ int tile_step(const KParams& p) { return p.cluster == 2 ? (int)(gridDim.x >> 1) : (int)gridDim.x; }
__device__ __forceinline__ int tile_linear(const KParams& p, int it, uint32_t rank) {
  if (p.cluster != 2) return it;
  const int m_pair = it / p.n_tiles, n_tile = it - m_pair * p.n_tiles;
  return (m_pair * 2 + (int)rank) * p.n_tiles + n_tile;
}

// ------------------------------------------------------------------------------------------- tile decode
struct TileCoord {
  int m_tile, n_tile;
  int img, oh0, ow0;  // mode 1
};
__device__ __forceinline__ TileCoord decode_tile(const KParams& p, int t) {
  TileCoord c;
  c.img = 0; c.oh0 = 0; c.ow0 = 0;
  if (p.nb > 1) {  // batched row problems (coarse matching): t -> (batch, m_tile, n_tile)
    c.img = t / p.tiles_per_batch;
    t -= c.img * p.tiles_per_batch;
  }
  c.m_tile = t / p.n_tiles;
  c.n_tile = t - c.m_tile * p.n_tiles;
  if (p.mode == 1) {
    c.img = c.m_tile / p.tiles_per_img;
    int r = c.m_tile - c.img * p.tiles_per_img;
    int th = r / p.tiles_w;
    c.oh0 = th * TH;
    c.ow0 = (r - th * p.tiles_w) * TW;
  }
  return c;
}


// output / residual selection of the store epilogue (kernel template parameter OUT)
enum { OUT_F32 = 1, OUT_PLANES = 2, OUT_RESIDUAL = 4, OUT_RES_PLANES = 8 };  // RES_PLANES: residual given as fp16 planes

// Copy one of the four register-resident 32-column accumulator groups into the warp's smem tile.  The group loop in
// the epilogues is deliberately NOT unrolled (one copy of the per-group code keeps the SASS small enough for the
// instruction cache); the switch gives every case static register indices.
__device__ __forceinline__ void stage_group(float* tb, int lane, const float (&acc)[4][32], int gi) {
  float4* dst = reinterpret_cast<float4*>(tb + lane * TBP);
#define GIMB_STAGE(G)                                                                                   \
  _Pragma("unroll") for (int j4 = 0; j4 < 8; ++j4)                                                      \
      dst[j4] = make_float4(acc[G][j4 * 4], acc[G][j4 * 4 + 1], acc[G][j4 * 4 + 2], acc[G][j4 * 4 + 3]);
  switch (gi) {
    case 0: GIMB_STAGE(0) break;
    case 1: GIMB_STAGE(1) break;
    case 2: GIMB_STAGE(2) break;
    default: GIMB_STAGE(3) break;
  }
#undef GIMB_STAGE
}


// =========================================================================================== TMA epilogue
// Row-per-lane epilogue: the accumulator block arrives with TMEM lane = row, and that is also how it leaves.  Each
// epilogue warp owns one staging tile per tensor ([32 rows][32 channels]; SWIZZLE_64B for the two fp16 planes,
// SWIZZLE_128B for fp32) in shared memory: the lane writes its own row with 128-bit stores (the swizzle spreads every
// quarter-warp over all 32 banks) and ONE bulk tensor store per tensor moves the block to global memory - no
// transposition round trip, no per-row address arithmetic or predicates (the tensor map clips rows / channels that
// fall outside the tensor), no LSU store traffic.  The residual block comes in the same way (bulk tensor load
// signalled on a per-warp mbarrier) and is requested one block ahead of its use - also across tile boundaries - so
// its DRAM latency hides behind a whole block of work.
constexpr int STG_BLOCK = 4096;  // bytes of one staged 32 x 32 block (fp32, or both fp16 planes back to back)
template <int OUT>
struct StageLayout {
  static constexpr bool kRes = (OUT & (OUT_RESIDUAL | OUT_RES_PLANES)) != 0;
  static constexpr int off_res = 0;
  static constexpr int off_f32 = kRes ? STG_BLOCK : 0;
  static constexpr int off_pl = off_f32 + ((OUT & OUT_F32) ? STG_BLOCK : 0);
  static constexpr int bytes = off_pl + ((OUT & OUT_PLANES) ? STG_BLOCK : 0);
};
__device__ __forceinline__ uint32_t sw64_off(int row, int chunk) { return (uint32_t)(row * 64 + ((chunk ^ ((row >> 1) & 3)) << 4)); }
__device__ __forceinline__ uint32_t sw128_off(int row, int chunk) { return (uint32_t)(row * 128 + ((chunk ^ (row & 7)) << 4)); }
__device__ __forceinline__ uint4 lds128(uint32_t a) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void sts128(uint32_t a, uint4 v) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// lane 0 of the warp: ask for the residual block (rows of quadrant q of tile tc, channels c .. c+31)
template <int OUT>
__device__ __forceinline__ void request_residual(const KParams& p, const TMaps& maps, uint32_t dst, uint32_t bar,
                                                 const TileCoord& tc, int q, int c) {
  mbar_expect_tx(bar, STG_BLOCK);
  if (p.mode == 0) {
    const int r0 = tc.m_tile * BM + q * 32;
    if (OUT & OUT_RES_PLANES) {
      tma_load_3d(dst, &maps.r_hi, bar, c, r0, 0);
      tma_load_3d(dst + STG_BLOCK / 2, &maps.r_lo, bar, c, r0, 0);
    } else {
      tma_load_3d(dst, &maps.r_f32, bar, c, r0, 0);
    }
  } else {
    const int oh = tc.oh0 + q * 2;
    if (OUT & OUT_RES_PLANES) {
      tma_load_4d(dst, &maps.r_hi, bar, c, tc.ow0, oh, tc.img);
      tma_load_4d(dst + STG_BLOCK / 2, &maps.r_lo, bar, c, tc.ow0, oh, tc.img);
    } else {
      tma_load_4d(dst, &maps.r_f32, bar, c, tc.ow0, oh, tc.img);
    }
  }
}

// One 32 x 32 block of the store epilogue, lane = row: v[j] is the (LayerNorm-ed) accumulator of channel c + j.
template <int OUT, bool kSlowAct>
__device__ __forceinline__ void finish_block(const KParams& p, const TMaps& maps, float (&v)[32], uint32_t wb, uint32_t rbar,
                                             uint32_t& rphase, bool has_res, int lane, int q, const TileCoord& tc, int c,
                                             bool keep, bool have_next, const TileCoord& tcn, int cn) {
  using SL = StageLayout<OUT>;
  // ---- per-channel affine (folded BatchNorm / bias / LayerNorm gamma, beta); channels >= N give exactly 0
  if (p.scale) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      float4 sc = make_float4(0.f, 0.f, 0.f, 0.f), bi = sc;
      if (c + 4 * k < p.N) {
        sc = __ldg(reinterpret_cast<const float4*>(p.scale + c + 4 * k));
        bi = __ldg(reinterpret_cast<const float4*>(p.bias + c + 4 * k));
      }
      v[4 * k + 0] = fmaf(v[4 * k + 0], sc.x, bi.x);
      v[4 * k + 1] = fmaf(v[4 * k + 1], sc.y, bi.y);
      v[4 * k + 2] = fmaf(v[4 * k + 2], sc.z, bi.z);
      v[4 * k + 3] = fmaf(v[4 * k + 3], sc.w, bi.w);
      // keeps the 16 parameter loads from being hoisted into one 64-register burst (the accumulate path already
      // holds 128 accumulator registers)
      if (k & 1) asm volatile("" ::: "memory");
    }
  }
  // ---- residual block (requested one block ago)
  if (SL::kRes && has_res) {
    mbar_wait(rbar, rphase);
    rphase ^= 1u;
    if (OUT & OUT_RES_PLANES) {
      // identity carried as fp16 planes: x = hi + lo * 2^-8 (exact to 2^-22 relative)
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) {
        const uint4 h = lds128(wb + SL::off_res + sw64_off(lane, ch));
        const uint4 l = lds128(wb + SL::off_res + STG_BLOCK / 2 + sw64_off(lane, ch));
        const uint32_t hw[4] = {h.x, h.y, h.z, h.w}, lw[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&hw[e]));
          const float2 lf = __half22float2(*reinterpret_cast<const __half2*>(&lw[e]));
          v[ch * 8 + e * 2 + 0] += fmaf(lf.x, 1.f / kSplitScale, hf.x);
          v[ch * 8 + e * 2 + 1] += fmaf(lf.y, 1.f / kSplitScale, hf.y);
        }
      }
    } else {
#pragma unroll
      for (int ch = 0; ch < 8; ++ch) {
        const uint4 f = lds128(wb + SL::off_res + sw128_off(lane, ch));
        v[ch * 4 + 0] += __uint_as_float(f.x);
        v[ch * 4 + 1] += __uint_as_float(f.y);
        v[ch * 4 + 2] += __uint_as_float(f.z);
        v[ch * 4 + 3] += __uint_as_float(f.w);
      }
    }
    __syncwarp();  // every lane has its row: the buffer can take the next block
    if (have_next && lane == 0) request_residual<OUT>(p, maps, wb + SL::off_res, rbar, tcn, q, cn);
  }
  // ---- activation (uniform over the block: act_split is a multiple of 32)
  const int act = c >= p.act_split ? p.act1 : p.act0;
  if (kSlowAct) {
    if (act == ACT_ELU1) {
      // elu(x) + 1 = x + 1 (x > 0), exp(x) (x <= 0).  exp through one MUFU.EX2 on x * log2(e): for x <= 0 the absolute
      // error is <= |x| e^x * 1.44 * 2^-24 + 2 ulp(e^x) < 4e-8 - below half an ulp of the values near 1 it is added to
      // downstream (expf's range reduction cost ~3x the instructions of the whole block epilogue).
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        float ex;
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ex) : "f"(fminf(v[j], 0.f) * 1.4426950408889634f));
        v[j] = v[j] > 0.f ? v[j] + 1.f : ex;
      }
    } else if (act == ACT_DIVS) {
      // x / div, correctly rounded: q = x * r with r = RN(1 / div), one residual correction (Markstein); the operands
      // here are O(1) activations and a token count, far from the under/overflow cases the library division guards
      const float d = p.div, r = __frcp_rn(d);
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float q0 = v[j] * r;
        v[j] = fmaf(fmaf(-q0, d, v[j]), r, q0);
      }
    } else if (act == ACT_RELU) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
    } else if (act == ACT_LEAKY) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.01f * v[j]);
    }
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (!keep || c + j >= p.N) v[j] = 0.f;  // masked rows; pad channels of the planes stay zero
  } else {
    if (act == ACT_RELU) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
    } else if (act == ACT_LEAKY) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.01f * v[j]);
    }
  }
  // ---- stage the row and hand the block to the TMA unit.  The previous block's stores must have finished READING
  // the staging tiles (they had this whole block's math to do so).
  if (lane == 0) bulk_wait_read0();
  __syncwarp();
  if (OUT & OUT_F32) {
#pragma unroll
    for (int ch = 0; ch < 8; ++ch)
      sts128(wb + SL::off_f32 + sw128_off(lane, ch),
             make_uint4(__float_as_uint(v[ch * 4]), __float_as_uint(v[ch * 4 + 1]), __float_as_uint(v[ch * 4 + 2]),
                        __float_as_uint(v[ch * 4 + 3])));
  }
  if (OUT & OUT_PLANES) {
#pragma unroll
    for (int ch = 0; ch < 4; ++ch) {
      uint32_t hw[4], lw[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float x0 = v[ch * 8 + e * 2], x1 = v[ch * 8 + e * 2 + 1];
        const __half2 h = __floats2half2_rn(x0, x1);
        const float2 hf = __half22float2(h);
        const __half2 l = __floats2half2_rn((x0 - hf.x) * kSplitScale, (x1 - hf.y) * kSplitScale);
        hw[e] = *reinterpret_cast<const uint32_t*>(&h);
        lw[e] = *reinterpret_cast<const uint32_t*>(&l);
      }
      sts128(wb + SL::off_pl + sw64_off(lane, ch), make_uint4(hw[0], hw[1], hw[2], hw[3]));
      sts128(wb + SL::off_pl + STG_BLOCK / 2 + sw64_off(lane, ch), make_uint4(lw[0], lw[1], lw[2], lw[3]));
    }
  }
  fence_async_smem();
  __syncwarp();
  if (lane == 0) {
    if (p.mode == 0) {
      const int r0 = tc.m_tile * BM + q * 32;
      if (OUT & OUT_F32) tma_store_3d(&maps.o_f32, wb + SL::off_f32, c, r0, 0);
      if (OUT & OUT_PLANES) {
        tma_store_3d(&maps.o_hi, wb + SL::off_pl, c, r0, 0);
        tma_store_3d(&maps.o_lo, wb + SL::off_pl + STG_BLOCK / 2, c, r0, 0);
      }
    } else {
      const int oh = tc.oh0 + q * 2;
      if (OUT & OUT_F32) tma_store_4d(&maps.o_f32, wb + SL::off_f32, c, tc.ow0, oh, tc.img);
      if (OUT & OUT_PLANES) {
        tma_store_4d(&maps.o_hi, wb + SL::off_pl, c, tc.ow0, oh, tc.img);
        tma_store_4d(&maps.o_lo, wb + SL::off_pl + STG_BLOCK / 2, c, tc.ow0, oh, tc.img);
      }
    }
    bulk_commit();
  }
}

// ---- per-group (32 rows x 32 columns per warp) bodies of the coarse-matching epilogues.  `tb` is the warp's
// padded smem tile holding the raw accumulators row-per-lane: tb[lane * 33 + j] = 2^8 * <f0[row], f1[col j]>.
__device__ __forceinline__ void corr_stats_group(const KParams& p, float* tb, float (&v)[32], int lane, int q, int img, int m_tile,
                                                 int cbase, bool row_ok, bool rmasked, float& rm, float& rs) {
  // row view (lane = row): v[j] = raw accumulator of (row, cbase + j) on entry -> sim, straight-line and register resident
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] *= p.sim_scale;
  if (p.mask1) {
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const int c = cbase + j;
      if (rmasked || (c < p.S && p.mask1[(long long)img * p.S + c] == 0)) v[j] = -1e9f;
    }
  } else if (rmasked) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = -1e9f;
  }
  float gm = -INFINITY;
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    if (cbase + j >= p.S || !row_ok) v[j] = -INFINITY;
    gm = fmaxf(gm, v[j]);
  }
  // row partial (softmax over dim 2): online (max, sum exp) over this thread's columns
  if (gm > rm) { rs *= __expf(rm - gm); rm = gm; }
  if (rm > -INFINITY) {
    float a0 = 0.f, a1 = 0.f;
#pragma unroll
    for (int j = 0; j < 32; j += 2) { a0 += __expf(v[j] - rm); a1 += __expf(v[j + 1] - rm); }
    rs += a0 + a1;
  }
  // column partial (softmax over dim 1): write the sim block back, read it column-per-lane (conflict-free with the
  // 36-float pitch) and reduce over the 32 rows in registers - no shuffles
  __syncwarp();
#pragma unroll
  for (int j4 = 0; j4 < 8; ++j4)
    *reinterpret_cast<float4*>(tb + lane * TBP + j4 * 4) = make_float4(v[j4 * 4], v[j4 * 4 + 1], v[j4 * 4 + 2], v[j4 * 4 + 3]);
  __syncwarp();
  float cm = -INFINITY;
#pragma unroll
  for (int r = 0; r < 32; ++r) {
    v[r] = tb[r * TBP + lane];
    cm = fmaxf(cm, v[r]);
  }
  float c0s = 0.f, c1s = 0.f;
  if (cm > -INFINITY) {
#pragma unroll
    for (int r = 0; r < 32; r += 2) { c0s += __expf(v[r] - cm); c1s += __expf(v[r + 1] - cm); }
  }
  const int c = cbase + lane;
  if (c < p.S) p.colpart[((long long)img * p.S + c) * p.col_parts + m_tile * 4 + q] = make_float2(cm, c0s + c1s);
}

__device__ __forceinline__ void corr_conf_group(const KParams& p, const float* tb, int lane, int img, int cbase, long long row,
                                                bool row_ok, bool rmasked, unsigned long long& best) {
  // called by the whole warp (shuffles inside); rows beyond L get rmax = +inf and never produce candidates
  const float2 rst = row_ok ? p.rowstat[row] : make_float2(INFINITY, 1.f);
  const long long gc0 = (long long)img * p.S + cbase;
  const int ncol = min(32, p.S - cbase);
  // conf = softmax_col * softmax_row <= exp((x - rmax) + (x - cmax)): only entries that can exceed the threshold are
  // evaluated exactly; all others can never be a match (coarse_matching.py:174-190).  Pass 1 builds the candidate
  // bit mask branch-free; the (rare) candidates are then evaluated with the reference's formula.
  unsigned cand = 0u;
  float xs[32];
#pragma unroll
  for (int j4 = 0; j4 < 8; ++j4) {
    const float4 t4 = *reinterpret_cast<const float4*>(tb + lane * TBP + j4 * 4);
    xs[j4 * 4] = t4.x; xs[j4 * 4 + 1] = t4.y; xs[j4 * 4 + 2] = t4.z; xs[j4 * 4 + 3] = t4.w;
  }
  // column maxima of this group: one coalesced load (lane = column), broadcast by shuffle
  const float my_cmax = (lane < ncol) ? __ldg(&p.colstat[gc0 + lane].x) : INFINITY;
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    float x = xs[j] * p.sim_scale;
    const float cmax = __shfl_sync(0xffffffffu, my_cmax, j);
    if (j < ncol && (rmasked || (p.mask1 && p.mask1[gc0 + j] == 0))) x = -1e9f;
    const float t = (x - rst.x) + (x - cmax);
    cand |= (t > p.thr_log) ? (1u << j) : 0u;
  }
  if (p.conf_out && row_ok) {
    // debug tap (tests): every entry with the same formula the candidates use below; lane = row, 32 columns
    float* dst = p.conf_out + row * (long long)p.S + cbase;
#pragma unroll 1
    for (int j = 0; j < ncol; ++j) {
      float x = tb[lane * TBP + j] * p.sim_scale;
      if (rmasked || (p.mask1 && p.mask1[gc0 + j] == 0)) x = -1e9f;
      const float2 cst = __ldg(&p.colstat[gc0 + j]);
      dst[j] = __fdiv_rn(expf(x - cst.x), cst.y) * __fdiv_rn(expf(x - rst.x), rst.y);
    }
  }
  while (cand) {
    const int j = __ffs(cand) - 1;
    cand &= cand - 1;
    float x = tb[lane * TBP + j] * p.sim_scale;
    if (rmasked || (p.mask1 && p.mask1[gc0 + j] == 0)) x = -1e9f;
    const float2 cst = __ldg(&p.colstat[gc0 + j]);
    const float conf = __fdiv_rn(expf(x - cst.x), cst.y) * __fdiv_rn(expf(x - rst.x), rst.y);
    const unsigned int bits = __float_as_uint(conf);
    const unsigned int c = (unsigned int)(cbase + j);
    const unsigned long long pk = ((unsigned long long)bits << 32) | (unsigned long long)(~c);
    best = pk > best ? pk : best;
    atomicMax(&p.colbest[gc0 + j], bits);
  }
}

// kPair: the tcgen05.mma.cta_group::2 build of the kernel (p.pair_mma == 1).  A separate instantiation because a kernel
// that contains cta_group::2 instructions can only be launched with a cluster of 2.
template <int EPI, int OUT, bool kSlowAct, bool kLN, bool kPair>
__global__ void __launch_bounds__(NUM_THREADS, 1) umma_gemm_kernel(const __grid_constant__ TMaps maps, const KParams p) {
  constexpr bool kTma = EPI == EPI_STORE;  // store kernels: TMA epilogue; correlation sweeps: statistics epilogues
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;                 // stage ring, 1024-byte aligned
  uint8_t* gen = smem_raw + (base - raw);
  const uint32_t bars = base + p.stages * p.stage_bytes;        // barrier block
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gen + p.stages * p.stage_bytes + 256);
  float* tbuf = reinterpret_cast<float*>(gen + p.stages * p.stage_bytes + 512);  // [NUM_EPI_WARPS][32][TBP] transpose tiles
  // [2][4 quadrants][2 halves][32 rows]; the TMA epilogue has no transpose tiles, its staging tiles start at p.stg_off
  float* lnstat = kTma ? tbuf : tbuf + NUM_EPI_WARPS * 32 * TBP;
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 64u + 8u * s; };
  auto tfull_bar = [&](int b) { return bars + 128u + 8u * b; };   // up to 4 accumulator buffers
  auto tempty_bar = [&](int b) { return bars + 160u + 8u * b; };
  auto res_bar = [&](int w) { return bars + 192u + 8u * w; };     // TMA epilogue: residual block landed (per epilogue warp)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t cta_rank = (p.cluster == 2) ? cluster_ctarank() : 0u;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&maps.a_hi[0]);
    tma_prefetch_desc(&maps.a_lo[0]);
    tma_prefetch_desc(&maps.b_lo);
    tma_prefetch_desc(&maps.b_hi);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < p.stages; ++s) {
        mbar_init(full_bar(s), 1);
        // multicast pairs: released by the MMA warp of every CTA that reads/writes the stage; pair MMA: by the leader's commit
        mbar_init(empty_bar(s), (p.cluster == 2 && !kPair) ? 2u : 1u);
      }
      for (int b = 0; b < 4; ++b) {
        mbar_init(tfull_bar(b), 1);
        // pair MMA: the leader's issuer waits for the epilogue warps of BOTH CTAs before reusing an accumulator buffer
        mbar_init(tempty_bar(b), kPair ? 2u * NUM_EPI_WARPS : (uint32_t)NUM_EPI_WARPS);
      }
      for (int w = 0; w < NUM_EPI_WARPS; ++w) mbar_init(res_bar(w), 1);
      fence_barrier_init();
    }
    __syncwarp();
    if constexpr (kPair) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TMEM_COLS)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TMEM_COLS)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  if (p.cluster == 2) cluster_sync_all();  // peer barriers are initialised before any remote arrive / multicast
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const uint32_t a_plane = (uint32_t)BM * (uint32_t)p.bk * 2u;    // bytes of one A plane tile
  const uint32_t b_plane = (uint32_t)p.bn * (uint32_t)p.bk * 2u;  // bytes of one B plane tile
  const bool wide = p.bk == 64;

  if (warp == 0) {
    // =============================================================== TMA producer
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGS_LOW));
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int it = tile_first(p); it < tile_count(p); it += tile_step(p)) {
        const TileCoord tc = decode_tile(p, tile_linear(p, it, cta_rank));
        const int n0 = tc.n_tile * p.bn;
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1);
          const uint32_t sA = base + stage * p.stage_bytes;
          const uint32_t sB = sA + 2 * a_plane;
          // pair MMA: both CTAs' loads complete on the LEADER's barrier, which its producer arms for both stages' bytes
          // (the peer's bytes may land before the leader arms: the phase cannot complete before the leader's arrival)
          constexpr bool pm = kPair;
          const uint32_t fb = pm ? mapa_cta(full_bar(stage), 0u) : full_bar(stage);
          if (!pm) mbar_expect_tx(fb, (uint32_t)p.stage_bytes);
          else if (cta_rank == 0) mbar_expect_tx(full_bar(stage), 2u * (uint32_t)p.stage_bytes);
          auto tma_load_3d = [&](uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2) {
            if constexpr (kPair) tma_load_3d_pair(dst, m, bar, c0, c1, c2);
            else tma_load_3d_cta(dst, m, bar, c0, c1, c2);
          };
          auto tma_load_4d = [&](uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2, int c3) {
            if constexpr (kPair) tma_load_4d_pair(dst, m, bar, c0, c1, c2, c3);
            else tma_load_4d_cta(dst, m, bar, c0, c1, c2, c3);
          };
          int bk;  // k coordinate into the weight planes
          if (p.mode == 0) {
            const int m0 = tc.m_tile * BM;
            if (kb < p.cb1) {
              bk = kb * p.bk;
              tma_load_3d(sA, &maps.a_hi[0], fb, kb * p.bk, m0, tc.img);
              tma_load_3d(sA + a_plane, &maps.a_lo[0], fb, kb * p.bk, m0, tc.img);
            } else {
              const int k2 = (kb - p.cb1) * p.bk;
              bk = p.K1 + k2;
              tma_load_3d(sA, &maps.a2_hi, fb, k2, m0, 0);
              tma_load_3d(sA + a_plane, &maps.a2_lo, fb, k2, m0, 0);
            }
          } else {
            const int tap = kb / p.cb1, cb = kb - tap * p.cb1;
            const int kh = tap / p.KW, kw = tap - kh * p.KW;
            int dh = kh - p.pad, dw = kw - p.pad, view = 0;
            if (p.stride == 2) {
              const int py = dh & 1, px = dw & 1;
              view = py * 2 + px;
              dh = (dh - py) / 2;
              dw = (dw - px) / 2;
            }
            bk = tap * p.ldk + cb * p.bk;
            tma_load_4d(sA, &maps.a_hi[view], fb, cb * p.bk, tc.ow0 + dw, tc.oh0 + dh, tc.img);
            tma_load_4d(sA + a_plane, &maps.a_lo[view], fb, cb * p.bk, tc.ow0 + dw, tc.oh0 + dh, tc.img);
          }
          const int bb = (p.mode == 0) ? tc.img : 0;  // weights: one matrix; coarse matching: f1 of the same pair
          if (pm) {
            // the CTA's half of the B rows (the pair MMA reads the other half from the peer's shared memory)
            const int nh = n0 + (int)cta_rank * (p.bn >> 1);
            tma_load_3d(sB, &maps.bh_hi, fb, bk, nh, bb);
            tma_load_3d(sB + (b_plane >> 1), &maps.bh_lo, fb, bk, nh, bb);
          } else if (p.cluster == 2) {
            // each CTA of the pair fetches half of the B rows and multicasts them into both CTAs' stage
            const uint32_t hoff = cta_rank * (b_plane >> 1);
            const int nh = n0 + (int)cta_rank * (p.bn >> 1);
            tma_load_3d_mc(sB + hoff, &maps.bh_hi, fb, bk, nh, bb, (uint16_t)3);
            tma_load_3d_mc(sB + b_plane + hoff, &maps.bh_lo, fb, bk, nh, bb, (uint16_t)3);
          } else {
            tma_load_3d(sB, &maps.b_hi, fb, bk, n0, bb);
            tma_load_3d(sB + b_plane, &maps.b_lo, fb, bk, n0, bb);
          }
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // =============================================================== MMA issuer
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGS_LOW));
    if (lane == 0 && !(kPair && cta_rank != 0)) {  // pair MMA: the leader CTA issues for both
      constexpr bool pm = kPair;
      int stage = 0;
      uint32_t phase = 0;
      uint32_t cc = 0;  // chunk counter across tiles: TMEM buffer = cc mod (number of buffers)
      const uint64_t dconst = make_desc(0u, wide);  // descriptor with a zero start address
      const int nk16 = p.bk >> 4;
      for (int it = tile_first(p); it < tile_count(p); it += tile_step(p)) {
        int kb = 0;
        for (int ch = 0; ch < p.num_chunks; ++ch, ++cc) {
          const int buf = cc & ((1u << p.nbuf_log2) - 1u);
          mbar_wait(tempty_bar(buf), ((cc >> p.nbuf_log2) & 1) ^ 1);
          tc_fence_after();
          const uint32_t tacc = tmem_base + buf * p.acc_cols;
          const int kb_end = min(p.num_kb, kb + p.chunk_kb);
          // pass 1 over the chunk's stages: the two cross terms  D  = sum_k (A_hi*B_lo + A_lo*B_hi)      [lo = 2^8 * residual]
          // pass 2 over the same stages:    the main term        D  = sum_k A_hi*B_hi + 2^-8 * D         (scale-input-d = 8
          // on the first MMA of the pass).  The cross sum is accumulated at its own (small) magnitude and scaled exactly.
          // One thread issues every MMA of the CTA, so the loop is kept lean: the descriptors of a stage differ only in
          // the start-address field (bits 0-13, 16-byte units) - one add per operand instead of rebuilding them
          // (at N = 64 an MMA retires in ~40 cycles; a 30-instruction issue sequence made the issuer the bottleneck).
          int st = stage;
          uint32_t ph = phase;
          bool first = true;
          for (int k2 = kb; k2 < kb_end; ++k2) {
            mbar_wait(full_bar(st), ph);
            tc_fence_after();
            const uint64_t dA_hi = dconst + (((base + st * p.stage_bytes) & 0x3FFFFu) >> 4);
            const uint64_t dA_lo = dA_hi + (a_plane >> 4);
            const uint64_t dB_hi = dA_hi + (a_plane >> 3);
            const uint64_t dB_lo = dB_hi + (pm ? (b_plane >> 5) : (b_plane >> 4));  // pair MMA: half-height B planes
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {  // 16 fp16 = 32 bytes = 2 address units inside the swizzled tile row
              if (kk < nk16) {
                if constexpr (kPair) {
                  umma_f16_pair(tacc, dA_hi + 2 * kk, dB_lo + 2 * kk, p.idesc, (first && kk == 0) ? 0u : 1u);
                  umma_f16_pair(tacc, dA_lo + 2 * kk, dB_hi + 2 * kk, p.idesc, 1);
                } else {
                  umma_f16(tacc, dA_hi + 2 * kk, dB_lo + 2 * kk, p.idesc, (first && kk == 0) ? 0u : 1u);
                  umma_f16(tacc, dA_lo + 2 * kk, dB_hi + 2 * kk, p.idesc, 1);
                }
              }
            }
            first = false;
            if (++st == p.stages) { st = 0; ph ^= 1; }
          }
          first = true;
          for (; kb < kb_end; ++kb) {
            const uint64_t dA_hi = dconst + (((base + stage * p.stage_bytes) & 0x3FFFFu) >> 4);
            const uint64_t dB_hi = dA_hi + (a_plane >> 3);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
              if (kk < nk16) {
                if constexpr (kPair) {
                  if (first && kk == 0) umma_f16_scaled8_pair(tacc, dA_hi, dB_hi, p.idesc);
                  else umma_f16_pair(tacc, dA_hi + 2 * kk, dB_hi + 2 * kk, p.idesc, 1);
                } else {
                  if (first && kk == 0) umma_f16_scaled8(tacc, dA_hi, dB_hi, p.idesc);
                  else umma_f16(tacc, dA_hi + 2 * kk, dB_hi + 2 * kk, p.idesc, 1);
                }
              }
            }
            first = false;
            // smem stage reusable once every MMA of both passes has read it (in both CTAs of a pair)
            if constexpr (kPair) umma_commit_pair(empty_bar(stage));
            else if (p.cluster == 2) umma_commit_mc(empty_bar(stage), (uint16_t)3);
            else umma_commit(empty_bar(stage));
            if (++stage == p.stages) { stage = 0; phase ^= 1; }
          }
          if constexpr (kPair) umma_commit_pair(tfull_bar(buf));  // chunk accumulator complete (in both CTAs' TMEM)
          else umma_commit(tfull_bar(buf));
        }
      }
    }
  } else if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGS_LOW));
  } else {
    // =============================================================== epilogue warps
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REGS_HIGH));
    const int q = warp & 3;               // TMEM lane quadrant this warp may access
    const int half = (warp - 4) >> 2;     // which alternate 32-column groups this warp owns
    const int r_in_tile = q * 32 + lane;  // accumulator row owned by this thread
    uint32_t cc = 0;
    // accumulator buffer drained: tell the MMA issuer (pair MMA: the leader CTA's, from both CTAs)
    auto arrive_tempty = [&](int b) {
      if (kPair && cta_rank != 0) mbar_arrive_cluster(mapa_cta(tempty_bar(b), 0u));
      else mbar_arrive(tempty_bar(b));
    };
    // TMA epilogue state: staging tiles of this warp, blocks it owns per tile, residual pipeline
    const uint32_t wb = base + (uint32_t)p.stg_off + (uint32_t)(warp - 4) * (uint32_t)StageLayout<OUT>::bytes;
    const int ng = (p.bn - half * 32 + 63) / 64;
    const bool has_res = (OUT & OUT_RES_PLANES) || ((OUT & OUT_RESIDUAL) && p.residual != nullptr);
    uint32_t rphase = 0;
    if (kTma && EPI == EPI_STORE && StageLayout<OUT>::kRes && has_res && ng > 0 && lane == 0 && tile_first(p) < tile_count(p)) {
      const TileCoord t0 = decode_tile(p, tile_linear(p, tile_first(p), cta_rank));
      request_residual<OUT>(p, maps, wb, res_bar(warp - 4), t0, q, t0.n_tile * p.bn + half * 32);
    }
    for (int it = tile_first(p); it < tile_count(p); it += tile_step(p)) {
      const int t = tile_linear(p, it, cta_rank);
      const TileCoord tc = decode_tile(p, t);
      const int n0 = tc.n_tile * p.bn;
      long long row;
      bool row_ok;
      if (p.mode == 0) {
        row = (long long)tc.m_tile * BM + r_in_tile;
        row_ok = row < p.M;
        if (EPI != EPI_STORE) row += (long long)tc.img * p.L;  // global row of the batched problem
      } else {
        const int oh = tc.oh0 + r_in_tile / TW, ow = tc.ow0 + r_in_tile % TW;
        row_ok = oh < p.OH && ow < p.OW && tc.img < p.n_imgs;
        row = ((long long)tc.img * p.OH + oh) * p.OW + ow;
      }

      // the residual block to request once block gi of this tile has been consumed: the next block of the tile, else
      // block 0 of this CTA's next tile
      auto next_block = [&](int gi, bool& have_next, TileCoord& tcn, int& cn) {
        have_next = false;
        tcn = tc;
        cn = 0;
        if (!(StageLayout<OUT>::kRes && has_res)) return;
        if (gi + 1 < ng) {
          have_next = true;
          cn = n0 + ((gi + 1) * 2 + half) * 32;
        } else {
          const int itn = it + tile_step(p);
          if (itn < tile_count(p)) {
            have_next = true;
            tcn = decode_tile(p, tile_linear(p, itn, cta_rank));
            cn = tcn.n_tile * p.bn + half * 32;
          }
        }
      };
      if constexpr (kTma && EPI == EPI_STORE && !kLN) {
        if (p.num_chunks == 1) {
          // ---- single-chunk tiles (K <= 128): stream the accumulator block by block straight from TMEM through the
          // store epilogue (32 live accumulator registers instead of 128); the TMEM buffer is released after the last block
          const bool keep = !kSlowAct || p.row_mask == nullptr || (row_ok && p.row_mask[row] != 0);
          const int buf = cc & ((1u << p.nbuf_log2) - 1u);
          mbar_wait(tfull_bar(buf), (cc >> p.nbuf_log2) & 1);
          tc_fence_after();
          const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + buf * p.acc_cols;
#pragma unroll 1
          for (int gi = 0; gi < ng; ++gi) {
            const int c0 = (gi * 2 + half) * 32;
            uint32_t raw32[32];
            tmem_ld32(taddr + c0, raw32);
            tmem_ld_wait();
            float v[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(raw32[j]);
            if (gi + 1 == ng) {  // last TMEM read of this warp: hand the buffer back before the block's epilogue
              tc_fence_before();
              __syncwarp();
              if (lane == 0) arrive_tempty(buf);
            }
            bool have_next;
            TileCoord tcn;
            int cn;
            next_block(gi, have_next, tcn, cn);
            finish_block<OUT, kSlowAct>(p, maps, v, wb, res_bar(warp - 4), rphase, has_res, lane, q, tc, n0 + c0, keep,
                                        have_next, tcn, cn);
          }
          if (ng == 0) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) arrive_tempty(buf);
          }
          ++cc;
          continue;
        }
      }

      // ---- drain the chunk accumulators into fp32 registers (round-to-nearest adds)
      float acc[4][32];
#pragma unroll
      for (int gi = 0; gi < 4; ++gi)
#pragma unroll
        for (int j = 0; j < 32; ++j) acc[gi][j] = 0.f;
      for (int ch = 0; ch < p.num_chunks; ++ch, ++cc) {
        const int buf = cc & ((1u << p.nbuf_log2) - 1u);
        mbar_wait(tfull_bar(buf), (cc >> p.nbuf_log2) & 1);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + buf * p.acc_cols;
#pragma unroll
        for (int gi = 0; gi < 4; ++gi) {
          const int c0 = (gi * 2 + half) * 32;
          if (c0 < p.bn) {
            uint32_t raw32[32];
            tmem_ld32(taddr + c0, raw32);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) acc[gi][j] += __uint_as_float(raw32[j]);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) arrive_tempty(buf);
      }

      // ---- coarse-matching sweeps (networks/loftr/utils/coarse_matching.py:111-118, 174-190).
      // The per-group math lives in __noinline__ functions fed from the warp's smem tile: fully unrolled, the
      // epilogue was ~100 KB of straight-line SASS and ran instruction-fetch bound (ncu: stalled_no_instructions).
      if constexpr (EPI == EPI_CORR_STATS) {
        float* tb = tbuf + (warp - 4) * (32 * TBP);
        const bool rmasked = p.mask0 && row_ok && p.mask0[row] == 0;
        float rm = -INFINITY, rs = 0.f;
#pragma unroll 1
        for (int gi = 0; gi < 4; ++gi) {
          const int c0 = (gi * 2 + half) * 32;
          if (c0 >= p.bn) break;              // warp-uniform
          float v[32];
#define GIMB_TAKE(G) _Pragma("unroll") for (int j = 0; j < 32; ++j) v[j] = acc[G][j];
          switch (gi) {  // static register indices for every case; the group loop stays rolled (instruction cache)
            case 0: GIMB_TAKE(0) break;
            case 1: GIMB_TAKE(1) break;
            case 2: GIMB_TAKE(2) break;
            default: GIMB_TAKE(3) break;
          }
#undef GIMB_TAKE
          corr_stats_group(p, tb, v, lane, q, tc.img, tc.m_tile, n0 + c0, row_ok, rmasked, rm, rs);
        }
        if (row_ok) p.rowpart[row * p.row_parts + tc.n_tile * 2 + half] = make_float2(rm, rs);
        continue;
      }
      if constexpr (EPI == EPI_CORR_CONF) {
        float* tb = tbuf + (warp - 4) * (32 * TBP);
        const bool rmasked = p.mask0 && row_ok && p.mask0[row] == 0;
        unsigned long long best = 0ull;
#pragma unroll 1
        for (int gi = 0; gi < 4; ++gi) {
          const int c0 = (gi * 2 + half) * 32;
          if (c0 >= p.bn) break;
          __syncwarp();
          stage_group(tb, lane, acc, gi);
          corr_conf_group(p, tb, lane, tc.img, n0 + c0, row, row_ok, rmasked, best);
        }
        if (row_ok && best) atomicMax(&p.rowbest[row], best);
        continue;
      }

      // ---- store epilogue (multi-chunk tiles and fused LayerNorm): row-per-lane blocks through finish_block()
      if constexpr (kLN) {
        // fused LayerNorm over the full row (N == bn): this thread holds its row's columns of the alternate groups,
        // the partner warp of the quadrant the others; two-pass statistics exchanged through shared memory
        // (nn.LayerNorm eps 1e-5, networks/loftr/submodules/transformer.py:32-33; gamma/beta ride in scale/bias).
        const float inv_n = 1.f / (float)p.N;
        float* st1 = lnstat + (q * 2) * 32;
        float* st2 = lnstat + 256 + (q * 2) * 32;
        float sum = 0.f;
#pragma unroll
        for (int gi = 0; gi < 4; ++gi)
          if ((gi * 2 + half) * 32 < p.bn) {
#pragma unroll
            for (int j = 0; j < 32; ++j) sum += acc[gi][j];
          }
        st1[half * 32 + lane] = sum;
        asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");
        const float mean = (st1[lane] + st1[32 + lane]) * inv_n;
        float sq = 0.f;
#pragma unroll
        for (int gi = 0; gi < 4; ++gi)
          if ((gi * 2 + half) * 32 < p.bn) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float d = acc[gi][j] - mean;
              sq = fmaf(d, d, sq);
            }
          }
        st2[half * 32 + lane] = sq;
        asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");
        const float rstd = 1.f / sqrtf((st2[lane] + st2[32 + lane]) * inv_n + 1e-5f);
#pragma unroll
        for (int gi = 0; gi < 4; ++gi)
#pragma unroll
          for (int j = 0; j < 32; ++j) acc[gi][j] = (acc[gi][j] - mean) * rstd;
      }
      if constexpr (kTma) {
        const bool keep = !kSlowAct || p.row_mask == nullptr || (row_ok && p.row_mask[row] != 0);
#pragma unroll 1
        for (int gi = 0; gi < ng; ++gi) {  // rolled: one copy of the block epilogue (instruction cache)
          float v[32];
#define GIMB_TAKE(G) _Pragma("unroll") for (int j = 0; j < 32; ++j) v[j] = acc[G][j];
          switch (gi) {  // static register indices for every case
            case 0: GIMB_TAKE(0) break;
            case 1: GIMB_TAKE(1) break;
            case 2: GIMB_TAKE(2) break;
            default: GIMB_TAKE(3) break;
          }
#undef GIMB_TAKE
          bool have_next;
          TileCoord tcn;
          int cn;
          next_block(gi, have_next, tcn, cn);
          finish_block<OUT, kSlowAct>(p, maps, v, wb, res_bar(warp - 4), rphase, has_res, lane, q, tc,
                                      n0 + (gi * 2 + half) * 32, keep, have_next, tcn, cn);
        }
      }
    }
    if (kTma && EPI == EPI_STORE && lane == 0) bulk_wait0();  // all bulk stores of this warp are complete
  }

  // ---- teardown (cluster mode: no CTA may exit while its peer can still multicast into it / arrive on its barriers)
  tc_fence_before();
  __syncthreads();
  if (p.cluster == 2) cluster_sync_all();
  if (warp == 1) {
    if constexpr (kPair) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
  }
}

// ------------------------------------------------------------------------------------------- fp32 -> planes
__global__ void split_planes_kernel(const float* __restrict__ src, long long rows, int cols, int src_ld, __half* hi,
                                    __half* lo, __half* h8, int ld) {
  long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  long long total = rows * ld;
  if (idx >= total) return;
  long long r = idx / ld;
  int c = (int)(idx - r * ld);
  float x = c < cols ? src[r * src_ld + c] : 0.f;
  __half h = __float2half_rn(x);
  float hf = __half2float(h);
  hi[idx] = h;
  lo[idx] = __float2half_rn((x - hf) * kSplitScale);
  if (h8) h8[idx] = __float2half_rn(hf * kSplitScale);
}


// CTA-pair multicast of the B operand: on (2) by default, GIMB_CLUSTER=1 turns it off
int cluster_setting() {
  static int v = 0;
  if (!v) {
    const char* e = getenv("GIMB_CLUSTER");
    v = (e && atoi(e) == 1) ? 1 : 2;
  }
  return v;
}

// k-blocks per in-TMEM accumulation chunk (GIMB_CHUNK_KB overrides the default for experiments)
int pair_mma_setting() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("GIMB_PAIR");
    v = (e && strcmp(e, "mma") == 0) ? 1 : 0;
  }
  return v;
}
int chunk_kb_setting() {
  static int v = 0;
  if (!v) {
    const char* e = getenv("GIMB_CHUNK_KB");
    v = e ? atoi(e) : CHUNK_KB_DEFAULT;
    if (v < 1) v = CHUNK_KB_DEFAULT;
  }
  return v;
}

// K extent of a ring stage: 0 = choose per layer (default), GIMB_BK=32 / 64 force it where possible
int bk_setting() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("GIMB_BK");
    v = e ? atoi(e) : 0;
    if (v != 32 && v != 64) v = 0;
  }
  return v;
}

// Tensor map whose box is one epilogue-warp block: 32 channels x 32 rows (row mode) or 32 channels x 16 x 2 pixels
// (conv mode: the two image rows of the 8 x 16 patch that belong to one TMEM lane quadrant).  `width` channels are
// addressable (pitch `ld` elements); everything outside is clipped on stores and zero-filled on loads.
int block_map(CUtensorMap* m, const void* ptr, bool f32, int mode, uint64_t width, uint64_t ld, uint64_t M, int B, int OH,
              int OW) {
  const uint64_t es = f32 ? 4 : 2;
  if (mode == 0) {
    uint64_t dims[3] = {width, M, 1};
    uint64_t strides[2] = {ld * es, M * ld * es};
    uint32_t box[3] = {32, 32, 1};
    return make_map(m, ptr, 3, dims, strides, box, f32);
  }
  uint64_t dims[4] = {width, (uint64_t)OW, (uint64_t)OH, (uint64_t)B};
  uint64_t strides[3] = {ld * es, (uint64_t)OW * ld * es, (uint64_t)OH * OW * ld * es};
  uint32_t box[4] = {32, TW, 2, 1};
  return make_map(m, ptr, 4, dims, strides, box, f32);
}


}  // namespace

int split_planes(Ctx& ctx, const float* src, int64_t rows, int cols, int src_ld, const SplitPlanes& dst) {
  GIMB_CHECK(dst.ld % 8 == 0 && dst.ld >= cols, "split_planes: plane pitch must be a multiple of 8 and >= cols");
  if (ctx.dry || rows == 0) return 0;
  long long total = rows * dst.ld;
  split_planes_kernel<<<(unsigned)cdiv64(total, 256), 256, 0, ctx.stream>>>(src, rows, cols, src_ld, dst.hi, dst.lo, dst.h8, dst.ld);
  ctx.launches++;
  GIMB_LAUNCH_CHECK();
  return 0;
}

int umma_gemm(Ctx& ctx, const UmmaGemm& g) {
  GIMB_CHECK(g.a.hi && g.a.lo && g.b.hi && g.b.lo, "umma_gemm: operand planes missing");
  // N % 4 != 0 is allowed when scale/bias (read as float4) are allocated up to the next multiple of 4 with zeros
  GIMB_CHECK(g.N >= 8, "umma_gemm: N must be >= 8");
  GIMB_CHECK(g.a.ld % 8 == 0 && g.b.ld % 8 == 0, "umma_gemm: plane pitches must be multiples of 8");
  GIMB_CHECK(g.stride == 1 || g.stride == 2, "umma_gemm: stride 1 or 2");
  GIMB_CHECK((g.scale == nullptr) == (g.bias == nullptr), "umma_gemm: scale and bias go together");
  GIMB_CHECK(g.out_f32 || g.out.hi, "umma_gemm: no output requested");
  if (g.out.hi) GIMB_CHECK(g.out.lo && g.out.ld % 8 == 0 && g.out.ld >= g.N && g.out.ld - g.N < 32, "umma_gemm: bad output planes");
  if (ctx.dry) return 0;

  KParams p = {};
  TMaps maps;
  memset(&maps, 0, sizeof(maps));
  p.mode = g.mode;
  p.N = g.N;
  p.n_tiles = cdiv(g.N, 256);
  // several n tiles: whole 32-column epilogue blocks per tile (a block that straddles the tile edge would store the next
  // tile's columns); a single tile only needs the MMA granularity of 16
  p.bn = p.n_tiles > 1 ? cdiv(cdiv(g.N, p.n_tiles), 32) * 32 : cdiv(g.N, 16) * 16;
  p.KH = g.KH; p.KW = g.KW; p.stride = g.stride; p.pad = g.pad;
  p.K1 = g.K1;
  // ---- shared-memory plan.  Epilogue staging first (TMA epilogue: 4 KB per staged tensor per epilogue warp), then the
  // K extent of a ring stage: 64 (128-byte tile rows) when the layer is K-heavy and two such stages fit - the L2 -> SM
  // path delivers ~1.5x the bytes per clock for 128-byte box rows (tools/probe_tma.py: conv patches 22.6 -> 33.3
  // B/clk/SM, row tiles 41.9 -> 49.4), and the K-heavy layers are bound by exactly that path.
  const bool want_f32 = g.out_f32 != nullptr, want_planes = g.out.hi != nullptr;
  const bool any_res = g.residual != nullptr || g.residual_planes.hi != nullptr;
  // staging bytes per epilogue warp (must match StageLayout<OUT> of the launched variant); the kernel variants with a
  // residual slot reserve it even when no residual is given
  const bool res_slot = any_res || (want_f32 && want_planes);
  const int wbytes = STG_BLOCK * ((res_slot ? 1 : 0) + (want_f32 ? 1 : 0) + (want_planes ? 1 : 0));
  // [alignment slack 1024][ring][barriers 512 | LayerNorm exchange 2048 (fused-LN kernels only) | pad][staging tiles]
  const int fixed_tma = g.layernorm ? 3072 : 1024;
  const int extra_tma = 1024 + fixed_tma + NUM_EPI_WARPS * wbytes;
  // CTA pairs (cluster of 2) when the layer is big enough to fill them (GIMB_CLUSTER=1 disables).  Default: two M = 128
  // MMAs, every B tile multicast into both CTAs.  GIMB_PAIR=mma: one tcgen05.mma.cta_group::2 over the pair, each CTA
  // stages half of B - implemented and parity-clean, but measured 15-25 % SLOWER per layer with the chunked split scheme
  // (profiles/r01_notes.md), so it is opt-in until that is understood.
  const int m_tiles_plan = g.mode == 0 ? (int)cdiv64(std::max<int64_t>(g.M, 1), BM) : g.B * cdiv(g.OH, TH) * cdiv(g.OW, TW);
  p.cluster = (cluster_setting() == 2 && m_tiles_plan >= 2 * ctx.sm_count && p.bn % 16 == 0 && p.bn >= 32) ? 2 : 1;
  p.pair_mma = (p.cluster == 2 && pair_mma_setting()) ? 1 : 0;
  {
    const int ktot = (g.mode == 0 ? g.K1 + g.K2 : g.KH * g.KW * g.K1);
    // the choice must not depend on the batch size (pairs only form for big layers): results stay bit-identical whether
    // a pair of images is processed alone or in a batch
    const int stage64 = 2 * BM * 128 + 2 * p.bn * 128;
    const int pref = bk_setting();
    p.bk = 32;
    // channels that do not fill the last 64-wide block (196 -> 256 instead of 224): the extra weight traffic and MMA
    // work only pays for narrow tiles, where the A operand dominates (measured: 196 -> 196 slower, 196 -> 128 faster)
    const bool pad_ok = cdiv(g.K1, 64) * 64 == cdiv(g.K1, 32) * 32 || p.bn <= 128;
    if (pref != 32 && (pref == 64 || (ktot >= 256 && pad_ok)) && (g.K2 == 0 || g.K1 % 64 == 0) &&
        (SMEM_LIMIT - extra_tma) / stage64 >= 2)
      p.bk = 64;
  }
  const int bkk = p.bk;
  p.cb1 = cdiv(g.K1, bkk);
  int m_tiles;
  if (g.mode == 0) {
    GIMB_CHECK(g.M > 0, "umma_gemm: M must be positive");
    p.M = g.M;
    m_tiles = (int)cdiv64(g.M, BM);
    GIMB_TRY(rows_map(&maps.a_hi[0], g.a.hi, g.K1, g.M, g.a.ld, BM, 1, bkk));
    GIMB_TRY(rows_map(&maps.a_lo[0], g.a.lo, g.K1, g.M, g.a.ld, BM, 1, bkk));
    if (g.K2 > 0) {
      GIMB_CHECK(g.a2.hi && g.a2.lo, "umma_gemm: concat planes missing");
      p.cb2 = cdiv(g.K2, bkk);
      GIMB_TRY(rows_map(&maps.a2_hi, g.a2.hi, g.K2, g.M, g.a2.ld, BM, 1, bkk));
      GIMB_TRY(rows_map(&maps.a2_lo, g.a2.lo, g.K2, g.M, g.a2.ld, BM, 1, bkk));
    }
    p.num_kb = p.cb1 + p.cb2;
    p.ldk = 0;
    const uint64_t Kw = (uint64_t)g.K1 + g.K2;
    GIMB_CHECK((uint64_t)g.b.ld >= Kw, "umma_gemm: weight pitch smaller than K");
    GIMB_TRY(rows_map(&maps.b_lo, g.b.lo, Kw, g.N, g.b.ld, p.bn, 1, bkk));
    GIMB_TRY(rows_map(&maps.b_hi, g.b.hi, Kw, g.N, g.b.ld, p.bn, 1, bkk));
  } else {
    GIMB_CHECK(g.K2 == 0, "umma_gemm: concat only in row mode");
    // stride 2 with odd H / W is fine: the four parity views get their own extents below
    p.OH = g.OH; p.OW = g.OW;
    p.tiles_h = cdiv(g.OH, TH); p.tiles_w = cdiv(g.OW, TW);
    p.tiles_per_img = p.tiles_h * p.tiles_w;
    m_tiles = g.B * p.tiles_per_img;
    p.M = (long long)g.B * g.OH * g.OW;
    p.num_kb = g.KH * g.KW * p.cb1;
    p.ldk = g.ldk;
    GIMB_CHECK(g.ldk >= g.K1 && g.ldk % 8 == 0, "umma_gemm: bad per-tap weight pitch");
    const uint64_t ld = g.a.ld;
    const int nviews = g.stride == 2 ? 4 : 1;
    for (int v = 0; v < nviews; ++v) {
      const int py = v >> 1, px = v & 1;
      const uint64_t s = g.stride;
      // view (py, px) holds the pixels (2y' + py, 2x' + px): (H + 1 - py) / 2 rows, (W + 1 - px) / 2 columns (odd sizes: the
      // views differ by one); everything outside is zero-filled = the convolution's padding
      uint64_t dims[4] = {(uint64_t)g.K1, (uint64_t)(g.stride == 2 ? (g.W + 1 - px) / 2 : g.W),
                          (uint64_t)(g.stride == 2 ? (g.H + 1 - py) / 2 : g.H), (uint64_t)g.B};
      uint64_t strides[3] = {s * ld * 2, s * (uint64_t)g.W * ld * 2, (uint64_t)g.H * g.W * ld * 2};
      uint32_t box[4] = {(uint32_t)bkk, TW, TH, 1};
      const size_t off = ((size_t)py * g.W + px) * ld;
      GIMB_TRY(make_map(&maps.a_hi[v], g.a.hi + off, 4, dims, strides, box));
      GIMB_TRY(make_map(&maps.a_lo[v], g.a.lo + off, 4, dims, strides, box));
    }
    const uint64_t Kw = (uint64_t)g.KH * g.KW * g.ldk;
    GIMB_CHECK((uint64_t)g.b.ld >= Kw, "umma_gemm: weight pitch smaller than K");
    GIMB_TRY(rows_map(&maps.b_lo, g.b.lo, Kw, g.N, g.b.ld, p.bn, 1, bkk));
    GIMB_TRY(rows_map(&maps.b_hi, g.b.hi, Kw, g.N, g.b.ld, p.bn, 1, bkk));
  }
  p.stage_bytes = 2 * BM * bkk * 2 + (p.pair_mma ? 1 : 2) * p.bn * bkk * 2;  // pair MMA: each CTA stages half of B
  const int extra = extra_tma;  // shared memory beside the operand ring
  GIMB_CHECK(g.act_split % 32 == 0, "umma_gemm: act_split must be a multiple of 32");
  {
    const uint64_t Mrows = (uint64_t)(g.mode == 0 ? g.M : (int64_t)g.B * g.OH * g.OW);
    const uint64_t ldo = g.out_f32_ld > 0 ? g.out_f32_ld : g.N, ldres = g.residual_ld > 0 ? g.residual_ld : g.N;
    GIMB_CHECK((int)ldo >= g.N && (int)ldres >= g.N && ldo % 4 == 0 && ldres % 4 == 0 && (int)ldo - g.N < 32,
               "umma_gemm: fp32 pitches must be multiples of 4, >= N and < N + 32");
    if (want_f32) GIMB_TRY(block_map(&maps.o_f32, g.out_f32, true, g.mode, ldo, ldo, Mrows, g.B, g.OH, g.OW));
    if (want_planes) {
      GIMB_TRY(block_map(&maps.o_hi, g.out.hi, false, g.mode, g.out.ld, g.out.ld, Mrows, g.B, g.OH, g.OW));
      GIMB_TRY(block_map(&maps.o_lo, g.out.lo, false, g.mode, g.out.ld, g.out.ld, Mrows, g.B, g.OH, g.OW));
    }
    if (g.residual) GIMB_TRY(block_map(&maps.r_f32, g.residual, true, g.mode, g.N, ldres, Mrows, g.B, g.OH, g.OW));
    if (g.residual_planes.hi) {
      const uint64_t ldr = g.residual_planes.ld;
      GIMB_TRY(block_map(&maps.r_hi, g.residual_planes.hi, false, g.mode, ldr, ldr, Mrows, g.B, g.OH, g.OW));
      GIMB_TRY(block_map(&maps.r_lo, g.residual_planes.lo, false, g.mode, ldr, ldr, Mrows, g.B, g.OH, g.OW));
    }
  }
  p.stages = std::min(MAX_STAGES, (SMEM_LIMIT - extra) / p.stage_bytes);
  p.stages = std::max(2, std::min(p.stages, std::max(3, p.num_kb + 1)));
  p.stg_off = p.stages * p.stage_bytes + fixed_tma;
  p.num_tiles = m_tiles * p.n_tiles;
  // K <= 128: one in-TMEM chunk (24 accumulation steps keep the truncation bias at the fp32-FFMA level and save a
  // drain hand-shake per tile); longer K: chunks of CHUNK_KB k-blocks
  p.chunk_kb = p.num_kb * bkk <= 128 ? p.num_kb : std::max(1, chunk_kb_setting() * 32 / bkk);
  p.chunk_kb = std::max(1, std::min(p.chunk_kb, p.stages - 1));  // a chunk's stages stay resident for both MMA passes
  p.num_chunks = cdiv(p.num_kb, p.chunk_kb);
  p.idesc = (1u << 4) | ((unsigned)(p.bn >> 3) << 17) | ((unsigned)(BM >> 4) << 24);
  p.acc_cols = p.bn <= 128 ? 128 : ACC_COLS;
  p.nbuf_log2 = p.bn <= 128 ? 2 : 1;
  p.scale = g.scale; p.bias = g.bias; p.residual = g.residual; p.row_mask = g.row_mask;
  p.act0 = g.act0; p.act1 = g.act1; p.act_split = g.act_split; p.div = g.div;
  p.out_f32 = g.out_f32; p.out_hi = g.out.hi; p.out_lo = g.out.lo; p.out_h8 = g.out.h8; p.ldp = g.out.ld;
  p.res_hi = g.residual_planes.hi; p.res_lo = g.residual_planes.lo; p.ldr = g.residual_planes.ld;

  p.nb = 1;
  p.n_imgs = g.mode == 1 ? g.B : 1;
  p.m_tiles_real = m_tiles;
  GIMB_CHECK(m_tiles == m_tiles_plan, "umma_gemm: tile plan mismatch");
  if (p.pair_mma) p.idesc = (1u << 4) | ((unsigned)(p.bn >> 3) << 17) | ((unsigned)((2 * BM) >> 4) << 24);  // M = 256 over the pair
  if (p.cluster == 2) {
    const int m_even = (m_tiles + 1) / 2 * 2;  // an odd tile count gets one dummy tile (TMA OOB zero fill, masked stores)
    p.num_tiles = m_even * p.n_tiles;
    const uint64_t Kw = g.mode == 0 ? (uint64_t)g.K1 + g.K2 : (uint64_t)g.KH * g.KW * g.ldk;
    GIMB_TRY(rows_map(&maps.bh_lo, g.b.lo, Kw, g.N, g.b.ld, p.bn / 2, 1, bkk));
    GIMB_TRY(rows_map(&maps.bh_hi, g.b.hi, Kw, g.N, g.b.ld, p.bn / 2, 1, bkk));
  }
  const int smem = p.stages * p.stage_bytes + extra;
  GIMB_CHECK(smem <= SMEM_LIMIT, "umma_gemm: shared memory plan %d B exceeds the limit", smem);
  const bool slow = g.act0 >= ACT_ELU1 || g.act1 >= ACT_ELU1 || g.row_mask != nullptr;
  const bool f32 = g.out_f32 != nullptr, planes = g.out.hi != nullptr, res = g.residual != nullptr;
  const bool resp = g.residual_planes.hi != nullptr;
  GIMB_CHECK(!(res && resp), "umma_gemm: residual given twice");
  if (resp) GIMB_CHECK(g.residual_planes.lo && g.residual_planes.ld % 8 == 0 && g.residual_planes.ld >= g.N, "umma_gemm: bad residual planes");
  GIMB_CHECK(g.out.h8 == nullptr, "umma_gemm: the GEMM epilogue does not produce the h8 plane");
  int grid = std::min(p.num_tiles, ctx.sm_count);
  if (p.cluster == 2) grid = std::min(p.num_tiles, ctx.sm_count / 2 * 2);
  cudaError_t aerr = cudaSuccess;
  cudaLaunchConfig_t lcfg = {};
  cudaLaunchAttribute lattr[1];
  lcfg.gridDim = dim3(grid); lcfg.blockDim = dim3(NUM_THREADS); lcfg.dynamicSmemBytes = smem; lcfg.stream = ctx.stream;
  lattr[0].id = cudaLaunchAttributeClusterDimension;
  lattr[0].val.clusterDim.x = p.cluster; lattr[0].val.clusterDim.y = 1; lattr[0].val.clusterDim.z = 1;
  lcfg.attrs = lattr; lcfg.numAttrs = 1;
#define GIMB_LAUNCH_VARIANT_T(OUTV, SLOWV, LNV, PAIRV)                                                             \
  do {                                                                                                            \
    GIMB_SMEM_OPTIN((umma_gemm_kernel<EPI_STORE, OUTV, SLOWV, LNV, PAIRV>), SMEM_LIMIT);                         \
    if (StageLayout<OUTV>::bytes != wbytes) {                                                                     \
      set_error("umma_gemm: staging plan (%d B) does not match the kernel variant (%d B)", wbytes,                \
                (int)StageLayout<OUTV>::bytes);                                                                   \
      return 1;                                                                                                   \
    }                                                                                                             \
    if (aerr == cudaSuccess)                                                                                      \
      aerr = cudaLaunchKernelEx(&lcfg, umma_gemm_kernel<EPI_STORE, OUTV, SLOWV, LNV, PAIRV>, maps, p);            \
  } while (0)
#define GIMB_LAUNCH_VARIANT_LN(OUTV, SLOWV, LNV)                        \
  do {                                                                  \
    if (p.pair_mma) GIMB_LAUNCH_VARIANT_T(OUTV, SLOWV, LNV, true);      \
    else GIMB_LAUNCH_VARIANT_T(OUTV, SLOWV, LNV, false);                \
  } while (0)
#define GIMB_LAUNCH_VARIANT(OUTV, SLOWV) GIMB_LAUNCH_VARIANT_LN(OUTV, SLOWV, false)
  if (g.layernorm) {
    GIMB_CHECK(g.N == p.bn && g.N % 64 == 0 && g.scale && !slow && !resp, "umma_gemm: fused LayerNorm needs N == tile N, gamma/beta");
    if (!f32 && planes && !res) GIMB_LAUNCH_VARIANT_LN(OUT_PLANES, false, true);
    else if (f32 && planes && res) GIMB_LAUNCH_VARIANT_LN(OUT_F32 | OUT_PLANES | OUT_RESIDUAL, false, true);
    else {
      set_error("umma_gemm: unsupported fused-LayerNorm output combination");
      return 1;
    }
  } else if (resp) {
    if (!f32 && planes && !slow) GIMB_LAUNCH_VARIANT(OUT_PLANES | OUT_RES_PLANES, false);
    else {
      set_error("umma_gemm: residual planes are supported for plane-only output");
      return 1;
    }
  } else if (f32 && !planes && !res && slow) GIMB_LAUNCH_VARIANT(OUT_F32, true);
  else if (f32 && !planes && !res) GIMB_LAUNCH_VARIANT(OUT_F32, false);
  else if (!f32 && planes && !res && !slow) GIMB_LAUNCH_VARIANT(OUT_PLANES, false);
  else if (f32 && planes && !slow) GIMB_LAUNCH_VARIANT(OUT_F32 | OUT_PLANES | OUT_RESIDUAL, false);
  else if (f32 && !planes && res && !slow) GIMB_LAUNCH_VARIANT(OUT_F32 | OUT_RESIDUAL, false);
  else {
    set_error("umma_gemm: unsupported epilogue combination (f32=%d planes=%d residual=%d slow_act=%d)", f32, planes, res, slow);
    return 1;
  }
#undef GIMB_LAUNCH_VARIANT
#undef GIMB_LAUNCH_VARIANT_LN
#undef GIMB_LAUNCH_VARIANT_T
  GIMB_CUDA(aerr);
  ctx.launches++;
  GIMB_LAUNCH_CHECK();
  return 0;
}


// ------------------------------------------------------------------------------------------- TMA load-rate probe
// Measurement helper (tools/probe_tma.py): every CTA (one per SM, one thread) streams boxes from an L2-resident
// tensor through a 4-stage ring.  Answers one design question: what does the L2 -> SM path deliver for 64-byte box
// rows (BK = 32 fp16, SWIZZLE_64B - what the GEMM uses) against 128-byte rows (BK = 64, SWIZZLE_128B)?
namespace {
__global__ void __launch_bounds__(32, 1) tma_probe_kernel(const __grid_constant__ CUtensorMap map, int variant, int box_bytes,
                                                          int boxes, int iters, int n_tiles) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t bars = base + 4 * 49152;
  if (threadIdx.x != 0) return;
  for (int s = 0; s < 4; ++s) mbar_init(bars + 8 * s, 1);
  fence_barrier_init();
  const int stage_bytes = box_bytes * boxes;
  auto issue = [&](int i) {
    const int s = i & 3;
    const uint32_t dst = base + s * 49152, bar = bars + 8 * s;
    mbar_expect_tx(bar, (uint32_t)stage_bytes);
    const int t = (int)(((long long)blockIdx.x * 7919 + (long long)i * gridDim.x) % n_tiles);
    for (int b = 0; b < boxes; ++b) {
      if (variant < 2) {  // row mode: tensor [rows][256 ch]; boxes walk the k-blocks of a 128-row tile
        const int bk = variant == 0 ? 32 : 64;
        const int kb = (i * boxes + b) % (256 / bk);
        tma_load_3d(dst + b * box_bytes, &map, bar, kb * bk, t * 128, 0);
      } else {           // conv mode: tensor [4][240][320][64 ch]; boxes are taps of a TH x TW patch (TH * TW = 128)
        const int bk = (variant & 1) ? 64 : 32;
        const int tw = variant < 4 ? 16 : (variant < 6 ? 32 : 64), th = 128 / tw;
        const int tiles_w = 320 / tw, per_img = tiles_w * (240 / th);
        const int img = t / per_img, r = t % per_img, ty = r / tiles_w, tx = r % tiles_w;
        const int tap = (i * boxes + b) % (9 * (64 / bk));
        const int cb = tap % (64 / bk), kk = tap / (64 / bk);
        tma_load_4d(dst + b * box_bytes, &map, bar, cb * bk, tx * tw + kk % 3 - 1, ty * th + kk / 3 - 1, img);
      }
    }
  };
  for (int i = 0; i < 4 && i < iters; ++i) issue(i);
  for (int i = 0; i < iters; ++i) {
    mbar_wait(bars + 8 * (i & 3), (i >> 2) & 1);
    if (i + 4 < iters) issue(i + 4);
  }
}
}  // namespace

int tma_probe(Ctx& ctx, int variant, int iters, float* gbps) {
  GIMB_CHECK(variant >= 0 && variant < 8 && iters > 0, "tma_probe: bad arguments");
  const bool conv = variant >= 2, wide = variant & 1;
  const size_t elems = conv ? (size_t)4 * 240 * 320 * 64 : (size_t)65536 * 256;
  __half* buf = nullptr;
  GIMB_CUDA(cudaMalloc(&buf, elems * 2));
  GIMB_CUDA(cudaMemsetAsync(buf, 0, elems * 2, ctx.stream));
  EncodeTiledFn enc;
  GIMB_TRY(get_encode(&enc));
  CUtensorMap map;
  cuuint64_t gd[4], gs[3];
  cuuint32_t bx[4], es[4] = {1, 1, 1, 1};
  int rank, n_tiles;
  const cuuint32_t bk = wide ? 64 : 32;
  if (!conv) {
    rank = 3; n_tiles = 65536 / 128;
    gd[0] = 256; gd[1] = 65536; gd[2] = 1; gs[0] = 512; gs[1] = 512ull * 65536;
    bx[0] = bk; bx[1] = 128; bx[2] = 1;
  } else {
    const cuuint32_t tw = variant < 4 ? 16 : (variant < 6 ? 32 : 64);
    rank = 4; n_tiles = 4 * 600;
    gd[0] = 64; gd[1] = 320; gd[2] = 240; gd[3] = 4; gs[0] = 128; gs[1] = 128ull * 320; gs[2] = 128ull * 320 * 240;
    bx[0] = bk; bx[1] = tw; bx[2] = 128 / tw; bx[3] = 1;
  }
  CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, (cuuint32_t)rank, buf, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   wide ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  GIMB_CHECK(r == CUDA_SUCCESS, "tma_probe: cuTensorMapEncodeTiled failed with %d", (int)r);
  const int box_bytes = (int)bk * 2 * 128, boxes = wide ? 2 : 4;  // 32 KB per ring stage either way
  const int smem = 4 * 49152 + 1024 + 64;
  GIMB_CUDA(cudaFuncSetAttribute(tma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  cudaEvent_t e0, e1;
  GIMB_CUDA(cudaEventCreate(&e0));
  GIMB_CUDA(cudaEventCreate(&e1));
  tma_probe_kernel<<<ctx.sm_count, 32, smem, ctx.stream>>>(map, variant, box_bytes, boxes, iters, n_tiles);  // warm L2
  GIMB_CUDA(cudaEventRecord(e0, ctx.stream));
  tma_probe_kernel<<<ctx.sm_count, 32, smem, ctx.stream>>>(map, variant, box_bytes, boxes, iters, n_tiles);
  GIMB_CUDA(cudaEventRecord(e1, ctx.stream));
  GIMB_CUDA(cudaEventSynchronize(e1));
  float ms = 0.f;
  GIMB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
  *gbps = (float)((double)ctx.sm_count * iters * box_bytes * boxes / (ms * 1e-3) / 1e9);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(buf);
  return 0;
}

// N-tile width of the correlation sweeps: 128 -> four TMEM accumulator buffers, the MMA warp runs a whole tile ahead of
// the (expensive) statistics epilogue; 256 -> two buffers, half as many tiles (GIMB_CORR_BN overrides)
int corr_bn() {
  static int v = 0;
  if (!v) {
    const char* e = getenv("GIMB_CORR_BN");
    v = (e && atoi(e) == 256) ? 256 : 128;
  }
  return v;
}

void umma_corr_parts(int L, int S, int* row_parts, int* col_parts) {
  *row_parts = cdiv(S, corr_bn()) * 2;
  *col_parts = cdiv(L, BM) * 4;
}

int umma_corr(Ctx& ctx, const UmmaCorr& c, int pass) {
  GIMB_CHECK(c.f0.hi && c.f0.lo && c.f1.hi && c.f1.lo, "umma_corr: operand planes missing");
  GIMB_CHECK(c.C % 32 == 0, "umma_corr: C must be a multiple of 32");
  if (ctx.dry || c.N == 0) return 0;
  KParams p = {};
  TMaps maps;
  memset(&maps, 0, sizeof(maps));
  p.mode = 0;
  p.N = c.S;
  p.bn = corr_bn();
  p.n_tiles = cdiv(c.S, p.bn);
  if (c.S < p.bn) p.bn = cdiv(c.S, 16) * 16;
  p.bk = BK;
  p.K1 = c.C; p.cb1 = c.C / BK; p.cb2 = 0; p.num_kb = p.cb1;
  p.KH = p.KW = 1; p.stride = 1; p.pad = 0;
  p.M = c.L;
  const int m_tiles = cdiv(c.L, BM);
  p.nb = c.N; p.L = c.L; p.S = c.S;
  p.cluster = 1; p.n_imgs = 1; p.m_tiles_real = m_tiles;
  p.tiles_per_batch = m_tiles * p.n_tiles;
  p.num_tiles = c.N * p.tiles_per_batch;
  GIMB_TRY(rows_map(&maps.a_hi[0], c.f0.hi, c.C, c.L, c.f0.ld, BM, c.N));
  GIMB_TRY(rows_map(&maps.a_lo[0], c.f0.lo, c.C, c.L, c.f0.ld, BM, c.N));
  GIMB_TRY(rows_map(&maps.b_lo, c.f1.lo, c.C, c.S, c.f1.ld, p.bn, c.N));
  GIMB_TRY(rows_map(&maps.b_hi, c.f1.hi, c.C, c.S, c.f1.ld, p.bn, c.N));
  p.stage_bytes = 2 * A_TILE_BYTES + 2 * p.bn * BK * 2;
  p.stages = std::max(2, std::min(MAX_STAGES, (SMEM_LIMIT - SMEM_EXTRA) / p.stage_bytes));
  {
    // k-blocks accumulated inside the tensor core between fp32 drains: 4 (128 k, the same 24 accumulation steps the
    // K <= 128 GEMM layers use; C = 256 -> two drains per tile).  Ids stay bit-exact on every golden case and the sweeps
    // get ~9 % faster than with 64-k chunks (the drain is part of their bottleneck).  GIMB_CORR_CHUNK_KB overrides.
    static int corr_chunk = 0;
    if (!corr_chunk) {
      const char* e = getenv("GIMB_CORR_CHUNK_KB");
      corr_chunk = (e && atoi(e) > 0) ? atoi(e) : 4;
    }
    p.chunk_kb = std::max(1, std::min(corr_chunk, p.stages - 1));
  }
  p.num_chunks = cdiv(p.num_kb, p.chunk_kb);
  p.idesc = (1u << 4) | ((unsigned)(p.bn >> 3) << 17) | ((unsigned)(BM >> 4) << 24);
  p.acc_cols = p.bn <= 128 ? 128 : ACC_COLS;
  p.nbuf_log2 = p.bn <= 128 ? 2 : 1;
  p.mask0 = c.mask0; p.mask1 = c.mask1;
  p.inv_c = 1.f / (float)c.C;
  p.temperature = c.temperature;
  p.thr_log = c.thr > 0.f ? logf(c.thr) - 1e-3f : -INFINITY;
  // sim = <f0, f1> / C / T.  One multiply by the fp32-rounded constant
  // (differs from the reference's `/ T` by at most 1 ulp of sim, far below the fp32 noise of the dot product).
  p.sim_scale = (float)(1.0 / ((double)c.C * (double)c.temperature));
  p.rowpart = c.rowpart; p.colpart = c.colpart;
  umma_corr_parts(c.L, c.S, &p.row_parts, &p.col_parts);
  p.rowstat = c.rowstat; p.colstat = c.colstat; p.rowbest = c.rowbest; p.colbest = c.colbest;
  p.conf_out = c.conf_matrix;
  // the multiplexing TMA batch coordinate is tile.img for both operands (mode 0)
  const int smem = p.stages * p.stage_bytes + SMEM_EXTRA;
  GIMB_SMEM_OPTIN((umma_gemm_kernel<EPI_CORR_STATS, 0, false, false, false>), SMEM_LIMIT);
  GIMB_SMEM_OPTIN((umma_gemm_kernel<EPI_CORR_CONF, 0, false, false, false>), SMEM_LIMIT);
  const int grid = std::min(p.num_tiles, ctx.sm_count);
  if (pass == 0)
    umma_gemm_kernel<EPI_CORR_STATS, 0, false, false, false><<<grid, NUM_THREADS, smem, ctx.stream>>>(maps, p);
  else
    umma_gemm_kernel<EPI_CORR_CONF, 0, false, false, false><<<grid, NUM_THREADS, smem, ctx.stream>>>(maps, p);
  ctx.launches++;
  GIMB_LAUNCH_CHECK();
  return 0;
}

}  // namespace gimb
