// probe_mma.cu - measurement helper (tools/probe_mma.py, libgimb200_test.so only): how many SM cycles does one
// tcgen05.mma.kind::f16 (M = 128, K = 16) really take in the operand / accumulator patterns the split-fp16 kernels use?
// One thread per CTA issues `iters` groups of MMAs on resident (zero) shared-memory operands - no TMA, no epilogue - and
// times them with clock64 between the first issue and the completion of a tcgen05.commit.
#include <vector>

#include "umma_ptx.cuh"

namespace gimb {
namespace {

__global__ void __launch_bounds__(288, 1) mma_probe_kernel(int variant, int n, int iters, unsigned idesc, int ld_warps, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - raw);
  const int a_plane = 128 * 128, b_plane = n * 128;          // k = 64 per stage, 128-byte rows (SWIZZLE_128B)
  const int stage_bytes = 2 * a_plane + 2 * b_plane;
  const uint32_t bars = base + 2 * stage_bytes;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gen + 2 * stage_bytes + 64);
  for (int i = threadIdx.x; i < 2 * stage_bytes / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(gen)[i] = 0u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    if (lane == 0) { mbar_init(bars, 1); fence_barrier_init(); }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  volatile int* done_flag = reinterpret_cast<volatile int*>(gen + 2 * stage_bytes + 128);
  if (threadIdx.x == 0) *done_flag = 0;
  __syncthreads();
  if (warp >= 1 && warp <= ld_warps) {
    // concurrent epilogue-like TMEM reads (columns 384..511 of this warp's lane quadrant) while the MMAs run
    long long loads = 0;
    uint32_t r[32];
    float sink = 0.f;
    while (*done_flag == 0) {
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        tmem_ld32(tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + 384 + 32 * c, r);
        tmem_ld_wait();
        sink += __uint_as_float(r[0]) + __uint_as_float(r[31]);
      }
      loads += 4;
    }
    if (lane == 0) { out[blockIdx.x * 4 + 2] += loads; if (sink == 123.f) out[0] = 1; }
  }
  if (threadIdx.x == 0) {
    const uint64_t dconst = make_desc_sw128(0u);
    const uint32_t tX = tmem_base, tY = tmem_base + 256;
    long long count = 0;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const int st = it & 1;
      const uint64_t dA_hi = dconst + (((base + st * stage_bytes) & 0x3FFFFu) >> 4);
      const uint64_t dA_lo = dA_hi + (a_plane >> 4);
      const uint64_t dB_hi = dA_hi + (a_plane >> 3);
      const uint64_t dB_lo = dB_hi + (b_plane >> 4);
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const uint32_t acc = (it == 0 && kk == 0) ? 0u : 1u;
        if (variant == 0) {          // same operands, one accumulator
          umma_f16(tX, dA_hi, dB_hi, idesc, acc);
          umma_f16(tX, dA_hi, dB_hi, idesc, 1u);
          umma_f16(tX, dA_hi, dB_hi, idesc, 1u);
        } else if (variant == 1) {   // walk k, one accumulator, hi*hi only
          umma_f16(tX, dA_hi + 2 * kk, dB_hi + 2 * kk, idesc, acc);
          umma_f16(tX, dA_hi + 2 * kk, dB_hi + 2 * kk, idesc, 1u);
          umma_f16(tX, dA_hi + 2 * kk, dB_hi + 2 * kk, idesc, 1u);
        } else if (variant == 2) {   // split scheme, two accumulators interleaved (corr_sweep.cu)
          umma_f16(tX, dA_hi + 2 * kk, dB_lo + 2 * kk, idesc, acc);
          umma_f16(tX, dA_lo + 2 * kk, dB_hi + 2 * kk, idesc, 1u);
          umma_f16(tY, dA_hi + 2 * kk, dB_hi + 2 * kk, idesc, acc);
        } else if (variant == 3) {   // split scheme, one accumulator
          umma_f16(tX, dA_hi + 2 * kk, dB_lo + 2 * kk, idesc, acc);
          umma_f16(tX, dA_lo + 2 * kk, dB_hi + 2 * kk, idesc, 1u);
          umma_f16(tX, dA_hi + 2 * kk, dB_hi + 2 * kk, idesc, 1u);
        } else {                     // 4: three independent accumulators
          umma_f16(tX, dA_hi + 2 * kk, dB_lo + 2 * kk, idesc, acc);
          umma_f16(tX + 128, dA_lo + 2 * kk, dB_hi + 2 * kk, idesc, acc);
          umma_f16(tY, dA_hi + 2 * kk, dB_hi + 2 * kk, idesc, acc);
        }
        count += 3;
      }
    }
    umma_commit(bars);
    mbar_wait(bars, 0);
    const long long t1 = clock64();
    out[blockIdx.x * 4] = t1 - t0;
    out[blockIdx.x * 4 + 1] = count;
    *done_flag = 1;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
}

}  // namespace
}  // namespace gimb

extern "C" int gimb_probe_mma(int variant, int n, int iters, int grid, int ld_warps, float* cycles_per_mma, float* ms_out, float* ld_bytes_per_clk, void* stream) {
  using namespace gimb;
  GIMB_CHECK(cycles_per_mma && ms_out && (n == 64 || n == 128 || n == 256) && grid > 0 && iters > 0, "gimb_probe_mma: bad arguments");
  GIMB_CHECK(!(variant == 4 && n > 128) && !(variant == 2 && n > 256), "gimb_probe_mma: accumulators do not fit");
  cudaStream_t st = (cudaStream_t)stream;
  const int smem = 2 * (2 * 128 * 128 + 2 * n * 128) + 1024 + 256;
  GIMB_CUDA(cudaFuncSetAttribute(mma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  long long* d = nullptr;
  GIMB_CUDA(cudaMalloc(&d, (size_t)grid * 4 * sizeof(long long)));
  const unsigned idesc = (1u << 4) | ((unsigned)(n >> 3) << 17) | ((unsigned)(128 >> 4) << 24);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  mma_probe_kernel<<<grid, 288, smem, st>>>(variant, n, iters, idesc, ld_warps, d);
  GIMB_CUDA(cudaMemsetAsync(d, 0, (size_t)grid * 4 * sizeof(long long), st));
  cudaEventRecord(e0, st);
  mma_probe_kernel<<<grid, 288, smem, st>>>(variant, n, iters, idesc, ld_warps, d);
  cudaEventRecord(e1, st);
  GIMB_CUDA(cudaEventSynchronize(e1));
  cudaEventElapsedTime(ms_out, e0, e1);
  std::vector<long long> h((size_t)grid * 4);
  GIMB_CUDA(cudaMemcpy(h.data(), d, h.size() * sizeof(long long), cudaMemcpyDeviceToHost));
  double acc = 0;
  double ldb = 0;
  for (int g = 0; g < grid; ++g) {
    acc += (double)h[g * 4] / (double)h[g * 4 + 1];
    ldb += (double)h[g * 4 + 2] * 4096.0 / (double)h[g * 4];  // every load moves 32 lanes x 32 columns x 4 B
  }
  *cycles_per_mma = (float)(acc / grid);
  if (ld_bytes_per_clk) *ld_bytes_per_clk = (float)(ldb / grid);
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(d);
  return 0;
}
