// split.cuh - device helpers for the split-fp16 activation format (see umma_gemm.cuh).
#pragma once
#include <cuda_fp16.h>

#include "umma_gemm.cuh"

namespace gimb {

// raw device view of SplitPlanes for kernels
struct PlanesDev {
  __half* hi;
  __half* lo;
  __half* h8;
  int ld;
};
static inline PlanesDev dev(const SplitPlanes& s) { return PlanesDev{s.hi, s.lo, s.h8, s.ld}; }

// store 4 consecutive values (16-byte aligned group of 4 floats -> 8-byte groups of halves)
__device__ __forceinline__ void split4_store(const PlanesDev& p, size_t off, float a, float b, float c, float d) {
  const __half h0 = __float2half_rn(a), h1 = __float2half_rn(b), h2 = __float2half_rn(c), h3 = __float2half_rn(d);
  const float f0 = __half2float(h0), f1 = __half2float(h1), f2 = __half2float(h2), f3 = __half2float(h3);
  __half2 hi01 = __halves2half2(h0, h1), hi23 = __halves2half2(h2, h3);
  __half2 lo01 = __floats2half2_rn((a - f0) * kSplitScale, (b - f1) * kSplitScale);
  __half2 lo23 = __floats2half2_rn((c - f2) * kSplitScale, (d - f3) * kSplitScale);
  uint2 uh, ul;
  uh.x = *reinterpret_cast<unsigned int*>(&hi01); uh.y = *reinterpret_cast<unsigned int*>(&hi23);
  ul.x = *reinterpret_cast<unsigned int*>(&lo01); ul.y = *reinterpret_cast<unsigned int*>(&lo23);
  *reinterpret_cast<uint2*>(p.hi + off) = uh;
  *reinterpret_cast<uint2*>(p.lo + off) = ul;
  if (p.h8) {
    __half2 s01 = __floats2half2_rn(f0 * kSplitScale, f1 * kSplitScale);
    __half2 s23 = __floats2half2_rn(f2 * kSplitScale, f3 * kSplitScale);
    uint2 us;
    us.x = *reinterpret_cast<unsigned int*>(&s01); us.y = *reinterpret_cast<unsigned int*>(&s23);
    *reinterpret_cast<uint2*>(p.h8 + off) = us;
  }
}

}  // namespace gimb
