// test_hooks.cu (libgimb200_test.so only, include/gimb200_test.h) - C entry points used only by tests/ and tools/: run one GEMM-shaped layer through the tcgen05 kernel
// and through the fp32 CUDA-core kernel on the same device buffers, so the two can be compared with a
// float64 reference on the host side.
#include <vector>

#include "../../include/gimb200_test.h"
#include "ops.cuh"
#include "umma_gemm.cuh"

using namespace gimb;

namespace {
__global__ void planes_to_f32_kernel(const __half* hi, const __half* lo, long long rows, int cols, int ld, float* out) {
  long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= rows * cols) return;
  long long r = idx / cols;
  int c = (int)(idx - r * cols);
  out[idx] = __half2float(hi[r * ld + c]) + __half2float(lo[r * ld + c]) * (1.f / kSplitScale);
}
}  // namespace

extern "C" int gimb_test_conv(const float* in, const float* in2, int B, int H, int W, int C1, int C2, const float* w, int Cout,
                              int ksize, int stride, const float* scale, const float* bias, const float* residual,
                              const uint8_t* row_mask, int act0, int act1, int act_split, float div, float* out_umma,
                              float* out_umma_planes, float* out_simt, void* workspace, size_t workspace_bytes,
                              void* stream) {
  GIMB_CHECK(in && w && out_umma && out_simt && workspace, "gimb_test_conv: null argument");
  Ctx ctx;
  ctx.stream = (cudaStream_t)stream;
  int dev = 0;
  GIMB_CUDA(cudaGetDevice(&dev));
  GIMB_CUDA(cudaDeviceGetAttribute(&ctx.sm_count, cudaDevAttrMultiProcessorCount, dev));
  uintptr_t base = ((uintptr_t)workspace + Arena::kAlign - 1) / Arena::kAlign * Arena::kAlign;
  ctx.arena.base = (char*)base;
  ctx.arena.cap = workspace_bytes - (base - (uintptr_t)workspace);
  const int pad = ksize / 2;
  const int OH = (H + 2 * pad - ksize) / stride + 1, OW = (W + 2 * pad - ksize) / stride + 1;
  const long long M = (long long)B * OH * OW;
  const int Cin = C1 + C2;

  // ---- fp32 CUDA-core path
  ConvGemm c;
  c.in = in; c.in2 = in2; c.B = B; c.H = H; c.W = W; c.C1 = C1; c.C2 = C2;
  c.KH = c.KW = ksize; c.stride = stride; c.pad = pad; c.OH = OH; c.OW = OW;
  c.w = w; c.Cout = Cout; c.scale = scale; c.bias = bias; c.residual = residual; c.row_mask = row_mask;
  c.act0 = act0; c.act1 = act1; c.act_split = act_split; c.div = div; c.out = out_simt;
  GIMB_TRY(conv_gemm(ctx, c));

  // ---- tcgen05 path: split the operands, run, optionally re-assemble the output planes
  auto pitch8 = [](int v) { return (v + 7) / 8 * 8; };
  Arena& A = ctx.arena;
  SplitPlanes a, a2, b, o;
  const long long pix = (long long)B * H * W;
  a.ld = pitch8(C1);
  a.hi = A.alloc<__half>(pix * a.ld); a.lo = A.alloc<__half>(pix * a.ld);
  GIMB_TRY(split_planes(ctx, in, pix, C1, C1, a));
  if (in2) {
    a2.ld = pitch8(C2);
    a2.hi = A.alloc<__half>(pix * a2.ld); a2.lo = A.alloc<__half>(pix * a2.ld);
    GIMB_TRY(split_planes(ctx, in2, pix, C2, C2, a2));
  }
  // weights [Cout][taps][Cin] -> planes with per-tap pitch ldk
  const int taps = ksize * ksize;
  const int ldk = pitch8(Cin);
  b.ld = taps * ldk;
  b.hi = A.alloc<__half>((size_t)Cout * b.ld); b.lo = A.alloc<__half>((size_t)Cout * b.ld);
  {
    SplitPlanes bt = b;
    bt.ld = ldk;  // treat [Cout*taps] rows of Cin -> pitch ldk
    GIMB_TRY(split_planes(ctx, w, (long long)Cout * taps, Cin, Cin, bt));
  }
  o.ld = pitch8(Cout);
  if (out_umma_planes) { o.hi = A.alloc<__half>(M * o.ld); o.lo = A.alloc<__half>(M * o.ld); }
  GIMB_CHECK(!A.overflow, "gimb_test_conv: workspace too small");

  UmmaGemm g;
  g.a = a; g.a2 = a2; g.b = b; g.N = Cout;
  if (ksize == 1 && stride == 1) {
    g.mode = 0; g.M = M; g.K1 = C1; g.K2 = C2;
  } else {
    g.mode = 1; g.K1 = Cin; g.B = B; g.H = H; g.W = W; g.KH = g.KW = ksize; g.stride = stride; g.pad = pad;
    g.OH = OH; g.OW = OW; g.ldk = ldk;
  }
  g.scale = scale; g.bias = bias; g.residual = residual; g.row_mask = row_mask;
  g.act0 = act0; g.act1 = act1; g.act_split = act_split; g.div = div;
  g.out_f32 = out_umma; g.out = o;
  GIMB_TRY(umma_gemm(ctx, g));
  if (out_umma_planes) {
    planes_to_f32_kernel<<<(unsigned)cdiv64(M * Cout, 256), 256, 0, ctx.stream>>>(o.hi, o.lo, M, Cout, o.ld, out_umma_planes);
    GIMB_LAUNCH_CHECK();
  }
  return 0;
}

// Times the tcgen05 GEMM alone on one layer shape (operands pre-split, buffers allocated here, zero data except a
// constant fill): `iters` back-to-back launches between two CUDA events on `stream`.  flags: 1 = folded BN,
// 2 = fp32 residual, 4 = fp32 output, 8 = fp16-plane output, 16 = residual as fp16 planes.
extern "C" int gimb_bench_layer(int B, int H, int W, int C1, int C2, int Cout, int ksize, int stride, int flags, int act,
                                int iters, float* ms_out, void* stream) {
  GIMB_CHECK(ms_out && iters > 0, "gimb_bench_layer: bad argument");
  Ctx ctx;
  ctx.stream = (cudaStream_t)stream;
  int dev = 0;
  GIMB_CUDA(cudaGetDevice(&dev));
  GIMB_CUDA(cudaDeviceGetAttribute(&ctx.sm_count, cudaDevAttrMultiProcessorCount, dev));
  const int pad = ksize / 2;
  const int OH = (H + 2 * pad - ksize) / stride + 1, OW = (W + 2 * pad - ksize) / stride + 1;
  const long long M = (long long)B * OH * OW, pix = (long long)B * H * W;
  const int Cin = C1 + C2;
  auto pitch8 = [](int v) { return (v + 7) / 8 * 8; };
  std::vector<void*> bufs;
  auto dalloc = [&](size_t bytes, int fill) -> void* {
    void* p = nullptr;
    if (cudaMalloc(&p, bytes) != cudaSuccess) return nullptr;
    cudaMemset(p, fill, bytes);
    bufs.push_back(p);
    return p;
  };
  SplitPlanes a, a2, b, o;
  a.ld = pitch8(C1);
  a.hi = (__half*)dalloc(pix * a.ld * 2, 0x3c); a.lo = (__half*)dalloc(pix * a.ld * 2, 0);
  if (C2) { a2.ld = pitch8(C2); a2.hi = (__half*)dalloc(pix * a2.ld * 2, 0x3c); a2.lo = (__half*)dalloc(pix * a2.ld * 2, 0); }
  const int taps = ksize * ksize, ldk = pitch8(Cin);
  b.ld = taps * ldk;
  b.hi = (__half*)dalloc((size_t)Cout * b.ld * 2, 0x1c); b.lo = (__half*)dalloc((size_t)Cout * b.ld * 2, 0);
  float* scale = (flags & 1) ? (float*)dalloc(Cout * 4, 0) : nullptr;
  float* bias = (flags & 1) ? (float*)dalloc(Cout * 4, 0) : nullptr;
  float* res = (flags & 2) ? (float*)dalloc(M * Cout * 4, 0) : nullptr;
  float* of = (flags & 4) ? (float*)dalloc(M * Cout * 4, 0) : nullptr;
  if (flags & 8) { o.ld = pitch8(Cout); o.hi = (__half*)dalloc(M * o.ld * 2, 0); o.lo = (__half*)dalloc(M * o.ld * 2, 0); }
  SplitPlanes rp;
  if (flags & 16) { rp.ld = pitch8(Cout); rp.hi = (__half*)dalloc(M * rp.ld * 2, 0x3c); rp.lo = (__half*)dalloc(M * rp.ld * 2, 0); }
  int rc = 0;
  for (void* p : bufs) if (!p) rc = 1;
  if (rc) { for (void* p : bufs) if (p) cudaFree(p); set_error("gimb_bench_layer: out of memory"); return 1; }
  UmmaGemm g;
  g.a = a; g.a2 = a2; g.b = b; g.N = Cout;
  if (ksize == 1 && stride == 1) { g.mode = 0; g.M = M; g.K1 = C1; g.K2 = C2; }
  else { g.mode = 1; g.K1 = Cin; g.B = B; g.H = H; g.W = W; g.KH = g.KW = ksize; g.stride = stride; g.pad = pad; g.OH = OH; g.OW = OW; g.ldk = ldk; }
  g.scale = scale; g.bias = bias; g.residual = res; g.act0 = g.act1 = act; g.out_f32 = of; g.out = o;
  g.residual_planes = rp;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  rc = umma_gemm(ctx, g);  // warm-up
  if (!rc) {
    cudaEventRecord(e0, ctx.stream);
    for (int i = 0; i < iters && !rc; ++i) rc = umma_gemm(ctx, g);
    cudaEventRecord(e1, ctx.stream);
    if (cudaEventSynchronize(e1) != cudaSuccess) { set_error("gimb_bench_layer: %s", cudaGetErrorString(cudaGetLastError())); rc = 1; }
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    *ms_out = ms / iters;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  for (void* p : bufs) cudaFree(p);
  return rc;
}

extern "C" int gimb_probe_tma(int variant, int iters, float* gbps_out, void* stream) {
  GIMB_CHECK(gbps_out, "gimb_probe_tma: null argument");
  Ctx ctx;
  ctx.stream = (cudaStream_t)stream;
  int dev = 0;
  GIMB_CUDA(cudaGetDevice(&dev));
  GIMB_CUDA(cudaDeviceGetAttribute(&ctx.sm_count, cudaDevAttrMultiProcessorCount, dev));
  return tma_probe(ctx, variant, iters, gbps_out);
}
