// fine.cu - fine-level window crop and sub-pixel matching.
//   fine_gather : FinePreprocess.forward (networks/loftr/submodules/fine_preprocess.py:29-47) without the
//                 F.unfold im2col buffer: windows are read straight from the NHWC fine map.
//   fine_match  : FineMatching.forward / get_fine_match (networks/loftr/utils/fine_matching.py:43-72).
#include <algorithm>

#include "ops.cuh"
#include "split.cuh"

namespace gimb {
namespace {

// one CTA per (match, token-group); each thread moves one float4 of one token.
__global__ void __launch_bounds__(256) fine_gather_kernel(const float* __restrict__ feat, int hf, int wf, int C4,
                                                          int wc, int stride, int Wn, const int64_t* __restrict__ b_ids,
                                                          const int64_t* __restrict__ ids, int64_t m0, int64_t m,
                                                          float* __restrict__ out, const PlanesDev sp) {
  const int WW = Wn * Wn, pad = Wn / 2;
  const int per_match = WW * C4;
  for (int64_t mm = blockIdx.x; mm < m; mm += gridDim.x) {
    const int64_t g = m0 + mm;
    const int b = (int)b_ids[g];
    const int cell = (int)ids[g];
    const int cy = (cell / wc) * stride - pad, cx = (cell % wc) * stride - pad;
    const float4* base = reinterpret_cast<const float4*>(feat) + (size_t)b * hf * wf * C4;
    float4* o = reinterpret_cast<float4*>(out) + (size_t)mm * per_match;
    for (int e = threadIdx.x; e < per_match; e += blockDim.x) {
      int t = e / C4, c = e - t * C4;
      int y = cy + t / Wn, x = cx + t % Wn;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (y >= 0 && y < hf && x >= 0 && x < wf) v = base[((size_t)y * wf + x) * C4 + c];
      o[e] = v;
      if (sp.hi) split4_store(sp, ((size_t)mm * WW + t) * sp.ld + c * 4, v.x, v.y, v.z, v.w);
    }
  }
}

// one warp per match: heat = softmax(<f0[centre], f1[r]> / sqrt(C)) over the WW tokens, expectation on the
// [-1,1]^2 grid (kornia create_meshgrid / spatial_expectation2d semantics), std, final coordinates.
__global__ void __launch_bounds__(256) fine_match_kernel(const FineMatchArgs a) {
  const int lane = threadIdx.x & 31;
  const int64_t mm = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (mm >= a.m) return;
  const int64_t g = a.m0 + mm;
  const int C = a.C, WW = a.WW, Wn = a.Wn;
  const float* f0 = a.f0 + ((size_t)mm * WW + WW / 2) * C;
  const float* f1 = a.f1 + (size_t)mm * WW * C;
  float sim = -INFINITY;
  if (lane < WW) {
    float s = 0.f;
    const float4* p0 = reinterpret_cast<const float4*>(f0);
    const float4* p1 = reinterpret_cast<const float4*>(f1 + (size_t)lane * C);
    for (int c = 0; c < C / 4; ++c) {
      float4 x = p0[c], y = p1[c];
      s = fmaf(x.x, y.x, s); s = fmaf(x.y, y.y, s); s = fmaf(x.z, y.z, s); s = fmaf(x.w, y.w, s);
    }
    sim = a.sim_scale * s;
  }
  float mx = sim;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  float e = lane < WW ? expf(sim - mx) : 0.f;
  float sum = e;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  float heat = __fdiv_rn(e, sum);
  // grid = linspace(-1, 1, Wn): x varies fastest
  float gx = 0.f, gy = 0.f;
  if (lane < WW) {
    gx = ((float)(lane % Wn) / (float)(Wn - 1) - 0.5f) * 2.f;
    gy = ((float)(lane / Wn) / (float)(Wn - 1) - 0.5f) * 2.f;
  }
  float ex = heat * gx, ey = heat * gy, exx = heat * gx * gx, eyy = heat * gy * gy;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    ex += __shfl_xor_sync(0xffffffffu, ex, o);
    ey += __shfl_xor_sync(0xffffffffu, ey, o);
    exx += __shfl_xor_sync(0xffffffffu, exx, o);
    eyy += __shfl_xor_sync(0xffffffffu, eyy, o);
  }
  if (lane == 0) {
    float vx = exx - ex * ex, vy = eyy - ey * ey;
    float sd = sqrtf(fmaxf(vx, 1e-10f)) + sqrtf(fmaxf(vy, 1e-10f));
    a.expec_f[g * 3 + 0] = ex;
    a.expec_f[g * 3 + 1] = ey;
    a.expec_f[g * 3 + 2] = sd;
    float sx = a.fscale, sy = a.fscale;
    if (a.scale1) {
      int b = (int)a.b_ids[g];
      sx = a.fscale * a.scale1[b * 2];
      sy = a.fscale * a.scale1[b * 2 + 1];
    }
    float half = (float)(Wn / 2);
    a.mkpts0_f[g * 2 + 0] = a.mkpts0_c[g * 2 + 0];
    a.mkpts0_f[g * 2 + 1] = a.mkpts0_c[g * 2 + 1];
    a.mkpts1_f[g * 2 + 0] = a.mkpts1_c[g * 2 + 0] + ex * half * sx;
    a.mkpts1_f[g * 2 + 1] = a.mkpts1_c[g * 2 + 1] + ey * half * sy;
  }
}

}  // namespace

int fine_gather(Ctx& ctx, const float* feat_f, int hf, int wf, int C, int wc, int stride, int Wn,
                const int64_t* b_ids, const int64_t* ids, int64_t m0, int64_t m, float* out, const SplitPlanes* planes) {
  GIMB_CHECK(C % 4 == 0, "fine_gather: C %% 4 != 0");
  if (ctx.dry || m == 0) return 0;
  const PlanesDev sp = planes ? dev(*planes) : PlanesDev{nullptr, nullptr, nullptr, 0};
  int blocks = (int)std::min<int64_t>(m, (int64_t)ctx.sm_count * 32);
  fine_gather_kernel<<<blocks, 256, 0, ctx.stream>>>(feat_f, hf, wf, C / 4, wc, stride, Wn, b_ids, ids, m0, m, out, sp);
  ctx.launches++;
  GIMB_LAUNCH_CHECK();
  return 0;
}

int fine_match(Ctx& ctx, const FineMatchArgs& a) {
  GIMB_CHECK(a.WW <= 32 && a.C % 4 == 0, "fine_match: WW <= 32 and C %% 4 == 0 required");
  if (ctx.dry || a.m == 0) return 0;
  fine_match_kernel<<<(unsigned)cdiv64(a.m, 8), 256, 0, ctx.stream>>>(a);
  ctx.launches++;
  GIMB_LAUNCH_CHECK();
  return 0;
}

}  // namespace gimb
