// dkm_api.cu - the gim_dkm (DKMv3) part of the C ABI (include/gimb200.h): weight blob -> device model and the
// orchestration of RegressionMatcher.match (networks/dkm/models/dkm.py:655-752) with symmetric = True, batched = False.
//
//   images -> bilinear resize (h, w) -> ResNet-50 pyramid of cat(query, support) (encoders.py:46-62)
//          -> Decoder (dkm.py:454-534): scales 32, 16: 1x1 projection, GP regression on the cosine kernel (the "global 4-D
//             correlation"), DFN embedding decoder; scales 16..1: ConvRefiner with local correlation
//          -> second pass at upsample_res (scales 8..1 only: layers 3-4 of the encoder are not needed and skipped)
//          -> certainty attenuation, sigmoid, out-of-range / black-pixel masks, symmetric concat (dkm.py:684-752)
//
// Every convolution / Linear-shaped layer runs on the shared tcgen05 split-fp16 GEMM engine (engine.cuh, umma_gemm.cu):
// the ResNet trunk, the projections, the DFN's 1x1 / 3x3 layers, the ConvRefiner's pointwise C x C GEMMs (83 % of the
// FLOPs) and the GP Gram matrices.  Everything else is in dkm_kernels.cu.
#include <math.h>
#include <stdlib.h>

#include <algorithm>
#include <string>
#include <vector>

#include "dkm_ops.cuh"
#include "engine.cuh"

namespace gimb {
const char* last_error();
namespace {

struct DBottleneck {
  Conv c1, c2, c3, ds;
  bool has_ds = false;
  int stride = 1;
};
struct RRBw { Conv c1, c2, c3; };
struct Refiner {
  int cin = 0, hidden = 0, emb = 0, radius = 0, mult = 1, feat = 0;
  const float *emb_w = nullptr, *emb_b = nullptr;
  const float* dw_wt[9];  // the same transposed and padded [25][pitch8(C)] (+ padded scale / bias) for the vectorised kernel
  const float* dw_sp[9];
  const float* dw_bp[9];
  const float* dw_w[9];   // depthwise weights [C, 25]
  const float* dw_s[9];   // folded BatchNorm (+ conv bias)
  const float* dw_b[9];
  Conv pw[9];             // pointwise C x C + bias
  Conv out;               // C -> 3 (padded to 8 outputs)
};

}  // namespace
}  // namespace gimb

using namespace gimb;

struct gimb_dkm {
  int device = 0;
  int engine = ENGINE_TC;
  int sm_count = 148;
  WeightStore ws;
  Conv stem;
  std::vector<DBottleneck> layers[4];
  Conv proj[2];                  // [0] = scale 32, [1] = scale 16
  const float *pos_w[2], *pos_b[2];
  Conv dfn_feat[2], dfn_term[2];
  RRBw rrb_d[2], rrb_u[2];
  const float *cab_w1[2], *cab_b1[2], *cab_w2[2], *cab_b2[2];
  Refiner ref[5];                // scales 16, 8, 4, 2, 1
  uint64_t launches = 0;
};

namespace gimb {
namespace {

constexpr int GP_DIM = 256, DFN_DIM = 384, FEAT_DIM = 256;

int load_rrb(gimb_dkm* m, Ctx& ctx, const std::string& pre, RRBw* r) {
  GIMB_TRY(m->ws.load_conv(ctx, pre + ".c1", true, &r->c1));
  GIMB_TRY(m->ws.load_conv(ctx, pre + ".c2", true, &r->c2));
  GIMB_TRY(m->ws.load_conv(ctx, pre + ".c3", true, &r->c3));
  return 0;
}

int build_dkm(gimb_dkm* m, Ctx& ctx) {
  for (int li = 0; li < 4; ++li) m->layers[li].clear();
  GIMB_TRY(m->ws.load_conv(ctx, "enc.stem", true, &m->stem));
  const int nblk[4] = {3, 4, 6, 3};
  for (int li = 0; li < 4; ++li)
    for (int bi = 0; bi < nblk[li]; ++bi) {
      DBottleneck b;
      const std::string pre = "enc.l" + std::to_string(li + 1) + "." + std::to_string(bi);
      GIMB_TRY(m->ws.load_conv(ctx, pre + ".c1", true, &b.c1));
      GIMB_TRY(m->ws.load_conv(ctx, pre + ".c2", true, &b.c2));
      GIMB_TRY(m->ws.load_conv(ctx, pre + ".c3", true, &b.c3));
      b.has_ds = bi == 0;
      b.stride = (li > 0 && bi == 0) ? 2 : 1;
      if (b.has_ds) GIMB_TRY(m->ws.load_conv(ctx, pre + ".ds", true, &b.ds));
      m->layers[li].push_back(b);
    }
  const char* sc[2] = {"32", "16"};
  for (int i = 0; i < 2; ++i) {
    const std::string s = sc[i];
    GIMB_TRY(m->ws.load_conv(ctx, "proj." + s, true, &m->proj[i]));
    GIMB_TRY(m->ws.find("gp." + s + ".pos_w", &m->pos_w[i]));
    GIMB_TRY(m->ws.find("gp." + s + ".pos_b", &m->pos_b[i]));
    GIMB_TRY(m->ws.load_conv(ctx, "dfn." + s + ".feat", true, &m->dfn_feat[i]));
    GIMB_TRY(load_rrb(m, ctx, "dfn." + s + ".rrbd", &m->rrb_d[i]));
    GIMB_TRY(load_rrb(m, ctx, "dfn." + s + ".rrbu", &m->rrb_u[i]));
    GIMB_TRY(m->ws.find("dfn." + s + ".cab.w1", &m->cab_w1[i]));
    GIMB_TRY(m->ws.find("dfn." + s + ".cab.b1", &m->cab_b1[i]));
    GIMB_TRY(m->ws.find("dfn." + s + ".cab.w2", &m->cab_w2[i]));
    GIMB_TRY(m->ws.find("dfn." + s + ".cab.b2", &m->cab_b2[i]));
    GIMB_TRY(m->ws.load_conv(ctx, "dfn." + s + ".term", true, &m->dfn_term[i]));
  }
  const char* rs[5] = {"16", "8", "4", "2", "1"};
  const int rad[5] = {7, 3, 2, 0, 0}, feat[5] = {512, 512, 256, 64, 3};
  for (int i = 0; i < 5; ++i) {
    Refiner& r = m->ref[i];
    const std::string pre = std::string("ref.") + rs[i];
    std::vector<uint32_t> sh;
    GIMB_TRY(m->ws.find(pre + ".emb_w", &r.emb_w, &sh));
    GIMB_TRY(m->ws.find(pre + ".emb_b", &r.emb_b));
    GIMB_CHECK(!sh.empty() && sh[0] > 0, "weight blob: %s.emb_w has no shape", pre.c_str());
    r.emb = (int)sh[0];
    r.radius = rad[i]; r.feat = feat[i];
    r.cin = 2 * r.feat + r.emb + (r.radius ? (2 * r.radius + 1) * (2 * r.radius + 1) : 0);
    for (int k = 0; k < 9; ++k) {
      const std::string bp = pre + ".b" + std::to_string(k);
      std::vector<uint32_t> dsh;
      GIMB_TRY(m->ws.find(bp + ".dw_w", &r.dw_w[k], &dsh));
      GIMB_TRY(m->ws.find(bp + ".dw_s", &r.dw_s[k]));
      GIMB_TRY(m->ws.find(bp + ".dw_b", &r.dw_b[k]));
      GIMB_TRY(m->ws.find(bp + ".dw_wt", &r.dw_wt[k]));
      GIMB_TRY(m->ws.find(bp + ".dw_sp", &r.dw_sp[k]));
      GIMB_TRY(m->ws.find(bp + ".dw_bp", &r.dw_bp[k]));
      if (k == 0) {
        GIMB_CHECK(!dsh.empty() && r.cin > 0, "weight blob: %s.dw_w has no shape", bp.c_str());
        r.hidden = (int)dsh[0];
        r.mult = r.hidden / r.cin;
      }
      GIMB_TRY(m->ws.load_conv(ctx, bp + ".pw", true, &r.pw[k]));
    }
    GIMB_CHECK(r.hidden == r.cin * r.mult && r.pw[0].cin == r.hidden, "refiner %s: inconsistent channel counts", rs[i]);
    GIMB_TRY(m->ws.load_conv(ctx, pre + ".out", true, &r.out));
  }
  return 0;
}

// fp32 NHWC [rows, C] with pitch pitch8(C) -> ActT view (+ planes when given)
ActT act_f32(float* p, int C, int ld) {
  ActT a;
  a.f32 = p; a.C = C; a.ldf = ld == C ? 0 : ld;
  return a;
}

// input planes of a GEMM layer from an fp32 tensor (TC engine); the CUDA-core engine reads the fp32 tensor directly
int with_planes(Fwd& F, ActT& a, size_t rows) {
  if (!F.tc() || a.sp.hi) return 0;
  a.sp.ld = pitch8(a.C);
  a.sp.hi = F.ctx.arena.alloc<__half>(rows * a.sp.ld);
  a.sp.lo = F.ctx.arena.alloc<__half>(rows * a.sp.ld);
  GIMB_CHECK(F.ctx.dry || !F.ctx.arena.overflow, "dkm: workspace exhausted (planes)");
  return split_planes(F.ctx, a.f32, (int64_t)rows, a.C, a.pitch(), a.sp);
}

struct Pyramid {   // fp32 NHWC features of cat(query, support), batch 2
  float* f[6];     // index = log2(scale): f[0] = image (NHWC, C = 3), f[1] = 64 ch @1/2, ... f[5] = 2048 ch @1/32
  int C[6];
  int H[6], W[6];
};

// ResNet50.forward (networks/dkm/models/encoders.py:46-62) on a batch of 2 images (NCHW in); upto = 8 skips layers 3, 4
int encoder(Fwd& F, gimb_dkm* m, const float* nchw, int B, int H, int W, int upto, Pyramid* py) {
  Ctx& ctx = F.ctx;
  Arena& A = ctx.arena;
  const int chans[6] = {3, 64, 256, 512, 1024, 2048};
  // sizes like torchvision's ResNet: every stride-2 stage (7x7 stem, max-pool, 3x3 convs with padding 1) maps n -> ceil(n / 2);
  // only the stem kernel needs an even input (the ZEB harness runs 660 x 880: 330 -> 165 -> 83 -> 42 -> 21)
  GIMB_CHECK(H % 2 == 0 && W % 2 == 0, "dkm encoder: the resized image must have even height and width (got %dx%d)", H, W);
  for (int i = 0; i < 6; ++i) {
    py->C[i] = chans[i]; py->f[i] = nullptr;
    py->H[i] = i == 0 ? H : (py->H[i - 1] + 1) / 2;
    py->W[i] = i == 0 ? W : (py->W[i - 1] + 1) / 2;
  }
  const int top = upto == 8 ? 3 : 5;
  for (int i = 1; i <= top; ++i) py->f[i] = A.alloc<float>((size_t)B * py->H[i] * py->W[i] * chans[i]);
  py->f[0] = A.alloc<float>((size_t)B * H * W * 4);  // image as NHWC with pitch 4 (refiner "1")
  GIMB_CHECK(ctx.dry || !A.overflow, "dkm encoder: workspace exhausted");
  size_t mark = A.mark();
  // stem: 7x7 s2 + BN + ReLU (same kernel as the gim_loftr stem), fp32 out = feats[2]
  GIMB_TRY(stem_conv7x7(ctx, nchw, B, H, W, m->stem.wt.w, m->stem.s, m->stem.b, py->f[1], nullptr));
  ActT cur = F.alloc((size_t)B * py->H[2] * py->W[2], 64, false, true);
  GIMB_CHECK(ctx.dry || !A.overflow, "dkm encoder: workspace exhausted");
  GIMB_TRY(dkm_maxpool3x3s2(ctx, py->f[1], B, py->H[1], py->W[1], 64, nullptr, cur.planes()));
  int cH = py->H[2], cW = py->W[2];
  for (int li = 0; li < top - 1; ++li) {
    for (size_t bi = 0; bi < m->layers[li].size(); ++bi) {
      const DBottleneck& b = m->layers[li][bi];
      const int oH = b.stride == 2 ? (cH + 1) / 2 : cH, oW = b.stride == 2 ? (cW + 1) / 2 : cW;
      const bool last = bi + 1 == m->layers[li].size();
      // block output: planes (next GEMM operand + identity); the last block of a layer also as fp32 (pyramid level)
      ActT xo = F.alloc((size_t)B * oH * oW, b.c3.cout, false, true);
      if (last) xo.f32 = py->f[li + 2];
      size_t mk2 = A.mark();
      ActT u1 = F.alloc((size_t)B * cH * cW, b.c1.cout, false, true);
      ActT u2 = F.alloc((size_t)B * oH * oW, b.c2.cout, false, true);
      ActT idn = cur;
      if (b.has_ds) idn = F.alloc((size_t)B * oH * oW, b.c3.cout, false, true);
      GIMB_CHECK(ctx.dry || !A.overflow, "dkm encoder: workspace exhausted");
      GIMB_TRY(run_conv(F, b.c1, cur, B, cH, cW, 1, ACT_RELU, nullptr, u1));
      GIMB_TRY(run_conv(F, b.c2, u1, B, cH, cW, b.stride, ACT_RELU, nullptr, u2));
      if (b.has_ds) GIMB_TRY(run_conv(F, b.ds, cur, B, cH, cW, b.stride, ACT_NONE, nullptr, idn));
      if (last) {
        // fp32 + planes with the identity as planes is not an epilogue variant: planes first, then the fp32 copy
        ActT xp = xo; xp.f32 = nullptr;
        GIMB_TRY(run_conv(F, b.c3, u2, B, oH, oW, 1, ACT_RELU, nullptr, xp, idn.planes()));
        GIMB_TRY(planes_to_f32(ctx, xo.sp, (int64_t)B * oH * oW, b.c3.cout, xo.f32, b.c3.cout));
      } else {
        GIMB_TRY(run_conv(F, b.c3, u2, B, oH, oW, 1, ACT_RELU, nullptr, xo, idn.planes()));
      }
      A.release(mk2);
      cur = xo; cH = oH; cW = oW;
    }
  }
  A.release(mark);
  // the image itself (feats[1]) as NHWC, pitch 4
  GIMB_TRY(nchw_to_nhwc(ctx, nchw, B, 3, H, W, py->f[0], 4));
  return 0;
}

// a GEMM-shaped layer on fp32 NHWC in / out (planes made on the fly for the tensor-core engine)
int conv_f32(Fwd& F, const Conv& c, const float* in, int ld_in, int B, int H, int W, int act, const float* residual, float* out, int ld_out,
             SplitPlanes* out_planes = nullptr) {
  Arena& A = F.ctx.arena;
  size_t mark = A.mark();
  ActT a = act_f32(const_cast<float*>(in), c.cin, ld_in);
  GIMB_TRY(with_planes(F, a, (size_t)B * H * W));
  ActT o = act_f32(out, c.cout, ld_out);
  if (out_planes) o.sp = *out_planes;
  GIMB_TRY(run_conv(F, c, a, B, H, W, 1, act, residual, o));
  A.release(mark);
  return 0;
}

// RRB.forward (dkm.py:196-202): x = conv1(in); out = relu(x + conv3(relu(bn(conv2(x)))))
int rrb(Fwd& F, const RRBw& r, const float* in, int ld_in, int B, int H, int W, float* out) {
  Arena& A = F.ctx.arena;
  size_t mark = A.mark();
  const size_t P = (size_t)B * H * W;
  ActT x = F.alloc(P, DFN_DIM, true, true);
  ActT t = F.alloc(P, DFN_DIM, false, true);
  GIMB_CHECK(F.ctx.dry || !A.overflow, "dkm rrb: workspace exhausted");
  ActT a = act_f32(const_cast<float*>(in), r.c1.cin, ld_in);
  GIMB_TRY(with_planes(F, a, P));
  GIMB_TRY(run_conv(F, r.c1, a, B, H, W, 1, ACT_NONE, nullptr, x));
  GIMB_TRY(run_conv(F, r.c2, x, B, H, W, 1, ACT_RELU, nullptr, t));
  ActT o = act_f32(out, DFN_DIM, DFN_DIM);
  GIMB_TRY(run_conv(F, r.c3, t, B, H, W, 1, ACT_RELU, x.f32, o));
  A.release(mark);
  return 0;
}

// GP.forward (dkm.py:340-370) with CosKernel (:135-144): mu = K_xy (K_yy + 0.1 I)^-1 f, all [B = 2] problems at once
int gp_regression(Fwd& F, gimb_dkm* m, int si, const ActT& fx, const ActT& fy, int B, int h, int w, float* mu /*[B, hw, 256]*/) {
  Ctx& ctx = F.ctx;
  Arena& A = ctx.arena;
  size_t mark = A.mark();
  const int N = h * w, C = 512;
  float* Kyy = A.alloc<float>((size_t)B * N * N);
  float* Kxy = A.alloc<float>((size_t)B * N * N);
  float* f = A.alloc<float>((size_t)B * N * GP_DIM);
  GIMB_CHECK(ctx.dry || !A.overflow, "dkm gp: workspace exhausted");
  bool tc_gram = F.tc() && N >= 128 && N % 4 == 0;
  if (tc_gram) {
    // dot products on the tensor cores: a Linear layer whose "weights" are the other token set (per batch entry)
    for (int b = 0; b < B; ++b) {
      Wt wy;
      wy.wp.hi = fy.sp.hi + (size_t)b * N * fy.sp.ld; wy.wp.lo = fy.sp.lo + (size_t)b * N * fy.sp.ld; wy.wp.ld = fy.sp.ld;
      Epi e;
      const ActT xb = view_rows(fx, (size_t)b * N), yb = view_rows(fy, (size_t)b * N);
      GIMB_TRY(gemm(F, wy, C, 0, N, 1, 1, yb, nullptr, 1, N, 1, e, act_f32(Kyy + (size_t)b * N * N, N, N)));
      GIMB_TRY(gemm(F, wy, C, 0, N, 1, 1, xb, nullptr, 1, N, 1, e, act_f32(Kxy + (size_t)b * N * N, N, N)));
    }
    GIMB_TRY(dkm_cos_gram_finish(ctx, Kyy, fy.f32, fy.f32, B, N, N, C, fy.pitch(), 0.2f, 0.1f));
    GIMB_TRY(dkm_cos_gram_finish(ctx, Kxy, fx.f32, fy.f32, B, N, N, C, fy.pitch(), 0.2f, 0.f));
  } else {
    GIMB_TRY(dkm_cos_gram(ctx, fy.f32, fy.f32, B, N, N, C, fy.pitch(), 0.2f, 0.1f, Kyy));
    GIMB_TRY(dkm_cos_gram(ctx, fx.f32, fy.f32, B, N, N, C, fy.pitch(), 0.2f, 0.f, Kxy));
  }
  GIMB_TRY(dkm_pos_basis(ctx, m->pos_w[si], m->pos_b[si], B, h, w, GP_DIM, f));
  if (N > 2000) {
    // REFERENCE QUIRK (dkm.py:352-356), reproduced because parity is defined against the unmodified reference: above 2000
    // tokens the reference's per-sample inversion loop slices a batch-1 `sigma_noise` with [k:k+1]; for k >= 1 the slice is
    // empty, so only K_yy[0] is inverted and that inverse is broadcast over the batch.  f does not depend on the batch
    // entry, so one solve with K_yy[0] serves both.
    GIMB_TRY(dkm_chol_solve(ctx, Kyy, f, 1, N, GP_DIM));
    if (!ctx.dry)
      for (int b = 1; b < B; ++b)
        GIMB_CUDA(cudaMemcpyAsync(f + (size_t)b * N * GP_DIM, f, (size_t)N * GP_DIM * sizeof(float), cudaMemcpyDeviceToDevice, ctx.stream));
  } else {
    GIMB_TRY(dkm_chol_solve(ctx, Kyy, f, B, N, GP_DIM));
  }
  GIMB_TRY(dkm_matmul_nn(ctx, Kxy, f, B, N, N, GP_DIM, mu, GP_DIM));
  A.release(mark);
  return 0;
}

// ConvRefiner.forward (dkm.py:75-123): returns the head [B, hw, 8] = (certainty, dx, dy, 0...) in `head`
int conv_refiner(Fwd& F, const Refiner& r, const float* x, const float* y, int ld_xy, int B, int h, int w, const float* flow, float* head,
                 float* tap_in = nullptr, float* tap_dw = nullptr, float* tap_pw = nullptr, float* tap_out = nullptr) {
  Ctx& ctx = F.ctx;
  Arena& A = ctx.arena;
  size_t mark = A.mark();
  const size_t P = (size_t)B * h * w;
  const int ldd = pitch8(r.cin), ldh = pitch8(r.hidden);
  float* d = A.alloc<float>(P * ldd);
  ActT ta = F.alloc(P, r.hidden, false, true);   // depthwise output: planes for the pointwise GEMM
  float* tb = A.alloc<float>(P * ldh);            // pointwise output (fp32, padded pitch)
  GIMB_CHECK(ctx.dry || !A.overflow, "dkm refiner: workspace exhausted");
  // d = cat(x, x_hat, emb, local_corr); the pad channels up to the pitch are read (times zero weights) by the vectorised
  // depthwise kernel, so they must not hold NaN bit patterns left in the workspace
  GIMB_TRY(dkm_zero_pad_channels(ctx, d, (int64_t)P, r.cin, ldd));
  GIMB_TRY(dkm_copy_channels(ctx, x, (int64_t)P, r.feat, ld_xy, d, ldd, 0));
  GIMB_TRY(dkm_grid_sample(ctx, y, B, h, w, r.feat, ld_xy, flow, d, ldd, r.feat));
  GIMB_TRY(dkm_disp_emb(ctx, flow, B, h, w, r.emb_w, r.emb_b, r.emb, d, ldd, 2 * r.feat));
  if (r.radius) GIMB_TRY(dkm_local_corr(ctx, x, y, B, h, w, r.feat, ld_xy, flow, r.radius, d, ldd, 2 * r.feat + r.emb));
  if (tap_in && !ctx.dry) GIMB_TRY(dkm_copy_channels(ctx, d, (int64_t)P, r.cin, ldd, tap_in, r.cin, 0));
  const float* cur = d;
  int cur_c = r.cin, cur_ld = ldd;
  for (int k = 0; k < 9; ++k) {
    const int mult = k == 0 ? r.mult : 1;
    if (F.tc() && (mult == 1 || mult == 2)) {
      GIMB_TRY(dkm_depthwise5x5_v4(ctx, cur, B, h, w, r.hidden, cur_ld, r.dw_wt[k], r.dw_sp[k], r.dw_bp[k], ldh, nullptr, 0, ta.planes(), mult));
    } else if (F.tc()) {
      GIMB_TRY(dkm_depthwise5x5(ctx, cur, B, h, w, cur_c, cur_ld, mult, r.dw_w[k], r.dw_s[k], r.dw_b[k], nullptr, 0, ta.planes()));
    } else {
      GIMB_TRY(dkm_depthwise5x5(ctx, cur, B, h, w, cur_c, cur_ld, mult, r.dw_w[k], r.dw_s[k], r.dw_b[k], ta.f32, r.hidden, nullptr));
    }
    if (k == 0 && tap_dw && !ctx.dry && F.tc()) GIMB_TRY(planes_to_f32(ctx, ta.sp, (int64_t)P, r.hidden, tap_dw, r.hidden));
    ActT o = act_f32(tb, r.hidden, F.tc() ? ldh : r.hidden);
    GIMB_TRY(run_conv(F, r.pw[k], ta, B, h, w, 1, ACT_NONE, nullptr, o));
    if (k == 0 && tap_pw && !ctx.dry) GIMB_TRY(dkm_copy_channels(ctx, tb, (int64_t)P, r.hidden, o.pitch(), tap_pw, r.hidden, 0));
    if (k == 8 && tap_out && !ctx.dry) GIMB_TRY(dkm_copy_channels(ctx, tb, (int64_t)P, r.hidden, o.pitch(), tap_out, r.hidden, 0));
    cur = tb; cur_c = r.hidden; cur_ld = o.pitch();
  }
  GIMB_TRY(conv_f32(F, r.out, cur, cur_ld, B, h, w, ACT_NONE, nullptr, head, 8));
  A.release(mark);
  return 0;
}

struct DecoderOut {
  float* flow1;   // [2, h, w, 2] finest flow
  float* cert1;   // [2, h, w]
  float* cert16;  // [2, h16, w16] (first pass only)
  int h16, w16;
};

// Decoder.forward (dkm.py:454-534).  f2 = f1 with the two batch halves swapped (forward_symmetric, dkm.py:640-650).
int decoder(Fwd& F, gimb_dkm* m, const Pyramid& py, const Pyramid& sw, bool upsample, const float* in_flow, const float* in_cert, int in_h,
            int in_w, DecoderOut* out, const gimb_dkm_taps* taps) {
  Ctx& ctx = F.ctx;
  Arena& A = ctx.arena;
  const int B = 2;
  const int H = py.H[0], W = py.W[0];
  const int first = upsample ? 3 : 5;  // log2 of the coarsest scale
  float* flow = A.alloc<float>((size_t)B * H * W * 2);
  float* cert = A.alloc<float>((size_t)B * H * W);
  float* flow_n = A.alloc<float>((size_t)B * H * W * 2);
  float* cert_n = A.alloc<float>((size_t)B * H * W);
  float* head = A.alloc<float>((size_t)B * H * W * 8);
  float* old = A.alloc<float>((size_t)B * py.H[4] * py.W[4] * DFN_DIM * (upsample ? 0 : 1) + 4);
  float* old_n = A.alloc<float>((size_t)B * py.H[4] * py.W[4] * DFN_DIM * (upsample ? 0 : 1) + 4);
  out->cert16 = A.alloc<float>((size_t)B * py.H[4] * py.W[4]);
  out->h16 = py.H[4]; out->w16 = py.W[4];
  GIMB_CHECK(ctx.dry || !A.overflow, "dkm decoder: workspace exhausted");
  bool have_cert = false;
  if (!upsample) {
    GIMB_TRY(dkm_grid_flow(ctx, flow, B, py.H[5], py.W[5]));
    GIMB_TRY(dkm_fill(ctx, old, (size_t)B * py.H[5] * py.W[5] * DFN_DIM, 0.f));
  } else {
    GIMB_TRY(dkm_resize_nhwc(ctx, in_flow, B, in_h, in_w, 2, 2, flow, py.H[3], py.W[3], 2));
    GIMB_TRY(dkm_resize_nhwc(ctx, in_cert, B, in_h, in_w, 1, 1, cert, py.H[3], py.W[3], 1));
    have_cert = true;
  }
  for (int s = first; s >= 0; --s) {
    const int ins = 1 << s, hs = py.H[s], ws = py.W[s];
    const size_t P = (size_t)B * hs * ws;
    size_t mark = A.mark();
    const float* f1 = py.f[s];
    const float* f2 = sw.f[s];
    int ldf = s == 0 ? 4 : py.C[s];
    if (s >= 4) {
      // 1x1 projection to 512 (DKMv3.py:137-139), GP, DFN
      const int si = s == 5 ? 0 : 1;
      ActT p1 = F.alloc(P, 512, true, true), p2 = F.alloc(P, 512, true, true);
      float* gpo = A.alloc<float>(P * GP_DIM);
      float* cat = A.alloc<float>(P * (GP_DIM + FEAT_DIM));
      float* feats = A.alloc<float>(P * FEAT_DIM);
      float* emb = A.alloc<float>(P * DFN_DIM);
      float* ctxv = A.alloc<float>(P * DFN_DIM);
      float* scratch = A.alloc<float>((size_t)B * 4 * DFN_DIM);
      GIMB_CHECK(ctx.dry || !A.overflow, "dkm decoder: workspace exhausted");
      {
        size_t mk = A.mark();
        ActT a1 = act_f32(const_cast<float*>(f1), py.C[s], py.C[s]), a2 = act_f32(const_cast<float*>(f2), py.C[s], py.C[s]);
        GIMB_TRY(with_planes(F, a1, P));
        GIMB_TRY(with_planes(F, a2, P));
        GIMB_TRY(run_conv(F, m->proj[si], a1, B, hs, ws, 1, ACT_NONE, nullptr, p1));
        GIMB_TRY(run_conv(F, m->proj[si], a2, B, hs, ws, 1, ACT_NONE, nullptr, p2));
        A.release(mk);
      }
      if (s == 4) {  // old_stuff to the new size (dkm.py:489-491)
        GIMB_TRY(dkm_resize_nhwc(ctx, old, B, py.H[5], py.W[5], DFN_DIM, DFN_DIM, old_n, hs, ws, DFN_DIM));
        std::swap(old, old_n);
      }
      GIMB_TRY(gp_regression(F, m, si, p1, p2, B, hs, ws, gpo));
      if (taps && (si == 0 ? taps->gp32 : taps->gp16) && !ctx.dry)
        GIMB_CUDA(cudaMemcpyAsync(si == 0 ? taps->gp32 : taps->gp16, gpo, P * GP_DIM * sizeof(float), cudaMemcpyDeviceToDevice, ctx.stream));
      // DFN.forward (dkm.py:245-254)
      GIMB_TRY(run_conv(F, m->dfn_feat[si], p1, B, hs, ws, 1, ACT_NONE, nullptr, act_f32(feats, FEAT_DIM, FEAT_DIM)));
      GIMB_TRY(dkm_copy_channels(ctx, feats, (int64_t)P, FEAT_DIM, FEAT_DIM, cat, GP_DIM + FEAT_DIM, 0));
      GIMB_TRY(dkm_copy_channels(ctx, gpo, (int64_t)P, GP_DIM, GP_DIM, cat, GP_DIM + FEAT_DIM, FEAT_DIM));
      GIMB_TRY(rrb(F, m->rrb_d[si], cat, GP_DIM + FEAT_DIM, B, hs, ws, emb));
      GIMB_TRY(dkm_cab(ctx, old, emb, B, hs * ws, DFN_DIM, m->cab_w1[si], m->cab_b1[si], m->cab_w2[si], m->cab_b2[si], ctxv, scratch));
      GIMB_TRY(rrb(F, m->rrb_u[si], ctxv, DFN_DIM, B, hs, ws, old));
      GIMB_TRY(conv_f32(F, m->dfn_term[si], old, DFN_DIM, B, hs, ws, ACT_NONE, nullptr, head, 8));
      GIMB_TRY(dkm_split_head(ctx, head, 8, (int64_t)P, flow, cert));
      have_cert = true;
      if (s <= 4) {  // conv_refiner["16"]
        if (taps && taps->dfn_flow16 && !ctx.dry)
          GIMB_CUDA(cudaMemcpyAsync(taps->dfn_flow16, flow, P * 2 * sizeof(float), cudaMemcpyDeviceToDevice, ctx.stream));
        GIMB_TRY(conv_refiner(F, m->ref[0], p1.f32, p2.f32, 512, B, hs, ws, flow, head, taps ? taps->refiner_in16 : nullptr,
                              taps ? taps->refiner_dw16 : nullptr, taps ? taps->refiner_pw16 : nullptr, taps ? taps->refiner_out16 : nullptr));
        GIMB_TRY(dkm_apply_delta(ctx, flow, cert, true, head, 8, B, hs, ws, (float)ins, W, H));
      }
      if (s == 4 && !ctx.dry)
        GIMB_CUDA(cudaMemcpyAsync(out->cert16, cert, P * sizeof(float), cudaMemcpyDeviceToDevice, ctx.stream));
    } else {
      const int ri = 4 - s;  // ref[1] = "8", ref[2] = "4", ref[3] = "2", ref[4] = "1"
      GIMB_TRY(conv_refiner(F, m->ref[ri], f1, f2, ldf, B, hs, ws, flow, head));
      GIMB_TRY(dkm_apply_delta(ctx, flow, cert, have_cert, head, 8, B, hs, ws, (float)ins, W, H));
      have_cert = true;
    }
    if (taps && !ctx.dry) {
      float* tf = upsample ? taps->flow_up[s] : taps->flow[s];
      float* tcert = upsample ? taps->cert_up[s] : taps->cert[s];
      if (tf) GIMB_CUDA(cudaMemcpyAsync(tf, flow, P * 2 * sizeof(float), cudaMemcpyDeviceToDevice, ctx.stream));
      if (tcert) GIMB_CUDA(cudaMemcpyAsync(tcert, cert, P * sizeof(float), cudaMemcpyDeviceToDevice, ctx.stream));
    }
    A.release(mark);
    if (s > 0) {  // to the next (finer) scale (dkm.py:517-530)
      GIMB_TRY(dkm_resize_nhwc(ctx, flow, B, hs, ws, 2, 2, flow_n, py.H[s - 1], py.W[s - 1], 2));
      GIMB_TRY(dkm_resize_nhwc(ctx, cert, B, hs, ws, 1, 1, cert_n, py.H[s - 1], py.W[s - 1], 1));
      std::swap(flow, flow_n);
      std::swap(cert, cert_n);
    }
  }
  out->flow1 = flow;
  out->cert1 = cert;
  return 0;
}

struct MatchArgs {
  const float *im1, *im2;
  int H1, W1, H2, W2, h, w, upsample, uh, uw;
  float *warp, *cert;
  const gimb_dkm_taps* taps;
};

// pyramid with the two batch halves swapped (views for batch-major tensors need a copy: rows of image 1 then image 0)
int swapped(Ctx& ctx, const Pyramid& py, Pyramid* sw) {
  *sw = py;
  for (int i = 0; i < 6; ++i) {
    if (!py.f[i]) continue;
    const int ld = i == 0 ? 4 : py.C[i];
    const size_t half = (size_t)py.H[i] * py.W[i] * ld;
    float* d = ctx.arena.alloc<float>(2 * half);
    GIMB_CHECK(ctx.dry || !ctx.arena.overflow, "dkm: workspace exhausted (swapped pyramid)");
    sw->f[i] = d;
    if (!ctx.dry) {
      GIMB_CUDA(cudaMemcpyAsync(d, py.f[i] + half, half * sizeof(float), cudaMemcpyDeviceToDevice, ctx.stream));
      GIMB_CUDA(cudaMemcpyAsync(d + half, py.f[i], half * sizeof(float), cudaMemcpyDeviceToDevice, ctx.stream));
    }
  }
  return 0;
}

int match_impl(Ctx& ctx, gimb_dkm* m, const MatchArgs& a) {
  Arena& A = ctx.arena;
  Fwd F{ctx, m->engine};
  GIMB_CHECK(a.h >= 32 && a.w >= 32 && a.h % 2 == 0 && a.w % 2 == 0, "gimb_dkm_match: h_resized / w_resized must be even and >= 32 (got %dx%d)",
             a.h, a.w);
  GIMB_CHECK(!a.upsample || (a.uh >= 8 && a.uw >= 8 && a.uh % 2 == 0 && a.uw % 2 == 0), "gimb_dkm_match: upsample_res must be even");
  // ---- pass 1 at (h, w)
  float* batch = A.alloc<float>((size_t)2 * 3 * a.h * a.w);
  GIMB_CHECK(ctx.dry || !A.overflow, "gimb_dkm_match: workspace too small");
  GIMB_TRY(dkm_resize_nchw(ctx, a.im1, 1, 3, a.H1, a.W1, batch, a.h, a.w));
  GIMB_TRY(dkm_resize_nchw(ctx, a.im2, 1, 3, a.H2, a.W2, batch + (size_t)3 * a.h * a.w, a.h, a.w));
  Pyramid py, sw;
  GIMB_TRY(encoder(F, m, batch, 2, a.h, a.w, 32, &py));
  if (a.taps && !ctx.dry)
    for (int i = 1; i <= 5; ++i)
      if (a.taps->enc[i])
        GIMB_CUDA(cudaMemcpyAsync(a.taps->enc[i], py.f[i], (size_t)2 * py.H[i] * py.W[i] * py.C[i] * sizeof(float), cudaMemcpyDeviceToDevice,
                                  ctx.stream));
  GIMB_TRY(swapped(ctx, py, &sw));
  DecoderOut d1;
  GIMB_TRY(decoder(F, m, py, sw, false, nullptr, nullptr, 0, 0, &d1, a.taps));
  int hs = a.h, ws = a.w;
  const float *flow = d1.flow1, *cert = d1.cert1;
  float* low = nullptr;
  DecoderOut d2;
  if (a.upsample) {
    hs = a.uh; ws = a.uw;
    float* batch2 = A.alloc<float>((size_t)2 * 3 * hs * ws);
    GIMB_CHECK(ctx.dry || !A.overflow, "gimb_dkm_match: workspace too small");
    GIMB_TRY(dkm_resize_nchw(ctx, a.im1, 1, 3, a.H1, a.W1, batch2, hs, ws));
    GIMB_TRY(dkm_resize_nchw(ctx, a.im2, 1, 3, a.H2, a.W2, batch2 + (size_t)3 * hs * ws, hs, ws));
    Pyramid py2, sw2;
    GIMB_TRY(encoder(F, m, batch2, 2, hs, ws, 8, &py2));
    GIMB_TRY(swapped(ctx, py2, &sw2));
    GIMB_TRY(decoder(F, m, py2, sw2, true, d1.flow1, d1.cert1, a.h, a.w, &d2, a.taps));
    flow = d2.flow1; cert = d2.cert1;
  }
  // low-resolution certainty (scale 16 of pass 1) at the output size (dkm.py:686-693)
  low = A.alloc<float>((size_t)2 * hs * ws);
  GIMB_CHECK(ctx.dry || !A.overflow, "gimb_dkm_match: workspace too small");
  GIMB_TRY(dkm_resize_nhwc(ctx, d1.cert16, 2, d1.h16, d1.w16, 1, 1, low, hs, ws, 1));
  DkmFinalArgs fa;
  fa.flow = flow; fa.certainty = cert; fa.low_cert = low; fa.im1 = a.im1; fa.im2 = a.im2;
  fa.H1 = a.H1; fa.W1 = a.W1; fa.H2 = a.H2; fa.W2 = a.W2; fa.hs = hs; fa.ws = ws; fa.warp = a.warp; fa.cert_out = a.cert;
  GIMB_TRY(dkm_finalize(ctx, fa));
  return 0;
}

}  // namespace
}  // namespace gimb

// =================================================================================================
extern "C" {

int gimb_dkm_create(const void* blob, size_t nbytes, int device, gimb_dkm** out) {
  GIMB_CHECK(blob && out, "gimb_dkm_create: null argument");
  DeviceGuard guard(device);
  GIMB_CHECK(guard.ok, "gimb_dkm_create: cudaSetDevice(%d) failed", device);
  cudaDeviceProp prop;
  GIMB_CUDA(cudaGetDeviceProperties(&prop, device));
  GIMB_CHECK(prop.major == 10, "libgimb200 is built for sm_100a (B200) only; device %d is sm_%d%d", device, prop.major, prop.minor);
  gimb_dkm* m = new gimb_dkm();
  m->device = device;
  m->sm_count = prop.multiProcessorCount;
  m->ws.device = device;
  m->ws.sm_count = m->sm_count;
  Ctx cctx;
  cctx.sm_count = m->sm_count;
  if (m->ws.upload(blob, nbytes) != 0 || build_dkm(m, cctx) != 0 || m->ws.alloc_planes() != 0 || build_dkm(m, cctx) != 0 ||
      cudaDeviceSynchronize() != cudaSuccess) {
    m->ws.release();
    delete m;
    return 1;
  }
  const char* eng = getenv("GIMB_ENGINE");
  if (eng && std::string(eng) == "simt") m->engine = ENGINE_SIMT;
  *out = m;
  return 0;
}

void gimb_dkm_destroy(gimb_dkm* h) {
  if (!h) return;
  DeviceGuard guard(h->device);
  h->ws.release();
  delete h;
}

int gimb_dkm_set_engine(gimb_dkm* h, int engine) {
  GIMB_CHECK(h && (engine == ENGINE_SIMT || engine == ENGINE_TC), "gimb_dkm_set_engine: bad argument");
  h->engine = engine;
  return 0;
}
uint64_t gimb_dkm_launch_count(gimb_dkm* h) { return h ? h->launches : 0; }

static int dkm_run(gimb_dkm* h, const MatchArgs& a, void* workspace, size_t workspace_bytes, bool dry, size_t* need, void* stream) {
  Ctx ctx;
  ctx.stream = (cudaStream_t)stream;
  ctx.sm_count = h->sm_count;
  ctx.dry = dry;
  ctx.arena.dry = dry;
  // like gimb_loftr_forward: the arena starts at the next kAlign boundary of whatever the caller passes
  const uintptr_t base = ((uintptr_t)workspace + Arena::kAlign - 1) / Arena::kAlign * Arena::kAlign;
  const size_t skew = base - (uintptr_t)workspace;
  GIMB_CHECK(dry || workspace_bytes > skew, "gimb_dkm_match: workspace too small");
  ctx.arena.base = (char*)base;
  ctx.arena.cap = dry ? 0 : workspace_bytes - skew;
  int rc = match_impl(ctx, h, a);
  if (need) *need = ctx.arena.peak + 4096 + Arena::kAlign;
  h->launches += ctx.launches;
  return rc;
}

int gimb_dkm_workspace_bytes(gimb_dkm* h, int H1, int W1, int H2, int W2, int h_resized, int w_resized, int upsample_preds, int up_h,
                             int up_w, size_t* bytes) {
  GIMB_CHECK(h && bytes, "gimb_dkm_workspace_bytes: null argument");
  MatchArgs a = {};
  a.H1 = H1; a.W1 = W1; a.H2 = H2; a.W2 = W2; a.h = h_resized; a.w = w_resized; a.upsample = upsample_preds; a.uh = up_h; a.uw = up_w;
  return dkm_run(h, a, nullptr, 0, true, bytes, nullptr);
}

int gimb_kde_density(const float* points, int n, float std, float* density, void* stream) {
  GIMB_CHECK(points && density && n >= 0 && std > 0.f, "gimb_kde_density: bad argument");
  GIMB_CHECK(((uintptr_t)points & 15) == 0, "gimb_kde_density: points must be 16-byte aligned");
  Ctx ctx;
  ctx.stream = (cudaStream_t)stream;
  return dkm_kde(ctx, points, n, std, density);
}

int gimb_dkm_match(gimb_dkm* h, const float* im1, int H1, int W1, const float* im2, int H2, int W2, int h_resized, int w_resized,
                   int upsample_preds, int up_h, int up_w, void* workspace, size_t workspace_bytes, float* warp, float* certainty,
                   const gimb_dkm_taps* taps, void* stream) {
  GIMB_CHECK(h && im1 && im2 && warp && certainty && workspace, "gimb_dkm_match: null argument");
  GIMB_CHECK(H1 > 0 && W1 > 0 && H2 > 0 && W2 > 0, "gimb_dkm_match: bad image size");
  DeviceGuard guard(h->device);
  GIMB_CHECK(guard.ok, "gimb_dkm_match: cudaSetDevice failed");
  MatchArgs a = {};
  a.im1 = im1; a.im2 = im2; a.H1 = H1; a.W1 = W1; a.H2 = H2; a.W2 = W2; a.h = h_resized; a.w = w_resized;
  a.upsample = upsample_preds; a.uh = up_h; a.uw = up_w; a.warp = warp; a.cert = certainty; a.taps = taps;
  return dkm_run(h, a, workspace, workspace_bytes, false, nullptr, stream);
}

}  // extern "C"
