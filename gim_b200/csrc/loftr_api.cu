// loftr_api.cu - the C ABI of libgimb200 (include/gimb200.h): weight blob -> device model, workspace
// planning, and the orchestration of one gim_loftr forward (networks/loftr/loftr.py:43-91).
#include <math.h>
#include <stdlib.h>

#include <algorithm>
#include <map>
#include <string>
#include <utility>
#include <vector>

#include "engine.cuh"

namespace gimb {
const char* last_error();

namespace {

struct Bottleneck {
  Conv c1, c2, c3, ds;
  bool has_ds = false;
  int stride = 1;
};
struct EncLayer {
  Wt q, kv, merge, mlp0, mlp2;
  const float *n1g, *n1b, *n2g, *n2b;
};

constexpr int FINE_CHUNK = 16384;  // matches per fine-stage pass

}  // namespace
}  // namespace gimb

using namespace gimb;

struct gimb_loftr {
  int device = 0;
  int engine = ENGINE_TC;
  gimb_loftr_cfg cfg;
  WeightStore ws;              // packed fp32 tensors + fp16 weight planes of every GEMM layer
  Conv stem;
  std::vector<Bottleneck> layers[3];
  Conv l3out, l2out, l2c1, l2c2, l1out, l1c1, l1c2;
  EncLayer coarse[8], fine[2];
  std::map<std::pair<int, int>, float*> pe_cache;
  int64_t* host_count = nullptr;  // pinned
  uint64_t launches = 0;
  uint64_t corr_fallbacks = 0;    // forwards that repeated the coarse matching with the exact sweeps
  int sm_count = 148;
  bool profiling = false;
  std::vector<std::pair<std::string, float>> last_profile;
  std::vector<std::string> prof_names;
};

namespace gimb {
namespace {

int load_enc(gimb_loftr* m, Ctx& ctx, const std::string& pre, EncLayer* e) {
  GIMB_TRY(m->ws.load_linear(ctx, pre + ".q", &e->q));
  GIMB_TRY(m->ws.load_linear(ctx, pre + ".kv", &e->kv));
  GIMB_TRY(m->ws.load_linear(ctx, pre + ".merge", &e->merge));
  GIMB_TRY(m->ws.load_linear(ctx, pre + ".mlp0", &e->mlp0));
  GIMB_TRY(m->ws.load_linear(ctx, pre + ".mlp2", &e->mlp2));
  GIMB_TRY(m->ws.find(pre + ".n1g", &e->n1g));
  GIMB_TRY(m->ws.find(pre + ".n1b", &e->n1b));
  GIMB_TRY(m->ws.find(pre + ".n2g", &e->n2g));
  GIMB_TRY(m->ws.find(pre + ".n2b", &e->n2b));
  return 0;
}

int build_model(gimb_loftr* m, Ctx& ctx) {
  for (int li = 0; li < 3; ++li) m->layers[li].clear();
  GIMB_TRY(m->ws.load_conv(ctx, "stem", true, &m->stem));
  GIMB_CHECK(m->stem.k == 7 && m->stem.cin == 3 && m->stem.cout == 64, "stem must be 7x7 3->64");
  const int nblk[3] = {3, 4, 6};
  for (int li = 0; li < 3; ++li) {
    for (int bi = 0; bi < nblk[li]; ++bi) {
      Bottleneck b;
      std::string pre = "l" + std::to_string(li + 1) + "." + std::to_string(bi);
      GIMB_TRY(m->ws.load_conv(ctx, pre + ".c1", true, &b.c1));
      GIMB_TRY(m->ws.load_conv(ctx, pre + ".c2", true, &b.c2));
      GIMB_TRY(m->ws.load_conv(ctx, pre + ".c3", true, &b.c3));
      b.has_ds = (bi == 0);
      b.stride = (li > 0 && bi == 0) ? 2 : 1;
      if (b.has_ds) GIMB_TRY(m->ws.load_conv(ctx, pre + ".ds", true, &b.ds));
      m->layers[li].push_back(b);
    }
  }
  GIMB_TRY(m->ws.load_conv(ctx, "fpn.l3out", false, &m->l3out));
  GIMB_TRY(m->ws.load_conv(ctx, "fpn.l2out", false, &m->l2out));
  GIMB_TRY(m->ws.load_conv(ctx, "fpn.l2c1", true, &m->l2c1));
  GIMB_TRY(m->ws.load_conv(ctx, "fpn.l2c2", false, &m->l2c2));
  GIMB_TRY(m->ws.load_conv(ctx, "fpn.l1out", false, &m->l1out));
  GIMB_TRY(m->ws.load_conv(ctx, "fpn.l1c1", true, &m->l1c1));
  GIMB_TRY(m->ws.load_conv(ctx, "fpn.l1c2", false, &m->l1c2));
  for (int i = 0; i < 8; ++i) GIMB_TRY(load_enc(m, ctx, "coarse." + std::to_string(i), &m->coarse[i]));
  for (int i = 0; i < 2; ++i) GIMB_TRY(load_enc(m, ctx, "fine." + std::to_string(i), &m->fine[i]));
  return 0;
}

// ResNet trunk + FPN (networks/loftr/backbone/resnet.py:214-235, 306-329).  NCHW in; feat_c / feat_f are fp32 NHWC.
int backbone(Fwd& F, gimb_loftr* m, const float* color, int B, int H, int W, float* feat_c, float* feat_f) {
  Ctx& ctx = F.ctx;
  Arena& A = ctx.arena;
  const int H2 = H / 2, W2 = W / 2, H4 = H / 4, W4 = W / 4, H8 = H / 8, W8 = W / 8;
  const size_t P2 = (size_t)B * H2 * W2, P4 = (size_t)B * H4 * W4, P8 = (size_t)B * H8 * W8;
  size_t mark = A.mark();
  // block outputs.  tcgen05 engine: fp16 planes only - they are the next GEMM operand AND the identity of the
  // next block (x = hi + lo * 2^-8); the fp32 copy would add 8 B/element of HBM traffic to every block.
  ActT xs[3] = {F.alloc(P2, 256, false, true), F.alloc(P4, 512, false, true), F.alloc(P8, 1024, false, true)};
  {
    size_t mk = A.mark();
    ActT x0 = F.alloc(P2, 64, false, true);
    GIMB_CHECK(ctx.dry || !A.overflow, "backbone: workspace exhausted");
    GIMB_TRY(stem_conv7x7(ctx, color, B, H, W, m->stem.wt.w, m->stem.s, m->stem.b, x0.f32, x0.planes()));
    ActT cur = x0;
    int cH = H2, cW = W2;
    for (int li = 0; li < 3; ++li) {
      ActT& xo = xs[li];
      for (size_t bi = 0; bi < m->layers[li].size(); ++bi) {
        const Bottleneck& b = m->layers[li][bi];
        const int oH = cH / b.stride, oW = cW / b.stride;
        size_t mk2 = A.mark();
        ActT u1 = F.alloc((size_t)B * cH * cW, b.c1.cout, false, true);
        ActT u2 = F.alloc((size_t)B * oH * oW, b.c2.cout, false, true);
        GIMB_CHECK(ctx.dry || !A.overflow, "backbone: workspace exhausted");
        GIMB_TRY(run_conv(F, b.c1, cur, B, cH, cW, 1, ACT_RELU, nullptr, u1));
        GIMB_TRY(run_conv(F, b.c2, u1, B, cH, cW, b.stride, ACT_RELU, nullptr, u2));
        if (b.has_ds) GIMB_TRY(run_conv(F, b.ds, cur, B, cH, cW, b.stride, ACT_NONE, nullptr, xo));
        // the identity (xo) is read and overwritten element-wise by the same thread: in-place is safe
        GIMB_TRY(run_conv(F, b.c3, u2, B, oH, oW, 1, ACT_RELU, F.tc() ? nullptr : xo.f32, xo, xo.planes()));
        A.release(mk2);
        cur = xo; cH = oH; cW = oW;
      }
    }
    A.release(mk);
  }
  // FPN
  ActT fc; fc.f32 = feat_c; fc.C = 256;
  ActT ffm; ffm.f32 = feat_f; ffm.C = 128;
  GIMB_TRY(run_conv(F, m->l3out, xs[2], B, H8, W8, 1, ACT_NONE, nullptr, fc));
  ActT x2s = F.alloc(P4, 256, true, true);   // 1x1(x2) fp32, then (+ upsampled x3_out) as planes
  ActT x2t = F.alloc(P4, 256, false, true);
  ActT x2o = F.alloc(P4, 196, true, false);
  ActT x1s = F.alloc(P2, 196, true, true);
  ActT x1t = F.alloc(P2, 196, false, true);
  GIMB_CHECK(ctx.dry || !A.overflow, "backbone: workspace exhausted");
  {
    ActT o; o.f32 = x2s.f32; o.C = 256;
    GIMB_TRY(run_conv(F, m->l2out, xs[1], B, H4, W4, 1, ACT_NONE, nullptr, o));
  }
  GIMB_TRY(upsample2x_add(ctx, feat_c, B, H8, W8, 256, x2s.f32, x2s.planes()));
  GIMB_TRY(run_conv(F, m->l2c1, x2s, B, H4, W4, 1, ACT_LEAKY, nullptr, x2t));
  GIMB_TRY(run_conv(F, m->l2c2, x2t, B, H4, W4, 1, ACT_NONE, nullptr, x2o));
  {
    ActT o; o.f32 = x1s.f32; o.C = 196;
    GIMB_TRY(run_conv(F, m->l1out, xs[0], B, H2, W2, 1, ACT_NONE, nullptr, o));
  }
  GIMB_TRY(upsample2x_add(ctx, x2o.f32, B, H4, W4, 196, x1s.f32, x1s.planes()));
  GIMB_TRY(run_conv(F, m->l1c1, x1s, B, H2, W2, 1, ACT_LEAKY, nullptr, x1t));
  GIMB_TRY(run_conv(F, m->l1c2, x1t, B, H2, W2, 1, ACT_NONE, nullptr, ffm));
  A.release(mark);
  return 0;
}

int linear(Fwd& F, const Wt& wt, const ActT& x, const ActT* x2, int64_t rows, int C1, int C2, int cout, int act0, int act1,
           int split, float div, const uint8_t* row_mask, const ActT& out) {
  Epi e;
  e.act0 = act0; e.act1 = act1; e.act_split = split; e.div = div; e.row_mask = row_mask;
  return gemm(F, wt, C1, C2, cout, 1, 1, x, x2, 1, (int)rows, 1, e, out);
}

// LoFTREncoderLayer.forward (networks/loftr/submodules/transformer.py:35-58); x is updated in place.
// n sequences; x [n, L, C], src [n, S, C].  fine == true selects the per-match attention kernel.
int encoder_layer(Fwd& F, const EncLayer& e, const ActT& x, const ActT& src, int64_t n, int L, int S, int C, int nhead,
                  const uint8_t* xmask, const uint8_t* smask, bool fine) {
  Ctx& ctx = F.ctx;
  Arena& A = ctx.arena;
  size_t mark = A.mark();
  const int64_t RL = n * L, RS = n * S;
  ActT q = F.alloc(RL, C, true, false);
  ActT kv = F.alloc(RS, 2 * C, true, false);
  ActT msg = F.alloc(RL, C, false, true);       // attention output -> merge GEMM operand
  ActT mrg = F.alloc(RL, C, false, true);       // norm1(merge(msg)): planes on the tcgen05 engine, fp32 on the CUDA-core one
  ActT hid = F.alloc(RL, 2 * C, false, true);
  ActT o2;
  if (!F.tc()) o2 = F.alloc(RL, C, true, false);
  GIMB_CHECK(ctx.dry || !A.overflow, "encoder_layer: workspace exhausted");
  GIMB_TRY(linear(F, e.q, x, nullptr, RL, C, 0, C, ACT_ELU1, ACT_ELU1, 1 << 30, 1.f, xmask, q));
  GIMB_TRY(linear(F, e.kv, src, nullptr, RS, C, 0, 2 * C, ACT_ELU1, ACT_DIVS, C, (float)S, smask, kv));
  if (fine)
    GIMB_TRY(fine_attention(ctx, q.f32, kv.f32, n, L, C, nhead, msg.f32, msg.planes()));
  else
    GIMB_TRY(linear_attention(ctx, q.f32, kv.f32, (int)n, L, S, C, nhead, msg.f32, msg.planes()));
  if (F.tc()) {
    // merge + norm1 and mlp.2 + norm2 + residual: LayerNorm fused into the GEMM epilogue (the full row is one tile)
    Epi e1;
    e1.scale = e.n1g; e1.bias = e.n1b; e1.layernorm = true;
    ActT o1; o1.sp = mrg.sp; o1.C = C;
    GIMB_TRY(gemm(F, e.merge, C, 0, C, 1, 1, msg, nullptr, 1, (int)RL, 1, e1, o1));
    GIMB_TRY(linear(F, e.mlp0, x, &mrg, RL, C, C, 2 * C, ACT_RELU, ACT_RELU, 1 << 30, 1.f, nullptr, hid));
    Epi e2;
    e2.scale = e.n2g; e2.bias = e.n2b; e2.layernorm = true; e2.residual = x.f32;
    GIMB_TRY(gemm(F, e.mlp2, 2 * C, 0, C, 1, 1, hid, nullptr, 1, (int)RL, 1, e2, x));
  } else {
    ActT o; o.f32 = mrg.f32; o.C = C;
    GIMB_TRY(linear(F, e.merge, msg, nullptr, RL, C, 0, C, ACT_NONE, ACT_NONE, 1 << 30, 1.f, nullptr, o));
    GIMB_TRY(layernorm(ctx, mrg.f32, e.n1g, e.n1b, nullptr, RL, C, mrg.f32, nullptr));
    GIMB_TRY(linear(F, e.mlp0, x, &mrg, RL, C, C, 2 * C, ACT_RELU, ACT_RELU, 1 << 30, 1.f, nullptr, hid));
    GIMB_TRY(linear(F, e.mlp2, hid, nullptr, RL, 2 * C, 0, C, ACT_NONE, ACT_NONE, 1 << 30, 1.f, nullptr, o2));
    GIMB_TRY(layernorm(ctx, o2.f32, e.n2g, e.n2b, x.f32, RL, C, x.f32, nullptr));
  }
  A.release(mark);
  return 0;
}

// LocalFeatureTransformer.forward (transformer.py:80-101): (self, cross) x npairs.  `tok` holds both token sets
// back to back: rows [0, n*L) = feat0, rows [n*L, n*L + n*S) = feat1.  When L == S the self layers run as one batch.
int feature_transformer(Fwd& F, const EncLayer* layers, int npairs, const ActT& tok, int64_t n, int L, int S, int C,
                        int nhead, const uint8_t* m0, const uint8_t* m1, bool fine) {
  const ActT t0 = tok, t1 = view_rows(tok, (size_t)n * L);
  const bool batched = (L == S) && (m0 == nullptr || m1 == m0 + n * L);
  for (int i = 0; i < npairs; ++i) {
    const EncLayer& self = layers[2 * i];
    const EncLayer& cross = layers[2 * i + 1];
    if (batched) {
      GIMB_TRY(encoder_layer(F, self, t0, t0, 2 * n, L, L, C, nhead, m0, m0, fine));
    } else {
      GIMB_TRY(encoder_layer(F, self, t0, t0, n, L, L, C, nhead, m0, m0, fine));
      GIMB_TRY(encoder_layer(F, self, t1, t1, n, S, S, C, nhead, m1, m1, fine));
    }
    GIMB_TRY(encoder_layer(F, cross, t0, t1, n, L, S, C, nhead, m0, m1, fine));
    GIMB_TRY(encoder_layer(F, cross, t1, t0, n, S, L, C, nhead, m1, m0, fine));
  }
  return 0;
}

struct Prof : Marker {
  gimb_loftr* m;
  cudaStream_t st;
  std::vector<cudaEvent_t> ev;
  std::vector<std::string> names;
  bool on;
  Prof(gimb_loftr* m_, cudaStream_t s, bool enabled) : m(m_), st(s), on(enabled) {
    if (on) mark("start");
  }
  void mark(const char* name) override {
    if (!on) return;
    cudaEvent_t e;
    cudaEventCreate(&e);
    cudaEventRecord(e, st);
    ev.push_back(e);
    names.push_back(name);
  }
  void finish() {
    if (!on) return;
    cudaStreamSynchronize(st);
    m->last_profile.clear();
    for (size_t i = 1; i < ev.size(); ++i) {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, ev[i - 1], ev[i]);
      m->last_profile.push_back({names[i], ms});
    }
    for (auto e : ev) cudaEventDestroy(e);
    ev.clear();
  }
};

struct FwdArgs {
  const float *color0, *color1;
  const uint8_t *mask0, *mask1;
  const float *scale0, *scale1;
  int n, h0, w0, h1, w1;
  const gimb_loftr_out* out;
  const gimb_loftr_taps* taps;
};

int copy_tap(Ctx& ctx, float* dst, const float* src, size_t nfloat) {
  if (!dst || ctx.dry || nfloat == 0) return 0;
  GIMB_CUDA(cudaMemcpyAsync(dst, src, nfloat * sizeof(float), cudaMemcpyDeviceToDevice, ctx.stream));
  return 0;
}

// the whole forward; in dry mode only the arena is exercised (workspace planning).
int forward_impl(Ctx& ctx, gimb_loftr* m, const FwdArgs& f, int64_t* m_out) {
  Arena& A = ctx.arena;
  Fwd F{ctx, m->engine};
  const int n = f.n;
  const int h0c = f.h0 / 8, w0c = f.w0 / 8, h1c = f.h1 / 8, w1c = f.w1 / 8;
  const int h0f = f.h0 / 2, w0f = f.w0 / 2, h1f = f.h1 / 2, w1f = f.w1 / 2;
  const int L = h0c * w0c, S = h1c * w1c;
  const int C = 256, CF = 128;
  const bool same = (f.h0 == f.h1 && f.w0 == f.w1);
  const gimb_loftr_taps notaps = {};
  const gimb_loftr_taps& taps = f.taps ? *f.taps : notaps;
  Prof prof(m, ctx.stream, m->profiling && !ctx.dry);
  ctx.marker = &prof;

  // persistent buffers (bottom of the stack): coarse tokens (both images back to back) and the fine maps
  ActT tok = F.alloc((size_t)n * (L + S), C, true, true);
  float* fc0 = tok.f32;
  float* fc1 = tok.f32 + (size_t)n * L * C;
  float* ff = A.alloc<float>((size_t)n * ((size_t)h0f * w0f + (size_t)h1f * w1f) * CF);
  float* ff0 = ff;
  float* ff1 = ff + (size_t)n * h0f * w0f * CF;
  uint8_t* maskbuf = nullptr;
  if (f.mask0) maskbuf = A.alloc<uint8_t>((size_t)n * (L + S));
  GIMB_CHECK(ctx.dry || !A.overflow, "forward: workspace too small");

  // 1. backbone (loftr.py:58-63): one batched pass when both images have the same size
  if (same) {
    // cat([color0, color1]) needs a contiguous NCHW batch: stage it unless the caller already provides one
    const float* both = f.color0;
    size_t mk = A.mark();
    const size_t img = (size_t)3 * f.h0 * f.w0;
    if (ctx.dry || f.color1 != f.color0 + (size_t)n * img) {
      float* stage = A.alloc<float>(2 * n * img);
      GIMB_CHECK(ctx.dry || !A.overflow, "forward: workspace too small");
      if (!ctx.dry) {
        GIMB_CUDA(cudaMemcpyAsync(stage, f.color0, n * img * sizeof(float), cudaMemcpyDeviceToDevice, ctx.stream));
        GIMB_CUDA(cudaMemcpyAsync(stage + n * img, f.color1, n * img * sizeof(float), cudaMemcpyDeviceToDevice, ctx.stream));
      }
      both = stage;
    }
    GIMB_TRY(backbone(F, m, both, 2 * n, f.h0, f.w0, fc0, ff0));
    A.release(mk);
  } else {
    GIMB_TRY(backbone(F, m, f.color0, n, f.h0, f.w0, fc0, ff0));
    GIMB_TRY(backbone(F, m, f.color1, n, f.h1, f.w1, fc1, ff1));
  }
  prof.mark("backbone");
  GIMB_TRY(copy_tap(ctx, taps.feat_c_backbone0, fc0, (size_t)n * L * C));
  GIMB_TRY(copy_tap(ctx, taps.feat_c_backbone1, fc1, (size_t)n * S * C));
  GIMB_TRY(copy_tap(ctx, taps.feat_f0, ff0, (size_t)n * h0f * w0f * CF));
  GIMB_TRY(copy_tap(ctx, taps.feat_f1, ff1, (size_t)n * h1f * w1f * CF));

  // 2. positional encoding + coarse transformer (loftr.py:70-80)
  const float *pe0 = nullptr, *pe1 = nullptr;
  if (!ctx.dry) {
    auto i0 = m->pe_cache.find({h0c, w0c}), i1 = m->pe_cache.find({h1c, w1c});
    GIMB_CHECK(i0 != m->pe_cache.end() && i1 != m->pe_cache.end(),
               "position-encoding table for %dx%d / %dx%d not set (call gimb_loftr_set_pe)", h0c, w0c, h1c, w1c);
    pe0 = i0->second; pe1 = i1->second;
  }
  {
    const ActT t1 = view_rows(tok, (size_t)n * L);
    GIMB_TRY(add_pe(ctx, fc0, pe0, n, L, C, fc0, tok.planes()));
    GIMB_TRY(add_pe(ctx, fc1, pe1, n, S, C, fc1, t1.planes()));
  }
  const uint8_t *cm0 = nullptr, *cm1 = nullptr;
  if (f.mask0) {
    cm0 = maskbuf; cm1 = maskbuf + (size_t)n * L;
    if (!ctx.dry) {
      GIMB_CUDA(cudaMemcpyAsync(maskbuf, f.mask0, (size_t)n * L, cudaMemcpyDeviceToDevice, ctx.stream));
      GIMB_CUDA(cudaMemcpyAsync(maskbuf + (size_t)n * L, f.mask1, (size_t)n * S, cudaMemcpyDeviceToDevice, ctx.stream));
    }
  }
  GIMB_TRY(feature_transformer(F, m->coarse, 4, tok, n, L, S, C, 8, cm0, cm1, false));
  prof.mark("coarse_transformer");
  GIMB_TRY(copy_tap(ctx, taps.feat_c0, fc0, (size_t)n * L * C));
  GIMB_TRY(copy_tap(ctx, taps.feat_c1, fc1, (size_t)n * S * C));

  // 3. coarse matching (loftr.py:83)
  const gimb_loftr_out& o = *f.out;
  int64_t* dcount = A.alloc<int64_t>(2);  // [0] match count, [1] range flag of the fast correlation sweeps
  CoarseMatchArgs cm;
  cm.f0 = fc0; cm.f1 = fc1; cm.N = n; cm.L = L; cm.S = S; cm.C = C;
  cm.h0c = h0c; cm.w0c = w0c; cm.h1c = h1c; cm.w1c = w1c; cm.H0 = f.h0; cm.H1 = f.h1;
  cm.mask0 = cm0; cm.mask1 = cm1; cm.scale0 = f.scale0; cm.scale1 = f.scale1;
  cm.thr = m->cfg.thr; cm.temperature = m->cfg.dsmax_temperature; cm.border = m->cfg.border_rm;
  cm.b_ids = o.b_ids; cm.i_ids = o.i_ids; cm.j_ids = o.j_ids;
  cm.mconf = o.mconf; cm.mkpts0_c = o.mkpts0_c; cm.mkpts1_c = o.mkpts1_c;
  cm.count = dcount; cm.capacity = o.capacity; cm.conf_matrix = taps.conf_matrix;
  const ActT tok1 = view_rows(tok, (size_t)n * L);
  if (F.tc()) { cm.planes0 = tok.planes(); cm.planes1 = tok1.planes(); }
  cm.range_flag = reinterpret_cast<int*>(dcount + 1);
  GIMB_TRY(coarse_match(ctx, cm));
  prof.mark("select_compact");

  // the one host synchronisation of the forward: M sizes the fine stage (reference: torch.where)
  int64_t M = 0;
  if (!ctx.dry) {
    GIMB_CUDA(cudaMemcpyAsync(m->host_count, dcount, 2 * sizeof(int64_t), cudaMemcpyDeviceToHost, ctx.stream));
    GIMB_CUDA(cudaStreamSynchronize(ctx.stream));
    if (*reinterpret_cast<const int*>(m->host_count + 1) != 0) {
      // a softmax sum left the safe range of the fast sweeps' fixed exponent reference: repeat the coarse matching with
      // the exact online-max sweeps (same outputs, ~3x the sweep time; not observed on real features)
      m->corr_fallbacks++;
      cm.exact = true;
      GIMB_TRY(coarse_match(ctx, cm));
      GIMB_CUDA(cudaMemcpyAsync(m->host_count, dcount, 2 * sizeof(int64_t), cudaMemcpyDeviceToHost, ctx.stream));
      GIMB_CUDA(cudaStreamSynchronize(ctx.stream));
    }
    M = *m->host_count;
    GIMB_CHECK(M <= o.capacity, "forward: %lld matches exceed the output capacity %lld", (long long)M,
               (long long)o.capacity);
  } else {
    M = std::min<int64_t>((int64_t)n * std::min(L, S), FINE_CHUNK);  // plan for a full chunk
  }
  *m_out = M;

  // 4./5. fine level (loftr.py:86-91), in chunks of FINE_CHUNK matches
  const int Wn = m->cfg.fine_window, WW = Wn * Wn;
  const int stride = h0f / h0c;
  for (int64_t m0 = 0; m0 < M; m0 += FINE_CHUNK) {
    const int64_t mc = std::min<int64_t>(FINE_CHUNK, M - m0);
    size_t mk = A.mark();
    ActT win = F.alloc((size_t)2 * mc * WW, CF, true, true);  // windows of image0 then image1
    const ActT w1v = view_rows(win, (size_t)mc * WW);
    GIMB_CHECK(ctx.dry || !A.overflow, "forward: workspace too small (fine stage)");
    GIMB_TRY(fine_gather(ctx, ff0, h0f, w0f, CF, w0c, stride, Wn, o.b_ids, o.i_ids, m0, mc, win.f32, win.planes()));
    GIMB_TRY(fine_gather(ctx, ff1, h1f, w1f, CF, w1c, stride, Wn, o.b_ids, o.j_ids, m0, mc, w1v.f32, w1v.planes()));
    GIMB_TRY(feature_transformer(F, m->fine, 1, win, mc, WW, WW, CF, 8, nullptr, nullptr, true));
    GIMB_TRY(copy_tap(ctx, taps.fine_win0 ? taps.fine_win0 + (size_t)m0 * WW * CF : nullptr, win.f32, (size_t)mc * WW * CF));
    GIMB_TRY(copy_tap(ctx, taps.fine_win1 ? taps.fine_win1 + (size_t)m0 * WW * CF : nullptr, w1v.f32, (size_t)mc * WW * CF));
    FineMatchArgs fm;
    fm.f0 = win.f32; fm.f1 = w1v.f32; fm.m0 = m0; fm.m = mc; fm.WW = WW; fm.C = CF; fm.Wn = Wn;
    fm.fscale = (float)f.h0 / (float)h0f;
    fm.sim_scale = (float)(1.0 / sqrt((double)CF));
    fm.b_ids = o.b_ids; fm.scale1 = f.scale0 ? f.scale1 : nullptr;
    fm.mkpts0_c = o.mkpts0_c; fm.mkpts1_c = o.mkpts1_c;
    fm.mkpts0_f = o.mkpts0_f; fm.mkpts1_f = o.mkpts1_f; fm.expec_f = o.expec_f;
    GIMB_TRY(fine_match(ctx, fm));
    A.release(mk);
  }
  prof.mark("fine");
  prof.finish();
  return 0;
}

int check_shapes(int n, int h0, int w0, int h1, int w1) {
  GIMB_CHECK(n >= 1, "batch must be >= 1");
  GIMB_CHECK(h0 > 0 && w0 > 0 && h1 > 0 && w1 > 0 && h0 % 8 == 0 && w0 % 8 == 0 && h1 % 8 == 0 && w1 % 8 == 0,
             "image sizes must be positive multiples of 8 (got %dx%d, %dx%d)", h0, w0, h1, w1);
  return 0;
}

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// ---- GPU pre-processing: what the ZEB loader does on the host per image (datasets/utils.py:112-124) -
// uint8 HWC -> float / 255 -> CHW, zero padding at the bottom / right, and the padding mask at 1/8 resolution
// (datasets/kitti/kitti.py:115-123: nearest-neighbour 1/8 of the full-resolution mask = mask[8y, 8x]).
__global__ void u8_to_nchw_kernel(const uint8_t* __restrict__ src, int n, int ih, int iw, float* __restrict__ dst, int H, int W) {
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long total = (long long)n * 3 * H * W;
  if (idx >= total) return;
  const int x = (int)(idx % W), y = (int)((idx / W) % H), c = (int)((idx / ((long long)W * H)) % 3);
  const long long b = idx / ((long long)3 * H * W);
  float v = 0.f;
  if (y < ih && x < iw) v = __fdiv_rn((float)src[((b * ih + y) * iw + x) * 3 + c], 255.f);
  dst[idx] = v;
}
__global__ void pad_mask_kernel(uint8_t* __restrict__ mask, int n, int hc, int wc, int ih, int iw) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n * hc * wc) return;
  const int x = idx % wc, y = (idx / wc) % hc;
  mask[idx] = (8 * y < ih && 8 * x < iw) ? 1 : 0;
}

}  // namespace
}  // namespace gimb

// =================================================================================================
extern "C" {

const char* gimb_last_error(void) { return gimb::last_error(); }
int gimb_abi_version(void) { return GIMB_ABI_VERSION; }

int gimb_loftr_create(const void* blob, size_t nbytes, const gimb_loftr_cfg* cfg, int device, gimb_loftr** out) {
  GIMB_CHECK(blob && out && cfg, "gimb_loftr_create: null argument");
  GIMB_CHECK(nbytes >= sizeof(gimb_blob_header), "weight blob too small");
  GIMB_CHECK(((const gimb_blob_header*)blob)->magic == GIMB_BLOB_MAGIC, "weight blob: bad magic");
  GIMB_CHECK(cfg->fine_window == 5, "only fine_window_size 5 is built (got %d)", cfg->fine_window);
  int ndev = 0;
  GIMB_CUDA(cudaGetDeviceCount(&ndev));
  GIMB_CHECK(device >= 0 && device < ndev, "device %d not available (%d CUDA devices)", device, ndev);
  DeviceGuard guard(device);  // the caller's current device is restored on every exit path
  GIMB_CHECK(guard.ok, "cudaSetDevice(%d) failed", device);
  cudaDeviceProp prop;
  GIMB_CUDA(cudaGetDeviceProperties(&prop, device));
  GIMB_CHECK(prop.major == 10, "libgimb200 is built for sm_100a only; device %d is sm_%d%d", device, prop.major, prop.minor);
  gimb_loftr* m = new gimb_loftr();
  m->device = device;
  m->cfg = *cfg;
  m->sm_count = prop.multiProcessorCount;
  m->ws.device = device;
  m->ws.sm_count = m->sm_count;
  // pass 1 sizes the fp16 weight planes, pass 2 fills them (tcgen05 engine operands)
  Ctx cctx;
  cctx.sm_count = m->sm_count;
  if (m->ws.upload(blob, nbytes) != 0 || build_model(m, cctx) != 0 || m->ws.alloc_planes() != 0 ||
      build_model(m, cctx) != 0 || cudaDeviceSynchronize() != cudaSuccess) {
    gimb_loftr_destroy(m);
    return 1;
  }
  const char* eng = getenv("GIMB_ENGINE");
  if (eng && eng[0] == 's') m->engine = ENGINE_SIMT;
  if (cudaMallocHost(&m->host_count, 2 * sizeof(int64_t)) != cudaSuccess) {
    gimb_loftr_destroy(m);
    set_error("cudaMallocHost failed");
    return 1;
  }
  *out = m;
  return 0;
}

void gimb_loftr_destroy(gimb_loftr* h) {
  if (!h) return;
  DeviceGuard guard(h->device);
  h->ws.release();
  for (auto& kv : h->pe_cache) cudaFree(kv.second);
  if (h->host_count) cudaFreeHost(h->host_count);
  delete h;
}

int gimb_loftr_set_pe(gimb_loftr* h, int hc, int wc, const float* host_pe) {
  GIMB_CHECK(h && host_pe && hc > 0 && wc > 0, "gimb_loftr_set_pe: bad argument");
  DeviceGuard guard(h->device);
  GIMB_CHECK(guard.ok, "cudaSetDevice(%d) failed", h->device);
  auto key = std::make_pair(hc, wc);
  auto it = h->pe_cache.find(key);
  float* d = nullptr;
  if (it != h->pe_cache.end()) {
    d = it->second;
  } else {
    GIMB_CUDA(cudaMalloc(&d, (size_t)hc * wc * 256 * sizeof(float)));
    h->pe_cache[key] = d;
  }
  GIMB_CUDA(cudaMemcpy(d, host_pe, (size_t)hc * wc * 256 * sizeof(float), cudaMemcpyHostToDevice));
  return 0;
}

int gimb_loftr_workspace_bytes(gimb_loftr* h, int n, int h0, int w0, int h1, int w1, size_t* bytes) {
  GIMB_CHECK(h && bytes, "gimb_loftr_workspace_bytes: null argument");
  GIMB_TRY(check_shapes(n, h0, w0, h1, w1));
  Ctx ctx;
  ctx.dry = true;
  ctx.arena.dry = true;
  ctx.sm_count = h->sm_count;
  gimb_loftr_out o = {};
  o.capacity = (int64_t)n * std::min((h0 / 8) * (w0 / 8), (h1 / 8) * (w1 / 8));
  FwdArgs f = {nullptr, nullptr, (const uint8_t*)1, (const uint8_t*)1, nullptr, nullptr, n, h0, w0, h1, w1, &o, nullptr};
  int64_t M = 0;
  GIMB_TRY(forward_impl(ctx, h, f, &M));
  *bytes = ctx.arena.peak + Arena::kAlign;
  return 0;
}

int gimb_loftr_forward(gimb_loftr* h, const float* color0, const float* color1, const uint8_t* mask0,
                       const uint8_t* mask1, const float* scale0, const float* scale1, int n, int h0, int w0, int h1,
                       int w1, void* workspace, size_t workspace_bytes, const gimb_loftr_out* out,
                       const gimb_loftr_taps* taps, int64_t* m_out, void* stream) {
  GIMB_CHECK(h && color0 && color1 && workspace && out && m_out, "gimb_loftr_forward: null argument");
  GIMB_TRY(check_shapes(n, h0, w0, h1, w1));
  GIMB_CHECK((mask0 == nullptr) == (mask1 == nullptr), "mask0 and mask1 must be given together");
  GIMB_CHECK((scale0 == nullptr) == (scale1 == nullptr), "scale0 and scale1 must be given together");
  const int64_t need = (int64_t)n * std::min((h0 / 8) * (w0 / 8), (h1 / 8) * (w1 / 8));
  GIMB_CHECK(out->capacity >= need, "output capacity %lld < n*min(L,S) = %lld", (long long)out->capacity, (long long)need);
  GIMB_CHECK(out->b_ids && out->i_ids && out->j_ids && out->mconf && out->mkpts0_c && out->mkpts1_c && out->mkpts0_f &&
                 out->mkpts1_f && out->expec_f,
             "gimb_loftr_out: every output array is required");
  DeviceGuard guard(h->device);
  GIMB_CHECK(guard.ok, "cudaSetDevice(%d) failed", h->device);
  Ctx ctx;
  ctx.stream = (cudaStream_t)stream;
  ctx.sm_count = h->sm_count;
  uintptr_t base = ((uintptr_t)workspace + Arena::kAlign - 1) / Arena::kAlign * Arena::kAlign;
  ctx.arena.base = (char*)base;
  ctx.arena.cap = workspace_bytes - (base - (uintptr_t)workspace);
  FwdArgs f = {color0, color1, mask0, mask1, scale0, scale1, n, h0, w0, h1, w1, out, taps};
  int rc = forward_impl(ctx, h, f, m_out);
  h->launches += ctx.launches;
  return rc;
}

int gimb_loftr_host_staging_bytes(int n, int h0, int w0, int h1, int w1, int with_mask, int with_scale, size_t* bytes) {
  GIMB_CHECK(bytes, "null argument");
  GIMB_TRY(check_shapes(n, h0, w0, h1, w1));
  size_t b = 0;
  b += align_up((size_t)n * 3 * h0 * w0 * 4, 256) + align_up((size_t)n * 3 * h1 * w1 * 4, 256);
  if (with_mask) b += align_up((size_t)n * (h0 / 8) * (w0 / 8), 256) + align_up((size_t)n * (h1 / 8) * (w1 / 8), 256);
  if (with_scale) b += 2 * align_up((size_t)n * 2 * 4, 256);
  *bytes = b + 256;
  return 0;
}

int gimb_loftr_forward_host(gimb_loftr* h, const float* color0, const float* color1, const uint8_t* mask0,
                            const uint8_t* mask1, const float* scale0, const float* scale1, int n, int h0, int w0,
                            int h1, int w1, void* dev_inputs, size_t dev_inputs_bytes, void* workspace,
                            size_t workspace_bytes, const gimb_loftr_out* dev_out, const gimb_loftr_out* host_out,
                            int64_t* m_out, uint64_t* h2d_bytes, uint64_t* d2h_bytes, void* stream) {
  GIMB_CHECK(h && color0 && color1 && dev_inputs && dev_out && host_out && m_out, "gimb_loftr_forward_host: null argument");
  size_t need = 0;
  GIMB_TRY(gimb_loftr_host_staging_bytes(n, h0, w0, h1, w1, mask0 != nullptr, scale0 != nullptr, &need));
  GIMB_CHECK(dev_inputs_bytes >= need, "dev_inputs too small: %zu < %zu", dev_inputs_bytes, need);
  DeviceGuard guard(h->device);
  GIMB_CHECK(guard.ok, "cudaSetDevice(%d) failed", h->device);
  cudaStream_t st = (cudaStream_t)stream;
  char* p = (char*)(((uintptr_t)dev_inputs + 255) / 256 * 256);
  uint64_t up = 0;
  auto push = [&](const void* src, size_t bytes, const void** dst) -> int {
    GIMB_CUDA(cudaMemcpyAsync(p, src, bytes, cudaMemcpyHostToDevice, st));
    *dst = p;
    p += align_up(bytes, 256);
    up += bytes;
    return 0;
  };
  const void *d_c0, *d_c1, *d_m0 = nullptr, *d_m1 = nullptr, *d_s0 = nullptr, *d_s1 = nullptr;
  GIMB_TRY(push(color0, (size_t)n * 3 * h0 * w0 * 4, &d_c0));
  GIMB_TRY(push(color1, (size_t)n * 3 * h1 * w1 * 4, &d_c1));
  if (mask0) {
    GIMB_TRY(push(mask0, (size_t)n * (h0 / 8) * (w0 / 8), &d_m0));
    GIMB_TRY(push(mask1, (size_t)n * (h1 / 8) * (w1 / 8), &d_m1));
  }
  if (scale0) {
    GIMB_TRY(push(scale0, (size_t)n * 2 * 4, &d_s0));
    GIMB_TRY(push(scale1, (size_t)n * 2 * 4, &d_s1));
  }
  GIMB_TRY(gimb_loftr_forward(h, (const float*)d_c0, (const float*)d_c1, (const uint8_t*)d_m0, (const uint8_t*)d_m1,
                              (const float*)d_s0, (const float*)d_s1, n, h0, w0, h1, w1, workspace, workspace_bytes,
                              dev_out, nullptr, m_out, stream));
  const int64_t M = *m_out;
  GIMB_CHECK(M <= host_out->capacity, "host_out capacity %lld < M = %lld", (long long)host_out->capacity, (long long)M);
  uint64_t down = sizeof(int64_t);
  auto pull = [&](void* dst, const void* src, size_t bytes) -> int {
    if (!dst || bytes == 0) return 0;
    GIMB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, st));
    down += bytes;
    return 0;
  };
  GIMB_TRY(pull(host_out->b_ids, dev_out->b_ids, M * 8));
  GIMB_TRY(pull(host_out->i_ids, dev_out->i_ids, M * 8));
  GIMB_TRY(pull(host_out->j_ids, dev_out->j_ids, M * 8));
  GIMB_TRY(pull(host_out->mconf, dev_out->mconf, M * 4));
  GIMB_TRY(pull(host_out->mkpts0_c, dev_out->mkpts0_c, M * 8));
  GIMB_TRY(pull(host_out->mkpts1_c, dev_out->mkpts1_c, M * 8));
  GIMB_TRY(pull(host_out->mkpts0_f, dev_out->mkpts0_f, M * 8));
  GIMB_TRY(pull(host_out->mkpts1_f, dev_out->mkpts1_f, M * 8));
  GIMB_TRY(pull(host_out->expec_f, dev_out->expec_f, M * 12));
  GIMB_CUDA(cudaStreamSynchronize(st));
  if (h2d_bytes) *h2d_bytes = up;
  if (d2h_bytes) *d2h_bytes = down;
  return 0;
}

int gimb_loftr_host_u8_staging_bytes(int n, int ih0, int iw0, int ih1, int iw1, int h0, int w0, int h1, int w1, int with_scale,
                                     size_t* bytes) {
  GIMB_CHECK(bytes, "null argument");
  GIMB_TRY(check_shapes(n, h0, w0, h1, w1));
  GIMB_CHECK(ih0 > 0 && iw0 > 0 && ih1 > 0 && iw1 > 0 && ih0 <= h0 && iw0 <= w0 && ih1 <= h1 && iw1 <= w1,
             "u8 images must fit inside the padded sizes");
  size_t b = 256;
  b += align_up((size_t)n * ih0 * iw0 * 3, 256) + align_up((size_t)n * ih1 * iw1 * 3, 256);
  b += align_up((size_t)n * 3 * h0 * w0 * 4, 256) + align_up((size_t)n * 3 * h1 * w1 * 4, 256);
  b += align_up((size_t)n * (h0 / 8) * (w0 / 8), 256) + align_up((size_t)n * (h1 / 8) * (w1 / 8), 256);
  if (with_scale) b += 2 * align_up((size_t)n * 2 * 4, 256);
  *bytes = b + 256;
  return 0;
}

namespace {
// Layout of the u8 staging buffer (gimb_loftr_host_u8_staging_bytes): raw bytes, converted fp32 NCHW, masks, scales.
struct U8Staging {
  uint8_t *u0, *u1, *m0, *m1;
  float *c0, *c1, *s0, *s1;
  size_t b0, b1;
  bool padded;
};
U8Staging carve_u8_staging(void* dev_inputs, int n, int ih0, int iw0, int ih1, int iw1, int h0, int w0, int h1, int w1, bool with_scale) {
  char* p = (char*)(((uintptr_t)dev_inputs + 255) / 256 * 256);
  auto carve = [&](size_t bytes) { char* r = p; p += align_up(bytes, 256); return r; };
  U8Staging s;
  s.b0 = (size_t)n * ih0 * iw0 * 3;
  s.b1 = (size_t)n * ih1 * iw1 * 3;
  s.u0 = (uint8_t*)carve(s.b0);
  s.u1 = (uint8_t*)carve(s.b1);
  s.c0 = (float*)carve((size_t)n * 3 * h0 * w0 * 4);
  s.c1 = (float*)carve((size_t)n * 3 * h1 * w1 * 4);
  s.padded = ih0 != h0 || iw0 != w0 || ih1 != h1 || iw1 != w1;  // the loader returns a mask only when it pads
  s.m0 = (uint8_t*)carve((size_t)n * (h0 / 8) * (w0 / 8));
  s.m1 = (uint8_t*)carve((size_t)n * (h1 / 8) * (w1 / 8));
  s.s0 = s.s1 = nullptr;
  if (with_scale) {
    s.s0 = (float*)carve((size_t)n * 8);
    s.s1 = (float*)carve((size_t)n * 8);
  }
  return s;
}
}  // namespace

int gimb_loftr_stage_host_u8(gimb_loftr* h, const uint8_t* img0, int ih0, int iw0, const uint8_t* img1, int ih1, int iw1,
                             const float* scale0, const float* scale1, int n, int h0, int w0, int h1, int w1, void* dev_inputs,
                             size_t dev_inputs_bytes, uint64_t* h2d_bytes, void* stream) {
  GIMB_CHECK(h && img0 && img1 && dev_inputs, "gimb_loftr_stage_host_u8: null argument");
  GIMB_CHECK((scale0 == nullptr) == (scale1 == nullptr), "scale0/scale1 go together");
  size_t need = 0;
  GIMB_TRY(gimb_loftr_host_u8_staging_bytes(n, ih0, iw0, ih1, iw1, h0, w0, h1, w1, scale0 != nullptr, &need));
  GIMB_CHECK(dev_inputs_bytes >= need, "dev_inputs too small: %zu < %zu", dev_inputs_bytes, need);
  DeviceGuard guard(h->device);
  GIMB_CHECK(guard.ok, "cudaSetDevice(%d) failed", h->device);
  cudaStream_t st = (cudaStream_t)stream;
  const U8Staging s = carve_u8_staging(dev_inputs, n, ih0, iw0, ih1, iw1, h0, w0, h1, w1, scale0 != nullptr);
  uint64_t up = s.b0 + s.b1;
  GIMB_CUDA(cudaMemcpyAsync(s.u0, img0, s.b0, cudaMemcpyHostToDevice, st));
  GIMB_CUDA(cudaMemcpyAsync(s.u1, img1, s.b1, cudaMemcpyHostToDevice, st));
  if (scale0) {
    GIMB_CUDA(cudaMemcpyAsync(s.s0, scale0, (size_t)n * 8, cudaMemcpyHostToDevice, st));
    GIMB_CUDA(cudaMemcpyAsync(s.s1, scale1, (size_t)n * 8, cudaMemcpyHostToDevice, st));
    up += (uint64_t)n * 16;
  }
  const long long t0 = (long long)n * 3 * h0 * w0, t1 = (long long)n * 3 * h1 * w1;
  u8_to_nchw_kernel<<<(unsigned)((t0 + 255) / 256), 256, 0, st>>>(s.u0, n, ih0, iw0, s.c0, h0, w0);
  u8_to_nchw_kernel<<<(unsigned)((t1 + 255) / 256), 256, 0, st>>>(s.u1, n, ih1, iw1, s.c1, h1, w1);
  if (s.padded) {
    pad_mask_kernel<<<(n * (h0 / 8) * (w0 / 8) + 255) / 256, 256, 0, st>>>(s.m0, n, h0 / 8, w0 / 8, ih0, iw0);
    pad_mask_kernel<<<(n * (h1 / 8) * (w1 / 8) + 255) / 256, 256, 0, st>>>(s.m1, n, h1 / 8, w1 / 8, ih1, iw1);
  }
  GIMB_LAUNCH_CHECK();
  h->launches += s.padded ? 4 : 2;
  if (h2d_bytes) *h2d_bytes = up;
  return 0;
}

int gimb_loftr_forward_staged_u8(gimb_loftr* h, int ih0, int iw0, int ih1, int iw1, int with_scale, int n, int h0, int w0, int h1,
                                 int w1, void* dev_inputs, size_t dev_inputs_bytes, void* workspace, size_t workspace_bytes,
                                 const gimb_loftr_out* dev_out, const gimb_loftr_out* host_out, int64_t* m_out,
                                 uint64_t* d2h_bytes, void* stream) {
  GIMB_CHECK(h && dev_inputs && dev_out && host_out && m_out, "gimb_loftr_forward_staged_u8: null argument");
  size_t need = 0;
  GIMB_TRY(gimb_loftr_host_u8_staging_bytes(n, ih0, iw0, ih1, iw1, h0, w0, h1, w1, with_scale, &need));
  GIMB_CHECK(dev_inputs_bytes >= need, "dev_inputs too small: %zu < %zu", dev_inputs_bytes, need);
  DeviceGuard guard(h->device);
  GIMB_CHECK(guard.ok, "cudaSetDevice(%d) failed", h->device);
  cudaStream_t st = (cudaStream_t)stream;
  const U8Staging s = carve_u8_staging(dev_inputs, n, ih0, iw0, ih1, iw1, h0, w0, h1, w1, with_scale != 0);
  GIMB_TRY(gimb_loftr_forward(h, s.c0, s.c1, s.padded ? s.m0 : nullptr, s.padded ? s.m1 : nullptr, s.s0, s.s1, n, h0, w0, h1, w1,
                              workspace, workspace_bytes, dev_out, nullptr, m_out, stream));
  const int64_t M = *m_out;
  GIMB_CHECK(M <= host_out->capacity, "host_out capacity %lld < M = %lld", (long long)host_out->capacity, (long long)M);
  uint64_t down = sizeof(int64_t);
  auto pull = [&](void* dst, const void* src, size_t bytes) -> int {
    if (!dst || bytes == 0) return 0;
    GIMB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, st));
    down += bytes;
    return 0;
  };
  GIMB_TRY(pull(host_out->b_ids, dev_out->b_ids, M * 8));
  GIMB_TRY(pull(host_out->i_ids, dev_out->i_ids, M * 8));
  GIMB_TRY(pull(host_out->j_ids, dev_out->j_ids, M * 8));
  GIMB_TRY(pull(host_out->mconf, dev_out->mconf, M * 4));
  GIMB_TRY(pull(host_out->mkpts0_c, dev_out->mkpts0_c, M * 8));
  GIMB_TRY(pull(host_out->mkpts1_c, dev_out->mkpts1_c, M * 8));
  GIMB_TRY(pull(host_out->mkpts0_f, dev_out->mkpts0_f, M * 8));
  GIMB_TRY(pull(host_out->mkpts1_f, dev_out->mkpts1_f, M * 8));
  GIMB_TRY(pull(host_out->expec_f, dev_out->expec_f, M * 12));
  GIMB_CUDA(cudaStreamSynchronize(st));
  if (d2h_bytes) *d2h_bytes = down;
  return 0;
}

int gimb_loftr_forward_host_u8(gimb_loftr* h, const uint8_t* img0, int ih0, int iw0, const uint8_t* img1, int ih1, int iw1,
                               const float* scale0, const float* scale1, int n, int h0, int w0, int h1, int w1, void* dev_inputs,
                               size_t dev_inputs_bytes, void* workspace, size_t workspace_bytes, const gimb_loftr_out* dev_out,
                               const gimb_loftr_out* host_out, int64_t* m_out, uint64_t* h2d_bytes, uint64_t* d2h_bytes,
                               void* stream) {
  GIMB_CHECK(h && img0 && img1 && dev_inputs && dev_out && host_out && m_out, "gimb_loftr_forward_host_u8: null argument");
  GIMB_TRY(gimb_loftr_stage_host_u8(h, img0, ih0, iw0, img1, ih1, iw1, scale0, scale1, n, h0, w0, h1, w1, dev_inputs, dev_inputs_bytes,
                                    h2d_bytes, stream));
  return gimb_loftr_forward_staged_u8(h, ih0, iw0, ih1, iw1, scale0 != nullptr, n, h0, w0, h1, w1, dev_inputs, dev_inputs_bytes, workspace,
                                      workspace_bytes, dev_out, host_out, m_out, d2h_bytes, stream);
}

uint64_t gimb_loftr_launch_count(gimb_loftr* h) { return h ? h->launches : 0; }
uint64_t gimb_loftr_corr_fallbacks(gimb_loftr* h) { return h ? h->corr_fallbacks : 0; }

int gimb_loftr_set_engine(gimb_loftr* h, int engine) {
  GIMB_CHECK(h && (engine == ENGINE_SIMT || engine == ENGINE_TC), "gimb_loftr_set_engine: bad argument");
  h->engine = engine;
  return 0;
}

int gimb_loftr_set_profiling(gimb_loftr* h, int enabled) {
  GIMB_CHECK(h, "null handle");
  h->profiling = enabled != 0;
  return 0;
}

int gimb_loftr_last_profile(gimb_loftr* h, const char** names, float* ms, int* n_stages) {
  GIMB_CHECK(h && names && ms && n_stages, "null argument");
  int n = (int)std::min<size_t>(h->last_profile.size(), 32);
  h->prof_names.resize(n);
  for (int i = 0; i < n; ++i) {
    h->prof_names[i] = h->last_profile[i].first;
    names[i] = h->prof_names[i].c_str();
    ms[i] = h->last_profile[i].second;
  }
  *n_stages = n;
  return 0;
}

}  // extern "C"
