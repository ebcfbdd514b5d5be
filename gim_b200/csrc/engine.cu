// engine.cu - weight store and GEMM-layer dispatch shared by loftr_api.cu and dkm_api.cu (see engine.cuh).
#include "engine.cuh"

#include <string.h>

namespace gimb {

int WeightStore::upload(const void* blob, size_t nbytes) {
  GIMB_CHECK(blob, "weight blob: null pointer");
  GIMB_CHECK(nbytes >= sizeof(gimb_blob_header), "weight blob too small");
  const gimb_blob_header* hd = (const gimb_blob_header*)blob;
  GIMB_CHECK(hd->magic == GIMB_BLOB_MAGIC, "weight blob: bad magic");
  GIMB_CHECK(hd->version == 1, "weight blob: unsupported version %u", hd->version);
  GIMB_CHECK(hd->total_bytes <= nbytes && hd->data_offset <= hd->total_bytes, "weight blob: truncated");
  const size_t data_bytes = hd->total_bytes - hd->data_offset;
  dblob_bytes = data_bytes;
  if (cudaMalloc(&dblob, data_bytes) != cudaSuccess) {
    dblob = nullptr;
    set_error("cudaMalloc of %zu weight bytes failed", data_bytes);
    return 1;
  }
  GIMB_CUDA(cudaMemcpy(dblob, (const char*)blob + hd->data_offset, data_bytes, cudaMemcpyHostToDevice));
  const gimb_blob_entry* ent = (const gimb_blob_entry*)((const char*)blob + sizeof(gimb_blob_header));
  GIMB_CHECK(sizeof(gimb_blob_header) + (size_t)hd->n_entries * sizeof(gimb_blob_entry) <= hd->data_offset,
             "weight blob: entry table overlaps the data");
  for (uint32_t i = 0; i < hd->n_entries; ++i) {
    GIMB_CHECK(ent[i].ndim <= 4, "weight blob: entry %u has %u dims", i, ent[i].ndim);
    std::vector<uint32_t> sh(ent[i].shape, ent[i].shape + ent[i].ndim);
    GIMB_CHECK(ent[i].offset + ent[i].nbytes <= data_bytes, "weight blob: entry %u out of range", i);
    std::string name(ent[i].name, strnlen(ent[i].name, sizeof(ent[i].name)));
    tensors[name] = {(const float*)(dblob + ent[i].offset), sh};
  }
  return 0;
}

int WeightStore::alloc_planes() {
  dplanes_bytes = dplanes_top + 256;
  if (cudaMalloc(&dplanes, dplanes_bytes) != cudaSuccess) {
    dplanes = nullptr;
    set_error("cudaMalloc of %zu weight-plane bytes failed", dplanes_bytes);
    return 1;
  }
  GIMB_CUDA(cudaMemset(dplanes, 0, dplanes_bytes));
  dplanes_top = 0;
  return 0;
}

void WeightStore::release() {
  if (dblob) cudaFree(dblob);
  if (dplanes) cudaFree(dplanes);
  dblob = dplanes = nullptr;
}

int WeightStore::find(const std::string& name, const float** out, std::vector<uint32_t>* shape) const {
  auto it = tensors.find(name);
  GIMB_CHECK(it != tensors.end(), "weight blob: tensor '%s' missing", name.c_str());
  *out = it->second.first;
  if (shape) *shape = it->second.second;
  return 0;
}

int WeightStore::make_weight_planes(Ctx& ctx, Wt* wt, int cout, int taps, int cin) {
  wt->ldk = pitch8(cin);
  const size_t n = (size_t)cout * taps * wt->ldk;
  __half* ptr[2];
  for (int i = 0; i < 2; ++i) {
    dplanes_top = (dplanes_top + 255) / 256 * 256;
    ptr[i] = dplanes ? (__half*)(dplanes + dplanes_top) : nullptr;
    dplanes_top += n * sizeof(__half);
  }
  if (!dplanes) return 0;
  wt->wp.hi = ptr[0]; wt->wp.lo = ptr[1];
  SplitPlanes tap_view = wt->wp;
  tap_view.ld = wt->ldk;  // rows of cin values, pitch ldk
  GIMB_TRY(split_planes(ctx, wt->w, (int64_t)cout * taps, cin, cin, tap_view));
  wt->wp.ld = taps * wt->ldk;
  return 0;
}

int WeightStore::load_conv(Ctx& ctx, const std::string& name, bool affine, Conv* c) {
  std::vector<uint32_t> sh;
  GIMB_TRY(find(name + ".w", &c->wt.w, &sh));
  GIMB_CHECK(sh.size() == 4 && sh[1] == sh[2], "conv '%s': expected [Cout,k,k,Cin]", name.c_str());
  c->cout = sh[0]; c->k = sh[1]; c->cin = sh[3];
  if (affine) {
    GIMB_TRY(find(name + ".s", &c->s));
    GIMB_TRY(find(name + ".b", &c->b));
  }
  if (c->cin % 4 == 0 || c->k == 1) GIMB_TRY(make_weight_planes(ctx, &c->wt, c->cout, c->k * c->k, c->cin));
  return 0;
}

int WeightStore::load_linear(Ctx& ctx, const std::string& name, Wt* wt) {
  std::vector<uint32_t> sh;
  GIMB_TRY(find(name, &wt->w, &sh));
  GIMB_CHECK(sh.size() == 2, "linear '%s': expected [out,in]", name.c_str());
  return make_weight_planes(ctx, wt, sh[0], 1, sh[1]);
}

ActT Fwd::alloc(size_t rows, int C, bool want_f32, bool want_split, bool want_h8, bool padded) {
  ActT a;
  a.C = C;
  if (!tc()) { want_f32 = true; want_split = false; }
  if (want_f32) {
    a.ldf = padded ? pitch8(C) : 0;
    a.f32 = ctx.arena.alloc<float>(rows * (size_t)a.pitch());
  }
  if (want_split) {
    a.sp.ld = pitch8(C);
    a.sp.hi = ctx.arena.alloc<__half>(rows * a.sp.ld);
    a.sp.lo = ctx.arena.alloc<__half>(rows * a.sp.ld);
    if (want_h8) a.sp.h8 = ctx.arena.alloc<__half>(rows * a.sp.ld);
  }
  return a;
}

ActT view_rows(const ActT& a, size_t row0) {
  ActT v = a;
  if (a.f32) v.f32 = a.f32 + row0 * a.pitch();
  if (a.sp.hi) {
    v.sp.hi = a.sp.hi + row0 * a.sp.ld;
    v.sp.lo = a.sp.lo + row0 * a.sp.ld;
    if (a.sp.h8) v.sp.h8 = a.sp.h8 + row0 * a.sp.ld;
  }
  return v;
}

int gemm(Fwd& F, const Wt& wt, int cin1, int cin2, int cout, int k, int stride, const ActT& in, const ActT* in2, int B,
         int H, int W, const Epi& e, const ActT& out) {
  const int pad = k / 2;
  const int OH = (H + 2 * pad - k) / stride + 1, OW = (W + 2 * pad - k) / stride + 1;
  if (!F.tc()) {
    GIMB_CHECK(in.pitch() == in.C && out.pitch() == out.C, "gemm: the CUDA-core engine needs dense fp32 tensors");
    ConvGemm g;
    g.in = in.f32; g.in2 = in2 ? in2->f32 : nullptr;
    g.B = B; g.H = H; g.W = W; g.C1 = cin1; g.C2 = cin2;
    g.KH = g.KW = k; g.stride = stride; g.pad = pad; g.OH = OH; g.OW = OW;
    g.w = wt.w; g.Cout = cout; g.scale = e.scale; g.bias = e.bias; g.residual = e.residual; g.row_mask = e.row_mask;
    g.act0 = e.act0; g.act1 = e.act1; g.act_split = e.act_split; g.div = e.div; g.out = out.f32;
    return conv_gemm(F.ctx, g);
  }
  UmmaGemm g;
  g.a = in.sp;
  if (in2) g.a2 = in2->sp;
  g.b = wt.wp; g.N = cout;
  if (k == 1 && stride == 1) {
    g.mode = 0; g.M = (int64_t)B * H * W; g.K1 = cin1; g.K2 = cin2;
  } else {
    g.mode = 1; g.K1 = cin1; g.B = B; g.H = H; g.W = W; g.KH = g.KW = k; g.stride = stride; g.pad = pad;
    g.OH = OH; g.OW = OW; g.ldk = wt.ldk;
  }
  g.scale = e.scale; g.bias = e.bias; g.residual = e.residual; g.row_mask = e.row_mask;
  if (e.residual_planes) g.residual_planes = *e.residual_planes;
  g.act0 = e.act0; g.act1 = e.act1; g.act_split = e.act_split; g.div = e.div;
  g.layernorm = e.layernorm;
  g.out_f32 = out.f32; g.out_f32_ld = out.f32 ? out.pitch() : 0; g.residual_ld = g.out_f32_ld;
  g.out = out.sp;
  return umma_gemm(F.ctx, g);
}

int run_conv(Fwd& F, const Conv& c, const ActT& in, int B, int H, int W, int stride, int act, const float* residual,
             const ActT& out, const SplitPlanes* residual_planes) {
  Epi e;
  e.scale = c.s; e.bias = c.b; e.residual = residual; e.residual_planes = residual_planes; e.act0 = e.act1 = act;
  return gemm(F, c.wt, c.cin, 0, c.cout, c.k, stride, in, nullptr, B, H, W, e, out);
}

}  // namespace gimb
