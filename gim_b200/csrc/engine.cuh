// engine.cuh - model-agnostic pieces shared by the gim_loftr and gim_dkm orchestration code: the packed weight blob on
// the device, GEMM-shaped layers in both engine formats, activation tensors (fp32 and/or split fp16 planes) and the
// one-call wrapper that runs a convolution / Linear layer on the selected engine.
#pragma once
#include <map>
#include <string>
#include <utility>
#include <vector>

#include "../../include/gimb200.h"
#include "ops.cuh"

namespace gimb {

enum Engine : int { ENGINE_SIMT = 0, ENGINE_TC = 1 };

// weights of one GEMM-shaped layer in both engine formats
struct Wt {
  const float* w = nullptr;   // fp32 [cout][k*k*cin]            (CUDA-core engine)
  SplitPlanes wp;             // fp16 planes [cout][k*k*ldk]      (tcgen05 engine)
  int ldk = 0;                // per-tap K pitch of the planes
};
struct Conv {
  Wt wt;
  const float* s = nullptr;  // per-channel scale (folded BN; ones for a plain bias), null: no affine
  const float* b = nullptr;
  int cout = 0, cin = 0, k = 1;
};

// channel pitch of fp16 planes: multiples of 32 elements (64 B) so that every TMA box row (32 channels) is one
// aligned 64-byte segment
inline int pitch8(int c) { return (c + 31) / 32 * 32; }

// The packed weight blob (include/gimb200.h: header | entries | fp32 tensors) on the device, plus the fp16 planes of
// every GEMM layer.  Loading is two passes over the same model-building code: the first sizes the plane buffer
// (`dplanes == nullptr`), the second fills it.
struct WeightStore {
  int device = 0;
  int sm_count = 148;
  char* dblob = nullptr;
  size_t dblob_bytes = 0;
  char* dplanes = nullptr;
  size_t dplanes_bytes = 0, dplanes_top = 0;
  std::map<std::string, std::pair<const float*, std::vector<uint32_t>>> tensors;

  int upload(const void* blob, size_t nbytes);  // validates the header, copies the tensor data, fills `tensors`
  int alloc_planes();                           // after the sizing pass
  void release();
  bool has(const std::string& name) const { return tensors.count(name) != 0; }
  int find(const std::string& name, const float** out, std::vector<uint32_t>* shape = nullptr) const;
  // carve fp16 planes (hi, lo) for a weight [rows = cout*taps][cin] and fill them (sizing pass: only count)
  int make_weight_planes(Ctx& ctx, Wt* wt, int cout, int taps, int cin);
  int load_conv(Ctx& ctx, const std::string& name, bool affine, Conv* c);  // "<name>.w" [Cout,k,k,Cin] (+ ".s", ".b")
  int load_linear(Ctx& ctx, const std::string& name, Wt* wt);              // "<name>" [out,in]
};

// An activation tensor [rows, C]: fp32 (pitch `ldf`, default C) and/or split fp16 planes (pitch pitch8(C)).
struct ActT {
  float* f32 = nullptr;
  SplitPlanes sp;
  int C = 0;
  int ldf = 0;  // fp32 row pitch in elements; 0 = C
  int pitch() const { return ldf ? ldf : C; }
  const SplitPlanes* planes() const { return sp.hi ? &sp : nullptr; }
};

struct Fwd {  // per-forward context
  Ctx& ctx;
  int engine;
  bool tc() const { return engine == ENGINE_TC; }
  // Allocate an activation.  SIMT engine: always fp32 only.  TC engine: as requested.  `padded`: the fp32 tensor gets
  // the plane pitch (needed when C * 4 bytes is not a multiple of 16: TMA strides).
  ActT alloc(size_t rows, int C, bool want_f32, bool want_split, bool want_h8 = false, bool padded = false);
};

ActT view_rows(const ActT& a, size_t row0);

struct Epi {
  const float* scale = nullptr;
  const float* bias = nullptr;
  const float* residual = nullptr;
  const SplitPlanes* residual_planes = nullptr;  // tcgen05 engine: identity carried as fp16 planes
  const uint8_t* row_mask = nullptr;
  int act0 = ACT_NONE, act1 = ACT_NONE, act_split = 1 << 30;
  float div = 1.f;
  bool layernorm = false;  // tcgen05 engine only: LayerNorm fused into the epilogue (scale/bias = gamma/beta)
};

// one GEMM-shaped layer on the selected engine.  in2: channel concat (1x1 only).
int gemm(Fwd& F, const Wt& wt, int cin1, int cin2, int cout, int k, int stride, const ActT& in, const ActT* in2, int B,
         int H, int W, const Epi& e, const ActT& out);
int run_conv(Fwd& F, const Conv& c, const ActT& in, int B, int H, int W, int stride, int act, const float* residual,
             const ActT& out, const SplitPlanes* residual_planes = nullptr);

}  // namespace gimb
