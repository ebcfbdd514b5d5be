// corr_sweep.cuh - host-side description of the second-generation dual-softmax correlation sweeps (corr_sweep.cu).
#pragma once
#include "umma_gemm.cuh"

namespace gimb {

struct CorrSweep {
  SplitPlanes f0, f1;          // [N*L, C] / [N*S, C] split planes (hi, lo), dense (ld == C)
  const float* f0_f32 = nullptr;  // the same features in fp32 (row norms for the exponent reference)
  const float* f1_f32 = nullptr;
  int N = 0, L = 0, S = 0, C = 0;
  const uint8_t* mask0 = nullptr;
  const uint8_t* mask1 = nullptr;
  float temperature = 0.1f, thr = 0.2f;
  // workspace (sizes from corr_sweep_parts): all caller-allocated
  float* normsq = nullptr;     // [2 * N]
  float* rowpart = nullptr;    // [row_parts][N * Lp]
  float* colpart = nullptr;    // [col_parts][N * Sp]
  float2* rowstat = nullptr;   // [N * Lp]
  float* colthr = nullptr;     // [N * Sp]
  float* colsum = nullptr;     // [N * Sp]
  int* flag = nullptr;         // set to 1 when a softmax sum left the safe range: the caller must run the exact sweeps
  unsigned long long* rowbest = nullptr;  // [N * L]
  unsigned int* colbest = nullptr;        // [N * S]
  float* conf_matrix = nullptr;           // optional debug tap [N, L, S]
};

bool corr_sweep_supported(int C);
void corr_sweep_parts(const Ctx& ctx, int nb, int L, int S, int C, int* row_parts, int* col_parts, int* Lp, int* Sp);
int corr_sweeps(Ctx& ctx, const CorrSweep& c);  // norms, sweep 1, merge, sweep 2 (marks corr_stats / corr_merge / corr_conf)

}  // namespace gimb
