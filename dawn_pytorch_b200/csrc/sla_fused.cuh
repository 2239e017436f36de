// Fused spatial-linear-attention context (K/V projection + softmax over pixels + k^T v + out-projection compose); see sla_fused.cu.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <vector>

namespace dawn {

struct SlaCtxArgs {
  const float* x; int ldx;        // layer input, rows f*P + pixel, 64 channels
  int F, P;
  const uint16_t* Wkv;            // [hi|lo][8 heads x (k 32 | v 32)][64] fp16, LayerNorm gain folded, pre-scaled by 1/inv_wscale
  float inv_wscale;
  float* part;                    // scratch: sla_fused_part_floats(F, P) floats
  int px_per_cta;                 // set by the launcher
};

struct SlaOutArgs {
  const float* x; int ldx;        // layer input (also the residual)
  float* out; int ldo;            // may alias x
  int F, P;
  const uint16_t* Wq;             // [hi|lo][256][64] fp16 q projection, LayerNorm gain folded, pre-scaled by 1/inv_wscale
  float inv_wscale;
  const float* Bf; int ldb;       // [F][256][ldb] per-frame (context x out-projection) matrices
  const float* bias;              // [64]
  int px_per_cta;                 // set by the launcher
};
int launch_sla_out_fused(const SlaOutArgs& a, cudaStream_t st);
void sla_out_pack(const float* wqkv_folded, std::vector<uint16_t>& W, float* inv_wscale);

bool sla_fused_supported(int C, int P);
size_t sla_fused_part_floats(int F, int P);
// Bf[f][256][ldb] = per-frame (context x out-projection) matrices, as launch_sla_context produces
int launch_sla_ctx_fused(const SlaCtxArgs& a, const float* WoutT, float* Bf, int ldb, cudaStream_t st);
void sla_fused_pack(const float* wqkv_folded, std::vector<uint16_t>& W, float* inv_wscale);

}  // namespace dawn
