// tcgen05 temporal attention for the 64-channel levels (LayerNorm + QKV projection + rotary + banded attention with relative
// position bias + out-projection + residual, all on chip); see temporal_tc.cu.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <vector>

namespace dawn {

constexpr int kTtcWindowMax = 240;     // frames of one pixel held on chip per work unit (a 200-frame shard plus one 40-frame halo fits)
constexpr int kTtcBandMax = 40;        // |j - i| <= band, band <= 40 (reference win_width, config/DAWN_*.yaml)
constexpr int kTtcMaxSeg = 16;
constexpr int kTtcTable = 512;         // bias/mask table entries per head (index = key - query + kTtcTableZero)
constexpr int kTtcTableZero = 240;

// One work unit = one pixel x one segment.  A segment holds window frames [w0, w0 + wn) of the on-chip sequence (wn <= 240) and produces
// the outputs of query frames [qa, qb); every band key of those queries lies inside the window.
struct TtcSegment { int w0, wn, qa, qb; };

// Geometry of a segment's two row tiles (tile j = window rows [r0, r1) on TMEM lanes 0..; queries [q0, q1) of them; key columns start at
// window row kb).  Shared by the kernel and the host-side coverage test.
struct TtcTile { int r0, r1, q0, q1, kb; };
__host__ __device__ inline void ttc_tiles(const TtcSegment& s, int band, TtcTile (&t)[2]) {
  int R1 = (((s.wn + 1) >> 1) + 7) & ~7;
  if (R1 > 128) R1 = 128;
  const int qa = s.qa - s.w0, qb = s.qb - s.w0;
  for (int j = 0; j < 2; ++j) {
    TtcTile& x = t[j];
    x.r0 = j ? R1 : 0;
    x.r1 = j ? s.wn : (R1 < s.wn ? R1 : s.wn);
    if (x.r1 < x.r0) x.r1 = x.r0;
    x.q0 = qa > x.r0 ? qa : x.r0;
    x.q1 = qb < x.r1 ? qb : x.r1;
    if (x.q1 < x.q0) x.q1 = x.q0;
    int kb = x.q0 - band;
    if (kb < 0) kb = 0;
    kb &= ~15;
    if (kb > kTtcWindowMax - 160) kb = kTtcWindowMax - 160;
    x.kb = kb;
  }
}

struct TemporalTcArgs {
  const float* x; int ldx;        // layer input over F frames (own frames + neighbour halos when sharded): rows f*P + pixel
  const float* res; int ldr;      // residual rows of the owned frames: (f - q_lo)*P + pixel
  float* out; int ldo;            // output rows, same indexing as res
  int F, P;                       // sequence length, pixels per frame
  int q_lo, q_hi;                 // frames [q_lo, q_hi) produce output
  const uint8_t* Wqkv;            // [8 heads] shared-memory images: W'_h hi (96 x 128 B, swizzled) | lo; rows q 32, k 32, v 32
  const uint8_t* Wout;            // [8 heads] images: Wout_h (64 channels x [hi 32 | lo 32] halfs, swizzled)
  const float* rot;               // [F][16][2] cos/sin per frame
  const float* table;             // [8][kTtcTable] bias * log2(e) where |rel| <= band, -1e30 elsewhere
  int band;
  float inv_wscale, inv_oscale;
  int nseg;
  TtcSegment seg[kTtcMaxSeg];
  float* dbg;                     // optional: intermediates of unit 0 / head 0 (selftest)
  unsigned long long* trace;      // optional: 48 cycle counters of CTA 0 (selftest): [0,12) warpgroup 0, [16,28) warpgroup 1, [32,46) MMA issuer
};

bool temporal_tc_supported(int C, int F, int band, int q_lo, int q_hi);
// segments covering queries [q_lo, q_hi) of a sequence of F frames (returns the count, 0 if unsupported)
int temporal_tc_plan(int F, int band, int q_lo, int q_hi, TtcSegment* seg);
int launch_temporal_tc(const TemporalTcArgs& a, cudaStream_t st);
// host packing: wqkv [768][64] fp32 (gamma and q-scale folded), wout [64][256] -> swizzled fp16 hi|lo images
void temporal_tc_pack(const float* wqkv, const float* wout, std::vector<uint8_t>& Wq, std::vector<uint8_t>& Wo, float* inv_wscale,
                      float* inv_oscale);
// host: [8][2*band+1] relative bias -> [8][kTtcTable] table
void temporal_tc_table(const float* bias, int band, std::vector<float>& table);

}  // namespace dawn
