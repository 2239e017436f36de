// tcgen05 3x3 convolution with a shared-memory HALO tile (sm_100a) — the Block.proj of the reference (U:229, 234).
//
// tc_gemm.cu treats a 3x3 conv as 9 independent K panels per 64 channels: the same activations are gathered, split to
// fp16 hi/lo and stored 9 times.  Here a CTA stages the (16+2) x (8+2) pixel halo of its 16x8-pixel output tile ONCE per
// 64-channel chunk (6.3x less gather/convert/store work) and the 9 taps are 9 shifted operand windows over that tile:
// measured on B200, with descriptor base_offset = 0 the UMMA 128-byte swizzle follows ABSOLUTE shared-memory address bits,
// so a K-major operand may start at any 128-byte row and use any row-multiple stride between its 8-row groups:
//     window(dy,dx): start = halo + ((dy+1)*10 + (dx+1))*128 B,  stride between 8-pixel rows (SBO) = 10*128 B.
// Everything else follows tc_gemm.cu: FP16x3 split precision, TMEM double-buffered accumulators drained into RN fp32
// registers (once per 64-channel chunk = 108 MMAs), persistent warp-specialised CTA (8 producer warps, 4/8 epilogue warps,
// MMA issuer, weight loader), row-per-thread epilogue with bias + GroupNorm partial statistics.
#include <cuda.h>
#include <cuda_fp16.h>
#include <cstdlib>
#include <cstring>
#include <string>
#include <map>
#include <tuple>
#include "common.cuh"
#include "gemm.cuh"
#include "tc_common.cuh"
#include "tc_gemm.cuh"

namespace dawn {
namespace {

using namespace tc;

constexpr int TH = 16, TW = 8;                 // output tile (pixels) -> M = 128 rows, row m = y*8 + x
constexpr int HH = TH + 2, HW = TW + 2;        // halo tile
constexpr int HROWS = HH * HW;                 // 180 halo pixels = 180 rows of 128 B
constexpr int A_HALO = 23 * 1024;              // 180 * 128 = 23040 B, padded to 23 KB (keeps 1024-byte alignment)
constexpr int NPROD = 256;

template <int BN>
struct CCfg {
  static constexpr int NWG = BN / 64;
  static constexpr int B_PANEL = BN * 128;                     // one (tap, chunk) weight panel, hi or lo
  static constexpr int A_STAGES = 2;
  static constexpr int B_STAGES = (BN == 64) ? 4 : 3;
  static constexpr int A_BYTES = A_STAGES * 2 * A_HALO;        // hi + lo
  static constexpr int B_BYTES = B_STAGES * 2 * B_PANEL;
  static constexpr int EPI_STAGE = NWG * 4 * 32 * 20 * 4;
  static constexpr int SMEM_DYN = A_BYTES + B_BYTES + EPI_STAGE + 1024;
  static constexpr int NTHREADS = NPROD + 128 * NWG + 64;
  static constexpr int MMA_WARP = (NPROD + 128 * NWG) / 32;
  static constexpr int LOAD_WARP = MMA_WARP + 1;
  // BN == 64: one N=128 MMA multiplies A_hi with [B_hi | B_lo] (two 64-column accumulators, summed by the epilogue):
  // 2 instead of 3 instructions per k-step, and the N=64 instruction was issue-bound at about the cost of an N=128 one.
  static constexpr int ACC_COLS = (BN == 64) ? 128 : BN;
  static constexpr int TMEM_COLS = 2 * ACC_COLS;
};

// A-operand source: TMA = false: fp32 activations, gathered / split / swizzled by the 8 producer warps.  TMA = true: the activation exists as
// two dense fp16 planes (hi | lo, written by the producing kernel's epilogue) and ONE thread fetches the halo tile of a 64-channel chunk with
// two cp.async.bulk.tensor loads (4-D tiled map {C, W, H, F}, box {64, 10, 18, 1}, 128-byte swizzle, out-of-image rows zero-filled by the
// TMA unit): the smem image is byte-identical to what the producers write (row = hy * 10 + hx of 128 B, absolute-address swizzle).
template <int BN, bool TMA, bool TR>
__global__ void __launch_bounds__(CCfg<BN>::NTHREADS, 1) tc_conv3_kernel(const GemmParams p, const float* __restrict__ Bimg,
                                                                          int tiles_y, int tiles_x, int tiles_n,
                                                                          const __grid_constant__ CUtensorMap tm_hi,
                                                                          const __grid_constant__ CUtensorMap tm_lo) {
  using C = CCfg<BN>;
  constexpr int B_PANEL = C::B_PANEL, A_STAGES = C::A_STAGES, B_STAGES = C::B_STAGES, NWG = C::NWG;
  constexpr int MMA_WARP = C::MMA_WARP, LOAD_WARP = C::LOAD_WARP, TMEM_COLS = C::TMEM_COLS, ACC_COLS = C::ACC_COLS;
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t a_full[A_STAGES], a_free[A_STAGES], b_full[B_STAGES], b_free[B_STAGES], acc_full[2], acc_free[2];
  __shared__ uint32_t s_tmem_base;
  __shared__ float s_stat[NWG][16];
  __shared__ __align__(16) float s_bias[NWG][64];       // bias of this epilogue warpgroup's 64 columns (single n-tile: constant for the whole launch)

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);          // warp-uniform by construction (the MMA warp relies on it)
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* smemA = smem;
  uint8_t* smemB = smem + C::A_BYTES;
  uint8_t* smemE = smemB + C::B_BYTES;

  if (tid == 0) {
    for (int s = 0; s < A_STAGES; ++s) { mbar_init(&a_full[s], TMA ? 1 : NPROD); mbar_init(&a_free[s], 1); }
    for (int s = 0; s < B_STAGES; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_free[s], 1); }
    mbar_init(&acc_full[0], 1); mbar_init(&acc_full[1], 1);
    mbar_init(&acc_free[0], 128 * NWG); mbar_init(&acc_free[1], 128 * NWG);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == MMA_WARP) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem_base)), "r"(TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = s_tmem_base;

  const int H = p.IH, W = p.IW;
  const int F = p.M / (H * W);
  const int NCH = p.Cin / 64;                                  // 64-channel chunks
  const int tiles_sp = tiles_y * tiles_x;
  const int total_tiles = F * tiles_sp * tiles_n;              // n fastest: CTAs of one patch share its activations in L2
  auto decode = [&](int tile, int& f, int& y0, int& x0, int& nt) {
    nt = tile % tiles_n;
    const int sp = tile / tiles_n;
    f = sp / tiles_sp;
    const int r = sp - f * tiles_sp;
    y0 = (r / tiles_x) * TH; x0 = (r % tiles_x) * TW;
  };

  if (TMA && warp < 8) {
    // =============================================================== TMA producer: one thread, two tensor loads per (tile, chunk)
    if (tid == 0) {
      const uint64_t mh = reinterpret_cast<uint64_t>(&tm_hi), ml = reinterpret_cast<uint64_t>(&tm_lo);
      uint32_t ait = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        int f, y0, x0, nt;
        decode(tile, f, y0, x0, nt);
        for (int cc = 0; cc < NCH; ++cc, ++ait) {
          const int s = ait % A_STAGES;
          mbar_wait(&a_free[s], ((ait / A_STAGES) & 1) ^ 1);
          mbar_arrive_expect_tx(&a_full[s], 2 * HROWS * 128);
          const uint32_t dst = smem_u32(smemA + s * 2 * A_HALO), bar = smem_u32(&a_full[s]);
          const int c0 = cc * 64, c1 = x0 - 1, c2 = y0 - 1;
          asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                       ::"r"(dst), "l"(mh), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(f) : "memory");
          asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                       ::"r"(dst + A_HALO), "l"(ml), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(f) : "memory");
        }
      }
    }
  } else if (warp < 8) {
    // =============================================================== producers: halo tile of one 64-channel chunk
    const int c16 = tid & 7;
    const int r0 = tid >> 3;                                   // halo rows r0 + 32 q, q = 0..5 (< 180)
    uint32_t ait = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      int f, y0, x0, nt;
      decode(tile, f, y0, x0, nt);
      const float* img = p.A + (size_t)f * H * W * p.lda;
      for (int cc = 0; cc < NCH; ++cc, ++ait) {
        const int s = ait % A_STAGES;
        const uint32_t round = ait / A_STAGES;
        float4 v[12];
#pragma unroll
        for (int q = 0; q < 6; ++q) {
          const int r = r0 + 32 * q;
          const int hy = r / HW, hx = r - hy * HW;
          const int iy = y0 + hy - 1, ix = x0 + hx - 1;
          const bool ok = (r < HROWS) && (iy >= 0) && (iy < H) && (ix >= 0) && (ix < W);
          if (ok) {
            const float4* src = reinterpret_cast<const float4*>(img + (size_t)(iy * W + ix) * p.lda + cc * 64) + 2 * c16;
            v[2 * q] = __ldg(src);
            v[2 * q + 1] = __ldg(src + 1);
          } else {
            v[2 * q] = make_float4(0.f, 0.f, 0.f, 0.f);
            v[2 * q + 1] = make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
        mbar_wait(&a_free[s], (round & 1) ^ 1);
        uint8_t* a_hi = smemA + s * 2 * A_HALO;
        uint8_t* a_lo = a_hi + A_HALO;
#pragma unroll
        for (int q = 0; q < 6; ++q) {
          const int r = r0 + 32 * q;
          if (r < HROWS) {
            uint32_t h[4], l[4];
            split_f16x2(v[2 * q].x, v[2 * q].y, h[0], l[0]);
            split_f16x2(v[2 * q].z, v[2 * q].w, h[1], l[1]);
            split_f16x2(v[2 * q + 1].x, v[2 * q + 1].y, h[2], l[2]);
            split_f16x2(v[2 * q + 1].z, v[2 * q + 1].w, h[3], l[3]);
            const uint32_t off = swz(r, c16);                  // absolute-row swizzle (halo base is 1024-byte aligned)
            *reinterpret_cast<uint4*>(a_hi + off) = make_uint4(h[0], h[1], h[2], h[3]);
            *reinterpret_cast<uint4*>(a_lo + off) = make_uint4(l[0], l[1], l[2], l[3]);
          }
        }
        mbar_arrive_relaxed(&a_full[s]);                       // proxy fence runs on the consumer side (see tc_gemm.cu)
      }
    }
  } else if (warp == LOAD_WARP) {
    // =============================================================== weight loader: one (tap, chunk) panel pair per slot
    if (lane == 0) {
      uint32_t bit = 0;
      const int KC = 9 * NCH;                                  // panels per n-tile in the weight image: kc = tap*NCH + cc
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        int f, y0, x0, nt;
        decode(tile, f, y0, x0, nt);
        const uint8_t* src = reinterpret_cast<const uint8_t*>(Bimg) + (size_t)nt * KC * (2 * B_PANEL);
        for (int cc = 0; cc < NCH; ++cc)
          for (int tap = 0; tap < 9; ++tap, ++bit) {
            const int s = bit % B_STAGES;
            const uint32_t round = bit / B_STAGES;
            mbar_wait(&b_free[s], (round & 1) ^ 1);
            mbar_arrive_expect_tx(&b_full[s], 2 * B_PANEL);
            bulk_copy_g2s(smemB + s * 2 * B_PANEL, src + (size_t)(tap * NCH + cc) * (2 * B_PANEL), 2 * B_PANEL, &b_full[s]);
          }
      }
    }
  } else if (warp == MMA_WARP) {
    // =============================================================== MMA issuer: the whole warp, warp-uniform control flow, one elected
    // lane per MMA / commit (tc_common.cuh: elect_one)
    {
      const uint32_t tmem_base_u = __shfl_sync(0xffffffffu, tmem_base, 0);      // read from shared memory: make it a provably uniform value
      const uint32_t idesc = (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      const uint32_t idesc2 = (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)((2 * BN) >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      constexpr uint64_t SBO_HALO = (uint64_t)(HW * 128 / 16);   // 10 pixel rows of 128 B between 8-row groups
      uint32_t ait = 0, bit = 0, cg = 0;
      const int dt = (p.drain == 1 || p.drain == 3) ? p.drain : 9;   // taps accumulated in TMEM before a drain
      const bool tr = TR && (p.trace != nullptr) && blockIdx.x == 0;
      long long t_acc = 0, t_a = 0, t_b = 0, t_issue = 0, t0 = 0;
      const long long t_begin = clock64();
      uint32_t ntiles_done = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++ntiles_done) {
        for (int cc = 0; cc < NCH; ++cc, ++ait) {
          const int sa = ait % A_STAGES;
          if (tr) t0 = clock64();
          mbar_wait(&a_full[sa], (ait / A_STAGES) & 1);
          if (tr) { const long long t1 = clock64(); t_a += t1 - t0; t0 = t1; }
          fence_proxy_async();
          tc_fence_after();
          const uint32_t a_hi = smem_u32(smemA + sa * 2 * A_HALO), a_lo = a_hi + A_HALO;
          uint32_t buf = 0, d = 0;
          int in_group = 0;
          for (int tap = 0; tap < 9; ++tap, ++bit) {
            if (in_group == 0) {
              buf = cg & 1;
              if (tr) t0 = clock64();
              mbar_wait(&acc_free[buf], ((cg >> 1) & 1) ^ 1);
              if (tr) { const long long t1 = clock64(); t_acc += t1 - t0; t0 = t1; }
              tc_fence_after();
              d = tmem_base_u + buf * ACC_COLS;
            }
            const int sb = bit % B_STAGES;
            if (tr) t0 = clock64();
            mbar_wait(&b_full[sb], (bit / B_STAGES) & 1);
            if (tr) { const long long t1 = clock64(); t_b += t1 - t0; t0 = t1; }
            tc_fence_after();
            const int ky = tap / 3, kx = tap - ky * 3;          // window start: halo pixel (ky, kx)
            const uint32_t woff = (uint32_t)((ky * HW + kx) * 128);
            // K-major SW128 descriptors: start address, LBO = 1 (unused), SBO, version 1, layout SWIZZLE_128B, base_offset 0
            const uint64_t ahi = (uint64_t)(((a_hi + woff) >> 4) & 0x3FFF) | (1ull << 16) | (SBO_HALO << 32) | (1ull << 46) | (2ull << 61);
            const uint64_t alo = (uint64_t)(((a_lo + woff) >> 4) & 0x3FFF) | (1ull << 16) | (SBO_HALO << 32) | (1ull << 46) | (2ull << 61);
            const uint32_t sbaddr = smem_u32(smemB + sb * 2 * B_PANEL);
            const uint64_t bhi = make_desc(sbaddr), blo = make_desc(sbaddr + B_PANEL);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const uint64_t o = (uint64_t)(j * 2);
              if (BN == 64) {
                // the lo panel follows the hi panel in the stage: rows 64..127 of one N=128 operand
                tc_mma_f16_elected(d, ahi + o, bhi + o, idesc2, (in_group == 0 && j == 0) ? 0u : 1u);
                tc_mma_f16_elected(d, alo + o, bhi + o, idesc, 1u);
              } else {
                tc_mma_f16_elected(d, alo + o, bhi + o, idesc, (in_group == 0 && j == 0) ? 0u : 1u);
                tc_mma_f16_elected(d, ahi + o, blo + o, idesc, 1u);
                tc_mma_f16_elected(d, ahi + o, bhi + o, idesc, 1u);
              }
            }
            tc_commit_elected(&b_free[sb]);
            if (++in_group == dt) { tc_commit_elected(&acc_full[buf]); ++cg; in_group = 0; }
            if (tr) t_issue += clock64() - t0;
          }
          tc_commit_elected(&a_free[sa]);
        }
      }
      if (tr && lane == 0) {
        p.trace[0] = (unsigned long long)(clock64() - t_begin); p.trace[1] = ntiles_done;
        p.trace[2] = (unsigned long long)t_acc; p.trace[3] = (unsigned long long)t_a;
        p.trace[4] = (unsigned long long)t_b; p.trace[5] = (unsigned long long)t_issue;
      }
    }
  } else {
    // =============================================================== accumulate + epilogue
    const int wg = (warp - 8) >> 2, ew = (warp - 8) & 3;
    const int row_in_tile = ew * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(ew * 32) << 16;
    const int etid = (tid - NPROD) & 127;
    float* s_st = s_stat[wg];
    const int bar_id = 2 + wg;
    float* wbuf = reinterpret_cast<float*>(smemE) + ((wg * 4 + ew) * 32 * 20);
    uint32_t cg = 0;
    // GroupNorm partial sums (U:230).  Reducing them per tile (16 warp reductions, shared and global atomics, two barriers) cost 7.2k of the
    // 12.2k cycles a 64 -> 64 tile takes (cycle trace, profiles/r2_f_conv3_trace.md).  With a single n-tile every epilogue thread owns the same
    // 64 columns for the whole persistent loop, so it keeps fp32 running sums per 8-column block and the warps reduce them (in fp64) only
    // every 8 tiles and at the end: at most 64 values per fp32 partial sum.
    const bool defer_stats = (p.stats != nullptr) && tiles_n == 1;
    float gs[8], gss[8];
    int pending = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) { gs[i] = 0.f; gss[i] = 0.f; }
    // with one n-tile the 64 bias values never change: fetch them once instead of 16 L2 round trips per tile (1.0-1.7k cycles per tile in the trace)
    const bool bias_smem = (p.bias != nullptr) && tiles_n == 1;
    if (bias_smem) {
      if (etid < 64) s_bias[wg][etid] = __ldg(p.bias + wg * 64 + etid);
      asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
    }
    auto flush_stats = [&]() {
      const int n0f = wg * 64;                                    // tiles_n == 1: this thread's columns never change
#pragma unroll
      for (int b8 = 0; b8 < 8; ++b8) {
        double s = (double)gs[b8], ss = (double)gss[b8];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); ss += __shfl_xor_sync(0xffffffffu, ss, o); }
        if (lane == 0) {
          const int grp = (n0f + b8 * 8) / p.cpg;
          atomicAdd(&p.stats[2 * grp], s);
          atomicAdd(&p.stats[2 * grp + 1], ss);
        }
        gs[b8] = 0.f; gss[b8] = 0.f;
      }
      pending = 0;
    };
    const bool tre = TR && (p.trace != nullptr) && blockIdx.x == 0 && etid == 0 && wg == 0;
    long long te_wait = 0, te_drain = 0, te_final = 0, te0 = 0, te_bias = 0, te_store = 0, te1 = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      int f, y0, x0, nt;
      decode(tile, f, y0, x0, nt);
      const int n0 = nt * BN + wg * 64;
      float acc[64];
#pragma unroll
      for (int i = 0; i < 64; ++i) acc[i] = 0.f;
      const int ndrain = NCH * ((p.drain == 1 || p.drain == 3) ? 9 / p.drain : 1);
      for (int cc = 0; cc < ndrain; ++cc, ++cg) {
        const uint32_t buf = cg & 1;
        if (tre) te0 = clock64();
        mbar_wait(&acc_full[buf], (cg >> 1) & 1);
        if (tre) { const long long t1 = clock64(); te_wait += t1 - te0; te0 = t1; }
        tc_fence_after();
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float v[16];
          tmem_ld16(tmem_base + lane_addr + buf * ACC_COLS + wg * 64 + q * 16, v);
#pragma unroll
          for (int i = 0; i < 16; ++i) acc[q * 16 + i] += v[i];
          if (BN == 64) {                                         // hi*lo partial products live in the second 64 columns
            tmem_ld16(tmem_base + lane_addr + buf * ACC_COLS + 64 + q * 16, v);
#pragma unroll
            for (int i = 0; i < 16; ++i) acc[q * 16 + i] += v[i];
          }
        }
        tc_fence_before();
        mbar_arrive(&acc_free[buf]);
        if (tre) te_drain += clock64() - te0;
      }
      if (tre) te0 = clock64();
#pragma unroll
      for (int i = 0; i < 64; ++i) acc[i] *= p.tc_scale;
      const int oy = y0 + (row_in_tile >> 3), ox = x0 + (row_in_tile & 7);
      const bool rv = (oy < H) && (ox < W);
      size_t opix = (size_t)f * H * W + (size_t)(rv ? oy * W + ox : 0);
      int ocol = n0;
      if (p.up2) {
        // transposed conv as one 3x3 conv with 4 x 64 output columns: column block = output parity class (py, px)
        const int cls = n0 >> 6, py = cls >> 1, px = cls & 1;
        opix = (size_t)f * 4 * H * W + (size_t)(rv ? (2 * oy + py) * 2 * W + 2 * ox + px : 0);
        ocol = n0 & 63;
      }
      if (bias_smem) {
        const float4* bp = reinterpret_cast<const float4*>(s_bias[wg]);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float4 b = bp[i];
          acc[4 * i] += b.x; acc[4 * i + 1] += b.y; acc[4 * i + 2] += b.z; acc[4 * i + 3] += b.w;
        }
      } else if (p.bias) {
        const float4* bp = reinterpret_cast<const float4*>(p.bias + n0);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float4 b = __ldg(bp + i);
          acc[4 * i] += b.x; acc[4 * i + 1] += b.y; acc[4 * i + 2] += b.z; acc[4 * i + 3] += b.w;
        }
      }
      if (tre) { te1 = clock64(); te_bias += te1 - te0; }
      store_rows_coalesced(wbuf, acc, p.Out, opix, p.ldo, ocol, rv, lane);
      if (tre) te_store += clock64() - te1;
      if (defer_stats) {
        if (rv) {
#pragma unroll
          for (int b8 = 0; b8 < 8; ++b8) {
            float s = 0.f, ss = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) { const float x = acc[b8 * 8 + i]; s += x; ss += x * x; }
            gs[b8] += s; gss[b8] += ss;
          }
        }
        if (++pending == 8) flush_stats();
      } else if (p.stats != nullptr) {
        if (etid < 16) s_st[etid] = 0.f;
        asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
#pragma unroll
        for (int b8 = 0; b8 < 8; ++b8) {
          float s = 0.f, ss = 0.f;
          if (rv) {
#pragma unroll
            for (int i = 0; i < 8; ++i) { const float x = acc[b8 * 8 + i]; s += x; ss += x * x; }
          }
          s = warp_sum(s); ss = warp_sum(ss);
          if (lane == 0) {
            const int grp = (n0 + b8 * 8) / p.cpg;
            atomicAdd(&s_st[2 * grp], s);
            atomicAdd(&s_st[2 * grp + 1], ss);
          }
        }
        asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
        if (etid < 16) {
          const int grp = etid >> 1;
          const int glo = n0 / p.cpg, ghi = (n0 + 63) / p.cpg;
          if (grp >= glo && grp <= ghi) atomicAdd(&p.stats[etid], (double)s_st[etid]);
        }
      }
      if (tre) te_final += clock64() - te0;
    }
    if (defer_stats && pending > 0) flush_stats();
    if (tre) {
      p.trace[9] = (unsigned long long)te_wait; p.trace[12] = (unsigned long long)te_drain; p.trace[10] = (unsigned long long)te_final;
      p.trace[13] = (unsigned long long)te_bias; p.trace[14] = (unsigned long long)te_store;
    }
  }

  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == MMA_WARP) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS));
  }
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link dependency on libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int halo_tensor_map(const void* plane, int Cin, int W, int H, int F, CUtensorMap* out) {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    DAWN_CUDA_OK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q));
    if (q != cudaDriverEntryPointSuccess || !ptr) { set_last_error("cuTensorMapEncodeTiled is not available in this driver"); return -2; }
    fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  // one map per (plane, geometry): encoding costs microseconds but the activations of a handle live at fixed addresses
  static std::map<std::tuple<const void*, int, int, int, int>, CUtensorMap> cache;
  const auto key = std::make_tuple(plane, Cin, W, H, F);
  auto it = cache.find(key);
  if (it != cache.end()) { *out = it->second; return 0; }
  const cuuint64_t dims[4] = {(cuuint64_t)Cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)F};
  const cuuint64_t strides[3] = {(cuuint64_t)Cin * 2, (cuuint64_t)W * Cin * 2, (cuuint64_t)H * W * Cin * 2};
  const cuuint32_t box[4] = {64, (cuuint32_t)HW, (cuuint32_t)HH, 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  CUtensorMap m;
  const CUresult r = fn(&m, CU_TENSOR_MAP_DATA_TYPE_UINT16, 4, const_cast<void*>(plane), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_last_error("cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r)); return -2; }
  if (cache.size() > 4096) cache.clear();
  cache[key] = m;
  *out = m;
  return 0;
}

template <int BN>
int launch_c3(const GemmParams& p, const float* Bimg, cudaStream_t st) {
  using C = CCfg<BN>;
  static bool attr_set = false;
  static int num_sms = 0;
  if (!attr_set) {
    DAWN_CUDA_OK(cudaFuncSetAttribute(tc_conv3_kernel<BN, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_DYN));
    DAWN_CUDA_OK(cudaFuncSetAttribute(tc_conv3_kernel<BN, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_DYN));
    DAWN_CUDA_OK(cudaFuncSetAttribute(tc_conv3_kernel<BN, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_DYN));
    int dev = 0;
    DAWN_CUDA_OK(cudaGetDevice(&dev));
    DAWN_CUDA_OK(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
    attr_set = true;
  }
  const int tiles_y = p.IH / TH, tiles_x = p.IW / TW, tiles_n = p.N / BN;
  const int F = p.M / (p.IH * p.IW);
  const int grid = std::min(F * tiles_y * tiles_x * tiles_n, num_sms);
  if (p.A16h != nullptr && p.A16l != nullptr) {
    CUtensorMap mh, ml;
    if (halo_tensor_map(p.A16h, p.Cin, p.IW, p.IH, F, &mh) != 0 || halo_tensor_map(p.A16l, p.Cin, p.IW, p.IH, F, &ml) != 0) return -2;
    tc_conv3_kernel<BN, true, false><<<grid, C::NTHREADS, C::SMEM_DYN, st>>>(p, Bimg, tiles_y, tiles_x, tiles_n, mh, ml);
  } else {
    CUtensorMap dummy;
    memset(&dummy, 0, sizeof(dummy));
    if (p.trace) tc_conv3_kernel<BN, false, true><<<grid, C::NTHREADS, C::SMEM_DYN, st>>>(p, Bimg, tiles_y, tiles_x, tiles_n, dummy, dummy);
    else tc_conv3_kernel<BN, false, false><<<grid, C::NTHREADS, C::SMEM_DYN, st>>>(p, Bimg, tiles_y, tiles_x, tiles_n, dummy, dummy);
  }
  DAWN_LAUNCH_OK();
  return 0;
}

}  // namespace

// 3x3, stride 1, same padding, static weights, spatial size a multiple of the 16x8 tile, 64-channel chunks, EPI_PLAIN without residual
bool tc_conv3_supported(const GemmParams& p, int epi) {
  if (epi != EPI_PLAIN || p.Res != nullptr || p.perm_in || p.perm_out) return false;
  if (p.ntaps != 9 || p.in_stride != 1 || p.out_stride != 1 || p.oy0 != 0 || p.ox0 != 0) return false;
  if (p.IH != p.OH || p.IW != p.OW || p.OHs != p.OH || p.OWs != p.OW) return false;
  if (p.IH % TH != 0 || p.IW % TW != 0) return false;
  if (p.Cin % 64 != 0 || p.N % 64 != 0 || p.K != 9 * p.Cin) return false;
  if (p.b_batch_stride != 0 || p.rows_per_batch != p.M) return false;
  if ((p.lda & 3) || (p.ldo & 3)) return false;
  if (p.stats && (p.cpg % 8 != 0)) return false;
  if (p.up2 && (p.stats != nullptr || p.N != 256)) return false;
  return true;
}

int launch_tc_conv3(const GemmParams& p, const float* Bimg, cudaStream_t st) {
  if (!tc_conv3_supported(p, EPI_PLAIN)) { set_last_error("launch_tc_conv3: unsupported geometry"); return -1; }
  if (tc_tile_n(p.N) == 128) return launch_c3<128>(p, Bimg, st);
  return launch_c3<64>(p, Bimg, st);
}

}  // namespace dawn
