// Temporal attention of the 64-channel levels on tcgen05 (reference U:648-725 == LA:71-99, 275-342 inside Residual(PreNorm(...)),
// U:763-765).  One work unit = one pixel's frame window (<= 240 frames) and a query range inside it; a persistent CTA per SM walks
// the units.  Everything between the layer input and the layer output stays on chip:
//
//   x[w0 .. w0+wn, p, :] -> LayerNorm (the affine weight is folded into W'), fp16 hi|lo split -> X (shared, K-major, 128-B swizzle)
//   per head h (8):   PROJ   [q|k|v]_h = X . W'_h^T            UMMA 128 x 96 x 64 per row tile          (SS, TMEM fp32)
//                     E1     rotary(q, k), split -> Q_h, K_h (rows) and V_h^T (dims x keys) in shared memory
//                     S      S = Q_h . K_h^T                    UMMA 128 x 160 x 32 per row tile         (SS)
//                     E2     + relative bias / band mask (table), row softmax with warp-private rows, P = exp2(S - max) written
//                            back IN PLACE into the S columns as packed fp16 hi | lo
//                     PV     O = P . V_h                        UMMA 128 x 32 x 160, A operand from TMEM (TS)
//                     E3     O / rowsum, split -> O_h (shared)
//                     Y      Y += O_h . Wout_h^T                UMMA 128 x 64 x 32, accumulated over the 8 heads in TMEM
//   out = residual + Y
//
// Every product is the 3-term fp16 split (lo*hi + hi*lo + hi*hi, fp32 accumulate): SURVEY App. D.  The window's rows are cut into two
// balanced row tiles (each <= 112 rows on TMEM lanes 0..127); compute warpgroup j (4 warps = 128 lanes) owns tile j for the whole unit, so
// its softmax runs while the other tile's MMAs execute.  Roles: warps 0-7 compute (two warpgroups), warp 8 lane 0 issues every
// tcgen05.mma, warp 9 lane 0 streams the per-head weight images (cp.async.bulk).  TMEM (512 columns): S/P tiles 2 x 160 (the QKV
// accumulator aliases their first 96 columns), O 2 x 32, Y 2 x 64.  The sequence length is unbounded: long sequences are cut into
// segments whose windows overlap by the band (K/V of the overlap are re-projected).
#include <cuda_fp16.h>
#include <cmath>
#include <cstring>
#include <vector>
#include <algorithm>
#include "common.cuh"
#include "tc_common.cuh"
#include "temporal_tc.cuh"

namespace dawn {
namespace {

using namespace tc;

constexpr int WMAX = kTtcWindowMax;
constexpr int SN = 160;                    // key columns of an S tile
constexpr int CW = 128;                    // columns one softmax thread looks at (32 rows + 2*band <= 112, + 16 alignment)
constexpr int NTH = 384;                   // 8 compute warps + one auxiliary warpgroup (issuers in warps 8, 9)
constexpr int MMA_WARP = 8;             // warps 8, 9: MMA issuers of tile 0 / tile 1
constexpr float LOG2E = 1.4426950408889634f;

// shared-memory map (bytes from a 1024-aligned base); every UMMA operand starts on a 1024-byte swizzle atom
constexpr int XH_OFF = 0;                          // X hi: 240 rows x 128 B
constexpr int XL_OFF = XH_OFF + WMAX * 128;
constexpr int WQ_OFF = XL_OFF + WMAX * 128;        // W'_h hi (96 x 128 B) | lo
constexpr int WO_OFF = WQ_OFF + 2 * 96 * 128;      // Wout_h (64 x 128 B)
constexpr int K_OFF = WO_OFF + 64 * 128;           // K_h rows: [hi 32 | lo 32] halfs
constexpr int Q_OFF = K_OFF + WMAX * 128;
constexpr int O_OFF = Q_OFF + WMAX * 128;
constexpr int VH_OFF = O_OFF + WMAX * 128;         // V_h^T hi: 4 chunks of 64 keys, each 32 dims x 128 B
constexpr int VL_OFF = VH_OFF + 4 * 4096;
constexpr int TBL_OFF = VL_OFF + 4 * 4096;         // per-warpgroup bias/mask table of the current head
constexpr int SMEM_END = TBL_OFF + 2 * kTtcTable * 4;
constexpr int SMEM_DYN = SMEM_END + 1024;
constexpr int WQ_BYTES = 2 * 96 * 128, WO_BYTES = 64 * 128;

// TMEM columns
__device__ __forceinline__ constexpr uint32_t s_col(int j) { return 160u * j; }
__device__ __forceinline__ constexpr uint32_t o_col(int j) { return 320u + 32u * j; }
__device__ __forceinline__ constexpr uint32_t y_col(int j) { return 384u + 64u * j; }

constexpr uint32_t idesc_n(int n) { return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24); }   // D f32, A/B f16, K-major, M = 128

// (the P*V product takes its A operand from TMEM: fp16 pairs packed per 32-bit column -- mma_ts below)
// Every operand of this kernel is a K-major, 128-byte-swizzled panel: the descriptors differ only in their low word (start address
// >> 4 | LBO << 16); the high word (SBO = 1024 B, descriptor version 1, SWIZZLE_128B) is one constant.  Keeping 32-bit low words
// instead of 64-bit descriptors halves the issuer's (uniform-)register pressure.
constexpr uint32_t kDescHi = 64u | (1u << 14) | (2u << 29);
__device__ __forceinline__ uint32_t desc_lo(uint32_t saddr) { return ((saddr >> 4) & 0x3FFFu) | (1u << 16); }
// The issuer warps run their control flow WARP-UNIFORMLY -- all 32 lanes wait on the barriers and compute the operands -- and elect one
// lane per instruction (always the same lane: it also executes the commits that track its MMAs; tc_common.cuh: elect_one).  With provably uniform operands the
// descriptors live in uniform registers and an MMA costs one UTCHMMA; issued from inside an `if (lane == 0)` region every operand went
// through a per-lane R2UR loop (~75 cycles of issue per MMA, against 16 cycles of tensor-pipe time for a 128 x 32 x 16 MMA).
__device__ __forceinline__ void mma_ss(uint32_t tmem_d, uint32_t alo, uint32_t blo, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p, q;\n\t"
      ".reg .b64 da, db;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "mov.b64 da, {%1, %5};\n\t"
      "mov.b64 db, {%2, %5};\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n\t"
      "}" ::"r"(tmem_d), "r"(alo), "r"(blo), "r"(idesc), "r"(accumulate), "r"(kDescHi)
      : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t tmem_d, uint32_t tmem_a, uint32_t blo, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p, q;\n\t"
      ".reg .b64 db;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "mov.b64 db, {%2, %5};\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %3, p;\n\t"
      "}" ::"r"(tmem_d), "r"(tmem_a), "r"(blo), "r"(idesc), "r"(accumulate), "r"(kDescHi)
      : "memory");
}
__device__ __forceinline__ void tmem_ld32_async(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
      "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st8_zero(uint32_t taddr) {
  const uint32_t z = 0u;
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(taddr), "r"(z) : "memory");
}
__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void named_bar(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
// (x0, x1) -> packed fp16 hi pair and lo pair, hi = rn16(x), lo = rn16(x - hi): two packed conversions, two widenings, two subtractions
// (tc_common's split_f16x2 rounds in integer arithmetic first: two more instructions per pair; every value split in this kernel is far
// inside the fp16 range, so the direct conversion cannot overflow)
__device__ __forceinline__ void split2(float x0, float x1, uint32_t& hi, uint32_t& lo) {
  const __half2 h = __floats2half2_rn(x0, x1);
  const float2 hf = __half22float2(h);
  const __half2 l = __floats2half2_rn(x0 - hf.x, x1 - hf.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}
// 8 consecutive values of a row -> one 16-byte chunk of the hi half and one of the lo half of a [hi 32 | lo 32] operand row
__device__ __forceinline__ void store_row_chunk(uint8_t* base, int row, int c8, const float (&v)[8]) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) split2(v[2 * i], v[2 * i + 1], h[i], l[i]);
  *reinterpret_cast<uint4*>(base + swz(row, c8)) = make_uint4(h[0], h[1], h[2], h[3]);
  *reinterpret_cast<uint4*>(base + swz(row, 4 + c8)) = make_uint4(l[0], l[1], l[2], l[3]);
}

struct Bars {
  uint64_t x_ready, wq_ready, wq_free, wo_ready, wo_free, kv_ready, v_ready;
  uint64_t proj_ready[2], s_ready[2], p_ready[2], o_ready[2], oh_ready[2], y_ready[2];
};

template <bool TRACE>
__global__ void __launch_bounds__(NTH, 1) temporal_tc_kernel(const TemporalTcArgs a) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  __shared__ Bars bars;
  __shared__ uint32_t s_tmem_base;
  // round up inside the shared window (pointer arithmetic on the __shared__ array keeps the address space: LDS/STS, not generic LD/ST)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);          // warp-uniform by construction (the issuer warps rely on it)

  if (tid == 0) {
    mbar_init(&bars.x_ready, 256); mbar_init(&bars.kv_ready, 256); mbar_init(&bars.v_ready, 256);
    mbar_init(&bars.wq_ready, 1); mbar_init(&bars.wq_free, 2); mbar_init(&bars.wo_ready, 1); mbar_init(&bars.wo_free, 2);
    for (int j = 0; j < 2; ++j) {
      mbar_init(&bars.proj_ready[j], 1); mbar_init(&bars.s_ready[j], 1); mbar_init(&bars.o_ready[j], 1); mbar_init(&bars.y_ready[j], 1);
      mbar_init(&bars.p_ready[j], 128); mbar_init(&bars.oh_ready[j], 128);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == MMA_WARP) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem_base)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  // operand buffers start as zeros: rows beyond a window are multiplied (into masked or unused results) and must be finite
  {
    uint4* p0 = reinterpret_cast<uint4*>(smem + XH_OFF);
    for (int i = tid; i < (2 * WMAX * 128) / 16; i += NTH) p0[i] = make_uint4(0, 0, 0, 0);
    uint4* p1 = reinterpret_cast<uint4*>(smem + K_OFF);
    for (int i = tid; i < (TBL_OFF - K_OFF) / 16; i += NTH) p1[i] = make_uint4(0, 0, 0, 0);
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = s_tmem_base;

  const int nunits = a.P * a.nseg;
  const int band = a.band;

  // Registers are per scheduler partition (16K each, 3 warps per partition here: launch at 168 per thread).  The auxiliary
  // warpgroup hands most of its share to the two compute warps of its partition: 2 x 208 + 88 = 504 = the 3 x 168 this CTA owns per partition (asking for more than the CTA's pool blocks forever).
  if (warp < 8) {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 208;" ::: "memory");
    // ======================================================================= compute warpgroups
    const int j = warp >> 2;                               // tile owned by this warpgroup
    const int wq = warp & 3;                               // TMEM lane quarter
    const int l = wq * 32 + lane;                          // lane = row inside the tile
    const int wgtid = tid & 127;
    const uint32_t lane_addr = (uint32_t)(wq * 32) << 16;
    float* tb = reinterpret_cast<float*>(smem + TBL_OFF) + j * kTtcTable;
    const float fa = a.inv_wscale, faq = a.inv_wscale * LOG2E;      // scores live in the log2 domain
    // completed phases of the per-tile barriers (rows exist / queries exist), for this warpgroup's tile and for the other one
    uint32_t nact_own = 0, nact_oth = 0, nq_own = 0, nq_oth = 0;
    uint64_t* const proj_own = &bars.proj_ready[j];
    uint64_t* const proj_oth = &bars.proj_ready[1 - j];
    uint64_t* const o_oth = &bars.o_ready[1 - j];
    const bool tr = TRACE && (a.trace != nullptr) && blockIdx.x == 0 && wgtid == 0;   // cycle trace of CTA 0 (one thread per warpgroup)
    long long tc_[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, t0 = 0;
#define TTC_T(i) do { if (TRACE && tr) { const long long t1_ = clock64(); tc_[i] += t1_ - t0; t0 = t1_; } } while (0)

    float4 xpre[7];                                        // this thread's share of the next window's first 112 x rows, requested one unit ahead
    bool have_pre = false;
    for (int u = blockIdx.x; u < nunits; u += gridDim.x) {
      const int pix = u / a.nseg;
      const TtcSegment sg = a.seg[u - pix * a.nseg];
      TtcTile tl[2];
      ttc_tiles(sg, band, tl);
      const TtcTile T = j ? tl[1] : tl[0], To = j ? tl[0] : tl[1];
      const bool act_own = T.r1 > T.r0, act_oth = To.r1 > To.r0, hq_own = T.q1 > T.q0, hq_oth = To.q1 > To.q0;
      const int row = T.r0 + l;                            // window row of this thread
      const bool row_ok = row < T.r1;
      const bool dbg = (a.dbg != nullptr) && u == 0;

      if (tr) t0 = clock64();
      // ------------------------------------------------------------------ prologue: x rows -> LayerNorm -> fp16 split
      // 16 lanes x float4 = one 64-channel row, 16 rows per pass.  The rows were requested during the previous unit's last head
      // (xpre holds them); only the first unit of a CTA loads here.
      {
        const int l16 = tid & 15, rg = tid >> 4;
        const float* xb = a.x + ((size_t)sg.w0 * a.P + pix) * a.ldx + 4 * l16;
        const size_t fstride = (size_t)a.P * a.ldx;
        float4 xsec[8];                                    // rows 112 + : loaded now, consumed after the prefetched half
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int r = (7 + i) * 16 + rg;
          xsec[i] = (r < sg.wn) ? __ldg(reinterpret_cast<const float4*>(xb + (size_t)r * fstride)) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        if (!have_pre) {
#pragma unroll
          for (int i = 0; i < 7; ++i) {
            const int r = i * 16 + rg;
            xpre[i] = (r < sg.wn) ? __ldg(reinterpret_cast<const float4*>(xb + (size_t)r * fstride)) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
#pragma unroll
        for (int i = 0; i < 15; ++i) {
          const int r = i * 16 + rg;
          const float4 v = i < 7 ? xpre[i < 7 ? i : 0] : xsec[i < 7 ? 0 : i - 7];
          float s = (v.x + v.y) + (v.z + v.w);
#pragma unroll
          for (int o = 1; o < 16; o <<= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
          const float mu = s * (1.0f / 64.f);
          const float d0 = v.x - mu, d1 = v.y - mu, d2 = v.z - mu, d3 = v.w - mu;
          float ss = (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3);
#pragma unroll
          for (int o = 1; o < 16; o <<= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
          if (r < sg.wn) {
            // the normalised row goes into the projection (its affine weight is folded into W'): the accumulator needs no correction
            const float rstd = 1.0f / sqrtf(ss * (1.0f / 64.f) + 1e-5f);
            uint32_t h0, l0, h1, l1;
            split2(d0 * rstd, d1 * rstd, h0, l0); split2(d2 * rstd, d3 * rstd, h1, l1);
            const uint32_t off = swz(r, l16 >> 1) + (l16 & 1) * 8;
            *reinterpret_cast<uint2*>(smem + XH_OFF + off) = make_uint2(h0, h1);
            *reinterpret_cast<uint2*>(smem + XL_OFF + off) = make_uint2(l0, l1);
          }
        }
        fence_proxy_async();
        mbar_arrive(&bars.x_ready);
      }
      TTC_T(0);
      const float4* rotp = reinterpret_cast<const float4*>(a.rot + (size_t)(sg.w0 + (row_ok ? row : T.r0)) * 32);
      float inv_l = 1.f;

      for (int h = 0; h < 8; ++h) {
        // prefetch what this head needs from global memory before blocking on the projection: the (cos, sin) row of this thread's
        // frame (the shared-memory carve-out leaves ~9 KB of L1: these come from L2) and this thread's slice of the bias/mask table
        float4 cs4[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) cs4[i] = __ldg(rotp + i);
        float4 tab4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (hq_own) tab4 = __ldg(reinterpret_cast<const float4*>(a.table + h * kTtcTable) + wgtid);

        // -------------------------------------------------------------- E1: projection accumulator -> Q_h, K_h, V_h^T
        if (act_own) {
          mbar_wait(proj_own, nact_own & 1);
          TTC_T(1);
          tc_fence_after();
          uint32_t rq[32], rk[32], rv[32];
          const uint32_t ta = tmem_base + lane_addr + s_col(j);
          tmem_ld32_async(ta, rq); tmem_ld32_async(ta + 32, rk); tmem_ld32_async(ta + 64, rv);
          tmem_wait_ld();
          // K/V/Q of the previous head may still be read by the other tile's S / PV
          if (nq_oth > 0) mbar_wait(o_oth, (nq_oth - 1) & 1);
          TTC_T(2);
          if (dbg && h == 0 && row_ok) {
            float* d = a.dbg + (size_t)row * 96;
#pragma unroll
            for (int i = 0; i < 32; ++i) { d[i] = __uint_as_float(rq[i]); d[32 + i] = __uint_as_float(rk[i]); d[64 + i] = __uint_as_float(rv[i]); }
          }
          if (row_ok) {
#pragma unroll
            for (int c8 = 0; c8 < 4; ++c8) {
              float q8[8], k8[8];
              const float4 cs0 = cs4[2 * c8], cs1 = cs4[2 * c8 + 1];      // (cos, sin) of pairs 4 c8 .. 4 c8 + 3
              const float cs[8] = {cs0.x, cs0.y, cs0.z, cs0.w, cs1.x, cs1.y, cs1.z, cs1.w};
#pragma unroll
              for (int pp = 0; pp < 4; ++pp) {
                const int i = c8 * 8 + 2 * pp;
                const float q0 = faq * __uint_as_float(rq[i]), q1 = faq * __uint_as_float(rq[i + 1]);
                const float k0 = fa * __uint_as_float(rk[i]), k1 = fa * __uint_as_float(rk[i + 1]);
                const float co = cs[2 * pp], si = cs[2 * pp + 1];
                q8[2 * pp] = q0 * co - q1 * si; q8[2 * pp + 1] = q1 * co + q0 * si;
                k8[2 * pp] = k0 * co - k1 * si; k8[2 * pp + 1] = k1 * co + k0 * si;
              }
              store_row_chunk(smem + Q_OFF, row, c8, q8);
              store_row_chunk(smem + K_OFF, row, c8, k8);
            }
          }
          // K_h and Q_h are complete: S = Q K^T starts while V_h^T is still being written
          fence_proxy_async();
          mbar_arrive(&bars.kv_ready);
          if (row_ok) {
            // V_h^T: element (dim d, key = row) of chunk row >> 6
            uint8_t* vh = smem + VH_OFF + (row >> 6) * 4096 + (row & 7) * 2;
            uint8_t* vl = smem + VL_OFF + (row >> 6) * 4096 + (row & 7) * 2;
            const int cc = (row & 63) >> 3;
#pragma unroll
            for (int d = 0; d < 32; d += 2) {
              uint32_t hi, lo;
              split2(fa * __uint_as_float(rv[d]), fa * __uint_as_float(rv[d + 1]), hi, lo);
              const uint32_t o0 = swz(d, cc), o1 = swz(d + 1, cc);
              *reinterpret_cast<uint16_t*>(vh + o0) = (uint16_t)hi; *reinterpret_cast<uint16_t*>(vh + o1) = (uint16_t)(hi >> 16);
              *reinterpret_cast<uint16_t*>(vl + o0) = (uint16_t)lo; *reinterpret_cast<uint16_t*>(vl + o1) = (uint16_t)(lo >> 16);
            }
          }
          tc_fence_before();
          fence_proxy_async();
          ++nact_own;
        } else {
          if (nq_oth > 0) mbar_wait(o_oth, (nq_oth - 1) & 1);
          mbar_arrive(&bars.kv_ready);
        }
        mbar_arrive(&bars.v_ready);
        TTC_T(3);

        if (hq_own) {
          // table of this head (the previous head's readers are done: they arrived on p_ready before anyone could pass o_ready)
          *reinterpret_cast<float4*>(tb + 4 * wgtid) = tab4;
          named_bar(1 + j, 128);
          // ------------------------------------------------------------ E2: softmax of this thread's row
          mbar_wait(&bars.s_ready[j], nq_own & 1);
          TTC_T(4);
          tc_fence_after();
          int xw = T.r0 + 32 * wq - band - T.kb;
          if (xw < 0) xw = 0;
          int cstart = xw & ~15;
          if (cstart > SN - CW) cstart = SN - CW;
          const int rowc = row < WMAX - 1 ? row : WMAX - 1;
          const float* tp = tb + (T.kb + cstart - rowc + kTtcTableZero);
          const int climit = (sg.wn < T.kb + SN ? sg.wn : T.kb + SN) - T.kb - cstart;      // columns whose key lies inside the window
          const uint32_t ts = tmem_base + lane_addr + s_col(j);
          uint32_t sv[128];
          {
            uint32_t (*sv4)[32] = reinterpret_cast<uint32_t (*)[32]>(sv);
            tmem_ld32_async(ts + cstart, sv4[0]); tmem_ld32_async(ts + cstart + 32, sv4[1]);
            tmem_ld32_async(ts + cstart + 64, sv4[2]); tmem_ld32_async(ts + cstart + 96, sv4[3]);
            tmem_wait_ld();
          }
          if (dbg && h == 0 && row_ok) {
            float* d = a.dbg + (size_t)WMAX * 96 + (size_t)row * 130;
            d[0] = (float)(T.kb + cstart); d[1] = (float)climit;
#pragma unroll
            for (int c = 0; c < 128; ++c) d[2 + c] = __uint_as_float(sv[c]);
          }
          float m = -1e30f;
          if (climit >= CW) {
#pragma unroll
            for (int c = 0; c < CW; ++c) {
              const float s = __uint_as_float(sv[c]) + tp[c];
              sv[c] = __float_as_uint(s);
              m = fmaxf(m, s);
            }
          } else {
#pragma unroll
            for (int c = 0; c < CW; ++c) {
              const float s = (c < climit) ? __uint_as_float(sv[c]) + tp[c] : -1e30f;
              sv[c] = __float_as_uint(s);
              m = fmaxf(m, s);
            }
          }
          // P (fp16 pairs) over the S columns: hi at [0, 80), lo at [80, 160); zeros outside this warp's 128-key span.  Written 32 keys
          // at a time so that the stores overlap the remaining exponentials (every S column of this row is already in registers)
          const uint32_t c2 = (uint32_t)(cstart >> 1);
          float lsum = 0.f;
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) {
            uint32_t ph[16], pl[16];
#pragma unroll
            for (int c = 0; c < 32; c += 2) {
              const float p0 = ex2f(__uint_as_float(sv[32 * q4 + c]) - m), p1 = ex2f(__uint_as_float(sv[32 * q4 + c + 1]) - m);
              lsum += p0 + p1;
              split2(p0, p1, ph[c >> 1], pl[c >> 1]);
            }
            tmem_st16(ts + c2 + 16 * q4, ph);
            tmem_st16(ts + 80 + c2 + 16 * q4, pl);
          }
          inv_l = 1.0f / lsum;
#pragma unroll
          for (int b = 0; b < 10; ++b) {
            if (8 * b < (int)c2 || 8 * b >= (int)c2 + 64) { tmem_st8_zero(ts + 8 * b); tmem_st8_zero(ts + 80 + 8 * b); }
          }
          tmem_wait_st();
          tc_fence_before();
          mbar_arrive(&bars.p_ready[j]);
          TTC_T(5);
        }
        if (h == 7) {
          // request the next unit's x rows now: their DRAM latency hides behind this head's P*V, E3, Y and the epilogue
          const int un = u + (int)gridDim.x;
          have_pre = un < nunits;
          if (have_pre) {
            const int pixn = un / a.nseg;
            const TtcSegment sn = a.seg[un - pixn * a.nseg];
            const int l16 = tid & 15, rg = tid >> 4;
            const float* xb = a.x + ((size_t)sn.w0 * a.P + pixn) * a.ldx + 4 * l16;
            const size_t fstride = (size_t)a.P * a.ldx;
#pragma unroll
            for (int i = 0; i < 7; ++i) {
              const int r = i * 16 + rg;
              xpre[i] = (r < sn.wn) ? __ldg(reinterpret_cast<const float4*>(xb + (size_t)r * fstride)) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
          }
        }
        if (hq_own) {
          // ------------------------------------------------------------ E3: O / rowsum -> O_h
          mbar_wait(&bars.o_ready[j], nq_own & 1);
          TTC_T(6);
          tc_fence_after();
          {
            uint32_t ro[32];
            tmem_ld32_async(tmem_base + lane_addr + o_col(j), ro);
            tmem_wait_ld();
            if (dbg && h == 0 && row_ok) {
              float* d = a.dbg + (size_t)WMAX * (96 + 130) + (size_t)row * 33;
              d[0] = 1.0f / inv_l;
#pragma unroll
              for (int i = 0; i < 32; ++i) d[1 + i] = __uint_as_float(ro[i]);
            }
            if (row_ok) {
#pragma unroll
              for (int c8 = 0; c8 < 4; ++c8) {
                float o8[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) o8[i] = __uint_as_float(ro[c8 * 8 + i]) * inv_l;
                store_row_chunk(smem + O_OFF, row, c8, o8);
              }
            }
          }
          tc_fence_before();
          fence_proxy_async();
          mbar_arrive(&bars.oh_ready[j]);
          TTC_T(7);
          ++nq_own;
        }
        if (hq_oth) ++nq_oth;
        if (act_oth) ++nact_oth;
      }
      // the next prologue overwrites X: the other tile's last projection must have finished reading it
      if (act_oth) mbar_wait(proj_oth, (nact_oth - 1) & 1);

      // ------------------------------------------------------------------ unit epilogue: out = residual + Y
      if (hq_own) {
        const bool is_q = row >= T.q0 && row < T.q1;
        const size_t orow = (size_t)(sg.w0 + (is_q ? row : T.q0) - a.q_lo) * a.P + pix;
        const float4* rp = reinterpret_cast<const float4*>(a.res + orow * a.ldr);
        float4 rr[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) rr[i] = rp[i];                  // first half of the residual row in flight while the last Y finishes
        mbar_wait(&bars.y_ready[j], (nq_own / 8 - 1) & 1);       // y_ready completes once per unit with queries
        TTC_T(8);
        tc_fence_after();
        float4* op = reinterpret_cast<float4*>(a.out + orow * a.ldo);
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          uint32_t ry[32];
          tmem_ld32_async(tmem_base + lane_addr + y_col(j) + 32 * half, ry);
          tmem_wait_ld();
          float4 nx[8];
          if (half == 0) {
#pragma unroll
            for (int i = 0; i < 8; ++i) nx[i] = rp[8 + i];
          }
          if (is_q) {
#pragma unroll
            for (int i = 0; i < 8; ++i)
              op[8 * half + i] = make_float4(fmaf(__uint_as_float(ry[4 * i]), a.inv_oscale, rr[i].x), fmaf(__uint_as_float(ry[4 * i + 1]), a.inv_oscale, rr[i].y),
                                             fmaf(__uint_as_float(ry[4 * i + 2]), a.inv_oscale, rr[i].z), fmaf(__uint_as_float(ry[4 * i + 3]), a.inv_oscale, rr[i].w));
          }
          if (half == 0) {
#pragma unroll
            for (int i = 0; i < 8; ++i) rr[i] = nx[i];
          }
        }
        tc_fence_before();
        TTC_T(9);
      }
    }
    if (tr) for (int i = 0; i < 12; ++i) a.trace[16 * j + i] = (unsigned long long)tc_[i];
  } else {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 88;" ::: "memory");
  }
  if (warp == MMA_WARP || warp == MMA_WARP + 1) {
    // ======================================================================= MMA issuers: warp 8 -> tile 0 (+ weight images), warp 9 -> tile 1
    // (whole warps, warp-uniform control flow; one elected lane per MMA / commit / copy: see mma_ss)
    {
      const int j = warp - MMA_WARP;
      const uint32_t sb = smem_u32(smem);
      const uint32_t tmem_base_u = __shfl_sync(0xffffffffu, tmem_base, 0);      // read from shared memory: make it a provably uniform value
      uint32_t it = 0, ui = 0, nqj = 0;
      const bool tr = TRACE && (a.trace != nullptr) && blockIdx.x == 0;
      long long tc_[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, t0 = 0;
      const long long t_begin = clock64();
      constexpr uint32_t ID96 = idesc_n(96), ID160 = idesc_n(SN), ID32 = idesc_n(32), ID64 = idesc_n(64);
      const uint32_t my_units = (nunits - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
      const uint32_t total_it = my_units * 8;
      // weight images (tile 0's issuer): a buffer is refilled as soon as BOTH issuers have committed its last readers
      auto load_wq = [&](uint32_t itn) {
        if (j != 0 || itn >= total_it) return;
        mbar_wait(&bars.wq_free, (itn & 1) ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(&bars.wq_ready, WQ_BYTES);
          bulk_copy_g2s(smem + WQ_OFF, a.Wqkv + (size_t)(itn & 7) * WQ_BYTES, WQ_BYTES, &bars.wq_ready);
        }
        __syncwarp();
      };
      auto load_wo = [&](uint32_t itn) {
        if (j != 0 || itn >= total_it) return;
        mbar_wait(&bars.wo_free, (itn & 1) ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(&bars.wo_ready, WO_BYTES);
          bulk_copy_g2s(smem + WO_OFF, a.Wout + (size_t)(itn & 7) * WO_BYTES, WO_BYTES, &bars.wo_ready);
        }
        __syncwarp();
      };
      load_wq(0); load_wo(0);
      for (int u = blockIdx.x; u < nunits; u += gridDim.x, ++ui) {
        const int pix = u / a.nseg;
        const TtcSegment sg = a.seg[u - pix * a.nseg];
        TtcTile tl[2];
        ttc_tiles(sg, band, tl);
        const TtcTile T = j ? tl[1] : tl[0];
        const bool act = T.r1 > T.r0, hq = T.q1 > T.q0;
        // k-steps (16 keys) of P*V that can hold a band key of this tile's queries; P is zero beyond them
        int nks = ((sg.wn < T.q1 + band ? sg.wn : T.q1 + band) - T.kb + 15) >> 4;
        if (nks > SN / 16) nks = SN / 16;
        const uint32_t ro = (uint32_t)(T.r0 >> 3) * 1024u, ko = (uint32_t)(T.kb >> 3) * 1024u;
        const uint32_t x_hi = desc_lo(sb + XH_OFF + ro), x_lo = desc_lo(sb + XL_OFF + ro);
        const uint32_t w_hi = desc_lo(sb + WQ_OFF), w_lo = desc_lo(sb + WQ_OFF + 96 * 128);
        const uint32_t qd = desc_lo(sb + Q_OFF + ro), kd = desc_lo(sb + K_OFF + ko);
        const uint32_t od = desc_lo(sb + O_OFF + ro), wd = desc_lo(sb + WO_OFF);
        const uint32_t d_s = tmem_base_u + s_col(j), d_o = tmem_base_u + o_col(j), d_y = tmem_base_u + y_col(j);

        auto issue_proj = [&](uint32_t itn) {
          mbar_wait(&bars.wq_ready, itn & 1);
          TTC_T(1);
          tc_fence_after();
          if (act) {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              const uint32_t o = (uint32_t)(2 * ks);
              mma_ss(d_s, x_lo + o, w_hi + o, ID96, ks ? 1u : 0u);
              mma_ss(d_s, x_hi + o, w_lo + o, ID96, 1u);
              mma_ss(d_s, x_hi + o, w_hi + o, ID96, 1u);
            }
            tc_commit_elected(&bars.proj_ready[j]);
          }
          tc_commit_elected(&bars.wq_free);
          TTC_T(2);
        };

        if (tr) t0 = clock64();
        mbar_wait(&bars.x_ready, ui & 1);
        TTC_T(0);
        fence_proxy_async();
        issue_proj(it);
#pragma unroll 1
        for (int h = 0; h < 8; ++h, ++it) {
          load_wq(it + 1);                                 // both issuers have committed this head's projections by the time kv_ready can complete
          mbar_wait(&bars.kv_ready, it & 1);
          TTC_T(3);
          fence_proxy_async();
          tc_fence_after();
          if (hq) {
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
              const uint32_t hi = (uint32_t)(2 * ks), lo = (uint32_t)(4 + 2 * ks);
              mma_ss(d_s, qd + lo, kd + hi, ID160, ks ? 1u : 0u);
              mma_ss(d_s, qd + hi, kd + lo, ID160, 1u);
              mma_ss(d_s, qd + hi, kd + hi, ID160, 1u);
            }
            tc_commit_elected(&bars.s_ready[j]);
            TTC_T(4);
            mbar_wait(&bars.p_ready[j], nqj & 1);
            mbar_wait(&bars.v_ready, it & 1);
            TTC_T(5);
            fence_proxy_async();
            tc_fence_after();
            // a rolled loop on purpose: unrolled, the operand descriptors of all ten steps are precomputed and shuffled around for longer
            // than the tensor pipe needs for the MMAs
#pragma unroll 1
            for (int s = 0; s < nks; ++s) {
              // keys kb + 16 s ..: chunk (kb + 16 s) >> 6, 32-byte step inside the chunk
              const uint32_t key = (uint32_t)T.kb + 16u * s;
              const uint32_t vo = (key >> 6) * 4096u + (key & 63u) * 2u;
              const uint32_t vh = desc_lo(sb + VH_OFF + vo), vl = desc_lo(sb + VL_OFF + vo);
              mma_ts(d_o, d_s + 80 + 8 * s, vh, ID32, s ? 1u : 0u);
              mma_ts(d_o, d_s + 8 * s, vl, ID32, 1u);
              mma_ts(d_o, d_s + 8 * s, vh, ID32, 1u);
            }
            tc_commit_elected(&bars.o_ready[j]);
            TTC_T(7);
          }
          // next head's projection goes in behind P*V (its accumulator aliases the P columns); E3 / Y of this head overlap it
          if (h < 7) issue_proj(it + 1);
          mbar_wait(&bars.wo_ready, it & 1);
          TTC_T(8);
          if (hq) {
            mbar_wait(&bars.oh_ready[j], nqj & 1);
            TTC_T(9);
            fence_proxy_async();
            tc_fence_after();
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
              const uint32_t hi = (uint32_t)(2 * ks), lo = (uint32_t)(4 + 2 * ks);
              mma_ss(d_y, od + lo, wd + hi, ID64, (h || ks) ? 1u : 0u);
              mma_ss(d_y, od + hi, wd + lo, ID64, 1u);
              mma_ss(d_y, od + hi, wd + hi, ID64, 1u);
            }
            if (h == 7) tc_commit_elected(&bars.y_ready[j]);
            ++nqj;
            TTC_T(11);
          }
          tc_commit_elected(&bars.wo_free);
          load_wo(it + 1);
        }
      }
      if (tr && j == 0 && lane == 0) {
        for (int i = 0; i < 12; ++i) a.trace[32 + i] = (unsigned long long)tc_[i];
        a.trace[44] = (unsigned long long)(clock64() - t_begin); a.trace[45] = it;
      }
    }
  }
#undef TTC_T
  tc_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
}

}  // namespace

bool temporal_tc_supported(int C, int F, int band, int q_lo, int q_hi) {
  if (C != 64 || band < 1 || band > kTtcBandMax || F < 1 || q_lo < 0 || q_hi > F || q_lo >= q_hi) return false;
  TtcSegment seg[kTtcMaxSeg];
  return temporal_tc_plan(F, band, q_lo, q_hi, seg) > 0;
}

int temporal_tc_plan(int F, int band, int q_lo, int q_hi, TtcSegment* seg) {
  if (F <= kTtcWindowMax) {
    seg[0] = TtcSegment{0, F, q_lo, q_hi};
    return 1;
  }
  const int qmax = kTtcWindowMax - 2 * band;                  // queries per segment when the window needs a halo on both sides
  const int nq = q_hi - q_lo;
  const int nseg = (nq + qmax - 1) / qmax;
  if (nseg > kTtcMaxSeg) return 0;
  for (int s = 0; s < nseg; ++s) {
    const int qa = q_lo + (int)((long long)nq * s / nseg), qb = q_lo + (int)((long long)nq * (s + 1) / nseg);
    const int w0 = std::max(0, qa - band), w1 = std::min(F, qb + band);
    if (w1 - w0 > kTtcWindowMax) return 0;
    seg[s] = TtcSegment{w0, w1 - w0, qa, qb};
  }
  return nseg;
}

int launch_temporal_tc(const TemporalTcArgs& a_in, cudaStream_t st) {
  TemporalTcArgs a = a_in;
  if (a.band < 1 || a.band > kTtcBandMax) { set_last_error("temporal_tc: unsupported band"); return -1; }
  a.nseg = temporal_tc_plan(a.F, a.band, a.q_lo, a.q_hi, a.seg);
  if (a.nseg <= 0) { set_last_error("temporal_tc: unsupported shape"); return -1; }
  static int num_sms = 0;
  if (num_sms == 0) {
    int dev = 0;
    DAWN_CUDA_OK(cudaGetDevice(&dev));
    DAWN_CUDA_OK(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
    DAWN_CUDA_OK(cudaFuncSetAttribute(temporal_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_DYN));
    DAWN_CUDA_OK(cudaFuncSetAttribute(temporal_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_DYN));
  }
  const int grid = std::min(a.P * a.nseg, num_sms);
  if (a.trace) temporal_tc_kernel<true><<<grid, NTH, SMEM_DYN, st>>>(a);
  else temporal_tc_kernel<false><<<grid, NTH, SMEM_DYN, st>>>(a);
  DAWN_LAUNCH_OK();
  return 0;
}

// Host packing.  wqkv: [768][64] fp32 rows = output columns (q | k | v blocks of 256, gamma and q-scale folded), wout: [64][256].
void temporal_tc_pack(const float* wqkv, const float* wout, std::vector<uint8_t>& Wq, std::vector<uint8_t>& Wo, float* inv_wscale,
                      float* inv_oscale) {
  auto pow2scale = [](const float* p, size_t n) {
    float mx = 0.f;
    for (size_t i = 0; i < n; ++i) mx = std::max(mx, std::fabs(p[i]));
    int e = 0;
    if (mx > 0.f) std::frexp(mx, &e);
    return std::ldexp(1.0f, 11 - e);
  };
  const float sq = pow2scale(wqkv, (size_t)768 * 64), so = pow2scale(wout, (size_t)64 * 256);
  *inv_wscale = 1.0f / sq; *inv_oscale = 1.0f / so;
  auto put = [](uint8_t* hi, uint8_t* lo, float v) {
    const __half h = __float2half_rn(v);
    const __half l = __float2half_rn(v - __half2float(h));
    memcpy(hi, &h, 2); memcpy(lo, &l, 2);
  };
  auto swz_h = [](int r, int c) { return (size_t)((r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4)); };
  Wq.assign((size_t)8 * WQ_BYTES, 0);
  Wo.assign((size_t)8 * WO_BYTES, 0);
  for (int h = 0; h < 8; ++h) {
    uint8_t* qh = Wq.data() + (size_t)h * WQ_BYTES; uint8_t* ql = qh + 96 * 128;
    for (int part = 0; part < 3; ++part)
      for (int r = 0; r < 32; ++r)
        for (int k = 0; k < 64; ++k) {
          const int n = part * 32 + r;
          const size_t off = swz_h(n, k >> 3) + (size_t)(k & 7) * 2;
          put(qh + off, ql + off, wqkv[(size_t)(part * 256 + h * 32 + r) * 64 + k] * sq);
        }
    uint8_t* oi = Wo.data() + (size_t)h * WO_BYTES;
    for (int c = 0; c < 64; ++c)
      for (int d = 0; d < 32; ++d) {
        const size_t ohi = swz_h(c, d >> 3) + (size_t)(d & 7) * 2, olo = swz_h(c, 4 + (d >> 3)) + (size_t)(d & 7) * 2;
        put(oi + ohi, oi + olo, wout[(size_t)c * 256 + h * 32 + d] * so);
      }
  }
}

void temporal_tc_table(const float* bias, int band, std::vector<float>& table) {
  table.assign((size_t)8 * kTtcTable, -1e30f);
  for (int h = 0; h < 8; ++h)
    for (int rel = -band; rel <= band; ++rel)
      table[(size_t)h * kTtcTable + kTtcTableZero + rel] = bias[(size_t)h * (2 * band + 1) + rel + band] * 1.4426950408889634f;
}

}  // namespace dawn
