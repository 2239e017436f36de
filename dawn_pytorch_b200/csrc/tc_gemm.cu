// tcgen05 implicit GEMM for the DAWN UNet contractions (sm_100a): Out = epilogue(A (gathered, fp32) x B (weights)).
//
//   * 128-row position tile x 64/128-column output tile, K streamed in 64-element panels (= one 128-byte swizzle row of fp16).
//   * fp32-level parity needs the 3-term split  hi*hi + hi*lo + lo*hi  (SURVEY App. D).  The pieces are FP16
//     (kind::f16, K = 16 per instruction): an fp16 piece carries the same 11-bit significand as a TF32 piece, so
//     hi+lo keeps ~22 bits like 3xTF32, but every tcgen05.mma does twice the work and reads half the bytes
//     (measured: ~70 cycles of issue cost per tcgen05.mma regardless of N made the K=8 TF32 form issue-bound).
//     fp16's narrow exponent is handled with an exact power-of-two pre-scale of each weight matrix (undone in the
//     epilogue); activations stay unscaled (|x| < 65504; tiny lo pieces go subnormal, abs error <= 2^-25).  A is split on the fly by the producer warps, B is pre-split and pre-swizzled on the
//     host into ready-to-copy shared-memory images (1-D bulk copies, no tensor maps).
//   * The tensor core adds into its accumulator with round-toward-zero; chained over a long K that is a biased
//     error (measured 1.4e-4 at K=14112).  So TMEM holds TWO accumulator buffers; every CHUNK panels the issuer
//     flips buffers (first MMA overwrites) and the epilogue warps drain the finished buffer into fp32 registers
//     with ordinary round-to-nearest adds while the next chunk is already being multiplied.
//   * persistent CTAs, warp roles: 0-7 A producers (gather + split + swizzled st.shared, global loads prefetched
//     two panels ahead in registers), 8-11 accumulate/epilogue (each thread owns one output row), 12 MMA issuer
//     (one elected thread), 13 weight loader (one elected thread).
#include <cuda_fp16.h>
#include "common.cuh"
#include "gemm.cuh"
#include "tc_common.cuh"
#include "tc_gemm.cuh"

namespace dawn {
namespace {

constexpr int BM = 128;
constexpr int BKP = 64;                 // K elements per panel row (64 fp16 = 128 bytes)
constexpr int CHUNK = 4;                // panels accumulated inside TMEM before a drain (K = 128)
constexpr int A_PANEL_MIN = BM * 128;    // 16 KB
constexpr int NPROD = 256;              // producer threads (warps 0-7)

// BN = 64: one epilogue warpgroup, 4 stages of 48 KB.  BN = 128: two epilogue warpgroups (64 columns each), 3 stages of 64 KB.
template <int BN>
struct Cfg {
  static constexpr int NWG = BN / 64;                        // epilogue warpgroups
  static constexpr int A_PANEL = A_PANEL_MIN;
  static constexpr int B_PANEL = BN * 128;
  static constexpr int STAGE_BYTES = 2 * A_PANEL + 2 * B_PANEL;
  static constexpr int STAGES = (BN == 64) ? 4 : 3;
  static constexpr int NTHREADS = NPROD + 128 * NWG + 64;    // producers | epilogue | MMA warp | loader warp
  static constexpr int MMA_WARP = (NPROD + 128 * NWG) / 32;
  static constexpr int LOAD_WARP = MMA_WARP + 1;
  static constexpr int TMEM_COLS = 2 * BN;                   // two accumulator buffers
  static constexpr int EPI_STAGE = NWG * 4 * 32 * 20 * 4;     // per epilogue warp: 32 rows x (16 + 4 pad) floats
  static constexpr int SMEM_DYN = STAGES * STAGE_BYTES + EPI_STAGE + 1024;
};

using namespace tc;

struct RowInfo { int pix; int iy, ix; };     // per tile row: input frame base pixel, top-left input coordinate

template <int EPI, int BN>
__global__ void __launch_bounds__(Cfg<BN>::NTHREADS, 1) tc_gemm_kernel(const GemmParams p, const float* __restrict__ Bimg, int KC,
                                                                       int tiles_m, int tiles_n) {
  using C = Cfg<BN>;
  constexpr int STAGES = C::STAGES, STAGE_BYTES = C::STAGE_BYTES, B_PANEL = C::B_PANEL, TMEM_COLS = C::TMEM_COLS, A_PANEL = C::A_PANEL;
  constexpr int MMA_WARP = C::MMA_WARP, LOAD_WARP = C::LOAD_WARP, NWG = C::NWG;
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t a_full[STAGES], b_full[STAGES], slot_free[STAGES], acc_full[2], acc_free[2];
  __shared__ uint32_t s_tmem_base;
  __shared__ RowInfo s_rows[3][BM];
  __shared__ float2 s_ln[8][BM];              // (mu, rstd) per tile row, ring over this CTA's tiles (producers run ahead of the epilogue)
  __shared__ float s_stat[NWG][16];
  __shared__ __align__(16) float s_biasv[1024];  // the whole bias vector, fetched once per CTA (the per-tile __ldg round trip was 1-3k cycles of an epilogue-bound tile)

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);          // warp-uniform by construction (the MMA warp relies on it)
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&a_full[s], NPROD); mbar_init(&b_full[s], 1); mbar_init(&slot_free[s], 1); }
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

  const int total_tiles = tiles_m * tiles_n;
  const int Ps = p.OHs * p.OWs;
  const int chunks_per_tap = p.Cin / BKP;

  if (warp < 8) {
    // =============================================================== A producers
    // Work items are (tile, k-panel) pairs flattened over this CTA's tiles, so the register prefetch keeps running
    // across tile boundaries (with K = 64 a tile is a single panel: per-tile prologues exposed the full load latency).
    const int c16 = tid & 7;            // 16-byte chunk inside the 128-byte row
    const int r0 = tid >> 3;            // rows r0 + 32 q, q = 0..3
    const int my_tiles = (total_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const int n_items = my_tiles * KC;
    int last_table = -1;
    float ln_s[4] = {0.f, 0.f, 0.f, 0.f}, ln_ss[4] = {0.f, 0.f, 0.f, 0.f};   // LayerNorm partial sums of this thread's 4 rows
    uint32_t it = 0;
    long long tp_wait = 0, tp_work = 0, tp_load = 0;       // trace accumulators (registers; written once at the end)

    // row table of this CTA's T-th tile (ring of 3: prefetch runs at most 2 items = 2 tiles ahead of the stores)
    auto ensure_table = [&](int T) {
      if (T <= last_table) return;
      const int tile = blockIdx.x + T * gridDim.x;
      const int m0 = (tile / tiles_n) * BM;
      if (tid < BM) {
        const int m = m0 + tid;
        RowInfo ri;
        if (m < p.M && p.perm_in) {
          ri.pix = seq_blocked_pixel(m, p.perm_pb, p.perm_F, p.P); ri.iy = 0; ri.ix = 0;
        } else if (m < p.M) {
          const int f = m / Ps, rem = m - f * Ps;
          const int i = rem / p.OWs, j = rem - i * p.OWs;
          ri.pix = f * p.IH * p.IW; ri.iy = i * p.in_stride; ri.ix = j * p.in_stride;
        } else {
          ri.pix = -1; ri.iy = 0; ri.ix = 0;
        }
        s_rows[T % 3][tid] = ri;
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      last_table = T;
    };
    auto load_item = [&](float4 (&v)[8], int g) {
      const bool trl = (p.trace != nullptr) && blockIdx.x == 0 && tid == 0;
      const long long tl0 = trl ? clock64() : 0;
      const int T = g / KC, kc = g - T * KC;
      ensure_table(T);
      const RowInfo* rows = s_rows[T % 3];
      const int tap = kc / chunks_per_tap;
      const int c0 = (kc - tap * chunks_per_tap) * BKP;
      const int dy = p.dy[tap], dx = p.dx[tap];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const RowInfo ri = rows[r0 + 32 * q];
        const int iy = ri.iy + dy, ix = ri.ix + dx;
        const bool ok = (ri.pix >= 0) && (iy >= 0) && (iy < p.IH) && (ix >= 0) && (ix < p.IW);
        if (ok) {
          const float4* src = reinterpret_cast<const float4*>(p.A + (size_t)(ri.pix + iy * p.IW + ix) * p.lda + c0) + 2 * c16;
          v[2 * q] = __ldg(src);
          v[2 * q + 1] = __ldg(src + 1);
        } else {
          v[2 * q] = make_float4(0.f, 0.f, 0.f, 0.f);
          v[2 * q + 1] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
      if (trl) tp_load += clock64() - tl0;
    };
    auto store_item = [&](const float4 (&v)[8]) {
      const int s = it % STAGES;
      const uint32_t round = it / STAGES;
      const bool tr = (p.trace != nullptr) && blockIdx.x == 0 && tid == 0;
      long long t0 = 0;
      if (tr) t0 = clock64();
      mbar_wait(&slot_free[s], (round & 1) ^ 1);
      if (tr) { const long long t1 = clock64(); p.trace[6] += (unsigned long long)(t1 - t0); t0 = t1; }
      uint8_t* a_hi = smem + s * STAGE_BYTES;
      uint8_t* a_lo = a_hi + A_PANEL;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        uint32_t h[4], l[4];
        split_f16x2(v[2 * q].x, v[2 * q].y, h[0], l[0]);
        split_f16x2(v[2 * q].z, v[2 * q].w, h[1], l[1]);
        split_f16x2(v[2 * q + 1].x, v[2 * q + 1].y, h[2], l[2]);
        split_f16x2(v[2 * q + 1].z, v[2 * q + 1].w, h[3], l[3]);
        const uint32_t off = swz(r0 + 32 * q, c16);
        *reinterpret_cast<uint4*>(a_hi + off) = make_uint4(h[0], h[1], h[2], h[3]);
        *reinterpret_cast<uint4*>(a_lo + off) = make_uint4(l[0], l[1], l[2], l[3]);
        if (p.ln_inline) {
          const float4 a = v[2 * q], b = v[2 * q + 1];
          ln_s[q] += ((a.x + a.y) + (a.z + a.w)) + ((b.x + b.y) + (b.z + b.w));
          ln_ss[q] += ((a.x * a.x + a.y * a.y) + (a.z * a.z + a.w * a.w)) + ((b.x * b.x + b.y * b.y) + (b.z * b.z + b.w * b.w));
        }
      }
      if (p.ln_inline) {
        const int Ts = (int)(it / (uint32_t)KC), kcs = (int)(it - (uint32_t)Ts * KC);
        if (kcs == KC - 1) {                                   // the row is complete: reduce over the 8 lanes that share it
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            float sx = ln_s[q], sxx = ln_ss[q];
#pragma unroll
            for (int o = 1; o < 8; o <<= 1) { sx += __shfl_xor_sync(0xffffffffu, sx, o); sxx += __shfl_xor_sync(0xffffffffu, sxx, o); }
            if (c16 == 0) {
              const float inv = 1.0f / (float)p.K;
              const float mu = sx * inv;
              const float var = fmaxf(sxx * inv - mu * mu, 0.f);
              s_ln[Ts & 7][r0 + 32 * q] = make_float2(mu, 1.0f / sqrtf(var + 1e-5f));
            }
            ln_s[q] = 0.f; ln_ss[q] = 0.f;
          }
        }
      }
      // no proxy fence here: fence.proxy.async compiles to MEMBAR.ALL.CTA + FENCE.VIEW.ASYNC and would make every
      // producer thread drain its outstanding prefetch loads once per panel (measured ~1000 cycles).  The st.shared
      // above and this arrive retire in order through the same shared-memory pipe; the issuer thread runs the proxy
      // fence after acquiring a_full (it has no loads in flight), before the async-proxy reads of tcgen05.mma.
      mbar_arrive_relaxed(&a_full[s]);
      if (tr) p.trace[7] += (unsigned long long)(clock64() - t0);
      ++it;
    };

    // Pre-split A (fp16 hi | lo planes written by split_rows_kernel): the panel rows are copied global -> shared with cp.async, no
    // register staging and no conversion.  Each thread's copies of a stage signal a_full through cp.async.mbarrier.arrive.noinc, so
    // the producers run up to STAGES panels ahead.  Used where one A panel feeds several n-tiles and the conversion was the
    // bottleneck (the 8x8-level 3x3 convolutions: 4 n-tiles, producers at ~4.5k cycles per panel against ~0.85k of MMA issue).
    if (p.A16h != nullptr) {
      for (int gi = 0; gi < n_items; ++gi, ++it) {
        const int T = gi / KC, kc = gi - T * KC;
        ensure_table(T);
        const RowInfo* rows = s_rows[T % 3];
        const int tap = kc / chunks_per_tap;
        const int c0 = (kc - tap * chunks_per_tap) * BKP;
        const int dy = p.dy[tap], dx = p.dx[tap];
        const int s = it % STAGES;
        mbar_wait(&slot_free[s], ((it / STAGES) & 1) ^ 1);
        const uint32_t a_hi = smem_u32(smem + s * STAGE_BYTES), a_lo = a_hi + A_PANEL;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const RowInfo ri = rows[r0 + 32 * q];
          const int iy = ri.iy + dy, ix = ri.ix + dx;
          const bool ok = (ri.pix >= 0) && (iy >= 0) && (iy < p.IH) && (ix >= 0) && (ix < p.IW);
          const size_t e = ok ? (size_t)(ri.pix + iy * p.IW + ix) * p.Cin + c0 + 8 * c16 : 0;
          const uint32_t off = swz(r0 + 32 * q, c16), nbytes = ok ? 16u : 0u;        // src-size 0: zero fill
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(a_hi + off), "l"(p.A16h + e), "r"(nbytes) : "memory");
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(a_lo + off), "l"(p.A16l + e), "r"(nbytes) : "memory");
        }
        asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(&a_full[s])) : "memory");
      }
    } else
    // global loads run ahead of the split/store in statically indexed register buffers:
    // two items ahead (3 buffers) for BN = 64, one item ahead (2 buffers) for BN = 128 (96-register budget)
    if constexpr (BN == 64) {
      float4 v0[8], v1[8], v2[8];
      if (n_items > 0) load_item(v0, 0);
      if (n_items > 1) load_item(v1, 1);
      for (int g = 0; g < n_items; g += 3) {
        if (g + 2 < n_items) load_item(v2, g + 2);
        store_item(v0);
        if (g + 1 < n_items) {
          if (g + 3 < n_items) load_item(v0, g + 3);
          store_item(v1);
        }
        if (g + 2 < n_items) {
          if (g + 4 < n_items) load_item(v1, g + 4);
          store_item(v2);
        }
      }
    } else {
      float4 v0[8], v1[8];
      if (n_items > 0) load_item(v0, 0);
      for (int g = 0; g < n_items; g += 2) {
        if (g + 1 < n_items) load_item(v1, g + 1);
        store_item(v0);
        if (g + 1 < n_items) {
          if (g + 2 < n_items) load_item(v0, g + 2);
          store_item(v1);
        }
      }
    }
    if (p.trace != nullptr && blockIdx.x == 0 && tid == 0) {
      p.trace[6] = (unsigned long long)tp_wait; p.trace[7] = (unsigned long long)tp_work; p.trace[11] = (unsigned long long)tp_load;
    }
  } else if (warp == LOAD_WARP) {
    // =============================================================== weight loader (pre-swizzled hi|lo images)
    if (lane == 0) {
      uint32_t it = 0;
      long long tl_wait = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int nt = tile % tiles_n;
        const uint8_t* src = reinterpret_cast<const uint8_t*>(Bimg) + (size_t)nt * KC * (2 * B_PANEL);
        for (int kc = 0; kc < KC; ++kc, ++it) {
          const int s = it % STAGES;
          const uint32_t round = it / STAGES;
          const bool tr = (p.trace != nullptr) && blockIdx.x == 0;
          long long t0 = 0;
          if (tr) t0 = clock64();
          mbar_wait(&slot_free[s], (round & 1) ^ 1);
          if (tr) tl_wait += clock64() - t0;
          mbar_arrive_expect_tx(&b_full[s], 2 * B_PANEL);
          bulk_copy_g2s(smem + s * STAGE_BYTES + 2 * A_PANEL, src + (size_t)kc * (2 * B_PANEL), 2 * B_PANEL, &b_full[s]);
        }
      }
      if (p.trace != nullptr && blockIdx.x == 0) p.trace[8] = (unsigned long long)tl_wait;
    }
  } else if (warp == MMA_WARP) {
    // =============================================================== MMA issuer: the whole warp, warp-uniform control flow, one elected
    // lane per MMA / commit (tc_common.cuh: elect_one)
    {
      const uint32_t tmem_base_u = __shfl_sync(0xffffffffu, tmem_base, 0);      // read from shared memory: make it a provably uniform value
      const uint32_t idesc = (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);   // D f32, A/B f16, K-major
      const int CH = (p.drain > 0 && p.drain < CHUNK) ? p.drain : CHUNK;
      uint32_t it = 0, cg = 0;
      const bool tr = (p.trace != nullptr) && blockIdx.x == 0;
      long long t_acc = 0, t_a = 0, t_b = 0, t_issue = 0, t0 = 0, t_begin = clock64();
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        for (int kc = 0; kc < KC; ++kc, ++it) {
          const int s = it % STAGES;
          const uint32_t round = it / STAGES;
          const bool chunk_first = (kc % CH) == 0;
          const bool chunk_last = ((kc % CH) == CH - 1) || (kc == KC - 1);
          const uint32_t buf = cg & 1;
          if (tr) t0 = clock64();
          if (chunk_first) {
            mbar_wait(&acc_free[buf], ((cg >> 1) & 1) ^ 1);
            tc_fence_after();
          }
          if (tr) { const long long t1 = clock64(); t_acc += t1 - t0; t0 = t1; }
          mbar_wait(&a_full[s], round & 1);
          if (tr) { const long long t1 = clock64(); t_a += t1 - t0; t0 = t1; }
          mbar_wait(&b_full[s], round & 1);
          if (tr) { const long long t1 = clock64(); t_b += t1 - t0; t0 = t1; }
          fence_proxy_async();          // generic-proxy operand writes of the producers -> async proxy (see store_item)
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + s * STAGE_BYTES);
          const uint64_t ahi = make_desc(sa), alo = make_desc(sa + A_PANEL);
          const uint64_t bhi = make_desc(sa + 2 * A_PANEL), blo = make_desc(sa + 2 * A_PANEL + B_PANEL);
          const uint32_t d = tmem_base_u + buf * BN;
#pragma unroll
          for (int j = 0; j < BKP / 16; ++j) {
            const uint64_t o = (uint64_t)(j * 2);               // +32 bytes per k-step, in 16-byte units
            tc_mma_f16_elected(d, alo + o, bhi + o, idesc, (chunk_first && j == 0) ? 0u : 1u);
            tc_mma_f16_elected(d, ahi + o, blo + o, idesc, 1u);
            tc_mma_f16_elected(d, ahi + o, bhi + o, idesc, 1u);
          }
          tc_commit_elected(&slot_free[s]);
          if (chunk_last) { tc_commit_elected(&acc_full[buf]); ++cg; }
          if (tr) t_issue += clock64() - t0;
        }
      }
      if (tr && lane == 0) {
        p.trace[0] = (unsigned long long)(clock64() - t_begin); p.trace[1] = it;
        p.trace[2] = (unsigned long long)t_acc; p.trace[3] = (unsigned long long)t_a;
        p.trace[4] = (unsigned long long)t_b; p.trace[5] = (unsigned long long)t_issue;
      }
    }
  } else {
    // =============================================================== accumulate + epilogue (warps 8 .. 8+4*NWG-1)
    // warpgroup wg owns output columns [64 wg, 64 wg + 64) of the tile; inside it warp ew = warp % 4 reads TMEM lanes 32 ew ..
    const int wg = (warp - 8) >> 2;
    const int ew = (warp - 8) & 3;
    const int row_in_tile = ew * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(ew * 32) << 16;
    const int etid = (tid - NPROD) & 127;
    float* s_st = s_stat[wg];
    const int bar_id = 2 + wg;
    constexpr int EN = 64;                                 // columns per epilogue thread
    float* wbuf = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES) + ((wg * 4 + ew) * 32 * 20);
    uint32_t cg = 0;
    int Te = -1;
    const bool bias_s = (p.bias != nullptr) && p.N <= 1024;
    if (bias_s) {
      for (int i = (tid - NPROD); i < p.N; i += 128 * NWG) s_biasv[i] = __ldg(p.bias + i);
      asm volatile("bar.sync 4, %0;" ::"n"(128 * NWG) : "memory");
    }
    long long te_wait = 0, te_final = 0, te_drain = 0, te_store = 0, te_pre = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      ++Te;
      const int m0 = (tile / tiles_n) * BM, n0 = (tile % tiles_n) * BN + wg * EN;
      float acc[EN];
#pragma unroll
      for (int i = 0; i < EN; ++i) acc[i] = 0.f;
      const int CH = (p.drain > 0 && p.drain < CHUNK) ? p.drain : CHUNK;
      const int nchunks = (KC + CH - 1) / CH;
      for (int c = 0; c < nchunks; ++c, ++cg) {
        const uint32_t buf = cg & 1;
        const bool tr = (p.trace != nullptr) && blockIdx.x == 0 && etid == 0;
        long long t0 = 0;
        if (tr) t0 = clock64();
        mbar_wait(&acc_full[buf], (cg >> 1) & 1);
        if (tr) { const long long t1 = clock64(); te_wait += t1 - t0; t0 = t1; }
        tc_fence_after();
#pragma unroll
        for (int q = 0; q < EN / 16; ++q) {
          float v[16];
          tmem_ld16(tmem_base + lane_addr + buf * BN + wg * EN + q * 16, v);
#pragma unroll
          for (int i = 0; i < 16; ++i) acc[q * 16 + i] += v[i];
        }
        tc_fence_before();
        mbar_arrive(&acc_free[buf]);
        if (tr) te_drain += clock64() - t0;
      }

      // ---------------------------------------------------------- final epilogue: this thread owns row m
      const bool tr_e = (p.trace != nullptr) && blockIdx.x == 0 && etid == 0;
      const long long te0 = tr_e ? clock64() : 0;
      const int m = m0 + row_in_tile;
      bool rv = m < p.M;
      const int mc = rv ? m : (p.M - 1);
      int opx = 0;
      if (p.perm_out) {
        opx = seq_blocked_out_pixel(mc, p.perm_pb, p.perm_F, p.P, p.perm_f_lo, p.perm_f_hi);
        if (opx < 0) { rv = false; opx = 0; }
      }
      const int f = mc / Ps, rem = mc - f * Ps;
      const int oi = rem / p.OWs, oj = rem - oi * p.OWs;
      const size_t opix = p.perm_out ? (size_t)opx
                                     : (size_t)(f * p.OH + oi * p.out_stride + p.oy0) * p.OW + oj * p.out_stride + p.ox0;
      const int srow = p.perm_in ? seq_blocked_pixel(mc, p.perm_pb, p.perm_F, p.P) : mc;     // pixel behind this row

      if (EPI == EPI_PLAIN) {
        const float sc = p.tc_scale;                          // undoes the exact power-of-two weight pre-scale
        if (bias_s) {
          const float4* bp = reinterpret_cast<const float4*>(s_biasv + n0);
#pragma unroll
          for (int i = 0; i < EN / 4; ++i) {
            const float4 b = bp[i];
            acc[4 * i] = fmaf(acc[4 * i], sc, b.x); acc[4 * i + 1] = fmaf(acc[4 * i + 1], sc, b.y);
            acc[4 * i + 2] = fmaf(acc[4 * i + 2], sc, b.z); acc[4 * i + 3] = fmaf(acc[4 * i + 3], sc, b.w);
          }
        } else if (p.bias) {
          const float4* bp = reinterpret_cast<const float4*>(p.bias + n0);
#pragma unroll
          for (int i = 0; i < EN / 4; ++i) {
            const float4 b = __ldg(bp + i);
            acc[4 * i] = fmaf(acc[4 * i], sc, b.x); acc[4 * i + 1] = fmaf(acc[4 * i + 1], sc, b.y);
            acc[4 * i + 2] = fmaf(acc[4 * i + 2], sc, b.z); acc[4 * i + 3] = fmaf(acc[4 * i + 3], sc, b.w);
          }
        } else {
#pragma unroll
          for (int i = 0; i < EN; ++i) acc[i] *= sc;
        }
        if (rv && p.Res) {
          const float4* rp = reinterpret_cast<const float4*>(p.Res + opix * p.ldr + n0);
#pragma unroll
          for (int i = 0; i < EN / 4; ++i) {
            const float4 r = rp[i];
            acc[4 * i] += r.x; acc[4 * i + 1] += r.y; acc[4 * i + 2] += r.z; acc[4 * i + 3] += r.w;
          }
        }
        {
          const long long ts0 = tr_e ? clock64() : 0;
          if (tr_e) te_pre += ts0 - te0;
          if (!(p.exp_shift & 64)) store_rows_coalesced(wbuf, acc, p.Out, opix, p.ldo, n0, rv, lane);     // bit 6 of the debug field: skip the stores (timing experiment)
          if (tr_e) te_store += clock64() - ts0;
        }
        if (p.stats != nullptr) {
          // GroupNorm partial statistics (U:230): per 8-column sub-block, reduced over the warp's 32 rows
          if (etid < 16) s_st[etid] = 0.f;
          asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
#pragma unroll
          for (int b8 = 0; b8 < EN / 8; ++b8) {
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
            const int glo = n0 / p.cpg, ghi = (n0 + EN - 1) / p.cpg;
            if (grp >= glo && grp <= ghi) atomicAdd(&p.stats[etid], (double)s_st[etid]);
          }
        }
      } else {
        // LayerNorm fold (see gemm.cu): v = rstd * (acc - mu * colsum)
        float mu, rs;
        if (p.ln_inline) { const float2 st = s_ln[Te & 7][row_in_tile]; mu = st.x; rs = st.y; }
        else { mu = p.rowstats[2 * (size_t)srow]; rs = p.rowstats[2 * (size_t)srow + 1]; }
        {
          const float fa = rs * p.tc_scale, fb = -rs * mu;     // rstd * (scale*acc - mu*colsum)
          const float4* wp = reinterpret_cast<const float4*>(p.wsum + n0);
#pragma unroll
          for (int i = 0; i < EN / 4; ++i) {
            const float4 w = __ldg(wp + i);
            acc[4 * i] = fmaf(fb, w.x, fa * acc[4 * i]); acc[4 * i + 1] = fmaf(fb, w.y, fa * acc[4 * i + 1]);
            acc[4 * i + 2] = fmaf(fb, w.z, fa * acc[4 * i + 2]); acc[4 * i + 3] = fmaf(fb, w.w, fa * acc[4 * i + 3]);
          }
        }
        if (EPI == EPI_QKV_TEMPORAL) {
          if (n0 < 512) {
            const int fr = srow / p.P;
            // this row's 16 (cos, sin) pairs serve both heads of the 64-column slice
            const float4* rp = reinterpret_cast<const float4*>(p.rot + (size_t)fr * 32);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 cs = __ldg(rp + j);                    // pairs 2j, 2j+1
#pragma unroll
              for (int hd = 0; hd < EN / 32; ++hd) {
                const int i = hd * 32 + 4 * j;
                const float x0 = acc[i], x1 = acc[i + 1], x2 = acc[i + 2], x3 = acc[i + 3];
                acc[i] = x0 * cs.x - x1 * cs.y; acc[i + 1] = x1 * cs.x + x0 * cs.y;
                acc[i + 2] = x2 * cs.z - x3 * cs.w; acc[i + 3] = x3 * cs.z + x2 * cs.w;
              }
            }
          }
        } else if (EPI == EPI_QKV_SLA) {
          if (n0 < 256) {
#pragma unroll
            for (int hd = 0; hd < EN / 32; ++hd) {
              float mx = acc[hd * 32];
#pragma unroll
              for (int i = 1; i < 32; ++i) mx = fmaxf(mx, acc[hd * 32 + i]);
              float sum = 0.f;
#pragma unroll
              for (int i = 0; i < 32; ++i) { acc[hd * 32 + i] = expf(acc[hd * 32 + i] - mx); sum += acc[hd * 32 + i]; }
              const float inv = p.q_post_scale / sum;
#pragma unroll
              for (int i = 0; i < 32; ++i) acc[hd * 32 + i] *= inv;
            }
          }
        }
        if (EPI == EPI_CA_GATE) {
          const int fr = mc / p.P;
          const int ca = n0 >> 6;
          const float* kq = p.kq + ((size_t)fr * 3 + ca) * 64;
          const float* nk = p.nkq + ca * 8;
#pragma unroll
          for (int hd = 0; hd < 8; ++hd) {
            float nrm2 = 0.f, dr = 0.f, dn = 0.f;
#pragma unroll
            for (int d = 0; d < 8; ++d) {
              const float q = acc[hd * 8 + d];
              nrm2 += q * q; dr += q * kq[hd * 8 + d]; dn += q * nk[d];
            }
            const float inv = 8.0f / fmaxf(sqrtf(nrm2), 1e-12f);
            const float sr = dr * inv, sn = dn * inv;
            const float mx = fmaxf(sr, sn);
            const float er = expf(sr - mx), en = expf(sn - mx);
            if (rv) p.gates[(size_t)m * 24 + ca * 8 + hd] = er / (er + en);
          }
        } else {
          const long long ts0 = tr_e ? clock64() : 0;
          if (!(p.exp_shift & 64)) store_rows_coalesced(wbuf, acc, p.Out, opix, p.ldo, n0, rv, lane);     // bit 6 of the debug field: skip the stores (timing experiment)
          if (tr_e) te_store += clock64() - ts0;
        }
      }
      if (tr_e) te_final += clock64() - te0;
    }
    if (p.trace != nullptr && blockIdx.x == 0 && etid == 0 && wg == 0) {
      p.trace[9] = (unsigned long long)te_wait; p.trace[10] = (unsigned long long)te_final; p.trace[12] = (unsigned long long)te_drain;
      p.trace[14] = (unsigned long long)te_store; p.trace[13] = (unsigned long long)te_pre;
    }
  }

  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == MMA_WARP) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS));
  }
}

template <int EPI, int BN>
int launch_t(const GemmParams& p, const float* Bimg, cudaStream_t st) {
  using C = Cfg<BN>;
  static bool attr_set = false;
  static int num_sms = 0;
  if (!attr_set) {
    DAWN_CUDA_OK(cudaFuncSetAttribute(tc_gemm_kernel<EPI, BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_DYN));
    int dev = 0;
    DAWN_CUDA_OK(cudaGetDevice(&dev));
    DAWN_CUDA_OK(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
    attr_set = true;
  }
  const int tiles_m = (p.M + BM - 1) / BM, tiles_n = p.N / BN;
  const int KC = p.K / BKP;
  const int grid = std::min(tiles_m * tiles_n, num_sms);
  tc_gemm_kernel<EPI, BN><<<grid, C::NTHREADS, C::SMEM_DYN, st>>>(p, Bimg, KC, tiles_m, tiles_n);
  DAWN_LAUNCH_OK();
  return 0;
}

template <int EPI>
int launch_bn(const GemmParams& p, const float* Bimg, cudaStream_t st) {
  return tc_tile_n(p.N) == 128 ? launch_t<EPI, 128>(p, Bimg, st) : launch_t<EPI, 64>(p, Bimg, st);
}

}  // namespace

// output-tile width used for a problem with N columns (also decides the weight image layout)
int tc_tile_n(int N) { return (N % 128 == 0) ? 128 : 64; }

bool tc_gemm_supported(const GemmParams& p, int epi) {
  if (epi == EPI_GN_APPLY) return false;
  if (p.b_batch_stride != 0 || p.rows_per_batch != p.M) return false;
  if (p.N % 64 != 0 || p.K % BKP != 0 || p.Cin % BKP != 0) return false;      // 64-channel panels
  if ((p.lda & 3) || (p.ldo & 3) || (p.Res && (p.ldr & 3))) return false;
  if (p.M < BM) return false;
  if (epi == EPI_PLAIN && p.stats && (p.cpg % 8 != 0)) return false;
  return true;
}

// Host: [K][ldb] fp32 weights -> per (n-tile, k-panel) shared-memory images: hi panel (BN rows x 64 fp16, 128-byte
// swizzled) | lo panel.  Weights are multiplied by the exact power of two `*scale` first so that the lo pieces
// stay fp16-normal; returns the image size in floats.
size_t tc_pack_weights(const float* Bkn, int K, int N, int ldb, std::vector<float>& out, float* scale) {
  const int BN = tc_tile_n(N);
  const int KC = K / BKP, NT = N / BN, PH = BN * 64;           // panel size in fp16 elements
  float mx = 0.f;
  for (int k = 0; k < K; ++k)
    for (int n = 0; n < N; ++n) mx = std::max(mx, std::fabs(Bkn[(size_t)k * ldb + n]));
  int e = 0;
  if (mx > 0.f) { std::frexp(mx, &e); }                        // mx = m * 2^e, m in [0.5, 1)
  const float sc = std::ldexp(1.0f, 11 - e);                   // max |w| * sc in [1024, 2048)
  *scale = sc;
  out.assign(((size_t)NT * KC * 2 * PH + 1) / 2, 0.f);
  __half* base = reinterpret_cast<__half*>(out.data());
  for (int nt = 0; nt < NT; ++nt)
    for (int kc = 0; kc < KC; ++kc) {
      __half* hi = base + ((size_t)nt * KC + kc) * 2 * PH;
      __half* lo = hi + PH;
      for (int n = 0; n < BN; ++n)
        for (int k = 0; k < BKP; ++k) {
          const float w = Bkn[(size_t)(kc * BKP + k) * ldb + nt * BN + n] * sc;
          const __half h = __float2half_rn(w);
          const __half l = __float2half_rn(w - __half2float(h));
          const int off = (n >> 3) * 512 + (n & 7) * 64 + (((k >> 3) ^ (n & 7)) << 3) + (k & 7);   // in fp16 elements
          hi[off] = h; lo[off] = l;
        }
    }
  return out.size();
}

int launch_tc_gemm(const GemmParams& p, const float* Bimg, int epi, cudaStream_t st) {
  if (!tc_gemm_supported(p, epi)) { set_last_error("launch_tc_gemm: unsupported geometry"); return -1; }
  switch (epi) {
    case EPI_PLAIN: return launch_bn<EPI_PLAIN>(p, Bimg, st);
    case EPI_QKV_TEMPORAL: return launch_bn<EPI_QKV_TEMPORAL>(p, Bimg, st);
    case EPI_QKV_SLA: return launch_bn<EPI_QKV_SLA>(p, Bimg, st);
    case EPI_QKV_MID: return launch_bn<EPI_QKV_MID>(p, Bimg, st);
    case EPI_CA_GATE: return launch_bn<EPI_CA_GATE>(p, Bimg, st);
  }
  set_last_error("launch_tc_gemm: bad epilogue id");
  return -1;
}

}  // namespace dawn
