// Self-test of the tcgen05 temporal-attention kernel (temporal_tc.cu) on random data: every stage of unit 0 / head 0 (projection
// accumulator, scores, attention output) and the final output of pixel 0 against a double-precision host computation, and all pixels
// against the mma.sync kernel (temporal_fused.cu) where that one supports the shape.  Diagnostic entry point, not on the product path.
#include <cmath>
#include <cstring>
#include <string>
#include <vector>
#include <algorithm>
#include "common.cuh"
#include "temporal_fused.cuh"
#include "temporal_tc.cuh"

using namespace dawn;

// Host-side view of the kernel's work decomposition (no GPU needed): segments of a sequence and the two row tiles of each.
// out: per segment 14 ints {w0, wn, qa, qb, tile0{r0, r1, q0, q1, kb}, tile1{...}}; returns the segment count (0 = unsupported).
extern "C" int dawn_temporal_tc_plan(int F, int band, int q_lo, int q_hi, int* out) {
  if (band < 1 || band > kTtcBandMax || F < 1 || q_lo < 0 || q_hi > F || q_lo >= q_hi) return 0;
  TtcSegment seg[kTtcMaxSeg];
  const int n = temporal_tc_plan(F, band, q_lo, q_hi, seg);
  for (int s = 0; s < n && out; ++s) {
    TtcTile t[2];
    ttc_tiles(seg[s], band, t);
    int* o = out + 14 * s;
    o[0] = seg[s].w0; o[1] = seg[s].wn; o[2] = seg[s].qa; o[3] = seg[s].qb;
    for (int j = 0; j < 2; ++j) { o[4 + 5 * j] = t[j].r0; o[5 + 5 * j] = t[j].r1; o[6 + 5 * j] = t[j].q0; o[7 + 5 * j] = t[j].q1; o[8 + 5 * j] = t[j].kb; }
  }
  return n;
}

extern "C" int dawn_selftest_temporal_tc(int F, int P, int band, int q_lo, int q_hi, float* err, float* max_abs_ref, unsigned long long* trace48,
                                         float* ms) {
  if (!err || !max_abs_ref) { set_last_error("null argument"); return -1; }
  for (int i = 0; i < 6; ++i) err[i] = -1.f;
  if (!temporal_tc_supported(64, F, band, q_lo, q_hi)) { set_last_error("selftest: unsupported shape"); return -1; }
  uint32_t seed = 4242u + 13u * F + 7u * P;
  auto rnd = [&]() { seed = seed * 1664525u + 1013904223u; return ((seed >> 8) & 0xFFFF) / 32768.0f - 1.0f; };
  const int nb = 2 * band + 1, Fq = q_hi - q_lo;
  std::vector<float> x((size_t)F * P * 64), wqkv((size_t)768 * 64), wout((size_t)64 * 256), bias((size_t)8 * nb), rot((size_t)F * 32), wsum(768);
  for (auto& v : x) v = rnd() * 2.0f + 0.3f;
  for (auto& v : wqkv) v = rnd() * 0.25f;
  for (auto& v : wout) v = rnd() * 0.1f;
  for (auto& v : bias) v = rnd() * 2.0f;
  for (int f = 0; f < F; ++f)
    for (int i = 0; i < 16; ++i) { const float ang = 0.37f * f * std::pow(10000.f, -i / 16.f); rot[(size_t)f * 32 + 2 * i] = std::cos(ang); rot[(size_t)f * 32 + 2 * i + 1] = std::sin(ang); }
  for (int n = 0; n < 768; ++n) { double s = 0; for (int k = 0; k < 64; ++k) s += wqkv[(size_t)n * 64 + k]; wsum[n] = (float)s; }

  std::vector<uint8_t> Wq, Wo;
  float iw = 1.f, io = 1.f;
  temporal_tc_pack(wqkv.data(), wout.data(), Wq, Wo, &iw, &io);
  std::vector<float> table;
  temporal_tc_table(bias.data(), band, table);
  std::vector<uint16_t> Fq16, Fo16;
  float fiw = 1.f, fio = 1.f;
  temporal_fused_pack(wqkv.data(), wout.data(), Fq16, Fo16, &fiw, &fio);

  std::vector<void*> own;
  auto dalloc = [&](size_t bytes, void** p) { if (cudaMalloc(p, bytes) != cudaSuccess) return false; own.push_back(*p); return true; };
  auto cleanup = [&]() { for (void* p : own) cudaFree(p); };
  const size_t dbg_n = (size_t)kTtcWindowMax * (96 + 130 + 33);
  float *dx, *dres, *o1, *o2, *dws, *drot, *dtab, *dbias, *ddbg; uint8_t *dWq, *dWo; uint16_t *dFq, *dFo;
  const size_t orows = (size_t)Fq * P;
  bool ok = dalloc(x.size() * 4, (void**)&dx) && dalloc(orows * 64 * 4, (void**)&dres) && dalloc(orows * 64 * 4, (void**)&o1) &&
            dalloc(orows * 64 * 4, (void**)&o2) && dalloc(768 * 4, (void**)&dws) && dalloc(rot.size() * 4, (void**)&drot) &&
            dalloc(table.size() * 4, (void**)&dtab) && dalloc(bias.size() * 4, (void**)&dbias) && dalloc(dbg_n * 4, (void**)&ddbg) &&
            dalloc(Wq.size(), (void**)&dWq) && dalloc(Wo.size(), (void**)&dWo) && dalloc(Fq16.size() * 2, (void**)&dFq) &&
            dalloc(Fo16.size() * 2, (void**)&dFo);
  if (!ok) { cleanup(); set_last_error("selftest: cudaMalloc failed"); return -2; }
  cudaMemcpy(dx, x.data(), x.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dres, x.data() + (size_t)q_lo * P * 64, orows * 64 * 4, cudaMemcpyHostToDevice);     // residual = the owned frames of x
  cudaMemcpy(dws, wsum.data(), 768 * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(drot, rot.data(), rot.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dtab, table.data(), table.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dbias, bias.data(), bias.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dWq, Wq.data(), Wq.size(), cudaMemcpyHostToDevice);
  cudaMemcpy(dWo, Wo.data(), Wo.size(), cudaMemcpyHostToDevice);
  cudaMemcpy(dFq, Fq16.data(), Fq16.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dFo, Fo16.data(), Fo16.size() * 2, cudaMemcpyHostToDevice);
  cudaMemset(o1, 0, orows * 64 * 4); cudaMemset(o2, 0, orows * 64 * 4); cudaMemset(ddbg, 0, dbg_n * 4);

  TemporalTcArgs a{};
  a.x = dx; a.ldx = 64; a.res = dres; a.ldr = 64; a.out = o1; a.ldo = 64; a.F = F; a.P = P; a.q_lo = q_lo; a.q_hi = q_hi;
  a.Wqkv = dWq; a.Wout = dWo; a.rot = drot; a.table = dtab; a.band = band; a.inv_wscale = iw; a.inv_oscale = io; a.dbg = ddbg;
  int rc = launch_temporal_tc(a, 0);
  if (rc == 0 && cudaDeviceSynchronize() != cudaSuccess) { set_last_error(std::string("selftest tc: ") + cudaGetErrorString(cudaGetLastError())); rc = -2; }
  if (rc == 0 && trace48 && ms) {
    // second, timed run without the debug dump and with the cycle trace of CTA 0
    unsigned long long* dtr = nullptr;
    if (!dalloc(48 * 8, (void**)&dtr)) { cleanup(); set_last_error("selftest: cudaMalloc failed"); return -2; }
    cudaMemset(dtr, 0, 48 * 8);
    TemporalTcArgs t = a;
    t.dbg = nullptr; t.trace = dtr;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0, 0);
    rc = launch_temporal_tc(t, 0);
    cudaEventRecord(e1, 0);
    if (rc == 0 && cudaDeviceSynchronize() != cudaSuccess) { set_last_error(std::string("selftest tc(2): ") + cudaGetErrorString(cudaGetLastError())); rc = -2; }
    if (rc == 0) { cudaEventElapsedTime(ms, e0, e1); cudaMemcpy(trace48, dtr, 48 * 8, cudaMemcpyDeviceToHost); }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
  }
  bool have_old = false;
  if (rc == 0 && temporal_fused_supported(64, F, band, q_lo, q_hi)) {
    TemporalFusedArgs b{};
    b.x = dx; b.ldx = 64; b.res = dres; b.ldr = 64; b.out = o2; b.ldo = 64; b.F = F; b.P = P; b.q_lo = q_lo; b.q_hi = q_hi;
    b.Wqkv = dFq; b.Wout = dFo; b.wsum = dws; b.rot = drot; b.bias = dbias; b.band = band; b.inv_wscale = fiw; b.inv_oscale = fio;
    rc = launch_temporal_fused(b, 0);
    if (rc == 0 && cudaDeviceSynchronize() != cudaSuccess) { set_last_error(std::string("selftest fused: ") + cudaGetErrorString(cudaGetLastError())); rc = -2; }
    have_old = rc == 0;
  }
  if (rc != 0) { cleanup(); return rc; }

  std::vector<float> r1(orows * 64), r2(orows * 64), dbg(dbg_n);
  cudaMemcpy(r1.data(), o1, r1.size() * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(r2.data(), o2, r2.size() * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(dbg.data(), ddbg, dbg_n * 4, cudaMemcpyDeviceToHost);
  cleanup();

  // ---------------------------------------------------------------- host reference for pixel 0 (double precision)
  const int pix = 0;
  std::vector<double> mu(F), rs(F);
  for (int f = 0; f < F; ++f) {
    const float* xr = &x[((size_t)f * P + pix) * 64];
    double s = 0; for (int k = 0; k < 64; ++k) s += xr[k];
    s /= 64; double v = 0; for (int k = 0; k < 64; ++k) v += (xr[k] - s) * (xr[k] - s);
    mu[f] = s; rs[f] = 1.0 / std::sqrt(v / 64 + 1e-5);
  }
  std::vector<double> q((size_t)F * 256), k((size_t)F * 256), v((size_t)F * 256), raw((size_t)F * 768);
  for (int f = 0; f < F; ++f) {
    const float* xr = &x[((size_t)f * P + pix) * 64];
    for (int n = 0; n < 768; ++n) {
      double acc = 0; for (int c = 0; c < 64; ++c) acc += (double)xr[c] * wqkv[(size_t)n * 64 + c];
      const double val = rs[f] * (acc - mu[f] * wsum[n]);
      raw[(size_t)f * 768 + n] = val;                      // what the projection accumulates: the kernel normalises the row first
      (n < 256 ? q[(size_t)f * 256 + n] : n < 512 ? k[(size_t)f * 256 + n - 256] : v[(size_t)f * 256 + n - 512]) = val;
    }
    for (int n = 0; n < 256; n += 2) {
      const int i = (n & 31) >> 1;
      const double co = rot[(size_t)f * 32 + 2 * i], si = rot[(size_t)f * 32 + 2 * i + 1];
      double a0 = q[(size_t)f * 256 + n], a1 = q[(size_t)f * 256 + n + 1];
      q[(size_t)f * 256 + n] = a0 * co - a1 * si; q[(size_t)f * 256 + n + 1] = a1 * co + a0 * si;
      a0 = k[(size_t)f * 256 + n]; a1 = k[(size_t)f * 256 + n + 1];
      k[(size_t)f * 256 + n] = a0 * co - a1 * si; k[(size_t)f * 256 + n + 1] = a1 * co + a0 * si;
    }
  }
  std::vector<double> oh((size_t)F * 256, 0.0);
  for (int h = 0; h < 8; ++h)
    for (int i = q_lo; i < q_hi; ++i) {
      const int j0 = std::max(0, i - band), j1 = std::min(F - 1, i + band);
      std::vector<double> s(j1 - j0 + 1);
      double m = -1e300;
      for (int j = j0; j <= j1; ++j) {
        double d = 0; for (int c = 0; c < 32; ++c) d += q[(size_t)i * 256 + h * 32 + c] * k[(size_t)j * 256 + h * 32 + c];
        s[j - j0] = d + bias[(size_t)h * nb + (j - i) + band]; m = std::max(m, s[j - j0]);
      }
      double l = 0; for (auto& e : s) { e = std::exp(e - m); l += e; }
      for (int j = j0; j <= j1; ++j)
        for (int c = 0; c < 32; ++c) oh[(size_t)i * 256 + h * 32 + c] += s[j - j0] / l * v[(size_t)j * 256 + h * 32 + c];
    }
  // stage errors for unit 0 (pixel 0, segment 0), head 0
  TtcSegment segs[kTtcMaxSeg];
  temporal_tc_plan(F, band, q_lo, q_hi, segs);
  const TtcSegment sg = segs[0];
  TtcTile tl[2];
  ttc_tiles(sg, band, tl);
  const double sq = 1.0 / iw;
  double e_proj = 0, m_proj = 0, e_s = 0, m_s = 0, e_o = 0, m_o = 0;
  for (int r = 0; r < sg.wn; ++r) {
    const int f = sg.w0 + r;
    for (int part = 0; part < 3; ++part)
      for (int c = 0; c < 32; ++c) {
        const double ref = raw[(size_t)f * 768 + part * 256 + c] * sq, got = dbg[(size_t)r * 96 + part * 32 + c];
        e_proj = std::max(e_proj, std::fabs(got - ref)); m_proj = std::max(m_proj, std::fabs(ref));
      }
  }
  const float* dS = dbg.data() + (size_t)kTtcWindowMax * 96;
  const float* dO = dS + (size_t)kTtcWindowMax * 130;
  const double LOG2E = 1.4426950408889634;
  for (int j = 0; j < 2; ++j)
    for (int r = tl[j].q0; r < tl[j].q1; ++r) {
      const int i = sg.w0 + r;
      const int key0 = (int)dS[(size_t)r * 130], climit = (int)dS[(size_t)r * 130 + 1];
      for (int c = 0; c < 128 && c < climit; ++c) {
        const int jf = sg.w0 + key0 + c;
        double d = 0; for (int t = 0; t < 32; ++t) d += q[(size_t)i * 256 + t] * k[(size_t)jf * 256 + t];
        d *= LOG2E;
        e_s = std::max(e_s, std::fabs(dS[(size_t)r * 130 + 2 + c] - d)); m_s = std::max(m_s, std::fabs(d));
      }
      const double l = dO[(size_t)r * 33];
      for (int c = 0; c < 32; ++c) {
        const double ref = oh[(size_t)i * 256 + c], got = dO[(size_t)r * 33 + 1 + c] / l;
        e_o = std::max(e_o, std::fabs(got - ref)); m_o = std::max(m_o, std::fabs(ref));
      }
    }
  err[0] = (float)(e_proj / std::max(m_proj, 1e-30)); err[1] = (float)e_s; err[2] = (float)e_o;
  // final output of pixel 0
  double e_y = 0, m_y = 0;
  for (int i = q_lo; i < q_hi; ++i)
    for (int c = 0; c < 64; ++c) {
      double y = x[((size_t)i * P + pix) * 64 + c];
      for (int n = 0; n < 256; ++n) y += oh[(size_t)i * 256 + n] * wout[(size_t)c * 256 + n];
      const double got = r1[((size_t)(i - q_lo) * P + pix) * 64 + c];
      e_y = std::max(e_y, std::fabs(got - y)); m_y = std::max(m_y, std::fabs(y));
    }
  err[3] = (float)e_y;
  if (have_old) {
    float md = 0.f;
    for (size_t i = 0; i < r1.size(); ++i) { const float d = std::fabs(r1[i] - r2[i]); md = (d > md || d != d) ? d : md; }
    err[4] = md;
  }
  float nan_count = 0.f;
  for (float f : r1) if (f != f) nan_count += 1.f;
  err[5] = nan_count;
  *max_abs_ref = (float)m_y;
  (void)m_s; (void)m_o;
  return 0;
}
