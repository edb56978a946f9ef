// ORACLE (test infrastructure) — whole-frame driver (libjxl enc_frame.cc / enc_heuristics.cc /
// enc_group.cc order of operations) [UPSTREAM]. parity unpinned.
#include "jxo_frame.h"
#include "jxo_stages.h"

namespace jxo {

static EncTables* Tables() {
  // (function-local static: initialised once even when the first encodes start on several threads at a time)
  static EncTables* t = [] { EncTables* p = new EncTables(); p->Init(); return p; }();
  return t;
}

const EncTables& GetTables() { return *Tables(); }

static void HomogeneityMap(Frame* f) {
  const FrameDim& fd = f->fd;
  HomogConfig hc;
  hc.rows[0] = f->xyb[0].data(); hc.rows[1] = f->xyb[1].data(); hc.rows[2] = f->xyb[2].data();
  hc.stride = (size_t)fd.pitch; hc.ysize = (size_t)fd.ys_pad;
  f->homog.assign((size_t)fd.bxs * fd.bys * 3, 0.0f);
  for (int by = 0; by < fd.bys; ++by) for (int bx = 0; bx < fd.bxs; ++bx) {
    float* o = &f->homog[((size_t)by * fd.bxs + bx) * 3];
    CalculateHomogeneitySimilarityIndices((size_t)bx * 8, (size_t)by * 8, f->params.distance, hc, o, o + 1, o + 2);
  }
}

// transform + quantise every first block; fills coeffs / dc / nzeros, updates raw_qf
static void ComputeCoefficients(Frame* f) {
  const FrameDim& fd = f->fd;
  const EncTables& T = GetTables();
  const size_t nblk = (size_t)fd.bxs * fd.bys;
  for (int c = 0; c < 3; ++c) f->dc[c].assign(nblk, 0.0f);
  f->coeffs.assign((size_t)fd.num_groups * 1024 * 3 * 64, 0);
  f->nzeros.assign(3 * nblk, 0);
  f->nz_count.assign(3 * nblk, 0);
  std::vector<int32_t> q32[3];
  for (int by = 0; by < fd.bys; ++by) for (int bx = 0; bx < fd.bxs; ++bx) {
    const uint8_t a = f->acs[(size_t)by * fd.bxs + bx];
    if (!(a & 0x80)) continue;
    const int s = a & 0x7f;
    const int cx = kCoveredX[s], cy = kCoveredY[s], n = cx * cy, size = n * 64;
    const float* px[3]; float* dco[3]; int32_t* out[3];
    for (int c = 0; c < 3; ++c) {
      px[c] = &f->xyb[c][(size_t)by * 8 * fd.pitch + (size_t)bx * 8];
      dco[c] = &f->dc[c][(size_t)by * fd.bxs + bx];
      q32[c].assign(size, 0);
      out[c] = q32[c].data();
    }
    const int tx = bx / 8, ty = by / 8;
    const float x_factor = 0.0f + (float)f->cmap[(size_t)ty * fd.txs + tx] / 84.0f;
    const float b_factor = 1.0f + (float)f->cmap[(size_t)fd.txs * fd.tys + (size_t)ty * fd.txs + tx] / 84.0f;
    int32_t quant = f->raw_qf[(size_t)by * fd.bxs + bx];
    ComputeCoefficientsBlock(T, f->q, s, px, fd.pitch, x_factor, b_factor, &quant, dco, fd.bxs, out);
    for (int iy = 0; iy < cy; ++iy) for (int ix = 0; ix < cx; ++ix) f->raw_qf[(size_t)(by + iy) * fd.bxs + bx + ix] = quant;
    // store in scan order; chunk j (64 scan positions) of channel c lives in covered block j's slot
    const std::vector<uint16_t>& order = T.order[kStrategyOrder[s]];
    static const int chan_of_slot[3] = {1, 0, 2};  // token order Y, X, B
    for (int slot = 0; slot < 3; ++slot) {
      const int c = chan_of_slot[slot];
      int nz = 0;
      for (int k = 0; k < size; ++k) {
        const int32_t v = out[c][order[k]];
        if (k >= n && v != 0) nz++;
        const int j = k / 64;  // covered block index, raster inside the transform
        const int cbx = bx + (j % cx), cby = by + (j / cx);
        const int g = (cby / 32) * fd.gxs + (cbx / 32);
        const size_t blk = (size_t)g * 1024 + (size_t)(cby % 32) * 32 + (cbx % 32);
        f->coeffs[(blk * 3 + slot) * 64 + (k % 64)] = (int16_t)v;
      }
      const int log2n = FloorLog2((uint32_t)n);
      const int shared = (nz + n - 1) >> log2n;
      f->nz_count[(size_t)c * nblk + (size_t)by * fd.bxs + bx] = (uint16_t)nz;
      for (int iy = 0; iy < cy; ++iy) for (int ix = 0; ix < cx; ++ix)
        f->nzeros[(size_t)c * nblk + (size_t)(by + iy) * fd.bxs + bx + ix] = (uint8_t)shared;
    }
  }
  // DC
  const float* dcp[3] = {f->dc[0].data(), f->dc[1].data(), f->dc[2].data()};
  std::vector<int32_t> dq[3];
  int32_t* dqo[3];
  for (int c = 0; c < 3; ++c) { dq[c].assign(nblk, 0); dqo[c] = dq[c].data(); }
  QuantizeDc(f->q, dcp, nblk, dqo);
  f->dc_quant.assign(3 * nblk, 0);
  for (int c = 0; c < 3; ++c) for (size_t i = 0; i < nblk; ++i) {
    int32_t v = dq[c][i];
    v = v > 32767 ? 32767 : (v < -32768 ? -32768 : v);
    f->dc_quant[(size_t)c * nblk + i] = (int16_t)v;
  }
}

bool EncodeFrame(const uint8_t* rgb, int w, int h, size_t stride, const Params& p, Frame* f) {
  if (w <= 0 || h <= 0 || !rgb) { f->error = "invalid image"; return false; }
  if (!(p.distance >= 0.01f && p.distance <= 25.0f)) { f->error = "distance out of range [0.01, 25]"; return false; }
  if (p.effort < 1 || p.effort > 9) { f->error = "effort out of range [1, 9]"; return false; }
  if (p.proposal > 3) { f->error = "unknown proposal"; return false; }
  f->params = p;
  f->fd.Set(w, h);
  const FrameDim& fd = f->fd;
  const size_t plane = (size_t)fd.ys_pad * fd.pitch;
  const size_t nblk = (size_t)fd.bxs * fd.bys;
  for (int c = 0; c < 3; ++c) f->xyb[c].assign(plane, 0.0f);
  RgbToXyb(rgb, w, h, stride, fd, f->xyb[0].data(), f->xyb[1].data(), f->xyb[2].data());

  // quant field
  f->q = QuantState();
  f->q.adjust_quant = p.effort >= 5;
  {
    int xq = 2;
    if (p.distance > 1.25f) xq++;
    if (p.distance > 9.0f) xq++;
    if (p.distance < 0.299f) xq++;
    f->q.x_qm_scale = xq; f->q.b_qm_scale = 2;
    f->q.x_qm_mul = powf(1.25f, (float)(xq - 2));
    f->q.b_qm_mul = 1.0f;
  }
  f->mask.assign(nblk, 0.0f);
  f->mask1x1.assign(plane, 0.0f);
  if (p.flags & kFlagUniformQf) {
    f->qf_float.assign(nblk, 0.841f / p.distance);
  } else {
    InitialQuantField(f);
  }
  ComputeGlobalScale(f->qf_float.data(), nblk, InitialQuantDC(p.distance), &f->q);

  // Gaborish (opt-in): the quant field above relies on the pre-sharpening values (libjxl enc_heuristics.cc order
  // [UPSTREAM]); the homogeneity map, the search and the coefficients see the sharpened planes
  f->gab = (p.flags & kFlagGaborish) != 0;
  if (f->gab) {
    float* planes[3] = {f->xyb[0].data(), f->xyb[1].data(), f->xyb[2].data()};
    GaborishInverse(fd, planes);
  }
  HomogeneityMap(f);

  f->cmap.assign((size_t)2 * fd.txs * fd.tys, 0);
  if (p.flags & kFlagCfl) ChromaFromLumaFit(f);   // (opt-in; on the planes the search sees)
  f->acs.assign(nblk, 0x80 | DCT);
  f->acs_entropy.assign(nblk, 0.0f);
  if (p.flags & kFlagForcedAcs) {
    if (f->forced_acs.size() != nblk) { f->error = "forced strategy map of the wrong size"; return false; }
    f->acs = f->forced_acs;
  } else if (!(p.flags & kFlagFixedDct8)) {
    AcStrategySearch(f);
  }
  AdjustQuantField(f);
  f->raw_qf.assign(nblk, 0);
  SetRawQuantField(f->qf_float.data(), nblk, f->q, f->raw_qf.data());

  ComputeCoefficients(f);
  if (!EntropyCodeFrame(f)) return false;
  return true;
}

}  // namespace jxo
