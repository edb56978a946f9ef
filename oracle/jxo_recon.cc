// ORACLE (test infrastructure) — pixel reconstruction from the decoded integers (the second half of
// the self-decoder): dequantisation (libjxl dec_group.cc / quantizer-inl.h AdjustQuantBias), DC ->
// lowest frequencies, chroma-from-luma, inverse transforms, XYB -> linear -> sRGB (dec_xyb-inl.h,
// opsin_params) [UPSTREAM, recalled].  Used to report PSNR for the emitted codestreams (quality
// evidence for tier T2: the coded integers mean a sensible image).  Loop filters are off in the
// frames this repo emits, so nothing follows the inverse transform.  parity unpinned.
#include "jxo_frame.h"
#include "jxo_stages.h"

namespace jxo {

static inline float DequantBias(int c, int32_t q) {
  static const float kBias[4] = {1.0f - 0.05465007330715401f, 1.0f - 0.07005449891748593f,
                                 1.0f - 0.049935103337343655f, 0.145f};
  if (q == 0) return 0.0f;
  if (q == 1) return kBias[c];
  if (q == -1) return -kBias[c];
  return (float)q - kBias[3] / (float)q;
}

// sRGB code boundaries in linear light: code n is chosen when boundary[n-1] <= lin < boundary[n], with
// boundary[n] = EOTF((n + 0.5) / 255) evaluated by the same rational polynomial as the encoder's table.
void SrgbBoundaries(float b[255]) {
  static const float p[5] = {2.200248328e-04f, 1.043637593e-02f, 1.624820318e-01f, 7.961564959e-01f, 8.210152774e-01f};
  static const float q[5] = {2.631846970e-01f, 1.076976492e+00f, 4.987528350e-01f, -5.512498495e-02f, 6.521209011e-03f};
  for (int n = 0; n < 255; ++n) {
    const float x = ((float)n + 0.5f) / 255.0f;
    if (x > 0.04045f) {
      float yp = p[4], yq = q[4];
      for (int k = 3; k >= 0; --k) { yp = fmaf(yp, x, p[k]); yq = fmaf(yq, x, q[k]); }
      b[n] = yp / yq;
    } else {
      b[n] = x * (1.0f / 12.92f);
    }
  }
}

// inverse opsin matrix, computed in double from the forward matrix and rounded to float
void InverseOpsin(float inv[9]) {
  const double M[3][3] = {{0.30, 0.622, 0.078}, {0.23, 0.692, 0.078}, {0.24342268924547819, 0.20476744424496821, 0.55180986650955360}};
  const double a = M[0][0], b = M[0][1], c = M[0][2], d = M[1][0], e = M[1][1], g = M[1][2], h = M[2][0], i = M[2][1], j = M[2][2];
  const double det = a * (e * j - g * i) - b * (d * j - g * h) + c * (d * i - e * h);
  const double v[9] = {(e * j - g * i) / det, (c * i - b * j) / det, (b * g - c * e) / det,
                       (g * h - d * j) / det, (a * j - c * h) / det, (c * d - a * g) / det,
                       (d * i - e * h) / det, (b * h - a * i) / det, (a * e - b * d) / det};
  for (int k = 0; k < 9; ++k) inv[k] = (float)v[k];
}

// XYB sample -> three 8-bit sRGB codes; every step is a defined fp32 operation (shared with the CUDA path)
static inline void XybToSrgb8(float X, float Y, float B, const float inv[9], const float bnd[255], uint8_t out[3]) {
  const float kBias = 0.0037930732552754493f, kNegBiasCbrt = -0.15595420054924863f;
  const float l = (Y + X) - kNegBiasCbrt, m = (Y - X) - kNegBiasCbrt, s = B - kNegBiasCbrt;
  const float mix0 = (l * l) * l - kBias, mix1 = (m * m) * m - kBias, mix2 = (s * s) * s - kBias;
  for (int k = 0; k < 3; ++k) {
    const float lin = fmaf(inv[3 * k], mix0, fmaf(inv[3 * k + 1], mix1, inv[3 * k + 2] * mix2));
    int lo = 0, hi = 255;                       // number of boundaries <= lin
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (bnd[mid] <= lin) lo = mid + 1; else hi = mid; }
    out[k] = (uint8_t)lo;
  }
}

// f: a Frame holding dc_quant, acs, raw_qf, coeffs, cmap, q.* (from DecodeCodestream or from the encoder);
// rgb: h*w*3 bytes out
bool ReconstructRgb(const Frame& f, uint8_t* rgb) {
  const FrameDim& fd = f.fd;
  const EncTables& T = GetTables();
  const size_t nblk = (size_t)fd.bxs * fd.bys;
  const size_t plane = (size_t)fd.ys_pad * fd.pitch;
  std::vector<float> xyb[3];
  for (auto& p : xyb) p.assign(plane, 0.0f);
  const float inv_global_scale = 65536.0f / (float)f.q.global_scale;
  const float inv_quant_dc = inv_global_scale / (float)f.q.quant_dc;
  const float dc_step[3] = {inv_quant_dc * (1.0f / 4096.0f), inv_quant_dc * (1.0f / 512.0f), inv_quant_dc * (1.0f / 256.0f)};
  const float inv_qm[3] = {1.0f / powf(1.25f, (float)(f.q.x_qm_scale - 2)), 1.0f, 1.0f / powf(1.25f, (float)(f.q.b_qm_scale - 2))};
  // dequantised DC per block: Y, then X / B with the DC colour correlation (0 and 1.0 by default)
  std::vector<float> dc[3];
  for (auto& p : dc) p.assign(nblk, 0.0f);
  for (size_t i = 0; i < nblk; ++i) {
    const float y = (float)f.dc_quant[nblk + i] * dc_step[1];
    dc[1][i] = y;
    dc[0][i] = fmaf(0.0f, y, (float)f.dc_quant[i] * dc_step[0]);
    dc[2][i] = fmaf(1.0f, y, (float)f.dc_quant[2 * nblk + i] * dc_step[2]);
  }
  static const int slot_of_chan[3] = {1, 0, 2};   // coefficient slots are Y, X, B
  for (int by = 0; by < fd.bys; ++by) for (int bx = 0; bx < fd.bxs; ++bx) {
    const uint8_t a = f.acs[(size_t)by * fd.bxs + bx];
    if (!(a & 0x80)) continue;
    const int s = a & 0x7f, cx = kCoveredX[s], cy = kCoveredY[s], n = cx * cy, size = n * 64;
    const int kind = kQuantKind[s];
    if (T.dequant[kind].empty()) return false;
    const float* dq = T.dequant[kind].data();
    const std::vector<uint16_t>& order = T.order[kStrategyOrder[s]];
    const float inv_qac = inv_global_scale / (float)f.raw_qf[(size_t)by * fd.bxs + bx];
    const int tx = bx / 8, ty = by / 8;
    const float cfl[3] = {0.0f + (float)f.cmap[(size_t)ty * fd.txs + tx] / 84.0f, 0.0f,
                          1.0f + (float)f.cmap[(size_t)fd.txs * fd.tys + (size_t)ty * fd.txs + tx] / 84.0f};
    std::vector<float> coef[3];
    for (int c = 0; c < 3; ++c) coef[c].assign(size, 0.0f);
    for (int ci = 0; ci < 3; ++ci) {
      const int c = ci == 0 ? 1 : (ci == 1 ? 0 : 2);   // Y first: X and B add cfl * Y
      const int slot = slot_of_chan[c];
      const float mul = inv_qac * inv_qm[c];
      for (int k = 0; k < size; ++k) {
        const int j = k / 64;
        const int cbx = bx + (j % cx), cby = by + (j / cx);
        const int g = (cby / 32) * fd.gxs + (cbx / 32);
        const size_t blk = (size_t)g * 1024 + (size_t)(cby % 32) * 32 + (cbx % 32);
        const int32_t q = f.coeffs[(blk * 3 + slot) * 64 + (k % 64)];
        const int pos = order[k];
        float v = (DequantBias(c, q) * dq[(size_t)c * size + pos]) * mul;
        if (c != 1) v = fmaf(cfl[c], coef[1][pos], v);
        coef[c][pos] = v;
      }
    }
    for (int c = 0; c < 3; ++c) {
      // lowest frequencies from the DC image (after the chroma-from-luma term used the AC-only Y)
      float llf[64];
      LowestFrequenciesFromDc(s, &dc[c][(size_t)by * fd.bxs + bx], fd.bxs, llf);
      const int rows = cy * 8, cols = cx * 8;
      const bool transposed = rows >= cols;
      const int W = std::max(rows, cols);
      for (int vf = 0; vf < cy; ++vf) for (int hf = 0; hf < cx; ++hf) {
        const int pos = transposed ? hf * W + vf : vf * W + hf;
        coef[c][pos] = llf[vf * cx + hf];
      }
      TransformToPixels(s, coef[c].data(), &xyb[c][(size_t)by * 8 * fd.pitch + (size_t)bx * 8], fd.pitch);
    }
  }
  float inv[9], bnd[255];
  InverseOpsin(inv);
  SrgbBoundaries(bnd);
  for (int y = 0; y < fd.ysize; ++y) for (int x = 0; x < fd.xsize; ++x) {
    const size_t p = (size_t)y * fd.pitch + x;
    float v[3] = {xyb[0][p], xyb[1][p], xyb[2][p]};
    if (f.gab) for (int c = 0; c < 3; ++c) v[c] = GaborishBlurAt(xyb[c].data(), fd.pitch, fd.xsize, fd.ysize, x, y);
    XybToSrgb8(v[0], v[1], v[2], inv, bnd, &rgb[((size_t)y * fd.xsize + x) * 3]);
  }
  return true;
}

// Sum of squared errors per channel between the reconstruction and the original (the quantity behind
// calculate_mse / calculate_psnr of the harness, benchmark-jpegxl/src/image_reader.rs:555-606).
bool ReconstructionSse(const Frame& f, const uint8_t* orig, size_t stride, uint64_t sse[3]) {
  std::vector<uint8_t> rec((size_t)f.fd.xsize * f.fd.ysize * 3);
  if (!ReconstructRgb(f, rec.data())) return false;
  sse[0] = sse[1] = sse[2] = 0;
  for (int y = 0; y < f.fd.ysize; ++y) for (int x = 0; x < f.fd.xsize; ++x) for (int c = 0; c < 3; ++c) {
    const int d = (int)rec[((size_t)y * f.fd.xsize + x) * 3 + c] - (int)orig[(size_t)y * stride + 3 * x + c];
    sse[c] += (uint64_t)(d * d);
  }
  return true;
}

}  // namespace jxo
