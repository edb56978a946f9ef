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

// f: a Frame filled by DecodeCodestream (dc_quant, acs, raw_qf, coeffs, cmap, q.*); rgb: h*w*3 bytes out
bool ReconstructRgb(const Frame& f, uint8_t* rgb) {
  const FrameDim& fd = f.fd;
  const EncTables& T = GetTables();
  const size_t nblk = (size_t)fd.bxs * fd.bys;
  const size_t plane = (size_t)fd.ys_pad * fd.pitch;
  std::vector<float> xyb[3];
  for (auto& p : xyb) p.assign(plane, 0.0f);
  const float inv_global_scale = 65536.0f / (float)f.q.global_scale;
  const float inv_quant_dc = inv_global_scale / (float)f.q.quant_dc;
  const float dc_step[3] = {inv_quant_dc / 4096.0f, inv_quant_dc / 512.0f, inv_quant_dc / 256.0f};
  const float qm_mul[3] = {powf(1.25f, (float)(f.q.x_qm_scale - 2)), 1.0f, powf(1.25f, (float)(f.q.b_qm_scale - 2))};
  // dequantised DC per block: Y, then X / B with the DC colour correlation (0 and 1.0 by default)
  std::vector<float> dc[3];
  for (auto& p : dc) p.assign(nblk, 0.0f);
  for (size_t i = 0; i < nblk; ++i) {
    const float y = (float)f.dc_quant[nblk + i] * dc_step[1];
    dc[1][i] = y;
    dc[0][i] = (float)f.dc_quant[i] * dc_step[0] + 0.0f * y;
    dc[2][i] = (float)f.dc_quant[2 * nblk + i] * dc_step[2] + 1.0f * y;
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
      for (int k = 0; k < size; ++k) {
        const int j = k / 64;
        const int cbx = bx + (j % cx), cby = by + (j / cx);
        const int g = (cby / 32) * fd.gxs + (cbx / 32);
        const size_t blk = (size_t)g * 1024 + (size_t)(cby % 32) * 32 + (cbx % 32);
        const int32_t q = f.coeffs[(blk * 3 + slot) * 64 + (k % 64)];
        const int pos = order[k];
        float v = DequantBias(c, q) * dq[(size_t)c * size + pos] * inv_qac / qm_mul[c];
        if (c != 1) v += cfl[c] * coef[1][pos];
        coef[c][pos] = v;
      }
      // lowest frequencies from the DC image
      float llf[16];
      LowestFrequenciesFromDc(s, &dc[c][(size_t)by * fd.bxs + bx], fd.bxs, llf);
      const int rows = cy * 8, cols = cx * 8;
      const bool transposed = rows >= cols;
      const int W = std::max(rows, cols);
      for (int vf = 0; vf < cy; ++vf) for (int hf = 0; hf < cx; ++hf) {
        // coefficient layout has the long side horizontal: (hf, vf) swap for tall / square blocks
        const int pos = transposed ? hf * W + vf : vf * W + hf;
        coef[c][pos] = llf[vf * cx + hf];
      }
      TransformToPixels(s, coef[c].data(), &xyb[c][(size_t)by * 8 * fd.pitch + (size_t)bx * 8], fd.pitch);
    }
  }
  // XYB -> linear RGB -> sRGB
  const double M[3][3] = {{0.30, 0.622, 0.078}, {0.23, 0.692, 0.078}, {0.24342268924547819, 0.20476744424496821, 0.55180986650955360}};
  double inv[3][3];
  {
    const double a = M[0][0], b = M[0][1], c = M[0][2], d = M[1][0], e = M[1][1], g = M[1][2], h = M[2][0], i = M[2][1], j = M[2][2];
    const double det = a * (e * j - g * i) - b * (d * j - g * h) + c * (d * i - e * h);
    inv[0][0] = (e * j - g * i) / det; inv[0][1] = (c * i - b * j) / det; inv[0][2] = (b * g - c * e) / det;
    inv[1][0] = (g * h - d * j) / det; inv[1][1] = (a * j - c * h) / det; inv[1][2] = (c * d - a * g) / det;
    inv[2][0] = (d * i - e * h) / det; inv[2][1] = (b * h - a * i) / det; inv[2][2] = (a * e - b * d) / det;
  }
  const double bias = 0.0037930732552754493, cb = cbrt(bias);
  for (int y = 0; y < fd.ysize; ++y) for (int x = 0; x < fd.xsize; ++x) {
    const size_t p = (size_t)y * fd.pitch + x;
    const double X = xyb[0][p], Y = xyb[1][p], B = xyb[2][p];
    const double lms[3] = {Y + X + cb, Y - X + cb, B + cb};
    double mix[3];
    for (int k = 0; k < 3; ++k) mix[k] = lms[k] * lms[k] * lms[k] - bias;
    for (int k = 0; k < 3; ++k) {
      double lin = inv[k][0] * mix[0] + inv[k][1] * mix[1] + inv[k][2] * mix[2];
      lin = lin < 0 ? 0 : (lin > 1 ? 1 : lin);
      const double srgb = lin <= 0.0031308 ? 12.92 * lin : 1.055 * pow(lin, 1.0 / 2.4) - 0.055;
      const double v = srgb * 255.0 + 0.5;
      rgb[((size_t)y * fd.xsize + x) * 3 + k] = (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
    }
  }
  return true;
}

}  // namespace jxo
