// ORACLE (test infrastructure) — stage U1: sRGB u8 -> linear -> XYB.
// Restates libjxl lib/jxl/enc_xyb.cc + cms/opsin_params.h + cms/transfer_functions-inl.h
// [UPSTREAM, not under /root/reference; SURVEY.md section 8a row U1, Appendix U.16].
// parity unpinned.  Reached in the reference only through
// benchmark-jpegxl/src/docker_manager.rs:136 (the cjxl invocation).
#include "jxo.h"

namespace jxo {

const uint8_t kCoveredX[27] = {1, 1, 1, 1, 2, 4, 1, 2, 1, 4, 2, 4, 1, 1, 1, 1, 1, 1, 8, 4, 8, 16, 8, 16, 32, 16, 32};
const uint8_t kCoveredY[27] = {1, 1, 1, 1, 2, 4, 2, 1, 4, 1, 4, 2, 1, 1, 1, 1, 1, 1, 8, 8, 4, 16, 16, 8, 32, 32, 16};
const uint8_t kStrategyOrder[27] = {0, 1, 1, 1, 2, 3, 4, 4, 5, 5, 6, 6, 1, 1, 1, 1, 1, 1, 7, 8, 8, 9, 10, 10, 11, 12, 12};
// DCT, IDENTITY, DCT2X2, DCT4X4, DCT16X16, DCT32X32, DCT8X16(x2), DCT8X32(x2), DCT16X32(x2), DCT4X8(x2), AFV(x4), ...
const uint8_t kQuantKind[27] = {0, 1, 2, 3, 4, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 10, 10, 11, 12, 12, 13, 14, 14, 15, 16, 16};

// sRGB EOTF: 4/4 rational polynomial above 0.04045, linear segment below
// (libjxl TF_SRGB::DisplayFromEncoded; Horner with fused multiply-adds).
void SrgbLut(float lut[256]) {
  static const float p[5] = {2.200248328e-04f, 1.043637593e-02f, 1.624820318e-01f, 7.961564959e-01f, 8.210152774e-01f};
  static const float q[5] = {2.631846970e-01f, 1.076976492e+00f, 4.987528350e-01f, -5.512498495e-02f, 6.521209011e-03f};
  for (int i = 0; i < 256; ++i) {
    float x = (float)i / 255.0f;
    float r;
    if (x > 0.04045f) {
      float yp = p[4], yq = q[4];
      for (int k = 3; k >= 0; --k) { yp = fmaf(yp, x, p[k]); yq = fmaf(yq, x, q[k]); }
      r = yp / yq;
    } else {
      r = x * (1.0f / 12.92f);
    }
    lut[i] = r;
  }
}

// cube root for x >= 0: bit-hack estimate of x^(-1/3), three Newton steps, then x*r^2.
float CbrtPos(float x) {
  if (!(x > 0.0f)) return 0.0f;
  float r = u2f(0x54A21D2Au - f2u(x) / 3u);
  const float x3 = x * (1.0f / 3.0f);
  const float k43 = 4.0f / 3.0f;
  for (int i = 0; i < 3; ++i) {
    float r2 = r * r;
    float r4 = r2 * r2;
    r = fmaf(-x3, r4, k43 * r);
  }
  return (r * r) * x;
}

void RgbToXyb(const uint8_t* rgb, int w, int h, size_t stride, const FrameDim& fd, float* px, float* py, float* pb) {
  float lut[256];
  SrgbLut(lut);
  const float kBias = 0.0037930732552754493f;
  const float kNegBiasCbrt = -0.15595420054924863f;
  const float m00 = 0.30f, m01 = 0.622f, m02 = 0.078f;
  const float m10 = 0.23f, m11 = 0.692f, m12 = 0.078f;
  const float m20 = 0.24342268924547819f, m21 = 0.20476744424496821f, m22 = 0.55180986650955360f;
  for (int y = 0; y < fd.ys_pad; ++y) {
    const int sy = y < h ? y : h - 1;
    const uint8_t* row = rgb + (size_t)sy * stride;
    float* rx = px + (size_t)y * fd.pitch;
    float* ry = py + (size_t)y * fd.pitch;
    float* rb = pb + (size_t)y * fd.pitch;
    for (int x = 0; x < fd.pitch; ++x) {
      if (x >= fd.xs_pad) { rx[x] = ry[x] = rb[x] = 0.0f; continue; }
      const int sx = x < w ? x : w - 1;
      const float r = lut[row[3 * sx + 0]], g = lut[row[3 * sx + 1]], b = lut[row[3 * sx + 2]];
      float mix0 = fmaf(m00, r, fmaf(m01, g, fmaf(m02, b, kBias)));
      float mix1 = fmaf(m10, r, fmaf(m11, g, fmaf(m12, b, kBias)));
      float mix2 = fmaf(m20, r, fmaf(m21, g, fmaf(m22, b, kBias)));
      mix0 = mix0 > 0.0f ? mix0 : 0.0f;
      mix1 = mix1 > 0.0f ? mix1 : 0.0f;
      mix2 = mix2 > 0.0f ? mix2 : 0.0f;
      const float L = CbrtPos(mix0) + kNegBiasCbrt;
      const float M = CbrtPos(mix1) + kNegBiasCbrt;
      const float S = CbrtPos(mix2) + kNegBiasCbrt;
      rx[x] = 0.5f * (L - M);
      ry[x] = 0.5f * (L + M);
      rb[x] = S;
    }
  }
}

// ---- Gaborish (row U3, opt-in: kFlagGaborish) ------------------------------------------------------------------------
// Decoder side (ISO/IEC 18181-1 loop filter, default weights): every channel is convolved with the 3x3 kernel
// [w2 w1 w2; w1 1 w1; w2 w1 w2] / (1 + 4 w1 + 4 w2), w1 = 0.115169525, w2 = 0.061248592 [UPSTREAM, recalled], mirrored at the
// image borders.  Encoder side: libjxl sharpens the XYB planes with a hand-tuned 5x5 kernel before the search
// (enc_gaborish.cc, constants not available offline); any kernel is a legal encoder choice, and this one is the
// least-squares inverse of the decoder's kernel (tools/gen_gab_inverse.py), six weights by symmetry class.
// Both are defined with a fixed association so that the CUDA kernels reproduce them bit for bit.
static inline int Mirror(int i, int n) {
  while (i < 0 || i >= n) i = i < 0 ? -i - 1 : 2 * n - 1 - i;
  return i;
}

static const float kGabInv[6] = {1.758123398e+00f, -1.690203995e-01f, -7.491233945e-02f,
                                 2.443357371e-02f, 1.462947764e-02f, 7.093405002e-04f};

// in place on the xs_pad x ys_pad region of the three planes (the pitch padding beyond xs_pad stays zero)
void GaborishInverse(const FrameDim& fd, float* planes[3]) {
  const int W = fd.xs_pad, H = fd.ys_pad;
  std::vector<float> src((size_t)W * H);
  for (int c = 0; c < 3; ++c) {
    float* pl = planes[c];
    for (int y = 0; y < H; ++y) for (int x = 0; x < W; ++x) src[(size_t)y * W + x] = pl[(size_t)y * fd.pitch + x];
    auto at = [&](int x, int y) { return src[(size_t)Mirror(y, H) * W + Mirror(x, W)]; };
    for (int y = 0; y < H; ++y) for (int x = 0; x < W; ++x) {
      const float s0 = at(x, y);
      const float s1 = (at(x - 1, y) + at(x + 1, y)) + (at(x, y - 1) + at(x, y + 1));
      const float s2 = (at(x - 1, y - 1) + at(x + 1, y - 1)) + (at(x - 1, y + 1) + at(x + 1, y + 1));
      const float s3 = (at(x - 2, y) + at(x + 2, y)) + (at(x, y - 2) + at(x, y + 2));
      const float s4 = ((at(x - 2, y - 1) + at(x + 2, y - 1)) + (at(x - 2, y + 1) + at(x + 2, y + 1))) +
                       ((at(x - 1, y - 2) + at(x + 1, y - 2)) + (at(x - 1, y + 2) + at(x + 1, y + 2)));
      const float s5 = (at(x - 2, y - 2) + at(x + 2, y - 2)) + (at(x - 2, y + 2) + at(x + 2, y + 2));
      float v = kGabInv[0] * s0;
      v = fmaf(kGabInv[1], s1, v); v = fmaf(kGabInv[2], s2, v); v = fmaf(kGabInv[3], s3, v);
      v = fmaf(kGabInv[4], s4, v); v = fmaf(kGabInv[5], s5, v);
      pl[(size_t)y * fd.pitch + x] = v;
    }
  }
}

// the decoder's blur of pixel (x, y) of one plane, mirrored at the borders of the xsize x ysize image
float GaborishBlurAt(const float* plane, int pitch, int xsize, int ysize, int x, int y) {
  const float w1 = 0.115169525f, w2 = 0.061248592f;
  const float norm = 1.0f / (1.0f + 4.0f * w1 + 4.0f * w2);
  const float wc = norm, we = w1 * norm, wd = w2 * norm;
  auto at = [&](int xx, int yy) { return plane[(size_t)Mirror(yy, ysize) * pitch + Mirror(xx, xsize)]; };
  const float s1 = (at(x - 1, y) + at(x + 1, y)) + (at(x, y - 1) + at(x, y + 1));
  const float s2 = (at(x - 1, y - 1) + at(x + 1, y - 1)) + (at(x - 1, y + 1) + at(x + 1, y + 1));
  return fmaf(wd, s2, fmaf(we, s1, wc * at(x, y)));
}

}  // namespace jxo
