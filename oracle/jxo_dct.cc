// ORACLE (test infrastructure) — stage U5a: VarDCT forward / inverse transforms.
// Restates libjxl lib/jxl/dct-inl.h, dct_scales.h, enc_transforms-inl.h,
// dec_transforms-inl.h [UPSTREAM; SURVEY.md section 8a row U5, Appendix U.8].
// parity unpinned.
//
// Definition (Appendix U.8): forward 1-D  F[k] = (k ? sqrt2 : 1)/N * sum f[n] cos(pi (2n+1) k / 2N),
// computed with the recursive even/odd split libjxl uses (sum/difference halves, the odd half
// pre-multiplied by 1/(2 cos((i+1/2) pi / N)), "B" recombination).  2-D = horizontal pass
// then vertical pass.  Storage: rows >= cols -> out[hf*rows + vf], else out[vf*cols + hf]
// (long side horizontal, square blocks transposed).
// Numerics contract (round 2): the multiply-adds of the recombination steps are FUSED, as libjxl's
// MulAdd is on every FMA target: forward d[0] = fma(d[0], sqrt2, d[1]); inverse v[i] = fma(d[i], w, s[i]),
// v[n-1-i] = fma(-d[i], w, s[i]).  The CUDA path writes the same __fmaf_rn calls.
#include "jxo.h"

namespace jxo {

static const float kSqrt2 = 1.41421356237309504880f;

// 1 / (2 cos((i + 0.5) pi / n)) — literal tables printed by tools/gen_tables.py
static const float kWc4[2] = {5.411961e-01f, 1.306563e+00f};
static const float kWc8[4] = {5.097956e-01f, 6.013449e-01f, 8.999762e-01f, 2.5629156e+00f};
static const float kWc16[8] = {5.024193e-01f, 5.224986e-01f, 5.6694406e-01f, 6.468218e-01f, 7.881546e-01f, 1.0606776e+00f, 1.7224472e+00f, 5.1011486e+00f};
static const float kWc64[32] = {5.001506e-01f, 5.0135845e-01f, 5.037887e-01f, 5.0747114e-01f, 5.1245147e-01f, 5.187927e-01f, 5.265773e-01f, 5.3590983e-01f, 5.469204e-01f, 5.597698e-01f, 5.746552e-01f, 5.918185e-01f, 6.1155736e-01f, 6.3423896e-01f, 6.603198e-01f, 6.903721e-01f, 7.2512054e-01f, 7.6549417e-01f, 8.127021e-01f, 8.683447e-01f, 9.345836e-01f, 1.0144082e+00f, 1.1120716e+00f, 1.2338327e+00f, 1.3892939e+00f, 1.5939723e+00f, 1.874676e+00f, 2.2820501e+00f, 2.9246285e+00f, 4.084611e+00f, 6.7967505e+00f, 2.0373878e+01f};
static const float kWc32[16] = {5.00603e-01f, 5.0547093e-01f, 5.154473e-01f, 5.310426e-01f, 5.531039e-01f, 5.82935e-01f, 6.225041e-01f, 6.748083e-01f, 7.445363e-01f, 8.393496e-01f, 9.725682e-01f, 1.1694399e+00f, 1.4841646e+00f, 2.057781e+00f, 3.4076085e+00f, 1.0190008e+01f};
static float WcMul(int n, int i) {
  return n == 4 ? kWc4[i] : n == 8 ? kWc8[i] : n == 16 ? kWc16[i] : n == 32 ? kWc32[i] : kWc64[i];
}

// in-place-ish recursive DCT on contiguous buffer of n floats (unscaled)
static void DctRec(float* v, int n, float* tmp) {
  if (n == 1) return;
  if (n == 2) { float a = v[0] + v[1], b = v[0] - v[1]; v[0] = a; v[1] = b; return; }
  const int h = n / 2;
  float* s = tmp; float* d = tmp + h;
  for (int i = 0; i < h; ++i) { s[i] = v[i] + v[n - 1 - i]; d[i] = v[i] - v[n - 1 - i]; }
  for (int i = 0; i < h; ++i) d[i] = d[i] * WcMul(n, i);
  DctRec(s, h, tmp + n);
  DctRec(d, h, tmp + n);
  d[0] = fmaf(d[0], kSqrt2, d[1]);
  for (int i = 1; i + 1 < h; ++i) d[i] = d[i] + d[i + 1];
  for (int i = 0; i < h; ++i) { v[2 * i] = s[i]; v[2 * i + 1] = d[i]; }
}

static void IdctRec(float* v, int n, float* tmp) {
  if (n == 1) return;
  if (n == 2) { float a = v[0] + v[1], b = v[0] - v[1]; v[0] = a; v[1] = b; return; }
  const int h = n / 2;
  float* s = tmp; float* d = tmp + h;
  for (int i = 0; i < h; ++i) { s[i] = v[2 * i]; d[i] = v[2 * i + 1]; }
  IdctRec(s, h, tmp + n);
  for (int i = h - 1; i >= 1; --i) d[i] = d[i] + d[i - 1];
  d[0] = d[0] * kSqrt2;
  IdctRec(d, h, tmp + n);
  for (int i = 0; i < h; ++i) {
    const float w = WcMul(n, i);
    v[i] = fmaf(d[i], w, s[i]);
    v[n - 1 - i] = fmaf(-d[i], w, s[i]);
  }
}

void Dct1D(const float* in, int in_stride, float* out, int out_stride, int n) {
  float v[64], tmp[512];
  for (int i = 0; i < n; ++i) v[i] = in[i * in_stride];
  DctRec(v, n, tmp);
  const float sc = 1.0f / (float)n;
  for (int i = 0; i < n; ++i) out[i * out_stride] = v[i] * sc;
}

void Idct1D(const float* in, int in_stride, float* out, int out_stride, int n) {
  float v[64], tmp[512];
  for (int i = 0; i < n; ++i) v[i] = in[i * in_stride];
  IdctRec(v, n, tmp);
  for (int i = 0; i < n; ++i) out[i * out_stride] = v[i];
}

void Dct2D(const float* px, int px_stride, int rows, int cols, float* out) {
  std::vector<float> t((size_t)rows * cols);
  for (int y = 0; y < rows; ++y) Dct1D(px + (size_t)y * px_stride, 1, &t[(size_t)y * cols], 1, cols);  // t[y][hf]
  const bool transposed = rows >= cols;
  for (int hf = 0; hf < cols; ++hf) {
    if (transposed) Dct1D(&t[hf], cols, out + (size_t)hf * rows, 1, rows);          // out[hf*rows + vf]
    else            Dct1D(&t[hf], cols, out + hf, cols, rows);                     // out[vf*cols + hf]
  }
}

void Idct2D(const float* coef, int rows, int cols, float* px, int px_stride) {
  std::vector<float> t((size_t)rows * cols);  // t[y][hf]
  const bool transposed = rows >= cols;
  for (int hf = 0; hf < cols; ++hf) {
    if (transposed) Idct1D(coef + (size_t)hf * rows, 1, &t[hf], cols, rows);
    else            Idct1D(coef + hf, cols, &t[hf], cols, rows);
  }
  for (int y = 0; y < rows; ++y) Idct1D(&t[(size_t)y * cols], 1, px + (size_t)y * px_stride, 1, cols);
}

// resample scale between an N-point DCT's low frequencies and the M-point DCT of the
// N/M-box-averaged signal: sin(pi k / 2M) / ((N/M) sin(pi k / 2N)), 1 at k = 0.
static float ResampleScale(int n_from, int n_to, int k) {
  static const float kResample16_2[2] = {1.e+00f, 9.017642e-01f};
  static const float kResample32_4[4] = {1.e+00f, 9.7488683e-01f, 9.017642e-01f, 7.870549e-01f};
  static const float kResample64_8[8] = {1.e+00f, 9.936866e-01f, 9.7488683e-01f, 9.4401807e-01f, 9.017642e-01f, 8.490575e-01f, 7.870549e-01f, 7.1710813e-01f};
  if (n_to == 1) return 1.0f;
  return n_from == 16 ? kResample16_2[k] : n_from == 32 ? kResample32_4[k] : kResample64_8[k];
}

// libjxl DCT2TopBlock<S> on an 8-pitch block, in place: every 2x2 cell of the top-left SxS square becomes
// (sum, horizontal difference, vertical difference, diagonal difference) / 4, gathered into four S/2 quadrants
static void Dct2TopBlock(float* b, int S) {
  float t[64];
  const int h = S / 2;
  for (int y = 0; y < h; ++y) for (int x = 0; x < h; ++x) {
    const float c00 = b[y * 2 * 8 + x * 2], c01 = b[y * 2 * 8 + x * 2 + 1];
    const float c10 = b[(y * 2 + 1) * 8 + x * 2], c11 = b[(y * 2 + 1) * 8 + x * 2 + 1];
    t[y * 8 + x] = (c00 + c01 + c10 + c11) * 0.25f;
    t[y * 8 + h + x] = (c00 + c01 - c10 - c11) * 0.25f;
    t[(y + h) * 8 + x] = (c00 - c01 + c10 - c11) * 0.25f;
    t[(y + h) * 8 + h + x] = (c00 - c01 - c10 + c11) * 0.25f;
  }
  for (int y = 0; y < S; ++y) for (int x = 0; x < S; ++x) b[y * 8 + x] = t[y * 8 + x];
}
static void Idct2TopBlock(float* b, int S) {
  float t[64];
  const int h = S / 2;
  for (int y = 0; y < h; ++y) for (int x = 0; x < h; ++x) {
    const float c00 = b[y * 8 + x], c01 = b[y * 8 + h + x], c10 = b[(y + h) * 8 + x], c11 = b[(y + h) * 8 + h + x];
    t[y * 2 * 8 + x * 2] = c00 + c01 + c10 + c11;
    t[y * 2 * 8 + x * 2 + 1] = c00 + c01 - c10 - c11;
    t[(y * 2 + 1) * 8 + x * 2] = c00 - c01 + c10 - c11;
    t[(y * 2 + 1) * 8 + x * 2 + 1] = c00 - c01 - c10 + c11;
  }
  for (int y = 0; y < S; ++y) for (int x = 0; x < S; ++x) b[y * 8 + x] = t[y * 8 + x];
}

void TransformFromPixels(int s, const float* px, int ps, float* coef) {
  switch (s) {
    case DCT: Dct2D(px, ps, 8, 8, coef); return;
    case DCT16X16: Dct2D(px, ps, 16, 16, coef); return;
    case DCT32X32: Dct2D(px, ps, 32, 32, coef); return;
    case DCT16X8: Dct2D(px, ps, 16, 8, coef); return;
    case DCT8X16: Dct2D(px, ps, 8, 16, coef); return;
    case DCT32X16: Dct2D(px, ps, 32, 16, coef); return;
    case DCT16X32: Dct2D(px, ps, 16, 32, coef); return;
    case DCT32X8: Dct2D(px, ps, 32, 8, coef); return;
    case DCT8X32: Dct2D(px, ps, 8, 32, coef); return;
    case DCT64X64: Dct2D(px, ps, 64, 64, coef); return;
    case DCT64X32: Dct2D(px, ps, 64, 32, coef); return;
    case DCT32X64: Dct2D(px, ps, 32, 64, coef); return;
    case DCT2X2: {  // libjxl DCT2TopBlock<8>, <4>, <2>: three levels of 2x2 Hadamard averages
      for (int y = 0; y < 8; ++y) for (int x = 0; x < 8; ++x) coef[y * 8 + x] = px[y * ps + x];
      Dct2TopBlock(coef, 8); Dct2TopBlock(coef, 4); Dct2TopBlock(coef, 2);
      return;
    }
    case IDENTITY: {  // libjxl enc_transforms-inl.h Type::IDENTITY
      for (int y = 0; y < 2; ++y) for (int x = 0; x < 2; ++x) {
        float block_dc = 0.0f;
        for (int iy = 0; iy < 4; ++iy) for (int ix = 0; ix < 4; ++ix) block_dc += px[(y * 4 + iy) * ps + x * 4 + ix];
        block_dc *= 1.0f / 16;
        for (int iy = 0; iy < 4; ++iy) for (int ix = 0; ix < 4; ++ix) {
          if (ix == 1 && iy == 1) continue;
          coef[(y + iy * 2) * 8 + x + ix * 2] = px[(y * 4 + iy) * ps + x * 4 + ix] - px[(y * 4 + 1) * ps + x * 4 + 1];
        }
        coef[(y + 2) * 8 + x + 2] = coef[y * 8 + x];
        coef[y * 8 + x] = block_dc;
      }
      const float b00 = coef[0], b01 = coef[1], b10 = coef[8], b11 = coef[9];
      coef[0] = (b00 + b01 + b10 + b11) * 0.25f;
      coef[1] = (b00 + b01 - b10 - b11) * 0.25f;
      coef[8] = (b00 - b01 + b10 - b11) * 0.25f;
      coef[9] = (b00 - b01 - b10 + b11) * 0.25f;
      return;
    }
    case DCT4X4: {
      for (int y = 0; y < 2; ++y) for (int x = 0; x < 2; ++x) {
        float d[16];
        Dct2D(px + y * 4 * ps + x * 4, ps, 4, 4, d);
        for (int iy = 0; iy < 4; ++iy) for (int ix = 0; ix < 4; ++ix)
          coef[(y + iy * 2) * 8 + x + ix * 2] = d[iy * 4 + ix];
      }
      const float b00 = coef[0], b01 = coef[1], b10 = coef[8], b11 = coef[9];
      coef[0] = (b00 + b01 + b10 + b11) * 0.25f;
      coef[1] = (b00 + b01 - b10 - b11) * 0.25f;
      coef[8] = (b00 - b01 + b10 - b11) * 0.25f;
      coef[9] = (b00 - b01 - b10 + b11) * 0.25f;
      return;
    }
    case DCT4X8: {  // two 8-row x 4-col halves, side by side (left / right)
      for (int x = 0; x < 2; ++x) {
        float d[32];
        Dct2D(px + x * 4, ps, 8, 4, d);  // rows>=cols -> d[hf*8 + vf], 4 rows x 8 cols
        for (int iy = 0; iy < 4; ++iy) for (int ix = 0; ix < 8; ++ix)
          coef[(x + iy * 2) * 8 + ix] = d[iy * 8 + ix];
      }
      const float b0 = coef[0], b1 = coef[8];
      coef[0] = (b0 + b1) * 0.5f;
      coef[8] = (b0 - b1) * 0.5f;
      return;
    }
    case DCT8X4: {  // two 4-row x 8-col halves, stacked (top / bottom)
      for (int y = 0; y < 2; ++y) {
        float d[32];
        Dct2D(px + y * 4 * ps, ps, 4, 8, d);  // rows<cols -> d[vf*8 + hf]
        for (int iy = 0; iy < 4; ++iy) for (int ix = 0; ix < 8; ++ix)
          coef[(y + iy * 2) * 8 + ix] = d[iy * 8 + ix];
      }
      const float b0 = coef[0], b1 = coef[8];
      coef[0] = (b0 + b1) * 0.5f;
      coef[8] = (b0 - b1) * 0.5f;
      return;
    }
    default: return;  // AFV0-3 (basis not derivable offline) and 128+ are outside the emitted set
  }
}

void TransformToPixels(int s, const float* coef, float* px, int ps) {
  switch (s) {
    case DCT: Idct2D(coef, 8, 8, px, ps); return;
    case DCT16X16: Idct2D(coef, 16, 16, px, ps); return;
    case DCT32X32: Idct2D(coef, 32, 32, px, ps); return;
    case DCT16X8: Idct2D(coef, 16, 8, px, ps); return;
    case DCT8X16: Idct2D(coef, 8, 16, px, ps); return;
    case DCT32X16: Idct2D(coef, 32, 16, px, ps); return;
    case DCT16X32: Idct2D(coef, 16, 32, px, ps); return;
    case DCT32X8: Idct2D(coef, 32, 8, px, ps); return;
    case DCT8X32: Idct2D(coef, 8, 32, px, ps); return;
    case DCT64X64: Idct2D(coef, 64, 64, px, ps); return;
    case DCT64X32: Idct2D(coef, 64, 32, px, ps); return;
    case DCT32X64: Idct2D(coef, 32, 64, px, ps); return;
    case DCT2X2: {  // libjxl IDCT2TopBlock<2>, <4>, <8>
      float c[64];
      memcpy(c, coef, sizeof(c));
      Idct2TopBlock(c, 2); Idct2TopBlock(c, 4); Idct2TopBlock(c, 8);
      for (int y = 0; y < 8; ++y) for (int x = 0; x < 8; ++x) px[y * ps + x] = c[y * 8 + x];
      return;
    }
    case IDENTITY: {  // libjxl dec_transforms-inl.h Type::IDENTITY
      const float b00 = coef[0], b01 = coef[1], b10 = coef[8], b11 = coef[9];
      const float dcs[4] = {b00 + b01 + b10 + b11, b00 + b01 - b10 - b11, b00 - b01 + b10 - b11, b00 - b01 - b10 + b11};
      for (int y = 0; y < 2; ++y) for (int x = 0; x < 2; ++x) {
        const float block_dc = dcs[y * 2 + x];
        float residual_sum = 0.0f;
        for (int iy = 0; iy < 4; ++iy) for (int ix = 0; ix < 4; ++ix) {
          if (ix == 0 && iy == 0) continue;
          residual_sum += coef[(y + iy * 2) * 8 + x + ix * 2];
        }
        const float mid = block_dc - residual_sum * (1.0f / 16);
        px[(4 * y + 1) * ps + 4 * x + 1] = mid;
        for (int iy = 0; iy < 4; ++iy) for (int ix = 0; ix < 4; ++ix) {
          if (ix == 1 && iy == 1) continue;
          px[(y * 4 + iy) * ps + x * 4 + ix] = coef[(y + iy * 2) * 8 + x + ix * 2] + mid;
        }
        px[y * 4 * ps + x * 4] = coef[(y + 2) * 8 + x + 2] + mid;
      }
      return;
    }
    case DCT4X4: {
      float c[64];
      memcpy(c, coef, sizeof(c));
      const float b00 = c[0], b01 = c[1], b10 = c[8], b11 = c[9];
      c[0] = b00 + b01 + b10 + b11;
      c[1] = b00 + b01 - b10 - b11;
      c[8] = b00 - b01 + b10 - b11;
      c[9] = b00 - b01 - b10 + b11;
      for (int y = 0; y < 2; ++y) for (int x = 0; x < 2; ++x) {
        float d[16];
        for (int iy = 0; iy < 4; ++iy) for (int ix = 0; ix < 4; ++ix) d[iy * 4 + ix] = c[(y + iy * 2) * 8 + x + ix * 2];
        Idct2D(d, 4, 4, px + y * 4 * ps + x * 4, ps);
      }
      return;
    }
    case DCT4X8: {
      float c[64];
      memcpy(c, coef, sizeof(c));
      const float b0 = c[0], b1 = c[8];
      c[0] = b0 + b1; c[8] = b0 - b1;
      for (int x = 0; x < 2; ++x) {
        float d[32];
        for (int iy = 0; iy < 4; ++iy) for (int ix = 0; ix < 8; ++ix) d[iy * 8 + ix] = c[(x + iy * 2) * 8 + ix];
        Idct2D(d, 8, 4, px + x * 4, ps);
      }
      return;
    }
    case DCT8X4: {
      float c[64];
      memcpy(c, coef, sizeof(c));
      const float b0 = c[0], b1 = c[8];
      c[0] = b0 + b1; c[8] = b0 - b1;
      for (int y = 0; y < 2; ++y) {
        float d[32];
        for (int iy = 0; iy < 4; ++iy) for (int ix = 0; ix < 8; ++ix) d[iy * 8 + ix] = c[(y + iy * 2) * 8 + ix];
        Idct2D(d, 4, 8, px + y * 4 * ps, ps);
      }
      return;
    }
    default: return;
  }
}

// DC of every covered 8x8 block from the cy x cx lowest frequencies
// (libjxl DCFromLowestFrequencies / ReinterpretingIDCT).
void DcFromLowestFrequencies(int s, const float* coef, float* dc, int dc_stride) {
  const int cx = kCoveredX[s], cy = kCoveredY[s];
  if (cx == 1 && cy == 1) { dc[0] = coef[0]; return; }
  const int rows = cy * 8, cols = cx * 8;
  const bool transposed = rows >= cols;
  float llf[64];  // [vf][hf], cy x cx
  for (int vf = 0; vf < cy; ++vf) for (int hf = 0; hf < cx; ++hf) {
    const float c = transposed ? coef[(size_t)hf * rows + vf] : coef[(size_t)vf * cols + hf];
    llf[vf * cx + hf] = c * ResampleScale(rows, cy, vf) * ResampleScale(cols, cx, hf);
  }
  // inverse cy x cx DCT in plain [vf][hf] layout: horizontal-inverse last (mirror of Idct2D)
  float t[64];
  for (int hf = 0; hf < cx; ++hf) Idct1D(&llf[hf], cx, &t[hf], cx, cy);
  for (int y = 0; y < cy; ++y) Idct1D(&t[y * cx], 1, dc + (size_t)y * dc_stride, 1, cx);
}

void LowestFrequenciesFromDc(int s, const float* dc, int dc_stride, float* llf_out) {
  const int cx = kCoveredX[s], cy = kCoveredY[s];
  if (cx == 1 && cy == 1) { llf_out[0] = dc[0]; return; }
  const int rows = cy * 8, cols = cx * 8;
  float t[64], f[64];
  for (int y = 0; y < cy; ++y) Dct1D(dc + (size_t)y * dc_stride, 1, &t[y * cx], 1, cx);
  for (int hf = 0; hf < cx; ++hf) Dct1D(&t[hf], cx, &f[hf], cx, cy);
  for (int vf = 0; vf < cy; ++vf) for (int hf = 0; hf < cx; ++hf)
    llf_out[vf * cx + hf] = f[vf * cx + hf] / (ResampleScale(rows, cy, vf) * ResampleScale(cols, cx, hf));
}

}  // namespace jxo
