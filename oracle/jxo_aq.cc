// ORACLE (test infrastructure) — stage U2: adaptive quantisation field.
// Restates the structure of libjxl lib/jxl/enc_adaptive_quantization.cc
// (InitialQuantField -> AdaptiveQuantizationMap: per-pixel masking, 4x4 pre-erosion,
// fuzzy erosion, per-block ComputeMask + HF / colour / gamma modulations, exp2 mapping)
// [UPSTREAM; SURVEY.md section 8a row U2].  parity unpinned: the tuned constants are
// recalled, and libm calls (log1p) are replaced by the repo's FastLog2f so that the
// CUDA path can be bit-identical.
#include "jxo_frame.h"
#include "jxo_stages.h"

namespace jxo {

static inline float RatioCbrtToGamma(float v, bool invert) {
  const float kEps = 1e-2f;
  const float kNumMul = 1.1990657e+02f;   // kSGRetMul * 3 * kSGmul
  const float kVOffset = 5.4044867e+00f;  // kSGVOffset * ln2 + eps
  const float kDenMul = 1.5718648e+02f;   // ln2 * kSGmul
  v = v > 0.0f ? v : 0.0f;
  const float v2 = v * v;
  const float num = fmaf(kNumMul, v2, kEps);
  const float den = fmaf(kDenMul * v, v2, kVOffset);
  return invert ? num / den : den / num;
}

static inline float MaskingSqrt(float v) {
  const float kLogOffset = 26.481471032459346f;
  const float kSqrtMul = 1.4543302e+05f;  // sqrt(211.50759899638012 * 1e8)
  return 0.25f * sqrtf(fmaf(v, kSqrtMul, kLogOffset));
}

static inline float ComputeMask(float out_val) {
  const float kBase = -0.7647f, kMul4 = 9.4708735624378946f, kMul2 = 17.35036561631863f;
  const float kOffset2 = 302.59587815579727f, kMul3 = 6.7943250517376494f, kOffset3 = 3.7179635626140772f;
  const float kOffset4 = 0.25f * kOffset3, kMul0 = 0.80061762862741759f;
  float v1 = out_val * kMul0;
  v1 = v1 > 1e-3f ? v1 : 1e-3f;
  const float v2 = 1.0f / (v1 + kOffset2);
  const float v3 = 1.0f / fmaf(v1, v1, kOffset3);
  const float v4 = 1.0f / fmaf(v1, v1, kOffset4);
  return kBase + fmaf(kMul4, v4, fmaf(kMul2, v2, kMul3 * v3));
}

// halving tree over 8 partial sums: the association of a __shfl_xor butterfly (4, 2, 1)
static inline float Tree8(float* p) {
  for (int st = 4; st >= 1; st /= 2) for (int i = 0; i < st; ++i) p[i] = p[i] + p[i + st];
  return p[0];
}

static inline void StoreMin4(float v, float& m0, float& m1, float& m2, float& m3) {
  if (v < m3) {
    if (v < m0) { m3 = m2; m2 = m1; m1 = m0; m0 = v; }
    else if (v < m1) { m3 = m2; m2 = m1; m1 = v; }
    else if (v < m2) { m3 = m2; m2 = v; }
    else { m3 = v; }
  }
}

void InitialQuantField(Frame* f) {
  const FrameDim& fd = f->fd;
  const int xs = fd.xs_pad, ys = fd.ys_pad, pitch = fd.pitch;
  const float d = f->params.distance;
  const float* X = f->xyb[0].data();
  const float* Y = f->xyb[1].data();
  const float* B = f->xyb[2].data();
  const float kMatchGammaOffset = 0.019f;

  // --- per-pixel: mask1x1 and the masking "diff" -----------------------------------
  const int pw = xs / 4, ph = ys / 4;
  std::vector<float> pre((size_t)pw * ph, 0.0f);
  std::vector<float> acc((size_t)xs, 0.0f);
  for (int y = 0; y < ys; ++y) {
    const int y1 = y > 0 ? y - 1 : y, y2 = y + 1 < ys ? y + 1 : y;
    const float* r = Y + (size_t)y * pitch;
    const float* r1 = Y + (size_t)y1 * pitch;
    const float* r2 = Y + (size_t)y2 * pitch;
    float* m1 = &f->mask1x1[(size_t)y * pitch];
    for (int x = 0; x < xs; ++x) {
      const int x1 = x > 0 ? x - 1 : x, x2 = x + 1 < xs ? x + 1 : x;
      const float base = 0.25f * (((r2[x] + r1[x]) + r[x1]) + r[x2]);
      const float gammac = RatioCbrtToGamma(r[x] + kMatchGammaOffset, false);
      float diff = gammac * (r[x] - base);
      // mask1x1 (consumed by the AC-strategy loss term)
      const float l1 = FastLog2f(1.0f + fabsf(diff)) * 0.69314718f;
      m1[x] = 1.0f / (l1 + 0.01f);
      // 4x4 pre-erosion input
      diff = diff * diff;
      if (diff >= 0.2f) diff = 0.2f;
      diff = MaskingSqrt(diff);
      if ((y % 4) != 0) acc[x] += diff; else acc[x] = diff;
    }
    if (y % 4 == 3) {
      float* po = &pre[(size_t)(y / 4) * pw];
      for (int x = 0; x < pw; ++x) po[x] = (((acc[4 * x] + acc[4 * x + 1]) + acc[4 * x + 2]) + acc[4 * x + 3]) * 0.25f;
    }
  }

  // --- fuzzy erosion: weighted 4 smallest of each 3x3, 2x2 cells summed per block -------
  float mul = 0.0f;
  if (d < 2.0f) mul = (2.0f - d) * 0.5f;
  float k0 = 0.125f + mul * 0.0f, k1 = 0.10f + mul * -0.10f, k2 = 0.09f + mul * -0.09f, k3 = 0.06f + mul * -0.06f;
  const float norm = 0.29959705784054957f / (((k0 + k1) + k2) + k3);
  k0 *= norm; k1 *= norm; k2 *= norm; k3 *= norm;
  std::vector<float> aq((size_t)fd.bxs * fd.bys, 0.0f);
  for (int y = 0; y < ph; ++y) {
    const int ym1 = y >= 1 ? y - 1 : y, yp1 = y + 1 < ph ? y + 1 : y;
    const float* rt = &pre[(size_t)ym1 * pw];
    const float* r = &pre[(size_t)y * pw];
    const float* rb = &pre[(size_t)yp1 * pw];
    for (int x = 0; x < pw; ++x) {
      const int xm1 = x >= 1 ? x - 1 : x, xp1 = x + 1 < pw ? x + 1 : x;
      float m0 = r[x], m1 = r[xm1], m2 = r[xp1], m3 = rt[xm1];
      if (m0 > m1) std::swap(m0, m1);
      if (m0 > m2) std::swap(m0, m2);
      if (m0 > m3) std::swap(m0, m3);
      if (m1 > m2) std::swap(m1, m2);
      if (m1 > m3) std::swap(m1, m3);
      if (m2 > m3) std::swap(m2, m3);
      StoreMin4(rt[x], m0, m1, m2, m3);
      StoreMin4(rt[xp1], m0, m1, m2, m3);
      StoreMin4(rb[xm1], m0, m1, m2, m3);
      StoreMin4(rb[x], m0, m1, m2, m3);
      StoreMin4(rb[xp1], m0, m1, m2, m3);
      const float v = ((k0 * m0 + k1 * m1) + k2 * m2) + k3 * m3;
      float& o = aq[(size_t)(y / 2) * fd.bxs + (x / 2)];
      if (x % 2 == 0 && y % 2 == 0) o = v; else o += v;
    }
  }

  // --- per-block modulations ---------------------------------------------------------
  const float scale = 0.841f / d;  // kAcQuant / butteraugli_target
  const float base_level = 0.48f * scale;
  float dampen = 1.0f;
  if (d >= 2.0f) { dampen = 1.0f - ((d - 2.0f) / (14.0f - 2.0f)); if (dampen < 0) dampen = 0; }
  const float qmul = scale * dampen;
  const float qadd = (1.0f - dampen) * base_level;
  f->qf_float.assign((size_t)fd.bxs * fd.bys, 0.0f);
  for (int by = 0; by < fd.bys; ++by) for (int bx = 0; bx < fd.bxs; ++bx) {
    const size_t bi = (size_t)by * fd.bxs + bx;
    f->mask[bi] = 1.0f / (aq[bi] + 0.001f);  // ComputeMaskForAcStrategyUse
    float out_val = ComputeMask(aq[bi]);
    const int x0 = bx * 8, y0 = by * 8;
    // HfModulation: clamped absolute differences to the right and below, Y channel
    {
      // per-column accumulators over the 8 rows, then a halving tree over the columns
      float col[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      const float valmin = 0.020602694503245016f;
      for (int dy = 0; dy < 8; ++dy) {
        const float* r = Y + (size_t)(y0 + dy) * pitch + x0;
        const float* rn = dy == 7 ? r : r + pitch;
        for (int dx = 0; dx < 8; ++dx) {
          if (dx < 7) { const float a = fabsf(r[dx] - r[dx + 1]); col[dx] += a < valmin ? a : valmin; }
          const float b = fabsf(r[dx] - rn[dx]);
          col[dx] += b < valmin ? b : valmin;
        }
      }
      const float sum = Tree8(col);
      out_val = (sum + -1.110929106987477f) * -0.38078920620238305f + out_val;
    }
    // ColorModulation
    {
      const float strength = 3.0f * (1.0f - 0.25f * d);
      if (!(strength < 0)) {
        const float kRedRampStart = 0.0073200141118951231f, kRedRampLength = 0.019421555948474039f;
        const float kBlueRampLength = 0.086890611400405895f, kBlueRampStart = 0.26973418507870539f;
        const float red_strength = strength * 5.992297772961519f, blue_strength = strength;
        out_val = out_val + strength * -0.009174542291185913f;
        float redc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, bluec[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (int dy = 0; dy < 8; ++dy) {
          const size_t o = (size_t)(y0 + dy) * pitch + x0;
          for (int dx = 0; dx < 8; ++dx) {
            float pxv = X[o + dx] - kRedRampStart; pxv = pxv > 0.0f ? pxv : 0.0f;
            float pbv = B[o + dx] - (Y[o + dx] + kBlueRampStart); pbv = pbv > 0.0f ? pbv : 0.0f;
            bluec[dx] += pbv < kBlueRampLength ? pbv : kBlueRampLength;
            redc[dx] += pxv < kRedRampLength ? pxv : kRedRampLength;
          }
        }
        float red = Tree8(redc), blue = Tree8(bluec);
        const float ratio = 30.610615782142737f;
        red = red < ratio * kRedRampLength ? red : ratio * kRedRampLength;
        red = red * (red_strength / ratio);
        blue = blue < ratio * kBlueRampLength ? blue : ratio * kBlueRampLength;
        blue = blue * (blue_strength / ratio);
        out_val = red + (blue + out_val);
      }
    }
    // GammaModulation
    {
      float oc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      for (int dy = 0; dy < 8; ++dy) {
        const size_t o = (size_t)(y0 + dy) * pitch + x0;
        for (int dx = 0; dx < 8; ++dx) {
          const float iny = Y[o + dx] + 0.16f, inx = X[o + dx];
          const float rr = RatioCbrtToGamma(iny - inx, true);
          const float rg = RatioCbrtToGamma(iny + inx, true);
          oc[dx] += 0.5f * (rr + rg);
        }
      }
      float overall = Tree8(oc);
      overall = overall * (1.0f / 64.0f);
      out_val = fmaf(1.00561336e-01f, FastLog2f(overall), out_val);
    }
    f->qf_float[bi] = FastPow2f(out_val * 1.442695041f) * qmul + qadd;
  }
}

// AdjustQuantField: every block of a multi-block transform gets the maximum of the covered cells.
void AdjustQuantField(Frame* f) {
  const FrameDim& fd = f->fd;
  for (int by = 0; by < fd.bys; ++by) for (int bx = 0; bx < fd.bxs; ++bx) {
    const uint8_t a = f->acs[(size_t)by * fd.bxs + bx];
    if (!(a & 0x80)) continue;
    const int s = a & 0x7f, cx = kCoveredX[s], cy = kCoveredY[s];
    if (cx == 1 && cy == 1) continue;
    float m = f->qf_float[(size_t)by * fd.bxs + bx];
    for (int iy = 0; iy < cy; ++iy) for (int ix = 0; ix < cx; ++ix) {
      const float v = f->qf_float[(size_t)(by + iy) * fd.bxs + bx + ix];
      if (v > m) m = v;
    }
    for (int iy = 0; iy < cy; ++iy) for (int ix = 0; ix < cx; ++ix) f->qf_float[(size_t)(by + iy) * fd.bxs + bx + ix] = m;
  }
}

}  // namespace jxo
