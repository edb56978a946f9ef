// ORACLE (test infrastructure) — rows H1-H9: the thesis' homogeneity metric.
// Line-by-line restatement of proposals/homogeneity-partitioning.diff (identical
// ranges in homogeneity-factored-entropy.diff and combined.diff).  This is the only
// arithmetic on the hot path that exists under /root/reference; it is pinned by the
// hand-computed vectors in tests/golden/homogeneity_*.json (the reference has none).
//
// Defined behaviour for the reference's undefined / ambiguous spots:
//  * H4 (diff :91): `x + j - 1 < 0` / `y + i - 1 < 0` are vacuous on size_t, so the
//    reference reads index -1 at image column 0 / row 0.  Here: such pixels are SKIPPED
//    (the evident intent of the guard).
//  * unqualified abs() on float (diff :100-101) -> fabsf.
//  * unqualified sqrt() with a double literal 0.3 (diff :148-149) -> evaluated in double,
//    narrowed once on assignment.
//  * `acc += a * b` statements are contracted to one fused multiply-add (clang
//    -ffp-contract=on in an FMA-capable Highway target); written here as fmaf().
//  * horizontal bound is `src_stride` (row pitch), not the width (diff :73, :91): the
//    pitch padding (zero-filled in this build's planes) is read at the right edge.
#include "jxo.h"

namespace jxo {

static inline float Px(const HomogConfig& c, int ch, size_t x, size_t y) { return c.rows[ch][y * c.stride + x]; }

// proposals/homogeneity-partitioning.diff:17-55
size_t CalculateNumZeroCrossings(size_t xsize, size_t ysize, float threshold, const float* laplacian) {
  size_t num_h = 0;
  for (size_t i = 0; i < ysize; i++) {
    bool in_edge = false;
    for (size_t j = 0; j < xsize; j++) {
      const float v = laplacian[i * xsize + j];
      if (!in_edge && v > threshold) { num_h++; in_edge = true; }
      else if (in_edge && v <= threshold) { in_edge = false; }
    }
  }
  const float avg_h = num_h / (float)ysize;
  size_t num_v = 0;
  for (size_t i = 0; i < xsize; i++) {
    bool in_edge = false;
    for (size_t j = 0; j < ysize; j++) {
      const float v = laplacian[j * xsize + i];
      if (!in_edge && v > threshold) { num_v++; in_edge = true; }
      else if (in_edge && v <= threshold) { in_edge = false; }
    }
  }
  const float avg_v = num_v / (float)xsize;
  return (size_t)(avg_h + avg_v);  // float -> size_t truncation (diff :54)
}

// proposals/homogeneity-partitioning.diff:57-81
void CalculateLaplacianFilter(size_t x, size_t y, size_t xsize, size_t ysize, size_t bx, size_t by,
                              const HomogConfig& c, float* out) {
  static const float mask[3][3] = {{0, -1, 0}, {-1, -4, -1}, {0, -1, 0}};
  for (size_t i = by; i < ysize + by; i++) {
    for (size_t j = bx; j < xsize + bx; j++) {
      float sum = 0;
      for (int k = -1; k < 2; k++) {
        for (int l = -1; l < 2; l++) {
          const size_t cy = i + k, cx = j + l;  // wraps at 0 - 1; the `<` test then fails
          if (x + cx < c.stride && y + cy < c.ysize) {
            sum = fmaf(Px(c, 1, x + cx, y + cy), mask[k + 1][l + 1], sum);
          }
        }
      }
      out[(i - by) * xsize + (j - bx)] = sum;
    }
  }
}

// proposals/homogeneity-partitioning.diff:83-105
float CalculateSumModifiedLaplacian(size_t x, size_t y, size_t xsize, size_t ysize, size_t bx, size_t by,
                                    const HomogConfig& c) {
  float sum = 0;
  for (size_t i = by; i < ysize + by; i++) {
    for (size_t j = bx; j < xsize + bx; j++) {
      if (x + j + 1 >= c.stride || y + i + 1 >= c.ysize) continue;
      if (x + j == 0 || y + i == 0) continue;  // defined behaviour for the reference's UB
      const float p = Px(c, 1, x + j, y + i);
      const float pl = Px(c, 1, x + j - 1, y + i);
      const float pr = Px(c, 1, x + j + 1, y + i);
      const float pu = Px(c, 1, x + j, y + i - 1);
      const float pd = Px(c, 1, x + j, y + i + 1);
      sum += fabsf(2 * p - pl - pr) + fabsf(2 * p - pu - pd);
    }
  }
  return sum;
}

// proposals/homogeneity-partitioning.diff:107-151
float CalculateColorfulness(size_t x, size_t y, size_t xsize, size_t ysize, size_t bx, size_t by,
                            const HomogConfig& c) {
  const float n = (float)(xsize * ysize);
  float mean_x = 0;
  for (size_t i = by; i < ysize + by; i++) for (size_t j = bx; j < xsize + bx; j++) mean_x += Px(c, 0, x + j, y + i);
  mean_x /= n;
  float mean_b = 0;
  for (size_t i = by; i < ysize + by; i++) for (size_t j = bx; j < xsize + bx; j++) mean_b += Px(c, 2, x + j, y + i);
  mean_b /= n;
  float var_x = 0;
  for (size_t i = by; i < ysize + by; i++) for (size_t j = bx; j < xsize + bx; j++) {
    const float diff = Px(c, 0, x + j, y + i) - mean_x;
    var_x = fmaf(diff, diff, var_x);
  }
  var_x /= n;
  float var_b = 0;
  for (size_t i = by; i < ysize + by; i++) for (size_t j = bx; j < xsize + bx; j++) {
    const float diff = Px(c, 2, x + j, y + i) - mean_b;
    var_b = fmaf(diff, diff, var_b);
  }
  var_b /= n;
  const float s1 = var_x + var_b;
  const float s2 = fmaf(mean_x, mean_x, mean_b * mean_b);
  return (float)(sqrt((double)s1) + 0.3 * sqrt((double)s2));
}

// proposals/homogeneity-partitioning.diff:153-181
float CalculateHomogeneity(size_t x, size_t y, size_t xsize, size_t ysize, size_t bx, size_t by, float d,
                           const HomogConfig& c) {
  float lap[64];
  CalculateLaplacianFilter(x, y, xsize, ysize, bx, by, c, lap);
  float thr = 0.25f;
  if (d > 10.0f) thr = 0.40f; else if (d <= 2.0f) thr = 0.15f;
  const size_t crossings = CalculateNumZeroCrossings(xsize, ysize, thr, lap);
  const float sml = CalculateSumModifiedLaplacian(x, y, xsize, ysize, bx, by, c);
  const float col = CalculateColorfulness(x, y, xsize, ysize, bx, by, c);
  return ((float)crossings + sml) + col;
}

static inline float MaxF(float a, float b) { return (a < b) ? b : a; }  // std::max
static inline float MinF(float a, float b) { return (b < a) ? b : a; }  // std::min

// proposals/homogeneity-partitioning.diff:183-211
void CalculateHomogeneitySimilarityIndices(size_t x, size_t y, float d, const HomogConfig& c,
                                           float* r_h, float* r_v, float* r_d) {
  const float h1 = CalculateHomogeneity(x, y, 8, 4, 0, 0, d, c);
  const float h2 = CalculateHomogeneity(x, y, 8, 4, 0, 4, d, c);
  const float v1 = CalculateHomogeneity(x, y, 4, 8, 0, 0, d, c);
  const float v2 = CalculateHomogeneity(x, y, 4, 8, 4, 0, d, c);
  // only the second term of each diagonal sum is halved (operator precedence, diff :200-203)
  const float d1 = CalculateHomogeneity(x, y, 4, 4, 0, 0, d, c) + CalculateHomogeneity(x, y, 4, 4, 4, 4, d, c) / 2;
  const float d2 = CalculateHomogeneity(x, y, 4, 4, 0, 4, d, c) + CalculateHomogeneity(x, y, 4, 4, 4, 0, d, c) / 2;
  *r_h = MaxF(h1, h2) / MinF(h1, h2);
  *r_v = MaxF(v1, v2) / MinF(v1, v2);
  *r_d = MaxF(d1, d2) / MinF(d1, d2);
}

// proposals/homogeneity-partitioning.diff:213-235 (decision only; indices passed in)
uint8_t HomogeneityPartition(float r_h, float r_v, float r_d, float d) {
  float thr = 1.60f;
  if (d > 10.0f) thr = 1.80f; else if (d <= 3.0f) thr = 1.50f;
  if (r_d > thr) return DCT4X4;
  if (r_h > r_v && r_h > thr) return DCT8X4;
  if (r_v > r_h && r_v > thr) return DCT4X8;
  return DCT;
}

}  // namespace jxo
