// ORACLE (test infrastructure) — stage U5c: quantiser set-up and per-block
// transform + quantisation (libjxl lib/jxl/quantizer.cc, enc_group.cc:
// ComputeCoefficients / QuantizeBlockAC / AdjustQuantBlockAC / QuantizeRoundtripYBlockAC,
// enc_modular.cc: AddVarDCTDC) [UPSTREAM; SURVEY.md section 8a row U5].  parity unpinned.
#include "jxo_enc.h"

namespace jxo {

static inline float Clamp1(float v, float lo, float hi) { return v < lo ? lo : (v > hi ? hi : v); }

// Quantizer::SetQuantField / ComputeGlobalScaleAndQuant: global_scale from
// median - median-absolute-deviation of the float quant field.
void ComputeGlobalScale(const float* qf, size_t n, float quant_dc, QuantState* q) {
  std::vector<float> data(qf, qf + n);
  std::nth_element(data.begin(), data.begin() + n / 2, data.end());
  const float median = data[n / 2];
  std::vector<float> dev(n);
  for (size_t i = 0; i < n; ++i) dev[i] = fabsf(qf[i] - median);
  std::nth_element(dev.begin(), dev.begin() + n / 2, dev.end());
  const float mad = dev[n / 2];
  const float kQuantFieldTarget = 5.0f;
  float scale = 65536.0f * (median - mad) / kQuantFieldTarget;
  if (!(scale >= 1.0f)) scale = 1.0f;
  if (scale > 32768.0f) scale = 32768.0f;
  int gs = (int)scale;
  const int scaled_quant_dc = (int)(quant_dc * 4096.0f * 1.6f);
  if (gs > scaled_quant_dc) { gs = scaled_quant_dc; if (gs <= 0) gs = 1; }
  q->global_scale = gs;
  q->inv_global_scale = 65536.0f / (float)gs;
  float fval = quant_dc * q->inv_global_scale + 0.5f;
  if (fval > 65536.0f) fval = 65536.0f;
  q->quant_dc = (int)fval;
  if (q->quant_dc < 1) q->quant_dc = 1;
  q->scale = (float)gs * (1.0f / 65536.0f);
  q->median = median; q->mad = mad;
}

void SetRawQuantField(const float* qf, size_t n, const QuantState& q, int32_t* raw) {
  for (size_t i = 0; i < n; ++i) {
    int v = (int)(qf[i] * q.inv_global_scale + 0.5f);
    raw[i] = v < 1 ? 1 : (v > 256 ? 256 : v);
  }
}

// libjxl AdjustQuantBias
static inline float AdjustQuantBias(int c, int32_t q) {
  static const float kBias[4] = {1.0f - 0.05465007330715401f, 1.0f - 0.07005449891748593f,
                                 1.0f - 0.049935103337343655f, 0.145f};
  if (q == 0) return 0.0f;
  if (q == 1) return kBias[c];
  if (q == -1) return -kBias[c];
  const float fq = (float)q;
  return fq - kBias[3] / fq;
}

// QuantizeBlockAC: xs >= ys are the covered blocks of the coefficient layout.
static void QuantizeBlockAC(const float* qm, float qac_mul, int c, int xs, int ys, float* thr,
                            const float* in, int32_t* out) {
  if (c != 1 && xs * ys >= 4) {
    for (int i = 0; i < 4; ++i) {
      thr[i] -= 0.00744f * (float)(xs * ys);
      if (thr[i] < 0.5f) thr[i] = 0.5f;
    }
  }
  const int W = xs * 8, H = ys * 8;
  for (int y = 0; y < H; ++y) {
    const int yfix = (y >= H / 2) ? 2 : 0;
    for (int x = 0; x < W; ++x) {
      const float t = thr[yfix + (x >= W / 2 ? 1 : 0)];
      const float q = qm[y * W + x] * qac_mul;
      const float val = q * in[y * W + x];
      int32_t v = (fabsf(val) >= t) ? (int32_t)rintf(val) : 0;
      if (x < xs && y < ys) v = 0;  // LLF is carried by the DC image
      if (v > 32767) v = 32767;     // coefficients are stored as int16
      if (v < -32767) v = -32767;
      out[y * W + x] = v;
    }
  }
}

static void AdjustQuantBlockAC(const float* qm, float scale, int c, float qm_mul, int strategy,
                               int xs, int ys, float* thr, const float* in, int32_t* quant) {
  const uint32_t kPartial = (1u << IDENTITY) | (1u << DCT2X2) | (1u << DCT4X4) | (1u << DCT4X8) |
                            (1u << DCT8X4) | (1u << AFV0) | (1u << AFV1) | (1u << AFV2) | (1u << AFV3);
  if ((1u << strategy) & kPartial) return;
  const float qac = scale * (float)(*quant);
  if (xs > 1 || ys > 1) {
    for (int i = 0; i < 4; ++i) {
      thr[i] -= Clamp1(0.003f * (float)(xs * ys), 0.f, (c > 0 ? 0.08f : 0.12f));
      if (thr[i] < 0.54f) thr[i] = 0.54f;
    }
  }
  // Block-wide float sums are defined as per-lane partial sums followed by a halving tree over the lanes
  // (stride L/2 .. 1), the association a warp-shuffle butterfly produces (DESIGN.md "Numerics").  Lane =
  // horizontal frequency: the storage row of square / tall transforms (sequential over x), the storage
  // column of wide ones (sequential over y) — the index a lane of the CUDA path owns after its column pass.
  const int W = xs * 8, H = ys * 8;
  const bool wide = kCoveredX[strategy] > kCoveredY[strategy];
  const int L = wide ? W : H;
  std::vector<float> r_hf(L, 0.0f), r_err(L, 0.0f), r_vals(L, 0.0f), r_nz[4];
  for (int i = 0; i < 4; ++i) r_nz[i].assign(L, 0.0f);
  float hfMaxErr[4] = {0, 0, 0, 0};
  for (int y = 0; y < H; ++y) {
    for (int x = 0; x < W; ++x) {
      if (x < xs && y < ys) continue;
      const int pos = y * W + x;
      const int lane = wide ? x : y;
      const int hfix = (y >= H / 2 ? 2 : 0) + (x >= W / 2 ? 1 : 0);
      const float val = in[pos] * (qm[pos] * qac * qm_mul);
      const float v = (fabsf(val) < thr[hfix]) ? 0.0f : rintf(val);
      const float err = fabsf(val - v);
      r_err[lane] += err;
      r_vals[lane] += fabsf(v);
      if (c == 1 && v == 0.0f) { if (hfMaxErr[hfix] < err) hfMaxErr[hfix] = err; }
      if (v != 0.0f) {
        r_nz[hfix][lane] += fabsf(v);
        const bool in_corner = y >= 7 * ys && x >= 7 * xs;
        const bool on_border = y == H - 1 || x == W - 1;
        const bool in_larger_corner = x >= 4 * xs && y >= 4 * ys;
        if (in_corner || (on_border && in_larger_corner)) r_hf[lane] += fabsf(val);
      }
    }
  }
  auto tree = [L](std::vector<float>& p) { for (int st = L / 2; st >= 1; st /= 2) for (int y = 0; y < st; ++y) p[y] = p[y] + p[y + st]; return p[0]; };
  const float sum_hf_rc = tree(r_hf), sum_err = tree(r_err), sum_vals = tree(r_vals);
  float hfNZ[4];
  for (int i = 0; i < 4; ++i) hfNZ[i] = tree(r_nz[i]);
  if (c == 1 && sum_vals * 8 < (float)(xs * ys)) {
    const double kLimit = 0.46, kMul = 0.9999;
    const int32_t orig = *quant;
    int32_t nq = *quant;
    for (int i = 1; i < 4; ++i) {
      if (hfNZ[i] == 0.0f && (double)hfMaxErr[i] > kLimit) { nq = orig + 1; break; }
    }
    *quant = nq;
    if (hfNZ[3] == 0.0f && (double)hfMaxErr[3] > kLimit) {
      thr[3] = (float)(kMul * (double)hfMaxErr[3] * (double)nq / (double)orig);
    } else if ((hfNZ[1] == 0.0f && (double)hfMaxErr[1] > kLimit) || (hfNZ[2] == 0.0f && (double)hfMaxErr[2] > kLimit)) {
      const float m = hfMaxErr[1] > hfMaxErr[2] ? hfMaxErr[1] : hfMaxErr[2];
      thr[1] = (float)(kMul * (double)m * (double)nq / (double)orig);
      thr[2] = thr[1];
    } else if (hfNZ[0] == 0.0f && (double)hfMaxErr[0] > kLimit) {
      thr[0] = (float)(kMul * (double)hfMaxErr[0] * (double)nq / (double)orig);
    }
  }
  {
    const float all = hfNZ[0] + hfNZ[1] + hfNZ[2] + hfNZ[3] + 1;
    const float mul[3] = {70, 30, 60};
    if (mul[c] * sum_hf_rc >= all) {
      *quant = (int32_t)((float)(*quant) + mul[c] * sum_hf_rc / all);
      if (*quant >= 256) *quant = 255;
    }
  }
  if (strategy == DCT) {
    if (hfNZ[0] + hfNZ[1] + hfNZ[2] + hfNZ[3] < 11) {
      *quant += 1;
      if (*quant >= 256) *quant = 255;
    }
  }
  {
    static const double kMul1[4][3] = {{0.22080615753848404, 0.45797479824262011, 0.29859235095977965},
                                       {0.70109486510286834, 0.16185281305512639, 0.14387691730035473},
                                       {0.114985964456218638, 0.44656840441027695, 0.10587658215149048},
                                       {0.46849665264409396, 0.41239077937781954, 0.088667407767185444}};
    static const double kMul2[4][3] = {{0.27450281941822197, 1.1255766549984996, 0.98950459134128388},
                                       {0.4652168675598285, 0.40945807983455818, 0.36581899811751367},
                                       {0.28034972424715715, 0.9182653201929738, 1.5581531543057416},
                                       {0.26873118114033728, 0.68863712390392484, 1.2082185408666786}};
    const double kQuantNormalizer = 2.2942708343284721;
    const double se = (double)sum_err * kQuantNormalizer;
    const double sv = (double)sum_vals * kQuantNormalizer;
    if (strategy >= DCT16X16) {
      int ix = 3;
      if (strategy == DCT32X16 || strategy == DCT16X32) ix = 1;
      else if (strategy == DCT16X16) ix = 0;
      else if (strategy == DCT32X32) ix = 2;
      const double lim = kMul1[ix][c] * (double)(xs * ys * 64) + kMul2[ix][c] * sv;
      int step = (int)(se / lim);
      if (step >= 2) step = 2;
      if (step < 0) step = 0;
      if (se > lim) {
        *quant += step;
        if (*quant >= 256) *quant = 255;
      }
    }
  }
}

void ComputeCoefficientsBlock(const EncTables& T, const QuantState& q, int strategy,
                              const float* px[3], int ps, float x_factor, float b_factor,
                              int32_t* quant_io, float* dc_out[3], int dc_stride, int32_t* out[3]) {
  const int cx = kCoveredX[strategy], cy = kCoveredY[strategy];
  const int xs = cx > cy ? cx : cy, ys = cx > cy ? cy : cx;
  const int size = cx * cy * 64;
  std::vector<float> coef((size_t)3 * size);
  for (int c = 0; c < 3; ++c) {
    TransformFromPixels(strategy, px[c], ps, &coef[(size_t)c * size]);
    DcFromLowestFrequencies(strategy, &coef[(size_t)c * size], dc_out[c], dc_stride);
  }
  const float* qm = T.weights[kQuantKind[strategy]].data();
  const float* dq = T.dequant[kQuantKind[strategy]].data();
  float thres_y[4] = {0.58f, 0.64f, 0.64f, 0.64f};
  int32_t quant = *quant_io;
  if (q.adjust_quant) {
    int32_t max_quant = 0;
    const int32_t orig = quant;
    const float mulc[3] = {q.x_qm_mul, 1.0f, q.b_qm_mul};
    static const int order[3] = {1, 0, 2};
    for (int k = 0; k < 3; ++k) {
      const int c = order[k];
      float thr[4] = {0.58f, 0.64f, 0.64f, 0.64f};
      int32_t qq = orig;
      AdjustQuantBlockAC(qm + (size_t)c * size, q.scale, c, mulc[c], strategy, xs, ys, thr, &coef[(size_t)c * size], &qq);
      if (c == 1) for (int i = 0; i < 4; ++i) thres_y[i] = thr[i];
      if (qq > max_quant) max_quant = qq;
    }
    quant = max_quant;
  } else {
    thres_y[0] = 0.56f; thres_y[1] = thres_y[2] = thres_y[3] = 0.62f;
  }
  *quant_io = quant;
  const float qac = q.scale * (float)quant;
  QuantizeBlockAC(qm + (size_t)size, qac * 1.0f, 1, xs, ys, thres_y, &coef[(size_t)size], out[1]);
  // roundtrip Y, then remove the chroma-from-luma prediction from X and B
  const float inv_qac = q.inv_global_scale / (float)quant;
  for (int k = 0; k < size; ++k) {
    const float yrt = (AdjustQuantBias(1, out[1][k]) * dq[(size_t)size + k]) * inv_qac;
    coef[k] = fmaf(-x_factor, yrt, coef[k]);
    coef[(size_t)2 * size + k] = fmaf(-b_factor, yrt, coef[(size_t)2 * size + k]);
  }
  {
    float thr[4] = {0.58f, 0.64f, 0.64f, 0.64f};
    QuantizeBlockAC(qm, qac * q.x_qm_mul, 0, xs, ys, thr, &coef[0], out[0]);
  }
  {
    float thr[4] = {0.58f, 0.64f, 0.64f, 0.64f};
    QuantizeBlockAC(qm + (size_t)2 * size, qac * q.b_qm_mul, 2, xs, ys, thr, &coef[(size_t)2 * size], out[2]);
  }
}

// AddVarDCTDC quantisation: Y first, X/B after removing cfl * dequantised Y.
void QuantizeDc(const QuantState& q, const float* dc[3], size_t n, int32_t* out[3]) {
  static const float kInvDcQuant[3] = {4096.0f, 512.0f, 256.0f};
  static const float kDcQuant[3] = {1.0f / 4096.0f, 1.0f / 512.0f, 1.0f / 256.0f};
  const float gsq = q.scale * (float)q.quant_dc;                       // global_scale_float * quant_dc
  const float inv_quant_dc = q.inv_global_scale / (float)q.quant_dc;
  float inv_factor[3], cfl[3] = {0.0f, 0.0f, 1.0f};
  for (int c = 0; c < 3; ++c) inv_factor[c] = kInvDcQuant[c] * gsq;
  const float y_factor = inv_quant_dc * kDcQuant[1];
  for (size_t i = 0; i < n; ++i) {
    const float qy = roundf(dc[1][i] * inv_factor[1]);
    out[1][i] = (int32_t)qy;
    out[0][i] = (int32_t)roundf((dc[0][i] - qy * (y_factor * cfl[0])) * inv_factor[0]);
    out[2][i] = (int32_t)roundf((dc[2][i] - qy * (y_factor * cfl[2])) * inv_factor[2]);
  }
}

}  // namespace jxo
