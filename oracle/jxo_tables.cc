// ORACLE (test infrastructure) — tables: quant matrices per kind, natural coefficient
// orders (libjxl AcStrategy::ComputeNaturalCoeffOrder), initial DC quant
// (enc_adaptive_quantization.cc: InitialQuantDC) [UPSTREAM]. parity unpinned.
#include "jxo_enc.h"

namespace jxo {

void NaturalCoeffOrder(int strategy, std::vector<uint16_t>* order) {
  size_t cx = kCoveredX[strategy], cy = kCoveredY[strategy];
  if (cy > cx) std::swap(cx, cy);  // CoefficientLayout: cx >= cy
  order->assign(cx * cy * 64, 0);
  const size_t xs = cx / cy, xsm = xs - 1, xss = CeilLog2((uint32_t)xs);
  size_t cur = cx * cy;
  const size_t n = cx * 8;
  for (size_t i = 0; i < n; i++) {
    for (size_t j = 0; j <= i; j++) {
      size_t x = j, y = i - j;
      if (i % 2) std::swap(x, y);
      if ((y & xsm) != 0) continue;
      y >>= xss;
      size_t val;
      if (x < cx && y < cy) val = y * cx + x; else val = cur++;
      (*order)[val] = (uint16_t)(y * cx * 8 + x);
    }
  }
  for (size_t ip = n - 1; ip > 0; ip--) {
    const size_t i = ip - 1;
    for (size_t j = 0; j <= i; j++) {
      size_t x = n - 1 - (i - j), y = n - 1 - j;
      if (i % 2) std::swap(x, y);
      if ((y & xsm) != 0) continue;
      y >>= xss;
      const size_t val = cur++;
      (*order)[val] = (uint16_t)(y * cx * 8 + x);
    }
  }
}

void EncTables::Init() {
  for (int k = 0; k < 17; ++k) {
    ncoef[k] = QuantWeights(k, &weights[k]);
    dequant[k].resize(weights[k].size());
    for (size_t i = 0; i < weights[k].size(); ++i) dequant[k][i] = 1.0f / weights[k][i];
  }
  // one representative strategy per order class
  static const int rep[13] = {DCT, DCT4X4, DCT16X16, DCT32X32, DCT16X8, DCT32X8, DCT32X16, DCT64X64, 19, 21, 22, 24, 25};
  for (int o = 0; o < 13; ++o) NaturalCoeffOrder(rep[o], &order[o]);
}

float InitialQuantDC(float d) {
  const float kDcMul = 0.3f, kDcQuantPow = 0.83f, kDcQuant = 1.095924047623553f;
  const float a = kDcMul * powf((1.0f / kDcMul) * d, kDcQuantPow);
  const float m = d < a ? d : a;
  const float t = 0.5f * d > m ? 0.5f * d : m;
  const float r = kDcQuant / t;
  return r < 50.0f ? r : 50.0f;
}

}  // namespace jxo
