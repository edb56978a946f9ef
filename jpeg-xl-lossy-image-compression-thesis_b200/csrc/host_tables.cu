// jxlb200 — host-side constant tables uploaded once per context: sRGB EOTF table, default
// quantisation weights per table kind, natural coefficient orders, strategy geometry.
// (libjxl quant_weights.cc DequantMatrices::Library / GetQuantWeights, ac_strategy.cc
// ComputeNaturalCoeffOrder, cms/transfer_functions-inl.h [UPSTREAM]; SURVEY.md Appendix U.)
// Compiled with -Xcompiler -ffp-contract=off: fused multiply-adds only where fmaf() is written.
#include "host_tables.h"
#include <cmath>
#include <cstring>
#include <algorithm>

namespace jxlb {

const uint8_t kCoveredX[27] = {1, 1, 1, 1, 2, 4, 1, 2, 1, 4, 2, 4, 1, 1, 1, 1, 1, 1, 8, 4, 8, 16, 8, 16, 32, 16, 32};
const uint8_t kCoveredY[27] = {1, 1, 1, 1, 2, 4, 2, 1, 4, 1, 4, 2, 1, 1, 1, 1, 1, 1, 8, 8, 4, 16, 16, 8, 32, 32, 16};
const uint8_t kStrategyOrder[27] = {0, 1, 1, 1, 2, 3, 4, 4, 5, 5, 6, 6, 1, 1, 1, 1, 1, 1, 7, 8, 8, 9, 10, 10, 11, 12, 12};
const uint8_t kQuantKind[27] = {0, 1, 2, 3, 4, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 10, 10, 11, 12, 12, 13, 14, 14, 15, 16, 16};

static inline uint32_t fbits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float ffrom(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

void host_srgb_lut(float lut[256]) {
  static const float p[5] = {2.200248328e-04f, 1.043637593e-02f, 1.624820318e-01f, 7.961564959e-01f, 8.210152774e-01f};
  static const float q[5] = {2.631846970e-01f, 1.076976492e+00f, 4.987528350e-01f, -5.512498495e-02f, 6.521209011e-03f};
  for (int i = 0; i < 256; ++i) {
    const float x = (float)i / 255.0f;
    if (x > 0.04045f) {
      float yp = p[4], yq = q[4];
      for (int k = 3; k >= 0; --k) { yp = fmaf(yp, x, p[k]); yq = fmaf(yq, x, q[k]); }
      lut[i] = yp / yq;
    } else {
      lut[i] = x * (1.0f / 12.92f);
    }
  }
}

// tab[0..8]: inverse opsin matrix (computed in double from the forward matrix, rounded to float);
// tab[9 + n]: sRGB code boundary in linear light, EOTF((n + 0.5) / 255) by the encoder's rational polynomial —
// code n is chosen when boundary[n-1] <= linear < boundary[n] (oracle/jxo_recon.cc SrgbBoundaries / InverseOpsin)
void host_recon_tables(float tab[264]) {
  const double M[3][3] = {{0.30, 0.622, 0.078}, {0.23, 0.692, 0.078}, {0.24342268924547819, 0.20476744424496821, 0.55180986650955360}};
  const double a = M[0][0], b = M[0][1], c = M[0][2], d = M[1][0], e = M[1][1], g = M[1][2], h = M[2][0], i = M[2][1], j = M[2][2];
  const double det = a * (e * j - g * i) - b * (d * j - g * h) + c * (d * i - e * h);
  const double v[9] = {(e * j - g * i) / det, (c * i - b * j) / det, (b * g - c * e) / det,
                       (g * h - d * j) / det, (a * j - c * h) / det, (c * d - a * g) / det,
                       (d * i - e * h) / det, (b * h - a * i) / det, (a * e - b * d) / det};
  for (int k = 0; k < 9; ++k) tab[k] = (float)v[k];
  static const float p[5] = {2.200248328e-04f, 1.043637593e-02f, 1.624820318e-01f, 7.961564959e-01f, 8.210152774e-01f};
  static const float q[5] = {2.631846970e-01f, 1.076976492e+00f, 4.987528350e-01f, -5.512498495e-02f, 6.521209011e-03f};
  for (int n = 0; n < 255; ++n) {
    const float x = ((float)n + 0.5f) / 255.0f;
    if (x > 0.04045f) {
      float yp = p[4], yq = q[4];
      for (int k = 3; k >= 0; --k) { yp = fmaf(yp, x, p[k]); yq = fmaf(yq, x, q[k]); }
      tab[9 + n] = yp / yq;
    } else {
      tab[9 + n] = x * (1.0f / 12.92f);
    }
  }
}

static float h_log2(float x) {
  const int32_t xb = (int32_t)fbits(x);
  const int32_t es = (xb - 0x3f2aaaab) >> 23;
  const float m = ffrom((uint32_t)(xb - (es << 23))) - 1.0f;
  const float yp = fmaf(fmaf(7.4245873327820566E-01f, m, 1.4287160470083755E+00f), m, -1.8503833400518310E-06f);
  const float yq = fmaf(fmaf(1.7409343003366853E-01f, m, 1.0096718572241148E+00f), m, 9.9032814277590719E-01f);
  return yp / yq + (float)es;
}
static float h_pow2(float x) {
  const float fl = floorf(x);
  const float e = ffrom((uint32_t)(((int32_t)fl + 127) << 23));
  const float fr = x - fl;
  float num = fr + 1.01749063e+01f;
  num = fmaf(num, fr, 4.88687798e+01f);
  num = fmaf(num, fr, 9.85506591e+01f);
  num = num * e;
  float den = fmaf(fr, 2.10242958e-01f, -2.22328856e-02f);
  den = fmaf(den, fr, -1.94414990e+01f);
  den = fmaf(den, fr, 9.85506633e+01f);
  return num / den;
}

struct Bands { int n; float b[3][8]; };
static const Bands kB8 = {6, {{3150.0f, 0.0f, -0.4f, -0.4f, -0.4f, -2.0f}, {560.0f, 0.0f, -0.3f, -0.3f, -0.3f, -0.3f}, {512.0f, -2.0f, -1.0f, 0.0f, -1.0f, -2.0f}}};
static const Bands kB4 = {4, {{2200.0f, 0.0f, 0.0f, 0.0f}, {392.0f, 0.0f, 0.0f, 0.0f}, {112.0f, -0.25f, -0.25f, -0.5f}}};
static const Bands kB16 = {7, {{8996.8725711814115328f, -1.3000777393353804f, -0.49424529824571225f, -0.439093774457103443f, -0.6350101832695744f, -0.90177264050827612f, -1.6162099239887414f},
                               {3191.48366296844234752f, -0.67424582104194355f, -0.80745813428471001f, -0.44925837484843441f, -0.35865440981033403f, -0.31322389111877305f, -0.37615025315725483f},
                               {1157.50408145487200256f, -2.0531423165804414f, -1.4f, -0.50687130033378396f, -0.42708730624733904f, -1.4856834539296244f, -4.9209142884401604f}}};
static const Bands kB32 = {8, {{15718.40830982518931456f, -1.025f, -0.98f, -0.9012f, -0.4f, -0.48819395464f, -0.421064f, -0.27f},
                               {7305.7636810695983104f, -0.8041958212306401f, -0.7633036457487539f, -0.55660379990111464f, -0.49785304658857626f, -0.43699592683512467f, -0.40180866526242109f, -0.27321683125358037f},
                               {3803.53173721215041536f, -3.060733579805728f, -2.0413270132490346f, -2.0235650159727417f, -0.5495389509954993f, -0.4f, -0.4f, -0.3f}}};
static const Bands kB16x8 = {7, {{7240.7734393502f, -0.7f, -0.7f, -0.2f, -0.2f, -0.2f, -0.5f}, {1448.15468787004f, -0.5f, -0.5f, -0.5f, -0.2f, -0.2f, -0.2f}, {506.854140754517f, -1.4f, -0.2f, -0.5f, -0.5f, -1.5f, -3.6f}}};
static const Bands kB32x8 = {8, {{16283.2494710648897f, -1.7812845336559429f, -1.6309059012653515f, -1.0382179034313539f, -0.85f, -0.7f, -0.9f, -1.2360638576849587f},
                                 {5089.15750884921511936f, -0.320049391452786891f, -0.35362849922161446f, -0.30340000000000003f, -0.61f, -0.5f, -0.5f, -0.6f},
                                 {3397.77603275308720128f, -0.321327362693153371f, -0.34507619223117997f, -0.70340000000000003f, -0.9f, -1.0f, -1.0f, -1.1754605576265209f}}};
static const Bands kB32x16 = {8, {{13844.97076442300573f, -0.97113799999999995f, -0.658f, -0.42026f, -0.22712f, -0.2206f, -0.226f, -0.6f},
                                  {4798.964084220744293f, -0.61125308982767057f, -0.83770786552491361f, -0.79014862079498627f, -0.2692727459704829f, -0.38272769465388551f, -0.22924222653091453f, -0.20719098826199578f},
                                  {1807.236946760964614f, -1.2f, -1.2f, -0.7f, -0.7f, -0.7f, -0.4f, -0.5f}}};
static const Bands kB4x8 = {4, {{2198.050556016380522f, -0.96269623020744692f, -0.76194253026666783f, -0.6551140670773547f},
                                {764.3655248643528689f, -0.92630200888366945f, -0.9675229603596517f, -0.27845290869168118f},
                                {527.107573587542228f, -1.4594385811273854f, -1.450082094097871593f, -1.5843722511996204f}}};

static const Bands kB64 = {8, {{0.9f * 26629.073922049845f, -1.025f, -0.78f, -0.65012f, -0.19041574084286472f, -0.20819395464f, -0.421064f, -0.32733845535848671f},
                               {0.9f * 9311.3238710010046f, -0.3041958212306401f, -0.3633036457487539f, -0.35660379990111464f, -0.3443074455424403f, -0.33699592683512467f, -0.30180866526242109f, -0.27321683125358037f},
                               {0.9f * 4992.2486445538634f, -1.2f, -1.2f, -0.8f, -0.7f, -0.7f, -0.4f, -0.5f}}};
static const Bands kB32x64 = {8, {{0.65f * 23629.073922049845f, -1.025f, -0.78f, -0.65012f, -0.19041574084286472f, -0.20819395464f, -0.421064f, -0.32733845535848671f},
                                  {0.65f * 8611.3238710010046f, -0.3041958212306401f, -0.3633036457487539f, -0.35660379990111464f, -0.3443074455424403f, -0.33699592683512467f, -0.30180866526242109f, -0.27321683125358037f},
                                  {0.65f * 4492.2486445538634f, -1.2f, -1.2f, -0.8f, -0.7f, -0.7f, -0.4f, -0.5f}}};
// IDENTITY / DCT2X2 parameter sets (libjxl QuantEncoding::Identity / DCT2 library values)
static const float kIdW[3][3] = {{280.0f, 3160.0f, 3160.0f}, {60.0f, 864.0f, 864.0f}, {18.0f, 200.0f, 200.0f}};
static const float kDct2W[3][6] = {{3840.0f, 2560.0f, 1280.0f, 640.0f, 480.0f, 300.0f}, {960.0f, 640.0f, 320.0f, 180.0f, 140.0f, 120.0f},
                                   {640.0f, 320.0f, 128.0f, 64.0f, 32.0f, 16.0f}};

static void band_weights(int rows, int cols, const Bands& bp, float* out) {
  for (int c = 0; c < 3; ++c) {
    float bands[8];
    bands[0] = bp.b[c][0];
    for (int i = 1; i < bp.n; ++i) {
      const float v = bp.b[c][i];
      bands[i] = bands[i - 1] * (v > 0.0f ? 1.0f + v : 1.0f / (1.0f - v));
    }
    const float scale = (float)(bp.n - 1) / (1.41421356237309504880f + 1e-6f);
    const float rcpcol = scale / (float)(cols - 1), rcprow = scale / (float)(rows - 1);
    for (int y = 0; y < rows; ++y) {
      const float dy = (float)y * rcprow, dy2 = dy * dy;
      for (int x = 0; x < cols; ++x) {
        const float dx = (float)x * rcpcol;
        const float dist = sqrtf(fmaf(dx, dx, dy2));
        int idx = (int)dist;
        if (idx > bp.n - 2) idx = bp.n - 2;
        const float frac = dist - (float)idx;
        const float a = bands[idx], b = bands[idx + 1];
        out[(size_t)c * rows * cols + (size_t)y * cols + x] = a * h_pow2(h_log2(b / a) * frac);
      }
    }
  }
}

int host_quant_weights(int kind, std::vector<float>* w) {
  int rows = 8, cols = 8;
  const Bands* bp = nullptr;
  switch (kind) {
    case 0: bp = &kB8; break;
    case 4: bp = &kB16; rows = cols = 16; break;
    case 5: bp = &kB32; rows = cols = 32; break;
    case 6: bp = &kB16x8; rows = 8; cols = 16; break;
    case 7: bp = &kB32x8; rows = 8; cols = 32; break;
    case 8: bp = &kB32x16; rows = 16; cols = 32; break;
    case 11: bp = &kB64; rows = cols = 64; break;
    case 12: bp = &kB32x64; rows = 32; cols = 64; break;
    case 1: {
      w->assign(192, 0.0f);
      for (int c = 0; c < 3; ++c) {
        for (int i = 0; i < 64; ++i) (*w)[c * 64 + i] = kIdW[c][0];
        (*w)[c * 64 + 1] = kIdW[c][1]; (*w)[c * 64 + 8] = kIdW[c][1]; (*w)[c * 64 + 9] = kIdW[c][2];
      }
      return 64;
    }
    case 2: {
      w->assign(192, 0.0f);
      for (int c = 0; c < 3; ++c) {
        float* o = &(*w)[c * 64];
        o[0] = 2989.0f;   // the DC position is never quantised with this table
        o[1] = o[8] = kDct2W[c][0];
        o[9] = kDct2W[c][1];
        for (int y = 0; y < 2; ++y) for (int x = 2; x < 4; ++x) { o[y * 8 + x] = kDct2W[c][2]; o[x * 8 + y] = kDct2W[c][2]; }
        for (int y = 2; y < 4; ++y) for (int x = 2; x < 4; ++x) o[y * 8 + x] = kDct2W[c][3];
        for (int y = 0; y < 4; ++y) for (int x = 4; x < 8; ++x) { o[y * 8 + x] = kDct2W[c][4]; o[x * 8 + y] = kDct2W[c][4]; }
        for (int y = 4; y < 8; ++y) for (int x = 4; x < 8; ++x) o[y * 8 + x] = kDct2W[c][5];
      }
      return 64;
    }
    case 3: {
      float w4[48];
      band_weights(4, 4, kB4, w4);
      w->assign(192, 0.0f);
      for (int c = 0; c < 3; ++c) for (int y = 0; y < 8; ++y) for (int x = 0; x < 8; ++x)
        (*w)[c * 64 + y * 8 + x] = w4[c * 16 + (y / 2) * 4 + (x / 2)];
      return 64;
    }
    case 9: {
      float w48[96];
      band_weights(4, 8, kB4x8, w48);
      w->assign(192, 0.0f);
      for (int c = 0; c < 3; ++c) for (int y = 0; y < 8; ++y) for (int x = 0; x < 8; ++x)
        (*w)[c * 64 + y * 8 + x] = w48[c * 32 + (y / 2) * 8 + x];
      return 64;
    }
    default: w->clear(); return 0;
  }
  w->assign((size_t)3 * rows * cols, 0.0f);
  band_weights(rows, cols, *bp, w->data());
  return rows * cols;
}

static int ceil_log2(uint32_t v) { return v <= 1 ? 0 : 32 - __builtin_clz(v - 1); }

void host_natural_order(int strategy, std::vector<uint16_t>* order) {
  size_t cx = kCoveredX[strategy], cy = kCoveredY[strategy];
  if (cy > cx) std::swap(cx, cy);
  order->assign(cx * cy * 64, 0);
  const size_t xs = cx / cy, xsm = xs - 1, xss = (size_t)ceil_log2((uint32_t)xs);
  size_t cur = cx * cy;
  const size_t n = cx * 8;
  for (size_t i = 0; i < n; i++) for (size_t j = 0; j <= i; j++) {
    size_t x = j, y = i - j;
    if (i % 2) std::swap(x, y);
    if ((y & xsm) != 0) continue;
    y >>= xss;
    const size_t val = (x < cx && y < cy) ? y * cx + x : cur++;
    (*order)[val] = (uint16_t)(y * cx * 8 + x);
  }
  for (size_t ip = n - 1; ip > 0; ip--) {
    const size_t i = ip - 1;
    for (size_t j = 0; j <= i; j++) {
      size_t x = n - 1 - (i - j), y = n - 1 - j;
      if (i % 2) std::swap(x, y);
      if ((y & xsm) != 0) continue;
      y >>= xss;
      (*order)[cur++] = (uint16_t)(y * cx * 8 + x);
    }
  }
}

float host_initial_quant_dc(float d) {
  const float kDcMul = 0.3f, kDcQuantPow = 0.83f, kDcQuant = 1.095924047623553f;
  const float a = kDcMul * powf((1.0f / kDcMul) * d, kDcQuantPow);
  const float m = d < a ? d : a;
  const float t = 0.5f * d > m ? 0.5f * d : m;
  const float r = kDcQuant / t;
  return r < 50.0f ? r : 50.0f;
}

}  // namespace jxlb
