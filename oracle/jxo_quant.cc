// ORACLE (test infrastructure) — stage U5b: default quantisation tables.
// Restates libjxl lib/jxl/quant_weights.cc (DequantMatrices::Library, GetQuantWeights)
// and base/fast_math-inl.h (FastLog2f / FastPow2f) [UPSTREAM; SURVEY.md section 8a
// row U5, Appendix U.9].  parity unpinned: the band parameters are recalled, the
// reference tree holds no copy of them.
#include "jxo.h"

namespace jxo {

// log2(x), 2/2 rational polynomial on the mantissa (range reduced around 2/3).
float FastLog2f(float x) {
  const float p0 = -1.8503833400518310E-06f, p1 = 1.4287160470083755E+00f, p2 = 7.4245873327820566E-01f;
  const float q0 = 9.9032814277590719E-01f, q1 = 1.0096718572241148E+00f, q2 = 1.7409343003366853E-01f;
  const int32_t x_bits = (int32_t)f2u(x);
  const int32_t exp_bits = x_bits - 0x3f2aaaab;
  const int32_t exp_shifted = exp_bits >> 23;
  const float mantissa = u2f((uint32_t)(x_bits - (exp_shifted << 23)));
  const float exp_val = (float)exp_shifted;
  const float m = mantissa - 1.0f;
  float yp = fmaf(fmaf(p2, m, p1), m, p0);
  float yq = fmaf(fmaf(q2, m, q1), m, q0);
  return yp / yq + exp_val;
}

// 2^x, 3/3 rational polynomial on the fractional part.
float FastPow2f(float x) {
  const float floorx = floorf(x);
  const float exp = u2f((uint32_t)(((int32_t)floorx + 127) << 23));
  const float frac = x - floorx;
  float num = frac + 1.01749063e+01f;
  num = fmaf(num, frac, 4.88687798e+01f);
  num = fmaf(num, frac, 9.85506591e+01f);
  num = num * exp;
  float den = fmaf(frac, 2.10242958e-01f, -2.22328856e-02f);
  den = fmaf(den, frac, -1.94414990e+01f);
  den = fmaf(den, frac, 9.85506633e+01f);
  return num / den;
}

float FastPowf(float base, float e) { return FastPow2f(FastLog2f(base) * e); }

static float Mult(float v) { return v > 0.0f ? 1.0f + v : 1.0f / (1.0f - v); }

// weights for a rows x cols coefficient grid from per-channel distance bands
static void GetQuantWeights(int rows, int cols, const float bands_in[3][8], int num_bands, float* out) {
  for (int c = 0; c < 3; ++c) {
    float bands[8];
    bands[0] = bands_in[c][0];
    for (int i = 1; i < num_bands; ++i) bands[i] = bands[i - 1] * Mult(bands_in[c][i]);
    const float scale = (float)(num_bands - 1) / (1.41421356237309504880f + 1e-6f);
    const float rcpcol = scale / (float)(cols - 1);
    const float rcprow = scale / (float)(rows - 1);
    for (int y = 0; y < rows; ++y) {
      const float dy = (float)y * rcprow;
      const float dy2 = dy * dy;
      for (int x = 0; x < cols; ++x) {
        const float dx = (float)x * rcpcol;
        const float dist = sqrtf(fmaf(dx, dx, dy2));
        float w;
        if (num_bands == 1) {
          w = bands[0];
        } else {
          int idx = (int)dist;
          if (idx > num_bands - 2) idx = num_bands - 2;
          const float frac = dist - (float)idx;
          const float a = bands[idx], b = bands[idx + 1];
          w = a * FastPowf(b / a, frac);
        }
        out[(size_t)c * rows * cols + (size_t)y * cols + x] = w;
      }
    }
  }
}

struct BandParams { int num; float b[3][8]; };

// DequantMatrices::Library() parameters (recalled; DCT8 per SURVEY Appendix U.9)
static const BandParams kDct8 = {6, {{3150.0f, 0.0f, -0.4f, -0.4f, -0.4f, -2.0f},
                                     {560.0f, 0.0f, -0.3f, -0.3f, -0.3f, -0.3f},
                                     {512.0f, -2.0f, -1.0f, 0.0f, -1.0f, -2.0f}}};
static const BandParams kDct4 = {4, {{2200.0f, 0.0f, 0.0f, 0.0f},
                                     {392.0f, 0.0f, 0.0f, 0.0f},
                                     {112.0f, -0.25f, -0.25f, -0.5f}}};
static const BandParams kDct16 = {7, {{8996.8725711814115328f, -1.3000777393353804f, -0.49424529824571225f, -0.439093774457103443f, -0.6350101832695744f, -0.90177264050827612f, -1.6162099239887414f},
                                      {3191.48366296844234752f, -0.67424582104194355f, -0.80745813428471001f, -0.44925837484843441f, -0.35865440981033403f, -0.31322389111877305f, -0.37615025315725483f},
                                      {1157.50408145487200256f, -2.0531423165804414f, -1.4f, -0.50687130033378396f, -0.42708730624733904f, -1.4856834539296244f, -4.9209142884401604f}}};
static const BandParams kDct32 = {8, {{15718.40830982518931456f, -1.025f, -0.98f, -0.9012f, -0.4f, -0.48819395464f, -0.421064f, -0.27f},
                                      {7305.7636810695983104f, -0.8041958212306401f, -0.7633036457487539f, -0.55660379990111464f, -0.49785304658857626f, -0.43699592683512467f, -0.40180866526242109f, -0.27321683125358037f},
                                      {3803.53173721215041536f, -3.060733579805728f, -2.0413270132490346f, -2.0235650159727417f, -0.5495389509954993f, -0.4f, -0.4f, -0.3f}}};
static const BandParams kDct16x8 = {7, {{7240.7734393502f, -0.7f, -0.7f, -0.2f, -0.2f, -0.2f, -0.5f},
                                        {1448.15468787004f, -0.5f, -0.5f, -0.5f, -0.2f, -0.2f, -0.2f},
                                        {506.854140754517f, -1.4f, -0.2f, -0.5f, -0.5f, -1.5f, -3.6f}}};
static const BandParams kDct32x8 = {8, {{16283.2494710648897f, -1.7812845336559429f, -1.6309059012653515f, -1.0382179034313539f, -0.85f, -0.7f, -0.9f, -1.2360638576849587f},
                                        {5089.15750884921511936f, -0.320049391452786891f, -0.35362849922161446f, -0.30340000000000003f, -0.61f, -0.5f, -0.5f, -0.6f},
                                        {3397.77603275308720128f, -0.321327362693153371f, -0.34507619223117997f, -0.70340000000000003f, -0.9f, -1.0f, -1.0f, -1.1754605576265209f}}};
static const BandParams kDct32x16 = {8, {{13844.97076442300573f, -0.97113799999999995f, -0.658f, -0.42026f, -0.22712f, -0.2206f, -0.226f, -0.6f},
                                         {4798.964084220744293f, -0.61125308982767057f, -0.83770786552491361f, -0.79014862079498627f, -0.2692727459704829f, -0.38272769465388551f, -0.22924222653091453f, -0.20719098826199578f},
                                         {1807.236946760964614f, -1.2f, -1.2f, -0.7f, -0.7f, -0.7f, -0.4f, -0.5f}}};
static const BandParams kDct4x8 = {4, {{2198.050556016380522f, -0.96269623020744692f, -0.76194253026666783f, -0.6551140670773547f},
                                       {764.3655248643528689f, -0.92630200888366945f, -0.9675229603596517f, -0.27845290869168118f},
                                       {527.107573587542228f, -1.4594385811273854f, -1.450082094097871593f, -1.5843722511996204f}}};

static const BandParams kDct64 = {8, {{0.9f * 26629.073922049845f, -1.025f, -0.78f, -0.65012f, -0.19041574084286472f, -0.20819395464f, -0.421064f, -0.32733845535848671f},
                                      {0.9f * 9311.3238710010046f, -0.3041958212306401f, -0.3633036457487539f, -0.35660379990111464f, -0.3443074455424403f, -0.33699592683512467f, -0.30180866526242109f, -0.27321683125358037f},
                                      {0.9f * 4992.2486445538634f, -1.2f, -1.2f, -0.8f, -0.7f, -0.7f, -0.4f, -0.5f}}};
static const BandParams kDct32x64 = {8, {{0.65f * 23629.073922049845f, -1.025f, -0.78f, -0.65012f, -0.19041574084286472f, -0.20819395464f, -0.421064f, -0.32733845535848671f},
                                         {0.65f * 8611.3238710010046f, -0.3041958212306401f, -0.3633036457487539f, -0.35660379990111464f, -0.3443074455424403f, -0.33699592683512467f, -0.30180866526242109f, -0.27321683125358037f},
                                         {0.65f * 4492.2486445538634f, -1.2f, -1.2f, -0.8f, -0.7f, -0.7f, -0.4f, -0.5f}}};
// QuantEncoding::Identity / DCT2 parameter sets (per channel)
static const float kIdWeights[3][3] = {{280.0f, 3160.0f, 3160.0f}, {60.0f, 864.0f, 864.0f}, {18.0f, 200.0f, 200.0f}};
static const float kDct2Weights[3][6] = {{3840.0f, 2560.0f, 1280.0f, 640.0f, 480.0f, 300.0f},
                                         {960.0f, 640.0f, 320.0f, 180.0f, 140.0f, 120.0f},
                                         {640.0f, 320.0f, 128.0f, 64.0f, 32.0f, 16.0f}};

// kind: DequantMatrices::QuantTable index (1 IDENTITY, 2 DCT2X2, 11 DCT64X64, 12 DCT32X64) (0 DCT, 3 DCT4X4, 4 DCT16X16, 5 DCT32X32,
// 6 DCT8X16, 7 DCT8X32, 8 DCT16X32, 9 DCT4X8).  Weights are laid out like the
// coefficient block of the kind (long side horizontal).
int QuantWeights(int kind, std::vector<float>* w) {
  int rows = 8, cols = 8;
  const BandParams* bp = nullptr;
  switch (kind) {
    case 0: bp = &kDct8; break;
    case 4: bp = &kDct16; rows = cols = 16; break;
    case 5: bp = &kDct32; rows = cols = 32; break;
    case 6: bp = &kDct16x8; rows = 8; cols = 16; break;
    case 7: bp = &kDct32x8; rows = 8; cols = 32; break;
    case 8: bp = &kDct32x16; rows = 16; cols = 32; break;
    case 11: bp = &kDct64; rows = cols = 64; break;
    case 12: bp = &kDct32x64; rows = 32; cols = 64; break;
    case 1: {   // IDENTITY: one weight everywhere, two more for positions 1 / 8 and 9
      w->assign(3 * 64, 0.0f);
      for (int c = 0; c < 3; ++c) {
        for (int i = 0; i < 64; ++i) (*w)[c * 64 + i] = kIdWeights[c][0];
        (*w)[c * 64 + 1] = kIdWeights[c][1];
        (*w)[c * 64 + 8] = kIdWeights[c][1];
        (*w)[c * 64 + 9] = kIdWeights[c][2];
      }
      return 64;
    }
    case 2: {   // DCT2X2: one weight per Hadamard level and orientation
      w->assign(3 * 64, 0.0f);
      for (int c = 0; c < 3; ++c) {
        float* o = &(*w)[c * 64];
        o[0] = 2989.0f;   // 0xBAD: the DC position is never quantised with this table
        o[1] = o[8] = kDct2Weights[c][0];
        o[9] = kDct2Weights[c][1];
        for (int y = 0; y < 2; ++y) for (int x = 2; x < 4; ++x) { o[y * 8 + x] = kDct2Weights[c][2]; o[x * 8 + y] = kDct2Weights[c][2]; }
        for (int y = 2; y < 4; ++y) for (int x = 2; x < 4; ++x) o[y * 8 + x] = kDct2Weights[c][3];
        for (int y = 0; y < 4; ++y) for (int x = 4; x < 8; ++x) { o[y * 8 + x] = kDct2Weights[c][4]; o[x * 8 + y] = kDct2Weights[c][4]; }
        for (int y = 4; y < 8; ++y) for (int x = 4; x < 8; ++x) o[y * 8 + x] = kDct2Weights[c][5];
      }
      return 64;
    }
    case 3: {
      float w4[3 * 16];
      GetQuantWeights(4, 4, kDct4.b, kDct4.num, w4);
      w->assign(3 * 64, 0.0f);
      for (int c = 0; c < 3; ++c) {
        for (int y = 0; y < 8; ++y) for (int x = 0; x < 8; ++x)
          (*w)[c * 64 + y * 8 + x] = w4[c * 16 + (y / 2) * 4 + (x / 2)];
        // library dct4multipliers are all 1.0: w[1], w[8], w[9] unchanged
      }
      return 64;
    }
    case 9: {
      float w48[3 * 32];
      GetQuantWeights(4, 8, kDct4x8.b, kDct4x8.num, w48);
      w->assign(3 * 64, 0.0f);
      for (int c = 0; c < 3; ++c)
        for (int y = 0; y < 8; ++y) for (int x = 0; x < 8; ++x)
          (*w)[c * 64 + y * 8 + x] = w48[c * 32 + (y / 2) * 8 + x];
      return 64;
    }
    default: w->clear(); return 0;
  }
  w->assign((size_t)3 * rows * cols, 0.0f);
  GetQuantWeights(rows, cols, bp->b, bp->num, w->data());
  return rows * cols;
}

}  // namespace jxo
