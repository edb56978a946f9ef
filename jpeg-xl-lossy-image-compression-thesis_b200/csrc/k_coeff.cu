// K7 (general) — forward transform, DC extraction and adaptive quantisation for every AC strategy the
// search can pick (stage U5: libjxl enc_group.cc ComputeCoefficients / AdjustQuantBlockAC /
// QuantizeBlockAC / QuantizeRoundtripYBlockAC, enc_modular.cc AddVarDCTDC, dct_util
// DCFromLowestFrequencies [UPSTREAM]); same arithmetic and operation order as oracle/jxo_coef.cc.
// The DCT8-only frame (BASELINE config 2) keeps its specialised kernel (k_dct_quant.cu).
//
// One CTA per 32x32-pixel square: the XYB tile is staged in shared memory, every warp takes the
// transforms whose first block lies in the square round-robin.  Lane y of the transform's group owns
// coefficient row y (transforms.cuh); block-wide sums of the heuristics are xor-butterflies over the
// rows.  Quantised coefficients go to the group arena in scan order through the inverse natural
// order table of the strategy's class.
#include "transforms.cuh"
#include "kernels.h"

namespace jxlb {

#ifndef JXLB_COEFF_WARPS
#define JXLB_COEFF_WARPS 4
#endif
constexpr int kCoeffWarps = JXLB_COEFF_WARPS;

struct CoeffShared {
  float px[3][32 * kTPitch];
  float buf[kCoeffWarps][3][32 * kTPitch];   // per warp: X, Y, B coefficient rows (transforms and quantisation work in place)
  // channel-parallel path (32-lane transforms): per-channel adjusted quant, Y's thresholds, per-channel DC values
  int cp_q[3];
  float cp_thr[4];
  float cp_dc[3][16];
};

__device__ __forceinline__ float quant_bias(int c, int q) {
  const float b0 = 1.0f - 0.05465007330715401f, b1 = 1.0f - 0.07005449891748593f, b2 = 1.0f - 0.049935103337343655f;
  const float bc = c == 0 ? b0 : (c == 1 ? b1 : b2);
  if (q == 0) return 0.0f;
  if (q == 1) return bc;
  if (q == -1) return -bc;
  const float fq = (float)q;
  return fq - 0.145f / fq;
}

__device__ __forceinline__ float clamp1(float v, float lo, float hi) { return v < lo ? lo : (v > hi ? hi : v); }

// oracle AdjustQuantBlockAC for the group's lanes (lane y = coefficient row y); returns the adjusted quant
template <int S>
__device__ int adjust_quant(const float* coef, const float* __restrict__ qm, int c, float scale, float qm_mul, int quant,
                            float thr[4], int gl) {
  constexpr int R = StratDim<S>::R, C = StratDim<S>::C;
  constexpr int W = R > C ? R : C, H = R > C ? C : R, xs = W / 8, ys = H / 8;
  if constexpr (S == kStratDCT4X4 || S == kStratDCT4X8 || S == kStratDCT8X4) return quant;
  const float qac = scale * (float)quant;
  if (xs > 1 || ys > 1) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      thr[i] -= clamp1(0.003f * (float)(xs * ys), 0.f, (c > 0 ? 0.08f : 0.12f));
      if (thr[i] < 0.54f) thr[i] = 0.54f;
    }
  }
  float r_hf = 0.0f, r_err = 0.0f, r_vals = 0.0f, nzA = 0.0f, nzB = 0.0f, meA = 0.0f, meB = 0.0f;
  if (gl < H) {
    const int y = gl;
    const int yfix = y >= H / 2 ? 2 : 0;
    // (weights are read as 16-byte vectors: rows are 32-byte aligned; a quarter of the global-load instructions)
#pragma unroll 2
    for (int x4 = 0; x4 < W; x4 += 4) {
      const float4 w4 = __ldg(reinterpret_cast<const float4*>(qm + y * W + x4));
      const float wv[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
      const int x = x4 + e;
      if (x < xs && y < ys) continue;
      const int hfix = yfix + (x >= W / 2 ? 1 : 0);
      const float val = coef[y * kTPitch + x] * (wv[e] * qac * qm_mul);
      const float v = (fabsf(val) < thr[hfix]) ? 0.0f : rintf(val);
      const float err = fabsf(val - v);
      r_err += err;
      r_vals += fabsf(v);
      if (c == 1 && v == 0.0f) { if (x >= W / 2) { if (meB < err) meB = err; } else { if (meA < err) meA = err; } }
      if (v != 0.0f) {
        if (x >= W / 2) nzB += fabsf(v); else nzA += fabsf(v);
        const bool in_corner = y >= 7 * ys && x >= 7 * xs;
        const bool on_border = y == H - 1 || x == W - 1;
        const bool in_larger_corner = x >= 4 * xs && y >= 4 * ys;
        if (in_corner || (on_border && in_larger_corner)) r_hf += fabsf(val);
      }
      }
    }
  }
  const bool top = gl < H / 2;
  const float sum_hf_rc = group_sum<H>(r_hf), sum_err = group_sum<H>(r_err), sum_vals = group_sum<H>(r_vals);
  float hfNZ[4];
  hfNZ[0] = group_sum<H>(top ? nzA : 0.0f);
  hfNZ[1] = group_sum<H>(top ? nzB : 0.0f);
  hfNZ[2] = group_sum<H>(top ? 0.0f : nzA);
  hfNZ[3] = group_sum<H>(top ? 0.0f : nzB);
  if (c == 1) {
    float hfME[4];
    float m0 = top ? meA : 0.0f, m1 = top ? meB : 0.0f, m2 = top ? 0.0f : meA, m3 = top ? 0.0f : meB;
#pragma unroll
    for (int st = H / 2; st >= 1; st >>= 1) {
      m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, st)); m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, st));
      m2 = fmaxf(m2, __shfl_xor_sync(0xffffffffu, m2, st)); m3 = fmaxf(m3, __shfl_xor_sync(0xffffffffu, m3, st));
    }
    hfME[0] = m0; hfME[1] = m1; hfME[2] = m2; hfME[3] = m3;
    if (sum_vals * 8 < (float)(xs * ys)) {
      const double kLimit = 0.46, kMul = 0.9999;
      const int orig = quant;
      int nq = quant;
#pragma unroll
      for (int i = 1; i < 4; ++i) if (nq == orig && hfNZ[i] == 0.0f && (double)hfME[i] > kLimit) nq = orig + 1;
      quant = nq;
      if (hfNZ[3] == 0.0f && (double)hfME[3] > kLimit) {
        thr[3] = (float)(kMul * (double)hfME[3] * (double)nq / (double)orig);
      } else if ((hfNZ[1] == 0.0f && (double)hfME[1] > kLimit) || (hfNZ[2] == 0.0f && (double)hfME[2] > kLimit)) {
        const float m = hfME[1] > hfME[2] ? hfME[1] : hfME[2];
        thr[1] = (float)(kMul * (double)m * (double)nq / (double)orig);
        thr[2] = thr[1];
      } else if (hfNZ[0] == 0.0f && (double)hfME[0] > kLimit) {
        thr[0] = (float)(kMul * (double)hfME[0] * (double)nq / (double)orig);
      }
    }
  }
  {
    const float all = hfNZ[0] + hfNZ[1] + hfNZ[2] + hfNZ[3] + 1;
    const float mul = c == 0 ? 70.0f : (c == 1 ? 30.0f : 60.0f);
    if (mul * sum_hf_rc >= all) {
      quant = (int)((float)quant + mul * sum_hf_rc / all);
      if (quant >= 256) quant = 255;
    }
  }
  if constexpr (S == kStratDCT) {
    if (hfNZ[0] + hfNZ[1] + hfNZ[2] + hfNZ[3] < 11) { quant += 1; if (quant >= 256) quant = 255; }
  }
  if constexpr (S == kStratDCT16X16 || S == kStratDCT32X32 || S == kStratDCT32X16 || S == kStratDCT16X32 ||
                S == kStratDCT16X8 || S == kStratDCT8X16) {
    // (oracle: strategy >= DCT16X16, i.e. every multi-block transform of the emitted set)
    const double kMul1[4][3] = {{0.22080615753848404, 0.45797479824262011, 0.29859235095977965},
                                {0.70109486510286834, 0.16185281305512639, 0.14387691730035473},
                                {0.114985964456218638, 0.44656840441027695, 0.10587658215149048},
                                {0.46849665264409396, 0.41239077937781954, 0.088667407767185444}};
    const double kMul2[4][3] = {{0.27450281941822197, 1.1255766549984996, 0.98950459134128388},
                                {0.4652168675598285, 0.40945807983455818, 0.36581899811751367},
                                {0.28034972424715715, 0.9182653201929738, 1.5581531543057416},
                                {0.26873118114033728, 0.68863712390392484, 1.2082185408666786}};
    const double kQuantNormalizer = 2.2942708343284721;
    const double se = (double)sum_err * kQuantNormalizer;
    const double sv = (double)sum_vals * kQuantNormalizer;
    int ix = 3;
    if (S == kStratDCT32X16 || S == kStratDCT16X32) ix = 1;
    else if (S == kStratDCT16X16) ix = 0;
    else if (S == kStratDCT32X32) ix = 2;
    const double lim = kMul1[ix][c] * (double)(xs * ys * 64) + kMul2[ix][c] * sv;
    int step = (int)(se / lim);
    if (step >= 2) step = 2;
    if (step < 0) step = 0;
    if (se > lim) { quant += step; if (quant >= 256) quant = 255; }
  }
  return quant;
}

// oracle QuantizeBlockAC for row gl; quantised ints are written to `out` (same [H][kTPitch] layout, int bits)
template <int S>
__device__ __forceinline__ void quantize_rows(const float* coef, const float* __restrict__ qm, int c, float qac_mul, float thr[4],
                                              int* out, int gl) {
  constexpr int R = StratDim<S>::R, C = StratDim<S>::C;
  constexpr int W = R > C ? R : C, H = R > C ? C : R, xs = W / 8, ys = H / 8;
  if (c != 1 && xs * ys >= 4) {
#pragma unroll
    for (int i = 0; i < 4; ++i) { thr[i] -= 0.00744f * (float)(xs * ys); if (thr[i] < 0.5f) thr[i] = 0.5f; }
  }
  if (gl < H) {
    const int y = gl, yfix = (y >= H / 2) ? 2 : 0;
#pragma unroll 2
    for (int x4 = 0; x4 < W; x4 += 4) {
      const float4 w4 = __ldg(reinterpret_cast<const float4*>(qm + y * W + x4));
      const float wv[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int x = x4 + e;
        const float t = thr[yfix + (x >= W / 2 ? 1 : 0)];
        const float q = wv[e] * qac_mul;
        const float val = q * coef[y * kTPitch + x];
        int v = (fabsf(val) >= t) ? (int)rintf(val) : 0;
        if (x < xs && y < ys) v = 0;
        v = v > 32767 ? 32767 : (v < -32767 ? -32767 : v);
        out[y * kTPitch + x] = v;
      }
    }
  }
}

__device__ __forceinline__ float resample_scale(int n_from, int n_to, int k) {
  if (n_to == 1) return 1.0f;
  if (n_from == 16) return k == 0 ? 1.e+00f : 9.017642e-01f;
  return k == 0 ? 1.e+00f : (k == 1 ? 9.7488683e-01f : (k == 2 ? 9.017642e-01f : 7.870549e-01f));
}

template <int N> __device__ __forceinline__ void idct_small(float* v) { idct1d<N>(v); }

// oracle DcFromLowestFrequencies: dc[cy][cx] from the cy x cx lowest frequencies (serial, tiny)
template <int S>
__device__ void dc_from_llf(const float* coef, float* dc /*[cy*cx]*/) {
  constexpr int R = StratDim<S>::R, C = StratDim<S>::C, cy = R / 8, cx = C / 8;
  if constexpr (cx == 1 && cy == 1) { dc[0] = coef[0]; return; }
  else {
    float llf[16], t[16];
#pragma unroll
    for (int vf = 0; vf < cy; ++vf)
#pragma unroll
      for (int hf = 0; hf < cx; ++hf) {
        const float cv = (R >= C) ? coef[hf * kTPitch + vf] : coef[vf * kTPitch + hf];
        llf[vf * cx + hf] = cv * resample_scale(R, cy, vf) * resample_scale(C, cx, hf);
      }
#pragma unroll
    for (int hf = 0; hf < cx; ++hf) {
      float v[cy];
#pragma unroll
      for (int y = 0; y < cy; ++y) v[y] = llf[y * cx + hf];
      idct1d<cy>(v);
#pragma unroll
      for (int y = 0; y < cy; ++y) t[y * cx + hf] = v[y];
    }
#pragma unroll
    for (int y = 0; y < cy; ++y) {
      float v[cx];
#pragma unroll
      for (int x = 0; x < cx; ++x) v[x] = t[y * cx + x];
      idct1d<cx>(v);
#pragma unroll
      for (int x = 0; x < cx; ++x) dc[y * cx + x] = v[x];
    }
  }
}

struct CoeffArgs {
  FrameDim fd;
  const QuantDev* qd;
  AcsTables T;                      // weights / dequant per quant-table kind
  const uint16_t* inv_order[13];    // per order class: coefficient position -> scan index
  const int8_t* cmap;
  float x_qm_mul, b_qm_mul;
  int adjust;
  int32_t* raw_qf;
  int16_t* coeffs;
  int16_t* dc_quant;
  uint8_t* nzeros;
  uint16_t* nzcount;
  uint16_t* lastk;
};

// One lane group (GS = max(R, C) lanes) per transform, 32 / GS transforms side by side in a warp.
// Inactive groups run the same instruction stream on the tile origin with every store suppressed.
template <int S>
__device__ void process_transform(CoeffShared& sh, int warp, bool active, int ox, int oy, int bx, int by, const CoeffArgs& A,
                                  int kind, int order_class, int lane) {
  constexpr int R = StratDim<S>::R, C = StratDim<S>::C;
  constexpr int W = R > C ? R : C, H = R > C ? C : R, cxb = C / 8, cyb = R / 8, n = cxb * cyb, size = R * C;
  constexpr int GS = W;                                    // lanes per transform
  const FrameDim& fd = A.fd;
  const int gl = lane & (GS - 1), gbase = lane & ~(GS - 1), go = (lane / GS) * GS * kTPitch;
  // (every templated helper below has ONE call site inside a non-unrolled channel loop: the kernel holds ten
  // strategy instantiations and instruction-cache misses were 29 % of its stalls when each helper was inlined 3x)
  float* bufs[3] = {sh.buf[warp][0] + go, sh.buf[warp][1] + go, sh.buf[warp][2] + go};
  const int po = oy * kTPitch + ox;
#pragma unroll 1
  for (int c = 0; c < 3; ++c) fwd_transform<S>(sh.px[c] + po, kTPitch, bufs[c], bufs[c], gl);
  float* b0 = bufs[0]; float* b1 = bufs[1]; float* b2 = bufs[2];
  const size_t nblk = (size_t)fd.bxs * fd.bys;
  const size_t bi = (size_t)by * fd.bxs + bx;
  const float scale = A.qd->scale, inv_gs = A.qd->inv_global_scale;
  // ---- DC of every covered block (AddVarDCTDC quantisation: Y first, B with the 1.0 base correlation)
  if (gl == 0 && active) {
    float dc[3][16];
#pragma unroll 1
    for (int c = 0; c < 3; ++c) dc_from_llf<S>(bufs[c], dc[c]);
    const int quant_dc = A.qd->quant_dc;
    const float gsq = scale * (float)quant_dc;
    const float inv_quant_dc = inv_gs / (float)quant_dc;
    const float y_factor = inv_quant_dc * (1.0f / 512.0f);
    for (int j = 0; j < n; ++j) {
      const size_t bj = bi + (size_t)(j / cxb) * fd.bxs + (j % cxb);
      const float qy = roundf(dc[1][j] * (512.0f * gsq));
      const float qx = roundf((dc[0][j] - qy * (y_factor * 0.0f)) * (4096.0f * gsq));
      const float qb = roundf((dc[2][j] - qy * (y_factor * 1.0f)) * (256.0f * gsq));
      const int iy = (int)qy, ix = (int)qx, ib = (int)qb;
      A.dc_quant[0 * nblk + bj] = (int16_t)(ix > 32767 ? 32767 : (ix < -32768 ? -32768 : ix));
      A.dc_quant[1 * nblk + bj] = (int16_t)(iy > 32767 ? 32767 : (iy < -32768 ? -32768 : iy));
      A.dc_quant[2 * nblk + bj] = (int16_t)(ib > 32767 ? 32767 : (ib < -32768 ? -32768 : ib));
    }
  }
  // ---- quant adjust (channel order Y, X, B; the result is the maximum, Y's thresholds are kept)
  const float* qm = A.T.w[kind];
  const float* dq = A.T.dq[kind];
  int quant = active ? A.raw_qf[bi] : 1;
  float thr_y[4] = {0.58f, 0.64f, 0.64f, 0.64f};
  if (A.adjust) {
    const int orig = quant;
    int maxq = 0;
#pragma unroll 1
    for (int it = 0; it < 3; ++it) {
      const int c = it == 0 ? 1 : (it == 1 ? 0 : 2);
      float thr[4] = {0.58f, 0.64f, 0.64f, 0.64f};
      const float mulc = c == 0 ? A.x_qm_mul : (c == 1 ? 1.0f : A.b_qm_mul);
      maxq = max(maxq, adjust_quant<S>(bufs[c], qm + c * size, c, scale, mulc, orig, thr, gl));
      if (c == 1) { thr_y[0] = thr[0]; thr_y[1] = thr[1]; thr_y[2] = thr[2]; thr_y[3] = thr[3]; }
    }
    quant = __shfl_sync(0xffffffffu, maxq, gbase);
    thr_y[0] = __shfl_sync(0xffffffffu, thr_y[0], gbase); thr_y[1] = __shfl_sync(0xffffffffu, thr_y[1], gbase);
    thr_y[2] = __shfl_sync(0xffffffffu, thr_y[2], gbase); thr_y[3] = __shfl_sync(0xffffffffu, thr_y[3], gbase);
  } else {
    thr_y[0] = 0.56f; thr_y[1] = thr_y[2] = thr_y[3] = 0.62f;
  }
  // ---- quantise Y in place, roundtrip, remove chroma-from-luma, quantise X and B in place (all row-local)
  const float qac = scale * (float)quant;
  const int* qy = reinterpret_cast<const int*>(b1);
  const float inv_qac = inv_gs / (float)quant;
  const int tx = bx >> 3, ty = by >> 3;
  const float x_factor = 0.0f + (float)A.cmap[(size_t)ty * fd.txs + tx] / 84.0f;
  const float b_factor = 1.0f + (float)A.cmap[(size_t)fd.txs * fd.tys + (size_t)ty * fd.txs + tx] / 84.0f;
#pragma unroll 1
  for (int it = 0; it < 3; ++it) {
    const int c = it == 0 ? 1 : (it == 1 ? 0 : 2);
    if (it == 1 && gl < H) {
#pragma unroll 2
      for (int x4 = 0; x4 < W; x4 += 4) {
        const float4 d4 = __ldg(reinterpret_cast<const float4*>(dq + size + gl * W + x4));
        const float dv[4] = {d4.x, d4.y, d4.z, d4.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int x = x4 + e;
          const float yrt = (quant_bias(1, qy[gl * kTPitch + x]) * dv[e]) * inv_qac;
          b0[gl * kTPitch + x] = __fmaf_rn(-x_factor, yrt, b0[gl * kTPitch + x]);
          b2[gl * kTPitch + x] = __fmaf_rn(-b_factor, yrt, b2[gl * kTPitch + x]);
        }
      }
    }
    float thr[4] = {0.58f, 0.64f, 0.64f, 0.64f};
    if (c == 1) { thr[0] = thr_y[0]; thr[1] = thr_y[1]; thr[2] = thr_y[2]; thr[3] = thr_y[3]; }
    const float mulc = c == 0 ? A.x_qm_mul : (c == 1 ? 1.0f : A.b_qm_mul);
    quantize_rows<S>(bufs[c], qm + c * size, c, qac * mulc, thr, reinterpret_cast<int*>(bufs[c]), gl);
  }
  // ---- scan-order output + non-zero statistics
  const uint16_t* inv = A.inv_order[order_class];
  constexpr int log2n = n == 1 ? 0 : (n == 2 ? 1 : (n == 4 ? 2 : (n == 8 ? 3 : 4)));
#pragma unroll 1
  for (int slot = 0; slot < 3; ++slot) {
    const int c = slot == 0 ? 1 : (slot == 1 ? 0 : 2);
    const int* src = reinterpret_cast<const int*>(bufs[c]);
    int nz = 0, last = 0;
    uint2 inv4 = make_uint2(0u, 0u);
    if (gl < H && active) {
      for (int x = 0; x < W; ++x) {
        const int v = src[gl * kTPitch + x];
        // (scan indices of four positions per 8-byte load)
        if ((x & 3) == 0) inv4 = __ldg(reinterpret_cast<const uint2*>(inv + gl * W + x));
        const int k = (int)(((x & 2) ? inv4.y : inv4.x) >> ((x & 1) * 16)) & 0xFFFF;
        const int j = k >> 6;
        const int cbx = bx + (j % cxb), cby = by + (j / cxb);
        const int g = (cby >> 5) * fd.gxs + (cbx >> 5);
        const size_t blk = (size_t)g * kGroupBlocks + (size_t)(cby & 31) * 32 + (cbx & 31);
        A.coeffs[(blk * 3 + slot) * 64 + (k & 63)] = (int16_t)v;
        if (v != 0) { ++nz; last = max(last, k); }
      }
    }
    nz = group_isum<H>(nz);
#pragma unroll
    for (int st = H / 2; st >= 1; st >>= 1) last = max(last, __shfl_xor_sync(0xffffffffu, last, st));
    if (gl == 0 && active) {
      const int shared = (nz + n - 1) >> log2n;
      A.nzcount[(size_t)c * nblk + bi] = (uint16_t)nz;
      A.lastk[(size_t)c * nblk + bi] = (uint16_t)last;
      for (int j = 0; j < n; ++j) A.nzeros[(size_t)c * nblk + bi + (size_t)(j / cxb) * fd.bxs + (j % cxb)] = (uint8_t)shared;
    }
  }
  if (gl == 0 && active) for (int j = 0; j < n; ++j) A.raw_qf[bi + (size_t)(j / cxb) * fd.bxs + (j % cxb)] = quant;
  __syncwarp();
}

// Channel-parallel form for the 32-lane transforms (32x32, 32x16, 16x32): such a transform fills a warp, a 32x32 square
// holds one or two of them, and with one warp per transform the other warps of the CTA had nothing to do (ncu: 6.5 %
// active warps, 12 % issue utilisation, a 32x32 transform took 160 k cycles).  Here warps 0..2 take channels X, Y, B of the
// SAME transform and meet at CTA barriers where the channels depend on each other (the adjusted quant is the maximum
// over the channels; X and B need the quantised Y); warp 3 quantises the DC values and writes the per-block side data.
// Every per-channel step is the code of process_transform, so the results are bit-identical.  Called by all warps.
template <int S>
__device__ void process_transform_cp(CoeffShared& sh, int warp, int ox, int oy, int bx, int by, const CoeffArgs& A, int kind,
                                     int order_class, int lane) {
  constexpr int R = StratDim<S>::R, C = StratDim<S>::C;
  constexpr int W = R > C ? R : C, H = R > C ? C : R, cxb = C / 8, cyb = R / 8, n = cxb * cyb, size = R * C;
  static_assert(W == 32, "channel-parallel path is for 32-lane transforms");
  const FrameDim& fd = A.fd;
  const int gl = lane;
  const int c = warp;                                   // 0 = X, 1 = Y, 2 = B, 3 = helper
  float* buf = sh.buf[warp][0];                         // this warp's coefficient rows
  const int* qy = reinterpret_cast<const int*>(sh.buf[1][0]);
  const int po = oy * kTPitch + ox;
  const size_t nblk = (size_t)fd.bxs * fd.bys;
  const size_t bi = (size_t)by * fd.bxs + bx;
  const float scale = A.qd->scale, inv_gs = A.qd->inv_global_scale;
  const float* qm = A.T.w[kind];
  const float* dq = A.T.dq[kind];
  const int orig = A.raw_qf[bi];
  const float mulc = c == 0 ? A.x_qm_mul : (c == 1 ? 1.0f : A.b_qm_mul);
  // ---- phase 1: transform + quant adjust of the warp's channel
  if (c < 3) {
    fwd_transform<S>(sh.px[c] + po, kTPitch, buf, buf, gl);
    if (gl == 0) dc_from_llf<S>(buf, sh.cp_dc[c]);
    if (A.adjust) {
      float thr[4] = {0.58f, 0.64f, 0.64f, 0.64f};
      const int q = adjust_quant<S>(buf, qm + c * size, c, scale, mulc, orig, thr, gl);
      if (gl == 0) {
        sh.cp_q[c] = q;
        if (c == 1) { sh.cp_thr[0] = thr[0]; sh.cp_thr[1] = thr[1]; sh.cp_thr[2] = thr[2]; sh.cp_thr[3] = thr[3]; }
      }
    }
  }
  __syncthreads();
  int quant = orig;
  float thr_y[4] = {0.56f, 0.62f, 0.62f, 0.62f};
  if (A.adjust) {
    quant = max(sh.cp_q[1], max(sh.cp_q[0], sh.cp_q[2]));
    thr_y[0] = sh.cp_thr[0]; thr_y[1] = sh.cp_thr[1]; thr_y[2] = sh.cp_thr[2]; thr_y[3] = sh.cp_thr[3];
  }
  const float qac = scale * (float)quant;
  const float inv_qac = inv_gs / (float)quant;
  // ---- phase 2: Y quantised in place; the helper warp does the DC values and the integer quant field meanwhile
  if (c == 1) quantize_rows<S>(buf, qm + size, 1, qac * 1.0f, thr_y, reinterpret_cast<int*>(buf), gl);
  if (c == 3 && gl == 0) {
    const int quant_dc = A.qd->quant_dc;
    const float gsq = scale * (float)quant_dc;
    const float inv_quant_dc = inv_gs / (float)quant_dc;
    const float y_factor = inv_quant_dc * (1.0f / 512.0f);
    for (int j = 0; j < n; ++j) {
      const size_t bj = bi + (size_t)(j / cxb) * fd.bxs + (j % cxb);
      const float fy = roundf(sh.cp_dc[1][j] * (512.0f * gsq));
      const float fx = roundf((sh.cp_dc[0][j] - fy * (y_factor * 0.0f)) * (4096.0f * gsq));
      const float fb = roundf((sh.cp_dc[2][j] - fy * (y_factor * 1.0f)) * (256.0f * gsq));
      const int iy = (int)fy, ix = (int)fx, ib = (int)fb;
      A.dc_quant[0 * nblk + bj] = (int16_t)(ix > 32767 ? 32767 : (ix < -32768 ? -32768 : ix));
      A.dc_quant[1 * nblk + bj] = (int16_t)(iy > 32767 ? 32767 : (iy < -32768 ? -32768 : iy));
      A.dc_quant[2 * nblk + bj] = (int16_t)(ib > 32767 ? 32767 : (ib < -32768 ? -32768 : ib));
      A.raw_qf[bj] = quant;
    }
  }
  __syncthreads();
  // ---- phase 3: X and B remove their chroma-from-luma share of the dequantised Y and quantise; every channel goes out
  if (c == 0 || c == 2) {
    const int tx = bx >> 3, ty = by >> 3;
    const float factor = c == 0 ? 0.0f + (float)A.cmap[(size_t)ty * fd.txs + tx] / 84.0f
                                : 1.0f + (float)A.cmap[(size_t)fd.txs * fd.tys + (size_t)ty * fd.txs + tx] / 84.0f;
    if (gl < H) {
#pragma unroll 2
      for (int x4 = 0; x4 < W; x4 += 4) {
        const float4 d4 = __ldg(reinterpret_cast<const float4*>(dq + size + gl * W + x4));
        const float dv[4] = {d4.x, d4.y, d4.z, d4.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int x = x4 + e;
          const float yrt = (quant_bias(1, qy[gl * kTPitch + x]) * dv[e]) * inv_qac;
          buf[gl * kTPitch + x] = __fmaf_rn(-factor, yrt, buf[gl * kTPitch + x]);
        }
      }
    }
    float thr[4] = {0.58f, 0.64f, 0.64f, 0.64f};
    quantize_rows<S>(buf, qm + c * size, c, qac * mulc, thr, reinterpret_cast<int*>(buf), gl);
  }
  if (c < 3) {
    const uint16_t* inv = A.inv_order[order_class];
    constexpr int log2n = n == 1 ? 0 : (n == 2 ? 1 : (n == 4 ? 2 : (n == 8 ? 3 : 4)));
    const int slot = c == 1 ? 0 : (c == 0 ? 1 : 2);
    const int* src = reinterpret_cast<const int*>(buf);
    int nz = 0, last = 0;
    uint2 inv4 = make_uint2(0u, 0u);
    if (gl < H) {
      for (int x = 0; x < W; ++x) {
        const int v = src[gl * kTPitch + x];
        // (scan indices of four positions per 8-byte load)
        if ((x & 3) == 0) inv4 = __ldg(reinterpret_cast<const uint2*>(inv + gl * W + x));
        const int k = (int)(((x & 2) ? inv4.y : inv4.x) >> ((x & 1) * 16)) & 0xFFFF;
        const int j = k >> 6;
        const int cbx = bx + (j % cxb), cby = by + (j / cxb);
        const int g = (cby >> 5) * fd.gxs + (cbx >> 5);
        const size_t blk = (size_t)g * kGroupBlocks + (size_t)(cby & 31) * 32 + (cbx & 31);
        A.coeffs[(blk * 3 + slot) * 64 + (k & 63)] = (int16_t)v;
        if (v != 0) { ++nz; last = max(last, k); }
      }
    }
    nz = group_isum<H>(nz);
#pragma unroll
    for (int st = H / 2; st >= 1; st >>= 1) last = max(last, __shfl_xor_sync(0xffffffffu, last, st));
    if (gl == 0) {
      const int shared = (nz + n - 1) >> log2n;
      A.nzcount[(size_t)c * nblk + bi] = (uint16_t)nz;
      A.lastk[(size_t)c * nblk + bi] = (uint16_t)last;
      for (int j = 0; j < n; ++j) A.nzeros[(size_t)c * nblk + bi + (size_t)(j / cxb) * fd.bxs + (j % cxb)] = (uint8_t)shared;
    }
  }
  __syncthreads();   // the buffers and cp_* are reused by the next transform
}

// all transforms of strategy S whose first block lies in the square: `mask` has one bit per block of the square
template <int S>
__device__ void process_strategy(CoeffShared& sh, int warp, unsigned mask, int sbx, int sby, const CoeffArgs& A, int kind,
                                 int order_class, int lane) {
  constexpr int R = StratDim<S>::R, C = StratDim<S>::C;
  constexpr int GS = R > C ? R : C, GPW = 32 / GS;
  const int m = __popc(mask);
  if constexpr (GS == 32 && kCoeffWarps == 4) {
    // (mask is the same in every warp: the loop and the barriers inside are CTA-uniform)
    for (int idx = 0; idx < m; ++idx) {
      const int b = (int)__fns(mask, 0, idx + 1);
      const int lx = b & 3, ly = b >> 2;
      process_transform_cp<S>(sh, warp, lx * 8, ly * 8, sbx + lx, sby + ly, A, kind, order_class, lane);
    }
    return;
  }
  for (int base = warp * GPW; base < m; base += kCoeffWarps * GPW) {
    const int idx = base + lane / GS;
    const bool active = idx < m;
    const int b = active ? (int)__fns(mask, 0, idx + 1) : 0;
    const int lx = b & 3, ly = b >> 2;
    process_transform<S>(sh, warp, active, lx * 8, ly * 8, sbx + lx, sby + ly, A, kind, order_class, lane);
  }
}

__global__ void __launch_bounds__(kCoeffWarps * 32) k_coeff_general(const float* __restrict__ X, const float* __restrict__ Y,
                                                                    const float* __restrict__ B, const uint8_t* __restrict__ acs,
                                                                    CoeffArgs A) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  CoeffShared& sh = *reinterpret_cast<CoeffShared*>(smem_raw);
  const FrameDim& fd = A.fd;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int sbx = blockIdx.x * 4, sby = blockIdx.y * 4;
  const int bw = min(4, fd.bxs - sbx), bh = min(4, fd.bys - sby);
  for (int i = t; i < 32 * 32; i += kCoeffWarps * 32) {
    const int y = i >> 5, x = i & 31;
    const bool in = x < bw * 8 && y < bh * 8;
    const size_t g = (size_t)(sby * 8 + y) * fd.pitch + (size_t)sbx * 8 + x;
    sh.px[0][y * kTPitch + x] = in ? X[g] : 0.0f;
    sh.px[1][y * kTPitch + x] = in ? Y[g] : 0.0f;
    sh.px[2][y * kTPitch + x] = in ? B[g] : 0.0f;
  }
  // strategy of the square's blocks: lane b < 16 looks at block b; one ballot per strategy gives its transform list
  int my_s = -1;
  if (lane < 16) {
    const int lx = lane & 3, ly = lane >> 2;
    if (lx < bw && ly < bh) {
      const uint8_t a = acs[(size_t)(sby + ly) * fd.bxs + sbx + lx];
      if (a & 0x80) my_s = a & 0x7f;
    }
  }
  __syncthreads();
#define JXLB_STRATEGY(S, KIND, ORD) \
  { const unsigned mk = __ballot_sync(0xffffffffu, my_s == S); if (mk) process_strategy<S>(sh, warp, mk, sbx, sby, A, KIND, ORD, lane); }
  JXLB_STRATEGY(kStratDCT, 0, 0)
  JXLB_STRATEGY(kStratDCT4X4, 3, 1)
  JXLB_STRATEGY(kStratDCT4X8, 9, 1)
  JXLB_STRATEGY(kStratDCT8X4, 9, 1)
  JXLB_STRATEGY(kStratDCT16X8, 6, 4)
  JXLB_STRATEGY(kStratDCT8X16, 6, 4)
  JXLB_STRATEGY(kStratDCT16X16, 4, 2)
  JXLB_STRATEGY(kStratDCT32X16, 8, 6)
  JXLB_STRATEGY(kStratDCT16X32, 8, 6)
  JXLB_STRATEGY(kStratDCT32X32, 5, 3)
#undef JXLB_STRATEGY
}

void launch_coeff_general(const float* x, const float* y, const float* b, const uint8_t* acs, const FrameDim& fd,
                          const QuantDev* qd, const AcsTables& T, const uint16_t* const* inv_order, const int8_t* cmap,
                          float x_qm_mul, float b_qm_mul, int adjust, int32_t* raw_qf, int16_t* coeffs, int16_t* dc_quant,
                          uint8_t* nzeros, uint16_t* nzcount, uint16_t* lastk, cudaStream_t s) {
  // (function attributes are per device: set on every launch, a context may live on any GPU of the process)
  cudaFuncSetAttribute(k_coeff_general, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(CoeffShared));
  CoeffArgs A;
  A.fd = fd; A.qd = qd; A.T = T;
  for (int i = 0; i < 13; ++i) A.inv_order[i] = inv_order[i];
  A.cmap = cmap; A.x_qm_mul = x_qm_mul; A.b_qm_mul = b_qm_mul; A.adjust = adjust;
  A.raw_qf = raw_qf; A.coeffs = coeffs; A.dc_quant = dc_quant; A.nzeros = nzeros; A.nzcount = nzcount; A.lastk = lastk;
  ++g_kernel_launches;
  dim3 grid((fd.bxs + 3) / 4, (fd.bys + 3) / 4);
  k_coeff_general<<<grid, kCoeffWarps * 32, sizeof(CoeffShared), s>>>(x, y, b, acs, A);
}

}  // namespace jxlb
