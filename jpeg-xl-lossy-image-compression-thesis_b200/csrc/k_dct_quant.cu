// K7 (DCT8 path) — forward 8x8 DCT of X, Y, B + DC extraction + adaptive quantisation
// (stage U5: libjxl enc_group.cc ComputeCoefficients / QuantizeRoundtripYBlockAC /
// AdjustQuantBlockAC / QuantizeBlockAC and enc_modular.cc AddVarDCTDC [UPSTREAM]).
//
// Thread mapping: a warp owns 4 horizontally adjacent blocks, 8 lanes per block.  Lane r
// loads pixel row r of its block as two 128-bit loads (a warp reads 128 contiguous bytes
// per pixel row), runs the horizontal 8-point DCT in registers, the 8x8 tile is transposed
// across the 8 lanes with 12 warp shuffles, and the vertical DCT leaves lane h holding
// coefficient row h of the stored layout (index h*8 + v = hfreq*8 + vfreq).  Block-wide
// sums of the quantisation heuristics are xor-butterfly reductions over the 8 lanes.
// Quantised coefficients are staged in shared memory in scan order and leave the CTA as one
// contiguous 12 KB run of 128-bit stores (32 blocks x 3 channels x 64 x int16).
// HBM: 12 B/px in, 6 B/px out (+ ~0.3 B/px side data) -> 18.3 B/px.
#include "jxl_common.cuh"
#include "kernels.h"

namespace jxlb {

struct BlockSums { float hf, err, vals, nz[4], maxerr[4]; };

__device__ __forceinline__ float adjust_quant_bias(int c, int q) {
  const float b0 = 1.0f - 0.05465007330715401f, b1 = 1.0f - 0.07005449891748593f, b2 = 1.0f - 0.049935103337343655f;
  const float bc = c == 0 ? b0 : (c == 1 ? b1 : b2);
  if (q == 0) return 0.0f;
  if (q == 1) return bc;
  if (q == -1) return -bc;
  const float fq = (float)q;
  return fq - 0.145f / fq;
}

// AdjustQuantBlockAC for a DCT8 block; `in` is this lane's coefficient row h, `w` its weights.
// Returns the adjusted quant; thr[4] is updated in place.
__device__ __forceinline__ int adjust_quant_dct8(const float in[8], const float* w, int h, int c, float scale,
                                                 float qm_mul, int quant, float thr[4]) {
  const float qac = scale * (float)quant;
  float r_hf = 0.0f, r_err = 0.0f, r_vals = 0.0f, nzA = 0.0f, nzB = 0.0f, meA = 0.0f, meB = 0.0f;
  const int yfix = h >= 4 ? 2 : 0;
#pragma unroll
  for (int v = 0; v < 8; ++v) {
    if (h == 0 && v == 0) continue;
    const int hfix = yfix + (v >= 4 ? 1 : 0);
    const float val = in[v] * ((w[v] * qac) * qm_mul);
    const float vq = (fabsf(val) < thr[hfix]) ? 0.0f : rintf(val);
    const float err = fabsf(val - vq);
    r_err += err;
    r_vals += fabsf(vq);
    if (c == 1 && vq == 0.0f) { if (v >= 4) { if (meB < err) meB = err; } else { if (meA < err) meA = err; } }
    if (vq != 0.0f) {
      if (v >= 4) nzB += fabsf(vq); else nzA += fabsf(vq);
      const bool in_corner = h >= 7 && v >= 7;
      const bool on_border = h == 7 || v == 7;
      const bool in_larger = v >= 4 && h >= 4;
      if (in_corner || (on_border && in_larger)) r_hf += fabsf(val);
    }
  }
  const float sum_hf = tree8_sum(r_hf), sum_err = tree8_sum(r_err), sum_vals = tree8_sum(r_vals);
  (void)sum_err;
  float hfNZ[4], hfME[4];
  hfNZ[0] = tree8_sum(h < 4 ? nzA : 0.0f);
  hfNZ[1] = tree8_sum(h < 4 ? nzB : 0.0f);
  hfNZ[2] = tree8_sum(h < 4 ? 0.0f : nzA);
  hfNZ[3] = tree8_sum(h < 4 ? 0.0f : nzB);
  if (c == 1) {
    hfME[0] = tree8_max(h < 4 ? meA : 0.0f);
    hfME[1] = tree8_max(h < 4 ? meB : 0.0f);
    hfME[2] = tree8_max(h < 4 ? 0.0f : meA);
    hfME[3] = tree8_max(h < 4 ? 0.0f : meB);
    if (sum_vals * 8 < 1.0f) {
      const double kLimit = 0.46, kMul = 0.9999;
      const int orig = quant;
      int nq = quant;
#pragma unroll
      for (int i = 1; i < 4; ++i) {
        if (nq == orig && hfNZ[i] == 0.0f && (double)hfME[i] > kLimit) nq = orig + 1;
      }
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
    if (mul * sum_hf >= all) {
      quant = (int)((float)quant + mul * sum_hf / all);
      if (quant >= 256) quant = 255;
    }
  }
  if (hfNZ[0] + hfNZ[1] + hfNZ[2] + hfNZ[3] < 11) {
    quant += 1;
    if (quant >= 256) quant = 255;
  }
  return quant;
}

__device__ __forceinline__ void quantize_row(const float in[8], const float* w, int h, float qac_mul, const float thr[4],
                                             int out[8]) {
  const int yfix = h >= 4 ? 2 : 0;
#pragma unroll
  for (int v = 0; v < 8; ++v) {
    const float t = thr[yfix + (v >= 4 ? 1 : 0)];
    const float q = w[v] * qac_mul;
    const float val = q * in[v];
    int vq = (fabsf(val) >= t) ? (int)rintf(val) : 0;
    if (h == 0 && v == 0) vq = 0;
    vq = vq > 32767 ? 32767 : (vq < -32767 ? -32767 : vq);
    out[v] = vq;
  }
}

constexpr int kRowsPerCta = 4;  // block rows processed by one CTA (amortises table staging)

__global__ void __launch_bounds__(256, 3) k_dct8_quant(const float* __restrict__ X, const float* __restrict__ Y,
                                                    const float* __restrict__ B, FrameDim fd,
                                                    const QuantDev* __restrict__ qd, const float* __restrict__ weights,
                                                    const float* __restrict__ dequant_y, const uint8_t* __restrict__ izz,
                                                    const int8_t* __restrict__ cmap, float x_qm_mul, float b_qm_mul,
                                                    int adjust, int32_t* __restrict__ raw_qf,
                                                    int16_t* __restrict__ coeffs, int16_t* __restrict__ dc_quant,
                                                    uint8_t* __restrict__ nzeros, uint16_t* __restrict__ nzcount,
                                                    uint16_t* __restrict__ lastk) {
  __shared__ float s_w[3][8][9];
  __shared__ float s_dqy[8][9];
  __shared__ uint8_t s_izz[64];
  __shared__ __align__(16) int16_t s_out[32][3][64];
  const int t = threadIdx.x;
  for (int i = t; i < 192; i += 256) s_w[i / 64][(i % 64) / 8][i % 8] = weights[i];
  if (t < 64) { s_dqy[t / 8][t % 8] = dequant_y[t]; s_izz[t] = izz[t]; }
  __syncthreads();
  const int h = t & 7;            // lane within the block group = pixel row, later coefficient row
  const int bl = t >> 3;          // block within the CTA row (0..31)
  const int gx = blockIdx.x;      // group column
  const float scale = qd->scale, inv_gs = qd->inv_global_scale;
  const int quant_dc = qd->quant_dc;
  const size_t nblk = (size_t)fd.bxs * fd.bys;
  for (int rr = 0; rr < kRowsPerCta; ++rr) {
    const int by = blockIdx.y * kRowsPerCta + rr;
    if (by >= fd.bys) break;
    int bx = gx * 32 + bl;
    const bool active = bx < fd.bxs;
    if (!active) bx = fd.bxs - 1;
    // ---- load + forward DCT ---------------------------------------------------------------
    float c[3][8];
    const size_t po = (size_t)(by * 8 + h) * fd.pitch + (size_t)bx * 8;
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
      const float* P = ch == 0 ? X : (ch == 1 ? Y : B);
      const float4 a = *reinterpret_cast<const float4*>(P + po);
      const float4 b = *reinterpret_cast<const float4*>(P + po + 4);
      c[ch][0] = a.x; c[ch][1] = a.y; c[ch][2] = a.z; c[ch][3] = a.w;
      c[ch][4] = b.x; c[ch][5] = b.y; c[ch][6] = b.z; c[ch][7] = b.w;
    }
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
      dct8_scaled(c[ch]);       // horizontal: lane = pixel row y, regs = hfreq
      transpose8(c[ch], h);     // lane = hfreq, regs = pixel row y
      dct8_scaled(c[ch]);       // vertical: lane = hfreq h, regs = vfreq v
    }
    const float dcx = c[0][0], dcy = c[1][0], dcb = c[2][0];  // meaningful on lane h == 0
    // ---- quant adjust ---------------------------------------------------------------------
    const size_t bi = (size_t)by * fd.bxs + bx;
    int quant = raw_qf[bi];
    float thr_y[4] = {0.58f, 0.64f, 0.64f, 0.64f};
    if (adjust) {
      const int orig = quant;
      int maxq = 0;
      {
        float thr[4] = {0.58f, 0.64f, 0.64f, 0.64f};
        const int q1 = adjust_quant_dct8(c[1], &s_w[1][h][0], h, 1, scale, 1.0f, orig, thr);
        thr_y[0] = thr[0]; thr_y[1] = thr[1]; thr_y[2] = thr[2]; thr_y[3] = thr[3];
        maxq = q1;
      }
      {
        float thr[4] = {0.58f, 0.64f, 0.64f, 0.64f};
        const int q0 = adjust_quant_dct8(c[0], &s_w[0][h][0], h, 0, scale, x_qm_mul, orig, thr);
        maxq = max(maxq, q0);
      }
      {
        float thr[4] = {0.58f, 0.64f, 0.64f, 0.64f};
        const int q2 = adjust_quant_dct8(c[2], &s_w[2][h][0], h, 2, scale, b_qm_mul, orig, thr);
        maxq = max(maxq, q2);
      }
      quant = maxq;
    } else {
      thr_y[0] = 0.56f; thr_y[1] = thr_y[2] = thr_y[3] = 0.62f;
    }
    // ---- quantise Y, roundtrip, remove chroma-from-luma, quantise X and B -----------------
    const float qac = scale * (float)quant;
    int q[3][8];
    quantize_row(c[1], &s_w[1][h][0], h, qac * 1.0f, thr_y, q[1]);
    const float inv_qac = inv_gs / (float)quant;
    const int tx = bx >> 3, ty = by >> 3;
    const float x_factor = 0.0f + (float)cmap[(size_t)ty * fd.txs + tx] / 84.0f;
    const float b_factor = 1.0f + (float)cmap[(size_t)fd.txs * fd.tys + (size_t)ty * fd.txs + tx] / 84.0f;
#pragma unroll
    for (int v = 0; v < 8; ++v) {
      const float yrt = (adjust_quant_bias(1, q[1][v]) * s_dqy[h][v]) * inv_qac;
      c[0][v] = __fmaf_rn(-x_factor, yrt, c[0][v]);
      c[2][v] = __fmaf_rn(-b_factor, yrt, c[2][v]);
    }
    {
      const float thr[4] = {0.58f, 0.64f, 0.64f, 0.64f};
      quantize_row(c[0], &s_w[0][h][0], h, qac * x_qm_mul, thr, q[0]);
      quantize_row(c[2], &s_w[2][h][0], h, qac * b_qm_mul, thr, q[2]);
    }
    // ---- stage in scan order, count non-zeros --------------------------------------------
#pragma unroll
    for (int slot = 0; slot < 3; ++slot) {
      const int ch = slot == 0 ? 1 : (slot == 1 ? 0 : 2);  // token order Y, X, B
      int nz = 0, last = 0;
#pragma unroll
      for (int v = 0; v < 8; ++v) {
        const int k = s_izz[h * 8 + v];
        s_out[bl][slot][k] = (int16_t)q[ch][v];
        if (q[ch][v] != 0) { nz++; last = max(last, k); }
      }
      nz = tree8_isum(nz);
      last = tree8_imax(last);
      if (h == 0 && active) {
        nzeros[(size_t)ch * nblk + bi] = (uint8_t)nz;
        nzcount[(size_t)ch * nblk + bi] = (uint16_t)nz;
        lastk[(size_t)ch * nblk + bi] = (uint16_t)last;
      }
    }
    // ---- DC (AddVarDCTDC) + side data on lane 0 --------------------------------------------
    if (h == 0 && active) {
      raw_qf[bi] = quant;
      const float gsq = scale * (float)quant_dc;
      const float inv_quant_dc = inv_gs / (float)quant_dc;
      const float y_factor = inv_quant_dc * (1.0f / 512.0f);
      const float qy = roundf(dcy * (512.0f * gsq));
      const float qx = roundf((dcx - qy * (y_factor * 0.0f)) * (4096.0f * gsq));
      const float qb = roundf((dcb - qy * (y_factor * 1.0f)) * (256.0f * gsq));
      const int iy = (int)qy, ix = (int)qx, ib = (int)qb;
      dc_quant[0 * nblk + bi] = (int16_t)(ix > 32767 ? 32767 : (ix < -32768 ? -32768 : ix));
      dc_quant[1 * nblk + bi] = (int16_t)(iy > 32767 ? 32767 : (iy < -32768 ? -32768 : iy));
      dc_quant[2 * nblk + bi] = (int16_t)(ib > 32767 ? 32767 : (ib < -32768 ? -32768 : ib));
    }
    __syncthreads();
    // ---- coalesced copy-out: blocks of one group row are contiguous ------------------------
    {
      const int g = (by >> 5) * fd.gxs + gx;
      const int nvalid = min(32, fd.bxs - gx * 32);
      int16_t* dst = coeffs + ((size_t)g * kGroupBlocks + (size_t)(by & 31) * 32) * 192;
      const uint4* src4 = reinterpret_cast<const uint4*>(&s_out[0][0][0]);
      uint4* dst4 = reinterpret_cast<uint4*>(dst);
      const int n16 = nvalid * 24;  // 384 bytes per block = 24 x 16 B
      for (int i = t; i < n16; i += 256) dst4[i] = src4[i];
    }
    __syncthreads();
  }
}

void launch_dct8_quant(const float* x, const float* y, const float* b, const FrameDim& fd, const QuantDev* qd,
                       const float* weights, const float* dequant_y, const uint8_t* izz, const int8_t* cmap,
                       float x_qm_mul, float b_qm_mul, int adjust, int32_t* raw_qf, int16_t* coeffs, int16_t* dc_quant,
                       uint8_t* nzeros, uint16_t* nzcount, uint16_t* lastk, cudaStream_t s) {
  dim3 grid(fd.gxs, (fd.bys + kRowsPerCta - 1) / kRowsPerCta);
  ++g_kernel_launches;
  k_dct8_quant<<<grid, 256, 0, s>>>(x, y, b, fd, qd, weights, dequant_y, izz, cmap, x_qm_mul, b_qm_mul, adjust, raw_qf,
                                    coeffs, dc_quant, nzeros, nzcount, lastk);
}

}  // namespace jxlb
