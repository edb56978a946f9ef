// K7 (DCT8 frames, second design) — ONE THREAD PER 8x8 BLOCK.
//
// The first design (k_dct_quant.cu: 8 lanes per block, shuffle transposes, shuffle reductions) executes
// ~1800 warp instructions per 4 blocks and is issue-bound at 23 % of the HBM roofline (ncu: 58 M warp
// instructions for a 4K frame, profiles/).  Here a thread owns a whole block: both DCT passes, every
// block-wide sum of the quantisation heuristics and the scan-order packing are thread-local with
// compile-time register indices — no shuffles, no shared-memory transposes, no redundant control flow
// across lanes.  The float association is unchanged (per-row sequential sums, then the butterfly tree
// ((r0+r4)+(r2+r6))+((r1+r5)+(r3+r7)) the oracle uses), so the output is bit-identical.
//
// Two passes over the three channels: A = transform + AdjustQuantBlockAC (only the maximum quant and Y's
// thresholds survive), B = transform + quantise (Y first; X and B after removing the chroma-from-luma
// prediction of the dequantised Y).  The pixels are re-read in pass B (L1/L2 hits) instead of keeping
// 192 coefficients live.  Quantised coefficients are packed in scan order in registers (the zig-zag is
// a compile-time permutation) and leave as 128-bit stores, 384 contiguous bytes per block.
// Tables (weights, Y dequant) sit in __constant__ memory: all threads read the same entry.
#include "transforms.cuh"
#include "kernels.h"

namespace jxlb {

__constant__ float c_w8[192];     // quantisation weights X, Y, B (coefficient layout hf*8 + vf)
__constant__ float c_dqy8[64];    // Y dequant (1 / weight)

// natural coefficient order of an 8x8 block (JPEG zig-zag on the hf*8+vf layout; tests pin it): scan k -> position
__device__ constexpr int kZigzag8[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                                         41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                                         30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

__device__ __forceinline__ float tree8(const float r[8]) { return ((r[0] + r[4]) + (r[2] + r[6])) + ((r[1] + r[5]) + (r[3] + r[7])); }
__device__ __forceinline__ float half4_top(const float r[8]) { return ((r[0] + 0.0f) + (r[2] + 0.0f)) + ((r[1] + 0.0f) + (r[3] + 0.0f)); }
__device__ __forceinline__ float half4_bot(const float r[8]) { return ((0.0f + r[4]) + (0.0f + r[6])) + ((0.0f + r[5]) + (0.0f + r[7])); }

__device__ __forceinline__ float quant_bias_y(int q) {
  const float b1 = 1.0f - 0.07005449891748593f;
  if (q == 0) return 0.0f;
  if (q == 1) return b1;
  if (q == -1) return -b1;
  const float fq = (float)q;
  return fq - 0.145f / fq;
}

// loads the block's 8x8 pixels of one plane and transforms them: c[hf*8 + vf]
__device__ __forceinline__ void load_dct8(const float* __restrict__ P, size_t po, int pitch, float c[64]) {
  float t[64];
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(P + po + (size_t)r * pitch));
    const float4 b = __ldg(reinterpret_cast<const float4*>(P + po + (size_t)r * pitch + 4));
    float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    dct1d<8>(v);
#pragma unroll
    for (int x = 0; x < 8; ++x) t[r * 8 + x] = v[x];
  }
#pragma unroll
  for (int hf = 0; hf < 8; ++hf) {
    float v[8];
#pragma unroll
    for (int y = 0; y < 8; ++y) v[y] = t[y * 8 + hf];
    dct1d<8>(v);
#pragma unroll
    for (int vf = 0; vf < 8; ++vf) c[hf * 8 + vf] = v[vf];
  }
}

// oracle AdjustQuantBlockAC for a DCT8 block (xs = ys = 1), thread-local
__device__ __forceinline__ int adjust_quant_block(const float c[64], int ch, float scale, float qm_mul, int quant, float thr[4]) {
  const float qac = scale * (float)quant;
  float r_hf[8], r_err[8], r_vals[8], nzA[8], nzB[8];
  float me[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
  for (int y = 0; y < 8; ++y) {
    float hf = 0.0f, er = 0.0f, vs = 0.0f, a = 0.0f, b = 0.0f;
#pragma unroll
    for (int x = 0; x < 8; ++x) {
      if (y == 0 && x == 0) continue;
      const int hfix = (y >= 4 ? 2 : 0) + (x >= 4 ? 1 : 0);
      const float val = c[y * 8 + x] * ((c_w8[ch * 64 + y * 8 + x] * qac) * qm_mul);
      const float v = (fabsf(val) < thr[hfix]) ? 0.0f : rintf(val);
      const float err = fabsf(val - v);
      er += err;
      vs += fabsf(v);
      if (ch == 1 && v == 0.0f) { if (me[hfix] < err) me[hfix] = err; }
      if (v != 0.0f) {
        if (x >= 4) b += fabsf(v); else a += fabsf(v);
        const bool in_corner = y >= 7 && x >= 7;
        const bool on_border = y == 7 || x == 7;
        const bool in_larger = x >= 4 && y >= 4;
        if (in_corner || (on_border && in_larger)) hf += fabsf(val);
      }
    }
    r_hf[y] = hf; r_err[y] = er; r_vals[y] = vs; nzA[y] = a; nzB[y] = b;
  }
  const float sum_hf = tree8(r_hf), sum_vals = tree8(r_vals);
  float hfNZ[4];
  hfNZ[0] = half4_top(nzA); hfNZ[1] = half4_top(nzB); hfNZ[2] = half4_bot(nzA); hfNZ[3] = half4_bot(nzB);
  if (ch == 1) {
    if (sum_vals * 8 < 1.0f) {
      const double kLimit = 0.46, kMul = 0.9999;
      const int orig = quant;
      int nq = quant;
#pragma unroll
      for (int i = 1; i < 4; ++i) if (nq == orig && hfNZ[i] == 0.0f && (double)me[i] > kLimit) nq = orig + 1;
      quant = nq;
      if (hfNZ[3] == 0.0f && (double)me[3] > kLimit) {
        thr[3] = (float)(kMul * (double)me[3] * (double)nq / (double)orig);
      } else if ((hfNZ[1] == 0.0f && (double)me[1] > kLimit) || (hfNZ[2] == 0.0f && (double)me[2] > kLimit)) {
        const float m = me[1] > me[2] ? me[1] : me[2];
        thr[1] = (float)(kMul * (double)m * (double)nq / (double)orig);
        thr[2] = thr[1];
      } else if (hfNZ[0] == 0.0f && (double)me[0] > kLimit) {
        thr[0] = (float)(kMul * (double)me[0] * (double)nq / (double)orig);
      }
    }
  }
  {
    const float all = hfNZ[0] + hfNZ[1] + hfNZ[2] + hfNZ[3] + 1;
    const float mul = ch == 0 ? 70.0f : (ch == 1 ? 30.0f : 60.0f);
    if (mul * sum_hf >= all) {
      quant = (int)((float)quant + mul * sum_hf / all);
      if (quant >= 256) quant = 255;
    }
  }
  if (hfNZ[0] + hfNZ[1] + hfNZ[2] + hfNZ[3] < 11) { quant += 1; if (quant >= 256) quant = 255; }
  return quant;
}

// oracle QuantizeBlockAC (DCT8), thread-local; q[pos]
__device__ __forceinline__ void quantize_block(const float c[64], int ch, float qac_mul, const float thr[4], int q[64]) {
#pragma unroll
  for (int y = 0; y < 8; ++y)
#pragma unroll
    for (int x = 0; x < 8; ++x) {
      const float t = thr[(y >= 4 ? 2 : 0) + (x >= 4 ? 1 : 0)];
      const float qq = c_w8[ch * 64 + y * 8 + x] * qac_mul;
      const float val = qq * c[y * 8 + x];
      int v = (fabsf(val) >= t) ? (int)rintf(val) : 0;
      if (y == 0 && x == 0) v = 0;
      v = v > 32767 ? 32767 : (v < -32767 ? -32767 : v);
      q[y * 8 + x] = v;
    }
}

// packs q[] in scan order into 32 words, counts non-zeros and finds the last non-zero scan index, stores 128 B
__device__ __forceinline__ void emit_block(const int q[64], int16_t* __restrict__ dst, int& nz, int& last) {
  uint32_t w[32];
  nz = 0; last = 0;
#pragma unroll
  for (int k = 0; k < 64; k += 2) {
    const int a = q[kZigzag8[k]], b = q[kZigzag8[k + 1]];
    w[k >> 1] = ((uint32_t)a & 0xFFFFu) | ((uint32_t)b << 16);
    if (a != 0) { ++nz; last = k; }
    if (b != 0) { ++nz; last = k + 1; }
  }
  uint4* d4 = reinterpret_cast<uint4*>(dst);
#pragma unroll
  for (int i = 0; i < 8; ++i) d4[i] = make_uint4(w[4 * i], w[4 * i + 1], w[4 * i + 2], w[4 * i + 3]);
}

constexpr int kV2Threads = 64;

__global__ void __launch_bounds__(kV2Threads, 8) k_dct8_quant_v2(const float* __restrict__ X, const float* __restrict__ Y,
                                                             const float* __restrict__ B, FrameDim fd,
                                                             const QuantDev* __restrict__ qd, const int8_t* __restrict__ cmap,
                                                             float x_qm_mul, float b_qm_mul, int adjust,
                                                             int32_t* __restrict__ raw_qf, int16_t* __restrict__ coeffs,
                                                             int16_t* __restrict__ dc_quant, uint8_t* __restrict__ nzeros,
                                                             uint16_t* __restrict__ nzcount, uint16_t* __restrict__ lastk) {
  const int bx = blockIdx.x * kV2Threads + threadIdx.x, by = blockIdx.y;
  if (bx >= fd.bxs) return;
  const size_t po = (size_t)by * 8 * fd.pitch + (size_t)bx * 8;
  const size_t nblk = (size_t)fd.bxs * fd.bys, bi = (size_t)by * fd.bxs + bx;
  const float scale = qd->scale, inv_gs = qd->inv_global_scale;
  float c[64];
  int quant = raw_qf[bi];
  float thr_y[4] = {0.58f, 0.64f, 0.64f, 0.64f};
  // ---- pass A: quant adjust (Y, X, B; maximum wins, Y's thresholds are kept)
  if (adjust) {
    const int orig = quant;
    int maxq = 0;
#pragma unroll 1
    for (int it = 0; it < 3; ++it) {
      const int ch = it == 0 ? 1 : (it == 1 ? 0 : 2);
      load_dct8(ch == 0 ? X : (ch == 1 ? Y : B), po, fd.pitch, c);
      float thr[4] = {0.58f, 0.64f, 0.64f, 0.64f};
      const float mulc = ch == 0 ? x_qm_mul : (ch == 1 ? 1.0f : b_qm_mul);
      maxq = max(maxq, adjust_quant_block(c, ch, scale, mulc, orig, thr));
      if (ch == 1) { thr_y[0] = thr[0]; thr_y[1] = thr[1]; thr_y[2] = thr[2]; thr_y[3] = thr[3]; }
    }
    quant = maxq;
  } else {
    thr_y[0] = 0.56f; thr_y[1] = thr_y[2] = thr_y[3] = 0.62f;
  }
  // ---- pass B: quantise
  const float qac = scale * (float)quant;
  const float inv_qac = inv_gs / (float)quant;
  const int tx = bx >> 3, ty = by >> 3;
  const float x_factor = 0.0f + (float)cmap[(size_t)ty * fd.txs + tx] / 84.0f;
  const float b_factor = 1.0f + (float)cmap[(size_t)fd.txs * fd.tys + (size_t)ty * fd.txs + tx] / 84.0f;
  const int g = (by >> 5) * fd.gxs + (bx >> 5);
  int16_t* dst = coeffs + ((size_t)g * kGroupBlocks + (size_t)(by & 31) * 32 + (bx & 31)) * 192;
  int q[64];
  int nz, last;
  float dcv[3];
  // Y
  load_dct8(Y, po, fd.pitch, c);
  dcv[1] = c[0];
  quantize_block(c, 1, qac * 1.0f, thr_y, q);
  emit_block(q, dst, nz, last);
  nzeros[nblk + bi] = (uint8_t)nz; nzcount[nblk + bi] = (uint16_t)nz; lastk[nblk + bi] = (uint16_t)last;
  // dequantised Y for the chroma-from-luma term, kept as 64 floats in place of q
  float yrt[64];
#pragma unroll
  for (int i = 0; i < 64; ++i) yrt[i] = (quant_bias_y(q[i]) * c_dqy8[i]) * inv_qac;
  const float thr0[4] = {0.58f, 0.64f, 0.64f, 0.64f};
  // X
  load_dct8(X, po, fd.pitch, c);
  dcv[0] = c[0];
#pragma unroll
  for (int i = 0; i < 64; ++i) c[i] = __fmaf_rn(-x_factor, yrt[i], c[i]);
  quantize_block(c, 0, qac * x_qm_mul, thr0, q);
  emit_block(q, dst + 64, nz, last);
  nzeros[bi] = (uint8_t)nz; nzcount[bi] = (uint16_t)nz; lastk[bi] = (uint16_t)last;
  // B
  load_dct8(B, po, fd.pitch, c);
  dcv[2] = c[0];
#pragma unroll
  for (int i = 0; i < 64; ++i) c[i] = __fmaf_rn(-b_factor, yrt[i], c[i]);
  quantize_block(c, 2, qac * b_qm_mul, thr0, q);
  emit_block(q, dst + 128, nz, last);
  nzeros[2 * nblk + bi] = (uint8_t)nz; nzcount[2 * nblk + bi] = (uint16_t)nz; lastk[2 * nblk + bi] = (uint16_t)last;
  // ---- DC (AddVarDCTDC) + side data
  raw_qf[bi] = quant;
  {
    const int quant_dc = qd->quant_dc;
    const float gsq = scale * (float)quant_dc;
    const float inv_quant_dc = inv_gs / (float)quant_dc;
    const float y_factor = inv_quant_dc * (1.0f / 512.0f);
    const float qy = roundf(dcv[1] * (512.0f * gsq));
    const float qx = roundf((dcv[0] - qy * (y_factor * 0.0f)) * (4096.0f * gsq));
    const float qb = roundf((dcv[2] - qy * (y_factor * 1.0f)) * (256.0f * gsq));
    const int iy = (int)qy, ix = (int)qx, ib = (int)qb;
    dc_quant[0 * nblk + bi] = (int16_t)(ix > 32767 ? 32767 : (ix < -32768 ? -32768 : ix));
    dc_quant[1 * nblk + bi] = (int16_t)(iy > 32767 ? 32767 : (iy < -32768 ? -32768 : iy));
    dc_quant[2 * nblk + bi] = (int16_t)(ib > 32767 ? 32767 : (ib < -32768 ? -32768 : ib));
  }
}

bool dct8_v2_upload_tables(const float* weights192, const float* dequant_y64) {
  return cudaMemcpyToSymbol(c_w8, weights192, 192 * sizeof(float)) == cudaSuccess &&
         cudaMemcpyToSymbol(c_dqy8, dequant_y64, 64 * sizeof(float)) == cudaSuccess;
}

void launch_dct8_quant_v2(const float* x, const float* y, const float* b, const FrameDim& fd, const QuantDev* qd,
                          const int8_t* cmap, float x_qm_mul, float b_qm_mul, int adjust, int32_t* raw_qf, int16_t* coeffs,
                          int16_t* dc_quant, uint8_t* nzeros, uint16_t* nzcount, uint16_t* lastk, cudaStream_t s) {
  ++g_kernel_launches;
  dim3 grid((fd.bxs + kV2Threads - 1) / kV2Threads, fd.bys);
  k_dct8_quant_v2<<<grid, kV2Threads, 0, s>>>(x, y, b, fd, qd, cmap, x_qm_mul, b_qm_mul, adjust, raw_qf, coeffs, dc_quant,
                                             nzeros, nzcount, lastk);
}

}  // namespace jxlb
