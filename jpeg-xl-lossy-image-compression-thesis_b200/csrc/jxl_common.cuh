// jxlb200 — shared device/host definitions for the sm_100a kernels.
// Numerics contract (DESIGN.md "Numerics"): IEEE fp32, round-to-nearest, compiled with
// -fmad=false; a fused multiply-add happens exactly where __fmaf_rn()/fmaf() is written.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>

namespace jxlb {

constexpr int kGroupBlocks = 1024;  // 32 x 32 blocks per AC group slot

struct FrameDim {
  int xsize, ysize, xs_pad, ys_pad, pitch, bxs, bys, gxs, gys, num_groups, dgxs, dgys, num_dc_groups, txs, tys;
  __host__ __device__ void Set(int w, int h) {
    xsize = w; ysize = h;
    xs_pad = (w + 7) / 8 * 8; ys_pad = (h + 7) / 8 * 8;
    pitch = (xs_pad + 31) / 32 * 32;
    bxs = xs_pad / 8; bys = ys_pad / 8;
    gxs = (bxs + 31) / 32; gys = (bys + 31) / 32; num_groups = gxs * gys;
    dgxs = (bxs + 255) / 256; dgys = (bys + 255) / 256; num_dc_groups = dgxs * dgys;
    txs = (bxs + 7) / 8; tys = (bys + 7) / 8;
  }
};

// quantiser state living in device memory (written by k_quant_params, read by later kernels)
struct QuantDev {
  int global_scale;
  int quant_dc;
  float scale;             // global_scale / 65536
  float inv_global_scale;  // 65536 / global_scale
  float median, mad;
};

__device__ __forceinline__ float fast_log2f(float x) {
  const float p0 = -1.8503833400518310E-06f, p1 = 1.4287160470083755E+00f, p2 = 7.4245873327820566E-01f;
  const float q0 = 9.9032814277590719E-01f, q1 = 1.0096718572241148E+00f, q2 = 1.7409343003366853E-01f;
  const int x_bits = __float_as_int(x);
  const int exp_bits = x_bits - 0x3f2aaaab;
  const int exp_shifted = exp_bits >> 23;
  const float mantissa = __int_as_float(x_bits - (exp_shifted << 23));
  const float exp_val = (float)exp_shifted;
  const float m = mantissa - 1.0f;
  const float yp = __fmaf_rn(__fmaf_rn(p2, m, p1), m, p0);
  const float yq = __fmaf_rn(__fmaf_rn(q2, m, q1), m, q0);
  return yp / yq + exp_val;
}

__device__ __forceinline__ float fast_pow2f(float x) {
  const float floorx = floorf(x);
  const float e = __int_as_float(((int)floorx + 127) << 23);
  const float frac = x - floorx;
  float num = frac + 1.01749063e+01f;
  num = __fmaf_rn(num, frac, 4.88687798e+01f);
  num = __fmaf_rn(num, frac, 9.85506591e+01f);
  num = num * e;
  float den = __fmaf_rn(frac, 2.10242958e-01f, -2.22328856e-02f);
  den = __fmaf_rn(den, frac, -1.94414990e+01f);
  den = __fmaf_rn(den, frac, 9.85506633e+01f);
  return num / den;
}

// cube root for x >= 0 (bit-hack estimate of x^(-1/3), three Newton steps, x * r^2)
__device__ __forceinline__ float cbrt_pos(float x) {
  // (branch-free: x <= 0 or NaN is selected to 0 at the end; the early return cost a divergence bracket per call)
  float r = __uint_as_float(0x54A21D2Au - __float_as_uint(x) / 3u);
  const float x3 = x * (1.0f / 3.0f);
  const float k43 = 4.0f / 3.0f;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const float r2 = r * r;
    const float r4 = r2 * r2;
    r = __fmaf_rn(-x3, r4, k43 * r);
  }
  const float y = (r * r) * x;
  return x > 0.0f ? y : 0.0f;
}

// ---- 1-D scaled DCT-II on registers, same operation order as oracle/jxo_dct.cc ------------
// (recursive even/odd split; literal tables printed by tools/gen_tables.py)
__device__ __forceinline__ void dct2_raw(float& a, float& b) { const float s = a + b, d = a - b; a = s; b = d; }

__device__ __forceinline__ void dct4_raw(float v[4]) {
  float s0 = v[0] + v[3], s1 = v[1] + v[2];
  float d0 = v[0] - v[3], d1 = v[1] - v[2];
  d0 = d0 * 5.411961e-01f; d1 = d1 * 1.306563e+00f;
  dct2_raw(s0, s1);
  dct2_raw(d0, d1);
  d0 = d0 * 1.41421356237309504880f + d1;
  v[0] = s0; v[1] = d0; v[2] = s1; v[3] = d1;
}

__device__ __forceinline__ void dct8_raw(float v[8]) {
  float s[4], d[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) { s[i] = v[i] + v[7 - i]; d[i] = v[i] - v[7 - i]; }
  d[0] = d[0] * 5.097956e-01f; d[1] = d[1] * 6.013449e-01f; d[2] = d[2] * 8.999762e-01f; d[3] = d[3] * 2.5629156e+00f;
  dct4_raw(s);
  dct4_raw(d);
  d[0] = d[0] * 1.41421356237309504880f + d[1];
  d[1] = d[1] + d[2];
  d[2] = d[2] + d[3];
#pragma unroll
  for (int i = 0; i < 4; ++i) { v[2 * i] = s[i]; v[2 * i + 1] = d[i]; }
}

__device__ __forceinline__ void dct8_scaled(float v[8]) {
  dct8_raw(v);
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = v[i] * 0.125f;
}

// transpose an 8x8 tile held as one row of 8 registers per lane across 8 consecutive lanes
__device__ __forceinline__ void transpose8(float a[8], int lane8) {
  const unsigned full = 0xffffffffu;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const bool hi = lane8 & 4;
    const float send = hi ? a[i] : a[i + 4];
    const float recv = __shfl_xor_sync(full, send, 4);
    if (hi) a[i] = recv; else a[i + 4] = recv;
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int i = (k & 1) + ((k >> 1) << 2);  // 0,1,4,5
    const bool hi = lane8 & 2;
    const float send = hi ? a[i] : a[i + 2];
    const float recv = __shfl_xor_sync(full, send, 2);
    if (hi) a[i] = recv; else a[i + 2] = recv;
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int i = 2 * k;  // 0,2,4,6
    const bool hi = lane8 & 1;
    const float send = hi ? a[i] : a[i + 1];
    const float recv = __shfl_xor_sync(full, send, 1);
    if (hi) a[i] = recv; else a[i + 1] = recv;
  }
}

// butterfly sums over the 8 lanes of a block group (strides 4, 2, 1) — the association the
// oracle's halving tree uses
__device__ __forceinline__ float tree8_sum(float v) {
  v = v + __shfl_xor_sync(0xffffffffu, v, 4);
  v = v + __shfl_xor_sync(0xffffffffu, v, 2);
  v = v + __shfl_xor_sync(0xffffffffu, v, 1);
  return v;
}
__device__ __forceinline__ float tree8_max(float v) {
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 4));
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
  return v;
}
__device__ __forceinline__ int tree8_isum(int v) {
  v += __shfl_xor_sync(0xffffffffu, v, 4);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v;
}
__device__ __forceinline__ int tree8_imax(int v) {
  v = max(v, __shfl_xor_sync(0xffffffffu, v, 4));
  v = max(v, __shfl_xor_sync(0xffffffffu, v, 2));
  v = max(v, __shfl_xor_sync(0xffffffffu, v, 1));
  return v;
}

}  // namespace jxlb
