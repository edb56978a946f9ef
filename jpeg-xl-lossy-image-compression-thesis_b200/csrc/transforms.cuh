// jxlb200 — VarDCT transforms (stage U5a; libjxl dct-inl.h / enc_transforms-inl.h / dec_transforms-inl.h
// [UPSTREAM]).  Same operation order as oracle/jxo_dct.cc, so results are bit-identical: recursive even/odd
// 1-D DCT in registers, 2-D = horizontal pass then vertical pass (inverse: vertical, then horizontal).
// Numerics contract: the recombination multiply-adds are fused (__fmaf_rn), everything else is a separate
// IEEE operation (-fmad=false).
//
// Two execution shapes:
//  * 8x8 strategies (DCT, DCT4X4, DCT2X2, DCT4X8, DCT8X4, IDENTITY): ONE THREAD holds the 64 samples in
//    registers; the transposition between the passes is register renaming (fwd8x8 / inv8x8).
//  * 16 / 32 / 64-sized strategies: a LANE GROUP of N lanes works on an N x N pixel square that holds one
//    square transform, two tall halves (left / right) or two wide halves (top / bottom): lane r transforms
//    pixel row r, the rows go through an XOR-swizzled shared-memory square (16-byte row accesses and scalar
//    column accesses are both conflict-free), lane hf transforms column hf and ends up owning every vertical
//    frequency of its horizontal frequency — the quantisation / error / inverse-vertical steps of the callers
//    stay in registers (SquareXform).
#pragma once
#include "jxl_common.cuh"

namespace jxlb {

enum { kStratDCT = 0, kStratIDENTITY = 1, kStratDCT2X2 = 2, kStratDCT4X4 = 3, kStratDCT16X16 = 4, kStratDCT32X32 = 5,
       kStratDCT16X8 = 6, kStratDCT8X16 = 7, kStratDCT32X8 = 8, kStratDCT8X32 = 9, kStratDCT32X16 = 10, kStratDCT16X32 = 11,
       kStratDCT4X8 = 12, kStratDCT8X4 = 13, kStratDCT64X64 = 18, kStratDCT64X32 = 19, kStratDCT32X64 = 20 };

// 1 / (2 cos((i + 1/2) pi / N)) — literal tables printed by tools/gen_tables.py
template <int N> struct Wc;
template <> struct Wc<4> { static __device__ __forceinline__ float v(int i) { const float t[2] = {5.411961e-01f, 1.306563e+00f}; return t[i]; } };
template <> struct Wc<8> { static __device__ __forceinline__ float v(int i) { const float t[4] = {5.097956e-01f, 6.013449e-01f, 8.999762e-01f, 2.5629156e+00f}; return t[i]; } };
template <> struct Wc<16> { static __device__ __forceinline__ float v(int i) { const float t[8] = {5.024193e-01f, 5.224986e-01f, 5.6694406e-01f, 6.468218e-01f, 7.881546e-01f, 1.0606776e+00f, 1.7224472e+00f, 5.1011486e+00f}; return t[i]; } };
template <> struct Wc<32> { static __device__ __forceinline__ float v(int i) { const float t[16] = {5.00603e-01f, 5.0547093e-01f, 5.154473e-01f, 5.310426e-01f, 5.531039e-01f, 5.82935e-01f, 6.225041e-01f, 6.748083e-01f, 7.445363e-01f, 8.393496e-01f, 9.725682e-01f, 1.1694399e+00f, 1.4841646e+00f, 2.057781e+00f, 3.4076085e+00f, 1.0190008e+01f}; return t[i]; } };
template <> struct Wc<64> { static __device__ __forceinline__ float v(int i) { const float t[32] = {5.001506e-01f, 5.0135845e-01f, 5.037887e-01f, 5.0747114e-01f, 5.1245147e-01f, 5.187927e-01f, 5.265773e-01f, 5.3590983e-01f, 5.469204e-01f, 5.597698e-01f, 5.746552e-01f, 5.918185e-01f, 6.1155736e-01f, 6.3423896e-01f, 6.603198e-01f, 6.903721e-01f, 7.2512054e-01f, 7.6549417e-01f, 8.127021e-01f, 8.683447e-01f, 9.345836e-01f, 1.0144082e+00f, 1.1120716e+00f, 1.2338327e+00f, 1.3892939e+00f, 1.5939723e+00f, 1.874676e+00f, 2.2820501e+00f, 2.9246285e+00f, 4.084611e+00f, 6.7967505e+00f, 2.0373878e+01f}; return t[i]; } };

// unscaled forward DCT, oracle DctRec
template <int N> __device__ __forceinline__ void dct_rec(float* v) {
  if constexpr (N == 1) { return; }
  else if constexpr (N == 2) { const float a = v[0] + v[1], b = v[0] - v[1]; v[0] = a; v[1] = b; }
  else {
    constexpr int h = N / 2;
    float s[h], d[h];
#pragma unroll
    for (int i = 0; i < h; ++i) { s[i] = v[i] + v[N - 1 - i]; d[i] = v[i] - v[N - 1 - i]; }
#pragma unroll
    for (int i = 0; i < h; ++i) d[i] = d[i] * Wc<N>::v(i);
    dct_rec<h>(s);
    dct_rec<h>(d);
    d[0] = __fmaf_rn(d[0], 1.41421356237309504880f, d[1]);
#pragma unroll
    for (int i = 1; i + 1 < h; ++i) d[i] = d[i] + d[i + 1];
#pragma unroll
    for (int i = 0; i < h; ++i) { v[2 * i] = s[i]; v[2 * i + 1] = d[i]; }
  }
}
// oracle Dct1D: forward + 1/N scale
template <int N> __device__ __forceinline__ void dct1d(float* v) {
  dct_rec<N>(v);
#pragma unroll
  for (int i = 0; i < N; ++i) v[i] = v[i] * (1.0f / (float)N);
}
// oracle IdctRec / Idct1D
template <int N> __device__ __forceinline__ void idct1d(float* v) {
  if constexpr (N == 1) { return; }
  else if constexpr (N == 2) { const float a = v[0] + v[1], b = v[0] - v[1]; v[0] = a; v[1] = b; }
  else {
    constexpr int h = N / 2;
    float s[h], d[h];
#pragma unroll
    for (int i = 0; i < h; ++i) { s[i] = v[2 * i]; d[i] = v[2 * i + 1]; }
    idct1d<h>(s);
#pragma unroll
    for (int i = h - 1; i >= 1; --i) d[i] = d[i] + d[i - 1];
    d[0] = d[0] * 1.41421356237309504880f;
    idct1d<h>(d);
#pragma unroll
    for (int i = 0; i < h; ++i) {
      const float w = Wc<N>::v(i);
      v[i] = __fmaf_rn(d[i], w, s[i]);
      v[N - 1 - i] = __fmaf_rn(-d[i], w, s[i]);
    }
  }
}

// xor-butterfly sum over N consecutive lanes (N <= 32): the oracle's ButterflySum
template <int N> __device__ __forceinline__ float group_sum(float v) {
#pragma unroll
  for (int st = N / 2; st >= 1; st >>= 1) v = v + __shfl_xor_sync(0xffffffffu, v, st);
  return v;
}
template <int N> __device__ __forceinline__ int group_isum(int v) {
#pragma unroll
  for (int st = N / 2; st >= 1; st >>= 1) v += __shfl_xor_sync(0xffffffffu, v, st);
  return v;
}
template <int N> __device__ __forceinline__ int group_imax(int v) {
#pragma unroll
  for (int st = N / 2; st >= 1; st >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, st));
  return v;
}
template <int N> __device__ __forceinline__ float group_fmax(float v) {
#pragma unroll
  for (int st = N / 2; st >= 1; st >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, st));
  return v;
}
// the same association inside one thread: p[0] of the butterfly over 8 values
__device__ __forceinline__ float tree8(const float* p) {
  return ((p[0] + p[4]) + (p[2] + p[6])) + ((p[1] + p[5]) + (p[3] + p[7]));
}

// ====================================================================================================
// 8x8 strategies, one thread per block.  p: pixels row-major; c: coefficients in the strategy's storage
// layout (oracle TransformFromPixels / TransformToPixels).
// ====================================================================================================
__device__ __forceinline__ void hadamard4_fwd(float& a, float& b, float& c, float& d) {
  const float b00 = a, b01 = b, b10 = c, b11 = d;
  a = (b00 + b01 + b10 + b11) * 0.25f;
  b = (b00 + b01 - b10 - b11) * 0.25f;
  c = (b00 - b01 + b10 - b11) * 0.25f;
  d = (b00 - b01 - b10 + b11) * 0.25f;
}
__device__ __forceinline__ void hadamard4_inv(float& a, float& b, float& c, float& d) {
  const float b00 = a, b01 = b, b10 = c, b11 = d;
  a = b00 + b01 + b10 + b11;
  b = b00 + b01 - b10 - b11;
  c = b00 - b01 + b10 - b11;
  d = b00 - b01 - b10 + b11;
}

// libjxl DCT2TopBlock<S> / IDCT2TopBlock<S> on an 8-pitch block (in place through a copy of the SxS corner)
template <int S> __device__ __forceinline__ void dct2_top(float* b) {
  constexpr int h = S / 2;
  float t[S * S];
#pragma unroll
  for (int y = 0; y < h; ++y)
#pragma unroll
    for (int x = 0; x < h; ++x) {
      const float c00 = b[y * 2 * 8 + x * 2], c01 = b[y * 2 * 8 + x * 2 + 1];
      const float c10 = b[(y * 2 + 1) * 8 + x * 2], c11 = b[(y * 2 + 1) * 8 + x * 2 + 1];
      t[y * S + x] = (c00 + c01 + c10 + c11) * 0.25f;
      t[y * S + h + x] = (c00 + c01 - c10 - c11) * 0.25f;
      t[(y + h) * S + x] = (c00 - c01 + c10 - c11) * 0.25f;
      t[(y + h) * S + h + x] = (c00 - c01 - c10 + c11) * 0.25f;
    }
#pragma unroll
  for (int y = 0; y < S; ++y)
#pragma unroll
    for (int x = 0; x < S; ++x) b[y * 8 + x] = t[y * S + x];
}
template <int S> __device__ __forceinline__ void idct2_top(float* b) {
  constexpr int h = S / 2;
  float t[S * S];
#pragma unroll
  for (int y = 0; y < h; ++y)
#pragma unroll
    for (int x = 0; x < h; ++x) {
      const float c00 = b[y * 8 + x], c01 = b[y * 8 + h + x], c10 = b[(y + h) * 8 + x], c11 = b[(y + h) * 8 + h + x];
      t[y * 2 * S + x * 2] = c00 + c01 + c10 + c11;
      t[y * 2 * S + x * 2 + 1] = c00 + c01 - c10 - c11;
      t[(y * 2 + 1) * S + x * 2] = c00 - c01 + c10 - c11;
      t[(y * 2 + 1) * S + x * 2 + 1] = c00 - c01 - c10 + c11;
    }
#pragma unroll
  for (int y = 0; y < S; ++y)
#pragma unroll
    for (int x = 0; x < S; ++x) b[y * 8 + x] = t[y * S + x];
}

template <int S> __device__ __forceinline__ void fwd8x8(const float* p, float* c) {
  if constexpr (S == kStratDCT) {
    float t[64];
#pragma unroll
    for (int y = 0; y < 8; ++y) {
      float v[8];
#pragma unroll
      for (int x = 0; x < 8; ++x) v[x] = p[y * 8 + x];
      dct1d<8>(v);
#pragma unroll
      for (int x = 0; x < 8; ++x) t[y * 8 + x] = v[x];
    }
#pragma unroll
    for (int hf = 0; hf < 8; ++hf) {
      float u[8];
#pragma unroll
      for (int y = 0; y < 8; ++y) u[y] = t[y * 8 + hf];
      dct1d<8>(u);
#pragma unroll
      for (int vf = 0; vf < 8; ++vf) c[hf * 8 + vf] = u[vf];
    }
  } else if constexpr (S == kStratDCT4X4 || S == kStratDCT4X8) {
    float t[64];
#pragma unroll
    for (int y = 0; y < 8; ++y) {
      float a[4], b[4];
#pragma unroll
      for (int x = 0; x < 4; ++x) { a[x] = p[y * 8 + x]; b[x] = p[y * 8 + 4 + x]; }
      dct1d<4>(a); dct1d<4>(b);
#pragma unroll
      for (int x = 0; x < 4; ++x) { t[y * 8 + x] = a[x]; t[y * 8 + 4 + x] = b[x]; }
    }
#pragma unroll
    for (int col = 0; col < 8; ++col) {
      const int q = col >> 2, hf = col & 3;
      if constexpr (S == kStratDCT4X4) {
#pragma unroll
        for (int qy = 0; qy < 2; ++qy) {
          float u[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) u[i] = t[(qy * 4 + i) * 8 + col];
          dct1d<4>(u);
#pragma unroll
          for (int vf = 0; vf < 4; ++vf) c[(qy + hf * 2) * 8 + q + vf * 2] = u[vf];
        }
      } else {
        float u[8];
#pragma unroll
        for (int y = 0; y < 8; ++y) u[y] = t[y * 8 + col];
        dct1d<8>(u);
#pragma unroll
        for (int vf = 0; vf < 8; ++vf) c[(q + hf * 2) * 8 + vf] = u[vf];
      }
    }
    if constexpr (S == kStratDCT4X4) {
      hadamard4_fwd(c[0], c[1], c[8], c[9]);
    } else {
      const float b0 = c[0], b1 = c[8];
      c[0] = (b0 + b1) * 0.5f; c[8] = (b0 - b1) * 0.5f;
    }
  } else if constexpr (S == kStratDCT8X4) {
    float t[64];
#pragma unroll
    for (int y = 0; y < 8; ++y) {
      float v[8];
#pragma unroll
      for (int x = 0; x < 8; ++x) v[x] = p[y * 8 + x];
      dct1d<8>(v);
#pragma unroll
      for (int x = 0; x < 8; ++x) t[y * 8 + x] = v[x];
    }
#pragma unroll
    for (int hf = 0; hf < 8; ++hf)
#pragma unroll
      for (int hy = 0; hy < 2; ++hy) {
        float u[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) u[i] = t[(hy * 4 + i) * 8 + hf];
        dct1d<4>(u);
#pragma unroll
        for (int vf = 0; vf < 4; ++vf) c[(hy + vf * 2) * 8 + hf] = u[vf];
      }
    const float b0 = c[0], b1 = c[8];
    c[0] = (b0 + b1) * 0.5f; c[8] = (b0 - b1) * 0.5f;
  } else if constexpr (S == kStratDCT2X2) {
#pragma unroll
    for (int i = 0; i < 64; ++i) c[i] = p[i];
    dct2_top<8>(c); dct2_top<4>(c); dct2_top<2>(c);
  } else {   // IDENTITY
#pragma unroll
    for (int y = 0; y < 2; ++y)
#pragma unroll
      for (int x = 0; x < 2; ++x) {
        float block_dc = 0.0f;
#pragma unroll
        for (int iy = 0; iy < 4; ++iy)
#pragma unroll
          for (int ix = 0; ix < 4; ++ix) block_dc = block_dc + p[(y * 4 + iy) * 8 + x * 4 + ix];
        block_dc = block_dc * (1.0f / 16);
        const float mid = p[(y * 4 + 1) * 8 + x * 4 + 1];
#pragma unroll
        for (int iy = 0; iy < 4; ++iy)
#pragma unroll
          for (int ix = 0; ix < 4; ++ix) {
            if (ix == 1 && iy == 1) continue;
            c[(y + iy * 2) * 8 + x + ix * 2] = p[(y * 4 + iy) * 8 + x * 4 + ix] - mid;
          }
        c[(y + 2) * 8 + x + 2] = c[y * 8 + x];
        c[y * 8 + x] = block_dc;
      }
    hadamard4_fwd(c[0], c[1], c[8], c[9]);
  }
}

// c may be modified (the DC Hadamard of the split strategies is undone in place, as the oracle does on its copy)
template <int S> __device__ __forceinline__ void inv8x8(float* c, float* p) {
  if constexpr (S == kStratDCT) {
    float t[64];
#pragma unroll
    for (int hf = 0; hf < 8; ++hf) {
      float u[8];
#pragma unroll
      for (int vf = 0; vf < 8; ++vf) u[vf] = c[hf * 8 + vf];
      idct1d<8>(u);
#pragma unroll
      for (int y = 0; y < 8; ++y) t[y * 8 + hf] = u[y];
    }
#pragma unroll
    for (int y = 0; y < 8; ++y) {
      float v[8];
#pragma unroll
      for (int x = 0; x < 8; ++x) v[x] = t[y * 8 + x];
      idct1d<8>(v);
#pragma unroll
      for (int x = 0; x < 8; ++x) p[y * 8 + x] = v[x];
    }
  } else if constexpr (S == kStratDCT4X4 || S == kStratDCT4X8) {
    if constexpr (S == kStratDCT4X4) {
      hadamard4_inv(c[0], c[1], c[8], c[9]);
    } else {
      const float b0 = c[0], b1 = c[8];
      c[0] = b0 + b1; c[8] = b0 - b1;
    }
    float t[64];
#pragma unroll
    for (int col = 0; col < 8; ++col) {
      const int q = col >> 2, hf = col & 3;
      if constexpr (S == kStratDCT4X4) {
#pragma unroll
        for (int qy = 0; qy < 2; ++qy) {
          float u[4];
#pragma unroll
          for (int vf = 0; vf < 4; ++vf) u[vf] = c[(qy + hf * 2) * 8 + q + vf * 2];
          idct1d<4>(u);
#pragma unroll
          for (int i = 0; i < 4; ++i) t[(qy * 4 + i) * 8 + col] = u[i];
        }
      } else {
        float u[8];
#pragma unroll
        for (int vf = 0; vf < 8; ++vf) u[vf] = c[(q + hf * 2) * 8 + vf];
        idct1d<8>(u);
#pragma unroll
        for (int y = 0; y < 8; ++y) t[y * 8 + col] = u[y];
      }
    }
#pragma unroll
    for (int y = 0; y < 8; ++y) {
      float a[4], b[4];
#pragma unroll
      for (int x = 0; x < 4; ++x) { a[x] = t[y * 8 + x]; b[x] = t[y * 8 + 4 + x]; }
      idct1d<4>(a); idct1d<4>(b);
#pragma unroll
      for (int x = 0; x < 4; ++x) { p[y * 8 + x] = a[x]; p[y * 8 + 4 + x] = b[x]; }
    }
  } else if constexpr (S == kStratDCT8X4) {
    const float b0 = c[0], b1 = c[8];
    c[0] = b0 + b1; c[8] = b0 - b1;
    float t[64];
#pragma unroll
    for (int hf = 0; hf < 8; ++hf)
#pragma unroll
      for (int hy = 0; hy < 2; ++hy) {
        float u[4];
#pragma unroll
        for (int vf = 0; vf < 4; ++vf) u[vf] = c[(hy + vf * 2) * 8 + hf];
        idct1d<4>(u);
#pragma unroll
        for (int i = 0; i < 4; ++i) t[(hy * 4 + i) * 8 + hf] = u[i];
      }
#pragma unroll
    for (int y = 0; y < 8; ++y) {
      float v[8];
#pragma unroll
      for (int x = 0; x < 8; ++x) v[x] = t[y * 8 + x];
      idct1d<8>(v);
#pragma unroll
      for (int x = 0; x < 8; ++x) p[y * 8 + x] = v[x];
    }
  } else if constexpr (S == kStratDCT2X2) {
    idct2_top<2>(c); idct2_top<4>(c); idct2_top<8>(c);
#pragma unroll
    for (int i = 0; i < 64; ++i) p[i] = c[i];
  } else {   // IDENTITY
    float d0 = c[0], d1 = c[1], d2 = c[8], d3 = c[9];
    hadamard4_inv(d0, d1, d2, d3);
    const float dcs[4] = {d0, d1, d2, d3};
#pragma unroll
    for (int y = 0; y < 2; ++y)
#pragma unroll
      for (int x = 0; x < 2; ++x) {
        float residual_sum = 0.0f;
#pragma unroll
        for (int iy = 0; iy < 4; ++iy)
#pragma unroll
          for (int ix = 0; ix < 4; ++ix) {
            if (ix == 0 && iy == 0) continue;
            residual_sum = residual_sum + c[(y + iy * 2) * 8 + x + ix * 2];
          }
        const float mid = dcs[y * 2 + x] - residual_sum * (1.0f / 16);
#pragma unroll
        for (int iy = 0; iy < 4; ++iy)
#pragma unroll
          for (int ix = 0; ix < 4; ++ix) {
            if (ix == 1 && iy == 1) continue;
            p[(y * 4 + iy) * 8 + x * 4 + ix] = c[(y + iy * 2) * 8 + x + ix * 2] + mid;
          }
        p[(4 * y + 1) * 8 + 4 * x + 1] = mid;
        p[y * 4 * 8 + x * 4] = c[(y + 2) * 8 + x + 2] + mid;
      }
  }
}

// 16-byte asynchronous copy global -> shared (LDGSTS); !valid zero-fills the destination (source size 0)
__device__ __forceinline__ void cp_async16(float* smem_dst, const float* gsrc, bool valid) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int PENDING> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(PENDING) : "memory"); }

// ====================================================================================================
// N x N squares, N lanes.  MODE: one square transform, two tall halves (left | right, N rows x N/2 columns
// each) or two wide halves (top / bottom, N/2 rows x N columns each).
// ====================================================================================================
// kModeTall4 / kModeWide4: ONE N x N/4 (resp. N/4 x N) transform in the first quarter of the square (DCT32X8 / DCT8X32,
// which no search proposes: they are only reached through a caller-supplied strategy map); the rest of the square is zero
enum { kModeSq = 0, kModeTall2 = 1, kModeWide2 = 2, kModeTall4 = 3, kModeWide4 = 4 };

// XOR swizzle of the 16-byte chunks of a row: quarter-warps of row accesses and whole-warp column accesses hit
// 32 distinct banks (N = 16: two rows share a 128-byte line, so the row index is halved first)
template <int N> __device__ __forceinline__ constexpr int swz(int r) { return N == 16 ? ((r >> 1) & 3) : (r & 7); }

template <int N> struct SquareXform {
  static_assert(N == 16 || N == 32 || N == 64, "square sizes");
  static constexpr int kFloats = N * N;
  // group barrier: N <= 32 lanes live in one warp; N = 64 is two warps meeting at named barrier `bar_id`
  static __device__ __forceinline__ void sync(int bar_id) {
    if constexpr (N == 64) asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory");
    else __syncwarp();
  }
  // column addressing of lane l: element (y, l) lives at y * N + col_off[swz(y)]
  struct Col { int off[8]; };
  static __device__ __forceinline__ Col col_of(int l) {
    Col c;
    const int cq = l >> 2, cr = l & 3;
#pragma unroll
    for (int s = 0; s < 8; ++s) c.off[s] = ((cq ^ s) << 2) | cr;
    return c;
  }
  // lane r stores its row v[0..N)
  static __device__ __forceinline__ void store_row(float* t, int r, const float* v) {
    const int s = swz<N>(r);
#pragma unroll
    for (int j = 0; j < N / 4; ++j)
      *reinterpret_cast<float4*>(t + r * N + ((j ^ s) << 2)) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
  }
  static __device__ __forceinline__ void load_row(const float* t, int r, float* v) {
    const int s = swz<N>(r);
#pragma unroll
    for (int j = 0; j < N / 4; ++j) {
      const float4 q = *reinterpret_cast<const float4*>(t + r * N + ((j ^ s) << 2));
      v[4 * j] = q.x; v[4 * j + 1] = q.y; v[4 * j + 2] = q.z; v[4 * j + 3] = q.w;
    }
  }
  static __device__ __forceinline__ void load_col(const float* t, const Col& c, float* u) {
#pragma unroll
    for (int y = 0; y < N; ++y) u[y] = t[y * N + c.off[swz<N>(y)]];
  }
  static __device__ __forceinline__ void store_col(float* t, const Col& c, const float* u) {
#pragma unroll
    for (int y = 0; y < N; ++y) t[y * N + c.off[swz<N>(y)]] = u[y];
  }

  // forward: v = lane's pixel row in, u = lane's coefficient column out (u[vf] of horizontal frequency l;
  // wide halves: u[0..N/2) belongs to the top transform, u[N/2..N) to the bottom one)
  template <int MODE>
  static __device__ __forceinline__ void forward(float* t, int l, const Col& c, float* v, float* u, int bar_id) {
    if constexpr (MODE == kModeTall2) { dct1d<N / 2>(v); dct1d<N / 2>(v + N / 2); }
    else if constexpr (MODE == kModeTall4) dct1d<N / 4>(v);
    else dct1d<N>(v);
    store_row(t, l, v);
    sync(bar_id);
    load_col(t, c, u);
    if constexpr (MODE == kModeWide2) { dct1d<N / 2>(u); dct1d<N / 2>(u + N / 2); }
    else if constexpr (MODE == kModeWide4) dct1d<N / 4>(u);
    else dct1d<N>(u);
  }
  // inverse: u = lane's coefficient column in (destroyed), v = lane's pixel row out
  template <int MODE>
  static __device__ __forceinline__ void inverse(float* t, int l, const Col& c, float* u, float* v, int bar_id) {
    if constexpr (MODE == kModeWide2) { idct1d<N / 2>(u); idct1d<N / 2>(u + N / 2); }
    else if constexpr (MODE == kModeWide4) idct1d<N / 4>(u);
    else idct1d<N>(u);
    store_col(t, c, u);
    sync(bar_id);
    load_row(t, l, v);
    if constexpr (MODE == kModeTall2) { idct1d<N / 2>(v); idct1d<N / 2>(v + N / 2); }
    else if constexpr (MODE == kModeTall4) idct1d<N / 4>(v);
    else idct1d<N>(v);
  }
};

// resample scale between an N-point DCT's low frequencies and the N/8-point DCT of the block means
// (oracle ResampleScale)
__device__ __forceinline__ float resample_scale(int n_from, int k) {
  if (n_from == 8) return 1.0f;
  if (n_from == 16) return k == 0 ? 1.e+00f : 9.017642e-01f;
  if (n_from == 32) { const float t[4] = {1.e+00f, 9.7488683e-01f, 9.017642e-01f, 7.870549e-01f}; return t[k]; }
  const float t[8] = {1.e+00f, 9.936866e-01f, 9.7488683e-01f, 9.4401807e-01f, 9.017642e-01f, 8.490575e-01f, 7.870549e-01f, 7.1710813e-01f};
  return t[k];
}

}  // namespace jxlb
