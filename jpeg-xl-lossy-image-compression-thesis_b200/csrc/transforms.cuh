// jxlb200 — VarDCT transforms for lane groups (stage U5a; libjxl dct-inl.h / enc_transforms-inl.h /
// dec_transforms-inl.h [UPSTREAM]).  Same operation order as oracle/jxo_dct.cc, so results are
// bit-identical: recursive even/odd 1-D DCT in registers, 2-D = horizontal pass then vertical pass.
//
// A "group" is GS = max(R, C) consecutive lanes of a warp working on one R x C pixel rectangle
// (32 / GS transforms run side by side in a warp).  Pass 1: lane r transforms pixel row r; pass 2:
// lane hf transforms column hf; the coefficient block (H = min(R, C) rows of W = max(R, C), long
// side horizontal) lands in shared memory with a padded row pitch, after which lane y owns
// coefficient row y.  The inverse mirrors it and leaves pixel row r with lane r.
#pragma once
#include "jxl_common.cuh"

namespace jxlb {

constexpr int kTPitch = 33;   // row pitch (floats) of every 2-D scratch buffer: conflict-free column access

template <int N> struct Wc;
template <> struct Wc<4> { static __device__ __forceinline__ float v(int i) { const float t[2] = {5.411961e-01f, 1.306563e+00f}; return t[i]; } };
template <> struct Wc<8> { static __device__ __forceinline__ float v(int i) { const float t[4] = {5.097956e-01f, 6.013449e-01f, 8.999762e-01f, 2.5629156e+00f}; return t[i]; } };
template <> struct Wc<16> { static __device__ __forceinline__ float v(int i) { const float t[8] = {5.024193e-01f, 5.224986e-01f, 5.6694406e-01f, 6.468218e-01f, 7.881546e-01f, 1.0606776e+00f, 1.7224472e+00f, 5.1011486e+00f}; return t[i]; } };
template <> struct Wc<32> { static __device__ __forceinline__ float v(int i) { const float t[16] = {5.00603e-01f, 5.0547093e-01f, 5.154473e-01f, 5.310426e-01f, 5.531039e-01f, 5.82935e-01f, 6.225041e-01f, 6.748083e-01f, 7.445363e-01f, 8.393496e-01f, 9.725682e-01f, 1.1694399e+00f, 1.4841646e+00f, 2.057781e+00f, 3.4076085e+00f, 1.0190008e+01f}; return t[i]; } };

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
    d[0] = d[0] * 1.41421356237309504880f + d[1];
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
      const float m = d[i] * Wc<N>::v(i);
      v[i] = s[i] + m;
      v[N - 1 - i] = s[i] - m;
    }
  }
}

// ---- plain R x C DCT (strategies DCT, DCT16X16, DCT32X32, DCTrXc rectangles) --------------------
// px: pixel tile in shared memory (row pitch px_pitch); t / out: [.][kTPitch] buffers that MAY BE THE SAME
// buffer (every stage reads all of its inputs into registers, synchronises, then writes), so a transform
// needs one scratch buffer only.  gl = lane within the group.  All lanes of the warp must call.
template <int R, int C>
__device__ __forceinline__ void fwd_dct2d(const float* px, int px_pitch, float* t, float* out, int gl) {
  if (gl < R) {
    float v[C];
#pragma unroll
    for (int x = 0; x < C; ++x) v[x] = px[gl * px_pitch + x];
    dct1d<C>(v);
#pragma unroll
    for (int x = 0; x < C; ++x) t[gl * kTPitch + x] = v[x];      // t[r][hf]
  }
  __syncwarp();
  {
    float v[R];
    if (gl < C) {
#pragma unroll
      for (int y = 0; y < R; ++y) v[y] = t[y * kTPitch + gl];
    }
    __syncwarp();   // t and out may alias
    if (gl < C) {
      dct1d<R>(v);
      if constexpr (R >= C) {
#pragma unroll
        for (int y = 0; y < R; ++y) out[gl * kTPitch + y] = v[y];  // out[hf][vf]: coefficient row = hf
      } else {
#pragma unroll
        for (int y = 0; y < R; ++y) out[y * kTPitch + gl] = v[y];  // out[vf][hf]: coefficient row = vf
      }
    }
  }
  __syncwarp();
}

// coef: coefficient block [H][kTPitch]; t: scratch; px: output pixel tile [R][kTPitch]
template <int R, int C>
__device__ __forceinline__ void inv_dct2d(const float* coef, float* t, float* px, int gl) {
  {
    float v[R];
    if (gl < C) {
      if constexpr (R >= C) {
#pragma unroll
        for (int y = 0; y < R; ++y) v[y] = coef[gl * kTPitch + y];
      } else {
#pragma unroll
        for (int y = 0; y < R; ++y) v[y] = coef[y * kTPitch + gl];
      }
    }
    __syncwarp();   // coef and t may alias
    if (gl < C) {
      idct1d<R>(v);
#pragma unroll
      for (int y = 0; y < R; ++y) t[y * kTPitch + gl] = v[y];      // t[y][hf]
    }
  }
  __syncwarp();
  if (gl < R) {
    float v[C];
#pragma unroll
    for (int x = 0; x < C; ++x) v[x] = t[gl * kTPitch + x];
    idct1d<C>(v);
#pragma unroll
    for (int x = 0; x < C; ++x) px[gl * kTPitch + x] = v[x];
  }
  __syncwarp();
}

// ---- 8x8 special strategies (groups of 8 lanes) ---------------------------------------------------
enum { kStratDCT = 0, kStratDCT4X4 = 3, kStratDCT16X16 = 4, kStratDCT32X32 = 5, kStratDCT16X8 = 6, kStratDCT8X16 = 7,
       kStratDCT32X16 = 10, kStratDCT16X32 = 11, kStratDCT4X8 = 12, kStratDCT8X4 = 13 };

// All of them are alias-safe like the plain transform (t / out / px may be one buffer): every stage loads
// into registers, synchronises, then stores.

// DCT4X4: four 4x4 DCTs interleaved, then the 2x2 Hadamard of their DCs (oracle TransformFromPixels)
__device__ __forceinline__ void fwd_dct4x4(const float* px, int px_pitch, float* t, float* out, int gl) {
  float a[4], b[4];
  if (gl < 8) {
#pragma unroll
    for (int x = 0; x < 4; ++x) { a[x] = px[gl * px_pitch + x]; b[x] = px[gl * px_pitch + 4 + x]; }
    dct1d<4>(a); dct1d<4>(b);
#pragma unroll
    for (int x = 0; x < 4; ++x) { t[gl * kTPitch + x] = a[x]; t[gl * kTPitch + 4 + x] = b[x]; }
  }
  __syncwarp();
  if (gl < 8) {
#pragma unroll
    for (int y = 0; y < 4; ++y) { a[y] = t[y * kTPitch + gl]; b[y] = t[(4 + y) * kTPitch + gl]; }
  }
  __syncwarp();
  if (gl < 8) {
    const int x = gl >> 2, hf = gl & 3;     // column gl = quadrant column x, horizontal frequency hf
    dct1d<4>(a); dct1d<4>(b);
    // d[hf*4 + vf] of quadrant (y, x) -> coef[(y + hf*2)*8 + x + vf*2]
#pragma unroll
    for (int vf = 0; vf < 4; ++vf) {
      out[(0 + hf * 2) * kTPitch + x + vf * 2] = a[vf];
      out[(1 + hf * 2) * kTPitch + x + vf * 2] = b[vf];
    }
  }
  __syncwarp();
  if (gl == 0) {
    const float b00 = out[0], b01 = out[1], b10 = out[kTPitch], b11 = out[kTPitch + 1];
    out[0] = (b00 + b01 + b10 + b11) * 0.25f;
    out[1] = (b00 + b01 - b10 - b11) * 0.25f;
    out[kTPitch] = (b00 - b01 + b10 - b11) * 0.25f;
    out[kTPitch + 1] = (b00 - b01 - b10 + b11) * 0.25f;
  }
  __syncwarp();
}

__device__ __forceinline__ void inv_dct4x4(float* coef, float* t, float* px, int gl) {
  // (coef is modified in place: the DC Hadamard is undone first, as the oracle does on its copy)
  if (gl == 0) {
    const float b00 = coef[0], b01 = coef[1], b10 = coef[kTPitch], b11 = coef[kTPitch + 1];
    coef[0] = b00 + b01 + b10 + b11;
    coef[1] = b00 + b01 - b10 - b11;
    coef[kTPitch] = b00 - b01 + b10 - b11;
    coef[kTPitch + 1] = b00 - b01 - b10 + b11;
  }
  __syncwarp();
  float a[4], b[4];
  if (gl < 8) {
    const int x = gl >> 2, hf = gl & 3;
#pragma unroll
    for (int vf = 0; vf < 4; ++vf) { a[vf] = coef[(0 + hf * 2) * kTPitch + x + vf * 2]; b[vf] = coef[(1 + hf * 2) * kTPitch + x + vf * 2]; }
  }
  __syncwarp();
  if (gl < 8) {
    idct1d<4>(a); idct1d<4>(b);
#pragma unroll
    for (int y = 0; y < 4; ++y) { t[y * kTPitch + gl] = a[y]; t[(4 + y) * kTPitch + gl] = b[y]; }
  }
  __syncwarp();
  if (gl < 8) {
#pragma unroll
    for (int x = 0; x < 4; ++x) { a[x] = t[gl * kTPitch + x]; b[x] = t[gl * kTPitch + 4 + x]; }
    idct1d<4>(a); idct1d<4>(b);
#pragma unroll
    for (int x = 0; x < 4; ++x) { px[gl * kTPitch + x] = a[x]; px[gl * kTPitch + 4 + x] = b[x]; }
  }
  __syncwarp();
}

// DCT4X8: two 8-row x 4-col halves side by side; coef[(x + hf*2)*8 + vf]
__device__ __forceinline__ void fwd_dct4x8(const float* px, int px_pitch, float* t, float* out, int gl) {
  if (gl < 8) {
    float a[4], b[4];
#pragma unroll
    for (int x = 0; x < 4; ++x) { a[x] = px[gl * px_pitch + x]; b[x] = px[gl * px_pitch + 4 + x]; }
    dct1d<4>(a); dct1d<4>(b);
#pragma unroll
    for (int x = 0; x < 4; ++x) { t[gl * kTPitch + x] = a[x]; t[gl * kTPitch + 4 + x] = b[x]; }
  }
  __syncwarp();
  float v[8];
  if (gl < 8) {
#pragma unroll
    for (int y = 0; y < 8; ++y) v[y] = t[y * kTPitch + gl];
  }
  __syncwarp();
  if (gl < 8) {
    const int x = gl >> 2, hf = gl & 3;
    dct1d<8>(v);
#pragma unroll
    for (int vf = 0; vf < 8; ++vf) out[(x + hf * 2) * kTPitch + vf] = v[vf];
  }
  __syncwarp();
  if (gl == 0) {
    const float b0 = out[0], b1 = out[kTPitch];
    out[0] = (b0 + b1) * 0.5f;
    out[kTPitch] = (b0 - b1) * 0.5f;
  }
  __syncwarp();
}

__device__ __forceinline__ void inv_dct4x8(float* coef, float* t, float* px, int gl) {
  if (gl == 0) {
    const float b0 = coef[0], b1 = coef[kTPitch];
    coef[0] = b0 + b1; coef[kTPitch] = b0 - b1;
  }
  __syncwarp();
  float v[8];
  if (gl < 8) {
    const int x = gl >> 2, hf = gl & 3;
#pragma unroll
    for (int vf = 0; vf < 8; ++vf) v[vf] = coef[(x + hf * 2) * kTPitch + vf];
  }
  __syncwarp();
  if (gl < 8) {
    idct1d<8>(v);
#pragma unroll
    for (int y = 0; y < 8; ++y) t[y * kTPitch + gl] = v[y];
  }
  __syncwarp();
  if (gl < 8) {
    float a[4], b[4];
#pragma unroll
    for (int x = 0; x < 4; ++x) { a[x] = t[gl * kTPitch + x]; b[x] = t[gl * kTPitch + 4 + x]; }
    idct1d<4>(a); idct1d<4>(b);
#pragma unroll
    for (int x = 0; x < 4; ++x) { px[gl * kTPitch + x] = a[x]; px[gl * kTPitch + 4 + x] = b[x]; }
  }
  __syncwarp();
}

// DCT8X4: two 4-row x 8-col halves stacked; coef[(y + vf*2)*8 + hf]
__device__ __forceinline__ void fwd_dct8x4(const float* px, int px_pitch, float* t, float* out, int gl) {
  if (gl < 8) {
    float v[8];
#pragma unroll
    for (int x = 0; x < 8; ++x) v[x] = px[gl * px_pitch + x];
    dct1d<8>(v);
#pragma unroll
    for (int x = 0; x < 8; ++x) t[gl * kTPitch + x] = v[x];
  }
  __syncwarp();
  float a[4], b[4];
  if (gl < 8) {
#pragma unroll
    for (int y = 0; y < 4; ++y) { a[y] = t[y * kTPitch + gl]; b[y] = t[(4 + y) * kTPitch + gl]; }
  }
  __syncwarp();
  if (gl < 8) {
    dct1d<4>(a); dct1d<4>(b);
#pragma unroll
    for (int vf = 0; vf < 4; ++vf) { out[(0 + vf * 2) * kTPitch + gl] = a[vf]; out[(1 + vf * 2) * kTPitch + gl] = b[vf]; }
  }
  __syncwarp();
  if (gl == 0) {
    const float b0 = out[0], b1 = out[kTPitch];
    out[0] = (b0 + b1) * 0.5f;
    out[kTPitch] = (b0 - b1) * 0.5f;
  }
  __syncwarp();
}

__device__ __forceinline__ void inv_dct8x4(float* coef, float* t, float* px, int gl) {
  if (gl == 0) {
    const float b0 = coef[0], b1 = coef[kTPitch];
    coef[0] = b0 + b1; coef[kTPitch] = b0 - b1;
  }
  __syncwarp();
  float a[4], b[4];
  if (gl < 8) {
#pragma unroll
    for (int vf = 0; vf < 4; ++vf) { a[vf] = coef[(0 + vf * 2) * kTPitch + gl]; b[vf] = coef[(1 + vf * 2) * kTPitch + gl]; }
  }
  __syncwarp();
  if (gl < 8) {
    idct1d<4>(a); idct1d<4>(b);
#pragma unroll
    for (int y = 0; y < 4; ++y) { t[y * kTPitch + gl] = a[y]; t[(4 + y) * kTPitch + gl] = b[y]; }
  }
  __syncwarp();
  if (gl < 8) {
    float v[8];
#pragma unroll
    for (int x = 0; x < 8; ++x) v[x] = t[gl * kTPitch + x];
    idct1d<8>(v);
#pragma unroll
    for (int x = 0; x < 8; ++x) px[gl * kTPitch + x] = v[x];
  }
  __syncwarp();
}

// strategy-dispatched forward / inverse for compile-time (S) strategies
template <int S> struct StratDim;
template <> struct StratDim<kStratDCT> { static constexpr int R = 8, C = 8; };
template <> struct StratDim<kStratDCT4X4> { static constexpr int R = 8, C = 8; };
template <> struct StratDim<kStratDCT4X8> { static constexpr int R = 8, C = 8; };
template <> struct StratDim<kStratDCT8X4> { static constexpr int R = 8, C = 8; };
template <> struct StratDim<kStratDCT16X16> { static constexpr int R = 16, C = 16; };
template <> struct StratDim<kStratDCT32X32> { static constexpr int R = 32, C = 32; };
template <> struct StratDim<kStratDCT16X8> { static constexpr int R = 16, C = 8; };
template <> struct StratDim<kStratDCT8X16> { static constexpr int R = 8, C = 16; };
template <> struct StratDim<kStratDCT32X16> { static constexpr int R = 32, C = 16; };
template <> struct StratDim<kStratDCT16X32> { static constexpr int R = 16, C = 32; };

template <int S>
__device__ __forceinline__ void fwd_transform(const float* px, int px_pitch, float* t, float* out, int gl) {
  if constexpr (S == kStratDCT4X4) fwd_dct4x4(px, px_pitch, t, out, gl);
  else if constexpr (S == kStratDCT4X8) fwd_dct4x8(px, px_pitch, t, out, gl);
  else if constexpr (S == kStratDCT8X4) fwd_dct8x4(px, px_pitch, t, out, gl);
  else fwd_dct2d<StratDim<S>::R, StratDim<S>::C>(px, px_pitch, t, out, gl);
}
template <int S>
__device__ __forceinline__ void inv_transform(float* coef, float* t, float* px, int gl) {
  if constexpr (S == kStratDCT4X4) inv_dct4x4(coef, t, px, gl);
  else if constexpr (S == kStratDCT4X8) inv_dct4x8(coef, t, px, gl);
  else if constexpr (S == kStratDCT8X4) inv_dct8x4(coef, t, px, gl);
  else inv_dct2d<StratDim<S>::R, StratDim<S>::C>(coef, t, px, gl);
}

// xor-butterfly sum over the first `n` lanes of a group (n = 8, 16, 32): the oracle's ButterflySum
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

}  // namespace jxlb
