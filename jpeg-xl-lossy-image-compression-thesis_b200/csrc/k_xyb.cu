// K1 — RGB8 interleaved -> planar XYB f32 (stage U1; libjxl enc_xyb.cc [UPSTREAM]).
// HBM-bound: 3 B/px in, 12 B/px out.  One thread converts 4 consecutive pixels of one
// row and writes one float4 per plane (fully coalesced 128-bit stores); the 12 source
// bytes are fetched as three aligned 32-bit words when the row allows it.  The sRGB EOTF
// is a 256-entry table staged in shared memory (computed once on the host with the same
// rational polynomial as the oracle).
#include "jxl_common.cuh"
#include "kernels.h"

namespace jxlb {

__global__ void __launch_bounds__(128) k_rgb8_to_xyb(const uint8_t* __restrict__ rgb, size_t stride, int w, int h,
                                                     FrameDim fd, const float* __restrict__ lut_g,
                                                     float* __restrict__ px, float* __restrict__ py,
                                                     float* __restrict__ pb) {
  __shared__ float lut[256];
  for (int i = threadIdx.x; i < 256; i += blockDim.x) lut[i] = lut_g[i];
  __syncthreads();
  const int x4 = blockIdx.x * blockDim.x + threadIdx.x;  // group of 4 pixels
  const int y = blockIdx.y;
  if (x4 * 4 >= fd.pitch) return;
  const int sy = y < h ? y : h - 1;
  const uint8_t* row = rgb + (size_t)sy * stride;
  uint8_t v[12];
  const int x0 = x4 * 4;
  if (x0 + 3 < w && ((reinterpret_cast<uintptr_t>(row) & 3) == 0)) {
    const uint32_t* r32 = reinterpret_cast<const uint32_t*>(row) + x4 * 3;
    const uint32_t a = __ldg(r32), b = __ldg(r32 + 1), c = __ldg(r32 + 2);
    v[0] = a & 255; v[1] = (a >> 8) & 255; v[2] = (a >> 16) & 255; v[3] = a >> 24;
    v[4] = b & 255; v[5] = (b >> 8) & 255; v[6] = (b >> 16) & 255; v[7] = b >> 24;
    v[8] = c & 255; v[9] = (c >> 8) & 255; v[10] = (c >> 16) & 255; v[11] = c >> 24;
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int sx = x0 + i;
      sx = sx < w ? sx : w - 1;
      v[3 * i + 0] = __ldg(row + 3 * sx + 0);
      v[3 * i + 1] = __ldg(row + 3 * sx + 1);
      v[3 * i + 2] = __ldg(row + 3 * sx + 2);
    }
  }
  const float kBias = 0.0037930732552754493f;
  const float kNegBiasCbrt = -0.15595420054924863f;
  float ox[4], oy[4], ob[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float r = lut[v[3 * i]], g = lut[v[3 * i + 1]], b = lut[v[3 * i + 2]];
    float m0 = __fmaf_rn(0.30f, r, __fmaf_rn(0.622f, g, __fmaf_rn(0.078f, b, kBias)));
    float m1 = __fmaf_rn(0.23f, r, __fmaf_rn(0.692f, g, __fmaf_rn(0.078f, b, kBias)));
    float m2 = __fmaf_rn(0.24342268924547819f, r, __fmaf_rn(0.20476744424496821f, g, __fmaf_rn(0.55180986650955360f, b, kBias)));
    m0 = m0 > 0.0f ? m0 : 0.0f;
    m1 = m1 > 0.0f ? m1 : 0.0f;
    m2 = m2 > 0.0f ? m2 : 0.0f;
    const float L = cbrt_pos(m0) + kNegBiasCbrt;
    const float M = cbrt_pos(m1) + kNegBiasCbrt;
    const float S = cbrt_pos(m2) + kNegBiasCbrt;
    const bool pad = (x0 + i) >= fd.xs_pad;  // pitch padding is zero-filled
    ox[i] = pad ? 0.0f : 0.5f * (L - M);
    oy[i] = pad ? 0.0f : 0.5f * (L + M);
    ob[i] = pad ? 0.0f : S;
  }
  const size_t o = (size_t)y * fd.pitch + x0;
  *reinterpret_cast<float4*>(px + o) = make_float4(ox[0], ox[1], ox[2], ox[3]);
  *reinterpret_cast<float4*>(py + o) = make_float4(oy[0], oy[1], oy[2], oy[3]);
  *reinterpret_cast<float4*>(pb + o) = make_float4(ob[0], ob[1], ob[2], ob[3]);
}

void launch_rgb8_to_xyb(const uint8_t* d_rgb, size_t stride, int w, int h, const FrameDim& fd, const float* d_lut,
                        float* x, float* y, float* b, cudaStream_t s) {
  const int groups = fd.pitch / 4;
  dim3 grid((groups + 127) / 128, fd.ys_pad);
  ++g_kernel_launches;
  k_rgb8_to_xyb<<<grid, 128, 0, s>>>(d_rgb, stride, w, h, fd, d_lut, x, y, b);
}

}  // namespace jxlb
