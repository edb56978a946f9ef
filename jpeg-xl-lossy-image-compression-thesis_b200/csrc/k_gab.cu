// Gaborish, encoder side (row U3, opt-in: JXLB200_FLAG_GABORISH) — sharpens the three XYB planes with the 5x5 kernel that
// approximately inverts the decoder's default 3x3 Gaborish blur (oracle/jxo_xyb.cc GaborishInverse; weights from
// tools/gen_gab_inverse.py: libjxl's own hand-tuned constants are not available offline, and any kernel is a legal
// encoder choice).  Same association as the oracle: six symmetry-class sums, then five fused multiply-adds.
// One CTA = a 64 x 16 tile of one plane, staged with a 2-px mirrored halo; HBM: 12 B/px in, 12 B/px out.
#include "jxl_common.cuh"
#include "kernels.h"

namespace jxlb {

__device__ __forceinline__ int gab_mirror(int i, int n) {
  while (i < 0 || i >= n) i = i < 0 ? -i - 1 : 2 * n - 1 - i;
  return i;
}

__global__ void __launch_bounds__(256) k_gab_inverse(const float* __restrict__ src, float* __restrict__ dst, FrameDim fd) {
  __shared__ float t[20][68 + 1];
  const size_t plane = (size_t)fd.ys_pad * fd.pitch;
  const float* in = src + (size_t)blockIdx.z * plane;
  float* out = dst + (size_t)blockIdx.z * plane;
  const int x0 = blockIdx.x * 64, y0 = blockIdx.y * 16;
  const int W = fd.xs_pad, H = fd.ys_pad;
  for (int i = threadIdx.x; i < 20 * 68; i += 256) {
    const int r = i / 68, c = i % 68;
    t[r][c] = __ldg(in + (size_t)gab_mirror(y0 + r - 2, H) * fd.pitch + gab_mirror(x0 + c - 2, W));
  }
  __syncthreads();
  const int lx = threadIdx.x & 63, ly0 = threadIdx.x >> 6;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int ly = ly0 * 4 + k;
    const int x = x0 + lx, y = y0 + ly;
    if (y >= H || x >= fd.pitch) continue;
    float v = 0.0f;                                   // (pitch padding beyond xs_pad stays zero)
    if (x < W) {
      const int cx = lx + 2, cy = ly + 2;
      const float s0 = t[cy][cx];
      const float s1 = (t[cy][cx - 1] + t[cy][cx + 1]) + (t[cy - 1][cx] + t[cy + 1][cx]);
      const float s2 = (t[cy - 1][cx - 1] + t[cy - 1][cx + 1]) + (t[cy + 1][cx - 1] + t[cy + 1][cx + 1]);
      const float s3 = (t[cy][cx - 2] + t[cy][cx + 2]) + (t[cy - 2][cx] + t[cy + 2][cx]);
      const float s4 = ((t[cy - 1][cx - 2] + t[cy - 1][cx + 2]) + (t[cy + 1][cx - 2] + t[cy + 1][cx + 2])) +
                       ((t[cy - 2][cx - 1] + t[cy - 2][cx + 1]) + (t[cy + 2][cx - 1] + t[cy + 2][cx + 1]));
      const float s5 = (t[cy - 2][cx - 2] + t[cy - 2][cx + 2]) + (t[cy + 2][cx - 2] + t[cy + 2][cx + 2]);
      v = 1.758123398e+00f * s0;
      v = __fmaf_rn(-1.690203995e-01f, s1, v); v = __fmaf_rn(-7.491233945e-02f, s2, v); v = __fmaf_rn(2.443357371e-02f, s3, v);
      v = __fmaf_rn(1.462947764e-02f, s4, v); v = __fmaf_rn(7.093405002e-04f, s5, v);
    }
    out[(size_t)y * fd.pitch + x] = v;
  }
}

void launch_gab_inverse(const float* src, float* dst, const FrameDim& fd, cudaStream_t s) {
  dim3 grid((fd.pitch + 63) / 64, (fd.ys_pad + 15) / 16, 3);
  ++g_kernel_launches;
  k_gab_inverse<<<grid, 256, 0, s>>>(src, dst, fd);
}

}  // namespace jxlb
