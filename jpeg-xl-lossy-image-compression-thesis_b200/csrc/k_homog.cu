// K4 — homogeneity map (rows H1-H9): the thesis' proposals as one precomputed pass.
// Restates proposals/homogeneity-partitioning.diff:17-211 (CalculateNumZeroCrossings,
// CalculateLaplacianFilter, CalculateSumModifiedLaplacian, CalculateColorfulness,
// CalculateHomogeneity, CalculateHomogeneitySimilarityIndices).  The CPU reference
// recomputes these on every EstimateEntropy candidate (factored-entropy.diff:248-253);
// they are a pure function of (block, distance, image), so here they are computed once
// per 8x8 block: r_h, r_v, r_d.
//
// One CTA = 32 horizontally adjacent blocks (256 x 8 px), 256 threads.
//  phase 1  thread = pixel column: per-pixel Laplacian (only its comparison with the threshold is ever used: one
//           bit per pixel, gathered into a row mask and a column mask per block) and modified-Laplacian term
//  phase 2  warp = one of the eight sub-rectangles of CalculateHomogeneitySimilarityIndices, lane = block: the
//           loop bounds are compile-time constants per warp; zero crossings are counted on the bit masks
//           (rising edges = m & ~(m << 1)); the float sums run in the reference's sequential order
//  phase 3  ratios
// Shared-memory rows carry one pad float per 32 columns (index c + (c >> 5)): the column-per-lane accesses of
// phase 1 and the block-per-lane (stride 8) accesses of phase 2 are both conflict-free.
// HBM traffic: 12 B/px in (+ halo re-reads from L2), 12 B/block out.
#include "jxl_common.cuh"
#include "kernels.h"

namespace jxlb {

__device__ __forceinline__ float hmax(float a, float b) { return (a < b) ? b : a; }  // std::max
__device__ __forceinline__ float hmin(float a, float b) { return (b < a) ? b : a; }  // std::min

constexpr int kHPitch = 264 + 8;                                     // 258 columns + pads
__device__ __forceinline__ int hidx(int c) { return c + (c >> 5); }  // padded column index

// one sub-rectangle (XS x YS at (OX, OY) inside the block) of block `cb / 8`: CalculateHomogeneity (diff :153-181)
// (the offsets are run-time values: three copies of the code — 8x4, 4x8, 4x4 — instead of eight, for the instruction cache)
template <int XS, int YS>
__device__ __noinline__ float homogeneity_rect(const float (*ssml)[kHPitch], const float (*sx)[kHPitch], const float (*sb)[kHPitch],
                                                  const uint8_t* rowmask, const uint8_t* colmask, int cb, int OX, int OY) {
  // zero crossings (diff :17-55): rising edges of `laplacian > threshold` along the rows, then along the columns
  unsigned nh = 0, nv = 0;
#pragma unroll
  for (int i = 0; i < YS; ++i) {
    const unsigned m = ((unsigned)rowmask[OY + i] >> OX) & ((1u << XS) - 1);
    nh += __popc(m & ~(m << 1));
  }
#pragma unroll
  for (int i = 0; i < XS; ++i) {
    const unsigned m = ((unsigned)colmask[OX + i] >> OY) & ((1u << YS) - 1);
    nv += __popc(m & ~(m << 1));
  }
  const float avg_h = (float)nh / (float)YS;
  const float avg_v = (float)nv / (float)XS;
  const unsigned long long crossings = (unsigned long long)(avg_h + avg_v);
  // sum modified Laplacian (diff :83-105), colourfulness (diff :107-151): row-major sequential sums
  float sml = 0.0f, mean_x = 0.0f, mean_b = 0.0f, var_x = 0.0f, var_b = 0.0f;
  float vx[XS * YS], vb[XS * YS];
#pragma unroll
  for (int i = 0; i < YS; ++i)
#pragma unroll
    for (int j = 0; j < XS; ++j) {
      const int c = hidx(cb + OX + j);
      sml += ssml[OY + i][c];
      vx[i * XS + j] = sx[OY + i][c];
      vb[i * XS + j] = sb[OY + i][c];
    }
  const float n = (float)(XS * YS);
#pragma unroll
  for (int k = 0; k < XS * YS; ++k) mean_x += vx[k];
  mean_x /= n;
#pragma unroll
  for (int k = 0; k < XS * YS; ++k) mean_b += vb[k];
  mean_b /= n;
#pragma unroll
  for (int k = 0; k < XS * YS; ++k) { const float d = vx[k] - mean_x; var_x = __fmaf_rn(d, d, var_x); }
  var_x /= n;
#pragma unroll
  for (int k = 0; k < XS * YS; ++k) { const float d = vb[k] - mean_b; var_b = __fmaf_rn(d, d, var_b); }
  var_b /= n;
  const float s1 = var_x + var_b;
  const float s2 = __fmaf_rn(mean_x, mean_x, mean_b * mean_b);
  const float col = (float)(sqrt((double)s1) + 0.3 * sqrt((double)s2));
  return ((float)crossings + sml) + col;
}

__global__ void __launch_bounds__(256) k_homogeneity(const float* __restrict__ X, const float* __restrict__ Y,
                                                     const float* __restrict__ B, FrameDim fd, float distance,
                                                     float* __restrict__ out) {
  __shared__ float sy[10][kHPitch];   // rows -1..8, columns -1..256 stored at hidx(column + 1)
  __shared__ float sx[8][kHPitch];
  __shared__ float sb[8][kHPitch];
  __shared__ float ssml[8][kHPitch];
  __shared__ uint8_t sbits[8][256];   // per pixel: laplacian > threshold
  __shared__ uint8_t srow[32][8], scol[32][8];
  __shared__ float sh[32][8];
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int by = blockIdx.y;
  const int px0 = blockIdx.x * 256;
  const int gy0 = by * 8;
  // ---- load tiles (0 where the reference would not read): thread = column, every load of the thread in flight before the
  // first store (the rolled one-load-one-store loops waited for 40 % of the kernel's stall samples: profiles/r02k)
  {
    const int gx = px0 + t;
    const bool col_in = gx < fd.pitch;
    float vy[10], vx[8], vb[8];
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      const int gy = gy0 + r - 1;
      vy[r] = (col_in && gy >= 0 && gy < fd.ys_pad) ? __ldg(Y + (size_t)gy * fd.pitch + gx) : 0.0f;
    }
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      vx[r] = col_in ? __ldg(X + (size_t)(gy0 + r) * fd.pitch + gx) : 0.0f;
      vb[r] = col_in ? __ldg(B + (size_t)(gy0 + r) * fd.pitch + gx) : 0.0f;
    }
    float halo = 0.0f;                          // columns -1 and 256 of the Y tile: threads 0..19 = (row, side)
    const int hr = t >> 1, hc = (t & 1) ? 257 : 0;
    if (t < 20) {
      const int gy = gy0 + hr - 1, hx = px0 + hc - 1;
      if (gy >= 0 && gy < fd.ys_pad && hx >= 0 && hx < fd.pitch) halo = __ldg(Y + (size_t)gy * fd.pitch + hx);
    }
    const int ci = hidx(t), cy = hidx(t + 1);
#pragma unroll
    for (int r = 0; r < 10; ++r) sy[r][cy] = vy[r];
#pragma unroll
    for (int r = 0; r < 8; ++r) { sx[r][ci] = vx[r]; sb[r][ci] = vb[r]; }
    if (t < 20) sy[hr][hidx(hc)] = halo;
  }
  __syncthreads();
  float thr = 0.25f;
  if (distance > 10.0f) thr = 0.40f; else if (distance <= 2.0f) thr = 0.15f;
  // ---- phase 1: per-pixel terms, thread = pixel column ---------------------------------
  {
    const int c = t;
    const int gx = px0 + c;
    const int cl = hidx(c), cm = hidx(c + 1), cr = hidx(c + 2);
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const int gy = gy0 + r;
      const float p = sy[r + 1][cm], pl = sy[r + 1][cl], pr = sy[r + 1][cr], pu = sy[r][cm], pd = sy[r + 2][cm];
      // CalculateLaplacianFilter (diff :57-81): mask {{0,-1,0},{-1,-4,-1},{0,-1,0}}, k-major order
      float sum = 0.0f;
      if (gy - 1 >= 0 && gx < fd.pitch) sum = __fmaf_rn(pu, -1.0f, sum);
      if (gx - 1 >= 0) sum = __fmaf_rn(pl, -1.0f, sum);
      if (gx < fd.pitch) sum = __fmaf_rn(p, -4.0f, sum);
      if (gx + 1 < fd.pitch) sum = __fmaf_rn(pr, -1.0f, sum);
      if (gy + 1 < fd.ys_pad && gx < fd.pitch) sum = __fmaf_rn(pd, -1.0f, sum);
      sbits[r][c] = sum > thr ? 1 : 0;
      // CalculateSumModifiedLaplacian term (diff :83-105); skipped pixels contribute +0
      float term = 0.0f;
      if (!(gx + 1 >= fd.pitch || gy + 1 >= fd.ys_pad || gx == 0 || gy == 0))
        term = fabsf(2 * p - pl - pr) + fabsf(2 * p - pu - pd);
      ssml[r][cl] = term;
    }
  }
  __syncthreads();
  // ---- per block: eight row masks (bit j = column j) and eight column masks (bit i = row i)
  {
    const int b = t >> 3, k = t & 7;
    unsigned rm = 0, cmk = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) { rm |= (unsigned)sbits[k][b * 8 + j] << j; cmk |= (unsigned)sbits[j][b * 8 + k] << j; }
    srow[b][k] = (uint8_t)rm; scol[b][k] = (uint8_t)cmk;
  }
  __syncthreads();
  // ---- phase 2: warp = sub-rectangle of CalculateHomogeneitySimilarityIndices (diff :183-211), lane = block ------
  {
    const int cb = lane * 8;
    const uint8_t* rmk = srow[lane]; const uint8_t* cmk = scol[lane];
    float h;
    // (xsize, ysize, bx, by) of the eight calls: warps 0-1 8x4 at y = 0 / 4, warps 2-3 4x8 at x = 0 / 4, warps 4-7 4x4 at
    // (0,0), (4,4), (0,4), (4,0)
    if (warp < 2) h = homogeneity_rect<8, 4>(ssml, sx, sb, rmk, cmk, cb, 0, warp * 4);
    else if (warp < 4) h = homogeneity_rect<4, 8>(ssml, sx, sb, rmk, cmk, cb, (warp - 2) * 4, 0);
    else h = homogeneity_rect<4, 4>(ssml, sx, sb, rmk, cmk, cb, (warp == 5 || warp == 7) ? 4 : 0, (warp == 5 || warp == 6) ? 4 : 0);
    sh[lane][warp] = h;
  }
  __syncthreads();
  // ---- phase 3: similarity indices (diff :183-211) -------------------------------------
  if (t < 32) {
    const int bx = blockIdx.x * 32 + t;
    if (bx < fd.bxs) {
      const float h1 = sh[t][0], h2 = sh[t][1], v1 = sh[t][2], v2 = sh[t][3];
      const float d1 = sh[t][4] + sh[t][5] / 2;
      const float d2 = sh[t][6] + sh[t][7] / 2;
      float* o = out + ((size_t)by * fd.bxs + bx) * 3;
      o[0] = hmax(h1, h2) / hmin(h1, h2);
      o[1] = hmax(v1, v2) / hmin(v1, v2);
      o[2] = hmax(d1, d2) / hmin(d1, d2);
    }
  }
}

void launch_homogeneity(const float* x, const float* y, const float* b, const FrameDim& fd, float distance,
                        float* out, cudaStream_t s) {
  dim3 grid((fd.bxs + 31) / 32, fd.bys);
  ++g_kernel_launches;
  k_homogeneity<<<grid, 256, 0, s>>>(x, y, b, fd, distance, out);
}

}  // namespace jxlb
