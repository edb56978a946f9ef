// K4 — homogeneity map (rows H1-H9): the thesis' proposals as one precomputed pass.
// Restates proposals/homogeneity-partitioning.diff:17-211 (CalculateNumZeroCrossings,
// CalculateLaplacianFilter, CalculateSumModifiedLaplacian, CalculateColorfulness,
// CalculateHomogeneity, CalculateHomogeneitySimilarityIndices).  The CPU reference
// recomputes these on every EstimateEntropy candidate (factored-entropy.diff:248-253);
// they are a pure function of (block, distance, image), so here they are computed once
// per 8x8 block: r_h, r_v, r_d.
//
// One CTA = 32 horizontally adjacent blocks (256 x 8 px).  Phase 1: one thread per pixel
// column computes the per-pixel Laplacian and modified-Laplacian terms into shared
// memory.  Phase 2: one thread per (block, sub-rectangle) accumulates its eight
// homogeneity values in the reference's sequential order.  Phase 3: ratios.
// HBM traffic: 12 B/px in (+ halo re-reads from L2), 12 B/block out.
#include "jxl_common.cuh"
#include "kernels.h"

namespace jxlb {

__device__ __forceinline__ float hmax(float a, float b) { return (a < b) ? b : a; }  // std::max
__device__ __forceinline__ float hmin(float a, float b) { return (b < a) ? b : a; }  // std::min

__global__ void __launch_bounds__(256) k_homogeneity(const float* __restrict__ X, const float* __restrict__ Y,
                                                     const float* __restrict__ B, FrameDim fd, float distance,
                                                     float* __restrict__ out) {
  __shared__ float sy[10][260];   // rows -1..8, cols -1..256 (+pad)
  __shared__ float sx[8][256];
  __shared__ float sb[8][256];
  __shared__ float slap[8][256];
  __shared__ float ssml[8][256];
  __shared__ float sh[32][8];
  const int t = threadIdx.x;
  const int by = blockIdx.y;
  const int px0 = blockIdx.x * 256;
  const int gy0 = by * 8;
  // ---- load tiles (0 where the reference would not read) ------------------------------
  for (int i = t; i < 10 * 258; i += 256) {
    const int r = i / 258, c = i % 258;
    const int gy = gy0 + r - 1, gx = px0 + c - 1;
    float v = 0.0f;
    if (gy >= 0 && gy < fd.ys_pad && gx >= 0 && gx < fd.pitch) v = Y[(size_t)gy * fd.pitch + gx];
    sy[r][c] = v;
  }
  for (int i = t; i < 8 * 256; i += 256) {
    const int r = i >> 8, c = i & 255;
    const int gx = px0 + c;
    float vx = 0.0f, vb = 0.0f;
    if (gx < fd.pitch) { vx = X[(size_t)(gy0 + r) * fd.pitch + gx]; vb = B[(size_t)(gy0 + r) * fd.pitch + gx]; }
    sx[r][c] = vx; sb[r][c] = vb;
  }
  __syncthreads();
  // ---- phase 1: per-pixel terms, thread = pixel column ---------------------------------
  {
    const int c = t;
    const int gx = px0 + c;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const int gy = gy0 + r;
      const float p = sy[r + 1][c + 1], pl = sy[r + 1][c], pr = sy[r + 1][c + 2], pu = sy[r][c + 1], pd = sy[r + 2][c + 1];
      // CalculateLaplacianFilter (diff :57-81): mask {{0,-1,0},{-1,-4,-1},{0,-1,0}}, k-major order
      float sum = 0.0f;
      if (gy - 1 >= 0 && gx < fd.pitch) sum = __fmaf_rn(pu, -1.0f, sum);
      if (gx - 1 >= 0) sum = __fmaf_rn(pl, -1.0f, sum);
      if (gx < fd.pitch) sum = __fmaf_rn(p, -4.0f, sum);
      if (gx + 1 < fd.pitch) sum = __fmaf_rn(pr, -1.0f, sum);
      if (gy + 1 < fd.ys_pad && gx < fd.pitch) sum = __fmaf_rn(pd, -1.0f, sum);
      slap[r][c] = sum;
      // CalculateSumModifiedLaplacian term (diff :83-105); skipped pixels contribute +0
      float term = 0.0f;
      if (!(gx + 1 >= fd.pitch || gy + 1 >= fd.ys_pad || gx == 0 || gy == 0))
        term = fabsf(2 * p - pl - pr) + fabsf(2 * p - pu - pd);
      ssml[r][c] = term;
    }
  }
  __syncthreads();
  // ---- phase 2: thread = (block, sub-rectangle) ---------------------------------------
  {
    const int b = t >> 3, sr = t & 7;
    // (xsize, ysize, bx, by) of the eight calls in CalculateHomogeneitySimilarityIndices
    const int xs = (sr < 2) ? 8 : 4;
    const int ys = (sr < 2) ? 4 : ((sr < 4) ? 8 : 4);
    const int ox = (sr == 3 || sr == 5 || sr == 7) ? 4 : 0;
    const int oy = (sr == 1 || sr == 5 || sr == 6) ? 4 : 0;
    const int cb = b * 8;
    float thr = 0.25f;
    if (distance > 10.0f) thr = 0.40f; else if (distance <= 2.0f) thr = 0.15f;
    // zero crossings (diff :17-55)
    unsigned nh = 0, nv = 0;
    for (int i = 0; i < ys; ++i) {
      bool in_edge = false;
      for (int j = 0; j < xs; ++j) {
        const float v = slap[oy + i][cb + ox + j];
        if (!in_edge && v > thr) { nh++; in_edge = true; }
        else if (in_edge && v <= thr) { in_edge = false; }
      }
    }
    for (int i = 0; i < xs; ++i) {
      bool in_edge = false;
      for (int j = 0; j < ys; ++j) {
        const float v = slap[oy + j][cb + ox + i];
        if (!in_edge && v > thr) { nv++; in_edge = true; }
        else if (in_edge && v <= thr) { in_edge = false; }
      }
    }
    const float avg_h = (float)nh / (float)ys;
    const float avg_v = (float)nv / (float)xs;
    const unsigned long long crossings = (unsigned long long)(avg_h + avg_v);
    // sum modified Laplacian
    float sml = 0.0f;
    for (int i = 0; i < ys; ++i) for (int j = 0; j < xs; ++j) sml += ssml[oy + i][cb + ox + j];
    // colourfulness (diff :107-151)
    const float n = (float)(xs * ys);
    float mean_x = 0.0f, mean_b = 0.0f, var_x = 0.0f, var_b = 0.0f;
    for (int i = 0; i < ys; ++i) for (int j = 0; j < xs; ++j) mean_x += sx[oy + i][cb + ox + j];
    mean_x /= n;
    for (int i = 0; i < ys; ++i) for (int j = 0; j < xs; ++j) mean_b += sb[oy + i][cb + ox + j];
    mean_b /= n;
    for (int i = 0; i < ys; ++i) for (int j = 0; j < xs; ++j) {
      const float d = sx[oy + i][cb + ox + j] - mean_x;
      var_x = __fmaf_rn(d, d, var_x);
    }
    var_x /= n;
    for (int i = 0; i < ys; ++i) for (int j = 0; j < xs; ++j) {
      const float d = sb[oy + i][cb + ox + j] - mean_b;
      var_b = __fmaf_rn(d, d, var_b);
    }
    var_b /= n;
    const float s1 = var_x + var_b;
    const float s2 = __fmaf_rn(mean_x, mean_x, mean_b * mean_b);
    const float col = (float)(sqrt((double)s1) + 0.3 * sqrt((double)s2));
    sh[b][sr] = ((float)crossings + sml) + col;
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
