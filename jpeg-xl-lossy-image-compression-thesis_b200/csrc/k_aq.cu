// K2 — adaptive quantisation field (stage U2; libjxl enc_adaptive_quantization.cc [UPSTREAM]).
//   k_aq_pre   : per pixel — 5-point Laplacian of Y weighted by the cube-root->gamma ratio;
//                writes mask1x1 (f32/px) and the 4x4 "pre-erosion" cells (f32 per 16 px).
//   k_aq_block : per 8x8 block — fuzzy erosion of the 2x2 cells, ComputeMask, HF / colour /
//                gamma modulations, exp2 mapping -> quant_field (f32/block) and mask (f32/block).
//   k_quant_params : exact median / median-absolute-deviation of the quant field by radix
//                selection on the float bit patterns -> global_scale, quant_dc (Quantizer::
//                SetQuantField / ComputeGlobalScaleAndQuant).
//   k_raw_qf   : AdjustQuantField (max over a multi-block transform) + integer quant field.
// HBM traffic: k_aq_pre 4 B/px in + 4.25 B/px out; k_aq_block 12 B/px in (X, Y, B).
// Float sums follow the oracle's association: per-column accumulators, then the xor
// butterfly over the 8 lanes of a block group.
#include "jxl_common.cuh"
#include "kernels.h"

#include <cooperative_groups.h>
namespace cg = cooperative_groups;

namespace jxlb {

__device__ __forceinline__ float ratio_cbrt_to_gamma(float v, bool invert) {
  const float kEps = 1e-2f;
  const float kNumMul = 1.1990657e+02f;
  const float kVOffset = 5.4044867e+00f;
  const float kDenMul = 1.5718648e+02f;
  v = v > 0.0f ? v : 0.0f;
  const float v2 = v * v;
  const float num = __fmaf_rn(kNumMul, v2, kEps);
  const float den = __fmaf_rn(kDenMul * v, v2, kVOffset);
  return invert ? num / den : den / num;
}

__device__ __forceinline__ float masking_sqrt(float v) {
  return 0.25f * sqrtf(__fmaf_rn(v, 1.4543302e+05f, 26.481471032459346f));
}

__device__ __forceinline__ float compute_mask(float out_val) {
  const float kBase = -0.7647f, kMul4 = 9.4708735624378946f, kMul2 = 17.35036561631863f;
  const float kOffset2 = 302.59587815579727f, kMul3 = 6.7943250517376494f, kOffset3 = 3.7179635626140772f;
  const float kOffset4 = 0.25f * kOffset3, kMul0 = 0.80061762862741759f;
  float v1 = out_val * kMul0;
  v1 = v1 > 1e-3f ? v1 : 1e-3f;
  const float v2 = 1.0f / (v1 + kOffset2);
  const float v3 = 1.0f / __fmaf_rn(v1, v1, kOffset3);
  const float v4 = 1.0f / __fmaf_rn(v1, v1, kOffset4);
  return kBase + __fmaf_rn(kMul4, v4, __fmaf_rn(kMul2, v2, kMul3 * v3));
}

// tile: 32 x 8 cells of 4x4 px = 128 x 32 px, 256 threads (one per cell)
__global__ void __launch_bounds__(256) k_aq_pre(const float* __restrict__ Y, FrameDim fd, float* __restrict__ mask1x1,
                                                float* __restrict__ pre) {
  __shared__ float sy[34][132];
  const int t = threadIdx.x;
  const int gx0 = blockIdx.x * 128, gy0 = blockIdx.y * 32;
  const int xs = fd.xs_pad, ys = fd.ys_pad;
  {
    // (all of the thread's loads in flight before the first store: the rolled one-load-one-store loop was half of the
    // kernel's stall samples, profiles/r02k)
    float tmp[18];
#pragma unroll
    for (int k = 0; k < 18; ++k) {
      const int i = t + k * 256;
      tmp[k] = 0.0f;
      if (i < 34 * 130) {
        const int r = i / 130, c = i % 130;
        int gy = gy0 + r - 1, gx = gx0 + c - 1;
        gy = gy < 0 ? 0 : (gy > ys - 1 ? ys - 1 : gy);
        gx = gx < 0 ? 0 : (gx > xs - 1 ? xs - 1 : gx);
        tmp[k] = __ldg(Y + (size_t)gy * fd.pitch + gx);
      }
    }
#pragma unroll
    for (int k = 0; k < 18; ++k) {
      const int i = t + k * 256;
      if (i < 34 * 130) sy[i / 130][i % 130] = tmp[k];
    }
  }
  __syncthreads();
  const int cx = t & 31, cy = t >> 5;
  const int x0 = gx0 + cx * 4, y0 = gy0 + cy * 4;
  if (x0 >= fd.pitch || y0 >= ys) return;
  if (x0 >= xs) {  // pitch padding
    for (int r = 0; r < 4; ++r) *reinterpret_cast<float4*>(mask1x1 + (size_t)(y0 + r) * fd.pitch + x0) = make_float4(0, 0, 0, 0);
    return;
  }
  float acc[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    float m[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int lr = cy * 4 + r + 1, lc = cx * 4 + j + 1;
      const float p = sy[lr][lc];
      const float base = 0.25f * (((sy[lr + 1][lc] + sy[lr - 1][lc]) + sy[lr][lc - 1]) + sy[lr][lc + 1]);
      const float gammac = ratio_cbrt_to_gamma(p + 0.019f, false);
      float diff = gammac * (p - base);
      const float l1 = fast_log2f(1.0f + fabsf(diff)) * 0.69314718f;
      m[j] = 1.0f / (l1 + 0.01f);
      diff = diff * diff;
      if (diff >= 0.2f) diff = 0.2f;
      diff = masking_sqrt(diff);
      if (r == 0) acc[j] = diff; else acc[j] += diff;
    }
    *reinterpret_cast<float4*>(mask1x1 + (size_t)(y0 + r) * fd.pitch + x0) = make_float4(m[0], m[1], m[2], m[3]);
  }
  pre[(size_t)(y0 / 4) * (xs / 4) + (x0 / 4)] = (((acc[0] + acc[1]) + acc[2]) + acc[3]) * 0.25f;
}

__device__ __forceinline__ void store_min4(float v, float& m0, float& m1, float& m2, float& m3) {
  if (v < m3) {
    if (v < m0) { m3 = m2; m2 = m1; m1 = m0; m0 = v; }
    else if (v < m1) { m3 = m2; m2 = m1; m1 = v; }
    else if (v < m2) { m3 = m2; m2 = v; }
    else { m3 = v; }
  }
}
__device__ __forceinline__ void cswap(float& a, float& b) { if (a > b) { const float t = a; a = b; b = t; } }

// warp = 4 blocks x 8 lanes (lane = pixel column); CTA = 32 blocks of one block row
__global__ void __launch_bounds__(256) k_aq_block(const float* __restrict__ X, const float* __restrict__ Y,
                                                  const float* __restrict__ B, const float* __restrict__ pre,
                                                  FrameDim fd, float distance, float* __restrict__ qf,
                                                  float* __restrict__ mask) {
  const int t = threadIdx.x;
  const int l8 = t & 7;
  int bx = blockIdx.x * 32 + (t >> 3);
  const int by = blockIdx.y;
  const bool active = bx < fd.bxs;
  if (!active) bx = fd.bxs - 1;  // keep the warp converged for the shuffles
  const int pw = fd.xs_pad / 4, ph = fd.ys_pad / 4;
  // ---- fuzzy erosion on lanes 0..3 (cells (l8&1, l8>>1) of the block) ---------------------
  float mulv = 0.0f;
  if (distance < 2.0f) mulv = (2.0f - distance) * 0.5f;
  float k0 = 0.125f + mulv * 0.0f, k1 = 0.10f + mulv * -0.10f, k2 = 0.09f + mulv * -0.09f, k3 = 0.06f + mulv * -0.06f;
  const float norm = 0.29959705784054957f / (((k0 + k1) + k2) + k3);
  k0 *= norm; k1 *= norm; k2 *= norm; k3 *= norm;
  float v = 0.0f;
  if (l8 < 4) {
    const int x = 2 * bx + (l8 & 1), y = 2 * by + (l8 >> 1);
    const int xm1 = x >= 1 ? x - 1 : x, xp1 = x + 1 < pw ? x + 1 : x;
    const int ym1 = y >= 1 ? y - 1 : y, yp1 = y + 1 < ph ? y + 1 : y;
    const float* rt = pre + (size_t)ym1 * pw;
    const float* r = pre + (size_t)y * pw;
    const float* rb = pre + (size_t)yp1 * pw;
    float m0 = r[x], m1 = r[xm1], m2 = r[xp1], m3 = rt[xm1];
    cswap(m0, m1); cswap(m0, m2); cswap(m0, m3); cswap(m1, m2); cswap(m1, m3); cswap(m2, m3);
    store_min4(rt[x], m0, m1, m2, m3);
    store_min4(rt[xp1], m0, m1, m2, m3);
    store_min4(rb[xm1], m0, m1, m2, m3);
    store_min4(rb[x], m0, m1, m2, m3);
    store_min4(rb[xp1], m0, m1, m2, m3);
    v = ((k0 * m0 + k1 * m1) + k2 * m2) + k3 * m3;
  }
  const int base_lane = (threadIdx.x & 31) & ~7;
  const float v00 = __shfl_sync(0xffffffffu, v, base_lane + 0);
  const float v10 = __shfl_sync(0xffffffffu, v, base_lane + 1);
  const float v01 = __shfl_sync(0xffffffffu, v, base_lane + 2);
  const float v11 = __shfl_sync(0xffffffffu, v, base_lane + 3);
  const float aq = ((v00 + v10) + v01) + v11;

  // ---- column accumulators -------------------------------------------------------------
  const size_t o0 = (size_t)(by * 8) * fd.pitch + (size_t)bx * 8 + l8;
  float yv[8], xv[8], bv[8];
#pragma unroll
  for (int dy = 0; dy < 8; ++dy) {
    yv[dy] = Y[o0 + (size_t)dy * fd.pitch];
    xv[dy] = X[o0 + (size_t)dy * fd.pitch];
    bv[dy] = B[o0 + (size_t)dy * fd.pitch];
  }
  const float valmin = 0.020602694503245016f;
  float hf = 0.0f;
#pragma unroll
  for (int dy = 0; dy < 8; ++dy) {
    const float right = __shfl_down_sync(0xffffffffu, yv[dy], 1);
    if (l8 < 7) { const float a = fabsf(yv[dy] - right); hf += a < valmin ? a : valmin; }
    const float down = dy == 7 ? yv[7] : yv[dy + 1 > 7 ? 7 : dy + 1];
    const float b = fabsf(yv[dy] - down);
    hf += b < valmin ? b : valmin;
  }
  const float hf_sum = tree8_sum(hf);

  const float strength = 3.0f * (1.0f - 0.25f * distance);
  const float kRedRampStart = 0.0073200141118951231f, kRedRampLength = 0.019421555948474039f;
  const float kBlueRampLength = 0.086890611400405895f, kBlueRampStart = 0.26973418507870539f;
  float red = 0.0f, blue = 0.0f, gam = 0.0f;
#pragma unroll
  for (int dy = 0; dy < 8; ++dy) {
    float pxv = xv[dy] - kRedRampStart; pxv = pxv > 0.0f ? pxv : 0.0f;
    float pbv = bv[dy] - (yv[dy] + kBlueRampStart); pbv = pbv > 0.0f ? pbv : 0.0f;
    blue += pbv < kBlueRampLength ? pbv : kBlueRampLength;
    red += pxv < kRedRampLength ? pxv : kRedRampLength;
    const float iny = yv[dy] + 0.16f, inx = xv[dy];
    const float rr = ratio_cbrt_to_gamma(iny - inx, true);
    const float rg = ratio_cbrt_to_gamma(iny + inx, true);
    gam += 0.5f * (rr + rg);
  }
  float red_sum = tree8_sum(red);
  float blue_sum = tree8_sum(blue);
  float overall = tree8_sum(gam);

  if (l8 == 0 && active) {
    const float scale = 0.841f / distance;
    const float base_level = 0.48f * scale;
    float dampen = 1.0f;
    if (distance >= 2.0f) { dampen = 1.0f - ((distance - 2.0f) / (14.0f - 2.0f)); if (dampen < 0) dampen = 0; }
    const float qmul = scale * dampen;
    const float qadd = (1.0f - dampen) * base_level;
    float out_val = compute_mask(aq);
    out_val = (hf_sum + -1.110929106987477f) * -0.38078920620238305f + out_val;
    if (!(strength < 0)) {
      const float red_strength = strength * 5.992297772961519f, blue_strength = strength;
      out_val = out_val + strength * -0.009174542291185913f;
      const float ratio = 30.610615782142737f;
      red_sum = red_sum < ratio * kRedRampLength ? red_sum : ratio * kRedRampLength;
      red_sum = red_sum * (red_strength / ratio);
      blue_sum = blue_sum < ratio * kBlueRampLength ? blue_sum : ratio * kBlueRampLength;
      blue_sum = blue_sum * (blue_strength / ratio);
      out_val = red_sum + (blue_sum + out_val);
    }
    overall = overall * (1.0f / 64.0f);
    out_val = __fmaf_rn(1.00561336e-01f, fast_log2f(overall), out_val);
    const size_t bi = (size_t)by * fd.bxs + bx;
    qf[bi] = fast_pow2f(out_val * 1.442695041f) * qmul + qadd;
    mask[bi] = 1.0f / (aq + 0.001f);
  }
}

__global__ void k_fill(float* p, size_t n, float v) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

// ---- exact k-th smallest by 3-pass radix selection on positive-float bit patterns --------
// ONE THREAD-BLOCK CLUSTER of 8 CTAs (the field has one float per 8x8 block: 130k values for a 4K frame; a single CTA
// spent 207 us on the six passes).  Every CTA histograms one eighth of the values into its own shared memory; the bins
// are then summed slice-wise through distributed shared memory (CTA r owns bins [r * nb/8, (r+1) * nb/8) of all eight
// histograms), the slice sums are exchanged the same way, and the CTA whose slice holds the k-th element scans it and
// writes the selected digit into every CTA's shared memory.  Three cluster barriers per pass, no global scratch.
constexpr int kQpCtas = 8;
constexpr int kQpThreads = 1024;

struct SelectShared {
  uint32_t hist[4096];
  uint32_t tot[4096 / kQpCtas];
  uint32_t slice_sum[kQpCtas];
  uint32_t prefix, k;
};

template <bool kDeviation>
__device__ uint32_t radix_select(cg::cluster_group& cluster, const float* __restrict__ v, size_t n, size_t k, float center,
                                 SelectShared& sh) {
  const int rank = (int)cluster.block_rank(), t = threadIdx.x, lane = t & 31;
  uint32_t prefix = 0;  // selected high bits so far
  uint32_t kk = (uint32_t)k;
  const int shifts[3] = {20, 8, 0};
  const int bits[3] = {12, 12, 8};
  int done_bits = 0;
  for (int pass = 0; pass < 3; ++pass) {
    const int nb = 1 << bits[pass], per = nb / kQpCtas;
    for (int i = t; i < nb; i += kQpThreads) sh.hist[i] = 0;
    __syncthreads();
    // the field is narrow-ranged, so most values share a digit: aggregate equal digits inside the
    // warp (match.any) and let one lane per distinct digit do the shared-memory atomic
    for (size_t i0 = (size_t)rank * kQpThreads * 8; i0 < n; i0 += (size_t)kQpCtas * kQpThreads * 8) {
      float vals[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {       // eight independent loads in flight before any of them is consumed
        const size_t i = i0 + (size_t)u * kQpThreads + t;
        vals[u] = i < n ? v[i] : 0.0f;
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const size_t i = i0 + (size_t)u * kQpThreads + t;
        bool match = false;
        uint32_t digit = 0;
        if (i < n) {
          const float f = kDeviation ? fabsf(vals[u] - center) : vals[u];
          const uint32_t uu = __float_as_uint(f);
          match = done_bits == 0 ? true : ((uu >> (32 - done_bits)) == (prefix >> (32 - done_bits)));
          digit = (uu >> shifts[pass]) & (nb - 1);
        }
        const unsigned part = __ballot_sync(0xffffffffu, match);
        if (match) {
          const unsigned peers = __match_any_sync(part, digit);
          if (lane == (__ffs(peers) - 1)) atomicAdd(&sh.hist[digit], (uint32_t)__popc(peers));
        }
      }
    }
    cluster.sync();
    // my slice of the bins, summed over the eight histograms
    if (t < per) {
      uint32_t sum = 0;
#pragma unroll
      for (int r = 0; r < kQpCtas; ++r) sum += cluster.map_shared_rank(sh.hist, r)[rank * per + t];
      sh.tot[t] = sum;
    }
    __syncthreads();
    if (t < 32) {
      uint32_t run = 0;
      for (int b = lane; b < per; b += 32) run += sh.tot[b];
#pragma unroll
      for (int d = 16; d >= 1; d >>= 1) run += __shfl_xor_sync(0xffffffffu, run, d);
      if (lane < kQpCtas) cluster.map_shared_rank(sh.slice_sum, lane)[rank] = run;
    }
    cluster.sync();
    if (t < 32) {
      uint32_t excl = 0;
      for (int r = 0; r < rank; ++r) excl += sh.slice_sum[r];
      if (excl <= kk && kk < excl + sh.slice_sum[rank]) {
        // this CTA's slice holds the k-th element: each lane sums a contiguous run of per/32 bins (one bin per lane in the
        // 8-bit pass), a warp scan picks the run, the owning lane the bin
        const int run_len = per >= 32 ? per / 32 : 1;
        uint32_t run = 0;
        if (lane < per) for (int b = lane * run_len; b < (lane + 1) * run_len; ++b) run += sh.tot[b];
        uint32_t incl = run;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const uint32_t o = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += o; }
        const uint32_t lo = excl + incl - run;
        if (lane < per && lo <= kk && kk < lo + run) {
          uint32_t acc = lo; int b = lane * run_len;
          for (; b < (lane + 1) * run_len; ++b) { if (acc + sh.tot[b] > kk) break; acc += sh.tot[b]; }
          const uint32_t new_k = kk - acc, new_prefix = prefix | ((uint32_t)(rank * per + b) << shifts[pass]);
          for (int r = 0; r < kQpCtas; ++r) {
            SelectShared* rs = cluster.map_shared_rank(&sh, r);
            rs->k = new_k; rs->prefix = new_prefix;
          }
        }
      }
    }
    cluster.sync();
    kk = sh.k; prefix = sh.prefix;
    done_bits += bits[pass];
  }
  return prefix;
}

__global__ void __cluster_dims__(kQpCtas, 1, 1) __launch_bounds__(kQpThreads)
    k_quant_params(const float* __restrict__ qf, size_t n, float quant_dc, QuantDev* __restrict__ q) {
  __shared__ SelectShared sh;
  cg::cluster_group cluster = cg::this_cluster();
  const float median = __uint_as_float(radix_select<false>(cluster, qf, n, n / 2, 0.0f, sh));
  const float mad = __uint_as_float(radix_select<true>(cluster, qf, n, n / 2, median, sh));
  cluster.sync();   // nobody leaves while its shared memory can still be addressed by a peer
  if (cluster.block_rank() == 0 && threadIdx.x == 0) {
    float scale = 65536.0f * (median - mad) / 5.0f;
    if (!(scale >= 1.0f)) scale = 1.0f;
    if (scale > 32768.0f) scale = 32768.0f;
    int gs = (int)scale;
    const int scaled_quant_dc = (int)(quant_dc * 4096.0f * 1.6f);
    if (gs > scaled_quant_dc) { gs = scaled_quant_dc; if (gs <= 0) gs = 1; }
    const float inv_gs = 65536.0f / (float)gs;
    float fval = quant_dc * inv_gs + 0.5f;
    if (fval > 65536.0f) fval = 65536.0f;
    int qdc = (int)fval;
    if (qdc < 1) qdc = 1;
    q->global_scale = gs; q->quant_dc = qdc;
    q->scale = (float)gs * (1.0f / 65536.0f);
    q->inv_global_scale = inv_gs;
    q->median = median; q->mad = mad;
  }
}

// AdjustQuantField (max over a multi-block transform) + SetQuantFieldRect
__global__ void k_raw_qf(const float* __restrict__ qf, const uint8_t* __restrict__ acs, FrameDim fd,
                         const QuantDev* __restrict__ q, const uint8_t* __restrict__ covered_x,
                         const uint8_t* __restrict__ covered_y, int32_t* __restrict__ raw) {
  const int bx = blockIdx.x * blockDim.x + threadIdx.x, by = blockIdx.y;
  if (bx >= fd.bxs) return;
  const uint8_t a = acs[(size_t)by * fd.bxs + bx];
  if (!(a & 0x80)) return;
  const int s = a & 0x7f, cx = covered_x[s], cy = covered_y[s];
  float m = qf[(size_t)by * fd.bxs + bx];
  for (int iy = 0; iy < cy; ++iy) for (int ix = 0; ix < cx; ++ix) {
    const float v = qf[(size_t)(by + iy) * fd.bxs + bx + ix];
    if (v > m) m = v;
  }
  int val = (int)(m * q->inv_global_scale + 0.5f);
  val = val < 1 ? 1 : (val > 256 ? 256 : val);
  for (int iy = 0; iy < cy; ++iy) for (int ix = 0; ix < cx; ++ix) raw[(size_t)(by + iy) * fd.bxs + bx + ix] = val;
}

void launch_aq(const float* x, const float* y, const float* b, const FrameDim& fd, float distance, float* mask1x1,
               float* pre, float* qf, float* mask, cudaStream_t s) {
  dim3 g1((fd.pitch + 127) / 128, (fd.ys_pad + 31) / 32);
  ++g_kernel_launches;
  k_aq_pre<<<g1, 256, 0, s>>>(y, fd, mask1x1, pre);
  dim3 g2((fd.bxs + 31) / 32, fd.bys);
  ++g_kernel_launches;
  k_aq_block<<<g2, 256, 0, s>>>(x, y, b, pre, fd, distance, qf, mask);
}

void launch_fill(float* p, size_t n, float v, cudaStream_t s) {
  ++g_kernel_launches;
  k_fill<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(p, n, v);
}

void launch_quant_params(const float* qf, size_t n, float quant_dc, QuantDev* q, cudaStream_t s) {
  ++g_kernel_launches;
  k_quant_params<<<kQpCtas, kQpThreads, 0, s>>>(qf, n, quant_dc, q);
}

void launch_raw_qf(const float* qf, const uint8_t* acs, const FrameDim& fd, const QuantDev* q, const uint8_t* cvx,
                   const uint8_t* cvy, int32_t* raw, cudaStream_t s) {
  dim3 g((fd.bxs + 127) / 128, fd.bys);
  ++g_kernel_launches;
  k_raw_qf<<<g, 128, 0, s>>>(qf, acs, fd, q, cvx, cvy, raw);
}

}  // namespace jxlb
