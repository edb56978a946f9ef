// K13 — reconstruction error of the frame that was just coded (the quantity behind
// calculate_mse / calculate_psnr of the harness, benchmark-jpegxl/src/image_reader.rs:555-606, taken from the
// integers the codestream carries instead of a djxl round trip): dequantise the stored coefficients, put the
// lowest frequencies back from the DC image, chroma-from-luma, inverse transform, XYB -> linear -> 8-bit sRGB,
// squared error against the input pixels.  Same arithmetic and operation order as oracle/jxo_recon.cc
// (ReconstructRgb / ReconstructionSse), so the three 64-bit sums are bit-exact.
//
// Mirror of k_coeff.cu: one CTA per 32x32-pixel square, the transforms whose first block lies in the square are
// taken round-robin by the warps, lane y of a transform's lane group owns coefficient row y (transforms.cuh).
// The reconstructed XYB square lives in shared memory only; HBM traffic is the coefficients (6 B/px), the
// input pixels (3 B/px) and ~0.2 B/px of per-block data.  Enabled by JXLB200_FLAG_QUALITY.
#include "transforms.cuh"
#include "kernels.h"

namespace jxlb {
namespace {

constexpr int kReconWarps = 4;

struct ReconShared {
  float px[3][32 * kTPitch];
  float buf[kReconWarps][3][32 * kTPitch];
  float tab[9 + 255];               // inverse opsin matrix, sRGB code boundaries in linear light
};

struct ReconArgs {
  FrameDim fd;
  const QuantDev* qd;
  const float* dq[17];              // dequantisation matrices per quant-table kind
  const uint16_t* inv_order[13];    // per order class: coefficient position -> scan index
  const int8_t* cmap;
  float inv_qm_x, inv_qm_b;
  const uint8_t* acs;
  const int32_t* raw_qf;
  const int16_t* coeffs;
  const int16_t* dc_quant;
  const uint8_t* rgb;
  size_t stride;
  const float* tables;
  unsigned long long* sse;
};

__device__ __forceinline__ float dequant_bias(int c, int q) {
  const float b0 = 1.0f - 0.05465007330715401f, b1 = 1.0f - 0.07005449891748593f, b2 = 1.0f - 0.049935103337343655f;
  const float bc = c == 0 ? b0 : (c == 1 ? b1 : b2);
  if (q == 0) return 0.0f;
  if (q == 1) return bc;
  if (q == -1) return -bc;
  return (float)q - 0.145f / (float)q;
}

__device__ __forceinline__ float recon_resample_scale(int n_from, int n_to, int k) {
  if (n_to == 1) return 1.0f;
  if (n_from == 16) return k == 0 ? 1.e+00f : 9.017642e-01f;
  return k == 0 ? 1.e+00f : (k == 1 ? 9.7488683e-01f : (k == 2 ? 9.017642e-01f : 7.870549e-01f));
}

// oracle LowestFrequenciesFromDc: the cy x cx lowest frequencies of the transform from its blocks' DC values
template <int S>
__device__ void llf_from_dc(const float* dc /*[cy*cx]*/, float* llf /*[cy*cx]*/) {
  constexpr int R = StratDim<S>::R, C = StratDim<S>::C, cy = R / 8, cx = C / 8;
  if constexpr (cx == 1 && cy == 1) { llf[0] = dc[0]; return; }
  else {
    float t[16], f[16];
#pragma unroll
    for (int y = 0; y < cy; ++y) {
      float v[cx];
#pragma unroll
      for (int x = 0; x < cx; ++x) v[x] = dc[y * cx + x];
      dct1d<cx>(v);
#pragma unroll
      for (int x = 0; x < cx; ++x) t[y * cx + x] = v[x];
    }
#pragma unroll
    for (int hf = 0; hf < cx; ++hf) {
      float v[cy];
#pragma unroll
      for (int y = 0; y < cy; ++y) v[y] = t[y * cx + hf];
      dct1d<cy>(v);
#pragma unroll
      for (int y = 0; y < cy; ++y) f[y * cx + hf] = v[y];
    }
#pragma unroll
    for (int vf = 0; vf < cy; ++vf)
#pragma unroll
      for (int hf = 0; hf < cx; ++hf)
        llf[vf * cx + hf] = f[vf * cx + hf] / (recon_resample_scale(R, cy, vf) * recon_resample_scale(C, cx, hf));
  }
}

template <int S>
__device__ void recon_transform(ReconShared& sh, int warp, bool active, int ox, int oy, int bx, int by, const ReconArgs& A,
                                int kind, int order_class, int lane) {
  constexpr int R = StratDim<S>::R, C = StratDim<S>::C;
  constexpr int W = R > C ? R : C, H = R > C ? C : R, cxb = C / 8, cyb = R / 8, n = cxb * cyb, size = R * C;
  constexpr int GS = W;
  const FrameDim& fd = A.fd;
  const int gl = lane & (GS - 1), go = (lane / GS) * GS * kTPitch;
  float* bufs[3] = {sh.buf[warp][0] + go, sh.buf[warp][1] + go, sh.buf[warp][2] + go};
  const size_t nblk = (size_t)fd.bxs * fd.bys;
  const size_t bi = (size_t)by * fd.bxs + bx;
  const float inv_gs = A.qd->inv_global_scale;
  const float inv_qac = inv_gs / (float)(active ? A.raw_qf[bi] : 1);
  const int tx = bx >> 3, ty = by >> 3;
  const float cfl_x = 0.0f + (float)A.cmap[(size_t)ty * fd.txs + tx] / 84.0f;
  const float cfl_b = 1.0f + (float)A.cmap[(size_t)fd.txs * fd.tys + (size_t)ty * fd.txs + tx] / 84.0f;
  const uint16_t* inv = A.inv_order[order_class];
  const float* dq = A.dq[kind];
  // ---- dequantise: Y first, X and B add their chroma-from-luma share of the (AC-only) Y
#pragma unroll 1
  for (int it = 0; it < 3; ++it) {
    const int c = it == 0 ? 1 : (it == 1 ? 0 : 2), slot = it;
    const float mul = inv_qac * (c == 0 ? A.inv_qm_x : (c == 1 ? 1.0f : A.inv_qm_b));
    const float cfl = c == 0 ? cfl_x : cfl_b;
    if (gl < H) {
      uint2 inv4 = make_uint2(0u, 0u);
      float4 dq4 = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
      for (int x = 0; x < W; ++x) {
        const int pos = gl * W + x;
        if ((x & 3) == 0) {   // scan indices and dequantisation weights of four positions per vector load
          inv4 = __ldg(reinterpret_cast<const uint2*>(inv + pos));
          dq4 = __ldg(reinterpret_cast<const float4*>(dq + (size_t)c * size + pos));
        }
        const int k = (int)(((x & 2) ? inv4.y : inv4.x) >> ((x & 1) * 16)) & 0xFFFF;
        const float dqv = (x & 3) == 0 ? dq4.x : ((x & 3) == 1 ? dq4.y : ((x & 3) == 2 ? dq4.z : dq4.w));
        const int j = k >> 6;
        const int cbx = bx + (j % cxb), cby = by + (j / cxb);
        const int g = (cby >> 5) * fd.gxs + (cbx >> 5);
        const size_t blk = (size_t)g * kGroupBlocks + (size_t)(cby & 31) * 32 + (cbx & 31);
        const int q = active ? (int)A.coeffs[(blk * 3 + slot) * 64 + (k & 63)] : 0;
        float v = (dequant_bias(c, q) * dqv) * mul;
        if (c != 1) v = __fmaf_rn(cfl, bufs[1][gl * kTPitch + x], v);
        bufs[c][gl * kTPitch + x] = v;
      }
    }
  }
  __syncwarp();
  // ---- lowest frequencies from the dequantised DC of the covered blocks (Y, then X / B with the DC correlation)
  if (gl == 0 && active) {
    const float inv_quant_dc = inv_gs / (float)A.qd->quant_dc;
    const float step_x = inv_quant_dc * (1.0f / 4096.0f), step_y = inv_quant_dc * (1.0f / 512.0f), step_b = inv_quant_dc * (1.0f / 256.0f);
    float dc[3][16];
    for (int j = 0; j < n; ++j) {
      const size_t bj = bi + (size_t)(j / cxb) * fd.bxs + (j % cxb);
      const float y = (float)A.dc_quant[nblk + bj] * step_y;
      dc[1][j] = y;
      dc[0][j] = __fmaf_rn(0.0f, y, (float)A.dc_quant[bj] * step_x);
      dc[2][j] = __fmaf_rn(1.0f, y, (float)A.dc_quant[2 * nblk + bj] * step_b);
    }
#pragma unroll 1
    for (int c = 0; c < 3; ++c) {
      float llf[16];
      llf_from_dc<S>(dc[c], llf);
#pragma unroll
      for (int vf = 0; vf < cyb; ++vf)
#pragma unroll
        for (int hf = 0; hf < cxb; ++hf) {
          // coefficient layout has the long side horizontal: (hf, vf) swap for tall / square transforms
          if (R >= C) bufs[c][hf * kTPitch + vf] = llf[vf * cxb + hf];
          else bufs[c][vf * kTPitch + hf] = llf[vf * cxb + hf];
        }
    }
  }
  __syncwarp();
  // ---- inverse transform into the square (inactive groups transform their scratch in place)
#pragma unroll 1
  for (int c = 0; c < 3; ++c) inv_transform<S>(bufs[c], bufs[c], active ? sh.px[c] + oy * kTPitch + ox : bufs[c], gl);
  __syncwarp();
}

template <int S>
__device__ void recon_strategy(ReconShared& sh, int warp, unsigned mask, int sbx, int sby, const ReconArgs& A, int kind,
                               int order_class, int lane) {
  constexpr int R = StratDim<S>::R, C = StratDim<S>::C;
  constexpr int GS = R > C ? R : C, GPW = 32 / GS;
  const int m = __popc(mask);
  for (int base = warp * GPW; base < m; base += kReconWarps * GPW) {
    const int idx = base + lane / GS;
    const bool active = idx < m;
    const int b = active ? (int)__fns(mask, 0, idx + 1) : 0;
    const int lx = b & 3, ly = b >> 2;
    recon_transform<S>(sh, warp, active, lx * 8, ly * 8, sbx + lx, sby + ly, A, kind, order_class, lane);
  }
}

// XYB sample -> 8-bit sRGB code of channel k (oracle XybToSrgb8): boundaries[n] is the smallest linear value coded n+1
__device__ __forceinline__ int srgb_code(const float* tab, int k, float mix0, float mix1, float mix2) {
  const float lin = __fmaf_rn(tab[3 * k], mix0, __fmaf_rn(tab[3 * k + 1], mix1, tab[3 * k + 2] * mix2));
  const float* bnd = tab + 9;
  int lo = 0, hi = 255;
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const int mid = (lo + hi) >> 1;
    if (lo < hi) { if (bnd[mid] <= lin) lo = mid + 1; else hi = mid; }
  }
  return lo;
}

__global__ void __launch_bounds__(kReconWarps * 32) k_recon_sse(ReconArgs A) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  ReconShared& sh = *reinterpret_cast<ReconShared*>(smem_raw);
  const FrameDim& fd = A.fd;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int sbx = blockIdx.x * 4, sby = blockIdx.y * 4;
  const int bw = min(4, fd.bxs - sbx), bh = min(4, fd.bys - sby);
  for (int i = t; i < 9 + 255; i += kReconWarps * 32) sh.tab[i] = A.tables[i];
  int my_s = -1;
  if (lane < 16) {
    const int lx = lane & 3, ly = lane >> 2;
    if (lx < bw && ly < bh) {
      const uint8_t a = A.acs[(size_t)(sby + ly) * fd.bxs + sbx + lx];
      if (a & 0x80) my_s = a & 0x7f;
    }
  }
  __syncthreads();
#define JXLB_STRATEGY(S, KIND, ORD) \
  { const unsigned mk = __ballot_sync(0xffffffffu, my_s == S); if (mk) recon_strategy<S>(sh, warp, mk, sbx, sby, A, KIND, ORD, lane); }
  JXLB_STRATEGY(kStratDCT, 0, 0)
  JXLB_STRATEGY(kStratDCT4X4, 3, 1)
  JXLB_STRATEGY(kStratDCT4X8, 9, 1)
  JXLB_STRATEGY(kStratDCT8X4, 9, 1)
  JXLB_STRATEGY(kStratDCT16X8, 6, 4)
  JXLB_STRATEGY(kStratDCT8X16, 6, 4)
  JXLB_STRATEGY(kStratDCT16X16, 4, 2)
  JXLB_STRATEGY(kStratDCT32X16, 8, 6)
  JXLB_STRATEGY(kStratDCT16X32, 8, 6)
  JXLB_STRATEGY(kStratDCT32X32, 5, 3)
#undef JXLB_STRATEGY
  __syncthreads();
  // ---- XYB -> sRGB codes, squared error against the input
  const float kBias = 0.0037930732552754493f, kNegBiasCbrt = -0.15595420054924863f;
  unsigned sse[3] = {0u, 0u, 0u};
  for (int i = t; i < 32 * 32; i += kReconWarps * 32) {
    const int y = i >> 5, x = i & 31;
    const int gx = sbx * 8 + x, gy = sby * 8 + y;
    if (gx >= fd.xsize || gy >= fd.ysize) continue;
    const float X = sh.px[0][y * kTPitch + x], Y = sh.px[1][y * kTPitch + x], Bv = sh.px[2][y * kTPitch + x];
    const float l = (Y + X) - kNegBiasCbrt, m = (Y - X) - kNegBiasCbrt, s = Bv - kNegBiasCbrt;
    const float mix0 = (l * l) * l - kBias, mix1 = (m * m) * m - kBias, mix2 = (s * s) * s - kBias;
    const uint8_t* o = A.rgb + (size_t)gy * A.stride + 3 * (size_t)gx;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const int d = srgb_code(sh.tab, k, mix0, mix1, mix2) - (int)o[k];
      sse[k] += (unsigned)(d * d);
    }
  }
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const unsigned w = __reduce_add_sync(0xffffffffu, sse[k]);
    if (lane == 0 && w) atomicAdd(A.sse + k, (unsigned long long)w);
  }
}

}  // namespace

void launch_recon_sse(const FrameDim& fd, const QuantDev* qd, const AcsTables& T, const uint16_t* const* inv_order,
                      const int8_t* cmap, float inv_qm_x, float inv_qm_b, const uint8_t* acs, const int32_t* raw_qf,
                      const int16_t* coeffs, const int16_t* dc_quant, const uint8_t* rgb, size_t stride, const float* tables,
                      unsigned long long* sse3, cudaStream_t s) {
  cudaFuncSetAttribute(k_recon_sse, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ReconShared));
  ReconArgs A;
  A.fd = fd; A.qd = qd;
  for (int i = 0; i < 17; ++i) A.dq[i] = T.dq[i];
  for (int i = 0; i < 13; ++i) A.inv_order[i] = inv_order[i];
  A.cmap = cmap; A.inv_qm_x = inv_qm_x; A.inv_qm_b = inv_qm_b; A.acs = acs; A.raw_qf = raw_qf; A.coeffs = coeffs;
  A.dc_quant = dc_quant; A.rgb = rgb; A.stride = stride; A.tables = tables; A.sse = sse3;
  ++g_kernel_launches;
  dim3 grid((fd.bxs + 3) / 4, (fd.bys + 3) / 4);
  k_recon_sse<<<grid, kReconWarps * 32, sizeof(ReconShared), s>>>(A);
}

}  // namespace jxlb
