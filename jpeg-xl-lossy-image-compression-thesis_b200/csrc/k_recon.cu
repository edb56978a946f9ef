// K13 — reconstruction error of the frame that was just coded (the quantity behind
// calculate_mse / calculate_psnr of the harness, benchmark-jpegxl/src/image_reader.rs:555-606, taken from the
// integers the codestream carries instead of a djxl round trip): dequantise the stored coefficients, put the
// lowest frequencies back from the DC image, chroma-from-luma, inverse transform, XYB -> linear -> 8-bit sRGB,
// squared error against the input pixels.  Same arithmetic and operation order as oracle/jxo_recon.cc
// (ReconstructRgb / ReconstructionSse), so the three 64-bit sums are bit-exact.
//
// Mirror of k_coeff.cu: the per-strategy lists of first blocks are reused; 8x8 strategies are inverted by one thread
// per block in registers, larger ones by one lane group per transform (SquareXform::inverse).  The reconstructed XYB
// planes go to a scratch frame; k_recon_sse turns them into sRGB codes and sums the squared differences.
// Enabled by JXLB200_FLAG_QUALITY.
#include "transforms.cuh"
#include "kernels.h"

namespace jxlb {
namespace {

enum { kListDCT = 0, kListID, kList2X2, kList4X4, kList4X8, kList8X4, kList16Tall, kList16Wide, kList16Sq, kList32Tall, kList32Wide,
       kList32Sq, kList64Tall, kList64Wide, kList64Sq, kList32Tall4, kList32Wide4, kNumLists };
constexpr int kListHeader = 32;

struct ReconArgs {
  FrameDim fd;
  const QuantDev* qd;
  const float* dq;                  // dequantisation table of the strategy, lane order ([hf][vf])
  const uint16_t* inv;              // coefficient position (lane order) -> scan index
  const int8_t* cmap;
  float inv_qm_x, inv_qm_b;
  const int32_t* raw_qf;
  const int16_t* coeffs;
  const int16_t* dc_quant;
  float* out;                       // 3 planes of ys_pad x pitch floats
  const uint32_t* list; const uint32_t* count;
};

__device__ __forceinline__ float dequant_bias(int c, int q) {
  const float b0 = 1.0f - 0.05465007330715401f, b1 = 1.0f - 0.07005449891748593f, b2 = 1.0f - 0.049935103337343655f;
  const float bc = c == 0 ? b0 : (c == 1 ? b1 : b2);
  if (q == 0) return 0.0f;
  if (q == 1) return bc;
  if (q == -1) return -bc;
  return (float)q - 0.145f / (float)q;
}

// dequantised DC of block bj, channel c (Y first; X / B with the default 0 / 1.0 DC correlation)
__device__ __forceinline__ float dc_value(const ReconArgs& A, size_t nblk, size_t bj, int c) {
  const float inv_quant_dc = A.qd->inv_global_scale / (float)A.qd->quant_dc;
  const float y = (float)A.dc_quant[nblk + bj] * (inv_quant_dc * (1.0f / 512.0f));
  if (c == 1) return y;
  if (c == 0) return __fmaf_rn(0.0f, y, (float)A.dc_quant[bj] * (inv_quant_dc * (1.0f / 4096.0f)));
  return __fmaf_rn(1.0f, y, (float)A.dc_quant[2 * nblk + bj] * (inv_quant_dc * (1.0f / 256.0f)));
}

struct Zigzag8 { int pos[64]; };
constexpr Zigzag8 make_zigzag8() {
  Zigzag8 z{};
  int cur = 1;
  for (int i = 0; i < 8; ++i)
    for (int j = 0; j <= i; ++j) {
      int x = j, y = i - j;
      if (i % 2) { const int t = x; x = y; y = t; }
      const int val = (x < 1 && y < 1) ? 0 : cur++;
      z.pos[val] = y * 8 + x;
    }
  for (int ip = 7; ip > 0; --ip) {
    const int i = ip - 1;
    for (int j = 0; j <= i; ++j) {
      int x = 7 - (i - j), y = 7 - j;
      if (i % 2) { const int t = x; x = y; y = t; }
      z.pos[cur++] = y * 8 + x;
    }
  }
  return z;
}
__device__ constexpr Zigzag8 kZigzag8 = make_zigzag8();

template <int S>
__global__ void __launch_bounds__(64) k_recon8(ReconArgs A) {
  const FrameDim& fd = A.fd;
  const unsigned n = *A.count;
  const unsigned i = blockIdx.x * 64 + threadIdx.x;
  if (i >= n) return;
  const size_t nblk = (size_t)fd.bxs * fd.bys, plane = (size_t)fd.ys_pad * fd.pitch;
  const size_t bi = A.list[i];
  const int bx = (int)(bi % fd.bxs), by = (int)(bi / fd.bxs);
  const float inv_qac = A.qd->inv_global_scale / (float)A.raw_qf[bi];
  const int g = (by >> 5) * fd.gxs + (bx >> 5);
  const size_t cblk = (size_t)g * kGroupBlocks + (size_t)(by & 31) * 32 + (bx & 31);
  const int tx = bx >> 3, ty = by >> 3;
  const float cfl_x = 0.0f + (float)A.cmap[(size_t)ty * fd.txs + tx] / 84.0f;
  const float cfl_b = 1.0f + (float)A.cmap[(size_t)fd.txs * fd.tys + (size_t)ty * fd.txs + tx] / 84.0f;
  float ydq[64];
#pragma unroll 1
  for (int it = 0; it < 3; ++it) {
    const int c = it == 0 ? 1 : (it == 1 ? 0 : 2), slot = it;
    const float mul = inv_qac * (c == 0 ? A.inv_qm_x : (c == 1 ? 1.0f : A.inv_qm_b));
    const float cfl = c == 0 ? cfl_x : cfl_b;
    const uint4* src = reinterpret_cast<const uint4*>(A.coeffs + (cblk * 3 + slot) * 64);
    uint32_t words[32];
#pragma unroll
    for (int k = 0; k < 8; ++k) { const uint4 q4 = __ldg(src + k); words[4 * k] = q4.x; words[4 * k + 1] = q4.y; words[4 * k + 2] = q4.z; words[4 * k + 3] = q4.w; }
    float cf[64];
#pragma unroll
    for (int k = 0; k < 64; ++k) {
      const int q = (int)(int16_t)((k & 1) ? (words[k >> 1] >> 16) : (words[k >> 1] & 0xFFFFu));
      const int pos = kZigzag8.pos[k];
      float v = (dequant_bias(c, q) * __ldg(A.dq + c * 64 + pos)) * mul;
      if (c != 1) v = __fmaf_rn(cfl, ydq[pos], v);
      cf[pos] = v;
    }
    if (it == 0) {
#pragma unroll
      for (int k = 0; k < 64; ++k) ydq[k] = cf[k];
    }
    cf[0] = dc_value(A, nblk, bi, c);
    float p[64];
    inv8x8<S>(cf, p);
    float* dst = A.out + (size_t)c * plane + (size_t)by * 8 * fd.pitch + (size_t)bx * 8;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      *reinterpret_cast<float4*>(dst + (size_t)r * fd.pitch) = make_float4(p[r * 8], p[r * 8 + 1], p[r * 8 + 2], p[r * 8 + 3]);
      *reinterpret_cast<float4*>(dst + (size_t)r * fd.pitch + 4) = make_float4(p[r * 8 + 4], p[r * 8 + 5], p[r * 8 + 6], p[r * 8 + 7]);
    }
  }
}

template <int N, int MODE> struct ReconGeom {
  static constexpr bool kWide = MODE == kModeWide2 || MODE == kModeWide4;
  static constexpr int SHORT = MODE == kModeSq ? N : ((MODE == kModeTall2 || MODE == kModeWide2) ? N / 2 : N / 4);
  static constexpr int LANES = kWide ? N : SHORT;
  static constexpr int VALS = kWide ? SHORT : N;
  static constexpr int R = kWide ? SHORT : N, C = kWide ? N : SHORT;
  static constexpr int cxb = C / 8, cyb = R / 8;
};
template <int N> struct ReconSqGeom {
  static constexpr int kGroupsPerWarp = N == 16 ? 2 : 1;
  static constexpr int kThreads = N == 64 ? 64 : 128;
  static constexpr int kGroups = (N == 64 ? 1 : 4) * kGroupsPerWarp;
  static constexpr int kSq = N == 16 ? 256 + 16 : N * N;
  static constexpr int kGroupFloats = kSq + 64;
};

template <int N, int MODE>
__global__ void __launch_bounds__(ReconSqGeom<N>::kThreads) k_reconsq(ReconArgs A) {
  using G = ReconGeom<N, MODE>;
  using RG = ReconSqGeom<N>;
  using SX = SquareXform<N>;
  extern __shared__ __align__(16) float smem_f[];
  const FrameDim& fd = A.fd;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int l = N == 64 ? tid : (lane & (N - 1));
  const int grp = N == 64 ? 0 : warp * RG::kGroupsPerWarp + (N == 16 ? (lane >> 4) : 0);
  float* tbuf = smem_f + grp * RG::kGroupFloats;
  float* llf = tbuf + RG::kSq;
  const typename SX::Col col = SX::col_of(l);
  const unsigned n = *A.count;
  const size_t nblk = (size_t)fd.bxs * fd.bys, plane = (size_t)fd.ys_pad * fd.pitch;
  const bool owner = l < G::LANES;
  constexpr int bar_id = 1;
  for (unsigned item0 = blockIdx.x * RG::kGroups; item0 < n; item0 += gridDim.x * RG::kGroups) {
    const unsigned item = item0 + grp;
    const bool active = item < n;
    const size_t bi = active ? A.list[item] : 0;
    const int bx = (int)(bi % fd.bxs), by = (int)(bi / fd.bxs);
    const float inv_qac = A.qd->inv_global_scale / (float)(active ? A.raw_qf[bi] : 1);
    const int tx = bx >> 3, ty = by >> 3;
    const float cfl_x = 0.0f + (float)A.cmap[(size_t)ty * fd.txs + tx] / 84.0f;
    const float cfl_b = 1.0f + (float)A.cmap[(size_t)fd.txs * fd.tys + (size_t)ty * fd.txs + tx] / 84.0f;
    const uint16_t* invrow = A.inv + (size_t)l * G::VALS;
    float uY[G::VALS];
#pragma unroll 1
    for (int it = 0; it < 3; ++it) {
      const int c = it == 0 ? 1 : (it == 1 ? 0 : 2), slot = it;
      const float mul = inv_qac * (c == 0 ? A.inv_qm_x : (c == 1 ? 1.0f : A.inv_qm_b));
      const float cfl = c == 0 ? cfl_x : cfl_b;
      const float* dqrow = A.dq + (size_t)c * G::LANES * G::VALS + (size_t)l * G::VALS;
      float u[N], v[N];
#pragma unroll
      for (int j = 0; j < N; ++j) u[j] = 0.0f;
      if (owner) {
#pragma unroll
        for (int j4 = 0; j4 < G::VALS; j4 += 4) {
          const uint2 i4 = __ldg(reinterpret_cast<const uint2*>(invrow + j4));
          const uint32_t iv[4] = {i4.x & 0xFFFFu, i4.x >> 16, i4.y & 0xFFFFu, i4.y >> 16};
          const float4 d4 = __ldg(reinterpret_cast<const float4*>(dqrow + j4));
          const float dv[4] = {d4.x, d4.y, d4.z, d4.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int j = j4 + e;
            const int k = (int)iv[e];
            const int jj = k >> 6;
            const int cbx = bx + (jj % G::cxb), cby = by + (jj / G::cxb);
            const int g = (cby >> 5) * fd.gxs + (cbx >> 5);
            const size_t blk = (size_t)g * kGroupBlocks + (size_t)(cby & 31) * 32 + (cbx & 31);
            const int q = active ? (int)A.coeffs[(blk * 3 + slot) * 64 + (k & 63)] : 0;
            float val = (dequant_bias(c, q) * dv[e]) * mul;
            if (c != 1) val = __fmaf_rn(cfl, uY[j], val);
            u[j] = val;
          }
        }
      }
      if (it == 0) {
#pragma unroll
        for (int j = 0; j < G::VALS; ++j) uY[j] = u[j];
      }
      // lowest frequencies from the dequantised DC of the covered blocks (oracle LowestFrequenciesFromDc): lane y < cyb
      // transforms DC row y horizontally, lane hf < cxb then transforms column hf vertically and rescales
      if (l < G::cyb) {
        float t[G::cxb];
#pragma unroll
        for (int x = 0; x < G::cxb; ++x) t[x] = active ? dc_value(A, nblk, bi + (size_t)l * fd.bxs + x, c) : 0.0f;
        dct1d<G::cxb>(t);
#pragma unroll
        for (int x = 0; x < G::cxb; ++x) llf[l * G::cxb + x] = t[x];
      }
      SX::sync(bar_id);
      if (l < G::cxb) {
        float t[G::cyb];
#pragma unroll
        for (int y = 0; y < G::cyb; ++y) t[y] = llf[y * G::cxb + l];
        dct1d<G::cyb>(t);
#pragma unroll
        for (int vf = 0; vf < G::cyb; ++vf) u[vf] = t[vf] / (resample_scale(G::R, vf) * resample_scale(G::C, l));
      }
      SX::template inverse<MODE>(tbuf, l, col, u, v, bar_id);
      if (active && l < G::R) {
        float* dst = A.out + (size_t)c * plane + (size_t)(by * 8 + l) * fd.pitch + (size_t)bx * 8;
#pragma unroll
        for (int j = 0; j < G::C / 4; ++j) *reinterpret_cast<float4*>(dst + 4 * j) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
      }
      SX::sync(bar_id);   // the square and the LLF scratch are rewritten by the next channel
    }
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

// the decoder's default Gaborish blur of pixel (x, y) of one plane, mirrored at the image borders (oracle GaborishBlurAt)
__device__ __forceinline__ int recon_mirror(int i, int n) {
  while (i < 0 || i >= n) i = i < 0 ? -i - 1 : 2 * n - 1 - i;
  return i;
}
__device__ __forceinline__ float gab_blur_at(const float* __restrict__ pl, const FrameDim& fd, int x, int y) {
  const float w1 = 0.115169525f, w2 = 0.061248592f;
  const float norm = 1.0f / (1.0f + 4.0f * w1 + 4.0f * w2);
  const float wc = norm, we = w1 * norm, wd = w2 * norm;
  const int xm = recon_mirror(x - 1, fd.xsize), xp = recon_mirror(x + 1, fd.xsize);
  const size_t r0 = (size_t)recon_mirror(y - 1, fd.ysize) * fd.pitch, r1 = (size_t)y * fd.pitch, r2 = (size_t)recon_mirror(y + 1, fd.ysize) * fd.pitch;
  const float s1 = (pl[r1 + xm] + pl[r1 + xp]) + (pl[r0 + x] + pl[r2 + x]);
  const float s2 = (pl[r0 + xm] + pl[r0 + xp]) + (pl[r2 + xm] + pl[r2 + xp]);
  return __fmaf_rn(wd, s2, __fmaf_rn(we, s1, wc * pl[r1 + x]));
}

__global__ void __launch_bounds__(256) k_recon_sse(const float* __restrict__ xyb, FrameDim fd, const uint8_t* __restrict__ rgb, size_t stride,
                                                   const float* __restrict__ tables, unsigned long long* __restrict__ sse3, int gab) {
  __shared__ float tab[9 + 255];
  for (int i = threadIdx.x; i < 9 + 255; i += 256) tab[i] = tables[i];
  __syncthreads();
  const size_t plane = (size_t)fd.ys_pad * fd.pitch;
  const float kBias = 0.0037930732552754493f, kNegBiasCbrt = -0.15595420054924863f;
  unsigned sse[3] = {0u, 0u, 0u};
  const int gy = blockIdx.y;
  for (int gx = blockIdx.x * 256 + threadIdx.x; gx < fd.xsize; gx += gridDim.x * 256) {
    const size_t p = (size_t)gy * fd.pitch + gx;
    float X = xyb[p], Y = xyb[plane + p], Bv = xyb[2 * plane + p];
    if (gab) { X = gab_blur_at(xyb, fd, gx, gy); Y = gab_blur_at(xyb + plane, fd, gx, gy); Bv = gab_blur_at(xyb + 2 * plane, fd, gx, gy); }
    const float l = (Y + X) - kNegBiasCbrt, m = (Y - X) - kNegBiasCbrt, s = Bv - kNegBiasCbrt;
    const float mix0 = (l * l) * l - kBias, mix1 = (m * m) * m - kBias, mix2 = (s * s) * s - kBias;
    const uint8_t* o = rgb + (size_t)gy * stride + 3 * (size_t)gx;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const int d = srgb_code(tab, k, mix0, mix1, mix2) - (int)o[k];
      sse[k] += (unsigned)(d * d);
    }
  }
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const unsigned w = __reduce_add_sync(0xffffffffu, sse[k]);
    if (lane == 0 && w) atomicAdd(sse3 + k, (unsigned long long)w);
  }
}

template <int S>
void launch_recon8(ReconArgs A, int list_id, int kind, const uint32_t* lists, const AcsTables& T, size_t nblk, cudaStream_t s) {
  A.dq = T.dq[kind]; A.inv = nullptr;
  A.count = lists + list_id; A.list = lists + kListHeader + (size_t)list_id * nblk;
  ++g_kernel_launches;
  k_recon8<S><<<(unsigned)((nblk + 63) / 64), 64, 0, s>>>(A);
}
template <int N, int MODE>
void launch_reconsq(ReconArgs A, int list_id, const float* dq, const uint16_t* inv, const uint32_t* lists, size_t nblk, cudaStream_t s) {
  using RG = ReconSqGeom<N>;
  A.dq = dq; A.inv = inv;
  A.count = lists + list_id; A.list = lists + kListHeader + (size_t)list_id * nblk;
  const size_t smem = (size_t)RG::kGroups * RG::kGroupFloats * sizeof(float);
  cudaFuncSetAttribute(k_reconsq<N, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const size_t ncov = (size_t)ReconGeom<N, MODE>::cxb * ReconGeom<N, MODE>::cyb;
  size_t grid = (nblk / ncov + RG::kGroups) / RG::kGroups;
  if (grid > 148 * 8) grid = 148 * 8;
  ++g_kernel_launches;
  k_reconsq<N, MODE><<<(unsigned)grid, RG::kThreads, smem, s>>>(A);
}

}  // namespace

void launch_recon_sse(const FrameDim& fd, const QuantDev* qd, const AcsTables& T, const uint16_t* const* inv_order,
                      const int8_t* cmap, float inv_qm_x, float inv_qm_b, const uint8_t* acs, const int32_t* raw_qf,
                      const int16_t* coeffs, const int16_t* dc_quant, const uint8_t* rgb, size_t stride, const float* tables,
                      const uint32_t* lists, float* scratch_xyb, unsigned long long* sse3, int gab, cudaStream_t s) {
  (void)acs;
  const size_t nblk = (size_t)fd.bxs * fd.bys;
  ReconArgs A;
  A.fd = fd; A.qd = qd; A.cmap = cmap; A.inv_qm_x = inv_qm_x; A.inv_qm_b = inv_qm_b; A.raw_qf = raw_qf; A.coeffs = coeffs;
  A.dc_quant = dc_quant; A.out = scratch_xyb; A.dq = nullptr; A.inv = nullptr; A.list = nullptr; A.count = nullptr;
  launch_recon8<kStratDCT>(A, kListDCT, 0, lists, T, nblk, s);
  launch_recon8<kStratIDENTITY>(A, kListID, 1, lists, T, nblk, s);
  launch_recon8<kStratDCT2X2>(A, kList2X2, 2, lists, T, nblk, s);
  launch_recon8<kStratDCT4X4>(A, kList4X4, 3, lists, T, nblk, s);
  launch_recon8<kStratDCT4X8>(A, kList4X8, 9, lists, T, nblk, s);
  launch_recon8<kStratDCT8X4>(A, kList8X4, 9, lists, T, nblk, s);
  launch_reconsq<16, kModeTall2>(A, kList16Tall, T.dq[6], inv_order[4], lists, nblk, s);
  launch_reconsq<16, kModeWide2>(A, kList16Wide, T.dqT[6], inv_order[13], lists, nblk, s);
  launch_reconsq<16, kModeSq>(A, kList16Sq, T.dq[4], inv_order[2], lists, nblk, s);
  launch_reconsq<32, kModeTall2>(A, kList32Tall, T.dq[8], inv_order[6], lists, nblk, s);
  launch_reconsq<32, kModeWide2>(A, kList32Wide, T.dqT[8], inv_order[14], lists, nblk, s);
  launch_reconsq<32, kModeSq>(A, kList32Sq, T.dq[5], inv_order[3], lists, nblk, s);
  launch_reconsq<64, kModeTall2>(A, kList64Tall, T.dq[12], inv_order[8], lists, nblk, s);
  launch_reconsq<64, kModeWide2>(A, kList64Wide, T.dqT[12], inv_order[15], lists, nblk, s);
  launch_reconsq<64, kModeSq>(A, kList64Sq, T.dq[11], inv_order[7], lists, nblk, s);
  launch_reconsq<32, kModeTall4>(A, kList32Tall4, T.dq[7], inv_order[5], lists, nblk, s);
  launch_reconsq<32, kModeWide4>(A, kList32Wide4, T.dqT[7], inv_order[16], lists, nblk, s);
  ++g_kernel_launches;
  k_recon_sse<<<dim3((fd.xsize + 255) / 256, fd.ysize), 256, 0, s>>>(scratch_xyb, fd, rgb, stride, tables, sse3, gab);
}

}  // namespace jxlb
