// K7 (DCT8 frames, fourth design) — TWO THREADS PER 8x8 BLOCK, tiles staged with cp.async.
//
// History (profiles/README.md): the 8-lanes-per-block kernel (k_dct_quant.cu) spends most of its issue slots on
// shuffle transposes, shuffle reductions and replicated control flow (59.7 M warp instructions per 4K frame,
// 100 us).  One thread per block (k_dct8_v2.cu, v3) halves the instruction count but needs 255 registers and
// 768 B of shared memory per thread (256 threads per SM) and 146 KB of straight-line code: ncu shows 29 % of
// the stall samples waiting for instructions and 26 % for loads (87 us).
//
// Here a block is shared by two adjacent lanes (p = lane & 1).  A CTA owns a tile of 32 x ROWS blocks (one
// AC-group column wide); the three 8*ROWS x 256 px planes arrive as three cp.async commit groups (every warp
// copy is 512 contiguous bytes of a pixel row), so X and B are in flight while Y is transformed.
//   row pass     lane p transforms pixel rows 4p..4p+3 of its block and writes them back in place;
//   column pass  lane p reads columns 4p..4p+3 (one LDS.128 per pixel row) and transforms them, so it ends up
//                owning coefficient rows hf = 4p..4p+3 of the stored layout (index hf*8 + vf) — the top or the
//                bottom pair of quadrants of the quantisation heuristics;
//   stash        the 32 coefficients go back over the block's own pixels; the quantise pass reloads them.
// The transposition therefore costs no shuffles, and the per-quadrant sums need five shuffles per channel.
// 16-byte chunk j of tile row r is stored at chunk j ^ ((r >> 2) & 1): all three access patterns (rows 4p+k,
// both chunks / one row, chunk 2b+p / cp.async fill) are bank-conflict free.
// Scan-order packing: the two lanes own different scan positions, so each builds the 32-word image of its own
// values in a lane-specialised branch (compile-time positions), then one shuffle per word merges the halves;
// lane 0 stores scan positions 0..31, lane 1 stores 32..63 (64 contiguous bytes each).
// Arithmetic and association are the oracle's (oracle/jxo_coef.cc, jxo_dct.cc); see the exactness notes at
// v4_adjust below.  HBM: 12 B/px in, 6 B/px out (+ ~0.3 B/px side data) -> 18.3 B/px.
#include "transforms.cuh"
#include "kernels.h"

namespace jxlb {
namespace {

constexpr int kBiasN = 1024;   // entries of the Y dequantisation-bias table
constexpr int kWPitch = 36;    // floats between the p = 0 and p = 1 halves of a weight table (16-byte aligned, distinct banks)
constexpr int kTableFloats = kBiasN + 8 * kWPitch + 512;   // bias, weights (6 halves), Y dequant (2 halves), last-index LUT (2 KB)

// natural coefficient order of an 8x8 block: scan k -> position hf*8 + vf (tests pin it against the oracle)
__device__ constexpr int kScan8[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                                       41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                                       30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

__device__ __forceinline__ void cp_async16(float* smem_dst, const float* gsrc) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async4(float* smem_dst, const float* gsrc) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float a, float b, float c, float d) {
  *reinterpret_cast<float4*>(p) = make_float4(a, b, c, d);
}
__device__ __forceinline__ float f4get(const float4& v, int j) { return j == 0 ? v.x : (j == 1 ? v.y : (j == 2 ? v.z : v.w)); }

// Per-thread view of its block inside one staged plane.
struct BlockView {
  float* rows;   // first of the lane's own four tile rows (4p .. 4p+3 of the block)
  float* col;    // first tile row of the block
  int o0, o1;    // float offsets of the block's two 16-byte chunks in the lane's own rows
  int cA, cB;    // float offset of chunk 2b+p in block rows 0..3 / 4..7
};

// Forward 8x8 DCT of the lane's half of the block; c[j*8 + vf] = coefficient (hf = 4p+j, vf).  Unscaled passes and
// one exact scaling by 1/64 (the oracle scales by 1/8 after each pass; a power-of-two scale commutes with every
// rounding in between, so the bits agree).  The coefficients replace the pixels in shared memory.
__device__ __forceinline__ void v4_transform(const BlockView& bv, float c[32]) {
  {
    float4 a[4], b[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) { a[k] = ld4(bv.rows + k * 256 + bv.o0); b[k] = ld4(bv.rows + k * 256 + bv.o1); }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float* r = bv.rows + k * 256;
      float v[8] = {a[k].x, a[k].y, a[k].z, a[k].w, b[k].x, b[k].y, b[k].z, b[k].w};
      dct_rec<8>(v);
      st4(r + bv.o0, v[0], v[1], v[2], v[3]);
      st4(r + bv.o1, v[4], v[5], v[6], v[7]);
    }
  }
  __syncwarp();
  float4 m[8];
#pragma unroll
  for (int y = 0; y < 8; ++y) m[y] = ld4(bv.col + y * 256 + (y < 4 ? bv.cA : bv.cB));
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float v[8];
#pragma unroll
    for (int y = 0; y < 8; ++y) v[y] = f4get(m[y], j);
    dct_rec<8>(v);
#pragma unroll
    for (int vf = 0; vf < 8; ++vf) c[j * 8 + vf] = v[vf] * (1.0f / 64.0f);
  }
  __syncwarp();
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float* r = bv.rows + j * 256;
    st4(r + bv.o0, c[j * 8], c[j * 8 + 1], c[j * 8 + 2], c[j * 8 + 3]);
    st4(r + bv.o1, c[j * 8 + 4], c[j * 8 + 5], c[j * 8 + 6], c[j * 8 + 7]);
  }
}

__device__ __forceinline__ void v4_reload(const BlockView& bv, float c[32]) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float* r = bv.rows + j * 256;
    const float4 a = ld4(r + bv.o0), b = ld4(r + bv.o1);
    c[j * 8] = a.x; c[j * 8 + 1] = a.y; c[j * 8 + 2] = a.z; c[j * 8 + 3] = a.w;
    c[j * 8 + 4] = b.x; c[j * 8 + 5] = b.y; c[j * 8 + 6] = b.z; c[j * 8 + 7] = b.w;
  }
}

// AdjustQuantBlockAC (oracle jxo_coef.cc) for a DCT8 block entered with the default thresholds; both lanes return
// the same adjusted quant (and, for Y, the same thresholds).  Exactness of the rearrangements:
//  * the sums of |rounded value| are integer-valued and far below 2^24 for any 8-bit input, so their association
//    is free: quadrant sums replace the per-row sums + trees, and their total replaces sum_vals;
//  * the DC, which the oracle skips, enters as 0 and contributes +0 to every sum and maximum;
//  * `if (me < err) me = err` is fmaxf for the non-NaN operands that occur;
//  * the high-frequency sum keeps the oracle's association: rows 0-3 hold +0, so the tree is (r4+r6)+(r5+r7).
template <bool IS_Y>
__device__ __forceinline__ int v4_adjust(const float c[32], int p, const float* __restrict__ w, float scale, float qm_mul,
                                         float hf_mul, int quant, float thr[4]) {
  const float qac = scale * (float)quant;
  const float thr_lo = p ? 0.64f : 0.58f;
  float nzl = 0.0f, nzh = 0.0f, mel = 0.0f, meh = 0.0f;
  float hfr[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const float4 w4 = ld4(w + j * 8 + h * 4);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int x = h * 4 + e;
        float cv = c[j * 8 + x];
        if (j == 0 && x == 0) cv = p ? cv : 0.0f;
        const float wq = f4get(w4, e) * qac;
        const float val = cv * (IS_Y ? wq : wq * qm_mul);
        const float a = fabsf(val);
        const float v = (a < (x < 4 ? thr_lo : 0.64f)) ? 0.0f : rintf(val);
        if (x < 4) nzl += fabsf(v); else nzh += fabsf(v);
        if (IS_Y) {
          const float er = v == 0.0f ? fabsf(val - v) : 0.0f;
          if (x < 4) mel = fmaxf(mel, er); else meh = fmaxf(meh, er);
        }
        if ((x == 7 || j == 3) && x >= 4) hfr[j] += (p && v != 0.0f) ? a : 0.0f;
      }
    }
  }
  const unsigned full = 0xffffffffu;
  const float own_hf = (hfr[0] + hfr[2]) + (hfr[1] + hfr[3]);
  const float sum_hf = __shfl_sync(full, own_hf, (threadIdx.x & 31) | 1);
  const float onzl = __shfl_xor_sync(full, nzl, 1), onzh = __shfl_xor_sync(full, nzh, 1);
  float nz[4], me[4] = {0.0f, 0.0f, 0.0f, 0.0f};
  nz[0] = p ? onzl : nzl; nz[1] = p ? onzh : nzh; nz[2] = p ? nzl : onzl; nz[3] = p ? nzh : onzh;
  if (IS_Y) {
    const float omel = __shfl_xor_sync(full, mel, 1), omeh = __shfl_xor_sync(full, meh, 1);
    me[0] = p ? omel : mel; me[1] = p ? omeh : meh; me[2] = p ? mel : omel; me[3] = p ? meh : omeh;
  }
  const float sum_vals = nz[0] + nz[1] + nz[2] + nz[3];
  if (IS_Y) {
    if (sum_vals * 8 < 1.0f) {
      const double kLimit = 0.46, kMul = 0.9999;
      const int orig = quant;
      int nq = quant;
#pragma unroll
      for (int i = 1; i < 4; ++i) if (nq == orig && nz[i] == 0.0f && (double)me[i] > kLimit) nq = orig + 1;
      quant = nq;
      if (nz[3] == 0.0f && (double)me[3] > kLimit) {
        thr[3] = (float)(kMul * (double)me[3] * (double)nq / (double)orig);
      } else if ((nz[1] == 0.0f && (double)me[1] > kLimit) || (nz[2] == 0.0f && (double)me[2] > kLimit)) {
        const float m = me[1] > me[2] ? me[1] : me[2];
        thr[1] = (float)(kMul * (double)m * (double)nq / (double)orig);
        thr[2] = thr[1];
      } else if (nz[0] == 0.0f && (double)me[0] > kLimit) {
        thr[0] = (float)(kMul * (double)me[0] * (double)nq / (double)orig);
      }
    }
  }
  {
    const float all = nz[0] + nz[1] + nz[2] + nz[3] + 1;
    if (hf_mul * sum_hf >= all) {
      quant = (int)((float)quant + hf_mul * sum_hf / all);
      if (quant >= 256) quant = 255;
    }
  }
  if (nz[0] + nz[1] + nz[2] + nz[3] < 11) { quant += 1; if (quant >= 256) quant = 255; }
  return quant;
}

// The lane's 32-word scan-order image: words it keeps (its half of the scan) and words it sends to its partner.
template <int P>
__device__ __forceinline__ void v4_pack_half(const int q[32], uint32_t keep[16], uint32_t send[16]) {
#pragma unroll
  for (int wd = 0; wd < 32; ++wd) {
    const int plo = kScan8[2 * wd], phi = kScan8[2 * wd + 1];
    const bool own_lo = (plo >> 5) == P, own_hi = (phi >> 5) == P;
    uint32_t word = 0;
    if (own_lo || own_hi) {
      const int lo = own_lo ? q[plo & 31] : 0, hi = own_hi ? q[phi & 31] : 0;
      asm("cvt.pack.sat.s16.s32 %0, %1, %2;" : "=r"(word) : "r"(hi), "r"(lo));   // saturates to [-32768, 32767]
    }
    if ((wd < 16) == (P == 0)) keep[wd & 15] = word; else send[wd & 15] = word;
  }
}

// QuantizeBlockAC (DCT8) of the lane's 32 coefficients.  A quantised value is non-zero exactly when
// |val| >= threshold and |val| > 0.5 (round-half-even), so the non-zero mask comes from one compare.
// q holds int32 values (saturated by the conversion); the int16 clamp happens when the words are packed.
__device__ __forceinline__ uint32_t v4_quantize(const float c[32], int p, const float* __restrict__ w, float qac_mul,
                                                const float thr[4], int q[32]) {
  // |val| > 0.5 is |val| >= the next float above 0.5, so both conditions fold into one threshold per quadrant
  const float kAboveHalf = 0.50000006f;
  const float t_lo = fmaxf(p ? thr[2] : thr[0], kAboveHalf), t_hi = fmaxf(p ? thr[3] : thr[1], kAboveHalf);
  uint32_t mask = 0;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const float4 w4 = ld4(w + j * 8 + h * 4);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int x = h * 4 + e;
        const float val = (f4get(w4, e) * qac_mul) * c[j * 8 + x];
        const float a = fabsf(val);
        bool nonzero = a >= (h == 0 ? t_lo : t_hi);
        if (j == 0 && x == 0) nonzero = nonzero && (p != 0);
        q[j * 8 + x] = nonzero ? __float2int_rn(val) : 0;
        if (nonzero) mask |= 1u << (j * 8 + x);
      }
    }
  }
  return mask;
}

// merges the two lanes' images, applies the oracle's [-32767, 32767] clamp and stores the lane's 64 bytes;
// returns through nz / last the block's non-zero count and last non-zero scan index (valid on both lanes)
__device__ __forceinline__ void v4_emit(const int q[32], uint32_t mask, int p, const uint8_t* __restrict__ last_lut, bool active,
                                        int16_t* __restrict__ dst, int& nz, int& last) {
  const unsigned full = 0xffffffffu;
  uint32_t keep[16], send[16];
  if (p == 0) v4_pack_half<0>(q, keep, send); else v4_pack_half<1>(q, keep, send);
  int own_last = 0;
#pragma unroll
  for (int b = 0; b < 4; ++b) own_last = max(own_last, (int)last_lut[(p * 4 + b) * 256 + ((mask >> (8 * b)) & 255u)]);
  const int own_nz = __popc(mask);
  nz = own_nz + __shfl_xor_sync(full, own_nz, 1);
  last = max(own_last, __shfl_xor_sync(full, own_last, 1));
#pragma unroll
  for (int i = 0; i < 16; ++i) keep[i] = __vimax3_s16x2(keep[i] | __shfl_xor_sync(full, send[i], 1), 0x80018001u, 0x80018001u);
  if (active) {
    uint4* d4 = reinterpret_cast<uint4*>(dst + p * 32);
#pragma unroll
    for (int i = 0; i < 4; ++i) d4[i] = make_uint4(keep[4 * i], keep[4 * i + 1], keep[4 * i + 2], keep[4 * i + 3]);
  }
}

__device__ __forceinline__ float v4_bias_formula(int aq) {
  if (aq == 0) return 0.0f;
  if (aq == 1) return 1.0f - 0.07005449891748593f;
  const float fq = (float)aq;
  return fq - 0.145f / fq;
}

template <int ROWS, int TPS>   // block rows per CTA, resident threads per SM the register allocation aims at
__global__ void __launch_bounds__(64 * ROWS, TPS / (64 * ROWS)) k_dct8_quant_v4(
    const float* __restrict__ X, const float* __restrict__ Y, const float* __restrict__ B, FrameDim fd,
    const QuantDev* __restrict__ qd, const float* __restrict__ weights, const float* __restrict__ dequant_y,
    const float* __restrict__ bias_tab, const uint8_t* __restrict__ last_lut, const int8_t* __restrict__ cmap, float x_qm_mul,
    float b_qm_mul, int adjust, int32_t* __restrict__ raw_qf, int16_t* __restrict__ coeffs, int16_t* __restrict__ dc_quant,
    uint8_t* __restrict__ nzeros, uint16_t* __restrict__ nzcount, uint16_t* __restrict__ lastk) {
  extern __shared__ __align__(16) float s_mem[];
  constexpr int kThreads = 64 * ROWS;
  constexpr int kPlane = ROWS * 8 * 256;
  constexpr int kSmemLut = 3 * kPlane + kBiasN + 8 * kWPitch;   // 16-byte aligned
  float* const s_tile = s_mem;                       // [3 (Y, X, B)][ROWS * 8][256]
  float* const s_bias = s_tile + 3 * kPlane;         // [kBiasN]
  float* const s_w = s_bias + kBiasN;                // [3 (X, Y, B)][2 (p)][kWPitch]
  float* const s_dq = s_w + 6 * kWPitch;             // [2 (p)][kWPitch]   Y dequant
  uint8_t* const s_lut = reinterpret_cast<uint8_t*>(s_mem + kSmemLut);   // [2 (p)][4][256] last scan index per mask byte
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int p = lane & 1, bc = (warp & 1) * 16 + (lane >> 1), br = warp >> 1;
  const int gx = blockIdx.x, by0 = blockIdx.y * ROWS;
  // ---- stage the lookup tables and the three planes (Y first); nothing here waits for a load
  for (int k = t; k < kBiasN / 4; k += kThreads) cp_async16(s_bias + 4 * k, bias_tab + 4 * k);
  for (int k = t; k < 128; k += kThreads) cp_async16(s_mem + kSmemLut + 4 * k, reinterpret_cast<const float*>(last_lut) + 4 * k);
  for (int k = t; k < 192; k += kThreads) cp_async4(s_w + (k >> 5) * kWPitch + (k & 31), weights + k);
  for (int k = t; k < 64; k += kThreads) cp_async4(s_dq + (k >> 5) * kWPitch + (k & 31), dequant_y + k);
#pragma unroll
  for (int slot = 0; slot < 3; ++slot) {
    const float* __restrict__ P = slot == 0 ? Y : (slot == 1 ? X : B);
#pragma unroll
    for (int i = 0; i < (ROWS * 8 * 64) / kThreads; ++i) {
      const int idx = t + i * kThreads;
      const int row = idx >> 6, ck = idx & 63;
      const int py = min(by0 * 8 + row, fd.ys_pad - 1);
      const int px = min(gx * 256 + ck * 4, fd.pitch - 4);
      cp_async16(s_tile + slot * kPlane + row * 256 + ((ck ^ ((row >> 2) & 1)) << 2), P + (size_t)py * fd.pitch + px);
    }
    asm volatile("cp.async.commit_group;\n" ::: "memory");
  }
  const int bx = gx * 32 + bc, by = by0 + br;
  const bool active = bx < fd.bxs && by < fd.bys;
  const size_t nblk = (size_t)fd.bxs * fd.bys;
  const size_t bi = active ? (size_t)by * fd.bxs + bx : 0;
  const float scale = qd->scale, inv_gs = qd->inv_global_scale;
  const int quant_dc = qd->quant_dc;
  int quant = raw_qf[bi];
  const int tx = min(bx, fd.bxs - 1) >> 3, ty = min(by, fd.bys - 1) >> 3;
  const int cmap_x = cmap[(size_t)ty * fd.txs + tx], cmap_b = cmap[(size_t)fd.txs * fd.tys + (size_t)ty * fd.txs + tx];
  // the lane's view of its block in plane `slot` (0 = Y, 1 = X, 2 = B)
  auto view = [&](int slot) {
    BlockView v;
    v.col = s_tile + slot * kPlane + br * 8 * 256;
    v.rows = v.col + 4 * p * 256;
    v.o0 = ((2 * bc) ^ p) << 2;
    v.o1 = ((2 * bc + 1) ^ p) << 2;
    v.cA = (2 * bc + p) << 2;
    v.cB = ((2 * bc + p) ^ 1) << 2;
    return v;
  };
  // weight tables: s_w holds X, Y, B halves; the chroma passes below run as rolled two-trip loops (k = 0: X, 1: B)
  // so that their code is shared: the kernel's footprint in the instruction cache is what limits issue otherwise
  const float* const wY = s_w + (2 + p) * kWPitch;
  const float* const dqY = s_dq + p * kWPitch;
  float thr_y[4] = {0.58f, 0.64f, 0.64f, 0.64f};
  float c[32];
  // ---- pass A: transform, stash, quant adjust (the maximum over the channels wins; Y's thresholds are kept)
  {
    const int orig = quant;
    asm volatile("cp.async.wait_group 2;\n" ::: "memory");
    __syncthreads();
    v4_transform(view(0), c);
    int maxq = v4_adjust<true>(c, p, wY, scale, 1.0f, 30.0f, orig, thr_y);
#pragma unroll 1
    for (int k = 0; k < 2; ++k) {
      if (k == 0) asm volatile("cp.async.wait_group 1;\n" ::: "memory");
      else asm volatile("cp.async.wait_group 0;\n" ::: "memory");
      __syncthreads();
      float thr[4] = {0.58f, 0.64f, 0.64f, 0.64f};
      v4_transform(view(1 + k), c);
      maxq = max(maxq, v4_adjust<false>(c, p, s_w + (4 * k + p) * kWPitch, scale, k ? b_qm_mul : x_qm_mul, k ? 60.0f : 70.0f,
                                        orig, thr));
    }
    if (adjust) {
      quant = maxq;
    } else {
      thr_y[0] = 0.56f; thr_y[1] = thr_y[2] = thr_y[3] = 0.62f;
    }
  }
  // ---- pass B: quantise
  const float qac = scale * (float)quant;
  const float inv_qac = inv_gs / (float)quant;
  const float x_factor = 0.0f + (float)cmap_x / 84.0f;
  const float b_factor = 1.0f + (float)cmap_b / 84.0f;
  const int g = (by >> 5) * fd.gxs + gx;
  int16_t* dst = coeffs + ((size_t)g * kGroupBlocks + (size_t)(by & 31) * 32 + bc) * 192;
  const bool writer = active && p == 0;
  int q[32];
  int nz, last;
  float dcv[3];
  float yrt[32];
  {
    v4_reload(view(0), c);
    dcv[1] = c[0];
    const uint32_t mask = v4_quantize(c, p, wY, qac * 1.0f, thr_y, q);
    v4_emit(q, mask, p, s_lut, active, dst, nz, last);
    if (writer) { nzeros[nblk + bi] = (uint8_t)nz; nzcount[nblk + bi] = (uint16_t)nz; lastk[nblk + bi] = (uint16_t)last; }
    // dequantised Y for the chroma-from-luma term: +-bias[|q|] from the table.  A block holding a |q| beyond the
    // table (possible at very small distances) takes the formula everywhere, with q clamped like the stored int16.
    unsigned big = 0;
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      const unsigned aq = (unsigned)abs(q[i]);
      big = max(big, aq);
      const float m = s_bias[min(aq, (unsigned)(kBiasN - 1))];
      const float sm = __uint_as_float(__float_as_uint(m) | ((uint32_t)q[i] & 0x80000000u));
      yrt[i] = (sm * dqY[i]) * inv_qac;
    }
    if (big >= (unsigned)kBiasN) {
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const int qi = max(min(q[i], 32767), -32767);
        const float m = v4_bias_formula(abs(qi));
        yrt[i] = ((qi < 0 ? -m : m) * dqY[i]) * inv_qac;
      }
    }
  }
  const float thr0[4] = {0.58f, 0.64f, 0.64f, 0.64f};
#pragma unroll 1
  for (int k = 0; k < 2; ++k) {
    v4_reload(view(1 + k), c);
    if (k == 0) dcv[0] = c[0]; else dcv[2] = c[0];
    const float factor = k ? b_factor : x_factor;
#pragma unroll
    for (int i = 0; i < 32; ++i) c[i] = __fmaf_rn(-factor, yrt[i], c[i]);
    const uint32_t mask = v4_quantize(c, p, s_w + (4 * k + p) * kWPitch, qac * (k ? b_qm_mul : x_qm_mul), thr0, q);
    v4_emit(q, mask, p, s_lut, active, dst + 64 + 64 * k, nz, last);
    const size_t si = (size_t)(2 * k) * nblk + bi;
    if (writer) { nzeros[si] = (uint8_t)nz; nzcount[si] = (uint16_t)nz; lastk[si] = (uint16_t)last; }
  }
  // ---- DC (AddVarDCTDC) + side data, from the lane that holds coefficient 0
  if (writer) {
    raw_qf[bi] = quant;
    const float gsq = scale * (float)quant_dc;
    const float inv_quant_dc = inv_gs / (float)quant_dc;
    const float y_factor = inv_quant_dc * (1.0f / 512.0f);
    const float qy = roundf(dcv[1] * (512.0f * gsq));
    const float qx = roundf((dcv[0] - qy * (y_factor * 0.0f)) * (4096.0f * gsq));
    const float qb = roundf((dcv[2] - qy * (y_factor * 1.0f)) * (256.0f * gsq));
    const int iy = (int)qy, ix = (int)qx, ib = (int)qb;
    dc_quant[0 * nblk + bi] = (int16_t)(ix > 32767 ? 32767 : (ix < -32768 ? -32768 : ix));
    dc_quant[1 * nblk + bi] = (int16_t)(iy > 32767 ? 32767 : (iy < -32768 ? -32768 : iy));
    dc_quant[2 * nblk + bi] = (int16_t)(ib > 32767 ? 32767 : (ib < -32768 ? -32768 : ib));
  }
}

template <int ROWS, int TPS>
void launch_v4(const float* x, const float* y, const float* b, const FrameDim& fd, const QuantDev* qd, const float* weights,
               const float* dequant_y, const float* bias_tab, const uint8_t* last_lut, const int8_t* cmap, float x_qm_mul,
               float b_qm_mul, int adjust, int32_t* raw_qf, int16_t* coeffs, int16_t* dc_quant, uint8_t* nzeros, uint16_t* nzcount,
               uint16_t* lastk, cudaStream_t s) {
  constexpr size_t smem = ((size_t)3 * ROWS * 8 * 256 + kTableFloats) * sizeof(float);
  cudaFuncSetAttribute(k_dct8_quant_v4<ROWS, TPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  dim3 grid(fd.gxs, (fd.bys + ROWS - 1) / ROWS);
  k_dct8_quant_v4<ROWS, TPS><<<grid, 64 * ROWS, smem, s>>>(x, y, b, fd, qd, weights, dequant_y, bias_tab, last_lut, cmap, x_qm_mul,
                                                     b_qm_mul, adjust, raw_qf, coeffs, dc_quant, nzeros, nzcount, lastk);
}

}  // namespace

int dct8_v4_bias_entries() { return kBiasN; }

// host side of the two lookup tables: bias[k] (AdjustQuantBias of Y for |q| = k) and, per lane half p and byte b of
// the lane's non-zero mask, the largest scan index among the set bits (coefficient (4p+b)*8 + bit)
void dct8_v4_host_tables(const uint8_t* izz64, float* bias, uint8_t* last_lut) {
  for (int k = 0; k < kBiasN; ++k) {
    const float fk = (float)k;
    bias[k] = k == 0 ? 0.0f : (k == 1 ? 1.0f - 0.07005449891748593f : fk - 0.145f / fk);
  }
  for (int p = 0; p < 2; ++p)
    for (int b = 0; b < 4; ++b)
      for (int v = 0; v < 256; ++v) {
        int m = 0;
        for (int bit = 0; bit < 8; ++bit)
          if (v >> bit & 1) m = izz64[(4 * p + b) * 8 + bit] > m ? izz64[(4 * p + b) * 8 + bit] : m;
        last_lut[(p * 4 + b) * 256 + v] = (uint8_t)m;
      }
}

void launch_dct8_quant_v4(const float* x, const float* y, const float* b, const FrameDim& fd, const QuantDev* qd,
                          const float* weights, const float* dequant_y, const float* bias_tab, const uint8_t* last_lut,
                          const int8_t* cmap, float x_qm_mul, float b_qm_mul, int adjust, int rows_per_cta, int threads_per_sm,
                          int32_t* raw_qf, int16_t* coeffs, int16_t* dc_quant, uint8_t* nzeros, uint16_t* nzcount, uint16_t* lastk,
                          cudaStream_t s) {
  ++g_kernel_launches;
#define V4_ARGS x, y, b, fd, qd, weights, dequant_y, bias_tab, last_lut, cmap, x_qm_mul, b_qm_mul, adjust, raw_qf, coeffs, dc_quant, \
                nzeros, nzcount, lastk, s
  if (rows_per_cta == 2) {
    if (threads_per_sm >= 512) launch_v4<2, 512>(V4_ARGS); else launch_v4<2, 384>(V4_ARGS);
  } else {
    if (threads_per_sm >= 512) launch_v4<4, 512>(V4_ARGS); else launch_v4<4, 256>(V4_ARGS);
  }
#undef V4_ARGS
}

}  // namespace jxlb
