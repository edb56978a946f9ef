// K7 (general) — forward transform, DC extraction and adaptive quantisation for every AC strategy the search can pick
// (stage U5: libjxl enc_group.cc ComputeCoefficients / AdjustQuantBlockAC / QuantizeBlockAC / QuantizeRoundtripYBlockAC,
// enc_modular.cc AddVarDCTDC, dct_util DCFromLowestFrequencies [UPSTREAM]); same arithmetic and operation order as
// oracle/jxo_coef.cc.  The DCT8-only frame (BASELINE config 2) keeps its specialised kernel (k_dct8_v4.cu).
//
// The strategy map is first binned (k_coeff_lists: one list of first blocks per strategy, warp-aggregated appends), so
// that every warp of the transform kernels runs ONE strategy's code:
//   k_coeff8_lanes     DCT, DCT4X4, DCT4X8, DCT8X4: eight lanes per block, four blocks per warp (the geometry of
//                      k_acs_evalsq<8>); a block's three channels stay in 24 registers per lane, a channel leaves as
//                      128 contiguous bytes in scan order
//   k_coeff8_special   IDENTITY, DCT2X2 (rare): one thread per block, the block in registers (fwd8x8)
//   k_coeffsq_all<N>   16 / 32 / 64-sized strategies, one lane group of N lanes per transform: a channel's tile is the
//                      landing zone of its pixels (coalesced cp.async), the transposition buffer and then the lane-ordered
//                      home of the coefficients (lane hf owns every vertical frequency of horizontal frequency hf); the
//                      heuristics' block sums are per-lane sequential sums + xor-butterflies over the lanes; quant
//                      adjustment and quantisation are rolled loops over 16-byte chunks of the lane's row; only non-zero
//                      coefficients are scattered into the (pre-zeroed) scan-order arena
// A lone frame runs the five launches on three streams (StreamFork); in batch mode they share the image's stream.
#include "transforms.cuh"
#include "kernels.h"

namespace jxlb {

// list ids: six 8x8 strategies, then (tall, wide, square) of the three larger levels
enum { kListDCT = 0, kListID, kList2X2, kList4X4, kList4X8, kList8X4, kList16Tall, kList16Wide, kList16Sq, kList32Tall, kList32Wide,
       kList32Sq, kList64Tall, kList64Wide, kList64Sq, kList32Tall4, kList32Wide4, kNumLists };

constexpr int kListHeader = 32;   // counters, then kNumLists lists of nblk entries

__device__ __forceinline__ int list_of_strategy(int s) {
  switch (s) {
    case kStratDCT: return kListDCT;       case kStratIDENTITY: return kListID;    case kStratDCT2X2: return kList2X2;
    case kStratDCT4X4: return kList4X4;    case kStratDCT4X8: return kList4X8;     case kStratDCT8X4: return kList8X4;
    case kStratDCT16X8: return kList16Tall; case kStratDCT8X16: return kList16Wide; case kStratDCT16X16: return kList16Sq;
    case kStratDCT32X16: return kList32Tall; case kStratDCT16X32: return kList32Wide; case kStratDCT32X32: return kList32Sq;
    case kStratDCT64X32: return kList64Tall; case kStratDCT32X64: return kList64Wide; case kStratDCT64X64: return kList64Sq;
    case kStratDCT32X8: return kList32Tall4; case kStratDCT8X32: return kList32Wide4;
    default: return -1;
  }
}

size_t coeff_list_words(const FrameDim& fd) { return kListHeader + (size_t)kNumLists * fd.bxs * fd.bys; }

__global__ void __launch_bounds__(256) k_coeff_lists(const uint8_t* __restrict__ acs, int nblk, uint32_t* __restrict__ lists) {
  const int i = blockIdx.x * 256 + threadIdx.x, lane = threadIdx.x & 31;
  int li = -1;
  if (i < nblk) { const uint8_t a = acs[i]; if (a & 0x80) li = list_of_strategy(a & 0x7f); }
  // one atomic per (warp, list)
  unsigned todo = __ballot_sync(0xffffffffu, li >= 0);
  while (todo) {
    const int leader = __ffs(todo) - 1;
    const int l0 = __shfl_sync(0xffffffffu, li, leader);
    const unsigned same = __ballot_sync(0xffffffffu, li == l0);
    unsigned base = 0;
    if (lane == leader) base = atomicAdd(&lists[l0], (unsigned)__popc(same));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (li == l0) lists[kListHeader + (size_t)l0 * nblk + base + __popc(same & ((1u << lane) - 1))] = (uint32_t)i;
    todo &= ~same;
  }
}

struct CoeffArgs {
  const float* X; const float* Y; const float* B;
  FrameDim fd;
  const QuantDev* qd;
  const float* w; const float* dq;          // tables of the strategy in lane order ([hf][vf]; see AcsTables)
  const uint16_t* inv;                      // coefficient position (same lane order) -> scan index
  const int8_t* cmap;
  float x_qm_mul, b_qm_mul;
  int adjust;
  int32_t* raw_qf; int16_t* coeffs; int16_t* dc_quant; uint8_t* nzeros; uint16_t* nzcount; uint16_t* lastk;
  const uint32_t* list; const uint32_t* count;
};

__device__ __forceinline__ float quant_bias(int c, int q) {
  const float b0 = 1.0f - 0.05465007330715401f, b1 = 1.0f - 0.07005449891748593f, b2 = 1.0f - 0.049935103337343655f;
  const float bc = c == 0 ? b0 : (c == 1 ? b1 : b2);
  if (q == 0) return 0.0f;
  if (q == 1) return bc;
  if (q == -1) return -bc;
  const float fq = (float)q;
  return fq - 0.145f / fq;
}
__device__ __forceinline__ float clamp1(float v, float lo, float hi) { return v < lo ? lo : (v > hi ? hi : v); }
__device__ __forceinline__ int16_t sat16(int v) { return (int16_t)(v > 32767 ? 32767 : (v < -32768 ? -32768 : v)); }

// AddVarDCTDC quantisation of one block's DC values (Y first; X / B with the default 0 / 1.0 base correlation)
__device__ __forceinline__ void store_dc(const CoeffArgs& A, size_t bj, float dcX, float dcY, float dcB) {
  const size_t nblk = (size_t)A.fd.bxs * A.fd.bys;
  const float scale = A.qd->scale, inv_gs = A.qd->inv_global_scale;
  const int quant_dc = A.qd->quant_dc;
  const float gsq = scale * (float)quant_dc;
  const float inv_quant_dc = inv_gs / (float)quant_dc;
  const float y_factor = inv_quant_dc * (1.0f / 512.0f);
  const float qy = roundf(dcY * (512.0f * gsq));
  const float qx = roundf((dcX - qy * (y_factor * 0.0f)) * (4096.0f * gsq));
  const float qb = roundf((dcB - qy * (y_factor * 1.0f)) * (256.0f * gsq));
  A.dc_quant[0 * nblk + bj] = sat16((int)qx);
  A.dc_quant[1 * nblk + bj] = sat16((int)qy);
  A.dc_quant[2 * nblk + bj] = sat16((int)qb);
}

// closing rules of AdjustQuantBlockAC once the block sums are known (oracle/jxo_coef.cc)
template <int S>
__device__ __forceinline__ int adjust_close(int c, int quant, int ncov, float sum_hf_rc, float sum_err, float sum_vals, const float hfNZ[4],
                                            const float hfME[4], float thr[4]) {
  if (c == 1 && sum_vals * 8 < (float)ncov) {
    const double kLimit = 0.46, kMul = 0.9999;
    const int orig = quant;
    int nq = quant;
#pragma unroll
    for (int i = 1; i < 4; ++i) if (nq == orig && hfNZ[i] == 0.0f && (double)hfME[i] > kLimit) nq = orig + 1;
    quant = nq;
    if (hfNZ[3] == 0.0f && (double)hfME[3] > kLimit) {
      thr[3] = (float)(kMul * (double)hfME[3] * (double)nq / (double)orig);
    } else if ((hfNZ[1] == 0.0f && (double)hfME[1] > kLimit) || (hfNZ[2] == 0.0f && (double)hfME[2] > kLimit)) {
      const float m = hfME[1] > hfME[2] ? hfME[1] : hfME[2];
      thr[1] = (float)(kMul * (double)m * (double)nq / (double)orig);
      thr[2] = thr[1];
    } else if (hfNZ[0] == 0.0f && (double)hfME[0] > kLimit) {
      thr[0] = (float)(kMul * (double)hfME[0] * (double)nq / (double)orig);
    }
  }
  {
    const float all = hfNZ[0] + hfNZ[1] + hfNZ[2] + hfNZ[3] + 1;
    const float mul = c == 0 ? 70.0f : (c == 1 ? 30.0f : 60.0f);
    if (mul * sum_hf_rc >= all) {
      quant = (int)((float)quant + mul * sum_hf_rc / all);
      if (quant >= 256) quant = 255;
    }
  }
  if constexpr (S == kStratDCT) {
    if (hfNZ[0] + hfNZ[1] + hfNZ[2] + hfNZ[3] < 11) { quant += 1; if (quant >= 256) quant = 255; }
  }
  if constexpr (S >= kStratDCT16X16 && S != kStratDCT4X8 && S != kStratDCT8X4) {
    const double kMul1[4][3] = {{0.22080615753848404, 0.45797479824262011, 0.29859235095977965},
                                {0.70109486510286834, 0.16185281305512639, 0.14387691730035473},
                                {0.114985964456218638, 0.44656840441027695, 0.10587658215149048},
                                {0.46849665264409396, 0.41239077937781954, 0.088667407767185444}};
    const double kMul2[4][3] = {{0.27450281941822197, 1.1255766549984996, 0.98950459134128388},
                                {0.4652168675598285, 0.40945807983455818, 0.36581899811751367},
                                {0.28034972424715715, 0.9182653201929738, 1.5581531543057416},
                                {0.26873118114033728, 0.68863712390392484, 1.2082185408666786}};
    const double kQuantNormalizer = 2.2942708343284721;
    const double se = (double)sum_err * kQuantNormalizer;
    const double sv = (double)sum_vals * kQuantNormalizer;
    constexpr int ix = (S == kStratDCT32X16 || S == kStratDCT16X32) ? 1 : (S == kStratDCT16X16 ? 0 : (S == kStratDCT32X32 ? 2 : 3));
    const double lim = kMul1[ix][c] * (double)(ncov * 64) + kMul2[ix][c] * sv;
    int step = (int)(se / lim);
    if (step >= 2) step = 2;
    if (step < 0) step = 0;
    if (se > lim) { quant += step; if (quant >= 256) quant = 255; }
  }
  return quant;
}

// ====================================================================================================== 8x8 strategies
// natural (zig-zag) coefficient order of an 8x8 block (oracle NaturalCoeffOrder, cx = cy = 1), at compile time
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
constexpr Zigzag8 make_inv_zigzag8() {
  const Zigzag8 z = make_zigzag8();
  Zigzag8 inv{};
  for (int k = 0; k < 64; ++k) inv.pos[z.pos[k]] = k;
  return inv;
}
__device__ constexpr Zigzag8 kInvZigzag8 = make_inv_zigzag8();   // storage position -> scan index

// oracle QuantizeBlockAC for an 8x8 block in registers; returns the values in place (as ints)
__device__ __forceinline__ void quantize8(const float* cf, const float* __restrict__ qm, float qac_mul, const float thr[4], int* out) {
#pragma unroll
  for (int k = 0; k < 64; ++k) {
    const int y = k >> 3, x = k & 7;
    const float t = thr[(y >= 4 ? 2 : 0) + (x >= 4 ? 1 : 0)];
    const float q = __ldg(qm + k) * qac_mul;
    const float val = q * cf[k];
    int v = (fabsf(val) >= t) ? (int)rintf(val) : 0;
    if (k == 0) v = 0;
    v = v > 32767 ? 32767 : (v < -32767 ? -32767 : v);
    out[k] = v;
  }
}

// scan-order output of one channel of an 8x8 block + its non-zero statistics
__device__ __forceinline__ void emit8(const CoeffArgs& A, const int* q, size_t cblk, int slot, int c, size_t bi, size_t nblk) {
  int nz = 0, last = 0;
  uint32_t words[32];
#pragma unroll
  for (int k = 0; k < 64; k += 2) {
    const int a = q[kZigzag8.pos[k]], b = q[kZigzag8.pos[k + 1]];
    if (a != 0) { ++nz; last = k; }
    if (b != 0) { ++nz; last = k + 1; }
    words[k >> 1] = ((uint32_t)a & 0xFFFFu) | ((uint32_t)b << 16);
  }
  uint4* dst = reinterpret_cast<uint4*>(A.coeffs + (cblk * 3 + slot) * 64);
#pragma unroll
  for (int i = 0; i < 8; ++i) dst[i] = make_uint4(words[4 * i], words[4 * i + 1], words[4 * i + 2], words[4 * i + 3]);
  A.nzcount[(size_t)c * nblk + bi] = (uint16_t)nz;
  A.lastk[(size_t)c * nblk + bi] = (uint16_t)last;
  A.nzeros[(size_t)c * nblk + bi] = (uint8_t)nz;
}

__device__ __forceinline__ void load_block8(const float* __restrict__ plane, int pitch, int bx, int by, float* p) {
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    const float4* src = reinterpret_cast<const float4*>(plane + (size_t)(by * 8 + r) * pitch + bx * 8);
    const float4 a = __ldg(src), d = __ldg(src + 1);
    p[r * 8 + 0] = a.x; p[r * 8 + 1] = a.y; p[r * 8 + 2] = a.z; p[r * 8 + 3] = a.w;
    p[r * 8 + 4] = d.x; p[r * 8 + 5] = d.y; p[r * 8 + 6] = d.z; p[r * 8 + 7] = d.w;
  }
}

// oracle AdjustQuantBlockAC for the plain 8x8 DCT (the other 8x8 strategies are not adjusted): lane = storage row
__device__ __forceinline__ int adjust_quant8(const float* cf, const float* __restrict__ qm, int c, float scale, float qm_mul, int quant, float thr[4]) {
  const float qac = scale * (float)quant;
  float r_hf[8], r_err[8], r_vals[8], nz0[8], nz1[8], nz2[8], nz3[8];
  float hfME[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
  for (int y = 0; y < 8; ++y) {
    float a_hf = 0.0f, a_err = 0.0f, a_vals = 0.0f, a_nzA = 0.0f, a_nzB = 0.0f;
#pragma unroll
    for (int x = 0; x < 8; ++x) {
      if (x == 0 && y == 0) continue;
      const int k = y * 8 + x;
      const int hfix = (y >= 4 ? 2 : 0) + (x >= 4 ? 1 : 0);
      const float val = cf[k] * (__ldg(qm + k) * qac * qm_mul);
      const float v = (fabsf(val) < thr[hfix]) ? 0.0f : rintf(val);
      const float err = fabsf(val - v);
      a_err += err;
      a_vals += fabsf(v);
      if (c == 1 && v == 0.0f) { if (hfME[hfix] < err) hfME[hfix] = err; }
      if (v != 0.0f) {
        if (x >= 4) a_nzB += fabsf(v); else a_nzA += fabsf(v);
        const bool in_corner = y >= 7 && x >= 7;
        const bool on_border = y == 7 || x == 7;
        const bool in_larger_corner = x >= 4 && y >= 4;
        if (in_corner || (on_border && in_larger_corner)) a_hf += fabsf(val);
      }
    }
    r_hf[y] = a_hf; r_err[y] = a_err; r_vals[y] = a_vals;
    nz0[y] = y < 4 ? a_nzA : 0.0f; nz1[y] = y < 4 ? a_nzB : 0.0f; nz2[y] = y < 4 ? 0.0f : a_nzA; nz3[y] = y < 4 ? 0.0f : a_nzB;
  }
  const float hfNZ[4] = {tree8(nz0), tree8(nz1), tree8(nz2), tree8(nz3)};
  return adjust_close<kStratDCT>(c, quant, 1, tree8(r_hf), tree8(r_err), tree8(r_vals), hfNZ, hfME, thr);
}

template <int S>
__device__ __noinline__ void coeff8_body(const CoeffArgs& A, unsigned i) {
  const FrameDim& fd = A.fd;
  const size_t nblk = (size_t)fd.bxs * fd.bys;
  const size_t bi = A.list[i];
  const int bx = (int)(bi % fd.bxs), by = (int)(bi / fd.bxs);
  const float scale = A.qd->scale, inv_gs = A.qd->inv_global_scale;
  const int orig = A.raw_qf[bi];
  const float* planes[3] = {A.X, A.Y, A.B};
  float ycoef[64];
  float dc[3];
  int quant = orig;
  float thr_y[4] = {0.58f, 0.64f, 0.64f, 0.64f};
  // ---- pass A: transforms, DC, quant adjust (channel order Y, X, B; the result is the maximum over the channels)
  {
    int maxq = 0;
#pragma unroll 1
    for (int it = 0; it < 3; ++it) {
      const int c = it == 0 ? 1 : (it == 1 ? 0 : 2);
      float p[64], cf[64];
      load_block8(planes[c], fd.pitch, bx, by, p);
      fwd8x8<S>(p, cf);
      dc[c] = cf[0];
      if (it == 0) {
#pragma unroll
        for (int k = 0; k < 64; ++k) ycoef[k] = cf[k];
      }
      if (S == kStratDCT && A.adjust) {
        float thr[4] = {0.58f, 0.64f, 0.64f, 0.64f};
        const float mulc = c == 0 ? A.x_qm_mul : (c == 1 ? 1.0f : A.b_qm_mul);
        maxq = max(maxq, adjust_quant8(cf, A.w + c * 64, c, scale, mulc, orig, thr));
        if (c == 1) { thr_y[0] = thr[0]; thr_y[1] = thr[1]; thr_y[2] = thr[2]; thr_y[3] = thr[3]; }
      }
    }
    if (A.adjust) { if (S == kStratDCT) quant = maxq; }
    else { thr_y[0] = 0.56f; thr_y[1] = thr_y[2] = thr_y[3] = 0.62f; }
  }
  store_dc(A, bi, dc[0], dc[1], dc[2]);
  A.raw_qf[bi] = quant;
  // ---- pass B: quantise Y, roundtrip it, remove the chroma-from-luma share from X and B, quantise them
  const float qac = scale * (float)quant;
  const float inv_qac = inv_gs / (float)quant;
  const int g = (by >> 5) * fd.gxs + (bx >> 5);
  const size_t cblk = (size_t)g * kGroupBlocks + (size_t)(by & 31) * 32 + (bx & 31);
  const int tx = bx >> 3, ty = by >> 3;
  const float x_factor = 0.0f + (float)A.cmap[(size_t)ty * fd.txs + tx] / 84.0f;
  const float b_factor = 1.0f + (float)A.cmap[(size_t)fd.txs * fd.tys + (size_t)ty * fd.txs + tx] / 84.0f;
  int qy[64];
  quantize8(ycoef, A.w + 64, qac * 1.0f, thr_y, qy);
  emit8(A, qy, cblk, 0, 1, bi, nblk);
  // dequantised Y replaces the coefficients (only its products with the two factors are needed below)
#pragma unroll
  for (int k = 0; k < 64; ++k) ycoef[k] = (quant_bias(1, qy[k]) * __ldg(A.dq + 64 + k)) * inv_qac;
#pragma unroll 1
  for (int it = 0; it < 2; ++it) {
    const int c = it == 0 ? 0 : 2;
    const float factor = c == 0 ? x_factor : b_factor;
    float p[64], cf[64];
    load_block8(planes[c], fd.pitch, bx, by, p);
    fwd8x8<S>(p, cf);
#pragma unroll
    for (int k = 0; k < 64; ++k) cf[k] = __fmaf_rn(-factor, ycoef[k], cf[k]);
    const float thr[4] = {0.58f, 0.64f, 0.64f, 0.64f};
    int q[64];
    quantize8(cf, A.w + c * 64, qac * (c == 0 ? A.x_qm_mul : A.b_qm_mul), thr, q);
    emit8(A, q, cblk, c == 0 ? 1 : 2, c, bi, nblk);
  }
}

// ---------------------------------------------------------------------------------------------- 8x8, eight lanes per block
// DCT, DCT4X4, DCT4X8 and DCT8X4 (the bulk of a searched frame's 8x8 blocks) run with eight lanes per block, four blocks
// per warp — the geometry of k_acs_evalsq<8>: lane (q * 4 + hf) of a block owns the eight coefficients storage_index()
// gives it (oracle/jxo_acs.cc), three channels of them stay in registers, and the code of a strategy is a few hundred
// instructions.  (The one-thread-per-block version ran 5 800 instructions once per thread and spent 58 % of its stall
// samples on instruction fetches: ncu, profiles/r02k.)  A group's tile is 8 rows x 12 floats, 104 floats from the next
// group's: row accesses (16 bytes) and column accesses of the warp's four groups are conflict-free.
constexpr int kC8Pitch = 12, kC8Tile = 104, kC8WarpFloats = 3 * 4 * kC8Tile + 4 * 32;   // tiles [c][group], then 64 int16 per group

template <int S> __device__ __forceinline__ int pos8_of(int l, int j) {   // storage position of a lane's j-th value
  if constexpr (S == kStratDCT) return l * 8 + j;
  else if constexpr (S == kStratDCT4X4) return ((j >> 2) + (l & 3) * 2) * 8 + (l >> 2) + (j & 3) * 2;
  else if constexpr (S == kStratDCT4X8) return ((l >> 2) + (l & 3) * 2) * 8 + j;
  else return ((j >> 2) + (j & 3) * 2) * 8 + l;                              // DCT8X4
}

// rows, then columns, then the DC Hadamard of the split strategies (same operations as k_acs_evalsq<8>'s forward half)
template <int S>
__device__ __forceinline__ void fwd8_lanes(float* t, int l, float (&v)[8]) {
  constexpr bool row_full = S == kStratDCT || S == kStratDCT8X4, col_full = S == kStratDCT || S == kStratDCT4X8;
  constexpr int P = kC8Pitch;
  {
    const float4 a = *reinterpret_cast<const float4*>(t + l * P), b = *reinterpret_cast<const float4*>(t + l * P + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  }
  if constexpr (row_full) dct1d<8>(v); else { dct1d<4>(v); dct1d<4>(v + 4); }
  *reinterpret_cast<float4*>(t + l * P) = make_float4(v[0], v[1], v[2], v[3]);
  *reinterpret_cast<float4*>(t + l * P + 4) = make_float4(v[4], v[5], v[6], v[7]);
  __syncwarp();
#pragma unroll
  for (int y = 0; y < 8; ++y) v[y] = t[y * P + l];
  if constexpr (col_full) dct1d<8>(v); else { dct1d<4>(v); dct1d<4>(v + 4); }
  if constexpr (S != kStratDCT) {
    const float pa = __shfl_xor_sync(0xffffffffu, v[0], 4), pb = __shfl_xor_sync(0xffffffffu, v[4], 4);
    if constexpr (S == kStratDCT4X8) {
      if (l == 0) v[0] = (v[0] + pa) * 0.5f; else if (l == 4) v[0] = (pa - v[0]) * 0.5f;
    } else if constexpr (S == kStratDCT8X4) {
      if (l == 0) { const float b0 = v[0], b1 = v[4]; v[0] = (b0 + b1) * 0.5f; v[4] = (b0 - b1) * 0.5f; }
    } else {
      if (l == 0) { const float b00 = v[0], b01 = pa, b10 = v[4], b11 = pb; v[0] = (b00 + b01 + b10 + b11) * 0.25f; v[4] = (b00 - b01 + b10 - b11) * 0.25f; }
      else if (l == 4) { const float b00 = pa, b01 = v[0], b10 = pb, b11 = v[4]; v[0] = (b00 + b01 - b10 - b11) * 0.25f; v[4] = (b00 - b01 - b10 + b11) * 0.25f; }
    }
  }
}

// oracle AdjustQuantBlockAC for the plain 8x8 DCT, lane = storage row (adjust_quant8 with the rows spread over the lanes)
__device__ __forceinline__ int adjust_quant8_lanes(const float (&cf)[8], const float* __restrict__ wrow, int c, float scale, float qm_mul,
                                                   int quant, float thr[4], int l) {
  const float qac = scale * (float)quant;
  const bool first = l < 4;
  const float thrA = first ? thr[0] : thr[2], thrB = first ? thr[1] : thr[3];
  float a_hf = 0.0f, a_err = 0.0f, a_vals = 0.0f, nzA = 0.0f, nzB = 0.0f, meA = 0.0f, meB = 0.0f;
  const float4 w0 = __ldg(reinterpret_cast<const float4*>(wrow)), w1 = __ldg(reinterpret_cast<const float4*>(wrow + 4));
  const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
  for (int x = 0; x < 8; ++x) {
    if (x == 0 && l == 0) continue;
    const bool second = x >= 4;
    const float val = cf[x] * (wv[x] * qac * qm_mul);
    const float v = (fabsf(val) < (second ? thrB : thrA)) ? 0.0f : rintf(val);
    const float err = fabsf(val - v);
    a_err += err;
    a_vals += fabsf(v);
    if (c == 1 && v == 0.0f) { if (second) { if (meB < err) meB = err; } else { if (meA < err) meA = err; } }
    if (v != 0.0f) {
      if (second) nzB += fabsf(v); else nzA += fabsf(v);
      const bool in_corner = l >= 7 && x >= 7;
      const bool on_border = l == 7 || x == 7;
      const bool in_larger_corner = x >= 4 && l >= 4;
      if (in_corner || (on_border && in_larger_corner)) a_hf += fabsf(val);
    }
  }
  const float r_hf = group_sum<8>(a_hf), r_err = group_sum<8>(a_err), r_vals = group_sum<8>(a_vals);
  const float hfNZ[4] = {group_sum<8>(first ? nzA : 0.0f), group_sum<8>(first ? nzB : 0.0f), group_sum<8>(first ? 0.0f : nzA),
                         group_sum<8>(first ? 0.0f : nzB)};
  float hfME[4] = {0.0f, 0.0f, 0.0f, 0.0f};
  if (c == 1) {
    hfME[0] = group_fmax<8>(first ? meA : 0.0f); hfME[1] = group_fmax<8>(first ? meB : 0.0f);
    hfME[2] = group_fmax<8>(first ? 0.0f : meA); hfME[3] = group_fmax<8>(first ? 0.0f : meB);
  }
  return adjust_close<kStratDCT>(c, quant, 1, r_hf, r_err, r_vals, hfNZ, hfME, thr);
}

// one round of a warp: four list entries (blocks), eight lanes each
template <int S>
__device__ __noinline__ void coeff8_lanes_body(const CoeffArgs& A, float* smem_w, unsigned item0) {
  constexpr int P = kC8Pitch;
  const FrameDim& fd = A.fd;
  const int lane = threadIdx.x & 31, l = lane & 7, grp = lane >> 3;
  const unsigned n = *A.count;
  const unsigned item = item0 + grp;
  const bool active = item < n;
  const size_t nblk = (size_t)fd.bxs * fd.bys;
  const size_t bi = active ? A.list[item] : 0;
  const int bx = (int)(bi % fd.bxs), by = (int)(bi / fd.bxs);
  const float scale = A.qd->scale, inv_gs = A.qd->inv_global_scale;
  const int orig = active ? A.raw_qf[bi] : 1;
  int16_t* stage = reinterpret_cast<int16_t*>(smem_w + 3 * 4 * kC8Tile) + grp * 64;
  __syncwarp();                                // (the previous round's reads of the tiles)
  {
    // the group's eight lanes fetch their block: two 16-byte chunks per row, four rows per instruction
    const int fr = l >> 1, fc = l & 1;
    const size_t off = active ? (size_t)(by * 8 + fr) * fd.pitch + (size_t)bx * 8 + 4 * fc : 0;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const int c = k == 0 ? 1 : (k == 1 ? 0 : 2);
      const float* plane = c == 0 ? A.X : (c == 1 ? A.Y : A.B);
      float* dst = smem_w + (c * 4 + grp) * kC8Tile + fr * P + 4 * fc;
#pragma unroll
      for (int i = 0; i < 2; ++i) cp_async16(dst + i * 4 * P, active ? plane + off + (size_t)(4 * i) * fd.pitch : plane, active);
    }
    cp_async_commit();
  }
  cp_async_wait<0>();
  __syncwarp();
  float cf[3][8];
  float thr_y[4] = {0.58f, 0.64f, 0.64f, 0.64f};
  int maxq = 0;
#pragma unroll
  for (int it = 0; it < 3; ++it) {
    const int c = it == 0 ? 1 : (it == 1 ? 0 : 2);
    fwd8_lanes<S>(smem_w + (c * 4 + grp) * kC8Tile, l, cf[c]);
    if (S == kStratDCT && A.adjust) {
      float thr[4] = {0.58f, 0.64f, 0.64f, 0.64f};
      const float mulc = c == 0 ? A.x_qm_mul : (c == 1 ? 1.0f : A.b_qm_mul);
      maxq = max(maxq, adjust_quant8_lanes(cf[c], A.w + c * 64 + l * 8, c, scale, mulc, orig, thr, l));
      if (c == 1) { thr_y[0] = thr[0]; thr_y[1] = thr[1]; thr_y[2] = thr[2]; thr_y[3] = thr[3]; }
    }
  }
  int quant = orig;
  if (A.adjust) { if (S == kStratDCT) quant = maxq; }
  else { thr_y[0] = 0.56f; thr_y[1] = thr_y[2] = thr_y[3] = 0.62f; }
  {
    // the DC values sit in lane 0's first coefficient
    const float dX = __shfl_sync(0xffffffffu, cf[0][0], grp * 8), dY = __shfl_sync(0xffffffffu, cf[1][0], grp * 8),
                dB = __shfl_sync(0xffffffffu, cf[2][0], grp * 8);
    if (active && l == 0) { store_dc(A, bi, dX, dY, dB); A.raw_qf[bi] = quant; }
  }
  // ---- quantise Y, roundtrip it, remove the chroma-from-luma share from X and B, quantise them
  const float qac = scale * (float)quant;
  const float inv_qac = inv_gs / (float)quant;
  const size_t cblk = (size_t)((by >> 5) * fd.gxs + (bx >> 5)) * kGroupBlocks + (size_t)(by & 31) * 32 + (bx & 31);
  const int tx = bx >> 3, ty = by >> 3;
  const float x_factor = 0.0f + (float)A.cmap[(size_t)ty * fd.txs + tx] / 84.0f;
  const float b_factor = 1.0f + (float)A.cmap[(size_t)fd.txs * fd.tys + (size_t)ty * fd.txs + tx] / 84.0f;
  float ydq[8];
#pragma unroll
  for (int it = 0; it < 3; ++it) {
    const int c = it == 0 ? 1 : (it == 1 ? 0 : 2);
    const float4 w0 = __ldg(reinterpret_cast<const float4*>(A.w + c * 64 + l * 8)), w1 = __ldg(reinterpret_cast<const float4*>(A.w + c * 64 + l * 8 + 4));
    const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
    const float qac_mul = qac * (c == 0 ? A.x_qm_mul : (c == 1 ? 1.0f : A.b_qm_mul));
    const float factor = c == 0 ? x_factor : b_factor;
    float thr[4] = {0.58f, 0.64f, 0.64f, 0.64f};
    if (c == 1) { thr[0] = thr_y[0]; thr[1] = thr_y[1]; thr[2] = thr_y[2]; thr[3] = thr_y[3]; }
    int nz = 0, last = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int pos = pos8_of<S>(l, j);
      const int y = pos >> 3, x = pos & 7;
      const bool yh = y >= 4, xh = x >= 4;
      const float t = yh ? (xh ? thr[3] : thr[2]) : (xh ? thr[1] : thr[0]);
      const float in = c == 1 ? cf[1][j] : __fmaf_rn(-factor, ydq[j], cf[c][j]);
      const float val = (wv[j] * qac_mul) * in;
      int q = (fabsf(val) >= t) ? (int)rintf(val) : 0;
      if (pos == 0) q = 0;
      q = q > 32767 ? 32767 : (q < -32767 ? -32767 : q);
      const int k = kInvZigzag8.pos[pos];
      stage[k] = (int16_t)q;
      if (q != 0) { ++nz; last = max(last, k); }
      if (c == 1) cf[1][j] = (float)q;       // (kept for the roundtrip below)
    }
    if (c == 1) {
      const float4 d0 = __ldg(reinterpret_cast<const float4*>(A.dq + 64 + l * 8)), d1 = __ldg(reinterpret_cast<const float4*>(A.dq + 64 + l * 8 + 4));
      const float dv[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
#pragma unroll
      for (int j = 0; j < 8; ++j) ydq[j] = (quant_bias(1, (int)cf[1][j]) * dv[j]) * inv_qac;
    }
    nz = group_isum<8>(nz);
    last = group_imax<8>(last);
    __syncwarp();
    if (active) {
      // scan-order output of the block's channel: 128 contiguous bytes, 16 per lane
      reinterpret_cast<uint4*>(A.coeffs + (cblk * 3 + it) * 64)[l] = reinterpret_cast<const uint4*>(stage)[l];
      if (l == 0) {
        A.nzcount[(size_t)c * nblk + bi] = (uint16_t)nz;
        A.lastk[(size_t)c * nblk + bi] = (uint16_t)last;
        A.nzeros[(size_t)c * nblk + bi] = (uint8_t)nz;
      }
    }
    __syncwarp();                              // `stage` is rewritten by the next channel
  }
}

// ====================================================================================================== 16 / 32 / 64
// Lane geometry of a transform inside the N x N square (SquareXform modes): square = N lanes x N values; tall = the
// first N/2 lanes x N values (left half; the right half of the square is not loaded); wide = N lanes x the first N/2
// values (top half).  (x, y) below are storage coordinates (long side horizontal): square / tall: y = lane, x = j;
// wide: x = lane, y = j.
template <int N, int MODE> struct CoeffGeom {
  static constexpr bool kWide = MODE == kModeWide2 || MODE == kModeWide4;   // lane = storage column x, value index = storage row y
  static constexpr int SHORT = MODE == kModeSq ? N : ((MODE == kModeTall2 || MODE == kModeWide2) ? N / 2 : N / 4);   // short side
  static constexpr int H = N / 2;
  static constexpr int W = N;                                   // storage width (long side)
  static constexpr int HS = SHORT;                              // storage height
  static constexpr int LANES = kWide ? N : SHORT;               // lanes that own coefficients
  static constexpr int VALS = kWide ? SHORT : N;                // values per lane
  static constexpr int R = kWide ? SHORT : N, C = kWide ? N : SHORT;   // pixel rows / columns
  static constexpr int cxb = C / 8, cyb = R / 8, ncov = cxb * cyb, xs = W / 8, ys = HS / 8;
  static constexpr int S = N == 16 ? (MODE == kModeSq ? kStratDCT16X16 : MODE == kModeTall2 ? kStratDCT16X8 : kStratDCT8X16)
                         : N == 32 ? (MODE == kModeSq ? kStratDCT32X32 : MODE == kModeTall2 ? kStratDCT32X16 : MODE == kModeWide2 ? kStratDCT16X32
                                      : MODE == kModeTall4 ? kStratDCT32X8 : kStratDCT8X32)
                                   : (MODE == kModeSq ? kStratDCT64X64 : MODE == kModeTall2 ? kStratDCT64X32 : kStratDCT32X64);
};

// sums / maxima over the LANES lanes that own coefficients (N = 64 square / wide: two warps through `xch`)
template <int N, int LANES, int K>
__device__ __forceinline__ void lane_sums(float (&v)[K], float* xch, int l, int bar_id) {
  if constexpr (LANES == 64) {
#pragma unroll
    for (int k = 0; k < K; ++k) xch[k * 64 + l] = v[k];
    SquareXform<N>::sync(bar_id);
#pragma unroll
    for (int k = 0; k < K; ++k) v[k] = v[k] + xch[k * 64 + (l ^ 32)];
    SquareXform<N>::sync(bar_id);
#pragma unroll
    for (int k = 0; k < K; ++k) v[k] = group_sum<32>(v[k]);
  } else {
#pragma unroll
    for (int k = 0; k < K; ++k) v[k] = group_sum<LANES>(v[k]);
  }
}
template <int N, int LANES, int K>
__device__ __forceinline__ void lane_maxs(float (&v)[K], float* xch, int l, int bar_id) {
  if constexpr (LANES == 64) {
#pragma unroll
    for (int k = 0; k < K; ++k) xch[k * 64 + l] = v[k];
    SquareXform<N>::sync(bar_id);
#pragma unroll
    for (int k = 0; k < K; ++k) v[k] = fmaxf(v[k], xch[k * 64 + (l ^ 32)]);
    SquareXform<N>::sync(bar_id);
#pragma unroll
    for (int k = 0; k < K; ++k) v[k] = group_fmax<32>(v[k]);
  } else {
#pragma unroll
    for (int k = 0; k < K; ++k) v[k] = group_fmax<LANES>(v[k]);
  }
}

// Shared memory of one lane group: three N x (N + 4) tiles (one per channel), the lowest-frequency scratch of the DC
// extraction, and (N = 64) the exchange rows of the two-warp sums.  A channel's tile is first the landing zone of its
// pixels (coalesced 16-byte cp.async: one instruction = four whole rows), then the transposition buffer of the transform,
// then the lane-ordered home of the coefficients: lane hf keeps its N values in row hf.  The pitch is 4 floats past a
// multiple of 32: 16-byte accesses to a lane's own row and 4-byte accesses down a column are both conflict-free (N = 16:
// the two groups of a warp are 16 banks apart).
// Everything after the transform runs as rolled loops over 16-byte chunks of the lane's row (tables are stored in the same
// chunks, [c][chunk][lane][4], so a warp's table load is one contiguous run): the code of a transform is a few thousand
// instructions that stay in the instruction cache, where the fully unrolled version streamed 35 000 of them once per item
// and spent 40 % of its time waiting for instruction fetches (ncu, profiles/r02k).
template <int N> struct CoeffSqGeom {
  static constexpr int kGroupsPerWarp = N == 16 ? 2 : 1;
  static constexpr int kThreads = N == 64 ? 64 : 128;
  static constexpr int kGroups = N == 64 ? 1 : 4 * kGroupsPerWarp;          // transforms per CTA pass
  static constexpr int kPitch = N + 4;
  static constexpr int kSq = N * kPitch;
  static constexpr int kLlf = N == 16 ? 16 : 3 * (N / 8) * (N / 8);          // [c][cyb][cxb]
  static constexpr int kXch = N == 64 ? 7 * 64 : 0;
  static constexpr int kGroupFloats = 3 * kSq + kLlf + kXch;                 // (N = 16: 976 = 16 mod 32)
  static constexpr int kSmemFloats = kGroups * kGroupFloats;
};

// oracle AdjustQuantBlockAC for a lane group; row = the lane's coefficients, wl = the lane's slot of the chunked table
template <int N, int MODE>
__device__ __forceinline__ int adjust_quant_sq(const float* row, const float* __restrict__ wl, int c, float scale, float qm_mul, int quant,
                                               float thr[4], int l, bool owner, float* xch, int bar_id) {
  using G = CoeffGeom<N, MODE>;
  const float qac = scale * (float)quant;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    thr[i] -= clamp1(0.003f * (float)(G::xs * G::ys), 0.f, (c > 0 ? 0.08f : 0.12f));
    if (thr[i] < 0.54f) thr[i] = 0.54f;
  }
  float r_hf = 0.0f, r_err = 0.0f, r_vals = 0.0f, nzA = 0.0f, nzB = 0.0f, meA = 0.0f, meB = 0.0f;
  // quadrant index hfix = (y half) * 2 + (x half); `first` = the lane sits in the first half of the lane-indexed axis;
  // A / B = first / second half of the axis the lane's own values run along
  const bool first = G::kWide ? (l < G::W / 2) : (l < G::HS / 2);
  const float thrA = G::kWide ? (first ? thr[0] : thr[1]) : (first ? thr[0] : thr[2]);
  const float thrB = G::kWide ? (first ? thr[2] : thr[3]) : (first ? thr[1] : thr[3]);
  if (owner) {
    const float wmul = qac;   // (w * qac) * qm_mul, in this order
#pragma unroll 1
    for (int j4 = 0; j4 < G::VALS; j4 += 4) {
      const float4 w4 = __ldg(reinterpret_cast<const float4*>(wl + (size_t)(j4 >> 2) * G::LANES * 4));
      const float4 u4 = *reinterpret_cast<const float4*>(row + j4);
      const float wv[4] = {w4.x, w4.y, w4.z, w4.w}, uv[4] = {u4.x, u4.y, u4.z, u4.w};
      const bool second = j4 >= G::VALS / 2;
      const float t = second ? thrB : thrA;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int j = j4 + e;
        const int x = G::kWide ? l : j, y = G::kWide ? j : l;
        if (x < G::xs && y < G::ys) continue;
        const float val = uv[e] * (wv[e] * wmul * qm_mul);
        const float v = (fabsf(val) < t) ? 0.0f : rintf(val);
        const float err = fabsf(val - v);
        r_err += err;
        r_vals += fabsf(v);
        if (c == 1 && v == 0.0f) { if (second) { if (meB < err) meB = err; } else { if (meA < err) meA = err; } }
        if (v != 0.0f) {
          if (second) nzB += fabsf(v); else nzA += fabsf(v);
          const bool in_corner = y >= 7 * G::ys && x >= 7 * G::xs;
          const bool on_border = y == G::HS - 1 || x == G::W - 1;
          const bool in_larger_corner = x >= 4 * G::xs && y >= 4 * G::ys;
          if (in_corner || (on_border && in_larger_corner)) r_hf += fabsf(val);
        }
      }
    }
  }
  float s[7];
  s[0] = r_hf; s[1] = r_err; s[2] = r_vals;
  if constexpr (G::kWide) {   // lane = x: first -> quadrants 0 (A: top) and 2 (B: bottom); else 1 and 3
    s[3] = first ? nzA : 0.0f; s[4] = first ? 0.0f : nzA; s[5] = first ? nzB : 0.0f; s[6] = first ? 0.0f : nzB;
  } else {                              // lane = y: first -> quadrants 0 (A: left) and 1 (B: right); else 2 and 3
    s[3] = first ? nzA : 0.0f; s[4] = first ? nzB : 0.0f; s[5] = first ? 0.0f : nzA; s[6] = first ? 0.0f : nzB;
  }
  lane_sums<N, G::LANES, 7>(s, xch, l, bar_id);
  const float hfNZ[4] = {s[3], s[4], s[5], s[6]};
  float hfME[4] = {0.0f, 0.0f, 0.0f, 0.0f};
  if (c == 1) {
    float m[4];
    if constexpr (G::kWide) { m[0] = first ? meA : 0.0f; m[1] = first ? 0.0f : meA; m[2] = first ? meB : 0.0f; m[3] = first ? 0.0f : meB; }
    else { m[0] = first ? meA : 0.0f; m[1] = first ? meB : 0.0f; m[2] = first ? 0.0f : meA; m[3] = first ? 0.0f : meB; }
    lane_maxs<N, G::LANES, 4>(m, xch, l, bar_id);
    hfME[0] = m[0]; hfME[1] = m[1]; hfME[2] = m[2]; hfME[3] = m[3];
  }
  return adjust_close<G::S>(c, quant, G::ncov, s[0], s[1], s[2], hfNZ, hfME, thr);
}

template <int N, int MODE>
__device__ __noinline__ void coeffsq_body(const CoeffArgs& A, float* smem_f, unsigned item0) {
  using G = CoeffGeom<N, MODE>;
  using CG = CoeffSqGeom<N>;
  using SX = SquareXform<N>;
  constexpr int P = CG::kPitch;
  const FrameDim& fd = A.fd;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int l = N == 64 ? tid : (lane & (N - 1));
  const int grp = N == 64 ? 0 : warp * CG::kGroupsPerWarp + (N == 16 ? (lane >> 4) : 0);
  float* base = smem_f + grp * CG::kGroupFloats;
  float* llf = base + 3 * CG::kSq;          // [c][cyb][cxb]: lowest frequencies, then the DC values
  float* xch = llf + CG::kLlf;              // [7][64] (N = 64)
  const unsigned n = *A.count;
  const size_t nblk = (size_t)fd.bxs * fd.bys;
  const float scale = A.qd->scale, inv_gs = A.qd->inv_global_scale;
  const bool owner = l < G::LANES;
  const int bar_id = 1;
  const unsigned item = item0 + grp;
  const bool active = item < n;
  const size_t bi = active ? A.list[item] : 0;
  const int bx = (int)(bi % fd.bxs), by = (int)(bi / fd.bxs);
  const int orig = active ? A.raw_qf[bi] : 1;
  // ---- all three channels' pixels start their way to shared memory (only the transform's own R x C pixels; the rest of
  // each tile is zero-filled)
  SX::sync(bar_id);                          // (the previous item's last reads of the tiles)
  {
    constexpr int CR = N / 4;                // 16-byte chunks per tile row; the group's N lanes cover four rows per instruction
    const int fr = l / CR, fc = l % CR;
    const bool col_ok = active && 4 * fc < G::C;
    const size_t off = col_ok ? (size_t)(by * 8 + fr) * fd.pitch + (size_t)bx * 8 + 4 * fc : 0;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const int c = k == 0 ? 1 : (k == 1 ? 0 : 2);
      const float* plane = c == 0 ? A.X : (c == 1 ? A.Y : A.B);
      float* dst = base + c * CG::kSq + fr * P + 4 * fc;
#pragma unroll
      for (int i = 0; i < N / 4; ++i) {
        const bool ok = col_ok && 4 * i + fr < G::R;
        cp_async16(dst + i * 4 * P, ok ? plane + off + (size_t)(4 * i) * fd.pitch : plane, ok);
      }
      cp_async_commit();
    }
  }
  float thr_y[4] = {0.58f, 0.64f, 0.64f, 0.64f};
  int maxq = 0;
  // ---- pass A: per channel transform, DC from the lowest frequencies, quant adjust
#pragma unroll 1
  for (int it = 0; it < 3; ++it) {
    const int c = it == 0 ? 1 : (it == 1 ? 0 : 2);
    float* T = base + c * CG::kSq;
    if (it == 0) cp_async_wait<2>(); else if (it == 1) cp_async_wait<1>(); else cp_async_wait<0>();
    SX::sync(bar_id);
    float u[N];
    {
      // rows (lane = pixel row), then columns (lane = horizontal frequency)
      float v[N];
#pragma unroll
      for (int j = 0; j < N / 4; ++j) {
        const float4 q4 = *reinterpret_cast<const float4*>(T + l * P + 4 * j);
        v[4 * j] = q4.x; v[4 * j + 1] = q4.y; v[4 * j + 2] = q4.z; v[4 * j + 3] = q4.w;
      }
      if (l < G::R) {   // (the other rows are zero and stay zero)
        if constexpr (MODE == kModeTall2) dct1d<N / 2>(v);
        else if constexpr (MODE == kModeTall4) dct1d<N / 4>(v);
        else dct1d<N>(v);
      }
#pragma unroll
      for (int j = 0; j < N / 4; ++j) *reinterpret_cast<float4*>(T + l * P + 4 * j) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    }
    SX::sync(bar_id);
#pragma unroll
    for (int y = 0; y < N; ++y) u[y] = T[y * P + l];
    if (owner) {
      if constexpr (MODE == kModeWide2) dct1d<N / 2>(u);
      else if constexpr (MODE == kModeWide4) dct1d<N / 4>(u);
      else dct1d<N>(u);
    }
    // lowest frequencies -> DC of the covered blocks (oracle DcFromLowestFrequencies): lane hf < cxb scales its cyb
    // values and inverts them vertically, lane y < cyb then inverts row y horizontally
    float* lc = llf + c * (G::cxb * G::cyb);
    if (l < G::cxb) {
      float t[G::cyb];
#pragma unroll
      for (int vf = 0; vf < G::cyb; ++vf) t[vf] = u[vf] * resample_scale(G::R, vf) * resample_scale(G::C, l);
      idct1d<G::cyb>(t);
#pragma unroll
      for (int y = 0; y < G::cyb; ++y) lc[y * G::cxb + l] = t[y];
    }
    SX::sync(bar_id);                        // (every lane has also read its column by now: the tile can take the coefficients)
    if (l < G::cyb) {
      float t[G::cxb];
#pragma unroll
      for (int x = 0; x < G::cxb; ++x) t[x] = lc[l * G::cxb + x];
      idct1d<G::cxb>(t);
#pragma unroll
      for (int x = 0; x < G::cxb; ++x) lc[l * G::cxb + x] = t[x];     // the DC values of block row l
    }
    if (owner) {
#pragma unroll
      for (int j = 0; j < G::VALS / 4; ++j) *reinterpret_cast<float4*>(T + l * P + 4 * j) = make_float4(u[4 * j], u[4 * j + 1], u[4 * j + 2], u[4 * j + 3]);
    }
    if (A.adjust) {
      float thr[4] = {0.58f, 0.64f, 0.64f, 0.64f};
      const float mulc = c == 0 ? A.x_qm_mul : (c == 1 ? 1.0f : A.b_qm_mul);
      const int q = adjust_quant_sq<N, MODE>(T + l * P, A.w + (size_t)c * G::LANES * G::VALS + (size_t)l * 4, c, scale, mulc, orig, thr, l, owner, xch, bar_id);
      maxq = max(maxq, q);
      if (c == 1) { thr_y[0] = thr[0]; thr_y[1] = thr[1]; thr_y[2] = thr[2]; thr_y[3] = thr[3]; }
    }
  }
  int quant = orig;
  if (A.adjust) {
    // every lane computed the same maximum (the sums are group-wide)
    quant = maxq;
  } else { thr_y[0] = 0.56f; thr_y[1] = thr_y[2] = thr_y[3] = 0.62f; }
  // ---- DC values and the integer quant field of the covered blocks
  SX::sync(bar_id);
  if (active && l < G::cyb) {
#pragma unroll
    for (int x = 0; x < G::cxb; ++x) {
      const size_t bj = bi + (size_t)l * fd.bxs + x;
      const int o = l * G::cxb + x;
      store_dc(A, bj, llf[0 * (G::cxb * G::cyb) + o], llf[1 * (G::cxb * G::cyb) + o], llf[2 * (G::cxb * G::cyb) + o]);
      A.raw_qf[bj] = quant;
    }
  }
  // ---- pass B: quantise Y; its dequantised value feeds X and B; non-zero values are scattered in scan order
  const float qac = scale * (float)quant;
  const float inv_qac = inv_gs / (float)quant;
  const int tx = bx >> 3, ty = by >> 3;
  const float x_factor = 0.0f + (float)A.cmap[(size_t)ty * fd.txs + tx] / 84.0f;
  const float b_factor = 1.0f + (float)A.cmap[(size_t)fd.txs * fd.tys + (size_t)ty * fd.txs + tx] / 84.0f;
  const uint16_t* invl = A.inv + (size_t)l * 4;
  const float* dqY = A.dq + (size_t)1 * G::LANES * G::VALS + (size_t)l * 4;
  float* rowY = base + 1 * CG::kSq + l * P;
  constexpr int log2n = G::ncov == 2 ? 1 : (G::ncov == 4 ? 2 : (G::ncov == 8 ? 3 : (G::ncov == 16 ? 4 : (G::ncov == 32 ? 5 : 6))));
  constexpr int log2cxb = G::cxb == 1 ? 0 : (G::cxb == 2 ? 1 : (G::cxb == 4 ? 2 : 3));
  // (a transform never straddles an AC group: its blocks share the group, block jj sits jj / cxb rows and jj % cxb columns in)
  const size_t blk0 = (size_t)((by >> 5) * fd.gxs + (bx >> 5)) * kGroupBlocks + (size_t)(by & 31) * 32 + (bx & 31);
  const bool first = G::kWide ? (l < G::W / 2) : (l < G::HS / 2);
#pragma unroll 1
  for (int it = 0; it < 3; ++it) {
    const int c = it == 0 ? 1 : (it == 1 ? 0 : 2);
    const int slot = it;
    float thr[4] = {0.58f, 0.64f, 0.64f, 0.64f};
    if (c == 1) { thr[0] = thr_y[0]; thr[1] = thr_y[1]; thr[2] = thr_y[2]; thr[3] = thr_y[3]; }
    else if (G::xs * G::ys >= 4) {
#pragma unroll
      for (int i = 0; i < 4; ++i) { thr[i] -= 0.00744f * (float)(G::xs * G::ys); if (thr[i] < 0.5f) thr[i] = 0.5f; }
    }
    const float thrA = G::kWide ? (first ? thr[0] : thr[1]) : (first ? thr[0] : thr[2]);
    const float thrB = G::kWide ? (first ? thr[2] : thr[3]) : (first ? thr[1] : thr[3]);
    const float qac_mul = qac * (c == 0 ? A.x_qm_mul : (c == 1 ? 1.0f : A.b_qm_mul));
    const float factor = c == 0 ? x_factor : b_factor;
    const float* wl = A.w + (size_t)c * G::LANES * G::VALS + (size_t)l * 4;
    const float* row = base + c * CG::kSq + l * P;
    int16_t* cdst = A.coeffs + (blk0 * 3 + slot) * 64;
    int nz = 0, last = 0;
    if (owner) {
#pragma unroll 1
      for (int j4 = 0; j4 < G::VALS; j4 += 4) {
        const size_t co = (size_t)(j4 >> 2) * G::LANES * 4;
        const float4 w4 = __ldg(reinterpret_cast<const float4*>(wl + co));
        const uint2 i4 = __ldg(reinterpret_cast<const uint2*>(invl + co));
        const float4 y4 = *reinterpret_cast<const float4*>(rowY + j4);
        float4 d4 = make_float4(0.0f, 0.0f, 0.0f, 0.0f), s4 = y4;
        if (c == 1) d4 = __ldg(reinterpret_cast<const float4*>(dqY + co));
        else s4 = *reinterpret_cast<const float4*>(row + j4);
        const float wv[4] = {w4.x, w4.y, w4.z, w4.w}, dv[4] = {d4.x, d4.y, d4.z, d4.w};
        const float yv[4] = {y4.x, y4.y, y4.z, y4.w}, sv[4] = {s4.x, s4.y, s4.z, s4.w};
        const uint32_t iv[4] = {i4.x & 0xFFFFu, i4.x >> 16, i4.y & 0xFFFFu, i4.y >> 16};
        const float t = j4 >= G::VALS / 2 ? thrB : thrA;
        float deq[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int j = j4 + e;
          const int x = G::kWide ? l : j, y = G::kWide ? j : l;
          float in;
          if (c == 1) in = yv[e];
          else in = __fmaf_rn(-factor, yv[e], sv[e]);                   // the Y row holds the dequantised Y by now
          const float val = (wv[e] * qac_mul) * in;
          int q = (fabsf(val) >= t) ? (int)rintf(val) : 0;
          if (x < G::xs && y < G::ys) q = 0;
          q = q > 32767 ? 32767 : (q < -32767 ? -32767 : q);
          deq[e] = c == 1 ? (quant_bias(1, q) * dv[e]) * inv_qac : 0.0f;
          if (q != 0 && active) {
            const int k = (int)iv[e];
            const int jj = k >> 6;
            cdst[(size_t)((jj >> log2cxb) * 32 + (jj & (G::cxb - 1))) * 192 + (k & 63)] = (int16_t)q;
            ++nz; last = max(last, k);
          }
        }
        if (c == 1) *reinterpret_cast<float4*>(rowY + j4) = make_float4(deq[0], deq[1], deq[2], deq[3]);
      }
    }
    float cnt[1] = {(float)nz};                                       // <= 4096: exact in float
    lane_sums<N, G::LANES, 1>(cnt, xch, l, bar_id);
    float lastf[1] = {(float)last};
    lane_maxs<N, G::LANES, 1>(lastf, xch, l, bar_id);
    if (active && l == 0) {
      const int nzt = (int)cnt[0];
      const int shared = (nzt + G::ncov - 1) >> log2n;
      A.nzcount[(size_t)c * nblk + bi] = (uint16_t)nzt;
      A.lastk[(size_t)c * nblk + bi] = (uint16_t)(int)lastf[0];
      for (int jj = 0; jj < G::ncov; ++jj) A.nzeros[(size_t)c * nblk + bi + (size_t)(jj / G::cxb) * fd.bxs + (jj % G::cxb)] = (uint8_t)shared;
    }
  }
}

// ------------------------------------------------------------------------------------------------ kernels
// One launch covers all the lists of a size class: a CTA walks "virtual CTAs" (list, chunk of the list), so the
// strategies of a frame are transformed side by side instead of one short launch after the other.
struct CoeffAllArgs {
  CoeffArgs base;                                   // everything but the per-list fields
  const float* w[kNumLists]; const float* dq[kNumLists]; const uint16_t* inv[kNumLists];
  const uint32_t* lists;                            // [16 counters][kNumLists][nblk]
  unsigned nblk;
};

__device__ __forceinline__ CoeffArgs list_args(const CoeffAllArgs& AA, int li) {
  CoeffArgs A = AA.base;
  A.w = AA.w[li]; A.dq = AA.dq[li]; A.inv = AA.inv[li];
  A.count = AA.lists + li; A.list = AA.lists + kListHeader + (size_t)li * AA.nblk;
  return A;
}

// DCT2X2 and IDENTITY (rare; not separable along lanes): one thread per block
__global__ void __launch_bounds__(64) k_coeff8_special(CoeffAllArgs AA) {
  unsigned ctas[2], total = 0;
#pragma unroll
  for (int m = 0; m < 2; ++m) { ctas[m] = (AA.lists[kListID + m] + 63) / 64; total += ctas[m]; }
  for (unsigned v = blockIdx.x; v < total; v += gridDim.x) {
    const int m = v >= ctas[0] ? 1 : 0;
    const unsigned i = (v - (m ? ctas[0] : 0)) * 64 + threadIdx.x;
    if (i >= AA.lists[kListID + m]) continue;
    const CoeffArgs A = list_args(AA, kListID + m);
    if (m == 0) coeff8_body<kStratIDENTITY>(A, i); else coeff8_body<kStratDCT2X2>(A, i);
  }
}

// DCT, DCT4X4, DCT4X8, DCT8X4: a warp takes four list entries per round; virtual CTAs of 16 entries walk the lists
__global__ void __launch_bounds__(128) k_coeff8_lanes(CoeffAllArgs AA) {
  extern __shared__ __align__(16) float smem_f[];
  const int lists[4] = {kListDCT, kList4X4, kList4X8, kList8X4};
  unsigned ctas[4], total = 0;
#pragma unroll
  for (int m = 0; m < 4; ++m) { ctas[m] = (AA.lists[lists[m]] + 15) / 16; total += ctas[m]; }
  float* smem_w = smem_f + (threadIdx.x >> 5) * kC8WarpFloats;
  for (unsigned v = blockIdx.x; v < total; v += gridDim.x) {
    unsigned rel = v;
    int m = 0;
    while (rel >= ctas[m]) { rel -= ctas[m]; ++m; }
    const unsigned item0 = rel * 16 + (threadIdx.x >> 5) * 4;
    if (item0 >= AA.lists[lists[m]]) continue;       // (warp-uniform)
    const CoeffArgs A = list_args(AA, lists[m]);
    if (m == 0) coeff8_lanes_body<kStratDCT>(A, smem_w, item0);
    else if (m == 1) coeff8_lanes_body<kStratDCT4X4>(A, smem_w, item0);
    else if (m == 2) coeff8_lanes_body<kStratDCT4X8>(A, smem_w, item0);
    else coeff8_lanes_body<kStratDCT8X4>(A, smem_w, item0);
  }
}

template <int N>
__global__ void __launch_bounds__(CoeffSqGeom<N>::kThreads) k_coeffsq_all(CoeffAllArgs AA) {
  using CG = CoeffSqGeom<N>;
  extern __shared__ __align__(16) float smem_f[];
  constexpr int first = N == 16 ? kList16Tall : (N == 32 ? kList32Tall : kList64Tall);   // tall, wide, square
  constexpr int nlists = N == 32 ? 5 : 3;                                                  // N = 32: + DCT32X8, DCT8X32
  unsigned ctas[5], total = 0;
#pragma unroll
  for (int m = 0; m < nlists; ++m) { ctas[m] = (AA.lists[m < 3 ? first + m : kList32Tall4 + m - 3] + CG::kGroups - 1) / CG::kGroups; total += ctas[m]; }
  for (unsigned v = blockIdx.x; v < total; v += gridDim.x) {
    unsigned rel = v;
    int m = 0;
    while (rel >= ctas[m]) { rel -= ctas[m]; ++m; }
    const CoeffArgs A = list_args(AA, m < 3 ? first + m : kList32Tall4 + m - 3);
    if (m == 0) coeffsq_body<N, kModeTall2>(A, smem_f, rel * CG::kGroups);
    else if (m == 1) coeffsq_body<N, kModeWide2>(A, smem_f, rel * CG::kGroups);
    else if (m == 2) coeffsq_body<N, kModeSq>(A, smem_f, rel * CG::kGroups);
    else if constexpr (N == 32) {
      if (m == 3) coeffsq_body<N, kModeTall4>(A, smem_f, rel * CG::kGroups);
      else coeffsq_body<N, kModeWide4>(A, smem_f, rel * CG::kGroups);
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------ host side
void launch_coeff_lists(const uint8_t* acs, const FrameDim& fd, uint32_t* lists, cudaStream_t s) {
  const size_t nblk = (size_t)fd.bxs * fd.bys;
  cudaMemsetAsync(lists, 0, kListHeader * 4, s);
  ++g_kernel_launches;
  k_coeff_lists<<<(unsigned)((nblk + 255) / 256), 256, 0, s>>>(acs, (int)nblk, lists);
}

template <int N>
static void launch_coeffsq_all(const CoeffAllArgs& AA, cudaStream_t s) {
  using CG = CoeffSqGeom<N>;
  const size_t smem = CG::kSmemFloats * sizeof(float);
  cudaFuncSetAttribute(k_coeffsq_all<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const size_t max_items = (size_t)AA.nblk / (N * N / 256) + 5;          // the smallest transforms of the class are N x N/4
  size_t grid = (max_items + CG::kGroups - 1) / CG::kGroups;
  if (grid > 148 * 8) grid = 148 * 8;
  ++g_kernel_launches;
  k_coeffsq_all<N><<<(unsigned)grid, CG::kThreads, smem, s>>>(AA);
}

void launch_coeff_general(const float* x, const float* y, const float* b, const uint8_t* acs, const FrameDim& fd,
                          const QuantDev* qd, const AcsTables& T, const uint16_t* const* inv_order, const int8_t* cmap,
                          float x_qm_mul, float b_qm_mul, int adjust, int32_t* raw_qf, int16_t* coeffs, int16_t* dc_quant,
                          uint8_t* nzeros, uint16_t* nzcount, uint16_t* lastk, uint32_t* lists, cudaStream_t s,
                          const StreamFork* fork) {
  const size_t nblk = (size_t)fd.bxs * fd.bys;
  cudaMemsetAsync(coeffs, 0, (size_t)fd.num_groups * kGroupBlocks * 192 * sizeof(int16_t), s);
  launch_coeff_lists(acs, fd, lists, s);
  CoeffAllArgs AA;
  CoeffArgs& A = AA.base;
  A.X = x; A.Y = y; A.B = b; A.fd = fd; A.qd = qd; A.cmap = cmap; A.x_qm_mul = x_qm_mul; A.b_qm_mul = b_qm_mul; A.adjust = adjust;
  A.raw_qf = raw_qf; A.coeffs = coeffs; A.dc_quant = dc_quant; A.nzeros = nzeros; A.nzcount = nzcount; A.lastk = lastk;
  A.w = nullptr; A.dq = nullptr; A.inv = nullptr; A.list = nullptr; A.count = nullptr;
  AA.lists = lists; AA.nblk = (unsigned)nblk;
  // tables in lane order per list; inv_order: [order class] natural, [13..15] transposed for the wide strategy of class 4 / 6 / 8
  // 8x8 lists: DCT, DCT4X4, DCT4X8, DCT8X4 in the lane order of the eight-lane kernel (AcsTables::w8); IDENTITY, DCT2X2 stored order
  AA.w[kListDCT] = T.w8[0]; AA.dq[kListDCT] = T.dq8[0]; AA.w[kList4X4] = T.w8[1]; AA.dq[kList4X4] = T.dq8[1];
  AA.w[kList4X8] = T.w8[2]; AA.dq[kList4X8] = T.dq8[2]; AA.w[kList8X4] = T.w8[3]; AA.dq[kList8X4] = T.dq8[3];
  AA.w[kListID] = T.w[1]; AA.dq[kListID] = T.dq[1]; AA.w[kList2X2] = T.w[2]; AA.dq[kList2X2] = T.dq[2];
  for (int li = 0; li < 6; ++li) AA.inv[li] = nullptr;
  // 16 / 32 / 64-sized lists: tables and inverse scan orders in 16-byte chunks, [c][chunk][lane][4] (AcsTables::wJ)
  for (int li = kList16Tall; li < kNumLists; ++li) { AA.w[li] = T.wJ[li - kList16Tall]; AA.dq[li] = T.dqJ[li - kList16Tall]; AA.inv[li] = T.invJ[li - kList16Tall]; }
  // the five launches are independent of each other (disjoint transforms): a lone frame runs them on three streams
  cudaStream_t s8 = s, s32 = s;
  if (fork) {
    cudaEventRecord(fork->fork, s);
    cudaStreamWaitEvent(fork->aux[0], fork->fork, 0);
    cudaStreamWaitEvent(fork->aux[1], fork->fork, 0);
    s8 = fork->aux[1]; s32 = fork->aux[0];
  }
  g_kernel_launches += 2;
  size_t g8 = (nblk + 15) / 16 + 4;
  if (g8 > 148 * 16) g8 = 148 * 16;
  launch_coeffsq_all<64>(AA, s);     // (longest dependency chains first)
  launch_coeffsq_all<32>(AA, s32);
  launch_coeffsq_all<16>(AA, s32);
  k_coeff8_lanes<<<(unsigned)g8, 128, 4 * kC8WarpFloats * sizeof(float), s8>>>(AA);
  k_coeff8_special<<<(unsigned)((nblk + 63) / 64 + 2 > 148 * 8 ? 148 * 8 : (nblk + 63) / 64 + 2), 64, 0, s8>>>(AA);
  if (fork) {
    cudaEventRecord(fork->join[0], fork->aux[0]);
    cudaEventRecord(fork->join[1], fork->aux[1]);
    cudaStreamWaitEvent(s, fork->join[0], 0);
    cudaStreamWaitEvent(s, fork->join[1], 0);
  }
}

}  // namespace jxlb
