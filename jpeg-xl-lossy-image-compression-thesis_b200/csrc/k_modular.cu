// K11 — modular sub-bitstreams of the VarDCT frame: quantised DC of every 2048x2048 DC group and
// the AC metadata (chroma-from-luma map, AC strategy, quant field, EPF sharpness), coded with one
// fixed global MA tree and prefix (Huffman) codes so that every token's bit length is known
// independently and the streams are written fully in parallel (an rANS chain over the ~4e5
// tokens of a DC group would serialise for milliseconds).  Format: libjxl enc_modular.cc /
// modular/encoding / dec_huffman.cc [UPSTREAM]; tree and code construction: DESIGN.md.
//
// Element space: every DC group owns a fixed-capacity run of "elements"
//   [head][3*w*h DC residuals Y,X,B][mid][tw*th ytox][tw*th ytob][w*h strategy][w*h qf][w*h epf]
// (strategy / qf slots beyond the group's first-block count are invalid = 0 bits), so a single
// exclusive scan of element bit lengths places every element of every LfGroup section.
#include "entropy.cuh"
#include "kernels.h"

namespace jxlb {

constexpr uint32_t kTokHead = 0xFE000000u;
constexpr uint32_t kTokMid = 0xFD000000u;
constexpr int kScanTile = 2048;

__device__ __forceinline__ int clamped_gradient(int w, int n, int nw) {
  const int m = min(w, n), M = max(w, n), g = w + n - nw;
  return g < m ? m : (g > M ? M : g);
}

__device__ __forceinline__ int find_dg(const DcGroupInfo* __restrict__ dgs, int num_dg, uint32_t idx) {
  int dg = 0;
  while (dg + 1 < num_dg && idx >= dgs[dg + 1].elem_base) ++dg;
  return dg;
}

// ---- first-block ranks inside each DC group (raster order) and the compacted strategy / qf rows
__global__ void __launch_bounds__(1024) k_mod_ranks(const uint8_t* __restrict__ acs, const int32_t* __restrict__ raw_qf,
                                                    FrameDim fd, const DcGroupInfo* __restrict__ dgs,
                                                    int32_t* __restrict__ strat_c, int32_t* __restrict__ qf_c,
                                                    uint32_t* __restrict__ first_count,
                                                    unsigned long long* __restrict__ acs_hist) {
  __shared__ uint32_t s_warp[32];
  __shared__ uint32_t s_carry;
  __shared__ uint32_t s_acs_hist[32];   // first blocks per AcStrategy code (jxlb200_stats.acs_histogram)
  const DcGroupInfo d = dgs[blockIdx.x];
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int total = d.w * d.h;
  if (t == 0) s_carry = 0;
  if (t < 32) s_acs_hist[t] = 0;
  __syncthreads();
  // 4 consecutive raster positions per thread and iteration (their loads are issued together)
  for (int i0 = 0; i0 < total; i0 += 4096) {
    uint8_t a[4]; size_t bi[4];
    uint32_t cnt = 0;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = i0 + t * 4 + u;
      a[u] = 0; bi[u] = 0;
      if (i < total) {
        const int y = i / d.w, x = i - y * d.w;
        bi[u] = (size_t)(d.y0 + y) * fd.bxs + d.x0 + x;
        a[u] = acs[bi[u]];
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) cnt += (a[u] >> 7) & 1;
    uint32_t incl = cnt;
#pragma unroll
    for (int dd = 1; dd < 32; dd <<= 1) { const uint32_t o = __shfl_up_sync(0xffffffffu, incl, dd); if (lane >= dd) incl += o; }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    uint32_t base = s_carry + incl - cnt;
    for (int w = 0; w < warp; ++w) base += s_warp[w];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (a[u] & 0x80) {
        strat_c[d.block_base + base] = a[u] & 0x7f;
        qf_c[d.block_base + base] = raw_qf[bi[u]] - 1;
        atomicAdd(&s_acs_hist[a[u] & 31], 1u);
        ++base;
      }
    }
    __syncthreads();
    if (t == 1023) s_carry = base;
    __syncthreads();
  }
  if (t == 0) first_count[blockIdx.x] = s_carry;
  if (t < 27 && s_acs_hist[t]) atomicAdd(&acs_hist[t], (unsigned long long)s_acs_hist[t]);
}

// ---- tokens (leaf << 24 | packed residual) + per-leaf histograms
__global__ void __launch_bounds__(256) k_mod_tokens(const int16_t* __restrict__ dc_quant, const int8_t* __restrict__ cmap,
                                                    const int32_t* __restrict__ strat_c, const int32_t* __restrict__ qf_c,
                                                    const uint32_t* __restrict__ first_count, FrameDim fd,
                                                    const DcGroupInfo* __restrict__ dgs, int num_dg, uint32_t total_elems,
                                                    uint32_t* __restrict__ tokens, uint32_t* __restrict__ mod_hist) {
  __shared__ uint32_t s_hist[kNumModularCtx * kModAlphabet];
  for (int i = threadIdx.x; i < kNumModularCtx * kModAlphabet; i += 256) s_hist[i] = 0;
  __syncthreads();
  const size_t nblk = (size_t)fd.bxs * fd.bys;
  for (uint32_t idx = blockIdx.x * 256 + threadIdx.x; idx < total_elems; idx += gridDim.x * 256) {
    const int dg = find_dg(dgs, num_dg, idx);
    const DcGroupInfo d = dgs[dg];
    uint32_t e = idx - d.elem_base;
    const uint32_t wh = (uint32_t)(d.w * d.h), tt = (uint32_t)(d.tw * d.th);
    uint32_t tok = kInvalidToken;
    if (e == 0) tok = kTokHead;
    else if (e < 1 + 3 * wh) {
      e -= 1;
      const int ch = e / wh;               // modular channel: 0 = Y, 1 = X, 2 = B
      const uint32_t r = e - ch * wh;
      const int x = r % d.w, y = r / d.w;
      const int plane = ch == 0 ? 1 : (ch == 1 ? 0 : 2);
      const int16_t* p = dc_quant + (size_t)plane * nblk + (size_t)(d.y0 + y) * fd.bxs + d.x0 + x;
      const int v = p[0];
      const int W = x ? p[-1] : (y ? p[-fd.bxs] : 0);
      const int N = y ? p[-fd.bxs] : W;
      const int NW = (x && y) ? p[-fd.bxs - 1] : W;
      const int leaf = ch == 0 ? kLeafDcY : (ch == 1 ? kLeafDcX : kLeafDcB);
      tok = ((uint32_t)leaf << 24) | pack_signed(v - clamped_gradient(W, N, NW));
    } else if (e == 1 + 3 * wh) tok = kTokMid;
    else {
      e -= 2 + 3 * wh;
      if (e < 2 * tt) {
        const int m = e / tt;
        const uint32_t r = e - m * tt;
        const int x = r % d.tw, y = r / d.tw;
        const int v = cmap[(size_t)m * fd.txs * fd.tys + (size_t)((d.y0 >> 3) + y) * fd.txs + (d.x0 >> 3) + x];
        tok = ((uint32_t)(m == 0 ? kLeafYtoX : kLeafYtoB) << 24) | pack_signed(v);
      } else {
        e -= 2 * tt;
        const uint32_t count = first_count[dg];
        if (e < wh) { if (e < count) tok = ((uint32_t)kLeafAcs << 24) | pack_signed(strat_c[d.block_base + e]); }
        else if (e < 2 * wh) {
          e -= wh;
          if (e < count) {
            const int prev = e ? qf_c[d.block_base + e - 1] : strat_c[d.block_base];
            tok = ((uint32_t)kLeafQf << 24) | pack_signed(qf_c[d.block_base + e] - prev);
          }
        } else tok = ((uint32_t)kLeafEpf << 24) | 0u;
      }
    }
    tokens[idx] = tok;
    if (tok < 0xFD000000u) {
      uint32_t tk, nb, bits;
      hybrid_encode(tok & 0xFFFFFF, tk, nb, bits);
      atomicAdd(&s_hist[(tok >> 24) * kModAlphabet + tk], 1u);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < kNumModularCtx * kModAlphabet; i += 256) { const uint32_t v = s_hist[i]; if (v) atomicAdd(&mod_hist[i], v); }
}

// ---- Huffman codes of the 8 leaves (one warp each) and the LfGlobal section
// global MA tree + its single-histogram ANS code (constant given num_dc_groups; see the oracle's
// WriteGlobalTree for the node table).  Serial, ~80 tokens; runs on one lane.
__device__ void write_global_tree(int num_dc_groups, BitWriterDev& w, uint16_t* s_scratch, uint16_t* s_rmap, AnsSymInfo* s_info,
                                  uint32_t* s_back) {
  const int props[15] = {1, 0, 0, 0, 0, 0, -1, -1, 2, -1, -1, -1, -1, -1, -1};
  const int vals[15] = {2 * num_dc_groups, 1, 0, 2, 0, 1, 5, 0, 0, 0, 0, 5, 5, 1, 0};
  uint32_t toks[80];
  int nt = 0;
  for (int i = 0; i < 15; ++i) {
    if (props[i] < 0) { toks[nt++] = 0; toks[nt++] = (uint32_t)vals[i]; toks[nt++] = 0; toks[nt++] = 0; toks[nt++] = 0; }
    else { toks[nt++] = (uint32_t)(props[i] + 1); toks[nt++] = pack_signed(vals[i]); }
  }
  w.write(1, 0); w.write(1, 1); w.write(2, 0); w.write(1, 0); w.write(2, kLogAlphaSize - 5);
  w.write(4, 4); w.write(3, 2); w.write(2, 0);
  uint32_t counts[kAcAlphabet];
  for (int s = 0; s < kAcAlphabet; ++s) counts[s] = 0;
  for (int i = 0; i < nt; ++i) { uint32_t tk, nb, bits; hybrid_encode(toks[i], tk, nb, bits); counts[tk]++; }
  uint16_t norm[kAcAlphabet];
  normalize_counts(counts, kAcAlphabet, norm);
  write_ans_histogram(norm, kAcAlphabet, w);
  rmap_fill(s_scratch, s_info, alias_serial(norm, kAcAlphabet, s_scratch, s_info), s_rmap, 0, 1);
  BackWriterDev bw;
  bw.init(s_back, 128);
  uint32_t state = kAnsInitState;
  for (int i = nt - 1; i >= 0; --i) {
    uint32_t tk, nb, bits;
    hybrid_encode(toks[i], tk, nb, bits);
    bw.push((int)nb, bits, true);
    uint32_t o16;
    if (ans_put(state, s_info[tk].freq, s_info[tk].rcp, s_rmap + s_info[tk].base, o16)) bw.push(16, o16, true);
  }
  bw.push(32, state, true);
  const long long sb = bw.finish(true);
  // append bits [sb, 128*32) of s_back
  for (long long p = sb; p < 128 * 32;) {
    const int n = (int)min((long long)(32 - (p & 31)), 128 * 32 - p);
    w.write(n, (s_back[p >> 5] >> (p & 31)) & (n == 32 ? 0xFFFFFFFFu : ((1u << n) - 1)));
    p += n;
  }
}

// the tree blob depends only on num_dc_groups: built once per geometry and cached by the encoder
__global__ void __launch_bounds__(32) k_tree_blob(int num_dc_groups, uint32_t* __restrict__ tree_words, uint32_t* __restrict__ tree_bits) {
  __shared__ uint16_t s_scratch[1024];
  __shared__ uint16_t s_rmap[kAnsTabSize];
  __shared__ AnsSymInfo s_info[kAcAlphabet];
  __shared__ uint32_t s_back[128];
  if (threadIdx.x != 0) return;
  BitWriterDev w;
  w.init(tree_words);
  write_global_tree(num_dc_groups, w, s_scratch, s_rmap, s_info, s_back);
  w.flush();
  *tree_bits = w.bits();
}

__global__ void __launch_bounds__(256) k_mod_codes(const uint32_t* __restrict__ mod_hist, const QuantDev* __restrict__ qd,
                                                   const uint32_t* __restrict__ tree_words, const uint32_t* __restrict__ tree_bits,
                                                   uint8_t* __restrict__ code_len, uint16_t* __restrict__ code_bits,
                                                   uint32_t* __restrict__ lf_words, uint32_t* __restrict__ lf_bits) {
  __shared__ HuffScratch s_hs[kNumModularCtx];
  __shared__ uint8_t s_len[kNumModularCtx][kModAlphabet];
  __shared__ uint16_t s_bits[kNumModularCtx][kModAlphabet];
  __shared__ int s_alpha[kNumModularCtx];
  __shared__ uint32_t s_hdr[kNumModularCtx][20];
  __shared__ uint32_t s_hdr_bits[kNumModularCtx];
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  if (lane < 20) s_hdr[warp][lane] = 0;
  __syncwarp();
  build_prefix_code_warp(mod_hist + warp * kModAlphabet, s_hs[warp], s_len[warp], s_bits[warp], &s_alpha[warp], lane);
  const uint32_t hb = write_prefix_header_warp(s_len[warp], s_alpha[warp], s_hdr[warp], lane);
  if (lane == 0) s_hdr_bits[warp] = hb;
  __syncthreads();
  for (int i = t; i < kNumModularCtx * kModAlphabet; i += 256) { code_len[i] = s_len[i / kModAlphabet][i % kModAlphabet]; code_bits[i] = s_bits[i / kModAlphabet][i % kModAlphabet]; }
  if (t == 0) {
    BitWriterDev w;
    w.init(lf_words);
    w.write(1, 1);  // default DC dequantisation
    const uint32_t gs = (uint32_t)qd->global_scale;
    if (gs < 2049) { w.write(2, 0); w.write(11, gs - 1); }
    else if (gs < 4097) { w.write(2, 1); w.write(11, gs - 2049); }
    else if (gs < 8193) { w.write(2, 2); w.write(12, gs - 4097); }
    else { w.write(2, 3); w.write(16, gs - 8193); }
    const uint32_t q = (uint32_t)qd->quant_dc;
    if (q == 16) { w.write(2, 0); }
    else if (q <= 32) { w.write(2, 1); w.write(5, q - 1); }
    else if (q <= 256) { w.write(2, 2); w.write(8, q - 1); }
    else { w.write(2, 3); w.write(16, q - 1); }
    w.write(1, 1); w.write(1, 1);  // default block context map, default colour correlation
    w.write(1, 1);                 // global MA tree present
    w.append(tree_words, *tree_bits);
    w.write(1, 0);                 // lz77 disabled
    w.write(1, 1); w.write(2, 3);  // simple context map, 3 bits per entry
    for (int l = 0; l < kNumModularCtx; ++l) w.write(3, (uint32_t)l);
    w.write(1, 1);                 // prefix codes
    for (int l = 0; l < kNumModularCtx; ++l) { w.write(4, 4); w.write(3, 2); w.write(2, 0); }
    for (int l = 0; l < kNumModularCtx; ++l) w.var_len_uint16((uint32_t)(s_alpha[l] - 1));
    for (int l = 0; l < kNumModularCtx; ++l) w.append(s_hdr[l], s_hdr_bits[l]);
    w.flush();
    *lf_bits = w.bits();
  }
}

// ---- (length, bits) of every element + per-tile bit totals
__device__ __forceinline__ unsigned long long element_code(uint32_t tok, const uint8_t* __restrict__ code_len,
                                                           const uint16_t* __restrict__ code_bits, int wh, uint32_t count) {
  if (tok == kInvalidToken) return 0ull;
  if (tok == kTokHead) return (6ull << 32) | 12ull;
  if (tok == kTokMid) {
    const int L = wh <= 1 ? 0 : 32 - __clz(wh - 1);
    return ((unsigned long long)(L + 4) << 32) | (unsigned long long)((count - 1) | (3u << L));
  }
  uint32_t tk, nb, bits;
  hybrid_encode(tok & 0xFFFFFF, tk, nb, bits);
  const int leaf = tok >> 24;
  const uint32_t cl = code_len[leaf * kModAlphabet + tk];
  const unsigned long long v = (unsigned long long)code_bits[leaf * kModAlphabet + tk] | ((unsigned long long)bits << cl);
  return ((unsigned long long)(cl + nb) << 32) | v;   // cl <= 15, nb <= 22 -> v < 2^37: keep the low 32 + note below
}

// NOTE on widths: modular residuals here are < 2^19 (DC is clamped to int16, qf < 2^9), so nb <= 17 and
// cl + nb <= 32; the code word fits 32 bits.

__global__ void __launch_bounds__(256) k_mod_lengths(const uint32_t* __restrict__ tokens, const uint8_t* __restrict__ code_len,
                                                     const uint16_t* __restrict__ code_bits, const DcGroupInfo* __restrict__ dgs,
                                                     int num_dg, const uint32_t* __restrict__ first_count, uint32_t total_elems,
                                                     uint32_t* __restrict__ tile_sums) {
  __shared__ uint32_t s_warp[8];
  const uint32_t base = blockIdx.x * kScanTile;
  uint32_t sum = 0;
  for (int i = 0; i < kScanTile / 256; ++i) {
    const uint32_t idx = base + i * 256 + threadIdx.x;
    if (idx >= total_elems) break;
    const uint32_t tok = tokens[idx];
    int wh = 0; uint32_t count = 0;
    if (tok == kTokMid) { const int dg = find_dg(dgs, num_dg, idx); wh = dgs[dg].w * dgs[dg].h; count = first_count[dg]; }
    sum += (uint32_t)(element_code(tok, code_len, code_bits, wh, count) >> 32);
  }
#pragma unroll
  for (int d = 16; d >= 1; d >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, d);
  if ((threadIdx.x & 31) == 0) s_warp[threadIdx.x >> 5] = sum;
  __syncthreads();
  if (threadIdx.x == 0) { uint32_t s = 0; for (int w = 0; w < 8; ++w) s += s_warp[w]; tile_sums[blockIdx.x] = s; }
}

__global__ void __launch_bounds__(1024) k_scan_tiles(uint32_t* __restrict__ tile_sums, int num_tiles, uint32_t* __restrict__ total_bits) {
  __shared__ uint32_t s_warp[32];
  __shared__ uint32_t s_carry;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  if (t == 0) s_carry = 0;
  __syncthreads();
  for (int i0 = 0; i0 < num_tiles; i0 += 1024) {
    const int i = i0 + t;
    const uint32_t v = i < num_tiles ? tile_sums[i] : 0;
    uint32_t incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const uint32_t o = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += o; }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    uint32_t base = s_carry + incl - v;
    for (int w = 0; w < warp; ++w) base += s_warp[w];
    if (i < num_tiles) tile_sums[i] = base;
    __syncthreads();
    if (t == 1023) s_carry = base + v;
    __syncthreads();
  }
  if (t == 0) *total_bits = s_carry;
}

// ---- place every element: tile offset + in-tile exclusive scan, bits ORed into the stream words
__global__ void __launch_bounds__(256) k_mod_write(const uint32_t* __restrict__ tokens, const uint8_t* __restrict__ code_len,
                                                   const uint16_t* __restrict__ code_bits, const DcGroupInfo* __restrict__ dgs,
                                                   int num_dg, const uint32_t* __restrict__ first_count, uint32_t total_elems,
                                                   const uint32_t* __restrict__ tile_offsets, uint32_t* __restrict__ words,
                                                   uint32_t* __restrict__ dg_start_bit) {
  __shared__ uint32_t s_warp[8];
  __shared__ uint32_t s_run;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  if (t == 0) s_run = tile_offsets[blockIdx.x];
  __syncthreads();
  const uint32_t base = blockIdx.x * kScanTile;
  for (int i = 0; i < kScanTile / 256; ++i) {
    const uint32_t idx = base + i * 256 + t;
    unsigned long long code = 0;
    bool is_head = false;
    if (idx < total_elems) {
      const uint32_t tok = tokens[idx];
      int wh = 0; uint32_t count = 0;
      if (tok == kTokMid) { const int dg = find_dg(dgs, num_dg, idx); wh = dgs[dg].w * dgs[dg].h; count = first_count[dg]; }
      is_head = tok == kTokHead;
      code = element_code(tok, code_len, code_bits, wh, count);
    }
    const uint32_t len = (uint32_t)(code >> 32);
    uint32_t incl = len;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const uint32_t o = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += o; }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    uint32_t pos = s_run + incl - len;
    for (int w = 0; w < warp; ++w) pos += s_warp[w];
    if (is_head) dg_start_bit[find_dg(dgs, num_dg, idx)] = pos;
    if (len) {
      const unsigned long long v = (code & 0xFFFFFFFFull) << (pos & 31);
      atomicOr(&words[pos >> 5], (uint32_t)v);
      if ((pos & 31) + len > 32) atomicOr(&words[(pos >> 5) + 1], (uint32_t)(v >> 32));
    }
    __syncthreads();
    if (t == 255) s_run = pos + len;
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------ launchers
void launch_mod_ranks(const uint8_t* acs, const int32_t* raw_qf, const FrameDim& fd, const DcGroupInfo* dgs, int num_dg,
                      int32_t* strat_c, int32_t* qf_c, uint32_t* first_count, unsigned long long* acs_hist, cudaStream_t s) {
  ++g_kernel_launches;
  k_mod_ranks<<<num_dg, 1024, 0, s>>>(acs, raw_qf, fd, dgs, strat_c, qf_c, first_count, acs_hist);
}
void launch_mod_tokens(const int16_t* dc_quant, const int8_t* cmap, const int32_t* strat_c, const int32_t* qf_c,
                       const uint32_t* first_count, const FrameDim& fd, const DcGroupInfo* dgs, int num_dg,
                       uint32_t total_elems, uint32_t* tokens, uint32_t* mod_hist, cudaStream_t s) {
  ++g_kernel_launches;
  const uint32_t want = (total_elems + 255) / 256;
  const int grid = (int)(want < 148u * 8u ? want : 148u * 8u);
  k_mod_tokens<<<grid, 256, 0, s>>>(dc_quant, cmap, strat_c, qf_c, first_count, fd, dgs, num_dg, total_elems, tokens, mod_hist);
}
void launch_tree_blob(int num_dc_groups, uint32_t* tree_words, uint32_t* tree_bits, cudaStream_t s) {
  ++g_kernel_launches;
  k_tree_blob<<<1, 32, 0, s>>>(num_dc_groups, tree_words, tree_bits);
}
void launch_mod_codes(const uint32_t* mod_hist, const QuantDev* qd, const uint32_t* tree_words, const uint32_t* tree_bits,
                      uint8_t* code_len, uint16_t* code_bits, uint32_t* lf_words, uint32_t* lf_bits, cudaStream_t s) {
  ++g_kernel_launches;
  k_mod_codes<<<1, 256, 0, s>>>(mod_hist, qd, tree_words, tree_bits, code_len, code_bits, lf_words, lf_bits);
}
void launch_mod_write(const uint32_t* tokens, const uint8_t* code_len, const uint16_t* code_bits, const DcGroupInfo* dgs,
                      int num_dg, const uint32_t* first_count, uint32_t total_elems, uint32_t* tile_sums, uint32_t* total_bits,
                      uint32_t* words, uint32_t* dg_start_bit, cudaStream_t s) {
  const int tiles = (int)((total_elems + kScanTile - 1) / kScanTile);
  g_kernel_launches += 3;
  k_mod_lengths<<<tiles, 256, 0, s>>>(tokens, code_len, code_bits, dgs, num_dg, first_count, total_elems, tile_sums);
  k_scan_tiles<<<1, 1024, 0, s>>>(tile_sums, tiles, total_bits);
  k_mod_write<<<tiles, 256, 0, s>>>(tokens, code_len, code_bits, dgs, num_dg, first_count, total_elems, tile_sums, words,
                                    dg_start_bit);
}

}  // namespace jxlb
