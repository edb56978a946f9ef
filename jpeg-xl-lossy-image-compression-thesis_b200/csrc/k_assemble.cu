// K12 — frame assembly on the device (stage U9): HfGlobal (context map + cluster histograms),
// codestream / frame headers, TOC and the concatenation of every section into the output
// buffer.  Field orders follow libjxl headers.cc, frame_header.cc, enc_toc.cc, enc_frame.cc,
// enc_context_map.cc [UPSTREAM]; encoder choices are listed in DESIGN.md "Bitstream".
#include "entropy.cuh"
#include "kernels.h"

namespace jxlb {

// ---- HfGlobal: context map (prefix coded, no MTF: every entry's bit length is known independently,
// so the 7425 entries are placed by a CTA-wide scan and written in parallel) + the per-cluster
// uint configs and histogram headers, concatenated with a cooperative bit copy.
__global__ void __launch_bounds__(256) k_hf_global(const uint8_t* __restrict__ cmap, const int* __restrict__ num_clusters_p,
                                                   const uint32_t* __restrict__ hdr_bits, const uint32_t* __restrict__ hdr_len,
                                                   int num_groups, uint32_t* __restrict__ cm_words, uint32_t* __restrict__ hf_words,
                                                   uint32_t* __restrict__ hf_bits) {
  __shared__ uint32_t s_hist[kModAlphabet];
  __shared__ HuffScratch s_hs;
  __shared__ uint8_t s_len[kModAlphabet];
  __shared__ uint16_t s_code[kModAlphabet];
  __shared__ int s_alpha;
  __shared__ uint32_t s_pc_hdr[20];
  __shared__ uint32_t s_pc_hdr_bits;
  __shared__ uint32_t s_head[32], s_tail[16];
  __shared__ uint32_t s_head_bits, s_tail_bits, s_cm_bits;
  __shared__ uint32_t s_warp[8];
  __shared__ unsigned long long s_piece_dst[kMaxClusters + 1];
  constexpr int kCmWords = 8192;
  constexpr int kPer = (kNumAcContexts + 255) / 256;  // 30 entries per thread
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int K = *num_clusters_p;
  for (int i = t; i < kModAlphabet; i += 256) s_hist[i] = 0;
  if (t < 20) s_pc_hdr[t] = 0;
  for (int i = t; i < kCmWords; i += 256) cm_words[i] = 0;
  for (int i = t; i < 8192 + kMaxClusters * 64 + 256; i += 256) hf_words[i] = 0;
  __syncthreads();
  uint32_t cm_bits = 0;
  if (K > 1) {
    for (int i = t; i < kNumAcContexts; i += 256) {
      uint32_t tk, nb, bits;
      hybrid_encode(cmap[i], tk, nb, bits);
      atomicAdd(&s_hist[tk], 1u);
    }
    __syncthreads();
    if (warp == 0) {
      build_prefix_code_warp(s_hist, s_hs, s_len, s_code, &s_alpha, lane);
      const uint32_t hb = write_prefix_header_warp(s_len, s_alpha, s_pc_hdr, lane);
      if (lane == 0) s_pc_hdr_bits = hb;
    }
    __syncthreads();
    // per-thread chunk of consecutive entries: total length, CTA scan, then write
    const int i0 = min(t * kPer, kNumAcContexts), i1 = min(i0 + kPer, kNumAcContexts);
    uint32_t sum = 0;
    for (int i = i0; i < i1; ++i) { uint32_t tk, nb, bits; hybrid_encode(cmap[i], tk, nb, bits); sum += s_len[tk] + nb; }
    uint32_t incl = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const uint32_t o = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += o; }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    uint32_t pos = incl - sum;
    for (int w = 0; w < warp; ++w) pos += s_warp[w];
    if (t == 255) s_cm_bits = pos + sum;
    for (int i = i0; i < i1; ++i) {
      uint32_t tk, nb, bits;
      hybrid_encode(cmap[i], tk, nb, bits);
      const uint32_t cl = s_len[tk];
      const unsigned long long v = ((unsigned long long)s_code[tk] | ((unsigned long long)bits << cl)) << (pos & 31);
      atomicOr(&cm_words[pos >> 5], (uint32_t)v);
      if ((pos & 31) + cl + nb > 32) atomicOr(&cm_words[(pos >> 5) + 1], (uint32_t)(v >> 32));
      pos += cl + nb;
    }
    __syncthreads();
    cm_bits = s_cm_bits;
  }
  if (t == 0) {
    BitWriterDev w;
    w.init(s_head);
    w.write(1, 1);                                                  // default dequant matrices
    const int lg = num_groups <= 1 ? 0 : 32 - __clz(num_groups - 1);
    w.write(lg, 0);                                                 // num_histograms - 1
    w.write(2, 2);                                                  // used_orders = 0
    w.write(1, 0);                                                  // lz77 disabled
    if (K == 1) { w.write(1, 1); w.write(2, 0); }
    else {
      w.write(1, 0); w.write(1, 0);                                 // not simple, no MTF
      w.write(1, 0); w.write(1, 1);                                 // nested code: no lz77, prefix code
      w.write(4, 4); w.write(3, 2); w.write(2, 0);
      w.var_len_uint16((uint32_t)(s_alpha - 1));
      w.append(s_pc_hdr, s_pc_hdr_bits);
    }
    w.flush();
    s_head_bits = w.bits();
    BitWriterDev w2;
    w2.init(s_tail);
    w2.write(1, 0);                                                 // ANS
    w2.write(2, kLogAlphaSize - 5);
    for (int k = 0; k < K; ++k) { w2.write(4, 4); w2.write(3, 2); w2.write(2, 0); }
    w2.flush();
    s_tail_bits = w2.bits();
    unsigned long long pos = (unsigned long long)s_head_bits + cm_bits + s_tail_bits;
    for (int k = 0; k < K; ++k) { s_piece_dst[k] = pos; pos += hdr_len[k]; }
    s_piece_dst[K] = pos;
    *hf_bits = (uint32_t)pos;
  }
  __syncthreads();
  cta_bitcopy(hf_words, 0, s_head, 0, s_head_bits, t, 256);
  cta_bitcopy(hf_words, s_head_bits, cm_words, 0, cm_bits, t, 256);
  cta_bitcopy(hf_words, (unsigned long long)s_head_bits + cm_bits, s_tail, 0, s_tail_bits, t, 256);
  for (int k = warp; k < K; k += 8) cta_bitcopy(hf_words, s_piece_dst[k], hdr_bits + (size_t)k * 64, 0, hdr_len[k], lane, 32);
}

// ---- headers + TOC + section placement
__device__ void write_size_u32(BitWriterDev& w, uint32_t v) {
  const uint32_t m = v - 1;
  if (m < (1u << 9)) { w.write(2, 0); w.write(9, m); }
  else if (m < (1u << 13)) { w.write(2, 1); w.write(13, m); }
  else if (m < (1u << 18)) { w.write(2, 2); w.write(18, m); }
  else { w.write(2, 3); w.write(30, m); }
}

__global__ void __launch_bounds__(256) k_finalize(FrameDim fd, int x_qm_scale, int b_qm_scale, int gab, const uint32_t* __restrict__ lf_bits,
                                                 const uint32_t* __restrict__ dg_start_bit, const uint32_t* __restrict__ mod_total_bits,
                                                 const uint32_t* __restrict__ hf_bits, const unsigned long long* __restrict__ group_start_bit,
                                                 Section* __restrict__ sections, uint32_t* __restrict__ hdr_stage,
                                                 uint32_t* __restrict__ out_words, unsigned long long out_capacity_bits,
                                                 unsigned long long* __restrict__ out_info, const QuantDev* __restrict__ qd,
                                                 const uint32_t* __restrict__ token_counts,
                                                 const int* __restrict__ num_clusters_p) {
  if (threadIdx.x < 32) {  // statistics block read back by the host together with the size (one pinned copy)
    unsigned long long nt = 0;
    for (int g = threadIdx.x; g < fd.num_groups; g += 32) nt += token_counts[g];
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) nt += __shfl_xor_sync(0xffffffffu, nt, d);
    if (threadIdx.x == 0) {
      out_info[2] = (unsigned long long)qd->global_scale; out_info[3] = (unsigned long long)qd->quant_dc;
      out_info[4] = nt; out_info[5] = (unsigned long long)*num_clusters_p;
    }
  }
  // section sources and lengths: one thread per section (a serial loop paid one global-load round trip per section);
  // the lengths stay in shared memory for the TOC and the placement below
  extern __shared__ unsigned long long s_nbits[];   // [nsec] lengths, then [nsec] destination bits
  const int ndc = fd.num_dc_groups, ng = fd.num_groups;
  const bool small = ng == 1;
  const int nsec = 2 + ndc + ng;
  unsigned long long* s_dst = s_nbits + nsec;
  for (int i = threadIdx.x; i < nsec; i += blockDim.x) {
    Section s;
    if (i == 0) { s.kind = 0; s.index = 0; s.src_bit = 0; s.nbits = *lf_bits; }
    else if (i <= ndc) {
      const int dg = i - 1;
      const unsigned long long b0 = dg_start_bit[dg], b1 = dg + 1 < ndc ? dg_start_bit[dg + 1] : *mod_total_bits;
      s.kind = 1; s.index = dg; s.src_bit = b0; s.nbits = b1 - b0;
    } else if (i == ndc + 1) { s.kind = 2; s.index = 0; s.src_bit = 0; s.nbits = *hf_bits; }
    else {
      const int g = i - ndc - 2;
      s.kind = 3; s.index = g; s.src_bit = group_start_bit[g]; s.nbits = (unsigned long long)kTokensPerGroupMax * 32 - s.src_bit;
    }
    s.dst_bit = 0;
    sections[i] = s;
    s_nbits[i] = s.nbits;
  }
  __syncthreads();
  __shared__ unsigned long long s_pos;
  __shared__ uint32_t s_header_bits;
  __shared__ int s_overflow;
  if (threadIdx.x == 0) {
  // codestream headers, frame header and TOC into the staging words
  BitWriterDev w;
  w.init(hdr_stage);
  w.write(8, 0xFF); w.write(8, 0x0A);
  w.write(1, 0);                                  // SizeHeader: not small
  write_size_u32(w, (uint32_t)fd.ysize);
  w.write(3, 0);                                  // ratio 0: explicit xsize
  write_size_u32(w, (uint32_t)fd.xsize);
  w.write(1, 1); w.write(1, 1);                   // ImageMetadata.all_default, default_m
  if (w.bits() & 7) w.write(8 - (int)(w.bits() & 7), 0);
  w.write(1, 0); w.write(2, 0); w.write(1, 0);    // frame header: not all_default, regular frame, VarDCT
  w.write(2, 2); w.write(8, 128 - 17);            // flags = kSkipAdaptiveDCSmoothing (U64 selector 2)
  w.write(2, 0);                                  // upsampling 1
  w.write(3, (uint32_t)x_qm_scale); w.write(3, (uint32_t)b_qm_scale);
  w.write(2, 0); w.write(1, 0); w.write(2, 0);    // one pass, no crop, blend replace
  w.write(1, 1); w.write(2, 0);                   // is_last, no name
  w.write(1, 0); w.write(1, gab ? 1u : 0u);       // loop filter: not all_default, gab
  if (gab) w.write(1, 0);                         //   default Gaborish weights
  w.write(2, 0); w.write(2, 0);                   //   epf 0, no extensions
  w.write(2, 0);                                  // no frame-header extensions
  w.write(1, 0);                                  // TOC not permuted
  if (w.bits() & 7) w.write(8 - (int)(w.bits() & 7), 0);
  unsigned long long total_small_bits = 0;
  if (small) for (int i = 0; i < nsec; ++i) total_small_bits += s_nbits[i];
  const int ntoc = small ? 1 : nsec;
  for (int i = 0; i < ntoc; ++i) {
    const uint32_t size = (uint32_t)(((small ? total_small_bits : s_nbits[i]) + 7) >> 3);
    if (size < 1024) { w.write(2, 0); w.write(10, size); }
    else if (size < 17408) { w.write(2, 1); w.write(14, size - 1024); }
    else if (size < 4211712) { w.write(2, 2); w.write(22, size - 17408); }
    else { w.write(2, 3); w.write(30, size - 4211712); }
  }
  if (w.bits() & 7) w.write(8 - (int)(w.bits() & 7), 0);
  w.flush();
  const uint32_t header_bits = w.bits();
  unsigned long long pos = header_bits;
  for (int i = 0; i < nsec; ++i) {
    s_dst[i] = pos;
    pos += small ? s_nbits[i] : ((s_nbits[i] + 7) & ~7ull);
  }
  pos = (pos + 7) & ~7ull;
  const bool overflow = pos + 64 > out_capacity_bits;
  out_info[0] = overflow ? 0 : pos >> 3;
  out_info[1] = overflow ? 1 : 0;
  s_pos = pos; s_header_bits = header_bits; s_overflow = overflow ? 1 : 0;
  }
  __syncthreads();
  if (s_overflow) return;
  // the words a section covers only partially are ORed into by k_assemble: clear them first
  for (int i = threadIdx.x; i < nsec; i += blockDim.x) {
    const unsigned long long a = s_dst[i], b = a + s_nbits[i];
    sections[i].dst_bit = a;
    out_words[a >> 5] = 0;
    out_words[b >> 5] = 0;
  }
  if (threadIdx.x == 0) out_words[s_pos >> 5] = 0;
  __syncthreads();   // (the header words below overlap the first section's first word)
  for (uint32_t i = threadIdx.x; i < (s_header_bits + 31) / 32; i += blockDim.x) out_words[i] = hdr_stage[i];
}

__global__ void __launch_bounds__(256) k_assemble(const Section* __restrict__ sections, const uint32_t* __restrict__ lf_words,
                                                  const uint32_t* __restrict__ mod_words, const uint32_t* __restrict__ hf_words,
                                                  const uint32_t* __restrict__ group_arena, uint32_t* __restrict__ out_words,
                                                  const unsigned long long* __restrict__ out_info) {
  if (out_info[1]) return;
  const Section s = sections[blockIdx.y];
  if (s.nbits == 0) return;
  const uint32_t* src = s.kind == 0 ? lf_words : (s.kind == 1 ? mod_words : (s.kind == 2 ? hf_words
                                   : group_arena + (size_t)s.index * kTokensPerGroupMax));
  const unsigned long long d0 = s.dst_bit, d1 = s.dst_bit + s.nbits;
  const unsigned long long w0 = d0 >> 5, w1 = (d1 + 31) >> 5;
  for (unsigned long long wi = w0 + blockIdx.x * 256 + threadIdx.x; wi < w1; wi += (unsigned long long)gridDim.x * 256) {
    const unsigned long long lo = max(wi << 5, d0), hi = min((wi + 1) << 5, d1);
    const int n = (int)(hi - lo);
    const unsigned long long sp = s.src_bit + (lo - d0);
    const uint32_t a = src[sp >> 5];
    const int sh = (int)(sp & 31);
    uint32_t v = a >> sh;
    if (sh + n > 32) v |= src[(sp >> 5) + 1] << (32 - sh);
    if (n < 32) v &= (1u << n) - 1;
    v <<= (int)(lo & 31);
    if (n == 32) out_words[wi] = v; else atomicOr(&out_words[wi], v);
  }
}

// ------------------------------------------------------------------------------------------ launchers
void launch_hf_global(const uint8_t* cmap, const int* num_clusters, const uint32_t* hdr_bits, const uint32_t* hdr_len,
                      int num_groups, uint32_t* cm_back, uint32_t* hf_words, uint32_t* hf_bits, cudaStream_t s) {
  ++g_kernel_launches;
  k_hf_global<<<1, 256, 0, s>>>(cmap, num_clusters, hdr_bits, hdr_len, num_groups, cm_back, hf_words, hf_bits);
}

void launch_finalize(const FrameDim& fd, int x_qm_scale, int b_qm_scale, int gab, const uint32_t* lf_bits, const uint32_t* dg_start_bit,
                     const uint32_t* mod_total_bits, const uint32_t* hf_bits, const unsigned long long* group_start_bit,
                     Section* sections, uint32_t* hdr_stage, uint32_t* out_words, unsigned long long out_capacity_bits,
                     unsigned long long* out_info, const QuantDev* qd, const uint32_t* token_counts, const int* num_clusters,
                     cudaStream_t s) {
  ++g_kernel_launches;
  const size_t smem = (size_t)2 * (2 + fd.num_dc_groups + fd.num_groups) * sizeof(unsigned long long);
  if (smem > 48 * 1024) cudaFuncSetAttribute(k_finalize, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  k_finalize<<<1, 256, smem, s>>>(fd, x_qm_scale, b_qm_scale, gab, lf_bits, dg_start_bit, mod_total_bits, hf_bits, group_start_bit,
                             sections, hdr_stage, out_words, out_capacity_bits, out_info, qd, token_counts, num_clusters);
}

void launch_assemble(const Section* sections, int num_sections, const uint32_t* lf_words, const uint32_t* mod_words,
                     const uint32_t* hf_words, const uint32_t* group_arena, uint32_t* out_words,
                     const unsigned long long* out_info, cudaStream_t s) {
  ++g_kernel_launches;
  dim3 grid(8, num_sections);
  k_assemble<<<grid, 256, 0, s>>>(sections, lf_words, mod_words, hf_words, group_arena, out_words, out_info);
}

}  // namespace jxlb
