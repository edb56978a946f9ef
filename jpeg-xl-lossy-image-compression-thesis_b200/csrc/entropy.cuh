// jxlb200 — device-side building blocks of the entropy stage (stages U6-U9): hybrid-uint
// tokens, fixed-point log2 for the clustering cost, a forward bit writer, ANS histogram
// normalisation / header coding, alias tables + reverse maps.  Format side follows libjxl
// enc_ans.cc / ans_common.cc / dec_ans.cc (ISO/IEC 18181-1 Annex C) [UPSTREAM]; see DESIGN.md
// "Entropy stage".  Everything here is integer arithmetic, so results do not depend on the
// order in which threads combine partial sums.
#pragma once
#include "jxl_common.cuh"

namespace jxlb {

constexpr int kAnsLogTabSize = 12;
constexpr int kAnsTabSize = 1 << kAnsLogTabSize;
constexpr uint32_t kAnsInitState = 0x13u << 16;
constexpr int kLogAlphaSize = 8;
constexpr int kAcAlphabet = 64;
constexpr int kModAlphabet = 128;
constexpr int kMaxClusters = 24;  // 24 reverse maps x 8 KB = 192 KB: one SM's shared memory holds them all
constexpr int kNumOrders = 13;
constexpr int kNonZeroBuckets = 37;
constexpr int kZeroDensityContextCount = 458;
constexpr int kNumBlockCtx = 15;
constexpr int kNumAcContexts = kNumBlockCtx * (kNonZeroBuckets + kZeroDensityContextCount);  // 7425
constexpr int kNumModularCtx = 8;
constexpr int kTokensPerGroupMax = 3 * 1024 * 64;  // every block-channel: 1 nzeros token + at most 63 coefficients
constexpr uint32_t kInvalidToken = 0xFFFFFFFFu;

enum ModLeaf { kLeafDcY = 0, kLeafEpf = 1, kLeafYtoB = 2, kLeafYtoX = 3, kLeafDcB = 4, kLeafDcX = 5, kLeafQf = 6, kLeafAcs = 7 };

__host__ __device__ __forceinline__ uint32_t pack_signed(int v) { return v >= 0 ? 2u * (uint32_t)v : 2u * (uint32_t)(-v) - 1u; }

// hybrid uint config (4, 2, 0)
__device__ __forceinline__ void hybrid_encode(uint32_t v, uint32_t& tok, uint32_t& nbits, uint32_t& bits) {
  if (v < 16) { tok = v; nbits = 0; bits = 0; return; }
  const uint32_t n = 31 - __clz(v);
  const uint32_t m = v - (1u << n);
  tok = 16 + ((n - 4) << 2) + (m >> (n - 2));
  nbits = n - 2;
  bits = v & ((1u << (n - 2)) - 1);
}

// log2 in Q20 through a 1025-entry table (uploaded once by the host)
__device__ __forceinline__ long long log2_q20(uint32_t n, const int* __restrict__ lut) {
  const int e = 31 - __clz(n);
  const uint32_t m = n << (31 - e);
  const uint32_t idx = (m >> 21) & 1023;
  const uint32_t frac = (m >> 5) & 0xFFFF;
  const long long a = lut[idx], b = lut[idx + 1];
  return ((long long)e << 20) + a + (((b - a) * (long long)frac) >> 16);
}
__device__ __forceinline__ long long xlogx(uint32_t n, const int* __restrict__ lut) {
  return n ? (long long)n * log2_q20(n, lut) : 0;
}

// forward LSB-first bit writer used by single threads (headers, histograms)
struct BitWriterDev {
  uint32_t* words;
  unsigned long long acc;
  int cnt;
  uint32_t wpos;
  __device__ void init(uint32_t* w) { words = w; acc = 0; cnt = 0; wpos = 0; }
  __device__ void write(int nbits, uint32_t value) {
    acc |= (unsigned long long)value << cnt;
    cnt += nbits;
    if (cnt >= 32) { words[wpos++] = (uint32_t)acc; acc >>= 32; cnt -= 32; }
  }
  __device__ void var_len_uint8(uint32_t n) {
    if (n == 0) { write(1, 0); return; }
    write(1, 1);
    const int nb = 31 - __clz(n);
    write(3, (uint32_t)nb);
    write(nb, n - (1u << nb));
  }
  __device__ void var_len_uint16(uint32_t n) {
    if (n == 0) { write(1, 0); return; }
    write(1, 1);
    const int nb = 31 - __clz(n);
    write(4, (uint32_t)nb);
    write(nb, n - (1u << nb));
  }
  __device__ uint32_t bits() const { return wpos * 32 + (uint32_t)cnt; }
  __device__ void flush() { if (cnt > 0) words[wpos] = (uint32_t)acc; }
  // appends `nbits` bits of another word buffer (bit 0 of src first)
  __device__ void append(const uint32_t* src, uint32_t nbits) {
    uint32_t i = 0;
    for (; i + 32 <= nbits; i += 32) write(32, src[i >> 5]);
    if (i < nbits) write((int)(nbits - i), src[i >> 5] & ((1u << (nbits - i)) - 1));
  }
};

// counts -> frequencies summing to 4096; serial version (one thread), same rule as the oracle's
// NormalizeCounts: round to nearest with a floor of 1, then move the excess onto the largest bins
__device__ inline void normalize_counts(const uint32_t* counts, int alphabet, uint16_t* norm) {
  unsigned long long total = 0;
  for (int s = 0; s < alphabet; ++s) total += counts[s];
  for (int s = 0; s < alphabet; ++s) norm[s] = 0;
  if (total == 0) { norm[0] = kAnsTabSize; return; }
  long long sum = 0;
  for (int s = 0; s < alphabet; ++s) {
    if (!counts[s]) continue;
    unsigned long long t = ((unsigned long long)counts[s] * (2 * kAnsTabSize) + total) / (2 * total);
    if (t < 1) t = 1;
    norm[s] = (uint16_t)t;
    sum += (long long)t;
  }
  long long delta = kAnsTabSize - sum;
  while (delta != 0) {
    int best = 0;
    for (int s = 1; s < alphabet; ++s) if (norm[s] > norm[best]) best = s;
    if (delta > 0) { norm[best] = (uint16_t)(norm[best] + delta); delta = 0; }
    else {
      long long take = -delta < (long long)norm[best] - 1 ? -delta : (long long)norm[best] - 1;
      norm[best] = (uint16_t)(norm[best] - take);
      delta += take;
    }
  }
}

// ANS histogram header (shift = 13, no RLE); serial
__device__ inline void write_ans_histogram(const uint16_t* norm, int alphabet, BitWriterDev& w) {
  const uint8_t kLen[14] = {5, 4, 4, 4, 4, 4, 3, 3, 3, 3, 3, 6, 7, 7};
  const uint8_t kSym[14] = {17, 11, 15, 3, 9, 7, 4, 2, 5, 6, 0, 33, 1, 65};
  int nsym = 0, s0 = 0, s1 = 0, last = -1;
  for (int s = 0; s < alphabet; ++s) if (norm[s]) { if (nsym == 0) s0 = s; else if (nsym == 1) s1 = s; ++nsym; last = s; }
  if (nsym <= 2) {
    w.write(1, 1);
    w.write(1, nsym == 2 ? 1u : 0u);
    if (nsym == 0) { w.var_len_uint8(0); return; }
    w.var_len_uint8((uint32_t)s0);
    if (nsym == 2) { w.var_len_uint8((uint32_t)s1); w.write(kAnsLogTabSize, norm[s0]); }
    return;
  }
  w.write(1, 0); w.write(1, 0); w.write(3, 7); w.write(3, 6);
  const int length = last + 1;
  w.var_len_uint8((uint32_t)(length - 3));
  int omit_pos = -1, omit_log = -1;
  for (int i = 0; i < length; ++i) {
    const int lc = norm[i] ? (31 - __clz((uint32_t)norm[i])) + 1 : 0;
    if (lc > omit_log) { omit_log = lc; omit_pos = i; }
    w.write(kLen[lc], kSym[lc]);
  }
  for (int i = 0; i < length; ++i) {
    const int lc = norm[i] ? (31 - __clz((uint32_t)norm[i])) + 1 : 0;
    if (i == omit_pos || lc <= 1) continue;
    w.write(lc - 1, (uint32_t)norm[i] - (1u << (lc - 1)));
  }
}

// encoder-side symbol info: freq, start of the symbol's offsets in the reverse map, reciprocal
struct AnsSymInfo { uint16_t freq; uint16_t base; uint32_t rcp; };

// Alias table of libjxl's InitAliasTable (normative) for log_alpha_size = 8 and the encoder's
// per-symbol info.  Serial (one thread).  scratch: 4 x 256 uint16.  Returns the index of the
// symbol that owns the whole table (freq 4096) or -1.
__device__ inline int alias_serial(const uint16_t* norm, int alphabet, uint16_t* scratch, AnsSymInfo* info) {
  uint16_t* cutoffs = scratch;          // [256]
  uint16_t* right = scratch + 256;      // [256]
  uint16_t* offsets1 = scratch + 512;   // [256]
  uint16_t* stack = scratch + 768;      // underfull [0..), overfull grows down from 255
  const int entry_size = 1 << (kAnsLogTabSize - kLogAlphaSize);
  const int table_size = 1 << kLogAlphaSize;
  int n = alphabet;
  while (n > 0 && norm[n - 1] == 0) --n;
  int single = -1;
  for (int s = 0; s < n; ++s) if (norm[s] == kAnsTabSize) single = s;
  if (single < 0) {
    int nu = 0, no = 0;  // the two stacks share one 256-entry array: nu + no <= 256 always holds
    for (int i = 0; i < table_size; ++i) {
      const int c = i < n ? norm[i] : 0;
      cutoffs[i] = (uint16_t)c; right[i] = (uint16_t)i; offsets1[i] = 0;
      if (c > entry_size) { stack[255 - no] = (uint16_t)i; ++no; }
      else if (c < entry_size) { stack[nu++] = (uint16_t)i; }
    }
    while (no > 0) {
      const int o = stack[255 - (no - 1)]; --no;
      const int u = stack[nu - 1]; --nu;
      const int by = entry_size - cutoffs[u];
      cutoffs[o] = (uint16_t)(cutoffs[o] - by);
      right[u] = (uint16_t)o;
      offsets1[u] = cutoffs[o];
      if (cutoffs[o] < entry_size) stack[nu++] = (uint16_t)o;
      else if (cutoffs[o] > entry_size) { stack[255 - no] = (uint16_t)o; ++no; }
    }
  }
  uint32_t acc = 0;
  for (int s = 0; s < alphabet; ++s) {
    const uint32_t f = norm[s];
    info[s].freq = (uint16_t)f; info[s].base = (uint16_t)acc;
    info[s].rcp = f <= 1 ? 0xFFFFFFFFu : (uint32_t)(0x100000000ull / f);
    acc += f;
  }
  return single;
}

// reverse map slot(symbol, offset) for table slots v = first, first + step, ...
__device__ inline void rmap_fill(const uint16_t* scratch, const AnsSymInfo* info, int single, uint16_t* rmap, int first, int step) {
  const uint16_t* cutoffs = scratch; const uint16_t* right = scratch + 256; const uint16_t* offsets1 = scratch + 512;
  const int entry_size = 1 << (kAnsLogTabSize - kLogAlphaSize);
  for (int v = first; v < kAnsTabSize; v += step) {
    int sym, off;
    if (single >= 0) { sym = single; off = v; }
    else {
      const int i = v >> (kAnsLogTabSize - kLogAlphaSize), pos = v & (entry_size - 1);
      const int c = cutoffs[i];
      if (c == entry_size || pos < c) { sym = i; off = pos; }
      else { sym = right[i]; off = (int)offsets1[i] - c + pos; }
    }
    rmap[info[sym].base + off] = (uint16_t)v;
  }
}

// warp-cooperative: lane 0 builds the alias table, all lanes fill the reverse map
__device__ inline void build_reverse_map(const uint16_t* norm, int alphabet, uint16_t* scratch, uint16_t* rmap,
                                         AnsSymInfo* info, int lane) {
  int single = -1;
  if (lane == 0) single = alias_serial(norm, alphabet, scratch, info);
  __syncwarp();
  single = __shfl_sync(0xffffffffu, single, 0);
  rmap_fill(scratch, info, single, rmap, lane, 32);
  __syncwarp();
}

// one rANS step; returns true when 16 bits (`out16`) must be emitted before the symbol
__device__ __forceinline__ bool ans_put(uint32_t& state, uint32_t freq, uint32_t rcp, const uint16_t* __restrict__ rmap_sym,
                                        uint32_t& out16) {
  bool emit = false;
  if ((state >> (32 - kAnsLogTabSize)) >= freq) { out16 = state & 0xFFFF; state >>= 16; emit = true; }
  uint32_t q = __umulhi(state, rcp);
  uint32_t r = state - q * freq;
  if (r >= freq) { ++q; r -= freq; }
  state = (q << kAnsLogTabSize) + rmap_sym[r];
  return emit;
}

// ---- prefix (Huffman) codes, warp-cooperative ------------------------------------------------
// Deterministic Huffman, identical to the oracle's BuildPrefixCode: repeatedly merge the two
// smallest nodes by (weight, node index) — leaves are indexed by symbol (< 128), internal nodes by
// 128 + creation order, which is the oracle's (weight, id) order — with a doubling count floor until
// no code is longer than 15 bits; canonical codes, bit-reversed for the LSB-first stream.
struct HuffScratch { unsigned long long weight[2 * kModAlphabet]; short parent[2 * kModAlphabet]; uint8_t alive[2 * kModAlphabet]; };

__device__ __forceinline__ bool huff_less(unsigned long long wa, int ia, unsigned long long wb, int ib) {
  return wa < wb || (wa == wb && ia < ib);
}

__device__ inline void build_prefix_code_warp(const uint32_t* counts_in, HuffScratch& hs, uint8_t* length, uint16_t* code_bits,
                                              int* alphabet_out, int lane) {
  const unsigned full = 0xffffffffu;
  uint32_t cnt[4];
  int used = 0, last = -1;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int s = lane + 32 * j;
    cnt[j] = counts_in[s];
    length[s] = 0; code_bits[s] = 0;
    if (cnt[j]) { ++used; last = s; }
  }
#pragma unroll
  for (int d = 16; d >= 1; d >>= 1) { used += __shfl_xor_sync(full, used, d); last = max(last, __shfl_xor_sync(full, last, d)); }
  if (used == 0 || (used == 1 && last == 0)) { if (lane == 0) *alphabet_out = 1; __syncwarp(); return; }
  if (used == 1) { if (lane == 0) cnt[0] = 1; used = 2; }   // a complex prefix code needs two coded symbols
  if (lane == 0) *alphabet_out = last + 1;
  for (uint32_t floor_count = 1;; floor_count *= 2) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int s = lane + 32 * j;
      hs.alive[s] = cnt[j] ? 1 : 0;
      hs.weight[s] = cnt[j] > floor_count ? cnt[j] : floor_count;
      hs.parent[s] = -1;
      hs.alive[128 + s] = 0; hs.parent[128 + s] = -1;
    }
    __syncwarp();
    int created = 0;
    for (int live = used; live > 1; --live) {
      // per-lane two smallest among nodes lane, lane + 32, ...
      unsigned long long w1 = ~0ull, w2 = ~0ull; int i1 = 0x7fff, i2 = 0x7fff;
      const int nn = 128 + created;
      for (int i = lane; i < nn; i += 32) {
        if (!hs.alive[i]) continue;
        const unsigned long long w = hs.weight[i];
        if (huff_less(w, i, w1, i1)) { w2 = w1; i2 = i1; w1 = w; i1 = i; }
        else if (huff_less(w, i, w2, i2)) { w2 = w; i2 = i; }
      }
      unsigned long long wa = w1; int ia = i1;
#pragma unroll
      for (int d = 16; d >= 1; d >>= 1) {
        const unsigned long long ow = __shfl_xor_sync(full, wa, d); const int oi = __shfl_xor_sync(full, ia, d);
        if (huff_less(ow, oi, wa, ia)) { wa = ow; ia = oi; }
      }
      unsigned long long wb = (i1 == ia) ? w2 : w1; int ib = (i1 == ia) ? i2 : i1;
#pragma unroll
      for (int d = 16; d >= 1; d >>= 1) {
        const unsigned long long ow = __shfl_xor_sync(full, wb, d); const int oi = __shfl_xor_sync(full, ib, d);
        if (huff_less(ow, oi, wb, ib)) { wb = ow; ib = oi; }
      }
      if (lane == 0) {
        hs.weight[nn] = wa + wb; hs.parent[nn] = -1; hs.alive[nn] = 1;
        hs.parent[ia] = (short)nn; hs.parent[ib] = (short)nn; hs.alive[ia] = 0; hs.alive[ib] = 0;
      }
      ++created;
      __syncwarp();
    }
    int maxlen = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int s = lane + 32 * j;
      int dpt = 0;
      if (cnt[j]) for (int v = s; hs.parent[v] >= 0; v = hs.parent[v]) ++dpt;
      length[s] = (uint8_t)dpt;
      maxlen = max(maxlen, dpt);
    }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) maxlen = max(maxlen, __shfl_xor_sync(full, maxlen, d));
    __syncwarp();
    if (maxlen <= 15) break;
  }
  if (lane == 0) {
    uint32_t next_code[17];
    uint32_t count[17];
    for (int l = 0; l < 17; ++l) count[l] = 0;
    for (int s = 0; s <= last; ++s) count[length[s]]++;
    uint32_t code = 0;
    for (int l = 1; l <= 15; ++l) { next_code[l] = code; code = (code + count[l]) << 1; }
    for (int s = 0; s <= last; ++s) {
      const int l = length[s];
      if (l) code_bits[s] = (uint16_t)(__brev(next_code[l]++) >> (32 - l));
    }
  }
  __syncwarp();
}

// complex prefix-code description (see the oracle's WritePrefixCodeHeader): 2 + 36 fixed bits, then
// one bit-reversed 4-bit length per symbol.  hdr: >= 18 zeroed words.  Returns the bit count.
__device__ inline uint32_t write_prefix_header_warp(const uint8_t* length, int alphabet, uint32_t* hdr, int lane) {
  if (alphabet <= 1) return 0;
  if (lane == 0) {
    // hskip = 0 (2 bits), then code-length-code lengths in order {1,2,3,4,0,5,17,6,16,7..15}: "01" for 4, "00" for 0
    unsigned long long fixed = 0;
    const int order[18] = {1, 2, 3, 4, 0, 5, 17, 6, 16, 7, 8, 9, 10, 11, 12, 13, 14, 15};
    for (int i = 0; i < 18; ++i) if (order[i] < 16) fixed |= 1ull << (2 + 2 * i);
    atomicOr(&hdr[0], (uint32_t)fixed);
    atomicOr(&hdr[1], (uint32_t)(fixed >> 32));
  }
  for (int s = lane; s < alphabet; s += 32) {
    const uint32_t nib = __brev((uint32_t)length[s]) >> 28;
    const int pos = 38 + 4 * s;
    const unsigned long long v = (unsigned long long)nib << (pos & 31);
    atomicOr(&hdr[pos >> 5], (uint32_t)v);
    if ((pos & 31) > 28) atomicOr(&hdr[(pos >> 5) + 1], (uint32_t)(v >> 32));
  }
  __syncwarp();
  return 38 + 4 * (uint32_t)alphabet;
}

// CTA-cooperative bit copy with OR semantics (dst words must start out zero where not yet written)
__device__ inline void cta_bitcopy(uint32_t* dst, unsigned long long dst_bit, const uint32_t* src, unsigned long long src_bit,
                                   unsigned long long nbits, int t, int nthreads) {
  if (nbits == 0) return;
  const unsigned long long d0 = dst_bit, d1 = dst_bit + nbits;
  for (unsigned long long wi = (d0 >> 5) + t; wi < ((d1 + 31) >> 5); wi += nthreads) {
    const unsigned long long lo = max(wi << 5, d0), hi = min((wi + 1) << 5, d1);
    const int n = (int)(hi - lo);
    const unsigned long long sp = src_bit + (lo - d0);
    const int sh = (int)(sp & 31);
    uint32_t v = src[sp >> 5] >> sh;
    if (sh + n > 32) v |= src[(sp >> 5) + 1] << (32 - sh);
    if (n < 32) v &= (1u << n) - 1;
    atomicOr(&dst[wi], v << (int)(lo & 31));
  }
}

// backward bit writer: chunks are prepended, so the stream read forward lists them in the
// reverse of the order they were pushed.  The stream ends at word `end_word` of `words`.
struct BackWriterDev {
  uint32_t* words;
  unsigned long long acc;  // bit i = stream bit pos + i
  int cnt;                 // valid bits in acc (< 32 between pushes)
  long long wptr;          // next word index to store (exclusive, moving down)
  __device__ void init(uint32_t* w, long long end_word) { words = w; acc = 0; cnt = 0; wptr = end_word; }
  __device__ void push(int nbits, uint32_t value, bool store) {
    acc = (acc << nbits) | value;
    cnt += nbits;
    if (cnt >= 32) {
      --wptr;
      if (store) words[wptr] = (uint32_t)(acc >> (cnt - 32));
      cnt -= 32;
      acc &= (1ull << cnt) - 1;
    }
  }
  // flushes the partial word; returns the stream's first bit position (in bits from words[0])
  __device__ long long finish(bool store) {
    if (cnt > 0) { --wptr; if (store) words[wptr] = (uint32_t)(acc << (32 - cnt)); return wptr * 32 + (32 - cnt); }
    return wptr * 32;
  }
};

}  // namespace jxlb
