// ORACLE (test infrastructure) — entropy-coding building blocks shared by the AC token coder
// (stage U6-U8), the modular DC / metadata coder and the self-decoder.
// Restates libjxl enc_ans.cc / ans_common.cc / dec_ans.cc / enc_huffman.cc (format side:
// ISO/IEC 18181-1 Annex C entropy coding) [UPSTREAM, recalled]. parity unpinned.
// The clustering / normalisation / Huffman heuristics are this repo's own deterministic
// integer formulations (DESIGN.md "Entropy stage"); only the emitted FORMAT follows libjxl.
#pragma once
#include "jxo.h"
#include "jxo_bits.h"

namespace jxo {

constexpr int kAnsLogTabSize = 12;
constexpr int kAnsTabSize = 1 << kAnsLogTabSize;
constexpr uint32_t kAnsSignature = 0x13;  // initial state 0x130000
constexpr int kLogAlphaSize = 8;          // alias tables of 256 buckets x 16 slots
constexpr int kAcAlphabet = 64;           // hybrid-uint (4,2,0) tokens of values < 2^16
constexpr int kModAlphabet = 128;         // hybrid-uint (4,2,0) tokens of values < 2^32
constexpr int kMaxClusters = 24;         // all reverse maps of one frame fit one SM's shared memory (24 x 8 KB)

// block contexts (libjxl ac_context.h)
constexpr int kNumOrders = 13;
constexpr int kNonZeroBuckets = 37;
constexpr int kZeroDensityContextCount = 458;
constexpr int kNumBlockCtx = 15;  // default BlockCtxMap
constexpr int kNumAcContexts = kNumBlockCtx * (kNonZeroBuckets + kZeroDensityContextCount);  // 7425
extern const uint8_t kDefaultBlockCtxMap[39];
extern const uint16_t kCoeffFreqContext[64];
extern const uint16_t kCoeffNumNonzeroContext[64];

// hybrid unsigned integer, config split_exponent=4 msb_in_token=2 lsb_in_token=0
static inline void HybridEncode(uint32_t v, uint32_t* tok, uint32_t* nbits, uint32_t* bits) {
  if (v < 16) { *tok = v; *nbits = 0; *bits = 0; return; }
  const uint32_t n = (uint32_t)FloorLog2(v);
  const uint32_t m = v - (1u << n);
  *tok = 16 + ((n - 4) << 2) + (m >> (n - 2));
  *nbits = n - 2;
  *bits = v & ((1u << (n - 2)) - 1);
}
static inline uint32_t HybridNbits(uint32_t tok) { return tok < 16 ? 0 : ((tok - 16) >> 2) + 4 - 2; }
static inline uint32_t HybridDecode(uint32_t tok, uint32_t bits) {
  if (tok < 16) return tok;
  const uint32_t n = ((tok - 16) >> 2) + 4;
  return (1u << n) | ((tok & 3) << (n - 2)) | bits;
}

// fixed-point log2 (Q20) used by the clustering cost; table built once from libm
int64_t Log2Q20(uint32_t n);                     // n >= 1
static inline int64_t XLogX(uint32_t n) { return n ? (int64_t)n * Log2Q20(n) : 0; }

// counts -> frequencies summing to 4096 (own deterministic rule)
void NormalizeCounts(const uint32_t* counts, int alphabet, uint16_t* norm);
// ANS histogram header (libjxl enc_ans.cc EncodeCounts / dec_ans.cc ReadHistogram), shift = 13
void WriteAnsHistogram(const uint16_t* norm, int alphabet, BitWriter* w);
bool ReadAnsHistogram(BitReader* r, std::vector<int>* counts);

// alias table (libjxl ans_common.cc InitAliasTable) — normative
struct AliasEntry { uint8_t cutoff; uint8_t right_value; uint16_t freq0; uint16_t offsets1; uint16_t freq1_xor_freq0; };
void InitAliasTable(std::vector<int> distribution, int log_alpha_size, AliasEntry* a);
struct AliasSymbol { uint32_t value, offset, freq; };
static inline AliasSymbol AliasLookup(const AliasEntry* a, uint32_t v, int log_entry_size) {
  const uint32_t i = v >> log_entry_size, pos = v & ((1u << log_entry_size) - 1);
  const bool greater = pos >= a[i].cutoff;
  AliasSymbol s;
  s.value = greater ? a[i].right_value : i;
  s.offset = (greater ? a[i].offsets1 : 0) + pos;
  s.freq = greater ? (a[i].freq0 ^ a[i].freq1_xor_freq0) : a[i].freq0;
  return s;
}

// one ANS-coded entropy code: cluster histograms + per-symbol reverse maps
struct AnsCode {
  int num_clusters = 0;
  int alphabet = 0;
  std::vector<uint16_t> norm;     // [cluster][alphabet]
  std::vector<uint16_t> rmap;     // [cluster][4096]: (symbol, offset) -> slot via sym_base
  std::vector<uint16_t> sym_base; // [cluster][alphabet]: start of the symbol's offsets in rmap
  void Build();                   // fills rmap / sym_base from norm
};

// writes tokens (ctx << 16 | value) through `cmap` with `code`: 32-bit final state, then the
// chunks in forward order (libjxl enc_ans.cc WriteTokens)
void AnsWriteTokens(const uint32_t* tokens, size_t n, const uint8_t* cmap, const AnsCode& code, BitWriter* w);

// clustering of per-context histograms (own integer formulation of libjxl's FastClusterHistograms)
// hist: [num_ctx][alphabet]; returns the number of clusters; cmap[num_ctx]; cluster_hist [k][alphabet]
int ClusterHistograms(const uint32_t* hist, int num_ctx, int alphabet, int max_clusters, uint8_t* cmap,
                      std::vector<uint32_t>* cluster_hist);

// context map coding (libjxl enc_context_map.cc / dec_context_map.cc), no MTF
void WriteContextMap(const uint8_t* cmap, int n, int num_clusters, BitWriter* w);

// prefix (Huffman) code, length-limited to 15 (format: Brotli-style, libjxl dec_huffman.cc)
struct PrefixCode {
  int alphabet = 0;                 // coded alphabet size (max used symbol + 1, >= 1)
  std::vector<uint8_t> length;      // [kModAlphabet]
  std::vector<uint16_t> bits;       // [kModAlphabet] code bits, already bit-reversed for LSB-first writing
};
void BuildPrefixCode(const uint32_t* counts, int alphabet, PrefixCode* pc);
void WritePrefixCodeHeader(const PrefixCode& pc, BitWriter* w);   // the per-histogram Huffman description

// ------------------------------------------------------------------ modular streams (DC, AC metadata)
constexpr int kNumModularCtx = 8;   // leaves of the fixed global MA tree
enum ModLeaf { kLeafDcY = 0, kLeafEpf = 1, kLeafYtoB = 2, kLeafYtoX = 3, kLeafDcB = 4, kLeafDcX = 5, kLeafQf = 6, kLeafAcs = 7 };
// tokens of one DC group: DC stream (Y, X, B) and AC-metadata stream; each token = leaf << 24 | packed residual
void ModularTokensDcGroup(const struct Frame& f, int dg, std::vector<uint32_t>* dc_tokens, std::vector<uint32_t>* meta_tokens,
                          uint32_t* num_first_blocks);
void WriteGlobalTree(int num_dc_groups, BitWriter* w);  // MA tree + its own entropy code

}  // namespace jxo
