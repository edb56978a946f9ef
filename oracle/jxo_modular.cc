// ORACLE (test infrastructure) — modular sub-bitstreams of a VarDCT frame (stage U9 / N1):
// quantised DC per 2048x2048 DC group and the AC metadata (chroma-from-luma map, AC strategy,
// quant field, EPF sharpness), coded with one fixed global MA tree and prefix (Huffman) codes.
// Format side restates libjxl enc_modular.cc (AddVarDCTDC / AddACMetadata), modular/encoding
// (GroupHeader, MA tree, predictors), dec_huffman.cc [UPSTREAM, recalled]. parity unpinned.
// The tree, predictor choice and Huffman construction are this repo's own.
#include "jxo_entropy.h"
#include "jxo_frame.h"

namespace jxo {

static inline int32_t ClampedGradient(int32_t w, int32_t n, int32_t nw) {
  const int32_t m = std::min(w, n), M = std::max(w, n);
  const int32_t g = w + n - nw;
  return g < m ? m : (g > M ? M : g);
}

static inline uint32_t ModToken(int leaf, int32_t residual) { return ((uint32_t)leaf << 24) | PackSigned(residual); }

void ModularTokensDcGroup(const Frame& f, int dg, std::vector<uint32_t>* dc_tokens, std::vector<uint32_t>* meta_tokens,
                          uint32_t* num_first_blocks) {
  const FrameDim& fd = f.fd;
  const size_t nblk = (size_t)fd.bxs * fd.bys;
  const int x0 = (dg % fd.dgxs) * 256, y0 = (dg / fd.dgxs) * 256;
  const int w = std::min(256, fd.bxs - x0), h = std::min(256, fd.bys - y0);
  // ---- DC: modular channels Y, X, B (libjxl: channel[c < 2 ? c ^ 1 : c]); gradient predictor
  static const int plane_of_chan[3] = {1, 0, 2};
  static const int leaf_of_chan[3] = {kLeafDcY, kLeafDcX, kLeafDcB};
  dc_tokens->clear();
  for (int ch = 0; ch < 3; ++ch) {
    const int16_t* p = &f.dc_quant[(size_t)plane_of_chan[ch] * nblk];
    auto at = [&](int x, int y) -> int32_t { return p[(size_t)(y0 + y) * fd.bxs + x0 + x]; };
    for (int y = 0; y < h; ++y) for (int x = 0; x < w; ++x) {
      const int32_t W = x ? at(x - 1, y) : (y ? at(x, y - 1) : 0);
      const int32_t N = y ? at(x, y - 1) : W;
      const int32_t NW = (x && y) ? at(x - 1, y - 1) : W;
      dc_tokens->push_back(ModToken(leaf_of_chan[ch], at(x, y) - ClampedGradient(W, N, NW)));
    }
  }
  // ---- AC metadata
  meta_tokens->clear();
  const int tx0 = x0 >> 3, ty0 = y0 >> 3, tw = (w + 7) >> 3, th = (h + 7) >> 3;
  for (int m = 0; m < 2; ++m)  // channel 0 = ytox, 1 = ytob; zero predictor
    for (int y = 0; y < th; ++y) for (int x = 0; x < tw; ++x)
      meta_tokens->push_back(ModToken(m == 0 ? kLeafYtoX : kLeafYtoB,
                                      f.cmap[(size_t)m * fd.txs * fd.tys + (size_t)(ty0 + y) * fd.txs + tx0 + x]));
  // channel 2: row 0 = raw strategy of every first block (raster), row 1 = quant - 1
  std::vector<int32_t> strat, qf;
  for (int y = 0; y < h; ++y) for (int x = 0; x < w; ++x) {
    const size_t i = (size_t)(y0 + y) * fd.bxs + x0 + x;
    if (!(f.acs[i] & 0x80)) continue;
    strat.push_back(f.acs[i] & 0x7f);
    qf.push_back(f.raw_qf[i] - 1);
  }
  *num_first_blocks = (uint32_t)strat.size();
  for (size_t i = 0; i < strat.size(); ++i) meta_tokens->push_back(ModToken(kLeafAcs, strat[i]));          // zero predictor
  for (size_t i = 0; i < qf.size(); ++i) meta_tokens->push_back(ModToken(kLeafQf, qf[i] - (i ? qf[i - 1] : strat[0])));  // left predictor
  // channel 3: EPF sharpness, all zero (loop filter disabled); zero predictor
  for (int i = 0; i < w * h; ++i) meta_tokens->push_back(ModToken(kLeafEpf, 0));
}

// ------------------------------------------------------------------ fixed global MA tree
// BFS layout (libjxl dec_ma.cc DecodeTree): node = (property, splitval) or leaf (predictor).
// property 0 = channel, 1 = stream id, 2 = y.  "> splitval" goes to the left child.
// stream ids: DC group g -> 1 + g, AC metadata of g -> 1 + 2 * num_dc_groups + g.
void WriteGlobalTree(int num_dc_groups, BitWriter* w) {
  struct Node { int property; int value; };  // property -1: leaf, value = predictor
  const Node nodes[15] = {
      {1, 2 * num_dc_groups},  // 0: stream id > 2*ndc ? AC metadata : DC
      {0, 1},                  // 1: metadata: channel > 1 ?
      {0, 0},                  // 2: DC: channel > 0 ?
      {0, 2},                  // 3: metadata channel > 2 ? epf : acs/qf
      {0, 0},                  // 4: metadata channel > 0 ? ytob : ytox
      {0, 1},                  // 5: DC channel > 1 ? B : X
      {-1, 5},                 // 6: leaf 0 DC Y  (gradient)
      {-1, 0},                 // 7: leaf 1 EPF   (zero)
      {2, 0},                  // 8: y > 0 ? qf : strategy
      {-1, 0},                 // 9: leaf 2 ytob  (zero)
      {-1, 0},                 // 10: leaf 3 ytox (zero)
      {-1, 5},                 // 11: leaf 4 DC B (gradient)
      {-1, 5},                 // 12: leaf 5 DC X (gradient)
      {-1, 1},                 // 13: leaf 6 qf   (left)
      {-1, 0},                 // 14: leaf 7 strategy (zero)
  };
  std::vector<uint32_t> tokens;  // ctx << 16 | value; contexts: 0 splitval, 1 property, 2 predictor, 3 offset, 4 mul_log, 5 mul_bits
  for (const Node& n : nodes) {
    if (n.property < 0) {
      tokens.push_back((1u << 16) | 0);
      tokens.push_back((2u << 16) | (uint32_t)n.value);
      tokens.push_back((3u << 16) | 0);
      tokens.push_back((4u << 16) | 0);
      tokens.push_back((5u << 16) | 0);
    } else {
      tokens.push_back((1u << 16) | (uint32_t)(n.property + 1));
      tokens.push_back((0u << 16) | PackSigned(n.value));
    }
  }
  // entropy code of the tree: 6 contexts -> one ANS histogram
  w->Write(1, 0);                    // lz77 disabled
  w->Write(1, 1); w->Write(2, 0);    // context map: simple, 0 bits per entry
  w->Write(1, 0);                    // ANS
  w->Write(2, kLogAlphaSize - 5);
  w->Write(4, 4); w->Write(3, 2); w->Write(2, 0);
  std::vector<uint32_t> counts(kAcAlphabet, 0);
  for (uint32_t t : tokens) { uint32_t tok, nb, bits; HybridEncode(t & 0xFFFF, &tok, &nb, &bits); counts[tok]++; }
  AnsCode code;
  code.num_clusters = 1; code.alphabet = kAcAlphabet;
  code.norm.assign(kAcAlphabet, 0);
  NormalizeCounts(counts.data(), kAcAlphabet, code.norm.data());
  WriteAnsHistogram(code.norm.data(), kAcAlphabet, w);
  code.Build();
  const uint8_t cmap[6] = {0, 0, 0, 0, 0, 0};
  AnsWriteTokens(tokens.data(), tokens.size(), cmap, code, w);
}

// ------------------------------------------------------------------ prefix codes
// Deterministic Huffman: repeatedly merge the two smallest (weight, id) nodes; leaves have
// id = symbol, internal nodes id = 256 + creation order; counts are raised to a doubling
// floor until no code is longer than 15 bits (the retry rule of Brotli's encoder).
void BuildPrefixCode(const uint32_t* counts_in, int alphabet_max, PrefixCode* pc) {
  pc->length.assign(kModAlphabet, 0);
  pc->bits.assign(kModAlphabet, 0);
  uint32_t counts[kModAlphabet];
  int used = 0, last = -1;
  for (int s = 0; s < kModAlphabet; ++s) {
    counts[s] = s < alphabet_max ? counts_in[s] : 0;
    if (counts[s]) { ++used; last = s; }
  }
  if (used == 0 || (used == 1 && last == 0)) { pc->alphabet = 1; return; }
  if (used == 1) { counts[0] = 1; used = 2; }  // a complex prefix code needs two coded symbols
  pc->alphabet = last + 1;
  for (uint32_t floor_count = 1;; floor_count *= 2) {
    uint64_t weight[2 * kModAlphabet];
    int parent[2 * kModAlphabet], id[2 * kModAlphabet];
    bool alive[2 * kModAlphabet];
    int n = 0;
    int leaf_node[kModAlphabet];
    for (int s = 0; s <= last; ++s) {
      leaf_node[s] = -1;
      if (!counts[s]) continue;
      weight[n] = std::max(counts[s], floor_count); parent[n] = -1; id[n] = s; alive[n] = true;
      leaf_node[s] = n++;
    }
    int live = n, created = 0;
    while (live > 1) {
      int a = -1, b = -1;  // two smallest by (weight, id)
      for (int i = 0; i < n; ++i) {
        if (!alive[i]) continue;
        auto less = [&](int p, int q) { return weight[p] < weight[q] || (weight[p] == weight[q] && id[p] < id[q]); };
        if (a < 0 || less(i, a)) { b = a; a = i; }
        else if (b < 0 || less(i, b)) b = i;
      }
      weight[n] = weight[a] + weight[b]; parent[n] = -1; id[n] = 256 + created++; alive[n] = true;
      parent[a] = n; parent[b] = n; alive[a] = alive[b] = false;
      ++n; --live;
    }
    int maxlen = 0;
    for (int s = 0; s <= last; ++s) {
      if (leaf_node[s] < 0) { pc->length[s] = 0; continue; }
      int d = 0;
      for (int v = leaf_node[s]; parent[v] >= 0; v = parent[v]) ++d;
      pc->length[s] = (uint8_t)d;
      maxlen = std::max(maxlen, d);
    }
    if (maxlen <= 15) break;
  }
  // canonical codes: by increasing length, then symbol; written bit-reversed (LSB-first stream)
  uint32_t code = 0;
  for (int len = 1; len <= 15; ++len) {
    for (int s = 0; s <= last; ++s) {
      if (pc->length[s] != len) continue;
      uint32_t rev = 0;
      for (int i = 0; i < len; ++i) rev |= ((code >> i) & 1) << (len - 1 - i);
      pc->bits[s] = (uint16_t)rev;
      ++code;
    }
    code <<= 1;
  }
}

// Complex prefix-code description (Brotli RFC 7932 section 3.5 as used by libjxl dec_huffman.cc):
// no skipping, a flat 4-bit code-length code over lengths 0..15, every length literal, stopping
// at the last symbol (where the Kraft sum completes and the decoder stops reading).
void WritePrefixCodeHeader(const PrefixCode& pc, BitWriter* w) {
  if (pc.alphabet <= 1) return;
  w->Write(2, 0);  // hskip = 0
  // code-length-code lengths in the order {1,2,3,4,0,5,17,6,16,7,...,15}: 4 for lengths 0..15, 0 for 16 and 17
  static const int kOrder[18] = {1, 2, 3, 4, 0, 5, 17, 6, 16, 7, 8, 9, 10, 11, 12, 13, 14, 15};
  for (int i = 0; i < 18; ++i) {
    if (kOrder[i] >= 16) w->Write(2, 0);   // value 0 -> "00"
    else w->Write(2, 1);                   // value 4 -> bits 1,0
  }
  for (int s = 0; s < pc.alphabet; ++s) {
    const uint32_t v = pc.length[s];       // canonical 4-bit code of value v is v itself; bit-reverse for the stream
    const uint32_t rev = ((v & 1) << 3) | ((v & 2) << 1) | ((v & 4) >> 1) | ((v & 8) >> 3);
    w->Write(4, rev);
  }
}

}  // namespace jxo
