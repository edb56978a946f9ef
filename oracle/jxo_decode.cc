// ORACLE (test infrastructure) — self-decoder for the codestream subset this repo emits
// (no djxl exists offline, SURVEY.md section 8c: decodability tier T2).  Written against the
// FORMAT (ISO/IEC 18181-1 as implemented by libjxl dec_frame.cc, dec_group.cc, dec_ans.cc,
// dec_huffman.cc, dec_context_map.cc, dec_modular.cc, modular/encoding/dec_ma.cc) [UPSTREAM,
// recalled], independently of the encoder-side code: it parses headers, TOC, entropy codes,
// the MA tree and every token stream generically and rejects anything outside the subset.
#include "jxo_entropy.h"
#include "jxo_frame.h"
#include "jxo_stages.h"

namespace jxo {

namespace {

struct UintConfig { int split_exponent = 4, msb = 2, lsb = 0; };

struct Huffman {
  int alphabet = 1;
  int single_symbol = 0;
  std::vector<uint8_t> len;
  // canonical decode tables
  uint32_t first_code[17] = {0}, count[17] = {0}, offset[17] = {0};
  std::vector<uint16_t> sorted;
  void Build() {
    for (int l = 0; l < 17; ++l) count[l] = 0;
    for (int s = 0; s < (int)len.size(); ++s) if (len[s]) count[len[s]]++;
    uint32_t code = 0, off = 0;
    for (int l = 1; l <= 15; ++l) { first_code[l] = code; offset[l] = off; code = (code + count[l]) << 1; off += count[l]; }
    sorted.assign(off, 0);
    uint32_t pos[17];
    for (int l = 0; l < 17; ++l) pos[l] = offset[l];
    for (int s = 0; s < (int)len.size(); ++s) if (len[s]) sorted[pos[len[s]]++] = (uint16_t)s;
  }
  int Read(BitReader* r) const {
    if (alphabet <= 1) return single_symbol;
    uint32_t code = 0;
    for (int l = 1; l <= 15; ++l) {
      code = (code << 1) | (uint32_t)r->Read(1);
      if (count[l] && code >= first_code[l] && code - first_code[l] < count[l]) return sorted[offset[l] + code - first_code[l]];
    }
    return -1;
  }
};

struct EntropyCode {
  bool use_prefix = false;
  int log_alpha_size = 8;
  std::vector<uint8_t> context_map;
  std::vector<UintConfig> cfg;
  std::vector<std::vector<AliasEntry>> alias;
  std::vector<Huffman> huff;
};

struct SymbolReader {
  const EntropyCode* code;
  BitReader* r;
  uint32_t state;
  SymbolReader(const EntropyCode* c, BitReader* br) : code(c), r(br) {
    state = c->use_prefix ? (kAnsSignature << 16) : (uint32_t)br->Read(32);
  }
  int ReadSymbol(int histo) {
    if (code->use_prefix) return code->huff[histo].Read(r);
    const AliasSymbol s = AliasLookup(code->alias[histo].data(), state & (kAnsTabSize - 1), kAnsLogTabSize - code->log_alpha_size);
    state = s.freq * (state >> kAnsLogTabSize) + s.offset;
    if (state < (1u << 16)) state = (state << 16) | (uint32_t)r->Read(16);
    return (int)s.value;
  }
  uint32_t ReadHybridUint(int ctx) {
    const int histo = code->context_map[ctx];
    const int tok = ReadSymbol(histo);
    if (tok < 0) return 0xFFFFFFFFu;
    const UintConfig& c = code->cfg[histo];
    const uint32_t split = 1u << c.split_exponent;
    if ((uint32_t)tok < split) return (uint32_t)tok;
    const uint32_t nbits = c.split_exponent - (c.msb + c.lsb) + (((uint32_t)tok - split) >> (c.msb + c.lsb));
    const uint32_t low = (uint32_t)tok & ((1u << c.lsb) - 1);
    const uint32_t t = (uint32_t)tok >> c.lsb;
    const uint32_t bits = (uint32_t)r->Read((int)nbits);
    return (((((1u << c.msb) | (t & ((1u << c.msb) - 1))) << nbits) | bits) << c.lsb) | low;
  }
  bool CheckFinalState() const { return state == (kAnsSignature << 16); }
};

bool ReadUintConfig(BitReader* r, int log_alpha_size, UintConfig* c) {
  c->split_exponent = (int)r->Read(CeilLog2((uint32_t)log_alpha_size + 1));
  c->msb = c->lsb = 0;
  if (c->split_exponent != log_alpha_size) {
    c->msb = (int)r->Read(CeilLog2((uint32_t)c->split_exponent + 1));
    if (c->msb > c->split_exponent) return false;
    c->lsb = (int)r->Read(CeilLog2((uint32_t)(c->split_exponent - c->msb) + 1));
  }
  return c->lsb + c->msb <= c->split_exponent;
}

bool ReadHuffman(BitReader* r, int alphabet, Huffman* h) {
  h->alphabet = alphabet;
  h->len.assign(alphabet, 0);
  if (alphabet <= 1) return true;
  const int hskip = (int)r->Read(2);
  if (hskip == 1) {  // simple code
    int max_bits = 0;
    for (int m = alphabet - 1; m; m >>= 1) ++max_bits;
    const int num = (int)r->Read(2) + 1;
    int sym[4];
    for (int i = 0; i < num; ++i) { sym[i] = (int)r->Read(max_bits); if (sym[i] >= alphabet) return false; }
    for (int i = 0; i < num; ++i) for (int j = i + 1; j < num; ++j) if (sym[i] == sym[j]) return false;
    if (num == 1) { h->alphabet = 1; h->single_symbol = sym[0]; return true; }
    if (num == 2) { h->len[sym[0]] = 1; h->len[sym[1]] = 1; }
    else if (num == 3) { h->len[sym[0]] = 1; h->len[sym[1]] = 2; h->len[sym[2]] = 2; }
    else if (r->Read(1)) { h->len[sym[0]] = 1; h->len[sym[1]] = 2; h->len[sym[2]] = 3; h->len[sym[3]] = 3; }
    else { for (int i = 0; i < 4; ++i) h->len[sym[i]] = 2; }
    h->Build();
    return true;
  }
  static const int kOrder[18] = {1, 2, 3, 4, 0, 5, 17, 6, 16, 7, 8, 9, 10, 11, 12, 13, 14, 15};
  static const uint8_t kLen[16] = {2, 2, 2, 3, 2, 2, 2, 4, 2, 2, 2, 3, 2, 2, 2, 4};
  static const uint8_t kVal[16] = {0, 4, 3, 2, 0, 4, 3, 1, 0, 4, 3, 2, 0, 4, 3, 5};
  Huffman clc;
  clc.alphabet = 18; clc.len.assign(18, 0);
  int space = 32, num_codes = 0;
  for (int i = hskip; i < 18 && space > 0; ++i) {
    const int idx = (int)r->Peek(4);
    r->Skip(kLen[idx]);
    const int v = kVal[idx];
    clc.len[kOrder[i]] = (uint8_t)v;
    if (v) { space -= 32 >> v; ++num_codes; }
  }
  if (!(num_codes == 1 || space == 0)) return false;
  clc.Build();
  int single = -1;
  if (num_codes == 1) for (int i = 0; i < 18; ++i) if (clc.len[i]) single = i;
  int symbol = 0, prev_len = 8, repeat = 0, repeat_len = 0;
  int sp = 32768;
  while (symbol < alphabet && sp > 0) {
    const int cl = single >= 0 ? single : clc.Read(r);
    if (cl < 0) return false;
    if (cl < 16) {
      repeat = 0;
      h->len[symbol++] = (uint8_t)cl;
      if (cl) { prev_len = cl; sp -= 32768 >> cl; }
    } else {
      const int extra = cl - 14;
      const int new_len = cl == 16 ? prev_len : 0;
      if (repeat_len != new_len) { repeat = 0; repeat_len = new_len; }
      const int old = repeat;
      if (repeat > 0) { repeat -= 2; repeat <<= extra; }
      repeat += (int)r->Read(extra) + 3;
      const int delta = repeat - old;
      if (symbol + delta > alphabet) return false;
      for (int i = 0; i < delta; ++i) h->len[symbol++] = (uint8_t)repeat_len;
      if (repeat_len) sp -= delta << (15 - repeat_len);
    }
  }
  if (sp != 0) return false;
  h->Build();
  return true;
}

bool ReadEntropyCode(BitReader* r, int num_contexts, EntropyCode* code);

bool ReadContextMap(BitReader* r, std::vector<uint8_t>* cmap, int* num_histograms) {
  const bool simple = r->Read(1);
  if (simple) {
    const int bits = (int)r->Read(2);
    for (auto& e : *cmap) e = bits ? (uint8_t)r->Read(bits) : 0;
  } else {
    const bool mtf = r->Read(1);
    EntropyCode nested;
    if (!ReadEntropyCode(r, 1, &nested)) return false;
    SymbolReader sr(&nested, r);
    for (auto& e : *cmap) { const uint32_t v = sr.ReadHybridUint(0); if (v > 255) return false; e = (uint8_t)v; }
    if (!sr.CheckFinalState()) return false;
    if (mtf) {
      uint8_t list[256];
      for (int i = 0; i < 256; ++i) list[i] = (uint8_t)i;
      for (auto& e : *cmap) {
        const uint8_t idx = e, v = list[idx];
        e = v;
        for (int i = idx; i > 0; --i) list[i] = list[i - 1];
        list[0] = v;
      }
    }
  }
  int mx = 0;
  for (auto e : *cmap) mx = std::max<int>(mx, e);
  *num_histograms = mx + 1;
  return true;
}

bool ReadEntropyCode(BitReader* r, int num_contexts, EntropyCode* code) {
  if (r->Read(1)) return false;  // lz77 is outside the emitted subset
  int num_histograms = 1;
  code->context_map.assign(num_contexts, 0);
  if (num_contexts > 1 && !ReadContextMap(r, &code->context_map, &num_histograms)) return false;
  code->use_prefix = r->Read(1);
  code->log_alpha_size = code->use_prefix ? 15 : (int)r->Read(2) + 5;
  code->cfg.resize(num_histograms);
  for (auto& c : code->cfg) if (!ReadUintConfig(r, code->log_alpha_size, &c)) return false;
  if (code->use_prefix) {
    std::vector<int> sizes(num_histograms);
    for (auto& s : sizes) s = (int)r->ReadVarLenUint16() + 1;
    code->huff.resize(num_histograms);
    for (int i = 0; i < num_histograms; ++i) if (!ReadHuffman(r, sizes[i], &code->huff[i])) return false;
  } else {
    code->alias.resize(num_histograms);
    for (int i = 0; i < num_histograms; ++i) {
      std::vector<int> counts;
      if (!ReadAnsHistogram(r, &counts)) return false;
      if ((int)counts.size() > (1 << code->log_alpha_size)) return false;
      code->alias[i].resize((size_t)1 << code->log_alpha_size);
      InitAliasTable(counts, code->log_alpha_size, code->alias[i].data());
    }
  }
  return !r->Overrun();
}

// ------------------------------------------------------------------ modular
struct TreeNode { int property; int splitval; int lchild, rchild; int predictor; int64_t offset; uint32_t multiplier; int leaf_id; };

bool ReadTree(BitReader* r, std::vector<TreeNode>* tree) {
  EntropyCode code;
  if (!ReadEntropyCode(r, 6, &code)) return false;
  SymbolReader sr(&code, r);
  int leaf_id = 0;
  size_t to_decode = 1;
  tree->clear();
  while (to_decode > 0) {
    if (tree->size() > (1u << 20)) return false;
    to_decode--;
    const int property = (int)sr.ReadHybridUint(1) - 1;
    TreeNode n{};
    if (property == -1) {
      n.property = -1;
      n.predictor = (int)sr.ReadHybridUint(2);
      n.offset = UnpackSigned(sr.ReadHybridUint(3));
      const uint32_t mul_log = sr.ReadHybridUint(4);
      const uint32_t mul_bits = sr.ReadHybridUint(5);
      n.multiplier = (mul_bits + 1u) << mul_log;
      n.leaf_id = leaf_id++;
      tree->push_back(n);
      continue;
    }
    n.property = property;
    n.splitval = UnpackSigned(sr.ReadHybridUint(0));
    n.lchild = (int)(tree->size() + to_decode + 1);
    n.rchild = (int)(tree->size() + to_decode + 2);
    tree->push_back(n);
    to_decode += 2;
  }
  return sr.CheckFinalState() && !r->Overrun();
}

struct Channel { int w = 0, h = 0; std::vector<int32_t> px; };

// GroupHeader + pixel data of one modular stream coded with the global tree / code
bool ReadModularStream(BitReader* r, const std::vector<TreeNode>& tree, const EntropyCode& code, int stream_id,
                       std::vector<Channel>* chans) {
  if (!r->Read(1)) return false;          // use_global_tree
  if (!r->Read(1)) return false;          // wp_header.all_default
  if (r->Read(2) != 0) return false;      // nb_transforms = 0
  SymbolReader sr(&code, r);
  for (size_t ci = 0; ci < chans->size(); ++ci) {
    Channel& ch = (*chans)[ci];
    ch.px.assign((size_t)ch.w * ch.h, 0);
    for (int y = 0; y < ch.h; ++y) for (int x = 0; x < ch.w; ++x) {
      auto at = [&](int xx, int yy) -> int32_t { return ch.px[(size_t)yy * ch.w + xx]; };
      const int32_t W = x ? at(x - 1, y) : (y ? at(x, y - 1) : 0);
      const int32_t N = y ? at(x, y - 1) : W;
      const int32_t NW = (x && y) ? at(x - 1, y - 1) : W;
      int pos = 0;
      while (tree[pos].property >= 0) {
        int64_t pv;
        switch (tree[pos].property) {
          case 0: pv = (int64_t)ci; break;
          case 1: pv = stream_id; break;
          case 2: pv = y; break;
          case 3: pv = x; break;
          case 4: pv = N < 0 ? -(int64_t)N : N; break;
          case 5: pv = W < 0 ? -(int64_t)W : W; break;
          case 6: pv = N; break;
          case 7: pv = W; break;
          default: return false;  // outside the subset
        }
        pos = pv > tree[pos].splitval ? tree[pos].lchild : tree[pos].rchild;
      }
      int64_t pred;
      switch (tree[pos].predictor) {
        case 0: pred = 0; break;
        case 1: pred = W; break;
        case 2: pred = N; break;
        case 5: { const int64_t m = std::min(W, N), M = std::max(W, N), g = (int64_t)W + N - NW; pred = g < m ? m : (g > M ? M : g); break; }
        default: return false;
      }
      const uint32_t v = sr.ReadHybridUint(tree[pos].leaf_id);
      ch.px[(size_t)y * ch.w + x] = (int32_t)((int64_t)UnpackSigned(v) * tree[pos].multiplier + tree[pos].offset + pred);
    }
  }
  return sr.CheckFinalState() && !r->Overrun();
}

uint32_t ReadU32(BitReader* r, const int bits[4], const uint32_t offs[4]) {
  const int sel = (int)r->Read(2);
  return (uint32_t)r->Read(bits[sel]) + offs[sel];
}

}  // namespace

// Decodes `data` into frame-level integers laid out exactly like the encoder's stages.
bool DecodeCodestream(const uint8_t* data, size_t size, Frame* f) {
  BitReader r(data, size);
  auto fail = [&](const char* m) { f->error = m; return false; };
  if (r.Read(8) != 0xFF || r.Read(8) != 0x0A) return fail("signature");
  // SizeHeader
  int xs, ys;
  {
    const int kb[4] = {9, 13, 18, 30}; const uint32_t ko[4] = {1, 1, 1, 1};
    const bool small = r.Read(1);
    ys = small ? ((int)r.Read(5) + 1) * 8 : (int)ReadU32(&r, kb, ko);
    const int ratio = (int)r.Read(3);
    if (ratio == 0) xs = small ? ((int)r.Read(5) + 1) * 8 : (int)ReadU32(&r, kb, ko);
    else {
      static const int num[8] = {0, 1, 12, 4, 3, 16, 5, 2}, den[8] = {1, 1, 10, 3, 2, 9, 4, 1};
      xs = (int)((int64_t)ys * num[ratio] / den[ratio]);
    }
  }
  if (!r.Read(1)) return fail("ImageMetadata not all_default");
  if (!r.Read(1)) return fail("custom transform data");
  r.ZeroPadToByte();
  f->fd.Set(xs, ys);
  const FrameDim& fd = f->fd;
  const size_t nblk = (size_t)fd.bxs * fd.bys;
  // FrameHeader
  if (r.Read(1)) return fail("all_default frame header (gaborish/EPF on) is outside the subset");
  if (r.Read(2) != 0) return fail("frame_type");
  if (r.Read(1) != 0) return fail("not VarDCT");
  const uint64_t flags = r.ReadU64();
  if (flags & ~(uint64_t)128) return fail("frame flags");
  if (r.Read(2) != 0) return fail("upsampling");
  f->q.x_qm_scale = (int)r.Read(3);
  f->q.b_qm_scale = (int)r.Read(3);
  if (r.Read(2) != 0) return fail("passes");
  if (r.Read(1)) return fail("custom size");
  if (r.Read(2) != 0) return fail("blend mode");
  if (!r.Read(1)) return fail("not last frame");
  if (r.Read(2) != 0) return fail("name");
  if (r.Read(1)) return fail("default loop filter");
  f->gab = r.Read(1) != 0;
  if (f->gab && r.Read(1)) return fail("custom gaborish weights");
  if (r.Read(2) != 0) return fail("epf");
  if (r.ReadU64() != 0) return fail("lf extensions");
  if (r.ReadU64() != 0) return fail("extensions");
  // TOC
  if (r.Read(1)) return fail("permuted TOC");
  r.ZeroPadToByte();
  const bool small = fd.num_groups == 1;
  const int nsec = small ? 1 : 2 + fd.num_dc_groups + fd.num_groups;
  std::vector<uint32_t> sec_size(nsec);
  {
    const int kb[4] = {10, 14, 22, 30}; const uint32_t ko[4] = {0, 1024, 17408, 4211712};
    for (auto& s : sec_size) s = ReadU32(&r, kb, ko);
  }
  r.ZeroPadToByte();
  std::vector<size_t> sec_start(nsec + 1);
  sec_start[0] = r.Pos() / 8;
  for (int i = 0; i < nsec; ++i) sec_start[i + 1] = sec_start[i] + sec_size[i];
  if (sec_start[nsec] != size) return fail("TOC does not add up to the codestream size");
  auto section = [&](int idx) { if (!small) r.Seek(sec_start[idx] * 8); };
  auto end_section = [&](int idx) -> bool {
    if (small) return true;
    return (r.Pos() + 7) / 8 == sec_start[idx + 1];   // consumed exactly (up to the zero padding)
  };
  // ---- LfGlobal
  section(0);
  if (!r.Read(1)) return fail("dc dequant");
  {
    const int kb[4] = {11, 11, 12, 16}; const uint32_t ko[4] = {1, 2049, 4097, 8193};
    f->q.global_scale = (int)ReadU32(&r, kb, ko);
    const int kb2[4] = {0, 5, 8, 16}; const uint32_t ko2[4] = {16, 1, 1, 1};
    f->q.quant_dc = (int)ReadU32(&r, kb2, ko2);
  }
  if (!r.Read(1)) return fail("block ctx map");
  if (!r.Read(1)) return fail("cmap dc");
  if (!r.Read(1)) return fail("no global tree");
  std::vector<TreeNode> tree;
  if (!ReadTree(&r, &tree)) return fail("MA tree");
  EntropyCode mcode;
  if (!ReadEntropyCode(&r, (int)(tree.size() + 1) / 2, &mcode)) return fail("modular entropy code");
  if (!end_section(0)) return fail("LfGlobal size");
  // ---- LfGroups
  f->dc_quant.assign(3 * nblk, 0);
  f->acs.assign(nblk, 0);
  f->raw_qf.assign(nblk, 0);
  f->cmap.assign((size_t)2 * fd.txs * fd.tys, 0);
  for (int dg = 0; dg < fd.num_dc_groups; ++dg) {
    section(1 + dg);
    const int x0 = (dg % fd.dgxs) * 256, y0 = (dg / fd.dgxs) * 256;
    const int w = std::min(256, fd.bxs - x0), h = std::min(256, fd.bys - y0);
    if (r.Read(2) != 0) return fail("extra precision");
    std::vector<Channel> dc(3);
    for (auto& c : dc) { c.w = w; c.h = h; }
    if (!ReadModularStream(&r, tree, mcode, 1 + dg, &dc)) return fail("DC stream");
    static const int plane_of_chan[3] = {1, 0, 2};
    for (int ch = 0; ch < 3; ++ch) for (int y = 0; y < h; ++y) for (int x = 0; x < w; ++x) {
      const int32_t v = dc[ch].px[(size_t)y * w + x];
      if (v < -32768 || v > 32767) return fail("dc range");
      f->dc_quant[(size_t)plane_of_chan[ch] * nblk + (size_t)(y0 + y) * fd.bxs + x0 + x] = (int16_t)v;
    }
    const int count = (int)r.Read(CeilLog2((uint32_t)(w * h))) + 1;
    std::vector<Channel> meta(4);
    meta[0].w = meta[1].w = (w + 7) >> 3; meta[0].h = meta[1].h = (h + 7) >> 3;
    meta[2].w = count; meta[2].h = 2;
    meta[3].w = w; meta[3].h = h;
    if (!ReadModularStream(&r, tree, mcode, 1 + 2 * fd.num_dc_groups + dg, &meta)) return fail("AC metadata stream");
    for (int m = 0; m < 2; ++m) for (int y = 0; y < meta[m].h; ++y) for (int x = 0; x < meta[m].w; ++x) {
      const int32_t v = meta[m].px[(size_t)y * meta[m].w + x];
      if (v < -128 || v > 127) return fail("cmap range");
      f->cmap[(size_t)m * fd.txs * fd.tys + (size_t)((y0 >> 3) + y) * fd.txs + (x0 >> 3) + x] = (int8_t)v;
    }
    int num = 0;
    for (int y = 0; y < h; ++y) for (int x = 0; x < w; ++x) {
      const size_t i = (size_t)(y0 + y) * fd.bxs + x0 + x;
      if (f->acs[i] & 0x40) continue;  // already covered by an earlier transform
      if (num >= count) return fail("too few strategies");
      const int s = meta[2].px[num];
      if (s < 0 || s >= 27) return fail("strategy value");
      const int cx = kCoveredX[s], cy = kCoveredY[s];
      if (x + cx > w || y + cy > h) return fail("transform crosses the DC group");
      if ((x % 32) + cx > 32 || (y % 32) + cy > 32) return fail("transform crosses an AC group");
      const int32_t q = 1 + std::max(0, std::min(255, meta[2].px[(size_t)count + num]));
      for (int iy = 0; iy < cy; ++iy) for (int ix = 0; ix < cx; ++ix) {
        const size_t j = i + (size_t)iy * fd.bxs + ix;
        if (f->acs[j] & 0x40) return fail("overlapping transforms");
        f->acs[j] = (uint8_t)(s | 0x40 | ((ix == 0 && iy == 0) ? 0x80 : 0));
        f->raw_qf[j] = q;
      }
      ++num;
    }
    for (const int32_t e : meta[3].px) if (e < 0 || e > 7) return fail("epf sharpness");
    if (num != count) return fail("strategy count");
    if (!end_section(1 + dg)) return fail("LfGroup size");
  }
  for (auto& a : f->acs) a = (uint8_t)(a & ~0x40);
  // ---- HfGlobal
  section(1 + fd.num_dc_groups);
  if (!r.Read(1)) return fail("dequant matrices");
  if (r.Read(CeilLog2((uint32_t)fd.num_groups)) != 0) return fail("num_histograms");
  if (r.Read(2) != 2) return fail("coefficient orders");
  EntropyCode ac;
  if (!ReadEntropyCode(&r, kNumAcContexts, &ac)) return fail("AC entropy code");
  f->num_clusters = (int)ac.cfg.size();
  f->context_map = ac.context_map;
  if (!end_section(1 + fd.num_dc_groups)) return fail("HfGlobal size");
  // ---- PassGroups
  const EncTables& T = GetTables();
  f->coeffs.assign((size_t)fd.num_groups * 1024 * 3 * 64, 0);
  f->nzeros.assign(3 * nblk, 0);
  f->nz_count.assign(3 * nblk, 0);
  static const int chan_of_slot[3] = {1, 0, 2};
  for (int g = 0; g < fd.num_groups; ++g) {
    section(2 + fd.num_dc_groups + g);
    SymbolReader sr(&ac, &r);
    const int gx0 = (g % fd.gxs) * 32, gy0 = (g / fd.gxs) * 32;
    const int gx1 = std::min(gx0 + 32, fd.bxs), gy1 = std::min(gy0 + 32, fd.bys);
    for (int by = gy0; by < gy1; ++by) for (int bx = gx0; bx < gx1; ++bx) {
      const uint8_t a = f->acs[(size_t)by * fd.bxs + bx];
      if (!(a & 0x80)) continue;
      const int s = a & 0x7f, cx = kCoveredX[s], cy = kCoveredY[s], n = cx * cy, size = n * 64;
      const int log2n = FloorLog2((uint32_t)n), ord = kStrategyOrder[s];
      (void)T;
      for (int slot = 0; slot < 3; ++slot) {
        const int c = chan_of_slot[slot];
        const int block_ctx = kDefaultBlockCtxMap[(c < 2 ? c ^ 1 : 2) * kNumOrders + ord];
        uint8_t* nzp = &f->nzeros[(size_t)c * nblk];
        const int lx = bx - gx0, ly = by - gy0;
        int pred;
        if (lx == 0) pred = ly == 0 ? 32 : nzp[(size_t)(by - 1) * fd.bxs + bx];
        else if (ly == 0) pred = nzp[(size_t)by * fd.bxs + bx - 1];
        else pred = (nzp[(size_t)(by - 1) * fd.bxs + bx] + nzp[(size_t)by * fd.bxs + bx - 1] + 1) / 2;
        const int p = pred >= 64 ? 64 : pred;
        const int bucket = p < 8 ? p : 4 + p / 2;
        int nz = (int)sr.ReadHybridUint(bucket * kNumBlockCtx + block_ctx);
        if (nz > size - n) return fail("nzeros too large");
        f->nz_count[(size_t)c * nblk + (size_t)by * fd.bxs + bx] = (uint16_t)nz;
        const int shared = (nz + n - 1) >> log2n;
        for (int iy = 0; iy < cy; ++iy) for (int ix = 0; ix < cx; ++ix) nzp[(size_t)(by + iy) * fd.bxs + bx + ix] = (uint8_t)shared;
        const int histo_offset = kNumBlockCtx * kNonZeroBuckets + kZeroDensityContextCount * block_ctx;
        int prev = nz > size / 16 ? 0 : 1;
        for (int k = n; k < size && nz != 0; ++k) {
          const int nzl = (nz + n - 1) >> log2n;
          const int ctx = histo_offset + (kCoeffNumNonzeroContext[nzl] + kCoeffFreqContext[k >> log2n]) * 2 + prev;
          const int32_t coeff = UnpackSigned(sr.ReadHybridUint(ctx));
          const int j = k / 64;
          const int cbx = bx + (j % cx), cby = by + (j / cx);
          const size_t blk = (size_t)g * 1024 + (size_t)(cby % 32) * 32 + (cbx % 32);
          f->coeffs[(blk * 3 + slot) * 64 + (k % 64)] = (int16_t)coeff;
          prev = coeff != 0;
          nz -= prev;
        }
        if (nz != 0) return fail("non-zero count mismatch");
      }
    }
    if (!sr.CheckFinalState()) return fail("ANS final state");
    if (r.Overrun()) return fail("overrun");
    if (!end_section(2 + fd.num_dc_groups + g)) return fail("PassGroup size");
  }
  if (small && (r.Pos() + 7) / 8 != size) return fail("trailing data");
  return true;
}

}  // namespace jxo
