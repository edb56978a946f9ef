// ORACLE (test infrastructure) — stages U6-U8: coefficient tokenisation, per-context
// histograms, clustering, ANS histogram headers, alias tables and the reverse-order rANS
// writer.  Format side restates libjxl enc_entropy_coder.cc (TokenizeCoefficients),
// ac_context.h, enc_ans.cc (WriteTokens / EncodeCounts), ans_common.cc (InitAliasTable) and
// dec_ans.cc (ReadHistogram) [UPSTREAM, recalled; SURVEY.md Appendix U.10-12, U.17-18].
// parity unpinned.  Clustering / normalisation are own integer formulations.
#include "jxo_entropy.h"
#include "jxo_frame.h"
#include "jxo_stages.h"

namespace jxo {

const uint8_t kDefaultBlockCtxMap[39] = {0, 1, 2, 2, 3, 3, 4, 5, 6, 6, 6, 6, 6,
                                         7, 8, 9, 9, 10, 11, 12, 13, 14, 14, 14, 14, 14,
                                         7, 8, 9, 9, 10, 11, 12, 13, 14, 14, 14, 14, 14};
const uint16_t kCoeffFreqContext[64] = {
    0xBAD, 0,  1,  2,  3,  4,  5,  6,  7,  8,  9,  10, 11, 12, 13, 14, 15, 15, 16, 16, 17, 17,
    18,    18, 19, 19, 20, 20, 21, 21, 22, 22, 23, 23, 23, 23, 24, 24, 24, 24, 25, 25, 25, 25,
    26,    26, 26, 26, 27, 27, 27, 27, 28, 28, 28, 28, 29, 29, 29, 29, 30, 30, 30, 30};
const uint16_t kCoeffNumNonzeroContext[64] = {
    0xBAD, 0,   31,  62,  62,  93,  93,  93,  93,  123, 123, 123, 123, 152, 152, 152,
    152,   152, 152, 152, 152, 180, 180, 180, 180, 180, 180, 180, 180, 180, 180, 180,
    180,   206, 206, 206, 206, 206, 206, 206, 206, 206, 206, 206, 206, 206, 206, 206,
    206,   206, 206, 206, 206, 206, 206, 206, 206, 206, 206, 206, 206, 206, 206, 206};

// ------------------------------------------------------------------ fixed-point log2
static int32_t g_log2_lut[1025];
static bool InitLog2() {
  for (int i = 0; i <= 1024; ++i) g_log2_lut[i] = (int32_t)lrint(ldexp(log2(1.0 + (double)i / 1024.0), 20));
  return true;
}
int64_t Log2Q20(uint32_t n) {
  static const bool init = InitLog2();   // thread-safe one-time initialisation
  (void)init;
  const int e = FloorLog2(n);
  const uint32_t m = n << (31 - e);             // leading one at bit 31
  const uint32_t idx = (m >> 21) & 1023;        // next 10 bits
  const uint32_t frac = (m >> 5) & 0xFFFF;      // next 16 bits
  const int64_t a = g_log2_lut[idx], b = g_log2_lut[idx + 1];
  return ((int64_t)e << 20) + a + (((b - a) * (int64_t)frac) >> 16);
}

// ------------------------------------------------------------------ normalisation
void NormalizeCounts(const uint32_t* counts, int alphabet, uint16_t* norm) {
  uint64_t total = 0;
  for (int s = 0; s < alphabet; ++s) total += counts[s];
  for (int s = 0; s < alphabet; ++s) norm[s] = 0;
  if (total == 0) { norm[0] = kAnsTabSize; return; }
  int64_t sum = 0;
  for (int s = 0; s < alphabet; ++s) {
    if (!counts[s]) continue;
    uint64_t t = ((uint64_t)counts[s] * (2 * kAnsTabSize) + total) / (2 * total);
    if (t < 1) t = 1;
    norm[s] = (uint16_t)t;
    sum += (int64_t)t;
  }
  int64_t delta = kAnsTabSize - sum;
  while (delta != 0) {
    int best = 0;
    for (int s = 1; s < alphabet; ++s) if (norm[s] > norm[best]) best = s;  // largest, lowest index on ties
    if (delta > 0) { norm[best] = (uint16_t)(norm[best] + delta); delta = 0; }
    else {
      int64_t take = -delta < (int64_t)norm[best] - 1 ? -delta : (int64_t)norm[best] - 1;
      norm[best] = (uint16_t)(norm[best] - take);
      delta += take;
    }
  }
}

// ------------------------------------------------------------------ ANS histogram header
static const uint8_t kLogCountBitLengths[14] = {5, 4, 4, 4, 4, 4, 3, 3, 3, 3, 3, 6, 7, 7};
static const uint8_t kLogCountSymbols[14] = {17, 11, 15, 3, 9, 7, 4, 2, 5, 6, 0, 33, 1, 65};

void WriteAnsHistogram(const uint16_t* norm, int alphabet, BitWriter* w) {
  int nsym = 0, syms[2] = {0, 0}, last = -1;
  for (int s = 0; s < alphabet; ++s) if (norm[s]) { if (nsym < 2) syms[nsym] = s; ++nsym; last = s; }
  if (nsym <= 2) {
    w->Write(1, 1);                       // simple code
    w->Write(1, (uint64_t)(nsym == 2));   // num_symbols - 1
    if (nsym == 0) { w->WriteVarLenUint8(0); return; }
    w->WriteVarLenUint8((uint32_t)syms[0]);
    if (nsym == 2) { w->WriteVarLenUint8((uint32_t)syms[1]); w->Write(kAnsLogTabSize, norm[syms[0]]); }
    return;
  }
  w->Write(1, 0);  // not simple
  w->Write(1, 0);  // not flat
  w->Write(3, 7);  // shift: unary "111" -> log = 3
  w->Write(3, 6);  // (6 | 8) - 1 = 13 = full precision
  const int length = last + 1;
  w->WriteVarLenUint8((uint32_t)(length - 3));
  int logcounts[256], omit_pos = -1, omit_log = -1;
  for (int i = 0; i < length; ++i) {
    logcounts[i] = norm[i] ? FloorLog2(norm[i]) + 1 : 0;
    if (logcounts[i] > omit_log) { omit_log = logcounts[i]; omit_pos = i; }
  }
  for (int i = 0; i < length; ++i) w->Write(kLogCountBitLengths[logcounts[i]], kLogCountSymbols[logcounts[i]]);
  for (int i = 0; i < length; ++i) {
    if (i == omit_pos || logcounts[i] <= 1) continue;
    const int l = logcounts[i] - 1;  // population-count precision at shift 13 is the full l bits
    w->Write(l, (uint64_t)(norm[i] - (1u << l)));
  }
}

bool ReadAnsHistogram(BitReader* r, std::vector<int>* counts) {
  counts->clear();
  if (r->Read(1)) {  // simple
    const int num = (int)r->Read(1) + 1;
    int syms[2];
    for (int i = 0; i < num; ++i) syms[i] = (int)r->ReadVarLenUint8();
    const int mx = num == 2 ? std::max(syms[0], syms[1]) : syms[0];
    counts->assign(mx + 1, 0);
    if (num == 1) (*counts)[syms[0]] = kAnsTabSize;
    else {
      if (syms[0] == syms[1]) return false;
      (*counts)[syms[0]] = (int)r->Read(kAnsLogTabSize);
      (*counts)[syms[1]] = kAnsTabSize - (*counts)[syms[0]];
    }
    return true;
  }
  if (r->Read(1)) {  // flat
    const int n = (int)r->ReadVarLenUint8() + 1;
    counts->assign(n, kAnsTabSize / n);
    for (int i = 0; i < kAnsTabSize % n; ++i) (*counts)[i]++;
    return true;
  }
  int log = 0;
  for (; log < 3; ++log) if (!r->Read(1)) break;
  const int shift = (int)((r->Read(log) | (1u << log)) - 1);
  if (shift > kAnsLogTabSize + 1) return false;
  const int length = (int)r->ReadVarLenUint8() + 3;
  counts->assign(length, 0);
  std::vector<int> logcounts(length, 0), same(length, 0);
  int omit_log = -1, omit_pos = -1;
  // prefix code of the log-counts: decode by matching (length, symbol) pairs
  for (int i = 0; i < length; ++i) {
    int found = -1;
    for (int sym = 0; sym < 14 && found < 0; ++sym) {
      const int len = kLogCountBitLengths[sym];
      if ((int)r->Peek(len) == kLogCountSymbols[sym]) found = sym;
    }
    if (found < 0) return false;
    r->Skip(kLogCountBitLengths[found]);
    logcounts[i] = found;
    if (found == kAnsLogTabSize + 1) {
      const int rle = (int)r->ReadVarLenUint8();
      same[i] = rle + 5;
      i += rle + 3;
      continue;
    }
    if (found > omit_log) { omit_log = found; omit_pos = i; }
  }
  if (omit_pos < 0) return false;
  if (omit_pos + 1 < length && logcounts[omit_pos + 1] == kAnsLogTabSize + 1) return false;
  int prev = 0, numsame = 0, total = 0;
  for (int i = 0; i < length; ++i) {
    if (same[i]) { numsame = same[i] - 1; prev = i > 0 ? (*counts)[i - 1] : 0; }
    if (numsame > 0) { (*counts)[i] = prev; numsame--; }
    else {
      const int code = logcounts[i];
      if (i == omit_pos || code == 0) continue;
      if (code == 1) (*counts)[i] = 1;
      else {
        int bitcount = std::min(code - 1, shift - ((kAnsLogTabSize - (code - 1)) >> 1));
        if (bitcount < 0) bitcount = 0;
        (*counts)[i] = (1 << (code - 1)) + ((int)r->Read(bitcount) << (code - 1 - bitcount));
      }
    }
    total += (*counts)[i];
  }
  (*counts)[omit_pos] = kAnsTabSize - total;
  return (*counts)[omit_pos] > 0;
}

// ------------------------------------------------------------------ alias table
void InitAliasTable(std::vector<int> distribution, int log_alpha_size, AliasEntry* a) {
  const int range = kAnsTabSize;
  while (!distribution.empty() && distribution.back() == 0) distribution.pop_back();
  if (distribution.empty()) distribution.push_back(range);
  const int table_size = 1 << log_alpha_size;
  const int log_entry_size = kAnsLogTabSize - log_alpha_size;
  const int entry_size = 1 << log_entry_size;
  for (size_t sym = 0; sym < distribution.size(); ++sym) {
    if (distribution[sym] == range) {
      for (int i = 0; i < table_size; ++i) {
        a[i].right_value = (uint8_t)sym; a[i].cutoff = 0; a[i].offsets1 = (uint16_t)(entry_size * i);
        a[i].freq0 = 0; a[i].freq1_xor_freq0 = (uint16_t)range;
      }
      return;
    }
  }
  std::vector<uint32_t> underfull, overfull, cutoffs(table_size, 0), offsets1(table_size, 0), right(table_size, 0);
  for (size_t i = 0; i < distribution.size(); ++i) {
    cutoffs[i] = (uint32_t)distribution[i];
    if ((int)cutoffs[i] > entry_size) overfull.push_back((uint32_t)i);
    else if ((int)cutoffs[i] < entry_size) underfull.push_back((uint32_t)i);
  }
  for (int i = (int)distribution.size(); i < table_size; ++i) { cutoffs[i] = 0; underfull.push_back((uint32_t)i); }
  while (!overfull.empty()) {
    const uint32_t o = overfull.back(); overfull.pop_back();
    const uint32_t u = underfull.back(); underfull.pop_back();
    const uint32_t by = (uint32_t)entry_size - cutoffs[u];
    cutoffs[o] -= by;
    right[u] = o;
    offsets1[u] = cutoffs[o];
    if ((int)cutoffs[o] < entry_size) underfull.push_back(o);
    else if ((int)cutoffs[o] > entry_size) overfull.push_back(o);
  }
  for (int i = 0; i < table_size; ++i) {
    if ((int)cutoffs[i] == entry_size) { a[i].right_value = (uint8_t)i; a[i].offsets1 = 0; a[i].cutoff = 0; }
    else { a[i].right_value = (uint8_t)right[i]; a[i].offsets1 = (uint16_t)(offsets1[i] - cutoffs[i]); a[i].cutoff = (uint8_t)cutoffs[i]; }
    const int freq0 = i < (int)distribution.size() ? distribution[i] : 0;
    const int i1 = a[i].right_value;
    const int freq1 = i1 < (int)distribution.size() ? distribution[i1] : 0;
    a[i].freq0 = (uint16_t)freq0;
    a[i].freq1_xor_freq0 = (uint16_t)(freq1 ^ freq0);
  }
}

void AnsCode::Build() {
  rmap.assign((size_t)num_clusters * kAnsTabSize, 0);
  sym_base.assign((size_t)num_clusters * alphabet, 0);
  std::vector<AliasEntry> table(1 << kLogAlphaSize);
  for (int k = 0; k < num_clusters; ++k) {
    const uint16_t* nm = &norm[(size_t)k * alphabet];
    std::vector<int> dist(nm, nm + alphabet);
    InitAliasTable(dist, kLogAlphaSize, table.data());
    uint32_t acc = 0;
    for (int s = 0; s < alphabet; ++s) { sym_base[(size_t)k * alphabet + s] = (uint16_t)acc; acc += nm[s]; }
    for (uint32_t v = 0; v < (uint32_t)kAnsTabSize; ++v) {
      const AliasSymbol s = AliasLookup(table.data(), v, kAnsLogTabSize - kLogAlphaSize);
      rmap[(size_t)k * kAnsTabSize + sym_base[(size_t)k * alphabet + s.value] + s.offset] = (uint16_t)v;
    }
  }
}

// ------------------------------------------------------------------ rANS writer
void AnsWriteTokens(const uint32_t* tokens, size_t n, const uint8_t* cmap, const AnsCode& code, BitWriter* w) {
  // chunk list built back to front, written front to back
  std::vector<uint32_t> chunk_bits; std::vector<uint8_t> chunk_n;
  chunk_bits.reserve(2 * n); chunk_n.reserve(2 * n);
  uint32_t state = kAnsSignature << 16;
  for (size_t i = n; i-- > 0;) {
    const uint32_t ctx = tokens[i] >> 16, value = tokens[i] & 0xFFFF;
    const int k = cmap[ctx];
    uint32_t tok, nbits, bits;
    HybridEncode(value, &tok, &nbits, &bits);
    chunk_bits.push_back(bits); chunk_n.push_back((uint8_t)nbits);   // extra bits first: the list is reversed
    const uint32_t freq = code.norm[(size_t)k * code.alphabet + tok];
    if ((state >> (32 - kAnsLogTabSize)) >= freq) {
      chunk_bits.push_back(state & 0xFFFF); chunk_n.push_back(16);
      state >>= 16;
    } else { chunk_bits.push_back(0); chunk_n.push_back(0); }
    const uint32_t q = state / freq, r = state % freq;
    state = (q << kAnsLogTabSize) + code.rmap[(size_t)k * kAnsTabSize + code.sym_base[(size_t)k * code.alphabet + tok] + r];
  }
  w->Write(32, state);
  for (size_t i = chunk_bits.size(); i-- > 0;) if (chunk_n[i]) w->Write(chunk_n[i], chunk_bits[i]);
}

// ------------------------------------------------------------------ clustering
int ClusterHistograms(const uint32_t* hist, int num_ctx, int alphabet, int max_clusters, uint8_t* cmap,
                      std::vector<uint32_t>* cluster_hist) {
  std::vector<uint32_t> total(num_ctx, 0);
  std::vector<int64_t> dist(num_ctx, 0);
  std::vector<int> assign(num_ctx, 0);
  int first = -1;
  for (int c = 0; c < num_ctx; ++c) {
    uint32_t t = 0;
    for (int s = 0; s < alphabet; ++s) t += hist[(size_t)c * alphabet + s];
    total[c] = t;
    dist[c] = t ? INT64_MAX : -1;
    if (t && (first < 0 || t > total[first])) first = c;
  }
  int K = 0;
  if (first >= 0) {
    const int64_t kMinDist = (int64_t)64 << 20;
    int seed = first;
    while (true) {
      const int k = K++;
      const uint32_t* hb = &hist[(size_t)seed * alphabet];
      const uint32_t tb = total[seed];
      for (int c = 0; c < num_ctx; ++c) {
        if (!total[c]) continue;
        const uint32_t* ha = &hist[(size_t)c * alphabet];
        const uint32_t ta = total[c];
        int64_t d = XLogX(ta + tb) - XLogX(ta) - XLogX(tb);
        for (int s = 0; s < alphabet; ++s) {
          const uint32_t a = ha[s], b = hb[s];
          if (a && b) d -= XLogX(a + b) - XLogX(a) - XLogX(b);
        }
        if (c == seed) d = 0;
        if (d < dist[c]) { dist[c] = d; assign[c] = k; }
      }
      if (K >= max_clusters) break;
      int next = -1;
      for (int c = 0; c < num_ctx; ++c) if (total[c] && (next < 0 || dist[c] > dist[next])) next = c;
      if (dist[next] < kMinDist) break;
      seed = next;
    }
  } else {
    K = 1;
  }
  cluster_hist->assign((size_t)K * alphabet, 0);
  int prev = 0;
  for (int c = 0; c < num_ctx; ++c) {
    if (total[c]) {
      prev = assign[c];
      for (int s = 0; s < alphabet; ++s) (*cluster_hist)[(size_t)prev * alphabet + s] += hist[(size_t)c * alphabet + s];
    }
    cmap[c] = (uint8_t)prev;  // empty contexts inherit the previous context's cluster
  }
  return K;
}

// ------------------------------------------------------------------ context map
void WriteContextMap(const uint8_t* cmap, int n, int num_clusters, BitWriter* w) {
  if (num_clusters == 1) { w->Write(1, 1); w->Write(2, 0); return; }  // simple, 0 bits per entry
  w->Write(1, 0);  // not simple
  w->Write(1, 0);  // no move-to-front
  // nested single-context code over the hybrid-uint tokens of the entries: a PREFIX code, so that
  // the 7425 entries can be written in parallel (an rANS chain would serialise them)
  w->Write(1, 0);  // lz77 disabled
  w->Write(1, 1);  // prefix code
  w->Write(4, 4); w->Write(3, 2); w->Write(2, 0);  // uint config (4, 2, 0) at log_alpha_size 15
  std::vector<uint32_t> counts(kModAlphabet, 0);
  for (int i = 0; i < n; ++i) {
    uint32_t tok, nb, bits;
    HybridEncode(cmap[i], &tok, &nb, &bits);
    counts[tok]++;
  }
  PrefixCode pc;
  BuildPrefixCode(counts.data(), kModAlphabet, &pc);
  w->WriteVarLenUint16((uint32_t)(pc.alphabet - 1));
  WritePrefixCodeHeader(pc, w);
  for (int i = 0; i < n; ++i) {
    uint32_t tok, nb, bits;
    HybridEncode(cmap[i], &tok, &nb, &bits);
    w->Write(pc.length[tok], pc.bits[tok]);
    w->Write((int)nb, bits);
  }
}

// ------------------------------------------------------------------ coefficient tokens (U6)
static void TokenizeGroup(const Frame& f, int g, std::vector<uint32_t>* out) {
  const FrameDim& fd = f.fd;
  const size_t nblk = (size_t)fd.bxs * fd.bys;
  const int gx0 = (g % fd.gxs) * 32, gy0 = (g / fd.gxs) * 32;
  const int gx1 = std::min(gx0 + 32, fd.bxs), gy1 = std::min(gy0 + 32, fd.bys);
  static const int chan_of_slot[3] = {1, 0, 2};
  for (int by = gy0; by < gy1; ++by) for (int bx = gx0; bx < gx1; ++bx) {
    const uint8_t a = f.acs[(size_t)by * fd.bxs + bx];
    if (!(a & 0x80)) continue;
    const int s = a & 0x7f;
    const int cx = kCoveredX[s], cy = kCoveredY[s], n = cx * cy, size = n * 64;
    const int log2n = FloorLog2((uint32_t)n);
    const int ord = kStrategyOrder[s];
    for (int slot = 0; slot < 3; ++slot) {
      const int c = chan_of_slot[slot];
      const int block_ctx = kDefaultBlockCtxMap[(c < 2 ? c ^ 1 : 2) * kNumOrders + ord];
      const uint8_t* nzp = &f.nzeros[(size_t)c * nblk];
      const int lx = bx - gx0, ly = by - gy0;
      int pred;
      if (lx == 0) pred = ly == 0 ? 32 : nzp[(size_t)(by - 1) * fd.bxs + bx];
      else if (ly == 0) pred = nzp[(size_t)by * fd.bxs + bx - 1];
      else pred = (nzp[(size_t)(by - 1) * fd.bxs + bx] + nzp[(size_t)by * fd.bxs + bx - 1] + 1) / 2;
      int nz = f.nz_count[(size_t)c * nblk + (size_t)by * fd.bxs + bx];
      {
        int p = pred >= 64 ? 64 : pred;
        const int bucket = p < 8 ? p : 4 + p / 2;
        out->push_back(((uint32_t)(bucket * kNumBlockCtx + block_ctx) << 16) | (uint32_t)nz);
      }
      if (nz == 0) continue;
      const int histo_offset = kNumBlockCtx * kNonZeroBuckets + kZeroDensityContextCount * block_ctx;
      int prev = nz > size / 16 ? 0 : 1;
      for (int k = n; k < size && nz != 0; ++k) {
        const int j = k / 64;
        const int cbx = bx + (j % cx), cby = by + (j / cx);
        const int gg = (cby / 32) * fd.gxs + (cbx / 32);
        const size_t blk = (size_t)gg * 1024 + (size_t)(cby % 32) * 32 + (cbx % 32);
        const int coeff = f.coeffs[(blk * 3 + slot) * 64 + (k % 64)];
        const int nzl = (nz + n - 1) >> log2n;
        const int ctx = histo_offset + (kCoeffNumNonzeroContext[nzl] + kCoeffFreqContext[k >> log2n]) * 2 + prev;
        out->push_back(((uint32_t)ctx << 16) | PackSigned(coeff));
        prev = coeff != 0;
        nz -= prev;
      }
    }
  }
}

void TokenizeFrame(Frame* f) {
  const FrameDim& fd = f->fd;
  f->token_offsets.assign(fd.num_groups + 1, 0);
  f->tokens.clear();
  for (int g = 0; g < fd.num_groups; ++g) {
    TokenizeGroup(*f, g, &f->tokens);
    f->token_offsets[g + 1] = (uint32_t)f->tokens.size();
  }
  f->histograms.assign((size_t)kNumAcContexts * kAcAlphabet, 0);
  for (uint32_t t : f->tokens) {
    uint32_t tok, nb, bits;
    HybridEncode(t & 0xFFFF, &tok, &nb, &bits);
    f->histograms[(size_t)(t >> 16) * kAcAlphabet + tok]++;
  }
}

}  // namespace jxo
