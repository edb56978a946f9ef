// ORACLE (test infrastructure) — stage U9: frame assembly.  Codestream headers, frame header,
// TOC and the LfGlobal / LfGroup / HfGlobal / PassGroup sections of a single-pass VarDCT frame.
// Restates libjxl enc_frame.cc, frame_header.cc, headers.cc, enc_toc.cc, enc_modular.cc,
// quantizer.cc (field orders and codings) [UPSTREAM, recalled; SURVEY.md Appendix U.1, U.13, U.14,
// U.19]. parity unpinned.  Encoder choices (all legal per Appendix U.19): gaborish off, EPF off,
// adaptive DC smoothing skipped, default quant tables / coefficient orders / block-context map,
// one histogram set, ANS for AC tokens, prefix codes for the modular streams.
#include "jxo_entropy.h"
#include "jxo_frame.h"
#include "jxo_stages.h"

namespace jxo {

void TokenizeFrame(Frame* f);  // jxo_entropy.cc

static void WriteU32Size(BitWriter* w, uint32_t v) {  // SizeHeader U32(BitsOffset(9,1), (13,1), (18,1), (30,1))
  const uint32_t m = v - 1;
  if (m < (1u << 9)) { w->Write(2, 0); w->Write(9, m); }
  else if (m < (1u << 13)) { w->Write(2, 1); w->Write(13, m); }
  else if (m < (1u << 18)) { w->Write(2, 2); w->Write(18, m); }
  else { w->Write(2, 3); w->Write(30, m); }
}

void WriteCodestreamHeaders(const Frame& f, BitWriter* w) {
  w->Write(8, 0xFF); w->Write(8, 0x0A);
  // SizeHeader: small = 0, ysize, ratio = 0, xsize
  w->Write(1, 0);
  WriteU32Size(w, (uint32_t)f.fd.ysize);
  w->Write(3, 0);
  WriteU32Size(w, (uint32_t)f.fd.xsize);
  w->Write(1, 1);  // ImageMetadata.all_default: 8-bit, sRGB, XYB-encoded, no extra channels
  w->Write(1, 1);  // default_m (CustomTransformData.all_default)
  w->ZeroPadToByte();
}

void WriteFrameHeader(const Frame& f, BitWriter* w) {
  w->Write(1, 0);      // all_default = false
  w->Write(2, 0);      // frame_type = kRegularFrame
  w->Write(1, 0);      // encoding = VarDCT
  w->WriteU64(128);    // flags = kSkipAdaptiveDCSmoothing
  w->Write(2, 0);      // upsampling = 1
  w->Write(3, (uint64_t)f.q.x_qm_scale);
  w->Write(3, (uint64_t)f.q.b_qm_scale);
  w->Write(2, 0);      // passes.num_passes = 1
  w->Write(1, 0);      // no custom size or origin
  w->Write(2, 0);      // blending_info.mode = kReplace
  w->Write(1, 1);      // is_last
  w->Write(2, 0);      // name length 0
  w->Write(1, 0);      // loop_filter.all_default = false
  w->Write(1, f.gab ? 1 : 0);      //   gab
  if (f.gab) w->Write(1, 0);       //   gab_custom = false (default weights)
  w->Write(2, 0);      //   epf_iters = 0
  w->Write(2, 0);      //   loop_filter extensions = 0
  w->Write(2, 0);      // frame header extensions = 0
}

static void WriteQuantizer(const QuantState& q, BitWriter* w) {
  const uint32_t gs = (uint32_t)q.global_scale;
  if (gs < 2049) { w->Write(2, 0); w->Write(11, gs - 1); }
  else if (gs < 4097) { w->Write(2, 1); w->Write(11, gs - 2049); }
  else if (gs < 8193) { w->Write(2, 2); w->Write(12, gs - 4097); }
  else { w->Write(2, 3); w->Write(16, gs - 8193); }
  const uint32_t qd = (uint32_t)q.quant_dc;
  if (qd == 16) { w->Write(2, 0); }
  else if (qd <= 32) { w->Write(2, 1); w->Write(5, qd - 1); }
  else if (qd <= 256) { w->Write(2, 2); w->Write(8, qd - 1); }
  else { w->Write(2, 3); w->Write(16, qd - 1); }
}

static void WriteModularTokens(const std::vector<uint32_t>& tokens, const PrefixCode* codes, BitWriter* w) {
  for (uint32_t t : tokens) {
    const PrefixCode& pc = codes[t >> 24];
    uint32_t tok, nbits, bits;
    HybridEncode(t & 0xFFFFFF, &tok, &nbits, &bits);
    w->Write(pc.length[tok], pc.bits[tok]);
    w->Write((int)nbits, bits);
  }
}

static void WriteTocEntry(BitWriter* w, uint32_t size) {
  if (size < 1024) { w->Write(2, 0); w->Write(10, size); }
  else if (size < 17408) { w->Write(2, 1); w->Write(14, size - 1024); }
  else if (size < 4211712) { w->Write(2, 2); w->Write(22, size - 17408); }
  else { w->Write(2, 3); w->Write(30, size - 4211712); }
}

bool EntropyCodeFrame(Frame* f) {
  const FrameDim& fd = f->fd;
  // ---- AC tokens, histograms, clustering, codes
  TokenizeFrame(f);
  f->context_map.assign(kNumAcContexts, 0);
  std::vector<uint32_t> cluster_hist;
  AnsCode ac;
  ac.alphabet = kAcAlphabet;
  ac.num_clusters = ClusterHistograms(f->histograms.data(), kNumAcContexts, kAcAlphabet, kMaxClusters,
                                      f->context_map.data(), &cluster_hist);
  ac.norm.assign((size_t)ac.num_clusters * kAcAlphabet, 0);
  for (int k = 0; k < ac.num_clusters; ++k)
    NormalizeCounts(&cluster_hist[(size_t)k * kAcAlphabet], kAcAlphabet, &ac.norm[(size_t)k * kAcAlphabet]);
  ac.Build();
  f->num_clusters = ac.num_clusters;

  // ---- modular tokens of every DC group, global prefix codes
  std::vector<std::vector<uint32_t>> dc_tok(fd.num_dc_groups), meta_tok(fd.num_dc_groups);
  std::vector<uint32_t> nfirst(fd.num_dc_groups, 0);
  std::vector<uint32_t> mod_hist((size_t)kNumModularCtx * kModAlphabet, 0);
  for (int dg = 0; dg < fd.num_dc_groups; ++dg) {
    ModularTokensDcGroup(*f, dg, &dc_tok[dg], &meta_tok[dg], &nfirst[dg]);
    for (const auto* v : {&dc_tok[dg], &meta_tok[dg]})
      for (uint32_t t : *v) {
        uint32_t tok, nb, bits;
        HybridEncode(t & 0xFFFFFF, &tok, &nb, &bits);
        mod_hist[(size_t)(t >> 24) * kModAlphabet + tok]++;
      }
  }
  PrefixCode pcs[kNumModularCtx];
  for (int l = 0; l < kNumModularCtx; ++l) BuildPrefixCode(&mod_hist[(size_t)l * kModAlphabet], kModAlphabet, &pcs[l]);

  // ---- sections
  const bool small = fd.num_groups == 1;
  std::vector<BitWriter> sec(small ? 1 : 2 + fd.num_dc_groups + fd.num_groups);
  auto out = [&](int idx) -> BitWriter& { return sec[small ? 0 : idx]; };
  {  // LfGlobal
    BitWriter& w = out(0);
    w.Write(1, 1);              // default DC dequantisation
    WriteQuantizer(f->q, &w);
    w.Write(1, 1);              // default block context map
    w.Write(1, 1);              // default colour correlation (colour factor 84, base x 0, base b 1, dc 0)
    w.Write(1, 1);              // global MA tree present
    WriteGlobalTree(fd.num_dc_groups, &w);
    // entropy code of the modular streams: 8 contexts, identity context map, prefix codes
    w.Write(1, 0);              // lz77 disabled
    w.Write(1, 1); w.Write(2, 3);
    for (int l = 0; l < kNumModularCtx; ++l) w.Write(3, (uint64_t)l);
    w.Write(1, 1);              // prefix codes
    for (int l = 0; l < kNumModularCtx; ++l) { w.Write(4, 4); w.Write(3, 2); w.Write(2, 0); }
    for (int l = 0; l < kNumModularCtx; ++l) w.WriteVarLenUint16((uint32_t)(pcs[l].alphabet - 1));
    for (int l = 0; l < kNumModularCtx; ++l) WritePrefixCodeHeader(pcs[l], &w);
  }
  for (int dg = 0; dg < fd.num_dc_groups; ++dg) {  // LfGroup
    BitWriter& w = out(1 + dg);
    const int x0 = (dg % fd.dgxs) * 256, y0 = (dg / fd.dgxs) * 256;
    const int bw = std::min(256, fd.bxs - x0), bh = std::min(256, fd.bys - y0);
    w.Write(2, 0);              // extra_precision = 0
    w.Write(1, 1); w.Write(1, 1); w.Write(2, 0);   // GroupHeader: global tree, default wp, no transforms
    WriteModularTokens(dc_tok[dg], pcs, &w);
    w.Write(CeilLog2((uint32_t)(bw * bh)), nfirst[dg] - 1);
    w.Write(1, 1); w.Write(1, 1); w.Write(2, 0);
    WriteModularTokens(meta_tok[dg], pcs, &w);
  }
  {  // HfGlobal
    BitWriter& w = out(1 + fd.num_dc_groups);
    w.Write(1, 1);              // default dequant matrices
    w.Write(CeilLog2((uint32_t)fd.num_groups), 0);  // num_histograms - 1
    w.Write(2, 2);              // used_orders = 0: natural coefficient orders
    w.Write(1, 0);              // lz77 disabled
    WriteContextMap(f->context_map.data(), kNumAcContexts, ac.num_clusters, &w);
    w.Write(1, 0);              // ANS
    w.Write(2, kLogAlphaSize - 5);
    for (int k = 0; k < ac.num_clusters; ++k) { w.Write(4, 4); w.Write(3, 2); w.Write(2, 0); }
    for (int k = 0; k < ac.num_clusters; ++k) WriteAnsHistogram(&ac.norm[(size_t)k * kAcAlphabet], kAcAlphabet, &w);
  }
  f->group_offsets.assign(fd.num_groups + 1, 0);
  f->group_streams.clear();
  for (int g = 0; g < fd.num_groups; ++g) {  // PassGroup
    BitWriter gw;
    AnsWriteTokens(&f->tokens[f->token_offsets[g]], f->token_offsets[g + 1] - f->token_offsets[g],
                   f->context_map.data(), ac, &gw);
    const std::vector<uint8_t>& gb = gw.Bytes();
    f->group_streams.insert(f->group_streams.end(), gb.begin(), gb.end());
    f->group_offsets[g + 1] = (uint32_t)f->group_streams.size();
    BitWriter& w = out(2 + fd.num_dc_groups + g);
    if (small) {  // single TOC entry: sections are bit-concatenated
      const size_t nbits = gw.BitsWritten();
      for (size_t i = 0; i < nbits; i += 8) w.Write((int)std::min<size_t>(8, nbits - i), gb[i / 8]);
    } else {
      w.AppendBytes(gb);
    }
  }
  // ---- codestream
  BitWriter cs;
  WriteCodestreamHeaders(*f, &cs);
  WriteFrameHeader(*f, &cs);
  cs.Write(1, 0);  // TOC not permuted
  cs.ZeroPadToByte();
  for (auto& s : sec) WriteTocEntry(&cs, (uint32_t)s.Bytes().size());
  cs.ZeroPadToByte();
  for (auto& s : sec) cs.AppendBytes(s.Bytes());
  f->codestream = cs.Bytes();
  return true;
}

}  // namespace jxo
