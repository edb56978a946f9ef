// ORACLE (test infrastructure) — whole-frame driver: runs every stage in order and
// keeps every intermediate so tests can compare them with the CUDA path stage by stage.
#pragma once
#include "jxo_enc.h"

namespace jxo {

// mirrors jxlb200_params (include/jxlb200.h)
struct Params {
  float distance = 1.0f;
  uint32_t effort = 7;
  uint32_t proposal = 0;  // 0 none, 1 partitioning, 2 factored-entropy, 3 combined
  uint32_t flags = 0;     // bit0: fixed DCT8 strategy; bit1: uniform quant field
};
enum : uint32_t { kFlagFixedDct8 = 1u, kFlagUniformQf = 2u, kFlagForcedAcs = 8u, kFlagGaborish = 16u, kFlagCfl = 32u };

// stage ids — identical to JXLB200_STAGE_* in include/jxlb200.h
enum Stage : int {
  kStageXyb = 1, kStageQfFloat = 2, kStageMask1x1 = 3, kStageHomog = 4, kStageAcs = 5,
  kStageRawQf = 6, kStageQuantParams = 7, kStageCoeffs = 8, kStageDcQuant = 9, kStageNzeros = 10,
  kStageTokens = 11, kStageHistograms = 12, kStageContextMap = 13, kStageGroupStreams = 14,
  kStageCodestream = 15, kStageMask = 16, kStageCmap = 17, kStageTokenOffsets = 18,
  kStageGroupOffsets = 19, kStageAcsEntropy = 20
};

struct Token { uint32_t ctx; uint32_t value; };

struct Frame {
  FrameDim fd;
  Params params;
  QuantState q;
  bool gab = false;                   // loop filter: Gaborish with the default weights (kFlagGaborish / decoded header)
  std::vector<float> xyb[3];          // ys_pad * pitch
  std::vector<float> qf_float;        // bys * bxs
  std::vector<float> mask;            // bys * bxs
  std::vector<float> mask1x1;         // ys_pad * pitch
  std::vector<float> homog;           // blocks * 3 (r_h, r_v, r_d)
  std::vector<uint8_t> forced_acs;    // kFlagForcedAcs: the strategy map to code with instead of searching (parity tap)
  std::vector<uint8_t> acs;           // bys * bxs : raw strategy | 0x80 if first block
  std::vector<float> acs_entropy;     // bys * bxs : entropy estimate left by the search
  std::vector<int32_t> raw_qf;        // bys * bxs
  std::vector<int8_t> cmap;           // 2 * tys * txs (ytox, ytob)
  std::vector<int16_t> coeffs;        // num_groups * 1024 blocks * 3 * 64, scan order
  std::vector<float> dc[3];           // bys * bxs
  std::vector<int16_t> dc_quant;      // 3 * bys * bxs (X, Y, B)
  std::vector<uint8_t> nzeros;        // 3 * bys * bxs (X, Y, B) stored per covered block (ceil-shared)
  std::vector<uint16_t> nz_count;     // 3 * bys * bxs: true non-zero count at first blocks
  std::vector<uint32_t> token_offsets;  // num_groups + 1
  std::vector<uint32_t> tokens;         // (ctx << 16) | value
  std::vector<uint32_t> histograms;     // num_ctx * kAlphabet
  std::vector<uint8_t> context_map;     // num_ctx
  std::vector<uint32_t> group_offsets;  // num_groups + 1 (bytes)
  std::vector<uint8_t> group_streams;
  std::vector<uint8_t> codestream;
  int num_clusters = 0;
  std::string error;
};

bool EncodeFrame(const uint8_t* rgb, int w, int h, size_t stride, const Params& p, Frame* f);

}  // namespace jxo
