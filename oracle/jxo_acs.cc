// ORACLE (test infrastructure) — stage U4: AC-strategy (block partition) search with the thesis'
// two hooks.  Restates libjxl lib/jxl/enc_ac_strategy.cc (the file the proposals patch) as it looks
// in the diffs' pre-image (blob 4aafd7a5, proposals/*.diff:2): the function bodies are recalled
// [UPSTREAM], the call structure is pinned by the diff context lines cited below.
//
//  * H8  proposals/homogeneity-partitioning.diff:213-235, hook :272-276 (combined.diff:270-274): after the
//        8x8 search a block whose winner is plain DCT8 is overridden by HomogeneityPartition(); its
//        entropy estimate is NOT recomputed.
//  * H9  proposals/homogeneity-factored-entropy.diff:248-253 (combined.diff:248-253): every EstimateEntropy()
//        result is multiplied by 0.8 * (r_h + r_v + r_d) / 3 of the candidate's top-left 8x8 block (double
//        multiply, narrowed on return).  A NaN candidate loses the 8x8 argmin (`entropy < best`, :266) and
//        every `<` / std::min test of FindBestFirstLevelDivisionForSquare, but is ACCEPTED by TryMergeAcs
//        (`if (entropy_candidate >= entropy_current) return;`, combined.diff context "@@ -602,7 +835,7").
//  * H10 control flow of ProcessRectACS (combined.diff context @@ -911 .. -1010): per 64x64 tile, 8x8 search ->
//        merge table {16X8, 8X16, 16X32, 32X16, 64X32, 32X64} where the aligned 2-, 4- and 8-block squares go
//        through FindBestFirstLevelDivisionForSquare(2 | 4 | 8, ...) and everything the squares do not cover
//        (last column / row of ragged tiles, the two 64X32 halves, the lower 32X64) through
//        TryMergeAcs(type, ..., priority) -> non-aligned 16-level squares ((cy | cx) % 2 != 0) -> non-aligned
//        32-level squares (step 2 below glacier).  Efforts: < 5 no search; 5 (hare) no DCT4X8 / DCT8X4
//        candidates and no non-aligned passes; 6..9 the full search.
//
// Defined behaviour where the recalled control flow would leave an invalid map: TryMergeAcs only looks at
// `priority`, which FindBestFirstLevelDivisionForSquare never sets, so the lower 32X64 candidate could be
// accepted on top of a 64X32 / 64X64 the square pass chose (always when it is NaN under H9).  This
// restatement rejects a TryMergeAcs candidate whose rectangle is straddled by an existing transform.
//
// Candidate set: DCT, DCT4X4, DCT2X2, DCT4X8, DCT8X4, IDENTITY at the 8x8 level (AFV0-3 are left out: their
// 16x16 basis is a table of constants that cannot be derived offline), every DCT size of the merge table.
// parity unpinned (see jxo.h).
//
// Numerics contract (DESIGN.md "Numerics"): a sum over a transform is defined as per-lane sequential sums
// followed by an xor-butterfly over the lanes.  Entropy term: lane = horizontal frequency hf, sequential
// over vf (DCT4X4 / DCT4X8 / DCT8X4: the hf of each sub-block column, see storage_index; DCT2X2 / IDENTITY:
// lane = storage row, sequential over the row).  Loss term: lane = pixel row, sequential over x.
#include "jxo_frame.h"
#include "jxo_stages.h"

#include <cfloat>

namespace jxo {

namespace {

struct AcsConfig {
  float info_loss_multiplier, zeros_mul, cost_delta;
  float distance;
  int speed_tier;          // 10 - effort (libjxl SpeedTier: hare = 5, squirrel = 3, tortoise = 1)
  bool factored_entropy;   // H9 active
  bool partitioning;       // H8 active
};

float ButterflySum(const float* lanes, int n) {
  float p[64], q[64];
  for (int i = 0; i < n; ++i) p[i] = lanes[i];
  for (int st = n / 2; st >= 1; st /= 2) {
    for (int i = 0; i < n; ++i) q[i] = p[i] + p[i ^ st];
    for (int i = 0; i < n; ++i) p[i] = q[i];
  }
  return p[0];
}

}  // namespace

// Chroma from luma (row U3, opt-in: kFlagCfl) — per 64x64 tile the factors ytox / ytob (int8, units of 1/84) that the
// search, the coefficient stage and the decoder apply as X -= (0 + ytox/84) Y, B -= (1 + ytob/84) Y.  libjxl fits them in
// enc_chroma_from_luma.cc with a robust iteration whose constants are not available offline; any map is a legal encoder
// choice.  This is the ridge least-squares fit on the quantiser-weighted DCT8 AC coefficients of the tile's blocks:
//   a_k = w_c[k] Y_k,  r_k = w_c[k] (C_k - base_c Y_k),  factor_c = sum(a r) / (sum(a a) + 0.25 n),  n = 63 x blocks,
// every block's partial sums sequential over k = 1..63 (fused multiply-adds), the tile's sums a butterfly over its
// 64 block slots (absent blocks add 0) — the order the CUDA kernel (k_cfl_fit) reproduces bit for bit.
void ChromaFromLumaFit(Frame* f) {
  const FrameDim& fd = f->fd;
  const EncTables& T = GetTables();
  const float* w = T.weights[0].data();   // DCT8, [c * 64 + position]
  for (int ty = 0; ty < fd.tys; ++ty) for (int tx = 0; tx < fd.txs; ++tx) {
    float sxy[64], sxx[64], sby[64], sbb[64];
    int blocks = 0;
    for (int i = 0; i < 64; ++i) {
      sxy[i] = sxx[i] = sby[i] = sbb[i] = 0.0f;
      const int bx = tx * 8 + (i & 7), by = ty * 8 + (i >> 3);
      if (bx >= fd.bxs || by >= fd.bys) continue;
      ++blocks;
      float cf[3][64];
      for (int c = 0; c < 3; ++c) TransformFromPixels(DCT, &f->xyb[c][(size_t)by * 8 * fd.pitch + (size_t)bx * 8], fd.pitch, cf[c]);
      float axy = 0.0f, axx = 0.0f, aby = 0.0f, abb = 0.0f;
      for (int k = 1; k < 64; ++k) {
        const float ax = w[k] * cf[1][k], rx = w[k] * cf[0][k];
        axy = fmaf(ax, rx, axy); axx = fmaf(ax, ax, axx);
        const float ab = w[128 + k] * cf[1][k], rb = w[128 + k] * (cf[2][k] - cf[1][k]);
        aby = fmaf(ab, rb, aby); abb = fmaf(ab, ab, abb);
      }
      sxy[i] = axy; sxx[i] = axx; sby[i] = aby; sbb[i] = abb;
    }
    const float ridge = 0.25f * (float)(63 * blocks);
    const float fx = ButterflySum(sxy, 64) / (ButterflySum(sxx, 64) + ridge);
    const float fb = ButterflySum(sby, 64) / (ButterflySum(sbb, 64) + ridge);
    auto to_i8 = [](float v) { const float r = rintf(v * 84.0f); return (int8_t)(r < -128.0f ? -128.0f : (r > 127.0f ? 127.0f : r)); };
    f->cmap[(size_t)ty * fd.txs + tx] = to_i8(fx);
    f->cmap[(size_t)fd.txs * fd.tys + (size_t)ty * fd.txs + tx] = to_i8(fb);
  }
}

namespace {

inline bool IsPlainDct(int s) {
  return s == DCT || s == DCT16X16 || s == DCT32X32 || (s >= DCT16X8 && s <= DCT16X32) || (s >= DCT64X64 && s <= DCT32X64);
}

// libjxl enc_ac_strategy.cc EstimateEntropy (restated; last lines pinned by combined.diff:245-253) + H9
float EstimateEntropy(const Frame& f, const AcsConfig& cfg, int s, float entropy_mul, int bx, int by) {
  const FrameDim& fd = f.fd;
  const EncTables& T = GetTables();
  const int cx = kCoveredX[s], cy = kCoveredY[s], n = cx * cy;
  const int rows = cy * 8, cols = cx * 8, size = rows * cols;
  // quant_norm16: the block's quant, the larger of two, or the 16-norm mean over more
  float quant_norm16 = 0.0f;
  if (n == 1) {
    quant_norm16 = f.qf_float[(size_t)by * fd.bxs + bx];
  } else if (n == 2) {
    const float a = f.qf_float[(size_t)by * fd.bxs + bx];
    const float b = cy == 2 ? f.qf_float[(size_t)(by + 1) * fd.bxs + bx] : f.qf_float[(size_t)by * fd.bxs + bx + 1];
    quant_norm16 = std::max(a, b);
  } else {
    for (int iy = 0; iy < cy; ++iy) for (int ix = 0; ix < cx; ++ix) {
      float qval = f.qf_float[(size_t)(by + iy) * fd.bxs + bx + ix];
      qval *= qval; qval *= qval; qval *= qval;
      quant_norm16 += qval * qval;
    }
    quant_norm16 /= (float)n;
    quant_norm16 = FastPowf(quant_norm16, 1.0f / 16.0f);
  }
  const int kind = kQuantKind[s];
  const float* weights = T.weights[kind].data();
  const float* dequant = T.dequant[kind].data();
  std::vector<float> coef((size_t)3 * size), mem(size), pix((size_t)rows * cols);
  for (int c = 0; c < 3; ++c)
    TransformFromPixels(s, &f.xyb[c][(size_t)by * 8 * fd.pitch + (size_t)bx * 8], fd.pitch, &coef[(size_t)c * size]);
  const int tx = bx / 8, ty = by / 8;
  const float cmapf[3] = {0.0f + (float)f.cmap[(size_t)ty * fd.txs + tx] / 84.0f, 0.0f,
                          1.0f + (float)f.cmap[(size_t)fd.txs * fd.tys + (size_t)ty * fd.txs + tx] / 84.0f};
  // pow(kChannelMul[c], 8.0) of {10.2, 1.0, 1.03}, narrowed to float
  static const float kChannelMul8[3] = {1.1716594e+08f, 1.0f, 1.2667701e+00f};
  // lane / sequence position of storage index k (see the numerics note above)
  const bool plain = IsPlainDct(s);
  const int lanes = plain ? cols : 8;
  const int per_lane = size / lanes;
  auto storage_index = [&](int lane, int j) {
    if (plain) return rows >= cols ? lane * rows + j : j * cols + lane;   // (hf = lane, vf = j)
    // split 8x8 strategies: the lane that owns horizontal frequency hf of sub-block column q (lane = q * 4 + hf)
    if (s == DCT4X4) return ((j >> 2) + (lane & 3) * 2) * 8 + (lane >> 2) + (j & 3) * 2;   // j = sub-block row * 4 + vf
    if (s == DCT4X8) return ((lane >> 2) + (lane & 3) * 2) * 8 + j;                        // j = vf
    if (s == DCT8X4) return ((j >> 2) + (j & 3) * 2) * 8 + lane;                           // lane = hf, j = half * 4 + vf
    return lane * 8 + j;                                                                   // DCT2X2, IDENTITY: storage rows
  };
  float entropy = 0.0f, loss_sum = 0.0f;
  for (int c = 0; c < 3; ++c) {
    float ent_lane[64];
    int nz = 0;
    for (int l = 0; l < lanes; ++l) {
      float acc = 0.0f;
      for (int j = 0; j < per_lane; ++j) {
        const int k = storage_index(l, j);
        const float in = coef[(size_t)c * size + k];
        const float v_in = cmapf[c] == 0.0f ? in : fmaf(-cmapf[c], coef[(size_t)size + k], in);
        const float val = v_in * (weights[(size_t)c * size + k] * quant_norm16);
        const float rval = rintf(val);
        const float diff = val - rval;
        mem[k] = dequant[(size_t)c * size + k] * diff;
        acc += sqrtf(fabsf(rval));
        nz += rval != 0.0f;
      }
      ent_lane[l] = acc;
    }
    float ent = ButterflySum(ent_lane, lanes) * cfg.cost_delta;
    const int nbits = CeilLog2((uint32_t)nz + 1) + 1;
    ent += cfg.zeros_mul * (float)(CeilLog2((uint32_t)nbits + 17) + nbits);
    entropy += ent;
    // information loss: masked 8-norm of the quantisation error in the pixel domain
    TransformToPixels(s, mem.data(), pix.data(), cols);
    float loss_row[64];
    for (int r = 0; r < rows; ++r) {
      float acc = 0.0f;
      const float* m = &f.mask1x1[(size_t)(by * 8 + r) * fd.pitch + (size_t)bx * 8];
      for (int x = 0; x < cols; ++x) {
        const float t = fabsf(m[x]) * pix[(size_t)r * cols + x];
        const float t2 = t * t, t4 = t2 * t2;
        acc += t4 * t4;
      }
      loss_row[r] = acc;
    }
    loss_sum += kChannelMul8[c] * ButterflySum(loss_row, rows);
  }
  const float npx = (float)(n * 64);
  const float loss_scalar = sqrtf(sqrtf(sqrtf(loss_sum / npx))) * npx / quant_norm16;
  float ret = entropy * entropy_mul;
  ret += cfg.info_loss_multiplier * loss_scalar;
  if (cfg.factored_entropy) {
    const float* r = &f.homog[((size_t)by * fd.bxs + bx) * 3];
    const float avg_r = (r[0] + r[1] + r[2]) / 3;
    ret = (float)(((double)ret * 0.8) * (double)avg_r);
  }
  return ret;
}

// ---- AcStrategyImage view of Frame::acs ------------------------------------------------------------
struct AcsImage {
  Frame* f;
  int xsize() const { return f->fd.bxs; }
  int ysize() const { return f->fd.bys; }
  bool IsFirst(int x, int y) const { return f->acs[(size_t)y * f->fd.bxs + x] & 0x80; }
  int Raw(int x, int y) const { return f->acs[(size_t)y * f->fd.bxs + x] & 0x7f; }
  void Set(int x, int y, int s) {
    for (int iy = 0; iy < kCoveredY[s]; ++iy) for (int ix = 0; ix < kCoveredX[s]; ++ix)
      f->acs[(size_t)(y + iy) * f->fd.bxs + x + ix] = (uint8_t)(s | ((ix == 0 && iy == 0) ? 0x80 : 0));
  }
};

// libjxl MultiBlockTransformCrossesHorizontalBoundary: does a transform that started above row y reach into it
// somewhere in [start_x, end_x)?
bool CrossesHorizontalBoundary(const AcsImage& a, int start_x, int y, int end_x) {
  if (start_x >= a.xsize() || y >= a.ysize()) return false;
  if (y % 8 == 0) return false;   // nothing crosses 64x64 boundaries
  end_x = std::min(end_x, a.xsize());
  const int start_x_limit = start_x & ~7;
  while (start_x != start_x_limit && !a.IsFirst(start_x, y)) --start_x;
  for (int x = start_x; x < end_x;) {
    if (a.IsFirst(x, y)) x += kCoveredX[a.Raw(x, y)];
    else return true;
  }
  return false;
}

bool CrossesVerticalBoundary(const AcsImage& a, int x, int start_y, int end_y) {
  if (x >= a.xsize() || start_y >= a.ysize()) return false;
  if (x % 8 == 0) return false;
  end_y = std::min(end_y, a.ysize());
  const int start_y_limit = start_y & ~7;
  while (start_y != start_y_limit && !a.IsFirst(x, start_y)) --start_y;
  for (int y = start_y; y < end_y;) {
    if (a.IsFirst(x, y)) y += kCoveredY[a.Raw(x, y)];
    else return true;
  }
  return false;
}

void SetEntropyForTransform(int cx, int cy, int s, float entropy, float* entropy_estimate) {
  for (int dy = 0; dy < kCoveredY[s]; ++dy) for (int dx = 0; dx < kCoveredX[s]; ++dx) entropy_estimate[(cy + dy) * 8 + cx + dx] = 0.0f;
  entropy_estimate[cy * 8 + cx] = entropy;
}

inline float StdMin(float a, float b) { return (b < a) ? b : a; }   // std::min(a, b): NaN in `a` stays

// libjxl FindBest8x8Transform (candidate table recalled, hook H8 pinned by combined.diff:270-274)
int FindBest8x8Transform(const Frame& f, const AcsConfig& cfg, int bx, int by, float* entropy_out) {
  struct Try { int type; int tier_max; double mul; };
  static const Try kTransforms8x8[] = {
      {DCT, 9, 0.8}, {DCT4X4, 5, 1.08}, {DCT2X2, 5, 0.95}, {DCT4X8, 4, 0.85931637428340035},
      {DCT8X4, 4, 0.85931637428340035}, {IDENTITY, 5, 1.0427542510634957},
      // {AFV0..3, 4, 0.81779489591359944}: not built (basis constants not derivable offline)
  };
  double best = 1e30;
  int best_tx = DCT;
  const float d = cfg.distance;
  for (const Try& tx : kTransforms8x8) {
    if (tx.tier_max < cfg.speed_tier) continue;
    float entropy_mul = (float)(tx.mul / kTransforms8x8[0].mul);
    if ((tx.type == DCT2X2 || tx.type == IDENTITY) && d < 5.0f) {
      const float kFavor2X2AtHighQuality = 0.4f;
      const float w = (5.0f - d) / 5.0f;
      entropy_mul -= kFavor2X2AtHighQuality * (w * w);
    }
    if (tx.type != DCT && tx.type != DCT2X2 && tx.type != IDENTITY && d > 4.0f) {
      const float kAvoidEntropyOfTransforms = 0.5f;
      float mul = 1.0f;
      if (d < 12.0f) mul *= (12.0f - 4.0f) / (d - 4.0f);
      entropy_mul += kAvoidEntropyOfTransforms * mul;      // (combined.diff context "@@ -566,7 +793,7")
    }
    const float entropy = EstimateEntropy(f, cfg, tx.type, entropy_mul, bx, by);
    if ((double)entropy < best) { best_tx = tx.type; best = (double)entropy; }
  }
  *entropy_out = (float)best;
  if (cfg.partitioning && best_tx == DCT) {
    const float* r = &f.homog[((size_t)by * f.fd.bxs + bx) * 3];
    best_tx = HomogeneityPartition(r[0], r[1], r[2], d);
  }
  return best_tx;
}

// libjxl TryMergeAcs (combined.diff context "@@ -586,7 +819,7" .. "@@ -602,7 +835,7")
void TryMergeAcs(Frame* f, const AcsConfig& cfg, int s, int bx, int by, int cx, int cy, float entropy_mul,
                 uint8_t candidate_priority, uint8_t* priority, float* entropy_estimate) {
  AcsImage a{f};
  const int cvx = kCoveredX[s], cvy = kCoveredY[s];
  float entropy_current = 0.0f;
  for (int iy = 0; iy < cvy; ++iy) for (int ix = 0; ix < cvx; ++ix) {
    if (priority[(cy + iy) * 8 + cx + ix] >= candidate_priority) return;   // would reuse allocated blocks
    entropy_current += entropy_estimate[(cy + iy) * 8 + cx + ix];
  }
  // defined behaviour (see the header): an existing transform must not straddle the candidate's rectangle
  if (CrossesHorizontalBoundary(a, bx + cx, by + cy, bx + cx + cvx) || CrossesHorizontalBoundary(a, bx + cx, by + cy + cvy, bx + cx + cvx) ||
      CrossesVerticalBoundary(a, bx + cx, by + cy, by + cy + cvy) || CrossesVerticalBoundary(a, bx + cx + cvx, by + cy, by + cy + cvy)) return;
  const float entropy_candidate = EstimateEntropy(*f, cfg, s, entropy_mul, bx + cx, by + cy);
  if (entropy_candidate >= entropy_current) return;
  for (int iy = 0; iy < cvy; ++iy) for (int ix = 0; ix < cvx; ++ix) {
    entropy_estimate[(cy + iy) * 8 + cx + ix] = 0.0f;
    priority[(cy + iy) * 8 + cx + ix] = candidate_priority;
  }
  a.Set(bx + cx, by + cy, s);
  entropy_estimate[cy * 8 + cx] = entropy_candidate;
}

// libjxl FindBestFirstLevelDivisionForSquare (combined.diff context "@@ -669,7 +902,7" .. "@@ -748,7 +981,7")
void FindBestFirstLevelDivisionForSquare(Frame* f, const AcsConfig& cfg, int blocks, bool allow_square_transform, int bx, int by,
                                         int cx, int cy, float entropy_mul_JXK, float entropy_mul_JXJ, float* entropy_estimate) {
  AcsImage a{f};
  const int blocks_half = blocks / 2;
  const int acs_rawJXK = blocks == 2 ? DCT16X8 : (blocks == 4 ? DCT32X16 : DCT64X32);   // J rows x K columns
  const int acs_rawKXJ = blocks == 2 ? DCT8X16 : (blocks == 4 ? DCT16X32 : DCT32X64);
  const int acs_rawJXJ = blocks == 2 ? DCT16X16 : (blocks == 4 ? DCT32X32 : DCT64X64);
  // can a JXJ block be considered here at all? (needed for the 'floating' positions)
  if (CrossesHorizontalBoundary(a, bx + cx, by + cy, bx + cx + blocks) ||
      CrossesHorizontalBoundary(a, bx + cx, by + cy + blocks, bx + cx + blocks) ||
      CrossesVerticalBoundary(a, bx + cx, by + cy, by + cy + blocks) ||
      CrossesVerticalBoundary(a, bx + cx + blocks, by + cy, by + cy + blocks)) return;
  const bool allow_JXK = !CrossesVerticalBoundary(a, bx + cx + blocks_half, by + cy, by + cy + blocks);
  const bool allow_KXJ = !CrossesHorizontalBoundary(a, bx + cx, by + cy + blocks_half, bx + cx + blocks);
  float entropy[2][2] = {};
  for (int dy = 0; dy < blocks; ++dy) for (int dx = 0; dx < blocks; ++dx)
    entropy[dy / blocks_half][dx / blocks_half] += entropy_estimate[(cy + dy) * 8 + cx + dx];
  float entropy_JXK_left = FLT_MAX, entropy_JXK_right = FLT_MAX, entropy_KXJ_top = FLT_MAX, entropy_KXJ_bottom = FLT_MAX,
        entropy_JXJ = FLT_MAX;
  if (allow_JXK) {
    if (a.Raw(bx + cx, by + cy) != acs_rawJXK)
      entropy_JXK_left = EstimateEntropy(*f, cfg, acs_rawJXK, entropy_mul_JXK, bx + cx, by + cy);
    if (a.Raw(bx + cx + blocks_half, by + cy) != acs_rawJXK)
      entropy_JXK_right = EstimateEntropy(*f, cfg, acs_rawJXK, entropy_mul_JXK, bx + cx + blocks_half, by + cy);
  }
  if (allow_KXJ) {
    if (a.Raw(bx + cx, by + cy) != acs_rawKXJ)
      entropy_KXJ_top = EstimateEntropy(*f, cfg, acs_rawKXJ, entropy_mul_JXK, bx + cx, by + cy);
    if (a.Raw(bx + cx, by + cy + blocks_half) != acs_rawKXJ)
      entropy_KXJ_bottom = EstimateEntropy(*f, cfg, acs_rawKXJ, entropy_mul_JXK, bx + cx, by + cy + blocks_half);
  }
  if (allow_square_transform)
    entropy_JXJ = EstimateEntropy(*f, cfg, acs_rawJXJ, entropy_mul_JXJ, bx + cx, by + cy);
  // the square can have JXK or KXJ transforms, not both
  const float costJxN = StdMin(entropy_JXK_left, entropy[0][0] + entropy[1][0]) + StdMin(entropy_JXK_right, entropy[0][1] + entropy[1][1]);
  const float costNxJ = StdMin(entropy_KXJ_top, entropy[0][0] + entropy[0][1]) + StdMin(entropy_KXJ_bottom, entropy[1][0] + entropy[1][1]);
  if (entropy_JXJ < costJxN && entropy_JXJ < costNxJ) {
    a.Set(bx + cx, by + cy, acs_rawJXJ);
    SetEntropyForTransform(cx, cy, acs_rawJXJ, entropy_JXJ, entropy_estimate);
  } else if (costJxN < costNxJ) {
    if (entropy_JXK_left < entropy[0][0] + entropy[1][0]) {
      a.Set(bx + cx, by + cy, acs_rawJXK);
      SetEntropyForTransform(cx, cy, acs_rawJXK, entropy_JXK_left, entropy_estimate);
    }
    if (entropy_JXK_right < entropy[0][1] + entropy[1][1]) {
      a.Set(bx + cx + blocks_half, by + cy, acs_rawJXK);
      SetEntropyForTransform(cx + blocks_half, cy, acs_rawJXK, entropy_JXK_right, entropy_estimate);
    }
  } else {
    if (entropy_KXJ_top < entropy[0][0] + entropy[0][1]) {
      a.Set(bx + cx, by + cy, acs_rawKXJ);
      SetEntropyForTransform(cx, cy, acs_rawKXJ, entropy_KXJ_top, entropy_estimate);
    }
    if (entropy_KXJ_bottom < entropy[1][0] + entropy[1][1]) {
      a.Set(bx + cx, by + cy + blocks_half, acs_rawKXJ);
      SetEntropyForTransform(cx, cy + blocks_half, acs_rawKXJ, entropy_KXJ_bottom, entropy_estimate);
    }
  }
}

// libjxl ProcessRectACS for one 64x64 tile: (bx, by) first block, rxs x rys blocks (combined.diff context @@ -911 .. -1010)
void ProcessRectACS(Frame* f, const AcsConfig& cfg, int bx, int by, int rxs, int rys) {
  AcsImage a{f};
  const FrameDim& fd = f->fd;
  float entropy_estimate[64] = {};
  const float mul8x8 = 1.0f + -0.4f / (cfg.distance + 1.4f);
  for (int iy = 0; iy < rys; ++iy) for (int ix = 0; ix < rxs; ++ix) {
    float entropy = 0.0f;
    const int best_of_8x8s = FindBest8x8Transform(*f, cfg, bx + ix, by + iy, &entropy);
    a.Set(bx + ix, by + iy, best_of_8x8s);
    entropy_estimate[iy * 8 + ix] = entropy * mul8x8;
  }
  struct MergeTry { int type; uint8_t priority; uint8_t decoding_speed_tier_max_limit; float entropy_mul; };
  const float entropy_mul16X8 = 1.25f, entropy_mul16X16 = 1.35f, entropy_mul16X32 = 1.5f, entropy_mul32X32 = 1.5f,
              entropy_mul64X32 = 2.26f, entropy_mul64X64 = 2.26f;
  const MergeTry kTransformsForMerge[6] = {
      {DCT16X8, 2, 4, entropy_mul16X8},   {DCT8X16, 2, 4, entropy_mul16X8},   {DCT16X32, 4, 4, entropy_mul16X32},
      {DCT32X16, 4, 4, entropy_mul16X32}, {DCT64X32, 6, 1, entropy_mul64X32}, {DCT32X64, 6, 1, entropy_mul64X32},
  };
  uint8_t priority[64] = {};
  const int decoding_speed_tier = 0;
  const bool enable_32x32 = decoding_speed_tier < 4;
  for (const MergeTry& tx : kTransformsForMerge) {
    if (tx.decoding_speed_tier_max_limit < decoding_speed_tier) continue;
    const int cvx = kCoveredX[tx.type], cvy = kCoveredY[tx.type];
    for (int cy = 0; cy + cvy - 1 < rys; cy += cvy) {
      for (int cx = 0; cx + cvx - 1 < rxs; cx += cvx) {
        if (cy + 7 < rys && cx + 7 < rxs) {
          if (decoding_speed_tier < 4 && tx.type == DCT32X64) {
            if ((cy | cx) % 8 == 0)
              FindBestFirstLevelDivisionForSquare(f, cfg, 8, true, bx, by, cx, cy, tx.entropy_mul, entropy_mul64X64, entropy_estimate);
            continue;
          } else if (tx.type == DCT32X16) {
            continue;
          }
        }
        if ((tx.type == DCT16X32 && cy % 4 != 0) || (tx.type == DCT32X16 && cx % 4 != 0)) continue;   // covered by the 32x32 squares
        if (cy + 3 < rys && cx + 3 < rxs) {
          if (tx.type == DCT16X32) {
            if ((cy | cx) % 4 == 0)
              FindBestFirstLevelDivisionForSquare(f, cfg, 4, enable_32x32, bx, by, cx, cy, tx.entropy_mul, entropy_mul32X32, entropy_estimate);
            continue;
          } else if (tx.type == DCT32X16) {
            continue;
          }
        }
        if ((tx.type == DCT16X32 && cy % 4 != 0) || (tx.type == DCT32X16 && cx % 4 != 0)) continue;
        if (cy + 1 < rys && cx + 1 < rxs) {
          if (tx.type == DCT8X16) {
            if ((cy | cx) % 2 == 0)
              FindBestFirstLevelDivisionForSquare(f, cfg, 2, true, bx, by, cx, cy, tx.entropy_mul, entropy_mul16X16, entropy_estimate);
            continue;
          } else if (tx.type == DCT16X8) {
            continue;
          }
        }
        // no square covers this position: the normal integral transform merging process
        TryMergeAcs(f, cfg, tx.type, bx, by, cx, cy, tx.entropy_mul, tx.priority, priority, entropy_estimate);
      }
    }
  }
  if (cfg.speed_tier < 5) {   // `if (cparams.speed_tier >= SpeedTier::kHare) return;`
    // non-aligned matching: a few more 16X8, 8X16 and 16X16 between the non-2-aligned blocks
    for (int cy = 0; cy + 1 < rys; ++cy) for (int cx = 0; cx + 1 < rxs; ++cx) {
      if ((cy | cx) % 2 != 0)
        FindBestFirstLevelDivisionForSquare(f, cfg, 2, true, bx, by, cx, cy, entropy_mul16X8, entropy_mul16X16, entropy_estimate);
    }
    // non-aligned matching for 32X32, 16X32 and 32X16
    const int step = cfg.speed_tier >= 1 ? 2 : 1;   // kTortoise and faster
    for (int cy = 0; cy + 3 < rys; cy += step) for (int cx = 0; cx + 3 < rxs; cx += step) {
      if ((cy | cx) % 4 == 0) continue;   // already tried with the aligned loop
      FindBestFirstLevelDivisionForSquare(f, cfg, 4, enable_32x32, bx, by, cx, cy, entropy_mul16X32, entropy_mul32X32, entropy_estimate);
    }
  }
  for (int iy = 0; iy < rys; ++iy) for (int ix = 0; ix < rxs; ++ix)
    f->acs_entropy[(size_t)(by + iy) * fd.bxs + bx + ix] = entropy_estimate[iy * 8 + ix];
}

}  // namespace

void AcStrategySearch(Frame* f) {
  const FrameDim& fd = f->fd;
  const Params& p = f->params;
  if (p.effort < 5) return;   // AcStrategyHeuristics::ProcessRect: DCT8 everywhere at cheetah and faster
  AcsConfig cfg;
  const float ratio = (p.distance + 0.1373f) / 1.1373f;
  cfg.info_loss_multiplier = 1.2f * powf(ratio, 0.33677806662454718f);
  cfg.zeros_mul = 9.3089171683409026f * powf(ratio, 0.50990926717963703f);
  cfg.cost_delta = 10.833273317067883f * powf(ratio, 0.36702940662370243f);
  cfg.distance = p.distance;
  cfg.speed_tier = 10 - (int)p.effort;
  cfg.partitioning = p.proposal == 1 || p.proposal == 3;
  cfg.factored_entropy = p.proposal == 2 || p.proposal == 3;
  for (int ty = 0; ty < fd.tys; ++ty) for (int tx = 0; tx < fd.txs; ++tx)
    ProcessRectACS(f, cfg, tx * 8, ty * 8, std::min(8, fd.bxs - tx * 8), std::min(8, fd.bys - ty * 8));
}

}  // namespace jxo
