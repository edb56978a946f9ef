// ORACLE (test infrastructure) — stage U4: AC-strategy (block partition) search with the thesis'
// two hooks.
//
//  * H8 (proposals/homogeneity-partitioning.diff:213-235, hook :272-276; combined.diff:270-274):
//    after the 8x8 search, a block whose winner is plain DCT8 is overridden by
//    HomogeneityPartition(); its entropy estimate is NOT recomputed.
//  * H9 (proposals/homogeneity-factored-entropy.diff:248-253; combined.diff:248-253): every
//    EstimateEntropy() result is multiplied by 0.8 * (r_h + r_v + r_d) / 3 of the candidate's
//    top-left 8x8 block (double multiply, narrowed on return).  NaN candidates lose the 8x8
//    argmin (`entropy < best` is false) but WIN merges (`candidate >= current` is false), exactly
//    as in the patched libjxl (diff :266, :296).
//  * H10: the control-flow shape of ProcessRectACS in the diff context (:259-401): 8x8 search ->
//    aligned 16x16 squares (square vs the two half splits) -> aligned 32x32 squares.
//
// Everything else (the cost model EstimateEntropy, the candidate multipliers) restates libjxl
// enc_ac_strategy.cc from recall [UPSTREAM, SURVEY.md Appendix U.4-7] over the transform set this
// repo emits (DCT8, 4x4, 4x8, 8x4, 16x8, 8x16, 16x16, 32x16, 16x32, 32x32); the 64-sized
// transforms, IDENTITY / DCT2X2 / AFV and the non-aligned re-tries are out of scope.
// parity unpinned.
//
// Numerics contract: block-wide float sums are per-row sequential sums followed by an xor-butterfly
// over the rows (the association a warp-shuffle reduction produces); see DESIGN.md "Numerics".
#include "jxo_frame.h"
#include "jxo_stages.h"

namespace jxo {

namespace {

struct AcsConfig {
  float info_loss_multiplier, zeros_mul, cost_delta;
  float distance;
  bool factored_entropy;   // H9 active
  bool partitioning;       // H8 active
};

float ButterflySum(const float* rows, int n) {
  float p[32], q[32];
  for (int i = 0; i < n; ++i) p[i] = rows[i];
  for (int st = n / 2; st >= 1; st /= 2) {
    for (int i = 0; i < n; ++i) q[i] = p[i] + p[i ^ st];
    for (int i = 0; i < n; ++i) p[i] = q[i];
  }
  return p[0];
}

// libjxl enc_ac_strategy.cc EstimateEntropy (restated) + H9
float EstimateEntropy(const Frame& f, const AcsConfig& cfg, int s, float entropy_mul, int bx, int by) {
  const FrameDim& fd = f.fd;
  const EncTables& T = GetTables();
  const int cx = kCoveredX[s], cy = kCoveredY[s], n = cx * cy;
  const int rows = cy * 8, cols = cx * 8, size = rows * cols;
  const int W = std::max(rows, cols), H = std::min(rows, cols);   // coefficient block: H rows of W
  const int xs = W / 8, ys = H / 8;
  float q = f.qf_float[(size_t)by * fd.bxs + bx];
  for (int iy = 0; iy < cy; ++iy) for (int ix = 0; ix < cx; ++ix) q = std::max(q, f.qf_float[(size_t)(by + iy) * fd.bxs + bx + ix]);
  const float inv_q = 1.0f / q;
  const int kind = kQuantKind[s];
  const float* weights = T.weights[kind].data();
  const float* dequant = T.dequant[kind].data();
  std::vector<float> coef((size_t)3 * size), err(size), pix((size_t)rows * cols);
  for (int c = 0; c < 3; ++c)
    TransformFromPixels(s, &f.xyb[c][(size_t)by * 8 * fd.pitch + (size_t)bx * 8], fd.pitch, &coef[(size_t)c * size]);
  const int tx = bx / 8, ty = by / 8;
  const float cmapf[3] = {0.0f + (float)f.cmap[(size_t)ty * fd.txs + tx] / 84.0f, 0.0f,
                          1.0f + (float)f.cmap[(size_t)fd.txs * fd.tys + (size_t)ty * fd.txs + tx] / 84.0f};
  static const float kChannelMul[3] = {10.2f, 1.0f, 1.03f};
  float entropy = 0.0f, loss = 0.0f;
  for (int c = 0; c < 3; ++c) {
    float ent_row[32];
    int nz = 0;
    for (int y = 0; y < H; ++y) {
      float acc = 0.0f;
      for (int x = 0; x < W; ++x) {
        const int k = y * W + x;
        if (x < xs && y < ys) { err[k] = 0.0f; continue; }
        const float v_in = c == 1 ? coef[(size_t)size + k] : fmaf(-cmapf[c], coef[(size_t)size + k], coef[(size_t)c * size + k]);
        const float val = v_in * (weights[(size_t)c * size + k] * q);
        const float rval = rintf(val);
        const float diff = val - rval;
        acc += sqrtf(fabsf(rval));
        nz += rval != 0.0f;
        err[k] = diff * (dequant[(size_t)c * size + k] * inv_q);
      }
      ent_row[y] = acc;
    }
    float ent = ButterflySum(ent_row, H) * cfg.cost_delta;
    const int nbits = CeilLog2((uint32_t)nz + 1) + 1;
    ent += cfg.zeros_mul * (float)(CeilLog2((uint32_t)nbits + 17) + nbits);
    entropy += ent;
    // information loss: masked 8-norm of the quantisation error in the pixel domain
    TransformToPixels(s, err.data(), pix.data(), cols);
    float loss_row[32];
    for (int r = 0; r < rows; ++r) {
      float acc = 0.0f;
      const float* m = &f.mask1x1[(size_t)(by * 8 + r) * fd.pitch + (size_t)bx * 8];
      for (int x = 0; x < cols; ++x) {
        const float t = pix[(size_t)r * cols + x] * m[x];
        const float t2 = t * t, t4 = t2 * t2;
        acc += t4 * t4;
      }
      loss_row[r] = acc;
    }
    const float mean8 = ButterflySum(loss_row, rows) / (float)(rows * cols);
    loss += kChannelMul[c] * sqrtf(sqrtf(sqrtf(mean8)));
  }
  const float loss_scalar = loss * (float)(n * 64) * inv_q;
  float ret = entropy * entropy_mul + cfg.info_loss_multiplier * loss_scalar;
  if (cfg.factored_entropy) {
    const float* r = &f.homog[((size_t)by * fd.bxs + bx) * 3];
    const float avg_r = (r[0] + r[1] + r[2]) / 3;
    ret = (float)(((double)ret * 0.8) * (double)avg_r);
  }
  return ret;
}

void SetStrategy(Frame* f, int s, int bx, int by, float est) {
  const FrameDim& fd = f->fd;
  const int cx = kCoveredX[s], cy = kCoveredY[s];
  for (int iy = 0; iy < cy; ++iy) for (int ix = 0; ix < cx; ++ix) {
    const size_t i = (size_t)(by + iy) * fd.bxs + bx + ix;
    f->acs[i] = (uint8_t)(s | ((ix == 0 && iy == 0) ? 0x80 : 0));
    f->acs_entropy[i] = (ix == 0 && iy == 0) ? est : 0.0f;
  }
}

// One aligned square of `blocks` x `blocks` (2 or 4): the square transform against the two ways of
// halving it against what is there now (libjxl FindBestFirstLevelDivisionForSquare, simplified to
// aligned candidates).  "Horizontal" halves are wide transforms (top / bottom), "vertical" halves
// are tall ones (left / right).
void MergeSquare(Frame* f, const AcsConfig& cfg, int blocks, int sx, int sy) {
  const FrameDim& fd = f->fd;
  const int half = blocks / 2;
  const int s_wide = blocks == 2 ? DCT8X16 : DCT16X32;    // half rows x full cols
  const int s_tall = blocks == 2 ? DCT16X8 : DCT32X16;    // full rows x half cols
  const int s_sq = blocks == 2 ? DCT16X16 : DCT32X32;
  const float mul_half = blocks == 2 ? 1.25f : 1.5f;
  const float mul_sq = blocks == 2 ? 1.35f : 1.5f;
  auto region = [&](int x0, int y0, int w, int h) {
    float acc = 0.0f;
    for (int y = 0; y < h; ++y) for (int x = 0; x < w; ++x) acc += f->acs_entropy[(size_t)(sy + y0 + y) * fd.bxs + sx + x0 + x];
    return acc;
  };
  float cur_h[2], cur_v[2], e_h[2], e_v[2];
  for (int i = 0; i < 2; ++i) {
    cur_h[i] = region(0, i * half, blocks, half);
    cur_v[i] = region(i * half, 0, half, blocks);
    e_h[i] = EstimateEntropy(*f, cfg, s_wide, mul_half, sx, sy + i * half);
    e_v[i] = EstimateEntropy(*f, cfg, s_tall, mul_half, sx + i * half, sy);
  }
  const float e_s = EstimateEntropy(*f, cfg, s_sq, mul_sq, sx, sy);
  bool take_h[2], take_v[2];
  float cost_h = 0.0f, cost_v = 0.0f;
  for (int i = 0; i < 2; ++i) {
    take_h[i] = !(e_h[i] >= cur_h[i]);   // NaN candidates are accepted (H9 semantics, diff :296)
    take_v[i] = !(e_v[i] >= cur_v[i]);
    cost_h += take_h[i] ? e_h[i] : cur_h[i];
    cost_v += take_v[i] ? e_v[i] : cur_v[i];
  }
  float best = cur_h[0] + cur_h[1];
  int choice = 0;
  if ((take_h[0] || take_h[1]) && !(cost_h >= best)) { best = cost_h; choice = 1; }
  if ((take_v[0] || take_v[1]) && !(cost_v >= best)) { best = cost_v; choice = 2; }
  if (!(e_s >= best)) { best = e_s; choice = 3; }
  if (choice == 1) {
    for (int i = 0; i < 2; ++i) if (take_h[i]) SetStrategy(f, s_wide, sx, sy + i * half, e_h[i]);
  } else if (choice == 2) {
    for (int i = 0; i < 2; ++i) if (take_v[i]) SetStrategy(f, s_tall, sx + i * half, sy, e_v[i]);
  } else if (choice == 3) {
    SetStrategy(f, s_sq, sx, sy, e_s);
  }
}

}  // namespace

void AcStrategySearch(Frame* f) {
  const FrameDim& fd = f->fd;
  const Params& p = f->params;
  if (p.effort < 5) return;   // ProcessRectACS returns early for tiers faster than hare (Appendix U.3)
  AcsConfig cfg;
  const float ratio = (p.distance + 0.1373f) / 1.1373f;
  cfg.info_loss_multiplier = 1.2f * powf(ratio, 0.33677806662454718f);
  cfg.zeros_mul = 9.3089171683409026f * powf(ratio, 0.50990926717963703f);
  cfg.cost_delta = 10.833273317067883f * powf(ratio, 0.36702940662370243f);
  cfg.distance = p.distance;
  cfg.partitioning = p.proposal == 1 || p.proposal == 3;
  cfg.factored_entropy = p.proposal == 2 || p.proposal == 3;
  const float mul8x8 = 1.0f - 0.4f / (p.distance + 1.4f);
  // ---- FindBest8x8Transform for every block
  static const int kCand[4] = {DCT, DCT4X4, DCT4X8, DCT8X4};
  static const float kCandMul[4] = {0.8f, 1.08f, 0.8593f, 0.8593f};
  for (int by = 0; by < fd.bys; ++by) for (int bx = 0; bx < fd.bxs; ++bx) {
    float best = 1e30f;
    int best_tx = DCT;
    for (int i = 0; i < 4; ++i) {
      float mul = kCandMul[i] / 0.8f;
      if (i != 0 && p.distance > 4.0f) mul += 0.5f;     // kAvoidEntropyOfTransforms (diff context :260)
      const float e = EstimateEntropy(*f, cfg, kCand[i], mul, bx, by);
      if (e < best) { best = e; best_tx = kCand[i]; }
    }
    if (cfg.partitioning && best_tx == DCT) {
      const float* r = &f->homog[((size_t)by * fd.bxs + bx) * 3];
      best_tx = HomogeneityPartition(r[0], r[1], r[2], p.distance);
    }
    SetStrategy(f, best_tx, bx, by, best * mul8x8);
  }
  // ---- merges: aligned 16x16 squares, then aligned 32x32 squares
  for (int sy = 0; sy + 2 <= fd.bys; sy += 2) for (int sx = 0; sx + 2 <= fd.bxs; sx += 2) MergeSquare(f, cfg, 2, sx, sy);
  for (int sy = 0; sy + 4 <= fd.bys; sy += 4) for (int sx = 0; sx + 4 <= fd.bxs; sx += 4) {
    // a half of the square can only be replaced when no existing transform straddles it: after the
    // 16-level every transform lies inside one 16x16 square, hence inside one half
    MergeSquare(f, cfg, 4, sx, sy);
  }
}

}  // namespace jxo
