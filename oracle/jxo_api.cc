// ORACLE (test infrastructure) — C entry points for ctypes (tests/, bench.py cpu_baseline).
#include "jxo_frame.h"
#include "jxo_stages.h"

using namespace jxo;

template <typename T>
static size_t CopyOut(const std::vector<T>& v, void* dst, size_t cap) {
  const size_t bytes = v.size() * sizeof(T);
  if (dst && cap >= bytes && bytes) memcpy(dst, v.data(), bytes);
  return bytes;
}

extern "C" {

struct JxoParams { float distance; uint32_t effort; uint32_t proposal; uint32_t flags; };

void* jxo_encode(const uint8_t* rgb, int w, int h, size_t stride, const JxoParams* p) {
  Frame* f = new Frame();
  Params pp; pp.distance = p->distance; pp.effort = p->effort; pp.proposal = p->proposal; pp.flags = p->flags;
  EncodeFrame(rgb, w, h, stride, pp, f);
  return f;
}
// encode with a caller-supplied strategy map (bys * bxs bytes, raw strategy | 0x80 on first blocks) instead of the search
void* jxo_encode_forced(const uint8_t* rgb, int w, int h, size_t stride, const JxoParams* p, const uint8_t* acs, size_t n) {
  Frame* f = new Frame();
  Params pp; pp.distance = p->distance; pp.effort = p->effort; pp.proposal = p->proposal; pp.flags = p->flags | kFlagForcedAcs;
  f->forced_acs.assign(acs, acs + n);
  EncodeFrame(rgb, w, h, stride, pp, f);
  return f;
}
// self-decoder: parses a codestream back into the integer stages (dc_quant, acs, raw_qf, coeffs, ...)
void* jxo_decode(const uint8_t* data, size_t size) {
  Frame* f = new Frame();
  DecodeCodestream(data, size, f);
  return f;
}
// reconstructs h*w*3 sRGB bytes from a decoded frame (jxo_decode); returns 1 on success
int jxo_reconstruct(void* h, uint8_t* rgb) { return ReconstructRgb(*(Frame*)h, rgb) ? 1 : 0; }
// per-channel sum of squared errors of the frame's reconstruction against the original image
int jxo_sse(void* h, const uint8_t* orig, size_t stride, uint64_t* sse3) { return ReconstructionSse(*(Frame*)h, orig, stride, sse3) ? 1 : 0; }
const char* jxo_error(void* h) { return ((Frame*)h)->error.c_str(); }
void jxo_free(void* h) { delete (Frame*)h; }

// fills dims[16]: xsize ysize xs_pad ys_pad pitch bxs bys gxs gys num_groups dgxs dgys num_dc_groups txs tys 0
void jxo_dims(int w, int h, int32_t* dims) {
  FrameDim fd; fd.Set(w, h);
  const int32_t v[16] = {fd.xsize, fd.ysize, fd.xs_pad, fd.ys_pad, fd.pitch, fd.bxs, fd.bys, fd.gxs, fd.gys,
                         fd.num_groups, fd.dgxs, fd.dgys, fd.num_dc_groups, fd.txs, fd.tys, 0};
  memcpy(dims, v, sizeof(v));
}

// returns the stage's size in bytes; copies when dst has room
size_t jxo_dump(void* h, int stage, void* dst, size_t cap) {
  Frame* f = (Frame*)h;
  switch (stage) {
    case kStageXyb: {
      const size_t pb = f->xyb[0].size() * 4;
      if (dst && cap >= 3 * pb) for (int c = 0; c < 3; ++c) memcpy((char*)dst + c * pb, f->xyb[c].data(), pb);
      return 3 * pb;
    }
    case kStageQfFloat: return CopyOut(f->qf_float, dst, cap);
    case kStageMask1x1: return CopyOut(f->mask1x1, dst, cap);
    case kStageMask: return CopyOut(f->mask, dst, cap);
    case kStageHomog: return CopyOut(f->homog, dst, cap);
    case kStageAcs: return CopyOut(f->acs, dst, cap);
    case kStageAcsEntropy: return CopyOut(f->acs_entropy, dst, cap);
    case kStageRawQf: return CopyOut(f->raw_qf, dst, cap);
    case kStageQuantParams: {
      const int32_t v[4] = {f->q.global_scale, f->q.quant_dc, f->q.x_qm_scale, f->q.b_qm_scale};
      if (dst && cap >= sizeof(v)) memcpy(dst, v, sizeof(v));
      return sizeof(v);
    }
    case kStageCoeffs: return CopyOut(f->coeffs, dst, cap);
    case kStageDcQuant: return CopyOut(f->dc_quant, dst, cap);
    case kStageNzeros: return CopyOut(f->nzeros, dst, cap);
    case kStageCmap: return CopyOut(f->cmap, dst, cap);
    case kStageTokenOffsets: return CopyOut(f->token_offsets, dst, cap);
    case kStageTokens: return CopyOut(f->tokens, dst, cap);
    case kStageHistograms: return CopyOut(f->histograms, dst, cap);
    case kStageContextMap: return CopyOut(f->context_map, dst, cap);
    case kStageGroupOffsets: return CopyOut(f->group_offsets, dst, cap);
    case kStageGroupStreams: return CopyOut(f->group_streams, dst, cap);
    case kStageCodestream: return CopyOut(f->codestream, dst, cap);
    case 21: { const int32_t v[1] = {f->num_clusters}; if (dst && cap >= 4) memcpy(dst, v, 4); return 4; }
    default: return 0;
  }
}

// ---- direct taps for the hand-computed golden vectors (H rows) and unit tests -----
void jxo_homogeneity_indices(const float* x, const float* y, const float* b, size_t stride, size_t ysize,
                             size_t px, size_t py, float d, float* out3) {
  HomogConfig hc; hc.rows[0] = x; hc.rows[1] = y; hc.rows[2] = b; hc.stride = stride; hc.ysize = ysize;
  CalculateHomogeneitySimilarityIndices(px, py, d, hc, out3, out3 + 1, out3 + 2);
}
float jxo_homogeneity(const float* x, const float* y, const float* b, size_t stride, size_t ysize,
                      size_t px, size_t py, size_t xs, size_t ys, size_t bx, size_t by, float d) {
  HomogConfig hc; hc.rows[0] = x; hc.rows[1] = y; hc.rows[2] = b; hc.stride = stride; hc.ysize = ysize;
  return CalculateHomogeneity(px, py, xs, ys, bx, by, d, hc);
}
int jxo_homogeneity_partition(float r_h, float r_v, float r_d, float d) { return HomogeneityPartition(r_h, r_v, r_d, d); }
// H9: ret = ret * 0.8 * avg_r  (proposals/homogeneity-factored-entropy.diff:248-253)
float jxo_factored_entropy(float ret, float r_h, float r_v, float r_d) {
  const float avg_r = (r_h + r_v + r_d) / 3;
  return (float)(((double)ret * 0.8) * (double)avg_r);
}
void jxo_dct2d(const float* px, int stride, int rows, int cols, float* out) { Dct2D(px, stride, rows, cols, out); }
void jxo_idct2d(const float* coef, int rows, int cols, float* px, int stride) { Idct2D(coef, rows, cols, px, stride); }
void jxo_transform(int strategy, const float* px, int stride, float* coef) { TransformFromPixels(strategy, px, stride, coef); }
void jxo_inverse_transform(int strategy, const float* coef, float* px, int stride) { TransformToPixels(strategy, coef, px, stride); }
int jxo_quant_weights(int kind, float* out, size_t cap) {
  std::vector<float> w; int n = QuantWeights(kind, &w);
  if (out && cap >= w.size()) memcpy(out, w.data(), w.size() * 4);
  return n;
}
int jxo_natural_order(int strategy, uint16_t* out, size_t cap) {
  std::vector<uint16_t> o; NaturalCoeffOrder(strategy, &o);
  if (out && cap >= o.size()) memcpy(out, o.data(), o.size() * 2);
  return (int)o.size();
}
float jxo_cbrt(float x) { return CbrtPos(x); }
void jxo_srgb_lut(float* lut) { SrgbLut(lut); }

}  // extern "C"
