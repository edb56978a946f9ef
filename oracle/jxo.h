// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product.
//
// Scalar CPU restatement of the JPEG XL lossy (VarDCT) encode hot path that the
// thesis workspace pscoro/JPEG-XL-Lossy-Image-Compression-Thesis benchmarks
// (benchmark-jpegxl/src/docker_manager.rs:100-137 -> `cjxl in out --distance=D
// --effort=E`) and patches (proposals/*.diff).  Only tests/, bench.py's
// cpu_baseline / --impl reference legs and __graft_entry__.smoke() may use it.
//
// PARITY STATUS
//  * H-rows (homogeneity metric + hooks): restated line by line from
//    proposals/homogeneity-partitioning.diff:17-235, :272-276 and
//    proposals/homogeneity-factored-entropy.diff:248-253.  Pinned by the
//    hand-computed vectors in tests/golden/ (the reference ships none).
//  * U-rows (XYB, adaptive quant, transforms, quantisation, tokens, histograms,
//    ANS, frame assembly): the arithmetic lives in libjxl, which is NOT under
//    /root/reference (cloned unpinned at docker build time,
//    benchmark-jpegxl/Dockerfile:40, default "main" benchmark-jpegxl/src/context.rs:17;
//    best identification v0.10.x from the blob ids in proposals/*.diff:2).  These
//    stages restate libjxl's published algorithm / ISO 18181-1 from recall:
//    **parity unpinned** — no golden vector of the reference exists for them.
//
// Numerics contract shared with the CUDA path (DESIGN.md "Numerics"): IEEE fp32,
// round-to-nearest, no contraction (-ffp-contract=off / -fmad=false); a fused
// multiply-add happens exactly where fmaf() is written.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <cmath>
#include <vector>
#include <string>
#include <algorithm>

namespace jxo {

// ---------------------------------------------------------------- geometry
constexpr int kBlockDim = 8;
constexpr int kDCTBlockSize = 64;
constexpr int kGroupDim = 256;          // AC group, pixels
constexpr int kGroupDimInBlocks = 32;
constexpr int kDcGroupDimInBlocks = 256;  // 2048 px
constexpr int kColorTileDimInBlocks = 8;  // 64 px (ACS / CfL tile)

struct FrameDim {
  int xsize = 0, ysize = 0;        // original pixels
  int xs_pad = 0, ys_pad = 0;      // padded to x8
  int pitch = 0;                   // floats per plane row (xs_pad rounded up to 32)
  int bxs = 0, bys = 0;            // blocks
  int gxs = 0, gys = 0, num_groups = 0;        // AC groups
  int dgxs = 0, dgys = 0, num_dc_groups = 0;   // DC groups
  int txs = 0, tys = 0;            // 64x64 tiles
  void Set(int w, int h) {
    xsize = w; ysize = h;
    xs_pad = (w + 7) / 8 * 8; ys_pad = (h + 7) / 8 * 8;
    pitch = (xs_pad + 31) / 32 * 32;
    bxs = xs_pad / 8; bys = ys_pad / 8;
    gxs = (bxs + 31) / 32; gys = (bys + 31) / 32; num_groups = gxs * gys;
    dgxs = (bxs + 255) / 256; dgys = (bys + 255) / 256; num_dc_groups = dgxs * dgys;
    txs = (bxs + 7) / 8; tys = (bys + 7) / 8;
  }
};

// ---------------------------------------------------------------- bit helpers
static inline uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static inline int FloorLog2(uint32_t v) { return 31 - __builtin_clz(v); }          // v > 0
static inline int CeilLog2(uint32_t v) { return v <= 1 ? 0 : FloorLog2(v - 1) + 1; }  // v > 0
static inline uint32_t PackSigned(int32_t v) { return v >= 0 ? 2u * (uint32_t)v : 2u * (uint32_t)(-v) - 1u; }
static inline int32_t UnpackSigned(uint32_t u) { return (u & 1) ? -(int32_t)((u + 1) >> 1) : (int32_t)(u >> 1); }

// ---------------------------------------------------------------- AcStrategy
enum Strategy : uint8_t {
  DCT = 0, IDENTITY = 1, DCT2X2 = 2, DCT4X4 = 3, DCT16X16 = 4, DCT32X32 = 5,
  DCT16X8 = 6, DCT8X16 = 7, DCT32X8 = 8, DCT8X32 = 9, DCT32X16 = 10, DCT16X32 = 11,
  DCT4X8 = 12, DCT8X4 = 13, AFV0 = 14, AFV1 = 15, AFV2 = 16, AFV3 = 17,
  DCT64X64 = 18, DCT64X32 = 19, DCT32X64 = 20, kNumStrategies = 27
};
extern const uint8_t kCoveredX[27];   // covered blocks, horizontal
extern const uint8_t kCoveredY[27];   // covered blocks, vertical
extern const uint8_t kStrategyOrder[27];
// quant-table kind per strategy (libjxl DequantMatrices::kQuantTable)
extern const uint8_t kQuantKind[27];

// ---------------------------------------------------------------- stage: XYB (U1)
void SrgbLut(float lut[256]);
float CbrtPos(float x);
// rgb: h*w*3 interleaved, row stride in bytes.  Output: 3 planes of ys_pad*pitch
// floats (edge-replicated to the padded size, pitch padding zero-filled).
void RgbToXyb(const uint8_t* rgb, int w, int h, size_t stride, const FrameDim& fd, float* x, float* y, float* b);
// Gaborish (jxo_xyb.cc): the encoder's sharpening of the padded XYB planes, the decoder's blur of one pixel
void GaborishInverse(const FrameDim& fd, float* planes[3]);
float GaborishBlurAt(const float* plane, int pitch, int xsize, int ysize, int x, int y);

// ---------------------------------------------------------------- stage: transforms (U5)
// 1-D scaled DCT-II / its inverse over `n` floats with stride; n in {2,4,8,16,32,64}
void Dct1D(const float* in, int in_stride, float* out, int out_stride, int n);
void Idct1D(const float* in, int in_stride, float* out, int out_stride, int n);
// 2-D transform of a rows x cols pixel rectangle; output has the long side
// horizontal: rows>=cols -> out[hf*rows + vf] (transposed), else out[vf*cols + hf].
void Dct2D(const float* px, int px_stride, int rows, int cols, float* out);
void Idct2D(const float* coef, int rows, int cols, float* px, int px_stride);
// Full per-strategy forward/inverse transform on a covered rectangle
void TransformFromPixels(int strategy, const float* px, int px_stride, float* coef);
void TransformToPixels(int strategy, const float* coef, float* px, int px_stride);
void DcFromLowestFrequencies(int strategy, const float* coef, float* dc, int dc_stride);
void LowestFrequenciesFromDc(int strategy, const float* dc, int dc_stride, float* llf);

// ---------------------------------------------------------------- stage: quant tables (U5)
// weights (== inverse dequant matrix) for table `kind` (0..10 covered), 3 channels;
// returns number of coefficients per channel
int QuantWeights(int kind, std::vector<float>* w);
float FastLog2f(float x);
float FastPow2f(float x);
float FastPowf(float base, float e);   // FastPow2f(FastLog2f(base) * e)

// ---------------------------------------------------------------- stage: homogeneity (H1-H9)
struct HomogConfig {
  const float* rows[3];   // X, Y, B plane base (rect origin)
  size_t stride;          // floats per row (pitch) — H1 quirk: used as the horizontal bound
  size_t ysize;           // rect_in.ysize()
};
size_t CalculateNumZeroCrossings(size_t xsize, size_t ysize, float threshold, const float* laplacian);
void CalculateLaplacianFilter(size_t x, size_t y, size_t xsize, size_t ysize, size_t bx, size_t by, const HomogConfig& c, float* out);
float CalculateSumModifiedLaplacian(size_t x, size_t y, size_t xsize, size_t ysize, size_t bx, size_t by, const HomogConfig& c);
float CalculateColorfulness(size_t x, size_t y, size_t xsize, size_t ysize, size_t bx, size_t by, const HomogConfig& c);
float CalculateHomogeneity(size_t x, size_t y, size_t xsize, size_t ysize, size_t bx, size_t by, float d, const HomogConfig& c);
void CalculateHomogeneitySimilarityIndices(size_t x, size_t y, float d, const HomogConfig& c, float* r_h, float* r_v, float* r_d);
uint8_t HomogeneityPartition(float r_h, float r_v, float r_d, float d);

}  // namespace jxo
