// ORACLE (test infrastructure) — encoder-side state shared by the oracle stages.
// See jxo.h for the parity status of each stage.
#pragma once
#include "jxo.h"

namespace jxo {

struct QuantState {
  int global_scale = 1;
  int quant_dc = 1;
  float scale = 0;             // global_scale / 65536
  float inv_global_scale = 0;  // 65536 / global_scale
  float x_qm_mul = 1.0f, b_qm_mul = 1.0f;  // 1.25^(x_qm_scale-2), 1.25^(b_qm_scale-2)
  int x_qm_scale = 2, b_qm_scale = 2;
  bool adjust_quant = true;    // effort >= 5 (speed tier <= hare): AdjustQuantBlockAC
  float median = 0, mad = 0;
};

struct EncTables {
  std::vector<float> weights[17];   // inverse dequant matrices (quantisation weights), 3 channels each
  std::vector<float> dequant[17];   // 1 / weights
  int ncoef[17] = {0};
  std::vector<uint16_t> order[13];  // natural coefficient order per strategy order class (position list)
  void Init();
};

void ComputeGlobalScale(const float* qf, size_t n, float quant_dc, QuantState* q);
void SetRawQuantField(const float* qf, size_t n, const QuantState& q, int32_t* raw);
void ComputeCoefficientsBlock(const EncTables& T, const QuantState& q, int strategy,
                              const float* px[3], int ps, float x_factor, float b_factor,
                              int32_t* quant_io, float* dc_out[3], int dc_stride, int32_t* out[3]);
void QuantizeDc(const QuantState& q, const float* dc[3], size_t n, int32_t* out[3]);

// natural (zig-zag) coefficient order of a strategy: order[k] = position in the coefficient block
void NaturalCoeffOrder(int strategy, std::vector<uint16_t>* order);

float InitialQuantDC(float distance);

}  // namespace jxo
