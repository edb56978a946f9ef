// jxlb200 — host-side constant tables (see host_tables.cu)
#pragma once
#include <stdint.h>
#include <vector>

namespace jxlb {
extern const uint8_t kCoveredX[27];
extern const uint8_t kCoveredY[27];
extern const uint8_t kStrategyOrder[27];
extern const uint8_t kQuantKind[27];
void host_srgb_lut(float lut[256]);
void host_recon_tables(float tab[264]);
int host_quant_weights(int kind, std::vector<float>* w);
void host_natural_order(int strategy, std::vector<uint16_t>* order);
float host_initial_quant_dc(float distance);
}  // namespace jxlb
