// ORACLE (test infrastructure) — stage entry points used by the frame driver.
#pragma once
#include "jxo_frame.h"

namespace jxo {
const EncTables& GetTables();
void InitialQuantField(Frame* f);       // U2  (jxo_aq.cc)
void AdjustQuantField(Frame* f);        // U2  (jxo_aq.cc)
void ChromaFromLumaFit(Frame* f);      // U3 (opt-in): per-tile ytox / ytob (jxo_acs.cc)
void AcStrategySearch(Frame* f);        // U4 + H8/H9 (jxo_acs.cc)
bool EntropyCodeFrame(Frame* f);        // U6-U9 (jxo_entropy.cc, jxo_modular.cc, jxo_bitstream.cc)
bool DecodeCodestream(const uint8_t* data, size_t size, Frame* f);  // self-decoder (jxo_decode.cc)
bool ReconstructRgb(const Frame& f, uint8_t* rgb);                  // decoded integers -> sRGB pixels (jxo_recon.cc)
bool ReconstructionSse(const Frame& f, const uint8_t* orig, size_t stride, uint64_t sse[3]);
void SrgbBoundaries(float b[255]);
void InverseOpsin(float inv[9]);
}  // namespace jxo
