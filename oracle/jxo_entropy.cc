// ORACLE (test infrastructure) — stages U6-U9 (placeholder until slice 2).
#include "jxo_frame.h"
#include "jxo_stages.h"
namespace jxo {
bool EntropyCodeFrame(Frame* f) { (void)f; return true; }
}  // namespace jxo
