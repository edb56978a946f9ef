// ORACLE (test infrastructure) — stage U4: AC-strategy search (placeholder until slice 4).
#include "jxo_frame.h"
#include "jxo_stages.h"
namespace jxo {
void AcStrategySearch(Frame* f) { (void)f; }
}  // namespace jxo
