// Dormand-Prince variants (Layer::dormand_prince_step, junction.hpp:866-1002); single layer.
// Adaptive time: dynamic drivers are always evaluated inline (F_DYNI).
#include "variants.h"
namespace cmtjb {
static const Variant kTable[] = {
    CMTJB_DOPRI(0),
    CMTJB_DOPRI(F_SOT | F_K2 | F_DYNI),
    CMTJB_DOPRI(F_ALL),
};
int variants_dopri(const Variant **out) {
  *out = kTable;
  return (int)(sizeof(kTable) / sizeof(kTable[0]));
}
}  // namespace cmtjb
