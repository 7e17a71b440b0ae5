// Dormand-Prince variants (Layer::dormand_prince_step, junction.hpp:866-1002); single layer.
#include "variants.h"
namespace cmtjb {
static const Variant kTable[] = {
    CMTJB_DOPRI(0),
    CMTJB_DOPRI(F_ALL),
};
int variants_dopri(const Variant **out) {
  *out = kTable;
  return (int)(sizeof(kTable) / sizeof(kTable[0]));
}
}  // namespace cmtjb
