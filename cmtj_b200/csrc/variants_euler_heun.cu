// Euler-Heun variants (Junction::eulerHeunSolverStep, junction.hpp:1709-1754).
#include "variants.h"
namespace cmtjb {
static const Variant kTable[] = {
    CMTJB_FIXED(1, CMTJB_EULER_HEUN, 0),
    CMTJB_FIXED(1, CMTJB_EULER_HEUN, F_STT | F_THERMAL),
    CMTJB_FIXED(1, CMTJB_EULER_HEUN, F_STT | F_THERMAL | F_OVR),
    CMTJB_FIXED(1, CMTJB_EULER_HEUN, F_ONEF),                       // 1/f^alpha noise alone (sensor noise floors)
    CMTJB_FIXED(1, CMTJB_EULER_HEUN, F_STT | F_THERMAL | F_ONEF),   // ... on top of thermal noise
    CMTJB_FIXED(1, CMTJB_EULER_HEUN, F_ALL_TAB),
    CMTJB_FIXED(1, CMTJB_EULER_HEUN, F_ALL),
    CMTJB_FIXED(2, CMTJB_EULER_HEUN, F_IEC),
    CMTJB_FIXED(2, CMTJB_EULER_HEUN, F_IEC | F_THERMAL),          // two coupled layers at temperature
    CMTJB_FIXED(2, CMTJB_EULER_HEUN, F_IEC | F_STT | F_THERMAL),  // thermally assisted STT switching of a two-layer junction
    CMTJB_FIXED(2, CMTJB_EULER_HEUN, F_IEC | F_THERMAL | F_ONEF),
    CMTJB_FIXED(2, CMTJB_EULER_HEUN, F_ALL_TAB),
    CMTJB_FIXED(2, CMTJB_EULER_HEUN, F_ALL),
    CMTJB_FIXED(3, CMTJB_EULER_HEUN, F_IEC),
    CMTJB_FIXED(3, CMTJB_EULER_HEUN, F_ALL_TAB),
    CMTJB_FIXED(3, CMTJB_EULER_HEUN, F_ALL),
    CMTJB_FIXED(4, CMTJB_EULER_HEUN, F_ALL_TAB),
    CMTJB_FIXED(4, CMTJB_EULER_HEUN, F_ALL),
};
int variants_euler_heun(const Variant **out) {
  *out = kTable;
  return (int)(sizeof(kTable) / sizeof(kTable[0]));
}
}  // namespace cmtjb
