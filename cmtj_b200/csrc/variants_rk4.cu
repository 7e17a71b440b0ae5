// RK4 variants (Layer::rk4_step, junction.hpp:1017-1041).
#include "variants.h"
namespace cmtjb {
static const Variant kTable[] = {
    // lean: folded K1+demag diagonal operator + constant fields
    CMTJB_FIXED(1, CMTJB_RK4, 0),
    CMTJB_FIXED(1, CMTJB_RK4, F_DYN),
    CMTJB_FIXED(1, CMTJB_RK4, F_FAST),           // one field driver per layer with per-instance waveform (VSD map)
    CMTJB_FIXED(1, CMTJB_RK4, F_STT),
    CMTJB_FIXED(1, CMTJB_RK4, F_STT | F_DYNS),   // current-driven switching: time-dependent current, batch-uniform waveform
    CMTJB_FIXED(1, CMTJB_RK4, F_SOT | F_DYNS),
    CMTJB_FIXED(1, CMTJB_RK4, F_NOFOLD | F_DYN | F_DYNS | F_K2 | F_SOT | F_STT | F_DEMAG_FULL),
    CMTJB_FIXED(1, CMTJB_RK4, F_ALL_TAB),
    CMTJB_FIXED(1, CMTJB_RK4, F_ALL),
    CMTJB_FIXED(2, CMTJB_RK4, F_IEC),
    CMTJB_FIXED(2, CMTJB_RK4, F_IEC | F_DYN),
    CMTJB_FIXED(2, CMTJB_RK4, F_IEC | F_FAST),
    CMTJB_FIXED(2, CMTJB_RK4, F_IEC | F_STT),            // two-layer STT junction (free + reference layer)
    CMTJB_FIXED(2, CMTJB_RK4, F_IEC | F_STT | F_DYNS),
    CMTJB_FIXED(2, CMTJB_RK4, F_IEC | F_DYN | F_DYNS | F_NOFOLD),
    CMTJB_FIXED(2, CMTJB_RK4, F_ALL_TAB | CMTJB_SPLIT(1)),
    CMTJB_FIXED(2, CMTJB_RK4, F_ALL_TAB),
    CMTJB_FIXED(2, CMTJB_RK4, F_ALL),
    CMTJB_FIXED(3, CMTJB_RK4, F_IEC),
    CMTJB_FIXED(3, CMTJB_RK4, F_IEC | F_DYN),
    CMTJB_FIXED(3, CMTJB_RK4, F_ALL_TAB | CMTJB_SPLIT(1)),  // torque / drivers / overrides on the free layer only
    CMTJB_FIXED(3, CMTJB_RK4, F_ALL_TAB),
    CMTJB_FIXED(3, CMTJB_RK4, F_ALL),
    CMTJB_FIXED(4, CMTJB_RK4, F_ALL_TAB),
    CMTJB_FIXED(4, CMTJB_RK4, F_ALL),
};
int variants_rk4(const Variant **out) {
  *out = kTable;
  return (int)(sizeof(kTable) / sizeof(kTable[0]));
}
}  // namespace cmtjb
