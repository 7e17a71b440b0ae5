// RK4 variants carrying the fused post-processing of the sweep procedures (F_POST: projection log
// and spin-diode IIR of cmtjb_post_t; cmtj/utils/procedures.py:42-331 always runs the default
// solver).  Kept in their own translation unit so the plain variants stay untouched.
#include "variants.h"
namespace cmtjb {
static const Variant kTable[] = {
    CMTJB_FIXED(1, CMTJB_RK4, F_PROJ),            // projection log only (PIMM): relaxation segment, no spin-diode state
    CMTJB_FIXED(2, CMTJB_RK4, F_IEC | F_PROJ),    // ... of the two-layer MTJ (the S1 sweep through PIMM_batch)
    CMTJB_FIXED(1, CMTJB_RK4, F_POST),            // the relaxation segment after the pulse window (constant drivers)
    CMTJB_FIXED(1, CMTJB_RK4, F_DYN | F_POST),    // PIMM / VSD of a single layer, batch-uniform excitation
    CMTJB_FIXED(1, CMTJB_RK4, F_FAST | F_POST),           // VSD with a per-instance excitation frequency (resolved once)
    CMTJB_FIXED(1, CMTJB_RK4, F_DYNI | F_OVR | F_POST),   // ... the general inline-evaluating form
    CMTJB_FIXED(1, CMTJB_RK4, F_ALL | F_POST),
    CMTJB_FIXED(2, CMTJB_RK4, F_IEC | F_POST),           // relaxation segment of the two-layer PIMM sweep
    CMTJB_FIXED(2, CMTJB_RK4, F_IEC | F_DYN | F_POST),   // PIMM of the two-layer MTJ (bench S1)
    CMTJB_FIXED(2, CMTJB_RK4, F_IEC | F_FAST | F_POST),
    CMTJB_FIXED(2, CMTJB_RK4, F_IEC | F_DYNI | F_OVR | F_POST),
    CMTJB_FIXED(2, CMTJB_RK4, F_ALL | F_POST),
    CMTJB_FIXED(3, CMTJB_RK4, F_ALL_TAB | F_POST),
    CMTJB_FIXED(4, CMTJB_RK4, F_ALL_TAB | F_POST),
};
int variants_post(const Variant **out) {
  *out = kTable;
  return (int)(sizeof(kTable) / sizeof(kTable[0]));
}
}  // namespace cmtjb
