// variants.h -- helper macros for the per-solver variant tables (variants_*.cu).
// A variant is one fully specialised instantiation of the fused time-loop kernel:
// (layer count, solver, compile-time feature mask).  cabi.cu picks, per run, the compiled
// variant with the fewest superfluous terms that still covers the model.
#pragma once
#include "llgs_kernels.cuh"

#define CMTJB_FIXED(NL, SOLVER, FEAT) \
  { NL, SOLVER, (uint32_t)(FEAT), &cmtjb::launch_fixed<NL, SOLVER, (uint32_t)(FEAT)>, #SOLVER "/L" #NL "/" #FEAT }
#define CMTJB_DOPRI(FEAT) \
  { 1, CMTJB_DORMAND_PRINCE, (uint32_t)(FEAT), &cmtjb::launch_dopri<(uint32_t)(FEAT)>, "CMTJB_DORMAND_PRINCE/L1/" #FEAT }
