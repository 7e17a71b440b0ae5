// dev_model.h -- derived, device-resident form of a junction model.
//
// The raw records of include/cmtj_b200.h are turned into this struct on the host
// (cabi.cu: derive_model) with every reciprocal and product that does not depend on the
// magnetisation hoisted (SURVEY 7 "Target arithmetic").  The whole struct is passed to the
// kernels BY VALUE as a __grid_constant__ parameter, so uniform parameters live in
// constant bank 0 and reach the FP64 pipe as free c[0x0][..] operands: no registers, no
// loads.  Only per-instance quantities (state, summed constant field, override table
// entries) occupy registers.
#pragma once
#include <stdint.h>

#include "../../include/cmtj_b200.h"

namespace cmtjb {

// compile-time feature bits of a kernel variant (terms that are compiled in)
enum : uint32_t {
  F_K2 = 1u << 0,         // second-order anisotropy          (junction.hpp:645-652)
  F_IEC = 1u << 1,        // bilinear + biquadratic IEC       (junction.hpp:654-675)
  F_IDMI = 1u << 2,       // interlayer DMI                   (junction.hpp:677-701)
  F_DIPOLE = 1u << 3,     // top/bottom dipole tensors        (junction.hpp:583-585)
  F_DEMAG_FULL = 1u << 4, // off-diagonal demag entries (diagonal is always compiled in)
  F_STT = 1u << 5,        // Slonczewski STT                  (junction.hpp:743-769)
  F_SOT = 1u << 6,        // SOT primary + secondary          (junction.hpp:770-812)
  F_THERMAL = 1u << 7,    // Langevin field                   (junction.hpp:1111-1128)
  F_ONEF = 1u << 8,       // pre-generated 1/f^alpha vector   (junction.hpp:1130-1136)
  F_DYN = 1u << 9,        // time-dependent FIELD drivers (Hext/Hoe/Hdmi/Hres): per-layer sum read from a
                          // per-launch table (uniform in the batch)
  F_OVR = 1u << 10,       // per-instance overrides other than constant field components
  F_DYNI = 1u << 11,      // time-dependent drivers evaluated inline per thread (per-instance
                          // frequency/phase/timing overrides, adaptive time); wins over F_DYN
  F_NOFOLD = 1u << 12,    // K1 anisotropy and demag kept as separate terms; without this bit the
                          // kernel applies the host-folded DIAGONAL operator lin = (2K1/Ms) a a^T - N Ms/mu0
  F_DYNS = 1u << 13,      // time-dependent SCALAR drivers (K, J, I, torques, T) read from the same table
  F_ALL = (1u << 14) - 1,
  F_ALL_TAB = F_ALL & ~F_DYNI,
  // fused post-processing of the sweep procedures (projection log, spin-diode IIR; cmtjb_post_t).
  // Not part of F_ALL: only the variants of variants_post.cu carry it.
  F_POST = 1u << 14,
  // Split variants: only the layers named in the 4-bit heavy set (bits 15..18, bit l = layer l) carry the
  // full term set; the others are compiled with the field terms only (F_LIGHTSET: constant, batch-uniform
  // parameters, no torque, no noise, no time-dependent driver, no override).  A CIMS stack has its SOT
  // drive, per-instance amplitudes and step drivers on the free layer alone: the two pinned layers then
  // cost what they cost in the lean sweep kernel, and the loop body shrinks by a third.
  F_SPLIT = 1u << 19,
  F_HEAVY_SHIFT = 15,
  F_LIGHTSET = F_K2 | F_IEC | F_IDMI | F_DIPOLE | F_DEMAG_FULL | F_NOFOLD,
  F_BASEMASK = (1u << 15) - 1,
  // Fast-field variants (RK4): every time-dependent driver is a FIELD driver, at most one per layer, and every
  // per-instance waveform parameter belongs to it (a VSD map: Oersted sine with per-instance frequency).  The
  // driver is resolved once per launch (resolve_fast) and added to the field at the three stage times, so
  // the loop body is compiled WITHOUT inline-driver and override machinery (F_DYNI / F_OVR).
  F_FAST = 1u << 20,
  // Projection-only post-processing: the projection log of F_POST without the spin-diode filter and its per-thread
  // state (a PIMM sweep writes one projected double per step and nothing else: 87 registers like the plain RK4 kernel
  // instead of the 108 of the F_POST form, whose double-precision sine also brings a stack frame).
  F_PROJ = 1u << 21
};
#define CMTJB_SPLIT(heavy_layers) (cmtjb::F_SPLIT | ((uint32_t)(heavy_layers) << cmtjb::F_HEAVY_SHIFT))
// the term set compiled into layer l of a variant
#ifdef __CUDACC__
__host__ __device__
#endif
constexpr uint32_t layer_mask(uint32_t F, int l) {
  return ((F & F_SPLIT) != 0 && ((F >> F_HEAVY_SHIFT >> l) & 1u) == 0) ? (F & F_LIGHTSET) : (F & F_BASEMASK);
}

struct DevDriver {
  int32_t type;  // cmtjb_update
  int32_t pad;   // CMTJB_UPD_CUSTOM: column of RunParams::custom
  double amplitude, frequency, phase, period, cycle, timeStart, timeStop, edgeTime, steadyTime;
};

struct DevLayer {
  double cst[CMTJB_NSLOTS];  // constantValue of every slot
  double anis[3];
  double cK1, cK2;           // 2/Ms, 4/Ms
  double inv_Mst;            // 1/(Ms*thickness)
  double demag[9];           // demagTensor * (Ms/mu0)
  double dipB[9], dipT[9];   // dipole tensors * (Ms/mu0)
  double alpha;
  double llg;                // -gamma/(1+alpha^2)
  double aJ_c;               // hbar/(e*Ms*thickness)
  double lam, lam2, P, beta, kappa;
  double stt_num, stt_a, stt_b;  // Slonczewski eta = stt_num / (stt_a (m.p) + stt_b): P lam^2, lam^2 - 1, lam^2 + 1 (or P, lam, 1)
  double sigma0;                 // sqrt(sigma_c * T) of a constant, batch-uniform temperature
  double flT, dlT;           // static SOT efficiencies
  double ref[3], ref2[3];
  double sigma_c;            // 2*alpha*kB/(Ms*V*gamma); sigma = sqrt(sigma_c*T)
  double Rx0, Ry0, AMR_X, SMR_X, AHE_half, AMRSMR_Y;
  // Dormand-Prince controller
  double abs_tol, rel_tol, max_factor, min_factor, safety_factor, ki, kp, kd, prev_error_ratio,
      integral_error;
  int32_t refType;           // cmtjb_reference
  int32_t stt, sot, dynamicSOT, altSTT, thermal, onef, pid;
  int32_t includeSOT;        // raw flag: the energy log writes Hfl/Hdl for these layers (junction.hpp:1604-1609)
  double lin[3];             // folded diagonal linear operator (valid when the variant lacks F_NOFOLD)
  // 1/f^alpha noise generated on the device (onef.cu): first series of this layer in RunParams::onef_gen, the
  // VectorAlphaNoise axis (cmtjb_axis; ALL = three components, normalised per sample) and scale (noise.hpp:291-301)
  int32_t onef_slot, onef_axis;
  double onef_scale;
  // per-launch products (run_typed): h = llg*dt and the IEC coefficients J/(Ms t), 2 J2/(Ms t) of
  // constant, batch-uniform couplings (used by variants without F_DYNS / F_DYNI / F_OVR)
  double h, h_half, h_sixth;
  double iecB_a, iecB_b, iecT_a, iecT_b;
  uint32_t dynmask;          // bit s: slot s is time dependent
  int8_t dyncol[CMTJB_NSLOTS];     // column of scalar slot s in the per-launch driver table, or -1
  int32_t dynH;              // 1: the layer has time-dependent field drivers (table columns 3l..3l+2)
  uint32_t ovmask;           // bit s: slot s has a per-instance override of some field
  int8_t ovidx[CMTJB_NSLOTS][10];  // override table row of (slot, field) or -1
  DevDriver drv[CMTJB_NSLOTS];
};

template <int NL> struct DevModel {
  DevLayer L[NL];
  double Rp, Rhalf;          // Rp, (Rap-Rp)/2
  int32_t mr_mode;
  int32_t n_layers;
};

constexpr int kMaxOverrides = 24;

// device form of cmtjb_post_t (zeroed = no post-processing)
struct PostParams {
  double *proj_log;         // [n_proj][proj_n_samples][B], this launch's sample s lands in row proj_sample0 + s
  long long proj_n_samples, proj_sample0;
  const double *sd_omega;   // [B]
  double *sd_state;         // [4][B]: z0, z1, z2, sum(y)
  int32_t n_proj, sd_mode;
  double proj[CMTJB_MAX_PROJ][3 * CMTJB_POST_MAX_LAYERS];
  double res[CMTJB_POST_MAX_LAYERS][CMTJB_RES_NFIELDS];
  double sd_b[4], sd_a[4];
};

// the ten Philox4x32 round keys of a run's seed (key + r * Weyl constant)
struct PhiloxKeys {
  uint32_t x[10], y[10];
};
inline PhiloxKeys make_philox_keys(unsigned long long seed) {
  PhiloxKeys K;
  uint32_t kx = (uint32_t)seed, ky = (uint32_t)(seed >> 32);
  for (int r = 0; r < 10; r++) {
    K.x[r] = kx;
    K.y[r] = ky;
    kx += 0x9E3779B9u;
    ky += 0xBB67AE85u;
  }
  return K;
}

// run parameters (also constant bank)
struct RunParams {
  double *m_state;           // [NL*3][B]
  double *log;               // [n_channels][n_samples][B] or null
  int32_t *status;           // [B]
  int32_t *n_samples_out;    // [B] (DP)
  double *time_log;          // [n_samples][B] (DP)
  const double *ovr[kMaxOverrides];  // row k: [B] values of override k
  const double *onef;        // caller-provided, already scaled: [NL*3][total_iterations][B] or null
  const float *onef_gen;     // generated on the device, raw float32: [n_slots][total_iterations][onef_B] or null
  long long onef_i0, onef_B; // onef_gen covers instances onef_i0 .. onef_i0 + onef_B - 1
  long long i_lo, i_hi;      // instances this launch integrates (a sub-batch of 0..B when noise buffers are generated in pieces)
  const double *custom;      // [3*total_iterations][n_custom] pre-evaluated CMTJB_UPD_CUSTOM drivers or null
  long long n_custom;
  // [it_end-it_begin][3][dyn_ncols]: time-dependent driver values at the three stage times
  // t, t+dt/2, t+dt of every step.  Columns 3l..3l+2: summed dynamic field of layer l (A/m);
  // further columns: unit-amplitude waveform of each dynamic scalar slot (DevLayer::dyncol).
  const double *dyn_tab;
  long long dyn_ncols;
  long long B;
  long long inst_offset;
  long long n_samples;       // row count of the log
  long long it_begin, it_end;  // step range of this launch
  long long total_iterations;
  long long write_every;
  long long first_sample;    // sample index of the first write of this launch
  long long elog_ch0;        // first CMTJB_LOG_ENERGY channel, or -1 (general variants only)
  double dt, sqrt_dt, total_time, write_frequency;
  unsigned long long seed;
  PhiloxKeys pk;             // round keys of `seed` (thermal noise)
  int normals64;             // 1: double-precision thermal normals (normal3_f64), 0: the float32 SFU stream
  int32_t tune_flags;        // CMTJB_TUNE bits (experiments): 1 = stateless thermal normals in the sweep kernels
  PostParams post;
};

}  // namespace cmtjb
