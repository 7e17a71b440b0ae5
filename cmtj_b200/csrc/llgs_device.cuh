// llgs_device.cuh -- FP64 device functions of the fused (s)LLGS time-loop kernels.
//
// One thread integrates one sweep instance; its magnetisation vectors live in registers for
// the whole time loop.  Uniform parameters are read from the __grid_constant__ DevModel
// (constant bank), time-dependent drivers are evaluated inline, and thermal noise comes from
// a counter-based Philox4x32-10 stream keyed by (seed, instance, layer, step).
//
// Arithmetic is re-associated relative to the reference (hoisted reciprocals, -gamma/(1+a^2)
// folded into the stage updates, FMA contraction); the parity bar is <= 1e-9 relative on m
// and R against oracle/llgs_oracle.cpp, which itself is bit-exact against the reference.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

#include "dev_model.h"

namespace cmtjb {

struct d3 {
  double x, y, z;
};
__device__ __forceinline__ d3 mk(double x, double y, double z) { return d3{x, y, z}; }
__device__ __forceinline__ d3 operator+(d3 a, d3 b) { return mk(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ d3 operator-(d3 a, d3 b) { return mk(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ d3 operator*(d3 a, double s) { return mk(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ double dot(d3 a, d3 b) { return fma(a.z, b.z, fma(a.y, b.y, a.x * b.x)); }
__device__ __forceinline__ d3 cross(d3 a, d3 b) {
  return mk(fma(a.y, b.z, -(a.z * b.y)), fma(a.z, b.x, -(a.x * b.z)), fma(a.x, b.y, -(a.y * b.x)));
}
// r + a*s
__device__ __forceinline__ d3 axpy(d3 a, double s, d3 r) { return mk(fma(a.x, s, r.x), fma(a.y, s, r.y), fma(a.z, s, r.z)); }
__device__ __forceinline__ d3 ld3(const double *p) { return mk(p[0], p[1], p[2]); }
// r - T*v  (row-major 3x3)
__device__ __forceinline__ d3 sub_matvec(const double *T, d3 v, d3 r) {
  return mk(fma(-T[0], v.x, fma(-T[1], v.y, fma(-T[2], v.z, r.x))), fma(-T[3], v.x, fma(-T[4], v.y, fma(-T[5], v.z, r.y))),
            fma(-T[6], v.x, fma(-T[7], v.y, fma(-T[8], v.z, r.z))));
}

// CVector::normalize (cvector.hpp:185-192): no-op for the zero vector.  With e = |m|^2 - 1 the
// scale is (1+e)^(-1/2).  After a regular RK4 step |e| ~ 1e-11, where 1 - e/2 is exact to
// 3/8 e^2 < 4e-17; first-order steppers and thermal kicks land in |e| < 2e-3, where the
// degree-6 Taylor polynomial is exact to |e|^7 < 2e-19; anything else (first step from an
// un-normalised input) takes the full-precision path.
__device__ __forceinline__ d3 normalize3(d3 v) {
  const double l2 = fma(v.z, v.z, fma(v.y, v.y, v.x * v.x));
  const double e = l2 - 1.0;
  const double ae = fabs(e);
  double s;
  if (ae < 1e-8) {
    s = fma(-0.5, e, 1.0);
  } else if (ae < 2e-3) {
    s = fma(fma(fma(fma(fma(fma(0.2255859375, e, -0.24609375), e, 0.2734375), e, -0.3125), e, 0.375), e, -0.5), e, 1.0);
  } else {
    if (l2 == 0.0) return v;
    s = 1.0 / sqrt(l2);
  }
  return mk(v.x * s, v.y * s, v.z * s);
}

// ---------------------------------------------------------------------------------------
// per-instance overrides
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ double ovr_always(const DevLayer &Lr, const RunParams &R, int s, int field, long long i,
                                             double dflt) {
  if ((Lr.ovmask >> s) & 1u) {
    const int k = Lr.ovidx[s][field];
    if (k >= 0) return __ldg(R.ovr[k] + i);
  }
  return dflt;
}
template <uint32_t F>
__device__ __forceinline__ double ovr_or(const DevLayer &Lr, const RunParams &R, int s, int field, long long i, double dflt) {
  if constexpr ((F & F_OVR) != 0) return ovr_always(Lr, R, s, field, i, dflt);
  return dflt;
}

// ---------------------------------------------------------------------------------------
// drivers: the time-dependent part of ScalarDriver::getCurrentScalarValue
// (drivers.hpp:262-305, helpers :73-108); constantValue is added by the caller.
// ---------------------------------------------------------------------------------------
// unit-amplitude waveform s(t) with  dynamic part = post(A * s(t));  post = |.| for posine,
// min(.,0) for halfsine, identity otherwise.
template <uint32_t F>
__device__ __forceinline__ double driver_shape(const DevLayer &Lr, const RunParams &R, int s, double t, long long i) {
  const DevDriver &d = Lr.drv[s];
  switch (d.type) {
  case CMTJB_UPD_PULSE: {
    const double period = ovr_or<F>(Lr, R, s, CMTJB_FIELD_PERIOD, i, d.period);
    const double cycle = ovr_or<F>(Lr, R, s, CMTJB_FIELD_CYCLE, i, d.cycle);
    const int n = (int)(t / period);
    const double nT = n * period;
    return (nT <= t && t <= (nT + cycle * period)) ? 1.0 : 0.0;
  }
  case CMTJB_UPD_SINE:
  case CMTJB_UPD_POSINE:
  case CMTJB_UPD_HALFSINE: {
    const double f = ovr_or<F>(Lr, R, s, CMTJB_FIELD_FREQUENCY, i, d.frequency);
    const double ph = ovr_or<F>(Lr, R, s, CMTJB_FIELD_PHASE, i, d.phase);
    return sin(2 * M_PI * t * f + ph);
  }
  case CMTJB_UPD_STEP: {
    const double t0 = ovr_or<F>(Lr, R, s, CMTJB_FIELD_TIMESTART, i, d.timeStart);
    const double t1 = ovr_or<F>(Lr, R, s, CMTJB_FIELD_TIMESTOP, i, d.timeStop);
    return (t >= t0 && t <= t1) ? 1.0 : 0.0;
  }
  case CMTJB_UPD_TRAPEZOID: {
    const double t0 = ovr_or<F>(Lr, R, s, CMTJB_FIELD_TIMESTART, i, d.timeStart);
    const double e = ovr_or<F>(Lr, R, s, CMTJB_FIELD_EDGETIME, i, d.edgeTime);
    const double st = ovr_or<F>(Lr, R, s, CMTJB_FIELD_STEADYTIME, i, d.steadyTime);
    if (t < t0) return 0.0;
    if (t <= t0 + e) return (t - t0) / e;
    if (t <= t0 + e + st) return 1.0;
    if (t <= t0 + 2 * e + st) return 1.0 - (t - (t0 + e + st)) / e;
    return 0.0;
  }
  case CMTJB_UPD_GAUSSIMPULSE: {
    const double t0 = ovr_or<F>(Lr, R, s, CMTJB_FIELD_TIMESTART, i, d.timeStart);
    const double e = ovr_or<F>(Lr, R, s, CMTJB_FIELD_EDGETIME, i, d.edgeTime);
    const double x = t - t0;
    return exp(-(x * x) / (2 * (e * e)));
  }
  case CMTJB_UPD_GAUSSSTEP: {
    const double t0 = ovr_or<F>(Lr, R, s, CMTJB_FIELD_TIMESTART, i, d.timeStart);
    const double e = ovr_or<F>(Lr, R, s, CMTJB_FIELD_EDGETIME, i, d.edgeTime);
    return 0.5 * (1 + erf((t - t0) / (sqrt(2.0) * e)));
  }
  default:
    return 0.0;
  }
}
__device__ __forceinline__ double driver_post(int type, double v) {
  if (type == CMTJB_UPD_POSINE) return fabs(v);
  if (type == CMTJB_UPD_HALFSINE) return v <= 0 ? v : 0.0;
  return v;
}
template <uint32_t F>
__device__ __noinline__ double driver_dynamic_inline(const DevLayer &Lr, const RunParams &R, int s, double t, long long i) {
  const double A = ovr_or<F>(Lr, R, s, CMTJB_FIELD_AMPLITUDE, i, Lr.drv[s].amplitude);
  return driver_post(Lr.drv[s].type, A * driver_shape<F>(Lr, R, s, t, i));
}

// Where in time a right-hand-side evaluation sits: the step index and which of the three
// stage times of that step (0: t, 1: t+dt/2, 2: t+dt).  Fixed-step solvers only ever
// evaluate drivers at these times, which are the same for every instance of a batch, so
// F_DYN variants read the waveform from a table built once per launch.
struct StageTime {
  double t;
  // this stage time's row of RunParams::dyn_tab (F_DYN / F_DYNS variants).  Inline-evaluating variants
  // (F_DYNI) have no such table; for them the pointer walks the caller's custom_table instead (same
  // [step][stage][column] layout), which is how CMTJB_UPD_CUSTOM drivers are read there; null otherwise.
  const double *row;
};
// unit waveform of slot s in an inline-evaluating variant: pre-evaluated callbacks from the custom table
template <uint32_t F>
__device__ __forceinline__ double dynamic_inline_at(const DevLayer &Lr, const RunParams &R, int s, StageTime st, long long i) {
  if (Lr.drv[s].type == CMTJB_UPD_CUSTOM) return st.row ? __ldg(st.row + Lr.drv[s].pad) : 0.0;
  return driver_dynamic_inline<F>(Lr, R, s, st.t, i);
}

template <uint32_t F>
__device__ __forceinline__ double driver_dynamic(const DevLayer &Lr, const RunParams &R, int s, StageTime st, long long i) {
  if constexpr ((F & F_DYNI) != 0) {
    return dynamic_inline_at<F>(Lr, R, s, st, i);
  } else {
    const double A = ovr_or<F>(Lr, R, s, CMTJB_FIELD_AMPLITUDE, i, Lr.drv[s].amplitude);
    return driver_post(Lr.drv[s].type, A * __ldg(st.row + Lr.dyncol[s]));
  }
}

// value of a scalar slot at time t
template <uint32_t F>
__device__ __forceinline__ double slot_value(const DevLayer &Lr, const RunParams &R, int s, StageTime t, long long i) {
  double v = ovr_or<F>(Lr, R, s, CMTJB_FIELD_CONSTANT, i, Lr.cst[s]);
  if constexpr ((F & (F_DYNS | F_DYNI)) != 0) {
    if ((Lr.dynmask >> s) & 1u) v += driver_dynamic<F>(Lr, R, s, t, i);
  }
  return v;
}
template <uint32_t F>
__device__ __forceinline__ d3 slot_vec(const DevLayer &Lr, const RunParams &R, int s0, StageTime t, long long i) {
  return mk(slot_value<F>(Lr, R, s0, t, i), slot_value<F>(Lr, R, s0 + 1, t, i), slot_value<F>(Lr, R, s0 + 2, t, i));
}

// Sum of the CONSTANT parts of the four plain field drivers (Hext + Hoe + Hdmi + Hreserved,
// junction.hpp:607-620), including per-instance overrides; evaluated once per launch.
__device__ __forceinline__ d3 constant_field(const DevLayer &Lr, const RunParams &R, long long i) {
  double h[3];
#pragma unroll
  for (int c = 0; c < 3; c++) {
    double v = 0.0;
#pragma unroll
    for (int g = 0; g < 4; g++) v += ovr_always(Lr, R, g * 3 + c, CMTJB_FIELD_CONSTANT, i, Lr.cst[g * 3 + c]);
    h[c] = v;
  }
  return mk(h[0], h[1], h[2]);
}

// time-dependent parts of the same four drivers.  l is the layer index (table columns 3l..3l+2).
template <uint32_t F>
__device__ __forceinline__ d3 dynamic_field(const DevLayer &Lr, const RunParams &R, int l, StageTime t, long long i, d3 H) {
  if constexpr ((F & F_DYNI) != 0) {
    if (Lr.dynmask & 0xFFFu) {
      double h[3] = {H.x, H.y, H.z};
#pragma unroll
      for (int g = 0; g < 4; g++)
#pragma unroll
        for (int c = 0; c < 3; c++)
          if ((Lr.dynmask >> (g * 3 + c)) & 1u) h[c] += dynamic_inline_at<F>(Lr, R, g * 3 + c, t, i);
      return mk(h[0], h[1], h[2]);
    }
  } else if constexpr ((F & F_DYN) != 0) {
    if (Lr.dynH) {
      const double *p = t.row + 3 * l;
      return mk(H.x + __ldg(p), H.y + __ldg(p + 1), H.z + __ldg(p + 2));
    }
  }
  return H;
}

// ---------------------------------------------------------------------------------------
// F_DYNI fast path.  Inline evaluation (per-instance frequency / phase / timing: a VSD map, one
// instance per (field, frequency) pair) walks 12 field slots and their override tables at every
// right-hand-side evaluation: ~1900 bookkeeping instructions per step against ~400 of arithmetic
// (ncu, VSD 256 x 64).  Sweeps excite ONE field component per layer, so the driver of that slot is
// resolved once per launch into registers (type, component, amplitude and three waveform parameters,
// per-instance overrides applied) and the steppers add its value to the constant field at each of the
// step's distinct stage times (three for RK4, not four).  Layers with several time-dependent field
// slots keep the general path (type = kFastGeneric).
// ---------------------------------------------------------------------------------------
constexpr int kFastNone = -2, kFastGeneric = -1;
struct FastDrv {
  int type;  // cmtjb_update of the resolved slot, kFastNone (no time-dependent field) or kFastGeneric
  int comp;  // field component 0..2
  double amp, p0, p1, p2;
};

template <uint32_t F> __device__ __forceinline__ FastDrv resolve_fast(const DevLayer &Lr, const RunParams &R, long long i) {
  FastDrv f{kFastNone, 0, 0.0, 0.0, 0.0, 0.0};
  if constexpr ((F & (F_DYNI | F_FAST)) != 0) {
    const uint32_t mask = Lr.dynmask & 0xFFFu;
    if (mask == 0) return f;
    if (mask & (mask - 1)) {
      f.type = kFastGeneric;
      return f;
    }
    const int s = __ffs((int)mask) - 1;
    const DevDriver &d = Lr.drv[s];
    f.type = d.type;
    f.comp = s % 3;
    f.amp = ovr_always(Lr, R, s, CMTJB_FIELD_AMPLITUDE, i, d.amplitude);
    switch (d.type) {
    case CMTJB_UPD_PULSE:
      f.p0 = ovr_always(Lr, R, s, CMTJB_FIELD_PERIOD, i, d.period);
      f.p1 = ovr_always(Lr, R, s, CMTJB_FIELD_CYCLE, i, d.cycle);
      break;
    case CMTJB_UPD_SINE:
    case CMTJB_UPD_POSINE:
    case CMTJB_UPD_HALFSINE:
      f.p0 = ovr_always(Lr, R, s, CMTJB_FIELD_FREQUENCY, i, d.frequency);
      f.p1 = ovr_always(Lr, R, s, CMTJB_FIELD_PHASE, i, d.phase);
      break;
    case CMTJB_UPD_STEP:
      f.p0 = ovr_always(Lr, R, s, CMTJB_FIELD_TIMESTART, i, d.timeStart);
      f.p1 = ovr_always(Lr, R, s, CMTJB_FIELD_TIMESTOP, i, d.timeStop);
      break;
    case CMTJB_UPD_TRAPEZOID:
      f.p0 = ovr_always(Lr, R, s, CMTJB_FIELD_TIMESTART, i, d.timeStart);
      f.p1 = ovr_always(Lr, R, s, CMTJB_FIELD_EDGETIME, i, d.edgeTime);
      f.p2 = ovr_always(Lr, R, s, CMTJB_FIELD_STEADYTIME, i, d.steadyTime);
      break;
    case CMTJB_UPD_GAUSSIMPULSE:
    case CMTJB_UPD_GAUSSSTEP:
      f.p0 = ovr_always(Lr, R, s, CMTJB_FIELD_TIMESTART, i, d.timeStart);
      f.p1 = ovr_always(Lr, R, s, CMTJB_FIELD_EDGETIME, i, d.edgeTime);
      break;
    default: f.type = kFastGeneric; break;  // custom etc.: not evaluable inline
    }
  }
  return f;
}

// same expressions as driver_shape, on the resolved parameters
__device__ __forceinline__ double fast_shape(const FastDrv &f, double t) {
  switch (f.type) {
  case CMTJB_UPD_PULSE: {
    const int n = (int)(t / f.p0);
    const double nT = n * f.p0;
    return (nT <= t && t <= (nT + f.p1 * f.p0)) ? 1.0 : 0.0;
  }
  case CMTJB_UPD_SINE:
  case CMTJB_UPD_POSINE:
  case CMTJB_UPD_HALFSINE: return sin(2 * M_PI * t * f.p0 + f.p1);
  case CMTJB_UPD_STEP: return (t >= f.p0 && t <= f.p1) ? 1.0 : 0.0;
  case CMTJB_UPD_TRAPEZOID: {
    const double t0 = f.p0, e = f.p1, st = f.p2;
    if (t < t0) return 0.0;
    if (t <= t0 + e) return (t - t0) / e;
    if (t <= t0 + e + st) return 1.0;
    if (t <= t0 + 2 * e + st) return 1.0 - (t - (t0 + e + st)) / e;
    return 0.0;
  }
  case CMTJB_UPD_GAUSSIMPULSE: {
    const double x = t - f.p0;
    return exp(-(x * x) / (2 * (f.p1 * f.p1)));
  }
  case CMTJB_UPD_GAUSSSTEP: return 0.5 * (1 + erf((t - f.p0) / (sqrt(2.0) * f.p1)));
  default: return 0.0;
  }
}

// constant field + time-dependent field of layer l at time t (F_DYNI variants)
template <uint32_t F>
__device__ __forceinline__ d3 stage_field(const DevLayer &Lr, const RunParams &R, int l, long long i, StageTime st, d3 Hc,
                                          const FastDrv &f) {
  const double t = st.t;
  if (f.type == kFastNone) return Hc;
  if (f.type == kFastGeneric) return dynamic_field<F>(Lr, R, l, st, i, Hc);
  const double v = driver_post(f.type, f.amp * fast_shape(f, t));
  return mk(Hc.x + (f.comp == 0 ? v : 0.0), Hc.y + (f.comp == 1 ? v : 0.0), Hc.z + (f.comp == 2 ? v : 0.0));
}

// ---------------------------------------------------------------------------------------
// External per-thread inputs of the coupled-system kernels (coupled_kernels.cuh): a Stack
// replaces every layer's current driver each step (stack.hpp:79-91).  NoExt = plain junction.
// ---------------------------------------------------------------------------------------
struct NoExt {
  static constexpr bool kCurrent = false;
  __device__ __forceinline__ double current(double) const { return 0.0; }
};

// ---------------------------------------------------------------------------------------
// u(t, m) with dm/dt = Lr.llg * u : field assembly (calculateHeff, junction.hpp:579-622)
// fused with solveLLG (junction.hpp:716-815).
//   HAS_BOT / HAS_TOP: the layer has a neighbour on that side (else the reference passes 0).
// ---------------------------------------------------------------------------------------
//   PRE: Hc already holds the complete field of the four plain drivers at this stage time (stage_field).
template <uint32_t F, bool HAS_BOT, bool HAS_TOP, class Ext = NoExt, bool PRE = false>
__device__ __forceinline__ d3 llg_u(const DevLayer &Lr, const RunParams &R, int l, long long i, StageTime t, d3 m, d3 bot,
                                    d3 top, d3 Hc, const Ext &ext = Ext()) {
  static_assert((F & F_DEMAG_FULL) == 0 || (F & F_NOFOLD) != 0, "a full demag tensor cannot be folded into the diagonal operator");
  d3 H = PRE ? Hc : dynamic_field<F>(Lr, R, l, t, i, Hc);
  const d3 a = ld3(Lr.anis);
  if constexpr ((F & F_NOFOLD) != 0) {  // K1 (junction.hpp:638-643)
    const double K = slot_value<F>(Lr, R, CMTJB_SLOT_K1, t, i);
    H = axpy(a, (Lr.cK1 * K) * dot(a, m), H);
  } else {
    // K1 anisotropy and demag folded on the host into one diagonal operator (constant K1,
    // axis-aligned easy axis, diagonal demag tensor): three FMAs instead of ten.
    H = mk(fma(Lr.lin[0], m.x, H.x), fma(Lr.lin[1], m.y, H.y), fma(Lr.lin[2], m.z, H.z));
  }
  if constexpr ((F & F_K2) != 0) {  // (junction.hpp:645-652)
    const double K2 = slot_value<F>(Lr, R, CMTJB_SLOT_K2, t, i);
    const double am = dot(a, m);
    H = axpy(a, (Lr.cK2 * K2) * (am * am * am), H);
  }
  if constexpr ((F & F_IEC) != 0) {  // (junction.hpp:654-675)
    if constexpr ((F & (F_DYNS | F_DYNI | F_OVR)) == 0) {
      // constant, batch-uniform couplings: J/(Ms t) and 2 J2/(Ms t) hoisted to the host
      if constexpr (HAS_BOT) H = axpy(bot, fma(Lr.iecB_b, dot(bot, m), Lr.iecB_a), H);
      if constexpr (HAS_TOP) H = axpy(top, fma(Lr.iecT_b, dot(top, m), Lr.iecT_a), H);
    } else {
      if constexpr (HAS_BOT) {
        const double J = slot_value<F>(Lr, R, CMTJB_SLOT_IEC_BOTTOM, t, i);
        const double J2 = slot_value<F>(Lr, R, CMTJB_SLOT_IECQ_BOTTOM, t, i);
        H = axpy(bot, fma(2 * J2 * Lr.inv_Mst, dot(bot, m), J * Lr.inv_Mst), H);
      }
      if constexpr (HAS_TOP) {
        const double J = slot_value<F>(Lr, R, CMTJB_SLOT_IEC_TOP, t, i);
        const double J2 = slot_value<F>(Lr, R, CMTJB_SLOT_IECQ_TOP, t, i);
        H = axpy(top, fma(2 * J2 * Lr.inv_Mst, dot(top, m), J * Lr.inv_Mst), H);
      }
    }
  }
  if constexpr ((F & F_IDMI) != 0) {  // (junction.hpp:677-701)
    if constexpr (HAS_BOT) {
      const d3 D = slot_vec<F>(Lr, R, CMTJB_SLOT_IDMI_BOTTOM_X, t, i);
      H = axpy(cross(D, bot), -Lr.inv_Mst, H);
    }
    if constexpr (HAS_TOP) {
      const d3 D = slot_vec<F>(Lr, R, CMTJB_SLOT_IDMI_TOP_X, t, i);
      H = axpy(cross(D, top), -Lr.inv_Mst, H);
    }
  }
  // demag (junction.hpp:598-599); off-diagonal entries only in F_DEMAG_FULL variants
  if constexpr ((F & F_DEMAG_FULL) != 0) {
    H = sub_matvec(Lr.demag, m, H);
  } else if constexpr ((F & F_NOFOLD) != 0) {
    H = mk(fma(-Lr.demag[0], m.x, H.x), fma(-Lr.demag[4], m.y, H.y), fma(-Lr.demag[8], m.z, H.z));
  }
  if constexpr ((F & F_DIPOLE) != 0) {  // (junction.hpp:583-585)
    if constexpr (HAS_BOT) H = sub_matvec(Lr.dipB, bot, H);
    if constexpr (HAS_TOP) H = sub_matvec(Lr.dipT, top, H);
  }

  // --- solveLLG ------------------------------------------------------------------------
  const d3 p = cross(m, H);
  const d3 am3 = m * Lr.alpha;
  d3 u = mk(fma(am3.y, p.z, fma(-am3.z, p.y, p.x)), fma(am3.z, p.x, fma(-am3.x, p.z, p.y)),
            fma(am3.x, p.y, fma(-am3.y, p.x, p.z)));
  if constexpr ((F & (F_STT | F_SOT)) != 0) {
    d3 ref = mk(0, 0, 0);
    if (Lr.refType == CMTJB_REF_FIXED) ref = ld3(Lr.ref);
    else if (Lr.refType == CMTJB_REF_TOP) ref = top;
    else if (Lr.refType == CMTJB_REF_BOTTOM) ref = bot;
    if constexpr ((F & F_STT) != 0) {
      if (Lr.stt) {
        const double I = Ext::kCurrent ? ext.current(t.t) : slot_value<F>(Lr, R, CMTJB_SLOT_CURRENT, t, i);
        const double mp = dot(m, ref);
        // eta = P lam^2 / (lam^2 + 1 + (lam^2 - 1) m.p)  or  P / (1 + lam m.p)  (junction.hpp:748-757), the
        // constants folded on the host; reciprocal + multiply instead of the library division
        const double eta = Lr.stt_num * (1.0 / fma(Lr.stt_a, mp, Lr.stt_b));
        const double ae = Lr.aJ_c * I * eta;
        const d3 fl = cross(m, ref);
        const d3 dl = cross(m, fl);
        u = axpy(dl, ae * Lr.kappa, u);
        u = axpy(fl, -(ae * Lr.beta), u);
        return u;
      }
    }
    if constexpr ((F & F_SOT) != 0) {
      if (Lr.sot) {
        double Hdl, Hfl, Hdl2 = 0, Hfl2 = 0;
        if (Lr.dynamicSOT) {
          Hdl = slot_value<F>(Lr, R, CMTJB_SLOT_DL, t, i);
          Hfl = slot_value<F>(Lr, R, CMTJB_SLOT_FL, t, i);
          Hdl2 = slot_value<F>(Lr, R, CMTJB_SLOT_DL2, t, i);
          Hfl2 = slot_value<F>(Lr, R, CMTJB_SLOT_FL2, t, i);
        } else {
          const double I = Ext::kCurrent ? ext.current(t.t) : slot_value<F>(Lr, R, CMTJB_SLOT_CURRENT, t, i);
          Hdl = Lr.dlT * I;
          Hfl = Lr.flT * I;
        }
        const d3 cm1 = cross(m, ref);
        const d3 ccm1 = cross(m, cm1);
        u = axpy(cm1, fma(-Lr.alpha, Hdl, Hfl), u);
        u = axpy(ccm1, fma(Lr.alpha, Hfl, Hdl), u);
        if (Lr.dynamicSOT) {
          const d3 ref2 = ld3(Lr.ref2);
          const d3 cm2 = cross(m, ref2);
          const d3 ccm2 = cross(m, cm2);
          u = axpy(cm2, fma(-Lr.alpha, Hdl2, Hfl2), u);
          u = axpy(ccm2, fma(Lr.alpha, Hfl2, Hdl2), u);
        }
      }
    }
  }
  return u;
}

// stochasticTorque (junction.hpp:824-831) in the same "u" units: g = Lr.llg * (m x dW + a m x (m x dW))
__device__ __forceinline__ d3 stoch_u(const DevLayer &Lr, d3 m, d3 dW) {
  const d3 c1 = cross(m, dW);
  const d3 c2 = cross(m, c1);
  return axpy(c2, Lr.alpha, c1);
}

// ---------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al., SC'11) + Box-Muller.  Normals are produced in FP32 (the
// FP32/SFU pipes are otherwise idle) and widened to FP64; validated distributionally in
// tests/test_stochastic_gpu.py.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; r++) {
    const uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
    const uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += W0;
    key.y += W1;
  }
  return ctr;
}

// Three standard normals from one Philox block.  Everything after the counter hash runs on the FP32 /
// SFU pipes, which the FP64 time loop leaves idle: uniforms are built from the top 23 bits by exponent
// splicing (no I2F), Box-Muller uses the SFU approximations (MUFU.LG2 / RSQ / SIN / COS, absolute
// error ~2^-21: six orders of magnitude below what any statistic of a Langevin run can resolve; the
// stream is validated distributionally in tests/test_stochastic_gpu.py -- moments, tails, KS, PSD).
// u = 1 - (r + 0.5) 2^-32 in (0, 1): top 23 bits spliced into [1, 2), the low 9 bits into an integer-
// valued float; FP32's relative precision keeps all 32 bits where they matter, next to 0 (the tail of
// the radius reaches 6.7 sigma as with a 32-bit uniform).
__device__ __forceinline__ float unit_open_low(uint32_t r) {
  const float hi = 2.0f - __uint_as_float(0x3f800000u | (r >> 9));              // k 2^-23, k = 1 .. 2^23
  const float lo = __uint_as_float(0x4b000000u | (r & 0x1ffu)) - 8388608.0f;    // 0 .. 511
  return fmaf(lo, -2.3283064365386963e-10f, hi) - 1.1641532182693481e-10f;      // - lo 2^-32 - 2^-33
}
// the same hash with the ten round keys taken from the constant bank (RunParams::pk, filled on the host from the
// seed): the key schedule costs nothing in the time loop
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, const PhiloxKeys &K) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
#pragma unroll
  for (int r = 0; r < 10; r++) {
    const uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
    const uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ K.x[r], lo1, hi0 ^ ctr.w ^ K.y[r], lo0);
  }
  return ctr;
}
__device__ __forceinline__ PhiloxKeys philox_keys(unsigned long long seed) {
  PhiloxKeys K;
  uint32_t kx = (uint32_t)seed, ky = (uint32_t)(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; r++) {
    K.x[r] = kx;
    K.y[r] = ky;
    kx += 0x9E3779B9u;
    ky += 0xBB67AE85u;
  }
  return K;
}

// four N(0,1) float32 values from one Philox block: two Box-Muller pairs, radius sqrt(-2 ln u) with u in (0, 1],
// t = -2 ln 2 * log2(u) >= 0 and sqrt(t) = t * rsqrt(t) (0 at t = 0); angle uniform on [-pi, pi), where the SFU sine
// and cosine are most accurate
__device__ __forceinline__ float4 normal4(uint4 r) {
  const float t0 = -1.3862943611198906f * __log2f(unit_open_low(r.x));
  const float t1 = -1.3862943611198906f * __log2f(unit_open_low(r.z));
  const float r0 = t0 * rsqrtf(fmaxf(t0, 1e-30f)), r1 = t1 * rsqrtf(fmaxf(t1, 1e-30f));
  const float a0 = 6.283185307179586f * (__uint_as_float(0x3f800000u | (r.y >> 9)) - 1.5f);
  const float a1 = 6.283185307179586f * (__uint_as_float(0x3f800000u | (r.w >> 9)) - 1.5f);
  return make_float4(r0 * __cosf(a0), r0 * __sinf(a0), r1 * __cosf(a1), r1 * __sinf(a1));
}

// ---- the thermal stream ----------------------------------------------------------------------------------------
// Normal n = 3 j + c (component c of step s = 4 q + j) of a (seed, instance, layer) stream is element n % 4 of
// Philox block (q, n / 4): the twelve normals of four consecutive steps come from THREE blocks, every normal of
// every block is used, and each value is a pure function of (seed, instance, layer, step, component) -- shards,
// sub-batches and the persistent kernel's task boundaries reproduce the same numbers.
__device__ __forceinline__ float4 thermal_block(const PhiloxKeys &K, unsigned long long inst, unsigned layer, unsigned long long q,
                                                unsigned sub) {
  const uint4 ctr = make_uint4((uint32_t)inst, (uint32_t)(inst >> 32), (uint32_t)q, ((uint32_t)(q >> 32) << 8) | (sub << 4) | (layer & 0xFu));
  return normal4(philox4x32_10(ctr, K));
}
__device__ __forceinline__ float pick4(float4 v, unsigned k) { return k == 0 ? v.x : (k == 1 ? v.y : (k == 2 ? v.z : v.w)); }

// stateless form: one or two blocks per step (coupled and lane-per-layer kernels, test hook)
__device__ __forceinline__ d3 normal3(const PhiloxKeys &K, unsigned long long inst, unsigned layer, unsigned long long step) {
  const unsigned n0 = 3u * (unsigned)(step & 3ull);
  const unsigned long long q = step >> 2;
  const unsigned s0 = n0 >> 2, s1 = (n0 + 2u) >> 2;
  const float4 b0 = thermal_block(K, inst, layer, q, s0);
  float4 b1 = b0;
  if (s1 != s0) b1 = thermal_block(K, inst, layer, q, s1);
  float v[3];
#pragma unroll
  for (unsigned c = 0; c < 3; c++) v[c] = pick4(((n0 + c) >> 2) == s0 ? b0 : b1, (n0 + c) & 3u);
  return mk((double)v[0], (double)v[1], (double)v[2]);
}

// carried form (time loops that own a layer for consecutive steps): the spare normals of a block wait in three
// registers for the following steps, so a block is hashed exactly once -- 0.75 Philox blocks per layer-step.
struct NoiseCarry {
  float a, b, c;
};
// the spares as they stand BEFORE step `step` (a time loop that starts in the middle of a quad)
__device__ __forceinline__ NoiseCarry carry_before(const PhiloxKeys &K, unsigned long long inst, unsigned layer, unsigned long long step) {
  NoiseCarry nc{0.f, 0.f, 0.f};
  const unsigned j = (unsigned)(step & 3ull);
  if (j != 0) {
    const float4 v = thermal_block(K, inst, layer, step >> 2, j - 1);
    if (j == 1) nc.a = v.w;                  // normal 3
    else if (j == 2) nc.a = v.z, nc.b = v.w; // normals 6, 7
    else nc.a = v.y, nc.b = v.z, nc.c = v.w; // normals 9, 10, 11
  }
  return nc;
}
__device__ __forceinline__ d3 normal3_carried(const PhiloxKeys &K, unsigned long long inst, unsigned layer, unsigned long long step,
                                              NoiseCarry &nc) {
  const unsigned j = (unsigned)(step & 3ull);  // uniform across the warp: no divergence
  if (j == 3) return mk((double)nc.a, (double)nc.b, (double)nc.c);
  const float4 v = thermal_block(K, inst, layer, step >> 2, j);
  d3 out;
  if (j == 0) {
    out = mk((double)v.x, (double)v.y, (double)v.z);
    nc.a = v.w;
  } else if (j == 1) {
    out = mk((double)nc.a, (double)v.x, (double)v.y);
    nc.a = v.z;
    nc.b = v.w;
  } else {
    out = mk((double)nc.a, (double)nc.b, (double)v.x);
    nc.a = v.y;
    nc.b = v.z;
    nc.c = v.w;
  }
  return out;
}

// ---- double-precision alternative (cmtjb_run_args_t::thermal_normals = 1) ---------------------------------------------
// The reference draws std::normal_distribution<double>.  The default stream above is float32 (absolute error 2^-21 per
// normal, radius tail ending at 6.7 sigma): more than any Langevin statistic tested resolves, but a caller who studies rare
// events can ask for this one instead: two Philox blocks per layer-step (sub-counters 8 and 9, disjoint from the float32
// stream's 0..2), 53-bit uniforms u = (k + 1) 2^-53 in (0, 1], Box-Muller with the double-precision library log / sqrt /
// sincospi: relative error ~1e-16, tail to 8.6 sigma.  Out of line and cold: the default path pays one predicate for it.
static __device__ __noinline__ d3 normal3_f64(uint32_t kx, uint32_t ky, unsigned long long inst, unsigned layer, unsigned long long step) {
  double n[4];
#pragma unroll
  for (unsigned b = 0; b < 2; b++) {
    const uint4 ctr = make_uint4((uint32_t)inst, (uint32_t)(inst >> 32), (uint32_t)step,
                                 ((uint32_t)(step >> 32) << 8) | ((8u + b) << 4) | (layer & 0xFu));
    const uint4 r = philox4x32_10(ctr, make_uint2(kx, ky));
    const unsigned long long k1 = (((unsigned long long)r.x << 32) | r.y) >> 11, k2 = (((unsigned long long)r.z << 32) | r.w) >> 11;
    const double u1 = ((double)k1 + 1.0) * 1.1102230246251565e-16;  // (0, 1]
    const double u2 = (double)k2 * 1.1102230246251565e-16;          // [0, 1)
    const double rad = sqrt(-2.0 * log(u1));
    double sn, cs;
    sincospi(2.0 * u2, &sn, &cs);
    n[2 * b] = rad * cs;
    n[2 * b + 1] = rad * sn;
  }
  return mk(n[0], n[1], n[2]);
}

// the step's noise vector dW (junction.hpp:1121-1136): Langevin + 1/f^alpha buffer
template <uint32_t F, int NL>
__device__ __forceinline__ d3 noise_dW(const DevLayer &Lr, const RunParams &R, int l, long long i, long long it, StageTime t,
                                       NoiseCarry *nc = nullptr) {
  d3 dW = mk(0, 0, 0);
  if constexpr ((F & F_THERMAL) != 0) {
    if (Lr.thermal) {
      double sigma;
      if constexpr ((F & (F_OVR | F_DYNS | F_DYNI)) == 0) {
        sigma = Lr.sigma0;  // constant, batch-uniform temperature: the square root is taken on the host
      } else {
        const double T = slot_value<F>(Lr, R, CMTJB_SLOT_TEMPERATURE, t, i);
        sigma = sqrt(Lr.sigma_c * T);
      }
      const unsigned long long inst = (unsigned long long)(R.inst_offset + i);
      if (R.normals64) dW = normal3_f64(R.pk.x[0], R.pk.y[0], inst, (unsigned)l, (unsigned long long)it) * sigma;
      else dW = (nc ? normal3_carried(R.pk, inst, (unsigned)l, (unsigned long long)it, *nc) : normal3(R.pk, inst, (unsigned)l, (unsigned long long)it)) * sigma;
    }
  }
  if constexpr ((F & F_ONEF) != 0) {
    if (Lr.onef) {
      if (R.onef) {  // caller's buffer: VectorAlphaNoise::tickVector values, scaled
        const double *b = R.onef + ((long long)(l * 3) * R.total_iterations + it) * R.B + i;
        const long long cs = R.total_iterations * R.B;
        dW = dW + mk(__ldg(b), __ldg(b + cs), __ldg(b + 2 * cs));
      } else if (R.onef_gen) {
        // tickVector (noise.hpp:291-301) on the raw component buffers: widen the float32 samples, normalise the
        // vector when all three axes carry noise (CVector::normalize: no-op for the zero vector), scale
        const float *b = R.onef_gen + ((long long)Lr.onef_slot * R.total_iterations + it) * R.onef_B + (i - R.onef_i0);
        if (Lr.onef_axis == CMTJB_AXIS_ALL) {
          const long long cs = R.total_iterations * R.onef_B;
          const d3 v = mk((double)__ldg(b), (double)__ldg(b + cs), (double)__ldg(b + 2 * cs));
          const double l2 = dot(v, v);
          const double s = l2 != 0.0 ? Lr.onef_scale / sqrt(l2) : Lr.onef_scale;
          dW = axpy(v, s, dW);
        } else {
          const double x = (double)__ldg(b) * Lr.onef_scale;
          dW = mk(dW.x + (Lr.onef_axis == CMTJB_AXIS_X ? x : 0.0), dW.y + (Lr.onef_axis == CMTJB_AXIS_Y ? x : 0.0),
                  dW.z + (Lr.onef_axis == CMTJB_AXIS_Z ? x : 0.0));
        }
      }
    }
  }
  return dW;
}

}  // namespace cmtjb
