/*
 * llgs_oracle.cpp -- CPU restatement of the reference's (s)LLGS junction integration.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the parity oracle for the CUDA path in
 * cmtj_b200/csrc; it is compiled by oracle/Makefile into oracle/liboracle.so and may be
 * loaded only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg.  The
 * product never links or falls back to it.
 *
 * Parity status: PINNED.  oracle/_ref/_cmtj*.so is the unmodified reference extension
 * (python/cmtj.cpp) compiled from /root/reference by oracle/Makefile; tests/test_oracle_vs_ref.py
 * checks this restatement BIT-FOR-BIT against it on deterministic runs, and tests/golden/
 * holds vectors generated from it (tests/golden/make_golden.py).
 *
 * The arithmetic below deliberately keeps the reference's operation order (scalar, no FMA
 * contraction on baseline x86-64) so the comparison can be exact; it is written from the
 * algorithm, with each function citing the reference lines it restates.
 *
 * It consumes the same raw model records as the C-ABI (include/cmtj_b200.h).
 */
#include "../include/cmtj_b200.h"

#include <cmath>
#include <cstdio>
#include <cstring>
#include <random>
#include <string>
#include <vector>

namespace {

thread_local std::string g_err;

int fail(int code, const std::string &msg) {
  g_err = msg;
  return code;
}

struct V3 {
  double x, y, z;
};
inline V3 v3(double x, double y, double z) { return V3{x, y, z}; }
inline V3 operator+(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline V3 operator-(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline V3 operator*(V3 a, double s) { return v3(a.x * s, a.y * s, a.z * s); }
inline V3 operator*(double s, V3 a) { return v3(s * a.x, s * a.y, s * a.z); }
inline V3 operator/(V3 a, double s) { return v3(a.x / s, a.y / s, a.z / s); }
/* core/junction.hpp:74-84 */
inline V3 cross(V3 a, V3 b) {
  return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
inline double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
/* core/cvector.hpp:179-192 */
inline double vlen(V3 a) { return std::sqrt(std::pow(a.x, 2) + std::pow(a.y, 2) + std::pow(a.z, 2)); }
inline V3 normalized(V3 a) {
  const double l = vlen(a);
  if (l != 0) return v3(a.x / l, a.y / l, a.z / l);
  return a;
}
inline V3 arr3(const double *p) { return v3(p[0], p[1], p[2]); }

/* CMTJB_UPD_CUSTOM (drivers.hpp:294-303: the Python callback's value IS the driver's value).  A callback
 * cannot cross the C boundary; the caller pre-evaluates it at the stage times of every step
 * (cmtjb_run_args_t::custom_table) and the time loop publishes which step it is in, so that the stage can
 * be recognised from the time argument by exact comparison with the reference's own time expressions. */
struct CustomCtx {
  const double *table = nullptr; /* [3 * total_iterations][n_custom] */
  int64_t n_custom = 0, step = 0;
  double t0 = 0, dt = 0;
};
thread_local CustomCtx g_custom;

double custom_value(const cmtjb_driver_t &d, double time) {
  const CustomCtx &c = g_custom;
  if (!c.table || d.custom_column < 0 || d.custom_column >= c.n_custom) return 0.0;
  const int stage = (time == c.t0) ? 0 : (time == c.t0 + 0.5 * c.dt ? 1 : 2);
  return c.table[(3 * c.step + stage) * c.n_custom + d.custom_column];
}

/* core/drivers.hpp:262-305 with helpers :73-108 */
double driver_value(const cmtjb_driver_t &d, double time) {
  if (d.update == CMTJB_UPD_CUSTOM) return custom_value(d, time);
  double ret = d.constantValue;
  switch (d.update) {
  case CMTJB_UPD_PULSE: {
    const int n = static_cast<int>(time / d.period);
    const double dT = d.cycle * d.period;
    const double nT = n * d.period;
    if (nT <= time && time <= (nT + dT)) ret += d.amplitude;
    break;
  }
  case CMTJB_UPD_SINE:
    ret += d.amplitude * std::sin(2 * M_PI * time * d.frequency + d.phase);
    break;
  case CMTJB_UPD_POSINE:
    ret += std::abs(d.amplitude * std::sin(2 * M_PI * time * d.frequency + d.phase));
    break;
  case CMTJB_UPD_HALFSINE: {
    const double tamp = d.amplitude * std::sin(2 * M_PI * time * d.frequency + d.phase);
    if (tamp <= 0) ret += tamp;
    break;
  }
  case CMTJB_UPD_STEP:
    if (time >= d.timeStart && time <= d.timeStop) ret += d.amplitude;
    break;
  case CMTJB_UPD_TRAPEZOID: {
    double v = 0;
    if (time < d.timeStart) {
      v = 0;
    } else if (time <= d.timeStart + d.edgeTime) {
      v = (d.amplitude / d.edgeTime) * (time - d.timeStart);
    } else if (time <= d.timeStart + d.edgeTime + d.steadyTime) {
      v = d.amplitude;
    } else if (time <= d.timeStart + 2 * d.edgeTime + d.steadyTime) {
      v = d.amplitude - (d.amplitude / d.edgeTime) * (time - (d.timeStart + d.edgeTime + d.steadyTime));
    }
    ret += v;
    break;
  }
  case CMTJB_UPD_GAUSSIMPULSE:
    ret += d.amplitude * std::exp(-std::pow(time - d.timeStart, 2) / (2 * std::pow(d.edgeTime, 2)));
    break;
  case CMTJB_UPD_GAUSSSTEP:
    ret += 0.5 * d.amplitude * (1 + std::erf((time - d.timeStart) / (std::sqrt(2) * d.edgeTime)));
    break;
  default:
    break;
  }
  return ret;
}

/* One layer of one instance: raw record + this instance's driver set (overrides applied). */
struct LayerSim {
  const cmtjb_layer_t *L;
  const cmtjb_junction_t *J;
  cmtjb_driver_t drv[CMTJB_NSLOTS];
  V3 mag;
  std::mt19937 rng;
  std::normal_distribution<double> normal{0.0, 1.0};
  const double *onef; /* this instance's [3][n_iter] view base, stride info below */
  int64_t onef_comp_stride, onef_iter_stride;
  /* adaptive controller state (core/junction.hpp:136-138) */
  double integral_error, prev_error_ratio;
  /* values cached by the last RHS evaluation, logged when calculateEnergies (core/junction.hpp:226-236) */
  mutable double K_log = 0, K2_log = 0, Jbottom_log = 0, Jtop_log = 0, I_log = 0;
  mutable V3 cHext{0, 0, 0}, cHIEC{0, 0, 0}, cHAnis{0, 0, 0}, cHdemag{0, 0, 0}, cHfl{0, 0, 0}, cHdl{0, 0, 0};

  double slot(int s, double t) const { return driver_value(drv[s], t); }
  V3 axial(int s0, double t) const { return v3(slot(s0, t), slot(s0 + 1, t), slot(s0 + 2, t)); }
};

/* core/junction.hpp:54-62 */
V3 tensor_interaction(V3 m, const double *T, double Ms, double mu0) {
  V3 r = v3(T[0] * m.x + T[1] * m.y + T[2] * m.z, T[3] * m.x + T[4] * m.y + T[5] * m.z,
            T[6] * m.x + T[7] * m.y + T[8] * m.z);
  return r * (Ms / mu0);
}

/* core/junction.hpp:654-662 */
V3 iec_term(const LayerSim &s, double Jv, double J2v, V3 m, V3 coupled) {
  return coupled * (Jv + 2 * J2v * dot(coupled, m)) / (s.L->Ms * s.L->thickness);
}
/* core/junction.hpp:677-688 */
V3 idmi_term(const LayerSim &s, V3 D, V3 coupled) {
  const V3 c = -1.0 * cross(D, coupled);
  return c / (s.L->Ms * s.L->thickness);
}

/* core/junction.hpp:579-622 (calculateHeff + calculateHeffDipoleInjection), terms :624-701 */
V3 heff(const LayerSim &s, double t, V3 m, V3 bottom, V3 top) {
  const cmtjb_layer_t &L = *s.L;
  const double mu0 = s.J->MAGNETIC_PERMEABILITY;
  const V3 Hdipole = tensor_interaction(bottom, L.dipoleBottom, L.Ms, mu0) +
                     tensor_interaction(top, L.dipoleTop, L.Ms, mu0);
  const V3 Hext = s.axial(CMTJB_SLOT_HEXT_X, t);
  const V3 Hoe = s.axial(CMTJB_SLOT_HOE_X, t);
  const V3 Hdemag = tensor_interaction(m, L.demagTensor, L.Ms, mu0);
  const V3 HIEC = iec_term(s, s.slot(CMTJB_SLOT_IEC_BOTTOM, t), s.slot(CMTJB_SLOT_IECQ_BOTTOM, t), m, bottom) +
                  iec_term(s, s.slot(CMTJB_SLOT_IEC_TOP, t), s.slot(CMTJB_SLOT_IECQ_TOP, t), m, top);
  const V3 Hidmi = idmi_term(s, s.axial(CMTJB_SLOT_IDMI_BOTTOM_X, t), bottom) +
                   idmi_term(s, s.axial(CMTJB_SLOT_IDMI_TOP_X, t), top);
  const V3 anis = arr3(L.anis);
  const double K = s.slot(CMTJB_SLOT_K1, t);
  const V3 HAnis = anis * ((2 * K) * dot(anis, m) / L.Ms);
  const V3 Hdmi = s.axial(CMTJB_SLOT_HDMI_X, t);
  const V3 Hres = s.axial(CMTJB_SLOT_HRES_X, t);
  const double K2 = s.slot(CMTJB_SLOT_K2, t);
  const V3 HAnis2 = anis * ((4 * K2) * std::pow(dot(anis, m), 3) / L.Ms);
  const V3 zero = v3(0, 0, 0);
  s.cHext = Hext;
  s.cHdemag = Hdemag;
  s.cHIEC = HIEC;
  s.cHAnis = HAnis;
  s.K_log = K;
  s.K2_log = K2;
  s.Jbottom_log = s.slot(CMTJB_SLOT_IEC_BOTTOM, t);
  s.Jtop_log = s.slot(CMTJB_SLOT_IEC_TOP, t);
  return Hext + HAnis + HAnis2 + HIEC + Hidmi + Hoe + Hdmi + zero - Hdemag - Hdipole + Hres;
}

/* core/junction.hpp:716-815 */
V3 solve_llg(const LayerSim &s, double t, V3 m, V3 bottom, V3 top, V3 H) {
  const cmtjb_layer_t &L = *s.L;
  const double GYRO = s.J->GYRO;
  const V3 prod = cross(m, H);
  const V3 prod2 = cross(m, prod);
  const double conv = 1 / (1 + std::pow(L.damping, 2));
  const V3 dmdt = prod + prod2 * L.damping;
  V3 ref = v3(0, 0, 0);
  switch (L.referenceType) {
  case CMTJB_REF_FIXED: ref = arr3(L.referenceLayer); break;
  case CMTJB_REF_TOP: ref = top; break;
  case CMTJB_REF_BOTTOM: ref = bottom; break;
  default: break;
  }
  if (L.includeSTT) {
    const double I = s.slot(CMTJB_SLOT_CURRENT, t);
    s.I_log = I;
    const double aJ = s.J->HBAR * I / (s.J->ELECTRON_CHARGE * L.Ms * L.thickness);
    double eta;
    if (L.alternativeSTT) {
      eta = (L.spinPolarisation) / (1 + L.SlonczewskiSpacerLayerParameter * dot(m, ref));
    } else {
      const double sl2 = std::pow(L.SlonczewskiSpacerLayerParameter, 2);
      eta = (L.spinPolarisation * sl2) / (sl2 + 1 + (sl2 - 1) * dot(m, ref));
    }
    const double stt = GYRO * aJ * eta;
    const V3 fl = cross(m, ref);
    const V3 dl = cross(m, fl);
    return (dmdt * -GYRO + dl * -stt * L.kappa + fl * stt * L.beta) * conv;
  } else if (L.includeSOT) {
    double Hdl = 0, Hfl = 0, Hdl2 = 0, Hfl2 = 0;
    if (L.dynamicSOT) {
      Hdl = s.slot(CMTJB_SLOT_DL, t);
      Hfl = s.slot(CMTJB_SLOT_FL, t);
      Hdl2 = s.slot(CMTJB_SLOT_DL2, t);
      Hfl2 = s.slot(CMTJB_SLOT_FL2, t);
    } else {
      const double I = s.slot(CMTJB_SLOT_CURRENT, t);
      s.I_log = I;
      Hdl = L.dampingLikeTorque * I;
      Hfl = L.fieldLikeTorque * I;
    }
    s.cHfl = ref * (Hfl - L.damping * Hdl); /* core/junction.hpp:786-787 */
    s.cHdl = ref * (Hdl + L.damping * Hfl);
    const V3 ref2 = arr3(L.secondaryReferenceLayer);
    const V3 cm1 = cross(m, ref);
    const V3 ccm1 = cross(m, cm1);
    const V3 cm2 = cross(m, ref2);
    const V3 ccm2 = cross(m, cm2);
    const V3 fl1 = cm1 * (Hfl - L.damping * Hdl);
    const V3 dl1 = ccm1 * (Hdl + L.damping * Hfl);
    const V3 fl2 = cm2 * (Hfl2 - L.damping * Hdl2);
    const V3 dl2 = ccm2 * (Hdl2 + L.damping * Hfl2);
    return (dmdt + fl1 + dl1 + fl2 + dl2) * -GYRO * conv;
  }
  return dmdt * -GYRO * conv;
}

/* core/junction.hpp:855-864 */
V3 llg_rhs(const LayerSim &s, double t, V3 m, V3 bottom, V3 top) {
  return solve_llg(s, t, m, bottom, top, heff(s, t, m, bottom, top));
}

/* core/junction.hpp:824-831 */
V3 stochastic_torque(const LayerSim &s, V3 m, V3 dW) {
  const double conv = -s.J->GYRO / (1. + std::pow(s.L->damping, 2));
  const V3 c1 = cross(m, dW);
  const V3 c2 = cross(m, c1);
  return (c1 + c2 * s.L->damping) * conv;
}

/* core/junction.hpp:1111-1128 + :1130-1136: the step's noise vector dW */
V3 draw_noise(LayerSim &s, double t, int64_t iter) {
  V3 dW = v3(0, 0, 0);
  if (s.L->temperatureSet) {
    const double T = s.slot(CMTJB_SLOT_TEMPERATURE, t);
    const double V = s.L->cellSurface * s.L->thickness;
    const double sigma = std::sqrt((2 * s.L->damping * s.J->BOLTZMANN_CONST * T) / (s.L->Ms * V * s.J->GYRO));
    const double nx = s.normal(s.rng), ny = s.normal(s.rng), nz = s.normal(s.rng);
    dW = v3(nx, ny, nz) * sigma;
  }
  if (s.L->noiseScale != 0 && s.onef) {
    const double *b = s.onef + iter * s.onef_iter_stride;
    dW = dW + v3(b[0], b[s.onef_comp_stride], b[2 * s.onef_comp_stride]);
  }
  return dW;
}

/* core/junction.hpp:1017-1041 */
bool rk4_layer(LayerSim &s, double t, double dt, V3 bottom, V3 top) {
  const V3 m = s.mag;
  const V3 k1 = llg_rhs(s, t, m, bottom, top) * dt;
  const V3 k2 = llg_rhs(s, t + 0.5 * dt, m + k1 * 0.5, bottom, top) * dt;
  const V3 k3 = llg_rhs(s, t + 0.5 * dt, m + k2 * 0.5, bottom, top) * dt;
  const V3 k4 = llg_rhs(s, t + dt, m + k3, bottom, top) * dt;
  V3 mn = m + (k1 + (k2 * 2.0) + (k3 * 2.0) + k4) / 6.0;
  s.mag = normalized(mn);
  return !std::isnan(s.mag.x);
}

/* core/junction.hpp:866-1002.  Returns accepted; *nan set on NaN magnetisation. */
bool dopri_layer(LayerSim &s, double t, double &h, V3 bottom, V3 top, bool *nan) {
  const cmtjb_layer_t &L = *s.L;
  const V3 m = s.mag;
  const double c2 = 1.0 / 5.0, c3 = 3.0 / 10.0, c4 = 4.0 / 5.0, c5 = 8.0 / 9.0, c6 = 1.0, c7 = 1.0;
  const V3 k1 = llg_rhs(s, t, m, bottom, top) * h;
  const V3 k2 = llg_rhs(s, t + c2 * h, m + k1 * (c2), bottom, top) * h;
  const V3 k3 = llg_rhs(s, t + c3 * h, m + k1 * (3.0 / 40.0) + k2 * (9.0 / 40.0), bottom, top) * h;
  const V3 k4 = llg_rhs(s, t + c4 * h, m + k1 * (44.0 / 45.0) + k2 * (-56.0 / 15.0) + k3 * (32.0 / 9.0),
                        bottom, top) * h;
  const V3 k5 = llg_rhs(s, t + c5 * h,
                        m + k1 * (19372.0 / 6561.0) + k2 * (-25360.0 / 2187.0) + k3 * (64448.0 / 6561.0) +
                            k4 * (-212.0 / 729.0),
                        bottom, top) * h;
  const V3 k6 = llg_rhs(s, t + c6 * h,
                        m + k1 * (9017.0 / 3168.0) + k2 * (-355.0 / 33.0) + k3 * (46732.0 / 5247.0) +
                            k4 * (49.0 / 176.0) + k5 * (-5103.0 / 18656.0),
                        bottom, top) * h;
  const V3 k7 = llg_rhs(s, t + c7 * h,
                        m + k1 * (35.0 / 384.0) + k3 * (500.0 / 1113.0) + k4 * (125.0 / 192.0) +
                            k5 * (-2187.0 / 6784.0) + k6 * (11.0 / 84.0),
                        bottom, top) * h;
  V3 m5 = m + k1 * (35.0 / 384.0) + k3 * (500.0 / 1113.0) + k4 * (125.0 / 192.0) + k5 * (-2187.0 / 6784.0) +
          k6 * (11.0 / 84.0);
  const V3 m4 = m + k1 * (5179.0 / 57600.0) + k3 * (7571.0 / 16695.0) + k4 * (393.0 / 640.0) +
                k5 * (-92097.0 / 339200.0) + k6 * (187.0 / 2100.0) + k7 * (1.0 / 40.0);
  const double err = vlen(m5 - m4);
  const double tol = L.abs_tol + L.rel_tol * vlen(m);
  double ratio = err / tol;
  if (tol < 1e-15 || std::isnan(ratio) || std::isinf(ratio)) ratio = 10.0;
  double factor;
  if (L.use_pid_control) {
    const double e_n = std::log(ratio);
    s.integral_error += e_n;
    factor = std::exp(L.kp * e_n + L.ki * s.integral_error + L.kd * (e_n - s.prev_error_ratio));
    s.prev_error_ratio = e_n;
  } else {
    factor = L.safety_factor * std::pow(1.0 / ratio, 0.2);
  }
  if (std::isnan(factor) || std::isinf(factor)) factor = L.min_factor;
  factor = std::min(factor, L.max_factor);
  factor = std::max(factor, L.min_factor);
  double nh = h * factor;
  if (std::isnan(nh) || std::isinf(nh) || nh <= 0) nh = h * L.min_factor;
  if (ratio <= 1.0) {
    s.mag = normalized(m5);
    if (std::isnan(s.mag.x)) *nan = true;
    h = nh;
    return true;
  }
  h = nh;
  return false;
}

struct JunctionSim {
  const cmtjb_junction_t *J;
  std::vector<LayerSim> layers;
  int n;
};

/* core/junction.hpp:1668-1707 (fixed-step use: every layer accepts) */
bool step_rk4(JunctionSim &j, double t, double dt) {
  bool ok = true;
  if (j.n == 1) return rk4_layer(j.layers[0], t, dt, v3(0, 0, 0), v3(0, 0, 0));
  std::vector<V3> snap(j.n + 2, v3(0, 0, 0));
  for (int i = 0; i < j.n; i++) snap[i + 1] = j.layers[i].mag;
  for (int i = 0; i < j.n; i++) {
    if (!rk4_layer(j.layers[i], t, dt, snap[i], snap[i + 2])) {
      ok = false;
      break; /* the reference throws here */
    }
  }
  return ok;
}

/* core/junction.hpp:1709-1754 */
void step_euler_heun(JunctionSim &j, double t, double dt, int64_t iter) {
  std::vector<V3> mPrime(j.n);
  for (int i = 0; i < j.n; i++) {
    LayerSim &s = j.layers[i];
    const V3 dW = draw_noise(s, t, iter);
    const V3 bottom = (i == 0) ? v3(0, 0, 0) : j.layers[i - 1].mag;
    const V3 top = (i == j.n - 1) ? v3(0, 0, 0) : j.layers[i + 1].mag;
    const V3 fn = llg_rhs(s, t, s.mag, bottom, top);
    const V3 gn = stochastic_torque(s, s.mag, dW);
    const V3 mNext = s.mag + gn * std::sqrt(dt);
    const V3 gnP = stochastic_torque(s, mNext, dW);
    mPrime[i] = s.mag + fn * dt + 0.5 * (gn + gnP) * std::sqrt(dt);
  }
  for (int i = 0; i < j.n; i++) j.layers[i].mag = normalized(mPrime[i]);
}

/* core/junction.hpp:1756-1815 */
void step_heun(JunctionSim &j, double t, double dt, int64_t iter) {
  std::vector<V3> fn(j.n), gn(j.n), dW(j.n), mNext(j.n);
  for (int i = 0; i < j.n; i++) {
    LayerSim &s = j.layers[i];
    const V3 bottom = (i == 0) ? v3(0, 0, 0) : j.layers[i - 1].mag;
    const V3 top = (i == j.n - 1) ? v3(0, 0, 0) : j.layers[i + 1].mag;
    fn[i] = llg_rhs(s, t, s.mag, bottom, top);
    dW[i] = draw_noise(s, t, iter);
    gn[i] = stochastic_torque(s, s.mag, dW[i]);
    mNext[i] = s.mag + fn[i] * dt + gn[i] * std::sqrt(dt);
  }
  for (int i = 0; i < j.n; i++) {
    LayerSim &s = j.layers[i];
    const V3 bottom = (i == 0) ? v3(0, 0, 0) : mNext[i - 1];
    const V3 top = (i == j.n - 1) ? v3(0, 0, 0) : mNext[i + 1];
    const V3 m = s.mag + 0.5 * dt * (fn[i] + llg_rhs(s, t + dt, mNext[i], bottom, top)) +
                 0.5 * (gn[i] + stochastic_torque(s, mNext[i], dW[i])) * std::sqrt(dt);
    s.mag = normalized(m);
  }
}

/* core/junction.hpp:1835-1867 */
void write_sample(const JunctionSim &j, double *log, int64_t n_samples, int64_t B, int64_t sample, int64_t inst) {
  auto put = [&](int ch, double v) { log[(ch * n_samples + sample) * B + inst] = v; };
  int ch = 0;
  for (int i = 0; i < j.n; i++) {
    put(ch++, j.layers[i].mag.x);
    put(ch++, j.layers[i].mag.y);
    put(ch++, j.layers[i].mag.z);
  }
  const cmtjb_junction_t &J = *j.J;
  if (J.mr_mode == CMTJB_MR_CLASSIC) {
    const double c = (j.n == 1) ? dot(j.layers[0].mag, arr3(j.layers[0].L->referenceLayer))
                                : dot(j.layers[0].mag, j.layers[1].mag);
    put(ch++, J.Rp + (((J.Rap - J.Rp) / 2.0) * (1.0 - c)));
  } else if (J.mr_mode == CMTJB_MR_STRIP) {
    double Rx_acc = 0.0, Ry_acc = 0.0;
    for (int i = 0; i < j.n; i++) {
      const cmtjb_layer_t &L = *j.layers[i].L;
      const V3 m = j.layers[i].mag;
      const double Rx = L.Rx0 + L.AMR_X * (m.x * m.x) + L.SMR_X * (m.y * m.y);
      const double Ry = L.Ry0 + 0.5 * L.AHE * m.z + (L.AMR_Y + L.SMR_Y) * m.x * m.y;
      Rx_acc += 1. / Rx;
      Ry_acc += 1. / Ry;
    }
    put(ch++, 1 / Rx_acc);
    put(ch++, 1 / Ry_acc);
    put(ch++, 0.);
  }
}

/* core/junction.hpp:1584-1613 (calculateEnergies); channels follow the LOG_FULL ones */
void write_energy_sample(const JunctionSim &j, double *log, int64_t n_samples, int64_t B, int64_t sample, int64_t inst, int ch) {
  auto put = [&](double v) { log[((int64_t)ch++ * n_samples + sample) * B + inst] = v; };
  auto put3 = [&](V3 v) {
    put(v.x);
    put(v.y);
    put(v.z);
  };
  for (int i = 0; i < j.n; i++) {
    const LayerSim &s = j.layers[i];
    put(s.K_log);
    put(s.K2_log);
    put(s.Jbottom_log);
    put(s.Jtop_log);
    put(s.I_log);
    put3(s.cHext);
    put3(s.cHIEC);
    put3(s.cHAnis);
    put3(s.cHdemag);
    put3(v3(0, 0, 0)); /* Hfluctuation is never passed non-default (SURVEY 8a note 6) */
    if (s.L->includeSOT) {
      put3(s.cHfl);
      put3(s.cHdl);
    }
  }
}

int energy_channels(const cmtjb_junction_t *jn, const cmtjb_layer_t *layers) {
  int n = 0;
  for (int i = 0; i < jn->n_layers; i++) n += layers[i].includeSOT ? CMTJB_ENERGY_CHANNELS_SOT : CMTJB_ENERGY_CHANNELS;
  return n;
}

int validate(const cmtjb_junction_t *jn, const cmtjb_layer_t *layers, const cmtjb_run_args_t *a, cmtjb_plan_t *p) {
  if (!jn || !layers || !a || !p) return fail(CMTJB_EINVAL, "null argument");
  if (jn->n_layers <= 0) return fail(CMTJB_EINVAL, "Passed a zero length Layer vector!");
  if (jn->n_layers > CMTJB_MAX_LAYERS) return fail(CMTJB_ENOTSUP, "too many layers for this build");
  if (a->timeStep > a->writeFrequency)
    return fail(CMTJB_ERUNTIME, "The time step cannot be larger than write frequency!");
  std::memset(p, 0, sizeof(*p));
  p->n_layers = jn->n_layers;
  /* core/junction.hpp:1956-1959: unsigned casts of the FP64 quotients */
  p->total_iterations = (int64_t) static_cast<unsigned int>(a->totalTime / a->timeStep);
  p->write_every = (int64_t) static_cast<unsigned int>(a->writeFrequency / a->timeStep);
  int mode = a->solverMode;
  bool any_custom = false;
  for (int i = 0; i < jn->n_layers; i++) {
    const cmtjb_layer_t &L = layers[i];
    if (L.temperatureSet || L.noiseScale != 0) {
      if (mode != CMTJB_HEUN && mode != CMTJB_EULER_HEUN) {
        mode = CMTJB_EULER_HEUN;
        p->solver_overridden = 1;
      }
    }
    if (L.temperatureSet && L.cellSurface * L.thickness == 0.0)
      return fail(CMTJB_ERUNTIME, "Cell surface cannot be 0 during temp. calculations!");
    if (L.Ms * L.thickness == 0) return fail(CMTJB_ERUNTIME, "Failed to divide vector by zero!");
    for (int s = 0; s < CMTJB_NSLOTS; s++)
      if (L.drivers[s].update == CMTJB_UPD_CUSTOM) any_custom = true;
  }
  p->effective_solver = mode;
  if (any_custom && mode == CMTJB_DORMAND_PRINCE)
    return fail(CMTJB_ENOTSUP, "custom (Python callback) drivers need a fixed-step solver: Dormand-Prince evaluates drivers at adaptive times");
  p->n_channels = 3 * jn->n_layers + (jn->mr_mode == CMTJB_MR_CLASSIC ? 1 : jn->mr_mode == CMTJB_MR_STRIP ? 3 : 0);
  if (a->log_mode == CMTJB_LOG_ENERGY) p->n_channels += energy_channels(jn, layers);
  if (mode == CMTJB_DORMAND_PRINCE) {
    if (jn->n_layers > 1)
      return fail(CMTJB_ENOTSUP,
                  "Dormand-Prince with more than one layer diverges in the reference (dt=-1 loop) and is not supported");
    p->variable_samples = 1;
    p->n_samples = (int64_t)(a->totalTime / a->writeFrequency) + 2;
  } else {
    if (p->write_every <= 0) return fail(CMTJB_ERUNTIME, "The time step cannot be larger than write frequency!");
    const int64_t first = a->start_iteration;
    /* samples at iterations i in [first, total) with i % write_every == 0 */
    const int64_t total = p->total_iterations;
    auto count_upto = [&](int64_t n) { return n <= 0 ? 0 : (n - 1) / p->write_every + 1; };
    p->n_samples = count_upto(total) - count_upto(first);
  }
  std::snprintf(p->variant, sizeof(p->variant), "oracle-scalar");
  return CMTJB_OK;
}

} // namespace

extern "C" {

const char *oracle_last_error(void) { return g_err.c_str(); }

int oracle_plan(const cmtjb_junction_t *jn, const cmtjb_layer_t *layers, const cmtjb_run_args_t *args,
                cmtjb_plan_t *plan) {
  return validate(jn, layers, args, plan);
}

int oracle_eval_driver(const cmtjb_driver_t *driver, const double *times, int64_t n, double *out) {
  for (int64_t i = 0; i < n; i++) out[i] = driver_value(*driver, times[i]);
  return CMTJB_OK;
}

/* Same contract as cmtjb_run_host (include/cmtj_b200.h); restates
 * Junction::runSimulation, core/junction.hpp:1948-2056. */
int oracle_run_host(const cmtjb_junction_t *jn, const cmtjb_layer_t *layers, const cmtjb_override_t *ovr,
                    int32_t n_ovr, const cmtjb_run_args_t *args, double *m_state, double *log, int32_t *status,
                    int32_t *n_samples_out, double *time_log) {
  cmtjb_plan_t plan;
  int rc = validate(jn, layers, args, &plan);
  if (rc) return rc;
  const int64_t B = args->n_instances;
  const int n = jn->n_layers;
  const double dt = args->timeStep;
  const int64_t NS = plan.n_samples;
  const bool do_log = (args->log_mode == CMTJB_LOG_FULL || args->log_mode == CMTJB_LOG_ENERGY) && log;
  const bool do_energy = args->log_mode == CMTJB_LOG_ENERGY;
  const int base_ch = plan.n_channels - (do_energy ? energy_channels(jn, layers) : 0);
  for (int k = 0; k < n_ovr; k++) {
    if (ovr[k].layer >= n || ovr[k].slot < 0 || ovr[k].slot >= CMTJB_NSLOTS || ovr[k].field < 0 ||
        ovr[k].field > CMTJB_FIELD_STEADYTIME)
      return fail(CMTJB_EINVAL, "override out of range");
  }
  int worst = CMTJB_OK;
  for (int64_t inst = 0; inst < B; inst++) {
    JunctionSim j;
    j.J = jn;
    j.n = n;
    j.layers.resize(n);
    for (int l = 0; l < n; l++) {
      LayerSim &s = j.layers[l];
      s.L = &layers[l];
      s.J = jn;
      std::memcpy(s.drv, layers[l].drivers, sizeof(s.drv));
      s.mag = v3(m_state[(l * 3 + 0) * B + inst], m_state[(l * 3 + 1) * B + inst], m_state[(l * 3 + 2) * B + inst]);
      std::seed_seq seq{(uint32_t)(args->seed), (uint32_t)(args->seed >> 32),
                        (uint32_t)(args->instance_offset + inst), (uint32_t)l, 0x5eedu};
      s.rng.seed(seq);
      s.onef = args->onef_buffer ? args->onef_buffer + ((int64_t)(l * 3) * plan.total_iterations) * B + inst : nullptr;
      s.onef_comp_stride = plan.total_iterations * B;
      s.onef_iter_stride = B;
      s.integral_error = layers[l].integral_error;
      s.prev_error_ratio = layers[l].prev_error_ratio;
    }
    for (int k = 0; k < n_ovr; k++) {
      for (int l = 0; l < n; l++) {
        if (ovr[k].layer >= 0 && ovr[k].layer != l) continue;
        double *f = &j.layers[l].drv[ovr[k].slot].constantValue;
        f[ovr[k].field] = ovr[k].values[inst];
      }
    }
    int st = 0;
    if (plan.effective_solver != CMTJB_DORMAND_PRINCE) {
      int64_t sample = 0;
      for (int64_t i = args->start_iteration; i < plan.total_iterations; i++) {
        const double t = i * dt;
        g_custom.table = args->custom_table;
        g_custom.n_custom = args->n_custom;
        g_custom.step = i;
        g_custom.t0 = t;
        g_custom.dt = dt;
        bool ok = true;
        switch (plan.effective_solver) {
        case CMTJB_RK4: ok = step_rk4(j, t, dt); break;
        case CMTJB_HEUN: step_heun(j, t, dt, i); break;
        default: step_euler_heun(j, t, dt, i); break;
        }
        if (!ok) {
          st = 1;
          break;
        }
        if (!(i % plan.write_every)) {
          if (do_log) write_sample(j, log, NS, B, sample, inst);
          if (do_log && do_energy) write_energy_sample(j, log, NS, B, sample, inst, base_ch);
          sample++;
        }
      }
    } else {
      /* core/junction.hpp:1983-2044 */
      double t = 0.0, next_write = 0.0, h = dt;
      int consecutive_rejections = 0;
      const int max_rejections = 10;
      const double min_h = 1e-15;
      int64_t sample = 0;
      LayerSim &s = j.layers[0];
      while (t < args->totalTime) {
        if (t + h > args->totalTime) h = args->totalTime - t;
        consecutive_rejections = 0;
        if (std::isnan(h) || std::isinf(h) || h <= 0) {
          h = dt;
          consecutive_rejections++;
        }
        bool nan = false;
        bool accepted = dopri_layer(s, t, h, v3(0, 0, 0), v3(0, 0, 0), &nan);
        if (nan) {
          st = 1;
          break;
        }
        if (!accepted && (h < min_h || consecutive_rejections >= max_rejections)) {
          accepted = true;
          consecutive_rejections = 0;
        }
        if (accepted) {
          t += h;
          consecutive_rejections = 0;
          if (t >= next_write) {
            if (do_log && sample < NS) {
              write_sample(j, log, NS, B, sample, inst);
              if (do_energy) write_energy_sample(j, log, NS, B, sample, inst, base_ch);
              if (time_log) time_log[sample * B + inst] = t;
            }
            sample++;
            next_write = ((static_cast<unsigned int>(t / args->writeFrequency) + 1) * args->writeFrequency);
          }
        } else {
          consecutive_rejections++;
        }
      }
      if (n_samples_out) n_samples_out[inst] = (int32_t)std::min<int64_t>(sample, NS);
    }
    for (int l = 0; l < n; l++) {
      m_state[(l * 3 + 0) * B + inst] = j.layers[l].mag.x;
      m_state[(l * 3 + 1) * B + inst] = j.layers[l].mag.y;
      m_state[(l * 3 + 2) * B + inst] = j.layers[l].mag.z;
    }
    if (status) status[inst] = st;
    if (st) worst = CMTJB_ENAN;
  }
  if (worst == CMTJB_ENAN) return fail(CMTJB_ENAN, "NAN magnetisation");
  return CMTJB_OK;
}

} /* extern "C" */

/* ------------------------------------------------------------------------------------------------
 * Coupled systems (include/cmtj_b200.h "coupled systems"): Stack and GroupInteraction.
 * ---------------------------------------------------------------------------------------------- */
namespace {

struct SystemSim {
  std::vector<JunctionSim> j;
  int nl = 0;
  int solver = CMTJB_RK4;
  cmtjb_plan_t plan;
};

int build_system(int n, const cmtjb_junction_t *jns, const cmtjb_layer_t *layers, const cmtjb_override_t *ovr, int n_ovr,
                 const cmtjb_run_args_t *a, const double *m_state, int64_t E, int64_t inst, SystemSim &S) {
  S.nl = jns[0].n_layers;
  S.j.resize(n);
  cmtjb_run_args_t aj = *a;
  aj.solverMode = CMTJB_RK4; /* stack.hpp:50, reservoir.hpp:200 */
  aj.onef_buffer = nullptr;
  for (int k = 0; k < n; k++) {
    if (jns[k].n_layers != S.nl) return fail(CMTJB_ENOTSUP, "junctions of a coupled system must all have the same number of layers in this build");
    cmtjb_plan_t p;
    int rc = validate(&jns[k], layers + (size_t)k * S.nl, &aj, &p);
    if (rc) return rc;
    for (int l = 0; l < S.nl; l++)
      for (int s = 0; s < CMTJB_NSLOTS; s++)
        if (layers[(size_t)k * S.nl + l].drivers[s].update == CMTJB_UPD_CUSTOM)
          return fail(CMTJB_ENOTSUP, "custom (Python callback) drivers are not supported inside a coupled system");
    if (k == 0) {
      S.plan = p;
      S.solver = p.effective_solver;
    } else if (p.effective_solver != S.solver) {
      return fail(CMTJB_ERUNTIME, "Junctions have different solver modes! Set the same solver mode for all junctions explicitly. Do not mix stochastic and deterministic solvers!");
    }
    JunctionSim &J = S.j[k];
    J.J = &jns[k];
    J.n = S.nl;
    J.layers.resize(S.nl);
    const int64_t e = inst * n + k;
    for (int l = 0; l < S.nl; l++) {
      LayerSim &s = J.layers[l];
      s.L = &layers[(size_t)k * S.nl + l];
      s.J = &jns[k];
      std::memcpy(s.drv, s.L->drivers, sizeof(s.drv));
      s.mag = v3(m_state[(l * 3 + 0) * E + e], m_state[(l * 3 + 1) * E + e], m_state[(l * 3 + 2) * E + e]);
      std::seed_seq seq{(uint32_t)(a->seed), (uint32_t)(a->seed >> 32), (uint32_t)(a->instance_offset + e), (uint32_t)l, 0x5eedu};
      s.rng.seed(seq);
      s.onef = nullptr;
      s.onef_comp_stride = s.onef_iter_stride = 0;
      s.integral_error = s.L->integral_error;
      s.prev_error_ratio = s.L->prev_error_ratio;
    }
    for (int q = 0; q < n_ovr; q++) {
      if (ovr[q].layer == CMTJB_LAYER_SYSTEM) continue;
      for (int l = 0; l < S.nl; l++) {
        if (ovr[q].layer >= 0 && ovr[q].layer != l) continue;
        double *f = &J.layers[l].drv[ovr[q].slot].constantValue;
        f[ovr[q].field] = ovr[q].values[e];
      }
    }
  }
  return CMTJB_OK;
}

bool step_junction(JunctionSim &J, int solver, double t, double dt, int64_t iter) {
  if (solver == CMTJB_RK4) return step_rk4(J, t, dt);
  step_euler_heun(J, t, dt, iter);
  return true;
}

/* Junction::getMagnetoresistance()[0], CLASSIC (core/junction.hpp:1865-1889) */
double classic_resistance(const JunctionSim &J) {
  const cmtjb_junction_t &H = *J.J;
  const double c = (J.n == 1) ? dot(J.layers[0].mag, arr3(J.layers[0].L->referenceLayer)) : dot(J.layers[0].mag, J.layers[1].mag);
  return H.Rp + (((H.Rap - H.Rp) / 2.0) * (1.0 - c));
}

void store_system(const SystemSim &S, double *m_state, int64_t E, int64_t inst) {
  const int n = (int)S.j.size();
  for (int k = 0; k < n; k++)
    for (int l = 0; l < S.nl; l++) {
      const int64_t e = inst * n + k;
      m_state[(l * 3 + 0) * E + e] = S.j[k].layers[l].mag.x;
      m_state[(l * 3 + 1) * E + e] = S.j[k].layers[l].mag.y;
      m_state[(l * 3 + 2) * E + e] = S.j[k].layers[l].mag.z;
    }
}

int stack_checks(const cmtjb_stack_t *st, const cmtjb_junction_t *jns, const cmtjb_layer_t *layers) {
  if (!st || !jns || !layers) return fail(CMTJB_EINVAL, "null argument");
  const int n = st->n_junctions;
  if (n < 2) return fail(CMTJB_ERUNTIME, "Stack must have at least 2 junctions!");
  for (int k = 0; k < n; k++)
    if (jns[k].mr_mode != CMTJB_MR_CLASSIC)
      return fail(CMTJB_ERUNTIME, "Junction has a non-classic magnetoresitance mode! Define the junction with Rp and Rap resistance values.");
  for (int k = 0; k < n; k++)
    if (jns[k].n_layers >= 3) return fail(CMTJB_ERUNTIME, "At least one junction has more than 2 layers! It's not supported now.");
  if (!st->coupling) return fail(CMTJB_ERUNTIME, "Coupling strength is not set!");
  const int nl = jns[0].n_layers;
  if (!st->delayed && nl != 2) return fail(CMTJB_ERUNTIME, "Non delayed coupling is only supported for 2 layer stacks!");
  if (nl == 1) {
    const double *p0 = layers[st->top_layer].referenceLayer;
    for (int k = 1; k < n; k++) {
      const double *pk = layers[(size_t)k * nl + st->top_layer].referenceLayer;
      if (pk[0] != p0[0] || pk[1] != p0[1] || pk[2] != p0[2]) return fail(CMTJB_ERUNTIME, "Polarisation vectors are not equal in stack");
    }
    if (vlen(arr3(p0)) == 0) return fail(CMTJB_ERUNTIME, "Polarisation is not set!");
  }
  return CMTJB_OK;
}

/* reservoir.hpp:79-97 */
V3 dipole_kanao(V3 r1, V3 r2, const LayerSim &l2) {
  const V3 rij = r1 - r2;
  const double len = vlen(rij);
  const double r3 = std::pow(len, 3), r5 = std::pow(len, 5);
  const V3 m1 = l2.mag;
  const double V = l2.L->thickness * l2.L->cellSurface;
  const double prefactor = (l2.L->Ms / l2.J->MAGNETIC_PERMEABILITY) * V;
  return prefactor * ((3 * dot(m1, rij) * rij) / r5 - m1 / r3);
}
/* reservoir.hpp:36-56, :108-120 */
V3 dipole_nomura(V3 r1, V3 r2, const LayerSim &l2) {
  const V3 rij = r1 - r2;
  const double len = vlen(rij);
  const double r_mag = std::pow(len, 2);
  const double mult = 3 / (4 * M_PI * std::pow(len, 5));
  const V3 row0 = v3(std::pow(rij.x, 2) - (r_mag / 3), rij.x * rij.y, rij.x * rij.z) * mult;
  const V3 row1 = v3(rij.x * rij.y, std::pow(rij.y, 2) - (r_mag / 3), rij.y * rij.z) * mult;
  const V3 row2 = v3(rij.x * rij.z, rij.y * rij.z, std::pow(rij.z, 2) - (r_mag / 3)) * mult;
  const double T[9] = {row0.x, row0.y, row0.z, row1.x, row1.y, row1.z, row2.x, row2.y, row2.z};
  const double V = l2.L->thickness * l2.L->cellSurface;
  return tensor_interaction(l2.mag, T, l2.L->Ms * V, l2.J->MAGNETIC_PERMEABILITY);
}

} // namespace

extern "C" {

int oracle_stack_plan(const cmtjb_stack_t *st, const cmtjb_junction_t *jns, const cmtjb_layer_t *layers,
                      const cmtjb_run_args_t *args, cmtjb_coupled_plan_t *plan) {
  int rc = stack_checks(st, jns, layers);
  if (rc) return rc;
  if (args->timeStep > args->writeFrequency) return fail(CMTJB_ERUNTIME, "The time step cannot be larger than write frequency!");
  SystemSim S;
  std::vector<double> m0((size_t)jns[0].n_layers * 3 * st->n_junctions, 0.0);
  rc = build_system(st->n_junctions, jns, layers, nullptr, 0, args, m0.data(), st->n_junctions, 0, S);
  if (rc) return rc;
  std::memset(plan, 0, sizeof(*plan));
  plan->effective_solver = S.solver;
  plan->solver_overridden = S.plan.solver_overridden;
  plan->n_layers = S.nl;
  plan->n_junction_channels = S.plan.n_channels;
  plan->n_system_channels = 1 + st->n_junctions + (st->n_junctions - 1);
  plan->total_iterations = S.plan.total_iterations;
  plan->write_every = S.plan.write_every;
  plan->n_samples = S.plan.n_samples;
  return CMTJB_OK;
}

/* Stack::runSimulationKCL / NonKCL (core/stack.hpp:317-423) */
int oracle_stack_run_host(const cmtjb_stack_t *st, const cmtjb_junction_t *jns, const cmtjb_layer_t *layers,
                          const cmtjb_override_t *ovr, int32_t n_ovr, const cmtjb_run_args_t *args, double *m_state,
                          double *jlog, double *slog, int32_t *status) {
  cmtjb_coupled_plan_t plan;
  int rc = oracle_stack_plan(st, jns, layers, args, &plan);
  if (rc) return rc;
  const int N = st->n_junctions;
  const int64_t n_inst = args->n_instances, E = n_inst * N, NS = plan.n_samples;
  const double dt = args->timeStep;
  const bool do_log = args->log_mode == CMTJB_LOG_FULL;
  const bool two = plan.n_layers == 2;
  int worst = CMTJB_OK;
  for (int64_t inst = 0; inst < n_inst; inst++) {
    SystemSim S;
    rc = build_system(N, jns, layers, ovr, n_ovr, args, m_state, E, inst, S);
    if (rc) return rc;
    cmtjb_driver_t cur = st->currentDriver;
    for (int q = 0; q < n_ovr; q++)
      if (ovr[q].layer == CMTJB_LAYER_SYSTEM) (&cur.constantValue)[ovr[q].field] = ovr[q].values[inst];
    std::vector<double> Rt(N), It(N), Ct(N - 1);
    std::vector<V3> fm(N), fp(N);
    const V3 pol = arr3(layers[st->top_layer].referenceLayer);
    auto coupling = [&](int order, V3 m1, V3 m2, V3 p1, V3 p2) {
      const double a = dot(m1, p1), b = dot(m2, p2);
      return st->coupling[order] * (st->kind == CMTJB_STACK_SERIES ? (a + b) : (a - b));
    };
    auto phase_offset = [&](int order) { return st->kind == CMTJB_STACK_SERIES ? st->phaseOffset * order : st->phaseOffset; };
    auto top = [&](int k) { return S.j[k].layers[st->top_layer].mag; };
    auto bot = [&](int k) { return S.j[k].layers[st->bottom_layer].mag; };
    int64_t sample = 0;
    bool nan = false;
    for (int64_t i = args->start_iteration; i < plan.total_iterations && !nan; i++) {
      const double t = i * dt;
      for (int k = 0; k < N; k++) {
        fm[k] = top(k);
        fp[k] = two ? bot(k) : pol;
      }
      auto exec = [&](int k, const cmtjb_driver_t &local) {
        for (int l = 0; l < S.nl; l++) S.j[k].layers[l].drv[CMTJB_SLOT_CURRENT] = local;
        if (!step_junction(S.j[k], S.solver, t, dt, i)) nan = true;
        Rt[k] = classic_resistance(S.j[k]);
        It[k] = driver_value(local, t);
      };
      if (st->use_kcl) {
        const double I0 = driver_value(cur, t);
        double total = I0;
        for (int k = 0; k < N - 1; k++) {
          Ct[k] = coupling(k, fm[k], fm[k + 1], fp[k], fp[k + 1]);
          total += Ct[k] * I0;
        }
        for (int k = 0; k < N && !nan; k++) {
          cmtjb_driver_t local;
          std::memset(&local, 0, sizeof(local));
          local.update = CMTJB_UPD_CONSTANT;
          local.constantValue = total;
          local.frequency = local.period = local.cycle = local.timeStart = local.timeStop = local.edgeTime = local.steadyTime = -1;
          local.phase += phase_offset(k);
          exec(k, local);
        }
      } else {
        double eff = 1;
        for (int k = 0; k < N && !nan; k++) {
          if (k > 0) {
            if (st->delayed) eff *= (1 + coupling(k - 1, fm[k - 1], fm[k], fp[k - 1], fp[k]));
            else eff *= (1 + coupling(k - 1, top(k - 1), top(k), bot(k - 1), bot(k)));
            Ct[k - 1] = eff;
          }
          cmtjb_driver_t local = cur;
          local.constantValue = cur.constantValue * eff;
          local.amplitude = cur.amplitude * eff;
          local.phase += phase_offset(k);
          exec(k, local);
        }
      }
      if (nan) break;
      if (!(i % plan.write_every)) {
        if (do_log && slog) {
          double res;
          if (st->kind == CMTJB_STACK_SERIES) {
            res = 0.0;
            for (int k = 0; k < N; k++) res += Rt[k];
          } else {
            double inv = 0.0;
            for (int k = 0; k < N; k++) inv += 1.0 / Rt[k];
            res = 1. / inv;
          }
          auto put = [&](int ch, double v) { slog[((int64_t)ch * NS + sample) * n_inst + inst] = v; };
          put(0, res);
          for (int k = 0; k < N; k++) put(1 + k, It[k]);
          for (int k = 0; k < N - 1; k++) put(1 + N + k, Ct[k]);
        }
        if (do_log && jlog)
          for (int k = 0; k < N; k++) write_sample(S.j[k], jlog, NS, E, sample, inst * N + k);
        sample++;
      }
    }
    store_system(S, m_state, E, inst);
    for (int k = 0; k < N; k++)
      if (status) status[inst * N + k] = nan ? 1 : 0;
    if (nan) worst = CMTJB_ENAN;
  }
  if (worst) return fail(CMTJB_ENAN, "NAN magnetisation");
  return CMTJB_OK;
}

/* CMTJB_INT_TABLE: the tabulated affine interaction c[a][b] + T[a][b] m_b (include/cmtj_b200.h) */
static V3 pair_table_field(const cmtjb_group_t *g, int a, int b, V3 m) {
  const size_t n = (size_t)g->n_junctions;
  const double *t = g->pair_table + (size_t)b * 12 * n + (size_t)a;
  auto e = [&](int k) { return t[(size_t)k * n]; };
  return v3(e(9) + (e(0) * m.x + e(1) * m.y + e(2) * m.z), e(10) + (e(3) * m.x + e(4) * m.y + e(5) * m.z),
            e(11) + (e(6) * m.x + e(7) * m.y + e(8) * m.z));
}

int oracle_group_plan(const cmtjb_group_t *g, const cmtjb_junction_t *jns, const cmtjb_layer_t *layers,
                      const cmtjb_run_args_t *args, cmtjb_coupled_plan_t *plan) {
  if (!g || !jns || !layers || !args || !plan) return fail(CMTJB_EINVAL, "null argument");
  const int n = g->n_junctions;
  if (n <= 0 || !g->coordinates) return fail(CMTJB_ERUNTIME, "Coordinate matrix and junction list cannot be empty!");
  if (args->timeStep > args->writeFrequency) return fail(CMTJB_ERUNTIME, "The time step cannot be larger than write frequency!");
  const double lim = (g->interaction == CMTJB_INT_DIPOLE_NOMURA) ? 1e-14 : 1e-10;
  for (int i = 0; i < n; i++)
    for (int j = i + 1; j < n; j++) {
      const V3 d = arr3(g->coordinates + 3 * i) - arr3(g->coordinates + 3 * j);
      if (d.x == 0 && d.y == 0 && d.z == 0) return fail(CMTJB_ERUNTIME, "Coordinate vectors must be unique!");
      if ((g->interaction == CMTJB_INT_DIPOLE || g->interaction == CMTJB_INT_DIPOLE_NOMURA) && vlen(d) < lim)
        return fail(CMTJB_ERUNTIME, "Points are too close for stable dipole calculation");
    }
  SystemSim S;
  std::vector<double> m0((size_t)jns[0].n_layers * 3 * n, 0.0);
  int rc = build_system(n, jns, layers, nullptr, 0, args, m0.data(), n, 0, S);
  if (rc) return rc;
  std::memset(plan, 0, sizeof(*plan));
  plan->effective_solver = S.solver;
  plan->solver_overridden = S.plan.solver_overridden;
  plan->n_layers = S.nl;
  plan->n_junction_channels = S.plan.n_channels;
  plan->total_iterations = S.plan.total_iterations;
  plan->write_every = S.plan.write_every;
  plan->n_samples = S.plan.n_samples;
  return CMTJB_OK;
}

/* GroupInteraction::runSimulation + stepFunctionalSolver (core/reservoir.hpp:129-152, :185-226) */
int oracle_group_run_host(const cmtjb_group_t *g, const cmtjb_junction_t *jns, const cmtjb_layer_t *layers,
                          const cmtjb_override_t *ovr, int32_t n_ovr, const cmtjb_run_args_t *args, double *m_state,
                          double *jlog, int32_t *status) {
  cmtjb_coupled_plan_t plan;
  int rc = oracle_group_plan(g, jns, layers, args, &plan);
  if (rc) return rc;
  const int N = g->n_junctions;
  const int64_t n_inst = args->n_instances, E = n_inst * N, NS = plan.n_samples;
  const double dt = args->timeStep;
  const bool do_log = args->log_mode == CMTJB_LOG_FULL && jlog;
  int worst = CMTJB_OK;
  for (int64_t inst = 0; inst < n_inst; inst++) {
    SystemSim S;
    rc = build_system(N, jns, layers, ovr, n_ovr, args, m_state, E, inst, S);
    if (rc) return rc;
    int64_t sample = 0;
    bool nan = false;
    for (int64_t i = args->start_iteration; i < plan.total_iterations && !nan; i++) {
      const double t = i * dt;
      for (int a = 0; a < N; a++) {
        V3 H = v3(0, 0, 0);
        for (int b = 0; b < N; b++) {
          if (a == b) continue;
          const V3 ra = arr3(g->coordinates + 3 * a), rb = arr3(g->coordinates + 3 * b);
          const LayerSim &lb = S.j[b].layers[g->top_layer];
          if (g->interaction == CMTJB_INT_DIPOLE) H = H + dipole_kanao(ra, rb, lb);
          else if (g->interaction == CMTJB_INT_DIPOLE_NOMURA) H = H + dipole_nomura(ra, rb, lb);
          else if (g->interaction == CMTJB_INT_TABLE) H = H + pair_table_field(g, a, b, lb.mag);
        }
        LayerSim &la = S.j[a].layers[g->top_layer];
        const double h[3] = {H.x, H.y, H.z};
        for (int c = 0; c < 3; c++) {
          cmtjb_driver_t &d = la.drv[CMTJB_SLOT_HRES_X + c];
          std::memset(&d, 0, sizeof(d));
          d.update = CMTJB_UPD_CONSTANT;
          d.constantValue = h[c];
        }
      }
      for (int a = 0; a < N && !nan; a++)
        if (!step_junction(S.j[a], S.solver, t, dt, i)) nan = true;
      if (nan) break;
      if (!(i % plan.write_every)) {
        if (do_log)
          for (int a = 0; a < N; a++) write_sample(S.j[a], jlog, NS, E, sample, inst * N + a);
        sample++;
      }
    }
    store_system(S, m_state, E, inst);
    for (int a = 0; a < N; a++)
      if (status) status[inst * N + a] = nan ? 1 : 0;
    if (nan) worst = CMTJB_ENAN;
  }
  if (worst) return fail(CMTJB_ENAN, "NAN magnetisation");
  return CMTJB_OK;
}

/* The sharded form of cmtjb_group_run_sharded (include/cmtj_b200.h) with HOST pointers: rank `rank` steps elements
 * [rank*per, rank*per + own) of reservoir.hpp:129-152's loop and reads every partner's top-layer magnetisation from the
 * exchange buffer [world][3][per] that `exchange` all-gathers after each step.  Stand-in for the device library in the
 * world_size-2 gloo test of the host-side plumbing (tests/test_sharded_group.py); `device` and `stream` are ignored. */
typedef int (*oracle_exchange_fn)(void *ctx, double *buf, int64_t count, int64_t step, void *stream);
int oracle_group_run_sharded(const cmtjb_group_t *g, const cmtjb_junction_t *jns, const cmtjb_layer_t *layers,
                             const cmtjb_run_args_t *args, int32_t rank, int32_t world, oracle_exchange_fn exchange, void *ctx,
                             double *m_state, double *jlog, int32_t *status, int device, void *stream) {
  (void)device;
  if (!exchange || world < 1 || rank < 0 || rank >= world) return fail(CMTJB_EINVAL, "bad rank / world / exchange callback");
  if (!args || args->n_instances != 1) return fail(CMTJB_EINVAL, "a sharded array is one instance");
  cmtjb_coupled_plan_t plan;
  int rc = oracle_group_plan(g, jns, layers, args, &plan);
  if (rc) return rc;
  const int N = g->n_junctions, nl = plan.n_layers, top = g->top_layer;
  const int per = (N + world - 1) / world;
  const int e_lo = std::min(N, rank * per), e_hi = std::min(N, e_lo + per), own = e_hi - e_lo;
  const int64_t NS = plan.n_samples;
  const bool do_log = args->log_mode == CMTJB_LOG_FULL && jlog;
  /* the whole array is built so that partners carry their Ms / thickness / surface; only own elements hold real state */
  std::vector<double> m_all((size_t)nl * 3 * N, 0.0);
  for (int r = 0; r < nl * 3; r++)
    for (int a = 0; a < own; a++) m_all[(size_t)r * N + e_lo + a] = m_state[(size_t)r * own + a];
  SystemSim S;
  rc = build_system(N, jns, layers, nullptr, 0, args, m_all.data(), N, 0, S);
  if (rc) return rc;
  std::vector<double> xbuf((size_t)world * 3 * per, 0.0);
  auto publish = [&]() {
    for (int a = 0; a < own; a++) {
      const V3 m = S.j[e_lo + a].layers[top].mag;
      double *q = xbuf.data() + (size_t)rank * 3 * per + a;
      q[0] = m.x, q[per] = m.y, q[2 * per] = m.z;
    }
  };
  auto absorb = [&]() {
    for (int b = 0; b < N; b++) {
      if (b >= e_lo && b < e_hi) continue;
      const double *q = xbuf.data() + (size_t)(b / per) * 3 * per + (b % per);
      S.j[b].layers[top].mag = v3(q[0], q[per], q[2 * per]);
    }
  };
  publish();
  if (exchange(ctx, xbuf.data(), (int64_t)3 * per, -1, stream)) return fail(CMTJB_ERUNTIME, "exchange callback failed");
  absorb();
  int64_t sample = 0;
  bool nan = false;
  const double dt = args->timeStep;
  for (int64_t i = args->start_iteration; i < plan.total_iterations; i++) {
    const double t = i * dt;
    for (int a = e_lo; a < e_hi && !nan; a++) {
      V3 H = v3(0, 0, 0);
      for (int b = 0; b < N; b++) {
        if (a == b) continue;
        const V3 ra = arr3(g->coordinates + 3 * a), rb = arr3(g->coordinates + 3 * b);
        const LayerSim &lb = S.j[b].layers[top];
        if (g->interaction == CMTJB_INT_DIPOLE) H = H + dipole_kanao(ra, rb, lb);
        else if (g->interaction == CMTJB_INT_DIPOLE_NOMURA) H = H + dipole_nomura(ra, rb, lb);
        else if (g->interaction == CMTJB_INT_TABLE) H = H + pair_table_field(g, a, b, lb.mag);
      }
      LayerSim &la = S.j[a].layers[top];
      const double h[3] = {H.x, H.y, H.z};
      for (int c = 0; c < 3; c++) {
        cmtjb_driver_t &d = la.drv[CMTJB_SLOT_HRES_X + c];
        std::memset(&d, 0, sizeof(d));
        d.update = CMTJB_UPD_CONSTANT;
        d.constantValue = h[c];
      }
    }
    for (int a = e_lo; a < e_hi && !nan; a++)
      if (!step_junction(S.j[a], S.solver, t, dt, i)) nan = true;
    if (!(i % plan.write_every)) {
      if (do_log && !nan)
        for (int a = 0; a < own; a++) write_sample(S.j[e_lo + a], jlog, NS, own, sample, a);
      sample++;
    }
    /* every rank keeps calling the collective for the full run, NaN or not (the others are waiting in it) */
    publish();
    if (exchange(ctx, xbuf.data(), (int64_t)3 * per, i, stream)) return fail(CMTJB_ERUNTIME, "exchange callback failed");
    absorb();
  }
  for (int l = 0; l < nl; l++)
    for (int a = 0; a < own; a++) {
      const V3 m = S.j[e_lo + a].layers[l].mag;
      m_state[(size_t)(l * 3 + 0) * own + a] = m.x;
      m_state[(size_t)(l * 3 + 1) * own + a] = m.y;
      m_state[(size_t)(l * 3 + 2) * own + a] = m.z;
    }
  for (int a = 0; a < own; a++)
    if (status) status[a] = nan ? 1 : 0;
  if (nan) return fail(CMTJB_ENAN, "NAN magnetisation");
  return CMTJB_OK;
}


} /* extern "C" */
