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

/* core/drivers.hpp:262-305 with helpers :73-108 */
double driver_value(const cmtjb_driver_t &d, double time) {
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
      Hdl = L.dampingLikeTorque * I;
      Hfl = L.fieldLikeTorque * I;
    }
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
      if (L.drivers[s].update == CMTJB_UPD_CUSTOM)
        return fail(CMTJB_ENOTSUP, "custom (Python callback) drivers are not supported on the device path");
  }
  p->effective_solver = mode;
  p->n_channels = 3 * jn->n_layers + (jn->mr_mode == CMTJB_MR_CLASSIC ? 1 : jn->mr_mode == CMTJB_MR_STRIP ? 3 : 0);
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
  const bool do_log = (args->log_mode == CMTJB_LOG_FULL) && log;
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
