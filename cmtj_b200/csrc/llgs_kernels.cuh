// llgs_kernels.cuh -- fused time-loop kernels: one launch integrates every step of every
// instance of a batch (Junction::runSimulation, junction.hpp:1948-2056), one thread per
// instance, magnetisation in registers, decimated log written with coalesced streaming
// stores ([channel][sample][instance], instance fastest).
#pragma once
#include "llgs_device.cuh"

namespace cmtjb {

// 64-thread blocks: a 1e5-instance sweep is 1563 blocks = 10.6 per SM, so the busiest SM
// carries 11 (4 % over the mean) instead of 6 vs 5.3 (14 %) with 128-thread blocks.
constexpr int kBlock = 64;

template <int NL> struct State {
  d3 m[NL];
};

// ---- log writer: logLayerParams (junction.hpp:1580-1636), MR :1835-1867 ---------------------
template <int NL>
__device__ __forceinline__ void write_sample(const DevModel<NL> &M, const RunParams &R, const State<NL> &S, long long sample,
                                             long long i) {
  double *p = R.log + sample * R.B + i;
  const long long cs = R.n_samples * R.B;  // channel stride
#pragma unroll
  for (int l = 0; l < NL; l++) {
    __stcs(p + (3 * l + 0) * cs, S.m[l].x);
    __stcs(p + (3 * l + 1) * cs, S.m[l].y);
    __stcs(p + (3 * l + 2) * cs, S.m[l].z);
  }
  if (M.mr_mode == CMTJB_MR_CLASSIC) {
    double c;
    if constexpr (NL == 1) c = dot(S.m[0], ld3(M.L[0].ref));
    else c = dot(S.m[0], S.m[1]);
    __stcs(p + (3 * NL) * cs, fma(M.Rhalf, 1.0 - c, M.Rp));
  } else if (M.mr_mode == CMTJB_MR_STRIP) {
    double rx = 0.0, ry = 0.0;
#pragma unroll
    for (int l = 0; l < NL; l++) {
      const DevLayer &Lr = M.L[l];
      const d3 m = S.m[l];
      const double Rx = fma(Lr.SMR_X, m.y * m.y, fma(Lr.AMR_X, m.x * m.x, Lr.Rx0));
      const double Ry = fma(Lr.AMRSMR_Y * m.x, m.y, fma(Lr.AHE_half, m.z, Lr.Ry0));
      rx += 1.0 / Rx;
      ry += 1.0 / Ry;
    }
    __stcs(p + (3 * NL + 0) * cs, 1.0 / rx);
    __stcs(p + (3 * NL + 1) * cs, 1.0 / ry);
    __stcs(p + (3 * NL + 2) * cs, 0.0);
  }
}

// ---- fused post-processing of the sweep procedures (cmtjb_post_t, include/cmtj_b200.h) -----------
// Per-thread carry of the spin-diode reduction: the three delays of scipy's transposed direct
// form II order-3 IIR, the running sum of its output and the instance's angular frequency.
struct PostState {
  double z0, z1, z2, sum, omega;
};

// R[n] of compute_sd: the junction's own MR channel (junction.hpp:1865-1867) or Rx / Ry of
// calculate_resistance_series / _parallel (cmtj/utils/resistance.py:25-66, :96-149).
template <int NL>
__device__ __forceinline__ double sd_resistance(const DevModel<NL> &M, const PostParams &P, const State<NL> &S) {
  if (P.sd_mode == CMTJB_SD_RZ) {
    double c;
    if constexpr (NL == 1) c = dot(S.m[0], ld3(M.L[0].ref));
    else c = dot(S.m[0], S.m[1]);
    return fma(M.Rhalf, 1.0 - c, M.Rp);
  }
  const bool want_x = P.sd_mode == CMTJB_SD_RX_SERIES || P.sd_mode == CMTJB_SD_RX_PARALLEL;
  const bool parallel = P.sd_mode == CMTJB_SD_RX_PARALLEL || P.sd_mode == CMTJB_SD_RY_PARALLEL;
  double acc = 0.0;
#pragma unroll
  for (int l = 0; l < NL; l++) {
    const double *r = P.res[l < CMTJB_POST_MAX_LAYERS ? l : 0];
    const d3 m = S.m[l];
    const double s = want_x ? r[CMTJB_RES_RXX0] + (r[CMTJB_RES_AMR] * (m.x * m.x) + r[CMTJB_RES_SMR] * (m.y * m.y))
                            : r[CMTJB_RES_RXY0] + 0.5 * r[CMTJB_RES_AHE] * m.z +
                                  r[CMTJB_RES_W_OVER_L] * (r[CMTJB_RES_SMR] + r[CMTJB_RES_AMR]) * m.x * m.y;
    acc += parallel ? 1.0 / s : s;
  }
  return parallel ? 1.0 / acc : acc;
}

template <int NL, bool SD = true>
__device__ __forceinline__ void post_sample(const DevModel<NL> &M, const RunParams &R, const State<NL> &S, PostState &ps,
                                            double t, long long sample, long long i) {
  const PostParams &P = R.post;
  // projection channels (PIMM: sum_l w_l m_l,c ; procedures.py:171-181)
  for (int k = 0; k < P.n_proj; k++) {
    double v = 0.0;
#pragma unroll
    for (int l = 0; l < NL; l++) {
      const double *w = P.proj[k] + 3 * (l < CMTJB_POST_MAX_LAYERS ? l : 0);
      v = fma(w[2], S.m[l].z, fma(w[1], S.m[l].y, fma(w[0], S.m[l].x, v)));
    }
    __stcs(P.proj_log + ((long long)k * P.proj_n_samples + P.proj_sample0 + sample) * R.B + i, v);
  }
  // spin-diode mixing + low-pass (compute_sd, resistance.py:12-22): scipy's lfilter recurrence
  // in its own operation order, without FMA contraction (the filter is ill-conditioned: its poles
  // sit 6e-6 from the unit circle, so the operation order is visible in the fifth digit).
  if (SD && P.sd_mode != CMTJB_SD_NONE) {
    const double x = -sin(ps.omega * t) * sd_resistance<NL>(M, P, S);
    const double y = __dadd_rn(ps.z0, __dmul_rn(P.sd_b[0], x));
    ps.z0 = __dsub_rn(__dadd_rn(ps.z1, __dmul_rn(x, P.sd_b[1])), __dmul_rn(y, P.sd_a[1]));
    ps.z1 = __dsub_rn(__dadd_rn(ps.z2, __dmul_rn(x, P.sd_b[2])), __dmul_rn(y, P.sd_a[2]));
    ps.z2 = __dsub_rn(__dmul_rn(x, P.sd_b[3]), __dmul_rn(y, P.sd_a[3]));
    ps.sum += y;
  }
}

// ---- calculateEnergies = true (junction.hpp:1584-1613): the field terms of the LAST right-hand-side
// evaluation of the step, recomputed at the point (t_eval, Sm, Snb) where the reference cached them.
// General (F_ALL) variants only; off the hot path (__noinline__, runs on written samples).
template <int NL, uint32_t F>
__device__ __noinline__ void write_energy(const DevModel<NL> &M, const RunParams &R, const State<NL> &Sm, const State<NL> &Snb,
                                          double t_eval, const double *row_eval, long long sample, long long i) {
  const long long cs = R.n_samples * R.B;
  double *p = R.log + R.elog_ch0 * cs + sample * R.B + i;
  auto put = [&](double v) {
    __stcs(p, v);
    p += cs;
  };
  auto put3 = [&](d3 v) {
    put(v.x);
    put(v.y);
    put(v.z);
  };
  const StageTime t{t_eval, row_eval};  // row_eval: the custom-table row of that evaluation (inline variants)
  const d3 zero = mk(0, 0, 0);
  for (int l = 0; l < NL; l++) {
    const DevLayer &Lr = M.L[l];
    const d3 m = Sm.m[l];
    const d3 bot = l > 0 ? Snb.m[l - 1] : zero;
    const d3 top = l < NL - 1 ? Snb.m[l + 1] : zero;
    const double K = slot_value<F>(Lr, R, CMTJB_SLOT_K1, t, i);
    const double Jb = slot_value<F>(Lr, R, CMTJB_SLOT_IEC_BOTTOM, t, i), Jt = slot_value<F>(Lr, R, CMTJB_SLOT_IEC_TOP, t, i);
    put(K);
    put(slot_value<F>(Lr, R, CMTJB_SLOT_K2, t, i));
    put(Jb);
    put(Jt);
    put((Lr.stt || (Lr.sot && !Lr.dynamicSOT)) ? slot_value<F>(Lr, R, CMTJB_SLOT_CURRENT, t, i) : 0.0);
    put3(slot_vec<F>(Lr, R, CMTJB_SLOT_HEXT_X, t, i));
    const double J2b = slot_value<F>(Lr, R, CMTJB_SLOT_IECQ_BOTTOM, t, i), J2t = slot_value<F>(Lr, R, CMTJB_SLOT_IECQ_TOP, t, i);
    put3(bot * ((Jb + 2 * J2b * dot(bot, m)) * Lr.inv_Mst) + top * ((Jt + 2 * J2t * dot(top, m)) * Lr.inv_Mst));
    const d3 a = ld3(Lr.anis);
    put3(a * ((Lr.cK1 * K) * dot(a, m)));
    put3(mk(0, 0, 0) - sub_matvec(Lr.demag, m, zero));
    put3(zero);  // Hfluctuation: never passed non-default in the reference
    if (Lr.includeSOT) {
      d3 fl = zero, dl = zero;
      if (Lr.sot) {
        d3 ref = zero;
        if (Lr.refType == CMTJB_REF_FIXED) ref = ld3(Lr.ref);
        else if (Lr.refType == CMTJB_REF_TOP) ref = top;
        else if (Lr.refType == CMTJB_REF_BOTTOM) ref = bot;
        double Hdl, Hfl;
        if (Lr.dynamicSOT) {
          Hdl = slot_value<F>(Lr, R, CMTJB_SLOT_DL, t, i);
          Hfl = slot_value<F>(Lr, R, CMTJB_SLOT_FL, t, i);
        } else {
          const double I = slot_value<F>(Lr, R, CMTJB_SLOT_CURRENT, t, i);
          Hdl = Lr.dlT * I;
          Hfl = Lr.flT * I;
        }
        fl = ref * (Hfl - Lr.alpha * Hdl);
        dl = ref * (Hdl + Lr.alpha * Hfl);
      }
      put3(fl);
      put3(dl);
    }
  }
}

// ---- steppers -------------------------------------------------------------------------------------
// Each stepper is written per layer with compile-time layer index so that every DevLayer
// access is a constant-bank operand.

template <int NL, uint32_t F, int l, class Ext = NoExt>
__device__ __forceinline__ d3 layer_u(const DevModel<NL> &M, const RunParams &R, long long i, StageTime t, d3 m,
                                      const State<NL> &nb, const d3 *Hc, const Ext &ext = Ext()) {
  const d3 zero = mk(0, 0, 0);
  const d3 bot = (l > 0) ? nb.m[l > 0 ? l - 1 : 0] : zero;
  const d3 top = (l < NL - 1) ? nb.m[l < NL - 1 ? l + 1 : 0] : zero;
  return llg_u<layer_mask(F, l), (l > 0), (l < NL - 1), Ext>(M.L[l], R, l, i, t, m, bot, top, Hc[l], ext);
}
// the same with the complete plain-driver field of this stage time handed in (F_DYNI fast path)
template <int NL, uint32_t F, int l, class Ext = NoExt>
__device__ __forceinline__ d3 layer_u_pre(const DevModel<NL> &M, const RunParams &R, long long i, StageTime t, d3 m,
                                          const State<NL> &nb, d3 Hstage, const Ext &ext = Ext()) {
  const d3 zero = mk(0, 0, 0);
  const d3 bot = (l > 0) ? nb.m[l > 0 ? l - 1 : 0] : zero;
  const d3 top = (l < NL - 1) ? nb.m[l < NL - 1 ? l + 1 : 0] : zero;
  return llg_u<layer_mask(F, l), (l > 0), (l < NL - 1), Ext, true>(M.L[l], R, l, i, t, m, bot, top, Hstage, ext);
}

// RK4: Layer::rk4_step (junction.hpp:1017-1041) under runMultiLayerSolver (:1682-1696):
// neighbours frozen at their pre-step values for all four stages (Jacobi).
template <int NL, uint32_t F, int l, class Ext = NoExt>
__device__ __forceinline__ d3 rk4_layer(const DevModel<NL> &M, const RunParams &R, long long i, const double *row, double t,
                                        const State<NL> &S0, const d3 *Hc, d3 &m4, const Ext &ext = Ext(),
                                        const FastDrv *fd = nullptr) {
  const DevLayer &Lr = M.L[l];
  const double h = Lr.h;  // llg * dt: dm = h * u
  const d3 m0 = S0.m[l];
  const StageTime t0{t, row}, t1{t + 0.5 * R.dt, row + R.dyn_ncols}, t2{t + R.dt, row + 2 * R.dyn_ncols};
  d3 u1, u2, u3, u4;
  if (((layer_mask(F, l) & F_DYNI) != 0 || (F & F_FAST) != 0) && fd != nullptr) {
    // the plain-driver field at the step's three distinct stage times, from the resolved driver
    constexpr uint32_t FL = layer_mask(F, l);
    const d3 H0 = stage_field<FL>(Lr, R, l, i, t0, Hc[l], fd[l]), H1 = stage_field<FL>(Lr, R, l, i, t1, Hc[l], fd[l]),
             H2 = stage_field<FL>(Lr, R, l, i, t2, Hc[l], fd[l]);
    u1 = layer_u_pre<NL, F, l, Ext>(M, R, i, t0, m0, S0, H0, ext);
    u2 = layer_u_pre<NL, F, l, Ext>(M, R, i, t1, axpy(u1, Lr.h_half, m0), S0, H1, ext);
    u3 = layer_u_pre<NL, F, l, Ext>(M, R, i, t1, axpy(u2, Lr.h_half, m0), S0, H1, ext);
    m4 = axpy(u3, h, m0);
    u4 = layer_u_pre<NL, F, l, Ext>(M, R, i, t2, m4, S0, H2, ext);
  } else {
    u1 = layer_u<NL, F, l, Ext>(M, R, i, t0, m0, S0, Hc, ext);
    u2 = layer_u<NL, F, l, Ext>(M, R, i, t1, axpy(u1, Lr.h_half, m0), S0, Hc, ext);
    u3 = layer_u<NL, F, l, Ext>(M, R, i, t1, axpy(u2, Lr.h_half, m0), S0, Hc, ext);
    m4 = axpy(u3, h, m0);  // where the reference's cached log values come from (energy log)
    u4 = layer_u<NL, F, l, Ext>(M, R, i, t2, m4, S0, Hc, ext);
  }
  const d3 acc = mk(fma(2.0, u2.x + u3.x, u1.x + u4.x), fma(2.0, u2.y + u3.y, u1.y + u4.y), fma(2.0, u2.z + u3.z, u1.z + u4.z));
  return normalize3(axpy(acc, Lr.h_sixth, m0));
}

template <int NL, uint32_t F, int l = 0, class Ext = NoExt>
__device__ __forceinline__ void rk4_all(const DevModel<NL> &M, const RunParams &R, long long i, const double *row, double t,
                                        const State<NL> &S0, State<NL> &S1, const d3 *Hc, State<NL> &M4, const Ext &ext = Ext(),
                                        const FastDrv *fd = nullptr) {
  S1.m[l] = rk4_layer<NL, F, l, Ext>(M, R, i, row, t, S0, Hc, M4.m[l], ext, fd);
  if constexpr (l + 1 < NL) rk4_all<NL, F, l + 1, Ext>(M, R, i, row, t, S0, S1, Hc, M4, ext, fd);
}

// Euler-Heun: Junction::eulerHeunSolverStep (junction.hpp:1709-1754)
template <int NL, uint32_t F, int l = 0, class Ext = NoExt>
__device__ __forceinline__ void euler_heun_all(const DevModel<NL> &M, const RunParams &R, long long i, long long it,
                                               const double *row, double t, const State<NL> &S0, State<NL> &S1, const d3 *Hc,
                                               const Ext &ext = Ext(), NoiseCarry *nc = nullptr) {
  const DevLayer &Lr = M.L[l];
  const d3 m0 = S0.m[l];
  const StageTime t0{t, row};
  const d3 f = layer_u<NL, F, l, Ext>(M, R, i, t0, m0, S0, Hc, ext);
  d3 mn = axpy(f, Lr.llg * R.dt, m0);
  if constexpr ((layer_mask(F, l) & (F_THERMAL | F_ONEF)) != 0) {
    const d3 dW = noise_dW<layer_mask(F, l), NL>(Lr, R, l, i, it, t0, nc ? nc + l : nullptr);
    const double gs = Lr.llg * R.sqrt_dt;
    const d3 g = stoch_u(Lr, m0, dW);
    const d3 gp = stoch_u(Lr, axpy(g, gs, m0), dW);
    mn = axpy(g + gp, 0.5 * gs, mn);
  }
  S1.m[l] = normalize3(mn);
  if constexpr (l + 1 < NL) euler_heun_all<NL, F, l + 1, Ext>(M, R, i, it, row, t, S0, S1, Hc, ext, nc);
}

// Heun: Junction::heunSolverStep (junction.hpp:1756-1815); predictor for every layer first,
// corrector with the PREDICTED neighbours.
template <int NL, uint32_t F, int l = 0>
__device__ __forceinline__ void heun_predict(const DevModel<NL> &M, const RunParams &R, long long i, long long it,
                                             const double *row, double t, const State<NL> &S0, State<NL> &P, State<NL> &Fn,
                                             State<NL> &Gn, State<NL> &dWs, const d3 *Hc, NoiseCarry *nc = nullptr) {
  const DevLayer &Lr = M.L[l];
  const d3 m0 = S0.m[l];
  const StageTime t0{t, row};
  const d3 f = layer_u<NL, F, l>(M, R, i, t0, m0, S0, Hc);
  Fn.m[l] = f;
  d3 mp = axpy(f, Lr.llg * R.dt, m0);
  if constexpr ((layer_mask(F, l) & (F_THERMAL | F_ONEF)) != 0) {
    const d3 dW = noise_dW<layer_mask(F, l), NL>(Lr, R, l, i, it, t0, nc ? nc + l : nullptr);
    const d3 g = stoch_u(Lr, m0, dW);
    dWs.m[l] = dW;
    Gn.m[l] = g;
    mp = axpy(g, Lr.llg * R.sqrt_dt, mp);
  }
  P.m[l] = mp;
  if constexpr (l + 1 < NL) heun_predict<NL, F, l + 1>(M, R, i, it, row, t, S0, P, Fn, Gn, dWs, Hc, nc);
}
template <int NL, uint32_t F, int l = 0>
__device__ __forceinline__ void heun_correct(const DevModel<NL> &M, const RunParams &R, long long i, const double *row, double t,
                                             const State<NL> &S0, const State<NL> &P, const State<NL> &Fn, const State<NL> &Gn,
                                             const State<NL> &dWs, State<NL> &S1, const d3 *Hc) {
  const DevLayer &Lr = M.L[l];
  const StageTime t2{t + R.dt, row + 2 * R.dyn_ncols};
  const d3 f2 = layer_u<NL, F, l>(M, R, i, t2, P.m[l], P, Hc);
  d3 mn = axpy(Fn.m[l] + f2, 0.5 * (Lr.llg * R.dt), S0.m[l]);
  if constexpr ((layer_mask(F, l) & (F_THERMAL | F_ONEF)) != 0) {
    const d3 g2 = stoch_u(Lr, P.m[l], dWs.m[l]);
    mn = axpy(Gn.m[l] + g2, 0.5 * (Lr.llg * R.sqrt_dt), mn);
  }
  S1.m[l] = normalize3(mn);
  if constexpr (l + 1 < NL) heun_correct<NL, F, l + 1>(M, R, i, row, t, S0, P, Fn, Gn, dWs, S1, Hc);
}

// One step.  Em / Enb receive the point of the step's LAST right-hand-side evaluation (stage
// magnetisations and the neighbour set it used); only the energy log of the general variants reads
// them, everywhere else they are dead code.
template <int NL, int SOLVER, uint32_t F>
__device__ __forceinline__ void do_step(const DevModel<NL> &M, const RunParams &R, long long i, long long it,
                                        const double *row, double t, State<NL> &S, const d3 *Hc, State<NL> &Em, State<NL> &Enb,
                                        const FastDrv *fd, NoiseCarry *nc) {
  State<NL> S1;
  if constexpr (SOLVER == CMTJB_RK4) {
    rk4_all<NL, F>(M, R, i, row, t, S, S1, Hc, Em, NoExt(), fd);
    Enb = S;
  } else if constexpr (SOLVER == CMTJB_EULER_HEUN) {
    euler_heun_all<NL, F>(M, R, i, it, row, t, S, S1, Hc, NoExt(), nc);
    Em = S;
    Enb = S;
  } else {
    State<NL> P, Fn, Gn, dWs;
    heun_predict<NL, F>(M, R, i, it, row, t, S, P, Fn, Gn, dWs, Hc, nc);
    heun_correct<NL, F>(M, R, i, row, t, S, P, Fn, Gn, dWs, S1, Hc);
    Em = P;
    Enb = P;
  }
  S = S1;
}

// ---- fixed-step time loop (RK4 / Heun / Euler-Heun) ----------------------------------------------
// Steps [it0, it1) of instance i, a sub-range of the launch's [R.it_begin, R.it_end).  State is
// read from and written back to m_state (and sd_state), so consecutive sub-ranges of one instance
// may be integrated by different warps (llgs_balanced_kernel); loads go through L2 (__ldcg) because
// an earlier sub-range may have been written from another SM.
template <int NL, int SOLVER, uint32_t F>
__device__ __forceinline__ void run_span(const DevModel<NL> &M, const RunParams &R, long long i, long long it0, long long it1) {
  State<NL> S;
  d3 Hc[NL];
#pragma unroll
  for (int l = 0; l < NL; l++) {
    S.m[l] = mk(__ldcg(R.m_state + (3 * l + 0) * R.B + i), __ldcg(R.m_state + (3 * l + 1) * R.B + i),
                __ldcg(R.m_state + (3 * l + 2) * R.B + i));
    Hc[l] = constant_field(M.L[l], R, i);
  }
  // F_DYNI RK4 variants (VSD maps): the single time-dependent field driver of each layer, resolved once
  FastDrv fdrv[NL];
  const FastDrv *fd = nullptr;
  if constexpr ((F & (F_DYNI | F_FAST)) != 0 && SOLVER == CMTJB_RK4 && F != F_ALL) {
#pragma unroll
    for (int l = 0; l < NL; l++) fdrv[l] = resolve_fast<F>(M.L[l], R, i);
    fd = fdrv;
  }
  // sample k is written after step k*write_every (junction.hpp:1973-1980): step `it` writes when
  // it % write_every == 0, so ceil(x / write_every) samples precede step x
  const long long we = R.write_every;
  long long until_write = (we - (it0 % we)) % we;
  long long sample = R.first_sample + ((it0 + we - 1) / we - (R.it_begin + we - 1) / we);
  const bool logging = R.log != nullptr;
  // advances by one step (3 stage rows) per iteration
  const double *row = R.dyn_tab ? R.dyn_tab + (it0 - R.it_begin) * 3 * R.dyn_ncols : nullptr;
  PostState ps{0.0, 0.0, 0.0, 0.0, 0.0};
  if constexpr ((F & F_POST) != 0) {
    if (R.post.sd_mode != CMTJB_SD_NONE) {
      ps.z0 = __ldcg(R.post.sd_state + i);
      ps.z1 = __ldcg(R.post.sd_state + R.B + i);
      ps.z2 = __ldcg(R.post.sd_state + 2 * R.B + i);
      ps.sum = __ldcg(R.post.sd_state + 3 * R.B + i);
      ps.omega = R.post.sd_omega[i];
    }
  }
  // thermal noise: the spare normals of each layer's current Philox block (normal3_carried)
  NoiseCarry carry[NL];
  NoiseCarry *nc = nullptr;
  if constexpr ((F & F_THERMAL) != 0 && SOLVER != CMTJB_RK4) {
#pragma unroll
    for (int l = 0; l < NL; l++)
      carry[l] = (M.L[l].thermal && (layer_mask(F, l) & F_THERMAL)) ? carry_before(R.pk, (unsigned long long)(R.inst_offset + i), (unsigned)l, (unsigned long long)it0)
                                                                     : NoiseCarry{0.f, 0.f, 0.f};
    nc = (R.tune_flags & 1) ? nullptr : carry;
  }
  for (long long it = it0; it < it1; ++it) {
    const double t = (double)it * R.dt;
    State<NL> Em, Enb;
    do_step<NL, SOLVER, F>(M, R, i, it, row, t, S, Hc, Em, Enb, fd, nc);
    if constexpr ((F & (F_DYN | F_DYNS | F_DYNI)) != 0) {
      if (row) row += 3 * R.dyn_ncols;  // table variants: the stage table; inline variants: the custom table, if any
    }
    if (until_write == 0) {
      if (logging) write_sample<NL>(M, R, S, sample, i);
      if constexpr ((F & F_POST) != 0) post_sample<NL, true>(M, R, S, ps, t, sample, i);
      else if constexpr ((F & F_PROJ) != 0) post_sample<NL, false>(M, R, S, ps, t, sample, i);
      if constexpr (F == F_ALL) {
        if (logging && R.elog_ch0 >= 0)
          write_energy<NL, F>(M, R, Em, Enb, SOLVER == CMTJB_EULER_HEUN ? t : t + R.dt,
                              row ? row - (SOLVER == CMTJB_EULER_HEUN ? 3 : 1) * R.dyn_ncols : nullptr, sample, i);
      }
      sample++;
      until_write = we;
    }
    until_write--;
  }
#pragma unroll
  for (int l = 0; l < NL; l++) {
    R.m_state[(3 * l + 0) * R.B + i] = S.m[l].x;
    R.m_state[(3 * l + 1) * R.B + i] = S.m[l].y;
    R.m_state[(3 * l + 2) * R.B + i] = S.m[l].z;
  }
  if constexpr ((F & F_POST) != 0) {
    if (R.post.sd_mode != CMTJB_SD_NONE) {
      R.post.sd_state[i] = ps.z0;
      R.post.sd_state[R.B + i] = ps.z1;
      R.post.sd_state[2 * R.B + i] = ps.z2;
      R.post.sd_state[3 * R.B + i] = ps.sum;
    }
  }
  // RK4 throws "NAN magnetisation" on isnan(mag.x) (junction.hpp:1037-1039); NaN is absorbing,
  // so testing the state at the end of a sub-range is equivalent.
  if (R.status) {
    int st = 0;
    if constexpr (SOLVER == CMTJB_RK4) {
#pragma unroll
      for (int l = 0; l < NL; l++) st |= isnan(S.m[l].x) ? 1 : 0;
    }
    if (it0 == 0 || st) R.status[i] = st;
  }
}

// Register budget: left to the compiler for every variant.  Tried and dropped (round 2): capping the stochastic single-
// layer kernels at 64 registers for 8 warps per sub-partition instead of 5 (ncu: `wait` is their top stall) made
// them slower, 7.3 -> 6.5e10 (Heun) and 9.0 -> 8.2e10 (Euler-Heun) layer-steps/s: the spilled state costs more than
// the extra warps hide; at 80 registers (6 warps, 20-160 bytes of spills) still 6 % slower, 7.29 -> 6.87e10 and 8.55 -> 8.10e10.  NB a second __launch_bounds__ argument of 1 is NOT neutral: it lifts the S1 kernel from 87 to
// 116 registers and the 3-layer Heun kernel from 120 to 192, 9 % slower each.
constexpr int min_blocks(int, int, int, uint32_t) { return 0; }  // 0 = no constraint

// One thread per instance, whole step range of the launch.
template <int NL, int SOLVER, uint32_t F>
__global__ void __launch_bounds__(kBlock, min_blocks(kBlock, NL, SOLVER, F)) llgs_fixed_kernel(const __grid_constant__ DevModel<NL> M,
                                                            const __grid_constant__ RunParams R) {
  const long long i = R.i_lo + (long long)blockIdx.x * kBlock + threadIdx.x;
  if (i >= R.i_hi) return;
  run_span<NL, SOLVER, F>(M, R, i, R.it_begin, R.it_end);
}

// ---- SMSP-balanced persistent form ---------------------------------------------------------------------
// The FP64 pipe is per SM sub-partition (4 x 148 = 592 on B200) and a warp is bound to one sub-
// partition for its lifetime.  A sweep of W = ceil(B/32) warps that is not a multiple of 592 leaves
// the run time set by the sub-partitions carrying ceil(W/592) warps while the others idle at the
// end: W = 3125 (the 1e5-point PIMM sweep) is 5.28 per sub-partition, i.e. 12 % lost to the ones
// that carry 6.  Here the run is cut into (round, instance-warp) tasks of `chunk` steps; resident
// warps draw tasks from a ticket counter in (round-major) order, so every sub-partition stays busy
// until the queue drains.  Task (r, w) may start once (r-1, w) has published the state of its 32
// instances: a per-instance-warp round counter, release/acquire at gpu scope.  Tickets are handed
// out in dependency order, so a waiting warp only ever waits on a warp that is already running.
//
// Tried and dropped (round 2, S1 sweep, same box): claiming any unheld instance-warp instead of
// waiting for the ticket's predecessor removes the `sleeping` stall (1.07 -> 0.02 per issue) but the
// lane-0 search executes 19 % more warp instructions and the kernel gets 8 % slower (376 vs 348 ms).
struct Balance {
  unsigned long long *ticket;  // zeroed before the launch
  int *done;                   // [n_wtasks] rounds completed per instance-warp, zeroed before the launch
  long long chunk;             // steps per task
  long long n_rounds, n_wtasks;
};
constexpr int kBalBlock = 128;  // one warp per sub-partition per block

__device__ __forceinline__ int ld_acquire_gpu(const int *p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu(int *p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

template <int NL, int SOLVER, uint32_t F>
__global__ void __launch_bounds__(kBalBlock, min_blocks(kBalBlock, NL, SOLVER, F)) llgs_balanced_kernel(const __grid_constant__ DevModel<NL> M,
                                                                                                        const __grid_constant__ RunParams R,
                                                                                                        const __grid_constant__ Balance Q) {
  const unsigned lane = threadIdx.x & 31u;
  const unsigned long long total = (unsigned long long)Q.n_wtasks * (unsigned long long)Q.n_rounds;
  for (;;) {
    unsigned long long id = 0;
    if (lane == 0) id = atomicAdd(Q.ticket, 1ULL);
    id = __shfl_sync(0xffffffffu, id, 0);
    if (id >= total) break;
    const long long round = (long long)(id / (unsigned long long)Q.n_wtasks);
    const long long w = (long long)(id - (unsigned long long)round * (unsigned long long)Q.n_wtasks);
    if (round > 0) {
      if (lane == 0)
        while (ld_acquire_gpu(Q.done + w) < (int)round) __nanosleep(256);
      __syncwarp();
    }
    const long long i = R.i_lo + w * 32 + lane;
    const long long it0 = R.it_begin + round * Q.chunk;
    const long long it1 = (it0 + Q.chunk < R.it_end) ? it0 + Q.chunk : R.it_end;
    if (i < R.i_hi) run_span<NL, SOLVER, F>(M, R, i, it0, it1);
    __threadfence();
    __syncwarp();
    if (lane == 0) st_release_gpu(Q.done + w, (int)round + 1);
  }
}

// ---- junctions of 5..8 layers: one LANE per layer -------------------------------------------------------
// The per-thread kernels above keep every layer of an instance in one thread's registers, which stops
// paying beyond four layers (the general variant of a 4-layer junction is already 14k instructions).
// Here an instance is a group of kWideGroup = 8 consecutive lanes, lane l integrating layer l; the
// neighbours' magnetisations (runMultiLayerSolver's snapshot, junction.hpp:1682-1696) arrive by warp
// shuffle, the layer models are read from a device array indexed by the lane.  General term set with
// inline drivers (F_ALL); same arithmetic per layer as the per-thread kernels, same Philox keys.
constexpr int kWideGroup = 8;
struct WideModel {
  const DevLayer *L;  // device [n_layers]
  int32_t n_layers, mr_mode;
  double Rp, Rhalf;
};

__device__ __forceinline__ d3 shfl3(d3 v, int src_lane_in_group) {
  return mk(__shfl_sync(0xffffffffu, v.x, src_lane_in_group, kWideGroup), __shfl_sync(0xffffffffu, v.y, src_lane_in_group, kWideGroup),
            __shfl_sync(0xffffffffu, v.z, src_lane_in_group, kWideGroup));
}

template <int SOLVER>
__global__ void __launch_bounds__(128) llgs_wide_kernel(const __grid_constant__ WideModel W, const __grid_constant__ RunParams R) {
  constexpr uint32_t F = F_ALL;
  const long long gt = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long i_raw = R.i_lo + gt / kWideGroup;
  const int l_raw = (int)(gt % kWideGroup);
  const int NL = W.n_layers;
  const bool live = i_raw < R.i_hi && l_raw < NL;  // idle lanes still take part in the shuffles
  const long long i = i_raw < R.i_hi ? i_raw : R.i_hi - 1;
  const int l = l_raw < NL ? l_raw : NL - 1;
  const DevLayer &Lr = W.L[l];
  const d3 zero = mk(0, 0, 0);
  d3 m = mk(__ldcg(R.m_state + (3 * l + 0) * R.B + i), __ldcg(R.m_state + (3 * l + 1) * R.B + i), __ldcg(R.m_state + (3 * l + 2) * R.B + i));
  const d3 Hc = constant_field(Lr, R, i);
  // neighbours of the lane's layer; the reference passes the zero vector beyond the ends (junction.hpp:1686-1690)
  auto neighbours = [&](d3 v, d3 &bot, d3 &top) {
    const d3 b = shfl3(v, l_raw - 1 < 0 ? 0 : l_raw - 1), t = shfl3(v, l_raw + 1 >= kWideGroup ? kWideGroup - 1 : l_raw + 1);
    bot = l > 0 ? b : zero;
    top = l < NL - 1 ? t : zero;
  };
  // `stage` (0, 1, 2 = t, t + dt/2, t + dt) selects the row of the caller's custom-driver table for this step, if any
  const double *crow = nullptr;
  auto rhs = [&](double t, int stage, d3 mm, d3 bot, d3 top) {
    return llg_u<F, true, true>(Lr, R, l, i, StageTime{t, crow ? crow + stage * R.n_custom : nullptr}, mm, bot, top, Hc);
  };
  const long long we = R.write_every;
  long long until_write = (we - (R.it_begin % we)) % we;
  long long sample = R.first_sample;
  const long long cs = R.n_samples * R.B;
  for (long long it = R.it_begin; it < R.it_end; ++it) {
    const double t = (double)it * R.dt;
    crow = R.custom ? R.custom + 3 * it * R.n_custom : nullptr;
    d3 bot, top;
    neighbours(m, bot, top);
    if constexpr (SOLVER == CMTJB_RK4) {
      const d3 u1 = rhs(t, 0, m, bot, top);
      const d3 u2 = rhs(t + 0.5 * R.dt, 1, axpy(u1, Lr.h_half, m), bot, top);
      const d3 u3 = rhs(t + 0.5 * R.dt, 1, axpy(u2, Lr.h_half, m), bot, top);
      const d3 u4 = rhs(t + R.dt, 2, axpy(u3, Lr.h, m), bot, top);
      const d3 acc = mk(fma(2.0, u2.x + u3.x, u1.x + u4.x), fma(2.0, u2.y + u3.y, u1.y + u4.y), fma(2.0, u2.z + u3.z, u1.z + u4.z));
      m = normalize3(axpy(acc, Lr.h_sixth, m));
    } else if constexpr (SOLVER == CMTJB_EULER_HEUN) {
      const StageTime t0{t, crow};
      const d3 f = rhs(t, 0, m, bot, top);
      d3 mn = axpy(f, Lr.llg * R.dt, m);
      const d3 dW = noise_dW<F, 1>(Lr, R, l, i, it, t0);
      const double gs = Lr.llg * R.sqrt_dt;
      const d3 g = stoch_u(Lr, m, dW);
      const d3 gp = stoch_u(Lr, axpy(g, gs, m), dW);
      mn = axpy(g + gp, 0.5 * gs, mn);
      m = normalize3(mn);
    } else {  // Heun: predictor with the current neighbours, corrector with the PREDICTED ones (junction.hpp:1756-1815)
      const StageTime t0{t, crow};
      const d3 f = rhs(t, 0, m, bot, top);
      const d3 dW = noise_dW<F, 1>(Lr, R, l, i, it, t0);
      const d3 g = stoch_u(Lr, m, dW);
      const d3 mp = axpy(g, Lr.llg * R.sqrt_dt, axpy(f, Lr.llg * R.dt, m));
      d3 pb, pt;
      neighbours(mp, pb, pt);
      const d3 f2 = rhs(t + R.dt, 2, mp, pb, pt);
      const d3 g2 = stoch_u(Lr, mp, dW);
      d3 mn = axpy(f + f2, 0.5 * (Lr.llg * R.dt), m);
      mn = axpy(g + g2, 0.5 * (Lr.llg * R.sqrt_dt), mn);
      m = normalize3(mn);
    }
    if (until_write == 0) {
      if (R.log != nullptr) {  // logLayerParams (junction.hpp:1580-1636), MR :1835-1867
        double *p = R.log + sample * R.B + i;
        if (live) {
          __stcs(p + (3 * l + 0) * cs, m.x);
          __stcs(p + (3 * l + 1) * cs, m.y);
          __stcs(p + (3 * l + 2) * cs, m.z);
        }
        if (W.mr_mode == CMTJB_MR_CLASSIC) {
          const d3 m1 = shfl3(m, 1);  // layer 1 of this instance (NL >= 5 here)
          if (live && l == 0) __stcs(p + (3 * NL) * cs, fma(W.Rhalf, 1.0 - dot(m, m1), W.Rp));
        } else if (W.mr_mode == CMTJB_MR_STRIP) {
          double rx = 0.0, ry = 0.0;
          if (l_raw < NL) {
            rx = 1.0 / fma(Lr.SMR_X, m.y * m.y, fma(Lr.AMR_X, m.x * m.x, Lr.Rx0));
            ry = 1.0 / fma(Lr.AMRSMR_Y * m.x, m.y, fma(Lr.AHE_half, m.z, Lr.Ry0));
          }
          for (int off = kWideGroup >> 1; off > 0; off >>= 1) {
            rx += __shfl_xor_sync(0xffffffffu, rx, off, kWideGroup);
            ry += __shfl_xor_sync(0xffffffffu, ry, off, kWideGroup);
          }
          if (live && l == 0) {
            __stcs(p + (3 * NL + 0) * cs, 1.0 / rx);
            __stcs(p + (3 * NL + 1) * cs, 1.0 / ry);
            __stcs(p + (3 * NL + 2) * cs, 0.0);
          }
        }
      }
      sample++;
      until_write = we;
    }
    until_write--;
  }
  if (live) {
    R.m_state[(3 * l + 0) * R.B + i] = m.x;
    R.m_state[(3 * l + 1) * R.B + i] = m.y;
    R.m_state[(3 * l + 2) * R.B + i] = m.z;
  }
  if (R.status) {  // RK4 throws "NAN magnetisation" on isnan(mag.x) of any layer (junction.hpp:1037-1039)
    int st = (SOLVER == CMTJB_RK4 && live && isnan(m.x)) ? 1 : 0;
    for (int off = kWideGroup >> 1; off > 0; off >>= 1) st |= __shfl_xor_sync(0xffffffffu, st, off, kWideGroup);
    if (live && l == 0 && (R.it_begin == 0 || st)) R.status[i] = st;
  }
}

template <int SOLVER> cudaError_t launch_wide(const WideModel &W, const RunParams &R, cudaStream_t stream) {
  const long long threads = (R.i_hi - R.i_lo) * kWideGroup;
  llgs_wide_kernel<SOLVER><<<(unsigned)((threads + 127) / 128), 128, 0, stream>>>(W, R);
  return cudaGetLastError();
}

// ---- Dormand-Prince (single layer): Layer::dormand_prince_step (junction.hpp:866-1002) and the
// adaptive loop of runSimulation (:1983-2044), including its quirks: time advances by the NEW
// step size after an accepted step, and consecutive_rejections is reset every iteration.
template <uint32_t F>
__global__ void __launch_bounds__(kBlock) llgs_dopri_kernel(const __grid_constant__ DevModel<1> M,
                                                            const __grid_constant__ RunParams R) {
  const long long i = (long long)blockIdx.x * kBlock + threadIdx.x;
  if (i >= R.B) return;
  const DevLayer &Lr = M.L[0];
  State<1> S;
  S.m[0] = mk(R.m_state[i], R.m_state[R.B + i], R.m_state[2 * R.B + i]);
  d3 Hc[1];
  Hc[0] = constant_field(Lr, R, i);
  double t = 0.0, next_write = 0.0, h = R.dt;
  double integral_error = Lr.integral_error, prev_error_ratio = Lr.prev_error_ratio;
  long long sample = 0;
  int st = 0;
  const bool logging = R.log != nullptr;
  static_assert((F & (F_DYN | F_DYNS)) == 0 || (F & F_DYNI) != 0, "adaptive time needs inline driver evaluation");
  auto f = [&](double tt, d3 mm) { return layer_u<1, F, 0>(M, R, i, StageTime{tt, nullptr}, mm, S, Hc) * Lr.llg; };
  while (t < R.total_time) {
    if (t + h > R.total_time) h = R.total_time - t;
    if (isnan(h) || isinf(h) || h <= 0) h = R.dt;
    const d3 m = S.m[0];
    const d3 k1 = f(t, m) * h;
    const d3 k2 = f(t + (1.0 / 5.0) * h, axpy(k1, 1.0 / 5.0, m)) * h;
    const d3 k3 = f(t + (3.0 / 10.0) * h, axpy(k2, 9.0 / 40.0, axpy(k1, 3.0 / 40.0, m))) * h;
    const d3 k4 = f(t + (4.0 / 5.0) * h, axpy(k3, 32.0 / 9.0, axpy(k2, -56.0 / 15.0, axpy(k1, 44.0 / 45.0, m)))) * h;
    const d3 k5 = f(t + (8.0 / 9.0) * h, axpy(k4, -212.0 / 729.0,
                                              axpy(k3, 64448.0 / 6561.0, axpy(k2, -25360.0 / 2187.0, axpy(k1, 19372.0 / 6561.0, m))))) *
                  h;
    const d3 k6 = f(t + h, axpy(k5, -5103.0 / 18656.0,
                                axpy(k4, 49.0 / 176.0,
                                     axpy(k3, 46732.0 / 5247.0, axpy(k2, -355.0 / 33.0, axpy(k1, 9017.0 / 3168.0, m)))))) *
                  h;
    const d3 m5 = axpy(k6, 11.0 / 84.0,
                       axpy(k5, -2187.0 / 6784.0, axpy(k4, 125.0 / 192.0, axpy(k3, 500.0 / 1113.0, axpy(k1, 35.0 / 384.0, m)))));
    const d3 k7 = f(t + h, m5) * h;
    const double t_last = t + h;  // last right-hand-side evaluation of this attempt (energy log)
    const d3 m4 = axpy(k7, 1.0 / 40.0,
                       axpy(k6, 187.0 / 2100.0,
                            axpy(k5, -92097.0 / 339200.0,
                                 axpy(k4, 393.0 / 640.0, axpy(k3, 7571.0 / 16695.0, axpy(k1, 5179.0 / 57600.0, m))))));
    const d3 e = m5 - m4;
    const double err = sqrt(dot(e, e));
    const double tol = fma(Lr.rel_tol, sqrt(dot(m, m)), Lr.abs_tol);
    double ratio = err / tol;
    if (tol < 1e-15 || isnan(ratio) || isinf(ratio)) ratio = 10.0;
    double factor;
    if (Lr.pid) {
      const double e_n = log(ratio);
      integral_error += e_n;
      factor = exp(Lr.kp * e_n + Lr.ki * integral_error + Lr.kd * (e_n - prev_error_ratio));
      prev_error_ratio = e_n;
    } else {
      factor = Lr.safety_factor * pow(1.0 / ratio, 0.2);
    }
    if (isnan(factor) || isinf(factor)) factor = Lr.min_factor;
    factor = fmin(factor, Lr.max_factor);
    factor = fmax(factor, Lr.min_factor);
    double nh = h * factor;
    if (isnan(nh) || isinf(nh) || nh <= 0) nh = h * Lr.min_factor;
    bool accepted = ratio <= 1.0;
    if (accepted) {
      const double l2 = dot(m5, m5);
      S.m[0] = (l2 != 0.0) ? m5 * (1.0 / sqrt(l2)) : m5;
      if (isnan(S.m[0].x)) {
        st = 1;
        break;
      }
    }
    h = nh;
    if (!accepted && h < 1e-15) accepted = true;  // forced acceptance (:2020-2027)
    if (accepted) {
      t += h;
      if (t >= next_write) {
        if (logging && sample < R.n_samples) {
          write_sample<1>(M, R, S, sample, i);
          if constexpr (F == F_ALL) {
            if (R.elog_ch0 >= 0) {
              State<1> Em, Enb;
              Em.m[0] = m5;
              Enb.m[0] = m5;
              write_energy<1, F>(M, R, Em, Enb, t_last, nullptr, sample, i);
            }
          }
          if (R.time_log) R.time_log[sample * R.B + i] = t;
        }
        sample++;
        next_write = ((double)((unsigned int)(t / R.write_frequency) + 1u)) * R.write_frequency;
      }
    }
  }
  R.m_state[i] = S.m[0].x;
  R.m_state[R.B + i] = S.m[0].y;
  R.m_state[2 * R.B + i] = S.m[0].z;
  if (R.n_samples_out) R.n_samples_out[i] = (int)(sample < R.n_samples ? sample : R.n_samples);
  if (R.status) R.status[i] = st;
}

// ---- per-launch driver table (F_DYN / F_DYNS variants) -----------------------------------------------
// tab[((k*3 + stage) * ncols) + col] at the stage times of step it_begin + k, using exactly the
// reference's time expressions i*dt, i*dt + 0.5*dt, i*dt + dt (junction.hpp:1974, :1020-1033).
// Columns 3l..3l+2: sum over layer l's time-dependent field drivers (amplitude and posine/halfsine
// rectification applied); then one column per time-dependent scalar slot holding its
// unit-amplitude waveform (amplitude may be per instance).
template <int NL>
__global__ void build_dyn_table_kernel(const __grid_constant__ DevModel<NL> M, const __grid_constant__ RunParams R, double *tab) {
  const long long n = (R.it_end - R.it_begin) * 3;
  const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n) return;
  const long long k = q / 3;
  const int stage = (int)(q - 3 * k);
  const double t0 = (double)(R.it_begin + k) * R.dt;
  const double t = stage == 0 ? t0 : (stage == 1 ? t0 + 0.5 * R.dt : t0 + R.dt);
  double *o = tab + q * R.dyn_ncols;
  // a custom driver's waveform is the caller's pre-evaluated value at this (absolute step, stage)
  const double *crow = R.custom ? R.custom + (3 * (R.it_begin + k) + stage) * R.n_custom : nullptr;
  auto shape = [&](const DevLayer &Lr, int s) {
    if (Lr.drv[s].type == CMTJB_UPD_CUSTOM) return crow ? crow[Lr.drv[s].pad] : 0.0;
    return driver_shape<0>(Lr, R, s, t, 0);
  };
  for (int l = 0; l < NL; l++) {
    const DevLayer &Lr = M.L[l];
    double h[3] = {0.0, 0.0, 0.0};
    for (int s = 0; s < CMTJB_SLOT_K1; s++)
      if ((Lr.dynmask >> s) & 1u) h[s % 3] += driver_post(Lr.drv[s].type, Lr.drv[s].amplitude * shape(Lr, s));
    o[3 * l + 0] = h[0];
    o[3 * l + 1] = h[1];
    o[3 * l + 2] = h[2];
    for (int s = CMTJB_SLOT_K1; s < CMTJB_NSLOTS; s++)
      if (Lr.dyncol[s] >= 0) o[Lr.dyncol[s]] = shape(Lr, s);
  }
}
template <int NL> cudaError_t launch_build_dyn_table(const DevModel<NL> &M, const RunParams &R, double *tab, cudaStream_t stream) {
  const long long n = (R.it_end - R.it_begin) * 3;
  if (n <= 0) return cudaSuccess;
  build_dyn_table_kernel<NL><<<(unsigned)((n + 127) / 128), 128, 0, stream>>>(M, R, tab);
  return cudaGetLastError();
}

// ---- host-side launch shims ---------------------------------------------------------------------
typedef cudaError_t (*LaunchFn)(const void *model, const RunParams *R, cudaStream_t stream);

// Decide between the plain and the SMSP-balanced form (see llgs_balanced_kernel).  CMTJB_BALANCE=0/1
// forces the choice (tests run both and compare bit-for-bit).
struct BalancePlan {
  bool use;
  long long chunk, n_rounds, n_wtasks;
  int blocks;
  int per_sm;  // resident blocks (= warps per sub-partition) the plan assumes on EVERY SM
};
BalancePlan plan_balance(const RunParams &R, const void *balanced_kernel, const void *plain_kernel);

template <int NL, int SOLVER, uint32_t F>
cudaError_t launch_fixed(const void *model, const RunParams *R, cudaStream_t stream) {
  const DevModel<NL> *M = static_cast<const DevModel<NL> *>(model);
  const BalancePlan bp = plan_balance(*R, (const void *)&llgs_balanced_kernel<NL, SOLVER, F>, (const void *)&llgs_fixed_kernel<NL, SOLVER, F>);
  if (bp.use) {
    const size_t bytes = 256 + (size_t)bp.n_wtasks * sizeof(int);
    char *scratch = nullptr;
    cudaError_t e = cudaMallocAsync((void **)&scratch, bytes, stream);
    if (e != cudaSuccess) return e;
    cudaMemsetAsync(scratch, 0, bytes, stream);
    Balance Q;
    Q.ticket = (unsigned long long *)scratch;
    Q.done = (int *)(scratch + 256);
    Q.chunk = bp.chunk;
    Q.n_rounds = bp.n_rounds;
    Q.n_wtasks = bp.n_wtasks;
    // The plan counts on exactly per_sm blocks on every SM (one warp per sub-partition each).  The block scheduler
    // only guarantees that when no SM can take more: each block asks for 1/per_sm of the SM's shared memory, which
    // it never touches, so that a (per_sm + 1)-th block does not fit.  Without this some SMs carry 5 warps per
    // sub-partition and others 3; the slow ones then publish late and every waiter of theirs sleeps (S1: 354-362 ms
    // unpinned, 348 ms pinned).
    size_t pin = 0;
    if (bp.per_sm > 0 && bp.blocks >= bp.per_sm * 8) {
      pin = ((size_t)227 * 1024 / (size_t)bp.per_sm - 1024) & ~(size_t)1023;
      if (cudaFuncSetAttribute(llgs_balanced_kernel<NL, SOLVER, F>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pin) != cudaSuccess) {
        cudaGetLastError();
        pin = 0;
      }
    }
    llgs_balanced_kernel<NL, SOLVER, F><<<(unsigned)bp.blocks, kBalBlock, pin, stream>>>(*M, *R, Q);
    e = cudaGetLastError();
    cudaFreeAsync(scratch, stream);
    return e;
  }
  const long long blocks = (R->i_hi - R->i_lo + kBlock - 1) / kBlock;
  llgs_fixed_kernel<NL, SOLVER, F><<<(unsigned)blocks, kBlock, 0, stream>>>(*M, *R);
  return cudaGetLastError();
}
template <uint32_t F> cudaError_t launch_dopri(const void *model, const RunParams *R, cudaStream_t stream) {
  const DevModel<1> *M = static_cast<const DevModel<1> *>(model);
  const long long blocks = (R->B + kBlock - 1) / kBlock;
  llgs_dopri_kernel<F><<<(unsigned)blocks, kBlock, 0, stream>>>(*M, *R);
  return cudaGetLastError();
}

struct Variant {
  int nl;
  int solver;
  uint32_t features;
  LaunchFn launch;
  const char *name;
};

}  // namespace cmtjb
