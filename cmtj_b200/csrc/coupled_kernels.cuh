// coupled_kernels.cuh -- junctions that exchange information every time step:
//
//   stack_kernel   Stack<T>::runSimulation{KCL,NonKCL}   (core/stack.hpp:317-423)
//   group_kernel   GroupInteraction::runSimulation       (core/reservoir.hpp:129-226)
//
// One thread owns one junction (element) of one system instance for the whole time loop; its
// magnetisation stays in registers.  Junction models may differ inside a system, so they are read
// from a device array (L1/L2-cached) instead of the constant bank.  Two term sets are compiled: the
// lean one (F_COUPLED_LEAN: folded K1 + demag, constant drivers, IEC, STT / static SOT, thermal
// noise -- what stacks of STT junctions and plain macrospin arrays need) and the fully general one
// (F_ALL: every field/torque term, inline drivers, per-element overrides); coupled.cu takes the
// union of the junctions' needs and picks.  The
// per-step exchange is a block-level scan (stack) or an all-pairs field sum with m staged in
// shared memory (group); arrays larger than one block use a cooperative launch with one
// grid-wide barrier per step.
#pragma once
#include <cooperative_groups.h>

#include "llgs_kernels.cuh"

namespace cmtjb {

namespace cg = cooperative_groups;

constexpr uint32_t F_COUPLED_LEAN = F_IEC | F_STT | F_SOT | F_THERMAL;
constexpr int kCoupledBlockMax = 512;  // 128 registers per thread for the general term set

// waveform of an explicit driver record (same expressions as driver_shape, drivers.hpp:262-305)
__device__ __forceinline__ double shape_of(const DevDriver &d, double phase, double t) {
  switch (d.type) {
  case CMTJB_UPD_PULSE: {
    const int n = (int)(t / d.period);
    const double nT = n * d.period;
    return (nT <= t && t <= (nT + d.cycle * d.period)) ? 1.0 : 0.0;
  }
  case CMTJB_UPD_SINE:
  case CMTJB_UPD_POSINE:
  case CMTJB_UPD_HALFSINE: return sin(2 * M_PI * t * d.frequency + phase);
  case CMTJB_UPD_STEP: return (t >= d.timeStart && t <= d.timeStop) ? 1.0 : 0.0;
  case CMTJB_UPD_TRAPEZOID: {
    const double t0 = d.timeStart, e = d.edgeTime, st = d.steadyTime;
    if (t < t0) return 0.0;
    if (t <= t0 + e) return (t - t0) / e;
    if (t <= t0 + e + st) return 1.0;
    if (t <= t0 + 2 * e + st) return 1.0 - (t - (t0 + e + st)) / e;
    return 0.0;
  }
  case CMTJB_UPD_GAUSSIMPULSE: {
    const double x = t - d.timeStart;
    return exp(-(x * x) / (2 * (d.edgeTime * d.edgeTime)));
  }
  case CMTJB_UPD_GAUSSSTEP: return 0.5 * (1 + erf((t - d.timeStart) / (sqrt(2.0) * d.edgeTime)));
  default: return 0.0;
  }
}

// The current a Stack pushes into every layer of a junction for one step (stack.hpp:370-371,
// :413-415): constant + post(amplitude * waveform(t; phase)).  A step only ever asks for it at its
// three stage times t, t + dt/2, t + dt (Euler-Heun: t), so the values are computed once per step
// (one call site of the waveform code) and llg_u selects by comparing the stage time it was given.
struct StackCurrent {
  static constexpr bool kCurrent = true;
  double t1, t2;      // thresholds t + dt/4, t + 3dt/4 separating the stage times t, t + dt/2, t + dt
  double I0, I1, I2;  // current at the three stage times
  __device__ __forceinline__ double current(double t) const { return t >= t2 ? I2 : (t >= t1 ? I1 : I0); }
};
template <int SOLVER>
__device__ __forceinline__ StackCurrent stack_current(const DevDriver &d, double cst, double amp, double phase, double t, double dt) {
  StackCurrent c;
  c.t1 = t + 0.25 * dt;
  c.t2 = t + 0.75 * dt;
  if (amp == 0.0) {
    c.I0 = c.I1 = c.I2 = cst;
    return c;
  }
  double v[3];
  const double ts[3] = {t, t + 0.5 * dt, t + dt};  // the reference's stage-time expressions (junction.hpp:1020-1033)
#pragma unroll 1
  for (int k = 0; k < (SOLVER == CMTJB_RK4 ? 3 : 1); k++) v[k] = cst + driver_post(d.type, amp * shape_of(d, phase, ts[k]));
  c.I0 = v[0];
  c.I1 = SOLVER == CMTJB_RK4 ? v[1] : v[0];
  c.I2 = SOLVER == CMTJB_RK4 ? v[2] : v[0];
  return c;
}

template <int NL, int SOLVER, uint32_t F, class Ext>
__device__ __forceinline__ void coupled_step(const DevModel<NL> &M, const RunParams &R, long long i, long long it, double t,
                                             State<NL> &S, const d3 *Hc, const Ext &ext) {
  State<NL> S1, M4;
  if constexpr (SOLVER == CMTJB_RK4) rk4_all<NL, F, 0, Ext>(M, R, i, nullptr, t, S, S1, Hc, M4, ext);
  else euler_heun_all<NL, F, 0, Ext>(M, R, i, it, nullptr, t, S, S1, Hc, ext);
  S = S1;
}

template <int NL> __device__ __forceinline__ void load_state(const RunParams &R, long long i, State<NL> &S) {
#pragma unroll
  for (int l = 0; l < NL; l++)
    S.m[l] = mk(R.m_state[(3 * l + 0) * R.B + i], R.m_state[(3 * l + 1) * R.B + i], R.m_state[(3 * l + 2) * R.B + i]);
}
template <int NL> __device__ __forceinline__ void store_state(const RunParams &R, long long i, const State<NL> &S) {
#pragma unroll
  for (int l = 0; l < NL; l++) {
    R.m_state[(3 * l + 0) * R.B + i] = S.m[l].x;
    R.m_state[(3 * l + 1) * R.B + i] = S.m[l].y;
    R.m_state[(3 * l + 2) * R.B + i] = S.m[l].z;
  }
}
template <int NL, int SOLVER> __device__ __forceinline__ int nan_status(const State<NL> &S) {
  int st = 0;
  if constexpr (SOLVER == CMTJB_RK4) {
#pragma unroll
    for (int l = 0; l < NL; l++) st |= isnan(S.m[l].x) ? 1 : 0;
  }
  return st;
}

// ====================================================================================================
// Stack
// ====================================================================================================
struct StackParams {
  int N;              // junctions per stack
  int spb;            // stacks per block (blockDim.x = spb * N)
  int kind, use_kcl, delayed;
  int top, bottom;    // layer indices of topId / bottomId (bottom = -1: single-layer junctions)
  long long n_inst;
  double phaseOffset;
  double pol[3];
  const double *coupling;      // device [N-1]
  DevDriver cur;               // coupled current driver
  double cur_cst;
  const double *cur_ovr[10];   // per-instance override of a current-driver field, or null
  double *slog;                // [1 + N + (N-1)][n_samples][n_inst]
};

// inclusive segmented Hillis-Steele scan over the N threads of one stack (sum or product)
template <bool MUL> __device__ __forceinline__ double seg_scan(double v, double *a, int tid, int j, int N) {
  a[tid] = v;
  __syncthreads();
  for (int off = 1; off < N; off <<= 1) {
    const double o = (j >= off) ? a[tid - off] : (MUL ? 1.0 : 0.0);
    __syncthreads();
    v = MUL ? v * o : v + o;
    a[tid] = v;
    __syncthreads();
  }
  return v;
}

// Sum of v over the N threads of one stack, returned to all of them.  When stacks are warp-aligned
// (N a multiple of 32, or a power of two <= 32) it is a shuffle reduction plus, for N > 32, one
// shared-memory exchange of the per-warp partials: ONE block barrier instead of the scan's
// 1 + 2 log2 N.  `part` is double-buffered by the caller (step parity), so no trailing barrier is needed.
__device__ __forceinline__ double stack_sum(double v, double *scan, double *part, int tid, int j, int sl, int N, bool aligned) {
  if (!aligned) {
    seg_scan<false>(v, scan, tid, j, N);
    const double tot = scan[sl * N + N - 1];
    __syncthreads();  // scan[] is rewritten by the next reduction
    return tot;
  }
  const int width = N < 32 ? N : 32;
  for (int off = width >> 1; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off, width);
  if (N <= 32) return v;
  const int wps = N >> 5;  // warps per stack
  if ((tid & 31) == 0) part[tid >> 5] = v;
  __syncthreads();
  double tot = 0.0;
  const int w0 = sl * wps;
  for (int k = 0; k < wps; k++) tot += part[w0 + k];
  return tot;
}

template <int NL> __device__ __forceinline__ d3 layer_m(const State<NL> &S, int idx) {
  if constexpr (NL == 1) return S.m[0];
  else return idx == 0 ? S.m[0] : S.m[1];  // stacks hold junctions of at most two layers (stack.hpp:298-308)
}

// top-layer pick for the group kernels (junctions of 1 or 2 layers in this build), without dynamic
// indexing of the register-resident state
template <int NL> __device__ __forceinline__ d3 layer_m2(const State<NL> &S, int idx) {
  d3 v = S.m[0];
#pragma unroll
  for (int l = 1; l < NL; l++)
    if (idx == l) v = S.m[l];
  return v;
}

// MODE: which of the three coupling schemes of Stack::runSimulation this instantiation integrates
enum { STACK_KCL = 0, STACK_DELAYED = 1, STACK_SEQUENTIAL = 2 };

// UNIFORM: every junction of the system has the same model (the usual `[junction] * N` construction);
// it is then passed by value like in the sweep kernels and its parameters are constant-bank operands
// instead of registers loaded from models[j].
template <int NL, int SOLVER, uint32_t F, bool UNIFORM, int MODE>
__global__ void __maxnreg__((F == F_COUPLED_LEAN && UNIFORM) ? 88 : 128) stack_kernel(const DevModel<NL> *__restrict__ models,
                                                                 const __grid_constant__ DevModel<NL> Mu,
                                                                 const __grid_constant__ StackParams SP,
                                                                 const __grid_constant__ RunParams R) {
  extern __shared__ double sm[];
  const int N = SP.N;
  const int tid = threadIdx.x;
  const int sl = tid / N, j = tid - sl * N;
  const long long s = (long long)blockIdx.x * SP.spb + sl;
  const bool active = s < SP.n_inst;  // idle threads of the last block still take part in the barriers
  const long long i = active ? s * N + j : 0;
  const DevModel<NL> &M = UNIFORM ? Mu : models[j];
  double *dots = sm;                    // [2][blockDim]  m_top . p of every junction (double-buffered)
  double *scan = sm + 2 * blockDim.x;   // [blockDim]     scan workspace
  double *part = sm + 3 * blockDim.x;   // [2][blockDim / 32 + 1] per-warp partial sums (double-buffered)
  const int part_stride = (blockDim.x >> 5) + 1;
  const bool aligned = (N % 32 == 0) || (N <= 32 && (N & (N - 1)) == 0);

  State<NL> S;
  d3 Hc[NL];
  load_state<NL>(R, i, S);
#pragma unroll
  for (int l = 0; l < NL; l++) Hc[l] = constant_field(M.L[l], R, i);

  // this instance's coupled current driver (per-instance overrides of its fields)
  auto cur_field = [&](int f, double dflt) { return (SP.cur_ovr[f] && active) ? __ldg(SP.cur_ovr[f] + s) : dflt; };
  DevDriver cd = SP.cur;
  const double cur_cst = cur_field(CMTJB_FIELD_CONSTANT, SP.cur_cst);
  cd.amplitude = cur_field(CMTJB_FIELD_AMPLITUDE, cd.amplitude);
  cd.frequency = cur_field(CMTJB_FIELD_FREQUENCY, cd.frequency);
  cd.phase = cur_field(CMTJB_FIELD_PHASE, cd.phase);
  cd.period = cur_field(CMTJB_FIELD_PERIOD, cd.period);
  cd.cycle = cur_field(CMTJB_FIELD_CYCLE, cd.cycle);
  cd.timeStart = cur_field(CMTJB_FIELD_TIMESTART, cd.timeStart);
  cd.timeStop = cur_field(CMTJB_FIELD_TIMESTOP, cd.timeStop);
  cd.edgeTime = cur_field(CMTJB_FIELD_EDGETIME, cd.edgeTime);
  cd.steadyTime = cur_field(CMTJB_FIELD_STEADYTIME, cd.steadyTime);
  // getPhaseOffset: series phaseOffset * order, parallel phaseOffset (stack.hpp:441-443, :470-472)
  const double my_phase = cd.phase + (SP.kind == CMTJB_STACK_SERIES ? SP.phaseOffset * j : SP.phaseOffset);
  const double sign = (SP.kind == CMTJB_STACK_SERIES) ? 1.0 : -1.0;
  const double c_here = (j < N - 1) ? __ldg(SP.coupling + j) : 0.0;      // coupling(j)   : pair (j, j+1)
  const double c_prev = (j > 0) ? __ldg(SP.coupling + j - 1) : 0.0;      // coupling(j-1) : pair (j-1, j)
  const d3 pol = mk(SP.pol[0], SP.pol[1], SP.pol[2]);

  auto my_dot = [&]() {
    const d3 mt = layer_m<NL>(S, SP.top);
    if (NL >= 2 && SP.bottom >= 0) return dot(mt, layer_m<NL>(S, SP.bottom));
    return dot(mt, pol);
  };
  auto resistance = [&]() {  // getMagnetoresistance()[0], classic (junction.hpp:1865-1889)
    double c;
    if constexpr (NL == 1) c = dot(S.m[0], ld3(M.L[0].ref));
    else c = dot(S.m[0], S.m[1]);
    return fma(M.Rhalf, 1.0 - c, M.Rp);
  };

  long long until_write = (R.write_every - (R.it_begin % R.write_every)) % R.write_every;
  long long sample = R.first_sample;
  const long long scs = R.n_samples * SP.n_inst;  // stack-log channel stride
  int buf = 0;
  for (long long it = R.it_begin; it < R.it_end; ++it) {
    const double t = (double)it * R.dt;
    StackCurrent ext;
    double c_log = 0.0;  // what the reference logs as C_(j-1)_j (non-KCL) or C_j_(j+1) (KCL)
    if constexpr (MODE == STACK_KCL) {
      // stack.hpp:397-408: I_total = I0(t) (1 + sum_j C_j), C_j = c_j (m_j.p_j +- m_(j+1).p_(j+1))
      double *db = dots + buf * blockDim.x;
      db[tid] = my_dot();
      __syncthreads();
      const double C = (j < N - 1) ? c_here * (db[tid] + sign * db[tid + 1]) : 0.0;
      const double sumC = stack_sum(C, scan, part + buf * part_stride, tid, j, sl, N, aligned);
      const double I0 = cur_cst + (cd.amplitude == 0.0 ? 0.0 : driver_post(cd.type, cd.amplitude * shape_of(cd, cd.phase, t)));
      const double I = fma(sumC, I0, I0);
      ext.t1 = ext.t2 = t;
      ext.I0 = ext.I1 = ext.I2 = I;  // held over the step
      c_log = C;
      buf ^= 1;
      coupled_step<NL, SOLVER, F>(M, R, i, it, t, S, Hc, ext);
    } else if constexpr (MODE == STACK_DELAYED) {
      // stack.hpp:335-354: eff_j = prod_{k<=j, k>0} (1 + C_(k-1)) from the frozen magnetisations
      dots[tid] = my_dot();
      __syncthreads();
      const double f = (j > 0) ? 1.0 + c_prev * (dots[tid - 1] + sign * dots[tid]) : 1.0;
      const double eff = seg_scan<true>(f, scan, tid, j, N);
      ext = stack_current<SOLVER>(cd, cur_cst * eff, cd.amplitude * eff, my_phase, t, R.dt);
      c_log = eff;
      __syncthreads();  // dots/scan are rewritten next step
      coupled_step<NL, SOLVER, F>(M, R, i, it, t, S, Hc, ext);
    } else {
      // stack.hpp:356-365: Gauss-Seidel chain -- junction j sees the ALREADY STEPPED j-1.  One call site
      // of the step: every thread walks the chain, only thread jj commits.
      for (int jj = 0; jj < N; jj++) {
        if (j == jj) {
          double eff = 1.0;
          if (jj > 0) eff = scan[tid - 1] * (1.0 + c_prev * (dots[tid - 1] + sign * my_dot()));
          ext = stack_current<SOLVER>(cd, cur_cst * eff, cd.amplitude * eff, my_phase, t, R.dt);
          c_log = eff;
          coupled_step<NL, SOLVER, F>(M, R, i, it, t, S, Hc, ext);
          dots[tid] = my_dot();
          scan[tid] = eff;
        }
        __syncthreads();
      }
    }
    if (until_write == 0) {
      // logSimulationStep (stack.hpp:93-103): stack resistance, per-junction current and coupling
      const double Rj = resistance();
      __syncthreads();
      const double tot = seg_scan<false>(SP.kind == CMTJB_STACK_SERIES ? Rj : 1.0 / Rj, scan, tid, j, N);
      if (active && R.log) write_sample<NL>(M, R, S, sample, i);
      if (active && SP.slog) {
        double *p = SP.slog + sample * SP.n_inst + s;
        if (j == N - 1) __stcs(p, SP.kind == CMTJB_STACK_SERIES ? tot : 1.0 / tot);
        __stcs(p + (1 + j) * scs, ext.current(t));
        if (MODE == STACK_KCL) {
          if (j < N - 1) __stcs(p + (1 + N + j) * scs, c_log);
        } else if (j > 0) {
          __stcs(p + (1 + N + j - 1) * scs, c_log);
        }
      }
      __syncthreads();
      sample++;
      until_write = R.write_every;
    }
    until_write--;
  }
  if (active) {
    store_state<NL>(R, i, S);
    if (R.status) R.status[i] = nan_status<NL, SOLVER>(S);
  }
}

// ====================================================================================================
// GroupInteraction
// ====================================================================================================
struct GroupParams {
  int N;             // elements per group
  int P;             // threads cooperating on one element's pair sum (block kernel: a power of two <= 32)
  int A, tile;       // grid kernel: elements per block, partners per shared-memory tile
  int interaction;   // cmtjb_interaction
  int top;           // layer index of topId
  long long n_inst;
  const double *pos;   // device [3][N]
  const double *pref;  // device [N]: (Ms_b / mu0) * thickness_b * cellSurface_b  (reservoir.hpp:91-94)
  double *mbuf;        // cooperative kernel: [2][3][N] top-layer magnetisation exchange
  long long inst;      // cooperative kernel: which instance of the batch this launch integrates
};

// field at a from the dipole at b (reservoir.hpp:79-97); r = r_a - r_b:
//   H += pref (3 (m.r) r / r^5 - m / r^3)
// 1/sqrt(x) without the library routine's special-case branches (BSSY/BSYNC + slow path around every
// call): the SFU seed rsqrt.approx.ftz.f64 (MUFU.RSQ64H, relative error ~2^-22) refined by ONE third-order
// (Halley) step  y <- y (1 + e/2 + 3 e^2/8),  e = 1 - x y^2:  the residual is (5/16) e^3 ~ 2^-65, i.e. rounding
// level, in five FP64 instructions (two Newton steps take seven).  Inputs are squared pair distances, validated
// on the host to be >= 1e-28 (reservoir.hpp:84-87, :39-42 guards): no zero / inf / NaN.
__device__ __forceinline__ double rsqrt_nr(double x) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double e = fma(-(x * y), y, 1.0);
  return fma(y, fma(0.375, e, 0.5) * e, y);
}

// `self`: the pair (a, a) contributes nothing; handled without a branch (r2 -> 1, pref -> 0) so that the
// pair loop unrolls into independent chains.  25 FP64 instructions per pair:
//   H += p3 (g r - m),   p3 = pref / r^3,   g = 3 (m.r) / r^2
__device__ __forceinline__ void pair_accumulate(d3 &H, d3 r, d3 mb, double pref, bool self) {
  const double r2 = self ? 1.0 : dot(r, r);
  pref = self ? 0.0 : pref;
  const double inv = rsqrt_nr(r2);
  const double inv2 = inv * inv;
  const double p3 = (pref * inv) * inv2;
  const double g = (3.0 * inv2) * dot(mb, r);
  H = mk(fma(p3, fma(g, r.x, -mb.x), H.x), fma(p3, fma(g, r.y, -mb.y), H.y), fma(p3, fma(g, r.z, -mb.z), H.z));
}

// GroupInteraction, one group per block: blockDim.x = N * P rounded up to a warp multiple.  Two roles per
// thread: (a = tid / P, p = tid % P) sums the pair field of element a over b = p, p+P, ...; thread
// e = tid < N owns junction e (state in registers) and steps it.  The reduced fields pass through shared
// memory, so the steps run on N consecutive, fully populated lanes instead of one lane in P.
template <int NL, int SOLVER, uint32_t F, bool UNIFORM>
__global__ void __launch_bounds__(kCoupledBlockMax) group_block_kernel(const DevModel<NL> *__restrict__ models,
                                                                       const __grid_constant__ DevModel<NL> Mu,
                                                                       const __grid_constant__ GroupParams GP,
                                                                       const __grid_constant__ RunParams R) {
  extern __shared__ double sm[];
  const int N = GP.N, P = GP.P;
  const int tid = threadIdx.x;
  const int a_raw = tid / P, p = tid - a_raw * P;
  const int a = a_raw < N ? a_raw : N - 1;  // padding lanes of the last warp shadow element N-1
  const long long g = blockIdx.x;
  const bool stepper = tid < N;
  const int e = stepper ? tid : N - 1;
  const long long i = g * N + e;
  const DevModel<NL> &M = UNIFORM ? Mu : models[e];
  double *px = sm, *py = px + N, *pz = py + N, *pf = pz + N;  // static geometry + prefactors
  double *mb0 = pf + N;                                      // [2][3][N] double-buffered m_top
  double *Hs = mb0 + 6 * N;                                  // [3][N] reduced pair fields of this step
  for (int k = tid; k < N; k += blockDim.x) {
    px[k] = GP.pos[k];
    py[k] = GP.pos[N + k];
    pz[k] = GP.pos[2 * N + k];
    pf[k] = GP.pref[k];
  }
  State<NL> S;
  d3 Hc[NL];
  if (stepper) {
    load_state<NL>(R, i, S);
#pragma unroll
    for (int l = 0; l < NL; l++) Hc[l] = constant_field(M.L[l], R, i);
    const d3 mt = layer_m2<NL>(S, GP.top);
    mb0[e] = mt.x;
    mb0[N + e] = mt.y;
    mb0[2 * N + e] = mt.z;
  }
  __syncthreads();
  const d3 ra = mk(px[a], py[a], pz[a]);
  const double scale = (GP.interaction == CMTJB_INT_DIPOLE_NOMURA) ? 1.0 / (4.0 * M_PI) : 1.0;
  long long until_write = (R.write_every - (R.it_begin % R.write_every)) % R.write_every;
  long long sample = R.first_sample;
  int cur = 0;
  for (long long it = R.it_begin; it < R.it_end; ++it) {
    const double t = (double)it * R.dt;
    const double *mx = mb0 + cur * 3 * N, *my = mx + N, *mz = my + N;
    if (GP.interaction != CMTJB_INT_NULL) {
      d3 H = mk(0, 0, 0);
#pragma unroll 4
      for (int b = p; b < N; b += P)
        pair_accumulate(H, mk(ra.x - px[b], ra.y - py[b], ra.z - pz[b]), mk(mx[b], my[b], mz[b]), pf[b], b == a);
      for (int off = P >> 1; off > 0; off >>= 1) {
        H.x += __shfl_xor_sync(0xffffffffu, H.x, off);
        H.y += __shfl_xor_sync(0xffffffffu, H.y, off);
        H.z += __shfl_xor_sync(0xffffffffu, H.z, off);
      }
      if (p == 0 && a_raw < N) {
        Hs[a] = H.x;
        Hs[N + a] = H.y;
        Hs[2 * N + a] = H.z;
      }
      __syncthreads();
    }
    if (stepper) {
      // setLayerReservedInteractionField(topId, AxialDriver(H_extra)) (reservoir.hpp:145-146)
      const d3 H = (GP.interaction != CMTJB_INT_NULL) ? mk(Hs[e], Hs[N + e], Hs[2 * N + e]) : mk(0, 0, 0);
      d3 Hl[NL];
#pragma unroll
      for (int l = 0; l < NL; l++) Hl[l] = (l == GP.top) ? axpy(H, scale, Hc[l]) : Hc[l];
      coupled_step<NL, SOLVER, F>(M, R, i, it, t, S, Hl, NoExt());
      double *nx = mb0 + (cur ^ 1) * 3 * N;
      const d3 mt = layer_m2<NL>(S, GP.top);
      nx[e] = mt.x;
      nx[N + e] = mt.y;
      nx[2 * N + e] = mt.z;
      if (until_write == 0 && R.log) write_sample<NL>(M, R, S, sample, i);
    }
    if (until_write == 0) {
      sample++;
      until_write = R.write_every;
    }
    until_write--;
    cur ^= 1;
    __syncthreads();
  }
  if (stepper) {
    store_state<NL>(R, i, S);
    if (R.status) R.status[i] = nan_status<NL, SOLVER>(S);
  }
}

// GroupInteraction, one large group across the whole grid (cooperative launch, ONE grid-wide barrier per step).
//
// Decomposition: block k owns A = ceil(N / gridDim) consecutive elements and sums their pair fields over ALL N
// partners -- no reduction across blocks, and with gridDim = the SM count every SM carries the same 1/gridDim of
// the N^2 pairs.  Inside the block thread (a, p) = (tid % A, tid / A) walks partners p, p + P, ... (P = 512 / A lanes
// per element; the lanes of a warp mostly share the partner, so its seven shared-memory loads are broadcasts),
// the P partial fields meet in shared memory, and the A elements are then stepped on A CONSECUTIVE lanes (one
// element per lane, state in registers) -- not on one lane in P with the others parked.
//
// Shared memory holds the partners' x, y, z, prefactor and top-layer m: 7 x 8 B x TB.  Up to TB = 4096 partners
// the geometry is staged ONCE and only the three m rows are re-staged per step (N = 4096 exactly fills the 227 KB);
// larger arrays walk tiles of 4096 partners and re-stage all seven rows per tile.  The top-layer magnetisation is
// exchanged through a double-buffered global array.
constexpr int kGroupThreads = 512;   // general term set (128 registers)
constexpr int kGroupThreadsLean = 768;  // lean term set: 6 warps per sub-partition at <= 85 registers
constexpr int kGroupTileMax = 4096;

template <int NL, int SOLVER, uint32_t F, bool UNIFORM, int THREADS>
__global__ void __launch_bounds__(THREADS, 1) group_grid_kernel(const DevModel<NL> *__restrict__ models,
                                                                      const __grid_constant__ DevModel<NL> Mu,
                                                                      const __grid_constant__ GroupParams GP,
                                                                      const __grid_constant__ RunParams R) {
  cg::grid_group grid = cg::this_grid();
  extern __shared__ double sm[];
  const int N = GP.N, A = GP.A, P = GP.P;
  constexpr int TB = kGroupTileMax;  // compile-time row stride: the seven loads of a pair share one address register
  double *tx = sm, *ty = tx + TB, *tz = ty + TB, *tp = tz + TB;  // static rows
  double *tm = tp + TB;                                          // [3][TB] partners' m; reused for the P partial sums
  double *Hs = tm + 3 * TB;                                      // [3][A] reduced fields of this block's elements
  const int tid = threadIdx.x;
  const int a_loc = tid % A, p = tid / A;
  const int a0 = blockIdx.x * A;
  const bool pair_lane = p < P;                  // the last (blockDim - A P) threads have no pairs
  const int a_raw = a0 + a_loc;
  const int a = a_raw < N ? a_raw : N - 1;       // lanes past the end shadow element N-1 (their sums are dropped)
  const bool stepper = tid < A && a0 + tid < N;  // element a0 + tid is stepped by thread tid
  const int e = stepper ? a0 + tid : N - 1;
  const long long i = GP.inst * N + e;
  const DevModel<NL> &M = UNIFORM ? Mu : models[e];
  const bool resident = N <= TB;
  State<NL> S;
  d3 Hc[NL];
  if (stepper) {
    load_state<NL>(R, i, S);
#pragma unroll
    for (int l = 0; l < NL; l++) Hc[l] = constant_field(M.L[l], R, i);
    const d3 mt = layer_m2<NL>(S, GP.top);
    GP.mbuf[e] = mt.x;
    GP.mbuf[N + e] = mt.y;
    GP.mbuf[2 * N + e] = mt.z;
  }
  const d3 ra = mk(__ldg(GP.pos + a), __ldg(GP.pos + N + a), __ldg(GP.pos + 2 * N + a));
  const double scale = (GP.interaction == CMTJB_INT_DIPOLE_NOMURA) ? 1.0 / (4.0 * M_PI) : 1.0;
  const bool interacting = GP.interaction != CMTJB_INT_NULL;
  if (resident && interacting) {
    for (int k = tid; k < N; k += blockDim.x) {
      tx[k] = __ldg(GP.pos + k);
      ty[k] = __ldg(GP.pos + N + k);
      tz[k] = __ldg(GP.pos + 2 * N + k);
      tp[k] = __ldg(GP.pref + k);
    }
  }
  long long until_write = (R.write_every - (R.it_begin % R.write_every)) % R.write_every;
  long long sample = R.first_sample;
  int cur = 0;
  grid.sync();
  for (long long it = R.it_begin; it < R.it_end; ++it) {
    const double t = (double)it * R.dt;
    const double *mcur = GP.mbuf + (size_t)cur * 3 * N;
    if (interacting) {
      d3 H = mk(0, 0, 0);
      for (int b0 = 0; b0 < N; b0 += TB) {
        const int nb = min(TB, N - b0);
        if (b0 > 0) __syncthreads();  // the previous tile is still being read
        for (int k = tid; k < nb; k += blockDim.x) {
          if (!resident) {
            tx[k] = __ldg(GP.pos + b0 + k);
            ty[k] = __ldg(GP.pos + N + b0 + k);
            tz[k] = __ldg(GP.pos + 2 * N + b0 + k);
            tp[k] = __ldg(GP.pref + b0 + k);
          }
          tm[k] = __ldcg(mcur + b0 + k);  // written by other SMs last step: read through L2
          tm[TB + k] = __ldcg(mcur + N + b0 + k);
          tm[2 * TB + k] = __ldcg(mcur + 2 * N + b0 + k);
        }
        __syncthreads();
        if (pair_lane) {
          // two partners per iteration into separate accumulators: two independent dependency chains per thread
          d3 H2 = mk(0, 0, 0);
          int k = p;
          for (; k + P < nb; k += 2 * P) {
            const int k2 = k + P;
            pair_accumulate(H, mk(ra.x - tx[k], ra.y - ty[k], ra.z - tz[k]), mk(tm[k], tm[TB + k], tm[2 * TB + k]), tp[k], b0 + k == a);
            pair_accumulate(H2, mk(ra.x - tx[k2], ra.y - ty[k2], ra.z - tz[k2]), mk(tm[k2], tm[TB + k2], tm[2 * TB + k2]), tp[k2], b0 + k2 == a);
          }
          if (k < nb) pair_accumulate(H, mk(ra.x - tx[k], ra.y - ty[k], ra.z - tz[k]), mk(tm[k], tm[TB + k], tm[2 * TB + k]), tp[k], b0 + k == a);
          H = H + H2;
        }
      }
      __syncthreads();  // every lane is done with the m rows: their space takes the partial sums
      if (pair_lane) {
        double *q = tm + (size_t)(p * A + a_loc) * 3;
        q[0] = H.x;
        q[1] = H.y;
        q[2] = H.z;
      }
      __syncthreads();
      if (tid < 3 * A) {  // component c of element a_loc: sum over the P lanes in lane order (deterministic)
        const int c = tid / A, al = tid - c * A;
        double acc = 0.0;
        for (int q = 0; q < P; q++) acc += tm[(size_t)(q * A + al) * 3 + c];
        Hs[c * A + al] = acc;
      }
      __syncthreads();
    }
    if (stepper) {
      // setLayerReservedInteractionField(topId, AxialDriver(H_extra)) (reservoir.hpp:145-146)
      const d3 H = interacting ? mk(Hs[tid], Hs[A + tid], Hs[2 * A + tid]) : mk(0, 0, 0);
      d3 Hl[NL];
#pragma unroll
      for (int l = 0; l < NL; l++) Hl[l] = (l == GP.top) ? axpy(H, scale, Hc[l]) : Hc[l];
      coupled_step<NL, SOLVER, F>(M, R, i, it, t, S, Hl, NoExt());
      double *nx = GP.mbuf + (size_t)(cur ^ 1) * 3 * N;
      const d3 mt = layer_m2<NL>(S, GP.top);
      __stcg(nx + e, mt.x);
      __stcg(nx + N + e, mt.y);
      __stcg(nx + 2 * N + e, mt.z);
      if (until_write == 0 && R.log) write_sample<NL>(M, R, S, sample, i);
    }
    if (until_write == 0) {
      sample++;
      until_write = R.write_every;
    }
    until_write--;
    cur ^= 1;
    grid.sync();
  }
  if (stepper) {
    store_state<NL>(R, i, S);
    if (R.status) R.status[i] = nan_status<NL, SOLVER>(S);
  }
}

// ====================================================================================================
// GroupInteraction, one large group: two kernels per time step instead of one persistent cooperative kernel.
// EXPERIMENT, selected with CMTJB_GROUP_TWO_KERNEL=1; parity-tested, but measured SLOWER than the cooperative
// kernel on B200 at N = 4096 (79 vs 61 us per step): the pair kernel takes 45-48 us (FP64 pipe 61 % busy at
// 41 % occupancy) and the 4096-thread step kernel is pure single-thread latency (15-20 us).
//
// The cooperative form above inlines the junction's step next to the pair loop, which pins it at 128
// registers, 16 warps per SM and one grid barrier per step (ncu: FP64 pipe 39 % busy, the rest latency).
// Here the pair sum is its own kernel -- ~40 registers, every SM full, no shared memory, no barrier: a warp
// is 32 CONSECUTIVE elements a sharing one partner slice, so every partner load is a warp-wide broadcast
// served by L1 -- and the step is a second, small kernel (one thread per element, state in m_state).  The stream orders
// them; 2 launches per step cost less than the grid barrier they replace.  Partial sums are written per
// slice and added in slice order by the step kernel: deterministic, no atomics.
// ====================================================================================================
struct PairParams {
  int N, P;                // elements, partner slices
  const double *pos;       // device [3][N]
  const double *pref;      // device [N]
  const double *mx, *my, *mz;  // device [N] each: top-layer magnetisation at the start of the step (rows of m_state)
  double *Hpart;           // device [P][3][N] partial pair fields
};

__global__ void __launch_bounds__(256) group_pair_kernel(const __grid_constant__ PairParams Q) {
  const long long gt = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int N = Q.N;
  const int tiles = (N + 31) / 32;       // warps per slice
  const long long w = gt >> 5;
  const int p = (int)(w / tiles);
  if (p >= Q.P) return;
  const int a_raw = (int)(w - (long long)p * tiles) * 32 + (int)(gt & 31);
  const int a = a_raw < N ? a_raw : N - 1;
  const d3 ra = mk(__ldg(Q.pos + a), __ldg(Q.pos + N + a), __ldg(Q.pos + 2 * N + a));
  const int per = (N + Q.P - 1) / Q.P;
  const int b0 = p * per, b1 = (b0 + per < N) ? b0 + per : N;
  d3 H = mk(0, 0, 0);
#pragma unroll 4
  for (int b = b0; b < b1; b++) {
    const d3 rb = mk(__ldg(Q.pos + b), __ldg(Q.pos + N + b), __ldg(Q.pos + 2 * N + b));
    const d3 mb = mk(Q.mx[b], Q.my[b], Q.mz[b]);  // rewritten by the step kernel between launches: plain loads
    pair_accumulate(H, ra - rb, mb, __ldg(Q.pref + b), b == a);
  }
  if (a_raw < N) {
    double *o = Q.Hpart + (size_t)p * 3 * N + a;
    o[0] = H.x;
    o[N] = H.y;
    o[2 * N] = H.z;
  }
}

// CMTJB_INT_TABLE: a tabulated affine interaction, H_a += c[a][b] + T[a][b] m_b (include/cmtj_b200.h).  Same decomposition
// as group_pair_kernel (a warp = 32 consecutive elements a sharing one partner slice); the table is [b][12][a], so the 12
// loads of a pair are coalesced over the warp's elements.  The a == b entries are zero by construction.
struct PairTableParams {
  int N, P;
  const double *table;         // device [N b][12][N a]
  const double *mx, *my, *mz;  // device [N] each
  double *Hpart;               // device [P][3][N]
};
__global__ void __launch_bounds__(256) group_pair_table_kernel(const __grid_constant__ PairTableParams Q) {
  const long long gt = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int N = Q.N;
  const int tiles = (N + 31) / 32;
  const long long w = gt >> 5;
  const int p = (int)(w / tiles);
  if (p >= Q.P) return;
  const int a_raw = (int)(w - (long long)p * tiles) * 32 + (int)(gt & 31);
  const int a = a_raw < N ? a_raw : N - 1;
  const int per = (N + Q.P - 1) / Q.P;
  const int b0 = p * per, b1 = (b0 + per < N) ? b0 + per : N;
  d3 H = mk(0, 0, 0);
  for (int b = b0; b < b1; b++) {
    const double *t = Q.table + (size_t)b * 12 * N + a;
    const d3 m = mk(Q.mx[b], Q.my[b], Q.mz[b]);
    double e[12];
#pragma unroll
    for (int k = 0; k < 12; k++) e[k] = __ldg(t + (size_t)k * N);
    // the oracle's operation order: offset + (row . m), no contraction across the parenthesis
    H.x += e[9] + __dadd_rn(__dadd_rn(__dmul_rn(e[0], m.x), __dmul_rn(e[1], m.y)), __dmul_rn(e[2], m.z));
    H.y += e[10] + __dadd_rn(__dadd_rn(__dmul_rn(e[3], m.x), __dmul_rn(e[4], m.y)), __dmul_rn(e[5], m.z));
    H.z += e[11] + __dadd_rn(__dadd_rn(__dmul_rn(e[6], m.x), __dmul_rn(e[7], m.y)), __dmul_rn(e[8], m.z));
  }
  if (a_raw < N) {
    double *o = Q.Hpart + (size_t)p * 3 * N + a;
    o[0] = H.x;
    o[N] = H.y;
    o[2 * N] = H.z;
  }
}

// ---- peer-memory exchange of a sharded array (no collective library, no host in the time loop) ---------------------------
// Every rank owns an exchange AREA in its own device memory that all ranks of the box can address (CUDA IPC over NVLink):
//     double             X[2][world][3][per]   top-layer m of every element, double-buffered by step parity
//     unsigned long long arrived[world]        arrived[r] = number of slices rank r has delivered into THIS area
//     unsigned int       blocks_done, error
// A rank's step kernel stores its new m rows into X[(s + 1) % 2][rank] of EVERY area (remote stores for the peers') and the
// last of its blocks to finish -- system-scope fences on both sides of a block counter -- raises arrived[rank] to s + 2 in
// every area.  Before the pair kernel of step s a one-warp kernel waits until arrived[r] >= s + 1 for all r in its OWN area.
// Buffer parity is safe with two buffers: a rank writes X[(s + 1) % 2] only after its own wait for step s, i.e. after every
// peer has finished step s - 1, whose pair kernel was the last reader of that buffer.
constexpr int kMaxPeers = 16;
struct PeerPush {
  int n;                               // ranks (0: no peer exchange)
  int rank;
  double *x[kMaxPeers];                // this rank's rows [3][per] of the buffer being written, in every area
  unsigned long long *arrived[kMaxPeers];  // &arrived[rank] in every area
  unsigned long long epoch;            // value to publish
  unsigned int *blocks_done;           // own area
};

__device__ __forceinline__ void peer_signal(const PeerPush &P) {
  // every thread's remote stores are ordered before its block's increment of the counter ...
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int total = gridDim.x * gridDim.y;
    if (atomicAdd(P.blocks_done, 1u) == total - 1) {
      *P.blocks_done = 0;  // next launch (stream-ordered)
      __threadfence_system();  // ... and the last block's flag stores after everybody's data
      for (int r = 0; r < P.n; r++) *reinterpret_cast<volatile unsigned long long *>(P.arrived[r]) = P.epoch;
    }
  }
}

// initial state: the top-layer rows of m_state into X[0][rank] of every area
__global__ void __launch_bounds__(128) group_peer_publish_kernel(const double *__restrict__ mx, const double *__restrict__ my,
                                                                 const double *__restrict__ mz, int own, int per,
                                                                 const __grid_constant__ PeerPush P) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a < own) {
    const double x = mx[a], y = my[a], z = mz[a];
    for (int r = 0; r < P.n; r++) {
      P.x[r][a] = x;
      P.x[r][per + a] = y;
      P.x[r][2 * per + a] = z;
    }
  }
  peer_signal(P);
}

// wait until every rank has delivered slice number `epoch` into this area (one thread per rank); gives up after timeout_ns
__global__ void group_peer_wait_kernel(const unsigned long long *arrived, int world, unsigned long long epoch, unsigned int *error,
                                       unsigned long long timeout_ns) {
  const int r = threadIdx.x;
  if (r >= world) return;
  if (*reinterpret_cast<volatile unsigned int *>(error)) return;  // an earlier wait of this call already gave up: do not wait again
  const volatile unsigned long long *f = arrived + r;
  unsigned long long t0 = 0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  unsigned int spins = 0;
  while (*f < epoch) {
    if ((++spins & 1023u) == 0) {
      unsigned long long t1;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
      if (t1 - t0 > timeout_ns) {
        atomicExch(error, 1u + (unsigned)r);
        break;
      }
    }
  }
  __threadfence_system();  // the data behind the flag is read after this point
}

struct GroupStepParams {
  int N, P, interaction, top;
  long long inst;          // instance of the batch (state rows inst*N .. inst*N + N - 1)
  long long it, sample;    // this step's index and, if `write`, the log row it produces
  int write;
  const double *Hpart;     // [P][3][N]
  double *xout;            // sharded arrays: this rank's rows [3][xpitch] of the exchange buffer (else null)
  int xpitch;
  PeerPush push;           // sharded arrays, peer-memory exchange: where else the rows go, and whom to tell (n = 0: nobody)
};

// ---- one array sharded over several GPUs (reservoir.hpp:129-152 with the elements split into contiguous ranges) ---------
// Rank r owns elements [e_lo, e_lo + own); the top-layer magnetisation of EVERY element lives in an exchange buffer
// [world][3][per] (per = elements per rank, the last rank's tail unused) that the host all-gathers after each step
// (NCCL, in place: each rank's step kernel writes its own rows).  Pair sums: a warp is 32 consecutive own elements
// sharing one slice of the N partners, partner loads are warp-wide broadcasts; partial sums per slice, added in slice
// order by the step kernel (deterministic, independent of the number of ranks because slices are ranges of the global index).
struct PairShardParams {
  int N, P, own, e_lo, per;
  const double *pos;   // device [3][N]
  const double *pref;  // device [N]
  const double *xbuf;  // device [world][3][per]
  double *Hpart;       // device [P][3][own]
};

__global__ void __launch_bounds__(256) group_pair_shard_kernel(const __grid_constant__ PairShardParams Q) {
  const long long gt = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int N = Q.N, own = Q.own, per = Q.per;
  const int tiles = (own + 31) / 32;
  const long long w = gt >> 5;
  const int p = (int)(w / tiles);
  if (p >= Q.P) return;
  const int a_raw = (int)(w - (long long)p * tiles) * 32 + (int)(gt & 31);
  const int a = Q.e_lo + (a_raw < own ? a_raw : own - 1);
  const d3 ra = mk(__ldg(Q.pos + a), __ldg(Q.pos + N + a), __ldg(Q.pos + 2 * N + a));
  const int len = (N + Q.P - 1) / Q.P;
  const int b0 = p * len, b1 = (b0 + len < N) ? b0 + len : N;
  int r = b0 / per, o = b0 - r * per;
  const double *base = Q.xbuf + (size_t)r * 3 * per;
  d3 H = mk(0, 0, 0);
  for (int b = b0; b < b1; b++) {
    const d3 rb = mk(__ldg(Q.pos + b), __ldg(Q.pos + N + b), __ldg(Q.pos + 2 * N + b));
    const d3 mb = mk(base[o], base[per + o], base[2 * per + o]);  // rewritten between launches: plain loads
    pair_accumulate(H, ra - rb, mb, __ldg(Q.pref + b), b == a);
    if (++o == per) {
      o = 0;
      base += (size_t)3 * per;
    }
  }
  if (a_raw < own) {
    double *q = Q.Hpart + (size_t)p * 3 * own + a_raw;
    q[0] = H.x;
    q[own] = H.y;
    q[2 * own] = H.z;
  }
}

// Launched once per time step: its parameter block is kept to a few words (the models and the run
// parameters are read through pointers) so that the two launches of a step cost the host less than the
// kernels cost the device.
template <int NL, int SOLVER, uint32_t F, bool UNIFORM>
__global__ void __launch_bounds__(128) group_step_kernel(const DevModel<NL> *__restrict__ models, const RunParams *__restrict__ Rp,
                                                         const __grid_constant__ GroupStepParams G) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a < G.N) {
    const int N = G.N;
    const long long i = G.inst * N + a;
    const RunParams &R = *Rp;
    const DevModel<NL> &M = models[UNIFORM ? 0 : a];
    State<NL> S;
    load_state<NL>(R, i, S);
    d3 H = mk(0, 0, 0);
    if (G.interaction != CMTJB_INT_NULL) {
      for (int p = 0; p < G.P; p++) {
        const double *q = G.Hpart + (size_t)p * 3 * N + a;
        H = H + mk(q[0], q[N], q[2 * N]);
      }
      if (G.interaction == CMTJB_INT_DIPOLE_NOMURA) H = H * (1.0 / (4.0 * M_PI));
    }
    d3 Hl[NL];
  #pragma unroll
    for (int l = 0; l < NL; l++) {
      const d3 Hc = constant_field(M.L[l], R, i);
      Hl[l] = (l == G.top) ? Hc + H : Hc;  // setLayerReservedInteractionField(topId, ...) (reservoir.hpp:145-146)
    }
    coupled_step<NL, SOLVER, F>(M, R, i, G.it, (double)G.it * R.dt, S, Hl, NoExt());
    store_state<NL>(R, i, S);  // the next pair kernel reads the top layer's rows from here
    if (G.xout) {
      const d3 mt = layer_m2<NL>(S, G.top);
      G.xout[a] = mt.x;
      G.xout[G.xpitch + a] = mt.y;
      G.xout[2 * G.xpitch + a] = mt.z;
    }
    if (G.push.n > 0) {
      const d3 mt = layer_m2<NL>(S, G.top);
      for (int r = 0; r < G.push.n; r++) {  // NVLink stores into the peers' areas (and a local one into our own)
        G.push.x[r][a] = mt.x;
        G.push.x[r][G.xpitch + a] = mt.y;
        G.push.x[r][2 * G.xpitch + a] = mt.z;
      }
    }
    if (G.write && R.log) write_sample<NL>(M, R, S, G.sample, i);
    if (R.status) {
      const int st = nan_status<NL, SOLVER>(S);
      if (G.it == 0 || st) R.status[i] = st;
    }
  }
  if (G.push.n > 0) peer_signal(G.push);  // every thread of the block, owners of an element or not
}

}  // namespace cmtjb
