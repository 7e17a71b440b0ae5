// cabi.cu -- extern "C" entry points of libcmtj_b200.so (include/cmtj_b200.h).
//
// Host side of the library: validates a raw junction model the way the reference's
// constructors / runSimulation do, derives the constant-bank DevModel, picks the leanest
// compiled kernel variant that covers the model's terms, and launches it.  No torch, no
// Python, no C++ types across the boundary.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "dev_model.h"
#include "llgs_kernels.cuh"

namespace cmtjb {
// variant tables, one per translation unit (variants_*.cu)
int variants_rk4(const Variant **out);
int variants_heun(const Variant **out);
int variants_euler_heun(const Variant **out);
int variants_dopri(const Variant **out);

namespace {

thread_local std::string g_err;

int fail(int code, const std::string &msg) {
  g_err = msg;
  return code;
}
int cuda_fail(cudaError_t e, const char *what) {
  return fail(CMTJB_ECUDA, std::string("CUDA error in ") + what + ": " + cudaGetErrorString(e));
}
#define CUDA_TRY(expr)                                   \
  do {                                                   \
    cudaError_t _e = (expr);                             \
    if (_e != cudaSuccess) return cuda_fail(_e, #expr);  \
  } while (0)

bool slot_dynamic(const cmtjb_driver_t &d) { return d.update != CMTJB_UPD_CONSTANT; }

// Mirrors Junction::runSimulation's guards (junction.hpp:1952-1959), getSolver's solver
// override (:1891-1931) and the run-time throws reachable from the hot loop.
int make_plan(const cmtjb_junction_t *jn, const cmtjb_layer_t *layers, const cmtjb_run_args_t *a, cmtjb_plan_t *p) {
  if (!jn || !layers || !a || !p) return fail(CMTJB_EINVAL, "null argument");
  if (jn->n_layers <= 0) return fail(CMTJB_EINVAL, "Passed a zero length Layer vector!");
  if (jn->n_layers > CMTJB_MAX_LAYERS) return fail(CMTJB_ENOTSUP, "too many layers for this build");
  if (a->timeStep > a->writeFrequency)
    return fail(CMTJB_ERUNTIME, "The time step cannot be larger than write frequency!");
  std::memset(p, 0, sizeof(*p));
  p->n_layers = jn->n_layers;
  p->total_iterations = (int64_t) static_cast<unsigned int>(a->totalTime / a->timeStep);
  p->write_every = (int64_t) static_cast<unsigned int>(a->writeFrequency / a->timeStep);
  int mode = a->solverMode;
  if (mode != CMTJB_RK4 && mode != CMTJB_HEUN && mode != CMTJB_EULER_HEUN && mode != CMTJB_DORMAND_PRINCE)
    return fail(CMTJB_EINVAL, "unknown solver mode");
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
    for (int s = 0; s < CMTJB_NSLOTS; s++) {
      const int u = L.drivers[s].update;
      if (u == CMTJB_UPD_CUSTOM)
        return fail(CMTJB_ENOTSUP, "custom (Python callback) drivers are not supported on the device path");
      if (u < CMTJB_UPD_CONSTANT || u > CMTJB_UPD_GAUSSSTEP) return fail(CMTJB_EINVAL, "unknown driver update type");
    }
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
    auto count_upto = [&](int64_t n) { return n <= 0 ? (int64_t)0 : (n - 1) / p->write_every + 1; };
    p->n_samples = count_upto(p->total_iterations) - count_upto(a->start_iteration);
  }
  return CMTJB_OK;
}

bool any_nonzero(const double *p, int n) {
  for (int i = 0; i < n; i++)
    if (p[i] != 0.0) return true;
  return false;
}

// which terms does this model need compiled in?
uint32_t needed_features(const cmtjb_junction_t *jn, const cmtjb_layer_t *layers, const cmtjb_override_t *ovr, int n_ovr,
                         const cmtjb_run_args_t *a, int solver) {
  uint32_t F = 0;
  const int n = jn->n_layers;
  // a slot is "live" if its constant is non-zero, it is time dependent, or it is overridden
  auto live = [&](int l, int s) {
    const cmtjb_driver_t &d = layers[l].drivers[s];
    if (d.constantValue != 0.0 || slot_dynamic(d)) return true;
    for (int k = 0; k < n_ovr; k++)
      if ((ovr[k].layer < 0 || ovr[k].layer == l) && ovr[k].slot == s) return true;
    return false;
  };
  for (int l = 0; l < n; l++) {
    const cmtjb_layer_t &L = layers[l];
    if (live(l, CMTJB_SLOT_K2)) F |= F_K2;
    for (int s = CMTJB_SLOT_IEC_TOP; s <= CMTJB_SLOT_IECQ_BOTTOM; s++)
      if (live(l, s)) F |= F_IEC;
    for (int s = CMTJB_SLOT_IDMI_TOP_X; s <= CMTJB_SLOT_IDMI_BOTTOM_Z; s++)
      if (live(l, s)) F |= F_IDMI;
    if (any_nonzero(L.dipoleBottom, 9) || any_nonzero(L.dipoleTop, 9)) F |= F_DIPOLE;
    const double *T = L.demagTensor;
    if (T[1] != 0 || T[2] != 0 || T[3] != 0 || T[5] != 0 || T[6] != 0 || T[7] != 0) F |= F_DEMAG_FULL;
    if (L.includeSTT) F |= F_STT;
    else if (L.includeSOT) F |= F_SOT;
    if (solver == CMTJB_HEUN || solver == CMTJB_EULER_HEUN) {
      if (L.temperatureSet) F |= F_THERMAL;
      if (L.noiseScale != 0 && a->onef_buffer) F |= F_ONEF;
    }
    for (int s = 0; s < CMTJB_NSLOTS; s++)
      if (slot_dynamic(L.drivers[s])) F |= F_DYN;
  }
  for (int k = 0; k < n_ovr; k++)
    if (!(ovr[k].field == CMTJB_FIELD_CONSTANT && ovr[k].slot < CMTJB_SLOT_K1)) F |= F_OVR;
  return F;
}

const Variant *pick_variant(int nl, int solver, uint32_t need) {
  const Variant *tab = nullptr;
  int n = 0;
  switch (solver) {
  case CMTJB_RK4: n = variants_rk4(&tab); break;
  case CMTJB_HEUN: n = variants_heun(&tab); break;
  case CMTJB_EULER_HEUN: n = variants_euler_heun(&tab); break;
  case CMTJB_DORMAND_PRINCE: n = variants_dopri(&tab); break;
  }
  const Variant *best = nullptr;
  int best_extra = 1 << 30;
  for (int i = 0; i < n; i++) {
    const Variant &v = tab[i];
    if (v.nl != nl || (v.features & need) != need) continue;
    const int extra = __builtin_popcount(v.features & ~need);
    if (extra < best_extra) {
      best_extra = extra;
      best = &v;
    }
  }
  return best;
}

template <int NL> struct ModelBox {
  DevModel<NL> M;
};

void derive_layer(const cmtjb_junction_t &J, const cmtjb_layer_t &L, DevLayer &D) {
  std::memset(&D, 0, sizeof(D));
  const double mu0 = J.MAGNETIC_PERMEABILITY, gyro = J.GYRO;
  for (int s = 0; s < CMTJB_NSLOTS; s++) {
    const cmtjb_driver_t &d = L.drivers[s];
    D.cst[s] = d.constantValue;
    DevDriver &o = D.drv[s];
    o.type = d.update;
    o.amplitude = d.amplitude;
    o.frequency = d.frequency;
    o.phase = d.phase;
    o.period = d.period;
    o.cycle = d.cycle;
    o.timeStart = d.timeStart;
    o.timeStop = d.timeStop;
    o.edgeTime = d.edgeTime;
    o.steadyTime = d.steadyTime;
    if (slot_dynamic(d)) D.dynmask |= (1u << s);
    for (int f = 0; f < 10; f++) D.ovidx[s][f] = -1;
  }
  for (int c = 0; c < 3; c++) {
    D.anis[c] = L.anis[c];
    D.ref[c] = L.referenceLayer[c];
    D.ref2[c] = L.secondaryReferenceLayer[c];
  }
  D.cK1 = 2.0 / L.Ms;
  D.cK2 = 4.0 / L.Ms;
  D.inv_Mst = 1.0 / (L.Ms * L.thickness);
  const double Msmu = L.Ms / mu0;
  for (int k = 0; k < 9; k++) {
    D.demag[k] = L.demagTensor[k] * Msmu;
    D.dipB[k] = L.dipoleBottom[k] * Msmu;
    D.dipT[k] = L.dipoleTop[k] * Msmu;
  }
  D.alpha = L.damping;
  D.llg = -gyro / (1.0 + L.damping * L.damping);
  D.aJ_c = J.HBAR / (J.ELECTRON_CHARGE * L.Ms * L.thickness);
  D.lam = L.SlonczewskiSpacerLayerParameter;
  D.lam2 = D.lam * D.lam;
  D.P = L.spinPolarisation;
  D.beta = L.beta;
  D.kappa = L.kappa;
  D.flT = L.fieldLikeTorque;
  D.dlT = L.dampingLikeTorque;
  const double V = L.cellSurface * L.thickness;
  D.sigma_c = (V != 0.0) ? (2.0 * L.damping * J.BOLTZMANN_CONST) / (L.Ms * V * gyro) : 0.0;
  D.Rx0 = L.Rx0;
  D.Ry0 = L.Ry0;
  D.AMR_X = L.AMR_X;
  D.SMR_X = L.SMR_X;
  D.AHE_half = 0.5 * L.AHE;
  D.AMRSMR_Y = L.AMR_Y + L.SMR_Y;
  D.abs_tol = L.abs_tol;
  D.rel_tol = L.rel_tol;
  D.max_factor = L.max_factor;
  D.min_factor = L.min_factor;
  D.safety_factor = L.safety_factor;
  D.ki = L.ki;
  D.kp = L.kp;
  D.kd = L.kd;
  D.prev_error_ratio = L.prev_error_ratio;
  D.integral_error = L.integral_error;
  D.refType = L.referenceType;
  D.stt = L.includeSTT ? 1 : 0;
  D.sot = (!L.includeSTT && L.includeSOT) ? 1 : 0;
  D.dynamicSOT = L.dynamicSOT ? 1 : 0;
  D.altSTT = L.alternativeSTT ? 1 : 0;
  D.thermal = L.temperatureSet ? 1 : 0;
  D.onef = (L.noiseScale != 0) ? 1 : 0;
  D.pid = L.use_pid_control ? 1 : 0;
}

template <int NL>
int run_typed(const cmtjb_junction_t *jn, const cmtjb_layer_t *layers, const cmtjb_override_t *ovr, int n_ovr,
              const cmtjb_run_args_t *a, const cmtjb_plan_t &plan, const Variant *v, double *m_state, double *log,
              int32_t *status, int32_t *n_samples_out, double *time_log, cudaStream_t stream) {
  static_assert(sizeof(DevModel<NL>) + sizeof(RunParams) < 32000, "kernel parameters exceed the 32 KB limit");
  DevModel<NL> M;
  std::memset(&M, 0, sizeof(M));
  for (int l = 0; l < NL; l++) derive_layer(*jn, layers[l], M.L[l]);
  M.Rp = jn->Rp;
  M.Rhalf = (jn->Rap - jn->Rp) / 2.0;
  M.mr_mode = jn->mr_mode;
  M.n_layers = NL;
  RunParams R;
  std::memset(&R, 0, sizeof(R));
  for (int k = 0; k < n_ovr; k++) {
    R.ovr[k] = ovr[k].values;
    for (int l = 0; l < NL; l++) {
      if (ovr[k].layer >= 0 && ovr[k].layer != l) continue;
      M.L[l].ovmask |= (1u << ovr[k].slot);
      M.L[l].ovidx[ovr[k].slot][ovr[k].field] = (int8_t)k;
    }
  }
  R.m_state = m_state;
  R.log = (a->log_mode == CMTJB_LOG_FULL) ? log : nullptr;
  R.status = status;
  R.n_samples_out = n_samples_out;
  R.time_log = time_log;
  R.onef = a->onef_buffer;
  R.B = a->n_instances;
  R.inst_offset = a->instance_offset;
  R.n_samples = plan.n_samples;
  R.it_begin = a->start_iteration;
  R.it_end = plan.total_iterations;
  R.total_iterations = plan.total_iterations;
  R.write_every = plan.write_every > 0 ? plan.write_every : 1;
  R.first_sample = 0;
  R.dt = a->timeStep;
  R.sqrt_dt = std::sqrt(a->timeStep);
  R.total_time = a->totalTime;
  R.write_frequency = a->writeFrequency;
  R.seed = a->seed;
  if (R.B <= 0) return CMTJB_OK;
  cudaError_t e = v->launch(&M, &R, stream);
  if (e != cudaSuccess) return cuda_fail(e, v->name);
  return CMTJB_OK;
}

int check_overrides(const cmtjb_junction_t *jn, const cmtjb_override_t *ovr, int n_ovr) {
  if (n_ovr < 0 || n_ovr > kMaxOverrides) return fail(CMTJB_EINVAL, "too many per-instance overrides (max 24)");
  if (n_ovr > 0 && !ovr) return fail(CMTJB_EINVAL, "null override table");
  for (int k = 0; k < n_ovr; k++) {
    if (ovr[k].layer >= jn->n_layers || ovr[k].slot < 0 || ovr[k].slot >= CMTJB_NSLOTS || ovr[k].field < 0 ||
        ovr[k].field > CMTJB_FIELD_STEADYTIME || !ovr[k].values)
      return fail(CMTJB_EINVAL, "override out of range");
  }
  return CMTJB_OK;
}

int run_device_impl(const cmtjb_junction_t *jn, const cmtjb_layer_t *layers, const cmtjb_override_t *ovr, int n_ovr,
                    const cmtjb_run_args_t *a, double *m_state, double *log, int32_t *status, int32_t *n_samples_out,
                    double *time_log, cudaStream_t stream, cmtjb_plan_t *plan_out) {
  cmtjb_plan_t plan;
  int rc = make_plan(jn, layers, a, &plan);
  if (rc) return rc;
  rc = check_overrides(jn, ovr, n_ovr);
  if (rc) return rc;
  if (!m_state) return fail(CMTJB_EINVAL, "m_state is null");
  const uint32_t need = needed_features(jn, layers, ovr, n_ovr, a, plan.effective_solver);
  const Variant *v = pick_variant(jn->n_layers, plan.effective_solver, need);
  if (!v) {
    char buf[160];
    std::snprintf(buf, sizeof(buf), "no compiled kernel variant for %d layers, solver %d, features 0x%x", jn->n_layers,
                  plan.effective_solver, need);
    return fail(CMTJB_ENOTSUP, buf);
  }
  plan.feature_mask = (int32_t)v->features;
  std::snprintf(plan.variant, sizeof(plan.variant), "%s", v->name);
  if (plan_out) *plan_out = plan;
  switch (jn->n_layers) {
  case 1: return run_typed<1>(jn, layers, ovr, n_ovr, a, plan, v, m_state, log, status, n_samples_out, time_log, stream);
  case 2: return run_typed<2>(jn, layers, ovr, n_ovr, a, plan, v, m_state, log, status, n_samples_out, time_log, stream);
  case 3: return run_typed<3>(jn, layers, ovr, n_ovr, a, plan, v, m_state, log, status, n_samples_out, time_log, stream);
  case 4: return run_typed<4>(jn, layers, ovr, n_ovr, a, plan, v, m_state, log, status, n_samples_out, time_log, stream);
  default: return fail(CMTJB_ENOTSUP, "junctions with more than 4 layers need the warp-per-instance kernel (not built yet)");
  }
}

// ---- small utility kernels ---------------------------------------------------------------------
__global__ void eval_driver_kernel(const __grid_constant__ DevLayer Lr, const __grid_constant__ RunParams R,
                                   const double *times, long long n, double *out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = slot_value<F_DYN>(Lr, R, 0, times[i], 0);
}

// 8 independent DFMA chains per thread; 2 flop per DFMA
__global__ void __launch_bounds__(256) fp64_peak_kernel(double *out, long long iters, double a, double b) {
  double x0 = threadIdx.x * 1e-3, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (long long k = 0; k < iters; k++) {
#pragma unroll
    for (int u = 0; u < 8; u++) {
      x0 = fma(x0, a, b);
      x1 = fma(x1, a, b);
      x2 = fma(x2, a, b);
      x3 = fma(x3, a, b);
      x4 = fma(x4, a, b);
      x5 = fma(x5, a, b);
      x6 = fma(x6, a, b);
      x7 = fma(x7, a, b);
    }
  }
  const double s = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
  if (s == 123.456) out[0] = s;
}

}  // namespace
}  // namespace cmtjb

using namespace cmtjb;

extern "C" {

int cmtjb_abi_version(void) { return CMTJB_ABI_VERSION; }
const char *cmtjb_last_error(void) { return g_err.c_str(); }

int cmtjb_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

int cmtjb_plan(const cmtjb_junction_t *jn, const cmtjb_layer_t *layers, const cmtjb_run_args_t *args, cmtjb_plan_t *plan) {
  int rc = make_plan(jn, layers, args, plan);
  if (rc) return rc;
  const uint32_t need = needed_features(jn, layers, nullptr, 0, args, plan->effective_solver);
  const Variant *v = pick_variant(jn->n_layers, plan->effective_solver, need);
  if (v) {
    plan->feature_mask = (int32_t)v->features;
    std::snprintf(plan->variant, sizeof(plan->variant), "%s", v->name);
  }
  return CMTJB_OK;
}

int cmtjb_run_device(const cmtjb_junction_t *jn, const cmtjb_layer_t *layers, const cmtjb_override_t *overrides,
                     int32_t n_overrides, const cmtjb_run_args_t *args, double *m_state, double *log, int32_t *status,
                     int32_t *n_samples_out, double *time_log, int device, void *stream) {
  CUDA_TRY(cudaSetDevice(device));
  return run_device_impl(jn, layers, overrides, n_overrides, args, m_state, log, status, n_samples_out, time_log,
                         (cudaStream_t)stream, nullptr);
}

int cmtjb_run_host(const cmtjb_junction_t *jn, const cmtjb_layer_t *layers, const cmtjb_override_t *overrides,
                   int32_t n_overrides, const cmtjb_run_args_t *args, double *m_state, double *log, int32_t *status,
                   int32_t *n_samples_out, double *time_log, int device) {
  cmtjb_plan_t plan;
  int rc = make_plan(jn, layers, args, &plan);
  if (rc) return rc;
  rc = check_overrides(jn, overrides, n_overrides);
  if (rc) return rc;
  if (!m_state) return fail(CMTJB_EINVAL, "m_state is null");
  CUDA_TRY(cudaSetDevice(device));
  const long long B = args->n_instances;
  if (B <= 0) return CMTJB_OK;
  const size_t nm = (size_t)jn->n_layers * 3 * B * sizeof(double);
  const bool want_log = (args->log_mode == CMTJB_LOG_FULL) && log && plan.n_samples > 0;
  const size_t nlog = want_log ? (size_t)plan.n_channels * plan.n_samples * B * sizeof(double) : 0;
  std::vector<void *> to_free;
  auto dalloc = [&](size_t bytes, void **p) -> cudaError_t {
    cudaError_t e = cudaMalloc(p, bytes ? bytes : 8);
    if (e == cudaSuccess) to_free.push_back(*p);
    return e;
  };
  struct Cleanup {
    std::vector<void *> &v;
    ~Cleanup() {
      for (void *p : v) cudaFree(p);
    }
  } cleanup{to_free};

  double *d_m = nullptr, *d_log = nullptr, *d_tlog = nullptr, *d_onef = nullptr;
  int32_t *d_status = nullptr, *d_ns = nullptr;
  CUDA_TRY(dalloc(nm, (void **)&d_m));
  CUDA_TRY(cudaMemcpy(d_m, m_state, nm, cudaMemcpyHostToDevice));
  if (want_log) CUDA_TRY(dalloc(nlog, (void **)&d_log));
  CUDA_TRY(dalloc(B * sizeof(int32_t), (void **)&d_status));
  CUDA_TRY(cudaMemset(d_status, 0, B * sizeof(int32_t)));
  if (plan.variable_samples) {
    CUDA_TRY(dalloc(B * sizeof(int32_t), (void **)&d_ns));
    if (want_log && time_log) CUDA_TRY(dalloc((size_t)plan.n_samples * B * sizeof(double), (void **)&d_tlog));
  }
  std::vector<cmtjb_override_t> d_ovr(overrides, overrides + n_overrides);
  for (int k = 0; k < n_overrides; k++) {
    double *p = nullptr;
    CUDA_TRY(dalloc(B * sizeof(double), (void **)&p));
    CUDA_TRY(cudaMemcpy(p, overrides[k].values, B * sizeof(double), cudaMemcpyHostToDevice));
    d_ovr[k].values = p;
  }
  cmtjb_run_args_t a = *args;
  if (args->onef_buffer) {
    const size_t nb = (size_t)jn->n_layers * 3 * plan.total_iterations * B * sizeof(double);
    CUDA_TRY(dalloc(nb, (void **)&d_onef));
    CUDA_TRY(cudaMemcpy(d_onef, args->onef_buffer, nb, cudaMemcpyHostToDevice));
    a.onef_buffer = d_onef;
  }
  rc = run_device_impl(jn, layers, d_ovr.data(), n_overrides, &a, d_m, d_log, d_status, d_ns, d_tlog, nullptr, nullptr);
  if (rc) return rc;
  CUDA_TRY(cudaDeviceSynchronize());
  CUDA_TRY(cudaMemcpy(m_state, d_m, nm, cudaMemcpyDeviceToHost));
  if (want_log) CUDA_TRY(cudaMemcpy(log, d_log, nlog, cudaMemcpyDeviceToHost));
  if (d_tlog) CUDA_TRY(cudaMemcpy(time_log, d_tlog, (size_t)plan.n_samples * B * sizeof(double), cudaMemcpyDeviceToHost));
  if (d_ns && n_samples_out) CUDA_TRY(cudaMemcpy(n_samples_out, d_ns, B * sizeof(int32_t), cudaMemcpyDeviceToHost));
  std::vector<int32_t> st_local;
  int32_t *st = status;
  if (!st) {
    st_local.resize(B);
    st = st_local.data();
  }
  CUDA_TRY(cudaMemcpy(st, d_status, B * sizeof(int32_t), cudaMemcpyDeviceToHost));
  for (long long i = 0; i < B; i++)
    if (st[i]) return fail(CMTJB_ENAN, "NAN magnetisation");
  return CMTJB_OK;
}

int cmtjb_eval_driver_host(const cmtjb_driver_t *driver, const double *times, int64_t n, double *out, int device) {
  if (!driver || !times || !out) return fail(CMTJB_EINVAL, "null argument");
  if (driver->update == CMTJB_UPD_CUSTOM) return fail(CMTJB_ENOTSUP, "custom drivers are host-only");
  CUDA_TRY(cudaSetDevice(device));
  cmtjb_junction_t J;
  std::memset(&J, 0, sizeof(J));
  J.MAGNETIC_PERMEABILITY = 1;
  J.GYRO = 1;
  cmtjb_layer_t L;
  std::memset(&L, 0, sizeof(L));
  L.Ms = 1;
  L.thickness = 1;
  L.drivers[0] = *driver;
  static DevLayer D;  // large; keep off the stack
  derive_layer(J, L, D);
  RunParams R;
  std::memset(&R, 0, sizeof(R));
  R.B = 1;
  double *d_t = nullptr, *d_o = nullptr;
  CUDA_TRY(cudaMalloc(&d_t, n * sizeof(double)));
  CUDA_TRY(cudaMalloc(&d_o, n * sizeof(double)));
  cudaMemcpy(d_t, times, n * sizeof(double), cudaMemcpyHostToDevice);
  eval_driver_kernel<<<(unsigned)((n + 127) / 128), 128>>>(D, R, d_t, n, d_o);
  cudaError_t e = cudaDeviceSynchronize();
  if (e == cudaSuccess) e = cudaMemcpy(out, d_o, n * sizeof(double), cudaMemcpyDeviceToHost);
  cudaFree(d_t);
  cudaFree(d_o);
  if (e != cudaSuccess) return cuda_fail(e, "eval_driver_kernel");
  return CMTJB_OK;
}

int cmtjb_fp64_peak(int device, int64_t iters, double *flops, double *ms) {
  CUDA_TRY(cudaSetDevice(device));
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, device));
  const int blocks = prop.multiProcessorCount * 4, threads = 256;
  double *d = nullptr;
  CUDA_TRY(cudaMalloc(&d, 8));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  fp64_peak_kernel<<<blocks, threads>>>(d, iters / 8 + 1, 0.999999, 1e-7);  // warm-up
  cudaEventRecord(e0);
  fp64_peak_kernel<<<blocks, threads>>>(d, iters, 0.999999, 1e-7);
  cudaEventRecord(e1);
  cudaError_t e = cudaEventSynchronize(e1);
  float t = 0;
  cudaEventElapsedTime(&t, e0, e1);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d);
  if (e != cudaSuccess) return cuda_fail(e, "fp64_peak_kernel");
  const double n_fma = (double)blocks * threads * (double)iters * 64.0;
  if (flops) *flops = 2.0 * n_fma / (t * 1e-3);
  if (ms) *ms = t;
  return CMTJB_OK;
}

} /* extern "C" */
