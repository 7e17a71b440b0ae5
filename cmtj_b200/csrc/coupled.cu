// coupled.cu -- extern "C" entry points for Stack and GroupInteraction batches
// (include/cmtj_b200.h, "coupled systems"); kernels in coupled_kernels.cuh.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "coupled_kernels.cuh"
#include "host_common.h"
#include "onef.h"

namespace cmtjb {
namespace {

struct Common {
  int nl = 0;
  int solver = CMTJB_RK4;
  int overridden = 0;
  int n_jch = 0;
  cmtjb_plan_t plan0;
};

// Shared validation: same layer count everywhere, per-junction runSimulation guards, one solver
// mode for the whole system (stack.hpp:36-64, reservoir.hpp:196-213).
int validate_system(int n, const cmtjb_junction_t *jns, const cmtjb_layer_t *layers, const cmtjb_run_args_t *a, Common *c) {
  if (!jns || !layers || !a) return fail(CMTJB_EINVAL, "null argument");
  if (a->timeStep > a->writeFrequency) return fail(CMTJB_ERUNTIME, "The time step cannot be larger than write frequency!");
  c->nl = jns[0].n_layers;
  cmtjb_run_args_t aj = *a;
  aj.solverMode = CMTJB_RK4;
  aj.onef_buffer = nullptr;
  for (int j = 0; j < n; j++) {
    if (jns[j].n_layers != c->nl)
      return fail(CMTJB_ENOTSUP, "junctions of a coupled system must all have the same number of layers in this build");
    cmtjb_plan_t p;
    int rc = make_plan(&jns[j], layers + (size_t)j * c->nl, &aj, &p);
    if (rc) return rc;
    for (int l = 0; l < c->nl; l++) {
      for (int s = 0; s < CMTJB_NSLOTS; s++)
        if (layers[(size_t)j * c->nl + l].drivers[s].update == CMTJB_UPD_CUSTOM)
          return fail(CMTJB_ENOTSUP, "custom (Python callback) drivers are not supported inside a coupled system");
    }
    if (j == 0) {
      c->plan0 = p;
      c->solver = p.effective_solver;
      c->overridden = p.solver_overridden;
    } else if (p.effective_solver != c->solver) {
      return fail(CMTJB_ERUNTIME,
                  "Junctions have different solver modes! Set the same solver mode for all junctions explicitly. Do not "
                  "mix stochastic and deterministic solvers!");
    }
  }
  c->n_jch = c->plan0.n_channels;
  return CMTJB_OK;
}

void fill_plan(const Common &c, int n_sys_ch, cmtjb_coupled_plan_t *p) {
  std::memset(p, 0, sizeof(*p));
  p->effective_solver = c.solver;
  p->solver_overridden = c.overridden;
  p->n_layers = c.nl;
  p->n_junction_channels = c.n_jch;
  p->n_system_channels = n_sys_ch;
  p->total_iterations = c.plan0.total_iterations;
  p->write_every = c.plan0.write_every;
  p->n_samples = c.plan0.n_samples;
}

// Union of the term sets the junctions of a system need (cabi.cu: needed_features).  The coupled
// kernels have no per-launch driver table, so any time-dependent driver means inline evaluation;
// system-level overrides (the stack's current driver) are not junction terms.
// `stack`: the layers' own current drivers are never evaluated -- the stack pushes its coupled current into
// every layer each step (stack.hpp:79-91) -- so they do not count as time-dependent terms.
uint32_t system_features(int n, int nl, const cmtjb_junction_t *jns, const cmtjb_layer_t *layers, const cmtjb_override_t *ovr,
                         int n_ovr, const cmtjb_run_args_t *a, int solver, bool stack) {
  std::vector<cmtjb_override_t> jo;
  for (int k = 0; k < n_ovr; k++)
    if (ovr[k].layer != CMTJB_LAYER_SYSTEM) jo.push_back(ovr[k]);
  cmtjb_run_args_t aj = *a;
  aj.onef_buffer = nullptr;
  uint32_t F = 0;
  std::vector<cmtjb_layer_t> tmp(nl);
  for (int j = 0; j < n; j++) {
    for (int l = 0; l < nl; l++) {
      tmp[l] = layers[(size_t)j * nl + l];
      if (stack) {
        std::memset(&tmp[l].drivers[CMTJB_SLOT_CURRENT], 0, sizeof(cmtjb_driver_t));
        tmp[l].drivers[CMTJB_SLOT_CURRENT].update = CMTJB_UPD_CONSTANT;
      }
    }
    F |= model_features(&jns[j], tmp.data(), jo.data(), (int)jo.size(), &aj, solver);
  }
  if (F & (F_DYN | F_DYNS)) F = (F & ~(F_DYN | F_DYNS)) | F_DYNI;
  const char *force = std::getenv("CMTJB_FORCE_VARIANT");  // "full": always the most general kernel (testing)
  if (force && std::strcmp(force, "full") == 0) F = F_ALL;
  return F;
}
inline bool lean_ok(uint32_t need) { return (need & ~F_COUPLED_LEAN) == 0; }
// every junction carries the same derived model (bytewise): pass it once, by value
template <int NL> bool models_uniform(const std::vector<DevModel<NL>> &m) {
  for (size_t j = 1; j < m.size(); j++)
    if (std::memcmp(&m[0], &m[j], sizeof(DevModel<NL>)) != 0) return false;
  return true;
}

// 1/f^alpha noise of a coupled system (every junction's layers tick their own VectorAlphaNoise inside the
// junction's stepper, junction.hpp:1725-1727): one series per (element, layer, axis).  The buffer has one slot
// layout for the whole system, defined per layer index by the first junction that carries noise there; the
// others must use the same axis, alpha and std (scale may differ, 0 = none).
struct SystemNoise {
  OnefSlot slots[3 * CMTJB_MAX_LAYERS];
  int first[CMTJB_MAX_LAYERS];
  int n_slots = 0;
};
int system_noise(int n, int nl, const cmtjb_layer_t *layers, int solver, SystemNoise *sn) {
  sn->n_slots = 0;
  if (solver != CMTJB_EULER_HEUN && solver != CMTJB_HEUN) return CMTJB_OK;
  std::vector<cmtjb_layer_t> proto(nl);
  for (int l = 0; l < nl; l++) {
    proto[l] = layers[l];
    proto[l].noiseScale = 0;
    for (int j = 0; j < n; j++) {
      const cmtjb_layer_t &L = layers[(size_t)j * nl + l];
      if (L.noiseScale == 0) continue;
      if (proto[l].noiseScale == 0) {
        proto[l] = L;
      } else if (L.noiseAxis != proto[l].noiseAxis || L.noiseAlpha != proto[l].noiseAlpha || L.noiseStd != proto[l].noiseStd) {
        return fail(CMTJB_ENOTSUP, "1/f^alpha noise in a coupled system: junctions must share axis, alpha and std per layer in this build");
      }
    }
  }
  const int ns = onef_slots_of(proto.data(), nl, sn->slots, sn->first);
  if (ns < 0) return fail(CMTJB_ERUNTIME, "Invalid axis specified");
  sn->n_slots = ns;
  return CMTJB_OK;
}

// generate the system's noise buffers for all n_elements elements (stream-ordered); *gen is freed by the caller
int system_noise_generate(const SystemNoise &sn, const cmtjb_run_args_t *a, long long n_iter, long long n_elements, int n_junctions,
                          RunParams &R, float **gen, cudaStream_t stream) {
  *gen = nullptr;
  if (sn.n_slots == 0 || n_iter <= 0) return CMTJB_OK;
  const size_t bytes = (size_t)sn.n_slots * (size_t)n_iter * sizeof(float) * (size_t)n_elements;
  size_t free_b = 0, total_b = 0;
  if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess && bytes > free_b / 2) {
    char buf[256];
    std::snprintf(buf, sizeof(buf), "1/f^alpha noise: %zu bytes of noise buffers (%d series x %lld iterations x %lld elements x 4 B) exceed half of the %zu bytes free on the device; run fewer instances per call",
                  bytes, sn.n_slots, n_iter, n_elements, free_b);
    return fail(CMTJB_ERUNTIME, buf);
  }
  cudaError_t e = cudaMallocAsync((void **)gen, bytes, stream);
  if (e != cudaSuccess) return cuda_fail(e, "cudaMallocAsync(1/f noise buffers)");
  int dev = 0;
  cudaGetDevice(&dev);
  int rc = onef_generate(dev, stream, n_iter, sn.slots, sn.n_slots, a->seed, (unsigned long long)a->instance_offset * (unsigned long long)n_junctions,
                         n_elements, *gen);
  if (rc) {
    cudaFreeAsync(*gen, stream);
    *gen = nullptr;
    return rc;
  }
  R.onef_gen = *gen;
  R.onef_i0 = 0;
  R.onef_B = n_elements;
  return CMTJB_OK;
}

template <int NL>
void build_models(int n, const cmtjb_junction_t *jns, const cmtjb_layer_t *layers, const cmtjb_override_t *ovr, int n_ovr,
                  int clear_hres_layer, double dt, std::vector<DevModel<NL>> &out, RunParams &R, const SystemNoise *sn = nullptr) {
  out.resize(n);
  for (int j = 0; j < n; j++) {
    DevModel<NL> &M = out[j];
    std::memset(&M, 0, sizeof(M));
    for (int l = 0; l < NL; l++) {
      cmtjb_layer_t L = layers[(size_t)j * NL + l];
      if (l == clear_hres_layer) {
        // GroupInteraction replaces the reserved-field driver of the top layer every step
        for (int s = CMTJB_SLOT_HRES_X; s <= CMTJB_SLOT_HRES_Z; s++) {
          std::memset(&L.drivers[s], 0, sizeof(L.drivers[s]));
          L.drivers[s].update = CMTJB_UPD_CONSTANT;
        }
      }
      derive_layer(jns[j], L, M.L[l]);
      set_step_products(M.L[l], dt);
      if (sn && sn->n_slots > 0) {
        M.L[l].onef_slot = sn->first[l] < 0 ? 0 : sn->first[l];
        M.L[l].onef_axis = L.noiseAxis;
        M.L[l].onef_scale = L.noiseScale;
      }
    }
    M.Rp = jns[j].Rp;
    M.Rhalf = (jns[j].Rap - jns[j].Rp) / 2.0;
    M.mr_mode = jns[j].mr_mode;
    M.n_layers = NL;
    int k_row = 0;
    for (int k = 0; k < n_ovr; k++) {
      if (ovr[k].layer == CMTJB_LAYER_SYSTEM) continue;
      for (int l = 0; l < NL; l++) {
        if (ovr[k].layer >= 0 && ovr[k].layer != l) continue;
        M.L[l].ovmask |= (1u << ovr[k].slot);
        M.L[l].ovidx[ovr[k].slot][ovr[k].field] = (int8_t)k_row;
      }
      k_row++;
    }
  }
  int k_row = 0;
  for (int k = 0; k < n_ovr; k++) {
    if (ovr[k].layer == CMTJB_LAYER_SYSTEM) continue;
    R.ovr[k_row++] = ovr[k].values;
  }
}

int check_coupled_overrides(int nl, const cmtjb_override_t *ovr, int n_ovr) {
  if (n_ovr < 0 || n_ovr > kMaxOverrides) return fail(CMTJB_EINVAL, "too many per-instance overrides (max 24)");
  if (n_ovr > 0 && !ovr) return fail(CMTJB_EINVAL, "null override table");
  for (int k = 0; k < n_ovr; k++) {
    if (ovr[k].field < 0 || ovr[k].field > CMTJB_FIELD_STEADYTIME || !ovr[k].values) return fail(CMTJB_EINVAL, "override out of range");
    if (ovr[k].layer == CMTJB_LAYER_SYSTEM) continue;
    if (ovr[k].layer >= nl || ovr[k].layer < -1 || ovr[k].slot < 0 || ovr[k].slot >= CMTJB_NSLOTS)
      return fail(CMTJB_EINVAL, "override out of range");
  }
  return CMTJB_OK;
}

void fill_run_params(RunParams &R, const Common &c, const cmtjb_run_args_t *a, long long n_elements, double *m_state, double *jlog,
                     int32_t *status) {
  R.m_state = m_state;
  R.log = (a->log_mode == CMTJB_LOG_FULL) ? jlog : nullptr;
  R.status = status;
  R.B = n_elements;
  R.i_lo = 0;
  R.i_hi = n_elements;
  R.inst_offset = a->instance_offset;
  R.n_samples = c.plan0.n_samples;
  R.it_begin = a->start_iteration;
  R.it_end = c.plan0.total_iterations;
  R.total_iterations = c.plan0.total_iterations;
  R.write_every = c.plan0.write_every > 0 ? c.plan0.write_every : 1;
  R.first_sample = 0;
  R.elog_ch0 = -1;
  R.dt = a->timeStep;
  R.sqrt_dt = std::sqrt(a->timeStep);
  R.total_time = a->totalTime;
  R.write_frequency = a->writeFrequency;
  R.seed = a->seed;
  R.pk = make_philox_keys(a->seed);
  R.normals64 = a->thermal_normals == 1 ? 1 : 0;
}

// stream-ordered upload of a small host array
template <class T> cudaError_t upload(const T *src, size_t n, T **dst, cudaStream_t st) {
  keep_pool_memory();
  cudaError_t e = cudaMallocAsync((void **)dst, std::max<size_t>(n, 1) * sizeof(T), st);
  if (e != cudaSuccess) return e;
  if (n) e = cudaMemcpyAsync(*dst, src, n * sizeof(T), cudaMemcpyHostToDevice, st);
  return e;
}

// ---------------------------------------------------------------------------------------------------
// Stack
// ---------------------------------------------------------------------------------------------------
int validate_stack(const cmtjb_stack_t *st, const cmtjb_junction_t *jns, const cmtjb_layer_t *layers, const cmtjb_run_args_t *a,
                   Common *c) {
  if (!st) return fail(CMTJB_EINVAL, "null argument");
  const int n = st->n_junctions;
  if (n < 2) return fail(CMTJB_ERUNTIME, "Stack must have at least 2 junctions!");
  for (int j = 0; j < n; j++)
    if (jns[j].mr_mode != CMTJB_MR_CLASSIC)
      return fail(CMTJB_ERUNTIME,
                  "Junction has a non-classic magnetoresitance mode! Define the junction with Rp and Rap resistance values.");
  for (int j = 0; j < n; j++)
    if (jns[j].n_layers >= 3) return fail(CMTJB_ERUNTIME, "At least one junction has more than 2 layers! It's not supported now.");
  int rc = validate_system(n, jns, layers, a, c);
  if (rc) return rc;
  if (!st->coupling) return fail(CMTJB_ERUNTIME, "Coupling strength is not set!");
  if (st->top_layer < 0 || st->top_layer >= c->nl || st->bottom_layer >= c->nl) return fail(CMTJB_EINVAL, "layer index out of range");
  if (c->nl == 2 && st->bottom_layer < 0) return fail(CMTJB_EINVAL, "two-layer stack needs the bottom layer index");
  if (!st->delayed && c->nl != 2) return fail(CMTJB_ERUNTIME, "Non delayed coupling is only supported for 2 layer stacks!");
  if (c->nl == 1) {
    // getPolarisationVector (stack.hpp:285-296)
    const double *p0 = layers[st->top_layer].referenceLayer;
    for (int j = 1; j < n; j++) {
      const double *pj = layers[(size_t)j * c->nl + st->top_layer].referenceLayer;
      if (pj[0] != p0[0] || pj[1] != p0[1] || pj[2] != p0[2]) return fail(CMTJB_ERUNTIME, "Polarisation vectors are not equal in stack");
    }
    if (std::sqrt(p0[0] * p0[0] + p0[1] * p0[1] + p0[2] * p0[2]) == 0) return fail(CMTJB_ERUNTIME, "Polarisation is not set!");
  }
  if (st->currentDriver.update == CMTJB_UPD_CUSTOM) return fail(CMTJB_ENOTSUP, "custom (Python callback) drivers are not supported on the device path");
  if (n > kCoupledBlockMax) return fail(CMTJB_ENOTSUP, "stacks of more than 512 junctions are not supported by this build");
  return CMTJB_OK;
}

template <int NL>
int stack_launch(const cmtjb_stack_t *st, const cmtjb_junction_t *jns, const cmtjb_layer_t *layers, const cmtjb_override_t *ovr,
                 int n_ovr, const cmtjb_run_args_t *a, const Common &c, double *m_state, double *jlog, double *slog,
                 int32_t *status, cudaStream_t stream) {
  const int N = st->n_junctions;
  const long long n_inst = a->n_instances;
  if (n_inst <= 0) return CMTJB_OK;
  RunParams R;
  std::memset(&R, 0, sizeof(R));
  fill_run_params(R, c, a, n_inst * N, m_state, jlog, status);
  SystemNoise sn;
  int rc_n = system_noise(N, NL, layers, c.solver, &sn);
  if (rc_n) return rc_n;
  std::vector<DevModel<NL>> models;
  build_models<NL>(N, jns, layers, ovr, n_ovr, -1, a->timeStep, models, R, &sn);
  StackParams SP;
  std::memset(&SP, 0, sizeof(SP));
  SP.N = N;
  // stacks per block: the per-step barriers are block-wide, so a block holds as few stacks as fill two warps
  // (CMTJB_STACK_BLOCK overrides the thread target; tuning knob)
  int block_target = 64;
  if (const char *c = std::getenv("CMTJB_STACK_BLOCK")) block_target = std::max(32, std::atoi(c));
  SP.spb = std::max(1, block_target / N);
  SP.kind = st->kind;
  SP.use_kcl = st->use_kcl;
  SP.delayed = st->delayed;
  SP.top = st->top_layer;
  SP.bottom = (NL == 2) ? st->bottom_layer : -1;
  SP.n_inst = n_inst;
  SP.phaseOffset = st->phaseOffset;
  if (NL == 1)
    for (int k = 0; k < 3; k++) SP.pol[k] = layers[st->top_layer].referenceLayer[k];
  {
    cmtjb_junction_t J0 = jns[0];
    cmtjb_layer_t L;
    std::memset(&L, 0, sizeof(L));
    L.Ms = 1;
    L.thickness = 1;
    for (int s = 0; s < CMTJB_NSLOTS; s++) L.drivers[s].update = CMTJB_UPD_CONSTANT;
    L.drivers[CMTJB_SLOT_CURRENT] = st->currentDriver;
    static thread_local DevLayer D;
    derive_layer(J0, L, D);
    SP.cur = D.drv[CMTJB_SLOT_CURRENT];
    SP.cur_cst = D.cst[CMTJB_SLOT_CURRENT];
    if (st->currentDriver.update == CMTJB_UPD_CONSTANT) SP.cur.amplitude = 0.0;
  }
  for (int k = 0; k < n_ovr; k++)
    if (ovr[k].layer == CMTJB_LAYER_SYSTEM) SP.cur_ovr[ovr[k].field] = ovr[k].values;
  SP.slog = (a->log_mode == CMTJB_LOG_FULL) ? slog : nullptr;
  DevModel<NL> *d_models = nullptr;
  double *d_coupling = nullptr;
  cudaError_t e = upload(models.data(), models.size(), &d_models, stream);
  if (e == cudaSuccess) e = upload(st->coupling, (size_t)(N - 1), &d_coupling, stream);
  if (e != cudaSuccess) return cuda_fail(e, "stack parameter upload");
  SP.coupling = d_coupling;
  float *d_gen = nullptr;
  rc_n = system_noise_generate(sn, a, c.plan0.total_iterations, n_inst * N, N, R, &d_gen, stream);
  if (rc_n) {
    cudaFreeAsync(d_models, stream);
    cudaFreeAsync(d_coupling, stream);
    return rc_n;
  }
  const int threads = SP.spb * N;
  const long long blocks = (n_inst + SP.spb - 1) / SP.spb;
  const size_t smem = (3 * (size_t)threads + 2 * ((size_t)threads / 32 + 1)) * sizeof(double);
  const bool lean = lean_ok(system_features(N, NL, jns, layers, ovr, n_ovr, a, c.solver, true));
  const bool uni = lean && models_uniform<NL>(models);
  const DevModel<NL> &M0 = models[0];
#define CMTJB_STACK_LAUNCH2(SOLVER, MODE)                                                                                                \
  do {                                                                                                                                   \
    if (uni) stack_kernel<NL, SOLVER, F_COUPLED_LEAN, true, MODE><<<(unsigned)blocks, threads, smem, stream>>>(d_models, M0, SP, R);       \
    else if (lean) stack_kernel<NL, SOLVER, F_COUPLED_LEAN, false, MODE><<<(unsigned)blocks, threads, smem, stream>>>(d_models, M0, SP, R); \
    else stack_kernel<NL, SOLVER, F_ALL, false, MODE><<<(unsigned)blocks, threads, smem, stream>>>(d_models, M0, SP, R);                   \
  } while (0)
#define CMTJB_STACK_LAUNCH(SOLVER)                                    \
  do {                                                                \
    if (SP.use_kcl) CMTJB_STACK_LAUNCH2(SOLVER, STACK_KCL);           \
    else if (SP.delayed) CMTJB_STACK_LAUNCH2(SOLVER, STACK_DELAYED);  \
    else CMTJB_STACK_LAUNCH2(SOLVER, STACK_SEQUENTIAL);               \
  } while (0)
  if (c.solver == CMTJB_RK4) CMTJB_STACK_LAUNCH(CMTJB_RK4);
  else CMTJB_STACK_LAUNCH(CMTJB_EULER_HEUN);
#undef CMTJB_STACK_LAUNCH2
#undef CMTJB_STACK_LAUNCH
  e = cudaGetLastError();
  g_launches++;
  cudaFreeAsync(d_models, stream);
  cudaFreeAsync(d_coupling, stream);
  if (d_gen) cudaFreeAsync(d_gen, stream);
  if (e != cudaSuccess) return cuda_fail(e, "stack_kernel");
  return CMTJB_OK;
}

int stack_run(const cmtjb_stack_t *st, const cmtjb_junction_t *jns, const cmtjb_layer_t *layers, const cmtjb_override_t *ovr,
              int n_ovr, const cmtjb_run_args_t *a, double *m_state, double *jlog, double *slog, int32_t *status,
              cudaStream_t stream) {
  Common c;
  int rc = validate_stack(st, jns, layers, a, &c);
  if (rc) return rc;
  rc = check_coupled_overrides(c.nl, ovr, n_ovr);
  if (rc) return rc;
  if (!m_state) return fail(CMTJB_EINVAL, "m_state is null");
  if (c.nl == 1) return stack_launch<1>(st, jns, layers, ovr, n_ovr, a, c, m_state, jlog, slog, status, stream);
  return stack_launch<2>(st, jns, layers, ovr, n_ovr, a, c, m_state, jlog, slog, status, stream);
}

// ---------------------------------------------------------------------------------------------------
// GroupInteraction
// ---------------------------------------------------------------------------------------------------
int validate_group(const cmtjb_group_t *g, const cmtjb_junction_t *jns, const cmtjb_layer_t *layers, const cmtjb_run_args_t *a,
                   Common *c) {
  if (!g) return fail(CMTJB_EINVAL, "null argument");
  const int n = g->n_junctions;
  if (n <= 0 || !g->coordinates) return fail(CMTJB_ERUNTIME, "Coordinate matrix and junction list cannot be empty!");
  int rc = validate_system(n, jns, layers, a, c);
  if (rc) return rc;
  if (c->nl > 4) return fail(CMTJB_ENOTSUP, "coupled junctions with more than 4 layers are not supported by this build");
  if (g->top_layer < 0 || g->top_layer >= c->nl) return fail(CMTJB_EINVAL, "layer index out of range");
  if (g->interaction < CMTJB_INT_NULL || g->interaction > CMTJB_INT_TABLE) return fail(CMTJB_EINVAL, "unknown interaction function");
  if (g->interaction == CMTJB_INT_TABLE && !g->pair_table) return fail(CMTJB_EINVAL, "CMTJB_INT_TABLE without a pair table");
  // reservoir.hpp:171-177 (unique coordinates) and the per-pair distance guards (:84-87, :39-42): the reference tests
  // every pair; points closer than `lim` are adjacent in x within `lim`, so a sort and a sliding window find them
  const double lim = (g->interaction == CMTJB_INT_DIPOLE_NOMURA) ? 1e-14 : 1e-10;
  const double *r = g->coordinates;
  std::vector<int> order(n);
  for (int i = 0; i < n; i++) order[i] = i;
  std::sort(order.begin(), order.end(), [&](int a, int b) {
    if (r[3 * a] != r[3 * b]) return r[3 * a] < r[3 * b];
    if (r[3 * a + 1] != r[3 * b + 1]) return r[3 * a + 1] < r[3 * b + 1];
    return r[3 * a + 2] < r[3 * b + 2];
  });
  bool too_close = false;
  for (int u = 0; u < n; u++)
    for (int v = u + 1; v < n; v++) {
      const int i = order[u], j = order[v];
      const double dx = r[3 * j] - r[3 * i], dy = r[3 * i + 1] - r[3 * j + 1], dz = r[3 * i + 2] - r[3 * j + 2];
      if (dx > lim && dx != 0) break;
      if (dx == 0 && dy == 0 && dz == 0) return fail(CMTJB_ERUNTIME, "Coordinate vectors must be unique!");
      if ((g->interaction == CMTJB_INT_DIPOLE || g->interaction == CMTJB_INT_DIPOLE_NOMURA) && std::sqrt(dx * dx + dy * dy + dz * dz) < lim)
        too_close = true;
    }
  if (too_close) return fail(CMTJB_ERUNTIME, "Points are too close for stable dipole calculation");
  return CMTJB_OK;
}

int pow2_floor(long long v) {
  int p = 1;
  while ((long long)p * 2 <= v) p *= 2;
  return p;
}

template <int NL> bool group_uses_grid(int N) { return N * 1 > kCoupledBlockMax / 2; }

template <int NL>
int group_launch(const cmtjb_group_t *g, const cmtjb_junction_t *jns, const cmtjb_layer_t *layers, const cmtjb_override_t *ovr,
                 int n_ovr, const cmtjb_run_args_t *a, const Common &c, double *m_state, double *jlog, int32_t *status,
                 int device, cudaStream_t stream) {
  const int N = g->n_junctions;
  const long long n_inst = a->n_instances;
  if (n_inst <= 0) return CMTJB_OK;
  RunParams R;
  std::memset(&R, 0, sizeof(R));
  fill_run_params(R, c, a, n_inst * N, m_state, jlog, status);
  SystemNoise sn;
  int rc_n = system_noise(N, NL, layers, c.solver, &sn);
  if (rc_n) return rc_n;
  std::vector<DevModel<NL>> models;
  build_models<NL>(N, jns, layers, ovr, n_ovr, g->top_layer, a->timeStep, models, R, &sn);
  std::vector<double> pos(3 * (size_t)N), pref(N);
  for (int k = 0; k < N; k++) {
    for (int d = 0; d < 3; d++) pos[(size_t)d * N + k] = g->coordinates[3 * k + d];
    const cmtjb_layer_t &L = layers[(size_t)k * NL + g->top_layer];
    pref[k] = (L.Ms / jns[k].MAGNETIC_PERMEABILITY) * (L.thickness * L.cellSurface);
  }
  GroupParams GP;
  std::memset(&GP, 0, sizeof(GP));
  GP.N = N;
  GP.interaction = g->interaction;
  GP.top = g->top_layer;
  GP.n_inst = n_inst;
  DevModel<NL> *d_models = nullptr;
  double *d_pos = nullptr, *d_pref = nullptr, *d_mbuf = nullptr;
  cudaError_t e = upload(models.data(), models.size(), &d_models, stream);
  if (e == cudaSuccess) e = upload(pos.data(), pos.size(), &d_pos, stream);
  if (e == cudaSuccess) e = upload(pref.data(), pref.size(), &d_pref, stream);
  if (e != cudaSuccess) return cuda_fail(e, "group parameter upload");
  GP.pos = d_pos;
  GP.pref = d_pref;
  float *d_gen = nullptr;
  int rc = system_noise_generate(sn, a, c.plan0.total_iterations, n_inst * N, N, R, &d_gen, stream);
  if (rc) {
    cudaFreeAsync(d_models, stream);
    cudaFreeAsync(d_pos, stream);
    cudaFreeAsync(d_pref, stream);
    return rc;
  }
  const bool rk4 = (c.solver == CMTJB_RK4);
  const bool lean = lean_ok(system_features(N, NL, jns, layers, ovr, n_ovr, a, c.solver, false));
  const bool uni = lean && models_uniform<NL>(models);
  DevModel<NL> M0 = models[0];
  const bool table = g->interaction == CMTJB_INT_TABLE;  // tabulated interaction: the two-kernel form for every N
  if (!table && !group_uses_grid<NL>(N)) {
    GP.P = std::min(32, pow2_floor(kCoupledBlockMax / N));
    const int threads = ((N * GP.P + 31) / 32) * 32;
    const size_t smem = 13 * (size_t)N * sizeof(double);
#define CMTJB_GROUP_LAUNCH(SOLVER)                                                                                              \
  do {                                                                                                                          \
    if (uni) group_block_kernel<NL, SOLVER, F_COUPLED_LEAN, true><<<(unsigned)n_inst, threads, smem, stream>>>(d_models, M0, GP, R); \
    else if (lean) group_block_kernel<NL, SOLVER, F_COUPLED_LEAN, false><<<(unsigned)n_inst, threads, smem, stream>>>(d_models, M0, GP, R); \
    else group_block_kernel<NL, SOLVER, F_ALL, false><<<(unsigned)n_inst, threads, smem, stream>>>(d_models, M0, GP, R);         \
  } while (0)
    if (rk4) CMTJB_GROUP_LAUNCH(CMTJB_RK4);
    else CMTJB_GROUP_LAUNCH(CMTJB_EULER_HEUN);
#undef CMTJB_GROUP_LAUNCH
    e = cudaGetLastError();
    g_launches++;
    if (e != cudaSuccess) rc = cuda_fail(e, "group_block_kernel");
  } else {
    const char *two_env = std::getenv("CMTJB_GROUP_TWO_KERNEL");
    if (table || (two_env && two_env[0] == '1')) {
      // EXPERIMENT (off by default, measured slower: 79 vs 61 us per step at N = 4096): two kernels per step,
      // the pair sum at full occupancy, then the per-element step (coupled_kernels.cuh)
      int sms = 148;
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
      const int tiles = (N + 31) / 32;
      const long long want_threads = (long long)sms * 2048;  // every SM full
      const int P = (int)std::max<long long>(1, std::min<long long>(std::min<long long>(64, N), want_threads / ((long long)tiles * 32)));
      double *d_hpart = nullptr, *d_table = nullptr;
      RunParams *d_R = nullptr;
      e = cudaMallocAsync((void **)&d_hpart, (size_t)P * 3 * N * sizeof(double), stream);
      if (e == cudaSuccess) e = upload(&R, 1, &d_R, stream);
      if (e == cudaSuccess && table) e = upload(g->pair_table, (size_t)N * 12 * N, &d_table, stream);
      if (e != cudaSuccess) rc = cuda_fail(e, "cudaMallocAsync(pair-field partials)");
      const long long we = R.write_every;
      const long long Btot = R.B;
      for (long long inst = 0; inst < n_inst && rc == CMTJB_OK; inst++) {
        PairParams Q;
        Q.N = N;
        Q.P = P;
        Q.pos = d_pos;
        Q.pref = d_pref;
        Q.mx = m_state + (size_t)(3 * g->top_layer + 0) * Btot + inst * N;
        Q.my = m_state + (size_t)(3 * g->top_layer + 1) * Btot + inst * N;
        Q.mz = m_state + (size_t)(3 * g->top_layer + 2) * Btot + inst * N;
        Q.Hpart = d_hpart;
        GroupStepParams G;
        std::memset(&G, 0, sizeof(G));
        G.N = N;
        G.P = P;
        G.interaction = g->interaction;
        G.top = g->top_layer;
        G.inst = inst;
        G.Hpart = d_hpart;
        long long sample = R.first_sample;
        const long long pair_blocks = ((long long)tiles * 32 * P + 255) / 256;
        for (long long it = R.it_begin; it < R.it_end && rc == CMTJB_OK; it++) {
          if (table) {
            PairTableParams QT;
            QT.N = N, QT.P = P, QT.table = d_table, QT.mx = Q.mx, QT.my = Q.my, QT.mz = Q.mz, QT.Hpart = d_hpart;
            group_pair_table_kernel<<<(unsigned)pair_blocks, 256, 0, stream>>>(QT);
            g_launches++;
          } else if (g->interaction != CMTJB_INT_NULL) {
            group_pair_kernel<<<(unsigned)pair_blocks, 256, 0, stream>>>(Q);
            g_launches++;
          }
          G.it = it;
          G.write = (it % we == 0) ? 1 : 0;
          G.sample = sample;
          if (G.write) sample++;
          const unsigned sb = (unsigned)((N + 127) / 128);
#define CMTJB_GSTEP(SOLVER)                                                                                              \
  do {                                                                                                                   \
    if (uni) group_step_kernel<NL, SOLVER, F_COUPLED_LEAN, true><<<sb, 128, 0, stream>>>(d_models, d_R, G);              \
    else if (lean) group_step_kernel<NL, SOLVER, F_COUPLED_LEAN, false><<<sb, 128, 0, stream>>>(d_models, d_R, G);       \
    else group_step_kernel<NL, SOLVER, F_ALL, false><<<sb, 128, 0, stream>>>(d_models, d_R, G);                          \
  } while (0)
          if (rk4) CMTJB_GSTEP(CMTJB_RK4);
          else CMTJB_GSTEP(CMTJB_EULER_HEUN);
#undef CMTJB_GSTEP
          g_launches++;
          e = cudaGetLastError();
          if (e != cudaSuccess) rc = cuda_fail(e, "group pair/step kernels");
        }
      }
      if (d_hpart) cudaFreeAsync(d_hpart, stream);
      if (d_R) cudaFreeAsync(d_R, stream);
      if (d_table) cudaFreeAsync(d_table, stream);
    } else {
    // one array across the grid: cooperative launch, one block per SM, every block the same share of the N^2 pairs
    // lean single-layer arrays (the reservoir case) run 768 threads per block; CMTJB_GROUP_THREADS=512 forces the 512 form
    const char *thr_env = std::getenv("CMTJB_GROUP_THREADS");
    const bool wide_block = NL == 1 && uni && !(thr_env && std::atoi(thr_env) == 512);
    const int n_threads = wide_block ? kGroupThreadsLean : kGroupThreads;
#define CMTJB_GRID_FN(SOLVER)                                                                                          \
  (wide_block ? (const void *)group_grid_kernel<NL, SOLVER, F_COUPLED_LEAN, true, kGroupThreadsLean>                   \
   : uni      ? (const void *)group_grid_kernel<NL, SOLVER, F_COUPLED_LEAN, true, kGroupThreads>                       \
   : lean     ? (const void *)group_grid_kernel<NL, SOLVER, F_COUPLED_LEAN, false, kGroupThreads>                      \
              : (const void *)group_grid_kernel<NL, SOLVER, F_ALL, false, kGroupThreads>)
    const void *fn = rk4 ? CMTJB_GRID_FN(CMTJB_RK4) : CMTJB_GRID_FN(CMTJB_EULER_HEUN);
#undef CMTJB_GRID_FN
    int per_sm = 0, sms = 0, coop = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, device);
    // elements per block: at least 8 (tiny arrays use fewer blocks: the barrier, not the pair sum, is their cost)
    int n_blocks = std::max(1, std::min(sms, (N + 7) / 8));
    GP.A = (N + n_blocks - 1) / n_blocks;
    n_blocks = (N + GP.A - 1) / GP.A;
    GP.P = GP.A <= n_threads ? n_threads / GP.A : 0;
    GP.tile = kGroupTileMax;
    const size_t smem = ((size_t)7 * GP.tile + 3 * (size_t)GP.A) * sizeof(double);
    if (coop && GP.P > 0) e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess && coop && GP.P > 0) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, n_threads, smem);
    if (!coop || e != cudaSuccess || (GP.P > 0 && per_sm <= 0)) {
      cudaGetLastError();
      rc = fail(CMTJB_ECUDA, "cooperative launch is not available for the group kernel on this device");
    } else if (GP.P == 0 || n_blocks > per_sm * sms) {
      rc = fail(CMTJB_ENOTSUP, "dipole-coupled array exceeds one GPU's co-resident thread capacity (" + std::to_string((long long)sms * n_threads) +
                                   " elements); shard it over several GPUs (cmtj_b200.reservoir.ShardedGroup)");
    } else {
      e = cudaMallocAsync((void **)&d_mbuf, 6 * (size_t)N * sizeof(double), stream);
      if (e != cudaSuccess) rc = cuda_fail(e, "cudaMallocAsync(group exchange buffer)");
      GP.mbuf = d_mbuf;
      for (long long inst = 0; inst < n_inst && rc == CMTJB_OK; inst++) {
        GP.inst = inst;
        void *args[] = {(void *)&d_models, (void *)&M0, (void *)&GP, (void *)&R};
        e = cudaLaunchCooperativeKernel(fn, dim3((unsigned)n_blocks), dim3(n_threads), args, smem, stream);
        g_launches++;
        if (e != cudaSuccess) rc = cuda_fail(e, "group_grid_kernel");
      }
    }
    }
  }
  cudaFreeAsync(d_models, stream);
  cudaFreeAsync(d_pos, stream);
  cudaFreeAsync(d_pref, stream);
  if (d_mbuf) cudaFreeAsync(d_mbuf, stream);
  if (d_gen) cudaFreeAsync(d_gen, stream);
  return rc;
}

// One array sharded over `world` ranks (one process per GPU): this rank integrates elements [e_lo, e_hi) and exchanges
// the top-layer magnetisation of all elements once per step through `exchange` (an in-place all-gather of the
// [world][3][per] device buffer, stream-ordered; NCCL in cmtj_b200.reservoir.ShardedGroup).  Device pointers, sized
// for this rank's elements only.
template <int NL>
int group_shard_launch(const cmtjb_group_t *g, const cmtjb_junction_t *jns, const cmtjb_layer_t *layers, const cmtjb_run_args_t *a,
                       const Common &c, int rank, int world, cmtjb_exchange_fn exchange, void *ctx, void *const *areas, double *m_state,
                       double *jlog, int32_t *status, int device, cudaStream_t stream) {
  const int N = g->n_junctions;
  const int per = (N + world - 1) / world;
  const bool peer = areas != nullptr;  // peer-memory exchange (coupled_kernels.cuh: PeerPush) instead of the callback
  const int e_lo = std::min(N, rank * per), e_hi = std::min(N, e_lo + per), own = e_hi - e_lo;
  RunParams R;
  std::memset(&R, 0, sizeof(R));
  fill_run_params(R, c, a, own, m_state, jlog, status);
  R.inst_offset = a->instance_offset + e_lo;  // Philox streams are keyed by the global element index
  std::vector<DevModel<NL>> models;
  if (own > 0) build_models<NL>(own, jns + e_lo, layers + (size_t)e_lo * NL, nullptr, 0, g->top_layer, a->timeStep, models, R, nullptr);
  std::vector<double> pos(3 * (size_t)N), pref(N);
  for (int k = 0; k < N; k++) {
    for (int d = 0; d < 3; d++) pos[(size_t)d * N + k] = g->coordinates[3 * k + d];
    const cmtjb_layer_t &L = layers[(size_t)k * NL + g->top_layer];
    pref[k] = (L.Ms / jns[k].MAGNETIC_PERMEABILITY) * (L.thickness * L.cellSurface);
  }
  DevModel<NL> *d_models = nullptr;
  double *d_pos = nullptr, *d_pref = nullptr, *d_x = nullptr, *d_hpart = nullptr;
  RunParams *d_R = nullptr;
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  const int tiles = std::max(1, (own + 31) / 32);
  // partner slices: a function of N alone (the sum order must not depend on the number of ranks), enough of them to fill
  // one GPU's share of the machine when the array is small
  const int P = (int)std::max<long long>(1, std::min<long long>(std::min<long long>(64, N), ((long long)sms * 2048) / (((long long)N + 31) / 32 * 32)));
  cudaError_t e = upload(models.data(), models.size(), &d_models, stream);
  if (e == cudaSuccess) e = upload(pos.data(), pos.size(), &d_pos, stream);
  if (e == cudaSuccess) e = upload(pref.data(), pref.size(), &d_pref, stream);
  if (e == cudaSuccess) e = upload(&R, 1, &d_R, stream);
  if (e == cudaSuccess && !peer) e = cudaMallocAsync((void **)&d_x, (size_t)world * 3 * per * sizeof(double), stream);
  if (e == cudaSuccess && !peer) e = cudaMemsetAsync(d_x, 0, (size_t)world * 3 * per * sizeof(double), stream);
  if (e == cudaSuccess) e = cudaMallocAsync((void **)&d_hpart, (size_t)P * 3 * std::max(own, 1) * sizeof(double), stream);
  int rc = (e == cudaSuccess) ? CMTJB_OK : cuda_fail(e, "sharded group set-up");
  double *x_own = d_x ? d_x + (size_t)rank * 3 * per : nullptr;
  // layout of an exchange area (peer mode): X[2][world][3][per], arrived[world], blocks_done, error
  const size_t x_doubles = (size_t)world * 3 * per;
  auto area_x = [&](int r, int parity) { return reinterpret_cast<double *>(areas[r]) + (size_t)parity * x_doubles; };
  auto area_arrived = [&](int r) { return reinterpret_cast<unsigned long long *>(reinterpret_cast<double *>(areas[r]) + 2 * x_doubles); };
  unsigned int *own_words = peer ? reinterpret_cast<unsigned int *>(area_arrived(rank) + world) : nullptr;  // blocks_done, error
  unsigned long long timeout_ns = 20ull * 1000000000ull;
  if (const char *cto = std::getenv("CMTJB_PEER_TIMEOUT_MS")) timeout_ns = (unsigned long long)std::max(1, std::atoi(cto)) * 1000000ull;
  PeerPush push;
  std::memset(&push, 0, sizeof(push));
  auto aim = [&](int parity, unsigned long long epoch) {  // this rank's rows of buffer `parity` in every area
    push.n = world, push.rank = rank, push.epoch = epoch, push.blocks_done = own_words;
    for (int r = 0; r < world; r++) {
      push.x[r] = area_x(r, parity) + (size_t)rank * 3 * per;
      push.arrived[r] = area_arrived(r) + rank;
    }
  };
  // initial exchange: the top-layer rows of the incoming state
  if (peer) {
    if (rc == CMTJB_OK) {
      aim(0, 1);
      const double *mx = m_state + (size_t)(3 * g->top_layer) * own;
      group_peer_publish_kernel<<<(unsigned)std::max(1, (own + 127) / 128), 128, 0, stream>>>(mx, mx + own, mx + 2 * (size_t)own, own, per, push);
      g_launches++;
      e = cudaGetLastError();
      if (e != cudaSuccess) rc = cuda_fail(e, "group_peer_publish_kernel");
    }
  } else {
    for (int cpt = 0; cpt < 3 && rc == CMTJB_OK && own > 0; cpt++) {
      e = cudaMemcpyAsync(x_own + (size_t)cpt * per, m_state + (size_t)(3 * g->top_layer + cpt) * own, (size_t)own * sizeof(double),
                          cudaMemcpyDeviceToDevice, stream);
      if (e != cudaSuccess) rc = cuda_fail(e, "sharded group initial state");
    }
    if (rc == CMTJB_OK && exchange(ctx, d_x, (int64_t)3 * per, (int64_t)-1, (void *)stream) != 0) rc = fail(CMTJB_ERUNTIME, "exchange callback failed");
  }
  const bool rk4 = (c.solver == CMTJB_RK4);
  const bool lean = own > 0 && lean_ok(system_features(own, NL, jns + e_lo, layers + (size_t)e_lo * NL, nullptr, 0, a, c.solver, false));
  PairShardParams Q;
  Q.N = N, Q.P = P, Q.own = own, Q.e_lo = e_lo, Q.per = per;
  Q.pos = d_pos, Q.pref = d_pref, Q.xbuf = d_x, Q.Hpart = d_hpart;
  GroupStepParams G;
  std::memset(&G, 0, sizeof(G));
  G.N = own, G.P = P, G.interaction = g->interaction, G.top = g->top_layer, G.inst = 0, G.Hpart = d_hpart;
  G.xout = peer ? nullptr : x_own, G.xpitch = per;
  long long sample = R.first_sample;
  long long step_no = 0;  // steps of THIS call (exchange epochs count from the call's initial publish)
  const long long we = R.write_every;
  const long long pair_blocks = ((long long)tiles * 32 * P + 255) / 256;
  const bool trace = std::getenv("CMTJB_TRACE_SHARD") != nullptr;  // device time of the step loop, printed per rank
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  if (trace) {
    cudaEventCreate(&ev0), cudaEventCreate(&ev1);
    cudaEventRecord(ev0, stream);
  }
  for (long long it = R.it_begin; it < R.it_end && rc == CMTJB_OK; it++, step_no++) {
    if (peer) {
      // slices number step_no + 1 of every rank must be in before the pair sums read buffer step_no % 2
      group_peer_wait_kernel<<<1, 32, 0, stream>>>(area_arrived(rank), world, (unsigned long long)step_no + 1, own_words + 1, timeout_ns);
      g_launches++;
      Q.xbuf = area_x(rank, (int)(step_no & 1));
      aim((int)((step_no + 1) & 1), (unsigned long long)step_no + 2);
      G.push = push;
    }
    if (own > 0) {
      if (g->interaction != CMTJB_INT_NULL) {
        group_pair_shard_kernel<<<(unsigned)pair_blocks, 256, 0, stream>>>(Q);
        g_launches++;
      }
      G.it = it;
      G.write = (it % we == 0) ? 1 : 0;
      G.sample = sample;
      const unsigned sb = (unsigned)((own + 127) / 128);
      if (rk4) {
        if (lean) group_step_kernel<NL, CMTJB_RK4, F_COUPLED_LEAN, false><<<sb, 128, 0, stream>>>(d_models, d_R, G);
        else group_step_kernel<NL, CMTJB_RK4, F_ALL, false><<<sb, 128, 0, stream>>>(d_models, d_R, G);
      } else {
        if (lean) group_step_kernel<NL, CMTJB_EULER_HEUN, F_COUPLED_LEAN, false><<<sb, 128, 0, stream>>>(d_models, d_R, G);
        else group_step_kernel<NL, CMTJB_EULER_HEUN, F_ALL, false><<<sb, 128, 0, stream>>>(d_models, d_R, G);
      }
      g_launches++;
      e = cudaGetLastError();
      if (e != cudaSuccess) rc = cuda_fail(e, "sharded group pair/step kernels");
    }
    else if (peer) {  // a rank without elements still takes part in the hand-shake
      group_peer_publish_kernel<<<1, 128, 0, stream>>>(nullptr, nullptr, nullptr, 0, per, push);
      g_launches++;
    }
    if (it % we == 0) sample++;
    if (!peer && rc == CMTJB_OK && exchange(ctx, d_x, (int64_t)3 * per, (int64_t)it, (void *)stream) != 0)
      rc = fail(CMTJB_ERUNTIME, "exchange callback failed");
  }
  if (trace) {
    cudaEventRecord(ev1, stream);
    float ms = 0.f;
    if (cudaEventSynchronize(ev1) == cudaSuccess && cudaEventElapsedTime(&ms, ev0, ev1) == cudaSuccess && step_no > 0)
      std::fprintf(stderr, "[cmtjb] sharded group rank %d/%d (%s): %lld steps, %.2f us per step on the device\n", rank, world,
                   peer ? "peer memory" : "exchange callback", step_no, 1e3 * ms / (double)step_no);
    cudaEventDestroy(ev0), cudaEventDestroy(ev1);
  }
  if (peer && rc == CMTJB_OK) {
    // the last slices must have left before anybody frees an area, and a peer that never arrived is an error, not a hang
    group_peer_wait_kernel<<<1, 32, 0, stream>>>(area_arrived(rank), world, (unsigned long long)step_no + 1, own_words + 1, timeout_ns);
    g_launches++;
    unsigned int err = 0;
    e = cudaMemcpyAsync(&err, own_words + 1, sizeof(err), cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    if (e != cudaSuccess) rc = cuda_fail(e, "sharded group (peer exchange)");
    else if (err) {
      char buf[128];
      std::snprintf(buf, sizeof(buf), "peer exchange: rank %u did not deliver its slice within the time limit", err - 1);
      rc = fail(CMTJB_ERUNTIME, buf);
    }
  }
  cudaFreeAsync(d_models, stream);
  cudaFreeAsync(d_pos, stream);
  cudaFreeAsync(d_pref, stream);
  if (d_R) cudaFreeAsync(d_R, stream);
  if (d_x) cudaFreeAsync(d_x, stream);
  if (d_hpart) cudaFreeAsync(d_hpart, stream);
  return rc;
}

int group_run(const cmtjb_group_t *g, const cmtjb_junction_t *jns, const cmtjb_layer_t *layers, const cmtjb_override_t *ovr,
              int n_ovr, const cmtjb_run_args_t *a, double *m_state, double *jlog, int32_t *status, int device,
              cudaStream_t stream) {
  Common c;
  int rc = validate_group(g, jns, layers, a, &c);
  if (rc) return rc;
  rc = check_coupled_overrides(c.nl, ovr, n_ovr);
  if (rc) return rc;
  for (int k = 0; k < n_ovr; k++)
    if (ovr[k].layer == CMTJB_LAYER_SYSTEM) return fail(CMTJB_EINVAL, "a group has no system-level driver to override");
  if (!m_state) return fail(CMTJB_EINVAL, "m_state is null");
  switch (c.nl) {
  case 1: return group_launch<1>(g, jns, layers, ovr, n_ovr, a, c, m_state, jlog, status, device, stream);
  case 2: return group_launch<2>(g, jns, layers, ovr, n_ovr, a, c, m_state, jlog, status, device, stream);
  default: return fail(CMTJB_ENOTSUP, "dipole-coupled junctions with more than 2 layers are not supported by this build");
  }
}

// host-pointer wrapper shared by both systems: stage, run, copy back
struct HostStage {
  std::vector<void *> allocs;
  ~HostStage() {
    for (void *p : allocs) cudaFree(p);
  }
  template <class T> int in(const T *h, size_t n, T **d) {
    *d = nullptr;
    cudaError_t e = cudaMalloc((void **)d, std::max<size_t>(n, 1) * sizeof(T));
    if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc(coupled host staging)");
    allocs.push_back(*d);
    if (h && n) {
      e = cudaMemcpy(*d, h, n * sizeof(T), cudaMemcpyHostToDevice);
      if (e != cudaSuccess) return cuda_fail(e, "cudaMemcpy H2D");
    }
    return CMTJB_OK;
  }
};

int finish_status(const int32_t *d_status, int32_t *status, long long n) {
  std::vector<int32_t> st(n);
  cudaError_t e = cudaMemcpy(st.data(), d_status, n * sizeof(int32_t), cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) return cuda_fail(e, "cudaMemcpy D2H(status)");
  if (status) std::memcpy(status, st.data(), n * sizeof(int32_t));
  for (long long i = 0; i < n; i++)
    if (st[i]) return fail(CMTJB_ENAN, "NAN magnetisation");
  return CMTJB_OK;
}

}  // namespace
}  // namespace cmtjb

using namespace cmtjb;

extern "C" {

int cmtjb_stack_plan(const cmtjb_stack_t *st, const cmtjb_junction_t *jns, const cmtjb_layer_t *layers,
                     const cmtjb_run_args_t *args, cmtjb_coupled_plan_t *plan) {
  if (!plan) return fail(CMTJB_EINVAL, "null argument");
  Common c;
  int rc = validate_stack(st, jns, layers, args, &c);
  if (rc) return rc;
  fill_plan(c, 1 + st->n_junctions + (st->n_junctions - 1), plan);
  return CMTJB_OK;
}

int cmtjb_stack_run_device(const cmtjb_stack_t *st, const cmtjb_junction_t *jns, const cmtjb_layer_t *layers,
                           const cmtjb_override_t *overrides, int32_t n_overrides, const cmtjb_run_args_t *args,
                           double *m_state, double *junction_log, double *stack_log, int32_t *status, int device,
                           void *stream) {
  CUDA_TRY(cudaSetDevice(device));
  return stack_run(st, jns, layers, overrides, n_overrides, args, m_state, junction_log, stack_log, status, (cudaStream_t)stream);
}

int cmtjb_stack_run_host(const cmtjb_stack_t *st, const cmtjb_junction_t *jns, const cmtjb_layer_t *layers,
                         const cmtjb_override_t *overrides, int32_t n_overrides, const cmtjb_run_args_t *args,
                         double *m_state, double *junction_log, double *stack_log, int32_t *status, int device) {
  cmtjb_coupled_plan_t plan;
  int rc = cmtjb_stack_plan(st, jns, layers, args, &plan);
  if (rc) return rc;
  rc = check_coupled_overrides(plan.n_layers, overrides, n_overrides);
  if (rc) return rc;
  if (!m_state) return fail(CMTJB_EINVAL, "m_state is null");
  CUDA_TRY(cudaSetDevice(device));
  const long long n_inst = args->n_instances, N = st->n_junctions, E = n_inst * N;
  if (E <= 0) return CMTJB_OK;
  const bool want_log = args->log_mode == CMTJB_LOG_FULL && plan.n_samples > 0;
  const size_t nm = (size_t)plan.n_layers * 3 * E;
  const size_t njl = (want_log && junction_log) ? (size_t)plan.n_junction_channels * plan.n_samples * E : 0;
  const size_t nsl = (want_log && stack_log) ? (size_t)plan.n_system_channels * plan.n_samples * n_inst : 0;
  HostStage hs;
  double *d_m, *d_jl, *d_sl;
  int32_t *d_st;
  if ((rc = hs.in(m_state, nm, &d_m))) return rc;
  if ((rc = hs.in((const double *)nullptr, njl, &d_jl))) return rc;
  if ((rc = hs.in((const double *)nullptr, nsl, &d_sl))) return rc;
  if ((rc = hs.in((const int32_t *)nullptr, (size_t)E, &d_st))) return rc;
  CUDA_TRY(cudaMemset(d_st, 0, E * sizeof(int32_t)));
  std::vector<cmtjb_override_t> d_ovr(overrides, overrides + n_overrides);
  for (int k = 0; k < n_overrides; k++) {
    double *p;
    const size_t n = (overrides[k].layer == CMTJB_LAYER_SYSTEM) ? (size_t)n_inst : (size_t)E;
    if ((rc = hs.in(overrides[k].values, n, &p))) return rc;
    d_ovr[k].values = p;
  }
  rc = stack_run(st, jns, layers, d_ovr.data(), n_overrides, args, d_m, njl ? d_jl : nullptr, nsl ? d_sl : nullptr, d_st, nullptr);
  if (rc) return rc;
  CUDA_TRY(cudaDeviceSynchronize());
  CUDA_TRY(cudaMemcpy(m_state, d_m, nm * sizeof(double), cudaMemcpyDeviceToHost));
  if (njl) CUDA_TRY(cudaMemcpy(junction_log, d_jl, njl * sizeof(double), cudaMemcpyDeviceToHost));
  if (nsl) CUDA_TRY(cudaMemcpy(stack_log, d_sl, nsl * sizeof(double), cudaMemcpyDeviceToHost));
  return finish_status(d_st, status, E);
}

int cmtjb_group_plan(const cmtjb_group_t *g, const cmtjb_junction_t *jns, const cmtjb_layer_t *layers,
                     const cmtjb_run_args_t *args, cmtjb_coupled_plan_t *plan) {
  if (!plan) return fail(CMTJB_EINVAL, "null argument");
  Common c;
  int rc = validate_group(g, jns, layers, args, &c);
  if (rc) return rc;
  fill_plan(c, 0, plan);
  plan->cooperative = (g->n_junctions > kCoupledBlockMax / 2) ? 1 : 0;
  return CMTJB_OK;
}

int cmtjb_group_run_device(const cmtjb_group_t *g, const cmtjb_junction_t *jns, const cmtjb_layer_t *layers,
                           const cmtjb_override_t *overrides, int32_t n_overrides, const cmtjb_run_args_t *args,
                           double *m_state, double *junction_log, int32_t *status, int device, void *stream) {
  CUDA_TRY(cudaSetDevice(device));
  return group_run(g, jns, layers, overrides, n_overrides, args, m_state, junction_log, status, device, (cudaStream_t)stream);
}

int cmtjb_group_run_host(const cmtjb_group_t *g, const cmtjb_junction_t *jns, const cmtjb_layer_t *layers,
                         const cmtjb_override_t *overrides, int32_t n_overrides, const cmtjb_run_args_t *args,
                         double *m_state, double *junction_log, int32_t *status, int device) {
  cmtjb_coupled_plan_t plan;
  int rc = cmtjb_group_plan(g, jns, layers, args, &plan);
  if (rc) return rc;
  rc = check_coupled_overrides(plan.n_layers, overrides, n_overrides);
  if (rc) return rc;
  if (!m_state) return fail(CMTJB_EINVAL, "m_state is null");
  CUDA_TRY(cudaSetDevice(device));
  const long long n_inst = args->n_instances, N = g->n_junctions, E = n_inst * N;
  if (E <= 0) return CMTJB_OK;
  const bool want_log = args->log_mode == CMTJB_LOG_FULL && plan.n_samples > 0;
  const size_t nm = (size_t)plan.n_layers * 3 * E;
  const size_t njl = (want_log && junction_log) ? (size_t)plan.n_junction_channels * plan.n_samples * E : 0;
  HostStage hs;
  double *d_m, *d_jl;
  int32_t *d_st;
  if ((rc = hs.in(m_state, nm, &d_m))) return rc;
  if ((rc = hs.in((const double *)nullptr, njl, &d_jl))) return rc;
  if ((rc = hs.in((const int32_t *)nullptr, (size_t)E, &d_st))) return rc;
  CUDA_TRY(cudaMemset(d_st, 0, E * sizeof(int32_t)));
  std::vector<cmtjb_override_t> d_ovr(overrides, overrides + n_overrides);
  for (int k = 0; k < n_overrides; k++) {
    double *p;
    if ((rc = hs.in(overrides[k].values, (size_t)E, &p))) return rc;
    d_ovr[k].values = p;
  }
  rc = group_run(g, jns, layers, d_ovr.data(), n_overrides, args, d_m, njl ? d_jl : nullptr, d_st, device, nullptr);
  if (rc) return rc;
  CUDA_TRY(cudaDeviceSynchronize());
  CUDA_TRY(cudaMemcpy(m_state, d_m, nm * sizeof(double), cudaMemcpyDeviceToHost));
  if (njl) CUDA_TRY(cudaMemcpy(junction_log, d_jl, njl * sizeof(double), cudaMemcpyDeviceToHost));
  return finish_status(d_st, status, E);
}

int cmtjb_group_run_sharded(const cmtjb_group_t *g, const cmtjb_junction_t *jns, const cmtjb_layer_t *layers,
                            const cmtjb_run_args_t *args, int32_t rank, int32_t world, cmtjb_exchange_fn exchange, void *ctx,
                            double *m_state, double *junction_log, int32_t *status, int device, void *stream) {
  if (!exchange || world < 1 || rank < 0 || rank >= world) return fail(CMTJB_EINVAL, "bad rank / world / exchange callback");
  if (!args || args->n_instances != 1) return fail(CMTJB_EINVAL, "a sharded array is one instance");
  Common c;
  int rc = validate_group(g, jns, layers, args, &c);
  if (rc) return rc;
  if (!m_state) return fail(CMTJB_EINVAL, "m_state is null");
  if (args->onef_buffer) return fail(CMTJB_ENOTSUP, "1/f noise buffers are not supported for sharded arrays");
  if (g->interaction == CMTJB_INT_TABLE) return fail(CMTJB_ENOTSUP, "tabulated interaction functions are not supported for sharded arrays");
  for (long long k = 0; k < (long long)g->n_junctions * c.nl; k++)
    if (layers[k].noiseScale != 0.0)
      return fail(CMTJB_ENOTSUP, "1/f noise is not supported for sharded arrays");
  CUDA_TRY(cudaSetDevice(device));
  switch (c.nl) {
  case 1: return group_shard_launch<1>(g, jns, layers, args, c, rank, world, exchange, ctx, nullptr, m_state, junction_log, status, device, (cudaStream_t)stream);
  case 2: return group_shard_launch<2>(g, jns, layers, args, c, rank, world, exchange, ctx, nullptr, m_state, junction_log, status, device, (cudaStream_t)stream);
  default: return fail(CMTJB_ENOTSUP, "dipole-coupled junctions with more than 2 layers are not supported by this build");
  }
}

int64_t cmtjb_group_peer_area_bytes(int32_t n_junctions, int32_t world) {
  if (n_junctions <= 0 || world <= 0) return 0;
  const int64_t per = (n_junctions + world - 1) / world;
  return 2 * (int64_t)world * 3 * per * (int64_t)sizeof(double) + (int64_t)world * (int64_t)sizeof(unsigned long long) + 64;
}

int cmtjb_group_run_sharded_peer(const cmtjb_group_t *g, const cmtjb_junction_t *jns, const cmtjb_layer_t *layers,
                                 const cmtjb_run_args_t *args, int32_t rank, int32_t world, void *const *areas, double *m_state,
                                 double *junction_log, int32_t *status, int device, void *stream) {
  if (!areas || world < 1 || world > kMaxPeers || rank < 0 || rank >= world) return fail(CMTJB_EINVAL, "bad rank / world (at most 16 ranks) / areas");
  for (int r = 0; r < world; r++)
    if (!areas[r]) return fail(CMTJB_EINVAL, "null exchange area");
  if (!args || args->n_instances != 1) return fail(CMTJB_EINVAL, "a sharded array is one instance");
  Common c;
  int rc = validate_group(g, jns, layers, args, &c);
  if (rc) return rc;
  if (!m_state) return fail(CMTJB_EINVAL, "m_state is null");
  if (args->onef_buffer) return fail(CMTJB_ENOTSUP, "1/f noise buffers are not supported for sharded arrays");
  if (g->interaction == CMTJB_INT_TABLE) return fail(CMTJB_ENOTSUP, "tabulated interaction functions are not supported for sharded arrays");
  for (long long k = 0; k < (long long)g->n_junctions * c.nl; k++)
    if (layers[k].noiseScale != 0.0) return fail(CMTJB_ENOTSUP, "1/f noise is not supported for sharded arrays");
  CUDA_TRY(cudaSetDevice(device));
  switch (c.nl) {
  case 1: return group_shard_launch<1>(g, jns, layers, args, c, rank, world, nullptr, nullptr, areas, m_state, junction_log, status, device, (cudaStream_t)stream);
  case 2: return group_shard_launch<2>(g, jns, layers, args, c, rank, world, nullptr, nullptr, areas, m_state, junction_log, status, device, (cudaStream_t)stream);
  default: return fail(CMTJB_ENOTSUP, "dipole-coupled junctions with more than 2 layers are not supported by this build");
  }
}

/* device memory every process of the box can map (cudaMalloc, zero-filled: stream-ordered pool memory cannot be exported) */
int cmtjb_peer_alloc(int device, int64_t bytes, void **ptr) {
  if (!ptr || bytes <= 0) return fail(CMTJB_EINVAL, "bad argument");
  CUDA_TRY(cudaSetDevice(device));
  CUDA_TRY(cudaMalloc(ptr, (size_t)bytes));
  CUDA_TRY(cudaMemset(*ptr, 0, (size_t)bytes));
  CUDA_TRY(cudaDeviceSynchronize());
  return CMTJB_OK;
}
int cmtjb_peer_free(int device, void *ptr) {
  CUDA_TRY(cudaSetDevice(device));
  if (ptr) CUDA_TRY(cudaFree(ptr));
  return CMTJB_OK;
}
int cmtjb_peer_export(int device, void *ptr, unsigned char *handle64) {
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
  if (!ptr || !handle64) return fail(CMTJB_EINVAL, "bad argument");
  CUDA_TRY(cudaSetDevice(device));
  cudaIpcMemHandle_t h;
  CUDA_TRY(cudaIpcGetMemHandle(&h, ptr));
  std::memcpy(handle64, &h, 64);
  return CMTJB_OK;
}
int cmtjb_peer_open(int device, const unsigned char *handle64, void **ptr) {
  if (!ptr || !handle64) return fail(CMTJB_EINVAL, "bad argument");
  CUDA_TRY(cudaSetDevice(device));
  cudaIpcMemHandle_t h;
  std::memcpy(&h, handle64, 64);
  CUDA_TRY(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return CMTJB_OK;
}
int cmtjb_peer_close(int device, void *ptr) {
  CUDA_TRY(cudaSetDevice(device));
  if (ptr) CUDA_TRY(cudaIpcCloseMemHandle(ptr));
  return CMTJB_OK;
}

} /* extern "C" */
