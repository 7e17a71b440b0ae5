// cabi.cu -- extern "C" entry points of libcmtj_b200.so (include/cmtj_b200.h).
//
// Host side of the library: validates a raw junction model the way the reference's
// constructors / runSimulation do, derives the constant-bank DevModel, picks the leanest
// compiled kernel variant that covers the model's terms, and launches it.  No torch, no
// Python, no C++ types across the boundary.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <mutex>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "dev_model.h"
#include "host_common.h"
#include "llgs_kernels.cuh"
#include "onef.h"

namespace cmtjb {
// variant tables, one per translation unit (variants_*.cu)
int variants_rk4(const Variant **out);
int variants_heun(const Variant **out);
int variants_euler_heun(const Variant **out);
int variants_dopri(const Variant **out);
int variants_post(const Variant **out);

thread_local std::string g_err;
std::atomic<long long> g_launches{0};
thread_local std::string g_last_variant;  // name of the kernel variant of the calling thread's last sweep launch  // kernels launched by this library (bench.py: gpu_launches)

int fail(int code, const std::string &msg) {
  g_err = msg;
  return code;
}
void keep_pool_memory() {
  static std::atomic<unsigned long long> done{0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return;
  if (done.load(std::memory_order_relaxed) & (1ull << dev)) return;
  cudaMemPool_t pool = nullptr;
  if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess && pool) {
    // enough for every per-launch scratch block; the large stream-ordered allocations (generated 1/f^alpha buffers, spectrum
    // slabs) still go back to the OS at the next synchronisation, where torch's own allocator can have them
    unsigned long long keep = 256ull << 20;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
  }
  cudaGetLastError();
  done.fetch_or(1ull << dev, std::memory_order_relaxed);
}

int cuda_fail(cudaError_t e, const char *what) {
  return fail(CMTJB_ECUDA, std::string("CUDA error in ") + what + ": " + cudaGetErrorString(e));
}

namespace {

bool slot_dynamic(const cmtjb_driver_t &d) { return d.update != CMTJB_UPD_CONSTANT; }

}  // namespace

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
  bool any_custom = false;
  for (int i = 0; i < jn->n_layers; i++) {
    const cmtjb_layer_t &L = layers[i];
    if (L.temperatureSet || L.noiseScale != 0) {
      if (mode != CMTJB_HEUN && mode != CMTJB_EULER_HEUN) {
        mode = CMTJB_EULER_HEUN;
        p->solver_overridden = 1;
      }
    }
    if (L.noiseScale != 0) {
      // setAlphaNoise (junction.hpp:406-414) / VectorAlphaNoise (noise.hpp:285-287)
      if (L.noiseAlpha < 0 || L.noiseAlpha > 2) return fail(CMTJB_ERUNTIME, "alpha must be between 0 and 2");
      if (L.noiseAxis < CMTJB_AXIS_X || L.noiseAxis > CMTJB_AXIS_ALL)
        return fail(CMTJB_ERUNTIME, "Invalid axis specified: " + std::to_string(L.noiseAxis));
    }
    if (L.temperatureSet && L.cellSurface * L.thickness == 0.0)
      return fail(CMTJB_ERUNTIME, "Cell surface cannot be 0 during temp. calculations!");
    if (L.Ms * L.thickness == 0) return fail(CMTJB_ERUNTIME, "Failed to divide vector by zero!");
    for (int s = 0; s < CMTJB_NSLOTS; s++) {
      const int u = L.drivers[s].update;
      if (u == CMTJB_UPD_CUSTOM) {
        // pre-evaluated by the caller at the stage times of a FIXED-step solver (cmtjb_run_args_t::custom_table)
        any_custom = true;
        if (L.drivers[s].custom_column < 0) return fail(CMTJB_EINVAL, "custom driver without a table column");
        continue;
      }
      if (u < CMTJB_UPD_CONSTANT || u > CMTJB_UPD_GAUSSSTEP) return fail(CMTJB_EINVAL, "unknown driver update type");
    }
  }
  p->effective_solver = mode;
  if (any_custom && mode == CMTJB_DORMAND_PRINCE)
    return fail(CMTJB_ENOTSUP, "custom (Python callback) drivers need a fixed-step solver: Dormand-Prince evaluates drivers at adaptive times");
  p->n_channels = 3 * jn->n_layers + (jn->mr_mode == CMTJB_MR_CLASSIC ? 1 : jn->mr_mode == CMTJB_MR_STRIP ? 3 : 0);
  if (a->log_mode == CMTJB_LOG_ENERGY)
    for (int i = 0; i < jn->n_layers; i++) p->n_channels += layers[i].includeSOT ? CMTJB_ENERGY_CHANNELS_SOT : CMTJB_ENERGY_CHANNELS;
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

namespace {

bool any_nonzero(const double *p, int n) {
  for (int i = 0; i < n; i++)
    if (p[i] != 0.0) return true;
  return false;
}

// K1 anisotropy + demag fold into one DIAGONAL operator when K1 is constant and uniform, the
// easy axis is axis-aligned and the demag tensor is diagonal (DevLayer::lin).
bool layer_foldable(const cmtjb_layer_t &L, int l, const cmtjb_override_t *ovr, int n_ovr) {
  if (slot_dynamic(L.drivers[CMTJB_SLOT_K1])) return false;
  for (int k = 0; k < n_ovr; k++)
    if ((ovr[k].layer < 0 || ovr[k].layer == l) && ovr[k].slot == CMTJB_SLOT_K1) return false;
  const int nz = (L.anis[0] != 0) + (L.anis[1] != 0) + (L.anis[2] != 0);
  if (nz > 1) return false;
  const double *T = L.demagTensor;
  return T[1] == 0 && T[2] == 0 && T[3] == 0 && T[5] == 0 && T[6] == 0 && T[7] == 0;
}

// Fast-field eligibility (dev_model.h: F_FAST): every time-dependent driver is a plain FIELD driver evaluable on
// the device, at most one per layer, and every override other than a constant field component addresses such
// a driver.  Then nothing but that one driver needs inline evaluation or the override tables in the time loop.
bool fast_field_eligible(const cmtjb_junction_t *jn, const cmtjb_layer_t *layers, const cmtjb_override_t *ovr, int n_ovr) {
  bool any = false;
  for (int l = 0; l < jn->n_layers; l++) {
    int n_dyn = 0;
    for (int s = 0; s < CMTJB_NSLOTS; s++) {
      const cmtjb_driver_t &d = layers[l].drivers[s];
      if (!slot_dynamic(d)) continue;
      if (s >= CMTJB_SLOT_K1 || d.update == CMTJB_UPD_CUSTOM) return false;
      n_dyn++;
    }
    if (n_dyn > 1) return false;
    any = any || n_dyn == 1;
  }
  for (int k = 0; k < n_ovr; k++) {
    const bool field_slot = ovr[k].slot < CMTJB_SLOT_K1;
    if (ovr[k].field == CMTJB_FIELD_CONSTANT && field_slot) continue;  // summed into the constant field once per launch
    if (!field_slot) return false;
    for (int l = 0; l < jn->n_layers; l++)
      if ((ovr[k].layer < 0 || ovr[k].layer == l) && !slot_dynamic(layers[l].drivers[ovr[k].slot])) return false;
  }
  return any;
}

// which terms does this model need compiled in?  Returns the union over the layers; `per_layer`
// (optional, [n_layers]) receives each layer's own set (split variants, dev_model.h: F_SPLIT).
uint32_t needed_features(const cmtjb_junction_t *jn, const cmtjb_layer_t *layers, const cmtjb_override_t *ovr, int n_ovr,
                         const cmtjb_run_args_t *a, int solver, uint32_t *per_layer = nullptr) {
  const int n = jn->n_layers;
  uint32_t Fl[CMTJB_MAX_LAYERS] = {0};
  // a slot is "live" if its constant is non-zero, it is time dependent, or it is overridden
  auto live = [&](int l, int s) {
    const cmtjb_driver_t &d = layers[l].drivers[s];
    if (d.constantValue != 0.0 || slot_dynamic(d)) return true;
    for (int k = 0; k < n_ovr; k++)
      if ((ovr[k].layer < 0 || ovr[k].layer == l) && ovr[k].slot == s) return true;
    return false;
  };
  for (int l = 0; l < n && l < CMTJB_MAX_LAYERS; l++) {
    const cmtjb_layer_t &L = layers[l];
    uint32_t F = 0;
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
      if (L.noiseScale != 0) F |= F_ONEF;  // the caller's buffer, or generated on the device (onef.cu)
    }
    for (int s = 0; s < CMTJB_NSLOTS; s++)
      if (slot_dynamic(L.drivers[s])) F |= (s < CMTJB_SLOT_K1) ? F_DYN : F_DYNS;
    if (!layer_foldable(L, l, ovr, n_ovr)) F |= F_NOFOLD;
    for (int k = 0; k < n_ovr; k++) {
      if (!(ovr[k].layer < 0 || ovr[k].layer == l)) continue;
      if (!(ovr[k].field == CMTJB_FIELD_CONSTANT && ovr[k].slot < CMTJB_SLOT_K1)) F |= F_OVR;
      if (ovr[k].field == CMTJB_FIELD_CONSTANT) continue;
      // a per-instance waveform parameter makes the driver non-uniform across the batch: the shared
      // stage table cannot hold it.  Scalar slots keep the table for per-instance AMPLITUDE only.
      const bool field_slot = ovr[k].slot < CMTJB_SLOT_K1;
      if (!field_slot && ovr[k].field == CMTJB_FIELD_AMPLITUDE) continue;
      if (slot_dynamic(L.drivers[ovr[k].slot])) F |= F_DYNI;
    }
    Fl[l] = F;
  }
  uint32_t U = 0;
  for (int l = 0; l < n && l < CMTJB_MAX_LAYERS; l++) U |= Fl[l];
  if ((U & (F_DYN | F_DYNS)) && solver == CMTJB_DORMAND_PRINCE) U |= F_DYNI;  // adaptive time: no stage table
  // one evaluation mode per launch: if any layer evaluates inline, every time-dependent driver does
  if (U & F_DYNI) {
    for (int l = 0; l < n && l < CMTJB_MAX_LAYERS; l++)
      if (Fl[l] & (F_DYN | F_DYNS | F_DYNI)) Fl[l] = (Fl[l] & ~(F_DYN | F_DYNS)) | F_DYNI;
    U &= ~(F_DYN | F_DYNS);
  }
  if (per_layer)
    for (int l = 0; l < n && l < CMTJB_MAX_LAYERS; l++) per_layer[l] = Fl[l];
  return U;
}

// `need`: union of the layers' needs plus launch-wide bits (F_POST, forced F_ALL); `per_layer`: each layer's own
// needs, or null.  A plain variant covers the model if it contains `need`; a split variant (F_SPLIT) if every
// layer's needs are contained in the term set that variant compiles for that layer.
const Variant *pick_variant_exact(int nl, int solver, uint32_t need, const uint32_t *per_layer);
const Variant *pick_variant(int nl, int solver, uint32_t need, const uint32_t *per_layer = nullptr) {
  const Variant *v = pick_variant_exact(nl, solver, need, per_layer);
  // projection-only needs are also covered by the full post-processing variants
  if (!v && (need & F_PROJ)) v = pick_variant_exact(nl, solver, (need & ~(uint32_t)F_PROJ) | F_POST, per_layer);
  return v;
}
const Variant *pick_variant_exact(int nl, int solver, uint32_t need, const uint32_t *per_layer) {
  const Variant *tab = nullptr;
  int n = 0;
  switch (solver) {
  case CMTJB_RK4: n = (need & (F_POST | F_PROJ)) ? variants_post(&tab) : variants_rk4(&tab); break;
  case CMTJB_HEUN: n = variants_heun(&tab); break;
  case CMTJB_EULER_HEUN: n = variants_euler_heun(&tab); break;
  case CMTJB_DORMAND_PRINCE: n = variants_dopri(&tab); break;
  }
  const Variant *best = nullptr;
  int best_extra = 1 << 30;
  for (int i = 0; i < n; i++) {
    const Variant &v = tab[i];
    if (v.nl != nl) continue;
    int extra = 0;
    if (v.features & F_SPLIT) {
      if (!per_layer) continue;
      const uint32_t launch_wide = need & ~F_BASEMASK;  // none today besides what F_BASEMASK already holds
      if (launch_wide || (need & F_POST) != (v.features & F_POST) || (need & F_PROJ)) continue;
      bool ok = true;
      for (int l = 0; l < nl && ok; l++) {
        const uint32_t have = layer_mask(v.features, l) & ~(uint32_t)F_POST;
        ok = (per_layer[l] & ~have) == 0;
        extra += __builtin_popcount(have & ~per_layer[l]);
      }
      // the launch-wide evaluation mode must match: table variants need a table, inline ones must not mix
      if (ok && ((need & F_DYNI) != 0) != ((v.features & F_DYNI) != 0)) ok = false;
      if (!ok) continue;
    } else {
      if ((v.features & need) != need) continue;
      extra = nl * __builtin_popcount(v.features & ~need);  // every layer carries the superfluous terms
    }
    if (extra < best_extra) {
      best_extra = extra;
      best = &v;
    }
  }
  return best;
}

// sub-range of a run executed by one launch (time-chunked pipelining in cmtjb_run_host)
struct Span {
  long long it_begin = 0, it_end = 0, first_sample = 0;
  long long n_samples = 0;  // > 0: `log` is a buffer of this many sample rows per channel (a time chunk), not the whole log
  // instance sub-range of this launch (i_hi < 0: all) and the device-generated 1/f^alpha buffer covering it (onef.cu)
  long long i_lo = 0, i_hi = -1;
  const float *onef_gen = nullptr;
  long long onef_i0 = 0, onef_B = 0;
};

}  // namespace

// needed_features for the other translation units (coupled.cu picks its kernel variant with it)
uint32_t model_features(const cmtjb_junction_t *jn, const cmtjb_layer_t *layers, const cmtjb_override_t *ovr, int n_ovr,
                        const cmtjb_run_args_t *a, int solver) {
  return needed_features(jn, layers, ovr, n_ovr, a, solver);
}

// When does the SMSP-balanced persistent form (llgs_balanced_kernel) pay?  With W = ceil(B / 32) warps
// over S = 4 * numSMs sub-partitions, the plain form lasts as long as its fullest sub-partition:
//   * the whole sweep co-resident (W / S <= resident warps per sub-partition): each sub-partition holds
//     floor or ceil of W / S warps for the whole run  ->  efficiency (W/S) / ceil(W/S);
//   * several waves of blocks: every block runs the same number of steps, so waves end together and the
//     last, partly filled one costs a full wave  ->  efficiency waves / ceil(waves).
// Balance when that model loses more than 3 % and the launch is long enough to be cut into >= 8 rounds.
BalancePlan plan_balance(const RunParams &R, const void *balanced_kernel, const void *plain_kernel) {
  BalancePlan bp{false, 0, 0, 0, 0, 0};
  const char *force = std::getenv("CMTJB_BALANCE");
  if (force && force[0] == '0') return bp;
  const bool forced = force && force[0] == '1';
  const long long steps = R.it_end - R.it_begin;
  const long long n_inst = R.i_hi - R.i_lo;  // instances of this launch (a sub-batch when noise buffers are generated in pieces)
  if (steps <= 0 || n_inst <= 0) return bp;
  static int n_sm[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return bp;
  if (n_sm[dev] == 0 && cudaDeviceGetAttribute(&n_sm[dev], cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return bp;
  int bal_blocks = 0, plain_blocks = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bal_blocks, balanced_kernel, kBalBlock, 0) != cudaSuccess || bal_blocks <= 0 ||
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&plain_blocks, plain_kernel, kBlock, 0) != cudaSuccess || plain_blocks <= 0) {
    cudaGetLastError();
    return bp;
  }
  bal_blocks = std::min(bal_blocks, 8);  // kBalBlock = 4 warps: at most 8 resident warps per sub-partition
  const long long W = (n_inst + 31) / 32;
  const double mean = (double)W / (4.0 * n_sm[dev]);
  const double cap = plain_blocks * (kBlock / 32) / 4.0;  // resident warps per sub-partition, plain form
  const double waves = mean / cap;
  const double eff = mean <= cap ? mean / std::ceil(mean) : waves / std::ceil(waves);
  // a task is a whole number of log periods ...
  // (nothing is written between tasks in a final-state-only run: any length will do)
  const bool writes = R.log != nullptr || R.post.n_proj > 0 || R.post.sd_mode != CMTJB_SD_NONE;
  const long long we = (writes && R.write_every > 0) ? R.write_every : 1;
  // ... and the launch is cut into about 64 rounds (tail <= 1/64 of the quantisation loss), 128 to 400 steps
  // each (measured on B200, S1: 100..400 steps per task within 0.1 % of each other, 1000+ slower)
  long long target = std::min<long long>(400, std::max<long long>(128, steps / 64));
  if (const char *c = std::getenv("CMTJB_BALANCE_CHUNK")) target = std::max<long long>(1, std::atoll(c));  // tuning knob
  const long long chunk = we * ((target + we - 1) / we);
  const long long rounds = (steps + chunk - 1) / chunk;
  // Resident warps per sub-partition of the balanced form: at most 0.9 W / S, so that a warp drawing a
  // ticket (almost) always finds the predecessor task finished -- tickets are issued round-major, and
  // the W - S k instance-warps not being integrated at any moment are the slack that keeps the queue
  // from stalling (measured on B200, S1 with W / S = 5.28: k = 4 -> 1.15e11 layer-steps/s, k = 5 ->
  // 0.6e11, plain -> 1.04e11).  Below four the FP64 pipe is latency- rather than throughput-bound
  // and time-slicing would serialise what the plain form overlaps.
  long long k = std::min<long long>(bal_blocks, (long long)std::floor(0.9 * mean));
  if (const char *c = std::getenv("CMTJB_BALANCE_WARPS")) k = std::max<long long>(1, std::min<long long>(std::atoll(c), bal_blocks));
  if (!forced && (eff > 0.97 || rounds < 8 || k < 4)) return bp;
  k = std::max<long long>(k, 1);
  bp.use = true;
  bp.chunk = forced ? std::max<long long>(1, std::min(chunk, (steps + 2) / 3)) : chunk;  // forced runs (tests) always see >= 3 rounds
  bp.n_rounds = (steps + bp.chunk - 1) / bp.chunk;
  bp.n_wtasks = W;
  bp.blocks = (int)std::min<long long>(k * n_sm[dev], (W + 3) / 4);
  bp.per_sm = std::getenv("CMTJB_BALANCE_NOPIN") ? 0 : (int)k;
  return bp;
}

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
    if (d.update == CMTJB_UPD_CUSTOM) {
      // the callback's value IS the driver's value (drivers.hpp:294-303 returns it directly)
      D.cst[s] = 0.0;
      o.amplitude = 1.0;
      o.pad = d.custom_column;
    }
    if (slot_dynamic(d)) D.dynmask |= (1u << s);
    D.dyncol[s] = -1;
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
  for (int c = 0; c < 3; c++)
    D.lin[c] = D.cK1 * L.drivers[CMTJB_SLOT_K1].constantValue * L.anis[c] * L.anis[c] - D.demag[4 * c];
  D.alpha = L.damping;
  D.llg = -gyro / (1.0 + L.damping * L.damping);
  D.aJ_c = J.HBAR / (J.ELECTRON_CHARGE * L.Ms * L.thickness);
  D.lam = L.SlonczewskiSpacerLayerParameter;
  D.lam2 = D.lam * D.lam;
  D.stt_num = L.alternativeSTT ? L.spinPolarisation : L.spinPolarisation * D.lam2;
  D.stt_a = L.alternativeSTT ? D.lam : D.lam2 - 1.0;
  D.stt_b = L.alternativeSTT ? 1.0 : D.lam2 + 1.0;
  D.P = L.spinPolarisation;
  D.beta = L.beta;
  D.kappa = L.kappa;
  D.flT = L.fieldLikeTorque;
  D.dlT = L.dampingLikeTorque;
  const double V = L.cellSurface * L.thickness;
  D.sigma_c = (V != 0.0) ? (2.0 * L.damping * J.BOLTZMANN_CONST) / (L.Ms * V * gyro) : 0.0;
  D.sigma0 = std::sqrt(D.sigma_c * L.drivers[CMTJB_SLOT_TEMPERATURE].constantValue);
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
  D.includeSOT = L.includeSOT ? 1 : 0;
  D.iecB_a = L.drivers[CMTJB_SLOT_IEC_BOTTOM].constantValue * D.inv_Mst;
  D.iecB_b = 2 * L.drivers[CMTJB_SLOT_IECQ_BOTTOM].constantValue * D.inv_Mst;
  D.iecT_a = L.drivers[CMTJB_SLOT_IEC_TOP].constantValue * D.inv_Mst;
  D.iecT_b = 2 * L.drivers[CMTJB_SLOT_IECQ_TOP].constantValue * D.inv_Mst;
}

void set_step_products(DevLayer &D, double dt) {
  D.h = D.llg * dt;
  D.h_half = 0.5 * D.h;
  D.h_sixth = D.h * (1.0 / 6.0);
}

// where each layer's generated 1/f^alpha series sit in RunParams::onef_gen, and how tickVector combines them
void set_onef_slots(const cmtjb_layer_t *layers, int n_layers, DevLayer *D) {
  OnefSlot slots[3 * CMTJB_MAX_LAYERS];
  int first[CMTJB_MAX_LAYERS];
  if (onef_slots_of(layers, n_layers, slots, first) < 0) return;  // invalid axis: rejected in make_plan
  for (int l = 0; l < n_layers; l++) {
    D[l].onef_slot = first[l] < 0 ? 0 : first[l];
    D[l].onef_axis = layers[l].noiseAxis;
    D[l].onef_scale = layers[l].noiseScale;
  }
}

namespace {

// RunParams of one launch of a single-junction batch (everything except the override rows)
void fill_run_params(RunParams &R, const cmtjb_junction_t *jn, int NL, const cmtjb_run_args_t *a, const cmtjb_plan_t &plan,
                     double *m_state, double *log, int32_t *status, int32_t *n_samples_out, double *time_log, const Span *span,
                     const cmtjb_post_t *post) {
  R.m_state = m_state;
  R.log = (a->log_mode == CMTJB_LOG_FULL || a->log_mode == CMTJB_LOG_ENERGY) ? log : nullptr;
  R.elog_ch0 = -1;
  if (a->log_mode == CMTJB_LOG_ENERGY)
    R.elog_ch0 = 3 * NL + (jn->mr_mode == CMTJB_MR_CLASSIC ? 1 : jn->mr_mode == CMTJB_MR_STRIP ? 3 : 0);
  R.status = status;
  R.n_samples_out = n_samples_out;
  R.time_log = time_log;
  R.onef = a->onef_buffer;
  R.custom = a->custom_table;
  R.n_custom = a->n_custom;
  R.B = a->n_instances;
  R.inst_offset = a->instance_offset;
  const bool ranged = span != nullptr;
  R.n_samples = (span && span->n_samples > 0) ? span->n_samples : plan.n_samples;
  R.it_begin = ranged ? span->it_begin : a->start_iteration;
  R.it_end = ranged ? span->it_end : plan.total_iterations;
  R.total_iterations = plan.total_iterations;
  R.write_every = plan.write_every > 0 ? plan.write_every : 1;
  R.first_sample = ranged ? span->first_sample : 0;
  R.i_lo = span ? span->i_lo : 0;
  R.i_hi = (span && span->i_hi >= 0) ? span->i_hi : a->n_instances;
  if (span && span->onef_gen) {
    R.onef_gen = span->onef_gen;
    R.onef_i0 = span->onef_i0;
    R.onef_B = span->onef_B;
  }
  R.dt = a->timeStep;
  R.sqrt_dt = std::sqrt(a->timeStep);
  R.total_time = a->totalTime;
  R.write_frequency = a->writeFrequency;
  R.seed = a->seed;
  R.pk = make_philox_keys(a->seed);
  R.normals64 = a->thermal_normals == 1 ? 1 : 0;
  if (const char *c = std::getenv("CMTJB_TUNE")) R.tune_flags = std::atoi(c);
  if (post) {
    PostParams &P = R.post;
    P.proj_log = post->proj_log;
    P.proj_n_samples = post->proj_n_samples > 0 ? post->proj_n_samples : plan.n_samples;
    P.proj_sample0 = post->proj_sample0;
    P.sd_omega = post->sd_omega;
    P.sd_state = post->sd_state;
    P.n_proj = post->n_proj;
    P.sd_mode = post->sd_mode;
    std::memcpy(P.proj, post->proj, sizeof(P.proj));
    std::memcpy(P.res, post->res, sizeof(P.res));
    std::memcpy(P.sd_b, post->sd_b, sizeof(P.sd_b));
    std::memcpy(P.sd_a, post->sd_a, sizeof(P.sd_a));
  }
}

template <int NL>
int run_typed(const cmtjb_junction_t *jn, const cmtjb_layer_t *layers, const cmtjb_override_t *ovr, int n_ovr,
              const cmtjb_run_args_t *a, const cmtjb_plan_t &plan, const Variant *v, double *m_state, double *log,
              int32_t *status, int32_t *n_samples_out, double *time_log, cudaStream_t stream, const Span *span,
              const cmtjb_post_t *post) {
  static_assert(sizeof(DevModel<NL>) + sizeof(RunParams) < 32000, "kernel parameters exceed the 32 KB limit");
  DevModel<NL> M;
  std::memset(&M, 0, sizeof(M));
  for (int l = 0; l < NL; l++) {
    derive_layer(*jn, layers[l], M.L[l]);
    set_step_products(M.L[l], a->timeStep);
  }
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
  fill_run_params(R, jn, NL, a, plan, m_state, log, status, n_samples_out, time_log, span, post);
  if (R.B <= 0 || R.i_hi <= R.i_lo) return CMTJB_OK;
  set_onef_slots(layers, NL, M.L);
  // table of the time-dependent drivers at this launch's stage times (F_DYN / F_DYNS variants)
  double *tab = nullptr;
  if ((v->features & (F_DYN | F_DYNS)) && !(v->features & F_DYNI) && R.it_end > R.it_begin) {
    int ncols = 3 * NL;
    bool any = false;
    for (int l = 0; l < NL; l++) {
      M.L[l].dynH = (M.L[l].dynmask & 0xFFFu) ? 1 : 0;
      any = any || M.L[l].dynmask != 0;
      for (int s = CMTJB_SLOT_K1; s < CMTJB_NSLOTS; s++)
        if ((M.L[l].dynmask >> s) & 1u) M.L[l].dyncol[s] = (int8_t)ncols++;
    }
    R.dyn_ncols = ncols;
    if (any) {
      const size_t bytes = (size_t)ncols * (size_t)(R.it_end - R.it_begin) * 3 * sizeof(double);
      cudaError_t e = cudaMallocAsync((void **)&tab, bytes, stream);
      if (e != cudaSuccess) return cuda_fail(e, "cudaMallocAsync(driver table)");
      R.dyn_tab = tab;
      e = launch_build_dyn_table<NL>(M, R, tab, stream);
      g_launches++;
      if (e != cudaSuccess) {
        cudaFreeAsync(tab, stream);
        return cuda_fail(e, "build_dyn_table_kernel");
      }
    }
  }
  if ((v->features & F_DYNI) && R.custom && R.n_custom > 0) {
    // inline-evaluating variants read pre-evaluated callbacks straight from the caller's table, which has the
    // stage table's [step][stage][column] layout: the row pointer of the time loop walks it instead
    R.dyn_tab = R.custom + 3 * R.it_begin * R.n_custom;
    R.dyn_ncols = R.n_custom;
  }
  cudaError_t e = v->launch(&M, &R, stream);
  g_launches++;
  if (tab) cudaFreeAsync(tab, stream);
  if (e != cudaSuccess) return cuda_fail(e, v->name);
  return CMTJB_OK;
}

// Junctions of 5..CMTJB_MAX_LAYERS layers: the lane-per-layer kernel (llgs_wide_kernel), general term set.
int run_wide(const cmtjb_junction_t *jn, const cmtjb_layer_t *layers, const cmtjb_override_t *ovr, int n_ovr,
             const cmtjb_run_args_t *a, const cmtjb_plan_t &plan, double *m_state, double *log, int32_t *status,
             cudaStream_t stream, const Span *span, const cmtjb_post_t *post) {
  const int NL = jn->n_layers;
  if (NL > kWideGroup || NL > CMTJB_MAX_LAYERS) return fail(CMTJB_ENOTSUP, "junctions with more than 8 layers are not supported by this build");
  if (post) return fail(CMTJB_ENOTSUP, "post-processing supports at most 4 layers");
  if (a->log_mode == CMTJB_LOG_ENERGY) return fail(CMTJB_ENOTSUP, "the energy log supports junctions of at most 4 layers");
  if (plan.effective_solver == CMTJB_DORMAND_PRINCE) return fail(CMTJB_ENOTSUP, "Dormand-Prince is single-layer only");
  std::vector<DevLayer> L(NL);
  for (int l = 0; l < NL; l++) {
    for (int s = 0; s < CMTJB_NSLOTS; s++)
      if (layers[l].drivers[s].update == CMTJB_UPD_CUSTOM &&
          (!a->custom_table || layers[l].drivers[s].custom_column < 0 || layers[l].drivers[s].custom_column >= a->n_custom))
        return fail(CMTJB_EINVAL, "custom driver: custom_table is null or has too few columns");
    derive_layer(*jn, layers[l], L[l]);
    set_step_products(L[l], a->timeStep);
  }
  RunParams R;
  std::memset(&R, 0, sizeof(R));
  for (int k = 0; k < n_ovr; k++) {
    R.ovr[k] = ovr[k].values;
    for (int l = 0; l < NL; l++) {
      if (ovr[k].layer >= 0 && ovr[k].layer != l) continue;
      L[l].ovmask |= (1u << ovr[k].slot);
      L[l].ovidx[ovr[k].slot][ovr[k].field] = (int8_t)k;
    }
  }
  fill_run_params(R, jn, NL, a, plan, m_state, log, status, nullptr, nullptr, span, nullptr);
  if (R.B <= 0 || R.it_end <= R.it_begin || R.i_hi <= R.i_lo) return CMTJB_OK;
  set_onef_slots(layers, NL, L.data());
  DevLayer *dL = nullptr;
  cudaError_t e = cudaMallocAsync((void **)&dL, sizeof(DevLayer) * NL, stream);
  if (e != cudaSuccess) return cuda_fail(e, "cudaMallocAsync(layer models)");
  e = cudaMemcpyAsync(dL, L.data(), sizeof(DevLayer) * NL, cudaMemcpyHostToDevice, stream);
  if (e == cudaSuccess) {
    WideModel W;
    W.L = dL;
    W.n_layers = NL;
    W.mr_mode = jn->mr_mode;
    W.Rp = jn->Rp;
    W.Rhalf = (jn->Rap - jn->Rp) / 2.0;
    switch (plan.effective_solver) {
    case CMTJB_RK4: e = launch_wide<CMTJB_RK4>(W, R, stream); break;
    case CMTJB_HEUN: e = launch_wide<CMTJB_HEUN>(W, R, stream); break;
    default: e = launch_wide<CMTJB_EULER_HEUN>(W, R, stream); break;
    }
    g_launches++;
  }
  cudaFreeAsync(dL, stream);
  if (e != cudaSuccess) return cuda_fail(e, "llgs_wide_kernel");
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

// One launch over the step range `span` (or the whole run) with the leanest variant covering `layers`.
int launch_span(const cmtjb_junction_t *jn, const cmtjb_layer_t *layers, const cmtjb_override_t *ovr, int n_ovr,
                const cmtjb_run_args_t *a, cmtjb_plan_t &plan, double *m_state, double *log, int32_t *status,
                int32_t *n_samples_out, double *time_log, cudaStream_t stream, const Span *span,
                const cmtjb_post_t *post = nullptr) {
  if (jn->n_layers > 4) {
    std::snprintf(plan.variant, sizeof(plan.variant), "wide/L%d/F_ALL", jn->n_layers);
    plan.feature_mask = (int32_t)F_ALL;
    g_last_variant = plan.variant;
    return run_wide(jn, layers, ovr, n_ovr, a, plan, m_state, log, status, stream, span, post);
  }
  uint32_t per_layer[CMTJB_MAX_LAYERS] = {0};
  uint32_t need = needed_features(jn, layers, ovr, n_ovr, a, plan.effective_solver, per_layer);
  bool split_ok = true;  // per-layer matching only when nothing below rewrites the need
  if (need & (F_DYN | F_DYNS)) {
    // the stage table costs 24 B per column per step; beyond 2 GiB evaluate inline instead
    size_t n_cols = 3 * (size_t)jn->n_layers;
    for (int l = 0; l < jn->n_layers; l++)
      for (int s = CMTJB_SLOT_K1; s < CMTJB_NSLOTS; s++) n_cols += slot_dynamic(layers[l].drivers[s]) ? 1 : 0;
    const long long steps = span ? span->it_end - span->it_begin : plan.total_iterations - a->start_iteration;
    if (n_cols * (size_t)std::max<long long>(steps, 0) * 24 > ((size_t)2 << 30)) {
      need = (need & ~(F_DYN | F_DYNS)) | F_DYNI;
      split_ok = false;
    }
  }
  uint32_t need_fast = 0;  // alternative: the fast-field form of an inline-evaluating RK4 model
  if ((need & F_DYNI) && plan.effective_solver == CMTJB_RK4 && a->log_mode != CMTJB_LOG_ENERGY &&
      fast_field_eligible(jn, layers, ovr, n_ovr))
    need_fast = (need & ~(uint32_t)(F_DYNI | F_OVR)) | F_FAST;
  const char *force = std::getenv("CMTJB_FORCE_VARIANT");  // "full": always the most general kernel; "nosplit": no split variants (testing)
  if (force && (std::strcmp(force, "full") == 0 || std::strcmp(force, "nofast") == 0)) need_fast = 0;
  if (force && std::strcmp(force, "full") == 0) need = F_ALL, split_ok = false;
  if (force && std::strcmp(force, "nosplit") == 0) split_ok = false;
  if (a->log_mode == CMTJB_LOG_ENERGY) need = F_ALL, split_ok = false;  // the energy log lives in the general variants only
  // projection log without spin-diode reduction: the lean F_PROJ form (pick_variant falls back to F_POST variants)
  const uint32_t post_bit = (post && post->sd_mode == CMTJB_SD_NONE && post->n_proj > 0) ? F_PROJ : F_POST;
  if (post) need |= post_bit;
  if (post && need_fast) need_fast |= post_bit;
  {
    int max_col = -1;
    for (int l = 0; l < jn->n_layers; l++)
      for (int s = 0; s < CMTJB_NSLOTS; s++)
        if (layers[l].drivers[s].update == CMTJB_UPD_CUSTOM) max_col = std::max(max_col, layers[l].drivers[s].custom_column);
    if (max_col >= 0) {
      if (!a->custom_table || max_col >= a->n_custom) return fail(CMTJB_EINVAL, "custom driver: custom_table is null or has too few columns");
    }
  }
  const Variant *v = need_fast ? pick_variant(jn->n_layers, plan.effective_solver, need_fast) : nullptr;
  if (!v) v = pick_variant(jn->n_layers, plan.effective_solver, need, split_ok ? per_layer : nullptr);
  if (!v) {
    char buf[160];
    std::snprintf(buf, sizeof(buf), "no compiled kernel variant for %d layers, solver %d, features 0x%x", jn->n_layers,
                  plan.effective_solver, need);
    return fail(CMTJB_ENOTSUP, buf);
  }
  g_last_variant = v->name;
  // the plan reports the variant of the (last) segment that still evaluates time-dependent drivers
  if (plan.variant[0] == 0 || (v->features & (F_DYN | F_DYNS | F_DYNI))) {
    plan.feature_mask = (int32_t)v->features;
    std::snprintf(plan.variant, sizeof(plan.variant), "%s", v->name);
  }
  switch (jn->n_layers) {
  case 1: return run_typed<1>(jn, layers, ovr, n_ovr, a, plan, v, m_state, log, status, n_samples_out, time_log, stream, span, post);
  case 2: return run_typed<2>(jn, layers, ovr, n_ovr, a, plan, v, m_state, log, status, n_samples_out, time_log, stream, span, post);
  case 3: return run_typed<3>(jn, layers, ovr, n_ovr, a, plan, v, m_state, log, status, n_samples_out, time_log, stream, span, post);
  case 4: return run_typed<4>(jn, layers, ovr, n_ovr, a, plan, v, m_state, log, status, n_samples_out, time_log, stream, span, post);
  default: return fail(CMTJB_ENOTSUP, "unreachable: junctions with more than 4 layers run the lane-per-layer kernel");
  }
}

// Support of a driver's time-dependent part: step and trapezoid drivers add their amplitude only
// inside a finite window (drivers.hpp:73-108, :262-305); everything else is treated as unbounded.
bool driver_window(const cmtjb_driver_t &d, double *t0, double *t1) {
  switch (d.update) {
  case CMTJB_UPD_STEP:
    *t0 = d.timeStart;
    *t1 = d.timeStop;
    return true;
  case CMTJB_UPD_TRAPEZOID:
    *t0 = d.timeStart;
    *t1 = d.timeStart + 2 * d.edgeTime + d.steadyTime;
    return true;
  default: return false;
  }
}

// Steps [span.it_begin, span.it_end) of instances [span.i_lo, span.i_hi): one launch, or three when the
// time-dependent drivers have finite support (below).
int run_device_spans(const cmtjb_junction_t *jn, const cmtjb_layer_t *layers, const cmtjb_override_t *ovr, int n_ovr,
                     const cmtjb_run_args_t *a, cmtjb_plan_t &plan, double *m_state, double *log, int32_t *status,
                     int32_t *n_samples_out, double *time_log, cudaStream_t stream, const Span *span, const cmtjb_post_t *post) {
  int rc = CMTJB_OK;
  // Time-window specialisation: when every time-dependent driver has finite support and none of its
  // timing fields varies per instance, only the steps inside the union of the windows run the
  // driver-evaluating variant; the steps before and after run the lean constant-driver variant
  // (a PIMM sweep excites for 3 of its 2e5 steps).  State chains through m_state between launches.
  const int n = jn->n_layers;
  bool windowed = plan.effective_solver != CMTJB_DORMAND_PRINCE && a->log_mode != CMTJB_LOG_ENERGY;
  bool any_dynamic = false;
  double w0 = 0, w1 = 0;
  for (int l = 0; l < n && windowed; l++)
    for (int s = 0; s < CMTJB_NSLOTS && windowed; s++) {
      const cmtjb_driver_t &d = layers[l].drivers[s];
      if (!slot_dynamic(d)) continue;
      double t0, t1;
      if (!driver_window(d, &t0, &t1) || !(t1 >= t0)) {
        windowed = false;
        break;
      }
      for (int k = 0; k < n_ovr; k++)
        if ((ovr[k].layer < 0 || ovr[k].layer == l) && ovr[k].slot == s && ovr[k].field >= CMTJB_FIELD_TIMESTART) windowed = false;
      w0 = any_dynamic ? std::min(w0, t0) : t0;
      w1 = any_dynamic ? std::max(w1, t1) : t1;
      any_dynamic = true;
    }
  const long long it0 = span->it_begin, it1 = span->it_end, fs0 = span->first_sample;
  if (!windowed || !any_dynamic || it1 <= it0)
    return launch_span(jn, layers, ovr, n_ovr, a, plan, m_state, log, status, n_samples_out, time_log, stream, span, post);
  // step `it` touches times [it*dt, it*dt + dt]; two steps of margin absorb rounding at the edges
  const double dt = a->timeStep;
  long long a_on = (long long)std::floor(w0 / dt) - 2, a_off = (long long)std::ceil(w1 / dt) + 2;
  a_on = std::min(std::max(a_on, it0), it1);
  a_off = std::min(std::max(a_off + 1, a_on), it1);
  std::vector<cmtjb_layer_t> frozen(layers, layers + n);  // drivers reduced to their constant part
  for (int l = 0; l < n; l++)
    for (int s = 0; s < CMTJB_NSLOTS; s++)
      if (slot_dynamic(frozen[l].drivers[s])) {
        frozen[l].drivers[s].update = CMTJB_UPD_CONSTANT;
        frozen[l].drivers[s].amplitude = 0.0;
      }
  const long long we = plan.write_every > 0 ? plan.write_every : 1;
  auto count_upto = [&](long long m) { return m <= 0 ? 0LL : (m - 1) / we + 1; };
  const long long cuts[4] = {it0, a_on, a_off, it1};
  for (int seg = 0; seg < 3 && rc == CMTJB_OK; seg++) {
    Span sp = *span;
    sp.it_begin = cuts[seg];
    sp.it_end = cuts[seg + 1];
    if (sp.it_end <= sp.it_begin) continue;
    sp.first_sample = fs0 + count_upto(sp.it_begin) - count_upto(it0);
    rc = launch_span(jn, seg == 1 ? layers : frozen.data(), ovr, n_ovr, a, plan, m_state, log, status, n_samples_out, time_log,
                     stream, &sp, post);
  }
  return rc;
}

// Device memory a run may spend on generated 1/f^alpha buffers: half of what is free now (CMTJB_ONEF_BUDGET_MB
// overrides, tests use it to force sub-batching).
size_t onef_budget_bytes() {
  if (const char *c = std::getenv("CMTJB_ONEF_BUDGET_MB")) return (size_t)std::max(1, std::atoi(c)) << 20;
  size_t free_b = 0, total_b = 0;
  if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) free_b = (size_t)8 << 30;
  return free_b / 2;
}

// does this run draw 1/f^alpha noise that the library has to generate?
int onef_needed(const cmtjb_junction_t *jn, const cmtjb_layer_t *layers, const cmtjb_run_args_t *a, const cmtjb_plan_t &plan,
                OnefSlot *slots, int *n_slots) {
  *n_slots = 0;
  if (a->onef_buffer) return CMTJB_OK;
  if (plan.effective_solver != CMTJB_HEUN && plan.effective_solver != CMTJB_EULER_HEUN) return CMTJB_OK;
  int first[CMTJB_MAX_LAYERS];
  const int ns = onef_slots_of(layers, jn->n_layers, slots, first);
  if (ns < 0) return fail(CMTJB_ERUNTIME, "Invalid axis specified");
  *n_slots = ns;
  return CMTJB_OK;
}

int run_device_impl(const cmtjb_junction_t *jn, const cmtjb_layer_t *layers, const cmtjb_override_t *ovr, int n_ovr,
                    const cmtjb_run_args_t *a, double *m_state, double *log, int32_t *status, int32_t *n_samples_out,
                    double *time_log, cudaStream_t stream, cmtjb_plan_t *plan_out, const Span *span_in = nullptr,
                    const cmtjb_post_t *post = nullptr) {
  keep_pool_memory();
  cmtjb_plan_t plan;
  int rc = make_plan(jn, layers, a, &plan);
  if (rc) return rc;
  rc = check_overrides(jn, ovr, n_ovr);
  if (rc) return rc;
  if (!m_state) return fail(CMTJB_EINVAL, "m_state is null");
  plan.variant[0] = 0;
  Span span;
  if (span_in) {
    span = *span_in;
  } else {
    span.it_begin = a->start_iteration;
    span.it_end = plan.total_iterations;
    span.first_sample = 0;
  }
  if (span.i_hi < 0) span.i_hi = a->n_instances;
  OnefSlot slots[3 * CMTJB_MAX_LAYERS];
  int n_slots = 0;
  rc = onef_needed(jn, layers, a, plan, slots, &n_slots);
  if (rc) return rc;
  if (n_slots == 0 || span.onef_gen || span.i_hi <= span.i_lo || plan.total_iterations <= 0) {
    rc = run_device_spans(jn, layers, ovr, n_ovr, a, plan, m_state, log, status, n_samples_out, time_log, stream, &span, post);
    if (plan_out) *plan_out = plan;
    return rc;
  }
  // 1/f^alpha noise (junction.hpp:1905-1916): one float32 buffer of total_iterations samples per (instance, layer,
  // axis).  When the buffers of the whole batch do not fit the budget, the batch is integrated in sub-batches of
  // consecutive instances (a multiple of 32), each with its own generation; results do not depend on the split.
  int dev = 0;
  cudaGetDevice(&dev);
  const long long n = plan.total_iterations, n_inst = span.i_hi - span.i_lo;
  const size_t per_inst = (size_t)n_slots * (size_t)n * sizeof(float);
  const size_t budget = onef_budget_bytes();
  long long Bc = (long long)std::min<size_t>((size_t)n_inst, budget / std::max<size_t>(per_inst, 1));
  if (Bc < n_inst) Bc = (Bc / 32) * 32;
  if (Bc <= 0) {
    char buf[256];
    std::snprintf(buf, sizeof(buf),
                  "1/f^alpha noise: %zu bytes of noise buffer per instance (%d series x %lld iterations x 4 B) do not fit the "
                  "%zu bytes available for 32 instances at a time; shorten the run or free device memory",
                  per_inst, n_slots, n, budget);
    return fail(CMTJB_ERUNTIME, buf);
  }
  float *gen = nullptr;
  cudaError_t e = cudaMallocAsync((void **)&gen, per_inst * (size_t)Bc, stream);
  if (e != cudaSuccess) {
    char buf[200];
    std::snprintf(buf, sizeof(buf), "1/f^alpha noise: cudaMallocAsync of %zu bytes failed (%s)", per_inst * (size_t)Bc, cudaGetErrorString(e));
    return fail(CMTJB_ECUDA, buf);
  }
  for (long long i0 = span.i_lo; i0 < span.i_hi && rc == CMTJB_OK; i0 += Bc) {
    Span sp = span;
    sp.i_lo = i0;
    sp.i_hi = std::min(i0 + Bc, span.i_hi);
    rc = onef_generate(dev, stream, n, slots, n_slots, a->seed, (unsigned long long)(a->instance_offset + i0), sp.i_hi - sp.i_lo, gen);
    if (rc) break;
    sp.onef_gen = gen;
    sp.onef_i0 = i0;
    sp.onef_B = sp.i_hi - sp.i_lo;
    rc = run_device_spans(jn, layers, ovr, n_ovr, a, plan, m_state, log, status, n_samples_out, time_log, stream, &sp, post);
  }
  cudaFreeAsync(gen, stream);
  if (plan_out) *plan_out = plan;
  return rc;
}

// Device scratch reused across cmtjb_run_host calls (one slab per device, grown on demand):
// a sweep loop calling run_host repeatedly should not pay cudaMalloc + first-touch page
// mapping of a multi-GB log buffer on every call.
struct Slab {
  void *p = nullptr;
  size_t bytes = 0;
};
// Where the last cmtjb_run_host call of this thread spent its time (cmtjb_host_stats; bench.py puts it into e2e).
struct HostStats {
  double total_ms = 0, integrate_ms = 0, ring_wait_ms = 0, copy_ms = 0, tail_copy_ms = 0, pipeline_ms = 0, log_bytes = 0;
  int n_chunks = 1, n_ring = 1;
};
thread_local HostStats g_host_stats;
std::mutex g_slab_mu[64];  // one per device: host-entry calls on different devices run concurrently
Slab g_slab[64];

int slab_reserve(int device, size_t bytes, void **out) {
  if (device < 0 || device >= 64) return fail(CMTJB_EINVAL, "device index out of range");
  Slab &s = g_slab[device];
  if (s.bytes < bytes) {
    if (s.p) cudaFree(s.p);
    s.p = nullptr;
    s.bytes = 0;
    cudaError_t e = cudaMalloc(&s.p, bytes);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc(run_host scratch)");
    s.bytes = bytes;
  }
  *out = s.p;
  return CMTJB_OK;
}

// ---- small utility kernels ---------------------------------------------------------------------
__global__ void eval_driver_kernel(const __grid_constant__ DevLayer Lr, const __grid_constant__ RunParams R,
                                   const double *times, long long n, double *out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = slot_value<F_DYNI>(Lr, R, CMTJB_SLOT_K1, StageTime{times[i], nullptr}, 0);
}

// RNG test hooks: the raw Philox block for a given counter/key, and the normal3 stream the
// steppers draw from
__global__ void philox_kernel(uint4 ctr, uint2 key, uint32_t *out) {
  const uint4 r = philox4x32_10(ctr, key);
  out[0] = r.x;
  out[1] = r.y;
  out[2] = r.z;
  out[3] = r.w;
}
__global__ void normals_kernel(unsigned long long seed, long long inst0, long long n_inst, unsigned layer, long long step0,
                               long long n_steps, double *out, int f64) {
  const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n_inst * n_steps) return;
  const long long st = q / n_inst, i = q - st * n_inst;
  const PhiloxKeys K = philox_keys(seed);
  const d3 v = f64 ? normal3_f64(K.x[0], K.y[0], (unsigned long long)(inst0 + i), layer, (unsigned long long)(step0 + st))
                   : normal3(K, (unsigned long long)(inst0 + i), layer, (unsigned long long)(step0 + st));
  out[3 * q + 0] = v.x;
  out[3 * q + 1] = v.y;
  out[3 * q + 2] = v.z;
}

// CH independent DFMA chains per thread, 8 links per chain and iteration; 2 flop per DFMA
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// out[1..4]: SM cycle counter and global nanosecond timer at the start and the end of block 0 (the SM clock the
// microbenchmark actually ran at: sustained DFMA issue is power-limited on this part)
template <int CH> __global__ void __launch_bounds__(256) fp64_peak_kernel(double *out, long long iters, double a, double b) {
  long long c0 = 0;
  unsigned long long n0 = 0;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    c0 = clock64();
    n0 = global_ns();
  }
  double x[CH];
#pragma unroll
  for (int c = 0; c < CH; c++) x[c] = threadIdx.x * 1e-3 + c;
  for (long long k = 0; k < iters; k++) {
#pragma unroll
    for (int u = 0; u < 8; u++) {
#pragma unroll
      for (int c = 0; c < CH; c++) x[c] = fma(x[c], a, b);
    }
  }
  double s = 0.0;
#pragma unroll
  for (int c = 0; c < CH; c++) s += x[c];
  if (s == 123.456) out[0] = s;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    out[1] = (double)(clock64() - c0);
    out[2] = (double)(global_ns() - n0);
  }
}


// FP64 pipe probe (cmtjb_fp64_pipe_probe): what fraction of the pipe's issue slots (64 lanes per SM and clock) a stream of
// independent chains sustains at a given number of warps per sub-partition, for the operand shapes the integration kernels use.
//   KIND 0: x = fma(x, a, b), a and b kernel parameters (constant bank): one register source operand
//   KIND 1: x = fma(x, y, z), all three in registers (y, z differ per chain)
//   KIND 2: the S1 step's instruction mix on three-register operands: 12 DFMA : 3 DMUL : 1 DADD
//   KIND 3: the step's own dependency shape: cross products, dot products and axpy on 3-vectors (ILP 3), neighbour and field
//           terms included (28 FP64 instructions per round, like one right-hand-side evaluation of the two-layer IEC model)
//   KIND 4: KIND 3 on two independent vectors per thread, interleaved (what two layers, or two instances, per thread give)
//   KIND 5: independent chains like KIND 1, but every DFMA reads three registers of THREE different chains, rotating from
//           instruction to instruction (operand variety of real code without its short dependencies)
template <int KIND, int CH> __global__ void __launch_bounds__(128) fp64_probe_kernel(double *out, const double *seed, long long iters, double a, double b) {
  long long c0 = 0;
  unsigned long long n0 = 0;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    c0 = clock64();
    n0 = global_ns();
  }
  double x[CH], y[CH], z[CH];
#pragma unroll
  for (int c = 0; c < CH; c++) {
    x[c] = seed[c] + threadIdx.x * 1e-6;
    y[c] = seed[CH + c];
    z[c] = seed[2 * CH + c];
  }
  if constexpr (KIND == 3) {
    // state m (x[0..2]), neighbour (y[0..2]), field (z[0..2]); constants from registers
    double mx = x[0], my = x[1], mz = x[2];
    const double nx = y[0], ny = y[1], nz = y[2], hx = z[0], hy = z[1], hz = z[2];
    const double l0 = y[3], l1 = y[4], l2 = y[5], ib = z[3], ia = z[4], al = z[5], h = z[6];
    for (long long k = 0; k < iters; k++) {
#pragma unroll
      for (int u = 0; u < 4; u++) {
        double Hx = fma(l0, mx, hx), Hy = fma(l1, my, hy), Hz = fma(l2, mz, hz);               // 3
        const double d = fma(nz, mz, fma(ny, my, nx * mx));                                        // 3
        const double sc = fma(ib, d, ia);                                                          // 1
        Hx = fma(nx, sc, Hx), Hy = fma(ny, sc, Hy), Hz = fma(nz, sc, Hz);                           // 3
        const double px = fma(my, Hz, -(mz * Hy)), py = fma(mz, Hx, -(mx * Hz)), pz = fma(mx, Hy, -(my * Hx));  // 6
        const double ax = al * mx, ay = al * my, az = al * mz;                                     // 3
        const double ux = fma(ay, pz, fma(-az, py, px)), uy = fma(az, px, fma(-ax, pz, py)), uz = fma(ax, py, fma(-ay, px, pz));  // 6
        mx = fma(ux, h, mx), my = fma(uy, h, my), mz = fma(uz, h, mz);                              // 3  -> 28
      }
    }
    x[0] = mx, x[1] = my, x[2] = mz;
  } else if constexpr (KIND == 4) {
    double mx = x[0], my = x[1], mz = x[2], qx = x[2], qy = x[0], qz = x[1];
    const double nx = y[0], ny = y[1], nz = y[2], hx = z[0], hy = z[1], hz = z[2];
    const double l0 = y[3], l1 = y[4], l2 = y[5], ib = z[3], ia = z[4], al = z[5], h = z[6];
    auto rhs = [&](double &ax_, double &ay_, double &az_) {
      double Hx = fma(l0, ax_, hx), Hy = fma(l1, ay_, hy), Hz = fma(l2, az_, hz);
      const double d = fma(nz, az_, fma(ny, ay_, nx * ax_));
      const double sc = fma(ib, d, ia);
      Hx = fma(nx, sc, Hx), Hy = fma(ny, sc, Hy), Hz = fma(nz, sc, Hz);
      const double px = fma(ay_, Hz, -(az_ * Hy)), py = fma(az_, Hx, -(ax_ * Hz)), pz = fma(ax_, Hy, -(ay_ * Hx));
      const double ax = al * ax_, ay = al * ay_, az = al * az_;
      const double ux = fma(ay, pz, fma(-az, py, px)), uy = fma(az, px, fma(-ax, pz, py)), uz = fma(ax, py, fma(-ay, px, pz));
      ax_ = fma(ux, h, ax_), ay_ = fma(uy, h, ay_), az_ = fma(uz, h, az_);
    };
    for (long long k = 0; k < iters; k++) {
#pragma unroll
      for (int u = 0; u < 2; u++) {
        rhs(mx, my, mz);
        rhs(qx, qy, qz);
      }
    }
    x[0] = mx + qx, x[1] = my + qy, x[2] = mz + qz;
  } else if constexpr (KIND == 5) {
    for (long long k = 0; k < iters; k++) {
#pragma unroll
      for (int u = 0; u < 16; u++) {
#pragma unroll
        for (int c = 0; c < CH; c++) x[c] = fma(y[(c + u) % CH], z[(c + 3 * u + 1) % CH], x[c]);
      }
    }
  } else {
    for (long long k = 0; k < iters; k++) {
#pragma unroll
      for (int u = 0; u < 16; u++) {
#pragma unroll
        for (int c = 0; c < CH; c++) {
          if constexpr (KIND == 0) x[c] = fma(x[c], a, b);
          else if constexpr (KIND == 1) x[c] = fma(x[c], y[c], z[c]);
          else {
            if (u % 16 == 15) x[c] = x[c] + z[c];
            else if (u % 4 == 3) x[c] = x[c] * y[c];
            else x[c] = fma(x[c], y[c], z[c]);
          }
        }
      }
    }
  }
  double s = 0.0;
#pragma unroll
  for (int c = 0; c < CH; c++) s += x[c];
  if (s == 123.456) out[0] = s;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    out[1] = (double)(clock64() - c0);
    out[2] = (double)(global_ns() - n0);
  }
}

thread_local double g_peak_mhz = 0.0;

}  // namespace
}  // namespace cmtjb

using namespace cmtjb;

extern "C" {

int cmtjb_abi_version(void) { return CMTJB_ABI_VERSION; }
const char *cmtjb_last_error(void) { return g_err.c_str(); }

long long cmtjb_launch_count(void) { return g_launches.load(); }
const char *cmtjb_last_variant(void) { return g_last_variant.c_str(); }

int cmtjb_select_variant(const cmtjb_junction_t *jn, const cmtjb_layer_t *layers, const cmtjb_override_t *overrides,
                         int32_t n_overrides, const cmtjb_run_args_t *args, char *name, int32_t name_size) {
  if (!name || name_size <= 0) return fail(CMTJB_EINVAL, "null argument");
  cmtjb_plan_t plan;
  int rc = make_plan(jn, layers, args, &plan);
  if (rc) return rc;
  rc = check_overrides(jn, overrides, n_overrides);
  if (rc) return rc;
  if (jn->n_layers > 4) {
    std::snprintf(name, name_size, "wide/L%d/F_ALL", jn->n_layers);
    return CMTJB_OK;
  }
  uint32_t per_layer[CMTJB_MAX_LAYERS] = {0};
  uint32_t need = needed_features(jn, layers, overrides, n_overrides, args, plan.effective_solver, per_layer);
  if (args->log_mode == CMTJB_LOG_ENERGY) need = F_ALL;
  const Variant *v = nullptr;
  if ((need & F_DYNI) && plan.effective_solver == CMTJB_RK4 && args->log_mode != CMTJB_LOG_ENERGY &&
      fast_field_eligible(jn, layers, overrides, n_overrides))
    v = pick_variant(jn->n_layers, plan.effective_solver, (need & ~(uint32_t)(F_DYNI | F_OVR)) | F_FAST);
  if (!v) v = pick_variant(jn->n_layers, plan.effective_solver, need, args->log_mode == CMTJB_LOG_ENERGY ? nullptr : per_layer);
  if (!v) return fail(CMTJB_ENOTSUP, "no compiled kernel variant covers this model");
  std::snprintf(name, name_size, "%s", v->name);
  return CMTJB_OK;
}

int cmtjb_release_cache(void) {
  for (int d = 0; d < 64; d++) {
    std::lock_guard<std::mutex> lock(g_slab_mu[d]);
    if (g_slab[d].p) {
      cudaSetDevice(d);
      cudaFree(g_slab[d].p);
      g_slab[d] = Slab();
    }
  }
  return CMTJB_OK;
}

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
  if (jn->n_layers > 4) {  // lane-per-layer kernel (launch_span)
    plan->feature_mask = (int32_t)F_ALL;
    std::snprintf(plan->variant, sizeof(plan->variant), "wide/L%d/F_ALL", jn->n_layers);
    return CMTJB_OK;
  }
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

int cmtjb_run_device_post(const cmtjb_junction_t *jn, const cmtjb_layer_t *layers, const cmtjb_override_t *overrides,
                          int32_t n_overrides, const cmtjb_run_args_t *args, const cmtjb_post_t *post, double *m_state,
                          double *log, int32_t *status, int device, void *stream) {
  if (!post || !jn || !args) return fail(CMTJB_EINVAL, "null argument");
  if (post->n_proj < 0 || post->n_proj > CMTJB_MAX_PROJ) return fail(CMTJB_EINVAL, "n_proj out of range");
  if (post->sd_mode < CMTJB_SD_NONE || post->sd_mode > CMTJB_SD_RY_PARALLEL) return fail(CMTJB_EINVAL, "unknown sd_mode");
  if (post->n_proj > 0 && !post->proj_log) return fail(CMTJB_EINVAL, "proj_log is null");
  if (post->proj_sample0 < 0 || post->proj_n_samples < 0) return fail(CMTJB_EINVAL, "negative projection log extent");
  if (post->sd_mode != CMTJB_SD_NONE && (!post->sd_state || !post->sd_omega))
    return fail(CMTJB_EINVAL, "sd_state / sd_omega is null");
  if (jn->n_layers > CMTJB_POST_MAX_LAYERS) return fail(CMTJB_ENOTSUP, "post-processing supports at most 4 layers");
  if (post->sd_mode == CMTJB_SD_RZ && (jn->mr_mode != CMTJB_MR_CLASSIC || jn->n_layers > 2))
    return fail(CMTJB_EINVAL, "Rz can only be used for 2 layer junctions. Use Rx or Ry instead.");
  if (args->log_mode == CMTJB_LOG_ENERGY) return fail(CMTJB_ENOTSUP, "the energy log cannot be combined with post-processing");
  cmtjb_plan_t plan;
  int rc = make_plan(jn, layers, args, &plan);
  if (rc) return rc;
  if (post->n_proj > 0 && post->proj_n_samples > 0 && post->proj_sample0 + plan.n_samples > post->proj_n_samples)
    return fail(CMTJB_EINVAL, "projection log too small for this call");
  if (plan.effective_solver != CMTJB_RK4)
    return fail(CMTJB_ENOTSUP, "post-processing is fused into the RK4 kernels only (the procedures run the default solver)");
  CUDA_TRY(cudaSetDevice(device));
  return run_device_impl(jn, layers, overrides, n_overrides, args, m_state, log, status, nullptr, nullptr,
                         (cudaStream_t)stream, nullptr, nullptr, post);
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
  const bool trace_host = std::getenv("CMTJB_TRACE_HOST") != nullptr;
  const auto wall0 = std::chrono::steady_clock::now();
  auto wall_ms = [&]() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - wall0).count(); };
  const size_t nm = (size_t)jn->n_layers * 3 * B * sizeof(double);
  const bool want_log = (args->log_mode == CMTJB_LOG_FULL || args->log_mode == CMTJB_LOG_ENERGY) && log && plan.n_samples > 0;
  const size_t nlog = want_log ? (size_t)plan.n_channels * plan.n_samples * B * sizeof(double) : 0;
  // time chunking of large fixed-step logs (see below): at most ~768 MB per chunk and at least 32 chunks (the
  // copy of the LAST chunk is the only one not hidden behind compute), each of >= 1 sample
  int n_chunks = 1;
  if (want_log && !plan.variable_samples && nlog > ((size_t)64 << 20) && plan.n_samples >= 16) {
    const long long by_size = (long long)((nlog + ((size_t)768 << 20) - 1) / ((size_t)768 << 20));
    n_chunks = (int)std::min<long long>(plan.n_samples, std::max<long long>(std::min<long long>(32, plan.n_samples / 4), by_size));
    if (const char *c = std::getenv("CMTJB_HOST_CHUNKS")) n_chunks = (int)std::max<long long>(2, std::min<long long>(std::atoll(c), plan.n_samples));  // tuning knob
  }
  const bool chunked = n_chunks > 1;
  const long long chunk_samples = chunked ? (plan.n_samples + n_chunks - 1) / n_chunks + 1 : plan.n_samples;
  const size_t chunk_elems = (size_t)plan.n_channels * chunk_samples * B;
  // ring of chunk buffers: as many as fit in a quarter of the free device memory (at most 24 GB), at least two.
  // With the whole log in the ring the integration never waits for a copy, so a transient slowdown of the host
  // side (other tenants of the host's memory) only delays the copies; with two buffers it still streams.
  int n_buf = 1;
  if (chunked) {
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) free_b = (size_t)8 << 30;
    const size_t budget = std::min<size_t>(free_b / 4, (size_t)24 << 30);
    n_buf = (int)std::max<size_t>(2, std::min<size_t>((size_t)n_chunks, budget / std::max<size_t>(chunk_elems * sizeof(double), 1)));
    if (const char *c = std::getenv("CMTJB_HOST_RING")) n_buf = std::max(2, std::min(n_chunks, std::atoi(c)));  // tests: force buffer reuse
  }
  const size_t nlog_dev = chunked ? (size_t)n_buf * chunk_elems * sizeof(double) : nlog;
  // carve every device buffer of this call out of the per-device slab (256-byte aligned)
  auto al = [](size_t b) { return (b + 255) & ~(size_t)255; };
  const size_t n_tlog = (plan.variable_samples && want_log && time_log) ? (size_t)plan.n_samples * B * sizeof(double) : 0;
  const size_t n_onef = args->onef_buffer ? (size_t)jn->n_layers * 3 * plan.total_iterations * B * sizeof(double) : 0;
  const size_t n_cust = (args->custom_table && args->n_custom > 0) ? (size_t)3 * plan.total_iterations * args->n_custom * sizeof(double) : 0;
  const size_t total = al(nm) + al(nlog_dev) + 2 * al(B * sizeof(int32_t)) + al(n_tlog) + al(n_onef) + al(n_cust) +
                       (size_t)n_overrides * al(B * sizeof(double)) + 256;
  if (device < 0 || device >= 64) return fail(CMTJB_EINVAL, "device index out of range");
  std::lock_guard<std::mutex> lock(g_slab_mu[device]);
  void *base = nullptr;
  rc = slab_reserve(device, total, &base);
  if (rc) return rc;
  char *cur = (char *)base;
  auto take = [&](size_t bytes) {
    void *p = cur;
    cur += al(bytes);
    return p;
  };
  double *d_m = (double *)take(nm);
  double *d_log = want_log ? (double *)take(nlog_dev) : nullptr;
  int32_t *d_status = (int32_t *)take(B * sizeof(int32_t));
  int32_t *d_ns = plan.variable_samples ? (int32_t *)take(B * sizeof(int32_t)) : nullptr;
  double *d_tlog = n_tlog ? (double *)take(n_tlog) : nullptr;
  double *d_onef = n_onef ? (double *)take(n_onef) : nullptr;
  double *d_cust = n_cust ? (double *)take(n_cust) : nullptr;
  CUDA_TRY(cudaMemcpy(d_m, m_state, nm, cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMemset(d_status, 0, B * sizeof(int32_t)));
  std::vector<cmtjb_override_t> d_ovr(overrides, overrides + n_overrides);
  for (int k = 0; k < n_overrides; k++) {
    double *p = (double *)take(B * sizeof(double));
    CUDA_TRY(cudaMemcpy(p, overrides[k].values, B * sizeof(double), cudaMemcpyHostToDevice));
    d_ovr[k].values = p;
  }
  cmtjb_run_args_t a = *args;
  if (d_onef) {
    CUDA_TRY(cudaMemcpy(d_onef, args->onef_buffer, n_onef, cudaMemcpyHostToDevice));
    a.onef_buffer = d_onef;
  }
  if (d_cust) {
    CUDA_TRY(cudaMemcpy(d_cust, args->custom_table, n_cust, cudaMemcpyHostToDevice));
    a.custom_table = d_cust;
  }
  // the staging copies above ran on the legacy stream; a pageable cudaMemcpy may return before its DMA has
  // landed, and the streams below do not synchronise with the legacy stream
  CUDA_TRY(cudaDeviceSynchronize());
  if (trace_host) std::fprintf(stderr, "[cmtjb] host: inputs staged at %.2f ms\n", wall_ms());
  HostStats hs_stats;
  hs_stats.n_chunks = n_chunks;
  hs_stats.n_ring = n_buf;
  // 1/f^alpha noise generated by the library: once for the whole run when it fits the budget (time chunks then
  // share the buffer), in instance sub-batches by run_device_impl otherwise (the log is then kept resident)
  float *d_gen = nullptr;
  Span gen_span;  // carries the pre-generated buffer into every time chunk
  {
    OnefSlot slots[3 * CMTJB_MAX_LAYERS];
    int n_slots = 0;
    rc = onef_needed(jn, layers, &a, plan, slots, &n_slots);
    if (rc) return rc;
    if (n_slots > 0 && chunked) {
      const size_t need = (size_t)n_slots * (size_t)plan.total_iterations * sizeof(float) * (size_t)B;
      if (need > onef_budget_bytes())
        return fail(CMTJB_ERUNTIME, "1/f^alpha noise: the noise buffers of this batch (" + std::to_string(need) +
                                        " bytes) and a log of " + std::to_string(nlog) +
                                        " bytes do not fit the device together; run fewer instances per call or decimate the log");
      CUDA_TRY(cudaMalloc((void **)&d_gen, need));
      rc = onef_generate(device, nullptr, plan.total_iterations, slots, n_slots, a.seed, (unsigned long long)a.instance_offset, B, d_gen);
      if (rc == CMTJB_OK && cudaDeviceSynchronize() != cudaSuccess) rc = fail(CMTJB_ECUDA, "1/f^alpha noise generation failed");
      if (rc) {
        cudaFree(d_gen);
        return rc;
      }
      gen_span.onef_gen = d_gen;
      gen_span.onef_i0 = 0;
      gen_span.onef_B = B;
    }
  }
  if (!chunked) {
    const double t_staged = wall_ms();
    rc = run_device_impl(jn, layers, d_ovr.data(), n_overrides, &a, d_m, d_log, d_status, d_ns, d_tlog, nullptr, nullptr);
    if (rc) return rc;
    CUDA_TRY(cudaDeviceSynchronize());
    hs_stats.integrate_ms = wall_ms() - t_staged;
    if (want_log) CUDA_TRY(cudaMemcpy(log, d_log, nlog, cudaMemcpyDeviceToHost));
  } else {
    // Large fixed-step logs: the run is cut into time chunks; chunk c integrates into buffer c % n_buf of a ring
    // while the copy engine streams the finished chunks to the host (effective when `log` is pinned memory).
    // Device memory for the log is bounded by the ring whatever the length of the run, so a log larger than HBM
    // (full-rate PIMM logs: 56 B per junction-step) streams through.
    cudaStream_t sc = nullptr, sd = nullptr;
    cudaError_t es = cudaStreamCreateWithFlags(&sc, cudaStreamNonBlocking);
    if (es == cudaSuccess) es = cudaStreamCreateWithFlags(&sd, cudaStreamNonBlocking);
    if (es != cudaSuccess) {
      if (sc) cudaStreamDestroy(sc);
      if (d_gen) cudaFree(d_gen);
      return cuda_fail(es, "cudaStreamCreateWithFlags");
    }
    const long long we = plan.write_every, S0 = a.start_iteration;
    const long long i0 = ((S0 + we - 1) / we) * we;  // iteration of this run's sample 0
    // per chunk: integration started / finished (compute stream), copy finished (copy stream)
    std::vector<cudaEvent_t> ev_begin(n_chunks, nullptr), ev_end(n_chunks, nullptr), ev_copied(n_chunks, nullptr);
    for (int c = 0; c < n_chunks; c++) {
      cudaEventCreate(&ev_begin[c]);
      cudaEventCreate(&ev_end[c]);
      cudaEventCreate(&ev_copied[c]);
    }
    int enqueued = 0;
    for (int c = 0; c < n_chunks && rc == CMTJB_OK; c++) {
      const long long s0 = plan.n_samples * c / n_chunks, s1 = plan.n_samples * (c + 1) / n_chunks;
      const int bi = c % n_buf;
      double *buf = d_log + (size_t)bi * chunk_elems;
      Span sp = gen_span;
      sp.it_begin = (c == 0) ? S0 : i0 + s0 * we;
      sp.it_end = (c == n_chunks - 1) ? plan.total_iterations : i0 + s1 * we;
      sp.first_sample = 0;
      sp.n_samples = chunk_samples;
      if (c >= n_buf) cudaStreamWaitEvent(sc, ev_copied[c - n_buf], 0);  // buffer bi has been drained to the host
      cudaEventRecord(ev_begin[c], sc);
      rc = run_device_impl(jn, layers, d_ovr.data(), n_overrides, &a, d_m, buf, d_status, d_ns, d_tlog, sc, nullptr, &sp);
      if (rc) break;
      cudaEventRecord(ev_end[c], sc);
      cudaStreamWaitEvent(sd, ev_end[c], 0);
      for (int ch = 0; ch < plan.n_channels; ch++)
        cudaMemcpyAsync(log + ((size_t)ch * plan.n_samples + s0) * B, buf + (size_t)ch * chunk_samples * B,
                        (size_t)(s1 - s0) * B * sizeof(double), cudaMemcpyDeviceToHost, sd);
      cudaEventRecord(ev_copied[c], sd);
      enqueued = c + 1;
    }
    if (trace_host) std::fprintf(stderr, "[cmtjb] host: all chunks enqueued at %.2f ms\n", wall_ms());
    cudaError_t e1 = cudaStreamSynchronize(sc), e2 = cudaStreamSynchronize(sd);
    if (trace_host) std::fprintf(stderr, "[cmtjb] host: streams drained at %.2f ms\n", wall_ms());
    if (rc == CMTJB_OK && e1 == cudaSuccess && e2 == cudaSuccess) {
      // where the time went: integration, gaps on the compute stream (waiting for a ring buffer to drain), and the
      // copy engine -- busy, or idle waiting for the next chunk
      float ms = 0;
      double prev_end = 0, prev_copied = 0;
      for (int c = 0; c < enqueued; c++) {
        cudaEventElapsedTime(&ms, ev_begin[0], ev_begin[c]);
        const double tb = ms;
        cudaEventElapsedTime(&ms, ev_begin[0], ev_end[c]);
        const double te = ms;
        cudaEventElapsedTime(&ms, ev_begin[0], ev_copied[c]);
        const double tc = ms;
        hs_stats.integrate_ms += te - tb;
        if (c > 0) hs_stats.ring_wait_ms += std::max(0.0, tb - prev_end);
        hs_stats.copy_ms += tc - std::max(te, prev_copied);
        if (trace_host) std::fprintf(stderr, "[cmtjb] chunk %d: integration %.1f..%.1f ms, copied %.1f ms\n", c, tb, te, tc);
        prev_end = te;
        prev_copied = tc;
        hs_stats.pipeline_ms = tc;
        hs_stats.tail_copy_ms = tc - te;
      }
    }
    for (int c = 0; c < n_chunks; c++) {
      cudaEventDestroy(ev_begin[c]);
      cudaEventDestroy(ev_end[c]);
      cudaEventDestroy(ev_copied[c]);
    }
    cudaStreamDestroy(sc);
    cudaStreamDestroy(sd);
    if (d_gen) cudaFree(d_gen);
    if (rc) return rc;
    if (e1 != cudaSuccess) return cuda_fail(e1, "chunked run (compute stream)");
    if (e2 != cudaSuccess) return cuda_fail(e2, "chunked run (copy stream)");
  }
  hs_stats.log_bytes = (double)nlog;
  hs_stats.total_ms = wall_ms();
  g_host_stats = hs_stats;
  CUDA_TRY(cudaMemcpy(m_state, d_m, nm, cudaMemcpyDeviceToHost));
  if (d_tlog) CUDA_TRY(cudaMemcpy(time_log, d_tlog, (size_t)plan.n_samples * B * sizeof(double), cudaMemcpyDeviceToHost));
  if (d_ns && n_samples_out) CUDA_TRY(cudaMemcpy(n_samples_out, d_ns, B * sizeof(int32_t), cudaMemcpyDeviceToHost));
  std::vector<int32_t> st_local;
  int32_t *st = status;
  if (!st) {
    st_local.resize(B);
    st = st_local.data();
  }
  CUDA_TRY(cudaMemcpy(st, d_status, B * sizeof(int32_t), cudaMemcpyDeviceToHost));
  if (trace_host) std::fprintf(stderr, "[cmtjb] host: done at %.2f ms\n", wall_ms());
  for (long long i = 0; i < B; i++)
    if (st[i]) return fail(CMTJB_ENAN, "NAN magnetisation");
  return CMTJB_OK;
}

int cmtjb_host_stats(double *out, int32_t n) {
  if (!out || n <= 0) return fail(CMTJB_EINVAL, "null argument");
  const HostStats &h = g_host_stats;
  const double v[CMTJB_HOST_NSTATS] = {h.total_ms,     h.integrate_ms, h.ring_wait_ms,     h.copy_ms,         h.tail_copy_ms,
                                       h.pipeline_ms,  h.log_bytes,    (double)h.n_chunks, (double)h.n_ring};
  for (int k = 0; k < n && k < CMTJB_HOST_NSTATS; k++) out[k] = v[k];
  return CMTJB_OK;
}

double cmtjb_fp64_peak_sm_mhz(void) { return g_peak_mhz; }

int cmtjb_fp64_pipe_probe(int device, int kind, int warps_per_subpartition, int64_t iters, double *issue_slot_frac,
                          double *sm_mhz) {
  if (kind < 0 || kind > 5 || warps_per_subpartition < 1 || warps_per_subpartition > 16 || iters <= 0 || !issue_slot_frac)
    return fail(CMTJB_EINVAL, "bad argument");
  CUDA_TRY(cudaSetDevice(device));
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, device));
  constexpr int CH = 8;
  double h_seed[3 * CH];
  for (int c = 0; c < CH; c++) {
    h_seed[c] = 0.5 + 0.01 * c;
    h_seed[CH + c] = 0.999999 - 1e-7 * c;
    h_seed[2 * CH + c] = 1e-7 * (c + 1);
  }
  if (kind == 3 || kind == 4) {  // a slowly precessing unit vector: field 1, damping 0.01, h = 1e-3
    const double m[3] = {0.6, 0.0, 0.8}, nb[3] = {0.0, 0.6, 0.8}, H[3] = {0.0, 0.0, 1.0};
    for (int c = 0; c < 3; c++) h_seed[c] = m[c], h_seed[CH + c] = nb[c], h_seed[2 * CH + c] = H[c];
    h_seed[CH + 3] = -0.1, h_seed[CH + 4] = -0.1, h_seed[CH + 5] = -1.0;
    h_seed[2 * CH + 3] = 1e-3, h_seed[2 * CH + 4] = 1e-2, h_seed[2 * CH + 5] = 0.01, h_seed[2 * CH + 6] = 1e-3;
  }
  double *d = nullptr;
  CUDA_TRY(cudaMalloc(&d, (4 + 3 * CH) * sizeof(double)));
  cudaMemcpy(d + 4, h_seed, sizeof(h_seed), cudaMemcpyHostToDevice);
  const int blocks = prop.multiProcessorCount * warps_per_subpartition;
  auto launch = [&](long long n) {
    switch (kind) {
    case 0: fp64_probe_kernel<0, CH><<<blocks, 128>>>(d, d + 4, n, 0.999999, 1e-7); break;
    case 1: fp64_probe_kernel<1, CH><<<blocks, 128>>>(d, d + 4, n, 0.0, 0.0); break;
    case 2: fp64_probe_kernel<2, CH><<<blocks, 128>>>(d, d + 4, n, 0.0, 0.0); break;
    case 3: fp64_probe_kernel<3, CH><<<blocks, 128>>>(d, d + 4, n, 0.0, 0.0); break;
    case 4: fp64_probe_kernel<4, CH><<<blocks, 128>>>(d, d + 4, n, 0.0, 0.0); break;
    default: fp64_probe_kernel<5, CH><<<blocks, 128>>>(d, d + 4, n, 0.0, 0.0); break;
    }
  };
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  launch(iters / 8 + 1);
  cudaEventRecord(e0);
  launch(iters);
  cudaEventRecord(e1);
  cudaError_t e = cudaEventSynchronize(e1);
  float t_ms = 0;
  cudaEventElapsedTime(&t_ms, e0, e1);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  double probe[4] = {0, 0, 0, 0};
  if (e == cudaSuccess) e = cudaMemcpy(probe, d, sizeof(probe), cudaMemcpyDeviceToHost);
  cudaFree(d);
  if (e != cudaSuccess) return cuda_fail(e, "fp64_probe_kernel");
  // the kernel's duration (events; the warp scheduler favours the oldest warp, so one block's own cycle count says nothing
  // about the others) at the clock block 0 saw (cycle counter against the global nanosecond timer)
  const double mhz = probe[2] > 0 ? probe[1] / probe[2] * 1e3 : 0.0;
  const double cycles = (double)t_ms * 1e-3 * mhz * 1e6;
  const double per_thread = (kind == 3 || kind == 4) ? 28.0 * 4.0 * (double)iters : 16.0 * CH * (double)iters;
  const double per_sm = per_thread * 128.0 * warps_per_subpartition;  // thread-instructions on one SM
  *issue_slot_frac = cycles > 0 ? per_sm / (64.0 * cycles) : 0.0;
  if (sm_mhz) *sm_mhz = mhz;
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
  L.drivers[CMTJB_SLOT_K1] = *driver;
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

int cmtjb_philox_host(const uint32_t *ctr, const uint32_t *key, uint32_t *out, int device) {
  if (!ctr || !key || !out) return fail(CMTJB_EINVAL, "null argument");
  CUDA_TRY(cudaSetDevice(device));
  uint32_t *d = nullptr;
  CUDA_TRY(cudaMalloc(&d, 16));
  philox_kernel<<<1, 1>>>(make_uint4(ctr[0], ctr[1], ctr[2], ctr[3]), make_uint2(key[0], key[1]), d);
  cudaError_t e = cudaMemcpy(out, d, 16, cudaMemcpyDeviceToHost);
  cudaFree(d);
  if (e != cudaSuccess) return cuda_fail(e, "philox_kernel");
  return CMTJB_OK;
}

int cmtjb_normals_host(uint64_t seed, int64_t inst0, int64_t n_inst, int32_t layer, int64_t step0, int64_t n_steps,
                       double *out, int device) {
  return cmtjb_normals_host_ex(seed, inst0, n_inst, layer, step0, n_steps, 0, out, device);
}

int cmtjb_normals_host_ex(uint64_t seed, int64_t inst0, int64_t n_inst, int32_t layer, int64_t step0, int64_t n_steps,
                          int32_t thermal_normals, double *out, int device) {
  if (!out || n_inst <= 0 || n_steps <= 0) return fail(CMTJB_EINVAL, "bad argument");
  CUDA_TRY(cudaSetDevice(device));
  const long long n = n_inst * n_steps;
  double *d = nullptr;
  CUDA_TRY(cudaMalloc(&d, 3 * n * sizeof(double)));
  normals_kernel<<<(unsigned)((n + 255) / 256), 256>>>(seed, inst0, n_inst, (unsigned)layer, step0, n_steps, d, thermal_normals == 1 ? 1 : 0);
  cudaError_t e = cudaMemcpy(out, d, 3 * n * sizeof(double), cudaMemcpyDeviceToHost);
  cudaFree(d);
  if (e != cudaSuccess) return cuda_fail(e, "normals_kernel");
  return CMTJB_OK;
}

int cmtjb_fp64_peak(int device, int64_t iters, double *flops, double *ms) {
  CUDA_TRY(cudaSetDevice(device));
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, device));
  double *d = nullptr;
  CUDA_TRY(cudaMalloc(&d, 4 * sizeof(double)));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  // the FP64 pipe only saturates with enough independent DFMAs per sub-partition: 8 or 16 chains per thread at 8 or
  // 16 warps per sub-partition; the best of the four shapes is the measured peak
  double best = 0.0, best_ms = 0.0;
  cudaError_t e = cudaSuccess;
  for (int shape = 0; shape < 4 && e == cudaSuccess; shape++) {
    const int chains = (shape & 1) ? 16 : 8, blocks_per_sm = (shape & 2) ? 8 : 4;
    const int blocks = prop.multiProcessorCount * blocks_per_sm, threads = 256;
    const long long it = std::max<long long>(1, iters * 8 / chains);
    auto launch = [&](long long n) {
      if (chains == 8) fp64_peak_kernel<8><<<blocks, threads>>>(d, n, 0.999999, 1e-7);
      else fp64_peak_kernel<16><<<blocks, threads>>>(d, n, 0.999999, 1e-7);
    };
    launch(it / 8 + 1);  // warm-up
    cudaEventRecord(e0);
    launch(it);
    cudaEventRecord(e1);
    e = cudaEventSynchronize(e1);
    float t = 0;
    cudaEventElapsedTime(&t, e0, e1);
    const double n_fma = (double)blocks * threads * (double)it * 8.0 * chains;
    const double f = 2.0 * n_fma / (t * 1e-3);
    if (e == cudaSuccess && f > best) {
      best = f;
      best_ms = t;
      double probe[4] = {0, 0, 0, 0};
      if (cudaMemcpy(probe, d, sizeof(probe), cudaMemcpyDeviceToHost) == cudaSuccess && probe[2] > 0) g_peak_mhz = probe[1] / probe[2] * 1e3;
    }
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d);
  if (e != cudaSuccess) return cuda_fail(e, "fp64_peak_kernel");
  if (flops) *flops = best;
  if (ms) *ms = best_ms;
  return CMTJB_OK;
}

} /* extern "C" */
