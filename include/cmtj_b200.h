/*
 * cmtj_b200.h -- C-ABI of the B200-native (s)LLGS junction integrator.
 *
 * This is the drop-in boundary that replaces the pybind11 translation unit of the
 * reference (python/cmtj.cpp:291-355 `Junction`, :358-426 `stack`, :429-445 `reservoir`)
 * for ONE path: the per-timestep (s)LLGS integration of multilayer junctions, batched
 * across parameter-sweep instances.  Plain C types only: no C++ types, no exceptions,
 * no Python.h, no torch types.  Every entry point returns 0 on success or a negative
 * CMTJB_E* code; cmtjb_last_error() returns a thread-local message whose wording follows
 * the reference's exception texts so that a Python facade can re-raise the same
 * ValueError / RuntimeError the reference's tests match on.
 *
 * Data model
 * ----------
 *  - A *junction model* is a `cmtjb_junction_t` header plus `n_layers` `cmtjb_layer_t`
 *    records.  It carries the RAW physical description, exactly what the reference's
 *    `Layer<T>` / `Junction<T>` members hold (core/junction.hpp:145-260, :1139-1156);
 *    all hoisting of reciprocals / derived coefficients happens inside the library.
 *  - Every time-dependent quantity of a layer is a *driver slot* (`cmtjb_slot`), holding
 *    the reference's `ScalarDriver<T>` fields (core/drivers.hpp:31-72).
 *  - A *batch* is `n_instances` copies of the model.  Instances differ by their
 *    magnetisation state and by optional per-instance *overrides* of driver fields
 *    (e.g. a different external field vector or current amplitude per instance).
 *  - State and logs are SoA with the instance index fastest:
 *        m_state[(layer*3 + comp) * n_instances + inst]
 *        log[(channel * n_samples + sample) * n_instances + inst]
 *    Channels: for each layer mx,my,mz; then R (CLASSIC) or Rx,Ry,Rz (STRIP)
 *    (core/junction.hpp:1614-1635).
 */
#ifndef CMTJ_B200_H_
#define CMTJ_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CMTJB_ABI_VERSION 6
#define CMTJB_MAX_LAYERS 8

/* ---- error codes --------------------------------------------------------------- */
enum cmtjb_status {
  CMTJB_OK = 0,
  CMTJB_EINVAL = -1,      /* reference: std::invalid_argument -> ValueError   */
  CMTJB_ERUNTIME = -2,    /* reference: std::runtime_error    -> RuntimeError */
  CMTJB_ENOTSUP = -3,     /* valid in the reference, not supported by this build */
  CMTJB_ECUDA = -4,       /* CUDA runtime failure (no device, launch error...) */
  CMTJB_ENAN = -5         /* "NAN magnetisation" (core/junction.hpp:1037-1039) */
};

/* ---- drivers: core/drivers.hpp:18-29 (UpdateType) ------------------------------- */
enum cmtjb_update {
  CMTJB_UPD_CONSTANT = 0,
  CMTJB_UPD_PULSE = 1,
  CMTJB_UPD_SINE = 2,
  CMTJB_UPD_STEP = 3,
  CMTJB_UPD_POSINE = 4,
  CMTJB_UPD_HALFSINE = 5,
  CMTJB_UPD_TRAPEZOID = 6,
  CMTJB_UPD_GAUSSIMPULSE = 7,
  CMTJB_UPD_GAUSSSTEP = 8,
  CMTJB_UPD_CUSTOM = 100 /* host callback (drivers.hpp:294-303), pre-evaluated by the caller: see custom_table */
};

/* core/drivers.hpp:34-36 (Driver fields) + :68-69 (edgeTime, steadyTime) */
typedef struct cmtjb_driver {
  int32_t update;
  int32_t custom_column; /* CMTJB_UPD_CUSTOM: column of cmtjb_run_args_t::custom_table; else unused */
  double constantValue, amplitude, frequency, phase, period, cycle, timeStart, timeStop,
      edgeTime, steadyTime;
} cmtjb_driver_t;

/* field selector for per-instance overrides (index into the doubles above) */
enum cmtjb_field {
  CMTJB_FIELD_CONSTANT = 0,
  CMTJB_FIELD_AMPLITUDE = 1,
  CMTJB_FIELD_FREQUENCY = 2,
  CMTJB_FIELD_PHASE = 3,
  CMTJB_FIELD_PERIOD = 4,
  CMTJB_FIELD_CYCLE = 5,
  CMTJB_FIELD_TIMESTART = 6,
  CMTJB_FIELD_TIMESTOP = 7,
  CMTJB_FIELD_EDGETIME = 8,
  CMTJB_FIELD_STEADYTIME = 9
};

/* driver slots of one layer: the ScalarDriver / AxialDriver members of Layer<T>
 * (core/junction.hpp:147-163) */
enum cmtjb_slot {
  CMTJB_SLOT_HEXT_X = 0, CMTJB_SLOT_HEXT_Y, CMTJB_SLOT_HEXT_Z,       /* externalFieldDriver */
  CMTJB_SLOT_HOE_X = 3, CMTJB_SLOT_HOE_Y, CMTJB_SLOT_HOE_Z,          /* HoeDriver */
  CMTJB_SLOT_HDMI_X = 6, CMTJB_SLOT_HDMI_Y, CMTJB_SLOT_HDMI_Z,       /* HdmiDriver */
  CMTJB_SLOT_HRES_X = 9, CMTJB_SLOT_HRES_Y, CMTJB_SLOT_HRES_Z,       /* HreservedInteractionFieldDriver */
  CMTJB_SLOT_K1 = 12,                                                 /* anisotropyDriver */
  CMTJB_SLOT_K2 = 13,                                                 /* secondOrderAnisotropyDriver */
  CMTJB_SLOT_IEC_TOP = 14, CMTJB_SLOT_IEC_BOTTOM = 15,
  CMTJB_SLOT_IECQ_TOP = 16, CMTJB_SLOT_IECQ_BOTTOM = 17,
  CMTJB_SLOT_IDMI_TOP_X = 18, CMTJB_SLOT_IDMI_TOP_Y, CMTJB_SLOT_IDMI_TOP_Z,
  CMTJB_SLOT_IDMI_BOTTOM_X = 21, CMTJB_SLOT_IDMI_BOTTOM_Y, CMTJB_SLOT_IDMI_BOTTOM_Z,
  CMTJB_SLOT_CURRENT = 24,
  CMTJB_SLOT_FL = 25, CMTJB_SLOT_DL = 26, CMTJB_SLOT_FL2 = 27, CMTJB_SLOT_DL2 = 28,
  CMTJB_SLOT_TEMPERATURE = 29,
  CMTJB_NSLOTS = 30
};

/* core/junction.hpp:141 */
enum cmtjb_reference { CMTJB_REF_NONE = 0, CMTJB_REF_FIXED = 1, CMTJB_REF_TOP = 2, CMTJB_REF_BOTTOM = 3 };
/* core/junction.hpp:143 */
enum cmtjb_solver { CMTJB_EULER_HEUN = 0, CMTJB_RK4 = 1, CMTJB_DORMAND_PRINCE = 2, CMTJB_HEUN = 3 };
/* core/junction.hpp:1144 */
enum cmtjb_mr_mode { CMTJB_MR_NONE = 0, CMTJB_MR_CLASSIC = 1, CMTJB_MR_STRIP = 2 };
/* core/cvector.hpp:11 */
enum cmtjb_axis { CMTJB_AXIS_X = 0, CMTJB_AXIS_Y = 1, CMTJB_AXIS_Z = 2, CMTJB_AXIS_ALL = 3, CMTJB_AXIS_NONE = 4 };

/* One magnetic layer: the data members of Layer<T> (core/junction.hpp:145-260). */
typedef struct cmtjb_layer {
  double mag[3];  /* as stored by the reference (NOT normalised by the ctor, see SURVEY 8a note 10) */
  double anis[3];
  double Ms, thickness, cellSurface, damping;
  double demagTensor[9], dipoleBottom[9], dipoleTop[9]; /* row-major */
  double fieldLikeTorque, dampingLikeTorque;            /* static SOT (createSOTLayer) */
  double SlonczewskiSpacerLayerParameter, beta, kappa, spinPolarisation; /* STT */
  double referenceLayer[3], secondaryReferenceLayer[3];
  double Rx0, Ry0, AMR_X, AMR_Y, SMR_X, SMR_Y, AHE;     /* STRIP magnetoresistance, per layer */
  double noiseAlpha, noiseStd, noiseScale;              /* setAlphaNoise (core/junction.hpp:406-414) */
  /* AdaptiveIntegrationParams (core/junction.hpp:125-139) */
  double abs_tol, rel_tol, max_factor, min_factor, safety_factor, ki, kp, kd,
      prev_error_ratio, integral_error;
  int32_t referenceType;
  int32_t includeSTT, includeSOT, dynamicSOT, alternativeSTT;
  int32_t temperatureSet;
  int32_t noiseAxis;
  int32_t use_pid_control;
  cmtjb_driver_t drivers[CMTJB_NSLOTS];
} cmtjb_layer_t;

/* Junction-level data (core/junction.hpp:1146-1155) + a snapshot of the mutable
 * PhysicalConstants (core/constants.hpp:39-44), taken at call time. */
typedef struct cmtjb_junction {
  int32_t n_layers;
  int32_t mr_mode;
  double Rp, Rap;
  double MAGNETIC_PERMEABILITY, GYRO, TtoAm, HBAR, ELECTRON_CHARGE, BOLTZMANN_CONST;
} cmtjb_junction_t;

/* Per-instance override of one driver field: values[inst] replaces
 * layers[layer].drivers[slot].<field> for instance `inst`; layer = -1 applies the same
 * per-instance value to every layer (the reference's layerID == "all"). */
typedef struct cmtjb_override {
  int32_t layer;
  int32_t slot;
  int32_t field;
  int32_t reserved;
  const double *values; /* [n_instances]; host or device pointer according to the entry point */
} cmtjb_override_t;

/* Arguments of Junction::runSimulation (core/junction.hpp:1948-1950) for a batch. */
typedef struct cmtjb_run_args {
  double totalTime, timeStep, writeFrequency;
  int32_t solverMode;       /* requested cmtjb_solver; may be overridden like getSolver (:1891-1931) */
  int32_t log_mode;         /* CMTJB_LOG_* */
  uint64_t seed;            /* Philox key for thermal noise */
  int64_t n_instances;
  int64_t instance_offset;  /* global index of instance 0 (RNG stream id; multi-GPU shards) */
  int64_t start_iteration;  /* first step index; >0 continues a run in time chunks */
  /* 1/f^alpha noise of layers with noiseScale != 0 (setAlphaNoise, core/junction.hpp:406-414; one fresh
   * VectorAlphaNoise of bufferSize = total_iterations per layer and run, getSolver :1905-1916).
   * NULL (the normal case): the library generates the buffers on the device -- one independent series per
   * (instance, layer, axis): Philox4x32-10 white noise keyed by (seed, instance_offset + instance, layer,
   * axis, sample), Kasdin's filter in float32 exactly as BufferedAlphaNoise::fillBuffer does it
   * (core/noise.hpp:190-227, 2n-point complex FFT, one-sided spectrum with both end bins halved), then
   * tickVector's normalisation and scale (noise.hpp:291-301) inside the time loop.  Batches whose buffers
   * exceed the free device memory are integrated in sub-batches of instances; a run that cannot fit even
   * 32 instances fails with CMTJB_ERUNTIME and the byte counts.
   * Non-NULL: the caller's own tickVector values, already scaled, double [n_layers*3][total_iterations]
   * [n_instances] (host or device pointer according to the entry point; sweep entry points only). */
  const double *onef_buffer;
  /* CMTJB_UPD_CUSTOM drivers (ScalarDriver::getCustomDriver, drivers.hpp:247-260, :294-303).  A callback
   * cannot cross this boundary, but a fixed-step solver only ever asks a driver for its value at the
   * three stage times of each step, which are the same for every instance.  The caller evaluates each
   * callback there and passes [3 * total_iterations][n_custom] values: row 3*i + s is stage s of step i,
   * at the reference's own time expressions  i*dt,  i*dt + 0.5*dt,  i*dt + dt  (junction.hpp:1974,
   * :1020-1033; Heun reads stages 0 and 2, Euler-Heun stage 0).  As in the reference the callback's
   * value IS the driver's value (constantValue and amplitude are ignored).  Host or device pointer
   * according to the entry point; NULL when the model has no custom driver.  Not available for
   * Dormand-Prince, Stack and GroupInteraction (CMTJB_ENOTSUP). */
  const double *custom_table;
  int64_t n_custom;
  /* Thermal (Langevin) normals, replacing std::normal_distribution<double> over mt19937 (core/junction.hpp:174-175,
   * :1111-1119).  0 (default): Philox4x32-10 + Box-Muller in float32 on the SFU -- absolute error 2^-21 per normal, radius
   * tail ending at 6.7 sigma (a 32-bit uniform); 0.75 hashes per layer-step.  1: the same generator with 53-bit uniforms
   * and double-precision log / sqrt / sincospi -- relative error ~1e-16, tail to 8.6 sigma; two hashes and ~300 more
   * instructions per layer-step.  Both are pure functions of (seed, global instance, layer, step); the two streams are
   * different numbers. */
  int32_t thermal_normals;
  int32_t reserved0;
} cmtjb_run_args_t;

enum cmtjb_log_mode {
  CMTJB_LOG_FULL = 0,   /* every channel at every written sample */
  CMTJB_LOG_NONE = 1,   /* final state only: m_state is updated, no samples written */
  /* calculateEnergies = true (core/junction.hpp:1584-1613): after the LOG_FULL channels, per layer
   * K, K2, Jbottom, Jtop, I, Hext[3], Hiec[3], Hanis[3], Hdemag[3], Hth[3] and, for layers with
   * includeSOT, Hfl[3], Hdl[3] -- the values the reference caches in its LAST right-hand-side
   * evaluation of the step (RK4: stage 4 at t+dt; Heun: corrector; Euler-Heun: the evaluation at t). */
  CMTJB_LOG_ENERGY = 2
};
#define CMTJB_ENERGY_CHANNELS 20      /* per layer */
#define CMTJB_ENERGY_CHANNELS_SOT 26  /* per layer with includeSOT */

/* What a run will do; filled by cmtjb_plan() so the caller can size buffers. */
typedef struct cmtjb_plan {
  int32_t effective_solver;  /* after the thermal / 1/f-noise override */
  int32_t solver_overridden; /* 1 if the reference would print its [WARNING] */
  int32_t n_layers;
  int32_t n_channels;        /* 3*n_layers + {0,1,3} */
  int64_t total_iterations;  /* (unsigned)(totalTime/timeStep) */
  int64_t write_every;       /* (unsigned)(writeFrequency/timeStep) */
  int64_t n_samples;         /* fixed-step: ceil(total_iterations/write_every); DP: capacity */
  int32_t variable_samples;  /* 1 for Dormand-Prince: per-instance n_samples_out/time_log used */
  int32_t feature_mask;      /* kernel variant chosen (diagnostic) */
  char variant[64];
} cmtjb_plan_t;

/* ---- entry points ---------------------------------------------------------------- */

int cmtjb_abi_version(void);
const char *cmtjb_last_error(void);
/* number of CUDA devices visible to the library (0 when none) */
int cmtjb_device_count(void);

/* Validates the model the way the reference's constructors and runSimulation do
 * (core/junction.hpp:1164-1243, :1952-1955, :1111-1114) and fills `plan`. */
int cmtjb_plan(const cmtjb_junction_t *jn, const cmtjb_layer_t *layers,
               const cmtjb_run_args_t *args, cmtjb_plan_t *plan);

/* Device entry: replaces Junction::runSimulation (core/junction.hpp:1948-2056) for a batch.
 * All pointers below are DEVICE pointers on `device`; the launch is asynchronous on
 * `stream` (a cudaStream_t, NULL = legacy default stream).
 *   m_state        in/out [n_layers*3][n_instances]
 *   log            out    [n_channels][n_samples][n_instances]   (may be NULL for LOG_NONE)
 *   status         out    [n_instances] int32: 0 ok, 1 NaN magnetisation
 *   n_samples_out  out    [n_instances] int32 (Dormand-Prince only, else may be NULL)
 *   time_log       out    [n_samples][n_instances]  (Dormand-Prince only, else may be NULL)
 *   dt_state       in/out [n_instances] current adaptive step (DP only; NULL = start at timeStep)
 */
int cmtjb_run_device(const cmtjb_junction_t *jn, const cmtjb_layer_t *layers,
                     const cmtjb_override_t *overrides, int32_t n_overrides,
                     const cmtjb_run_args_t *args, double *m_state, double *log,
                     int32_t *status, int32_t *n_samples_out, double *time_log,
                     int device, void *stream);

/* Host entry: same contract with HOST pointers; the library stages inputs to the device,
 * runs, and copies state/logs/status back before returning (synchronous).  This is the
 * call `Junction.runSimulation` of the facade makes, and what `e2e` in bench.py times. */
int cmtjb_run_host(const cmtjb_junction_t *jn, const cmtjb_layer_t *layers,
                   const cmtjb_override_t *overrides, int32_t n_overrides,
                   const cmtjb_run_args_t *args, double *m_state, double *log,
                   int32_t *status, int32_t *n_samples_out, double *time_log, int device);

/* ---- fused post-processing of the callers of runSimulation ------------------
 *
 * The reference's sweep procedures (cmtj/utils/procedures.py: PIMM_procedure :42-204,
 * VSD_procedure :207-331) log every layer's m at every step, pull the log through pybind11 as
 * Python lists and reduce it with numpy/scipy.  cmtjb_run_device_post runs the same integration
 * and performs those reductions inside the time loop, so the full-rate log never exists:
 *
 *  - projection channels (PIMM, procedures.py:142-150, :171-181): at every logged sample
 *        proj_log[k][proj_sample0 + sample][inst] = sum_{l,c} proj[k][3*l + c] * m_{l,c}
 *    e.g. one row holding thickness_l*Ms_l/sum(thickness*Ms) on the extraction component gives the
 *    `mixed_sum` signal whose FFT is the PIMM spectrum; one row per layer gives the per-layer spectra.
 *
 *  - spin-diode voltage (VSD, procedures.py:327-329 + cmtj/utils/resistance.py:12-22):
 *        SD[n] = -sin(sd_omega[inst] * t_n) * R[n],   t_n = n*timeStep (the sample's time label)
 *        y = lfilter(sd_b, sd_a, SD)   order-3 IIR, transposed direct form II in scipy's operation
 *                                      order (scipy/signal/_lfilter.c.in)
 *        sd_state[3][inst] += sum_n y[n]    (the caller divides by the sample count: np.mean)
 *    R[n] is the junction's own magnetoresistance channel (CMTJB_SD_RZ, junction.hpp:1865-1867) or
 *    Rx / Ry of calculate_resistance_series / _parallel (resistance.py:25-66, :96-149) from the
 *    per-layer parameters `res`.
 *
 * Requires a fixed-step RK4 run (the procedures always request the default solver) and at most
 * CMTJB_POST_MAX_LAYERS layers.  A LOG_FULL `log` may be requested in the same call. */
#define CMTJB_MAX_PROJ 4
#define CMTJB_POST_MAX_LAYERS 4

enum cmtjb_sd_mode {
  CMTJB_SD_NONE = 0,
  CMTJB_SD_RZ = 1,          /* R / R_<id0>_<id1> channel of the junction (MR_CLASSIC) */
  CMTJB_SD_RX_SERIES = 2,   /* calculate_resistance_series(...)[0] */
  CMTJB_SD_RY_SERIES = 3,   /* calculate_resistance_series(...)[1] */
  CMTJB_SD_RX_PARALLEL = 4, /* calculate_resistance_parallel(...)[0] */
  CMTJB_SD_RY_PARALLEL = 5  /* calculate_resistance_parallel(...)[1] */
};

/* per-layer ResistanceParameters (procedures.py:16-26): Rxx0, Rxy0, Ramr, Rahe, Rsmr, w/l */
enum cmtjb_res_field { CMTJB_RES_RXX0 = 0, CMTJB_RES_RXY0, CMTJB_RES_AMR, CMTJB_RES_AHE, CMTJB_RES_SMR, CMTJB_RES_W_OVER_L, CMTJB_RES_NFIELDS };

typedef struct cmtjb_post {
  int32_t n_proj;   /* 0..CMTJB_MAX_PROJ projection channels */
  int32_t sd_mode;  /* cmtjb_sd_mode */
  double proj[CMTJB_MAX_PROJ][3 * CMTJB_POST_MAX_LAYERS];
  double res[CMTJB_POST_MAX_LAYERS][CMTJB_RES_NFIELDS];
  double sd_b[4], sd_a[4];  /* scipy.signal.butter(3, ...) coefficients, a[0] == 1 */
  const double *sd_omega;   /* DEVICE [n_instances]: 2*pi*frequency of the mixing current */
  double *proj_log;         /* DEVICE out [n_proj][proj_n_samples][n_instances]; this call writes rows
                               proj_sample0 .. proj_sample0 + (samples of this call) - 1 */
  int64_t proj_n_samples;   /* row count of proj_log per channel; 0 = the samples of this call */
  int64_t proj_sample0;     /* first row this call writes (lets chained calls fill one buffer) */
  double *sd_state;         /* DEVICE in/out [4][n_instances]: filter delays z0,z1,z2 and sum(y); zero it before
                               the first call, chain it through calls that continue a run (start_iteration) */
} cmtjb_post_t;

/* cmtjb_run_device plus the fused reductions above.  All pointers are DEVICE pointers. */
int cmtjb_run_device_post(const cmtjb_junction_t *jn, const cmtjb_layer_t *layers,
                          const cmtjb_override_t *overrides, int32_t n_overrides,
                          const cmtjb_run_args_t *args, const cmtjb_post_t *post, double *m_state,
                          double *log, int32_t *status, int device, void *stream);

/* calculate_resistance_series / _parallel (cmtj/utils/resistance.py:96-149) over a device log:
 * m is [n_layers*3][n] (n = samples x instances, any flattening), weights[l] scales layer l's m
 * first (PIMM passes thickness*Ms/normalising_factor, procedures.py:152-162; VSD passes 1);
 * Rx, Ry are [n].  parallel = 0 -> series. */
int cmtjb_strip_resistance_device(int32_t n_layers, const double *res /* [n_layers][6] host */,
                                  const double *weights /* [n_layers] host */, int32_t parallel,
                                  const double *m, int64_t n, double *Rx, double *Ry, int device, void *stream);

/* compute_sd (cmtj/utils/resistance.py:12-22) for B series at once: R is DEVICE [n][B]; I is DEVICE [n]
 * (shared by all series) or [n][B] when current_per_series != 0; b, a are the HOST order-3 IIR
 * coefficients of Filters.butter_lowpass_filter (filters.py:33-44); out is DEVICE [B] = mean(lfilter(b, a, -I*R)). */
int cmtjb_compute_sd_device(const double *R, const double *I, int32_t current_per_series, int64_t n, int64_t B,
                            const double *b, const double *a, double *out, int device, void *stream);
/* compute_gmr (resistance.py:68-74): m1, m2 DEVICE [3][n], out DEVICE [n]. */
int cmtjb_compute_gmr_device(double Rp, double Rap, const double *m1, const double *m2, int64_t n, double *out,
                             int device, void *stream);

/* compute_spectrum_strip (cmtj/utils/procedures.py:29-40) for B series at once: x is DEVICE, time-major,
 * x[t * ld + b] for t < n, b < B (the projection log of cmtjb_run_device_post, or any row range of it);
 * out (DEVICE, [B][K]) receives |FFT(x_b)[k]| for the K lowest bins, K <= n/2 + 1, in double precision.
 * Strongly pruned spectra (K << n/2, the sweep-procedure case) reduce blocks of consecutive samples to Chebyshev
 * moments in one pass over x and transform the moments (truncation below 2e-17 of the series' l1 norm); otherwise a
 * single-pass pruned mixed-radix transform when n factors into 2, 3, 4, 5, else cuFFT over column slices.  No
 * transposed copy of the input in any case. */
int cmtjb_rfft_abs_device(const double *x, int64_t n, int64_t B, int64_t ld, int64_t K, double *out, int device,
                          void *stream);
/* The same with a caller-owned DEVICE workspace (32-byte aligned) for the moments of a slab of series, so that repeated
 * calls do not allocate: cmtjb_rfft_abs_workspace reports the bytes one series needs (0: this (n, K) takes a path that
 * needs no workspace); any multiple of it is accepted, series are processed workspace_bytes / bytes_per_series at a time. */
int cmtjb_rfft_abs_workspace(int64_t n, int64_t K, int device, int64_t *bytes_per_series);
int cmtjb_rfft_abs_device_ws(const double *x, int64_t n, int64_t B, int64_t ld, int64_t K, double *out, int device,
                             void *stream, void *workspace, int64_t workspace_bytes);

/* The same spectrum for series that are produced in TIME CHUNKS and never resident as a whole (a sweep's projection log at
 * full rate is 160 GB for 1e5 points x 2e5 steps).  cmtjb_spectrum_stream_plan reports the block length L the series must
 * be cut into and the bytes of moment storage one series needs (block_len = 0: no streaming plan for this (n, K); use
 * cmtjb_rfft_abs_device).  The caller owns moments[B * bytes_per_series] (DEVICE, 32-byte aligned) and, for every chunk of
 * nq whole blocks starting at block q0 (q0 even), passes the chunk's rows x_chunk[(t - q0 * L) * ld + b] to
 * cmtjb_spectrum_stream_moments (one pass over the chunk, any order of chunks, stream-ordered); when blocks
 * 0 .. n / L - 1 are all in, cmtjb_spectrum_stream_finish writes |FFT(x_b)[k]|, k < K, to out [B][K] (DEVICE).
 * background != 0 selects a low-footprint form of the reduction (128 threads, <= 128 registers per block) meant to run on a
 * second stream WHILE the next chunk integrates: it fits into the registers a resident time-loop kernel leaves free. */
int cmtjb_spectrum_stream_plan(int64_t n, int64_t K, int device, int64_t *block_len, int64_t *bytes_per_series);
int cmtjb_spectrum_stream_moments(const double *x_chunk, int64_t ld, int64_t B, int64_t n, int64_t K, int64_t q0, int64_t nq,
                                  void *moments, int32_t background, int device, void *stream);
int cmtjb_spectrum_stream_finish(const void *moments, int64_t n, int64_t B, int64_t K, double *out, int device, void *stream);

/* Number of kernels this library has launched in this process (monotonic). */
long long cmtjb_launch_count(void);
/* Name of the kernel variant (layer count / solver / compiled term set) the calling thread's last sweep launch
 * selected; diagnostic.  cmtjb_plan() cannot know it in general: per-instance overrides change the choice. */
const char *cmtjb_last_variant(void);
/* Which compiled variant a run of this model with these overrides would launch for its driver-evaluating
 * segment (no CUDA call; host logic only, used by the CPU tests of the selection rules).  `overrides[k].values`
 * are not dereferenced. */
int cmtjb_select_variant(const cmtjb_junction_t *jn, const cmtjb_layer_t *layers, const cmtjb_override_t *overrides,
                         int32_t n_overrides, const cmtjb_run_args_t *args, char *name, int32_t name_size);
/* cmtjb_run_host keeps one device scratch slab per device between calls; this frees them. */
int cmtjb_release_cache(void);

/* Where the calling thread's last cmtjb_run_host call spent its time, in the order of cmtjb_host_stat
 * (milliseconds unless named otherwise).  A time-chunked run reports device-timed figures: integration, gaps
 * on the compute stream while a ring buffer was still draining, the copy engine's busy time and the part of
 * the last chunk's copy that nothing could hide. */
enum cmtjb_host_stat {
  CMTJB_HOST_TOTAL_MS = 0, CMTJB_HOST_INTEGRATE_MS, CMTJB_HOST_RING_WAIT_MS, CMTJB_HOST_COPY_MS, CMTJB_HOST_TAIL_COPY_MS,
  CMTJB_HOST_PIPELINE_MS, CMTJB_HOST_LOG_BYTES, CMTJB_HOST_N_CHUNKS, CMTJB_HOST_N_RING, CMTJB_HOST_NSTATS
};
int cmtjb_host_stats(double *out, int32_t n);

/* BufferedAlphaNoise::fillBuffer (core/noise.hpp:190-227) for n_series independent series of n samples on the
 * device: out (HOST, float32 [n_series][n]) receives each series' `result` buffer (getFullBuffer), unscaled.
 * Series k is the one a sweep instance series0 + k draws for layer 0, x axis, under `seed`. */
int cmtjb_alpha_noise_host(double alpha, double std, int64_t n, uint64_t seed, int64_t series0, int64_t n_series,
                           float *out, int device);

/* ScalarDriver::getCurrentScalarValue (core/drivers.hpp:262-305) evaluated on the device
 * for n times (host pointers); used by the parity tests of the inline driver code. */
int cmtjb_eval_driver_host(const cmtjb_driver_t *driver, const double *times, int64_t n,
                           double *out, int device);

/* ---- coupled systems: Stack (core/stack.hpp) and GroupInteraction (core/reservoir.hpp) ---------
 *
 * A coupled system is `n_junctions` junction models that all have the same number of layers,
 * stepped together with a per-step exchange.  A batch is `args->n_instances` copies of the whole
 * system; an *element* is one junction of one copy, index e = instance * n_junctions + junction.
 *   jns     [n_junctions]               junction headers (all MR_CLASSIC for a stack)
 *   layers  [n_junctions * n_layers]    layer records, junction-major
 *   m_state [n_layers*3][n_elements]    in/out
 *   junction_log [3*n_layers+1|+3|+0][n_samples][n_elements]   per-junction logLayerParams channels
 * Overrides: `values` has one entry per ELEMENT; layer = CMTJB_LAYER_SYSTEM targets the stack's
 * coupled current driver instead, with one entry per INSTANCE (slot is ignored).
 * Solver: the reference always requests RK4 here (stack.hpp:50, reservoir.hpp:200), which
 * getSolver turns into Euler-Heun when any layer has a temperature; mixing is an error. */
#define CMTJB_LAYER_SYSTEM (-2)

enum cmtjb_stack_kind { CMTJB_STACK_SERIES = 0, CMTJB_STACK_PARALLEL = 1 };

/* Stack<T> members (core/stack.hpp:20-22, :106-110) */
typedef struct cmtjb_stack {
  int32_t n_junctions;
  int32_t kind;          /* cmtjb_stack_kind */
  int32_t use_kcl;       /* ctor argument useKCL (stack.hpp:131-139) */
  int32_t delayed;       /* setDelayed (stack.hpp:166-172) */
  int32_t top_layer;     /* index of the layer called topId inside each junction */
  int32_t bottom_layer;  /* index of bottomId, or -1 for single-layer junctions */
  int32_t reserved0, reserved1;
  double phaseOffset;
  double polarisation[3]; /* getPolarisationVector() (stack.hpp:285-296); used when junctions have one layer */
  const double *coupling; /* HOST pointer, [n_junctions-1] */
  cmtjb_driver_t currentDriver; /* setCoupledCurrentDriver (stack.hpp:214-220) */
} cmtjb_stack_t;

/* reservoir.hpp:63-120 */
enum cmtjb_interaction {
  CMTJB_INT_NULL = 0,
  CMTJB_INT_DIPOLE = 1,
  CMTJB_INT_DIPOLE_NOMURA = 2,
  /* Any interaction function that is AFFINE in the partner's magnetisation (every linear-response coupling is):
   *     H_a += c[a][b] + T[a][b] m_b        for every partner b != a,
   * with the 3x3 tensors and offsets tabulated by the caller -- the binding obtains them by probing the user's callable
   * (GroupInteraction::setInteractionFunction, reservoir.hpp:181; python/cmtj.cpp:438) with m_b = 0, e_x, e_y, e_z once
   * per ordered pair instead of calling it N^2 times per step.  cmtjb_group_t::pair_table. */
  CMTJB_INT_TABLE = 3
};

typedef struct cmtjb_group {
  int32_t n_junctions;
  int32_t interaction;   /* cmtjb_interaction */
  int32_t top_layer;     /* index of the layer called topId inside each junction */
  int32_t reserved;
  const double *coordinates; /* HOST pointer, [n_junctions][3] */
  /* CMTJB_INT_TABLE only (else NULL): HOST pointer, [n_junctions b][12][n_junctions a] doubles -- for the field at a from
   * partner b: rows 0..8 the tensor T[a][b] row-major, rows 9..11 the offset c[a][b]; the a == b entries are ignored. */
  const double *pair_table;
} cmtjb_group_t;

typedef struct cmtjb_coupled_plan {
  int32_t effective_solver;
  int32_t solver_overridden;
  int32_t n_layers;
  int32_t n_junction_channels; /* rows of junction_log */
  int32_t n_system_channels;   /* stack: 1 + n_junctions + (n_junctions-1); group: 0 */
  int32_t cooperative;         /* group: 1 when the array spans the grid (one launch, grid-wide barrier per step) */
  int64_t total_iterations, write_every, n_samples;
} cmtjb_coupled_plan_t;

int cmtjb_stack_plan(const cmtjb_stack_t *st, const cmtjb_junction_t *jns, const cmtjb_layer_t *layers,
                     const cmtjb_run_args_t *args, cmtjb_coupled_plan_t *plan);
/* Stack::runSimulation (core/stack.hpp:311-423).  stack_log: [n_system_channels][n_samples][n_instances]
 * with channel 0 = "Resistance", 1+j = "I_j", 1+n_junctions+j = "C_j_(j+1)" (stack.hpp:255-266).
 * Host pointers; synchronous. */
int cmtjb_stack_run_host(const cmtjb_stack_t *st, const cmtjb_junction_t *jns, const cmtjb_layer_t *layers,
                         const cmtjb_override_t *overrides, int32_t n_overrides, const cmtjb_run_args_t *args,
                         double *m_state, double *junction_log, double *stack_log, int32_t *status, int device);
/* Same with DEVICE pointers for m_state / logs / status / override values; asynchronous on stream. */
int cmtjb_stack_run_device(const cmtjb_stack_t *st, const cmtjb_junction_t *jns, const cmtjb_layer_t *layers,
                           const cmtjb_override_t *overrides, int32_t n_overrides, const cmtjb_run_args_t *args,
                           double *m_state, double *junction_log, double *stack_log, int32_t *status, int device,
                           void *stream);

int cmtjb_group_plan(const cmtjb_group_t *g, const cmtjb_junction_t *jns, const cmtjb_layer_t *layers,
                     const cmtjb_run_args_t *args, cmtjb_coupled_plan_t *plan);
/* GroupInteraction::runSimulation (core/reservoir.hpp:185-226). */
int cmtjb_group_run_host(const cmtjb_group_t *g, const cmtjb_junction_t *jns, const cmtjb_layer_t *layers,
                         const cmtjb_override_t *overrides, int32_t n_overrides, const cmtjb_run_args_t *args,
                         double *m_state, double *junction_log, int32_t *status, int device);
int cmtjb_group_run_device(const cmtjb_group_t *g, const cmtjb_junction_t *jns, const cmtjb_layer_t *layers,
                           const cmtjb_override_t *overrides, int32_t n_overrides, const cmtjb_run_args_t *args,
                           double *m_state, double *junction_log, int32_t *status, int device, void *stream);

/* One dipole-coupled array sharded over `world` ranks, one process per GPU (reservoir.hpp:129-152 with the elements
 * split into contiguous ranges of ceil(n_junctions / world)): this rank integrates its own elements and, after the
 * initial state (step = -1) and after every step, calls `exchange` to all-gather the top-layer magnetisations in place:
 * buf is a DEVICE buffer [world][count] doubles whose slice `rank` this rank has just written on `stream`; the callback
 * must enqueue the all-gather on that stream (NCCL: ncclAllGather(buf + rank*count, buf, count, ncclDouble, comm, stream))
 * and return 0.  g / jns / layers describe the WHOLE array on every rank; m_state [3*n_layers][own], junction_log
 * [channels][n_samples][own] and status [own] are DEVICE pointers sized for this rank's elements.  The pair sums are
 * accumulated in an order that depends on n_junctions alone, so the result is bit-identical for every world size.
 * args->n_instances must be 1; thermal noise is keyed by the global element index; 1/f noise is not supported here. */
typedef int (*cmtjb_exchange_fn)(void *ctx, double *buf, int64_t count, int64_t step, void *stream);
int cmtjb_group_run_sharded(const cmtjb_group_t *g, const cmtjb_junction_t *jns, const cmtjb_layer_t *layers,
                            const cmtjb_run_args_t *args, int32_t rank, int32_t world, cmtjb_exchange_fn exchange, void *ctx,
                            double *m_state, double *junction_log, int32_t *status, int device, void *stream);

/* The same array with the exchange done by the kernels themselves over PEER MEMORY (NVLink / NVSwitch): no collective
 * library and no host in the time loop.  Every rank owns an exchange area of cmtjb_group_peer_area_bytes(n_junctions, world)
 * bytes in its own device memory (cmtjb_peer_alloc: cudaMalloc, zero-filled) that the other ranks have mapped
 * (cmtjb_peer_export -> 64-byte CUDA IPC handle, sent by any means -> cmtjb_peer_open in the other processes; ranks that are
 * threads of one process pass each other's pointers directly).  areas[r] is rank r's area as addressable from THIS process,
 * areas[rank] the own one.  Each step's kernel stores the rank's new top-layer magnetisations into every area and raises a
 * per-rank arrival counter there behind a system-scope fence; a one-warp kernel waits on the own area's counters before the
 * next pair sums (no deadlock: a rank never waits for anything a peer has not already been able to launch; a peer that
 * does not arrive within CMTJB_PEER_TIMEOUT_MS, default 20000, is CMTJB_ERUNTIME).  All areas must be allocated (and zero)
 * before any rank calls, and freed only after every rank has returned: two barriers of the caller's.  An area serves ONE call.
 * At most 16 ranks.  Results are bit-identical to cmtjb_group_run_sharded and to a single rank. */
int64_t cmtjb_group_peer_area_bytes(int32_t n_junctions, int32_t world);
int cmtjb_group_run_sharded_peer(const cmtjb_group_t *g, const cmtjb_junction_t *jns, const cmtjb_layer_t *layers,
                                 const cmtjb_run_args_t *args, int32_t rank, int32_t world, void *const *areas, double *m_state,
                                 double *junction_log, int32_t *status, int device, void *stream);
int cmtjb_peer_alloc(int device, int64_t bytes, void **ptr);
int cmtjb_peer_free(int device, void *ptr);
int cmtjb_peer_export(int device, void *ptr, unsigned char *handle64);
int cmtjb_peer_open(int device, const unsigned char *handle64, void **ptr);
int cmtjb_peer_close(int device, void *ptr);

/* RNG test hooks.  cmtjb_philox_host: one Philox4x32-10 block for counter ctr[4] / key key[2]
 * (known-answer tests).  cmtjb_normals_host: the N(0,1) triples the stochastic steppers draw for
 * instances inst0..inst0+n_inst-1, layer `layer`, steps step0..step0+n_steps-1 under `seed`
 * (replaces std::normal_distribution + mt19937, core/junction.hpp:174-175);
 * out is [n_steps][n_inst][3]. */
int cmtjb_philox_host(const uint32_t *ctr, const uint32_t *key, uint32_t *out, int device);
int cmtjb_normals_host(uint64_t seed, int64_t inst0, int64_t n_inst, int32_t layer, int64_t step0, int64_t n_steps,
                       double *out, int device);
/* the same for either stream: thermal_normals as in cmtjb_run_args_t */
int cmtjb_normals_host_ex(uint64_t seed, int64_t inst0, int64_t n_inst, int32_t layer, int64_t step0, int64_t n_steps,
                          int32_t thermal_normals, double *out, int device);

/* FP64 FMA-pipe microbenchmark used for the roofline denominator: runs `iters` dependent
 * DFMA chains on every SM and returns achieved FLOP/s in *flops (2 flop per DFMA). */
int cmtjb_fp64_peak(int device, int64_t iters, double *flops, double *ms);
/* SM clock (MHz) at which the calling thread's last cmtjb_fp64_peak actually ran, from the cycle counter and the global
 * nanosecond timer read inside the kernel (0 before the first call). */
double cmtjb_fp64_peak_sm_mhz(void);
/* FP64 pipe probe: the fraction of the pipe's issue slots (64 lanes per SM and clock, measured with the in-kernel cycle
 * counter) that independent chains sustain at `warps_per_subpartition` resident warps.  kind 0: DFMA with two constant-bank
 * operands; 1: DFMA with three register operands; 2: the S1 step's 12 DFMA : 3 DMUL : 1 DADD mix on register operands;
 * 3: the dependency shape of one right-hand-side evaluation (3-vector cross/dot/axpy, 28 instructions); 4: the same on two
 * independent vectors per thread; 5: independent chains whose DFMAs read registers of three different chains.  Diagnostics
 * only: it bounds what the integration kernels can reach. */
int cmtjb_fp64_pipe_probe(int device, int kind, int warps_per_subpartition, int64_t iters, double *issue_slot_frac,
                          double *sm_mhz);

#ifdef __cplusplus
}
#endif
#endif /* CMTJ_B200_H_ */
