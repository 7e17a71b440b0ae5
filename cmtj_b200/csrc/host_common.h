// host_common.h -- host-side helpers shared by the translation units of libcmtj_b200.so.
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <string>

#include "dev_model.h"

namespace cmtjb {

extern thread_local std::string g_err;
extern std::atomic<long long> g_launches;

int fail(int code, const std::string &msg);
int cuda_fail(cudaError_t e, const char *what);
#define CUDA_TRY(expr)                                          \
  do {                                                          \
    cudaError_t _e = (expr);                                    \
    if (_e != cudaSuccess) return ::cmtjb::cuda_fail(_e, #expr); \
  } while (0)

// The library's small per-launch scratch (models, stage tables, scheduler tickets) comes from the device's stream-ordered
// pool.  By default that pool hands its memory back to the OS at every synchronisation, so each call of a sweep loop pays
// the driver's map / unmap again; this lets the current device's pool keep up to 256 MB (once per device and process).
void keep_pool_memory();

// validation + derived quantities (cabi.cu)
int make_plan(const cmtjb_junction_t *jn, const cmtjb_layer_t *layers, const cmtjb_run_args_t *a, cmtjb_plan_t *p);
void derive_layer(const cmtjb_junction_t &J, const cmtjb_layer_t &L, DevLayer &D);
void set_step_products(DevLayer &D, double dt);  // h = llg*dt etc. for this launch's time step
// feature bits (dev_model.h: F_*) a junction model needs compiled into its kernel
uint32_t model_features(const cmtjb_junction_t *jn, const cmtjb_layer_t *layers, const cmtjb_override_t *ovr, int n_ovr,
                        const cmtjb_run_args_t *a, int solver);

}  // namespace cmtjb
