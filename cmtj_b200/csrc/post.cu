// post.cu -- stand-alone post-processing kernels of the sweep procedures:
// the resistance models of cmtj/utils/resistance.py evaluated on a device-resident log.
#include <cuda_runtime.h>

#include <cstring>

#include "host_common.h"

namespace cmtjb {
namespace {

struct ResParams {
  double res[CMTJB_POST_MAX_LAYERS][CMTJB_RES_NFIELDS];
  double w[CMTJB_POST_MAX_LAYERS];
  int32_t n_layers, parallel;
};

// compute_resistance (resistance.py:25-66) + the series / parallel combination (:96-149):
//   Sx_l = Rxx0 + (AMR mx^2 + SMR my^2),  Sy_l = Rxy0 + 0.5 AHE mz + (w/l)(SMR + AMR) mx my
//   series: sum_l ; parallel: 1 / sum_l 1/S_l
// One thread per (sample, instance) element; m is [n_layers*3][n] so every load is coalesced.
__global__ void __launch_bounds__(256) strip_resistance_kernel(const __grid_constant__ ResParams P, const double *__restrict__ m,
                                                               long long n, double *__restrict__ Rx, double *__restrict__ Ry) {
  const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n) return;
  double ax = 0.0, ay = 0.0;
  for (int l = 0; l < P.n_layers; l++) {
    const double *r = P.res[l];
    const double mx = __ldcs(m + (3 * l + 0) * n + q) * P.w[l];
    const double my = __ldcs(m + (3 * l + 1) * n + q) * P.w[l];
    const double mz = __ldcs(m + (3 * l + 2) * n + q) * P.w[l];
    const double sx = r[CMTJB_RES_RXX0] + (r[CMTJB_RES_AMR] * (mx * mx) + r[CMTJB_RES_SMR] * (my * my));
    const double sy = r[CMTJB_RES_RXY0] + 0.5 * r[CMTJB_RES_AHE] * mz + r[CMTJB_RES_W_OVER_L] * (r[CMTJB_RES_SMR] + r[CMTJB_RES_AMR]) * mx * my;
    ax += P.parallel ? 1.0 / sx : sx;
    ay += P.parallel ? 1.0 / sy : sy;
  }
  __stcs(Rx + q, P.parallel ? 1.0 / ax : ax);
  __stcs(Ry + q, P.parallel ? 1.0 / ay : ay);
}

// compute_sd (resistance.py:12-22) for B independent series stored [n][B]: SD = -I*R, order-3 IIR in
// scipy's lfilter operation order (no FMA contraction), mean of the output.  One thread per series;
// the recurrence is sequential in n, the B axis is the parallel one (coalesced loads).
struct SdParams {
  double b[4], a[4];
};
__global__ void __launch_bounds__(128) sd_filter_kernel(const __grid_constant__ SdParams P, const double *__restrict__ R,
                                                        const double *__restrict__ I, long long i_stride, long long n, long long B,
                                                        double *__restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B) return;
  double z0 = 0.0, z1 = 0.0, z2 = 0.0, sum = 0.0;
  for (long long k = 0; k < n; k++) {
    const double cur = i_stride ? __ldg(I + k * i_stride + i) : __ldg(I + k);
    const double x = -cur * __ldcs(R + k * B + i);
    const double y = __dadd_rn(z0, __dmul_rn(P.b[0], x));
    z0 = __dsub_rn(__dadd_rn(z1, __dmul_rn(x, P.b[1])), __dmul_rn(y, P.a[1]));
    z1 = __dsub_rn(__dadd_rn(z2, __dmul_rn(x, P.b[2])), __dmul_rn(y, P.a[2]));
    z2 = __dsub_rn(__dmul_rn(x, P.b[3]), __dmul_rn(y, P.a[3]));
    sum += y;
  }
  out[i] = sum / (double)n;
}

// compute_gmr (resistance.py:68-74): Rp + 0.5 (Rap - Rp)(1 - m1.m2); m1, m2 are [3][n]
__global__ void __launch_bounds__(256) gmr_kernel(double Rp, double Rhalf, const double *__restrict__ m1,
                                                  const double *__restrict__ m2, long long n, double *__restrict__ out) {
  const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n) return;
  const double c = m1[q] * m2[q] + m1[n + q] * m2[n + q] + m1[2 * n + q] * m2[2 * n + q];
  out[q] = Rp + Rhalf * (1.0 - c);
}

}  // namespace
}  // namespace cmtjb

using namespace cmtjb;

extern "C" int cmtjb_compute_sd_device(const double *R, const double *I, int32_t current_per_series, int64_t n, int64_t B,
                                       const double *b, const double *a, double *out, int device, void *stream) {
  if (!R || !I || !b || !a || !out) return fail(CMTJB_EINVAL, "null argument");
  if (n <= 0 || B <= 0) return CMTJB_OK;
  CUDA_TRY(cudaSetDevice(device));
  SdParams P;
  for (int k = 0; k < 4; k++) {
    P.b[k] = b[k] / a[0];
    P.a[k] = a[k] / a[0];
  }
  sd_filter_kernel<<<(unsigned)((B + 127) / 128), 128, 0, (cudaStream_t)stream>>>(P, R, I, current_per_series ? (long long)B : 0LL,
                                                                                   (long long)n, (long long)B, out);
  g_launches++;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "sd_filter_kernel");
  return CMTJB_OK;
}

extern "C" int cmtjb_compute_gmr_device(double Rp, double Rap, const double *m1, const double *m2, int64_t n, double *out,
                                        int device, void *stream) {
  if (!m1 || !m2 || !out) return fail(CMTJB_EINVAL, "null argument");
  if (n <= 0) return CMTJB_OK;
  CUDA_TRY(cudaSetDevice(device));
  gmr_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(Rp, 0.5 * (Rap - Rp), m1, m2, (long long)n, out);
  g_launches++;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "gmr_kernel");
  return CMTJB_OK;
}

extern "C" int cmtjb_strip_resistance_device(int32_t n_layers, const double *res, const double *weights, int32_t parallel,
                                             const double *m, int64_t n, double *Rx, double *Ry, int device, void *stream) {
  if (!res || !m || !Rx || !Ry) return fail(CMTJB_EINVAL, "null argument");
  if (n_layers <= 0 || n_layers > CMTJB_POST_MAX_LAYERS) return fail(CMTJB_ENOTSUP, "post-processing supports 1..4 layers");
  if (n <= 0) return CMTJB_OK;
  CUDA_TRY(cudaSetDevice(device));
  ResParams P;
  std::memset(&P, 0, sizeof(P));
  std::memcpy(P.res, res, sizeof(double) * CMTJB_RES_NFIELDS * n_layers);
  for (int l = 0; l < n_layers; l++) P.w[l] = weights ? weights[l] : 1.0;
  P.n_layers = n_layers;
  P.parallel = parallel ? 1 : 0;
  strip_resistance_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(P, m, (long long)n, Rx, Ry);
  g_launches++;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "strip_resistance_kernel");
  return CMTJB_OK;
}
