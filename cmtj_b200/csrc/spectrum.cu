// spectrum.cu -- |FFT| of a whole sweep's projected signals (compute_spectrum_strip, cmtj/utils/procedures.py:29-40).
//
// The fused procedures log one projected double per instance and step, TIME-major ([t][instance], the layout the
// time loop writes coalesced).  The reference transforms one series at a time with scipy; here cuFFT runs a
// double-precision real-to-complex transform over column slices of that buffer IN PLACE OF ANY TRANSPOSE (advanced
// layout: element stride = the row pitch, consecutive series one double apart), and a second kernel keeps
// |X[k]| of the K lowest bins per series, transposed to [instance][K] -- the orientation the caller hands back.
#include <cuda_runtime.h>
#include <cufft.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <vector>

#include "host_common.h"

namespace cmtjb {
namespace {

struct SpecPlan {
  int device;
  long long n, cols, ld;
  cufftHandle h;
  size_t work;
};
std::mutex g_spec_mu;
std::vector<SpecPlan> g_spec_plans;

// half[k * cols + c] (complex) -> out[(c0 + c) * K + k] = |half|, k < K
__global__ void abs_transpose_kernel(const double2 *half, long long cols, long long K, double *out, long long c0) {
  __shared__ double tile[32][33];
  const long long kb = (long long)blockIdx.x * 32, cb = (long long)blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const long long k = kb + r, c = cb + threadIdx.x;
    double v = 0.0;
    if (k < K && c < cols) {
      const double2 z = half[k * cols + c];
      v = hypot(z.x, z.y);
    }
    tile[r][threadIdx.x] = v;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const long long c = cb + r, k = kb + threadIdx.x;
    if (c < cols && k < K) out[(c0 + c) * K + k] = tile[threadIdx.x][r];
  }
}

}  // namespace
}  // namespace cmtjb

using namespace cmtjb;

extern "C" int cmtjb_rfft_abs_device(const double *x, int64_t n, int64_t B, int64_t ld, int64_t K, double *out, int device,
                                     void *stream_) {
  if (!x || !out || n <= 0 || B <= 0 || ld < B || K < 0 || K > n / 2 + 1) return fail(CMTJB_EINVAL, "bad argument");
  if (K == 0) return CMTJB_OK;
  if (n > (1LL << 30) || ld > (1LL << 30)) return fail(CMTJB_ENOTSUP, "series too long for the 32-bit cuFFT plan interface");
  CUDA_TRY(cudaSetDevice(device));
  cudaStream_t stream = (cudaStream_t)stream_;
  const long long nh = n / 2 + 1;
  // column slices: the complex half-spectra of a slice stay below ~1 GiB (CMTJB_SPECTRUM_SLICE_MB overrides)
  size_t slice_bytes = (size_t)1 << 30;
  if (const char *c = std::getenv("CMTJB_SPECTRUM_SLICE_MB")) slice_bytes = (size_t)std::max(1, std::atoi(c)) << 20;
  long long cols = (long long)std::max<size_t>(1, slice_bytes / ((size_t)nh * sizeof(double2)));
  cols = std::min<long long>(cols, B);
  if (cols >= 32) cols = (cols / 32) * 32;
  double2 *half = nullptr;
  cudaError_t e = cudaMallocAsync((void **)&half, (size_t)nh * cols * sizeof(double2), stream);
  if (e != cudaSuccess) return cuda_fail(e, "cudaMallocAsync(half spectra)");
  int rc = CMTJB_OK;
  std::lock_guard<std::mutex> lock(g_spec_mu);  // cached plans carry their stream and work area
  for (long long c0 = 0; c0 < B && rc == CMTJB_OK; c0 += cols) {
    const long long nc = std::min(cols, B - c0);
    cufftHandle h = 0;
    bool found = false;
    for (const SpecPlan &p : g_spec_plans)
      if (p.device == device && p.n == n && p.cols == nc && p.ld == ld) {
        h = p.h;
        found = true;
      }
    if (!found) {
      if (g_spec_plans.size() >= 6) {
        cudaDeviceSynchronize();
        cufftDestroy(g_spec_plans.front().h);
        g_spec_plans.erase(g_spec_plans.begin());
      }
      int nn[1] = {(int)n}, inembed[1] = {(int)n}, onembed[1] = {(int)nh};
      // input: series c is x[t * ld + c] (stride ld, distance 1); output: half[k * nc + c] (stride nc, distance 1)
      const cufftResult r = cufftPlanMany(&h, 1, nn, inembed, (int)ld, 1, onembed, (int)nc, 1, CUFFT_D2Z, (int)nc);
      if (r != CUFFT_SUCCESS) {
        char buf[160];
        std::snprintf(buf, sizeof(buf), "cufftPlanMany(D2Z, %lld points x %lld series) failed with cufftResult %d", (long long)n, nc, (int)r);
        rc = fail(CMTJB_ECUDA, buf);
        break;
      }
      g_spec_plans.push_back(SpecPlan{device, (long long)n, nc, (long long)ld, h, 0});
    }
    if (cufftSetStream(h, stream) != CUFFT_SUCCESS) rc = fail(CMTJB_ECUDA, "cufftSetStream failed");
    if (rc == CMTJB_OK && cufftExecD2Z(h, (cufftDoubleReal *)(x + c0), (cufftDoubleComplex *)half) != CUFFT_SUCCESS)
      rc = fail(CMTJB_ECUDA, "cufftExecD2Z failed");
    if (rc == CMTJB_OK) {
      const dim3 grid((unsigned)((K + 31) / 32), (unsigned)((nc + 31) / 32));
      abs_transpose_kernel<<<grid, dim3(32, 8), 0, stream>>>(half, nc, K, out, c0);
      g_launches++;
      e = cudaGetLastError();
      if (e != cudaSuccess) rc = cuda_fail(e, "abs_transpose_kernel");
    }
  }
  cudaFreeAsync(half, stream);
  return rc;
}
