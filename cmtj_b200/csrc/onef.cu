// onef.cu -- 1/f^alpha noise on the device: BufferedAlphaNoise::fillBuffer (core/noise.hpp:190-227) for
// thousands of independent series at once.
//
// The reference gives every layer with scaleNoise != 0 one VectorAlphaNoise per run (getSolver,
// junction.hpp:1905-1916): up to three BufferedAlphaNoise components of bufferSize = totalIterations
// float32 samples each, filled at the first tick by Kasdin's filter through a 2n-point kissfft:
//     white[0..n) ~ N(0, std) (float), zero padded to 2n;      h[0] = 1, h[i] = h[i-1] (alpha/2 + i - 1) / i
//     P[k] = FFT(h)[k] * FFT(white)[k] for k < n,  P[0] /= 2,  P[n-1] /= 2,  P[k >= n] = 0
//     result[t] = Re(IFFT_unnormalised(P)[t]) / n                    (the one-sided-spectrum form, quirks kept)
// A sweep needs one such series per (instance, layer, axis).  Here: white noise from Philox4x32-10 keyed by
// (seed, global instance, layer, axis, sample) -- so shards of a sweep reproduce the single-GPU batch --,
// Box-Muller in float32, the same spectrum product, cuFFT C2C in float32 (set-up work), and a
// tiled transpose into the layout the time loop reads coalesced:  out[(slot * n + t) * Bc + instance].
#include <cuda_runtime.h>
#include <cufft.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <string>
#include <vector>

#include "host_common.h"
#include "llgs_device.cuh"
#include "onef.h"

namespace cmtjb {
namespace {

// W[g][t] = (std * N(0,1), 0) for t < n, 0 for n <= t < 2n.  One thread per four consecutive samples.
// Series g belongs to global element elem0 + g (sweeps: instances; coupled systems: junctions of all instances).
__global__ void onef_fill_kernel(float2 *W, long long n, long long G, unsigned long long seed, unsigned long long elem0,
                                 unsigned tag, float sd) {
  const long long n2 = 2 * n;
  const long long q4 = n2 / 4 + ((n2 % 4) ? 1 : 0);  // blocks of four samples per series
  const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= q4 * G) return;
  const long long g = q / q4, b = q - g * q4;
  const long long t0 = 4 * b;
  float v[4] = {0.f, 0.f, 0.f, 0.f};
  if (t0 < n) {
    const unsigned long long e = elem0 + (unsigned long long)g;
    const uint4 ctr = make_uint4((uint32_t)e, (uint32_t)(e >> 32), (uint32_t)b, (((uint32_t)((unsigned long long)b >> 32)) << 8) | tag);
    const float4 z = normal4(philox4x32_10(ctr, make_uint2((uint32_t)seed, (uint32_t)(seed >> 32))));
    v[0] = z.x * sd;
    v[1] = z.y * sd;
    v[2] = z.z * sd;
    v[3] = z.w * sd;
  }
  float2 *w = W + g * n2 + t0;
#pragma unroll
  for (int k = 0; k < 4; k++)
    if (t0 + k < n2) w[k] = make_float2(t0 + k < n ? v[k] : 0.f, 0.f);
}

// pad the filter weights to 2n complex
__global__ void onef_pad_kernel(float2 *C, const float *h, long long n) {
  const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k < 2 * n) C[k] = make_float2(k < n ? h[k] : 0.f, 0.f);
}

// noise.hpp:206-219: P = FFT(h) .* FFT(white) on the lower half, both "end" bins halved, upper half cleared
__global__ void onef_mul_kernel(float2 *W, const float2 *C, long long n, long long G) {
  const long long n2 = 2 * n;
  const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n2 * G) return;
  const long long k = q % n2;
  float2 o = make_float2(0.f, 0.f);
  if (k < n) {
    const float2 c = C[k], w = W[q];
    o = make_float2(c.x * w.x - c.y * w.y, c.x * w.y + c.y * w.x);
    if (k == 0) o = make_float2(o.x * 0.5f, o.y * 0.5f);
    if (k == n - 1) o = make_float2(o.x * 0.5f, o.y * 0.5f);
  }
  W[q] = o;
}

// result[t] = Re(y[t]) / n (noise.hpp:222-226), transposed: series-major FFT rows -> [t][instance]
__global__ void onef_store_kernel(const float2 *W, long long n, long long Gc, float *out, long long ld) {
  __shared__ float tile[32][33];
  const long long t0 = (long long)blockIdx.x * 32, g0 = (long long)blockIdx.y * 32;
  const float nf = (float)n;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const long long g = g0 + r, t = t0 + threadIdx.x;
    tile[r][threadIdx.x] = (g < Gc && t < n) ? W[g * 2 * n + t].x / nf : 0.f;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const long long t = t0 + r, g = g0 + threadIdx.x;
    if (t < n && g < Gc) out[t * ld + g] = tile[threadIdx.x][r];
  }
}

// cuFFT plans are kept for the life of the process (creating one costs milliseconds, and destroying one
// with work in flight would free its work area under the kernels)
struct PlanKey {
  int device;
  long long n2, batch;
  cufftHandle h;
};
std::mutex g_plan_mu;
std::vector<PlanKey> g_plans;

int get_plan(int device, long long n2, long long batch, cufftHandle *out) {
  std::lock_guard<std::mutex> lock(g_plan_mu);
  for (const PlanKey &p : g_plans)
    if (p.device == device && p.n2 == n2 && p.batch == batch) {
      *out = p.h;
      return CMTJB_OK;
    }
  if (g_plans.size() >= 8) {  // evict the oldest; nothing may still be running on it
    cudaDeviceSynchronize();
    cufftDestroy(g_plans.front().h);
    g_plans.erase(g_plans.begin());
  }
  if (n2 > (1LL << 27)) return fail(CMTJB_ENOTSUP, "1/f^alpha noise: more than 2^26 iterations per run exceed cuFFT's single-precision 1-D limit");
  cufftHandle h;
  int nn[1] = {(int)n2};
  const cufftResult r = cufftPlanMany(&h, 1, nn, nullptr, 1, (int)n2, nullptr, 1, (int)n2, CUFFT_C2C, (int)batch);
  if (r != CUFFT_SUCCESS) {
    char buf[160];
    std::snprintf(buf, sizeof(buf), "cufftPlanMany(%lld points x %lld series) failed with cufftResult %d", n2, batch, (int)r);
    return fail(CMTJB_ECUDA, buf);
  }
  g_plans.push_back(PlanKey{device, n2, batch, h});
  *out = h;
  return CMTJB_OK;
}

}  // namespace

void kasdin_weights(long long n, double alpha, std::vector<float> &h) {
  // noise.hpp:173-177, in the reference's types: (float)(0.5*alpha + (float)(i-1)) / (float)i, float product
  h.resize((size_t)std::max<long long>(n, 0));
  if (n <= 0) return;
  h[0] = 1.0f;
  for (long long i = 1; i < n; i++) {
    const float weight = (float)(0.5 * alpha + (double)((float)(i - 1))) / ((float)i);
    h[i] = h[i - 1] * weight;
  }
}

size_t onef_scratch_bytes(long long n, long long group) { return (size_t)group * 2 * (size_t)n * sizeof(float2); }

long long onef_group_size(long long n, long long n_series) {
  // FFT rows in flight: about 1 GiB of complex float32 (cuFFT's own work area is of the same order)
  size_t budget = (size_t)1 << 30;
  if (const char *c = std::getenv("CMTJB_ONEF_SCRATCH_MB")) budget = (size_t)std::max(1, std::atoi(c)) << 20;
  const long long per = (long long)(2 * (size_t)n * sizeof(float2));
  return std::max<long long>(1, std::min<long long>(n_series, (long long)(budget / (size_t)std::max<long long>(per, 1))));
}

std::mutex g_gen_mu;  // cached plans carry their stream: one generation is enqueued at a time

int onef_generate(int device, cudaStream_t stream, long long n, const OnefSlot *slots, int n_slots, unsigned long long seed,
                  unsigned long long elem_global0, long long Bc, float *out) {
  if (n <= 0 || Bc <= 0 || n_slots <= 0) return CMTJB_OK;
  std::lock_guard<std::mutex> gen_lock(g_gen_mu);
  const long long n2 = 2 * n;
  const long long G = onef_group_size(n, Bc);
  cufftHandle plan, plan1;
  int rc = get_plan(device, n2, G, &plan);
  if (rc) return rc;
  rc = get_plan(device, n2, 1, &plan1);
  if (rc) return rc;
  if (cufftSetStream(plan, stream) != CUFFT_SUCCESS || cufftSetStream(plan1, stream) != CUFFT_SUCCESS)
    return fail(CMTJB_ECUDA, "cufftSetStream failed");
  float2 *W = nullptr, *C = nullptr;
  float *d_h = nullptr;
  cudaError_t e = cudaMallocAsync((void **)&W, onef_scratch_bytes(n, G), stream);
  if (e == cudaSuccess) e = cudaMallocAsync((void **)&C, (size_t)n2 * sizeof(float2), stream);
  if (e == cudaSuccess) e = cudaMallocAsync((void **)&d_h, (size_t)n * sizeof(float), stream);
  if (e != cudaSuccess) {
    if (W) cudaFreeAsync(W, stream);
    if (C) cudaFreeAsync(C, stream);
    char buf[200];
    std::snprintf(buf, sizeof(buf), "1/f^alpha noise: cannot allocate %zu bytes of FFT scratch (%s)", onef_scratch_bytes(n, G) + (size_t)n2 * 8,
                  cudaGetErrorString(e));
    return fail(CMTJB_ECUDA, buf);
  }
  std::vector<float> h;
  double last_alpha = -1.0;
  rc = CMTJB_OK;
  for (int s = 0; s < n_slots && rc == CMTJB_OK; s++) {
    if (slots[s].alpha != last_alpha) {
      // spectrum of the filter weights: once per distinct alpha (pageable copy: staged before the call returns)
      kasdin_weights(n, slots[s].alpha, h);
      e = cudaMemcpyAsync(d_h, h.data(), (size_t)n * sizeof(float), cudaMemcpyHostToDevice, stream);
      if (e == cudaSuccess) {
        onef_pad_kernel<<<(unsigned)((n2 + 255) / 256), 256, 0, stream>>>(C, d_h, n);
        g_launches++;
        if (cufftExecC2C(plan1, C, C, CUFFT_FORWARD) != CUFFT_SUCCESS) rc = fail(CMTJB_ECUDA, "cufftExecC2C(filter weights) failed");
      } else {
        rc = cuda_fail(e, "cudaMemcpyAsync(Kasdin weights)");
      }
      last_alpha = slots[s].alpha;
    }
    const unsigned tag = 0x80u | ((unsigned)(slots[s].axis & 3) << 4) | ((unsigned)slots[s].layer & 0xFu);
    for (long long g0 = 0; g0 < Bc && rc == CMTJB_OK; g0 += G) {
      const long long Gc = std::min(G, Bc - g0);
      const long long q4 = n2 / 4 + ((n2 % 4) ? 1 : 0);
      onef_fill_kernel<<<(unsigned)((q4 * Gc + 255) / 256), 256, 0, stream>>>(W, n, Gc, seed, elem_global0 + (unsigned long long)g0, tag,
                                                                           (float)slots[s].std);
      if (cufftExecC2C(plan, W, W, CUFFT_FORWARD) != CUFFT_SUCCESS) rc = fail(CMTJB_ECUDA, "cufftExecC2C(forward) failed");
      onef_mul_kernel<<<(unsigned)((n2 * Gc + 255) / 256), 256, 0, stream>>>(W, C, n, Gc);
      if (rc == CMTJB_OK && cufftExecC2C(plan, W, W, CUFFT_INVERSE) != CUFFT_SUCCESS) rc = fail(CMTJB_ECUDA, "cufftExecC2C(inverse) failed");
      const dim3 grid((unsigned)((n + 31) / 32), (unsigned)((Gc + 31) / 32));
      onef_store_kernel<<<grid, dim3(32, 8), 0, stream>>>(W, n, Gc, out + (size_t)s * n * Bc + g0, Bc);
      g_launches += 3;
      e = cudaGetLastError();
      if (e != cudaSuccess && rc == CMTJB_OK) rc = cuda_fail(e, "1/f^alpha noise kernels");
    }
  }
  cudaFreeAsync(W, stream);
  cudaFreeAsync(C, stream);
  cudaFreeAsync(d_h, stream);
  return rc;
}

int onef_slots_of(const cmtjb_layer_t *layers, int n_layers, OnefSlot *slots /* [3 * n_layers] */, int *slot_of_layer /* [n_layers] */) {
  int ns = 0;
  for (int l = 0; l < n_layers; l++) {
    slot_of_layer[l] = -1;
    const cmtjb_layer_t &L = layers[l];
    if (L.noiseScale == 0) continue;
    slot_of_layer[l] = ns;
    if (L.noiseAxis == CMTJB_AXIS_ALL) {
      for (int c = 0; c < 3; c++) slots[ns++] = OnefSlot{l, c, L.noiseAlpha, L.noiseStd};
    } else if (L.noiseAxis >= CMTJB_AXIS_X && L.noiseAxis <= CMTJB_AXIS_Z) {
      slots[ns++] = OnefSlot{l, L.noiseAxis, L.noiseAlpha, L.noiseStd};
    } else {
      return -1;  // VectorAlphaNoise: "Invalid axis specified" (noise.hpp:285-287)
    }
  }
  return ns;
}

}  // namespace cmtjb

using namespace cmtjb;

extern "C" int cmtjb_alpha_noise_host(double alpha, double std, int64_t n, uint64_t seed, int64_t series0, int64_t n_series,
                                      float *out, int device) {
  if (!out || n <= 0 || n_series <= 0) return fail(CMTJB_EINVAL, "bad argument");
  if (alpha < 0 || alpha > 2) return fail(CMTJB_ERUNTIME, "alpha must be between 0 and 2");
  CUDA_TRY(cudaSetDevice(device));
  float *d = nullptr;
  CUDA_TRY(cudaMalloc((void **)&d, (size_t)n * n_series * sizeof(float)));
  const OnefSlot slot{0, 0, alpha, std};
  int rc = onef_generate(device, nullptr, n, &slot, 1, seed, (unsigned long long)series0, n_series, d);
  if (rc == CMTJB_OK) {
    // device layout is [t][series]; the caller gets [series][t] like BufferedAlphaNoise::getFullBuffer per object
    std::vector<float> tmp((size_t)n * n_series);
    cudaError_t e = cudaMemcpy(tmp.data(), d, tmp.size() * sizeof(float), cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) rc = cuda_fail(e, "cudaMemcpy D2H(1/f noise)");
    else
      for (int64_t t = 0; t < n; t++)
        for (int64_t s = 0; s < n_series; s++) out[(size_t)s * n + t] = tmp[(size_t)t * n_series + s];
  }
  cudaFree(d);
  return rc;
}
