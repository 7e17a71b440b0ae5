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
#include <cmath>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cstring>
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


// ---- single-pass pruned DFT -------------------------------------------------------------------------------------
// A sweep's spectra keep only the K lowest bins of n-sample series (PIMM: 1601 of 100 001 at 20 ns / 80 GHz), but a
// library FFT computes them all, in several passes over a buffer that is most of the GPU's memory (160 GB).  With
// n = N1 N2 and N1 >= K, Cooley-Tukey's first split needs no second index for those bins:
//     X[k] = sum_{n2 < N2} W_n^(n2 k) Y_n2[k],    Y_n2 = FFT_N1( x[N2 n1 + n2], n1 < N1 ),    k < K <= N1
// so one block can walk ONE series once: for each pair of decimated sequences (n2, n2 + 1), packed as one complex
// sequence, an N1-point mixed-radix FFT in shared memory (in place: samples are scattered to their digit-reversed
// slots as they arrive), the two real-sequence spectra un-packed by Hermitian symmetry, and K running sums with the
// twiddle carried by recurrence (W_n^(2k) per pair, N2 / 2 multiplications: 1e-14).  Every sample is read exactly
// once, nothing but |X[k]|, k < K, is written; the transform itself is exact to rounding like any FFT.
constexpr int kPrunedThreads = 256;
constexpr int kMaxStages = 12;
struct PrunedPlan {
  int n1, n2, K, n_stages;
  int radix[kMaxStages];  // outermost first: n1 = radix[0] * radix[1] * ...
  // four-instance kernel: ceil(2^32 / d) for d = n1 / radix (butterflies per buffer) and d = M (sub-transform size) of each stage
  uint32_t magic_per[kMaxStages], magic_M[kMaxStages];
};

__device__ __forceinline__ double2 cmul(double2 a, double2 b) { return make_double2(fma(a.x, b.x, -(a.y * b.y)), fma(a.x, b.y, a.y * b.x)); }
__device__ __forceinline__ double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ double2 mul_mi(double2 a) { return make_double2(a.y, -a.x); }  // a * (-i)

// r-point DFT (forward, e^{-2 pi i jk / r}) of v[0..r) in place
template <int R> __device__ __forceinline__ void dft_small(double2 *v);
template <> __device__ __forceinline__ void dft_small<2>(double2 *v) {
  const double2 a = v[0], b = v[1];
  v[0] = cadd(a, b);
  v[1] = csub(a, b);
}
template <> __device__ __forceinline__ void dft_small<3>(double2 *v) {
  const double c = -0.5, s = 0.86602540378443864676;  // cos, sin of 2 pi / 3
  const double2 t1 = cadd(v[1], v[2]), t2 = csub(v[1], v[2]);
  const double2 m = make_double2(fma(c, t1.x, v[0].x), fma(c, t1.y, v[0].y));
  const double2 r = make_double2(s * t2.y, -s * t2.x);  // -i s t2
  v[0] = cadd(v[0], t1);
  v[1] = cadd(m, r);
  v[2] = csub(m, r);
}
template <> __device__ __forceinline__ void dft_small<4>(double2 *v) {
  const double2 a = cadd(v[0], v[2]), b = csub(v[0], v[2]), c = cadd(v[1], v[3]), d = mul_mi(csub(v[1], v[3]));
  v[0] = cadd(a, c);
  v[1] = cadd(b, d);
  v[2] = csub(a, c);
  v[3] = csub(b, d);
}
template <> __device__ __forceinline__ void dft_small<5>(double2 *v) {
  const double c1 = 0.30901699437494742410, c2 = -0.80901699437494742410;  // cos 2pi/5, cos 4pi/5
  const double s1 = 0.95105651629515357212, s2 = 0.58778525229247312917;   // sin 2pi/5, sin 4pi/5
  const double2 a1 = cadd(v[1], v[4]), b1 = csub(v[1], v[4]), a2 = cadd(v[2], v[3]), b2 = csub(v[2], v[3]);
  const double2 m1 = make_double2(fma(c1, a1.x, fma(c2, a2.x, v[0].x)), fma(c1, a1.y, fma(c2, a2.y, v[0].y)));
  const double2 m2 = make_double2(fma(c2, a1.x, fma(c1, a2.x, v[0].x)), fma(c2, a1.y, fma(c1, a2.y, v[0].y)));
  // -i (s1 b1 + s2 b2) and -i (s2 b1 - s1 b2)
  const double2 q1 = make_double2(fma(s1, b1.y, s2 * b2.y), -fma(s1, b1.x, s2 * b2.x));
  const double2 q2 = make_double2(fma(s2, b1.y, -(s1 * b2.y)), -fma(s2, b1.x, -(s1 * b2.x)));
  v[0] = cadd(v[0], cadd(a1, a2));
  v[1] = cadd(m1, q1);
  v[4] = csub(m1, q1);
  v[2] = cadd(m2, q2);
  v[3] = csub(m2, q2);
}

// one in-place decimation-in-time stage: sub-transforms of size M (already done, stored blockwise) are combined R at a time
template <int R> __device__ __forceinline__ void stage(double2 *buf, int n1, int M, const double2 *__restrict__ wtab, int tid) {
  const int step = n1 / (R * M);  // W_{RM}^j = W_n1^(j * step)
  for (int u = tid; u < n1 / R; u += kPrunedThreads) {
    const int blk = u / M, k1 = u - blk * M;
    double2 *p = buf + (size_t)blk * R * M + k1;
    double2 v[R];
    v[0] = p[0];
#pragma unroll
    for (int sgm = 1; sgm < R; sgm++) v[sgm] = (k1 == 0) ? p[sgm * M] : cmul(p[sgm * M], __ldg(wtab + (size_t)sgm * k1 * step));
    dft_small<R>(v);
#pragma unroll
    for (int sgm = 0; sgm < R; sgm++) p[sgm * M] = v[sgm];
  }
}

__global__ void __launch_bounds__(kPrunedThreads) pruned_dft_kernel(const double *__restrict__ x, long long B, long long ld,
                                                                    const __grid_constant__ PrunedPlan P,
                                                                    const double2 *__restrict__ wtab,   // W_n1^j, j < n1
                                                                    const double2 *__restrict__ wk,     // W_n^k,  k < K
                                                                    double *__restrict__ out) {
  extern __shared__ double2 sm2[];
  double2 *buf = sm2;              // [n1]  the packed pair, transformed in place
  double2 *acc = buf + P.n1;       // [K]   running sums X[k]
  double2 *tw = acc + P.K;         // [K]   W_n^(n2 k) of the current pair's even member
  unsigned short *perm = (unsigned short *)(tw + P.K);  // [n1] digit-reversed slot of sample n1
  const int tid = threadIdx.x;
  const long long b = blockIdx.x;
  const int n1 = P.n1, n2 = P.n2, K = P.K;
  for (int q = tid; q < n1; q += kPrunedThreads) {
    // sample q = r0 q0' + s0 goes to block s0 of size n1 / r0, and so on inwards
    int rem = q, span = n1, pos = 0;
    for (int st = 0; st < P.n_stages; st++) {
      const int r = P.radix[st];
      span /= r;
      pos += (rem % r) * span;
      rem /= r;
    }
    perm[q] = (unsigned short)pos;
  }
  for (int k = tid; k < K; k += kPrunedThreads) {
    acc[k] = make_double2(0.0, 0.0);
    tw[k] = make_double2(1.0, 0.0);
  }
  __syncthreads();
  const double *xb = x + b;
  for (int j = 0; j < n2; j += 2) {
    const bool has_odd = j + 1 < n2;
    for (int q = tid; q < n1; q += kPrunedThreads) {
      const double *p = xb + ((long long)n2 * q + j) * ld;
      buf[perm[q]] = make_double2(__ldcs(p), has_odd ? __ldcs(p + ld) : 0.0);
    }
    __syncthreads();
    int M = 1;
    for (int st = P.n_stages - 1; st >= 0; st--) {
      const int r = P.radix[st];
      if (r == 4) stage<4>(buf, n1, M, wtab, tid);
      else if (r == 5) stage<5>(buf, n1, M, wtab, tid);
      else if (r == 2) stage<2>(buf, n1, M, wtab, tid);
      else stage<3>(buf, n1, M, wtab, tid);
      M *= r;
      __syncthreads();
    }
    // un-pack the two real spectra (Z = E + i O:  E[k] = (Z[k] + conj Z[n1-k]) / 2,  O[k] = (Z[k] - conj Z[n1-k]) / 2i) and add
    // W^(j k) E[k] + W^((j+1) k) O[k]
    for (int k = tid; k < K; k += kPrunedThreads) {
      const double2 zk = buf[k], zm = buf[k == 0 ? 0 : n1 - k];
      const double2 E = make_double2(0.5 * (zk.x + zm.x), 0.5 * (zk.y - zm.y));
      const double2 O = make_double2(0.5 * (zk.y + zm.y), -0.5 * (zk.x - zm.x));
      const double2 w = __ldg(wk + k);
      const double2 t = tw[k];
      acc[k] = cadd(acc[k], cmul(t, cadd(E, cmul(w, O))));
      tw[k] = cmul(t, cmul(w, w));
    }
    __syncthreads();
  }
  for (int k = tid; k < K; k += kPrunedThreads) out[b * K + k] = hypot(acc[k].x, acc[k].y);
}


// ---- the same transform, four instances per block ------------------------------------------------------------------
// pruned_dft_kernel reads ONE double of every 64-byte DRAM burst of the time-major log (ncu: 5.9 bytes fetched per
// byte used, 8 % FP64 pipe, top stall long_scoreboard).  Here a block owns FOUR neighbouring instances, i.e. one whole
// 32-byte sector of every row it touches (the neighbouring block, running beside it, takes the other half of the
// burst), packs them as two complex sequences (instances 0/1 and 2/3), and walks the decimated sequences j < n2 one
// per pass: an N1-point FFT of both packed buffers, the four real spectra un-packed by Hermitian symmetry, and
// X_g[k] += W_n^(j k) Y_g,j[k mod N1] with the twiddle (shared by the four instances) carried by recurrence.  Y_j is
// N1-periodic in k, so N1 need not reach K; the plan still prefers N1 >= K (fewest passes).  The samples of pass
// j + 1 are fetched into registers while pass j is transformed, so the DRAM latency hides behind the butterflies.
constexpr int kG = 4;
constexpr int kGThreads = 512;
constexpr int kGSlots = 16;  // register-staged samples per thread and pass: kG * n1 <= kGSlots * kGThreads

// n / d for n < 2^14, d < 2^14 by one multiply-high: magic = ceil(2^32 / d) (d >= 2; d == 1 is handled by the caller)
__device__ __forceinline__ int div_magic(int n, uint32_t magic) { return (int)__umulhi((uint32_t)n, magic); }

// One decimation-in-time stage over n_buf buffers, sub-transforms of size M combined R at a time.  Two butterflies per
// thread and trip (independent loads in flight), index arithmetic by multiply-high, the R - 1 twiddles W_{RM}^(s k1)
// from ONE table read and powers of it; the first stage (M == 1) has no twiddles at all.
template <int R, int THREADS = 512>
__device__ __forceinline__ void stage_multi(double2 *buf, int n1, int n_buf, int M, uint32_t magic_per, uint32_t magic_M,
                                            const double2 *__restrict__ wtab, int tid) {
  const int per = n1 / R, total = n_buf * per;
  const int step = per / M;  // W_{RM}^j = W_n1^(j * step)
  for (int ub = tid; ub < total; ub += 2 * THREADS) {
    double2 v[2][R];
    double2 *p[2];
    bool on[2];
#pragma unroll
    for (int h = 0; h < 2; h++) {
      const int u = ub + h * THREADS;
      on[h] = u < total;
      const int uc = on[h] ? u : 0;
      const int b = div_magic(uc, magic_per), uu = uc - b * per;
      const int blk = (M == 1) ? uu : div_magic(uu, magic_M), k1 = uu - blk * M;
      p[h] = buf + b * n1 + blk * R * M + k1;
      double2 w1 = make_double2(1.0, 0.0);
      if (M > 1) w1 = __ldg(wtab + k1 * step);
#pragma unroll
      for (int sgm = 0; sgm < R; sgm++) v[h][sgm] = p[h][sgm * M];
      if (M > 1) {
        const double2 w2 = cmul(w1, w1);
        v[h][1] = cmul(v[h][1], w1);
        if (R > 2) v[h][2] = cmul(v[h][2], w2);
        if (R > 3) v[h][3] = cmul(v[h][3], cmul(w2, w1));
        if (R > 4) v[h][4] = cmul(v[h][4], cmul(w2, w2));
      }
    }
#pragma unroll
    for (int h = 0; h < 2; h++) {
      dft_small<R>(v[h]);
      if (on[h]) {
#pragma unroll
        for (int sgm = 0; sgm < R; sgm++) p[h][sgm * M] = v[h][sgm];
      }
    }
  }
}

__global__ void __launch_bounds__(kGThreads, 1) pruned_dft4_kernel(const double *__restrict__ x, long long B, long long ld,
                                                                   const __grid_constant__ PrunedPlan P,
                                                                   const double2 *__restrict__ wtab,   // W_n1^j, j < n1
                                                                   const double2 *__restrict__ wk,     // W_n^k,  k < K
                                                                   double *__restrict__ out) {
  extern __shared__ double2 sm2[];
  const int n1 = P.n1, n2 = P.n2, K = P.K;
  double2 *buf = sm2;                    // [2][n1]  instances (0,1) and (2,3), each packed re/im, transformed in place
  double2 *acc = buf + 2 * (size_t)n1;   // [kG][K]  running sums X_g[k]
  double2 *tw = acc + (size_t)kG * K;    // [K]      W_n^(j k) of the current pass
  const int tid = threadIdx.x;
  const long long b0 = (long long)blockIdx.x * kG;
  const int g = tid & (kG - 1), q0 = tid >> 2;
  const bool g_ok = b0 + g < B;
  // slot s of this thread: sample q = q0 + 128 s of instance g; its digit-reversed place in buf (in doubles, < 2^16), two
  // places per register, 0xffff = no such sample
  uint32_t dst2[kGSlots / 2];
#pragma unroll
  for (int s = 0; s < kGSlots; s++) {
    const int q = q0 + s * (kGThreads / kG);
    int rem = q, span = n1, pos = 0;
    for (int st = 0; st < P.n_stages; st++) {
      const int r = P.radix[st];
      span /= r;
      pos += (rem % r) * span;
      rem /= r;
    }
    const uint32_t d = (q < n1) ? (uint32_t)(2 * ((g >> 1) * n1 + pos) + (g & 1)) : 0xffffu;
    if (s & 1) dst2[s >> 1] |= d << 16;
    else dst2[s >> 1] = d;
  }
  auto dst = [&](int s) { return (dst2[s >> 1] >> ((s & 1) * 16)) & 0xffffu; };
  for (int k = tid; k < K; k += kGThreads) {
    tw[k] = make_double2(1.0, 0.0);
#pragma unroll
    for (int i = 0; i < kG; i++) acc[(size_t)i * K + k] = make_double2(0.0, 0.0);
  }
  const double *xg = x + b0 + g + (long long)n2 * q0 * ld;
  const long long slot_stride = (long long)(kGThreads / kG) * n2 * ld;
  double v[kGSlots];
#pragma unroll
  for (int s = 0; s < kGSlots; s++) v[s] = (g_ok && dst(s) != 0xffffu) ? __ldcs(xg + s * slot_stride) : 0.0;
  for (int j = 0; j < n2; j++) {
    double *bd = reinterpret_cast<double *>(buf);
#pragma unroll
    for (int s = 0; s < kGSlots; s++)
      if (dst(s) != 0xffffu) bd[dst(s)] = v[s];
    __syncthreads();
    if (j + 1 < n2) {
      const double *xn = xg + (long long)(j + 1) * ld;
#pragma unroll
      for (int s = 0; s < kGSlots; s++) v[s] = (g_ok && dst(s) != 0xffffu) ? __ldcs(xn + s * slot_stride) : 0.0;
    }
    int M = 1;
    for (int st = P.n_stages - 1; st >= 0; st--) {
      const int r = P.radix[st];
      if (r == 4) stage_multi<4>(buf, n1, 2, M, P.magic_per[st], P.magic_M[st], wtab, tid);
      else if (r == 5) stage_multi<5>(buf, n1, 2, M, P.magic_per[st], P.magic_M[st], wtab, tid);
      else if (r == 2) stage_multi<2>(buf, n1, 2, M, P.magic_per[st], P.magic_M[st], wtab, tid);
      else stage_multi<3>(buf, n1, 2, M, P.magic_per[st], P.magic_M[st], wtab, tid);
      M *= r;
      __syncthreads();
    }
    // Z = Y_a + i Y_b of two real sequences:  2 Y_a[k] = Z[k] + conj Z[-k],  2 Y_b[k] = -i (Z[k] - conj Z[-k])   (the 1/2 is applied at the end)
    for (int k = tid; k < K; k += kGThreads) {
      int km = k;
      if (K > n1) km = k % n1;
      const int kr = (km == 0) ? 0 : n1 - km;
      const double2 t = tw[k];
#pragma unroll
      for (int pr = 0; pr < 2; pr++) {
        const double2 zk = buf[(size_t)pr * n1 + km], zm = buf[(size_t)pr * n1 + kr];
        const double2 Ya = make_double2(zk.x + zm.x, zk.y - zm.y);
        const double2 Yb = make_double2(zk.y + zm.y, zm.x - zk.x);
        double2 *a = acc + (size_t)(2 * pr) * K + k;
        a[0] = cadd(a[0], cmul(t, Ya));
        a[K] = cadd(a[K], cmul(t, Yb));
      }
      tw[k] = cmul(t, __ldg(wk + k));
    }
    __syncthreads();
  }
  for (int k = tid; k < K; k += kGThreads)
#pragma unroll
    for (int i = 0; i < kG; i++)
      if (b0 + i < B) out[(b0 + i) * K + k] = 0.5 * hypot(acc[(size_t)i * K + k].x, acc[(size_t)i * K + k].y);
}

// n = n1 n2 for the four-instance kernel: n1 a product of 2, 3, 4, 5 with n1 <= n1_max; the smallest n1 >= K if there is
// one, else the largest
bool plan_grouped(long long n, long long K, int n1_max, PrunedPlan *P) {
  auto factor = [&](long long n1) {
    long long rem = n1;
    int st = 0, radix[kMaxStages];
    while (rem % 4 == 0 && st < kMaxStages) radix[st++] = 4, rem /= 4;
    while (rem % 2 == 0 && st < kMaxStages) radix[st++] = 2, rem /= 2;
    while (rem % 5 == 0 && st < kMaxStages) radix[st++] = 5, rem /= 5;
    while (rem % 3 == 0 && st < kMaxStages) radix[st++] = 3, rem /= 3;
    if (rem != 1) return false;
    P->n1 = (int)n1;
    P->n2 = (int)(n / n1);
    P->K = (int)K;
    P->n_stages = st;
    for (int i = 0; i < st; i++) P->radix[i] = radix[i];
    long long M = 1;
    auto magic = [](long long d) { return d <= 1 ? 0u : (uint32_t)((((unsigned long long)1 << 32) + d - 1) / d); };
    for (int i = st - 1; i >= 0; i--) {  // the kernel runs the stages innermost (M = 1) first
      P->magic_per[i] = magic(n1 / radix[i]);
      P->magic_M[i] = magic(M);
      M *= radix[i];
    }
    return true;
  };
  for (long long n1 = std::max<long long>(K, 8); n1 <= n1_max && n1 <= n; n1++)
    if (n % n1 == 0 && factor(n1)) return true;
  for (long long n1 = std::min<long long>(n1_max, n); n1 >= 8; n1--)
    if (n % n1 == 0 && factor(n1)) return true;
  return false;
}


// ---- pruned DFT through Chebyshev moments ---------------------------------------------------------------------------
// When K << n / 2 (PIMM keeps 1601 of 100 001 bins) most of the single-pass kernels' work is wasted on resolution the
// output never sees.  Cut the series into N1 blocks of N2 consecutive samples, t = N2 q + j.  Inside a block the kernel
// e^{-2 pi i k j / n} turns by at most phi = pi k (N2 - 1) / n (a fraction of a cycle to a few cycles), so with
// s = (2 j - (N2 - 1)) / (N2 - 1) in [-1, 1] the Jacobi-Anger expansion
//     e^{-i phi s} = J_0(phi) + 2 sum_{d >= 1} (-i)^d J_d(phi) T_d(s)
// converges super-exponentially once d > phi: D = 20 terms leave 1e-17 at phi = 2.5.  Hence, exactly up to that tail,
//     X[k] = e^{-i phi_k} sum_{d < D} eps_d (-i)^d J_d(phi_k) C_d[k mod N1],   C_d = FFT_N1( c[q][d] ),
//     c[q][d] = sum_{j < N2} T_d(s_j) x[N2 q + j]                              (the block's Chebyshev moments)
// i.e. D real sequences of N1 points per series instead of N2 of them (20 instead of 100 for PIMM), and the moments are
// a streaming reduction: one coalesced pass over the time-major log (cheb_moments_kernel, HBM-bound), after which the
// transforms run on a fifth of the data laid out per series (cheb_combine_kernel: two moments packed per complex FFT,
//     S[k] += alpha_p[k] Z_p[k] + beta_p[k] conj Z_p[-k],  alpha, beta = (-1)^p (a_2p -+ a_2p+1) / 2,  a_d = eps_d J_d(phi_k)
// ).  The tail bound is evaluated on the host (Miller recurrence in long double) and D chosen for 2e-17.
constexpr int kMomThreads = 256;
constexpr int kBgThreads = 128;
constexpr int kMomMaxD = 32;

struct MomentPlan {
  PrunedPlan fft;  // n1 (FFT length), n2 (block length), K, radices
  int D;           // moments per block (even)
};

// J_0 .. J_{count-1} at x by Miller's backward recurrence, normalised with J_0 + 2 sum J_2m = 1
void bessel_j(long double x, int count, std::vector<long double> &J) {
  J.assign((size_t)count, 0.0L);
  if (x == 0.0L) {
    J[0] = 1.0L;
    return;
  }
  const int start = count + 60 + (int)(2.0L * x);
  long double jp1 = 0.0L, jc = 1e-300L, norm = 0.0L;
  for (int d = start; d >= 1; d--) {
    const long double jm1 = (2.0L * d / x) * jc - jp1;  // J_{d-1}
    jp1 = jc;
    jc = jm1;
    if (d - 1 < count) J[(size_t)(d - 1)] = jc;
    if ((d - 1) % 2 == 0) norm += (d - 1 == 0 ? 1.0L : 2.0L) * jc;
    if (fabsl(jc) > 1e250L) {  // rescale
      const long double sc = 1e-250L;
      jc *= sc, jp1 *= sc, norm *= sc;
      for (int i = d - 1; i < count; i++) J[(size_t)i] *= sc;
    }
  }
  for (auto &v : J) v /= norm;
}

bool factor_radices(long long n1, PrunedPlan *P) {
  long long rem = n1;
  int st = 0;
  while (rem % 4 == 0 && st < kMaxStages) P->radix[st++] = 4, rem /= 4;
  while (rem % 2 == 0 && st < kMaxStages) P->radix[st++] = 2, rem /= 2;
  while (rem % 5 == 0 && st < kMaxStages) P->radix[st++] = 5, rem /= 5;
  while (rem % 3 == 0 && st < kMaxStages) P->radix[st++] = 3, rem /= 3;
  if (rem != 1) return false;
  P->n_stages = st;
  long long M = 1;
  auto magic = [](long long d) { return d <= 1 ? 0u : (uint32_t)((((unsigned long long)1 << 32) + d - 1) / d); };
  for (int i = st - 1; i >= 0; i--) {  // the kernels run the stages innermost (M = 1) first
    P->magic_per[i] = magic(n1 / P->radix[i]);
    P->magic_M[i] = magic(M);
    M *= P->radix[i];
  }
  return true;
}

// block length N2 | n (8 <= N2 <= 1024) with N1 = n / N2 a product of 2, 3, 4, 5 that fits shared memory, D <= min(32, N2 / 2):
// the plan with the least estimated work, or false
bool plan_moments(long long n, long long K, size_t max_smem, MomentPlan *out) {
  const long double pi = 3.14159265358979323846264338327950288L;
  double best = 0.0;
  bool found = false;
  long long only_n2 = 0;  // CMTJB_SPECTRUM_N2: pin the block length (tests)
  if (const char *c = std::getenv("CMTJB_SPECTRUM_N2")) only_n2 = std::atoll(c);
  for (long long n2 = 8; n2 <= 1024 && n2 * 16 <= n; n2++) {
    if (n % n2 || (only_n2 > 0 && n2 != only_n2)) continue;
    const long long n1 = n / n2;
    if (n1 < 16 || n1 > 32768) continue;
    MomentPlan M;
    if (!factor_radices(n1, &M.fft)) continue;
    const size_t smem = (2 * (size_t)n1 + (size_t)K) * sizeof(double2) + (size_t)n1 * sizeof(unsigned short) + 16;
    if (smem > max_smem || n1 >= 16384) continue;
    const long double phi = pi * (long double)(K - 1) * (long double)(n2 - 1) / (long double)n;
    std::vector<long double> J;
    bessel_j(phi, kMomMaxD + 48, J);
    int D = -1;
    long double tail = 0.0L;
    for (int d = kMomMaxD + 47; d >= 2; d--) {
      tail += 2.0L * fabsl(J[(size_t)d]);
      if (tail > 2e-17L) {
        D = d + 1;
        break;
      }
    }
    if (D < 0) D = 2;
    D = (D + 3) / 4 * 4;
    if (D > kMomMaxD || 2 * D > n2) continue;
    double lg = std::log2((double)n1);
    // measured on B200 (PIMM shape): the moments pass is bound by the bytes it moves (the log once, the moments written and
    // read back), the transforms by their butterflies at about twice the cost per unit
    const double cost = (double)n * (1.0 + 2.5 * D / (double)n2) + 2.0 * (0.5 * D) * (double)n1 * lg + 2.0 * D * (double)K;
    if (!found || cost < best) {
      best = cost;
      found = true;
      M.fft.n1 = (int)n1, M.fft.n2 = (int)n2, M.fft.K = (int)K;
      M.D = D;
      *out = M;
    }
  }
  return found;
}

// Host tables of a plan: T[(n2 + 1) / 2][D] (T_d at s_j, j < n2 / 2, then the middle sample of an odd block) and
// ab[D / 2][K] = (alpha_p[k], beta_p[k])
void moment_tables(long long n, const MomentPlan &M, std::vector<double> &T, std::vector<double2> &ab) {
  const long double pi = 3.14159265358979323846264338327950288L;
  const int n2 = M.fft.n2, D = M.D, K = M.fft.K;
  const int rows = (n2 + 1) / 2;
  T.resize((size_t)rows * D);
  for (int j = 0; j < rows; j++) {
    const long double sj = (long double)(2 * j - (n2 - 1)) / (long double)(n2 - 1);
    const long double th = acosl(sj);
    for (int d = 0; d < D; d++) T[(size_t)j * D + d] = (double)cosl((long double)d * th);
  }
  ab.resize((size_t)(D / 2) * K);
  std::vector<long double> J;
  for (int k = 0; k < K; k++) {
    const long double phi = pi * (long double)k * (long double)(n2 - 1) / (long double)n;
    bessel_j(phi, D, J);
    for (int p = 0; p < D / 2; p++) {
      const long double a0 = (2 * p == 0 ? 1.0L : 2.0L) * J[(size_t)(2 * p)], a1 = 2.0L * J[(size_t)(2 * p + 1)];
      const long double sg = (p % 2) ? -0.5L : 0.5L;
      ab[(size_t)p * K + k] = make_double2((double)(sg * (a0 - a1)), (double)(sg * (a0 + a1)));
    }
  }
}

// Moments of one block q of one series per thread: lane = series (rows of the time-major log are read coalesced), warp =
// block.  Samples j and N2 - 1 - j share |T_d| (T_d(-s) = (-1)^d T_d(s)): their sum feeds the even moments, their difference
// the odd ones, D / 2 + 1 FP64 instructions per sample.  mom[series][p][q] = (c[q][2p], c[q][2p+1]); the 16-byte halves of
// a sector come from neighbouring warps of one block and merge in L2.
template <int D>
__global__ void __launch_bounds__(kMomThreads, (D > 24 ? 2 : 3)) cheb_moments_kernel(const double *__restrict__ x, long long ld, long long b_lo,
                                                                      long long b_hi, int n1, int n2, int q0, int nq, int q_per_block,
                                                                      const double *__restrict__ Ttab, double2 *__restrict__ mom) {
  extern __shared__ double sT[];  // [(n2 + 1) / 2][D]
  const int rows = (n2 + 1) / 2;
  for (int i = threadIdx.x; i < rows * D; i += kMomThreads) sT[i] = Ttab[i];
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long b = b_lo + (long long)blockIdx.x * 32 + lane;
  if (b >= b_hi) return;
  const int half = n2 / 2;
  for (int it = 0; it < q_per_block; it += kMomThreads / 32) {
    const int q = blockIdx.y * q_per_block + it + warp;  // relative to the chunk: x's row 0 is the first sample of block q0
    if (q >= nq) break;
    const double *lo = x + (long long)n2 * q * ld + b;
    const double *hi = lo + (long long)(n2 - 1) * ld;
    double c[D];
#pragma unroll
    for (int d = 0; d < D; d++) c[d] = 0.0;
#pragma unroll 4
    for (int j = 0; j < half; j++) {
      const double a0 = __ldcs(lo), a1 = __ldcs(hi);
      lo += ld;
      hi -= ld;
      const double e = a0 + a1, o = a0 - a1;
      const double2 *t2 = reinterpret_cast<const double2 *>(sT + j * D);
#pragma unroll
      for (int d = 0; d < D; d += 2) {
        const double2 t = t2[d / 2];  // (T_d, T_{d+1}) at s_j, d even
        c[d] = fma(t.x, e, c[d]);
        c[d + 1] = fma(t.y, o, c[d + 1]);
      }
    }
    if (n2 & 1) {  // the middle sample, s = 0: T_d(0) = 0 for odd d
      const double a0 = __ldcs(lo);
#pragma unroll
      for (int d = 0; d < D; d += 2) c[d] = fma(sT[half * D + d], a0, c[d]);
    }
    double2 *m = mom + ((b - b_lo) * (D / 2)) * (long long)n1 + q0 + q;
#pragma unroll
    for (int p = 0; p < D / 2; p++) m[(long long)p * n1] = make_double2(c[2 * p], c[2 * p + 1]);
  }
}

// ---- the BACKGROUND form: reduces chunk c on a second stream while the time loop integrates chunk c + 1 ---------------
// EXPERIMENT (CMTJB_PIMM_OVERLAP=1; measured, not the default: see DESIGN.md section 7).  llgs_balanced_kernel keeps 16 warps
// on every SM and pins its own occupancy by claiming the SM's shared memory (launch_fixed), which leaves a quarter of the
// registers and no shared memory.  This kernel is sized for that hole: ONE block of 128 threads x <= 160 registers per SM (a persistent grid of one block per SM; a second block does not fit
// beside the time loop), no shared memory (the table is read through L1, one broadcast load per warp and entry), a
// grid-stride walk over the (32-series tile, block q) tasks with q fastest, so that the 16-byte halves of a moment sector
// still come from neighbouring warps.  Its FP64 instructions fill issue slots the time loop leaves idle (~21 % of them).
template <int D>
__global__ void __maxnreg__(160) cheb_moments_bg_kernel(const double *__restrict__ x, long long ld, long long B,
                                                                                      int n1, int n2, int q0, int nq,
                                                                                      const double *__restrict__ Ttab, double2 *__restrict__ mom) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long tiles = (B + 31) / 32, total = tiles * nq;
  const int half = n2 / 2;
  for (long long u = (long long)blockIdx.x * (kBgThreads / 32) + warp; u < total; u += (long long)gridDim.x * (kBgThreads / 32)) {
    const long long tile = u / nq;
    const int q = (int)(u - tile * nq);
    const long long b_raw = tile * 32 + lane;
    const long long b = b_raw < B ? b_raw : B - 1;
    const double *lo = x + (long long)n2 * q * ld + b;
    const double *hi = lo + (long long)(n2 - 1) * ld;
    double c[D];
#pragma unroll
    for (int d = 0; d < D; d++) c[d] = 0.0;
#pragma unroll 8
    for (int j = 0; j < half; j++) {
      const double a0 = __ldcs(lo), a1 = __ldcs(hi);
      lo += ld;
      hi -= ld;
      const double e = a0 + a1, o = a0 - a1;
      const double2 *t2 = reinterpret_cast<const double2 *>(Ttab + j * D);
#pragma unroll
      for (int d = 0; d < D; d += 2) {
        const double2 t = __ldg(t2 + d / 2);
        c[d] = fma(t.x, e, c[d]);
        c[d + 1] = fma(t.y, o, c[d + 1]);
      }
    }
    if (n2 & 1) {
      const double a0 = __ldcs(lo);
#pragma unroll
      for (int d = 0; d < D; d += 2) c[d] = fma(__ldg(Ttab + half * D + d), a0, c[d]);
    }
    if (b_raw < B) {
      double2 *m = mom + (b * (D / 2)) * (long long)n1 + q0 + q;
#pragma unroll
      for (int p = 0; p < D / 2; p++) m[(long long)p * n1] = make_double2(c[2 * p], c[2 * p + 1]);
    }
  }
}

// ---- the same reduction with the log staged through shared memory by asynchronous copies ------------------------------
// The register version above keeps at most a few KB per warp in flight and stalls on DRAM latency (ncu: 35 % of the HBM
// peak, long_scoreboard 24 per issue).  Here a warp owns a ring of kMomStages shared-memory stages; one stage holds up to
// S front rows j and their mirror rows N2 - 1 - j of a 32-series tile (256 bytes per row, 16 bytes per lane and
// cp.async, two rows per warp instruction, one commit group per stage), so a block of 8 warps keeps ~200 KB of loads in
// flight with no registers tied up.  Each warp is its own producer and consumer; warps never synchronise with each other.
// A warp walks an even number of consecutive blocks q and writes the moments of a pair (q, q + 1) as whole 32-byte sectors.
// (Tried first: one 256-byte cp.async.bulk per row, completion on an mbarrier -- the bulk-copy unit retires about one such
// operation per 35 cycles and SM, 2.1 TB/s in all, slower than the register version.)
constexpr int kMomStages = 4;

__device__ __forceinline__ void cp_async16(void *dst, const void *src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int D>
__global__ void __launch_bounds__(kMomThreads, 1) cheb_moments_async_kernel(const double *__restrict__ x, long long ld, int n1, int n2,
                                                                            int q0, int nq, int S, int q_per_warp,
                                                                            const double *__restrict__ Ttab, double2 *__restrict__ mom) {
  extern __shared__ __align__(128) unsigned char smraw[];
  const int rows_T = (n2 + 1) / 2;
  const int stage_rows = 2 * S + 1;
  double *sT = reinterpret_cast<double *>(smraw);                                            // [(n2 + 1) / 2][D]
  const size_t t_bytes = (((size_t)rows_T * D * sizeof(double)) + 127) / 128 * 128;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double *ring = reinterpret_cast<double *>(smraw + t_bytes) + (size_t)warp * kMomStages * stage_rows * 32;  // [stage][row][32]
  for (int i = threadIdx.x; i < rows_T * D; i += kMomThreads) sT[i] = Ttab[i];
  __syncthreads();
  const int half = n2 / 2;
  const int nst = (half + S - 1) / S;  // stages per block of samples
  const int q_begin = (blockIdx.y * (kMomThreads / 32) + warp) * q_per_warp;
  const int q_end = min(nq, q_begin + q_per_warp);  // block indices relative to the chunk (q0 is even: pairs stay sector-aligned)
  if (q_begin >= q_end) return;
  const int total = (q_end - q_begin) * nst;
  const double *xt = x + (long long)blockIdx.x * 32 + 2 * (lane & 15);  // this lane's 16 bytes of the tile's rows
  const int hw = lane >> 4;
  auto issue = [&](int t) {
    const int q = q_begin + t / nst, k = t - (t / nst) * nst;
    const int j0 = k * S, cnt = min(S, half - j0);
    const int rows = 2 * cnt + (((n2 & 1) && k == nst - 1) ? 1 : 0);
    double *dst = ring + (size_t)(t % kMomStages) * stage_rows * 32 + 2 * (lane & 15);
    const long long r0 = (long long)n2 * q;
    for (int ri = hw; ri < rows; ri += 2) {
      long long row;
      int slot;
      if (ri < cnt) row = r0 + j0 + ri, slot = ri;
      else if (ri < 2 * cnt) row = r0 + (n2 - 1 - (j0 + ri - cnt)), slot = S + ri - cnt;
      else row = r0 + half, slot = 2 * S;
      cp_async16(dst + slot * 32, xt + row * ld);
    }
  };
  for (int t = 0; t < kMomStages; t++) {
    if (t < total) issue(t);
    cp_async_commit();
  }
  double c[D], cprev[D];
#pragma unroll
  for (int d = 0; d < D; d++) c[d] = 0.0, cprev[d] = 0.0;
  double2 *mrow = mom + ((long long)blockIdx.x * 32 + lane) * (D / 2) * (long long)n1 + q0;
  for (int t = 0; t < total; t++) {
    const int q = q_begin + t / nst, k = t - (t / nst) * nst;
    const int j0 = k * S, cnt = min(S, half - j0);
    cp_async_wait<kMomStages - 1>();  // this lane's copies of stage t have landed ...
    __syncwarp();                     // ... and so have the other lanes'
    const double *st = ring + (size_t)(t % kMomStages) * stage_rows * 32 + lane;
#pragma unroll 2
    for (int jj = 0; jj < cnt; jj++) {
      const double a0 = st[jj * 32], a1 = st[(S + jj) * 32];
      const double e = a0 + a1, o = a0 - a1;
      const double2 *t2 = reinterpret_cast<const double2 *>(sT + (j0 + jj) * D);
#pragma unroll
      for (int d = 0; d < D; d += 2) {
        const double2 tt = t2[d / 2];
        c[d] = fma(tt.x, e, c[d]);
        c[d + 1] = fma(tt.y, o, c[d + 1]);
      }
    }
    if ((n2 & 1) && k == nst - 1) {
      const double a0 = st[2 * S * 32];
#pragma unroll
      for (int d = 0; d < D; d += 2) c[d] = fma(sT[half * D + d], a0, c[d]);
    }
    __syncwarp();  // every lane has read the stage: it may be refilled
    if (t + kMomStages < total) issue(t + kMomStages);
    cp_async_commit();
    if (k == nst - 1) {  // block q complete
      const bool second = ((q - q_begin) & 1) != 0;
      if (second) {
#pragma unroll
        for (int p = 0; p < D / 2; p++) {
          double2 *m = mrow + (long long)p * n1 + (q - 1);
          m[0] = make_double2(cprev[2 * p], cprev[2 * p + 1]);
          m[1] = make_double2(c[2 * p], c[2 * p + 1]);
        }
      } else if (q + 1 >= q_end) {
#pragma unroll
        for (int p = 0; p < D / 2; p++) mrow[(long long)p * n1 + q] = make_double2(c[2 * p], c[2 * p + 1]);
      }
#pragma unroll
      for (int d = 0; d < D; d++) cprev[d] = c[d], c[d] = 0.0;
    }
  }
  cp_async_wait<0>();
}

// One series per block: the D / 2 packed moment sequences are transformed two at a time in shared memory (mixed-radix, in
// place, loaded straight to their digit-reversed slots; two butterflies per thread and trip, index arithmetic by
// multiply-high, one twiddle read per butterfly) and folded into the K running sums.
__global__ void __launch_bounds__(kPrunedThreads, 3) cheb_combine_kernel(const double2 *__restrict__ mom, long long b_lo, int n_passes,
                                                                      const __grid_constant__ PrunedPlan P,
                                                                      const double2 *__restrict__ wtab,  // W_n1^j, j < n1
                                                                      const double2 *__restrict__ ab,    // [n_passes][K]
                                                                      double *__restrict__ out) {
  extern __shared__ double2 sm2[];
  const int n1 = P.n1, K = P.K;
  double2 *buf = sm2;                                       // [2][n1]
  double2 *acc = buf + 2 * n1;                              // [K]
  unsigned short *perm = (unsigned short *)(acc + K);       // [n1]
  const int tid = threadIdx.x;
  for (int q = tid; q < n1; q += kPrunedThreads) {
    int rem = q, span = n1, pos = 0;
    for (int st = 0; st < P.n_stages; st++) {
      const int r = P.radix[st];
      span /= r;
      pos += (rem % r) * span;
      rem /= r;
    }
    perm[q] = (unsigned short)pos;
  }
  for (int k = tid; k < K; k += kPrunedThreads) acc[k] = make_double2(0.0, 0.0);
  __syncthreads();
  const double2 *mb = mom + (long long)blockIdx.x * n_passes * n1;
  for (int p = 0; p < n_passes; p += 2) {  // n_passes is even (D is a multiple of 4)
    const double2 *z = mb + (long long)p * n1;
    for (int q = tid; q < n1; q += kPrunedThreads) {
      const int d = perm[q];
      buf[d] = __ldcs(z + q);
      buf[n1 + d] = __ldcs(z + n1 + q);
    }
    __syncthreads();
    int M = 1;
    for (int st = P.n_stages - 1; st >= 0; st--) {
      const int r = P.radix[st];
      if (r == 4) stage_multi<4, kPrunedThreads>(buf, n1, 2, M, P.magic_per[st], P.magic_M[st], wtab, tid);
      else if (r == 5) stage_multi<5, kPrunedThreads>(buf, n1, 2, M, P.magic_per[st], P.magic_M[st], wtab, tid);
      else if (r == 2) stage_multi<2, kPrunedThreads>(buf, n1, 2, M, P.magic_per[st], P.magic_M[st], wtab, tid);
      else stage_multi<3, kPrunedThreads>(buf, n1, 2, M, P.magic_per[st], P.magic_M[st], wtab, tid);
      M *= r;
      __syncthreads();
    }
    const double2 *abp = ab + (long long)p * K;
    for (int k = tid; k < K; k += kPrunedThreads) {
      const int km = (K > n1) ? k % n1 : k;
      const int kr = km == 0 ? 0 : n1 - km;
      double2 a = acc[k];
#pragma unroll
      for (int h = 0; h < 2; h++) {
        const double2 zk = buf[h * n1 + km], zm = buf[h * n1 + kr];
        const double2 c = __ldg(abp + (long long)h * K + k);
        a.x = fma(c.x, zk.x, fma(c.y, zm.x, a.x));
        a.y = fma(c.x, zk.y, fma(-c.y, zm.y, a.y));
      }
      acc[k] = a;
    }
    __syncthreads();
  }
  const long long b = b_lo + blockIdx.x;
  for (int k = tid; k < K; k += kPrunedThreads) out[b * K + k] = hypot(acc[k].x, acc[k].y);
}

template <int D>
cudaError_t launch_moments(const double *x, long long ld, long long b_lo, long long b_hi, int n1, int n2, int q0, int nq,
                           const double *Ttab, double2 *mom, int max_smem, bool allow_tma, cudaStream_t stream, bool background = false) {
  const int rows_T = (n2 + 1) / 2;
  long long full_tiles = 0;
  if (background) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    // the time loop beside it runs with the SM's carve-out at its maximum (its occupancy pin); an SM only changes its
    // carve-out when idle, so this kernel must ask for the same one or the two kernels exclude each other
    cudaFuncSetAttribute(cheb_moments_bg_kernel<D>, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
    cheb_moments_bg_kernel<D><<<(unsigned)sms, kBgThreads, 0, stream>>>(x + b_lo, ld, b_hi - b_lo, n1, n2, q0, nq, Ttab, mom);
    g_launches++;
    return cudaGetLastError();
  }
  // asynchronous-copy form: whole 32-series tiles whose rows start on 16-byte boundaries
  if (allow_tma && (ld % 2) == 0 && (reinterpret_cast<uintptr_t>(x + b_lo) % 16) == 0 && n2 >= 4 && (q0 % 2) == 0) {
    const size_t t_bytes = (((size_t)rows_T * D * sizeof(double)) + 127) / 128 * 128;
    const size_t bar_bytes = 0;
    const size_t avail = (size_t)max_smem > t_bytes + bar_bytes + 1024 ? (size_t)max_smem - t_bytes - bar_bytes - 1024 : 0;
    // per warp: kMomStages stages of (2 S + 1) rows of 256 bytes
    int S = (int)std::min<size_t>(15, (avail / ((size_t)(kMomThreads / 32) * kMomStages * 256) - 1) / 2);
    const int half = n2 / 2;
    if (S >= 2) {
      S = std::min(S, half);
      const int nst = (half + S - 1) / S;
      S = (half + nst - 1) / nst;  // even out the stages of a block
      const size_t smem = t_bytes + (size_t)(kMomThreads / 32) * kMomStages * (2 * S + 1) * 256 + bar_bytes;
      cudaError_t e = cudaFuncSetAttribute(cheb_moments_async_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return e;
      full_tiles = (b_hi - b_lo) / 32;
      if (full_tiles > 0) {
        // an even number of blocks per warp, enough warps for every SM: ~4 waves of 148 blocks at least when the slab allows
        int q_per_warp = 8;
        while (q_per_warp > 2 && full_tiles * ((nq + 8 * q_per_warp - 1) / (8 * q_per_warp)) < 4 * 148) q_per_warp -= 2;
        const dim3 grid((unsigned)full_tiles, (unsigned)((nq + 8 * q_per_warp - 1) / (8 * q_per_warp)));
        cheb_moments_async_kernel<D><<<grid, kMomThreads, smem, stream>>>(x + b_lo, ld, n1, n2, q0, nq, S, q_per_warp, Ttab, mom);
        e = cudaGetLastError();
        if (e != cudaSuccess) return e;
        g_launches++;
      }
    }
  }
  const long long b_rest = b_lo + full_tiles * 32;
  if (b_rest >= b_hi) return cudaSuccess;
  const size_t smem = (size_t)rows_T * D * sizeof(double);
  cudaError_t e = cudaFuncSetAttribute(cheb_moments_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  const int per_block = 4 * (kMomThreads / 32);  // four blocks of samples per warp: the table load is amortised over 32 of them
  const dim3 grid((unsigned)((b_hi - b_rest + 31) / 32), (unsigned)((nq + per_block - 1) / per_block));
  cheb_moments_kernel<D><<<grid, kMomThreads, smem, stream>>>(x, ld, b_rest, b_hi, n1, n2, q0, nq, per_block, Ttab,
                                                              mom + full_tiles * 32 * (D / 2) * (long long)n1);
  g_launches++;
  return cudaGetLastError();
}

cudaError_t launch_moments_D(int D, const double *x, long long ld, long long b_lo, long long b_hi, int n1, int n2, int q0, int nq,
                             const double *Ttab, double2 *mom, int max_smem, bool allow_tma, cudaStream_t stream, bool background = false) {
  switch (D) {
  case 4: return launch_moments<4>(x, ld, b_lo, b_hi, n1, n2, q0, nq, Ttab, mom, max_smem, allow_tma, stream, background);
  case 8: return launch_moments<8>(x, ld, b_lo, b_hi, n1, n2, q0, nq, Ttab, mom, max_smem, allow_tma, stream, background);
  case 12: return launch_moments<12>(x, ld, b_lo, b_hi, n1, n2, q0, nq, Ttab, mom, max_smem, allow_tma, stream, background);
  case 16: return launch_moments<16>(x, ld, b_lo, b_hi, n1, n2, q0, nq, Ttab, mom, max_smem, allow_tma, stream, background);
  case 20: return launch_moments<20>(x, ld, b_lo, b_hi, n1, n2, q0, nq, Ttab, mom, max_smem, allow_tma, stream, background);
  case 24: return launch_moments<24>(x, ld, b_lo, b_hi, n1, n2, q0, nq, Ttab, mom, max_smem, allow_tma, stream, background);
  case 28: return launch_moments<28>(x, ld, b_lo, b_hi, n1, n2, q0, nq, Ttab, mom, max_smem, allow_tma, stream, background);
  default: return launch_moments<32>(x, ld, b_lo, b_hi, n1, n2, q0, nq, Ttab, mom, max_smem, allow_tma, stream, background);
  }
}

// n = n1 n2 with K <= n1 <= n1_max and n1 a product of 2, 3, 4, 5: the smallest such n1, or false
bool plan_pruned(long long n, long long K, int n1_max, PrunedPlan *P) {
  for (long long n1 = std::max<long long>(K, 2); n1 <= n1_max && n1 <= n; n1++) {
    if (n % n1) continue;
    long long rem = n1;
    int st = 0, radix[kMaxStages];
    while (rem % 4 == 0 && st < kMaxStages) radix[st++] = 4, rem /= 4;
    while (rem % 2 == 0 && st < kMaxStages) radix[st++] = 2, rem /= 2;
    while (rem % 5 == 0 && st < kMaxStages) radix[st++] = 5, rem /= 5;
    while (rem % 3 == 0 && st < kMaxStages) radix[st++] = 3, rem /= 3;
    if (rem != 1) continue;
    P->n1 = (int)n1;
    P->n2 = (int)(n / n1);
    P->K = (int)K;
    P->n_stages = st;
    for (int i = 0; i < st; i++) P->radix[i] = radix[i];
    return true;
  }
  return false;
}

// e^{-2 pi i j / m}, j < count, with the argument reduced exactly (j mod m in integers, octant symmetry left to sincospi)
void unit_roots(long long m, long long count, std::vector<double2> &w) {
  w.resize((size_t)count);
  for (long long j = 0; j < count; j++) {
    const long double a = -2.0L * 3.14159265358979323846264338327950288L * (long double)(j % m) / (long double)m;
    w[(size_t)j] = make_double2((double)cosl(a), (double)sinl(a));
  }
}

}  // namespace
}  // namespace cmtjb

using namespace cmtjb;

extern "C" int cmtjb_rfft_abs_device_ws(const double *x, int64_t n, int64_t B, int64_t ld, int64_t K, double *out, int device,
                                        void *stream_, void *workspace, int64_t workspace_bytes);

namespace {
// device tables of the moment plans used so far (a sweep procedure calls with the same shape again and again)
struct MomentTables {
  int device;
  long long n, K, pinned_n2;
  bool found;
  MomentPlan M;
  double *d_T;
  double2 *d_ab, *d_w;
};
std::mutex g_mom_mu;
std::vector<MomentTables> g_mom_tables;

// plan + tables for (n, K) on `device`, built on first use; nullptr when the moment path does not apply
const MomentTables *moment_tables_for(long long n, long long K, int device, int max_smem, cudaError_t *err) {
  *err = cudaSuccess;
  long long pinned = 0;
  if (const char *c = std::getenv("CMTJB_SPECTRUM_N2")) pinned = std::atoll(c);
  std::lock_guard<std::mutex> lock(g_mom_mu);
  for (const MomentTables &t : g_mom_tables)
    if (t.device == device && t.n == n && t.K == K && t.pinned_n2 == pinned) return t.found ? &t : nullptr;
  MomentTables t;
  std::memset(&t, 0, sizeof(t));
  t.device = device, t.n = n, t.K = K, t.pinned_n2 = pinned;
  t.found = n >= 512 && K >= 1 && plan_moments(n, K, (size_t)max_smem, &t.M);
  if (t.found) {
    std::vector<double> T;
    std::vector<double2> ab, wt;
    moment_tables(n, t.M, T, ab);
    unit_roots(t.M.fft.n1, t.M.fft.n1, wt);
    cudaError_t e = cudaMalloc((void **)&t.d_T, T.size() * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc((void **)&t.d_ab, ab.size() * sizeof(double2));
    if (e == cudaSuccess) e = cudaMalloc((void **)&t.d_w, wt.size() * sizeof(double2));
    if (e == cudaSuccess) e = cudaMemcpy(t.d_T, T.data(), T.size() * sizeof(double), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(t.d_ab, ab.data(), ab.size() * sizeof(double2), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(t.d_w, wt.data(), wt.size() * sizeof(double2), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
      cudaFree(t.d_T), cudaFree(t.d_ab), cudaFree(t.d_w);
      *err = e;
      return nullptr;
    }
  }
  if (g_mom_tables.size() >= 8) {  // oldest out; its kernels may still be running
    cudaDeviceSynchronize();
    cudaFree(g_mom_tables.front().d_T), cudaFree(g_mom_tables.front().d_ab), cudaFree(g_mom_tables.front().d_w);
    g_mom_tables.erase(g_mom_tables.begin());
  }
  g_mom_tables.reserve(9);  // pointers into the vector stay valid while at most 8 + 1 entries live
  g_mom_tables.push_back(t);
  return t.found ? &g_mom_tables.back() : nullptr;
}

bool moments_wanted() {
  const char *force = std::getenv("CMTJB_SPECTRUM");
  return !force || force[0] == 'm' || force[0] == '\0';
}
}  // namespace

extern "C" int cmtjb_rfft_abs_workspace(int64_t n, int64_t K, int device, int64_t *bytes_per_series) {
  if (!bytes_per_series || n <= 0 || K < 0) return fail(CMTJB_EINVAL, "bad argument");
  *bytes_per_series = 0;
  if (K == 0 || !moments_wanted()) return CMTJB_OK;
  CUDA_TRY(cudaSetDevice(device));
  int max_smem = 0;
  cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
  cudaError_t e;
  const MomentTables *t = moment_tables_for(n, K, device, max_smem, &e);
  if (e != cudaSuccess) return cuda_fail(e, "moment tables");
  if (t) *bytes_per_series = (int64_t)(t->M.D / 2) * t->M.fft.n1 * (int64_t)sizeof(double2);
  return CMTJB_OK;
}

extern "C" int cmtjb_rfft_abs_device(const double *x, int64_t n, int64_t B, int64_t ld, int64_t K, double *out, int device,
                                     void *stream_) {
  return cmtjb_rfft_abs_device_ws(x, n, B, ld, K, out, device, stream_, nullptr, 0);
}

// ---- the same spectrum, streamed: the series arrive in time chunks and are never resident as a whole --------------------
extern "C" int cmtjb_spectrum_stream_plan(int64_t n, int64_t K, int device, int64_t *block_len, int64_t *bytes_per_series) {
  if (!block_len || !bytes_per_series || n <= 0 || K < 0) return fail(CMTJB_EINVAL, "bad argument");
  *block_len = 0, *bytes_per_series = 0;
  if (K == 0 || !moments_wanted()) return CMTJB_OK;
  CUDA_TRY(cudaSetDevice(device));
  int max_smem = 0;
  cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
  cudaError_t e;
  const MomentTables *t = moment_tables_for(n, K, device, max_smem, &e);
  if (e != cudaSuccess) return cuda_fail(e, "moment tables");
  if (t) {
    *block_len = t->M.fft.n2;
    *bytes_per_series = (int64_t)(t->M.D / 2) * t->M.fft.n1 * (int64_t)sizeof(double2);
  }
  return CMTJB_OK;
}

extern "C" int cmtjb_spectrum_stream_moments(const double *x_chunk, int64_t ld, int64_t B, int64_t n, int64_t K, int64_t q0,
                                             int64_t nq, void *moments, int32_t background, int device, void *stream_) {
  if (!x_chunk || !moments || B <= 0 || ld < B || n <= 0 || K <= 0 || q0 < 0 || nq < 0) return fail(CMTJB_EINVAL, "bad argument");
  if (nq == 0) return CMTJB_OK;
  CUDA_TRY(cudaSetDevice(device));
  int max_smem = 0;
  cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
  cudaError_t e;
  const MomentTables *mt = moment_tables_for(n, K, device, max_smem, &e);
  if (e != cudaSuccess) return cuda_fail(e, "moment tables");
  if (!mt) return fail(CMTJB_ENOTSUP, "no streaming plan for this (n, K): see cmtjb_spectrum_stream_plan");
  const MomentPlan &M = mt->M;
  if (q0 + nq > M.fft.n1) return fail(CMTJB_EINVAL, "block range beyond the series");
  if ((reinterpret_cast<uintptr_t>(moments) % 32) != 0) return fail(CMTJB_EINVAL, "moments must be 32-byte aligned");
  const char *tma_env = std::getenv("CMTJB_SPECTRUM_TMA");
  const bool allow_tma = !(tma_env && tma_env[0] == '0');
  e = launch_moments_D(M.D, x_chunk, ld, 0, B, M.fft.n1, M.fft.n2, (int)q0, (int)nq, mt->d_T, reinterpret_cast<double2 *>(moments), max_smem,
                       allow_tma, (cudaStream_t)stream_, background != 0);
  if (e != cudaSuccess) return cuda_fail(e, "cheb_moments_kernel (streamed)");
  return CMTJB_OK;
}

extern "C" int cmtjb_spectrum_stream_finish(const void *moments, int64_t n, int64_t B, int64_t K, double *out, int device, void *stream_) {
  if (!moments || !out || B <= 0 || n <= 0 || K <= 0) return fail(CMTJB_EINVAL, "bad argument");
  CUDA_TRY(cudaSetDevice(device));
  int max_smem = 0;
  cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
  cudaError_t e;
  const MomentTables *mt = moment_tables_for(n, K, device, max_smem, &e);
  if (e != cudaSuccess) return cuda_fail(e, "moment tables");
  if (!mt) return fail(CMTJB_ENOTSUP, "no streaming plan for this (n, K): see cmtjb_spectrum_stream_plan");
  const MomentPlan &M = mt->M;
  const int n1 = M.fft.n1;
  const size_t smem_c = (2 * (size_t)n1 + (size_t)K) * sizeof(double2) + (size_t)n1 * sizeof(unsigned short) + 16;
  e = cudaFuncSetAttribute(cheb_combine_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_c);
  const long long step = 1 << 20;  // grid.x limit is 2^31 - 1; keep launches modest
  for (long long b_lo = 0; b_lo < B && e == cudaSuccess; b_lo += step) {
    const long long cnt = std::min<long long>(step, B - b_lo);
    cheb_combine_kernel<<<(unsigned)cnt, kPrunedThreads, smem_c, (cudaStream_t)stream_>>>(
        reinterpret_cast<const double2 *>(moments) + b_lo * (long long)(M.D / 2) * n1, b_lo, M.D / 2, M.fft, mt->d_w, mt->d_ab, out);
    g_launches++;
    e = cudaGetLastError();
  }
  if (e != cudaSuccess) return cuda_fail(e, "cheb_combine_kernel (streamed)");
  return CMTJB_OK;
}

extern "C" int cmtjb_rfft_abs_device_ws(const double *x, int64_t n, int64_t B, int64_t ld, int64_t K, double *out, int device,
                                        void *stream_, void *workspace, int64_t workspace_bytes) {
  if (!x || !out || n <= 0 || B <= 0 || ld < B || K < 0 || K > n / 2 + 1) return fail(CMTJB_EINVAL, "bad argument");
  if (K == 0) return CMTJB_OK;
  if (n > (1LL << 30) || ld > (1LL << 30)) return fail(CMTJB_ENOTSUP, "series too long for the 32-bit cuFFT plan interface");
  CUDA_TRY(cudaSetDevice(device));
  cudaStream_t stream = (cudaStream_t)stream_;
  if (moments_wanted()) {
    // strongly pruned spectra (K << n / 2): Chebyshev moments of blocks of consecutive samples, then short transforms of
    // the moments (see cheb_moments_kernel).  CMTJB_SPECTRUM=moments or unset: used where a plan exists; any other value skips it.
    int max_smem = 0;
    cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
    cudaError_t e;
    const MomentTables *mt = moment_tables_for(n, K, device, max_smem, &e);
    if (e != cudaSuccess) return cuda_fail(e, "moment tables");
    if (mt) {
      const MomentPlan &M = mt->M;
      const int n1 = M.fft.n1, n2 = M.fft.n2, D = M.D, Pn = D / 2;
      const char *tma_env = std::getenv("CMTJB_SPECTRUM_TMA");  // =0: register-staged moments kernel only
      const bool allow_tma = !(tma_env && tma_env[0] == '0');
      // slab of series whose packed moments fit the caller's workspace, else an allocation of this call
      // (CMTJB_SPECTRUM_SLAB_MB; default: a quarter of the free memory, <= 2 GiB)
      const size_t per_series = (size_t)Pn * n1 * sizeof(double2);
      double2 *d_mom = reinterpret_cast<double2 *>(workspace);
      size_t budget = (size_t)std::max<int64_t>(0, workspace_bytes);
      const bool own = !workspace || budget < per_series || (reinterpret_cast<uintptr_t>(workspace) % 32) != 0;
      if (own) {
        size_t free_b = 0, total_b = 0;
        cudaMemGetInfo(&free_b, &total_b);
        budget = std::min<size_t>(free_b / 4, (size_t)2 << 30);
        if (const char *c = std::getenv("CMTJB_SPECTRUM_SLAB_MB")) budget = (size_t)std::max(1, std::atoi(c)) << 20;
      }
      long long slab = (long long)std::max<size_t>(1, budget / per_series);
      if (slab >= 32) slab = slab / 32 * 32;
      slab = std::min<long long>(slab, B);
      if (own) {
        d_mom = nullptr;
        e = cudaMallocAsync((void **)&d_mom, (size_t)slab * per_series, stream);
        if (e != cudaSuccess) return cuda_fail(e, "cudaMallocAsync(moment slab)");
      }
      const double *d_T = mt->d_T;
      const size_t smem_c = (2 * (size_t)n1 + (size_t)K) * sizeof(double2) + (size_t)n1 * sizeof(unsigned short) + 16;
      e = cudaFuncSetAttribute(cheb_combine_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_c);
      for (long long b_lo = 0; b_lo < B && e == cudaSuccess; b_lo += slab) {
        const long long b_hi = std::min<long long>(B, b_lo + slab);
        e = launch_moments_D(D, x, ld, b_lo, b_hi, n1, n2, 0, n1, d_T, d_mom, max_smem, allow_tma, stream);
        if (e != cudaSuccess) break;
        cheb_combine_kernel<<<(unsigned)(b_hi - b_lo), kPrunedThreads, smem_c, stream>>>(d_mom, b_lo, Pn, M.fft, mt->d_w, mt->d_ab, out);
        g_launches++;
        e = cudaGetLastError();
      }
      if (own && d_mom) cudaFreeAsync(d_mom, stream);
      if (e != cudaSuccess) return cuda_fail(e, "cheb_moments_kernel / cheb_combine_kernel");
      return CMTJB_OK;
    }
  }
  {
    // single pass over the buffer when n factors suitably: four instances per block (pruned_dft4_kernel) when its shared
    // memory fits, else one (pruned_dft_kernel); CMTJB_SPECTRUM=cufft forces the library path, =single the one-instance kernel
    PrunedPlan P;
    const char *force = std::getenv("CMTJB_SPECTRUM");
    const bool want = !(force && force[0] == 'c');
    const bool want_grouped = !(force && force[0] == 's');
    int max_smem = 0;
    cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
    if (force && force[0] == 'm') force = nullptr;
    bool grouped = false;
    size_t smem = 0;
    if (want && want_grouped && n >= 64 && plan_grouped(n, K, kGSlots * kGThreads / kG, &P) && (long long)P.n1 * 8 >= K) {
      smem = (2 * (size_t)P.n1 + (kG + 1) * (size_t)P.K) * sizeof(double2);
      grouped = smem <= (size_t)max_smem;
    }
    bool single = false;
    if (want && !grouped && n >= 64 && plan_pruned(n, K, 8192, &P)) {
      smem = ((size_t)P.n1 + 2 * (size_t)P.K) * sizeof(double2) + (size_t)P.n1 * sizeof(unsigned short) + 16;
      single = smem <= (size_t)max_smem && P.n1 < 65536;
    }
    if (grouped || single) {
      std::vector<double2> wt, wkv;
      unit_roots(P.n1, P.n1, wt);
      unit_roots(n, K, wkv);
      double2 *d_w = nullptr;
      cudaError_t e = cudaMallocAsync((void **)&d_w, (wt.size() + wkv.size()) * sizeof(double2), stream);
      if (e != cudaSuccess) return cuda_fail(e, "cudaMallocAsync(twiddles)");
      e = cudaMemcpyAsync(d_w, wt.data(), wt.size() * sizeof(double2), cudaMemcpyHostToDevice, stream);
      if (e == cudaSuccess) e = cudaMemcpyAsync(d_w + wt.size(), wkv.data(), wkv.size() * sizeof(double2), cudaMemcpyHostToDevice, stream);
      if (e == cudaSuccess) e = cudaStreamSynchronize(stream);  // the host vectors go out of scope
      if (e == cudaSuccess)
        e = grouped ? cudaFuncSetAttribute(pruned_dft4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
                    : cudaFuncSetAttribute(pruned_dft_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e == cudaSuccess) {
        if (grouped) pruned_dft4_kernel<<<(unsigned)((B + kG - 1) / kG), kGThreads, smem, stream>>>(x, B, ld, P, d_w, d_w + wt.size(), out);
        else pruned_dft_kernel<<<(unsigned)B, kPrunedThreads, smem, stream>>>(x, B, ld, P, d_w, d_w + wt.size(), out);
        g_launches++;
        e = cudaGetLastError();
      }
      cudaFreeAsync(d_w, stream);
      if (e != cudaSuccess) return cuda_fail(e, grouped ? "pruned_dft4_kernel" : "pruned_dft_kernel");
      return CMTJB_OK;
    }
  }
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
