// Longest-first scheduling of the solver's sample queue (K1).
//
// The lanes of gync_loglik_tm_kernel pull samples from a queue; a batch ends when the last-started solves finish, and
// with ~5 solves per lane whose step counts spread 810 .. 7981 (sigma 22 %) that tail is ~10 % of the kernel time
// (DESIGN.md §8).  Handing out the EXPENSIVE samples first shrinks it.  The cost of a sample is not known in advance, so
// the library learns it: after every batch a ridge regression of the measured step counts on the log-parameters (117
// features) is refitted on the device, and the next batch is ordered by its prediction (R^2 ~ 0.6 out of sample on the
// near set).  Scheduling only: every sample's arithmetic is independent of the lane and of the order it runs in, so
// results are bit-identical with and without it (tests: permutation invariance; GYNC_NO_SCHED=1 switches it off).
// Everything is stream-ordered on the device; nothing synchronises with the host.
#pragma once
#include <cstdint>

#define GYNC_SCHED_NF 117            /* 1 + 116 log-parameters */
#define GYNC_SCHED_NB 1024           /* cost buckets of the counting sort */
#define GYNC_SCHED_MAXS 4096         /* samples used for a refit (117 unknowns) */

struct GyncSchedModel {              // device-resident
  double beta[GYNC_SCHED_NF];
  double ref[GYNC_SCHED_NF];         // feature offset: log|x| of the first training sample (conditioning)
  double ref_new[GYNC_SCHED_NF];     // offset of the fit in progress: becomes `ref` only together with its `beta`
};

__device__ __forceinline__ double gync_sched_feature(double x) { return log(fmax(fabs(x), 1e-300)); }
// feature relative to the offset, clamped: an exact zero (log -> -690) must not become a leverage point of the fit
__device__ __forceinline__ double gync_sched_rel(double x, double ref) {
  const double v = gync_sched_feature(x) - ref;
  return (v == v) ? fmin(fmax(v, -20.0), 20.0) : 0.0;
}

// predicted cost per sample + its range (as order-preserving integers: predictions are made non-negative)
__global__ void __launch_bounds__(256)
gync_sched_predict_kernel(const double* __restrict__ X, int64_t N, const GyncSchedModel* __restrict__ mdl,
                          double* __restrict__ pred, unsigned long long* __restrict__ minmax) {
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  double p = 0.0;
  if (n < N) {
    p = mdl->beta[0];
    double chk = 0.0;
    for (int f = 1; f < GYNC_SCHED_NF; ++f) {
      const double x = __ldg(X + n + N * (f - 1));
      chk += x * 0.0;                              // NaN / Inf anywhere -> NaN
      p = fma(mdl->beta[f], gync_sched_rel(x, mdl->ref[f]), p);
    }
    // samples that fail at load time (non-finite input, e.g. Metropolis proposals outside the prior) cost nothing:
    // they go last, where they fill the tail for free
    if (!(chk == 0.0) || !(p == p)) p = 0.0;
    p = (p > 0.0) ? fmin(p, 1e300) : 0.0;          // non-negative (and never -0.0): the bit patterns order like the values
    pred[n] = p;
  }
  // block range, then two atomics per block
  __shared__ double smn[8], smx[8];
  double mn = (n < N) ? p : 1e300, mx = (n < N) ? p : 0.0;
  for (int o = 16; o > 0; o >>= 1) {
    mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
  if ((threadIdx.x & 31) == 0) { smn[threadIdx.x >> 5] = mn; smx[threadIdx.x >> 5] = mx; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) { mn = fmin(mn, smn[w]); mx = fmax(mx, smx[w]); }
    atomicMin(minmax, (unsigned long long)__double_as_longlong(mn));
    atomicMax(minmax + 1, (unsigned long long)__double_as_longlong(mx));
  }
}

__device__ __forceinline__ int gync_sched_bucket(double p, const unsigned long long* minmax) {
  const double mn = __longlong_as_double((long long)minmax[0]), mx = __longlong_as_double((long long)minmax[1]);
  const double w = mx - mn;
  if (!(w > 0.0)) return 0;
  const int b = (int)((mx - p) / w * (double)(GYNC_SCHED_NB - 1));      // bucket 0 = the most expensive
  return min(max(b, 0), GYNC_SCHED_NB - 1);
}

__global__ void __launch_bounds__(256)
gync_sched_hist_kernel(const double* __restrict__ pred, int64_t N, const unsigned long long* __restrict__ minmax, int* __restrict__ hist) {
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n < N) atomicAdd(hist + gync_sched_bucket(pred[n], minmax), 1);
}

// exclusive scan of the bucket counts (one CTA of GYNC_SCHED_NB threads) -> cursors
__global__ void __launch_bounds__(GYNC_SCHED_NB) gync_sched_scan_kernel(int* __restrict__ hist) {
  __shared__ int s[GYNC_SCHED_NB];
  const int t = threadIdx.x;
  s[t] = hist[t];
  __syncthreads();
  for (int o = 1; o < GYNC_SCHED_NB; o <<= 1) {
    const int v = (t >= o) ? s[t - o] : 0;
    __syncthreads();
    s[t] += v;
    __syncthreads();
  }
  hist[t] = s[t] - hist[t];
}

__global__ void __launch_bounds__(256)
gync_sched_scatter_kernel(const double* __restrict__ pred, int64_t N, const unsigned long long* __restrict__ minmax,
                          int* __restrict__ cursor, int* __restrict__ order) {
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n < N) order[atomicAdd(cursor + gync_sched_bucket(pred[n], minmax), 1)] = (int)n;
}

// ---- refit: features of a strided subsample, normal equations, ridge solve -----------------------------------
// Phi[s][f] (S x 117, sample-major), target[s] = step attempts of sample n_s = s * stride
__global__ void __launch_bounds__(128)
gync_sched_features_kernel(const double* __restrict__ X, int64_t N, int S, int64_t stride, const int* __restrict__ nsteps /* N x 2 */,
                           double* __restrict__ Phi, double* __restrict__ target, GyncSchedModel* __restrict__ mdl) {
  const int s = blockIdx.x;
  const int64_t n = (int64_t)s * stride;
  // a sample with a non-finite entry never runs a step: it carries no information about the cost of the others and is
  // left out of the fit (an all-zero row)
  double chk = 0.0;
  for (int f = threadIdx.x; f < GYNC_SCHED_NF - 1; f += blockDim.x) chk += __ldg(X + n + N * f) * 0.0;
  const int bad = __syncthreads_or(!(chk == 0.0));
  for (int f = threadIdx.x; f < GYNC_SCHED_NF; f += blockDim.x) {
    double v = bad ? 0.0 : 1.0;
    if (f > 0) {
      const double r = gync_sched_feature(__ldg(X + N * (f - 1)));          // sample 0 is the offset
      const double rr = (r == r) ? r : 0.0;
      v = bad ? 0.0 : gync_sched_rel(__ldg(X + n + N * (f - 1)), rr);
      if (s == 0) mdl->ref_new[f] = rr;
    } else if (s == 0) {
      mdl->ref_new[0] = 0.0;
    }
    Phi[(size_t)s * GYNC_SCHED_NF + f] = v;
  }
  if (threadIdx.x == 0) target[s] = bad ? 0.0 : (double)(nsteps[n] + nsteps[n + N]);
}

// G[i][j] = sum_s Phi[s][i] Phi[s][j],  b[i] = sum_s Phi[s][i] target[s]   (one CTA per row i, fixed order)
__global__ void __launch_bounds__(128)
gync_sched_normal_kernel(const double* __restrict__ Phi, const double* __restrict__ target, int S, double* __restrict__ G,
                         double* __restrict__ b) {
  const int i = blockIdx.x, j = threadIdx.x;
  if (j > GYNC_SCHED_NF) return;
  double acc = 0.0;
  for (int s = 0; s < S; ++s) {
    const double pi = Phi[(size_t)s * GYNC_SCHED_NF + i];
    acc = fma(pi, (j < GYNC_SCHED_NF) ? Phi[(size_t)s * GYNC_SCHED_NF + j] : target[s], acc);
  }
  if (j < GYNC_SCHED_NF) G[i * GYNC_SCHED_NF + j] = acc;
  else b[i] = acc;
}

// (G + lambda diag) beta = b by Cholesky in shared memory (one CTA); on a non-positive pivot the old model is kept
__global__ void __launch_bounds__(128)
gync_sched_fit_kernel(const double* __restrict__ G, const double* __restrict__ b, GyncSchedModel* __restrict__ mdl) {
  extern __shared__ double sm_fit[];                      // n x (n+1) matrix + rhs
  const int n = GYNC_SCHED_NF, ld = GYNC_SCHED_NF + 1;
  double* A = sm_fit;
  double* x = sm_fit + n * ld;
  __shared__ int ok;
  const int t = threadIdx.x;
  double tr = 0.0;
  for (int i = 0; i < n; ++i) tr += G[i * n + i];
  const double lam = 1e-8 * tr / n + 1e-12;
  for (int e = t; e < n * n; e += blockDim.x) A[(e / n) * ld + (e % n)] = G[e] + (((e / n) == (e % n)) ? lam : 0.0);
  for (int i = t; i < n; i += blockDim.x) x[i] = b[i];
  if (t == 0) ok = 1;
  __syncthreads();
  for (int k = 0; k < n; ++k) {
    const double piv = A[k * ld + k];
    if (!(piv > 0.0)) { if (t == 0) ok = 0; break; }     // uniform: every thread reads the same pivot
    const double r = sqrt(piv);
    __syncthreads();
    if (t == k) A[k * ld + k] = r;
    else if (t > k && t < n) A[t * ld + k] /= r;
    __syncthreads();
    if (t > k && t < n) {
      const double l = A[t * ld + k];
      for (int j = k + 1; j <= t; ++j) A[t * ld + j] = fma(-l, A[j * ld + k], A[t * ld + j]);
    }
    __syncthreads();
  }
  __syncthreads();
  if (t == 0 && ok) {                                      // two triangular solves, serial: 117^2 flops
    for (int i = 0; i < n; ++i) {
      double s = x[i];
      for (int j = 0; j < i; ++j) s -= A[i * ld + j] * x[j];
      x[i] = s / A[i * ld + i];
    }
    for (int i = n - 1; i >= 0; --i) {
      double s = x[i];
      for (int j = i + 1; j < n; ++j) s -= A[j * ld + i] * x[j];
      x[i] = s / A[i * ld + i];
    }
    bool fin = true;
    for (int i = 0; i < n; ++i) fin = fin && (x[i] == x[i]) && fabs(x[i]) < 1e300;
    if (fin) for (int i = 0; i < n; ++i) { mdl->beta[i] = x[i]; mdl->ref[i] = mdl->ref_new[i]; }   // (beta, ref) change together or not at all
  }
}
