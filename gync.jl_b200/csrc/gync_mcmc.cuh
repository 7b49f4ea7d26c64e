// K2/K3: log-prior and the many-chain Metropolis step; K5: WeightedChain normalisation.
#pragma once
#include "gync_kernels.cuh"

#include "gync_philox.h"

// ------------------------------------------------------------------------------------------
// logprior(c,x) — model.jl:111-117.  GMM: uniform mixture of 31 diagonal normals (model.jl:73-77);
// box: Truncated(Flat(),0,upper_i) -> 0 inside [0,upper_i], -Inf outside (its unknown additive
// constant cancels everywhere on the hot path); period: Normal(mu_T, sd_T).
struct GyncPriorDev {
  double gmm_means[31 * 33];
  double gmm_stds[33];
  double upper[82];
  double mu_T, sd_T;
};

GYNC_HD static double gync_logprior(const double* x, const GyncPriorDev* __restrict__ pr) {
  const double ninf = -INFINITY;
  for (int i = 0; i < GYNC_NS; ++i)
    if (!(x[i] >= 0.0 && x[i] <= pr->upper[i])) return ninf;
  double istd[GYNC_NY];
  double logdet = 0.0;
  for (int j = 0; j < GYNC_NY; ++j) {
    istd[j] = 1.0 / pr->gmm_stds[j];
    logdet += log(pr->gmm_stds[j]);
  }
  double lp[31];
  double mx = ninf;
  for (int i = 0; i < 31; ++i) {
    double s = 0.0;
    for (int j = 0; j < GYNC_NY; ++j) {
      const double z = (x[GYNC_NS + j] - pr->gmm_means[i * GYNC_NY + j]) * istd[j];
      s = fma(z, z, s);
    }
    lp[i] = -0.5 * s;
    mx = fmax(mx, lp[i]);
  }
  if (!(mx > ninf)) return (mx == mx) ? ninf : mx;   // all components underflow / NaN input
  double se = 0.0;
  for (int i = 0; i < 31; ++i) se += exp(lp[i] - mx);
  const double LOG2PI = 1.8378770664093454835606594728112;
  double l = mx + log(se) - log(31.0) - logdet - 0.5 * GYNC_NY * LOG2PI;
  const double zt = (x[GYNC_NX - 1] - pr->mu_T) / pr->sd_T;
  l += -0.5 * zt * zt - log(pr->sd_T) - 0.5 * LOG2PI;
  return l;
}

__global__ void gync_logprior_kernel(const double* __restrict__ X, int64_t N, const GyncPriorDev* __restrict__ pr,
                                     double* __restrict__ LP) {
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  double x[GYNC_NX];
  for (int k = 0; k < GYNC_NX; ++k) x[k] = X[n + N * k];
  LP[n] = gync_logprior(x, pr);
}

// ------------------------------------------------------------------------------------------
// K3: Metropolis.  One chain per thread; one forward solve per iteration.
struct GyncOnePatientSink {
  const double* __restrict__ data;     // 31 x 4 (this chain's patient)
  const double* __restrict__ iscale;   // 4
  double acc;
  __device__ __forceinline__ void operator()(int, int day, const double* y) {
    const double ym[GYNC_NMEAS] = {y[1], y[6], y[23], y[24]};
#pragma unroll
    for (int h = 0; h < GYNC_NMEAS; ++h) {
      const double d = __ldg(data + day + GYNC_NDAYS * h);
      if (d == d) {
        const double r = (d - ym[h]) * __ldg(iscale + h);
        acc = fma(r, r, acc);
      }
    }
  }
};

// logf(x) = logpost(c, exp(x)) + sum(x)   (model.jl:102, 119-124)
template <class WK>
__device__ __forceinline__ double gync_logf(WK& W, const double* xlog, const GyncPriorDev* pr, const double* data,
                                            const double* iscale, int allnan, const GyncSolverOpts& o) {
  double xs[GYNC_NX];
  double sumx = 0.0;
  for (int k = 0; k < GYNC_NX; ++k) {
    xs[k] = exp(xlog[k]);
    sumx += xlog[k];
  }
  const double lp = gync_logprior(xs, pr);
  if (lp == -INFINITY) return lp;                 // model.jl:121: llh skipped
  double T;
  gync_load_sample(W, [&](int k) { return xs[k]; }, c_refallparms, c_sampledinds, &T);
  GyncOnePatientSink sink = {data, iscale, 0.0};
  const int s = gync_run_meas_grid(W, T, o, sink, (GyncStepStats*)nullptr);
  if (s != GYNC_ST_OK || allnan) return -INFINITY;
  return lp + (-0.5 * sink.acc) + sumx;
}

// One chain per thread, grid-stride over chains; the initial log-target and every proposal go
// through the same (single) inlined copy of the solver: iteration -1 evaluates the current state.
template <int BLOCK>
__global__ void __launch_bounds__(BLOCK, 1)
gync_mcmc_kernel(double* __restrict__ state, double* __restrict__ logf, int64_t C, const double* __restrict__ chol,
                 double prop_sd, const double* __restrict__ data, const double* __restrict__ iscale,
                 const int* __restrict__ allnan, const int* __restrict__ patient_of_chain,
                 const GyncPriorDev* __restrict__ pr, GyncSolverOpts o, int64_t iters, int thin, uint64_t seed,
                 uint64_t iter_offset, const double* __restrict__ inj_z, const double* __restrict__ inj_u,
                 double* __restrict__ samples_out, long long* __restrict__ naccept) {
  extern __shared__ double gync_sm[];
  GyncWorkSm<BLOCK> W;
  W.bind(gync_sm, threadIdx.x);
  for (int64_t c = (int64_t)blockIdx.x * BLOCK + threadIdx.x; c < C; c += (int64_t)gridDim.x * BLOCK) {
    const int pat = patient_of_chain ? patient_of_chain[c] : 0;
    const double* pdata = data + (size_t)pat * GYNC_NDAYS * GYNC_NMEAS;
    const double* pisc = iscale + (size_t)pat * GYNC_NMEAS;
    const int pnan = allnan[pat];
    double x[GYNC_NX], prop[GYNC_NX], z[GYNC_NX];
    for (int k = 0; k < GYNC_NX; ++k) x[k] = state[c + C * k];
    double lf = logf[c];
    const int64_t n_out = iters / thin;
    long long nacc = 0;
    for (int64_t it = (lf == lf) ? 0 : -1; it < iters; ++it) {
      double u = 0.0;
      if (it < 0) {
        for (int k = 0; k < GYNC_NX; ++k) prop[k] = x[k];
      } else {
        const uint64_t git = iter_offset + (uint64_t)it;
        if (inj_z) {
          for (int k = 0; k < GYNC_NX; ++k) z[k] = inj_z[it + iters * (k + (int64_t)GYNC_NX * c)];
          u = inj_u[it + iters * c];
        } else {
          for (int b = 0; b < GYNC_NX / 2; ++b) gync_draw_n2(seed, (uint64_t)c, git, (uint32_t)b, z[2 * b], z[2 * b + 1]);
          double u1;
          gync_draw_u2(seed, (uint64_t)c, git, (uint32_t)(GYNC_NX / 2), u, u1);
        }
        if (chol) {
          for (int d = 0; d < GYNC_NX; ++d) {
            double s = x[d];
            for (int e = 0; e <= d; ++e) s = fma(__ldg(chol + d + GYNC_NX * e), z[e], s);
            prop[d] = s;
          }
        } else {
          for (int d = 0; d < GYNC_NX; ++d) prop[d] = fma(prop_sd, z[d], x[d]);
        }
      }
      const double lfp = gync_logf(W, prop, pr, pdata, pisc, pnan, o);
      if (it < 0) {
        lf = lfp;
        continue;
      }
      if (u < exp(lfp - lf)) {     // Mamba: rand() < exp(logf(x') - logf(x))
        for (int k = 0; k < GYNC_NX; ++k) x[k] = prop[k];
        lf = lfp;
        ++nacc;
      }
      if ((it + 1) % thin == 0) {
        const int64_t row = (it + 1) / thin - 1;
        if (samples_out && row < n_out)
          for (int k = 0; k < GYNC_NX; ++k) samples_out[row + n_out * (k + (int64_t)GYNC_NX * c)] = exp(x[k]);   // sampling.jl:74
      }
    }
    for (int k = 0; k < GYNC_NX; ++k) state[c + C * k] = x[k];
    logf[c] = lf;
    if (naccept) naccept[c] = nacc;
  }
}

// ------------------------------------------------------------------------------------------
// K5: WeightedChain normalisation (weightedchain.jl:59-67).  One block per column for the
// column max / column sum passes, row-parallel for the final scale + density.
#define GYNC_WC_THREADS 1024

__device__ __forceinline__ double wc_block_reduce(double v, bool is_max, double* s) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    const double o = __shfl_xor_sync(0xffffffffu, v, off);
    v = is_max ? fmax(v, o) : v + o;
  }
  if (lane == 0) s[warp] = v;
  __syncthreads();
  if (warp == 0) {
    v = (lane < (int)(blockDim.x >> 5)) ? s[lane] : (is_max ? -INFINITY : 0.0);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      const double o = __shfl_xor_sync(0xffffffffu, v, off);
      v = is_max ? fmax(v, o) : v + o;
    }
    if (lane == 0) s[0] = v;
  }
  __syncthreads();
  const double r = s[0];
  __syncthreads();
  return r;
}

// blocks 0..M-1: column m of LL -> L = exp(LL - max), colsum[m]; block M: max of logprior;
// block M+1: sum of counts.
__global__ void __launch_bounds__(GYNC_WC_THREADS)
gync_wc_colpass_kernel(const double* __restrict__ LL, const double* __restrict__ logprior,
                       const long long* __restrict__ counts, int64_t K, int M, double* __restrict__ L,
                       double* __restrict__ colsum, double* __restrict__ scal /* [0]=lpmax [1]=sumcounts */) {
  __shared__ double s[32];
  const int b = blockIdx.x;
  if (b < M) {
    const double* col = LL + (size_t)b * K;
    double* out = L + (size_t)b * K;
    double mx = -INFINITY;
    for (int64_t k = threadIdx.x; k < K; k += blockDim.x) {
      mx = fmax(mx, col[k]);
    }
    mx = wc_block_reduce(mx, true, s);
    double sm = 0.0;
    for (int64_t k = threadIdx.x; k < K; k += blockDim.x) {
      const double e = exp(col[k] - mx);
      out[k] = e;
      sm += e;
    }
    sm = wc_block_reduce(sm, false, s);
    if (threadIdx.x == 0) colsum[b] = sm;
  } else if (b == M) {
    double mx = -INFINITY;
    for (int64_t k = threadIdx.x; k < K; k += blockDim.x) mx = fmax(mx, logprior[k]);
    mx = wc_block_reduce(mx, true, s);
    if (threadIdx.x == 0) scal[0] = mx;
  } else {
    double sm = 0.0;
    for (int64_t k = threadIdx.x; k < K; k += blockDim.x) sm += (double)counts[k];
    sm = wc_block_reduce(sm, false, s);
    if (threadIdx.x == 0) scal[1] = sm;
  }
}

__global__ void __launch_bounds__(256)
gync_wc_rowpass_kernel(double* __restrict__ L, const double* __restrict__ logprior, const long long* __restrict__ counts,
                       int64_t K, int M, const double* __restrict__ colsum, const double* __restrict__ scal,
                       double* __restrict__ weights, double* __restrict__ density, double* __restrict__ dens_partial) {
  __shared__ double s[32];
  double part = 0.0;
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < K; k += (int64_t)gridDim.x * blockDim.x) {
    double rs = 0.0;
    for (int m = 0; m < M; ++m) {
      const double v = L[k + K * m] / colsum[m];
      L[k + K * m] = v;
      rs += v;
    }
    const double d = (rs / (double)M) * exp(logprior[k] - scal[0]);
    density[k] = d;
    weights[k] = (double)counts[k] / scal[1];
    part += d;
  }
  part = wc_block_reduce(part, false, s);
  if (threadIdx.x == 0) dens_partial[blockIdx.x] = part;
}

__global__ void __launch_bounds__(256)
gync_wc_scale_kernel(double* __restrict__ density, int64_t K, const double* __restrict__ dens_partial, int nb) {
  __shared__ double tot;
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int b = 0; b < nb; ++b) t += dens_partial[b];
    tot = t;
  }
  __syncthreads();
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < K; k += (int64_t)gridDim.x * blockDim.x)
    density[k] /= tot;
}
