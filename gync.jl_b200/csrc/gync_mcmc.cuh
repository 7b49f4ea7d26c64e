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
  // 31 quadratic forms of 33 terms each.  Species-outer / component-inner: every component keeps its own summation order
  // (bit-identical to the component-outer loop), but the 31 chains of dependent FMAs now run side by side — this function
  // is executed by a single lane while the other 31 of its warp wait (gync_mcmc_persist.cuh).
  double lp[31];
#pragma unroll
  for (int i = 0; i < 31; ++i) lp[i] = 0.0;
  for (int j = 0; j < GYNC_NY; ++j) {
    const double xj = x[GYNC_NS + j], is = istd[j];
#pragma unroll
    for (int i = 0; i < 31; ++i) {
      const double z = (xj - pr->gmm_means[i * GYNC_NY + j]) * is;
      lp[i] = fma(z, z, lp[i]);
    }
  }
  double mx = ninf;
#pragma unroll
  for (int i = 0; i < 31; ++i) {
    lp[i] = -0.5 * lp[i];
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
// K3: Metropolis with SPECULATIVE (prefetching) evaluation.
//
// A chain is sequential and one likelihood costs ~5000 Rosenbrock steps, so one thread per chain
// leaves the GPU idle (measured: 7 iterations/s for a single chain, 8.6 k chain-iterations/s for
// 4096 chains).  With a random-walk proposal most steps are rejections, and the draws of iteration
// i are addressable (Philox counter = (seed, chain, i)): so from the current state the proposals
// of iterations i, i+1, .., i+K-1 are generated at once, their K likelihoods go through the
// batched solver kernel, and a scan accepts the first success — everything up to and including it
// is EXACTLY what the sequential algorithm would have produced; later candidates (which assumed
// "still rejected") are discarded.  Expected iterations advanced per round: (1-(1-a)^K)/a.
//
// propose: one thread per (chain, slot).  mode 1 evaluates the current states (first call).
__global__ void __launch_bounds__(128)
gync_mcmc_propose_kernel(const double* __restrict__ state /* C x 116 log-space */, const long long* __restrict__ itc,
                         int64_t C, int K, int64_t iters, int mode, const double* __restrict__ chol, double prop_sd,
                         const GyncPriorDev* __restrict__ pr, uint64_t seed, uint64_t iter_offset,
                         const double* __restrict__ inj_z, const int* __restrict__ patient_of_chain,
                         double* __restrict__ XL /* N x 116 proposals, log space */, double* __restrict__ X /* exp */,
                         double* __restrict__ lpr /* N: logprior + sum(x), -Inf outside the prior */,
                         int* __restrict__ pat_of_sample,
                         const unsigned long long* __restrict__ chain_ids = nullptr /* Philox stream per chain; NULL: c */) {
  const int64_t N = C * K;
  const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= N) return;
  const int64_t c = s / K;
  const int k = (int)(s % K);
  const long long it = itc[c] + k;
  pat_of_sample[s] = patient_of_chain ? patient_of_chain[c] : 0;
  double x[GYNC_NX], z[GYNC_NX], prop[GYNC_NX];
  if (it >= iters && mode == 0) {            // beyond the requested iterations: no work for this slot
    for (int d = 0; d < GYNC_NX; ++d) { X[s + N * d] = NAN; XL[s + N * d] = NAN; }
    lpr[s] = -INFINITY;
    return;
  }
  for (int d = 0; d < GYNC_NX; ++d) x[d] = state[c + C * d];
  if (mode == 1) {
    for (int d = 0; d < GYNC_NX; ++d) prop[d] = x[d];
  } else {
    if (inj_z) {
      for (int d = 0; d < GYNC_NX; ++d) z[d] = inj_z[it + iters * (d + (int64_t)GYNC_NX * c)];
    } else {
      const uint64_t git = iter_offset + (uint64_t)it;
      const uint64_t cid = chain_ids ? (uint64_t)chain_ids[c] : (uint64_t)c;
      for (int b = 0; b < GYNC_NX / 2; ++b) gync_draw_n2(seed, cid, git, (uint32_t)b, z[2 * b], z[2 * b + 1]);
    }
    if (chol) {
      for (int d = 0; d < GYNC_NX; ++d) {
        double acc = x[d];
        for (int e = 0; e <= d; ++e) acc = fma(__ldg(chol + d + GYNC_NX * e), z[e], acc);
        prop[d] = acc;
      }
    } else {
      for (int d = 0; d < GYNC_NX; ++d) prop[d] = fma(prop_sd, z[d], x[d]);
    }
  }
  double sumx = 0.0;
  for (int d = 0; d < GYNC_NX; ++d) { sumx += prop[d]; x[d] = exp(prop[d]); }
  const double lp = gync_logprior(x, pr);            // logf(x) = logpost(c, exp(x)) + sum(x)   (model.jl:102)
  const bool inside = (lp > -INFINITY);
  for (int d = 0; d < GYNC_NX; ++d) {
    XL[s + N * d] = prop[d];
    X[s + N * d] = inside ? x[d] : NAN;              // outside the prior the likelihood is skipped (model.jl:121)
  }
  lpr[s] = inside ? lp + sumx : -INFINITY;
}

// accept scan: one thread per chain walks its K slots in iteration order.
__global__ void __launch_bounds__(128)
gync_mcmc_accept_kernel(double* __restrict__ state, double* __restrict__ logf, long long* __restrict__ itc, int64_t C, int K,
                        int64_t iters, int thin, int mode, const double* __restrict__ XL, const double* __restrict__ lpr,
                        const double* __restrict__ LL /* N */, uint64_t seed, uint64_t iter_offset,
                        const double* __restrict__ inj_u, double* __restrict__ samples_out,
                        long long* __restrict__ naccept, long long* __restrict__ min_it,
                        const unsigned long long* __restrict__ chain_ids = nullptr) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const int64_t N = C * K;
  const int64_t n_out = iters / thin;
  if (mode == 1) {
    const int64_t s = c * K;
    if (!(logf[c] == logf[c])) logf[c] = (lpr[s] > -INFINITY) ? lpr[s] + LL[s] : -INFINITY;
    return;
  }
  long long it = itc[c];
  double lf = logf[c];
  long long nacc = naccept[c];
  for (int k = 0; k < K && it < iters; ++k, ++it) {
    const int64_t s = c * K + k;
    const double lfp = (lpr[s] > -INFINITY) ? lpr[s] + LL[s] : -INFINITY;
    double u, u1;
    if (inj_u) u = inj_u[it + iters * c];
    else gync_draw_u2(seed, chain_ids ? (uint64_t)chain_ids[c] : (uint64_t)c, iter_offset + (uint64_t)it, (uint32_t)(GYNC_NX / 2), u, u1);
    const bool acc = u < exp(lfp - lf);              // Mamba: rand() < exp(logf(x') - logf(x))
    if (acc) {
      for (int d = 0; d < GYNC_NX; ++d) state[c + C * d] = XL[s + N * d];
      lf = lfp;
      ++nacc;
    }
    if ((it + 1) % thin == 0) {
      const int64_t row = (it + 1) / thin - 1;
      if (samples_out && row < n_out)
        for (int d = 0; d < GYNC_NX; ++d) samples_out[row + n_out * (d + (int64_t)GYNC_NX * c)] = exp(state[c + C * d]);   // sampling.jl:74
    }
    if (acc) { ++it; break; }                        // later slots were proposed from the old state
  }
  itc[c] = it;
  logf[c] = lf;
  naccept[c] = nacc;
  if (min_it) atomicMin(min_it, it);
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

// ---- K5 row-sharded: the same normalisation with the collectives between the phases left to the
// caller: colstats -> allreduce(max) of [colmax(M), lpmax] and allreduce(sum) of sumcounts -> expsum -> allreduce(sum)
// of colsum(M) -> rowpass (above) -> total -> allreduce(sum) of the density total -> scale.
// blocks 0..M-1: column maxima; block M: max of logprior; block M+1: sum of counts.  out: M + 2 doubles.
__global__ void __launch_bounds__(GYNC_WC_THREADS)
gync_wc_colstats_kernel(const double* __restrict__ LL, const double* __restrict__ logprior, const long long* __restrict__ counts,
                        int64_t K, int M, double* __restrict__ out) {
  __shared__ double s[32];
  const int b = blockIdx.x;
  double v = (b <= M) ? -INFINITY : 0.0;
  if (b < M) {
    const double* col = LL + (size_t)b * K;
    for (int64_t k = threadIdx.x; k < K; k += blockDim.x) v = fmax(v, col[k]);
  } else if (b == M) {
    for (int64_t k = threadIdx.x; k < K; k += blockDim.x) v = fmax(v, logprior[k]);
  } else {
    for (int64_t k = threadIdx.x; k < K; k += blockDim.x) v += (double)counts[k];
  }
  v = wc_block_reduce(v, b <= M, s);
  if (threadIdx.x == 0) out[b] = v;
}

// L = exp(LL - colmax[m]) in place, colsum[m] = sum over the local rows
__global__ void __launch_bounds__(GYNC_WC_THREADS)
gync_wc_expsum_kernel(double* __restrict__ LL, int64_t K, int M, const double* __restrict__ colmax, double* __restrict__ colsum) {
  __shared__ double s[32];
  const int b = blockIdx.x;
  double* col = LL + (size_t)b * K;
  const double mx = colmax[b];
  double sm = 0.0;
  for (int64_t k = threadIdx.x; k < K; k += blockDim.x) {
    const double e = exp(col[k] - mx);
    col[k] = e;
    sm += e;
  }
  sm = wc_block_reduce(sm, false, s);
  if (threadIdx.x == 0) colsum[b] = sm;
}

__global__ void __launch_bounds__(256) gync_wc_total_kernel(const double* __restrict__ dens_partial, int nb, double* __restrict__ total) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    double t = 0.0;
    for (int b = 0; b < nb; ++b) t += dens_partial[b];
    total[0] = t;
  }
}

// ---- duplicate collapse of WeightedChain(samplings) — weightedchain.jl:31-46: consecutive equal rows of the
// concatenated sample matrix (K x 116, column-major) become one row with a count.  flag -> scan -> scatter.
__global__ void __launch_bounds__(256)
gync_dedupe_flag_kernel(const double* __restrict__ X, int64_t K, int* __restrict__ flag) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K) return;
  int f = (k == 0);
  if (k > 0)
    for (int d = 0; d < GYNC_NX; ++d) f |= (X[k + K * d] != X[k - 1 + K * d]);     // `==` of the reference: NaN != NaN
  flag[k] = f;
}

// inclusive scan of the flags by ONE CTA (chunks of blockDim, carry across chunks): pos[k] = index of the row group
__global__ void __launch_bounds__(1024)
gync_dedupe_scan_kernel(const int* __restrict__ flag, int64_t K, int* __restrict__ pos, int* __restrict__ ngroups) {
  __shared__ int s[1024];
  __shared__ int carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int64_t base = 0; base < K; base += 1024) {
    const int64_t k = base + threadIdx.x;
    s[threadIdx.x] = (k < K) ? flag[k] : 0;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
      const int v = ((int)threadIdx.x >= o) ? s[threadIdx.x - o] : 0;
      __syncthreads();
      s[threadIdx.x] += v;
      __syncthreads();
    }
    if (k < K) pos[k] = carry + s[threadIdx.x] - 1;
    __syncthreads();
    if (threadIdx.x == 1023) carry += s[1023];
    __syncthreads();
  }
  if (threadIdx.x == 0) *ngroups = carry;
}

// rows: group g <- first row of the group (out is G x 116 with leading dimension K: the caller trims); counts[g]
__global__ void __launch_bounds__(256)
gync_dedupe_scatter_kernel(const double* __restrict__ X, int64_t K, const int* __restrict__ flag, const int* __restrict__ pos,
                           double* __restrict__ out, long long* __restrict__ counts) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K) return;
  const int g = pos[k];
  if (flag[k]) {
    for (int d = 0; d < GYNC_NX; ++d) out[g + K * d] = X[k + K * d];
    int64_t e = k + 1;
    while (e < K && !flag[e]) ++e;                 // run length (runs are short compared with K)
    counts[g] = (long long)(e - k);
  }
}
