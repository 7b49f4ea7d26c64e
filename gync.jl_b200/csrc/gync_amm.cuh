// K3a: Mamba's ADAPTIVE mixture Metropolis (AMMTune with adapt = true; call sites model.jl:104-106, sampling.jl:72;
// behaviour restated in SURVEY.md Appendix D) inside the speculative-evaluation scheme of gync_mcmc.cuh.
//
//   m += 1
//   x' = x + (m > 2d && u_mix >= beta ? SigmaLm : SigmaL) z          mixture of the adapted and the initial proposal
//   accept iff u < exp(logf(x') - logf(x))
//   p = m/(m+1);  Mv = p Mv + (1-p) v;  Mvv = p Mvv + (1-p) v v'      v = the state after the accept/reject
//   Sigma = (scale^2 / d / p) (Mvv - Mv Mv');  if Sigma factorises (full rank): SigmaLm = its factor (initially SigmaL)
//
// Speculation stays exact: under "everything so far was rejected" the tuning state of iteration i+j is determined (j
// updates with the unchanged state), so one CTA per chain walks the K slots sequentially on a scratch copy of
// (Mv, Mvv), factorises each covariance in shared memory and emits the K proposals; after the accept scan a second
// kernel replays the consumed iterations on the real tuning state.  Mamba factorises with a pivoted Cholesky and uses
// P L; here the factor is the unpivoted L of the same Sigma (same proposal distribution; draws cannot equal Mamba's
// MersenneTwister stream anyway) and "full rank" is the LAPACK dpstrf criterion pivot > d eps max(diag).
#pragma once
#include "gync_mcmc.cuh"

#define GYNC_AMM_THREADS 256
#define GYNC_AMM_LD (GYNC_NX + 1)      /* padded row length of the factor in shared memory */

struct GyncAmmDev {
  long long* m;        // C
  double* Mv;          // C x 116       (chain-major: one CTA owns one chain)
  double* Mvv;         // C x 116 x 116
  double* SigmaLm;     // C x 116 x 116 lower factors, row-major [i][j]; SigmaL until the first full-rank estimate
  double beta, scale;
};

struct GyncAmmSmem {
  double buf[2][GYNC_NX * GYNC_AMM_LD];
  double x[GYNC_NX], z[GYNC_NX], prop[GYNC_NX], ex[GYNC_NX], mv[GYNC_NX];
  double red[GYNC_AMM_THREADS / 32];
  int ok;
};

// Mv <- p Mv + (1-p) v ; Mvv <- p Mvv + (1-p) v v'   (global, this chain's block); v in shared memory
__device__ __forceinline__ void gync_amm_update(double* __restrict__ Mv, double* __restrict__ Mvv, const double* v, double p) {
  const double q = 1.0 - p;
  for (int e = threadIdx.x; e < GYNC_NX * GYNC_NX; e += GYNC_AMM_THREADS) {
    const int i = e / GYNC_NX, j = e % GYNC_NX;
    Mvv[e] = p * Mvv[e] + q * (v[i] * v[j]);
  }
  for (int i = threadIdx.x; i < GYNC_NX; i += GYNC_AMM_THREADS) Mv[i] = p * Mv[i] + q * v[i];
  __syncthreads();
}

// S (shared, lower triangle) = (scale^2/d/p) (Mvv - Mv Mv'), then in-place Cholesky.  Returns 1 on full rank.
__device__ __forceinline__ int gync_amm_factor(const double* __restrict__ Mv, const double* __restrict__ Mvv, double p,
                                               double scale, double* S, GyncAmmSmem& sm) {
  const double f = scale * scale / (double)GYNC_NX / p;
  for (int i = threadIdx.x; i < GYNC_NX; i += GYNC_AMM_THREADS) sm.mv[i] = Mv[i];
  __syncthreads();
  double dmax = 0.0;
  for (int e = threadIdx.x; e < GYNC_NX * GYNC_NX; e += GYNC_AMM_THREADS) {
    const int i = e / GYNC_NX, j = e % GYNC_NX;
    const double s = f * (Mvv[e] - sm.mv[i] * sm.mv[j]);
    S[i * GYNC_AMM_LD + j] = (j <= i) ? s : 0.0;
    if (i == j) dmax = fmax(dmax, s);
  }
  for (int o = 16; o > 0; o >>= 1) dmax = fmax(dmax, __shfl_xor_sync(0xffffffffu, dmax, o));
  if ((threadIdx.x & 31) == 0) sm.red[threadIdx.x >> 5] = dmax;
  __syncthreads();
  dmax = sm.red[0];
  for (int w = 1; w < GYNC_AMM_THREADS / 32; ++w) dmax = fmax(dmax, sm.red[w]);
  const double tol = (double)GYNC_NX * 2.220446049250313e-16 * dmax;
  const int i = threadIdx.x & 127, half = threadIdx.x >> 7;     // two threads per row: even / odd columns
  for (int k = 0; k < GYNC_NX; ++k) {
    const double piv = S[k * GYNC_AMM_LD + k];
    if (!(piv > tol)) return 0;                                 // uniform: every thread reads the same pivot
    const double r = sqrt(piv);
    __syncthreads();                                            // everyone has read the pivot before it is overwritten
    if (half == 0 && i < GYNC_NX) {
      if (i == k) S[k * GYNC_AMM_LD + k] = r;
      else if (i > k) S[i * GYNC_AMM_LD + k] /= r;
    }
    __syncthreads();
    if (i > k && i < GYNC_NX) {
      const double lik = S[i * GYNC_AMM_LD + k];
      for (int j = k + 1 + ((k + 1 + half) & 1); j <= i; j += 2)      // columns j > k with j = half (mod 2)
        S[i * GYNC_AMM_LD + j] = fma(-lik, S[j * GYNC_AMM_LD + k], S[i * GYNC_AMM_LD + j]);
    }
    __syncthreads();
  }
  return 1;
}

// One CTA per chain: the K speculative proposals of this round (see the header comment).
__global__ void __launch_bounds__(GYNC_AMM_THREADS, 1)
gync_amm_propose_kernel(const double* __restrict__ state, const long long* __restrict__ itc, const long long* __restrict__ nacc,
                        int64_t C, int K, int64_t iters, const double* __restrict__ chol, double prop_sd,
                        const GyncPriorDev* __restrict__ pr, uint64_t seed, uint64_t iter_offset,
                        const double* __restrict__ inj_z, const double* __restrict__ inj_umix,
                        const int* __restrict__ patient_of_chain, GyncAmmDev amm, double* __restrict__ virt /* C x (116 + 116^2) */,
                        double* __restrict__ xold /* C x 116 */, long long* __restrict__ it0, long long* __restrict__ nacc0,
                        double* __restrict__ XL, double* __restrict__ X, double* __restrict__ lpr, int* __restrict__ pat_of_sample,
                        const unsigned long long* __restrict__ chain_ids = nullptr /* Philox stream per chain; NULL: c */) {
  extern __shared__ __align__(16) unsigned char gync_amm_raw[];
  GyncAmmSmem& sm = *reinterpret_cast<GyncAmmSmem*>(gync_amm_raw);
  const int64_t c = blockIdx.x;
  const int64_t N = C * K;
  const int tid = threadIdx.x;
  double* Mv_v = virt + (size_t)c * (GYNC_NX + GYNC_NX * GYNC_NX);
  double* Mvv_v = Mv_v + GYNC_NX;
  const double* Mv = amm.Mv + (size_t)c * GYNC_NX;
  const double* Mvv = amm.Mvv + (size_t)c * GYNC_NX * GYNC_NX;
  const double* Lm = amm.SigmaLm + (size_t)c * GYNC_NX * GYNC_NX;
  for (int e = tid; e < GYNC_NX * GYNC_NX; e += GYNC_AMM_THREADS) {
    Mvv_v[e] = Mvv[e];
    sm.buf[0][(e / GYNC_NX) * GYNC_AMM_LD + (e % GYNC_NX)] = Lm[e];
  }
  for (int d = tid; d < GYNC_NX; d += GYNC_AMM_THREADS) {
    Mv_v[d] = Mv[d];
    const double xv = state[c + C * d];
    sm.x[d] = xv;
    xold[(size_t)c * GYNC_NX + d] = xv;
  }
  if (tid == 0) { it0[c] = itc[c]; nacc0[c] = nacc[c]; }
  __syncthreads();
  int good = 0;                      // which shared buffer holds the current SigmaLm
  long long m = amm.m[c];
  const long long itb = itc[c];
  for (int k = 0; k < K; ++k) {
    const int64_t s = c * K + k;
    const long long it = itb + k;
    if (tid == 0) pat_of_sample[s] = patient_of_chain ? patient_of_chain[c] : 0;
    if (it >= iters) {               // beyond the requested iterations: no work for this slot
      for (int d = tid; d < GYNC_NX; d += GYNC_AMM_THREADS) { X[s + N * d] = NAN; XL[s + N * d] = NAN; }
      if (tid == 0) lpr[s] = -INFINITY;
      continue;
    }
    m += 1;
    const uint64_t git = iter_offset + (uint64_t)it;
    double umix;
    if (inj_z) {
      for (int d = tid; d < GYNC_NX; d += GYNC_AMM_THREADS) sm.z[d] = inj_z[it + iters * (d + (int64_t)GYNC_NX * c)];
      umix = inj_umix ? inj_umix[it + iters * c] : 1.0;
    } else {
      const uint64_t cid = chain_ids ? (uint64_t)chain_ids[c] : (uint64_t)c;
      if (tid < GYNC_NX / 2) gync_draw_n2(seed, cid, git, (uint32_t)tid, sm.z[2 * tid], sm.z[2 * tid + 1]);
      double u;
      gync_draw_u2(seed, cid, git, (uint32_t)(GYNC_NX / 2), u, umix);   // u is the accept draw (accept kernel)
    }
    __syncthreads();
    const bool adapted = (m > 2 * GYNC_NX) && (umix >= amm.beta);
    if (tid < GYNC_NX) {
      const int d = tid;
      double acc = sm.x[d];
      if (adapted) {
        const double* F = sm.buf[good] + d * GYNC_AMM_LD;
        for (int e = 0; e <= d; ++e) acc = fma(F[e], sm.z[e], acc);
      } else if (chol) {
        for (int e = 0; e <= d; ++e) acc = fma(__ldg(chol + d + GYNC_NX * e), sm.z[e], acc);
      } else {
        acc = fma(prop_sd, sm.z[d], acc);
      }
      sm.prop[d] = acc;
      sm.ex[d] = exp(acc);
    }
    __syncthreads();
    if (tid == 0) {
      double sumx = 0.0;
      for (int d = 0; d < GYNC_NX; ++d) sumx += sm.prop[d];
      const double lp = gync_logprior(sm.ex, pr);
      lpr[s] = (lp > -INFINITY) ? lp + sumx : -INFINITY;
      sm.ok = (lp > -INFINITY);
    }
    __syncthreads();
    const bool inside = sm.ok != 0;
    for (int d = tid; d < GYNC_NX; d += GYNC_AMM_THREADS) {
      XL[s + N * d] = sm.prop[d];
      X[s + N * d] = inside ? sm.ex[d] : NAN;
    }
    // the update this iteration would make if its proposal is rejected
    const double p = (double)m / ((double)m + 1.0);
    gync_amm_update(Mv_v, Mvv_v, sm.x, p);
    if (m >= GYNC_NX && k + 1 < K) {             // a covariance of m+1 points has rank <= m
      if (gync_amm_factor(Mv_v, Mvv_v, p, amm.scale, sm.buf[good ^ 1], sm)) good ^= 1;
      __syncthreads();
    }
  }
}

// After the accept scan: replay the consumed iterations on the real tuning state.
__global__ void __launch_bounds__(GYNC_AMM_THREADS, 1)
gync_amm_commit_kernel(const double* __restrict__ state, const long long* __restrict__ itc, const long long* __restrict__ nacc,
                       int64_t C, GyncAmmDev amm, const double* __restrict__ xold, const long long* __restrict__ it0,
                       const long long* __restrict__ nacc0) {
  extern __shared__ __align__(16) unsigned char gync_amm_raw[];
  GyncAmmSmem& sm = *reinterpret_cast<GyncAmmSmem*>(gync_amm_raw);
  const int64_t c = blockIdx.x;
  const int tid = threadIdx.x;
  const int n = (int)(itc[c] - it0[c]);
  if (n <= 0) return;
  const bool accepted = nacc[c] != nacc0[c];
  double* Mv = amm.Mv + (size_t)c * GYNC_NX;
  double* Mvv = amm.Mvv + (size_t)c * GYNC_NX * GYNC_NX;
  double* Lm = amm.SigmaLm + (size_t)c * GYNC_NX * GYNC_NX;
  for (int d = tid; d < GYNC_NX; d += GYNC_AMM_THREADS) {
    sm.x[d] = xold[(size_t)c * GYNC_NX + d];
    sm.z[d] = state[c + C * d];                 // the state after the scan (== xold unless accepted)
  }
  __syncthreads();
  long long m = amm.m[c];
  for (int q = 0; q < n; ++q) {
    m += 1;
    const double p = (double)m / ((double)m + 1.0);
    const double* v = (accepted && q == n - 1) ? sm.z : sm.x;
    gync_amm_update(Mv, Mvv, v, p);
    if (m >= GYNC_NX) {
      if (gync_amm_factor(Mv, Mvv, p, amm.scale, sm.buf[0], sm)) {
        for (int e = tid; e < GYNC_NX * GYNC_NX; e += GYNC_AMM_THREADS)
          Lm[e] = sm.buf[0][(e / GYNC_NX) * GYNC_AMM_LD + (e % GYNC_NX)];
      }
      __syncthreads();
    }
  }
  if (tid == 0) amm.m[c] = m;
}

// setadapt!(v, true) on a fresh variate (chains with m < 0 on entry): m = 0, Mv = v, Mvv = v v', SigmaLm = SigmaL — until
// the running covariance has full rank the "adapted" component of the mixture is the initial proposal (a zero SigmaLm
// would propose x' = x, accept it with probability one and freeze the chain, which the reference's production runs
// with adapt = true contradict).
__global__ void __launch_bounds__(GYNC_AMM_THREADS)
gync_amm_init_kernel(const double* __restrict__ state, int64_t C, const double* __restrict__ chol, double prop_sd, GyncAmmDev amm) {
  const int64_t c = blockIdx.x;
  if (amm.m[c] >= 0) return;
  double* Mv = amm.Mv + (size_t)c * GYNC_NX;
  double* Mvv = amm.Mvv + (size_t)c * GYNC_NX * GYNC_NX;
  double* Lm = amm.SigmaLm + (size_t)c * GYNC_NX * GYNC_NX;
  for (int e = threadIdx.x; e < GYNC_NX * GYNC_NX; e += GYNC_AMM_THREADS) {
    const int i = e / GYNC_NX, j = e % GYNC_NX;
    Mvv[e] = state[c + C * i] * state[c + C * j];
    Lm[e] = chol ? ((j <= i) ? chol[i + GYNC_NX * j] : 0.0) : ((i == j) ? prop_sd : 0.0);
  }
  for (int d = threadIdx.x; d < GYNC_NX; d += GYNC_AMM_THREADS) Mv[d] = state[c + C * d];
  __syncthreads();
  if (threadIdx.x == 0) amm.m[c] = 0;
}
