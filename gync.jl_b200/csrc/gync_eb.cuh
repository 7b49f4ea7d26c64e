// Empirical-Bayes callers of the likelihood matrix, B200-native:
//   K6  likelihood matrix with missing values   eb/likelihoodmat.jl:34-95 (likelihoodmat_nanfast, sqeucdist_nan)
//   K7  column-major matvec pair                 eb/regularizers.jl:5-10 (logl, dlogl), :30-79 (hz, dhzdiscr), :149-160
//   K8  projected ascent step                    eb/projectsimplex.jl:11-46, eb/gradientascent.jl:1-25
// All reductions are deterministic (fixed tree / fixed partial order, no floating-point atomics).
#pragma once
#include <cooperative_groups.h>
#include <cstdint>

namespace cg = cooperative_groups;

// ------------------------------------------------------------------------------------------ K6
// L[i,j] = exp(-sq[i,j]/2) / norm[i,j]   with
//   sq[i,j]   = sum_k v_ijk ((x_ik - y_jk)/sigma_k)^2          v_ijk = both observed (not NaN)
//   norm[i,j] = sqrt(prod_k sigma_k^v * (2 pi)^sum v)           (likelihoodmat.jl:47 — sigma, not sigma^2)
// and, with renormalize, sq -= min(sq), norm /= max(norm) over the whole matrix (likelihoodmat.jl:53-54).
// The reference forms sq through A'B and two correction products (sqeucdist_nan); here it is accumulated by its
// definition, which has no cancellation, and log(norm) = g/2 with g = sum_k v_ijk (log sigma_k + log 2 pi).
// A 64 x 64 tile of pairs per CTA, the inner dimension staged through shared memory in chunks of 32; each thread owns a
// 4 x 4 micro-tile.  FP64-pipe bound: 2 FP64 instructions per (i,j,k) when the staged chunk has no missing value,
// 5 otherwise (HBM traffic is one read of A and B per tile row/column and one write of L).
#define GYNC_LM_T 64
#define GYNC_LM_KC 32
#define GYNC_LM_TP (GYNC_LM_T + 1)
#define GYNC_LM_THREADS 256

struct GyncLmSmem {
  double a[GYNC_LM_KC][GYNC_LM_TP];    // x/sigma, NaN -> 0   (rows padded: staging along k is conflict-free)
  double am[GYNC_LM_KC][GYNC_LM_TP];   // 1 observed / 0 missing
  double b[GYNC_LM_KC][GYNC_LM_TP];
  double bm[GYNC_LM_KC][GYNC_LM_TP];
  double c[GYNC_LM_KC];               // log sigma_k + log 2 pi
  double red[2][GYNC_LM_THREADS / 32];
};

// Staging of one k-chunk of a tile in two halves so that the global loads of chunk c+1 are in flight while chunk c is
// being multiplied: fetch (global -> registers, raw values, NaN kept) and put (registers -> shared memory: scaled by
// 1/sigma_k, NaN -> value 0 / mask 0).  Element e of the tile maps to (k, col) with the faster index along the smaller
// stride so that the loads coalesce; both halves use the same mapping.
#define GYNC_LM_EPT (GYNC_LM_KC * GYNC_LM_T / GYNC_LM_THREADS)      /* 8 elements per thread and tile */
#define GYNC_LM_MAXD 512                                             /* 1/sigma table in shared memory */

__device__ __forceinline__ void gync_lm_map(int e, bool kfast, int& kk, int& cc) {
  if (kfast) { kk = e % GYNC_LM_KC; cc = e / GYNC_LM_KC; }
  else       { cc = e % GYNC_LM_T;  kk = e / GYNC_LM_T; }
}

__device__ __forceinline__ void gync_lm_fetch(const double* __restrict__ X, int64_t n, int64_t s_k, int64_t s_c, int64_t c0,
                                              int k0, int D, double (&r)[GYNC_LM_EPT]) {
#pragma unroll
  for (int q = 0; q < GYNC_LM_EPT; ++q) {
    int kk, cc;
    gync_lm_map(threadIdx.x + q * GYNC_LM_THREADS, s_k == 1, kk, cc);
    const int k = k0 + kk;
    const int64_t col = c0 + cc;
    r[q] = (k < D && col < n) ? __ldg(X + k * s_k + col * s_c) : 0.0;      // padding: observed zeros (never force the slow path)
  }
}

__device__ __forceinline__ void gync_lm_put(const double (&r)[GYNC_LM_EPT], bool kfast, int k0, int D, const double* isig,
                                            const double* __restrict__ sigmas, double (*v)[GYNC_LM_TP], double (*m)[GYNC_LM_TP],
                                            int& any_missing) {
#pragma unroll
  for (int q = 0; q < GYNC_LM_EPT; ++q) {
    int kk, cc;
    gync_lm_map(threadIdx.x + q * GYNC_LM_THREADS, kfast, kk, cc);
    const int k = min(k0 + kk, D - 1);
    const double x = r[q];
    const bool obs = (x == x);
    if (!obs) any_missing = 1;
    // likelihoodmat.jl:35-36 divides by sigma; the table holds 1/sigma (one rounding apart); D > table: divide
    const double sc = (D <= GYNC_LM_MAXD) ? x * isig[k] : x / sigmas[k];
    v[kk][cc] = obs ? sc : 0.0;
    m[kk][cc] = obs ? 1.0 : 0.0;
  }
}

__global__ void __launch_bounds__(GYNC_LM_THREADS, 2)
gync_lmat_kernel(const double* __restrict__ A, int64_t I, int64_t sa_k, int64_t sa_i,
                 const double* __restrict__ B, int64_t J, int64_t sb_k, int64_t sb_j, int D,
                 const double* __restrict__ sigmas, int renormalize, double* __restrict__ L,
                 double* __restrict__ blk_min, double* __restrict__ blk_max) {
  extern __shared__ __align__(16) unsigned char gync_lm_raw[];
  GyncLmSmem& sm = *reinterpret_cast<GyncLmSmem*>(gync_lm_raw);
  __shared__ double isig[GYNC_LM_MAXD];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int64_t i0 = (int64_t)blockIdx.x * GYNC_LM_T, j0 = (int64_t)blockIdx.y * GYNC_LM_T;
  double sq[4][4], g[4][4];
#pragma unroll
  for (int u = 0; u < 4; ++u)
#pragma unroll
    for (int v = 0; v < 4; ++v) { sq[u][v] = 0.0; g[u][v] = 0.0; }
  double gfast = 0.0;
  if (D <= GYNC_LM_MAXD)
    for (int k = tid; k < D; k += GYNC_LM_THREADS) isig[k] = 1.0 / sigmas[k];
  double ra[GYNC_LM_EPT], rb[GYNC_LM_EPT];
  gync_lm_fetch(A, I, sa_k, sa_i, i0, 0, D, ra);
  gync_lm_fetch(B, J, sb_k, sb_j, j0, 0, D, rb);
  __syncthreads();

  for (int k0 = 0; k0 < D; k0 += GYNC_LM_KC) {
    int miss = 0;
    gync_lm_put(ra, sa_k == 1, k0, D, isig, sigmas, sm.a, sm.am, miss);
    gync_lm_put(rb, sb_k == 1, k0, D, isig, sigmas, sm.b, sm.bm, miss);
    if (tid < GYNC_LM_KC) sm.c[tid] = (k0 + tid < D) ? log(sigmas[k0 + tid]) + 1.8378770664093453 : 0.0;
    const int slow = __syncthreads_or(miss);
    if (k0 + GYNC_LM_KC < D) {                     // next chunk's loads fly while this one is multiplied
      gync_lm_fetch(A, I, sa_k, sa_i, i0, k0 + GYNC_LM_KC, D, ra);
      gync_lm_fetch(B, J, sb_k, sb_j, j0, k0 + GYNC_LM_KC, D, rb);
    }
    const int kn = min(GYNC_LM_KC, D - k0);
    if (!slow) {
      for (int kk = 0; kk < kn; ++kk) {
        double a[4], b[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) { a[u] = sm.a[kk][tx + 16 * u]; b[u] = sm.b[kk][ty + 16 * u]; }
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
          for (int v = 0; v < 4; ++v) {
            const double d = a[u] - b[v];
            sq[u][v] = fma(d, d, sq[u][v]);
          }
        gfast += sm.c[kk];               // every pair of the tile observes entry k: one scalar instead of 16 adds
      }
    } else {
      for (int kk = 0; kk < kn; ++kk) {
        double a[4], b[4], am[4], bm[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          a[u] = sm.a[kk][tx + 16 * u]; am[u] = sm.am[kk][tx + 16 * u];
          b[u] = sm.b[kk][ty + 16 * u]; bm[u] = sm.bm[kk][ty + 16 * u];
        }
        const double c = sm.c[kk];
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
          for (int v = 0; v < 4; ++v) {
            const double w = am[u] * bm[v];
            const double d = a[u] - b[v];
            sq[u][v] = fma(d * w, d, sq[u][v]);
            g[u][v] = fma(w, c, g[u][v]);
          }
      }
    }
    __syncthreads();
  }

  double mn = INFINITY, mx = -INFINITY;
#pragma unroll
  for (int v = 0; v < 4; ++v) {
    const int64_t j = j0 + ty + 16 * v;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int64_t i = i0 + tx + 16 * u;
      if (i < I && j < J) {
        if (renormalize) {
          const double gt = g[u][v] + gfast;
          L[i + I * j] = -0.5 * (sq[u][v] + gt);           // log of the unnormalised entry; finished by gync_lmat_exp_kernel
          mn = fmin(mn, sq[u][v]);
          mx = fmax(mx, gt);
        } else {
          L[i + I * j] = exp(-0.5 * sq[u][v]);             // normalization = ones (likelihoodmat.jl:40)
        }
      }
    }
  }
  if (renormalize) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
      mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    if ((tid & 31) == 0) { sm.red[0][tid >> 5] = mn; sm.red[1][tid >> 5] = mx; }
    __syncthreads();
    if (tid == 0) {
      for (int w = 1; w < GYNC_LM_THREADS / 32; ++w) { mn = fmin(mn, sm.red[0][w]); mx = fmax(mx, sm.red[1][w]); }
      const int64_t blk = (int64_t)blockIdx.y * gridDim.x + blockIdx.x;
      blk_min[blk] = mn;
      blk_max[blk] = mx;
    }
  }
}

// scal[0] = min sq, scal[1] = max g over all tiles
__global__ void __launch_bounds__(1024) gync_lmat_minmax_kernel(const double* __restrict__ blk_min,
                                                                const double* __restrict__ blk_max, int64_t nblk,
                                                                double* __restrict__ scal) {
  __shared__ double red[2][32];
  double mn = INFINITY, mx = -INFINITY;
  for (int64_t b = threadIdx.x; b < nblk; b += blockDim.x) { mn = fmin(mn, blk_min[b]); mx = fmax(mx, blk_max[b]); }
  for (int o = 16; o > 0; o >>= 1) {
    mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = mn; red[1][threadIdx.x >> 5] = mx; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) { mn = fmin(mn, red[0][w]); mx = fmax(mx, red[1][w]); }
    scal[0] = mn;
    scal[1] = mx;
  }
}

// L = exp(logL + (min sq + max g)/2)  ==  exp(-(sq - min sq)/2) / (norm / max norm)
__global__ void __launch_bounds__(256) gync_lmat_exp_kernel(double* __restrict__ L, int64_t n, const double* __restrict__ scal) {
  const double shift = 0.5 * (scal[0] + scal[1]);
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x)
    L[e] = exp(L[e] + shift);
}

// ------------------------------------------------------------------------------------------ K7
// Matvec pair over a column-major R x C matrix (leading dimension ld), HBM-bound: every entry is read once per product.
// Both are 2-D decompositions with per-tile partial results, summed in a fixed order by gync_mv_finish_kernel.
#define GYNC_MV_THREADS 256
#define GYNC_MV_CCHUNK 256     /* columns per CTA in the row-sum product */
#define GYNC_MV_RCHUNK 2048    /* rows per CTA in the column-sum product */
#define GYNC_MV_TCOLS 32       /* columns per CTA in the column-sum product */

// part[chunk * R + r] = sum_{c in chunk} A[r, c] * x[c]        (one thread per row, coalesced over rows)
__global__ void __launch_bounds__(GYNC_MV_THREADS)
gync_gemv_n_kernel(const double* __restrict__ A, int64_t R, int64_t C, int64_t ld, const double* __restrict__ x,
                   double* __restrict__ part) {
  __shared__ double xs[GYNC_MV_CCHUNK];
  const int64_t c0 = (int64_t)blockIdx.y * GYNC_MV_CCHUNK;
  const int cn = (int)min((int64_t)GYNC_MV_CCHUNK, C - c0);
  for (int c = threadIdx.x; c < cn; c += GYNC_MV_THREADS) xs[c] = x[c0 + c];
  __syncthreads();
  const int64_t r = (int64_t)blockIdx.x * GYNC_MV_THREADS + threadIdx.x;
  if (r >= R) return;
  const double* __restrict__ p = A + r + ld * c0;
  double acc = 0.0;
  int c = 0;
  for (; c + 8 <= cn; c += 8) {
    double v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = __ldg(p + ld * (c + u));
#pragma unroll
    for (int u = 0; u < 8; ++u) acc = fma(v[u], xs[c + u], acc);
  }
  for (; c < cn; ++c) acc = fma(__ldg(p + ld * c), xs[c], acc);
  part[(int64_t)blockIdx.y * R + r] = acc;
}

// part[rb * C + c] = sum_{r in row block rb} A[r, c] * x[r]     (one warp per column, lanes over rows)
__global__ void __launch_bounds__(GYNC_MV_THREADS)
gync_gemv_t_kernel(const double* __restrict__ A, int64_t R, int64_t C, int64_t ld, const double* __restrict__ x,
                   int rchunk /* <= GYNC_MV_RCHUNK */, double* __restrict__ part) {
  __shared__ double xs[GYNC_MV_RCHUNK];
  const int64_t r0 = (int64_t)blockIdx.x * rchunk;
  const int rn = (int)min((int64_t)rchunk, R - r0);
  for (int r = threadIdx.x; r < rn; r += GYNC_MV_THREADS) xs[r] = x[r0 + r];
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t cend = min(C, ((int64_t)blockIdx.y + 1) * GYNC_MV_TCOLS);
  for (int64_t c = (int64_t)blockIdx.y * GYNC_MV_TCOLS + warp; c < cend; c += GYNC_MV_THREADS / 32) {
    const double* __restrict__ p = A + r0 + ld * c;
    double acc = 0.0;
    int r = lane;
    for (; r + 32 * 7 < rn; r += 32 * 8) {
      double v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = __ldg(p + r + 32 * u);
#pragma unroll
      for (int u = 0; u < 8; ++u) acc = fma(v[u], xs[r + 32 * u], acc);
    }
    for (; r < rn; r += 32) acc = fma(__ldg(p + r), xs[r], acc);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) part[(int64_t)blockIdx.x * C + c] = acc;
  }
}

// out[i] = sum_p part[p * n + i] (fixed order), with the epilogue the caller needs
enum GyncMvEpi {
  GYNC_EPI_PLAIN = 0,     // out = s
  GYNC_EPI_RECIP = 1,     // out = s, out2 = 1/s                                 (norms and 1/norms: dlogl, EM)
  GYNC_EPI_DHZ_DISCR = 2, // out = -s - (1/zmult) sum_m log(rho[i + n m])          (regularizers.jl:70-75; local z only)
  GYNC_EPI_DHZ_CONT = 3,  // out = -1 - s                                          (regularizers.jl:149-160)
  GYNC_EPI_EM = 4         // out = out / M * s      (in place on w; em.jl:33-38)
};
#define GYNC_FIN_SLICES 4     /* threads per output element: slice s adds partials s, s+4, ..; slices are combined in order */
__global__ void __launch_bounds__(256)
gync_mv_finish_kernel(const double* __restrict__ part, int nparts, int64_t n, int epi, double* __restrict__ out,
                      double* __restrict__ out2, const double* __restrict__ rho, int zmult, int* __restrict__ nan_flag,
                      int64_t z_off, int64_t z_len /* rho holds rows [z_off, z_off + z_len) of the z samples (a shard) */) {
  __shared__ double sl[GYNC_FIN_SLICES][256 / GYNC_FIN_SLICES];
  const int col = threadIdx.x % (256 / GYNC_FIN_SLICES), slice = threadIdx.x / (256 / GYNC_FIN_SLICES);
  const int64_t i = (int64_t)blockIdx.x * (256 / GYNC_FIN_SLICES) + col;
  double s = 0.0;
  if (i < n)
    for (int p = slice; p < nparts; p += GYNC_FIN_SLICES) s += part[(int64_t)p * n + i];
  sl[slice][col] = s;
  __syncthreads();
  if (slice != 0 || i >= n) return;
  s = sl[0][col];
#pragma unroll
  for (int q = 1; q < GYNC_FIN_SLICES; ++q) s += sl[q][col];
  double r = s;
  if (epi == GYNC_EPI_RECIP) {
    out2[i] = 1.0 / s;
  } else if (epi == GYNC_EPI_DHZ_DISCR) {
    r = -s;
    for (int m = 0; m < zmult; ++m) {
      const int64_t zl = i + n * m - z_off;
      if (zl >= 0 && zl < z_len) r -= log(rho[zl]) / (double)zmult;
    }
    if (r != r && nan_flag) atomicOr(nan_flag, 1);          // regularizers.jl:77
  } else if (epi == GYNC_EPI_DHZ_CONT) {
    r = ((z_off == 0) ? -1.0 : 0.0) - s;                    // the constant belongs to one shard only
    if (r != r && nan_flag) atomicOr(nan_flag, 1);
  } else if (epi == GYNC_EPI_EM) {
    r = out[i] / (double)zmult * s;                         // zmult carries M here
  }
  out[i] = r;
}

// f[z] = wz[z] / rho[z]   (dhzdiscr, regularizers.jl:66)   or   wz[z] log(rho[z]) / rho[z]   (dhzloop2, :153)
// with wz = repeatweights(w, zs) = repmat(w, zmult) / zmult   (regularizers.jl:16-19)
__global__ void __launch_bounds__(256)
gync_dhz_factors_kernel(const double* __restrict__ w, int64_t K, int zmult, const double* __restrict__ rho, int64_t Z,
                        int64_t z_off, int cont, double* __restrict__ f) {
  const int64_t z = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (z >= Z) return;
  const double wz = w[(z + z_off) % K] / (double)zmult;
  const double r = rho[z];
  f[z] = cont ? wz * log(r) / r : wz / r;
}

// out[0] = scale * sum_i log(v[i]) * wt_i     wt_i = wt ? wt[i % wmod] / wdiv : 1;  entries with v == 0 skipped when
// skipzero (hz, regularizers.jl:40-43); logl (regularizers.jl:5) is the unweighted form.  One CTA, fixed order.
__global__ void __launch_bounds__(1024)
gync_logsum_kernel(const double* __restrict__ v, int64_t n, const double* __restrict__ wt, int64_t wmod, int64_t ioff,
                   double wdiv, double scale, int skipzero, double* __restrict__ out) {
  __shared__ double red[32];
  double acc = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
    const double r = v[i];
    if (skipzero && r == 0.0) continue;
    acc += wt ? log(r) * (wt[(i + ioff) % wmod] / wdiv) : log(r);
  }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[w];
    out[0] = scale * s;
  }
}

// ------------------------------------------------------------------------------------------ K8
// One projected-ascent step, fused in one cooperative launch:
//   y = w + h (ca ga + cb gb)                                  gradientascent.jl:3, likelihoodmodel.jl:56-64
//   p = projectsimplex(y)                                      projectsimplex.jl:11-46
//   w = any(p == 0) ? (1-rel) w + rel p : p                    gradientascent.jl:19-25 (movefromboundary)
// The projection threshold t solves sum_i max(y_i - t, 0) = 1.  The reference finds it by sorting; here by Michelot's
// fixed point (t <- (sum_{y_i > t} y_i - 1) / #{y_i > t}, starting from all entries), which ends at the same active
// set, needs only reductions, and terminates because the active set shrinks strictly until it is stable.
#define GYNC_PS_THREADS 1024

struct GyncPsPart { double s; long long n; };

__device__ __forceinline__ void gync_ps_reduce(double s, long long n, GyncPsPart* parts, double* sh_s, long long* sh_n,
                                               cg::grid_group& grid, double& S, long long& N) {
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    n += __shfl_xor_sync(0xffffffffu, n, o);
  }
  if ((threadIdx.x & 31) == 0) { sh_s[threadIdx.x >> 5] = s; sh_n[threadIdx.x >> 5] = n; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double bs = 0.0; long long bn = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { bs += sh_s[w]; bn += sh_n[w]; }
    parts[blockIdx.x].s = bs;
    parts[blockIdx.x].n = bn;
  }
  grid.sync();
  // every CTA adds the partials in the same order -> the same t everywhere
  if (threadIdx.x == 0) {
    double ts = 0.0; long long tn = 0;
    for (int b = 0; b < (int)gridDim.x; ++b) { ts += parts[b].s; tn += parts[b].n; }
    sh_s[0] = ts; sh_n[0] = tn;
  }
  __syncthreads();
  S = sh_s[0]; N = sh_n[0];
  __syncthreads();
}

__global__ void __launch_bounds__(GYNC_PS_THREADS, 1)
gync_ascent_step_kernel(double* __restrict__ w, const double* __restrict__ ga, double ca, const double* __restrict__ gb,
                        double cb, double h, int64_t K, double relstep, double* __restrict__ y, GyncPsPart* parts /* 2 x grid */,
                        int* __restrict__ zero_flag, double* __restrict__ hist, int* __restrict__ npass_out) {
  cg::grid_group grid = cg::this_grid();
  __shared__ double sh_s[32];
  __shared__ long long sh_n[32];
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t t0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  double s = 0.0; long long n = 0;
  for (int64_t i = t0; i < K; i += stride) {
    double v = w[i];
    if (ga || gb) {
      const double grad = (ga ? ca * ga[i] : 0.0) + (gb ? cb * gb[i] : 0.0);
      v = v + h * grad;
    }
    y[i] = v;
    s += v; n += 1;
  }
  double S; long long N;
  int pass = 0;
  gync_ps_reduce(s, n, parts + (size_t)(pass & 1) * gridDim.x, sh_s, sh_n, grid, S, N);
  double t = (S - 1.0) / (double)N;
  for (;;) {
    ++pass;
    s = 0.0; n = 0;
    for (int64_t i = t0; i < K; i += stride) {
      const double v = y[i];
      if (v > t) { s += v; n += 1; }
    }
    long long N2;
    gync_ps_reduce(s, n, parts + (size_t)(pass & 1) * gridDim.x, sh_s, sh_n, grid, S, N2);
    if (N2 == N || N2 == 0) break;
    N = N2;
    t = (S - 1.0) / (double)N;
  }
  int anyzero = 0;
  for (int64_t i = t0; i < K; i += stride) {
    const double p = fmax(y[i] - t, 0.0);
    y[i] = p;
    anyzero |= (p == 0.0);
  }
  if (__syncthreads_or(anyzero) && threadIdx.x == 0) atomicOr(zero_flag, 1);
  grid.sync();
  const int flag = *((volatile int*)zero_flag);
  for (int64_t i = t0; i < K; i += stride) {
    const double p = y[i];
    const double r = flag ? (1.0 - relstep) * w[i] + relstep * p : p;
    w[i] = r;
    if (hist) hist[i] = r;
  }
  if (npass_out && t0 == 0) *npass_out = pass;
}

// phi(x) = forwardsol(x, 0:30)[:, measuredinds] (eb/gync.jl:11, model.jl:1): pick the 4 measured species out of the
// solver's N x nT x 33 output into vec(31 x 4) columns, out[k + 124 n], k = t + 31 h.
__global__ void __launch_bounds__(256) gync_pick_measured_kernel(const double* __restrict__ Y, int64_t N, int nT,
                                                                 double* __restrict__ out) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= N * 124) return;
  const int64_t n = e % N;
  const int k = (int)(e / N);
  const int t = k % 31, h = k / 31;
  const int sp = (h == 0) ? 1 : (h == 1) ? 6 : (h == 2) ? 23 : 24;
  out[k + 124 * n] = Y[n + N * ((int64_t)t + (int64_t)nT * sp)];
}

// ------------------------------------------------------------------------------------------ K7b
// dhz in ONE pass over HBM.  The two products of dhzdiscr (rho = Lz w, then Lz' (wz ./ rho), regularizers.jl:65-72)
// cannot be fused entry by entry — a row's factor needs the whole row first — but they can be fused ROW by row when a
// z-row is contiguous: the matrix is therefore kept as Lzt (K x Z column-major: element (z,k) at k + K z; the
// likelihood matrix is symmetric in its arguments, so it is simply built with the operands swapped).  A CTA owns a set
// of z-rows; for four rows at a time it streams them once from HBM for the dot products with w (w staged in shared
// memory), reduces, forms the four factors and immediately re-reads the same 4 K doubles — now L2 hits — to
// accumulate f_z Lz[z,:] into its private K-vector in shared memory.  Per-CTA vectors are summed in a fixed order by
// gync_mv_finish_kernel.  HBM traffic: 8 Z K bytes per gradient instead of 16 Z K.
// (A software-pipelined variant that overlaps the L2 re-read of group g with the HBM stream of group g+1 measured
// the same 177-183 us at K = Z = 1e4, so the simpler form is kept.)
#define GYNC_DHZ_THREADS 512
#define GYNC_DHZ_ROWS 4

__global__ void __launch_bounds__(GYNC_DHZ_THREADS, 1)
gync_dhz_fused_kernel(const double* __restrict__ Lzt, int64_t K, int64_t Z, int64_t z_off, const double* __restrict__ w, int zmult,
                      int cont, double* __restrict__ rho_out, double* __restrict__ part /* gridDim.x x K */) {
  extern __shared__ __align__(16) double gync_dhz_sm[];
  double* acc = gync_dhz_sm;            // K
  double* ws = gync_dhz_sm + K;         // K
  __shared__ double red[GYNC_DHZ_ROWS][GYNC_DHZ_THREADS / 32];
  __shared__ double fz[GYNC_DHZ_ROWS];
  const int tid = threadIdx.x;
  for (int64_t k = tid; k < K; k += GYNC_DHZ_THREADS) { acc[k] = 0.0; ws[k] = w[k]; }
  __syncthreads();
  const int64_t K2 = K >> 1;            // K is even on this path (16-byte row alignment)
  const double2* __restrict__ ws2 = reinterpret_cast<const double2*>(ws);
  double2* acc2 = reinterpret_cast<double2*>(acc);
  for (int64_t z0 = (int64_t)blockIdx.x * GYNC_DHZ_ROWS; z0 < Z; z0 += (int64_t)gridDim.x * GYNC_DHZ_ROWS) {
    const double2* __restrict__ row[GYNC_DHZ_ROWS];
#pragma unroll
    for (int r = 0; r < GYNC_DHZ_ROWS; ++r)
      row[r] = reinterpret_cast<const double2*>(Lzt + K * min(z0 + r, Z - 1));     // rows past Z: re-read the last one, unused
    double dot[GYNC_DHZ_ROWS];
#pragma unroll
    for (int r = 0; r < GYNC_DHZ_ROWS; ++r) dot[r] = 0.0;
    int64_t k = tid;
    for (; k + GYNC_DHZ_THREADS < K2; k += 2 * GYNC_DHZ_THREADS) {
      double2 a[GYNC_DHZ_ROWS], b[GYNC_DHZ_ROWS];
#pragma unroll
      for (int r = 0; r < GYNC_DHZ_ROWS; ++r) { a[r] = row[r][k]; b[r] = row[r][k + GYNC_DHZ_THREADS]; }
      const double2 wa = ws2[k], wb = ws2[k + GYNC_DHZ_THREADS];
#pragma unroll
      for (int r = 0; r < GYNC_DHZ_ROWS; ++r) {
        dot[r] = fma(a[r].x, wa.x, dot[r]); dot[r] = fma(a[r].y, wa.y, dot[r]);
        dot[r] = fma(b[r].x, wb.x, dot[r]); dot[r] = fma(b[r].y, wb.y, dot[r]);
      }
    }
    for (; k < K2; k += GYNC_DHZ_THREADS) {
      const double2 wa = ws2[k];
#pragma unroll
      for (int r = 0; r < GYNC_DHZ_ROWS; ++r) { const double2 a = row[r][k]; dot[r] = fma(a.x, wa.x, dot[r]); dot[r] = fma(a.y, wa.y, dot[r]); }
    }
#pragma unroll
    for (int r = 0; r < GYNC_DHZ_ROWS; ++r) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) dot[r] += __shfl_xor_sync(0xffffffffu, dot[r], o);
      if ((tid & 31) == 0) red[r][tid >> 5] = dot[r];
    }
    __syncthreads();
    if (tid < GYNC_DHZ_ROWS) {
      double s = 0.0;
      for (int wp = 0; wp < GYNC_DHZ_THREADS / 32; ++wp) s += red[tid][wp];
      const int64_t z = z0 + tid;
      double f = 0.0;
      if (z < Z) {
        rho_out[z] = s;
        const double wz = ws[(z + z_off) % K] / (double)zmult;
        f = cont ? wz * log(s) / s : wz / s;
      }
      fz[tid] = f;
    }
    __syncthreads();
    double f[GYNC_DHZ_ROWS];
#pragma unroll
    for (int r = 0; r < GYNC_DHZ_ROWS; ++r) f[r] = fz[r];
    const int nrow = (int)min((int64_t)GYNC_DHZ_ROWS, Z - z0);
#pragma unroll
    for (int r = 0; r < GYNC_DHZ_ROWS; ++r) if (r >= nrow) f[r] = 0.0;
    const bool full = (nrow == GYNC_DHZ_ROWS);      // padding rows re-read the last row: 0 * Inf must not leak in
    k = tid;
    if (full) {
      for (; k + GYNC_DHZ_THREADS < K2; k += 2 * GYNC_DHZ_THREADS) {      // two column pairs per trip: 8 L2 loads in flight
        double2 a[GYNC_DHZ_ROWS], b[GYNC_DHZ_ROWS];
#pragma unroll
        for (int r = 0; r < GYNC_DHZ_ROWS; ++r) { a[r] = row[r][k]; b[r] = row[r][k + GYNC_DHZ_THREADS]; }
        double2 c = acc2[k], d = acc2[k + GYNC_DHZ_THREADS];
#pragma unroll
        for (int r = 0; r < GYNC_DHZ_ROWS; ++r) {
          c.x = fma(f[r], a[r].x, c.x); c.y = fma(f[r], a[r].y, c.y);
          d.x = fma(f[r], b[r].x, d.x); d.y = fma(f[r], b[r].y, d.y);
        }
        acc2[k] = c;
        acc2[k + GYNC_DHZ_THREADS] = d;
      }
    }
    for (; k < K2; k += GYNC_DHZ_THREADS) {
      double2 a[GYNC_DHZ_ROWS];
#pragma unroll
      for (int r = 0; r < GYNC_DHZ_ROWS; ++r) a[r] = row[r][k];
      double2 c = acc2[k];
#pragma unroll
      for (int r = 0; r < GYNC_DHZ_ROWS; ++r)
        if (r < nrow) { c.x = fma(f[r], a[r].x, c.x); c.y = fma(f[r], a[r].y, c.y); }
      acc2[k] = c;
    }
    // red / fz are rewritten only after the next __syncthreads pair; acc entries are private to their thread
  }
  __syncthreads();
  for (int64_t k = tid; k < K; k += GYNC_DHZ_THREADS) part[(int64_t)blockIdx.x * K + k] = acc[k];
}

// out (C x R column-major) = in' (in: R x C column-major); 32 x 32 tiles through shared memory
__global__ void __launch_bounds__(256) gync_transpose_kernel(const double* __restrict__ in, int64_t R, int64_t C, double* __restrict__ out) {
  __shared__ double tile[32][33];
  const int64_t r0 = (int64_t)blockIdx.x * 32, c0 = (int64_t)blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int j = ty; j < 32; j += 8)
    if (r0 + tx < R && c0 + j < C) tile[j][tx] = in[(r0 + tx) + R * (c0 + j)];
  __syncthreads();
  for (int j = ty; j < 32; j += 8)
    if (c0 + tx < C && r0 + j < R) out[(c0 + tx) + C * (r0 + j)] = tile[tx][j];
}
