/* CPU ORACLE (test infrastructure, NOT product code) — plain C restatement of the GynC.jl hot path.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * build, link or call this.  Each function cites the reference file:line it follows (paths
 * relative to the reference repo root).
 *
 * PARITY UNPINNED at the solver level: the reference's integrator is SUNDIALS CVODE reached through
 * Sundials.jl (REQUIRE:9, no version pin, not vendored) and its tests hold no numeric goldens on
 * this path.  The RHS is pinned (tests/golden/rhs_golden.npz, generated from the reference's own
 * source text).  Two integrators live here:
 *   - "reference-tolerance mode" (orc_bdf_*): variable-order (1-5) variable-step BDF/NDF in
 *     backward-difference form, Newton with dense partially-pivoted LU, DIFFERENCE-QUOTIENT
 *     Jacobian, integrate past each output time and interpolate, 500 internal steps per output
 *     (the published behaviour of CVODE's CV_BDF/CV_NEWTON/CVDense/CV_NORMAL defaults that
 *     Sundials.cvode uses, gyncycle.jl:15) — the stand-in for the reference's CPU path;
 *   - "converged mode" (orc_seulex_*): extrapolated linearly-implicit Euler (order 8), stepping
 *     exactly onto each output time — independent of the product's RODAS4, used as the
 *     converged parity target.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define NY 33
#define NP 114
#define NX 116
#define NS 82

/* model.jl:3-4 — sampledinds = deleteat!(collect(1:103), hillinds), 0-based here */
static const int SAMPLEDINDS[NS] = {0, 1, 2, 4, 6, 7, 8, 10, 11, 12, 13, 14, 15, 16, 18, 20, 22, 23, 24, 26, 27, 28,
    29, 30, 31, 33, 34, 36, 37, 39, 40, 41, 43, 44, 45, 47, 49, 50, 52, 53, 55, 56, 57, 59, 60, 61, 62, 63, 65, 66,
    67, 68, 69, 70, 71, 72, 73, 74, 75, 76, 77, 78, 79, 80, 81, 82, 83, 84, 85, 86, 87, 88, 89, 90, 91, 92, 93, 95,
    96, 98, 99, 101};
static const int MEASUREDINDS[4] = {1, 6, 23, 24};          /* model.jl:1 */
static const double MEASMAG[4] = {120., 10., 400., 15.};    /* model.jl:10 */

static double g_refallparms[NP];   /* set by orc_set_reference (data file src/gync/data/refparms.jl) */

void orc_set_reference(const double* refallparms) { memcpy(g_refallparms, refallparms, sizeof(g_refallparms)); }
/* bench.py pins the thread count itself: launchers such as torchrun export OMP_NUM_THREADS=1 */
void orc_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}
int orc_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* ---------------------------------------------------------------- RHS (gyncycle.jl:41-46, 50-244) */
#include <complex.h>
static inline double ipow(double x, int n) { double r = 1.0; for (int i = 0; i < n; ++i) r *= x; return r; }
static inline double hplus(double s, double t, int n) { double q = ipow(s / t, n); return q / (1.0 + q); }   /* :41-44 */
static inline double hminus(double s, double t, int n) { return 1.0 / (1.0 + ipow(s / t, n)); }              /* :46 */
static inline double _Complex cipow(double _Complex x, int n) { double _Complex r = 1.0; for (int i = 0; i < n; ++i) r *= x; return r; }
static inline double _Complex chplus(double _Complex s, double t, int n) { double _Complex q = cipow(s / t, n); return q / (1.0 + q); }
static inline double _Complex chminus(double _Complex s, double t, int n) { return 1.0 / (1.0 + cipow(s / t, n)); }

#define Y(i) y[(i)-1]
#define P(i) p[(i)-1]
#define F(i) f[(i)-1]
void orc_rhs(const double* y, const double* p, double* f) {
#define REAL double
#define HPLUS hplus
#define HMINUS hminus
#define IPOW ipow
#define RPOW3689(r) pow((r), 3.689)      /* negative base -> NaN (Julia: DomainError -> NaN trajectory) */
#include "gync_rhs_body.inc"
#undef REAL
#undef HPLUS
#undef HMINUS
#undef IPOW
#undef RPOW3689
}
/* the same arithmetic on complex y: used only for the complex-step Jacobian below */
static void rhs_complex(const double _Complex* y, const double* p, double _Complex* f) {
#define REAL double _Complex
#define HPLUS chplus
#define HMINUS chminus
#define IPOW cipow
#define RPOW3689(r) cexp(3.689 * clog(r))
#include "gync_rhs_body.inc"
#undef REAL
#undef HPLUS
#undef HMINUS
#undef IPOW
#undef RPOW3689
}
#undef Y
#undef P
#undef F

/* exact (to rounding) Jacobian by complex-step differentiation of the restated RHS; J row-major */
void orc_jac_exact(const double* y, const double* p, double* J) {
  double _Complex yc[NY], fc[NY];
  for (int i = 0; i < NY; ++i) yc[i] = y[i];
  for (int j = 0; j < NY; ++j) {
    yc[j] = y[j] + 1e-30 * I;
    rhs_complex(yc, p, fc);
    for (int i = 0; i < NY; ++i) J[i * NY + j] = cimag(fc[i]) / 1e-30;
    yc[j] = y[j];
  }
}

/* allparms (model.jl:20-24) + slicing (model.jl:82-84) */
static void split_x(const double* x, double* p, double* y0, double* T) {
  memcpy(p, g_refallparms, sizeof(double) * NP);
  for (int k = 0; k < NS; ++k) p[SAMPLEDINDS[k]] = x[k];
  memcpy(y0, x + NS, sizeof(double) * NY);
  *T = x[NX - 1];
}

/* ---------------------------------------------------------------- dense linear algebra */
static int lu_factor(double* A, int* piv) {   /* row-major NY x NY, partial pivoting */
  for (int k = 0; k < NY; ++k) {
    int m = k; double mx = fabs(A[k * NY + k]);
    for (int i = k + 1; i < NY; ++i) if (fabs(A[i * NY + k]) > mx) { mx = fabs(A[i * NY + k]); m = i; }
    piv[k] = m;
    if (!(mx > 0.0) || !isfinite(mx)) return -1;
    if (m != k) for (int j = 0; j < NY; ++j) { double t = A[k * NY + j]; A[k * NY + j] = A[m * NY + j]; A[m * NY + j] = t; }
    const double inv = 1.0 / A[k * NY + k];
    for (int i = k + 1; i < NY; ++i) {
      const double l = A[i * NY + k] * inv;
      A[i * NY + k] = l;
      if (l != 0.0) for (int j = k + 1; j < NY; ++j) A[i * NY + j] -= l * A[k * NY + j];
    }
  }
  return 0;
}
static void lu_solve(const double* A, const int* piv, double* b) {
  /* all row interchanges first (the stored multipliers are in final row order), then L, then U */
  for (int k = 0; k < NY; ++k) if (piv[k] != k) { double t = b[k]; b[k] = b[piv[k]]; b[piv[k]] = t; }
  for (int k = 0; k < NY; ++k) for (int i = k + 1; i < NY; ++i) b[i] -= A[i * NY + k] * b[k];
  for (int k = NY - 1; k >= 0; --k) {
    double s = b[k];
    for (int j = k + 1; j < NY; ++j) s -= A[k * NY + j] * b[j];
    b[k] = s / A[k * NY + k];
  }
}
static int all_finite(const double* v, int n) { for (int i = 0; i < n; ++i) if (!isfinite(v[i])) return 0; return 1; }
static double rms_scaled(const double* v, const double* sc) { double s = 0; for (int i = 0; i < NY; ++i) { double r = v[i] / sc[i]; s += r * r; } return sqrt(s / NY); }

/* difference-quotient Jacobian, forward differences (CVDense default): J row-major */
static void dq_jac(const double* y, const double* fy, const double* p, double rtol, double atol, double* J) {
  double yy[NY], f1[NY];
  memcpy(yy, y, sizeof(yy));
  const double srur = sqrt(2.220446049250313e-16);
  for (int j = 0; j < NY; ++j) {
    const double ewt_inv = rtol * fabs(y[j]) + atol;        /* 1/ewt_j */
    double inc = fmax(srur * fabs(y[j]), srur * ewt_inv);
    if (inc == 0.0) inc = srur;
    yy[j] = y[j] + inc;
    orc_rhs(yy, p, f1);
    const double invinc = 1.0 / (yy[j] - y[j]);
    for (int i = 0; i < NY; ++i) J[i * NY + j] = (f1[i] - fy[i]) * invinc;
    yy[j] = y[j];
  }
}

/* ---------------------------------------------------------------- reference-tolerance mode: BDF */
#define MAXORD 5
#define NEWTON_MAXITER 4
typedef struct { long nfe, nje, nlu, nsteps, nrej; } orc_stats;

static void compute_R(int order, double factor, double R[MAXORD + 1][MAXORD + 1]) {
  double M[MAXORD + 1][MAXORD + 1];
  memset(M, 0, sizeof(M));
  for (int i = 1; i <= order; ++i) for (int j = 1; j <= order; ++j) M[i][j] = (i - 1 - factor * j) / i;
  for (int j = 0; j <= order; ++j) M[0][j] = 1.0;
  for (int j = 0; j <= order; ++j) { double c = 1.0; for (int i = 0; i <= order; ++i) { c *= M[i][j]; R[i][j] = c; } }
}
static void change_D(double D[MAXORD + 3][NY], int order, double factor) {
  double R[MAXORD + 1][MAXORD + 1], U[MAXORD + 1][MAXORD + 1], RU[MAXORD + 1][MAXORD + 1], T[MAXORD + 1][NY];
  compute_R(order, factor, R);
  compute_R(order, 1.0, U);
  for (int i = 0; i <= order; ++i) for (int j = 0; j <= order; ++j) { double s = 0; for (int k = 0; k <= order; ++k) s += R[i][k] * U[k][j]; RU[i][j] = s; }
  for (int i = 0; i <= order; ++i) for (int n = 0; n < NY; ++n) { double s = 0; for (int k = 0; k <= order; ++k) s += RU[k][i] * D[k][n]; T[i][n] = s; }
  for (int i = 0; i <= order; ++i) memcpy(D[i], T[i], sizeof(double) * NY);
}

/* Integrate y' = f(y,p) from t[0] (y(t[0]) = y0) and return y at every t[k] (row-major nT x NY),
 * stepping past each output and interpolating (CV_NORMAL).  Returns 0 on success; on any failure
 * the whole output is NaN (gyncycle.jl:18-19). */
int orc_bdf_solve(const double* y0, const double* p, const double* t, int nT, double rtol, double atol,
                  int mxstep, double* Yout, orc_stats* st) {
  static const double kappa[6] = {0, -0.1850, -1.0 / 9, -0.0823, -0.0415, 0};
  double gam[7], alpha[7], errc[7];
  gam[0] = 0; for (int k = 1; k <= 6; ++k) gam[k] = gam[k - 1] + 1.0 / k;
  for (int k = 0; k <= 5; ++k) { alpha[k] = (1 - kappa[k]) * gam[k]; errc[k] = kappa[k] * gam[k] + 1.0 / (k + 1); }
  errc[6] = 1.0 / 7;
  double D[MAXORD + 3][NY], J[NY * NY], LU[NY * NY], y[NY], f[NY], scale[NY], ypred[NY], psi[NY], d[NY], dy[NY], ynew[NY];
  int piv[NY];
  orc_stats loc = {0, 0, 0, 0, 0};
#define FAIL() do { for (int i = 0; i < nT * NY; ++i) Yout[i] = NAN; if (st) *st = loc; return 1; } while (0)
  if (!all_finite(y0, NY) || !all_finite(p, NP)) FAIL();
  memcpy(y, y0, sizeof(y));
  for (int i = 0; i < NY; ++i) Yout[i] = y0[i];
  if (nT == 1) { if (st) *st = loc; return 0; }
  double tcur = t[0];
  const double tend = t[nT - 1];
  orc_rhs(y, p, f); loc.nfe++;
  if (!all_finite(f, NY)) FAIL();
  /* initial step (Hairer's rule as in scipy select_initial_step, order 1) */
  double h;
  {
    double d0 = 0, d1 = 0, y1[NY], f1[NY];
    for (int i = 0; i < NY; ++i) { scale[i] = atol + rtol * fabs(y[i]); d0 += (y[i] / scale[i]) * (y[i] / scale[i]); d1 += (f[i] / scale[i]) * (f[i] / scale[i]); }
    d0 = sqrt(d0 / NY); d1 = sqrt(d1 / NY);
    double h0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * d0 / d1;
    for (int i = 0; i < NY; ++i) y1[i] = y[i] + h0 * f[i];
    orc_rhs(y1, p, f1); loc.nfe++;
    double d2 = 0; for (int i = 0; i < NY; ++i) { double r = (f1[i] - f[i]) / scale[i]; d2 += r * r; }
    d2 = sqrt(d2 / NY) / h0;
    double h1 = (d1 <= 1e-15 && d2 <= 1e-15) ? fmax(1e-6, h0 * 1e-3) : pow(0.01 / fmax(d1, d2), 0.5);
    h = fmin(100 * h0, h1);
    if (!(h > 0) || !isfinite(h)) h = 1e-6;
    if (tend > tcur) h = fmin(h, tend - tcur);
  }
  memset(D, 0, sizeof(D));
  memcpy(D[0], y, sizeof(y));
  for (int i = 0; i < NY; ++i) D[1][i] = f[i] * h;
  int order = 1, n_equal = 0, have_lu = 0, current_jac = 1;
  dq_jac(y, f, p, rtol, atol, J); loc.nje++; loc.nfe += NY;
  const double newton_tol = fmax(10 * 2.220446049250313e-16 / rtol, fmin(0.03, sqrt(rtol)));
  int kout = 1;
  while (kout < nT && t[kout] <= tcur) { memcpy(Yout + kout * NY, y, sizeof(y)); ++kout; }
  long steps_this_output = 0;
  while (kout < nT) {
    /* ---- one accepted step */
    if (++steps_this_output > mxstep) FAIL();            /* CVODE: mxstep internal steps per CVode call */
    double tnew = 0, safety = 0.9, err_norm = 0;
    int accepted = 0;
    while (!accepted) {
      if (!(h > 10 * 2.220446049250313e-16 * fmax(1.0, fabs(tcur)))) FAIL();
      tnew = tcur + h;
      for (int i = 0; i < NY; ++i) { double s = 0, q = 0; for (int j = 0; j <= order; ++j) s += D[j][i]; for (int j = 1; j <= order; ++j) q += gam[j] * D[j][i]; ypred[i] = s; psi[i] = q / alpha[order]; scale[i] = atol + rtol * fabs(s); }
      const double c = h / alpha[order];
      int converged = 0, n_iter = 0;
      for (;;) {
        if (!have_lu) {
          for (int i = 0; i < NY; ++i) for (int j = 0; j < NY; ++j) LU[i * NY + j] = (i == j ? 1.0 : 0.0) - c * J[i * NY + j];
          loc.nlu++;
          if (lu_factor(LU, piv)) { converged = 0; n_iter = 0; have_lu = 0; goto newton_done; }
          have_lu = 1;
        }
        {   /* simplified Newton (scipy solve_bdf_system) */
          memset(d, 0, sizeof(d)); memcpy(ynew, ypred, sizeof(ynew));
          double dy_old = -1; converged = 0;
          for (int k = 0; k < NEWTON_MAXITER; ++k) {
            n_iter = k + 1;
            orc_rhs(ynew, p, f); loc.nfe++;
            if (!all_finite(f, NY)) break;
            for (int i = 0; i < NY; ++i) dy[i] = c * f[i] - psi[i] - d[i];
            lu_solve(LU, piv, dy);
            const double dyn = rms_scaled(dy, scale);
            const double rate = (dy_old >= 0) ? dyn / dy_old : -1;
            if (rate >= 0 && (rate >= 1 || pow(rate, NEWTON_MAXITER - k) / (1 - rate) * dyn > newton_tol)) break;
            for (int i = 0; i < NY; ++i) { ynew[i] += dy[i]; d[i] += dy[i]; }
            if (dyn == 0 || (rate >= 0 && rate / (1 - rate) * dyn < newton_tol)) { converged = 1; break; }
            dy_old = dyn;
          }
        }
      newton_done:
        if (converged) break;
        if (current_jac) break;
        orc_rhs(ypred, p, f); loc.nfe++;
        if (!all_finite(f, NY)) break;
        dq_jac(ypred, f, p, rtol, atol, J); loc.nje++; loc.nfe += NY;
        have_lu = 0; current_jac = 1;
      }
      if (!converged) { h *= 0.5; change_D(D, order, 0.5); n_equal = 0; have_lu = 0; loc.nrej++; continue; }
      safety = 0.9 * (2 * NEWTON_MAXITER + 1) / (2.0 * NEWTON_MAXITER + n_iter);
      double e[NY];
      for (int i = 0; i < NY; ++i) { scale[i] = atol + rtol * fabs(ynew[i]); e[i] = errc[order] * d[i]; }
      err_norm = rms_scaled(e, scale);
      if (err_norm > 1) {
        const double factor = fmax(0.2, safety * pow(err_norm, -1.0 / (order + 1)));
        h *= factor; change_D(D, order, factor); n_equal = 0; have_lu = 0; loc.nrej++;
      } else accepted = 1;
    }
    loc.nsteps++; n_equal++;
    tcur = tnew; memcpy(y, ynew, sizeof(y));
    current_jac = 0;
    for (int i = 0; i < NY; ++i) { D[order + 2][i] = d[i] - D[order + 1][i]; D[order + 1][i] = d[i]; }
    for (int j = order; j >= 0; --j) for (int i = 0; i < NY; ++i) D[j][i] += D[j + 1][i];
    /* ---- outputs passed by this step: interpolate (dense output of the BDF polynomial) */
    while (kout < nT && t[kout] <= tcur) {
      double* out = Yout + kout * NY;
      memcpy(out, D[0], sizeof(double) * NY);
      double pr = 1.0;
      for (int j = 0; j < order; ++j) { pr *= (t[kout] - (tcur - h * j)) / (h * (1 + j)); for (int i = 0; i < NY; ++i) out[i] += D[j + 1][i] * pr; }
      ++kout; steps_this_output = 0;
    }
    if (kout >= nT) break;
    if (n_equal < order + 1) continue;
    /* ---- order / step selection */
    double em = INFINITY, ep = INFINITY, v[NY];
    if (order > 1) { for (int i = 0; i < NY; ++i) v[i] = errc[order - 1] * D[order][i]; em = rms_scaled(v, scale); }
    if (order < MAXORD) { for (int i = 0; i < NY; ++i) v[i] = errc[order + 1] * D[order + 2][i]; ep = rms_scaled(v, scale); }
    double fm = pow(em, -1.0 / order), f0 = pow(err_norm, -1.0 / (order + 1)), fp = pow(ep, -1.0 / (order + 2));
    if (em == INFINITY) fm = 0; if (ep == INFINITY) fp = 0; if (err_norm == 0) f0 = 10;
    int delta = 0; double best = f0;
    if (fm > best) { best = fm; delta = -1; }
    if (fp > best) { best = fp; delta = 1; }
    order += delta;
    const double factor = fmin(10.0, safety * best);
    h *= factor; change_D(D, order, factor); n_equal = 0; have_lu = 0;
  }
  if (st) *st = loc;
  if (!all_finite(Yout, nT * NY)) FAIL();
  return 0;
#undef FAIL
}

/* ---------------------------------------------------------------- converged mode: extrapolation */
#define KEX 8
int orc_seulex_solve(const double* y0, const double* p, const double* t, int nT, double rtol, double atol,
                     long max_steps, double* Yout, orc_stats* st) {
  double y[NY], f0[NY], J[NY * NY], W[NY * NY], T[KEX][KEX][NY], yy[NY], f[NY], sc[NY];
  int piv[NY];
  orc_stats loc = {0, 0, 0, 0, 0};
#define FAIL() do { for (int i = 0; i < nT * NY; ++i) Yout[i] = NAN; if (st) *st = loc; return 1; } while (0)
  if (!all_finite(y0, NY) || !all_finite(p, NP)) FAIL();
  memcpy(y, y0, sizeof(y));
  double tcur = t[0], H = 1e-3;
  long steps = 0;
  for (int kout = 0; kout < nT; ++kout) {
    const double tout = t[kout];
    while (tcur < tout) {
      if (++steps > max_steps) FAIL();
      double h = fmin(H, tout - tcur);
      const int last = (h >= tout - tcur);
      if (!(h > 1e-14 * fmax(1.0, fabs(tcur)))) FAIL();
      orc_rhs(y, p, f0); loc.nfe++;
      if (!all_finite(f0, NY)) FAIL();
      orc_jac_exact(y, p, J); loc.nje++;   /* exact J: the stiff error expansion of the extrapolation needs it */
      int bad = 0;
      for (int j = 0; j < KEX && !bad; ++j) {
        const int n = j + 1; const double hh = h / n;
        for (int a = 0; a < NY; ++a) for (int b = 0; b < NY; ++b) W[a * NY + b] = (a == b ? 1.0 / hh : 0.0) - J[a * NY + b];
        loc.nlu++;
        if (lu_factor(W, piv)) { bad = 1; break; }
        memcpy(yy, y, sizeof(yy));
        for (int i = 0; i < n; ++i) {
          if (i == 0) memcpy(f, f0, sizeof(f)); else { orc_rhs(yy, p, f); loc.nfe++; }
          lu_solve(W, piv, f);
          for (int a = 0; a < NY; ++a) yy[a] += f[a];
        }
        if (!all_finite(yy, NY)) { bad = 1; break; }
        memcpy(T[j][0], yy, sizeof(yy));
        for (int k = 1; k <= j; ++k) {
          const double den = (double)(j + 1) / (double)(j + 1 - k) - 1.0;
          for (int a = 0; a < NY; ++a) T[j][k][a] = T[j][k - 1][a] + (T[j][k - 1][a] - T[j - 1][k - 1][a]) / den;
        }
      }
      double err = INFINITY;
      if (!bad) {
        double e[NY];
        for (int a = 0; a < NY; ++a) { sc[a] = atol + rtol * fmax(fabs(y[a]), fabs(T[KEX - 1][KEX - 1][a])); e[a] = T[KEX - 1][KEX - 1][a] - T[KEX - 1][KEX - 2][a]; }
        err = rms_scaled(e, sc);
      }
      double fac = isfinite(err) ? fmax(0.25, fmin(5.0, pow(err, 1.0 / KEX) / 0.9)) : 5.0;
      if (isfinite(err) && err <= 1.0) {
        tcur = last ? tout : tcur + h;
        memcpy(y, T[KEX - 1][KEX - 1], sizeof(y));
        loc.nsteps++;
        const double hn = h / fac;
        H = last ? fmax(H, hn) : hn;
      } else { loc.nrej++; H = h / fac; }
    }
    memcpy(Yout + kout * NY, y, sizeof(y));
  }
  if (st) *st = loc;
  return 0;
#undef FAIL
}

/* ---------------------------------------------------------------- llh (model.jl:126-163) */
/* head of llh (model.jl:129-136): times, stable sortperm, solve, invperm, measured species.
 * mode 0: reference-tolerance BDF, mode 1: converged extrapolation.
 * traj: 62 x 4 row-major in [period0; period1] order, all-NaN on failure. Returns 1 on failure. */
int orc_meas_traj(const double* x, int mode, double rtol, double atol, double* traj, orc_stats* st) {
  double p[NP], y0[NY], T;
  split_x(x, p, y0, &T);
  double tt[62], ts[62];
  int perm[62];
  for (int q = 0; q < 2; ++q) for (int d = 0; d < 31; ++d) tt[q * 31 + d] = (double)(d + 1) + T * q;    /* meastimes :126 */
  for (int i = 0; i < 62; ++i) perm[i] = i;
  for (int i = 1; i < 62; ++i) { int k = perm[i]; int j = i - 1; while (j >= 0 && tt[perm[j]] > tt[k]) { perm[j + 1] = perm[j]; --j; } perm[j + 1] = k; }   /* stable sortperm :135 */
  for (int i = 0; i < 62; ++i) ts[i] = tt[perm[i]];
  double sol[62 * NY];
  int fail = !isfinite(T);
  if (!fail) fail = mode == 0 ? orc_bdf_solve(y0, p, ts, 62, rtol, atol, 500, sol, st) : orc_seulex_solve(y0, p, ts, 62, rtol, atol, 2000000, sol, st);
  for (int i = 0; i < 62; ++i) for (int h = 0; h < 4; ++h) traj[perm[i] * 4 + h] = fail ? NAN : sol[i * NY + MEASUREDINDS[h]];   /* invperm + measuredinds :136 */
  return fail;
}

/* tail of llh (model.jl:138-155) with llh_measerror (:158-163) */
double orc_llh_from_traj(const double* traj, const double* data /*31x4 column-major*/, double sigma_rho) {
  for (int i = 0; i < 248; ++i) if (isnan(traj[i])) return -INFINITY;   /* :138-141 */
  double l = 0.0;
  for (int q = 0; q < 2; ++q) {                      /* :148-152 */
    double s = 0; int any = 0;
    for (int h = 0; h < 4; ++h) for (int d = 0; d < 31; ++d) {      /* scaled[nonnan], column-major order */
      const double e = (data[d + 31 * h] - traj[(q * 31 + d) * 4 + h]) / (MEASMAG[h] * sigma_rho);
      if (!isnan(e)) { s += e * e; any = 1; }
    }
    if (!any) return -INFINITY;                      /* :161 */
    l += -s / 2;
  }
  return l;
}

double orc_llh(const double* x, const double* data, double sigma_rho, int mode, double rtol, double atol,
               double* traj_or_null, orc_stats* st) {
  double tr[248];
  orc_meas_traj(x, mode, rtol, atol, tr, st);
  if (traj_or_null) memcpy(traj_or_null, tr, sizeof(tr));
  return orc_llh_from_traj(tr, data, sigma_rho);
}

/* batched: X (N x 116 column-major), data 31x4xM, LL (N x M column-major).  One solve per sample
 * serves all M patients; `redundant` != 0 re-solves per (sample, patient) as the reference does
 * (weightedchain.jl:53-55) — same numbers, M x the cost.  One OpenMP thread per sample, the
 * analogue of the reference's pmap (sampling.jl:32). */
void orc_llh_batch(const double* X, int64_t N, const double* data, int M, const double* sigma_rho, int mode,
                   double rtol, double atol, int redundant, double* LL, double* traj /* N x 248, nullable */,
                   long* work /* nullable: [nfe, nje, nlu, nsteps] summed */) {
  long w0 = 0, w1 = 0, w2 = 0, w3 = 0;
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : w0, w1, w2, w3)
  for (int64_t n = 0; n < N; ++n) {
    double x[NX], tr[248];
    for (int k = 0; k < NX; ++k) x[k] = X[n + N * k];
    for (int m = 0; m < M; ++m) {
      if (m == 0 || redundant) {
        orc_stats s = {0, 0, 0, 0, 0};
        orc_meas_traj(x, mode, rtol, atol, tr, &s);
        w0 += s.nfe; w1 += s.nje; w2 += s.nlu; w3 += s.nsteps;
        if (traj && m == 0) memcpy(traj + n * 248, tr, sizeof(tr));
      }
      LL[n + N * m] = orc_llh_from_traj(tr, data + 124 * m, sigma_rho[m]);
    }
  }
  if (work) { work[0] = w0; work[1] = w1; work[2] = w2; work[3] = w3; }
}

/* generic trajectory (gync(y0,p,t), gyncycle.jl:8-26): Y row-major nT x 33 */
int orc_gync(const double* y0, const double* p, const double* t, int nT, int mode, double rtol, double atol, double* Y) {
  return mode == 0 ? orc_bdf_solve(y0, p, t, nT, rtol, atol, 500, Y, NULL) : orc_seulex_solve(y0, p, t, nT, rtol, atol, 2000000, Y, NULL);
}

/* ---------------------------------------------------------------- emiteration! (em.jl:27-41) */
void orc_emiteration(double* w, const double* L /* K x M column-major */, int64_t K, int M, double* norms /* M scratch */) {
  for (int m = 0; m < M; ++m) { double s = 0; const double* c = L + (size_t)m * K; for (int64_t k = 0; k < K; ++k) s += c[k] * w[k]; norms[m] = s; }   /* norms = L'w :31 */
  for (int64_t k = 0; k < K; ++k) {                                                                   /* :33-39 */
    double s = 0.;
    for (int m = 0; m < M; ++m) s += L[k + (size_t)K * m] / norms[m];
    w[k] = w[k] / M * s;
  }
}
/* same arithmetic with the rows split over OpenMP threads (the gemv is what OpenBLAS threads in the reference) */
void orc_emiteration_mt(double* w, const double* L, int64_t K, int M, double* norms) {
#pragma omp parallel for schedule(static)
  for (int m = 0; m < M; ++m) { double s = 0; const double* c = L + (size_t)m * K; for (int64_t k = 0; k < K; ++k) s += c[k] * w[k]; norms[m] = s; }
#pragma omp parallel for schedule(static)
  for (int64_t k = 0; k < K; ++k) {
    double s = 0.;
    for (int m = 0; m < M; ++m) s += L[k + (size_t)K * m] / norms[m];
    w[k] = w[k] / M * s;
  }
}
