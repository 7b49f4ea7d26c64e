// prototype: global error estimation by propagating the embedded local error estimates
#include <math.h>
#include <string.h>
#include <stdint.h>
#include "gync_llh.cuh"
static const double REFALLPARMS[GYNC_NP] = {GYNC_REFALLPARMS_LIST};
static const unsigned char SAMPLEDINDS[GYNC_NS] = {GYNC_SAMPLEDINDS0_LIST};
static const int MEASINDS[4] = {GYNC_MEASINDS0_LIST};
extern "C" int eps_traj(const double* x, double rtol, double atol, int variant, double* traj, double* epstraj, int* nacc) {
  GyncWork W; double T;
  gync_load_sample(W, [&](int k) { return x[k]; }, REFALLPARMS, SAMPLEDINDS, &T);
  GyncSolverOpts o = {rtol, atol, 0.0, 0.0, 200000};
  GyncCtrl c; gync_ctrl_init(c, gync_initial_step(W, rtol, atol));
  double eps[GYNC_NY] = {0};
  int i0 = 0, i1 = 0; const double big = 1e300; double t = fmin(1.0, 1.0 + T);
  while (i0 < 31 || i1 < 31) {
    const double tA = (i0 < 31) ? (double)(i0 + 1) : big, tB = (i1 < 31) ? (double)(i1 + 1) + T : big;
    const double tn = fmin(tA, tB);
    while (tn > t) {
      double h = c.h; const double rem = tn - t; const bool last = (h * 1.0001 >= rem); if (last) h = rem;
      const double hprop = c.h;
      const double err = gync_rodas4_step(W, h, rtol, atol);
      const bool fin = gync_all_finite(W.yn, GYNC_NY);
      if (gync_ctrl_update(c, h, err, fin)) {
        // accepted: W.e = k6 = local error estimate, W.A = LU of (1/(g h)) I - J(y_n)
        double b[GYNC_NY];
        if (variant == 2) {      // 4 implicit Euler substeps of g*h, le injected first
          for (int i = 0; i < GYNC_NY; ++i) b[i] = eps[i] + W.e[i];
          const double s = 1.0 / (RD_G * h);
          for (int m = 0; m < 4; ++m) { gync_lusolve(W.A, b); for (int i = 0; i < GYNC_NY; ++i) b[i] *= s; }
        } else if (variant == 3) {   // le injected after propagation
          for (int i = 0; i < GYNC_NY; ++i) b[i] = eps[i];
          const double s = 1.0 / (RD_G * h);
          for (int m = 0; m < 4; ++m) { gync_lusolve(W.A, b); for (int i = 0; i < GYNC_NY; ++i) b[i] *= s; }
          for (int i = 0; i < GYNC_NY; ++i) b[i] += W.e[i];
        } else {                 // no propagation: plain sum of the local estimates
          for (int i = 0; i < GYNC_NY; ++i) b[i] = eps[i] + W.e[i];
        }
        for (int i = 0; i < GYNC_NY; ++i) eps[i] = b[i];
        t = last ? tn : t + h;
        for (int i = 0; i < GYNC_NY; ++i) W.y[i] = W.yn[i];
        if (last && h < hprop) c.h = fmax(c.h, hprop);
      }
      if (c.nacc + c.nrej > 190000) return 1;
    }
    if (tA == tn) { for (int h = 0; h < 4; ++h) { traj[(i0) * 4 + h] = W.y[MEASINDS[h]]; epstraj[i0 * 4 + h] = eps[MEASINDS[h]]; } ++i0; }
    if (tB == tn) { for (int h = 0; h < 4; ++h) { traj[(31 + i1) * 4 + h] = W.y[MEASINDS[h]]; epstraj[(31 + i1) * 4 + h] = eps[MEASINDS[h]]; } ++i1; }
  }
  *nacc = c.nacc;
  return 0;
}
