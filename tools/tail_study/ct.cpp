#include <math.h>
#include <string.h>
#include <stdint.h>
#include "gync_llh.cuh"
static const double REFALLPARMS[GYNC_NP] = {GYNC_REFALLPARMS_LIST};
static const unsigned char SAMPLEDINDS[GYNC_NS] = {GYNC_SAMPLEDINDS0_LIST};
static const int MEASINDS[4] = {GYNC_MEASINDS0_LIST};
static bool ctrl(GyncCtrl& c, const double h, double err, const bool finite, double SAFE, double FAC1, double expo, int gus) {
  const double FAC2 = 1.0 / 6.0;
  if (!finite || !(err == err)) { c.h = 0.2 * h; c.nrej++; c.rej_last = 1; return false; }
  double fac = fmax(FAC2, fmin(1.0/ (1.0/FAC1) , pow(err, expo) * (1.0 / SAFE)));
  fac = fmax(1.0/FAC1, fmin(1.0/FAC2, pow(err, expo) / SAFE));   // h_new = h / fac ; fac in [1/FAC1, 6]
  double hnew = h / fac;
  if (err <= 1.0) {
    if (c.nacc > 0 && gus) {
      double facgus = (c.hacc / h) * pow(err * err / c.erracc, expo) / SAFE;
      facgus = fmax(1.0/FAC1, fmin(1.0/FAC2, facgus));
      fac = fmax(fac, facgus);
      hnew = h / fac;
    }
    c.hacc = h; c.erracc = fmax(1.0e-2, err);
    if (c.rej_last) hnew = fmin(hnew, h);
    c.rej_last = 0; c.nacc++; c.h = hnew; return true;
  }
  c.nrej++; c.rej_last = 1; c.h = hnew; return false;
}
extern "C" int ct_traj(const double* x, double rtol, double atol, double SAFE, double FAC1, double expo, int gus, double* traj, int* nsteps) {
  GyncWork W; double T;
  gync_load_sample(W, [&](int k) { return x[k]; }, REFALLPARMS, SAMPLEDINDS, &T);
  GyncCtrl c; gync_ctrl_init(c, gync_initial_step(W, rtol, atol));
  int i0 = 0, i1 = 0; const double big = 1e300; double t = fmin(1.0, 1.0 + T);
  while (i0 < 31 || i1 < 31) {
    const double tA = (i0 < 31) ? (double)(i0 + 1) : big, tB = (i1 < 31) ? (double)(i1 + 1) + T : big;
    const double tn = fmin(tA, tB);
    while (tn > t) {
      double h = c.h; const double rem = tn - t; const bool last = (h * 1.0001 >= rem); if (last) h = rem;
      const double hprop = c.h;
      const double err = gync_rodas4_step(W, h, rtol, atol);
      const bool fin = gync_all_finite(W.yn, GYNC_NY);
      if (ctrl(c, h, err, fin, SAFE, FAC1, expo, gus)) {
        t = last ? tn : t + h;
        for (int i = 0; i < GYNC_NY; ++i) W.y[i] = W.yn[i];
        if (last && h < hprop) c.h = fmax(c.h, hprop);
      }
      if (c.nacc + c.nrej > 400000) return 1;
    }
    if (tA == tn) { for (int h = 0; h < 4; ++h) traj[(i0) * 4 + h] = W.y[MEASINDS[h]]; ++i0; }
    if (tB == tn) { for (int h = 0; h < 4; ++h) traj[(31 + i1) * 4 + h] = W.y[MEASINDS[h]]; ++i1; }
  }
  *nsteps = c.nacc + c.nrej;
  return 0;
}
