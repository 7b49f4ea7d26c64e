#include <math.h>
#include <string.h>
#include <stdint.h>
#include "gync_llh.cuh"
static const double REFALLPARMS[GYNC_NP] = {GYNC_REFALLPARMS_LIST};
static const unsigned char SAMPLEDINDS[GYNC_NS] = {GYNC_SAMPLEDINDS0_LIST};
static const int MEASINDS[4] = {GYNC_MEASINDS0_LIST};
// norm: 0 = RMS (product), 1 = max
extern "C" int av_traj(const double* x, double rtol, const double* atolv, int norm, double* traj, int* nsteps, double* ymin, double* ymax) {
  GyncWork W; double T;
  gync_load_sample(W, [&](int k) { return x[k]; }, REFALLPARMS, SAMPLEDINDS, &T);
  GyncCtrl c; gync_ctrl_init(c, gync_initial_step(W, rtol, atolv[0]));
  for (int i = 0; i < GYNC_NY; ++i) { ymin[i] = fabs(W.y[i]); ymax[i] = fabs(W.y[i]); }
  int i0 = 0, i1 = 0; const double big = 1e300; double t = fmin(1.0, 1.0 + T);
  while (i0 < 31 || i1 < 31) {
    const double tA = (i0 < 31) ? (double)(i0 + 1) : big, tB = (i1 < 31) ? (double)(i1 + 1) + T : big;
    const double tn = fmin(tA, tB);
    while (tn > t) {
      double h = c.h; const double rem = tn - t; const bool last = (h * 1.0001 >= rem); if (last) h = rem;
      const double hprop = c.h;
      gync_rodas4_step(W, h, rtol, atolv[0]);
      double s = 0, mx = 0;
      for (int i = 0; i < GYNC_NY; ++i) { const double sc = atolv[i] + rtol * fmax(fabs(W.y[i]), fabs(W.yn[i])); const double r = W.e[i] / sc; s += r * r; mx = fmax(mx, fabs(r)); }
      const double err = norm ? mx : sqrt(s / GYNC_NY);
      const bool fin = gync_all_finite(W.yn, GYNC_NY);
      if (gync_ctrl_update(c, h, err, fin)) {
        t = last ? tn : t + h;
        for (int i = 0; i < GYNC_NY; ++i) { W.y[i] = W.yn[i]; ymin[i] = fmin(ymin[i], fabs(W.y[i])); ymax[i] = fmax(ymax[i], fabs(W.y[i])); }
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
