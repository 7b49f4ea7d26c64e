// global Richardson extrapolation prototype: pass A adaptive (logs accepted h), pass B = merged pairs of A's steps
#include <math.h>
#include <string.h>
#include <stdint.h>
#include <vector>
#include "gync_llh.cuh"
static const double REFALLPARMS[GYNC_NP] = {GYNC_REFALLPARMS_LIST};
static const unsigned char SAMPLEDINDS[GYNC_NS] = {GYNC_SAMPLEDINDS0_LIST};
static const int MEASINDS[4] = {GYNC_MEASINDS0_LIST};
extern "C" int rx_traj(const double* x, double rtol, double atol, int merge, double* trajA, double* trajB, int* nA, int* nB) {
  GyncWork W; double T;
  gync_load_sample(W, [&](int k) { return x[k]; }, REFALLPARMS, SAMPLEDINDS, &T);
  double y0[GYNC_NY]; for (int i = 0; i < GYNC_NY; ++i) y0[i] = W.y[i];
  GyncCtrl c; gync_ctrl_init(c, gync_initial_step(W, rtol, atol));
  std::vector<std::vector<double>> hs(62);
  int i0 = 0, i1 = 0; const double big = 1e300; double t = fmin(1.0, 1.0 + T);
  int seg = 0; std::vector<int> segA, segB; std::vector<double> segt;
  while (i0 < 31 || i1 < 31) {
    const double tA = (i0 < 31) ? (double)(i0 + 1) : big, tB = (i1 < 31) ? (double)(i1 + 1) + T : big;
    const double tn = fmin(tA, tB);
    std::vector<double> hl;
    while (tn > t) {
      double h = c.h; const double rem = tn - t; const bool last = (h * 1.0001 >= rem); if (last) h = rem;
      const double hprop = c.h;
      const double err = gync_rodas4_step(W, h, rtol, atol);
      const bool fin = gync_all_finite(W.yn, GYNC_NY);
      if (gync_ctrl_update(c, h, err, fin)) {
        hl.push_back(h);
        t = last ? tn : t + h;
        for (int i = 0; i < GYNC_NY; ++i) W.y[i] = W.yn[i];
        if (last && h < hprop) c.h = fmax(c.h, hprop);
      }
      if (c.nacc + c.nrej > 400000) return 1;
    }
    hs[seg] = hl;
    if (tA == tn) { for (int h = 0; h < 4; ++h) trajA[(i0) * 4 + h] = W.y[MEASINDS[h]]; segA.push_back(i0); ++i0; } else segA.push_back(-1);
    if (tB == tn) { for (int h = 0; h < 4; ++h) trajA[(31 + i1) * 4 + h] = W.y[MEASINDS[h]]; segB.push_back(i1); ++i1; } else segB.push_back(-1);
    ++seg;
  }
  *nA = c.nacc + c.nrej;
  // pass B
  for (int i = 0; i < GYNC_NY; ++i) W.y[i] = y0[i];
  int nb = 0;
  for (int s = 0; s < seg; ++s) {
    const std::vector<double>& hl = hs[s];
    size_t j = 0;
    while (j < hl.size()) {
      double h = 0; int m = 0;
      while (m < merge && j < hl.size()) { h += hl[j++]; ++m; }
      gync_rodas4_step(W, h, rtol, atol);
      for (int i = 0; i < GYNC_NY; ++i) W.y[i] = W.yn[i];
      ++nb;
    }
    if (segA[s] >= 0) for (int h = 0; h < 4; ++h) trajB[(segA[s]) * 4 + h] = W.y[MEASINDS[h]];
    if (segB[s] >= 0) for (int h = 0; h < 4; ++h) trajB[(31 + segB[s]) * 4 + h] = W.y[MEASINDS[h]];
  }
  *nB = nb;
  return 0;
}
