#include <math.h>
#include <string.h>
#include <stdint.h>
#include "gync_llh.cuh"
static const double REFALLPARMS[GYNC_NP] = {GYNC_REFALLPARMS_LIST};
static const unsigned char SAMPLEDINDS[GYNC_NS] = {GYNC_SAMPLEDINDS0_LIST};
static const int MEASINDS[4] = {GYNC_MEASINDS0_LIST};
struct Sink { double* traj; double* dtraj; GyncWork* W;
  void operator()(int period, int day, const double* y) {
    double f[GYNC_NY]; gync_rhs(y, W->q, f);
    for (int h = 0; h < 4; ++h) { traj[(period * 31 + day) * 4 + h] = y[MEASINDS[h]]; dtraj[(period * 31 + day) * 4 + h] = f[MEASINDS[h]]; } } };
extern "C" int ph_traj(const double* x, double rtol, double atol, double* traj, double* dtraj) {
  GyncWork W; double T;
  gync_load_sample(W, [&](int k) { return x[k]; }, REFALLPARMS, SAMPLEDINDS, &T);
  GyncSolverOpts o = {rtol, atol, 0.0, 0.0, 400000};
  Sink s = {traj, dtraj, &W}; GyncStepStats st = {0, 0};
  return gync_run_meas_grid(W, T, o, s, &st);
}
