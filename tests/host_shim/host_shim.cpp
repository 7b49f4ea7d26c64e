// TEST-ONLY host build of the device solver core (gync.jl_b200/csrc/*.cuh compiled as plain C++),
// so the generated RHS / Jacobian / sparse LU and the RODAS4 driver can be checked against the
// oracle on a machine without a GPU.  Never loaded by the product path.
#include <math.h>
#include <string.h>
#include <stdint.h>
#include "gync_llh.cuh"
#include "gync_philox.h"

static const double REFALLPARMS[GYNC_NP] = {GYNC_REFALLPARMS_LIST};
static const unsigned char SAMPLEDINDS[GYNC_NS] = {GYNC_SAMPLEDINDS0_LIST};
static const int MEASINDS[4] = {GYNC_MEASINDS0_LIST};
static const double MEASMAG[4] = {GYNC_MEASMAG_LIST};

extern "C" {

int hs_dims(int* nq, int* nlu) { *nq = GYNC_NQ; *nlu = GYNC_NLU; return GYNC_NY; }

void hs_rhs(const double* y, const double* p, double* f) {
  double q[GYNC_NQ];
  gync_pack_params(p, q);
  gync_rhs(y, q, f);
}

// dense W = fac*I - J (row-major 33x33) from the sparse builder, and f
void hs_rhs_w(const double* y, const double* p, double fac, double* f, double* Wd) {
  double q[GYNC_NQ], A[GYNC_NLU];
  gync_pack_params(p, q);
  gync_rhs_w(y, q, fac, f, A);
  memset(Wd, 0, sizeof(double) * GYNC_NY * GYNC_NY);
  for (int k = 0; k < GYNC_NLU; ++k) Wd[GYNC_LU_ROW[k] * GYNC_NY + GYNC_LU_COL[k]] = A[k];
}

// b <- W^{-1} b through the generated sparse LU
void hs_lusolve(const double* y, const double* p, double fac, double* b) {
  double q[GYNC_NQ], A[GYNC_NLU], f[GYNC_NY];
  gync_pack_params(p, q);
  gync_rhs_w(y, q, fac, f, A);
  gync_lu(A);
  gync_lusolve(A, b);
}

struct MeasSink {
  double* traj;   // 62 x 4 row-major: [period*31+day][h]
  void operator()(int period, int day, const double* y) {
    for (int h = 0; h < 4; ++h) traj[(period * 31 + day) * 4 + h] = y[MEASINDS[h]];
  }
};

// one sample x[116]: measured trajectory (62x4) at the measurement grid; returns status
int hs_meas_traj(const double* x, double rtol, double atol, int max_steps, double* traj, int* nacc, int* nrej) {
  GyncWork W;
  double T;
  gync_load_sample(W, [&](int k) { return x[k]; }, REFALLPARMS, SAMPLEDINDS, &T);
  GyncSolverOpts o = {rtol, atol, 0.0, 0.0, max_steps};
  MeasSink s = {traj};
  GyncStepStats st = {0, 0};
  int status = gync_run_meas_grid(W, T, o, s, &st);
  *nacc = st.nacc; *nrej = st.nrej;
  if (status) for (int i = 0; i < 248; ++i) traj[i] = NAN;
  return status;
}

// the same with a step-size cap (order tests: a loose tolerance and hmax = h give fixed steps)
int hs_meas_traj_hmax(const double* x, double rtol, double atol, double hmax, int max_steps, double* traj, int* nacc, int* nrej) {
  GyncWork W;
  double T;
  gync_load_sample(W, [&](int k) { return x[k]; }, REFALLPARMS, SAMPLEDINDS, &T);
  GyncSolverOpts o = {rtol, atol, hmax, hmax, max_steps};
  MeasSink s = {traj};
  GyncStepStats st = {0, 0};
  int status = gync_run_meas_grid(W, T, o, s, &st);
  *nacc = st.nacc; *nrej = st.nrej;
  if (status) for (int i = 0; i < 248; ++i) traj[i] = NAN;
  return status;
}

struct GridSink {
  double* Y;
  void operator()(int k, const double* y) { for (int i = 0; i < GYNC_NY; ++i) Y[k * GYNC_NY + i] = y[i]; }
};

void hs_philox(uint64_t seed, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t* out) { Philox::gen(seed, c0, c1, c2, c3, out); }
void hs_draw(uint64_t seed, uint64_t chain, uint64_t iter, double* z116, double* u) {
  for (int b = 0; b < 58; ++b) gync_draw_n2(seed, chain, iter, (uint32_t)b, z116[2 * b], z116[2 * b + 1]);
  double u1;
  gync_draw_u2(seed, chain, iter, 58u, *u, u1);
}

// gync(y0,p,t): full trajectory nT x 33 row-major
int hs_solve(const double* y0, const double* p, const double* t, int nT, double rtol, double atol, int max_steps, double* Y) {
  GyncWork W;
  gync_pack_params(p, W.q);
  for (int i = 0; i < GYNC_NY; ++i) W.y[i] = y0[i];
  GyncSolverOpts o = {rtol, atol, 0.0, 0.0, max_steps};
  GridSink s = {Y};
  int status = gync_run_time_grid(W, t, nT, o, s, nullptr);
  if (status) for (int i = 0; i < nT * GYNC_NY; ++i) Y[i] = NAN;
  return status;
}
}
