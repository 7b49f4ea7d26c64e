// Measurement-grid driver and Gaussian log-likelihood epilogue, fused onto the forward solve.
//
// Follows llh(c,x,periods=2) (reference src/gync/model.jl:128-155), meastimes (:126) and
// llh_measerror (:158-163):
//   t = [(1:31) ; (1:31)+T], sorted for the solver, y(t[first]) = y0 (cvode semantics,
//   gyncycle.jl:15), measured species [2,7,24,25] (model.jl:1), residuals against the SAME
//   31x4 patient table for both periods, scaled by [120,10,400,15]*sigma_rho (model.jl:10,159),
//   NaN data = missing, all-missing table -> -Inf, failed solve -> -Inf.
// Instead of sortperm/invperm the two ascending day sequences are merged on the fly; equal times
// (integer T) are served from the same state, as CVODE does for a repeated tout.
#pragma once
#include "gync_solver.cuh"

// Sink concept:  void operator()(int period, int day /*0..30*/, const double* y /*33*/)
template <class WK, class Sink>
GYNC_HD static int gync_run_meas_grid(WK& W, const double T, const GyncSolverOpts& o,
                                             Sink& sink, GyncStepStats* st) {
  if (!gync_all_finite(W.y, GYNC_NY) || !gync_all_finite(W.q, GYNC_NQ) || !(T * 0.0 == 0.0))
    return GYNC_ST_NONFINITE;
  GyncCtrl c;
  c.h = (o.h0 > 0.0) ? o.h0 : gync_initial_step(W, o.rtol, o.atol);
  c.hacc = c.h; c.rootacc = 1.0; c.nacc = 0; c.nrej = 0; c.rej_last = 0;
  int budget = o.max_steps;
  int i0 = 0, i1 = 0;
  const double big = 1.0e300;
  double t = fmin(1.0, 1.0 + T);
  int status = GYNC_ST_OK;
  while (i0 < GYNC_NDAYS || i1 < GYNC_NDAYS) {
    const double tA = (i0 < GYNC_NDAYS) ? (double)(i0 + 1) : big;
    const double tB = (i1 < GYNC_NDAYS) ? (double)(i1 + 1) + T : big;
    const double tn = fmin(tA, tB);
    if (tn > t) {
      status = gync_advance(W, c, t, tn, o, budget);
      if (status != GYNC_ST_OK) break;
    }
    if (tA == tn) { sink(0, i0, W.y); ++i0; }
    if (tB == tn) { sink(1, i1, W.y); ++i1; }
  }
  if (st) { st->nacc = c.nacc; st->nrej = c.nrej; }
  return status;
}

// Generic sorted time grid (gync(y0,p,t): gyncycle.jl:8-26).  Sink: void(int k, const double* y).
template <class WK, class Sink>
GYNC_HD static int gync_run_time_grid(WK& W, const double* tgrid, const int nT,
                                             const GyncSolverOpts& o, Sink& sink, GyncStepStats* st) {
  if (!gync_all_finite(W.y, GYNC_NY) || !gync_all_finite(W.q, GYNC_NQ)) return GYNC_ST_NONFINITE;
  GyncCtrl c;
  c.h = (o.h0 > 0.0) ? o.h0 : gync_initial_step(W, o.rtol, o.atol);
  c.hacc = c.h; c.rootacc = 1.0; c.nacc = 0; c.nrej = 0; c.rej_last = 0;
  int budget = o.max_steps;
  double t = tgrid[0];
  int status = GYNC_ST_OK;
  for (int k = 0; k < nT; ++k) {
    const double tn = tgrid[k];
    if (!(tn == tn)) { status = GYNC_ST_NONFINITE; break; }
    if (tn > t) {
      status = gync_advance(W, c, t, tn, o, budget);
      if (status != GYNC_ST_OK) break;
    }
    sink(k, W.y);
  }
  if (st) { st->nacc = c.nacc; st->nrej = c.nrej; }
  return status;
}
