/* gync_b200 — C ABI of the B200-native GynC.jl hot path.
 *
 * The reference (axsk/GynC.jl) has no FFI layer: the seam is Julia-method level.  Each entry
 * point below replaces the body of one reference method (cited as file:line, paths relative to
 * the reference repo) and is what a `ccall` from that method binds; INTEGRATION.md shows the
 * Julia stubs.  Conventions:
 *   - all matrices are COLUMN-MAJOR, as Julia stores them; `int64_t` for Julia `Int` sizes,
 *     `int32_t` for `Cint`;
 *   - the caller owns every host buffer; nothing is retained past return;
 *   - return 0 on success, <0 on argument / CUDA errors (text via gync_last_error()); never
 *     throws.  A NUMERICAL failure of one sample is not an error: its status is non-zero, its
 *     trajectory NaN and its log-likelihood -Inf (gyncycle.jl:18-19, model.jl:138-141);
 *   - calls block until results are in the output buffers; one host thread at a time;
 *   - there is NO CPU fallback: without a CUDA device every compute call returns GYNC_ERR_CUDA.
 *   - `*_dev` variants take DEVICE pointers and a cudaStream_t (as void*) and do not
 *     synchronise; they are what a host that keeps data resident in HBM calls.
 *
 * Environment switches (read per call; for A/B measurements and debugging, never needed for correctness):
 *   GYNC_NO_SCHED=1        run the solver's sample queue in input order (default: longest-first by a cost model the
 *                          library refits on the device after every batch; results are bit-identical either way)
 *   GYNC_MCMC_SPEC=K       proposals each chain keeps in flight, fixed (default: its share of 1.1 per solver lane, by remaining
 *                          iterations); the chains are bit-identical for every value
 *   GYNC_MCMC_OVERSUB=f    that 1.1
 *   GYNC_MCMC_ROUNDS=1     non-adaptive chains through the round-1 loop of propose / solve / accept launches instead of the
 *                          persistent kernel (A/B reference; same chains)
 *   GYNC_MCMC_MAXSTEPS=n   step budget of the likelihood solves inside gync_mcmc_run* (default: the caller's opts)
 *   GYNC_DHZ_TWO_PASS=1    dhz as the reference's two matvecs instead of the one-pass row-fused kernel
 *   GYNC_EM_NO_RESIDENT=1  EM always streams L from HBM (default: shared-memory resident when this GPU's rows fit)
 *   GYNC_EM_NO_HYBRID=1    no partly resident shard: when at least half (but not all) of a block's rows fit in shared
 *                          memory the EM kernel keeps those and streams the rest; with this set it streams everything
 *   GYNC_SOLVE_SM=1        gync/forwardsol through the older shared-memory solver kernel
 *   GYNC_TRACE_CALLS=1     gync_loglik_batch prints host time stamps and device event times of the call on stderr
 */
#ifndef GYNC_B200_H
#define GYNC_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GYNC_OK          0
#define GYNC_ERR_ARG    -1
#define GYNC_ERR_CUDA   -2
#define GYNC_ERR_NOMEM  -3
#define GYNC_ERR_NUMERIC -4   /* the reference would have raised (e.g. "encountered NaN in dhz") */

/* per-sample solver status bits */
#define GYNC_STATUS_OK        0
#define GYNC_STATUS_MAXSTEPS  1
#define GYNC_STATUS_HMIN      2
#define GYNC_STATUS_NONFINITE 4

/* ---- lifecycle ------------------------------------------------------------------------ */
int  gync_init(int device);              /* select device, create the library stream */
/* Drive several GPUs of one NVLink domain from ONE process (the Julia host is a single process): the host-buffer
 * entry points gync_loglik_batch, gync_mcmc_run(_ids) and gync_em_iterate then shard samples / chains / rows over the
 * devices (the EM through the fused peer-memory kernel).  devices == NULL means 0..ndev-1. */
int  gync_init_multi(int32_t ndev, const int32_t* devices);
int  gync_ndev(void);
void gync_shutdown(void);
const char* gync_last_error(void);
int  gync_abi_version(void);
/* model dimensions: ny=33, np=114, nx=116, nq (derived params), nlu (LU nonzeros) */
void gync_dims(int32_t* ny, int32_t* np, int32_t* nx, int32_t* nq, int32_t* nlu);

/* solver options shared by the solve / likelihood / MCMC entry points.
 * rtol/atol are the `reltol`/`abstol` keywords of gync() (gyncycle.jl:8).  max_steps<=0 -> default.
 * rtol/atol <= 0 (or opts == NULL) select the library defaults below: the tolerance at which every
 * log-likelihood of the validation sets is within 1e-8 relative of the converged solution
 * (the reference's own default, 1e-4, is ~1e-3 away from it). */
#define GYNC_DEFAULT_RTOL 7e-10
#define GYNC_DEFAULT_ATOL 1e-11
typedef struct {
  double  rtol, atol;
  double  h0;          /* initial step, <=0 automatic */
  double  hmax;        /* <=0 none */
  int32_t max_steps;   /* step attempts per solve (the reference's CVODE cap is 500 per output) */
  int32_t reserved;
} gync_solver_opts;

/* ---- gync(y0,p,t;reltol,abstol)  — gyncycle.jl:8-26 --------------------------------------
 * Y0: N x 33, P: N x 114 (full parameter vectors), t: nT sorted times, t[0] = initial time.
 * Y:  N x nT x 33 (for N=1 this is exactly the reference's nT x 33 matrix). */
int gync_solve(const double* Y0, const double* P, int64_t N, const double* t, int32_t nT,
               const gync_solver_opts* opts, double* Y, int32_t* status);

/* ---- forwardsol(x,tspan) — model.jl:90 (allparms: model.jl:20-24) ------------------------
 * X: N x 116 samples [82 parms | 33 y0 | period]. */
int gync_forwardsol(const double* X, int64_t N, const double* t, int32_t nT,
                    const gync_solver_opts* opts, double* Y, int32_t* status);

/* ---- llh(c,x) / llh(s::Sampling) / llh block of WeightedChain(...) ----------------------
 * model.jl:128-163, sampling.jl:32, weightedchain.jl:53-56.
 * data: 31 x 4 x M patient tables (NaN = missing), sigma_rho: M.
 * LL: N x M log-likelihoods (one forward solve per sample serves all M patients).
 * Ymeas (nullable): N x 62 x 4 measured-species trajectory at meastimes(31,2,T) in
 * [period0; period1] order (model.jl:126,136).  nsteps (nullable): N x 2 accepted/rejected. */
int gync_loglik_batch(const double* X, int64_t N, const double* data, int32_t M,
                      const double* sigma_rho, const gync_solver_opts* opts,
                      double* LL, double* Ymeas, int32_t* status, int32_t* nsteps);
int gync_loglik_batch_dev(const double* dX, int64_t N, const double* d_data, int32_t M,
                          const double* d_sigma_rho, const gync_solver_opts* opts,
                          double* dLL, double* dYmeas, int32_t* d_status, int32_t* d_nsteps,
                          void* stream);

/* ---- logprior(c,x) — model.jl:111-117, priors model.jl:69-77 ---------------------------- */
typedef struct {
  double gmm_means[31 * 33];   /* row-major: component i, species j (rows of referencesolution()) */
  double gmm_stds[33];
  double upper[82];            /* box (0, upper_i): Truncated(Flat(),0,5*refparms_i) */
  double mu_T, sd_T;           /* Normal(28.9, 3.4) */
} gync_prior;
int gync_logprior_batch(const double* X, int64_t N, const gync_prior* prior, double* LP);

/* ---- sample!(s::Sampling, iters) for many chains — sampling.jl:62-79, model.jl:97-107 ----
 * Non-adaptive Mamba AMM step: x' = x + SigmaL z, accept iff u < exp(logf(x')-logf(x)),
 * logf(x) = logpost(c, exp(x)) + sum(x).
 * state: C x 116 log-space chain states (in/out); logf: C current log-target (in/out; pass NaN
 *   to have it computed); chol: 116 x 116 lower Cholesky factor of propvar, or NULL with
 *   `prop_sd` > 0 for the isotropic default (model.jl:14-16: sd = sqrt(log(1.01)));
 * patient_of_chain: C indices into data (0-based);
 * samples_out: n_out x 116 x C with n_out = iters/thin rows of exp(state) (sampling.jl:74);
 * rng: Philox4x32-10 keyed by (seed, chain), counter = (iteration offset) so runs are resumable;
 * inj_z (iters x 116 x C) / inj_u (iters x C): if non-NULL, use these draws instead (tests). */
int gync_mcmc_run(double* state, double* logf, int64_t C, const double* chol, double prop_sd,
                  const double* data, int32_t M, const double* sigma_rho,
                  const int32_t* patient_of_chain, const gync_prior* prior,
                  const gync_solver_opts* opts, int64_t iters, int32_t thin,
                  uint64_t seed, uint64_t iter_offset,
                  const double* inj_z, const double* inj_u,
                  double* samples_out, int64_t* naccept);

/* The same with an explicit Philox stream id per chain (chain_ids[c]; NULL = position c in this call), so that a chain
 * keeps its own random stream when it is advanced in a different group or alone (resume, batch.jl:26-46).
 * Non-adaptive chains run as ONE persistent launch per device (propose / solve / accept / re-propose on the device, no
 * host round trip between rounds); under gync_init_multi the chains are sharded over the devices in contiguous blocks. */
int gync_mcmc_run_ids(double* state, double* logf, int64_t C, const double* chol, double prop_sd,
                      const double* data, int32_t M, const double* sigma_rho,
                      const int32_t* patient_of_chain, const gync_prior* prior,
                      const gync_solver_opts* opts, int64_t iters, int32_t thin,
                      uint64_t seed, uint64_t iter_offset, const uint64_t* chain_ids,
                      const double* inj_z, const double* inj_u,
                      double* samples_out, int64_t* naccept);
/* Likelihood solves started and Rosenbrock step attempts of the last gync_mcmc_run* call (all devices), and the number of
 * undecided iterations each chain keeps in flight (its window): iterations / solves is the efficiency of the speculation. */
void gync_mcmc_last_stats(uint64_t* solves, uint64_t* steps, int32_t* slots_per_round);
/* solves dropped or given up because their chain had accepted an earlier proposal */
uint64_t gync_mcmc_last_cancelled(void);
/* solves stopped early because their residual sum had already passed what acceptance allows (the decision is exact) */
uint64_t gync_mcmc_last_early(void);
/* completed solves of the last run by floor(log2(step attempts)) (24 bins), and how many of them failed */
void gync_mcmc_last_histogram(uint64_t* hist24, uint64_t* failed);

/* ---- the same with Mamba's adaptation switched on (Config.adapt = true; model.jl:104-106, sampling.jl:72) ----------
 * AMMTune state per chain, in/out so that a Sampling can be continued (sampling.jl:62-68) and propadapt() read
 * (sampling.jl:26-29).  m[c] < 0 on entry: adaptation starts here (m = 0, Mv = v, Mvv = v v', SigmaLm = SigmaL).
 * Per iteration: m += 1; x' = x + (m > 2*116 && u_mix >= beta ? SigmaLm : SigmaL) z; accept; p = m/(m+1);
 * Mv = p Mv + (1-p) v; Mvv = p Mvv + (1-p) v v'; Sigma = scale^2/116/p (Mvv - Mv Mv'); if full rank SigmaLm = chol. */
typedef struct {
  double   beta, scale;   /* 0.05, 2.38 (model.jl:105-106) */
  int64_t* m;             /* C */
  double*  Mv;            /* 116 x C */
  double*  Mvv;           /* 116 x 116 x C */
  double*  SigmaLm;       /* 116 x 116 x C, chain c's lower factor stored ROW-major ([i][j] at i*116 + j) */
} gync_amm_tune;
int gync_mcmc_run_amm(double* state, double* logf, int64_t C, const double* chol, double prop_sd,
                      const double* data, int32_t M, const double* sigma_rho,
                      const int32_t* patient_of_chain, const gync_prior* prior,
                      const gync_solver_opts* opts, int64_t iters, int32_t thin,
                      uint64_t seed, uint64_t iter_offset,
                      const double* inj_z, const double* inj_u, const double* inj_umix /* iters x C */,
                      double* samples_out, int64_t* naccept, const gync_amm_tune* tune);
int gync_mcmc_run_amm_ids(double* state, double* logf, int64_t C, const double* chol, double prop_sd,
                          const double* data, int32_t M, const double* sigma_rho,
                          const int32_t* patient_of_chain, const gync_prior* prior,
                          const gync_solver_opts* opts, int64_t iters, int32_t thin,
                          uint64_t seed, uint64_t iter_offset, const uint64_t* chain_ids,
                          const double* inj_z, const double* inj_u, const double* inj_umix /* iters x C */,
                          double* samples_out, int64_t* naccept, const gync_amm_tune* tune);

/* ---- emiteration!(w,L) / emiterations — em.jl:27-41, em.jl:5 ----------------------------
 * L: K x M, w: K (updated in place, niter times).  history (nullable): K x niter, column i =
 * weights after iteration i+1. */
int gync_em_iterate(double* w, const double* L, int64_t K, int32_t M, int32_t niter, double* history);
int gync_em_iterate_dev(double* dw, const double* dL, int64_t K, int32_t M, int32_t niter,
                        double* d_history, void* stream);
int gync_em_iterate_multi(double* w, const double* L, int64_t K, int32_t M, int32_t niter);   /* needs gync_init_multi */
/* emiteration!(w) driven one iteration per call from a host loop (em.jl:7): keep L in HBM between the calls.  The handle
 * copies L (K x M, column-major) to the current device once; every gync_em_iterate_handle then moves only the K weights.
 * A WeightedChain owns one handle for its likelihood matrix. */
typedef struct gync_em_handle gync_em_handle;
int  gync_em_create(const double* L, int64_t K, int32_t M, gync_em_handle** out);
int  gync_em_iterate_handle(gync_em_handle* h, double* w, int32_t niter, double* history);
void gync_em_destroy(gync_em_handle* h);
/* Sharded form (one process per GPU): one EM step on the local rows given the GLOBAL
 * denominators d_norms_in (M); writes this shard's partial denominators of the NEXT iteration
 * to d_norms_out (M) for the caller's allreduce.  em_norms_dev computes the first partials. */
int gync_em_norms_dev(const double* dw, const double* dL, int64_t K, int32_t M, double* d_norms_out, void* stream);
int gync_em_step_dev(double* dw, const double* dL, int64_t K, int32_t M, const double* d_norms_in,
                     double* d_norms_out, void* stream);
/* Fused multi-GPU EM (one process per GPU, <= 8 ranks of one NVLink domain): the step and the
 * allreduce of the M denominators run as ONE persistent kernel per rank; partial denominators
 * are stored straight into the peers' mailboxes over NVLink (cudaIpc-mapped), no NCCL inside the
 * loop.  create -> exchange the 64-byte handles among ranks (any transport) -> connect.
 * Every rank must call gync_em_iterate_fused_dev with the same M and niter. */
typedef struct gync_em_comm gync_em_comm;
int  gync_em_comm_create(int32_t world, int32_t rank, void* handle_out /* 64 bytes */, gync_em_comm** comm);
int  gync_em_comm_connect(gync_em_comm* comm, const void* all_handles /* world x 64 bytes, rank order */);
void gync_em_comm_destroy(gync_em_comm* comm);
int  gync_em_iterate_fused_dev(gync_em_comm* comm, double* dw_local, const double* dL_local, int64_t K_local,
                               int32_t M, int32_t niter, void* stream);
int  gync_em_comm_error(gync_em_comm* comm);   /* 1 if a peer wait timed out */
/* launches of our kernels since init (bench bookkeeping) */
int64_t gync_launch_count(void);

/* ---- WeightedChain normalisation — weightedchain.jl:59-67 -------------------------------
 * LL: K x M log-likelihoods (in) -> L (out, may alias): exp(LL - colmax) ./ colsum;
 * logprior: K, counts: K (int64) -> weights = counts/sum, density = rowmean(L).*exp(lp-max), normalised. */
int gync_wc_normalize(const double* LL, const double* logprior, const int64_t* counts, int64_t K, int32_t M,
                      double* L, double* weights, double* density);
/* device-resident form: dLL is normalised in place (LL -> L), so K1 -> K5 -> K4 never leave HBM */
int gync_wc_normalize_dev(double* dLL_inout, const double* d_logprior, const int64_t* d_counts, int64_t K, int32_t M,
                          double* d_weights, double* d_density, void* stream);

/* ================= empirical-Bayes callers of the likelihood matrix (src/eb) ===================
 * ---- likelihoodmat_nanfast(xs, ys, d::MatrixNormalCentered; renormalize) — eb/likelihoodmat.jl:34-58
 *      (with sqeucdist_nan, :61-95).  xs: D x I (column i = vec(xs[i]), NaN = missing), ys: D x J, sigmas: D
 *      (vec(d.sigmas)).  L: I x J.  The _dev form takes element strides (k, column) so that sample-major device
 *      arrays need no transpose. */
int gync_likelihoodmat(const double* xs, int64_t I, const double* ys, int64_t J, int32_t D, const double* sigmas,
                       int32_t renormalize, double* L);
int gync_likelihoodmat_dev(const double* dA, int64_t I, int64_t sa_k, int64_t sa_i, const double* dB, int64_t J,
                           int64_t sb_k, int64_t sb_j, int32_t D, const double* d_sigmas, int32_t renormalize, double* dL,
                           void* stream);
/* ---- phi(x) = forwardsol(x)[:, measuredinds] on t = 0:30 — eb/gync.jl:11, eb/regularizers.jl:215.
 *      X: N x 116 -> Yphi: 124 x N (column n = vec of the 31 x 4 matrix), NaN column on solver failure. */
int gync_phi_batch(const double* X, int64_t N, const gync_solver_opts* opts, double* Yphi, int32_t* status);
/* ---- projectsimplex!(y) — eb/projectsimplex.jl:8-46 (in place) */
int gync_projectsimplex(double* y, int64_t K);
/* ---- LikelihoodModel — eb/likelihoodmodel.jl:3-14.  Builds and keeps in HBM
 *        Ld = likelihoodmat(ys, datas, measerr)     K x M   (em: likelihoodmodel.jl:17-20; logl/dlogl use its transpose)
 *        Lz = likelihoodmat(zs, ys, zsampledistr)   Z x K   (hz/dhz: regularizers.jl:26,60), Z = zmult * K, may be 0
 *      ys: D x K, datas: D x M, zs: D x Z, sig_*: D (sig_z NULL -> sig_meas, likelihoodmodel.jl:12-14). */
typedef struct gync_ebmodel gync_ebmodel;
int  gync_ebmodel_create(const double* ys, int64_t K, const double* datas, int64_t M, const double* zs, int64_t Z,
                         int32_t D, const double* sig_meas, const double* sig_z, gync_ebmodel** out);
int  gync_ebmodel_from_matrices(const double* Ld, int64_t K, int64_t M, const double* Lz, int64_t Z, gync_ebmodel** out);
void gync_ebmodel_destroy(gync_ebmodel* m);
int  gync_ebmodel_dims(gync_ebmodel* m, int64_t* K, int64_t* M, int64_t* Z);
int  gync_ebmodel_matrices(gync_ebmodel* m, double* Ld /* K x M or NULL */, double* Lz /* Z x K or NULL */);
/* resident matrices: Ld (K x M) and Lzt = Lz' (K x Z column-major: every z-row contiguous) */
int  gync_ebmodel_device_ptrs(gync_ebmodel* m, double** dLd, double** dLzt);
/* logl(m,w) = sum(log(L*w)), dlogl(m,w) = sum(L./(L*w),1) with L = Ld' — eb/regularizers.jl:3-10 */
int  gync_ebmodel_logl(gync_ebmodel* m, const double* w, double* out);
int  gync_ebmodel_dlogl(gync_ebmodel* m, const double* w, double* d /* K */);
/* hz(wx, Lzx, wz) — regularizers.jl:30-45; wz NULL -> repeatweights(w, zs) (:16-19) */
int  gync_ebmodel_hz(gync_ebmodel* m, const double* w, const double* wz, double* out);
/* dhz(m,w): cont == 0 dhzdiscr (:59-79, the reference's DHZCONT = false), else dhzcont (:50-57,149-160).
 * GYNC_ERR_NUMERIC when the result holds a NaN (:77). */
int  gync_ebmodel_dhz(gync_ebmodel* m, const double* w, int32_t cont, double* d /* K */);
/* em(m, w0, niter) — likelihoodmodel.jl:17-20: niter emiteration! steps on Ld, w in place */
int  gync_ebmodel_em(gync_ebmodel* m, double* w, int32_t niter, double* history);
/* gradientascent(dmple_obj(m,reg), w0, n, h) — eb/gradientascent.jl:1-25, likelihoodmodel.jl:33-64: nsteps times
 *   w <- movefromboundary(w, projectsimplex(w + h (reg dhz(w) + (1-reg) dlogl(w)))), all on the device.
 * history (nullable): K x nsteps, column i = weights after step i+1. */
int  gync_ebmodel_mple(gync_ebmodel* m, double* w, double reg, double h, int32_t nsteps, int32_t cont, double* history);
/* Device-level forms (device pointers, caller's stream, no synchronisation).  On a z-sharded model the results of
 * dhz_dev / hz_dev are PARTIAL sums over this device's z-rows: allreduce(sum) them over the ranks. */
int  gync_ebmodel_dhz_dev(gync_ebmodel* m, const double* dw, int32_t cont, double* d_out /* K */, void* stream);
int  gync_ebmodel_dlogl_dev(gync_ebmodel* m, const double* dw, double* d_out /* K */, void* stream);
int  gync_ebmodel_hz_dev(gync_ebmodel* m, const double* dw, double* d_out /* 1 */, void* stream);
/* one step w <- movefromboundary(w, projectsimplex(w + h (ca ga + cb gb))) — gradientascent.jl:3,19-25 */
int  gync_ebmodel_ascent_dev(gync_ebmodel* m, double* dw, const double* d_ga, double ca, const double* d_gb, double cb,
                             double h, void* stream);
int  gync_ebmodel_nan_flag(gync_ebmodel* m, void* stream);   /* 1 if a dhz held a NaN since the last call (regularizers.jl:77) */
/* z-sharded LikelihoodModel (one process per GPU): this device holds rows [z_off, z_off + Z_local) of the Z_total x K
 * z-matrix, the K x M data matrix is replicated.  The renormalisation (likelihoodmat.jl:53-54) is global: create returns
 * the local (min sq, max g) in scal_out; combine them over the ranks (min, max) and call finish. */
int  gync_ebmodel_create_zshard(const double* ys, int64_t K, const double* datas, int64_t M, const double* zs_local,
                                int64_t Z_local, int64_t z_off, int64_t Z_total, int32_t D, const double* sig_meas,
                                const double* sig_z, double* scal_out /* 2 */, gync_ebmodel** out);
int  gync_ebmodel_zshard_finish(gync_ebmodel* m, const double* scal /* 2 */);

/* ---- duplicate collapse at the head of WeightedChain(samplings, burnin) — weightedchain.jl:31-46 --------------------
 * X: K x 116, the concatenated chains after burn-in.  Consecutive equal rows become one row with a count.
 * rows_out: K x 116 buffer (leading dimension K) whose first *ngroups rows are the collapsed samples; counts_out: K. */
int gync_collapse_duplicates(const double* X, int64_t K, double* rows_out, int64_t* counts_out, int64_t* ngroups);

/* Row-sharded form (one process per GPU): each rank holds K_local rows; the caller's collectives sit
 * between the phases:  colstats -> allreduce(max) of d_stats[0..M] (column maxima, max log-prior) and allreduce(sum) of
 * d_stats[M+1] (count total) -> expsum(d_colmax = d_stats) -> allreduce(sum) of the M column sums -> rows (d_stats =
 * [colsum(M) | lpmax | sumcounts]) -> allreduce(sum) of d_total -> scale. */
int gync_wc_colstats_dev(const double* dLL, const double* d_logprior, const int64_t* d_counts, int64_t K, int32_t M,
                         double* d_stats /* M + 2 */, void* stream);
int gync_wc_expsum_dev(double* dLL_inout, int64_t K, int32_t M, const double* d_colmax, double* d_colsum, void* stream);
int gync_wc_rows_dev(double* dL_inout, const double* d_logprior, const int64_t* d_counts, int64_t K, int32_t M,
                     const double* d_stats, double* d_weights, double* d_density, double* d_total, void* stream);
int gync_wc_scale_dev(double* d_density, int64_t K, const double* d_total, void* stream);

/* FP64 flops of one Rosenbrock step attempt (counted from the generated code; roofline bookkeeping), and the order of
 * the method the library was built with (5: Rodas5P, the default; 4: RODAS4P). */
double gync_flops_per_step(void);
int gync_method_order(void);
/* FP64 FMA peak microbenchmark (TFLOP/s) used as the solver's roofline denominator. */
int gync_fp64_peak(double* tflops);

#ifdef __cplusplus
}
#endif
#endif
