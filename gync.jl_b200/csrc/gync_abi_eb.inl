// C ABI, part 3: empirical-Bayes callers of the likelihood matrix:
// likelihoodmat_nanfast, LikelihoodModel + logl/dlogl/hz/dhz, projectsimplex, gradient ascent (mple), gyncmodel's phi.
// (included at the end of gync_b200.cu)
#include "gync_eb.cuh"

namespace {

void lmat_finish(double* dL, int64_t n, const double* d_scal, cudaStream_t s) {
  const unsigned nb = (unsigned)std::min<int64_t>((n + 255) / 256, (int64_t)g_num_sms * 16);
  gync_lmat_exp_kernel<<<nb, 256, 0, s>>>(dL, n, d_scal);
  g_launches += 1;
}

// d_scal_out != NULL: stop after the (min sq, max g) reduction — the matrix then holds the logs of the unnormalised
// entries and the caller finishes it with lmat_finish once the scalars are global (row-sharded builds).
int lmat_launch(const double* dA, int64_t I, int64_t sa_k, int64_t sa_i, const double* dB, int64_t J, int64_t sb_k,
                int64_t sb_j, int D, const double* d_sig, int renorm, double* dL, cudaStream_t s, double* d_scal_out = nullptr) {
  if (I == 0 || J == 0) return GYNC_OK;
  const int64_t gx = (I + GYNC_LM_T - 1) / GYNC_LM_T, gy = (J + GYNC_LM_T - 1) / GYNC_LM_T;
  if (gy > 65535) return fail(GYNC_ERR_ARG, "likelihoodmat: more than 4.19 M columns");
  const int64_t nblk = gx * gy;
  ABuf red;
  CK(red.alloc(sizeof(double) * (2 * nblk + 2), s));
  double* bmin = red.as<double>();
  double* bmax = bmin + nblk;
  double* scal = bmax + nblk;
  CK(cudaFuncSetAttribute(gync_lmat_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(GyncLmSmem)));
  gync_lmat_kernel<<<dim3((unsigned)gx, (unsigned)gy), GYNC_LM_THREADS, sizeof(GyncLmSmem), s>>>(
      dA, I, sa_k, sa_i, dB, J, sb_k, sb_j, D, d_sig, renorm, dL, bmin, bmax);
  g_launches += 1;
  if (renorm) {
    gync_lmat_minmax_kernel<<<1, 1024, 0, s>>>(bmin, bmax, nblk, d_scal_out ? d_scal_out : scal);
    g_launches += 1;
    if (!d_scal_out) lmat_finish(dL, I * J, scal, s);
  }
  CK(cudaGetLastError());
  return GYNC_OK;
}

// y = A x (row sums): partials in `part`, number of partials returned
int mv_n_launch(const double* A, int64_t R, int64_t C, const double* x, double* part, int* nparts, cudaStream_t s) {
  const int64_t chunks = (C + GYNC_MV_CCHUNK - 1) / GYNC_MV_CCHUNK;
  if (chunks > 65535) return fail(GYNC_ERR_ARG, "matvec: too many columns");
  gync_gemv_n_kernel<<<dim3(blocks_for(R, GYNC_MV_THREADS), (unsigned)chunks), GYNC_MV_THREADS, 0, s>>>(A, R, C, R, x, part);
  g_launches += 1;
  *nparts = (int)chunks;
  CK(cudaGetLastError());
  return GYNC_OK;
}

int mv_t_rchunk(int64_t R, int64_t C) {
  // enough CTAs to fill the GPU four times over when the matrix allows it
  const int64_t colblocks = (C + GYNC_MV_TCOLS - 1) / GYNC_MV_TCOLS;
  const int64_t want_rb = std::max<int64_t>(1, (4 * (int64_t)g_num_sms + colblocks - 1) / colblocks);
  int64_t rc = (R + want_rb - 1) / want_rb;
  rc = (rc + 255) / 256 * 256;
  return (int)std::min<int64_t>(GYNC_MV_RCHUNK, std::max<int64_t>(256, rc));
}

// y = A' x (column sums)
int mv_t_launch(const double* A, int64_t R, int64_t C, const double* x, double* part, int* nparts, cudaStream_t s) {
  const int rchunk = mv_t_rchunk(R, C);
  const int64_t rb = (R + rchunk - 1) / rchunk, cb = (C + GYNC_MV_TCOLS - 1) / GYNC_MV_TCOLS;
  if (cb > 65535) return fail(GYNC_ERR_ARG, "matvec: too many columns");
  gync_gemv_t_kernel<<<dim3((unsigned)rb, (unsigned)cb), GYNC_MV_THREADS, 0, s>>>(A, R, C, R, x, rchunk, part);
  g_launches += 1;
  *nparts = (int)rb;
  CK(cudaGetLastError());
  return GYNC_OK;
}

int mv_finish(const double* part, int nparts, int64_t n, int epi, double* out, double* out2, const double* rho, int zmult,
              int* nan_flag, cudaStream_t s, int64_t z_off = 0, int64_t z_len = -1) {
  if (z_len < 0) z_len = n * (int64_t)std::max(zmult, 1);
  gync_mv_finish_kernel<<<blocks_for(n, 256 / GYNC_FIN_SLICES), 256, 0, s>>>(part, nparts, n, epi, out, out2, rho, zmult, nan_flag,
                                                                              z_off, z_len);
  g_launches += 1;
  CK(cudaGetLastError());
  return GYNC_OK;
}

size_t mv_part_doubles(int64_t R, int64_t C) {
  const int64_t n_parts = (C + GYNC_MV_CCHUNK - 1) / GYNC_MV_CCHUNK;
  const int64_t t_parts = (R + 255) / 256;     // upper bound: the smallest row chunk is 256
  return (size_t)std::max<int64_t>(n_parts * R, t_parts * C);
}

}  // namespace

// LikelihoodModel (eb/likelihoodmodel.jl:3-14) with both likelihood matrices resident in HBM:
//   Ld = likelihoodmat(ys, datas, measerr)      K x M   (its transpose is what logl/dlogl use, regularizers.jl:3,8)
//   Lz = likelihoodmat(zs, ys, zsampledistr)    Z x K   (hz/dhz, regularizers.jl:26,60) — held TRANSPOSED (Lzt, K x Z
//        column-major: every z-row contiguous) so that dhz can run in one pass over HBM (gync_dhz_fused_kernel)
struct gync_ebmodel {
  int device = -1;
  int64_t K = 0, M = 0, Z = 0;     // Z: rows of the z-matrix held HERE
  int64_t z_off = 0;               // z-sharded model: this device holds rows [z_off, z_off + Z) of Z_total = zmult K
  int zmult = 0;
  double *Ld = nullptr, *Lzt = nullptr;
  double *part = nullptr, *rho = nullptr, *fac = nullptr, *norms = nullptr, *rnorm = nullptr;
  double *ga = nullptr, *gb = nullptr, *y = nullptr, *w = nullptr, *scal = nullptr;
  GyncPsPart* ps = nullptr;
  int* flags = nullptr;     // [0] zero flag of the ascent step, [1] NaN flag of dhz, [2] projection passes
  int ps_grid = 0;
};

namespace {

void ebmodel_free(gync_ebmodel* m) {
  if (!m) return;
  int cur = 0;
  cudaGetDevice(&cur);
  if (m->device >= 0) cudaSetDevice(m->device);
  for (double* p : {m->Ld, m->Lzt, m->part, m->rho, m->fac, m->norms, m->rnorm, m->ga, m->gb, m->y, m->w, m->scal})
    if (p) cudaFree(p);
  if (m->ps) cudaFree(m->ps);
  if (m->flags) cudaFree(m->flags);
  cudaSetDevice(cur);
  delete m;
}

int ps_grid_size(int64_t K, int* grid) {
  int per_sm = 0;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, gync_ascent_step_kernel, GYNC_PS_THREADS, 0));
  if (per_sm < 1) return fail(GYNC_ERR_CUDA, "ascent kernel does not fit on an SM");
  *grid = (int)std::max<int64_t>(1, std::min<int64_t>((int64_t)g_num_sms, (K + GYNC_PS_THREADS - 1) / GYNC_PS_THREADS));
  return GYNC_OK;
}

int ebmodel_alloc(gync_ebmodel* m) {
  m->device = g_device;
  const int64_t K = m->K, M = m->M, Z = m->Z;
  CK(cudaMalloc((void**)&m->Ld, sizeof(double) * std::max<int64_t>(K * M, 1)));
  if (Z > 0) CK(cudaMalloc((void**)&m->Lzt, sizeof(double) * Z * K));
  size_t pd = mv_part_doubles(K, std::max<int64_t>(M, 1));
  if (Z > 0) pd = std::max(std::max(pd, mv_part_doubles(K, Z)), (size_t)g_num_sms * (size_t)K);
  CK(cudaMalloc((void**)&m->part, sizeof(double) * pd));
  CK(cudaMalloc((void**)&m->rho, sizeof(double) * std::max<int64_t>(Z, 1)));
  CK(cudaMalloc((void**)&m->fac, sizeof(double) * std::max<int64_t>(Z, 1)));
  CK(cudaMalloc((void**)&m->norms, sizeof(double) * std::max<int64_t>(M, 1)));
  CK(cudaMalloc((void**)&m->rnorm, sizeof(double) * std::max<int64_t>(M, 1)));
  CK(cudaMalloc((void**)&m->ga, sizeof(double) * K));
  CK(cudaMalloc((void**)&m->gb, sizeof(double) * K));
  CK(cudaMalloc((void**)&m->y, sizeof(double) * K));
  CK(cudaMalloc((void**)&m->w, sizeof(double) * K));
  CK(cudaMalloc((void**)&m->scal, sizeof(double) * 4));
  if (int rc = ps_grid_size(K, &m->ps_grid)) return rc;
  CK(cudaMalloc((void**)&m->ps, sizeof(GyncPsPart) * 2 * m->ps_grid));
  CK(cudaMalloc((void**)&m->flags, sizeof(int) * 4));
  CK(cudaMemset(m->flags, 0, sizeof(int) * 4));
  return GYNC_OK;
}

// norms = Ld' w, rnorm = 1 ./ norms
int eb_norms(gync_ebmodel* m, const double* dw, cudaStream_t s) {
  int np = 0;
  if (int rc = mv_t_launch(m->Ld, m->K, m->M, dw, m->part, &np, s)) return rc;
  return mv_finish(m->part, np, m->M, GYNC_EPI_RECIP, m->norms, m->rnorm, nullptr, 0, nullptr, s);
}

// dlogl(w) = sum(L ./ (L*w), 1) with L = Ld' (regularizers.jl:7-10)  ->  out (K)
int eb_dlogl(gync_ebmodel* m, const double* dw, double* out, cudaStream_t s) {
  if (int rc = eb_norms(m, dw, s)) return rc;
  int np = 0;
  if (int rc = mv_n_launch(m->Ld, m->K, m->M, m->rnorm, m->part, &np, s)) return rc;
  return mv_finish(m->part, np, m->K, GYNC_EPI_PLAIN, out, nullptr, nullptr, 0, nullptr, s);
}

// rho = Lz w (regularizers.jl:36,65): column sums of Lzt
int eb_rho(gync_ebmodel* m, const double* dw, cudaStream_t s) {
  int np = 0;
  if (int rc = mv_t_launch(m->Lzt, m->K, m->Z, dw, m->part, &np, s)) return rc;
  return mv_finish(m->part, np, m->Z, GYNC_EPI_PLAIN, m->rho, nullptr, nullptr, 0, nullptr, s);
}

// one-pass form available?  (even K for 16-byte rows; accumulator + w in shared memory)
bool eb_dhz_fusable(const gync_ebmodel* m, size_t* smem) {
  *smem = sizeof(double) * 2 * (size_t)m->K;
  int max_optin = 0;
  if (cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, g_device) != cudaSuccess) return false;
  return (m->K % 2 == 0) && *smem + 2048 <= (size_t)max_optin && !getenv("GYNC_DHZ_TWO_PASS");
}

// dhzdiscr / dhzcont (regularizers.jl:59-79, 50-57 + 149-160)  ->  out (K); sets flags[1] on NaN
int eb_dhz(gync_ebmodel* m, const double* dw, int cont, double* out, cudaStream_t s) {
  size_t smem = 0;
  if (eb_dhz_fusable(m, &smem)) {
    const int nb = (int)std::min<int64_t>(g_num_sms, (m->Z + GYNC_DHZ_ROWS - 1) / GYNC_DHZ_ROWS);
    CK(cudaFuncSetAttribute(gync_dhz_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    gync_dhz_fused_kernel<<<nb, GYNC_DHZ_THREADS, smem, s>>>(m->Lzt, m->K, m->Z, m->z_off, dw, m->zmult, cont, m->rho, m->part);
    g_launches += 1;
    CK(cudaGetLastError());
    return mv_finish(m->part, nb, m->K, cont ? GYNC_EPI_DHZ_CONT : GYNC_EPI_DHZ_DISCR, out, nullptr, m->rho, m->zmult,
                     m->flags + 1, s, m->z_off, m->Z);
  }
  if (int rc = eb_rho(m, dw, s)) return rc;
  gync_dhz_factors_kernel<<<blocks_for(m->Z, 256), 256, 0, s>>>(dw, m->K, m->zmult, m->rho, m->Z, m->z_off, cont, m->fac);
  g_launches += 1;
  int np = 0;
  if (int rc = mv_n_launch(m->Lzt, m->K, m->Z, m->fac, m->part, &np, s)) return rc;
  return mv_finish(m->part, np, m->K, cont ? GYNC_EPI_DHZ_CONT : GYNC_EPI_DHZ_DISCR, out, nullptr, m->rho, m->zmult,
                   m->flags + 1, s, m->z_off, m->Z);
}

// one projected step on the device weights (in place); ga/gb may be NULL (pure projection)
int eb_ascent_step(gync_ebmodel* m, double* dw, const double* ga, double ca, const double* gb, double cb, double h,
                   double relstep, double* d_hist, cudaStream_t s) {
  CK(cudaMemsetAsync(m->flags, 0, sizeof(int), s));
  int64_t K = m->K;
  GyncPsPart* ps = m->ps;
  int* zf = m->flags;
  int* npass = m->flags + 2;
  double* y = m->y;
  void* args[] = {&dw, &ga, &ca, &gb, &cb, &h, &K, &relstep, &y, &ps, &zf, &d_hist, &npass};
  CK(cudaLaunchCooperativeKernel((void*)gync_ascent_step_kernel, dim3(m->ps_grid), dim3(GYNC_PS_THREADS), args, 0, s));
  g_launches += 1;
  return GYNC_OK;
}

// host-level hz / dhz / mple return complete results: on a z-sharded model they would silently be partial sums
int eb_check_unsharded(const gync_ebmodel* m, const char* what) {
  if (m->Z != (int64_t)m->zmult * m->K)
    return fail(GYNC_ERR_ARG, std::string(what) + ": z-sharded model — use the *_dev entry points and allreduce the partial results");
  return GYNC_OK;
}

int eb_check_model(gync_ebmodel* m, const void* w, const char* what) {
  if (!m || !w) return fail(GYNC_ERR_ARG, std::string(what) + ": NULL argument");
  if (int rc = ensure_init()) return rc;
  if (m->device != g_device) return fail(GYNC_ERR_ARG, std::string(what) + ": model lives on another device");
  return GYNC_OK;
}

}  // namespace

// emiteration! for more than 64 data columns (em.jl:27-41 as written: norms = L'w, then w[k] <- w[k]/M * (L (1 ./ norms))[k]);
// the persistent single-pass kernel keeps M/2 accumulators per thread and stops at M = 64.
extern "C" int gync_em_iterate_generic(double* dw, const double* dL, int64_t K, int M, int niter, double* d_hist, cudaStream_t s) {
  ABuf part, nr;
  CK(part.alloc(sizeof(double) * mv_part_doubles(K, M), s));
  CK(nr.alloc(sizeof(double) * (2 * (size_t)M + (size_t)K), s));
  double* norms = nr.as<double>();
  double* rnorm = norms + M;
  for (int it = 0; it < niter; ++it) {
    int np = 0;
    if (int rc = mv_t_launch(dL, K, M, dw, part.as<double>(), &np, s)) return rc;
    if (int rc = mv_finish(part.as<double>(), np, M, GYNC_EPI_RECIP, norms, rnorm, nullptr, 0, nullptr, s)) return rc;
    if (int rc = mv_n_launch(dL, K, M, rnorm, part.as<double>(), &np, s)) return rc;
    if (int rc = mv_finish(part.as<double>(), np, K, GYNC_EPI_EM, dw, nullptr, nullptr, M, nullptr, s)) return rc;
    if (d_hist) CK(cudaMemcpyAsync(d_hist + (size_t)it * K, dw, sizeof(double) * K, cudaMemcpyDeviceToDevice, s));
  }
  return GYNC_OK;
}

extern "C" {

// ---- likelihoodmat_nanfast ---------------------------------------------------------------------
int gync_likelihoodmat_dev(const double* dA, int64_t I, int64_t sa_k, int64_t sa_i, const double* dB, int64_t J,
                           int64_t sb_k, int64_t sb_j, int32_t D, const double* d_sigmas, int32_t renormalize, double* dL,
                           void* stream) {
  if (I < 0 || J < 0 || D < 1 || !d_sigmas || (I > 0 && J > 0 && (!dA || !dB || !dL)))
    return fail(GYNC_ERR_ARG, "gync_likelihoodmat: bad arguments");
  if (int rc = ensure_init()) return rc;
  return lmat_launch(dA, I, sa_k, sa_i, dB, J, sb_k, sb_j, D, d_sigmas, renormalize, dL, (cudaStream_t)stream);
}

int gync_likelihoodmat(const double* xs, int64_t I, const double* ys, int64_t J, int32_t D, const double* sigmas,
                       int32_t renormalize, double* L) {
  if (I < 0 || J < 0 || D < 1 || !sigmas || (I > 0 && J > 0 && (!xs || !ys || !L)))
    return fail(GYNC_ERR_ARG, "gync_likelihoodmat: bad arguments");
  if (int rc = ensure_init()) return rc;
  if (I == 0 || J == 0) return GYNC_OK;
  DBuf dA, dB, dS, dL;
  CK(dA.alloc(sizeof(double) * D * I));
  CK(dB.alloc(sizeof(double) * D * J));
  CK(dS.alloc(sizeof(double) * D));
  CK(dL.alloc(sizeof(double) * I * J));
  CK(cudaMemcpyAsync(dA.p, xs, sizeof(double) * D * I, cudaMemcpyHostToDevice, g_stream));
  CK(cudaMemcpyAsync(dB.p, ys, sizeof(double) * D * J, cudaMemcpyHostToDevice, g_stream));
  CK(cudaMemcpyAsync(dS.p, sigmas, sizeof(double) * D, cudaMemcpyHostToDevice, g_stream));
  if (int rc = lmat_launch(dA.as<double>(), I, 1, D, dB.as<double>(), J, 1, D, D, dS.as<double>(), renormalize,
                           dL.as<double>(), g_stream)) return rc;
  CK(cudaMemcpyAsync(L, dL.p, sizeof(double) * I * J, cudaMemcpyDeviceToHost, g_stream));
  CK(wait_stream(g_stream));
  return GYNC_OK;
}

// ---- phi(x) = forwardsol(x)[:, measuredinds] on t = 0:30 (eb/gync.jl:11, regularizers.jl:215) ----
int gync_phi_batch(const double* X, int64_t N, const gync_solver_opts* opts, double* Yphi, int32_t* status) {
  if (N < 0 || (N > 0 && (!X || !Yphi))) return fail(GYNC_ERR_ARG, "gync_phi_batch: bad arguments");
  if (int rc = ensure_init()) return rc;
  if (N == 0) return GYNC_OK;
  const int nT = 31;
  double t[31];
  for (int k = 0; k < nT; ++k) t[k] = (double)k;
  DBuf dX, dT, dY, dSt, dP;
  CK(dX.alloc(sizeof(double) * N * GYNC_NX));
  CK(dT.alloc(sizeof(double) * nT));
  CK(dY.alloc(sizeof(double) * N * nT * GYNC_NY));
  CK(dSt.alloc(sizeof(int32_t) * N));
  CK(dP.alloc(sizeof(double) * N * 124));
  CK(cudaMemcpyAsync(dX.p, X, sizeof(double) * N * GYNC_NX, cudaMemcpyHostToDevice, g_stream));
  CK(cudaMemcpyAsync(dT.p, t, sizeof(double) * nT, cudaMemcpyHostToDevice, g_stream));
  if (int rc = launch_solve_grid(dX.as<double>(), nullptr, nullptr, N, dT.as<double>(), nT, to_opts(opts), dY.as<double>(),
                                 dSt.as<int32_t>(), g_stream)) return rc;
  gync_pick_measured_kernel<<<blocks_for(N * 124, 256), 256, 0, g_stream>>>(dY.as<double>(), N, nT, dP.as<double>());
  g_launches += 1;
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(Yphi, dP.p, sizeof(double) * N * 124, cudaMemcpyDeviceToHost, g_stream));
  if (status) CK(cudaMemcpyAsync(status, dSt.p, sizeof(int32_t) * N, cudaMemcpyDeviceToHost, g_stream));
  CK(wait_stream(g_stream));
  return GYNC_OK;
}

// ---- projectsimplex!(y) ------------------------------------------------------------------------
int gync_projectsimplex(double* yv, int64_t K) {
  if (K < 1 || !yv) return fail(GYNC_ERR_ARG, "gync_projectsimplex: bad arguments");
  if (int rc = ensure_init()) return rc;
  gync_ebmodel tmp;
  tmp.K = K;
  tmp.device = g_device;
  DBuf dw, dy, dps, dfl;
  if (int rc = ps_grid_size(K, &tmp.ps_grid)) return rc;
  CK(dw.alloc(sizeof(double) * K));
  CK(dy.alloc(sizeof(double) * K));
  CK(dps.alloc(sizeof(GyncPsPart) * 2 * tmp.ps_grid));
  CK(dfl.alloc(sizeof(int) * 4));
  tmp.y = dy.as<double>();
  tmp.ps = dps.as<GyncPsPart>();
  tmp.flags = dfl.as<int>();
  CK(cudaMemcpyAsync(dw.p, yv, sizeof(double) * K, cudaMemcpyHostToDevice, g_stream));
  // relstep = 1: movefromboundary is the identity, the result is the projection itself
  if (int rc = eb_ascent_step(&tmp, dw.as<double>(), nullptr, 0.0, nullptr, 0.0, 0.0, 1.0, nullptr, g_stream)) return rc;
  CK(cudaMemcpyAsync(yv, dw.p, sizeof(double) * K, cudaMemcpyDeviceToHost, g_stream));
  CK(wait_stream(g_stream));
  return GYNC_OK;
}

// ---- LikelihoodModel ---------------------------------------------------------------------------
void gync_ebmodel_destroy(gync_ebmodel* m) { ebmodel_free(m); }

static int ebmodel_create_impl(const double* ys, int64_t K, const double* datas, int64_t M, const double* zs, int64_t Z,
                               int64_t z_off, int64_t Z_total, int32_t D, const double* sig_meas, const double* sig_z,
                               double* scal_out, gync_ebmodel** out) {
  if (K < 1 || M < 0 || Z < 0 || D < 1 || !ys || !sig_meas || !out || (M > 0 && !datas) || (Z > 0 && !zs) || (Z_total % K) != 0 ||
      z_off < 0 || z_off + Z > Z_total)
    return fail(GYNC_ERR_ARG, "gync_ebmodel_create: bad arguments (length(zs) must be a multiple of length(ys))");
  if (int rc = ensure_init()) return rc;
  gync_ebmodel* m = new gync_ebmodel();
  m->K = K; m->M = M; m->Z = Z; m->z_off = z_off; m->zmult = (int)(Z_total / K);
  int rc = ebmodel_alloc(m);
  DBuf dY, dB, dS;
  auto up = [&](DBuf& b, const double* h, size_t n) -> int {
    CK(b.alloc(sizeof(double) * n));
    CK(cudaMemcpyAsync(b.p, h, sizeof(double) * n, cudaMemcpyHostToDevice, g_stream));
    return GYNC_OK;
  };
  if (!rc) rc = up(dY, ys, (size_t)D * K);
  if (!rc) rc = up(dS, sig_meas, (size_t)D);
  if (!rc && M > 0) {
    rc = up(dB, datas, (size_t)D * M);
    if (!rc) rc = lmat_launch(dY.as<double>(), K, 1, D, dB.as<double>(), M, 1, D, D, dS.as<double>(), 1, m->Ld, g_stream);
  }
  if (!rc && Z > 0) {
    DBuf dZ, dSz;
    rc = up(dZ, zs, (size_t)D * Z);
    if (!rc && sig_z) rc = up(dSz, sig_z, (size_t)D);
    // likelihoodmat(zs, ys)' == likelihoodmat(ys, zs): built directly in the transposed layout
    if (!rc) rc = lmat_launch(dY.as<double>(), K, 1, D, dZ.as<double>(), Z, 1, D, D, sig_z ? dSz.as<double>() : dS.as<double>(),
                              1, m->Lzt, g_stream, scal_out ? m->scal + 2 : nullptr);
    if (!rc && scal_out && cudaMemcpyAsync(scal_out, m->scal + 2, 2 * sizeof(double), cudaMemcpyDeviceToHost, g_stream) != cudaSuccess)
      rc = fail(GYNC_ERR_CUDA, "gync_ebmodel_create: copy of the renormalisation scalars failed");
    if (!rc && cudaStreamSynchronize(g_stream) != cudaSuccess) rc = fail(GYNC_ERR_CUDA, "gync_ebmodel_create: kernel failed");
  }
  if (!rc && cudaStreamSynchronize(g_stream) != cudaSuccess) rc = fail(GYNC_ERR_CUDA, "gync_ebmodel_create: kernel failed");
  if (rc) { ebmodel_free(m); return rc; }
  *out = m;
  return GYNC_OK;
}

int gync_ebmodel_create(const double* ys, int64_t K, const double* datas, int64_t M, const double* zs, int64_t Z,
                        int32_t D, const double* sig_meas, const double* sig_z, gync_ebmodel** out) {
  return ebmodel_create_impl(ys, K, datas, M, zs, Z, 0, Z, D, sig_meas, sig_z, nullptr, out);
}

// z-sharded LikelihoodModel (one process per GPU): this device holds rows [z_off, z_off + Z_local) of the z-matrix, the
// (small) data matrix is replicated.  The renormalisation of likelihoodmat.jl:53-54 is global, so the build stops after the
// local (min sq, max g) — scal_out — and gync_ebmodel_zshard_finish takes the values combined over ranks (min, max).
int gync_ebmodel_create_zshard(const double* ys, int64_t K, const double* datas, int64_t M, const double* zs_local,
                               int64_t Z_local, int64_t z_off, int64_t Z_total, int32_t D, const double* sig_meas,
                               const double* sig_z, double* scal_out /* 2 */, gync_ebmodel** out) {
  if (!scal_out || Z_local < 1) return fail(GYNC_ERR_ARG, "gync_ebmodel_create_zshard: bad arguments");
  return ebmodel_create_impl(ys, K, datas, M, zs_local, Z_local, z_off, Z_total, D, sig_meas, sig_z, scal_out, out);
}

int gync_ebmodel_zshard_finish(gync_ebmodel* m, const double* scal /* 2: global min sq, global max g */) {
  if (int rc = eb_check_model(m, scal, "gync_ebmodel_zshard_finish")) return rc;
  CK(cudaMemcpyAsync(m->scal + 2, scal, 2 * sizeof(double), cudaMemcpyHostToDevice, g_stream));
  lmat_finish(m->Lzt, m->Z * m->K, m->scal + 2, g_stream);
  CK(cudaGetLastError());
  CK(wait_stream(g_stream));
  return GYNC_OK;
}

int gync_ebmodel_from_matrices(const double* Ld, int64_t K, int64_t M, const double* Lz, int64_t Z, gync_ebmodel** out) {
  if (K < 1 || M < 0 || Z < 0 || !out || (M > 0 && !Ld) || (Z > 0 && !Lz) || (Z % K) != 0)
    return fail(GYNC_ERR_ARG, "gync_ebmodel_from_matrices: bad arguments");
  if (int rc = ensure_init()) return rc;
  gync_ebmodel* m = new gync_ebmodel();
  m->K = K; m->M = M; m->Z = Z; m->zmult = (int)(Z / K);
  int rc = ebmodel_alloc(m);
  if (!rc && M > 0 && cudaMemcpyAsync(m->Ld, Ld, sizeof(double) * K * M, cudaMemcpyHostToDevice, g_stream) != cudaSuccess) rc = fail(GYNC_ERR_CUDA, "upload Ld");
  DBuf tmp;
  if (!rc && Z > 0) {
    if (tmp.alloc(sizeof(double) * Z * K) != cudaSuccess ||
        cudaMemcpyAsync(tmp.p, Lz, sizeof(double) * Z * K, cudaMemcpyHostToDevice, g_stream) != cudaSuccess) rc = fail(GYNC_ERR_CUDA, "upload Lz");
    else {
      gync_transpose_kernel<<<dim3(blocks_for(Z, 32), blocks_for(K, 32)), 256, 0, g_stream>>>(tmp.as<double>(), Z, K, m->Lzt);
      g_launches += 1;
    }
  }
  if (!rc && cudaStreamSynchronize(g_stream) != cudaSuccess) rc = fail(GYNC_ERR_CUDA, "gync_ebmodel_from_matrices: copy failed");
  if (rc) { ebmodel_free(m); return rc; }
  *out = m;
  return GYNC_OK;
}

int gync_ebmodel_dims(gync_ebmodel* m, int64_t* K, int64_t* M, int64_t* Z) {
  if (!m) return fail(GYNC_ERR_ARG, "gync_ebmodel_dims: NULL");
  if (K) *K = m->K;
  if (M) *M = m->M;
  if (Z) *Z = m->Z;
  return GYNC_OK;
}

int gync_ebmodel_matrices(gync_ebmodel* m, double* Ld, double* Lz) {
  if (int rc = eb_check_model(m, m, "gync_ebmodel_matrices")) return rc;
  if (Ld && m->M > 0) CK(cudaMemcpyAsync(Ld, m->Ld, sizeof(double) * m->K * m->M, cudaMemcpyDeviceToHost, g_stream));
  DBuf tmp;
  if (Lz && m->Z > 0) {
    CK(tmp.alloc(sizeof(double) * m->Z * m->K));
    gync_transpose_kernel<<<dim3(blocks_for(m->K, 32), blocks_for(m->Z, 32)), 256, 0, g_stream>>>(m->Lzt, m->K, m->Z, tmp.as<double>());
    g_launches += 1;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(Lz, tmp.p, sizeof(double) * m->Z * m->K, cudaMemcpyDeviceToHost, g_stream));
  }
  CK(wait_stream(g_stream));
  return GYNC_OK;
}

// device pointers of the resident matrices (for callers that keep working on the device: EM, bench)
int gync_ebmodel_device_ptrs(gync_ebmodel* m, double** dLd, double** dLzt) {
  if (!m) return fail(GYNC_ERR_ARG, "gync_ebmodel_device_ptrs: NULL");
  if (dLd) *dLd = m->Ld;
  if (dLzt) *dLzt = m->Lzt;
  return GYNC_OK;
}

int gync_ebmodel_logl(gync_ebmodel* m, const double* w, double* out) {
  if (int rc = eb_check_model(m, w, "gync_ebmodel_logl")) return rc;
  if (!out || m->M < 1) return fail(GYNC_ERR_ARG, "gync_ebmodel_logl: no data matrix / NULL output");
  CK(cudaMemcpyAsync(m->w, w, sizeof(double) * m->K, cudaMemcpyHostToDevice, g_stream));
  if (int rc = eb_norms(m, m->w, g_stream)) return rc;
  gync_logsum_kernel<<<1, 1024, 0, g_stream>>>(m->norms, m->M, nullptr, 1, 0, 1.0, 1.0, 0, m->scal);
  g_launches += 1;
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(out, m->scal, sizeof(double), cudaMemcpyDeviceToHost, g_stream));
  CK(wait_stream(g_stream));
  return GYNC_OK;
}

int gync_ebmodel_dlogl(gync_ebmodel* m, const double* w, double* d) {
  if (int rc = eb_check_model(m, w, "gync_ebmodel_dlogl")) return rc;
  if (!d || m->M < 1) return fail(GYNC_ERR_ARG, "gync_ebmodel_dlogl: no data matrix / NULL output");
  CK(cudaMemcpyAsync(m->w, w, sizeof(double) * m->K, cudaMemcpyHostToDevice, g_stream));
  if (int rc = eb_dlogl(m, m->w, m->ga, g_stream)) return rc;
  CK(cudaMemcpyAsync(d, m->ga, sizeof(double) * m->K, cudaMemcpyDeviceToHost, g_stream));
  CK(wait_stream(g_stream));
  return GYNC_OK;
}

// hz(wx, Lzx, wz) (regularizers.jl:30-45); wz == NULL -> repeatweights(w, zs)
int gync_ebmodel_hz(gync_ebmodel* m, const double* w, const double* wz, double* out) {
  if (int rc = eb_check_model(m, w, "gync_ebmodel_hz")) return rc;
  if (!out || m->Z < 1) return fail(GYNC_ERR_ARG, "gync_ebmodel_hz: no z matrix / NULL output");
  if (int rc = eb_check_unsharded(m, "gync_ebmodel_hz")) return rc;
  CK(cudaMemcpyAsync(m->w, w, sizeof(double) * m->K, cudaMemcpyHostToDevice, g_stream));
  if (int rc = eb_rho(m, m->w, g_stream)) return rc;
  if (wz) {
    CK(cudaMemcpyAsync(m->fac, wz, sizeof(double) * m->Z, cudaMemcpyHostToDevice, g_stream));
    gync_logsum_kernel<<<1, 1024, 0, g_stream>>>(m->rho, m->Z, m->fac, m->Z, 0, 1.0, -1.0, 1, m->scal);
  } else {
    gync_logsum_kernel<<<1, 1024, 0, g_stream>>>(m->rho, m->Z, m->w, m->K, m->z_off, (double)m->zmult, -1.0, 1, m->scal);
  }
  g_launches += 1;
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(out, m->scal, sizeof(double), cudaMemcpyDeviceToHost, g_stream));
  CK(wait_stream(g_stream));
  return GYNC_OK;
}

// dhz(w) = DHZCONT ? dhzcont : dhzdiscr (regularizers.jl:47-79); a NaN in the result is an error (:77)
int gync_ebmodel_dhz(gync_ebmodel* m, const double* w, int32_t cont, double* d) {
  if (int rc = eb_check_model(m, w, "gync_ebmodel_dhz")) return rc;
  if (!d || m->Z < 1) return fail(GYNC_ERR_ARG, "gync_ebmodel_dhz: no z matrix / NULL output");
  if (int rc = eb_check_unsharded(m, "gync_ebmodel_dhz")) return rc;
  CK(cudaMemcpyAsync(m->w, w, sizeof(double) * m->K, cudaMemcpyHostToDevice, g_stream));
  CK(cudaMemsetAsync(m->flags + 1, 0, sizeof(int), g_stream));
  if (int rc = eb_dhz(m, m->w, cont, m->gb, g_stream)) return rc;
  int nanflag = 0;
  CK(cudaMemcpyAsync(d, m->gb, sizeof(double) * m->K, cudaMemcpyDeviceToHost, g_stream));
  CK(cudaMemcpyAsync(&nanflag, m->flags + 1, sizeof(int), cudaMemcpyDeviceToHost, g_stream));
  CK(wait_stream(g_stream));
  if (nanflag) return fail(GYNC_ERR_NUMERIC, "encountered NaN in dhz");
  return GYNC_OK;
}

// em(m, w0, niter) (likelihoodmodel.jl:17-20): emiterations on the resident data matrix; w in place, niter steps
int gync_ebmodel_em(gync_ebmodel* m, double* w, int32_t niter, double* history) {
  if (int rc = eb_check_model(m, w, "gync_ebmodel_em")) return rc;
  if (niter < 0 || m->M < 1) return fail(GYNC_ERR_ARG, "gync_ebmodel_em: no data matrix / niter < 0");
  if (niter == 0) return GYNC_OK;
  DBuf dH;
  if (history) CK(dH.alloc(sizeof(double) * m->K * niter));
  CK(cudaMemcpyAsync(m->w, w, sizeof(double) * m->K, cudaMemcpyHostToDevice, g_stream));
  if (int rc = gync_em_iterate_dev(m->w, m->Ld, m->K, (int32_t)m->M, niter, history ? dH.as<double>() : nullptr, g_stream)) return rc;
  CK(cudaMemcpyAsync(w, m->w, sizeof(double) * m->K, cudaMemcpyDeviceToHost, g_stream));
  if (history) CK(cudaMemcpyAsync(history, dH.p, sizeof(double) * m->K * niter, cudaMemcpyDeviceToHost, g_stream));
  CK(wait_stream(g_stream));
  return GYNC_OK;
}

// gradientascent(dmple_obj(m, reg), w0, niter, h) (gradientascent.jl:1-5, likelihoodmodel.jl:33-64), device-resident:
// nsteps iterations of  w <- movefromboundary(w, projectsimplex(w + h (reg dhz(w) + (1-reg) dlogl(w)))).
// history (nullable): K x nsteps, column i = weights after step i+1.
int gync_ebmodel_mple(gync_ebmodel* m, double* w, double reg, double h, int32_t nsteps, int32_t cont, double* history) {
  if (int rc = eb_check_model(m, w, "gync_ebmodel_mple")) return rc;
  if (nsteps < 0 || (reg != 0.0 && m->Z < 1) || (reg != 1.0 && m->M < 1))
    return fail(GYNC_ERR_ARG, "gync_ebmodel_mple: the model lacks the matrix this regularisation weight needs");
  if (reg != 0.0) if (int rc = eb_check_unsharded(m, "gync_ebmodel_mple")) return rc;
  if (nsteps == 0) return GYNC_OK;
  cudaStream_t s = g_stream;
  DBuf dH;
  if (history) CK(dH.alloc(sizeof(double) * m->K * nsteps));
  CK(cudaMemcpyAsync(m->w, w, sizeof(double) * m->K, cudaMemcpyHostToDevice, s));
  CK(cudaMemsetAsync(m->flags + 1, 0, sizeof(int), s));
  for (int it = 0; it < nsteps; ++it) {
    const double* ga = nullptr;
    const double* gb = nullptr;
    if (reg != 0.0) { if (int rc = eb_dhz(m, m->w, cont, m->ga, s)) return rc; ga = m->ga; }
    if (reg != 1.0) { if (int rc = eb_dlogl(m, m->w, m->gb, s)) return rc; gb = m->gb; }
    const double ca = (reg == 1.0) ? 1.0 : reg, cb = (reg == 0.0) ? 1.0 : 1.0 - reg;   // likelihoodmodel.jl:57-63
    if (int rc = eb_ascent_step(m, m->w, ga, ca, gb, cb, h, 0.5, history ? dH.as<double>() + (size_t)it * m->K : nullptr, s)) return rc;
  }
  int nanflag = 0;
  CK(cudaMemcpyAsync(w, m->w, sizeof(double) * m->K, cudaMemcpyDeviceToHost, s));
  if (history) CK(cudaMemcpyAsync(history, dH.p, sizeof(double) * m->K * nsteps, cudaMemcpyDeviceToHost, s));
  CK(cudaMemcpyAsync(&nanflag, m->flags + 1, sizeof(int), cudaMemcpyDeviceToHost, s));
  CK(wait_stream(s));
  if (nanflag) return fail(GYNC_ERR_NUMERIC, "encountered NaN in dhz");
  return GYNC_OK;
}

// ---- device-level entry points (device pointers, caller's stream, no synchronisation): what a host that keeps the
// weights on the GPU calls — bench.py, and the z-sharded loop of dist.ShardedEB where the results of dhz / hz are PARTIAL
// sums over this device's z-rows, to be allreduce(sum)-ed by the caller.
int gync_ebmodel_dhz_dev(gync_ebmodel* m, const double* dw, int32_t cont, double* d_out, void* stream) {
  if (int rc = eb_check_model(m, dw, "gync_ebmodel_dhz_dev")) return rc;
  if (!d_out || m->Z < 1) return fail(GYNC_ERR_ARG, "gync_ebmodel_dhz_dev: no z matrix / NULL output");
  return eb_dhz(m, dw, cont, d_out, (cudaStream_t)stream);
}

int gync_ebmodel_dlogl_dev(gync_ebmodel* m, const double* dw, double* d_out, void* stream) {
  if (int rc = eb_check_model(m, dw, "gync_ebmodel_dlogl_dev")) return rc;
  if (!d_out || m->M < 1) return fail(GYNC_ERR_ARG, "gync_ebmodel_dlogl_dev: no data matrix / NULL output");
  return eb_dlogl(m, dw, d_out, (cudaStream_t)stream);
}

int gync_ebmodel_hz_dev(gync_ebmodel* m, const double* dw, double* d_out /* 1 */, void* stream) {
  if (int rc = eb_check_model(m, dw, "gync_ebmodel_hz_dev")) return rc;
  if (!d_out || m->Z < 1) return fail(GYNC_ERR_ARG, "gync_ebmodel_hz_dev: no z matrix / NULL output");
  cudaStream_t s = (cudaStream_t)stream;
  if (int rc = eb_rho(m, dw, s)) return rc;
  gync_logsum_kernel<<<1, 1024, 0, s>>>(m->rho, m->Z, dw, m->K, m->z_off, (double)m->zmult, -1.0, 1, d_out);
  g_launches += 1;
  CK(cudaGetLastError());
  return GYNC_OK;
}

// one step  w <- movefromboundary(w, projectsimplex(w + h (ca ga + cb gb)))  on device vectors (gb / ga may be NULL)
int gync_ebmodel_ascent_dev(gync_ebmodel* m, double* dw, const double* d_ga, double ca, const double* d_gb, double cb, double h,
                            void* stream) {
  if (int rc = eb_check_model(m, dw, "gync_ebmodel_ascent_dev")) return rc;
  return eb_ascent_step(m, dw, d_ga, ca, d_gb, cb, h, 0.5, nullptr, (cudaStream_t)stream);
}

// 1 if a NaN appeared in a dhz evaluation since the last call (regularizers.jl:77); synchronises the stream
int gync_ebmodel_nan_flag(gync_ebmodel* m, void* stream) {
  if (!m) return -1;
  int f = 0;
  if (cudaMemcpyAsync(&f, m->flags + 1, sizeof(int), cudaMemcpyDeviceToHost, (cudaStream_t)stream) != cudaSuccess) return -2;
  if (cudaMemsetAsync(m->flags + 1, 0, sizeof(int), (cudaStream_t)stream) != cudaSuccess) return -2;
  if (cudaStreamSynchronize((cudaStream_t)stream) != cudaSuccess) return -2;
  return f;
}

}  // extern "C"
