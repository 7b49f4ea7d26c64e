// C ABI, part 2: log-prior, many-chain Metropolis, WeightedChain normalisation.
// (included at the end of gync_b200.cu)
extern "C" {

static_assert(sizeof(gync_prior) == sizeof(GyncPriorDev), "gync_prior layout");

int gync_logprior_batch(const double* X, int64_t N, const gync_prior* prior, double* LP) {
  if (N < 0 || !prior || (N > 0 && (!X || !LP))) return fail(GYNC_ERR_ARG, "gync_logprior_batch: bad arguments");
  if (int rc = ensure_init()) return rc;
  if (N == 0) return GYNC_OK;
  DBuf dX, dP, dL;
  CK(dX.alloc(sizeof(double) * N * GYNC_NX));
  CK(dP.alloc(sizeof(gync_prior)));
  CK(dL.alloc(sizeof(double) * N));
  CK(cudaMemcpyAsync(dX.p, X, sizeof(double) * N * GYNC_NX, cudaMemcpyHostToDevice, g_stream));
  CK(cudaMemcpyAsync(dP.p, prior, sizeof(gync_prior), cudaMemcpyHostToDevice, g_stream));
  gync_logprior_kernel<<<blocks_for(N, 128), 128, 0, g_stream>>>(dX.as<double>(), N, dP.as<GyncPriorDev>(), dL.as<double>());
  g_launches += 1;
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(LP, dL.p, sizeof(double) * N, cudaMemcpyDeviceToHost, g_stream));
  CK(wait_stream(g_stream));
  return GYNC_OK;
}

// ---- K3p host side: one persistent launch per device (gync_mcmc_persist.cuh).  Chains shard over the devices of
// gync_init_multi in contiguous blocks (independent chains, no data-path collective); every device runs
// its block asynchronously and the host only waits once, at the end.
struct McmcStats { unsigned long long solves = 0, steps = 0, cancelled = 0, early = 0, failed = 0, hist[24] = {0}; };
static McmcStats g_mcmc_stats;          // last gync_mcmc_run: likelihood solves started / step attempts (all devices)
static int g_mcmc_K = 0;                // slots per chain round of the last run

namespace {
struct McmcJob {
  int64_t lo = 0, n = 0;
  int K = 1;
  DBuf dS, dLf, dIt, dTl, dAcc, dGen, dPend, dIss, dNc, dLfp, dSt, dDt, dQ, dCtl, dCh, dD, dSig, dPat, dPr, dZ, dU, dOut, dIsc, dNan, dIds;
  // scratch of the initial log-target evaluation
  DBuf dXL, dX, dLpr, dLL, dPs;
  GyncMcmcP P;
};
}  // namespace

static int mcmc_persist_impl(double* state, double* logf, int64_t C, const double* chol, double prop_sd, const double* data,
                             int32_t M, const double* sigma_rho, const int32_t* patient_of_chain, const gync_prior* prior,
                             const gync_solver_opts* opts, int64_t iters, int32_t thin, uint64_t seed, uint64_t iter_offset,
                             const uint64_t* chain_ids, const double* inj_z, const double* inj_u, double* samples_out,
                             int64_t* naccept) {
  const int64_t n_out = iters / thin;
  const int nd = (g_devs.size() > 1 && C >= (int64_t)g_devs.size()) ? (int)g_devs.size() : 1;
  const int home = [&] { for (size_t i = 0; i < g_devs.size(); ++i) if (g_devs[i].device == g_device) return (int)i; return 0; }();
  GyncSolverOpts so = to_opts(opts);
  if (const char* v = getenv("GYNC_MCMC_MAXSTEPS")) so.max_steps = std::max(100, atoi(v));
  bool need_init = false;
  for (int64_t c = 0; c < C; ++c) need_init |= !(logf[c] == logf[c]);
  std::vector<McmcJob> jobs(nd);
  g_mcmc_stats = McmcStats();
  for (int i = 0; i < nd; ++i) {
    McmcJob& J = jobs[i];
    J.lo = C * i / nd;
    J.n = C * (i + 1) / nd - J.lo;
    if (J.n == 0) continue;
    if (nd > 1) if (int rc = use_device(i)) return rc;
    cudaStream_t s = g_stream;
    const int64_t n = J.n, lanes = (int64_t)g_num_sms * GYNC_TM_NSLOT;
    // proposals in flight per chain: 1.1 per lane in total (a lane only idles when the queue is empty; more in flight
    // only adds solves that the next acceptance cancels); look-ahead ring: 4x that
    double over = 1.1;
    if (const char* v = getenv("GYNC_MCMC_OVERSUB")) over = std::max(0.25, atof(v));
    int K = (int)std::max<int64_t>(1, std::min<int64_t>(1024, (int64_t)(over * (double)lanes / (double)n + 0.5)));
    if (const char* v = getenv("GYNC_MCMC_SPEC")) K = std::max(1, std::min(4096, atoi(v)));
    K = (int)std::min<int64_t>(K, std::max<int64_t>(iters, 1));
    J.K = K;
    g_mcmc_K = K;
    const bool fixedK = getenv("GYNC_MCMC_SPEC") != nullptr;             // tests pin the depth: no growth
    const int64_t ring_cap = std::max<int64_t>(8, (512LL << 20) / (48 * n));        // ring storage <= 512 MB per device
    if (K > ring_cap / 2) { K = (int)(ring_cap / 2); J.K = K; g_mcmc_K = K; }
    const int Kmax = fixedK ? K : (int)std::min<int64_t>(std::min<int64_t>(1024, ring_cap / 2), std::max<int64_t>(iters, 1));
    const int R = (int)std::min<int64_t>(std::max<int64_t>(2 * (int64_t)Kmax, std::max<int64_t>(4 * (int64_t)K, 8)), std::max<int64_t>(iters, 1));
    int qshift = 1;
    while ((1LL << qshift) < 4 * (n * K + (int64_t)(over * (double)lanes) + n) + 2) ++qshift;
    CK(J.dS.alloc(sizeof(double) * n * GYNC_NX));
    CK(J.dLf.alloc(sizeof(double) * n));
    CK(J.dIt.alloc(sizeof(long long) * n));
    CK(J.dAcc.alloc(sizeof(long long) * n));
    CK(J.dTl.alloc(sizeof(long long) * n));
    CK(J.dGen.alloc(sizeof(int) * n));
    CK(J.dPend.alloc(sizeof(int) * n));
    CK(J.dIss.alloc(sizeof(int) * n));
    CK(J.dNc.alloc(sizeof(int) * 2 * n));
    CK(J.dLfp.alloc(sizeof(double) * n * 2 * R));
    CK(J.dSt.alloc(sizeof(unsigned long long) * n * 2 * R));
    CK(J.dDt.alloc(sizeof(unsigned long long) * n * 2 * R));
    CK(J.dQ.alloc(sizeof(unsigned long long) << qshift));
    CK(J.dCtl.alloc(sizeof(GyncMcmcCtl)));
    CK(J.dD.alloc(sizeof(double) * 124 * M));
    CK(J.dSig.alloc(sizeof(double) * M));
    CK(J.dPr.alloc(sizeof(gync_prior)));
    CK(J.dIsc.alloc(sizeof(double) * 4 * M));
    CK(J.dNan.alloc(sizeof(int) * M));
    CK(cudaMemcpy2DAsync(J.dS.p, sizeof(double) * n, state + J.lo, sizeof(double) * C, sizeof(double) * n, GYNC_NX,
                         cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(J.dLf.p, logf + J.lo, sizeof(double) * n, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(J.dD.p, data, sizeof(double) * 124 * M, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(J.dSig.p, sigma_rho, sizeof(double) * M, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(J.dPr.p, prior, sizeof(gync_prior), cudaMemcpyHostToDevice, s));
    CK(cudaMemsetAsync(J.dAcc.p, 0, sizeof(long long) * n, s));
    CK(cudaMemsetAsync(J.dIt.p, 0, sizeof(long long) * n, s));
    CK(cudaMemsetAsync(J.dTl.p, 0, sizeof(long long) * n, s));
    CK(cudaMemsetAsync(J.dGen.p, 0, sizeof(int) * n, s));
    CK(cudaMemsetAsync(J.dPend.p, 0, sizeof(int) * n, s));
    CK(cudaMemsetAsync(J.dIss.p, 0, sizeof(int) * n, s));
    CK(cudaMemsetAsync(J.dNc.p, 0, sizeof(int) * 2 * n, s));
    CK(cudaMemsetAsync(J.dSt.p, 0, sizeof(unsigned long long) * n * 2 * R, s));
    CK(cudaMemsetAsync(J.dDt.p, 0, sizeof(unsigned long long) * n * 2 * R, s));
    CK(cudaMemsetAsync(J.dQ.p, 0, sizeof(unsigned long long) << qshift, s));
    CK(cudaMemsetAsync(J.dCtl.p, 0, sizeof(GyncMcmcCtl), s));
    if (chol) {
      CK(J.dCh.alloc(sizeof(double) * GYNC_NX * GYNC_NX));
      CK(cudaMemcpyAsync(J.dCh.p, chol, sizeof(double) * GYNC_NX * GYNC_NX, cudaMemcpyHostToDevice, s));
    }
    if (patient_of_chain) {
      CK(J.dPat.alloc(sizeof(int32_t) * n));
      CK(cudaMemcpyAsync(J.dPat.p, patient_of_chain + J.lo, sizeof(int32_t) * n, cudaMemcpyHostToDevice, s));
    }
    if (chain_ids) {
      CK(J.dIds.alloc(sizeof(uint64_t) * n));
      CK(cudaMemcpyAsync(J.dIds.p, chain_ids + J.lo, sizeof(uint64_t) * n, cudaMemcpyHostToDevice, s));
    }
    if (inj_z) {       // iters x 116 x C and iters x C: chain-major, so a block of chains is contiguous
      CK(J.dZ.alloc(sizeof(double) * iters * GYNC_NX * n));
      CK(J.dU.alloc(sizeof(double) * iters * n));
      CK(cudaMemcpyAsync(J.dZ.p, inj_z + iters * GYNC_NX * J.lo, sizeof(double) * iters * GYNC_NX * n, cudaMemcpyHostToDevice, s));
      CK(cudaMemcpyAsync(J.dU.p, inj_u + iters * J.lo, sizeof(double) * iters * n, cudaMemcpyHostToDevice, s));
    }
    if (samples_out && n_out > 0) CK(J.dOut.alloc(sizeof(double) * n_out * GYNC_NX * n));
    gync_prep_patients_kernel<<<blocks_for(M, 64), 64, 0, s>>>(J.dD.as<double>(), J.dSig.as<double>(), M, J.dIsc.as<double>(), J.dNan.as<int>());
    g_launches += 1;
    GyncMcmcP& P = J.P;
    memset(&P, 0, sizeof(P));
    P.state = J.dS.as<double>(); P.logf = J.dLf.as<double>(); P.itc = J.dIt.as<long long>(); P.naccept = J.dAcc.as<long long>();
    P.tailc = J.dTl.as<long long>(); P.gen = J.dGen.as<int>(); P.pending = J.dPend.as<int>(); P.issued = J.dIss.as<int>(); P.ncomp = J.dNc.as<int>(); P.lfp = J.dLfp.as<double>();
    P.slot_tag = J.dSt.as<unsigned long long>(); P.done_tag = J.dDt.as<unsigned long long>();
    P.idle_limit = 60LL * 1900000000LL;      // ~60 s of SM clock
    P.queue = J.dQ.as<unsigned long long>(); P.qmask = (1ull << qshift) - 1ull; P.qshift = qshift; P.ctl = J.dCtl.as<GyncMcmcCtl>();
    P.C = n; P.W = K; P.Wmax = Kmax; P.inflight_target = fixedK ? (long long)n * K : (long long)std::max<double>(over * (double)lanes, (double)n);
    P.total_iters = (unsigned long long)n * (unsigned long long)iters; P.R = R; P.iters = iters; P.thin = thin;
    P.chol = chol ? J.dCh.as<double>() : nullptr; P.prop_sd = prop_sd; P.pr = J.dPr.as<GyncPriorDev>();
    P.seed = seed; P.iter_offset = iter_offset;
    P.chain_ids = chain_ids ? J.dIds.as<unsigned long long>() : nullptr; P.chain_id_base = (unsigned long long)J.lo;
    P.inj_z = inj_z ? J.dZ.as<double>() : nullptr; P.inj_u = inj_u ? J.dU.as<double>() : nullptr;
    P.patient_of_chain = patient_of_chain ? J.dPat.as<int>() : nullptr;
    P.data = J.dD.as<double>(); P.iscale = J.dIsc.as<double>(); P.allnan = J.dNan.as<int>();
    P.samples_out = (samples_out && n_out > 0) ? J.dOut.as<double>() : nullptr; P.n_out = n_out;
    P.o = so;
    if (need_init) {
      // log-target of the CURRENT states (a fresh variate carries NaN): one batch through the ordinary kernels
      CK(J.dXL.alloc(sizeof(double) * n * GYNC_NX));
      CK(J.dX.alloc(sizeof(double) * n * GYNC_NX));
      CK(J.dLpr.alloc(sizeof(double) * n));
      CK(J.dLL.alloc(sizeof(double) * n));
      CK(J.dPs.alloc(sizeof(int) * n));
      gync_mcmc_propose_kernel<<<blocks_for(n, 128), 128, 0, s>>>(
          P.state, P.itc, n, 1, iters, 1, P.chol, prop_sd, P.pr, seed, iter_offset, nullptr, P.patient_of_chain,
          J.dXL.as<double>(), J.dX.as<double>(), J.dLpr.as<double>(), J.dPs.as<int>());
      g_launches += 1;
      CK(cudaGetLastError());
      if (int rc = launch_loglik(J.dX.as<double>(), n, P.data, P.iscale, P.allnan, M, J.dPs.as<int>(), so, J.dLL.as<double>(),
                                 nullptr, nullptr, nullptr, s)) return rc;
      gync_mcmc_accept_kernel<<<blocks_for(n, 128), 128, 0, s>>>(
          P.state, P.logf, P.itc, n, 1, iters, thin, 1, J.dXL.as<double>(), J.dLpr.as<double>(), J.dLL.as<double>(), seed,
          iter_offset, nullptr, nullptr, P.naccept, nullptr);
      g_launches += 1;
    }
    if (int rc = prep_solver_kernel(gync_mcmc_persist_kernel, GyncWorkTm::kSmemBytes, 128)) return rc;
    gync_mcmc_persist_init_kernel<<<blocks_for(n, 128), 128, 0, s>>>(P);
    const int nb = (int)std::min<int64_t>((n * K + GYNC_TM_NSLOT - 1) / GYNC_TM_NSLOT, (int64_t)g_num_sms);
    gync_mcmc_persist_kernel<<<nb, 128, GyncWorkTm::kSmemBytes, s>>>(P);
    g_launches += 2;
    CK(cudaGetLastError());
    // results come back on the same stream, behind the kernel
    CK(cudaMemcpy2DAsync(state + J.lo, sizeof(double) * C, J.dS.p, sizeof(double) * n, sizeof(double) * n, GYNC_NX,
                         cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(logf + J.lo, J.dLf.p, sizeof(double) * n, cudaMemcpyDeviceToHost, s));
    if (samples_out && n_out > 0)
      CK(cudaMemcpyAsync(samples_out + n_out * GYNC_NX * J.lo, J.dOut.p, sizeof(double) * n_out * GYNC_NX * n, cudaMemcpyDeviceToHost, s));
    if (naccept) CK(cudaMemcpyAsync(naccept + J.lo, J.dAcc.p, sizeof(long long) * n, cudaMemcpyDeviceToHost, s));
  }
  int rc_all = GYNC_OK;
  for (int i = 0; i < nd; ++i) {
    McmcJob& J = jobs[i];
    if (J.n == 0) continue;
    if (nd > 1) if (int rc = use_device(i)) return rc;
    GyncMcmcCtl ctl;
    cudaError_t e = cudaMemcpyAsync(&ctl, J.dCtl.p, sizeof(ctl), cudaMemcpyDeviceToHost, g_stream);
    if (e == cudaSuccess) e = wait_stream(g_stream);
#ifdef GYNC_MCMC_PROFILE
    if (e == cudaSuccess) fprintf(stderr, "[mcmc profile] device slot %d: lane-trips active %llu idle %llu | sleeps %llu | cycles/lane-trip: fetch %.0f step %.0f complete %.0f\n",
            i, ctl.trips_active, ctl.trips_idle, ctl.sleeps,
            (double)ctl.cyc_fetch / (double)(ctl.trips_active + ctl.trips_idle + ctl.sleeps + 1),
            (double)ctl.cyc_step / (double)(ctl.trips_active + ctl.trips_idle + 1),
            (double)ctl.cyc_complete / (double)(ctl.trips_active + ctl.trips_idle + 1));
    if (e == cudaSuccess) {
      fprintf(stderr, "[mcmc profile] proposals by logf + log u - logf' (<0 accept | 0-1 | 1-3 | 3-10 | 10-30 | 30-100 | 100-300 | 300-1e3 | 1e3-1e4 | >1e4 | -Inf):");
      for (int b = 0; b < 11; ++b) fprintf(stderr, " %llu", ctl.deficit[b]);
      fprintf(stderr, "\n");
    }
#endif
    if (e != cudaSuccess) rc_all = fail(GYNC_ERR_CUDA, std::string("gync_mcmc_run: ") + cudaGetErrorString(e));
    else if (ctl.abort) rc_all = fail(GYNC_ERR_CUDA, "gync_mcmc_run: the persistent kernel's watchdog fired (no progress)");
    else { g_mcmc_stats.solves += ctl.solves; g_mcmc_stats.steps += ctl.steps; g_mcmc_stats.cancelled += ctl.cancelled; g_mcmc_stats.early += ctl.early;
           g_mcmc_stats.failed += ctl.failed; for (int b = 0; b < 24; ++b) g_mcmc_stats.hist[b] += ctl.hist[b]; }
    J = McmcJob();                // free on the owning device
  }
  if (nd > 1) if (int rc = use_device(home)) return rc;
  return rc_all;
}

static int mcmc_run_impl(double* state, double* logf, int64_t C, const double* chol, double prop_sd, const double* data,
                  int32_t M, const double* sigma_rho, const int32_t* patient_of_chain, const gync_prior* prior,
                  const gync_solver_opts* opts, int64_t iters, int32_t thin, uint64_t seed, uint64_t iter_offset,
                  const double* inj_z, const double* inj_u, double* samples_out, int64_t* naccept,
                  const gync_amm_tune* amm, const double* inj_umix, const uint64_t* chain_ids) {
  if (C < 0 || M < 1 || iters < 0 || thin < 1 || !prior || (C > 0 && (!state || !logf || !data || !sigma_rho)))
    return fail(GYNC_ERR_ARG, "gync_mcmc_run: bad arguments");
  if (!chol && !(prop_sd > 0.0)) return fail(GYNC_ERR_ARG, "gync_mcmc_run: need chol or prop_sd > 0");
  if ((inj_z == nullptr) != (inj_u == nullptr)) return fail(GYNC_ERR_ARG, "gync_mcmc_run: inj_z and inj_u go together");
  if (patient_of_chain)
    for (int64_t c = 0; c < C; ++c)
      if (patient_of_chain[c] < 0 || patient_of_chain[c] >= M) return fail(GYNC_ERR_ARG, "gync_mcmc_run: patient index out of range");
  if (int rc = ensure_init()) return rc;
  if (C == 0) return GYNC_OK;
  if (!amm && !getenv("GYNC_MCMC_ROUNDS"))      // non-adaptive chains: one persistent launch (gync_mcmc_persist.cuh)
    return mcmc_persist_impl(state, logf, C, chol, prop_sd, data, M, sigma_rho, patient_of_chain, prior, opts, iters, thin, seed,
                             iter_offset, chain_ids, inj_z, inj_u, samples_out, naccept);
  const int64_t n_out = iters / thin;
  // speculation depth: fill the solver kernel's sample slots (112 per SM) about twice per round
  int K = (int)std::max<int64_t>(1, std::min<int64_t>(1024, (2 * (int64_t)g_num_sms * GYNC_TM_NSLOT) / C));
  if (const char* v = getenv("GYNC_MCMC_SPEC")) K = std::max(1, std::min(4096, atoi(v)));
  if (amm && C > 1) K = std::min(K, 32);     // every speculative slot costs a 116 x 116 factorisation (gync_amm.cuh)
  K = (int)std::min<int64_t>(K, std::max<int64_t>(iters, 1));
  const int64_t N = C * K;
  cudaStream_t s = g_stream;
  DBuf dS, dLf, dCh, dD, dSig, dPat, dPr, dZ, dU, dOut, dAcc, dIsc, dNan, dIt, dXL, dX, dLpr, dLL, dPs, dMin, dIds;
  CK(dS.alloc(sizeof(double) * C * GYNC_NX));
  CK(dLf.alloc(sizeof(double) * C));
  CK(dD.alloc(sizeof(double) * 124 * M));
  CK(dSig.alloc(sizeof(double) * M));
  CK(dPr.alloc(sizeof(gync_prior)));
  CK(dIsc.alloc(sizeof(double) * 4 * M));
  CK(dNan.alloc(sizeof(int) * M));
  CK(dAcc.alloc(sizeof(long long) * C));
  CK(dIt.alloc(sizeof(long long) * C));
  CK(dXL.alloc(sizeof(double) * N * GYNC_NX));
  CK(dX.alloc(sizeof(double) * N * GYNC_NX));
  CK(dLpr.alloc(sizeof(double) * N));
  CK(dLL.alloc(sizeof(double) * N));
  CK(dPs.alloc(sizeof(int) * N));
  CK(dMin.alloc(sizeof(long long)));
  CK(cudaMemcpyAsync(dS.p, state, sizeof(double) * C * GYNC_NX, cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(dLf.p, logf, sizeof(double) * C, cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(dD.p, data, sizeof(double) * 124 * M, cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(dSig.p, sigma_rho, sizeof(double) * M, cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(dPr.p, prior, sizeof(gync_prior), cudaMemcpyHostToDevice, s));
  CK(cudaMemsetAsync(dAcc.p, 0, sizeof(long long) * C, s));
  CK(cudaMemsetAsync(dIt.p, 0, sizeof(long long) * C, s));
  if (chol) {
    CK(dCh.alloc(sizeof(double) * GYNC_NX * GYNC_NX));
    CK(cudaMemcpyAsync(dCh.p, chol, sizeof(double) * GYNC_NX * GYNC_NX, cudaMemcpyHostToDevice, s));
  }
  if (patient_of_chain) {
    CK(dPat.alloc(sizeof(int32_t) * C));
    CK(cudaMemcpyAsync(dPat.p, patient_of_chain, sizeof(int32_t) * C, cudaMemcpyHostToDevice, s));
  }
  if (chain_ids) {
    CK(dIds.alloc(sizeof(uint64_t) * C));
    CK(cudaMemcpyAsync(dIds.p, chain_ids, sizeof(uint64_t) * C, cudaMemcpyHostToDevice, s));
  }
  const unsigned long long* d_ids = chain_ids ? dIds.as<unsigned long long>() : nullptr;
  if (inj_z) {
    CK(dZ.alloc(sizeof(double) * iters * GYNC_NX * C));
    CK(dU.alloc(sizeof(double) * iters * C));
    CK(cudaMemcpyAsync(dZ.p, inj_z, sizeof(double) * iters * GYNC_NX * C, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(dU.p, inj_u, sizeof(double) * iters * C, cudaMemcpyHostToDevice, s));
  }
  if (samples_out && n_out > 0) CK(dOut.alloc(sizeof(double) * n_out * GYNC_NX * C));
  gync_prep_patients_kernel<<<blocks_for(M, 64), 64, 0, s>>>(dD.as<double>(), dSig.as<double>(), M, dIsc.as<double>(), dNan.as<int>());
  g_launches += 1;
  // adaptive proposal: tuning state on the device (chain-major), scratch for the speculative replay
  DBuf aM, aMv, aMvv, aLm, aVirt, aXold, aIt0, aNacc0, dUm;
  GyncAmmDev ad;
  memset(&ad, 0, sizeof(ad));
  const size_t d1 = GYNC_NX, d2 = (size_t)GYNC_NX * GYNC_NX;
  if (amm) {
    if (!amm->m || !amm->Mv || !amm->Mvv || !amm->SigmaLm) return fail(GYNC_ERR_ARG, "gync_mcmc_run_amm: NULL tuning arrays");
    CK(aM.alloc(sizeof(long long) * C));
    CK(aMv.alloc(sizeof(double) * C * d1));
    CK(aMvv.alloc(sizeof(double) * C * d2));
    CK(aLm.alloc(sizeof(double) * C * d2));
    CK(aVirt.alloc(sizeof(double) * C * (d1 + d2)));
    CK(aXold.alloc(sizeof(double) * C * d1));
    CK(aIt0.alloc(sizeof(long long) * C));
    CK(aNacc0.alloc(sizeof(long long) * C));
    CK(cudaMemcpyAsync(aM.p, amm->m, sizeof(long long) * C, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(aMv.p, amm->Mv, sizeof(double) * C * d1, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(aMvv.p, amm->Mvv, sizeof(double) * C * d2, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(aLm.p, amm->SigmaLm, sizeof(double) * C * d2, cudaMemcpyHostToDevice, s));
    if (inj_umix) {
      CK(dUm.alloc(sizeof(double) * iters * C));
      CK(cudaMemcpyAsync(dUm.p, inj_umix, sizeof(double) * iters * C, cudaMemcpyHostToDevice, s));
    }
    ad.m = aM.as<long long>(); ad.Mv = aMv.as<double>(); ad.Mvv = aMvv.as<double>(); ad.SigmaLm = aLm.as<double>();
    ad.beta = amm->beta; ad.scale = amm->scale;
    CK(cudaFuncSetAttribute(gync_amm_propose_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(GyncAmmSmem)));
    CK(cudaFuncSetAttribute(gync_amm_commit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(GyncAmmSmem)));
    gync_amm_init_kernel<<<(unsigned)C, GYNC_AMM_THREADS, 0, s>>>(dS.as<double>(), C, chol ? dCh.as<double>() : nullptr, prop_sd, ad);
    g_launches += 1;
    CK(cudaGetLastError());
  }
  const GyncSolverOpts so = to_opts(opts);
  bool need_init = false;
  for (int64_t c = 0; c < C; ++c) need_init |= !(logf[c] == logf[c]);
  // round: propose K slots per chain -> batched likelihoods -> accept scan
  auto round = [&](int mode, int Kr) -> int {
    const int64_t Nr = C * Kr;
    const bool adaptive = amm && mode == 0;
    if (adaptive) {
      gync_amm_propose_kernel<<<(unsigned)C, GYNC_AMM_THREADS, sizeof(GyncAmmSmem), s>>>(
          dS.as<double>(), dIt.as<long long>(), dAcc.as<long long>(), C, Kr, iters, chol ? dCh.as<double>() : nullptr, prop_sd,
          dPr.as<GyncPriorDev>(), seed, iter_offset, inj_z ? dZ.as<double>() : nullptr, inj_umix ? dUm.as<double>() : nullptr,
          patient_of_chain ? dPat.as<int>() : nullptr, ad, aVirt.as<double>(), aXold.as<double>(), aIt0.as<long long>(),
          aNacc0.as<long long>(), dXL.as<double>(), dX.as<double>(), dLpr.as<double>(), dPs.as<int>(), d_ids);
    } else {
      gync_mcmc_propose_kernel<<<blocks_for(Nr, 128), 128, 0, s>>>(
          dS.as<double>(), dIt.as<long long>(), C, Kr, iters, mode, chol ? dCh.as<double>() : nullptr, prop_sd,
          dPr.as<GyncPriorDev>(), seed, iter_offset, inj_z ? dZ.as<double>() : nullptr,
          patient_of_chain ? dPat.as<int>() : nullptr, dXL.as<double>(), dX.as<double>(), dLpr.as<double>(), dPs.as<int>(), d_ids);
    }
    g_launches += 1;
    CK(cudaGetLastError());
    if (int rc = launch_loglik(dX.as<double>(), Nr, dD.as<double>(), dIsc.as<double>(), dNan.as<int>(), M, dPs.as<int>(), so,
                               dLL.as<double>(), nullptr, nullptr, nullptr, s)) return rc;
    long long big = (long long)iters;
    CK(cudaMemcpyAsync(dMin.p, &big, sizeof(long long), cudaMemcpyHostToDevice, s));
    gync_mcmc_accept_kernel<<<blocks_for(C, 128), 128, 0, s>>>(
        dS.as<double>(), dLf.as<double>(), dIt.as<long long>(), C, Kr, iters, thin, mode, dXL.as<double>(), dLpr.as<double>(),
        dLL.as<double>(), seed, iter_offset, inj_u ? dU.as<double>() : nullptr,
        (samples_out && n_out > 0) ? dOut.as<double>() : nullptr, dAcc.as<long long>(), dMin.as<long long>(), d_ids);
    g_launches += 1;
    if (adaptive) {
      gync_amm_commit_kernel<<<(unsigned)C, GYNC_AMM_THREADS, sizeof(GyncAmmSmem), s>>>(
          dS.as<double>(), dIt.as<long long>(), dAcc.as<long long>(), C, ad, aXold.as<double>(), aIt0.as<long long>(),
          aNacc0.as<long long>());
      g_launches += 1;
    }
    CK(cudaGetLastError());
    return GYNC_OK;
  };
  if (need_init) if (int rc = round(1, 1)) return rc;
  long long min_it = 0;
  int Kr = K;
  if (amm && C == 1) Kr = std::min(K, 32);
  while (min_it < iters) {
    const long long before = min_it;
    if (int rc = round(0, Kr)) return rc;
    CK(cudaMemcpyAsync(&min_it, dMin.p, sizeof(long long), cudaMemcpyDeviceToHost, s));
    CK(wait_stream(s));
    if (amm && C == 1) {     // every adaptive slot costs a factorisation: follow the acceptance rate with the depth
      const long long used = min_it - before;
      if (used >= Kr) Kr = std::min(K, 2 * Kr);
      else if (4 * used < Kr) Kr = std::max(8, Kr / 2);
    }
  }
  CK(cudaMemcpyAsync(state, dS.p, sizeof(double) * C * GYNC_NX, cudaMemcpyDeviceToHost, s));
  CK(cudaMemcpyAsync(logf, dLf.p, sizeof(double) * C, cudaMemcpyDeviceToHost, s));
  if (samples_out && n_out > 0)
    CK(cudaMemcpyAsync(samples_out, dOut.p, sizeof(double) * n_out * GYNC_NX * C, cudaMemcpyDeviceToHost, s));
  if (naccept) CK(cudaMemcpyAsync(naccept, dAcc.p, sizeof(long long) * C, cudaMemcpyDeviceToHost, s));
  if (amm) {
    CK(cudaMemcpyAsync(amm->m, aM.p, sizeof(long long) * C, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(amm->Mv, aMv.p, sizeof(double) * C * d1, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(amm->Mvv, aMvv.p, sizeof(double) * C * d2, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(amm->SigmaLm, aLm.p, sizeof(double) * C * d2, cudaMemcpyDeviceToHost, s));
  }
  CK(wait_stream(s));
  return GYNC_OK;
}

int gync_mcmc_run(double* state, double* logf, int64_t C, const double* chol, double prop_sd, const double* data,
                  int32_t M, const double* sigma_rho, const int32_t* patient_of_chain, const gync_prior* prior,
                  const gync_solver_opts* opts, int64_t iters, int32_t thin, uint64_t seed, uint64_t iter_offset,
                  const double* inj_z, const double* inj_u, double* samples_out, int64_t* naccept) {
  return mcmc_run_impl(state, logf, C, chol, prop_sd, data, M, sigma_rho, patient_of_chain, prior, opts, iters, thin, seed,
                       iter_offset, inj_z, inj_u, samples_out, naccept, nullptr, nullptr, nullptr);
}

int gync_mcmc_run_ids(double* state, double* logf, int64_t C, const double* chol, double prop_sd, const double* data,
                      int32_t M, const double* sigma_rho, const int32_t* patient_of_chain, const gync_prior* prior,
                      const gync_solver_opts* opts, int64_t iters, int32_t thin, uint64_t seed, uint64_t iter_offset,
                      const uint64_t* chain_ids, const double* inj_z, const double* inj_u, double* samples_out,
                      int64_t* naccept) {
  return mcmc_run_impl(state, logf, C, chol, prop_sd, data, M, sigma_rho, patient_of_chain, prior, opts, iters, thin, seed,
                       iter_offset, inj_z, inj_u, samples_out, naccept, nullptr, nullptr, chain_ids);
}

int gync_mcmc_run_amm(double* state, double* logf, int64_t C, const double* chol, double prop_sd, const double* data,
                      int32_t M, const double* sigma_rho, const int32_t* patient_of_chain, const gync_prior* prior,
                      const gync_solver_opts* opts, int64_t iters, int32_t thin, uint64_t seed, uint64_t iter_offset,
                      const double* inj_z, const double* inj_u, const double* inj_umix, double* samples_out,
                      int64_t* naccept, const gync_amm_tune* tune) {
  return gync_mcmc_run_amm_ids(state, logf, C, chol, prop_sd, data, M, sigma_rho, patient_of_chain, prior, opts, iters, thin,
                               seed, iter_offset, nullptr, inj_z, inj_u, inj_umix, samples_out, naccept, tune);
}

int gync_mcmc_run_amm_ids(double* state, double* logf, int64_t C, const double* chol, double prop_sd, const double* data,
                          int32_t M, const double* sigma_rho, const int32_t* patient_of_chain, const gync_prior* prior,
                          const gync_solver_opts* opts, int64_t iters, int32_t thin, uint64_t seed, uint64_t iter_offset,
                          const uint64_t* chain_ids, const double* inj_z, const double* inj_u, const double* inj_umix,
                          double* samples_out, int64_t* naccept, const gync_amm_tune* tune) {
  if (!tune) return fail(GYNC_ERR_ARG, "gync_mcmc_run_amm: NULL tune");
  if (C > 65535LL * 32768LL) return fail(GYNC_ERR_ARG, "gync_mcmc_run_amm: too many chains");
  return mcmc_run_impl(state, logf, C, chol, prop_sd, data, M, sigma_rho, patient_of_chain, prior, opts, iters, thin, seed,
                       iter_offset, inj_z, inj_u, samples_out, naccept, tune, inj_umix, chain_ids);
}

// bookkeeping of the last persistent Metropolis run on this thread's library context: likelihood solves started, their
// Rosenbrock step attempts (summed over the devices), slots per chain round
void gync_mcmc_last_stats(uint64_t* solves, uint64_t* steps, int32_t* slots_per_round) {
  if (solves) *solves = g_mcmc_stats.solves;
  if (steps) *steps = g_mcmc_stats.steps;
  if (slots_per_round) *slots_per_round = g_mcmc_K;
}
uint64_t gync_mcmc_last_cancelled(void) { return g_mcmc_stats.cancelled; }
uint64_t gync_mcmc_last_early(void) { return g_mcmc_stats.early; }
// completed solves of the last run by floor(log2(step attempts)) (24 bins), and how many of them failed
void gync_mcmc_last_histogram(uint64_t* hist24, uint64_t* failed) {
  if (hist24) for (int b = 0; b < 24; ++b) hist24[b] = g_mcmc_stats.hist[b];
  if (failed) *failed = g_mcmc_stats.failed;
}

int gync_wc_normalize_dev(double* dLL_inout, const double* d_logprior, const int64_t* d_counts, int64_t K, int32_t M,
                          double* d_weights, double* d_density, void* stream) {
  if (K < 1 || M < 1 || !dLL_inout || !d_logprior || !d_counts || !d_weights || !d_density)
    return fail(GYNC_ERR_ARG, "gync_wc_normalize: bad arguments");
  if (int rc = ensure_init()) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  const int nb = g_num_sms * 4;
  double* scratch = nullptr;      // colsum[M] | scal[2] | partial[nb]
  CK(cudaMallocAsync((void**)&scratch, sizeof(double) * (M + 2 + nb), s));
  double* dCs = scratch;
  double* dSc = scratch + M;
  double* dPart = scratch + M + 2;
  gync_wc_colpass_kernel<<<M + 2, GYNC_WC_THREADS, 0, s>>>(dLL_inout, d_logprior, (const long long*)d_counts, K, M, dLL_inout, dCs, dSc);
  gync_wc_rowpass_kernel<<<nb, 256, 0, s>>>(dLL_inout, d_logprior, (const long long*)d_counts, K, M, dCs, dSc, d_weights, d_density, dPart);
  gync_wc_scale_kernel<<<nb, 256, 0, s>>>(d_density, K, dPart, nb);
  g_launches += 3;
  CK(cudaGetLastError());
  CK(cudaFreeAsync(scratch, s));
  return GYNC_OK;
}

// Duplicate collapse of WeightedChain(samplings, burnin) (weightedchain.jl:31-46) on the device.
// X: K x 116 (the concatenated chains after burn-in).  rows_out: K x 116 buffer whose first *ngroups rows (leading
// dimension K) are the collapsed samples; counts_out: K (first *ngroups valid).
int gync_collapse_duplicates(const double* X, int64_t K, double* rows_out, int64_t* counts_out, int64_t* ngroups) {
  if (K < 0 || !ngroups || (K > 0 && (!X || !rows_out || !counts_out)) || K >= (1LL << 31))
    return fail(GYNC_ERR_ARG, "gync_collapse_duplicates: bad arguments");
  if (int rc = ensure_init()) return rc;
  *ngroups = 0;
  if (K == 0) return GYNC_OK;
  DBuf dX, dOut, dCnt, dFlag, dPos, dN;
  CK(dX.alloc(sizeof(double) * K * GYNC_NX));
  CK(dOut.alloc(sizeof(double) * K * GYNC_NX));
  CK(dCnt.alloc(sizeof(long long) * K));
  CK(dFlag.alloc(sizeof(int) * K));
  CK(dPos.alloc(sizeof(int) * K));
  CK(dN.alloc(sizeof(int)));
  CK(cudaMemcpyAsync(dX.p, X, sizeof(double) * K * GYNC_NX, cudaMemcpyHostToDevice, g_stream));
  gync_dedupe_flag_kernel<<<blocks_for(K, 256), 256, 0, g_stream>>>(dX.as<double>(), K, dFlag.as<int>());
  gync_dedupe_scan_kernel<<<1, 1024, 0, g_stream>>>(dFlag.as<int>(), K, dPos.as<int>(), dN.as<int>());
  gync_dedupe_scatter_kernel<<<blocks_for(K, 256), 256, 0, g_stream>>>(dX.as<double>(), K, dFlag.as<int>(), dPos.as<int>(),
                                                                       dOut.as<double>(), dCnt.as<long long>());
  g_launches += 3;
  CK(cudaGetLastError());
  int n = 0;
  CK(cudaMemcpyAsync(&n, dN.p, sizeof(int), cudaMemcpyDeviceToHost, g_stream));
  CK(wait_stream(g_stream));
  *ngroups = n;
  CK(cudaMemcpy2DAsync(rows_out, sizeof(double) * K, dOut.p, sizeof(double) * K, sizeof(double) * n, GYNC_NX,
                       cudaMemcpyDeviceToHost, g_stream));
  CK(cudaMemcpyAsync(counts_out, dCnt.p, sizeof(long long) * n, cudaMemcpyDeviceToHost, g_stream));
  CK(wait_stream(g_stream));
  return GYNC_OK;
}

// Row-sharded normalisation (one process per GPU): three phases, the caller's collectives in between.
int gync_wc_colstats_dev(const double* dLL, const double* d_logprior, const int64_t* d_counts, int64_t K, int32_t M,
                         double* d_stats /* M + 2 */, void* stream) {
  if (K < 0 || M < 1 || !d_stats || (K > 0 && (!dLL || !d_logprior || !d_counts))) return fail(GYNC_ERR_ARG, "gync_wc_colstats: bad arguments");
  if (int rc = ensure_init()) return rc;
  gync_wc_colstats_kernel<<<M + 2, GYNC_WC_THREADS, 0, (cudaStream_t)stream>>>(dLL, d_logprior, (const long long*)d_counts, K, M, d_stats);
  g_launches += 1;
  CK(cudaGetLastError());
  return GYNC_OK;
}

int gync_wc_expsum_dev(double* dLL_inout, int64_t K, int32_t M, const double* d_colmax, double* d_colsum, void* stream) {
  if (K < 0 || M < 1 || !d_colmax || !d_colsum || (K > 0 && !dLL_inout)) return fail(GYNC_ERR_ARG, "gync_wc_expsum: bad arguments");
  if (int rc = ensure_init()) return rc;
  gync_wc_expsum_kernel<<<M, GYNC_WC_THREADS, 0, (cudaStream_t)stream>>>(dLL_inout, K, M, d_colmax, d_colsum);
  g_launches += 1;
  CK(cudaGetLastError());
  return GYNC_OK;
}

// d_stats: [colsum(M) | lpmax | sumcounts] (global values).  Writes the local weights, the unnormalised local density and
// its local total (d_total[0]); gync_wc_scale_dev divides by the global total.
int gync_wc_rows_dev(double* dL_inout, const double* d_logprior, const int64_t* d_counts, int64_t K, int32_t M,
                     const double* d_stats, double* d_weights, double* d_density, double* d_total, void* stream) {
  if (K < 0 || M < 1 || !d_stats || !d_total || (K > 0 && (!dL_inout || !d_logprior || !d_counts || !d_weights || !d_density)))
    return fail(GYNC_ERR_ARG, "gync_wc_rows: bad arguments");
  if (int rc = ensure_init()) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  const int nb = g_num_sms * 4;
  double* part = nullptr;
  CK(cudaMallocAsync((void**)&part, sizeof(double) * nb, s));
  gync_wc_rowpass_kernel<<<nb, 256, 0, s>>>(dL_inout, d_logprior, (const long long*)d_counts, K, M, d_stats, d_stats + M, d_weights,
                                            d_density, part);
  gync_wc_total_kernel<<<1, 32, 0, s>>>(part, nb, d_total);
  g_launches += 2;
  CK(cudaGetLastError());
  CK(cudaFreeAsync(part, s));
  return GYNC_OK;
}

int gync_wc_scale_dev(double* d_density, int64_t K, const double* d_total, void* stream) {
  if (K < 0 || !d_total || (K > 0 && !d_density)) return fail(GYNC_ERR_ARG, "gync_wc_scale: bad arguments");
  if (int rc = ensure_init()) return rc;
  gync_wc_scale_kernel<<<g_num_sms * 4, 256, 0, (cudaStream_t)stream>>>(d_density, K, d_total, 1);
  g_launches += 1;
  CK(cudaGetLastError());
  return GYNC_OK;
}

int gync_wc_normalize(const double* LL, const double* logprior, const int64_t* counts, int64_t K, int32_t M,
                      double* L, double* weights, double* density) {
  if (K < 1 || M < 1 || !LL || !logprior || !counts || !L || !weights || !density)
    return fail(GYNC_ERR_ARG, "gync_wc_normalize: bad arguments");
  if (int rc = ensure_init()) return rc;
  DBuf dLL, dLp, dCn, dW, dDn;
  CK(dLL.alloc(sizeof(double) * K * M));
  CK(dLp.alloc(sizeof(double) * K));
  CK(dCn.alloc(sizeof(long long) * K));
  CK(dW.alloc(sizeof(double) * K));
  CK(dDn.alloc(sizeof(double) * K));
  CK(cudaMemcpyAsync(dLL.p, LL, sizeof(double) * K * M, cudaMemcpyHostToDevice, g_stream));
  CK(cudaMemcpyAsync(dLp.p, logprior, sizeof(double) * K, cudaMemcpyHostToDevice, g_stream));
  CK(cudaMemcpyAsync(dCn.p, counts, sizeof(long long) * K, cudaMemcpyHostToDevice, g_stream));
  if (int rc = gync_wc_normalize_dev(dLL.as<double>(), dLp.as<double>(), dCn.as<int64_t>(), K, M, dW.as<double>(),
                                     dDn.as<double>(), g_stream)) return rc;
  CK(cudaMemcpyAsync(L, dLL.p, sizeof(double) * K * M, cudaMemcpyDeviceToHost, g_stream));
  CK(cudaMemcpyAsync(weights, dW.p, sizeof(double) * K, cudaMemcpyDeviceToHost, g_stream));
  CK(cudaMemcpyAsync(density, dDn.p, sizeof(double) * K, cudaMemcpyDeviceToHost, g_stream));
  CK(wait_stream(g_stream));
  return GYNC_OK;
}

// ------------------------------------------------------------------------------------------
// emiteration!(w, L) called ONCE PER ITERATION from a host loop — the way the reference's scripts drive it
// (`for i in 1:n emiteration!(w) end`, em.jl:7) — would upload the K x M matrix on every call: 32 MB at the cfg-3 shape,
// milliseconds over PCIe for microseconds of device work.  A handle keeps L (and a staging vector for w) in HBM: per call
// only the K weights travel.  A WeightedChain owns one handle for its likelihood matrix (julia/GynCB200.jl, api.py).
struct gync_em_handle {
  int device = -1;
  int64_t K = 0;
  int32_t M = 0;
  double *dL = nullptr, *dw = nullptr, *dH = nullptr;
  int64_t hist_cap = 0;
};

int gync_em_create(const double* L, int64_t K, int32_t M, gync_em_handle** out) {
  if (!out || K < 1 || M < 1 || !L) return fail(GYNC_ERR_ARG, "gync_em_create: bad arguments");
  if (int rc = ensure_init()) return rc;
  gync_em_handle* h = new gync_em_handle();
  h->device = g_device; h->K = K; h->M = M;
  auto body = [&]() -> int {
    CK(cudaMalloc((void**)&h->dL, sizeof(double) * K * M));
    CK(cudaMalloc((void**)&h->dw, sizeof(double) * K));
    CK(cudaMemcpyAsync(h->dL, L, sizeof(double) * K * M, cudaMemcpyHostToDevice, g_stream));
    CK(wait_stream(g_stream));
    return GYNC_OK;
  };
  if (int rc = body()) {
    const std::string why = g_err;
    gync_em_destroy(h);
    g_err = why;
    return rc;
  }
  *out = h;
  return GYNC_OK;
}

// w: K host doubles, updated in place `niter` times; history (nullable): K x niter, column i = weights after iteration i+1
int gync_em_iterate_handle(gync_em_handle* h, double* w, int32_t niter, double* history) {
  if (!h || !w || niter < 0) return fail(GYNC_ERR_ARG, "gync_em_iterate_handle: bad arguments");
  if (niter == 0) return GYNC_OK;
  if (h->device != g_device) return fail(GYNC_ERR_ARG, "gync_em_iterate_handle: the handle belongs to another device");
  if (history && h->hist_cap < (int64_t)niter) {
    if (h->dH) cudaFree(h->dH);
    h->dH = nullptr; h->hist_cap = 0;
    CK(cudaMalloc((void**)&h->dH, sizeof(double) * h->K * niter));
    h->hist_cap = niter;
  }
  CK(cudaMemcpyAsync(h->dw, w, sizeof(double) * h->K, cudaMemcpyHostToDevice, g_stream));
  if (int rc = gync_em_iterate_dev(h->dw, h->dL, h->K, h->M, niter, history ? h->dH : nullptr, g_stream)) return rc;
  CK(cudaMemcpyAsync(w, h->dw, sizeof(double) * h->K, cudaMemcpyDeviceToHost, g_stream));
  if (history) CK(cudaMemcpyAsync(history, h->dH, sizeof(double) * h->K * niter, cudaMemcpyDeviceToHost, g_stream));
  CK(wait_stream(g_stream));
  return GYNC_OK;
}

void gync_em_destroy(gync_em_handle* h) {
  if (!h) return;
  int cur = 0;
  cudaGetDevice(&cur);
  if (h->device >= 0) cudaSetDevice(h->device);
  if (h->dL) cudaFree(h->dL);
  if (h->dw) cudaFree(h->dw);
  if (h->dH) cudaFree(h->dH);
  cudaSetDevice(cur);
  delete h;
}

// ------------------------------------------------------------------------------------------
// Fused multi-GPU EM: mailboxes in cudaMalloc'd memory exported with cudaIpc, so that each rank's
// kernel can store its partial denominators directly into every peer's HBM over NVLink.
struct gync_em_comm {
  int world = 1, rank = 0, MT = 64;
  double* mail = nullptr;                 // local mailbox [2][world][MT]
  unsigned long long* flag = nullptr;     // local flags   [2][world]
  void* peer_base[8] = {nullptr};         // opened peer allocations (nullptr for self)
  GyncEmPeer peer;
  EmLaunchState st;                       // sequence numbers start at 1: mailbox flags start at 0
  size_t flag_off = 0;
};

void gync_em_comm_destroy(gync_em_comm* c);
int gync_em_comm_create(int32_t world, int32_t rank, void* handle_out /* 64 bytes */, gync_em_comm** out) {
  if (world < 1 || world > 8 || rank < 0 || rank >= world || !handle_out || !out)
    return fail(GYNC_ERR_ARG, "gync_em_comm_create: bad arguments (world <= 8)");
  if (int rc = ensure_init()) return rc;
  gync_em_comm* c = new gync_em_comm();
  c->world = world; c->rank = rank;
  const size_t mail_bytes = sizeof(double) * 2 * world * c->MT * 2;      // [2 slots][world][MT] values x 2 self-validating words
  c->flag_off = (mail_bytes + 255) / 256 * 256;
  const size_t total = c->flag_off + sizeof(unsigned long long) * 2 * world;
  // every failure path below releases what was allocated so far (gync_em_comm_destroy copes with a half-built object)
  auto body = [&]() -> int {
    void* base = nullptr;
    CK(cudaMalloc(&base, total));
    c->mail = (double*)base;
    c->flag = (unsigned long long*)((char*)base + c->flag_off);
    CK(cudaMemset(base, 0, total));
    if (int rc = c->st.alloc()) return rc;
    cudaIpcMemHandle_t h;
    CK(cudaIpcGetMemHandle(&h, base));
    static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t size");
    memcpy(handle_out, &h, 64);
    c->peer.world = world; c->peer.rank = rank;
    c->peer.mail[rank] = c->mail; c->peer.flag[rank] = c->flag;
    CK(cudaDeviceSynchronize());
    return GYNC_OK;
  };
  if (int rc = body()) {
    const std::string why = g_err;
    gync_em_comm_destroy(c);
    g_err = why;
    return rc;
  }
  *out = c;
  return GYNC_OK;
}

int gync_em_comm_connect(gync_em_comm* c, const void* all_handles /* world x 64 bytes, rank order */) {
  if (!c || !all_handles) return fail(GYNC_ERR_ARG, "gync_em_comm_connect: NULL");
  for (int p = 0; p < c->world; ++p) {
    if (p == c->rank) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, (const char*)all_handles + 64 * p, 64);
    void* base = nullptr;
    CK(cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess));
    c->peer_base[p] = base;
    c->peer.mail[p] = (double*)base;
    c->peer.flag[p] = (unsigned long long*)((char*)base + c->flag_off);
  }
  return GYNC_OK;
}

void gync_em_comm_destroy(gync_em_comm* c) {
  if (!c) return;
  for (int p = 0; p < c->world; ++p) if (c->peer_base[p]) cudaIpcCloseMemHandle(c->peer_base[p]);
  if (c->mail) cudaFree(c->mail);
  c->st.release();
  delete c;
}

}  // extern "C"

extern "C" {

int gync_em_iterate_fused_dev(gync_em_comm* c, double* dw, const double* dL, int64_t K, int32_t M, int32_t niter,
                              void* stream) {
  if (!c || K < 0 || M < 1 || niter < 0 || (K > 0 && (!dw || !dL))) return fail(GYNC_ERR_ARG, "gync_em_iterate_fused: bad arguments");
  cudaStream_t s = (cudaStream_t)stream;
#define CALL(MH, F) em_launch<MH, F>(c->st, c->peer, dw, dL, K, M, niter, nullptr, s)
  EM_DISPATCH(M, CALL);
#undef CALL
}

#ifdef GYNC_EM_PHASES
// profile builds only: read and clear the phase counters of the current device
int gync_em_phase_read(uint64_t* out8) {
  CK(cudaMemcpyFromSymbol(out8, gync_em_phase, sizeof(unsigned long long) * 8));
  unsigned long long z[8] = {0};
  CK(cudaMemcpyToSymbol(gync_em_phase, z, sizeof(z)));
  return GYNC_OK;
}
#endif

int gync_em_comm_error(gync_em_comm* c) {
  if (!c) return -1;
  int e = 0;
  if (cudaMemcpy(&e, c->st.err, sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) return -2;
  return e;
}


// Single-process multi-GPU EM (gync_init_multi): rows of (w, L) are sharded over the devices and every device runs
// the fused step+exchange kernel; the mailboxes are wired with plain peer pointers (same process, peer access on).
int gync_em_iterate_multi(double* w, const double* L, int64_t K, int32_t M, int32_t niter) {
  const int nd = (int)g_devs.size();
  if (nd < 2) return fail(GYNC_ERR_ARG, "gync_em_iterate_multi: call gync_init_multi with >= 2 devices first");
  if (K < nd || M < 1 || niter < 0 || !w || !L) return fail(GYNC_ERR_ARG, "gync_em_iterate_multi: bad arguments");
  struct Part { int64_t lo = 0, n = 0; DBuf dw, dL; gync_em_comm* comm = nullptr; };
  std::vector<Part> parts(nd);
  char handle[64];
  int rc = GYNC_OK;
  for (int i = 0; i < nd && !rc; ++i) {
    Part& p = parts[i];
    p.lo = K * i / nd;
    p.n = K * (i + 1) / nd - p.lo;
    if ((rc = use_device(i))) break;
    if ((rc = gync_em_comm_create(nd, i, handle, &p.comm))) break;
    if (p.dw.alloc(sizeof(double) * p.n) != cudaSuccess || p.dL.alloc(sizeof(double) * p.n * M) != cudaSuccess) { rc = fail(GYNC_ERR_NOMEM, "gync_em_iterate_multi: out of device memory"); break; }
    cudaMemcpyAsync(p.dw.p, w + p.lo, sizeof(double) * p.n, cudaMemcpyHostToDevice, g_stream);
    cudaMemcpy2DAsync(p.dL.p, sizeof(double) * p.n, L + p.lo, sizeof(double) * K, sizeof(double) * p.n, M, cudaMemcpyHostToDevice, g_stream);
  }
  if (!rc) {
    for (int i = 0; i < nd; ++i)                   // wire the mailboxes: direct peer pointers
      for (int j = 0; j < nd; ++j) {
        parts[i].comm->peer.mail[j] = parts[j].comm->mail;
        parts[i].comm->peer.flag[j] = parts[j].comm->flag;
      }
    int launched = 0;
    for (int i = 0; i < nd && !rc; ++i) {          // all launches are asynchronous: the kernels meet on the GPUs
      if ((rc = use_device(i))) break;
      rc = gync_em_iterate_fused_dev(parts[i].comm, parts[i].dw.as<double>(), parts[i].dL.as<double>(), parts[i].n, M, niter, g_stream);
      if (!rc) launched = i + 1;
    }
    if (rc && launched > 0) {
      // a later device failed to launch: tell the kernels already running (they are waiting for its mailbox) to stop
      // now instead of running into their spin time-out — their error word is polled inside every spin
      const std::string why = g_err;
      for (int i = 0; i < launched; ++i) {
        if (use_device(i)) continue;
        cudaStream_t side = nullptr;
        if (cudaStreamCreateWithFlags(&side, cudaStreamNonBlocking) == cudaSuccess) {
          const int one = 1;
          cudaMemcpyAsync(parts[i].comm->st.err, &one, sizeof(int), cudaMemcpyHostToDevice, side);
          cudaStreamSynchronize(side);
          cudaStreamDestroy(side);
        }
      }
      g_err = why;
    }
    for (int i = 0; i < nd && !rc; ++i) {
      if ((rc = use_device(i))) break;
      cudaMemcpyAsync(w + parts[i].lo, parts[i].dw.p, sizeof(double) * parts[i].n, cudaMemcpyDeviceToHost, g_stream);
      if (cudaStreamSynchronize(g_stream) != cudaSuccess) rc = fail(GYNC_ERR_CUDA, "gync_em_iterate_multi: kernel failed");
      else if (gync_em_comm_error(parts[i].comm)) rc = fail(GYNC_ERR_CUDA, "gync_em_iterate_multi: peer wait timed out");
    }
  }
  for (int i = 0; i < nd; ++i) {
    use_device(i);
    cudaStreamSynchronize(g_stream);
    if (parts[i].comm) gync_em_comm_destroy(parts[i].comm);
    parts[i] = Part();
  }
  use_device(0);
  return rc;
}

}  // extern "C"
