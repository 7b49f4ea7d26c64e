// K3p: many-chain Metropolis as ONE persistent kernel — propose, solve, score, accept and re-propose on the device, with
// no kernel boundary, host copy or synchronisation between rounds.
//
// Replaces the round loop of Mamba.sample! behind sample!(s::Sampling, iters) (reference src/gync/sampling.jl:62-79,
// target src/gync/model.jl:97-107) for non-adaptive chains.  Same speculative scheme as gync_mcmc.cuh (the proposals of
// iterations i .. i+K-1 of a chain are evaluated at once from its current state; an ordered scan keeps everything up to
// and including the first acceptance — exactly the sequential chain, for every K), but the rounds of different chains
// are no longer in lock step:
//
//   * the solver lanes (the tensor-memory layout of K1: 128 lanes per SM, one sample each, persistent) pull WORK ITEMS
//     (chain c, slot k) from a ring queue in global memory;
//   * the lane that picks an item up builds the proposal itself: x' = x_c + SigmaL z(seed, chain, iteration) with the
//     addressable Philox draws, exp, log-prior, parameter packing — then integrates it like any other sample and
//     accumulates the patient's residuals in a register;
//   * the lane that finishes the LAST outstanding slot of a chain's round runs that chain's accept scan, writes the
//     thinned output rows, and immediately enqueues the chain's next round.  Nobody waits for the slowest chain; a lane
//     is idle only when the queue is empty.
//
// Round 1 ran propose / solve / accept as separate launches per round with a device->host copy and a host
// synchronisation after each (gync_abi_more.inl), every round ending with the slowest of its C*K solves.
#pragma once
#include "gync_mcmc.cuh"

struct GyncMcmcCtl {                 // one per launch, zero-initialised
  unsigned long long head;           // consumer tickets handed out
  unsigned long long tail;           // queue entries reserved by producers
  int chains_done;                   // chains that have completed all their iterations
  int pad;
  unsigned long long solves;         // likelihood solves started (bookkeeping: speculative efficiency)
  unsigned long long steps;          // Rosenbrock step attempts of those solves
};

struct GyncMcmcP {
  // chains (state is C x 116 log space, column-major like everywhere else)
  double* state;
  double* logf;
  long long* itc;                    // next iteration of each chain
  long long* naccept;
  int* round_left;                   // outstanding slots of the chain's current round
  int* round_k;                      // slots of the current round
  double* lfp;                       // C x K: log-target of the proposals of the current round
  unsigned long long* queue;         // ring of (generation + 1) << 40 | item,  item = c * K + k
  unsigned long long qmask;          // capacity - 1 (power of two, > 2 C K)
  int qshift;                        // log2(capacity)
  GyncMcmcCtl* ctl;
  int64_t C;
  int K;
  int64_t iters;
  int thin;
  const double* chol;                // 116 x 116 lower factor, column-major; NULL: isotropic prop_sd
  double prop_sd;
  const GyncPriorDev* pr;
  uint64_t seed, iter_offset;
  const unsigned long long* chain_ids;   // Philox stream of each chain; NULL: chain c draws from stream chain_id_base + c
  unsigned long long chain_id_base;
  const double* inj_z;               // test hook: iters x 116 x C normals, iters x C uniforms
  const double* inj_u;
  const int* patient_of_chain;       // NULL: patient 0
  const double* data;                // 31 x 4 x M
  const double* iscale;              // 4 x M
  const int* allnan;                 // M
  double* samples_out;               // n_out x 116 x C or NULL
  int64_t n_out;
  GyncSolverOpts o;
};

__device__ __forceinline__ unsigned long long gync_ld_vol_u64(const unsigned long long* p) { return *(const volatile unsigned long long*)p; }

__device__ __forceinline__ uint64_t gync_mcmc_chain_id(const GyncMcmcP& P, int64_t c) {
  return P.chain_ids ? (uint64_t)P.chain_ids[c] : (uint64_t)(P.chain_id_base + (unsigned long long)c);
}

// proposal of iteration `it` of chain c from its current state (log space) -> prop[116]
__device__ __noinline__ void gync_mcmc_proposal(const GyncMcmcP& P, int64_t c, long long it, double* prop) {
  double z[GYNC_NX];
  if (P.inj_z) {
    for (int d = 0; d < GYNC_NX; ++d) z[d] = P.inj_z[it + P.iters * (d + (int64_t)GYNC_NX * c)];
  } else {
    const uint64_t git = P.iter_offset + (uint64_t)it, cid = gync_mcmc_chain_id(P, c);
    for (int b = 0; b < GYNC_NX / 2; ++b) gync_draw_n2(P.seed, cid, git, (uint32_t)b, z[2 * b], z[2 * b + 1]);
  }
  if (P.chol) {
    for (int d = 0; d < GYNC_NX; ++d) {
      double acc = __ldcg(P.state + c + P.C * d);
      for (int e = 0; e <= d; ++e) acc = fma(__ldg(P.chol + d + GYNC_NX * e), z[e], acc);
      prop[d] = acc;
    }
  } else {
    for (int d = 0; d < GYNC_NX; ++d) prop[d] = fma(P.prop_sd, z[d], __ldcg(P.state + c + P.C * d));
  }
}

// enqueue the next round of chain c (iterations it .. it+kk-1); caller has published the chain state
__device__ __forceinline__ void gync_mcmc_push_round(const GyncMcmcP& P, int64_t c, long long it) {
  const int kk = (int)min((long long)P.K, (long long)P.iters - it);
  P.round_k[c] = kk;
  P.round_left[c] = kk;
  __threadfence();
  const unsigned long long t0 = atomicAdd(&P.ctl->tail, (unsigned long long)kk);
  for (int j = 0; j < kk; ++j) {
    const unsigned long long t = t0 + j;
    *(volatile unsigned long long*)(P.queue + (t & P.qmask)) = (((t >> P.qshift) + 1ull) << 40) | (unsigned long long)(c * P.K + j);
  }
}

// accept scan of chain c's finished round (the sequential Metropolis rule over the slots, in iteration order)
__device__ __noinline__ void gync_mcmc_scan(const GyncMcmcP& P, int64_t c) {
  // chain words were last written by another lane, possibly on another SM: read them from L2, never from this SM's L1
  long long it = __ldcg(P.itc + c);
  double lf = __ldcg(P.logf + c);
  long long nacc = __ldcg(P.naccept + c);
  const int kk = __ldcg(P.round_k + c);
  const uint64_t cid = gync_mcmc_chain_id(P, c);
  for (int k = 0; k < kk; ++k, ++it) {
    const double lfp = __ldcg(P.lfp + c * P.K + k);
    double u, u1;
    if (P.inj_u) u = P.inj_u[it + P.iters * c];
    else gync_draw_u2(P.seed, cid, P.iter_offset + (uint64_t)it, (uint32_t)(GYNC_NX / 2), u, u1);
    const bool acc = u < exp(lfp - lf);              // Mamba: rand() < exp(logf(x') - logf(x))
    if (acc) {
      double prop[GYNC_NX];
      gync_mcmc_proposal(P, c, it, prop);            // the same draws the slot was built from
      for (int d = 0; d < GYNC_NX; ++d) P.state[c + P.C * d] = prop[d];
      lf = lfp;
      ++nacc;
    }
    if ((it + 1) % P.thin == 0) {
      const int64_t row = (it + 1) / P.thin - 1;
      if (P.samples_out && row < P.n_out)
        for (int d = 0; d < GYNC_NX; ++d)
          P.samples_out[row + P.n_out * (d + (int64_t)GYNC_NX * c)] = exp(__ldcg(P.state + c + P.C * d));   // sampling.jl:74
    }
    if (acc) { ++it; break; }                        // later slots were proposed from the old state
  }
  P.itc[c] = it;
  P.logf[c] = lf;
  P.naccept[c] = nacc;
  if (it >= P.iters) {
    __threadfence();
    atomicAdd(&P.ctl->chains_done, 1);
  } else {
    gync_mcmc_push_round(P, c, it);
  }
}

// a slot's log-target is known: record it; the last finisher of the round scans and re-proposes
__device__ __forceinline__ void gync_mcmc_complete(const GyncMcmcP& P, int64_t c, int k, double lfp) {
  __stcg(P.lfp + c * P.K + k, lfp);
  __threadfence();
  if (atomicSub(P.round_left + c, 1) == 1) {
    __threadfence();
    gync_mcmc_scan(P, c);
  }
}

// first round of every chain (after the initial log-targets are known)
__global__ void gync_mcmc_persist_init_kernel(GyncMcmcP P) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= P.C) return;
  if (P.itc[c] >= P.iters) atomicAdd(&P.ctl->chains_done, 1);
  else gync_mcmc_push_round(P, c, P.itc[c]);
}

struct GyncLlhSinkReg {              // one patient per sample, residuals accumulated in a register
  const double* __restrict__ data;   // 31 x 4 of this sample's patient
  const double* __restrict__ iscale; // 4
  double acc;
  __device__ __forceinline__ void operator()(int, int day, const double* y) {
    const double ym[GYNC_NMEAS] = {y[1], y[6], y[23], y[24]};   // measuredinds (model.jl:1)
    double a = 0.0;                                             // same summation order as GyncLlhSink
#pragma unroll
    for (int h = 0; h < GYNC_NMEAS; ++h) {
      const double d = __ldg(data + day + GYNC_NDAYS * h);
      if (d == d) {
        const double r = (d - ym[h]) * __ldg(iscale + h);
        a = fma(r, r, a);
      }
    }
    acc += a;
  }
};

__global__ void __launch_bounds__(128, 1)
gync_mcmc_persist_kernel(const GyncMcmcP P) {
  extern __shared__ double gync_sm[];
  __shared__ uint32_t s_tmem;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;"
                 :: "r"((uint32_t)__cvta_generic_to_shared(&s_tmem)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = s_tmem;
  GyncWorkTm W;
  const int slot = warp * 32 + lane;
  W.q.p = gync_sm + slot;
  W.A.p = gync_sm + (size_t)GYNC_NQ * GYNC_TM_NSLOT + slot;
  W.tm = tmem_base + ((uint32_t)(warp * 32) << 16);
  W.A.tm = W.tm;
#pragma unroll
  for (int i = 0; i < GYNC_NY; ++i) { W.y[i] = 1.0; W.yn[i] = 1.0; W.e[i] = 0.0; }

  const GyncSolverOpts o = P.o;
  const double big = 1.0e300;
  bool active = false, done = false;
  long long ticket = -1;
  int64_t c = 0;
  int kslot = 0;
  double T = 0.0, t = 0.0, tn = 0.0, lpr = 0.0;
  int i0 = 0, i1 = 0, budget = 0, st = GYNC_ST_OK, mpat = 0;
  GyncCtrl ctl;
  gync_ctrl_init(ctl, 1e-3);
  GyncLlhSinkReg sink = {P.data, P.iscale, 0.0};
  for (;;) {
    if (!active && !done) {
      if (ticket < 0) ticket = (long long)atomicAdd(&P.ctl->head, 1ull);
      const unsigned long long ent = gync_ld_vol_u64(P.queue + ((unsigned long long)ticket & P.qmask));
      if ((ent >> 40) == (((unsigned long long)ticket >> P.qshift) + 1ull)) {
        __threadfence();
        const int64_t item = (int64_t)(ent & ((1ull << 40) - 1ull));
        ticket = -1;
        c = item / P.K;
        kslot = (int)(item % P.K);
        const long long it = __ldcg((const long long*)P.itc + c) + kslot;
        double x[GYNC_NX];
        gync_mcmc_proposal(P, c, it, x);
        double sumx = 0.0;
        for (int d = 0; d < GYNC_NX; ++d) { sumx += x[d]; x[d] = exp(x[d]); }
        const double lp = gync_logprior(x, P.pr);          // logf(x) = logpost(c, exp(x)) + sum(x)   (model.jl:102)
        mpat = P.patient_of_chain ? P.patient_of_chain[c] : 0;
        if (!(lp > -INFINITY) || P.allnan[mpat]) {
          // outside the prior the likelihood is skipped (model.jl:121); an all-missing table scores -Inf (model.jl:161)
          gync_mcmc_complete(P, c, kslot, -INFINITY);
        } else {
          lpr = lp + sumx;
          gync_load_sample(W, [&](int k) { return x[k]; }, c_refallparms, c_sampledinds, &T);
          sink.data = P.data + (size_t)GYNC_NDAYS * GYNC_NMEAS * mpat;
          sink.iscale = P.iscale + GYNC_NMEAS * mpat;
          sink.acc = 0.0;
          i0 = i1 = 0;
          budget = o.max_steps;
          t = fmin(1.0, 1.0 + T);
          tn = t;
          st = (gync_all_finite(W.y, GYNC_NY) && gync_all_finite(W.q, GYNC_NQ) && (T * 0.0 == 0.0)) ? GYNC_ST_OK
                                                                                                   : GYNC_ST_NONFINITE;
          gync_ctrl_init(ctl, (st == GYNC_ST_OK) ? ((o.h0 > 0.0) ? o.h0 : gync_initial_step(W, o.rtol, o.atol)) : 1.0);
          active = true;
          atomicAdd(&P.ctl->solves, 1ull);
          if (st == GYNC_ST_OK) {
            for (;;) {
              const double tA = (i0 < GYNC_NDAYS) ? (double)(i0 + 1) : big;
              const double tB = (i1 < GYNC_NDAYS) ? (double)(i1 + 1) + T : big;
              tn = fmin(tA, tB);
              if (tn > t) break;
              if (tA == tn) { sink(0, i0, W.y); ++i0; }
              if (tB == tn) { sink(1, i1, W.y); ++i1; }
            }
          }
        }
      } else if (*(volatile int*)&P.ctl->chains_done >= (int)P.C) {
        done = true;
      }
    }
    __syncwarp();
    if (__all_sync(0xffffffffu, done)) break;
    // ---- one step attempt, executed by ALL lanes of the warp (TMEM ops are warp-collective)
    const bool stepping = active && st == GYNC_ST_OK && tn > t;
    double h = ctl.h;
    if (o.hmax > 0.0) h = fmin(h, o.hmax);
    const double rem = tn - t;
    const bool last = stepping && (h * 1.0001 >= rem);
    if (last) h = rem;
    if (!stepping || !(h > 0.0)) h = 1.0;                  // idle lanes: any finite positive step
    const double hprop = ctl.h;
    const double err = gync_ros_step_tm(W, h, o.rtol, o.atol);
    __syncwarp();
    if (stepping) {
      if (budget-- <= 0) {
        st = GYNC_ST_MAXSTEPS;
      } else if (!(h > 1e-14 * fmax(1.0, fabs(t)))) {
        st = GYNC_ST_HMIN;
      } else {
        const bool fin = gync_all_finite(W.yn, GYNC_NY);
        if (gync_ctrl_update(ctl, h, err, fin)) {
          t = last ? tn : t + h;
#pragma unroll
          for (int i = 0; i < GYNC_NY; ++i) W.y[i] = W.yn[i];
          if (last && h < hprop) ctl.h = fmax(ctl.h, hprop);
        }
      }
    }
    if (active) {
      bool finished = (st != GYNC_ST_OK);
      if (!finished && t >= tn) {
        for (;;) {
          const double tA = (i0 < GYNC_NDAYS) ? (double)(i0 + 1) : big;
          const double tB = (i1 < GYNC_NDAYS) ? (double)(i1 + 1) + T : big;
          tn = fmin(tA, tB);
          if (tn > t) break;
          if (tA == tn) { sink(0, i0, W.y); ++i0; }
          if (tB == tn) { sink(1, i1, W.y); ++i1; }
        }
        if (i0 >= GYNC_NDAYS && i1 >= GYNC_NDAYS) finished = true;
      }
      if (finished) {
        atomicAdd(&P.ctl->steps, (unsigned long long)(ctl.nacc + ctl.nrej));
        // failed solve -> NaN trajectory -> -Inf (gyncycle.jl:18-19, model.jl:138-141)
        gync_mcmc_complete(P, c, kslot, (st == GYNC_ST_OK) ? __dadd_rn(lpr, __dmul_rn(-0.5, sink.acc)) : -INFINITY);
        active = false;
#pragma unroll
        for (int i = 0; i < GYNC_NY; ++i) if (!(W.y[i] * 0.0 == 0.0)) W.y[i] = 1.0;
      }
    }
  }
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" :: "r"(tmem_base) : "memory");
  }
}
