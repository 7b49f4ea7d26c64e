// K3p: many-chain Metropolis as ONE persistent kernel — propose, solve, score, accept and re-propose on the device, with
// no kernel boundary, host copy or synchronisation anywhere inside the run.
//
// Replaces the loop of Mamba.sample! behind sample!(s::Sampling, iters) (reference src/gync/sampling.jl:62-79, target
// src/gync/model.jl:97-107) for non-adaptive chains.  The speculation is the one of gync_mcmc.cuh — the proposals of
// the next iterations of a chain are evaluated from its current state, and an ORDERED scan keeps everything up to and
// including the first acceptance, which is exactly the sequential chain for every window size — but nothing here is
// in lock step:
//
//   * every chain keeps W proposals IN FLIGHT (queued or being integrated), taken from a look-ahead ring of R >= W
//     undecided iterations.  The solver lanes (the tensor-memory layout of K1: 128 lanes per SM, one sample each,
//     persistent) pull work items (chain, iteration) from a ring queue in global memory;
//   * the lane that picks an item up builds the proposal itself: x' = x_c + SigmaL z(seed, stream, iteration) with the
//     addressable Philox draws, exp, log-prior, parameter packing — then integrates it like any other sample and
//     accumulates the patient's residuals in a register;
//   * a lane that finishes an item records the proposal's log-target and ADVANCES its chain: while the oldest
//     undecided iteration has its result, apply the Metropolis rule to it (in iteration order, so the chain is the
//     sequential one); then top the number of proposals in flight up to W again from the far end of the ring.  Issue is
//     decoupled from retirement: a slow solve at the head of the ring delays the decisions behind it, not the work —
//     results that arrive out of order wait in their ring slots and cost nothing.  An acceptance moves the chain and
//     bumps its generation, which cancels every younger item of the old generation: not-yet-started ones are dropped
//     at pick-up, running ones give up at their next step; W proposals from the new state are issued at once.
//     Advancing is serialised per chain by a combining counter (the lane that finds it busy leaves its result for the
//     current advancer), never by a lock that spins.
//
//   * W is not fixed: a chain's share of the ~1.1 proposals in flight per lane is proportional to its REMAINING iterations.
//     Chains differ in cost per iteration by tens of percent; with equal shares the cheap ones finish early and leave their
//     lanes idle, this way a chain that falls behind gets more lanes and all finish together.
//
//   * a proposal is integrated only as far as its fate is open.  The log-likelihood is -1/2 times a sum of squares that
//     only grows as the measurement times go by, and acceptance needs  logprior' + sum(x') - 1/2 S > logf + log u  with
//     everything but S known when the item is picked up (u is addressable like every draw): once S has passed
//     2 (logprior' + sum(x') - logf - log u) the proposal IS rejected, whatever the rest of the trajectory does — the lane
//     stops there.  The decisions, and therefore the chains, are exactly those of the full solves (most proposals of a
//     random-walk chain miss acceptance by hundreds of log-units: they stop about half-way).
//
// No lane ever waits for another chain, for the slowest solve of a batch, or for the host; a lane is idle only when
// the queue is empty.  Round 1 ran propose / solve / accept as separate launches per round with a device->host copy
// and a host synchronisation after each (still in gync_abi_more.inl, GYNC_MCMC_ROUNDS=1, and used by the adaptive
// sampler), every round ending with the slowest of its C*K solves.
#pragma once
#include "gync_mcmc.cuh"

struct GyncMcmcCtl {                 // one per launch, zero-initialised
  unsigned long long head;           // consumer tickets handed out
  unsigned long long tail;           // queue entries reserved by producers
  unsigned long long decided;        // iterations decided so far, all chains
  int chains_done;                   // chains that have decided all their iterations
  int abort;                         // watchdog: a lane saw no progress for too long (never expected)
  unsigned long long solves;         // likelihood solves started (bookkeeping: cost of the speculation)
  unsigned long long steps;          // Rosenbrock step attempts spent, cancelled solves included
  unsigned long long cancelled;      // solves given up or dropped because their chain had moved on
  unsigned long long early;          // solves stopped early: their residual sum had already passed what acceptance allows
  unsigned long long failed;         // completed solves with a non-zero status (their proposal scores -Inf)
  unsigned long long hist[24];       // completed solves by floor(log2(step attempts))
  // -DGYNC_MCMC_PROFILE builds only: where the lanes' time goes (summed over lanes)
  unsigned long long trips_active, trips_idle, cyc_fetch, cyc_complete, cyc_step, sleeps;
  unsigned long long deficit[12];    // completed proposals by how far they miss acceptance: logf + log u - logf' in
                                     // (<0 = accept, 0-1, 1-3, 3-10, 10-30, 30-100, 100-300, 300-1e3, 1e3-1e4, >1e4, -Inf, unused)
};
#ifdef GYNC_MCMC_PROFILE
#  define GYNC_MP(...) __VA_ARGS__
#else
#  define GYNC_MP(...)
#endif

#define GYNC_MCMC_GEN_MASK 0xFFFFFFull          /* generation counter, 24 bits inside a slot tag */

struct GyncMcmcP {
  // chains (state is C x 116 log space, column-major like everywhere else)
  double* state;
  double* logf;
  long long* itc;                    // head: next iteration to decide = iterations decided so far
  long long* tailc;                  // next iteration to enqueue (head <= tail <= head + R)
  int* gen;                          // bumped by every acceptance
  int* pending;                      // combining counter of the advance
  int* issued;                       // items enqueued under the current generation (written by the advancer only)
  int* ncomp;                        // C x 2, by generation parity: items of that generation that have reported back
                                     // (result or cancellation); issued - ncomp = proposals in flight
  long long* naccept;
  // ring slots: 2 R per chain, slot = (generation & 1) * R + iteration mod R.  The items an acceptance cancels cover
  // the same iterations as the items issued after it; the generation parity keeps a cancelled item that is just
  // finishing from writing into its successor's slot.
  double* lfp;                       // C x 2R: log-target of the proposal
  unsigned long long* slot_tag;      // C x 2R: identity (generation << 40 | iteration + 1) of the item that owns the slot
  unsigned long long* done_tag;      // C x 2R: same identity once the slot's result is valid
  unsigned long long* queue;         // ring of (ticket generation + 1) << 40 | item,  item = c * 2R + slot
  unsigned long long qmask;          // capacity - 1 (power of two, > 4 C W: at most two generations of a chain are queued)
  int qshift;                        // log2(capacity)
  GyncMcmcCtl* ctl;
  int64_t C;
  int W;                             // proposals in flight per chain when all chains have equally far to go
  int Wmax;                          // ... and at most
  long long inflight_target;         // proposals in flight over all chains (a little more than there are lanes)
  unsigned long long total_iters;    // C * iters minus what was already decided at launch
  int R;                             // look-ahead ring (undecided iterations per chain), >= 2 Wmax
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
  long long idle_limit;              // watchdog: clock64 ticks a lane may see the queue's tail unchanged before it raises ctl->abort
  GyncSolverOpts o;
};

__device__ __forceinline__ unsigned long long gync_ld_vol_u64(const unsigned long long* p) { return *(const volatile unsigned long long*)p; }
__device__ __forceinline__ unsigned long long gync_mcmc_tag(int gen, long long it) {
  return (((unsigned long long)gen & GYNC_MCMC_GEN_MASK) << 40) | (unsigned long long)(it + 1);
}
__device__ __forceinline__ int64_t gync_mcmc_slot(const GyncMcmcP& P, int64_t c, int gen, long long it) {
  return c * (2 * (int64_t)P.R) + (int64_t)(gen & 1) * P.R + (int64_t)(it % P.R);
}
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
#pragma unroll 4                     // independent draws side by side: one lane runs this while its warp waits
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

// enqueue iterations [t0, t1) of chain c under generation g (the chain's state and generation are already published)
__device__ __forceinline__ void gync_mcmc_enqueue(const GyncMcmcP& P, int64_t c, int g, long long t0, long long t1) {
  if (t1 <= t0) return;
  for (long long it = t0; it < t1; ++it) __stcg(P.slot_tag + gync_mcmc_slot(P, c, g, it), gync_mcmc_tag(g, it));
  __threadfence();
  const unsigned long long q0 = atomicAdd(&P.ctl->tail, (unsigned long long)(t1 - t0));
  for (long long it = t0; it < t1; ++it) {
    const unsigned long long t = q0 + (unsigned long long)(it - t0);
    *(volatile unsigned long long*)(P.queue + (t & P.qmask)) =
        (((t >> P.qshift) + 1ull) << 40) | (unsigned long long)gync_mcmc_slot(P, c, g, it);
  }
}

// Decide as many iterations of chain c as have their result (the sequential Metropolis rule, in iteration order), then
// top the proposals in flight up to W.  Called by exactly one lane at a time per chain (gync_mcmc_try_advance).
__device__ __noinline__ void gync_mcmc_advance(const GyncMcmcP& P, int64_t c) {
  // chain words were last written by another lane, possibly on another SM: read them from L2, never from this SM's L1
  long long h = __ldcg(P.itc + c);
  if (h >= P.iters) return;
  long long t = __ldcg(P.tailc + c);
  int g = __ldcg(P.gen + c);
  int issued = __ldcg(P.issued + c);
  double lf = __ldcg(P.logf + c);
  long long nacc = __ldcg(P.naccept + c);
  const uint64_t cid = gync_mcmc_chain_id(P, c);
  const long long h0 = h;
  while (h < P.iters) {
    const int64_t slot = gync_mcmc_slot(P, c, g, h);
    if (h >= t || __ldcg(P.done_tag + slot) != gync_mcmc_tag(g, h)) break;
    const double lfp = __ldcg(P.lfp + slot);
    double u, u1;
    if (P.inj_u) u = P.inj_u[h + P.iters * c];
    else gync_draw_u2(P.seed, cid, P.iter_offset + (uint64_t)h, (uint32_t)(GYNC_NX / 2), u, u1);
    const bool acc = u < exp(lfp - lf);              // Mamba: rand() < exp(logf(x') - logf(x))
    if (acc) {
      g = (int)(((unsigned)g + 1u) & (unsigned)GYNC_MCMC_GEN_MASK);
      __stcg(P.ncomp + 2 * c + (g & 1), 0);          // the new generation's report counter (its last users are long gone)
      __threadfence();
      __stcg(P.gen + c, g);                          // cancels every younger item of the old generation
      __threadfence();
      double prop[GYNC_NX];
      gync_mcmc_proposal(P, c, h, prop);             // the same draws the item was built from
      for (int d = 0; d < GYNC_NX; ++d) __stcg(P.state + c + P.C * d, prop[d]);
      lf = lfp;
      ++nacc;
      issued = 0;
    }
    if ((h + 1) % P.thin == 0) {
      const int64_t row = (h + 1) / P.thin - 1;
      if (P.samples_out && row < P.n_out)
        for (int d = 0; d < GYNC_NX; ++d)
          P.samples_out[row + P.n_out * (d + (int64_t)GYNC_NX * c)] = exp(__ldcg(P.state + c + P.C * d));   // sampling.jl:74
    }
    ++h;
    if (acc) { t = h; break; }                       // nothing of the new generation exists yet
  }
  if (h != h0) {
    __stcg(P.logf + c, lf);
    __stcg(P.naccept + c, nacc);
    __stcg(P.itc + c, h);
    atomicAdd(&P.ctl->decided, (unsigned long long)(h - h0));
  }
  if (h >= P.iters) {
    if (h != h0) { __threadfence(); atomicAdd(&P.ctl->chains_done, 1); }
    return;
  }
  // issue: keep this chain's share of proposals in flight, no further ahead of the oldest undecided iteration than the
  // ring allows.  The share is proportional to the chain's REMAINING iterations: chains differ in cost per iteration by
  // tens of percent, a fixed share would let the cheap ones finish early and leave their lanes idle while each slow chain
  // could use no more than its own share; this way a chain that falls behind gets more lanes and all finish together.
  const unsigned long long dec = gync_ld_vol_u64(&P.ctl->decided);
  const double rem_all = (double)(P.total_iters > dec ? P.total_iters - dec : 1ull);
  const long long share = (long long)ceil((double)P.inflight_target * (double)(P.iters - h) / rem_all);
  const long long want = min((long long)P.Wmax, max(1ll, share));
  const int inflight = issued - __ldcg(P.ncomp + 2 * c + (g & 1));
  const long long room = min(h + (long long)P.R, (long long)P.iters) - t;
  const long long n_new = min(want - (long long)inflight, room);
  if (n_new > 0) {
    __stcg(P.tailc + c, t + n_new);
    __stcg(P.issued + c, issued + (int)n_new);
    __threadfence();
    gync_mcmc_enqueue(P, c, g, t, t + n_new);
  } else if (h != h0) {
    __stcg(P.tailc + c, t);
    __stcg(P.issued + c, issued);
  }
}

// a lane has news for chain c: become the advancer, or leave the news to the lane that already is
__device__ __forceinline__ void gync_mcmc_try_advance(const GyncMcmcP& P, int64_t c) {
  if (atomicAdd(P.pending + c, 1) != 0) return;
  for (;;) {
    const int seen = *(volatile int*)(P.pending + c);
    __threadfence();
    gync_mcmc_advance(P, c);
    __threadfence();
    if (atomicSub(P.pending + c, seen) == seen) break;
  }
}

// an item's log-target is known
__device__ __forceinline__ void gync_mcmc_complete(const GyncMcmcP& P, int64_t item, unsigned long long tag, double lfp) {
  const int64_t c = item / (2 * (int64_t)P.R);
  __stcg(P.lfp + item, lfp);
  __threadfence();
  __stcg(P.done_tag + item, tag);
  __threadfence();
  atomicAdd(P.ncomp + 2 * c + (int)((tag >> 40) & 1ull), 1);
  GYNC_MP({
    const long long it = (long long)(tag & ((1ull << 40) - 1ull)) - 1;
    double u, u1;
    if (P.inj_u) u = P.inj_u[it + P.iters * c];
    else gync_draw_u2(P.seed, gync_mcmc_chain_id(P, c), P.iter_offset + (uint64_t)it, (uint32_t)(GYNC_NX / 2), u, u1);
    const double d = __ldcg(P.logf + c) + log(u) - lfp;
    const int b = !(lfp > -INFINITY) ? 10 : d < 0 ? 0 : d < 1 ? 1 : d < 3 ? 2 : d < 10 ? 3 : d < 30 ? 4 : d < 100 ? 5 : d < 300 ? 6 : d < 1e3 ? 7 : d < 1e4 ? 8 : 9;
    atomicAdd(&P.ctl->deficit[b], 1ull);
  })
  gync_mcmc_try_advance(P, c);
}
// an item was dropped or given up: it only reports that it is no longer in flight
__device__ __forceinline__ void gync_mcmc_cancelled(const GyncMcmcP& P, int64_t c, int gen) {
  atomicAdd(P.ncomp + 2 * c + (gen & 1), 1);
  atomicAdd(&P.ctl->cancelled, 1ull);
}

// first window of every chain (after the initial log-targets are known)
__global__ void gync_mcmc_persist_init_kernel(GyncMcmcP P) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= P.C) return;
  const long long h = P.itc[c];
  if (h >= P.iters) { atomicAdd(&P.ctl->chains_done, 1); return; }
  const long long t1 = min(h + (long long)P.W, (long long)P.iters);
  P.tailc[c] = t1;
  P.issued[c] = (int)(t1 - h);
  gync_mcmc_enqueue(P, c, P.gen[c], h, t1);
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
  long long ticket = -1, idle_since = 0;
  unsigned long long idle_tail = ~0ull;
  int64_t c = 0, item = 0;
  unsigned long long tag = 0;
  int mygen = 0;
  double T = 0.0, t = 0.0, tn = 0.0, lpr = 0.0, accmax = INFINITY;
  int i0 = 0, i1 = 0, budget = 0, st = GYNC_ST_OK, mpat = 0;
  GyncCtrl ctl;
  gync_ctrl_init(ctl, 1e-3);
  GyncLlhSinkReg sink = {P.data, P.iscale, 0.0};
  GYNC_MP(unsigned long long p_act = 0, p_idle = 0, p_fetch = 0, p_comp = 0, p_step = 0, p_sleep = 0;)
  for (;;) {
    GYNC_MP(const long long p_t0 = clock64();)
    if (!active && !done) {
      if (ticket < 0) ticket = (long long)atomicAdd(&P.ctl->head, 1ull);
      const unsigned long long ent = gync_ld_vol_u64(P.queue + ((unsigned long long)ticket & P.qmask));
      if ((ent >> 40) == (((unsigned long long)ticket >> P.qshift) + 1ull)) {
        __threadfence();
        item = (int64_t)(ent & ((1ull << 40) - 1ull));
        ticket = -1;
        idle_tail = ~0ull;
        c = item / (2 * (int64_t)P.R);
        tag = __ldcg(P.slot_tag + item);
        mygen = (int)((tag >> 40) & GYNC_MCMC_GEN_MASK);
        const long long it = (long long)(tag & ((1ull << 40) - 1ull)) - 1;
        bool live = (__ldcg(P.gen + c) == mygen);
        double x[GYNC_NX];
        if (live) {
          gync_mcmc_proposal(P, c, it, x);
          __threadfence();
          live = (*(volatile int*)(P.gen + c) == mygen);      // the state was not replaced while it was being read
        }
        if (!live) {
          gync_mcmc_cancelled(P, c, mygen);                   // its chain has moved on: nothing to solve
        } else {
          double sumx = 0.0;
          for (int d = 0; d < GYNC_NX; ++d) sumx += x[d];
#pragma unroll 4
          for (int d = 0; d < GYNC_NX; ++d) x[d] = exp(x[d]);
          const double lp = gync_logprior(x, P.pr);          // logf(x) = logpost(c, exp(x)) + sum(x)   (model.jl:102)
          mpat = P.patient_of_chain ? P.patient_of_chain[c] : 0;
          if (!(lp > -INFINITY) || P.allnan[mpat]) {
            // outside the prior the likelihood is skipped (model.jl:121); an all-missing table scores -Inf (model.jl:161)
            gync_mcmc_complete(P, item, tag, -INFINITY);
          } else {
            lpr = lp + sumx;
            {
              // the residual sum beyond which this proposal is rejected for certain (a hair above the exact bound, so that the
              // rounding of exp / log in the rule itself cannot matter); +Inf when the chain sits at -Inf and accepts anything
              double u, u1;
              if (P.inj_u) u = P.inj_u[it + P.iters * c];
              else gync_draw_u2(P.seed, gync_mcmc_chain_id(P, c), P.iter_offset + (uint64_t)it, (uint32_t)(GYNC_NX / 2), u, u1);
              const double b = 2.0 * (lpr - __ldcg(P.logf + c) - log(u));
              accmax = (b == b) ? b + 1e-9 * fabs(b) + 1e-9 : INFINITY;
            }
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
        }
      } else {
        if (*(volatile int*)&P.ctl->chains_done >= (int)P.C || *(volatile int*)&P.ctl->abort) {
          done = true;
        } else {
          // watchdog (never expected to fire): nothing was enqueued anywhere for idle_limit ticks while work is outstanding
          const unsigned long long qt = gync_ld_vol_u64(&P.ctl->tail);
          const long long now = clock64();
          if (qt != idle_tail) { idle_tail = qt; idle_since = now; }
          else if (now - idle_since > P.idle_limit) { atomicExch(&P.ctl->abort, 1); done = true; }
        }
      }
    }
    GYNC_MP(p_fetch += (unsigned long long)(clock64() - p_t0);)
    __syncwarp();
    if (__all_sync(0xffffffffu, done)) break;
    if (!__any_sync(0xffffffffu, active)) {               // nobody in this warp has a sample: do not burn a step on it
      __nanosleep(2000);
      GYNC_MP(++p_sleep;)
      continue;
    }
    GYNC_MP(if (active) ++p_act; else ++p_idle;)
    GYNC_MP(const long long p_t1 = clock64();)
    // ---- one step attempt, executed by ALL lanes of the warp (TMEM ops are warp-collective)
    const bool stepping = active && st == GYNC_ST_OK && tn > t;
    double h = ctl.h;
    if (o.hmax > 0.0) h = fmin(h, o.hmax);
    const double rem = tn - t;
    const bool last = stepping && (h * 1.0001 >= rem);
    if (last) h = rem;
    if (!stepping || !(h > 0.0)) h = 1.0;                  // idle lanes: any finite positive step
    const double hprop = ctl.h;
    // has this item's chain moved on?  (asked before the step, read after it: the L2 round trip hides behind the step)
    const int gen_now = active ? __ldcg(P.gen + c) : mygen;
    const double err = gync_ros_step_tm(W, h, o.rtol, o.atol);
    __syncwarp();
    GYNC_MP(p_step += (unsigned long long)(clock64() - p_t1);)
    GYNC_MP(const long long p_t2 = clock64();)
    if (stepping) {
      if (budget-- <= 0) {
        st = GYNC_ST_MAXSTEPS;
      } else if (!(h > 1e-14 * fmax(1.0, fabs(t)))) {
        st = GYNC_ST_HMIN;
      } else {
        if (gync_ctrl_update(ctl, h, err, true)) {       // a non-finite candidate arrives as err = NaN
          t = last ? tn : t + h;
#pragma unroll
          for (int i = 0; i < GYNC_NY; ++i) W.y[i] = W.yn[i];
          if (last && h < hprop) ctl.h = fmax(ctl.h, hprop);
        }
      }
    }
    if (active) {
      bool finished = (st != GYNC_ST_OK);
      const bool cancelled = (gen_now != mygen) || *(volatile int*)&P.ctl->abort;
      if (!finished && !cancelled && t >= tn) {
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
      const bool hopeless = sink.acc > accmax;             // already past what acceptance allows: rejected for certain
      if (hopeless && !finished && !cancelled) atomicAdd(&P.ctl->early, 1ull);
      if (finished || cancelled || hopeless) {
        atomicAdd(&P.ctl->steps, (unsigned long long)(ctl.nacc + ctl.nrej));
        if (!cancelled) {
          atomicAdd(&P.ctl->hist[min(23, 31 - __clz(max(1, ctl.nacc + ctl.nrej)))], 1ull);
          if (st != GYNC_ST_OK) atomicAdd(&P.ctl->failed, 1ull);
        }
        if (cancelled || __ldcg(P.gen + c) != mygen) gync_mcmc_cancelled(P, c, mygen);
        // failed solve -> NaN trajectory -> -Inf (gyncycle.jl:18-19, model.jl:138-141)
        else gync_mcmc_complete(P, item, tag, (st == GYNC_ST_OK && !hopeless) ? __dadd_rn(lpr, __dmul_rn(-0.5, sink.acc)) : -INFINITY);
        active = false;
#pragma unroll
        for (int i = 0; i < GYNC_NY; ++i) if (!(W.y[i] * 0.0 == 0.0)) W.y[i] = 1.0;
      }
    }
    GYNC_MP(__syncwarp(); p_comp += (unsigned long long)(clock64() - p_t2);)
  }
  GYNC_MP(atomicAdd(&P.ctl->trips_active, p_act); atomicAdd(&P.ctl->trips_idle, p_idle); atomicAdd(&P.ctl->cyc_fetch, p_fetch);
          atomicAdd(&P.ctl->cyc_complete, p_comp); atomicAdd(&P.ctl->cyc_step, p_step); atomicAdd(&P.ctl->sleeps, p_sleep);)
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" :: "r"(tmem_base) : "memory");
  }
}
