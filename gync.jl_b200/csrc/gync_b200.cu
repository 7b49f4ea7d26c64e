// C ABI of the B200-native GynC.jl hot path (declared in include/gync_b200.h).
// Host wrappers: argument checks, H2D/D2H staging, kernel launches.  No CPU fallback.
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <string>
#include <vector>
#include "gync_kernels.cuh"
#include "gync_mcmc.cuh"
#include "gync_mcmc_persist.cuh"
#include "gync_amm.cuh"
#include "gync_sched.cuh"
#include "../../include/gync_b200.h"

extern "C" int gync_init(int device);
extern "C" void gync_shutdown(void);

namespace {

thread_local std::string g_err;
cudaStream_t g_stream = nullptr;
int g_device = -1;
int g_num_sms = 0;
std::atomic<int64_t> g_launches{0};

// Devices this process drives (gync_init: one; gync_init_multi: several).  The g_* variables above always
// describe the device that is current; use_device() switches them.
struct DevCtx {
  int device = -1;
  cudaStream_t stream = nullptr;
  int num_sms = 0;
};
std::vector<DevCtx> g_devs;

int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}

#define CK(expr)                                                                                   \
  do {                                                                                             \
    cudaError_t _e = (expr);                                                                       \
    if (_e != cudaSuccess) {                                                                       \
      return fail(GYNC_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));              \
    }                                                                                              \
  } while (0)

int ensure_init() {
  if (g_device >= 0) return GYNC_OK;
  return gync_init(0);
}

int use_device(int i) {
  if (i < 0 || i >= (int)g_devs.size()) return fail(GYNC_ERR_ARG, "device slot out of range");
  CK(cudaSetDevice(g_devs[i].device));
  g_device = g_devs[i].device;
  g_stream = g_devs[i].stream;
  g_num_sms = g_devs[i].num_sms;
  return GYNC_OK;
}

// RAII stream-ordered temporary (freed on every exit path)
struct ABuf {
  void* p = nullptr;
  cudaStream_t s = nullptr;
  ~ABuf() { if (p) cudaFreeAsync(p, s); }
  cudaError_t alloc(size_t bytes, cudaStream_t st) { s = st; return cudaMallocAsync(&p, bytes ? bytes : 8, st); }
  template <class T> T* as() { return static_cast<T*>(p); }
};

// RAII device buffer
struct DBuf {
  void* p = nullptr;
  ~DBuf() { if (p) cudaFree(p); }
  cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 8); }
  template <class T> T* as() { return static_cast<T*>(p); }
};

// Grow-only staging buffers of the blocking host-buffer entry points (one set per device slot): repeated calls of the
// same size reuse them instead of paying cudaMalloc / cudaFree (both synchronise the device) around every batch.
struct StagePool {
  void* p[8] = {nullptr};
  size_t cap[8] = {0};
  int device = -1;
  void* get(int id, size_t bytes, int dev) {
    if (device != dev) { release(); device = dev; }
    if (cap[id] < bytes || !p[id]) {
      if (p[id]) cudaFree(p[id]);
      p[id] = nullptr; cap[id] = 0;
      const size_t want = bytes ? bytes + bytes / 8 : 8;
      if (cudaMalloc(&p[id], want) != cudaSuccess) { (void)cudaGetLastError(); return nullptr; }
      cap[id] = want;
    }
    return p[id];
  }
  void release() {
    if (device >= 0) {
      int cur = 0;
      cudaGetDevice(&cur);
      cudaSetDevice(device);
      for (int i = 0; i < 8; ++i) { if (p[i]) cudaFree(p[i]); p[i] = nullptr; cap[i] = 0; }
      cudaSetDevice(cur);
    }
    device = -1;
  }
};
StagePool g_stage[8];

GyncSolverOpts to_opts(const gync_solver_opts* o) {
  GyncSolverOpts r;
  r.rtol = (o && o->rtol > 0) ? o->rtol : GYNC_DEFAULT_RTOL;
  r.atol = (o && o->atol > 0) ? o->atol : GYNC_DEFAULT_ATOL;
  r.h0 = o ? o->h0 : 0.0;
  r.hmax = o ? o->hmax : 0.0;
  r.max_steps = (o && o->max_steps > 0) ? o->max_steps : 200000;
  return r;
}

inline unsigned blocks_for(int64_t n, int bs) { return (unsigned)((n + bs - 1) / bs); }

// Wait for a stream without spinning AND without the driver's blocking-sync wake-up.  Round 1 replaced the spinning
// cudaStreamSynchronize (a core burnt for hundreds of milliseconds; on CPU-quota'd hosts that showed up as stalls of the
// calling process) by a cudaEventBlockingSync event.  Traced in round 2 (GYNC_TRACE_CALLS=1): on the shared hosts of this
// pool that wake-up itself takes ~10 ms when it is fast and 100 - 350 ms when it is not, AFTER the device has finished —
// the "sporadic stalls" of the end-to-end figure.  So: poll cudaStreamQuery and sleep in between (20 us at first — short
// kernels —, then 200 us: about 1 % of a core while a 480 ms batch runs).
cudaError_t wait_stream(cudaStream_t s) {
  for (int i = 0;; ++i) {
    const cudaError_t e = cudaStreamQuery(s);
    if (e != cudaErrorNotReady) return e;
    timespec ts = {0, i < 64 ? 20000L : 200000L};
    nanosleep(&ts, nullptr);
  }
}

// opt in to the large dynamic shared-memory carve-out the solver kernels need
template <class K>
int prep_solver_kernel(K kern, size_t smem, int block) {
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  (void)block;
  return GYNC_OK;
}

}  // namespace

extern "C" int gync_init(int device);

namespace {
// Longest-first ordering of the sample queue, learned from the step counts of earlier batches (gync_sched.cuh).
struct SchedState {
  GyncSchedModel* mdl = nullptr;
  double *G = nullptr, *b = nullptr, *Phi = nullptr, *target = nullptr, *pred = nullptr;
  int *order = nullptr, *hist = nullptr, *steps = nullptr;
  unsigned long long* minmax = nullptr;
  cudaEvent_t busy = nullptr;      // last use of these buffers: calls on other streams wait for it
  int64_t cap = 0;
  bool fitted = false;
  int device = -1;
  void release() {
    if (device >= 0) {
      int cur = 0;
      cudaGetDevice(&cur);
      cudaSetDevice(device);
      for (void* p : {(void*)mdl, (void*)G, (void*)b, (void*)Phi, (void*)target, (void*)pred, (void*)order, (void*)hist, (void*)steps, (void*)minmax})
        if (p) cudaFree(p);
      if (busy) cudaEventDestroy(busy);
      cudaSetDevice(cur);
    }
    *this = SchedState();
  }
  int ensure(int64_t N, int dev) {
    if (device == dev && cap >= N) return GYNC_OK;
    const bool keep = (device == dev) && fitted;
    GyncSchedModel* old = keep ? mdl : nullptr;
    if (keep) mdl = nullptr;                      // survive the reallocation below
    const bool was_fitted = keep;
    release();
    device = dev;
    cap = N + N / 4;
    if (old) mdl = old; else { CK(cudaMalloc((void**)&mdl, sizeof(GyncSchedModel))); CK(cudaMemset(mdl, 0, sizeof(GyncSchedModel))); }
    fitted = was_fitted;
    CK(cudaMalloc((void**)&G, sizeof(double) * GYNC_SCHED_NF * GYNC_SCHED_NF));
    CK(cudaMalloc((void**)&b, sizeof(double) * GYNC_SCHED_NF));
    CK(cudaMalloc((void**)&Phi, sizeof(double) * GYNC_SCHED_MAXS * GYNC_SCHED_NF));
    CK(cudaMalloc((void**)&target, sizeof(double) * GYNC_SCHED_MAXS));
    CK(cudaMalloc((void**)&pred, sizeof(double) * cap));
    CK(cudaMalloc((void**)&order, sizeof(int) * cap));
    CK(cudaMalloc((void**)&steps, sizeof(int) * 2 * cap));
    CK(cudaMalloc((void**)&hist, sizeof(int) * GYNC_SCHED_NB));
    CK(cudaMalloc((void**)&minmax, sizeof(unsigned long long) * 2));
    CK(cudaEventCreateWithFlags(&busy, cudaEventDisableTiming));
    return GYNC_OK;
  }
};
SchedState g_sched[8];

// Order the queue of a batch of N samples X (N x 116) by predicted cost, if a model has been fitted; returns the state to
// hand to sched_refit after the solver kernel, or NULL when scheduling is off / does not pay / cannot allocate.
// *steps is replaced by an internal buffer when the caller did not ask for step counts (the refit needs them).
SchedState* sched_prepare(const double* dX, int64_t N, cudaStream_t s, int** order, int32_t** steps) {
  *order = nullptr;
  // scheduling pays once there is more than one solve per lane; below that every lane starts at once anyway
  if (getenv("GYNC_NO_SCHED") || !dX || 4 * N < 5 * (int64_t)g_num_sms * GYNC_TM_NSLOT || N >= (1LL << 31)) return nullptr;
  int slot = 0;
  for (size_t i = 0; i < g_devs.size(); ++i) if (g_devs[i].device == g_device) slot = (int)i;
  SchedState* st = &g_sched[slot];
  if (st->ensure(N, g_device) != GYNC_OK) {         // no memory for the scheduler: run unordered
    (void)cudaGetLastError();
    st->release();
    return nullptr;
  }
  if (cudaStreamWaitEvent(s, st->busy, 0) != cudaSuccess) return nullptr;     // the buffers are shared by every stream of this device
  if (!*steps) *steps = st->steps;
  if (st->fitted) {
    cudaMemsetAsync(st->minmax, 0xFF, sizeof(unsigned long long), s);
    cudaMemsetAsync(st->minmax + 1, 0, sizeof(unsigned long long), s);
    cudaMemsetAsync(st->hist, 0, sizeof(int) * GYNC_SCHED_NB, s);
    gync_sched_predict_kernel<<<blocks_for(N, 256), 256, 0, s>>>(dX, N, st->mdl, st->pred, st->minmax);
    gync_sched_hist_kernel<<<blocks_for(N, 256), 256, 0, s>>>(st->pred, N, st->minmax, st->hist);
    gync_sched_scan_kernel<<<1, GYNC_SCHED_NB, 0, s>>>(st->hist);
    gync_sched_scatter_kernel<<<blocks_for(N, 256), 256, 0, s>>>(st->pred, N, st->minmax, st->hist, st->order);
    g_launches += 4;
    *order = st->order;
  }
  return st;
}

// refit the cost model on this batch's step counts (N x 2: accepted, rejected)
int sched_refit(SchedState* st, const double* dX, int64_t N, const int32_t* steps, cudaStream_t s) {
  const int S = (int)std::min<int64_t>(N, GYNC_SCHED_MAXS);
  const size_t fit_smem = sizeof(double) * (GYNC_SCHED_NF * (GYNC_SCHED_NF + 1) + GYNC_SCHED_NF);
  CK(cudaFuncSetAttribute(gync_sched_fit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fit_smem));
  gync_sched_features_kernel<<<S, 128, 0, s>>>(dX, N, S, N / S, steps, st->Phi, st->target, st->mdl);
  gync_sched_normal_kernel<<<GYNC_SCHED_NF, 128, 0, s>>>(st->Phi, st->target, S, st->G, st->b);
  gync_sched_fit_kernel<<<1, 128, fit_smem, s>>>(st->G, st->b, st->mdl);
  g_launches += 3;
  st->fitted = true;
  CK(cudaEventRecord(st->busy, s));
  return GYNC_OK;
}

// One launch of the solver kernel: persistent lanes, one CTA (4 warps x 32 samples, state in shared
// + tensor memory) per SM.  `pat` (nullable) gives every sample its own patient -> LL is N x 1.
int launch_loglik(const double* dX, int64_t N, const double* d_data, const double* iscale, const int* allnan, int M,
                  const int* pat, const GyncSolverOpts& so, double* dLL, double* dYmeas, int32_t* d_status,
                  int32_t* d_nsteps, cudaStream_t s) {
  ABuf cnt;
  CK(cnt.alloc(sizeof(unsigned long long), s));
  unsigned long long* counter = cnt.as<unsigned long long>();
  CK(cudaMemsetAsync(counter, 0, sizeof(unsigned long long), s));
  if (int rc = prep_solver_kernel(gync_loglik_tm_kernel, GyncWorkTm::kSmemBytes, 128)) return rc;
  const int64_t want = (N + GYNC_TM_NSLOT - 1) / GYNC_TM_NSLOT;
  const int nb = (int)std::min<int64_t>(want, (int64_t)g_num_sms);
  int* order = nullptr;
  int32_t* steps = d_nsteps;
  SchedState* st = sched_prepare(dX, N, s, &order, &steps);
  gync_loglik_tm_kernel<<<nb, 128, GyncWorkTm::kSmemBytes, s>>>(dX, N, d_data, iscale, allnan, M, pat, so, dLL, dYmeas,
                                                                d_status, steps, counter, order);
  g_launches += 1;
  if (st) if (int rc = sched_refit(st, dX, N, steps, s)) return rc;
  CK(cudaGetLastError());
  return GYNC_OK;
}
}  // namespace

namespace {
// gync / forwardsol on a shared time grid: the tensor-memory kernel (persistent lanes, 128 samples per SM); the older
// shared-memory kernel (64 per SM, static assignment) is kept behind GYNC_SOLVE_SM for A/B runs.
int launch_solve_grid(const double* dX, const double* dY0, const double* dP, int64_t N, const double* d_t, int nT,
                      const GyncSolverOpts& so, double* dY, int32_t* d_status, cudaStream_t s) {
  if (getenv("GYNC_SOLVE_SM")) {
    if (int rc = prep_solver_kernel(gync_solve_kernel<GYNC_SOLVE_BLOCK>, GyncWorkSm<GYNC_SOLVE_BLOCK>::kSmemBytes, GYNC_SOLVE_BLOCK)) return rc;
    gync_solve_kernel<GYNC_SOLVE_BLOCK><<<std::min<unsigned>(blocks_for(N, GYNC_SOLVE_BLOCK), g_num_sms), GYNC_SOLVE_BLOCK,
                                           GyncWorkSm<GYNC_SOLVE_BLOCK>::kSmemBytes, s>>>(dX, dY0, dP, N, d_t, nT, so, dY, d_status);
  } else {
    ABuf cnt;
    CK(cnt.alloc(sizeof(unsigned long long), s));
    CK(cudaMemsetAsync(cnt.p, 0, sizeof(unsigned long long), s));
    if (int rc = prep_solver_kernel(gync_solve_tm_kernel, GyncWorkTm::kSmemBytes, 128)) return rc;
    const int nb = (int)std::min<int64_t>((N + GYNC_TM_NSLOT - 1) / GYNC_TM_NSLOT, (int64_t)g_num_sms);
    int* order = nullptr;
    int32_t* steps = nullptr;
    SchedState* st = sched_prepare(dX, N, s, &order, &steps);      // samples given as X only (the cost model's features)
    gync_solve_tm_kernel<<<nb, 128, GyncWorkTm::kSmemBytes, s>>>(dX, dY0, dP, N, d_t, nT, so, dY, d_status, cnt.as<unsigned long long>(),
                                                                 order, steps);
    if (st) if (int rc = sched_refit(st, dX, N, steps, s)) return rc;
  }
  g_launches += 1;
  CK(cudaGetLastError());
  return GYNC_OK;
}
}  // namespace

// EM launch helpers (templates: outside extern "C")
namespace {

// Unified EM launch: one cooperative launch of gync_em_kernel for all iterations; L resident in shared
// memory when this GPU's rows fit.  `peer` describes the other ranks (world 1: none).
struct EmLaunchState {          // synchronisation words + scratch of one (device, communicator)
  unsigned long long* sync_words = nullptr;    // [0] arrive, [1] ready
  double* partials = nullptr;                  // 1024 x 64
  double* gnorm = nullptr;                     // 2 x 64
  int* err = nullptr;
  unsigned long long seq = 1, arrive0 = 0;
  int device = -1;
  // The arrival counter, the partials and the published denominators are ONE set per (device, communicator): two launches
  // that overlap on different streams would share them.  Every launch therefore waits for the previous one's event and
  // records its own — launches through this state are serialised on the device whatever stream they were given.
  cudaEvent_t busy = nullptr;
  int alloc() {
    CK(cudaMalloc((void**)&sync_words, 2 * sizeof(unsigned long long)));
    CK(cudaMemset(sync_words, 0, 2 * sizeof(unsigned long long)));
    CK(cudaMalloc((void**)&partials, sizeof(double) * 1024 * 64));
    CK(cudaMalloc((void**)&gnorm, sizeof(double) * 2 * 64));
    CK(cudaMalloc((void**)&err, sizeof(int)));
    CK(cudaMemset(err, 0, sizeof(int)));
    CK(cudaEventCreateWithFlags(&busy, cudaEventDisableTiming));
    CK(cudaGetDevice(&device));
    return GYNC_OK;
  }
  // blocking callers: did a spin time out inside the last launches?  (the kernel then stops updating and raises this)
  int check_error(cudaStream_t s, const char* who) {
    int e = 0;
    CK(cudaMemcpyAsync(&e, err, sizeof(int), cudaMemcpyDeviceToHost, s));
    CK(wait_stream(s));
    if (e) {
      cudaMemsetAsync(err, 0, sizeof(int), s);
      return fail(GYNC_ERR_CUDA, std::string(who) + ": a synchronisation spin inside the EM kernel timed out");
    }
    return GYNC_OK;
  }
  void release() {
    if (device >= 0) { int cur; cudaGetDevice(&cur); cudaSetDevice(device);
      cudaFree(sync_words); cudaFree(partials); cudaFree(gnorm); cudaFree(err); if (busy) cudaEventDestroy(busy); cudaSetDevice(cur); }
    *this = EmLaunchState();
  }
};
EmLaunchState g_em_local[8];    // single-GPU calls: one per device slot

template <int MH, bool FULL, bool RESIDENT>
int em_launch_variant(EmLaunchState& st, const GyncEmPeer& peer, double* dw, const double* dL, int64_t K, int M, int niter,
                      double* d_hist, cudaStream_t s, bool* done) {
  *done = false;
  const int64_t rows_per_pass = GYNC_EM_WARPS * 16;
  int nb;
  int64_t R = 0, RR = 0;
  size_t smem = sizeof(EmSmem<MH>) + 16;
  if (RESIDENT) {
    nb = g_num_sms;
    R = (K + nb - 1) / nb;
    int max_optin = 0;
    CK(cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, g_device));
    // rows of a block that shared memory (and the three register-held weights per thread) can hold; the rest streams.
    // Worth it while at least half of the rows are resident; otherwise the caller streams everything.
    const int64_t cap = std::min<int64_t>(((int64_t)max_optin - (int64_t)smem) / (int64_t)(sizeof(double) * M), 3 * rows_per_pass);
    RR = std::min<int64_t>(R, cap);
    if (RR < 1 || R - RR > RR || (RR < R && getenv("GYNC_EM_NO_HYBRID"))) return GYNC_OK;
    smem += sizeof(double) * (size_t)M * RR;
  } else {
    nb = (int)std::min<int64_t>(g_num_sms, std::max<int64_t>((K + rows_per_pass - 1) / rows_per_pass, 1));
  }
  const void* kern = !RESIDENT ? ((peer.world > 1) ? (const void*)gync_em_kernel<MH, FULL, 0, true> : (const void*)gync_em_kernel<MH, FULL, 0, false>)
                     : RR == R ? ((peer.world > 1) ? (const void*)gync_em_kernel<MH, FULL, 1, true> : (const void*)gync_em_kernel<MH, FULL, 1, false>)
                               : ((peer.world > 1) ? (const void*)gync_em_kernel<MH, FULL, 2, true> : (const void*)gync_em_kernel<MH, FULL, 2, false>);
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 0;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, GYNC_EM_THREADS, smem));
  if (per_sm < 1) return RESIDENT ? GYNC_OK : fail(GYNC_ERR_CUDA, "EM kernel does not fit on an SM");
  GyncEmSync sync = {st.sync_words, st.sync_words + 1, st.seq, st.arrive0};
  GyncEmPeer pr = peer;
  int Ri = (int)R, RRi = (int)RR;
  void* args[] = {&dw, &dL, &K, &M, &niter, &Ri, &RRi, &st.partials, &st.gnorm, &sync, &pr, &st.err, &d_hist};
  if (st.busy) CK(cudaStreamWaitEvent(s, st.busy, 0));
  CK(cudaLaunchCooperativeKernel(kern, dim3(nb), dim3(GYNC_EM_THREADS), args, smem, s));
  if (st.busy) CK(cudaEventRecord(st.busy, s));
  st.seq += (unsigned long long)niter + 1ull;
  st.arrive0 += (unsigned long long)nb * ((unsigned long long)niter + 1ull);
  g_launches += 1;
  *done = true;
  return GYNC_OK;
}

template <int MH, bool FULL>
int em_launch(EmLaunchState& st, const GyncEmPeer& peer, double* dw, const double* dL, int64_t K, int M, int niter,
              double* d_hist, cudaStream_t s) {
  bool done = false;
  if (niter >= 2 && !getenv("GYNC_EM_NO_RESIDENT")) {
    if (int rc = em_launch_variant<MH, FULL, true>(st, peer, dw, dL, K, M, niter, d_hist, s, &done)) return rc;
    if (done) return GYNC_OK;
  }
  return em_launch_variant<MH, FULL, false>(st, peer, dw, dL, K, M, niter, d_hist, s, &done);
}

template <int MH, bool FULL>
int em_persistent_launch(double* dw, const double* dL, int64_t K, int M, int niter, double* d_hist, cudaStream_t s) {
  int slot = 0;
  for (size_t i = 0; i < g_devs.size(); ++i) if (g_devs[i].device == g_device) slot = (int)i;
  EmLaunchState& st = g_em_local[slot];
  if (st.device != g_device) { st.release(); if (int rc = st.alloc()) return rc; }
  GyncEmPeer peer;
  memset(&peer, 0, sizeof(peer));
  peer.world = 1;
  return em_launch<MH, FULL>(st, peer, dw, dL, K, M, niter, d_hist, s);
}

struct EmScratch {   // scratch for the sharded step kernels (on the device that was current when allocated)
  double* partials = nullptr;
  size_t doubles = 0;
  int device = -1;
  cudaEvent_t busy = nullptr;   // one scratch per device: users on different streams are serialised through this event
} g_em_scratch;

int em_grid(int64_t K) {
  const int64_t rows_per_block = GYNC_EM_WARPS * 16;
  int64_t want = (K + rows_per_block - 1) / rows_per_block;
  return (int)std::min<int64_t>((int64_t)g_num_sms, std::max<int64_t>(want, 1));
}

int em_scratch(size_t doubles) {
  if (g_em_scratch.partials && g_em_scratch.doubles >= doubles && g_em_scratch.device == g_device) return GYNC_OK;
  if (g_em_scratch.partials) {
    cudaSetDevice(g_em_scratch.device);
    cudaFree(g_em_scratch.partials);              // synchronises the device: no user of the old buffer is left
    if (g_em_scratch.busy) cudaEventDestroy(g_em_scratch.busy);
    cudaSetDevice(g_device);
  }
  g_em_scratch.partials = nullptr;
  g_em_scratch.busy = nullptr;
  CK(cudaEventCreateWithFlags(&g_em_scratch.busy, cudaEventDisableTiming));
  CK(cudaMalloc((void**)&g_em_scratch.partials, sizeof(double) * doubles));
  g_em_scratch.doubles = doubles;
  g_em_scratch.device = g_device;
  return GYNC_OK;
}

template <int MH, bool FULL>
int em_norms_launch(const double* dw, const double* dL, int64_t K, int M, double* d_out, cudaStream_t s) {
  const int nb = em_grid(K);
  if (int rc = em_scratch((size_t)nb * 2 * MH)) return rc;
  CK(cudaStreamWaitEvent(s, g_em_scratch.busy, 0));
  gync_em_norms_kernel<MH, FULL><<<nb, GYNC_EM_THREADS, 0, s>>>(dw, dL, K, M, g_em_scratch.partials);
  gync_em_finish_kernel<MH><<<1, GYNC_EM_THREADS, 0, s>>>(g_em_scratch.partials, nb, M, d_out);
  CK(cudaEventRecord(g_em_scratch.busy, s));
  g_launches += 2;
  CK(cudaGetLastError());
  return GYNC_OK;
}

template <int MH, bool FULL>
int em_step_launch(double* dw, const double* dL, int64_t K, int M, const double* d_in, double* d_out, cudaStream_t s) {
  const int nb = em_grid(K);
  if (int rc = em_scratch((size_t)nb * 2 * MH)) return rc;
  CK(cudaStreamWaitEvent(s, g_em_scratch.busy, 0));
  gync_em_step_kernel<MH, FULL><<<nb, GYNC_EM_THREADS, 0, s>>>(dw, dL, K, M, d_in, g_em_scratch.partials);
  gync_em_finish_kernel<MH><<<1, GYNC_EM_THREADS, 0, s>>>(g_em_scratch.partials, nb, M, d_out);
  CK(cudaEventRecord(g_em_scratch.busy, s));
  g_launches += 2;
  CK(cudaGetLastError());
  return GYNC_OK;
}

// columns are split over two threads per row: MH = columns per thread; FULL when M == 2*MH
#define EM_DISPATCH(M, CALL)                                        \
  do {                                                              \
    if ((M) == 40) return CALL(20, true);                           \
    if ((M) <= 8) return CALL(4, false);                            \
    if ((M) <= 16) return CALL(8, false);                           \
    if ((M) <= 24) return CALL(12, false);                          \
    if ((M) <= 32) return CALL(16, false);                          \
    if ((M) <= 40) return CALL(20, false);                          \
    if ((M) <= 64) return CALL(32, false);                          \
    return fail(GYNC_ERR_ARG, "EM: M > 64 patients not supported"); \
  } while (0)

}  // namespace

extern "C" {

int gync_abi_version(void) { return 1; }

const char* gync_last_error(void) { return g_err.c_str(); }

int64_t gync_launch_count(void) { return g_launches.load(); }

// FP64 flops of one step attempt of the method the library was built with: generated-code counts (add+mul; a
// reciprocal = seed + 3 DFMA = 6, the real power = exp+log ~ 60) + the hand-written stage algebra of the step function.
//   Rodas5P (gync_rodas5p_step_tm): 8 solves, 7 RHS (their real power by series); per element of the 33-vectors: newest-k terms 7 x 3, stored-k sums
//   (0+1+2+3+4+5+5) x 4, adding the parked correction 7 x 2, absorbing k6 2 x 2, error norm 13.
//   RODAS4P (gync_rodas4_step_tm): 6 solves, 5 RHS; 33 x [2+2+4+5+6+7+8+9+1+11+1] + 33 x 13.
double gync_flops_per_step(void) {
  const double rcp = 6.0, pw = 60.0;
  const double rhs = GYNC_OPS_RHS_FLOP + rcp * GYNC_OPS_RHS_RCP + pw * GYNC_OPS_RHS_POW;
  const double rhsw = GYNC_OPS_RHSW_FLOP + rcp * GYNC_OPS_RHSW_RCP + pw * GYNC_OPS_RHSW_POW;
  const double lu = GYNC_OPS_LU_FLOP + rcp * GYNC_OPS_LU_RCP;
#if GYNC_ROS_ORDER == 5
  const double stage = 33.0 * (7 * 3 + 20 * 4 + 7 * 2 + 4) + 33.0 * 13.0;
  // the seven stage evaluations take the real power from the step's start by a 10-term series (13 FMAs = 26 flops instead of
  // the log + exp counted as 60; gync_pow0689), the first evaluation computes it in full and one extra reciprocal
  const double rhs_stage = rhs - (pw - 26.0) * GYNC_OPS_RHS_POW;
  return rhsw + rcp + 7.0 * rhs_stage + lu + 8.0 * GYNC_OPS_SOLVE_FLOP + stage;
#else
  const double stage = 33.0 * (2 + 2 + 4 + 5 + 6 + 7 + 8 + 9 + 1 + 11 + 1) + 33.0 * 13.0;
  return rhsw + 5.0 * rhs + lu + 6.0 * GYNC_OPS_SOLVE_FLOP + stage;
#endif
}

// 5: Rodas5P (default build), 4: RODAS4P (-DGYNC_ROS_ORDER=4, A/B runs)
int gync_method_order(void) { return GYNC_ROS_ORDER; }

void gync_dims(int32_t* ny, int32_t* np, int32_t* nx, int32_t* nq, int32_t* nlu) {
  if (ny) *ny = GYNC_NY;
  if (np) *np = GYNC_NP;
  if (nx) *nx = GYNC_NX;
  if (nq) *nq = GYNC_NQ;
  if (nlu) *nlu = GYNC_NLU;
}

int gync_ndev(void) { return (int)g_devs.size(); }

int gync_init_multi(int32_t ndev, const int32_t* devices) {
  if (ndev < 1 || ndev > 8) return fail(GYNC_ERR_ARG, "gync_init_multi: 1..8 devices");
  gync_shutdown();
  std::vector<DevCtx> devs;
  for (int i = 0; i < ndev; ++i) {
    const int d = devices ? devices[i] : i;
    if (int rc = gync_init(d)) return rc;          // creates the stream of device d
    devs.push_back({g_device, g_stream, g_num_sms});
    g_stream = nullptr;                            // keep it alive: owned by the table now
    g_devs.clear();
  }
  for (int i = 0; i < ndev; ++i)                   // NVLink peer access for the fused EM exchange
    for (int j = 0; j < ndev; ++j)
      if (i != j) {
        int can = 0;
        CK(cudaDeviceCanAccessPeer(&can, devs[i].device, devs[j].device));
        if (!can) return fail(GYNC_ERR_CUDA, "gync_init_multi: devices without peer access");
        CK(cudaSetDevice(devs[i].device));
        cudaError_t e = cudaDeviceEnablePeerAccess(devs[j].device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) CK(e);
        (void)cudaGetLastError();
      }
  g_devs = devs;
  return use_device(0);
}

int gync_init(int device) {
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0)
    return fail(GYNC_ERR_CUDA, std::string("no CUDA device (this library has no CPU fallback): ") +
                                   (e != cudaSuccess ? cudaGetErrorString(e) : "device count 0"));
  if (device < 0 || device >= count) return fail(GYNC_ERR_ARG, "device index out of range");
  CK(cudaSetDevice(device));
  if (g_stream && g_device != device) { cudaStreamDestroy(g_stream); g_stream = nullptr; }
  if (!g_stream) CK(cudaStreamCreateWithFlags(&g_stream, cudaStreamNonBlocking));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, device));
  g_num_sms = prop.multiProcessorCount;
  g_device = device;
  g_devs.assign(1, DevCtx{g_device, g_stream, g_num_sms});
  return GYNC_OK;
}

void gync_shutdown(void) {
  for (EmLaunchState& st : g_em_local) st.release();
  for (SchedState& st : g_sched) st.release();
  for (StagePool& sp : g_stage) sp.release();
  if (g_em_scratch.partials) {
    if (g_em_scratch.device >= 0) cudaSetDevice(g_em_scratch.device);
    cudaFree(g_em_scratch.partials);
    if (g_em_scratch.busy) cudaEventDestroy(g_em_scratch.busy);
    g_em_scratch = EmScratch();
  }
  for (DevCtx& d : g_devs)
    if (d.stream) { cudaSetDevice(d.device); cudaStreamDestroy(d.stream); if (d.stream == g_stream) g_stream = nullptr; }
  g_devs.clear();
  if (g_stream) { cudaStreamDestroy(g_stream); g_stream = nullptr; }
  g_device = -1;
}

// ------------------------------------------------------------------------------------------ K1
int gync_loglik_batch_dev(const double* dX, int64_t N, const double* d_data, int32_t M,
                          const double* d_sigma_rho, const gync_solver_opts* opts, double* dLL,
                          double* dYmeas, int32_t* d_status, int32_t* d_nsteps, void* stream) {
  if (N < 0 || M < 1 || !dLL || (N > 0 && (!dX || !d_data || !d_sigma_rho)))
    return fail(GYNC_ERR_ARG, "gync_loglik_batch: bad arguments");
  if (int rc = ensure_init()) return rc;
  if (N == 0) return GYNC_OK;
  cudaStream_t s = (cudaStream_t)stream;   // NULL = the legacy default stream, as torch uses
  ABuf bIsc, bNan, bY;
  CK(bIsc.alloc(sizeof(double) * 4 * M, s));
  CK(bNan.alloc(sizeof(int) * M, s));
  double* iscale = bIsc.as<double>();
  int* allnan = bNan.as<int>();
  gync_prep_patients_kernel<<<blocks_for(M, 64), 64, 0, s>>>(d_data, d_sigma_rho, M, iscale, allnan);
  g_launches += 1;
  CK(cudaGetLastError());
  if (M <= GYNC_FUSED_MAX_M) {       // fused epilogue: the likelihood is accumulated at every measurement hit
    return launch_loglik(dX, N, d_data, iscale, allnan, M, nullptr, to_opts(opts), dLL, dYmeas, d_status, d_nsteps, s);
  }
  // many patients: solve once per sample, then score all (sample, patient) pairs in one pass
  if (!dYmeas) CK(bY.alloc(sizeof(double) * N * GYNC_NT_MEAS * GYNC_NMEAS, s));
  double* ym = dYmeas ? dYmeas : bY.as<double>();
  if (int rc = launch_loglik(dX, N, d_data, iscale, allnan, 0, nullptr, to_opts(opts), dLL, ym, d_status, d_nsteps, s)) return rc;
  gync_llh_epilogue_kernel<<<blocks_for(N, GYNC_EPI_SAMPLES), 256, 0, s>>>(ym, N, d_data, iscale, allnan, M, dLL);
  g_launches += 1;
  CK(cudaGetLastError());
  return GYNC_OK;
}

// Several devices in one process (gync_init_multi): contiguous blocks of samples per device, no
// data-path collective.  Column-major slices are moved with 2-D copies; all devices run concurrently.
static int loglik_batch_multi(const double* X, int64_t N, const double* data, int32_t M, const double* sigma_rho,
                              const gync_solver_opts* opts, double* LL, double* Ymeas, int32_t* status, int32_t* nsteps) {
  const int nd = (int)g_devs.size();
  // staging buffers come from the grow-only pool of each device slot (as in the single-device path): no cudaMalloc /
  // cudaFree — both synchronise the device — around a repeated call of the same size
  struct Part { int64_t lo = 0, n = 0; double *dX = nullptr, *dD = nullptr, *dS = nullptr, *dLL = nullptr, *dY = nullptr;
                int32_t *dSt = nullptr, *dNs = nullptr; };
  std::vector<Part> parts(nd);
  for (int i = 0; i < nd; ++i) {
    Part& p = parts[i];
    p.lo = N * i / nd;
    p.n = N * (i + 1) / nd - p.lo;
    if (p.n == 0) continue;
    if (int rc = use_device(i)) return rc;
    StagePool& sp = g_stage[i];
    p.dX = (double*)sp.get(0, sizeof(double) * p.n * GYNC_NX, g_device);
    p.dD = (double*)sp.get(1, sizeof(double) * 124 * M, g_device);
    p.dS = (double*)sp.get(2, sizeof(double) * M, g_device);
    p.dLL = (double*)sp.get(3, sizeof(double) * p.n * M, g_device);
    p.dY = Ymeas ? (double*)sp.get(4, sizeof(double) * p.n * GYNC_NT_MEAS * GYNC_NMEAS, g_device) : nullptr;
    p.dSt = (int32_t*)sp.get(5, sizeof(int32_t) * p.n, g_device);
    p.dNs = nsteps ? (int32_t*)sp.get(6, sizeof(int32_t) * p.n * 2, g_device) : nullptr;
    if (!p.dX || !p.dD || !p.dS || !p.dLL || !p.dSt || (Ymeas && !p.dY) || (nsteps && !p.dNs))
      return fail(GYNC_ERR_NOMEM, "gync_loglik_batch: out of device memory");
    CK(cudaMemcpy2DAsync(p.dX, sizeof(double) * p.n, X + p.lo, sizeof(double) * N, sizeof(double) * p.n, GYNC_NX,
                         cudaMemcpyHostToDevice, g_stream));
    CK(cudaMemcpyAsync(p.dD, data, sizeof(double) * 124 * M, cudaMemcpyHostToDevice, g_stream));
    CK(cudaMemcpyAsync(p.dS, sigma_rho, sizeof(double) * M, cudaMemcpyHostToDevice, g_stream));
    if (int rc = gync_loglik_batch_dev(p.dX, p.n, p.dD, M, p.dS, opts, p.dLL, p.dY, p.dSt, p.dNs, g_stream)) return rc;
  }
  for (int i = 0; i < nd; ++i) {
    Part& p = parts[i];
    if (p.n == 0) continue;
    if (int rc = use_device(i)) return rc;
    CK(cudaMemcpy2DAsync(LL + p.lo, sizeof(double) * N, p.dLL, sizeof(double) * p.n, sizeof(double) * p.n, M,
                         cudaMemcpyDeviceToHost, g_stream));
    if (Ymeas) CK(cudaMemcpy2DAsync(Ymeas + p.lo, sizeof(double) * N, p.dY, sizeof(double) * p.n, sizeof(double) * p.n,
                                    GYNC_NT_MEAS * GYNC_NMEAS, cudaMemcpyDeviceToHost, g_stream));
    if (status) CK(cudaMemcpyAsync(status + p.lo, p.dSt, sizeof(int32_t) * p.n, cudaMemcpyDeviceToHost, g_stream));
    if (nsteps) CK(cudaMemcpy2DAsync(nsteps + p.lo, sizeof(int32_t) * N, p.dNs, sizeof(int32_t) * p.n, sizeof(int32_t) * p.n, 2,
                                     cudaMemcpyDeviceToHost, g_stream));
  }
  for (int i = 0; i < nd; ++i) {
    if (parts[i].n == 0) continue;
    if (int rc = use_device(i)) return rc;
    CK(wait_stream(g_stream));
  }
  return use_device(0);
}

int gync_loglik_batch(const double* X, int64_t N, const double* data, int32_t M, const double* sigma_rho,
                      const gync_solver_opts* opts, double* LL, double* Ymeas, int32_t* status,
                      int32_t* nsteps) {
  if (N < 0 || M < 1 || !LL || (N > 0 && (!X || !data || !sigma_rho)))
    return fail(GYNC_ERR_ARG, "gync_loglik_batch: bad arguments");
  if (int rc = ensure_init()) return rc;
  if (N == 0) return GYNC_OK;
  if (g_devs.size() > 1 && N >= (int64_t)g_devs.size() * 256)
    return loglik_batch_multi(X, N, data, M, sigma_rho, opts, LL, Ymeas, status, nsteps);
  int slot = 0;
  for (size_t i = 0; i < g_devs.size(); ++i) if (g_devs[i].device == g_device) slot = (int)i;
  StagePool& sp = g_stage[slot];
  double* dX = (double*)sp.get(0, sizeof(double) * N * GYNC_NX, g_device);
  double* dD = (double*)sp.get(1, sizeof(double) * 124 * M, g_device);
  double* dS = (double*)sp.get(2, sizeof(double) * M, g_device);
  double* dLL = (double*)sp.get(3, sizeof(double) * N * M, g_device);
  double* dY = Ymeas ? (double*)sp.get(4, sizeof(double) * N * GYNC_NT_MEAS * GYNC_NMEAS, g_device) : nullptr;
  int32_t* dSt = (int32_t*)sp.get(5, sizeof(int32_t) * N, g_device);
  int32_t* dNs = nsteps ? (int32_t*)sp.get(6, sizeof(int32_t) * N * 2, g_device) : nullptr;
  if (!dX || !dD || !dS || !dLL || !dSt || (Ymeas && !dY) || (nsteps && !dNs))
    return fail(GYNC_ERR_NOMEM, "gync_loglik_batch: out of device memory");
  // GYNC_TRACE_CALLS=1: host-side time stamps of this call and device-side event times on stderr (where does a slow call
  // spend its time: enqueueing, the copies, the kernels, or waking up?)
  const bool trace = getenv("GYNC_TRACE_CALLS") != nullptr;
  auto now = [] { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6; };
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  const double t0 = now();
  if (trace) { for (auto& e : ev) cudaEventCreate(&e); cudaEventRecord(ev[0], g_stream); }
  CK(cudaMemcpyAsync(dX, X, sizeof(double) * N * GYNC_NX, cudaMemcpyHostToDevice, g_stream));
  CK(cudaMemcpyAsync(dD, data, sizeof(double) * 124 * M, cudaMemcpyHostToDevice, g_stream));
  CK(cudaMemcpyAsync(dS, sigma_rho, sizeof(double) * M, cudaMemcpyHostToDevice, g_stream));
  const double t1 = now();
  if (trace) cudaEventRecord(ev[1], g_stream);
  int rc = gync_loglik_batch_dev(dX, N, dD, M, dS, opts, dLL, dY, dSt, dNs, g_stream);
  if (rc) return rc;
  const double t2 = now();
  if (trace) cudaEventRecord(ev[2], g_stream);
  CK(cudaMemcpyAsync(LL, dLL, sizeof(double) * N * M, cudaMemcpyDeviceToHost, g_stream));
  if (Ymeas) CK(cudaMemcpyAsync(Ymeas, dY, sizeof(double) * N * GYNC_NT_MEAS * GYNC_NMEAS, cudaMemcpyDeviceToHost, g_stream));
  if (status) CK(cudaMemcpyAsync(status, dSt, sizeof(int32_t) * N, cudaMemcpyDeviceToHost, g_stream));
  if (nsteps) CK(cudaMemcpyAsync(nsteps, dNs, sizeof(int32_t) * N * 2, cudaMemcpyDeviceToHost, g_stream));
  const double t3 = now();
  if (trace) cudaEventRecord(ev[3], g_stream);
  CK(wait_stream(g_stream));
  if (trace) {
    const double t4 = now();
    float d01 = 0, d12 = 0, d23 = 0;
    cudaEventElapsedTime(&d01, ev[0], ev[1]); cudaEventElapsedTime(&d12, ev[1], ev[2]); cudaEventElapsedTime(&d23, ev[2], ev[3]);
    fprintf(stderr, "[gync trace] loglik_batch N=%lld host ms: h2d-enqueue %.2f launch-enqueue %.2f d2h-enqueue %.2f wait %.2f total %.2f | "
                    "device ms: h2d %.2f kernels %.2f d2h %.2f\n", (long long)N, t1 - t0, t2 - t1, t3 - t2, t4 - t3, t4 - t0, d01, d12, d23);
    for (auto& e : ev) cudaEventDestroy(e);
  }
  return GYNC_OK;
}

static int solve_common(const double* X, const double* Y0, const double* P, int64_t N, const double* t, int32_t nT,
                        const gync_solver_opts* opts, double* Y, int32_t* status) {
  if (N < 0 || nT < 1 || !t || !Y || (N > 0 && !X && (!Y0 || !P)))
    return fail(GYNC_ERR_ARG, "gync_solve: bad arguments");
  for (int k = 1; k < nT; ++k)
    if (t[k] < t[k - 1]) return fail(GYNC_ERR_ARG, "gync_solve: time grid must be sorted ascending");
  if (int rc = ensure_init()) return rc;
  if (N == 0) return GYNC_OK;
  DBuf dX, dY0, dP, dT, dY, dSt;
  if (X) {
    CK(dX.alloc(sizeof(double) * N * GYNC_NX));
    CK(cudaMemcpyAsync(dX.p, X, sizeof(double) * N * GYNC_NX, cudaMemcpyHostToDevice, g_stream));
  } else {
    CK(dY0.alloc(sizeof(double) * N * GYNC_NY));
    CK(dP.alloc(sizeof(double) * N * GYNC_NP));
    CK(cudaMemcpyAsync(dY0.p, Y0, sizeof(double) * N * GYNC_NY, cudaMemcpyHostToDevice, g_stream));
    CK(cudaMemcpyAsync(dP.p, P, sizeof(double) * N * GYNC_NP, cudaMemcpyHostToDevice, g_stream));
  }
  CK(dT.alloc(sizeof(double) * nT));
  CK(dY.alloc(sizeof(double) * N * nT * GYNC_NY));
  CK(dSt.alloc(sizeof(int32_t) * N));
  CK(cudaMemcpyAsync(dT.p, t, sizeof(double) * nT, cudaMemcpyHostToDevice, g_stream));
  if (int rc = launch_solve_grid(X ? dX.as<double>() : nullptr, dY0.as<double>(), dP.as<double>(), N, dT.as<double>(), nT, to_opts(opts),
                                 dY.as<double>(), dSt.as<int32_t>(), g_stream)) return rc;
  CK(cudaMemcpyAsync(Y, dY.p, sizeof(double) * N * nT * GYNC_NY, cudaMemcpyDeviceToHost, g_stream));
  if (status) CK(cudaMemcpyAsync(status, dSt.p, sizeof(int32_t) * N, cudaMemcpyDeviceToHost, g_stream));
  CK(wait_stream(g_stream));
  return GYNC_OK;
}

int gync_solve(const double* Y0, const double* P, int64_t N, const double* t, int32_t nT,
               const gync_solver_opts* opts, double* Y, int32_t* status) {
  return solve_common(nullptr, Y0, P, N, t, nT, opts, Y, status);
}

int gync_forwardsol(const double* X, int64_t N, const double* t, int32_t nT, const gync_solver_opts* opts,
                    double* Y, int32_t* status) {
  if (N > 0 && !X) return fail(GYNC_ERR_ARG, "gync_forwardsol: X is NULL");
  return solve_common(X, nullptr, nullptr, N, t, nT, opts, Y, status);
}

// ------------------------------------------------------------------------------------------ K4
int gync_em_iterate_generic(double* dw, const double* dL, int64_t K, int M, int niter, double* d_hist, cudaStream_t s);

int gync_em_iterate_dev(double* dw, const double* dL, int64_t K, int32_t M, int32_t niter, double* d_history,
                        void* stream) {
  if (K < 0 || M < 1 || niter < 0 || (K > 0 && (!dw || !dL))) return fail(GYNC_ERR_ARG, "gync_em_iterate: bad arguments");
  if (int rc = ensure_init()) return rc;
  if (K == 0 || niter == 0) return GYNC_OK;
  cudaStream_t s = (cudaStream_t)stream;   // NULL = the legacy default stream, as torch uses
  if (M > 64) return gync_em_iterate_generic(dw, dL, K, M, niter, d_history, s);
#define CALL(MH, F) em_persistent_launch<MH, F>(dw, dL, K, M, niter, d_history, s)
  EM_DISPATCH(M, CALL);
#undef CALL
}

int gync_em_norms_dev(const double* dw, const double* dL, int64_t K, int32_t M, double* d_norms_out, void* stream) {
  if (K < 0 || M < 1 || !d_norms_out) return fail(GYNC_ERR_ARG, "gync_em_norms: bad arguments");
  if (int rc = ensure_init()) return rc;
  cudaStream_t s = (cudaStream_t)stream;   // NULL = the legacy default stream, as torch uses
#define CALL(MH, F) em_norms_launch<MH, F>(dw, dL, K, M, d_norms_out, s)
  EM_DISPATCH(M, CALL);
#undef CALL
}

int gync_em_step_dev(double* dw, const double* dL, int64_t K, int32_t M, const double* d_norms_in,
                     double* d_norms_out, void* stream) {
  if (K < 0 || M < 1 || !d_norms_in || !d_norms_out) return fail(GYNC_ERR_ARG, "gync_em_step: bad arguments");
  if (int rc = ensure_init()) return rc;
  cudaStream_t s = (cudaStream_t)stream;   // NULL = the legacy default stream, as torch uses
#define CALL(MH, F) em_step_launch<MH, F>(dw, dL, K, M, d_norms_in, d_norms_out, s)
  EM_DISPATCH(M, CALL);
#undef CALL
}

extern "C" int gync_em_iterate_multi(double* w, const double* L, int64_t K, int32_t M, int32_t niter);

int gync_em_iterate(double* w, const double* L, int64_t K, int32_t M, int32_t niter, double* history) {
  if (K < 0 || M < 1 || niter < 0 || (K > 0 && (!w || !L))) return fail(GYNC_ERR_ARG, "gync_em_iterate: bad arguments");
  if (int rc = ensure_init()) return rc;
  if (K == 0 || niter == 0) return GYNC_OK;
  if (g_devs.size() > 1 && !history && K >= (int64_t)g_devs.size() * 4096 && M <= 64)
    return gync_em_iterate_multi(w, L, K, M, niter);
  DBuf dw, dL, dH;
  CK(dw.alloc(sizeof(double) * K));
  CK(dL.alloc(sizeof(double) * K * M));
  if (history) CK(dH.alloc(sizeof(double) * K * niter));
  CK(cudaMemcpyAsync(dw.p, w, sizeof(double) * K, cudaMemcpyHostToDevice, g_stream));
  CK(cudaMemcpyAsync(dL.p, L, sizeof(double) * K * M, cudaMemcpyHostToDevice, g_stream));
  int rc = gync_em_iterate_dev(dw.as<double>(), dL.as<double>(), K, M, niter, history ? dH.as<double>() : nullptr, g_stream);
  if (rc) return rc;
  CK(cudaMemcpyAsync(w, dw.p, sizeof(double) * K, cudaMemcpyDeviceToHost, g_stream));
  if (history) CK(cudaMemcpyAsync(history, dH.p, sizeof(double) * K * niter, cudaMemcpyDeviceToHost, g_stream));
  CK(wait_stream(g_stream));
  {
    int slot = 0;
    for (size_t i = 0; i < g_devs.size(); ++i) if (g_devs[i].device == g_device) slot = (int)i;
    if (g_em_local[slot].device == g_device)
      if (int rc2 = g_em_local[slot].check_error(g_stream, "gync_em_iterate")) return rc2;
  }
  return GYNC_OK;
}

// ------------------------------------------------------------------------------------------ misc
int gync_fp64_peak(double* tflops) {
  if (!tflops) return fail(GYNC_ERR_ARG, "gync_fp64_peak: NULL");
  if (int rc = ensure_init()) return rc;
  const int threads = 256, blocks = g_num_sms * 8, iters = 1 << 16;
  DBuf out;
  CK(out.alloc(sizeof(double) * threads * blocks));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  gync_fp64_peak_kernel<<<blocks, threads, 0, g_stream>>>(out.as<double>(), 1024);
  double best = 0.0;
  for (int rep = 0; rep < 3; ++rep) {
    CK(cudaEventRecord(e0, g_stream));
    gync_fp64_peak_kernel<<<blocks, threads, 0, g_stream>>>(out.as<double>(), iters);
    CK(cudaEventRecord(e1, g_stream));
    CK(cudaEventSynchronize(e1));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    const double fl = 2.0 * 8.0 * (double)iters * threads * blocks;
    best = std::max(best, fl / (ms * 1e-3) * 1e-12);
  }
  g_launches += 4;
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  *tflops = best;
  return GYNC_OK;
}

}  // extern "C"

#include "gync_abi_more.inl"
#include "gync_abi_eb.inl"
