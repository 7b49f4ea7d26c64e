// sm_100a kernels of the GynC.jl hot path.  See DESIGN.md for the layout and roofline of each.
#pragma once
#include <cuda_runtime.h>
#include <cooperative_groups.h>
#include "gync_llh.cuh"

namespace cg = cooperative_groups;

__constant__ double        c_refallparms[GYNC_NP] = {GYNC_REFALLPARMS_LIST};
__constant__ unsigned char c_sampledinds[GYNC_NS] = {GYNC_SAMPLEDINDS0_LIST};
__constant__ double        c_measmag[GYNC_NMEAS]  = {GYNC_MEASMAG_LIST};

#define GYNC_NT_MEAS 62   /* 2 periods x 31 days (model.jl:126,128) */

// ------------------------------------------------------------------------------------------
// Patient tables -> per-(patient,hormone) inverse scale and the all-missing flag
// (llh_measerror, model.jl:158-163: scale = model_measmagnitude * sigma_rho; all-NaN -> -Inf)
__global__ void gync_prep_patients_kernel(const double* __restrict__ data, const double* __restrict__ sigma_rho,
                                          int M, double* __restrict__ iscale, int* __restrict__ allnan) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  int any = 0;
  for (int i = 0; i < GYNC_NDAYS * GYNC_NMEAS; ++i) {
    const double d = data[(size_t)m * GYNC_NDAYS * GYNC_NMEAS + i];
    any |= (d == d);
  }
  allnan[m] = !any;
  for (int h = 0; h < GYNC_NMEAS; ++h) iscale[GYNC_NMEAS * m + h] = 1.0 / (c_measmag[h] * sigma_rho[m]);
}

// ------------------------------------------------------------------------------------------
// Per-thread solver state, B200 layout.
//
// One sample per thread needs ~520 doubles of live state per step (derived parameters q[93], the
// LU factors A[165], five RODAS stage vectors, y/y_new/work vector) — four times what a thread
// can hold in registers, and local-memory spills of that size miss L1 and serialise on L2
// latency (measured: 14 k solves/s).  So the big arrays live in SHARED memory, interleaved
// [element][thread] (a warp touches 32 consecutive doubles: conflict-free), and the small hot
// vectors (y, y_new, work) in registers.  The straight-line generated code indexes every array
// with compile-time constants, so each access is one LDS/STS with an immediate offset.
// GyncWorkSm keeps everything in shared memory (64 samples per CTA: 216.6 KB of the 227 KB) and is
// used by the generic time-grid kernel; the likelihood kernel uses the TMEM layout further down.
#ifndef GYNC_SOLVE_BLOCK
#define GYNC_SOLVE_BLOCK 48
#endif

// Element proxy: reads are VOLATILE shared loads.  Without that the compiler forwards the values
// it stored (it can prove the addresses) and, being out of registers, spills them to LOCAL
// memory — i.e. right back to the L1/L2 path this layout exists to avoid.
struct GyncSmRef {
  double* p;
  __device__ __forceinline__ operator double() const { return *(volatile double*)p; }
  __device__ __forceinline__ const GyncSmRef& operator=(double v) const { *p = v; return *this; }
  __device__ __forceinline__ const GyncSmRef& operator+=(double v) const { *p = *(volatile double*)p + v; return *this; }
  __device__ __forceinline__ const GyncSmRef& operator-=(double v) const { *p = *(volatile double*)p - v; return *this; }
};

template <int BLOCK>
struct GyncSmView {
  double* p;
  __device__ __forceinline__ GyncSmRef operator[](int i) const { return GyncSmRef{p + i * BLOCK}; }
};

// BLOCK = sample slots per CTA (one thread each).
template <int BLOCK>
struct GyncWorkSm {
  GyncSmView<BLOCK> q, k, A;
  double y[GYNC_NY], yn[GYNC_NY], e[GYNC_NY];
  static constexpr int kDoublesPerThread = GYNC_NQ + 6 * GYNC_NY + GYNC_NLU;
  static constexpr size_t kSmemBytes = sizeof(double) * kDoublesPerThread * BLOCK;
  __device__ __forceinline__ void bind(double* sm, int slot) {
    double* b = sm + slot;
    q.p = b;  b += GYNC_NQ * BLOCK;
    k.p = b;  b += 6 * GYNC_NY * BLOCK;
    A.p = b;
  }
};

// ------------------------------------------------------------------------------------------
// K1: forward solve + fused Gaussian log-likelihood epilogue.
// Persistent lanes: each thread pulls sample indices from a global counter and walks its sample
// through the merged measurement grid one step attempt per loop trip, so lanes of a warp stay
// busy although their samples need different numbers of (adaptive) steps.
struct GyncLlhSink {
  const double* __restrict__ data;     // 31 x 4 x M  (offset to the first patient this sample is scored against)
  const double* __restrict__ iscale;   // 4 x M
  int M;                               // number of patients scored (1 when every sample has its own patient)
  int64_t N;
  double* acc;                         // LL + n   (stride N over patients)
  double* ymeas;                       // Ymeas + n (N x 62 x 4) or nullptr
  bool enabled = true;                 // mirror lanes of the TMEM kernel compute but do not write
  double r0 = 0.0;                     // M == 1 (the bench shape, and every Metropolis sample): the sum stays in a register
                                       // instead of a global read-modify-write per measurement hit
  __device__ __forceinline__ void operator()(int period, int day, const double* y) {
    if (!enabled) return;
    const double ym[GYNC_NMEAS] = {y[1], y[6], y[23], y[24]};   // measuredinds (model.jl:1)
    if (ymeas) {
#pragma unroll
      for (int h = 0; h < GYNC_NMEAS; ++h) ymeas[N * ((period * GYNC_NDAYS + day) + GYNC_NT_MEAS * h)] = ym[h];
    }
    for (int m = 0; m < M; ++m) {
      double a = 0.0;
#pragma unroll
      for (int h = 0; h < GYNC_NMEAS; ++h) {
        const double d = __ldg(data + day + GYNC_NDAYS * h + GYNC_NDAYS * GYNC_NMEAS * m);
        if (d == d) {
          const double r = (d - ym[h]) * __ldg(iscale + GYNC_NMEAS * m + h);
          a = fma(r, r, a);
        }
      }
      if (M == 1) r0 += a; else acc[N * m] += a;
    }
  }
};

// ------------------------------------------------------------------------------------------
// K1, TMEM variant: the five RODAS stage vectors live in TENSOR MEMORY.
//
// The solver is bound by per-thread latency and by how many samples an SM can hold, and shared
// memory caps that at 64 (3.4 KB each).  Blackwell's 256 KB of tensor memory is per-thread
// addressable scratch in the 32x32b access shape (thread i of warp w <-> TMEM lane 32(w%4)+i,
// 512 columns x 32 bit each = 256 doubles per thread), with a 12-cycle load latency.  Moving
// k1..k5 (330 columns) plus a 66-column scratch vector there leaves q + A = 2064 B per sample in
// shared memory; with the 32 LU entries that the factorisation never touches (written once by the
// Jacobian code, read once per triangular solve) also in TMEM it is 1808 B: 128 samples per SM as
// 4 warps x 32 lanes — one full warp per SM sub-partition (396 + 64 = 460 of the 512 columns used).
// tcgen05.ld/st are warp-collective (.sync.aligned), so every lane of a warp executes the step
// (lanes without a sample run on their stale state).
#define GYNC_TM_NSLOT 128
#define GYNC_TM_LPW 32
#define GYNC_TM_COL_K 0
#define GYNC_TM_COL_V (5 * 2 * GYNC_NY)
#define GYNC_TM_COL_TAIL (6 * 2 * GYNC_NY)   /* the 32 LU entries the factorisation never touches (GYNC_NLU_BODY..) */

// If GYNC_TM_LPW < 32 the upper lanes of each warp have no shared-memory slot of their own (an earlier
// layout ran 28 samples per warp).  They then MIRROR the top real lanes: same sample index, same inputs,
// same instruction stream, hence bit-identical values — so they may issue the same shared-memory stores
// to the same addresses (benign same-value races) and no store needs a predicate (predicated stores
// turned into branches that chopped the straight-line LU / Jacobian code into small basic blocks and
// caused spills).  Only real lanes touch global memory.  With GYNC_TM_LPW == 32 every lane is real.
template <int NSLOT>
struct GyncSmViewP {
  double* p;
  __device__ __forceinline__ GyncSmRef operator[](int i) const { return GyncSmRef{p + i * NSLOT}; }
};

// Read-only view for phases that only READ an array (q in the RHS, A in the triangular solves):
// plain `ld.shared.f64` through non-volatile asm on a laundered address.  The compiler cannot
// forward earlier stores into them (the address is opaque after the barrier in gync_sm_ro) but
// ptxas may schedule them freely inside the phase; volatile loads stay pinned in program order.
// The view must be re-created inside every loop iteration that uses it, otherwise the loads are
// loop-invariant and get hoisted out of the loop wholesale (258 live doubles -> spills).
template <int NSLOT>
struct GyncSmViewRO {
  uint32_t sa;
  __device__ __forceinline__ double operator[](int i) const {
    double v;
    asm("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(sa + (uint32_t)(i * NSLOT * 8)));
    return v;
  }
};
template <int NSLOT>
__device__ __forceinline__ GyncSmViewRO<NSLOT> gync_sm_ro(const double* p) {
  uint32_t sa = (uint32_t)__cvta_generic_to_shared(p);
  asm volatile("" : "+r"(sa) :: "memory");       // earlier stores are emitted before; provenance forgotten
  return GyncSmViewRO<NSLOT>{sa};
}

template <int B>
__device__ __forceinline__ GyncSmViewRO<B> gync_ro(const GyncSmView<B>& v) { return gync_sm_ro<B>(v.p); }

// LU storage of the TMEM layout: body [0, GYNC_NLU_BODY) in shared memory, tail in tensor memory.
template <int NSLOT>
struct GyncSmViewT {
  double* p;
  uint32_t tm;     // TMEM address of this warp's lane quadrant, column 0
  __device__ __forceinline__ GyncSmRef operator[](int i) const { return GyncSmRef{p + i * NSLOT}; }
};
template <int NSLOT>
struct GyncSmViewROT {
  uint32_t sa, tm;
  __device__ __forceinline__ double operator[](int i) const {
    double v;
    asm("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(sa + (uint32_t)(i * NSLOT * 8)));
    return v;
  }
};

#define TM_R16(r) "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), \
                  "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
#define TM_W16(r) "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), \
                  "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
#define TM_I16(r) "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), \
                  "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])

// 8 doubles (16 columns) TMEM -> registers; the wait carries the registers as in/out operands so
// no consumer can be scheduled ahead of it.
__device__ __forceinline__ void gync_tm_ld8(uint32_t ta, double (&v)[8]) {
  uint32_t r[16];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : TM_R16(r) : "r"(ta));
  asm volatile("tcgen05.wait::ld.sync.aligned;" : TM_W16(r) :: "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __hiloint2double((int)r[2 * i + 1], (int)r[2 * i]);
}
__device__ __forceinline__ double gync_tm_ld1(uint32_t ta) {
  uint32_t lo, hi;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0,%1}, [%2];" : "=r"(lo), "=r"(hi) : "r"(ta));
  asm volatile("tcgen05.wait::ld.sync.aligned;" : "+r"(lo), "+r"(hi) :: "memory");
  return __hiloint2double((int)hi, (int)lo);
}
__device__ __forceinline__ void gync_tm_st8(uint32_t ta, const double* v) {
  uint32_t r[16];
#pragma unroll
  for (int i = 0; i < 8; ++i) { r[2 * i] = (uint32_t)__double2loint(v[i]); r[2 * i + 1] = (uint32_t)__double2hiint(v[i]); }
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%16], {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15};"
               :: TM_I16(r), "r"(ta) : "memory");
}
__device__ __forceinline__ void gync_tm_st1(uint32_t ta, double v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%2], {%0,%1};"
               :: "r"((uint32_t)__double2loint(v)), "r"((uint32_t)__double2hiint(v)), "r"(ta) : "memory");
}
__device__ __forceinline__ void gync_tm_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// store a 33-vector to TMEM columns [col, col+66)
// One x2 store per double: a double already sits in an aligned register pair, so nothing has to be marshalled.
// (x16 stores need 16 consecutive registers; the allocator keeps the vectors elsewhere and ptxas emitted 16 moves in
// front of every x16 store — 640 IMAD.MOV per RODAS step, seen in the ncu source page.  Removing them changed the
// time by +0.5 %: the kernel is bound by dependency latency, not by issue slots.)
__device__ __forceinline__ void gync_tm_store_vec(uint32_t ta, const double* v) {
#pragma unroll
  for (int i = 0; i < GYNC_NY; ++i) gync_tm_st1(ta + 2 * i, v[i]);
  gync_tm_wait_st();
}

// tail chunk c (8 doubles = 16 columns) of the LU storage: overloads of the generated code's hooks
template <int NSLOT>
__device__ __forceinline__ void gync_tail_store8(GyncSmViewT<NSLOT>& A, const int c, const double v0, const double v1,
                                                 const double v2, const double v3, const double v4, const double v5,
                                                 const double v6, const double v7) {
  const double v[8] = {v0, v1, v2, v3, v4, v5, v6, v7};
#pragma unroll
  for (int i = 0; i < 8; ++i) gync_tm_st1(A.tm + GYNC_TM_COL_TAIL + 16 * c + 2 * i, v[i]);
  gync_tm_wait_st();     // chunks are emitted in the order their values become available, not by index
}
template <int NSLOT>
__device__ __forceinline__ void gync_tail_load8(const GyncSmViewROT<NSLOT>& A, const int c, double (&v)[8]) {
  gync_tm_ld8(A.tm + GYNC_TM_COL_TAIL + 16 * c, v);
}
template <int NSLOT>
__device__ __forceinline__ void gync_tail_load8(const GyncSmViewT<NSLOT>& A, const int c, double (&v)[8]) {   // during the LU
  gync_tm_ld8(A.tm + GYNC_TM_COL_TAIL + 16 * c, v);
}
template <int NSLOT>
__device__ __forceinline__ GyncSmViewROT<NSLOT> gync_ro(const GyncSmViewT<NSLOT>& A) {
  return GyncSmViewROT<NSLOT>{gync_sm_ro<NSLOT>(A.p).sa, A.tm};
}

struct GyncWorkTm {
  GyncSmViewP<GYNC_TM_NSLOT> q;
  GyncSmViewT<GYNC_TM_NSLOT> A;
  double y[GYNC_NY], yn[GYNC_NY], e[GYNC_NY];
  uint32_t tm;   // TMEM address of this warp's lane quadrant, column 0
  static constexpr size_t kSmemBytes = sizeof(double) * (GYNC_NQ + GYNC_NLU_BODY) * GYNC_TM_NSLOT;
};

// RODAS4 step with k in TMEM (same arithmetic as gync_rodas4_step up to the order of the stage
// sums).  Each stage reads every earlier k once, producing the stage point (registers) and the
// correction sum_j c_sj k_j, which is parked in the TMEM scratch vector while the RHS runs.
__device__ __forceinline__ double gync_rodas4_step_tm(GyncWorkTm& W, const double h, const double rtol, const double atol) {
  const double ih = GYNC_RCP(h);
  GYNC_TICK_INIT
  gync_rhs_w(W.y, gync_sm_ro<GYNC_TM_NSLOT>(W.q.p), ih * (1.0 / RD_G), W.e, W.A);
  GYNC_TICK(0)
  gync_lu(W.A);
  GYNC_TICK(1)
  gync_lusolve(gync_ro(W.A), W.e);
  GYNC_TICK(2)
#pragma unroll 1
  for (int s = 0; s < 5; ++s) {
    // park k_{s+1} for the later stages (k5 is consumed right here and never read again); its own contribution to
    // this stage comes from the registers it still sits in — a third of the stage sums' tensor-memory loads
    if (s < 4) gync_tm_store_vec(W.tm + GYNC_TM_COL_K + 2 * GYNC_NY * s, W.e);
    {
      const double a = RD_TA[s][s], c = RD_TC[s][s];
#pragma unroll
      for (int i = 0; i < GYNC_NY; ++i) { W.yn[i] = fma(a, W.e[i], W.y[i]); W.e[i] = c * W.e[i]; }
    }
#pragma unroll 1
    for (int j = 0; j < s; ++j) {
      const double a = RD_TA[s][j], c = RD_TC[s][j];
      const uint32_t ta = W.tm + GYNC_TM_COL_K + 2 * GYNC_NY * j;
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) {
        double k[8];
        gync_tm_ld8(ta + 16 * ch, k);
#pragma unroll
        for (int i = 0; i < 8; ++i) { W.yn[8 * ch + i] = fma(a, k[i], W.yn[8 * ch + i]); W.e[8 * ch + i] = fma(c, k[i], W.e[8 * ch + i]); }
      }
      const double k32 = gync_tm_ld1(ta + 64);
      W.yn[32] = fma(a, k32, W.yn[32]);
      W.e[32] = fma(c, k32, W.e[32]);
    }
    gync_tm_store_vec(W.tm + GYNC_TM_COL_V, W.e);                          // park the correction
    GYNC_TICK(3)
    gync_rhs(W.yn, gync_sm_ro<GYNC_TM_NSLOT>(W.q.p), W.e);
    GYNC_TICK(4)
#pragma unroll
    for (int ch = 0; ch < 4; ++ch) {
      double v[8];
      gync_tm_ld8(W.tm + GYNC_TM_COL_V + 16 * ch, v);
#pragma unroll
      for (int i = 0; i < 8; ++i) W.e[8 * ch + i] = fma(ih, v[i], W.e[8 * ch + i]);
    }
    W.e[32] = fma(ih, gync_tm_ld1(W.tm + GYNC_TM_COL_V + 64), W.e[32]);
    GYNC_TICK(5)
    gync_lusolve(gync_ro(W.A), W.e);
    GYNC_TICK(2)
  }
  const double err = gync_finish_step(W, rtol, atol);
  GYNC_TICK(6)
  return err;
}

// Rodas5P step with the stage storage in TMEM: same arithmetic as gync_rodas5p_step (gync_solver.cuh explains why five
// stored vectors are enough) up to the order of the sums, and the same code shape as the RODAS4 form above — one rolled
// stage loop whose body is the proven RHS + triangular solve + two short tensor-memory sums.
__device__ __forceinline__ double gync_rodas5p_step_tm(GyncWorkTm& W, const double h, const double rtol, const double atol) {
  const double ih = GYNC_RCP(h);
  double pc[2];                      // the model's real power at the step's start and 1/base (gync_pow0689, gync_solver.cuh)
  GYNC_TICK_INIT
  gync_rhs_w(W.y, GyncPowRec<GyncSmViewRO<GYNC_TM_NSLOT>>{gync_sm_ro<GYNC_TM_NSLOT>(W.q.p), pc}, ih * (1.0 / R5_G), W.e, W.A);
  GYNC_TICK(0)
  gync_lu(W.A);
  GYNC_TICK(1)
  gync_lusolve(gync_ro(W.A), W.e);
  GYNC_TICK(2)
#pragma unroll 1
  for (int s = 0; s < 7; ++s) {
    if (s < R5_NSTORE) {
      // park k_{s+1} for the later stages; its own contribution to this stage comes from the registers it still sits in
      gync_tm_store_vec(W.tm + GYNC_TM_COL_K + 2 * GYNC_NY * s, W.e);
    } else if (s == R5_NSTORE) {
      // absorb k6 into the stored k4 and k5 (see gync_solver.cuh)
#pragma unroll 1
      for (int j = 3; j < 5; ++j) {
        const double f = (j == 3) ? R5_BETA : R5_ALPHA;
        const uint32_t ta = W.tm + GYNC_TM_COL_K + 2 * GYNC_NY * j;
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          double k[8];
          gync_tm_ld8(ta + 16 * ch, k);
#pragma unroll
          for (int i = 0; i < 8; ++i) gync_tm_st1(ta + 16 * ch + 2 * i, fma(f, W.e[8 * ch + i], k[i]));
        }
        gync_tm_st1(ta + 64, fma(f, W.e[32], gync_tm_ld1(ta + 64)));
      }
      gync_tm_wait_st();
    }
    {
      const double a = R5_OA[s], c = R5_OC[s];
#pragma unroll
      for (int i = 0; i < GYNC_NY; ++i) { W.yn[i] = fma(a, W.e[i], W.y[i]); W.e[i] = c * W.e[i]; }
    }
    const int ns = s < R5_NSTORE ? s : R5_NSTORE;
#pragma unroll 1
    for (int j = 0; j < ns; ++j) {
      const double a = R5_TA[s][j], c = R5_TC[s][j];
      const uint32_t ta = W.tm + GYNC_TM_COL_K + 2 * GYNC_NY * j;
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) {
        double k[8];
        gync_tm_ld8(ta + 16 * ch, k);
#pragma unroll
        for (int i = 0; i < 8; ++i) { W.yn[8 * ch + i] = fma(a, k[i], W.yn[8 * ch + i]); W.e[8 * ch + i] = fma(c, k[i], W.e[8 * ch + i]); }
      }
      const double k32 = gync_tm_ld1(ta + 64);
      W.yn[32] = fma(a, k32, W.yn[32]);
      W.e[32] = fma(c, k32, W.e[32]);
    }
    gync_tm_store_vec(W.tm + GYNC_TM_COL_V, W.e);                          // park the correction
    GYNC_TICK(3)
    gync_rhs(W.yn, GyncPowUse<GyncSmViewRO<GYNC_TM_NSLOT>>{gync_sm_ro<GYNC_TM_NSLOT>(W.q.p), pc}, W.e);
    GYNC_TICK(4)
#pragma unroll
    for (int ch = 0; ch < 4; ++ch) {
      double v[8];
      gync_tm_ld8(W.tm + GYNC_TM_COL_V + 16 * ch, v);
#pragma unroll
      for (int i = 0; i < 8; ++i) W.e[8 * ch + i] = fma(ih, v[i], W.e[8 * ch + i]);
    }
    W.e[32] = fma(ih, gync_tm_ld1(W.tm + GYNC_TM_COL_V + 64), W.e[32]);
    GYNC_TICK(5)
    gync_lusolve(gync_ro(W.A), W.e);
    GYNC_TICK(2)
  }
  const double err = gync_finish_step(W, rtol, atol);
  GYNC_TICK(6)
  return err;
}

__device__ __forceinline__ double gync_ros_step_tm(GyncWorkTm& W, const double h, const double rtol, const double atol) {
#if GYNC_ROS_ORDER == 5
  return gync_rodas5p_step_tm(W, h, rtol, atol);
#else
  return gync_rodas4_step_tm(W, h, rtol, atol);
#endif
}

__global__ void __launch_bounds__(128, 1)
gync_loglik_tm_kernel(const double* __restrict__ X, int64_t N, const double* __restrict__ data,
                      const double* __restrict__ iscale, const int* __restrict__ allnan, int M,
                      const int* __restrict__ pat /* nullable: patient of each sample -> LL is N x 1 */,
                      GyncSolverOpts o, double* __restrict__ LL, double* __restrict__ Ymeas,
                      int* __restrict__ status, int* __restrict__ nsteps, unsigned long long* __restrict__ counter,
                      const int* __restrict__ order /* nullable: queue position -> sample (gync_sched.cuh) */) {
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

  const bool real = lane < GYNC_TM_LPW;                    // lanes >= LPW (none at LPW = 32) mirror the top real lanes
  const int src_lane = real ? lane : lane - (32 - GYNC_TM_LPW);
  GyncWorkTm W;
  const int slot = warp * GYNC_TM_LPW + src_lane;
  W.q.p = gync_sm + slot;
  W.A.p = gync_sm + (size_t)GYNC_NQ * GYNC_TM_NSLOT + slot;
  W.tm = tmem_base + ((uint32_t)(warp * 32) << 16);
  W.A.tm = W.tm;
#pragma unroll
  for (int i = 0; i < GYNC_NY; ++i) { W.y[i] = 1.0; W.yn[i] = 1.0; W.e[i] = 0.0; }

  const double big = 1.0e300, ninf = -INFINITY;
  bool active = false, done = false;
  int64_t n = 0;
  double T = 0.0, t = 0.0, tn = 0.0;
  int i0 = 0, i1 = 0, budget = 0, st = GYNC_ST_OK, mlo = 0, mcnt = M;
  GyncCtrl c;
  gync_ctrl_init(c, 1e-3);
  GyncLlhSink sink = {data, iscale, M, N, nullptr, nullptr, real};
  for (;;) {
    // fetch: real lanes draw an index, mirror lanes copy it (collectives stay outside divergent code)
    const bool want = !active && !done;
    long long drawn = -1;
    if (want && real) drawn = (long long)atomicAdd(counter, 1ull);
    drawn = __shfl_sync(0xffffffffu, drawn, src_lane);
    if (want) {
      n = (int64_t)drawn;
      if (n >= N) {
        done = true;
      } else {
        if (order) n = order[n];
        gync_load_sample(W, [&](int k) { return __ldg(X + n + N * k); }, c_refallparms, c_sampledinds, &T);
        mlo = pat ? pat[n] : 0;
        mcnt = pat ? 1 : M;
        sink.data = data + (size_t)GYNC_NDAYS * GYNC_NMEAS * mlo;
        sink.iscale = iscale + GYNC_NMEAS * mlo;
        sink.M = mcnt;
        if (real && mcnt > 1) for (int m = 0; m < mcnt; ++m) LL[n + N * m] = 0.0;
        sink.r0 = 0.0;
        sink.acc = LL + n;
        sink.ymeas = Ymeas ? Ymeas + n : nullptr;
        i0 = i1 = 0;
        budget = o.max_steps;
        t = fmin(1.0, 1.0 + T);
        tn = t;
        st = (gync_all_finite(W.y, GYNC_NY) && gync_all_finite(W.q, GYNC_NQ) && (T * 0.0 == 0.0)) ? GYNC_ST_OK
                                                                                                 : GYNC_ST_NONFINITE;
        gync_ctrl_init(c, (st == GYNC_ST_OK) ? ((o.h0 > 0.0) ? o.h0 : gync_initial_step(W, o.rtol, o.atol)) : 1.0);
        active = true;
        // serve the measurement times equal to the start time
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
    __syncwarp();
    if (__all_sync(0xffffffffu, done)) break;
    // ---- one step attempt, executed by ALL lanes of the warp (TMEM ops are warp-collective)
    const bool stepping = active && st == GYNC_ST_OK && tn > t;
    double h = c.h;
    if (o.hmax > 0.0) h = fmin(h, o.hmax);
    const double rem = tn - t;
    const bool last = stepping && (h * 1.0001 >= rem);
    if (last) h = rem;
    if (!stepping || !(h > 0.0)) h = 1.0;                  // dummy lanes: any finite positive step
    const double hprop = c.h;
    const double err = gync_ros_step_tm(W, h, o.rtol, o.atol);
    __syncwarp();
    if (stepping) {
      if (budget-- <= 0) {
        st = GYNC_ST_MAXSTEPS;
      } else if (!(h > 1e-14 * fmax(1.0, fabs(t)))) {
        st = GYNC_ST_HMIN;
      } else {
        if (gync_ctrl_update(c, h, err, true)) {         // a non-finite candidate arrives as err = NaN
          t = last ? tn : t + h;
#pragma unroll
          for (int i = 0; i < GYNC_NY; ++i) W.y[i] = W.yn[i];
          if (last && h < hprop) c.h = fmax(c.h, hprop);
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
        if (real) {
          for (int m = 0; m < mcnt; ++m) {
            const double a = (mcnt == 1) ? sink.r0 : LL[n + N * m];
            LL[n + N * m] = (st != GYNC_ST_OK || allnan[mlo + m]) ? ninf : -0.5 * a;
          }
          if (st != GYNC_ST_OK && Ymeas)
            for (int k = 0; k < GYNC_NT_MEAS * GYNC_NMEAS; ++k) Ymeas[n + N * k] = NAN;
          if (status) status[n] = st;
          if (nsteps) { nsteps[n] = c.nacc; nsteps[n + N] = c.nrej; }
        }
        active = false;
        // keep the idle state benign for the lock-step iterations that may follow
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

// Many-patient epilogue (M > GYNC_FUSED_MAX_M): the solver kernel only records the 62 x 4 measured
// values per sample; this kernel scores every sample against every patient table.  In the fused
// form each measurement hit costs M global read-modify-writes by ONE lane while the other 31 lanes
// of the warp wait (measured: +62 % kernel time at M = 40); here the cost of the solve no longer
// depends on M and the likelihood matrix is one coalesced pass: a CTA stages the measured
// trajectories of 16 samples in shared memory and its threads walk (sample, patient) pairs.
#define GYNC_FUSED_MAX_M 4
#define GYNC_EPI_SAMPLES 16
__global__ void __launch_bounds__(256)
gync_llh_epilogue_kernel(const double* __restrict__ Ymeas /* N x 62 x 4 */, int64_t N, const double* __restrict__ data,
                         const double* __restrict__ iscale, const int* __restrict__ allnan, int M,
                         double* __restrict__ LL /* N x M */) {
  __shared__ double sy[GYNC_NT_MEAS * GYNC_NMEAS][GYNC_EPI_SAMPLES + 1];
  __shared__ int sbad[GYNC_EPI_SAMPLES];
  const int64_t n0 = (int64_t)blockIdx.x * GYNC_EPI_SAMPLES;
  if (threadIdx.x < GYNC_EPI_SAMPLES) sbad[threadIdx.x] = 0;
  __syncthreads();
  for (int idx = threadIdx.x; idx < GYNC_NT_MEAS * GYNC_NMEAS * GYNC_EPI_SAMPLES; idx += blockDim.x) {
    const int s = idx % GYNC_EPI_SAMPLES, k = idx / GYNC_EPI_SAMPLES;      // consecutive threads: consecutive samples
    const double v = (n0 + s < N) ? Ymeas[n0 + s + N * k] : 0.0;
    sy[k][s] = v;
    if (v != v) sbad[s] = 1;                                               // NaN trajectory -> -Inf (model.jl:138-141)
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < GYNC_EPI_SAMPLES * M; idx += blockDim.x) {
    const int s = idx % GYNC_EPI_SAMPLES, m = idx / GYNC_EPI_SAMPLES;
    if (n0 + s >= N) continue;
    double acc = 0.0;
    const double* dm = data + (size_t)GYNC_NDAYS * GYNC_NMEAS * m;
    for (int q = 0; q < 2; ++q)
      for (int h = 0; h < GYNC_NMEAS; ++h) {
        const double isc = iscale[GYNC_NMEAS * m + h];
        for (int d = 0; d < GYNC_NDAYS; ++d) {
          const double dv = __ldg(dm + d + GYNC_NDAYS * h);
          if (dv == dv) {
            const double r = (dv - sy[(q * GYNC_NDAYS + d) + GYNC_NT_MEAS * h][s]) * isc;
            acc = fma(r, r, acc);
          }
        }
      }
    LL[n0 + s + N * m] = (sbad[s] || allnan[m]) ? -INFINITY : -0.5 * acc;
  }
}

// gync(y0,p,t) / forwardsol(x,t): full 33-species trajectory on a shared sorted time grid.
struct GyncGridSink {
  double* Y;        // Y + n ; layout N x nT x 33
  int64_t N;
  int nT;
  __device__ __forceinline__ void operator()(int k, const double* y) {
#pragma unroll
    for (int i = 0; i < GYNC_NY; ++i) Y[N * (k + (int64_t)nT * i)] = y[i];
  }
};

template <int BLOCK>
__global__ void __launch_bounds__(BLOCK, 1)
gync_solve_kernel(const double* __restrict__ X, const double* __restrict__ Y0, const double* __restrict__ P,
                  int64_t N, const double* __restrict__ tgrid, int nT, GyncSolverOpts o,
                  double* __restrict__ Y, int* __restrict__ status) {
  extern __shared__ double gync_sm[];
  GyncWorkSm<BLOCK> W;
  W.bind(gync_sm, threadIdx.x);
  for (int64_t n = (int64_t)blockIdx.x * BLOCK + threadIdx.x; n < N; n += (int64_t)gridDim.x * BLOCK) {
    if (X) {
      double T;
      gync_load_sample(W, [&](int k) { return __ldg(X + n + N * k); }, c_refallparms, c_sampledinds, &T);
    } else {
      double p[GYNC_NP];
#pragma unroll
      for (int k = 0; k < GYNC_NP; ++k) p[k] = __ldg(P + n + N * k);
      gync_pack_params(p, W.q);
#pragma unroll
      for (int i = 0; i < GYNC_NY; ++i) W.y[i] = __ldg(Y0 + n + N * i);
    }
    GyncGridSink sink = {Y + n, N, nT};
    const int s = gync_run_time_grid(W, tgrid, nT, o, sink, (GyncStepStats*)nullptr);
    if (s != GYNC_ST_OK) {
      for (int64_t k = 0; k < (int64_t)nT * GYNC_NY; ++k) Y[n + N * k] = NAN;   // gyncycle.jl:18-19
    }
    if (status) status[n] = s;
  }
}

// gync(y0,p,t) / forwardsol(x,t) in the tensor-memory layout of K1 (128 samples per SM, persistent lanes): the same step
// function, a per-lane cursor into the shared sorted time grid instead of the merged measurement sequences.  Row 0 of
// the output is y0 at t[0]; repeated grid times are served from the same state (CVODE's behaviour for a repeated tout).
__global__ void __launch_bounds__(128, 1)
gync_solve_tm_kernel(const double* __restrict__ X, const double* __restrict__ Y0, const double* __restrict__ P,
                     int64_t N, const double* __restrict__ tgrid, int nT, GyncSolverOpts o,
                     double* __restrict__ Y, int* __restrict__ status, unsigned long long* __restrict__ counter,
                     const int* __restrict__ order /* nullable, gync_sched.cuh */, int* __restrict__ nsteps /* nullable, N x 2 */) {
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
  bool active = false, done = false;
  int64_t n = 0;
  double t = 0.0, tn = 0.0;
  int kidx = 0, budget = 0, st = GYNC_ST_OK;
  GyncCtrl c;
  gync_ctrl_init(c, 1e-3);
  GyncGridSink sink = {Y, N, nT};
  for (;;) {
    if (!active && !done) {
      n = (int64_t)atomicAdd(counter, 1ull);
      if (n >= N) {
        done = true;
      } else {
        if (order) n = order[n];
        if (X) {
          double T;
          gync_load_sample(W, [&](int k) { return __ldg(X + n + N * k); }, c_refallparms, c_sampledinds, &T);
        } else {
          double p[GYNC_NP];
#pragma unroll
          for (int k = 0; k < GYNC_NP; ++k) p[k] = __ldg(P + n + N * k);
          gync_pack_params(p, W.q);
#pragma unroll
          for (int i = 0; i < GYNC_NY; ++i) W.y[i] = __ldg(Y0 + n + N * i);
        }
        sink.Y = Y + n;
        kidx = 0;
        budget = o.max_steps;
        t = tgrid[0];
        tn = t;
        st = (gync_all_finite(W.y, GYNC_NY) && gync_all_finite(W.q, GYNC_NQ) && t == t) ? GYNC_ST_OK : GYNC_ST_NONFINITE;
        gync_ctrl_init(c, (st == GYNC_ST_OK) ? ((o.h0 > 0.0) ? o.h0 : gync_initial_step(W, o.rtol, o.atol)) : 1.0);
        active = true;
        if (st == GYNC_ST_OK) {
          while (kidx < nT && (tn = tgrid[kidx]) <= t) { sink(kidx, W.y); ++kidx; }
          if (kidx < nT && !(tn == tn)) st = GYNC_ST_NONFINITE;
        }
      }
    }
    __syncwarp();
    if (__all_sync(0xffffffffu, done)) break;
    // ---- one step attempt, executed by ALL lanes of the warp (TMEM ops are warp-collective)
    const bool stepping = active && st == GYNC_ST_OK && kidx < nT && tn > t;
    double h = c.h;
    if (o.hmax > 0.0) h = fmin(h, o.hmax);
    const double rem = tn - t;
    const bool last = stepping && (h * 1.0001 >= rem);
    if (last) h = rem;
    if (!stepping || !(h > 0.0)) h = 1.0;                  // idle lanes: any finite positive step
    const double hprop = c.h;
    const double err = gync_ros_step_tm(W, h, o.rtol, o.atol);
    __syncwarp();
    if (stepping) {
      if (budget-- <= 0) {
        st = GYNC_ST_MAXSTEPS;
      } else if (!(h > 1e-14 * fmax(1.0, fabs(t)))) {
        st = GYNC_ST_HMIN;
      } else {
        if (gync_ctrl_update(c, h, err, true)) {         // a non-finite candidate arrives as err = NaN
          t = last ? tn : t + h;
#pragma unroll
          for (int i = 0; i < GYNC_NY; ++i) W.y[i] = W.yn[i];
          if (last && h < hprop) c.h = fmax(c.h, hprop);
        }
      }
    }
    if (active) {
      bool finished = (st != GYNC_ST_OK);
      if (!finished && t >= tn) {
        while (kidx < nT && (tn = tgrid[kidx]) <= t) { sink(kidx, W.y); ++kidx; }
        if (kidx < nT && !(tn == tn)) { st = GYNC_ST_NONFINITE; finished = true; }
      }
      if (kidx >= nT) finished = true;
      if (finished) {
        if (st != GYNC_ST_OK)
          for (int64_t k = 0; k < (int64_t)nT * GYNC_NY; ++k) Y[n + N * k] = NAN;   // gyncycle.jl:18-19
        if (status) status[n] = st;
        if (nsteps) { nsteps[n] = c.nacc; nsteps[n + N] = c.nrej; }
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

// ------------------------------------------------------------------------------------------
// K4: EM / self-consistency iteration  w_k <- w_k/M * sum_m L_km / (L'w)_m   (em.jl:27-41)
//
// Fused single-pass form: L is read ONCE per iteration.  While row k's new weight is produced,
// the NEXT iteration's denominators sum_k L_km w_k' are accumulated in registers.  HBM-bound:
// 8KM + 16K + 16M algorithmic bytes per iteration.
//   * two threads per row, each owning half of the columns (MH = MT/2 values of L in flight per
//     thread, 512 threads per SM: ~80 KB of loads in flight per SM to cover HBM latency);
//     lane = 16*half + r, so for one column a half-warp reads 16 consecutive rows = one 128 B line;
//   * the row sum needs one xor-16 shuffle; the per-thread accumulators are reduced over the 16
//     row-lanes by shuffles, over the 16 warps through shared memory (fixed order);
//   * block partials go to a double-buffered global scratch; after ONE grid-wide barrier every
//     block re-reduces them in a fixed order (deterministic, no atomics), 12 blocks per thread
//     group in parallel.  All `niter` iterations run inside one cooperative launch.
#define GYNC_EM_THREADS 512
#define GYNC_EM_WARPS (GYNC_EM_THREADS / 32)

template <int MH>
struct EmSmem {
  double red[GYNC_EM_WARPS * 2 * MH];
  double rinv[2 * MH];
  double grp[(GYNC_EM_THREADS / (2 * MH) + 1) * 2 * MH];
  double part[8 * 2 * MH];       // multi-GPU: every rank's partial denominators as read from the mailbox
  int last;                      // multi-GPU: this block was the last to arrive and does the send
};

// acc[j] (column half*MH + j) summed over the block -> out[m], m < M
template <int MH>
__device__ __forceinline__ void em_block_reduce(double (&acc)[MH], EmSmem<MH>& sm, double* out, int M) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, half = lane >> 4;
#pragma unroll
  for (int j = 0; j < MH; ++j) {
    double v = acc[j];
#pragma unroll
    for (int off = 8; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    if ((lane & 15) == 0) sm.red[warp * 2 * MH + half * MH + j] = v;
  }
  __syncthreads();
  if (threadIdx.x < M) {
    double v = 0.0;
#pragma unroll
    for (int w = 0; w < GYNC_EM_WARPS; ++w) v += sm.red[w * 2 * MH + threadIdx.x];
    out[threadIdx.x] = v;
  }
  __syncthreads();
}

// out[m] = sum_b partials[b][m] in a fixed order, using all threads of the block
template <int MH>
__device__ __forceinline__ void em_gather_partials(const double* __restrict__ partials, int nb, int M, EmSmem<MH>& sm,
                                                   double* out_sm /* [2*MH] in shared memory */, bool reciprocal) {
  constexpr int MT = 2 * MH, NG = GYNC_EM_THREADS / MT;
  const int m = threadIdx.x % MT, g = threadIdx.x / MT;
  if (g < NG) {
    // up to 16 independent loads in flight per thread (fixed summation order)
    double v[16];
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      const int b = g + u * NG;
      v[u] = (b < nb) ? __ldcg(partials + (size_t)b * MT + m) : 0.0;
    }
    double acc = ((v[0] + v[1]) + (v[2] + v[3])) + ((v[4] + v[5]) + (v[6] + v[7]));
    acc += ((v[8] + v[9]) + (v[10] + v[11])) + ((v[12] + v[13]) + (v[14] + v[15]));
    for (int b = g + 16 * NG; b < nb; b += NG) acc += __ldcg(partials + (size_t)b * MT + m);
    sm.grp[g * MT + m] = acc;
  }
  __syncthreads();
  if (threadIdx.x < MT) {
    double v = 0.0;
    for (int gg = 0; gg < NG; ++gg) v += sm.grp[gg * MT + threadIdx.x];
    out_sm[threadIdx.x] = (threadIdx.x < M) ? (reciprocal ? 1.0 / v : v) : 0.0;   // unused columns: exact zeros
  }
  __syncthreads();
}

// partial denominators of this block's rows: out[m] = sum_k L[k,m] w[k]
template <int MH, bool FULL>
__device__ __forceinline__ void em_norms_pass(const double* __restrict__ w, const double* __restrict__ L,
                                              int64_t K, int M, EmSmem<MH>& sm, double* out) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, half = lane >> 4, r = lane & 15;
  double acc[MH];
#pragma unroll
  for (int j = 0; j < MH; ++j) acc[j] = 0.0;
  const int64_t stride = (int64_t)gridDim.x * GYNC_EM_WARPS * 16;
  for (int64_t k = ((int64_t)blockIdx.x * GYNC_EM_WARPS + warp) * 16 + r; k < K; k += stride) {
    const double wk = w[k];
    const double* p = L + k + K * (int64_t)(half * MH);
#pragma unroll
    for (int j = 0; j < MH; ++j) {
      if (FULL || half * MH + j < M) acc[j] = fma(__ldg(p), wk, acc[j]);
      p += K;
    }
  }
  em_block_reduce<MH>(acc, sm, out, M);
}

// one EM step over this block's rows given sm.rinv[m] = 1/norms[m]; accumulates next partials
template <int MH, bool FULL>
__device__ __forceinline__ void em_step_pass(double* __restrict__ w, const double* __restrict__ L, int64_t K, int M,
                                             EmSmem<MH>& sm, double* out, double* hist) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, half = lane >> 4, r = lane & 15;
  double acc[MH];
#pragma unroll
  for (int j = 0; j < MH; ++j) acc[j] = 0.0;
  const double* ri = sm.rinv + half * MH;      // reciprocal denominators (shared memory, broadcast reads)
  const double invM = 1.0 / (double)M;
  const int64_t stride = (int64_t)gridDim.x * GYNC_EM_WARPS * 16;
  const int64_t k0 = ((int64_t)blockIdx.x * GYNC_EM_WARPS + warp) * 16;
  for (int64_t kb = k0; kb < K; kb += stride) {          // warp-uniform loop (the shuffle needs all lanes)
    const int64_t k = kb + r;
    const bool valid = k < K;
    double l[MH];
    const double* p = L + (valid ? k : 0) + K * (int64_t)(half * MH);
#pragma unroll
    for (int j = 0; j < MH; ++j) {
      l[j] = (valid && (FULL || half * MH + j < M)) ? __ldg(p) : 0.0;
      p += K;
    }
    const double wk = valid ? w[k] : 0.0;
    double s0 = 0.0, s1 = 0.0;
#pragma unroll
    for (int j = 0; j + 1 < MH; j += 2) { s0 = fma(l[j], ri[j], s0); s1 = fma(l[j + 1], ri[j + 1], s1); }
    if (MH & 1) s0 = fma(l[MH - 1], ri[MH - 1], s0);
    double s = s0 + s1;
    s += __shfl_xor_sync(0xffffffffu, s, 16);
    const double wn = wk * invM * s;
    if (valid && half == 0) {
      w[k] = wn;
      if (hist) hist[k] = wn;
    }
#pragma unroll
    for (int j = 0; j < MH; ++j) acc[j] = fma(l[j], wn, acc[j]);
  }
  em_block_reduce<MH>(acc, sm, out, M);
}

// ------------------------------------------------------------------------------------------
// K4 across GPUs as ONE kernel: compute step + the allreduce of the M denominators fused over
// NVLink peer memory.  Every rank runs this persistent cooperative kernel on its row shard.  Per
// iteration: local fused pass -> grid barrier -> block 0 sums the block partials and STORES the M
// partial denominators straight into every peer's mailbox (cudaIpc-mapped peer pointers, i.e.
// st.global over NVLink/NVSwitch), publishes a sequence flag, spins on the flags of all peers,
// adds the `world` contributions in rank order (bit-identical on every rank) -> grid barrier ->
// next pass.  No NCCL call, no kernel boundary, no host round trip inside the iteration loop.
#define GYNC_EM_SPIN_LIMIT (4000000000ll)      /* ~2 s of SM clock */
struct GyncEmPeer {
  double* mail[8];                 // mail[p]: rank p's mailbox [2 slots][world][MT][2] 64-bit words (peer pointer; mail[rank] local)
  unsigned long long* flag[8];     // (round-1 flag words; unused by the self-validating exchange)
  int world, rank;
};

// Self-validating mailbox words (the "LL" idea of NCCL's low-latency protocol, for doubles): a double travels as two
// naturally aligned 8-byte words, each (sequence number << 32 | 32 bits of the value).  An 8-byte store is single-copy
// atomic — over NVLink too — so the receiver needs neither a separate flag nor a system-wide fence: it polls a word until
// its upper half is the sequence number it is waiting for, and then holds valid data.
__device__ __forceinline__ void gync_em_ll_store(double* mail, size_t idx, double v, unsigned long long seq) {
  volatile unsigned long long* p = reinterpret_cast<volatile unsigned long long*>(mail) + 2 * idx;
  const unsigned long long tag = (seq & 0xFFFFFFFFull) << 32;
  p[0] = tag | (unsigned long long)(unsigned)__double2loint(v);
  p[1] = tag | (unsigned long long)(unsigned)__double2hiint(v);
}
__device__ __forceinline__ bool gync_em_ll_load(const double* mail, size_t idx, unsigned long long seq, int* err_flag, double* out) {
  const volatile unsigned long long* p = reinterpret_cast<const volatile unsigned long long*>(mail) + 2 * idx;
  const unsigned long long want = seq & 0xFFFFFFFFull;
  const long long t0 = clock64();
  int polls = 0;
  unsigned long long a, b;
  for (;;) {
    a = p[0];
    b = p[1];
    if ((a >> 32) == want && (b >> 32) == want) break;
    if ((++polls & 255) == 0 && (*(volatile int*)err_flag || clock64() - t0 > GYNC_EM_SPIN_LIMIT)) {
      *(volatile int*)err_flag = 1;
      __threadfence_system();
      *out = 0.0;
      return false;
    }
  }
  *out = __hiloint2double((int)(unsigned)(b & 0xFFFFFFFFull), (int)(unsigned)(a & 0xFFFFFFFFull));
  return true;
}

// ------------------------------------------------------------------------------------------
// K4, unified persistent kernel (single GPU and fused multi-GPU), flag-based synchronisation.
//
// A grid-wide barrier is more than this loop needs: only ONE block has to see all partials, and the
// other blocks only have to learn the new denominators.  So per iteration every block does its pass,
// writes its partials and ARRIVES (one atomic); block 0 waits for the arrivals, sums the partials in a
// fixed order, exchanges with the peer GPUs if there are any (mailboxes over NVLink, rank order), and
// publishes 1/denominators plus a sequence flag; all blocks poll that flag.  Compared with two
// cooperative-groups grid barriers plus 148 redundant gathers this removes ~5 us per iteration.
// RESIDENT: this GPU's rows of L are loaded into shared memory once (when they fit).
// Bounded spin on a monotonically increasing word.  Gives up — and raises *err_flag for everybody — after ~2 s of SM
// clock, or as soon as somebody else has raised the flag: a rank that failed to launch, a peer that died or a mismatched
// iteration count must not leave the other blocks / ranks spinning for (iterations x timeout).
__device__ __forceinline__ bool gync_em_spin(const volatile unsigned long long* word, unsigned long long want, int* err_flag) {
  const long long t0 = clock64();
  int polls = 0;
  while (*word < want) {
    if ((++polls & 255) == 0 && (*(volatile int*)err_flag || clock64() - t0 > GYNC_EM_SPIN_LIMIT)) {
      *(volatile int*)err_flag = 1;
      __threadfence_system();
      return false;
    }
  }
  return true;
}

struct GyncEmSync {
  unsigned long long* arrive;   // monotonically increasing arrival counter (device memory)
  unsigned long long* ready;    // sequence number of the denominators currently published in gnorm
  unsigned long long seq0;      // first sequence number of this launch (ready < seq0 on entry)
  unsigned long long arrive0;   // value of *arrive when this launch starts
};

// -DGYNC_EM_PHASES builds only (tools/em_phase_breakdown.py): SM-clock ticks per phase, summed over blocks and iterations.
// [0] local pass  [1] arrival -> all contributions read (what a block waits)  [2] gather by the last-arriving block
// [3] its stores into the mailboxes  [4] rank-order sum + reciprocals  [5] block-iterations counted  [6] sends counted
#ifdef GYNC_EM_PHASES
__device__ unsigned long long gync_em_phase[8];
#  define GYNC_EP(...) __VA_ARGS__
#else
#  define GYNC_EP(...)
#endif

// MULTI: the exchange with peer GPUs is compiled in.  A separate instantiation, because the exchange code in the same
// body cost the single-GPU streaming pass 9 % (60.3 vs 54.9 us per iteration at K = 1e6: two registers over, a worse
// schedule of the pass's loads) although it never executed there.
template <int MH, bool FULL, int RES /* 0 streaming, 1 all rows resident, 2 partly resident */, bool MULTI>
__global__ void __launch_bounds__(GYNC_EM_THREADS, 1)
gync_em_kernel(double* __restrict__ w, const double* __restrict__ L, int64_t K, int M, int niter, int rows_per_block,
               int rows_resident /* RESIDENT: rows of each block held in shared memory (<= rows_per_block; the rest streams) */,
               double* __restrict__ partials /* 2 x gridDim.x x 2MH */, double* __restrict__ gnorm /* 2 x 2MH */,
               GyncEmSync sync, GyncEmPeer peer, int* __restrict__ err_flag, double* __restrict__ history) {
  constexpr bool RESIDENT = RES != 0;
  extern __shared__ double em_dyn[];            // RESIDENT: [M][rows_per_block] then EmSmem ; else EmSmem only
  constexpr int MT = 2 * MH;
  // RESIDENT: a block owns rows [row0, row0 + rows_per_block); the first `rows_resident` of them live in shared memory for
  // the whole launch, the rest (none when everything fits) is streamed from L2 / HBM in every iteration.  The hybrid
  // exists for the 8-GPU shard of cfg 5 (125 000 rows x 40 = 40 MB per GPU against 148 x ~215 KB = 31 MB of shared
  // memory): 80 % of the pass then runs at shared-memory speed.
  const int R = RES == 2 ? rows_resident : rows_per_block;
  double* sL = em_dyn;
  EmSmem<MH>& sm = *reinterpret_cast<EmSmem<MH>*>(em_dyn + (RESIDENT ? (size_t)M * R : 0));
  const int nb = gridDim.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, half = lane >> 4, rl = lane & 15;
  const int64_t row0 = (int64_t)blockIdx.x * rows_per_block;
  const int nrows_all = RESIDENT ? (int)max((int64_t)0, min((int64_t)rows_per_block, K - row0)) : 0;
  const int nrows = RES == 2 ? min(nrows_all, R) : nrows_all;
  if (RESIDENT) {
    for (int m = 0; m < M; ++m)
      for (int r = threadIdx.x; r < nrows; r += GYNC_EM_THREADS) sL[(size_t)m * R + r] = __ldg(L + row0 + r + K * (int64_t)m);
    __syncthreads();
  }
  const double invM = 1.0 / (double)M;
  // RESIDENT: rows_per_block <= kSweeps * 256 (the host checks it); each thread keeps its rows' weights in registers
  constexpr int kSweeps = 3;
  double wreg[kSweeps];
#pragma unroll
  for (int sw = 0; sw < kSweeps; ++sw) {
    const int r = warp * 16 + sw * GYNC_EM_WARPS * 16 + rl;
    wreg[sw] = (RESIDENT && r < nrows) ? w[row0 + r] : 0.0;
  }
  for (int it = -1; it < niter; ++it) {
    double* pbuf = partials + (size_t)((it + 1) & 1) * nb * MT;      // double-buffered block partials
    GYNC_EP(const long long ep0 = clock64();)
    // ---- local pass (it == -1: first denominators only)
    if (RESIDENT) {
      double acc[MH];
#pragma unroll
      for (int j = 0; j < MH; ++j) acc[j] = 0.0;
      const double* ri = sm.rinv + half * MH;
#pragma unroll
      for (int sw = 0; sw < kSweeps; ++sw) {      // a thread owns the same <= kSweeps rows in every iteration: w stays in registers
        const int rb = warp * 16 + sw * GYNC_EM_WARPS * 16;
        if (rb < nrows) {                          // warp-uniform
          const int r = rb + rl;
          const bool valid = r < nrows;
          const double* p = sL + (size_t)(half * MH) * R + (valid ? r : 0);
          double l[MH];
#pragma unroll
          for (int j = 0; j < MH; ++j) l[j] = (valid && (FULL || half * MH + j < M)) ? p[(size_t)j * R] : 0.0;
          double wk = wreg[sw];
          if (it >= 0) {
            double s0 = 0.0, s1 = 0.0;
#pragma unroll
            for (int j = 0; j + 1 < MH; j += 2) { s0 = fma(l[j], ri[j], s0); s1 = fma(l[j + 1], ri[j + 1], s1); }
            if (MH & 1) s0 = fma(l[MH - 1], ri[MH - 1], s0);
            double s = s0 + s1;
            s += __shfl_xor_sync(0xffffffffu, s, 16);
            wk = wk * invM * s;
            wreg[sw] = wk;
            if (valid && half == 0) {
              w[row0 + r] = wk;
              if (history) history[(size_t)it * K + row0 + r] = wk;
            }
          }
#pragma unroll
          for (int j = 0; j < MH; ++j) acc[j] = fma(l[j], wk, acc[j]);
        }
      }
      // the rows of this block that did not fit: same arithmetic, operands from global memory.  (Its own instantiation:
      // compiled into the all-resident kernel, this never-entered loop cost that kernel 4 %, 6.95 vs 6.70 us at K = 1e5.)
      if constexpr (RES == 2)
      for (int rb = R + warp * 16; rb < nrows_all; rb += GYNC_EM_WARPS * 16) {      // warp-uniform bounds (the shuffle needs all lanes)
        const int r = rb + rl;
        const bool valid = r < nrows_all;
        const int64_t k = row0 + (valid ? r : 0);
        const double* p = L + k + K * (int64_t)(half * MH);
        double l[MH];
#pragma unroll
        for (int j = 0; j < MH; ++j) { l[j] = (valid && (FULL || half * MH + j < M)) ? __ldg(p) : 0.0; p += K; }
        double wk = valid ? w[k] : 0.0;
        if (it >= 0) {
          double s0 = 0.0, s1 = 0.0;
#pragma unroll
          for (int j = 0; j + 1 < MH; j += 2) { s0 = fma(l[j], ri[j], s0); s1 = fma(l[j + 1], ri[j + 1], s1); }
          if (MH & 1) s0 = fma(l[MH - 1], ri[MH - 1], s0);
          double s = s0 + s1;
          s += __shfl_xor_sync(0xffffffffu, s, 16);
          wk = wk * invM * s;
          if (valid && half == 0) {
            w[k] = wk;
            if (history) history[(size_t)it * K + k] = wk;
          }
        }
#pragma unroll
        for (int j = 0; j < MH; ++j) acc[j] = fma(l[j], wk, acc[j]);
      }
      em_block_reduce<MH>(acc, sm, pbuf + (size_t)blockIdx.x * MT, M);
    } else if (it < 0) {
      em_norms_pass<MH, FULL>(w, L, K, M, sm, pbuf + (size_t)blockIdx.x * MT);
    } else {
      em_step_pass<MH, FULL>(w, L, K, M, sm, pbuf + (size_t)blockIdx.x * MT, history ? history + (size_t)it * K : nullptr);
    }
    // ---- arrive
    const unsigned long long seq = sync.seq0 + (unsigned long long)(it + 1);
    const int slot = (int)(seq & 1ull);
    const unsigned long long want = sync.arrive0 + (unsigned long long)nb * (unsigned long long)(it + 2);
    GYNC_EP(const long long ep1 = clock64();)
    if (threadIdx.x == 0) {
      __threadfence();
      const unsigned long long prev = atomicAdd(sync.arrive, 1ull);
      if constexpr (MULTI) sm.last = (prev + 1ull == want);
    }
    if constexpr (!MULTI) {
      // single GPU: every block waits for all arrivals and gathers the partials itself (same fixed order
      // everywhere) — two L2 round trips shorter than publish-and-poll through block 0
      if (threadIdx.x == 0) {
        gync_em_spin(sync.arrive, want, err_flag);
        __threadfence();
      }
      if (__syncthreads_or(threadIdx.x == 0 && *(volatile int*)err_flag != 0)) break;
      em_gather_partials<MH>(partials + (size_t)((it + 1) & 1) * nb * MT, nb, M, sm, sm.rinv, true);
    } else {
    // several GPUs: ONE-SHOT exchange.  The block that arrives LAST (it knows: the arrival counter told it) sums the
    // block partials in the fixed order and stores this GPU's M partial denominators straight into the mailbox of every
    // rank — its own included — as self-validating words over NVLink.  EVERY block of every rank then polls its own GPU's
    // mailbox until all `world` contributions of this sequence number are there and adds them in rank order (bit-identical
    // everywhere).  No block waits for another block of its own GPU except through the data it needs; there is no
    // gather -> publish -> poll relay through block 0, no separate flag, no system-wide fence.  (Round 1: block 0 waited
    // for the arrivals, gathered, stored, fenced, raised flags, waited for the peers' flags, summed, published, and the
    // other 147 blocks polled the published copy: ~12 us of a 19 us iteration on 8 GPUs.)
    // Slot reuse is safe with two slots: a peer can only write sequence n+2 into this slot after it has seen every
    // rank's n+1, which this rank sends only once ALL its blocks have arrived for n+1, i.e. are done reading n.
    __syncthreads();
    if (sm.last) {
      GYNC_EP(const long long eg0 = clock64();)
      __threadfence();
      em_gather_partials<MH>(pbuf, nb, M, sm, sm.rinv, false);                // this GPU's denominators (fixed order)
      GYNC_EP(const long long eg1 = clock64();)
      for (int idx = threadIdx.x; idx < peer.world * MT; idx += GYNC_EM_THREADS) {
        const int p = idx / MT, m = idx % MT;
        gync_em_ll_store(peer.mail[p], ((size_t)slot * peer.world + peer.rank) * MT + m, sm.rinv[m], seq);
      }
      GYNC_EP(if (threadIdx.x == 0) { atomicAdd(&gync_em_phase[2], (unsigned long long)(eg1 - eg0));
                                      atomicAdd(&gync_em_phase[3], (unsigned long long)(clock64() - eg1)); atomicAdd(&gync_em_phase[6], 1ull); })
    }
    bool ok = true;
    for (int idx = threadIdx.x; idx < peer.world * MT; idx += GYNC_EM_THREADS)
      ok &= gync_em_ll_load(peer.mail[peer.rank], (size_t)slot * peer.world * MT + idx, seq, err_flag, &sm.part[idx]);
    // a timed-out wait anywhere (this GPU or, through the flag, a peer that gave up) ends the loop for every block
    if (__syncthreads_or(!ok || (threadIdx.x == 0 && *(volatile int*)err_flag != 0))) break;
    GYNC_EP(const long long ep2 = clock64();)
    if (threadIdx.x < MT) {
      double v = 0.0;
      for (int r = 0; r < peer.world; ++r) v += sm.part[r * MT + threadIdx.x];     // rank order: identical everywhere
      sm.rinv[threadIdx.x] = (threadIdx.x < M) ? 1.0 / v : 0.0;
    }
    __syncthreads();
    GYNC_EP(if (threadIdx.x == 0) { atomicAdd(&gync_em_phase[0], (unsigned long long)(ep1 - ep0)); atomicAdd(&gync_em_phase[1], (unsigned long long)(ep2 - ep1));
                                    atomicAdd(&gync_em_phase[4], (unsigned long long)(clock64() - ep2)); atomicAdd(&gync_em_phase[5], 1ull); })
    }
  }
}

// Sharded form (one process per GPU): block partials -> scratch, then a one-block finishing
// kernel sums them; the caller allreduces the M denominators across ranks.
template <int MH, bool FULL>
__global__ void __launch_bounds__(GYNC_EM_THREADS, 1)
gync_em_norms_kernel(const double* __restrict__ w, const double* __restrict__ L, int64_t K, int M,
                     double* __restrict__ partials) {
  __shared__ EmSmem<MH> sm;
  em_norms_pass<MH, FULL>(w, L, K, M, sm, partials + (size_t)blockIdx.x * 2 * MH);
}

template <int MH, bool FULL>
__global__ void __launch_bounds__(GYNC_EM_THREADS, 1)
gync_em_step_kernel(double* __restrict__ w, const double* __restrict__ L, int64_t K, int M,
                    const double* __restrict__ norms_in, double* __restrict__ partials) {
  __shared__ EmSmem<MH> sm;
  if (threadIdx.x < 2 * MH) sm.rinv[threadIdx.x] = (threadIdx.x < M) ? 1.0 / norms_in[threadIdx.x] : 0.0;
  __syncthreads();
  em_step_pass<MH, FULL>(w, L, K, M, sm, partials + (size_t)blockIdx.x * 2 * MH, nullptr);
}

template <int MH>
__global__ void __launch_bounds__(GYNC_EM_THREADS, 1)
gync_em_finish_kernel(const double* __restrict__ partials, int nb, int M, double* __restrict__ out) {
  __shared__ EmSmem<MH> sm;
  em_gather_partials<MH>(partials, nb, M, sm, sm.rinv, false);
  if (threadIdx.x < M) out[threadIdx.x] = sm.rinv[threadIdx.x];
}

// ------------------------------------------------------------------------------------------
// FP64 FMA peak microbenchmark: 8 independent DFMA chains per thread.
__global__ void gync_fp64_peak_kernel(double* out, int iters) {
  double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const double b = 1.0000001, c = 1e-9;
  for (int i = 0; i < iters; ++i) {
    a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
    a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
  }
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}
