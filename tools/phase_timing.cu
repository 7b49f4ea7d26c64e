// Debug tool: per-phase clock64 attribution of the Rosenbrock step the library is built with (thread 0 of block 0; Rodas5P by
// default, -DGYNC_ROS_ORDER=4 for RODAS4P).  slots: 0 rhs_w, 1 lu, 2 triangular solves, 3 stage sums (TMEM), 4 rhs, 5 e += v, 6 norm
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I gync.jl_b200/csrc tools/phase_timing.cu -o tools/phase_timing.bin
#define GYNC_PHASE_TIMING
#include <cstdio>
#include "gync_kernels.cuh"
int main() {
  const int N = 112 * 148;
  double *X, *LL, *D, *S; double* isc; int* an; unsigned long long* cnt;
  cudaMallocManaged(&X, sizeof(double) * N * 116); cudaMallocManaged(&LL, sizeof(double) * N);
  cudaMallocManaged(&D, sizeof(double) * 124); cudaMallocManaged(&S, 8); cudaMallocManaged(&isc, 32); cudaMallocManaged(&an, 4);
  cudaMallocManaged(&cnt, 8); *cnt = 0;
  static const double refp[114] = {GYNC_REFALLPARMS_LIST}; static const double refy[33] = {GYNC_REFY0_LIST};
  static const unsigned char si[82] = {GYNC_SAMPLEDINDS0_LIST};
  for (int n = 0; n < N; ++n) { for (int k = 0; k < 82; ++k) X[n + (size_t)N * k] = refp[si[k]] * (1.0 + 0.001 * (n % 7));
    for (int i = 0; i < 33; ++i) X[n + (size_t)N * (82 + i)] = refy[i]; X[n + (size_t)N * 115] = 28.0; }
  for (int i = 0; i < 124; ++i) D[i] = 1.0; S[0] = 0.1;
  gync_prep_patients_kernel<<<1, 64>>>(D, S, 1, isc, an);
  GyncSolverOpts o = {1e-6, 1e-6, 0, 0, 100000};
  cudaFuncSetAttribute(gync_loglik_tm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GyncWorkTm::kSmemBytes);
  gync_loglik_tm_kernel<<<148, 128, GyncWorkTm::kSmemBytes>>>(X, N, D, isc, an, 1, nullptr, o, LL, nullptr, nullptr, nullptr, cnt, nullptr);
  cudaError_t e = cudaDeviceSynchronize(); printf("%s LL0=%f\n", cudaGetErrorString(e), LL[0]);
  long long h[8]; cudaMemcpyFromSymbol(h, gync_phase_clk, sizeof(h));
  #if GYNC_ROS_ORDER == 5
  const char* nm[] = {"rhs_w (RHS + Jacobian + W)", "lu", "lusolve x8", "stage sums (TMEM) x7", "rhs x7", "e += v (TMEM) x7", "error norm", "outside step"};
#else
  const char* nm[] = {"rhs_w (RHS + Jacobian + W)", "lu", "lusolve x6", "stage sums (TMEM) x5", "rhs x5", "e += v (TMEM) x5", "error norm", "outside step"};
#endif
  long long tot = 0; for (int i = 0; i < 7; ++i) tot += h[i];
  for (int i = 0; i < 7; ++i) printf("%-40s %12lld  %.1f%%\n", nm[i], h[i], 100.0 * h[i] / tot);
  return 0;
}
