// FP64 microbenchmarks on one B200 (evidence for DESIGN.md §5: K1's latency bound, K6's "DMMA has no more throughput
// than the FMA pipe" claim).  Prints one JSON object.
//   (a) dependent DFMA chain: cycles per DFMA seen by one warp (the latency a straight-line solver step pays)
//   (b) DFMA issue rate of ONE warp per SM sub-partition at ILP 1..8, and of 2/4/8 warps per sub-partition
//   (c) DMMA (mma.sync.m8n8k4.f64) throughput at 1..8 warps per sub-partition, in FP64 flop/s, vs the DFMA peak
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 tools/fp64_microbench.cu -o tools/fp64_microbench.bin
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void dfma_kernel(double* out, int iters, double a, double b, long long* cyc) {
  double x[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) x[i] = threadIdx.x * 1e-3 + i;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 16; ++r)
#pragma unroll
      for (int i = 0; i < ILP; ++i) x[i] = fma(x[i], a, b);
  }
  const long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int ILP>
__global__ void dmma_kernel(double* out, int iters, long long* cyc) {
  double a = threadIdx.x * 1e-3, b = 1.0 - threadIdx.x * 1e-3;
  double c[ILP][2];
#pragma unroll
  for (int i = 0; i < ILP; ++i) { c[i][0] = i; c[i][1] = -i; }
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int i = 0; i < ILP; ++i)
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                     : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  }
  const long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <class F>
static float time_ms(F launch) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  launch();                       // warm-up
  cudaEventRecord(e0);
  launch();
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  return ms;
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const int sms = p.multiProcessorCount;
  double* out; long long* cyc;
  cudaMalloc(&out, sizeof(double) * sms * 1024); cudaMallocManaged(&cyc, 8);
  const int iters = 20000;
  printf("{\"device\": \"%s\", \"sms\": %d,\n", p.name, sms);
  // (a) latency: ILP 1, one warp
  dfma_kernel<1><<<1, 32>>>(out, iters, 1.0000001, 1e-9, cyc); cudaDeviceSynchronize();
  printf(" \"dfma_dependent_latency_cycles\": %.2f,\n", (double)*cyc / (16.0 * iters));
  // (b) one warp per sub-partition (128 threads/SM), ILP sweep; then more warps
  printf(" \"dfma_tflops\": {");
#define RUN_DFMA(ILP, THREADS) { float ms = time_ms([&] { dfma_kernel<ILP><<<sms, THREADS>>>(out, iters, 1.0000001, 1e-9, cyc); }); \
    printf("\"ilp%d_warps%d\": %.2f, ", ILP, THREADS / 128, 2.0 * 16 * ILP * iters * (double)THREADS * sms / (ms * 1e-3) * 1e-12); }
  RUN_DFMA(1, 128) RUN_DFMA(2, 128) RUN_DFMA(4, 128) RUN_DFMA(8, 128)
  RUN_DFMA(1, 256) RUN_DFMA(2, 256) RUN_DFMA(4, 256) RUN_DFMA(1, 512) RUN_DFMA(4, 512) RUN_DFMA(4, 1024)
  printf("\"unit\": \"TFLOP/s\"},\n");
  // (c) DMMA m8n8k4: 8*8*4*2 = 512 flop per warp instruction
  dmma_kernel<1><<<1, 32>>>(out, iters, cyc); cudaDeviceSynchronize();
  printf(" \"dmma_dependent_latency_cycles\": %.2f,\n", (double)*cyc / (8.0 * iters));
  printf(" \"dmma_tflops\": {");
#define RUN_DMMA(ILP, THREADS) { float ms = time_ms([&] { dmma_kernel<ILP><<<sms, THREADS>>>(out, iters, cyc); }); \
    printf("\"ilp%d_warps%d\": %.2f, ", ILP, THREADS / 128, 512.0 * 8 * ILP * iters * (double)(THREADS / 32) * sms / (ms * 1e-3) * 1e-12); }
  RUN_DMMA(1, 128) RUN_DMMA(4, 128) RUN_DMMA(8, 128) RUN_DMMA(4, 256) RUN_DMMA(4, 512) RUN_DMMA(8, 1024)
  printf("\"unit\": \"TFLOP/s\"}}\n");
  return 0;
}
