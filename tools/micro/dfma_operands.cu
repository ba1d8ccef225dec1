// Microbenchmark: FP64 issue rate vs number of distinct register source operands per instruction.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dfma_operands dfma_operands.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256) k(double* out, int iters, double m, double c) {
  double a[8], b[8], d[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    a[i] = threadIdx.x * 1e-9 + i;
    b[i] = 1.0 + 1e-7 * (i + 1) + 1e-12 * threadIdx.x;
    d[i] = 1e-9 * (i + 2) + 1e-13 * threadIdx.x;
  }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (MODE == 0) a[i] = fma(a[i], m, c);               // 1 new register operand (m, c reused)
        if (MODE == 1) a[i] = fma(b[i], d[i], a[i]);         // 3 distinct register operands
        if (MODE == 2) a[i] = fma(b[i], m, a[i]);            // 2 distinct + 1 shared
        if (MODE == 3) a[i] = a[i] + b[i];                   // DADD 2 distinct
        if (MODE == 4) a[i] = a[i] * b[i];                   // DMUL 2 distinct
        if (MODE == 5) a[i] = fma(b[i], d[(i + r) & 7], a[i]);  // 3 distinct, rotating
        if (MODE == 6) a[i] = fma(b[r], d[i], a[i]);         // b shared across the 8 consecutive FMAs
        if (MODE == 7) { double t = a[i] + b[i]; a[i] = b[i]; b[i] = t; }          // DADD, destination differs from both sources
        if (MODE == 8) { double t = fma(a[i], b[i], d[i]); d[i] = a[i]; a[i] = b[i]; b[i] = t; }  // DFMA, 3 distinct + distinct dest
        if (MODE == 9) { double t = a[i] * b[i]; a[i] = b[i]; b[i] = t; }          // DMUL distinct dest
      }
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name, int warps_per_sm) {
  int sms;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int threads = 256, blocks = sms * warps_per_sm * 32 / threads * 1, iters = 2048;
  double* buf;
  cudaMalloc(&buf, sizeof(double) * blocks * threads);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  k<MODE><<<blocks, threads>>>(buf, 16, 1.0000001, 1e-9);
  float best = 1e30f;
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0);
    k<MODE><<<blocks, threads>>>(buf, iters, 1.0000001, 1e-9);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  const double inst = 64.0 * iters * blocks * threads;
  printf("%-44s warps/SM=%2d  %.2f Ginstr/s  (%.2f TFLOP/s if FMA)  lanes/clk/SM@1.9GHz=%.1f\n", name, warps_per_sm,
         inst / best / 1e6, 2 * inst / best / 1e9, inst / best / 1e6 / 148 / 1.9);
  cudaFree(buf);
}

int main() {
  for (int w : {8, 32}) {
    run<0>("DFMA a=fma(a,m,c)      [1 new reg operand]", w);
    run<1>("DFMA a=fma(b,d,a)      [3 distinct]", w);
    run<2>("DFMA a=fma(b,m,a)      [2 distinct + shared]", w);
    run<3>("DADD a=a+b             [2 distinct]", w);
    run<4>("DMUL a=a*b             [2 distinct]", w);
    run<5>("DFMA a=fma(b,d[rot],a) [3 distinct]", w);
    run<6>("DFMA a=fma(b[r],d,a)   [b shared over 8]", w);
    run<7>("DADD t=a+b             [distinct dest]", w);
    run<8>("DFMA t=fma(a,b,d)      [3 distinct, distinct dest]", w);
    run<9>("DMUL t=a*b             [distinct dest]", w);
  }
  return 0;
}
