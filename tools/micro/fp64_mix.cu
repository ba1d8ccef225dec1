// Microbenchmark (B200): throughput of an FP64 instruction stream mixed with other instruction types, all streams
// INDEPENDENT (no dependent chains between the mixed-in instructions), inline PTX so that nothing is merged or removed.
// Reports SM cycles per loop iteration per scheduler (SMSP) from the wall time of a grid that puts W warps on every SMSP.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_mix fp64_mix.cu
#include <cstdio>
#include <cuda_runtime.h>

#define FMA2(i) asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(a[i]) : "d"(b[i]), "d"(m));
#define FMA3(i) asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(a[i]) : "d"(b[i]), "d"(d[i]));
#define FMA3R(i, r) asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(a[i]) : "d"(b[r]), "d"(d[i]));
#define ADD2(i) asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(a[i]) : "d"(b[i]));
#define LDS(j) asm volatile("ld.shared.f64 %0, [%1];" : "=d"(l[j]) : "r"(saddr + 1024 * (j)));
#define STS(j) asm volatile("st.shared.f64 [%0], %1;" ::"r"(saddr + 1024 * (j)), "d"(b[j]));
#define IAD(j) asm volatile("add.s32 %0, %0, %1;" : "+r"(x[j]) : "r"(y0));
#define FFM(j) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[j]) : "f"(f1), "f"(f2));

// MODE: FP64 flavour (0: FMA 2 reg src, 1: FMA 3 reg src, 2: FMA 3 src with first reused over 8, 3: DADD)
// EXTRA: mixed-in instruction (0 none, 1 LDS.64, 2 STS.64, 3 IADD, 4 FFMA), K of them per 8 FP64 instructions
template <int MODE, int EXTRA, int K>
__global__ void __launch_bounds__(128) k(double* out, int iters, double m, int y0, float f1, float f2) {
  __shared__ double sm[8 * 128];
  double a[8], b[8], d[8], l[8];
  int x[8];
  float f[8];
  for (int i = 0; i < 8; ++i) {
    a[i] = threadIdx.x * 1e-9 + i; b[i] = 1.0 + 1e-7 * (i + 1) + 1e-12 * threadIdx.x; d[i] = 1e-9 * (i + 2) + 1e-13 * threadIdx.x;
    l[i] = 0; x[i] = i + threadIdx.x; f[i] = i * 0.5f; sm[i * 128 + threadIdx.x] = i;
  }
  const unsigned saddr = (unsigned)__cvta_generic_to_shared(sm + threadIdx.x);
  __syncthreads();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (MODE == 0) { FMA2(i) }
        if (MODE == 1) { FMA3(i) }
        if (MODE == 2) { FMA3R(i, r) }
        if (MODE == 3) { ADD2(i) }
        if (i < K) {
          if (EXTRA == 1) { LDS(i) }
          if (EXTRA == 2) { STS(i) }
          if (EXTRA == 3) { IAD(i) }
          if (EXTRA == 4) { FFM(i) }
        }
      }
    }
  }
  double s = 0;
  for (int i = 0; i < 8; ++i) s += a[i] + l[i] + x[i] + f[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

static double* buf;
template <int MODE, int EXTRA, int K>
void run(const char* name) {
  for (int w : {1, 2, 4}) {
    const int blocks = 148 * w, iters = 2000;
    k<MODE, EXTRA, K><<<blocks, 128>>>(buf, 10, 1.0000001, 3, 1.0001f, 0.5f);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
      cudaEventRecord(e0);
      k<MODE, EXTRA, K><<<blocks, 128>>>(buf, iters, 1.0000001, 3, 1.0001f, 0.5f);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      if (ms < best) best = ms;
    }
    // cycles at 1.965 GHz per group of 8 FP64 (+K extra) instructions per SMSP, all w warps of the SMSP together
    const double cyc = best * 1e-3 * 1.965e9 / ((double)iters * 8 * w);
    printf("%-46s w=%d  %.2f cycles per [8 FP64 + %d extra] per SMSP  (8 FP64 alone = 16)\n", name, w, cyc, K);
  }
}

int main() {
  cudaMalloc(&buf, sizeof(double) * 148 * 4 * 128);
  run<0, 0, 0>("DFMA 2 reg src");
  run<1, 0, 0>("DFMA 3 reg src");
  run<2, 0, 0>("DFMA 3 reg src, first reused over 8");
  run<3, 0, 0>("DADD 2 reg src");
  run<0, 1, 2>("DFMA 2 reg src + LDS.64 x2");
  run<0, 1, 4>("DFMA 2 reg src + LDS.64 x4");
  run<0, 1, 8>("DFMA 2 reg src + LDS.64 x8");
  run<1, 1, 4>("DFMA 3 reg src + LDS.64 x4");
  run<0, 2, 4>("DFMA 2 reg src + STS.64 x4");
  run<0, 3, 4>("DFMA 2 reg src + IADD x4");
  run<0, 3, 8>("DFMA 2 reg src + IADD x8");
  run<1, 3, 8>("DFMA 3 reg src + IADD x8");
  run<0, 4, 8>("DFMA 2 reg src + FFMA x8");
  run<3, 1, 4>("DADD 2 reg src + LDS.64 x4");
  return 0;
}
