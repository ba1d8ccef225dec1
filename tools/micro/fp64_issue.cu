// Microbenchmark (B200): what keeps the FP64 pipe from its 2-cycles-per-warp-instruction rate?
//   A. latency / ILP: N independent DFMA chains per warp, 1..4 warps per SMSP  -> cycles per DFMA per SMSP
//   B. interference: 8 DFMAs (2 register sources + constant) interleaved with K other instructions per group
//      (LDS.64, IADD3/LOP, FFMA, MOV-like) -> does the DFMA stream slow down?
//   C. three-register-source DFMAs with and without operand reuse, mixed with LDS
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_issue fp64_issue.cu ; cycles measured with clock64 per warp.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int N>
__global__ void __launch_bounds__(128) k_chain(double* out, long long* cyc, int iters, double m, double c) {
  double a[N];
#pragma unroll
  for (int i = 0; i < N; ++i) a[i] = threadIdx.x * 1e-9 + i;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 64 / N; ++r)
#pragma unroll
      for (int i = 0; i < N; ++i) a[i] = fma(a[i], m, c);
  }
  const long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < N; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if ((threadIdx.x & 31) == 0) cyc[(blockIdx.x * blockDim.x + threadIdx.x) >> 5] = t1 - t0;
}

// MODE: 0 none, 1 LDS.64, 2 integer (IADD3 / LOP3), 3 FFMA, 4 LDS.64 + integer
template <int MODE, int K, int SRC3>
__global__ void __launch_bounds__(128) k_mix(double* out, long long* cyc, int iters, double m, double c) {
  __shared__ double sm[8 * 128];
  double a[8], b[8], d[8];
  unsigned x = threadIdx.x * 2654435761u, y = 12345u + threadIdx.x;
  float f = threadIdx.x * 1e-3f;
  double l = 0.0;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    a[i] = threadIdx.x * 1e-9 + i;
    b[i] = 1.0 + 1e-7 * (i + 1) + 1e-12 * threadIdx.x;
    d[i] = 1e-9 * (i + 2) + 1e-13 * threadIdx.x;
    sm[i * 128 + threadIdx.x] = 1e-12 * i;
  }
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (SRC3 == 0) a[i] = fma(b[i], m, a[i]);          // 2 register sources + constant
        if (SRC3 == 1) a[i] = fma(b[i], d[i], a[i]);       // 3 distinct register sources
        if (SRC3 == 2) a[i] = fma(b[r], d[i], a[i]);       // 3 sources, first one reused by 8 consecutive DFMAs
        if (i < K) {
          if (MODE == 1 || MODE == 4) l += ((volatile double*)sm)[((i + r) & 7) * 128 + threadIdx.x];
          if (MODE == 2 || MODE == 4) { x = (x + y) ^ (x >> 3); y += x; }
          if (MODE == 3) f = fmaf(f, 1.0001f, 0.5f);
        }
      }
    }
  }
  const long long t1 = clock64();
  double s = l + x + y + f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if ((threadIdx.x & 31) == 0) cyc[(blockIdx.x * blockDim.x + threadIdx.x) >> 5] = t1 - t0;
}

static double* buf;
static long long* cyc;
static long long hc[148 * 64];

template <class F>
void run(const char* name, int warps_per_smsp, int fp64_per_iter, int iters, F launch) {
  const int blocks = 148 * warps_per_smsp;  // 128 threads = 4 warps = one per SMSP
  launch(blocks, 4);
  launch(blocks, iters);
  cudaDeviceSynchronize();
  cudaMemcpy(hc, cyc, sizeof(long long) * blocks * 4, cudaMemcpyDeviceToHost);
  double mean = 0;
  for (int i = 0; i < blocks * 4; ++i) mean += hc[i];
  mean /= blocks * 4;
  // each warp ran fp64_per_iter*iters FP64 instructions in `mean` cycles while sharing its SMSP with warps_per_smsp-1 others
  const double cyc_per_inst = mean / ((double)fp64_per_iter * iters * warps_per_smsp);
  printf("%-58s warps/SMSP=%d  cycles per FP64 warp-instr per SMSP = %.3f  (pipe rate 2.0)\n", name, warps_per_smsp, cyc_per_inst);
}

int main() {
  cudaMalloc(&buf, sizeof(double) * 148 * 8 * 128);
  cudaMalloc(&cyc, sizeof(long long) * 148 * 8 * 4);
  const int iters = 512;
  for (int w : {1, 2, 3, 4}) {
    run("A chain ILP=1", w, 64, iters, [&](int b, int it) { k_chain<1><<<b, 128>>>(buf, cyc, it, 1.0000001, 1e-9); });
    run("A chain ILP=2", w, 64, iters, [&](int b, int it) { k_chain<2><<<b, 128>>>(buf, cyc, it, 1.0000001, 1e-9); });
    run("A chain ILP=4", w, 64, iters, [&](int b, int it) { k_chain<4><<<b, 128>>>(buf, cyc, it, 1.0000001, 1e-9); });
    run("A chain ILP=8", w, 64, iters, [&](int b, int it) { k_chain<8><<<b, 128>>>(buf, cyc, it, 1.0000001, 1e-9); });
  }
// LDS modes add one DADD per load (the loaded value is consumed): count it as an FP64 instruction
#define MIX(MODE, K, S, label) \
  run(label, w, 64 + ((MODE == 1 || MODE == 4) ? 8 * K : 0), iters, [&](int b, int it) { k_mix<MODE, K, S><<<b, 128>>>(buf, cyc, it, 1.0000001, 1e-9); })
  for (int w : {2, 4}) {
    MIX(0, 0, 0, "B 8 DFMA(2 reg src)");
    MIX(1, 2, 0, "B 8 DFMA(2 reg src) + 2 LDS.64");
    MIX(1, 4, 0, "B 8 DFMA(2 reg src) + 4 LDS.64");
    MIX(1, 8, 0, "B 8 DFMA(2 reg src) + 8 LDS.64");
    MIX(2, 2, 0, "B 8 DFMA(2 reg src) + 2x3 int ops");
    MIX(2, 4, 0, "B 8 DFMA(2 reg src) + 4x3 int ops");
    MIX(2, 8, 0, "B 8 DFMA(2 reg src) + 8x3 int ops");
    MIX(3, 4, 0, "B 8 DFMA(2 reg src) + 4 FFMA");
    MIX(3, 8, 0, "B 8 DFMA(2 reg src) + 8 FFMA");
    MIX(4, 4, 0, "B 8 DFMA(2 reg src) + 4 LDS.64 + 4x3 int");
    MIX(0, 0, 1, "C 8 DFMA(3 reg src)");
    MIX(1, 4, 1, "C 8 DFMA(3 reg src) + 4 LDS.64");
    MIX(2, 4, 1, "C 8 DFMA(3 reg src) + 4x3 int ops");
    MIX(0, 0, 2, "C 8 DFMA(3 reg src, 1st reused x8)");
    MIX(1, 4, 2, "C 8 DFMA(3 reg src, 1st reused x8) + 4 LDS.64");
    MIX(2, 4, 2, "C 8 DFMA(3 reg src, 1st reused x8) + 4x3 int ops");
  }
  return 0;
}
