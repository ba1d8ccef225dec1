// Microbenchmark: FP64 RED (atomicAdd without return) throughput vs. address pattern within a warp.
//   mode 0: every lane its own random 8-byte slot (different sectors)
//   mode 1: groups of G consecutive lanes hit G consecutive doubles (same 32 B sector for G<=4, aligned)
//   mode 2: all 32 lanes consecutive doubles (256 B contiguous)
#include <cstdio>
#include <cuda_runtime.h>
#include <cstdint>

__device__ __forceinline__ uint32_t hash(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x;
}

template <int G>
__global__ void k(double* buf, uint32_t nslots, int iters) {
  const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t lane = threadIdx.x & 31, warp = tid >> 5;
  for (int it = 0; it < iters; ++it) {
    // one random group base per (warp, group, iteration); group = lane / G
    const uint32_t grp = lane / G;
    const uint32_t base = (hash(warp * 1315423911u + grp * 2654435761u + it * 97u) % (nslots / G)) * G;
    atomicAdd(buf + base + (lane % G), 1.0);
  }
}

template <int G>
void run(double* buf, uint32_t nslots) {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int blocks = sms * 16, threads = 256, iters = 256;
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  k<G><<<blocks, threads>>>(buf, nslots, 4);
  cudaEventRecord(a);
  k<G><<<blocks, threads>>>(buf, nslots, iters);
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  const double n = (double)blocks * threads * iters;
  printf("G=%2d consecutive doubles per group: %.1f G RED/s  (%.1f G groups/s)\n", G, n / ms / 1e6, n / G / ms / 1e6);
}

int main() {
  const uint32_t nslots = 24u << 20;  // 192 MB of doubles: larger than L2
  double* buf; cudaMalloc(&buf, sizeof(double) * nslots); cudaMemset(buf, 0, sizeof(double) * nslots);
  run<1>(buf, nslots); run<2>(buf, nslots); run<3>(buf, nslots); run<4>(buf, nslots); run<8>(buf, nslots); run<16>(buf, nslots); run<32>(buf, nslots);
  const uint32_t small = 2u << 20;  // 16 MB: L2 resident
  printf("-- 16 MB target (L2 resident)\n");
  run<1>(buf, small); run<3>(buf, small); run<4>(buf, small); run<32>(buf, small);
  return 0;
}
