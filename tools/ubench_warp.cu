// Developer micro-benchmark: latency / throughput of the warp-collective primitives the FPS round
// loop can be built from (redux.sync, vote.ballot, shfl, bar.sync, LDS), for 1..8 warps per SMSP.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/_build/ubench_warp tools/ubench_warp.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int OP>
__global__ void k(int iters, long long* out, int* sink) {
  __shared__ int sm[64];
  int v = threadIdx.x * 7 + 3;
  unsigned u = threadIdx.x;
  sm[threadIdx.x & 63] = v;
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    if (OP == 0) v = __reduce_max_sync(0xffffffffu, v + i);
    if (OP == 1) u = __ballot_sync(0xffffffffu, (u + i) & 1);
    if (OP == 2) v = __shfl_xor_sync(0xffffffffu, v + i, 16);
    if (OP == 3) { __syncthreads(); v += i; }
    if (OP == 4) v = sm[(v + i) & 63];
    if (OP == 5) u = __reduce_min_sync(0xffffffffu, u + i);
    if (OP == 6) { v = __shfl_xor_sync(0xffffffffu, v + i, 16); v = max(v, __shfl_xor_sync(0xffffffffu, v, 8)); }
    if (OP == 7) u = __match_any_sync(0xffffffffu, (u + i) & 3);
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  if (v == 123456789 && u == 77) *sink = v;
}

template <int OP>
void run(const char* name) {
  long long* d; int* s;
  cudaMalloc(&d, 8 * 148); cudaMalloc(&s, 4);
  const int iters = 2000;
  printf("%-22s", name);
  for (int threads : {32, 128, 256, 512, 1024}) {
    k<OP><<<1, threads>>>(iters, d, s);
    cudaDeviceSynchronize();
    k<OP><<<1, threads>>>(iters, d, s);
    long long h; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    printf("  %4d thr: %6.1f", threads, double(h) / iters);
  }
  printf("   cycles/iter (thread 0)\n");
}

int main() {
  run<0>("redux.max.s32");
  run<5>("redux.min.u32");
  run<1>("vote.ballot");
  run<2>("shfl.bfly");
  run<6>("shfl x2 + max");
  run<7>("match.any");
  run<3>("bar.sync");
  run<4>("lds dependent");
  return 0;
}
