// Developer micro-benchmark: issue rate of scalar FP32 (FADD/FFMA/FMNMX, 3-register forms) versus the
// packed add/mul/fma.f32x2 forms of sm_100, for the FPS / ball-query inner loops.
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned long long f2fma(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long d;
  asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ unsigned long long f2add(unsigned long long a, unsigned long long b) {
  unsigned long long d;
  asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}

template <int OP>
__global__ void k(int iters, long long* out, float* sink, float s) {
  float a[8];
  unsigned long long p[8];
  for (int i = 0; i < 8; ++i) { a[i] = threadIdx.x * 0.001f + i; p[i] = (unsigned long long)(threadIdx.x + i) * 0x3f8000013f800001ull; }
  float b = s * 1.0001f, c = s * 0.5f;
  unsigned long long pb = 0x3f8000003f800000ull, pc = 0x3a0000003a000000ull;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (OP == 0) a[i] = __fmaf_rn(a[i], b, c);             // FFMA 3-reg
      if (OP == 1) a[i] = __fadd_rn(a[i], b);                // FADD
      if (OP == 2) a[i] = fminf(a[i], b + i);                // FMNMX (alu pipe)
      if (OP == 3) p[i] = f2fma(p[i], pb, pc);               // FFMA2
      if (OP == 4) p[i] = f2add(p[i], pb);                   // FADD2
      if (OP == 5) { a[i] = __fmaf_rn(a[i], a[i], c); }      // FFMA with 2 distinct regs
    }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  float r = 0; for (int i = 0; i < 8; ++i) r += a[i] + __uint_as_float((unsigned)p[i]);
  if (r == 12345.678f) *sink = r;
}

template <int OP>
void run(const char* name) {
  long long* d; float* s;
  cudaMalloc(&d, 8 * 148); cudaMalloc(&s, 4);
  const int iters = 2000;
  printf("%-28s", name);
  for (int threads : {32, 128, 256, 512, 1024}) {
    k<OP><<<1, threads>>>(iters, d, s, 1.0f);
    cudaDeviceSynchronize();
    k<OP><<<1, threads>>>(iters, d, s, 1.0f);
    long long h; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    // warp-instructions per SMSP per cycle
    const double warps_per_smsp = threads / 32 / 4.0 < 1 ? 1 : threads / 32 / 4.0;
    printf("  %4d thr: %5.2f cyc/instr/SMSP", threads, double(h) / (iters * 8.0 * warps_per_smsp));
  }
  printf("\n");
}

int main() {
  run<0>("FFMA (a*b+c, 3 regs)");
  run<5>("FFMA (a*a+c)");
  run<1>("FADD");
  run<2>("FMNMX");
  run<3>("fma.rn.f32x2 (2 FMAs)");
  run<4>("add.rn.f32x2 (2 adds)");
  return 0;
}
