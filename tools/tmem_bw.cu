// Developer micro-benchmark: TMEM read throughput of one SM as the fused SA body uses it
// (tcgen05.ld.32x32b.xN from 4 / 8 / 16 warps, each warp in its own lane quadrant).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/_bin/tmem_bw tools/tmem_bw.cu && tools/_bin/tmem_bw
// Prints bytes per cycle per SM for x16 / x32 / x64 loads, with the wait after every load or after four.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

template <int N>
__device__ __forceinline__ void ld(uint32_t taddr, uint32_t (&r)[N]);
template <>
__device__ __forceinline__ void ld<16>(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr) : "memory");
}
template <>
__device__ __forceinline__ void ld<32>(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// DEPTH loads of N columns in flight per wait
template <int N, int DEPTH>
__global__ void tmem_read_kernel(int iters, long long* cycles, uint32_t* sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = slot + (static_cast<uint32_t>((warp & 3) * 32) << 16) + ((warp >> 2) * 128) % 512;
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    uint32_t r[DEPTH][N];
#pragma unroll
    for (int d = 0; d < DEPTH; ++d) ld<N>(base + (d * N) % 128, r[d]);
    wait_ld();
#pragma unroll
    for (int d = 0; d < DEPTH; ++d)
#pragma unroll
      for (int i = 0; i < N; ++i) acc ^= r[d][i];
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  if (acc == 0x12345u) sink[0] = acc;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(512) : "memory");
}

template <int N, int DEPTH>
void run(int warps, int ctas) {
  long long* cyc;
  uint32_t* sink;
  cudaMalloc(&cyc, sizeof(long long) * ctas);
  cudaMalloc(&sink, 4);
  const int iters = 2000;
  for (int rep = 0; rep < 2; ++rep) tmem_read_kernel<N, DEPTH><<<ctas, warps * 32>>>(iters, cyc, sink);
  cudaError_t e = cudaDeviceSynchronize();
  long long h = 0;
  cudaMemcpy(&h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  const double bytes = double(iters) * DEPTH * N * 4.0 * 32 * warps;
  printf("x%-3d depth %d warps %2d ctas %3d: %8lld cycles, %7.1f B/cycle/SM, %6.1f cycles per load per warp  %s\n", N, DEPTH,
         warps, ctas, h, bytes / h, double(h) / (double(iters) * DEPTH), e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(cyc); cudaFree(sink);
}

int main() {
  for (int ctas : {1, 148}) {
    for (int warps : {4, 8, 16}) {
      run<16, 1>(warps, ctas);
      run<16, 4>(warps, ctas);
      run<32, 1>(warps, ctas);
      run<32, 2>(warps, ctas);
    }
  }
  return 0;
}
