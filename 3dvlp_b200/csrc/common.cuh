// Shared helpers for the sm_100a kernels behind include/pointnet2_b200.h.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <cmath>
#include <cstdarg>
#include <cstdio>

#include "../../include/pointnet2_b200.h"

namespace pn2 {

// ---- error reporting: every entry point returns a status, nothing ever calls exit() ------------
// (the reference aborts the process on a launch failure, cuda_utils.h:35-44)
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);

#define PN2_CUDA(expr)                                              \
  do {                                                              \
    cudaError_t _e = (expr);                                        \
    if (_e != cudaSuccess) return ::pn2::cuda_fail(_e, #expr);      \
  } while (0)

#define PN2_CHECK_LAUNCH(name)                                      \
  do {                                                              \
    cudaError_t _e = cudaGetLastError();                            \
    if (_e != cudaSuccess) return ::pn2::cuda_fail(_e, name);       \
  } while (0)

#define PN2_REQUIRE(cond, ...)                                      \
  do {                                                              \
    if (!(cond)) {                                                  \
      ::pn2::set_error(__VA_ARGS__);                                \
      return PN2_ERR_INVALID;                                       \
    }                                                               \
  } while (0)

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// cuda_utils.h:20-24 of the reference: the block size its launchers pick.  The FPS tie-break is a
// function of this number, so the host computes it with the very same double-precision expression.
inline int opt_n_threads(int work_size) {
  const int pow_2 = static_cast<int>(std::log(static_cast<double>(work_size)) / std::log(2.0));
  int t = 1 << pow_2;
  if (t > 512) t = 512;
  if (t < 1) t = 1;
  return t;
}

int sm_count();
// Raise the release threshold of the current device's default memory pool once (thread-safe): the
// default of 0 hands freed workspace back to the driver at every synchronisation, so the next
// cudaMallocAsync would pay for a real allocation.
void keep_workspace_pool();

// run-time options (pn2_set_option); initial values come from the environment
enum Option { OPT_FPS_BUCKET = 0, OPT_BQ_GRID = 1, OPT_SA_TILES_PER_GROUP = 2, OPT_WGRAD_TMA = 3, OPT_COUNT = 4 };
int option_value(Option o);

// host-side launch statistics, read back with pn2_get_stat(name): the parity tests assert through them
// that the kernel variant they mean to cover is the one that ran
enum Stat { STAT_FPS_CLUSTER = 0, STAT_FPS_BUCKET, STAT_FPS_PAIR, STAT_BQ_GRID, STAT_BQ_SMALL_GRID, STAT_BQ_SCAN,
            STAT_NN_SMALL, STAT_NN_TILED, STAT_WGRAD_TMA, STAT_WGRAD_CPASYNC, STAT_COUNT };
void stat_add(Stat s, long long v = 1);

// NVTX range around every C-ABI entry point (header-only NVTX3: the tools library is looked up at run
// time, there is no link dependency and no cost when no tool is attached)
struct NvtxRange {
  explicit NvtxRange(const char* name);
  ~NvtxRange();
};
#define PN2_RANGE(name) ::pn2::NvtxRange _pn2_range(name)

// ---- device helpers -------------------------------------------------------------------------

// Squared distance exactly as nvcc contracts the reference expression
// `dx*dx + dy*dy + dz*dz` (ball_query_gpu.cu:36-37, interpolate_gpu.cu:38, sampling_gpu.cu:108-109)
// for sm_100a at -O2:  FMUL dy,dy ; FFMA dx,dx ; FFMA dz,dz.  Written with explicit intrinsics so
// no compiler version or flag can re-associate it.
__device__ __forceinline__ float sqdist3(float dx, float dy, float dz) {
  return __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}

__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n"
               "barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}

__device__ __forceinline__ uint32_t mapa_u32(uint32_t local_smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(rank));
  return r;
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}

__device__ __forceinline__ void fence_mbar_init_cluster() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}

// Bounded wait: a protocol bug must surface as an error flag, never as a hung GPU box.
// Returns false after ~2^26 polls (seconds) so the caller can bail out.
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity) {
#pragma unroll 1
  for (uint32_t it = 0; it < (1u << 26); ++it) {
    if (mbar_try_wait(bar, parity)) return true;
  }
  return false;
}

}  // namespace pn2
