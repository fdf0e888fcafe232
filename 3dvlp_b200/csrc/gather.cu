// Bandwidth-bound index ops: gather_points, group_points, three_interpolate and their scatter-add
// gradients.  One thread per output element, the fastest-varying output index on threadIdx.x so the
// index reads and the stores are coalesced; the grid covers (elements, channels, scenes) instead of
// the reference's one CTA per scene (sampling_gpu.cu:13-52, group_points_gpu.cu:13-69,
// interpolate_gpu.cu:77-148).
//
// The reference layout (B,C,N) makes every gathered element its own 32-byte sector, so the direct
// kernels above are bound by L2 sector traffic, not by HBM.  Whenever the indexed row (N floats) or a
// tile of it fits in shared memory, the "staged" kernels below read each row ONCE with a TMA bulk copy
// (or plain coalesced loads when the row is not 16-byte aligned), gather / accumulate in shared memory
// and write the result with coalesced 16-byte streaming stores: DRAM traffic equals the algorithmic
// bytes.  Results are identical to the direct kernels (pure copies; the interpolation keeps its fma
// chain); the gradients only differ in summation order, which the reference's atomics leave open too.
#include "common.cuh"
#include "tc.cuh"

namespace pn2 {
namespace {

constexpr int T = 256;

// out[b,c,q] = points[b,c,idx[b,q]]     (q over npoints, or over npoints*nsample for grouping)
__global__ void __launch_bounds__(T)
gather_rows_kernel(int c, int n, long long q_total, const float* __restrict__ points,
                   const int* __restrict__ idx, float* __restrict__ out) {
  const long long q = static_cast<long long>(blockIdx.x) * T + threadIdx.x;
  if (q >= q_total) return;
  const int scene = blockIdx.z;
  const int a = __ldg(idx + static_cast<size_t>(scene) * q_total + q);
  for (int l = blockIdx.y; l < c; l += gridDim.y) {
    const size_t row = static_cast<size_t>(scene) * c + l;
    out[row * q_total + q] = __ldg(points + row * n + a);
  }
}

// grad_points[b,c,idx[b,q]] += grad_out[b,c,q]
__global__ void __launch_bounds__(T)
scatter_rows_kernel(int c, int n, long long q_total, const float* __restrict__ grad_out,
                    const int* __restrict__ idx, float* __restrict__ grad_points) {
  const long long q = static_cast<long long>(blockIdx.x) * T + threadIdx.x;
  if (q >= q_total) return;
  const int scene = blockIdx.z;
  const int a = __ldg(idx + static_cast<size_t>(scene) * q_total + q);
  for (int l = blockIdx.y; l < c; l += gridDim.y) {
    const size_t row = static_cast<size_t>(scene) * c + l;
    atomicAdd(grad_points + row * n + a, __ldg(grad_out + row * q_total + q));
  }
}

// out[b,l,j] = fma(p3,w3, fma(p1,w1, p2*w2))  -- the contraction nvcc picks for
// interpolate_gpu.cu:103-104 (SASS of the reference build); pinned for reproducibility.
__global__ void __launch_bounds__(T)
three_interpolate_kernel(int c, int m, int n, const float* __restrict__ points,
                         const int* __restrict__ idx, const float* __restrict__ weight,
                         float* __restrict__ out) {
  const int j = blockIdx.x * T + threadIdx.x;
  if (j >= n) return;
  const int scene = blockIdx.z;
  const size_t o = (static_cast<size_t>(scene) * n + j) * 3;
  const int i1 = __ldg(idx + o), i2 = __ldg(idx + o + 1), i3 = __ldg(idx + o + 2);
  const float w1 = __ldg(weight + o), w2 = __ldg(weight + o + 1), w3 = __ldg(weight + o + 2);
  for (int l = blockIdx.y; l < c; l += gridDim.y) {
    const float* p = points + (static_cast<size_t>(scene) * c + l) * m;
    const float v = __fmaf_rn(__ldg(p + i3), w3, __fmaf_rn(__ldg(p + i1), w1, __fmul_rn(__ldg(p + i2), w2)));
    out[(static_cast<size_t>(scene) * c + l) * n + j] = v;
  }
}

__global__ void __launch_bounds__(T)
three_interpolate_grad_kernel(int c, int n, int m, const float* __restrict__ grad_out,
                              const int* __restrict__ idx, const float* __restrict__ weight,
                              float* __restrict__ grad_points) {
  const int j = blockIdx.x * T + threadIdx.x;
  if (j >= n) return;
  const int scene = blockIdx.z;
  const size_t o = (static_cast<size_t>(scene) * n + j) * 3;
  const int i1 = __ldg(idx + o), i2 = __ldg(idx + o + 1), i3 = __ldg(idx + o + 2);
  const float w1 = __ldg(weight + o), w2 = __ldg(weight + o + 1), w3 = __ldg(weight + o + 2);
  for (int l = blockIdx.y; l < c; l += gridDim.y) {
    float* g = grad_points + (static_cast<size_t>(scene) * c + l) * m;
    const float go = __ldg(grad_out + (static_cast<size_t>(scene) * c + l) * n + j);
    atomicAdd(g + i1, go * w1);
    atomicAdd(g + i2, go * w2);
    atomicAdd(g + i3, go * w3);
  }
}


// ---------------------------------------------------------------------------------------------
// Staged variants: one CTA per (scene, channel) row held in shared memory.
// ---------------------------------------------------------------------------------------------
constexpr int kStagedSmemMax = 227 * 1024 - 64;

__device__ __forceinline__ bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// Bring src[0..n) into srow.  TMA bulk copies when the row is 16-byte aligned and n % 4 == 0.
__device__ __forceinline__ void stage_row(float* srow, const float* __restrict__ src, int n, uint64_t* bar, int* err) {
  if (aligned16(src) && (n & 3) == 0) {
    if (threadIdx.x == 0) {
      mbar_init(bar, 1);
      fence_mbar_init_cluster();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      const uint32_t bytes = static_cast<uint32_t>(n) * 4u;
      mbar_arrive_expect_tx(bar, bytes);
      for (uint32_t o = 0; o < bytes; o += 32768u)
        tc::bulk_g2s(reinterpret_cast<char*>(srow) + o, reinterpret_cast<const char*>(src) + o,
                     bytes - o < 32768u ? bytes - o : 32768u, bar);
    }
    if (!mbar_wait(bar, 0) && err) atomicExch(err, 3);
  } else {
    for (int i = threadIdx.x; i < n; i += blockDim.x) srow[i] = __ldg(src + i);
    __syncthreads();
  }
}

// out[b,l,q] = points[b,l,idx[b,q]] with the row points[b,l,:] in shared memory.
__global__ void __launch_bounds__(1024)
gather_rows_staged_kernel(int c, int n, long long q_total, const float* __restrict__ points,
                          const int* __restrict__ idx, float* __restrict__ out, int* err) {
  extern __shared__ __align__(16) float srow[];
  __shared__ uint64_t bar;
  const int l = blockIdx.x, scene = blockIdx.y;
  const size_t row = static_cast<size_t>(scene) * c + l;
  const int* ip = idx + static_cast<size_t>(scene) * q_total;
  float* op = out + row * q_total;
  stage_row(srow, points + row * n, n, &bar, err);
  if ((q_total & 3) == 0 && aligned16(ip) && aligned16(op)) {
    const int4* ip4 = reinterpret_cast<const int4*>(ip);
    float4* op4 = reinterpret_cast<float4*>(op);
    const long long q4 = q_total >> 2;
    // four index vectors in flight per thread: the loop is bound by the latency of the index stream
    const long long bd = blockDim.x;
    for (long long q = threadIdx.x; q < q4; q += 4 * bd) {
      int4 a[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) a[u] = q + u * bd < q4 ? __ldg(ip4 + q + u * bd) : make_int4(0, 0, 0, 0);
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (q + u * bd < q4) __stcs(op4 + q + u * bd, make_float4(srow[a[u].x], srow[a[u].y], srow[a[u].z], srow[a[u].w]));
    }
  } else {
    for (long long q = threadIdx.x; q < q_total; q += blockDim.x) op[q] = srow[__ldg(ip + q)];
  }
}

// grad_points[b,l,idx[b,q]] += grad_out[b,l,q], accumulated in a shared-memory copy of the row.
// Ball-query padding repeats one index over the tail of a ball: runs of equal indices inside a
// thread's four elements are summed in registers before they reach the shared-memory atomic.
__global__ void __launch_bounds__(1024)
scatter_rows_staged_kernel(int c, int n, long long q_total, const float* __restrict__ grad_out,
                           const int* __restrict__ idx, float* __restrict__ grad_points) {
  extern __shared__ __align__(16) float srow[];
  const int l = blockIdx.x, scene = blockIdx.y;
  const size_t row = static_cast<size_t>(scene) * c + l;
  const int* ip = idx + static_cast<size_t>(scene) * q_total;
  const float* gp = grad_out + row * q_total;
  for (int i = threadIdx.x; i < n; i += blockDim.x) srow[i] = 0.f;
  __syncthreads();
  if ((q_total & 3) == 0 && aligned16(ip) && aligned16(gp)) {
    const int4* ip4 = reinterpret_cast<const int4*>(ip);
    const float4* gp4 = reinterpret_cast<const float4*>(gp);
    const long long q4 = q_total >> 2;
    const long long bd = blockDim.x;
    for (long long q0 = threadIdx.x; q0 < q4; q0 += 4 * bd) {
      int4 av[4];
      float4 vv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const bool in = q0 + u * bd < q4;
        av[u] = in ? __ldg(ip4 + q0 + u * bd) : make_int4(0, 0, 0, 0);
        vv[u] = in ? __ldcs(gp4 + q0 + u * bd) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
      if (q0 + u * bd >= q4) break;
      const int4 a = av[u];
      const float4 v = vv[u];
      int cur = a.x;
      float acc = v.x;
      if (a.y == cur) acc += v.y; else { atomicAdd(srow + cur, acc); cur = a.y; acc = v.y; }
      if (a.z == cur) acc += v.z; else { atomicAdd(srow + cur, acc); cur = a.z; acc = v.z; }
      if (a.w == cur) acc += v.w; else { atomicAdd(srow + cur, acc); cur = a.w; acc = v.w; }
      atomicAdd(srow + cur, acc);
      }
    }
  } else {
    for (long long q = threadIdx.x; q < q_total; q += blockDim.x) atomicAdd(srow + __ldg(ip + q), __ldg(gp + q));
  }
  __syncthreads();
  float* dst = grad_points + row * n;  // this CTA owns the row: plain read-modify-write
  for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] += srow[i];
}

// three_interpolate with a channel tile of the known features transposed into shared memory as
// tile[m][CT + 4]: one 16-byte shared load per neighbour fetches four channels, so the three gathers
// cost 3 LDS.128 per four outputs.  Warp = 32 consecutive unknowns x 4 channels.
template <int CT>
__global__ void __launch_bounds__(1024)
three_interpolate_staged_kernel(int c, int m, int n, int jchunk, const float* __restrict__ points,
                                const int* __restrict__ idx, const float* __restrict__ weight,
                                float* __restrict__ out) {
  extern __shared__ __align__(16) float tile[];
  constexpr int LD = CT + 4, G = CT / 4;
  const int scene = blockIdx.z, c0 = blockIdx.y * CT;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const int SETS = nwarps / G;
  for (int ch = warp; ch < CT; ch += nwarps) {
    const bool live = c0 + ch < c;
    const float* src = points + (static_cast<size_t>(scene) * c + (live ? c0 + ch : 0)) * m;
    for (int i = lane; i < m; i += 32) tile[i * LD + ch] = live ? __ldg(src + i) : 0.f;
  }
  __syncthreads();
  const int cg = warp % G, jw = warp / G;
  const int j_end = min(n, (blockIdx.x + 1) * jchunk);
  const float4* t4 = reinterpret_cast<const float4*>(tile);
  // two unknowns per thread and iteration, all index / weight loads issued before the first use:
  // the loop is bound by the latency of that stream, not by the shared-memory gathers
  constexpr int U = 2;
  for (int jb = blockIdx.x * jchunk + jw * 32; jb < j_end; jb += U * SETS * 32) {
    int ia[U][3];
    float wa[U][3];
    bool in[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int j = jb + u * SETS * 32 + lane;
      in[u] = j < j_end;
      const size_t o = (static_cast<size_t>(scene) * n + (in[u] ? j : 0)) * 3;
#pragma unroll
      for (int t = 0; t < 3; ++t) { ia[u][t] = __ldg(idx + o + t); wa[u][t] = __ldg(weight + o + t); }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (!in[u]) continue;
      const int j = jb + u * SETS * 32 + lane;
      const float w1 = wa[u][0], w2 = wa[u][1], w3 = wa[u][2];
      const float4 p1 = t4[ia[u][0] * (LD / 4) + cg], p2 = t4[ia[u][1] * (LD / 4) + cg], p3 = t4[ia[u][2] * (LD / 4) + cg];
      const float v[4] = {__fmaf_rn(p3.x, w3, __fmaf_rn(p1.x, w1, __fmul_rn(p2.x, w2))),
                          __fmaf_rn(p3.y, w3, __fmaf_rn(p1.y, w1, __fmul_rn(p2.y, w2))),
                          __fmaf_rn(p3.z, w3, __fmaf_rn(p1.z, w1, __fmul_rn(p2.z, w2))),
                          __fmaf_rn(p3.w, w3, __fmaf_rn(p1.w, w1, __fmul_rn(p2.w, w2)))};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int l = c0 + cg * 4 + k;
        if (l < c) __stcs(out + (static_cast<size_t>(scene) * c + l) * n + j, v[k]);
      }
    }
  }
}

// Gradient of the above: the tile accumulates grad_out * w per known point in shared memory
// (odd row pitch: the scalar atomics of a warp spread over all banks), then joins grad_points.
template <int CT>
__global__ void __launch_bounds__(1024)
three_interpolate_grad_staged_kernel(int c, int n, int m, int jchunk, const float* __restrict__ grad_out,
                                     const int* __restrict__ idx, const float* __restrict__ weight,
                                     float* __restrict__ grad_points) {
  extern __shared__ __align__(16) float tile[];
  constexpr int LD = CT + 1, G = CT / 4;
  const int scene = blockIdx.z, c0 = blockIdx.y * CT;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const int SETS = nwarps / G;
  for (int e = threadIdx.x; e < m * LD; e += blockDim.x) tile[e] = 0.f;
  __syncthreads();
  const int cg = warp % G, jw = warp / G;
  const int j_end = min(n, (blockIdx.x + 1) * jchunk);
  for (int jb = blockIdx.x * jchunk + jw * 32; jb < j_end; jb += SETS * 32) {
    const int j = jb + lane;
    if (j >= j_end) continue;
    const size_t o = (static_cast<size_t>(scene) * n + j) * 3;
    const int i1 = __ldg(idx + o), i2 = __ldg(idx + o + 1), i3 = __ldg(idx + o + 2);
    const float w1 = __ldg(weight + o), w2 = __ldg(weight + o + 1), w3 = __ldg(weight + o + 2);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int ch = cg * 4 + k;
      if (c0 + ch < c) {
        const float go = __ldcs(grad_out + (static_cast<size_t>(scene) * c + c0 + ch) * n + j);
        atomicAdd(tile + i1 * LD + ch, go * w1);
        atomicAdd(tile + i2 * LD + ch, go * w2);
        atomicAdd(tile + i3 * LD + ch, go * w3);
      }
    }
  }
  __syncthreads();
  for (int ch = warp; ch < CT; ch += nwarps) {
    if (c0 + ch >= c) break;
    float* dst = grad_points + (static_cast<size_t>(scene) * c + c0 + ch) * m;
    for (int i = lane; i < m; i += 32) atomicAdd(dst + i, tile[i * LD + ch]);
  }
}

// channels per tile for the staged interpolation: the largest that fits; 0 -> use the direct kernel
inline int interp_tile_channels(int c, int m, int pad) {
  for (int ct = 32; ct >= 4; ct >>= 1) {
    if (ct / 2 >= c && ct > 4) continue;  // do not stage dead channels
    if (static_cast<size_t>(m) * (ct + pad) * 4 <= static_cast<size_t>(kStagedSmemMax)) return ct;
  }
  return 0;
}

// unknowns per CTA: enough CTAs to fill the machine, few enough that staging the tile amortises
inline int interp_jchunk(int n, int m, int tiles) {
  int chunks = (2 * sm_count() + tiles - 1) / tiles;
  const int max_chunks = (n + 4 * m - 1) / (4 * m) > 0 ? (n + 4 * m - 1) / (4 * m) : 1;
  if (chunks > max_chunks) chunks = max_chunks;
  if (chunks < 1) chunks = 1;
  int jchunk = (n + chunks - 1) / chunks;
  return (jchunk + 1023) / 1024 * 1024;
}
// one CTA per SM when the tile is large: give it all the warps the SM can hold
inline int interp_threads(size_t smem) { return smem > 110 * 1024 ? 1024 : (smem > 54 * 1024 ? 512 : 256); }

inline dim3 grid_for(long long q_total, int c, int b) {
  return dim3(static_cast<unsigned>((q_total + T - 1) / T), static_cast<unsigned>(c < 65535 ? c : 65535),
              static_cast<unsigned>(b));
}

// shared-memory row staging pays when the row is re-used: at least n/2 gathered elements per row
inline bool staged_rows_ok(int n, long long q_total, int b) {
  return static_cast<size_t>(n) * 4 + 16 <= static_cast<size_t>(kStagedSmemMax) && q_total * 2 >= n && b <= 65535;
}
inline int staged_threads(int n, long long q_total) {
  if (q_total < 4096) return 256;
  return static_cast<size_t>(n) * 4 > 100 * 1024 ? 1024 : 512;
}

}  // namespace
int* device_err_flag();  // abi.cu
}  // namespace pn2

using namespace pn2;

static int launch_gather(const char* what, int b, int c, int n, long long q, const float* points, const int* idx,
                         float* out, cudaStream_t st) {
  if (staged_rows_ok(n, q, b)) {
    const size_t smem = (static_cast<size_t>(n) * 4 + 15) & ~static_cast<size_t>(15);
    PN2_CUDA(cudaFuncSetAttribute(gather_rows_staged_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  static_cast<int>(smem)));
    gather_rows_staged_kernel<<<dim3(c, b), staged_threads(n, q), smem, st>>>(c, n, q, points, idx, out,
                                                                              device_err_flag());
  } else {
    gather_rows_kernel<<<grid_for(q, c, b), T, 0, st>>>(c, n, q, points, idx, out);
  }
  PN2_CHECK_LAUNCH(what);
  return PN2_OK;
}

static int launch_scatter(const char* what, int b, int c, int n, long long q, const float* grad_out, const int* idx,
                          float* grad_points, cudaStream_t st) {
  if (staged_rows_ok(n, q, b)) {
    const size_t smem = (static_cast<size_t>(n) * 4 + 15) & ~static_cast<size_t>(15);
    PN2_CUDA(cudaFuncSetAttribute(scatter_rows_staged_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  static_cast<int>(smem)));
    scatter_rows_staged_kernel<<<dim3(c, b), staged_threads(n, q), smem, st>>>(c, n, q, grad_out, idx, grad_points);
  } else {
    scatter_rows_kernel<<<grid_for(q, c, b), T, 0, st>>>(c, n, q, grad_out, idx, grad_points);
  }
  PN2_CHECK_LAUNCH(what);
  return PN2_OK;
}

template <int CT>
static int launch_interp(int b, int c, int m, int n, const float* points, const int* idx, const float* weight,
                         float* out, cudaStream_t st) {
  const int tiles = (c + CT - 1) / CT;
  const int jchunk = interp_jchunk(n, m, tiles * b);
  const size_t smem = static_cast<size_t>(m) * (CT + 4) * 4;
  PN2_CUDA(cudaFuncSetAttribute(three_interpolate_staged_kernel<CT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                static_cast<int>(smem)));
  three_interpolate_staged_kernel<CT><<<dim3((n + jchunk - 1) / jchunk, tiles, b), interp_threads(smem), smem, st>>>(
      c, m, n, jchunk, points, idx, weight, out);
  PN2_CHECK_LAUNCH("three_interpolate");
  return PN2_OK;
}

template <int CT>
static int launch_interp_grad(int b, int c, int n, int m, const float* grad_out, const int* idx, const float* weight,
                              float* grad_points, cudaStream_t st) {
  const int tiles = (c + CT - 1) / CT;
  const int jchunk = interp_jchunk(n, m, tiles * b);
  const size_t smem = static_cast<size_t>(m) * (CT + 1) * 4;
  PN2_CUDA(cudaFuncSetAttribute(three_interpolate_grad_staged_kernel<CT>,
                                cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  three_interpolate_grad_staged_kernel<CT><<<dim3((n + jchunk - 1) / jchunk, tiles, b), interp_threads(smem), smem, st>>>(
      c, n, m, jchunk, grad_out, idx, weight, grad_points);
  PN2_CHECK_LAUNCH("three_interpolate_grad");
  return PN2_OK;
}

extern "C" int pn2_gather_points(int b, int c, int n, int npoints, const float* points, const int* idx,
                                 float* out, void* stream) {
  PN2_RANGE("pn2_gather_points");
  if (b <= 0 || c <= 0 || npoints <= 0) return PN2_OK;
  PN2_REQUIRE(points && idx && out && n > 0 && b <= 65535, "gather_points: bad arguments");
  return launch_gather("gather_points", b, c, n, npoints, points, idx, out, as_stream(stream));
}

extern "C" int pn2_gather_points_grad(int b, int c, int n, int npoints, const float* grad_out,
                                      const int* idx, float* grad_points, void* stream) {
  PN2_RANGE("pn2_gather_points_grad");
  if (b <= 0 || c <= 0 || npoints <= 0) return PN2_OK;
  PN2_REQUIRE(grad_out && idx && grad_points && n > 0 && b <= 65535, "gather_points_grad: bad arguments");
  return launch_scatter("gather_points_grad", b, c, n, npoints, grad_out, idx, grad_points, as_stream(stream));
}

extern "C" int pn2_group_points(int b, int c, int n, int npoints, int nsample, const float* points,
                                const int* idx, float* out, void* stream) {
  PN2_RANGE("pn2_group_points");
  if (b <= 0 || c <= 0 || npoints <= 0 || nsample <= 0) return PN2_OK;
  PN2_REQUIRE(points && idx && out && n > 0 && b <= 65535, "group_points: bad arguments");
  const long long q = static_cast<long long>(npoints) * nsample;
  return launch_gather("group_points", b, c, n, q, points, idx, out, as_stream(stream));
}

extern "C" int pn2_group_points_grad(int b, int c, int n, int npoints, int nsample, const float* grad_out,
                                     const int* idx, float* grad_points, void* stream) {
  PN2_RANGE("pn2_group_points_grad");
  if (b <= 0 || c <= 0 || npoints <= 0 || nsample <= 0) return PN2_OK;
  PN2_REQUIRE(grad_out && idx && grad_points && n > 0 && b <= 65535, "group_points_grad: bad arguments");
  const long long q = static_cast<long long>(npoints) * nsample;
  return launch_scatter("group_points_grad", b, c, n, q, grad_out, idx, grad_points, as_stream(stream));
}

extern "C" int pn2_three_interpolate(int b, int c, int m, int n, const float* points, const int* idx,
                                     const float* weight, float* out, void* stream) {
  PN2_RANGE("pn2_three_interpolate");
  if (b <= 0 || c <= 0 || n <= 0) return PN2_OK;
  PN2_REQUIRE(points && idx && weight && out && m > 0 && b <= 65535, "three_interpolate: bad arguments");
  switch (n >= 2 * m ? interp_tile_channels(c, m, 4) : 0) {
    case 32: return launch_interp<32>(b, c, m, n, points, idx, weight, out, as_stream(stream));
    case 16: return launch_interp<16>(b, c, m, n, points, idx, weight, out, as_stream(stream));
    case 8: return launch_interp<8>(b, c, m, n, points, idx, weight, out, as_stream(stream));
    case 4: return launch_interp<4>(b, c, m, n, points, idx, weight, out, as_stream(stream));
    default: break;
  }
  three_interpolate_kernel<<<grid_for(n, c, b), T, 0, as_stream(stream)>>>(c, m, n, points, idx, weight, out);
  PN2_CHECK_LAUNCH("three_interpolate");
  return PN2_OK;
}

extern "C" int pn2_three_interpolate_grad(int b, int c, int n, int m, const float* grad_out, const int* idx,
                                          const float* weight, float* grad_points, void* stream) {
  PN2_RANGE("pn2_three_interpolate_grad");
  if (b <= 0 || c <= 0 || n <= 0) return PN2_OK;
  PN2_REQUIRE(grad_out && idx && weight && grad_points && m > 0 && b <= 65535,
              "three_interpolate_grad: bad arguments");
  switch (n >= 8 * m ? interp_tile_channels(c, m, 1) : 0) {  // small n/m: L2 atomics win
    case 32: return launch_interp_grad<32>(b, c, n, m, grad_out, idx, weight, grad_points, as_stream(stream));
    case 16: return launch_interp_grad<16>(b, c, n, m, grad_out, idx, weight, grad_points, as_stream(stream));
    case 8: return launch_interp_grad<8>(b, c, n, m, grad_out, idx, weight, grad_points, as_stream(stream));
    case 4: return launch_interp_grad<4>(b, c, n, m, grad_out, idx, weight, grad_points, as_stream(stream));
    default: break;
  }
  three_interpolate_grad_kernel<<<grid_for(n, c, b), T, 0, as_stream(stream)>>>(c, n, m, grad_out, idx, weight,
                                                                                grad_points);
  PN2_CHECK_LAUNCH("three_interpolate_grad");
  return PN2_OK;
}
