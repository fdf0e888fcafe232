// Bandwidth-bound index ops: gather_points, group_points, three_interpolate and their scatter-add
// gradients.  One thread per output element, the fastest-varying output index on threadIdx.x so the
// index reads and the stores are coalesced; the grid covers (elements, channels, scenes) instead of
// the reference's one CTA per scene (sampling_gpu.cu:13-52, group_points_gpu.cu:13-69,
// interpolate_gpu.cu:77-148).
#include "common.cuh"

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

inline dim3 grid_for(long long q_total, int c, int b) {
  return dim3(static_cast<unsigned>((q_total + T - 1) / T), static_cast<unsigned>(c < 65535 ? c : 65535),
              static_cast<unsigned>(b));
}

}  // namespace
}  // namespace pn2

using namespace pn2;

extern "C" int pn2_gather_points(int b, int c, int n, int npoints, const float* points, const int* idx,
                                 float* out, void* stream) {
  if (b <= 0 || c <= 0 || npoints <= 0) return PN2_OK;
  PN2_REQUIRE(points && idx && out && n > 0 && b <= 65535, "gather_points: bad arguments");
  gather_rows_kernel<<<grid_for(npoints, c, b), T, 0, as_stream(stream)>>>(c, n, npoints, points, idx, out);
  PN2_CHECK_LAUNCH("gather_points");
  return PN2_OK;
}

extern "C" int pn2_gather_points_grad(int b, int c, int n, int npoints, const float* grad_out,
                                      const int* idx, float* grad_points, void* stream) {
  if (b <= 0 || c <= 0 || npoints <= 0) return PN2_OK;
  PN2_REQUIRE(grad_out && idx && grad_points && n > 0 && b <= 65535, "gather_points_grad: bad arguments");
  scatter_rows_kernel<<<grid_for(npoints, c, b), T, 0, as_stream(stream)>>>(c, n, npoints, grad_out, idx,
                                                                           grad_points);
  PN2_CHECK_LAUNCH("gather_points_grad");
  return PN2_OK;
}

extern "C" int pn2_group_points(int b, int c, int n, int npoints, int nsample, const float* points,
                                const int* idx, float* out, void* stream) {
  if (b <= 0 || c <= 0 || npoints <= 0 || nsample <= 0) return PN2_OK;
  PN2_REQUIRE(points && idx && out && n > 0 && b <= 65535, "group_points: bad arguments");
  const long long q = static_cast<long long>(npoints) * nsample;
  gather_rows_kernel<<<grid_for(q, c, b), T, 0, as_stream(stream)>>>(c, n, q, points, idx, out);
  PN2_CHECK_LAUNCH("group_points");
  return PN2_OK;
}

extern "C" int pn2_group_points_grad(int b, int c, int n, int npoints, int nsample, const float* grad_out,
                                     const int* idx, float* grad_points, void* stream) {
  if (b <= 0 || c <= 0 || npoints <= 0 || nsample <= 0) return PN2_OK;
  PN2_REQUIRE(grad_out && idx && grad_points && n > 0 && b <= 65535, "group_points_grad: bad arguments");
  const long long q = static_cast<long long>(npoints) * nsample;
  scatter_rows_kernel<<<grid_for(q, c, b), T, 0, as_stream(stream)>>>(c, n, q, grad_out, idx, grad_points);
  PN2_CHECK_LAUNCH("group_points_grad");
  return PN2_OK;
}

extern "C" int pn2_three_interpolate(int b, int c, int m, int n, const float* points, const int* idx,
                                     const float* weight, float* out, void* stream) {
  if (b <= 0 || c <= 0 || n <= 0) return PN2_OK;
  PN2_REQUIRE(points && idx && weight && out && m > 0 && b <= 65535, "three_interpolate: bad arguments");
  three_interpolate_kernel<<<grid_for(n, c, b), T, 0, as_stream(stream)>>>(c, m, n, points, idx, weight, out);
  PN2_CHECK_LAUNCH("three_interpolate");
  return PN2_OK;
}

extern "C" int pn2_three_interpolate_grad(int b, int c, int n, int m, const float* grad_out, const int* idx,
                                          const float* weight, float* grad_points, void* stream) {
  if (b <= 0 || c <= 0 || n <= 0) return PN2_OK;
  PN2_REQUIRE(grad_out && idx && weight && grad_points && m > 0 && b <= 65535,
              "three_interpolate_grad: bad arguments");
  three_interpolate_grad_kernel<<<grid_for(n, c, b), T, 0, as_stream(stream)>>>(c, n, m, grad_out, idx, weight,
                                                                                grad_points);
  PN2_CHECK_LAUNCH("three_interpolate_grad");
  return PN2_OK;
}
