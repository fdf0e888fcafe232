// nn_distance: dense pairwise cost + row/column arg-min, forward and backward, for sm_100a.
//
// The reference (utils/nn_distance.py:45-59) materialises two (B,N,M,C) copies with .repeat, a
// (B,N,M,C) difference, a (B,N,M) cost and runs two torch.min over it (~8 kernels, O(B*N*M*C) HBM
// traffic several times over).  Here one thread owns one query point, streams the other cloud
// through shared memory and keeps the running minimum in registers: algorithmic traffic only
// (4C(N+M) bytes in, 12(N+M) bytes out).  The second direction is the same kernel with the roles of
// the clouds swapped (the three costs are even functions of the difference).
#include "common.cuh"

namespace pn2 {
namespace {

constexpr int T = 128;
constexpr int TILE_FLOATS = 6144;  // 24 KB of the scanned cloud per tile

template <int MODE>
__device__ __forceinline__ float pair_cost(float e, float delta) {
  if (MODE == PN2_NN_L2) return __fmul_rn(e, e);
  const float a = fabsf(e);
  if (MODE == PN2_NN_L1) return a;
  // huber_loss, nn_distance.py:13-30: 0.5*min(|e|,d)^2 + d*(|e| - min(|e|,d))
  const float q = fminf(a, delta);
  return __fadd_rn(__fmul_rn(0.5f, __fmul_rn(q, q)), __fmul_rn(delta, a - q));
}

// For every point j of cloud A (na points): min over cloud B (nb points) of sum_c cost(A[j]-B[k]).
template <int MODE, int C>
__global__ void __launch_bounds__(T)
nn_min_kernel(int na, int nb, int c_rt, const float* __restrict__ A, const float* __restrict__ Bp,
              float delta, float* __restrict__ dist, long long* __restrict__ idx) {
  extern __shared__ float tile[];
  const int c = C > 0 ? C : c_rt;
  const int scene = blockIdx.y;
  A += static_cast<size_t>(scene) * na * c;
  Bp += static_cast<size_t>(scene) * nb * c;
  const int j = blockIdx.x * T + threadIdx.x;
  const bool live = j < na;
  float a_reg[C > 0 ? C : 1];
  if (C > 0 && live) {
#pragma unroll
    for (int q = 0; q < C; ++q) a_reg[q] = __ldg(A + static_cast<size_t>(j) * C + q);
  }
  const int tile_pts = max(1, TILE_FLOATS / c);
  float best = INFINITY;
  long long besti = 0;
  bool any = false;
  for (int t0 = 0; t0 < nb; t0 += tile_pts) {
    const int tn = min(tile_pts, nb - t0);
    __syncthreads();
    for (int q = threadIdx.x; q < tn * c; q += T) tile[q] = __ldg(Bp + static_cast<size_t>(t0) * c + q);
    __syncthreads();
    if (live) {
      for (int k = 0; k < tn; ++k) {
        float s = 0.f;
        if (C > 0) {
#pragma unroll
          for (int q = 0; q < C; ++q) s = __fadd_rn(s, pair_cost<MODE>(a_reg[q] - tile[k * C + q], delta));
        } else {
          for (int q = 0; q < c; ++q)
            s = __fadd_rn(s, pair_cost<MODE>(__ldg(A + static_cast<size_t>(j) * c + q) - tile[k * c + q], delta));
        }
        // torch.min keeps the first minimal index; a NaN cost propagates like torch (NaN wins once)
        if (!any || s < best || (s != s && best == best)) { best = s; besti = t0 + k; any = true; }
      }
    }
  }
  if (live) {
    dist[static_cast<size_t>(scene) * na + j] = best;
    idx[static_cast<size_t>(scene) * na + j] = besti;
  }
}

template <int MODE>
__device__ __forceinline__ float pair_cost_grad(float e, float delta) {
  if (MODE == PN2_NN_L2) return 2.f * e;
  if (MODE == PN2_NN_L1) return (e > 0.f) ? 1.f : ((e < 0.f) ? -1.f : 0.f);
  return fminf(fmaxf(e, -delta), delta);
}

// side 0: thread per point j of pc1, partner k = idx1[j], upstream g = grad_dist1[j]
// side 1: thread per point k of pc2, partner j = idx2[k], upstream g = grad_dist2[k]
template <int MODE>
__global__ void __launch_bounds__(T)
nn_grad_kernel(int n, int m, int c, int side, const float* __restrict__ pc1, const float* __restrict__ pc2,
               float delta, const long long* __restrict__ idx, const float* __restrict__ gdist,
               float* __restrict__ g1, float* __restrict__ g2) {
  const int scene = blockIdx.y;
  const int cnt = side == 0 ? n : m;
  const int t = blockIdx.x * T + threadIdx.x;
  if (t >= cnt) return;
  const long long partner = idx[static_cast<size_t>(scene) * cnt + t];
  const float g = gdist[static_cast<size_t>(scene) * cnt + t];
  const size_t j = side == 0 ? t : partner, k = side == 0 ? partner : t;
  const size_t o1 = (static_cast<size_t>(scene) * n + j) * c, o2 = (static_cast<size_t>(scene) * m + k) * c;
  for (int q = 0; q < c; ++q) {
    const float e = pc1[o1 + q] - pc2[o2 + q];
    const float d = g * pair_cost_grad<MODE>(e, delta);
    if (g1) atomicAdd(g1 + o1 + q, d);
    if (g2) atomicAdd(g2 + o2 + q, -d);
  }
}

template <int MODE>
int launch_min(int b, int na, int nb, int c, const float* A, const float* Bp, float delta, float* dist,
               long long* idx, cudaStream_t s) {
  dim3 grid((na + T - 1) / T, b);
  const size_t smem = static_cast<size_t>(max(1, TILE_FLOATS / c)) * c * sizeof(float);
  if (c == 3) nn_min_kernel<MODE, 3><<<grid, T, smem, s>>>(na, nb, c, A, Bp, delta, dist, idx);
  else nn_min_kernel<MODE, 0><<<grid, T, smem, s>>>(na, nb, c, A, Bp, delta, dist, idx);
  PN2_CHECK_LAUNCH("nn_min_kernel");
  return PN2_OK;
}

}  // namespace
}  // namespace pn2

using namespace pn2;

extern "C" int pn2_nn_distance(int b, int n, int m, int c, const float* pc1, const float* pc2, int mode,
                               float delta, float* dist1, int64_t* idx1, float* dist2, int64_t* idx2,
                               void* stream) {
  if (b <= 0) return PN2_OK;
  PN2_REQUIRE(n > 0 && m > 0 && c > 0, "nn_distance: empty cloud (n=%d m=%d c=%d); torch.min would raise", n, m, c);
  PN2_REQUIRE(c <= TILE_FLOATS, "nn_distance: c=%d too large", c);
  PN2_REQUIRE(pc1 && pc2 && dist1 && idx1 && dist2 && idx2 && b <= 65535, "nn_distance: bad arguments");
  PN2_REQUIRE(mode >= 0 && mode <= 2, "nn_distance: unknown mode %d", mode);
  cudaStream_t s = as_stream(stream);
  long long* i1 = reinterpret_cast<long long*>(idx1);
  long long* i2 = reinterpret_cast<long long*>(idx2);
  int rc;
  switch (mode) {
    case PN2_NN_L2:
      if ((rc = launch_min<PN2_NN_L2>(b, n, m, c, pc1, pc2, delta, dist1, i1, s))) return rc;
      return launch_min<PN2_NN_L2>(b, m, n, c, pc2, pc1, delta, dist2, i2, s);
    case PN2_NN_L1:
      if ((rc = launch_min<PN2_NN_L1>(b, n, m, c, pc1, pc2, delta, dist1, i1, s))) return rc;
      return launch_min<PN2_NN_L1>(b, m, n, c, pc2, pc1, delta, dist2, i2, s);
    default:
      if ((rc = launch_min<PN2_NN_HUBER>(b, n, m, c, pc1, pc2, delta, dist1, i1, s))) return rc;
      return launch_min<PN2_NN_HUBER>(b, m, n, c, pc2, pc1, delta, dist2, i2, s);
  }
}

extern "C" int pn2_nn_distance_grad(int b, int n, int m, int c, const float* pc1, const float* pc2, int mode,
                                    float delta, const int64_t* idx1, const int64_t* idx2,
                                    const float* grad_dist1, const float* grad_dist2, float* grad_pc1,
                                    float* grad_pc2, void* stream) {
  if (b <= 0 || n <= 0 || m <= 0 || c <= 0) return PN2_OK;
  PN2_REQUIRE(pc1 && pc2 && b <= 65535 && mode >= 0 && mode <= 2, "nn_distance_grad: bad arguments");
  cudaStream_t s = as_stream(stream);
  for (int side = 0; side < 2; ++side) {
    const float* g = side == 0 ? grad_dist1 : grad_dist2;
    const long long* idx = reinterpret_cast<const long long*>(side == 0 ? idx1 : idx2);
    if (!g) continue;
    PN2_REQUIRE(idx != nullptr, "nn_distance_grad: indices missing for side %d", side);
    const int cnt = side == 0 ? n : m;
    dim3 grid((cnt + T - 1) / T, b);
    if (mode == PN2_NN_L2)
      nn_grad_kernel<PN2_NN_L2><<<grid, T, 0, s>>>(n, m, c, side, pc1, pc2, delta, idx, g, grad_pc1, grad_pc2);
    else if (mode == PN2_NN_L1)
      nn_grad_kernel<PN2_NN_L1><<<grid, T, 0, s>>>(n, m, c, side, pc1, pc2, delta, idx, g, grad_pc1, grad_pc2);
    else
      nn_grad_kernel<PN2_NN_HUBER><<<grid, T, 0, s>>>(n, m, c, side, pc1, pc2, delta, idx, g, grad_pc1, grad_pc2);
    PN2_CHECK_LAUNCH("nn_grad_kernel");
  }
  return PN2_OK;
}
