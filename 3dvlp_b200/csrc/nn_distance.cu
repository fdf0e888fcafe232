// nn_distance: dense pairwise cost + row/column arg-min, forward and backward, for sm_100a.
//
// The reference (utils/nn_distance.py:45-59) materialises two (B,N,M,C) copies with .repeat, a
// (B,N,M,C) difference, a (B,N,M) cost and runs two torch.min over it (~8 kernels, O(B*N*M*C) HBM
// traffic several times over).  Here every query point keeps its running minimum in registers:
// algorithmic traffic only (4C(N+M) bytes in, 12(N+M) bytes out).  The second direction is the same
// kernel with the roles of the clouds swapped (the three costs are even functions of the difference).
//
// Three launch geometries, picked per direction from the shape (all flatten (scene, point) over a 1-D
// grid, so the batch is not limited by gridDim.y and tiny clouds of many scenes share a CTA -- the
// vote loss calls this with B*1024 "scenes" of 1 x 3 points, lib/loss_helper/loss_detection.py:57-66):
//   small  partner cloud of <= 32 points: one thread per (scene, query), partner read through L1;
//   warp   few queries, larger partner cloud: one warp per (scene, query), lanes stride the partner
//          cloud and the 32 partial arg-mins are merged by (value, index) -- first index on ties;
//   tiled  many queries: two queries per thread, the partner cloud streamed through shared memory.
// The arithmetic of a cost is the same in all three: separate rounded products summed in channel
// order (torch's reduction over the last dimension), so distances and indices do not depend on the
// geometry.
#include "common.cuh"

namespace pn2 {
namespace {

constexpr int T = 128;
constexpr int TILE_FLOATS = 6144;  // 24 KB of the scanned cloud per tile
constexpr int SMALL_NB = 32;

template <int MODE>
__device__ __forceinline__ float pair_cost(float e, float delta) {
  if (MODE == PN2_NN_L2) return __fmul_rn(e, e);
  const float a = fabsf(e);
  if (MODE == PN2_NN_L1) return a;
  // huber_loss, nn_distance.py:13-30: 0.5*min(|e|,d)^2 + d*(|e| - min(|e|,d))
  const float q = fminf(a, delta);
  return __fadd_rn(__fmul_rn(0.5f, __fmul_rn(q, q)), __fmul_rn(delta, a - q));
}

// torch.min keeps the first minimal index; a NaN cost propagates like torch (the first NaN wins)
__device__ __forceinline__ bool better(float s, float best) { return s < best || (s != s && best == best); }

// lexicographic merge of two partial results: NaN beats numbers, then the smaller value, then the smaller index
__device__ __forceinline__ bool better_pair(float s, long long si, float best, long long bi) {
  const bool sn = s != s, bn = best != best;
  if (sn != bn) return sn;
  if (sn) return si < bi;
  return s < best || (s == best && si < bi);
}

template <int MODE, int C>
__device__ __forceinline__ float cost_of(const float* __restrict__ a, const float* __restrict__ b, int c, float delta) {
  float s = 0.f;
  if (C > 0) {
#pragma unroll
    for (int q = 0; q < C; ++q) s = __fadd_rn(s, pair_cost<MODE>(a[q] - b[q], delta));
  } else {
    for (int q = 0; q < c; ++q) s = __fadd_rn(s, pair_cost<MODE>(a[q] - b[q], delta));
  }
  return s;
}

// ---- small: thread per (scene, query), flattened over the grid ------------------------------------
template <int MODE, int C>
__global__ void __launch_bounds__(T)
nn_min_small_kernel(long long total, int na, int nb, int c_rt, const float* __restrict__ A,
                    const float* __restrict__ Bp, float delta, float* __restrict__ dist,
                    long long* __restrict__ idx) {
  const int c = C > 0 ? C : c_rt;
  const long long t = static_cast<long long>(blockIdx.x) * T + threadIdx.x;
  if (t >= total) return;
  const long long scene = t / na;
  const float* a = A + t * c;
  const float* bp = Bp + scene * nb * c;
  float a_reg[C > 0 ? C : 1];
  if (C > 0) {
#pragma unroll
    for (int q = 0; q < C; ++q) a_reg[q] = __ldg(a + q);
  }
  float best = INFINITY;
  long long besti = 0;
  for (int k = 0; k < nb; ++k) {
    float b_reg[C > 0 ? C : 1];
    float s;
    if (C > 0) {
#pragma unroll
      for (int q = 0; q < C; ++q) b_reg[q] = __ldg(bp + static_cast<size_t>(k) * C + q);
      s = cost_of<MODE, C>(a_reg, b_reg, c, delta);
    } else {
      s = cost_of<MODE, 0>(a, bp + static_cast<size_t>(k) * c, c, delta);
    }
    if (k == 0 || better(s, best)) { best = s; besti = k; }
  }
  dist[t] = best;
  idx[t] = besti;
}

// ---- warp: one warp per (scene, query) --------------------------------------------------------------
template <int MODE, int C>
__global__ void __launch_bounds__(T)
nn_min_warp_kernel(long long total, int na, int nb, int c_rt, const float* __restrict__ A,
                   const float* __restrict__ Bp, float delta, float* __restrict__ dist,
                   long long* __restrict__ idx) {
  const int c = C > 0 ? C : c_rt;
  const int lane = threadIdx.x & 31;
  const long long t = static_cast<long long>(blockIdx.x) * (T / 32) + (threadIdx.x >> 5);
  if (t >= total) return;  // warp-uniform
  const long long scene = t / na;
  const float* a = A + t * c;
  const float* bp = Bp + scene * nb * c;
  float a_reg[C > 0 ? C : 1];
  if (C > 0) {
#pragma unroll
    for (int q = 0; q < C; ++q) a_reg[q] = __ldg(a + q);
  }
  float best = INFINITY;
  long long besti = 0x7fffffffffffffffLL;  // a lane that saw nothing loses every merge
  bool any = false;
  for (int k = lane; k < nb; k += 32) {
    float b_reg[C > 0 ? C : 1];
    float s;
    if (C > 0) {
#pragma unroll
      for (int q = 0; q < C; ++q) b_reg[q] = __ldg(bp + static_cast<size_t>(k) * C + q);
      s = cost_of<MODE, C>(a_reg, b_reg, c, delta);
    } else {
      s = cost_of<MODE, 0>(a, bp + static_cast<size_t>(k) * c, c, delta);
    }
    if (!any || better(s, best)) { best = s; besti = k; any = true; }
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, d);
    const long long oi = __shfl_xor_sync(0xffffffffu, besti, d);
    const bool oany = __shfl_xor_sync(0xffffffffu, static_cast<int>(any), d) != 0;
    if (oany && (!any || better_pair(ob, oi, best, besti))) { best = ob; besti = oi; any = true; }
  }
  if (lane == 0) {
    dist[t] = best;
    idx[t] = besti;
  }
}

// ---- tiled: Q queries per thread, partner cloud through shared memory -------------------------------
// Grid: one CTA per (scene, chunk of Q*T queries), flattened.  A tile whose values are all finite (and
// whose queries are) cannot produce a NaN cost, so its inner loop drops the NaN clause of `better`.
constexpr int Q = 2;

template <int MODE, int C, bool NANS>
__device__ __forceinline__ void scan_tile(const float* __restrict__ tile, int tn, int t0, int c, float delta,
                                          const float (&a_reg)[Q][C > 0 ? C : 1], const float* const (&a_ptr)[Q],
                                          float (&best)[Q], int (&besti)[Q]) {
#pragma unroll 4
  for (int k = 0; k < tn; ++k) {
    const float* b = tile + k * c;
#pragma unroll
    for (int u = 0; u < Q; ++u) {
      const float s = C > 0 ? cost_of<MODE, C>(a_reg[u], b, c, delta) : cost_of<MODE, 0>(a_ptr[u], b, c, delta);
      if (NANS ? better(s, best[u]) : (s < best[u])) { best[u] = s; besti[u] = t0 + k; }
    }
  }
}

template <int MODE, int C>
__global__ void __launch_bounds__(T)
nn_min_tiled_kernel(int chunks, int na, int nb, int c_rt, const float* __restrict__ A, const float* __restrict__ Bp,
                    float delta, float* __restrict__ dist, long long* __restrict__ idx) {
  extern __shared__ float tile[];
  const int c = C > 0 ? C : c_rt;
  const long long scene = blockIdx.x / chunks;
  const int chunk = blockIdx.x - static_cast<int>(scene) * chunks;
  A += scene * na * c;
  Bp += scene * nb * c;
  float a_reg[Q][C > 0 ? C : 1];
  const float* a_ptr[Q];
  int j[Q];
  bool a_finite = true;
#pragma unroll
  for (int u = 0; u < Q; ++u) {
    j[u] = chunk * (Q * T) + u * T + threadIdx.x;
    const int jj = j[u] < na ? j[u] : na - 1;  // surplus threads shadow the last query and do not store
    a_ptr[u] = A + static_cast<size_t>(jj) * c;
    if (C > 0) {
#pragma unroll
      for (int q = 0; q < C; ++q) { a_reg[u][q] = __ldg(a_ptr[u] + q); a_finite &= isfinite(a_reg[u][q]); }
    } else {
      for (int q = 0; q < c; ++q) a_finite &= isfinite(__ldg(a_ptr[u] + q));
    }
  }
  const int tile_pts = max(1, TILE_FLOATS / c);
  float best[Q];
  int besti[Q];
  // seeded with partner 0 (`better` never replaces a NaN, and a +inf cost must still yield index 0)
#pragma unroll
  for (int u = 0; u < Q; ++u) { best[u] = INFINITY; besti[u] = 0; }
  bool first = true;
  for (int t0 = 0; t0 < nb; t0 += tile_pts) {
    const int tn = min(tile_pts, nb - t0);
    __syncthreads();
    bool fin = a_finite;
    for (int q = threadIdx.x; q < tn * c; q += T) {
      const float v = __ldg(Bp + static_cast<size_t>(t0) * c + q);
      tile[q] = v;
      fin &= isfinite(v);
    }
    const bool all_finite = __syncthreads_and(fin) != 0;
    if (first) {  // partner 0 seeds the minimum exactly like torch.min's first element
#pragma unroll
      for (int u = 0; u < Q; ++u)
        best[u] = C > 0 ? cost_of<MODE, C>(a_reg[u], tile, c, delta) : cost_of<MODE, 0>(a_ptr[u], tile, c, delta);
      first = false;
    }
    if (all_finite) scan_tile<MODE, C, false>(tile, tn, t0, c, delta, a_reg, a_ptr, best, besti);
    else scan_tile<MODE, C, true>(tile, tn, t0, c, delta, a_reg, a_ptr, best, besti);
  }
#pragma unroll
  for (int u = 0; u < Q; ++u) {
    if (j[u] < na) {
      dist[scene * na + j[u]] = best[u];
      idx[scene * na + j[u]] = besti[u];
    }
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
// (scene, point) flattened over the grid.
template <int MODE>
__global__ void __launch_bounds__(T)
nn_grad_kernel(long long total, int n, int m, int c, int side, const float* __restrict__ pc1,
               const float* __restrict__ pc2, float delta, const long long* __restrict__ idx,
               const float* __restrict__ gdist, float* __restrict__ g1, float* __restrict__ g2) {
  const int cnt = side == 0 ? n : m;
  const long long t = static_cast<long long>(blockIdx.x) * T + threadIdx.x;
  if (t >= total) return;
  const long long scene = t / cnt;
  const long long self = t - scene * cnt;
  const long long partner = idx[t];
  const float g = gdist[t];
  const size_t j = side == 0 ? self : partner, k = side == 0 ? partner : self;
  const size_t o1 = (static_cast<size_t>(scene) * n + j) * c, o2 = (static_cast<size_t>(scene) * m + k) * c;
  for (int q = 0; q < c; ++q) {
    const float e = pc1[o1 + q] - pc2[o2 + q];
    const float d = g * pair_cost_grad<MODE>(e, delta);
    if (g1) atomicAdd(g1 + o1 + q, d);
    if (g2) atomicAdd(g2 + o2 + q, -d);
  }
}

template <int MODE, int C>
int launch_min_c(int b, int na, int nb, int c, const float* A, const float* Bp, float delta, float* dist,
                 long long* idx, cudaStream_t s) {
  const long long total = static_cast<long long>(b) * na;
  if (nb <= SMALL_NB) {
    const long long grid = (total + T - 1) / T;
    PN2_REQUIRE(grid < (1LL << 31), "nn_distance: b * n = %lld too large", total);
    stat_add(STAT_NN_SMALL);
    nn_min_small_kernel<MODE, C><<<static_cast<unsigned>(grid), T, 0, s>>>(total, na, nb, c, A, Bp, delta, dist, idx);
    PN2_CHECK_LAUNCH("nn_min_small_kernel");
    return PN2_OK;
  }
  // too few queries for a thread each to fill the machine: a warp per query
  if (total < 2LL * sm_count() * T) {
    const long long grid = (total + T / 32 - 1) / (T / 32);
    stat_add(STAT_NN_SMALL);
    nn_min_warp_kernel<MODE, C><<<static_cast<unsigned>(grid), T, 0, s>>>(total, na, nb, c, A, Bp, delta, dist, idx);
    PN2_CHECK_LAUNCH("nn_min_warp_kernel");
    return PN2_OK;
  }
  const int chunks = (na + Q * T - 1) / (Q * T);
  const long long grid = static_cast<long long>(b) * chunks;
  PN2_REQUIRE(grid < (1LL << 31), "nn_distance: b * n = %lld too large", total);
  const size_t smem = static_cast<size_t>(max(1, TILE_FLOATS / c)) * c * sizeof(float);
  stat_add(STAT_NN_TILED);
  nn_min_tiled_kernel<MODE, C><<<static_cast<unsigned>(grid), T, smem, s>>>(chunks, na, nb, c, A, Bp, delta, dist, idx);
  PN2_CHECK_LAUNCH("nn_min_tiled_kernel");
  return PN2_OK;
}

template <int MODE>
int launch_min(int b, int na, int nb, int c, const float* A, const float* Bp, float delta, float* dist,
               long long* idx, cudaStream_t s) {
  if (c == 3) return launch_min_c<MODE, 3>(b, na, nb, c, A, Bp, delta, dist, idx, s);
  return launch_min_c<MODE, 0>(b, na, nb, c, A, Bp, delta, dist, idx, s);
}

}  // namespace
}  // namespace pn2

using namespace pn2;

extern "C" int pn2_nn_distance(int b, int n, int m, int c, const float* pc1, const float* pc2, int mode,
                               float delta, float* dist1, int64_t* idx1, float* dist2, int64_t* idx2,
                               void* stream) {
  PN2_RANGE("pn2_nn_distance");
  if (b <= 0) return PN2_OK;
  PN2_REQUIRE(n > 0 && m > 0 && c > 0, "nn_distance: empty cloud (n=%d m=%d c=%d); torch.min would raise", n, m, c);
  PN2_REQUIRE(c <= TILE_FLOATS, "nn_distance: c=%d too large", c);
  PN2_REQUIRE(pc1 && pc2 && dist1 && idx1 && dist2 && idx2, "nn_distance: null pointer");
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
  PN2_RANGE("pn2_nn_distance_grad");
  if (b <= 0 || n <= 0 || m <= 0 || c <= 0) return PN2_OK;
  PN2_REQUIRE(pc1 && pc2 && mode >= 0 && mode <= 2, "nn_distance_grad: bad arguments");
  cudaStream_t s = as_stream(stream);
  for (int side = 0; side < 2; ++side) {
    const float* g = side == 0 ? grad_dist1 : grad_dist2;
    const long long* idx = reinterpret_cast<const long long*>(side == 0 ? idx1 : idx2);
    if (!g) continue;
    PN2_REQUIRE(idx != nullptr, "nn_distance_grad: indices missing for side %d", side);
    const long long total = static_cast<long long>(b) * (side == 0 ? n : m);
    const long long grid = (total + T - 1) / T;
    PN2_REQUIRE(grid < (1LL << 31), "nn_distance_grad: b * n = %lld too large", total);
    const unsigned gx = static_cast<unsigned>(grid);
    if (mode == PN2_NN_L2)
      nn_grad_kernel<PN2_NN_L2><<<gx, T, 0, s>>>(total, n, m, c, side, pc1, pc2, delta, idx, g, grad_pc1, grad_pc2);
    else if (mode == PN2_NN_L1)
      nn_grad_kernel<PN2_NN_L1><<<gx, T, 0, s>>>(total, n, m, c, side, pc1, pc2, delta, idx, g, grad_pc1, grad_pc2);
    else
      nn_grad_kernel<PN2_NN_HUBER><<<gx, T, 0, s>>>(total, n, m, c, side, pc1, pc2, delta, idx, g, grad_pc1, grad_pc2);
    PN2_CHECK_LAUNCH("nn_grad_kernel");
  }
  return PN2_OK;
}
