// Neighbour searches: ball query and three-nearest-neighbours, for sm_100a.
//
// ball_query replaces query_ball_point_kernel (reference ball_query_gpu.cu:14-49: one CTA per scene,
// one thread per centre, a serial scan of all n points from global memory with a divergent early
// exit).  Here a warp owns CPW centres (kept in registers) and its 32 lanes test 32 consecutive
// points per step, so the scan order of the reference -- ascending point index -- becomes
// ballot + popc ordered compaction: the t-th set bit of the ballot is the t-th hit.  Each point is
// loaded once per warp step and tested against CPW centres; the exit test is warp-uniform.
// The grid is (centres / (CPW * warps), scenes): thousands of CTAs instead of B.
//
// three_nn replaces three_nn_kernel (interpolate_gpu.cu:14-64): thread per unknown point, the known
// set staged through shared memory in tiles, same strict-`<` cascade so ties keep the earlier index.
#include <cstdlib>

#include "common.cuh"

namespace pn2 {
namespace {

constexpr int BQ_CPW = 4;      // centres per warp
constexpr int BQ_WARPS = 8;    // warps per CTA
constexpr int BQ_AHEAD = 4;    // iterations whose points are in flight

__global__ void __launch_bounds__(BQ_WARPS * 32)
ball_query_kernel(int n, int m, float radius2, int nsample, const float* __restrict__ new_xyz,
                  const float* __restrict__ xyz, int* __restrict__ idx) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int scene = blockIdx.y;
  const int c0 = (blockIdx.x * BQ_WARPS + warp) * BQ_CPW;
  if (c0 >= m) return;
  xyz += static_cast<size_t>(scene) * n * 3;
  new_xyz += static_cast<size_t>(scene) * m * 3;
  idx += static_cast<size_t>(scene) * m * nsample;

  float cx[BQ_CPW], cy[BQ_CPW], cz[BQ_CPW];
  int cnt[BQ_CPW], first[BQ_CPW];
#pragma unroll
  for (int c = 0; c < BQ_CPW; ++c) {
    const int j = min(c0 + c, m - 1);
    cx[c] = __ldg(new_xyz + 3 * j); cy[c] = __ldg(new_xyz + 3 * j + 1); cz[c] = __ldg(new_xyz + 3 * j + 2);
    cnt[c] = (c0 + c < m) ? 0 : nsample;  // out-of-range centres are "already full"
    first[c] = 0;
  }
  const uint32_t lt = (1u << lane) - 1u;

  // The scan stops as soon as every centre of the warp is full, so the compiler cannot hoist the next
  // iteration's loads above the exit test: the loop ran one global round trip per 32 points (45 % of its
  // stall samples).  The points of the next BQ_AHEAD iterations are requested ahead, in registers.
  float px[BQ_AHEAD], py[BQ_AHEAD], pz[BQ_AHEAD];
#pragma unroll
  for (int d = 0; d < BQ_AHEAD; ++d) {
    const int k = d * 32 + lane;
    px[d] = py[d] = pz[d] = 0.f;
    if (k < n) { px[d] = __ldg(xyz + 3 * k); py[d] = __ldg(xyz + 3 * k + 1); pz[d] = __ldg(xyz + 3 * k + 2); }
  }
  for (int base = 0; base < n; base += 32) {
    const int k = base + lane;
    const float x = px[0], y = py[0], z = pz[0];
    const bool in = k < n;
#pragma unroll
    for (int d = 0; d + 1 < BQ_AHEAD; ++d) { px[d] = px[d + 1]; py[d] = py[d + 1]; pz[d] = pz[d + 1]; }
    {
      const int ka = k + BQ_AHEAD * 32;
      px[BQ_AHEAD - 1] = py[BQ_AHEAD - 1] = pz[BQ_AHEAD - 1] = 0.f;
      if (ka < n) {
        px[BQ_AHEAD - 1] = __ldg(xyz + 3 * ka); py[BQ_AHEAD - 1] = __ldg(xyz + 3 * ka + 1);
        pz[BQ_AHEAD - 1] = __ldg(xyz + 3 * ka + 2);
      }
    }
    bool all_full = true;
#pragma unroll
    for (int c = 0; c < BQ_CPW; ++c) {
      // ball_query_gpu.cu:36-38: (new - p), strict `<`; NaN never hits
      const float d2 = sqdist3(cx[c] - x, cy[c] - y, cz[c] - z);
      const bool hit = in && (d2 < radius2);
      const uint32_t mask = __ballot_sync(0xffffffffu, hit);
      if (mask != 0u && cnt[c] < nsample) {
        if (cnt[c] == 0) first[c] = base + __ffs(mask) - 1;
        const int pos = cnt[c] + __popc(mask & lt);
        if (hit && pos < nsample) idx[static_cast<size_t>(c0 + c) * nsample + pos] = k;
        cnt[c] += __popc(mask);
      }
      all_full = all_full && (cnt[c] >= nsample);
    }
    if (all_full) break;
  }
  // padding: slots [cnt, nsample) repeat the first hit; an empty ball is a row of zeros
  // (ball_query_gpu.cu:40-44 + the zero-initialised output of ball_query.cpp:24-26)
#pragma unroll
  for (int c = 0; c < BQ_CPW; ++c) {
    if (c0 + c < m) {
      const int have = min(cnt[c], nsample);
      for (int s = have + lane; s < nsample; s += 32)
        idx[static_cast<size_t>(c0 + c) * nsample + s] = first[c];
    }
  }
}

constexpr int NN_THREADS = 128;
constexpr int NN_TILE = 1024;  // known points per shared-memory tile (16 KB as float4)

// U unknown points per thread: one broadcast LDS.128 of a known point feeds U distance tests.
template <int U>
__global__ void __launch_bounds__(NN_THREADS)
three_nn_kernel(int n, int m, const float* __restrict__ unknown, const float* __restrict__ known,
                float* __restrict__ dist2, int* __restrict__ idx) {
  __shared__ float4 sk[NN_TILE];
  const int scene = blockIdx.y;
  unknown += static_cast<size_t>(scene) * n * 3;
  known += static_cast<size_t>(scene) * m * 3;
  dist2 += static_cast<size_t>(scene) * n * 3;
  idx += static_cast<size_t>(scene) * n * 3;
  float ux[U], uy[U], uz[U];
  // interpolate_gpu.cu:32-33 keeps doubles initialised to 1e40; a float compared after promotion
  // orders exactly like the float itself, and (float)1e40 == +inf is what an unfilled slot stores.
  float b1[U], b2[U], b3[U];
  int i1[U], i2[U], i3[U];
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const int j = (blockIdx.x * U + u) * NN_THREADS + threadIdx.x;
    const bool live = j < n;
    ux[u] = live ? __ldg(unknown + 3 * j) : 0.f;
    uy[u] = live ? __ldg(unknown + 3 * j + 1) : 0.f;
    uz[u] = live ? __ldg(unknown + 3 * j + 2) : 0.f;
    b1[u] = b2[u] = b3[u] = INFINITY;
    i1[u] = i2[u] = i3[u] = 0;
  }
  for (int t0 = 0; t0 < m; t0 += NN_TILE) {
    const int tn = min(NN_TILE, m - t0);
    __syncthreads();
    for (int q = threadIdx.x; q < tn; q += NN_THREADS)
      sk[q] = make_float4(__ldg(known + 3 * (t0 + q)), __ldg(known + 3 * (t0 + q) + 1), __ldg(known + 3 * (t0 + q) + 2), 0.f);
    __syncthreads();
    for (int k = 0; k < tn; ++k) {
      const float4 p = sk[k];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const float d = sqdist3(ux[u] - p.x, uy[u] - p.y, uz[u] - p.z);
        if (d < b3[u]) {  // b1 <= b2 <= b3 always: one compare rejects almost every candidate
          if (d < b1[u]) {
            b3[u] = b2[u]; i3[u] = i2[u]; b2[u] = b1[u]; i2[u] = i1[u]; b1[u] = d; i1[u] = t0 + k;
          } else if (d < b2[u]) {
            b3[u] = b2[u]; i3[u] = i2[u]; b2[u] = d; i2[u] = t0 + k;
          } else {
            b3[u] = d; i3[u] = t0 + k;
          }
        }
      }
    }
  }
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const int j = (blockIdx.x * U + u) * NN_THREADS + threadIdx.x;
    if (j < n) {
      dist2[3 * j] = b1[u]; dist2[3 * j + 1] = b2[u]; dist2[3 * j + 2] = b3[u];
      idx[3 * j] = i1[u]; idx[3 * j + 1] = i2[u]; idx[3 * j + 2] = i3[u];
    }
  }
}

}  // namespace

int ball_query_grid(int b, int n, int m, float radius, int nsample, const float* new_xyz, const float* xyz,
                    int* idx, cudaStream_t s);  // ball_query_grid.cu
int ball_query_small(int b, int n, int m, float radius, int nsample, const float* new_xyz, const float* xyz,
                     int* idx, cudaStream_t s);  // ball_query_grid.cu: grid in shared memory, 2048 <= n < 4096
bool ball_query_small_supported(int n);

}  // namespace pn2

extern "C" int pn2_ball_query(int b, int n, int m, float radius, int nsample, const float* new_xyz,
                              const float* xyz, int* idx, void* stream) {
  PN2_RANGE("pn2_ball_query");
  using namespace pn2;
  if (b <= 0 || m <= 0 || nsample <= 0) return PN2_OK;
  PN2_REQUIRE(n >= 0 && new_xyz && idx && (xyz || n == 0), "ball_query: bad arguments");
  PN2_REQUIRE(b <= 65535, "ball_query: batch %d exceeds the grid limit", b);
  const float radius2 = radius * radius;  // fp32 product, ball_query_gpu.cu:27
  // large clouds: hashed-grid search (identical output, ~100x fewer distance tests);
  // pn2_set_option("bq_grid", 0) / PN2_BQ_GRID=0 forces the linear scan
  if (option_value(OPT_BQ_GRID) && n >= 4096 && radius2 > 0.f && std::isfinite(radius) && fabsf(radius) > 1e-6f) {
    stat_add(STAT_BQ_GRID);
    return ball_query_grid(b, n, m, radius, nsample, new_xyz, xyz, idx, as_stream(stream));
  }
  // mid-sized clouds (SA2): the same search out of a grid built in shared memory, one launch; a centre
  // whose stencil would hold more than n/8 candidates scans linearly inside that kernel
  if (option_value(OPT_BQ_GRID) && ball_query_small_supported(n) && radius2 > 0.f && std::isfinite(radius) &&
      fabsf(radius) > 1e-6f) {
    stat_add(STAT_BQ_SMALL_GRID);
    return ball_query_small(b, n, m, radius, nsample, new_xyz, xyz, idx, as_stream(stream));
  }
  stat_add(STAT_BQ_SCAN);
  const int per_cta = BQ_CPW * BQ_WARPS;
  dim3 grid((m + per_cta - 1) / per_cta, b);
  ball_query_kernel<<<grid, BQ_WARPS * 32, 0, as_stream(stream)>>>(n, m, radius2, nsample, new_xyz, xyz, idx);
  PN2_CHECK_LAUNCH("ball_query_kernel");
  return PN2_OK;
}

extern "C" int pn2_three_nn(int b, int n, int m, const float* unknown, const float* known, float* dist2,
                            int* idx, void* stream) {
  PN2_RANGE("pn2_three_nn");
  using namespace pn2;
  if (b <= 0 || n <= 0) return PN2_OK;
  PN2_REQUIRE(m >= 0 && unknown && dist2 && idx && (known || m == 0), "three_nn: bad arguments");
  PN2_REQUIRE(b <= 65535, "three_nn: batch %d exceeds the grid limit", b);
  // two unknowns per thread once there are enough of them to fill the machine twice over
  if (static_cast<long long>(n) * b >= 4LL * sm_count() * NN_THREADS * 2) {
    dim3 grid((n + 2 * NN_THREADS - 1) / (2 * NN_THREADS), b);
    three_nn_kernel<2><<<grid, NN_THREADS, 0, as_stream(stream)>>>(n, m, unknown, known, dist2, idx);
  } else {
    dim3 grid((n + NN_THREADS - 1) / NN_THREADS, b);
    three_nn_kernel<1><<<grid, NN_THREADS, 0, as_stream(stream)>>>(n, m, unknown, known, dist2, idx);
  }
  PN2_CHECK_LAUNCH("three_nn_kernel");
  return PN2_OK;
}
