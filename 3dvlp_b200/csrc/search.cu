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

  for (int base = 0; base < n; base += 32) {
    const int k = base + lane;
    float x = 0.f, y = 0.f, z = 0.f;
    const bool in = k < n;
    if (in) { x = __ldg(xyz + 3 * k); y = __ldg(xyz + 3 * k + 1); z = __ldg(xyz + 3 * k + 2); }
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
constexpr int NN_TILE = 1024;  // known points per shared-memory tile (12 KB)

__global__ void __launch_bounds__(NN_THREADS)
three_nn_kernel(int n, int m, const float* __restrict__ unknown, const float* __restrict__ known,
                float* __restrict__ dist2, int* __restrict__ idx) {
  __shared__ float sk[NN_TILE * 3];
  const int scene = blockIdx.y;
  unknown += static_cast<size_t>(scene) * n * 3;
  known += static_cast<size_t>(scene) * m * 3;
  dist2 += static_cast<size_t>(scene) * n * 3;
  idx += static_cast<size_t>(scene) * n * 3;
  const int j = blockIdx.x * NN_THREADS + threadIdx.x;
  const bool live = j < n;
  float ux = 0.f, uy = 0.f, uz = 0.f;
  if (live) { ux = __ldg(unknown + 3 * j); uy = __ldg(unknown + 3 * j + 1); uz = __ldg(unknown + 3 * j + 2); }
  // interpolate_gpu.cu:32-33 keeps doubles initialised to 1e40; a float compared after promotion
  // orders exactly like the float itself, and (float)1e40 == +inf is what an unfilled slot stores.
  float b1 = INFINITY, b2 = INFINITY, b3 = INFINITY;
  int i1 = 0, i2 = 0, i3 = 0;
  for (int t0 = 0; t0 < m; t0 += NN_TILE) {
    const int tn = min(NN_TILE, m - t0);
    __syncthreads();
    for (int q = threadIdx.x; q < tn * 3; q += NN_THREADS) sk[q] = __ldg(known + 3 * t0 + q);
    __syncthreads();
    if (live) {
      for (int k = 0; k < tn; ++k) {
        const float d = sqdist3(ux - sk[3 * k], uy - sk[3 * k + 1], uz - sk[3 * k + 2]);
        if (d < b1) {
          b3 = b2; i3 = i2; b2 = b1; i2 = i1; b1 = d; i1 = t0 + k;
        } else if (d < b2) {
          b3 = b2; i3 = i2; b2 = d; i2 = t0 + k;
        } else if (d < b3) {
          b3 = d; i3 = t0 + k;
        }
      }
    }
  }
  if (live) {
    dist2[3 * j] = b1; dist2[3 * j + 1] = b2; dist2[3 * j + 2] = b3;
    idx[3 * j] = i1; idx[3 * j + 1] = i2; idx[3 * j + 2] = i3;
  }
}

}  // namespace

int ball_query_grid(int b, int n, int m, float radius, int nsample, const float* new_xyz, const float* xyz,
                    int* idx, cudaStream_t s);  // ball_query_grid.cu

}  // namespace pn2

extern "C" int pn2_ball_query(int b, int n, int m, float radius, int nsample, const float* new_xyz,
                              const float* xyz, int* idx, void* stream) {
  using namespace pn2;
  if (b <= 0 || m <= 0 || nsample <= 0) return PN2_OK;
  PN2_REQUIRE(n >= 0 && new_xyz && idx && (xyz || n == 0), "ball_query: bad arguments");
  PN2_REQUIRE(b <= 65535, "ball_query: batch %d exceeds the grid limit", b);
  const float radius2 = radius * radius;  // fp32 product, ball_query_gpu.cu:27
  // large clouds: hashed-grid search (identical output, ~100x fewer distance tests);
  // pn2_set_option("bq_grid", 0) / PN2_BQ_GRID=0 forces the linear scan
  if (option_value(OPT_BQ_GRID) && n >= 4096 && radius2 > 0.f && std::isfinite(radius) && fabsf(radius) > 1e-6f)
    return ball_query_grid(b, n, m, radius, nsample, new_xyz, xyz, idx, as_stream(stream));
  const int per_cta = BQ_CPW * BQ_WARPS;
  dim3 grid((m + per_cta - 1) / per_cta, b);
  ball_query_kernel<<<grid, BQ_WARPS * 32, 0, as_stream(stream)>>>(n, m, radius2, nsample, new_xyz, xyz, idx);
  PN2_CHECK_LAUNCH("ball_query_kernel");
  return PN2_OK;
}

extern "C" int pn2_three_nn(int b, int n, int m, const float* unknown, const float* known, float* dist2,
                            int* idx, void* stream) {
  using namespace pn2;
  if (b <= 0 || n <= 0) return PN2_OK;
  PN2_REQUIRE(m >= 0 && unknown && dist2 && idx && (known || m == 0), "three_nn: bad arguments");
  PN2_REQUIRE(b <= 65535, "three_nn: batch %d exceeds the grid limit", b);
  dim3 grid((n + NN_THREADS - 1) / NN_THREADS, b);
  three_nn_kernel<<<grid, NN_THREADS, 0, as_stream(stream)>>>(n, m, unknown, known, dist2, idx);
  PN2_CHECK_LAUNCH("three_nn_kernel");
  return PN2_OK;
}
