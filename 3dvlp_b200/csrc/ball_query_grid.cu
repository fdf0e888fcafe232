// Ball query for large clouds through a hashed uniform grid -- same output as the linear scan,
// two orders of magnitude fewer distance tests.
//
// Reference semantics (ball_query_gpu.cu:14-49): for every centre, the FIRST nsample point indices in
// ascending order with d2 < r^2, padded with the first hit; an empty ball is a row of zeros.  The
// reference (and csrc/search.cu for small clouds) scans all n points per centre: N*npoint = 82 M
// tests per scene for SA1.  Here:
//   1. points are counting-sorted into buckets of a hashed grid with cell edge h = 1.001*r
//      (count -> per-scene exclusive scan -> scatter; x,y,z,index packed as float4 for 16-byte loads);
//   2. a warp owns a centre: lanes 0..26 fetch the ranges of the 27 neighbouring cells in parallel,
//      a warp scan flattens them into one candidate list, lanes stride over it with LDG.128, discard
//      hash-collision strangers (the candidate's own cell must equal the probed cell, which also
//      guarantees no point is seen twice) and apply the SAME fp32 test as the reference
//      (d2 = fma(dz,dz,fma(dx,dx,dy*dy)) < r*r, centre minus point), so the hit SET is bit-identical;
//   3. the reference's ORDER is restored by ranking the hits by point index (rank = number of hits with
//      a smaller index; indices are unique), slot[rank] = index for rank < nsample, padding with the
//      rank-0 hit.
// The 0.1 % slack in h makes the 27-cell stencil a superset of the ball under fp32 rounding of the
// cell coordinates.  A centre with more than BQ_CAP hits falls back to the ordered linear scan inside
// the same kernel (rare: dense clusters), so the result is exact for every input.
#include "common.cuh"

namespace pn2 {
namespace {

constexpr int BQ_CAP = 256;      // hits buffered per centre before the linear-scan fallback
constexpr int BQ_QWARPS = 8;     // centres (warps) per CTA in the query kernel

__device__ __forceinline__ uint32_t cell_hash(int ix, int iy, int iz, uint32_t mask) {
  return ((static_cast<uint32_t>(ix) * 73856093u) ^ (static_cast<uint32_t>(iy) * 19349663u) ^
          (static_cast<uint32_t>(iz) * 83492791u)) & mask;
}
__device__ __forceinline__ int cell_of(float x, float inv_h) { return static_cast<int>(floorf(x * inv_h)); }

__global__ void bq_count_kernel(int n, float inv_h, uint32_t mask, const float* __restrict__ xyz,
                                int* __restrict__ cnt) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const int scene = blockIdx.y;
  const float* p = xyz + (static_cast<size_t>(scene) * n + k) * 3;
  const uint32_t bkt = cell_hash(cell_of(__ldg(p), inv_h), cell_of(__ldg(p + 1), inv_h), cell_of(__ldg(p + 2), inv_h), mask);
  atomicAdd(cnt + static_cast<size_t>(scene) * (mask + 2) + bkt, 1);
}

// one CTA per scene: exclusive scan of the T bucket counts (start[T] = n), cursor = start.
// Coalesced: the 1024 threads sweep a group of 65 536 buckets as 16 chunks of 4096 with lane-contiguous
// 4-element pieces; the 16 per-thread chunk sums are scanned across the block together (one shuffle
// scan + two shared-memory hops for all 16 chunks), then the pieces are re-read and the exclusive
// prefixes written back with the same coalesced pattern.
__global__ void __launch_bounds__(1024) bq_scan_kernel(int n, uint32_t T, int* __restrict__ start,
                                                        int* __restrict__ cursor) {
  constexpr int NCH = 16;              // chunks per group
  __shared__ int wsum[NCH][33];        // per chunk: inclusive scan of the 32 warp sums
  __shared__ int cpre[NCH + 1];        // exclusive prefix of the chunk totals (+ group total)
  const int scene = blockIdx.x;
  int* s = start + static_cast<size_t>(scene) * (T + 1);
  int* c = cursor + static_cast<size_t>(scene) * T;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int group_carry = 0;
  for (uint32_t g0 = 0; g0 < T; g0 += NCH * 4096) {
    int ts[NCH], inc[NCH];
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) {
      const int* p = s + g0 + ch * 4096 + tid * 4;
      ts[ch] = p[0] + p[1] + p[2] + p[3];
      inc[ch] = ts[ch];
    }
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
#pragma unroll
      for (int ch = 0; ch < NCH; ++ch) {
        const int t = __shfl_up_sync(0xffffffffu, inc[ch], o);
        if (lane >= o) inc[ch] += t;
      }
    }
    if (lane == 31) {
#pragma unroll
      for (int ch = 0; ch < NCH; ++ch) wsum[ch][warp] = inc[ch];
    }
    __syncthreads();
    if (warp < NCH) {  // warp `ch` scans the 32 warp sums of chunk `ch`
      int w = wsum[warp][lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, w, o);
        if (lane >= o) w += t;
      }
      wsum[warp][lane] = w;
    }
    __syncthreads();
    if (tid == 0) {
      int run = 0;
      for (int ch = 0; ch < NCH; ++ch) { cpre[ch] = run; run += wsum[ch][31]; }
      cpre[NCH] = run;
    }
    __syncthreads();
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) {
      int run = group_carry + cpre[ch] + (warp ? wsum[ch][warp - 1] : 0) + inc[ch] - ts[ch];
      int* p = s + g0 + ch * 4096 + tid * 4;
      int* q = c + g0 + ch * 4096 + tid * 4;
      const int v0 = p[0], v1 = p[1], v2 = p[2], v3 = p[3];
      p[0] = run; q[0] = run; run += v0;
      p[1] = run; q[1] = run; run += v1;
      p[2] = run; q[2] = run; run += v2;
      p[3] = run; q[3] = run; (void)v3;
    }
    group_carry += cpre[NCH];
    __syncthreads();
  }
  if (tid == 0) s[T] = n;
}

__global__ void bq_scatter_kernel(int n, float inv_h, uint32_t mask, const float* __restrict__ xyz,
                                  int* __restrict__ cursor, float4* __restrict__ sorted) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const int scene = blockIdx.y;
  const float* p = xyz + (static_cast<size_t>(scene) * n + k) * 3;
  const float x = __ldg(p), y = __ldg(p + 1), z = __ldg(p + 2);
  const uint32_t bkt = cell_hash(cell_of(x, inv_h), cell_of(y, inv_h), cell_of(z, inv_h), mask);
  const int pos = atomicAdd(cursor + static_cast<size_t>(scene) * (mask + 1) + bkt, 1);
  sorted[static_cast<size_t>(scene) * n + pos] = make_float4(x, y, z, __int_as_float(k));
}

// One centre, one warp.  SMEM: `st` / `pts` are the shared-memory grid of bq_small_kernel, otherwise the
// global workspace of ball_query_grid.  A centre whose 27-cell stencil holds more than `max_total`
// candidates, or more than BQ_CAP hits, takes the ordered linear scan over the original cloud.
template <bool SMEM>
__device__ __forceinline__ void bq_query_centre(int n, int m, int scene, int j, float radius2, float inv_h, uint32_t mask,
                                                int nsample, int max_total, const float* __restrict__ new_xyz,
                                                const float* __restrict__ xyz, const int* st, const float4* pts,
                                                int* __restrict__ idx, int* s_begin, int* s_off, int* s_hits) {
  const int lane = threadIdx.x & 31;
  const float* c = new_xyz + (static_cast<size_t>(scene) * m + j) * 3;
  const float cx = __ldg(c), cy = __ldg(c + 1), cz = __ldg(c + 2);
  int* row = idx + (static_cast<size_t>(scene) * m + j) * nsample;
  const int bx = cell_of(cx, inv_h), by = cell_of(cy, inv_h), bz = cell_of(cz, inv_h);

  // lanes 0..26: range of one neighbouring cell each.  Every BUCKET is scanned once: when two of the 27
  // cells hash to the same bucket only the lowest lane keeps it.  A point within the radius lives in one
  // of the 27 cells (cell edge > radius), hence in exactly one of the scanned buckets, so it is found
  // exactly once; strangers that share a bucket through a hash collision fail the distance test like
  // any other far point.  (No per-candidate "is it in the probed cell" test: that was 19 % of the
  // kernel's instructions.)
  int begin = 0, count = 0;
  uint32_t bkt = 0x80000000u | static_cast<uint32_t>(lane);  // lanes 27..31: unique, matches nothing
  if (lane < 27) bkt = cell_hash(bx + lane % 3 - 1, by + (lane / 3) % 3 - 1, bz + lane / 9 - 1, mask);
  const uint32_t same = __match_any_sync(0xffffffffu, bkt);
  if (lane < 27 && (same & ((1u << lane) - 1u)) == 0u) {
    begin = SMEM ? st[bkt] : __ldg(st + bkt);
    count = (SMEM ? st[bkt + 1] : __ldg(st + bkt + 1)) - begin;
  }
  int incl = count;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  const int total = __shfl_sync(0xffffffffu, incl, 31);  // lanes 27..31 hold no range
  // the NON-EMPTY ranges, compacted: their exclusive offsets are strictly increasing
  const uint32_t lt = (1u << lane) - 1u;
  const uint32_t nonempty = __ballot_sync(0xffffffffu, count > 0);
  const int nr = __popc(nonempty);
  if (count > 0) {
    const int slot = __popc(nonempty & lt);
    s_begin[slot] = begin;
    s_off[slot] = incl - count;
  }
  __syncwarp();
  const int my_off = lane < nr ? s_off[lane] : 0x7fffffff;

  int nhit = 0;
  const bool use_grid = total <= max_total;  // warp-uniform
  if (!use_grid) nhit = BQ_CAP + 1;          // -> the ordered linear scan below
  int ranges_before = 0;                     // ranges that start before q0
  for (int q0 = 0; use_grid && q0 < total; q0 += 32) {
    const int q = q0 + lane;
    // which range holds candidate q: every range that starts inside this chunk of 32 candidates sets one bit
    // (one redux.or for the warp); a lane's range is the last one that starts at or before its candidate --
    // a population count instead of a five-step binary search over shared memory per candidate
    const uint32_t rel = static_cast<uint32_t>(my_off - q0);
    const uint32_t marks = __reduce_or_sync(0xffffffffu, rel < 32u ? (1u << rel) : 0u);
    bool hit = false;
    int pidx = 0;
    if (q < total) {
      const int lo = ranges_before + __popc(marks & (0xffffffffu >> (31 - lane))) - 1;
      const float4* pp = pts + s_begin[lo] + (q - s_off[lo]);
      const float4 p = SMEM ? *pp : __ldg(pp);
      hit = sqdist3(cx - p.x, cy - p.y, cz - p.z) < radius2;
      pidx = __float_as_int(p.w);
    }
    ranges_before += __popc(marks);
    const uint32_t mk = __ballot_sync(0xffffffffu, hit);
    if (hit) {
      const int pos = nhit + __popc(mk & lt);
      if (pos < BQ_CAP) s_hits[pos] = pidx;
    }
    nhit += __popc(mk);
  }
  __syncwarp();

  if (nhit == 0) {
    for (int s = lane; s < nsample; s += 32) row[s] = 0;  // empty ball: zeros (ball_query.cpp:24-26)
    return;
  }
  if (nhit <= BQ_CAP) {
    // rank the hits by point index: slot[rank] = index, padding with the rank-0 hit
    int first = 0x7fffffff;
    if (lane < 4) s_hits[nhit + lane] = 0x7fffffff;  // pad to a multiple of four: never "smaller"
    __syncwarp();
    const int4* h4 = reinterpret_cast<const int4*>(s_hits);
    const int n4 = (nhit + 3) >> 2;
    for (int i = lane; i < nhit; i += 32) {
      const int mine = s_hits[i];
      int rank = 0;
      for (int t = 0; t < n4; ++t) {  // one broadcast LDS.128 per four hits
        const int4 h = h4[t];
        rank += (h.x < mine) + (h.y < mine) + (h.z < mine) + (h.w < mine);
      }
      if (rank < nsample) row[rank] = mine;
      if (rank == 0) first = mine;
    }
    first = __reduce_min_sync(0xffffffffu, first);
    for (int s = nhit + lane; s < nsample; s += 32) row[s] = first;
    return;
  }
  // ---- overflow: ordered linear scan of the whole scene for this centre (exact, rare) -----------
  const float* P = xyz + static_cast<size_t>(scene) * n * 3;
  int cnt = 0, first = 0;
  for (int base = 0; base < n && cnt < nsample; base += 32) {
    const int k = base + lane;
    bool hit = false;
    if (k < n) hit = sqdist3(cx - __ldg(P + 3 * k), cy - __ldg(P + 3 * k + 1), cz - __ldg(P + 3 * k + 2)) < radius2;
    const uint32_t mk = __ballot_sync(0xffffffffu, hit);
    if (mk) {
      if (cnt == 0) first = base + __ffs(mk) - 1;
      const int pos = cnt + __popc(mk & lt);
      if (hit && pos < nsample) row[pos] = k;
      cnt += __popc(mk);
    }
  }
  for (int s = min(cnt, nsample) + lane; s < nsample; s += 32) row[s] = first;
}

__global__ void __launch_bounds__(BQ_QWARPS * 32)
bq_query_kernel(int n, int m, float radius2, float inv_h, uint32_t mask, int nsample,
                const float* __restrict__ new_xyz, const float* __restrict__ xyz, const int* __restrict__ start,
                const float4* __restrict__ sorted, int* __restrict__ idx) {
  __shared__ int s_begin[BQ_QWARPS][27], s_off[BQ_QWARPS][28];
  __shared__ __align__(16) int s_hits[BQ_QWARPS][BQ_CAP + 4];
  const int warp = threadIdx.x >> 5;
  const int scene = blockIdx.y;
  const int j = blockIdx.x * BQ_QWARPS + warp;
  if (j >= m) return;  // warp-uniform
  bq_query_centre<false>(n, m, scene, j, radius2, inv_h, mask, nsample, 0x7fffffff, new_xyz, xyz,
                         start + static_cast<size_t>(scene) * (mask + 2), sorted + static_cast<size_t>(scene) * n, idx,
                         s_begin[warp], s_off[warp], s_hits[warp]);
}

// ---- small clouds (2048 <= n <= 4096: SA2 of the backbone) -------------------------------------------
// The linear scan of search.cu tests all n points per centre and only stops early when a ball fills up;
// SA2's balls hold 7 of 32 samples, so every centre pays the full scan (12.8 M warp-instructions per
// step, the third-largest kernel by instruction count).  For clouds this small the whole grid fits in
// shared memory: a CTA of 1024 threads counting-sorts its scene into SM_T hashed buckets (count ->
// block scan -> scatter, all in shared memory), then each of its 32 warps answers SM_CPW centres out
// of it with the same per-centre routine as the large-cloud kernel.  A centre whose stencil holds more
// than n/8 candidates (large radius relative to the scene: the grid would not pay) takes the ordered
// linear scan instead, so the choice is made per centre, on the device, with identical output.
constexpr int SM_T = 2048;       // buckets of the shared-memory grid
constexpr int SM_WARPS = 32;
constexpr int SM_CPW = 8;        // centres per warp
constexpr int SM_MAXN = 4096;

__global__ void __launch_bounds__(SM_WARPS * 32)
bq_small_kernel(int n, int m, float radius2, float inv_h, int nsample, const float* __restrict__ new_xyz,
                const float* __restrict__ xyz, int* __restrict__ idx) {
  extern __shared__ __align__(16) unsigned char sm_raw[];
  float4* s_pts = reinterpret_cast<float4*>(sm_raw);                       // [n]
  int* s_start = reinterpret_cast<int*>(s_pts + n);                         // [SM_T + 1]
  int* s_cur = s_start + SM_T + 1;                                          // [SM_T]
  int* s_wsum = s_cur + SM_T;                                               // [32]
  int* s_begin = s_wsum + 32;                                               // [SM_WARPS][27]
  int* s_off = s_begin + SM_WARPS * 27;                                     // [SM_WARPS][28]
  int* s_hits = s_off + SM_WARPS * 28 + 3;                                  // [SM_WARPS][BQ_CAP + 4], 16-byte aligned
  s_hits = reinterpret_cast<int*>((reinterpret_cast<uintptr_t>(s_hits)) & ~static_cast<uintptr_t>(15));
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int scene = blockIdx.y;
  const float* P = xyz + static_cast<size_t>(scene) * n * 3;
  constexpr uint32_t mask = SM_T - 1;

  for (int i = tid; i < SM_T; i += SM_WARPS * 32) s_cur[i] = 0;
  __syncthreads();
  float px[4], py[4], pz[4];
  uint32_t pb[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int k = tid + u * SM_WARPS * 32;
    if (k < n) {
      px[u] = __ldg(P + 3 * k); py[u] = __ldg(P + 3 * k + 1); pz[u] = __ldg(P + 3 * k + 2);
      pb[u] = cell_hash(cell_of(px[u], inv_h), cell_of(py[u], inv_h), cell_of(pz[u], inv_h), mask);
      atomicAdd(s_cur + pb[u], 1);
    }
  }
  __syncthreads();
  {  // exclusive scan of the SM_T counts: two per thread, warp scan, scan of the 32 warp sums
    const int c0 = s_cur[2 * tid], c1 = s_cur[2 * tid + 1];
    int inc = c0 + c1;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    if (lane == 31) s_wsum[warp] = inc;
    __syncthreads();
    if (warp == 0) {
      int w = s_wsum[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, w, o);
        if (lane >= o) w += t;
      }
      s_wsum[lane] = w;
    }
    __syncthreads();
    const int excl = (warp ? s_wsum[warp - 1] : 0) + inc - (c0 + c1);
    s_start[2 * tid] = excl; s_start[2 * tid + 1] = excl + c0;
    s_cur[2 * tid] = excl; s_cur[2 * tid + 1] = excl + c0;
    if (tid == 0) s_start[SM_T] = n;
  }
  __syncthreads();
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int k = tid + u * SM_WARPS * 32;
    if (k < n) s_pts[atomicAdd(s_cur + pb[u], 1)] = make_float4(px[u], py[u], pz[u], __int_as_float(k));
  }
  __syncthreads();

  const int j0 = blockIdx.x * (SM_WARPS * SM_CPW);
  for (int i = 0; i < SM_CPW; ++i) {
    const int j = j0 + i * SM_WARPS + warp;
    if (j >= m) break;  // warp-uniform
    bq_query_centre<true>(n, m, scene, j, radius2, inv_h, mask, nsample, n >> 3, new_xyz, xyz, s_start, s_pts, idx,
                          s_begin + warp * 27, s_off + warp * 28, s_hits + warp * (BQ_CAP + 4));
    __syncwarp();
  }
}

}  // namespace

// Called by pn2_ball_query (search.cu) for large clouds.  Workspace comes from the stream-ordered
// allocator, so the call stays asynchronous, stateless and re-entrant like the reference launcher.
int ball_query_grid(int b, int n, int m, float radius, int nsample, const float* new_xyz, const float* xyz,
                    int* idx, cudaStream_t s) {
  const uint32_t T = n <= 131072 ? 65536u : 262144u;
  const uint32_t mask = T - 1;
  const float h = fabsf(radius) * 1.001f;
  const float inv_h = 1.0f / h;
  const float radius2 = radius * radius;
  const size_t start_bytes = static_cast<size_t>(b) * (T + 1) * sizeof(int);
  const size_t cursor_bytes = static_cast<size_t>(b) * T * sizeof(int);
  const size_t sorted_off = (start_bytes + cursor_bytes + 255) & ~static_cast<size_t>(255);
  const size_t bytes = sorted_off + static_cast<size_t>(b) * n * sizeof(float4);
  keep_workspace_pool();  // abi.cu: freed workspace stays cached in the device's default pool
  uint8_t* ws = nullptr;
  PN2_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&ws), bytes, s));
  int* start = reinterpret_cast<int*>(ws);
  int* cursor = reinterpret_cast<int*>(ws + start_bytes);
  float4* sorted = reinterpret_cast<float4*>(ws + sorted_off);
  int rc = PN2_OK;
  do {
    if (cudaMemsetAsync(start, 0, start_bytes, s) != cudaSuccess) { rc = cuda_fail(cudaGetLastError(), "memset"); break; }
    dim3 gp((n + 255) / 256, b);
    bq_count_kernel<<<gp, 256, 0, s>>>(n, inv_h, mask, xyz, start);
    bq_scan_kernel<<<b, 1024, 0, s>>>(n, T, start, cursor);
    bq_scatter_kernel<<<gp, 256, 0, s>>>(n, inv_h, mask, xyz, cursor, sorted);
    dim3 gq((m + BQ_QWARPS - 1) / BQ_QWARPS, b);
    bq_query_kernel<<<gq, BQ_QWARPS * 32, 0, s>>>(n, m, radius2, inv_h, mask, nsample, new_xyz, xyz, start, sorted, idx);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) rc = cuda_fail(e, "ball_query_grid kernels");
  } while (0);
  cudaFreeAsync(ws, s);
  return rc;
}

// Called by pn2_ball_query (search.cu) for clouds of 2048..4096 points: one launch, grid in shared memory.
int ball_query_small(int b, int n, int m, float radius, int nsample, const float* new_xyz, const float* xyz,
                     int* idx, cudaStream_t s) {
  const float h = fabsf(radius) * 1.001f;
  const size_t smem = static_cast<size_t>(n) * sizeof(float4) +
                      (static_cast<size_t>(SM_T) * 2 + 1 + 32 + SM_WARPS * 55 + 4 + SM_WARPS * (BQ_CAP + 4)) * sizeof(int) + 16;
  PN2_CUDA(cudaFuncSetAttribute(bq_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
  dim3 grid((m + SM_WARPS * SM_CPW - 1) / (SM_WARPS * SM_CPW), b);
  bq_small_kernel<<<grid, SM_WARPS * 32, smem, s>>>(n, m, radius * radius, 1.0f / h, nsample, new_xyz, xyz, idx);
  PN2_CHECK_LAUNCH("bq_small_kernel");
  return PN2_OK;
}

bool ball_query_small_supported(int n) { return n >= 2048 && n <= SM_MAXN; }

}  // namespace pn2
