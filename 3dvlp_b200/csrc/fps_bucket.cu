// Bucketed furthest point sampling: ONE CTA per scene, exact, with spatial pruning.
//
// csrc/fps.cu spreads a scene over a 16-CTA cluster and touches all N points for every accepted sample.
// Late in a run almost none of those updates change anything: a new sample only lowers the running
// minimum `temp` of points closer to it than sqrt(temp).  This kernel exploits that while keeping the
// reference's output bit for bit (same fp32 distance expression, same (temp, key) arg-max as
// sampling_gpu.cu:75-178 -- see the header of fps.cu for the key):
//
//   prologue  the CTA counting-sorts its scene into a coarse Morton grid (shared-memory histogram,
//             no extra launches) and cuts the sorted sequence into BUCKETS of 64 consecutive points
//             (x, y, z, original index as float4 in a global scratch that stays L2-resident);
//             running minima live in shared memory (4 B/point), buckets carry {max temp, key and
//             xyz of that point, second-largest temp}; 32 consecutive buckets form a GROUP = one warp.
//   round     K centres were accepted in the previous round (batched acceptance as in fps.cu):
//     1. prune  lane c of every warp tests centre c against the warp's group box; only buckets of hit
//               groups test their own box.  A (bucket, centre) pair is dropped iff
//               0.9999 * d2(centre, box) >= max temp of the bucket: no point of the bucket can change.
//     2. update the surviving (bucket, centre-mask) items are spread over the 32 warps: lane = point,
//               16-byte coalesced load, fminf chain, redux.sync for the bucket's new max / key / second.
//     3. select thread = bucket, warp = group: redux arg-max -> one candidate per warp (32 candidates)
//               plus the bound that dominates every non-candidate.
//     4. accept warp 0 replays greedy FPS on the 32 candidates (exact; fps.cu header) and publishes
//               up to KMAX new centres.
// Work per sample drops from N distance updates to ~0.5 k, and a scene needs 1 SM instead of 16, so a
// batch no longer monopolises the GPU: other kernels and other batches run beside it.
#include <cstdlib>

#include "common.cuh"

namespace pn2 {
namespace {

constexpr int FB_THREADS = 1024;
constexpr int FB_WARPS = 32;
constexpr int FB_KMAX = 16;       // centres accepted per round at most
// Candidate groups per warp: the select step can cut a warp's buckets into FB_CPW lane segments, each reporting its
// own candidate and bound (segmented redux.sync), for 32 * FB_CPW candidates instead of 32: the bound -- the best
// value no group reported -- then comes from smaller groups and more samples are accepted per round.  Measured on
// 8 x 40 000 points (tools/fb_debug.py): 418 / 348 / 300 rounds for 1 / 2 / 4, but the segmented select
// (0.9 k -> 1.7 k -> 2.3 k cycles) and the single-warp greedy replay over more candidates (1.5 k -> 2.1 k -> 4.5 k)
// cost more than the saved rounds: 3.9 M / 4.2 M / 4.9 M cycles per scene.  Kept at 1.
#ifndef PN2_FB_CPW
#define PN2_FB_CPW 1
#endif
constexpr int FB_CPW = PN2_FB_CPW, FB_GL = 32 / FB_CPW, FB_NC = FB_WARPS * FB_CPW;
constexpr int FB_CELLS = 8192;    // sort grid: 32 x 32 (Morton) x 8
// Points per bucket = 32 * PPL (PPL points per lane).  Scenes of up to 50 176 points use PPL = 2 with the running
// minima in shared memory; larger scenes (up to 131 072 points: thread = bucket caps the buckets at 1024) use
// PPL = 4 with the minima in an L2-resident global scratch (GT), fetched together with the bucket's points.

struct FbArgs {
  const float* xyz;   // (b, n, 3)
  float4* sorted;     // (b, npad) scratch
  float* temps_g;     // (b, npad) running minima when they do not fit shared memory, else nullptr
  int* idxs;          // (b, m)
  float* new_xyz;     // (b, m, 3) or nullptr
  int n, m, log2bs, nb, npad;
};

struct Box { float lx, ly, lz, hx, hy, hz; };

__device__ __forceinline__ float box_d2(const Box& b, float x, float y, float z) {
  const float dx = fmaxf(fmaxf(b.lx - x, x - b.hx), 0.f);
  const float dy = fmaxf(fmaxf(b.ly - y, y - b.hy), 0.f);
  const float dz = fmaxf(fmaxf(b.lz - z, z - b.hz), 0.f);
  return dx * dx + dy * dy + dz * dz;
}

__device__ __forceinline__ uint32_t part1by1(uint32_t v) {  // spread the low 5 bits
  v &= 0x1fu;
  v = (v | (v << 4)) & 0x0f0fu;   // not needed beyond 5 bits but harmless
  v = (v | (v << 2)) & 0x3333u;
  v = (v | (v << 1)) & 0x5555u;
  return v;
}

#ifndef PN2_FB_MAXREG
#define PN2_FB_MAXREG 64
#endif
template <int PPL, bool GT>
__global__ void __maxnreg__(PN2_FB_MAXREG) fps_bucket_kernel(const FbArgs a) {
  constexpr int FB_BK = 32 * PPL;
  extern __shared__ __align__(16) uint8_t smem_raw[];
  __shared__ float4 ctr[FB_KMAX];     // accepted centres (x, y, z, key bits)
  __shared__ int ctr_count;
  __shared__ int nwork;
  __shared__ uint2 w_vk[FB_NC];
  __shared__ float4 w_xyz[FB_NC];
  __shared__ float red[FB_WARPS][6];
  __shared__ float s_org[3], s_scale[3];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int scene = blockIdx.x;
  const int n = a.n, m = a.m, log2bs = a.log2bs, nb = a.nb, npad = a.npad;
  const float* __restrict__ xyz = a.xyz + static_cast<size_t>(scene) * n * 3;
  float4* __restrict__ sorted = a.sorted + static_cast<size_t>(scene) * npad;
  int* __restrict__ idxs = a.idxs + static_cast<size_t>(scene) * m;
  float* __restrict__ oxyz = a.new_xyz ? a.new_xyz + static_cast<size_t>(scene) * m * 3 : nullptr;

  float* temps = GT ? a.temps_g + static_cast<size_t>(scene) * npad : reinterpret_cast<float*>(smem_raw);  // npad
  // bucket records: behind the minima in shared memory, or (GT) at its start -- the sort's histogram, which
  // aliases this region, is dead by the time they are first written
  float4* bb_a = GT ? reinterpret_cast<float4*>(smem_raw) : reinterpret_cast<float4*>(reinterpret_cast<float*>(smem_raw) + npad);  // nb: {max temp, key bits, x, y}
  float2* bb_b = reinterpret_cast<float2*>(bb_a + nb);               // nb: {z, second temp}
  uint32_t* work = reinterpret_cast<uint32_t*>(bb_b + nb);           // nb: bucket | centre mask << 16
  int* hist = reinterpret_cast<int*>(smem_raw);                      // aliases temps / the records during the sort
  int* cursor = hist + FB_CELLS;

  const uint32_t bsmask = (1u << log2bs) - 1u;
  const uint32_t rowmask = log2bs ? ((1u << (32 - log2bs)) - 1u) : 0xffffffffu;
  const uint32_t lt = (1u << lane) - 1u;
  const float x0 = __ldg(xyz), y0 = __ldg(xyz + 1), z0 = __ldg(xyz + 2);
#ifdef PN2_FB_DEBUG
  const long long dbg_t0 = clock64();
#endif

  // ================= prologue: counting sort into the coarse Morton grid ==========================
  {
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int k = tid; k < n; k += FB_THREADS) {
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        const float v = __ldg(xyz + 3 * k + d);
        if (fabsf(v) <= 3.0e38f) { lo[d] = fminf(lo[d], v); hi[d] = fmaxf(hi[d], v); }  // finite only
      }
    }
#pragma unroll
    for (int d = 0; d < 3; ++d) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        lo[d] = fminf(lo[d], __shfl_xor_sync(0xffffffffu, lo[d], o));
        hi[d] = fmaxf(hi[d], __shfl_xor_sync(0xffffffffu, hi[d], o));
      }
      if (lane == 0) { red[warp][d] = lo[d]; red[warp][3 + d] = hi[d]; }
    }
    for (int i = tid; i < 2 * FB_CELLS; i += FB_THREADS) hist[i] = 0;
    __syncthreads();
    if (tid < 3) {
      float l = INFINITY, h = -INFINITY;
      for (int w = 0; w < FB_WARPS; ++w) { l = fminf(l, red[w][tid]); h = fmaxf(h, red[w][3 + tid]); }
      if (!(h >= l)) { l = 0.f; h = 1.f; }
      const float cells = tid == 2 ? 8.f : 32.f;
      s_org[tid] = l;
      s_scale[tid] = cells / fmaxf((h - l) * 1.0001f, 1e-20f);
    }
    __syncthreads();
    const float ox = s_org[0], oy = s_org[1], oz = s_org[2], sx = s_scale[0], sy = s_scale[1], sz = s_scale[2];
    auto cell_of = [&](float x, float y, float z) -> int {
      const int ix = min(31, max(0, static_cast<int>((x - ox) * sx)));
      const int iy = min(31, max(0, static_cast<int>((y - oy) * sy)));
      const int iz = min(7, max(0, static_cast<int>((z - oz) * sz)));
      return static_cast<int>(((part1by1(ix) | (part1by1(iy) << 1)) << 3) | static_cast<uint32_t>(iz));
    };
    for (int k = tid; k < n; k += FB_THREADS)
      atomicAdd(&hist[cell_of(__ldg(xyz + 3 * k), __ldg(xyz + 3 * k + 1), __ldg(xyz + 3 * k + 2))], 1);
    __syncthreads();
    // exclusive scan of the 8192 counts: 8 per thread
    int v[8], sum = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) { v[i] = hist[tid * 8 + i]; sum += v[i]; }
    int incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    __shared__ int wtot[FB_WARPS];
    if (lane == 31) wtot[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      int w = wtot[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, w, o);
        if (lane >= o) w += t;
      }
      wtot[lane] = w;
    }
    __syncthreads();
    int run = incl - sum + (warp ? wtot[warp - 1] : 0);
#pragma unroll
    for (int i = 0; i < 8; ++i) { cursor[tid * 8 + i] = run; run += v[i]; }
    __syncthreads();
    for (int k = tid; k < n; k += FB_THREADS) {
      const float x = __ldg(xyz + 3 * k), y = __ldg(xyz + 3 * k + 1), z = __ldg(xyz + 3 * k + 2);
      const int pos = atomicAdd(&cursor[cell_of(x, y, z)], 1);
      // .w carries the point's tie-break KEY (fps.cu header), not its index: every bucket update needs the key,
      // only an accepted sample needs the index back (the accept step inverts the key)
      const uint32_t u = static_cast<uint32_t>(k);
      sorted[pos] = make_float4(x, y, z, __uint_as_float((log2bs ? __brev(u & bsmask) : 0u) | (u >> log2bs)));
    }
    for (int k = n + tid; k < npad; k += FB_THREADS) sorted[k] = make_float4(0.f, 0.f, 0.f, __uint_as_float(0xffffffffu));  // no point: no key
    __syncthreads();  // the CTA's own global writes are visible to all its threads from here on
  }

  // one bucket (32 * PPL points, PPL per lane): apply the centres in `mask`, refresh temps and the record.
  // `tin` = the running minima of the lane's points (ignored when `init`).
  auto process_bucket = [&](int b, uint32_t mask, bool init, const float4 (&p)[PPL], const float (&tin)[PPL]) {
    float t[PPL], told[PPL];
    uint32_t key[PPL];
#pragma unroll
    for (int q = 0; q < PPL; ++q) {
      key[q] = __float_as_uint(p[q].w);  // 0xffffffff = padding
      if (init) {
        // sampling_gpu.cu:105-106: `if (mag <= 1e-3) continue;` -- a double comparison
        const float mg = __fmaf_rn(p[q].z, p[q].z, __fmaf_rn(p[q].x, p[q].x, __fmul_rn(p[q].y, p[q].y)));
        t[q] = (key[q] != 0xffffffffu && !(static_cast<double>(mg) <= 1e-3)) ? 1e10f : -1.f;
      } else {
        t[q] = GT ? tin[q] : temps[b * FB_BK + 32 * q + lane];  // shared memory: read on the spot; global: prefetched
      }
      told[q] = t[q];
    }
    while (mask) {
      const int c = __ffs(mask) - 1;
      mask &= mask - 1;
      const float4 oc = ctr[c];
#pragma unroll
      for (int q = 0; q < PPL; ++q) t[q] = fminf(sqdist3(p[q].x - oc.x, p[q].y - oc.y, p[q].z - oc.z), t[q]);
    }
    if (!init) {
      // the box test is conservative: when no running minimum of the bucket moved, its minima and its record
      // (a function of them alone) are already right -- skip the stores and the record's three reductions
      bool moved = false;
#pragma unroll
      for (int q = 0; q < PPL; ++q) moved |= t[q] != told[q];
      if (!__any_sync(0xffffffffu, moved)) return;
    }
#pragma unroll
    for (int q = 0; q < PPL; ++q) temps[b * FB_BK + 32 * q + lane] = t[q];
    // lane-local best of its points: value, then smaller key
    int v[PPL];
#pragma unroll
    for (int q = 0; q < PPL; ++q) {
      v[q] = __float_as_int(t[q]);  // t >= 0 or exactly -1: int order ok
    }
    int vb = v[0];
    uint32_t kb = key[0];
    float bx = p[0].x, by = p[0].y, bz = p[0].z;
#pragma unroll
    for (int q = 1; q < PPL; ++q) {
      if (v[q] > vb || (v[q] == vb && key[q] < kb)) { vb = v[q]; kb = key[q]; bx = p[q].x; by = p[q].y; bz = p[q].z; }
    }
    const int vmax = __reduce_max_sync(0xffffffffu, vb);
    const uint32_t kk = (vb == vmax) ? kb : 0xffffffffu;
    const uint32_t kmin = __reduce_min_sync(0xffffffffu, kk);
    const bool win = (kk == kmin) && (vb == vmax);
    const int first = __ffs(__ballot_sync(0xffffffffu, win)) - 1;  // unique unless the bucket is empty
    // `second` = best value among the points the bucket does NOT report.  Exact coordinate duplicates
    // of the winner are left out: they share its temp history, have larger keys (the winner is the
    // smallest key among equals) and drop to exactly 0 the moment the winner is accepted, so they can
    // never be the reference's next pick ahead of a candidate.  Morton order puts duplicates (dataset
    // sampling with replacement, lib/joint/dataset.py:609-612) next to each other, so without this
    // every duplicated point would cap the accepted batch at one sample.
    const float wx = __shfl_sync(0xffffffffu, bx, first), wy = __shfl_sync(0xffffffffu, by, first),
                wz = __shfl_sync(0xffffffffu, bz, first);
    const int none = __float_as_int(-1.f);
    int rest = none;
#pragma unroll
    for (int q = 0; q < PPL; ++q) {
      const bool twin = (v[q] == vmax) && (p[q].x == wx) && (p[q].y == wy) && (p[q].z == wz);  // incl. the winner
      rest = max(rest, twin ? none : v[q]);
    }
    const int sec = __reduce_max_sync(0xffffffffu, rest);
    if (lane == first) {
      bb_a[b] = make_float4(__int_as_float(vmax), __uint_as_float(kmin), bx, by);
      bb_b[b] = make_float2(bz, __int_as_float(sec));
    }
  };
  auto load_bucket = [&](int b, float4 (&p)[PPL], float (&t)[PPL], bool with_temps) {
    const float4* src = sorted + b * FB_BK + lane;
#pragma unroll
    for (int q = 0; q < PPL; ++q) {
      p[q] = src[32 * q];
      t[q] = (GT && with_temps) ? temps[b * FB_BK + 32 * q + lane] : 0.f;
    }
  };

  // ================= init: temps, bucket records, bucket and group boxes ==========================
  for (int b = warp; b < nb; b += FB_WARPS) {
    float4 p[PPL];
    float t[PPL];
    load_bucket(b, p, t, false);
    process_bucket(b, 0u, true, p, t);
  }
  // thread = bucket; a warp holds FB_CPW candidate groups of GB consecutive buckets each (lane segments of FB_GL
  // lanes).  GB = ceil(nb / FB_NC) spreads the buckets over all warps; the prune step works on the warp's
  // whole range (one box per warp), the select step on the segments.
  const int GB = (nb + FB_NC - 1) / FB_NC;  // <= FB_GL: nb <= 1024
  const int sub = lane / FB_GL, sl = lane % FB_GL;
  const uint32_t smask = FB_CPW == 1 ? 0xffffffffu : (((1u << FB_GL) - 1u) << (sub * FB_GL));
  const int cgrp = warp * FB_CPW + sub;
  const int myb = (sl < GB && cgrp * GB + sl < nb) ? cgrp * GB + sl : -1;
  Box bbox = {INFINITY, INFINITY, INFINITY, -INFINITY, -INFINITY, -INFINITY};
  if (myb >= 0) {
    for (int i = 0; i < FB_BK; ++i) {
      const float4 p = sorted[myb * FB_BK + i];
      if (__float_as_uint(p.w) != 0xffffffffu) {
        bbox.lx = fminf(bbox.lx, p.x); bbox.ly = fminf(bbox.ly, p.y); bbox.lz = fminf(bbox.lz, p.z);
        bbox.hx = fmaxf(bbox.hx, p.x); bbox.hy = fmaxf(bbox.hy, p.y); bbox.hz = fmaxf(bbox.hz, p.z);
      }
    }
  }
  Box gbox = bbox;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    gbox.lx = fminf(gbox.lx, __shfl_xor_sync(0xffffffffu, gbox.lx, o)); gbox.ly = fminf(gbox.ly, __shfl_xor_sync(0xffffffffu, gbox.ly, o));
    gbox.lz = fminf(gbox.lz, __shfl_xor_sync(0xffffffffu, gbox.lz, o)); gbox.hx = fmaxf(gbox.hx, __shfl_xor_sync(0xffffffffu, gbox.hx, o));
    gbox.hy = fmaxf(gbox.hy, __shfl_xor_sync(0xffffffffu, gbox.hy, o)); gbox.hz = fmaxf(gbox.hz, __shfl_xor_sync(0xffffffffu, gbox.hz, o));
  }
  if (tid == 0) {
    ctr[0] = make_float4(x0, y0, z0, 0.f);  // sample 0 is index 0 (sampling_gpu.cu:89-90)
    ctr_count = 1;
    nwork = 0;
    idxs[0] = 0;
    if (oxyz) { oxyz[0] = x0; oxyz[1] = y0; oxyz[2] = z0; }
  }
  __syncthreads();
  float bmax = myb >= 0 ? bb_a[myb].x : -1.f;
  float gmax = __int_as_float(__reduce_max_sync(0xffffffffu, __float_as_int(bmax)));
  const bool has_group = warp * FB_CPW * GB < nb;  // warp-uniform

#ifdef PN2_FB_DEBUG
  long long dbg_items = 0, dbg_t[5] = {0, 0, 0, 0, 0}, dbg_prev = clock64();
  printf(tid == 0 ? "fps_bucket[%d]: prologue+init %lld cycles\n" : "", scene, dbg_prev - dbg_t0);
  int dbg_rounds = 0;
#define FB_T(i) { const long long _t = clock64(); dbg_t[i] += _t - dbg_prev; dbg_prev = _t; }
#else
#define FB_T(i)
#endif
  int j = 1;
  while (j < m) {
    // ---------------- 1. prune: which (bucket, centre) pairs can change anything ------------------
    const int K = ctr_count;
    if (has_group && gmax >= 0.f) {
      float4 myc = make_float4(0.f, 0.f, 0.f, 0.f);
      if (lane < K) myc = ctr[lane];
      const bool gaff = lane < K && !(box_d2(gbox, myc.x, myc.y, myc.z) * 0.9999f >= gmax);
      uint32_t gmask = __ballot_sync(0xffffffffu, gaff);
      if (gmask != 0u) {
        uint32_t mask = 0u;
        const bool live = myb >= 0 && bmax >= 0.f;
        while (gmask) {
          const int c = __ffs(gmask) - 1;
          gmask &= gmask - 1;
          const float4 o = ctr[c];
          if (live && !(box_d2(bbox, o.x, o.y, o.z) * 0.9999f >= bmax)) mask |= 1u << c;
        }
        const uint32_t any = __ballot_sync(0xffffffffu, mask != 0u);
        if (any) {
          int base = 0;
          if (lane == 0) base = atomicAdd(&nwork, __popc(any));
          base = __shfl_sync(0xffffffffu, base, 0);
          if (mask) work[base + __popc(any & lt)] = static_cast<uint32_t>(myb) | (mask << 16);
        }
      }
    }
    __syncthreads();
    FB_T(0)
    // ---------------- 2. update the surviving buckets (next item's points are fetched while the
    // current one is processed: the L2 latency of the sorted array is off the critical path) --------
    const int nw = nwork;
    {
      uint32_t item = 0u;
      float4 pn[PPL];
      float tn[PPL];
#pragma unroll
      for (int q = 0; q < PPL; ++q) { pn[q] = make_float4(0.f, 0.f, 0.f, 0.f); tn[q] = 0.f; }
      if (warp < nw) {
        item = work[warp];
        load_bucket(static_cast<int>(item & 0xffffu), pn, tn, true);
      }
      for (int i = warp; i < nw; i += FB_WARPS) {
        const uint32_t cur = item;
        float4 pc[PPL];
        float tcur[PPL];
#pragma unroll
        for (int q = 0; q < PPL; ++q) { pc[q] = pn[q]; tcur[q] = tn[q]; }
        if (i + FB_WARPS < nw) {  // a bucket appears once per round: its minima cannot change under the prefetch
          item = work[i + FB_WARPS];
          load_bucket(static_cast<int>(item & 0xffffu), pn, tn, true);
        }
        process_bucket(static_cast<int>(cur & 0xffffu), cur >> 16, false, pc, tcur);
      }
    }
    __syncthreads();
    FB_T(1)
#ifdef PN2_FB_DEBUG
    dbg_items += nw; ++dbg_rounds;
#endif
    // ---------------- 3. select: thread = bucket, warp = group ---------------------------------------
    {
      int best = __float_as_int(-1.f), second = __float_as_int(-1.f);
      uint32_t bkey = 0xffffffffu;
      float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
      float rz = 0.f;
      if (myb >= 0) {
        r = bb_a[myb];
        const float2 q = bb_b[myb];
        best = __float_as_int(r.x); bkey = __float_as_uint(r.y);
        second = __float_as_int(q.y); rz = q.x;
      }
      bmax = __int_as_float(best);
      gmax = __int_as_float(__reduce_max_sync(0xffffffffu, best));  // of the warp's whole range (prune)
      const int vmax = __reduce_max_sync(smask, best);               // of this lane's candidate group
      const uint32_t kk = (best == vmax) ? bkey : 0xffffffffu;
      const uint32_t kmin = __reduce_min_sync(smask, kk);
      const bool win = (kk == kmin) && (best == vmax);
      const int first = __ffs(__ballot_sync(0xffffffffu, win) & smask) - 1;
      const float wx = __shfl_sync(0xffffffffu, r.z, first), wy = __shfl_sync(0xffffffffu, r.w, first),
                  wz = __shfl_sync(0xffffffffu, rz, first);
      // a neighbouring bucket whose best is an exact coordinate twin of the group's winner (duplicates
      // split by a bucket boundary) contributes its `second`, not its best -- same argument as above
      const bool twin = (best == vmax) && (r.z == wx) && (r.w == wy) && (rz == wz);
      const int wsec = __reduce_max_sync(smask, (lane == first || twin) ? second : best);
      if (lane == first) {
        w_vk[cgrp] = make_uint2(static_cast<uint32_t>(vmax), kmin);
        w_xyz[cgrp] = make_float4(r.z, r.w, rz, __int_as_float(wsec));
      }
      if (tid == 0) nwork = 0;
    }
    __syncthreads();
    FB_T(2)
    // ---------------- 4. accept: greedy FPS on the 32 warp candidates (exact) ------------------------
    if (warp == 0) {
      // lane holds candidates lane + 32 q
      int cv[FB_CPW];
      uint32_t ck[FB_CPW];
      float cx[FB_CPW], cy[FB_CPW], cz[FB_CPW];
      int lb = __float_as_int(-1.f);
#pragma unroll
      for (int q = 0; q < FB_CPW; ++q) {
        const uint2 r = w_vk[lane + 32 * q];
        const float4 w = w_xyz[lane + 32 * q];
        cv[q] = static_cast<int>(r.x); ck[q] = r.y;
        cx[q] = w.x; cy[q] = w.y; cz[q] = w.z;
        lb = max(lb, __float_as_int(w.w));
      }
      const int bnd = max(__reduce_max_sync(0xffffffffu, lb), 0);
      const int room = min(FB_KMAX, m - j);
      int k = 0;
      // the lane's best candidate: value, then smaller key
      int lv;
      uint32_t lk;
      float lx, ly, lz;
      auto lane_best = [&]() {
        lv = cv[0]; lk = ck[0]; lx = cx[0]; ly = cy[0]; lz = cz[0];
#pragma unroll
        for (int q = 1; q < FB_CPW; ++q)
          if (cv[q] > lv || (cv[q] == lv && ck[q] < lk)) { lv = cv[q]; lk = ck[q]; lx = cx[q]; ly = cy[q]; lz = cz[q]; }
      };
      lane_best();
      int rv = __reduce_max_sync(0xffffffffu, lv);
      if (rv < 0) {
        // no eligible point anywhere: every reference thread reports besti = 0 (key of index 0 is 0)
        if (lane == 0) ctr[0] = make_float4(x0, y0, z0, __uint_as_float(0u));
        k = 1;
      } else {
        for (;;) {
          uint32_t tie = __ballot_sync(0xffffffffu, lv == rv);
          if (tie & (tie - 1)) {  // equal values: the smallest key wins (keys are unique)
            const uint32_t kk = (lv == rv) ? lk : 0xffffffffu;
            const uint32_t kmin = __reduce_min_sync(0xffffffffu, kk);
            tie = __ballot_sync(0xffffffffu, kk == kmin);
          }
          const int src = __ffs(tie) - 1;
          const float ax = __shfl_sync(0xffffffffu, lx, src), ay = __shfl_sync(0xffffffffu, ly, src),
                      az = __shfl_sync(0xffffffffu, lz, src);
          const uint32_t ak = __shfl_sync(0xffffffffu, lk, src);
          if (lane == 0) ctr[k] = make_float4(ax, ay, az, __uint_as_float(ak));
          if (++k >= room) break;
#pragma unroll
          for (int q = 0; q < FB_CPW; ++q)
            cv[q] = __float_as_int(fminf(sqdist3(cx[q] - ax, cy[q] - ay, cz[q] - az), __int_as_float(cv[q])));
          lane_best();
          rv = __reduce_max_sync(0xffffffffu, lv);
          if (!(rv > bnd)) break;
        }
      }
      __syncwarp();
      if (lane < k) {
        const float4 cc = ctr[lane];
        const uint32_t mykey = __float_as_uint(cc.w);
        const uint32_t row = mykey & rowmask;
        const uint32_t vv = log2bs ? __brev(mykey & ~rowmask) : 0u;
        idxs[j + lane] = static_cast<int>(vv + (row << log2bs));
        if (oxyz) { oxyz[3 * (j + lane)] = cc.x; oxyz[3 * (j + lane) + 1] = cc.y; oxyz[3 * (j + lane) + 2] = cc.z; }
      }
      if (lane == 0) ctr_count = k;
    }
    __syncthreads();
    FB_T(3)
    j += ctr_count;
  }
#ifdef PN2_FB_DEBUG
  if (tid == 0)
    printf("fps_bucket[%d]: total %lld cycles; rounds", scene, clock64() - dbg_t0);
  if (tid == 0)
    printf(" rounds %d items %lld (%.1f/round) cycles/round: prune %lld update %lld select %lld accept %lld\n", dbg_rounds,
           dbg_items, double(dbg_items) / dbg_rounds, dbg_t[0] / dbg_rounds, dbg_t[1] / dbg_rounds, dbg_t[2] / dbg_rounds,
           dbg_t[3] / dbg_rounds);
#endif
}

}  // namespace

// Scenes of up to 50 176 points keep the running minima on chip (4 B/point + 28 B per 64-point bucket within the
// 227 KB of shared memory of an sm_100a CTA); up to 131 072 points (1024 buckets of 128) the minima live in an
// L2-resident scratch and only the bucket records stay in shared memory.
constexpr int kSmemPoints = 50176, kMaxPoints = 131072;
bool fps_bucket_supported(int n) { return n >= 16384 && n <= kMaxPoints; }

int fps_bucket_run(int b, int n, int m, int log2bs, const float* dataset, int* idxs, float* new_xyz,
                   cudaStream_t stream) {
  FbArgs a;
  a.xyz = dataset; a.idxs = idxs; a.new_xyz = new_xyz;
  a.n = n; a.m = m; a.log2bs = log2bs;
  const bool big = n > kSmemPoints;
  const int bk = big ? 128 : 64;
  a.nb = (n + bk - 1) / bk;
  a.npad = a.nb * bk;
  PN2_REQUIRE(a.nb <= FB_THREADS, "fps_bucket: too many buckets");
  const size_t records = static_cast<size_t>(a.nb) * (16 + 8 + 4) + 64;
  const size_t sort_bytes = 2 * FB_CELLS * sizeof(int);
  const size_t smem = big ? (records > sort_bytes ? records : sort_bytes) : static_cast<size_t>(a.npad) * 4 + records;
  PN2_REQUIRE(smem <= 227 * 1024 && (big || static_cast<size_t>(a.npad) * 4 >= sort_bytes),
              "fps_bucket: n=%d outside the supported range", n);
  uint8_t* ws = nullptr;
  const size_t sorted_bytes = static_cast<size_t>(b) * a.npad * sizeof(float4);
  const size_t temps_bytes = big ? static_cast<size_t>(b) * a.npad * sizeof(float) : 0;
  PN2_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&ws), sorted_bytes + temps_bytes, stream));
  a.sorted = reinterpret_cast<float4*>(ws);
  a.temps_g = big ? reinterpret_cast<float*>(ws + sorted_bytes) : nullptr;
  cudaError_t e;
  if (big) {
    e = cudaFuncSetAttribute(fps_bucket_kernel<4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e == cudaSuccess) {
      fps_bucket_kernel<4, true><<<b, FB_THREADS, smem, stream>>>(a);
      e = cudaGetLastError();
    }
  } else {
    e = cudaFuncSetAttribute(fps_bucket_kernel<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e == cudaSuccess) {
      fps_bucket_kernel<2, false><<<b, FB_THREADS, smem, stream>>>(a);
      e = cudaGetLastError();
    }
  }
  cudaFreeAsync(ws, stream);
  if (e != cudaSuccess) return cuda_fail(e, "fps_bucket_kernel");
  return PN2_OK;
}

}  // namespace pn2
