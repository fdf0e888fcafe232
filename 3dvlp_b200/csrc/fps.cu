// Furthest point sampling for sm_100a: one thread-block CLUSTER per scene, every point resident in
// registers for the whole run, one associative packed-key arg-max per round.
//
// Replaces furthest_point_sampling_kernel (reference lib/pointnet2/_ext_src/src/sampling_gpu.cu:75-178),
// which runs ONE 512-thread CTA per scene and re-reads xyz and the running minima from global
// memory in each of the m-1 dependent rounds.
//
// Bit-exactness.  The reference's selection rule is a pure function of the point index
// (verified against a literal simulation in oracle/pointnet2_oracle.c):
// with BS = opt_n_threads(n) reference threads, point k belongs to thread k mod BS, a thread keeps
// its smallest maximal k, and the shared-memory halving tree prefers the thread with the smallest
// bit-reversed id.  So   next = argmax_k ( temp[k], -key32(k) ),
//     key32(k) = bitrev_{log2 BS}(k mod BS) << (32 - log2 BS)  |  (k div BS)
// and any decomposition of that max (lane -> warp -> CTA -> cluster) returns the same index.
//
// Ownership.  Thread g of the cluster (g = cta_rank * THREADS + tid) plays reference thread
// v = g mod BS for the contiguous row segment [s*P, (s+1)*P), s = g div BS, of that thread's points
// k = v + BS*row.  Within a thread rows ascend, so the reference's strict `>` scan order is kept;
// between threads the packed key decides.  For a fixed row, consecutive tid read consecutive k
// (coalesced one-time load).
//
// Per round: P distance updates per thread (registers only) -> redux.sync max/min inside the warp ->
// one __syncthreads -> warp 0 reduces the warp records and posts the CTA's candidate {value, key,
// x, y, z} into every peer's shared memory with st.async (DSMEM store that completes a transaction
// on the peer's mbarrier) -> every CTA waits on its own mbarrier and reduces the CLUSTER
// candidates redundantly, so all CTAs know the new centre's coordinates without touching global
// memory.  There is no cluster-wide barrier inside the loop.
#include "common.cuh"

namespace pn2 {
namespace {

struct FpsArgs {
  const float* xyz;  // (b, n, 3)
  int* idxs;         // (b, m)
  float* new_xyz;    // (b, m, 3) or nullptr
  int* err;          // device error flag (protocol time-out)
  int n, m, log2bs, rows;
};

__device__ __forceinline__ void st_async_v4(uint32_t raddr, uint32_t a, uint32_t b, uint32_t c,
                                            uint32_t d, uint32_t rbar) {
  asm volatile(
      "st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];"
      ::"r"(raddr), "r"(a), "r"(b), "r"(c), "r"(d), "r"(rbar)
      : "memory");
}
__device__ __forceinline__ void st_async_b32(uint32_t raddr, uint32_t a, uint32_t rbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(raddr),
               "r"(a), "r"(rbar)
               : "memory");
}

// Reduce NR (<= 32) records held one per lane (lanes >= NR pass INT_MIN / ~0u): returns the lane
// that holds the arg-max of (value, -key); value/key of the winner are returned by reference.
__device__ __forceinline__ int warp_argmax(int& v, uint32_t& k) {
  const int vmax = __reduce_max_sync(0xffffffffu, v);
  const uint32_t kk = (v == vmax) ? k : 0xffffffffu;
  const uint32_t kmin = __reduce_min_sync(0xffffffffu, kk);
  const uint32_t win = __ballot_sync(0xffffffffu, kk == kmin && v == vmax);
  v = vmax;
  k = kmin;
  return __ffs(win) - 1;
}

template <int P, int THREADS, int CLUSTER>
__global__ void __launch_bounds__(THREADS, 1) fps_kernel(const FpsArgs a) {
  constexpr int NW = THREADS / 32;
  constexpr int NC = CLUSTER > 1 ? CLUSTER : 1;
  __shared__ uint2 w_vk[2][NW];
  __shared__ float4 w_xyz[2][NW];
  __shared__ __align__(16) uint4 cand[2][NC];  // {value bits, key32, x, y}
  __shared__ float cand_z[2][NC];
  __shared__ __align__(8) uint64_t mbar[2];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t rank = CLUSTER > 1 ? cluster_ctarank() : 0u;
  const int scene = blockIdx.x / CLUSTER;
  const int n = a.n, m = a.m, log2bs = a.log2bs;
  const float* __restrict__ xyz = a.xyz + static_cast<size_t>(scene) * n * 3;
  int* __restrict__ idxs = a.idxs + static_cast<size_t>(scene) * m;
  float* __restrict__ oxyz = a.new_xyz ? a.new_xyz + static_cast<size_t>(scene) * m * 3 : nullptr;

  const uint32_t bsmask = (1u << log2bs) - 1u;
  const uint32_t g = rank * THREADS + tid;
  const uint32_t v = g & bsmask;
  const uint32_t seg = g >> log2bs;
  const uint32_t rowmask = log2bs ? ((1u << (32 - log2bs)) - 1u) : 0xffffffffu;
  const uint32_t keybase = (log2bs ? __brev(v) : 0u) | (seg * P);  // key of slot i = keybase + i

  // ---- one-time load: coordinates and running minima live in registers from here on ----------
  float px[P], py[P], pz[P], pt[P];
#pragma unroll
  for (int i = 0; i < P; ++i) {
    const uint32_t row = seg * P + i;
    const long long k = static_cast<long long>(v) + (static_cast<long long>(row) << log2bs);
    px[i] = py[i] = pz[i] = 0.f;
    pt[i] = -1.f;  // never selectable, never raised by fminf
    if (row < static_cast<uint32_t>(a.rows) && k < n) {
      const float x = __ldg(xyz + 3 * k), y = __ldg(xyz + 3 * k + 1), z = __ldg(xyz + 3 * k + 2);
      const float mag = __fmaf_rn(z, z, __fmaf_rn(x, x, __fmul_rn(y, y)));
      px[i] = x; py[i] = y; pz[i] = z;
      // sampling_gpu.cu:105-106: `if (mag <= 1e-3) continue;` -- a double comparison
      if (!(static_cast<double>(mag) <= 1e-3)) pt[i] = 1e10f;
    }
  }
  const float x0 = __ldg(xyz), y0 = __ldg(xyz + 1), z0 = __ldg(xyz + 2);
  float ox = x0, oy = y0, oz = z0;

  if (CLUSTER > 1) {
    if (tid == 0) {
      mbar_init(&mbar[0], 1);
      mbar_init(&mbar[1], 1);
      fence_mbar_init_cluster();
    }
    cluster_sync_all();
  }
  if (rank == 0 && tid == 0) {
    idxs[0] = 0;
    if (oxyz) { oxyz[0] = x0; oxyz[1] = y0; oxyz[2] = z0; }
  }

  for (int j = 1; j < m; ++j) {
    const int par = j & 1;
    if (CLUSTER > 1 && tid == 0) mbar_arrive_expect_tx(&mbar[par], CLUSTER * 20u);

    // ---- A: update the running minima of the owned points, thread-local arg-max -------------
    float best = -1.f;
    int bslot = 0;
#pragma unroll
    for (int i = 0; i < P; ++i) {
      const float d = sqdist3(px[i] - ox, py[i] - oy, pz[i] - oz);
      const float t = fminf(d, pt[i]);
      pt[i] = t;
      if (t > best) { best = t; bslot = i; }
    }
    // ---- B: warp arg-max with two redux.sync ---------------------------------------------------
    int vb = __float_as_int(best);  // best >= 0 or exactly -1.0f: signed-int order == float order
    uint32_t key = keybase + bslot;
    {
      const int vmax = __reduce_max_sync(0xffffffffu, vb);
      const uint32_t kk = (vb == vmax) ? key : 0xffffffffu;
      const uint32_t kmin = __reduce_min_sync(0xffffffffu, kk);
      if (kk == kmin) {  // exactly one lane: keys are unique
        float bx = px[0], by = py[0], bz = pz[0];
#pragma unroll
        for (int i = 1; i < P; ++i)
          if (bslot == i) { bx = px[i]; by = py[i]; bz = pz[i]; }
        w_vk[par][warp] = make_uint2(static_cast<uint32_t>(vmax), kmin);
        w_xyz[par][warp] = make_float4(bx, by, bz, 0.f);
      }
    }
    __syncthreads();

    int gv;
    uint32_t gk;
    float gx, gy, gz;
    if (CLUSTER == 1) {
      // ---- C: every warp reduces the NW warp records redundantly (no second barrier) ---------
      int rv = INT_MIN;
      uint32_t rk = 0xffffffffu;
      if (lane < NW) { const uint2 r = w_vk[par][lane]; rv = static_cast<int>(r.x); rk = r.y; }
      const int src = warp_argmax(rv, rk);
      const float4 w = w_xyz[par][src];
      gv = rv; gk = rk; gx = w.x; gy = w.y; gz = w.z;
    } else {
      // ---- C+D: warp 0 reduces the CTA and posts its candidate to all peers -------------------
      if (warp == 0) {
        int rv = INT_MIN;
        uint32_t rk = 0xffffffffu;
        if (lane < NW) { const uint2 r = w_vk[par][lane]; rv = static_cast<int>(r.x); rk = r.y; }
        const int src = warp_argmax(rv, rk);
        const float4 w = w_xyz[par][src];
        if (lane < CLUSTER) {
          const uint32_t rbar = mapa_u32(smem_u32(&mbar[par]), lane);
          st_async_v4(mapa_u32(smem_u32(&cand[par][rank]), lane), static_cast<uint32_t>(rv), rk,
                      __float_as_uint(w.x), __float_as_uint(w.y), rbar);
          st_async_b32(mapa_u32(smem_u32(&cand_z[par][rank]), lane), __float_as_uint(w.z), rbar);
        }
      }
      // ---- E: wait for the CLUSTER candidates, reduce them redundantly in every warp ----------
      if (!mbar_wait(&mbar[par], ((j - 1) >> 1) & 1)) {
        if (tid == 0) atomicExch(a.err, 1);
        break;
      }
      int rv = INT_MIN;
      uint32_t rk = 0xffffffffu;
      if (lane < CLUSTER) { const uint4 c = cand[par][lane]; rv = static_cast<int>(c.x); rk = c.y; }
      const int src = warp_argmax(rv, rk);
      const uint4 c = cand[par][src];
      gv = rv; gk = rk; gx = __uint_as_float(c.z); gy = __uint_as_float(c.w); gz = cand_z[par][src];
    }

    int old;
    if (gv < 0) {  // no eligible point at all: every reference thread reports besti = 0
      old = 0; ox = x0; oy = y0; oz = z0;
    } else {
      const uint32_t row = gk & rowmask;
      const uint32_t vv = log2bs ? __brev(gk & ~rowmask) : 0u;
      old = static_cast<int>(vv + (row << log2bs));
      ox = gx; oy = gy; oz = gz;
    }
    if (rank == 0 && tid == 0) {
      idxs[j] = old;
      if (oxyz) { oxyz[3 * j] = ox; oxyz[3 * j + 1] = oy; oxyz[3 * j + 2] = oz; }
    }
  }
  // peers may still be storing into this CTA's shared memory: leave together
  if (CLUSTER > 1) cluster_sync_all();
}

template <int P, int THREADS, int CLUSTER>
int launch(const FpsArgs& a, int b, cudaStream_t stream) {
  auto kern = fps_kernel<P, THREADS, CLUSTER>;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(static_cast<unsigned>(b) * CLUSTER);
  cfg.blockDim = dim3(THREADS);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  int nattr = 0;
  if (CLUSTER > 1) {
    if (CLUSTER > 8) PN2_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CLUSTER;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    nattr = 1;
  }
  cfg.attrs = attr;
  cfg.numAttrs = nattr;
  PN2_CUDA(cudaLaunchKernelEx(&cfg, kern, a));
  return PN2_OK;
}

template <int THREADS, int CLUSTER>
int dispatch_p(int p, const FpsArgs& a, int b, cudaStream_t s) {
  switch (p) {
    case 1: return launch<1, THREADS, CLUSTER>(a, b, s);
    case 2: return launch<2, THREADS, CLUSTER>(a, b, s);
    case 3: return launch<3, THREADS, CLUSTER>(a, b, s);
    case 4: return launch<4, THREADS, CLUSTER>(a, b, s);
    case 5: return launch<5, THREADS, CLUSTER>(a, b, s);
    case 6: return launch<6, THREADS, CLUSTER>(a, b, s);
    case 7: case 8: return launch<8, THREADS, CLUSTER>(a, b, s);
    case 9: case 10: case 11: case 12: return launch<12, THREADS, CLUSTER>(a, b, s);
    case 13: case 14: case 15: case 16: return launch<16, THREADS, CLUSTER>(a, b, s);
    default: break;
  }
  set_error("furthest_point_sampling: n too large for the register-resident kernel (P=%d)", p);
  return PN2_ERR_INVALID;
}

}  // namespace

int* device_err_flag();  // abi.cu

// Tunables (read once): PN2_FPS_CLUSTER = 1 | 8 | 16, PN2_FPS_THREADS = 512 | 1024.
static int env_int(const char* name, int dflt) {
  const char* s = getenv(name);
  return s ? atoi(s) : dflt;
}

int fps_run(int b, int n, int m, const float* dataset, int* idxs, float* new_xyz, cudaStream_t stream) {
  if (m <= 0 || b <= 0) return PN2_OK;  // sampling_gpu.cu:78
  PN2_REQUIRE(n > 0, "furthest_point_sampling: n must be positive (got %d)", n);
  PN2_REQUIRE(dataset && idxs, "furthest_point_sampling: null pointer");
  const int bs = opt_n_threads(n);
  int log2bs = 0;
  while ((1 << log2bs) < bs) ++log2bs;
  FpsArgs a;
  a.xyz = dataset; a.idxs = idxs; a.new_xyz = new_xyz;
  a.n = n; a.m = m; a.log2bs = log2bs; a.rows = (n + bs - 1) / bs;
  a.err = device_err_flag();
  PN2_REQUIRE(a.err != nullptr, "furthest_point_sampling: cannot allocate the device error flag");

  static const int force_cluster = env_int("PN2_FPS_CLUSTER", 0);
  static const int force_threads = env_int("PN2_FPS_THREADS", 0);
  int cluster = force_cluster ? force_cluster : (n > 8192 ? 16 : 1);
  int threads = force_threads ? force_threads : (cluster == 1 ? (n <= 512 ? 512 : 1024) : 512);
  if (cluster == 16 && threads == 512 && a.rows > 16 * 16) threads = 1024;
  const int segs = cluster * threads / bs;  // >= 1: threads >= 512 >= bs
  const int p = (a.rows + segs - 1) / segs;
  if (cluster == 1 && threads == 512) return dispatch_p<512, 1>(p, a, b, stream);
  if (cluster == 1) return dispatch_p<1024, 1>(p, a, b, stream);
  if (cluster == 8 && threads == 512) return dispatch_p<512, 8>(p, a, b, stream);
  if (cluster == 8) return dispatch_p<1024, 8>(p, a, b, stream);
  if (cluster == 16 && threads == 512) return dispatch_p<512, 16>(p, a, b, stream);
  if (cluster == 16) return dispatch_p<1024, 16>(p, a, b, stream);
  set_error("furthest_point_sampling: unsupported PN2_FPS_CLUSTER=%d", cluster);
  return PN2_ERR_INVALID;
}

}  // namespace pn2

extern "C" int pn2_furthest_point_sampling(int b, int n, int m, const float* dataset, float* temp,
                                           int* idxs, void* stream) {
  (void)temp;  // running minima stay in registers; the reference's scratch is not needed
  return pn2::fps_run(b, n, m, dataset, idxs, nullptr, pn2::as_stream(stream));
}

extern "C" int pn2_furthest_point_sampling_gather(int b, int n, int m, const float* dataset, int* idxs,
                                                  float* new_xyz, void* stream) {
  return pn2::fps_run(b, n, m, dataset, idxs, new_xyz, pn2::as_stream(stream));
}
