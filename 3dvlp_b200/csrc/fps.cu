// Furthest point sampling for sm_100a: one thread-block CLUSTER per scene, every point resident in
// registers for the whole run, one associative packed-key arg-max per round, and SEVERAL samples
// accepted per round.
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
// Batched acceptance (what makes the serial chain short).  The reduction keeps, per group (lane,
// warp, CTA), the best point AND the value of the best point it did NOT report ("second").  After
// one exchange every CTA holds G candidates -- one per CTA of the cluster (or one per warp in the
// single-CTA kernel) -- and  bound = max of the groups' seconds  dominates every point that is not a
// candidate.  Running minima only decrease, so warp 0 can replay greedy FPS on the G candidates
// alone: accept the arg-max, lower the other candidates by their exact distance to it (same fp32
// expression as the bulk update), and accept the next arg-max as long as its value stays strictly
// above `bound` and above zero -- then it is provably the point the reference would pick next.  The
// accepted centres (up to KMAX per round) are applied to all points in the next bulk update.  On
// ScanNet-like scenes ~3-5 samples are accepted per exchange, so the number of dependent
// exchange rounds drops from m-1 to roughly (m-1)/4 with bit-identical output.
//
// Per round: K x P distance updates per thread (registers only) -> redux.sync inside the warp ->
// __syncthreads -> warp 0 reduces the warp records and posts the CTA's candidate into every peer's
// shared memory with st.async (DSMEM store completing a transaction on the peer's mbarrier) ->
// warp 0 waits on its own mbarrier, runs the acceptance loop, publishes the new centres ->
// __syncthreads.  There is no cluster-wide barrier inside the loop.
#include "common.cuh"

namespace pn2 {
#ifndef PN2_FPS_TIMING
namespace {
#endif

constexpr int KMAX = 16;     // samples accepted per exchange round at most
constexpr int MAXC = 16;     // largest cluster

struct FpsArgs {
  const float* xyz;  // (b, n, 3)
  int* idxs;         // (b, m)
  float* new_xyz;    // (b, m, 3) or nullptr
  int* err;          // device error flag (protocol time-out)
  const int* in_order;  // per scene: 1 = the input is verified to be in FPS order (fps_prefix.cu), or nullptr
  int n, m, log2bs, rows, csize;
  long long* dbg;    // PN2_FPS_TIMING builds only
};

__device__ __forceinline__ void st_async_v4(uint32_t raddr, uint32_t a, uint32_t b, uint32_t c,
                                            uint32_t d, uint32_t rbar) {
  asm volatile(
      "st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];"
      ::"r"(raddr), "r"(a), "r"(b), "r"(c), "r"(d), "r"(rbar)
      : "memory");
}
__device__ __forceinline__ void st_async_v2(uint32_t raddr, uint32_t a, uint32_t b, uint32_t rbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.b32 [%0], {%1, %2}, [%3];" ::"r"(
                   raddr),
               "r"(a), "r"(b), "r"(rbar)
               : "memory");
}

// arg-max of (value, -key) over the lanes of a warp; lanes without a record pass v = INT_MIN.
// Returns the source lane; v / k become the winner's.  Value ties fall back to the keys, which
// are unique.
__device__ __forceinline__ int warp_argmax(int& v, uint32_t& k) {
  const int vmax = __reduce_max_sync(0xffffffffu, v);
  const uint32_t kk = (v == vmax) ? k : 0xffffffffu;
  const uint32_t kmin = __reduce_min_sync(0xffffffffu, kk);
  const uint32_t win = __ballot_sync(0xffffffffu, kk == kmin && v == vmax);
  v = vmax;
  k = kmin;
  return __ffs(win) - 1;
}

struct Centres {          // published by warp 0 each round
  float4 c[KMAX];         // accepted centres (x, y, z, -)
  int count;
};

template <int P, int THREADS, bool CLUSTERED>
// P <= 5 fits 64 registers per thread: two CTAs (of different clusters / scenes) share an SM, so the
// idle issue slots of one scene's serial phase run another scene's bulk update.
__global__ void __launch_bounds__(THREADS, (P <= 5 && THREADS <= 512) ? 2 : 1) fps_kernel(const FpsArgs a) {
  constexpr int NW = THREADS / 32;
  __shared__ uint2 w_vk[NW];                      // warp record: value bits, key
  __shared__ float4 w_xyz[NW];                    // x, y, z, second
  __shared__ __align__(16) uint4 cand[2][32];     // {value bits, key32, x, y}, slot = rank * T + t
  __shared__ __align__(8) uint2 cand_zs[2][32];   // {z, CTA bound}
  __shared__ __align__(8) uint64_t mbar[2];
  __shared__ Centres ctr;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int csize = CLUSTERED ? a.csize : 1;
  const uint32_t rank = CLUSTERED ? cluster_ctarank() : 0u;
  const int scene = blockIdx.x / csize;
  const int n = a.n, m = a.m, log2bs = a.log2bs;
  const float* __restrict__ xyz = a.xyz + static_cast<size_t>(scene) * n * 3;
  int* __restrict__ idxs = a.idxs + static_cast<size_t>(scene) * m;
  float* __restrict__ oxyz = a.new_xyz ? a.new_xyz + static_cast<size_t>(scene) * m * 3 : nullptr;

  if (!CLUSTERED && a.in_order != nullptr && a.in_order[scene] != 0) {
    // verified on the device: the reference's samples are 0, 1, ..., m-1
    for (int k = tid; k < m; k += THREADS) {
      idxs[k] = k;
      if (oxyz) { oxyz[3 * k] = xyz[3 * k]; oxyz[3 * k + 1] = xyz[3 * k + 1]; oxyz[3 * k + 2] = xyz[3 * k + 2]; }
    }
    return;
  }
  const uint32_t bsmask = (1u << log2bs) - 1u;
  // Thread -> reference-thread map: consecutive reference threads (consecutive point indices) are
  // dealt to DIFFERENT warps, so that when the input is already in sampling order (SA2..SA4 run FPS
  // on the previous layer's FPS output, whose next samples are k, k+1, k+2, ...) the per-warp
  // candidates of one round are distinct upcoming samples and the batch acceptance stays long.
  const uint32_t g = rank * THREADS + static_cast<uint32_t>(lane) * NW + warp;
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

  if (tid == 0) {
    if (CLUSTERED) {
      mbar_init(&mbar[0], 1);
      mbar_init(&mbar[1], 1);
      fence_mbar_init_cluster();
    }
    ctr.c[0] = make_float4(x0, y0, z0, 0.f);  // sample 0 is index 0 (sampling_gpu.cu:89-90)
    ctr.count = 1;
    if (rank == 0) {
      idxs[0] = 0;
      if (oxyz) { oxyz[0] = x0; oxyz[1] = y0; oxyz[2] = z0; }
    }
  }
  if (CLUSTERED) cluster_sync_all(); else __syncthreads();

#ifdef PN2_FPS_TIMING
  long long tacc[6] = {0, 0, 0, 0, 0, 0};
  long long nrounds = 0;
#define PN2_T(i) { const long long _t = clock64(); tacc[i] += _t - tprev; tprev = _t; }
  long long tprev = clock64();
#else
#define PN2_T(i)
#endif

  int j = 1;            // samples emitted so far
  uint32_t round = 0;   // exchange rounds done (mbarrier parity bookkeeping)
  while (j < m) {
    const uint32_t par = round & 1u;
    if (CLUSTERED && tid == 0)
      mbar_arrive_expect_tx(&mbar[par], static_cast<uint32_t>(csize * min(3, 32 / csize)) * 24u);

    // ---- A: apply the centres accepted last round; thread-local best and second --------------
    const int kc = ctr.count;
    for (int c = 0; c < kc; ++c) {
      const float4 o = ctr.c[c];
#pragma unroll
      for (int i = 0; i < P; ++i)
        pt[i] = fminf(sqdist3(px[i] - o.x, py[i] - o.y, pz[i] - o.z), pt[i]);
    }
    float best = -1.f, second = -1.f;
    int bslot = 0;
#pragma unroll
    for (int i = 0; i < P; ++i) {
      const float t = pt[i];
      if (t > best) { second = best; best = t; bslot = i; }
      else second = fmaxf(second, t);
    }
    PN2_T(0)
    // all threads have read ctr (count + centres) before warp 0 may overwrite it below
    // (ordering is provided by the __syncthreads that follows phase B)
    // ---- B: warp record = arg-max + best value among everything else in the warp -------------
    {
      const int vb = __float_as_int(best);  // best >= 0 or exactly -1.0f: int order == float order
      const uint32_t key = keybase + bslot;
      const int vmax = __reduce_max_sync(0xffffffffu, vb);
      const uint32_t kk = (vb == vmax) ? key : 0xffffffffu;
      const uint32_t kmin = __reduce_min_sync(0xffffffffu, kk);
      const bool win = (kk == kmin);  // exactly one lane: keys are unique
      const int sec = __reduce_max_sync(0xffffffffu, __float_as_int(win ? second : best));
      if (win) {
        float bx = px[0], by = py[0], bz = pz[0];
#pragma unroll
        for (int i = 1; i < P; ++i)
          if (bslot == i) { bx = px[i]; by = py[i]; bz = pz[i]; }
        w_vk[warp] = make_uint2(static_cast<uint32_t>(vmax), kmin);
        w_xyz[warp] = make_float4(bx, by, bz, __int_as_float(sec));
      }
    }
    PN2_T(1)
    __syncthreads();
    PN2_T(2)

    if (warp == 0) {
      // candidate list of this warp: one record per lane
      int cv = INT_MIN;                // value bits (INT_MIN = no record)
      uint32_t ck = 0xffffffffu;
      float cx = 0.f, cy = 0.f, cz = 0.f;
      int csec = INT_MIN;
      bool ok = true;
      if (lane < NW) {
        const uint2 r = w_vk[lane];
        const float4 w = w_xyz[lane];
        cv = static_cast<int>(r.x); ck = r.y; cx = w.x; cy = w.y; cz = w.z; csec = __float_as_int(w.w);
      }
      if (CLUSTERED) {
        // ---- C+D: the CTA's top-T warp records (T = 32 / cluster size, at most 3) go to every peer,
        // together with the CTA bound = best value among everything the CTA does NOT report.  More
        // candidates per exchange -> longer accepted batches -> fewer dependent exchange rounds.
        const int T = min(3, 32 / csize);
        int rem = cv;                       // bests still unreported
        int tv[3] = {INT_MIN, INT_MIN, INT_MIN};
        uint32_t tk[3] = {0xffffffffu, 0xffffffffu, 0xffffffffu};
        float tx[3] = {0.f, 0.f, 0.f}, ty[3] = {0.f, 0.f, 0.f}, tz[3] = {0.f, 0.f, 0.f};
        bool taken = false;
#pragma unroll
        for (int t = 0; t < 3; ++t) {
          if (t < T) {
            int rv = rem;
            uint32_t rk = ck;
            const int src = warp_argmax(rv, rk);
            tv[t] = rv; tk[t] = rk;
            tx[t] = __shfl_sync(0xffffffffu, cx, src); ty[t] = __shfl_sync(0xffffffffu, cy, src);
            tz[t] = __shfl_sync(0xffffffffu, cz, src);
            if (lane == src) { rem = INT_MIN; taken = true; }
          }
        }
        // reported warps contribute their "second", the others their (unreported) best
        const int cta_sec = __reduce_max_sync(0xffffffffu, taken ? csec : cv);
        if (lane < csize * T) {
          const int peer = lane / T, t = lane - peer * T;
          const int sv = t == 0 ? tv[0] : (t == 1 ? tv[1] : tv[2]);
          const uint32_t sk = t == 0 ? tk[0] : (t == 1 ? tk[1] : tk[2]);
          const float sx = t == 0 ? tx[0] : (t == 1 ? tx[1] : tx[2]);
          const float sy = t == 0 ? ty[0] : (t == 1 ? ty[1] : ty[2]);
          const float sz = t == 0 ? tz[0] : (t == 1 ? tz[1] : tz[2]);
          const uint32_t rbar = mapa_u32(smem_u32(&mbar[par]), peer);
          const uint32_t slot = rank * T + t;
          st_async_v4(mapa_u32(smem_u32(&cand[par][slot]), peer), static_cast<uint32_t>(sv), sk,
                      __float_as_uint(sx), __float_as_uint(sy), rbar);
          st_async_v2(mapa_u32(smem_u32(&cand_zs[par][slot]), peer), __float_as_uint(sz),
                      static_cast<uint32_t>(cta_sec), rbar);
        }
        PN2_T(3)
        ok = mbar_wait(&mbar[par], (round >> 1) & 1u);
        PN2_T(4)
        cv = INT_MIN; ck = 0xffffffffu; csec = INT_MIN;
        if (ok && lane < csize * T) {
          const uint4 c = cand[par][lane];
          const uint2 zs = cand_zs[par][lane];
          cv = static_cast<int>(c.x); ck = c.y; cx = __uint_as_float(c.z); cy = __uint_as_float(c.w);
          cz = __uint_as_float(zs.x); csec = static_cast<int>(zs.y);
        }
      }
      // ---- E: greedy acceptance on the candidates (exact, see header) --------------------------
      // Warp 0 is alone on the critical path here, so the loop carries only what the next
      // iteration needs: arg-max, broadcast of the winner's xyz, the candidates' update.  Lane k
      // keeps the k-th accepted record; decoding, the shared-memory publish and the global stores
      // happen once after the loop, one lane per sample.
      // bnd folds both acceptance conditions (value > bound and value > 0) into one compare
      const int bnd = max(__reduce_max_sync(0xffffffffu, csec), 0);
      const int room = min(KMAX, m - j);
      int k = 0;
      if (ok) {
        int rv = __reduce_max_sync(0xffffffffu, cv);
        if (rv < 0) {
          // no eligible point anywhere: every reference thread reports besti = 0 (key of index 0 is 0)
          if (lane == 0) ctr.c[0] = make_float4(x0, y0, z0, __uint_as_float(0u));
          k = 1;
        } else {
          for (;;) {
            uint32_t tie = __ballot_sync(0xffffffffu, cv == rv);
            if (tie & (tie - 1)) {  // equal values: the smallest key wins (keys are unique)
              const uint32_t kk = (cv == rv) ? ck : 0xffffffffu;
              const uint32_t kmin = __reduce_min_sync(0xffffffffu, kk);
              tie = __ballot_sync(0xffffffffu, kk == kmin);
            }
            const int src = __ffs(tie) - 1;
            const float ax = __shfl_sync(0xffffffffu, cx, src), ay = __shfl_sync(0xffffffffu, cy, src),
                        az = __shfl_sync(0xffffffffu, cz, src);
            const uint32_t ak = __shfl_sync(0xffffffffu, ck, src);
            if (lane == 0) ctr.c[k] = make_float4(ax, ay, az, __uint_as_float(ak));  // key rides in .w
            if (++k >= room) break;
            // lower every candidate -- the accepted one included -- by its exact distance to the new
            // centre: a finite point drops to 0 and can never be re-accepted in this round; a point
            // with inf/NaN coordinates keeps its value (d is NaN) exactly as in the reference.  Lanes
            // without a record hold -0.0f / -1.0f, which fminf leaves untouched.
            cv = __float_as_int(fminf(sqdist3(cx - ax, cy - ay, cz - az), __int_as_float(cv)));
            rv = __reduce_max_sync(0xffffffffu, cv);
            if (!(rv > bnd)) break;  // int compare == float compare for values >= 0
          }
        }
      }
      __syncwarp();
      if (ok && lane < k && rank == 0) {
        const float4 cc = ctr.c[lane];
        const uint32_t mykey = __float_as_uint(cc.w);
        const uint32_t row = mykey & rowmask;
        const uint32_t vv = log2bs ? __brev(mykey & ~rowmask) : 0u;
        idxs[j + lane] = static_cast<int>(vv + (row << log2bs));
        if (oxyz) { oxyz[3 * (j + lane)] = cc.x; oxyz[3 * (j + lane) + 1] = cc.y; oxyz[3 * (j + lane) + 2] = cc.z; }
      }
      if (lane == 0) ctr.count = ok ? k : -1;
    }
    __syncthreads();
    const int got = ctr.count;
    if (got < 0) {
      if (tid == 0) atomicExch(a.err, 1);
      break;
    }
    j += got;
    ++round;
    PN2_T(5)
#ifdef PN2_FPS_TIMING
    ++nrounds;
#endif
  }
#ifdef PN2_FPS_TIMING
  if (blockIdx.x == 0 && tid == 0 && a.dbg) {
    for (int i = 0; i < 6; ++i) a.dbg[i] = tacc[i];
    a.dbg[6] = nrounds;
  }
#endif
  // peers may still be storing into this CTA's shared memory: leave together
  if (CLUSTERED) cluster_sync_all();
}

template <int P, int THREADS, bool CLUSTERED>
int launch(const FpsArgs& a, int b, cudaStream_t stream, int* max_active) {
  auto kern = fps_kernel<P, THREADS, CLUSTERED>;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(static_cast<unsigned>(b) * (CLUSTERED ? a.csize : 1));
  cfg.blockDim = dim3(THREADS);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  int nattr = 0;
  if (CLUSTERED) {
    if (a.csize > 8) PN2_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = a.csize;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    nattr = 1;
  }
  cfg.attrs = attr;
  cfg.numAttrs = nattr;
  if (max_active) {  // occupancy probe only
    PN2_CUDA(cudaOccupancyMaxActiveClusters(max_active, kern, &cfg));
    return PN2_OK;
  }
  PN2_CUDA(cudaLaunchKernelEx(&cfg, kern, a));
  return PN2_OK;
}

template <int THREADS, bool CLUSTERED>
int dispatch_p(int p, const FpsArgs& a, int b, cudaStream_t s, int* max_active) {
  switch (p) {
    case 1: return launch<1, THREADS, CLUSTERED>(a, b, s, max_active);
    case 2: return launch<2, THREADS, CLUSTERED>(a, b, s, max_active);
    case 3: return launch<3, THREADS, CLUSTERED>(a, b, s, max_active);
    case 4: return launch<4, THREADS, CLUSTERED>(a, b, s, max_active);
    case 5: return launch<5, THREADS, CLUSTERED>(a, b, s, max_active);
    case 6: return launch<6, THREADS, CLUSTERED>(a, b, s, max_active);
    case 7: case 8: return launch<8, THREADS, CLUSTERED>(a, b, s, max_active);
    case 9: case 10: return launch<10, THREADS, CLUSTERED>(a, b, s, max_active);
    case 11: case 12: return launch<12, THREADS, CLUSTERED>(a, b, s, max_active);
    case 13: case 14: case 15: case 16: return launch<16, THREADS, CLUSTERED>(a, b, s, max_active);
    default: break;
  }
  set_error("furthest_point_sampling: n too large for the register-resident kernel (P=%d)", p);
  return PN2_ERR_INVALID;
}

#ifndef PN2_FPS_TIMING
}  // namespace
#endif

int* device_err_flag();  // abi.cu
bool fps_bucket_supported(int n);  // fps_bucket.cu
int fps_bucket_run(int b, int n, int m, int log2bs, const float* dataset, int* idxs, float* new_xyz,
                   cudaStream_t stream);
#ifdef PN2_FPS_TIMING
long long* fps_dbg_ptr = nullptr;
#endif

// Tunables (read once): PN2_FPS_CLUSTER = 1 | 2..16, PN2_FPS_THREADS = 512 | 1024.
static int env_int(const char* name, int dflt) {
  const char* s = getenv(name);
  return s ? atoi(s) : dflt;
}

static int run_cfg(int cluster, int threads, FpsArgs a, int bs, int b, cudaStream_t stream, int* max_active) {
  a.csize = cluster;
  const int segs = cluster * threads / bs;  // >= 1: threads >= 512 >= bs
  const int p = (a.rows + segs - 1) / segs;
  if (cluster == 1) {
    if (threads == 512) return dispatch_p<512, false>(p, a, b, stream, max_active);
    return dispatch_p<1024, false>(p, a, b, stream, max_active);
  }
  if (threads == 512) return dispatch_p<512, true>(p, a, b, stream, max_active);
  return dispatch_p<1024, true>(p, a, b, stream, max_active);
}

int fps_prefix_check(int b, int n, int m, int log2bs, const float* dataset, void* ws, const int** verdict,
                     cudaStream_t stream);  // fps_prefix.cu

int fps_run(int b, int n, int m, const float* dataset, int* idxs, float* new_xyz, cudaStream_t stream,
            bool maybe_in_order = false) {
  if (m <= 0 || b <= 0) return PN2_OK;  // sampling_gpu.cu:78
  PN2_REQUIRE(n > 0, "furthest_point_sampling: n must be positive (got %d)", n);
  PN2_REQUIRE(dataset && idxs, "furthest_point_sampling: null pointer");
  const int bs = opt_n_threads(n);
  int log2bs = 0;
  while ((1 << log2bs) < bs) ++log2bs;
  FpsArgs a;
  a.xyz = dataset; a.idxs = idxs; a.new_xyz = new_xyz;
  a.n = n; a.m = m; a.log2bs = log2bs; a.rows = (n + bs - 1) / bs; a.csize = 1;
  a.err = device_err_flag();
  a.in_order = nullptr;
  a.dbg = nullptr;
#ifdef PN2_FPS_TIMING
  if (!fps_dbg_ptr) cudaMalloc(&fps_dbg_ptr, 8 * sizeof(long long));
  a.dbg = fps_dbg_ptr;
#endif
  PN2_REQUIRE(a.err != nullptr, "furthest_point_sampling: cannot allocate the device error flag");

  // Two algorithms for large scenes (both bit-exact):
  //   cluster kernel (default)  16 SMs per scene, lowest latency for a single batch;
  //   bucketed kernel           1 SM per scene with spatial pruning: ~2x the latency but 1/16 of the
  //                             SM footprint, the right choice when several batches are in flight.
  // Selected with pn2_set_option("fps_bucket", 0|1) (initial value: env PN2_FPS_BUCKET, default 0).
  if (option_value(OPT_FPS_BUCKET) && fps_bucket_supported(n)) {
    stat_add(STAT_FPS_BUCKET);
    return fps_bucket_run(b, n, m, log2bs, dataset, idxs, new_xyz, stream);
  }

  static const int force_cluster = env_int("PN2_FPS_CLUSTER", 0);
  static const int force_threads = env_int("PN2_FPS_THREADS", 0);
  int cluster, threads;
  if (force_cluster) {
    cluster = force_cluster;
    threads = force_threads ? force_threads : 512;
  } else if (n <= 8192) {
    cluster = 1;
    threads = force_threads ? force_threads : (n <= 2048 ? 512 : 1024);
    if (maybe_in_order && m >= 2 && m <= n && m <= 8192) {
      // the caller believes the input is a previous FPS output: verify that in parallel (fps_prefix.cu);
      // the kernel below then only writes 0..m-1 for the scenes where it holds
      void* ws = nullptr;
      PN2_CUDA(cudaMallocAsync(&ws, static_cast<size_t>(b) * m * sizeof(float) + b * sizeof(int), stream));
      int rc = fps_prefix_check(b, n, m, log2bs, dataset, ws, &a.in_order, stream);
      if (rc == PN2_OK) rc = run_cfg(cluster, threads, a, bs, b, stream, nullptr);
      cudaFreeAsync(ws, stream);
      return rc;
    }
  } else {
    // The largest cluster whose co-resident count covers the whole batch in ONE wave (a B200 GPC
    // holds a single 16-CTA cluster and one GPC is short of 16 SMs: 7 such clusters fit, 15 of 8).
    threads = force_threads ? force_threads : 512;
    static int active_cache[MAXC + 1] = {0};
    cluster = 8;
    for (int cs = 16; cs >= 8; cs -= 2) {
      if (a.rows > cs * (threads / bs) * 16) continue;  // would not fit the register budget
      int cached = __atomic_load_n(&active_cache[cs], __ATOMIC_RELAXED);  // idempotent: racing threads store the same value
      if (!cached) {
        int act = 0;
        if (run_cfg(cs, threads, a, bs, b, stream, &act) != PN2_OK || act <= 0) act = -1;
        __atomic_store_n(&active_cache[cs], act, __ATOMIC_RELAXED);
        cached = act;
      }
      if (cached >= b) { cluster = cs; break; }
    }
    if (a.rows > cluster * (threads / bs) * 16) { cluster = 16; threads = 1024; }
  }
  PN2_REQUIRE(cluster >= 1 && cluster <= MAXC && (threads == 512 || threads == 1024),
              "furthest_point_sampling: unsupported cluster=%d threads=%d", cluster, threads);
  stat_add(STAT_FPS_CLUSTER);
  return run_cfg(cluster, threads, a, bs, b, stream, nullptr);
}

}  // namespace pn2

#ifdef PN2_FPS_TIMING
// developer build only (tools/fps_timing.py): per-phase cycle sums + round count of the last launch
extern "C" int pn2_fps_timing(long long* out7) {
  if (!pn2::fps_dbg_ptr) return 1;
  cudaDeviceSynchronize();
  return cudaMemcpy(out7, pn2::fps_dbg_ptr, 7 * sizeof(long long), cudaMemcpyDeviceToHost) != cudaSuccess;
}
#endif

extern "C" int pn2_furthest_point_sampling(int b, int n, int m, const float* dataset, float* temp,
                                           int* idxs, void* stream) {
  PN2_RANGE("pn2_furthest_point_sampling");
  (void)temp;  // running minima stay in registers; the reference's scratch is not needed
  return pn2::fps_run(b, n, m, dataset, idxs, nullptr, pn2::as_stream(stream));
}

extern "C" int pn2_furthest_point_sampling_gather(int b, int n, int m, const float* dataset, int* idxs,
                                                  float* new_xyz, void* stream) {
  PN2_RANGE("pn2_furthest_point_sampling_gather");
  return pn2::fps_run(b, n, m, dataset, idxs, new_xyz, pn2::as_stream(stream));
}

extern "C" int pn2_furthest_point_sampling_gather_ordered(int b, int n, int m, const float* dataset, int* idxs,
                                                          float* new_xyz, void* stream) {
  PN2_RANGE("pn2_furthest_point_sampling_gather_ordered");
  return pn2::fps_run(b, n, m, dataset, idxs, new_xyz, pn2::as_stream(stream), true);
}
