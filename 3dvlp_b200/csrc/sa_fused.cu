// Fused set-abstraction body: grouping + shared MLP (3 layers) + max-pool in ONE kernel (sm_100a).
//
// Reference path (pointnet2_modules.py:243-257): QueryAndGroup materialises the grouped tensor
// (B, 3+C, npoint, nsample) in HBM (group_points_gpu.cu:13-33, 70.8 MB/scene for SA1 at 135
// channels), then three cuDNN 1x1 convolutions + BatchNorm + ReLU read and write (B, C, npoint,
// nsample) activations, then max_pool2d.  Here nothing but the pooled (B, C3, npoint) result leaves
// the SM:
//
//   layer 1  is split algebraically.  A 1x1 convolution commutes with the neighbourhood gather, so
//            the feature half  W1[:,3:] . f  is evaluated once per INPUT point (pn2_linear_bf16 ->
//            `g`, or directly from <= 4 raw channels) and only gathered here; the geometric half
//            W1[:,0:3] . (p - centre)/radius  is three FMAs per channel in fp32.  One warp builds
//            32 rows of the 128-row tile: lane = channels, rows broadcast by shuffle, the row of g
//            is read coalesced, weights stay in registers.  Result -> ReLU -> bf16 -> shared memory
//            in the SWIZZLE_128B K-major operand layout.
//   layer 2  D2[128 rows x C2] = X1 . W2^T   tcgen05.mma, rows on the TMEM lanes; the epilogue
//            (thread = row) adds bias, ReLU, rounds to bf16 and overwrites the tile as X2.
//   layer 3  transposed: D3[C3 channels x 128 rows] = W3 . X2^T, i.e. the weights are the A operand
//            and the activation tile the B operand (both K-major, same shared-memory image), so
//            output channels land on TMEM lanes and the nsample neighbours of a centre are
//            consecutive TMEM columns: max-pooling is a register-only reduction after tcgen05.ld.
//            bias + ReLU commute with max and are applied once per (channel, centre).
//
// Persistent CTAs (grid = #SMs x resident CTAs), folded weights fetched once per CTA with TMA bulk
// copies (cp.async.bulk + mbarrier), accumulators in TMEM (D3 aliases D2).
#include "tc.cuh"

namespace pn2 {
namespace {

// max of three floats in one instruction (FMNMX3, sm_100+); NaN operands are ignored like fmaxf
__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}

constexpr int SA_THREADS = 128;
constexpr int kMaxScenes = 64;  // scenes per launch (the tile prefix scan holds two per lane of one warp)

struct SaArgs {
  const float* xyz;      // (B, N, 3)
  const float* new_xyz;  // (B, npoint, 3)
  const int* idx;        // (B, npoint, ns)
  const float* g;        // (B*N, C1) fp32 or nullptr
  const float* feat;     // (B, CD, N) fp32 or nullptr
  const float* w1;       // (C1, 3 + CD) fp32, BatchNorm folded
  const float* b1;       // (C1)
  const uint8_t* w2p;    // packed (C2, C1)
  const float* b2;
  const uint8_t* w3p;    // packed (C3, C2)
  const float* b3;
  float* out_cm;         // (B, C3, npoint) fp32
  int* err;
  int* tile_counter;     // next tile, zeroed by sa_count_kernel: tiles are handed out dynamically
  const int* row_src;    // (B, cap) compacted rows: index into idx (centre * ns + slot) or -1; see sa_plan_kernel
  const int* tiles_b;    // (B) tiles of each scene's compacted row list
  int cap;               // row capacity of a scene's region: npoint * ns rounded up to 128
  int B, N, npoint, ns, ns_shift, C1, C2, C3;  // ns = 1 << ns_shift
  float radius_div;      // divisor applied to the recentred xyz (radius, or 1 when not normalised)
  long long total_rows;  // B * npoint * ns
};

// NG = independent 128-thread groups per CTA.  Each group runs the tile loop on its own activation tile,
// TMEM columns and mbarrier; the folded weights in shared memory are shared.  When the weights leave room
// for only one CTA per SM (SA2-SA4: 96 KB) two groups give the SM a second tile to work on while the
// first waits for its gather / MMA / TMEM round trips.
//
// MMA1 (raw-channel mode, no g): layer 1 runs on the tensor cores too.  Its fp32 inputs
// v = ((p - centre)/radius, raw channels) and fp32 weights are each split into THREE bf16 terms
// (h = bf16(v), m = bf16(v - h), l = bf16(v - h - m): 24 mantissa bits); the six products whose
// weight is >= 2^-16 of w.v  (w_h v_h, w_h v_m, w_m v_h, w_h v_l, w_l v_h, w_m v_m)  and the bias
// (b_h + b_m + b_l against three columns of ones) become one K = 32 or 48 GEMM whose operand a thread
// builds for its own row: ~60 instructions per row instead of ~24 per row AND lane.  Every product is
// exact in the fp32 accumulator and the dropped terms are ~2^-24 relative, i.e. the result matches the
// fp32 FMA chain to fp32 rounding -- the bf16 rounding that follows sees the same values.
template <int CPL, int CD, int NG, bool MMA1>  // C1 = 32 * CPL, raw feature channels, groups, layer 1 on MMA
__global__ void __maxnreg__(128) sa_fused_kernel(const SaArgs a) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_w, bar_mma_g[NG];
  __shared__ uint32_t tmem_slot;
  __shared__ int s_next[NG][2];  // next tile of each group, double-buffered by iteration parity
  __shared__ uint32_t s_end[NG][2][4];    // bit r: row r of the group's tile is the last row of its centre in this tile
  __shared__ int s_jg0[NG][2];            // global centre index of the tile's row 0
  __shared__ int s_tpre[kMaxScenes + 1];  // prefix sums of tiles_b: tile id -> scene
  __shared__ float s_b2[128];

#ifdef PN2_SA_TIMING
  const long long dbg_entry = clock64();
  long long dbg_p[5] = {0, 0, 0, 0, 0};
#define SA_P(i) dbg_p[i] = clock64() - dbg_entry;
  unsigned long long dbg_g0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(dbg_g0));
#endif
  constexpr int C1 = 32 * CPL;
  const int grp = threadIdx.x / SA_THREADS;            // group of this thread
  const int tid = threadIdx.x % SA_THREADS, lane = tid & 31, warp = tid >> 5;  // within the group
  const int C2 = a.C2, C3 = a.C3, ns = a.ns;
  uint64_t& bar_mma = bar_mma_g[grp];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t w2_bytes = static_cast<uint32_t>(C2) * C1 * 2, w3_bytes = static_cast<uint32_t>(C3) * C2 * 2;
  const uint32_t x_bytes = 128u * static_cast<uint32_t>(C1 > C2 ? C1 : C2) * 2u;
  constexpr int NT = 6 * (3 + CD) + 3;          // layer-1 product terms (MMA1): 6 per input + 3 for the bias
  constexpr int K0 = (NT + 15) / 16 * 16;        // 32 (CD <= 1) or 48: inside the first 64-wide K panel
  const uint32_t w1_bytes = MMA1 ? static_cast<uint32_t>(C1) * 128u : 0u;
  const uint32_t s_w2 = base, s_w3 = base + w2_bytes, s_w1 = s_w3 + w3_bytes,
                 s_x = s_w1 + w1_bytes + grp * x_bytes;  // 1 KB multiples
  constexpr uint32_t group_cols = 128u;  // D2 (C2 <= 128 columns) and each 128-channel half of D3
  const uint32_t tmem_cols = group_cols * NG < 32u ? 32u : group_cols * NG;

  if (threadIdx.x == 0) {
    mbar_init(&bar_w, 1);
    for (int g = 0; g < NG; ++g) mbar_init(&bar_mma_g[g], 1);
    fence_mbar_init_cluster();
  }
  for (int i = threadIdx.x; i < C2; i += SA_THREADS * NG) s_b2[i] = a.b2[i];
  if (MMA1) {
    // B operand of layer 1: row = channel, K-major: [w_h, w_h, w_m, w_h, w_l, w_m] per input (against
    // [v_h, v_m, v_h, v_l, v_h, v_m]), [b_h, b_m, b_l] against ones, zeros up to K0
    // the weights of a batch of WB elements are requested before the first is converted: the loop costs one or
    // two global round trips per CTA instead of one per element (16 -- 13 k cycles, more than a tile -- in SA1)
    constexpr int WB = 16;
    auto w1_term = [&](int e, int& part) -> const float* {  // element e of the panel: which weight, which split term
      const int ch = e / K0, k = e - ch * K0;
      part = 0;  // 0 = h, 1 = m, 2 = l
      if (k < NT - 3) {
        const int t = k % 6;
        part = t == 2 || t == 5 ? 1 : (t == 4 ? 2 : 0);
        return a.w1 + static_cast<size_t>(ch) * (3 + CD) + k / 6;
      }
      if (k < NT) { part = k - (NT - 3); return a.b1 + ch; }
      return nullptr;  // zero padding up to K0
    };
    for (int e0 = threadIdx.x; e0 < C1 * K0; e0 += SA_THREADS * NG * WB) {
      float wv[WB];
#pragma unroll
      for (int u = 0; u < WB; ++u) {
        const int e = e0 + u * SA_THREADS * NG;
        int part;
        const float* src = e < C1 * K0 ? w1_term(e, part) : nullptr;
        wv[u] = src ? __ldg(src) : 0.f;
      }
#pragma unroll
      for (int u = 0; u < WB; ++u) {
        const int e = e0 + u * SA_THREADS * NG;
        if (e >= C1 * K0) break;
        int part;
        w1_term(e, part);
        const float w = wv[u];
        const __nv_bfloat16 h = __float2bfloat16_rn(w);
        const float r1 = w - __bfloat162float(h);
        const __nv_bfloat16 m = __float2bfloat16_rn(r1);
        const __nv_bfloat16 l = __float2bfloat16_rn(r1 - __bfloat162float(m));
        const __nv_bfloat16 val = part == 0 ? h : (part == 1 ? m : l);
        const int ch = e / K0, k = e - ch * K0;
        const uint32_t addr = s_w1 + tc::swz_off(ch, static_cast<uint32_t>(k) >> 3) + (k & 7) * 2;
        asm volatile("st.shared.b16 [%0], %1;" ::"r"(addr), "h"(__bfloat16_as_ushort(val)) : "memory");
      }
    }
    fence_proxy_async_smem();
  }
#ifdef PN2_SA_TIMING
  SA_P(0)
#endif
  // Tiles come from a global counter, not from a static stride: when other kernels hold some SMs
  // (several batches in flight) part of the grid starts late, and a late CTA must find the work
  // already done instead of a full static share of it.
  // the first tile of a group is static (no round trip before the first gather); the rest come from the counter
  if (tid == 0) s_next[grp][0] = blockIdx.x * NG + grp;
  if (threadIdx.x < 32) {
    // prefix sums of the scenes' tile counts by one warp: a single global round trip (kMaxScenes = 64 = 2 per lane)
    const int l = threadIdx.x;
    int i0 = l < a.B ? __ldg(a.tiles_b + l) : 0, i1 = l + 32 < a.B ? __ldg(a.tiles_b + l + 32) : 0;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int o0 = __shfl_up_sync(0xffffffffu, i0, d), o1 = __shfl_up_sync(0xffffffffu, i1, d);
      if (l >= d) { i0 += o0; i1 += o1; }
    }
    i1 += __shfl_sync(0xffffffffu, i0, 31);
    if (l == 0) s_tpre[0] = 0;
    if (l < a.B) s_tpre[l + 1] = i0;
    if (l + 32 < a.B) s_tpre[l + 33] = i1;
#ifdef PN2_SA_TIMING
    SA_P(1)
#endif
    tc::tmem_alloc(&tmem_slot, tmem_cols);
#ifdef PN2_SA_TIMING
    SA_P(2)
#endif
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
#ifdef PN2_SA_TIMING
  SA_P(3)
#endif
  const uint32_t tmem_base = tmem_slot;
  const uint32_t tmem_d = tmem_base + grp * group_cols;
  auto group_sync = [&]() { asm volatile("bar.sync %0, %1;" ::"r"(1 + grp), "n"(SA_THREADS) : "memory"); };
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(&bar_w, w2_bytes + w3_bytes);
    uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
    tc::bulk_g2s(gen, a.w2p, w2_bytes, &bar_w);
    tc::bulk_g2s(gen + w2_bytes, a.w3p, w3_bytes, &bar_w);
  }

  // layer-1 weights of this lane's channels, in registers for the whole kernel
  float wx[CPL], wy[CPL], wz[CPL], wb[CPL], wf[CPL][CD > 0 ? CD : 1];
#pragma unroll
  for (int q = 0; q < CPL && !MMA1; ++q) {
    const int c = lane * CPL + q;
    const float* row = a.w1 + static_cast<size_t>(c) * (3 + CD);
    wx[q] = row[0]; wy[q] = row[1]; wz[q] = row[2]; wb[q] = a.b1[c];
#pragma unroll
    for (int d = 0; d < CD; ++d) wf[q][d] = row[3 + d];
  }

  const uint32_t idesc1 = tc::idesc_bf16_f32(128, C1);
  const uint32_t idesc2 = tc::idesc_bf16_f32(128, C2);
  const uint32_t idesc3 = tc::idesc_bf16_f32(128, 128);
  const long long n_tiles = s_tpre[a.B];  // tiles of the compacted row lists of all scenes
  uint32_t phase = 0;
  bool weights_ready = false, failed = false;

  // Inputs of this thread's row: compacted row -> (centre, slot) -> neighbour index -> point coordinates,
  // centre, raw channels: a dependent chain of global round trips, requested for the NEXT tile while the
  // current one is being processed.
  float in_p[3], in_c[3], in_f[CD > 0 ? CD : 1];
  long long in_grow = 0;
  bool in_ok = false;
  int in_scene = 0;
  // The chain is requested in three stages, each consumed a phase after it was issued, so that none of its two
  // dependent global round trips is ever waited for on the spot:
  //   A (after the tile's first group_sync)  compacted-row entries of the NEXT tile: row_src[tid], row_src[tid + 1]
  //   B (before the layer-3 group_sync)      -> centre id, end-of-centre mask; requests idx[src] and the centre
  //   C (after the layer-3 MMA is issued)    -> point id; requests the point's coordinates / raw channels
  int nx_src = -1, nx_nxt = -1, nx_pi = 0, nx_jg = -1;
  bool nx_valid = false;
  auto fetch_a = [&](long long t) {
    nx_valid = t < n_tiles;  // uniform over the group
    nx_src = nx_nxt = -1;
    if (!nx_valid) return;
    int b = in_scene;                // a group's tiles only grow: search on from the current scene
    while (s_tpre[b + 1] <= t) ++b;  // scene of the tile (few scenes: linear search)
    in_scene = b;
    const int* rs = a.row_src + static_cast<long long>(b) * a.cap + (t - s_tpre[b]) * 128;
    nx_src = __ldg(rs + tid);
    nx_nxt = tid < 127 ? __ldg(rs + tid + 1) : -1;
  };
  auto fetch_b = [&](int par) {
    in_ok = false;
    in_grow = 0;
#pragma unroll
    for (int d = 0; d < 3; ++d) in_p[d] = in_c[d] = 0.f;
#pragma unroll
    for (int d = 0; d < (CD > 0 ? CD : 1); ++d) in_f[d] = 0.f;
    if (!nx_valid) return;
    const int src = nx_src;
    nx_jg = src >= 0 ? (src >> a.ns_shift) : -1;  // global centre index b*npoint + j (< 2^31, host-checked)
    // a centre may continue in the next tile: each tile flushes its part with an atomic max
    const bool last = src >= 0 && (nx_nxt < 0 || (nx_nxt >> a.ns_shift) != nx_jg);
    const uint32_t m = __ballot_sync(0xffffffffu, last);
    if (lane == 0) s_end[grp][par][warp] = m;
    if (tid == 0) s_jg0[grp][par] = nx_jg;  // row 0 of a claimed tile always exists
    in_ok = src >= 0;
    if (in_ok) {
      nx_pi = __ldg(a.idx + src);
      const float* c = a.new_xyz + static_cast<long long>(nx_jg) * 3;
#pragma unroll
      for (int d = 0; d < 3; ++d) in_c[d] = __ldg(c + d);
    }
  };
  auto fetch_c = [&]() {
    if (!nx_valid || !in_ok) return;
    const int pi = nx_pi;
    in_grow = static_cast<long long>(in_scene) * a.N + pi;  // row of g / point id within the batch-flattened cloud
    const float* p = a.xyz + in_grow * 3;
#pragma unroll
    for (int d = 0; d < 3; ++d) in_p[d] = __ldg(p + d);
#pragma unroll
    for (int d = 0; d < CD; ++d) in_f[d] = __ldg(a.feat + (static_cast<size_t>(in_scene) * CD + d) * a.N + pi);
  };
  long long tile = s_next[grp][0], next_tile = n_tiles;
  fetch_a(tile); fetch_b(0); fetch_c();  // the first tile's inputs: requested back to back
#ifdef PN2_SA_TIMING
  SA_P(4)
#endif
#ifdef PN2_SA_TIMING
  long long dbg_t[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, dbg_prev = clock64(), dbg_start = dbg_prev;
  int dbg_tiles = 0, dbg_nld = 0;
  long long dbg_ld = 0;
#define SA_T(i) { const long long _t = clock64(); dbg_t[i] += _t - dbg_prev; dbg_prev = _t; }
#else
#define SA_T(i)
#endif

  for (int it = 0; tile < n_tiles && !failed; ++it, tile = next_tile) {
    SA_T(0)
#ifdef PN2_SA_TIMING
    ++dbg_tiles;
#endif
    const int scene = in_scene;  // of THIS tile (fetch_row for the next one overwrites these)
    // claim the tile after this one; the group reads it after the next group_sync (the other parity
    // slot was last read an iteration ago, several group_syncs back)
    if (tid == 0) s_next[grp][(it + 1) & 1] = static_cast<int>(gridDim.x) * NG + atomicAdd(a.tile_counter, 1);
    // pointnet2_utils.py:345-348: grouped_xyz -= new_xyz; grouped_xyz /= radius  (fp32 divide)
    const bool ok = in_ok;
    const long long grow = in_grow;
    const float rx = ok ? __fdiv_rn(in_p[0] - in_c[0], a.radius_div) : 0.f;
    const float ry = ok ? __fdiv_rn(in_p[1] - in_c[1], a.radius_div) : 0.f;
    const float rz = ok ? __fdiv_rn(in_p[2] - in_c[2], a.radius_div) : 0.f;
    float fd[CD > 0 ? CD : 1];
#pragma unroll
    for (int d = 0; d < (CD > 0 ? CD : 1); ++d) fd[d] = in_f[d];
    if (MMA1) {
      // ---------------- layer 1 on the tensor cores: thread = row builds its split-bf16 operand ----
      float v[3 + CD];
      v[0] = rx; v[1] = ry; v[2] = rz;
#pragma unroll
      for (int d = 0; d < CD; ++d) v[3 + d] = fd[d];
      uint32_t e[K0];  // bf16 bit patterns of the row: [h, m, h, l, h, m] per input, 1, 1, 1, zeros
#pragma unroll
      for (int k = 0; k < K0; ++k) e[k] = 0u;
#pragma unroll
      for (int d = 0; d < 3 + CD; ++d) {
        const __nv_bfloat16 h = __float2bfloat16_rn(v[d]);
        const float r1 = v[d] - __bfloat162float(h);
        const __nv_bfloat16 m = __float2bfloat16_rn(r1);
        const __nv_bfloat16 l = __float2bfloat16_rn(r1 - __bfloat162float(m));
        e[6 * d] = e[6 * d + 2] = e[6 * d + 4] = __bfloat16_as_ushort(h);
        e[6 * d + 1] = e[6 * d + 5] = __bfloat16_as_ushort(m);
        e[6 * d + 3] = __bfloat16_as_ushort(l);
      }
      e[NT - 3] = e[NT - 2] = e[NT - 1] = 0x3F80u;  // 1.0
#pragma unroll
      for (int ck = 0; ck < K0 / 8; ++ck) {
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(s_x + tc::swz_off(tid, ck)),
                     "r"(e[8 * ck] | (e[8 * ck + 1] << 16)), "r"(e[8 * ck + 2] | (e[8 * ck + 3] << 16)),
                     "r"(e[8 * ck + 4] | (e[8 * ck + 5] << 16)), "r"(e[8 * ck + 6] | (e[8 * ck + 7] << 16))
                     : "memory");
      }
      fence_proxy_async_smem();
      tc::fence_before_sync();
      SA_T(1)
      group_sync();
      if (!weights_ready) {
        if (!mbar_wait(&bar_w, 0)) { failed = true; break; }
        weights_ready = true;
      }
      next_tile = s_next[grp][(it + 1) & 1];  // claimed by thread 0 before the group_sync above
      fetch_a(next_tile);
      SA_T(2)
      if (tid == 0) {
        tc::fence_after_sync();
#pragma unroll
        for (int kk = 0; kk < K0 / 16; ++kk)
          tc::mma_bf16(tmem_d, tc::smem_desc_sw128(s_x + kk * 32), tc::smem_desc_sw128(s_w1 + kk * 32), idesc1,
                       kk > 0 ? 1u : 0u);
        tc::mma_commit(&bar_mma);
      }
      if (!mbar_wait(&bar_mma, phase)) { failed = true; break; }
      phase ^= 1;
      tc::fence_after_sync();
      SA_T(3)
      // epilogue 1: thread = row; ReLU folded into the bf16 conversion; X1 overwrites the operand rows
      const uint32_t lane_base = tmem_d + (static_cast<uint32_t>(warp * 32) << 16);
#pragma unroll
      for (int c0 = 0; c0 < C1; c0 += 16) {
        float t[16];
        tc::tmem_ld16(lane_base + c0, t);
        const uint32_t chunk = c0 >> 3;
        const uint32_t pa = s_x + (chunk >> 3) * (128u * 128u);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(pa + tc::swz_off(tid, chunk & 7u)),
                     "r"(tc::pack_bf16_relu(t[0], t[1])), "r"(tc::pack_bf16_relu(t[2], t[3])),
                     "r"(tc::pack_bf16_relu(t[4], t[5])), "r"(tc::pack_bf16_relu(t[6], t[7]))
                     : "memory");
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(pa + tc::swz_off(tid, (chunk + 1) & 7u)),
                     "r"(tc::pack_bf16_relu(t[8], t[9])), "r"(tc::pack_bf16_relu(t[10], t[11])),
                     "r"(tc::pack_bf16_relu(t[12], t[13])), "r"(tc::pack_bf16_relu(t[14], t[15]))
                     : "memory");
      }
    } else {
      // ---------------- layer 1: build X1 (bf16) for rows warp*32 .. warp*32+31 -------------------
      // rows in batches of GB: in gather mode the batch's g rows are all requested before the first is
      // used (the loop is bound by that latency: 4 loads in flight per warp left the kernel stalled
      // on `v += g` for a third of its samples)
      constexpr int GB = CPL == 4 ? 8 : 16;
      constexpr int GW = CPL == 4 ? 4 : 2;
#pragma unroll 1
      for (int i0 = 0; i0 < 32; i0 += GB) {
        float gq[CD == 0 ? GB : 1][GW];
        if (CD == 0 && a.g != nullptr) {  // g and raw channels are alternatives (the host picks one)
#pragma unroll
          for (int u = 0; u < GB; ++u) {
            const long long gr = __shfl_sync(0xffffffffu, grow, i0 + u);  // rows past the end carry row 0
            if (CPL == 4) {
              const float4 t = __ldg(reinterpret_cast<const float4*>(a.g + gr * C1) + lane);
              gq[CD == 0 ? u : 0][0] = t.x; gq[CD == 0 ? u : 0][1] = t.y; gq[CD == 0 ? u : 0][GW - 2] = t.z; gq[CD == 0 ? u : 0][GW - 1] = t.w;
            } else {
              const float2 t = __ldg(reinterpret_cast<const float2*>(a.g + gr * C1) + lane);
              gq[CD == 0 ? u : 0][0] = t.x; gq[CD == 0 ? u : 0][1] = t.y;
            }
          }
        }
#pragma unroll
        for (int u = 0; u < GB; ++u) {
          const int i = i0 + u;
          const float x = __shfl_sync(0xffffffffu, rx, i), y = __shfl_sync(0xffffffffu, ry, i),
                      z = __shfl_sync(0xffffffffu, rz, i);
          const bool rok = __shfl_sync(0xffffffffu, static_cast<int>(ok), i) != 0;
          float v[CPL];
#pragma unroll
          for (int q = 0; q < CPL; ++q) v[q] = __fmaf_rn(wz[q], z, __fmaf_rn(wy[q], y, __fmaf_rn(wx[q], x, wb[q])));
          if (CD > 0) {
#pragma unroll
            for (int d = 0; d < CD; ++d) {
              const float f = __shfl_sync(0xffffffffu, fd[d], i);
#pragma unroll
              for (int q = 0; q < CPL; ++q) v[q] = __fmaf_rn(wf[q][d], f, v[q]);
            }
          }
          if (CD == 0 && a.g != nullptr) {
#pragma unroll
            for (int q = 0; q < CPL; ++q) v[q] += gq[CD == 0 ? u : 0][q];
          }
          const uint32_t row = warp * 32 + i;
          const uint32_t c0 = lane * CPL;              // first channel of this lane
          const uint32_t chunk = c0 >> 3;               // 16-byte chunk along K
          const uint32_t addr = s_x + (chunk >> 3) * (128u * 128u) + tc::swz_off(row, chunk & 7u) + (c0 & 7u) * 2u;
          // ReLU folded into the conversion; rows past the end of the batch are stored as zeros
          if (CPL == 4) {
            asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(rok ? tc::pack_bf16_relu(v[0], v[1]) : 0u),
                         "r"(rok ? tc::pack_bf16_relu(v[2], v[3]) : 0u)
                         : "memory");
          } else {
            asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(rok ? tc::pack_bf16_relu(v[0], v[1]) : 0u) : "memory");
          }
        }
      }
    }
    fence_proxy_async_smem();
    tc::fence_before_sync();
    SA_T(4)
    group_sync();
    if (!weights_ready) {
      if (!mbar_wait(&bar_w, 0)) { failed = true; break; }
      weights_ready = true;
    }
    if (!MMA1) {  // this is the tile's first group_sync in gather mode
      next_tile = s_next[grp][(it + 1) & 1];
      fetch_a(next_tile);
    }
    SA_T(5)
    // ---------------- layer 2: D2 = X1 . W2^T --------------------------------------------------
    if (tid == 0) {
      tc::fence_after_sync();
#pragma unroll
      for (int kk = 0; kk < C1 / 16; ++kk) {
        const uint32_t po = (kk >> 2), ko = (kk & 3) * 32;
        tc::mma_bf16(tmem_d, tc::smem_desc_sw128(s_x + po * (128u * 128u) + ko),
                     tc::smem_desc_sw128(s_w2 + po * (static_cast<uint32_t>(C2) * 128u) + ko), idesc2, kk > 0 ? 1u : 0u);
      }
      tc::mma_commit(&bar_mma);
    }
    // next tile: centre ids / end mask from the row entries (requested a phase ago), request idx and the centre --
    // here, not after the epilogue, so that idx has the whole epilogue to arrive before fetch_c needs it
    fetch_b((it + 1) & 1);
    if (!mbar_wait(&bar_mma, phase)) { failed = true; break; }
    phase ^= 1;
    tc::fence_after_sync();
    SA_T(6)
    // ---------------- epilogue 2: thread = row; bias + ReLU -> bf16 X2 over X1 ----------------
    {
      const uint32_t lane_base = tmem_d + (static_cast<uint32_t>(warp * 32) << 16);
      for (int c0 = 0; c0 < C2; c0 += 16) {
        float v[16];
        tc::tmem_ld16(lane_base + c0, v);
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] += s_b2[c0 + i];  // the ReLU rides on the bf16 conversion
        const uint32_t chunk = c0 >> 3;
        const uint32_t pa = s_x + (chunk >> 3) * (128u * 128u);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(pa + tc::swz_off(tid, chunk & 7u)),
                     "r"(tc::pack_bf16_relu(v[0], v[1])), "r"(tc::pack_bf16_relu(v[2], v[3])),
                     "r"(tc::pack_bf16_relu(v[4], v[5])), "r"(tc::pack_bf16_relu(v[6], v[7]))
                     : "memory");
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(pa + tc::swz_off(tid, (chunk + 1) & 7u)),
                     "r"(tc::pack_bf16_relu(v[8], v[9])), "r"(tc::pack_bf16_relu(v[10], v[11])),
                     "r"(tc::pack_bf16_relu(v[12], v[13])), "r"(tc::pack_bf16_relu(v[14], v[15]))
                     : "memory");
      }
    }
    fence_proxy_async_smem();
    tc::fence_before_sync();
    SA_T(7)
    group_sync();
    SA_T(8)
    // ---------------- layer 3 (transposed), one 128-channel half at a time ----------------------
    // D3[h] = W3[h*128 .. +128] . X2^T always lands in the group's first 128 TMEM columns (aliasing D2),
    // so a group needs 128 columns whatever C3 is and up to four groups share an SM.
    for (int h = 0; h < C3 / 128 && !failed; ++h) {
      if (h > 0) {  // every warp of the group has drained half h-1 from TMEM before it is overwritten
        tc::fence_before_sync();
        group_sync();
      }
      if (tid == 0) {
        tc::fence_after_sync();
        for (int kk = 0; kk < C2 / 16; ++kk) {
          const uint32_t po = (kk >> 2), ko = (kk & 3) * 32;
          tc::mma_bf16(tmem_d,
                       tc::smem_desc_sw128(s_w3 + po * (static_cast<uint32_t>(C3) * 128u) + h * (128u * 128u) + ko),
                       tc::smem_desc_sw128(s_x + po * (128u * 128u) + ko), idesc3, kk > 0 ? 1u : 0u);
        }
        tc::mma_commit(&bar_mma);
      }
      if (h == 0) fetch_c();  // next tile: the points' coordinates / raw channels travel under this tile's epilogue
      SA_T(9)
      if (!mbar_wait(&bar_mma, phase)) { failed = true; break; }
      phase ^= 1;
      tc::fence_after_sync();
      SA_T(10)
      // ---- epilogue 3: thread = channel; running max over the columns, flushed where a centre ends ----
      const uint32_t lane_base = tmem_d + (static_cast<uint32_t>(warp * 32) << 16);
      // The rows of a centre are consecutive TMEM columns; the end mask is the same for every thread, so
      // the flush branch never diverges.  A centre that continues in another tile is completed by that
      // tile's flush: outputs are ReLU'd (>= 0), out_cm starts at zero, and an integer max on the bit
      // patterns of non-negative floats is the float max.
      const int ch = h * 128 + tid;
      const float bias = __ldg(a.b3 + ch);
      const int par = it & 1;
      uint32_t endm[4];
#pragma unroll
      for (int w = 0; w < 4; ++w) endm[w] = s_end[grp][par][w];
      const int last_col = endm[3] ? 127 - __clz(endm[3]) : (endm[2] ? 95 - __clz(endm[2])
                         : (endm[1] ? 63 - __clz(endm[1]) : (endm[0] ? 31 - __clz(endm[0]) : -1)));
      // Every centre owns at least one compacted row and the rows of a scene are listed in centre order, so the
      // centres that end in a tile are consecutive: the flush address is a running pointer, no lookup.
      int* out_p = reinterpret_cast<int*>(a.out_cm) + (static_cast<long long>(scene) * C3 + ch) * a.npoint -
                   static_cast<long long>(scene) * a.npoint + s_jg0[grp][par];
      float mx = -INFINITY;
      // NOT unrolled: the body (a fast path and a per-column path for each of four column quads) is ~4 KB of
      // code; eight copies made the tile loop larger than the 32 KB instruction cache the SM's four groups
      // share, and every skipped slow path was an instruction-fetch miss
#pragma unroll 1
      for (int c0 = 0; c0 <= last_col; c0 += 16) {
        float v[16];
#ifdef PN2_SA_TIMING
        const long long _l0 = clock64();
#endif
#ifdef PN2_SA_EXP_NOLD
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = __int_as_float(c0 + i + tid + it);
#else
        tc::tmem_ld16(lane_base + c0, v);
#endif
#ifdef PN2_SA_TIMING
        dbg_ld += clock64() - _l0; ++dbg_nld;
#endif
#ifdef PN2_SA_EXP_NOMAX
        mx = fmaxf(mx, v[0] + v[15]);
        if (c0 == 112) asm volatile("red.global.max.s32 [%0], %1;" ::"l"(out_p), "r"(__float_as_int(fmaxf(mx + bias, 0.f))) : "memory");
        continue;
#endif
        // (selects, not endm[c0 >> 5]: a dynamically indexed array would live in local memory)
        const uint32_t e_lo = (c0 & 64) ? endm[2] : endm[0], e_hi = (c0 & 64) ? endm[3] : endm[1];
        const uint32_t bits = (((c0 & 32) ? e_hi : e_lo) >> (c0 & 31)) & 0xffffu;
        // four columns at a time: where no centre ends inside them (the common case: 28 / 7 / 14 rows per
        // centre in SA1 / SA2 / SA3-4) the running max is two 3-input FMNMX3 and one warp-uniform test
        // instead of four max + four tests; max is exact, so the grouping cannot change a bit
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          if (((bits >> (4 * q4)) & 0xfu) == 0u) {
            mx = fmax3(fmax3(mx, v[4 * q4], v[4 * q4 + 1]), v[4 * q4 + 2], v[4 * q4 + 3]);
          } else {
#pragma unroll
            for (int i = 4 * q4; i < 4 * q4 + 4; ++i) {
              mx = fmaxf(mx, v[i]);
              if (bits & (1u << i)) {
#ifdef PN2_SA_EXP_NORED  // experiment: how much of the epilogue is the throughput of the scattered reductions
                if (__float_as_int(fmaxf(mx + bias, 0.f)) == 0x7fc12345)
#endif
                asm volatile("red.global.max.s32 [%0], %1;" ::"l"(out_p), "r"(__float_as_int(fmaxf(mx + bias, 0.f))) : "memory");
                ++out_p;
                mx = -INFINITY;
              }
            }
          }
        }
      }
    }
    if (failed) break;
    tc::fence_before_sync();  // ordered before the next tile's __syncthreads / MMA into the same TMEM
    SA_T(11)
  }
#ifdef PN2_SA_TIMING
  unsigned long long dbg_g1;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(dbg_g1));
  if (tid == 0 && grp == 0 && (blockIdx.x == 0 || blockIdx.x == gridDim.x / 2 || blockIdx.x == gridDim.x - 1) && dbg_tiles > 0)
    printf("sa_fused[cta %d of %d, C1 %d C2 %d C3 %d MMA1 %d NG %d]: globaltimer entry %llu ns, lifetime %llu ns; %d tiles, %lld cycles/tile; prologue %lld cycles (w1 panel %lld, tile prefix %lld, tmem alloc %lld, sync %lld, first fetch chain %lld); per tile: "
           "loop-top %lld build %lld sync1 %lld mma1 %lld epi1 %lld sync2 %lld mma2 %lld epi2 %lld sync3 %lld "
           "issue3+prefetch %lld mma3-wait %lld epi3 %lld (of which %d tmem loads: %lld)\n",
           blockIdx.x, gridDim.x, C1, C2, C3, int(MMA1), NG, dbg_g0 % 100000000ull, dbg_g1 - dbg_g0, dbg_tiles,
           (clock64() - dbg_start) / dbg_tiles, dbg_start - dbg_entry, dbg_p[0], dbg_p[1], dbg_p[2], dbg_p[3], dbg_p[4],
           dbg_t[0] / dbg_tiles, dbg_t[1] / dbg_tiles, dbg_t[2] / dbg_tiles, dbg_t[3] / dbg_tiles, dbg_t[4] / dbg_tiles,
           dbg_t[5] / dbg_tiles, dbg_t[6] / dbg_tiles, dbg_t[7] / dbg_tiles, dbg_t[8] / dbg_tiles, dbg_t[9] / dbg_tiles,
           dbg_t[10] / dbg_tiles, dbg_t[11] / dbg_tiles, dbg_nld / dbg_tiles, dbg_ld / dbg_tiles);
#endif
  if (failed && tid == 0) atomicExch(a.err, 2);
  // a group that found no tile never waited for the weights: the bulk copies must have landed before
  // this CTA's shared memory can be handed to another CTA
  if (!weights_ready && threadIdx.x == 0) mbar_wait(&bar_w, 0);
  tc::fence_before_sync();
  __syncthreads();
  if (threadIdx.x < 32) tc::tmem_dealloc(tmem_base, tmem_cols);
}


// ---------------------------------------------------------------------------------------------
// Row plan.  ball_query pads every ball to nsample entries by repeating its first hit
// (ball_query_gpu.cu:40-44); a repeated row has the inputs of slot 0, so its activations are a copy of
// slot 0's and the max-pool cannot see it.  On ScanNet-density scenes balls hold 28 of 64 (SA1) and 7
// of 32 (SA2) distinct points on average.  sa_count_kernel counts each centre's distinct hits (the
// first repeat of idx[0] ends them: points are unique); one CTA per scene then takes the exclusive
// prefix sum and writes, per
// compacted row, the index of its entry in idx (-1 for the tail of the last tile); the fused kernel then
// runs ceil(sum / 128) tiles per scene instead of npoint * nsample / 128, with bit-identical output.
// The counting kernel also zero-fills out_cm (the fused kernel combines with an atomic max).
// ---------------------------------------------------------------------------------------------
// distinct hits of every centre (warp per centre, all scenes at once) + zero-fill of out_cm
__global__ void __launch_bounds__(256)
sa_count_kernel(int n_centres, int ns, const int* __restrict__ idx, int* __restrict__ cnt_out,
                float4* __restrict__ out4, long long n_out4, int* __restrict__ tile_counter) {
  const int lane = threadIdx.x & 31;
  if (blockIdx.x == 0 && threadIdx.x == 0) *tile_counter = 0;
  const int j = blockIdx.x * 8 + (threadIdx.x >> 5);
  for (long long i = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x; i < n_out4; i += static_cast<long long>(gridDim.x) * 256)
    out4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (j >= n_centres) return;
  const int* row = idx + static_cast<long long>(j) * ns;
  const int first = __ldg(row);
  int v[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) v[k] = lane + 32 * k < ns ? __ldg(row + lane + 32 * k) : -2;  // ns <= 128
  int cnt = ns;
#pragma unroll
  for (int k = 3; k >= 0; --k) {
    const int sl = lane + 32 * k;
    const uint32_t m = __ballot_sync(0xffffffffu, sl > 0 && sl < ns && v[k] == first);
    if (m) cnt = 32 * k + __ffs(m) - 1;  // descending k: the lowest repeat wins
  }
  if (lane == 0) cnt_out[j] = cnt;
}

// per scene: exclusive scan of the counts, then one warp per centre writes its compacted rows
__global__ void __launch_bounds__(1024)
sa_plan_kernel(int npoint, int ns_shift, int cap, const int* __restrict__ cnt_in, int* __restrict__ row_src,
               int* __restrict__ tiles_b) {
  extern __shared__ int s_plan[];  // off[npoint + 1], then 33 words of scan scratch
  int* s_off = s_plan;
  int* s_scan = s_plan + npoint + 1;
  const int b = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int* cntb = cnt_in + static_cast<long long>(b) * npoint;
  const int per = (npoint + blockDim.x - 1) / blockDim.x;  // <= 4 (npoint <= 4096)
  const int j0 = threadIdx.x * per;
  int mine[4], local = 0;
#pragma unroll
  for (int q = 0; q < 4; ++q) { mine[q] = (q < per && j0 + q < npoint) ? __ldg(cntb + j0 + q) : 0; local += mine[q]; }
  int incl = local;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int o = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += o;
  }
  if (lane == 31) s_scan[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    int w = lane < nwarps ? s_scan[lane] : 0;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int o = __shfl_up_sync(0xffffffffu, w, d);
      if (lane >= d) w += o;
    }
    s_scan[lane] = w;  // inclusive over warps
    if (lane == 31) s_scan[32] = w;
  }
  __syncthreads();
  int run = incl - local + (warp > 0 ? s_scan[warp - 1] : 0);
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    if (q < per && j0 + q < npoint) { s_off[j0 + q] = run; run += mine[q]; }
  }
  const int total = s_scan[32];
  if (threadIdx.x == 0) { s_off[npoint] = total; tiles_b[b] = (total + 127) >> 7; }
  __syncthreads();
  int* dst = row_src + static_cast<long long>(b) * cap;
  for (int j = warp; j < npoint; j += nwarps) {
    const int off = s_off[j], c = s_off[j + 1] - off;
    const int base = (b * npoint + j) << ns_shift;
    for (int sl = lane; sl < c; sl += 32) dst[off + sl] = base + sl;
  }
  const int padded = (total + 127) & ~127;
  for (int i = total + threadIdx.x; i < padded; i += blockDim.x) dst[i] = -1;
}

template <int CPL, int NG>
int launch_cd(int cd, const SaArgs& a, int grid, size_t smem, cudaStream_t s) {
#define PN2_SA_LAUNCH(CDV, M1)                                                                             \
  {                                                                                                        \
    PN2_CUDA(cudaFuncSetAttribute(sa_fused_kernel<CPL, CDV, NG, M1>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                  static_cast<int>(smem)));                                                \
    sa_fused_kernel<CPL, CDV, NG, M1><<<grid, SA_THREADS * NG, smem, s>>>(a);                              \
  }
  if (a.g != nullptr) {  // per-point layer-1 rows are gathered: FMA build (cd == 0, checked by the caller)
    PN2_SA_LAUNCH(0, false)
  } else {               // raw channels: layer 1 on the tensor cores
    switch (cd) {
      case 0: PN2_SA_LAUNCH(0, true) break;
      case 1: PN2_SA_LAUNCH(1, true) break;
      case 2: PN2_SA_LAUNCH(2, true) break;
      case 3: PN2_SA_LAUNCH(3, true) break;
      case 4: PN2_SA_LAUNCH(4, true) break;
      default: set_error("sa_group_mlp: direct feature channels must be <= 4 (got %d)", cd); return PN2_ERR_INVALID;
    }
  }
#undef PN2_SA_LAUNCH
  PN2_CHECK_LAUNCH("sa_fused_kernel");
  return PN2_OK;
}

}  // namespace

int* device_err_flag();

}  // namespace pn2

using namespace pn2;

extern "C" int pn2_sa_group_mlp(int b, int n, int npoint, int nsample, float radius_div, const float* xyz,
                                const float* new_xyz, const int* idx, int c_direct, const float* feat,
                                const float* g, int c1, int c2, int c3, const float* w1, const float* b1,
                                const void* w2_packed, const float* b2, const void* w3_packed, const float* b3,
                                float* out_cm, void* out_pm_bf16, void* stream) {
  PN2_RANGE("pn2_sa_group_mlp");
  if (b <= 0 || npoint <= 0) return PN2_OK;
  PN2_REQUIRE(nsample == 16 || nsample == 32 || nsample == 64 || nsample == 128,
              "sa_group_mlp: nsample=%d must be 16, 32, 64 or 128", nsample);
  PN2_REQUIRE((c1 == 64 || c1 == 128) && (c2 == 64 || c2 == 128) && (c3 == 128 || c3 == 256),
              "sa_group_mlp: unsupported widths %d/%d/%d (c1,c2 in {64,128}, c3 in {128,256})", c1, c2, c3);
  PN2_REQUIRE(xyz && new_xyz && idx && w1 && b1 && w2_packed && b2 && w3_packed && b3 && out_cm && n > 0,
              "sa_group_mlp: null pointer");
  PN2_REQUIRE(c_direct >= 0 && c_direct <= 4 && (c_direct == 0 || feat), "sa_group_mlp: bad direct features");
  PN2_REQUIRE(!(g && c_direct > 0), "sa_group_mlp: pass either g (per-point layer-1 rows) or raw channels, not both");
  PN2_REQUIRE(radius_div != 0.f, "sa_group_mlp: radius_div must be non-zero");
  SaArgs a;
  a.xyz = xyz; a.new_xyz = new_xyz; a.idx = idx; a.g = g; a.feat = feat; a.w1 = w1; a.b1 = b1;
  a.out_cm = out_cm;
  a.w2p = static_cast<const uint8_t*>(w2_packed); a.b2 = b2; a.w3p = static_cast<const uint8_t*>(w3_packed); a.b3 = b3;
  a.B = b; a.N = n; a.npoint = npoint; a.ns = nsample; a.C1 = c1; a.C2 = c2; a.C3 = c3;
  a.radius_div = radius_div;
  a.ns_shift = nsample == 16 ? 4 : (nsample == 32 ? 5 : (nsample == 64 ? 6 : 7));
  PN2_REQUIRE(static_cast<long long>(b) * npoint < (1LL << 31), "sa_group_mlp: b * npoint must stay below 2^31");
  a.total_rows = static_cast<long long>(b) * npoint * nsample;
  a.err = device_err_flag();
  PN2_REQUIRE(a.err != nullptr, "sa_group_mlp: cannot allocate the device error flag");
  // ---- row plan (stream-ordered workspace: the call stays asynchronous and capturable) -----------------
  PN2_REQUIRE(b <= kMaxScenes, "sa_group_mlp: at most %d scenes per call (got %d)", kMaxScenes, b);
  PN2_REQUIRE(npoint <= 4096, "sa_group_mlp: npoint=%d exceeds the row plan's 4096", npoint);
  PN2_REQUIRE(static_cast<long long>(b) * npoint * nsample < (1LL << 31), "sa_group_mlp: b * npoint * nsample must stay below 2^31");
  const int cap = (npoint * nsample + 127) & ~127;
  const size_t plan_smem = (static_cast<size_t>(npoint) + 40) * sizeof(int);
  keep_workspace_pool();
  int* ws = nullptr;
  const size_t n_centres = static_cast<size_t>(b) * npoint;
  PN2_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&ws), (static_cast<size_t>(b) * cap + b + n_centres + 2) * sizeof(int),
                           as_stream(stream)));
  int* tiles_b = ws + static_cast<size_t>(b) * cap;
  int* cnt_ws = tiles_b + b;
  // the dynamic tile counter lives in the call's own workspace and is zeroed by sa_count_kernel: no
  // process-wide ring, nothing to recycle when a captured graph is destroyed
  a.tile_counter = cnt_ws + n_centres;
  a.row_src = ws; a.tiles_b = tiles_b; a.cap = cap;
  {
    const long long n_out4 = (static_cast<long long>(b) * c3 * npoint) >> 2;  // c3 is a multiple of 128
    sa_count_kernel<<<static_cast<unsigned>((n_centres + 7) / 8), 256, 0, as_stream(stream)>>>(
        static_cast<int>(n_centres), nsample, idx, cnt_ws, reinterpret_cast<float4*>(out_cm), n_out4, a.tile_counter);
    sa_plan_kernel<<<b, 1024, plan_smem, as_stream(stream)>>>(npoint, a.ns_shift, cap, cnt_ws, ws, tiles_b);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { cudaFreeAsync(ws, as_stream(stream)); return cuda_fail(e, "sa_plan_kernel"); }
  }
  const size_t w_bytes = static_cast<size_t>(c2) * c1 * 2 + static_cast<size_t>(c3) * c2 * 2 +
                         (g == nullptr ? static_cast<size_t>(c1) * 128 : 0);  // + the layer-1 operand panel
  const size_t x_bytes = 128 * static_cast<size_t>(c1 > c2 ? c1 : c2) * 2;
  const long long tiles = (a.total_rows + 127) / 128;
  // Groups per CTA: as many independent 128-thread groups (own tile, own 128 TMEM columns) as shared
  // memory allows, up to 4 (4 x 128 = all 512 TMEM columns), in ONE CTA per SM: the per-CTA prologue
  // (layer-1 operand panel, weight TMA, TMEM allocation: ~13 k cycles, more than a tile) is then paid once
  // per SM instead of once per group (SA1: +5 % on the whole step against four single-group CTAs per SM).
  // Only calls too small to give every SM four tiles keep single-group CTAs, spread over more SMs.
  const size_t limit = 227 * 1024 - 2048;  // leave room for the static shared memory
  const size_t smem1 = w_bytes + x_bytes + 1024;
  int ng = 1, resident = 1;
  if (smem1 <= 56 * 1024 && tiles < 4LL * sm_count()) {
    resident = 4;  // four single-group CTAs per SM (SA1-sized weights, small call)
  } else {
    ng = 4;
    while (ng > 1 && w_bytes + ng * x_bytes + 1024 > limit) ng >>= 1;
    if (ng == 1 && smem1 <= 113 * 1024) resident = 2;
  }
  const size_t smem = w_bytes + ng * x_bytes + 1024;
  long long grid = static_cast<long long>(sm_count()) * resident;
  if (grid * ng > tiles) grid = (tiles + ng - 1) / ng;
  // pn2_set_option("sa_tiles_per_group", k): fewer persistent CTAs, each group claiming >= k tiles (counted
  // before padded rows are dropped), so the per-CTA prologue (weight TMA, TMEM allocation) is paid less
  // often -- for several batches in flight, where SM-time and not the latency of one launch counts
  if (const int k = option_value(OPT_SA_TILES_PER_GROUP)) {
    const long long capped = (tiles + static_cast<long long>(ng) * k - 1) / (static_cast<long long>(ng) * k);
    if (capped < grid) grid = capped < 1 ? 1 : capped;
  }
  cudaStream_t s = as_stream(stream);
  int rc;
  if (ng == 4) {
    rc = c1 == 128 ? launch_cd<4, 4>(c_direct, a, static_cast<int>(grid), smem, s)
                   : launch_cd<2, 4>(c_direct, a, static_cast<int>(grid), smem, s);
  } else if (ng == 2) {
    rc = c1 == 128 ? launch_cd<4, 2>(c_direct, a, static_cast<int>(grid), smem, s)
                   : launch_cd<2, 2>(c_direct, a, static_cast<int>(grid), smem, s);
  } else {
    rc = c1 == 128 ? launch_cd<4, 1>(c_direct, a, static_cast<int>(grid), smem, s)
                   : launch_cd<2, 1>(c_direct, a, static_cast<int>(grid), smem, s);
  }
  cudaFreeAsync(ws, s);
  // the bf16 point-major twin (next layer's GEMM operand) is a transpose of the finished out_cm
  if (rc == PN2_OK && out_pm_bf16) rc = pn2_pack_rows_bf16(b, c3, npoint, c3, out_cm, out_pm_bf16, stream);
  return rc;
}
