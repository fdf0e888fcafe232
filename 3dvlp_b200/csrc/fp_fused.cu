// Feature-propagation front end: three_nn + inverse-distance weights + three_interpolate + skip
// concatenation, fused into ONE kernel that writes the bf16 point-major operand of the FP shared MLP.
//
// Reference path (pointnet2_modules.py:393-412): three_nn kernel -> sqrt -> 1/(d+1e-8) -> sum -> div
// (4 torch kernels) -> three_interpolate kernel -> torch.cat -> unsqueeze -> cuDNN 1x1 convs.  Here a
// CTA owns 32 unknown points and 128 channels of the output: its four warps find the three nearest
// known points (known xyz staged in shared memory; the strict-`<` cascade of interpolate_gpu.cu:39-55
// per quarter, merged by (distance, index) so ties keep the earlier index) and the weights -- the
// search is cheap enough to repeat per channel block, which buys 4x the CTAs for a launch that is
// otherwise latency-bound -- then interpolate 64 channels at a time (lane = point, so the skip
// features are read coalesced from the reference's (B, C, n) layout), transpose through shared
// memory and store 128-byte row segments of  X[b*n + j][0 : C2 + C1]  in bf16 -- the K-major
// activation matrix that pn2_linear_bf16 (tcgen05) consumes.  Interpolated channels come first,
// skip channels second (torch.cat([interpolated, unknow_feats], dim=1)).
#include "tc.cuh"

namespace pn2 {
namespace {

constexpr int FP_THREADS = 128;
constexpr int FP_KTILE = 1024;  // known points staged per pass (16 KB as float4)
constexpr int FP_CH_PER_CTA = 128;  // channels of X one CTA writes (grid.z splits the K dimension)

// (d, i) ordered lexicographically: the order in which the reference's strict-`<` cascade over
// ascending indices ranks candidates (an equal distance never displaces an earlier index).
__device__ __forceinline__ bool before(float d, int i, float bd, int bi) { return d < bd || (d == bd && i < bi); }

__device__ __forceinline__ void insert3(float d, int i, float& b1, float& b2, float& b3, int& i1, int& i2, int& i3) {
  if (before(d, i, b1, i1)) { b3 = b2; i3 = i2; b2 = b1; i2 = i1; b1 = d; i1 = i; }
  else if (before(d, i, b2, i2)) { b3 = b2; i3 = i2; b2 = d; i2 = i; }
  else if (before(d, i, b3, i3)) { b3 = d; i3 = i; }
}

__global__ void __launch_bounds__(FP_THREADS)
fp_prepare_kernel(int n, int m, int c1, int c2, int kpad, const float* __restrict__ unknown,
                  const float* __restrict__ known, const float* __restrict__ uf, const float* __restrict__ kf,
                  const int* __restrict__ pre_idx, const float* __restrict__ pre_w, __nv_bfloat16* __restrict__ out) {
  __shared__ float4 sk[FP_KTILE];
  __shared__ float s_b[4][32][3];
  __shared__ int s_i[4][32][3];
  __shared__ int s_idx[32][3];
  __shared__ float s_w[32][3];
  __shared__ float tile[32][65];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int scene = blockIdx.y, p0 = blockIdx.x * 32;
  unknown += static_cast<size_t>(scene) * n * 3;
  known += static_cast<size_t>(scene) * m * 3;
  const float* kfb = kf + static_cast<size_t>(scene) * c2 * m;
  const float* ufb = uf ? uf + static_cast<size_t>(scene) * c1 * n : nullptr;

  const int j = p0 + lane;
  int a1, a2, a3;
  float w1, w2, w3;
  if (pre_idx != nullptr) {
    // neighbours and weights were computed ahead of time (pn2_three_nn_weights: they depend on the
    // coordinates only, so the backbone evaluates them on its sampling stream, off the critical path)
    const size_t o = (static_cast<size_t>(scene) * n + (j < n ? j : 0)) * 3;
    a1 = __ldg(pre_idx + o); a2 = __ldg(pre_idx + o + 1); a3 = __ldg(pre_idx + o + 2);
    w1 = __ldg(pre_w + o); w2 = __ldg(pre_w + o + 1); w3 = __ldg(pre_w + o + 2);
  } else {
    // ---- three nearest known points + weights for the CTA's 32 unknown points --------------------
    // lane = unknown point; the four warps each scan a quarter of every staged tile of known points
    // and warp 0 merges the four sorted triples by (distance, index).
    float ux = 0.f, uy = 0.f, uz = 0.f;
    if (j < n) { ux = __ldg(unknown + 3 * j); uy = __ldg(unknown + 3 * j + 1); uz = __ldg(unknown + 3 * j + 2); }
    float b1 = INFINITY, b2 = INFINITY, b3 = INFINITY;
    int i1 = 0, i2 = 0, i3 = 0;
    for (int t0 = 0; t0 < m; t0 += FP_KTILE) {
      const int tn = min(FP_KTILE, m - t0);
      __syncthreads();
      for (int q = tid; q < tn; q += FP_THREADS)
        sk[q] = make_float4(__ldg(known + 3 * (t0 + q)), __ldg(known + 3 * (t0 + q) + 1), __ldg(known + 3 * (t0 + q) + 2), 0.f);
      __syncthreads();
      const int quarter = (tn + 3) >> 2;
      const int k_end = min(tn, (warp + 1) * quarter);
      for (int k = warp * quarter; k < k_end; ++k) {
        const float4 p = sk[k];
        const float d = sqdist3(ux - p.x, uy - p.y, uz - p.z);
        if (d < b3) {  // b1 <= b2 <= b3 always: one compare rejects almost every candidate
          if (d < b1) { b3 = b2; i3 = i2; b2 = b1; i2 = i1; b1 = d; i1 = t0 + k; }
          else if (d < b2) { b3 = b2; i3 = i2; b2 = d; i2 = t0 + k; }
          else { b3 = d; i3 = t0 + k; }
        }
      }
    }
    s_b[warp][lane][0] = b1; s_b[warp][lane][1] = b2; s_b[warp][lane][2] = b3;
    s_i[warp][lane][0] = i1; s_i[warp][lane][1] = i2; s_i[warp][lane][2] = i3;
    __syncthreads();
    if (warp == 0) {
  #pragma unroll
      for (int w = 1; w < 4; ++w)
  #pragma unroll
        for (int t = 0; t < 3; ++t) insert3(s_b[w][lane][t], s_i[w][lane][t], b1, b2, b3, i1, i2, i3);
      // pointnet2_utils.py:138-140 (sqrt) and pointnet2_modules.py:395-397 (1/(d+1e-8), normalise)
      const float r1 = __fdiv_rn(1.0f, __fsqrt_rn(b1) + 1e-8f), r2 = __fdiv_rn(1.0f, __fsqrt_rn(b2) + 1e-8f),
                  r3 = __fdiv_rn(1.0f, __fsqrt_rn(b3) + 1e-8f);
      const float norm = (r1 + r2) + r3;
      s_idx[lane][0] = i1; s_idx[lane][1] = i2; s_idx[lane][2] = i3;
      s_w[lane][0] = __fdiv_rn(r1, norm); s_w[lane][1] = __fdiv_rn(r2, norm); s_w[lane][2] = __fdiv_rn(r3, norm);
    }
    __syncthreads();
    a1 = s_idx[lane][0]; a2 = s_idx[lane][1]; a3 = s_idx[lane][2];
    w1 = s_w[lane][0]; w2 = s_w[lane][1]; w3 = s_w[lane][2];
  }
  const bool live = j < n;

  // ---- 64 channels per pass: lane = point, warp = 16-channel slice; transpose; store ----------
  const int ch_begin = blockIdx.z * FP_CH_PER_CTA, ch_stop = min(kpad, ch_begin + FP_CH_PER_CTA);
  for (int ch0 = ch_begin; ch0 < ch_stop; ch0 += 64) {
    // every gather of the 16 channels is requested before the first is used (predicated loads, no branches): one
    // or two global round trips per pass instead of sixteen -- the kernel was 78 % long-scoreboard stalls
    float x1[16], x2[16], x3[16], xu[16];
#pragma unroll
    for (int cc = 0; cc < 16; ++cc) {
      const int c = ch0 + warp * 16 + cc;
      const bool isk = live && c < c2, isu = live && c >= c2 && c < c2 + c1;
      const float* p = kfb + static_cast<size_t>(isk ? c : 0) * m;
      x1[cc] = isk ? __ldg(p + a1) : 0.f;
      x2[cc] = isk ? __ldg(p + a2) : 0.f;
      x3[cc] = isk ? __ldg(p + a3) : 0.f;
      xu[cc] = isu ? __ldg(ufb + static_cast<size_t>(c - c2) * n + j) : 0.f;
    }
#pragma unroll
    for (int cc = 0; cc < 16; ++cc) {
      const int c = ch0 + warp * 16 + cc;
      // zero weights on dead lanes / channels would still be exact, but keep the selects explicit
      const float v = (live && c < c2) ? __fmaf_rn(x3[cc], w3, __fmaf_rn(x1[cc], w1, __fmul_rn(x2[cc], w2))) : xu[cc];
      tile[lane][warp * 16 + cc] = v;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int q = tid + FP_THREADS * i;
      const int pt = q >> 3, piece = q & 7;
      if (p0 + pt < n) {
        const float* t = &tile[pt][piece * 8];
        const uint4 pk = make_uint4(tc::pack_bf16(t[0], t[1]), tc::pack_bf16(t[2], t[3]), tc::pack_bf16(t[4], t[5]),
                                    tc::pack_bf16(t[6], t[7]));
        *reinterpret_cast<uint4*>(out + (static_cast<size_t>(scene) * n + p0 + pt) * kpad + ch0 + piece * 8) = pk;
      }
    }
    __syncthreads();
  }
}


// Three nearest known points of every unknown point and their normalised inverse-distance weights,
// one WARP per unknown point: each lane keeps the best three of its stride of the known set with the
// reference's strict-`<` cascade (interpolate_gpu.cu:39-55), then three rounds of a warp-wide
// lexicographic (distance, index) minimum pop the global best three -- the order the serial cascade
// over ascending indices produces.  Weights as pointnet2_utils.py:138-140 + pointnet2_modules.py:395-397.
__global__ void __launch_bounds__(256)
three_nn_weights_kernel(int n, int m, const float* __restrict__ unknown, const float* __restrict__ known,
                        int* __restrict__ idx, float* __restrict__ weight) {
  const int lane = threadIdx.x & 31, scene = blockIdx.y;
  const int j = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (j >= n) return;
  const float* u = unknown + (static_cast<size_t>(scene) * n + j) * 3;
  const float* kn = known + static_cast<size_t>(scene) * m * 3;
  const float ux = __ldg(u), uy = __ldg(u + 1), uz = __ldg(u + 2);
  float b1 = INFINITY, b2 = INFINITY, b3 = INFINITY;
  int i1 = 0, i2 = 0, i3 = 0;
  for (int k = lane; k < m; k += 32) {
    const float d = sqdist3(ux - __ldg(kn + 3 * k), uy - __ldg(kn + 3 * k + 1), uz - __ldg(kn + 3 * k + 2));
    if (d < b3) {
      if (d < b1) { b3 = b2; i3 = i2; b2 = b1; i2 = i1; b1 = d; i1 = k; }
      else if (d < b2) { b3 = b2; i3 = i2; b2 = d; i2 = k; }
      else { b3 = d; i3 = k; }
    }
  }
  float rd[3];
  int ri[3];
#pragma unroll
  for (int t = 0; t < 3; ++t) {
    // distances are >= 0 (or +inf for an empty slot): their bit patterns order like the floats
    const uint32_t dbits = __float_as_uint(b1);
    const uint32_t dmin = __reduce_min_sync(0xffffffffu, dbits);
    const uint32_t cand = dbits == dmin ? static_cast<uint32_t>(i1) : 0xffffffffu;
    const uint32_t imin = __reduce_min_sync(0xffffffffu, cand);
    rd[t] = __uint_as_float(dmin);
    ri[t] = static_cast<int>(imin);
    if (dbits == dmin && static_cast<uint32_t>(i1) == imin && dmin != 0x7f800000u) {  // pop the winner's head
      b1 = b2; i1 = i2; b2 = b3; i2 = i3; b3 = INFINITY; i3 = 0;
    }
  }
  if (lane == 0) {
    const float r1 = __fdiv_rn(1.0f, __fsqrt_rn(rd[0]) + 1e-8f), r2 = __fdiv_rn(1.0f, __fsqrt_rn(rd[1]) + 1e-8f),
                r3 = __fdiv_rn(1.0f, __fsqrt_rn(rd[2]) + 1e-8f);
    const float norm = (r1 + r2) + r3;
    const size_t o = (static_cast<size_t>(scene) * n + j) * 3;
    idx[o] = ri[0]; idx[o + 1] = ri[1]; idx[o + 2] = ri[2];
    weight[o] = __fdiv_rn(r1, norm); weight[o + 1] = __fdiv_rn(r2, norm); weight[o + 2] = __fdiv_rn(r3, norm);
  }
}

}  // namespace
}  // namespace pn2

using namespace pn2;

extern "C" int pn2_fp_prepare(int b, int n, int m, int c1, int c2, int kpad, const float* unknown,
                              const float* known, const float* unknow_feats, const float* known_feats,
                              void* out_bf16, void* stream) {
  PN2_RANGE("pn2_fp_prepare");
  if (b <= 0 || n <= 0) return PN2_OK;
  PN2_REQUIRE(m > 0 && c2 > 0 && c1 >= 0 && kpad >= c1 + c2 && kpad % 64 == 0, "fp_prepare: bad sizes");
  PN2_REQUIRE(unknown && known && known_feats && out_bf16 && (c1 == 0 || unknow_feats) && b <= 65535,
              "fp_prepare: bad arguments");
  dim3 grid((n + 31) / 32, b, (kpad + FP_CH_PER_CTA - 1) / FP_CH_PER_CTA);
  fp_prepare_kernel<<<grid, FP_THREADS, 0, as_stream(stream)>>>(n, m, c1, c2, kpad, unknown, known, unknow_feats,
                                                               known_feats, nullptr, nullptr,
                                                               static_cast<__nv_bfloat16*>(out_bf16));
  PN2_CHECK_LAUNCH("fp_prepare_kernel");
  return PN2_OK;
}

extern "C" int pn2_three_nn_weights(int b, int n, int m, const float* unknown, const float* known, int* idx,
                                    float* weight, void* stream) {
  PN2_RANGE("pn2_three_nn_weights");
  if (b <= 0 || n <= 0) return PN2_OK;
  PN2_REQUIRE(m > 0 && unknown && known && idx && weight && b <= 65535, "three_nn_weights: bad arguments");
  dim3 grid((n + 7) / 8, b);
  three_nn_weights_kernel<<<grid, 256, 0, as_stream(stream)>>>(n, m, unknown, known, idx, weight);
  PN2_CHECK_LAUNCH("three_nn_weights_kernel");
  return PN2_OK;
}

extern "C" int pn2_fp_prepare_indexed(int b, int n, int m, int c1, int c2, int kpad, const int* idx,
                                      const float* weight, const float* unknow_feats, const float* known_feats,
                                      void* out_bf16, void* stream) {
  PN2_RANGE("pn2_fp_prepare_indexed");
  if (b <= 0 || n <= 0) return PN2_OK;
  PN2_REQUIRE(m > 0 && c2 > 0 && c1 >= 0 && kpad >= c1 + c2 && kpad % 64 == 0, "fp_prepare_indexed: bad sizes");
  PN2_REQUIRE(idx && weight && known_feats && out_bf16 && (c1 == 0 || unknow_feats) && b <= 65535,
              "fp_prepare_indexed: bad arguments");
  dim3 grid((n + 31) / 32, b, (kpad + FP_CH_PER_CTA - 1) / FP_CH_PER_CTA);
  fp_prepare_kernel<<<grid, FP_THREADS, 0, as_stream(stream)>>>(n, m, c1, c2, kpad, nullptr, nullptr, unknow_feats,
                                                               known_feats, idx, weight,
                                                               static_cast<__nv_bfloat16*>(out_bf16));
  PN2_CHECK_LAUNCH("fp_prepare_kernel");
  return PN2_OK;
}
