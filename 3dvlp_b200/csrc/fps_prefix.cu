// "Is the input already in furthest-point order?" -- a sync-free shortcut for the FPS calls of SA2..SA4.
//
// The backbone runs FPS on the OUTPUT of the previous layer's FPS (pointnet2_modules.py:233-240 applied
// layer after layer, backbone_module.py:101-113).  A furthest-point ordering is hereditary: the point
// that is furthest from {p_0..p_{k-1}} in the full cloud is still the furthest among the sampled subset,
// so FPS over (p_0, p_1, ...) returns 0, 1, 2, ... -- the reference itself relies on this
// (`fp2_inds = sa1_inds[:, :num_seed]`, backbone_module.py:134).  Ties and the reference's
// block-size-dependent tie-break can break the rule, so it is VERIFIED, not assumed:
//
//   sample k is the reference's choice  <=>  for every eligible point p != k
//        (t_p, -key(p)) < (r_k, -key(k))      with  t_p = min_{j<k} d(p, x_j),  r_k = t_k
//
// (the packed arg-max of fps.cu: larger running minimum wins, smaller key32 breaks ties; points with
// |x|^2 <= 1e-3 never take part, sampling_gpu.cu:105-106).  r_k only needs the first k points
// (kernel A: thread = k), and then every point checks its own inequality for all k with NO step-wise
// synchronisation (kernel B: thread = p).  That is the same n*m distance evaluations as FPS itself but
// embarrassingly parallel: ~20 us instead of 220 us for 2048 -> 1024.  The per-scene verdict stays on
// the device: the regular FPS kernel reads it and either writes 0..m-1 or runs as usual.
#include "common.cuh"

namespace pn2 {
namespace {

__device__ __forceinline__ uint32_t key32(uint32_t p, int log2bs) {
  return log2bs ? (__brev(p & ((1u << log2bs) - 1u)) | (p >> log2bs)) : p;
}
__device__ __forceinline__ bool eligible(float x, float y, float z) {
  const float mag = __fmaf_rn(z, z, __fmaf_rn(x, x, __fmul_rn(y, y)));
  return !(static_cast<double>(mag) <= 1e-3);
}

// r[k] = min_{j<k} d(x_k, x_j) for k in [1, m); verdict[scene] starts as "in order".
// The serial part of either kernel is one fminf per step; the distance evaluations of eight steps are
// independent, so the loops are unrolled by eight to keep a lone warp per scheduler busy.
__global__ void __launch_bounds__(128)
prefix_radius_kernel(int n, int m, const float* __restrict__ xyz, float* __restrict__ r, int* __restrict__ verdict) {
  extern __shared__ float4 s_pt[];  // the points every thread of this CTA may need: j < (blockIdx.x + 1) * 128
  const int scene = blockIdx.y, k = blockIdx.x * 128 + threadIdx.x;
  const float* x = xyz + static_cast<size_t>(scene) * n * 3;
  const int need = min(m, (blockIdx.x + 1) * 128);
  for (int j = threadIdx.x; j < need; j += 128)
    s_pt[j] = make_float4(__ldg(x + 3 * j), __ldg(x + 3 * j + 1), __ldg(x + 3 * j + 2), 0.f);
  if (k == 0) verdict[scene] = 1;
  __syncthreads();
  if (k >= m) return;
  const float4 me = s_pt[k];
  float t = 1e10f;  // sampling.cpp:79: temp filled with 1e10
  int j = 0;
  for (; j + 8 <= k; j += 8) {
    float d[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const float4 q = s_pt[j + u];
      d[u] = sqdist3(me.x - q.x, me.y - q.y, me.z - q.z);
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) t = fminf(d[u], t);
  }
  for (; j < k; ++j) {
    const float4 q = s_pt[j];
    t = fminf(sqdist3(me.x - q.x, me.y - q.y, me.z - q.z), t);
  }
  r[static_cast<size_t>(scene) * m + k] = t;
}

// Every point checks its inequality for all k.  (value, key) pairs are compared as one 64-bit integer:
// running minima are >= 0, so their bit patterns order like the floats, and the complement of the key
// makes "smaller key wins" a plain `>`; for p == k both sides are equal, so no special case.
__global__ void __launch_bounds__(128)
prefix_check_kernel(int n, int m, int log2bs, const float* __restrict__ xyz, const float* __restrict__ r,
                    int* __restrict__ verdict) {
  extern __shared__ float4 s_pt[];  // [m] points (x, y, z, -), then [m] packed winners as uint2 {~key_k, bits(r_k)}
  uint2* s_win = reinterpret_cast<uint2*>(s_pt + m);
  const int scene = blockIdx.y, p = blockIdx.x * 128 + threadIdx.x;
  const float* x = xyz + static_cast<size_t>(scene) * n * 3;
  for (int k = threadIdx.x; k < m; k += 128) {
    s_pt[k] = make_float4(__ldg(x + 3 * k), __ldg(x + 3 * k + 1), __ldg(x + 3 * k + 2), 0.f);
    s_win[k] = make_uint2(~key32(static_cast<uint32_t>(k), log2bs),
                          k > 0 ? __float_as_uint(__ldg(r + static_cast<size_t>(scene) * m + k)) : 0u);
  }
  __syncthreads();
  if (p >= n) return;
  const float px = __ldg(x + 3 * p), py = __ldg(x + 3 * p + 1), pz = __ldg(x + 3 * p + 2);
  bool bad = !(isfinite(px) && isfinite(py) && isfinite(pz));  // non-finite input: let the real kernel decide
  const bool elig = eligible(px, py, pz);
  if (p >= 1 && p < m && !elig) bad = true;  // an ineligible point is never the reference's pick
  const uint32_t nkp = ~key32(static_cast<uint32_t>(p), log2bs);
  float t = 1e10f;
  if (elig) {
    int k = 1;
    for (; k + 8 <= m; k += 8) {
      float d[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const float4 q = s_pt[k + u - 1];  // sample k+u-1 is the newest centre when sample k+u is chosen
        d[u] = sqdist3(px - q.x, py - q.y, pz - q.z);
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        t = fminf(d[u], t);
        const uint2 w = s_win[k + u];
        const unsigned long long mine = (static_cast<unsigned long long>(__float_as_uint(t)) << 32) | nkp;
        const unsigned long long win = (static_cast<unsigned long long>(w.y) << 32) | w.x;
        bad |= mine > win;
      }
    }
    for (; k < m; ++k) {
      const float4 q = s_pt[k - 1];
      t = fminf(sqdist3(px - q.x, py - q.y, pz - q.z), t);
      const uint2 w = s_win[k];
      bad |= ((static_cast<unsigned long long>(__float_as_uint(t)) << 32) | nkp) >
             ((static_cast<unsigned long long>(w.y) << 32) | w.x);
    }
  }
  if (bad) verdict[scene] = 0;
}

}  // namespace

// ws: m*b floats (r) followed by b ints (verdict); returns the verdict pointer through *verdict
int fps_prefix_check(int b, int n, int m, int log2bs, const float* dataset, void* ws, const int** verdict,
                     cudaStream_t stream) {
  float* r = static_cast<float*>(ws);
  int* v = reinterpret_cast<int*>(r + static_cast<size_t>(b) * m);
  const size_t smem_a = static_cast<size_t>(m) * sizeof(float4);
  const size_t smem_b = static_cast<size_t>(m) * (sizeof(float4) + sizeof(uint2));
  if (smem_a > 48 * 1024)
    PN2_CUDA(cudaFuncSetAttribute(prefix_radius_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem_a)));
  if (smem_b > 48 * 1024)
    PN2_CUDA(cudaFuncSetAttribute(prefix_check_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem_b)));
  prefix_radius_kernel<<<dim3((m + 127) / 128, b), 128, smem_a, stream>>>(n, m, dataset, r, v);
  prefix_check_kernel<<<dim3((n + 127) / 128, b), 128, smem_b, stream>>>(n, m, log2bs, dataset, r, v);
  PN2_CHECK_LAUNCH("fps_prefix_check");
  *verdict = v;
  return PN2_OK;
}

}  // namespace pn2
