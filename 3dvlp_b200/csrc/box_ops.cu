// Axis-aligned 3D box IoU / DIoU, batched on the GPU.
//
// Reference: utils/box_util.py:488-528 `box3d_diou_batch_tensor` -- ~45 elementwise torch kernels per
// call, called B x lang_num_max times per step from Python loops with host syncs
// (lib/loss_helper/loss_grounding.py:205-231), and pytorch3d's `box3d_overlap` on boxes that are
// axis-aligned by construction (models/constrast_module/constrast_module.py:9-16,105).
// Here: ONE launch for all (scene, sentence, proposal) triples, and one for the gradient.
//
// The forward restates the reference's fp32 expression tree with explicit round-to-nearest
// intrinsics in the reference's operator order (no FMA contraction), so IoU values -- and every label
// thresholded from them (>= 0.25, arg-max) -- are bit-identical to the torch ops.
#include "common.cuh"

namespace pn2 {
namespace {

// x < lo ? lo : x written so that a NaN passes through, like torch.clamp
__device__ __forceinline__ float clamp_min0(float x) { return x < 0.f ? 0.f : x; }
__device__ __forceinline__ float vmax(float a, float b) { return a > b ? a : b; }
__device__ __forceinline__ float vmin(float a, float b) { return a < b ? a : b; }

struct Axis {
  float lo1, hi1, lo2, hi2;
};

__device__ __forceinline__ Axis axis_of(float c1, float s1, float c2, float s2) {
  const float h1 = __fmul_rn(s1, 0.5f), h2 = __fmul_rn(s2, 0.5f);  // s/2 is exact either way
  return {__fsub_rn(c1, h1), __fadd_rn(c1, h1), __fsub_rn(c2, h2), __fadd_rn(c2, h2)};
}

// box_util.py:505-527
__device__ __forceinline__ void diou_forward(const float c1[3], const float s1[3], const float c2[3],
                                             const float s2[3], float& iou, float& diou) {
  const float vol1 = __fmul_rn(__fmul_rn(s1[0], s1[1]), s1[2]);
  const float vol2 = __fmul_rn(__fmul_rn(s2[0], s2[1]), s2[2]);
  float e[3], o2[3], d2[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const Axis x = axis_of(c1[a], s1[a], c2[a], s2[a]);
    e[a] = clamp_min0(__fsub_rn(vmin(x.hi1, x.hi2), vmax(x.lo1, x.lo2)));
    const float o = clamp_min0(__fsub_rn(vmax(x.hi1, x.hi2), vmin(x.lo1, x.lo2)));
    o2[a] = __fmul_rn(o, o);
    const float d = __fsub_rn(c1[a], c2[a]);
    d2[a] = __fmul_rn(d, d);
  }
  const float inter = __fmul_rn(__fmul_rn(e[0], e[1]), e[2]);
  const float uni = __fsub_rn(__fadd_rn(vol1, vol2), inter);
  iou = __fdiv_rn(inter, uni);
  const float dcen = __fadd_rn(__fadd_rn(d2[0], d2[1]), d2[2]);
  const float odiag = __fadd_rn(__fadd_rn(o2[0], o2[1]), o2[2]);
  const float raw = __fsub_rn(iou, __fdiv_rn(__fmul_rn(1.5f, dcen), odiag));
  diou = raw < -1.f ? -1.f : (raw > 1.f ? 1.f : raw);
}

// d max(a,b)/da and d min(a,b)/da with torch's even split on ties (derivatives.yaml: maximum/minimum)
__device__ __forceinline__ float w_max(float a, float b) { return a > b ? 1.f : (a == b ? 0.5f : 0.f); }
__device__ __forceinline__ float w_min(float a, float b) { return a < b ? 1.f : (a == b ? 0.5f : 0.f); }

// Gradients of (iou, diou) w.r.t. both boxes given upstream g_iou, g_diou; accumulates into g*.
__device__ __forceinline__ void diou_backward(const float c1[3], const float s1[3], const float c2[3],
                                              const float s2[3], float g_iou, float g_diou, float gc1[3],
                                              float gs1[3], float gc2[3], float gs2[3]) {
  const float vol1 = s1[0] * s1[1] * s1[2], vol2 = s2[0] * s2[1] * s2[2];
  Axis x[3];
  float r[3], e[3], ro[3], o[3], d[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    x[a] = axis_of(c1[a], s1[a], c2[a], s2[a]);
    r[a] = vmin(x[a].hi1, x[a].hi2) - vmax(x[a].lo1, x[a].lo2);
    e[a] = clamp_min0(r[a]);
    ro[a] = vmax(x[a].hi1, x[a].hi2) - vmin(x[a].lo1, x[a].lo2);
    o[a] = clamp_min0(ro[a]);
    d[a] = c1[a] - c2[a];
  }
  const float inter = e[0] * e[1] * e[2];
  const float uni = vol1 + vol2 - inter;
  const float iou = inter / uni;
  const float dcen = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
  const float odiag = o[0] * o[0] + o[1] * o[1] + o[2] * o[2];
  const float raw = iou - 1.5f * dcen / odiag;
  const float g_raw = (raw >= -1.f && raw <= 1.f) ? g_diou : 0.f;  // clamp passes its boundary
  const float g_tot = g_iou + g_raw;
  const float g_inter = g_tot * (1.f / uni + inter / (uni * uni));
  const float g_vol = -g_tot * inter / (uni * uni);
  const float g_dcen = -1.5f * g_raw / odiag;
  const float g_odiag = 1.5f * g_raw * dcen / (odiag * odiag);
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const float others = a == 0 ? e[1] * e[2] : (a == 1 ? e[0] * e[2] : e[0] * e[1]);
    const float g_r = r[a] >= 0.f ? g_inter * others : 0.f;          // into iB - iA
    const float g_ro = ro[a] >= 0.f ? g_odiag * 2.f * o[a] : 0.f;    // into oB - oA
    // iB = min(hi1, hi2), iA = max(lo1, lo2), oB = max(hi1, hi2), oA = min(lo1, lo2)
    const float g_hi1 = g_r * w_min(x[a].hi1, x[a].hi2) + g_ro * w_max(x[a].hi1, x[a].hi2);
    const float g_hi2 = g_r * w_min(x[a].hi2, x[a].hi1) + g_ro * w_max(x[a].hi2, x[a].hi1);
    const float g_lo1 = -g_r * w_max(x[a].lo1, x[a].lo2) - g_ro * w_min(x[a].lo1, x[a].lo2);
    const float g_lo2 = -g_r * w_max(x[a].lo2, x[a].lo1) - g_ro * w_min(x[a].lo2, x[a].lo1);
    const float g_d = g_dcen * 2.f * d[a];
    const float v1o = a == 0 ? s1[1] * s1[2] : (a == 1 ? s1[0] * s1[2] : s1[0] * s1[1]);
    const float v2o = a == 0 ? s2[1] * s2[2] : (a == 1 ? s2[0] * s2[2] : s2[0] * s2[1]);
    gc1[a] += g_lo1 + g_hi1 + g_d;
    gc2[a] += g_lo2 + g_hi2 - g_d;
    gs1[a] += 0.5f * (g_hi1 - g_lo1) + g_vol * v1o;
    gs2[a] += 0.5f * (g_hi2 - g_lo2) + g_vol * v2o;
  }
}

__device__ __forceinline__ void load3(const float* p, size_t i, float v[3]) {
  v[0] = __ldg(p + 3 * i); v[1] = __ldg(p + 3 * i + 1); v[2] = __ldg(p + 3 * i + 2);
}

// all (scene, l, k) triples: box1 = (b, k) of K proposals, box2 = (b, l) of L targets
__global__ void __launch_bounds__(256)
diou_pairs_kernel(int L, int K, long long total, const float* __restrict__ c1, const float* __restrict__ s1,
                  const float* __restrict__ c2, const float* __restrict__ s2, float* __restrict__ iou,
                  float* __restrict__ diou) {
  const long long t = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
  if (t >= total) return;
  const long long bl = t / K;
  const int k = static_cast<int>(t - bl * K);
  const long long b = bl / L;
  float a[3], as[3], g[3], gs[3], vi, vd;
  load3(c1, b * K + k, a); load3(s1, b * K + k, as);
  load3(c2, bl, g); load3(s2, bl, gs);
  diou_forward(a, as, g, gs, vi, vd);
  if (iou) iou[t] = vi;
  if (diou) diou[t] = vd;
}

// gradient w.r.t. the proposals: thread = (scene, k), serial over the L targets (no atomics)
__global__ void __launch_bounds__(128)
diou_pairs_grad_kernel(int B, int L, int K, const float* __restrict__ c1, const float* __restrict__ s1,
                       const float* __restrict__ c2, const float* __restrict__ s2,
                       const float* __restrict__ g_iou, const float* __restrict__ g_diou,
                       float* __restrict__ gc1, float* __restrict__ gs1) {
  const long long t = static_cast<long long>(blockIdx.x) * 128 + threadIdx.x;
  if (t >= static_cast<long long>(B) * K) return;
  const long long b = t / K;
  const int k = static_cast<int>(t - b * K);
  float a[3], as[3], ga[3] = {0.f, 0.f, 0.f}, gas[3] = {0.f, 0.f, 0.f};
  load3(c1, t, a); load3(s1, t, as);
  for (int l = 0; l < L; ++l) {
    float g[3], gs[3], gg[3] = {0.f, 0.f, 0.f}, ggs[3] = {0.f, 0.f, 0.f};
    load3(c2, b * L + l, g); load3(s2, b * L + l, gs);
    const size_t o = (static_cast<size_t>(b) * L + l) * K + k;
    diou_backward(a, as, g, gs, g_iou ? __ldg(g_iou + o) : 0.f, g_diou ? __ldg(g_diou + o) : 0.f, ga, gas, gg, ggs);
  }
#pragma unroll
  for (int i = 0; i < 3; ++i) { gc1[3 * t + i] = ga[i]; gs1[3 * t + i] = gas[i]; }
}

// element-wise form of box_util.py:488 (P independent pairs), gradients for both boxes
__global__ void __launch_bounds__(256)
diou_elementwise_grad_kernel(long long P, const float* __restrict__ c1, const float* __restrict__ s1,
                             const float* __restrict__ c2, const float* __restrict__ s2,
                             const float* __restrict__ g_iou, const float* __restrict__ g_diou,
                             float* __restrict__ gc1, float* __restrict__ gs1, float* __restrict__ gc2,
                             float* __restrict__ gs2) {
  const long long t = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
  if (t >= P) return;
  float a[3], as[3], g[3], gs[3], ga[3] = {0, 0, 0}, gas[3] = {0, 0, 0}, gg[3] = {0, 0, 0}, ggs[3] = {0, 0, 0};
  load3(c1, t, a); load3(s1, t, as); load3(c2, t, g); load3(s2, t, gs);
  diou_backward(a, as, g, gs, g_iou ? __ldg(g_iou + t) : 0.f, g_diou ? __ldg(g_diou + t) : 0.f, ga, gas, gg, ggs);
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    if (gc1) gc1[3 * t + i] = ga[i];
    if (gs1) gs1[3 * t + i] = gas[i];
    if (gc2) gc2[3 * t + i] = gg[i];
    if (gs2) gs2[3 * t + i] = ggs[i];
  }
}

}  // namespace
}  // namespace pn2

using namespace pn2;

extern "C" int pn2_box3d_diou_pairs(int b, int l, int k, const float* center1, const float* size1,
                                    const float* center2, const float* size2, float* iou, float* diou,
                                    void* stream) {
  PN2_RANGE("pn2_box3d_diou_pairs");
  if (b <= 0 || l <= 0 || k <= 0) return PN2_OK;
  PN2_REQUIRE(center1 && size1 && center2 && size2 && (iou || diou), "box3d_diou_pairs: bad arguments");
  const long long total = static_cast<long long>(b) * l * k;
  diou_pairs_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, as_stream(stream)>>>(
      l, k, total, center1, size1, center2, size2, iou, diou);
  PN2_CHECK_LAUNCH("box3d_diou_pairs");
  return PN2_OK;
}

extern "C" int pn2_box3d_diou_pairs_grad(int b, int l, int k, const float* center1, const float* size1,
                                         const float* center2, const float* size2, const float* grad_iou,
                                         const float* grad_diou, float* grad_center1, float* grad_size1,
                                         void* stream) {
  PN2_RANGE("pn2_box3d_diou_pairs_grad");
  if (b <= 0 || k <= 0) return PN2_OK;
  PN2_REQUIRE(l >= 0 && center1 && size1 && (l == 0 || (center2 && size2)) && grad_center1 && grad_size1,
              "box3d_diou_pairs_grad: bad arguments");
  const long long total = static_cast<long long>(b) * k;
  diou_pairs_grad_kernel<<<static_cast<unsigned>((total + 127) / 128), 128, 0, as_stream(stream)>>>(
      b, l, k, center1, size1, center2, size2, grad_iou, grad_diou, grad_center1, grad_size1);
  PN2_CHECK_LAUNCH("box3d_diou_pairs_grad");
  return PN2_OK;
}

extern "C" int pn2_box3d_diou_grad(long long p, const float* center1, const float* size1, const float* center2,
                                   const float* size2, const float* grad_iou, const float* grad_diou,
                                   float* grad_center1, float* grad_size1, float* grad_center2,
                                   float* grad_size2, void* stream) {
  PN2_RANGE("pn2_box3d_diou_grad");
  if (p <= 0) return PN2_OK;
  PN2_REQUIRE(center1 && size1 && center2 && size2, "box3d_diou_grad: bad arguments");
  diou_elementwise_grad_kernel<<<static_cast<unsigned>((p + 255) / 256), 256, 0, as_stream(stream)>>>(
      p, center1, size1, center2, size2, grad_iou, grad_diou, grad_center1, grad_size1, grad_center2, grad_size2);
  PN2_CHECK_LAUNCH("box3d_diou_grad");
  return PN2_OK;
}
