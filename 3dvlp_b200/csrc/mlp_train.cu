// Training-mode SharedMLP (+ max-pool) on own kernels: batch-statistics BatchNorm, backward, weight gradients.
//
// Reference: pytorch_utils.py:11-36 (SharedMLP = per layer Conv2d 1x1 without bias -> BatchNorm2d -> ReLU),
// pointnet2_modules.py:251-257 (max_pool2d over nsample), executed by cuDNN in train() mode with batch
// statistics over all (B, npoint, nsample) positions (lib/joint/solver_3dvlp.py:580-584).
//
// Layout: activations are point-major rows, (P, C) with P = B * npoint * nsample positions.
//   forward, per layer   Y = X W^T on tcgen05 (pn2_linear_bf16: bf16 operands, fp32 accumulate, Y kept in fp32)
//                        pn2_column_stats      per-channel sum / sum of squares of Y, accumulated in double
//                        pn2_bn_finalize       mean, biased variance -> scale / shift; running statistics
//                        pn2_bn_relu_rows      X' = bf16(relu(Y scale + shift))          (hidden layers)
//                        pn2_bn_relu_maxpool   out[b,c,j] = max_s relu(Y scale + shift)  (last layer) + arg-max slot
//   backward, per layer  pn2_bn_bwd_stats      s1 = sum dYhat, s2 = sum dYhat xhat  (dYhat = upstream masked by ReLU;
//                                              the last layer's upstream is the pooled gradient at the arg-max slot)
//                        pn2_bn_bwd_rows       dY = gamma invstd (dYhat - s1/P - xhat s2/P)  -> bf16 rows
//                        pn2_wgrad_bf16        dW = dY^T X on tcgen05: MN-major operands, the rows fetched as they are by
//                                              tiled TMA (tensor maps {channel, term, position}; cp.async fallback),
//                                              split over K = P across persistent CTAs
//                        pn2_linear_bf16       dX = dY W (the same linear kernel, W^T packed as its weight)
//   pn2_unpack_rows      (P, ld) fp32 rows -> the reference's (B, C, n) layout (gradient w.r.t. the grouped tensor)
// The elementwise / reduction kernels are HBM-bound: 16-byte loads with a thread's channels fixed for its lifetime
// (per-channel terms in registers), rows strided over a CTA's contiguous range; each exists in a scalar form for
// rows that are not 16-byte aligned.
#include <cstdlib>

#include <cuda.h>  // CUtensorMap (types only: the encoder is fetched through the runtime, libcuda is not linked)

#include "tc.cuh"

namespace pn2 {
namespace {

constexpr int RT = 256;  // threads of the reduction / elementwise kernels

// ---- per-channel reductions over rows: thread = (row lane, channel), fp32 partials per thread over a
// bounded row range, shared-memory combine over the row lanes, ONE double atomic per channel per CTA ----
template <typename F>
__device__ __forceinline__ void column_reduce2(long long rows, int c, long long rows_per_cta, F f, double* o1, double* o2) {
  __shared__ float s_a[RT], s_b[RT];
  const int lanes = RT / c;                    // row lanes per CTA (c <= RT)
  const int tx = threadIdx.x % c, ty = threadIdx.x / c;
  const long long r0 = static_cast<long long>(blockIdx.x) * rows_per_cta;
  const long long r1 = r0 + rows_per_cta < rows ? r0 + rows_per_cta : rows;
  float a = 0.f, b = 0.f;
  if (ty < lanes) {
    for (long long r = r0 + ty; r < r1; r += lanes) {
      float va, vb;
      f(r, tx, va, vb);
      a += va; b += vb;
    }
  }
  s_a[threadIdx.x] = a; s_b[threadIdx.x] = b;
  __syncthreads();
  if (ty == 0) {
    double da = 0.0, db = 0.0;
    for (int l = 0; l < lanes; ++l) { da += s_a[l * c + tx]; db += s_b[l * c + tx]; }
    atomicAdd(o1 + tx, da);
    atomicAdd(o2 + tx, db);
  }
}

// The same reduction with thread = (row lane, FOUR channels): 16-byte loads, four rows in flight per thread.  A
// thread's channels are fixed, so per-channel constants of the functor live in registers.  Needs c % 4 == 0 and
// 16-byte aligned rows (checked by the callers, which fall back to the scalar kernels otherwise).
__device__ __forceinline__ float4 f4_add(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }

template <typename F>
__device__ __forceinline__ void column_reduce2_v4(long long rows, int c, long long rows_per_cta, F f, double* o1, double* o2) {
  __shared__ float4 s_a[RT], s_b[RT];
  const int cq = c >> 2;                       // channel quads (cq <= RT / 4... <= RT)
  const int lanes = RT / cq;
  const int tx = threadIdx.x % cq, ty = threadIdx.x / cq;
  const long long r0 = static_cast<long long>(blockIdx.x) * rows_per_cta;
  const long long r1 = r0 + rows_per_cta < rows ? r0 + rows_per_cta : rows;
  float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
  if (ty < lanes) {
    long long r = r0 + ty;
    for (; r + 3 * lanes < r1; r += 4 * lanes) {  // four rows in flight
      float4 a0, b0, a1, b1, a2, b2, a3, b3;
      f(r, a0, b0);
      f(r + lanes, a1, b1);
      f(r + 2 * lanes, a2, b2);
      f(r + 3 * lanes, a3, b3);
      a = f4_add(a, f4_add(f4_add(a0, a1), f4_add(a2, a3)));
      b = f4_add(b, f4_add(f4_add(b0, b1), f4_add(b2, b3)));
    }
    for (; r < r1; r += lanes) {
      float4 a0, b0;
      f(r, a0, b0);
      a = f4_add(a, a0);
      b = f4_add(b, b0);
    }
  }
  s_a[threadIdx.x] = a; s_b[threadIdx.x] = b;
  __syncthreads();
  if (ty == 0) {
    double da[4] = {0.0, 0.0, 0.0, 0.0}, db[4] = {0.0, 0.0, 0.0, 0.0};
    for (int l = 0; l < lanes; ++l) {
      const float4 va = s_a[l * cq + tx], vb = s_b[l * cq + tx];
      da[0] += va.x; da[1] += va.y; da[2] += va.z; da[3] += va.w;
      db[0] += vb.x; db[1] += vb.y; db[2] += vb.z; db[3] += vb.w;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) { atomicAdd(o1 + 4 * tx + k, da[k]); atomicAdd(o2 + 4 * tx + k, db[k]); }
  }
}

__global__ void __launch_bounds__(RT, 4)
column_stats_v4_kernel(long long rows, int c, int ld, long long rows_per_cta, const float* __restrict__ y,
                       double* __restrict__ sum, double* __restrict__ sumsq) {
  const int ch0 = (threadIdx.x % (c >> 2)) << 2;
  column_reduce2_v4(rows, c, rows_per_cta, [&](long long r, float4& a, float4& b) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(y + r * ld + ch0));
    a = v; b = make_float4(v.x * v.x, v.y * v.y, v.z * v.z, v.w * v.w);
  }, sum, sumsq);
}

__global__ void __launch_bounds__(RT)
column_stats_kernel(long long rows, int c, int ld, long long rows_per_cta, const float* __restrict__ y,
                    double* __restrict__ sum, double* __restrict__ sumsq) {
  column_reduce2(rows, c, rows_per_cta, [&](long long r, int ch, float& a, float& b) {
    const float v = __ldg(y + r * ld + ch);
    a = v; b = v * v;
  }, sum, sumsq);
}

__global__ void bn_finalize_kernel(int c, double count, const double* __restrict__ sum, const double* __restrict__ sumsq,
                                   const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                                   float momentum, float* __restrict__ running_mean, float* __restrict__ running_var,
                                   float* __restrict__ scale, float* __restrict__ shift, float* __restrict__ mean_out,
                                   float* __restrict__ invstd_out) {
  const int ch = blockIdx.x * blockDim.x + threadIdx.x;
  if (ch >= c) return;
  const double mean = sum[ch] / count;
  double var = sumsq[ch] / count - mean * mean;  // biased (what normalisation uses)
  if (var < 0.0) var = 0.0;
  const float invstd = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
  const float g = gamma ? gamma[ch] : 1.f, bt = beta ? beta[ch] : 0.f;
  scale[ch] = g * invstd;
  shift[ch] = bt - static_cast<float>(mean) * g * invstd;
  mean_out[ch] = static_cast<float>(mean);
  invstd_out[ch] = invstd;
  if (running_mean) {  // torch.nn.BatchNorm: running_var takes the UNBIASED batch variance
    const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
    running_mean[ch] = (1.f - momentum) * running_mean[ch] + momentum * static_cast<float>(mean);
    running_var[ch] = (1.f - momentum) * running_var[ch] + momentum * static_cast<float>(unbiased);
  }
}

// eval-mode statistics (gradients through a frozen BatchNorm): scale / shift from the running buffers
__global__ void bn_frozen_kernel(int c, const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                                 const float* __restrict__ running_mean, const float* __restrict__ running_var,
                                 float* __restrict__ scale, float* __restrict__ shift, float* __restrict__ mean_out,
                                 float* __restrict__ invstd_out) {
  const int ch = blockIdx.x * blockDim.x + threadIdx.x;
  if (ch >= c) return;
  const float invstd = 1.f / sqrtf(running_var[ch] + eps);
  const float g = gamma ? gamma[ch] : 1.f, bt = beta ? beta[ch] : 0.f;
  scale[ch] = g * invstd;
  shift[ch] = bt - running_mean[ch] * g * invstd;
  mean_out[ch] = running_mean[ch];
  invstd_out[ch] = invstd;
}

// Operand rows of the split-precision contractions.  A value v is carried as `terms` bf16 terms
// (h = bf16(v), m = bf16(v - h), l = bf16(v - h - m): 8, 16, 24 mantissa bits) laid out along K so that ONE
// K-concatenated GEMM against a weight laid out in the matching order evaluates the leading products, each
// exact in the fp32 accumulator:
//   terms 1   K blocks [h]                      x [w_h]                                       plain bf16
//   terms 2   K blocks [h | m | h]              x [w_h | w_h | w_m]                                    ~2^-16
//   terms 3   K blocks [h | m | h | l | h | m]  x [w_h | w_h | w_m | w_h | w_l | w_m]                   ~2^-24 (fp32 level)
// The WEIGHTS are stored with their K blocks spelled out (they are small); the ROWS store each term once
// ([h], [h | m], [h | m | l]) and the kernels map a logical K block to the stored one.
__device__ __forceinline__ void store_rows8(__nv_bfloat16* __restrict__ dst, long long r, int c, int ch0, const float (&o)[8],
                                            int terms) {
  uint32_t h[4], m[4], l[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const __nv_bfloat16 h0 = __float2bfloat16_rn(o[2 * e]), h1 = __float2bfloat16_rn(o[2 * e + 1]);
    const float r0 = o[2 * e] - __bfloat162float(h0), r1 = o[2 * e + 1] - __bfloat162float(h1);
    const __nv_bfloat16 m0 = __float2bfloat16_rn(r0), m1 = __float2bfloat16_rn(r1);
    h[e] = static_cast<uint32_t>(__bfloat16_as_ushort(h0)) | (static_cast<uint32_t>(__bfloat16_as_ushort(h1)) << 16);
    m[e] = static_cast<uint32_t>(__bfloat16_as_ushort(m0)) | (static_cast<uint32_t>(__bfloat16_as_ushort(m1)) << 16);
    l[e] = tc::pack_bf16(r0 - __bfloat162float(m0), r1 - __bfloat162float(m1));
  }
  const uint4 hv = make_uint4(h[0], h[1], h[2], h[3]), mv = make_uint4(m[0], m[1], m[2], m[3]),
              lv = make_uint4(l[0], l[1], l[2], l[3]);
  // each term is stored ONCE, [h], [h | m] or [h | m | l]: the contractions address the blocks through the product
  // tables (linear.cu kLinRowBlock, kWgProd* below) instead of reading duplicated copies
  __nv_bfloat16* row = dst + r * (terms <= 1 ? 1 : terms) * c + ch0;
  *reinterpret_cast<uint4*>(row) = hv;
  if (terms >= 2) *reinterpret_cast<uint4*>(row + c) = mv;
  if (terms >= 3) *reinterpret_cast<uint4*>(row + 2 * c) = lv;
}

// X' = relu(Y * scale + shift) as operand rows, thread = 8 channels of one row
__global__ void __launch_bounds__(RT)
bn_relu_rows_kernel(long long rows, int c, int ld, const float* __restrict__ y, const float* __restrict__ scale,
                    const float* __restrict__ shift, __nv_bfloat16* __restrict__ x, int terms) {
  const int chunks = c >> 3;
  const long long t = static_cast<long long>(blockIdx.x) * RT + threadIdx.x;
  if (t >= rows * chunks) return;
  const long long r = t / chunks;
  const int ch0 = static_cast<int>(t - r * chunks) << 3;
  const float4 a = __ldg(reinterpret_cast<const float4*>(y + r * ld + ch0));
  const float4 b = __ldg(reinterpret_cast<const float4*>(y + r * ld + ch0) + 1);
  const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
  float o[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) o[i] = fmaxf(__fmaf_rn(v[i], __ldg(scale + ch0 + i), __ldg(shift + ch0 + i)), 0.f);
  store_rows8(x, r, c, ch0, o, terms);
}

// out[b, ch, j] = max over the ns rows of group g = b * npoint + j of relu(y * scale + shift); arg-max = the FIRST
// slot that attains it (F.max_pool2d: strict '>' scan), stored point-major as argmax[g][ch]
__global__ void __launch_bounds__(RT)
bn_relu_maxpool_kernel(int groups, int ns, int c, int ld, int npoint, const float* __restrict__ y,
                       const float* __restrict__ scale, const float* __restrict__ shift, float* __restrict__ out_cm,
                       int* __restrict__ argmax) {
  const long long t = static_cast<long long>(blockIdx.x) * RT + threadIdx.x;
  if (t >= static_cast<long long>(groups) * c) return;
  const int g = static_cast<int>(t / c), ch = static_cast<int>(t - static_cast<long long>(g) * c);
  const float sc = __ldg(scale + ch), sh = __ldg(shift + ch);
  const float* row = y + static_cast<long long>(g) * ns * ld + ch;
  float best = -INFINITY;
  int bi = 0;
  for (int s = 0; s < ns; ++s) {
    const float v = fmaxf(__fmaf_rn(__ldg(row + static_cast<long long>(s) * ld), sc, sh), 0.f);
    if (s == 0 || v > best) { best = v; bi = s; }
  }
  const int b = g / npoint, j = g - b * npoint;
  out_cm[(static_cast<long long>(b) * c + ch) * npoint + j] = best;
  argmax[t] = bi;  // point-major (group, channel): coalesced here and in the backward kernels
}

// ---- backward ----------------------------------------------------------------------------------------
// dense upstream: dYhat = dA * [relu input > 0]
__global__ void __launch_bounds__(RT)
bn_bwd_stats_dense_kernel(long long rows, int c, int ld_da, int ld_y, long long rows_per_cta, const float* __restrict__ da,
                          const float* __restrict__ y, const float* __restrict__ scale, const float* __restrict__ shift,
                          const float* __restrict__ mean, const float* __restrict__ invstd, double* __restrict__ s1,
                          double* __restrict__ s2) {
  column_reduce2(rows, c, rows_per_cta, [&](long long r, int ch, float& a, float& b) {
    const float yv = __ldg(y + r * ld_y + ch);
    const float act = __fmaf_rn(yv, __ldg(scale + ch), __ldg(shift + ch));
    const float d = act > 0.f ? __ldg(da + r * ld_da + ch) : 0.f;
    a = d; b = d * (yv - __ldg(mean + ch)) * __ldg(invstd + ch);
  }, s1, s2);
}

__global__ void __launch_bounds__(RT, 4)
bn_bwd_stats_dense_v4_kernel(long long rows, int c, int ld_da, int ld_y, long long rows_per_cta, const float* __restrict__ da,
                             const float* __restrict__ y, const float* __restrict__ scale, const float* __restrict__ shift,
                             const float* __restrict__ mean, const float* __restrict__ invstd, double* __restrict__ s1,
                             double* __restrict__ s2) {
  const int ch0 = (threadIdx.x % (c >> 2)) << 2;
  const float4 sc = __ldg(reinterpret_cast<const float4*>(scale + ch0)), sh = __ldg(reinterpret_cast<const float4*>(shift + ch0));
  const float4 mu = __ldg(reinterpret_cast<const float4*>(mean + ch0)), is = __ldg(reinterpret_cast<const float4*>(invstd + ch0));
  column_reduce2_v4(rows, c, rows_per_cta, [&](long long r, float4& a, float4& b) {
    const float4 yv = __ldg(reinterpret_cast<const float4*>(y + r * ld_y + ch0));
    const float4 dv = __ldg(reinterpret_cast<const float4*>(da + r * ld_da + ch0));
    a.x = __fmaf_rn(yv.x, sc.x, sh.x) > 0.f ? dv.x : 0.f; b.x = a.x * (yv.x - mu.x) * is.x;
    a.y = __fmaf_rn(yv.y, sc.y, sh.y) > 0.f ? dv.y : 0.f; b.y = a.y * (yv.y - mu.y) * is.y;
    a.z = __fmaf_rn(yv.z, sc.z, sh.z) > 0.f ? dv.z : 0.f; b.z = a.z * (yv.z - mu.z) * is.z;
    a.w = __fmaf_rn(yv.w, sc.w, sh.w) > 0.f ? dv.w : 0.f; b.w = a.w * (yv.w - mu.w) * is.w;
  }, s1, s2);
}

// pooled upstream: the gradient of group g reaches only its arg-max row; "rows" of the reduction are groups
__global__ void __launch_bounds__(RT)
bn_bwd_stats_pool_kernel(long long groups, int ns, int c, int ld_y, int npoint, long long groups_per_cta,
                         const float* __restrict__ dout_pm, const int* __restrict__ argmax, const float* __restrict__ y,
                         const float* __restrict__ scale, const float* __restrict__ shift, const float* __restrict__ mean,
                         const float* __restrict__ invstd, double* __restrict__ s1, double* __restrict__ s2) {
  column_reduce2(groups, c, groups_per_cta, [&](long long g, int ch, float& a, float& b) {
    const long long o = g * c + ch;  // upstream gradient and arg-max are point-major (group, channel)
    const long long r = g * ns + __ldg(argmax + o);
    const float yv = __ldg(y + r * ld_y + ch);
    const float act = __fmaf_rn(yv, __ldg(scale + ch), __ldg(shift + ch));
    const float d = act > 0.f ? __ldg(dout_pm + o) : 0.f;
    a = d; b = d * (yv - __ldg(mean + ch)) * __ldg(invstd + ch);
  }, s1, s2);
}

// dY = gamma invstd (dYhat - (s1 + xhat s2) / count)  [count > 0: batch statistics]  or  gamma invstd dYhat
// [count == 0: frozen statistics] -> bf16 rows.  Upstream dense (da) or pooled (dout_pm + argmax, point-major).
__global__ void __launch_bounds__(RT)
bn_bwd_rows_kernel(long long rows, int c, int ld_da, int ld_y, const float* __restrict__ da,
                   const float* __restrict__ dout_pm, const int* __restrict__ argmax, int ns, int npoint,
                   const float* __restrict__ y, const float* __restrict__ scale, const float* __restrict__ shift,
                   const float* __restrict__ mean, const float* __restrict__ invstd, const float* __restrict__ gamma,
                   const double* __restrict__ s1, const double* __restrict__ s2, double count,
                   __nv_bfloat16* __restrict__ dy, int terms) {
  const int chunks = c >> 3;
  const long long t = static_cast<long long>(blockIdx.x) * RT + threadIdx.x;
  if (t >= rows * chunks) return;
  const long long r = t / chunks;
  const int ch0 = static_cast<int>(t - r * chunks) << 3;
  const float inv_count = count > 0.0 ? static_cast<float>(1.0 / count) : 0.f;
  long long g = 0;
  int slot = 0;
  if (!da) { g = r / ns; slot = static_cast<int>(r - g * ns); }
  float o[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int ch = ch0 + i;
    const float yv = __ldg(y + r * ld_y + ch);
    const float act = __fmaf_rn(yv, __ldg(scale + ch), __ldg(shift + ch));
    float up;
    if (da) {
      up = __ldg(da + r * ld_da + ch);
    } else {
      const long long oo = g * c + ch;
      up = __ldg(argmax + oo) == slot ? __ldg(dout_pm + oo) : 0.f;
    }
    const float d = act > 0.f ? up : 0.f;
    const float is = __ldg(invstd + ch);
    const float xhat = (yv - __ldg(mean + ch)) * is;
    const float corr = (static_cast<float>(s1[ch]) + xhat * static_cast<float>(s2[ch])) * inv_count;
    o[i] = (gamma ? __ldg(gamma + ch) : 1.f) * is * (d - corr);
  }
  store_rows8(dy, r, c, ch0, o, terms);
}

// The same with a thread's EIGHT channels fixed for its lifetime and its rows strided over the CTA's range: the
// per-channel terms collapse into three coefficients held in registers (dY = ka dYhat + kb y + kc), the rows are
// read and the upstream gradient fetched with 16-byte loads, two rows in flight.  Needs 16-byte aligned rows.
__global__ void __launch_bounds__(RT)
bn_bwd_rows_v8_kernel(long long rows, int c, int ld_da, int ld_y, long long rows_per_cta, const float* __restrict__ da,
                      const float* __restrict__ dout_pm, const int* __restrict__ argmax, int ns,
                      const float* __restrict__ y, const float* __restrict__ scale, const float* __restrict__ shift,
                      const float* __restrict__ mean, const float* __restrict__ invstd, const float* __restrict__ gamma,
                      const double* __restrict__ s1, const double* __restrict__ s2, double count,
                      __nv_bfloat16* __restrict__ dy, int terms) {
  const int chunks = c >> 3, lanes = RT / chunks;
  const int cx = threadIdx.x % chunks, ry = threadIdx.x / chunks;
  if (ry >= lanes) return;
  const int ch0 = cx << 3;
  const float inv_count = count > 0.0 ? static_cast<float>(1.0 / count) : 0.f;
  float sc[8], sh[8], ka[8], kb[8], kc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int ch = ch0 + i;
    const float is = __ldg(invstd + ch);
    const float a = (gamma ? __ldg(gamma + ch) : 1.f) * is;
    sc[i] = __ldg(scale + ch); sh[i] = __ldg(shift + ch);
    // a (d - (s1 + (y - mean) is s2) / count)  =  a d  -  [a is s2 / count] y  +  [a is s2 mean / count - a s1 / count]
    const float t2 = a * inv_count * static_cast<float>(s2[ch]) * is;
    ka[i] = a; kb[i] = -t2;
    kc[i] = __fmaf_rn(t2, __ldg(mean + ch), -(a * inv_count * static_cast<float>(s1[ch])));
  }
  const long long r0 = static_cast<long long>(blockIdx.x) * rows_per_cta;
  const long long r1 = r0 + rows_per_cta < rows ? r0 + rows_per_cta : rows;
#pragma unroll 2
  for (long long r = r0 + ry; r < r1; r += lanes) {
    const float4 ya = __ldg(reinterpret_cast<const float4*>(y + r * ld_y + ch0));
    const float4 yb = __ldg(reinterpret_cast<const float4*>(y + r * ld_y + ch0) + 1);
    const float yv[8] = {ya.x, ya.y, ya.z, ya.w, yb.x, yb.y, yb.z, yb.w};
    float up[8];
    if (da) {
      const float4 d0 = __ldg(reinterpret_cast<const float4*>(da + r * ld_da + ch0));
      const float4 d1 = __ldg(reinterpret_cast<const float4*>(da + r * ld_da + ch0) + 1);
      up[0] = d0.x; up[1] = d0.y; up[2] = d0.z; up[3] = d0.w; up[4] = d1.x; up[5] = d1.y; up[6] = d1.z; up[7] = d1.w;
    } else {
      const long long g = r / ns;
      const int slot = static_cast<int>(r - g * ns);
      const long long oo = g * c + ch0;
      const int4 a0 = __ldg(reinterpret_cast<const int4*>(argmax + oo)), a1 = __ldg(reinterpret_cast<const int4*>(argmax + oo) + 1);
      const float4 p0 = __ldg(reinterpret_cast<const float4*>(dout_pm + oo)), p1 = __ldg(reinterpret_cast<const float4*>(dout_pm + oo) + 1);
      up[0] = a0.x == slot ? p0.x : 0.f; up[1] = a0.y == slot ? p0.y : 0.f; up[2] = a0.z == slot ? p0.z : 0.f;
      up[3] = a0.w == slot ? p0.w : 0.f; up[4] = a1.x == slot ? p1.x : 0.f; up[5] = a1.y == slot ? p1.y : 0.f;
      up[6] = a1.z == slot ? p1.z : 0.f; up[7] = a1.w == slot ? p1.w : 0.f;
    }
    float o[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float d = __fmaf_rn(yv[i], sc[i], sh[i]) > 0.f ? up[i] : 0.f;
      o[i] = __fmaf_rn(ka[i], d, __fmaf_rn(kb[i], yv[i], kc[i]));
    }
    store_rows8(dy, r, c, ch0, o, terms);
  }
}

// (b * n, ld) fp32 rows, columns [0, c) -> (b, c, n) slices of a channel-major tensor with batch stride `bstride`
__global__ void __launch_bounds__(256)
unpack_rows_kernel(int c, int n, int ld, long long bstride, const float* __restrict__ rows, float* __restrict__ out) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z, p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int i = ty; i < 32; i += 8) {
    const int p = p0 + i, ch = c0 + tx;
    tile[i][tx] = (p < n && ch < c) ? __ldg(rows + (static_cast<long long>(b) * n + p) * ld + ch) : 0.f;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int ch = c0 + i, p = p0 + tx;
    if (ch < c && p < n) out[static_cast<long long>(b) * bstride + static_cast<long long>(ch) * n + p] = tile[tx][i];
  }
}

// ---- weight gradient: dW[co][ci] += sum_p dY[p][co] * X[p][ci] on tcgen05 -----------------------------
// Both operands arrive as point-major rows, i.e. with the contraction index p as the SLOW dimension and the
// channels contiguous: exactly the MN-MAJOR operand form of tcgen05.mma (instruction-descriptor bits a_major /
// b_major = 1).  A 64-position x 64-channel panel of either operand is copied row by row with 16-byte cp.async
// into the SWIZZLE_128B image (row = position, 128 bytes = 64 channels, chunk ^= row & 7) -- no transposition by
// the threads (round 2 first wrote K-major tiles through eight 4-byte shared-memory stores per 16-byte load; that
// store storm, not HBM, bounded the kernel).  Canonical MN-major SW128 layout (cute/atom/mma_traits_sm100.hpp):
// ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units -- 64 channels contiguous, the next 64-channel panel LBO bytes
// on, eight positions 128 bytes apart, the next eight SBO = 1024 bytes on; one MMA (K = 16 positions) advances
// the start address by 2048 bytes.
// CTA = one [128 x NT] block of dW (A rows = 128 output channels, zero rows beyond c_out) over a slice of P;
// 64 positions per step, two shared-memory stages, accumulator in TMEM, fp32 atomics into dW at the end.
constexpr int WG_THREADS = 128;

struct WgArgs {
  const __nv_bfloat16* dy;  // (rows, ld_dy), columns [0, c_out)
  const __nv_bfloat16* x;   // (rows, ld_x)
  float* dw;                // (c_out, c_in)
  int* err;
  long long rows;
  int c_out, c_in, ld_dy, ld_x, nt;  // nt = columns of this launch's N tile (multiple of 16, <= 256)
  long long chunks_per_cta;
  // split-precision operands: `terms` bf16 terms (h, m, l) of dy and of x, each term a column block of its row
  // layout ([h], [h|m], [h|m|l]); all leading (dy term, x term) products accumulate into ONE TMEM tile
  int terms, blk_dy, blk_x;          // block widths in elements (c_out, ld_x / terms)
};

// leading products (dy term, x term) per number of terms: 1, 3, 6 of them
__device__ __constant__ int kWgProdDy[6] = {0, 1, 0, 2, 0, 1};
__device__ __constant__ int kWgProdX[6] = {0, 0, 1, 0, 2, 1};

// 64-bit shared-memory descriptor of an MN-major SWIZZLE_128B operand: panels of 64 channels 8 KB apart (LBO),
// groups of eight positions 1 KB apart (SBO)
__device__ __forceinline__ uint64_t smem_desc_mn_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(8192 >> 4) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

__global__ void __launch_bounds__(WG_THREADS) wgrad_kernel(const WgArgs a) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_mma[2];
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  // a.nt is a multiple of 64 here: whole 64-channel panels of 64 positions x 128 bytes = 8 KB
  const uint32_t a_bytes = 2u * 8192u, b_bytes = static_cast<uint32_t>(a.nt >> 6) * 8192u;
  const uint32_t stage_bytes = static_cast<uint32_t>(a.terms) * (a_bytes + b_bytes);  // terms A tiles, then terms B tiles
  const int n_prod = a.terms == 1 ? 1 : (a.terms == 2 ? 3 : 6);
  const int co0 = blockIdx.y * 128, ci0 = blockIdx.z * a.nt;  // nz > 1 only with full nt-wide tiles
  const int co_cnt = min(128, a.c_out - co0);                 // multiple of 8
  const int ci_cnt = min(a.nt, ((a.c_in - ci0) + 7) & ~7);      // channels with data (x blocks are padded to 8)
  uint32_t tmem_cols = 32;
  while (tmem_cols < static_cast<uint32_t>(a.nt)) tmem_cols <<= 1;
  if (tid == 0) {
    mbar_init(&bar_mma[0], 1); mbar_init(&bar_mma[1], 1);
    fence_mbar_init_cluster();
  }
  if (warp == 0) tc::tmem_alloc(&tmem_slot, tmem_cols);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_d = tmem_slot;
  // M = 128, N = nt, both operands MN-major (bits 15 / 16)
  const uint32_t idesc = tc::idesc_bf16_f32(128, a.nt) | (1u << 15) | (1u << 16);
  const long long total_chunks = (a.rows + 63) / 64;
  const long long c_first = static_cast<long long>(blockIdx.x) * a.chunks_per_cta;
  long long c_last = c_first + a.chunks_per_cta;
  if (c_last > total_chunks) c_last = total_chunks;
  const int n_chunks = c_last > c_first ? static_cast<int>(c_last - c_first) : 0;
  const int b_chunks = a.nt >> 3;  // 16-byte chunks per position of a B tile

  // every term of both operands is read ONCE per chunk of 64 positions
  auto load = [&](int i, int s) {
    const long long row0 = (c_first + i) * 64;
    const uint32_t sa = base + s * stage_bytes, sb = sa + static_cast<uint32_t>(a.terms) * a_bytes;
    for (int t = 0; t < a.terms; ++t) {
      const __nv_bfloat16* dy_t = a.dy + t * a.blk_dy + co0;
      const __nv_bfloat16* x_t = a.x + t * a.blk_x + ci0;
#pragma unroll
      for (int u = 0; u < 8; ++u) {  // A: 64 positions x 16 chunks
        const int q = u * WG_THREADS + tid, k = q >> 4, j = q & 15;
        const bool ok = row0 + k < a.rows && j * 8 < co_cnt;
        const __nv_bfloat16* src = ok ? dy_t + (row0 + k) * a.ld_dy + j * 8 : a.dy;
        tc::cp_async16(sa + t * a_bytes + (j >> 3) * 8192u + tc::swz_off(k, j & 7), src, ok);
      }
      for (int q = tid; q < 64 * b_chunks; q += WG_THREADS) {  // B: 64 positions x nt / 8 chunks
        const int k = q / b_chunks, j = q - k * b_chunks;
        const bool ok = row0 + k < a.rows && j * 8 < ci_cnt;
        const __nv_bfloat16* src = ok ? x_t + (row0 + k) * a.ld_x + j * 8 : a.x;
        tc::cp_async16(sb + t * b_bytes + (j >> 3) * 8192u + tc::swz_off(k, j & 7), src, ok);
      }
    }
    tc::cp_async_commit();
  };

  bool failed = false;
  int n_issued = 0;
  if (n_chunks > 0) load(0, 0);
  for (int i = 0; i < n_chunks; ++i, ++n_issued) {
    const int s = i & 1;
    if (i + 1 < n_chunks) {
      // stage s^1 was last read by the MMAs of chunk i-1
      if (i >= 1 && !mbar_wait(&bar_mma[s ^ 1], ((i - 1) >> 1) & 1)) { failed = true; break; }
      load(i + 1, s ^ 1);
      tc::cp_async_wait<1>();
    } else {
      tc::cp_async_wait<0>();
    }
    fence_proxy_async_smem();
    tc::fence_before_sync();
    __syncthreads();
    if (tid == 0) {
      tc::fence_after_sync();
      const uint32_t sa = base + s * stage_bytes, sb = sa + static_cast<uint32_t>(a.terms) * a_bytes;
      for (int pr = 0; pr < n_prod; ++pr) {
        const uint32_t ta = sa + kWgProdDy[pr] * a_bytes, tb = sb + kWgProdX[pr] * b_bytes;
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)  // 16 positions per MMA = two 1 KB groups of eight
          tc::mma_bf16(tmem_d, smem_desc_mn_sw128(ta + kk * 2048), smem_desc_mn_sw128(tb + kk * 2048), idesc,
                       (i > 0 || pr > 0 || kk > 0) ? 1u : 0u);
      }
      tc::mma_commit(&bar_mma[s]);
    }
  }
  if (!failed && n_issued > 0) {
    const int last = n_issued - 1;
    if (!mbar_wait(&bar_mma[last & 1], (last >> 1) & 1)) failed = true;
    // the other stage's commit completed earlier (commits complete in order)
  }
  if (failed) {
    if (tid == 0) atomicExch(a.err, 2);
  } else if (n_issued > 0) {
    tc::fence_after_sync();
    const int co = co0 + tid;  // thread = TMEM lane = output channel
    const uint32_t lane_base = tmem_d + (static_cast<uint32_t>(warp * 32) << 16);
    for (int c0 = 0; c0 < a.nt; c0 += 16) {
      float v[16];
      tc::tmem_ld16(lane_base + c0, v);
      if (co < a.c_out) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int ci = ci0 + c0 + i;
          if (ci < a.c_in) atomicAdd(a.dw + static_cast<long long>(co) * a.c_in + ci, v[i]);
        }
      }
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem_d, tmem_cols);
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// ---- the same contraction with the operand tiles fetched by TILED TMA (cp.async.bulk.tensor, SASS UTMALDG) --------
// Each operand is described to the TMA unit as a 3-D tensor {channel, term, position} (strides: 2 bytes, one term
// block, one row); a box of 64 channels x 1 term x 64 positions with SWIZZLE_128B lands in shared memory as exactly
// the MN-major panel image above -- address generation, swizzle and the zero fill of channels >= c_out / c_in and of
// positions >= rows happen in the copy engine.  Warp-specialised: lane 0 of warp 0 only issues copies (it runs as far
// ahead as the ring has free stages), lane 0 of warp 1 only issues MMAs; full / empty mbarriers per stage
// (expect_tx / tcgen05.commit) replace the per-chunk __syncthreads of the cp.async version, whose 128 threads spent
// their time on address arithmetic for 16-byte copies and kept at most two chunks in flight (2-3 TB/s on the
// SA1-sized layers).
constexpr int WG_MAX_STAGES = 6;

__global__ void __launch_bounds__(WG_THREADS)
wgrad_tma_kernel(const __grid_constant__ CUtensorMap tm_dy, const __grid_constant__ CUtensorMap tm_x, const WgArgs a,
                 const int stages) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_full[WG_MAX_STAGES], bar_empty[WG_MAX_STAGES], bar_done;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_bytes = 2u * 8192u, b_bytes = static_cast<uint32_t>(a.nt >> 6) * 8192u;
  const uint32_t stage_bytes = static_cast<uint32_t>(a.terms) * (a_bytes + b_bytes);
  const int n_prod = a.terms == 1 ? 1 : (a.terms == 2 ? 3 : 6);
  const int co0 = blockIdx.y * 128, ci0 = blockIdx.z * a.nt;
  uint32_t tmem_cols = 32;
  while (tmem_cols < static_cast<uint32_t>(a.nt)) tmem_cols <<= 1;
  if (tid == 0) {
    for (int s = 0; s < stages; ++s) { mbar_init(&bar_full[s], 1); mbar_init(&bar_empty[s], 1); }
    mbar_init(&bar_done, 1);
    fence_mbar_init_cluster();
    tc::tma_prefetch_desc(&tm_dy);
    tc::tma_prefetch_desc(&tm_x);
  }
  if (warp == 0) tc::tmem_alloc(&tmem_slot, tmem_cols);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_d = tmem_slot;
  const uint32_t idesc = tc::idesc_bf16_f32(128, a.nt) | (1u << 15) | (1u << 16);  // both operands MN-major
  const long long total_chunks = (a.rows + 63) / 64;
  const long long c_first = static_cast<long long>(blockIdx.x) * a.chunks_per_cta;
  long long c_last = c_first + a.chunks_per_cta;
  if (c_last > total_chunks) c_last = total_chunks;
  const int n_chunks = c_last > c_first ? static_cast<int>(c_last - c_first) : 0;
  bool failed = false;
  if (warp == 0) {
    if (lane == 0) {  // ---- copy issue ----
      for (int i = 0; i < n_chunks; ++i) {
        const int s = i % stages, round = i / stages;
        if (round > 0 && !mbar_wait(&bar_empty[s], (round - 1) & 1)) { failed = true; break; }
        mbar_arrive_expect_tx(&bar_full[s], stage_bytes);
        const int row0 = static_cast<int>((c_first + i) * 64);
        const uint32_t sa = base + s * stage_bytes, sb = sa + static_cast<uint32_t>(a.terms) * a_bytes;
        for (int t = 0; t < a.terms; ++t) {
          tc::tma_load_3d(sa + t * a_bytes, &tm_dy, co0, t, row0, &bar_full[s]);
          tc::tma_load_3d(sa + t * a_bytes + 8192u, &tm_dy, co0 + 64, t, row0, &bar_full[s]);
          for (int pnl = 0; pnl < (a.nt >> 6); ++pnl)
            tc::tma_load_3d(sb + t * b_bytes + pnl * 8192u, &tm_x, ci0 + 64 * pnl, t, row0, &bar_full[s]);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {  // ---- MMA issue ----
      for (int i = 0; i < n_chunks; ++i) {
        const int s = i % stages, round = i / stages;
        if (!mbar_wait(&bar_full[s], round & 1)) { failed = true; break; }
        tc::fence_after_sync();
        const uint32_t sa = base + s * stage_bytes, sb = sa + static_cast<uint32_t>(a.terms) * a_bytes;
        for (int pr = 0; pr < n_prod; ++pr) {
          const uint32_t ta = sa + kWgProdDy[pr] * a_bytes, tb = sb + kWgProdX[pr] * b_bytes;
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            tc::mma_bf16(tmem_d, smem_desc_mn_sw128(ta + kk * 2048), smem_desc_mn_sw128(tb + kk * 2048), idesc,
                         (i > 0 || pr > 0 || kk > 0) ? 1u : 0u);
        }
        tc::mma_commit(&bar_empty[s]);  // the stage is free once these MMAs have read it
      }
      if (!failed) tc::mma_commit(&bar_done);
    }
    __syncwarp();
  }
  if (failed) atomicExch(a.err, 2);  // (a lane that bailed out; the others below time out on bar_done)
  if (n_chunks > 0) {
    if (!mbar_wait(&bar_done, 0)) {
      if (tid == 0) atomicExch(a.err, 2);
    } else {
      tc::fence_after_sync();
      const int co = co0 + tid;  // thread = TMEM lane = output channel
      const uint32_t lane_base = tmem_d + (static_cast<uint32_t>(warp * 32) << 16);
      for (int c0 = 0; c0 < a.nt; c0 += 16) {
        float v[16];
        tc::tmem_ld16(lane_base + c0, v);
        if (co < a.c_out) {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int ci = ci0 + c0 + i;
            if (ci < a.c_in) atomicAdd(a.dw + static_cast<long long>(co) * a.c_in + ci, v[i]);
          }
        }
      }
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem_d, tmem_cols);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// cuTensorMapEncodeTiled through the runtime's driver entry point query (no link dependency on libcuda)
EncodeTiledFn tensor_map_encoder() {
  static void* cached = nullptr;
  void* p = __atomic_load_n(&cached, __ATOMIC_ACQUIRE);
  if (p) return reinterpret_cast<EncodeTiledFn>(p);
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
      q != cudaDriverEntryPointSuccess || !p) {
    cudaGetLastError();
    return nullptr;
  }
  __atomic_store_n(&cached, p, __ATOMIC_RELEASE);
  return reinterpret_cast<EncodeTiledFn>(p);
}

// {channel, term, position} view of split-precision operand rows: `valid` channels per term block, blocks `blk`
// elements apart, rows `ld` elements apart; boxes of 64 channels x 1 term x 64 positions, SWIZZLE_128B
bool encode_rows_map(CUtensorMap* tm, const void* ptr, long long rows, int valid, int terms, int blk, int ld) {
  EncodeTiledFn enc = tensor_map_encoder();
  if (!enc || rows <= 0 || rows >= (1LL << 31) || valid <= 0) return false;
  const cuuint64_t dims[3] = {static_cast<cuuint64_t>(valid), static_cast<cuuint64_t>(terms), static_cast<cuuint64_t>(rows)};
  const cuuint64_t strides[2] = {static_cast<cuuint64_t>(blk) * 2, static_cast<cuuint64_t>(ld) * 2};
  const cuuint32_t box[3] = {64, 1, 64}, estr[3] = {1, 1, 1};
  return enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// the 16-byte kernels: ONE wave of four CTAs per SM -- every CTA ends with a double atomic per channel on the same
// 2 c addresses, and those serialise in L2, so fewer, longer CTAs (a thread still sums < 4096 values in fp32)
long long pick_rows_per_cta_v4(long long rows, int c) {
  const long long lanes = RT / (c / 4);
  long long per = (rows + 4LL * sm_count() - 1) / (4LL * sm_count());
  const long long cap = 4096 * lanes;
  if (per > cap) per = cap;
  if (per < lanes) per = lanes;
  return per;
}

long long pick_rows_per_cta(long long rows, int c) {
  // enough CTAs to fill the machine, at most ~4096 values per thread partial (fp32 accumulation error)
  const long long lanes = RT / c > 0 ? RT / c : 1;
  long long per = (rows + 8LL * sm_count() - 1) / (8LL * sm_count());
  const long long cap = 4096 * lanes;
  if (per > cap) per = cap;
  if (per < lanes) per = lanes;
  return per;
}

}  // namespace

int* device_err_flag();

}  // namespace pn2

using namespace pn2;

extern "C" int pn2_column_stats(long long rows, int c, int ld, const float* y, double* sum, double* sumsq, void* stream) {
  PN2_RANGE("pn2_column_stats");
  if (rows <= 0) return PN2_OK;
  PN2_REQUIRE(c > 0 && c <= RT && ld >= c && y && sum && sumsq, "column_stats: bad arguments (c=%d must be <= %d)", c, RT);
  if (c % 4 == 0 && ld % 4 == 0 && aligned16(y)) {
    const long long per = pick_rows_per_cta_v4(rows, c);
    column_stats_v4_kernel<<<static_cast<unsigned>((rows + per - 1) / per), RT, 0, as_stream(stream)>>>(rows, c, ld, per, y, sum, sumsq);
  } else {
    const long long per = pick_rows_per_cta(rows, c);
    column_stats_kernel<<<static_cast<unsigned>((rows + per - 1) / per), RT, 0, as_stream(stream)>>>(rows, c, ld, per, y, sum, sumsq);
  }
  PN2_CHECK_LAUNCH("column_stats_kernel");
  return PN2_OK;
}

extern "C" int pn2_bn_finalize(int c, long long count, const double* sum, const double* sumsq, const float* gamma,
                               const float* beta, float eps, float momentum, float* running_mean, float* running_var,
                               float* scale, float* shift, float* mean, float* invstd, void* stream) {
  PN2_RANGE("pn2_bn_finalize");
  PN2_REQUIRE(c > 0 && scale && shift && mean && invstd, "bn_finalize: bad arguments");
  if (count > 0) {
    PN2_REQUIRE(sum && sumsq, "bn_finalize: batch statistics need sum / sumsq");
    bn_finalize_kernel<<<(c + 127) / 128, 128, 0, as_stream(stream)>>>(c, static_cast<double>(count), sum, sumsq, gamma, beta,
                                                                        eps, momentum, running_mean, running_var, scale, shift,
                                                                        mean, invstd);
  } else {  // count == 0: frozen (eval-mode) statistics from the running buffers
    PN2_REQUIRE(running_mean && running_var, "bn_finalize: frozen statistics need the running buffers");
    bn_frozen_kernel<<<(c + 127) / 128, 128, 0, as_stream(stream)>>>(c, gamma, beta, eps, running_mean, running_var, scale,
                                                                      shift, mean, invstd);
  }
  PN2_CHECK_LAUNCH("bn_finalize_kernel");
  return PN2_OK;
}

extern "C" int pn2_bn_relu_rows(long long rows, int c, int ld, const float* y, const float* scale, const float* shift,
                                int terms, void* x_bf16, void* stream) {
  PN2_RANGE("pn2_bn_relu_rows");
  if (rows <= 0) return PN2_OK;
  PN2_REQUIRE(c > 0 && c % 8 == 0 && ld >= c && ld % 4 == 0 && y && scale && shift && x_bf16 && terms >= 1 && terms <= 3,
              "bn_relu_rows: bad arguments");
  const long long total = rows * (c / 8);
  bn_relu_rows_kernel<<<static_cast<unsigned>((total + RT - 1) / RT), RT, 0, as_stream(stream)>>>(
      rows, c, ld, y, scale, shift, static_cast<__nv_bfloat16*>(x_bf16), terms);
  PN2_CHECK_LAUNCH("bn_relu_rows_kernel");
  return PN2_OK;
}

extern "C" int pn2_bn_relu_maxpool(int groups, int ns, int c, int ld, int npoint, const float* y, const float* scale,
                                   const float* shift, float* out_cm, int* argmax, void* stream) {
  PN2_RANGE("pn2_bn_relu_maxpool");
  if (groups <= 0) return PN2_OK;
  PN2_REQUIRE(ns > 0 && c > 0 && ld >= c && npoint > 0 && groups % npoint == 0 && y && scale && shift && out_cm && argmax,
              "bn_relu_maxpool: bad arguments");
  const long long total = static_cast<long long>(groups) * c;
  bn_relu_maxpool_kernel<<<static_cast<unsigned>((total + RT - 1) / RT), RT, 0, as_stream(stream)>>>(
      groups, ns, c, ld, npoint, y, scale, shift, out_cm, argmax);
  PN2_CHECK_LAUNCH("bn_relu_maxpool_kernel");
  return PN2_OK;
}

extern "C" int pn2_bn_bwd_stats(long long rows, int c, int ld_da, int ld_y, const float* da, const float* dout_pm,
                                const int* argmax, int ns, int npoint, const float* y, const float* scale,
                                const float* shift, const float* mean, const float* invstd, double* s1, double* s2,
                                void* stream) {
  PN2_RANGE("pn2_bn_bwd_stats");
  if (rows <= 0) return PN2_OK;
  PN2_REQUIRE(c > 0 && c <= RT && y && scale && shift && mean && invstd && s1 && s2, "bn_bwd_stats: bad arguments");
  if (da) {
    if (c % 4 == 0 && ld_da % 4 == 0 && ld_y % 4 == 0 && aligned16(da) && aligned16(y) && aligned16(scale) &&
        aligned16(shift) && aligned16(mean) && aligned16(invstd)) {
      const long long per = pick_rows_per_cta_v4(rows, c);
      bn_bwd_stats_dense_v4_kernel<<<static_cast<unsigned>((rows + per - 1) / per), RT, 0, as_stream(stream)>>>(
          rows, c, ld_da, ld_y, per, da, y, scale, shift, mean, invstd, s1, s2);
    } else {
      const long long per = pick_rows_per_cta(rows, c);
      bn_bwd_stats_dense_kernel<<<static_cast<unsigned>((rows + per - 1) / per), RT, 0, as_stream(stream)>>>(
          rows, c, ld_da, ld_y, per, da, y, scale, shift, mean, invstd, s1, s2);
    }
  } else {
    PN2_REQUIRE(dout_pm && argmax && ns > 0 && npoint > 0 && rows % ns == 0, "bn_bwd_stats: pooled upstream needs dout / argmax");
    const long long groups = rows / ns;
    const long long per = pick_rows_per_cta(groups, c);
    bn_bwd_stats_pool_kernel<<<static_cast<unsigned>((groups + per - 1) / per), RT, 0, as_stream(stream)>>>(
        groups, ns, c, ld_y, npoint, per, dout_pm, argmax, y, scale, shift, mean, invstd, s1, s2);
  }
  PN2_CHECK_LAUNCH("bn_bwd_stats_kernel");
  return PN2_OK;
}

extern "C" int pn2_bn_bwd_rows(long long rows, int c, int ld_da, int ld_y, const float* da, const float* dout_pm,
                               const int* argmax, int ns, int npoint, const float* y, const float* scale,
                               const float* shift, const float* mean, const float* invstd, const float* gamma,
                               const double* s1, const double* s2, long long count, int terms, void* dy_bf16,
                               void* stream) {
  PN2_RANGE("pn2_bn_bwd_rows");
  if (rows <= 0) return PN2_OK;
  PN2_REQUIRE(c > 0 && c % 8 == 0 && y && scale && shift && mean && invstd && s1 && s2 && dy_bf16 && terms >= 1 && terms <= 3,
              "bn_bwd_rows: bad arguments");
  PN2_REQUIRE(da || (dout_pm && argmax && ns > 0 && npoint > 0), "bn_bwd_rows: no upstream gradient");
  const long long total = rows * (c / 8);
  if (c <= 8 * RT && ld_y % 4 == 0 && aligned16(y) && aligned16(dy_bf16) &&
      (da ? (ld_da % 4 == 0 && aligned16(da)) : (aligned16(dout_pm) && aligned16(argmax)))) {
    // a CTA covers a contiguous range of rows: RT / (c / 8) of them per pass, enough CTAs for ~8 per SM
    const long long lanes = RT / (c / 8);
    long long per = (rows + 8LL * sm_count() - 1) / (8LL * sm_count());
    per = (per + lanes - 1) / lanes * lanes;
    if (per < 4 * lanes) per = 4 * lanes;
    bn_bwd_rows_v8_kernel<<<static_cast<unsigned>((rows + per - 1) / per), RT, 0, as_stream(stream)>>>(
        rows, c, ld_da, ld_y, per, da, dout_pm, argmax, ns, y, scale, shift, mean, invstd, gamma, s1, s2,
        static_cast<double>(count), static_cast<__nv_bfloat16*>(dy_bf16), terms);
  } else {
    bn_bwd_rows_kernel<<<static_cast<unsigned>((total + RT - 1) / RT), RT, 0, as_stream(stream)>>>(
        rows, c, ld_da, ld_y, da, dout_pm, argmax, ns, npoint, y, scale, shift, mean, invstd, gamma, s1, s2,
        static_cast<double>(count), static_cast<__nv_bfloat16*>(dy_bf16), terms);
  }
  PN2_CHECK_LAUNCH("bn_bwd_rows_kernel");
  return PN2_OK;
}

extern "C" int pn2_unpack_rows(int b, int c, int n, int ld, long long batch_stride, const float* rows, float* out,
                               void* stream) {
  PN2_RANGE("pn2_unpack_rows");
  if (b <= 0 || c <= 0 || n <= 0) return PN2_OK;
  PN2_REQUIRE(ld >= c && rows && out && b <= 65535 && batch_stride >= static_cast<long long>(c) * n, "unpack_rows: bad arguments");
  dim3 grid((n + 31) / 32, (c + 31) / 32, b);
  unpack_rows_kernel<<<grid, 256, 0, as_stream(stream)>>>(c, n, ld, batch_stride, rows, out);
  PN2_CHECK_LAUNCH("unpack_rows_kernel");
  return PN2_OK;
}

static int wgrad_launch(long long rows, int c_out, int c_in, int ld_dy, int ld_x, const void* dy_bf16,
                        const void* x_bf16, int terms, float* dw, void* stream, const char* who) {
  if (rows <= 0) return PN2_OK;
  PN2_REQUIRE(terms >= 1 && terms <= 3, "%s: terms=%d must be 1, 2 or 3", who, terms);
  const int nblk = terms;  // stored blocks per row
  PN2_REQUIRE(c_out > 0 && c_out % 8 == 0 && c_in > 0 && ld_x % (8 * nblk) == 0 && ld_x / nblk >= ((c_in + 7) & ~7) &&
                  ld_dy % 8 == 0 && ld_dy >= nblk * c_out && dy_bf16 && x_bf16 && dw,
              "%s: bad arguments (c_out %% 8, ld_dy / ld_x %% 8, every x block >= c_in rounded up to 8)", who);
  PN2_REQUIRE(((reinterpret_cast<uintptr_t>(dy_bf16) | reinterpret_cast<uintptr_t>(x_bf16)) & 15) == 0,
              "%s: operands must be 16-byte aligned", who);
  WgArgs a;
  a.dy = static_cast<const __nv_bfloat16*>(dy_bf16); a.x = static_cast<const __nv_bfloat16*>(x_bf16); a.dw = dw;
  a.rows = rows; a.c_out = c_out; a.c_in = c_in; a.ld_dy = ld_dy; a.ld_x = ld_x;
  a.terms = terms; a.blk_dy = c_out; a.blk_x = ld_x / nblk;
  a.err = device_err_flag();
  PN2_REQUIRE(a.err != nullptr, "%s: cannot allocate the device error flag", who);
  // N tile: up to 256 input channels, 128 with three terms (two stages of 3 x (16 + 16) KB must fit shared memory)
  const int nt_max = terms == 3 ? 128 : 256;
  const int nt = c_in >= nt_max ? nt_max : ((c_in + 63) & ~63);  // whole 64-channel MN-major panels
  const int nz = (c_in + nt - 1) / nt, ny = (c_out + 127) / 128;
  a.nt = nt;
  const long long chunks = (rows + 63) / 64;
  long long ctas = 2LL * sm_count() / (ny * nz);
  if (ctas < 1) ctas = 1;
  if (ctas > chunks) ctas = chunks;
  a.chunks_per_cta = (chunks + ctas - 1) / ctas;
  ctas = (chunks + a.chunks_per_cta - 1) / a.chunks_per_cta;
  const size_t stage = static_cast<size_t>(terms) * (2 * 8192 + static_cast<size_t>(nt / 64) * 8192);
  dim3 grid(static_cast<unsigned>(ctas), ny, nz);
  // tiled TMA path: the operands as {channel, term, position} tensor maps (needs rows laid out as whole term blocks)
  CUtensorMap tm_dy, tm_x;
  const int ci_valid = (c_in + 7) & ~7;
  if (option_value(OPT_WGRAD_TMA) && ld_dy % nblk == 0 && a.blk_dy * nblk <= ld_dy &&
      encode_rows_map(&tm_dy, dy_bf16, rows, c_out, terms, a.blk_dy, ld_dy) &&
      encode_rows_map(&tm_x, x_bf16, rows, ci_valid, terms, a.blk_x, ld_x)) {
    int stages = static_cast<int>((225 * 1024) / stage);
    if (stages > WG_MAX_STAGES) stages = WG_MAX_STAGES;
    if (stages >= 2) {
      const size_t smem = stages * stage + 1024;
      PN2_CUDA(cudaFuncSetAttribute(wgrad_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
      wgrad_tma_kernel<<<grid, WG_THREADS, smem, as_stream(stream)>>>(tm_dy, tm_x, a, stages);
      PN2_CHECK_LAUNCH("wgrad_tma_kernel");
      stat_add(STAT_WGRAD_TMA);
      return PN2_OK;
    }
  }
  const size_t smem = 2 * stage + 1024;
  PN2_CUDA(cudaFuncSetAttribute(wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  wgrad_kernel<<<grid, WG_THREADS, smem, as_stream(stream)>>>(a);
  PN2_CHECK_LAUNCH("wgrad_kernel");
  stat_add(STAT_WGRAD_CPASYNC);
  return PN2_OK;
}

extern "C" int pn2_wgrad_bf16(long long rows, int c_out, int c_in, int ld_dy, int ld_x, const void* dy_bf16,
                              const void* x_bf16, float* dw, void* stream) {
  PN2_RANGE("pn2_wgrad_bf16");
  return wgrad_launch(rows, c_out, c_in, ld_dy, ld_x, dy_bf16, x_bf16, 1, dw, stream, "wgrad_bf16");
}

extern "C" int pn2_wgrad_split_bf16(long long rows, int c_out, int c_in, int ld_dy, int ld_x, const void* dy_bf16,
                                    const void* x_bf16, int terms, float* dw, void* stream) {
  PN2_RANGE("pn2_wgrad_split_bf16");
  return wgrad_launch(rows, c_out, c_in, ld_dy, ld_x, dy_bf16, x_bf16, terms, dw, stream, "wgrad_split_bf16");
}
