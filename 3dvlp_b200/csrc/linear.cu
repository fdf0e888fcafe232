// Point-wise linear layer (1x1 convolution over points) on tcgen05 tensor cores:
//     Y[m][n] = act( sum_k X[m][k] * W[n][k] + bias[n] ),   bf16 operands, fp32 accumulation in TMEM.
//
// This is the dense contraction behind SharedMLP (reference pytorch_utils.py:11-36: Conv2d 1x1 ->
// BatchNorm2d -> ReLU, executed there by cuDNN in fp32/TF32).  BatchNorm (eval mode) is folded into
// W and bias on the host.  Used for (a) the per-point half of every SA layer's first convolution
// -- a 1x1 convolution commutes with the neighbourhood gather, so it is applied ONCE per input point
// here instead of once per (centre, neighbour) pair after grouping -- and (b) the FP-module MLPs.
//
// One CTA computes a 128-row tile: UMMA M = 128 (rows on the 128 TMEM lanes), N = n (<= 256 fp32
// columns), K streamed in 64-wide chunks through a 2-stage shared-memory ring.  The weight chunk is
// one contiguous pre-swizzled panel fetched with a 1-D TMA bulk copy (cp.async.bulk + mbarrier);
// the activation chunk is copied with 16-byte cp.async straight into the SWIZZLE_128B layout.
// A single elected thread issues tcgen05.mma and commits to an mbarrier that releases the stage.
#include "tc.cuh"

namespace pn2 {
namespace {

constexpr int LIN_THREADS = 128;

struct LinArgs {
  const __nv_bfloat16* x;  // (m, k) row-major
  const uint8_t* w;        // packed panels: [k/64][n rows][128 B swizzled]
  const float* bias;       // (n) or nullptr
  void* y;
  int* err;
  int m, n, k, relu, out_mode, rows_per_batch;
  // split-precision rows (training, csrc/mlp_train.cu): x stores `terms` blocks of kb columns, [h|m] or [h|m|l];
  // the logical K of the contraction is 3 or 6 blocks -- logical block i reads stored block kLinRowBlock[i]
  int terms, kb, ldx;
};

__device__ __constant__ int kLinRowBlock[6] = {0, 1, 0, 2, 0, 1};  // the first three entries serve terms = 2

__global__ void __launch_bounds__(LIN_THREADS) linear_kernel(const LinArgs a) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_w[2], bar_mma[2];
  __shared__ uint32_t tmem_slot;
  __shared__ float s_bias[256];

  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_bytes = 128u * 128u, w_bytes = static_cast<uint32_t>(a.n) * 128u;
  const uint32_t stage_bytes = a_bytes + ((w_bytes + 1023u) & ~1023u);
  const int row0 = blockIdx.x * 128;
  const int nk = a.k / 64;
  uint32_t tmem_cols = 32;
  while (tmem_cols < static_cast<uint32_t>(a.n)) tmem_cols <<= 1;

  if (tid == 0) {
    mbar_init(&bar_w[0], 1); mbar_init(&bar_w[1], 1);
    mbar_init(&bar_mma[0], 1); mbar_init(&bar_mma[1], 1);
    fence_mbar_init_cluster();
  }
  for (int i = tid; i < a.n; i += LIN_THREADS) s_bias[i] = a.bias ? a.bias[i] : 0.f;
  if (warp == 0) tc::tmem_alloc(&tmem_slot, tmem_cols);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_d = tmem_slot;
  const uint32_t idesc = tc::idesc_bf16_f32(128, a.n);

  auto load = [&](int kc, int s) {
    const uint32_t sa = base + s * stage_bytes;
    int col = kc * 64;  // first stored column of the chunk
    if (a.terms > 1) {
      const int cpb = a.kb >> 6, blk = kc / cpb;
      col = kLinRowBlock[blk] * a.kb + (kc - blk * cpb) * 64;
    }
    if (tid == 0) {
      mbar_arrive_expect_tx(&bar_w[s], w_bytes);
      // generic pointer of the stage's W region: offset from the dynamic smem base
      tc::bulk_g2s(smem_raw + (sa + a_bytes - smem_u32(smem_raw)), a.w + static_cast<size_t>(kc) * w_bytes,
                   w_bytes, &bar_w[s]);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int q = i * LIN_THREADS + tid;
      const int r = q >> 3, j = q & 7;
      const bool ok = row0 + r < a.m;
      const __nv_bfloat16* src = a.x + static_cast<size_t>(ok ? row0 + r : 0) * a.ldx + col + j * 8;
      tc::cp_async16(sa + tc::swz_off(r, j), src, ok);
    }
    tc::cp_async_commit();
  };

  bool failed = false;
  load(0, 0);
  for (int kc = 0; kc < nk; ++kc) {
    const int s = kc & 1;
    if (kc + 1 < nk) {
      // stage s^1 was last read by the MMAs of chunk kc-1
      if (kc >= 1 && !mbar_wait(&bar_mma[s ^ 1], ((kc - 1) >> 1) & 1)) { failed = true; break; }
      load(kc + 1, s ^ 1);
      tc::cp_async_wait<1>();
    } else {
      tc::cp_async_wait<0>();
    }
    fence_proxy_async_smem();
    if (!mbar_wait(&bar_w[s], (kc >> 1) & 1)) { failed = true; break; }
    __syncthreads();
    if (tid == 0) {
      tc::fence_after_sync();
      const uint32_t sa = base + s * stage_bytes;
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        tc::mma_bf16(tmem_d, tc::smem_desc_sw128(sa + kk * 32), tc::smem_desc_sw128(sa + a_bytes + kk * 32), idesc,
                     (kc > 0 || kk > 0) ? 1u : 0u);
      }
      tc::mma_commit(&bar_mma[s]);
    }
  }
  if (!failed && !mbar_wait(&bar_mma[(nk - 1) & 1], ((nk - 1) >> 1) & 1)) failed = true;
  if (failed) {
    if (tid == 0) atomicExch(a.err, 2);
  } else {
    tc::fence_after_sync();
    // ---- epilogue: thread t <-> row t (TMEM lane), 16 columns at a time ----------------------
    const int r = row0 + tid;
    const uint32_t lane_base = tmem_d + (static_cast<uint32_t>(warp * 32) << 16);
    int c_begin = 0;
    if (a.out_mode == 0) {
      // fp32 rows, 32 columns at a time through a per-warp 4 KB staging tile in the (now idle) operand ring: a
      // thread holds ONE row, so storing straight from registers makes every warp store 32 separate 16-byte
      // pieces 4 n bytes apart -- the L2 write-request rate, not HBM, bounded the wide training layers (537 MB
      // of Y in 303 us).  Staged, eight lanes write one row's 128 contiguous bytes: 4 full lines per instruction.
      // (All MMAs have completed -- the wait above -- so nothing reads the ring any more.)
      const int lane = tid & 31;
      const uint32_t st = base + static_cast<uint32_t>(warp) * 4096u;
      float* yrow = static_cast<float*>(a.y);
      for (int c0 = 0; c0 + 32 <= a.n; c0 += 32) {
        float v[32];
        tc::tmem_ld16(lane_base + c0, *reinterpret_cast<float(*)[16]>(v));
        tc::tmem_ld16(lane_base + c0 + 16, *reinterpret_cast<float(*)[16]>(v + 16));
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          v[i] += s_bias[c0 + i];
          if (a.relu) v[i] = fmaxf(v[i], 0.f);
        }
#pragma unroll
        for (int ch = 0; ch < 8; ++ch)
          asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(st + lane * 128u + (((ch ^ lane) & 7) << 4)),
                       "f"(v[4 * ch]), "f"(v[4 * ch + 1]), "f"(v[4 * ch + 2]), "f"(v[4 * ch + 3])
                       : "memory");
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int rr = i * 4 + (lane >> 3), ch = lane & 7;
          float4 o;
          asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                       : "=f"(o.x), "=f"(o.y), "=f"(o.z), "=f"(o.w)
                       : "r"(st + rr * 128u + (((ch ^ rr) & 7) << 4))
                       : "memory");
          const int gr = row0 + warp * 32 + rr;
          if (gr < a.m) *reinterpret_cast<float4*>(yrow + static_cast<size_t>(gr) * a.n + c0 + ch * 4) = o;
        }
        __syncwarp();
        c_begin = c0 + 32;
      }
    }
    for (int c0 = c_begin; c0 < a.n; c0 += 16) {
      float v[16];
      tc::tmem_ld16(lane_base + c0, v);
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        v[i] += s_bias[c0 + i];
        if (a.relu) v[i] = fmaxf(v[i], 0.f);
      }
      if (r < a.m) {
        if (a.out_mode == 0) {
          float4* dst = reinterpret_cast<float4*>(static_cast<float*>(a.y) + static_cast<size_t>(r) * a.n + c0);
#pragma unroll
          for (int i = 0; i < 4; ++i) dst[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
        } else if (a.out_mode == 1) {
          uint4* dst = reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(a.y) + static_cast<size_t>(r) * a.n + c0);
          dst[0] = make_uint4(tc::pack_bf16(v[0], v[1]), tc::pack_bf16(v[2], v[3]), tc::pack_bf16(v[4], v[5]),
                              tc::pack_bf16(v[6], v[7]));
          dst[1] = make_uint4(tc::pack_bf16(v[8], v[9]), tc::pack_bf16(v[10], v[11]), tc::pack_bf16(v[12], v[13]),
                              tc::pack_bf16(v[14], v[15]));
        } else {  // channel-major (batch, n, rows_per_batch): consecutive lanes -> consecutive points
          const int b = r / a.rows_per_batch, p = r - b * a.rows_per_batch;
          float* dst = static_cast<float*>(a.y) + (static_cast<size_t>(b) * a.n + c0) * a.rows_per_batch + p;
#pragma unroll
          for (int i = 0; i < 16; ++i) dst[static_cast<size_t>(i) * a.rows_per_batch] = v[i];
        }
      }
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem_d, tmem_cols);
}

// W (n, k) fp32 row-major -> bf16 panels [kpad/64][n][128 B swizzled], zero-padded in K.
__global__ void pack_weight_kernel(int n, int k, int kpad, const float* __restrict__ w, uint8_t* __restrict__ out) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;  // one 16-byte chunk (8 elements) per thread
  const int chunks = kpad / 8;
  if (idx >= n * chunks) return;
  const int r = idx / chunks, ch = idx % chunks;
  const int panel = ch / 8, j = ch % 8;
  uint32_t p[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const int k0 = ch * 8 + 2 * e;
    const float lo = k0 < k ? w[static_cast<size_t>(r) * k + k0] : 0.f;
    const float hi = k0 + 1 < k ? w[static_cast<size_t>(r) * k + k0 + 1] : 0.f;
    p[e] = tc::pack_bf16(lo, hi);
  }
  *reinterpret_cast<uint4*>(out + static_cast<size_t>(panel) * n * 128 + tc::swz_off(r, j)) =
      make_uint4(p[0], p[1], p[2], p[3]);
}

// (B, C, N) fp32 channel-major -> (B*N, kpad) bf16 point-major rows (zero padded), via a smem transpose
__global__ void __launch_bounds__(256)
pack_rows_kernel(int c, int n, int kpad, const float* __restrict__ f, __nv_bfloat16* __restrict__ out) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z, p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  for (int i = ty; i < 32; i += 8) {
    const int ch = c0 + i, p = p0 + tx;
    tile[i][tx] = (ch < c && p < n) ? f[(static_cast<size_t>(b) * c + ch) * n + p] : 0.f;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int p = p0 + i, ch = c0 + tx;
    if (p < n && ch < kpad) out[(static_cast<size_t>(b) * n + p) * kpad + ch] = __float2bfloat16_rn(tile[tx][i]);
  }
}

// The same transpose with a 64-channel x 32-point tile: a thread converts eight consecutive channels of
// one point and writes them with ONE 16-byte store (eight lanes cover a point's 128 contiguous bytes),
// instead of four 2-byte stores: 0.27 instead of 1.3 warp-instructions per element.  kpad % 8 == 0.
__global__ void __launch_bounds__(256)
pack_rows64_kernel(int c, int n, int kpad, const float* __restrict__ f, __nv_bfloat16* __restrict__ out) {
  __shared__ float tile[64][33];
  const int b = blockIdx.z, p0 = blockIdx.x * 32, c0 = blockIdx.y * 64;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const float* src = f + static_cast<size_t>(b) * c * n;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int ch = c0 + ty + 8 * i, p = p0 + tx;
    tile[ty + 8 * i][tx] = (ch < c && p < n) ? __ldg(src + static_cast<size_t>(ch) * n + p) : 0.f;
  }
  __syncthreads();
  const int chunk = threadIdx.x & 7, pl = threadIdx.x >> 3;
  const int p = p0 + pl, ch0 = c0 + chunk * 8;
  if (p < n && ch0 < kpad) {
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = tile[chunk * 8 + i][pl];
    *reinterpret_cast<uint4*>(out + (static_cast<size_t>(b) * n + p) * kpad + ch0) =
        make_uint4(tc::pack_bf16(v[0], v[1]), tc::pack_bf16(v[2], v[3]), tc::pack_bf16(v[4], v[5]), tc::pack_bf16(v[6], v[7]));
  }
}

// pack_rows64_kernel in the split layouts of csrc/mlp_train.cu (store_rows8): terms 2 -> (b*n, 2*kpad) rows
// [h | m]; terms 3 -> (b*n, 3*kpad) rows [h | m | l]
__global__ void __launch_bounds__(256)
pack_rows64_split_kernel(int c, int n, int kpad, int terms, const float* __restrict__ f, __nv_bfloat16* __restrict__ out) {
  __shared__ float tile[64][33];
  const int b = blockIdx.z, p0 = blockIdx.x * 32, c0 = blockIdx.y * 64;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const float* src = f + static_cast<size_t>(b) * c * n;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int ch = c0 + ty + 8 * i, p = p0 + tx;
    tile[ty + 8 * i][tx] = (ch < c && p < n) ? __ldg(src + static_cast<size_t>(ch) * n + p) : 0.f;
  }
  __syncthreads();
  const int chunk = threadIdx.x & 7, pl = threadIdx.x >> 3;
  const int p = p0 + pl, ch0 = c0 + chunk * 8;
  if (p < n && ch0 < kpad) {
    uint32_t h[4], m[4], l[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float v0 = tile[chunk * 8 + 2 * e][pl], v1 = tile[chunk * 8 + 2 * e + 1][pl];
      const __nv_bfloat16 h0 = __float2bfloat16_rn(v0), h1 = __float2bfloat16_rn(v1);
      const float r0 = v0 - __bfloat162float(h0), r1 = v1 - __bfloat162float(h1);
      const __nv_bfloat16 m0 = __float2bfloat16_rn(r0), m1 = __float2bfloat16_rn(r1);
      h[e] = static_cast<uint32_t>(__bfloat16_as_ushort(h0)) | (static_cast<uint32_t>(__bfloat16_as_ushort(h1)) << 16);
      m[e] = static_cast<uint32_t>(__bfloat16_as_ushort(m0)) | (static_cast<uint32_t>(__bfloat16_as_ushort(m1)) << 16);
      l[e] = tc::pack_bf16(r0 - __bfloat162float(m0), r1 - __bfloat162float(m1));
    }
    const uint4 hv = make_uint4(h[0], h[1], h[2], h[3]), mv = make_uint4(m[0], m[1], m[2], m[3]),
                lv = make_uint4(l[0], l[1], l[2], l[3]);
    __nv_bfloat16* row = out + (static_cast<size_t>(b) * n + p) * terms * kpad + ch0;
    *reinterpret_cast<uint4*>(row) = hv;
    *reinterpret_cast<uint4*>(row + kpad) = mv;
    if (terms == 3) *reinterpret_cast<uint4*>(row + 2 * kpad) = lv;
  }
}

// point-major rows of the input cloud, (rows, ld) fp32 with the features in columns [col0, col0+c),
// -> (rows, kpad) bf16 zero padded: thread = 8 output channels (one 16-byte store)
__global__ void __launch_bounds__(256)
pack_point_rows_kernel(long long rows, int ld, int col0, int c, int kpad, const float* __restrict__ in,
                       __nv_bfloat16* __restrict__ out) {
  const int chunks = kpad >> 3;
  const long long t = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
  if (t >= rows * chunks) return;
  const long long r = t / chunks;
  const int ch0 = static_cast<int>(t - r * chunks) << 3;
  const float* src = in + r * ld + col0 + ch0;
  float v[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = ch0 + i < c ? __ldg(src + i) : 0.f;
  *reinterpret_cast<uint4*>(out + r * kpad + ch0) = make_uint4(tc::pack_bf16(v[0], v[1]), tc::pack_bf16(v[2], v[3]),
                                                               tc::pack_bf16(v[4], v[5]), tc::pack_bf16(v[6], v[7]));
}

}  // namespace

int* device_err_flag();  // defined in abi.cu

}  // namespace pn2

using namespace pn2;

extern "C" int pn2_pack_weight_bf16(int n, int k, int kpad, const float* w, void* out, void* stream) {
  PN2_RANGE("pn2_pack_weight_bf16");
  PN2_REQUIRE(n > 0 && k > 0 && kpad >= k && kpad % 64 == 0 && w && out, "pack_weight_bf16: bad arguments");
  const int total = n * (kpad / 8);
  pack_weight_kernel<<<(total + 255) / 256, 256, 0, as_stream(stream)>>>(n, k, kpad, w, static_cast<uint8_t*>(out));
  PN2_CHECK_LAUNCH("pack_weight_kernel");
  return PN2_OK;
}

extern "C" int pn2_pack_rows_bf16(int b, int c, int n, int kpad, const float* features, void* out, void* stream) {
  PN2_RANGE("pn2_pack_rows_bf16");
  PN2_REQUIRE(b > 0 && c > 0 && n > 0 && kpad >= c && features && out && b <= 65535, "pack_rows_bf16: bad arguments");
  if (kpad % 8 == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0) {
    dim3 grid((n + 31) / 32, (kpad + 63) / 64, b);
    pack_rows64_kernel<<<grid, 256, 0, as_stream(stream)>>>(c, n, kpad, features, static_cast<__nv_bfloat16*>(out));
  } else {
    dim3 grid((n + 31) / 32, (kpad + 31) / 32, b);
    pack_rows_kernel<<<grid, 256, 0, as_stream(stream)>>>(c, n, kpad, features, static_cast<__nv_bfloat16*>(out));
  }
  PN2_CHECK_LAUNCH("pack_rows_kernel");
  return PN2_OK;
}

extern "C" int pn2_pack_rows_split_bf16(int b, int c, int n, int kpad, int terms, const float* features, void* out,
                                        void* stream) {
  PN2_RANGE("pn2_pack_rows_split_bf16");
  PN2_REQUIRE(b > 0 && c > 0 && n > 0 && kpad >= c && kpad % 8 == 0 && features && out && b <= 65535 &&
                  (reinterpret_cast<uintptr_t>(out) & 15) == 0 && (terms == 2 || terms == 3),
              "pack_rows_split_bf16: bad arguments");
  dim3 grid((n + 31) / 32, (kpad + 63) / 64, b);
  pack_rows64_split_kernel<<<grid, 256, 0, as_stream(stream)>>>(c, n, kpad, terms, features, static_cast<__nv_bfloat16*>(out));
  PN2_CHECK_LAUNCH("pack_rows64_split_kernel");
  return PN2_OK;
}

extern "C" int pn2_pack_point_rows_bf16(long long rows, int ld, int col0, int c, int kpad, const float* in, void* out,
                                        void* stream) {
  PN2_RANGE("pn2_pack_point_rows_bf16");
  if (rows <= 0) return PN2_OK;
  PN2_REQUIRE(c > 0 && col0 >= 0 && ld >= col0 + c && kpad >= c && kpad % 8 == 0 && in && out,
              "pack_point_rows_bf16: bad arguments");
  const long long total = rows * (kpad / 8);
  pack_point_rows_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, as_stream(stream)>>>(
      rows, ld, col0, c, kpad, in, static_cast<__nv_bfloat16*>(out));
  PN2_CHECK_LAUNCH("pack_point_rows_kernel");
  return PN2_OK;
}

static int linear_launch(int m, int n, int k, int terms, int kb, const void* x_bf16, const void* w_packed,
                         const float* bias, int relu, int out_mode, int rows_per_batch, void* y, void* stream,
                         const char* who) {
  if (m <= 0) return PN2_OK;
  PN2_REQUIRE(n >= 16 && n <= 256 && n % 16 == 0, "%s: n=%d must be a multiple of 16 in [16,256]", who, n);
  PN2_REQUIRE(k >= 64 && k % 64 == 0, "%s: k=%d must be a positive multiple of 64", who, k);
  PN2_REQUIRE(x_bf16 && w_packed && y && out_mode >= 0 && out_mode <= 2, "%s: bad arguments", who);
  PN2_REQUIRE(out_mode != 2 || rows_per_batch > 0, "%s: rows_per_batch required for channel-major output", who);
  LinArgs a;
  a.x = static_cast<const __nv_bfloat16*>(x_bf16); a.w = static_cast<const uint8_t*>(w_packed); a.bias = bias;
  a.y = y; a.m = m; a.n = n; a.k = k; a.relu = relu; a.out_mode = out_mode; a.rows_per_batch = rows_per_batch;
  a.terms = terms; a.kb = kb; a.ldx = terms > 1 ? terms * kb : k;
  a.err = device_err_flag();
  PN2_REQUIRE(a.err != nullptr, "%s: cannot allocate the device error flag", who);
  const size_t stage = 128 * 128 + ((static_cast<size_t>(n) * 128 + 1023) & ~static_cast<size_t>(1023));
  const size_t smem = 2 * stage + 1024;
  PN2_CUDA(cudaFuncSetAttribute(linear_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  linear_kernel<<<(m + 127) / 128, LIN_THREADS, smem, as_stream(stream)>>>(a);
  PN2_CHECK_LAUNCH("linear_kernel");
  return PN2_OK;
}

extern "C" int pn2_linear_bf16(int m, int n, int k, const void* x_bf16, const void* w_packed, const float* bias,
                               int relu, int out_mode, int rows_per_batch, void* y, void* stream) {
  PN2_RANGE("pn2_linear_bf16");
  return linear_launch(m, n, k, 1, k, x_bf16, w_packed, bias, relu, out_mode, rows_per_batch, y, stream, "linear_bf16");
}

extern "C" int pn2_linear_split_bf16(int m, int n, int kb, int terms, const void* x_bf16, const void* w_packed,
                                     const float* bias, int relu, int out_mode, int rows_per_batch, void* y,
                                     void* stream) {
  PN2_RANGE("pn2_linear_split_bf16");
  PN2_REQUIRE(terms >= 1 && terms <= 3 && kb >= 64 && kb % 64 == 0,
              "linear_split_bf16: terms=%d must be 1..3 and kb=%d a positive multiple of 64", terms, kb);
  const int nblk = terms == 1 ? 1 : (terms == 2 ? 3 : 6);
  return linear_launch(m, n, nblk * kb, terms, kb, x_bf16, w_packed, bias, relu, out_mode, rows_per_batch, y, stream,
                       "linear_split_bf16");
}
