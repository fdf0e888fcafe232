/*
 * pointnet2_oracle.c -- CPU restatement of the reference op semantics.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under 3dvlp_b200/ may import, link or call this file; it is
 * the checker used by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs.  The product path (the .cu files under 3dvlp_b200/csrc) fails loudly when its CUDA library is missing and
 * never falls back to this code.
 *
 * Every function restates one kernel of the reference extension
 * (/root/reference/lib/pointnet2/_ext_src/src, the .cu files) with the arithmetic pinned to what nvcc 12.9
 * -O2 emits for sm_100a (SASS inspected with cuobjdump on oracle/_ref/pointnet2_ref_ext.so):
 *
 *     d2  = fma(dz,dz, fma(dx,dx, rn(dy*dy)))      ball_query_gpu.cu:36-37, interpolate_gpu.cu:38,
 *                                                   sampling_gpu.cu:108-109 (and the FPS `mag`)
 *     out = fma(p3,w3, fma(p1,w1, rn(p2*w2)))      interpolate_gpu.cu:103-104
 *
 * fmaf() is exact (single rounding); the file is compiled with -ffp-contract=off so the compiler
 * adds no contractions of its own.
 *
 * Parity status: the reference ships no tests for these ops.  The oracle is
 * pinned (a) on the GPU box against the unmodified reference extension compiled by
 * oracle/build_ref.py (tests/test_ref_ext_gpu.py) and (b) by golden vectors under tests/golden/
 * produced by driving the reference's own pointnet2_utils.py / pointnet2_modules.py.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#if defined(__x86_64__) && defined(__GNUC__) && !defined(__clang__)
#define HOT __attribute__((target_clones("fma", "default")))
#else
#define HOT
#endif

/* cuda_utils.h:20-24 -- block size chosen by the reference launchers (host-side double log). */
int oracle_opt_n_threads(int work_size) {
    const int pow_2 = (int)(log((double)work_size) / log(2.0));
    int t = 1 << pow_2;
    if (t > 512) t = 512;
    if (t < 1) t = 1;
    return t;
}

static inline float sqdist(float ax, float ay, float az, float bx, float by, float bz) {
    const float dx = ax - bx, dy = ay - by, dz = az - bz;
    float t = dy * dy;
    t = fmaf(dx, dx, t);
    return fmaf(dz, dz, t);
}

/*
 * Furthest point sampling -- sampling_gpu.cu:75-178 (launcher :180-234, caller sampling.cpp:70-91).
 * dataset (b,n,3) f32, temp (b,n) f32 scratch pre-filled with 1e10 by the caller, idxs (b,m) i32.
 * The block of `bs` threads is simulated literally: thread t owns k = t, t+bs, ...; per-thread
 * strict-greater arg-max; then the shared-memory halving tree with `v2 > v1 ? i2 : i1`.
 */
HOT void oracle_furthest_point_sampling(int b, int n, int m, const float *dataset, float *temp,
                                        int *idxs) {
    if (m <= 0) return;
    const int bs = oracle_opt_n_threads(n);
    for (int bi = 0; bi < b; ++bi) {
        const float *xyz = dataset + (size_t)bi * n * 3;
        float *tmp = temp + (size_t)bi * n;
        int *out = idxs + (size_t)bi * m;
        float *dists = (float *)malloc(sizeof(float) * bs);
        int *dists_i = (int *)malloc(sizeof(int) * bs);
        unsigned char *skip = (unsigned char *)malloc(n);
        for (int k = 0; k < n; ++k) {
            const float x = xyz[3 * k], y = xyz[3 * k + 1], z = xyz[3 * k + 2];
            float mag = y * y;
            mag = fmaf(x, x, mag);
            mag = fmaf(z, z, mag);
            skip[k] = ((double)mag <= 1e-3) ? 1 : 0; /* sampling_gpu.cu:105-106 */
        }
        int old = 0;
        out[0] = 0;
        for (int j = 1; j < m; ++j) {
            const float x1 = xyz[3 * old], y1 = xyz[3 * old + 1], z1 = xyz[3 * old + 2];
            for (int t = 0; t < bs; ++t) { dists[t] = -1.0f; dists_i[t] = 0; }
            for (int k = 0; k < n; ++k) {
                if (skip[k]) continue;
                const float d = sqdist(xyz[3 * k], xyz[3 * k + 1], xyz[3 * k + 2], x1, y1, z1);
                const float d2 = fminf(d, tmp[k]);
                tmp[k] = d2;
                const int t = k % bs;
                if (d2 > dists[t]) { dists[t] = d2; dists_i[t] = k; }
            }
            for (int s = bs >> 1; s >= 1; s >>= 1) {
                for (int t = 0; t < s; ++t) {
                    const float v1 = dists[t], v2 = dists[t + s];
                    const int i1 = dists_i[t], i2 = dists_i[t + s];
                    dists[t] = fmaxf(v1, v2); /* CUDA max(float,float) */
                    dists_i[t] = v2 > v1 ? i2 : i1;
                }
            }
            old = dists_i[0];
            out[j] = old;
        }
        free(dists); free(dists_i); free(skip);
    }
}

/*
 * The same recurrence with the selection restated as ONE associative max over packed keys
 * (SURVEY.md appendix A-1): thread t of the reference block owns k = t (mod bs) and keeps its
 * smallest maximal k; the halving tree prefers, among equal values, the thread with the smallest
 * bit-reversed id.  Hence winner = argmax_k (temp[k], -key32(k)) with
 *     key32(k) = bitrev_{log2 bs}(k mod bs) * 2^(32-log2 bs) ... packed as (rev << S) | (k / bs).
 * temp >= 0 so its fp32 bit pattern orders like an unsigned integer.  No eligible point -> 0.
 * This is the formulation the CUDA kernel (3dvlp_b200/csrc/fps.cu) implements; tests check it
 * against the literal simulation above on ties, duplicates, lattices and near-origin points.
 */
static inline uint32_t bitrev(uint32_t v, int bits) {
    uint32_t r = 0;
    for (int i = 0; i < bits; ++i) r |= ((v >> i) & 1u) << (bits - 1 - i);
    return r;
}

HOT void oracle_furthest_point_sampling_keyed(int b, int n, int m, const float *dataset,
                                              float *temp, int *idxs) {
    if (m <= 0) return;
    const int bs = oracle_opt_n_threads(n);
    int bits = 0;
    while ((1 << bits) < bs) ++bits;
    const int shift = 32 - bits; /* k / bs < 2^shift as long as n < 2^31 / ... (n <= 2^22 * bs) */
    for (int bi = 0; bi < b; ++bi) {
        const float *xyz = dataset + (size_t)bi * n * 3;
        float *tmp = temp + (size_t)bi * n;
        int *out = idxs + (size_t)bi * m;
        int old = 0;
        out[0] = 0;
        for (int j = 1; j < m; ++j) {
            const float x1 = xyz[3 * old], y1 = xyz[3 * old + 1], z1 = xyz[3 * old + 2];
            uint64_t best = 0; int besti = 0; int any = 0;
            for (int k = 0; k < n; ++k) {
                const float x = xyz[3 * k], y = xyz[3 * k + 1], z = xyz[3 * k + 2];
                float mag = y * y;
                mag = fmaf(x, x, mag);
                mag = fmaf(z, z, mag);
                if ((double)mag <= 1e-3) continue;
                const float d2 = fminf(sqdist(x, y, z, x1, y1, z1), tmp[k]);
                tmp[k] = d2;
                uint32_t vb; memcpy(&vb, &d2, 4);
                const uint32_t key32 = bits ? ((bitrev((uint32_t)(k % bs), bits) << shift) | (uint32_t)(k / bs))
                                            : (uint32_t)k;
                const uint64_t packed = ((uint64_t)vb << 32) | (uint64_t)(0xFFFFFFFFu - key32);
                if (!any || packed > best) { best = packed; besti = k; any = 1; }
            }
            old = any ? besti : 0;
            out[j] = old;
        }
    }
}

/* gather_points -- sampling_gpu.cu:13-25.  points (b,c,n), idx (b,m) -> out (b,c,m). */
void oracle_gather_points(int b, int c, int n, int m, const float *points, const int *idx,
                          float *out) {
    for (int i = 0; i < b; ++i)
        for (int l = 0; l < c; ++l)
            for (int j = 0; j < m; ++j)
                out[((size_t)i * c + l) * m + j] = points[((size_t)i * c + l) * n + idx[(size_t)i * m + j]];
}

/* gather_points_grad -- sampling_gpu.cu:39-52.  grad_points must be zero-filled by the caller. */
void oracle_gather_points_grad(int b, int c, int n, int m, const float *grad_out, const int *idx,
                               float *grad_points) {
    for (int i = 0; i < b; ++i)
        for (int l = 0; l < c; ++l)
            for (int j = 0; j < m; ++j)
                grad_points[((size_t)i * c + l) * n + idx[(size_t)i * m + j]] +=
                    grad_out[((size_t)i * c + l) * m + j];
}

/*
 * ball_query -- ball_query_gpu.cu:14-49.  new_xyz (b,m,3), xyz (b,n,3) -> idx (b,m,nsample) i32.
 * idx must be zero-filled by the caller (an empty ball leaves its row untouched, ball_query.cpp:24).
 */
HOT void oracle_ball_query(int b, int n, int m, float radius, int nsample, const float *new_xyz,
                           const float *xyz, int *idx) {
    const float radius2 = radius * radius;
    for (int bi = 0; bi < b; ++bi) {
        for (int j = 0; j < m; ++j) {
            const float *P = xyz + (size_t)bi * n * 3;
            const float *c = new_xyz + ((size_t)bi * m + j) * 3;
            int *row = idx + ((size_t)bi * m + j) * nsample;
            int cnt = 0;
            for (int k = 0; k < n && cnt < nsample; ++k) {
                const float d2 = sqdist(c[0], c[1], c[2], P[3 * k], P[3 * k + 1], P[3 * k + 2]);
                if (d2 < radius2) {
                    if (cnt == 0)
                        for (int l = 0; l < nsample; ++l) row[l] = k;
                    row[cnt] = k;
                    ++cnt;
                }
            }
        }
    }
}

/* group_points -- group_points_gpu.cu:13-33.  points (b,c,n), idx (b,np,ns) -> out (b,c,np,ns). */
void oracle_group_points(int b, int c, int n, int npoints, int nsample, const float *points,
                         const int *idx, float *out) {
    for (int i = 0; i < b; ++i)
        for (int l = 0; l < c; ++l) {
            const float *src = points + ((size_t)i * c + l) * n;
            const int *id = idx + (size_t)i * npoints * nsample;
            float *dst = out + ((size_t)i * c + l) * npoints * nsample;
            for (size_t q = 0; q < (size_t)npoints * nsample; ++q) dst[q] = src[id[q]];
        }
}

/* group_points_grad -- group_points_gpu.cu:48-69.  grad_points zero-filled by the caller. */
void oracle_group_points_grad(int b, int c, int n, int npoints, int nsample, const float *grad_out,
                              const int *idx, float *grad_points) {
    for (int i = 0; i < b; ++i)
        for (int l = 0; l < c; ++l) {
            float *dst = grad_points + ((size_t)i * c + l) * n;
            const int *id = idx + (size_t)i * npoints * nsample;
            const float *src = grad_out + ((size_t)i * c + l) * npoints * nsample;
            for (size_t q = 0; q < (size_t)npoints * nsample; ++q) dst[id[q]] += src[q];
        }
}

/*
 * three_nn -- interpolate_gpu.cu:14-64.  unknown (b,n,3), known (b,m,3) -> dist2 (b,n,3) f32,
 * idx (b,n,3) i32.  best* are doubles initialised to 1e40 (so an unfilled slot stores +inf).
 */
HOT void oracle_three_nn(int b, int n, int m, const float *unknown, const float *known,
                         float *dist2, int *idx) {
    for (int bi = 0; bi < b; ++bi) {
        for (int j = 0; j < n; ++j) {
            const float *u = unknown + ((size_t)bi * n + j) * 3;
            const float *K = known + (size_t)bi * m * 3;
            double best1 = 1e40, best2 = 1e40, best3 = 1e40;
            int i1 = 0, i2 = 0, i3 = 0;
            for (int k = 0; k < m; ++k) {
                const double d = (double)sqdist(u[0], u[1], u[2], K[3 * k], K[3 * k + 1], K[3 * k + 2]);
                if (d < best1) {
                    best3 = best2; i3 = i2; best2 = best1; i2 = i1; best1 = d; i1 = k;
                } else if (d < best2) {
                    best3 = best2; i3 = i2; best2 = d; i2 = k;
                } else if (d < best3) {
                    best3 = d; i3 = k;
                }
            }
            float *o = dist2 + ((size_t)bi * n + j) * 3;
            int *oi = idx + ((size_t)bi * n + j) * 3;
            o[0] = (float)best1; o[1] = (float)best2; o[2] = (float)best3;
            oi[0] = i1; oi[1] = i2; oi[2] = i3;
        }
    }
}

/* three_interpolate -- interpolate_gpu.cu:77-106.  points (b,c,m), idx/weight (b,n,3) -> (b,c,n). */
HOT void oracle_three_interpolate(int b, int c, int m, int n, const float *points, const int *idx,
                                  const float *weight, float *out) {
    for (int bi = 0; bi < b; ++bi)
        for (int l = 0; l < c; ++l) {
            const float *p = points + ((size_t)bi * c + l) * m;
            for (int j = 0; j < n; ++j) {
                const float *w = weight + ((size_t)bi * n + j) * 3;
                const int *id = idx + ((size_t)bi * n + j) * 3;
                float t = p[id[1]] * w[1];
                t = fmaf(p[id[0]], w[0], t);
                out[((size_t)bi * c + l) * n + j] = fmaf(p[id[2]], w[2], t);
            }
        }
}

/*
 * three_interpolate_grad -- the mathematically correct scatter-add of interpolate_gpu.cu:121-148.
 * (The reference binding interpolate.cpp:95 launches the FORWARD kernel by mistake -- SURVEY.md
 * appendix C-1; this oracle and the product implement the intended gradient.)
 */
void oracle_three_interpolate_grad(int b, int c, int n, int m, const float *grad_out,
                                   const int *idx, const float *weight, float *grad_points) {
    for (int bi = 0; bi < b; ++bi)
        for (int l = 0; l < c; ++l) {
            float *g = grad_points + ((size_t)bi * c + l) * m;
            for (int j = 0; j < n; ++j) {
                const float *w = weight + ((size_t)bi * n + j) * 3;
                const int *id = idx + ((size_t)bi * n + j) * 3;
                const float go = grad_out[((size_t)bi * c + l) * n + j];
                g[id[0]] += go * w[0];
                g[id[1]] += go * w[1];
                g[id[2]] += go * w[2];
            }
        }
}

/*
 * nn_distance -- utils/nn_distance.py:32-59 (huber_loss :13-30).  pc1 (b,n,c), pc2 (b,m,c).
 * mode 0: sum diff^2, 1: sum |diff| (l1=True), 2: smooth-L1 / Huber with `delta` (l1smooth=True).
 * The per-pair cost is accumulated over c left to right in fp32 (torch.sum over a short last
 * dimension); ties in the two mins keep the first index (torch.min).  Indices are int64.
 */
void oracle_nn_distance(int b, int n, int m, int c, const float *pc1, const float *pc2, int mode,
                        float delta, float *dist1, int64_t *idx1, float *dist2, int64_t *idx2) {
    for (int bi = 0; bi < b; ++bi) {
        const float *A = pc1 + (size_t)bi * n * c, *B = pc2 + (size_t)bi * m * c;
        float *d1 = dist1 + (size_t)bi * n, *d2 = dist2 + (size_t)bi * m;
        int64_t *i1 = idx1 + (size_t)bi * n, *i2 = idx2 + (size_t)bi * m;
        for (int k = 0; k < m; ++k) { d2[k] = INFINITY; i2[k] = 0; }
        for (int j = 0; j < n; ++j) {
            float bestd = INFINITY; int64_t besti = 0;
            for (int k = 0; k < m; ++k) {
                float s = 0.0f;
                for (int q = 0; q < c; ++q) {
                    const float e = A[(size_t)j * c + q] - B[(size_t)k * c + q];
                    float v;
                    if (mode == 0) v = e * e;
                    else if (mode == 1) v = fabsf(e);
                    else {
                        const float a = fabsf(e);
                        const float quad = a < delta ? a : delta; /* clamp(max=delta) */
                        const float lin = a - quad;
                        v = 0.5f * (quad * quad) + delta * lin;
                    }
                    s += v;
                }
                if (k == 0 || s < bestd) { bestd = s; besti = k; }
                if (j == 0 || s < d2[k]) { d2[k] = s; i2[k] = j; }
            }
            d1[j] = bestd; i1[j] = besti;
        }
    }
}
