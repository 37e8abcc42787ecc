// ALiBi Transformer-MIL *training* step on one GPU: forward that checkpoints what the backward
// needs, backward (all parameter gradients, optionally d bags), soft-target cross entropy, the
// training-mode running-mean statistic, and a fused AdamW over a flat parameter buffer.
//
// replaces: LitTileClassifier._step / training_step (src/stamp/modeling/models/__init__.py:239-286:
//   logits = model(bags, coords=coords, mask=None); cross_entropy(logits, targets, weight=class_weights))
//   + loss.backward() through VisionTransformer.forward (vision_tranformer.py:332-384 and everything
//   it calls, incl. nn.Dropout in project_features :314-318 and feed_forward :157-169 and the
//   training-mode update of _RunningMeanScaler :23-31) + optim.AdamW.step (models/__init__.py:133-141).
//
// Numerics: bf16 tensor-core operands, fp32 accumulation, fp32 residual stream, fp32 master
// parameters and gradients (BASELINE.json configs[3] "training bf16").
//
// Checkpointed activations (ctx; M = B*(N+1) token rows, d = dim_model):
//   bags16 [B*N,F] bf16   z0 [B*N,d] f32 (pre-GELU)   coords_s [M] float2
//   x[0..2L] [M,d] f32 residual stream before each LayerNorm / after the last block
//   per layer: xn1, att, xn2 [M,d] bf16; qkv [M,3d] bf16; osm, odv [M,d] f32; lse2 [B,H,S] f32;
//              z1 [M,ff] f32 (pre-GELU); h [M,ff] bf16
// Dropout masks are never stored: keep(i) = hash(seed, site, i) >= p * 2^32, regenerated in backward.
#include <math.h>

#include "attention.cuh"
#include "attention_train.cuh"
#include "common.cuh"
#include "gemm.cuh"
#include "mil_common.cuh"
#include "rowops.cuh"
#include "stamp_b200.h"
#include "wgrad_tc.cuh"

namespace sb {
namespace {

// ---------------------------------------------------------------------------------------------
// small helpers
// ---------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint64_t splitmix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__device__ __forceinline__ bool keep_elem(uint64_t site_seed, uint64_t idx, uint32_t thresh) {
    return static_cast<uint32_t>(splitmix64(site_seed ^ (idx * 0xD6E8FEB86659FD93ull)) >> 32) >= thresh;
}
inline uint64_t site_seed(uint64_t seed, int site) { return splitmix64(seed + 0x632BE59BD9B4E019ull * static_cast<uint64_t>(site + 1)); }
inline uint32_t drop_thresh(float p) {
    if (!(p > 0.f)) return 0u;
    const double t = static_cast<double>(p) * 4294967296.0;
    return t >= 4294967295.0 ? 4294967295u : static_cast<uint32_t>(t);
}
inline float inv_keep(float p) { return p > 0.f ? 1.0f / (1.0f - p) : 1.0f; }

__device__ __forceinline__ float gelu_exact(float z) { return 0.5f * z * (1.0f + erff(z * 0.70710678118654752f)); }
__device__ __forceinline__ float gelu_grad(float z) {
    return 0.5f * (1.0f + erff(z * 0.70710678118654752f)) + z * 0.3989422804014327f * __expf(-0.5f * z * z);
}
__device__ __forceinline__ float bf16_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }

inline int grid1d(long long total, int block) {
    long long g = (total + block - 1) / block;
    const long long cap = 148LL * 16;
    return static_cast<int>(g < cap ? (g > 0 ? g : 1) : cap);
}

__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 "
        "{%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// ---------------------------------------------------------------------------------------------
// element-wise / row kernels (HBM-bound)
// ---------------------------------------------------------------------------------------------
// fp32 -> bf16, optional dropout mask + scale (thresh 0: plain cast). n % 4 == 0.
__global__ void __launch_bounds__(256)
mask_cast_kernel(const float* __restrict__ in, uint16_t* __restrict__ out, long long n4, uint32_t thresh,
                 float scale, uint64_t sseed) {
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n4;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        float4 v = *reinterpret_cast<const float4*>(in + i * 4);
        if (thresh != 0u) {
            v.x = keep_elem(sseed, i * 4 + 0, thresh) ? v.x * scale : 0.f;
            v.y = keep_elem(sseed, i * 4 + 1, thresh) ? v.y * scale : 0.f;
            v.z = keep_elem(sseed, i * 4 + 2, thresh) ? v.z * scale : 0.f;
            v.w = keep_elem(sseed, i * 4 + 3, thresh) ? v.w * scale : 0.f;
        }
        uint2 pk;
        pk.x = pack_bf16(v.x, v.y);
        pk.y = pack_bf16(v.z, v.w);
        *reinterpret_cast<uint2*>(out + i * 4) = pk;
    }
}

// out = dropout(GELU(z)); z fp32 [R, C] compact.  out32 (fp32, row-remapped: row = (r/gin)*gout+goff+r%gin)
// or out16 (bf16, compact).  Element index for the mask = r*C + c.
__global__ void __launch_bounds__(256)
gelu_drop_fwd_kernel(const float* __restrict__ z, long long R, int C, float* __restrict__ out32,
                     uint16_t* __restrict__ out16, int gin, int gout, int goff, uint32_t thresh, float scale,
                     uint64_t sseed) {
    const int c4n = C >> 2;
    const long long total = R * c4n;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long r = i / c4n;
        const int c = static_cast<int>(i % c4n) * 4;
        const float4 v = *reinterpret_cast<const float4*>(z + r * C + c);
        float y[4] = {gelu_exact(v.x), gelu_exact(v.y), gelu_exact(v.z), gelu_exact(v.w)};
        if (thresh != 0u) {
#pragma unroll
            for (int e = 0; e < 4; ++e) y[e] = keep_elem(sseed, r * C + c + e, thresh) ? y[e] * scale : 0.f;
        }
        if (out32 != nullptr) {
            const long long row = gin > 0 ? (r / gin) * gout + goff + (r % gin) : r;
            *reinterpret_cast<float4*>(out32 + row * C + c) = make_float4(y[0], y[1], y[2], y[3]);
        } else {
            uint2 pk;
            pk.x = pack_bf16(y[0], y[1]);
            pk.y = pack_bf16(y[2], y[3]);
            *reinterpret_cast<uint2*>(out16 + r * C + c) = pk;
        }
    }
}

// dz = dy * keep * scale * GELU'(z) -> bf16 compact [R, C]; dy fp32, optionally row-remapped
__global__ void __launch_bounds__(256)
gelu_drop_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ z, long long R, int C,
                     uint16_t* __restrict__ dz, int gin, int gout, int goff, uint32_t thresh, float scale,
                     uint64_t sseed) {
    const int c4n = C >> 2;
    const long long total = R * c4n;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long r = i / c4n;
        const int c = static_cast<int>(i % c4n) * 4;
        const long long row = gin > 0 ? (r / gin) * gout + goff + (r % gin) : r;
        const float4 g = *reinterpret_cast<const float4*>(dy + row * C + c);
        const float4 v = *reinterpret_cast<const float4*>(z + r * C + c);
        float y[4] = {g.x * gelu_grad(v.x), g.y * gelu_grad(v.y), g.z * gelu_grad(v.z), g.w * gelu_grad(v.w)};
        if (thresh != 0u) {
#pragma unroll
            for (int e = 0; e < 4; ++e) y[e] = keep_elem(sseed, r * C + c + e, thresh) ? y[e] * scale : 0.f;
        }
        uint2 pk;
        pk.x = pack_bf16(y[0], y[1]);
        pk.y = pack_bf16(y[2], y[3]);
        *reinterpret_cast<uint2*>(dz + r * C + c) = pk;
    }
}

// xout = xin + dropout(y)     (feed_forward's trailing Dropout + the residual add, :169, :292)
__global__ void __launch_bounds__(256)
resid_drop_fwd_kernel(const float* __restrict__ xin, const float* __restrict__ y, float* __restrict__ xout,
                      long long n4, uint32_t thresh, float scale, uint64_t sseed) {
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n4;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const float4 a = *reinterpret_cast<const float4*>(xin + i * 4);
        float4 v = *reinterpret_cast<const float4*>(y + i * 4);
        if (thresh != 0u) {
            v.x = keep_elem(sseed, i * 4 + 0, thresh) ? v.x * scale : 0.f;
            v.y = keep_elem(sseed, i * 4 + 1, thresh) ? v.y * scale : 0.f;
            v.z = keep_elem(sseed, i * 4 + 2, thresh) ? v.z * scale : 0.f;
            v.w = keep_elem(sseed, i * 4 + 3, thresh) ? v.w * scale : 0.f;
        }
        *reinterpret_cast<float4*>(xout + i * 4) = make_float4(a.x + v.x, a.y + v.y, a.z + v.z, a.w + v.w);
    }
}

__global__ void __launch_bounds__(256)
dropout_mask_kernel(uint8_t* __restrict__ out, long long n, uint32_t thresh, uint64_t sseed) {
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * blockDim.x)
        out[i] = keep_elem(sseed, i, thresh) ? 1 : 0;
}

// LayerNorm backward, one warp per row:  dx += rstd * (g - mean(g) - xhat * mean(g * xhat)),  g = dy * gamma;
// dgamma += sum_rows dy * xhat,  dbeta += sum_rows dy  (register partials per lane, smem reduce, atomics).
// out16 (optional): the updated dx as bf16 through a dropout mask (thresh 0: plain cast) -- the operand of the next
// weight-gradient / data-gradient GEMMs, which would otherwise be a separate pass over dx (mask_cast_kernel).
template <int MAXV>
__global__ void __launch_bounds__(256)
ln_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ gamma,
              float* __restrict__ dx, float* __restrict__ dgamma, float* __restrict__ dbeta, int rows, int d,
              float eps, uint16_t* __restrict__ out16, uint32_t thresh, float mscale, uint64_t sseed) {
    extern __shared__ float sh[];   // 2 * d
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int nvec = d >> 2;
    float4 ag[MAXV], ab[MAXV];
#pragma unroll
    for (int i = 0; i < MAXV; ++i) { ag[i] = make_float4(0.f, 0.f, 0.f, 0.f); ab[i] = ag[i]; }
    for (int i = threadIdx.x; i < 2 * d; i += blockDim.x) sh[i] = 0.f;
    __syncthreads();
    for (long long row = static_cast<long long>(blockIdx.x) * nwarps + warp; row < rows;
         row += static_cast<long long>(gridDim.x) * nwarps) {
        const float* xr = x + row * d;
        const float* gr = dy + row * d;
        float4 v[MAXV], g[MAXV];
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < MAXV; ++i) {
            const int idx = lane + i * 32;
            if (idx < nvec) {
                v[i] = *reinterpret_cast<const float4*>(xr + idx * 4);
                g[i] = *reinterpret_cast<const float4*>(gr + idx * 4);
                s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
            }
        }
        s = warp_sum(s);
        const float mean = s / static_cast<float>(d);
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < MAXV; ++i) {
            const int idx = lane + i * 32;
            if (idx < nvec) {
                v[i].x -= mean; v[i].y -= mean; v[i].z -= mean; v[i].w -= mean;
                q += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
            }
        }
        q = warp_sum(q);
        const float rstd = rsqrtf(q / static_cast<float>(d) + eps);
        float m1 = 0.f, m2 = 0.f;
#pragma unroll
        for (int i = 0; i < MAXV; ++i) {
            const int idx = lane + i * 32;
            if (idx < nvec) {
                const float4 w = __ldg(reinterpret_cast<const float4*>(gamma) + idx);
                v[i].x *= rstd; v[i].y *= rstd; v[i].z *= rstd; v[i].w *= rstd;   // xhat
                ag[i].x = fmaf(g[i].x, v[i].x, ag[i].x); ag[i].y = fmaf(g[i].y, v[i].y, ag[i].y);
                ag[i].z = fmaf(g[i].z, v[i].z, ag[i].z); ag[i].w = fmaf(g[i].w, v[i].w, ag[i].w);
                ab[i].x += g[i].x; ab[i].y += g[i].y; ab[i].z += g[i].z; ab[i].w += g[i].w;
                g[i].x *= w.x; g[i].y *= w.y; g[i].z *= w.z; g[i].w *= w.w;       // g = dy * gamma
                m1 += (g[i].x + g[i].y) + (g[i].z + g[i].w);
                m2 += (g[i].x * v[i].x + g[i].y * v[i].y) + (g[i].z * v[i].z + g[i].w * v[i].w);
            }
        }
        m1 = warp_sum(m1) / static_cast<float>(d);
        m2 = warp_sum(m2) / static_cast<float>(d);
        float* dr = dx + row * d;
#pragma unroll
        for (int i = 0; i < MAXV; ++i) {
            const int idx = lane + i * 32;
            if (idx < nvec) {
                float4 o = *reinterpret_cast<float4*>(dr + idx * 4);
                o.x += rstd * (g[i].x - m1 - v[i].x * m2);
                o.y += rstd * (g[i].y - m1 - v[i].y * m2);
                o.z += rstd * (g[i].z - m1 - v[i].z * m2);
                o.w += rstd * (g[i].w - m1 - v[i].w * m2);
                *reinterpret_cast<float4*>(dr + idx * 4) = o;
                if (out16 != nullptr) {
                    const long long e0 = row * d + idx * 4;     // element index of the mask = position in [M, d]
                    if (thresh != 0u) {
                        o.x = keep_elem(sseed, e0 + 0, thresh) ? o.x * mscale : 0.f;
                        o.y = keep_elem(sseed, e0 + 1, thresh) ? o.y * mscale : 0.f;
                        o.z = keep_elem(sseed, e0 + 2, thresh) ? o.z * mscale : 0.f;
                        o.w = keep_elem(sseed, e0 + 3, thresh) ? o.w * mscale : 0.f;
                    }
                    *reinterpret_cast<uint2*>(out16 + e0) = make_uint2(pack_bf16(o.x, o.y), pack_bf16(o.z, o.w));
                }
            }
        }
    }
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
        const int idx = lane + i * 32;
        if (idx < nvec) {
            atomicAdd(sh + idx * 4 + 0, ag[i].x); atomicAdd(sh + idx * 4 + 1, ag[i].y);
            atomicAdd(sh + idx * 4 + 2, ag[i].z); atomicAdd(sh + idx * 4 + 3, ag[i].w);
            atomicAdd(sh + d + idx * 4 + 0, ab[i].x); atomicAdd(sh + d + idx * 4 + 1, ab[i].y);
            atomicAdd(sh + d + idx * 4 + 2, ab[i].z); atomicAdd(sh + d + idx * 4 + 3, ab[i].w);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < d; i += blockDim.x) {
        atomicAdd(dgamma + i, sh[i]);
        atomicAdd(dbeta + i, sh[d + i]);
    }
}

// out[c] += sum_r a[r * ld + c]; bf16 or fp32 rows.  block (32, 8): 2 columns per thread.
template <bool BF16IN>
__global__ void __launch_bounds__(256)
colsum_kernel(const void* __restrict__ a, long long ld, long long rows, int cols, float* __restrict__ out) {
    __shared__ float red[8][64];
    const int c = blockIdx.x * 64 + threadIdx.x * 2;
    float s0 = 0.f, s1 = 0.f;
    if (c < cols) {
        for (long long r = static_cast<long long>(blockIdx.y) * 8 + threadIdx.y; r < rows;
             r += static_cast<long long>(gridDim.y) * 8) {
            if constexpr (BF16IN) {
                const uint32_t w = *reinterpret_cast<const uint32_t*>(static_cast<const uint16_t*>(a) + r * ld + c);
                s0 += bf16_lo(w); s1 += bf16_hi(w);
            } else {
                const float2 w = *reinterpret_cast<const float2*>(static_cast<const float*>(a) + r * ld + c);
                s0 += w.x; s1 += w.y;
            }
        }
    }
    red[threadIdx.y][threadIdx.x * 2] = s0;
    red[threadIdx.y][threadIdx.x * 2 + 1] = s1;
    __syncthreads();
    if (threadIdx.y == 0 && c < cols) {
#pragma unroll
        for (int i = 1; i < 8; ++i) { s0 += red[i][threadIdx.x * 2]; s1 += red[i][threadIdx.x * 2 + 1]; }
        atomicAdd(out + c, s0);
        atomicAdd(out + c + 1, s1);
    }
}

// W fp32 [R, C] -> bf16 copy [R, C] and bf16 transpose [C, R] (either may be null)
__global__ void __launch_bounds__(256)
cast_transpose_kernel(const float* __restrict__ w, int R, int C, uint16_t* __restrict__ w16,
                      uint16_t* __restrict__ w16t) {
    __shared__ float tile[32][33];
    const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += 8) {
        const int r = r0 + i, c = c0 + threadIdx.x;
        float v = 0.f;
        if (r < R && c < C) {
            v = w[static_cast<long long>(r) * C + c];
            if (w16 != nullptr) w16[static_cast<long long>(r) * C + c] = __bfloat16_as_ushort(__float2bfloat16_rn(v));
        }
        tile[i][threadIdx.x] = v;
    }
    __syncthreads();
    if (w16t != nullptr)
        for (int i = threadIdx.y; i < 32; i += 8) {
            const int c = c0 + i, r = r0 + threadIdx.x;
            if (r < R && c < C)
                w16t[static_cast<long long>(c) * R + r] = __bfloat16_as_ushort(__float2bfloat16_rn(tile[threadIdx.x][i]));
        }
}

// sum over bags of all pairwise token distances (class token at (0,0) included): the statistic the
// training-mode _RunningMeanScaler consumes (vision_tranformer.py:23-31: mean over the [B,S,S] cdist).
// grid (query blocks of 256, bags, key splits): the keys of a split are staged 256 at a time in shared memory (every
// lane reads the same key: broadcast), four independent fp32 partial sums per thread, sqrt.approx (2^-22 relative:
// the statistic is a mean over S^2 terms), fp64 from the 256-key partials upwards.
constexpr int DIST_SPLITS = 8;
__device__ __forceinline__ float sqrt_approx(float x) {
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;\n" : "=f"(r) : "f"(x));
    return r;
}
__global__ void __launch_bounds__(256)
dist_sum_kernel(const float2* __restrict__ coords_s, int S, double* __restrict__ out) {
    __shared__ float2 ck[256];
    __shared__ double red[8];
    const int b = blockIdx.y;
    const float2* c = coords_s + static_cast<long long>(b) * S;
    const int q = blockIdx.x * 256 + threadIdx.x;
    const float2 cq = (q < S) ? __ldg(c + q) : make_float2(0.f, 0.f);
    const int per = ((S + DIST_SPLITS - 1) / DIST_SPLITS + 255) / 256 * 256;
    const int k_begin = blockIdx.z * per, k_end = min(S, k_begin + per);
    double tot = 0.0;
    for (int k0 = k_begin; k0 < k_end; k0 += 256) {
        __syncthreads();
        const int k = k0 + threadIdx.x;
        ck[threadIdx.x] = (k < k_end) ? __ldg(c + k) : make_float2(0.f, 0.f);
        __syncthreads();
        const int n = min(256, k_end - k0);
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
        for (int j = n & ~3; j < n; ++j) {
            const float dx = cq.x - ck[j].x, dy = cq.y - ck[j].y;
            s0 += sqrt_approx(fmaf(dx, dx, dy * dy));
        }
#pragma unroll 4
        for (int j = 0; j < (n & ~3); j += 4) {
            const float2 a = ck[j], a1 = ck[j + 1], a2 = ck[j + 2], a3 = ck[j + 3];
            float dx = cq.x - a.x, dy = cq.y - a.y;
            s0 += sqrt_approx(fmaf(dx, dx, dy * dy));
            dx = cq.x - a1.x; dy = cq.y - a1.y;
            s1 += sqrt_approx(fmaf(dx, dx, dy * dy));
            dx = cq.x - a2.x; dy = cq.y - a2.y;
            s2 += sqrt_approx(fmaf(dx, dx, dy * dy));
            dx = cq.x - a3.x; dy = cq.y - a3.y;
            s3 += sqrt_approx(fmaf(dx, dx, dy * dy));
        }
        if (q < S) tot += static_cast<double>((s0 + s1) + (s2 + s3));
    }
    tot = warp_sum_d(tot);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = tot;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int i = 0; i < 8; ++i) t += red[i];
        atomicAdd(out, t);
    }
}

__global__ void dist_mean_finish_kernel(const double* __restrict__ sum, double count, float* __restrict__ mean) {
    *mean = static_cast<float>(*sum / count);
}

// backward of logits = head(LayerNorm(x[b*S, :])): one CTA per bag; writes dx of the class-token row
// (the rest of dx must be zero), accumulates d head / d norm with atomics.
__global__ void __launch_bounds__(256)
cls_head_bwd_kernel(const float* __restrict__ x, long long bag_stride, int d, const float* __restrict__ nw,
                    const float* __restrict__ nb, const float* __restrict__ hw, int C, float eps,
                    const float* __restrict__ dlogits, float* __restrict__ dx, float* __restrict__ dnw,
                    float* __restrict__ dnb, float* __restrict__ dhw, float* __restrict__ dhb) {
    extern __shared__ float sh[];   // xhat[d], g[d], red[32]
    float* xh = sh;
    float* gb = sh + d;
    float* red = sh + 2 * d;
    const float* xr = x + blockIdx.x * bag_stride;
    const float* dl = dlogits + static_cast<long long>(blockIdx.x) * C;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw_ = blockDim.x >> 5;
    auto block_sum = [&](float v) {
        v = warp_sum(v);
        __syncthreads();
        if (lane == 0) red[warp] = v;
        __syncthreads();
        float t = 0.f;
        for (int i = 0; i < nw_; ++i) t += red[i];
        return t;
    };
    float s = 0.f;
    for (int i = tid; i < d; i += blockDim.x) s += xr[i];
    const float mean = block_sum(s) / d;
    float q = 0.f;
    for (int i = tid; i < d; i += blockDim.x) { const float t = xr[i] - mean; q += t * t; }
    const float rstd = rsqrtf(block_sum(q) / d + eps);
    float m1 = 0.f, m2 = 0.f;
    for (int i = tid; i < d; i += blockDim.x) {
        const float xhat = (xr[i] - mean) * rstd;
        const float y = fmaf(xhat, nw[i], nb[i]);
        float dy = 0.f;
        for (int c = 0; c < C; ++c) {
            const float dlc = dl[c];
            dy = fmaf(dlc, __ldg(hw + static_cast<long long>(c) * d + i), dy);
            atomicAdd(dhw + static_cast<long long>(c) * d + i, dlc * y);
        }
        atomicAdd(dnw + i, dy * xhat);
        atomicAdd(dnb + i, dy);
        const float g = dy * nw[i];
        xh[i] = xhat;
        gb[i] = g;
        m1 += g;
        m2 += g * xhat;
    }
    m1 = block_sum(m1) / d;
    m2 = block_sum(m2) / d;
    float* dr = dx + blockIdx.x * bag_stride;
    for (int i = tid; i < d; i += blockDim.x) dr[i] = rstd * (gb[i] - m1 - xh[i] * m2);
    for (int c = tid; c < C; c += blockDim.x) atomicAdd(dhb + c, dl[c]);
}

// soft-target cross entropy with class weights, mean over the batch (F.cross_entropy with probability
// targets, src/stamp/modeling/models/__init__.py:254-258):
//   loss = 1/B sum_b sum_c -w_c y_bc log softmax(l_b)_c;   dl_bc = gscale/B * (p_bc * sum_c' w_c' y_bc' - w_c y_bc)
// one CTA, one warp per bag row.
__global__ void __launch_bounds__(256)
ce_loss_kernel(const float* __restrict__ logits, const float* __restrict__ targets, const float* __restrict__ cw,
               int B, int C, float gscale, float* __restrict__ loss, float* __restrict__ dlogits) {
    __shared__ float red[8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float acc = 0.f;
    for (int b = warp; b < B; b += 8) {
        const float* l = logits + static_cast<long long>(b) * C;
        const float* y = targets + static_cast<long long>(b) * C;
        float mx = -INFINITY;
        for (int c = lane; c < C; c += 32) mx = fmaxf(mx, l[c]);
        mx = warp_max(mx);
        float se = 0.f, wy = 0.f;
        for (int c = lane; c < C; c += 32) {
            se += expf(l[c] - mx);
            wy += (cw != nullptr ? cw[c] : 1.f) * y[c];
        }
        se = warp_sum(se);
        wy = warp_sum(wy);
        const float lse = mx + logf(se);
        for (int c = lane; c < C; c += 32) {
            const float w = (cw != nullptr ? cw[c] : 1.f) * y[c];
            const float logp = l[c] - lse;
            acc -= w * logp;
            if (dlogits != nullptr) dlogits[static_cast<long long>(b) * C + c] = gscale / B * (expf(logp) * wy - w);
        }
    }
    acc = warp_sum(acc);
    if (lane == 0) red[warp] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int i = 0; i < 8; ++i) t += red[i];
        *loss = t / B;
    }
}

// torch.optim.AdamW (decoupled weight decay, bias correction), one fused pass over flat buffers
__global__ void __launch_bounds__(256)
adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
             long long n, float lr, float beta1, float beta2, float eps, float wd, float bc1, float bc2_sqrt,
             float gscale) {
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const float gi = g[i] * gscale;
        float pi = p[i] * (1.0f - lr * wd);
        const float mi = beta1 * m[i] + (1.0f - beta1) * gi;
        const float vi = beta2 * v[i] + (1.0f - beta2) * gi * gi;
        m[i] = mi;
        v[i] = vi;
        pi -= (lr / bc1) * mi / (sqrtf(vi) / bc2_sqrt + eps);
        p[i] = pi;
    }
}

// ---------------------------------------------------------------------------------------------
// weight gradient: dW[Nout, Kin] += dY[M, Nout]^T . X[M, Kin]   (bf16 in, fp32 atomics out, split over M)
// Both operands are read in their natural token-major layout; the contraction runs over token rows,
// so A (= dY^T) and B (= X) fragments both come from transposing ldmatrix loads.  mma.sync tiles:
// CTA 128 x 128, 8 warps of 32 x 64, 32 token rows per cp.async stage.
// ---------------------------------------------------------------------------------------------
constexpr int WG_T = 128, WG_BK = 32, WG_LD = WG_T + 8, WG_THREADS = 256;

__global__ void __launch_bounds__(WG_THREADS)
wgrad_kernel(const uint16_t* __restrict__ dY, long long ldy, const uint16_t* __restrict__ X, long long ldx,
             float* __restrict__ dW, long long ldw, int M, int Nout, int Kin, int rows_per_split) {
    __shared__ __align__(16) uint16_t Ys[2][WG_BK * WG_LD];
    __shared__ __align__(16) uint16_t Xs[2][WG_BK * WG_LD];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t4 = lane & 3, mi = lane >> 3;
    const int n0 = blockIdx.x * WG_T, k0 = blockIdx.y * WG_T;
    const long long m_begin = static_cast<long long>(blockIdx.z) * rows_per_split;
    const long long m_end = (m_begin + rows_per_split < M) ? m_begin + rows_per_split : M;
    if (m_begin >= m_end) return;
    const int wm = warp & 3, wn = warp >> 2;
    const int nsteps = static_cast<int>((m_end - m_begin + WG_BK - 1) / WG_BK);

    auto load = [&](int step, int buf) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int c = tid + i * WG_THREADS;      // 512 16-byte chunks per tile
            const int r = c >> 4, ch = c & 15;
            const long long row = m_begin + static_cast<long long>(step) * WG_BK + r;
            const bool okr = row < m_end;
            {
                const int col = n0 + ch * 8;
                const bool ok = okr && col < Nout;
                cp_async_16(&Ys[buf][r * WG_LD + ch * 8], dY + (ok ? row * ldy + col : 0), ok);
            }
            {
                const int col = k0 + ch * 8;
                const bool ok = okr && col < Kin;
                cp_async_16(&Xs[buf][r * WG_LD + ch * 8], X + (ok ? row * ldx + col : 0), ok);
            }
        }
    };

    float acc[2][8][4];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 8; ++b) { acc[a][b][0] = acc[a][b][1] = acc[a][b][2] = acc[a][b][3] = 0.f; }

    load(0, 0);
    cp_async_commit();
    for (int st = 0; st < nsteps; ++st) {
        const int buf = st & 1;
        if (st + 1 < nsteps) {
            load(st + 1, buf ^ 1);
            cp_async_commit();
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
#pragma unroll
        for (int ks = 0; ks < WG_BK / 16; ++ks) {
            uint32_t af[2][4];
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
                const uint16_t* a = &Ys[buf][(ks * 16 + (mi >> 1) * 8 + (lane & 7)) * WG_LD + wm * 32 + mt * 16 + (mi & 1) * 8];
                ldmatrix_x4_trans(af[mt][0], af[mt][1], af[mt][2], af[mt][3], smem_u32(a));
            }
#pragma unroll
            for (int ntp = 0; ntp < 4; ++ntp) {
                uint32_t b0, b1, b2, b3;
                const uint16_t* a = &Xs[buf][(ks * 16 + (mi & 1) * 8 + (lane & 7)) * WG_LD + wn * 64 + ntp * 16 + (mi >> 1) * 8];
                ldmatrix_x4_trans(b0, b1, b2, b3, smem_u32(a));
#pragma unroll
                for (int mt = 0; mt < 2; ++mt) {
                    mma_bf16(acc[mt][2 * ntp], af[mt], b0, b1);
                    mma_bf16(acc[mt][2 * ntp + 1], af[mt], b2, b3);
                }
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            const int r = n0 + wm * 32 + mt * 16 + g;
            const int c = k0 + wn * 64 + nt * 8 + 2 * t4;
            if (c < Kin) {   // Kin % 8 == 0: c + 1 is in range too
                if (r < Nout) {
                    atomicAdd(dW + static_cast<long long>(r) * ldw + c, acc[mt][nt][0]);
                    atomicAdd(dW + static_cast<long long>(r) * ldw + c + 1, acc[mt][nt][1]);
                }
                if (r + 8 < Nout) {
                    atomicAdd(dW + static_cast<long long>(r + 8) * ldw + c, acc[mt][nt][2]);
                    atomicAdd(dW + static_cast<long long>(r + 8) * ldw + c + 1, acc[mt][nt][3]);
                }
            }
        }
}

// ---------------------------------------------------------------------------------------------
// host-side launch helpers
// ---------------------------------------------------------------------------------------------
#define SB_TRY(expr) do { const int rc__ = (expr); if (rc__ != SB_OK) return rc__; } while (0)

inline int last_status() { return cudaGetLastError() == cudaSuccess ? SB_OK : SB_ERR_CUDA; }

// scratch: per-split partial tiles for the tcgen05 kernel (part of the caller's ctx buffer), may be null
int wgrad(const uint16_t* dY, long long ldy, const uint16_t* X, long long ldx, float* dW, long long ldw, int M,
          int Nout, int Kin, float* scratch, size_t scratch_bytes, cudaStream_t stream) {
    if (M <= 0 || (Nout % 8) != 0 || (Kin % 8) != 0 || (ldy % 8) != 0 || (ldx % 8) != 0) return SB_ERR_BAD_ARG;
    {
        const int rc = wgrad_tc(dY, ldy, X, ldx, dW, ldw, M, Nout, Kin, scratch, scratch_bytes,
                                stream);   // long token dimension: tcgen05
        if (rc != SB_ERR_UNSUPPORTED) return rc;
    }
    const int tiles = ((Nout + WG_T - 1) / WG_T) * ((Kin + WG_T - 1) / WG_T);
    int splits = (2 * 148 + tiles - 1) / tiles;
    const int max_splits = (M + 4 * WG_BK - 1) / (4 * WG_BK);
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    int rows_per_split = (M + splits - 1) / splits;
    rows_per_split = (rows_per_split + WG_BK - 1) / WG_BK * WG_BK;
    splits = (M + rows_per_split - 1) / rows_per_split;
    dim3 grid((Nout + WG_T - 1) / WG_T, (Kin + WG_T - 1) / WG_T, splits);
    ProfScope prof(PROF_GEMM, 2.0 * M * static_cast<double>(Nout) * Kin, stream);
    wgrad_kernel<<<grid, WG_THREADS, 0, stream>>>(dY, ldy, X, ldx, dW, ldw, M, Nout, Kin, rows_per_split);
    count_launch();
    return last_status();
}

int colsum_bf16(const uint16_t* a, long long ld, long long rows, int cols, float* out, cudaStream_t stream) {
    if (rows <= 0 || (cols % 2) != 0) return SB_ERR_BAD_ARG;
    long long gy = (rows + 255) / 256;
    if (gy > 592) gy = 592;
    dim3 grid((cols + 63) / 64, static_cast<unsigned>(gy)), block(32, 8);
    ProfScope prof(PROF_ROWOP, rows * static_cast<double>(cols) * 2.0, stream);
    colsum_kernel<true><<<grid, block, 0, stream>>>(a, ld, rows, cols, out);
    count_launch();
    return last_status();
}
int colsum_f32(const float* a, long long ld, long long rows, int cols, float* out, cudaStream_t stream) {
    if (rows <= 0 || (cols % 2) != 0) return SB_ERR_BAD_ARG;
    long long gy = (rows + 255) / 256;
    if (gy > 592) gy = 592;
    dim3 grid((cols + 63) / 64, static_cast<unsigned>(gy)), block(32, 8);
    colsum_kernel<false><<<grid, block, 0, stream>>>(a, ld, rows, cols, out);
    count_launch();
    return last_status();
}

int cast_transpose(const float* w, int R, int C, uint16_t* w16, uint16_t* w16t, cudaStream_t stream) {
    dim3 grid((C + 31) / 32, (R + 31) / 32), block(32, 8);
    cast_transpose_kernel<<<grid, block, 0, stream>>>(w, R, C, w16, w16t);
    count_launch();
    return last_status();
}

int mask_cast(const float* in, uint16_t* out, long long n, uint32_t thresh, float scale, uint64_t sseed,
              cudaStream_t stream) {
    ProfScope prof(PROF_ROWOP, n * 6.0, stream);
    mask_cast_kernel<<<grid1d(n / 4, 256), 256, 0, stream>>>(in, out, n / 4, thresh, scale, sseed);
    count_launch();
    return last_status();
}

int ln_bwd(const float* dy, const float* x, const float* gamma, float* dx, float* dgamma, float* dbeta,
           int rows, int d, cudaStream_t stream, uint16_t* out16 = nullptr, uint32_t thresh = 0u, float mscale = 1.f,
           uint64_t sseed = 0) {
    if ((d % 4) != 0 || d > 1024) return SB_ERR_UNSUPPORTED;
    int blocks = (rows + 7) / 8;
    if (blocks > 148 * 4) blocks = 148 * 4;
    const size_t sh = 2 * static_cast<size_t>(d) * sizeof(float);
    ProfScope prof(PROF_ROWOP, static_cast<double>(rows) * d * (out16 != nullptr ? 18.0 : 16.0), stream);
    if (d <= 512)
        ln_bwd_kernel<4><<<blocks, 256, sh, stream>>>(dy, x, gamma, dx, dgamma, dbeta, rows, d, 1e-5f, out16, thresh, mscale, sseed);
    else
        ln_bwd_kernel<8><<<blocks, 256, sh, stream>>>(dy, x, gamma, dx, dgamma, dbeta, rows, d, 1e-5f, out16, thresh, mscale, sseed);
    count_launch();
    return last_status();
}

// ---------------------------------------------------------------------------------------------
// ctx layout
// ---------------------------------------------------------------------------------------------
struct TrainLayout {
    long long S, M, BN;
    size_t off_bags16, off_z0, off_coords, off_x /* (2L+1) x [M,d] f32 */, x_stride;
    size_t off_layer, layer_stride;   // per layer block, offsets inside:
    size_t l_xn1, l_qkv, l_att, l_osm, l_odv, l_lse, l_xn2, l_z1, l_h;
    size_t l_w_qkv, l_w_qkvT, l_w_fc, l_w_fcT, l_w_ff1, l_w_ff1T, l_w_ff2, l_w_ff2T;
    size_t off_w_proj, off_w_projT;
    size_t off_y32 /* [M, max(d,ff)] f32 scratch */, off_dx /* [M,d] f32 */, off_g16a /* [M, max(d,ff)] bf16 */,
        off_g16b /* [M,d] bf16 */, off_g16c /* [M,3d] bf16 */, off_delta, off_dsum, off_wscratch, wscratch_bytes,
        off_dist /* bf16 [B,S,ld] token distances (long ALiBi bags) */, off_dist_scale /* [B][2] */, total;
    bool dist;
};

inline size_t au(size_t v) { return (v + 255) / 256 * 256; }

bool make_train_layout(const StampMilConfig* c, int B, int N, TrainLayout* L) {
    if (c == nullptr || B <= 0 || N <= 0 || c->dim_input <= 0 || (c->dim_input % 8) != 0 || (c->dim_model % 8) != 0 ||
        (c->dim_ff % 8) != 0 || c->dim_model > 1024 || c->n_heads <= 0 || (c->dim_model % c->n_heads) != 0 ||
        c->dim_output <= 0 || c->n_layers <= 0)
        return false;
    const int hd = c->dim_model / c->n_heads;
    if (hd != 64 && hd != 32) return false;
    const size_t d = c->dim_model, F = c->dim_input, ff = c->dim_ff, H = c->n_heads;
    L->S = N + 1LL; L->M = L->S * B; L->BN = static_cast<long long>(B) * N;
    const size_t M = L->M, BN = L->BN, mx = d > ff ? d : ff;
    size_t o = 0;
    L->off_bags16 = o; o = au(o + BN * F * 2);
    L->off_z0 = o;     o = au(o + BN * d * 4);
    L->off_coords = o; o = au(o + M * 8);
    L->x_stride = au(M * d * 4);
    L->off_x = o;      o += L->x_stride * (2 * static_cast<size_t>(c->n_layers) + 1);
    size_t q = 0;
    L->l_xn1 = q; q = au(q + M * d * 2);
    L->l_qkv = q; q = au(q + M * 3 * d * 2);
    L->l_att = q; q = au(q + M * d * 2);
    L->l_osm = q; q = au(q + M * d * 4);
    L->l_odv = q; q = au(q + M * d * 4);
    L->l_lse = q; q = au(q + static_cast<size_t>(B) * H * L->S * 4);
    L->l_xn2 = q; q = au(q + M * d * 2);
    L->l_z1 = q;  q = au(q + M * ff * 4);
    L->l_h = q;   q = au(q + M * ff * 2);
    L->l_w_qkv = q;  q = au(q + 3 * d * d * 2);
    L->l_w_qkvT = q; q = au(q + 3 * d * d * 2);
    L->l_w_fc = q;   q = au(q + d * d * 2);
    L->l_w_fcT = q;  q = au(q + d * d * 2);
    L->l_w_ff1 = q;  q = au(q + ff * d * 2);
    L->l_w_ff1T = q; q = au(q + ff * d * 2);
    L->l_w_ff2 = q;  q = au(q + ff * d * 2);
    L->l_w_ff2T = q; q = au(q + ff * d * 2);
    L->layer_stride = q;
    L->off_layer = o;  o += q * c->n_layers;
    L->off_w_proj = o;  o = au(o + d * F * 2);
    L->off_w_projT = o; o = au(o + d * F * 2);
    L->off_y32 = o;  o = au(o + M * mx * 4);
    L->off_dx = o;   o = au(o + M * d * 4);
    L->off_g16a = o; o = au(o + M * mx * 2);
    L->off_g16b = o; o = au(o + M * d * 2);
    L->off_g16c = o; o = au(o + M * 3 * d * 2);
    L->off_delta = o; o = au(o + static_cast<size_t>(B) * H * L->S * 4);
    L->off_dsum = o;  o = au(o + 64);
    {   // per-split partial tiles of the largest weight gradient (wgrad_tc.cu)
        size_t w = wgrad_tc_scratch_bytes(static_cast<int>(3 * d), static_cast<int>(d));
        const size_t cand[4] = {wgrad_tc_scratch_bytes(static_cast<int>(d), static_cast<int>(F)), wgrad_tc_scratch_bytes(static_cast<int>(d), static_cast<int>(d)),
                                wgrad_tc_scratch_bytes(static_cast<int>(ff), static_cast<int>(d)), wgrad_tc_scratch_bytes(static_cast<int>(d), static_cast<int>(ff))};
        for (size_t c : cand) w = c > w ? c : w;
        L->wscratch_bytes = w;
    }
    L->off_wscratch = o; o = au(o + L->wscratch_bytes);
    // distance matrix of the third-generation attention forward (attention_mil_v3.cu), shared by all layers
    L->dist = c->use_alibi && hd == 64 && L->S > 256 && L->S <= 65535 && B <= 65535;
    L->off_dist_scale = o; o = au(o + static_cast<size_t>(B) * 8 + 16 + mil_dist16_scratch_bytes(B));
    L->off_dist = (o + 1023) / 1024 * 1024; o = L->off_dist;
    if (L->dist) o = au(o + mil_dist16_bytes(B, static_cast<int>(L->S)));
    L->total = o;
    return true;
}

GemmParams gp_bf16(int M, int N, int K, int store, void* out, long long ldo, const float* bias) {
    GemmParams p{};
    p.M = M; p.N = N; p.K = K; p.store = store; p.bf16 = 1; p.out = out; p.ldo = ldo; p.bias = bias;
    return p;
}

}  // namespace
}  // namespace sb

extern "C" {

size_t stamp_mil_train_ctx_bytes(const StampMilConfig* cfg, int B, int N) {
    sb::TrainLayout L;
    if (!sb::make_train_layout(cfg, B, N, &L)) return 0;
    return L.total;
}

int stamp_pairwise_dist_mean(const float* coords, int B, int N, float* mean_out, void* workspace,
                             size_t workspace_bytes, void* stream_) {
    using namespace sb;
    if (coords == nullptr || mean_out == nullptr || workspace == nullptr || B <= 0 || N <= 0) return SB_ERR_BAD_ARG;
    const long long S = N + 1LL, M = S * B;
    const size_t need = au(static_cast<size_t>(M) * 8) + 256;
    if (workspace_bytes < need) return SB_ERR_WORKSPACE;
    if ((reinterpret_cast<uintptr_t>(workspace) & 255) != 0) return SB_ERR_BAD_ARG;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    float2* coords_s = static_cast<float2*>(workspace);
    double* dsum = reinterpret_cast<double*>(static_cast<uint8_t*>(workspace) + au(static_cast<size_t>(M) * 8));
    SB_TRY(mil_prepare(coords, nullptr, coords_s, nullptr, B, N, stream));
    if (cudaMemsetAsync(dsum, 0, sizeof(double), stream) != cudaSuccess) return SB_ERR_CUDA;
    dim3 grid(static_cast<unsigned>((S + 255) / 256), B, DIST_SPLITS);
    dist_sum_kernel<<<grid, 256, 0, stream>>>(coords_s, static_cast<int>(S), dsum);
    dist_mean_finish_kernel<<<1, 1, 0, stream>>>(dsum, static_cast<double>(B) * S * S, mean_out);
    count_launch(2);
    return last_status();
}

size_t stamp_pairwise_dist_mean_workspace_bytes(int B, int N) {
    if (B <= 0 || N <= 0) return 0;
    return sb::au(static_cast<size_t>(B) * (N + 1) * 8) + 256;
}

int stamp_mil_train_dropout_mask(unsigned long long seed, int site, long long n, float p, uint8_t* keep_out,
                                 void* stream_) {
    using namespace sb;
    if (keep_out == nullptr || n <= 0 || site < 0 || p < 0.f || p >= 1.f) return SB_ERR_BAD_ARG;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    dropout_mask_kernel<<<grid1d(n, 256), 256, 0, stream>>>(keep_out, n, drop_thresh(p), site_seed(seed, site));
    count_launch();
    return last_status();
}

int stamp_mil_train_forward(const StampMilConfig* cfg, const StampMilTrainTop* top, const StampMilTrainLayer* layers,
                            const StampMilTrainStep* step, const float* bags, const float* coords, float* logits,
                            int B, int N, void* ctx, size_t ctx_bytes, void* stream_) {
    using namespace sb;
    TrainLayout L;
    if (!make_train_layout(cfg, B, N, &L) || top == nullptr || layers == nullptr || step == nullptr ||
        bags == nullptr || coords == nullptr || logits == nullptr || ctx == nullptr ||
        (cfg->use_alibi && step->inv_rm == nullptr) || step->p_drop_proj < 0.f || step->p_drop_proj >= 1.f || step->p_drop_ff < 0.f || step->p_drop_ff >= 1.f)
        return SB_ERR_BAD_ARG;
    if (ctx_bytes < L.total) return SB_ERR_WORKSPACE;
    if ((reinterpret_cast<uintptr_t>(ctx) & 255) != 0) return SB_ERR_BAD_ARG;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    uint8_t* ws = static_cast<uint8_t*>(ctx);
    const int d = cfg->dim_model, F = cfg->dim_input, ff = cfg->dim_ff, H = cfg->n_heads, hd = d / H;
    const int S = static_cast<int>(L.S), M = static_cast<int>(L.M), BN = static_cast<int>(L.BN);
    uint16_t* bags16 = reinterpret_cast<uint16_t*>(ws + L.off_bags16);
    float* z0 = reinterpret_cast<float*>(ws + L.off_z0);
    float2* coords_s = reinterpret_cast<float2*>(ws + L.off_coords);
    float* y32 = reinterpret_cast<float*>(ws + L.off_y32);
    uint16_t* w_proj = reinterpret_cast<uint16_t*>(ws + L.off_w_proj);
    uint16_t* w_projT = reinterpret_cast<uint16_t*>(ws + L.off_w_projT);
    auto xbuf = [&](int i) { return reinterpret_cast<float*>(ws + L.off_x + L.x_stride * i); };

    // bf16 operand copies of the weights (+ transposes for the data-gradient GEMMs)
    SB_TRY(cast_transpose(top->proj_w, d, F, w_proj, w_projT, stream));
    for (int l = 0; l < cfg->n_layers; ++l) {
        uint8_t* lb = ws + L.off_layer + L.layer_stride * l;
        const StampMilTrainLayer& y = layers[l];
        SB_TRY(cast_transpose(y.qkv_w, 3 * d, d, reinterpret_cast<uint16_t*>(lb + L.l_w_qkv), reinterpret_cast<uint16_t*>(lb + L.l_w_qkvT), stream));
        SB_TRY(cast_transpose(y.fc_w, d, d, reinterpret_cast<uint16_t*>(lb + L.l_w_fc), reinterpret_cast<uint16_t*>(lb + L.l_w_fcT), stream));
        SB_TRY(cast_transpose(y.ff1_w, ff, d, reinterpret_cast<uint16_t*>(lb + L.l_w_ff1), reinterpret_cast<uint16_t*>(lb + L.l_w_ff1T), stream));
        SB_TRY(cast_transpose(y.ff2_w, d, ff, reinterpret_cast<uint16_t*>(lb + L.l_w_ff2), reinterpret_cast<uint16_t*>(lb + L.l_w_ff2T), stream));
    }

    // project_features: Linear -> GELU -> Dropout into rows 1.. of every bag; class token into row 0
    const long long nb = static_cast<long long>(BN) * F;
    SB_TRY(mask_cast(bags, bags16, nb, 0u, 1.f, 0, stream));
    {
        GemmParams p = gp_bf16(BN, d, F, ST_32, z0, d, top->proj_b);
        SB_TRY(gemm_tn(bags16, F, w_proj, F, p, stream));
    }
    float* x = xbuf(0);
    {
        ProfScope prof(PROF_ROWOP, static_cast<double>(BN) * d * 8.0, stream);
        gelu_drop_fwd_kernel<<<grid1d(static_cast<long long>(BN) * d / 4, 256), 256, 0, stream>>>(
            z0, BN, d, x, nullptr, N, S, 1, drop_thresh(step->p_drop_proj), inv_keep(step->p_drop_proj),
            site_seed(step->seed, 0));
        count_launch();
    }
    SB_TRY(fill_rows(x, d, B, S, 0, top->class_token, d, nullptr, 0, 1, d, stream));
    SB_TRY(mil_prepare(coords, nullptr, coords_s, nullptr, B, N, stream));
    uint16_t* dist16 = reinterpret_cast<uint16_t*>(ws + L.off_dist);
    float* dist_scale = reinterpret_cast<float*>(ws + L.off_dist_scale);
    if (L.dist) SB_TRY(mil_dist16(reinterpret_cast<const float*>(coords_s), B, S, 1, dist_scale, dist16, dist_scale + ((2 * static_cast<size_t>(B) + 3) & ~static_cast<size_t>(3)), stream));

    const float scale = 1.0f / sqrtf(static_cast<float>(hd));
    for (int l = 0; l < cfg->n_layers; ++l) {
        uint8_t* lb = ws + L.off_layer + L.layer_stride * l;
        const StampMilTrainLayer& y = layers[l];
        uint16_t* xn1 = reinterpret_cast<uint16_t*>(lb + L.l_xn1);
        uint16_t* qkv = reinterpret_cast<uint16_t*>(lb + L.l_qkv);
        uint16_t* att = reinterpret_cast<uint16_t*>(lb + L.l_att);
        uint16_t* xn2 = reinterpret_cast<uint16_t*>(lb + L.l_xn2);
        float* z1 = reinterpret_cast<float*>(lb + L.l_z1);
        uint16_t* h = reinterpret_cast<uint16_t*>(lb + L.l_h);
        float* x_in = xbuf(2 * l);
        float* x_mid = xbuf(2 * l + 1);
        float* x_out = xbuf(2 * l + 2);

        SB_TRY(layernorm(x_in, d, y.ln1_w, y.ln1_b, xn1, nullptr, d, M, d, 1e-5f, 1, stream));
        {
            GemmParams p = gp_bf16(M, 3 * d, d, ST_16, qkv, 3 * d, y.qkv_b);
            SB_TRY(gemm_tn(xn1, d, lb + L.l_w_qkv, d, p, stream));
        }
        AttnTrainParams a{};
        a.q = qkv; a.k = qkv + d; a.v = qkv + 2 * d;
        a.row_stride = 3LL * d; a.batch_stride = 3LL * d * S;
        a.out = att; a.osm = reinterpret_cast<float*>(lb + L.l_osm); a.odv = reinterpret_cast<float*>(lb + L.l_odv); a.lse2 = reinterpret_cast<float*>(lb + L.l_lse);
        a.out_row_stride = d; a.out_batch_stride = static_cast<long long>(d) * S;
        a.B = B; a.S = S; a.H = H; a.scale = scale; a.scale_log2 = scale * 1.4426950408889634f;
        if (cfg->use_alibi) {   // else: plain softmax attention (nn.MultiheadAttention, vision_tranformer.py:191,218-228)
            a.coords = coords_s; a.beta = y.bias_scale; a.inv_rm = step->inv_rm + static_cast<size_t>(l) * H;
            if (L.dist) { a.dist16 = dist16; a.dist_scale = dist_scale; }
        }
        SB_TRY(attention_train_fwd(a, hd, stream));
        // x_mid = x_in + fc(att)
        if (cudaMemcpyAsync(x_mid, x_in, static_cast<size_t>(M) * d * 4, cudaMemcpyDeviceToDevice, stream) != cudaSuccess)
            return SB_ERR_CUDA;
        {
            GemmParams p = gp_bf16(M, d, d, ST_RESID32, x_mid, d, y.fc_b);
            SB_TRY(gemm_tn(att, d, lb + L.l_w_fc, d, p, stream));
        }
        SB_TRY(layernorm(x_mid, d, y.ln2_w, y.ln2_b, xn2, nullptr, d, M, d, 1e-5f, 1, stream));
        {
            GemmParams p = gp_bf16(M, ff, d, ST_32, z1, ff, y.ff1_b);
            SB_TRY(gemm_tn(xn2, d, lb + L.l_w_ff1, d, p, stream));
        }
        {
            ProfScope prof(PROF_ROWOP, static_cast<double>(M) * ff * 6.0, stream);
            gelu_drop_fwd_kernel<<<grid1d(static_cast<long long>(M) * ff / 4, 256), 256, 0, stream>>>(
                z1, M, ff, nullptr, h, 0, 0, 0, drop_thresh(step->p_drop_ff), inv_keep(step->p_drop_ff),
                site_seed(step->seed, 1 + 2 * l));
            count_launch();
        }
        {
            GemmParams p = gp_bf16(M, d, ff, ST_32, y32, d, y.ff2_b);
            SB_TRY(gemm_tn(h, ff, lb + L.l_w_ff2, ff, p, stream));
        }
        {
            ProfScope prof(PROF_ROWOP, static_cast<double>(M) * d * 12.0, stream);
            resid_drop_fwd_kernel<<<grid1d(static_cast<long long>(M) * d / 4, 256), 256, 0, stream>>>(
                x_mid, y32, x_out, static_cast<long long>(M) * d / 4, drop_thresh(step->p_drop_ff),
                inv_keep(step->p_drop_ff), site_seed(step->seed, 2 + 2 * l));
            count_launch();
        }
    }
    SB_TRY(cls_head(xbuf(2 * cfg->n_layers), static_cast<long long>(S) * d, d, top->norm_w, top->norm_b, top->head_w,
                    top->head_b, cfg->dim_output, B, logits, stream));
    return last_status();
}

int stamp_mil_train_backward(const StampMilConfig* cfg, const StampMilTrainTop* top, const StampMilTrainLayer* layers,
                             const StampMilTrainStep* step, const float* dlogits, StampMilTrainTop* gtop,
                             StampMilTrainLayer* glayers, float* dbags, int B, int N, void* ctx, size_t ctx_bytes,
                             void* stream_) {
    using namespace sb;
    TrainLayout L;
    if (!make_train_layout(cfg, B, N, &L) || top == nullptr || layers == nullptr || step == nullptr ||
        dlogits == nullptr || gtop == nullptr || glayers == nullptr || ctx == nullptr ||
        (cfg->use_alibi && step->inv_rm == nullptr))
        return SB_ERR_BAD_ARG;
    if (ctx_bytes < L.total) return SB_ERR_WORKSPACE;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    uint8_t* ws = static_cast<uint8_t*>(ctx);
    const int d = cfg->dim_model, F = cfg->dim_input, ff = cfg->dim_ff, H = cfg->n_heads, hd = d / H;
    const int S = static_cast<int>(L.S), M = static_cast<int>(L.M), BN = static_cast<int>(L.BN);
    uint16_t* bags16 = reinterpret_cast<uint16_t*>(ws + L.off_bags16);
    float* z0 = reinterpret_cast<float*>(ws + L.off_z0);
    float2* coords_s = reinterpret_cast<float2*>(ws + L.off_coords);
    float* g32 = reinterpret_cast<float*>(ws + L.off_y32);
    float* dx = reinterpret_cast<float*>(ws + L.off_dx);
    uint16_t* g16a = reinterpret_cast<uint16_t*>(ws + L.off_g16a);
    uint16_t* g16b = reinterpret_cast<uint16_t*>(ws + L.off_g16b);
    uint16_t* g16c = reinterpret_cast<uint16_t*>(ws + L.off_g16c);
    float* delta = reinterpret_cast<float*>(ws + L.off_delta);
    auto xbuf = [&](int i) { return reinterpret_cast<float*>(ws + L.off_x + L.x_stride * i); };
    const long long Md = static_cast<long long>(M) * d;
    float* wsc = reinterpret_cast<float*>(ws + L.off_wscratch);
    const size_t wsb = L.wscratch_bytes;
    const uint32_t th_ff = drop_thresh(step->p_drop_ff);
    const float ik_ff = inv_keep(step->p_drop_ff);

    // head + final LayerNorm: only the class-token rows carry gradient
    if (cudaMemsetAsync(dx, 0, static_cast<size_t>(Md) * 4, stream) != cudaSuccess) return SB_ERR_CUDA;
    cls_head_bwd_kernel<<<B, 256, (2 * d + 32) * sizeof(float), stream>>>(
        xbuf(2 * cfg->n_layers), static_cast<long long>(S) * d, d, top->norm_w, top->norm_b, top->head_w,
        cfg->dim_output, 1e-5f, dlogits, dx, gtop->norm_w, gtop->norm_b, gtop->head_w, gtop->head_b);
    count_launch();

    const float scale = 1.0f / sqrtf(static_cast<float>(hd));
    for (int l = cfg->n_layers - 1; l >= 0; --l) {
        uint8_t* lb = ws + L.off_layer + L.layer_stride * l;
        const StampMilTrainLayer& y = layers[l];
        StampMilTrainLayer& gy = glayers[l];
        uint16_t* xn1 = reinterpret_cast<uint16_t*>(lb + L.l_xn1);
        uint16_t* qkv = reinterpret_cast<uint16_t*>(lb + L.l_qkv);
        uint16_t* att = reinterpret_cast<uint16_t*>(lb + L.l_att);
        uint16_t* xn2 = reinterpret_cast<uint16_t*>(lb + L.l_xn2);
        float* z1 = reinterpret_cast<float*>(lb + L.l_z1);
        uint16_t* h = reinterpret_cast<uint16_t*>(lb + L.l_h);
        float* x_in = xbuf(2 * l);
        float* x_mid = xbuf(2 * l + 1);

        // ---- feed-forward block: x_out = x_mid + drop(ff2(drop(gelu(ff1(LN2(x_mid)))))) ----
        // dy [M,d] = bf16(dropout mask . dx): written by the LayerNorm backward of the layer above (below), by a
        // pass of its own only for the top layer
        if (l == cfg->n_layers - 1)
            SB_TRY(mask_cast(dx, g16b, Md, th_ff, ik_ff, site_seed(step->seed, 2 + 2 * l), stream));
        SB_TRY(wgrad(g16b, d, h, ff, gy.ff2_w, ff, M, d, ff, wsc, wsb, stream));
        SB_TRY(colsum_bf16(g16b, d, M, d, gy.ff2_b, stream));
        {
            GemmParams p = gp_bf16(M, ff, d, ST_32, g32, ff, nullptr);                              // dh [M,ff]
            SB_TRY(gemm_tn(g16b, d, lb + L.l_w_ff2T, d, p, stream));
        }
        {
            ProfScope prof(PROF_ROWOP, static_cast<double>(M) * ff * 10.0, stream);
            gelu_drop_bwd_kernel<<<grid1d(static_cast<long long>(M) * ff / 4, 256), 256, 0, stream>>>(
                g32, z1, M, ff, g16a, 0, 0, 0, th_ff, ik_ff, site_seed(step->seed, 1 + 2 * l));    // dz1 [M,ff]
            count_launch();
        }
        SB_TRY(wgrad(g16a, ff, xn2, d, gy.ff1_w, d, M, ff, d, wsc, wsb, stream));
        SB_TRY(colsum_bf16(g16a, ff, M, ff, gy.ff1_b, stream));
        {
            GemmParams p = gp_bf16(M, d, ff, ST_32, g32, d, nullptr);                               // d xn2 [M,d]
            SB_TRY(gemm_tn(g16a, ff, lb + L.l_w_ff1T, ff, p, stream));
        }
        // dx = d x_mid, and its bf16 copy for the attention block: x_mid = x_in + fc(attn(qkv(LN1(x_in))))
        SB_TRY(ln_bwd(g32, x_mid, y.ln2_w, dx, gy.ln2_w, gy.ln2_b, M, d, stream, g16b));
        SB_TRY(wgrad(g16b, d, att, d, gy.fc_w, d, M, d, d, wsc, wsb, stream));
        SB_TRY(colsum_bf16(g16b, d, M, d, gy.fc_b, stream));
        {
            GemmParams p = gp_bf16(M, d, d, ST_32, g32, d, nullptr);                                // d att [M,d] fp32
            SB_TRY(gemm_tn(g16b, d, lb + L.l_w_fcT, d, p, stream));
        }
        AttnTrainParams a{};
        a.q = qkv; a.k = qkv + d; a.v = qkv + 2 * d;
        a.row_stride = 3LL * d; a.batch_stride = 3LL * d * S;
        a.out = att; a.osm = reinterpret_cast<float*>(lb + L.l_osm); a.odv = reinterpret_cast<float*>(lb + L.l_odv); a.lse2 = reinterpret_cast<float*>(lb + L.l_lse);
        a.out_row_stride = d; a.out_batch_stride = static_cast<long long>(d) * S;
        a.B = B; a.S = S; a.H = H; a.scale = scale; a.scale_log2 = scale * 1.4426950408889634f;
        if (cfg->use_alibi) {
            a.coords = coords_s; a.beta = y.bias_scale; a.inv_rm = step->inv_rm + static_cast<size_t>(l) * H;
            a.dbeta = gy.bias_scale;
        }
        a.dout32 = g32; a.dout = g16a; a.delta = delta;
        a.dq = g16c; a.dk = g16c + d; a.dv = g16c + 2 * d;
        SB_TRY(attention_train_bwd(a, hd, stream));
        SB_TRY(wgrad(g16c, 3LL * d, xn1, d, gy.qkv_w, d, M, 3 * d, d, wsc, wsb, stream));
        SB_TRY(colsum_bf16(g16c, 3LL * d, M, 3 * d, gy.qkv_b, stream));
        {
            GemmParams p = gp_bf16(M, d, 3 * d, ST_32, g32, d, nullptr);                            // d xn1 [M,d]
            SB_TRY(gemm_tn(g16c, 3LL * d, lb + L.l_w_qkvT, 3LL * d, p, stream));
        }
        // dx = d x_in; for the layer below also dy of its feed-forward block (its output dropout mask applied)
        if (l > 0)
            SB_TRY(ln_bwd(g32, x_in, y.ln1_w, dx, gy.ln1_w, gy.ln1_b, M, d, stream, g16b, th_ff, ik_ff,
                          site_seed(step->seed, 2 + 2 * (l - 1))));
        else
            SB_TRY(ln_bwd(g32, x_in, y.ln1_w, dx, gy.ln1_w, gy.ln1_b, M, d, stream));
    }

    // class token rows, then project_features
    SB_TRY(colsum_f32(dx, static_cast<long long>(S) * d, B, d, gtop->class_token, stream));
    {
        ProfScope prof(PROF_ROWOP, static_cast<double>(BN) * d * 10.0, stream);
        gelu_drop_bwd_kernel<<<grid1d(static_cast<long long>(BN) * d / 4, 256), 256, 0, stream>>>(
            dx, z0, BN, d, g16a, N, S, 1, drop_thresh(step->p_drop_proj), inv_keep(step->p_drop_proj),
            site_seed(step->seed, 0));                                                              // dz0 [BN,d]
        count_launch();
    }
    SB_TRY(wgrad(g16a, d, bags16, F, gtop->proj_w, F, BN, d, F, wsc, wsb, stream));
    SB_TRY(colsum_bf16(g16a, d, BN, d, gtop->proj_b, stream));
    if (dbags != nullptr) {
        GemmParams p = gp_bf16(BN, F, d, ST_32, dbags, F, nullptr);
        SB_TRY(gemm_tn(g16a, d, ws + L.off_w_projT, d, p, stream));
    }
    return last_status();
}

int stamp_cross_entropy(const float* logits, const float* targets, const float* class_weights, int B, int C,
                        float grad_scale, float* loss_out, float* dlogits_out, void* stream_) {
    using namespace sb;
    if (logits == nullptr || targets == nullptr || loss_out == nullptr || B <= 0 || C <= 0) return SB_ERR_BAD_ARG;
    ce_loss_kernel<<<1, 256, 0, static_cast<cudaStream_t>(stream_)>>>(logits, targets, class_weights, B, C, grad_scale,
                                                                      loss_out, dlogits_out);
    count_launch();
    return last_status();
}

int stamp_adamw_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, long long n, float lr,
                     float beta1, float beta2, float eps, float weight_decay, int step, float grad_scale,
                     void* stream_) {
    using namespace sb;
    if (params == nullptr || grads == nullptr || exp_avg == nullptr || exp_avg_sq == nullptr || n <= 0 || step < 1)
        return SB_ERR_BAD_ARG;
    const float bc1 = 1.0f - powf(beta1, static_cast<float>(step));
    const float bc2 = 1.0f - powf(beta2, static_cast<float>(step));
    ProfScope prof(PROF_ROWOP, n * 28.0, static_cast<cudaStream_t>(stream_));
    adamw_kernel<<<grid1d(n, 256), 256, 0, static_cast<cudaStream_t>(stream_)>>>(
        params, grads, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, weight_decay, bc1, sqrtf(bc2), grad_scale);
    count_launch();
    return last_status();
}

}  // extern "C"
