// Row-wise HBM-bound kernels: LayerNorm (fp32 residual stream -> 16-bit GEMM operand or fp32),
// token-row fills, and the uint8 tile -> normalised patch-matrix transform.
//
// reference: LayerNorm = timm Block.norm1/norm2/norm (eps 1e-6) and the MIL aggregator's
// nn.LayerNorm (src/stamp/modeling/models/vision_tranformer.py:161,186,277, eps 1e-5);
// tile transform = ToTensor + Normalize applied per tile in
// src/stamp/preprocessing/__init__.py:94 (Extractor.transform).
#include "rowops.cuh"

#include "common.cuh"

namespace sb {

namespace {

// One warp per row; row cached in registers (cols <= 32 * 4 * MAXV), two-pass variance
// (mean first, then sum of squared deviations) to match torch's numerics.
template <int MAXV>
__global__ void __launch_bounds__(256)
layernorm_kernel(const float* __restrict__ x, long long ldx, const float* __restrict__ w,
                 const float* __restrict__ b, void* __restrict__ out, uint16_t* __restrict__ out_lo,
                 long long ldo, int rows, int cols, int cols_real, float eps, int out_kind /*0 fp16, 1 bf16, 2 fp32*/) {
    // cols_real <= cols: statistics over the first cols_real columns only (zero-padded widths: the padding columns
    // hold zeros and have zero weight / bias, so they come out as zeros)
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= rows) return;
    const float* xr = x + static_cast<long long>(warp) * ldx;
    const int nvec = cols >> 2;  // cols % 4 == 0 enforced by the launcher
    float4 v[MAXV];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
        const int idx = lane + i * 32;
        if (idx < nvec) {
            v[i] = *reinterpret_cast<const float4*>(xr + idx * 4);
            s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
        }
    }
    s = warp_sum(s);
    const float mean = s / static_cast<float>(cols_real);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
        const int idx = lane + i * 32;
        if (idx < nvec) {
            float dx = v[i].x - mean, dy = v[i].y - mean, dz = v[i].z - mean, dw = v[i].w - mean;
            if (cols_real < cols) {      // (warp-uniform) padded width: only real columns count
                const int c0 = idx * 4;
                if (c0 >= cols_real) dx = 0.f;
                if (c0 + 1 >= cols_real) dy = 0.f;
                if (c0 + 2 >= cols_real) dz = 0.f;
                if (c0 + 3 >= cols_real) dw = 0.f;
            }
            q += (dx * dx + dy * dy) + (dz * dz + dw * dw);
        }
    }
    q = warp_sum(q);
    const float rstd = rsqrtf(q / static_cast<float>(cols_real) + eps);
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
        const int idx = lane + i * 32;
        if (idx < nvec) {
            float4 ww = __ldg(reinterpret_cast<const float4*>(w) + idx);
            float4 bb = __ldg(reinterpret_cast<const float4*>(b) + idx);
            float y0 = (v[i].x - mean) * rstd * ww.x + bb.x;
            float y1 = (v[i].y - mean) * rstd * ww.y + bb.y;
            float y2 = (v[i].z - mean) * rstd * ww.z + bb.z;
            float y3 = (v[i].w - mean) * rstd * ww.w + bb.w;
            if (out_kind == 2) {
                float* o = reinterpret_cast<float*>(out) + static_cast<long long>(warp) * ldo;
                *reinterpret_cast<float4*>(o + idx * 4) = make_float4(y0, y1, y2, y3);
            } else {
                uint16_t* o = reinterpret_cast<uint16_t*>(out) + static_cast<long long>(warp) * ldo;
                uint2 pk;
                pk.x = pack_16(y0, y1, out_kind == 1);
                pk.y = pack_16(y2, y3, out_kind == 1);
                *reinterpret_cast<uint2*>(o + idx * 4) = pk;
                if (out_lo != nullptr) {
                    // split precision: y = hi + lo with hi = fp16(y), lo = fp16(y - hi)
                    const __half2 h01 = *reinterpret_cast<__half2*>(&pk.x), h23 = *reinterpret_cast<__half2*>(&pk.y);
                    uint2 lo;
                    lo.x = pack_f16(y0 - __low2float(h01), y1 - __high2float(h01));
                    lo.y = pack_f16(y2 - __low2float(h23), y3 - __high2float(h23));
                    *reinterpret_cast<uint2*>(out_lo + static_cast<long long>(warp) * ldo + idx * 4) = lo;
                }
            }
        }
    }
}

// x[g * rows_per_group + row_off + r, :] = src[r, :] (+ add[r, :])     fp32 rows
__global__ void fill_rows_kernel(float* __restrict__ x, long long ldx, int groups,
                                 int rows_per_group, int row_off, const float* __restrict__ src,
                                 long long lds, const float* __restrict__ add, long long lda,
                                 int nrows, int cols) {
    const long long total = static_cast<long long>(groups) * nrows * (cols >> 2);
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int c4 = static_cast<int>(i % (cols >> 2));
        const long long t = i / (cols >> 2);
        const int r = static_cast<int>(t % nrows);
        const int g = static_cast<int>(t / nrows);
        float4 v = __ldg(reinterpret_cast<const float4*>(src + r * lds) + c4);
        if (add != nullptr) {
            float4 a = __ldg(reinterpret_cast<const float4*>(add + r * lda) + c4);
            v.x += a.x; v.y += a.y; v.z += a.z; v.w += a.w;
        }
        float* o = x + (static_cast<long long>(g) * rows_per_group + row_off + r) * ldx;
        reinterpret_cast<float4*>(o)[c4] = v;
    }
}

// uint8 tiles [B, img, img, 3] (HWC) -> 16-bit patch matrix [B * (img/P)^2, 3 * P * P] with
// column order (c, ky, kx), i.e. the flattening of a Conv2d(3, D, P, stride=P) weight, after
// x/255 -> (x - mean) / std.  One thread handles 8 consecutive kx of one (patch, c, ky) row:
// 24 bytes of input (3 channels interleaved), one 16-byte store.
__global__ void __launch_bounds__(256)
tiles_to_patches_kernel(const uint8_t* __restrict__ tiles, uint16_t* __restrict__ patches, int B,
                        int img, int P, float3 scale, float3 shift, int bf16) {
    const int gp = img / P;           // patches per side
    const int kx8 = P / 8;            // 8-pixel groups per patch row (P = 16 -> 2; P = 14 handled below)
    const long long total = static_cast<long long>(B) * gp * gp * P * kx8;
    const int K = 3 * P * P;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        long long t = i;
        const int xg = static_cast<int>(t % kx8); t /= kx8;
        const int ky = static_cast<int>(t % P); t /= P;
        const int px = static_cast<int>(t % gp); t /= gp;
        const int py = static_cast<int>(t % gp); t /= gp;
        const int b = static_cast<int>(t);
        const int y = py * P + ky;
        const int x0 = px * P + xg * 8;
        const uint8_t* src = tiles + ((static_cast<long long>(b) * img + y) * img + x0) * 3;
        uint8_t px24[24];
        // 24 bytes, 8-byte aligned because x0 % 8 == 0 and row pitch img*3 is a multiple of 8 for img % 8 == 0
        const uint2* s2 = reinterpret_cast<const uint2*>(src);
        uint2 a0 = __ldg(s2), a1 = __ldg(s2 + 1), a2 = __ldg(s2 + 2);
        *reinterpret_cast<uint2*>(px24) = a0;
        *reinterpret_cast<uint2*>(px24 + 8) = a1;
        *reinterpret_cast<uint2*>(px24 + 16) = a2;
        const long long prow = (static_cast<long long>(b) * gp + py) * gp + px;
        const float sc[3] = {scale.x, scale.y, scale.z};
        const float sh[3] = {shift.x, shift.y, shift.z};
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            float f[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) f[j] = fmaf(static_cast<float>(px24[j * 3 + c]), sc[c], sh[c]);
            uint4 w;
            w.x = pack_16(f[0], f[1], bf16);
            w.y = pack_16(f[2], f[3], bf16);
            w.z = pack_16(f[4], f[5], bf16);
            w.w = pack_16(f[6], f[7], bf16);
            uint16_t* o = patches + prow * K + (c * P + ky) * P + xg * 8;
            *reinterpret_cast<uint4*>(o) = w;
        }
    }
}

// generic (any P, e.g. 14): one thread per output element pair; slower but only used for P % 8 != 0
__global__ void __launch_bounds__(256)
tiles_to_patches_generic_kernel(const uint8_t* __restrict__ tiles, uint16_t* __restrict__ patches,
                                int B, int img, int P, int Kpad, float3 scale, float3 shift,
                                int bf16) {
    const int gp = img / P;
    const int K = 3 * P * P;
    const long long total = static_cast<long long>(B) * gp * gp * Kpad;
    const float sc[3] = {scale.x, scale.y, scale.z};
    const float sh[3] = {shift.x, shift.y, shift.z};
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int k = static_cast<int>(i % Kpad);
        const long long prow = i / Kpad;
        float f = 0.f;
        if (k < K) {
            const int kx = k % P, ky = (k / P) % P, c = k / (P * P);
            const int px = static_cast<int>(prow % gp);
            const int py = static_cast<int>((prow / gp) % gp);
            const long long b = prow / (gp * gp);
            const uint8_t v = __ldg(tiles + ((b * img + py * P + ky) * img + px * P + kx) * 3 + c);
            f = fmaf(static_cast<float>(v), sc[c], sh[c]);
        }
        uint32_t w = pack_16(f, 0.f, bf16);
        patches[i] = static_cast<uint16_t>(w & 0xFFFF);
    }
}

// fp32 -> fp16 hi (+ optional lo = fp16(v - hi)), 8 elements per thread, 128-bit loads/stores
__global__ void __launch_bounds__(256)
cast_split_kernel(const float* __restrict__ in, __half* __restrict__ hi, __half* __restrict__ lo,
                  long long n8, long long n) {
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n8;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const uint4 a = ld_nc_v4(in + i * 8), b = ld_nc_v4(in + i * 8 + 4);
        const float v[8] = {__uint_as_float(a.x), __uint_as_float(a.y), __uint_as_float(a.z), __uint_as_float(a.w),
                            __uint_as_float(b.x), __uint_as_float(b.y), __uint_as_float(b.z), __uint_as_float(b.w)};
        uint4 w;
        w.x = pack_f16(v[0], v[1]); w.y = pack_f16(v[2], v[3]); w.z = pack_f16(v[4], v[5]); w.w = pack_f16(v[6], v[7]);
        *reinterpret_cast<uint4*>(hi + i * 8) = w;
        if (lo != nullptr) {
            const __half2* h = reinterpret_cast<const __half2*>(&w);
            uint4 l;
            l.x = pack_f16(v[0] - __low2float(h[0]), v[1] - __high2float(h[0]));
            l.y = pack_f16(v[2] - __low2float(h[1]), v[3] - __high2float(h[1]));
            l.z = pack_f16(v[4] - __low2float(h[2]), v[5] - __high2float(h[2]));
            l.w = pack_f16(v[6] - __low2float(h[3]), v[7] - __high2float(h[3]));
            *reinterpret_cast<uint4*>(lo + i * 8) = l;
        }
    }
    if (blockIdx.x == 0 && threadIdx.x < (n - n8 * 8)) {
        const long long j = n8 * 8 + threadIdx.x;
        const __half h = __float2half_rn(in[j]);
        hi[j] = h;
        if (lo != nullptr) lo[j] = __float2half_rn(in[j] - __half2float(h));
    }
}

inline int grid_for(long long total, int block) {
    long long g = (total + block - 1) / block;
    const long long cap = 148LL * 16;
    return static_cast<int>(g < cap ? (g > 0 ? g : 1) : cap);
}

}  // namespace

int layernorm(const float* x, long long ldx, const float* w, const float* b, void* out, void* out_lo,
              long long ldo, int rows, int cols, float eps, int out_kind, cudaStream_t stream) {
    return layernorm_padded(x, ldx, w, b, out, out_lo, ldo, rows, cols, cols, eps, out_kind, stream);
}

int layernorm_padded(const float* x, long long ldx, const float* w, const float* b, void* out, void* out_lo,
                     long long ldo, int rows, int cols, int cols_real, float eps, int out_kind, cudaStream_t stream) {
    if (cols_real <= 0 || cols_real > cols) return SB_ERR_BAD_ARG;
    if (rows <= 0 || cols <= 0 || (cols % 4) != 0 || (ldx % 4) != 0 || (ldo % 4) != 0 ||
        out_kind < 0 || out_kind > 2 || (out_lo != nullptr && out_kind != 0))
        return SB_ERR_BAD_ARG;
    uint16_t* lo = static_cast<uint16_t*>(out_lo);
    const int threads = 256;
    const int blocks = (rows * 32 + threads - 1) / threads;
    ProfScope prof(PROF_ROWOP, static_cast<double>(rows) * cols * (4.0 + (out_kind == 2 ? 4.0 : 2.0)), stream);
    if (cols <= 512)
        layernorm_kernel<4><<<blocks, threads, 0, stream>>>(x, ldx, w, b, out, lo, ldo, rows, cols, cols_real, eps, out_kind);
    else if (cols <= 1024)
        layernorm_kernel<8><<<blocks, threads, 0, stream>>>(x, ldx, w, b, out, lo, ldo, rows, cols, cols_real, eps, out_kind);
    else if (cols <= 2048)
        layernorm_kernel<16><<<blocks, threads, 0, stream>>>(x, ldx, w, b, out, lo, ldo, rows, cols, cols_real, eps, out_kind);
    else
        return SB_ERR_UNSUPPORTED;
    count_launch();
    return cudaGetLastError() == cudaSuccess ? SB_OK : SB_ERR_CUDA;
}

int fill_rows(float* x, long long ldx, int groups, int rows_per_group, int row_off,
              const float* src, long long lds, const float* add, long long lda, int nrows, int cols,
              cudaStream_t stream) {
    if (groups <= 0 || nrows <= 0 || (cols % 4) != 0) return SB_ERR_BAD_ARG;
    const long long total = static_cast<long long>(groups) * nrows * (cols / 4);
    fill_rows_kernel<<<grid_for(total, 256), 256, 0, stream>>>(x, ldx, groups, rows_per_group,
                                                              row_off, src, lds, add, lda, nrows, cols);
    count_launch();
    return cudaGetLastError() == cudaSuccess ? SB_OK : SB_ERR_CUDA;
}

int tiles_to_patches(const uint8_t* tiles, void* patches, int B, int img, int P, int Kpad,
                     const float mean[3], const float stdv[3], int bf16, cudaStream_t stream) {
    if (B <= 0 || img <= 0 || P <= 0 || (img % P) != 0 || Kpad < 3 * P * P) return SB_ERR_BAD_ARG;
    float3 scale = make_float3(1.f / (255.f * stdv[0]), 1.f / (255.f * stdv[1]), 1.f / (255.f * stdv[2]));
    float3 shift = make_float3(-mean[0] / stdv[0], -mean[1] / stdv[1], -mean[2] / stdv[2]);
    const int gp = img / P;
    if ((P % 8) == 0 && (img % 8) == 0 && Kpad == 3 * P * P &&
        (reinterpret_cast<uintptr_t>(tiles) & 7) == 0) {
        const long long total = static_cast<long long>(B) * gp * gp * P * (P / 8);
        tiles_to_patches_kernel<<<grid_for(total, 256), 256, 0, stream>>>(
            tiles, reinterpret_cast<uint16_t*>(patches), B, img, P, scale, shift, bf16);
    count_launch();
    } else {
        const long long total = static_cast<long long>(B) * gp * gp * Kpad;
        tiles_to_patches_generic_kernel<<<grid_for(total, 256), 256, 0, stream>>>(
            tiles, reinterpret_cast<uint16_t*>(patches), B, img, P, Kpad, scale, shift, bf16);
    count_launch();
    }
    return cudaGetLastError() == cudaSuccess ? SB_OK : SB_ERR_CUDA;
}

int cast_f32_f16(const float* in, __half* hi, __half* lo, long long n, cudaStream_t stream) {
    if (in == nullptr || hi == nullptr || n <= 0 || (reinterpret_cast<uintptr_t>(in) & 15) != 0 ||
        (reinterpret_cast<uintptr_t>(hi) & 15) != 0)
        return SB_ERR_BAD_ARG;
    ProfScope prof(PROF_ROWOP, n * (lo != nullptr ? 8.0 : 6.0), stream);
    cast_split_kernel<<<grid_for(n / 8, 256), 256, 0, stream>>>(in, hi, lo, n / 8, n);
    count_launch();
    return cudaGetLastError() == cudaSuccess ? SB_OK : SB_ERR_CUDA;
}

}  // namespace sb
