// Mean over the tiles of a bag: the pooling step of the reference's MLP / Linear aggregators.
//
// replaces: `x.mean(dim=1)` of MLP.forward / Linear.forward, src/stamp/modeling/models/mlp.py:40-43 and :57-60
// ([B, T, F] bag of tile features -> [B, F]); the layers after it are a handful of [B, F] x [F, H] products
// (stamp_sgemm_batched_f32).  HBM-bound: every feature is read once (B * T * F * sizeof(element) bytes), nothing else
// moves.  Layout: 16-byte vectors along the feature axis (4 fp32 / 8 fp16 per thread, a warp reads 512 contiguous
// bytes of one row), the 8 warps of a CTA take every 8th row of their row range, `splits` CTAs share the rows of one
// bag so that the grid fills the 148 SMs whatever B is.  Two deterministic stages (no atomics): per-split partial sums
// in a fixed order, then their sum in split order times 1 / T.
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstdint>

#include "common.cuh"
#include "stamp_b200.h"

namespace sb {
namespace {

template <typename T, int VEC> struct VecLoad;
template <> struct VecLoad<float, 4> {
    static __device__ __forceinline__ void add(const float* p, float (&a)[4]) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(p));
        a[0] += v.x; a[1] += v.y; a[2] += v.z; a[3] += v.w;
    }
};
template <> struct VecLoad<float, 1> {
    static __device__ __forceinline__ void add(const float* p, float (&a)[1]) { a[0] += __ldg(p); }
};
template <> struct VecLoad<__half, 8> {
    static __device__ __forceinline__ void add(const __half* p, float (&a)[8]) {
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(p));
        const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float2 f = __half22float2(h[i]);
            a[2 * i] += f.x; a[2 * i + 1] += f.y;
        }
    }
};
template <> struct VecLoad<__half, 1> {
    static __device__ __forceinline__ void add(const __half* p, float (&a)[1]) { a[0] += __half2float(p[0]); }
};

constexpr int BM_WARPS = 8;
inline unsigned blocks_for(long long n, int per) { return static_cast<unsigned>((n + per - 1) / per); }

// grid (column chunks of 32 * VEC, splits, B); partial[split][b][F] (or out itself, scaled, when splits == 1)
template <typename T, int VEC>
__global__ void __launch_bounds__(BM_WARPS * 32)
bag_mean_partial_kernel(const T* __restrict__ x, long long ldx, long long stride_bag, int n_tiles, int F, int rows_per_split,
                        float* __restrict__ dst, long long ldd, long long stride_split, float scale) {
    __shared__ float red[BM_WARPS][32 * VEC + 1];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c0 = (blockIdx.x * 32 + lane) * VEC;
    const int r0 = blockIdx.y * rows_per_split;
    const int r1 = min(r0 + rows_per_split, n_tiles);
    float acc[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) acc[i] = 0.f;
    if (c0 < F) {
        const T* p = x + blockIdx.z * stride_bag + c0;
        int r = r0 + warp;
        for (; r + 3 * BM_WARPS < r1; r += 4 * BM_WARPS) {          // four rows in flight per thread
            float a1[VEC], a2[VEC], a3[VEC];
#pragma unroll
            for (int i = 0; i < VEC; ++i) a1[i] = a2[i] = a3[i] = 0.f;
            VecLoad<T, VEC>::add(p + r * ldx, acc);
            VecLoad<T, VEC>::add(p + (r + BM_WARPS) * ldx, a1);
            VecLoad<T, VEC>::add(p + (r + 2 * BM_WARPS) * ldx, a2);
            VecLoad<T, VEC>::add(p + (r + 3 * BM_WARPS) * ldx, a3);
#pragma unroll
            for (int i = 0; i < VEC; ++i) acc[i] += (a1[i] + a2[i]) + a3[i];
        }
        for (; r < r1; r += BM_WARPS) VecLoad<T, VEC>::add(p + r * ldx, acc);
    }
#pragma unroll
    for (int i = 0; i < VEC; ++i) red[warp][lane * VEC + i] = acc[i];
    __syncthreads();
    for (int i = threadIdx.x; i < 32 * VEC; i += BM_WARPS * 32) {
        const int c = blockIdx.x * 32 * VEC + i;
        if (c >= F) continue;
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < BM_WARPS; ++w) s += red[w][i];
        dst[blockIdx.y * stride_split + blockIdx.z * ldd + c] = s * scale;
    }
}

__global__ void __launch_bounds__(256)
bag_mean_final_kernel(const float* __restrict__ partial, int splits, long long stride_split, int B, int F, float scale,
                      float* __restrict__ out, long long ldo) {
    const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (i >= static_cast<long long>(B) * F) return;
    float s = 0.f;
    for (int k = 0; k < splits; ++k) s += partial[k * stride_split + i];
    out[(i / F) * ldo + (i % F)] = s * scale;
}

template <typename T, int VEC>
int launch(const T* x, long long ldx, long long stride_bag, int B, int n_tiles, int F, float* out, long long ldo, float* scratch,
           int splits, cudaStream_t stream) {
    const int chunks = (F + 32 * VEC - 1) / (32 * VEC);
    const int rows_per_split = (n_tiles + splits - 1) / splits;
    const dim3 grid(chunks, splits, B);
    const float inv = 1.0f / static_cast<float>(n_tiles);
    if (splits == 1) {
        bag_mean_partial_kernel<T, VEC><<<grid, BM_WARPS * 32, 0, stream>>>(x, ldx, stride_bag, n_tiles, F, rows_per_split, out, ldo, 0, inv);
        count_launch();
    } else {
        const long long ss = static_cast<long long>(B) * F;
        bag_mean_partial_kernel<T, VEC><<<grid, BM_WARPS * 32, 0, stream>>>(x, ldx, stride_bag, n_tiles, F, rows_per_split, scratch, F, ss, 1.0f);
        bag_mean_final_kernel<<<blocks_for(ss, 256), 256, 0, stream>>>(scratch, splits, ss, B, F, inv, out, ldo);
        count_launch(2);
    }
    return cudaGetLastError() == cudaSuccess ? SB_OK : SB_ERR_CUDA;
}

}  // namespace
}  // namespace sb

extern "C" {

int stamp_bag_mean_splits(int B, int n_tiles, int F, int is_half) {
    if (B <= 0 || n_tiles <= 0 || F <= 0) return 1;
    const int vec = is_half ? 8 : 4;
    const long long ctas = static_cast<long long>(B) * ((F + 32 * vec - 1) / (32 * vec));
    const long long want = 148LL * 8;                               // eight resident CTAs per SM
    long long s = (want + ctas - 1) / ctas;
    const long long most = (n_tiles + 4 * sb::BM_WARPS - 1) / (4 * sb::BM_WARPS);   // at least 32 rows per split
    if (s > most) s = most;
    if (s > 65535) s = 65535;
    return s < 1 ? 1 : static_cast<int>(s);
}

int stamp_bag_mean(const void* x, int is_half, long long ldx, long long stride_bag, int B, int n_tiles, int F, float* out,
                   long long ldo, float* scratch, int splits, void* stream_) {
    using namespace sb;
    if (x == nullptr || out == nullptr || B <= 0 || B > 65535 || n_tiles <= 0 || F <= 0 || splits <= 0 || splits > 65535 ||
        (splits > 1 && scratch == nullptr) || ldx < F || ldo < F)
        return SB_ERR_BAD_ARG;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const bool aligned = (reinterpret_cast<uintptr_t>(x) & 15) == 0;
    if (!is_half) {
        const float* p = static_cast<const float*>(x);
        if (aligned && F % 4 == 0 && ldx % 4 == 0 && stride_bag % 4 == 0)
            return launch<float, 4>(p, ldx, stride_bag, B, n_tiles, F, out, ldo, scratch, splits, stream);
        return launch<float, 1>(p, ldx, stride_bag, B, n_tiles, F, out, ldo, scratch, splits, stream);
    }
    const __half* p = static_cast<const __half*>(x);
    if (aligned && F % 8 == 0 && ldx % 8 == 0 && stride_bag % 8 == 0)
        return launch<__half, 8>(p, ldx, stride_bag, B, n_tiles, F, out, ldo, scratch, splits, stream);
    return launch<__half, 1>(p, ldx, stride_bag, B, n_tiles, F, out, ldo, scratch, splits, stream);
}

}  // extern "C"
