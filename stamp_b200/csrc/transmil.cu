// fp32 building blocks of the TransMIL aggregator's Nystrom attention and its position layer.
//
// replaces (inference): NystromAttention.forward and moore_penrose_iter_pinv, src/stamp/modeling/models/trans_mil.py:
// 25-160 (landmark means, the three softmax kernels, the iterative pseudo-inverse, (attn1 @ pinv) @ (attn3 @ v), the
// depth-wise residual convolution over the tokens) and PPEG.forward :253-273 (three depth-wise 2-D convolutions + identity).
// The dense layers around them (_fc1, to_qkv, to_out) run on the tcgen05 GEMM (gemm.cu); everything here is the
// small-matrix side -- 256 landmarks, 64-wide heads, a few GFLOP per slide -- whose results go through a truncated
// Newton-Schulz pseudo-inverse that amplifies operand rounding, so it stays in fp32 on the CUDA cores.
//
// Layouts: token-major fp32 matrices [rows, ld]; head h of a [rows, heads * 64] matrix = columns h*64 .. h*64+63.
#include <cuda_runtime.h>
#include <math.h>

#include <cstdint>

#include "common.cuh"
#include "profile.cuh"
#include "stamp_b200.h"

namespace sb {
namespace {

// out[j, c] = (1 / group) * sum_{t < group} x[(j * group + t), c]
__global__ void __launch_bounds__(256)
landmark_mean_kernel(const float* __restrict__ x, long long ldx, int group, float* __restrict__ out, long long ldo,
                     int n_landmarks, int cols) {
    const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (i >= static_cast<long long>(n_landmarks) * cols) return;
    const int c = static_cast<int>(i % cols);
    const long long j = i / cols;
    float acc = 0.f;
    for (int t = 0; t < group; ++t) acc += x[(j * group + t) * ldx + c];
    out[j * ldo + c] = acc / static_cast<float>(group);
}

// C[b] (+)= alpha * A[b] @ op(B[b]) + bias,  op(B) = B^T (trans_b) or B or (eye * I - B); 64 x 64 tile, 256 threads, 4 x 4 per thread
constexpr int SG_T = 64, SG_K = 16, SG_LD = SG_T + 4;       // row stride 272 B: 128-bit shared-memory reads stay aligned
// The next K-slice is fetched into registers while the current one is multiplied out of shared memory: the small products
// of the pseudo-inverse run one CTA per SM (8 heads x 16 tiles), where nothing else hides the load latency.
template <bool TRANS_B>
__global__ void __launch_bounds__(256)
sgemm_batched_kernel(const float* __restrict__ A, long long lda, long long sa, const float* __restrict__ B, long long ldb,
                     long long sb_, float* __restrict__ C, long long ldc, long long sc, int M, int N, int K,
                     float alpha, float eye, const float* __restrict__ bias, int accumulate) {
    __shared__ __align__(16) float As[SG_K][SG_LD];
    __shared__ __align__(16) float Bs[SG_K][SG_LD];
    const float* a = A + blockIdx.z * sa;
    const float* b = B + blockIdx.z * sb_;
    float* c = C + blockIdx.z * sc;
    const int m0 = blockIdx.y * SG_T, n0 = blockIdx.x * SG_T;
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    // row-major operands (A, and B when the product is A B^T): thread -> 16 consecutive k of one row (64-byte segments);
    // B of A B: thread -> 64 consecutive columns of one row of B (coalesced)
    const int a_kk = tid & 15, a_r = tid >> 4;           // + 16 j
    const int b_kk = TRANS_B ? a_kk : tid >> 6;          // + 4 j when not transposed
    const int b_r = TRANS_B ? a_r : tid & 63;
    float ra[4], rb[4];
    auto fetch = [&](int k0) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int m = m0 + a_r + 16 * j, k = k0 + a_kk;
            ra[j] = (m < M && k < K) ? a[static_cast<long long>(m) * lda + k] : 0.f;
            if (TRANS_B) {
                const int n = n0 + b_r + 16 * j;
                rb[j] = (n < N && k < K) ? b[static_cast<long long>(n) * ldb + k] : 0.f;
            } else {
                const int n = n0 + b_r, kb = k0 + b_kk + 4 * j;
                float v = 0.f;
                if (n < N && kb < K) {
                    v = b[static_cast<long long>(kb) * ldb + n];
                    if (eye != 0.f) v = (kb == n ? eye : 0.f) - v;
                }
                rb[j] = v;
            }
        }
    };
    float acc[4][4] = {};
    fetch(0);
    for (int k0 = 0; k0 < K; k0 += SG_K) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            As[a_kk][a_r + 16 * j] = ra[j];
            if (TRANS_B) Bs[b_kk][b_r + 16 * j] = rb[j];
            else Bs[b_kk + 4 * j][b_r] = rb[j];
        }
        __syncthreads();
        if (k0 + SG_K < K) fetch(k0 + SG_K);
#pragma unroll
        for (int kk = 0; kk < SG_K; ++kk) {
            const float4 a4 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
            const float4 b4 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
            const float av[4] = {a4.x, a4.y, a4.z, a4.w}, bv[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int m = m0 + ty * 4 + i, n = n0 + tx * 4 + j;
            if (m < M && n < N) {
                float r = alpha * acc[i][j] + (bias != nullptr ? __ldg(bias + n) : 0.f);
                if (accumulate & 2) r = fmaxf(r, 0.f);          // bit 1: ReLU
                float* dst = c + static_cast<long long>(m) * ldc + n;
                *dst = (accumulate & 1) ? *dst + r : r;         // bit 0: C +=
            }
        }
}

// in-place softmax over the rows of [batch][rows, cols] (one warp per row)
__global__ void __launch_bounds__(256)
softmax_rows_kernel(float* __restrict__ x, long long ld, long long stride, int rows, int cols, int batch) {
    const long long w = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= static_cast<long long>(rows) * batch) return;
    float* r = x + (w / rows) * stride + (w % rows) * ld;
    float mx = -INFINITY;
    for (int c = lane; c < cols; c += 32) mx = fmaxf(mx, r[c]);
    mx = warp_max(mx);
    float s = 0.f;
    for (int c = lane; c < cols; c += 32) { const float e = expf(r[c] - mx); r[c] = e; s += e; }
    s = warp_sum(s);
    const float inv = 1.0f / s;
    for (int c = lane; c < cols; c += 32) r[c] *= inv;
}

// moore_penrose_iter_pinv's start: z = x^T / (max_i sum_j |x_ij| * max_j sum_i |x_ij|), the maxima over ALL matrices
// of the batch (torch.max over the whole tensor, trans_mil.py:31-35).  Pass 1: abs row / column sums -> two maxima
// (positive floats compare like their bit patterns); pass 2: transpose and scale.
__global__ void __launch_bounds__(256)
pinv_norms_kernel(const float* __restrict__ x, int n, int batch, unsigned int* __restrict__ maxima) {
    const long long w = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= 2LL * n * batch) return;
    const bool column = w >= static_cast<long long>(n) * batch;
    const long long v = column ? w - static_cast<long long>(n) * batch : w;
    const float* m = x + (v / n) * n * n;
    const int i = static_cast<int>(v % n);
    float s = 0.f;
    for (int j = lane; j < n; j += 32) s += fabsf(column ? m[static_cast<long long>(j) * n + i] : m[static_cast<long long>(i) * n + j]);
    s = warp_sum(s);
    if (lane == 0) atomicMax(maxima + (column ? 1 : 0), __float_as_uint(s));
}
__global__ void __launch_bounds__(256)
pinv_init_kernel(const float* __restrict__ x, float* __restrict__ z, int n, int batch, const unsigned int* __restrict__ maxima) {
    const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (i >= static_cast<long long>(batch) * n * n) return;
    const float denom = __uint_as_float(maxima[0]) * __uint_as_float(maxima[1]);
    const long long b = i / (static_cast<long long>(n) * n);
    const int r = static_cast<int>((i / n) % n), c = static_cast<int>(i % n);
    z[i] = x[b * n * n + static_cast<long long>(c) * n + r] / denom;
}

// O = softmax(scale * Q K^T) V per head (head dimension 64), fp32, online softmax: one thread per query row, 64 rows per CTA,
// key tiles of 32 rows staged in shared memory.  gridDim.z > 1 splits the keys (few queries over many keys: the landmark
// rows of attn3 @ v): CTA z takes `keys_per_split` keys and leaves its unnormalised accumulator, row maximum and row sum
// in part[((z * heads + h) * nq + row) * 66 ..], attention_f32_combine_kernel merges them.
constexpr int FA_Q = 64, FA_K = 32, FA_PART = 66;
__global__ void __launch_bounds__(FA_Q)
attention_f32_kernel(const float* __restrict__ Q, long long ldq, const float* __restrict__ K, long long ldk,
                     const float* __restrict__ V, long long ldv, float* __restrict__ O, long long ldo, int nq, int nk_all,
                     float scale, int keys_per_split, float* __restrict__ part) {
    __shared__ float Ks[FA_K][64];
    __shared__ float Vs[FA_K][64];
    const int h = blockIdx.y;
    const int row = blockIdx.x * FA_Q + threadIdx.x;
    float q[64], o[64];
    const float* qp = Q + static_cast<long long>(min(row, nq - 1)) * ldq + h * 64;
#pragma unroll
    for (int d = 0; d < 64; ++d) { q[d] = qp[d] * scale; o[d] = 0.f; }
    float mx = -INFINITY, l = 0.f;
    const int nk = min(nk_all, static_cast<int>(blockIdx.z + 1) * keys_per_split);
    for (int k0 = blockIdx.z * keys_per_split; k0 < nk; k0 += FA_K) {
        for (int i = threadIdx.x; i < FA_K * 64; i += FA_Q) {
            const int r = i >> 6, d = i & 63;
            const bool ok = k0 + r < nk;
            Ks[r][d] = ok ? K[static_cast<long long>(k0 + r) * ldk + h * 64 + d] : 0.f;
            Vs[r][d] = ok ? V[static_cast<long long>(k0 + r) * ldv + h * 64 + d] : 0.f;
        }
        __syncthreads();
        const int nv = min(FA_K, nk - k0);
        float s[FA_K];
        float tmx = mx;
#pragma unroll
        for (int r = 0; r < FA_K; ++r) {
            float a = 0.f;
#pragma unroll
            for (int d = 0; d < 64; ++d) a = fmaf(q[d], Ks[r][d], a);
            s[r] = (r < nv) ? a : -INFINITY;
            tmx = fmaxf(tmx, s[r]);
        }
        const float f = (mx == -INFINITY) ? 0.f : expf(mx - tmx);
        l *= f;
#pragma unroll
        for (int d = 0; d < 64; ++d) o[d] *= f;
#pragma unroll
        for (int r = 0; r < FA_K; ++r) {
            const float p = (r < nv) ? expf(s[r] - tmx) : 0.f;
            l += p;
#pragma unroll
            for (int d = 0; d < 64; ++d) o[d] = fmaf(p, Vs[r][d], o[d]);
        }
        mx = tmx;
        __syncthreads();
    }
    if (row >= nq) return;
    if (gridDim.z == 1) {
        const float inv = 1.0f / l;
        float* op = O + static_cast<long long>(row) * ldo + h * 64;
#pragma unroll
        for (int d = 0; d < 64; ++d) op[d] = o[d] * inv;
    } else {
        float* pp = part + ((static_cast<long long>(blockIdx.z) * gridDim.y + h) * nq + row) * FA_PART;
#pragma unroll
        for (int d = 0; d < 64; ++d) pp[d] = o[d];
        pp[64] = mx;
        pp[65] = l;
    }
}

// one warp per (head, query row): O = sum_z o_z e^(m_z - M) / sum_z l_z e^(m_z - M), M = max_z m_z
__global__ void __launch_bounds__(256)
attention_f32_combine_kernel(const float* __restrict__ part, int splits, int heads, int nq, float* __restrict__ O, long long ldo) {
    const long long w = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= static_cast<long long>(heads) * nq) return;
    const int h = static_cast<int>(w / nq), row = static_cast<int>(w % nq);
    float M = -INFINITY;
    for (int z = 0; z < splits; ++z) M = fmaxf(M, part[((static_cast<long long>(z) * heads + h) * nq + row) * FA_PART + 64]);
    float L = 0.f, o0 = 0.f, o1 = 0.f;
    for (int z = 0; z < splits; ++z) {
        const float* pp = part + ((static_cast<long long>(z) * heads + h) * nq + row) * FA_PART;
        const float f = (pp[64] == -INFINITY) ? 0.f : expf(pp[64] - M);
        L = fmaf(pp[65], f, L);
        o0 = fmaf(pp[lane], f, o0);
        o1 = fmaf(pp[lane + 32], f, o1);
    }
    float* op = O + static_cast<long long>(row) * ldo + h * 64;
    op[lane] = o0 / L;
    op[lane + 32] = o1 / L;
}

// out[i, h*64 + d] += sum_t w[h, t] * v[i + t - taps/2, h*64 + d]   (zero padding; nn.Conv2d(heads, heads, (taps, 1), groups=heads))
__global__ void __launch_bounds__(256)
dwconv1d_add_kernel(const float* __restrict__ v, long long ldv, const float* __restrict__ w, float* __restrict__ out,
                    long long ldo, int n, int heads, int taps) {
    const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    const int cols = heads * 64;
    if (i >= static_cast<long long>(n) * cols) return;
    const int c = static_cast<int>(i % cols);
    const int r = static_cast<int>(i / cols);
    const float* wh = w + (c >> 6) * taps;
    float acc = 0.f;
    for (int t = 0; t < taps; ++t) {
        const int rr = r + t - taps / 2;
        if (rr >= 0 && rr < n) acc = fmaf(__ldg(wh + t), v[static_cast<long long>(rr) * ldv + c], acc);
    }
    out[static_cast<long long>(r) * ldo + c] += acc;
}

// out[(y, x), c] = bias[c] + sum_{dy, dx} k[c, dy, dx] * in[(y + dy - R, x + dx - R), c]   (zero padding), tokens on an H x W grid
__global__ void __launch_bounds__(256)
dwconv2d_kernel(const float* __restrict__ in, long long ldi, const float* __restrict__ k, const float* __restrict__ bias,
                float* __restrict__ out, long long ldo, int H, int W, int C, int ksize) {
    const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (i >= static_cast<long long>(H) * W * C) return;
    const int c = static_cast<int>(i % C);
    const int p = static_cast<int>(i / C);
    const int y = p / W, x = p % W, R = ksize / 2;
    const float* kc = k + static_cast<long long>(c) * ksize * ksize;
    float acc = __ldg(bias + c);
    for (int dy = 0; dy < ksize; ++dy) {
        const int yy = y + dy - R;
        if (yy < 0 || yy >= H) continue;
        for (int dx = 0; dx < ksize; ++dx) {
            const int xx = x + dx - R;
            if (xx < 0 || xx >= W) continue;
            acc = fmaf(__ldg(kc + dy * ksize + dx), in[(static_cast<long long>(yy) * W + xx) * ldi + c], acc);
        }
    }
    out[static_cast<long long>(p) * ldo + c] = acc;
}

inline unsigned blocks_for(long long n, int per) { return static_cast<unsigned>((n + per - 1) / per); }

}  // namespace
}  // namespace sb

extern "C" {

int stamp_landmark_mean_f32(const float* x, long long ldx, int group, float* out, long long ldo, int n_landmarks, int cols,
                            void* stream) {
    using namespace sb;
    if (x == nullptr || out == nullptr || group <= 0 || n_landmarks <= 0 || cols <= 0) return SB_ERR_BAD_ARG;
    landmark_mean_kernel<<<blocks_for(static_cast<long long>(n_landmarks) * cols, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        x, ldx, group, out, ldo, n_landmarks, cols);
    count_launch();
    return cudaGetLastError() == cudaSuccess ? SB_OK : SB_ERR_CUDA;
}

int stamp_sgemm_batched_f32(const float* A, long long lda, long long stride_a, const float* B, long long ldb, long long stride_b,
                            float* C, long long ldc, long long stride_c, int M, int N, int K, int batch, int trans_b, float alpha,
                            float eye_minus_b, const float* bias, int accumulate, void* stream) {
    using namespace sb;
    if (A == nullptr || B == nullptr || C == nullptr || M <= 0 || N <= 0 || K <= 0 || batch <= 0 || batch > 65535 ||
        (eye_minus_b != 0.f && (trans_b || K != N)))
        return SB_ERR_BAD_ARG;
    const dim3 grid((N + SG_T - 1) / SG_T, (M + SG_T - 1) / SG_T, batch);
    if (trans_b)
        sgemm_batched_kernel<true><<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(A, lda, stride_a, B, ldb, stride_b, C, ldc, stride_c,
                                                                                       M, N, K, alpha, eye_minus_b, bias, accumulate);
    else
        sgemm_batched_kernel<false><<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(A, lda, stride_a, B, ldb, stride_b, C, ldc, stride_c,
                                                                                        M, N, K, alpha, eye_minus_b, bias, accumulate);
    count_launch();
    return cudaGetLastError() == cudaSuccess ? SB_OK : SB_ERR_CUDA;
}

int stamp_softmax_rows_f32(float* x, long long ld, long long stride, int rows, int cols, int batch, void* stream) {
    using namespace sb;
    if (x == nullptr || rows <= 0 || cols <= 0 || batch <= 0) return SB_ERR_BAD_ARG;
    softmax_rows_kernel<<<blocks_for(static_cast<long long>(rows) * batch * 32, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        x, ld, stride, rows, cols, batch);
    count_launch();
    return cudaGetLastError() == cudaSuccess ? SB_OK : SB_ERR_CUDA;
}

int stamp_pinv_init_f32(const float* x, float* z, int n, int batch, unsigned int* scratch2, void* stream_) {
    using namespace sb;
    if (x == nullptr || z == nullptr || scratch2 == nullptr || n <= 0 || batch <= 0) return SB_ERR_BAD_ARG;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (cudaMemsetAsync(scratch2, 0, 2 * sizeof(unsigned int), stream) != cudaSuccess) return SB_ERR_CUDA;
    pinv_norms_kernel<<<blocks_for(2LL * n * batch * 32, 256), 256, 0, stream>>>(x, n, batch, scratch2);
    pinv_init_kernel<<<blocks_for(static_cast<long long>(batch) * n * n, 256), 256, 0, stream>>>(x, z, n, batch, scratch2);
    count_launch(2);
    return cudaGetLastError() == cudaSuccess ? SB_OK : SB_ERR_CUDA;
}

int stamp_attention_f32(const float* Q, long long ldq, const float* K, long long ldk, const float* V, long long ldv, float* O,
                        long long ldo, int nq, int nk, int heads, float scale, float* scratch, int splits, void* stream_) {
    using namespace sb;
    if (Q == nullptr || K == nullptr || V == nullptr || O == nullptr || nq <= 0 || nk <= 0 || heads <= 0 || heads > 65535 ||
        splits <= 0 || splits > 65535 || (splits > 1 && scratch == nullptr))
        return SB_ERR_BAD_ARG;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const int tiles = (nk + FA_K - 1) / FA_K;
    if (splits > tiles) splits = tiles;
    const int keys_per_split = ((tiles + splits - 1) / splits) * FA_K;
    splits = (nk + keys_per_split - 1) / keys_per_split;             // no empty split
    const dim3 grid((nq + FA_Q - 1) / FA_Q, heads, splits);
    attention_f32_kernel<<<grid, FA_Q, 0, stream>>>(Q, ldq, K, ldk, V, ldv, O, ldo, nq, nk, scale, keys_per_split, scratch);
    count_launch();
    if (splits > 1) {
        attention_f32_combine_kernel<<<blocks_for(static_cast<long long>(heads) * nq * 32, 256), 256, 0, stream>>>(scratch, splits, heads,
                                                                                                                  nq, O, ldo);
        count_launch();
    }
    return cudaGetLastError() == cudaSuccess ? SB_OK : SB_ERR_CUDA;
}

int stamp_dwconv1d_add_f32(const float* v, long long ldv, const float* w, float* out, long long ldo, int n, int heads, int taps,
                           void* stream) {
    using namespace sb;
    if (v == nullptr || w == nullptr || out == nullptr || n <= 0 || heads <= 0 || taps <= 0 || (taps & 1) == 0) return SB_ERR_BAD_ARG;
    dwconv1d_add_kernel<<<blocks_for(static_cast<long long>(n) * heads * 64, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        v, ldv, w, out, ldo, n, heads, taps);
    count_launch();
    return cudaGetLastError() == cudaSuccess ? SB_OK : SB_ERR_CUDA;
}

int stamp_dwconv2d_f32(const float* in, long long ldi, const float* k, const float* bias, float* out, long long ldo, int H, int W,
                       int C, int ksize, void* stream) {
    using namespace sb;
    if (in == nullptr || k == nullptr || bias == nullptr || out == nullptr || H <= 0 || W <= 0 || C <= 0 || ksize <= 0 ||
        (ksize & 1) == 0)
        return SB_ERR_BAD_ARG;
    dwconv2d_kernel<<<blocks_for(static_cast<long long>(H) * W * C, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        in, ldi, k, bias, out, ldo, H, W, C, ksize);
    count_launch();
    return cudaGetLastError() == cudaSuccess ? SB_OK : SB_ERR_CUDA;
}

}  // extern "C"
