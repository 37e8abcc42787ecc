// Slide-level pooling on the device: CHIEF's gated-attention MIL pooling, exact top-k, and the
// EAGLE "mean of the top-k tiles" gather.
//
// replaces: CHIEFModel.forward + Attn_Net_Gated.forward, src/stamp/encoding/encoder/chief.py:74-89,
//   :255-275 (h = ReLU(W1 x); A = Wc (tanh(Wa h) * sigmoid(Wb h)); A = softmax over tiles;
//   slide embedding = A @ x_original), Eagle._generate_slide_embedding's torch.topk + mean,
//   src/stamp/encoding/encoder/eagle.py:104-120, and scores.topk in
//   src/stamp/heatmaps/__init__.py:216-229.
//
// The attention scores feed a top-k whose indices must be identical to the fp32 reference, so the
// score-producing chain runs in split precision: every dense layer is three tcgen05 passes over
// fp16 (hi, lo) operand pairs with fp32 accumulation (hi.hi + lo.hi + hi.lo), which is fp32-grade.
// The pooling itself is HBM-bound: one pass over the fp32 tile features (N x D x 4 bytes) with
// 128-bit loads, warp-shuffle softmax statistics, deterministic two-stage column sums.
#include <math.h>

#include "common.cuh"
#include "gemm.cuh"
#include "rowops.cuh"
#include "stamp_b200.h"

namespace sb {
namespace {

constexpr int POOL_ROWS_PER_CTA = 128;

// score[n] = dot(a_hi[n,:] + a_lo[n,:], w) + b     one warp per row
__global__ void __launch_bounds__(256)
rowdot_kernel(const __half* __restrict__ a_hi, const __half* __restrict__ a_lo, const float* __restrict__ w,
              float bias, int N, int D, float* __restrict__ score) {
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= N) return;
    float acc = 0.f;
    for (int i = lane * 2; i < D; i += 64) {
        const float2 h = __half22float2(*reinterpret_cast<const __half2*>(a_hi + static_cast<long long>(row) * D + i));
        float2 v = h;
        if (a_lo != nullptr) {
            const float2 l = __half22float2(*reinterpret_cast<const __half2*>(a_lo + static_cast<long long>(row) * D + i));
            v.x += l.x; v.y += l.y;
        }
        acc = fmaf(v.x, __ldg(w + i), acc);
        acc = fmaf(v.y, __ldg(w + i + 1), acc);
    }
    acc = warp_sum(acc);
    if (lane == 0) score[row] = acc + bias;
}

// stats[0] = max_n score, stats[1] = sum_n exp(score - max)      single CTA
__global__ void __launch_bounds__(1024)
softmax_stats_kernel(const float* __restrict__ score, int N, float* __restrict__ stats) {
    __shared__ float red[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float m = -INFINITY;
    for (int i = threadIdx.x; i < N; i += blockDim.x) m = fmaxf(m, score[i]);
    m = warp_max(m);
    if (lane == 0) red[warp] = m;
    __syncthreads();
    m = red[0];
    for (int i = 1; i < 32; ++i) m = fmaxf(m, red[i]);
    __syncthreads();
    float s = 0.f;
    for (int i = threadIdx.x; i < N; i += blockDim.x) s += expf(score[i] - m);
    s = warp_sum(s);
    if (lane == 0) red[warp] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int i = 0; i < 32; ++i) t += red[i];
        stats[0] = m;
        stats[1] = t;
    }
}

// partial[cta, :] = sum over this CTA's rows of softmax(score)[n] * x[n, :]
__global__ void __launch_bounds__(256)
weighted_pool_kernel(const float* __restrict__ x, const float* __restrict__ score,
                     const float* __restrict__ stats, int N, int D, float* __restrict__ partial) {
    const int r0 = blockIdx.x * POOL_ROWS_PER_CTA;
    const int r1 = min(N, r0 + POOL_ROWS_PER_CTA);
    const float m = stats[0], inv = 1.0f / stats[1];
    const int nvec = D >> 2;
    for (int c = threadIdx.x; c < nvec; c += blockDim.x) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int r = r0; r < r1; ++r) {
            const float wgt = expf(__ldg(score + r) - m) * inv;
            const uint4 u = ld_nc_v4(x + static_cast<long long>(r) * D + c * 4);
            acc.x = fmaf(wgt, __uint_as_float(u.x), acc.x);
            acc.y = fmaf(wgt, __uint_as_float(u.y), acc.y);
            acc.z = fmaf(wgt, __uint_as_float(u.z), acc.z);
            acc.w = fmaf(wgt, __uint_as_float(u.w), acc.w);
        }
        *reinterpret_cast<float4*>(partial + static_cast<long long>(blockIdx.x) * D + c * 4) = acc;
    }
}

// out[c] = sum_p partial[p, c] (fixed order -> deterministic)
__global__ void __launch_bounds__(256)
column_sum_kernel(const float* __restrict__ partial, int P, int D, float scale, float* __restrict__ out) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= D) return;
    float acc = 0.f;
    for (int p = 0; p < P; ++p) acc += partial[static_cast<long long>(p) * D + c];
    out[c] = acc * scale;
}

// out[c] = mean over j < k of feats[idx[j], c]
__global__ void __launch_bounds__(256)
gather_mean_kernel(const float* __restrict__ feats, const long long* __restrict__ idx, int k, int D,
                   float* __restrict__ out) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= D) return;
    float acc = 0.f;
    for (int j = 0; j < k; ++j) acc += __ldg(feats + idx[j] * D + c);
    out[c] = acc / static_cast<float>(k);
}

__device__ __forceinline__ unsigned int fkey(float f) {
    unsigned int u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// Exact top-k of N fp32 scores (largest first; ties broken by the lower index), single CTA:
// 4-pass 8-bit radix select of the k-th largest key, ordered gather, bitonic sort of the k winners.
constexpr int TOPK_THREADS = 1024;
constexpr int TOPK_MAX = 1024;

__global__ void __launch_bounds__(TOPK_THREADS)
topk_kernel(const float* __restrict__ score, int N, int k, int largest, long long* __restrict__ idx_out,
            float* __restrict__ val_out) {
    __shared__ unsigned int hist[256];
    __shared__ unsigned int s_prefix, s_rank, s_count;
    __shared__ unsigned int keys[TOPK_MAX];
    __shared__ int ids[TOPK_MAX];
    __shared__ unsigned int scan[TOPK_THREADS / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // smallest-k is top-k of the negated scores (heatmaps use (-scores).topk)
    auto key_of = [&](int i) { const float v = score[i]; return fkey(largest ? v : -v); };

    if (tid == 0) { s_prefix = 0; s_rank = static_cast<unsigned int>(k - 1); }  // rank among descending keys
    __syncthreads();
    for (int pass = 0; pass < 4; ++pass) {
        const int shift = 24 - 8 * pass;
        for (int i = tid; i < 256; i += blockDim.x) hist[i] = 0;
        __syncthreads();
        const unsigned int prefix = s_prefix;
        for (int i = tid; i < N; i += blockDim.x) {
            const unsigned int kk = key_of(i);
            if (pass == 0 || (kk >> (shift + 8)) == prefix) atomicAdd(&hist[(kk >> shift) & 255u], 1u);
        }
        __syncthreads();
        if (tid == 0) {
            unsigned int r = s_rank, cum = 0;
            int b = 255;
            for (; b >= 0; --b) {  // descending: largest digits first
                if (r < cum + hist[b]) break;
                cum += hist[b];
            }
            if (b < 0) b = 0;
            s_rank = r - cum;
            s_prefix = (prefix << 8) | static_cast<unsigned int>(b);
        }
        __syncthreads();
    }
    const unsigned int T = s_prefix;  // key of the k-th largest element
    // gather: everything above T in any order, then elements equal to T in index order
    if (tid == 0) s_count = 0;
    __syncthreads();
    for (int i = tid; i < N; i += blockDim.x) {
        const unsigned int kk = key_of(i);
        if (kk > T) {
            const unsigned int slot = atomicAdd(&s_count, 1u);
            if (slot < static_cast<unsigned int>(k)) { keys[slot] = kk; ids[slot] = i; }
        }
    }
    __syncthreads();
    unsigned int filled = min(s_count, static_cast<unsigned int>(k));
    for (int base = 0; base < N && filled < static_cast<unsigned int>(k); base += blockDim.x) {
        const int i = base + tid;
        const bool eq = (i < N) && (key_of(i) == T);
        const unsigned int bal = __ballot_sync(0xffffffffu, eq);
        if (lane == 0) scan[warp] = __popc(bal);
        __syncthreads();
        unsigned int before = 0, total = 0;
        for (int w = 0; w < TOPK_THREADS / 32; ++w) { if (w < warp) before += scan[w]; total += scan[w]; }
        if (eq) {
            const unsigned int slot = filled + before + __popc(bal & ((1u << lane) - 1u));
            if (slot < static_cast<unsigned int>(k)) { keys[slot] = T; ids[slot] = i; }
        }
        filled = min(static_cast<unsigned int>(k), filled + total);
        __syncthreads();
    }
    // pad to a power of two and bitonic-sort by (key desc, index asc)
    int n2 = 1;
    while (n2 < k) n2 <<= 1;
    for (int i = k + tid; i < n2; i += blockDim.x) { keys[i] = 0; ids[i] = 0x7FFFFFFF; }
    __syncthreads();
    for (int size = 2; size <= n2; size <<= 1)
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int i = tid; i < n2; i += blockDim.x) {
                const int j = i ^ stride;
                if (j > i) {
                    const bool desc = (i & size) == 0;
                    const bool i_first = (keys[i] > keys[j]) || (keys[i] == keys[j] && ids[i] < ids[j]);
                    if (i_first != desc) {
                        const unsigned int tk = keys[i]; keys[i] = keys[j]; keys[j] = tk;
                        const int ti = ids[i]; ids[i] = ids[j]; ids[j] = ti;
                    }
                }
            }
            __syncthreads();
        }
    for (int i = tid; i < k; i += blockDim.x) {
        idx_out[i] = ids[i];
        if (val_out != nullptr) val_out[i] = score[ids[i]];
    }
}

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct ChiefLayout {
    size_t off_xhi, off_xlo, off_acc, off_hhi, off_hlo, off_ghi, off_glo, off_stats, off_partial, total;
    int P;
};

bool chief_layout(int N, int D, int L, int Dh, ChiefLayout* c) {
    if (N <= 0 || D <= 0 || D % 8 || L % 8 || Dh % 8) return false;
    size_t o = 0;
    const size_t n = N;
    c->off_xhi = o; o = align_up(o + n * D * 2, 256);
    c->off_xlo = o; o = align_up(o + n * D * 2, 256);
    c->off_acc = o; o = align_up(o + n * (L > 2 * Dh ? L : 2 * Dh) * 4, 256);
    c->off_hhi = o; o = align_up(o + n * L * 2, 256);
    c->off_hlo = o; o = align_up(o + n * L * 2, 256);
    c->off_ghi = o; o = align_up(o + n * Dh * 2, 256);
    c->off_glo = o; o = align_up(o + n * Dh * 2, 256);
    c->off_stats = o; o = align_up(o + 16, 256);
    c->P = (N + POOL_ROWS_PER_CTA - 1) / POOL_ROWS_PER_CTA;
    c->off_partial = o; o = align_up(o + static_cast<size_t>(c->P) * D * 4, 256);
    c->total = o;
    return true;
}

// three-pass split-precision layer: out = epilogue(A.W^T), A = a_hi + a_lo, W = w_hi + w_lo
int split_gemm(const __half* a_hi, const __half* a_lo, const void* w_hi, const void* w_lo, int M, int N,
               int K, const float* bias, float* acc, int act, int store, void* out, void* out_lo,
               long long ldo, cudaStream_t stream) {
    GemmParams p{};
    p.M = M; p.N = N; p.K = K;
    p.store = ST_32; p.out = acc; p.ldo = N; p.bias = bias;
    int rc = gemm_tn(a_hi, K, w_hi, K, p, stream);
    if (rc != SB_OK) return rc;
    p.store = ST_RESID32; p.bias = nullptr;
    rc = gemm_tn(a_lo, K, w_hi, K, p, stream);
    if (rc != SB_OK) return rc;
    p.act = act; p.store = store; p.out = out; p.out_lo = out_lo; p.ldo = ldo;
    p.table = acc; p.ldt = N;
    return gemm_tn(a_hi, K, w_lo, K, p, stream);
}

}  // namespace
}  // namespace sb

extern "C" {

int stamp_topk_f32(const float* scores, int N, int k, int largest, long long* idx_out, float* val_out,
                   void* stream) {
    if (scores == nullptr || idx_out == nullptr || N <= 0 || k <= 0 || k > N || k > sb::TOPK_MAX)
        return sb::SB_ERR_BAD_ARG;
    sb::ProfScope prof(sb::PROF_POOL, 4.0 * N, static_cast<cudaStream_t>(stream));
    sb::topk_kernel<<<1, sb::TOPK_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(scores, N, k, largest,
                                                                                    idx_out, val_out);
    sb::count_launch();
    return cudaGetLastError() == cudaSuccess ? sb::SB_OK : sb::SB_ERR_CUDA;
}

int stamp_gather_mean_f32(const float* feats, const long long* idx, int k, int D, float* out, void* stream) {
    if (feats == nullptr || idx == nullptr || out == nullptr || k <= 0 || D <= 0) return sb::SB_ERR_BAD_ARG;
    sb::gather_mean_kernel<<<(D + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(feats, idx, k, D, out);
    sb::count_launch();
    return cudaGetLastError() == cudaSuccess ? sb::SB_OK : sb::SB_ERR_CUDA;
}

size_t stamp_gated_attn_pool_workspace_bytes(int N, int D, int L, int Dh) {
    sb::ChiefLayout c;
    return sb::chief_layout(N, D, L, Dh, &c) ? c.total : 0;
}

int stamp_gated_attn_pool(const StampGatedAttnWeights* w, const float* x, int N, int D, int L, int Dh,
                          float* attn_raw, float* pooled, void* workspace, size_t workspace_bytes,
                          void* stream_) {
    using namespace sb;
    ChiefLayout c;
    if (w == nullptr || x == nullptr || attn_raw == nullptr || pooled == nullptr || workspace == nullptr ||
        !chief_layout(N, D, L, Dh, &c))
        return SB_ERR_BAD_ARG;
    if (workspace_bytes < c.total) return SB_ERR_WORKSPACE;
    if ((reinterpret_cast<uintptr_t>(workspace) & 255) != 0) return SB_ERR_BAD_ARG;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    uint8_t* ws = static_cast<uint8_t*>(workspace);
    __half* x_hi = reinterpret_cast<__half*>(ws + c.off_xhi);
    __half* x_lo = reinterpret_cast<__half*>(ws + c.off_xlo);
    float* acc = reinterpret_cast<float*>(ws + c.off_acc);
    __half* h_hi = reinterpret_cast<__half*>(ws + c.off_hhi);
    __half* h_lo = reinterpret_cast<__half*>(ws + c.off_hlo);
    __half* g_hi = reinterpret_cast<__half*>(ws + c.off_ghi);
    __half* g_lo = reinterpret_cast<__half*>(ws + c.off_glo);
    float* stats = reinterpret_cast<float*>(ws + c.off_stats);
    float* partial = reinterpret_cast<float*>(ws + c.off_partial);

    int rc = cast_f32_f16(x, x_hi, x_lo, static_cast<long long>(N) * D, stream);
    if (rc != SB_OK) return rc;
    // h = ReLU(W1 x + b1)                                   chief.py:45 (fc) -- dropout inactive in eval
    rc = split_gemm(x_hi, x_lo, w->fc_w_hi, w->fc_w_lo, N, L, D, w->fc_b, acc, ACT_RELU, ST_16, h_hi, h_lo, L, stream);
    if (rc != SB_OK) return rc;
    // g = tanh(Wa h + ba) * sigmoid(Wb h + bb)              chief.py:270-273 (rows of Wa, Wb interleaved)
    rc = split_gemm(h_hi, h_lo, w->ab_w_hi, w->ab_w_lo, N, 2 * Dh, L, w->ab_b, acc, ACT_NONE, ST_GATED16, g_hi, g_lo, Dh, stream);
    if (rc != SB_OK) return rc;
    {
        ProfScope prof(PROF_POOL, static_cast<double>(N) * D * 4.0, stream);  // algorithmic: one read of x
        // A_raw = Wc g + bc                                  chief.py:274
        rowdot_kernel<<<(N * 32 + 255) / 256, 256, 0, stream>>>(g_hi, g_lo, w->c_w, w->c_b, N, Dh, attn_raw);
        // A = softmax over tiles; slide embedding = A @ x    chief.py:78-81
        softmax_stats_kernel<<<1, 1024, 0, stream>>>(attn_raw, N, stats);
        weighted_pool_kernel<<<c.P, 256, 0, stream>>>(x, attn_raw, stats, N, D, partial);
        column_sum_kernel<<<(D + 255) / 256, 256, 0, stream>>>(partial, c.P, D, 1.0f, pooled);
        count_launch(4);
    }
    return cudaGetLastError() == cudaSuccess ? SB_OK : SB_ERR_CUDA;
}

}  // extern "C"
