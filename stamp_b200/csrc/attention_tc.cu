// tcgen05 attention for short sequences (the ViT tile encoder: T = 197 tokens, head_dim 64).
//
//   O = softmax(Q K^T * scale) V      per (tile, head); no mask, no ALiBi, <= 256 keys
//
// replaces timm Attention.forward's F.scaled_dot_product_attention inside the ViT blocks run by
// src/stamp/preprocessing/__init__.py:325 (restated in oracle/vit_oracle.py: block_forward).
//
// One CTA = 128 query rows of one (tile, head):
//   warp 0 : TMA  -- Q [128 x 64], K [NK x 64], V [NK x 64] tiles (3-D tensor map over
//            [tile][token][3D], zero fill past the last token), 128B swizzle
//   warp 1 : MMA  -- S = Q K^T  (tcgen05.mma M=128, N=NK, 4 x K=16)  -> TMEM columns [0, NK)
//                    O = P V    (M=128, N=64, NK/16 x K=16, V as MN-major B operand) -> TMEM [0, 64)
//   warps 2-5 : one thread per query row (= TMEM lane): row max and exp2 straight from TMEM (no
//            shuffles: the whole row lives in one lane), P written as fp16 directly in the
//            K-major 128B-swizzled UMMA layout, 1/l applied when O is read back.
// Because all keys fit one pass there is no online-softmax rescale of the accumulator.
// Shared memory: the Q and K tiles are dead once S is committed, so P blocks 3 and 0 reuse them:
// Q 16 KB + K 26 KB + P1,P2 32 KB + V 26 KB = 100 KB at T = 197 -> two CTAs per SM, 2 x 256 TMEM
// columns.
#include <math.h>

#include "attention.cuh"
#include "common.cuh"
#include "gemm.cuh"

namespace sb {
namespace {

constexpr int TC_THREADS = 320;  // TMA warp, MMA warp, 8 softmax warps (two threads per query row)
constexpr int P_BLOCK_BYTES = 128 * 128;  // 128 rows x 64 keys fp16

struct TcSmem {
    int off_q, off_k, off_p12, off_v, off_bar, total;
};

inline TcSmem tc_layout(int nk) {
    TcSmem s;
    const int kv_bytes = nk * 128;
    s.off_q = 0;                                   // also P block 3
    s.off_k = P_BLOCK_BYTES;                       // also P block 0
    const int k_region = kv_bytes > P_BLOCK_BYTES ? kv_bytes : P_BLOCK_BYTES;
    s.off_p12 = s.off_k + ((k_region + 1023) / 1024) * 1024;
    s.off_v = s.off_p12 + 2 * P_BLOCK_BYTES;
    s.off_bar = s.off_v + ((kv_bytes + 1023) / 1024) * 1024;
    s.total = s.off_bar + 64 + 1024 + 1024;        // barriers, [2][128] float exchange buffer, alignment slack
    return s;
}

__global__ void __launch_bounds__(TC_THREADS, 2)
vit_attn_tc_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_kv,
                   __half* __restrict__ out, long long out_row_stride, long long out_batch_stride,
                   int S, int H, int D, int nk, float scale_log2, TcSmem L, long long* trace) {
    const long long t_start = clock64();
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>(
        (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    uint8_t* sQ = smem + L.off_q;
    uint8_t* sK = smem + L.off_k;
    uint8_t* sP12 = smem + L.off_p12;
    uint8_t* sV = smem + L.off_v;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L.off_bar);
    uint64_t* bar_load = bars;      // TMA bytes landed
    uint64_t* bar_s = bars + 1;     // S committed to TMEM
    uint64_t* bar_p = bars + 2;     // P written to smem by the 4 softmax warps
    uint64_t* bar_o = bars + 3;     // O committed to TMEM
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // the M-tiles of one (tile, head) are adjacent CTAs: the second read of its K/V hits L2
    const int bh = blockIdx.y;
    const int b = bh / H, h = bh % H;
    const int q0 = blockIdx.x * 128;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tm_q);
        tma_prefetch_desc(&tm_kv);
        mbar_init(bar_load, 1);
        mbar_init(bar_s, 1);
        mbar_init(bar_p, 8);
        mbar_init(bar_o, 1);
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, 256);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            mbar_expect_tx(bar_load, 128 * 128 + 2 * nk * 128);
            tma_load_3d(sQ, &tm_q, bar_load, h * 64, q0, b);            // Q rows q0.. of head h
            tma_load_3d(sK, &tm_kv, bar_load, D + h * 64, 0, b);        // K
            tma_load_3d(sV, &tm_kv, bar_load, 2 * D + h * 64, 0, b);    // V
        }
    } else if (warp == 1) {
        if (lane == 0) {
            mbar_wait(bar_load, 0);
            tc_fence_after();
            {   // S = Q K^T : both operands K-major, N = nk keys
                const uint32_t idesc = umma_idesc_f16(128, nk, false, false, false);
                const uint64_t a = umma_desc_k128(smem_u32(sQ)), bd = umma_desc_k128(smem_u32(sK));
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_f16_ss(tmem, a + 2 * k, bd + 2 * k, idesc, k != 0);
                umma_commit(bar_s);
            }
            mbar_wait(bar_p, 0);
            tc_fence_after();
            {   // O = P V : A = P (K-major, 64-key blocks), B = V rows (MN-major: features contiguous)
                const uint32_t idesc = umma_idesc_f16(128, 64, false, false, true);
                const int steps = nk / 16;
                for (int s = 0; s < steps; ++s) {
                    const int blk = s >> 2;
                    const uint8_t* pb = (blk == 0) ? sK : (blk == 3 ? sQ : sP12 + (blk - 1) * P_BLOCK_BYTES);
                    const uint64_t a = umma_desc_k128(smem_u32(pb)) + 2 * (s & 3);
                    const uint64_t bd = umma_desc_mn128(smem_u32(sV + s * 2048), 0);
                    umma_f16_ss(tmem, a, bd, idesc, s != 0);
                }
                umma_commit(bar_o);
            }
        }
    } else {
        // ---- softmax / epilogue: two threads per query row (= TMEM lane); the 32-key chunks of the
        //      row alternate between them, row statistics are exchanged through shared memory ----
        const int quarter = warp & 3;
        const int half = (warp - 2) >> 2;
        const int r = quarter * 32 + lane;              // row inside the 128-row tile
        const uint32_t t_row = tmem + (static_cast<uint32_t>(quarter * 32) << 16);
        const int nch = (nk + 31) >> 5;                 // 32-key chunks, the last one may hold 16 keys
        // warps whose 32 rows all lie past the last token only keep the barrier protocol alive
        const bool warp_rows_valid = (q0 + quarter * 32) < S;
        float* sX = reinterpret_cast<float*>(smem + L.off_bar + 64);   // [2][128] exchange buffer
        auto load_chunk = [&](uint32_t (&v)[32], int c) {
            if (nk - c * 32 >= 32) tmem_ld_32x32b_x32(t_row + c * 32, v);
            else tmem_ld_32x32b_x16(t_row + c * 32, *reinterpret_cast<uint32_t(*)[16]>(v));
        };
        mbar_wait(bar_s, 0);
        tc_fence_after();
        const long long t_s = clock64();
        // four independent max / sum accumulators: one long dependent chain costs ~4 cycles a link
        float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
        for (int c = half; warp_rows_valid && c < nch; c += 4) {
            uint32_t v0[32], v1[32];
            const bool two = c + 2 < nch;
            load_chunk(v0, c);
            if (two) load_chunk(v1, c + 2);
            tmem_ld_wait();
            const int n0 = min(32, min(nk, S) - c * 32), n1 = min(32, min(nk, S) - (c + 2) * 32);
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                if (j < n0) mx4[j & 3] = fmaxf(mx4[j & 3], __uint_as_float(v0[j]));
                if (two && j < n1) mx4[j & 3] = fmaxf(mx4[j & 3], __uint_as_float(v1[j]));
            }
        }
        float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
        sX[half * 128 + r] = mx;
        asm volatile("bar.sync 1, 256;\n" ::: "memory");
        mx = fmaxf(mx, sX[(half ^ 1) * 128 + r]);
        const float ms = mx * scale_log2;
        const long long t_m = clock64();
        float l4[4] = {0.f, 0.f, 0.f, 0.f};
        // exp2 + fp16 pack + swizzled store of 16 keys (= two 16-byte chunks of row r in 64-key block)
        auto emit16 = [&](const uint32_t* v, int key0) {
            float p[16];
            if (key0 + 16 <= S) {  // (uniform) every key of the group is a real token
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    p[j] = ex2_approx(fmaf(__uint_as_float(v[j]), scale_log2, -ms));
                    l4[j & 3] += p[j];
                }
            } else {
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    p[j] = (key0 + j < S) ? ex2_approx(fmaf(__uint_as_float(v[j]), scale_log2, -ms)) : 0.f;
                    l4[j & 3] += p[j];
                }
            }
            const int blk = key0 >> 6;
            const uint32_t pb = smem_u32((blk == 0) ? sK : (blk == 3 ? sQ : sP12 + (blk - 1) * P_BLOCK_BYTES));
            const int ch0 = (key0 & 63) >> 3;
            uint4 w0, w1;
            w0.x = pack_f16(p[0], p[1]);  w0.y = pack_f16(p[2], p[3]);  w0.z = pack_f16(p[4], p[5]);  w0.w = pack_f16(p[6], p[7]);
            w1.x = pack_f16(p[8], p[9]);  w1.y = pack_f16(p[10], p[11]); w1.z = pack_f16(p[12], p[13]); w1.w = pack_f16(p[14], p[15]);
            sts_v4(pb + r * 128 + ((ch0 ^ (r & 7)) * 16), w0);
            sts_v4(pb + r * 128 + (((ch0 + 1) ^ (r & 7)) * 16), w1);
        };
        for (int c = half; warp_rows_valid && c < nch; c += 4) {
            uint32_t v0[32], v1[32];
            const bool two = c + 2 < nch;
            load_chunk(v0, c);
            if (two) load_chunk(v1, c + 2);
            tmem_ld_wait();
            emit16(v0, c * 32);
            if (nk - c * 32 >= 32) emit16(v0 + 16, c * 32 + 16);
            if (two) {
                emit16(v1, (c + 2) * 32);
                if (nk - (c + 2) * 32 >= 32) emit16(v1 + 16, (c + 2) * 32 + 16);
            }
        }
        // make the generic-proxy smem writes visible to the tensor core (async proxy), free S in TMEM
        fence_proxy_async();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_p);
        const long long t_p = clock64();
        float l = (l4[0] + l4[1]) + (l4[2] + l4[3]);
        asm volatile("bar.sync 1, 256;\n" ::: "memory");   // everyone has read the max exchange
        sX[half * 128 + r] = l;
        asm volatile("bar.sync 1, 256;\n" ::: "memory");
        l += sX[(half ^ 1) * 128 + r];

        mbar_wait(bar_o, 0);
        tc_fence_after();
        const long long t_o = clock64();
        const float inv = 1.0f / l;
        const int row = q0 + r;
        __half* o = out + b * out_batch_stride + static_cast<long long>(row) * out_row_stride + h * 64 + half * 32;
        {
            uint32_t v[32];
            tmem_ld_32x32b_x32(t_row + half * 32, v);   // this thread's 32 of the row's 64 output columns
            tmem_ld_wait();
            if (row < S) {
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    uint4 w;
                    w.x = pack_f16(__uint_as_float(v[8 * c]) * inv, __uint_as_float(v[8 * c + 1]) * inv);
                    w.y = pack_f16(__uint_as_float(v[8 * c + 2]) * inv, __uint_as_float(v[8 * c + 3]) * inv);
                    w.z = pack_f16(__uint_as_float(v[8 * c + 4]) * inv, __uint_as_float(v[8 * c + 5]) * inv);
                    w.w = pack_f16(__uint_as_float(v[8 * c + 6]) * inv, __uint_as_float(v[8 * c + 7]) * inv);
                    *reinterpret_cast<uint4*>(o + c * 8) = w;
                }
            }
        }
        if (trace != nullptr && warp == 2 && lane == 0) {  // debug: cycles per phase of this CTA
            long long* tr = trace + (static_cast<long long>(blockIdx.y) * gridDim.x + blockIdx.x) * 6;
            tr[0] = t_s - t_start; tr[1] = t_p - t_s; tr[5] = t_m - t_s; tr[2] = t_o - t_p; tr[3] = clock64() - t_o; tr[4] = t_start;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem, 256);
    }
}

}  // namespace

static int g_tc_enabled = 1;
static long long* g_trace = nullptr;
void attention_tc_set_trace(long long* buf) { g_trace = buf; }
void attention_tc_enable(int on) { g_tc_enabled = on; }

// returns SB_ERR_UNSUPPORTED when the shape is outside this kernel's envelope (caller falls back
// to the general flash kernel in attention.cu -- same arithmetic, legacy tensor path)
int attention_tc_fwd(const AttnParams& p, int head_dim, cudaStream_t stream) {
    if (!g_tc_enabled || head_dim != 64 || p.coords != nullptr || p.mask != nullptr || p.out_f32 || p.S > 256 || p.S < 1)
        return SB_ERR_UNSUPPORTED;
    if (p.q == nullptr || p.k != p.q + static_cast<long long>(p.H) * 64 || p.v != p.q + 2LL * p.H * 64 ||
        p.v_row_stride != 0 || p.row_stride != 3LL * p.H * 64 || (p.out_row_stride % 8) != 0 ||
        (reinterpret_cast<uintptr_t>(p.q) & 15) != 0 || (reinterpret_cast<uintptr_t>(p.out) & 15) != 0)
        return SB_ERR_UNSUPPORTED;  // expects the packed [.., 3, H, 64] projection layout
    const int D = p.H * 64;
    const int nk = (p.S + 15) / 16 * 16;
    CUtensorMap tm_q, tm_kv;
    int rc = make_tmap_3d_f16(&tm_q, p.q, 3 * D, p.S, p.B, p.row_stride, p.batch_stride, 64, 128);
    if (rc != SB_OK) return rc;
    rc = make_tmap_3d_f16(&tm_kv, p.q, 3 * D, p.S, p.B, p.row_stride, p.batch_stride, 64, nk);
    if (rc != SB_OK) return rc;
    const TcSmem L = tc_layout(nk);
    static int configured_bytes = 0;
    if (L.total > configured_bytes) {
        if (cudaFuncSetAttribute(vit_attn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, L.total) != cudaSuccess)
            return SB_ERR_CUDA;
        configured_bytes = L.total;
    }
    if (static_cast<long long>(p.B) * p.H > 65535) return SB_ERR_UNSUPPORTED;
    const int nq = (p.q_rows > 0 && p.q_rows < p.S) ? p.q_rows : p.S;
    dim3 grid((nq + 127) / 128, p.B * p.H);
    ProfScope prof(PROF_ATTN, 4.0 * p.B * p.H * static_cast<double>(p.S) * p.S * 64, stream);
    vit_attn_tc_kernel<<<grid, TC_THREADS, L.total, stream>>>(
        tm_q, tm_kv, static_cast<__half*>(p.out), p.out_row_stride, p.out_batch_stride, p.S, p.H, D, nk,
        p.scale_log2, L, g_trace);
    count_launch();
    return cudaGetLastError() == cudaSuccess ? SB_OK : SB_ERR_CUDA;
}

}  // namespace sb
