// Persistent tcgen05 attention for the ViT tile encoder (T = 197 tokens, head_dim 64).
//
//   O = softmax(Q K^T * scale) V      per (tile, head); no mask, no ALiBi, <= 240 keys
//
// replaces timm Attention.forward's F.scaled_dot_product_attention inside the ViT blocks run by
// src/stamp/preprocessing/__init__.py:325 (restated in oracle/vit_oracle.py: block_forward).
//
// Why this shape: per-phase clock traces of the non-persistent kernel (attention_tc.cu) showed the
// softmax phase bound by TMEM READ bandwidth (~64 B/clk/SM): a row-max pass plus an exp pass read
// every score twice.  Here each score is read ONCE into registers, which needs one CTA per SM, so
// the CTA is persistent and hides the loads itself:
//   warp 0     TMA: Q/K/V of item i+1 stream into the other smem stage while item i is processed
//   warp 1     MMA: S(i+1) = Q K^T into the other TMEM S buffer during softmax(i); O(i) = P V
//   warps 2-17 softmax: FOUR threads per query row (= TMEM lane), each owning every 4th 16-key chunk
//              (<= 64 scores in registers): max -> smem exchange -> exp2 -> fp16 P in the swizzled UMMA
//              layout; the O epilogue of item i-1 is interleaved before the P store of item i.
// TMEM: S0 | S1 | O = 2 * nk + 64 <= 512 columns.  smem: 2 x (Q 16 KB + K + V) + P 64 KB.
#include <math.h>

#include "attention.cuh"
#include "common.cuh"
#include "gemm.cuh"

namespace sb {
namespace {

constexpr int VP_WARPS = 16;   // softmax warps: 4 per TMEM lane quarter
constexpr int VP_THREADS = 64 + VP_WARPS * 32;
constexpr int VP_SOFTMAX = VP_WARPS * 32;
constexpr int BLK = 128 * 128;  // 128 rows x 64 halfs (one swizzled tile)

struct VpSmem {
    int kv_bytes, stage_bytes, off_p, off_x, off_bar, total;
};

inline VpSmem vp_layout(int nk) {
    VpSmem s;
    s.kv_bytes = ((nk * 128 + 1023) / 1024) * 1024;
    s.stage_bytes = BLK + 2 * s.kv_bytes;         // Q | K | V
    s.off_p = 2 * s.stage_bytes;                  // 4 x 64-key P blocks
    s.off_x = s.off_p + 4 * BLK;                  // exchange: 2 parities x ([4][128] maxima + [4][128] sums)
    s.off_bar = s.off_x + 2 * 1024 * 4;
    s.total = s.off_bar + 128 + 1024;
    return s;
}

__global__ void __launch_bounds__(VP_THREADS, 1)
vit_attn_persist_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_kv,
                        __half* __restrict__ out, long long out_row_stride, long long out_batch_stride,
                        int S, int H, int D, int nk, int n_items, int nmt, float scale_log2, VpSmem L, long long* trace) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>(
        (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    uint8_t* sP = smem + L.off_p;
    float* sX = reinterpret_cast<float*>(smem + L.off_x);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L.off_bar);
    uint64_t* full = bars;        // [2] stage loaded              TMA -> MMA
    uint64_t* empty = bars + 2;   // [2] stage consumed            MMA -> TMA
    uint64_t* sfull = bars + 4;   // [2] S in TMEM                 MMA -> softmax
    uint64_t* sfree = bars + 6;   // [2] S copied to registers     softmax -> MMA
    uint64_t* pfull = bars + 8;   //     P in smem                 softmax -> MMA
    uint64_t* pfree = bars + 9;   //     P consumed, O in TMEM     MMA -> softmax
    uint64_t* ofree = bars + 10;  //     O copied out              softmax -> MMA
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 11);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int my_items = (n_items - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tm_q);
        tma_prefetch_desc(&tm_kv);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], 1);
            mbar_init(&sfull[i], 1);
            mbar_init(&sfree[i], VP_WARPS);
        }
        mbar_init(pfull, VP_WARPS);
        mbar_init(pfree, 1);
        mbar_init(ofree, VP_WARPS);
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t col_o = 2 * nk;

    if (warp == 0) {
        // ------------------------------------ TMA producer ------------------------------------
        if (lane == 0) {
            for (int i = 0; i < my_items; ++i) {
                const int w = blockIdx.x + i * gridDim.x;
                const int mt = w % nmt, bh = w / nmt, b = bh / H, h = bh % H;
                const int st = i & 1;
                uint8_t* sQ = smem + st * L.stage_bytes;
                mbar_wait(&empty[st], ((i >> 1) & 1) ^ 1);
                mbar_expect_tx(&full[st], BLK + 2 * nk * 128);
                tma_load_3d(sQ, &tm_q, &full[st], h * 64, mt * 128, b);
                tma_load_3d(sQ + BLK, &tm_kv, &full[st], D + h * 64, 0, b);
                tma_load_3d(sQ + BLK + L.kv_bytes, &tm_kv, &full[st], 2 * D + h * 64, 0, b);
            }
        }
    } else if (warp == 1) {
        // ------------------------------------ MMA issuer --------------------------------------
        if (lane == 0) {
            const uint32_t idesc_s = umma_idesc_f16(128, nk, false, false, false);
            const uint32_t idesc_o = umma_idesc_f16(128, 64, false, false, true);  // V: MN-major B
            const int steps = nk / 16;
            auto issue_s = [&](int i) {
                const int st = i & 1;
                const uint32_t ph = (i >> 1) & 1;
                mbar_wait(&full[st], ph);
                mbar_wait(&sfree[st], ph ^ 1);
                tc_fence_after();
                uint8_t* sQ = smem + st * L.stage_bytes;
                const uint64_t a = umma_desc_k128(smem_u32(sQ)), bd = umma_desc_k128(smem_u32(sQ + BLK));
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_f16_ss(tmem + st * nk, a + 2 * k, bd + 2 * k, idesc_s, k != 0);
                umma_commit(&sfull[st]);
            };
            if (my_items > 0) issue_s(0);
            for (int i = 0; i < my_items; ++i) {
                if (i + 1 < my_items) issue_s(i + 1);
                const int st = i & 1;
                mbar_wait_spin(pfull, i & 1);
                mbar_wait_spin(ofree, (i & 1) ^ 1);   // O of item i-1 has been copied out
                tc_fence_after();
                const uint8_t* sV = smem + st * L.stage_bytes + BLK + L.kv_bytes;
                for (int s = 0; s < steps; ++s) {
                    const uint64_t a = umma_desc_k128(smem_u32(sP + (s >> 2) * BLK)) + 2 * (s & 3);
                    const uint64_t bd = umma_desc_mn128(smem_u32(sV + s * 2048), 0);
                    umma_f16_ss(tmem + col_o, a, bd, idesc_o, s != 0);
                }
                umma_commit(pfree);
                umma_commit(&empty[st]);
            }
        }
    } else {
        // ------ softmax: FOUR threads per query row (= TMEM lane), 16-key chunks c = part + 4j ------
        const int quarter = warp & 3;
        const int part = (warp - 2) >> 2;               // 0..3
        const int r = quarter * 32 + lane;
        const uint32_t t_lane = tmem + (static_cast<uint32_t>(quarter * 32) << 16);
        const int nch = nk >> 4;                        // 16-key chunks in a row (<= 15)
        const uint32_t sP_addr = smem_u32(sP);
        float l_prev = 1.f;                             // row sum of the previous item (for its epilogue)
        int row_prev = 0;
        long long obase_prev = 0;

        auto epilogue_prev = [&]() {
            // O(i-1) = cols [col_o, col_o + 64): every part copies 16 columns
            uint32_t o0[16];
            tmem_ld_32x32b_x16(t_lane + col_o + part * 16, o0);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(ofree);
            if (row_prev < S) {
                const float inv = 1.0f / l_prev;
                __half* dst = out + obase_prev + part * 16;
                auto pack8 = [&](const uint32_t* o) {
                    return make_uint4(pack_f16(__uint_as_float(o[0]) * inv, __uint_as_float(o[1]) * inv),
                                      pack_f16(__uint_as_float(o[2]) * inv, __uint_as_float(o[3]) * inv),
                                      pack_f16(__uint_as_float(o[4]) * inv, __uint_as_float(o[5]) * inv),
                                      pack_f16(__uint_as_float(o[6]) * inv, __uint_as_float(o[7]) * inv));
                };
                *reinterpret_cast<uint4*>(dst) = pack8(o0);
                *reinterpret_cast<uint4*>(dst + 8) = pack8(o0 + 8);
            }
        };

        for (int i = 0; i < my_items; ++i) {
            const int w = blockIdx.x + i * gridDim.x;
            const int mt = w % nmt, bh = w / nmt, b = bh / H, h = bh % H;
            const int st = i & 1;
            const int row = mt * 128 + r;
            float* xm = sX + (i & 1) * 1024;           // exchange buffers, double-buffered by item parity:
            float* xl = xm + 512;                      // [4][128] maxima | [4][128] sums

            // ---- A. the whole row share of this thread: S -> registers, one TMEM pass ----
            mbar_wait_spin(&sfull[st], (i >> 1) & 1);
            tc_fence_after();
            uint32_t v[4][16];
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (part + 4 * j < nch) tmem_ld_32x32b_x16(t_lane + st * nk + (part + 4 * j) * 16, v[j]);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&sfree[st]);     // MMA may overwrite this S buffer (item i+2)

            // ---- B. row max (3 partial threads -> smem exchange), exp2 + fp16 pack in place ----
            float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int key0 = (part + 4 * j) * 16;
                if (part + 4 * j < nch) {
#pragma unroll
                    for (int e = 0; e < 16; ++e)
                        if (key0 + e < S) mx4[e & 3] = fmaxf(mx4[e & 3], __uint_as_float(v[j][e]));
                }
            }
            float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
            xm[part * 128 + r] = mx;
            asm volatile("bar.sync 1, %0;\n" ::"n"(VP_SOFTMAX) : "memory");
            mx = fmaxf(fmaxf(xm[r], xm[128 + r]), fmaxf(xm[256 + r], xm[384 + r]));
            const float ms = mx * scale_log2;
            float l4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int key0 = (part + 4 * j) * 16;
                if (part + 4 * j < nch) {
#pragma unroll
                    for (int e = 0; e < 16; e += 2) {
                        const float p0 = (key0 + e < S) ? ex2_approx(fmaf(__uint_as_float(v[j][e]), scale_log2, -ms)) : 0.f;
                        const float p1 = (key0 + e + 1 < S) ? ex2_approx(fmaf(__uint_as_float(v[j][e + 1]), scale_log2, -ms)) : 0.f;
                        l4[(e >> 1) & 3] += p0 + p1;
                        v[j][e >> 1] = pack_f16(p0, p1);     // packed P reuses the score registers
                    }
                }
            }
            float l = (l4[0] + l4[1]) + (l4[2] + l4[3]);
            xl[part * 128 + r] = l;
            asm volatile("bar.sync 2, %0;\n" ::"n"(VP_SOFTMAX) : "memory");
            l = (xl[r] + xl[128 + r]) + (xl[256 + r] + xl[384 + r]);

            // ---- C/D. P buffer free <=> P V of item i-1 retired <=> O(i-1) complete: copy it out ----
            mbar_wait_spin(pfree, (i & 1) ^ 1);
            tc_fence_after();
            if (i > 0) epilogue_prev();   // (completion #k of ofree == O(k) copied out; item 0 has no predecessor)

            // ---- E. P(i) -> smem (K-major, 128B swizzle), hand over to the MMA warp ----
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int key0 = (part + 4 * j) * 16;
                if (part + 4 * j < nch) {
                    const uint32_t pb = sP_addr + (key0 >> 6) * BLK + r * 128;
                    const int ch0 = (key0 & 63) >> 3;
                    sts_v4(pb + ((ch0 ^ (r & 7)) * 16), make_uint4(v[j][0], v[j][1], v[j][2], v[j][3]));
                    sts_v4(pb + (((ch0 + 1) ^ (r & 7)) * 16), make_uint4(v[j][4], v[j][5], v[j][6], v[j][7]));
                }
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(pfull);
            l_prev = l;
            row_prev = row;
            obase_prev = b * out_batch_stride + static_cast<long long>(row) * out_row_stride + h * 64;
        }
        if (my_items > 0) {
            mbar_wait_spin(pfree, (my_items & 1) ^ 1);     // P V of the last item retired
            tc_fence_after();
            epilogue_prev();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem, 512);
    }
}

int g_vp_enabled = 0;  // opt-in (stamp_b200_attention_tc_enable bit 1): equal to attention_tc.cu end to end on B200
long long* g_vp_trace = nullptr;

}  // namespace

void attention_vit_persist_enable(int on) { g_vp_enabled = on; }
void attention_vit_persist_set_trace(long long* buf) { g_vp_trace = buf; }

// SB_ERR_UNSUPPORTED when outside the envelope -> caller tries attention_tc_fwd, then the general kernel
int attention_vit_persist_fwd(const AttnParams& p, int head_dim, cudaStream_t stream) {
    if (!g_vp_enabled || head_dim != 64 || p.coords != nullptr || p.mask != nullptr || p.out_f32 || p.S > 240 || p.S < 1)
        return SB_ERR_UNSUPPORTED;
    if (p.q == nullptr || p.k != p.q + static_cast<long long>(p.H) * 64 || p.v != p.q + 2LL * p.H * 64 ||
        p.v_row_stride != 0 || p.row_stride != 3LL * p.H * 64 || (p.out_row_stride % 8) != 0 ||
        (reinterpret_cast<uintptr_t>(p.q) & 15) != 0 || (reinterpret_cast<uintptr_t>(p.out) & 15) != 0)
        return SB_ERR_UNSUPPORTED;  // expects the packed [.., 3, H, 64] projection layout
    const int D = p.H * 64;
    const int nk = (p.S + 15) / 16 * 16;
    const VpSmem L = vp_layout(nk);
    if (L.total > 232448 || 2 * nk + 64 > 512) return SB_ERR_UNSUPPORTED;
    CUtensorMap tm_q, tm_kv;
    int rc = make_tmap_3d_f16(&tm_q, p.q, 3 * D, p.S, p.B, p.row_stride, p.batch_stride, 64, 128);
    if (rc != SB_OK) return rc;
    rc = make_tmap_3d_f16(&tm_kv, p.q, 3 * D, p.S, p.B, p.row_stride, p.batch_stride, 64, nk);
    if (rc != SB_OK) return rc;
    static int configured_bytes = 0;
    if (L.total > configured_bytes) {
        if (cudaFuncSetAttribute(vit_attn_persist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, L.total) != cudaSuccess)
            return SB_ERR_CUDA;
        configured_bytes = L.total;
    }
    const int nmt = (p.S + 127) / 128;
    const long long n_items_ll = static_cast<long long>(p.B) * p.H * nmt;
    if (n_items_ll > 2000000000LL) return SB_ERR_UNSUPPORTED;
    const int n_items = static_cast<int>(n_items_ll);
    const int sms = gemm_num_sms();
    const int grid = n_items < sms ? n_items : sms;
    ProfScope prof(PROF_ATTN, 4.0 * p.B * p.H * static_cast<double>(p.S) * p.S * 64, stream);
    vit_attn_persist_kernel<<<grid, VP_THREADS, L.total, stream>>>(
        tm_q, tm_kv, static_cast<__half*>(p.out), p.out_row_stride, p.out_batch_stride, p.S, p.H, D, nk, n_items,
        nmt, p.scale_log2, L, g_vp_trace);
    count_launch();
    return cudaGetLastError() == cudaSuccess ? SB_OK : SB_ERR_CUDA;
}

}  // namespace sb
