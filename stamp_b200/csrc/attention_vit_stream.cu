// Persistent streaming tcgen05 attention for the ViT tile encoder (T = 197 tokens, head_dim 64).
//
//   O = softmax(Q K^T * scale) V      per (tile, head); no mask, no ALiBi
//
// replaces timm Attention.forward's F.scaled_dot_product_attention inside the ViT blocks run by
// src/stamp/preprocessing/__init__.py:325 (restated in oracle/vit_oracle.py: block_forward).
//
// The one-shot kernel of the first round (attention_tc.cu: one CTA per 128 query rows, all keys in one pass) spent
// most of its 13 400 cycles per CTA outside the softmax: 3 600 cycles from launch to the first scores (barrier
// init, TMEM allocation, 68 KB of TMA loads), 2 400 waiting for P V, and it read every score twice from tensor
// memory (row maximum, then exp), which is the scarce resource of the softmax side (64 B / clock / SM).  This
// kernel is the long-bag kernel (attention_mil_v3.cu) turned persistent:
//   * work units (tile, head, 128-row query block) stream through ONE resident CTA after another: the TMA warp
//     keeps loading the Q block of the next unit and the K / V tiles of 64 keys into a 4-stage ring while the
//     current unit is in flight -- no per-unit launch, allocation or pipeline fill;
//   * keys are processed 64 at a time with an online softmax: each score is read from tensor memory ONCE into the
//     registers of the thread that owns the query row, the accumulator is rescaled lazily (only when a row maximum
//     grows by more than 2^8; warp-uniform branch, tcgen05.ld / .st are collective);
//   * P goes back to tensor memory over the scores it came from and feeds P V as a TMEM operand (tcgen05.mma TS
//     form): no shared-memory round trip;
//   * the output accumulator is double buffered (S0 | S1 | Oa | Ob = 256 TMEM columns, two CTAs per SM), so the
//     first products of the next unit do not wait for the epilogue of this one; the last key tile of a unit (5 of
//     197 keys) is computed, read and exponentiated at 16 columns instead of 64.
// Warps: 0 TMA, 1 tcgen05.mma issue, 2-5 softmax (one thread per query row = TMEM lane).
#include <math.h>

#include "attention.cuh"
#include "common.cuh"
#include "gemm.cuh"

namespace sb {
namespace {

constexpr int VS_THREADS = 192;
constexpr int SUB_Q = 128 * 128;       // one 64-column sub-tile of a query block: 128 rows x 64 halfs
constexpr int SUB_K = 64 * 128;        // one 64-column sub-tile of a key / value tile: 64 keys x 64 halfs
// Fetching the next tile's scores during this tile's exponentials (two 64-register tiles in flight) was measured
// SLOWER on B200 (6.76 k vs 7.35 k tiles/s end to end: 168 registers with spills); kept behind this switch.
constexpr bool VS_PREFETCH = false;

// Head dimension 64 (ViT-L/16, UNI2-h, H-optimus): everything double buffered.  Head dimension 80 (ViT-H/14: Virchow2)
// is handled as TWO 64-column sub-tiles per operand, the second one loaded as a full 64-column TMA box of which only
// the first 16 columns belong to the head (the rest is never multiplied for Q / K; for V it lands in accumulator
// columns 80..127 that are never read): the contraction is 5 steps of 16 instead of 4, the P V product one N = 128
// MMA over both sub-tiles (MN-major operand of two 64-column blocks).  Shared memory doubles, so Q and the output
// accumulator are single-buffered there and the K / V ring has two stages -- still two CTAs per SM.
template <int HD>
struct VsCfg {
    static constexpr int NSUB = (HD + 63) / 64;
    static constexpr int KSTEPS = HD / 16;              // of the Q K^T contraction
    static constexpr int OCOLS = NSUB * 64;             // TMEM columns of one output accumulator
    static constexpr int QBUFS = (NSUB == 1) ? 2 : 1;
    static constexpr int OBUFS = (NSUB == 1) ? 2 : 1;
    static constexpr int STAGES = (NSUB == 1) ? 4 : 2;
    static constexpr int QB_BYTES = NSUB * SUB_Q;
    static constexpr int KT_BYTES = NSUB * SUB_K;       // K tile (and V tile)
    static constexpr int off_q = 0;
    static constexpr int off_kv = QBUFS * QB_BYTES;     // STAGES x (K tile | V tile)
    static constexpr int off_bar = off_kv + STAGES * 2 * KT_BYTES;
    static constexpr int total = off_bar + 256 + 1024;
    static_assert(HD % 16 == 0 && HD <= 128, "head dimension: multiple of 16, at most 128");
    static_assert(128 + OBUFS * OCOLS <= 256, "TMEM: S0 | S1 | accumulators within 256 columns");
};

template <int HD>
__global__ void __launch_bounds__(VS_THREADS, 2)
vit_attn_stream_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_kv,
                       __half* __restrict__ out, long long out_row_stride, long long out_batch_stride,
                       int S, int H, int D, int n_units, int nqb, float scale_log2, float rescale_margin) {
    using C = VsCfg<HD>;
    constexpr int VS_STAGES = C::STAGES, QB_BYTES = C::QB_BYTES, KT_BYTES = C::KT_BYTES;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>(
        (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    uint8_t* sQ = smem + C::off_q;
    uint8_t* sKV = smem + C::off_kv;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::off_bar);
    uint64_t* full = bars;                       // [VS_STAGES] TMA -> MMA   (K, V tile)
    uint64_t* empty = bars + VS_STAGES;          // [VS_STAGES] MMA -> TMA
    uint64_t* qfull = bars + 2 * VS_STAGES;      // [2] TMA -> MMA   (Q block of a unit)
    uint64_t* qempty = qfull + 2;                // [2] MMA -> TMA   (all S products of the unit retired)
    uint64_t* sfull = qempty + 2;                // [2] MMA -> softmax (S tile in TMEM)
    uint64_t* ofull = sfull + 2;                 // [2] MMA -> softmax (all P V of a unit retired)
    uint64_t* oempty = ofull + 2;                // [2] softmax -> MMA (accumulator copied out)
    uint64_t* pfull = oempty + 2;                // [2] softmax -> MMA (P tile in TMEM); one barrier per tile parity: a
                                                 //     warp that runs ahead (S(g+1) is ready early) must not complete
                                                 //     tile g's phase with its arrival for tile g+1
    uint64_t* pvdone = pfull + 2;                //     MMA -> softmax (P V of the tile retired)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pvdone + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nkt = (S + 63) / 64;
    const int my_units = (n_units - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
    // keys of the last tile, rounded up to the 16-key MMA step (197 tokens: 3 full tiles + 16 columns)
    const int last_cols = ((S - (nkt - 1) * 64) + 15) / 16 * 16;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tm_q);
        tma_prefetch_desc(&tm_kv);
        for (int i = 0; i < VS_STAGES; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&qfull[i], 1);
            mbar_init(&qempty[i], 1);
            mbar_init(&sfull[i], 1);
            mbar_init(&ofull[i], 1);
            mbar_init(&oempty[i], 4);
            mbar_init(&pfull[i], 4);
        }
        mbar_init(pvdone, 1);
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
    constexpr uint32_t COL_O = 128;    // S0 at column 0, S1 at 64, accumulator(s) from 128
    auto qbuf = [&](int i) { return (C::QBUFS == 2) ? (i & 1) : 0; };
    auto obuf = [&](int i) { return (C::OBUFS == 2) ? (i & 1) : 0; };
    // completion count a waiter of unit i's buffer has to see: with one buffer every unit uses it
    auto qpar = [&](int i) { return (C::QBUFS == 2) ? ((i >> 1) & 1) : (i & 1); };
    auto opar = [&](int i) { return (C::OBUFS == 2) ? ((i >> 1) & 1) : (i & 1); };

    // unit index -> (image b, head h, query block): the query blocks of one (b, h) are adjacent units, so the second
    // read of its K / V tiles hits L2
    auto unit_of = [&](int i, int& b, int& h, int& q0) {
        const int u = static_cast<int>(blockIdx.x) + i * static_cast<int>(gridDim.x);
        const int qb = u % nqb, bh = u / nqb;
        b = bh / H; h = bh % H; q0 = qb * 128;
    };

    if (warp == 0) {
        // ------------------------------------ TMA producer ------------------------------------
        if (lane == 0) {
            int g = 0;   // running key-tile index of this CTA
            for (int i = 0; i < my_units; ++i) {
                int b, h, q0;
                unit_of(i, b, h, q0);
                mbar_wait(&qempty[qbuf(i)], qpar(i) ^ 1);
                mbar_expect_tx(&qfull[qbuf(i)], QB_BYTES);
#pragma unroll
                for (int sub = 0; sub < C::NSUB; ++sub)
                    tma_load_3d(sQ + qbuf(i) * QB_BYTES + sub * SUB_Q, &tm_q, &qfull[qbuf(i)], h * HD + sub * 64, q0, b);
                for (int kt = 0; kt < nkt; ++kt, ++g) {
                    const int st = g % VS_STAGES;
                    uint8_t* dst = sKV + st * 2 * KT_BYTES;
                    mbar_wait(&empty[st], ((g / VS_STAGES) & 1) ^ 1);
                    mbar_expect_tx(&full[st], 2 * KT_BYTES);
#pragma unroll
                    for (int sub = 0; sub < C::NSUB; ++sub) {
                        tma_load_3d(dst + sub * SUB_K, &tm_kv, &full[st], D + h * HD + sub * 64, kt * 64, b);                 // K
                        tma_load_3d(dst + KT_BYTES + sub * SUB_K, &tm_kv, &full[st], 2 * D + h * HD + sub * 64, kt * 64, b);  // V
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------ MMA issuer --------------------------------------
        if (lane == 0) {
            const uint32_t idesc_o = umma_idesc_f16(128, C::OCOLS, false, false, true);   // V: MN-major B operand
            const int total = my_units * nkt;
            auto issue_s = [&](int g) {
                const int i = g / nkt, kt = g - i * nkt;
                const int st = g % VS_STAGES;
                if (kt == 0) mbar_wait(&qfull[qbuf(i)], qpar(i));
                mbar_wait(&full[st], (g / VS_STAGES) & 1);
                tc_fence_after();
                const int ncols = (kt == nkt - 1) ? last_cols : 64;
                const uint32_t idesc_s = umma_idesc_f16(128, ncols, false, false, false);
#pragma unroll
                for (int k = 0; k < C::KSTEPS; ++k) {
                    // 16 head dimensions per step; steps 4.. come from the second 64-column sub-tile
                    const uint64_t q_desc = umma_desc_k128(smem_u32(sQ + qbuf(i) * QB_BYTES + (k >> 2) * SUB_Q)) + 2 * (k & 3);
                    const uint64_t k_desc = umma_desc_k128(smem_u32(sKV + st * 2 * KT_BYTES + (k >> 2) * SUB_K)) + 2 * (k & 3);
                    umma_f16_ss(tmem + (g & 1) * 64, q_desc, k_desc, idesc_s, k != 0);
                }
                umma_commit(&sfull[g & 1]);
                if (kt == nkt - 1) umma_commit(&qempty[qbuf(i)]);    // the unit's Q block is free once its S tiles exist
            };
            if (total > 0) issue_s(0);
            for (int g = 0; g < total; ++g) {
                if (g + 1 < total) issue_s(g + 1);
                const int i = g / nkt, kt = g - i * nkt;
                const int st = g % VS_STAGES;
                mbar_wait(&pfull[g & 1], (g >> 1) & 1);
                if (kt == 0) mbar_wait(&oempty[obuf(i)], opar(i) ^ 1);   // the accumulator's previous unit has been copied out
                tc_fence_after();
                const int ksteps = (kt == nkt - 1) ? last_cols / 16 : 4;
                for (int k = 0; k < ksteps; ++k) {
                    // 16 keys per step: 16 rows x 128 B of the V tile as an MN-major operand; P: 8 TMEM columns
                    // (N spans the 64-column sub-tiles of the V tile: leading byte offset = one sub-tile)
                    const uint64_t v_desc = umma_desc_mn128(smem_u32(sKV + st * 2 * KT_BYTES + KT_BYTES + k * 2048), SUB_K);
                    umma_f16_ts(tmem + COL_O + obuf(i) * C::OCOLS, tmem + (g & 1) * 64 + k * 8, v_desc, idesc_o, (kt | k) != 0);
                }
                umma_commit(&empty[st]);
                umma_commit(pvdone);
                if (kt == nkt - 1) umma_commit(&ofull[obuf(i)]);
            }
        }
    } else {
        // ---------------- softmax: one thread per query row (= TMEM lane), 64 keys per tile ----------------
        const int quarter = warp & 3;
        const int r = quarter * 32 + lane;
        const uint32_t t_lane = tmem + (static_cast<uint32_t>(quarter * 32) << 16);
        int g = 0;
        for (int i = 0; i < my_units; ++i) {
            int b, h, q0;
            unit_of(i, b, h, q0);
            const int row = q0 + r;
            // warps whose 32 rows all lie past the last token still walk the barrier protocol, without the arithmetic
            const bool warp_live = (q0 + quarter * 32) < S;
            float ms = -INFINITY;      // reference maximum of the row, already multiplied by scale * log2(e)
            float l4[4] = {0.f, 0.f, 0.f, 0.f};
            // Software pipeline over the full key tiles of the unit: the scores of tile kt+1 are fetched from tensor
            // memory (tcgen05.ld is asynchronous until tcgen05.wait::ld) WHILE tile kt is exponentiated -- a warp's
            // 64-column load keeps its sub-partition's TMEM port busy for ~512 cycles, its 64 ex2 per thread the MUFU
            // for another ~512.  `have` tells whether the tile's scores are already in `cur`.
            auto full_tile = [&](int kt, uint32_t (&cur)[64], uint32_t (&nxt)[64], bool have) {
                const int s = g & 1;
                if (!have) {
                    mbar_wait(&sfull[s], (g >> 1) & 1);
                    tc_fence_after();
                }
                if (warp_live) {
                    if (!have) {
                        tmem_ld_32x32b_x64(t_lane + s * 64, cur);
                        tmem_ld_wait();
                    }
                    float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
                    for (int j = 0; j < 64; ++j) mx4[j & 3] = fmaxf(mx4[j & 3], __uint_as_float(cur[j]));
                    const float cand = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3])) * scale_log2;
                    const bool need = cand > ms + rescale_margin;
                    const bool resc = need && ms != -INFINITY;
                    if (__any_sync(0xffffffffu, resc)) {
                        const float f = resc ? ex2_approx(ms - cand) : 1.0f;
                        l4[0] *= f; l4[1] *= f; l4[2] *= f; l4[3] *= f;
                        mbar_wait(pvdone, (g & 1) ^ 1);       // P V of the previous tile has retired
                        tc_fence_after();
#pragma unroll 1
                        for (int c = 0; c < (HD + 31) / 32; ++c) {
                            uint32_t o[32];
                            tmem_ld_32x32b_x32(t_lane + COL_O + obuf(i) * C::OCOLS + c * 32, o);
                            tmem_ld_wait();
#pragma unroll
                            for (int j = 0; j < 32; ++j) o[j] = __float_as_uint(__uint_as_float(o[j]) * f);
                            tmem_st_32x32b_x32(t_lane + COL_O + obuf(i) * C::OCOLS + c * 32, o);
                        }
                        tmem_st_wait();
                    }
                    if (need) ms = cand;
                }
                const bool prefetch = VS_PREFETCH && kt + 1 < nkt - 1;   // the next tile is another full tile of this unit
                if (prefetch) {
                    mbar_wait(&sfull[s ^ 1], ((g + 1) >> 1) & 1);
                    tc_fence_after();
                    // (two 32-column loads, the second issued half-way through the exponentials: 64 destination
                    //  registers reserved from the start would push the kernel into spills)
                    if (warp_live) tmem_ld_32x32b_x32_at<0>(t_lane + (s ^ 1) * 64, nxt);
                }
                if (warp_live) {
                    // exp2 and fp16 packing in place: pair j lands in cur[j / 2], which has been read by then
#pragma unroll
                    for (int j = 0; j < 64; j += 2) {
                        if (j == 32 && prefetch)
                            tmem_ld_32x32b_x32_at<32>(t_lane + (s ^ 1) * 64 + 32, nxt);
                        const float p0 = ex2_approx(fmaf(__uint_as_float(cur[j]), scale_log2, -ms));
                        const float p1 = ex2_approx(fmaf(__uint_as_float(cur[j + 1]), scale_log2, -ms));
                        l4[(j >> 1) & 3] += p0 + p1;
                        cur[j >> 1] = pack_f16(p0, p1);
                    }
                    tmem_ld_wait();                           // the next tile's scores have landed (if any were requested)
                    tmem_st_32x32b_x32_lo(t_lane + s * 64, cur);
                    tmem_st_wait();
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&pfull[g & 1]);
                ++g;
                return prefetch;
            };
            uint32_t va[64], vb[64];
            {
                bool have = false;
                for (int kt = 0; kt + 1 < nkt; kt += 2) {
                    have = full_tile(kt, va, vb, have);
                    if (kt + 2 < nkt) have = full_tile(kt + 1, vb, va, have);
                }
            }
            {
                const int kt = nkt - 1;
                const int s = g & 1;
                mbar_wait(&sfull[s], (g >> 1) & 1);
                tc_fence_after();
                {
                    // last tile of the unit: `last_cols` (16 ... 64) columns, the tail keys masked
                    const int nvalid = S - kt * 64;
                    if (warp_live) {
                        float mx = -INFINITY;
                        for (int c = 0; c < last_cols; c += 16) {
                            uint32_t v[16];
                            tmem_ld_32x32b_x16(t_lane + s * 64 + c, v);
                            tmem_ld_wait();
#pragma unroll
                            for (int j = 0; j < 16; ++j)
                                if (c + j < nvalid) mx = fmaxf(mx, __uint_as_float(v[j]));
                        }
                        const float cand = mx * scale_log2;
                        const bool need = cand > ms + rescale_margin;
                        const bool resc = need && ms != -INFINITY;
                        if (__any_sync(0xffffffffu, resc)) {
                            const float f = resc ? ex2_approx(ms - cand) : 1.0f;
                            l4[0] *= f; l4[1] *= f; l4[2] *= f; l4[3] *= f;
                            mbar_wait(pvdone, (g & 1) ^ 1);
                            tc_fence_after();
#pragma unroll 1
                            for (int c = 0; c < (HD + 31) / 32; ++c) {
                                uint32_t o[32];
                                tmem_ld_32x32b_x32(t_lane + COL_O + obuf(i) * C::OCOLS + c * 32, o);
                                tmem_ld_wait();
#pragma unroll
                                for (int j = 0; j < 32; ++j) o[j] = __float_as_uint(__uint_as_float(o[j]) * f);
                                tmem_st_32x32b_x32(t_lane + COL_O + obuf(i) * C::OCOLS + c * 32, o);
                            }
                            tmem_st_wait();
                        }
                        if (need) ms = cand;
                        // second (short) read of the <= 64 columns: the scores of the last tile are not kept in registers
                        for (int c = 0; c < last_cols; c += 16) {
                            uint32_t v[16];
                            tmem_ld_32x32b_x16(t_lane + s * 64 + c, v);
                            tmem_ld_wait();
                            uint32_t q8[8];
#pragma unroll
                            for (int j = 0; j < 16; j += 2) {
                                const float p0 = (c + j < nvalid) ? ex2_approx(fmaf(__uint_as_float(v[j]), scale_log2, -ms)) : 0.f;
                                const float p1 = (c + j + 1 < nvalid) ? ex2_approx(fmaf(__uint_as_float(v[j + 1]), scale_log2, -ms)) : 0.f;
                                l4[(j >> 1) & 3] += p0 + p1;
                                q8[j >> 1] = pack_f16(p0, p1);
                            }
                            // P columns [c/2, c/2 + 8) of the tile
                            asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};\n" ::"r"(
                                             t_lane + s * 64 + (c >> 1)),
                                         "r"(q8[0]), "r"(q8[1]), "r"(q8[2]), "r"(q8[3]), "r"(q8[4]), "r"(q8[5]), "r"(q8[6]), "r"(q8[7])
                                         : "memory");
                        }
                        tmem_st_wait();
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&pfull[g & 1]);
                ++g;
            }
            // ---- epilogue of the unit: O / l -> fp16, this thread's 64 output columns (128 contiguous bytes) ----
            mbar_wait(&ofull[obuf(i)], opar(i));
            tc_fence_after();
            if (warp_live) {
                const float inv = 1.0f / ((l4[0] + l4[1]) + (l4[2] + l4[3]));
                __half* o = out + b * out_batch_stride + static_cast<long long>(row) * out_row_stride + h * HD;
#pragma unroll 1
                for (int c = 0; c < (HD + 31) / 32; ++c) {
                    uint32_t v[32];
                    tmem_ld_32x32b_x32(t_lane + COL_O + obuf(i) * C::OCOLS + c * 32, v);
                    tmem_ld_wait();
                    if (row < S) {
#pragma unroll
                        for (int j = 0; j < 32; j += 8) {
                            if (c * 32 + j >= HD) break;       // (head dimension 80: the third chunk holds 16 columns)
                            uint4 w;
                            w.x = pack_f16(__uint_as_float(v[j]) * inv, __uint_as_float(v[j + 1]) * inv);
                            w.y = pack_f16(__uint_as_float(v[j + 2]) * inv, __uint_as_float(v[j + 3]) * inv);
                            w.z = pack_f16(__uint_as_float(v[j + 4]) * inv, __uint_as_float(v[j + 5]) * inv);
                            w.w = pack_f16(__uint_as_float(v[j + 6]) * inv, __uint_as_float(v[j + 7]) * inv);
                            *reinterpret_cast<uint4*>(o + c * 32 + j) = w;
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&oempty[obuf(i)]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem, 256);
    }
}

int g_vs_enabled = 1;
int g_vs_eager = 0;

}  // namespace

void attention_vit_stream_enable(int on) {
    g_vs_enabled = on & 1;
    g_vs_eager = (on >> 1) & 1;
}

// returns SB_ERR_UNSUPPORTED when the shape is outside this kernel's envelope (the caller falls back to the
// one-shot kernel in attention_tc.cu, then to the general kernel in attention.cu -- same arithmetic)
namespace {

template <int HD>
int launch_vs(const AttnParams& p, cudaStream_t stream) {
    using C = VsCfg<HD>;
    const int D = p.H * HD;
    CUtensorMap tm_q, tm_kv;
    int rc = make_tmap_3d_f16(&tm_q, p.q, 3 * D, p.S, p.B, p.row_stride, p.batch_stride, 64, 128);
    if (rc != SB_OK) return rc;
    rc = make_tmap_3d_f16(&tm_kv, p.q, 3 * D, p.S, p.B, p.row_stride, p.batch_stride, 64, 64);
    if (rc != SB_OK) return rc;
    static bool configured = false;
    if (!configured) {
        if (cudaFuncSetAttribute(vit_attn_stream_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::total) != cudaSuccess)
            return SB_ERR_CUDA;
        configured = true;
    }
    const int nq = (p.q_rows > 0 && p.q_rows < p.S) ? p.q_rows : p.S;
    const int nqb = (nq + 127) / 128;
    const long long units_ll = static_cast<long long>(p.B) * p.H * nqb;
    if (units_ll > 2000000000LL) return SB_ERR_UNSUPPORTED;
    const int n_units = static_cast<int>(units_ll);
    const int slots = 2 * gemm_num_sms();
    const int grid = n_units < slots ? n_units : slots;
    ProfScope prof(PROF_ATTN, 4.0 * p.B * p.H * static_cast<double>(p.S) * p.S * HD, stream);
    vit_attn_stream_kernel<HD><<<grid, VS_THREADS, C::total, stream>>>(
        tm_q, tm_kv, static_cast<__half*>(p.out), p.out_row_stride, p.out_batch_stride, p.S, p.H, D, n_units, nqb,
        p.scale_log2, g_vs_eager ? 0.f : 8.f);
    count_launch();
    return cudaGetLastError() == cudaSuccess ? SB_OK : SB_ERR_CUDA;
}

}  // namespace

// returns SB_ERR_UNSUPPORTED when the shape is outside this kernel's envelope (the caller falls back to the
// one-shot kernel in attention_tc.cu, then to the general kernel in attention.cu -- same arithmetic)
int attention_vit_stream_fwd(const AttnParams& p, int head_dim, cudaStream_t stream) {
    if (!g_vs_enabled || (head_dim != 64 && head_dim != 80) || p.coords != nullptr || p.mask != nullptr || p.out_f32 ||
        p.S > 1024 || p.S < 1)
        return SB_ERR_UNSUPPORTED;
    const long long hw = static_cast<long long>(p.H) * head_dim;
    if (p.q == nullptr || p.k != p.q + hw || p.v != p.q + 2 * hw || p.v_row_stride != 0 || p.row_stride != 3 * hw ||
        (p.out_row_stride % 8) != 0 || (reinterpret_cast<uintptr_t>(p.q) & 15) != 0 ||
        (reinterpret_cast<uintptr_t>(p.out) & 15) != 0)
        return SB_ERR_UNSUPPORTED;  // expects the packed [.., 3, H, head_dim] projection layout
    return head_dim == 64 ? launch_vs<64>(p, stream) : launch_vs<80>(p, stream);
}

}  // namespace sb
