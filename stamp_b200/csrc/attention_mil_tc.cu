// tcgen05 attention for long feature bags, second generation: distances recomputed from the [S, 2] coordinates
// inside the kernel.  The MIL aggregator itself runs the third generation (attention_mil_v3.cu, distance tiles
// pre-computed and fed by TMA); this kernel remains (a) behind the stand-alone C-ABI primitive stamp_attention_fwd,
// whose callers pass coordinates and no distance matrix, and (b) as an independent implementation the tests compare
// the third generation against (stamp_b200_attention_tc_enable bit 5).
//
//   plain : O = softmax(Q K^T * scale) V
//   ALiBi : O = softmax(Q K^T * scale) V  -  c_h * Dist V          (bias subtracted AFTER the softmax,
//           src/stamp/modeling/models/vision_tranformer.py:58-72)
//
// One CTA = 128 query rows of one (bag, head), single pass over 64-key tiles (details above mil_attn_tc2_kernel).
#include <math.h>

#include "attention.cuh"
#include "attention_train.cuh"
#include "common.cuh"
#include "gemm.cuh"

namespace sb {
namespace {

constexpr int MT_THREADS = 320;  // TMA warp, MMA warp, 8 softmax warps (2 threads per query row)
constexpr int TILE_BYTES = 128 * 128;  // 128 rows x 64 halfs

__device__ __forceinline__ float2 lds_f2(uint32_t addr) {
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];\n" : "=f"(v.x), "=f"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts_f2(uint32_t addr, float2 v) {
    asm volatile("st.shared.v2.f32 [%0], {%1, %2};\n" ::"r"(addr), "f"(v.x), "f"(v.y) : "memory");
}
__device__ __forceinline__ void sts_u4(uint32_t addr, uint4 v) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};\n" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ float sqrt_approx(float x) {
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;\n" : "=f"(r) : "f"(x));
    return r;
}


// Training variant (TRAIN): bf16 operands; Dhat = dist * inv_rm_h, O = O1 / l - beta_h * O2; besides O (bf16)
// it stores what the backward needs -- Osm = O1 / l and Dhat V = O2 in fp32 and the row log-sum-exp (log2
// domain) -- see attention_train.cu.
struct MilTrainOut {
    uint16_t* out16;
    float* osm;
    float* odv;
    float* lse2;
    const float* beta;
    const float* inv_rm;
};

template <bool TRAIN>
__device__ __forceinline__ uint32_t pack_op(float a, float b) {
    if constexpr (TRAIN) return pack_bf16(a, b);
    else return pack_f16(a, b);
}

// ------------------------------------------------------------------------------------------------------
// Two-CTAs-per-SM variant of the single-pass kernel (default).  One CTA alone leaves every pipe below 55 %:
// each tile is a dependent chain  MMA -> TMEM load -> exp / sqrt -> smem stores -> MMA  and only eight softmax
// warps are resident.  With 64-key tiles the CTA needs S | O1a | O1b | O2 = 4 x 64 = 256 TMEM columns, 84 KB of
// shared memory and < 102 registers per thread, so two CTAs share an SM and cover each other's hand-offs.
// Same algorithm as above: per-thread reference maxima, lazy warp-uniform rescale, halves merged in the epilogue;
// a row's two threads own keys [0,32) and [32,64) of every tile.
constexpr int KV_BYTES = 64 * 128;     // 64 keys x 64 halfs

struct Mt2Smem {
    static constexpr int off_q = 0;
    static constexpr int off_k = TILE_BYTES;                 // 2 stages x 64 keys
    static constexpr int off_v = off_k + 2 * KV_BYTES;       // 2 stages
    static constexpr int off_p = off_v + 2 * KV_BYTES;       // [128 x 64] P tile
    static constexpr int off_d = off_p + TILE_BYTES;         // [128 x 64] D tile
    static constexpr int off_c = off_d + TILE_BYTES;         // key coordinates, 2 x 64 float2
    static constexpr int off_x = off_c + 2 * 64 * 8;         // row-statistics exchange, 4 x 128 floats
    static constexpr int off_bar = off_x + 4 * 128 * 4;
    static constexpr int total = off_bar + 128 + 1024;
};

template <bool ALIBI, bool TRAIN>
__global__ void __launch_bounds__(MT_THREADS, 2)
mil_attn_tc2_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                    const __grid_constant__ CUtensorMap tm_v,
                    const AttnParams p, int k_col0, const MilTrainOut t, float rescale_margin) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>(
        (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    uint8_t* sQ = smem + Mt2Smem::off_q;
    uint8_t* sK = smem + Mt2Smem::off_k;
    uint8_t* sV = smem + Mt2Smem::off_v;
    uint8_t* sP = smem + Mt2Smem::off_p;
    uint8_t* sD = smem + Mt2Smem::off_d;
    float2* sC = reinterpret_cast<float2*>(smem + Mt2Smem::off_c);
    float* sX = reinterpret_cast<float*>(smem + Mt2Smem::off_x);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Mt2Smem::off_bar);
    uint64_t* kfull = bars;        // [2] TMA -> MMA
    uint64_t* kempty = bars + 2;   // [2] MMA -> TMA
    uint64_t* sfull = bars + 4;    // [2] MMA -> softmax   (S tile in TMEM)
    uint64_t* sempty = bars + 6;   // [2] softmax -> MMA
    uint64_t* pfull = bars + 8;    // softmax -> MMA       (P/D tiles in smem)
    uint64_t* pempty = bars + 9;   // MMA -> softmax
    uint64_t* ofull = bars + 10;
    uint64_t* qfull = bars + 11;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.x / p.H, h = blockIdx.x % p.H;
    const int q0 = blockIdx.y * 128;
    const int S = p.S;
    const int nkt = (S + 63) / 64;              // key / value tiles of 64 rows

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tm_q);
        tma_prefetch_desc(&tm_k);
        tma_prefetch_desc(&tm_v);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&kfull[i], 1);
            mbar_init(&kempty[i], 1);
            mbar_init(&sfull[i], 1);      // only stage 0 is used: S is single-buffered in TMEM
            mbar_init(&sempty[i], 8);
        }
        mbar_init(pfull, 8);
        mbar_init(pempty, 1);
        mbar_init(ofull, 1);
        mbar_init(qfull, 1);
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
    constexpr uint32_t COL_O1A = 64, COL_O1B = 128, COL_O2 = 192;   // S at column 0

    if (warp == 0) {
        if (lane == 0) {
            mbar_expect_tx(qfull, TILE_BYTES);
            tma_load_3d(sQ, &tm_q, qfull, h * 64, q0, b);
            for (int kt = 0; kt < nkt; ++kt) {
                const int s = kt & 1;
                mbar_wait(&kempty[s], ((kt >> 1) & 1) ^ 1);
                mbar_expect_tx(&kfull[s], 2 * KV_BYTES);
                tma_load_3d(sK + s * KV_BYTES, &tm_k, &kfull[s], k_col0 + h * 64, kt * 64, b);
                tma_load_3d(sV + s * KV_BYTES, &tm_v, &kfull[s], h * 64, kt * 64, b);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc_s = umma_idesc_f16(128, 64, TRAIN, false, false);
            const uint32_t idesc_o = umma_idesc_f16(128, 64, TRAIN, false, true);  // V: MN-major B
            const uint64_t q_desc = umma_desc_k128(smem_u32(sQ));
            mbar_wait(qfull, 0);
            auto issue_s = [&](int kt) {
                const int s = kt & 1;                       // K / V smem tiles: two stages
                mbar_wait(&kfull[s], (kt >> 1) & 1);
                mbar_wait(&sempty[0], (kt & 1) ^ 1);        // S in TMEM: one stage, read into registers early
                tc_fence_after();
                const uint64_t k_desc = umma_desc_k128(smem_u32(sK + s * KV_BYTES));
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_f16_ss(tmem, q_desc + 2 * k, k_desc + 2 * k, idesc_s, k != 0);
                umma_commit(&sfull[0]);
            };
            issue_s(0);
            for (int kt = 0; kt < nkt; ++kt) {
                if (kt + 1 < nkt) issue_s(kt + 1);
                const int s = kt & 1;
                mbar_wait(pfull, kt & 1);
                tc_fence_after();
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint64_t v_desc = umma_desc_mn128(smem_u32(sV + s * KV_BYTES + k * 2048), 0);
                    const uint64_t p_desc = umma_desc_k128(smem_u32(sP)) + 2 * k;
                    // keys 0-31 of the tile accumulate into O1a, keys 32-63 into O1b (separate reference maxima)
                    umma_f16_ss(tmem + (k < 2 ? COL_O1A : COL_O1B), p_desc, v_desc, idesc_o, (kt | (k & 1)) != 0);
                    if constexpr (ALIBI) {
                        const uint64_t d_desc = umma_desc_k128(smem_u32(sD)) + 2 * k;
                        umma_f16_ss(tmem + COL_O2, d_desc, v_desc, idesc_o, (kt | k) != 0);
                    }
                }
                umma_commit(pempty);
                umma_commit(&kempty[s]);
            }
            umma_commit(ofull);
        }
    } else {
        const int quarter = warp & 3;
        const int half = (warp - 2) >> 2;          // keys [32*half, 32*half + 32) of every 64-key tile
        const int r = quarter * 32 + lane;
        const int st = threadIdx.x - 64;
        const int row = q0 + r;
        const uint32_t t_lane = tmem + (static_cast<uint32_t>(quarter * 32) << 16);
        const uint32_t t_o1 = t_lane + (half == 0 ? COL_O1A : COL_O1B);
        const float sl2 = p.scale_log2;
        const uint32_t sC_addr = smem_u32(sC), sP_addr = smem_u32(sP), sD_addr = smem_u32(sD);

        float2 cq = make_float2(0.f, 0.f);
        float slope = 0.f, descale = 1.f;
        const float2* cb = nullptr;
        if constexpr (ALIBI) {
            cb = reinterpret_cast<const float2*>(p.coords) + static_cast<long long>(b) * S;
            if (row < S) cq = __ldg(cb + row);
            if constexpr (TRAIN) {
                slope = __ldg(t.inv_rm + h);
                descale = __ldg(t.beta + h);
            } else {
                slope = __ldg(p.slope + h) * __ldg(p.dscale + 2 * b);
                descale = __ldg(p.dscale + 2 * b + 1);
            }
        }
        float ms = -INFINITY;      // reference maximum of this thread's keys, already multiplied by scale*log2(e)
        float l = 0.f;
        const uint32_t pb = sP_addr + r * 128;
        const uint32_t db = sD_addr + r * 128;
        for (int kt = 0; kt < nkt; ++kt) {
            const int s = kt & 1;
            const int kv_valid = min(64, S - kt * 64) - half * 32;   // valid keys among this thread's 32
            if constexpr (ALIBI) {
                if (st < 64) {
                    const int key = kt * 64 + st;
                    sts_f2(sC_addr + ((kt & 1) * 64 + st) * 8, (key < S) ? __ldg(cb + key) : make_float2(0.f, 0.f));
                }
                asm volatile("bar.sync 1, 256;\n" ::: "memory");
            }
            mbar_wait(&sfull[0], kt & 1);
            tc_fence_after();
            uint32_t v0[32];
            tmem_ld_32x32b_x32(t_lane + half * 32, v0);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&sempty[0]);     // S is in registers: the MMA warp may issue the next tile's S
            float mx = -INFINITY;
#pragma unroll
            for (int j = 0; j < 32; ++j)
                if (j < kv_valid) mx = fmaxf(mx, __uint_as_float(v0[j]));
            mbar_wait(pempty, (kt & 1) ^ 1);            // previous tile's P V / D V retired: O1 and sP / sD are ours
            tc_fence_after();
            {
                // raise the reference maximum (lazily: only past the margin) and rescale l and this thread's O1 row.
                // tcgen05.ld / .st are warp-collective (.sync.aligned): the branch is taken by the whole warp as
                // soon as one row needs it, rows that do not rescale by 1.
                const bool need = mx * sl2 > ms + rescale_margin;
                const float ms_new = need ? mx * sl2 : ms;
                const bool resc = need && kt > 0 && ms != -INFINITY;
                const float f = resc ? ex2_approx(ms - ms_new) : 1.0f;
                if (__any_sync(0xffffffffu, resc)) {
                    l *= f;
#pragma unroll 1
                    for (int c = 0; c < 2; ++c) {
                        uint32_t o[32];
                        tmem_ld_32x32b_x32(t_o1 + c * 32, o);
                        tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 32; ++j) o[j] = __float_as_uint(__uint_as_float(o[j]) * f);
                        tmem_st_32x32b_x32(t_o1 + c * 32, o);
                    }
                    tmem_st_wait();
                }
                ms = ms_new;
            }
#pragma unroll
            for (int c = 0; c < 2; ++c) {               // 16 keys at a time: registers for two CTAs per SM
                uint32_t pw[8], dw[8];
#pragma unroll
                for (int j = 0; j < 16; j += 2) {
                    float pv[2], dv[2];
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int kl = c * 16 + j + e;
                        const bool valid = kl < kv_valid;
                        const float sv = __uint_as_float(v0[c * 16 + j + e]);
                        pv[e] = valid ? ex2_approx(fmaf(sv, sl2, -ms)) : 0.f;
                        l += pv[e];
                        if constexpr (ALIBI) {
                            const float2 ck = lds_f2(sC_addr + ((kt & 1) * 64 + half * 32 + kl) * 8);
                            const float dx = cq.x - ck.x, dy = cq.y - ck.y;
                            dv[e] = valid ? sqrt_approx(fmaf(dx, dx, dy * dy)) * slope : 0.f;
                        }
                    }
                    pw[j >> 1] = pack_op<TRAIN>(pv[0], pv[1]);
                    if constexpr (ALIBI) dw[j >> 1] = pack_op<TRAIN>(dv[0], dv[1]);
                }
#pragma unroll
                for (int q = 0; q < 2; ++q) {               // 16 keys = two 16-byte chunks of the 128-byte tile row
                    const int off = ((half * 4 + c * 2 + q) ^ (r & 7)) * 16;
                    sts_u4(pb + off, make_uint4(pw[4 * q], pw[4 * q + 1], pw[4 * q + 2], pw[4 * q + 3]));
                    if constexpr (ALIBI)
                        sts_u4(db + off, make_uint4(dw[4 * q], dw[4 * q + 1], dw[4 * q + 2], dw[4 * q + 3]));
                }
            }
            fence_proxy_async();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(pfull);
        }
        // ---- merge the two halves of every row: weights w_h = exp2(m_h - max(m_a, m_b)) ----
        sX[half * 128 + r] = ms;
        sX[256 + half * 128 + r] = l;
        asm volatile("bar.sync 1, 256;\n" ::: "memory");
        const float ms_o = sX[(half ^ 1) * 128 + r], l_o = sX[256 + (half ^ 1) * 128 + r];
        const float M = fmaxf(ms, ms_o);
        const float w_me = (ms == -INFINITY) ? 0.f : ex2_approx(ms - M);
        const float w_ot = (ms_o == -INFINITY) ? 0.f : ex2_approx(ms_o - M);
        const float w_a = half == 0 ? w_me : w_ot, w_b = half == 0 ? w_ot : w_me;
        const float lt = l * w_me + l_o * w_ot;
        const bool has_b = true;                   // both accumulators are written by every tile (zeros past the bag)

        mbar_wait(ofull, 0);
        tc_fence_after();
        const float inv = 1.0f / lt;
        const long long obase = b * p.out_batch_stride + static_cast<long long>(row) * p.out_row_stride + h * 64 + half * 32;
        {
            uint32_t oa[32], ob[32], o2[32];
            tmem_ld_32x32b_x32(t_lane + COL_O1A + half * 32, oa);
            tmem_ld_32x32b_x32(t_lane + COL_O1B + half * 32, ob);
            if constexpr (ALIBI) tmem_ld_32x32b_x32(t_lane + COL_O2 + half * 32, o2);
            tmem_ld_wait();
            if (row < S) {
                float sm[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    float v = __uint_as_float(oa[j]) * w_a;
                    if (has_b) v = fmaf(__uint_as_float(ob[j]), w_b, v);
                    sm[j] = v * inv;
                }
                if constexpr (TRAIN) {
                    if (half == 0) t.lse2[(static_cast<long long>(b) * p.H + h) * S + row] = M + log2f(lt);
#pragma unroll
                    for (int j = 0; j < 32; j += 8) {
                        float y[8];
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            y[e] = sm[j + e];
                            if constexpr (ALIBI) y[e] = fmaf(-descale, __uint_as_float(o2[j + e]), sm[j + e]);
                        }
                        *reinterpret_cast<uint4*>(t.out16 + obase + j) =
                            make_uint4(pack_bf16(y[0], y[1]), pack_bf16(y[2], y[3]), pack_bf16(y[4], y[5]), pack_bf16(y[6], y[7]));
                        *reinterpret_cast<float4*>(t.osm + obase + j) = make_float4(sm[j], sm[j + 1], sm[j + 2], sm[j + 3]);
                        *reinterpret_cast<float4*>(t.osm + obase + j + 4) = make_float4(sm[j + 4], sm[j + 5], sm[j + 6], sm[j + 7]);
                        if constexpr (ALIBI) {
                            *reinterpret_cast<float4*>(t.odv + obase + j) =
                                make_float4(__uint_as_float(o2[j]), __uint_as_float(o2[j + 1]), __uint_as_float(o2[j + 2]), __uint_as_float(o2[j + 3]));
                            *reinterpret_cast<float4*>(t.odv + obase + j + 4) =
                                make_float4(__uint_as_float(o2[j + 4]), __uint_as_float(o2[j + 5]), __uint_as_float(o2[j + 6]), __uint_as_float(o2[j + 7]));
                        }
                    }
                } else {
                    float y[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        y[j] = sm[j];
                        if constexpr (ALIBI) y[j] = fmaf(-descale, __uint_as_float(o2[j]), y[j]);
                    }
                    if (p.out_f32) {
                        float* of = reinterpret_cast<float*>(p.out) + obase;
                        float* ol = (p.out_lo != nullptr) ? p.out_lo + obase : nullptr;
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            const float4 hi = make_float4(round_tf32(y[j]), round_tf32(y[j + 1]), round_tf32(y[j + 2]), round_tf32(y[j + 3]));
                            *reinterpret_cast<float4*>(of + j) = hi;
                            if (ol != nullptr)
                                *reinterpret_cast<float4*>(ol + j) = make_float4(round_tf32(y[j] - hi.x), round_tf32(y[j + 1] - hi.y),
                                                                                 round_tf32(y[j + 2] - hi.z), round_tf32(y[j + 3] - hi.w));
                        }
                    } else {
                        __half* oh = reinterpret_cast<__half*>(p.out) + obase;
#pragma unroll
                        for (int j = 0; j < 32; j += 8)
                            *reinterpret_cast<uint4*>(oh + j) = make_uint4(pack_f16(y[j], y[j + 1]), pack_f16(y[j + 2], y[j + 3]),
                                                                           pack_f16(y[j + 4], y[j + 5]), pack_f16(y[j + 6], y[j + 7]));
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem, 256);
    }
}


int g_mil_tc_enabled = 1;
int g_mil_eager_rescale = 0;   // tests: rescale whenever a row maximum grows

template <bool ALIBI, bool TRAIN>
int launch_mil(const CUtensorMap& tm_qk, const CUtensorMap& tm_v, const CUtensorMap& tm_k64, const CUtensorMap& tm_v64,
               const AttnParams& p, int k_col0, const MilTrainOut& t, cudaStream_t stream) {
    static bool configured = false;
    if (!configured) {
        if (cudaFuncSetAttribute(mil_attn_tc2_kernel<ALIBI, TRAIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Mt2Smem::total) != cudaSuccess)
            return SB_ERR_CUDA;
        configured = true;
    }
    dim3 grid(p.B * p.H, (p.S + 127) / 128);
    ProfScope prof(PROF_ATTN, 4.0 * p.B * p.H * static_cast<double>(p.S) * p.S * 64, stream);
    const float margin = g_mil_eager_rescale ? 0.f : 8.f;
    mil_attn_tc2_kernel<ALIBI, TRAIN><<<grid, MT_THREADS, Mt2Smem::total, stream>>>(tm_qk, tm_k64, tm_v64, p, k_col0, t, margin);
    count_launch();
    return cudaGetLastError() == cudaSuccess ? SB_OK : SB_ERR_CUDA;
}

}  // namespace

void attention_mil_tc_enable(int on) {
    g_mil_tc_enabled = on & 1;
    g_mil_eager_rescale = (on >> 3) & 1;
}

// SB_ERR_UNSUPPORTED: outside this kernel's envelope (masked calls, head_dim != 64, short
// sequences that the single-pass kernel covers) -> the caller uses the general kernel
int attention_mil_tc_fwd(const AttnParams& p, int head_dim, cudaStream_t stream) {
    if (!g_mil_tc_enabled || head_dim != 64 || p.mask != nullptr || p.S <= 256) return SB_ERR_UNSUPPORTED;
    const long long koff = p.k - p.q;
    const long long v_rs = p.v_row_stride ? p.v_row_stride : p.row_stride;
    const long long v_bs = p.v_row_stride ? p.v_batch_stride : p.batch_stride;
    if (koff < 0 || koff + static_cast<long long>(p.H) * 64 > p.row_stride || (koff % 8) != 0 ||
        (reinterpret_cast<uintptr_t>(p.q) & 15) != 0 || (reinterpret_cast<uintptr_t>(p.v) & 15) != 0 ||
        (reinterpret_cast<uintptr_t>(p.out) & 15) != 0 || (p.out_row_stride % 8) != 0 ||
        static_cast<long long>(p.H) * 64 > v_rs || (p.S + 127) / 128 > 65535)
        return SB_ERR_UNSUPPORTED;
    const bool alibi = p.coords != nullptr;
    if (alibi != (p.out_f32 != 0)) return SB_ERR_UNSUPPORTED;
    CUtensorMap tm_qk, tm_v;
    int rc = make_tmap_3d_f16(&tm_qk, p.q, static_cast<int>(p.row_stride), p.S, p.B, p.row_stride, p.batch_stride, 64, 128);
    if (rc != SB_OK) return rc;
    rc = make_tmap_3d_f16(&tm_v, p.v, static_cast<int>(v_rs), p.S, p.B, v_rs, v_bs, 64, 128);
    if (rc != SB_OK) return rc;
    CUtensorMap tm_k64, tm_v64;
    rc = make_tmap_3d_f16(&tm_k64, p.q, static_cast<int>(p.row_stride), p.S, p.B, p.row_stride, p.batch_stride, 64, 64);
    if (rc != SB_OK) return rc;
    rc = make_tmap_3d_f16(&tm_v64, p.v, static_cast<int>(v_rs), p.S, p.B, v_rs, v_bs, 64, 64);
    if (rc != SB_OK) return rc;
    const MilTrainOut none{};
    return alibi ? launch_mil<true, false>(tm_qk, tm_v, tm_k64, tm_v64, p, static_cast<int>(koff), none, stream)
                 : launch_mil<false, false>(tm_qk, tm_v, tm_k64, tm_v64, p, static_cast<int>(koff), none, stream);
}

// training forward on the same kernel (bf16, extra outputs); SB_ERR_UNSUPPORTED -> mma.sync kernel
int attention_mil_tc_train_fwd(const AttnTrainParams& tp, int head_dim, cudaStream_t stream) {
    if (!g_mil_tc_enabled || head_dim != 64 || tp.S <= 256) return SB_ERR_UNSUPPORTED;
    const long long koff = tp.k - tp.q, voff = tp.v - tp.q;
    if (koff < 0 || koff + static_cast<long long>(tp.H) * 64 > tp.row_stride || (koff % 8) != 0 || voff < 0 ||
        (reinterpret_cast<uintptr_t>(tp.q) & 15) != 0 || (reinterpret_cast<uintptr_t>(tp.v) & 15) != 0 ||
        (reinterpret_cast<uintptr_t>(tp.out) & 15) != 0 || (reinterpret_cast<uintptr_t>(tp.osm) & 15) != 0 ||
        (tp.odv != nullptr && (reinterpret_cast<uintptr_t>(tp.odv) & 15) != 0) || (tp.out_row_stride % 8) != 0 ||
        (tp.out_batch_stride % 8) != 0 || (tp.S + 127) / 128 > 65535)
        return SB_ERR_UNSUPPORTED;
    AttnParams p{};
    p.q = reinterpret_cast<const __half*>(tp.q);
    p.k = reinterpret_cast<const __half*>(tp.k);
    p.v = reinterpret_cast<const __half*>(tp.v);
    p.row_stride = tp.row_stride; p.batch_stride = tp.batch_stride;
    p.out = tp.out; p.out_row_stride = tp.out_row_stride; p.out_batch_stride = tp.out_batch_stride;
    p.B = tp.B; p.S = tp.S; p.H = tp.H; p.scale_log2 = tp.scale_log2;
    p.coords = reinterpret_cast<const float*>(tp.coords);
    const MilTrainOut t{tp.out, tp.osm, tp.odv, tp.lse2, tp.beta, tp.inv_rm};
    CUtensorMap tm_qk, tm_v;
    int rc = make_tmap_3d_f16(&tm_qk, tp.q, static_cast<int>(tp.row_stride), tp.S, tp.B, tp.row_stride, tp.batch_stride, 64, 128);
    if (rc != SB_OK) return rc;
    rc = make_tmap_3d_f16(&tm_v, tp.v, static_cast<int>(tp.row_stride), tp.S, tp.B, tp.row_stride, tp.batch_stride, 64, 128);
    if (rc != SB_OK) return rc;
    CUtensorMap tm_k64, tm_v64;
    rc = make_tmap_3d_f16(&tm_k64, tp.q, static_cast<int>(tp.row_stride), tp.S, tp.B, tp.row_stride, tp.batch_stride, 64, 64);
    if (rc != SB_OK) return rc;
    rc = make_tmap_3d_f16(&tm_v64, tp.v, static_cast<int>(tp.row_stride), tp.S, tp.B, tp.row_stride, tp.batch_stride, 64, 64);
    if (rc != SB_OK) return rc;
    return tp.coords != nullptr ? launch_mil<true, true>(tm_qk, tm_v, tm_k64, tm_v64, p, static_cast<int>(koff), t, stream)
                                : launch_mil<false, true>(tm_qk, tm_v, tm_k64, tm_v64, p, static_cast<int>(koff), t, stream);
}

}  // namespace sb
