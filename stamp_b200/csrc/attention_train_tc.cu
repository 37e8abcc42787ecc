// tcgen05 backward of the MIL training attention for long bags (head_dim 64, S > 256); same math as the
// mma.sync kernels in attention_train.cu (see its header), same warp-specialised skeleton as the forward in
// attention_mil_tc.cu (warp 0 TMA, warp 1 MMA issue, warps 2..: NPART threads per TMEM lane).
//
// One kernel template, two roles.  A CTA owns a block of 128 rows (TMEM lanes) and streams 64-row tiles:
//
//   DKV : rows = keys.    block = (K, V),  stream = (Q, dO) tiles of 64 queries
//         X = K Q^T (S^T),  Y = V dO^T (dW^T)                      tcgen05, TMEM, double buffered
//         P^T = exp2(X*scale - lse_q),  W^T = P^T - beta Dhat^T,  dS^T = P^T (Y - delta_q)
//                                                                   -> bf16 tiles in the K-major 128B-swizzled UMMA layout
//         dV += W^T dO,   dK += dS^T Q                              accumulate in TMEM over the whole stream
//   DQ  : rows = queries. block = (Q, dO), stream = (K, V) tiles of 64 keys
//         X = Q K^T,  Y = dO V^T,  dS = P (Y - delta_row)           dQ += dS K
//
// The streamed tile is used twice from the same shared-memory bytes: as K-major B operand of X / Y
// (contraction over head_dim) and as MN-major B operand of the accumulating products (contraction over
// the 64 streamed rows).  TMEM: X | Y | acc0 | acc1 (64 columns each) = 256 columns and <= 102 registers per
// thread, so TWO CTAs share an SM: X / Y are single-buffered inside a CTA (read into registers, stage released
// before the arithmetic) and the other CTA's arithmetic covers the hand-offs -- one CTA alone left every
// pipe below 55 % (issue, MUFU, tensor) because each tile is a chain MMA -> TMEM load -> math -> smem -> MMA.
// Per (row, column) pair the CUDA cores spend one ex2 (+ one sqrt for ALiBi in DKV) -- as in the
// forward, the MUFU pipe bounds the kernel, the four tensor-core products hide under it.
#include <math.h>

#include "attention_train.cuh"
#include "common.cuh"
#include "gemm.cuh"

namespace sb {
namespace {

constexpr int NPART = 2;                       // compute threads per row (TMEM lane): 64 / NPART streamed columns each
constexpr int CW = 64 / NPART;                 // columns per compute thread
constexpr int BT_THREADS = 64 + NPART * 128;
constexpr int BLK_BYTES = 128 * 128;   // 128 rows x 64 bf16
constexpr int STR_BYTES = 64 * 128;    // 64 rows x 64 bf16

__device__ __forceinline__ float4 lds_f4(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];\n" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts_f4(uint32_t addr, float4 v) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};\n" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void sts_u4b(uint32_t addr, uint4 v) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};\n" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ float sqrt_apx(float x) {
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;\n" : "=f"(r) : "f"(x));
    return r;
}

__device__ __forceinline__ void tmem_ld_cols(uint32_t taddr, uint32_t (&r)[32]) { tmem_ld_32x32b_x32(taddr, r); }
__device__ __forceinline__ void tmem_ld_cols(uint32_t taddr, uint32_t (&r)[16]) { tmem_ld_32x32b_x16(taddr, r); }

struct BtSmem {
    static constexpr int off_b0 = 0;                              // block tile 0 (K or Q)
    static constexpr int off_b1 = BLK_BYTES;                      // block tile 1 (V or dO)
    static constexpr int off_s0 = 2 * BLK_BYTES;                  // streamed tile 0 (Q or K), 2 stages
    static constexpr int off_s1 = off_s0 + 2 * STR_BYTES;         // streamed tile 1 (dO or V), 2 stages
    static constexpr int off_w = off_s1 + 2 * STR_BYTES;          // W^T tile  [128 x 64] bf16 (DKV)
    static constexpr int off_ds = off_w + BLK_BYTES;              // dS tile   [128 x 64] bf16
    static constexpr int off_col = off_ds + BLK_BYTES;            // per-column {lse, delta, cx, cy}: 2 x 64 float4
    static constexpr int off_bar = off_col + 2 * 64 * 16;
    static constexpr int total = off_bar + 128 + 1024;
};

struct BtMaps {
    CUtensorMap qkv_blk, do_blk, qkv_str, do_str;
};

template <bool DKV, bool ALIBI>
__global__ void __launch_bounds__(BT_THREADS, 2)
attn_bwd_tc_kernel(const __grid_constant__ BtMaps tm, const AttnTrainParams p, int k_col0, int v_col0) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>(
        (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    uint8_t* sB0 = smem + BtSmem::off_b0;
    uint8_t* sB1 = smem + BtSmem::off_b1;
    uint8_t* sS0 = smem + BtSmem::off_s0;
    uint8_t* sS1 = smem + BtSmem::off_s1;
    uint8_t* sW = smem + BtSmem::off_w;
    uint8_t* sDS = smem + BtSmem::off_ds;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + BtSmem::off_bar);
    uint64_t* tfull = bars;        // [2] TMA -> MMA      (streamed tiles)
    uint64_t* tempty = bars + 2;   // [2] MMA -> TMA
    uint64_t* sfull = bars + 4;    // [2] MMA -> compute  (X, Y in TMEM)
    uint64_t* sempty = bars + 6;   // [2] compute -> MMA
    uint64_t* pfull = bars + 8;    // compute -> MMA      (W / dS tiles in smem)
    uint64_t* pempty = bars + 9;   // MMA -> compute
    uint64_t* ofull = bars + 10;
    uint64_t* bfull = bars + 11;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.x / p.H, h = blockIdx.x % p.H;
    const int r0 = blockIdx.y * 128;
    const int S = p.S;
    const int nt = (S + 63) / 64;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tm.qkv_blk);
        tma_prefetch_desc(&tm.do_blk);
        tma_prefetch_desc(&tm.qkv_str);
        tma_prefetch_desc(&tm.do_str);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tfull[i], 1);
            mbar_init(&tempty[i], 1);
            mbar_init(&sfull[i], 1);
            mbar_init(&sempty[i], NPART * 4);
        }
        mbar_init(pfull, NPART * 4);
        mbar_init(pempty, 1);
        mbar_init(ofull, 1);
        mbar_init(bfull, 1);
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
    constexpr uint32_t COL_Y = 64, COL_A0 = 128, COL_A1 = 192;   // X at column 0 (single stage, see header)

    if (warp == 0) {
        // ------------------------------------ TMA producer ------------------------------------
        if (lane == 0) {
            mbar_expect_tx(bfull, 2 * BLK_BYTES);
            if constexpr (DKV) {
                tma_load_3d(sB0, &tm.qkv_blk, bfull, k_col0 + h * 64, r0, b);   // K block
                tma_load_3d(sB1, &tm.qkv_blk, bfull, v_col0 + h * 64, r0, b);   // V block
            } else {
                tma_load_3d(sB0, &tm.qkv_blk, bfull, h * 64, r0, b);            // Q block
                tma_load_3d(sB1, &tm.do_blk, bfull, h * 64, r0, b);             // dO block
            }
            for (int j = 0; j < nt; ++j) {
                const int s = j & 1;
                mbar_wait(&tempty[s], ((j >> 1) & 1) ^ 1);
                mbar_expect_tx(&tfull[s], 2 * STR_BYTES);
                if constexpr (DKV) {
                    tma_load_3d(sS0 + s * STR_BYTES, &tm.qkv_str, &tfull[s], h * 64, j * 64, b);           // Q tile
                    tma_load_3d(sS1 + s * STR_BYTES, &tm.do_str, &tfull[s], h * 64, j * 64, b);            // dO tile
                } else {
                    tma_load_3d(sS0 + s * STR_BYTES, &tm.qkv_str, &tfull[s], k_col0 + h * 64, j * 64, b);  // K tile
                    tma_load_3d(sS1 + s * STR_BYTES, &tm.qkv_str, &tfull[s], v_col0 + h * 64, j * 64, b);  // V tile
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------ MMA issuer --------------------------------------
        if (lane == 0) {
            const uint32_t idesc_xy = umma_idesc_f16(128, 64, true, false, false);
            const uint32_t idesc_acc = umma_idesc_f16(128, 64, true, false, true);   // streamed tile: MN-major B
            const uint64_t b0_desc = umma_desc_k128(smem_u32(sB0));
            const uint64_t b1_desc = umma_desc_k128(smem_u32(sB1));
            const uint64_t w_desc = umma_desc_k128(smem_u32(sW));
            const uint64_t ds_desc = umma_desc_k128(smem_u32(sDS));
            mbar_wait(bfull, 0);
            auto issue_xy = [&](int j) {
                const int s = j & 1;                       // streamed smem tiles: two stages
                mbar_wait(&tfull[s], (j >> 1) & 1);
                mbar_wait(&sempty[0], (j & 1) ^ 1);        // X / Y in TMEM: one stage, read into registers early
                tc_fence_after();
                const uint64_t s0_desc = umma_desc_k128(smem_u32(sS0 + s * STR_BYTES));
                const uint64_t s1_desc = umma_desc_k128(smem_u32(sS1 + s * STR_BYTES));
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_f16_ss(tmem, b0_desc + 2 * k, s0_desc + 2 * k, idesc_xy, k != 0);
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_f16_ss(tmem + COL_Y, b1_desc + 2 * k, s1_desc + 2 * k, idesc_xy, k != 0);
                umma_commit(&sfull[0]);
            };
            issue_xy(0);
            for (int j = 0; j < nt; ++j) {
                if (j + 1 < nt) issue_xy(j + 1);
                const int s = j & 1;
                mbar_wait(pfull, j & 1);
                tc_fence_after();
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    // 16 streamed rows per step: 16 x 128 B = 2048 B of the tile, read as an MN-major operand
                    const uint64_t s0_mn = umma_desc_mn128(smem_u32(sS0 + s * STR_BYTES + k * 2048), 0);
                    umma_f16_ss(tmem + COL_A1, ds_desc + 2 * k, s0_mn, idesc_acc, (j | k) != 0);
                    if constexpr (DKV) {
                        const uint64_t s1_mn = umma_desc_mn128(smem_u32(sS1 + s * STR_BYTES + k * 2048), 0);
                        umma_f16_ss(tmem + COL_A0, w_desc + 2 * k, s1_mn, idesc_acc, (j | k) != 0);
                    }
                }
                umma_commit(pempty);
                umma_commit(&tempty[s]);
            }
            umma_commit(ofull);
        }
    } else {
        // ---------- compute: NPART threads per row (= TMEM lane), CW streamed columns each ----------
        const int quarter = warp & 3;
        const int half = (warp - 2) >> 2;          // which CW-column part of the 64 streamed columns
        const int r = quarter * 32 + lane;
        const int st = threadIdx.x - 64;           // 0 .. NPART*128-1 among the compute threads
        const int row = r0 + r;
        const uint32_t t_lane = tmem + (static_cast<uint32_t>(quarter * 32) << 16);
        const float sl2 = p.scale_log2;
        const uint32_t col_addr = smem_u32(smem + BtSmem::off_col);
        const uint32_t w_row = smem_u32(sW) + r * 128, ds_row = smem_u32(sDS) + r * 128;
        const long long sbase = (static_cast<long long>(b) * p.H + h) * S;
        const float2* cb = ALIBI ? p.coords + static_cast<long long>(b) * S : nullptr;

        float lse_r = INFINITY, dl_r = 0.f;        // DQ: per-row statistics
        float2 ck = make_float2(0.f, 0.f);         // DKV: this key's coordinates
        float inv_rm = 0.f, beta = 0.f;
        if constexpr (!DKV) {
            if (row < S) {
                lse_r = __ldg(p.lse2 + sbase + row);
                dl_r = __ldg(p.delta + sbase + row);
            }
        } else if constexpr (ALIBI) {
            if (row < S) ck = __ldg(cb + row);
            inv_rm = __ldg(p.inv_rm + h);
            beta = __ldg(p.beta + h);
        }

        for (int j = 0; j < nt; ++j) {
            const int s = j & 1;
            const int c_valid = min(64, S - j * 64) - half * CW;   // valid columns among this thread's CW
            if constexpr (DKV) {
                if (st < 64) {
                    const int q = j * 64 + st;
                    float4 v = make_float4(INFINITY, 0.f, 0.f, 0.f);
                    if (q < S) {
                        v.x = __ldg(p.lse2 + sbase + q);
                        v.y = __ldg(p.delta + sbase + q);
                        if constexpr (ALIBI) {
                            const float2 c = __ldg(cb + q);
                            v.z = c.x; v.w = c.y;
                        }
                    }
                    sts_f4(col_addr + ((j & 1) * 64 + st) * 16, v);
                }
                asm volatile("bar.sync 1, %0;\n" ::"n"(NPART * 128) : "memory");
            }
            mbar_wait(&sfull[0], j & 1);
            tc_fence_after();
#pragma unroll
            for (int ch = 0; ch < CW / 16; ++ch) {      // 16 columns at a time: registers for two CTAs per SM
                uint32_t x[16], y[16];
                const int col0 = half * CW + ch * 16;
                tmem_ld_32x32b_x16(t_lane + col0, x);
                tmem_ld_32x32b_x16(t_lane + COL_Y + col0, y);
                tmem_ld_wait();
                if (ch == CW / 16 - 1) {                 // X / Y of this tile are in registers: free the TMEM stage
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&sempty[0]);
                }
                uint32_t ww[8], dw[8];
#pragma unroll
                for (int c = 0; c < 16; c += 2) {
                    float wv[2], dv[2];
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const bool valid = (ch * 16 + c + e) < c_valid;
                        float lse = lse_r, dl = dl_r;
                        float4 cv;
                        if constexpr (DKV) {
                            cv = lds_f4(col_addr + ((j & 1) * 64 + col0 + c + e) * 16);
                            lse = cv.x; dl = cv.y;
                        }
                        const float pv = valid ? ex2_approx(fmaf(__uint_as_float(x[c + e]), sl2, -lse)) : 0.f;
                        dv[e] = pv * (__uint_as_float(y[c + e]) - dl);
                        if constexpr (DKV) {
                            float w = pv;
                            if constexpr (ALIBI) {
                                const float dx = ck.x - cv.z, dy = ck.y - cv.w;
                                const float dh = valid ? sqrt_apx(fmaf(dx, dx, dy * dy)) * inv_rm : 0.f;
                                w = fmaf(-beta, dh, pv);
                            }
                            wv[e] = w;
                        }
                    }
                    dw[c >> 1] = pack_bf16(dv[0], dv[1]);
                    if constexpr (DKV) ww[c >> 1] = pack_bf16(wv[0], wv[1]);
                }
                // the W / dS tiles are free once the accumulating products of the previous tile have retired;
                // waiting here (not before the arithmetic) lets them overlap this tile's exp / sqrt work
                if (ch == 0) mbar_wait(pempty, (j & 1) ^ 1);
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const int off = (((col0 >> 3) + q) ^ (r & 7)) * 16;   // 128 B per row, 128B swizzle
                    sts_u4b(ds_row + off, make_uint4(dw[4 * q], dw[4 * q + 1], dw[4 * q + 2], dw[4 * q + 3]));
                    if constexpr (DKV)
                        sts_u4b(w_row + off, make_uint4(ww[4 * q], ww[4 * q + 1], ww[4 * q + 2], ww[4 * q + 3]));
                }
            }
            fence_proxy_async();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(pfull);
        }

        // ---- epilogue: accumulators -> bf16 global; this thread stores CW of the row's 64 columns ----
        mbar_wait(ofull, 0);
        tc_fence_after();
        uint32_t a1[CW], a0[CW];
        tmem_ld_cols(t_lane + COL_A1 + half * CW, a1);
        if constexpr (DKV) tmem_ld_cols(t_lane + COL_A0 + half * CW, a0);
        tmem_ld_wait();
        if (row < S) {
            const long long o = b * p.batch_stride + static_cast<long long>(row) * p.row_stride + h * 64 + half * CW;
            uint16_t* d1 = (DKV ? p.dk : p.dq) + o;
#pragma unroll
            for (int jj = 0; jj < CW; jj += 8) {
                float v[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) v[e] = __uint_as_float(a1[jj + e]) * p.scale;
                *reinterpret_cast<uint4*>(d1 + jj) =
                    make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
                if constexpr (DKV) {
#pragma unroll
                    for (int e = 0; e < 8; ++e) v[e] = __uint_as_float(a0[jj + e]);
                    *reinterpret_cast<uint4*>(p.dv + o + jj) =
                        make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
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

template <bool DKV, bool ALIBI>
int launch_bt(const BtMaps& tm, const AttnTrainParams& p, int k_col0, int v_col0, cudaStream_t stream) {
    static bool configured = false;
    if (!configured) {
        if (cudaFuncSetAttribute(attn_bwd_tc_kernel<DKV, ALIBI>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 BtSmem::total) != cudaSuccess)
            return SB_ERR_CUDA;
        configured = true;
    }
    dim3 grid(p.B * p.H, (p.S + 127) / 128);
    attn_bwd_tc_kernel<DKV, ALIBI><<<grid, BT_THREADS, BtSmem::total, stream>>>(tm, p, k_col0, v_col0);
    count_launch();
    return cudaGetLastError() == cudaSuccess ? SB_OK : SB_ERR_CUDA;
}

int g_bwd_tc_enabled = 1;

}  // namespace

void attention_train_tc_enable(int on) { g_bwd_tc_enabled = on; }

// dK/dV and dQ on tcgen05; p.dout (bf16 dO) and p.delta must already be filled (attn_delta_kernel).
// SB_ERR_UNSUPPORTED: outside the envelope -> the caller runs the mma.sync kernels.
int attention_train_tc_bwd(const AttnTrainParams& p, int head_dim, cudaStream_t stream) {
    if (!g_bwd_tc_enabled || head_dim != 64 || p.S <= 256) return SB_ERR_UNSUPPORTED;
    const long long koff = p.k - p.q, voff = p.v - p.q;
    auto mis = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) != 0; };
    if (koff < 0 || voff < 0 || koff + static_cast<long long>(p.H) * 64 > p.row_stride ||
        voff + static_cast<long long>(p.H) * 64 > p.row_stride || (koff % 8) != 0 || (voff % 8) != 0 ||
        mis(p.q) || mis(p.dout) || mis(p.dq) || mis(p.dk) || mis(p.dv) || (p.row_stride % 8) != 0 ||
        (p.batch_stride % 8) != 0 || (p.out_row_stride % 8) != 0 || (p.out_batch_stride % 8) != 0 ||
        static_cast<long long>(p.H) * 64 > p.out_row_stride || (p.S + 127) / 128 > 65535)
        return SB_ERR_UNSUPPORTED;
    BtMaps tm;
    int rc = make_tmap_3d_f16(&tm.qkv_blk, p.q, static_cast<int>(p.row_stride), p.S, p.B, p.row_stride, p.batch_stride, 64, 128);
    if (rc != SB_OK) return rc;
    rc = make_tmap_3d_f16(&tm.qkv_str, p.q, static_cast<int>(p.row_stride), p.S, p.B, p.row_stride, p.batch_stride, 64, 64);
    if (rc != SB_OK) return rc;
    rc = make_tmap_3d_f16(&tm.do_blk, p.dout, static_cast<int>(p.out_row_stride), p.S, p.B, p.out_row_stride, p.out_batch_stride, 64, 128);
    if (rc != SB_OK) return rc;
    rc = make_tmap_3d_f16(&tm.do_str, p.dout, static_cast<int>(p.out_row_stride), p.S, p.B, p.out_row_stride, p.out_batch_stride, 64, 64);
    if (rc != SB_OK) return rc;
    const bool alibi = p.coords != nullptr;
    const int kc = static_cast<int>(koff), vc = static_cast<int>(voff);
    // algorithmic FLOPs of the reference backward: dV, dW, dQ, dK = four [S,S]x[S,hd] products per head
    ProfScope prof(PROF_ATTN, 8.0 * p.B * p.H * static_cast<double>(p.S) * p.S * 64, stream);
    rc = alibi ? launch_bt<true, true>(tm, p, kc, vc, stream) : launch_bt<true, false>(tm, p, kc, vc, stream);
    if (rc != SB_OK) return rc;
    return launch_bt<false, false>(tm, p, kc, vc, stream);
}

}  // namespace sb
