// Persistent, warp-specialised tcgen05 GEMM for sm_100a.
//
//   C[M,N] = epilogue( A[M,K] . B[N,K]^T )        A, B: K-major fp16/bf16, fp32 accumulate
//
// One CTA per SM loops over 128 x BN output tiles.
//   warp 0      : TMA producer  (cp.async.bulk.tensor, 128B swizzle, STAGES-deep mbarrier ring)
//   warp 1      : MMA issuer    (one lane issues tcgen05.mma kind::f16, accumulators in TMEM,
//                                two accumulator stages so the epilogue overlaps the next tile)
//   warps 2..9  : epilogue      (tcgen05.ld -> registers -> bias / activation / residual -> HBM)
//
// This is the contraction engine behind every dense layer of the hot path:
// the reference issues them as cuBLAS/ATen addmm calls from timm's ViT blocks
// (reference: src/stamp/preprocessing/__init__.py:325) and from the MIL aggregator
// (reference: src/stamp/modeling/models/vision_tranformer.py:137-139,153,163-167,314-318).
#include "gemm.cuh"

#include <mutex>

#include "common.cuh"

namespace sb {

namespace {

constexpr int BM = 128;
constexpr int BK_BYTES = 128;  // one 128-byte swizzle row of K per tile row: 64 x 16-bit or 32 x tf32
constexpr int UMMA_K_BYTES = 32;  // K extent of one tcgen05.mma: 16 x 16-bit or 8 x tf32
constexpr int NUM_EPI_WARPS = 8;
constexpr int NUM_THREADS = 64 + NUM_EPI_WARPS * 32;

template <int BN>
struct Cfg {
    static constexpr int STAGES = (BN == 256) ? 4 : 6;
    static constexpr int A_BYTES = BM * BK_BYTES;
    static constexpr int B_BYTES = BN * BK_BYTES;
    static constexpr int TMEM_COLS = 2 * BN;
    static constexpr int BAR_BYTES = (2 * STAGES + 4) * 8 + 16;
    static constexpr int SMEM_BYTES = STAGES * (A_BYTES + B_BYTES) + BAR_BYTES + 1024;
};

__device__ __forceinline__ float4 ldg4(const float* p) {
    return __ldg(reinterpret_cast<const float4*>(p));
}

// Epilogue for 32 consecutive columns [n0, n0+32) of one output row.
__device__ __forceinline__ void epilogue_chunk(const GemmParams& p, const uint32_t (&acc)[32],
                                               int m, long long row, int n0) {
    float v[32];
    const bool full = (n0 + 32 <= p.N);
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(acc[j]);

    if (p.bias != nullptr) {
        if (full) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
                float4 b = ldg4(p.bias + n0 + j);
                v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
            }
        } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
                if (n0 + j < p.N) v[j] += __ldg(p.bias + n0 + j);
        }
    }
    if (p.table != nullptr) {
        // fp32 addend applied BEFORE the activation: position table of the patch embedding, or the
        // running fp32 partial of a split-precision (hi + lo) product
        const float* t = p.table + static_cast<long long>(p.gin > 0 ? (m % p.gin) : m) * p.ldt + n0;
        if (full && ((reinterpret_cast<uintptr_t>(t) & 15) == 0)) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
                float4 tt = ldg4(t + j);
                v[j] += tt.x; v[j + 1] += tt.y; v[j + 2] += tt.z; v[j + 3] += tt.w;
            }
        } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
                if (n0 + j < p.N) v[j] += __ldg(t + j);
        }
    }
    if (p.act == ACT_GELU) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = gelu_erf(v[j]);
    } else if (p.act == ACT_RELU) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.0f);
    }

    const bool bf = p.bf16 != 0;
    switch (p.store) {
    case ST_16: {
        uint16_t* o = reinterpret_cast<uint16_t*>(p.out) + row * p.ldo + n0;
        if (full && ((reinterpret_cast<uintptr_t>(o) & 15) == 0)) {
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
                uint4 w;
                w.x = pack_16(v[j], v[j + 1], bf);
                w.y = pack_16(v[j + 2], v[j + 3], bf);
                w.z = pack_16(v[j + 4], v[j + 5], bf);
                w.w = pack_16(v[j + 6], v[j + 7], bf);
                *reinterpret_cast<uint4*>(o + j) = w;
                if (p.out_lo != nullptr) {  // low half of a split-precision operand (fp16 only)
                    const __half2* h = reinterpret_cast<const __half2*>(&w);
                    uint4 lo;
                    lo.x = pack_f16(v[j] - __low2float(h[0]), v[j + 1] - __high2float(h[0]));
                    lo.y = pack_f16(v[j + 2] - __low2float(h[1]), v[j + 3] - __high2float(h[1]));
                    lo.z = pack_f16(v[j + 4] - __low2float(h[2]), v[j + 5] - __high2float(h[2]));
                    lo.w = pack_f16(v[j + 6] - __low2float(h[3]), v[j + 7] - __high2float(h[3]));
                    *reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(p.out_lo) + row * p.ldo + n0 + j) = lo;
                }
            }
        } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
                if (n0 + j < p.N) {
                    uint32_t w = pack_16(v[j], 0.f, bf);
                    o[j] = static_cast<uint16_t>(w & 0xFFFF);
                }
        }
        break;
    }
    case ST_32: {
        float* o = reinterpret_cast<float*>(p.out) + row * p.ldo + n0;
        if (full && ((reinterpret_cast<uintptr_t>(o) & 15) == 0)) {
#pragma unroll
            for (int j = 0; j < 32; j += 4)
                *reinterpret_cast<float4*>(o + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
                if (n0 + j < p.N) o[j] = v[j];
        }
        break;
    }
    case ST_RESID32: {
        float* o = reinterpret_cast<float*>(p.out) + row * p.ldo + n0;
        if (full && ((reinterpret_cast<uintptr_t>(o) & 15) == 0)) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
                float4 x = *reinterpret_cast<const float4*>(o + j);
                float4 g = (p.gamma != nullptr) ? ldg4(p.gamma + n0 + j)
                                                : make_float4(1.f, 1.f, 1.f, 1.f);
                x.x = fmaf(g.x, v[j], x.x);
                x.y = fmaf(g.y, v[j + 1], x.y);
                x.z = fmaf(g.z, v[j + 2], x.z);
                x.w = fmaf(g.w, v[j + 3], x.w);
                *reinterpret_cast<float4*>(o + j) = x;
            }
        } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
                if (n0 + j < p.N) {
                    float g = (p.gamma != nullptr) ? __ldg(p.gamma + n0 + j) : 1.0f;
                    o[j] = fmaf(g, v[j], o[j]);
                }
        }
        break;
    }
    case ST_SWIGLU16:
    case ST_GATED16: {
        float g[16];
        if (p.store == ST_SWIGLU16) {
#pragma unroll
            for (int j = 0; j < 16; ++j) g[j] = silu(v[2 * j]) * v[2 * j + 1];
        } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) g[j] = tanhf(v[2 * j]) * sigmoidf_(v[2 * j + 1]);
        }
        uint16_t* o = reinterpret_cast<uint16_t*>(p.out) + row * p.ldo + (n0 >> 1);
        if (full && ((reinterpret_cast<uintptr_t>(o) & 15) == 0)) {
#pragma unroll
            for (int j = 0; j < 16; j += 8) {
                uint4 w;
                w.x = pack_16(g[j], g[j + 1], bf);
                w.y = pack_16(g[j + 2], g[j + 3], bf);
                w.z = pack_16(g[j + 4], g[j + 5], bf);
                w.w = pack_16(g[j + 6], g[j + 7], bf);
                *reinterpret_cast<uint4*>(o + j) = w;
                if (p.out_lo != nullptr) {
                    const __half2* h = reinterpret_cast<const __half2*>(&w);
                    uint4 lo;
                    lo.x = pack_f16(g[j] - __low2float(h[0]), g[j + 1] - __high2float(h[0]));
                    lo.y = pack_f16(g[j + 2] - __low2float(h[1]), g[j + 3] - __high2float(h[1]));
                    lo.z = pack_f16(g[j + 4] - __low2float(h[2]), g[j + 5] - __high2float(h[2]));
                    lo.w = pack_f16(g[j + 6] - __low2float(h[3]), g[j + 7] - __high2float(h[3]));
                    *reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(p.out_lo) + row * p.ldo + (n0 >> 1) + j) = lo;
                }
            }
        } else {
#pragma unroll
            for (int j = 0; j < 16; ++j)
                if (n0 + 2 * j + 1 < p.N) {
                    uint32_t w = pack_16(g[j], 0.f, bf);
                    o[j] = static_cast<uint16_t>(w & 0xFFFF);
                }
        }
        break;
    }
    default:
        break;
    }
}

template <int BN, bool TF32>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tn_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const GemmParams p) {
    using C = Cfg<BN>;
    extern __shared__ uint8_t smem_raw[];
    // 128B swizzle needs 1024-byte aligned tiles
    uint8_t* smem = reinterpret_cast<uint8_t*>(
        (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    uint8_t* sA = smem;
    uint8_t* sB = smem + C::STAGES * C::A_BYTES;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + C::STAGES * (C::A_BYTES + C::B_BYTES));
    uint64_t* empty_bar = full_bar + C::STAGES;
    uint64_t* tfull_bar = empty_bar + C::STAGES;
    uint64_t* tempty_bar = tfull_bar + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < C::STAGES; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tfull_bar[i], 1);
            mbar_init(&tempty_bar[i], NUM_EPI_WARPS);
        }
        fence_barrier_init();
    }
    if (warp == 1) {
        __syncwarp();
        tmem_alloc(tmem_slot, C::TMEM_COLS);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int num_m = (p.M + BM - 1) / BM;
    const int num_n = (p.N + BN - 1) / BN;
    const int num_tiles = num_m * num_n;
    constexpr int BK = TF32 ? 32 : 64;  // elements of K per pipeline stage
    const int num_kb = (p.K + BK - 1) / BK;

    if (warp == 0) {
        // ------------------------------ TMA producer ------------------------------
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
                const int m_blk = t / num_n, n_blk = t % num_n;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    mbar_expect_tx(&full_bar[stage], C::A_BYTES + C::B_BYTES);
                    tma_load_2d(sA + stage * C::A_BYTES, &tmA, &full_bar[stage], kb * BK, m_blk * BM);
                    tma_load_2d(sB + stage * C::B_BYTES, &tmB, &full_bar[stage], kb * BK, n_blk * BN);
                    if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------ MMA issuer --------------------------------
        if (lane == 0) {
            const uint32_t idesc = TF32 ? umma_idesc_tf32(BM, BN) : umma_idesc_f16(BM, BN, p.bf16 != 0, false, false);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
                mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BN;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint64_t a_desc = umma_desc_k128(smem_u32(sA + stage * C::A_BYTES));
                    const uint64_t b_desc = umma_desc_k128(smem_u32(sB + stage * C::B_BYTES));
#pragma unroll
                    for (int k = 0; k < BK_BYTES / UMMA_K_BYTES; ++k) {
                        // advancing 32 B along K inside the swizzle row: +2 in the
                        // 16-byte-granular start-address field
                        if constexpr (TF32)
                            umma_tf32_ss(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
                        else
                            umma_f16_ss(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
                    }
                    umma_commit(&empty_bar[stage]);  // smem slot reusable once these MMAs retire
                    if (kb == num_kb - 1) umma_commit(&tfull_bar[acc]);
                    if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
                }
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else {
        // ------------------------------ epilogue ----------------------------------
        const int ew = warp - 2;
        const int quarter = warp & 3;      // TMEM lane quarter this warp may access
        const int chalf = ew >> 2;         // which half of the BN columns
        constexpr int COLS_PER_WARP = BN / 2;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
            const int m_blk = t / num_n, n_blk = t % num_n;
            mbar_wait(&tfull_bar[acc], acc_phase);
            tc_fence_after();
            const int m = m_blk * BM + quarter * 32 + lane;
            long long row = m;
            if (p.gin > 0) row = static_cast<long long>(m / p.gin) * p.gout + p.goff + (m % p.gin);
            const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * BN;
#pragma unroll 1
            for (int c = 0; c < COLS_PER_WARP / 32; ++c) {
                const int col = chalf * COLS_PER_WARP + c * 32;
                const int n0 = n_blk * BN + col;
                if (n0 >= p.N) break;  // warp-uniform
                uint32_t r[32];
                tmem_ld_32x32b_x32(t_row + col, r);
                tmem_ld_wait();
                if (m < p.M) epilogue_chunk(p, r, m, row, n0);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty_bar[acc]);
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, C::TMEM_COLS);
    }
}

// ------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* sym = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) ==
                cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(sym);
    });
    return fn;
}

int make_tmap(CUtensorMap* tm, const void* ptr, int rows, int cols, long long ld, int box_rows,
              int kind /*0 fp16, 1 bf16, 2 fp32(tf32)*/) {
    EncodeTiledFn fn = get_encode_fn();
    if (fn == nullptr) return SB_ERR_DRIVER;
    cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
    const int esz = (kind == 2) ? 4 : 2;
    cuuint64_t gstr[1] = {static_cast<cuuint64_t>(ld) * esz};
    cuuint32_t box[2] = {static_cast<cuuint32_t>(BK_BYTES / esz), static_cast<cuuint32_t>(box_rows)};
    cuuint32_t estr[2] = {1, 1};
    const CUtensorMapDataType dt = (kind == 2)   ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                                   : (kind == 1) ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
                                                 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
    CUresult r = fn(tm, dt, 2,
                    const_cast<void*>(ptr), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? SB_OK : SB_ERR_DRIVER;
}

int g_num_sms = 0;

template <int BN, bool TF32>
int launch(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmParams& p, int tiles,
           cudaStream_t stream) {
    static bool configured = false;
    if (!configured) {
        if (cudaFuncSetAttribute(gemm_tn_kernel<BN, TF32>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 Cfg<BN>::SMEM_BYTES) != cudaSuccess)
            return SB_ERR_CUDA;
        configured = true;
    }
    int grid = tiles < g_num_sms ? tiles : g_num_sms;
    ProfScope prof(PROF_GEMM, 2.0 * p.M * static_cast<double>(p.N) * p.K, stream);
    gemm_tn_kernel<BN, TF32><<<grid, NUM_THREADS, Cfg<BN>::SMEM_BYTES, stream>>>(tmA, tmB, p);
    count_launch();
    return cudaGetLastError() == cudaSuccess ? SB_OK : SB_ERR_CUDA;
}

}  // namespace

int gemm_num_sms() {
    if (g_num_sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
        if (g_num_sms <= 0) g_num_sms = 148;
    }
    return g_num_sms;
}

int gemm_tn(const void* A, long long lda, const void* B, long long ldb, const GemmParams& p,
            cudaStream_t stream) {
    if (p.M <= 0 || p.N <= 0 || p.K <= 0 || A == nullptr || B == nullptr || p.out == nullptr)
        return SB_ERR_BAD_ARG;
    const int ealign = p.tf32 ? 4 : 8;  // 16-byte row pitch
    if ((lda % ealign) != 0 || (ldb % ealign) != 0 || (reinterpret_cast<uintptr_t>(A) & 15) != 0 ||
        (reinterpret_cast<uintptr_t>(B) & 15) != 0)
        return SB_ERR_BAD_ARG;
    if ((p.store == ST_SWIGLU16 || p.store == ST_GATED16) && (p.N % 2) != 0) return SB_ERR_BAD_ARG;
    gemm_num_sms();

    const int num_m = (p.M + BM - 1) / BM;
    // 128 x 256 tiles unless that leaves most SMs without a tile
    const int tiles256 = num_m * ((p.N + 255) / 256);
    const bool use256 = (p.N >= 256) && (tiles256 >= g_num_sms);
    const int bn = use256 ? 256 : 128;

    CUtensorMap tmA, tmB;
    const int kind = p.tf32 ? 2 : (p.bf16 ? 1 : 0);
    int rc = make_tmap(&tmA, A, p.M, p.K, lda, BM, kind);
    if (rc != SB_OK) return rc;
    rc = make_tmap(&tmB, B, p.N, p.K, ldb, bn, kind);
    if (rc != SB_OK) return rc;

    const int tiles128 = num_m * ((p.N + 127) / 128);
    if (p.tf32) return use256 ? launch<256, true>(tmA, tmB, p, tiles256, stream) : launch<128, true>(tmA, tmB, p, tiles128, stream);
    return use256 ? launch<256, false>(tmA, tmB, p, tiles256, stream) : launch<128, false>(tmA, tmB, p, tiles128, stream);
}

}  // namespace sb
