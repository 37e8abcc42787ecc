// Persistent, warp-specialised tcgen05 GEMM for sm_100a.
//
//   C[M,N] = epilogue( A[M,K] . B[N,K]^T )     A, B: K-major fp16 / bf16 / tf32, fp32 accumulate
//
// Two tile shapes share one kernel template:
//   CTAS = 1 : one CTA per SM, 128 x BN tile (BN = 128 | 256), tcgen05.mma.cta_group::1
//   CTAS = 2 : a CTA pair (cluster of 2, one TPC) owns a 256 x 256 tile: each CTA loads its own 128
//              rows of A and HALF of the B tile, the leader issues tcgen05.mma.cta_group::2
//              (M = 256) which reads both halves -- per-SM operand traffic from L2 drops by a third.
// Roles per CTA:
//   warp 0      : TMA producer  (cp.async.bulk.tensor, 128B swizzle, mbarrier ring)
//   warp 1      : MMA issuer    (one lane; accumulators in TMEM, two accumulator stages so the
//                                epilogue of tile i overlaps the mainloop of tile i+1)
//   warps 2..9  : epilogue      (tcgen05.ld -> per-warp swizzled smem transpose -> every global
//                                access is a full 128-byte line: bias / GELU / ReLU / SwiGLU /
//                                gate / LayerScale + residual RMW / fp32 addend / hi-lo split)
//
// This is the contraction engine behind every dense layer of the hot path: the reference issues
// them as cuBLAS/ATen addmm calls from timm's ViT blocks (src/stamp/preprocessing/__init__.py:325)
// and from the MIL aggregator (src/stamp/modeling/models/vision_tranformer.py:137-139,153,163-167,
// 314-318).
#include "gemm.cuh"

#include <cstdio>
#include <mutex>

#include "common.cuh"

namespace sb {

namespace {

constexpr int BM = 128;            // rows of A per CTA
constexpr int BK_BYTES = 128;      // one 128-byte swizzle row of K per tile row: 64 x 16-bit or 32 x tf32
constexpr int UMMA_K_BYTES = 32;   // K extent of one tcgen05.mma: 16 x 16-bit or 8 x tf32
constexpr int NUM_EPI_WARPS = 8;
constexpr int NUM_THREADS = 64 + NUM_EPI_WARPS * 32;
// Warp roles.  The single-thread TMA and MMA warps take the HIGHEST warp ids: a sub-partition's issue arbiter favours
// higher warp ids, and as warps 0 / 1 the two pipeline drivers were starved by an activation-heavy epilogue (fc1 + GELU:
// the epilogue warps waited for accumulators a quarter of the time while the tensor pipe idled at 62 %).
constexpr int WARP_TMA = NUM_EPI_WARPS, WARP_MMA = NUM_EPI_WARPS + 1;
constexpr int EPI_STAGE_BYTES = 32 * 32 * 4;  // per epilogue warp: one 32 x 32 fp32 block

template <int BN, int CTAS, int EPI = 0>
struct Cfg {
    static constexpr int B_ROWS = BN / CTAS;  // rows of the B tile this CTA loads
    static constexpr int A_BYTES = BM * BK_BYTES;
    static constexpr int B_BYTES = B_ROWS * BK_BYTES;
    // TMA epilogues double-buffer their staging tile (a store must have READ its tile before the tile is rewritten;
    // with one buffer that wait sat on the critical path of every chunk pair), paid for with one pipeline stage
    static constexpr int EPI_BUFS = (EPI != 0) ? 2 : 1;
    static constexpr int STAGES = ((A_BYTES + B_BYTES == 48 * 1024) ? 4 : 6) - (EPI != 0 ? 1 : 0);
    static constexpr int TMEM_COLS = 2 * BN;
    static constexpr int EPI_BYTES = NUM_EPI_WARPS * EPI_STAGE_BYTES * EPI_BUFS;
    static constexpr int BAR_BYTES = (2 * STAGES + 4) * 8 + 16;
    static constexpr int SMEM_BYTES = STAGES * (A_BYTES + B_BYTES) + EPI_BYTES + BAR_BYTES + 1024;
};

__device__ __forceinline__ float4 ldg4(const float* p) {
    return __ldg(reinterpret_cast<const float4*>(p));
}

// ---- cluster helpers (2-CTA mode) ------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
// shared::cluster address of `local` (a shared::cta address) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t map_to_cta(uint32_t local, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;\n" : "=r"(r) : "r"(local), "r"(rank));
    return r;
}
// relaxed: the only thing the leader's MMA thread must observe is that this warp's tcgen05.ld of
// the accumulator stage completed (tcgen05.wait::ld + tcgen05.fence before this arrive); a
// cluster-scope release would additionally wait for all of the warp's outstanding global stores
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];\n" ::"r"(cluster_addr)
                 : "memory");
}
__device__ __forceinline__ void tma_load_2d_cta2(uint32_t smem_dst, const CUtensorMap* tm,
                                                 uint32_t leader_bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4}], [%2];\n" ::"r"(smem_dst),
        "l"(reinterpret_cast<uint64_t>(tm)), "r"(leader_bar), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* smem_result, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(
                     smem_u32(smem_result)),
                 "r"(ncols)
                 : "memory");
}
__device__ __forceinline__ void tmem_relinquish2() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;\n" ::"r"(taddr), "r"(ncols)
                 : "memory");
}
__device__ __forceinline__ void umma_f16_ss_cta2(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                                 uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive (once) on the barrier at the same smem offset in both CTAs of the pair
__device__ __forceinline__ void umma_commit_cta2(uint64_t* bar) {
    const uint16_t mask = 0x3;
    asm volatile(
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 "
        "[%0], %1;\n" ::"r"(smem_u32(bar)),
        "h"(mask)
        : "memory");
}

// ---- TMA stores of the epilogue (bulk async-group completion) -----------------------------------
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* tm, uint32_t smem_src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];\n" ::"l"(
                     reinterpret_cast<uint64_t>(tm)),
                 "r"(smem_src), "r"(c0), "r"(c1)
                 : "memory");
}
// global[tile] += smem[tile], element type of the tensor map (fp32); performed by the L2 atomics units
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* tm, uint32_t smem_src, int c0, int c1) {
    asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.bulk_group [%0, {%2, %3}], [%1];\n" ::"l"(
                     reinterpret_cast<uint64_t>(tm)),
                 "r"(smem_src), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;\n" ::: "memory"); }
// the staging tile may be overwritten once the previous bulk operations have READ it
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;\n" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory"); }

// ---- epilogue for 4 consecutive columns [n, n+4) of output row `orow` (input row m) ------------
// (slow path: only the ragged right edge of a matrix with N % 4 != 0 comes through here)
// (inlined exactly once, in a non-unrolled loop: a real call would force the kernel parameter
//  struct into local memory and turn every p.field read of the hot path into an LDL)
__device__ __forceinline__ void epilogue4(const GemmParams& p, float4 acc, int m, long long orow, int n,
                                          float4 bias4, float4 gamma4) {
    float v[4] = {acc.x + bias4.x, acc.y + bias4.y, acc.z + bias4.z, acc.w + bias4.w};
    const bool full = (n + 4 <= p.N);
    if (p.table != nullptr) {
        // fp32 addend applied BEFORE the activation: position table of the patch embedding, or the
        // running fp32 partial of a split-precision (hi + lo) product
        const float* t = p.table + static_cast<long long>(p.gin > 0 ? (m % p.gin) : m) * p.ldt + n;
        if (full && ((reinterpret_cast<uintptr_t>(t) & 15) == 0)) {
            const float4 tt = ldg4(t);
            v[0] += tt.x; v[1] += tt.y; v[2] += tt.z; v[3] += tt.w;
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (n + j < p.N) v[j] += __ldg(t + j);
        }
    }
    if (p.act == ACT_GELU) {
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = gelu_erf(v[j]);
    } else if (p.act == ACT_RELU) {
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = fmaxf(v[j], 0.0f);
    }
    const bool bf = p.bf16 != 0;
    switch (p.store) {
    case ST_16: {
        uint16_t* o = reinterpret_cast<uint16_t*>(p.out) + orow * p.ldo + n;
        if (full && ((reinterpret_cast<uintptr_t>(o) & 7) == 0)) {
            uint2 w;
            w.x = pack_16(v[0], v[1], bf);
            w.y = pack_16(v[2], v[3], bf);
            *reinterpret_cast<uint2*>(o) = w;
            if (p.out_lo != nullptr) {  // low half of a split-precision operand (fp16 only)
                const __half2* h = reinterpret_cast<const __half2*>(&w);
                uint2 lo;
                lo.x = pack_f16(v[0] - __low2float(h[0]), v[1] - __high2float(h[0]));
                lo.y = pack_f16(v[2] - __low2float(h[1]), v[3] - __high2float(h[1]));
                *reinterpret_cast<uint2*>(reinterpret_cast<uint16_t*>(p.out_lo) + orow * p.ldo + n) = lo;
            }
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (n + j < p.N) {
                    const uint32_t w = pack_16(v[j], 0.f, bf);
                    o[j] = static_cast<uint16_t>(w & 0xFFFF);
                    if (p.out_lo != nullptr) {
                        const float hi = __half2float(__ushort_as_half(static_cast<unsigned short>(w & 0xFFFF)));
                        reinterpret_cast<__half*>(p.out_lo)[orow * p.ldo + n + j] = __float2half_rn(v[j] - hi);
                    }
                }
        }
        break;
    }
    case ST_32: {
        float* o = reinterpret_cast<float*>(p.out) + orow * p.ldo + n;
        if (full && ((reinterpret_cast<uintptr_t>(o) & 15) == 0)) {
            *reinterpret_cast<float4*>(o) = make_float4(v[0], v[1], v[2], v[3]);
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (n + j < p.N) o[j] = v[j];
        }
        break;
    }
    case ST_RESID32: {
        float* o = reinterpret_cast<float*>(p.out) + orow * p.ldo + n;
        if (full && ((reinterpret_cast<uintptr_t>(o) & 15) == 0)) {
            float4 x = *reinterpret_cast<const float4*>(o);
            x.x = fmaf(gamma4.x, v[0], x.x);
            x.y = fmaf(gamma4.y, v[1], x.y);
            x.z = fmaf(gamma4.z, v[2], x.z);
            x.w = fmaf(gamma4.w, v[3], x.w);
            *reinterpret_cast<float4*>(o) = x;
        } else {
            const float g[4] = {gamma4.x, gamma4.y, gamma4.z, gamma4.w};
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (n + j < p.N) o[j] = fmaf(g[j], v[j], o[j]);
        }
        break;
    }
    case ST_SWIGLU16:
    case ST_GATED16: {
        float g0, g1;
        if (p.store == ST_SWIGLU16) {
            g0 = silu(v[0]) * v[1];
            g1 = silu(v[2]) * v[3];
        } else {
            g0 = tanhf(v[0]) * sigmoidf_(v[1]);
            g1 = tanhf(v[2]) * sigmoidf_(v[3]);
        }
        uint16_t* o = reinterpret_cast<uint16_t*>(p.out) + orow * p.ldo + (n >> 1);
        if (full) {
            const uint32_t w = pack_16(g0, g1, bf);
            *reinterpret_cast<uint32_t*>(o) = w;
            if (p.out_lo != nullptr) {
                const __half2 h = *reinterpret_cast<const __half2*>(&w);
                *reinterpret_cast<uint32_t*>(reinterpret_cast<uint16_t*>(p.out_lo) + orow * p.ldo + (n >> 1)) =
                    pack_f16(g0 - __low2float(h), g1 - __high2float(h));
            }
        } else {
            const float g[2] = {g0, g1};
#pragma unroll
            for (int j = 0; j < 2; ++j)
                if (n + 2 * j + 1 < p.N) {
                    const uint32_t w = pack_16(g[j], 0.f, bf);
                    o[j] = static_cast<uint16_t>(w & 0xFFFF);
                    if (p.out_lo != nullptr) {
                        const float hi = __half2float(__ushort_as_half(static_cast<unsigned short>(w & 0xFFFF)));
                        reinterpret_cast<__half*>(p.out_lo)[orow * p.ldo + (n >> 1) + j] = __float2half_rn(g[j] - hi);
                    }
                }
        }
        break;
    }
    default:
        break;
    }
}

template <int BN, bool TF32, int CTAS, int EPI>
// 10 warps: one SM sub-partition (16K registers) hosts 3 of them -> at most 168 registers per thread
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tn_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmC, const GemmParams p) {
    using C = Cfg<BN, CTAS, EPI>;
    extern __shared__ uint8_t smem_raw[];
    // 128B swizzle needs 1024-byte aligned tiles (identical offset in both CTAs of a pair)
    uint8_t* smem = reinterpret_cast<uint8_t*>(
        (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    uint8_t* sA = smem;
    uint8_t* sB = smem + C::STAGES * C::A_BYTES;
    float* sEpi = reinterpret_cast<float*>(smem + C::STAGES * (C::A_BYTES + C::B_BYTES));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + C::STAGES * (C::A_BYTES + C::B_BYTES) + C::EPI_BYTES);
    uint64_t* empty_bar = full_bar + C::STAGES;
    uint64_t* tfull_bar = empty_bar + C::STAGES;
    uint64_t* tempty_bar = tfull_bar + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t cta_rank = (CTAS == 2) ? cluster_ctarank() : 0u;
    const bool leader = cta_rank == 0;

    if (warp == WARP_TMA && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        if constexpr (EPI != 0) tma_prefetch_desc(&tmC);
    }
    if (warp == WARP_MMA && lane == 0) {
        for (int i = 0; i < C::STAGES; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tfull_bar[i], 1);
            mbar_init(&tempty_bar[i], NUM_EPI_WARPS * CTAS);  // the leader hears both CTAs' epilogues
        }
        fence_barrier_init();
    }
    if (warp == WARP_MMA) {
        __syncwarp();
        if constexpr (CTAS == 2) { tmem_alloc2(tmem_slot, C::TMEM_COLS); tmem_relinquish2(); }
        else { tmem_alloc(tmem_slot, C::TMEM_COLS); tmem_relinquish(); }
    }
    tc_fence_before();
    if constexpr (CTAS == 2) cluster_sync_all(); else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    constexpr int BK = TF32 ? 32 : 64;  // elements of K per pipeline stage
    constexpr int TILE_M = BM * CTAS;
    const int num_m = (p.M + TILE_M - 1) / TILE_M;
    const int num_n = (p.N + BN - 1) / BN;
    const int num_tiles = num_m * num_n;
    const int num_kb = (p.K + BK - 1) / BK;
    const int first_tile = blockIdx.x / CTAS;
    const int tile_step = gridDim.x / CTAS;

    if (warp == WARP_TMA) {
        // ------------------------------ TMA producer ------------------------------
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int t = first_tile; t < num_tiles; t += tile_step) {
                const int m_blk = t / num_n, n_blk = t % num_n;
                const int row_a = m_blk * TILE_M + static_cast<int>(cta_rank) * BM;
                const int row_b = n_blk * BN + static_cast<int>(cta_rank) * C::B_ROWS;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    // K-concatenated split-precision products read A = [hi | lo] as hi, lo, hi
                    const int ka = (p.a_kwrap > 0) ? (kb * BK) % p.a_kwrap : kb * BK;
                    if constexpr (CTAS == 2) {
                        // both CTAs' bytes complete on the LEADER's barrier, which expects all of them
                        const uint32_t lbar = map_to_cta(smem_u32(&full_bar[stage]), 0);
                        if (leader) mbar_expect_tx(&full_bar[stage], 2 * (C::A_BYTES + C::B_BYTES));
                        tma_load_2d_cta2(smem_u32(sA + stage * C::A_BYTES), &tmA, lbar, ka, row_a);
                        tma_load_2d_cta2(smem_u32(sB + stage * C::B_BYTES), &tmB, lbar, kb * BK, row_b);
                    } else {
                        mbar_expect_tx(&full_bar[stage], C::A_BYTES + C::B_BYTES);
                        tma_load_2d(sA + stage * C::A_BYTES, &tmA, &full_bar[stage], ka, row_a);
                        tma_load_2d(sB + stage * C::B_BYTES, &tmB, &full_bar[stage], kb * BK, row_b);
                    }
                    if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == WARP_MMA) {
        // ------------------------------ MMA issuer (leader CTA) --------------------
        if (lane == 0 && leader) {
            const uint32_t idesc = TF32 ? umma_idesc_tf32(TILE_M, BN)
                                        : umma_idesc_f16(TILE_M, BN, p.bf16 != 0, false, false);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int t = first_tile; t < num_tiles; t += tile_step) {
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
                        const uint32_t accum = (kb | k) != 0 ? 1u : 0u;
                        if constexpr (CTAS == 2) umma_f16_ss_cta2(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, accum);
                        else if constexpr (TF32) umma_tf32_ss(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, accum);
                        else umma_f16_ss(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, accum);
                    }
                    // smem slot reusable (in both CTAs) once these MMAs retire
                    if constexpr (CTAS == 2) umma_commit_cta2(&empty_bar[stage]); else umma_commit(&empty_bar[stage]);
                    if (kb == num_kb - 1) {
                        if constexpr (CTAS == 2) umma_commit_cta2(&tfull_bar[acc]); else umma_commit(&tfull_bar[acc]);
                    }
                    if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
                }
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else {
        // ------------------------------ epilogue ----------------------------------
        const int ew = warp;
        const int quarter = warp & 3;      // TMEM lane quarter this warp may access
        const int chalf = ew >> 2;         // which half of the BN columns
        constexpr int COLS_PER_WARP = BN / 2;
        const uint32_t stg = smem_u32(sEpi) + ew * EPI_STAGE_BYTES * C::EPI_BUFS;  // shared-space byte address
        const int seg = lane & 7;          // 16-byte column segment this lane owns when reading back
        const int rsub = lane >> 3;        // row (mod 4) this lane owns when reading back
        const uint32_t tempty_leader = (CTAS == 2) ? map_to_cta(smem_u32(&tempty_bar[0]), 0) : 0u;
        int acc = 0;
        uint32_t acc_phase = 0;
        [[maybe_unused]] uint32_t epi_seq = 0;   // bulk operations issued by this warp so far (staging-tile parity)
        for (int t = first_tile; t < num_tiles; t += tile_step) {
            const int m_blk = t / num_n, n_blk = t % num_n;
            mbar_wait(&tfull_bar[acc], acc_phase);
            tc_fence_after();
            const int m_base = m_blk * TILE_M + static_cast<int>(cta_rank) * BM + quarter * 32;
            // output row offsets in elements (32-bit row index x ldo fits 63 bits; kept as
            // row indices to save registers: the epilogue is right at the 168-register cap)
            int orow[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int m = m_base + rsub + 4 * k;
                orow[k] = (p.gin > 0) ? (m / p.gin) * p.gout + p.goff + (m % p.gin) : m;
            }
            const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * BN;
            if constexpr (EPI != 0) {
                // ---- one thread per accumulator row; the output leaves through a TMA store / reduce-add ----
                // Two staging tiles per warp, used alternately: before a tile is rewritten only the bulk operation
                // issued BEFORE the most recent one has to have read its source (wait_group.read 1).
                const int sw = lane & 7;                   // 128B swizzle: 16-byte chunk index ^ (row & 7)
                const bool swiglu = (EPI == 1) && p.store == ST_SWIGLU16;
                // chunks of 32 columns this warp owns inside the matrix (warp-uniform)
                const int n_first = n_blk * BN + chalf * COLS_PER_WARP;
                const int nch = (p.N <= n_first) ? 0 : min(COLS_PER_WARP / 32, (p.N - n_first + 31) / 32);
                // The accumulator chunk c+1 is fetched from tensor memory WHILE chunk c goes through the activation
                // (tcgen05.ld is asynchronous until tcgen05.wait::ld; under a running mainloop a 4 KB load takes far
                // longer than its 256 cycles of port time -- ncu: two thirds of the epilogue warps' samples sat in that
                // scoreboard), and the TMEM stage is handed back as soon as the LAST chunk is in registers.
                auto chunk = [&](int c, uint32_t (&r)[32], uint32_t (&rn)[32]) {
                    const int col = chalf * COLS_PER_WARP + c * 32;
                    const int n0 = n_blk * BN + col;
                    tmem_ld_wait();
                    if (c + 1 < nch) {
                        tmem_ld_32x32b_x32(t_row + col + 32, rn);
                    } else {
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) {
                            if constexpr (CTAS == 2) mbar_arrive_cluster(tempty_leader + acc * 8);
                            else mbar_arrive(&tempty_bar[acc]);
                        }
                    }
                    // the bias of the 32 columns is the same for every lane: uniform (broadcast) loads
                    float4 b4[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        b4[j] = (p.bias != nullptr && n0 + 4 * j < p.N) ? ldg4(p.bias + n0 + 4 * j)
                                                                         : make_float4(0.f, 0.f, 0.f, 0.f);
                    float v[32];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        v[4 * j] = __uint_as_float(r[4 * j]) + b4[j].x;
                        v[4 * j + 1] = __uint_as_float(r[4 * j + 1]) + b4[j].y;
                        v[4 * j + 2] = __uint_as_float(r[4 * j + 2]) + b4[j].z;
                        v[4 * j + 3] = __uint_as_float(r[4 * j + 3]) + b4[j].w;
                    }
                    if (p.act == ACT_GELU) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = gelu_erf(v[j]);
                    } else if (p.act == ACT_RELU) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.0f);
                    }
                    if constexpr (EPI == 1) {
                        const bool bf = p.bf16 != 0;
                        if (swiglu) {
                            // column pairs (x1, x2) -> one output: 16 outputs = 32 bytes of the row per chunk;
                            // the four chunks of this warp fill one 64-output (128-byte) row
                            const uint32_t sbuf = stg + (epi_seq & 1) * EPI_STAGE_BYTES;
                            const uint32_t srow = sbuf + lane * 128;
                            if (c == 0) {
                                if (lane == 0) bulk_wait_read1();
                                __syncwarp();
                            }
                            uint32_t w[8];
#pragma unroll
                            for (int j = 0; j < 8; ++j)
                                w[j] = pack_16(silu(v[4 * j]) * v[4 * j + 1], silu(v[4 * j + 2]) * v[4 * j + 3], bf);
                            sts_v4(srow + (((2 * c) ^ sw) * 16), make_uint4(w[0], w[1], w[2], w[3]));
                            sts_v4(srow + (((2 * c + 1) ^ sw) * 16), make_uint4(w[4], w[5], w[6], w[7]));
                            const bool last = (c + 1 == nch);
                            if (last) {
                                fence_proxy_async();
                                __syncwarp();
                                if (lane == 0) {
                                    tma_store_2d(&tmC, sbuf, (n_blk * BN + chalf * COLS_PER_WARP) >> 1, m_base);
                                    bulk_commit();
                                }
                                ++epi_seq;
                            }
                        } else {
                            // 32 outputs = 64 bytes: two chunks fill one 64-column (128-byte) staging row
                            const uint32_t sbuf = stg + (epi_seq & 1) * EPI_STAGE_BYTES;
                            const uint32_t srow = sbuf + lane * 128;
                            if ((c & 1) == 0) {
                                if (lane == 0) bulk_wait_read1();
                                __syncwarp();
                            }
                            uint32_t w[16];
#pragma unroll
                            for (int j = 0; j < 16; ++j) w[j] = pack_16(v[2 * j], v[2 * j + 1], bf);
#pragma unroll
                            for (int j = 0; j < 4; ++j)
                                sts_v4(srow + ((((c & 1) * 4 + j) ^ sw) * 16),
                                       make_uint4(w[4 * j], w[4 * j + 1], w[4 * j + 2], w[4 * j + 3]));
                            if ((c & 1) == 1 || c + 1 == nch) {
                                fence_proxy_async();
                                __syncwarp();
                                if (lane == 0) {
                                    tma_store_2d(&tmC, sbuf, n0 - (c & 1) * 32, m_base);
                                    bulk_commit();
                                }
                                ++epi_seq;
                            }
                        }
                    } else {
                        // x += gamma * (acc + bias): 32 fp32 columns = one 128-byte staging row per chunk
                        if (p.gamma != nullptr) {
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                const float4 g4 = (n0 + 4 * j < p.N) ? ldg4(p.gamma + n0 + 4 * j) : make_float4(0.f, 0.f, 0.f, 0.f);
                                v[4 * j] *= g4.x; v[4 * j + 1] *= g4.y; v[4 * j + 2] *= g4.z; v[4 * j + 3] *= g4.w;
                            }
                        }
                        const uint32_t sbuf = stg + (epi_seq & 1) * EPI_STAGE_BYTES;
                        const uint32_t srow = sbuf + lane * 128;
                        if (lane == 0) bulk_wait_read1();
                        __syncwarp();
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            sts_v4(srow + ((j ^ sw) * 16), make_uint4(__float_as_uint(v[4 * j]), __float_as_uint(v[4 * j + 1]),
                                                                      __float_as_uint(v[4 * j + 2]), __float_as_uint(v[4 * j + 3])));
                        fence_proxy_async();
                        __syncwarp();
                        if (lane == 0) {
                            tma_reduce_add_2d(&tmC, sbuf, n0, m_base);
                            bulk_commit();
                        }
                        ++epi_seq;
                    }
                };
                uint32_t ra[32], rb[32];
                if (nch > 0) {
                    tmem_ld_32x32b_x32(t_row + chalf * COLS_PER_WARP, ra);
#pragma unroll 1
                    for (int c = 0; c < nch; c += 2) {
                        chunk(c, ra, rb);
                        if (c + 1 < nch) chunk(c + 1, rb, ra);
                    }
                } else {   // tile column range entirely outside the matrix: only hand the stage back
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) {
                        if constexpr (CTAS == 2) mbar_arrive_cluster(tempty_leader + acc * 8);
                        else mbar_arrive(&tempty_bar[acc]);
                    }
                }
            } else {
#pragma unroll 1
            for (int c = 0; c < COLS_PER_WARP / 32; ++c) {
                const int col = chalf * COLS_PER_WARP + c * 32;
                const int n0 = n_blk * BN + col;
                if (n0 >= p.N) break;  // warp-uniform
                uint32_t r[32];
                tmem_ld_32x32b_x32(t_row + col, r);
                tmem_ld_wait();
                // transpose through smem: lane = row writes 8 x 16 B, segment index XOR-swizzled by row
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int pj = j ^ (lane & 7);
                    sts_v4(stg + lane * 128 + pj * 16, make_uint4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]));
                }
                __syncwarp();
                const int n = n0 + seg * 4;
                float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f), gamma4 = make_float4(1.f, 1.f, 1.f, 1.f);
                if (n + 4 <= p.N) {
                    if (p.bias != nullptr) bias4 = ldg4(p.bias + n);
                    if (p.gamma != nullptr) gamma4 = ldg4(p.gamma + n);
                } else if (n < p.N) {
                    float b[4] = {0.f, 0.f, 0.f, 0.f}, g[4] = {1.f, 1.f, 1.f, 1.f};
                    for (int j = 0; j < 4; ++j)
                        if (n + j < p.N) {
                            if (p.bias != nullptr) b[j] = __ldg(p.bias + n + j);
                            if (p.gamma != nullptr) g[j] = __ldg(p.gamma + n + j);
                        }
                    bias4 = make_float4(b[0], b[1], b[2], b[3]);
                    gamma4 = make_float4(g[0], g[1], g[2], g[3]);
                }
                if (n + 4 <= p.N) {
                    // fast path: this lane owns 4 full columns of 8 rows.  All loads of a phase are
                    // issued before the first dependent use, mode dispatch happens once per chunk.
                    float v[32];
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const int rl = rsub + 4 * k;  // row inside the 32 x 32 block
                        const float4 t4 = lds_v4f(stg + rl * 128 + ((seg ^ (rl & 7)) * 16));
                        v[4 * k] = t4.x + bias4.x; v[4 * k + 1] = t4.y + bias4.y;
                        v[4 * k + 2] = t4.z + bias4.z; v[4 * k + 3] = t4.w + bias4.w;
                    }
                    const int rows_valid = p.M - (m_base + rsub);   // row k valid iff 4k < rows_valid
                    if (p.table != nullptr) {
                        float4 tt[8];
#pragma unroll
                        for (int k = 0; k < 8; ++k) {
                            const int m = m_base + rsub + 4 * k;
                            tt[k] = (4 * k < rows_valid)
                                        ? ldg4(p.table + static_cast<long long>(p.gin > 0 ? (m % p.gin) : m) * p.ldt + n)
                                        : make_float4(0.f, 0.f, 0.f, 0.f);
                        }
#pragma unroll
                        for (int k = 0; k < 8; ++k) {
                            v[4 * k] += tt[k].x; v[4 * k + 1] += tt[k].y; v[4 * k + 2] += tt[k].z; v[4 * k + 3] += tt[k].w;
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
                    if (p.store == ST_RESID32) {
                        float* o = reinterpret_cast<float*>(p.out) + n;
                        float4 x[8];
#pragma unroll
                        for (int k = 0; k < 8; ++k)
                            if (4 * k < rows_valid) x[k] = *reinterpret_cast<const float4*>(o + orow[k] * p.ldo);
#pragma unroll
                        for (int k = 0; k < 8; ++k)
                            if (4 * k < rows_valid) {
                                x[k].x = fmaf(gamma4.x, v[4 * k], x[k].x);
                                x[k].y = fmaf(gamma4.y, v[4 * k + 1], x[k].y);
                                x[k].z = fmaf(gamma4.z, v[4 * k + 2], x[k].z);
                                x[k].w = fmaf(gamma4.w, v[4 * k + 3], x[k].w);
                                *reinterpret_cast<float4*>(o + orow[k] * p.ldo) = x[k];
                            }
                    } else if (p.store == ST_16) {
                        uint16_t* o = reinterpret_cast<uint16_t*>(p.out) + n;
#pragma unroll
                        for (int k = 0; k < 8; ++k)
                            if (4 * k < rows_valid) {
                                uint2 w;
                                w.x = pack_16(v[4 * k], v[4 * k + 1], bf);
                                w.y = pack_16(v[4 * k + 2], v[4 * k + 3], bf);
                                *reinterpret_cast<uint2*>(o + orow[k] * p.ldo) = w;
                                if (p.out_lo != nullptr) {  // low half of a split-precision operand (fp16)
                                    const __half2* h = reinterpret_cast<const __half2*>(&w);
                                    uint2 lo;
                                    lo.x = pack_f16(v[4 * k] - __low2float(h[0]), v[4 * k + 1] - __high2float(h[0]));
                                    lo.y = pack_f16(v[4 * k + 2] - __low2float(h[1]), v[4 * k + 3] - __high2float(h[1]));
                                    *reinterpret_cast<uint2*>(reinterpret_cast<uint16_t*>(p.out_lo) + n + orow[k] * p.ldo) = lo;
                                }
                            }
                    } else if (p.store == ST_32) {
                        float* o = reinterpret_cast<float*>(p.out) + n;
#pragma unroll
                        for (int k = 0; k < 8; ++k)
                            if (4 * k < rows_valid)
                                *reinterpret_cast<float4*>(o + orow[k] * p.ldo) = make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
                    } else {  // ST_SWIGLU16 / ST_GATED16: column pairs (x1, x2) -> one output
                        uint16_t* o = reinterpret_cast<uint16_t*>(p.out) + (n >> 1);
                        const bool swiglu = p.store == ST_SWIGLU16;
#pragma unroll
                        for (int k = 0; k < 8; ++k)
                            if (4 * k < rows_valid) {
                                const float g0 = swiglu ? silu(v[4 * k]) * v[4 * k + 1] : tanhf(v[4 * k]) * sigmoidf_(v[4 * k + 1]);
                                const float g1 = swiglu ? silu(v[4 * k + 2]) * v[4 * k + 3] : tanhf(v[4 * k + 2]) * sigmoidf_(v[4 * k + 3]);
                                const uint32_t w = pack_16(g0, g1, bf);
                                *reinterpret_cast<uint32_t*>(o + orow[k] * p.ldo) = w;
                                if (p.out_lo != nullptr) {
                                    const __half2 h = *reinterpret_cast<const __half2*>(&w);
                                    *reinterpret_cast<uint32_t*>(reinterpret_cast<uint16_t*>(p.out_lo) + (n >> 1) + orow[k] * p.ldo) =
                                        pack_f16(g0 - __low2float(h), g1 - __high2float(h));
                                }
                            }
                    }
                } else if (n < p.N) {
                    // ragged right edge (N % 4 != 0): generic scalar path, one row at a time
                    // (not unrolled, and the row remap is recomputed: a dynamic index into orow[]
                    //  would push that array into local memory for every path)
#pragma unroll 1
                    for (int k = 0; k < 8; ++k) {
                        const int rl = rsub + 4 * k;
                        const float4 t4 = lds_v4f(stg + rl * 128 + ((seg ^ (rl & 7)) * 16));
                        const int m = m_base + rl;
                        const long long orow_k =
                            (p.gin > 0) ? static_cast<long long>(m / p.gin) * p.gout + p.goff + (m % p.gin) : m;
                        if (m < p.M) epilogue4(p, t4, m, orow_k, n, bias4, gamma4);
                    }
                }
                __syncwarp();
            }
            }  // EPI == 0
            if constexpr (EPI == 0) {
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    if constexpr (CTAS == 2) mbar_arrive_cluster(tempty_leader + acc * 8);
                    else mbar_arrive(&tempty_bar[acc]);
                }
            }
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
        if constexpr (EPI != 0) {
            if (lane == 0) bulk_wait0();   // every store / reduce of this warp has completed before the CTA exits
        }
    }

    tc_fence_before();
    if constexpr (CTAS == 2) cluster_sync_all(); else __syncthreads();
    if (warp == WARP_MMA) {
        tc_fence_after();
        if constexpr (CTAS == 2) tmem_dealloc2(tmem_base, C::TMEM_COLS); else tmem_dealloc(tmem_base, C::TMEM_COLS);
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
    const int esz = (kind == 2) ? 4 : 2;
    cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
    cuuint64_t gstr[1] = {static_cast<cuuint64_t>(ld) * esz};
    cuuint32_t box[2] = {static_cast<cuuint32_t>(BK_BYTES / esz), static_cast<cuuint32_t>(box_rows)};
    cuuint32_t estr[2] = {1, 1};
    const CUtensorMapDataType dt = (kind == 2)   ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                                   : (kind == 1) ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
                                                 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
    CUresult r = fn(tm, dt, 2, const_cast<void*>(ptr), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? SB_OK : SB_ERR_DRIVER;
}

}  // namespace

// fp16 [batches][rows][inner] view with zero fill outside `rows`: per-(bag, head) Q/K/V tiles of the
// tcgen05 attention (attention_tc.cu); 128-byte swizzle, box = {box_inner, box_rows, 1}
int make_tmap_3d_f16(CUtensorMap* tm, const void* ptr, int inner, int rows, int batches,
                     long long row_stride, long long batch_stride, int box_inner, int box_rows) {
    EncodeTiledFn fn = get_encode_fn();
    if (fn == nullptr) return SB_ERR_DRIVER;
    cuuint64_t gdim[3] = {static_cast<cuuint64_t>(inner), static_cast<cuuint64_t>(rows), static_cast<cuuint64_t>(batches)};
    cuuint64_t gstr[2] = {static_cast<cuuint64_t>(row_stride) * 2, static_cast<cuuint64_t>(batch_stride) * 2};
    cuuint32_t box[3] = {static_cast<cuuint32_t>(box_inner), static_cast<cuuint32_t>(box_rows), 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(ptr), gdim, gstr, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? SB_OK : SB_ERR_DRIVER;
}

namespace {

int g_num_sms = 0;
int g_force_mode = 0;  // bits 0-1: 0 auto, 1 never use the 2-CTA kernel, 2 always (when legal); bit 2: legacy epilogue

template <int BN, bool TF32, int CTAS, int EPI>
int launch(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, const GemmParams& p, int tiles,
           cudaStream_t stream) {
    using C = Cfg<BN, CTAS, EPI>;
    static bool configured = false;
    if (!configured) {
        if (cudaFuncSetAttribute(gemm_tn_kernel<BN, TF32, CTAS, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 C::SMEM_BYTES) != cudaSuccess)
            return SB_ERR_CUDA;
        configured = true;
    }
    const int units = g_num_sms / CTAS;  // CTAs (or CTA pairs) that fit the chip
    const int grid = (tiles < units ? tiles : units) * CTAS;
    ProfScope prof(PROF_GEMM, 2.0 * p.M * static_cast<double>(p.N) * p.K, stream);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(NUM_THREADS);
    cfg.dynamicSmemBytes = C::SMEM_BYTES;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CTAS;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    const cudaError_t e = cudaLaunchKernelEx(&cfg, gemm_tn_kernel<BN, TF32, CTAS, EPI>, tmA, tmB, tmC, p);
    count_launch();
    if (e != cudaSuccess) {
        fprintf(stderr, "stamp_b200: gemm_tn_kernel<%d,%d,%d,%d> launch failed: %s (grid %d, smem %d)\n", BN,
                static_cast<int>(TF32), CTAS, EPI, cudaGetErrorString(e), grid, C::SMEM_BYTES);
        cudaGetLastError();
    }
    return e == cudaSuccess ? SB_OK : SB_ERR_CUDA;
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

void gemm_force_mode(int mode) { g_force_mode = mode; }

int gemm_tn(const void* A, long long lda, const void* B, long long ldb, const GemmParams& p,
            cudaStream_t stream) {
    if (p.M <= 0 || p.N <= 0 || p.K <= 0 || A == nullptr || B == nullptr || p.out == nullptr)
        return SB_ERR_BAD_ARG;
    const int ealign = p.tf32 ? 4 : 8;  // 16-byte row pitch
    if ((lda % ealign) != 0 || (ldb % ealign) != 0 || (reinterpret_cast<uintptr_t>(A) & 15) != 0 ||
        (reinterpret_cast<uintptr_t>(B) & 15) != 0)
        return SB_ERR_BAD_ARG;
    if ((p.store == ST_SWIGLU16 || p.store == ST_GATED16) && (p.N % 2) != 0) return SB_ERR_BAD_ARG;
    // the epilogue moves 4 columns per lane with vector accesses: 16-byte aligned bases, pitches % 4
    auto misaligned = [](const void* q) { return q != nullptr && (reinterpret_cast<uintptr_t>(q) & 15) != 0; };
    if ((p.ldo % 4) != 0 || misaligned(p.out) || misaligned(p.out_lo) || misaligned(p.bias) ||
        misaligned(p.gamma) || misaligned(p.table) || (p.table != nullptr && (p.ldt % 4) != 0))
        return SB_ERR_BAD_ARG;
    gemm_num_sms();

    const int num_m = (p.M + BM - 1) / BM;
    const int tiles256 = num_m * ((p.N + 255) / 256);
    const int tiles128 = num_m * ((p.N + 127) / 128);
    // CTA-pair tiles (256 x 256) once there are at least two waves of them; else single-CTA tiles,
    // 128 x 256 unless that leaves most SMs without a tile
    const int tiles_pair = ((p.M + 255) / 256) * ((p.N + 255) / 256);
    bool use_pair = !p.tf32 && p.N >= 256 && tiles_pair >= 2 * (g_num_sms / 2);
    if ((g_force_mode & 3) == 1) use_pair = false;
    if ((g_force_mode & 3) == 2 && !p.tf32) use_pair = true;
    const bool use256 = (p.N >= 256) && (tiles256 >= g_num_sms);
    const int bn = (use_pair || use256) ? 256 : 128;

    const int kind = p.tf32 ? 2 : (p.bf16 ? 1 : 0);
    CUtensorMap tmA, tmB, tmC;
    if (p.a_kwrap < 0 || (p.a_kwrap % 64) != 0) return SB_ERR_BAD_ARG;
    int rc = make_tmap(&tmA, A, p.M, p.a_kwrap > 0 ? p.a_kwrap : p.K, lda, BM, kind);
    if (rc != SB_OK) return rc;
    rc = make_tmap(&tmB, B, p.N, p.K, ldb, use_pair ? 128 : bn, kind);
    if (rc != SB_OK) return rc;

    // epilogue variant: TMA store (16-bit outputs) / TMA reduce-add (fp32 residual) when the output is a plain
    // matrix; the generic register path otherwise
    int epi = 0;
    const bool plain = p.gin == 0 && p.table == nullptr && p.out_lo == nullptr && (g_force_mode & 4) == 0;
    if (plain && !p.tf32 && p.store == ST_16 && (p.ldo % 8) == 0) epi = 1;
    if (plain && !p.tf32 && p.store == ST_SWIGLU16 && bn == 256 && (p.ldo % 8) == 0 && p.act == ACT_NONE) epi = 1;
    if (plain && p.store == ST_RESID32 && p.act == ACT_NONE) epi = 2;
    tmC = tmA;
    if (epi == 1)
        rc = make_tmap(&tmC, p.out, p.M, p.store == ST_SWIGLU16 ? p.N / 2 : p.N, p.ldo, 32, p.bf16 ? 1 : 0);
    else if (epi == 2)
        rc = make_tmap(&tmC, p.out, p.M, p.N, p.ldo, 32, 2);
    if (rc != SB_OK) return rc;

#define SB_GEMM_LAUNCH(BN_, TF_, CT_, TILES_)                                                             \
    (epi == 1 ? launch<BN_, TF_, CT_, (TF_) ? 0 : 1>(tmA, tmB, tmC, p, TILES_, stream)                     \
     : epi == 2 ? launch<BN_, TF_, CT_, 2>(tmA, tmB, tmC, p, TILES_, stream)                               \
                : launch<BN_, TF_, CT_, 0>(tmA, tmB, tmC, p, TILES_, stream))
    if (use_pair) return SB_GEMM_LAUNCH(256, false, 2, tiles_pair);
    if (p.tf32) return use256 ? SB_GEMM_LAUNCH(256, true, 1, tiles256) : SB_GEMM_LAUNCH(128, true, 1, tiles128);
    return use256 ? SB_GEMM_LAUNCH(256, false, 1, tiles256) : SB_GEMM_LAUNCH(128, false, 1, tiles128);
#undef SB_GEMM_LAUNCH
}

}  // namespace sb
