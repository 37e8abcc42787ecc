// tcgen05 attention for long feature bags, third generation: the ALiBi distance tiles come PRE-COMPUTED from
// HBM / L2 by TMA, the softmax weights go back to tensor memory and feed the second product from there.
//
//   plain : O = softmax(Q K^T * scale) V
//   ALiBi : O = softmax(Q K^T * scale) V  -  c_h * Dist V        (bias subtracted AFTER the softmax:
//           src/stamp/modeling/models/vision_tranformer.py:58-72)
//
// What bounded the previous generations (attention_mil_tc.cu; ncu in profiles/r1_ncu_full_mil_deploy_attention_4096.csv)
// was the CUDA-core side: ~20 thread instructions and two MUFU ops (ex2, sqrt) per query-key pair, every head
// recomputing the same distances, plus 32 KB of shared-memory stores and reloads per tile for P and D.  Here:
//   * Dist is a property of the bag, not of the head or the layer: `dist16_kernel` writes it once per forward as a
//     16-bit [S, S] matrix (33.5 MB for 4096 tiles: it lives in the 126 MB L2 while the eight heads of a query tile
//     stream it), scaled per bag by a power of two into the fp16 range; the per-head slope c_h multiplies the fp32
//     accumulator in the epilogue instead of every element.  The D tile lands in shared memory by TMA already in
//     the K-major 128B-swizzled UMMA layout: no sqrt, no coordinate loads, no packing, no smem stores by threads.
//   * P never touches shared memory: the softmax threads write it with tcgen05.st over the S tile they have just
//     read, and P V is a tcgen05.mma with the A operand in tensor memory (TS form).
//   * One thread per query row (= TMEM lane) holds its 64 scores of a tile in registers: no partial row maxima, no
//     exchange through shared memory, no second accumulator; the accumulator is rescaled lazily (only when a row
//     maximum grows by more than 2^8, warp-uniform because tcgen05.ld/st are collective).
// Left per pair: max, fma, ex2, add, half a pack -- one MUFU op.
//
// One CTA = 128 query rows of one (bag, head), 6 warps: warp 0 TMA (K, V, D tiles of 64 keys, 2 stages), warp 1
// tcgen05.mma issue, warps 2-5 softmax.  TMEM: S0 | S1 | O1 | O2 = 4 x 64 = 256 columns, 81 KB of shared memory
// -> two CTAs per SM.  The tensor pipe executes in issue order, which is all the protection the aliased S/P columns
// need: S(t+2) is issued after P(t) V, which is issued after the softmax threads have handed P(t) over.
//
// Training variant (TRAIN): bf16 operands, Dhat = Dist * inv_rm_h, O = O1 / l - beta_h * Dhat V; also stores
// Osm = O1 / l, Dhat V (fp32) and the row log-sum-exp (log2 domain) for the backward (attention_train*.cu).
#include <math.h>

#include "attention.cuh"
#include "attention_train.cuh"
#include "common.cuh"
#include "gemm.cuh"

namespace sb {
namespace {

constexpr int V3_THREADS = 192;
constexpr int QT_BYTES = 128 * 128;    // 128 rows x 64 halfs
constexpr int KV_BYTES = 64 * 128;     // 64 keys x 64 halfs
// Compile-time switch for fetching the next tile's scores during this tile's exponentials (see `tile` below).
// Measured on B200 at the deploy shape: OFF 0.356 ms / bag, ON 0.401 ms / bag (168 registers with spills, and S(t+1)
// queues behind the previous tile's products anyway) -> off.
#ifndef SB_V3_PREFETCH
#define SB_V3_PREFETCH 0
#endif
constexpr bool V3_PREFETCH = SB_V3_PREFETCH != 0;
constexpr int K_STAGES = 4;                            // K tiles run ahead: S(t+1) must never wait for a load
constexpr int VD_BYTES = KV_BYTES + QT_BYTES;          // V | D, two stages: freed when the tile's products retire

struct V3Smem {
    static constexpr int off_q = 0;
    static constexpr int off_k = QT_BYTES;                        // K_STAGES K tiles
    static constexpr int off_vd = off_k + K_STAGES * KV_BYTES;    // 2 stages of (V, D)
    static constexpr int off_bar = off_vd + 2 * VD_BYTES;
    static constexpr int total = off_bar + 256 + 1024;
};

struct V3Out {             // training-only outputs (null for inference)
    uint16_t* out16;
    float* osm;
    float* odv;
    float* lse2;
    const float* beta;     // [H] bias_scale_h
    const float* inv_rm;   // [H] 1 / running_mean_h
};

template <bool TRAIN>
__device__ __forceinline__ uint32_t pack_op(float a, float b) {
    if constexpr (TRAIN) return pack_bf16(a, b);
    else return pack_f16(a, b);
}

__device__ __forceinline__ float sqrt_apx(float x) {
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;\n" : "=f"(r) : "f"(x));
    return r;
}

constexpr int DIST_PARTS = 32;   // partial bounding boxes per bag (one per block of the first pass)

// pass 1, grid (DIST_PARTS, B): partial bounding boxes of the token coordinates, bbox[b][part] = {xmin, xmax, ymin, ymax}
__global__ void __launch_bounds__(256)
dist_bbox_kernel(const float2* __restrict__ coords, int S, float4* __restrict__ bbox, const int* __restrict__ seq_off) {
    const int b = blockIdx.y;
    const float2* c = coords + static_cast<long long>(b) * S;
    if (seq_off != nullptr) {   // ragged batch: bag b = rows seq_off[b] .. seq_off[b+1]
        c = coords + seq_off[b];
        S = seq_off[b + 1] - seq_off[b];
    }
    float xmin = INFINITY, xmax = -INFINITY, ymin = INFINITY, ymax = -INFINITY;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < S; i += gridDim.x * blockDim.x) {
        const float2 v = __ldg(c + i);
        xmin = fminf(xmin, v.x); xmax = fmaxf(xmax, v.x);
        ymin = fminf(ymin, v.y); ymax = fmaxf(ymax, v.y);
    }
    __shared__ float red[4][8];
    xmin = -warp_max(-xmin); xmax = warp_max(xmax);
    ymin = -warp_max(-ymin); ymax = warp_max(ymax);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) { red[0][w] = xmin; red[1][w] = xmax; red[2][w] = ymin; red[3][w] = ymax; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 1; i < 8; ++i) {
            xmin = fminf(xmin, red[0][i]); xmax = fmaxf(xmax, red[1][i]);
            ymin = fminf(ymin, red[2][i]); ymax = fmaxf(ymax, red[3][i]);
        }
        bbox[b * gridDim.x + blockIdx.x] = make_float4(xmin, xmax, ymin, ymax);
    }
}

// power of two that puts the largest possible distance of the bag (the diagonal of its bounding box) into
// (8192, 16384]: returns e, the matrix holds dist * 2^-e
__device__ __forceinline__ int dist_exponent(const float4* __restrict__ bbox_b) {
    float4 bb = __ldg(bbox_b + (threadIdx.x & 31) % DIST_PARTS);
    static_assert(DIST_PARTS == 32, "one partial box per lane");
    const float xmin = -warp_max(-bb.x), xmax = warp_max(bb.y), ymin = -warp_max(-bb.z), ymax = warp_max(bb.w);
    const float diag = sqrtf((xmax - xmin) * (xmax - xmin) + (ymax - ymin) * (ymax - ymin));
    int e = 0;
    if (diag > 0.f && isfinite(diag)) e = static_cast<int>(ceilf(log2f(diag / 16384.f)));
    return max(-100, min(100, e));
}

// pass 2: Dist[b, q, k] = |x_q - x_k| * 2^-e_b as 16-bit (fp16 for inference, bf16 for training); row pitch `ld`.
// One block = 64 query rows x 256 keys: a thread keeps 8 key coordinates in registers and walks 8 rows (one 16-byte
// store per row; a warp writes 512 contiguous bytes).  Block (0, 0) of every bag also publishes scale[b] = {2^-e, 2^e}.
__global__ void __launch_bounds__(256)
dist16_kernel(const float2* __restrict__ coords, const float4* __restrict__ bbox, float* __restrict__ scale,
              uint16_t* __restrict__ out, int S, long long ld, long long batch_stride, int bf16,
              const int* __restrict__ seq_off) {
    const int b = blockIdx.z;
    long long row0 = static_cast<long long>(b) * S;
    if (seq_off != nullptr) {   // ragged batch: common row pitch, bag b starts at row seq_off[b]
        row0 = seq_off[b];
        S = seq_off[b + 1] - seq_off[b];
        batch_stride = 0;
        out += row0 * ld;
        if (blockIdx.y * 64 >= S) return;
    }
    const int e = dist_exponent(bbox + b * DIST_PARTS);     // every warp computes the same value
    const float g = exp2f(static_cast<float>(-e));
    if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) {
        scale[2 * b] = g;
        scale[2 * b + 1] = exp2f(static_cast<float>(e));
    }
    const int k0 = (blockIdx.x * 32 + (threadIdx.x & 31)) * 8;
    if (k0 >= ld) return;
    const float2* c = coords + row0;
    float2 ck[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) ck[j] = (k0 + j < S) ? __ldg(c + k0 + j) : make_float2(0.f, 0.f);
    const int q0 = blockIdx.y * 64 + (threadIdx.x >> 5) * 8;
    uint16_t* o = out + b * batch_stride + k0;
#pragma unroll 2
    for (int r = 0; r < 8; ++r) {
        const int q = q0 + r;
        if (q >= S) break;
        const float2 cq = __ldg(c + q);
        float d[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float dx = cq.x - ck[j].x, dy = cq.y - ck[j].y;
            d[j] = (k0 + j < S) ? sqrt_apx(fmaf(dx, dx, dy * dy)) * g : 0.f;
        }
        uint4 w;
        w.x = pack_16(d[0], d[1], bf16 != 0); w.y = pack_16(d[2], d[3], bf16 != 0);
        w.z = pack_16(d[4], d[5], bf16 != 0); w.w = pack_16(d[6], d[7], bf16 != 0);
        *reinterpret_cast<uint4*>(o + static_cast<long long>(q) * ld) = w;
    }
}

template <bool ALIBI, bool TRAIN>
__global__ void __launch_bounds__(V3_THREADS, 2)
mil_attn_v3_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                   const __grid_constant__ CUtensorMap tm_v, const __grid_constant__ CUtensorMap tm_d,
                   const AttnParams p, int k_col0, const V3Out t, float rescale_margin) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>(
        (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    uint8_t* sQ = smem + V3Smem::off_q;
    uint8_t* sK = smem + V3Smem::off_k;
    uint8_t* sVD = smem + V3Smem::off_vd;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + V3Smem::off_bar);
    uint64_t* kfull = bars + 12;   // [K_STAGES] TMA -> MMA (K tile)
    uint64_t* kempty = kfull + K_STAGES;   // [K_STAGES] MMA -> TMA (S of the tile retired)
    uint64_t* full = bars;         // [2] TMA -> MMA      (V, D of a key tile)
    uint64_t* empty = bars + 2;    // [2] MMA -> TMA      (the tile's products have retired)
    uint64_t* sfull = bars + 4;    // [2] MMA -> softmax  (S tile in TMEM)
    uint64_t* pfull = bars + 6;    // [2] softmax -> MMA  (P tile in TMEM, over the S tile); one barrier per tile parity:
                                   //     a warp running ahead (S(t+1) is ready early) must not complete tile t's phase
                                   //     with its arrival for tile t+1
    uint64_t* pvdone = bars + 8;   //     MMA -> softmax  (P V of the tile retired: O1 may be rescaled)
    uint64_t* ofull = bars + 9;
    uint64_t* qfull = bars + 10;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 11);   // (bars 12 .. 12 + 2 K_STAGES: the K ring)

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // the heads of one (bag, query tile) are adjacent CTAs: they stream the same D tiles through L2 together
    const int b = blockIdx.x / p.H, h = blockIdx.x % p.H;
    const int q0 = blockIdx.y * 128;
    int S = p.S;
    int row0 = 0, bc = b;           // first row of the bag inside the tensor maps, batch coordinate
    if (!TRAIN && p.seq_off != nullptr) {
        // ragged batch: the bags are concatenated along the row axis of 2-D views.  Rows past a bag's end belong to
        // the next bag (not zero-filled like in the per-bag maps): their scores are masked by `nvalid`, their V rows
        // meet P = 0 and zero distance columns, query rows past the end are computed and never stored.
        row0 = __ldg(p.seq_off + b);
        S = __ldg(p.seq_off + b + 1) - row0;
        bc = 0;
        if (q0 >= S) return;        // uniform for the CTA, before any barrier / tensor-memory allocation
    }
    const int nkt = (S + 63) / 64;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tm_q);
        tma_prefetch_desc(&tm_k);
        tma_prefetch_desc(&tm_v);
        if constexpr (ALIBI) tma_prefetch_desc(&tm_d);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], 1);
            mbar_init(&sfull[i], 1);
            mbar_init(&pfull[i], 4);
        }
        for (int i = 0; i < K_STAGES; ++i) {
            mbar_init(&kfull[i], 1);
            mbar_init(&kempty[i], 1);
        }
        mbar_init(pvdone, 1);
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
    constexpr uint32_t COL_O1 = 128, COL_O2 = 192;   // S0 at column 0, S1 at 64

    if (warp == 0) {
        // ------------------------------------ TMA producer ------------------------------------
        if (lane == 0) {
            mbar_expect_tx(qfull, QT_BYTES);
            tma_load_3d(sQ, &tm_q, qfull, h * 64, row0 + q0, bc);
            // two rings served by one thread, whichever has a free slot: K tiles run up to K_STAGES ahead of the
            // products (the next S must never wait for a load), V / D tiles recycle as their products retire
            int ki = 0, vi = 0;
            while (ki < nkt || vi < nkt) {
                if (ki < nkt && mbar_try_wait(&kempty[ki % K_STAGES], ((ki / K_STAGES) & 1) ^ 1)) {
                    mbar_expect_tx(&kfull[ki % K_STAGES], KV_BYTES);
                    tma_load_3d(sK + (ki % K_STAGES) * KV_BYTES, &tm_k, &kfull[ki % K_STAGES], k_col0 + h * 64, row0 + ki * 64, bc);
                    ++ki;
                    continue;
                }
                if (vi < nkt && mbar_try_wait(&empty[vi & 1], ((vi >> 1) & 1) ^ 1)) {
                    uint8_t* st = sVD + (vi & 1) * VD_BYTES;
                    mbar_expect_tx(&full[vi & 1], ALIBI ? VD_BYTES : KV_BYTES);
                    tma_load_3d(st, &tm_v, &full[vi & 1], h * 64, row0 + vi * 64, bc);
                    if constexpr (ALIBI) tma_load_3d(st + KV_BYTES, &tm_d, &full[vi & 1], vi * 64, row0 + q0, bc);
                    ++vi;
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------ MMA issuer --------------------------------------
        if (lane == 0) {
            const uint32_t idesc_s = umma_idesc_f16(128, 64, TRAIN, false, false);
            const uint32_t idesc_o = umma_idesc_f16(128, 64, TRAIN, false, true);   // V: MN-major B operand
            const uint64_t q_desc = umma_desc_k128(smem_u32(sQ));
            mbar_wait(qfull, 0);
            auto issue_s = [&](int kt) {
                const int s = kt & 1, ks = kt % K_STAGES;
                mbar_wait(&kfull[ks], (kt / K_STAGES) & 1);
                tc_fence_after();
                const uint64_t k_desc = umma_desc_k128(smem_u32(sK + ks * KV_BYTES));
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_f16_ss(tmem + s * 64, q_desc + 2 * k, k_desc + 2 * k, idesc_s, k != 0);
                umma_commit(&sfull[s]);
                umma_commit(&kempty[ks]);
            };
            issue_s(0);
            for (int kt = 0; kt < nkt; ++kt) {
                if (kt + 1 < nkt) issue_s(kt + 1);
                const int s = kt & 1;
                const uint8_t* st = sVD + s * VD_BYTES;
                mbar_wait(&full[s], (kt >> 1) & 1);
                mbar_wait(&pfull[kt & 1], (kt >> 1) & 1);
                tc_fence_after();
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    // 16 keys per step: 16 rows x 128 B of the V tile read as an MN-major operand; P: 8 TMEM columns
                    const uint64_t v_desc = umma_desc_mn128(smem_u32(st + k * 2048), 0);
                    umma_f16_ts(tmem + COL_O1, tmem + s * 64 + k * 8, v_desc, idesc_o, (kt | k) != 0);
                    if constexpr (ALIBI) {
                        const uint64_t d_desc = umma_desc_k128(smem_u32(st + KV_BYTES)) + 2 * k;
                        umma_f16_ss(tmem + COL_O2, d_desc, v_desc, idesc_o, (kt | k) != 0);
                    }
                }
                umma_commit(&empty[s]);
                umma_commit(pvdone);
            }
            umma_commit(ofull);
        }
    } else {
        // ---------------- softmax: one thread per query row (= TMEM lane), 64 keys per tile ----------------
        const int quarter = warp & 3;
        const int r = quarter * 32 + lane;
        const int row = q0 + r;
        const uint32_t t_lane = tmem + (static_cast<uint32_t>(quarter * 32) << 16);
        const float sl2 = p.scale_log2;
        float ms = -INFINITY;      // reference maximum of the row, already multiplied by scale * log2(e)
        float l4[4] = {0.f, 0.f, 0.f, 0.f};
        // Software pipeline over the key tiles: the scores of tile kt+1 are fetched from tensor memory (tcgen05.ld is
        // asynchronous until tcgen05.wait::ld) WHILE tile kt is exponentiated.  A warp's 64-column load occupies its
        // sub-partition's TMEM port for ~512 cycles and the 64 ex2 per thread its MUFU for another ~512; issued back
        // to back they were serialised (ncu: 40 % of the warp samples in the TMEM-load scoreboard).
        uint32_t va[64], vb[64];     // the two tiles in flight swap roles (no copies): loop unrolled by two
        // one key tile: `cur` holds its scores (already in registers), `nxt` receives the next tile's
        auto tile = [&](int kt, uint32_t (&cur)[64], uint32_t (&nxt)[64], bool have) {
            const int s = kt & 1;
            if (!have) {
                mbar_wait(&sfull[s], (kt >> 1) & 1);
                tc_fence_after();
                tmem_ld_32x32b_x64(t_lane + s * 64, cur);
                tmem_ld_wait();
            }
            const int nvalid = S - kt * 64;            // >= 64 except on the last tile
            float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
            if (nvalid >= 64) {
#pragma unroll
                for (int j = 0; j < 64; ++j) mx4[j & 3] = fmaxf(mx4[j & 3], __uint_as_float(cur[j]));
            } else {
#pragma unroll
                for (int j = 0; j < 64; ++j)
                    if (j < nvalid) mx4[j & 3] = fmaxf(mx4[j & 3], __uint_as_float(cur[j]));
            }
            const float cand = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3])) * sl2;
            {
                // raise the reference maximum lazily (only past the margin) and rescale l and the O1 row.  tcgen05.ld /
                // .st are warp-collective: the branch is taken by the whole warp as soon as one row needs it, rows that
                // do not rescale by 1.
                const bool need = cand > ms + rescale_margin;
                const bool resc = need && ms != -INFINITY;
                if (__any_sync(0xffffffffu, resc)) {
                    const float f = resc ? ex2_approx(ms - cand) : 1.0f;
                    l4[0] *= f; l4[1] *= f; l4[2] *= f; l4[3] *= f;
                    mbar_wait(pvdone, (kt & 1) ^ 1);    // P V of tile kt-1 has retired: O1 is stable
                    tc_fence_after();
#pragma unroll 1
                    for (int c = 0; c < 2; ++c) {
                        uint32_t o[32];
                        tmem_ld_32x32b_x32(t_lane + COL_O1 + c * 32, o);
                        tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 32; ++j) o[j] = __float_as_uint(__uint_as_float(o[j]) * f);
                        tmem_st_32x32b_x32(t_lane + COL_O1 + c * 32, o);
                    }
                    tmem_st_wait();
                }
                if (need) ms = cand;
            }
            // The next tile's scores are fetched DURING this tile's exponentials if they exist by then (S(kt+1) queues
            // behind P(kt-1) V and D V in the tensor pipe, so it is tested a quarter of the way in, warp-uniformly);
            // otherwise they are loaded at the top of the next tile as usual.
            bool got = false;
            auto exp_pair = [&](int j, bool masked) {
                if (V3_PREFETCH && j == 16 && kt + 1 < nkt) {
                    got = __shfl_sync(0xffffffffu, mbar_test(&sfull[s ^ 1], ((kt + 1) >> 1) & 1) ? 1 : 0, 0) != 0;
                    if (got) {
                        tc_fence_after();
                        tmem_ld_32x32b_x32_at<0>(t_lane + (s ^ 1) * 64, nxt);
                    }
                }
                if (j == 40 && got) tmem_ld_32x32b_x32_at<32>(t_lane + (s ^ 1) * 64 + 32, nxt);
                const float p0 = (!masked || j < nvalid) ? ex2_approx(fmaf(__uint_as_float(cur[j]), sl2, -ms)) : 0.f;
                const float p1 = (!masked || j + 1 < nvalid) ? ex2_approx(fmaf(__uint_as_float(cur[j + 1]), sl2, -ms)) : 0.f;
                l4[(j >> 1) & 3] += p0 + p1;
                cur[j >> 1] = pack_op<TRAIN>(p0, p1);   // in place: pair j lands in cur[j / 2], which has been read by then
            };
            if (nvalid >= 64) {
#pragma unroll
                for (int j = 0; j < 64; j += 2) exp_pair(j, false);
            } else {
#pragma unroll
                for (int j = 0; j < 64; j += 2) exp_pair(j, true);
            }
            if (got) tmem_ld_wait();     // the next tile's scores have landed
            // P over the first 32 columns of the S tile it came from (16-bit pairs, K-major A operand in TMEM)
            tmem_st_32x32b_x32_lo(t_lane + s * 64, cur);
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&pfull[kt & 1]);
            return got;
        };
        {
            bool have = false;
            for (int kt = 0; kt < nkt; kt += 2) {
                have = tile(kt, va, vb, have);
                if (kt + 1 < nkt) have = tile(kt + 1, vb, va, have);
            }
        }
        const float lt = (l4[0] + l4[1]) + (l4[2] + l4[3]);

        mbar_wait(ofull, 0);
        tc_fence_after();
        const float inv = 1.0f / lt;
        // O2 holds (Dist * 2^-e) V: the head's slope and the bag's 2^e come back in here, in fp32
        float coef = 0.f;
        if constexpr (ALIBI) {
            const float unscale = __ldg(p.dscale + 2 * b + 1);
            coef = (TRAIN ? __ldg(t.inv_rm + h) : __ldg(p.slope + h)) * unscale;
        }
        const long long obase = bc * p.out_batch_stride + static_cast<long long>(row0 + row) * p.out_row_stride + h * 64;
        if constexpr (TRAIN) {
            if (row < S) t.lse2[(static_cast<long long>(b) * p.H + h) * S + row] = ms + log2f(lt);
        }
#pragma unroll 1
        for (int c = 0; c < 2; ++c) {
            uint32_t o1[32], o2[32];
            tmem_ld_32x32b_x32(t_lane + COL_O1 + c * 32, o1);
            if constexpr (ALIBI) tmem_ld_32x32b_x32(t_lane + COL_O2 + c * 32, o2);
            tmem_ld_wait();
            if (row < S) {
                float sm[32], dv[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    sm[j] = __uint_as_float(o1[j]) * inv;
                    dv[j] = ALIBI ? __uint_as_float(o2[j]) * coef : 0.f;
                }
                const long long ob = obase + c * 32;
                if constexpr (TRAIN) {
                    const float beta = ALIBI ? __ldg(t.beta + h) : 0.f;
#pragma unroll
                    for (int j = 0; j < 32; j += 8) {
                        float y[8];
#pragma unroll
                        for (int e = 0; e < 8; ++e) y[e] = ALIBI ? fmaf(-beta, dv[j + e], sm[j + e]) : sm[j + e];
                        *reinterpret_cast<uint4*>(t.out16 + ob + j) =
                            make_uint4(pack_bf16(y[0], y[1]), pack_bf16(y[2], y[3]), pack_bf16(y[4], y[5]), pack_bf16(y[6], y[7]));
                        *reinterpret_cast<float4*>(t.osm + ob + j) = make_float4(sm[j], sm[j + 1], sm[j + 2], sm[j + 3]);
                        *reinterpret_cast<float4*>(t.osm + ob + j + 4) = make_float4(sm[j + 4], sm[j + 5], sm[j + 6], sm[j + 7]);
                        if constexpr (ALIBI) {
                            *reinterpret_cast<float4*>(t.odv + ob + j) = make_float4(dv[j], dv[j + 1], dv[j + 2], dv[j + 3]);
                            *reinterpret_cast<float4*>(t.odv + ob + j + 4) = make_float4(dv[j + 4], dv[j + 5], dv[j + 6], dv[j + 7]);
                        }
                    }
                } else {
                    float y[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) y[j] = sm[j] - dv[j];
                    if (p.out_f32) {
                        float* of = reinterpret_cast<float*>(p.out) + ob;
                        float* ol = (p.out_lo != nullptr) ? p.out_lo + ob : nullptr;
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            const float4 hi = make_float4(round_tf32(y[j]), round_tf32(y[j + 1]), round_tf32(y[j + 2]), round_tf32(y[j + 3]));
                            *reinterpret_cast<float4*>(of + j) = hi;
                            if (ol != nullptr)
                                *reinterpret_cast<float4*>(ol + j) = make_float4(round_tf32(y[j] - hi.x), round_tf32(y[j + 1] - hi.y),
                                                                                 round_tf32(y[j + 2] - hi.z), round_tf32(y[j + 3] - hi.w));
                        }
                    } else {
                        __half* oh = reinterpret_cast<__half*>(p.out) + ob;
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

int g_v3_enabled = 1;
int g_v3_eager = 0;   // tests: rescale the accumulator whenever a row maximum grows (exercises the TMEM read-modify-write)

template <bool ALIBI, bool TRAIN>
int launch_v3(const CUtensorMap& tm_q, const CUtensorMap& tm_k, const CUtensorMap& tm_v, const CUtensorMap& tm_d,
              const AttnParams& p, int k_col0, const V3Out& t, cudaStream_t stream) {
    static bool configured = false;
    if (!configured) {
        if (cudaFuncSetAttribute(mil_attn_v3_kernel<ALIBI, TRAIN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 V3Smem::total) != cudaSuccess)
            return SB_ERR_CUDA;
        configured = true;
    }
    const int s_grid = (p.seq_off != nullptr) ? p.S_max : p.S;
    dim3 grid(p.B * p.H, (s_grid + 127) / 128);
    ProfScope prof(PROF_ATTN, 4.0 * p.B * p.H * static_cast<double>(p.S) * p.S * 64, stream);
    mil_attn_v3_kernel<ALIBI, TRAIN><<<grid, V3_THREADS, V3Smem::total, stream>>>(tm_q, tm_k, tm_v, tm_d, p, k_col0, t,
                                                                                  g_v3_eager ? 0.f : 8.f);
    count_launch();
    return cudaGetLastError() == cudaSuccess ? SB_OK : SB_ERR_CUDA;
}

}  // namespace

void attention_mil_v3_enable(int on) {
    g_v3_enabled = on & 1;
    g_v3_eager = (on >> 1) & 1;
}

size_t mil_dist16_bytes(int B, int S) {
    const size_t ld = (static_cast<size_t>(S) + 63) / 64 * 64;
    return static_cast<size_t>(B) * S * ld * 2;
}

// scale [B][2] = {2^-e, 2^e};  dist16 [B, S, ld] with ld = S rounded up to 64; bbox_scratch: B * 32 * 16 bytes
int mil_dist16(const float* coords_s, int B, int S, int bf16, float* scale, uint16_t* dist16, void* bbox_scratch,
               cudaStream_t stream) {
    if (coords_s == nullptr || scale == nullptr || dist16 == nullptr || bbox_scratch == nullptr || B <= 0 || S <= 0)
        return SB_ERR_BAD_ARG;
    if (B > 65535 || S > 65535 * 64) return SB_ERR_UNSUPPORTED;   // grid limits
    const long long ld = (static_cast<long long>(S) + 63) / 64 * 64;
    ProfScope prof(PROF_ROWOP, static_cast<double>(B) * S * ld * 2.0, stream);
    dist_bbox_kernel<<<dim3(DIST_PARTS, B), 256, 0, stream>>>(reinterpret_cast<const float2*>(coords_s), S,
                                                              static_cast<float4*>(bbox_scratch), nullptr);
    count_launch();
    dim3 grid(static_cast<unsigned>((ld + 255) / 256), static_cast<unsigned>((S + 63) / 64), B);
    dist16_kernel<<<grid, 256, 0, stream>>>(reinterpret_cast<const float2*>(coords_s), static_cast<const float4*>(bbox_scratch),
                                            scale, dist16, S, ld, static_cast<long long>(S) * ld, bf16, nullptr);
    count_launch();
    return cudaGetLastError() == cudaSuccess ? SB_OK : SB_ERR_CUDA;
}

int mil_dist16_ragged(const float* coords_s, const int* seq_off, int B, int S_max, long long ld, float* scale,
                      uint16_t* dist16, void* bbox_scratch, cudaStream_t stream) {
    if (coords_s == nullptr || seq_off == nullptr || scale == nullptr || dist16 == nullptr || bbox_scratch == nullptr ||
        B <= 0 || S_max <= 0 || ld < S_max || (ld % 64) != 0)
        return SB_ERR_BAD_ARG;
    if (B > 65535 || S_max > 65535 * 64) return SB_ERR_UNSUPPORTED;
    ProfScope prof(PROF_ROWOP, static_cast<double>(B) * S_max * ld * 2.0, stream);
    dist_bbox_kernel<<<dim3(DIST_PARTS, B), 256, 0, stream>>>(reinterpret_cast<const float2*>(coords_s), 0,
                                                              static_cast<float4*>(bbox_scratch), seq_off);
    count_launch();
    dim3 grid(static_cast<unsigned>((ld + 255) / 256), static_cast<unsigned>((S_max + 63) / 64), B);
    dist16_kernel<<<grid, 256, 0, stream>>>(reinterpret_cast<const float2*>(coords_s), static_cast<const float4*>(bbox_scratch),
                                            scale, dist16, 0, ld, 0, 0, seq_off);
    count_launch();
    return cudaGetLastError() == cudaSuccess ? SB_OK : SB_ERR_CUDA;
}

size_t mil_dist16_scratch_bytes(int B) { return static_cast<size_t>(B) * DIST_PARTS * 16; }

namespace {

int make_maps(const void* q, const void* v, long long row_stride, long long batch_stride, long long v_rs, long long v_bs,
              const uint16_t* dist16, int B, int S, CUtensorMap* tm_q, CUtensorMap* tm_k, CUtensorMap* tm_v, CUtensorMap* tm_d,
              long long ragged_dist_ld = 0) {
    int rc = make_tmap_3d_f16(tm_q, q, static_cast<int>(row_stride), S, B, row_stride, batch_stride, 64, 128);
    if (rc != SB_OK) return rc;
    rc = make_tmap_3d_f16(tm_k, q, static_cast<int>(row_stride), S, B, row_stride, batch_stride, 64, 64);
    if (rc != SB_OK) return rc;
    rc = make_tmap_3d_f16(tm_v, v, static_cast<int>(v_rs), S, B, v_rs, v_bs, 64, 64);
    if (rc != SB_OK) return rc;
    *tm_d = *tm_q;
    if (dist16 != nullptr) {
        const long long ld = (static_cast<long long>(S) + 63) / 64 * 64;
        // inner extent S (not ld): key columns past the bag are zero-filled by the TMA unit
        if (ragged_dist_ld > 0)   // ragged: one tall matrix [total rows, ld], zeros stored past each bag's length
            rc = make_tmap_3d_f16(tm_d, dist16, static_cast<int>(ragged_dist_ld), S, 1, ragged_dist_ld,
                                  static_cast<long long>(S) * ragged_dist_ld, 64, 128);
        else
            rc = make_tmap_3d_f16(tm_d, dist16, S, S, B, ld, static_cast<long long>(S) * ld, 64, 128);
    }
    return rc;
}

}  // namespace

// SB_ERR_UNSUPPORTED: outside this kernel's envelope (masked calls, head_dim != 64, short sequences, ALiBi without
// a distance matrix) -> the caller uses the previous kernels
int attention_mil_v3_fwd(const AttnParams& p, int head_dim, cudaStream_t stream) {
    const bool ragged = p.seq_off != nullptr;
    if (!g_v3_enabled || head_dim != 64 || p.mask != nullptr || (!ragged && p.S <= 256)) return SB_ERR_UNSUPPORTED;
    if (ragged && (p.S_max <= 0 || p.S_max > p.S || (p.coords != nullptr && (p.dist_ld < p.S_max || (p.dist_ld % 64) != 0))))
        return SB_ERR_BAD_ARG;
    const bool alibi = p.coords != nullptr;
    if (alibi && (p.dist16 == nullptr || p.dscale == nullptr || p.slope == nullptr)) return SB_ERR_UNSUPPORTED;
    const long long koff = p.k - p.q;
    const long long v_rs = p.v_row_stride ? p.v_row_stride : p.row_stride;
    const long long v_bs = p.v_row_stride ? p.v_batch_stride : p.batch_stride;
    if (koff < 0 || koff + static_cast<long long>(p.H) * 64 > p.row_stride || (koff % 8) != 0 ||
        (reinterpret_cast<uintptr_t>(p.q) & 15) != 0 || (reinterpret_cast<uintptr_t>(p.v) & 15) != 0 ||
        (reinterpret_cast<uintptr_t>(p.out) & 15) != 0 || (p.out_row_stride % 8) != 0 ||
        static_cast<long long>(p.H) * 64 > v_rs || ((ragged ? p.S_max : p.S) + 127) / 128 > 65535 ||
        (reinterpret_cast<uintptr_t>(p.dist16) & 15) != 0)
        return SB_ERR_UNSUPPORTED;
    if (alibi != (p.out_f32 != 0)) return SB_ERR_UNSUPPORTED;
    CUtensorMap tm_q, tm_k, tm_v, tm_d;
    // ragged: ONE "bag" of p.S rows in the maps, the kernel adds each bag's first row to the coordinates
    const int rc = ragged ? make_maps(p.q, p.v, p.row_stride, p.row_stride * p.S, v_rs, v_rs * p.S, alibi ? p.dist16 : nullptr,
                                      1, p.S, &tm_q, &tm_k, &tm_v, &tm_d, p.dist_ld)
                          : make_maps(p.q, p.v, p.row_stride, p.batch_stride, v_rs, v_bs, alibi ? p.dist16 : nullptr, p.B, p.S,
                                      &tm_q, &tm_k, &tm_v, &tm_d);
    if (rc != SB_OK) return rc;
    const V3Out none{};
    return alibi ? launch_v3<true, false>(tm_q, tm_k, tm_v, tm_d, p, static_cast<int>(koff), none, stream)
                 : launch_v3<false, false>(tm_q, tm_k, tm_v, tm_d, p, static_cast<int>(koff), none, stream);
}

// training forward (bf16, extra outputs); tp.dist16 = bf16 distance matrix for ALiBi, tp.dist_scale = [B][2]
int attention_mil_v3_train_fwd(const AttnTrainParams& tp, int head_dim, cudaStream_t stream) {
    if (!g_v3_enabled || head_dim != 64 || tp.S <= 256) return SB_ERR_UNSUPPORTED;
    const bool alibi = tp.coords != nullptr;
    if (alibi && (tp.dist16 == nullptr || tp.dist_scale == nullptr)) return SB_ERR_UNSUPPORTED;
    const long long koff = tp.k - tp.q, voff = tp.v - tp.q;
    if (koff < 0 || koff + static_cast<long long>(tp.H) * 64 > tp.row_stride || (koff % 8) != 0 || voff < 0 ||
        (reinterpret_cast<uintptr_t>(tp.q) & 15) != 0 || (reinterpret_cast<uintptr_t>(tp.v) & 15) != 0 ||
        (reinterpret_cast<uintptr_t>(tp.out) & 15) != 0 || (reinterpret_cast<uintptr_t>(tp.osm) & 15) != 0 ||
        (tp.odv != nullptr && (reinterpret_cast<uintptr_t>(tp.odv) & 15) != 0) || (tp.out_row_stride % 8) != 0 ||
        (tp.out_batch_stride % 8) != 0 || (tp.S + 127) / 128 > 65535 || (reinterpret_cast<uintptr_t>(tp.dist16) & 15) != 0)
        return SB_ERR_UNSUPPORTED;
    AttnParams p{};
    p.q = reinterpret_cast<const __half*>(tp.q);
    p.k = reinterpret_cast<const __half*>(tp.k);
    p.v = reinterpret_cast<const __half*>(tp.v);
    p.row_stride = tp.row_stride; p.batch_stride = tp.batch_stride;
    p.out = tp.out; p.out_row_stride = tp.out_row_stride; p.out_batch_stride = tp.out_batch_stride;
    p.B = tp.B; p.S = tp.S; p.H = tp.H; p.scale_log2 = tp.scale_log2;
    p.coords = reinterpret_cast<const float*>(tp.coords);
    p.dscale = tp.dist_scale;
    const V3Out t{tp.out, tp.osm, tp.odv, tp.lse2, tp.beta, tp.inv_rm};
    CUtensorMap tm_q, tm_k, tm_v, tm_d;
    const int rc = make_maps(tp.q, tp.v, tp.row_stride, tp.batch_stride, tp.row_stride, tp.batch_stride,
                             alibi ? tp.dist16 : nullptr, tp.B, tp.S, &tm_q, &tm_k, &tm_v, &tm_d);
    if (rc != SB_OK) return rc;
    return alibi ? launch_v3<true, true>(tm_q, tm_k, tm_v, tm_d, p, static_cast<int>(koff), t, stream)
                 : launch_v3<false, true>(tm_q, tm_k, tm_v, tm_d, p, static_cast<int>(koff), t, stream);
}

}  // namespace sb
