// Weight gradients on tcgen05:  dW[Nout, Kin] += dY[M, Nout]^T . X[M, Kin]   (bf16 in, fp32 out)
//
// replaces the parameter-gradient half of autograd's LinearBackward for every nn.Linear of the MIL
// aggregator (src/stamp/modeling/models/vision_tranformer.py:137-139,153,163-167,314-318) in the training
// step (src/stamp/modeling/models/__init__.py:239-286).
//
// Both operands stay in their natural token-major layout: the contraction runs over token rows, so dY^T and
// X are *MN-major* UMMA operands -- a TMA box of 64 tokens x 64 channels with the 128-byte swizzle is read
// by the tensor core with the channel dimension as M (resp. N) and the token rows as K; no transposed
// copies of the activations are ever written.  CTA tile 128 (Nout) x 128 (Kin), fp32 accumulator in 128
// TMEM columns, 4-stage TMA ring over 64-token chunks.  The token range is split over gridDim.z; every split
// stores its fp32 tile with plain vector stores into a scratch slab [split][Nout][Kin] and a second,
// bandwidth-bound kernel adds the slabs into dW (deterministic; atomics from ~18 splits finishing at the same
// time onto the same addresses cost more than the contraction itself).  Without scratch the tile is reduced
// with red.global.add.v4.f32.  warp 0 TMA, warp 1 MMA issue, warps 2-5 epilogue (one TMEM lane = one dW row).
#include "common.cuh"
#include "gemm.cuh"
#include "wgrad_tc.cuh"

namespace sb {
namespace {

constexpr int WT_THREADS = 192;
constexpr int WT_STAGES = 4;
constexpr int WT_BLOCK = 64 * 128;             // 64 tokens x 64 channels bf16
constexpr int WT_STAGE_BYTES = 4 * WT_BLOCK;   // dY: 2 channel blocks, X: 2 channel blocks
constexpr int WT_SMEM = WT_STAGES * WT_STAGE_BYTES + 256 + 1024;

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};\n" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__global__ void __launch_bounds__(WT_THREADS, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap tm_y, const __grid_constant__ CUtensorMap tm_x,
                float* __restrict__ dW, long long ldw, int M, int Nout, int Kin, int chunks_per_split,
                float* __restrict__ partial) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>(
        (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + WT_STAGES * WT_STAGE_BYTES);
    uint64_t* full = bars;                    // [WT_STAGES] TMA -> MMA
    uint64_t* empty = bars + WT_STAGES;       // [WT_STAGES] MMA -> TMA
    uint64_t* accfull = bars + 2 * WT_STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * WT_STAGES + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n0 = blockIdx.x * 128, k0 = blockIdx.y * 128;
    const int total_chunks = (M + 63) / 64;
    const int c_begin = blockIdx.z * chunks_per_split;
    const int c_end = min(total_chunks, c_begin + chunks_per_split);
    const int nc = c_end - c_begin;           // > 0 by construction of the grid

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tm_y);
        tma_prefetch_desc(&tm_x);
        for (int i = 0; i < WT_STAGES; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], 1);
        }
        mbar_init(accfull, 1);
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, 128);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            for (int i = 0; i < nc; ++i) {
                const int s = i % WT_STAGES;
                mbar_wait(&empty[s], ((i / WT_STAGES) & 1) ^ 1);
                mbar_expect_tx(&full[s], WT_STAGE_BYTES);
                uint8_t* st = smem + s * WT_STAGE_BYTES;
                const int row = (c_begin + i) * 64;
                tma_load_3d(st, &tm_y, &full[s], n0, row, 0);
                tma_load_3d(st + WT_BLOCK, &tm_y, &full[s], n0 + 64, row, 0);
                tma_load_3d(st + 2 * WT_BLOCK, &tm_x, &full[s], k0, row, 0);
                tma_load_3d(st + 3 * WT_BLOCK, &tm_x, &full[s], k0 + 64, row, 0);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_f16(128, 128, true, true, true);   // bf16, A and B MN-major
            for (int i = 0; i < nc; ++i) {
                const int s = i % WT_STAGES;
                mbar_wait(&full[s], (i / WT_STAGES) & 1);
                tc_fence_after();
                const uint32_t st = smem_u32(smem + s * WT_STAGE_BYTES);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    // 16 token rows per step = 2048 B inside each 64-channel block; the two 64-channel blocks of
                    // an operand are WT_BLOCK bytes apart (leading-dimension byte offset of the MN-major layout)
                    const uint64_t a_desc = umma_desc_mn128(st + k * 2048, WT_BLOCK);
                    const uint64_t b_desc = umma_desc_mn128(st + 2 * WT_BLOCK + k * 2048, WT_BLOCK);
                    umma_f16_ss(tmem, a_desc, b_desc, idesc, (i | k) != 0);
                }
                umma_commit(&empty[s]);
            }
            umma_commit(accfull);
        }
    } else {
        // epilogue: warp w may touch TMEM lanes 32 * (w % 4) ...; lane = dW row inside the tile
        const int quarter = warp & 3;
        const int row = n0 + quarter * 32 + lane;
        mbar_wait(accfull, 0);
        tc_fence_after();
        const uint32_t t_lane = tmem + (static_cast<uint32_t>(quarter * 32) << 16);
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
            uint32_t v[32];
            tmem_ld_32x32b_x32(t_lane + c * 32, v);
            tmem_ld_wait();
            if (row < Nout) {
                if (partial != nullptr) {
                    float* dst = partial + (static_cast<long long>(blockIdx.z) * Nout + row) * Kin + k0 + c * 32;
#pragma unroll
                    for (int j = 0; j < 32; j += 4)
                        if (k0 + c * 32 + j < Kin)      // Kin % 4 == 0: a quad is inside or outside as a whole
                            *reinterpret_cast<float4*>(dst + j) = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]),
                                                                             __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
                } else {
                    float* dst = dW + static_cast<long long>(row) * ldw + k0 + c * 32;
#pragma unroll
                    for (int j = 0; j < 32; j += 4)
                        if (k0 + c * 32 + j < Kin)
                            red_add_v4(dst + j, __uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                                       __uint_as_float(v[j + 3]));
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem, 128);
    }
}

// dW[r, c] += sum_z partial[z][r][c]
__global__ void __launch_bounds__(256)
wgrad_reduce_kernel(const float* __restrict__ partial, int splits, float* __restrict__ dW, long long ldw, int Nout, int Kin) {
    const int kq = Kin >> 2;
    const long long total = static_cast<long long>(Nout) * kq;
    const long long slab = static_cast<long long>(Nout) * Kin;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long r = i / kq;
        const int c = static_cast<int>(i % kq) * 4;
        float4 acc = *reinterpret_cast<float4*>(dW + r * ldw + c);
        for (int z = 0; z < splits; ++z) {
            const float4 v = *reinterpret_cast<const float4*>(partial + z * slab + r * Kin + c);
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        *reinterpret_cast<float4*>(dW + r * ldw + c) = acc;
    }
}

int g_wgrad_tc_enabled = 1;

}  // namespace

void wgrad_tc_enable(int on) { g_wgrad_tc_enabled = on; }

size_t wgrad_tc_scratch_bytes(int Nout, int Kin) {
    const long long tiles = static_cast<long long>((Nout + 127) / 128) * ((Kin + 127) / 128);
    const long long max_splits = (2 * 148 + tiles - 1) / tiles;
    return static_cast<size_t>(max_splits) * Nout * Kin * sizeof(float);
}

int wgrad_tc(const uint16_t* dY, long long ldy, const uint16_t* X, long long ldx, float* dW, long long ldw, int M,
             int Nout, int Kin, float* scratch, size_t scratch_bytes, cudaStream_t stream) {
    auto mis = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) != 0; };
    if (!g_wgrad_tc_enabled || M < 256 || (Nout % 8) != 0 || (Kin % 8) != 0 || (ldy % 8) != 0 || (ldx % 8) != 0 ||
        (ldw % 4) != 0 || mis(dY) || mis(X) || mis(dW))
        return SB_ERR_UNSUPPORTED;
    static bool configured = false;
    if (!configured) {
        if (cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WT_SMEM) != cudaSuccess)
            return SB_ERR_CUDA;
        configured = true;
    }
    CUtensorMap tm_y, tm_x;
    int rc = make_tmap_3d_f16(&tm_y, dY, Nout, M, 1, ldy, ldy * static_cast<long long>(M), 64, 64);
    if (rc != SB_OK) return rc;
    rc = make_tmap_3d_f16(&tm_x, X, Kin, M, 1, ldx, ldx * static_cast<long long>(M), 64, 64);
    if (rc != SB_OK) return rc;
    const int tiles = ((Nout + 127) / 128) * ((Kin + 127) / 128);
    const int total_chunks = (M + 63) / 64;
    int splits = (2 * 148 + tiles - 1) / tiles;
    if (splits > total_chunks / 4) splits = total_chunks / 4;
    if (splits < 1) splits = 1;
    const int chunks_per_split = (total_chunks + splits - 1) / splits;
    splits = (total_chunks + chunks_per_split - 1) / chunks_per_split;
    dim3 grid((Nout + 127) / 128, (Kin + 127) / 128, splits);
    float* partial = nullptr;
    if (scratch != nullptr && (reinterpret_cast<uintptr_t>(scratch) & 15) == 0 && (Kin % 4) == 0 &&
        static_cast<size_t>(splits) * Nout * Kin * sizeof(float) <= scratch_bytes)
        partial = scratch;
    ProfScope prof(PROF_GEMM, 2.0 * M * static_cast<double>(Nout) * Kin, stream);
    wgrad_tc_kernel<<<grid, WT_THREADS, WT_SMEM, stream>>>(tm_y, tm_x, dW, ldw, M, Nout, Kin, chunks_per_split, partial);
    count_launch();
    if (partial != nullptr) {
        const long long quads = static_cast<long long>(Nout) * (Kin / 4);
        const int blocks = static_cast<int>(quads / 256 + 1 < 148 * 8 ? quads / 256 + 1 : 148 * 8);
        wgrad_reduce_kernel<<<blocks, 256, 0, stream>>>(partial, splits, dW, ldw, Nout, Kin);
        count_launch();
    }
    return cudaGetLastError() == cudaSuccess ? SB_OK : SB_ERR_CUDA;
}

}  // namespace sb
