// ALiBi attention for MIL *training*: forward that keeps what the backward needs, and the backward.
//
//   forward  : O = (P - beta_h * Dhat) V,  P = softmax(Q K^T * scale),  Dhat = |c_q - c_k|_2 * inv_rm_h
//              (bias subtracted AFTER the softmax, vision_tranformer.py:58-72; inv_rm_h = 1 / running_mean
//              of _RunningMeanScaler :23-31, beta_h = _ALiBi.bias_scale :40); mask=None branch only --
//              the one every Lightning step takes (src/stamp/modeling/models/__init__.py:286,293,300).
//              Besides O (bf16, operand of mhsa.fc) it stores Osm = P V and Dhat V (fp32) and the row
//              log-sum-exp (log2 domain) so that the backward never materialises an [S,S] tensor either.
//   backward : with dW = dO V^T,  delta_q = dO_q . Osm_q,  dS = P * (dW - delta_q):
//              dV = (P - beta Dhat)^T dO      dK = scale * dS^T Q      dQ = scale * dS K
//              dbeta_h = - sum_{b,q,k} Dhat_qk dW_qk = - sum_{b,q} dO_q . (Dhat V)_q   (forward keeps Dhat V)
//              (what autograd derives for torch.softmax / einsum / cdist in _ALiBi.forward; coordinates
//              and running_mean carry no gradient).
//
// Two backward kernels, both O(S) memory: `dkv` owns a 64-key block and streams query tiles
// (transposed tiles S^T = K Q^T, dW^T = V dO^T, so P^T / dS^T come out in the A-fragment layout of
// the dV / dK products), `dq` owns a 64-query block and streams key tiles.  bf16 operands, fp32
// accumulation, legacy tensor path (mma.sync m16n8k16): 16 x 64 tiles per warp as in attention.cu.
#include "attention_train.cuh"

#include <math.h>

#include "common.cuh"

namespace sb {

namespace {

constexpr int TQ = 64;
constexpr int TK = 64;
constexpr int TTHREADS = 128;

__device__ __forceinline__ void mma_16816_bf16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 "
        "{%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// 64 rows x HD bf16 -> padded smem tile; rows >= rows_end are zero filled
template <int HD>
__device__ __forceinline__ void load_tile_async(uint16_t* dst, const uint16_t* src_base, long long row_stride,
                                                int row0, int rows_end, int tid) {
    constexpr int LDS = HD + 8;
    constexpr int CPR = HD / 8;
    constexpr int TOTAL = 64 * CPR;
#pragma unroll
    for (int i = 0; i < (TOTAL + TTHREADS - 1) / TTHREADS; ++i) {
        const int c = tid + i * TTHREADS;
        if (c < TOTAL) {
            const int r = c / CPR, ch = c % CPR;
            const int row = row0 + r;
            const bool ok = row < rows_end;
            const int rs = ok ? row : (rows_end - 1);
            cp_async_16(dst + r * LDS + ch * 8, src_base + static_cast<long long>(rs) * row_stride + ch * 8, ok);
        }
    }
}

// A fragments (16 rows of this warp x HD) from a padded [64][LDS] tile
template <int HD>
__device__ __forceinline__ void load_a_frags(uint32_t (&f)[HD / 16][4], const uint16_t* tile, int warp, int lane) {
    constexpr int LDS = HD + 8;
    const int mi = lane >> 3;
#pragma unroll
    for (int kk = 0; kk < HD / 16; ++kk) {
        const uint16_t* a = tile + (warp * 16 + (lane & 7) + (mi & 1) * 8) * LDS + kk * 16 + (mi >> 1) * 8;
        ldmatrix_x4(f[kk][0], f[kk][1], f[kk][2], f[kk][3], smem_u32(a));
    }
}

// acc[16 x 64] = A(frags, 16 x HD) . T^T, T = padded tile of 64 rows x HD (rows = output columns)
template <int HD>
__device__ __forceinline__ void mma_a_tileT(float (&acc)[8][4], const uint32_t (&af)[HD / 16][4],
                                            const uint16_t* tile, int lane) {
    constexpr int LDS = HD + 8;
    const int mi = lane >> 3;
#pragma unroll
    for (int ntp = 0; ntp < 4; ++ntp) {
#pragma unroll
        for (int kk = 0; kk < HD / 16; ++kk) {
            uint32_t b0, b1, b2, b3;
            const uint16_t* a = tile + (ntp * 16 + (mi >> 1) * 8 + (lane & 7)) * LDS + kk * 16 + (mi & 1) * 8;
            ldmatrix_x4(b0, b1, b2, b3, smem_u32(a));
            mma_16816_bf16(acc[2 * ntp], af[kk], b0, b1);
            mma_16816_bf16(acc[2 * ntp + 1], af[kk], b2, b3);
        }
    }
}

// acc[16 x HD] += F(frags, 16 x 64) . T, T = padded tile of 64 rows (contraction) x HD
template <int HD>
__device__ __forceinline__ void mma_frag_tile(float (&acc)[HD / 8][4], const uint32_t (&f)[4][4],
                                              const uint16_t* tile, int lane) {
    constexpr int LDS = HD + 8;
    const int mi = lane >> 3;
#pragma unroll
    for (int kk2 = 0; kk2 < 4; ++kk2) {
#pragma unroll
        for (int ntp = 0; ntp < HD / 16; ++ntp) {
            uint32_t b0, b1, b2, b3;
            const uint16_t* a = tile + (kk2 * 16 + (mi & 1) * 8 + (lane & 7)) * LDS + ntp * 16 + (mi >> 1) * 8;
            ldmatrix_x4_trans(b0, b1, b2, b3, smem_u32(a));
            mma_16816_bf16(acc[2 * ntp], f[kk2], b0, b1);
            mma_16816_bf16(acc[2 * ntp + 1], f[kk2], b2, b3);
        }
    }
}

template <int N>
__device__ __forceinline__ void zero_acc(float (&a)[N][4]) {
#pragma unroll
    for (int i = 0; i < N; ++i) { a[i][0] = a[i][1] = a[i][2] = a[i][3] = 0.f; }
}

// ------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------
template <int HD>
struct FwdSmem {
    static constexpr int LDS = HD + 8;
    static constexpr int TILE = 64 * LDS;
    static constexpr int BYTES = 5 * TILE * 2 + 2 * TK * 8;
};

template <int HD, bool ALIBI>
__global__ void __launch_bounds__(TTHREADS)
attn_train_fwd_kernel(const AttnTrainParams p) {
    using SM = FwdSmem<HD>;
    constexpr int ONT = HD / 8;
    extern __shared__ __align__(16) uint8_t smem[];
    uint16_t* Qs = reinterpret_cast<uint16_t*>(smem);
    uint16_t* Ks = Qs + SM::TILE;
    uint16_t* Vs = Ks + 2 * SM::TILE;
    float2* Cs = reinterpret_cast<float2*>(Vs + 2 * SM::TILE);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t4 = lane & 3;
    const int b = blockIdx.x / p.H, h = blockIdx.x % p.H;
    const int q0 = blockIdx.y * TQ;
    const int S = p.S;
    const uint16_t* qb = p.q + b * p.batch_stride + h * HD;
    const uint16_t* kb = p.k + b * p.batch_stride + h * HD;
    const uint16_t* vb = p.v + b * p.batch_stride + h * HD;
    const float2* cb = ALIBI ? p.coords + static_cast<long long>(b) * S : nullptr;
    const int ntiles = (S + TK - 1) / TK;

    auto load_kv = [&](int kt, int buf) {
        load_tile_async<HD>(Ks + buf * SM::TILE, kb, p.row_stride, kt * TK, S, tid);
        load_tile_async<HD>(Vs + buf * SM::TILE, vb, p.row_stride, kt * TK, S, tid);
        if constexpr (ALIBI) {
            if (tid < TK) {
                const int key = kt * TK + tid;
                Cs[buf * TK + tid] = (key < S) ? __ldg(cb + key) : make_float2(0.f, 0.f);
            }
        }
    };
    load_tile_async<HD>(Qs, qb, p.row_stride, q0, S, tid);
    load_kv(0, 0);
    cp_async_commit();

    const int row_a = q0 + warp * 16 + g, row_b = row_a + 8;
    const bool warp_active = (q0 + warp * 16) < S;
    float2 cq_a = make_float2(0.f, 0.f), cq_b = make_float2(0.f, 0.f);
    float inv_rm = 0.f, beta = 0.f;
    if constexpr (ALIBI) {
        if (row_a < S) cq_a = __ldg(cb + row_a);
        if (row_b < S) cq_b = __ldg(cb + row_b);
        inv_rm = __ldg(p.inv_rm + h);
        beta = __ldg(p.beta + h);
    }

    uint32_t qf[HD / 16][4];
    float o1[ONT][4];
    float o2[ALIBI ? ONT : 1][4];
    zero_acc(o1);
    zero_acc(o2);
    float m_a = -INFINITY, m_b = -INFINITY, l_a = 0.f, l_b = 0.f;
    const float sl2 = p.scale_log2;

    for (int kt = 0; kt < ntiles; ++kt) {
        const int buf = kt & 1;
        if (kt + 1 < ntiles) {
            load_kv(kt + 1, buf ^ 1);
            cp_async_commit();
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        if (warp_active) {
            if (kt == 0) load_a_frags<HD>(qf, Qs, warp, lane);
            const uint16_t* Kt = Ks + buf * SM::TILE;
            const uint16_t* Vt = Vs + buf * SM::TILE;
            const int kv_valid = min(TK, S - kt * TK);
            float s[8][4];
            zero_acc(s);
            mma_a_tileT<HD>(s, qf, Kt, lane);

            float mx_a = -INFINITY, mx_b = -INFINITY;
#pragma unroll
            for (int nt = 0; nt < 8; ++nt)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const bool valid = (nt * 8 + 2 * t4 + e) < kv_valid;
                    if (!valid) { s[nt][e] = -INFINITY; s[nt][2 + e] = -INFINITY; }
                    mx_a = fmaxf(mx_a, s[nt][e]);
                    mx_b = fmaxf(mx_b, s[nt][2 + e]);
                }
            mx_a = fmaxf(mx_a, __shfl_xor_sync(0xffffffffu, mx_a, 1));
            mx_a = fmaxf(mx_a, __shfl_xor_sync(0xffffffffu, mx_a, 2));
            mx_b = fmaxf(mx_b, __shfl_xor_sync(0xffffffffu, mx_b, 1));
            mx_b = fmaxf(mx_b, __shfl_xor_sync(0xffffffffu, mx_b, 2));
            const float mn_a = fmaxf(m_a, mx_a), mn_b = fmaxf(m_b, mx_b);   // finite: >= 1 valid key per tile
            const float ms_a = mn_a * sl2, ms_b = mn_b * sl2;
            const float sc_a = exp2f(m_a * sl2 - ms_a), sc_b = exp2f(m_b * sl2 - ms_b);
            m_a = mn_a; m_b = mn_b;
            l_a *= sc_a; l_b *= sc_b;
#pragma unroll
            for (int i = 0; i < ONT; ++i) { o1[i][0] *= sc_a; o1[i][1] *= sc_a; o1[i][2] *= sc_b; o1[i][3] *= sc_b; }

            uint32_t pfrag[4][4];
            uint32_t dfrag[ALIBI ? 4 : 1][4];
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) {
                float pv[4], dv[4];
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int kl = nt * 8 + 2 * t4 + e;
                    const bool valid = kl < kv_valid;
                    const float pa = exp2f(fmaf(s[nt][e], sl2, -ms_a));      // -inf -> 0
                    const float pb = exp2f(fmaf(s[nt][2 + e], sl2, -ms_b));
                    l_a += pa; l_b += pb;
                    pv[e] = pa; pv[2 + e] = pb;
                    if constexpr (ALIBI) {
                        const float2 ck = Cs[buf * TK + kl];
                        const float dxa = cq_a.x - ck.x, dya = cq_a.y - ck.y;
                        const float dxb = cq_b.x - ck.x, dyb = cq_b.y - ck.y;
                        dv[e] = valid ? sqrtf(fmaf(dxa, dxa, dya * dya)) * inv_rm : 0.f;
                        dv[2 + e] = valid ? sqrtf(fmaf(dxb, dxb, dyb * dyb)) * inv_rm : 0.f;
                    }
                }
                const int kk2 = nt >> 1, hi = (nt & 1) * 2;
                pfrag[kk2][hi] = pack_bf16(pv[0], pv[1]);
                pfrag[kk2][hi + 1] = pack_bf16(pv[2], pv[3]);
                if constexpr (ALIBI) {
                    dfrag[kk2][hi] = pack_bf16(dv[0], dv[1]);
                    dfrag[kk2][hi + 1] = pack_bf16(dv[2], dv[3]);
                }
            }
            mma_frag_tile<HD>(o1, pfrag, Vt, lane);
            if constexpr (ALIBI) mma_frag_tile<HD>(o2, dfrag, Vt, lane);
        }
        __syncthreads();
    }
    if (!warp_active) return;
    l_a += __shfl_xor_sync(0xffffffffu, l_a, 1);
    l_a += __shfl_xor_sync(0xffffffffu, l_a, 2);
    l_b += __shfl_xor_sync(0xffffffffu, l_b, 1);
    l_b += __shfl_xor_sync(0xffffffffu, l_b, 2);
    const float inv_a = 1.0f / l_a, inv_b = 1.0f / l_b;
    if (t4 == 0) {
        float* L = p.lse2 + (static_cast<long long>(b) * p.H + h) * S;
        if (row_a < S) L[row_a] = fmaf(m_a, sl2, log2f(l_a));
        if (row_b < S) L[row_b] = fmaf(m_b, sl2, log2f(l_b));
    }
    const long long obase = b * p.out_batch_stride + h * HD;
#pragma unroll
    for (int nt = 0; nt < ONT; ++nt) {
        const float sa0 = o1[nt][0] * inv_a, sa1 = o1[nt][1] * inv_a;
        const float sb0 = o1[nt][2] * inv_b, sb1 = o1[nt][3] * inv_b;
        float ya0 = sa0, ya1 = sa1, yb0 = sb0, yb1 = sb1;
        if constexpr (ALIBI) {
            ya0 = fmaf(-beta, o2[nt][0], ya0); ya1 = fmaf(-beta, o2[nt][1], ya1);
            yb0 = fmaf(-beta, o2[nt][2], yb0); yb1 = fmaf(-beta, o2[nt][3], yb1);
        }
        const int col = nt * 8 + 2 * t4;
        if (row_a < S) {
            const long long o = obase + static_cast<long long>(row_a) * p.out_row_stride + col;
            *reinterpret_cast<uint32_t*>(p.out + o) = pack_bf16(ya0, ya1);
            *reinterpret_cast<float2*>(p.osm + o) = make_float2(sa0, sa1);
            if constexpr (ALIBI) *reinterpret_cast<float2*>(p.odv + o) = make_float2(o2[nt][0], o2[nt][1]);
        }
        if (row_b < S) {
            const long long o = obase + static_cast<long long>(row_b) * p.out_row_stride + col;
            *reinterpret_cast<uint32_t*>(p.out + o) = pack_bf16(yb0, yb1);
            *reinterpret_cast<float2*>(p.osm + o) = make_float2(sb0, sb1);
            if constexpr (ALIBI) *reinterpret_cast<float2*>(p.odv + o) = make_float2(o2[nt][2], o2[nt][3]);
        }
    }
}

// Per token row and head, from the fp32 dO the fc data-gradient GEMM produced:
//   delta[b,h,q] = dO . Osm        dbeta[h] -= dO . (Dhat V)        dO16 = bf16(dO) for the tensor-core kernels
// dbeta uses the forward's own fp32 Dhat.V accumulator: exactly the derivative of the computed forward,
// without a second pass over the [S,S] distances.  One warp per token row; per-CTA partials, then atomics.
__global__ void __launch_bounds__(256)
attn_delta_kernel(const float* __restrict__ dout32, const float* __restrict__ osm, const float* __restrict__ odv,
                  long long row_stride, uint16_t* __restrict__ dout16, float* __restrict__ delta,
                  float* __restrict__ dbeta, int B, int S, int H, int hd) {
    __shared__ float part[64];   // H <= 64
    const int lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    for (int i = threadIdx.x; i < H; i += blockDim.x) part[i] = 0.f;
    __syncthreads();
    const long long rows = static_cast<long long>(B) * S;
    for (long long row = static_cast<long long>(blockIdx.x) * nwarps + (threadIdx.x >> 5); row < rows;
         row += static_cast<long long>(gridDim.x) * nwarps) {
        const int b = static_cast<int>(row / S), q = static_cast<int>(row % S);
        const float* d = dout32 + row * row_stride;
        const float* o = osm + row * row_stride;
        const float* v = (odv != nullptr) ? odv + row * row_stride : nullptr;
        uint16_t* d16 = dout16 + row * row_stride;
        for (int h = 0; h < H; ++h) {
            float a = 0.f, bsum = 0.f;
            for (int c = lane * 2; c < hd; c += 64) {
                const float2 dv = *reinterpret_cast<const float2*>(d + h * hd + c);
                const float2 ov = *reinterpret_cast<const float2*>(o + h * hd + c);
                // delta must cancel against sum_k P dW, and dW is formed from the bf16-rounded dO:
                // use the same rounded values here (dS = P (dW - delta) is a difference of near-equal terms)
                const uint32_t pk = pack_bf16(dv.x, dv.y);
                a = fmaf(__uint_as_float(pk << 16), ov.x, a);
                a = fmaf(__uint_as_float(pk & 0xffff0000u), ov.y, a);
                if (v != nullptr) {
                    const float2 vv = *reinterpret_cast<const float2*>(v + h * hd + c);
                    bsum = fmaf(dv.x, vv.x, bsum);
                    bsum = fmaf(dv.y, vv.y, bsum);
                }
                *reinterpret_cast<uint32_t*>(d16 + h * hd + c) = pk;
            }
            a = warp_sum(a);
            bsum = warp_sum(bsum);
            if (lane == 0) {
                delta[(static_cast<long long>(b) * H + h) * S + q] = a;
                if (v != nullptr) atomicAdd(part + h, bsum);
            }
        }
    }
    __syncthreads();
    if (odv != nullptr)
        for (int i = threadIdx.x; i < H; i += blockDim.x) atomicAdd(dbeta + i, -part[i]);
}

// ------------------------------------------------------------------------------------------------
// backward: dQ.  CTA = 64 queries of one (bag, head); streams key tiles.
// ------------------------------------------------------------------------------------------------
template <int HD>
struct DqSmem {
    static constexpr int LDS = HD + 8;
    static constexpr int TILE = 64 * LDS;
    static constexpr int BYTES = 6 * TILE * 2;   // Q, dO, K x2, V x2
};

template <int HD>
__global__ void __launch_bounds__(TTHREADS)
attn_bwd_dq_kernel(const AttnTrainParams p) {
    using SM = DqSmem<HD>;
    constexpr int ONT = HD / 8;
    extern __shared__ __align__(16) uint8_t smem[];
    uint16_t* Qs = reinterpret_cast<uint16_t*>(smem);
    uint16_t* Os = Qs + SM::TILE;
    uint16_t* Ks = Os + SM::TILE;
    uint16_t* Vs = Ks + 2 * SM::TILE;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t4 = lane & 3;
    const int b = blockIdx.x / p.H, h = blockIdx.x % p.H;
    const int q0 = blockIdx.y * TQ;
    const int S = p.S;
    const uint16_t* qb = p.q + b * p.batch_stride + h * HD;
    const uint16_t* kb = p.k + b * p.batch_stride + h * HD;
    const uint16_t* vb = p.v + b * p.batch_stride + h * HD;
    const uint16_t* dob = p.dout + b * p.out_batch_stride + h * HD;
    const int ntiles = (S + TK - 1) / TK;

    auto load_kv = [&](int kt, int buf) {
        load_tile_async<HD>(Ks + buf * SM::TILE, kb, p.row_stride, kt * TK, S, tid);
        load_tile_async<HD>(Vs + buf * SM::TILE, vb, p.row_stride, kt * TK, S, tid);
    };
    load_tile_async<HD>(Qs, qb, p.row_stride, q0, S, tid);
    load_tile_async<HD>(Os, dob, p.out_row_stride, q0, S, tid);
    load_kv(0, 0);
    cp_async_commit();

    const int row_a = q0 + warp * 16 + g, row_b = row_a + 8;
    const bool warp_active = (q0 + warp * 16) < S;
    const long long sbase = (static_cast<long long>(b) * p.H + h) * S;
    const float lse_a = (row_a < S) ? __ldg(p.lse2 + sbase + row_a) : INFINITY;
    const float lse_b = (row_b < S) ? __ldg(p.lse2 + sbase + row_b) : INFINITY;
    const float dl_a = (row_a < S) ? __ldg(p.delta + sbase + row_a) : 0.f;
    const float dl_b = (row_b < S) ? __ldg(p.delta + sbase + row_b) : 0.f;

    uint32_t qf[HD / 16][4], dof[HD / 16][4];
    float dq[ONT][4];
    zero_acc(dq);
    const float sl2 = p.scale_log2;

    for (int kt = 0; kt < ntiles; ++kt) {
        const int buf = kt & 1;
        if (kt + 1 < ntiles) {
            load_kv(kt + 1, buf ^ 1);
            cp_async_commit();
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        if (warp_active) {
            if (kt == 0) {
                load_a_frags<HD>(qf, Qs, warp, lane);
                load_a_frags<HD>(dof, Os, warp, lane);
            }
            const uint16_t* Kt = Ks + buf * SM::TILE;
            const uint16_t* Vt = Vs + buf * SM::TILE;
            const int kv_valid = min(TK, S - kt * TK);
            float s[8][4], dw[8][4];
            zero_acc(s);
            zero_acc(dw);
            mma_a_tileT<HD>(s, qf, Kt, lane);
            mma_a_tileT<HD>(dw, dof, Vt, lane);
            uint32_t dsf[4][4];
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) {
                float ds[4];
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const bool valid = (nt * 8 + 2 * t4 + e) < kv_valid;
                    const float pa = valid ? exp2f(fmaf(s[nt][e], sl2, -lse_a)) : 0.f;
                    const float pb = valid ? exp2f(fmaf(s[nt][2 + e], sl2, -lse_b)) : 0.f;
                    ds[e] = pa * (dw[nt][e] - dl_a);
                    ds[2 + e] = pb * (dw[nt][2 + e] - dl_b);
                }
                const int kk2 = nt >> 1, hi = (nt & 1) * 2;
                dsf[kk2][hi] = pack_bf16(ds[0], ds[1]);
                dsf[kk2][hi + 1] = pack_bf16(ds[2], ds[3]);
            }
            mma_frag_tile<HD>(dq, dsf, Kt, lane);
        }
        __syncthreads();
    }
    if (!warp_active) return;
    uint16_t* dqb = p.dq + b * p.batch_stride + h * HD;
#pragma unroll
    for (int nt = 0; nt < ONT; ++nt) {
        const int col = nt * 8 + 2 * t4;
        if (row_a < S)
            *reinterpret_cast<uint32_t*>(dqb + static_cast<long long>(row_a) * p.row_stride + col) =
                pack_bf16(dq[nt][0] * p.scale, dq[nt][1] * p.scale);
        if (row_b < S)
            *reinterpret_cast<uint32_t*>(dqb + static_cast<long long>(row_b) * p.row_stride + col) =
                pack_bf16(dq[nt][2] * p.scale, dq[nt][3] * p.scale);
    }
}

// ------------------------------------------------------------------------------------------------
// backward: dK, dV.  CTA = 64 keys of one (bag, head); streams query tiles.
// ------------------------------------------------------------------------------------------------
template <int HD>
struct DkvSmem {
    static constexpr int LDS = HD + 8;
    static constexpr int TILE = 64 * LDS;
    // K, V, (Q, dO) x 2, then per stage: lse[64], delta[64], coords[64]
    static constexpr int BYTES = 6 * TILE * 2 + 2 * TQ * (4 + 4 + 8);
};

template <int HD, bool ALIBI>
__global__ void __launch_bounds__(TTHREADS)
attn_bwd_dkv_kernel(const AttnTrainParams p) {
    using SM = DkvSmem<HD>;
    constexpr int ONT = HD / 8;
    extern __shared__ __align__(16) uint8_t smem[];
    uint16_t* Ks = reinterpret_cast<uint16_t*>(smem);
    uint16_t* Vs = Ks + SM::TILE;
    uint16_t* Qs = Vs + SM::TILE;          // 2 stages
    uint16_t* Os = Qs + 2 * SM::TILE;      // 2 stages
    float2* Cs = reinterpret_cast<float2*>(Os + 2 * SM::TILE);   // 2 x 64
    float* Ls = reinterpret_cast<float*>(Cs + 2 * TQ);           // 2 x 64
    float* Dl = Ls + 2 * TQ;                                      // 2 x 64

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t4 = lane & 3;
    const int b = blockIdx.x / p.H, h = blockIdx.x % p.H;
    const int k0 = blockIdx.y * TK;
    const int S = p.S;
    const uint16_t* qb = p.q + b * p.batch_stride + h * HD;
    const uint16_t* kb = p.k + b * p.batch_stride + h * HD;
    const uint16_t* vb = p.v + b * p.batch_stride + h * HD;
    const uint16_t* dob = p.dout + b * p.out_batch_stride + h * HD;
    const float2* cb = ALIBI ? p.coords + static_cast<long long>(b) * S : nullptr;
    const long long sbase = (static_cast<long long>(b) * p.H + h) * S;
    const int ntiles = (S + TQ - 1) / TQ;

    auto load_q = [&](int qt, int buf) {
        load_tile_async<HD>(Qs + buf * SM::TILE, qb, p.row_stride, qt * TQ, S, tid);
        load_tile_async<HD>(Os + buf * SM::TILE, dob, p.out_row_stride, qt * TQ, S, tid);
        if (tid < TQ) {
            const int q = qt * TQ + tid;
            const bool ok = q < S;
            Ls[buf * TQ + tid] = ok ? __ldg(p.lse2 + sbase + q) : INFINITY;
            Dl[buf * TQ + tid] = ok ? __ldg(p.delta + sbase + q) : 0.f;
            if constexpr (ALIBI) Cs[buf * TQ + tid] = ok ? __ldg(cb + q) : make_float2(0.f, 0.f);
        }
    };
    load_tile_async<HD>(Ks, kb, p.row_stride, k0, S, tid);
    load_tile_async<HD>(Vs, vb, p.row_stride, k0, S, tid);
    load_q(0, 0);
    cp_async_commit();

    const int key_a = k0 + warp * 16 + g, key_b = key_a + 8;
    const bool warp_active = (k0 + warp * 16) < S;
    float2 ck_a = make_float2(0.f, 0.f), ck_b = make_float2(0.f, 0.f);
    float inv_rm = 0.f, beta = 0.f;
    if constexpr (ALIBI) {
        if (key_a < S) ck_a = __ldg(cb + key_a);
        if (key_b < S) ck_b = __ldg(cb + key_b);
        inv_rm = __ldg(p.inv_rm + h);
        beta = __ldg(p.beta + h);
    }

    uint32_t kf[HD / 16][4], vf[HD / 16][4];
    float dk[ONT][4], dv[ONT][4];
    zero_acc(dk);
    zero_acc(dv);
    const float sl2 = p.scale_log2;

    for (int qt = 0; qt < ntiles; ++qt) {
        const int buf = qt & 1;
        if (qt + 1 < ntiles) {
            load_q(qt + 1, buf ^ 1);
            cp_async_commit();
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        if (warp_active) {
            if (qt == 0) {
                load_a_frags<HD>(kf, Ks, warp, lane);
                load_a_frags<HD>(vf, Vs, warp, lane);
            }
            const uint16_t* Qt = Qs + buf * SM::TILE;
            const uint16_t* Ot = Os + buf * SM::TILE;
            const int q_valid = min(TQ, S - qt * TQ);
            float st[8][4], dwt[8][4];     // [key row][query col]
            zero_acc(st);
            zero_acc(dwt);
            mma_a_tileT<HD>(st, kf, Qt, lane);
            mma_a_tileT<HD>(dwt, vf, Ot, lane);
            uint32_t wf[4][4], dsf[4][4];
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) {
                float w[4], ds[4];
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int ql = nt * 8 + 2 * t4 + e;
                    const bool valid = ql < q_valid;
                    const float lse = Ls[buf * TQ + ql], dl = Dl[buf * TQ + ql];
                    const float pa = valid ? exp2f(fmaf(st[nt][e], sl2, -lse)) : 0.f;
                    const float pb = valid ? exp2f(fmaf(st[nt][2 + e], sl2, -lse)) : 0.f;
                    float wa = pa, wb = pb;
                    if constexpr (ALIBI) {
                        const float2 cq = Cs[buf * TQ + ql];
                        const float dxa = ck_a.x - cq.x, dya = ck_a.y - cq.y;
                        const float dxb = ck_b.x - cq.x, dyb = ck_b.y - cq.y;
                        const float da = valid ? sqrtf(fmaf(dxa, dxa, dya * dya)) * inv_rm : 0.f;
                        const float db = valid ? sqrtf(fmaf(dxb, dxb, dyb * dyb)) * inv_rm : 0.f;
                        wa = fmaf(-beta, da, pa);
                        wb = fmaf(-beta, db, pb);
                    }
                    w[e] = wa; w[2 + e] = wb;
                    ds[e] = pa * (dwt[nt][e] - dl);
                    ds[2 + e] = pb * (dwt[nt][2 + e] - dl);
                }
                const int kk2 = nt >> 1, hi = (nt & 1) * 2;
                wf[kk2][hi] = pack_bf16(w[0], w[1]);
                wf[kk2][hi + 1] = pack_bf16(w[2], w[3]);
                dsf[kk2][hi] = pack_bf16(ds[0], ds[1]);
                dsf[kk2][hi + 1] = pack_bf16(ds[2], ds[3]);
            }
            mma_frag_tile<HD>(dv, wf, Ot, lane);
            mma_frag_tile<HD>(dk, dsf, Qt, lane);
        }
        __syncthreads();
    }
    if (!warp_active) return;
    uint16_t* dkb = p.dk + b * p.batch_stride + h * HD;
    uint16_t* dvb = p.dv + b * p.batch_stride + h * HD;
#pragma unroll
    for (int nt = 0; nt < ONT; ++nt) {
        const int col = nt * 8 + 2 * t4;
        if (key_a < S) {
            const long long o = static_cast<long long>(key_a) * p.row_stride + col;
            *reinterpret_cast<uint32_t*>(dkb + o) = pack_bf16(dk[nt][0] * p.scale, dk[nt][1] * p.scale);
            *reinterpret_cast<uint32_t*>(dvb + o) = pack_bf16(dv[nt][0], dv[nt][1]);
        }
        if (key_b < S) {
            const long long o = static_cast<long long>(key_b) * p.row_stride + col;
            *reinterpret_cast<uint32_t*>(dkb + o) = pack_bf16(dk[nt][2] * p.scale, dk[nt][3] * p.scale);
            *reinterpret_cast<uint32_t*>(dvb + o) = pack_bf16(dv[nt][2], dv[nt][3]);
        }
    }
}

template <typename K>
int set_smem(K kernel, int bytes) {
    return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes) == cudaSuccess ? SB_OK : SB_ERR_CUDA;
}

bool params_ok(const AttnTrainParams& p) {
    if (p.B <= 0 || p.S <= 0 || p.H <= 0 || p.q == nullptr || p.k == nullptr || p.v == nullptr) return false;
    if (static_cast<long long>(p.B) * p.H > 2147483647LL || (p.S + TQ - 1) / TQ > 65535) return false;
    if ((p.row_stride % 8) != 0 || (p.batch_stride % 8) != 0 || (p.out_row_stride % 8) != 0 || (p.out_batch_stride % 8) != 0)
        return false;
    if (p.coords != nullptr && (p.beta == nullptr || p.inv_rm == nullptr)) return false;
    return true;
}

template <int HD>
int fwd_hd(const AttnTrainParams& p, cudaStream_t stream) {
    const bool alibi = p.coords != nullptr;
    constexpr int bytes = FwdSmem<HD>::BYTES;
    static bool configured = false;
    if (!configured) {
        if (set_smem(attn_train_fwd_kernel<HD, true>, bytes) != SB_OK || set_smem(attn_train_fwd_kernel<HD, false>, bytes) != SB_OK)
            return SB_ERR_CUDA;
        configured = true;
    }
    dim3 grid(p.B * p.H, (p.S + TQ - 1) / TQ);
    ProfScope prof(PROF_ATTN, 4.0 * p.B * p.H * static_cast<double>(p.S) * p.S * HD, stream);
    if (alibi) attn_train_fwd_kernel<HD, true><<<grid, TTHREADS, bytes, stream>>>(p);
    else attn_train_fwd_kernel<HD, false><<<grid, TTHREADS, bytes, stream>>>(p);
    count_launch();
    return cudaGetLastError() == cudaSuccess ? SB_OK : SB_ERR_CUDA;
}

template <int HD>
int bwd_hd(const AttnTrainParams& p, cudaStream_t stream) {
    const bool alibi = p.coords != nullptr;
    static bool configured = false;
    if (!configured) {
        if (set_smem(attn_bwd_dq_kernel<HD>, DqSmem<HD>::BYTES) != SB_OK ||
            set_smem(attn_bwd_dkv_kernel<HD, true>, DkvSmem<HD>::BYTES) != SB_OK ||
            set_smem(attn_bwd_dkv_kernel<HD, false>, DkvSmem<HD>::BYTES) != SB_OK)
            return SB_ERR_CUDA;
        configured = true;
    }
    {
        const long long rows = static_cast<long long>(p.B) * p.S;
        long long blocks = (rows + 7) / 8;
        if (blocks > 148 * 8) blocks = 148 * 8;
        ProfScope prof(PROF_ROWOP, rows * p.H * HD * (alibi ? 14.0 : 10.0), stream);
        attn_delta_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(
            p.dout32, p.osm, alibi ? p.odv : nullptr, p.out_row_stride, p.dout, p.delta, p.dbeta, p.B, p.S, p.H, HD);
        count_launch();
    }
    {
        const int rc = attention_train_tc_bwd(p, HD, stream);   // long bags, head_dim 64: tcgen05 kernels
        if (rc != SB_ERR_UNSUPPORTED) return rc;
    }
    dim3 grid(p.B * p.H, (p.S + TQ - 1) / TQ);
    // algorithmic FLOPs of the reference backward: dV, dW, dQ, dK = four [S,S]x[S,hd] products per head
    ProfScope prof(PROF_ATTN, 8.0 * p.B * p.H * static_cast<double>(p.S) * p.S * HD, stream);
    if (alibi) attn_bwd_dkv_kernel<HD, true><<<grid, TTHREADS, DkvSmem<HD>::BYTES, stream>>>(p);
    else attn_bwd_dkv_kernel<HD, false><<<grid, TTHREADS, DkvSmem<HD>::BYTES, stream>>>(p);
    attn_bwd_dq_kernel<HD><<<grid, TTHREADS, DqSmem<HD>::BYTES, stream>>>(p);
    count_launch(2);
    return cudaGetLastError() == cudaSuccess ? SB_OK : SB_ERR_CUDA;
}

}  // namespace

int attention_train_fwd(const AttnTrainParams& p, int head_dim, cudaStream_t stream) {
    if (!params_ok(p) || p.out == nullptr || p.osm == nullptr || p.lse2 == nullptr ||
        (p.coords != nullptr && p.odv == nullptr))
        return SB_ERR_BAD_ARG;
    {
        int rc = attention_mil_v3_train_fwd(p, head_dim, stream);
        if (rc != SB_ERR_UNSUPPORTED) return rc;
        rc = attention_mil_tc_train_fwd(p, head_dim, stream);
        if (rc != SB_ERR_UNSUPPORTED) return rc;
    }
    if (head_dim == 64) return fwd_hd<64>(p, stream);
    if (head_dim == 32) return fwd_hd<32>(p, stream);
    return SB_ERR_UNSUPPORTED;
}

int attention_train_bwd(const AttnTrainParams& p, int head_dim, cudaStream_t stream) {
    if (!params_ok(p) || p.dout == nullptr || p.dout32 == nullptr || p.osm == nullptr || p.lse2 == nullptr ||
        p.delta == nullptr || p.dq == nullptr || p.dk == nullptr || p.dv == nullptr || p.H > 64 ||
        (p.coords != nullptr && (p.dbeta == nullptr || p.odv == nullptr)))
        return SB_ERR_BAD_ARG;
    if (head_dim == 64) return bwd_hd<64>(p, stream);
    if (head_dim == 32) return bwd_hd<32>(p, stream);
    return SB_ERR_UNSUPPORTED;
}

}  // namespace sb
