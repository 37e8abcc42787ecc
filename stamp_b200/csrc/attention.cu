// Fused multi-head attention for short-to-very-long token sequences with O(S) memory.
//
//   plain  : O = softmax(Q K^T * scale) V                      (ViT tile encoder; MIL use_alibi=False)
//   ALiBi  : O = (softmax(Q K^T * scale) - c_h * Dist) V        (MIL aggregator, bias subtracted AFTER
//            the softmax exactly as the reference does)         Dist[q,k] = |coords_q - coords_k|_2
//
// reference: src/stamp/modeling/models/vision_tranformer.py:42-74 (_ALiBi.forward),
// :123-154 (MultiHeadALiBi.forward), :354-379 (mask construction); the reference materialises
// >= 5 fp32 [S,S] tensors per head, this kernel keeps one 16 x 64 tile per warp in registers and
// recomputes distances from the [S,2] coordinates on the fly.
//
// One CTA = 64 query rows of one (bag, head); 4 warps x 16 rows; key/value tiles of 64 rows are
// streamed through a cp.async double buffer; S = QK^T and the two products P.V and Dist.V run on
// the legacy tensor path (mma.sync m16n8k16, fp16 in / fp32 accumulate) because the tiles are
// 16 x 64 per warp -- the dense layers around it are the tcgen05 GEMMs.
#include "attention.cuh"

#include <math.h>

#include "common.cuh"

namespace sb {

namespace {

constexpr int BQ = 64;
constexpr int BKV = 64;
constexpr int ATT_THREADS = 128;

template <int HD>
struct AttSmem {
    static constexpr int LDS = HD + 8;  // padded row length (halfs): conflict-free ldmatrix
    static constexpr int Q_HALFS = BQ * LDS;
    static constexpr int KV_HALFS = BKV * LDS;
    static constexpr int BYTES = (Q_HALFS + 4 * KV_HALFS) * 2 + 2 * BKV * 8 + 2 * BKV;
};

template <int HD>
__device__ __forceinline__ void load_rows_async(__half* dst, const __half* src_base,
                                                long long row_stride, int row0, int nrows_valid_end,
                                                int tid) {
    constexpr int LDS = HD + 8;
    constexpr int CHUNKS_PER_ROW = HD / 8;
    constexpr int TOTAL = 64 * CHUNKS_PER_ROW;
#pragma unroll
    for (int i = 0; i < (TOTAL + ATT_THREADS - 1) / ATT_THREADS; ++i) {
        const int c = tid + i * ATT_THREADS;
        if (c < TOTAL) {
            const int r = c / CHUNKS_PER_ROW, ch = c % CHUNKS_PER_ROW;
            const int row = row0 + r;
            const bool ok = row < nrows_valid_end;
            const int rs = ok ? row : (nrows_valid_end - 1);
            cp_async_16(dst + r * LDS + ch * 8, src_base + static_cast<long long>(rs) * row_stride + ch * 8, ok);
        }
    }
}

template <int HD, bool ALIBI>
__global__ void __launch_bounds__(ATT_THREADS)
flash_attn_kernel(const AttnParams p) {
    using SM = AttSmem<HD>;
    constexpr int LDS = SM::LDS;
    constexpr int KSTEPS = HD / 16;
    constexpr int ONT = HD / 8;
    extern __shared__ __align__(16) uint8_t smem[];
    __half* Qs = reinterpret_cast<__half*>(smem);
    __half* Ks = Qs + SM::Q_HALFS;
    __half* Vs = Ks + 2 * SM::KV_HALFS;
    float2* Cs = reinterpret_cast<float2*>(Vs + 2 * SM::KV_HALFS);
    uint8_t* Ms = reinterpret_cast<uint8_t*>(Cs + 2 * BKV);

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t4 = lane & 3;
    const int b = blockIdx.x / p.H, h = blockIdx.x % p.H;
    const int q0 = blockIdx.y * BQ;
    const int S = p.S;

    const __half* qb = p.q + b * p.batch_stride + h * HD;
    const __half* kb = p.k + b * p.batch_stride + h * HD;
    const long long v_rs = p.v_row_stride ? p.v_row_stride : p.row_stride;
    const __half* vb = p.v + b * (p.v_row_stride ? p.v_batch_stride : p.batch_stride) + h * HD;
    const float2* cb = ALIBI ? reinterpret_cast<const float2*>(p.coords) + static_cast<long long>(b) * S : nullptr;
    const int mask_mode = (p.mask != nullptr) ? p.mask_mode : 0;
    // mask_mode 2 keeps a reference quirk (vision_tranformer.py:222-226): the mask is expanded with
    // .repeat(H,1,1) (rows ordered head-major) but nn.MultiheadAttention indexes it bag-major, so
    // (bag b, head h) uses the mask row of bag (b*H + h) % B.
    const int mrow = (mask_mode == 2) ? (b * p.H + h) % p.B : b;
    const uint8_t* mb = (p.mask != nullptr) ? p.mask + static_cast<long long>(mrow) * S : nullptr;

    const int ntiles = (S + BKV - 1) / BKV;

    auto load_kv = [&](int kt, int buf) {
        load_rows_async<HD>(Ks + buf * SM::KV_HALFS, kb, p.row_stride, kt * BKV, S, tid);
        load_rows_async<HD>(Vs + buf * SM::KV_HALFS, vb, v_rs, kt * BKV, S, tid);
        if (tid < BKV) {
            const int key = kt * BKV + tid;
            if constexpr (ALIBI) Cs[buf * BKV + tid] = (key < S) ? __ldg(cb + key) : make_float2(0.f, 0.f);
            if (mask_mode != 0) Ms[buf * BKV + tid] = (key < S) ? __ldg(mb + key) : 1;
        }
    };

    load_rows_async<HD>(Qs, qb, p.row_stride, q0, S, tid);
    load_kv(0, 0);
    cp_async_commit();

    const int row_a = q0 + warp * 16 + g;  // this thread's two query rows
    const int row_b = row_a + 8;
    const bool warp_active = (q0 + warp * 16) < S;

    float2 cq_a = make_float2(0.f, 0.f), cq_b = make_float2(0.f, 0.f);
    float slope = 0.f, descale = 1.f;
    if constexpr (ALIBI) {
        if (row_a < S) cq_a = __ldg(cb + row_a);
        if (row_b < S) cq_b = __ldg(cb + row_b);
        const float ds = __ldg(p.dscale + 2 * b);
        descale = __ldg(p.dscale + 2 * b + 1);
        slope = __ldg(p.slope + h) * ds;
    }
    bool mq_a = false, mq_b = false;
    if (mask_mode != 0) {
        mq_a = (row_a < S) ? (__ldg(mb + row_a) != 0) : true;
        mq_b = (row_b < S) ? (__ldg(mb + row_b) != 0) : true;
    }

    uint32_t qf[KSTEPS][4];
    float o1[ONT][4];
    float o2[ALIBI ? ONT : 1][4];
#pragma unroll
    for (int i = 0; i < ONT; ++i) { o1[i][0] = o1[i][1] = o1[i][2] = o1[i][3] = 0.f; }
#pragma unroll
    for (int i = 0; i < (ALIBI ? ONT : 1); ++i) { o2[i][0] = o2[i][1] = o2[i][2] = o2[i][3] = 0.f; }
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
            if (kt == 0) {
                const int mi = lane >> 3;
#pragma unroll
                for (int kk = 0; kk < KSTEPS; ++kk) {
                    const __half* a = Qs + (warp * 16 + (lane & 7) + (mi & 1) * 8) * LDS + kk * 16 + (mi >> 1) * 8;
                    ldmatrix_x4(qf[kk][0], qf[kk][1], qf[kk][2], qf[kk][3], smem_u32(a));
                }
            }
            const __half* Kt = Ks + buf * SM::KV_HALFS;
            const __half* Vt = Vs + buf * SM::KV_HALFS;
            const int kv_valid = min(BKV, S - kt * BKV);   // valid keys in this tile
            const int nt_valid = (kv_valid + 7) >> 3;      // n-tiles (8 keys) holding valid keys

            // ---- S = Q K^T -------------------------------------------------------------
            float s[8][4];
#pragma unroll
            for (int i = 0; i < 8; ++i) { s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f; }
            {
                const int mi = lane >> 3;
#pragma unroll
                for (int ntp = 0; ntp < 4; ++ntp) {
                    if (ntp * 2 < nt_valid) {
#pragma unroll
                        for (int kk = 0; kk < KSTEPS; ++kk) {
                            uint32_t b0, b1, b2, b3;
                            const __half* a = Kt + (ntp * 16 + (mi >> 1) * 8 + (lane & 7)) * LDS + kk * 16 + (mi & 1) * 8;
                            ldmatrix_x4(b0, b1, b2, b3, smem_u32(a));
                            mma_16816_f16(s[2 * ntp], qf[kk], b0, b1);
                            mma_16816_f16(s[2 * ntp + 1], qf[kk], b2, b3);
                        }
                    }
                }
            }

            // ---- masks, ALiBi distances, online softmax ---------------------------------
            uint32_t dfrag[ALIBI ? 4 : 1][4];
            float mx_a = -INFINITY, mx_b = -INFINITY;
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int kl = nt * 8 + 2 * t4 + e;       // key index inside the tile
                    const int key = kt * BKV + kl;
                    const bool valid = kl < kv_valid;
                    bool mk = false;
                    if (mask_mode != 0) mk = Ms[buf * BKV + kl] != 0;
                    // attention mask of the reference's masked branch:
                    //   masked(q,k) = (m_q & m_k) | (q >= 1 & k == 0)
                    const bool am_a = (mask_mode != 0) && ((mq_a && mk) || (row_a >= 1 && key == 0));
                    const bool am_b = (mask_mode != 0) && ((mq_b && mk) || (row_b >= 1 && key == 0));
                    float sa = s[nt][e], sb_ = s[nt][2 + e];
                    if (!valid || (mask_mode == 2 && am_a)) sa = -INFINITY;
                    if (!valid || (mask_mode == 2 && am_b)) sb_ = -INFINITY;
                    s[nt][e] = sa;
                    s[nt][2 + e] = sb_;
                    mx_a = fmaxf(mx_a, sa);
                    mx_b = fmaxf(mx_b, sb_);
                }
            }
            mx_a = fmaxf(mx_a, __shfl_xor_sync(0xffffffffu, mx_a, 1));
            mx_a = fmaxf(mx_a, __shfl_xor_sync(0xffffffffu, mx_a, 2));
            mx_b = fmaxf(mx_b, __shfl_xor_sync(0xffffffffu, mx_b, 1));
            mx_b = fmaxf(mx_b, __shfl_xor_sync(0xffffffffu, mx_b, 2));
            const float mn_a = fmaxf(m_a, mx_a), mn_b = fmaxf(m_b, mx_b);
            const float ms_a = (mn_a == -INFINITY) ? 0.f : mn_a * sl2;
            const float ms_b = (mn_b == -INFINITY) ? 0.f : mn_b * sl2;
            const float sc_a = exp2f(m_a * sl2 - ms_a);   // m = -inf -> 0
            const float sc_b = exp2f(m_b * sl2 - ms_b);
            m_a = mn_a;
            m_b = mn_b;
            l_a *= sc_a;
            l_b *= sc_b;
#pragma unroll
            for (int i = 0; i < ONT; ++i) {
                o1[i][0] *= sc_a; o1[i][1] *= sc_a; o1[i][2] *= sc_b; o1[i][3] *= sc_b;
            }

            uint32_t pfrag[4][4];
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) {
                float pv[4], dv[4];
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int kl = nt * 8 + 2 * t4 + e;
                    const int key = kt * BKV + kl;
                    const bool valid = kl < kv_valid;
                    float pa = exp2f(fmaf(s[nt][e], sl2, -ms_a));
                    float pb = exp2f(fmaf(s[nt][2 + e], sl2, -ms_b));
                    l_a += pa;
                    l_b += pb;
                    bool am_a = false, am_b = false;
                    if (mask_mode == 1) {
                        const bool mk = Ms[buf * BKV + kl] != 0;
                        am_a = (mq_a && mk) || (row_a >= 1 && key == 0);
                        am_b = (mq_b && mk) || (row_b >= 1 && key == 0);
                        // post-softmax masking: the weight is zeroed but stays in the denominator
                        if (am_a) pa = 0.f;
                        if (am_b) pb = 0.f;
                    }
                    pv[e] = pa;
                    pv[2 + e] = pb;
                    if constexpr (ALIBI) {
                        const float2 ck = Cs[buf * BKV + kl];
                        float dxa = cq_a.x - ck.x, dya = cq_a.y - ck.y;
                        float dxb = cq_b.x - ck.x, dyb = cq_b.y - ck.y;
                        float da = sqrtf(fmaf(dxa, dxa, dya * dya)) * slope;
                        float db = sqrtf(fmaf(dxb, dxb, dyb * dyb)) * slope;
                        // alibi mask of the masked branch: no bias on the class-token row / column
                        const bool al_a = (mask_mode == 1) && (row_a == 0 || key == 0);
                        const bool al_b = (mask_mode == 1) && (row_b == 0 || key == 0);
                        if (!valid || am_a || al_a) da = 0.f;
                        if (!valid || am_b || al_b) db = 0.f;
                        dv[e] = da;
                        dv[2 + e] = db;
                    }
                }
                const int kk2 = nt >> 1, hi = (nt & 1) * 2;
                pfrag[kk2][hi] = pack_f16(pv[0], pv[1]);
                pfrag[kk2][hi + 1] = pack_f16(pv[2], pv[3]);
                if constexpr (ALIBI) {
                    dfrag[kk2][hi] = pack_f16(dv[0], dv[1]);
                    dfrag[kk2][hi + 1] = pack_f16(dv[2], dv[3]);
                }
            }

            // ---- O1 += P V ;  O2 += Dist V ------------------------------------------------
            {
                const int mi = lane >> 3;
#pragma unroll
                for (int kk2 = 0; kk2 < 4; ++kk2) {
                    if (kk2 * 2 < nt_valid) {
#pragma unroll
                        for (int ntp = 0; ntp < ONT / 2; ++ntp) {
                            uint32_t b0, b1, b2, b3;
                            const __half* a = Vt + (kk2 * 16 + (mi & 1) * 8 + (lane & 7)) * LDS + ntp * 16 + (mi >> 1) * 8;
                            ldmatrix_x4_trans(b0, b1, b2, b3, smem_u32(a));
                            mma_16816_f16(o1[2 * ntp], pfrag[kk2], b0, b1);
                            mma_16816_f16(o1[2 * ntp + 1], pfrag[kk2], b2, b3);
                            if constexpr (ALIBI) {
                                mma_16816_f16(o2[2 * ntp], dfrag[kk2], b0, b1);
                                mma_16816_f16(o2[2 * ntp + 1], dfrag[kk2], b2, b3);
                            }
                        }
                    }
                }
            }
        }
        __syncthreads();
    }

    if (!warp_active) return;
    l_a += __shfl_xor_sync(0xffffffffu, l_a, 1);
    l_a += __shfl_xor_sync(0xffffffffu, l_a, 2);
    l_b += __shfl_xor_sync(0xffffffffu, l_b, 1);
    l_b += __shfl_xor_sync(0xffffffffu, l_b, 2);
    const float inv_a = 1.0f / l_a, inv_b = 1.0f / l_b;
    const long long obase = b * p.out_batch_stride + h * HD;
#pragma unroll
    for (int nt = 0; nt < ONT; ++nt) {
        float ya0 = o1[nt][0] * inv_a, ya1 = o1[nt][1] * inv_a;
        float yb0 = o1[nt][2] * inv_b, yb1 = o1[nt][3] * inv_b;
        if constexpr (ALIBI) {
            ya0 = fmaf(-descale, o2[nt][0], ya0);
            ya1 = fmaf(-descale, o2[nt][1], ya1);
            yb0 = fmaf(-descale, o2[nt][2], yb0);
            yb1 = fmaf(-descale, o2[nt][3], yb1);
        }
        const int col = nt * 8 + 2 * t4;
        if (p.out_f32) {
            // fp32 output pre-rounded (round-to-nearest) to TF32: it is the A operand of a
            // kind::tf32 GEMM, whose hardware conversion would otherwise truncate
            float* of = reinterpret_cast<float*>(p.out) + obase;
            const float ha0 = round_tf32(ya0), ha1 = round_tf32(ya1), hb0 = round_tf32(yb0), hb1 = round_tf32(yb1);
            if (row_a < S)
                *reinterpret_cast<float2*>(of + static_cast<long long>(row_a) * p.out_row_stride + col) = make_float2(ha0, ha1);
            if (row_b < S)
                *reinterpret_cast<float2*>(of + static_cast<long long>(row_b) * p.out_row_stride + col) = make_float2(hb0, hb1);
            if (p.out_lo != nullptr) {
                float* ol = p.out_lo + obase;
                if (row_a < S)
                    *reinterpret_cast<float2*>(ol + static_cast<long long>(row_a) * p.out_row_stride + col) =
                        make_float2(round_tf32(ya0 - ha0), round_tf32(ya1 - ha1));
                if (row_b < S)
                    *reinterpret_cast<float2*>(ol + static_cast<long long>(row_b) * p.out_row_stride + col) =
                        make_float2(round_tf32(yb0 - hb0), round_tf32(yb1 - hb1));
            }
        } else {
            __half* oh = reinterpret_cast<__half*>(p.out) + obase;
            if (row_a < S)
                *reinterpret_cast<uint32_t*>(oh + static_cast<long long>(row_a) * p.out_row_stride + col) = pack_f16(ya0, ya1);
            if (row_b < S)
                *reinterpret_cast<uint32_t*>(oh + static_cast<long long>(row_b) * p.out_row_stride + col) = pack_f16(yb0, yb1);
        }
    }
}

// per bag: bounding box of the coordinates -> power-of-two scale keeping c * dist inside fp16
__global__ void __launch_bounds__(256)
alibi_scale_kernel(const float* __restrict__ coords, const float* __restrict__ slope, int H, int S,
                   float* __restrict__ dscale) {
    const int b = blockIdx.x;
    const float2* c = reinterpret_cast<const float2*>(coords) + static_cast<long long>(b) * S;
    float xmin = INFINITY, xmax = -INFINITY, ymin = INFINITY, ymax = -INFINITY;
    for (int i = threadIdx.x; i < S; i += blockDim.x) {
        float2 v = __ldg(c + i);
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
        xmin = fminf(red[0][0], xmin); xmax = fmaxf(red[1][0], xmax);
        ymin = fminf(red[2][0], ymin); ymax = fmaxf(red[3][0], ymax);
        const float diag = sqrtf((xmax - xmin) * (xmax - xmin) + (ymax - ymin) * (ymax - ymin));
        float cmax = 0.f;
        for (int h = 0; h < H; ++h) cmax = fmaxf(cmax, fabsf(slope[h]));
        const float bound = diag * cmax;  // largest |c_h * dist| in this bag
        int e = 0;
        if (bound > 16384.f) e = static_cast<int>(ceilf(log2f(bound / 16384.f)));
        dscale[2 * b] = exp2f(static_cast<float>(-e));
        dscale[2 * b + 1] = exp2f(static_cast<float>(e));
    }
}

template <int HD, bool ALIBI>
int launch_attn(const AttnParams& p, cudaStream_t stream) {
    static bool configured = false;
    constexpr int bytes = AttSmem<HD>::BYTES;
    if (!configured) {
        if (cudaFuncSetAttribute(flash_attn_kernel<HD, ALIBI>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes) != cudaSuccess)
            return SB_ERR_CUDA;
        configured = true;
    }
    const int nq = (p.q_rows > 0 && p.q_rows < p.S) ? p.q_rows : p.S;
    dim3 grid(p.B * p.H, (nq + BQ - 1) / BQ);  // (bag, head) on x: heatmaps call with B = n_tiles
    // algorithmic FLOPs of the reference math: QK^T and one [S,S]x[S,hd] product per head
    ProfScope prof(PROF_ATTN, 4.0 * p.B * p.H * static_cast<double>(p.S) * p.S * HD, stream);
    flash_attn_kernel<HD, ALIBI><<<grid, ATT_THREADS, bytes, stream>>>(p);
    count_launch();
    return cudaGetLastError() == cudaSuccess ? SB_OK : SB_ERR_CUDA;
}

}  // namespace

int attention_fwd(const AttnParams& p, int head_dim, cudaStream_t stream) {
    if (p.B <= 0 || p.S <= 0 || p.H <= 0 || p.q == nullptr || p.out == nullptr) return SB_ERR_BAD_ARG;
    if (static_cast<long long>(p.B) * p.H > 2147483647LL || (p.S + BQ - 1) / BQ > 65535) return SB_ERR_UNSUPPORTED;
    if ((p.row_stride % 8) != 0 || (p.batch_stride % 8) != 0 || (p.v_row_stride % 8) != 0 || (p.v_batch_stride % 8) != 0 || (p.out_row_stride % 2) != 0)
        return SB_ERR_BAD_ARG;
    // short unmasked sequences (ViT tiles): tcgen05 kernel; anything else: the general kernel below
    {
        // (head dimension 64 beyond 256 tokens is a feature bag: the long-bag kernels below; head dimension 80 is a
        //  ViT-H tile encoder with 261 ... 265 tokens)
        int rc = (p.S <= 256 || head_dim == 80) ? attention_vit_stream_fwd(p, head_dim, stream) : SB_ERR_UNSUPPORTED;
        if (rc != SB_ERR_UNSUPPORTED) return rc;
        rc = attention_tc_fwd(p, head_dim, stream);
        if (rc != SB_ERR_UNSUPPORTED) return rc;
    }
    // long unmasked bags (MIL aggregator, plain or ALiBi): tcgen05 kernels, newest generation first
    {
        int rc = attention_mil_v3_fwd(p, head_dim, stream);
        if (rc != SB_ERR_UNSUPPORTED) return rc;
        rc = attention_mil_tc_fwd(p, head_dim, stream);
        if (rc != SB_ERR_UNSUPPORTED) return rc;
    }
    const bool alibi = p.coords != nullptr;
    if (alibi && (p.slope == nullptr || p.dscale == nullptr)) return SB_ERR_BAD_ARG;
    if (head_dim == 64) return alibi ? launch_attn<64, true>(p, stream) : launch_attn<64, false>(p, stream);
    if (head_dim == 80 && !alibi) return launch_attn<80, false>(p, stream);
    if (head_dim == 32) return alibi ? launch_attn<32, true>(p, stream) : launch_attn<32, false>(p, stream);
    return SB_ERR_UNSUPPORTED;
}

int alibi_dist_scale(const float* coords, const float* slope, int B, int S, int H, float* dscale,
                     cudaStream_t stream) {
    if (B <= 0 || S <= 0 || H <= 0) return SB_ERR_BAD_ARG;
    alibi_scale_kernel<<<B, 256, 0, stream>>>(coords, slope, H, S, dscale);
    count_launch();
    return cudaGetLastError() == cudaSuccess ? SB_OK : SB_ERR_CUDA;
}

}  // namespace sb
