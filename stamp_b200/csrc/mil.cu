// ALiBi Transformer-MIL aggregator forward (inference): host-side sequencing of the sm_100a
// kernels for a batch of feature bags, plus the three small kernels only this path needs.
//
// replaces: VisionTransformer.forward and everything below it in
// src/stamp/modeling/models/vision_tranformer.py:332-384 (-> Transformer.forward :281-295 ->
// SelfAttention.forward :194-242 -> MultiHeadALiBi.forward :123-154 -> _ALiBi.forward :42-74,
// feed_forward :157-169), as called by LitTileClassifier.predict_step / validation_step
// (src/stamp/modeling/models/__init__.py:302-313) and heatmaps_ (src/stamp/heatmaps/__init__.py:392,419).
//
// Data layout in HBM (S = N + 1 tokens per bag, M = B * S rows, d = dim_model):
//   x    fp32 [M, d]   residual stream           xn  fp16 [M, d]   LayerNorm output
//   qkv  fp16 [M, 3d]  packed per-head q|k|v     att fp32 [M, d]   ALiBi attention output (TF32-rounded;
//                                                                   fp16 [M, d] for the nn.MultiheadAttention variant)
//   h    fp16 [M, ff]  feed-forward hidden       bags16 fp16 [B*N, F]
#include <math.h>

#include "attention.cuh"
#include "common.cuh"
#include "gemm.cuh"
#include "mil_common.cuh"
#include "rowops.cuh"
#include "stamp_b200.h"

namespace sb {
namespace {

// fp32 -> fp16 (round to nearest), 8 elements per thread, 128-bit loads/stores
__global__ void __launch_bounds__(256)
cast_f32_f16_kernel(const float* __restrict__ in, __half* __restrict__ out, long long n8, long long n) {
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n8;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const uint4 a = ld_nc_v4(in + i * 8), b = ld_nc_v4(in + i * 8 + 4);
        uint4 w;
        w.x = pack_f16(__uint_as_float(a.x), __uint_as_float(a.y));
        w.y = pack_f16(__uint_as_float(a.z), __uint_as_float(a.w));
        w.z = pack_f16(__uint_as_float(b.x), __uint_as_float(b.y));
        w.w = pack_f16(__uint_as_float(b.z), __uint_as_float(b.w));
        *reinterpret_cast<uint4*>(out + i * 8) = w;
    }
    // tail (n % 8 elements)
    if (blockIdx.x == 0 && threadIdx.x < (n - n8 * 8)) {
        const long long j = n8 * 8 + threadIdx.x;
        out[j] = __float2half_rn(in[j]);
    }
}

// coords [B,N,2] -> coords_s [B,S,2] with the class token at (0,0);  mask [B,N] -> mask_s [B,S]
// with the class token unmasked (vision_tranformer.py:349-351,360-362)
__global__ void __launch_bounds__(256)
mil_prepare_kernel(const float* __restrict__ coords, const uint8_t* __restrict__ mask,
                   float2* __restrict__ coords_s, uint8_t* __restrict__ mask_s, int B, int N) {
    const int S = N + 1;
    const long long total = static_cast<long long>(B) * S;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int s = static_cast<int>(i % S);
        const long long b = i / S;
        float2 c = make_float2(0.f, 0.f);
        uint8_t m = 0;
        if (s > 0) {
            const long long j = b * N + (s - 1);
            if (coords_s != nullptr) c = __ldg(reinterpret_cast<const float2*>(coords) + j);
            if (mask != nullptr) m = __ldg(mask + j);
        }
        if (coords_s != nullptr) coords_s[i] = c;
        if (mask_s != nullptr) mask_s[i] = m;
    }
}

// logits[b, :] = head(LayerNorm(x[b * S, :]))  -- the score-producing tail stays in fp32
// (transformer.norm + [:, 0] + mlp_head, vision_tranformer.py:293,382-384). One CTA per bag.
__global__ void __launch_bounds__(256)
cls_head_kernel(const float* __restrict__ x, long long row_stride, int d, int d_real, const float* __restrict__ nw,
                const float* __restrict__ nb, const float* __restrict__ hw, const float* __restrict__ hb,
                int C, float eps, float* __restrict__ logits, const int* __restrict__ row_index = nullptr) {
    extern __shared__ float sh[];  // d normalised values + 32 scratch
    float* y = sh;
    float* red = sh + d;
    // ragged batches: the class token of bag b sits in row row_index[b] (row pitch row_stride)
    const float* xr = x + (row_index != nullptr ? static_cast<long long>(row_index[blockIdx.x]) : blockIdx.x) * row_stride;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw_ = blockDim.x >> 5;
    float s = 0.f;
    for (int i = tid; i < d; i += blockDim.x) s += xr[i];
    s = warp_sum(s);
    if (lane == 0) red[warp] = s;
    __syncthreads();
    float tot = 0.f;
    for (int i = 0; i < nw_; ++i) tot += red[i];
    const float mean = tot / d_real;      // (zero-padded width: the padding channels are zeros)
    __syncthreads();
    float q = 0.f;
    for (int i = tid; i < d_real; i += blockDim.x) { const float t = xr[i] - mean; q += t * t; }
    q = warp_sum(q);
    if (lane == 0) red[warp] = q;
    __syncthreads();
    tot = 0.f;
    for (int i = 0; i < nw_; ++i) tot += red[i];
    const float rstd = rsqrtf(tot / d_real + eps);
    for (int i = tid; i < d; i += blockDim.x) y[i] = (xr[i] - mean) * rstd * nw[i] + nb[i];
    __syncthreads();
    for (int c = warp; c < C; c += nw_) {
        float a = 0.f;
        for (int i = lane; i < d; i += 32) a = fmaf(y[i], __ldg(hw + static_cast<long long>(c) * d + i), a);
        a = warp_sum(a);
        if (lane == 0) logits[static_cast<long long>(blockIdx.x) * C + c] = a + hb[c];
    }
}

// ragged batches: x[seq_off[b], :] = class_token for every bag b
__global__ void __launch_bounds__(256)
fill_cls_rows_kernel(float* __restrict__ x, int d, const int* __restrict__ seq_off, const float* __restrict__ cls) {
    float* row = x + static_cast<long long>(seq_off[blockIdx.x]) * d;
    for (int i = threadIdx.x; i < d; i += blockDim.x) row[i] = __ldg(cls + i);
}

inline int grid_for(long long total, int block) {
    long long g = (total + block - 1) / block;
    const long long cap = 148LL * 16;
    return static_cast<int>(g < cap ? (g > 0 ? g : 1) : cap);
}

struct Layout {
    long long M, S;
    size_t off_x, off_xn, off_qkv, off_att, off_h, off_bags, off_coords, off_mask, off_dscale, off_dist, off_bbox, total;
    bool v3;   // long unmasked ALiBi bags: third-generation attention kernel fed by a pre-computed distance matrix
};

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

bool make_layout(const StampMilConfig* c, int B, int N, Layout* L) {
    if (c == nullptr || B <= 0 || N < 0 || c->dim_input <= 0 || c->dim_input % 8 != 0 ||
        c->dim_model % 8 != 0 || c->dim_ff % 8 != 0 || c->n_heads <= 0 || c->dim_model % c->n_heads != 0 ||
        c->dim_output <= 0 || c->n_layers < 0)
        return false;
    const int hd = c->dim_model / c->n_heads;
    if (hd != 64 && hd != 32 && !(hd == 80 && !c->use_alibi)) return false;
    if (c->use_alibi && (c->dim_model % 32) != 0) return false;  // K-wrap of the split-precision GEMMs
    L->S = N + 1LL;
    L->M = L->S * B;
    const size_t d = c->dim_model, M = L->M;
    size_t o = 0;
    L->off_x = o;      o = align_up(o + M * d * 4, 256);
    L->off_xn = o;     o = align_up(o + M * 2 * d * 2, 256);   // [M, 2d] fp16: LayerNorm output hi | lo
    L->off_qkv = o;    o = align_up(o + M * 3 * d * 2, 256);
    L->off_att = o;    o = align_up(o + M * 2 * d * 4, 256);   // [M, 2d] fp32: attention output hi | lo
    L->off_h = o;      o = align_up(o + M * c->dim_ff * 2, 256);
    L->off_bags = o;   o = align_up(o + static_cast<size_t>(B) * N * c->dim_input * 2 + 16, 256);
    L->off_coords = o; o = align_up(o + M * 8, 256);
    L->off_mask = o;   o = align_up(o + M, 256);
    L->off_dscale = o; o = align_up(o + static_cast<size_t>(B) * 8 * (c->n_layers > 0 ? c->n_layers : 1), 256);
    // (sized for the unmasked call; a masked call of the same shape simply leaves it unused)
    L->v3 = c->use_alibi && hd == 64 && L->S > 256 && L->S <= 65535 && B <= 65535;
    L->off_dist = o;   if (L->v3) o = align_up(o + mil_dist16_bytes(B, static_cast<int>(L->S)), 1024);
    L->off_bbox = o;   if (L->v3) o = align_up(o + mil_dist16_scratch_bytes(B), 256);
    L->total = o;
    return true;
}

}  // namespace

int mil_prepare(const float* coords, const uint8_t* mask, float2* coords_s, uint8_t* mask_s, int B, int N,
                cudaStream_t stream) {
    if (B <= 0 || N < 0 || (coords_s != nullptr && coords == nullptr && N > 0)) return SB_ERR_BAD_ARG;
    mil_prepare_kernel<<<grid_for(static_cast<long long>(B) * (N + 1), 256), 256, 0, stream>>>(coords, mask, coords_s,
                                                                                          mask_s, B, N);
    count_launch();
    return cudaGetLastError() == cudaSuccess ? SB_OK : SB_ERR_CUDA;
}

int cls_head(const float* x, long long bag_stride, int d, const float* norm_w, const float* norm_b,
             const float* head_w, const float* head_b, int C, int B, float* logits, cudaStream_t stream) {
    if (x == nullptr || logits == nullptr || B <= 0 || d <= 0 || C <= 0) return SB_ERR_BAD_ARG;
    cls_head_kernel<<<B, 256, (d + 32) * sizeof(float), stream>>>(x, bag_stride, d, d, norm_w, norm_b, head_w, head_b, C,
                                                                 1e-5f, logits);
    count_launch();
    return cudaGetLastError() == cudaSuccess ? SB_OK : SB_ERR_CUDA;
}

}  // namespace sb

extern "C" {

size_t stamp_mil_workspace_bytes(const StampMilConfig* cfg, int B, int N) {
    sb::Layout L;
    if (!sb::make_layout(cfg, B, N, &L)) return 0;
    return L.total;
}

int stamp_mil_forward(const StampMilConfig* cfg, const StampMilWeights* w, const StampMilLayer* layers,
                      const void* bags, int bags_f16, const float* coords, const uint8_t* mask, float* logits,
                      int B, int N, void* workspace, size_t workspace_bytes, void* stream_) {
    using namespace sb;
    Layout L;
    if (!make_layout(cfg, B, N, &L) || w == nullptr || (cfg->n_layers > 0 && layers == nullptr) ||
        logits == nullptr || workspace == nullptr || (N > 0 && (bags == nullptr || coords == nullptr)))
        return SB_ERR_BAD_ARG;
    if (workspace_bytes < L.total) return SB_ERR_WORKSPACE;
    if ((reinterpret_cast<uintptr_t>(workspace) & 255) != 0) return SB_ERR_BAD_ARG;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    uint8_t* ws = static_cast<uint8_t*>(workspace);
    float* x = reinterpret_cast<float*>(ws + L.off_x);
    __half* xn = reinterpret_cast<__half*>(ws + L.off_xn);
    __half* qkv = reinterpret_cast<__half*>(ws + L.off_qkv);
    void* att = ws + L.off_att;
    __half* h = reinterpret_cast<__half*>(ws + L.off_h);
    __half* bags16 = reinterpret_cast<__half*>(ws + L.off_bags);
    float2* coords_s = reinterpret_cast<float2*>(ws + L.off_coords);
    uint8_t* mask_s = reinterpret_cast<uint8_t*>(ws + L.off_mask);
    float* dscale = reinterpret_cast<float*>(ws + L.off_dscale);

    const int d = cfg->dim_model, F = cfg->dim_input, H = cfg->n_heads, hd = d / H;
    const int S = static_cast<int>(L.S), M = static_cast<int>(L.M);
    const bool alibi = cfg->use_alibi != 0;
    int rc;

    // project_features: Linear + GELU into rows 1.. of every bag; class token into row 0
    if (N > 0) {
        const long long n = static_cast<long long>(B) * N * F;
        if (bags_f16) {
            // features as the .h5 files store them (fp16): they are the GEMM operand as they are
            if ((reinterpret_cast<uintptr_t>(bags) & 15) != 0) return SB_ERR_BAD_ARG;
            bags16 = const_cast<__half*>(static_cast<const __half*>(bags));
        } else {
            ProfScope prof(PROF_ROWOP, n * 6.0, stream);
            cast_f32_f16_kernel<<<grid_for(n / 8, 256), 256, 0, stream>>>(static_cast<const float*>(bags), bags16, n / 8, n);
            count_launch();
        }
        GemmParams p{};
        p.M = B * N; p.N = d; p.K = F;
        p.act = ACT_GELU; p.store = ST_32; p.out = x; p.ldo = d; p.bias = w->proj_b;
        p.gin = N; p.gout = S; p.goff = 1;
        rc = gemm_tn(bags16, F, w->proj_w, F, p, stream);
        if (rc != SB_OK) return rc;
    }
    rc = fill_rows(x, d, B, S, 0, w->class_token, d, nullptr, 0, 1, d, stream);
    if (rc != SB_OK) return rc;
    if (alibi || mask != nullptr) {
        mil_prepare_kernel<<<grid_for(L.M, 256), 256, 0, stream>>>(
            coords, mask, alibi ? coords_s : nullptr, mask != nullptr ? mask_s : nullptr, B, N);
        count_launch();
    }

    const int d_real = (cfg->dim_model_real > 0 && cfg->dim_model_real <= d) ? cfg->dim_model_real : d;
    const int hd_real = (cfg->head_dim_real > 0 && cfg->head_dim_real <= hd) ? cfg->head_dim_real : hd;
    const float scale_log2 = (1.0f / sqrtf(static_cast<float>(hd_real))) * 1.4426950408889634f;

    const bool v3 = L.v3 && mask == nullptr;
    uint16_t* dist16 = reinterpret_cast<uint16_t*>(ws + L.off_dist);
    if (v3) {
        // token distances once per call (they do not depend on the layer or the head): fp16 [B, S, S], scaled per
        // bag by a power of two; read by the attention kernel of every layer through TMA
        rc = mil_dist16(reinterpret_cast<const float*>(coords_s), B, S, 0, dscale, dist16, ws + L.off_bbox, stream);
        if (rc != SB_OK) return rc;
    } else if (alibi)  // power-of-two range scale of the distance operand, per layer (slopes differ) and bag
        for (int l = 0; l < cfg->n_layers; ++l) {
            rc = alibi_dist_scale(reinterpret_cast<const float*>(coords_s), layers[l].slope, B, S, H,
                                  dscale + static_cast<size_t>(l) * B * 2, stream);
            if (rc != SB_OK) return rc;
        }
    for (int l = 0; l < cfg->n_layers; ++l) {
        const StampMilLayer& y = layers[l];
        AttnParams a{};
        a.out = att; a.out_row_stride = d; a.out_batch_stride = static_cast<long long>(d) * S;
        a.B = B; a.S = S; a.H = H; a.scale_log2 = scale_log2;
        if (alibi) {
            // The reference's ALiBi term (unscaled distances times V) dominates the layer output by
            // orders of magnitude, so the two contractions it flows through -- the V projection and
            // fc -- run in split precision (operand = hi + lo, three tensor-core passes, fp32
            // accumulate); q/k and the softmax side are plain fp16.  Error budget: DESIGN.md.
            // LayerNorm -> xn = [hi | lo] (fp16, row pitch 2d)
            rc = layernorm_padded(x, d, y.ln1_w, y.ln1_b, xn, xn + d, 2LL * d, M, d, d_real, 1e-5f, 0, stream);
            if (rc != SB_OK) return rc;
            __half* qk = qkv;                                   // [M, 2d]
            __half* v16 = qkv + static_cast<size_t>(M) * 2 * d;  // [M, d]
            GemmParams p{};
            p.M = M; p.K = d;
            p.N = 2 * d; p.store = ST_16; p.out = qk; p.ldo = 2 * d; p.bias = y.qkv_b;
            rc = gemm_tn(xn, 2LL * d, y.qkv_w, d, p, stream);     // q | k heads from the hi half
            if (rc != SB_OK) return rc;
            // v heads, split precision in ONE launch: K-concatenated hi.Whi + lo.Whi + hi.Wlo
            p.N = d; p.K = 3 * d; p.a_kwrap = 2 * d; p.out = v16; p.ldo = d; p.bias = y.qkv_b + 2 * d;
            rc = gemm_tn(xn, 2LL * d, y.v_w3, 3LL * d, p, stream);
            if (rc != SB_OK) return rc;

            a.q = qk; a.k = qk + d; a.row_stride = 2LL * d; a.batch_stride = 2LL * d * S;
            a.v = v16; a.v_row_stride = d; a.v_batch_stride = static_cast<long long>(d) * S;
            // attention output [hi | lo] (TF32-rounded fp32, row pitch 2d)
            a.out_f32 = 1; a.out_lo = reinterpret_cast<float*>(att) + d;
            a.out_row_stride = 2LL * d; a.out_batch_stride = 2LL * d * S;
            a.coords = reinterpret_cast<const float*>(coords_s); a.slope = y.slope;
            a.dscale = v3 ? dscale : dscale + static_cast<size_t>(l) * B * 2;
            if (v3) a.dist16 = dist16;
            if (mask != nullptr) { a.mask = mask_s; a.mask_mode = 1; }
            rc = attention_fwd(a, hd, stream);
            if (rc != SB_OK) return rc;

            // x += fc(att): 3 x TF32 in one launch, att = [hi | lo] read as hi, lo, hi
            GemmParams f{};
            f.M = M; f.N = d; f.K = 3 * d; f.a_kwrap = 2 * d; f.tf32 = 1; f.store = ST_RESID32; f.out = x; f.ldo = d;
            f.bias = y.fc_b;
            rc = gemm_tn(att, 2LL * d, y.fc_w, 3LL * d, f, stream);
            if (rc != SB_OK) return rc;
        } else {
            rc = layernorm_padded(x, d, y.ln1_w, y.ln1_b, xn, nullptr, d, M, d, d_real, 1e-5f, 0, stream);
            if (rc != SB_OK) return rc;
            GemmParams p{};
            p.M = M; p.N = 3 * d; p.K = d;
            p.store = ST_16; p.out = qkv; p.ldo = 3 * d; p.bias = y.qkv_b;
            rc = gemm_tn(xn, d, y.qkv_w, d, p, stream);
            if (rc != SB_OK) return rc;
            a.q = qkv; a.k = qkv + d; a.v = qkv + 2 * d;
            a.row_stride = 3LL * d; a.batch_stride = 3LL * d * S;
            a.out_f32 = 0;
            if (mask != nullptr) { a.mask = mask_s; a.mask_mode = 2; }
            rc = attention_fwd(a, hd, stream);
            if (rc != SB_OK) return rc;
            GemmParams f{};
            f.M = M; f.N = d; f.K = d; f.store = ST_RESID32; f.out = x; f.ldo = d; f.bias = y.fc_b;
            rc = gemm_tn(att, d, y.fc_w, d, f, stream);
            if (rc != SB_OK) return rc;
        }
        rc = layernorm_padded(x, d, y.ln2_w, y.ln2_b, xn, nullptr, d, M, d, d_real, 1e-5f, 0, stream);
        if (rc != SB_OK) return rc;
        {
            GemmParams p{};
            p.M = M; p.N = cfg->dim_ff; p.K = d;
            p.act = ACT_GELU; p.store = ST_16; p.out = h; p.ldo = cfg->dim_ff; p.bias = y.ff1_b;
            rc = gemm_tn(xn, d, y.ff1_w, d, p, stream);
            if (rc != SB_OK) return rc;
        }
        {
            GemmParams p{};
            p.M = M; p.N = d; p.K = cfg->dim_ff;
            p.store = ST_RESID32; p.out = x; p.ldo = d; p.bias = y.ff2_b;
            rc = gemm_tn(h, cfg->dim_ff, y.ff2_w, cfg->dim_ff, p, stream);
            if (rc != SB_OK) return rc;
        }
    }
    cls_head_kernel<<<B, 256, (d + 32) * sizeof(float), stream>>>(
        x, static_cast<long long>(S) * d, d, d_real, w->norm_w, w->norm_b, w->head_w, w->head_b,
        cfg->dim_output, 1e-5f, logits);
    count_launch();
    return cudaGetLastError() == cudaSuccess ? SB_OK : SB_ERR_CUDA;
}

size_t stamp_mil_ragged_workspace_bytes(const StampMilConfig* cfg, int B, int total_rows, int S_max) {
    sb::Layout L;
    if (B <= 0 || total_rows < B || S_max <= 0 || S_max > total_rows || !sb::make_layout(cfg, 1, total_rows - 1, &L)) return 0;
    if (cfg->dim_model / cfg->n_heads != 64) return 0;
    size_t o = L.off_dist;   // everything up to the distance matrix is sized by the row count alone
    const size_t ld = (static_cast<size_t>(S_max) + 63) / 64 * 64;
    o = sb::align_up(o + static_cast<size_t>(B) * 8, 256);                       // per-bag distance scales
    const size_t dist = o;
    if (cfg->use_alibi) o = sb::align_up(o + static_cast<size_t>(total_rows) * ld * 2, 1024);
    o = sb::align_up(o + sb::mil_dist16_scratch_bytes(B), 256);
    (void)dist;
    return o;
}

int stamp_mil_forward_ragged(const StampMilConfig* cfg, const StampMilWeights* w, const StampMilLayer* layers,
                             const void* tokens_f16, const float* coords_s_in, const int* seq_off, int B, int total_rows,
                             int S_max, float* logits, void* workspace, size_t workspace_bytes, void* stream_) {
    using namespace sb;
    Layout L;
    if (B <= 0 || total_rows < B || S_max <= 0 || S_max > total_rows || !make_layout(cfg, 1, total_rows - 1, &L) ||
        w == nullptr || (cfg->n_layers > 0 && layers == nullptr) || tokens_f16 == nullptr || seq_off == nullptr ||
        logits == nullptr || workspace == nullptr || (cfg->use_alibi && coords_s_in == nullptr))
        return SB_ERR_BAD_ARG;
    const int d = cfg->dim_model, F = cfg->dim_input, H = cfg->n_heads, hd = d / H;
    if (hd != 64) return SB_ERR_UNSUPPORTED;                 // the long-bag tcgen05 kernel is the only ragged attention
    if (workspace_bytes < stamp_mil_ragged_workspace_bytes(cfg, B, total_rows, S_max)) return SB_ERR_WORKSPACE;
    if ((reinterpret_cast<uintptr_t>(workspace) & 255) != 0 || (reinterpret_cast<uintptr_t>(tokens_f16) & 15) != 0)
        return SB_ERR_BAD_ARG;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    uint8_t* ws = static_cast<uint8_t*>(workspace);
    float* x = reinterpret_cast<float*>(ws + L.off_x);
    __half* xn = reinterpret_cast<__half*>(ws + L.off_xn);
    __half* qkv = reinterpret_cast<__half*>(ws + L.off_qkv);
    void* att = ws + L.off_att;
    __half* h = reinterpret_cast<__half*>(ws + L.off_h);
    const long long ld = (static_cast<long long>(S_max) + 63) / 64 * 64;
    size_t o = L.off_dist;
    float* dscale = reinterpret_cast<float*>(ws + o);
    o = align_up(o + static_cast<size_t>(B) * 8, 256);
    uint16_t* dist16 = reinterpret_cast<uint16_t*>(ws + o);
    if (cfg->use_alibi) o = align_up(o + static_cast<size_t>(total_rows) * ld * 2, 1024);
    void* bbox = ws + o;
    const int M = total_rows;
    const bool alibi = cfg->use_alibi != 0;
    int rc;

    // project_features over every row (the class rows hold don't-care inputs), then the class token into row 0 of
    // every bag
    {
        GemmParams p{};
        p.M = M; p.N = d; p.K = F;
        p.act = ACT_GELU; p.store = ST_32; p.out = x; p.ldo = d; p.bias = w->proj_b;
        rc = gemm_tn(tokens_f16, F, w->proj_w, F, p, stream);
        if (rc != SB_OK) return rc;
    }
    fill_cls_rows_kernel<<<B, 256, 0, stream>>>(x, d, seq_off, w->class_token);
    count_launch();
    if (alibi) {
        rc = mil_dist16_ragged(coords_s_in, seq_off, B, S_max, ld, dscale, dist16, bbox, stream);
        if (rc != SB_OK) return rc;
    }
    const int d_real = (cfg->dim_model_real > 0 && cfg->dim_model_real <= d) ? cfg->dim_model_real : d;
    const int hd_real = (cfg->head_dim_real > 0 && cfg->head_dim_real <= hd) ? cfg->head_dim_real : hd;
    const float scale_log2 = (1.0f / sqrtf(static_cast<float>(hd_real))) * 1.4426950408889634f;

    for (int l = 0; l < cfg->n_layers; ++l) {
        const StampMilLayer& y = layers[l];
        AttnParams a{};
        a.out = att; a.B = B; a.S = M; a.S_max = S_max; a.seq_off = seq_off; a.H = H; a.scale_log2 = scale_log2;
        if (alibi) {   // same sequence as stamp_mil_forward (split-precision V projection and fc), rows = all bags
            rc = layernorm_padded(x, d, y.ln1_w, y.ln1_b, xn, xn + d, 2LL * d, M, d, d_real, 1e-5f, 0, stream);
            if (rc != SB_OK) return rc;
            __half* qk = qkv;
            __half* v16 = qkv + static_cast<size_t>(M) * 2 * d;
            GemmParams p{};
            p.M = M; p.K = d;
            p.N = 2 * d; p.store = ST_16; p.out = qk; p.ldo = 2 * d; p.bias = y.qkv_b;
            rc = gemm_tn(xn, 2LL * d, y.qkv_w, d, p, stream);
            if (rc != SB_OK) return rc;
            p.N = d; p.K = 3 * d; p.a_kwrap = 2 * d; p.out = v16; p.ldo = d; p.bias = y.qkv_b + 2 * d;
            rc = gemm_tn(xn, 2LL * d, y.v_w3, 3LL * d, p, stream);
            if (rc != SB_OK) return rc;
            a.q = qk; a.k = qk + d; a.row_stride = 2LL * d; a.batch_stride = 0;
            a.v = v16; a.v_row_stride = d; a.v_batch_stride = 0;
            a.out_f32 = 1; a.out_lo = reinterpret_cast<float*>(att) + d;
            a.out_row_stride = 2LL * d; a.out_batch_stride = 0;
            a.coords = coords_s_in; a.slope = y.slope; a.dscale = dscale; a.dist16 = dist16; a.dist_ld = ld;
            rc = attention_mil_v3_fwd(a, hd, stream);
            if (rc != SB_OK) return rc;
            GemmParams f{};
            f.M = M; f.N = d; f.K = 3 * d; f.a_kwrap = 2 * d; f.tf32 = 1; f.store = ST_RESID32; f.out = x; f.ldo = d;
            f.bias = y.fc_b;
            rc = gemm_tn(att, 2LL * d, y.fc_w, 3LL * d, f, stream);
            if (rc != SB_OK) return rc;
        } else {
            rc = layernorm_padded(x, d, y.ln1_w, y.ln1_b, xn, nullptr, d, M, d, d_real, 1e-5f, 0, stream);
            if (rc != SB_OK) return rc;
            GemmParams p{};
            p.M = M; p.N = 3 * d; p.K = d;
            p.store = ST_16; p.out = qkv; p.ldo = 3 * d; p.bias = y.qkv_b;
            rc = gemm_tn(xn, d, y.qkv_w, d, p, stream);
            if (rc != SB_OK) return rc;
            a.q = qkv; a.k = qkv + d; a.v = qkv + 2 * d;
            a.row_stride = 3LL * d; a.batch_stride = 0;
            a.out_f32 = 0; a.out_row_stride = d; a.out_batch_stride = 0;
            rc = attention_mil_v3_fwd(a, hd, stream);
            if (rc != SB_OK) return rc;
            GemmParams f{};
            f.M = M; f.N = d; f.K = d; f.store = ST_RESID32; f.out = x; f.ldo = d; f.bias = y.fc_b;
            rc = gemm_tn(att, d, y.fc_w, d, f, stream);
            if (rc != SB_OK) return rc;
        }
        rc = layernorm_padded(x, d, y.ln2_w, y.ln2_b, xn, nullptr, d, M, d, d_real, 1e-5f, 0, stream);
        if (rc != SB_OK) return rc;
        {
            GemmParams p{};
            p.M = M; p.N = cfg->dim_ff; p.K = d;
            p.act = ACT_GELU; p.store = ST_16; p.out = h; p.ldo = cfg->dim_ff; p.bias = y.ff1_b;
            rc = gemm_tn(xn, d, y.ff1_w, d, p, stream);
            if (rc != SB_OK) return rc;
        }
        {
            GemmParams p{};
            p.M = M; p.N = d; p.K = cfg->dim_ff;
            p.store = ST_RESID32; p.out = x; p.ldo = d; p.bias = y.ff2_b;
            rc = gemm_tn(h, cfg->dim_ff, y.ff2_w, cfg->dim_ff, p, stream);
            if (rc != SB_OK) return rc;
        }
    }
    cls_head_kernel<<<B, 256, (d + 32) * sizeof(float), stream>>>(x, d, d, d_real, w->norm_w, w->norm_b, w->head_w,
                                                                 w->head_b, cfg->dim_output, 1e-5f, logits, seq_off);
    count_launch();
    return cudaGetLastError() == cudaSuccess ? SB_OK : SB_ERR_CUDA;
}

}  // extern "C"
