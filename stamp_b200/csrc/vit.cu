// Tile-encoder ViT forward (UNI ViT-L/16, Virchow2 ViT-H/14 and any timm-style plain ViT):
// the host-side sequencing of the sm_100a kernels for one batch of uint8 tiles.
//
// replaces: model(tiles.to(device)) at src/stamp/preprocessing/__init__.py:322-327, i.e. timm's
// VisionTransformer.forward as constructed in src/stamp/preprocessing/extractor/uni.py:26-31 and
// virchow2.py:24-42 (arithmetic restated in oracle/vit_oracle.py).
//
// Data layout in HBM (M = B * T token rows, D = width):
//   x    fp32 [M, D]            residual stream (never rounded to 16 bit)
//   xn   fp16 [M, D]            LayerNorm output / attention output (GEMM A operands)
//   big  fp16 [M, max(3D, H')]  qkv projection, then the MLP hidden activations (also the patch matrix)
// Per block: LN -> GEMM(qkv) -> attention -> GEMM(proj, +LayerScale +residual) -> LN ->
//            GEMM(fc1, GELU | SwiGLU) -> GEMM(fc2, +LayerScale +residual).
#include <math.h>

#include "attention.cuh"
#include "common.cuh"
#include "gemm.cuh"
#include "rowops.cuh"
#include "stamp_b200.h"

namespace {

// feats[b] = [ normalised class token | mean over the other T-1 normalised tokens ]  (fp32 tokens [B, T, D] -> fp16)
// grid (D / 256 rounded up, B): a thread owns one channel and walks the tile's tokens (coalesced across the warp)
__global__ void __launch_bounds__(256)
cls_mean_pool_kernel(const float* __restrict__ tok, int T, int D, __half* __restrict__ feats) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= D) return;
    const float* t = tok + static_cast<long long>(blockIdx.y) * T * D + c;
    float acc = 0.f;
    for (int r = 1; r < T; ++r) acc += t[static_cast<long long>(r) * D];
    __half* o = feats + static_cast<long long>(blockIdx.y) * 2 * D;
    o[c] = __float2half_rn(t[0]);
    o[D + c] = __float2half_rn(acc / static_cast<float>(T - 1));
}

struct Layout {
    long long M, T, D, hid_out, big_cols;
    size_t off_x, off_xn, off_big, total;
};

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

bool make_layout(const StampVitConfig* c, int B, Layout* L) {
    if (c == nullptr || B <= 0 || c->patch <= 0 || c->img % c->patch != 0 || c->heads <= 0 ||
        c->dim % c->heads != 0 || c->dim % 8 != 0 || c->mlp_hidden % 8 != 0 || c->kpad % 8 != 0 ||
        c->kpad < 3 * c->patch * c->patch || c->pool < 0 || c->pool > 1)
        return false;
    const long long np = static_cast<long long>(c->img / c->patch) * (c->img / c->patch);
    L->T = np + 1 + c->reg_tokens;
    L->M = L->T * B;
    L->D = c->dim;
    L->hid_out = (c->mlp_kind == 1) ? c->mlp_hidden / 2 : c->mlp_hidden;
    long long big = 3LL * c->dim;
    if (c->mlp_hidden > big) big = c->mlp_hidden;
    if (c->kpad > big) big = c->kpad;
    L->big_cols = big;
    L->off_x = 0;
    L->off_xn = align_up(L->off_x + static_cast<size_t>(L->M) * L->D * 4, 1024);
    L->off_big = align_up(L->off_xn + static_cast<size_t>(L->M) * L->D * 2, 1024);
    L->total = align_up(L->off_big + static_cast<size_t>(L->M) * big * 2, 1024);
    return true;
}

}  // namespace

extern "C" {

size_t stamp_vit_workspace_bytes(const StampVitConfig* cfg, int B) {
    Layout L;
    if (!make_layout(cfg, B, &L)) return 0;
    return L.total;
}

int stamp_vit_forward(const StampVitConfig* cfg, const StampVitWeights* w,
                      const StampVitBlock* blocks, const uint8_t* tiles, void* feats16, int B,
                      void* workspace, size_t workspace_bytes, void* stream_) {
    using namespace sb;
    Layout L;
    if (!make_layout(cfg, B, &L) || w == nullptr || blocks == nullptr || tiles == nullptr ||
        feats16 == nullptr || workspace == nullptr)
        return SB_ERR_BAD_ARG;
    if (workspace_bytes < L.total) return SB_ERR_WORKSPACE;
    if ((reinterpret_cast<uintptr_t>(workspace) & 255) != 0) return SB_ERR_BAD_ARG;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    uint8_t* ws = static_cast<uint8_t*>(workspace);
    float* x = reinterpret_cast<float*>(ws + L.off_x);
    __half* xn = reinterpret_cast<__half*>(ws + L.off_xn);
    __half* big = reinterpret_cast<__half*>(ws + L.off_big);

    const int D = cfg->dim, T = static_cast<int>(L.T), M = static_cast<int>(L.M);
    const int np = T - 1 - cfg->reg_tokens, npre = 1 + cfg->reg_tokens;
    const int hd = D / cfg->heads;
    int rc;

    // tiles -> patch matrix -> patch-embed GEMM straight into the residual stream (+bias +pos_embed)
    rc = tiles_to_patches(tiles, big, B, cfg->img, cfg->patch, cfg->kpad, cfg->mean, cfg->std, 0, stream);
    if (rc != SB_OK) return rc;
    {
        GemmParams p{};
        p.M = B * np; p.N = D; p.K = cfg->kpad;
        p.act = ACT_NONE; p.store = ST_32;
        p.out = x; p.ldo = D;
        p.bias = w->patch_b;
        p.table = w->pos; p.ldt = D;
        p.gin = np; p.gout = T; p.goff = npre;
        rc = gemm_tn(big, cfg->kpad, w->patch_w, cfg->kpad, p, stream);
        if (rc != SB_OK) return rc;
    }
    rc = fill_rows(x, D, B, T, 0, w->prefix, D, nullptr, 0, npre, D, stream);
    if (rc != SB_OK) return rc;

    for (int l = 0; l < cfg->depth; ++l) {
        const StampVitBlock& b = blocks[l];
        // global_pool = 'token' (uni.py:26-31, virchow2.py:24-30 take x[:, 0]): the last block's outputs are only
        // read at the class-token rows, so after its K / V projections everything runs on those B rows alone
        // (row b*T of x and xn addressed with a row pitch of T*D; same arithmetic for the rows that are kept)
        const bool cls_only = (l == cfg->depth - 1) && cfg->pool == 0;
        const int Mr = cls_only ? B : M;                                  // rows from the attention output on
        const long long ldr = cls_only ? static_cast<long long>(T) * D : D;  // their pitch in x / xn
        rc = layernorm(x, D, b.ln1_w, b.ln1_b, xn, nullptr, D, M, D, cfg->ln_eps, 0, stream);
        if (rc != SB_OK) return rc;
        {
            GemmParams p{};
            p.M = M; p.N = 3 * D; p.K = D;
            p.store = ST_16; p.out = big; p.ldo = 3 * D; p.bias = b.qkv_b;
            rc = gemm_tn(xn, D, b.qkv_w, D, p, stream);
            if (rc != SB_OK) return rc;
        }
        {
            AttnParams a{};
            a.q = big; a.k = big + D; a.v = big + 2 * D;
            a.row_stride = 3LL * D; a.batch_stride = 3LL * D * T;
            a.out = xn; a.out_f32 = 0; a.out_row_stride = D; a.out_batch_stride = static_cast<long long>(D) * T;
            a.B = B; a.S = T; a.H = cfg->heads;
            a.q_rows = cls_only ? 1 : 0;
            a.scale_log2 = (1.0f / sqrtf(static_cast<float>(hd))) * 1.4426950408889634f;
            rc = attention_fwd(a, hd, stream);
            if (rc != SB_OK) return rc;
        }
        {
            GemmParams p{};
            p.M = Mr; p.N = D; p.K = D;
            p.store = ST_RESID32; p.out = x; p.ldo = ldr; p.bias = b.proj_b; p.gamma = b.ls1;
            rc = gemm_tn(xn, ldr, b.proj_w, D, p, stream);
            if (rc != SB_OK) return rc;
        }
        // (cls_only: the normalised class-token rows are written compactly, [B, D], over the head of xn -- the
        //  attention output there has been consumed by the projection above)
        rc = layernorm(x, ldr, b.ln2_w, b.ln2_b, xn, nullptr, D, Mr, D, cfg->ln_eps, 0, stream);
        if (rc != SB_OK) return rc;
        {
            GemmParams p{};
            p.M = Mr; p.N = cfg->mlp_hidden; p.K = D;
            p.act = (cfg->mlp_kind == 1) ? ACT_NONE : ACT_GELU;
            p.store = (cfg->mlp_kind == 1) ? ST_SWIGLU16 : ST_16;
            p.out = big; p.ldo = L.hid_out; p.bias = b.fc1_b;
            rc = gemm_tn(xn, D, b.fc1_w, D, p, stream);
            if (rc != SB_OK) return rc;
        }
        {
            GemmParams p{};
            p.M = Mr; p.N = D; p.K = static_cast<int>(L.hid_out);
            p.store = ST_RESID32; p.out = x; p.ldo = ldr; p.bias = b.fc2_b; p.gamma = b.ls2;
            rc = gemm_tn(big, L.hid_out, b.fc2_w, L.hid_out, p, stream);
            if (rc != SB_OK) return rc;
        }
    }
    if (cfg->pool == 1) {
        // every token through the final norm (fp32, into the big scratch: M x D x 4 <= M x 3D x 2 bytes), then the
        // class token and the mean of the rest side by side
        float* tok = reinterpret_cast<float*>(big);
        rc = layernorm(x, D, w->norm_w, w->norm_b, tok, nullptr, D, M, D, cfg->ln_eps, 2, stream);
        if (rc != SB_OK) return rc;
        cls_mean_pool_kernel<<<dim3((D + 255) / 256, B), 256, 0, stream>>>(tok, T, D, static_cast<__half*>(feats16));
        count_launch();
        return cudaGetLastError() == cudaSuccess ? SB_OK : SB_ERR_CUDA;
    }
    // final norm on the class-token rows only (global_pool='token'): row b*T of x
    return layernorm(x, static_cast<long long>(D) * T, w->norm_w, w->norm_b, feats16, nullptr, D, B, D,
                     cfg->ln_eps, 0, stream);
}

}  // extern "C"
