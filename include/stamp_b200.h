/*
 * stamp_b200 -- C ABI of the B200 (sm_100a) hot path behind STAMP's plugin interfaces.
 *
 * The reference (KatherLab/STAMP v2.5.0) has no FFI: its boundary is three Python plugin
 * interfaces (Extractor, MIL backbone, Encoder).  This header is the native boundary a
 * maintainer binds from Python with ctypes; every entry point cites the
 * reference call site whose arithmetic it replaces (paths relative to the reference root).
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless named host_*;
 *   - the callee never allocates: callers pass outputs and workspaces;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *   - return value: 0 = OK, negative = error (stamp_b200_strerror); nothing throws;
 *   - re-entrant per stream; one CUDA context per process (one process per GPU).
 *   - "16-bit" matrices are IEEE fp16 unless a `bf16` flag says otherwise.
 */
#ifndef STAMP_B200_H
#define STAMP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define STAMP_B200_ABI_VERSION 1

enum {
    STAMP_OK = 0,
    STAMP_ERR_BAD_ARG = -1,
    STAMP_ERR_CUDA = -2,
    STAMP_ERR_DRIVER = -3,
    STAMP_ERR_UNSUPPORTED = -4,
    STAMP_ERR_WORKSPACE = -5
};

int stamp_b200_abi_version(void);
const char* stamp_b200_strerror(int code);
/* kernels launched by this library since load / since the last reset (bench.py "gpu_launches") */
long long stamp_b200_launch_count(void);
void stamp_b200_reset_launch_count(void);
/* Optional CUDA-event timing of the library's own launches, on the launching stream, by category
 * (0 GEMM, 1 attention, 2 row ops, 3 Macenko, 4 pooling/top-k).  summary() synchronises, fills
 * host arrays ms[c] / work[c] (algorithmic FLOPs for 0-1, bytes otherwise) / count[c], clears. */
void stamp_b200_profile_enable(int on);
int stamp_b200_profile_summary(double* host_ms, double* host_work, long long* host_count, int ncat);

/* ---------------------------------------------------------------------------------------------
 * Dense layers: C[M,N] = epilogue(A[M,K] . W[N,K]^T), tcgen05/TMEM, TMA-fed.
 * replaces: every nn.Linear / Conv2d(patch) on the path --
 *   timm ViT blocks called from src/stamp/preprocessing/__init__.py:325;
 *   src/stamp/modeling/models/vision_tranformer.py:137-139 (per-head Q/K/V, packed to one GEMM),
 *   :153 (fc), :163-167 (feed_forward), :314-318 (project_features);
 *   src/stamp/encoding/encoder/chief.py:45,257-266 (fc + gated attention branches).
 * act:   0 none, 1 GELU(erf), 2 ReLU                         applied to acc + bias
 * store: 0 out16[row,n] = v          1 out32[row,n] = v (+ table[m % gin, n])
 *        2 out32[row,n] += gamma[n]*v (residual + LayerScale; gamma NULL = 1)
 *        3 out16[row,n/2] = silu(v[n]) * v[n+1]   (SwiGLU, weight rows interleaved x1,x2)
 *        4 out16[row,n/2] = tanh(v[n]) * sigmoid(v[n+1])   (gated attention)
 * row remap (prefix tokens): row = (m / gin) * gout + goff + m % gin; gin = 0 -> row = m.
 * dtype: operand type 0 fp16, 1 bf16, 2 fp32 consumed as TF32 (kind::tf32; pre-round operands
 *        to TF32 to avoid the hardware's truncation); 16-bit outputs are fp16 (bf16 if dtype 1).
 * lda/ldw/ldo/ldt in elements; A, W 16-byte aligned with 16-byte row pitch.
 * ------------------------------------------------------------------------------------------- */
/* tile-shape override for tests / tuning: 0 auto, 1 single-CTA tiles only, 2 CTA-pair
 * (tcgen05 cta_group::2, 256x256) tiles whenever legal */
void stamp_b200_gemm_force_mode(int mode);
/* bit 0 (default 1): unmasked head_dim-64 attention runs on the tcgen05 kernels (<= 256 tokens:
 * single-pass ViT kernel; longer: two-pass long-bag kernel, incl. the training forward / backward);
 * 0: always the general (mma.sync) kernels.
 * bit 1: prefer the persistent variant of the ViT kernel (tests / A-B timing).
 * bit 2: long-bag forward on the older two-pass kernel instead of the single-pass one (A-B timing).
 * bit 3: single-pass kernel rescales its accumulator whenever the row maximum grows (tests of that path).
 * bit 4: single-pass kernel with 128-key tiles and one CTA per SM instead of 64-key tiles and two (A-B timing). */
void stamp_b200_attention_tc_enable(int on);

int stamp_gemm_tn(const void* A, long long lda, const void* W, long long ldw, void* out,
                  long long ldo, int M, int N, int K, const float* bias, const float* gamma,
                  int act, int store, int dtype, const float* table, long long ldt, int gin,
                  int gout, int goff, void* stream);

/* LayerNorm over the last dim of an fp32 matrix; out_kind 0 fp16, 1 bf16, 2 fp32.
 * out_lo (fp16 output only, may be NULL) receives fp16(y - fp16(y)): the low half of a
 * split-precision GEMM operand (same leading dimension as out).
 * replaces: timm Block.norm1/norm2/norm (eps 1e-6); nn.LayerNorm in
 *   src/stamp/modeling/models/vision_tranformer.py:161,186,277 (eps 1e-5). */
int stamp_layernorm(const float* x, long long ldx, const float* weight, const float* bias,
                    void* out, void* out_lo, long long ldo, int rows, int cols, float eps,
                    int out_kind, void* stream);

/* x[g*rows_per_group + row_off + r, :] = src[r, :] (+ add[r, :]);  class / register token rows.
 * replaces: timm VisionTransformer._pos_embed (cls/reg token concat);
 *   src/stamp/modeling/models/vision_tranformer.py:347-348 (class token concat). */
int stamp_fill_rows(float* x, long long ldx, int groups, int rows_per_group, int row_off,
                    const float* src, long long lds, const float* add, long long lda, int nrows,
                    int cols, void* stream);

/* uint8 HWC tiles [B,img,img,3] -> 16-bit patch matrix [B*(img/P)^2, Kpad], columns (c,ky,kx),
 * values ((x/255) - mean[c]) / std[c]; columns >= 3*P*P are zero.  host_mean/host_std: 3 floats.
 * replaces: Extractor.transform (ToTensor + Normalize) applied per tile at
 *   src/stamp/preprocessing/__init__.py:94 and the H2D of fp32 tiles at :325, plus the im2col
 *   of timm PatchEmbed's Conv2d. */
int stamp_tiles_to_patches(const uint8_t* tiles, void* patches, int B, int img, int P, int Kpad,
                           const float* host_mean, const float* host_std, int bf16, void* stream);

/* Fused attention forward, fp16 in/out, O(S) memory.
 *   q,k,v: element pointers to (bag 0, token 0, head 0); head h at +h*head_dim;
 *   coords NULL  -> O = softmax(QK^T*scale) V                       (timm Attention / SDPA)
 *   coords given -> O = (softmax(QK^T*scale) - slope[h]*Dist) V     (reference ALiBi)
 *   mask [B,S] (1 = masked) with mask_mode 1 reproduces the reference's masked branch
 *   (attn_mask = m_q&m_k | (q>=1 & k==0) applied after the softmax, no ALiBi on row/col 0);
 *   mask_mode 2 = -inf before the softmax (nn.MultiheadAttention semantics).
 * replaces: src/stamp/modeling/models/vision_tranformer.py:42-74,123-154,218-228,354-379
 *   and timm Attention.forward's F.scaled_dot_product_attention.
 * out_f32: 0 -> fp16 output; 1 -> fp32 output rounded to TF32 (the ALiBi term of the reference is
 *   unscaled distances, |O| easily exceeds the fp16 range; it then feeds a dtype-2 GEMM).
 * dscale: [B,2] scratch filled by stamp_alibi_dist_scale (power-of-two range scale per bag). */
int stamp_attention_fwd(const void* q, const void* k, const void* v, long long row_stride,
                        long long batch_stride, void* out, long long out_row_stride,
                        long long out_batch_stride, int out_f32, int B, int S, int H,
                        int head_dim, float scale, const float* coords, const float* slope,
                        const float* dscale, const uint8_t* mask, int mask_mode, void* stream);
int stamp_alibi_dist_scale(const float* coords, const float* slope, int B, int S, int H,
                           float* dscale, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Tile-encoder ViT forward for one batch of uint8 tiles (UNI ViT-L/16, Virchow2 ViT-H/14, ...).
 * replaces: `model(tiles.to(device))` at src/stamp/preprocessing/__init__.py:322-327 == timm
 *   VisionTransformer.forward as built in src/stamp/preprocessing/extractor/uni.py:26-31 and
 *   virchow2.py:24-42 (incl. the wrapper's [:, 0]) together with Extractor.transform's
 *   ToTensor+Normalize.  Output: fp16 class-token features [B, dim] (the reference's `.half()`).
 * All structs are HOST structs holding DEVICE pointers; 16-bit weights are fp16, K-major [out,in].
 * ------------------------------------------------------------------------------------------- */
typedef struct {
    int img, patch, dim, depth, heads;
    int mlp_hidden;   /* fc1 output width (x1|x2 packed width for SwiGLU) */
    int mlp_kind;     /* 0: fc1-GELU(erf)-fc2   1: SwiGLUPacked (fc1 rows interleaved x1,x2) */
    int reg_tokens;
    int kpad;         /* patch vector length 3*patch^2 rounded up to a multiple of 8 */
    float ln_eps;
    float mean[3], std[3];
    int pool;         /* 0: class token (feats [B, dim]); 1: class token | mean of all other tokens after the final norm
                         (feats [B, 2*dim]; VirchowConcatenated, src/stamp/preprocessing/extractor/virchow_full.py:24-35) */
} StampVitConfig;

typedef struct {
    const void* patch_w;   /* fp16 [dim, kpad], Conv2d weight flattened (c,ky,kx), zero padded */
    const float* patch_b;  /* [dim] */
    const float* prefix;   /* [1+reg_tokens, dim]: cat(cls_token, reg_token) + pos_embed[:1+reg] */
    const float* pos;      /* [n_patches, dim]: pos_embed rows of the patch tokens */
    const float* norm_w;   /* final LayerNorm */
    const float* norm_b;
} StampVitWeights;

typedef struct {
    const float *ln1_w, *ln1_b;
    const void* qkv_w;  const float* qkv_b;   /* [3*dim, dim] */
    const void* proj_w; const float* proj_b;  /* [dim, dim] */
    const float* ls1;                         /* LayerScale gamma or NULL */
    const float *ln2_w, *ln2_b;
    const void* fc1_w;  const float* fc1_b;   /* [mlp_hidden, dim] */
    const void* fc2_w;  const float* fc2_b;   /* [dim, mlp_hidden (/2 for SwiGLU)] */
    const float* ls2;
} StampVitBlock;

/* bytes of 256-byte-aligned device workspace stamp_vit_forward needs for a batch of B tiles
 * (0 = invalid config) */
size_t stamp_vit_workspace_bytes(const StampVitConfig* cfg, int B);

int stamp_vit_forward(const StampVitConfig* cfg, const StampVitWeights* w,
                      const StampVitBlock* blocks /* [depth] */, const uint8_t* tiles /* [B,img,img,3] */,
                      void* feats16 /* [B, dim] (pool 0) or [B, 2*dim] (pool 1) */, int B, void* workspace, size_t workspace_bytes,
                      void* stream);

/* ---------------------------------------------------------------------------------------------
 * ALiBi Transformer-MIL aggregator forward (inference) for a batch of feature bags.
 * replaces: VisionTransformer.forward, src/stamp/modeling/models/vision_tranformer.py:332-384 and
 *   everything it calls (:15-295), as used by LitTileClassifier.predict_step/validation_step
 *   (src/stamp/modeling/models/__init__.py:288-313), deploy._predict (src/stamp/modeling/deploy.py:390-456)
 *   and heatmaps_ (src/stamp/heatmaps/__init__.py:392,419).
 * bags [B,N,F] fp32 (bags_f16 = 0) or fp16 as the feature files store them (bags_f16 = 1, 16-byte aligned: no
 * up-cast round trip), coords fp32 [B,N,2], mask uint8 [B,N] (1 = masked tile) or NULL,
 * logits fp32 [B,C].  mask == NULL takes the reference's mask=None branch (padding tiles attend and
 * are attended, ALiBi applies to the class token's (0,0) coordinate); a mask reproduces :359-379.
 * HOST structs holding DEVICE pointers.
 * ------------------------------------------------------------------------------------------- */
typedef struct {
    int dim_input, dim_model, n_layers, n_heads, dim_ff, dim_output;
    int use_alibi;   /* 1: MultiHeadALiBi, 0: nn.MultiheadAttention */
    /* Zero-padded shapes (0 = not padded).  The kernels want dim_input / dim_model / dim_ff in multiples of 8 and
     * heads of 32 or 64 (80) columns; a model outside that envelope (the reference's own unit tests use heads of 33
     * and 34, src/../tests/test_model.py) is run with zero-padded weights: dim_model = n_heads * padded head width
     * (residual stream = real channels followed by zeros), and these two fields carry what the padding must not
     * change -- the LayerNorm statistics and the softmax scale. */
    int dim_model_real;   /* LayerNorm normalises over the first dim_model_real channels */
    int head_dim_real;    /* softmax scale = head_dim_real ^ -0.5 */
} StampMilConfig;

typedef struct {
    const void* proj_w;          /* fp16 [dim_model, dim_input]   project_features.0.weight */
    const float* proj_b;
    const float* class_token;    /* [dim_model] */
    const float* norm_w;         /* transformer.norm */
    const float* norm_b;
    const float* head_w;         /* fp32 [dim_output, dim_model]  mlp_head.0 */
    const float* head_b;
} StampMilWeights;

typedef struct {
    const float *ln1_w, *ln1_b;              /* layers.l.0.norm */
    const void* qkv_w; const float* qkv_b;   /* fp16 [3d, d]: rows = q heads | k heads | v heads */
    const void* v_w3;                        /* ALiBi: fp16 [d, 3d] = [Wv_hi | Wv_hi | Wv_lo], Wv_lo =
                                                fp16(Wv - fp16(Wv)): K-concatenated split-precision V
                                                projection against LayerNorm output [hi | lo]; else NULL */
    const float* slope;                      /* [H] bias_scale_h / running_mean_h (ALiBi only) */
    const void* fc_w;  const float* fc_b;    /* ALiBi: fp32 [d, 3d] = [W_hi | W_hi | W_lo] rounded to TF32
                                                (mhsa.fc, 3 x TF32 against the attention output [hi | lo]);
                                                else fp16 [d, d] (mhsa.out_proj) */
    const float *ln2_w, *ln2_b;              /* layers.l.1.0 */
    const void* ff1_w; const float* ff1_b;   /* fp16 [ff, d]  layers.l.1.1 */
    const void* ff2_w; const float* ff2_b;   /* fp16 [d, ff]  layers.l.1.4 */
} StampMilLayer;

size_t stamp_mil_workspace_bytes(const StampMilConfig* cfg, int B, int N);

int stamp_mil_forward(const StampMilConfig* cfg, const StampMilWeights* w,
                      const StampMilLayer* layers /* [n_layers] */, const void* bags, int bags_f16,
                      const float* coords, const uint8_t* mask, float* logits, int B, int N,
                      void* workspace, size_t workspace_bytes, void* stream);

/* Ragged batches of the same forward (deploy over bags of different lengths in ONE pass: the dense layers see all
 * rows of all bags, the long-bag attention kernel walks every bag on its own): the bags are concatenated along the
 * token axis WITH their class-token row, bag b = rows seq_off[b] .. seq_off[b+1] (S_b = N_b + 1 tokens).
 *   tokens_f16 fp16 [total_rows, dim_input]: row seq_off[b] is a placeholder (any finite values), the tiles follow;
 *   coords_s   fp32 [total_rows, 2]: (0, 0) in the class-token rows (vision_tranformer.py:349-351); NULL without ALiBi;
 *   seq_off    int32 [B + 1] ON THE DEVICE; S_max = the longest bag in tokens; logits fp32 [B, dim_output].
 * Unmasked forwards with head dimension 64 only (STAMP_ERR_UNSUPPORTED otherwise: call stamp_mil_forward per bag).
 * Same arithmetic per bag as stamp_mil_forward (rows and bags are independent), hence the same results.
 * replaces: the per-patient loop of _predict, src/stamp/modeling/deploy.py:390-456. */
size_t stamp_mil_ragged_workspace_bytes(const StampMilConfig* cfg, int B, int total_rows, int S_max);
int stamp_mil_forward_ragged(const StampMilConfig* cfg, const StampMilWeights* w, const StampMilLayer* layers,
                             const void* tokens_f16, const float* coords_s, const int* seq_off, int B, int total_rows,
                             int S_max, float* logits, void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * ALiBi Transformer-MIL training step (forward with checkpoints, backward, loss, optimizer).
 * replaces: LitTileClassifier.training_step -> _step (src/stamp/modeling/models/__init__.py:239-286:
 *   logits = model(bags, coords=coords, mask=None); F.cross_entropy(logits, soft targets, class weights)),
 *   loss.backward() through VisionTransformer.forward (vision_tranformer.py:332-384, incl. nn.Dropout of
 *   project_features :314-318 and feed_forward :157-169, dropout 0.5 hard-wired :160), the training-mode
 *   _RunningMeanScaler statistic (:23-31) and optim.AdamW.step (models/__init__.py:133-141).
 * mask=None branch only (the branch every Lightning step takes); use_alibi=1 (MultiHeadALiBi) or 0
 * (nn.MultiheadAttention without attention dropout: qkv_w = in_proj_weight, fc = out_proj, bias_scale and
 * inv_rm unused).  bf16 tensor-core
 * operands, fp32 accumulation / residual stream / master parameters / gradients.
 * The two structs below hold fp32 DEVICE pointers in the reference's layouts; the same struct types
 * carry the gradients, which the backward ACCUMULATES into (zero them first, like optimizer.zero_grad()).
 * ------------------------------------------------------------------------------------------- */
typedef struct {
    float* proj_w;        /* [dim_model, dim_input]  project_features.0.weight */
    float* proj_b;
    float* class_token;   /* [dim_model] */
    float* norm_w;        /* transformer.norm */
    float* norm_b;
    float* head_w;        /* [dim_output, dim_model]  mlp_head.0 */
    float* head_b;
} StampMilTrainTop;

typedef struct {
    float *ln1_w, *ln1_b;                 /* layers.l.0.norm */
    float *qkv_w, *qkv_b;                 /* [3d, d] / [3d]: rows = query heads | key heads | value heads */
    float* bias_scale;                    /* [H]  mhsa.attentions.h.bias_scale */
    float *fc_w, *fc_b;                   /* mhsa.fc */
    float *ln2_w, *ln2_b;                 /* layers.l.1.0 */
    float *ff1_w, *ff1_b, *ff2_w, *ff2_b; /* layers.l.1.1 / layers.l.1.4 */
} StampMilTrainLayer;

typedef struct {
    float p_drop_proj;          /* project_features Dropout(p)            (0 = off) */
    float p_drop_ff;            /* both feed_forward Dropouts, reference 0.5 (0 = off) */
    unsigned long long seed;    /* dropout masks = f(seed, site, element): regenerated in the backward */
    const float* inv_rm;        /* DEVICE [n_layers, H]: 1 / scale_distance.running_mean (after its update) */
} StampMilTrainStep;

/* bytes of 256-byte-aligned device memory holding checkpoints + scratch between forward and backward
 * (0 = unsupported configuration: needs head dim 32/64, dims % 8 == 0, dim_model <= 1024) */
size_t stamp_mil_train_ctx_bytes(const StampMilConfig* cfg, int B, int N);
int stamp_mil_train_forward(const StampMilConfig* cfg, const StampMilTrainTop* params,
                            const StampMilTrainLayer* layers, const StampMilTrainStep* step,
                            const float* bags /* [B,N,F] */, const float* coords /* [B,N,2] */,
                            float* logits /* [B,C] */, int B, int N, void* ctx, size_t ctx_bytes, void* stream);
/* dbags: NULL or fp32 [B,N,F] (overwritten) -- the gradient heatmaps need (heatmaps/__init__.py:36-56) */
int stamp_mil_train_backward(const StampMilConfig* cfg, const StampMilTrainTop* params,
                             const StampMilTrainLayer* layers, const StampMilTrainStep* step,
                             const float* dlogits /* [B,C] */, StampMilTrainTop* grads,
                             StampMilTrainLayer* layer_grads, float* dbags, int B, int N, void* ctx,
                             size_t ctx_bytes, void* stream);
/* keep-mask (1 = kept) of dropout site `site` for element indices 0..n-1: site 0 = project_features
 * ([B*N, d] row-major), 1 + 2l = feed_forward hidden of layer l ([M, ff]), 2 + 2l = its output ([M, d]) */
int stamp_mil_train_dropout_mask(unsigned long long seed, int site, long long n, float p, uint8_t* keep_out,
                                 void* stream);
/* mean of all pairwise token distances of a batch of bags, class token at (0,0) included: the value
 * torch.cdist(...).mean() feeds the running mean in training mode (vision_tranformer.py:26-29) */
size_t stamp_pairwise_dist_mean_workspace_bytes(int B, int N);
int stamp_pairwise_dist_mean(const float* coords, int B, int N, float* mean_out /* device scalar */,
                             void* workspace, size_t workspace_bytes, void* stream);
/* loss = 1/B sum_b sum_c -w_c y_bc log softmax(logits_b)_c; dlogits_out (may be NULL) = grad_scale * dloss/dlogits */
int stamp_cross_entropy(const float* logits, const float* targets, const float* class_weights /* or NULL */,
                        int B, int C, float grad_scale, float* loss_out /* device scalar */,
                        float* dlogits_out, void* stream);
/* Losses of the regression and survival tasks with their gradients, one launch each, no host synchronisation.
 * replaces: nn.functional.l1_loss in LitBaseRegressor._compute_loss, src/stamp/modeling/models/__init__.py:420-422, and
 *   neg_partial_log_likelihood(log_hz, time, event), src/stamp/modeling/models/cox.py:107-268 (ties: Efron by default,
 *   Breslow when `breslow`; reduction "mean": over the distinct event times for Efron / no ties, over the events for
 *   Breslow).  No events: loss 0, gradient 0.  n <= STAMP_COX_MAX_SAMPLES (one CTA, O(n^2) compares in shared memory).
 *   d*_out (may be NULL) = grad_scale * dloss/dinput. */
#define STAMP_COX_MAX_SAMPLES 8192
int stamp_cox_loss(const float* log_hz, const float* time, const uint8_t* event, int n, int breslow, float grad_scale,
                   float* loss_out /* device scalar */, float* dlog_hz_out, void* stream);
int stamp_l1_loss(const float* pred, const float* target, long long n, float grad_scale, float* loss_out /* device scalar */,
                  float* dpred_out, void* stream);
/* torch.optim.AdamW step t (1-based) over flat fp32 buffers; grads are multiplied by grad_scale first */
int stamp_adamw_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, long long n,
                     float lr, float beta1, float beta2, float eps, float weight_decay, int step,
                     float grad_scale, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Macenko stain normalisation over a batch of uint8 RGB tiles (in/out [n_tiles,H,W,3]).
 * The reference snapshot has no Macenko code (README.md:35 only): the stage is named by
 * BASELINE.json's north_star; algorithm as specified in SURVEY.md 8c (oracle/macenko_oracle.py).
 * One stain matrix is fitted per group of `tiles_per_fit` consecutive tiles (<= 0: one fit over the
 * whole batch) and applied per pixel.  Optional outputs: he_out [G,3,2] (columns H, E),
 * maxc_out [G,2], valid_out [G] (0 = fewer than 16 tissue pixels: tiles passed through).
 * H*W*3 must be a multiple of 48; in/out 16-byte aligned; workspace 256-byte aligned; at most 65535 fit groups
 * per call.  Seven launches, all asynchronous on `stream`.
 * ------------------------------------------------------------------------------------------- */
size_t stamp_macenko_workspace_bytes(int n_tiles, int tiles_per_fit);
int stamp_macenko_u8(const uint8_t* in, uint8_t* out, int n_tiles, int H, int W, int tiles_per_fit,
                     float Io, float alpha, float beta, float* he_out, float* maxc_out,
                     int* valid_out, void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Tissue-texture filter of the tiling stage: per tile the number of Canny edge pixels of its grayscale
 * image, bit-exact with the reference's Pillow / OpenCV calls.
 * replaces: _has_enough_texture, src/stamp/preprocessing/tiling.py:279-291
 *   (tile.convert("L") -> cv2.Canny(gray, 40, 100) -> edges.mean() / 255 >= cutoff).
 * tiles uint8 [n_tiles, H, W, 3] RGB; edge_count int32 [n_tiles]; edges_out NULL or uint8 [n_tiles, H, W]
 * (0 / 255, cv2.Canny's output).  One tile must fit one SM's shared memory (224 x 224 tiles use 200 KB; up to about 240 x 224 pixels).
 * ------------------------------------------------------------------------------------------- */
int stamp_tile_texture_u8(const uint8_t* tiles, int n_tiles, int H, int W, int low, int high,
                          int* edge_count, uint8_t* edges_out, void* stream);

/* ---------------------------------------------------------------------------------------------
 * fp32 building blocks of the TransMIL aggregator (inference): the small-matrix side of its Nystrom attention and its
 * position layer; the dense layers around them use stamp_gemm_tn.
 * replaces: NystromAttention.forward / moore_penrose_iter_pinv, src/stamp/modeling/models/trans_mil.py:25-160, and
 *   PPEG.forward :253-273.  Token-major fp32 matrices [rows, ld]; head h = columns h*64 .. h*64+63.
 *   stamp_landmark_mean_f32   out[j, c] = mean of `group` consecutive rows of x             (:113-127, reduce "(n l) d -> n d")
 *   stamp_sgemm_batched_f32   C[b] (+)= alpha * A[b] @ op(B[b]) + bias; op = B^T (trans_b), B, or (eye_minus_b * I - B) for the
 *                             Newton-Schulz steps z <- 1/4 z (13 I - xz (15 I - xz (7 I - xz)))   (:25-42)
 *   stamp_softmax_rows_f32    in-place row softmax of [batch][rows, cols]
 *   stamp_pinv_init_f32       z0 = x^T / (max row abs-sum * max column abs-sum over the WHOLE batch); scratch2: 2 uint32
 *   stamp_attention_f32       O = softmax(scale * Q K^T) V per 64-wide head, any nq / nk (attn1 @ W and attn3 @ v); splits > 1
 *                             shares the keys among `splits` CTAs per query block (few landmark queries over many keys) through
 *                             scratch [splits * heads * nq * 66] floats, merged by a second kernel; splits = 1: scratch may be NULL
 *   stamp_dwconv1d_add_f32    out += depth-wise convolution of v along the tokens, one `taps`-vector per head  (res_conv, :82-90,153)
 *   stamp_dwconv2d_f32        depth-wise ksize x ksize convolution of tokens laid out on an H x W grid, zero padded  (PPEG)
 * ------------------------------------------------------------------------------------------- */
int stamp_landmark_mean_f32(const float* x, long long ldx, int group, float* out, long long ldo, int n_landmarks, int cols,
                            void* stream);
int stamp_sgemm_batched_f32(const float* A, long long lda, long long stride_a, const float* B, long long ldb, long long stride_b,
                            float* C, long long ldc, long long stride_c, int M, int N, int K, int batch, int trans_b, float alpha,
                            float eye_minus_b, const float* bias /* [N] or NULL */, int mode /* bit 0: C +=, bit 1: ReLU */, void* stream);
int stamp_softmax_rows_f32(float* x, long long ld, long long stride, int rows, int cols, int batch, void* stream);
int stamp_pinv_init_f32(const float* x, float* z, int n, int batch, unsigned int* scratch2, void* stream);
int stamp_attention_f32(const float* Q, long long ldq, const float* K, long long ldk, const float* V, long long ldv, float* O,
                        long long ldo, int nq, int nk, int heads, float scale, float* scratch, int splits, void* stream);
int stamp_dwconv1d_add_f32(const float* v, long long ldv, const float* w, float* out, long long ldo, int n, int heads, int taps,
                           void* stream);
int stamp_dwconv2d_f32(const float* in, long long ldi, const float* k, const float* bias, float* out, long long ldo, int H, int W,
                       int C, int ksize, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Mean over the tiles of a bag: the pooling step of the MLP / Linear aggregators (the layers after it are
 * stamp_sgemm_batched_f32 calls on [B, F] rows).
 * replaces: `x.mean(dim=1)` in MLP.forward / Linear.forward, src/stamp/modeling/models/mlp.py:40-43, :57-60.
 *   x        [B][n_tiles, ldx] fp32 or fp16 (is_half), bag b at x + b * stride_bag (elements)
 *   out      [B, ldo] fp32 means
 *   splits   CTAs sharing the rows of one bag (stamp_bag_mean_splits' suggestion fills 148 SMs); scratch holds
 *            splits * B * F floats when splits > 1 (two deterministic stages, no atomics), may be NULL otherwise
 * HBM-bound: B * n_tiles * F * sizeof(element) bytes read once.
 * ------------------------------------------------------------------------------------------- */
int stamp_bag_mean_splits(int B, int n_tiles, int F, int is_half);
int stamp_bag_mean(const void* x, int is_half, long long ldx, long long stride_bag, int B, int n_tiles, int F, float* out,
                   long long ldo, float* scratch, int splits, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Decode of cached JPEG tiles, bit-exact with Pillow (libjpeg-turbo defaults: islow inverse DCT, fancy chroma
 * up-sampling, fixed-point YCbCr -> RGB).
 * replaces: Image.open(tile_fp) + img.load() per cached tile in _tiles_from_cache_file,
 *   src/stamp/preprocessing/tiling.py:380-406.
 * Baseline / extended sequential Huffman JPEG, 8 bit, three components in one interleaved scan, 4:2:0 or 4:4:4
 * (what Pillow's Image.save(format="jpeg") writes); anything else returns STAMP_ERR_UNSUPPORTED.
 * Host side (no GPU work, re-entrant, callable from many threads):
 *   stamp_jpeg_read_header     geometry and quantisation tables of one file;
 *   stamp_jpeg_coef_count      int16 coefficients per tile (every block of every component, padded to whole MCUs);
 *   stamp_jpeg_entropy_decode  Huffman decode of one file into quantised coefficients, component-major
 *                              [Y blocks | Cb blocks | Cr blocks][64] in natural (row-major) order, and its
 *                              quantisation tables uint16 [3][64]; `expect` (may be NULL) pins the geometry.
 * Device side:
 *   stamp_jpeg_decode_coefs_u8 coef int16 [n_tiles][coef_count], quant uint16 [n_tiles][3][64] (both on the device,
 *                              16-byte aligned) -> out uint8 [n_tiles, height, width, 3]; workspace holds the
 *                              component planes (stamp_jpeg_workspace_bytes).
 * ------------------------------------------------------------------------------------------- */
typedef struct StampJpegInfo {
    int width, height, n_comp;
    int h[3], v[3];              /* sampling factors per component */
    int mcus_x, mcus_y;          /* MCUs per row / column */
    uint16_t quant[3][64];       /* per component, natural order */
} StampJpegInfo;
int stamp_jpeg_read_header(const uint8_t* data, size_t n, StampJpegInfo* info);
size_t stamp_jpeg_coef_count(const StampJpegInfo* info);
int stamp_jpeg_entropy_decode(const uint8_t* data, size_t n, const StampJpegInfo* expect, int16_t* coef,
                              uint16_t* quant);
size_t stamp_jpeg_workspace_bytes(const StampJpegInfo* info, int n_tiles);
int stamp_jpeg_decode_coefs_u8(const StampJpegInfo* info, const int16_t* coef, const uint16_t* quant, int n_tiles,
                               uint8_t* out, void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Resampling of uint8 RGB tiles (extractor transforms that resize the tile before the model), bit-exact with
 * Pillow's Image.resize for 8-bit images.
 * replaces: transforms.Resize(256, interpolation=BICUBIC) + transforms.CenterCrop(224),
 *   src/stamp/preprocessing/extractor/gigapath.py:20-27 (torchvision -> PIL Image.resize -> Resample.c:
 *   horizontal then vertical pass, 22-bit fixed-point coefficients, intermediate image rounded to uint8).
 * tiles uint8 [n_tiles, Hin, Win, 3]; out uint8 [n_tiles, Hc, Wc, 3] = rows crop_y .. crop_y+Hc and columns
 * crop_x .. crop_x+Wc of the resampled image.  coef_* int32 [out_size, ksize] (coefficients scaled by 2^22),
 * bounds_* int32 [out_size, 2] (first input index, tap count), both on the device, as Pillow's precompute_coeffs /
 * normalize_coeffs_8bpc produce them (stamp_b200/resize.py computes them in double).  rows_per_strip output rows
 * per CTA; max_in_rows = the most input rows any strip needs.  stamp_resize_u8_smem_bytes: dynamic shared memory
 * of the launch (0 for invalid arguments; must not exceed 227 KB).
 * ------------------------------------------------------------------------------------------- */
size_t stamp_resize_u8_smem_bytes(int Win, int Wc, int ksize_x, int max_in_rows);
int stamp_resize_u8(const uint8_t* tiles, int n_tiles, int Hin, int Win, uint8_t* out, int Hc, int Wc,
                    int crop_y, int crop_x, const int* coef_x, const int* bounds_x, int ksize_x,
                    const int* coef_y, const int* bounds_y, int ksize_y, int rows_per_strip, int max_in_rows,
                    void* stream);

/* ---------------------------------------------------------------------------------------------
 * Slide-level pooling.
 * stamp_gated_attn_pool -- CHIEF's gated-attention MIL pooling:
 *   h = ReLU(W1 x + b1); A_raw = Wc (tanh(Wa h + ba) * sigmoid(Wb h + bb)) + bc;
 *   pooled = softmax_over_tiles(A_raw) @ x           (x: the ORIGINAL fp32 features [N, D])
 *   replaces CHIEFModel.forward / Attn_Net_Gated.forward, src/stamp/encoding/encoder/chief.py:74-89,
 *   :255-275 (size 'small': D 768, L 512, Dh 256).  The score chain runs in split precision
 *   (hi/lo fp16 operand pairs, fp32 accumulate) so that top-k over attn_raw matches fp32.
 * stamp_topk_f32 -- exact top-k (k <= 1024) of fp32 scores, largest (or smallest) first, ties by
 *   lower index; replaces torch.topk at src/stamp/encoding/encoder/eagle.py:108-109 and
 *   src/stamp/heatmaps/__init__.py:216-229.
 * stamp_gather_mean_f32 -- mean of k gathered rows (EAGLE embedding, eagle.py:112-118).
 * ------------------------------------------------------------------------------------------- */
typedef struct {
    const void* fc_w_hi; const void* fc_w_lo; const float* fc_b;   /* fp16 [L, D] hi / lo; [L] */
    const void* ab_w_hi; const void* ab_w_lo; const float* ab_b;   /* fp16 [2*Dh, L], rows a_0,b_0,a_1,b_1,..; [2*Dh] */
    const float* c_w;                                              /* [Dh] attention_c.weight */
    float c_b;
} StampGatedAttnWeights;

size_t stamp_gated_attn_pool_workspace_bytes(int N, int D, int L, int Dh);
int stamp_gated_attn_pool(const StampGatedAttnWeights* w, const float* x, int N, int D, int L, int Dh,
                          float* attn_raw /* [N] */, float* pooled /* [D] */, void* workspace,
                          size_t workspace_bytes, void* stream);
int stamp_topk_f32(const float* scores, int N, int k, int largest, long long* idx_out,
                   float* val_out /* may be NULL */, void* stream);
int stamp_gather_mean_f32(const float* feats, const long long* idx, int k, int D, float* out,
                          void* stream);

#ifdef __cplusplus
}
#endif
#endif /* STAMP_B200_H */
