// Fused attention forward (attention.cu).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace sb {

struct AttnParams {
    const __half* q;  // head 0 of token 0 of bag 0; head h at +h*head_dim
    const __half* k;
    const __half* v;
    long long row_stride;    // elements between consecutive tokens
    long long batch_stride;  // elements between consecutive bags / tiles
    long long v_row_stride;  // same for v (0 -> use row_stride / batch_stride)
    long long v_batch_stride;
    void* out;               // [B, S, H*head_dim], head-major concat; fp16, or fp32 (TF32-rounded) if out_f32
    int out_f32;
    float* out_lo;           // out_f32 only, may be null: tf32(o - tf32(o)), the low half of a split-precision operand
    long long out_row_stride;
    long long out_batch_stride;
    int B, S, H;
    float scale_log2;        // softmax scale * log2(e)
    const float* coords;     // [B, S, 2] fp32 or null (no ALiBi)
    const float* slope;      // [H]   bias_scale_h / running_mean_h
    const float* dscale;     // [B, 2] {2^-e, 2^e} from alibi_dist_scale
    const uint8_t* mask;     // [B, S] 1 = masked token, or null
    int mask_mode;           // 1: reference ALiBi masked branch (post-softmax), 2: -inf before softmax
    // long-bag tcgen05 kernel, third generation (attention_mil_v3.cu): pre-computed 16-bit distance matrix
    // [B, S, ld = S rounded up to 64] scaled by dscale[2b]; slope[h] is then applied in fp32 in the epilogue
    const uint16_t* dist16;
    // ragged batches (attention_mil_v3.cu only): the bags are concatenated along the token axis, bag b owns the rows
    // seq_off[b] .. seq_off[b+1] of q / k / v / out (batch strides unused) and of the distance matrix, which then has
    // ONE row pitch for all bags (dist_ld); S_max = longest bag (grid size); S = total number of rows
    const int* seq_off;      // device, [B + 1], or null
    int S_max;
    long long dist_ld;
    int q_rows;              // 0: every token is a query; n > 0: only the first n tokens' outputs are needed (rows up
                             // to the end of their query tile may still be written)
};

int attention_fwd(const AttnParams& p, int head_dim, cudaStream_t stream);

// tcgen05 path for short unmasked sequences (attention_tc.cu); SB_ERR_UNSUPPORTED = not applicable
int attention_tc_fwd(const AttnParams& p, int head_dim, cudaStream_t stream);
void attention_tc_enable(int on);
void attention_tc_set_trace(long long* device_buf);  // debug: 5 x int64 per CTA (phase cycle counts)
// persistent streaming kernel for ViT tiles (attention_vit_stream.cu), tried before attention_tc_fwd
int attention_vit_stream_fwd(const AttnParams& p, int head_dim, cudaStream_t stream);
void attention_vit_stream_enable(int on);   // bit 0: use it, bit 1: rescale eagerly (tests)
// tcgen05 two-pass kernel for long unmasked bags, plain or ALiBi (attention_mil_tc.cu)
int attention_mil_tc_fwd(const AttnParams& p, int head_dim, cudaStream_t stream);
void attention_mil_tc_enable(int on);

// third-generation long-bag kernel (attention_mil_v3.cu) and its distance-matrix producer
int attention_mil_v3_fwd(const AttnParams& p, int head_dim, cudaStream_t stream);
void attention_mil_v3_enable(int on);   // bit 0: use it, bit 1: rescale eagerly (tests)
size_t mil_dist16_bytes(int B, int S);
// coords_s [B, S, 2] -> scale [B][2] = {2^-e, 2^e} (largest distance of the bag in (8192, 16384]) and
// dist16 [B, S, ld] = |x_q - x_k| * 2^-e as fp16 (bf16 != 0: bfloat16)
size_t mil_dist16_scratch_bytes(int B);
// ragged form: coords_s [total rows, 2], bag b = rows seq_off[b] .. seq_off[b+1]; dist16 [total rows, ld] with
// ld >= S_max rounded up to 64 (zeros past each bag's own length)
int mil_dist16_ragged(const float* coords_s, const int* seq_off, int B, int S_max, long long ld, float* scale,
                      uint16_t* dist16, void* bbox_scratch, cudaStream_t stream);
int mil_dist16(const float* coords_s, int B, int S, int bf16, float* scale, uint16_t* dist16, void* bbox_scratch,
               cudaStream_t stream);

int alibi_dist_scale(const float* coords, const float* slope, int B, int S, int H, float* dscale,
                     cudaStream_t stream);

}  // namespace sb
