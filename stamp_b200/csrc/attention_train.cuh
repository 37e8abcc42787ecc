// ALiBi attention forward-for-training and backward (attention_train.cu); bf16 operands.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace sb {

struct AttnTrainParams {
    // packed projections, bf16: head h of token s of bag b at  b*batch_stride + s*row_stride + h*head_dim
    const uint16_t* q;
    const uint16_t* k;
    const uint16_t* v;
    long long row_stride, batch_stride;
    // forward outputs / backward inputs, [B, S, H*head_dim] with out_row_stride / out_batch_stride
    uint16_t* out;          // bf16  O = (P - beta Dhat) V
    float* osm;             // fp32  P V            (same strides as out)
    float* odv;             // fp32  Dhat V         (same strides; ALiBi only)
    float* lse2;            // [B, H, S] log2-domain log-sum-exp of the scaled logits
    long long out_row_stride, out_batch_stride;
    int B, S, H;
    float scale;            // head_dim^-0.5
    float scale_log2;       // scale * log2(e)
    const float2* coords;   // [B, S] token coordinates, or null: plain softmax attention
    const float* beta;      // [H] bias_scale_h
    const float* inv_rm;    // [H] 1 / running_mean_h
    // optional (long bags, attention_mil_v3.cu): bf16 distance matrix [B, S, ld] scaled by dist_scale[2b] = 2^-e,
    // dist_scale[2b+1] = 2^e; null -> distances are recomputed from coords inside the kernels
    const uint16_t* dist16;
    const float* dist_scale;
    // backward only
    const float* dout32;    // fp32 dO, strides as out (input)
    uint16_t* dout;         // bf16 copy of dO written by the backward's first pass (scratch, strides as out)
    float* delta;           // [B, H, S] scratch: dO . Osm
    uint16_t* dq;           // bf16, strides as q/k/v
    uint16_t* dk;
    uint16_t* dv;
    float* dbeta;           // [H], accumulated with atomicAdd (caller zeroes / accumulates)
};

int attention_train_fwd(const AttnTrainParams& p, int head_dim, cudaStream_t stream);
// tcgen05 two-pass forward for long bags (attention_mil_tc.cu, training variant of the inference kernel);
// SB_ERR_UNSUPPORTED = outside its envelope (head_dim != 64, <= 256 tokens)
int attention_mil_tc_train_fwd(const AttnTrainParams& p, int head_dim, cudaStream_t stream);
// third-generation forward (TMA-fed distance tiles, P through tensor memory); needs p.dist16 for ALiBi
int attention_mil_v3_train_fwd(const AttnTrainParams& p, int head_dim, cudaStream_t stream);
// tcgen05 dK/dV + dQ kernels (attention_train_tc.cu); need p.dout / p.delta filled; SB_ERR_UNSUPPORTED = not applicable
int attention_train_tc_bwd(const AttnTrainParams& p, int head_dim, cudaStream_t stream);
void attention_train_tc_enable(int on);   // tests: 0 forces the mma.sync kernels
int attention_train_bwd(const AttnTrainParams& p, int head_dim, cudaStream_t stream);

}  // namespace sb
