// Small MIL kernels shared by the inference (mil.cu) and training (mil_train.cu) sequencing.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace sb {

// coords [B,N,2] -> coords_s [B,N+1] (class token at (0,0)); mask [B,N] -> mask_s [B,N+1] (class token
// unmasked).  Either output may be null.
int mil_prepare(const float* coords, const uint8_t* mask, float2* coords_s, uint8_t* mask_s, int B, int N,
                cudaStream_t stream);

// logits[b, :] = head(LayerNorm(x[b * bag_stride, :]))   fp32, one CTA per bag
int cls_head(const float* x, long long bag_stride, int d, const float* norm_w, const float* norm_b,
             const float* head_w, const float* head_b, int C, int B, float* logits, cudaStream_t stream);

}  // namespace sb
