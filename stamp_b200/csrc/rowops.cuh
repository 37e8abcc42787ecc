// Row-wise HBM-bound kernels (rowops.cu).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace sb {

// out_kind: 0 fp16, 1 bf16, 2 fp32; out_lo (fp16 only, may be null) receives fp16(y - fp16(y))
int layernorm(const float* x, long long ldx, const float* w, const float* b, void* out, void* out_lo,
              long long ldo, int rows, int cols, float eps, int out_kind, cudaStream_t stream);

// zero-padded widths: statistics over the first cols_real of cols columns (padding columns: zeros in, zeros out)
int layernorm_padded(const float* x, long long ldx, const float* w, const float* b, void* out, void* out_lo,
                     long long ldo, int rows, int cols, int cols_real, float eps, int out_kind, cudaStream_t stream);

// x[g * rows_per_group + row_off + r, :] = src[r, :] (+ add[r, :]) for g < groups, r < nrows
int fill_rows(float* x, long long ldx, int groups, int rows_per_group, int row_off,
              const float* src, long long lds, const float* add, long long lda, int nrows, int cols,
              cudaStream_t stream);

// uint8 HWC tiles -> normalised 16-bit patch matrix [B * (img/P)^2, Kpad], columns (c, ky, kx)
int tiles_to_patches(const uint8_t* tiles, void* patches, int B, int img, int P, int Kpad,
                     const float mean[3], const float stdv[3], int bf16, cudaStream_t stream);

// fp32 -> fp16 (round to nearest); lo (may be null) receives fp16(v - fp16(v))
int cast_f32_f16(const float* in, __half* hi, __half* lo, long long n, cudaStream_t stream);

}  // namespace sb
