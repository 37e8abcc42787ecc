// Host-visible declarations of the tcgen05 GEMM (gemm.cu).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace sb {

// activation applied to (acc + bias) before the store stage
enum GemmAct : int { ACT_NONE = 0, ACT_GELU = 1, ACT_RELU = 2 };

// store stage
enum GemmStore : int {
    ST_16 = 0,       // out16[row, n] = v
    ST_32 = 1,       // out32[row, n] = v
    ST_RESID32 = 2,  // out32[row, n] += gamma[n] * v           (gamma == null -> 1)
    ST_SWIGLU16 = 3, // out16[row, n/2] = silu(v[n]) * v[n+1]   (n even; weights row-interleaved)
    ST_GATED16 = 4,  // out16[row, n/2] = tanh(v[n]) * sigmoid(v[n+1])
};

struct GemmParams {
    int M, N, K;
    int act;
    int store;
    int bf16;            // 16-bit operands and 16-bit outputs are bf16 instead of fp16
    int tf32;            // operands are fp32 read as TF32 (kind::tf32); 16-bit outputs follow bf16
    void* out;           // fp16/bf16 or fp32 matrix
    void* out_lo;        // fp16 16-bit stores only, may be null: fp16(v - fp16(v)), same ldo (full tiles)
    long long ldo;       // leading dimension of out, elements
    const float* bias;   // [N] or null
    const float* gamma;  // [N] or null
    const float* table;  // [gin, ldt] fp32 addend applied before the activation, or null
    long long ldt;
    // optional row remap (token layouts with prefix rows):
    //   row = (m / gin) * gout + goff + (m % gin);   gin == 0 -> row = m
    int gin, gout, goff;
    // K-concatenated split-precision product: the A operand physically holds a_kwrap columns
    // ([hi | lo]) and is read at k % a_kwrap, i.e. hi, lo, hi against W' = [W_hi | W_hi | W_lo]; 0 = off
    int a_kwrap;
};

// C[M,N] = A[M,K] . B[N,K]^T, A and B K-major 16-bit, fp32 accumulate in TMEM.
// lda/ldb in elements with a 16-byte row pitch; pointers 16-byte aligned.
int gemm_tn(const void* A, long long lda, const void* B, long long ldb, const GemmParams& p,
            cudaStream_t stream);

// number of kernels the last gemm_tn call launched (always 1); kept for launch accounting
int gemm_num_sms();

}  // namespace sb

namespace sb {
// 0 auto, 1 single-CTA tiles only, 2 CTA-pair (cta_group::2) tiles whenever legal
void gemm_force_mode(int mode);
}  // namespace sb

namespace sb {
// fp16 [batches][rows][inner] tensor map, 128B swizzle, zero fill outside `rows` (gemm.cu)
int make_tmap_3d_f16(CUtensorMap* tm, const void* ptr, int inner, int rows, int batches,
                     long long row_stride, long long batch_stride, int box_inner, int box_rows);
}  // namespace sb
