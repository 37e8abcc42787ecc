// tcgen05 weight-gradient kernel (wgrad_tc.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace sb {

// dW[Nout, Kin] += dY[M, Nout]^T . X[M, Kin], bf16 operands in token-major layout, fp32 accumulation into dW.
// SB_ERR_UNSUPPORTED: outside the kernel's envelope (few tokens, unaligned) -> caller uses the mma.sync kernel.
// scratch (may be null): >= wgrad_tc_scratch_bytes(Nout, Kin) of 16-byte aligned device memory for the per-split
// partial tiles; with it the reduction into dW is a second deterministic pass instead of atomics.
size_t wgrad_tc_scratch_bytes(int Nout, int Kin);
int wgrad_tc(const uint16_t* dY, long long ldy, const uint16_t* X, long long ldx, float* dW, long long ldw, int M,
             int Nout, int Kin, float* scratch, size_t scratch_bytes, cudaStream_t stream);
void wgrad_tc_enable(int on);

}  // namespace sb
