// Separable fixed-point resampling of uint8 RGB tiles, bit-exact with Pillow's ImagingResample (8 bits per channel).
//
// reference call site: the extractor transforms that resample the 224 px tile before the model, e.g.
//   transforms.Resize(256, interpolation=BICUBIC) + CenterCrop(224), src/stamp/preprocessing/extractor/gigapath.py:20-27
//   (torchvision hands a PIL image to Image.resize; Pillow's Resample.c then runs a horizontal pass and a vertical
//   pass, each with 22-bit fixed-point coefficients, an accumulator preset to one half and a clip to 0..255).
// The coefficient tables are the host's business (double arithmetic identical to precompute_coeffs /
// normalize_coeffs_8bpc, stamp_b200/resize.py); the kernel is integer multiply-add only, so the result does not
// depend on the GPU's floating point.  Only the crop window of the resampled image is computed.
//
// One CTA per (strip of output rows, tile): the input rows the strip needs are staged in shared memory (coalesced
// 16-byte loads), the horizontal pass writes the intermediate uint8 rows next to them, the vertical pass reads four
// neighbouring bytes per thread and stores 32-bit words.  HBM traffic = the bytes of the tile in + the crop out.

#include <cuda_runtime.h>

#include <cstdint>

#include "common.cuh"
#include "profile.cuh"
#include "stamp_b200.h"

namespace sb {
constexpr int RS_ROW_SLACK = 5;                 // intermediate rows the fixed-tap vertical pass may read past the last one
namespace {

constexpr int RS_THREADS = 256;
constexpr int RS_PRECISION_BITS = 32 - 8 - 2;   // Pillow: PRECISION_BITS

__device__ __forceinline__ int clip8(int ss) {
    ss >>= RS_PRECISION_BITS;                   // arithmetic shift, as in C on the reference's platforms
    return ss < 0 ? 0 : (ss > 255 ? 255 : ss);
}

// KS > 0: both passes run exactly KS taps per output (the tables are zero-padded beyond each output's tap count, the
// staging buffers have KS rows / 32 bytes of slack, so the surplus taps read initialised-or-not bytes times zero);
// KS = 0: tap counts from the bounds table (any filter support).
template <int KS>
__global__ void __launch_bounds__(RS_THREADS)
resize_u8_kernel(const uint8_t* __restrict__ src, int Hin, int Win, uint8_t* __restrict__ dst, int Hc, int Wc,
                 int cy, int cx, const int* __restrict__ kx, const int* __restrict__ bx, int ksx,
                 const int* __restrict__ ky, const int* __restrict__ by, int ksy, int rows_per_strip,
                 int max_in_rows) {
    extern __shared__ __align__(16) uint8_t rs_smem[];
    const int tid = threadIdx.x;
    const int tile = blockIdx.y;
    const int r0 = blockIdx.x * rows_per_strip;
    const int nr = min(rows_per_strip, Hc - r0);
    const int y_first = by[(cy + r0) * 2];
    const int y_last_out = cy + r0 + nr - 1;
    const int nin = by[y_last_out * 2] + by[y_last_out * 2 + 1] - y_first;
    const int in_stride = Win * 3, out_stride = Wc * 3;

    uint8_t* tmp = rs_smem + ((static_cast<size_t>(max_in_rows) * in_stride + 32 + 15) & ~size_t(15));
    int* kxs = reinterpret_cast<int*>(tmp + ((static_cast<size_t>(max_in_rows + RS_ROW_SLACK) * out_stride + 15) & ~size_t(15)));
    int* bxs = kxs + Wc * ksx;

    // ---- stage the input rows (one contiguous byte range of the tile) and the strip's horizontal coefficients
    const uint8_t* g = src + (static_cast<size_t>(tile) * Hin + y_first) * in_stride;
    const int nbytes = nin * in_stride;
    const int mis = static_cast<int>(reinterpret_cast<uintptr_t>(g) & 15);
    uint8_t* s = rs_smem + mis;                  // same 16-byte phase as the source: body copies are aligned on both sides
    const int head = mis ? min(16 - mis, nbytes) : 0;
    for (int i = tid; i < head; i += RS_THREADS) s[i] = g[i];
    const int body = (nbytes - head) >> 4;
    const uint4* g4 = reinterpret_cast<const uint4*>(g + head);
    uint4* s4 = reinterpret_cast<uint4*>(s + head);
    for (int i = tid; i < body; i += RS_THREADS) s4[i] = __ldg(g4 + i);
    for (int i = head + (body << 4) + tid; i < nbytes; i += RS_THREADS) s[i] = g[i];
    for (int i = tid; i < Wc * ksx; i += RS_THREADS) kxs[i] = kx[cx * ksx + i];
    for (int i = tid; i < Wc * 2; i += RS_THREADS) bxs[i] = bx[cx * 2 + i];
    __syncthreads();

    // ---- horizontal pass: thread owns one (output column, channel), walks down the staged rows
    for (int col = tid; col < out_stride; col += RS_THREADS) {
        const int xo = col / 3, c = col - xo * 3;
        const int xmin = bxs[xo * 2], n = bxs[xo * 2 + 1];
        const int* k = kxs + xo * ksx;
        const uint8_t* p = s + xmin * 3 + c;
        if constexpr (KS > 0) {
            int kr[KS];
#pragma unroll
            for (int t = 0; t < KS; ++t) kr[t] = k[t];
            for (int i = 0; i < nin; ++i, p += in_stride) {
                int ss = 1 << (RS_PRECISION_BITS - 1);
#pragma unroll
                for (int t = 0; t < KS; ++t) ss += static_cast<int>(p[t * 3]) * kr[t];
                tmp[i * out_stride + col] = static_cast<uint8_t>(clip8(ss));
            }
        } else {
            for (int i = 0; i < nin; ++i, p += in_stride) {
                int ss = 1 << (RS_PRECISION_BITS - 1);
                for (int t = 0; t < n; ++t) ss += static_cast<int>(p[t * 3]) * k[t];
                tmp[i * out_stride + col] = static_cast<uint8_t>(clip8(ss));
            }
        }
    }
    __syncthreads();

    // ---- vertical pass
    uint8_t* d = dst + (static_cast<size_t>(tile) * Hc + r0) * out_stride;
    if ((out_stride & 3) == 0 && (reinterpret_cast<uintptr_t>(dst) & 3) == 0) {
        const int words = out_stride >> 2;
        for (int idx = tid; idx < nr * words; idx += RS_THREADS) {
            const int r = idx / words, w = idx - r * words;
            const int yo = cy + r0 + r;
            const int ymin = by[yo * 2] - y_first, n = by[yo * 2 + 1];
            const int* k = ky + yo * ksy;
            int a0 = 1 << (RS_PRECISION_BITS - 1), a1 = a0, a2 = a0, a3 = a0;
            const uint8_t* p = tmp + ymin * out_stride + w * 4;
            auto tap = [&](int kt) {
                const uint32_t v = *reinterpret_cast<const uint32_t*>(p);
                a0 += static_cast<int>(v & 255u) * kt;
                a1 += static_cast<int>((v >> 8) & 255u) * kt;
                a2 += static_cast<int>((v >> 16) & 255u) * kt;
                a3 += static_cast<int>(v >> 24) * kt;
                p += out_stride;
            };
            if constexpr (KS > 0) {
#pragma unroll
                for (int t = 0; t < KS; ++t) tap(__ldg(k + t));
            } else {
                for (int t = 0; t < n; ++t) tap(__ldg(k + t));
            }
            const uint32_t o = static_cast<uint32_t>(clip8(a0)) | (static_cast<uint32_t>(clip8(a1)) << 8) |
                               (static_cast<uint32_t>(clip8(a2)) << 16) | (static_cast<uint32_t>(clip8(a3)) << 24);
            *reinterpret_cast<uint32_t*>(d + static_cast<size_t>(r) * out_stride + w * 4) = o;
        }
    } else {
        for (int idx = tid; idx < nr * out_stride; idx += RS_THREADS) {
            const int r = idx / out_stride, col = idx - r * out_stride;
            const int yo = cy + r0 + r;
            const int ymin = by[yo * 2] - y_first, n = by[yo * 2 + 1];
            const int* k = ky + yo * ksy;
            int ss = 1 << (RS_PRECISION_BITS - 1);
            for (int t = 0; t < n; ++t) ss += static_cast<int>(tmp[(ymin + t) * out_stride + col]) * __ldg(k + t);
            d[idx] = static_cast<uint8_t>(clip8(ss));
        }
    }
}

}  // namespace
}  // namespace sb

extern "C" size_t stamp_resize_u8_smem_bytes(int Win, int Wc, int ksx, int max_in_rows) {
    if (Win <= 0 || Wc <= 0 || ksx <= 0 || max_in_rows <= 0) return 0;
    const size_t in_bytes = (static_cast<size_t>(max_in_rows) * Win * 3 + 32 + 15) & ~size_t(15);
    const size_t tmp_bytes = (static_cast<size_t>(max_in_rows + sb::RS_ROW_SLACK) * Wc * 3 + 15) & ~size_t(15);
    return in_bytes + tmp_bytes + static_cast<size_t>(Wc) * (ksx + 2) * sizeof(int);
}

extern "C" int stamp_resize_u8(const uint8_t* tiles, int n_tiles, int Hin, int Win, uint8_t* out, int Hc, int Wc,
                               int crop_y, int crop_x, const int* coef_x, const int* bounds_x, int ksize_x,
                               const int* coef_y, const int* bounds_y, int ksize_y, int rows_per_strip,
                               int max_in_rows, void* stream_) {
    using namespace sb;
    if (tiles == nullptr || out == nullptr || coef_x == nullptr || bounds_x == nullptr || coef_y == nullptr ||
        bounds_y == nullptr || n_tiles <= 0 || Hin <= 0 || Win <= 0 || Hc <= 0 || Wc <= 0 || crop_y < 0 || crop_x < 0 ||
        ksize_x <= 0 || ksize_y <= 0 || rows_per_strip <= 0 || max_in_rows <= 0 || max_in_rows > Hin)
        return SB_ERR_BAD_ARG;
    if (n_tiles > 65535) return SB_ERR_UNSUPPORTED;
    const size_t bytes = stamp_resize_u8_smem_bytes(Win, Wc, ksize_x, max_in_rows);
    if (bytes > 227 * 1024) return SB_ERR_UNSUPPORTED;
    const bool five = ksize_x == 5 && ksize_y == 5;           // bicubic / bilinear without down-sampling
    static size_t configured[2] = {0, 0};
    if (bytes > configured[five]) {
        const cudaError_t ce = five ? cudaFuncSetAttribute(resize_u8_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                           static_cast<int>(bytes))
                                    : cudaFuncSetAttribute(resize_u8_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                           static_cast<int>(bytes));
        if (ce != cudaSuccess) return SB_ERR_CUDA;
        configured[five] = bytes;
    }
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    ProfScope prof(PROF_MACENKO, static_cast<double>(n_tiles) * (static_cast<double>(Hin) * Win + static_cast<double>(Hc) * Wc) * 3.0,
                   stream);
    const dim3 grid((Hc + rows_per_strip - 1) / rows_per_strip, n_tiles);
    if (five)
        resize_u8_kernel<5><<<grid, RS_THREADS, bytes, stream>>>(tiles, Hin, Win, out, Hc, Wc, crop_y, crop_x, coef_x, bounds_x,
                                                                 ksize_x, coef_y, bounds_y, ksize_y, rows_per_strip, max_in_rows);
    else
        resize_u8_kernel<0><<<grid, RS_THREADS, bytes, stream>>>(tiles, Hin, Win, out, Hc, Wc, crop_y, crop_x, coef_x, bounds_x,
                                                                 ksize_x, coef_y, bounds_y, ksize_y, rows_per_strip, max_in_rows);
    count_launch();
    return cudaGetLastError() == cudaSuccess ? SB_OK : SB_ERR_CUDA;
}
