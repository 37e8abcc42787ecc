// Device half of the tile-cache JPEG decode: quantised DCT coefficients -> RGB tiles, bit-exact with what Pillow
// (libjpeg-turbo, library defaults) returns for the same file.
//
// reference call site: _tiles_from_cache_file, src/stamp/preprocessing/tiling.py:380-406 (Image.open + load per tile).
// libjpeg-turbo's default decompression path, restated in integer arithmetic (third-party library, not part of the
// reference tree; algorithm references are to its C sources):
//   jpeg_idct_islow      (jidctint.c)  dequantise, column pass (descale 11), row pass (descale 18), +128, clamp
//   h2v2_fancy_upsample  (jdsample.c)  triangle filter 3/4 - 1/4, vertical then horizontal, biases 8 / 7
//   ycc_rgb_convert      (jdcolor.c)   16-bit fixed-point YCbCr -> RGB
// Two kernels per batch: (1) one thread per 8x8 block writes the component planes (Y at full resolution, Cb / Cr at
// half or full resolution); (2) one thread per group of output pixels up-samples the chroma around it and converts.  The
// planes of a batch (225 KB per 224 px tile) stay in L2 between the two launches; HBM sees the coefficients once and
// the RGB tiles once.  (4:2:0: one thread per FOUR chroma samples = 2 x 8 output pixels.)
#include <cuda_runtime.h>

#include <cstdint>

#include "common.cuh"
#include "profile.cuh"
#include "stamp_b200.h"

namespace sb {
namespace {

constexpr int C_0_298631336 = 2446, C_0_390180644 = 3196, C_0_541196100 = 4433, C_0_765366865 = 6270;
constexpr int C_0_899976223 = 7373, C_1_175875602 = 9633, C_1_501321110 = 12299, C_1_847759065 = 15137;
constexpr int C_1_961570560 = 16069, C_2_053119869 = 16819, C_2_562915447 = 20995, C_3_072711026 = 25172;

// one 8-point pass of jpeg_idct_islow on v[0..7] (in place); SHIFT = descale of this pass
template <int SHIFT>
__device__ __forceinline__ void idct8(int (&v)[8]) {
    int z2 = v[2], z3 = v[6];
    int z1 = (z2 + z3) * C_0_541196100;
    const int tmp2 = z1 - z3 * C_1_847759065;
    const int tmp3 = z1 + z2 * C_0_765366865;
    const int tmp0 = (v[0] + v[4]) << 13;
    const int tmp1 = (v[0] - v[4]) << 13;
    const int t10 = tmp0 + tmp3, t13 = tmp0 - tmp3, t11 = tmp1 + tmp2, t12 = tmp1 - tmp2;
    int o0 = v[7], o1 = v[5], o2 = v[3], o3 = v[1];
    z1 = o0 + o3;
    z2 = o1 + o2;
    z3 = o0 + o2;
    int z4 = o1 + o3;
    const int z5 = (z3 + z4) * C_1_175875602;
    o0 *= C_0_298631336;
    o1 *= C_2_053119869;
    o2 *= C_3_072711026;
    o3 *= C_1_501321110;
    z1 *= -C_0_899976223;
    z2 *= -C_2_562915447;
    z3 = z3 * -C_1_961570560 + z5;
    z4 = z4 * -C_0_390180644 + z5;
    o0 += z1 + z3;
    o1 += z2 + z4;
    o2 += z2 + z3;
    o3 += z1 + z4;
    constexpr int R = 1 << (SHIFT - 1);
    v[0] = (t10 + o3 + R) >> SHIFT;
    v[7] = (t10 - o3 + R) >> SHIFT;
    v[1] = (t11 + o2 + R) >> SHIFT;
    v[6] = (t11 - o2 + R) >> SHIFT;
    v[2] = (t12 + o1 + R) >> SHIFT;
    v[5] = (t12 - o1 + R) >> SHIFT;
    v[3] = (t13 + o0 + R) >> SHIFT;
    v[4] = (t13 - o0 + R) >> SHIFT;
}

__device__ __forceinline__ int clamp8(int x) { return x < 0 ? 0 : (x > 255 ? 255 : x); }

struct JpegGeom {
    int bx[3], by[3];            // blocks per row / column of each component plane
    long long coef_off[3];       // first coefficient of the component inside a tile's coefficient record
    long long plane_off[3];      // first byte of the component inside a tile's plane record
    long long coefs_per_tile, plane_bytes_per_tile;
    int blocks_per_tile;
};

__global__ void __launch_bounds__(128)
jpeg_idct_kernel(const int16_t* __restrict__ coef, const uint16_t* __restrict__ quant, uint8_t* __restrict__ planes,
                 JpegGeom g, int n_tiles) {
    const long long gid = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (gid >= static_cast<long long>(n_tiles) * g.blocks_per_tile) return;
    const int tile = static_cast<int>(gid / g.blocks_per_tile);
    int b = static_cast<int>(gid - static_cast<long long>(tile) * g.blocks_per_tile);
    int c = 0;
    if (b >= g.bx[0] * g.by[0]) {
        b -= g.bx[0] * g.by[0];
        c = 1;
        if (b >= g.bx[1] * g.by[1]) {
            b -= g.bx[1] * g.by[1];
            c = 2;
        }
    }
    const int4* src = reinterpret_cast<const int4*>(coef + tile * g.coefs_per_tile + g.coef_off[c] + static_cast<long long>(b) * 64);
    const uint16_t* q = quant + (static_cast<long long>(tile) * 3 + c) * 64;
    int ws[8][8];   // [row][col]
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const int4 w = __ldg(src + r);
        const uint4 qa = __ldg(reinterpret_cast<const uint4*>(q) + r);
        const int cw[4] = {w.x, w.y, w.z, w.w};
        const uint32_t qw[4] = {qa.x, qa.y, qa.z, qa.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            ws[r][2 * k] = static_cast<int>(static_cast<int16_t>(cw[k] & 0xFFFF)) * static_cast<int>(qw[k] & 0xFFFFu);
            ws[r][2 * k + 1] = (cw[k] >> 16) * static_cast<int>(qw[k] >> 16);
        }
    }
#pragma unroll
    for (int col = 0; col < 8; ++col) {   // pass 1: columns
        int v[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) v[r] = ws[r][col];
        idct8<13 - 2>(v);
#pragma unroll
        for (int r = 0; r < 8; ++r) ws[r][col] = v[r];
    }
    const int by = b / g.bx[c], bx = b - by * g.bx[c];
    const int pitch = g.bx[c] * 8;
    uint8_t* dst = planes + tile * g.plane_bytes_per_tile + g.plane_off[c] + static_cast<long long>(by) * 8 * pitch + bx * 8;
#pragma unroll
    for (int r = 0; r < 8; ++r) {         // pass 2: rows
        idct8<13 + 2 + 3>(ws[r]);
        uint32_t lo = 0, hi = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            lo |= static_cast<uint32_t>(clamp8(ws[r][k] + 128)) << (8 * k);
            hi |= static_cast<uint32_t>(clamp8(ws[r][4 + k] + 128)) << (8 * k);
        }
        *reinterpret_cast<uint2*>(dst + r * pitch) = make_uint2(lo, hi);
    }
}

__device__ __forceinline__ void ycc_to_rgb(int y, int cb, int cr, uint8_t* o) {
    cb -= 128;
    cr -= 128;
    o[0] = static_cast<uint8_t>(clamp8(y + ((91881 * cr + 32768) >> 16)));                       // FIX(1.40200)
    o[1] = static_cast<uint8_t>(clamp8(y + ((-22554 * cb + 32768 - 46802 * cr) >> 16)));         // FIX(0.34414), FIX(0.71414)
    o[2] = static_cast<uint8_t>(clamp8(y + ((116130 * cb + 32768) >> 16)));                      // FIX(1.77200)
}

// 4:2:0 -- one thread per 4 chroma samples of a chroma row = 2 rows x 8 columns of output pixels: the vertical sums
// of the six chroma columns it touches are formed once, the luma and the output move as 64-bit words
__global__ void __launch_bounds__(256)
jpeg_color420_kernel(const uint8_t* __restrict__ planes, uint8_t* __restrict__ out, JpegGeom g, int n_tiles, int H, int W) {
    const int ds_w = (W + 1) >> 1, ds_h = (H + 1) >> 1;
    const int groups = (ds_w + 3) >> 2;
    const long long gid = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (gid >= static_cast<long long>(n_tiles) * groups * ds_h) return;
    const int tile = static_cast<int>(gid / (groups * ds_h));
    const int rem = static_cast<int>(gid - static_cast<long long>(tile) * groups * ds_h);
    const int cy = rem / groups, cx0 = (rem - cy * groups) * 4;
    const uint8_t* base = planes + tile * g.plane_bytes_per_tile;
    const int ypitch = g.bx[0] * 8, cpitch = g.bx[1] * 8;   // multiples of 8; the planes are padded to whole blocks
    // chroma rows with replicated edges (libjpeg's context rows), columns cx0-1 .. cx0+4 clamped to the plane
    const int ym = max(cy - 1, 0), yp = min(cy + 1, ds_h - 1);
    const int xl = max(cx0 - 1, 0);
    int up[2][2][8];   // [plane][output row][output column]
#pragma unroll
    for (int pl = 0; pl < 2; ++pl) {
        const uint8_t* cp = base + g.plane_off[1 + pl];
        int t[6], u[6];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const uint8_t* row = cp + (r == 0 ? ym : (r == 1 ? cy : yp)) * cpitch;
            const uint32_t w = *reinterpret_cast<const uint32_t*>(row + cx0);     // cx0 % 4 == 0, pitch % 8 == 0
            int v[6];
            v[0] = row[xl];
#pragma unroll
            for (int k = 0; k < 4; ++k) v[1 + k] = min(cx0 + k, ds_w - 1) == cx0 + k ? static_cast<int>((w >> (8 * k)) & 255u) : 0;
            // columns past the last chroma sample replicate it (right edge of the image)
#pragma unroll
            for (int k = 1; k < 4; ++k)
                if (cx0 + k > ds_w - 1) v[1 + k] = v[k];
            v[5] = (cx0 + 4 <= ds_w - 1) ? static_cast<int>(row[cx0 + 4]) : v[4];
#pragma unroll
            for (int k = 0; k < 6; ++k) {
                if (r == 0) t[k] = v[k];
                else if (r == 1) { t[k] += 3 * v[k]; u[k] = 3 * v[k]; }
                else u[k] += v[k];
            }
        }
        // horizontal: even column looks left (bias 8), odd column looks right (bias 7); at the plane's first / last
        // column the neighbour is the column itself, which reproduces libjpeg's special cases
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            up[pl][0][2 * k] = (3 * t[1 + k] + t[k] + 8) >> 4;
            up[pl][0][2 * k + 1] = (3 * t[1 + k] + t[2 + k] + 7) >> 4;
            up[pl][1][2 * k] = (3 * u[1 + k] + u[k] + 8) >> 4;
            up[pl][1][2 * k + 1] = (3 * u[1 + k] + u[2 + k] + 7) >> 4;
        }
    }
    const uint8_t* yp0 = base + g.plane_off[0];
    uint8_t* o = out + static_cast<long long>(tile) * H * W * 3;
    const int x0 = 2 * cx0;
    const int ncol = min(8, W - x0);
#pragma unroll
    for (int dy = 0; dy < 2; ++dy) {
        const int y = 2 * cy + dy;
        if (y >= H) break;
        const uint2 yw = *reinterpret_cast<const uint2*>(yp0 + y * ypitch + x0);   // x0 % 8 == 0
        uint8_t px[24];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int yy = static_cast<int>(((k < 4 ? yw.x : yw.y) >> (8 * (k & 3))) & 255u);
            ycc_to_rgb(yy, up[0][dy][k], up[1][dy][k], px + 3 * k);
        }
        uint8_t* d = o + (static_cast<long long>(y) * W + x0) * 3;
        if (ncol == 8 && (reinterpret_cast<uintptr_t>(d) & 7) == 0) {
            uint32_t w32[6];
#pragma unroll
            for (int k = 0; k < 6; ++k)
                w32[k] = px[4 * k] | (px[4 * k + 1] << 8) | (px[4 * k + 2] << 16) | (static_cast<uint32_t>(px[4 * k + 3]) << 24);
            reinterpret_cast<uint2*>(d)[0] = make_uint2(w32[0], w32[1]);
            reinterpret_cast<uint2*>(d)[1] = make_uint2(w32[2], w32[3]);
            reinterpret_cast<uint2*>(d)[2] = make_uint2(w32[4], w32[5]);
        } else {
            for (int k = 0; k < 3 * ncol; ++k) d[k] = px[k];
        }
    }
}

// 4:4:4 -- one thread per pixel
__global__ void __launch_bounds__(256)
jpeg_color444_kernel(const uint8_t* __restrict__ planes, uint8_t* __restrict__ out, JpegGeom g, int n_tiles, int H, int W) {
    const long long gid = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (gid >= static_cast<long long>(n_tiles) * H * W) return;
    const int tile = static_cast<int>(gid / (H * W));
    const int rem = static_cast<int>(gid - static_cast<long long>(tile) * H * W);
    const int y = rem / W, x = rem - y * W;
    const uint8_t* base = planes + tile * g.plane_bytes_per_tile;
    const int pitch = g.bx[0] * 8;
    uint8_t px[3];
    ycc_to_rgb(base[g.plane_off[0] + y * pitch + x], base[g.plane_off[1] + y * pitch + x], base[g.plane_off[2] + y * pitch + x], px);
    uint8_t* d = out + gid * 3;
    d[0] = px[0]; d[1] = px[1]; d[2] = px[2];
}

bool make_geom(const StampJpegInfo* info, JpegGeom* g) {
    if (info == nullptr || info->n_comp != 3 || info->width <= 0 || info->height <= 0 || info->mcus_x <= 0 || info->mcus_y <= 0)
        return false;
    const bool s420 = info->h[0] == 2 && info->v[0] == 2, s444 = info->h[0] == 1 && info->v[0] == 1;
    if (!(s420 || s444) || info->h[1] != 1 || info->v[1] != 1 || info->h[2] != 1 || info->v[2] != 1) return false;
    long long co = 0, po = 0;
    int blocks = 0;
    for (int c = 0; c < 3; ++c) {
        g->bx[c] = info->mcus_x * info->h[c];
        g->by[c] = info->mcus_y * info->v[c];
        g->coef_off[c] = co;
        g->plane_off[c] = po;
        co += static_cast<long long>(g->bx[c]) * g->by[c] * 64;
        po += static_cast<long long>(g->bx[c]) * g->by[c] * 64;
        blocks += g->bx[c] * g->by[c];
    }
    g->coefs_per_tile = co;
    g->plane_bytes_per_tile = (po + 15) / 16 * 16;
    g->blocks_per_tile = blocks;
    return true;
}

}  // namespace
}  // namespace sb

extern "C" size_t stamp_jpeg_workspace_bytes(const StampJpegInfo* info, int n_tiles) {
    sb::JpegGeom g;
    if (n_tiles <= 0 || !sb::make_geom(info, &g)) return 0;
    return static_cast<size_t>(g.plane_bytes_per_tile) * n_tiles;
}

extern "C" int stamp_jpeg_decode_coefs_u8(const StampJpegInfo* info, const int16_t* coef, const uint16_t* quant, int n_tiles,
                                          uint8_t* out, void* workspace, size_t workspace_bytes, void* stream_) {
    using namespace sb;
    JpegGeom g;
    if (coef == nullptr || quant == nullptr || out == nullptr || workspace == nullptr || n_tiles <= 0 || !make_geom(info, &g))
        return SB_ERR_BAD_ARG;
    if (workspace_bytes < static_cast<size_t>(g.plane_bytes_per_tile) * n_tiles) return SB_ERR_WORKSPACE;
    if ((reinterpret_cast<uintptr_t>(coef) & 15) != 0 || (reinterpret_cast<uintptr_t>(quant) & 15) != 0 ||
        (reinterpret_cast<uintptr_t>(workspace) & 15) != 0)
        return SB_ERR_BAD_ARG;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    uint8_t* planes = static_cast<uint8_t*>(workspace);
    const int H = info->height, W = info->width;
    ProfScope prof(PROF_MACENKO, static_cast<double>(n_tiles) * (g.coefs_per_tile * 2.0 + 3.0 * H * W), stream);
    const long long blocks = static_cast<long long>(n_tiles) * g.blocks_per_tile;
    jpeg_idct_kernel<<<static_cast<unsigned>((blocks + 127) / 128), 128, 0, stream>>>(coef, quant, planes, g, n_tiles);
    if (info->h[0] == 2) {
        const long long groups = static_cast<long long>(n_tiles) * (((W + 1) / 2 + 3) / 4) * ((H + 1) / 2);
        jpeg_color420_kernel<<<static_cast<unsigned>((groups + 255) / 256), 256, 0, stream>>>(planes, out, g, n_tiles, H, W);
    } else {
        const long long px = static_cast<long long>(n_tiles) * H * W;
        jpeg_color444_kernel<<<static_cast<unsigned>((px + 255) / 256), 256, 0, stream>>>(planes, out, g, n_tiles, H, W);
    }
    count_launch(2);
    return cudaGetLastError() == cudaSuccess ? SB_OK : SB_ERR_CUDA;
}
