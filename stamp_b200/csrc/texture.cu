// Tissue-texture filter for tile batches: grayscale + Canny edge fraction per tile, bit-exact with the
// Pillow / OpenCV calls of the reference.
//
// replaces: _has_enough_texture, src/stamp/preprocessing/tiling.py:279-291
//     tile.convert("L")  ->  cv2.Canny(gray, 40, 100)  ->  edges.mean() / 255 >= cutoff
// (the reference runs it per tile on the CPU inside the tiling workers.)
//
// Integer / byte work, one CTA per tile, the whole tile lives in shared memory (224 x 224: 50 KB gray,
// 102 KB 16-bit gradient magnitudes with a zero border, 51 KB state map), so HBM sees one read of the RGB
// bytes and 4 bytes of result per tile:
//   1. L = (19595 R + 38470 G + 7471 B + 0x8000) >> 16                          (Pillow rgb2l)
//   2. 3x3 Sobel with replicated borders, mag = |dx| + |dy|                       (cv::Canny, L1 gradient)
//   3. non-maximum suppression along the gradient direction quantised with TG22 = 13573 (2^15 tan 22.5 deg),
//      OpenCV's tie rules (">" towards the previous neighbour, ">=" towards the next one on the axes)
//   4. hysteresis as a monotone fixpoint in shared memory: weak pixels 8-connected to an edge become edges,
//      iterated over a compacted list of the weak pixels only
//   5. edge count (-> edges.mean() / 255 on the host, in the same double arithmetic as NumPy)
#include <math.h>

#include "common.cuh"
#include "stamp_b200.h"

namespace sb {
namespace {

constexpr int TX_THREADS = 512;

__device__ __forceinline__ void sobel_at(const uint8_t* __restrict__ g, int H, int W, int y, int x, int& dx, int& dy) {
    const int ym = y > 0 ? y - 1 : 0, yp = y < H - 1 ? y + 1 : H - 1;
    const int xm = x > 0 ? x - 1 : 0, xp = x < W - 1 ? x + 1 : W - 1;
    const int a = g[ym * W + xm], b = g[ym * W + x], c = g[ym * W + xp];
    const int d = g[y * W + xm], f = g[y * W + xp];
    const int p = g[yp * W + xm], q = g[yp * W + x], r = g[yp * W + xp];
    dx = (c + 2 * f + r) - (a + 2 * d + p);
    dy = (p + 2 * q + r) - (a + 2 * b + c);
}

__global__ void __launch_bounds__(TX_THREADS)
texture_kernel(const uint8_t* __restrict__ tiles, int H, int W, int low, int high, int* __restrict__ edge_count,
               uint8_t* __restrict__ edges_out) {
    extern __shared__ __align__(16) uint8_t smem[];
    const int HW = H * W, Wp = W + 2, HWp = (H + 2) * Wp;
    uint16_t* mag = reinterpret_cast<uint16_t*>(smem);              // [(H+2)][(W+2)], zero border
    uint8_t* map = smem + (static_cast<size_t>(HWp) * 2 + 15) / 16 * 16;   // same shape: 0 weak, 1 not an edge, 2 edge
    uint8_t* gray = map + (HWp + 15) / 16 * 16;                            // [H][W]
    const int tid = threadIdx.x;
    const uint8_t* src = tiles + static_cast<size_t>(blockIdx.x) * HW * 3;

    for (int i = tid; i < HWp; i += TX_THREADS) { mag[i] = 0; map[i] = 1; }
    // grayscale: 4 pixels = 12 bytes = three 32-bit words per step when the tile allows it
    if ((HW & 3) == 0 && (reinterpret_cast<uintptr_t>(src) & 3) == 0) {
        const uint32_t* s32 = reinterpret_cast<const uint32_t*>(src);
        for (int i = tid; i < HW / 4; i += TX_THREADS) {
            const uint32_t w0 = __ldg(s32 + 3 * i), w1 = __ldg(s32 + 3 * i + 1), w2 = __ldg(s32 + 3 * i + 2);
            const uint32_t r0 = w0 & 255, g0 = (w0 >> 8) & 255, b0 = (w0 >> 16) & 255;
            const uint32_t r1 = w0 >> 24, g1 = w1 & 255, b1 = (w1 >> 8) & 255;
            const uint32_t r2 = (w1 >> 16) & 255, g2 = w1 >> 24, b2 = w2 & 255;
            const uint32_t r3 = (w2 >> 8) & 255, g3 = (w2 >> 16) & 255, b3 = w2 >> 24;
            uchar4 o;
            o.x = static_cast<uint8_t>((r0 * 19595u + g0 * 38470u + b0 * 7471u + 0x8000u) >> 16);
            o.y = static_cast<uint8_t>((r1 * 19595u + g1 * 38470u + b1 * 7471u + 0x8000u) >> 16);
            o.z = static_cast<uint8_t>((r2 * 19595u + g2 * 38470u + b2 * 7471u + 0x8000u) >> 16);
            o.w = static_cast<uint8_t>((r3 * 19595u + g3 * 38470u + b3 * 7471u + 0x8000u) >> 16);
            reinterpret_cast<uchar4*>(gray)[i] = o;
        }
    } else {
        for (int i = tid; i < HW; i += TX_THREADS) {
            const uint32_t r = __ldg(src + 3 * i), g = __ldg(src + 3 * i + 1), b = __ldg(src + 3 * i + 2);
            gray[i] = static_cast<uint8_t>((r * 19595u + g * 38470u + b * 7471u + 0x8000u) >> 16);
        }
    }
    __syncthreads();

    // strided pixel loops keep (y, x) incrementally: no integer division per pixel
    const int sy = TX_THREADS / W, sx = TX_THREADS % W, y_first = tid / W, x_first = tid % W;
#define SB_NEXT_PIXEL() do { x += sx; y += sy; if (x >= W) { x -= W; ++y; } } while (0)
    for (int i = tid, y = y_first, x = x_first; i < HW; i += TX_THREADS) {
        int dx, dy;
        sobel_at(gray, H, W, y, x, dx, dy);
        mag[(y + 1) * Wp + x + 1] = static_cast<uint16_t>(abs(dx) + abs(dy));
        SB_NEXT_PIXEL();
    }
    __syncthreads();

    for (int i = tid, y = y_first, x = x_first; i < HW; i += TX_THREADS) {
        const int c = (y + 1) * Wp + x + 1;
        const int m = mag[c];
        SB_NEXT_PIXEL();
        if (m > low) {
            const int x0 = c % Wp - 1, y0 = c / Wp - 1;
            int dx, dy;
            sobel_at(gray, H, W, y0, x0, dx, dy);
            const int ax = abs(dx);
            const int ay = abs(dy) << 15;
            const int tg22x = ax * 13573;
            bool keep;
            if (ay < tg22x) {
                keep = m > mag[c - 1] && m >= mag[c + 1];
            } else {
                const int tg67x = tg22x + (ax << 16);
                if (ay > tg67x) {
                    keep = m > mag[c - Wp] && m >= mag[c + Wp];
                } else {
                    const int s = ((dx ^ dy) < 0) ? -1 : 1;
                    keep = m > mag[c - Wp - s] && m > mag[c + Wp + s];
                }
            }
            if (keep) map[c] = (m > high) ? 2 : 0;
        }
    }
    __syncthreads();

    // hysteresis.  Weak pixels are few (a few percent of a tissue tile): compact their positions into a list
    // (16-bit indices into the padded map, stored over the no longer needed grayscale buffer) and iterate the
    // monotone fixpoint "weak with an edge neighbour -> edge" over the list only; tiles with more weak pixels
    // than the list holds, or too large for 16-bit indices, sweep all pixels instead.
    __shared__ int n_weak;
    if (tid == 0) n_weak = 0;
    __syncthreads();
    const int per = (HW + TX_THREADS - 1) / TX_THREADS;
    const int beg = min(HW, tid * per), end = min(HW, beg + per);
    uint16_t* list = reinterpret_cast<uint16_t*>(gray);
    const int cap = (HWp < 65536) ? HW / 2 : 0;
    for (int i = beg, y = beg / W, x = beg % W; i < end; ++i) {
        const int c = (y + 1) * Wp + x + 1;
        if (map[c] == 0) {
            const int k = atomicAdd(&n_weak, 1);
            if (k < cap) list[k] = static_cast<uint16_t>(c);
        }
        if (++x == W) { x = 0; ++y; }
    }
    __syncthreads();
    auto promote = [&](int c) -> int {
        if (map[c] != 0) return 0;
        const bool hit = map[c - 1] == 2 || map[c + 1] == 2 || map[c - Wp] == 2 || map[c + Wp] == 2 ||
                         map[c - Wp - 1] == 2 || map[c - Wp + 1] == 2 || map[c + Wp - 1] == 2 || map[c + Wp + 1] == 2;
        if (hit) map[c] = 2;
        return hit ? 1 : 0;
    };
    const int nw = n_weak;
    if (nw <= cap) {
        if (nw > 0)
            for (;;) {
                int changed = 0;
                for (int k = tid; k < nw; k += TX_THREADS) changed |= promote(list[k]);
                for (int k = nw - 1 - tid; k >= 0; k -= TX_THREADS) changed |= promote(list[k]);
                if (!__syncthreads_or(changed)) break;
            }
    } else {
        for (;;) {   // this thread's contiguous run of pixels, swept forwards then backwards each round
            int changed = 0;
            for (int i = beg; i < end; ++i) { const int y = i / W; changed |= promote((y + 1) * Wp + (i - y * W) + 1); }
            for (int i = end - 1; i >= beg; --i) { const int y = i / W; changed |= promote((y + 1) * Wp + (i - y * W) + 1); }
            if (!__syncthreads_or(changed)) break;
        }
    }

    int cnt = 0;
    for (int i = tid, y = y_first, x = x_first; i < HW; i += TX_THREADS) {
        const bool e = map[(y + 1) * Wp + x + 1] == 2;
        cnt += e ? 1 : 0;
        if (edges_out != nullptr) edges_out[static_cast<size_t>(blockIdx.x) * HW + i] = e ? 255 : 0;
        SB_NEXT_PIXEL();
    }
#undef SB_NEXT_PIXEL
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    __shared__ int red[TX_THREADS / 32];
    if ((tid & 31) == 0) red[tid >> 5] = cnt;
    __syncthreads();
    if (tid == 0) {
        int t = 0;
        for (int i = 0; i < TX_THREADS / 32; ++i) t += red[i];
        edge_count[blockIdx.x] = t;
    }
}

inline size_t texture_smem(int H, int W) {
    const size_t hwp = static_cast<size_t>(H + 2) * (W + 2);
    return (hwp * 2 + 15) / 16 * 16 + (hwp + 15) / 16 * 16 + static_cast<size_t>(H) * W + 16;
}

}  // namespace
}  // namespace sb

extern "C" int stamp_tile_texture_u8(const uint8_t* tiles, int n_tiles, int H, int W, int low, int high,
                                     int* edge_count, uint8_t* edges_out, void* stream_) {
    using namespace sb;
    if (tiles == nullptr || edge_count == nullptr || n_tiles <= 0 || H <= 0 || W <= 0 || low < 0 || high < low)
        return SB_ERR_BAD_ARG;
    const size_t bytes = texture_smem(H, W);
    if (bytes > 227 * 1024 || H * W > (1 << 20)) return SB_ERR_UNSUPPORTED;   // tile must fit one SM's shared memory
    static size_t configured = 0;
    if (bytes > configured) {
        if (cudaFuncSetAttribute(texture_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(bytes)) != cudaSuccess)
            return SB_ERR_CUDA;
        configured = bytes;
    }
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    ProfScope prof(PROF_MACENKO, static_cast<double>(n_tiles) * H * W * 3.0, stream);
    texture_kernel<<<n_tiles, TX_THREADS, bytes, stream>>>(tiles, H, W, low, high, edge_count, edges_out);
    count_launch();
    return cudaGetLastError() == cudaSuccess ? SB_OK : SB_ERR_CUDA;
}
