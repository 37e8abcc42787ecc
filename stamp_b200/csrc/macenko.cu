// Macenko stain normalisation over a batch of uint8 RGB tiles, entirely on the device.
//
// The reference snapshot has no Macenko code (README.md:35 is the only mention); the stage is named
// by BASELINE.json's north_star and specified in SURVEY.md 8c / oracle/macenko_oracle.py.
//
// HBM-bound integer/byte work: every pass streams the uint8 pixels with 128-bit loads (16 pixels =
// 48 bytes per thread iteration), optical densities come from a 256-entry shared-memory table
// (OD depends on the byte value only), reductions are warp shuffles + one fp64 atomic per CTA, and
// the four exact percentiles (two of the stain angle, two of the stain concentrations, each needing
// the k-th and (k+1)-th order statistic for NumPy's linear interpolation) are found with a 2-pass
// radix select (11 + 11 of the 32 key bits; the last 10 bits are below output resolution) over
// order-preserving fixed-point keys: shared-memory histograms,
// no sort, no per-pixel intermediate in HBM.  A batch of tiles (<= ~800) stays in the 126 MB L2
// after the first pass, so DRAM traffic stays close to the algorithmic one read + one write.
//
//   stats -> eig -> 3 x (hist_phi, select) -> stain vectors -> 3 x (hist_conc, select) -> apply
#include <math.h>

#include "common.cuh"
#include "stamp_b200.h"

namespace sb {
namespace {

constexpr int NSEL = 4;          // order statistics searched concurrently per stage
constexpr int NBINS = 2048;
constexpr int MAC_THREADS = 256;
constexpr int PIX_PER_ITER = 16;  // 48 bytes
constexpr int MIN_TISSUE = 16;
constexpr int MAC_PASSES = 2;     // radix passes per order statistic: 11 + 11 key bits (a third would add the last 10)

struct MacGroup {
    double sum[3];
    double sq[6];                 // xx xy xz yy yz zz
    unsigned long long n_kept;
    unsigned long long n_all;
    unsigned long long rank[NSEL];
    unsigned int prefix[NSEL];
    float frac[2];                // interpolation weights of the two percentiles of the stage
    float E[6];                   // E[c*2 + j]: plane basis, j = 0 second-largest, 1 largest eigenvector
    float pinv[6];                // pinv[s*3 + c]
    float HE[6];                  // HE[c*2 + s], s = 0 haematoxylin, 1 eosin
    float maxC[2];
    float scale[2];               // maxCRef / maxC
    int valid;
    int pad;
};

// Order-preserving 32-bit fixed-point keys.  Uniform resolution spreads the values over the radix
// histograms (an IEEE-bit key would put a whole stain-angle distribution into 2-4 of the 2048 top-bit
// bins and serialise the shared-memory atomics).
// Stain angle: atan2 is replaced per pixel by the pseudo-angle pa(x, y) in [-2, 2], strictly monotone
// in atan2(y, x), so both have the same order statistics; the four selected keys are mapped back to
// radians in fp64 (pa_to_angle).  Resolution 2^-30 * 4 ~ 4e-9.
__device__ __forceinline__ float pseudo_angle(float x, float y) {
    const float ay = fabsf(y);
    const float r = __fdividef(ay, fabsf(x) + ay + 1e-30f);
    return copysignf(x >= 0.f ? r : 2.0f - r, y);
}
__device__ __forceinline__ unsigned int angle_key(float pa) {
    return static_cast<unsigned int>(fminf(fmaxf((pa + 2.0f) * 1073741824.0f, 0.f), 4294967040.f));
}
__device__ double pa_to_angle(unsigned int key) {
    const double pa = static_cast<double>(key) / 1073741824.0 - 2.0;
    const double a = fabs(pa);
    const double r = a <= 1.0 ? a : 2.0 - a;
    const double ang = atan2(r, a <= 1.0 ? 1.0 - r : -(1.0 - r));
    return pa < 0 ? -ang : ang;
}
// Stain concentrations: C = pinv . OD, |C| < 64 by a wide margin; resolution 2^-25 ~ 3e-8
__device__ __forceinline__ unsigned int conc_key(float c) {
    return static_cast<unsigned int>(fminf(fmaxf((c + 64.0f) * 33554432.0f, 0.f), 4294967040.f));
}
__device__ double key_to_conc(unsigned int key) { return static_cast<double>(key) / 33554432.0 - 64.0; }

__device__ __forceinline__ void load_lut(float* lut, float Io) {
    for (int i = threadIdx.x; i < 256; i += blockDim.x) lut[i] = -logf((static_cast<float>(i) + 1.0f) / Io);
}

// 16 pixels (48 bytes) -> rgb[16][3]
__device__ __forceinline__ void load_pixels(const uint8_t* p, uint8_t (&px)[48]) {
    const uint4 a = ld_nc_v4(p), b = ld_nc_v4(p + 16), c = ld_nc_v4(p + 32);
    *reinterpret_cast<uint4*>(px) = a;
    *reinterpret_cast<uint4*>(px + 16) = b;
    *reinterpret_cast<uint4*>(px + 32) = c;
}

// ---- pass 1: per-group sums for the covariance of the tissue optical densities ----------------
__global__ void __launch_bounds__(MAC_THREADS)
mac_stats_kernel(const uint8_t* __restrict__ img, long long n_chunks, long long chunks_per_group,
                 float Io, float beta, MacGroup* __restrict__ grp) {
    __shared__ float lut[256];
    __shared__ double red[MAC_THREADS / 32][10];
    load_lut(lut, Io);
    __syncthreads();
    // a CTA works on a contiguous span of chunks so that (almost) all of it belongs to one group
    const long long per_cta = (n_chunks + gridDim.x - 1) / gridDim.x;
    const long long c0 = blockIdx.x * per_cta;
    const long long c1 = min(n_chunks, c0 + per_cta);
    long long cur_group = -1;
    double acc[10];
#pragma unroll
    for (int i = 0; i < 10; ++i) acc[i] = 0.0;

    auto flush = [&](long long g) {  // called by all threads of the CTA
        if (g >= 0) {
#pragma unroll
            for (int i = 0; i < 10; ++i) acc[i] = warp_sum_d(acc[i]);
            const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
            if (l == 0)
                for (int i = 0; i < 10; ++i) red[w][i] = acc[i];
            __syncthreads();
            if (threadIdx.x < 10) {
                double t = 0.0;
                for (int k = 0; k < MAC_THREADS / 32; ++k) t += red[k][threadIdx.x];
                if (threadIdx.x < 3) atomicAdd(&grp[g].sum[threadIdx.x], t);
                else if (threadIdx.x < 9) atomicAdd(&grp[g].sq[threadIdx.x - 3], t);
                else atomicAdd(&grp[g].n_kept, static_cast<unsigned long long>(t + 0.5));
            }
            __syncthreads();
        }
#pragma unroll
        for (int i = 0; i < 10; ++i) acc[i] = 0.0;
    };

    for (long long base = c0; base < c1; base += MAC_THREADS) {
        const long long g_first = base / chunks_per_group;
        const long long g_last = (min(c1, base + MAC_THREADS) - 1) / chunks_per_group;
        const bool uniform = g_first == g_last;   // CTA-uniform
        if (uniform && g_first != cur_group) { flush(cur_group); cur_group = g_first; }
        const long long c = base + threadIdx.x;
        if (c < c1) {
            float f[10];
#pragma unroll
            for (int i = 0; i < 10; ++i) f[i] = 0.f;
            uint8_t px[48];
            load_pixels(img + c * 48, px);
#pragma unroll
            for (int k = 0; k < PIX_PER_ITER; ++k) {
                const float r = lut[px[3 * k]], g = lut[px[3 * k + 1]], b = lut[px[3 * k + 2]];
                if (r >= beta && g >= beta && b >= beta) {
                    f[0] += r; f[1] += g; f[2] += b;
                    f[3] += r * r; f[4] += r * g; f[5] += r * b;
                    f[6] += g * g; f[7] += g * b; f[8] += b * b;
                    f[9] += 1.f;
                }
            }
            if (uniform) {
#pragma unroll
                for (int i = 0; i < 10; ++i) acc[i] += static_cast<double>(f[i]);
            } else if (f[9] > 0.f) {
                // stripe straddling two fit groups (at most one per group boundary): direct atomics
                MacGroup& G = grp[c / chunks_per_group];
                for (int i = 0; i < 3; ++i) atomicAdd(&G.sum[i], static_cast<double>(f[i]));
                for (int i = 0; i < 6; ++i) atomicAdd(&G.sq[i], static_cast<double>(f[3 + i]));
                atomicAdd(&G.n_kept, static_cast<unsigned long long>(f[9] + 0.5f));
            }
        }
    }
    flush(cur_group);
}

// ---- 3x3 symmetric eigen-decomposition (cyclic Jacobi, fp64), plane basis, ranks for the phi stage
__device__ void jacobi3(double A[3][3], double V[3][3]) {
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) V[i][j] = (i == j) ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 24; ++sweep) {
        const double off = fabs(A[0][1]) + fabs(A[0][2]) + fabs(A[1][2]);
        if (off < 1e-300) break;
        for (int p = 0; p < 2; ++p)
            for (int q = p + 1; q < 3; ++q) {
                if (fabs(A[p][q]) < 1e-300) continue;
                const double theta = (A[q][q] - A[p][p]) / (2.0 * A[p][q]);
                const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < 3; ++k) {
                    const double akp = A[k][p], akq = A[k][q];
                    A[k][p] = c * akp - s * akq;
                    A[k][q] = s * akp + c * akq;
                }
                for (int k = 0; k < 3; ++k) {
                    const double apk = A[p][k], aqk = A[q][k];
                    A[p][k] = c * apk - s * aqk;
                    A[q][k] = s * apk + c * aqk;
                }
                for (int k = 0; k < 3; ++k) {
                    const double vkp = V[k][p], vkq = V[k][q];
                    V[k][p] = c * vkp - s * vkq;
                    V[k][q] = s * vkp + c * vkq;
                }
            }
    }
}

__device__ void set_ranks(MacGroup& g, unsigned long long n, double q_lo, double q_hi) {
    const double p0 = static_cast<double>(n - 1) * q_lo, p1 = static_cast<double>(n - 1) * q_hi;
    const unsigned long long k0 = static_cast<unsigned long long>(floor(p0));
    const unsigned long long k1 = static_cast<unsigned long long>(floor(p1));
    g.rank[0] = k0; g.rank[1] = min(k0 + 1, n - 1);
    g.rank[2] = k1; g.rank[3] = min(k1 + 1, n - 1);
    g.frac[0] = static_cast<float>(p0 - static_cast<double>(k0));
    g.frac[1] = static_cast<float>(p1 - static_cast<double>(k1));
    for (int i = 0; i < NSEL; ++i) g.prefix[i] = 0;
}

__global__ void mac_eig_kernel(MacGroup* __restrict__ grp, long long pixels_per_group,
                               long long n_pixels, float alpha) {
    MacGroup& g = grp[blockIdx.x];
    if (threadIdx.x != 0) return;
    const long long first = blockIdx.x * pixels_per_group;
    g.n_all = static_cast<unsigned long long>(min(pixels_per_group, n_pixels - first));
    const double n = static_cast<double>(g.n_kept);
    g.valid = g.n_kept >= MIN_TISSUE;
    if (!g.valid) return;
    const double mu[3] = {g.sum[0] / n, g.sum[1] / n, g.sum[2] / n};
    double A[3][3], V[3][3];
    const int idx[3][3] = {{0, 1, 2}, {1, 3, 4}, {2, 4, 5}};
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) A[i][j] = (g.sq[idx[i][j]] - n * mu[i] * mu[j]) / (n - 1.0);
    jacobi3(A, V);
    int order[3] = {0, 1, 2};  // ascending eigenvalues
    for (int i = 0; i < 2; ++i)
        for (int j = 0; j < 2 - i; ++j)
            if (A[order[j]][order[j]] > A[order[j + 1]][order[j + 1]]) { int t = order[j]; order[j] = order[j + 1]; order[j + 1] = t; }
    double e0[3], e1[3];  // second largest, largest
    for (int c = 0; c < 3; ++c) { e0[c] = V[c][order[1]]; e1[c] = V[c][order[2]]; }
    if (e1[0] + e1[1] + e1[2] < 0) for (int c = 0; c < 3; ++c) e1[c] = -e1[c];
    if (e0[0] < 0) for (int c = 0; c < 3; ++c) e0[c] = -e0[c];
    for (int c = 0; c < 3; ++c) { g.E[c * 2] = static_cast<float>(e0[c]); g.E[c * 2 + 1] = static_cast<float>(e1[c]); }
    set_ranks(g, g.n_kept, alpha / 100.0, 1.0 - alpha / 100.0);
}

// ---- radix-select histogram pass (stage 0: stain angle of tissue pixels; stage 1: concentrations)
template <int STAGE>
__global__ void __launch_bounds__(MAC_THREADS)
mac_hist_kernel(const uint8_t* __restrict__ img, long long n_chunks, long long chunks_per_group,
                float Io, float beta, const MacGroup* __restrict__ grp, unsigned int* __restrict__ hist,
                int pass) {
    __shared__ float lut[256];
    __shared__ unsigned int h[NSEL][NBINS];
    load_lut(lut, Io);
    for (int i = threadIdx.x; i < NSEL * NBINS; i += blockDim.x) (&h[0][0])[i] = 0;
    const long long per_cta = (n_chunks + gridDim.x - 1) / gridDim.x;
    const long long c0 = blockIdx.x * per_cta;
    const long long c1 = min(n_chunks, c0 + per_cta);
    // this pass looks at key bits [shift, shift + nbits); higher bits must equal the selection's prefix
    const int shift = (pass == 0) ? 21 : (pass == 1 ? 10 : 0);
    const int nbits = (pass == 2) ? 10 : 11;
    const int hi_shift = shift + nbits;
    long long cur_group = -1, my_group = -1;
    float e[6], pv[6];
    unsigned int prefix[NSEL];
    bool gvalid = false;

    auto flush = [&](long long g) {  // called by all threads of the CTA
        __syncthreads();
        if (g >= 0)
            for (int i = threadIdx.x; i < NSEL * NBINS; i += blockDim.x) {
                const unsigned int v = (&h[0][0])[i];
                if (v) {
                    atomicAdd(hist + g * (NSEL * NBINS) + i, v);
                    (&h[0][0])[i] = 0;
                }
            }
        __syncthreads();
    };

    __syncthreads();
    for (long long base = c0; base < c1; base += MAC_THREADS) {
        const long long g_first = base / chunks_per_group;
        const long long g_last = (min(c1, base + MAC_THREADS) - 1) / chunks_per_group;
        const bool uniform = g_first == g_last;   // CTA-uniform
        if (uniform && g_first != cur_group) { flush(cur_group); cur_group = g_first; }
        const long long c = base + threadIdx.x;
        if (c >= c1) continue;
        const long long g_mine = c / chunks_per_group;
        if (g_mine != my_group) {
            my_group = g_mine;
            const MacGroup& G = grp[g_mine];
            gvalid = G.valid != 0;
#pragma unroll
            for (int i = 0; i < 6; ++i) { e[i] = G.E[i]; pv[i] = G.pinv[i]; }
#pragma unroll
            for (int i = 0; i < NSEL; ++i) prefix[i] = G.prefix[i];
        }
        if (!gvalid) continue;
        uint8_t px[48];
        load_pixels(img + c * 48, px);
#pragma unroll
        for (int k = 0; k < PIX_PER_ITER; ++k) {
            const float r = lut[px[3 * k]], gg = lut[px[3 * k + 1]], b = lut[px[3 * k + 2]];
            unsigned int key[2];
            bool use = true;
            if (STAGE == 0) {
                use = (r >= beta && gg >= beta && b >= beta);
                const float t0 = r * e[0] + gg * e[2] + b * e[4];
                const float t1 = r * e[1] + gg * e[3] + b * e[5];
                key[0] = key[1] = angle_key(pseudo_angle(t0, t1));
            } else {
                key[0] = conc_key(pv[0] * r + pv[1] * gg + pv[2] * b);
                key[1] = conc_key(pv[3] * r + pv[4] * gg + pv[5] * b);
            }
            if (use) {
#pragma unroll
                for (int s = 0; s < NSEL; ++s) {
                    const unsigned int kk = key[s >> 1];
                    // pass 0 has no prefix yet: selections with the same key share the histogram of
                    // the first of them (stage 0: all four; stage 1: pairs) -- 4x / 2x fewer atomics
                    if (pass == 0 && (STAGE == 0 ? s != 0 : (s & 1) != 0)) continue;
                    const bool match = (pass == 0) || ((kk >> hi_shift) == prefix[s]);
                    if (match) {
                        const unsigned int bin = (kk >> shift) & ((1u << nbits) - 1u);
                        if (uniform) atomicAdd(&h[s][bin], 1u);
                        else atomicAdd(hist + g_mine * (NSEL * NBINS) + s * NBINS + bin, 1u);
                    }
                }
            }
        }
    }
    flush(cur_group);
}

// ---- after each histogram pass: locate the bin holding each rank, extend the prefix, clear hist
__global__ void __launch_bounds__(NSEL * 32)
mac_select_kernel(MacGroup* __restrict__ grp, unsigned int* __restrict__ hist, int pass, int stage, int last) {
    MacGroup& g = grp[blockIdx.x];
    unsigned int* hg = hist + static_cast<long long>(blockIdx.x) * (NSEL * NBINS);
    const int s = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nbits = (pass == 2) ? 10 : 11;
    const int nb = 1 << nbits;
    if (g.valid) {
        // pass 0: shared histograms (see mac_hist_kernel)
        const int hsel = (pass == 0) ? (stage == 0 ? 0 : (s & 2)) : s;
        const unsigned int* hs = hg + hsel * NBINS;
        const int per = nb / 32;
        unsigned long long local = 0;
        for (int i = 0; i < per; ++i) local += hs[lane * per + i];
        unsigned long long incl = local;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        const unsigned long long excl = incl - local;
        const unsigned long long r = g.rank[s];
        const bool here = (r >= excl) && (r < incl);
        const unsigned int ballot = __ballot_sync(0xffffffffu, here);
        const int owner = ballot ? (__ffs(ballot) - 1) : 31;
        if (lane == owner) {
            unsigned long long cum = excl;
            int bin = lane * per + per - 1;
            for (int i = 0; i < per; ++i) {
                const unsigned long long c = hs[lane * per + i];
                if (r < cum + c) { bin = lane * per + i; break; }
                cum += c;
            }
            g.rank[s] = r - cum;
            unsigned int pre = (g.prefix[s] << nbits) | static_cast<unsigned int>(bin);
            // two radix passes resolve the 22 high key bits (2^-20 of the pseudo-angle range, 3e-5 in
            // concentration units -- far below what moves an output byte): the 10 low bits get their midpoint
            if (last && pass == 1) pre = (pre << 10) | 0x200u;
            g.prefix[s] = pre;
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < NSEL * NBINS; i += blockDim.x) hg[i] = 0;
}

// ---- stain vectors from the two angle percentiles; pseudo-inverse; ranks for the concentration stage
__global__ void mac_vectors_kernel(MacGroup* __restrict__ grp) {
    MacGroup& g = grp[blockIdx.x];
    if (threadIdx.x != 0 || !g.valid) return;
    const double a0 = pa_to_angle(g.prefix[0]), a1 = pa_to_angle(g.prefix[1]);
    const double b0 = pa_to_angle(g.prefix[2]), b1 = pa_to_angle(g.prefix[3]);
    const double min_phi = a0 + static_cast<double>(g.frac[0]) * (a1 - a0);
    const double max_phi = b0 + static_cast<double>(g.frac[1]) * (b1 - b0);
    double vmin[3], vmax[3];
    for (int c = 0; c < 3; ++c) {
        vmin[c] = g.E[c * 2] * cos(min_phi) + g.E[c * 2 + 1] * sin(min_phi);
        vmax[c] = g.E[c * 2] * cos(max_phi) + g.E[c * 2 + 1] * sin(max_phi);
    }
    double h[3], e[3];
    const bool min_is_h = vmin[0] > vmax[0];
    for (int c = 0; c < 3; ++c) { h[c] = min_is_h ? vmin[c] : vmax[c]; e[c] = min_is_h ? vmax[c] : vmin[c]; }
    // least squares C = pinv(HE) . OD with pinv = (HE^T HE)^-1 HE^T
    const double hh = h[0] * h[0] + h[1] * h[1] + h[2] * h[2];
    const double ee = e[0] * e[0] + e[1] * e[1] + e[2] * e[2];
    const double he = h[0] * e[0] + h[1] * e[1] + h[2] * e[2];
    const double det = hh * ee - he * he;
    for (int c = 0; c < 3; ++c) {
        g.HE[c * 2] = static_cast<float>(h[c]);
        g.HE[c * 2 + 1] = static_cast<float>(e[c]);
        g.pinv[c] = static_cast<float>((ee * h[c] - he * e[c]) / det);
        g.pinv[3 + c] = static_cast<float>((hh * e[c] - he * h[c]) / det);
    }
    set_ranks(g, g.n_all, 0.99, 0.99);
    // selections 0,1 -> concentration 0 (k, k+1); 2,3 -> concentration 1 (k, k+1): same ranks
}

__global__ void mac_finalize_kernel(MacGroup* __restrict__ grp, float* __restrict__ he_out,
                                    float* __restrict__ maxc_out, int* __restrict__ valid_out) {
    MacGroup& g = grp[blockIdx.x];
    if (threadIdx.x != 0) return;
    if (g.valid) {
        for (int s = 0; s < 2; ++s) {
            const double v0 = key_to_conc(g.prefix[2 * s]), v1 = key_to_conc(g.prefix[2 * s + 1]);
            g.maxC[s] = static_cast<float>(v0 + static_cast<double>(g.frac[0]) * (v1 - v0));
        }
        g.scale[0] = 1.9705f / g.maxC[0];
        g.scale[1] = 1.0308f / g.maxC[1];
    }
    if (he_out != nullptr) for (int i = 0; i < 6; ++i) he_out[blockIdx.x * 6 + i] = g.valid ? g.HE[i] : 0.f;
    if (maxc_out != nullptr) for (int i = 0; i < 2; ++i) maxc_out[blockIdx.x * 2 + i] = g.valid ? g.maxC[i] : 0.f;
    if (valid_out != nullptr) valid_out[blockIdx.x] = g.valid;
}

// ---- apply: I' = Io * exp(-HERef . (pinv . OD * scale)), clipped and truncated to uint8
__global__ void __launch_bounds__(MAC_THREADS)
mac_apply_kernel(const uint8_t* __restrict__ img, uint8_t* __restrict__ out, long long n_chunks,
                 long long chunks_per_group, float Io, const MacGroup* __restrict__ grp) {
    __shared__ float lut[256];
    load_lut(lut, Io);
    __syncthreads();
    for (long long c = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; c < n_chunks;
         c += static_cast<long long>(gridDim.x) * blockDim.x) {
        const MacGroup& G = grp[c / chunks_per_group];
        uint8_t px[48];
        load_pixels(img + c * 48, px);
        if (G.valid) {
            const float p0 = G.pinv[0] * G.scale[0], p1 = G.pinv[1] * G.scale[0], p2 = G.pinv[2] * G.scale[0];
            const float p3 = G.pinv[3] * G.scale[1], p4 = G.pinv[4] * G.scale[1], p5 = G.pinv[5] * G.scale[1];
#pragma unroll
            for (int k = 0; k < PIX_PER_ITER; ++k) {
                const float r = lut[px[3 * k]], g = lut[px[3 * k + 1]], b = lut[px[3 * k + 2]];
                const float c0 = p0 * r + p1 * g + p2 * b;
                const float c1 = p3 * r + p4 * g + p5 * b;
                const float o0 = Io * expf(-(0.5626f * c0 + 0.2159f * c1));
                const float o1 = Io * expf(-(0.7201f * c0 + 0.8012f * c1));
                const float o2 = Io * expf(-(0.4062f * c0 + 0.5581f * c1));
                px[3 * k] = static_cast<uint8_t>(fminf(fmaxf(o0, 0.f), 255.f));
                px[3 * k + 1] = static_cast<uint8_t>(fminf(fmaxf(o1, 0.f), 255.f));
                px[3 * k + 2] = static_cast<uint8_t>(fminf(fmaxf(o2, 0.f), 255.f));
            }
        }
        uint4* o = reinterpret_cast<uint4*>(out + c * 48);
        o[0] = *reinterpret_cast<uint4*>(px);
        o[1] = *reinterpret_cast<uint4*>(px + 16);
        o[2] = *reinterpret_cast<uint4*>(px + 32);
    }
}

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace
}  // namespace sb

extern "C" {

size_t stamp_macenko_workspace_bytes(int n_tiles, int tiles_per_fit) {
    if (n_tiles <= 0) return 0;
    const int tpf = tiles_per_fit > 0 ? tiles_per_fit : n_tiles;
    const size_t G = (static_cast<size_t>(n_tiles) + tpf - 1) / tpf;
    return sb::align_up(G * sizeof(sb::MacGroup), 256) + G * sb::NSEL * sb::NBINS * sizeof(unsigned int);
}

int stamp_macenko_u8(const uint8_t* in, uint8_t* out, int n_tiles, int H, int W, int tiles_per_fit,
                     float Io, float alpha, float beta, float* he_out, float* maxc_out,
                     int* valid_out, void* workspace, size_t workspace_bytes, void* stream_) {
    using namespace sb;
    if (in == nullptr || out == nullptr || workspace == nullptr || n_tiles <= 0 || H <= 0 || W <= 0 ||
        Io <= 1.f || alpha < 0.f || alpha > 50.f)
        return SB_ERR_BAD_ARG;
    const long long tile_bytes = static_cast<long long>(H) * W * 3;
    if (tile_bytes % 48 != 0 || (reinterpret_cast<uintptr_t>(in) & 15) != 0 ||
        (reinterpret_cast<uintptr_t>(out) & 15) != 0 || (reinterpret_cast<uintptr_t>(workspace) & 255) != 0)
        return SB_ERR_BAD_ARG;
    const int tpf = tiles_per_fit > 0 ? tiles_per_fit : n_tiles;
    const int G = (n_tiles + tpf - 1) / tpf;
    if (workspace_bytes < stamp_macenko_workspace_bytes(n_tiles, tiles_per_fit)) return SB_ERR_WORKSPACE;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    MacGroup* grp = static_cast<MacGroup*>(workspace);
    unsigned int* hist = reinterpret_cast<unsigned int*>(static_cast<uint8_t*>(workspace) +
                                                         align_up(G * sizeof(MacGroup), 256));
    const long long n_chunks = static_cast<long long>(n_tiles) * (tile_bytes / 48);
    const long long chunks_per_group = static_cast<long long>(tpf) * (tile_bytes / 48);
    const long long n_pixels = n_chunks * PIX_PER_ITER, pixels_per_group = chunks_per_group * PIX_PER_ITER;
    const double bytes = static_cast<double>(n_tiles) * tile_bytes;

    if (cudaMemsetAsync(workspace, 0, stamp_macenko_workspace_bytes(n_tiles, tiles_per_fit), stream) != cudaSuccess)
        return SB_ERR_CUDA;
    const int sms = 148;
    const int grid = static_cast<int>(min(static_cast<long long>(sms) * 4, (n_chunks + MAC_THREADS - 1) / MAC_THREADS));
    {
        ProfScope prof(PROF_MACENKO, 2.0 * bytes, stream);  // algorithmic traffic: one read + one write
        mac_stats_kernel<<<grid, MAC_THREADS, 0, stream>>>(in, n_chunks, chunks_per_group, Io, beta, grp);
        mac_eig_kernel<<<G, 32, 0, stream>>>(grp, pixels_per_group, n_pixels, alpha);
        for (int pass = 0; pass < MAC_PASSES; ++pass) {
            mac_hist_kernel<0><<<grid, MAC_THREADS, 0, stream>>>(in, n_chunks, chunks_per_group, Io, beta, grp, hist, pass);
            mac_select_kernel<<<G, NSEL * 32, 0, stream>>>(grp, hist, pass, 0, pass == MAC_PASSES - 1);
        }
        mac_vectors_kernel<<<G, 32, 0, stream>>>(grp);
        for (int pass = 0; pass < MAC_PASSES; ++pass) {
            mac_hist_kernel<1><<<grid, MAC_THREADS, 0, stream>>>(in, n_chunks, chunks_per_group, Io, beta, grp, hist, pass);
            mac_select_kernel<<<G, NSEL * 32, 0, stream>>>(grp, hist, pass, 1, pass == MAC_PASSES - 1);
        }
        mac_finalize_kernel<<<G, 32, 0, stream>>>(grp, he_out, maxc_out, valid_out);
        const int agrid = static_cast<int>(min(static_cast<long long>(sms) * 8, (n_chunks + MAC_THREADS - 1) / MAC_THREADS));
        mac_apply_kernel<<<agrid, MAC_THREADS, 0, stream>>>(in, out, n_chunks, chunks_per_group, Io, grp);
        count_launch(5 + 4 * MAC_PASSES);
    }
    return cudaGetLastError() == cudaSuccess ? SB_OK : SB_ERR_CUDA;
}

}  // extern "C"
