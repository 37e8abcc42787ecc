// Macenko stain normalisation over a batch of uint8 RGB tiles, entirely on the device.
//
// The reference snapshot has no Macenko code (README.md:35 is the only mention); the stage is named
// by BASELINE.json's north_star and specified in SURVEY.md 8c / oracle/macenko_oracle.py.
//
// Byte work bound by instruction issue, not by HBM: six passes stream the uint8 pixels (a batch of <= ~800 tiles
// stays in the 126 MB L2 after the first), so what counts is the number of instructions -- and shared-memory
// wavefronts -- per pixel.  Optical density is a function of the byte value, OD = ln Io - ln 2 * log2(v + 1), and
// everything a pass needs from a pixel is LINEAR in the three log2 values L_c: one PRMT + FADD + MUFU.LG2 per channel
// (the byte is dropped into the mantissa of 2^23), then a few FFMAs against per-group coefficients that have all
// constants folded in.  (Table lookups instead of the MUFU cost 3.5-6 shared-memory wavefronts per gather on
// uncorrelated bytes and made every pass shared-memory bound at ~80 us; measured, profiles/r2_macenko.md.)
//
//   stats   : sums / products of L over the tissue pixels (largest byte <= threshold byte, the exact logf rule)
//             -> fp64; cov(OD) = ln2^2 cov(L): same eigenvectors.  The last CTA of a group (ticket counter)
//             does the 3x3 eigen-decomposition
//   angle   : t_j = E_j . OD  (6 FFMA); pseudo-angle -> 23-bit order-preserving key from the mantissa of (pa + 12.0f)
//   conc    : y_s = pinv_s . OD * kscale + 12  (6 FFMA) -> key from the mantissa
//   apply   : log2 out_k = Z_k + sum_c M_kc L_c  (9 FFMA, M = HERef . diag(scale) . pinv) -> ex2, truncation with add.rz
//
// The four exact percentiles (two of the stain angle, two of the concentrations, each needing the k-th and (k+1)-th
// order statistic for NumPy's linear interpolation) come from a 2-pass radix select (11 + 11 key bits) over
// shared-memory histograms: no sort, no per-pixel intermediate in HBM.  The bin search after a histogram pass is
// done by the last CTA of the pass, so one call is six launches:
//
//   stats(+eig) -> angle p0(+select) -> angle p1(+select, stain vectors) -> conc p0 -> conc p1(+finalize) -> apply
#include <math.h>

#include <algorithm>

#include "common.cuh"
#include "stamp_b200.h"

namespace sb {
namespace {

constexpr int NSEL = 4;          // order statistics searched concurrently per stage
constexpr int NBINS = 2048;
constexpr int MAC_THREADS = 256;
constexpr int PIX_PER_ITER = 16;  // 48 bytes
constexpr int MIN_TISSUE = 16;
// channels (counted from blue) whose log2 comes from the shared-memory table instead of the MUFU, per pass
constexpr int STATS_NLUT = 0, ANGLE_NLUT = 1, CONC_NLUT = 0, APPLY_NLUT = 2;

struct MacGroup {
    double sum[3];
    double sq[6];                 // xx xy xz yy yz zz
    unsigned long long n_kept;
    unsigned long long n_all;
    unsigned long long rank[NSEL];
    unsigned int prefix[NSEL];    // after pass 0: the 11 high key bits; after pass 1: the 22 high key bits
    float frac[2];                // interpolation weights of the two percentiles of the stage
    float E[6];                   // E[c*2 + j]: plane basis, j = 0 second-largest, 1 largest eigenvector
    float pinv[6];                // pinv[s*3 + c]
    float HE[6];                  // HE[c*2 + s], s = 0 haematoxylin, 1 eosin
    float maxC[2];
    float scale[2];               // maxCRef / maxC
    float kscale[2];              // concentration -> key: x = c * kscale + 4
    unsigned int ticket[5];       // CTAs of the group that finished the reducing pass i
    int valid;
};

// ---- order-preserving keys -------------------------------------------------------------------------------------
// x in [0, 8) -> K = round(x * 2^20), read off the mantissa of x + 8.  A radix pass resolves 11 bits: pass 0 bits
// 22..12, pass 1 bits 11..1; bit 0 is left at its midpoint.
// Stain angle: atan2 is replaced per pixel by the pseudo-angle pa(x, y) in [-2, 2], strictly monotone in
// atan2(y, x), so both have the same order statistics; x = pa + 4, resolved to 2^-19 (4e-6 rad), and the four
// selected keys are mapped back to radians in fp64.  Concentrations: x = c * kscale + 4 with kscale chosen per
// stain from the bound |c| <= sum|pinv| * max OD, so that the key cannot leave its range.
__device__ __forceinline__ float pseudo_angle(float x, float y) {
    const float ay = fabsf(y);
    const float r = __fdividef(ay, fabsf(x) + ay + 1e-30f);
    return copysignf(x >= 0.f ? r : 2.0f - r, y);
}
__device__ __forceinline__ unsigned int angle_key(float pa) { return __float_as_uint(pa + 12.0f) & 0x7fffffu; }
__device__ double key22_to_x(unsigned int key22) { return (static_cast<double>(key22) * 2.0 + 0.5) / 1048576.0; }
// (fp32 trigonometry: the key resolves the angle to 4e-6 rad, atan2f / sincosf are good to 2e-7)
__device__ float key22_to_angle(unsigned int key22) {
    const float pa = static_cast<float>(key22_to_x(key22) - 4.0);
    const float a = fabsf(pa);
    const float r = a <= 1.0f ? a : 2.0f - a;
    const float ang = atan2f(r, a <= 1.0f ? 1.0f - r : -(1.0f - r));
    return pa < 0 ? -ang : ang;
}

__device__ __forceinline__ float od_of(int v, float Io) { return -logf((static_cast<float>(v) + 1.0f) / Io); }
__device__ __forceinline__ float lg2_approx(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;\n" : "=f"(y) : "f"(x));
    return y;
}

// 16 pixels (48 bytes)
__device__ __forceinline__ void load_pixels(const uint8_t* p, uint32_t (&w)[12]) {
    const uint4 a = ld_nc_v4(p), b = ld_nc_v4(p + 16), c = ld_nc_v4(p + 32);
    w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w;
    w[4] = b.x; w[5] = b.y; w[6] = b.z; w[7] = b.w;
    w[8] = c.x; w[9] = c.y; w[10] = c.z; w[11] = c.w;
}
// byte j (compile-time) of the 48 as the float 2^23 + v: the byte becomes the low mantissa byte of 0x4B000000
template <int J>
__device__ __forceinline__ float biased_byte(const uint32_t (&w)[12]) {
    return __uint_as_float(__byte_perm(w[J >> 2], 0x4B000000u, 0x7650u + (J & 3)));
}
constexpr float BYTE_BIAS = 8388608.0f;
// log2(v + 1)
__device__ __forceinline__ float lg2_byte(float biased) { return lg2_approx(biased - (BYTE_BIAS - 1.0f)); }
// log2(v + 1) of channel CH: MUFU for the first 3 - NLUT channels, a 256-entry shared-memory table for the last
// NLUT ones.  The MUFU pipe (16 lanes/clk/SM) is the busiest unit of every pass; a gather of uncorrelated bytes costs
// ~3.5 shared-memory wavefronts, so moving one or two channels over balances the two units.
template <int CH, int NLUT>
__device__ __forceinline__ float lg2_ch(const float* ltab, float biased) {
    if (CH >= 3 - NLUT) return ltab[__float_as_uint(biased) & 0xffu];
    return lg2_byte(biased);
}
__device__ __forceinline__ void fill_lg2_table(float* ltab) {        // 256 threads; the caller synchronises
    ltab[threadIdx.x] = lg2_approx(static_cast<float>(threadIdx.x) + 1.0f);
}

// Work split: grid = (CTAs per group, groups).  A CTA never leaves its fit group.
struct Span {
    const uint8_t* base;      // first chunk of the group
    long long chunks;         // chunks of the group
};
__device__ __forceinline__ Span group_span(const uint8_t* img, long long n_chunks, long long chunks_per_group) {
    const long long first = static_cast<long long>(blockIdx.y) * chunks_per_group;
    Span s;
    s.base = img + first * 48;
    s.chunks = min(chunks_per_group, n_chunks - first);
    return s;
}

// Grid-stride loop over the 48-byte chunks of the CTA's group with the next chunk's three 128-bit loads in flight
// while the current one is processed.
template <typename F>
__device__ __forceinline__ void for_each_chunk(const Span& sp, F&& body) {
    const long long stride = static_cast<long long>(gridDim.x) * MAC_THREADS;
    long long c = blockIdx.x * static_cast<long long>(MAC_THREADS) + threadIdx.x;
    if (c >= sp.chunks) return;
    uint32_t w[12];
    load_pixels(sp.base + c * 48, w);
    while (true) {
        const long long nxt = c + stride;
        const bool more = nxt < sp.chunks;
        uint32_t wn[12];
        if (more) load_pixels(sp.base + nxt * 48, wn);
        body(w, c);
        if (!more) break;
#pragma unroll
        for (int i = 0; i < 12; ++i) w[i] = wn[i];
        c = nxt;
    }
}

// largest byte value v with OD(v) >= beta under the oracle's float rule, as 2^23 + v (2^23 - 1: none).  All 256
// threads of the CTA call it; it is a barrier.
__device__ __forceinline__ float tissue_threshold(float Io, float beta) {
    return BYTE_BIAS + static_cast<float>(__syncthreads_count(od_of(threadIdx.x, Io) >= beta) - 1);
}

// true in exactly one CTA of the group: the one that arrives last (its reads then see every other CTA's results)
__device__ bool last_cta_of_group(unsigned int* ticket) {
    __shared__ bool is_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) is_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    __syncthreads();
    if (is_last) __threadfence();
    return is_last;
}

// ---- 3x3 symmetric eigen-decomposition (cyclic Jacobi), plane basis, ranks for the angle stage.
// The covariance is formed in fp64 (sum of squares minus n * mean^2 cancels); the rotations run in fp32: eigenvector
// error ~ 6e-8 * |A| / gap, far below the 1e-4 the stain vectors are compared at, and one thread of one CTA does this
// while the GPU waits -- fp64 divisions and square roots made it a 38 us tail.
__device__ void jacobi3(float A[3][3], float V[3][3]) {
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) V[i][j] = (i == j) ? 1.0f : 0.0f;
    const float tiny = 1e-9f * (fabsf(A[0][0]) + fabsf(A[1][1]) + fabsf(A[2][2])) + 1e-37f;
    for (int sweep = 0; sweep < 12; ++sweep) {
        const float off = fabsf(A[0][1]) + fabsf(A[0][2]) + fabsf(A[1][2]);
        if (off < tiny) break;          // quadratic convergence: 4-5 sweeps
        for (int p = 0; p < 2; ++p)
            for (int q = p + 1; q < 3; ++q) {
                if (fabsf(A[p][q]) < 1e-37f) continue;
                const float theta = (A[q][q] - A[p][p]) / (2.0f * A[p][q]);
                const float t = (theta >= 0 ? 1.0f : -1.0f) / (fabsf(theta) + sqrtf(theta * theta + 1.0f));
                const float c = 1.0f / sqrtf(t * t + 1.0f), sn = t * c;
                for (int k = 0; k < 3; ++k) {
                    const float akp = A[k][p], akq = A[k][q];
                    A[k][p] = c * akp - sn * akq;
                    A[k][q] = sn * akp + c * akq;
                }
                for (int k = 0; k < 3; ++k) {
                    const float apk = A[p][k], aqk = A[q][k];
                    A[p][k] = c * apk - sn * aqk;
                    A[q][k] = sn * apk + c * aqk;
                }
                for (int k = 0; k < 3; ++k) {
                    const float vkp = V[k][p], vkq = V[k][q];
                    V[k][p] = c * vkp - sn * vkq;
                    V[k][q] = sn * vkp + c * vkq;
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

// one thread, after every CTA of the group has added its sums
__device__ __noinline__ void eig_of_group(MacGroup& g, unsigned long long n_all, float alpha) {
    const volatile MacGroup& v = g;
    g.n_all = n_all;
    const unsigned long long kept = v.n_kept;
    const double n = static_cast<double>(kept);
    g.valid = kept >= MIN_TISSUE;
    if (kept < MIN_TISSUE) return;
    const double sm[3] = {v.sum[0], v.sum[1], v.sum[2]};
    const double sq[6] = {v.sq[0], v.sq[1], v.sq[2], v.sq[3], v.sq[4], v.sq[5]};
    const double mu[3] = {sm[0] / n, sm[1] / n, sm[2] / n};
    float A[3][3], V[3][3];
    const int idx[3][3] = {{0, 1, 2}, {1, 3, 4}, {2, 4, 5}};
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) A[i][j] = static_cast<float>((sq[idx[i][j]] - n * mu[i] * mu[j]) / (n - 1.0));
    jacobi3(A, V);
    int order[3] = {0, 1, 2};  // ascending eigenvalues
    for (int i = 0; i < 2; ++i)
        for (int j = 0; j < 2 - i; ++j)
            if (A[order[j]][order[j]] > A[order[j + 1]][order[j + 1]]) { int t = order[j]; order[j] = order[j + 1]; order[j + 1] = t; }
    float e0[3], e1[3];  // second largest, largest
    for (int c = 0; c < 3; ++c) { e0[c] = V[c][order[1]]; e1[c] = V[c][order[2]]; }
    if (e1[0] + e1[1] + e1[2] < 0) for (int c = 0; c < 3; ++c) e1[c] = -e1[c];
    if (e0[0] < 0) for (int c = 0; c < 3; ++c) e0[c] = -e0[c];
    for (int c = 0; c < 3; ++c) { g.E[c * 2] = e0[c]; g.E[c * 2 + 1] = e1[c]; }
    set_ranks(g, kept, alpha / 100.0, 1.0 - alpha / 100.0);
}

// ---- pass 1: per-group sums for the covariance of the tissue optical densities ----------------
__global__ void __launch_bounds__(MAC_THREADS, 4)
mac_stats_kernel(const uint8_t* __restrict__ img, long long n_chunks, long long chunks_per_group,
                 float Io, float beta, float alpha, MacGroup* __restrict__ grp) {
    __shared__ double red[MAC_THREADS / 32][10];
    __shared__ float ltab[256];
    fill_lg2_table(ltab);
    const float vthr = tissue_threshold(Io, beta);
    const Span sp = group_span(img, n_chunks, chunks_per_group);
    double acc[10];     // sums over the tissue pixels of L_c = log2(v_c + 1), of their products, and the count
#pragma unroll
    for (int i = 0; i < 10; ++i) acc[i] = 0.0;
    for_each_chunk(sp, [&](const uint32_t (&w)[12], long long) {
        float f[10];
#pragma unroll
        for (int i = 0; i < 10; ++i) f[i] = 0.f;
        auto pixel = [&](float vr, float vg, float vb) {
            const float r = lg2_ch<0, STATS_NLUT>(ltab, vr), g = lg2_ch<1, STATS_NLUT>(ltab, vg), b = lg2_ch<2, STATS_NLUT>(ltab, vb);
            if (fmaxf(fmaxf(vr, vg), vb) <= vthr) {
                f[0] += r; f[1] += g; f[2] += b;
                f[3] = fmaf(r, r, f[3]); f[4] = fmaf(r, g, f[4]); f[5] = fmaf(r, b, f[5]);
                f[6] = fmaf(g, g, f[6]); f[7] = fmaf(g, b, f[7]); f[8] = fmaf(b, b, f[8]);
                f[9] += 1.f;
            }
        };
#define SB_PIX(K) pixel(biased_byte<3 * (K)>(w), biased_byte<3 * (K) + 1>(w), biased_byte<3 * (K) + 2>(w));
        SB_PIX(0) SB_PIX(1) SB_PIX(2) SB_PIX(3) SB_PIX(4) SB_PIX(5) SB_PIX(6) SB_PIX(7)
        SB_PIX(8) SB_PIX(9) SB_PIX(10) SB_PIX(11) SB_PIX(12) SB_PIX(13) SB_PIX(14) SB_PIX(15)
#pragma unroll
        for (int i = 0; i < 10; ++i) acc[i] += static_cast<double>(f[i]);
    });
#pragma unroll
    for (int i = 0; i < 10; ++i) acc[i] = warp_sum_d(acc[i]);
    const int wid = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0)
        for (int i = 0; i < 10; ++i) red[wid][i] = acc[i];
    __syncthreads();
    MacGroup& G = grp[blockIdx.y];
    if (threadIdx.x < 10) {
        double t = 0.0;
        for (int k = 0; k < MAC_THREADS / 32; ++k) t += red[k][threadIdx.x];
        if (threadIdx.x < 3) atomicAdd(&G.sum[threadIdx.x], t);
        else if (threadIdx.x < 9) atomicAdd(&G.sq[threadIdx.x - 3], t);
        else atomicAdd(&G.n_kept, static_cast<unsigned long long>(t + 0.5));
    }
    if (last_cta_of_group(&G.ticket[0]) && threadIdx.x == 0)
        eig_of_group(G, static_cast<unsigned long long>(sp.chunks) * PIX_PER_ITER, alpha);
}

// ---- stain vectors from the two angle percentiles; pseudo-inverse; key scale and ranks for the concentrations
__device__ __noinline__ void vectors_of_group(MacGroup& g, float Io) {
    const float a0 = key22_to_angle(g.prefix[0]), a1 = key22_to_angle(g.prefix[1]);
    const float b0 = key22_to_angle(g.prefix[2]), b1 = key22_to_angle(g.prefix[3]);
    const float min_phi = a0 + g.frac[0] * (a1 - a0);
    const float max_phi = b0 + g.frac[1] * (b1 - b0);
    float smin, cmin, smax, cmax;
    sincosf(min_phi, &smin, &cmin);
    sincosf(max_phi, &smax, &cmax);
    double vmin[3], vmax[3];
    for (int c = 0; c < 3; ++c) {
        vmin[c] = g.E[c * 2] * cmin + g.E[c * 2 + 1] * smin;
        vmax[c] = g.E[c * 2] * cmax + g.E[c * 2 + 1] * smax;
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
    const float od_max = od_of(0, Io);
    for (int s = 0; s < 2; ++s) {
        const float bound = (fabsf(g.pinv[3 * s]) + fabsf(g.pinv[3 * s + 1]) + fabsf(g.pinv[3 * s + 2])) * od_max;
        g.kscale[s] = 3.9f / fmaxf(bound, 1e-20f);
    }
    set_ranks(g, g.n_all, 0.99, 0.99);
    // selections 0,1 -> concentration 0 (k, k+1); 2,3 -> concentration 1 (k, k+1): same ranks
}

__device__ __noinline__ void finalize_group(MacGroup& g, int gi, float* he_out, float* maxc_out, int* valid_out) {
    if (g.valid) {
        for (int s = 0; s < 2; ++s) {
            const double ks = static_cast<double>(g.kscale[s]);
            const double v0 = (key22_to_x(g.prefix[2 * s]) - 4.0) / ks, v1 = (key22_to_x(g.prefix[2 * s + 1]) - 4.0) / ks;
            g.maxC[s] = static_cast<float>(v0 + static_cast<double>(g.frac[0]) * (v1 - v0));
        }
        g.scale[0] = 1.9705f / g.maxC[0];
        g.scale[1] = 1.0308f / g.maxC[1];
    }
    if (he_out != nullptr) for (int i = 0; i < 6; ++i) he_out[gi * 6 + i] = g.valid ? g.HE[i] : 0.f;
    if (maxc_out != nullptr) for (int i = 0; i < 2; ++i) maxc_out[gi * 2 + i] = g.valid ? g.maxC[i] : 0.f;
    if (valid_out != nullptr) valid_out[gi] = g.valid;
}

// groups without tissue: the fit outputs are still written (the histogram passes return at once for them)
__global__ void mac_invalid_outputs_kernel(MacGroup* __restrict__ grp, int G, float* he_out, float* maxc_out,
                                           int* valid_out) {
    const int gi = blockIdx.x * blockDim.x + threadIdx.x;
    if (gi < G && !grp[gi].valid) finalize_group(grp[gi], gi, he_out, maxc_out, valid_out);
}

// ---- radix-select histogram pass (stage 0: stain angle of tissue pixels; stage 1: concentrations)
// pass 0 has no prefix yet: selections with the same key share one histogram (stage 0: all four; stage 1: pairs)
template <int STAGE, int PASS>
__global__ void __launch_bounds__(MAC_THREADS, 4)
mac_hist_kernel(const uint8_t* __restrict__ img, long long n_chunks, long long chunks_per_group,
                float Io, float beta, MacGroup* __restrict__ grp, unsigned int* __restrict__ hist,
                float* he_out, float* maxc_out, int* valid_out) {
    constexpr int NH = PASS == 0 ? (STAGE == 0 ? 1 : 2) : NSEL;
    constexpr int NLUT = STAGE == 0 ? ANGLE_NLUT : CONC_NLUT;
    __shared__ unsigned int h[NSEL][NBINS];     // (NH of them used by the pixel loop, all four by the bin search)
    __shared__ float ltab[256];
    MacGroup& G = grp[blockIdx.y];
    if (!G.valid) return;
    fill_lg2_table(ltab);
    const float vthr = tissue_threshold(Io, beta);
    // v = k0 + k . L for the two projections (stage 0) / the two concentration keys (stage 1), L = log2(byte + 1):
    // OD_c = ln Io - ln2 L_c folded in
    float k[2][3], k0[2];
    {
        const float ln_io = logf(Io), ln2 = 0.6931471805599453f;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const float m = STAGE == 0 ? 1.0f : G.kscale[j];
            float sum = 0.f;
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) {
                const float coef = (STAGE == 0 ? G.E[ch * 2 + j] : G.pinv[3 * j + ch]) * m;
                k[j][ch] = -ln2 * coef;
                sum += coef;
            }
            k0[j] = ln_io * sum + (STAGE == 0 ? 0.f : 12.0f);      // keys: x = c * kscale + 4, read off x + 8
        }
    }
    for (int i = threadIdx.x; i < NH * NBINS; i += MAC_THREADS) (&h[0][0])[i] = 0;
    unsigned int prefix[NSEL];
#pragma unroll
    for (int i = 0; i < NSEL; ++i) prefix[i] = PASS == 0 ? 0u : G.prefix[i];
    __syncthreads();

    const Span sp = group_span(img, n_chunks, chunks_per_group);
    for_each_chunk(sp, [&](const uint32_t (&w)[12], long long) {
        auto count = [&](unsigned int key, int s0) {      // key: 23 bits
            const unsigned int top = key >> 12;
            if (PASS == 0) {
                atomicAdd(&h[s0 >> (STAGE == 0 ? 2 : 1)][top], 1u);
            } else {
                // the k-th and (k+1)-th order statistic are neighbours: no key lies strictly between their bins,
                // so one range test per pair filters (almost) every pixel out
                bool near = top - prefix[s0] <= prefix[s0 + 1] - prefix[s0];
                if (STAGE == 0) near |= top - prefix[2] <= prefix[3] - prefix[2];
                if (near) {
                    const unsigned int bin = (key >> 1) & (NBINS - 1);
#pragma unroll
                    for (int s = 0; s < (STAGE == 0 ? 4 : 2); ++s)
                        if (top == prefix[s0 + s]) atomicAdd(&h[s0 + s][bin], 1u);
                }
            }
        };
        // (the same instruction sequence in pass 0 and pass 1: a pixel must produce the same key in both)
        auto pixel = [&](float vr, float vg, float vb) {
            const float r = lg2_ch<0, NLUT>(ltab, vr), g = lg2_ch<1, NLUT>(ltab, vg), b = lg2_ch<2, NLUT>(ltab, vb);
            const float v0 = fmaf(k[0][0], r, fmaf(k[0][1], g, fmaf(k[0][2], b, k0[0])));
            const float v1 = fmaf(k[1][0], r, fmaf(k[1][1], g, fmaf(k[1][2], b, k0[1])));
            if (STAGE == 0) {
                if (fmaxf(fmaxf(vr, vg), vb) <= vthr) count(angle_key(pseudo_angle(v0, v1)), 0);
            } else {
                count(__float_as_uint(v0) & 0x7fffffu, 0);
                count(__float_as_uint(v1) & 0x7fffffu, 2);
            }
        };
        SB_PIX(0) SB_PIX(1) SB_PIX(2) SB_PIX(3) SB_PIX(4) SB_PIX(5) SB_PIX(6) SB_PIX(7)
        SB_PIX(8) SB_PIX(9) SB_PIX(10) SB_PIX(11) SB_PIX(12) SB_PIX(13) SB_PIX(14) SB_PIX(15)
    });
    __syncthreads();
    unsigned int* hg = hist + static_cast<long long>(blockIdx.y) * (NSEL * NBINS);
    for (int i = threadIdx.x; i < NH * NBINS; i += MAC_THREADS) {
        const unsigned int v = (&h[0][0])[i];
        if (v) atomicAdd(hg + i, v);
    }
    if (!last_cta_of_group(&G.ticket[1 + STAGE * 2 + PASS])) return;

    // ---- bin search by the last CTA: the group's histograms -> shared memory, one warp per order statistic ----
    for (int i = threadIdx.x; i < NH * NBINS; i += MAC_THREADS) {
        (&h[0][0])[i] = __ldcg(hg + i);
        hg[i] = 0;                                   // ready for the next pass
    }
    __syncthreads();
    const int s = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (s < NSEL) {
        const unsigned int* hs = h[PASS == 0 ? (STAGE == 0 ? 0 : (s >> 1)) : s];
        constexpr int per = NBINS / 32;
        unsigned long long local = 0;
        for (int i = 0; i < per; ++i) local += hs[lane * per + i];
        unsigned long long incl = local;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        const unsigned long long excl = incl - local;
        const unsigned long long r = G.rank[s];
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
            G.rank[s] = r - cum;
            G.prefix[s] = (prefix[s] << 11) | static_cast<unsigned int>(bin);
        }
    }
    if (PASS == 1) {
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) {
            if (STAGE == 0) vectors_of_group(G, Io);
            else finalize_group(G, blockIdx.y, he_out, maxc_out, valid_out);
        }
    }
}

// ---- apply: I' = Io * exp(-HERef . (pinv . OD * scale)), clipped and truncated to uint8
__global__ void __launch_bounds__(MAC_THREADS)
mac_apply_kernel(const uint8_t* __restrict__ img, uint8_t* __restrict__ out, long long n_chunks,
                 long long chunks_per_group, float Io, const MacGroup* __restrict__ grp) {
    __shared__ float ltab[256];
    fill_lg2_table(ltab);
    __syncthreads();
    const MacGroup& G = grp[blockIdx.y];
    const bool valid = G.valid != 0;
    const Span sp = group_span(img, n_chunks, chunks_per_group);
    uint8_t* obase = out + (sp.base - img);
    // log2 out_k = log2 Io - log2e * HERef_k . (scale * pinv . OD),  OD_c = ln Io - ln2 L_c
    //            = Z_k + sum_c M_kc L_c,   M_kc = sum_s HERef_ks scale_s pinv_sc,  Z_k = log2 Io * (1 - sum_c M_kc)
    float M[3][3], Z[3];
    if (valid) {
        const float href[3][2] = {{0.5626f, 0.2159f}, {0.7201f, 0.8012f}, {0.4062f, 0.5581f}};
        const float l2io = log2f(Io);
#pragma unroll
        for (int kk = 0; kk < 3; ++kk) {
            float sum = 0.f;
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) {
                M[kk][ch] = href[kk][0] * G.scale[0] * G.pinv[ch] + href[kk][1] * G.scale[1] * G.pinv[3 + ch];
                sum += M[kk][ch];
            }
            Z[kk] = l2io * (1.0f - sum);
        }
    }
    for_each_chunk(sp, [&](const uint32_t (&win)[12], long long c) {
        uint32_t w[12];
#pragma unroll
        for (int i = 0; i < 12; ++i) w[i] = win[i];
        if (valid) {
            uint32_t y[48];
            auto pixel = [&](float vr, float vg, float vb, uint32_t* yo) {
                const float r = lg2_ch<0, APPLY_NLUT>(ltab, vr), g = lg2_ch<1, APPLY_NLUT>(ltab, vg), b = lg2_ch<2, APPLY_NLUT>(ltab, vb);
#pragma unroll
                for (int kk = 0; kk < 3; ++kk) {
                    const float e = fmaf(M[kk][0], r, fmaf(M[kk][1], g, fmaf(M[kk][2], b, Z[kk])));
                    // exp2 > 0; truncation = round-toward-zero add of 2^23: the byte is the low mantissa byte
                    yo[kk] = __float_as_uint(__fadd_rz(fminf(ex2_approx(e), 255.f), BYTE_BIAS));
                }
            };
#define SB_APX(K) pixel(biased_byte<3 * (K)>(w), biased_byte<3 * (K) + 1>(w), biased_byte<3 * (K) + 2>(w), y + 3 * (K));
            SB_APX(0) SB_APX(1) SB_APX(2) SB_APX(3) SB_APX(4) SB_APX(5) SB_APX(6) SB_APX(7)
            SB_APX(8) SB_APX(9) SB_APX(10) SB_APX(11) SB_APX(12) SB_APX(13) SB_APX(14) SB_APX(15)
#undef SB_APX
#pragma unroll
            for (int i = 0; i < 12; ++i)
                w[i] = __byte_perm(__byte_perm(y[4 * i], y[4 * i + 1], 0x0040u),
                                   __byte_perm(y[4 * i + 2], y[4 * i + 3], 0x0040u), 0x5410u);
        }
        uint4* o = reinterpret_cast<uint4*>(obase + c * 48);
        o[0] = make_uint4(w[0], w[1], w[2], w[3]);
        o[1] = make_uint4(w[4], w[5], w[6], w[7]);
        o[2] = make_uint4(w[8], w[9], w[10], w[11]);
    });
}
#undef SB_PIX

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
    if (G > 65535) return SB_ERR_BAD_ARG;      // groups ride on gridDim.y
    if (workspace_bytes < stamp_macenko_workspace_bytes(n_tiles, tiles_per_fit)) return SB_ERR_WORKSPACE;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    MacGroup* grp = static_cast<MacGroup*>(workspace);
    unsigned int* hist = reinterpret_cast<unsigned int*>(static_cast<uint8_t*>(workspace) +
                                                         align_up(G * sizeof(MacGroup), 256));
    const long long n_chunks = static_cast<long long>(n_tiles) * (tile_bytes / 48);
    const long long chunks_per_group = static_cast<long long>(tpf) * (tile_bytes / 48);
    const double bytes = static_cast<double>(n_tiles) * tile_bytes;

    if (cudaMemsetAsync(workspace, 0, stamp_macenko_workspace_bytes(n_tiles, tiles_per_fit), stream) != cudaSuccess)
        return SB_ERR_CUDA;
    // grid = (CTAs per fit group, groups): about 4 CTAs per SM in total (the histogram kernels keep 38 KB of
    // shared memory each), never more CTAs than 256-chunk stripes in a group
    const int sms = 148;
    const long long stripes = (chunks_per_group + MAC_THREADS - 1) / MAC_THREADS;
    const int per_group = static_cast<int>(std::max(1LL, std::min(stripes, static_cast<long long>(sms * 4 / G))));
    const dim3 grid(per_group, G);
    const int per_group_apply = static_cast<int>(std::max(1LL, std::min(stripes, static_cast<long long>(sms * 8 / G))));
    {
        ProfScope prof(PROF_MACENKO, 2.0 * bytes, stream);  // algorithmic traffic: one read + one write
        mac_stats_kernel<<<grid, MAC_THREADS, 0, stream>>>(in, n_chunks, chunks_per_group, Io, beta, alpha, grp);
        mac_hist_kernel<0, 0><<<grid, MAC_THREADS, 0, stream>>>(in, n_chunks, chunks_per_group, Io, beta, grp, hist, he_out, maxc_out, valid_out);
        mac_hist_kernel<0, 1><<<grid, MAC_THREADS, 0, stream>>>(in, n_chunks, chunks_per_group, Io, beta, grp, hist, he_out, maxc_out, valid_out);
        mac_hist_kernel<1, 0><<<grid, MAC_THREADS, 0, stream>>>(in, n_chunks, chunks_per_group, Io, beta, grp, hist, he_out, maxc_out, valid_out);
        mac_hist_kernel<1, 1><<<grid, MAC_THREADS, 0, stream>>>(in, n_chunks, chunks_per_group, Io, beta, grp, hist, he_out, maxc_out, valid_out);
        mac_invalid_outputs_kernel<<<(G + 127) / 128, 128, 0, stream>>>(grp, G, he_out, maxc_out, valid_out);
        mac_apply_kernel<<<dim3(per_group_apply, G), MAC_THREADS, 0, stream>>>(in, out, n_chunks, chunks_per_group, Io, grp);
        count_launch(7);
    }
    return cudaGetLastError() == cudaSuccess ? SB_OK : SB_ERR_CUDA;
}

}  // extern "C"
