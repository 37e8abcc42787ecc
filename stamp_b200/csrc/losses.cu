// Losses of the regression and survival tasks, each with its gradient, in one single-CTA launch and without a host
// synchronisation (the scores of a batch are a few dozen to a few thousand floats; what matters is that the step
// never waits on the host).
//
// replaces: LitBaseRegressor._compute_loss = nn.functional.l1_loss, src/stamp/modeling/models/__init__.py:420-422, and
// neg_partial_log_likelihood, src/stamp/modeling/models/cox.py:107-268, as LitTileSurvival.training_step calls it
// (models/__init__.py:751-776: ties_method "efron", reduction "mean").  The reference sorts by time, takes the plain Cox
// partial likelihood when all times are distinct (:19-34, log-cumsum-exp over the risk sets) and otherwise loops in
// Python over the unique times with Efron's correction (:37-81) or Breslow's (:84-104).  Here every event sample
// gathers its own risk set, O(n^2) compares out of shared memory, no sort: for a time u with m events (H_u) and risk
// set R_u = {j: t_j >= u},
//     term_u = sum_{h in H_u} s_h - sum_{k<m} log(D_u - k/m T_u),   D_u = sum_{R_u} e^{s_j},  T_u = sum_{H_u} e^{s_h}
//     loss   = - mean over the times with m > 0 of term_u
// which is the Cox form when m = 1 everywhere.  Breslow: every event is its own term with T = 0, mean over the events.
// Scores are shifted by their maximum before exponentiation; sums in fp64.
#include <cuda_runtime.h>
#include <math.h>

#include <cstdint>

#include "common.cuh"
#include "stamp_b200.h"

namespace sb {
namespace {

constexpr int LOSS_THREADS = 1024;

__device__ double block_sum_d(double v, double* red) {
    v = warp_sum_d(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
    for (int w = 0; w < (blockDim.x >> 5); ++w) t += red[w];
    return t;
}

__global__ void __launch_bounds__(LOSS_THREADS)
cox_loss_kernel(const float* __restrict__ log_hz, const float* __restrict__ time, const uint8_t* __restrict__ event, int n,
                int breslow, float grad_scale, float* __restrict__ loss_out, float* __restrict__ dlog_hz) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* t = reinterpret_cast<float*>(smem_raw);
    float* s = t + n;
    float* w = s + n;
    float* g1 = w + n;
    float* g2 = g1 + n;
    uint8_t* e = reinterpret_cast<uint8_t*>(g2 + n);
    uint8_t* rep = e + n;
    __shared__ double red[LOSS_THREADS / 32];
    __shared__ float red_f[LOSS_THREADS / 32];

    float mx = -INFINITY;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        t[i] = time[i];
        s[i] = log_hz[i];
        e[i] = event[i] != 0;
        mx = fmaxf(mx, s[i]);
    }
    mx = warp_max(mx);
    if ((threadIdx.x & 31) == 0) red_f[threadIdx.x >> 5] = mx;
    __syncthreads();
    for (int k = 0; k < (blockDim.x >> 5); ++k) mx = fmaxf(mx, red_f[k]);
    for (int i = threadIdx.x; i < n; i += blockDim.x) w[i] = expf(s[i] - mx);
    __syncthreads();

    // pass 1: one term per event time (Efron; the first event sample of a tie group owns it) or per event (Breslow)
    double term_sum = 0.0, terms = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        float a1 = 0.f, a2 = 0.f;
        bool owner = false;
        if (e[i]) {
            const float ti = t[i];
            double D = 0.0, T = 0.0, S = 0.0;
            int m = 0;
            owner = true;
            for (int j = 0; j < n; ++j) {
                const float tj = t[j];
                if (tj >= ti) D += w[j];
                if (!breslow && tj == ti && e[j]) {
                    ++m;
                    T += w[j];
                    S += s[j];
                    if (j < i) owner = false;
                }
            }
            if (breslow) { m = 1; T = 0.0; S = s[i]; }
            if (owner) {
                double term = S - static_cast<double>(m) * mx, b1 = 0.0, b2 = 0.0;
                for (int k = 0; k < m; ++k) {
                    const double f = static_cast<double>(k) / m, den = D - f * T;
                    term -= log(den);
                    b1 += 1.0 / den;
                    b2 += f / den;
                }
                term_sum += term;
                terms += 1.0;
                a1 = static_cast<float>(b1);
                a2 = static_cast<float>(b2);
            }
        }
        rep[i] = owner;
        g1[i] = a1;
        g2[i] = a2;
    }
    term_sum = block_sum_d(term_sum, red);
    terms = block_sum_d(terms, red);            // block_sum_d starts with a barrier: g1 / g2 / rep are visible after it
    const double inv = terms > 0.0 ? 1.0 / terms : 0.0;
    if (threadIdx.x == 0) *loss_out = static_cast<float>(-term_sum * inv);
    if (dlog_hz == nullptr) return;

    // pass 2: d loss / d s_j = -1/J [ e_j - e^{s_j} sum_{u <= t_j} ( g1_u - [j in H_u] g2_u ) ]
    for (int j = threadIdx.x; j < n; j += blockDim.x) {
        const float tj = t[j];
        const bool ej = e[j];
        double acc = 0.0;
        for (int u = 0; u < n; ++u) {
            if (!rep[u] || t[u] > tj) continue;
            acc += g1[u];
            if (ej && t[u] == tj) acc -= g2[u];
        }
        dlog_hz[j] = grad_scale * static_cast<float>(-inv * ((ej ? 1.0 : 0.0) - static_cast<double>(w[j]) * acc));
    }
}

// loss = mean |pred - target|; dpred = grad_scale * sign(pred - target) / n   (torch: sign(0) = 0)
__global__ void __launch_bounds__(LOSS_THREADS)
l1_loss_kernel(const float* __restrict__ pred, const float* __restrict__ target, long long n, float grad_scale,
               float* __restrict__ loss_out, float* __restrict__ dpred) {
    __shared__ double red[LOSS_THREADS / 32];
    double acc = 0.0;
    const float gs = grad_scale / static_cast<float>(n);
    for (long long i = threadIdx.x; i < n; i += blockDim.x) {
        const float d = pred[i] - target[i];
        acc += fabsf(d);
        if (dpred != nullptr) dpred[i] = d > 0.f ? gs : (d < 0.f ? -gs : 0.f);
    }
    acc = block_sum_d(acc, red);
    if (threadIdx.x == 0) *loss_out = static_cast<float>(acc / static_cast<double>(n));
}

}  // namespace
}  // namespace sb

extern "C" {

int stamp_cox_loss(const float* log_hz, const float* time, const uint8_t* event, int n, int breslow, float grad_scale,
                   float* loss_out, float* dlog_hz_out, void* stream) {
    using namespace sb;
    if (log_hz == nullptr || time == nullptr || event == nullptr || loss_out == nullptr || n <= 0) return SB_ERR_BAD_ARG;
    if (n > STAMP_COX_MAX_SAMPLES) return SB_ERR_UNSUPPORTED;
    const size_t smem = static_cast<size_t>(n) * (5 * sizeof(float) + 2);
    if (smem > 48 * 1024 &&
        cudaFuncSetAttribute(cox_loss_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)) != cudaSuccess)
        return SB_ERR_CUDA;
    cox_loss_kernel<<<1, LOSS_THREADS, smem, static_cast<cudaStream_t>(stream)>>>(log_hz, time, event, n, breslow, grad_scale,
                                                                                  loss_out, dlog_hz_out);
    count_launch();
    return cudaGetLastError() == cudaSuccess ? SB_OK : SB_ERR_CUDA;
}

int stamp_l1_loss(const float* pred, const float* target, long long n, float grad_scale, float* loss_out, float* dpred_out,
                  void* stream) {
    using namespace sb;
    if (pred == nullptr || target == nullptr || loss_out == nullptr || n <= 0) return SB_ERR_BAD_ARG;
    l1_loss_kernel<<<1, LOSS_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(pred, target, n, grad_scale, loss_out, dpred_out);
    count_launch();
    return cudaGetLastError() == cudaSuccess ? SB_OK : SB_ERR_CUDA;
}

}  // extern "C"
