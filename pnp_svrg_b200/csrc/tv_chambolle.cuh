// Total-variation prox by Chambolle's dual projection (A. Chambolle, JMIV 2004) -- the "TV (Chambolle)"
// prox BASELINE.json's north star names.  Nothing in the reference computes it (its TVDenoiser is the
// wavelet BayesShrink of denoisers/TV.py:24,26), so it is an ADDITIVE mode, TVDenoiser(method='chambolle');
// the arithmetic follows scikit-image 0.18.2 restoration._denoise_tv_chambolle_nd (the library the
// reference's denoisers are built on) with a fixed number of iterations instead of the eps stop:
//
//     p = 0                                     (two components, one per axis)
//     repeat:   out = f - div p                 div p (i,j) = p0(i,j) - p0(i-1,j) + p1(i,j) - p1(i,j-1),  p(-1) = 0
//               g   = forward differences of out, 0 on the last row / column of its axis
//               p   = (p - tau g) / (1 + tau/weight * |g|),   tau = 1/4
//     result = out of the LAST pass, i.e. after n_iter - 1 updates of p.
//
// The update has a dependency radius of one pixel, so KB updates are fused per launch by temporal
// blocking: a CTA stages a (32 + 2h) x (64 + 2h) region (h = KB + 1) of f, p0, p1 in shared memory,
// applies KB updates on a region that shrinks by one pixel per update, and writes the p of its 32 x 64
// core (and, in the last launch, the result).  Per pixel and update that is ~4 B of global traffic
// instead of 20 B.  Pixels outside the image hold p = 0 and never contribute a difference.
// The operator is symmetric under transposition (the two axes just swap), so it runs directly on the
// device's transposed line layout.
#pragma once
#include <cuda_runtime.h>

#include "prox.cuh"

namespace pnp {

#define TV_KB 6                     // fused updates per launch
#define TV_TL 32                    // core tile: lines
#define TV_TP 64                    // core tile: samples along a line (contiguous)
#define TV_HALO (TV_KB + 1)
#define TV_RL (TV_TL + 2 * TV_HALO)
#define TV_RP (TV_TP + 2 * TV_HALO)
#define TV_PITCH (TV_RP + 1)
#define TV_SMEM_BYTES (4 * TV_RL * TV_PITCH * sizeof(float))

// f, p_in, p_out: [batch][nl][np] (p: two planes per image, [batch][2][nl][np]);  updates <= TV_KB.
// emit != 0: also write out = f - div p_new and accumulate the squared error against xrec.
__global__ void __launch_bounds__(256)
k_tv_chambolle(const float* __restrict__ f, const float* __restrict__ p_in, float* __restrict__ p_out,
               float* __restrict__ out, int nl, int np, int updates, int first, int emit, float weight,
               const double* __restrict__ sig_log, float sigma_modifier, float fallback_weight,
               const float* __restrict__ xrec, double* __restrict__ mse_log, const int* __restrict__ slot, int batch) {
    extern __shared__ float tv_smem[];                 // four planes of TV_RL x TV_PITCH floats (58 KiB)
    float (*s_f)[TV_PITCH] = reinterpret_cast<float (*)[TV_PITCH]>(tv_smem);
    float (*s_p0)[TV_PITCH] = s_f + TV_RL;
    float (*s_p1)[TV_PITCH] = s_p0 + TV_RL;
    float (*s_o)[TV_PITCH] = s_p1 + TV_RL;
    __shared__ float s_err[8];
    const int img = blockIdx.z;
    const long long n = (long long)nl * np;
    const float* fi = f + img * n;
    const float* pi = p_in + img * 2 * n;
    float* po = p_out + img * 2 * n;
    const int l0 = blockIdx.y * TV_TL - TV_HALO, c0 = blockIdx.x * TV_TP - TV_HALO;
    const float tau = 0.25f;
    // weight: fixed, or sigma_est * sigma_modifier from the loops' estimate (mean over lines), or the fallback;
    // weight <= 0 is the identity (no dual updates)
    float w = weight;
    if (!(w > 0.f)) {
        const double se = sig_log ? *slot_ptr(const_cast<double*>(sig_log), slot, batch, img) / (double)nl : 0.0;
        w = se > 0.0 ? (float)(se * (double)sigma_modifier) : fallback_weight;
    }
    if (!(w > 0.f)) updates = 0;
    const float tau_over_w = w > 0.f ? tau / w : 0.f;

    for (int i = threadIdx.x; i < TV_RL * TV_RP; i += 256) {
        const int rl = i / TV_RP, rp = i - rl * TV_RP;
        const int gl = l0 + rl, gp = c0 + rp;
        const bool in = gl >= 0 && gl < nl && gp >= 0 && gp < np;
        const long long o = (long long)gl * np + gp;
        s_f[rl][rp] = in ? fi[o] : 0.f;
        s_p0[rl][rp] = (in && !first) ? pi[o] : 0.f;
        s_p1[rl][rp] = (in && !first) ? pi[n + o] : 0.f;
    }
    __syncthreads();

    // after update u the valid region is [u+1, R-1-(u+1)) ... computed generously: every pass covers the whole
    // region minus a one-pixel rim; cells whose inputs are stale are never read by the core (halo = KB + 1).
    for (int u = 0; u < updates; ++u) {
        for (int i = threadIdx.x; i < TV_RL * TV_RP; i += 256) {
            const int rl = i / TV_RP, rp = i - rl * TV_RP;
            float d = -(s_p0[rl][rp] + s_p1[rl][rp]);
            if (rl > 0) d += s_p0[rl - 1][rp];
            if (rp > 0) d += s_p1[rl][rp - 1];
            s_o[rl][rp] = s_f[rl][rp] + d;
        }
        __syncthreads();
        for (int i = threadIdx.x; i < TV_RL * TV_RP; i += 256) {
            const int rl = i / TV_RP, rp = i - rl * TV_RP;
            const int gl = l0 + rl, gp = c0 + rp;
            if (gl < 0 || gl >= nl || gp < 0 || gp >= np) continue;          // p stays 0 outside the image
            const float o = s_o[rl][rp];
            const float g0 = (rl + 1 < TV_RL && gl + 1 < nl) ? s_o[rl + 1][rp] - o : 0.f;
            const float g1 = (rp + 1 < TV_RP && gp + 1 < np) ? s_o[rl][rp + 1] - o : 0.f;
            const float nrm = 1.f + tau_over_w * sqrtf(g0 * g0 + g1 * g1);
            s_p0[rl][rp] = (s_p0[rl][rp] - tau * g0) / nrm;
            s_p1[rl][rp] = (s_p1[rl][rp] - tau * g1) / nrm;
        }
        __syncthreads();
    }

    float err = 0.f;
    for (int i = threadIdx.x; i < TV_TL * TV_TP; i += 256) {
        const int tl = i / TV_TP, tp = i - tl * TV_TP;
        const int rl = tl + TV_HALO, rp = tp + TV_HALO;
        const int gl = l0 + rl, gp = c0 + rp;
        if (gl >= nl || gp >= np) continue;
        const long long o = (long long)gl * np + gp;
        po[o] = s_p0[rl][rp];
        po[n + o] = s_p1[rl][rp];
        if (emit) {
            const float d = -(s_p0[rl][rp] + s_p1[rl][rp]) + s_p0[rl - 1][rp] + s_p1[rl][rp - 1];
            const float v = s_f[rl][rp] + d;
            out[img * n + o] = v;
            if (xrec) { const float e = v - xrec[img * n + o]; err = fmaf(e, e, err); }
        }
    }
    if (emit && xrec && mse_log) {
        err = warp_sum_f(err);
        if ((threadIdx.x & 31) == 0) s_err[threadIdx.x >> 5] = err;
        __syncthreads();
        if (threadIdx.x == 0) {
            float t = 0.f;
            for (int k = 0; k < 8; ++k) t += s_err[k];
            atomicAdd(slot_ptr(mse_log, slot, batch, img), (double)t);
        }
    }
}

}  // namespace pnp
