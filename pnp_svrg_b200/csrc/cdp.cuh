// Coded-diffraction-pattern phase retrieval, intensity loss -- the "coded-diffraction |Ax|^2 - y Wirtinger
// gradient" BASELINE.json's north star names (config 3).  No counterpart in the reference, whose
// problems/PR.py:26,75-87 is a dense real Gaussian A with the amplitude loss (built in pr.cuh); this is the
// ADDITIVE mode PhaseRetrieval(model='cdp') of SURVEY section 8(a').
//
//   A_l x = fft2(d_l o x) / sqrt(N),  d_l in {1, i, -1, -i}^(HxW),  l < L;   y = |A x|^2 (+ noise),  M = L N
//   f(x)  = 1/(4M) sum_k (|a_k^H x|^2 - y_k)^2
//   grad  = Re( sum_l conj(d_l) o ifft2_unitary( m_l o (|A_l x|^2 - y_l) o A_l x ) )      (m = minibatch 0/1)
//         = Re( sum_l conj(d_l) o IFFT2( m_l o (|F_l|^2 / N - y_l) o F_l ) ) / N,  F_l = FFT2(d_l o x) unnormalised
//
// Complex 2-D transforms in three passes over S[l][line][sample] (complex64, the device's transposed line
// layout: line c = image column c; fft2 commutes with the transposition): lines forward, columns forward +
// pointwise residual + columns inverse, lines inverse + decode + sum over the masks.  All on the shared
// Stockham core (fft_core.cuh).
#pragma once
#include "fft_core.cuh"

namespace pnp {

__device__ __forceinline__ float2 cdp_code_mul(float v, int c) {             // v * i^c
    return c == 0 ? make_float2(v, 0.f) : c == 1 ? make_float2(0.f, v) : c == 2 ? make_float2(-v, 0.f) : make_float2(0.f, -v);
}
__device__ __forceinline__ float cdp_decode_re(float2 g, int c) {            // Re(conj(i^c) * g)
    return c == 0 ? g.x : c == 1 ? g.y : c == 2 ? -g.x : -g.y;
}
template <int L> __host__ __device__ constexpr int cdp_lines_per_cta() { return fft_threads<L>() >= 128 ? 1 : 128 / fft_threads<L>(); }
template <int L> __host__ __device__ constexpr int cdp_cols_per_cta() {
    return fft_threads<L>() >= 256 ? 1 : (fft_threads<L>() >= 128 ? 2 : (fft_threads<L>() >= 64 ? 4 : 8));
}

// minibatch: measurement ids m = l*N + ky*W + kx (the reference's flat (H, W) order per mask) -> bytes in the
// device layout mask[l][kx][ky]
__global__ void k_cdp_sel(unsigned char* __restrict__ mask, const int* __restrict__ idx, int count, int H, int W,
                          const int* __restrict__ cursor) {
    const int* src = idx + (long long)(cursor ? *cursor : 0) * count;
    const int N = H * W;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x) {
        const int m = src[i];
        const int l = m / N, rem = m - l * N;
        const int ky = rem / W, kx = rem - ky * W;
        mask[(long long)l * N + (long long)kx * H + ky] = 1;
    }
}

// pass 1: S[l][line][:] = FFT_L( i^code o u[line][:] )
template <int L>
__global__ void __launch_bounds__(cdp_lines_per_cta<L>() * fft_threads<L>())
k_cdp_lines_fwd(const float* __restrict__ u, const signed char* __restrict__ codes, float2* __restrict__ S, int nlines,
                const float* __restrict__ u2, float2* __restrict__ S2) {          // blockIdx.z = 1: the second point (u2 -> S2)
    constexpr int T = fft_threads<L>(), EPT = FftPlan<L>::EPT, PL = fft_plane<L>();
    using IX = FftIdx<L>;
    extern __shared__ __align__(16) float smem[];
    const int g = threadIdx.x / T, t = threadIdx.x % T;
    const int line = blockIdx.x * cdp_lines_per_cta<L>() + g, l = blockIdx.y;
    if (blockIdx.z) { u = u2; S = S2; }
    const bool active = line < nlines;
    const SmemBuf sb{smem + g * 2 * PL, smem + g * 2 * PL + PL};
    const long long lo = (long long)(active ? line : 0) * L, so = ((long long)l * nlines + (active ? line : 0)) * L;
    float2 x[EPT];
#pragma unroll
    for (int i = 0; i < EPT; ++i) {
        const int idx = IX::in(t, i);
        x[i] = active ? cdp_code_mul(u[lo + idx], codes[so + idx]) : make_float2(0.f, 0.f);
    }
    fft_regs<L>(t, sb, x);
    if (active) {
#pragma unroll
        for (int i = 0; i < EPT; ++i) S[so + IX::out(t, i)] = x[i];
    }
}

// pass 2, per mask l and block of NC adjacent samples ky: FFT along the lines, G = m (|F|^2/N - y) F, inverse FFT.
// Element (line c, sample ky) lives at S[(l*LW + c)*H + ky]: the NC columns of a CTA are NC*8 contiguous bytes.
template <int LW>
__global__ void __launch_bounds__(cdp_cols_per_cta<LW>() * fft_threads<LW>())
k_cdp_cols(float2* __restrict__ S, const float* __restrict__ y, unsigned char* __restrict__ mask, int H, float inv_n,
           int clear_mask, float2* __restrict__ S2) {      // S2: the second point's planes, done by the same CTA after S
    constexpr int T = fft_threads<LW>(), EPT = FftPlan<LW>::EPT, PL = fft_plane<LW>(), NC = cdp_cols_per_cta<LW>();
    using IX = FftIdx<LW>;
    extern __shared__ __align__(16) float smem[];
    const int g = threadIdx.x % NC, t = threadIdx.x / NC;
    const int ky = blockIdx.x * NC + g, l = blockIdx.y;
    const SmemBuf sb{smem + g * 2 * PL, smem + g * 2 * PL + PL};
    const long long base = (long long)l * LW * H + ky;
    const int npts = S2 ? 2 : 1;
    for (int pt = 0; pt < npts; ++pt) {
        float2* Sp = pt ? S2 : S;
        float2 x[EPT];
#pragma unroll
        for (int i = 0; i < EPT; ++i) x[i] = Sp[base + (long long)IX::in(t, i) * H];
        fft_regs<LW>(t, sb, x);
        float2 v[EPT];
#pragma unroll
        for (int m = 0; m < EPT; ++m) {
            const long long e = base + (long long)(t + T * m) * H;           // this thread owns kx = t + T*m before and after
            const float2 F = x[IX::out_slot(m)];
            float q = fmaf(F.x * F.x + F.y * F.y, inv_n, -y[e]);
            if (mask) {
                if (!mask[e]) q = 0.f;
                else if (clear_mask && pt == npts - 1) mask[e] = 0;           // single-use minibatch selection
            }
            v[IX::in_slot(m)] = make_float2(q * F.y, q * F.x);                // re/im swapped: inverse by the forward core
        }
        if (FftPlan<LW>::NS > 1) __syncthreads();
        fft_regs<LW>(t, sb, v);
#pragma unroll
        for (int i = 0; i < EPT; ++i) Sp[base + (long long)IX::out(t, i) * H] = cswap(v[i]);
        if (pt + 1 < npts) __syncthreads();       // exchange buffer free for the second point
    }
}

// pass 3: acc[line][:] (+)= sign * sum_l Re( conj(i^code) o IFFT_L(S[l][line][:]) )
template <int L>
__global__ void __launch_bounds__(cdp_lines_per_cta<L>() * fft_threads<L>())
k_cdp_lines_inv(const float2* __restrict__ S, const signed char* __restrict__ codes, float* __restrict__ acc, int nlines,
                int nmasks, float sign, int accumulate, const float2* __restrict__ S2) {   // S2: acc = sum(S) - sum(S2)
    constexpr int T = fft_threads<L>(), EPT = FftPlan<L>::EPT, PL = fft_plane<L>();
    using IX = FftIdx<L>;
    extern __shared__ __align__(16) float smem[];
    const int g = threadIdx.x / T, t = threadIdx.x % T;
    const int line = blockIdx.x * cdp_lines_per_cta<L>() + g;
    const bool active = line < nlines;
    const SmemBuf sb{smem + g * 2 * PL, smem + g * 2 * PL + PL};
    float sum[EPT];
#pragma unroll
    for (int i = 0; i < EPT; ++i) sum[i] = 0.f;
    for (int l = 0; l < nmasks; ++l) {
        const long long so = ((long long)l * nlines + (active ? line : 0)) * L;
        float2 x[EPT];
#pragma unroll
        for (int i = 0; i < EPT; ++i) x[i] = active ? cswap(S[so + IX::in(t, i)]) : make_float2(0.f, 0.f);
        fft_regs<L>(t, sb, x);
        if (active) {
#pragma unroll
            for (int i = 0; i < EPT; ++i) sum[i] += cdp_decode_re(cswap(x[i]), codes[so + IX::out(t, i)]);
        }
        __syncthreads();                          // exchange buffer free for the next mask
    }
    if (S2) {
        // the second point in the same launch: its masks are summed separately and subtracted once, which is what the
        // two-launch sequence (acc = sum, then acc = fma(-1, sum2, acc)) computes
        float sum2[EPT];
#pragma unroll
        for (int i = 0; i < EPT; ++i) sum2[i] = 0.f;
        for (int l = 0; l < nmasks; ++l) {
            const long long so = ((long long)l * nlines + (active ? line : 0)) * L;
            float2 x[EPT];
#pragma unroll
            for (int i = 0; i < EPT; ++i) x[i] = active ? cswap(S2[so + IX::in(t, i)]) : make_float2(0.f, 0.f);
            fft_regs<L>(t, sb, x);
            if (active) {
#pragma unroll
                for (int i = 0; i < EPT; ++i) sum2[i] += cdp_decode_re(cswap(x[i]), codes[so + IX::out(t, i)]);
            }
            __syncthreads();
        }
#pragma unroll
        for (int i = 0; i < EPT; ++i) sum[i] = sum[i] - sum2[i];
    }
    if (active) {
        const long long lo = (long long)line * L;
#pragma unroll
        for (int i = 0; i < EPT; ++i) {
            const long long e = lo + IX::out(t, i);
            acc[e] = accumulate ? fmaf(sign, sum[i], acc[e]) : sign * sum[i];
        }
    }
}

// g = acc * gscale ; v = g + vadd ; z_out = z_in - step * v
__global__ void __launch_bounds__(256)
k_cdp_epilogue(const float* __restrict__ acc, long long n, float gscale, float step, const float* __restrict__ step_ptr,
               float* __restrict__ g_out, const float* __restrict__ vadd, float* __restrict__ v_out,
               const float* __restrict__ z_in, float* __restrict__ z_out) {
    const float st = step_ptr ? *step_ptr : step;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float g = acc[i] * gscale;
        if (g_out) g_out[i] = g;
        const float v = vadd ? g + vadd[i] : g;
        if (v_out) v_out[i] = v;
        if (z_out) z_out[i] = z_in[i] - st * v;
    }
}

}  // namespace pnp
