// Deblur + super-resolution data-fidelity gradient  B^T S^T (S B z - y)  for sm_100a.
//
// Reference: problems/DeblurSR.py:119-120 (fft_blur: length-N circular convolution of the RAVELED
// image with a raveled image-sized kernel, times sqrt(N)), :126-132 (grad_full), :135-147
// (grad_stoch), :95-108 (pylops Bilinear sampling operator).
//
// The length-N = H*W 1-D FFT is done as a four-step transform on the transposed line layout:
//   n = W*n1 + n2,  k = k1 + H*k2
//   X[k1 + H k2] = sum_n2 w_W^(n2 k2) [ w_N^(n2 k1) sum_n1 x[W n1 + n2] w_H^(n1 k1) ]
// i.e. pass 1 (k_lines_r2c, real two-for-one FFT along every line), then per k1 a twiddle, an FFT
// across the lines, the product with the kernel spectrum, the inverse FFT and the conjugate
// twiddle (k_cols_conv, all in shared memory), then pass 3 (k_lines_c2r).  The raveled (helical,
// row-wrapping) boundary of the reference's 1-D convolution is reproduced exactly.
#pragma once
#include "fft_core.cuh"

namespace pnp {

// w_N^(c*k1) = w_H^(m / W) * w_N^(m % W),  m = c*k1 < N;  twn[j] = w_N^j for j < W
template <int L>
__device__ __forceinline__ float2 tw_big(int c, int k1, int H, const float2* __restrict__ twn) {
    const int m = c * k1;
    const int mh = m / L, ml = m - mh * L;                // L == W
    return cmul(g_tw[mh * (PNP_TW_N / H)], twn[ml]);
}

// task 0 = k1 0 (real part of the packed row 0), tasks 1..hp-1 = packed rows, task hp = k1 H/2
// (imaginary slot of packed row 0).  Bf: [hp+1][W] complex kernel spectrum * sqrt(N), row hp = Nyquist.
template <int L, int NC>
__global__ void __launch_bounds__(NC * (L / FftPlan<L>::EPT))
k_cols_conv(float2* __restrict__ S, const float2* __restrict__ Bf, const float2* __restrict__ twn,
            int H, int conj_kernel, long long bf_img_stride) {
    constexpr int T = fft_threads<L>();
    constexpr int EPT = FftPlan<L>::EPT;
    constexpr int PL = fft_plane<L>();
    extern __shared__ float smem[];
    const int hp = H / 2;
    const int g = threadIdx.x / T, t = threadIdx.x % T;
    const int task = blockIdx.x * NC + g;
    const bool active = task <= hp;
    const int img = blockIdx.y;
    const int k1 = task;                                    // hp for the Nyquist task = H/2
    const int row = (task == hp) ? 0 : task;
    const SmemBuf sb{smem + g * 2 * PL, smem + g * 2 * PL + PL};
    float2* Sc = S + ((long long)img * hp + row) * L;
    const float2* bf = Bf + (long long)img * bf_img_stride + (long long)task * L;

    auto ld = [&](int c) -> float2 {
        if (!active) return make_float2(0.f, 0.f);
        float2 v = Sc[c];
        if (task == 0) return make_float2(v.x, 0.f);
        if (task == hp) v = make_float2(v.y, 0.f);
        return cmul(v, tw_big<L>(c, k1, H, twn));
    };
    auto st = [&](int idx, float2 v) { sb.put(idx, v); };
    fft_forward<L, false>(t, sb, ld, st);
    __syncthreads();
    if (active) {
#pragma unroll 4
        for (int m = 0; m < EPT; ++m) {
            const int k2 = t + m * T;
            float2 b = bf[k2];
            if (conj_kernel) b.y = -b.y;
            sb.put(k2, cswap(cmul(sb.get(k2), b)));
        }
    }
    __syncthreads();
    auto ld2 = [&](int idx) -> float2 { return sb.get(idx); };
    auto st2 = [&](int c, float2 v) {
        if (!active) return;
        v = cswap(v);
        if (task == 0) { reinterpret_cast<float*>(Sc + c)[0] = v.x; return; }
        float2 w = tw_big<L>(c, k1, H, twn);
        w.y = -w.y;
        v = cmul(v, w);
        if (task == hp) reinterpret_cast<float*>(Sc + c)[1] = v.x;
        else Sc[c] = v;
    };
    fft_forward<L, true>(t, sb, ld2, st2);
}

// Bilinear sampling + residual + adjoint scatter for the selected measurements:
//   r_m = sum_taps w * x[tap] - (use_y ? y[m] : 0) ;  up[tap] += w * r_m
// x, up are line-layout images ([c][r] -> offset c*H + r).  sel lists measurement ids (null = all M).
__global__ void __launch_bounds__(256)
k_bilinear_residual(const float* __restrict__ x, float* __restrict__ up, const float* __restrict__ y,
                    const int* __restrict__ tl, const float* __restrict__ wts, const int* __restrict__ sel,
                    int count, int H, int identity, int use_y, long long img_stride, long long m_img_stride,
                    long long sel_img_stride, const int* __restrict__ cursor) {
    const int img = blockIdx.y;
    const float* xi = x + (long long)img * img_stride;
    float* ui = up + (long long)img * img_stride;
    const int* si = sel ? sel + (long long)img * sel_img_stride + (long long)(cursor ? *cursor : 0) * count : nullptr;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x) {
        const int m = si ? si[i] : i;
        const float ym = use_y ? y[(long long)img * m_img_stride + m] : 0.f;
        if (identity) {
            // scale_percent == 100: S = I on the raveled image; measurement m = r*W + c
            const int W = (int)(img_stride / H);
            const int r = m / W, c = m - r * W;
            const long long o = (long long)c * H + r;
            atomicAdd(ui + o, xi[o] - ym);
        } else {
            const int t = tl[2 * m], l = tl[2 * m + 1];
            const float wr = wts[2 * m], wc = wts[2 * m + 1];
            const long long o00 = (long long)l * H + t, o01 = o00 + H, o10 = o00 + 1, o11 = o01 + 1;
            const float w00 = (1.f - wr) * (1.f - wc), w01 = (1.f - wr) * wc, w10 = wr * (1.f - wc), w11 = wr * wc;
            const float r_m = w00 * xi[o00] + w01 * xi[o01] + w10 * xi[o10] + w11 * xi[o11] - ym;
            atomicAdd(ui + o00, w00 * r_m);
            atomicAdd(ui + o01, w01 * r_m);
            atomicAdd(ui + o10, w10 * r_m);
            atomicAdd(ui + o11, w11 * r_m);
        }
    }
}

// ---- direct path for blur kernels with a handful of taps (SURVEY section 8(a'): "Minimal" has 4, "Identity" 1) ----
// fft_blur(m, B)[n] = sqrt(N) * sum_p B[p] m[(n - p) mod N] on the RAVELED image (problems/DeblurSR.py:119-120), so a
// tap at raveled position p = pr*W + pc shifts by pr rows and pc columns with a borrow into the row index -- the
// helical boundary the FFT path reproduces.  dir = +1: the blur; dir = -1: its adjoint (kernel roll(flip(B), 1),
// DeblurSR.py:132,147), i.e. m[(n + p) mod N].  Images are in the transposed line layout (element (r, c) at
// c*H + r).  The adjoint call carries the variance-reduction epilogue of the gradient.
#define PNP_MAX_TAPS 16
struct TapList {
    int n;
    int pr[PNP_MAX_TAPS], pc[PNP_MAX_TAPS];
    float w[PNP_MAX_TAPS];                         // B[p] * sqrt(N)
};

__global__ void __launch_bounds__(256)
k_tap_conv(const float* __restrict__ a, const float* __restrict__ b, int H, int W, TapList taps, int dir, float gscale,
           float step, const float* __restrict__ step_ptr, float* __restrict__ g_out, const float* __restrict__ vadd,
           float* __restrict__ v_out, const float* __restrict__ z_in, float* __restrict__ z_out) {
    const long long N = (long long)H * W;
    const float st = step_ptr ? *step_ptr : step;
    for (long long d = (long long)blockIdx.x * blockDim.x + threadIdx.x; d < N; d += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(d / H), r = (int)(d - (long long)c * H);
        float acc = 0.f;
        for (int j = 0; j < taps.n; ++j) {
            int cc = c - dir * taps.pc[j], rr = r - dir * taps.pr[j];
            if (cc < 0) { cc += W; rr -= 1; }          // borrow / carry between the raveled rows
            if (cc >= W) { cc -= W; rr += 1; }
            rr = rr < 0 ? rr + H : (rr >= H ? rr - H : rr);
            const long long s = (long long)cc * H + rr;
            acc = fmaf(taps.w[j], b ? a[s] - b[s] : a[s], acc);
        }
        const float g = acc * gscale;
        if (g_out) g_out[d] = g;
        const float v = vadd ? g + vadd[d] : g;
        if (v_out) v_out[d] = v;
        if (z_out) z_out[d] = z_in[d] - st * v;
    }
}

}  // namespace pnp
