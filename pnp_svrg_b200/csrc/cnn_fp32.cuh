// 3x3 convolution stacks (DnCNN-17 / RealSN-DnCNN, SimpleCNN-4, MMO DnCNN_nobn-20) in fp32 on
// CUDA cores -- the EXACT-parity path (rel. error ~1e-6 per layer vs the torch fp32 forward).
// The tensor-core path (bf16, tcgen05) lives in cnn_tc.cuh.
//
// Reference: denoisers/RealSN_DnCNN.py:16-40 (min/max normalise, scale/shift, net, residual
// subtract, undo), denoisers/DeepDenoisers/model/models.py:5-22 and realSN_models.py:5-18
// (Conv3x3 no bias [+ BN eval] + ReLU), denoisers/MMODenoise.py:18-40,73-103 (clamp, conv + bias +
// LeakyReLU(0.01), out_conv(x) + x_in, clamp).
//
// Images are in the transposed line layout, viewed as PH = W rows (lines) of PW = H pixels; the
// packed weights have their two spatial axes swapped accordingly.  Activations are NHWC fp32
// [pixel][64].  BN (eval) is folded to a per-channel scale/shift applied after the convolution.
#pragma once
#include <cuda_runtime.h>

namespace pnp {

#define CNN_C 64

struct CnnAct {
    const float* scale;   // [64] or null (=1)
    const float* shift;   // [64] or null (=0): bias or folded BN shift
    float slope;          // 0 = ReLU, 0.01 = LeakyReLU, 1 = identity
};

__device__ __forceinline__ float act(float v, float slope) { return v > 0.f ? v : v * slope; }

// stats[0] = min, stats[1] = max of an image (RealSN_DnCNN.py:20-21); encoded as ordered ints
__device__ __forceinline__ int f2ord(float f) { int i = __float_as_int(f); return i >= 0 ? i : i ^ 0x7fffffff; }
__device__ __forceinline__ float ord2f(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

__global__ void k_minmax_init(int* __restrict__ stats) {
    if (threadIdx.x == 0) { stats[0] = 0x7fffffff; stats[1] = (int)0x80000000; }
}

__global__ void __launch_bounds__(256)
k_minmax(const float* __restrict__ x, long long n, int* __restrict__ stats) {
    float mn = 3.4e38f, mx = -3.4e38f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float v = x[i];
        mn = fminf(mn, v);
        mx = fmaxf(mx, v);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    if ((threadIdx.x & 31) == 0) { atomicMin(stats, f2ord(mn)); atomicMax(stats + 1, f2ord(mx)); }
}

// how the network input is formed from the image, and how its output becomes the denoised image
struct CnnIo {
    int mode;             // 0: DnCNN wrapper (min/max normalise, residual net); 1: MMO (clamp, net output + input, clamp)
    const int* stats;     // mode 0: ordered-int min/max
    float range, shift;   // mode 0: xt = (x - min) / (max - min) * range + shift
};

__device__ __forceinline__ float cnn_input(const CnnIo& io, float x) {
    if (io.mode == 0) {
        const float mn = ord2f(io.stats[0]), mx = ord2f(io.stats[1]);
        return (x - mn) / (mx - mn) * io.range + io.shift;
    }
    return fminf(fmaxf(x, 0.f), 1.f);
}

// first layer: 1 -> 64 channels.  w: [9][64]
__global__ void __launch_bounds__(256)
k_conv_first(const float* __restrict__ img, float* __restrict__ out, const float* __restrict__ w, CnnAct a, CnnIo io,
             int PH, int PW) {
    __shared__ float sw[9 * CNN_C];
    for (int i = threadIdx.x; i < 9 * CNN_C; i += blockDim.x) sw[i] = w[i];
    __syncthreads();
    const long long total = (long long)PH * PW * (CNN_C / 4);
    for (long long id = (long long)blockIdx.x * blockDim.x + threadIdx.x; id < total; id += (long long)gridDim.x * blockDim.x) {
        const int cg = (int)(id % (CNN_C / 4));
        const long long pix = id / (CNN_C / 4);
        const int l = (int)(pix / PW), p = (int)(pix % PW);
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int dl = -1; dl <= 1; ++dl)
#pragma unroll
            for (int dp = -1; dp <= 1; ++dp) {
                const int ll = l + dl, pp = p + dp;
                if (ll < 0 || ll >= PH || pp < 0 || pp >= PW) continue;
                const float v = cnn_input(io, img[(long long)ll * PW + pp]);
                const float* ww = sw + ((dl + 1) * 3 + (dp + 1)) * CNN_C + cg * 4;
#pragma unroll
                for (int k = 0; k < 4; ++k) acc[k] = fmaf(v, ww[k], acc[k]);
            }
        float4 o;
        float* op = &o.x;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int c = cg * 4 + k;
            float v = acc[k];
            if (a.scale) v *= a.scale[c];
            if (a.shift) v += a.shift[c];
            op[k] = act(v, a.slope);
        }
        reinterpret_cast<float4*>(out)[id] = o;
    }
}

// middle layers: 64 -> 64.  w: [9][64 ci][64 co].  Tile: 16 lines x 32 pixels, 2 pixels per thread.
#define CM_TL 16
#define CM_TP 32
#define CM_CK 8
__global__ void __launch_bounds__(256, 1)
k_conv_mid(const float* __restrict__ in, float* __restrict__ out, const float* __restrict__ w, CnnAct a, int PH, int PW) {
    __shared__ float s_in[CM_CK][CM_TL + 2][CM_TP + 2];
    __shared__ __align__(16) float s_w[9][CM_CK][CNN_C];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int l0 = blockIdx.y * CM_TL, p0 = blockIdx.x * CM_TP;
    // accumulators as channel pairs: fma.rn.f32x2 issues two FMAs per slot (each half rounds like fmaf)
    float2 acc0[CNN_C / 2], acc1[CNN_C / 2];
#pragma unroll
    for (int c = 0; c < CNN_C / 2; ++c) { acc0[c] = make_float2(0.f, 0.f); acc1[c] = make_float2(0.f, 0.f); }
    for (int c0 = 0; c0 < CNN_C; c0 += CM_CK) {
        __syncthreads();
        // input tile with halo: 8 channels = 32 contiguous bytes per pixel
        for (int i = threadIdx.x; i < (CM_TL + 2) * (CM_TP + 2) * 2; i += 256) {
            const int half = i & 1, pix = i >> 1;
            const int ll = pix / (CM_TP + 2), pp = pix % (CM_TP + 2);
            const int gl = l0 + ll - 1, gp = p0 + pp - 1;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (gl >= 0 && gl < PH && gp >= 0 && gp < PW)
                v = *reinterpret_cast<const float4*>(in + ((long long)gl * PW + gp) * CNN_C + c0 + half * 4);
            s_in[half * 4 + 0][ll][pp] = v.x;
            s_in[half * 4 + 1][ll][pp] = v.y;
            s_in[half * 4 + 2][ll][pp] = v.z;
            s_in[half * 4 + 3][ll][pp] = v.w;
        }
        for (int i = threadIdx.x; i < 9 * CM_CK * CNN_C / 4; i += 256) {
            const int co4 = i % (CNN_C / 4), r = i / (CNN_C / 4);
            const int ci = r % CM_CK, tap = r / CM_CK;
            reinterpret_cast<float4*>(&s_w[tap][ci][0])[co4] =
                reinterpret_cast<const float4*>(w + ((long long)tap * CNN_C + c0 + ci) * CNN_C)[co4];
        }
        __syncthreads();
#pragma unroll 1
        for (int tap = 0; tap < 9; ++tap) {
            const int dl = tap / 3, dp = tap % 3;
#pragma unroll
            for (int ci = 0; ci < CM_CK; ++ci) {
                const float x0 = s_in[ci][ty + dl][2 * tx + dp];
                const float x1 = s_in[ci][ty + dl][2 * tx + 1 + dp];
                const float2 xx0 = make_float2(x0, x0), xx1 = make_float2(x1, x1);
                const float4* wv = reinterpret_cast<const float4*>(&s_w[tap][ci][0]);
#pragma unroll
                for (int q = 0; q < CNN_C / 4; ++q) {
                    const float4 ww = wv[q];
                    const float2 wlo = make_float2(ww.x, ww.y), whi = make_float2(ww.z, ww.w);
                    acc0[2 * q + 0] = __ffma2_rn(xx0, wlo, acc0[2 * q + 0]); acc1[2 * q + 0] = __ffma2_rn(xx1, wlo, acc1[2 * q + 0]);
                    acc0[2 * q + 1] = __ffma2_rn(xx0, whi, acc0[2 * q + 1]); acc1[2 * q + 1] = __ffma2_rn(xx1, whi, acc1[2 * q + 1]);
                }
            }
        }
    }
    const int gl = l0 + ty;
    if (gl >= PH) return;
#pragma unroll
    for (int px = 0; px < 2; ++px) {
        const int gp = p0 + 2 * tx + px;
        if (gp >= PW) continue;
        float* o = out + ((long long)gl * PW + gp) * CNN_C;
#pragma unroll
        for (int q = 0; q < CNN_C / 4; ++q) {
            float v[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int c = 4 * q + k;
                const float2 pr = px == 0 ? acc0[c >> 1] : acc1[c >> 1];
                float t = (c & 1) ? pr.y : pr.x;
                if (a.scale) t *= a.scale[c];
                if (a.shift) t += a.shift[c];
                v[k] = act(t, a.slope);
            }
            reinterpret_cast<float4*>(o)[q] = make_float4(v[0], v[1], v[2], v[3]);
        }
    }
}

// last layer: 64 -> 1 and the wrapper's output map.  w: [9][64]; bias scalar.
//   mode 0: r = conv ; x = xt - r ; out = (x - shift) / range * (max - min) + min   (RealSN_DnCNN.py:36-40)
//   mode 1: out = clamp(conv + bias + clamp(x_in), 0, 1)                            (MMODenoise.py:30-32,101,128)
__global__ void __launch_bounds__(256)
k_conv_last(const float* __restrict__ in, const float* __restrict__ img, float* __restrict__ out,
            const float* __restrict__ w, float bias, CnnIo io, int PH, int PW, const float* __restrict__ xrec,
            double* __restrict__ mse_log, const int* __restrict__ slot) {
    __shared__ float sw[9 * CNN_C];
    __shared__ float s_err[8];
    for (int i = threadIdx.x; i < 9 * CNN_C; i += blockDim.x) sw[i] = w[i];
    __syncthreads();
    // one warp per pixel: lanes split the 64 channels (2 each), taps looped
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const long long npix = (long long)PH * PW;
    float err = 0.f;
    for (long long pix = (long long)blockIdx.x * 8 + wib; pix < npix; pix += (long long)gridDim.x * 8) {
        const int l = (int)(pix / PW), p = (int)(pix % PW);
        float acc = 0.f;
#pragma unroll
        for (int dl = -1; dl <= 1; ++dl)
#pragma unroll
            for (int dp = -1; dp <= 1; ++dp) {
                const int ll = l + dl, pp = p + dp;
                if (ll < 0 || ll >= PH || pp < 0 || pp >= PW) continue;
                const float2 v = reinterpret_cast<const float2*>(in + ((long long)ll * PW + pp) * CNN_C)[lane];
                const float* ww = sw + ((dl + 1) * 3 + (dp + 1)) * CNN_C + 2 * lane;
                acc = fmaf(v.x, ww[0], fmaf(v.y, ww[1], acc));
            }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) {
            const float x = img[pix];
            float res;
            if (io.mode == 0) {
                const float mn = ord2f(io.stats[0]), mx = ord2f(io.stats[1]);
                const float xt = (x - mn) / (mx - mn) * io.range + io.shift;
                res = ((xt - acc) - io.shift) / io.range * (mx - mn) + mn;
            } else {
                res = fminf(fmaxf(acc + bias + fminf(fmaxf(x, 0.f), 1.f), 0.f), 1.f);
            }
            out[pix] = res;
            if (xrec) { const float d = res - xrec[pix]; err = fmaf(d, d, err); }
        }
    }
    if (xrec && mse_log) {
        if (lane == 0) s_err[wib] = err;
        __syncthreads();
        if (threadIdx.x == 0) {
            float t = 0.f;
            for (int k = 0; k < 8; ++k) t += s_err[k];
            atomicAdd(mse_log + (slot ? *slot : 0), (double)t);
        }
    }
}

}  // namespace pnp
