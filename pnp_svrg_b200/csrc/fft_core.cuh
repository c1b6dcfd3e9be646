// Shared-memory Stockham FFT building blocks for sm_100a.
//
// One power-of-two transform of length L is done by T = L / EPT cooperating threads that keep
// EPT complex values each in registers.  Every stage is an in-register radix-R DFT (R in
// {4, 8, 16}); between stages the values are exchanged through a padded, split re/im
// shared-memory buffer (autosort / Stockham indexing, so the result is in natural order and
// no bit reversal pass exists).  The first stage reads through a caller supplied functor
// (global memory, or shared memory staged by the caller) and the last stage hands its results
// to a caller functor, so the surrounding kernel can fuse its own prologue / epilogue around
// the transform without an extra shared-memory round trip.
//
// The inverse transform is the forward one applied to re/im-swapped data
// (ifft(x) = swap(fft(swap(x))), unnormalised).
//
// Replaces np.fft.fft2 / ifft2 / fft / ifft at reference problems/CSMRI.py:77,81,85,89 and
// problems/DeblurSR.py:120.
#pragma once
#include <cuda_runtime.h>

#define PNP_TW_N 4096               // twiddle table: g_tw[m] = exp(-2*pi*i*m / PNP_TW_N)

__device__ float2 g_tw[PNP_TW_N];

namespace pnp {

__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x));
}
__device__ __forceinline__ float2 cswap(float2 a) { return make_float2(a.y, a.x); }

// cos/sin(2*pi*k/16), k = 0..7
__device__ constexpr float kC16[8] = {1.0f, 0.92387953251128674f, 0.70710678118654752f, 0.38268343236508977f,
                                      0.0f, -0.38268343236508977f, -0.70710678118654752f, -0.92387953251128674f};
__device__ constexpr float kS16[8] = {0.0f, 0.38268343236508977f, 0.70710678118654752f, 0.92387953251128674f,
                                      1.0f, 0.92387953251128674f, 0.70710678118654752f, 0.38268343236508977f};

// forward DFT of R values held in registers: v[k] = sum_n v[n] exp(-2*pi*i*n*k/R)
template <int R>
struct Dft {
    __device__ __forceinline__ static void run(float2 (&v)[R]) {
        float2 e[R / 2], o[R / 2];
#pragma unroll
        for (int k = 0; k < R / 2; ++k) { e[k] = v[2 * k]; o[k] = v[2 * k + 1]; }
        Dft<R / 2>::run(e);
        Dft<R / 2>::run(o);
#pragma unroll
        for (int k = 0; k < R / 2; ++k) {
            constexpr int S = 16 / R;
            float2 t;
            if (k == 0) {
                t = o[k];
            } else if (4 * k == R) {
                t = make_float2(o[k].y, -o[k].x);                        // * (-i)
            } else {
                t = cmul(o[k], make_float2(kC16[k * S], -kS16[k * S]));  // * exp(-2*pi*i*k/R)
            }
            v[k] = cadd(e[k], t);
            v[k + R / 2] = csub(e[k], t);
        }
    }
};
template <>
struct Dft<1> {
    __device__ __forceinline__ static void run(float2 (&)[1]) {}
};

// ------------------------------------------------------------------ plans
template <int L> struct FftPlan;
template <> struct FftPlan<8>    { static constexpr int NS = 1, R0 = 8,  R1 = 1,  R2 = 1,  EPT = 8;  };
template <> struct FftPlan<16>   { static constexpr int NS = 1, R0 = 16, R1 = 1,  R2 = 1,  EPT = 16; };
template <> struct FftPlan<32>   { static constexpr int NS = 2, R0 = 8,  R1 = 4,  R2 = 1,  EPT = 8;  };
template <> struct FftPlan<64>   { static constexpr int NS = 2, R0 = 8,  R1 = 8,  R2 = 1,  EPT = 8;  };
template <> struct FftPlan<128>  { static constexpr int NS = 2, R0 = 16, R1 = 8,  R2 = 1,  EPT = 16; };
template <> struct FftPlan<256>  { static constexpr int NS = 2, R0 = 16, R1 = 16, R2 = 1,  EPT = 16; };
template <> struct FftPlan<512>  { static constexpr int NS = 3, R0 = 8,  R1 = 8,  R2 = 8,  EPT = 8;  };
template <> struct FftPlan<1024> { static constexpr int NS = 3, R0 = 16, R1 = 8,  R2 = 8,  EPT = 16; };
template <> struct FftPlan<2048> { static constexpr int NS = 3, R0 = 16, R1 = 16, R2 = 8,  EPT = 16; };
template <> struct FftPlan<4096> { static constexpr int NS = 3, R0 = 16, R1 = 16, R2 = 16, EPT = 16; };

template <int L> __host__ __device__ constexpr int fft_threads() { return L / FftPlan<L>::EPT; }

// padded index into a split re/im plane (one extra word every 32 to spread power-of-two strides)
__host__ __device__ __forceinline__ constexpr int fpad(int i) { return i + (i >> 5); }
// floats in ONE plane of one transform's exchange buffer
template <int L> __host__ __device__ constexpr int fft_plane() { return L + (L >> 5) + 1; }

struct SmemBuf {
    float* re;
    float* im;
    __device__ __forceinline__ float2 get(int i) const { int p = fpad(i); return make_float2(re[p], im[p]); }
    __device__ __forceinline__ void put(int i, float2 v) const { int p = fpad(i); re[p] = v.x; im[p] = v.y; }
};

// One Stockham stage.  FIRST: inputs come from ld(idx); LAST: outputs go to st(idx, v).
// LD_SMEM says the FIRST-stage loader itself reads the exchange buffer (needs a barrier before
// the stage writes it).
template <int L, int R, int NSP, bool FIRST, bool LAST, bool LD_SMEM, class Load, class Store>
__device__ __forceinline__ void fft_stage(int t, const SmemBuf& sb, Load& ld, Store& st) {
    constexpr int EPT = FftPlan<L>::EPT;
    constexpr int T = L / EPT;
    constexpr int NB = EPT / R;
    constexpr int LR = L / R;
    float2 v[NB][R];
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        const int j = t + b * T;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int idx = j + r * LR;
            if (FIRST) v[b][r] = ld(idx); else v[b][r] = sb.get(idx);
        }
    }
    if (!FIRST || LD_SMEM) __syncthreads();      // every read done before the in-place writes
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        const int j = t + b * T;
        if (NSP > 1) {
            const int k = j % NSP;
#pragma unroll
            for (int r = 1; r < R; ++r) {
                const float2 w = g_tw[(r * k) * (PNP_TW_N / (NSP * R))];
                v[b][r] = cmul(v[b][r], w);
            }
        }
        Dft<R>::run(v[b]);
        const int base = (j / NSP) * (NSP * R) + (j % NSP);
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int idx = base + r * NSP;
            if (LAST) st(idx, v[b][r]); else sb.put(idx, v[b][r]);
        }
    }
    if (!LAST) __syncthreads();                  // exchange visible to the next stage
}

// Forward FFT of length L by threads t = 0 .. fft_threads<L>()-1 of a group.  Every thread of
// the CTA must call it (it contains __syncthreads()).  The exchange buffer must be free on entry
// unless LD_SMEM, in which case ld() may read it.
template <int L, bool LD_SMEM, class Load, class Store>
__device__ __forceinline__ void fft_forward(int t, const SmemBuf& sb, Load& ld, Store& st) {
    using P = FftPlan<L>;
    if constexpr (P::NS == 1) {
        fft_stage<L, P::R0, 1, true, true, LD_SMEM>(t, sb, ld, st);
    } else if constexpr (P::NS == 2) {
        fft_stage<L, P::R0, 1, true, false, LD_SMEM>(t, sb, ld, st);
        fft_stage<L, P::R1, P::R0, false, true, false>(t, sb, ld, st);
    } else {
        fft_stage<L, P::R0, 1, true, false, LD_SMEM>(t, sb, ld, st);
        fft_stage<L, P::R1, P::R0, false, false, false>(t, sb, ld, st);
        fft_stage<L, P::R2, P::R0 * P::R1, false, true, false>(t, sb, ld, st);
    }
}

}  // namespace pnp
