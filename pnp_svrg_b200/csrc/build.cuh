// Construction of a batch of CSMRI problems ON THE DEVICE with the package's own kernels (the step before the hot path,
// SURVEY section 8(f) rank 2; reference: problems/CSMRI.py:12-41 + problems/problem.py:58-61):
//
//     mask ~ Bernoulli(p)            Y0 = mask o fft2(X)          sigma = sqrt(||Y0||_2 / 10^(snr/10) / H / W)
//     Y = Y0 + mask o N(0, sigma)    Xinit = minmax(|ifft2(Y)|)   support = flatnonzero(mask)
//
// in the layouts of the iteration kernels: selection bytes [H/2][W] (csmri.cuh::set_sel_bits), Y1 / Y2 / Y1n / Y2n
// packed half planes, ascending support lists, line-layout Xinit.  Random numbers are counter based (a hash of seed,
// problem, stream and the flat k-space position), so every kernel can recompute the draw of any position -- the mask
// of a mirrored position, the noise of both members of a Hermitian pair -- without a stored random field.
//
//   k_build_bits        mask bits in the packed layout
//   k_support_count/scan/write   stream compaction of the mask in ascending position order + M0
//   k_cols_build        FFT over the lines of every packed row of the line pass's output (k_lines_r2c), i.e. the
//                       packed 2-D spectrum F, written to Y1 (row 0: DC row -> Y1, Nyquist row -> Y1n) + ||Y0||^2
//   k_build_noise       sigma from the norm; Y1 = m (F + sigma n), Y2 = m' (F + sigma n') (mirror position) and the
//                       rotated planes -iY1, +iY2 whose "gradient" is the imaginary part of ifft2(Y)
//   (pnp_b200.cu)       Re / Im of ifft2(Y) through the iteration's own column + inverse line passes on a zero spectrum
//   k_build_abs / k_build_norm   |.|, per-problem min / max, (x - min) / (max - min)
#pragma once
#include "csmri.cuh"

namespace pnp {

__device__ __forceinline__ unsigned build_hash(unsigned seed, unsigned img, unsigned stream, unsigned pos) {
    return mix32(seed ^ mix32(pos ^ mix32(img * 0x9e3779b9U + stream * 0x85ebca6bU + 0x27d4eb2fU)));
}
__device__ __forceinline__ bool build_mask_at(unsigned seed, int img, int pos, float p) {
    return (float)(build_hash(seed, (unsigned)img, 0u, (unsigned)pos) >> 8) * (1.0f / 16777216.0f) < p;
}
// N(0, 1) by Box-Muller on two 24-bit uniforms of the position
__device__ __forceinline__ float build_noise_at(unsigned seed, int img, int pos) {
    const float u1 = (float)((build_hash(seed, (unsigned)img, 1u, (unsigned)pos) >> 8) + 1u) * (1.0f / 16777216.0f);
    const float u2 = (float)(build_hash(seed, (unsigned)img, 2u, (unsigned)pos) >> 8) * (1.0f / 16777216.0f);
    return sqrtf(-2.0f * logf(u1)) * cospif(2.0f * u2);
}

__global__ void __launch_bounds__(256)
k_build_bits(unsigned char* __restrict__ bits, int H, int W, unsigned seed, const float* __restrict__ p) {
    const int img = blockIdx.y, hp = H / 2;
    const float pi = p[img];
    unsigned char* bi = bits + (long long)img * hp * W;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < hp * W; e += gridDim.x * blockDim.x) {
        const int kyp = e / W, kx = e - kyp * W;
        const int mx = (W - kx) % W, my = (H - kyp) % H;
        unsigned b = (build_mask_at(seed, img, kyp * W + kx, pi) ? 1u : 0u) | (build_mask_at(seed, img, my * W + mx, pi) ? 2u : 0u);
        if (kyp == 0) b |= (build_mask_at(seed, img, hp * W + kx, pi) ? 4u : 0u) | (build_mask_at(seed, img, hp * W + mx, pi) ? 8u : 0u);
        bi[e] = (unsigned char)b;
    }
}

// ---- support lists: positions of the mask in ascending order (np.flatnonzero), chunks of 1024 positions per CTA ----
#define BUILD_CHUNK 1024
__global__ void __launch_bounds__(256)
k_support_count(int* __restrict__ chunk_count, int N, unsigned seed, const float* __restrict__ p) {
    const int img = blockIdx.y, chunk = blockIdx.x, nchunks = gridDim.x;
    const float pi = p[img];
    int c = 0;
#pragma unroll
    for (int k = 0; k < BUILD_CHUNK / 256; ++k) {
        const int pos = chunk * BUILD_CHUNK + k * 256 + threadIdx.x;
        c += (pos < N && build_mask_at(seed, img, pos, pi)) ? 1 : 0;
    }
    __shared__ int s[8];
    c = __reduce_add_sync(0xffffffffu, c);
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int k = 0; k < 8; ++k) t += s[k];
        chunk_count[(long long)img * nchunks + chunk] = t;
    }
}
// exclusive scan of the chunk counts of one problem (one CTA per problem), M0 and 1 / M0
__global__ void __launch_bounds__(1024)
k_support_scan(int* __restrict__ chunk_count, int nchunks, int* __restrict__ m0, float* __restrict__ inv_m0) {
    const int img = blockIdx.x;
    int* cc = chunk_count + (long long)img * nchunks;
    __shared__ int s_warp[32];
    __shared__ int s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < nchunks; base += 1024) {
        const int i = base + threadIdx.x;
        const int v = i < nchunks ? cc[i] : 0;
        int x = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int y = __shfl_up_sync(0xffffffffu, x, d);
            if ((threadIdx.x & 31) >= d) x += y;
        }
        if ((threadIdx.x & 31) == 31) s_warp[threadIdx.x >> 5] = x;
        __syncthreads();
        if (threadIdx.x < 32) {
            int w = s_warp[threadIdx.x];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int y = __shfl_up_sync(0xffffffffu, w, d);
                if (threadIdx.x >= d) w += y;
            }
            s_warp[threadIdx.x] = w;
        }
        __syncthreads();
        const int incl = x + ((threadIdx.x >> 5) ? s_warp[(threadIdx.x >> 5) - 1] : 0) + s_carry;
        if (i < nchunks) cc[i] = incl - v;
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        m0[img] = s_carry;
        if (inv_m0) inv_m0[img] = 1.0f / (float)(s_carry > 0 ? s_carry : 1);
    }
}
__global__ void __launch_bounds__(256)
k_support_write(int* __restrict__ support, long long sup_stride, const int* __restrict__ chunk_off, int N, unsigned seed,
                const float* __restrict__ p) {
    const int img = blockIdx.y, chunk = blockIdx.x, nchunks = gridDim.x;
    const float pi = p[img];
    __shared__ int s_warp[8];
    int* out = support + (long long)img * sup_stride;
    int base = chunk_off[(long long)img * nchunks + chunk];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll 1
    for (int k = 0; k < BUILD_CHUNK / 256; ++k) {
        const int pos = chunk * BUILD_CHUNK + k * 256 + threadIdx.x;
        const bool m = pos < N && build_mask_at(seed, img, pos, pi);
        const unsigned bal = __ballot_sync(0xffffffffu, m);
        if (lane == 0) s_warp[warp] = __popc(bal);
        __syncthreads();
        int off = base, tot = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) {
            const int c = s_warp[w];
            if (w < warp) off += c;
            tot += c;
        }
        if (m) out[off + __popc(bal & ((1u << lane) - 1u))] = pos;
        base += tot;
        __syncthreads();
    }
}

// ---- packed 2-D spectrum of every problem from the line pass's output -------------------------------------------------
// S[kyp][c] (k_lines_r2c of the image, packed: row 0 = DC + i Nyquist) -> F[kyp][kx] = FFT over c.  Tasks 1 .. hp-1 are the
// packed rows; task 0 is packed row 0, whose transform C splits into the DC row F[0][kx] = (C[kx] + conj C[-kx]) / 2 and
// the Nyquist row F[H/2][kx] = (C[kx] - conj C[-kx]) / 2i.  ||mask o F||^2 over the FULL plane is accumulated from the
// bits: a packed entry stands for (kyp, kx) [bit0] and its Hermitian mirror [bit1] (kyp >= 1); row 0 counts bit0 (DC)
// and bit2 (Nyquist) only -- bits 1 / 3 there are the same rows seen from -kx.
template <int L, int NC>
__global__ void __launch_bounds__(NC * (L / FftPlan<L>::EPT))
k_cols_build(const float2* __restrict__ S, const unsigned char* __restrict__ bits, float2* __restrict__ Y1,
             float2* __restrict__ Y1n, int hp, double* __restrict__ norm2) {
    constexpr int T = fft_threads<L>();
    constexpr int EPT = FftPlan<L>::EPT;
    constexpr int PL = fft_plane<L>();
    extern __shared__ float smem[];
    __shared__ float s_red[32];
    const int g = threadIdx.x / T, t = threadIdx.x % T;
    const int task = blockIdx.x * NC + g;
    const bool active = task < hp;
    const int img = blockIdx.y;
    const SmemBuf sb{smem + g * 2 * PL, smem + g * 2 * PL + PL};
    const float2* Sc = S + ((long long)img * hp + (active ? task : 0)) * L;
    const unsigned char* bc = bits + ((long long)img * hp + (active ? task : 0)) * L;
    float2* y1 = Y1 + ((long long)img * hp + (active ? task : 0)) * L;
    auto ld = [&](int c) -> float2 { return active ? Sc[c] : make_float2(0.f, 0.f); };
    auto st = [&](int idx, float2 v) { sb.put(idx, v); };
    fft_forward<L, false>(t, sb, ld, st);
    __syncthreads();
    float acc = 0.f;
    if (active && task > 0) {
#pragma unroll 4
        for (int m = 0; m < EPT; ++m) {
            const int kx = t + m * T;
            const float2 f = sb.get(kx);
            const unsigned b = bc[kx];
            y1[kx] = f;
            acc += (float)((b & 1u) + ((b >> 1) & 1u)) * (f.x * f.x + f.y * f.y);
        }
    } else if (active) {
        float2* y1n = Y1n + (long long)img * L;
        for (int kx = t; kx < L; kx += T) {
            const int km = (L - kx) % L;
            const float2 ck = sb.get(kx), cm = sb.get(km);
            const float2 fdc = make_float2(0.5f * (ck.x + cm.x), 0.5f * (ck.y - cm.y));
            const float2 fny = make_float2(0.5f * (ck.y + cm.y), 0.5f * (cm.x - ck.x));
            const unsigned b = bc[kx];
            y1[kx] = fdc;
            y1n[kx] = fny;
            acc += (float)(b & 1u) * (fdc.x * fdc.x + fdc.y * fdc.y) + (float)((b >> 2) & 1u) * (fny.x * fny.x + fny.y * fny.y);
        }
    }
    acc = warp_sum_f(acc);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float tsum = 0.f;
        for (int k = 0; k < (int)(blockDim.x >> 5); ++k) tsum += s_red[k];
        atomicAdd(norm2 + img, (double)tsum);
    }
}

// sigma (problems/problem.py:58-61: the norm, not its square, as in the reference), measurements and their rotated copies.
// On entry Y1 / Y1n hold the unmasked spectrum F; rot = (Y1r, Y2r, Y1nr, Y2nr) receive -i Y1, +i Y2 (the Hermitian part
// of -iY is -i (Y1 - Y2) / 2: the real inverse of that spectrum is Im ifft2(Y)).
__global__ void __launch_bounds__(256)
k_build_noise(float2* __restrict__ Y1, float2* __restrict__ Y2, float2* __restrict__ Y1n, float2* __restrict__ Y2n,
              float2* __restrict__ Y1r, float2* __restrict__ Y2r, float2* __restrict__ Y1nr, float2* __restrict__ Y2nr,
              const unsigned char* __restrict__ bits, int H, int W, unsigned seed, const float* __restrict__ snr,
              const double* __restrict__ norm2, float* __restrict__ sigma_out) {
    const int img = blockIdx.y, hp = H / 2;
    const float sigma = (float)sqrt(sqrt(norm2[img]) / pow(10.0, (double)snr[img] / 10.0) / (double)H / (double)W);
    if (blockIdx.x == 0 && threadIdx.x == 0) sigma_out[img] = sigma;
    const long long pbase = (long long)img * hp * W;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < hp * W + W; e += gridDim.x * blockDim.x) {
        const bool nyq = e >= hp * W;
        const int kyp = nyq ? 0 : e / W, kx = nyq ? e - hp * W : e - kyp * W;
        const int mx = (W - kx) % W;
        const int ky = nyq ? hp : kyp, my = nyq ? hp : (H - kyp) % H;
        unsigned b = bits[pbase + (long long)kyp * W + kx];
        if (nyq) b >>= 2;
        float2* p1 = nyq ? Y1n + (long long)img * W + kx : Y1 + pbase + e;
        float2* p2 = nyq ? Y2n + (long long)img * W + kx : Y2 + pbase + e;
        float2* r1 = nyq ? Y1nr + (long long)img * W + kx : Y1r + pbase + e;
        float2* r2 = nyq ? Y2nr + (long long)img * W + kx : Y2r + pbase + e;
        const float2 f = *p1;
        float2 a = make_float2(0.f, 0.f), c = make_float2(0.f, 0.f);
        if (b & 1u) a = make_float2(f.x + sigma * build_noise_at(seed, img, ky * W + kx), f.y);
        if (b & 2u) c = make_float2(f.x + sigma * build_noise_at(seed, img, my * W + mx), f.y);
        *p1 = a;
        *p2 = c;
        *r1 = make_float2(a.y, -a.x);
        *r2 = make_float2(-c.y, c.x);
    }
}

// x0 = sqrt(re^2 + im^2) into `re`, min / max per problem (non-negative floats order like their bit patterns)
__global__ void __launch_bounds__(256)
k_build_abs(float* __restrict__ re, const float* __restrict__ im, long long N, unsigned* __restrict__ minmax) {
    const int img = blockIdx.y;
    float* r = re + (long long)img * N;
    const float* q = im + (long long)img * N;
    float lo = 3.0e38f, hi = 0.f;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < N; e += (long long)gridDim.x * blockDim.x) {
        const float a = r[e], b = q[e];
        const float v = sqrtf(a * a + b * b);
        r[e] = v;
        lo = fminf(lo, v);
        hi = fmaxf(hi, v);
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, d));
        hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, d));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMin(minmax + 2 * img, __float_as_uint(lo));
        atomicMax(minmax + 2 * img + 1, __float_as_uint(hi));
    }
}
__global__ void __launch_bounds__(256)
k_build_norm(float* __restrict__ x, long long N, const unsigned* __restrict__ minmax) {
    const int img = blockIdx.y;
    float* r = x + (long long)img * N;
    const float lo = __uint_as_float(minmax[2 * img]), hi = __uint_as_float(minmax[2 * img + 1]);
    const float inv = 1.0f / (hi - lo);
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < N; e += (long long)gridDim.x * blockDim.x)
        r[e] = (r[e] - lo) * inv;
}
__global__ void k_build_init(double* __restrict__ norm2, unsigned* __restrict__ minmax, int nb) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nb) {
        norm2[i] = 0.0;
        minmax[2 * i] = 0x7f7fffffu;
        minmax[2 * i + 1] = 0u;
    }
}

}  // namespace pnp
