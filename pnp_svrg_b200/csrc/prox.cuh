// Prox-step kernels on TRANSPOSED images (line c = original column c, contiguous, length H):
//
//   k_sigma_mad     skimage estimate_sigma(z0, multichannel=True, average_sigmas=True)
//                   (algorithms/pnp_svrg.py:71 and the five sibling call sites): per column the
//                   db2 detail coefficients (symmetric extension), median of |d| over d != 0,
//                   / Phi^-1(0.75); the mean over columns is accumulated in a double slot.
//   k_haar_bayes    skimage denoise_wavelet(BayesShrink, db1, soft, multichannel=True)
//                   (denoisers/TV.py:24,26): per column multi-level Haar, per-level threshold
//                   sigma^2 / sqrt(max(mean(d^2) - sigma^2, eps)), soft shrink, inverse -- one
//                   warp per column, pyramid in registers + warp shuffles; the squared error
//                   against the ground truth (problems/problem.py:33-35) is reduced in the epilogue.
//   k_sq_err        stand-alone sum((z - xrec)^2) for Problem.PSNR.
#pragma once
#include <cuda_runtime.h>

namespace pnp {

__device__ __forceinline__ int warp_sum_i(int v) { return __reduce_add_sync(0xffffffffu, v); }
__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// log slot addressing shared by the prox kernels: slot s of image i lives at base[s * batch + i]
__device__ __forceinline__ double* slot_ptr(double* base, const int* slot, int batch, int img) {
    return base + (long long)(slot ? *slot : 0) * batch + img;
}

// Warp-collective sigma estimate of ONE line x[0..L) (global or shared memory): db2 detail
// coefficients with half-sample symmetric extension, exact median of |d| over d != 0, / Phi^-1(0.75).
// The keys (float bit patterns of |d|, monotone in the value) stay in registers.  Selection: ONE pass builds a
// 256-bin histogram over the keys' own range [min, max] (bins = 2^s consecutive bit patterns, i.e. spaced like the
// exponent: ~21 bins per octave, so the bin that holds the median has ~1-2 % of the keys), a warp scan finds the
// bin of rank k1, the <= 32 keys of that bin are gathered and ranked directly.  If a bin holds more than 32 keys
// the bit search continues inside it (exact for any input; the 31-step search from the top bit that this replaces
// was 2/3 of the instructions of the estimate).
#define PNP_SIG_SCRATCH 320          // 32-bit words of shared scratch per warp: 256 bins + 32 candidates (+ pad)
// per-warp scratch (32-bit words) of the fused prox kernels: sigma selection, then Haar transposition (L/4 padded floats)
template <int L> __host__ __device__ constexpr int prox_scratch() { return L >= 2048 && L <= 2048 ? 640 : PNP_SIG_SCRATCH; }
template <int L>
__device__ __forceinline__ double line_sigma_mad(const float* __restrict__ x, int lane, unsigned* scratch /* PNP_SIG_SCRATCH words per warp, shared */) {
    constexpr int NO = (L + 3) / 2;            // db2 detail coefficients per line
    constexpr int PER = (NO + 31) / 32;
    // pywt dec_hi of db2; out[o] = sum_j h[j] * x_ext[2o + 1 - j]
    const float h0 = -0.48296291314469025f, h1 = 0.836516303737469f,
                h2 = -0.22414386804185735f, h3 = -0.12940952255092145f;
    unsigned a[PER];
    int nnz = 0;
    unsigned mn = 0xffffffffu, mx = 0u;
#pragma unroll
    for (int i = 0; i < PER; ++i) {
        const int o = lane + 32 * i;
        a[i] = 0xffffffffu;
        if (o < NO) {
            float d;
            if (i > 0 && 32 * i + 31 <= (L - 2) / 2) {
                // interior (compile-time test): samples 2o-2 .. 2o+1 are in range, two aligned 8-byte loads
                const float2 lo = *reinterpret_cast<const float2*>(x + 2 * o - 2), hi = *reinterpret_cast<const float2*>(x + 2 * o);
                d = fmaf(h0, hi.y, fmaf(h1, hi.x, fmaf(h2, lo.y, h3 * lo.x)));
            } else {
                int i0 = 2 * o + 1, i1 = 2 * o, i2 = 2 * o - 1, i3 = 2 * o - 2;
                // half-sample symmetric extension
                i0 = i0 >= L ? 2 * L - 1 - i0 : i0;
                i1 = i1 >= L ? 2 * L - 1 - i1 : i1;
                i2 = i2 < 0 ? -1 - i2 : (i2 >= L ? 2 * L - 1 - i2 : i2);
                i3 = i3 < 0 ? -1 - i3 : (i3 >= L ? 2 * L - 1 - i3 : i3);
                d = fmaf(h0, x[i0], fmaf(h1, x[i1], fmaf(h2, x[i2], h3 * x[i3])));
            }
            if (d != 0.f) {
                a[i] = __float_as_uint(fabsf(d));
                ++nnz;
                mn = min(mn, a[i]);
                mx = max(mx, a[i]);
            }
        }
    }
    nnz = warp_sum_i(nnz);
    if (nnz == 0) return __longlong_as_double(0x7ff8000000000000LL);      // median of nothing = NaN
    const int k1 = (nnz - 1) >> 1, k2 = nnz >> 1;
    mn = __reduce_min_sync(0xffffffffu, mn);
    mx = __reduce_max_sync(0xffffffffu, mx);
    // state of the search: #keys < res is c_lo, #keys < res + 2^(bit+1) is c_hi, the k1-th smallest key lies in between
    unsigned res;
    int c_lo, c_hi, bit;
    {
        const int dbits = 32 - __clz(mx ^ mn);           // the keys agree above bit dbits (0 when all are equal)
        const int s = dbits > 8 ? dbits - 8 : 0;         // (mx >> s) - (mn >> s) < 256
        const unsigned bin0 = mn >> s;
        unsigned* hist = scratch;
        reinterpret_cast<uint4*>(hist)[lane] = make_uint4(0u, 0u, 0u, 0u);
        reinterpret_cast<uint4*>(hist)[lane + 32] = make_uint4(0u, 0u, 0u, 0u);
        __syncwarp();
#pragma unroll
        for (int i = 0; i < PER; ++i)
            if (a[i] != 0xffffffffu) atomicAdd(&hist[(a[i] >> s) - bin0], 1u);
        __syncwarp();
        // lane owns bins 8*lane .. 8*lane + 7
        const uint4 ha = reinterpret_cast<const uint4*>(hist)[2 * lane], hb = reinterpret_cast<const uint4*>(hist)[2 * lane + 1];
        const int hh[8] = {(int)ha.x, (int)ha.y, (int)ha.z, (int)ha.w, (int)hb.x, (int)hb.y, (int)hb.z, (int)hb.w};
        const int part = hh[0] + hh[1] + hh[2] + hh[3] + hh[4] + hh[5] + hh[6] + hh[7];
        int incl = part;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        int below = incl - part;
        const unsigned hit = __ballot_sync(0xffffffffu, below <= k1 && k1 < incl);
        const int src = __ffs(hit) - 1;
        int b = 0, cnt = 0, found = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (!found && below + hh[j] > k1) { b = 8 * lane + j; cnt = hh[j]; found = 1; }
            if (!found) below += hh[j];
        }
        b = __shfl_sync(0xffffffffu, b, src);
        cnt = __shfl_sync(0xffffffffu, cnt, src);
        below = __shfl_sync(0xffffffffu, below, src);
        __syncwarp();                           // the histogram is dead: its words are reused for the candidates below
        res = (bin0 + (unsigned)b) << s;
        c_lo = below;
        c_hi = below + cnt;
        bit = s - 1;
    }
    for (; bit >= 0 && c_hi - c_lo > 32; --bit) {
        const unsigned cand = res | (1u << bit);
        int c = 0;
#pragma unroll
        for (int i = 0; i < PER; ++i) c += a[i] < cand;
        c = warp_sum_i(c);
        if (c <= k1) { res = cand; c_lo = c; } else { c_hi = c; }
    }
    unsigned v1;
    if (bit < 0) {
        v1 = res;                               // every bit decided
    } else {
        // candidates: keys in [res, res + 2^(bit+1)); exactly c_hi - c_lo <= 32 of them
        const unsigned width = 2u << bit;       // bit <= 29 here, no overflow; sentinels are >= 2^31 away from res
        unsigned* cand_buf = scratch + 256;
        int base = 0;
#pragma unroll
        for (int i = 0; i < PER; ++i) {
            const bool in = (a[i] - res) < width;
            const unsigned m = __ballot_sync(0xffffffffu, in);
            if (m) {
                if (in) cand_buf[base + __popc(m & ((1u << lane) - 1u))] = a[i];
                base += __popc(m);
            }
        }
        __syncwarp();
        const unsigned mine = lane < base ? cand_buf[lane] : 0xffffffffu;
        __syncwarp();
        // rank of every candidate among the candidates (ties broken by lane)
        int rnk = 0;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            const unsigned u = __shfl_sync(0xffffffffu, mine, j);
            rnk += (u < mine) || (u == mine && j < lane);
        }
        const int target = k1 - c_lo;
        const unsigned hit = __ballot_sync(0xffffffffu, rnk == target && mine != 0xffffffffu);
        v1 = __shfl_sync(0xffffffffu, mine, __ffs(hit) - 1);
    }
    unsigned v2 = v1;
    if (k2 != k1) {
        int c = 0;
        unsigned mnv = 0xffffffffu;
#pragma unroll
        for (int i = 0; i < PER; ++i) { c += a[i] <= v1; if (a[i] > v1) mnv = min(mnv, a[i]); }
        c = warp_sum_i(c);
        mnv = __reduce_min_sync(0xffffffffu, mnv);
        if (c < k2 + 1) v2 = mnv;
    }
    return 0.5 * ((double)__uint_as_float(v1) + (double)__uint_as_float(v2)) / 0.6744897501960817;
}

// NL lines by ONE warp, step by step side by side: the estimate of a short line is a chain of ~60 dependent warp
// collectives (reductions, scans, ballots, shuffles) whose latency nothing hides when a warp works on a single line;
// with the lines interleaved the chains overlap.  Same arithmetic, same result as line_sigma_mad, line by line.
// x[j]: line j (shared memory); scratch[j]: PNP_SIG_SCRATCH words each.
// NORM = false returns the median itself (the caller divides the SUM over its lines by Phi^-1(0.75) once: a double
// division per line is ~10 % of the estimate's instructions for a 256-sample line).
template <int L, int NL, bool NORM = true>
__device__ __forceinline__ void line_sigma_mad_n(const float* const (&x)[NL], int lane, unsigned* const (&scratch)[NL], double (&out)[NL]) {
    constexpr int NO = (L + 3) / 2;
    constexpr int PER = (NO + 31) / 32;
    const float h0 = -0.48296291314469025f, h1 = 0.836516303737469f,
                h2 = -0.22414386804185735f, h3 = -0.12940952255092145f;
    unsigned a[NL][PER], mn[NL], mx[NL], res[NL], v1[NL];
    int nnz[NL], c_lo[NL], c_hi[NL], bit[NL], k1[NL], k2[NL];
    bool empty[NL];
#pragma unroll
    for (int j = 0; j < NL; ++j) {
        nnz[j] = 0; mn[j] = 0xffffffffu; mx[j] = 0u;
#pragma unroll
        for (int i = 0; i < PER; ++i) {
            const int o = lane + 32 * i;
            a[j][i] = 0xffffffffu;
            if (o < NO) {
                float d;
                if (i > 0 && 32 * i + 31 <= (L - 2) / 2) {
                    const float2 lo = *reinterpret_cast<const float2*>(x[j] + 2 * o - 2), hi = *reinterpret_cast<const float2*>(x[j] + 2 * o);
                    d = fmaf(h0, hi.y, fmaf(h1, hi.x, fmaf(h2, lo.y, h3 * lo.x)));
                } else {
                    int i0 = 2 * o + 1, i1 = 2 * o, i2 = 2 * o - 1, i3 = 2 * o - 2;
                    i0 = i0 >= L ? 2 * L - 1 - i0 : i0;
                    i1 = i1 >= L ? 2 * L - 1 - i1 : i1;
                    i2 = i2 < 0 ? -1 - i2 : (i2 >= L ? 2 * L - 1 - i2 : i2);
                    i3 = i3 < 0 ? -1 - i3 : (i3 >= L ? 2 * L - 1 - i3 : i3);
                    d = fmaf(h0, x[j][i0], fmaf(h1, x[j][i1], fmaf(h2, x[j][i2], h3 * x[j][i3])));
                }
                if (d != 0.f) {
                    a[j][i] = __float_as_uint(fabsf(d));
                    ++nnz[j];
                    mn[j] = min(mn[j], a[j][i]);
                    mx[j] = max(mx[j], a[j][i]);
                }
            }
        }
    }
#pragma unroll
    for (int j = 0; j < NL; ++j) nnz[j] = warp_sum_i(nnz[j]);
#pragma unroll
    for (int j = 0; j < NL; ++j) {
        empty[j] = nnz[j] == 0;                 // median of nothing = NaN: run the steps on one dummy key, discard the result
        if (empty[j]) { nnz[j] = 1; if (lane == 0) { a[j][0] = 0x3f800000u; mn[j] = mx[j] = a[j][0]; } }
        k1[j] = (nnz[j] - 1) >> 1; k2[j] = nnz[j] >> 1;
    }
#pragma unroll
    for (int j = 0; j < NL; ++j) { mn[j] = __reduce_min_sync(0xffffffffu, mn[j]); mx[j] = __reduce_max_sync(0xffffffffu, mx[j]); }
    int s[NL];
    unsigned bin0[NL];
#pragma unroll
    for (int j = 0; j < NL; ++j) {
        const int dbits = 32 - __clz(mx[j] ^ mn[j]);
        s[j] = dbits > 8 ? dbits - 8 : 0;
        bin0[j] = mn[j] >> s[j];
        reinterpret_cast<uint4*>(scratch[j])[lane] = make_uint4(0u, 0u, 0u, 0u);
        reinterpret_cast<uint4*>(scratch[j])[lane + 32] = make_uint4(0u, 0u, 0u, 0u);
    }
    __syncwarp();
#pragma unroll
    for (int j = 0; j < NL; ++j) {
#pragma unroll
        for (int i = 0; i < PER; ++i)
            if (a[j][i] != 0xffffffffu) atomicAdd(&scratch[j][(a[j][i] >> s[j]) - bin0[j]], 1u);
    }
    __syncwarp();
    int hh[NL][8], part[NL], incl[NL];
#pragma unroll
    for (int j = 0; j < NL; ++j) {
        const uint4 ha = reinterpret_cast<const uint4*>(scratch[j])[2 * lane], hb = reinterpret_cast<const uint4*>(scratch[j])[2 * lane + 1];
        hh[j][0] = (int)ha.x; hh[j][1] = (int)ha.y; hh[j][2] = (int)ha.z; hh[j][3] = (int)ha.w;
        hh[j][4] = (int)hb.x; hh[j][5] = (int)hb.y; hh[j][6] = (int)hb.z; hh[j][7] = (int)hb.w;
        part[j] = hh[j][0] + hh[j][1] + hh[j][2] + hh[j][3] + hh[j][4] + hh[j][5] + hh[j][6] + hh[j][7];
        incl[j] = part[j];
    }
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
#pragma unroll
        for (int j = 0; j < NL; ++j) {
            const int v = __shfl_up_sync(0xffffffffu, incl[j], o);
            if (lane >= o) incl[j] += v;
        }
    }
#pragma unroll
    for (int j = 0; j < NL; ++j) {
        int below = incl[j] - part[j];
        const unsigned hit = __ballot_sync(0xffffffffu, below <= k1[j] && k1[j] < incl[j]);
        const int src = __ffs(hit) - 1;
        int b = 0, cnt = 0, found = 0;
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
            if (!found && below + hh[j][jj] > k1[j]) { b = 8 * lane + jj; cnt = hh[j][jj]; found = 1; }
            if (!found) below += hh[j][jj];
        }
        b = __shfl_sync(0xffffffffu, b, src);
        cnt = __shfl_sync(0xffffffffu, cnt, src);
        below = __shfl_sync(0xffffffffu, below, src);
        res[j] = (bin0[j] + (unsigned)b) << s[j];
        c_lo[j] = below;
        c_hi[j] = below + cnt;
        bit[j] = s[j] - 1;
    }
    __syncwarp();                               // the histograms are dead: their words are reused for the candidates below
#pragma unroll
    for (int j = 0; j < NL; ++j) {              // rare: more than 32 keys in the median's bin
        for (; bit[j] >= 0 && c_hi[j] - c_lo[j] > 32; --bit[j]) {
            const unsigned cand = res[j] | (1u << bit[j]);
            int c = 0;
#pragma unroll
            for (int i = 0; i < PER; ++i) c += a[j][i] < cand;
            c = warp_sum_i(c);
            if (c <= k1[j]) { res[j] = cand; c_lo[j] = c; } else { c_hi[j] = c; }
        }
    }
    int base[NL];
#pragma unroll
    for (int j = 0; j < NL; ++j) {
        const unsigned width = bit[j] >= 0 ? (2u << bit[j]) : 0u;       // bit < 0: every bit decided, no candidates
        unsigned* cand_buf = scratch[j] + 256;
        base[j] = 0;
#pragma unroll
        for (int i = 0; i < PER; ++i) {
            const bool in = (a[j][i] - res[j]) < width;
            const unsigned m = __ballot_sync(0xffffffffu, in);
            if (in) cand_buf[base[j] + __popc(m & ((1u << lane) - 1u))] = a[j][i];
            base[j] += __popc(m);
        }
    }
    __syncwarp();
    unsigned mine[NL];
    int rnk[NL];
#pragma unroll
    for (int j = 0; j < NL; ++j) { mine[j] = lane < base[j] ? scratch[j][256 + lane] : 0xffffffffu; rnk[j] = 0; }
    __syncwarp();
    int nb = base[0];
#pragma unroll
    for (int j = 1; j < NL; ++j) nb = max(nb, base[j]);
    for (int jj = 0; jj < nb; ++jj) {           // lanes >= base hold the sentinel, which ranks below nothing: stop at the last candidate
#pragma unroll
        for (int j = 0; j < NL; ++j) {
            const unsigned u = __shfl_sync(0xffffffffu, mine[j], jj);
            rnk[j] += (u < mine[j]) || (u == mine[j] && jj < lane);
        }
    }
#pragma unroll
    for (int j = 0; j < NL; ++j) {
        const int target = k1[j] - c_lo[j];
        const unsigned hit = __ballot_sync(0xffffffffu, rnk[j] == target && mine[j] != 0xffffffffu);
        const unsigned vc = __shfl_sync(0xffffffffu, mine[j], hit ? __ffs(hit) - 1 : 0);
        v1[j] = bit[j] < 0 ? res[j] : vc;
    }
#pragma unroll
    for (int j = 0; j < NL; ++j) {
        unsigned v2 = v1[j];
        int c = 0;
        unsigned mnv = 0xffffffffu;
#pragma unroll
        for (int i = 0; i < PER; ++i) { c += a[j][i] <= v1[j]; if (a[j][i] > v1[j]) mnv = min(mnv, a[j][i]); }
        c = warp_sum_i(c);
        mnv = __reduce_min_sync(0xffffffffu, mnv);
        if (k2[j] != k1[j] && c < k2[j] + 1) v2 = mnv;
        const double med = 0.5 * ((double)__uint_as_float(v1[j]) + (double)__uint_as_float(v2));
        out[j] = empty[j] ? __longlong_as_double(0x7ff8000000000000LL) : (NORM ? med / 0.6744897501960817 : med);
    }
}

template <int L>
__global__ void __launch_bounds__(128)
k_sigma_mad(const float* __restrict__ z, int nlines, long long img_stride, double* __restrict__ sig_log,
            const int* __restrict__ slot, int batch) {
    __shared__ __align__(16) unsigned scratch[4][PNP_SIG_SCRATCH];
    const int lane = threadIdx.x & 31;
    const int line = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int img = blockIdx.y;
    if (line >= nlines) return;
    const double sig = line_sigma_mad<L>(z + (long long)img * img_stride + (long long)line * L, lane, scratch[threadIdx.x >> 5]);
    if (lane == 0) atomicAdd(slot_ptr(sig_log, slot, batch, img), sig);
}

// sign(d) * max(|d| - thr, 0) as d - clamp(d, -thr, thr): the same value bit for bit in three instructions
// (thr >= 0; fminf / fmaxf drop a NaN threshold, so that case still yields 0 as the compare-and-select form did)
__device__ __forceinline__ float soft_shrink(float d, float thr) {
    return d - fminf(fmaxf(d, -thr), thr);
}

struct ShrinkParams {
    const double* sig_log;      // optional: sigma_est = slot value / nlines (mean over columns)
    float sigma_est;            // used when sig_log == nullptr
    float sigma_modifier;       // TVDenoiser.sigma_modifier
    float fallback_sigma;       // denoise_strength * decay**t, used when sigma_est <= 0 or NaN
};

// VPL consecutive samples per lane; a line of L samples is spread over WPL warps.  The first
// log2(VPL) levels stay inside a lane, the remaining XL levels are butterflies between lanes of
// one warp (shuffles); only the per-level sum of squares crosses warps (shared memory).
template <int L> struct HaarCfg {
    static constexpr int VPL = L >= 512 ? 16 : L / 32;
    static constexpr int WPL = L / (32 * VPL);
    static constexpr int LIN = VPL == 1 ? 0 : VPL == 2 ? 1 : VPL == 4 ? 2 : VPL == 8 ? 3 : 4;
    static constexpr int LOG2L = L == 32 ? 5 : L == 64 ? 6 : L == 128 ? 7 : L == 256 ? 8 : L == 512 ? 9
                                 : L == 1024 ? 10 : L == 2048 ? 11 : 12;
    static constexpr int LEVELS = LOG2L - 3;            // skimage skips the 3 coarsest scales
    static constexpr int XL = LEVELS - LIN;             // cross-lane levels (2..5)
};

template <int L>
__global__ void __launch_bounds__(HaarCfg<L>::WPL >= 4 ? 32 * HaarCfg<L>::WPL : 128)
k_haar_bayes(const float* __restrict__ zin, float* __restrict__ zout, const float* __restrict__ xrec,
             int nlines, long long img_stride, ShrinkParams sp, double* __restrict__ mse_log,
             const int* __restrict__ slot, int batch) {
    using C = HaarCfg<L>;
    constexpr int VPL = C::VPL, WPL = C::WPL, LIN = C::LIN, LEVELS = C::LEVELS, XL = C::XL;
    constexpr int LPB = WPL >= 4 ? 1 : 4 / WPL;          // lines per CTA
    constexpr float RS2 = 0.70710678118654752f;
    __shared__ float s_ss[LPB][LEVELS][WPL];
    __shared__ float s_err[LPB][WPL];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int lib = warp / WPL, wil = warp % WPL;          // line in block, warp in line
    const int line = blockIdx.x * LPB + lib;
    const int img = blockIdx.y;
    const bool active = line < nlines;
    const long long off = (long long)img * img_stride + (long long)(active ? line : 0) * L + (wil * 32 + lane) * VPL;

    float x[VPL];
    float xr[VPL];                               // ground truth, fetched up front so both streams overlap
    if (VPL >= 4) {
#pragma unroll
        for (int i = 0; i < VPL / 4; ++i) {
            const float4 q = reinterpret_cast<const float4*>(zin + off)[i];
            x[4 * i] = q.x; x[4 * i + 1] = q.y; x[4 * i + 2] = q.z; x[4 * i + 3] = q.w;
        }
        if (xrec) {
#pragma unroll
            for (int i = 0; i < VPL / 4; ++i) {
                const float4 r = reinterpret_cast<const float4*>(xrec + off)[i];
                xr[4 * i] = r.x; xr[4 * i + 1] = r.y; xr[4 * i + 2] = r.z; xr[4 * i + 3] = r.w;
            }
        }
    } else {
#pragma unroll
        for (int i = 0; i < VPL; ++i) { x[i] = zin[off + i]; if (xrec) xr[i] = xrec[off + i]; }
    }

    double se = sp.sig_log ? *slot_ptr(const_cast<double*>(sp.sig_log), slot, batch, img) / (double)nlines
                           : (double)sp.sigma_est;
    const float sigma = (se > 0.0) ? (float)(se * (double)sp.sigma_modifier) : sp.fallback_sigma;
    const float var = sigma * sigma;

    float ss[LEVELS];
#pragma unroll
    for (int l = 0; l < LEVELS; ++l) ss[l] = 0.f;

    // forward pyramid, in place: after level lv the approximation sits at multiples of 2^lv
#pragma unroll
    for (int lv = 1; lv <= LIN; ++lv) {
        const int stride = 1 << lv, half = stride >> 1;
#pragma unroll
        for (int i = 0; i < VPL / stride; ++i) {
            const float p = x[i * stride], q = x[i * stride + half];
            const float d = (p - q) * RS2;
            x[i * stride] = (p + q) * RS2;
            x[i * stride + half] = d;
            ss[lv - 1] = fmaf(d, d, ss[lv - 1]);
        }
    }
    float A[XL + 1], Dx[XL];
    A[0] = x[0];
#pragma unroll
    for (int q = 0; q < XL; ++q) {
        const float p = __shfl_xor_sync(0xffffffffu, A[q], 1 << q);
        const bool ev = (lane & (1 << q)) == 0;
        A[q + 1] = (A[q] + p) * RS2;
        Dx[q] = (ev ? (A[q] - p) : (p - A[q])) * RS2;
        ss[LIN + q] = (lane & ((2 << q) - 1)) == 0 ? Dx[q] * Dx[q] : 0.f;
    }

    float thr[LEVELS];
#pragma unroll
    for (int l = 0; l < LEVELS; ++l) ss[l] = warp_sum_f(ss[l]);
    if (WPL > 1) {
        if (lane == 0) {
#pragma unroll
            for (int l = 0; l < LEVELS; ++l) s_ss[lib][l][wil] = ss[l];
        }
        __syncthreads();
#pragma unroll
        for (int l = 0; l < LEVELS; ++l) {
            float t = 0.f;
#pragma unroll
            for (int w = 0; w < WPL; ++w) t += s_ss[lib][l][w];
            ss[l] = t;
        }
    }
    {   // lane l computes the threshold of level l (one division + root per lane instead of LEVELS), then broadcast
        float mine_ss = ss[0];
#pragma unroll
        for (int l = 1; l < LEVELS; ++l) mine_ss = lane == l ? ss[l] : mine_ss;
        const int lv = lane < LEVELS ? lane : 0;
        const float dvar = mine_ss / (float)(L >> (lv + 1));
        const float tk = var / sqrtf(fmaxf(dvar - var, 2.220446049250313e-16f));
#pragma unroll
        for (int l = 0; l < LEVELS; ++l) thr[l] = __shfl_sync(0xffffffffu, tk, l);
    }

    // inverse pyramid with soft-thresholded details
#pragma unroll
    for (int q = XL - 1; q >= 0; --q) {
        const float d = soft_shrink(Dx[q], thr[LIN + q]);
        const bool ev = (lane & (1 << q)) == 0;
        A[q] = (ev ? (A[q + 1] + d) : (A[q + 1] - d)) * RS2;
    }
    x[0] = A[0];
#pragma unroll
    for (int lv = LIN; lv >= 1; --lv) {
        const int stride = 1 << lv, half = stride >> 1;
#pragma unroll
        for (int i = 0; i < VPL / stride; ++i) {
            const float a = x[i * stride];
            const float d = soft_shrink(x[i * stride + half], thr[lv - 1]);
            x[i * stride] = (a + d) * RS2;
            x[i * stride + half] = (a - d) * RS2;
        }
    }

    float err = 0.f;
    if (active) {
        if (VPL >= 4) {
#pragma unroll
            for (int i = 0; i < VPL / 4; ++i) {
                reinterpret_cast<float4*>(zout + off)[i] = make_float4(x[4 * i], x[4 * i + 1], x[4 * i + 2], x[4 * i + 3]);
                if (xrec) {
                    const float e0 = x[4 * i] - xr[4 * i], e1 = x[4 * i + 1] - xr[4 * i + 1], e2 = x[4 * i + 2] - xr[4 * i + 2],
                                e3 = x[4 * i + 3] - xr[4 * i + 3];
                    err += e0 * e0 + e1 * e1 + e2 * e2 + e3 * e3;
                }
            }
        } else {
#pragma unroll
            for (int i = 0; i < VPL; ++i) {
                zout[off + i] = x[i];
                if (xrec) { const float e = x[i] - xr[i]; err = fmaf(e, e, err); }
            }
        }
    }
    if (xrec && mse_log) {
        err = warp_sum_f(err);
        if (lane == 0 && active) atomicAdd(slot_ptr(mse_log, slot, batch, img), (double)err);
    }
}

// sum((z - xrec)^2) per image -> out[slot][img]
__global__ void __launch_bounds__(256)
k_sq_err(const float* __restrict__ z, const float* __restrict__ xrec, long long n, long long img_stride,
         double* __restrict__ out, const int* __restrict__ slot, int batch) {
    const int img = blockIdx.y;
    const float* a = z + (long long)img * img_stride;
    const float* b = xrec + (long long)img * img_stride;
    float acc = 0.f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float e = a[i] - b[i];
        acc = fmaf(e, e, acc);
    }
    acc = warp_sum_f(acc);
    __shared__ float part[8];
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 8) {
        float v = part[threadIdx.x];
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) v += __shfl_xor_sync(0xffu, v, o);
        if (threadIdx.x == 0) atomicAdd(slot_ptr(out, slot, batch, img), (double)v);
    }
}

}  // namespace pnp

// ------------------------------------------------------------------------------------------------
// Fused prox: estimate_sigma + wavelet BayesShrink + PSNR in ONE cooperative launch.  Every CTA keeps
// its share of the image lines resident in shared memory (2048 lines x 8 KiB over 148 SMs = 113 KiB
// each), staged by TMA bulk copies: phase 1 computes the per-line sigma estimates from shared memory
// and adds them to the global slot, a grid-wide barrier publishes the mean, phase 2 shrinks the same
// resident lines (one warp per line, 512-sample sub-blocks: forward pass for the per-level energies,
// second pass forward + threshold + inverse) and streams them out with the squared error.  The iterate
// is read from HBM/L2 once instead of twice and two launches become one.
#include <cooperative_groups.h>
#include "fft_core.cuh"

namespace pnp {

template <int L> struct HaarSub {
    static constexpr int SB = L >= 512 ? 512 : L;          // samples per sub-block
    static constexpr int VPL = SB / 32;
    static constexpr int NSB = L / SB;
    static constexpr int LIN = VPL == 1 ? 0 : VPL == 2 ? 1 : VPL == 4 ? 2 : VPL == 8 ? 3 : 4;
    static constexpr int LEVELS = HaarCfg<L>::LEVELS;
    static constexpr int XL = LEVELS - LIN;
};

// forward pyramid of one sub-block held as VPL consecutive samples per lane; details stay in x[] / Dx[]
template <int L>
__device__ __forceinline__ void haar_sub_forward(float (&x)[HaarSub<L>::VPL], float (&A)[HaarSub<L>::XL + 1],
                                                 float (&Dx)[HaarSub<L>::XL], float (&ss)[HaarSub<L>::LEVELS], int lane) {
    using C = HaarSub<L>;
    constexpr float RS2 = 0.70710678118654752f;
#pragma unroll
    for (int lv = 1; lv <= C::LIN; ++lv) {
        const int stride = 1 << lv, half = stride >> 1;
#pragma unroll
        for (int i = 0; i < C::VPL / stride; ++i) {
            const float p = x[i * stride], q = x[i * stride + half];
            const float d = (p - q) * RS2;
            x[i * stride] = (p + q) * RS2;
            x[i * stride + half] = d;
            ss[lv - 1] = fmaf(d, d, ss[lv - 1]);
        }
    }
    A[0] = x[0];
#pragma unroll
    for (int q = 0; q < C::XL; ++q) {
        const float p = __shfl_xor_sync(0xffffffffu, A[q], 1 << q);
        const bool ev = (lane & (1 << q)) == 0;
        A[q + 1] = (A[q] + p) * RS2;
        Dx[q] = (ev ? (A[q] - p) : (p - A[q])) * RS2;
        if ((lane & ((2 << q) - 1)) == 0) ss[C::LIN + q] = fmaf(Dx[q], Dx[q], ss[C::LIN + q]);
    }
}

// ---- the two phases of the fused prox on lines resident in shared memory (16 warps, one warp per line) ----
template <int L>
__device__ __noinline__ void prox_phase_sigma(const float* lines, int mine, long long first, int nlines, int batch,
                                                 double* __restrict__ sig_log, int cur_slot, unsigned* scratch /* 16 x prox_scratch<L>() words */) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int l = warp; l < mine; l += 16) {
        const double sig = line_sigma_mad<L>(lines + (long long)l * L, lane, scratch + warp * prox_scratch<L>());
        const int img = (int)((first + l) / nlines);
        if (lane == 0) atomicAdd(sig_log + (long long)cur_slot * batch + img, sig);
    }
}

// ---- 512-sample sub-block in the chunk-cyclic register layout --------------------------------------------------
// Lane l holds the four float4 chunks f = c*32 + l (c < 4) of the sub-block, i.e. samples 4f .. 4f+3: every load and
// store of the warp is one contiguous 512-byte row (no shared-memory bank conflicts, fully coalesced global
// accesses; 16 consecutive samples per lane would be a 16-way conflict per scalar load).  Levels 1-2 stay inside a
// chunk, levels 3-7 are xor butterflies between lanes (four independent chains), the remaining one or two levels
// combine the four chunk tops inside a lane.
template <int LEVELS> struct HaarCC {
    static constexpr int XC = LEVELS - 2 < 5 ? LEVELS - 2 : 5;     // cross-lane levels
    static constexpr int TL = LEVELS - 2 - XC;                     // levels above them (0, 1 or 2)
};

template <int LEVELS, int NCH>
__device__ __forceinline__ void haar_cc_forward(float (&x)[NCH][4], float (&A)[NCH][6], float (&D)[NCH][5], float (&T)[4],
                                                float (&ss)[LEVELS], int lane) {
    using H = HaarCC<LEVELS>;
    static_assert(H::TL == 0 || NCH == 4, "levels above the lane butterflies need the four chunks of a 512-sample block");
    constexpr float RS2 = 0.70710678118654752f;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
        const float a0 = (x[c][0] + x[c][1]) * RS2, d0 = (x[c][0] - x[c][1]) * RS2;
        const float a1 = (x[c][2] + x[c][3]) * RS2, d1 = (x[c][2] - x[c][3]) * RS2;
        ss[0] = fmaf(d0, d0, fmaf(d1, d1, ss[0]));
        const float dd = (a0 - a1) * RS2;
        ss[1] = fmaf(dd, dd, ss[1]);
        x[c][1] = d0; x[c][3] = d1; x[c][2] = dd;                  // details stay where the inverse expects them
        A[c][0] = (a0 + a1) * RS2;
#pragma unroll
        for (int q = 0; q < H::XC; ++q) {
            const float p = __shfl_xor_sync(0xffffffffu, A[c][q], 1 << q);
            const bool ev = (lane & (1 << q)) == 0;
            A[c][q + 1] = (A[c][q] + p) * RS2;
            D[c][q] = (ev ? (A[c][q] - p) : (p - A[c][q])) * RS2;
            if ((lane & ((2 << q) - 1)) == 0) ss[2 + q] = fmaf(D[c][q], D[c][q], ss[2 + q]);
        }
    }
    if constexpr (H::TL >= 1 && NCH == 4) {                       // chunks (0,1) and (2,3) are adjacent 128-sample blocks
        const float t0 = A[0][H::XC], t1 = A[1][H::XC], t2 = A[2][H::XC], t3 = A[3][H::XC];
        T[0] = (t0 - t1) * RS2;
        T[1] = (t2 - t3) * RS2;
        const float u0 = (t0 + t1) * RS2, u1 = (t2 + t3) * RS2;
        if (lane == 0) ss[2 + H::XC] = fmaf(T[0], T[0], fmaf(T[1], T[1], ss[2 + H::XC]));
        T[2] = u0;
        T[3] = u1;
        if (H::TL >= 2) {
            const float dd = (u0 - u1) * RS2;
            if (lane == 0) ss[3 + H::XC] = fmaf(dd, dd, ss[3 + H::XC]);
            T[2] = (u0 + u1) * RS2;                                 // top approximation
            T[3] = dd;
        }
    }
}

template <int LEVELS, int NCH>
__device__ __forceinline__ void haar_cc_inverse(float (&x)[NCH][4], float (&A)[NCH][6], const float (&D)[NCH][5], const float (&T)[4],
                                                const float (&thr)[LEVELS], int lane) {
    using H = HaarCC<LEVELS>;
    constexpr float RS2 = 0.70710678118654752f;
    if constexpr (H::TL >= 1 && NCH == 4) {
        float u0 = T[2], u1 = T[3];
        if (H::TL >= 2) {
            const float dd = soft_shrink(T[3], thr[3 + H::XC]);
            u0 = (T[2] + dd) * RS2;
            u1 = (T[2] - dd) * RS2;
        }
        const float e0 = soft_shrink(T[0], thr[2 + H::XC]), e1 = soft_shrink(T[1], thr[2 + H::XC]);
        A[0][H::XC] = (u0 + e0) * RS2;
        A[1][H::XC] = (u0 - e0) * RS2;
        A[2][H::XC] = (u1 + e1) * RS2;
        A[3][H::XC] = (u1 - e1) * RS2;
    }
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
#pragma unroll
        for (int q = H::XC - 1; q >= 0; --q) {
            const float d = soft_shrink(D[c][q], thr[2 + q]);
            const bool ev = (lane & (1 << q)) == 0;
            A[c][q] = (ev ? (A[c][q + 1] + d) : (A[c][q + 1] - d)) * RS2;
        }
        const float dd = soft_shrink(x[c][2], thr[1]);
        const float a0 = (A[c][0] + dd) * RS2, a1 = (A[c][0] - dd) * RS2;
        const float d0 = soft_shrink(x[c][1], thr[0]), d1 = soft_shrink(x[c][3], thr[0]);
        x[c][0] = (a0 + d0) * RS2; x[c][1] = (a0 - d0) * RS2;
        x[c][2] = (a1 + d1) * RS2; x[c][3] = (a1 - d1) * RS2;
    }
}

// Lines of >= 512 samples: the 512-sample sub-blocks of all resident lines are independent tasks spread over the 16
// warps (no lock step between warps): pass A transforms every sub-block for its per-level detail energies, one CTA
// barrier, pass B transforms again, shrinks with the line's thresholds, inverts and streams out.  Shorter lines: one
// warp per line, a single pass.
#define PROX_MAX_TASKS 112          // sub-blocks per CTA: 200 KiB of lines / 2 KiB

// ---- one warp per line for 512 .. 2048 samples ------------------------------------------------------------------
// The lane holds NCH = L/128 float4 chunks in the chunk-cyclic layout (conflict-free shared-memory rows, coalesced
// global rows).  Levels 1-2 stay inside a chunk.  The L/4 level-2 approximations are then TRANSPOSED through the
// line's own shared memory (free once the samples sit in registers) so that every lane owns NCH consecutive ones:
// log2(NCH) more levels run inside the lane on a single chain, and only the last two levels are butterflies between
// lanes.  All per-level energies of the line live in one warp, so the thresholds need no second pass over the data
// and no cross-warp exchange: the sub-block version this replaces ran 4x the instructions (5 shuffle levels on four
// chains per 512 samples, the forward transform twice).
template <int L> struct HaarWL {
    static constexpr int NCH = L / 128;                                   // float4 chunks per lane: 4, 8, 16
    static constexpr int LG = NCH == 4 ? 2 : (NCH == 8 ? 3 : 4);          // in-lane levels after the transposition
    static constexpr int LEVELS = HaarCfg<L>::LEVELS;                     // = 2 + LG + 2
};
// index of approximation f in the transposition scratch: 4 pad floats per 16 keep the 16-byte block reads of a
// quarter warp on distinct banks
__device__ __forceinline__ int haar_wl_pad(int f) { return f + ((f >> 4) << 2); }

__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// What a warp keeps in REGISTERS across the grid barrier of the fused prox: the forward Haar pyramid of its line
// (details only: the transform does not need sigma), and the per-level energies (lane k holds level k).
template <int L> struct ProxLine {
    static constexpr int NCH = HaarWL<L>::NCH;
    float x[NCH][4];            // [c][1], [c][3]: level-1 details, [c][2]: level-2 detail of chunk c
    float v[NCH];               // in-lane pyramid over the lane's NCH level-2 approximations (v[0]: superseded by A)
    float A[3], Dx[2];          // the two cross-lane levels
    float level_ss;             // lane k < LEVELS: sum of squares of the level-(k+1) details of the line
};

// forward pyramid of the line at `sl` (shared memory; it becomes scratch)
template <int L>
__device__ __forceinline__ void prox_line_forward(float* sl, int lane, ProxLine<L>& st) {
    using W = HaarWL<L>;
    constexpr int NCH = W::NCH, LG = W::LG, LEVELS = W::LEVELS;
    constexpr float RS2 = 0.70710678118654752f;
    static_assert(LEVELS == LG + 4, "level split of the one-warp-per-line shrink");
    const float4* s4 = reinterpret_cast<const float4*>(sl);
    float ss[LEVELS];
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
        const float4 q = s4[c * 32 + lane];
        st.x[c][0] = q.x; st.x[c][1] = q.y; st.x[c][2] = q.z; st.x[c][3] = q.w;
    }
#pragma unroll
    for (int k = 0; k < LEVELS; ++k) ss[k] = 0.f;
    __syncwarp();                      // every lane holds its samples: the line's shared memory is scratch now
#pragma unroll
    for (int c = 0; c < NCH; ++c) {    // levels 1, 2; details stay where the inverse expects them
        const float a0 = (st.x[c][0] + st.x[c][1]) * RS2, d0 = (st.x[c][0] - st.x[c][1]) * RS2;
        const float a1 = (st.x[c][2] + st.x[c][3]) * RS2, d1 = (st.x[c][2] - st.x[c][3]) * RS2;
        ss[0] = fmaf(d0, d0, fmaf(d1, d1, ss[0]));
        const float dd = (a0 - a1) * RS2;
        ss[1] = fmaf(dd, dd, ss[1]);
        st.x[c][1] = d0; st.x[c][3] = d1; st.x[c][2] = dd;
        sl[haar_wl_pad(c * 32 + lane)] = (a0 + a1) * RS2;
    }
    __syncwarp();
    {
        const float4* b4 = reinterpret_cast<const float4*>(sl + haar_wl_pad(lane * NCH));
#pragma unroll
        for (int j = 0; j < NCH / 4; ++j) {
            const float4 q = b4[j];
            st.v[4 * j] = q.x; st.v[4 * j + 1] = q.y; st.v[4 * j + 2] = q.z; st.v[4 * j + 3] = q.w;
        }
    }
#pragma unroll
    for (int lv = 1; lv <= LG; ++lv) {
        const int stride = 1 << lv, half = stride >> 1;
#pragma unroll
        for (int i = 0; i < NCH / stride; ++i) {
            const float p = st.v[i * stride], q = st.v[i * stride + half];
            const float d = (p - q) * RS2;
            st.v[i * stride] = (p + q) * RS2;
            st.v[i * stride + half] = d;
            ss[1 + lv] = fmaf(d, d, ss[1 + lv]);
        }
    }
    st.A[0] = st.v[0];
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        const float p = __shfl_xor_sync(0xffffffffu, st.A[q], 1 << q);
        const bool ev = (lane & (1 << q)) == 0;
        st.A[q + 1] = (st.A[q] + p) * RS2;
        st.Dx[q] = (ev ? (st.A[q] - p) : (p - st.A[q])) * RS2;
        if ((lane & ((2 << q) - 1)) == 0) ss[2 + LG + q] = st.Dx[q] * st.Dx[q];
    }
    float mine_ss = 0.f;
#pragma unroll
    for (int k = 0; k < LEVELS; ++k) {
        const float e = warp_sum_f(ss[k]);
        mine_ss = lane == k ? e : mine_ss;
    }
    st.level_ss = mine_ss;
}

// thresholds from the energies, shrink, inverse pyramid, store; returns the lane's part of sum((out - xrec)^2).
// `tscr`: prox_scratch<L>() floats of shared scratch of this warp; `xr4`: ground truth of the line (shared or global).
// `keep4` (optional): shared-memory copy of the result in the line's own layout (may be the buffer xr4 points to: a lane
// reads a ground-truth chunk before it overwrites it)
template <int L>
__device__ __forceinline__ float prox_line_inverse(ProxLine<L>& st, float* tscr, int lane, float var, float4* zo4, const float4* xr4,
                                                   float4* keep4 = nullptr) {
    using W = HaarWL<L>;
    constexpr int NCH = W::NCH, LG = W::LG, LEVELS = W::LEVELS;
    constexpr float RS2 = 0.70710678118654752f;
    float thr[LEVELS];
    {   // lane k computes the threshold of level k, then it is broadcast
        const int k = lane < LEVELS ? lane : LEVELS - 1;
        const float dvar = st.level_ss / (float)(L >> (k + 1));
        const float tk = var / sqrtf(fmaxf(dvar - var, 2.220446049250313e-16f));
#pragma unroll
        for (int kk = 0; kk < LEVELS; ++kk) thr[kk] = __shfl_sync(0xffffffffu, tk, kk);
    }
    if (threadIdx.x < 32) trace(320);
#pragma unroll
    for (int q = 1; q >= 0; --q) {
        const float d = soft_shrink(st.Dx[q], thr[2 + LG + q]);
        const bool ev = (lane & (1 << q)) == 0;
        st.A[q] = (ev ? (st.A[q + 1] + d) : (st.A[q + 1] - d)) * RS2;
    }
    st.v[0] = st.A[0];
#pragma unroll
    for (int lv = LG; lv >= 1; --lv) {
        const int stride = 1 << lv, half = stride >> 1;
#pragma unroll
        for (int i = 0; i < NCH / stride; ++i) {
            const float a = st.v[i * stride];
            const float d = soft_shrink(st.v[i * stride + half], thr[1 + lv]);
            st.v[i * stride] = (a + d) * RS2;
            st.v[i * stride + half] = (a - d) * RS2;
        }
    }
    {   // back to the chunk-cyclic layout through the warp's scratch
        float4* b4 = reinterpret_cast<float4*>(tscr + haar_wl_pad(lane * NCH));
#pragma unroll
        for (int j = 0; j < NCH / 4; ++j) b4[j] = make_float4(st.v[4 * j], st.v[4 * j + 1], st.v[4 * j + 2], st.v[4 * j + 3]);
    }
    __syncwarp();
    if (threadIdx.x < 32) trace(321);
    float err = 0.f;
#ifdef PNP_TRACE
    for (int rep = 0; rep < (PNP_DBG(4) ? 2 : 1); ++rep) {       // experiment: is the second pass over the same code faster?
    if (rep && threadIdx.x < 32) trace(323);
#endif
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
        const float aa = tscr[haar_wl_pad(c * 32 + lane)];
        const float dd = soft_shrink(st.x[c][2], thr[1]);
        const float a0 = (aa + dd) * RS2, a1 = (aa - dd) * RS2;
        const float d0 = soft_shrink(st.x[c][1], thr[0]), d1 = soft_shrink(st.x[c][3], thr[0]);
        const float4 o = make_float4((a0 + d0) * RS2, (a0 - d0) * RS2, (a1 + d1) * RS2, (a1 - d1) * RS2);
        if (!PNP_DBG(1)) stg_keep(zo4 + c * 32 + lane, o);           // the new iterate: the next forward line pass reads it
        if (xr4 && !PNP_DBG(2)) {
            const float4 r = xr4[c * 32 + lane];
            const float e0 = o.x - r.x, e1 = o.y - r.y, e2 = o.z - r.z, e3 = o.w - r.w;
            err = fmaf(e0, e0, fmaf(e1, e1, fmaf(e2, e2, fmaf(e3, e3, err))));
        }
        if (keep4) keep4[c * 32 + lane] = o;
    }
#ifdef PNP_TRACE
    }
#endif
    __syncwarp();                      // the scratch may be rewritten by the caller's next line
    if (threadIdx.x < 32) trace(322);
    return err;
}

template <int L>
__device__ __noinline__ void prox_phase_shrink(float* lines, int mine, long long first, int nlines, int batch,
                                                  float* __restrict__ zout, const float* __restrict__ xrec,
                                                  float sigma_modifier, float fallback_sigma, const double* sig_log,
                                                  double* __restrict__ mse_log, int cur_slot) {
    using C = HaarSub<L>;
    constexpr int VPL = C::VPL, LEVELS = C::LEVELS, XL = C::XL, LIN = C::LIN, SB = C::SB, NSB = C::NSB;
    constexpr float RS2 = 0.70710678118654752f;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float err_acc = 0.f;
    int err_img = -1;
    auto flush_err = [&](int img) {
        if (err_img != img) {
            if (err_img >= 0 && xrec && mse_log) {
                const float e = warp_sum_f(err_acc);
                if (lane == 0) atomicAdd(mse_log + (long long)cur_slot * batch + err_img, (double)e);
            }
            err_acc = 0.f;
            err_img = img;
        }
    };
    auto sigma_var = [&](int img) {
        const double se = __ldcg(sig_log + (long long)cur_slot * batch + img) / (double)nlines;
        const float sigma = (se > 0.0) ? (float)(se * (double)sigma_modifier) : fallback_sigma;
        return sigma * sigma;
    };
    if constexpr (L >= 512 && L <= 2048) {
        using W = HaarWL<L>;
        constexpr int NCH = W::NCH, LG = W::LG;
        static_assert(LEVELS == LG + 4, "level split of the one-warp-per-line shrink");
        for (int l = warp; l < mine; l += 16) {
            const long long gl = first + l;
            const int img = (int)(gl / nlines);
            float* sl = lines + (long long)l * L;
            const float4* s4 = reinterpret_cast<const float4*>(sl);
            const float4* xr4 = xrec ? reinterpret_cast<const float4*>(xrec + gl * L) : nullptr;
            float4* zo4 = reinterpret_cast<float4*>(zout + gl * L);
            float x[NCH][4], ss[LEVELS], thr[LEVELS];
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                const float4 q = s4[c * 32 + lane];
                x[c][0] = q.x; x[c][1] = q.y; x[c][2] = q.z; x[c][3] = q.w;
            }
#pragma unroll
            for (int k = 0; k < LEVELS; ++k) ss[k] = 0.f;
            __syncwarp();                      // every lane holds its samples: the line's shared memory is scratch now
#pragma unroll
            for (int c = 0; c < NCH; ++c) {    // levels 1, 2; details stay where the inverse expects them
                const float a0 = (x[c][0] + x[c][1]) * RS2, d0 = (x[c][0] - x[c][1]) * RS2;
                const float a1 = (x[c][2] + x[c][3]) * RS2, d1 = (x[c][2] - x[c][3]) * RS2;
                ss[0] = fmaf(d0, d0, fmaf(d1, d1, ss[0]));
                const float dd = (a0 - a1) * RS2;
                ss[1] = fmaf(dd, dd, ss[1]);
                x[c][1] = d0; x[c][3] = d1; x[c][2] = dd;
                sl[haar_wl_pad(c * 32 + lane)] = (a0 + a1) * RS2;
            }
            __syncwarp();
            float v[NCH];                      // approximations lane * NCH .. lane * NCH + NCH - 1
            {
                const float4* b4 = reinterpret_cast<const float4*>(sl + haar_wl_pad(lane * NCH));
#pragma unroll
                for (int j = 0; j < NCH / 4; ++j) {
                    const float4 q = b4[j];
                    v[4 * j] = q.x; v[4 * j + 1] = q.y; v[4 * j + 2] = q.z; v[4 * j + 3] = q.w;
                }
            }
            float4 xr[2];                      // ground truth of the first two chunks: in flight during the pyramid top
            if (xrec) {
#pragma unroll
                for (int j = 0; j < 2; ++j) xr[j] = xr4[j * 32 + lane];
            }
#pragma unroll
            for (int lv = 1; lv <= LG; ++lv) {
                const int stride = 1 << lv, half = stride >> 1;
#pragma unroll
                for (int i = 0; i < NCH / stride; ++i) {
                    const float p = v[i * stride], q = v[i * stride + half];
                    const float d = (p - q) * RS2;
                    v[i * stride] = (p + q) * RS2;
                    v[i * stride + half] = d;
                    ss[1 + lv] = fmaf(d, d, ss[1 + lv]);
                }
            }
            float A[3], Dx[2];
            A[0] = v[0];
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const float p = __shfl_xor_sync(0xffffffffu, A[q], 1 << q);
                const bool ev = (lane & (1 << q)) == 0;
                A[q + 1] = (A[q] + p) * RS2;
                Dx[q] = (ev ? (A[q] - p) : (p - A[q])) * RS2;
                if ((lane & ((2 << q) - 1)) == 0) ss[2 + LG + q] = Dx[q] * Dx[q];
            }
            const float var = sigma_var(img);
            {   // lane k computes the threshold of level k, then it is broadcast
                float mine_ss = 0.f;
#pragma unroll
                for (int k = 0; k < LEVELS; ++k) {
                    const float e = warp_sum_f(ss[k]);
                    mine_ss = lane == k ? e : mine_ss;
                }
                const int k = lane < LEVELS ? lane : LEVELS - 1;
                const float dvar = mine_ss / (float)(L >> (k + 1));
                const float tk = var / sqrtf(fmaxf(dvar - var, 2.220446049250313e-16f));
#pragma unroll
                for (int kk = 0; kk < LEVELS; ++kk) thr[kk] = __shfl_sync(0xffffffffu, tk, kk);
            }
#pragma unroll
            for (int q = 1; q >= 0; --q) {
                const float d = soft_shrink(Dx[q], thr[2 + LG + q]);
                const bool ev = (lane & (1 << q)) == 0;
                A[q] = (ev ? (A[q + 1] + d) : (A[q + 1] - d)) * RS2;
            }
            v[0] = A[0];
#pragma unroll
            for (int lv = LG; lv >= 1; --lv) {
                const int stride = 1 << lv, half = stride >> 1;
#pragma unroll
                for (int i = 0; i < NCH / stride; ++i) {
                    const float a = v[i * stride];
                    const float d = soft_shrink(v[i * stride + half], thr[1 + lv]);
                    v[i * stride] = (a + d) * RS2;
                    v[i * stride + half] = (a - d) * RS2;
                }
            }
            {   // back to the chunk-cyclic layout (a lane rewrites the block only it has read)
                float4* b4 = reinterpret_cast<float4*>(sl + haar_wl_pad(lane * NCH));
#pragma unroll
                for (int j = 0; j < NCH / 4; ++j) b4[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            }
            __syncwarp();
            flush_err(img);
#pragma unroll
            for (int g = 0; g < NCH / 2; ++g) {
                float4 nx[2];
                if (xrec && g + 1 < NCH / 2) {
#pragma unroll
                    for (int j = 0; j < 2; ++j) nx[j] = xr4[((g + 1) * 2 + j) * 32 + lane];
                }
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const int c = g * 2 + j;
                    const float aa = sl[haar_wl_pad(c * 32 + lane)];
                    const float dd = soft_shrink(x[c][2], thr[1]);
                    const float a0 = (aa + dd) * RS2, a1 = (aa - dd) * RS2;
                    const float d0 = soft_shrink(x[c][1], thr[0]), d1 = soft_shrink(x[c][3], thr[0]);
                    const float4 o = make_float4((a0 + d0) * RS2, (a0 - d0) * RS2, (a1 + d1) * RS2, (a1 - d1) * RS2);
                    zo4[c * 32 + lane] = o;
                    if (xrec) {
                        const float e0 = o.x - xr[j].x, e1 = o.y - xr[j].y, e2 = o.z - xr[j].z, e3 = o.w - xr[j].w;
                        err_acc = fmaf(e0, e0, fmaf(e1, e1, fmaf(e2, e2, fmaf(e3, e3, err_acc))));
                    }
                }
                if (g + 1 < NCH / 2) {
#pragma unroll
                    for (int j = 0; j < 2; ++j) xr[j] = nx[j];
                }
            }
        }
    } else if constexpr (SB == 512) {
        __shared__ float s_ss[PROX_MAX_TASKS][LEVELS];
        const int tasks = mine * NSB;
        for (int t = warp; t < tasks; t += 16) {                                   // pass A: energies
            const float4* s4 = reinterpret_cast<const float4*>(lines + (long long)t * SB);
            float x[4][4], A[4][6], D[4][5], T[4], ss[LEVELS];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const float4 q = s4[c * 32 + lane];
                x[c][0] = q.x; x[c][1] = q.y; x[c][2] = q.z; x[c][3] = q.w;
            }
#pragma unroll
            for (int k = 0; k < LEVELS; ++k) ss[k] = 0.f;
            haar_cc_forward<LEVELS, 4>(x, A, D, T, ss, lane);
#pragma unroll
            for (int k = 0; k < LEVELS; ++k) {
                const float e = warp_sum_f(ss[k]);
                if (lane == 0) s_ss[t][k] = e;
            }
        }
        __syncthreads();
        for (int t = warp; t < tasks; t += 16) {                                   // pass B: shrink
            const int l = t / NSB;
            const long long gl = first + l;
            const int img = (int)(gl / nlines);
            const long long gbase = first * L + (long long)t * SB;                 // sub-blocks are contiguous in memory
            const float4* s4 = reinterpret_cast<const float4*>(lines + (long long)t * SB);
            float x[4][4], xr[4][4], A[4][6], D[4][5], T[4], dummy[LEVELS], thr[LEVELS];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                if (xrec) {                                                        // latency hides behind the transform
                    const float4 r = reinterpret_cast<const float4*>(xrec + gbase)[c * 32 + lane];
                    xr[c][0] = r.x; xr[c][1] = r.y; xr[c][2] = r.z; xr[c][3] = r.w;
                }
                const float4 q = s4[c * 32 + lane];
                x[c][0] = q.x; x[c][1] = q.y; x[c][2] = q.z; x[c][3] = q.w;
            }
            const float var = sigma_var(img);
            {   // lane k computes the threshold of level k, then it is broadcast (not 32 lanes x 8 divisions and roots)
                const int k = lane < LEVELS ? lane : LEVELS - 1;
                float e = 0.f;
#pragma unroll
                for (int w = 0; w < NSB; ++w) e += s_ss[l * NSB + w][k];
                const float dvar = e / (float)(L >> (k + 1));
                const float tk = var / sqrtf(fmaxf(dvar - var, 2.220446049250313e-16f));
#pragma unroll
                for (int kk = 0; kk < LEVELS; ++kk) { thr[kk] = __shfl_sync(0xffffffffu, tk, kk); dummy[kk] = 0.f; }
            }
            haar_cc_forward<LEVELS, 4>(x, A, D, T, dummy, lane);
            haar_cc_inverse<LEVELS, 4>(x, A, D, T, thr, lane);
            flush_err(img);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                reinterpret_cast<float4*>(zout + gbase)[c * 32 + lane] = make_float4(x[c][0], x[c][1], x[c][2], x[c][3]);
                if (xrec) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) { const float e = x[c][j] - xr[c][j]; err_acc = fmaf(e, e, err_acc); }
                }
            }
        }
    } else if constexpr (L >= 128) {
        // 128 or 256 samples: the line is one task of one warp, same chunk-cyclic layout with 1 or 2 chunks per lane
        constexpr int NCH = L / 128;
        for (int l = warp; l < mine; l += 16) {
            const long long gl = first + l;
            const int img = (int)(gl / nlines);
            const float4* s4 = reinterpret_cast<const float4*>(lines + (long long)l * L);
            float x[NCH][4], xr[NCH][4], A[NCH][6], D[NCH][5], T[4], ss[LEVELS], thr[LEVELS];
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                if (xrec) {
                    const float4 r = reinterpret_cast<const float4*>(xrec + gl * L)[c * 32 + lane];
                    xr[c][0] = r.x; xr[c][1] = r.y; xr[c][2] = r.z; xr[c][3] = r.w;
                }
                const float4 q = s4[c * 32 + lane];
                x[c][0] = q.x; x[c][1] = q.y; x[c][2] = q.z; x[c][3] = q.w;
            }
#pragma unroll
            for (int k = 0; k < LEVELS; ++k) ss[k] = 0.f;
            haar_cc_forward<LEVELS, NCH>(x, A, D, T, ss, lane);
            const float var = sigma_var(img);
            {
                float mine_ss = 0.f;
#pragma unroll
                for (int k = 0; k < LEVELS; ++k) {
                    const float e = warp_sum_f(ss[k]);
                    mine_ss = lane == k ? e : mine_ss;
                }
                const int k = lane < LEVELS ? lane : LEVELS - 1;
                const float dvar = mine_ss / (float)(L >> (k + 1));
                const float tk = var / sqrtf(fmaxf(dvar - var, 2.220446049250313e-16f));
#pragma unroll
                for (int kk = 0; kk < LEVELS; ++kk) thr[kk] = __shfl_sync(0xffffffffu, tk, kk);
            }
            haar_cc_inverse<LEVELS, NCH>(x, A, D, T, thr, lane);
            flush_err(img);
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                reinterpret_cast<float4*>(zout + gl * L)[c * 32 + lane] = make_float4(x[c][0], x[c][1], x[c][2], x[c][3]);
                if (xrec) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) { const float e = x[c][j] - xr[c][j]; err_acc = fmaf(e, e, err_acc); }
                }
            }
        }
    } else {
        for (int l = warp; l < mine; l += 16) {
            const long long gl = first + l;
            const int img = (int)(gl / nlines);
            const long long goff = gl * L + lane * VPL;
            float x[VPL], xr[VPL], A[XL + 1], Dx[XL], ss[LEVELS], thr[LEVELS];
            const float* s0 = lines + (long long)l * L + lane * VPL;
#pragma unroll
            for (int i = 0; i < VPL; ++i) {
                x[i] = s0[i];
                if (xrec) xr[i] = xrec[goff + i];
            }
#pragma unroll
            for (int k = 0; k < LEVELS; ++k) ss[k] = 0.f;
            haar_sub_forward<L>(x, A, Dx, ss, lane);
            const float var = sigma_var(img);
#pragma unroll
            for (int k = 0; k < LEVELS; ++k) {
                const float dvar = warp_sum_f(ss[k]) / (float)(L >> (k + 1));
                thr[k] = var / sqrtf(fmaxf(dvar - var, 2.220446049250313e-16f));
            }
#pragma unroll
            for (int q = XL - 1; q >= 0; --q) {
                const float d = soft_shrink(Dx[q], thr[LIN + q]);
                const bool ev = (lane & (1 << q)) == 0;
                A[q] = (ev ? (A[q + 1] + d) : (A[q + 1] - d)) * RS2;
            }
            x[0] = A[0];
#pragma unroll
            for (int lv = LIN; lv >= 1; --lv) {
                const int stride = 1 << lv, half = stride >> 1;
#pragma unroll
                for (int i = 0; i < VPL / stride; ++i) {
                    const float a = x[i * stride];
                    const float d = soft_shrink(x[i * stride + half], thr[lv - 1]);
                    x[i * stride] = (a + d) * RS2;
                    x[i * stride + half] = (a - d) * RS2;
                }
            }
            flush_err(img);
#pragma unroll
            for (int i = 0; i < VPL; ++i) {
                zout[goff + i] = x[i];
                if (xrec) { const float e = x[i] - xr[i]; err_acc = fmaf(e, e, err_acc); }
            }
        }
    }
    if (err_img >= 0 && xrec && mse_log) {
        const float e = warp_sum_f(err_acc);
        if (lane == 0) atomicAdd(mse_log + (long long)cur_slot * batch + err_img, (double)e);
    }
}

// Both fused kernels (k_prox_wavelet_fused, k_update_prox) end with this on lines resident in shared memory:
//   sigma estimate per line -> [grid barrier: mean over all lines] -> BayesShrink -> store + squared error.
// 512 .. 2048 samples with at most one line per warp: the warp also runs the FORWARD Haar pyramid before the barrier
// and keeps it in registers (it does not depend on sigma), so the lines' shared memory is free during the barrier:
// the ground-truth lines are pulled into it by TMA bulk copies meanwhile (one warp streaming them with plain loads
// had ~1 KiB in flight and was bound by DRAM latency), and after the barrier only thresholds + inverse remain.
// `scratch`: 16 x prox_scratch<L>() words; `xbar`: an initialised mbarrier (count 1, phase 0) not used otherwise.
template <int L>
__device__ __forceinline__ void prox_phases(float* lines, int mine, long long first, int nlines, int batch, float* __restrict__ zout,
                                            const float* __restrict__ xrec, float sigma_modifier, float fallback_sigma,
                                            double* __restrict__ sig_log, double* __restrict__ mse_log, int cur_slot,
                                            unsigned* scratch, unsigned long long* xbar, int* __restrict__ advance, int n_advance,
                                            unsigned* __restrict__ gbar /* software grid barrier workspace, or null: cooperative launch */,
                                            bool keep_lines = false /* leave the new lines in `lines` (one line per warp only) */) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    auto grid_sync = [&]() {
        if (gbar) {
            sw_grid_sync(gbar, gridDim.x);
        } else {
            __threadfence();
            cooperative_groups::this_grid().sync();
        }
    };
    if constexpr (L >= 512 && L <= 2048) {
        if (mine <= 16) {
            ProxLine<L> st;
            const bool active = warp < mine;
            const long long gl = first + warp;
            const int img = active ? (int)(gl / nlines) : 0;
            float* sl = lines + (long long)warp * L;
            if (active) {
                const double sig = line_sigma_mad<L>(sl, lane, scratch + warp * prox_scratch<L>());
                if (lane == 0) atomicAdd(sig_log + (long long)cur_slot * batch + img, sig);
                if (warp == 0) trace(302);
                prox_line_forward<L>(sl, lane, st);
            }
            __syncthreads();
            trace(305);
            if (xrec && mine > 0 && threadIdx.x == 0) {
                fence_proxy_async_smem();              // the lines were read / written through the generic proxy
                const unsigned total = (unsigned)(mine * L * sizeof(float));
                mbar_expect_tx(xbar, total);
                for (unsigned off = 0; off < total; off += 32768u) {
                    const unsigned n = total - off < 32768u ? total - off : 32768u;
                    bulk_g2s(reinterpret_cast<char*>(lines) + off, reinterpret_cast<const char*>(xrec + first * L) + off, n, xbar);
                }
            }
            grid_sync();
            trace(303);
            if (advance && blockIdx.x == 0 && threadIdx.x < n_advance) advance[threadIdx.x] += 1;
            if (active) {
                const double se = __ldcg(sig_log + (long long)cur_slot * batch + img) / (double)nlines;
                const float sigma = (se > 0.0) ? (float)(se * (double)sigma_modifier) : fallback_sigma;
                if (xrec) mbar_wait(xbar, 0);
                if (warp == 0) trace(306);
                float err = prox_line_inverse<L>(st, reinterpret_cast<float*>(scratch + warp * prox_scratch<L>()), lane, sigma * sigma,
                                                 reinterpret_cast<float4*>(zout + gl * L),
                                                 xrec ? reinterpret_cast<const float4*>(sl) : nullptr,
                                                 keep_lines ? reinterpret_cast<float4*>(sl) : nullptr);
                if (xrec && mse_log) {
                    err = warp_sum_f(err);
                    if (lane == 0) atomicAdd(mse_log + (long long)cur_slot * batch + img, (double)err);
                }
            }
            return;
        }
    }
    prox_phase_sigma<L>(lines, mine, first, nlines, batch, sig_log, cur_slot, scratch);
    __syncthreads();
    grid_sync();
    // end-of-iteration counters (pnp_advance) folded in: every CTA has read *slot before the barrier above, and
    // nothing else of this iteration reads them any more
    if (advance && blockIdx.x == 0 && threadIdx.x < n_advance) advance[threadIdx.x] += 1;
    prox_phase_shrink<L>(lines, mine, first, nlines, batch, zout, xrec, sigma_modifier, fallback_sigma, sig_log, mse_log, cur_slot);
}

template <int L>
__global__ void __launch_bounds__(512, 1)
k_prox_wavelet_fused(const float* __restrict__ zin, float* __restrict__ zout, const float* __restrict__ xrec,
                     int nlines, int batch, int lines_per_cta, float sigma_modifier, float fallback_sigma,
                     double* __restrict__ sig_log, double* __restrict__ mse_log, const int* __restrict__ slot,
                     unsigned* __restrict__ gbar) {
    using C = HaarSub<L>;
    constexpr int VPL = C::VPL, LEVELS = C::LEVELS, XL = C::XL, LIN = C::LIN, SB = C::SB, NSB = C::NSB;
    constexpr float RS2 = 0.70710678118654752f;
    extern __shared__ __align__(128) float lines[];                   // lines_per_cta x L
    __shared__ __align__(8) unsigned long long bar, xbar;
    __shared__ __align__(16) unsigned scratch[16 * prox_scratch<L>()];
    const long long total = (long long)nlines * batch;
    const long long first = (long long)blockIdx.x * lines_per_cta;
    long long mine = total - first;
    mine = mine < 0 ? 0 : (mine > lines_per_cta ? lines_per_cta : mine);
    griddep_wait();
    griddep_launch();
    const int cur_slot = slot ? *slot : 0;

    // ---- stage my lines (contiguous in memory) ----
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        mbar_init(&xbar, 1);
        mbar_fence_init();
        if (mine > 0) {
            // the CTA's lines are one contiguous block: a few large bulk copies instead of one per (short) line
            const unsigned total = (unsigned)(mine * L * sizeof(float));
            mbar_expect_tx(&bar, total);
            for (unsigned off = 0; off < total; off += 32768u) {
                const unsigned n = total - off < 32768u ? total - off : 32768u;
                bulk_g2s(reinterpret_cast<char*>(lines) + off, reinterpret_cast<const char*>(zin + first * L) + off, n, &bar);
            }
        }
    }
    __syncthreads();
    if (mine > 0) mbar_wait(&bar, 0);

    // ---- sigma estimate, grid-wide mean, BayesShrink of the resident lines ----
    prox_phases<L>(lines, (int)mine, first, nlines, batch, zout, xrec, sigma_modifier, fallback_sigma, sig_log, mse_log, cur_slot,
                   scratch, &xbar, nullptr, 0, gbar);
}

}  // namespace pnp
