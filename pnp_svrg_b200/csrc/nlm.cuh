// Non-local means (classic, fast_mode=False) for sm_100a.
//
// Reference: denoisers/NLM.py:22-27 -> skimage.restoration.denoise_nl_means(h, sigma,
// fast_mode=False, patch_size=4 -> 5, patch_distance=5) on a 2-D grey image, restated in
// oracle/skimage_port.py::denoise_nl_means: reflect padding by s/2, Gaussian patch weights of
// width (s-1)/4 normalised by (sum w) h^2, every squared difference reduced by 2 sigma^2, the
// running distance tested against 5.0 at the START of each patch row (early exit -> weight 0),
// weight exp(-max(0, distance)), search window clipped to the image.
//
// One thread per pixel, 16x16 pixel tile per CTA, the tile plus a (d + s/2)-pixel halo staged in
// shared memory.  The image is stored transposed (line layout); the kernel indexes it as
// img(r, c) = z[c*H + r] so that patch rows are the reference's rows (the early exit is per row).
#pragma once
#include <cuda_runtime.h>

namespace pnp {

#define NLM_TILE 16
#define NLM_MAX_S 7
#define NLM_MAX_HALO 16

struct NlmParams {
    int s;               // odd patch size
    int d;               // patch distance
    const double* sig_log;
    float sigma_est;     // used when sig_log == nullptr
    float sigma_modifier;
    float fallback_h;    // denoise_strength * decay**t when sigma_est <= 0 (then var = 0)
};

__device__ __forceinline__ int reflect_idx(int i, int n) {      // numpy pad mode 'reflect'
    if (i < 0) i = -i;
    if (i >= n) i = 2 * (n - 1) - i;
    return i;
}

__global__ void __launch_bounds__(NLM_TILE * NLM_TILE)
k_nlm(const float* __restrict__ zin, float* __restrict__ zout, const float* __restrict__ xrec, int H, int W,
      long long img_stride, NlmParams np_, double* __restrict__ mse_log, const int* __restrict__ slot, int batch) {
    extern __shared__ float tile[];
    __shared__ float wgt[NLM_MAX_S * NLM_MAX_S];
    __shared__ float s_err[NLM_TILE * NLM_TILE / 32];
    const int s = np_.s, d = np_.d, off = s / 2;
    const int halo = d + off;
    const int TW = NLM_TILE + 2 * halo;
    const int img = blockIdx.z;
    const float* zi = zin + (long long)img * img_stride;
    const int r0 = blockIdx.y * NLM_TILE, c0 = blockIdx.x * NLM_TILE;
    const int tid = threadIdx.y * NLM_TILE + threadIdx.x;

    double se = np_.sig_log ? *slot_ptr(const_cast<double*>(np_.sig_log), slot, batch, img) / (double)W
                            : (double)np_.sigma_est;
    float h, var;
    if (se > 0.0) {
        const float sg = (float)(se * (double)np_.sigma_modifier);
        h = sg;
        var = 2.f * sg * sg;
    } else {
        h = np_.fallback_h;
        var = 0.f;
    }
    // tile(rr, cc) = padded image at row r0 - halo + rr, col c0 - halo + cc  (reflect outside)
    for (int i = tid; i < TW * TW; i += NLM_TILE * NLM_TILE) {
        const int rr = i / TW, cc = i - rr * TW;
        const int r = reflect_idx(r0 - halo + rr, H), c = reflect_idx(c0 - halo + cc, W);
        tile[i] = zi[(long long)c * H + r];
    }
    if (tid < s * s) {
        const float A = (float)(s - 1) * 0.25f;
        float sum = 0.f;
        for (int q = 0; q < s * s; ++q) {
            const float dy = (float)(q / s - off), dx = (float)(q % s - off);
            sum += expf(-(dy * dy + dx * dx) / (2.f * A * A));
        }
        const float dy = (float)(tid / s - off), dx = (float)(tid % s - off);
        wgt[tid] = expf(-(dy * dy + dx * dx) / (2.f * A * A)) / (sum * h * h);
    }
    __syncthreads();
    const int r = r0 + threadIdx.y, c = c0 + threadIdx.x;
    float err = 0.f;
    if (r < H && c < W) {
        const int pr = threadIdx.y + halo, pc = threadIdx.x + halo;        // centre in tile coordinates
        float acc = 0.f, wsum = 0.f;
        const int i_lo = -min(d, r), i_hi = min(d + 1, H - r);             // clipped search window
        const int j_lo = -min(d, c), j_hi = min(d + 1, W - c);
        for (int di = i_lo; di < i_hi; ++di) {
            for (int dj = j_lo; dj < j_hi; ++dj) {
                float dist = 0.f;
                bool dead = false;
                for (int pi = 0; pi < s; ++pi) {
                    if (dist > 5.0f) { dead = true; break; }
                    const float* p1 = tile + (pr - off + pi) * TW + (pc - off);
                    const float* p2 = tile + (pr - off + pi + di) * TW + (pc - off + dj);
                    for (int pj = 0; pj < s; ++pj) {
                        const float df = p1[pj] - p2[pj];
                        dist = fmaf(wgt[pi * s + pj], fmaf(df, df, -var), dist);
                    }
                }
                if (!dead) {
                    const float wv = expf(-fmaxf(0.f, dist));
                    wsum += wv;
                    acc = fmaf(wv, tile[(pr + di) * TW + (pc + dj)], acc);
                }
            }
        }
        const float o = acc / wsum;
        const long long e = (long long)img * img_stride + (long long)c * H + r;
        zout[e] = o;
        if (xrec) { const float df = o - xrec[e]; err = df * df; }
    }
    if (xrec && mse_log) {
        err = warp_sum_f(err);
        if ((tid & 31) == 0) s_err[tid >> 5] = err;
        __syncthreads();
        if (tid == 0) {
            float t = 0.f;
            for (int k = 0; k < NLM_TILE * NLM_TILE / 32; ++k) t += s_err[k];
            atomicAdd(slot_ptr(mse_log, slot, batch, img), (double)t);
        }
    }
}

// ---- the reference's configuration (patch 4 -> 5, any patch distance <= 13): specialised -------------------------------------
// The generic kernel above spends ~16 instructions per (candidate, tap) pair -- two shared loads per tap for the two
// patches, a third for the weight, loop and address arithmetic of run-time s -- at 14 warps per SM for a 256 x 256 image
// (ncu: 102 M warp instructions, 169 us).  Here:
//   * the pixel's OWN patch (25 values) is read once into registers, the six distinct Gaussian weights (they depend on
//     dy^2 + dx^2 only) live in registers, the 5 x 5 loops are unrolled: 25 loads + 75 FP instructions per candidate;
//   * FOUR threads share a pixel: the candidate rows di are dealt round-robin over blockDim.z = 4, the partial sums
//     (sum w v, sum w) are added in a fixed order through shared memory: 4x the warps for the same image;
//   * tile 16 x 8 pixels, shared row stride 48 floats: the two tile rows of a warp sit 16 banks apart (no conflicts).
// Per candidate the arithmetic and its order are the generic kernel's (row by row, left to right, the same fmaf chain,
// the same early-exit test at the start of every patch row), so distances and exit decisions are bit-identical; only
// the order in which the 121 candidate terms are added differs (four partial sums).
#define NLM5_TX 16
#define NLM5_TY 8
#define NLM5_G 4
__device__ __forceinline__ constexpr int nlm5_cls(int pi, int pj) {       // class of dy^2 + dx^2 in {0, 1, 2, 4, 5, 8}
    const int q = (pi - 2) * (pi - 2) + (pj - 2) * (pj - 2);
    return q == 0 ? 0 : q == 1 ? 1 : q == 2 ? 2 : q == 4 ? 3 : q == 5 ? 4 : 5;
}

__global__ void __launch_bounds__(NLM5_TX * NLM5_TY * NLM5_G)
k_nlm5(const float* __restrict__ zin, float* __restrict__ zout, const float* __restrict__ xrec, int H, int W,
       long long img_stride, NlmParams np_, double* __restrict__ mse_log, const int* __restrict__ slot, int batch) {
    constexpr int S = 5, OFF = 2;
    extern __shared__ float tile[];                       // [TH][TWP]
    __shared__ float wgt[S * S];
    __shared__ float s_acc[NLM5_G][NLM5_TY][NLM5_TX], s_ws[NLM5_G][NLM5_TY][NLM5_TX];
    __shared__ float s_err[NLM5_TX * NLM5_TY / 32];
    const int d = np_.d, halo = d + OFF;
    const int TWd = NLM5_TX + 2 * halo, TH = NLM5_TY + 2 * halo;
    const int TWP = ((TWd + 31) / 32) * 32 + 16;          // == 16 (mod 32)
    const int img = blockIdx.z;
    const float* zi = zin + (long long)img * img_stride;
    const int r0 = blockIdx.y * NLM5_TY, c0 = blockIdx.x * NLM5_TX;
    const int tid = (threadIdx.z * NLM5_TY + threadIdx.y) * NLM5_TX + threadIdx.x;
    constexpr int NTH = NLM5_TX * NLM5_TY * NLM5_G;

    double se = np_.sig_log ? *slot_ptr(const_cast<double*>(np_.sig_log), slot, batch, img) / (double)W
                            : (double)np_.sigma_est;
    float h, var;
    if (se > 0.0) {
        const float sg = (float)(se * (double)np_.sigma_modifier);
        h = sg;
        var = 2.f * sg * sg;
    } else {
        h = np_.fallback_h;
        var = 0.f;
    }
    for (int i = tid; i < TH * TWd; i += NTH) {
        const int rr = i / TWd, cc = i - rr * TWd;
        const int r = reflect_idx(r0 - halo + rr, H), c = reflect_idx(c0 - halo + cc, W);
        tile[rr * TWP + cc] = zi[(long long)c * H + r];
    }
    if (tid < S * S) {
        const float A = (float)(S - 1) * 0.25f;
        float sum = 0.f;
        for (int q = 0; q < S * S; ++q) {
            const float dy = (float)(q / S - OFF), dx = (float)(q % S - OFF);
            sum += expf(-(dy * dy + dx * dx) / (2.f * A * A));
        }
        const float dy = (float)(tid / S - OFF), dx = (float)(tid % S - OFF);
        wgt[tid] = expf(-(dy * dy + dx * dx) / (2.f * A * A)) / (sum * h * h);
    }
    __syncthreads();
    const int r = r0 + threadIdx.y, c = c0 + threadIdx.x;
    const bool inside = r < H && c < W;
    float acc = 0.f, wsum = 0.f;
    if (inside) {
        // the six distinct weights: entries (2,2) (2,3) (3,3) (2,4) (3,4) (4,4) of the 5 x 5 table
        const float wc[6] = {wgt[12], wgt[13], wgt[18], wgt[14], wgt[19], wgt[24]};
        const int pr = threadIdx.y + halo, pc = threadIdx.x + halo;
        float p1[S][S];
#pragma unroll
        for (int pi = 0; pi < S; ++pi)
#pragma unroll
            for (int pj = 0; pj < S; ++pj) p1[pi][pj] = tile[(pr - OFF + pi) * TWP + (pc - OFF + pj)];
        const int i_lo = -min(d, r), i_hi = min(d + 1, H - r);
        const int j_lo = -min(d, c), j_hi = min(d + 1, W - c);
        for (int di = i_lo + (int)threadIdx.z; di < i_hi; di += NLM5_G) {
            const float* base = tile + (pr - OFF + di) * TWP + (pc - OFF);
            for (int dj = j_lo; dj < j_hi; ++dj) {
                const float* p2 = base + dj;
                float dist = 0.f;
                bool dead = false;
#pragma unroll
                for (int pi = 0; pi < S; ++pi) {
                    if (!dead) {
                        if (dist > 5.0f) {
                            dead = true;
                        } else {
#pragma unroll
                            for (int pj = 0; pj < S; ++pj) {
                                const float df = p1[pi][pj] - p2[pi * TWP + pj];
                                dist = fmaf(wc[nlm5_cls(pi, pj)], fmaf(df, df, -var), dist);
                            }
                        }
                    }
                }
                if (!dead) {
                    const float wv = expf(-fmaxf(0.f, dist));
                    wsum += wv;
                    acc = fmaf(wv, p2[OFF * TWP + OFF], acc);
                }
            }
        }
    }
    s_acc[threadIdx.z][threadIdx.y][threadIdx.x] = acc;
    s_ws[threadIdx.z][threadIdx.y][threadIdx.x] = wsum;
    __syncthreads();
    float err = 0.f;
    if (threadIdx.z == 0 && inside) {
        float a = 0.f, w = 0.f;
#pragma unroll
        for (int g = 0; g < NLM5_G; ++g) { a += s_acc[g][threadIdx.y][threadIdx.x]; w += s_ws[g][threadIdx.y][threadIdx.x]; }
        const float o = a / w;
        const long long e = (long long)img * img_stride + (long long)c * H + r;
        zout[e] = o;
        if (xrec) { const float df = o - xrec[e]; err = df * df; }
    }
    if (xrec && mse_log && threadIdx.z == 0) {            // warps 0 .. TX*TY/32 - 1 (whole warps: z == 0 covers tid < 128)
        err = warp_sum_f(err);
        if ((tid & 31) == 0) s_err[tid >> 5] = err;
    }
    __syncthreads();
    if (xrec && mse_log && tid == 0) {
        float t = 0.f;
        for (int k = 0; k < NLM5_TX * NLM5_TY / 32; ++k) t += s_err[k];
        atomicAdd(slot_ptr(mse_log, slot, batch, img), (double)t);
    }
}

}  // namespace pnp
