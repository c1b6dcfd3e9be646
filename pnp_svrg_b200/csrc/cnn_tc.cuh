// 64 -> 64 channel 3x3 convolution on the 5th-generation tensor cores (tcgen05 / TMEM / TMA), bf16
// operands, fp32 accumulation -- the throughput path of the DnCNN / MMO conv stacks
// (reference: denoisers/DeepDenoisers/model/models.py:5-22, realSN_models.py:5-18,
// denoisers/MMODenoise.py:73-103; cuDNN via torch in the reference, no sm_100 code there).
//
// Formulation.  Activations are bf16 NHWC with ONE zero pixel appended to every image line
// (pitch PW+1), so the whole image is a flat sequence of positions s = line*(PW+1) + pixel in which a
// shift by dp = +-1 stays inside the line or lands on a zero, and a shift by dl = +-1 is a shift by
// +-(PW+1).  For a tile of 128 consecutive positions the three tap COLUMNS are computed unshifted,
//
//     T_dp[s][co] = sum_{dl, ci} in[s + dl*(PW+1)][ci] * w[dl][dp][ci][co]          (dp = -1, 0, +1)
//
// as ONE GEMM  [128 x 192(dl,ci)] x [192(dl,ci) x 192(dp,co)]  (12 tcgen05.mma M128 N192 K16, operands
// staged by TMA in 128B-swizzled K-major tiles, accumulator 128 lanes x 192 columns in TMEM), and the
// pixel shift is applied to the OUTPUT in the epilogue:
//
//     out[s] = T_-1[s-1] + T_0[s] + T_+1[s+1]
//
// (neighbour TMEM lanes via warp shuffles, warp boundaries through shared memory).  Row 0 and row 127
// of a tile are halo, so a tile yields 126 outputs (98.4 % efficiency) and the activation tile is
// read 3x (once per dl) instead of 9x.  A shifted A operand cannot be expressed by a UMMA shared
// memory descriptor (rows are grouped by 8 in the canonical layouts), hence the output-side shift.
//
// Warp roles (320 threads, one persistent CTA per SM): warp 0 TMA producer, warp 1 MMA issuer + TMEM
// allocator, warps 2-9 epilogue (TMEM lane quadrant = warp_id % 4, two warps per quadrant split the channels).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "fft_core.cuh"

namespace pnp {

#define TC_M 128
#define TC_N 192
#define TC_KBLK 64               // one dl: 64 input channels = 128 bytes per row (one swizzle span)
#define TC_A_BYTES (TC_M * 128)  // 16 KiB
#define TC_B_BYTES (TC_N * 128)  // 24 KiB per K block
#define TC_STAGES 4
#define TC_OUT_PER_TILE 126
#define TC_THREADS 320           // warp 0 TMA, warp 1 MMA, warps 2-9 epilogue (2 per TMEM lane quadrant)

struct TcSmem {
    unsigned long long full[TC_STAGES], empty[TC_STAGES], bfull, tfull[2], tempty[2];
    unsigned tmem_base;
    __align__(16) float halo[2][2][4][2][16];  // [channel half][chunk parity][quadrant][0: lane31 T_-1, 1: lane0 T_+1][16 channels]
    float2 affine[64];           // per output channel (scale, shift): folded BatchNorm / bias
};

__device__ __forceinline__ void mbar_wait_bounded(unsigned long long* bar, unsigned parity, int backoff = 0) {
    unsigned done = 0;
    for (unsigned spin = 0; !done; ++spin) {
        if (backoff && spin) __nanosleep(backoff);
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (spin > (1u << 26)) __trap();          // a protocol bug must fail loudly, not hang the GPU
    }
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, unsigned long long* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
// shared memory matrix descriptor, K-major, 128-byte swizzle: rows of 128 B, 8-row groups 1024 B apart
__device__ __forceinline__ unsigned long long umma_desc_sw128(const void* smem_ptr) {
    unsigned long long d = 0;
    d |= (unsigned long long)((smem_u32(smem_ptr) & 0x3FFFF) >> 4);       // start address        bits [0,14)
    d |= (unsigned long long)1 << 16;                                      // leading byte offset  bits [16,30) (unused for swizzled K-major)
    d |= (unsigned long long)(1024 >> 4) << 32;                            // stride byte offset   bits [32,46)
    d |= (unsigned long long)1 << 46;                                      // descriptor version 1 (Blackwell)
    d |= (unsigned long long)2 << 61;                                      // layout type SWIZZLE_128B
    return d;
}
// instruction descriptor kind::f16: D fp32, A/B bf16, both K-major, N = 192, M = 128
__device__ __forceinline__ constexpr unsigned umma_idesc_bf16(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((unsigned)(N >> 3) << 17) | ((unsigned)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_f16(unsigned tmem_d, unsigned long long da, unsigned long long db, unsigned idesc, unsigned accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(unsigned long long* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(unsigned taddr, float (&v)[16]) {
    unsigned r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// in / out: bf16 [S][64] with S = PH*(PW+1) positions; tmA over `in` (box 64 x 128), tmB over the packed
// weights [192 rows (dp,co)][192 (dl,ci)] (box 64 x 192).
__global__ void __launch_bounds__(TC_THREADS, 1)
k_conv_mid_tc(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
              __nv_bfloat16* __restrict__ out, const float* __restrict__ scale, const float* __restrict__ shift,
              float slope, int PW, int S, int n_tiles, int dbg) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<unsigned long long>(smem_raw) + 1023ull) & ~1023ull);
    unsigned char* sB = base;                                   // 3 x 24 KiB
    unsigned char* sA = base + 3 * TC_B_BYTES;                  // TC_STAGES x 16 KiB
    TcSmem* ctl = reinterpret_cast<TcSmem*>(sA + TC_STAGES * TC_A_BYTES);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int pitch = PW + 1;

    if (threadIdx.x == 0) {
        for (int i = 0; i < TC_STAGES; ++i) { mbar_init(&ctl->full[i], 1); mbar_init(&ctl->empty[i], 1); }
        mbar_init(&ctl->bfull, 1);
        for (int i = 0; i < 2; ++i) { mbar_init(&ctl->tfull[i], 1); mbar_init(&ctl->tempty[i], 8); }
        mbar_fence_init();
    }
    if (threadIdx.x >= 64 && threadIdx.x < 128) {
        const int c = threadIdx.x - 64;
        ctl->affine[c] = make_float2(scale ? scale[c] : 1.f, shift ? shift[c] : 0.f);
    }
    if (warp == 1) {                                            // TMEM: 512 columns (2 accumulators of 192)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&ctl->tmem_base)), "r"(512u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const unsigned tmem = ctl->tmem_base;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            mbar_expect_tx(&ctl->bfull, 3 * TC_B_BYTES);
            for (int kb = 0; kb < 3; ++kb) tma_load_2d(sB + kb * TC_B_BYTES, &tmB, kb * TC_KBLK, 0, &ctl->bfull);
            int stage = 0;
            unsigned phase = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                const int s0 = tile * TC_OUT_PER_TILE - 1;
                for (int dl = -1; dl <= 1; ++dl) {
                    mbar_wait_bounded(&ctl->empty[stage], phase ^ 1, (dbg & 16) ? 100 : 0);
                    mbar_expect_tx(&ctl->full[stage], TC_A_BYTES);
                    tma_load_2d(sA + stage * TC_A_BYTES, &tmA, 0, s0 + dl * pitch, &ctl->full[stage]);
                    if (++stage == TC_STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (one thread) =====
        if (lane == 0) {
            const unsigned idesc = umma_idesc_bf16(TC_M, TC_N);
            mbar_wait_bounded(&ctl->bfull, 0);
            int stage = 0, acc = 0;
            unsigned phase = 0, aphase = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                mbar_wait_bounded(&ctl->tempty[acc], aphase ^ 1, (dbg & 16) ? 100 : 0);   // epilogue has drained this accumulator
                asm volatile("tcgen05.fence::after_thread_sync;");
                const unsigned d = tmem + acc * 256;                       // accumulators at columns 0 and 256
                for (int kb = 0; kb < 3; ++kb) {
                    mbar_wait_bounded(&ctl->full[stage], phase, (dbg & 16) ? 100 : 0);
                    asm volatile("tcgen05.fence::after_thread_sync;");
                    const unsigned long long da = umma_desc_sw128(sA + stage * TC_A_BYTES);
                    const unsigned long long db = umma_desc_sw128(sB + kb * TC_B_BYTES);
#pragma unroll
                    for (int k = 0; k < TC_KBLK / 16; ++k)                 // 32 bytes (= 2 x 16 B units) per K step
                        umma_f16(d, da + 2 * k, db + 2 * k, idesc, (kb | k) ? 1u : 0u);
                    umma_commit(&ctl->empty[stage]);                       // smem slot free when these MMAs retire
                    if (++stage == TC_STAGES) { stage = 0; phase ^= 1; }
                }
                umma_commit(&ctl->tfull[acc]);                             // accumulator ready for the epilogue
                if (++acc == 2) { acc = 0; aphase ^= 1; }
            }
        }
    } else {
        // ===== epilogue warps 2..9: TMEM lane quadrant q = warp % 4, channel half = (warp - 2) / 4 =====
        const int q = warp & 3;
        const int half = (warp - 2) >> 2;
        const int row = q * 32 + lane;                                     // position inside the tile
        int acc = 0;
        unsigned aphase = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const int s = tile * TC_OUT_PER_TILE - 1 + row;
            mbar_wait_bounded(&ctl->tfull[acc], aphase);
            asm volatile("tcgen05.fence::after_thread_sync;");
            const unsigned t0 = tmem + acc * 256 + ((unsigned)(q * 32) << 16);
            const bool valid = row >= 1 && row <= TC_OUT_PER_TILE && s >= 0 && s < S && (s % pitch) != PW && !(dbg & 1);
#pragma unroll 1
            for (int c = 32 * half; c < ((dbg & 4) ? 0 : 32 * half + 32); c += 16) {
                float tm[16], tz[16], tp[16];
                if (!(dbg & 8)) {
                    tmem_ld16(t0 + c, tm);              // T_-1 own row
                    tmem_ld16(t0 + 64 + c, tz);         // T_0
                    tmem_ld16(t0 + 128 + c, tp);        // T_+1
                    asm volatile("tcgen05.wait::ld.sync.aligned;");
                } else {
#pragma unroll
                    for (int i = 0; i < 16; ++i) { tm[i] = (float)(row + i); tz[i] = (float)(c + i); tp[i] = 1.f; }
                }
                if (dbg & 2) {
                    if (tm[0] + tz[1] + tp[2] == 12345.678f) out[0] = __float2bfloat16(tm[3]);
                    continue;
                }
                const int par = (c >> 4) & 1;
                if (lane == 31) {
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        reinterpret_cast<float4*>(ctl->halo[half][par][q][0])[i] = make_float4(tm[4 * i], tm[4 * i + 1], tm[4 * i + 2], tm[4 * i + 3]);
                }
                if (lane == 0) {
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        reinterpret_cast<float4*>(ctl->halo[half][par][q][1])[i] = make_float4(tp[4 * i], tp[4 * i + 1], tp[4 * i + 2], tp[4 * i + 3]);
                }
                if (half == 0) asm volatile("bar.sync 1, 128;" ::: "memory");     // the four warps of this channel half
                else asm volatile("bar.sync 2, 128;" ::: "memory");
                // neighbour warps' boundary rows, read by every lane (broadcast) so that no branch diverges
                float hu[16], hd[16];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float4 a4 = reinterpret_cast<const float4*>(ctl->halo[half][par][(q + 3) & 3][0])[i];   // previous quadrant's last row
                    const float4 b4 = reinterpret_cast<const float4*>(ctl->halo[half][par][(q + 1) & 3][1])[i];   // next quadrant's first row
                    hu[4 * i] = a4.x; hu[4 * i + 1] = a4.y; hu[4 * i + 2] = a4.z; hu[4 * i + 3] = a4.w;
                    hd[4 * i] = b4.x; hd[4 * i + 1] = b4.y; hd[4 * i + 2] = b4.z; hd[4 * i + 3] = b4.w;
                }
                float o[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    float up = __shfl_up_sync(0xffffffffu, tm[i], 1);       // T_-1 of row - 1
                    float dn = __shfl_down_sync(0xffffffffu, tp[i], 1);     // T_+1 of row + 1
                    up = lane == 0 ? hu[i] : up;                            // (quadrant 0 / 3 edges are halo rows: unused)
                    dn = lane == 31 ? hd[i] : dn;
                    const float2 af = ctl->affine[c + i];
                    const float v = fmaf(up + tz[i] + dn, af.x, af.y);
                    o[i] = v > 0.f ? v : v * slope;
                }
                if (valid) {
                    uint4 pk[2];
                    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(pk);
#pragma unroll
                    for (int i = 0; i < 8; ++i) h[i] = __floats2bfloat162_rn(o[2 * i], o[2 * i + 1]);
                    uint4* dst = reinterpret_cast<uint4*>(out + (long long)s * 64 + c);
                    dst[0] = pk[0];
                    dst[1] = pk[1];
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;");
            if (lane == 0) mbar_arrive(&ctl->tempty[acc]);                 // 8 arrivals (one per epilogue warp)
            if (++acc == 2) { acc = 0; aphase ^= 1; }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u));
    }
}

// ---- thin first / last layers for the bf16 path (CUDA cores; padded bf16 NHWC activations) ----
__global__ void __launch_bounds__(256)
k_conv_first_bf16(const float* __restrict__ img, __nv_bfloat16* __restrict__ out, const float* __restrict__ w, CnnAct a,
                  CnnIo io, int PH, int PW) {
    __shared__ float sw[9 * CNN_C];
    for (int i = threadIdx.x; i < 9 * CNN_C; i += blockDim.x) sw[i] = w[i];
    __syncthreads();
    const long long total = (long long)PH * PW * (CNN_C / 8);
    for (long long id = (long long)blockIdx.x * blockDim.x + threadIdx.x; id < total; id += (long long)gridDim.x * blockDim.x) {
        const int cg = (int)(id % (CNN_C / 8));
        const long long pix = id / (CNN_C / 8);
        const int l = (int)(pix / PW), p = (int)(pix % PW);
        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int dl = -1; dl <= 1; ++dl)
#pragma unroll
            for (int dp = -1; dp <= 1; ++dp) {
                const int ll = l + dl, pp = p + dp;
                if (ll < 0 || ll >= PH || pp < 0 || pp >= PW) continue;
                const float v = cnn_input(io, img[(long long)ll * PW + pp]);
                const float* ww = sw + ((dl + 1) * 3 + (dp + 1)) * CNN_C + cg * 8;
#pragma unroll
                for (int k = 0; k < 8; ++k) acc[k] = fmaf(v, ww[k], acc[k]);
            }
        uint4 pk;
        __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&pk);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            float v0 = acc[2 * k], v1 = acc[2 * k + 1];
            const int c = cg * 8 + 2 * k;
            if (a.scale) { v0 *= a.scale[c]; v1 *= a.scale[c + 1]; }
            if (a.shift) { v0 += a.shift[c]; v1 += a.shift[c + 1]; }
            h[k] = __floats2bfloat162_rn(act(v0, a.slope), act(v1, a.slope));
        }
        *reinterpret_cast<uint4*>(out + ((long long)l * (PW + 1) + p) * CNN_C + cg * 8) = pk;
    }
}

__global__ void __launch_bounds__(256)
k_conv_last_bf16(const __nv_bfloat16* __restrict__ in, const float* __restrict__ img, float* __restrict__ out,
                 const float* __restrict__ w, float bias, CnnIo io, int PH, int PW, const float* __restrict__ xrec,
                 double* __restrict__ mse_log, const int* __restrict__ slot) {
    __shared__ float sw[9 * CNN_C];
    __shared__ float s_err[8];
    for (int i = threadIdx.x; i < 9 * CNN_C; i += blockDim.x) sw[i] = w[i];
    __syncthreads();
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
                const __nv_bfloat162 v = reinterpret_cast<const __nv_bfloat162*>(in + ((long long)ll * (PW + 1) + pp) * CNN_C)[lane];
                const float* ww = sw + ((dl + 1) * 3 + (dp + 1)) * CNN_C + 2 * lane;
                acc = fmaf(__low2float(v), ww[0], fmaf(__high2float(v), ww[1], acc));
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
