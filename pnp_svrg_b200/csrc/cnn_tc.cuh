// 64 -> 64 (and 64 -> 1) channel 3x3 convolution on the 5th-generation tensor cores (tcgen05 / TMEM / TMA), bf16
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
// (neighbour TMEM lanes via warp shuffles).  Each 32-lane TMEM quadrant of a tile is loaded with its own
// 32 consecutive positions (a separate TMA box), overlapping the next quadrant's by two, so rows 0 and 31
// of every quadrant are halo, the shift never crosses a warp, and a tile yields 4 x 30 = 120 outputs
// (93.75 % efficiency); the activation tile is read 3x (once per dl) instead of 9x.  A shifted A operand
// would need N = 64 (one GEMM per tap), which makes the A re-read from shared memory the bound (4 KiB per
// 32-cycle MMA); the N = 192 output-shift form reads A once per 96-cycle MMA.
// The per-channel BatchNorm scale is folded into the bf16 weights by the host; the epilogue adds the
// shift and applies max(v, slope*v).
//
// The last layer (64 -> 1) is the same kernel with N = 16 (three used columns, one per dp): its
// epilogue sums the three shifted columns and applies the denoiser's output map
// (denoisers/RealSN_DnCNN.py:29-39, denoisers/MMODenoise.py:57-66).
//
// Warp roles (576 threads, one persistent CTA per SM): warp 0 TMA producer, warp 1 MMA issuer + TMEM
// allocator, warps 2-17 epilogue in two ping-pong groups of eight (TMEM lane quadrant = warp_id % 4).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cooperative_groups.h>

#include "fft_core.cuh"

namespace pnp {

#define TC_M 128
#define TC_KBLK 64               // one dl: 64 input channels = 128 bytes per row (one swizzle span)
#define TC_A_BYTES (TC_M * 128)  // 16 KiB
#define TC_STAGES 6               // two groups of three (one per dl)
#define TC_Q_ROWS 32             // rows per TMEM lane quadrant = rows per TMA box
#define TC_OUT_PER_Q 30          // rows 0 and 31 of every quadrant are halo
#define TC_OUT_PER_TILE (4 * TC_OUT_PER_Q)
#define TC_EPI_WARPS 16          // 4 per TMEM lane quadrant, 16 output channels each
#define TC_O_WARP_BYTES 1024     // output staging of one 16-channel chunk: 30 rows x 32 B, swizzle-aligned
#define TC_O_BYTES (TC_EPI_WARPS * 2 * TC_O_WARP_BYTES)
#define TC_THREADS (64 + 32 * TC_EPI_WARPS)   // warp 0 TMA, warp 1 MMA, warps 2-17 epilogue

template <int CO> struct TcCfg {                       // CO = 64: middle layers, CO = 1: last layer
    static constexpr int N = CO == 64 ? 192 : 16;      // (dp, co) columns, padded to a legal UMMA N
    static constexpr int B_BYTES = N * 128;            // per K block
    static constexpr int ACC_COLS = CO == 64 ? 256 : 32;
    static constexpr int TMEM_COLS = CO == 64 ? 512 : 64;
};

struct TcSmem {
    unsigned long long full[TC_STAGES], empty[TC_STAGES], bfull, tfull[2], tempty[2];
    unsigned tmem_base;
    float err[TC_EPI_WARPS];
    __align__(16) float shift[64];   // per output channel shift (bias / folded BatchNorm; the scale is folded into the weights)
};

// SPLIT (error-compensated mode, see k_conv_tc): hi AND lo halves of the weights stay resident (6 K blocks), and the
// activation ring shrinks to three stages so that a 64 -> 64 layer still fits the 227 KiB of an SM
#define TC_SPLIT_STAGES 3
template <int CO, bool SPLIT = false> constexpr size_t tc_smem() {
    return (((SPLIT ? 6 : 3) * TcCfg<CO>::B_BYTES + 1023) & ~(size_t)1023) + (SPLIT ? TC_SPLIT_STAGES : TC_STAGES) * TC_A_BYTES +
           (CO == 64 ? TC_O_BYTES : 0) + sizeof(TcSmem) + 1024;
}

__device__ __forceinline__ void mbar_wait_bounded(unsigned long long* bar, unsigned parity) {
    unsigned done = 0;
    for (unsigned spin = 0; !done; ++spin) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (spin > (1u << 26)) __trap();          // a protocol bug must fail loudly, not hang the GPU
    }
}
__device__ __forceinline__ bool elect_one() {
    unsigned pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred != 0;
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
// instruction descriptor kind::f16: D fp32, A/B bf16, both K-major
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
__device__ __forceinline__ void tmem_ld4(unsigned taddr, float (&v)[4]) {
    unsigned r[4];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr));
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = __uint_as_float(r[i]);
}

// Output side of the last layer (CO = 1): the wrapper arithmetic around the network
struct TcLast {
    const float* img;       // network input image (fp32 line layout, pitch PW)
    float* out;             // denoised image
    const float* xrec;      // optional ground truth for the PSNR
    double* mse_log;
    const int* slot;
    float bias;
    CnnIo io;
};

// Epilogue of one 128-position tile of a 64 -> 64 layer for one epilogue warp (TMEM quadrant of accumulator `t0`,
// channel half hf): two chunks of 16 channels -- TMEM -> registers, output shift by warp shuffles, + shift, activation,
// bf16 pack, swizzled staging and one TMA tensor store per chunk.  The accumulator is handed back (tempty) as soon as
// the second chunk has been read.  Shared by k_conv_tc<64> and k_conv_tc_stack.
// SPLIT: the fp32 result leaves as TWO bf16 planes, hi = bf16(v) through tmO and lo = bf16(v - hi) through tmOl (the
// error-compensated mode of k_conv_tc); the two staging buffers of the warp are then used alternately by the four stores.
template <bool SPLIT = false>
__device__ __forceinline__ void tc_epilogue_tile64(unsigned t0, TcSmem* ctl, unsigned char* sO, const CUtensorMap* tmO, int ew,
                                                   int wg, int hf, int lane, int s, bool pad, bool interior, bool relu,
                                                   float slope, int dbg, const CUtensorMap* tmOl = nullptr) {
#pragma unroll
    for (int j = 0; j < 2; ++j) {                                         // two chunks of 16 channels
        const int c = 32 * hf + 16 * j;
        float tm[16], tz[16], tp[16], sh[16];
        tmem_ld16(t0 + c, tm);              // T_-1 of this lane's row
        tmem_ld16(t0 + 64 + c, tz);         // T_0
        tmem_ld16(t0 + 128 + c, tp);        // T_+1
#pragma unroll
        for (int i = 0; i < 4; ++i) reinterpret_cast<float4*>(sh)[i] = reinterpret_cast<const float4*>(ctl->shift + c)[i];
        asm volatile("tcgen05.wait::ld.sync.aligned;");
        if (j == 1) {                                                     // accumulator read: hand it back before the math
            asm volatile("tcgen05.fence::before_thread_sync;");
            if (lane == 0) mbar_arrive(&ctl->tempty[wg]);
        }
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");   // staging buffer j is free again
        __syncwarp();
        if (dbg & 2) continue;
        unsigned hw[8], lw[8];                           // 16 bf16 channels as 8 packed words: hi plane, lo plane (SPLIT)
#pragma unroll
        for (int k = 0; k < 8; ++k) lw[k] = 0u;
        // Two channels per step on the packed fp32 pipe (add.f32x2: the three additions of a value cost 1.5 instructions), and
        // in the ReLU case the activation is the .relu of the bf16 conversion (cvt.rn.relu.bf16x2.f32) instead of a multiply and
        // a max per value: the epilogue's instructions, not the MMAs, fill most of the issue slots of a tile period.  The
        // two activation forms are separate loops under one uniform branch (if-converted, both would issue).
        float2 vv[8];
#pragma unroll
        for (int i = 0; i < 16; i += 2) {
            float2 up2, dn2;
            up2.x = __shfl_up_sync(0xffffffffu, tm[i], 1);                       // T_-1 of row - 1
            up2.y = __shfl_up_sync(0xffffffffu, tm[i + 1], 1);
            dn2.x = __shfl_down_sync(0xffffffffu, tp[i], 1);                     // T_+1 of row + 1
            dn2.y = __shfl_down_sync(0xffffffffu, tp[i + 1], 1);
            vv[i >> 1] = __fadd2_rn(__fadd2_rn(up2, make_float2(tz[i], tz[i + 1])), __fadd2_rn(dn2, make_float2(sh[i], sh[i + 1])));
        }
        if (!SPLIT && relu) {
#pragma unroll
            for (int k = 0; k < 8; ++k) asm volatile("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(hw[k]) : "f"(vv[k].y), "f"(vv[k].x));
        } else {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const float o0 = relu ? fmaxf(vv[k].x, 0.f) : fmaxf(vv[k].x, vv[k].x * slope);     // leaky ReLU needs slope <= 1
                const float o1 = relu ? fmaxf(vv[k].y, 0.f) : fmaxf(vv[k].y, vv[k].y * slope);
                const __nv_bfloat162 hb = __floats2bfloat162_rn(o0, o1);
                hw[k] = (unsigned)__bfloat16_as_ushort(hb.x) | ((unsigned)__bfloat16_as_ushort(hb.y) << 16);
                if (SPLIT) {
                    const float2 hf2 = __bfloat1622float2(hb);
                    const __nv_bfloat162 lb = __floats2bfloat162_rn(o0 - hf2.x, o1 - hf2.y);
                    lw[k] = (unsigned)__bfloat16_as_ushort(lb.x) | ((unsigned)__bfloat16_as_ushort(lb.y) << 16);
                }
            }
        }
        uint4 pk[2], pl[2];
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            pk[k] = pad ? make_uint4(0u, 0u, 0u, 0u) : make_uint4(hw[4 * k], hw[4 * k + 1], hw[4 * k + 2], hw[4 * k + 3]);
            pl[k] = pad ? make_uint4(0u, 0u, 0u, 0u) : make_uint4(lw[4 * k], lw[4 * k + 1], lw[4 * k + 2], lw[4 * k + 3]);
        }
        if (SPLIT) {
            // hi through staging buffer 0, lo through buffer 1, both chunks: a buffer is rewritten once at most one store
            // (the other buffer's) may still be reading
#pragma unroll
            for (int part = 0; part < 2; ++part) {
                unsigned char* stage_o = sO + (ew * 2 + part) * TC_O_WARP_BYTES;
                if (part == 1) {
                    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                    __syncwarp();
                }
                if (interior) {
                    const int r = lane - 1;
#pragma unroll
                    for (int k = 0; k < 2; ++k)
                        *reinterpret_cast<uint4*>(stage_o + r * 32 + ((k ^ ((r >> 2) & 1)) << 4)) = part ? pl[k] : pk[k];
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) {
                    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                                 ::"l"(part ? tmOl : tmO), "r"(smem_u32(stage_o)), "r"(c), "r"(s + 1) : "memory");
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
            }
            continue;
        }
        // 30 rows x 32 B per chunk; 16-byte chunk index XOR (row / 4) % 2  ==  CU_TENSOR_MAP_SWIZZLE_32B
        unsigned char* stage_o = sO + (ew * 2 + j) * TC_O_WARP_BYTES;
        if (interior) {
            const int r = lane - 1;
#pragma unroll
            for (int k = 0; k < 2; ++k) *reinterpret_cast<uint4*>(stage_o + r * 32 + ((k ^ ((r >> 2) & 1)) << 4)) = pk[k];
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
            asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                         ::"l"(tmO), "r"(smem_u32(stage_o)), "r"(c), "r"(s + 1) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
    }
}

// in / out: bf16 [S][64] with S = PH*(PW+1) positions.  tmA over `in` with box 64 x 32: every TMEM lane
// quadrant (32 rows) of a tile gets its own 32 consecutive positions, overlapping its neighbours by
// two, so that rows 0 and 31 of each quadrant are halo and the output shift never leaves a warp.
// tmB over the packed weights [N rows (dp, co)][192 (dl, ci)] (box 64 x N), BatchNorm scale folded in.
// tmO over `out` with box 16 x 30, 32-byte swizzle: every epilogue warp stages 30 rows x 16 channels in shared
// memory and stores them with one TMA tensor store (a per-lane 16-byte store to rows 128 B apart is 32 LSU
// wavefronts per instruction, which made the LSU data pipe the bound; staging + coalesced st.global was slower).
//
// SPLIT = true is the ERROR-COMPENSATED mode (precision 2, "bf16x3"): every fp32 activation and weight travels as two bf16
// numbers, x = hi + lo with hi = bf16(x), lo = bf16(x - hi) (16 mantissa bits together), and a tile accumulates the three
// products  Ah Wh + Ah Wl + Al Wh  into the same fp32 TMEM accumulator (the dropped Al Wl term is 2^-16 of the result):
// 36 MMAs per tile instead of 12, activations read as two planes (tmA hi, tmAl lo) through a ring of three stages in the
// order  Ah(dl=-1) Al(-1) Ah(0) Al(0) Ah(+1) Al(+1), weights hi (tmB) and lo (tmBl) both resident, results written as two
// planes (tmO, tmOl).  The reference runs these nets in fp32 (denoisers/RealSN_DnCNN.py:32-35); this mode gives its
// numbers to ~1e-5 on the tensor cores, the plain bf16 mode (~1e-2) stays the fast one.
template <int CO, bool SPLIT = false>
__global__ void __launch_bounds__(TC_THREADS, 1)
k_conv_tc(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
          const __grid_constant__ CUtensorMap tmO, const float* __restrict__ shift, float slope, int PW, int S, int n_tiles,
          TcLast last, int dbg, const __grid_constant__ CUtensorMap tmAl, const __grid_constant__ CUtensorMap tmBl,
          const __grid_constant__ CUtensorMap tmOl) {
    using CF = TcCfg<CO>;
    constexpr int NKB = SPLIT ? 6 : 3;                                         // resident weight blocks: (hi | lo) x dl
    constexpr int NST = SPLIT ? TC_SPLIT_STAGES : TC_STAGES;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<unsigned long long>(smem_raw) + 1023ull) & ~1023ull);
    unsigned char* sB = base;                                                  // 3 K blocks (SPLIT: hi blocks, then lo blocks)
    unsigned char* sA = base + ((NKB * CF::B_BYTES + 1023) & ~1023);           // NST x 16 KiB
    unsigned char* sO = sA + NST * TC_A_BYTES;                                 // output staging, 4 quadrants x 2 x 4 KiB (CO = 64)
    TcSmem* ctl = reinterpret_cast<TcSmem*>(sO + (CO == 64 ? TC_O_BYTES : 0));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int pitch = PW + 1;

    if (threadIdx.x == 0) {
        for (int i = 0; i < TC_STAGES; ++i) { mbar_init(&ctl->full[i], 1); mbar_init(&ctl->empty[i], 1); }
        mbar_init(&ctl->bfull, 1);
        for (int i = 0; i < 2; ++i) { mbar_init(&ctl->tfull[i], 1); mbar_init(&ctl->tempty[i], TC_EPI_WARPS / 2); }
        mbar_fence_init();
    }
    if (threadIdx.x < TC_EPI_WARPS) ctl->err[threadIdx.x] = 0.f;
    if (threadIdx.x >= 64 && threadIdx.x < 128) ctl->shift[threadIdx.x - 64] = (CO == 64 && shift) ? shift[threadIdx.x - 64] : 0.f;
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&ctl->tmem_base)), "r"((unsigned)CF::TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const unsigned tmem = ctl->tmem_base;

    // Pipeline indexing: tile iteration `it` of this CTA uses accumulator g = it & 1 and the three smem stages
    // 3g .. 3g+2 (one per dl); every barrier of group g is used once per two tiles, so all share the phase
    // (it >> 1) & 1.  The loops below are unrolled over g to keep stage indices and descriptors static, and run
    // by the WHOLE warp with one elected lane issuing, so that addresses and descriptors stay in uniform registers.
    if (warp == 0) {
        // ===== TMA producer =====
        if (elect_one()) {
            mbar_expect_tx(&ctl->bfull, NKB * CF::B_BYTES);
            for (int kb = 0; kb < 3; ++kb) tma_load_2d(sB + kb * CF::B_BYTES, &tmB, kb * TC_KBLK, 0, &ctl->bfull);
            if (SPLIT)
                for (int kb = 0; kb < 3; ++kb) tma_load_2d(sB + (3 + kb) * CF::B_BYTES, &tmBl, kb * TC_KBLK, 0, &ctl->bfull);
        }
        __syncwarp();
        unsigned ph = 0;
        if constexpr (SPLIT) {
            // six loads per tile through a ring of three stages: load q = 2 * kb + (lo plane ? 1 : 0) uses stage q % 3 for the
            // (q / 3)-th time in this tile, and every stage is used exactly twice per tile, so the barrier parities are static
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                const int s0 = tile * TC_OUT_PER_TILE - 1;
#pragma unroll
                for (int q6 = 0; q6 < 6; ++q6) {
                    const int stage = q6 % 3, use = q6 / 3, kb = q6 >> 1;
                    mbar_wait_bounded(&ctl->empty[stage], (unsigned)(use ^ 1));
                    if (elect_one()) {
                        mbar_expect_tx(&ctl->full[stage], TC_A_BYTES);
#pragma unroll
                        for (int q = 0; q < 4; ++q)
                            tma_load_2d(sA + stage * TC_A_BYTES + q * (TC_Q_ROWS * 128), (q6 & 1) ? &tmAl : &tmA, 0,
                                        s0 + q * TC_OUT_PER_Q + (kb - 1) * pitch, &ctl->full[stage]);
                    }
                    __syncwarp();
                }
            }
        } else
        for (int tile = blockIdx.x; tile < n_tiles; ph ^= 1) {
#pragma unroll
            for (int g = 0; g < 2; ++g) {
                if (tile < n_tiles) {
                    const int s0 = tile * TC_OUT_PER_TILE - 1;
#pragma unroll
                    for (int kb = 0; kb < 3; ++kb) {
                        const int stage = 3 * g + kb;
                        mbar_wait_bounded(&ctl->empty[stage], ph ^ 1);
                        if (elect_one()) {
                            mbar_expect_tx(&ctl->full[stage], (kb == 0 || !(dbg & 4)) ? TC_A_BYTES : 0);
#pragma unroll
                            for (int q = 0; q < 4; ++q)
                                if (kb == 0 || !(dbg & 4)) tma_load_2d(sA + stage * TC_A_BYTES + q * (TC_Q_ROWS * 128), &tmA, 0,
                                            s0 + q * TC_OUT_PER_Q + (kb - 1) * pitch, &ctl->full[stage]);
                        }
                        __syncwarp();
                    }
                }
                tile += gridDim.x;
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        const unsigned idesc = umma_idesc_bf16(TC_M, CF::N);
        unsigned long long da[NST], db[NKB];
#pragma unroll
        for (int i = 0; i < NST; ++i) da[i] = umma_desc_sw128(sA + i * TC_A_BYTES);
#pragma unroll
        for (int kb = 0; kb < NKB; ++kb) db[kb] = umma_desc_sw128(sB + kb * CF::B_BYTES);
        mbar_wait_bounded(&ctl->bfull, 0);
        unsigned ph = 0;
        for (int tile = blockIdx.x; tile < n_tiles; ph ^= 1) {
#pragma unroll
            for (int g = 0; g < 2; ++g) {
                if (tile < n_tiles) {
                    mbar_wait_bounded(&ctl->tempty[g], ph ^ 1);            // epilogue has drained this accumulator
                    asm volatile("tcgen05.fence::after_thread_sync;");
                    const unsigned d = tmem + g * CF::ACC_COLS;
                    if constexpr (SPLIT) {
#pragma unroll
                        for (int q6 = 0; q6 < 6; ++q6) {
                            const int stage = q6 % 3, use = q6 / 3, kb = q6 >> 1;
                            mbar_wait_bounded(&ctl->full[stage], (unsigned)use);
                            asm volatile("tcgen05.fence::after_thread_sync;");
                            if (elect_one()) {
                                if ((q6 & 1) == 0) {                       // Ah: against Wh and Wl
#pragma unroll
                                    for (int k = 0; k < TC_KBLK / 16; ++k)
                                        umma_f16(d, da[stage] + 2 * k, db[kb] + 2 * k, idesc, (kb | k) ? 1u : 0u);
#pragma unroll
                                    for (int k = 0; k < TC_KBLK / 16; ++k)
                                        umma_f16(d, da[stage] + 2 * k, db[3 + kb] + 2 * k, idesc, 1u);
                                } else {                                   // Al: against Wh
#pragma unroll
                                    for (int k = 0; k < TC_KBLK / 16; ++k)
                                        umma_f16(d, da[stage] + 2 * k, db[kb] + 2 * k, idesc, 1u);
                                }
                                umma_commit(&ctl->empty[stage]);
                                if (q6 == 5) umma_commit(&ctl->tfull[g]);
                            }
                            __syncwarp();
                        }
                    } else {
#pragma unroll
                    for (int kb = 0; kb < 3; ++kb) {
                        const int stage = 3 * g + kb;
                        mbar_wait_bounded(&ctl->full[stage], ph);
                        asm volatile("tcgen05.fence::after_thread_sync;");
                        if (elect_one()) {
#pragma unroll
                            for (int k = 0; k < TC_KBLK / 16; ++k)         // 32 bytes (= 2 x 16 B units) per K step
                                umma_f16(d, da[stage] + 2 * k, db[kb] + 2 * k, idesc, (kb | k) ? 1u : 0u);
                            umma_commit(&ctl->empty[stage]);               // smem slot free when these MMAs retire
                            if (kb == 2) umma_commit(&ctl->tfull[g]);      // accumulator ready for the epilogue
                        }
                        __syncwarp();
                    }
                    }
                }
                tile += gridDim.x;
            }
        }
    } else {
        // ===== epilogue warps 2..17 =====
        // Two groups of eight warps ping-pong over the tiles: group wg owns accumulator wg and takes the tiles with
        // it % 2 == wg, so its TMEM loads overlap the other group's arithmetic and a tile's epilogue may take up to
        // two MMA periods.  Inside a group: TMEM lane quadrant q = warp % 4 (a hardware rule), channel half hf.
        const int ew = warp - 2;
        const int wg = ew >> 3;
        const int q = warp & 3;
        const int hf = (ew >> 2) & 1;
        const bool interior = lane >= 1 && lane <= TC_OUT_PER_Q;
        const bool relu = slope == 0.f;
        unsigned aphase = 0;
        float err = 0.f;
        const unsigned t0 = tmem + wg * CF::ACC_COLS + ((unsigned)(q * 32) << 16);
        for (int tile = blockIdx.x + wg * gridDim.x; tile < n_tiles; tile += 2 * gridDim.x, aphase ^= 1) {
            const int s = tile * TC_OUT_PER_TILE + q * TC_OUT_PER_Q + lane - 1;       // position of this lane's row
            mbar_wait_bounded(&ctl->tfull[wg], aphase);
            asm volatile("tcgen05.fence::after_thread_sync;");
            if (CO == 64) {
                const bool pad = (s % pitch) == PW;                                   // the pad pixel of a line stays zero
                tc_epilogue_tile64<SPLIT>(t0, ctl, sO, &tmO, ew, wg, hf, lane, s, pad, interior, relu, slope, dbg, &tmOl);
            } else {
                // last layer: one output per position + the wrapper's output map
                float t[4];
                if (hf == 0) tmem_ld4(t0, t);
                asm volatile("tcgen05.wait::ld.sync.aligned;");
                asm volatile("tcgen05.fence::before_thread_sync;");
                if (lane == 0) mbar_arrive(&ctl->tempty[wg]);
                if (hf == 0) {
                    const float up = __shfl_up_sync(0xffffffffu, t[0], 1);
                    const float dn = __shfl_down_sync(0xffffffffu, t[2], 1);
                    const float conv = up + t[1] + dn;
                    if (interior && s < S && (s % pitch) != PW) {
                        const int l = s / pitch, p = s - l * pitch;
                        const long long pix = (long long)l * PW + p;
                        const float x = last.img[pix];
                        float res;
                        if (last.io.mode == 0) {
                            const float mn = ord2f(last.io.stats[0]), mx = ord2f(last.io.stats[1]);
                            const float xt = (x - mn) / (mx - mn) * last.io.range + last.io.shift;
                            res = ((xt - conv) - last.io.shift) / last.io.range * (mx - mn) + mn;
                        } else {
                            res = fminf(fmaxf(conv + last.bias + fminf(fmaxf(x, 0.f), 1.f), 0.f), 1.f);
                        }
                        last.out[pix] = res;
                        if (last.xrec) { const float df = res - last.xrec[pix]; err = fmaf(df, df, err); }
                    }
                }
            }
        }
        if (CO == 64 && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        if (CO == 1 && last.xrec && last.mse_log) {
            err = warp_sum_f(err);
            if (lane == 0) ctl->err[ew] = err;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (CO == 1 && last.xrec && last.mse_log && threadIdx.x == 0) {
        float t = 0.f;
        for (int k = 0; k < TC_EPI_WARPS; ++k) t += ctl->err[k];
        atomicAdd(last.mse_log + (last.slot ? *last.slot : 0), (double)t);
    }
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((unsigned)CF::TMEM_COLS));
    }
}

// ---- all 64 -> 64 layers of a stack in ONE cooperative launch (small images) ----------------------------------------
// At 256 x 256 a layer is 548 tiles = 3.7 per SM, ~2.3 us of MMA inside a ~10 us launch (launch latency, TMEM
// allocation, barrier init, pipeline fill and drain per layer).  Here the CTAs stay resident: per layer they reload
// the 72 KiB of weights, run the same TMA -> tcgen05 -> epilogue pipeline over their tiles (barrier phases simply
// continue), make their TMA stores visible and meet at a grid barrier; activations ping-pong between the two
// buffers (L2 resident at these sizes).
#define TC_MAX_LAYERS 30
struct TcStack {
    CUtensorMap a[2];                  // loads:  layer l reads  a[l & 1]   (a[0] over buffer 0, a[1] over buffer 1)
    CUtensorMap o[2];                  // stores: layer l writes o[l & 1]   (o[0] over buffer 1, o[1] over buffer 0)
    CUtensorMap b[TC_MAX_LAYERS];      // packed weights per layer
    const float* shift[TC_MAX_LAYERS];
    float slope[TC_MAX_LAYERS];
    int n_layers;
};

__global__ void __launch_bounds__(TC_THREADS, 1)
k_conv_tc_stack(const __grid_constant__ TcStack pm, int PW, int S, int n_tiles) {
    using CF = TcCfg<64>;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<unsigned long long>(smem_raw) + 1023ull) & ~1023ull);
    unsigned char* sB = base;
    unsigned char* sA = base + ((3 * CF::B_BYTES + 1023) & ~1023);
    unsigned char* sO = sA + TC_STAGES * TC_A_BYTES;
    TcSmem* ctl = reinterpret_cast<TcSmem*>(sO + TC_O_BYTES);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int pitch = PW + 1;
    cooperative_groups::grid_group grid = cooperative_groups::this_grid();

    if (threadIdx.x == 0) {
        for (int i = 0; i < TC_STAGES; ++i) { mbar_init(&ctl->full[i], 1); mbar_init(&ctl->empty[i], 1); }
        mbar_init(&ctl->bfull, 1);
        for (int i = 0; i < 2; ++i) { mbar_init(&ctl->tfull[i], 1); mbar_init(&ctl->tempty[i], TC_EPI_WARPS / 2); }
        mbar_fence_init();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&ctl->tmem_base)), "r"((unsigned)CF::TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const unsigned tmem = ctl->tmem_base;

    // pipeline state that carries over from layer to layer: the phase of the barriers of each group g (a layer with
    // an odd number of tiles uses group 0 once more than group 1)
    unsigned ph[2] = {0u, 0u};            // producer / issuer: smem stages 3g .. 3g+2 and accumulator g
    unsigned aphase = 0;                  // epilogue group: its accumulator
    const unsigned idesc = umma_idesc_bf16(TC_M, CF::N);
    unsigned long long da[TC_STAGES], db[3];
#pragma unroll
    for (int i = 0; i < TC_STAGES; ++i) da[i] = umma_desc_sw128(sA + i * TC_A_BYTES);
#pragma unroll
    for (int kb = 0; kb < 3; ++kb) db[kb] = umma_desc_sw128(sB + kb * CF::B_BYTES);

    for (int layer = 0; layer < pm.n_layers; ++layer) {
        if (layer > 0) {
            grid.sync();                                  // every CTA's outputs of the previous layer are in global memory
            asm volatile("fence.proxy.async.global;" ::: "memory");
        }
        if (threadIdx.x >= 64 && threadIdx.x < 128) ctl->shift[threadIdx.x - 64] = pm.shift[layer] ? pm.shift[layer][threadIdx.x - 64] : 0.f;
        __syncthreads();
        const CUtensorMap* tmA = &pm.a[layer & 1];
        const CUtensorMap* tmO = &pm.o[layer & 1];
        if (warp == 0) {
            // ===== TMA producer =====
            if (elect_one()) {
                mbar_expect_tx(&ctl->bfull, 3 * CF::B_BYTES);
                for (int kb = 0; kb < 3; ++kb) tma_load_2d(sB + kb * CF::B_BYTES, &pm.b[layer], kb * TC_KBLK, 0, &ctl->bfull);
            }
            __syncwarp();
            for (int tile = blockIdx.x; tile < n_tiles;) {
#pragma unroll
                for (int g = 0; g < 2; ++g) {
                    if (tile < n_tiles) {
                        const int s0 = tile * TC_OUT_PER_TILE - 1;
#pragma unroll
                        for (int kb = 0; kb < 3; ++kb) {
                            const int stage = 3 * g + kb;
                            mbar_wait_bounded(&ctl->empty[stage], ph[g] ^ 1);
                            if (elect_one()) {
                                mbar_expect_tx(&ctl->full[stage], TC_A_BYTES);
#pragma unroll
                                for (int q = 0; q < 4; ++q)
                                    tma_load_2d(sA + stage * TC_A_BYTES + q * (TC_Q_ROWS * 128), tmA, 0,
                                                s0 + q * TC_OUT_PER_Q + (kb - 1) * pitch, &ctl->full[stage]);
                            }
                            __syncwarp();
                        }
                        ph[g] ^= 1;
                    }
                    tile += gridDim.x;
                }
            }
        } else if (warp == 1) {
            // ===== MMA issuer =====
            mbar_wait_bounded(&ctl->bfull, layer & 1);
            for (int tile = blockIdx.x; tile < n_tiles;) {
#pragma unroll
                for (int g = 0; g < 2; ++g) {
                    if (tile < n_tiles) {
                        mbar_wait_bounded(&ctl->tempty[g], ph[g] ^ 1);
                        asm volatile("tcgen05.fence::after_thread_sync;");
                        const unsigned d = tmem + g * CF::ACC_COLS;
#pragma unroll
                        for (int kb = 0; kb < 3; ++kb) {
                            const int stage = 3 * g + kb;
                            mbar_wait_bounded(&ctl->full[stage], ph[g]);
                            asm volatile("tcgen05.fence::after_thread_sync;");
                            if (elect_one()) {
#pragma unroll
                                for (int k = 0; k < TC_KBLK / 16; ++k)
                                    umma_f16(d, da[stage] + 2 * k, db[kb] + 2 * k, idesc, (kb | k) ? 1u : 0u);
                                umma_commit(&ctl->empty[stage]);
                                if (kb == 2) umma_commit(&ctl->tfull[g]);
                            }
                            __syncwarp();
                        }
                        ph[g] ^= 1;
                    }
                    tile += gridDim.x;
                }
            }
        } else {
            // ===== epilogue warps: two ping-pong groups of eight, as in k_conv_tc =====
            const int ew = warp - 2;
            const int wg = ew >> 3;
            const int q = warp & 3;
            const int hf = (ew >> 2) & 1;
            const bool interior = lane >= 1 && lane <= TC_OUT_PER_Q;
            const float slope = pm.slope[layer];
            const bool relu = slope == 0.f;
            const unsigned t0 = tmem + wg * CF::ACC_COLS + ((unsigned)(q * 32) << 16);
            for (int tile = blockIdx.x + wg * gridDim.x; tile < n_tiles; tile += 2 * gridDim.x, aphase ^= 1) {
                const int s = tile * TC_OUT_PER_TILE + q * TC_OUT_PER_Q + lane - 1;
                mbar_wait_bounded(&ctl->tfull[wg], aphase);
                asm volatile("tcgen05.fence::after_thread_sync;");
                const bool pad = (s % pitch) == PW;
                tc_epilogue_tile64(t0, ctl, sO, tmO, ew, wg, hf, lane, s, pad, interior, relu, slope, 0);
            }
            if (lane == 0) {                              // this layer's stores are complete before the grid barrier
                asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
                asm volatile("fence.proxy.async.global;" ::: "memory");
            }
        }
        __threadfence();
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((unsigned)CF::TMEM_COLS));
    }
}

// ---- thin first layer for the bf16 path (CUDA cores; writes padded bf16 NHWC activations) ----
// first layer (1 -> 64).  A thread owns 8 output channels and keeps their 72 weights in REGISTERS; a warp is 4 adjacent
// pixels of a line x the 8 channel groups, so every store instruction of the warp writes 4 complete 128-byte pixels
// (512 contiguous bytes).  The thread walks down FL_R lines with a sliding 3x3 window (3 new inputs per step).  The
// version this replaces (one thread per pixel, weights re-read from shared memory for every pixel: 144 LDS.128 and
// eight 16-byte stores to eight different lines per pixel) ran 1107 warp instructions per 32 pixels, 40 % issue-active,
// 331 us at 2048^2 for a 537 MB write; this one runs ~600.
#define FL_R 16
__global__ void __launch_bounds__(256)
k_conv_first_bf16(const float* __restrict__ img, __nv_bfloat16* __restrict__ out, const float* __restrict__ w, CnnAct a,
                  CnnIo io, int PH, int PW, __nv_bfloat16* __restrict__ out_lo) {       // out_lo: the residual plane of the split mode
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int cg = lane & 7;                                   // channels 8*cg .. 8*cg + 7
    const int p = blockIdx.x * 32 + warp * 4 + (lane >> 3);    // pixel within the line
    const int l0 = blockIdx.y * FL_R;
    float2 wt[9][4];
#pragma unroll
    for (int t = 0; t < 9; ++t) {
        const float4 lo = *reinterpret_cast<const float4*>(w + t * CNN_C + cg * 8), hi = *reinterpret_cast<const float4*>(w + t * CNN_C + cg * 8 + 4);
        wt[t][0] = make_float2(lo.x, lo.y); wt[t][1] = make_float2(lo.z, lo.w);
        wt[t][2] = make_float2(hi.x, hi.y); wt[t][3] = make_float2(hi.z, hi.w);
    }
    float2 sc[4], sf[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        sc[k] = a.scale ? *reinterpret_cast<const float2*>(a.scale + cg * 8 + 2 * k) : make_float2(1.f, 1.f);
        sf[k] = a.shift ? *reinterpret_cast<const float2*>(a.shift + cg * 8 + 2 * k) : make_float2(0.f, 0.f);
    }
    // input map  x -> x * ia + ib  (mode 0: min/max normalisation to [shift, shift + range]; mode 1: clamp to [0, 1])
    float ia = 1.f, ib = 0.f;
    if (io.mode == 0) {
        const float mn = ord2f(io.stats[0]), mx = ord2f(io.stats[1]);
        ia = io.range / (mx - mn);
        ib = io.shift - mn * ia;
    }
    // the CTA's input tile (FL_R + 2 lines x 34 pixels, already transformed, zero outside the image) goes through
    // shared memory: one coalesced load phase instead of three dependent global loads per step of every thread
    __shared__ float tile[FL_R + 2][36];
    for (int i = threadIdx.x; i < (FL_R + 2) * 34; i += blockDim.x) {
        const int r = i / 34, c = i - r * 34;
        const int ll = l0 - 1 + r, pp = blockIdx.x * 32 - 1 + c;
        float v = 0.f;
        if (ll >= 0 && ll < PH && pp >= 0 && pp < PW) {
            const float x = img[(long long)ll * PW + pp];
            v = io.mode == 0 ? fmaf(x, ia, ib) : fminf(fmaxf(x, 0.f), 1.f);
        }
        tile[r][c] = v;
    }
    __syncthreads();
    const int pc = warp * 4 + (lane >> 3);                     // pixel within the tile; tile column pc + 1
    auto load3 = [&](int r, float (&o)[3]) {
#pragma unroll
        for (int d = 0; d < 3; ++d) o[d] = tile[r][pc + d];
    };
    if (p >= PW) return;
    float win[3][3];
    load3(0, win[0]);
    load3(1, win[1]);
    const float slope = a.slope;
#pragma unroll 2
    for (int l = l0; l < l0 + FL_R && l < PH; ++l) {
        load3(l - l0 + 2, win[2]);
        float2 acc[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) acc[k] = make_float2(0.f, 0.f);
#pragma unroll
        for (int dl = 0; dl < 3; ++dl)
#pragma unroll
            for (int dp = 0; dp < 3; ++dp) {
                const float2 xx = make_float2(win[dl][dp], win[dl][dp]);
#pragma unroll
                for (int k = 0; k < 4; ++k) acc[k] = __ffma2_rn(xx, wt[dl * 3 + dp][k], acc[k]);
            }
        uint4 pk, pl;
        __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&pk);
        __nv_bfloat162* hl = reinterpret_cast<__nv_bfloat162*>(&pl);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float2 v = __ffma2_rn(acc[k], sc[k], sf[k]);
            const float a0 = act(v.x, slope), a1 = act(v.y, slope);
            h[k] = __floats2bfloat162_rn(a0, a1);
            const float2 hf2 = __bfloat1622float2(h[k]);
            hl[k] = __floats2bfloat162_rn(a0 - hf2.x, a1 - hf2.y);
        }
        *reinterpret_cast<uint4*>(out + ((long long)l * (PW + 1) + p) * CNN_C + cg * 8) = pk;
        if (out_lo) *reinterpret_cast<uint4*>(out_lo + ((long long)l * (PW + 1) + p) * CNN_C + cg * 8) = pl;
#pragma unroll
        for (int d = 0; d < 3; ++d) { win[0][d] = win[1][d]; win[1][d] = win[2][d]; }
    }
}

}  // namespace pnp
