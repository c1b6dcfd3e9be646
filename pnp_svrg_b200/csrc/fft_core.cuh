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

// ------------------------------------------------------------------ optional event trace (builds with -DPNP_TRACE)
// thread 0 of a CTA stamps (%globaltimer, tag) into a small shared-memory list (no global traffic while the kernel
// works); trace_flush() at the end of the kernel appends the list to a global buffer with ONE atomic per CTA.
// Read back with pnp_debug_read(1, ...): 16-byte records (u64 ns, i32 tag, i16 blockIdx.x, i16 %smid).
struct TraceEv { unsigned long long t; int tag; short cta; short sm; };
#ifdef PNP_TRACE
#define PNP_TRACE_MAX (1 << 20)
#ifndef PNP_TRACE_CTA
#define PNP_TRACE_CTA 40
#endif
__device__ TraceEv g_trace[PNP_TRACE_MAX];
__device__ unsigned g_trace_n;
__device__ int g_dbg_flags;          // experiment switches of trace builds (pnp_debug_set key 3)
#define PNP_DBG(bit) (g_dbg_flags & (bit))
struct TraceLocal { unsigned long long t[PNP_TRACE_CTA]; int tag[PNP_TRACE_CTA]; int n; };
__device__ __forceinline__ TraceLocal* trace_local() {
    __shared__ TraceLocal s_tl;
    return &s_tl;
}
// first = true on the first call of a kernel (resets the CTA's list)
__device__ __forceinline__ void trace(int tag, bool first = false) {
    if (threadIdx.x == 0) {
        TraceLocal* tl = trace_local();
        if (first) tl->n = 0;
        const int i = tl->n;
        if (i < PNP_TRACE_CTA) {
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            tl->t[i] = t;
            tl->tag[i] = tag;
            tl->n = i + 1;
        }
    }
}
__device__ __forceinline__ void trace_flush() {
    if (threadIdx.x == 0) {
        TraceLocal* tl = trace_local();
        const int n = tl->n;
        unsigned sm;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(sm));
        const unsigned base = atomicAdd(&g_trace_n, (unsigned)n);
        for (int i = 0; i < n; ++i)
            if (base + i < PNP_TRACE_MAX) g_trace[base + i] = TraceEv{tl->t[i], tl->tag[i], (short)blockIdx.x, (short)sm};
        tl->n = 0;
    }
}
#else
__device__ __forceinline__ void trace(int, bool = false) {}
__device__ __forceinline__ void trace_flush() {}
#define PNP_DBG(bit) 0
#endif

// ------------------------------------------------------------------ TMA bulk copy + mbarrier (sm_90+)
// 1-D bulk copies global -> shared issued by ONE thread; completion is signalled on an mbarrier by
// byte count (complete_tx), every consumer thread waits on the barrier's phase parity.
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// L2 eviction policy for the image-sized data streams (iterate, snapshot, spectrum, ground truth ...): evict-first.
// One inner iteration at 2048^2 moves ~130 MB through an L2 whose useful capacity is about half of its 126 MB, so with
// the default policy NOTHING survives from one iteration to the next -- including the kernels' own code, which is
// straight-line and run once or twice per CTA: its instruction fetches then come from DRAM (measured: the same
// 450-instruction block took 4.1 us the first time and 0.8 us the second).  With the streams marked evict-first the
// code, the twiddle table and the small per-iteration state stay resident.
__device__ __forceinline__ unsigned long long l2_stream_policy() {
    unsigned long long p;
    asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
#ifndef PNP_NO_L2_HINTS
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(l2_stream_policy()) : "memory");
#else
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
#endif
}
// streaming global loads / stores with the same policy
__device__ __forceinline__ float4 ldg_stream(const float4* p) {
    float4 v;
#ifndef PNP_NO_L2_HINTS
    asm volatile("ld.global.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p), "l"(l2_stream_policy()));
#else
    v = *p;
#endif
    return v;
}
__device__ __forceinline__ float ldg_stream(const float* p) {
    float v;
#ifndef PNP_NO_L2_HINTS
    asm volatile("ld.global.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v) : "l"(p), "l"(l2_stream_policy()));
#else
    v = *p;
#endif
    return v;
}
__device__ __forceinline__ void stg_stream(float4* p, float4 v) {
#ifndef PNP_NO_L2_HINTS
    asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1, %2, %3, %4}, %5;"
                 ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "l"(l2_stream_policy()) : "memory");
#else
    *p = v;
#endif
}
__device__ __forceinline__ void stg_stream(float2* p, float2 v) {
#ifndef PNP_NO_L2_HINTS
    asm volatile("st.global.L2::cache_hint.v2.f32 [%0], {%1, %2}, %3;"
                 ::"l"(p), "f"(v.x), "f"(v.y), "l"(l2_stream_policy()) : "memory");
#else
    *p = v;
#endif
}
__device__ __forceinline__ void stg_stream(float* p, float v) {
#ifndef PNP_NO_L2_HINTS
    asm volatile("st.global.L2::cache_hint.f32 [%0], %1, %2;" ::"l"(p), "f"(v), "l"(l2_stream_policy()) : "memory");
#else
    *p = v;
#endif
}
// Hand-off data -- written by one pass and read by the NEXT one (the spectrum between the three passes, the iterate
// between the tail and the next forward line pass: 34 MB of the 126 MB L2 at 2048^2).  -DPNP_L2_KEEP gives it the
// default policy instead of evict-first so that it is still in L2 when its consumer starts -- measured on B200: 64.0 us
// per inner iteration against 63.2 with everything evict-first (the passes are latency / issue bound, and the data
// mostly survives anyway), so it is off.
#ifdef PNP_L2_KEEP
__device__ __forceinline__ void stg_keep(float4* p, float4 v) { *p = v; }
__device__ __forceinline__ void stg_keep(float2* p, float2 v) { *p = v; }
__device__ __forceinline__ void bulk_g2s_keep(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
#else
__device__ __forceinline__ void stg_keep(float4* p, float4 v) { stg_stream(p, v); }
__device__ __forceinline__ void stg_keep(float2* p, float2 v) { stg_stream(p, v); }
__device__ __forceinline__ void bulk_g2s_keep(void* dst, const void* src, unsigned bytes, unsigned long long* bar) { bulk_g2s(dst, src, bytes, bar); }
#endif
// bring `bytes` (multiple of 16) at `src` into L2 ahead of use; no destination, nothing to wait for
__device__ __forceinline__ void bulk_prefetch_l2(const void* src, unsigned bytes) {
#ifndef PNP_NO_L2_HINTS
    asm volatile("cp.async.bulk.prefetch.L2.global.L2::cache_hint [%0], %1, %2;" ::"l"(src), "r"(bytes), "l"(l2_stream_policy()) : "memory");
#else
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
#endif
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "MBAR_WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra MBAR_DONE_%=;\n\t"
        "bra MBAR_WAIT_%=;\n\t"
        "MBAR_DONE_%=:\n\t"
        "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

// ------------------------------------------------------------------ programmatic dependent launch + software grid barrier
// A pass launched with the programmatic-stream-serialization attribute may START while the previous kernel of the
// stream is still finishing (its CTAs become resident as SMs free up and run their prologue); griddep_wait() returns
// once that kernel has completed and flushed.  Rule used by every pass here: griddep_launch() only AFTER the own
// griddep_wait(), so a kernel that runs at all knows that everything older than its direct predecessor is complete
// and may read such data (the iterate of the previous iteration, the snapshot, mu, the ground truth) BEFORE the wait.
// Both are no-ops in a normal launch.
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void griddep_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// Grid-wide barrier of a persistent kernel whose CTAs are all co-resident (grid <= SMs x occupancy; the caller
// guarantees that no other kernel spinning on such a barrier shares the device).  ws: two 32-bit words used as ONE
// 64-bit arrival counter that only ever grows (zero before the first use; 2^64 arrivals never wrap): arrival k belongs
// to generation k / nblocks, and a CTA leaves when the counter has reached the end of its generation.  The workspace
// therefore belongs to ONE grid size (every launch that uses it has the same nblocks, so the counter is a multiple of
// nblocks between launches).  On the critical path of the last arriver there is one atomic on its way to L2 and one
// poll of the waiters -- the first version (arrival count + generation word: read the generation, add, reset, fence,
// release, poll) had four dependent L2 round trips there and was slower than the cooperative launch's barrier.
__device__ __forceinline__ void sw_grid_sync(unsigned* ws, unsigned nblocks) {
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long* c = reinterpret_cast<unsigned long long*>(ws);
        __threadfence();
        const unsigned long long ticket = atomicAdd(c, 1ull);
        const unsigned long long target = (ticket / nblocks + 1ull) * nblocks;
        unsigned long long v;
        do {
            asm volatile("ld.acquire.gpu.u64 %0, [%1];" : "=l"(v) : "l"(c) : "memory");
        } while (v < target);
    }
    __syncthreads();
}

// Complex arithmetic on the packed fp32 pipe of sm_100 (add/mul/fma.rn.f32x2): FADD2 takes a negated operand,
// FMUL2 / FFMA2 take a scalar broadcast (`.F32`), swapped halves (`.LO_HI`) and a per-half sign (`.NP`), so a complex
// add or subtract is ONE instruction and a complex multiply TWO (SASS checked) -- half the issue slots of the
// scalar forms, and the butterflies are what the FFT passes issue most.
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return __fadd2_rn(a, make_float2(-b.x, -b.y)); }
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    const float2 t = __fmul2_rn(a, make_float2(b.x, b.x));                       // (a.x b.x, a.y b.x)
    return __ffma2_rn(make_float2(-a.y, a.x), make_float2(b.y, b.y), t);         // + (-a.y b.y, a.x b.y)
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
// V = plan variant (0 everywhere but the cluster kernel of csrc/small.cuh): variant 1 of length 256 is 8 x 8 x 4 with 8
// values per thread, i.e. 32 threads = one whole warp per transform -- half the dependent work per thread of the 16 x 16
// plan (16 threads), for a phase that has only as many transforms as the CTA has warps.
template <int L, int V = 0> struct FftPlan;
template <> struct FftPlan<8, 0>    { static constexpr int NS = 1, R0 = 8,  R1 = 1,  R2 = 1,  EPT = 8;  };
template <> struct FftPlan<16, 0>   { static constexpr int NS = 1, R0 = 16, R1 = 1,  R2 = 1,  EPT = 16; };
template <> struct FftPlan<32, 0>   { static constexpr int NS = 2, R0 = 8,  R1 = 4,  R2 = 1,  EPT = 8;  };
template <> struct FftPlan<64, 0>   { static constexpr int NS = 2, R0 = 8,  R1 = 8,  R2 = 1,  EPT = 8;  };
template <> struct FftPlan<128, 0>  { static constexpr int NS = 2, R0 = 16, R1 = 8,  R2 = 1,  EPT = 16; };
template <> struct FftPlan<256, 0>  { static constexpr int NS = 2, R0 = 16, R1 = 16, R2 = 1,  EPT = 16; };
template <> struct FftPlan<256, 1>  { static constexpr int NS = 3, R0 = 8,  R1 = 8,  R2 = 4,  EPT = 8;  };
template <> struct FftPlan<512, 0>  { static constexpr int NS = 3, R0 = 8,  R1 = 8,  R2 = 8,  EPT = 8;  };
template <> struct FftPlan<1024, 0> { static constexpr int NS = 3, R0 = 16, R1 = 8,  R2 = 8,  EPT = 16; };
template <> struct FftPlan<2048, 0> { static constexpr int NS = 3, R0 = 16, R1 = 16, R2 = 8,  EPT = 16; };
template <> struct FftPlan<4096, 0> { static constexpr int NS = 3, R0 = 16, R1 = 16, R2 = 16, EPT = 16; };

template <int L, int V = 0> __host__ __device__ constexpr int fft_threads() { return L / FftPlan<L, V>::EPT; }

// padded index into the complex exchange buffer: one extra float2 every 16 spreads the
// power-of-two strides of the Stockham exchanges over the banks (64-bit accesses: 16 lanes per phase)
__host__ __device__ __forceinline__ constexpr int fpad(int i) { return i + (i >> 4); }
// floats in HALF of one transform's exchange buffer (the buffer holds 2 * fft_plane<L>() floats)
template <int L> __host__ __device__ constexpr int fft_plane() { return L + (L >> 4) + 2; }

struct SmemBuf {
    float* re;        // base of the buffer (float2 elements); `im` kept for layout compatibility
    float* im;
    __device__ __forceinline__ float2 get(int i) const { return reinterpret_cast<const float2*>(re)[fpad(i)]; }
    __device__ __forceinline__ void put(int i, float2 v) const { reinterpret_cast<float2*>(re)[fpad(i)] = v; }
};

// ---------------------------------------------------------------------------------- register API
// Thread t of a group owns, before AND after the transform, the elements  idx = t + T*m, m < EPT
// (only their order inside the register array differs), so element-wise work on the spectrum and
// a following inverse transform need no shared-memory exchange at all.
//   input  order: x[b*R0 + r]            holds element  (t + b*T) + r*(L/R0)
//   output order: x[b*RL + r] (RL last radix) holds element  (t + b*T) + r*(L/RL)
template <int L, int V = 0> struct FftIdx {
    using P = FftPlan<L, V>;
    static constexpr int EPT = P::EPT, T = L / P::EPT;
    static constexpr int RL = P::NS == 1 ? P::R0 : (P::NS == 2 ? P::R1 : P::R2);
    __host__ __device__ static constexpr int in(int t, int i) { return t + (i / P::R0) * T + (i % P::R0) * (L / P::R0); }
    __host__ __device__ static constexpr int out(int t, int i) { return t + (i / RL) * T + (i % RL) * (L / RL); }
    // register slot (input order) that holds the element with multiplier m (idx = t + T*m)
    __host__ __device__ static constexpr int in_slot(int m) {
        // m = b + r * (L/R0)/T  with b < EPT/R0
        return (m % (EPT / P::R0)) * P::R0 + m / (EPT / P::R0);
    }
    __host__ __device__ static constexpr int out_slot(int m) { return (m % (EPT / RL)) * RL + m / (EPT / RL); }
};

// w[r] = w1^r for r = 1..R-1 with multiplication depth <= 4 (error ~2e-7)
template <int R>
__device__ __forceinline__ void twiddle_powers(float2 w1, float2 (&w)[16]) {
    w[1] = w1;
    if (R > 2) { w[2] = cmul(w1, w1); w[3] = cmul(w[2], w1); }
    if (R > 4) { w[4] = cmul(w[2], w[2]); w[5] = cmul(w[4], w1); w[6] = cmul(w[4], w[2]); w[7] = cmul(w[4], w[3]); }
    if (R > 8) {
        w[8] = cmul(w[4], w[4]);
#pragma unroll
        for (int r = 1; r < 8; ++r) w[8 + r] = cmul(w[8], w[r]);
    }
}

// per-thread base twiddles exp(-2*pi*i*k/(NSP*R)), k = (t + b*T) % NSP, of the stages after the
// first.  They depend on the thread only, so persistent kernels load them once.
template <int L, int V = 0> struct FftTw {
    using P = FftPlan<L, V>;
    static constexpr int EPT = P::EPT, T = L / P::EPT;
    static constexpr int NB1 = P::NS >= 2 ? EPT / P::R1 : 1;
    static constexpr int NB2 = P::NS >= 3 ? EPT / P::R2 : 1;
    float2 s1[NB1], s2[NB2];
    __device__ __forceinline__ void init(int t) {
        if (P::NS >= 2) {
#pragma unroll
            for (int b = 0; b < NB1; ++b) s1[b] = g_tw[((t + b * T) % P::R0) * (PNP_TW_N / (P::R0 * P::R1))];
        }
        if (P::NS >= 3) {
#pragma unroll
            for (int b = 0; b < NB2; ++b) s2[b] = g_tw[((t + b * T) % (P::R0 * P::R1)) * (PNP_TW_N / (P::R0 * P::R1 * P::R2))];
        }
    }
};

template <int L, int R, int NSP, int NBW, int V = 0>
__device__ __forceinline__ void fft_reg_stage(int t, float2 (&x)[FftPlan<L, V>::EPT], const float2 (&w1)[NBW]) {
    constexpr int EPT = FftPlan<L, V>::EPT;
    constexpr int NB = EPT / R;
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        float2 v[R];
#pragma unroll
        for (int r = 0; r < R; ++r) v[r] = x[b * R + r];
        if (NSP > 1) {
            float2 w[16];
            twiddle_powers<R>(w1[b], w);
#pragma unroll
            for (int r = 1; r < R; ++r) v[r] = cmul(v[r], w[r]);
        }
        Dft<R>::run(v);
#pragma unroll
        for (int r = 0; r < R; ++r) x[b * R + r] = v[r];
    }
}

// exchange between a stage of radix RA (prefix NSP) and the next stage of radix RB
// WS: the T cooperating threads live in ONE warp (T <= 32): warp barriers instead of CTA barriers, so transforms of
// different warps need not run in lock step (csrc/small.cuh)
// barrier among the T threads of ONE transform: bar = 0 is the CTA barrier; bar > 0 a named barrier of `nthreads` threads (the
// transforms of a CTA that holds several -- GP line pairs, NC columns -- then run out of lock step: one group's exchange
// latency overlaps another group's butterflies instead of all 16 warps of the SM waiting at the same point)
#ifndef PNP_GROUP_BAR
#define PNP_GROUP_BAR 1
#endif
__device__ __forceinline__ void fft_sync(int bar, int nthreads) {
    if (bar) asm volatile("bar.sync %0, %1;" ::"r"(bar), "r"(nthreads) : "memory");
    else __syncthreads();
}
// named barrier of group g of a CTA with `groups` transforms of T threads each (0: use the CTA barrier)
template <int T> __device__ __forceinline__ int fft_group_bar(int g, int groups) {
    return (PNP_GROUP_BAR && T >= 64 && T % 32 == 0 && groups > 1 && groups <= 15) ? g + 1 : 0;
}

template <int L, int RA, int NSP, int RB, bool WS = false, int V = 0>
__device__ __forceinline__ void fft_reg_exchange(int t, const SmemBuf& sb, float2 (&x)[FftPlan<L, V>::EPT], int bar = 0) {
    constexpr int EPT = FftPlan<L, V>::EPT;
    constexpr int T = L / EPT;
#pragma unroll
    for (int b = 0; b < EPT / RA; ++b) {
        const int j = t + b * T;
        const int base = (j / NSP) * (NSP * RA) + (j % NSP);
#pragma unroll
        for (int r = 0; r < RA; ++r) sb.put(base + r * NSP, x[b * RA + r]);
    }
    if (WS) __syncwarp(); else fft_sync(bar, T);
#pragma unroll
    for (int b = 0; b < EPT / RB; ++b) {
        const int j = t + b * T;
#pragma unroll
        for (int r = 0; r < RB; ++r) x[b * RB + r] = sb.get(j + r * (L / RB));
    }
}

// In-register forward FFT.  The exchange buffer must be free on entry; on exit the LAST exchange's
// reads may still be in flight in other threads: callers sync before writing the buffer again.
template <int L, bool WS = false, int V = 0>
__device__ __forceinline__ void fft_regs(int t, const SmemBuf& sb, float2 (&x)[FftPlan<L, V>::EPT], const FftTw<L, V>& tw,
                                         int bar = 0) {
    using P = FftPlan<L, V>;
    static_assert(!WS || L / P::EPT <= 32, "warp-synchronous transform: the cooperating threads must fit one warp");
    fft_reg_stage<L, P::R0, 1, FftTw<L, V>::NB1, V>(t, x, tw.s1);
    if constexpr (P::NS >= 2) {
        fft_reg_exchange<L, P::R0, 1, P::R1, WS, V>(t, sb, x, bar);
        fft_reg_stage<L, P::R1, P::R0, FftTw<L, V>::NB1, V>(t, x, tw.s1);
    }
    if constexpr (P::NS >= 3) {
        if (WS) __syncwarp(); else fft_sync(bar, L / P::EPT);   // reads of the first exchange done before the second writes
        fft_reg_exchange<L, P::R1, P::R0, P::R2, WS, V>(t, sb, x, bar);
        fft_reg_stage<L, P::R2, P::R0 * P::R1, FftTw<L, V>::NB2, V>(t, x, tw.s2);
    }
}
template <int L>
__device__ __forceinline__ void fft_regs(int t, const SmemBuf& sb, float2 (&x)[FftPlan<L>::EPT]) {
    FftTw<L> tw;
    tw.init(t);
    fft_regs<L>(t, sb, x, tw);
}

// functor API kept for callers that stream through global / shared memory directly
template <int L, bool LD_SMEM, class Load, class Store>
__device__ __forceinline__ void fft_forward(int t, const SmemBuf& sb, Load& ld, Store& st) {
    constexpr int EPT = FftPlan<L>::EPT;
    float2 x[EPT];
#pragma unroll
    for (int i = 0; i < EPT; ++i) x[i] = ld(FftIdx<L>::in(t, i));
    if (LD_SMEM) __syncthreads();                // loader read the exchange buffer
    fft_regs<L>(t, sb, x);
#pragma unroll
    for (int i = 0; i < EPT; ++i) st(FftIdx<L>::out(t, i), x[i]);
}

}  // namespace pnp
