// CSMRI data-fidelity gradient  Re(ifft2(sel o fft2(x) - Y_sel))  fused with the
// variance-reduced update, for sm_100a.
//
// Reference: problems/CSMRI.py:76-81 (grad_full), :83-89 (grad_stoch); update lines
// algorithms/pnp_gd.py:32-35, pnp_sgd.py:32-36, pnp_svrg.py:53-57, pnp_sarah.py:72-75.
//
// DEVICE LAYOUT.  Images are kept TRANSPOSED: line c (c = original column, W lines) holds
// the H samples z[0..H-1][c] contiguously.  The per-column operators of the prox step
// (estimate_sigma / denoise_wavelet with multichannel=True) therefore work on contiguous
// lines, and fft2 of the transpose is the transpose of fft2.
//
//   pass 1  k_lines_r2c   two real lines -> one complex FFT of length H (two-for-one), unpacked
//                         to the packed half spectrum, stored TRANSPOSED as S[kyp][c], kyp in
//                         [0, H/2), c = line: the column pass then reads contiguous rows.  The real
//                         Nyquist term (ky = H/2) rides in the imaginary slot of the real DC term.
//                         Input may be a difference a - b (g_B(z) - g_B(w) is linear, Y cancels).
//   pass 2  k_cols_mask   per packed ky column: FFT over the W lines, multiply by the
//                         Hermitian-symmetrised selection  (sel[k] + sel[-k]) / 2, subtract the
//                         symmetrised measurements, inverse FFT -- the spectrum never leaves
//                         shared memory between the three steps.
//   pass 3  k_lines_c2r   two-for-one inverse, 1/(HW) scaling and the fused update epilogue
//                         v = g * gscale + vadd ;  z_out = z_in - step * v.
//
// Because sel is not Hermitian-symmetric, Re(ifft2(sel o Z)) is computed exactly as the
// real inverse of its Hermitian part:  ((sel[k] + sel[-k]) / 2) Z[k]  (Z[-k] = conj Z[k]).
#pragma once
#include "fft_core.cuh"
#include "prox.cuh"

namespace pnp {

struct GradEpilogue {
    float gscale;                 // multiplies the raw gradient (after 1/(HW))
    const float* gscale_ptr;      // optional per-image multiplier, replaces gscale
    float step;                   // z_out = z_in - step * v
    const float* step_ptr;        // optional per-image step, replaces step
    float* g_out;                 // optional: g * gscale
    const float* vadd;            // optional: v = g * gscale + vadd
    float* v_out;                 // optional: v
    const float* z_in;            // optional (with z_out)
    float* z_out;
};

// ------------------------------------------------------------------ selection bits
__device__ __forceinline__ void or_byte(unsigned char* base, long long byte_idx, unsigned v) {
    unsigned* w = reinterpret_cast<unsigned*>(base) + (byte_idx >> 2);
    atomicOr(w, v << (8 * (int)(byte_idx & 3)));
}

// k = ky * W + kx is an index into the reference's (H, W) k-space array (problems/CSMRI.py:66-74)
__device__ __forceinline__ void set_sel_bits(unsigned char* bits, int H, int W, int k) {
    const int ky = k / W, kx = k % W;
    const int hp = H / 2;
    const int kym = (H - ky) % H, kxm = (W - kx) % W;
    if (ky < hp) or_byte(bits, (long long)ky * W + kx, 1u);
    else if (ky == hp) or_byte(bits, kx, 4u);
    if (kym < hp) or_byte(bits, (long long)kym * W + kxm, 2u);
    else if (kym == hp) or_byte(bits, kxm, 8u);
}

// explicit minibatch:  idx[img][cursor][0..B)
// device-drawn minibatch: B distinct positions of the sampled support through a keyed
// cycle-walking Feistel permutation of [0, M0)  (bench / production mode; the reference draws
// with np.random.choice(..., replace=False), problems/CSMRI.py:72)
__device__ __forceinline__ unsigned mix32(unsigned x) {
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return x;
}
// Generalised (unbalanced) Feistel network on Z_b x Z_a with a = 2^hb >= sqrt(n) and b = ceil(n / a): the domain
// a*b exceeds n by less than a, so cycle walking almost never iterates (a balanced 2^(2 hb) domain needs up to 4
// passes per index).  Rounds alternate (l, r) -> (r, (l + F(r)) mod b) and (l, r) -> (r, (l + F(r)) mod a); the
// reduction of F to [0, b) is a multiply-high, so there is no division.  Host twins: host_sampler.cpp and
// engine.feistel_sample (NumPy) produce the same sequence bit for bit.
struct FeistelDom { int hb; unsigned hm, b; };          // domain of the permutation of [0, n): depends on n only
__device__ __forceinline__ FeistelDom feistel_domain(unsigned n) {
    int hb = 1;
    while ((1u << (2 * hb)) < n) ++hb;                 // a = 2^hb, a*a >= n
    const unsigned hm = (1u << hb) - 1u;
    return FeistelDom{hb, hm, (n + hm) >> hb};         // b = ceil(n / a) <= a
}
// one pass of the network over the domain a*b (a bijection of [0, a*b))
__device__ __forceinline__ unsigned feistel_pass(unsigned x, unsigned key, const FeistelDom& d) {
    const int hb = d.hb;
    const unsigned hm = d.hm, b = d.b;
    unsigned l = x >> hb, r = x & hm;                  // l in Z_b, r in Z_a
#pragma unroll
    for (int rd = 0; rd < 4; rd += 2) {
        const unsigned f0 = __umulhi(mix32(r ^ (key + 0x9e3779b9U * (rd + 1))), b);
        unsigned t = l + f0;                            // < 2b
        t = t >= b ? t - b : t;
        l = r;                                          // (l, r) now in Z_a x Z_b
        r = t;
        const unsigned f1 = mix32(r ^ (key + 0x9e3779b9U * (rd + 2))) & hm;
        t = (l + f1) & hm;
        l = r;                                          // back in Z_b x Z_a
        r = t;
    }
    return (l << hb) | r;
}
__device__ __forceinline__ unsigned feistel_perm(unsigned i, unsigned n, unsigned key, const FeistelDom& d) {
    unsigned x = i;
    do { x = feistel_pass(x, key, d); } while (x >= n);          // cycle walking
    return x;
}
__device__ __forceinline__ unsigned feistel_perm(unsigned i, unsigned n, unsigned key) {
    return feistel_perm(i, n, key, feistel_domain(n));
}


// Minibatch selection done INSIDE pass 1 (all threads of the persistent grid, while the first lines are still on
// their way from HBM): the stand-alone selection kernels cannot share an SM with a pass that holds every register,
// so on a parallel graph branch they only started when the first line CTAs retired and the column pass waited
// for them (6 us of an 82 us iteration).  Explicit indices idx[img][*cursor][0..count) when idx != null, else the
// keyed Feistel draw of `count` positions of support[img][0..m0[img]).  `bits` must be zero on entry.
struct SelJob {
    unsigned char* bits;          // null: no selection in this pass
    const int* idx;
    long long idx_img_stride;
    const int* cursor;
    const int* support;
    const int* m0;
    long long support_img_stride;
    int count;
    unsigned seed;
    const int* counter;
    int* idx_out;                 // optional [batch][count]: the positions drawn by the sampler
    int counter_add;              // added to *counter (a pass that draws the minibatch of the NEXT iteration before the counters advance)
};

__device__ __forceinline__ void run_sel_job(const SelJob& sj, int H, int W, int img, int tid, int nthreads) {
    unsigned char* bi = sj.bits + (long long)img * W * (H / 2);
    if (sj.idx) {
        const int* src = sj.idx + (long long)img * sj.idx_img_stride + (long long)(sj.cursor ? *sj.cursor : 0) * sj.count;
        for (int i = tid; i < sj.count; i += nthreads) set_sel_bits(bi, H, W, src[i]);
    } else {
        const unsigned key = mix32(sj.seed ^ mix32(((sj.counter ? (unsigned)*sj.counter : 0u) + (unsigned)sj.counter_add) * 0x632be5abU + (unsigned)img));
        const int* sup = sj.support + (long long)img * sj.support_img_stride;
        const unsigned n = (unsigned)sj.m0[img];
        for (int i = tid; i < sj.count; i += nthreads) {
            if ((unsigned)i >= n) break;              // count <= m0 is the caller's contract; never walk outside the domain
            const int k = sup[feistel_perm((unsigned)i, n, key)];
            if (sj.idx_out) sj.idx_out[(long long)img * sj.count + i] = k;
            set_sel_bits(bi, H, W, k);
        }
    }
}

// ------------------------------------------------------------------ pass 1
// group stride (floats) of the per-pair exchange buffers: 2 planes, skewed so that GP groups read
// at the same in-plane index hit different banks
template <int L, int GP> __host__ __device__ constexpr int group_stride() {
    return 2 * fft_plane<L>() + ((GP > 1 ? 32 / GP : 0) - (2 * fft_plane<L>()) % 32 + 32) % 32;
}
// floats of exchange planes in front of the TMA staging buffers (kept 128-byte aligned)
template <int L, int GP> __host__ __device__ constexpr int lines_stage_off() { return (GP * group_stride<L, GP>() + 31) & ~31; }

// floats between the staged line pairs of an item: with fewer than 32 threads per transform a warp holds several
// transforms, and their pairs, 2 L floats apart, would sit in the same banks (ncu: every excess shared-memory wavefront
// of the 256-point line passes was a read of the staged lines, 2-way).  T floats of padding put the pairs of a warp's
// transforms side by side in the banks; one bulk copy per pair then instead of one per item.
// resident CTAs per SM the passes are compiled for: 16 warps per SM whatever the CTA size (2 x 256, 4 x 128, 8 x 64 threads).
// CTAs of 64 threads (L = 256: 4 transforms of 16 threads) used to be compiled for 4 per SM: the inverse line pass took 198
// registers and ran with 6.5 warps per SM in the batched sweeps (ncu, profiles/README.md).
#ifndef PNP_SMALL_CTA_MINB
#define PNP_SMALL_CTA_MINB 8
#endif
__host__ __device__ constexpr int pass_min_ctas(int threads) { return threads >= 256 ? 2 : (threads >= 128 ? 4 : PNP_SMALL_CTA_MINB); }
template <int L> __host__ __device__ constexpr int stage_pad() {
    return fft_threads<L>() >= 32 ? 0 : (fft_threads<L>() < 4 ? 4 : fft_threads<L>());
}
template <int L> __host__ __device__ constexpr int stage_pair_stride() { return 2 * L + stage_pad<L>(); }

// Persistent: CTA b processes items b, b + gridDim.x, ... (an item = GP line pairs).  The 2*GP input
// lines of an item are contiguous in memory: ONE thread stages them into shared memory with TMA bulk
// copies (cp.async.bulk + mbarrier); the copy of item i+1 is in flight while item i is transformed.
// smem: [GP groups x exchange planes][stage a: 2*GP*L floats][stage b: 2*GP*L floats]
template <int L, int GP>
__global__ void __launch_bounds__(GP * (L / FftPlan<L>::EPT), pass_min_ctas(GP * (L / FftPlan<L>::EPT)))
k_lines_r2c(const float* __restrict__ a, const float* __restrict__ b, float2* __restrict__ S,
            int nlines, long long img_stride, SelJob sj, int row_lo, int row_hi) {
    constexpr int T = fft_threads<L>();
    constexpr int EPT = FftPlan<L>::EPT;
    constexpr int PL = fft_plane<L>();
    constexpr int GS = group_stride<L, GP>();
    extern __shared__ __align__(128) float smem[];
    __shared__ __align__(8) unsigned long long bar;
    constexpr int PSTR = stage_pair_stride<L>();
    float* stage_a = smem + lines_stage_off<L, GP>();
    float* stage_b = stage_a + GP * PSTR;
    const int g = threadIdx.x / T, t = threadIdx.x % T;
    const int gbar = fft_group_bar<T>(g, GP);        // the barriers inside a transform involve its T threads only
    const int npairs = nlines >> 1;
    const int items = (npairs + GP - 1) / GP;
    const long long ibase = (long long)blockIdx.y * img_stride;
    const SmemBuf sb{smem + g * GS, smem + g * GS + PL};
    const int W2 = nlines >> 1;                     // float4 per spectrum row

    auto issue = [&](int item) {                    // one thread: stage the item's lines
        int np = npairs - item * GP;
        np = np < GP ? np : GP;
        const unsigned bytes = (unsigned)(np * 2 * L * sizeof(float));
        mbar_expect_tx(&bar, b ? 2 * bytes : bytes);
        const long long off = ibase + (long long)(2 * item * GP) * L;
        if (stage_pad<L>() == 0) {
            bulk_g2s_keep(stage_a, a + off, bytes, &bar);     // the iterate: read again by the tail
            if (b) bulk_g2s(stage_b, b + off, bytes, &bar);
        } else {
            for (int p = 0; p < np; ++p) {                    // padded pairs: one copy each
                bulk_g2s_keep(stage_a + p * PSTR, a + off + (long long)p * 2 * L, (unsigned)(2 * L * sizeof(float)), &bar);
                if (b) bulk_g2s(stage_b + p * PSTR, b + off + (long long)p * 2 * L, (unsigned)(2 * L * sizeof(float)), &bar);
            }
        }
    };

    trace(100, true);
    if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    FftTw<L> tw;
    tw.init(t);
    __syncthreads();
    griddep_wait();                                 // the iterate comes from the previous kernel of the chain
    griddep_launch();
    int item = blockIdx.x;
    if (threadIdx.x == 0 && item < items) issue(item);
    // In-pass minibatch selection: when the items do not divide evenly, the CTAs with one item fewer do it AFTER their
    // loop (they would idle there anyway; at the start it sat in front of everybody's first transform: its dependent
    // global loads cost 3 us).  Otherwise every CTA takes a slice before its first item.
    const int n_heavy = items % (int)gridDim.x;                  // CTAs [0, n_heavy) have one item more
    const bool sel_late = sj.bits && n_heavy != 0;
    if (sj.bits && !sel_late) run_sel_job(sj, L, nlines, blockIdx.y, blockIdx.x * blockDim.x + threadIdx.x, gridDim.x * blockDim.x);
    unsigned parity = 0;
    for (; item < items; item += gridDim.x) {
        float2 x[EPT];
        mbar_wait(&bar, parity);
        trace(101);
        parity ^= 1;
        if (item * GP + g < npairs) {
            const float* la = stage_a + g * PSTR;
            const float* lb = stage_b + g * PSTR;
#pragma unroll
            for (int i = 0; i < EPT; ++i) {
                const int idx = FftIdx<L>::in(t, i);
                float re = la[idx], im = la[idx + L];
                if (b) { re -= lb[idx]; im -= lb[idx + L]; }
                x[i] = make_float2(re, im);
            }
        } else {
#pragma unroll
            for (int i = 0; i < EPT; ++i) x[i] = make_float2(0.f, 0.f);
        }
        __syncthreads();                            // staging consumed -> refill it for the next item
        if (threadIdx.x == 0 && item + (int)gridDim.x < items) issue(item + gridDim.x);
        fft_regs<L>(t, sb, x, tw, gbar);
        if (FftPlan<L>::NS > 1) fft_sync(gbar, T);
#pragma unroll
        for (int i = 0; i < EPT; ++i) sb.put(FftIdx<L>::out(t, i), x[i]);
        __syncthreads();
        // unpack the two real transforms; S[kyp][c]: the GP pairs of this CTA write 2*GP adjacent
        // complex values (16 bytes per pair) of row kyp
        const int pair0 = item * GP;
        float4* S4 = reinterpret_cast<float4*>(S + (ibase >> 1)) + pair0;
        for (int i = threadIdx.x; i < GP * (L / 2); i += GP * T) {
            const int gg = i % GP, k = i / GP;
            if (pair0 + gg >= npairs) continue;
            if (k < row_lo || k >= row_hi) continue;               // measurement shard: only the rank's packed ky rows are used
            const SmemBuf sg{smem + gg * GS, smem + gg * GS + PL};
            const float2 xk = sg.get(k);
            const float2 xm = sg.get(k == 0 ? L / 2 : L - k);
            float4 o;
            if (k == 0) {
                o = make_float4(xk.x, xm.x, xk.y, xm.y);       // A = (DC, Nyquist) line 2p ; B likewise line 2p+1
            } else {
                o = make_float4(0.5f * (xk.x + xm.x), 0.5f * (xk.y - xm.y), 0.5f * (xk.y + xm.y), 0.5f * (xm.x - xk.x));
            }
            stg_keep(S4 + (long long)k * W2 + gg, o);
        }
        __syncthreads();
        trace(102);
    }
    if (sel_late && (int)blockIdx.x >= n_heavy)
        run_sel_job(sj, L, nlines, blockIdx.y, ((int)blockIdx.x - n_heavy) * blockDim.x + threadIdx.x, ((int)gridDim.x - n_heavy) * blockDim.x);
    trace(109);
    trace_flush();
}

// ------------------------------------------------------------------ pass 2
// sel bits per packed entry [kyp][kx]:  bit0 = sel[kyp][kx], bit1 = sel[-kyp][-kx];
// for kyp == 0 additionally bit2 = sel[H/2][kx], bit3 = sel[H/2][-kx] (the packed Nyquist row).
__device__ __forceinline__ float2 apply_sel(float2 F, unsigned bb, bool use_y, const float2* y1,
                                            const float2* y2) {
    const float w = 0.5f * (float)((bb & 1u) + ((bb >> 1) & 1u));
    float2 o = make_float2(w * F.x, w * F.y);
    if (use_y) {
        if (bb & 1u) { const float2 y = *y1; o.x -= 0.5f * y.x; o.y -= 0.5f * y.y; }
        if (bb & 2u) { const float2 y = *y2; o.x -= 0.5f * y.x; o.y -= 0.5f * y.y; }
    }
    return o;
}

template <int L, int NC> __host__ __device__ constexpr int cols_stage_off() { return (NC * 2 * fft_plane<L>() + 31) & ~31; }

// Persistent: one group of T threads per packed column (= contiguous row S[kyp][0..L)), NC groups
// per CTA; the NC rows of an item are one contiguous block, staged by a TMA bulk copy that runs
// one item ahead.  Columns 1..hp-1 are element-wise in the spectrum: forward FFT, selection and
// inverse FFT happen in registers + the exchange buffer.  Column 0 (DC + i*Nyquist packed) needs
// C[kx] and C[-kx] together and is done by the last CTA after its loop.
#ifndef PNP_COLS_MINB
#define PNP_COLS_MINB 4
#endif
template <int L, int NC>
__global__ void __launch_bounds__(NC * (L / FftPlan<L>::EPT), (NC * (L / FftPlan<L>::EPT) >= 128 ? (NC * (L / FftPlan<L>::EPT) >= 256 ? 2 : PNP_COLS_MINB) : pass_min_ctas(NC * (L / FftPlan<L>::EPT))))
k_cols_mask(float2* __restrict__ S, const unsigned char* __restrict__ bits,
            const float2* __restrict__ Y1, const float2* __restrict__ Y2,
            const float2* __restrict__ Y1n, const float2* __restrict__ Y2n,
            int hp, long long bits_img_stride, long long y_img_stride, unsigned char* __restrict__ clear_bits,
            int row_lo, int row_hi) {
    constexpr int T = fft_threads<L>();
    constexpr int EPT = FftPlan<L>::EPT;
    constexpr int PL = fft_plane<L>();
    using IX = FftIdx<L>;
    extern __shared__ __align__(128) float smem[];
    __shared__ __align__(8) unsigned long long bar;
    const int g = threadIdx.x / T, t = threadIdx.x % T;
    const int gbar = fft_group_bar<T>(g, NC);        // the barriers inside a transform involve its T threads only
    const int img = blockIdx.y;
    const SmemBuf sb{smem + g * 2 * PL, smem + g * 2 * PL + PL};
    float2* stage = reinterpret_cast<float2*>(smem + cols_stage_off<L, NC>());      // NC columns of L complex
    const unsigned char* stage_bits = reinterpret_cast<const unsigned char*>(stage + NC * L);   // + NC rows of L bytes
    float2* Si = S + (long long)img * hp * L;
    const unsigned char* bi = bits + (long long)img * bits_img_stride;
    const bool use_y = Y1 != nullptr;
    const float2* y1i = use_y ? Y1 + (long long)img * y_img_stride : nullptr;
    const float2* y2i = use_y ? Y2 + (long long)img * y_img_stride : nullptr;
    // packed rows [c_lo, c_hi) (a measurement shard owns a band of them; everything for an unsharded problem)
    const int c_lo = row_lo > 1 ? row_lo : 1, c_hi = row_hi < hp ? row_hi : hp;
    const int items = c_hi > c_lo ? (c_hi - c_lo + NC - 1) / NC : 0;

    auto issue = [&](int item) {
        int nc = c_hi - c_lo - item * NC;
        nc = nc < NC ? nc : NC;
        const unsigned bytes = (unsigned)(nc * L * sizeof(float2));
        mbar_expect_tx(&bar, bytes + (unsigned)(nc * L));
        bulk_g2s_keep(stage, Si + (long long)(c_lo + item * NC) * L, bytes, &bar);
        bulk_g2s(const_cast<unsigned char*>(stage_bits), bi + (long long)(c_lo + item * NC) * L, (unsigned)(nc * L), &bar);
    };
    trace(200, true);
    if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    FftTw<L> tw;
    tw.init(t);
    __syncthreads();
    griddep_wait();
    griddep_launch();
    // CTA 0 of the grid owns packed column 0 (two transforms through a scalar split: as an appendix of an item CTA
    // it ended 2 us after everybody else); the others loop over the items
    const int nct = (int)gridDim.x - 1;
    int item = blockIdx.x > 0 ? (int)blockIdx.x - 1 : items;
    if (threadIdx.x == 0 && item < items) issue(item);
    unsigned parity = 0;
    for (; item < items; item += nct) {
        const int col = c_lo + item * NC + g;
        const bool active = col < c_hi;
        const long long crow = (long long)(active ? col : 0) * L;
        unsigned long long bbp = 0ull;                                      // 4 selection bits per element
        float2 x[EPT];
        mbar_wait(&bar, parity);
        trace(201);
        parity ^= 1;
        if (active) {
#pragma unroll
            for (int i = 0; i < EPT; ++i) x[i] = stage[g * L + IX::in(t, i)];
#pragma unroll
            for (int m = 0; m < EPT; ++m) bbp |= (unsigned long long)(stage_bits[g * L + t + T * m] & 0xFu) << (4 * m);
        } else {
#pragma unroll
            for (int i = 0; i < EPT; ++i) x[i] = make_float2(0.f, 0.f);
        }
        __syncthreads();                            // staging consumed -> refill it for the next item
        if (threadIdx.x == 0 && item + nct < items) issue(item + nct);
        if (clear_bits && active) {                 // minibatch selection is single use: leave the row zeroed
            unsigned char* cb = clear_bits + (long long)img * bits_img_stride + crow;
            for (int i = t; i < L / 16; i += T) reinterpret_cast<uint4*>(cb)[i] = make_uint4(0u, 0u, 0u, 0u);
        }
        fft_regs<L>(t, sb, x, tw, gbar);
        // selection in registers, then reorder (same elements t + T*m) into the inverse's input order
        float2 y[EPT];
#pragma unroll
        for (int m = 0; m < EPT; ++m) {
            const int kx = t + T * m;
            const float2 o = apply_sel(x[IX::out_slot(m)], (unsigned)(bbp >> (4 * m)) & 0xFu, use_y, y1i + crow + kx,
                                       y2i + crow + kx);
            y[IX::in_slot(m)] = cswap(o);
        }
        if (FftPlan<L>::NS > 1) fft_sync(gbar, T);
        fft_regs<L>(t, sb, y, tw, gbar);
        if (active) {
            float2* Sc = Si + crow;
#pragma unroll
            for (int i = 0; i < EPT; ++i) stg_keep(Sc + IX::out(t, i), cswap(y[i]));
        }
        if (FftPlan<L>::NS > 1) __syncthreads();
        trace(202);
    }
    trace(209);

    if (blockIdx.x != 0 || row_lo > 0) { trace_flush(); return; }
    // ---- packed column 0: C = FFT(DC + i * Nyq); split, select each row, re-pack ----
    {
        float2 x[EPT];
        if (g == 0) {
#pragma unroll
            for (int i = 0; i < EPT; ++i) x[i] = Si[IX::in(t, i)];
        } else {
#pragma unroll
            for (int i = 0; i < EPT; ++i) x[i] = make_float2(0.f, 0.f);
        }
        fft_regs<L>(t, sb, x);
        if (FftPlan<L>::NS > 1) __syncthreads();
#pragma unroll
        for (int i = 0; i < EPT; ++i) sb.put(IX::out(t, i), x[i]);
        __syncthreads();
        const float2* y1n = use_y ? Y1n + (long long)img * L : nullptr;
        const float2* y2n = use_y ? Y2n + (long long)img * L : nullptr;
        if (g == 0) {
            for (int kx = t; kx <= L / 2; kx += T) {
                const int km = (L - kx) % L;
                const float2 ck = sb.get(kx), cm = sb.get(km);
                const float2 fdc = make_float2(0.5f * (ck.x + cm.x), 0.5f * (ck.y - cm.y));
                const float2 fny = make_float2(0.5f * (ck.y + cm.y), 0.5f * (cm.x - ck.x));
                const unsigned bk = bi[kx], bm = bi[km];
                const float2 dk = apply_sel(fdc, bk, use_y, y1i + kx, y2i + kx);
                const float2 nk = apply_sel(fny, bk >> 2, use_y, y1n + kx, y2n + kx);
                const float2 dm = apply_sel(make_float2(fdc.x, -fdc.y), bm, use_y, y1i + km, y2i + km);
                const float2 nm = apply_sel(make_float2(fny.x, -fny.y), bm >> 2, use_y, y1n + km, y2n + km);
                // C'[k] = dc[k] + i * ny[k], stored re/im swapped for the inverse transform
                sb.put(kx, make_float2(dk.y + nk.x, dk.x - nk.y));
                if (km != kx) sb.put(km, make_float2(dm.y + nm.x, dm.x - nm.y));
            }
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < EPT; ++i) x[i] = sb.get(IX::in(t, i));
        __syncthreads();
        fft_regs<L>(t, sb, x);
        if (g == 0) {
#pragma unroll
            for (int i = 0; i < EPT; ++i) stg_keep(Si + IX::out(t, i), cswap(x[i]));
            if (clear_bits) {
                unsigned char* cb = clear_bits + (long long)img * bits_img_stride;
                for (int i = t; i < L / 16; i += T) reinterpret_cast<uint4*>(cb)[i] = make_uint4(0u, 0u, 0u, 0u);
            }
        }
    }
    trace(210);
    trace_flush();
}

// ------------------------------------------------------------------ pass 3
// Persistent like pass 1.  The spectrum entries of the next item are prefetched into registers (they
// are strided: 16*GP bytes per spectrum row); the epilogue operands (vadd and z_in lines, contiguous)
// are staged by TMA bulk copies issued at the start of the item, so they arrive while the inverse FFT
// runs and the epilogue never waits on a global load.
// UPD = true is the inner-iteration form (vadd, z_in and z_out given, nothing else written): the generic
// epilogue with its five optional pointers unrolls to >100 KiB of code and stalls on instruction fetch.
template <int L, int GP, bool UPD>
__global__ void __launch_bounds__(GP * (L / FftPlan<L>::EPT), pass_min_ctas(GP * (L / FftPlan<L>::EPT)))
k_lines_c2r(const float2* __restrict__ S, int nlines, long long img_stride, float inv_n, GradEpilogue ep, int row_lo, int row_hi) {
    constexpr int T = fft_threads<L>();
    constexpr int EPT = FftPlan<L>::EPT;
    constexpr int PL = fft_plane<L>();
    constexpr int GS = group_stride<L, GP>();
    constexpr int NQ = (L / 2) / T;                 // float4 spectrum entries per thread per item
    using IX = FftIdx<L>;
    extern __shared__ __align__(128) float smem[];
    __shared__ __align__(8) unsigned long long bar;
    constexpr int PSTR = stage_pair_stride<L>();
    float* stage_v = smem + lines_stage_off<L, GP>();      // vadd lines of the item
    float* stage_z = stage_v + GP * PSTR;                   // z_in lines of the item
    const int g = threadIdx.x / T, t = threadIdx.x % T;
    const int gbar = fft_group_bar<T>(g, GP);        // the barriers inside a transform involve its T threads only
    const int img = blockIdx.y;
    const int npairs = nlines >> 1;
    const int items = (npairs + GP - 1) / GP;
    const long long ibase = (long long)img * img_stride;
    const SmemBuf sb{smem + g * GS, smem + g * GS + PL};
    const int W2 = nlines >> 1;
    const float4* S4 = reinterpret_cast<const float4*>(S + (ibase >> 1));
    const float gs = inv_n * (ep.gscale_ptr ? ep.gscale_ptr[img] : ep.gscale);
    const float step = ep.step_ptr ? ep.step_ptr[img] : ep.step;
    const float* p_vadd = ep.vadd;
    const float* p_zin = ep.z_in;
    float* p_gout = UPD ? nullptr : ep.g_out;
    float* p_vout = UPD ? nullptr : ep.v_out;
    float* p_zout = ep.z_out;
    const bool staged = UPD || (p_vadd != nullptr) || (p_zout != nullptr);

    auto load_spec = [&](int item, float4 (&q)[NQ]) {
#pragma unroll
        for (int n = 0; n < NQ; ++n) {
            const int i = threadIdx.x + n * GP * T;
            const int gg = i % GP, k = i / GP;
            q[n] = (S && item * GP + gg < npairs && k >= row_lo && k < row_hi) ? ldg_stream(S4 + (long long)k * W2 + item * GP + gg)
                                                                               : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    };
    auto issue = [&](int item) {                    // one thread: TMA bulk copies of the epilogue operands
        int np = npairs - item * GP;
        np = np < GP ? np : GP;
        const unsigned bytes = (unsigned)(np * 2 * L * sizeof(float));
        const long long off = ibase + (long long)(2 * item * GP) * L;
        mbar_expect_tx(&bar, ((UPD || p_vadd) ? bytes : 0u) + ((UPD || p_zout) ? bytes : 0u));
        if (stage_pad<L>() == 0) {
            if (UPD || p_vadd) bulk_g2s(stage_v, p_vadd + off, bytes, &bar);
            if (UPD || p_zout) bulk_g2s(stage_z, p_zin + off, bytes, &bar);
        } else {
            for (int p = 0; p < np; ++p) {                    // padded pairs: one copy each
                if (UPD || p_vadd) bulk_g2s(stage_v + p * PSTR, p_vadd + off + (long long)p * 2 * L, (unsigned)(2 * L * sizeof(float)), &bar);
                if (UPD || p_zout) bulk_g2s(stage_z + p * PSTR, p_zin + off + (long long)p * 2 * L, (unsigned)(2 * L * sizeof(float)), &bar);
            }
        }
    };
    if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    __syncthreads();
    griddep_wait();
    griddep_launch();
    unsigned parity = 0;

    float4 qn[NQ];
    int item = blockIdx.x;
    if (item < items) load_spec(item, qn);
    FftTw<L> tw;
    tw.init(t);
    for (; item < items; item += gridDim.x) {
        if (staged && threadIdx.x == 0) issue(item);
        float2 x[EPT];
        if (UPD && !S) {
            // no spectrum (inner-iteration form only): the gradient term is exactly zero -- first inner iteration of an
            // SVRG epoch, z == w bit for bit, so g_B(z) - g_B(w) = 0 and v = vadd (k_update_prox does the same); the
            // transform of zeros is skipped, the epilogue below is unchanged
#pragma unroll
            for (int i = 0; i < EPT; ++i) x[i] = make_float2(0.f, 0.f);
        } else {
            // X[k] = A[k] + i B[k] of the two lines, written re/im swapped for the inverse transform
#pragma unroll
            for (int n = 0; n < NQ; ++n) {
                const int i = threadIdx.x + n * GP * T;
                const int gg = i % GP, k = i / GP;
                const SmemBuf sg{smem + gg * GS, smem + gg * GS + PL};
                const float4 q = qn[n];                                // A = (q.x, q.y) line 2p ; B = (q.z, q.w) line 2p+1
                if (k == 0) {
                    sg.put(0, make_float2(q.z, q.x));                  // X[0]   = A_dc + i B_dc   (swapped)
                    sg.put(L / 2, make_float2(q.w, q.y));              // X[L/2] = A_ny + i B_ny   (swapped)
                } else {
                    sg.put(k, make_float2(q.y + q.z, q.x - q.w));      // X[k]   = A + iB
                    sg.put(L - k, make_float2(q.z - q.y, q.x + q.w));  // X[L-k] = conj A + i conj B
                }
            }
            if (item + (int)gridDim.x < items) load_spec(item + gridDim.x, qn);
            __syncthreads();
#pragma unroll
            for (int i = 0; i < EPT; ++i) x[i] = sb.get(IX::in(t, i));
            fft_sync(gbar, T);
            fft_regs<L>(t, sb, x, tw, gbar);
        }
        if (staged) mbar_wait(&bar, parity);
        parity ^= 1;
        const int pair = item * GP + g;
        if (pair < npairs) {
            const long long base = ibase + (long long)(2 * pair) * L;
            const float* sv = stage_v + g * PSTR;
            const float* sz = stage_z + g * PSTR;
#pragma unroll
            for (int i = 0; i < EPT; ++i) {
                const int idx = IX::out(t, i);
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    // swapped output: .y = real part -> line 2*pair, .x = imag part -> line 2*pair + 1
                    const float gval = (h == 0 ? x[i].y : x[i].x) * gs;
                    const int sh = idx + (h ? L : 0);
                    const long long eh = base + sh;
                    if (UPD) {
                        stg_stream(p_zout + eh, sz[sh] - step * (gval + sv[sh]));
                    } else {
                        if (p_gout) stg_stream(p_gout + eh, gval);
                        float v = gval;
                        if (p_vadd) v += sv[sh];
                        if (p_vout) stg_stream(p_vout + eh, v);
                        if (p_zout) stg_stream(p_zout + eh, sz[sh] - step * v);
                    }
                }
            }
        }
        __syncthreads();                            // exchange and staging buffers free for the next item
    }
}

// ------------------------------------------------------------------ pass 3 + prox in one cooperative launch
// Inner-iteration tail  z <- Denoise(z - step * (g * gscale + vadd))  with the wavelet ("TV") prox, for images
// whose lines fit the SMs' shared memory (2048^2: 14 lines = 112 KiB per CTA).  Every CTA owns a contiguous
// block of line pairs.  Its z_in lines are staged ONCE by TMA bulk copies straight into a resident buffer; the
// inverse line transforms (GP pairs per round, same two-for-one scheme as k_lines_c2r) update that buffer in
// place; the sigma estimate, the grid-wide mean and the BayesShrink then run on the resident lines
// (prox_phase_* of prox.cuh) and only the denoised iterate is written.  Against k_lines_c2r +
// k_prox_wavelet_fused this drops one write and one read of the iterate and a launch boundary.
template <int L> __host__ __device__ constexpr int upd_gp() { return 512 / fft_threads<L>(); }
__device__ unsigned long long g_upd_phase_ns[8];      // PNP_PHASE_TIMING builds only: %globaltimer at the phase boundaries of CTA 0
__device__ __forceinline__ void upd_mark(int i) {
    trace(300 + i, i == 0);
#ifdef PNP_PHASE_TIMING
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        g_upd_phase_ns[i] = t;
    }
#endif
}

template <int L, bool NEXT = false>
__global__ void __launch_bounds__(512, 1)
k_update_prox(const float2* __restrict__ S, int nlines, float inv_n, float gscale, float step,
              const float* __restrict__ step_ptr, const float* __restrict__ vadd, const float* __restrict__ z_in,
              float* __restrict__ z_out, const float* __restrict__ xrec, int pairs_per_cta, float sigma_modifier,
              float fallback_sigma, double* __restrict__ sig_log, double* __restrict__ mse_log, const int* __restrict__ slot,
              int* __restrict__ advance, int n_advance, unsigned* __restrict__ gbar,
              const float* __restrict__ w_next, float2* __restrict__ S_next, SelJob sj_next) {
    constexpr int T = fft_threads<L>();
    constexpr int EPT = FftPlan<L>::EPT;
    constexpr int PL = fft_plane<L>();
    constexpr int GP = upd_gp<L>();
    constexpr int GS = group_stride<L, GP>();
    constexpr int NQ = (L / 2) / T;
    using IX = FftIdx<L>;
    extern __shared__ __align__(128) float smem[];
    __shared__ __align__(8) unsigned long long bars[4];         // z_in lines of round 0 / of the later rounds / ground truth
    float* lines = smem + lines_stage_off<L, GP>();             // resident: 2 * pairs_per_cta lines
    const int g = threadIdx.x / T, t = threadIdx.x % T;
    const int fbar = fft_group_bar<T>(g, GP);        // the barriers inside a transform involve its T threads only
    const int npairs = nlines >> 1;
    const int first_pair = blockIdx.x * pairs_per_cta;
    int mine = npairs - first_pair;
    mine = mine < 0 ? 0 : (mine > pairs_per_cta ? pairs_per_cta : mine);
    const long long first = 2ll * first_pair;                   // first line of this CTA
    const SmemBuf sb{smem + g * GS, smem + g * GS + PL};
    const float4* S4 = reinterpret_cast<const float4*>(S);
    const float gs = inv_n * gscale;

    upd_mark(0);
    if (threadIdx.x == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        mbar_init(&bars[2], 1);
        mbar_fence_init();
        if (mine > 0) {
            // the first round of transforms only waits for ITS lines (all CTAs start together: the whole iterate is
            // requested from HBM in the same microsecond, and round 0 used to wait for all of it)
            const int l0 = 2 * (mine < GP ? mine : GP);
            mbar_expect_tx(&bars[0], (unsigned)(l0 * L * sizeof(float)));
            for (int l = 0; l < l0; ++l)
                bulk_g2s(lines + (long long)l * L, z_in + (first + l) * L, (unsigned)(L * sizeof(float)), &bars[0]);
        }
    }
    // mu (the update's operand) comes from DRAM: start it towards L2 now, it is first touched after the first transform.
    // (The ground truth is pulled into shared memory by TMA during the grid barrier of the prox phases; requesting it
    // here as well put 16 MB in front of the data the first round of transforms waits for.)
    if (threadIdx.x >= 32 && threadIdx.x < 32 + 2 * (mine < GP ? mine : GP))
        bulk_prefetch_l2(vadd + (first + (threadIdx.x - 32)) * L, (unsigned)(L * sizeof(float)));
    FftTw<L> tw;
    tw.init(t);
    __syncthreads();
    // Everything above reads data that is older than the previous kernel of the chain (the iterate and mu were written
    // before the column pass started); the spectrum, the log slot and the step come from it or may change with it.
    griddep_wait();
    griddep_launch();
    const int cur_slot = slot ? *slot : 0;
    const float st = step_ptr ? *step_ptr : step;
    // minibatch selection of the NEXT iteration (its forward line pass is fused below): a slice per thread of the whole
    // grid, while the first lines and spectrum entries of this CTA are on their way.  The column pass of THIS iteration
    // has consumed and zeroed the selection bytes; the counters advance after the grid barrier, long after this read.
    if (NEXT && sj_next.bits) run_sel_job(sj_next, L, nlines, 0, blockIdx.x * blockDim.x + threadIdx.x, gridDim.x * blockDim.x);

    auto load_spec = [&](int round, float4 (&q)[NQ]) {
#pragma unroll
        for (int n = 0; n < NQ; ++n) {
            const int i = threadIdx.x + n * GP * T;
            const int gg = i % GP, k = i / GP;
            const int pl = round * GP + gg;
            q[n] = (S4 && pl < mine) ? ldg_stream(S4 + (long long)k * npairs + first_pair + pl) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    };
    const int rounds = (mine + GP - 1) / GP;
    float4 qn[NQ];
    if (rounds > 0) load_spec(0, qn);
    if (threadIdx.x == 0 && mine > GP) {            // lines of the later rounds: requested behind the first round's data
        const int l0 = 2 * GP;
        mbar_expect_tx(&bars[1], (unsigned)((2 * mine - l0) * L * sizeof(float)));
        for (int l = l0; l < 2 * mine; ++l)
            bulk_g2s(lines + (long long)l * L, z_in + (first + l) * L, (unsigned)(L * sizeof(float)), &bars[1]);
    }
    if (threadIdx.x >= 32 + 2 * GP && threadIdx.x < 32 + 2 * mine)
        bulk_prefetch_l2(vadd + (first + (threadIdx.x - 32)) * L, (unsigned)(L * sizeof(float)));
    for (int round = 0; round < rounds; ++round) {
        float2 x[EPT];
        if (S) {
            // X[k] = A[k] + i B[k] of the two lines, written re/im swapped for the inverse transform
#pragma unroll
            for (int n = 0; n < NQ; ++n) {
                const int i = threadIdx.x + n * GP * T;
                const int gg = i % GP, k = i / GP;
                const SmemBuf sg{smem + gg * GS, smem + gg * GS + PL};
                const float4 q = qn[n];
                if (k == 0) {
                    sg.put(0, make_float2(q.z, q.x));
                    sg.put(L / 2, make_float2(q.w, q.y));
                } else {
                    sg.put(k, make_float2(q.y + q.z, q.x - q.w));
                    sg.put(L - k, make_float2(q.z - q.y, q.x + q.w));
                }
            }
            if (round + 1 < rounds) load_spec(round + 1, qn);
            __syncthreads();
#pragma unroll
            for (int i = 0; i < EPT; ++i) x[i] = sb.get(IX::in(t, i));
            fft_sync(fbar, T);
            fft_regs<L>(t, sb, x, tw, fbar);
        } else {
            // no spectrum: the gradient term is exactly zero (first inner iteration of an SVRG epoch: z == w, so
            // g_B(z) - g_B(w) = 0 and v = mu); the transform of zeros is skipped, the update below is unchanged
#pragma unroll
            for (int i = 0; i < EPT; ++i) x[i] = make_float2(0.f, 0.f);
        }
        if (round <= 1) mbar_wait(&bars[round], 0);             // z_in lines of this round (round >= 1: of all later rounds) have landed
        const int pl = round * GP + g;
        if (pl < mine) {
            float* lz = lines + (long long)(2 * pl) * L;
            const float* gv = vadd + (first + 2 * pl) * L;
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                float va[EPT];
#pragma unroll
                for (int i = 0; i < EPT; ++i) va[i] = ldg_stream(gv + hh * L + IX::out(t, i));
#pragma unroll
                for (int i = 0; i < EPT; ++i) {
                    // swapped output: .y = real part -> line 2*pair, .x = imag part -> line 2*pair + 1
                    const float gval = (hh == 0 ? x[i].y : x[i].x) * gs;
                    float* pz = lz + hh * L + IX::out(t, i);
                    *pz = *pz - st * (gval + va[i]);
                }
            }
        }
        __syncthreads();                            // exchange buffers free for the next round; lines complete
        trace(310 + round);
    }

    upd_mark(1);
    // the exchange planes in front of the resident lines are idle from here on: they hold the per-warp scratch of the
    // prox phases (sigma selection, Haar transposition)
    static_assert(L < 512 || lines_stage_off<L, GP>() >= 16 * prox_scratch<L>(), "exchange planes too small for the prox scratch");
    // the snapshot lines of the fused forward pass below: towards L2 now (default policy: they are used ~15 us from here),
    // DRAM is idle during the sigma phase and the grid barrier
    if (NEXT && threadIdx.x < 2 * mine)
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(w_next + (first + threadIdx.x) * L), "r"((unsigned)(L * sizeof(float))) : "memory");
    prox_phases<L>(lines, 2 * mine, first, nlines, 1, z_out, xrec, sigma_modifier, fallback_sigma, sig_log, mse_log, cur_slot,
                   reinterpret_cast<unsigned*>(smem), &bars[2], advance, n_advance, gbar, NEXT);
    upd_mark(4);
    if constexpr (NEXT) {
        // ---- forward line pass of the NEXT inner iteration on the lines that are still in shared memory ----
        // (k_lines_r2c on z_new - w: same transform, same unpacking, same stores; what it saves is a kernel boundary --
        // drain of this grid, launch, cold start of a persistent pass, ~6 us at 2048^2 -- and the 4N-byte re-read of z)
        FftTw<L> tw2;                               // (reloaded: keeping the twiddles of the inverse rounds alive across the prox phases spills)
        tw2.init(t);
        __syncthreads();                            // every warp has written its new line; the prox scratch (exchange planes) is idle
        float4* S4o = reinterpret_cast<float4*>(S_next);
        for (int round = 0; round < rounds; ++round) {
            const int pl = round * GP + g;
            float2 x[EPT];
            if (pl < mine) {
                const float* la = lines + (long long)(2 * pl) * L;
                const float* lw = w_next + (first + 2 * pl) * L;
#pragma unroll
                for (int i = 0; i < EPT; ++i) {
                    const int idx = IX::in(t, i);
                    x[i] = make_float2(la[idx] - ldg_stream(lw + idx), la[idx + L] - ldg_stream(lw + idx + L));
                }
            } else {
#pragma unroll
                for (int i = 0; i < EPT; ++i) x[i] = make_float2(0.f, 0.f);
            }
            fft_regs<L>(t, sb, x, tw2, fbar);
            if (FftPlan<L>::NS > 1) fft_sync(fbar, T);
#pragma unroll
            for (int i = 0; i < EPT; ++i) sb.put(IX::out(t, i), x[i]);
            __syncthreads();
            const int pair0 = first_pair + round * GP;
            for (int i = threadIdx.x; i < GP * (L / 2); i += GP * T) {
                const int gg = i % GP, k = i / GP;
                if (round * GP + gg >= mine) continue;
                const SmemBuf sg{smem + gg * GS, smem + gg * GS + PL};
                const float2 xk = sg.get(k);
                const float2 xm = sg.get(k == 0 ? L / 2 : L - k);
                float4 o;
                if (k == 0) o = make_float4(xk.x, xm.x, xk.y, xm.y);
                else o = make_float4(0.5f * (xk.x + xm.x), 0.5f * (xk.y - xm.y), 0.5f * (xk.y + xm.y), 0.5f * (xm.x - xk.x));
                stg_keep(S4o + (long long)k * npairs + pair0 + gg, o);
            }
            __syncthreads();
            trace(330 + round);
        }
    }
    trace_flush();
}

// ------------------------------------------------------------------ selection kernels
__global__ void k_sel_from_indices(unsigned char* __restrict__ bits, int H, int W, const int* __restrict__ idx,
                                   int B, long long idx_img_stride, const int* __restrict__ cursor) {
    const int img = blockIdx.y;
    const int cur = cursor ? *cursor : 0;
    const int* src = idx + (long long)img * idx_img_stride + (long long)cur * B;
    unsigned char* bi = bits + (long long)img * W * (H / 2);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < B; i += gridDim.x * blockDim.x)
        set_sel_bits(bi, H, W, src[i]);
}

__global__ void k_sel_from_feistel(unsigned char* __restrict__ bits, int H, int W,
                                   const int* __restrict__ support, const int* __restrict__ m0,
                                   long long support_img_stride, int B, unsigned seed,
                                   const int* __restrict__ counter, int* __restrict__ idx_out) {
    const int img = blockIdx.y;
    trace(400, true);
    const unsigned key = mix32(seed ^ mix32((counter ? (unsigned)*counter : 0u) * 0x632be5abU + (unsigned)img));
    const int* sup = support + (long long)img * support_img_stride;
    const unsigned n = (unsigned)m0[img];
    unsigned char* bi = bits + (long long)img * W * (H / 2);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < B; i += gridDim.x * blockDim.x) {
        if ((unsigned)i >= n) break;                  // the cycle walk only terminates inside the domain [0, n)
        const int k = sup[feistel_perm((unsigned)i, n, key)];
        if (idx_out) idx_out[(long long)img * B + i] = k;
        set_sel_bits(bi, H, W, k);
    }
    trace(409);
    trace_flush();
}

__global__ void k_sample_indices(int* __restrict__ idx_out, int n, int count, unsigned seed,
                                 const int* __restrict__ counter) {
    const unsigned key = mix32(seed ^ mix32((counter ? (unsigned)*counter : 0u) * 0x632be5abU));
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < count && i < n; i += gridDim.x * blockDim.x)
        idx_out[i] = (int)feistel_perm((unsigned)i, (unsigned)n, key);
}

}  // namespace pnp
