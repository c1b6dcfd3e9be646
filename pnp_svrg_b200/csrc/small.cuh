// PnP-SVRG (paper-mode variance reduction) + wavelet prox for SMALL CSMRI images: one thread-block CLUSTER per image,
// the whole run -- every snapshot gradient and every inner iteration -- in ONE launch, the image resident in the
// cluster's shared memory.
//
// Reference: algorithms/pnp_svrg.py:26-95 (the loop), problems/CSMRI.py:76-89 (gradients), denoisers/TV.py:21-26 and
// the estimate_sigma call at pnp_svrg.py:71 (prox).  Same arithmetic as the three-pass path of csmri.cuh + prox.cuh
// (same unpacking, same selection, same epilogues, same sigma estimate and shrink; at 256 a transform plan with other
// radices), so the two paths agree to fp32 rounding.
//
// WHY.  A 256x256 iteration moves ~2 MB: on the multi-launch path it is five dependent grid-wide steps of ~4 us each
// (launch + drain + first-load latency), 1 % of the HBM roofline.  Here nothing leaves the SMs:
//   * cluster of C = 8 CTAs, CTA q owns the lines [q L/C, (q+1) L/C) of z, w and mu (shared memory, padded rows) and
//     the packed spectrum rows [q L/(2C), (q+1) L/(2C));
//   * the two transpositions of the 2-D transform are 16-byte stores into the shared memory of the CTA that owns the
//     destination row / line pair (distributed shared memory), followed by a cluster barrier;
//   * an FFT of length L <= 256 is done by at most 32 threads, i.e. inside one warp (length 256: the 8 x 8 x 4 plan
//     variant, one whole warp per transform, so that all 16 warps of a CTA of a cluster of 8 transform at once): the
//     Stockham exchanges need warp barriers only (fft_regs<L, true, V>), the warps of a phase do not run in lock step;
//   * the minibatch selection of iteration t is built in shared memory by the warps that have no transform in the
//     forward line phase (every CTA scans the B drawn positions and keeps the bits of its own rows);
//   * the sigma estimate's mean over lines is a cluster reduction: every CTA writes its partial sum into a slot of
//     every peer and all of them add the C slots in the same order (deterministic, identical in all CTAs).
// Per inner iteration: 3 cluster barriers, no global traffic except the ground truth (L2) for the PSNR log.
// Batches: blockIdx.x / C = image; the sweeps launch one cluster per reconstruction.
#pragma once
#include <cooperative_groups.h>
#include "csmri.cuh"
#include "prox.cuh"

namespace pnp {

struct SmallArgs {
    float* z;                         // [batch][L lines][L] iterate, in / out
    const float* xrec;                // optional ground truth, same layout (PSNR log)
    const float2 *Y1, *Y2, *Y1n, *Y2n;
    const unsigned char* bits_full;   // [batch][L/2][L] selection bytes of the full mask (set_sel_bits layout)
    const int* support;               // device sampler: sampled positions per image
    const int* m0;
    long long support_img_stride;
    const int* idx;                   // explicit minibatches (optional): idx[img * img_stride + t * iter_stride + i]
    long long idx_img_stride, idx_iter_stride;
    const float* snap_scale_ptr;      // per image 1/M0 (or null: snap_scale)
    float snap_scale;
    const float* step;                // per image step of the first epoch of this launch
    long long step_img_stride;        // 0: one step for all images
    double* sig_log;                  // [slot][batch] accumulated (+=), slots slot0 .. slot0 + n_inner - 1
    double* mse_log;
    const int* slot;                  // optional device counters: first log slot, first draw counter
    const int* draw_counter;
    int batch, n_inner, T2, B;
    unsigned seed;
    float lr_decay, sigma_modifier, fallback_sigma, fallback_decay;
};

template <int L, int C, int NT_ = 512> struct SmallCfg {
    static constexpr int V = L == 256 ? 1 : 0;              // transform plan variant (fft_core.cuh): 256 = 8 x 8 x 4 on 32 threads
    static constexpr int T = fft_threads<L, V>();
    static constexpr int EPT = FftPlan<L, V>::EPT;
    static constexpr int NT = NT_, NW = NT / 32;
    static constexpr int LPC = L / C;                       // lines per CTA
    static constexpr int PPC = LPC / 2;                     // line pairs per CTA = transforms per line phase
    static constexpr int RPC = (L / 2) / C;                 // packed spectrum rows per CTA (== PPC)
    static constexpr int LS = L + 8;                        // line stride (floats): the two lines of the half-warps of a warp sit 16 banks apart
    static constexpr int PL2 = 2 * fft_plane<L>();
    static constexpr int GS = PL2 + ((2 - PL2 % 32) + 32) % 32;   // exchange-buffer stride == 2 (mod 32): group g is shifted by g float2
    static constexpr int GPW = 32 / T;                      // transforms per warp
    static constexpr int FT = PPC * T;                      // threads that transform in a line phase
    // column phase: warp 0 holds packed row 0 of the CTA alone (in CTA 0 that row is the DC + i Nyquist packing and takes
    // another code path: a warp must not diverge around the warp barriers of the transforms), rows 1 .. RPC - 1 follow
    static constexpr int BT = ((RPC - 1 + GPW) * T + 31) / 32 * 32;     // threads of the column phase (whole warps)
    static constexpr int BG = BT / T;                       // its transform groups (the last ones may have no row)
    static constexpr int NEX = BG > PPC ? BG : PPC;
    static constexpr int NLW = LPC / NW;                    // lines per warp in the prox phases
    static constexpr int EXF = NEX * GS > NW * NLW * PNP_SIG_SCRATCH ? NEX * GS : NW * NLW * PNP_SIG_SCRATCH;
    static constexpr int OFF_Z = 0;
    static constexpr int OFF_W = OFF_Z + LPC * LS;
    static constexpr int OFF_MU = OFF_W + LPC * LS;
    static constexpr int OFF_S = OFF_MU + LPC * LS;         // float2 [RPC][L]
    static constexpr int OFF_T = OFF_S + RPC * L * 2;       // float4 [L/2][PPC]
    static constexpr int OFF_EX = OFF_T + (L / 2) * PPC * 4;
    static constexpr int OFF_BITS = (OFF_EX + EXF + 3) & ~3;          // 3 x RPC x L bytes: minibatch (double buffered), full mask
    static constexpr int OFF_RED = OFF_BITS + 3 * RPC * L / 4;        // doubles: [C] sigma slots, [NW] + [NW] per-warp partials
    static constexpr int FLOATS = OFF_RED + 2 * (C + 2 * NW);
    static constexpr size_t SMEM = sizeof(float) * (size_t)FLOATS;
    static_assert(PPC == RPC, "line pairs and packed rows per CTA");
    static_assert(T <= 32 && 32 % T == 0 && FT % 32 == 0 && FT <= NT && BT <= NT, "transform groups must tile whole warps of the CTA");
    static_assert(LPC % NW == 0 && NLW >= 1 && NLW <= 2, "one or two lines per warp in the prox phases");
    static_assert(L % 128 == 0 && L <= 256, "chunk-cyclic Haar layout: 128 or 256 samples per line");
    static_assert(OFF_RED % 2 == 0, "double alignment");
};

__device__ __forceinline__ void named_bar(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }

// see "Minibatch selection" in the kernel below
// or the bits of k-space position k (set_sel_bits of csmri.cuh) into the selection rows of their owner CTAs
template <int L, int C, int NT, class Cluster>
__device__ __forceinline__ void set_sel_bits_cluster(Cluster& cluster, unsigned char* bits, int k) {
    constexpr int hp = L / 2, RPC = SmallCfg<L, C, NT>::RPC;
    const int ky = k / L, kx = k % L;                              // L is a power of two: shifts and masks
    const int kym = (L - ky) % L, kxm = (L - kx) % L;
    auto put = [&](int row, int col, unsigned v) {
        const int byte_idx = (row % RPC) * L + col;
        unsigned* w = cluster.map_shared_rank(reinterpret_cast<unsigned*>(bits), row / RPC) + (byte_idx >> 2);
        atomicOr(w, v << (8 * (byte_idx & 3)));
    };
    if (ky < hp) put(ky, kx, 1u);
    else if (ky == hp) put(0, kx, 4u);
    if (kym < hp) put(kym, kxm, 2u);
    else if (kym == hp) put(0, kxm, 8u);
}

// CTA q of the cluster draws positions q, q + C, ... of minibatch `it` and ORs their bits into buffer it & 1 of the CTAs
// that own the rows (distributed shared-memory atomics; the buffers were zeroed a cluster barrier earlier).  Two halves
// around the transforms of the column phase: `fetch` (permutation + support lookup: a dependent global load whose latency
// the transforms hide) and `commit` (the atomics); every thread of the CTA takes positions q + C * (tid + n * NT).
template <int L, int C, int NT>
struct SmallSel {
    const int* src;
    const int* sup;
    unsigned key, n;
    int count, q;
    FeistelDom dom;
    __device__ __forceinline__ int fetch(int i) const {
        if (i >= count) return -1;
        return src ? src[i] : sup[feistel_perm((unsigned)i, n, key, dom)];
    }
};
template <int L, int C, int NT>
__device__ __forceinline__ void small_select_clear(unsigned char* bits_mb, int buf) {
    using K = SmallCfg<L, C, NT>;
    unsigned char* bits = bits_mb + buf * (K::RPC * L);
    for (int i = threadIdx.x; i < K::RPC * L / 16; i += NT) reinterpret_cast<uint4*>(bits)[i] = make_uint4(0u, 0u, 0u, 0u);
}

template <int L, int C, int NT = 512>
__global__ void __launch_bounds__(NT, NT <= 256 ? 2 : 1) k_csmri_svrg_small(SmallArgs a) {
    using K = SmallCfg<L, C, NT>;
    constexpr int V = K::V;
    using IX = FftIdx<L, V>;
    constexpr int T = K::T, EPT = K::EPT, LPC = K::LPC, PPC = K::PPC, RPC = K::RPC, LS = K::LS, GS = K::GS, FT = K::FT,
                  BT = K::BT, NW = K::NW, NLW = K::NLW, GPW = K::GPW;
    constexpr int LEVELS = HaarCfg<L>::LEVELS, NCH = L / 128;
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    const int q = (int)cluster.block_rank();
    const int img = blockIdx.x / C;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    extern __shared__ __align__(128) float smem[];
    float* zl = smem + K::OFF_Z;
    float* wl = smem + K::OFF_W;
    float* mul = smem + K::OFF_MU;
    float2* Sb = reinterpret_cast<float2*>(smem + K::OFF_S);
    float4* Tb = reinterpret_cast<float4*>(smem + K::OFF_T);
    float* ex = smem + K::OFF_EX;
    unsigned char* bits_mb = reinterpret_cast<unsigned char*>(smem + K::OFF_BITS);       // [2][RPC][L]
    unsigned char* bits_fu = bits_mb + 2 * RPC * L;
    double* red_sig = reinterpret_cast<double*>(smem + K::OFF_RED);
    double* red_w = red_sig + C;
    double* red_e = red_w + NW;

    const long long N = (long long)L * L;
    float* zg = a.z + (long long)img * N + (long long)q * LPC * L;
    const float* xg = a.xrec ? a.xrec + (long long)img * N + (long long)q * LPC * L : nullptr;
    const long long yoff = (long long)img * (L / 2) * L;
    const bool is_fft = tid < FT;
    const int g = is_fft ? tid / T : 0, t = tid % T;
    const SmemBuf sb{ex + g * GS, nullptr};
    FftTw<L, V> tw;
    tw.init(t);

    // ---- resident state: my lines of z, my rows of the full-mask selection ----
    for (int i = tid; i < LPC * (L / 4); i += K::NT) {
        const int l = i / (L / 4), c = i % (L / 4);
        reinterpret_cast<float4*>(zl + l * LS)[c] = reinterpret_cast<const float4*>(zg + (long long)l * L)[c];
    }
    {
        const uint4* src = reinterpret_cast<const uint4*>(a.bits_full + yoff + (long long)q * RPC * L);
        for (int i = tid; i < RPC * L / 16; i += K::NT) reinterpret_cast<uint4*>(bits_fu)[i] = src[i];
    }
    const int slot0 = a.slot ? *a.slot : 0;
    const unsigned draw0 = a.draw_counter ? (unsigned)*a.draw_counter : 0u;
    const float inv_n = (float)(1.0 / ((double)L * (double)L));
    const float gs_snap = inv_n * (a.snap_scale_ptr ? a.snap_scale_ptr[img] : a.snap_scale);
    const float gs_in = inv_n * (1.0f / (float)a.B);
    double step_d = (double)a.step[(long long)img * a.step_img_stride];
    float st = (float)step_d;
    trace(500, true);
    small_select_clear<L, C, NT>(bits_mb, 0);
    small_select_clear<L, C, NT>(bits_mb, 1);
    cluster.sync();                                              // every CTA of the cluster runs: remote stores may begin
    trace(501);

    // ================================================================= phases
    // forward line pass: two real lines -> one complex transform, unpacked rows sent to their owners (k_lines_r2c)
    auto phase_a = [&](bool sub) {
        float2 x[EPT];
        const float* la = zl + (2 * g) * LS;
        const float* lw = wl + (2 * g) * LS;
#pragma unroll
        for (int i = 0; i < EPT; ++i) {
            const int idx = IX::in(t, i);
            float re = la[idx], im = la[idx + LS];
            if (sub) { re -= lw[idx]; im -= lw[idx + LS]; }
            x[i] = make_float2(re, im);
        }
        fft_regs<L, true, V>(t, sb, x, tw);
        __syncwarp();
#pragma unroll
        for (int i = 0; i < EPT; ++i) sb.put(IX::out(t, i), x[i]);
        named_bar(1, FT);
        float4* S4 = reinterpret_cast<float4*>(Sb);
        for (int i = tid; i < PPC * (L / 2); i += FT) {
            const int gg = i % PPC, k = i / PPC;
            const SmemBuf sg{ex + gg * GS, nullptr};
            const float2 xk = sg.get(k);
            const float2 xm = sg.get(k == 0 ? L / 2 : L - k);
            float4 o;
            if (k == 0) o = make_float4(xk.x, xm.x, xk.y, xm.y);
            else o = make_float4(0.5f * (xk.x + xm.x), 0.5f * (xk.y - xm.y), 0.5f * (xk.y + xm.y), 0.5f * (xm.x - xk.x));
            float4* dst = cluster.map_shared_rank(S4, k / RPC) + (k % RPC) * (L / 2) + q * PPC + gg;
            *dst = o;
        }
    };

    // Minibatch selection of inner iteration `it`: selection bytes of my rows in buffer it & 1.  Built one iteration AHEAD,
    // during the column phase of iteration it - 1: every CTA draws 1/C of the positions (fetched before its transforms,
    // so that the dependent support lookup is hidden) and ORs the bits into the rows' owners through distributed
    // shared-memory atomics after them; the buffer was zeroed by its owner during the inverse line phase of iteration
    // it - 2, right after its last use.  Cluster barriers separate zeroing, filling and use.  (Inside the forward line
    // phase, every CTA scanning all B positions for its own rows, the selection took 3.3 us for B = 1000 against 1.9 us
    // of transforms; spread over the other phases the same work slowed those down by as much.)
    const unsigned sel_n = a.idx ? 0u : (unsigned)a.m0[img];
    const int sel_count = a.idx ? a.B : ((unsigned)a.B < sel_n ? a.B : (int)sel_n);
    SmallSel<L, C, NT> ss;
    ss.sup = a.support + (long long)img * a.support_img_stride;
    ss.n = sel_n; ss.count = sel_count; ss.q = q; ss.dom = feistel_domain(sel_n > 1u ? sel_n : 2u);
    auto sel_begin = [&](int it) {                               // -> first position of this thread (or -1)
        ss.src = a.idx ? a.idx + (long long)img * a.idx_img_stride + (long long)it * a.idx_iter_stride : nullptr;
        ss.key = mix32(a.seed ^ mix32((draw0 + (unsigned)it) * 0x632be5abU + (unsigned)img));
        return ss.fetch(q + C * (NT - 1 - tid));                 // from the last warp down: those have no transform when warps are spare
    };
    auto sel_end = [&](int it, int k0) {
        unsigned char* bits = bits_mb + (it & 1) * (RPC * L);
        if (k0 >= 0) set_sel_bits_cluster<L, C, NT>(cluster, bits, k0);
        for (int i = q + C * (NT - 1 - tid + NT); i < sel_count; i += C * NT) set_sel_bits_cluster<L, C, NT>(cluster, bits, ss.fetch(i));
    };

    // column pass on my packed rows: forward, selection (and measurements), inverse, sent to the owners of the lines
    auto phase_b = [&](const unsigned char* bits, bool use_y) {
        float2* T2f = reinterpret_cast<float2*>(Tb);             // [L/2][LPC]
        const int gb = tid / T;                                  // transform group of the column phase
        const int rl = gb < GPW ? (gb == 0 ? 0 : -1) : gb - (GPW - 1);      // my packed row of this CTA (warp 0: row 0 only)
        if (tid < BT && !(warp == 0 && q == 0)) {
            const bool active = rl >= 0 && rl < RPC;
            const int r = active ? rl : 0;
            const int kyp = q * RPC + r;
            const long long crow = yoff + (long long)kyp * L;
            const SmemBuf sbb{ex + gb * GS, nullptr};
            float2 x[EPT], y[EPT];
            unsigned long long bbp = 0ull;
#pragma unroll
            for (int i = 0; i < EPT; ++i) x[i] = active ? Sb[r * L + IX::in(t, i)] : make_float2(0.f, 0.f);
#pragma unroll
            for (int m = 0; m < EPT; ++m) bbp |= (unsigned long long)(bits[r * L + t + T * m] & 0xFu) << (4 * m);
            fft_regs<L, true, V>(t, sbb, x, tw);
#pragma unroll
            for (int m = 0; m < EPT; ++m) {
                const int kx = t + T * m;
                const float2 o = apply_sel(x[IX::out_slot(m)], active ? (unsigned)(bbp >> (4 * m)) & 0xFu : 0u, use_y, a.Y1 + crow + kx, a.Y2 + crow + kx);
                y[IX::in_slot(m)] = cswap(o);
            }
            __syncwarp();
            fft_regs<L, true, V>(t, sbb, y, tw);
            if (active) {
#pragma unroll
                for (int i = 0; i < EPT; ++i) {
                    const int c = IX::out(t, i);
                    float2* dst = cluster.map_shared_rank(T2f, c / LPC) + kyp * LPC + (c % LPC);
                    *dst = cswap(y[i]);
                }
            }
        } else if (warp == 0 && q == 0) {
            // packed row 0: C = FFT(DC + i * Nyquist); split, select each of the two rows, re-pack (k_cols_mask, column 0)
            const int g0 = lane / T;
            const SmemBuf s0{ex + g0 * GS, nullptr};
            float2 x[EPT];
#pragma unroll
            for (int i = 0; i < EPT; ++i) x[i] = g0 == 0 ? Sb[IX::in(t, i)] : make_float2(0.f, 0.f);
            fft_regs<L, true, V>(t, s0, x, tw);
            __syncwarp();
#pragma unroll
            for (int i = 0; i < EPT; ++i) s0.put(IX::out(t, i), x[i]);
            __syncwarp();
            {
                // all lanes of the warp split / select / re-pack the spectrum of group 0
                const SmemBuf sz{ex, nullptr};
                const float2* y1i = a.Y1 + yoff;
                const float2* y2i = a.Y2 + yoff;
                const float2* y1n = a.Y1n + (long long)img * L;
                const float2* y2n = a.Y2n + (long long)img * L;
                for (int kx = lane; kx <= L / 2; kx += 32) {
                    const int km = (L - kx) % L;
                    const float2 ck = sz.get(kx), cm = sz.get(km);
                    const float2 fdc = make_float2(0.5f * (ck.x + cm.x), 0.5f * (ck.y - cm.y));
                    const float2 fny = make_float2(0.5f * (ck.y + cm.y), 0.5f * (cm.x - ck.x));
                    const unsigned bk = bits[kx], bm = bits[km];
                    const float2 dk = apply_sel(fdc, bk, use_y, y1i + kx, y2i + kx);
                    const float2 nk = apply_sel(fny, bk >> 2, use_y, y1n + kx, y2n + kx);
                    const float2 dm = apply_sel(make_float2(fdc.x, -fdc.y), bm, use_y, y1i + km, y2i + km);
                    const float2 nm = apply_sel(make_float2(fny.x, -fny.y), bm >> 2, use_y, y1n + km, y2n + km);
                    sz.put(kx, make_float2(dk.y + nk.x, dk.x - nk.y));
                    if (km != kx) sz.put(km, make_float2(dm.y + nm.x, dm.x - nm.y));
                }
            }
            __syncwarp();
#pragma unroll
            for (int i = 0; i < EPT; ++i) x[i] = s0.get(IX::in(t, i));
            __syncwarp();
            fft_regs<L, true, V>(t, s0, x, tw);
            if (g0 == 0) {
#pragma unroll
                for (int i = 0; i < EPT; ++i) {
                    const int c = IX::out(t, i);
                    float2* dst = cluster.map_shared_rank(T2f, c / LPC) + (c % LPC);
                    *dst = cswap(x[i]);
                }
            }
        }
    };

    // inverse line pass + epilogue on my lines (k_lines_c2r): snap: mu = g * gs, w = z ; else z -= st * (g * gs + mu)
    auto phase_c = [&](bool snap, float gs) {
        if (!is_fft) return;
        constexpr int NQ = (L / 2) / T;
#pragma unroll
        for (int n = 0; n < NQ; ++n) {
            const int i = tid + n * FT;
            const int gg = i % PPC, k = i / PPC;
            const SmemBuf sg{ex + gg * GS, nullptr};
            const float4 v = Tb[k * PPC + gg];
            if (k == 0) {
                sg.put(0, make_float2(v.z, v.x));
                sg.put(L / 2, make_float2(v.w, v.y));
            } else {
                sg.put(k, make_float2(v.y + v.z, v.x - v.w));
                sg.put(L - k, make_float2(v.z - v.y, v.x + v.w));
            }
        }
        named_bar(1, FT);
        float2 x[EPT];
#pragma unroll
        for (int i = 0; i < EPT; ++i) x[i] = sb.get(IX::in(t, i));
        __syncwarp();
        fft_regs<L, true, V>(t, sb, x, tw);
        float* lz = zl + (2 * g) * LS;
        float* lm = mul + (2 * g) * LS;
        float* lw = wl + (2 * g) * LS;
#pragma unroll
        for (int i = 0; i < EPT; ++i) {
            const int idx = IX::out(t, i);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const float gval = (h == 0 ? x[i].y : x[i].x) * gs;          // swapped output: .y real -> line 2g, .x imag -> line 2g + 1
                const int o = idx + h * LS;
                if (snap) { lm[o] = gval; lw[o] = lz[o]; }
                else lz[o] = lz[o] - st * (gval + lm[o]);
            }
        }
        named_bar(1, FT);                                        // the exchange buffers become the prox scratch
    };

    // ================================================================= the run
    // One pass of the loop body is either a snapshot (mu = grad_full(z) * snap_scale, w = z; pnp_svrg.py:32-35) or an
    // inner iteration (v = g_B(z - w) / B + mu ; z <- prox(z - step * v); pnp_svrg.py:52-57, 71-76): both are the same
    // three phases with different operands, so every phase has ONE call site (and is inlined).
    if (a.n_inner > 0) sel_end(0, sel_begin(0));                 // the first minibatch (a snapshot pass and its barriers come before its use)
    bool need_snap = true;
    for (int it = 0; it < a.n_inner;) {
        const bool snap = need_snap;
        if (snap && it > 0) { step_d *= (double)a.lr_decay; st = (float)step_d; }
        // First inner iteration of an epoch: z == w bit for bit (the snapshot pass has just copied it), so the transform of
        // z - w, its selection and its inverse are exactly zero and v = mu: the three phases reduce to the update.  The two
        // cluster barriers stay (they order the selection buffers: zeroing, remote filling, use).
        const bool zero = !snap && it % a.T2 == 0;
        trace(510);
        if (is_fft && !zero) phase_a(!snap);
        trace(511);
        cluster.sync();
        trace(512);
        const bool sel_next = !snap && it + 1 < a.n_inner;
        const int k0 = sel_next ? sel_begin(it + 1) : -1;
        if (!zero) phase_b(snap ? bits_fu : bits_mb + (it & 1) * (RPC * L), snap);
        if (sel_next) sel_end(it + 1, k0);
        trace(513);
        cluster.sync();
        trace(514);
        if (zero) {
            for (int i = tid; i < LPC * L; i += K::NT) {
                const int o = (i / L) * LS + (i % L);
                zl[o] = zl[o] - st * (0.f * gs_in + mul[o]);
            }
        } else {
            phase_c(snap, snap ? gs_snap : gs_in);
        }
        if (!snap) small_select_clear<L, C, NT>(bits_mb, it & 1);   // used by the column phase above; filled again next iteration
        trace(515);
        __syncthreads();
        if (snap) { need_snap = false; continue; }

        // sigma estimate per line + forward Haar pyramid kept in registers (neither needs the mean)
        float hx[NLW][NCH][4], hA[NLW][NCH][6], hD[NLW][NCH][5], hT[NLW][4], level_ss[NLW];
        float4 xr[NLW][NCH];
        double sig_sum = 0.0;
        {
            const float* lx[NLW];
            unsigned* scr[NLW];
            double sg[NLW];
#pragma unroll
            for (int j = 0; j < NLW; ++j) {
                lx[j] = zl + (warp + NW * j) * LS;
                scr[j] = reinterpret_cast<unsigned*>(ex) + (warp * NLW + j) * PNP_SIG_SCRATCH;
            }
#ifndef PNP_SMALL_INTERLEAVE
            // (two lines side by side in one warp -- line_sigma_mad_n<L, 2> -- measured 4 % slower: register pressure)
#pragma unroll
            for (int j = 0; j < NLW; ++j) {
                const float* l1[1] = {lx[j]};
                unsigned* s1[1] = {scr[j]};
                double o1[1];
                line_sigma_mad_n<L, 1, false>(l1, lane, s1, o1);
                sg[j] = o1[0];
            }
#else
            line_sigma_mad_n<L, NLW, false>(lx, lane, scr, sg);
#endif
#pragma unroll
            for (int j = 0; j < NLW; ++j) sig_sum += sg[j];        // medians; / Phi^-1(0.75) once, after the cluster reduction
        }
#pragma unroll
        for (int j = 0; j < NLW; ++j) {
            const int l = warp + NW * j;
            const float* sl = zl + l * LS;
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                const float4 v = reinterpret_cast<const float4*>(sl)[c * 32 + lane];
                hx[j][c][0] = v.x; hx[j][c][1] = v.y; hx[j][c][2] = v.z; hx[j][c][3] = v.w;
                if (xg) xr[j][c] = reinterpret_cast<const float4*>(xg + (long long)l * L)[c * 32 + lane];
            }
        }
        float ss[NLW][LEVELS];
#pragma unroll
        for (int j = 0; j < NLW; ++j) {
#pragma unroll
            for (int k = 0; k < LEVELS; ++k) ss[j][k] = 0.f;
            haar_cc_forward<LEVELS, NCH>(hx[j], hA[j], hD[j], hT[j], ss[j], lane);
        }
#pragma unroll
        for (int j = 0; j < NLW; ++j) {
            float mine_ss = 0.f;
#pragma unroll
            for (int k = 0; k < LEVELS; ++k) {
                const float e = warp_sum_f(ss[j][k]);
                mine_ss = lane == k ? e : mine_ss;
            }
            level_ss[j] = mine_ss;
        }
        if (lane == 0) red_w[warp] = sig_sum;
        trace(516);
        __syncthreads();
        if (tid < C) {
            double s = 0.0;
#pragma unroll
            for (int w = 0; w < NW; ++w) s += red_w[w];
            *(cluster.map_shared_rank(red_sig, tid) + q) = s;    // my partial sum into slot q of CTA `tid`
        }
        trace(517);
        cluster.sync();
        trace(518);
        double tot = 0.0;
#pragma unroll
        for (int r = 0; r < C; ++r) tot += red_sig[r];
        tot /= 0.6744897501960817;
        const double se = tot / (double)L;
        const float fb = a.fallback_decay == 1.0f ? a.fallback_sigma : a.fallback_sigma * powf(a.fallback_decay, (float)it);
        const float sigma = (se > 0.0) ? (float)(se * (double)a.sigma_modifier) : fb;
        const float var = sigma * sigma;
        if (q == 0 && tid == 0) atomicAdd(a.sig_log + (long long)(slot0 + it) * a.batch + img, tot);

        // BayesShrink thresholds, inverse pyramid, new iterate into the resident lines, squared error
        float err = 0.f;
#pragma unroll
        for (int j = 0; j < NLW; ++j) {
            const int l = warp + NW * j;
            float thr[LEVELS];
            {
                const int k = lane < LEVELS ? lane : LEVELS - 1;
                const float dvar = level_ss[j] / (float)(L >> (k + 1));
                const float tk = var / sqrtf(fmaxf(dvar - var, 2.220446049250313e-16f));
#pragma unroll
                for (int kk = 0; kk < LEVELS; ++kk) thr[kk] = __shfl_sync(0xffffffffu, tk, kk);
            }
            haar_cc_inverse<LEVELS, NCH>(hx[j], hA[j], hD[j], hT[j], thr, lane);
            float* sl = zl + l * LS;
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                reinterpret_cast<float4*>(sl)[c * 32 + lane] = make_float4(hx[j][c][0], hx[j][c][1], hx[j][c][2], hx[j][c][3]);
                if (xg) {
                    const float e0 = hx[j][c][0] - xr[j][c].x, e1 = hx[j][c][1] - xr[j][c].y, e2 = hx[j][c][2] - xr[j][c].z,
                                e3 = hx[j][c][3] - xr[j][c].w;
                    err = fmaf(e0, e0, fmaf(e1, e1, fmaf(e2, e2, fmaf(e3, e3, err))));
                }
            }
        }
        if (xg && a.mse_log) {
            err = warp_sum_f(err);
            if (lane == 0) red_e[warp] = (double)err;
        }
        trace(519);
        __syncthreads();                                         // my lines are complete: the next line pass may read them
        if (xg && a.mse_log && tid == 0) {
            double s = 0.0;
#pragma unroll
            for (int w = 0; w < NW; ++w) s += red_e[w];
            atomicAdd(a.mse_log + (long long)(slot0 + it) * a.batch + img, s);
        }
        ++it;
        need_snap = it % a.T2 == 0;
    }

    for (int i = tid; i < LPC * (L / 4); i += K::NT) {
        const int l = i / (L / 4), c = i % (L / 4);
        reinterpret_cast<float4*>(zg + (long long)l * L)[c] = reinterpret_cast<const float4*>(zl + l * LS)[c];
    }
    trace(520);
    trace_flush();
}

}  // namespace pnp
