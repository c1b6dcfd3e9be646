// libpnp_b200.so -- C ABI (include/pnp_b200.h) over the sm_100a kernels of the PnP hot path.
// One translation unit: the kernels live in the .cuh files next to this one.
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <thread>
#include <tuple>
#include <vector>

#include "../../include/pnp_b200.h"
#include "host_sampler.h"
#include "csmri.cuh"
#include "prox.cuh"
#include "vr.cuh"
#include "deblur.cuh"
#include "pr.cuh"
#include "nlm.cuh"
#include "cnn_fp32.cuh"
#include "cnn_tc.cuh"
#include "tv_chambolle.cuh"
#include "cdp.cuh"
#include "small.cuh"
#include "build.cuh"

namespace {

thread_local char g_err[512] = "";
bool g_init[64] = {false};

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define CU_TRY(expr)                                                                            \
    do {                                                                                        \
        cudaError_t e__ = (expr);                                                               \
        if (e__ != cudaSuccess)                                                                 \
            return fail(PNP_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__),  \
                        __FILE__, __LINE__);                                                    \
    } while (0)

#define LAUNCH_CHECK() CU_TRY(cudaGetLastError())

bool pow2_ok(int n) { return n >= 32 && n <= 4096 && (n & (n - 1)) == 0; }

int check_init() {
    int dev = 0;
    CU_TRY(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || !g_init[dev]) return fail(PNP_ERR_NOT_INIT, "pnp_init() not called on device %d", dev);
    return PNP_OK;
}

// ---- per-size launch geometry ------------------------------------------------------------
#ifndef PNP_GP256
#define PNP_GP256 4                                    // line pairs per CTA at L = 256 (16 threads per transform)
#endif
template <int L> constexpr int lines_gp() {            // line pairs per CTA in passes 1 and 3
#ifdef PNP_LINES_GP1
    return pnp::fft_threads<L>() >= 128 ? 1 : 4;
#else
    return pnp::fft_threads<L>() >= 256 ? 1 : (pnp::fft_threads<L>() >= 128 ? 2 : (pnp::fft_threads<L>() == 16 ? PNP_GP256 : 4));
#endif
}
template <int L> constexpr int cols_nc() {             // packed columns per CTA in pass 2
    return pnp::fft_threads<L>() >= 128 ? 1 : (pnp::fft_threads<L>() >= 64 ? 2 : (64 / pnp::fft_threads<L>() > 8 ? 8 : 64 / pnp::fft_threads<L>()));
}
// exchange planes + two staging buffers of 2*GP lines each (TMA bulk copy targets)
template <int L> constexpr size_t lines_smem() {
    return sizeof(float) * (pnp::lines_stage_off<L, lines_gp<L>()>() + 2 * lines_gp<L>() * pnp::stage_pair_stride<L>());
}
// exchange planes + one staging buffer of NC columns (complex)
template <int L> constexpr size_t cols_smem() {
    return sizeof(float) * (pnp::cols_stage_off<L, cols_nc<L>()>() + 2 * cols_nc<L>() * L) + (size_t)cols_nc<L>() * L;
}
template <int L> constexpr size_t conv_smem() { return sizeof(float) * 2 * pnp::fft_plane<L>() * cols_nc<L>(); }

template <int L>
int set_attrs() {
    CU_TRY(cudaFuncSetAttribute(pnp::k_lines_r2c<L, lines_gp<L>()>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lines_smem<L>()));
    CU_TRY(cudaFuncSetAttribute(pnp::k_lines_c2r<L, lines_gp<L>(), true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lines_smem<L>()));
    CU_TRY(cudaFuncSetAttribute(pnp::k_lines_c2r<L, lines_gp<L>(), false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lines_smem<L>()));
    CU_TRY(cudaFuncSetAttribute(pnp::k_cols_mask<L, cols_nc<L>()>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cols_smem<L>()));
    cudaFuncAttributes fa;
    CU_TRY(cudaFuncGetAttributes(&fa, pnp::k_sigma_mad<L>));
    CU_TRY(cudaFuncGetAttributes(&fa, pnp::k_haar_bayes<L>));
    CU_TRY(cudaFuncSetAttribute(pnp::k_cols_conv<L, cols_nc<L>()>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)conv_smem<L>()));
    return PNP_OK;
}

template <int L>
int launch_conv(float2* S, const float2* Bf, const float2* twn, int H, int batch, int conj_kernel, cudaStream_t st) {
    constexpr int NC = cols_nc<L>();
    const int tasks = H / 2 + 1;
    dim3 grid((tasks + NC - 1) / NC, batch);
    pnp::k_cols_conv<L, NC><<<grid, NC * pnp::fft_threads<L>(), conv_smem<L>(), st>>>(S, Bf, twn, H, conj_kernel,
                                                                                     (long long)tasks * L);
    LAUNCH_CHECK();
    return PNP_OK;
}

int g_num_sms = 148;              // SM count of the device the calling thread last initialised / used
int g_sms_of[64] = {0};

int current_device() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = 0;
    return dev;
}
int num_sms() {
    const int n = g_sms_of[current_device()];
    return n > 0 ? n : g_num_sms;
}

// resident CTAs of a persistent kernel: occupancy is cached per (kernel, threads, shared memory, device)
int persistent_ctas(const void* kernel, int threads, size_t smem, int items, int batch) {
    static std::mutex mu;
    static std::map<std::tuple<const void*, int, size_t, int>, int> cache;
    const int dev = current_device();
    int per_sm = 0;
    {
        std::lock_guard<std::mutex> lock(mu);
        const auto key = std::make_tuple(kernel, threads, smem, dev);
        auto it = cache.find(key);
        if (it == cache.end()) {
            int n = 0;
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, threads, smem) != cudaSuccess || n < 1) n = 1;
            it = cache.emplace(key, n).first;
        }
        per_sm = it->second;
    }
    long long cap = (long long)num_sms() * per_sm / (batch > 0 ? batch : 1);
    if (cap < 1) cap = 1;
    return (int)(items < cap ? items : cap);
}

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) once per (kernel, device)
int raise_smem_limit(const void* kernel, int bytes) {
    static std::mutex mu;
    static std::map<std::pair<const void*, int>, int> done;
    const int dev = current_device();
    std::lock_guard<std::mutex> lock(mu);
    int& have = done[std::make_pair(kernel, dev)];
    if (have < bytes) {
        CU_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
        have = bytes;
    }
    return PNP_OK;
}

// Kernel launch with the optional programmatic-stream-serialization attribute (the kernel may begin while its
// predecessor on the stream is finishing; it calls griddepcontrol.wait before it touches that kernel's output).
int launch_ex(const void* kernel, dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool chain, void** args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    if (chain) {
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at;
        cfg.numAttrs = 1;
    }
    CU_TRY(cudaLaunchKernelExC(&cfg, kernel, args));
    return PNP_OK;
}

template <int L>
int launch_r2c(const pnp_csmri_grad_args& a, cudaStream_t st) {
    constexpr int GP = lines_gp<L>();
    const int pairs = a.W / 2;
    const int items = (pairs + GP - 1) / GP;
    dim3 grid(persistent_ctas((const void*)pnp::k_lines_r2c<L, GP>, GP * pnp::fft_threads<L>(), lines_smem<L>(), items, a.batch), a.batch);
    pnp::SelJob sj{};
    if (a.sel_count > 0) {
        sj.bits = const_cast<unsigned char*>(a.bits);
        sj.idx = a.sel_idx; sj.idx_img_stride = a.sel_idx_img_stride; sj.cursor = a.sel_cursor;
        sj.support = a.sel_support; sj.m0 = a.sel_m0; sj.support_img_stride = a.sel_support_img_stride;
        sj.count = a.sel_count; sj.seed = a.sel_seed; sj.counter = a.sel_counter; sj.idx_out = nullptr;
    }
    const float* pa = a.a;
    const float* pb = a.b;
    float2* pS = reinterpret_cast<float2*>(a.S);
    int nl = a.W;
    long long stride = (long long)a.H * a.W;
    int rlo = a.row_hi > a.row_lo ? a.row_lo : 0, rhi = a.row_hi > a.row_lo ? a.row_hi : a.H / 2;
    void* args[] = {(void*)&pa, (void*)&pb, (void*)&pS, (void*)&nl, (void*)&stride, (void*)&sj, (void*)&rlo, (void*)&rhi};
    return launch_ex((const void*)pnp::k_lines_r2c<L, GP>, grid, dim3(GP * pnp::fft_threads<L>()), lines_smem<L>(), st,
                     (a.flags & PNP_FLAG_CHAIN) != 0, args);
}

template <int L>
int launch_cols(const pnp_csmri_grad_args& a, cudaStream_t st) {
    constexpr int NC = cols_nc<L>();
    const int hp = a.H / 2;
    int rlo = a.row_hi > a.row_lo ? a.row_lo : 0, rhi = a.row_hi > a.row_lo ? a.row_hi : hp;
    const int c_lo = rlo > 1 ? rlo : 1, c_hi = rhi < hp ? rhi : hp;
    const int items = c_hi > c_lo ? (c_hi - c_lo + NC - 1) / NC : 0;
    // resident CTAs: one of them (CTA 0) owns packed column 0, the others loop over the items
    int ctas = persistent_ctas((const void*)pnp::k_cols_mask<L, NC>, NC * pnp::fft_threads<L>(), cols_smem<L>(), items + 1, a.batch);
    if (ctas < 2) ctas = 2;
    dim3 grid(ctas, a.batch);
    float2* pS = reinterpret_cast<float2*>(a.S);
    const unsigned char* pbits = a.bits;
    const float2 *y1 = reinterpret_cast<const float2*>(a.Y1), *y2 = reinterpret_cast<const float2*>(a.Y2),
                 *y1n = reinterpret_cast<const float2*>(a.Y1n), *y2n = reinterpret_cast<const float2*>(a.Y2n);
    int hp_ = hp;
    long long bstride = (long long)a.W * hp, ystride = (long long)a.W * hp;
    unsigned char* clr = a.clear_bits ? const_cast<unsigned char*>(a.bits) : nullptr;
    void* args[] = {(void*)&pS, (void*)&pbits, (void*)&y1, (void*)&y2, (void*)&y1n, (void*)&y2n, (void*)&hp_, (void*)&bstride,
                    (void*)&ystride, (void*)&clr, (void*)&rlo, (void*)&rhi};
    return launch_ex((const void*)pnp::k_cols_mask<L, NC>, grid, dim3(NC * pnp::fft_threads<L>()), cols_smem<L>(), st,
                     (a.flags & PNP_FLAG_CHAIN) != 0, args);
}

template <int L>
int launch_c2r(const pnp_csmri_grad_args& a, cudaStream_t st) {
    constexpr int GP = lines_gp<L>();
    const int pairs = a.W / 2;
    const int items = (pairs + GP - 1) / GP;
    pnp::GradEpilogue ep{a.gscale, a.gscale_ptr, a.step, a.step_ptr, a.g_out, a.vadd, a.v_out, a.z_in, a.z_out};
    const float inv_n = (float)(1.0 / ((double)a.H * (double)a.W));
    const float2* pS = reinterpret_cast<const float2*>(a.S);
    int nl = a.W;
    long long stride = (long long)a.H * a.W;
    float inv = inv_n;
    int rlo = a.row_hi > a.row_lo ? a.row_lo : 0, rhi = a.row_hi > a.row_lo ? a.row_hi : a.H / 2;
    void* args[] = {(void*)&pS, (void*)&nl, (void*)&stride, (void*)&inv, (void*)&ep, (void*)&rlo, (void*)&rhi};
    const bool chain = (a.flags & PNP_FLAG_CHAIN) != 0;
    if (a.vadd && a.z_in && a.z_out && !a.g_out && !a.v_out) {          // the inner-iteration update
        dim3 grid(persistent_ctas((const void*)pnp::k_lines_c2r<L, GP, true>, GP * pnp::fft_threads<L>(), lines_smem<L>(), items, a.batch), a.batch);
        return launch_ex((const void*)pnp::k_lines_c2r<L, GP, true>, grid, dim3(GP * pnp::fft_threads<L>()), lines_smem<L>(), st, chain, args);
    }
    dim3 grid(persistent_ctas((const void*)pnp::k_lines_c2r<L, GP, false>, GP * pnp::fft_threads<L>(), lines_smem<L>(), items, a.batch), a.batch);
    return launch_ex((const void*)pnp::k_lines_c2r<L, GP, false>, grid, dim3(GP * pnp::fft_threads<L>()), lines_smem<L>(), st, chain, args);
}

#define DISPATCH_POW2(n, FN, ...)                                 \
    switch (n) {                                                  \
        case 32: return FN<32>(__VA_ARGS__);                      \
        case 64: return FN<64>(__VA_ARGS__);                      \
        case 128: return FN<128>(__VA_ARGS__);                    \
        case 256: return FN<256>(__VA_ARGS__);                    \
        case 512: return FN<512>(__VA_ARGS__);                    \
        case 1024: return FN<1024>(__VA_ARGS__);                  \
        case 2048: return FN<2048>(__VA_ARGS__);                  \
        case 4096: return FN<4096>(__VA_ARGS__);                  \
        default: return fail(PNP_ERR_ARG, "size %d is not a power of two in [32, 4096]", (int)(n)); \
    }

int dispatch_r2c(int n, const pnp_csmri_grad_args& a, cudaStream_t st) { DISPATCH_POW2(n, launch_r2c, a, st) }
int dispatch_cols(int n, const pnp_csmri_grad_args& a, cudaStream_t st) { DISPATCH_POW2(n, launch_cols, a, st) }
int dispatch_c2r(int n, const pnp_csmri_grad_args& a, cudaStream_t st) { DISPATCH_POW2(n, launch_c2r, a, st) }
int dispatch_attrs(int n) { DISPATCH_POW2(n, set_attrs) }
int dispatch_conv(int n, float2* S, const float2* Bf, const float2* twn, int H, int batch, int cj, cudaStream_t st) {
    DISPATCH_POW2(n, launch_conv, S, Bf, twn, H, batch, cj, st)
}

template <int L>
int launch_sigma(const float* z, int W, int batch, double* sig_log, const int* slot, cudaStream_t st) {
    dim3 grid((W + 3) / 4, batch);
    pnp::k_sigma_mad<L><<<grid, 128, 0, st>>>(z, W, (long long)L * W, sig_log, slot, batch);
    LAUNCH_CHECK();
    return PNP_OK;
}
int dispatch_sigma(int n, const float* z, int W, int batch, double* sig_log, const int* slot, cudaStream_t st) {
    DISPATCH_POW2(n, launch_sigma, z, W, batch, sig_log, slot, st)
}

template <int L>
int launch_haar(const float* zin, float* zout, const float* xrec, int W, int batch, pnp::ShrinkParams sp,
                double* mse_log, const int* slot, cudaStream_t st) {
    constexpr int WPL = pnp::HaarCfg<L>::WPL;           // warps per line
    constexpr int LPB = WPL >= 4 ? 1 : 4 / WPL;         // lines per CTA
    dim3 grid((W + LPB - 1) / LPB, batch);
    pnp::k_haar_bayes<L><<<grid, 32 * WPL * LPB, 0, st>>>(zin, zout, xrec, W, (long long)L * W, sp, mse_log, slot, batch);
    LAUNCH_CHECK();
    return PNP_OK;
}
int dispatch_haar(int n, const float* zin, float* zout, const float* xrec, int W, int batch, pnp::ShrinkParams sp,
                  double* mse_log, const int* slot, cudaStream_t st) {
    DISPATCH_POW2(n, launch_haar, zin, zout, xrec, W, batch, sp, mse_log, slot, st)
}

int ew_blocks(long long n, int per_thread) {
    long long b = (n / per_thread + 255) / 256;
    if (b < 1) b = 1;
    if (b > 148 * 8) b = 148 * 8;
    return (int)b;
}

}  // namespace

namespace {
template <int L>
int launch_update_prox(const float* S, int W, float inv_n, float gscale, float step, const float* step_ptr, const float* vadd,
                       const float* z_in, float* z_out, const float* xrec, float sm, float fb, double* sig_log, double* mse_log,
                       const int* slot, int* adv, int n_adv, unsigned* gbar, int chain, const pnp_csmri_next_pass* next, cudaStream_t st) {
    constexpr int GP = pnp::upd_gp<L>();
    const int npairs = W / 2;
    int grid = num_sms() < npairs ? num_sms() : npairs;
    const int ppc = (npairs + grid - 1) / grid;
    grid = (npairs + ppc - 1) / ppc;
    const size_t smem = sizeof(float) * ((size_t)pnp::lines_stage_off<L, GP>() + (size_t)2 * ppc * L);
    // worth it only when the rounds are reasonably full and everything fits next to the exchange planes
    if (smem > 220 * 1024 || 2 * ppc < GP) return fail(PNP_ERR_UNSUPPORTED, "update+prox: image does not suit the resident-line kernel");
    { const int rc = raise_smem_limit((const void*)pnp::k_update_prox<L>, 220 * 1024); if (rc != PNP_OK) return rc; }
    int nl = W, p = ppc;
    const float* w_next = nullptr;
    float2* S_next = nullptr;
    pnp::SelJob sj{};
    if (next) {
        if (2 * ppc > 16) return fail(PNP_ERR_UNSUPPORTED, "update+prox: the fused next line pass needs at most one line per warp");
        w_next = next->w;
        S_next = reinterpret_cast<float2*>(next->S_out);
        if (next->sel_count > 0) {
            sj.bits = next->bits;
            sj.idx = next->sel_idx; sj.idx_img_stride = 0; sj.cursor = nullptr;
            sj.support = next->sel_support; sj.m0 = next->sel_m0; sj.support_img_stride = 0;
            sj.count = next->sel_count; sj.seed = next->sel_seed; sj.counter = next->sel_counter; sj.idx_out = nullptr;
            sj.counter_add = next->sel_counter_add;
        }
    }
    void* args[] = {(void*)&S, (void*)&nl, (void*)&inv_n, (void*)&gscale, (void*)&step, (void*)&step_ptr, (void*)&vadd, (void*)&z_in,
                    (void*)&z_out, (void*)&xrec, (void*)&p, (void*)&sm, (void*)&fb, (void*)&sig_log, (void*)&mse_log, (void*)&slot,
                    (void*)&adv, (void*)&n_adv, (void*)&gbar, (void*)&w_next, (void*)&S_next, (void*)&sj};
    // with a barrier workspace: plain launch + software grid barrier (all CTAs are co-resident: grid <= SM count, one
    // CTA per SM), which is what lets the kernel join a programmatic-dependent-launch chain; else cooperative launch
    const void* kernel = next ? (const void*)pnp::k_update_prox<L, true> : (const void*)pnp::k_update_prox<L, false>;
    if (next) { const int rc = raise_smem_limit(kernel, 220 * 1024); if (rc != PNP_OK) return rc; }
    if (gbar) return launch_ex(kernel, dim3(grid), dim3(512), smem, st, chain != 0, args);
    CU_TRY(cudaLaunchCooperativeKernel(kernel, dim3(grid), dim3(512), args, smem, st));
    return PNP_OK;
}
int dispatch_update_prox(int n, const float* S, int W, float inv_n, float gscale, float step, const float* step_ptr,
                         const float* vadd, const float* z_in, float* z_out, const float* xrec, float sm, float fb, double* sig_log,
                         double* mse_log, const int* slot, int* adv, int n_adv, unsigned* gbar, int chain,
                         const pnp_csmri_next_pass* next, cudaStream_t st) {
    DISPATCH_POW2(n, launch_update_prox, S, W, inv_n, gscale, step, step_ptr, vadd, z_in, z_out, xrec, sm, fb, sig_log, mse_log, slot, adv, n_adv, gbar, chain, next, st)
}
}  // namespace

namespace {
template <int L>
int launch_cdp_lines_fwd(const float* u, const signed char* codes, float2* S, int nlines, int nmasks, const float* u2, float2* S2,
                         cudaStream_t st) {
    constexpr int GL = pnp::cdp_lines_per_cta<L>();
    dim3 grid((nlines + GL - 1) / GL, nmasks, S2 ? 2 : 1);
    pnp::k_cdp_lines_fwd<L><<<grid, GL * pnp::fft_threads<L>(), sizeof(float) * GL * 2 * pnp::fft_plane<L>(), st>>>(u, codes, S, nlines,
                                                                                                                   u2, S2);
    LAUNCH_CHECK();
    return PNP_OK;
}
template <int LW>
int launch_cdp_cols(float2* S, const float* y, unsigned char* mask, int H, int nmasks, float inv_n, int clear, float2* S2,
                    cudaStream_t st) {
    constexpr int NC = pnp::cdp_cols_per_cta<LW>();
    dim3 grid(H / NC, nmasks);
    pnp::k_cdp_cols<LW><<<grid, NC * pnp::fft_threads<LW>(), sizeof(float) * NC * 2 * pnp::fft_plane<LW>(), st>>>(S, y, mask, H, inv_n, clear,
                                                                                                                 S2);
    LAUNCH_CHECK();
    return PNP_OK;
}
template <int L>
int launch_cdp_lines_inv(const float2* S, const signed char* codes, float* acc, int nlines, int nmasks, float sign, int accumulate,
                         const float2* S2, cudaStream_t st) {
    constexpr int GL = pnp::cdp_lines_per_cta<L>();
    pnp::k_cdp_lines_inv<L><<<(nlines + GL - 1) / GL, GL * pnp::fft_threads<L>(), sizeof(float) * GL * 2 * pnp::fft_plane<L>(), st>>>(
        S, codes, acc, nlines, nmasks, sign, accumulate, S2);
    LAUNCH_CHECK();
    return PNP_OK;
}
int dispatch_cdp_lines_fwd(int n, const float* u, const signed char* codes, float2* S, int nlines, int nmasks, const float* u2,
                           float2* S2, cudaStream_t st) {
    DISPATCH_POW2(n, launch_cdp_lines_fwd, u, codes, S, nlines, nmasks, u2, S2, st)
}
int dispatch_cdp_cols(int n, float2* S, const float* y, unsigned char* mask, int H, int nmasks, float inv_n, int clear, float2* S2,
                      cudaStream_t st) {
    DISPATCH_POW2(n, launch_cdp_cols, S, y, mask, H, nmasks, inv_n, clear, S2, st)
}
int dispatch_cdp_lines_inv(int n, const float2* S, const signed char* codes, float* acc, int nlines, int nmasks, float sign,
                           int accumulate, const float2* S2, cudaStream_t st) {
    DISPATCH_POW2(n, launch_cdp_lines_inv, S, codes, acc, nlines, nmasks, sign, accumulate, S2, st)
}
}  // namespace

namespace {
template <int L>
int launch_prox_fused(const float* zin, float* zout, const float* xrec, int W, int batch, float sm, float fb, double* sig_log,
                      double* mse_log, const int* slot, cudaStream_t st) {
    const long long total = (long long)W * batch;
    int grid = num_sms();
    if (total < grid) grid = (int)total;
    int lpc = (int)((total + grid - 1) / grid);
    grid = (int)((total + lpc - 1) / lpc);
    const size_t smem = (size_t)lpc * L * sizeof(float);
    if (smem > 176 * 1024) return fail(PNP_ERR_UNSUPPORTED, "fused prox: %zu bytes of lines per CTA do not fit shared memory", smem);
    { const int rc = raise_smem_limit((const void*)pnp::k_prox_wavelet_fused<L>, 176 * 1024); if (rc != PNP_OK) return rc; }
    int nl = W;
    unsigned* gbar = nullptr;
    void* args[] = {(void*)&zin, (void*)&zout, (void*)&xrec, (void*)&nl, (void*)&batch, (void*)&lpc, (void*)&sm, (void*)&fb,
                    (void*)&sig_log, (void*)&mse_log, (void*)&slot, (void*)&gbar};
    CU_TRY(cudaLaunchCooperativeKernel((const void*)pnp::k_prox_wavelet_fused<L>, dim3(grid), dim3(512), args, smem, st));
    return PNP_OK;
}
int dispatch_prox_fused(int n, const float* zin, float* zout, const float* xrec, int W, int batch, float sm, float fb,
                        double* sig_log, double* mse_log, const int* slot, cudaStream_t st) {
    DISPATCH_POW2(n, launch_prox_fused, zin, zout, xrec, W, batch, sm, fb, sig_log, mse_log, slot, st)
}
}  // namespace

// ---- device-side construction of CSMRI problem batches (csrc/build.cuh) ----------------------------------------------
namespace {
struct BuildWork {                 // carve-up of the caller's scratch (all offsets 256-byte aligned)
    size_t S, im, Y1r, Y2r, Y1nr, Y2nr, chunks, norm2, minmax, total;
    BuildWork(int H, int W, int batch) {
        const size_t N = (size_t)H * W, nb = (size_t)batch;
        auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
        size_t o = 0;
        S = o;      o = al(o + nb * N * 4);                     // packed half spectrum (complex64 [H/2][W])
        im = o;     o = al(o + nb * N * 4);
        Y1r = o;    o = al(o + nb * N * 4);
        Y2r = o;    o = al(o + nb * N * 4);
        Y1nr = o;   o = al(o + nb * (size_t)W * 8);
        Y2nr = o;   o = al(o + nb * (size_t)W * 8);
        chunks = o; o = al(o + nb * ((N + BUILD_CHUNK - 1) / BUILD_CHUNK) * 4);
        norm2 = o;  o = al(o + nb * 8);
        minmax = o; o = al(o + nb * 8);
        total = o;
    }
};

template <int L>
int launch_cols_build(const float2* S, const unsigned char* bits, float2* Y1, float2* Y1n, int hp, int batch, double* norm2,
                      cudaStream_t st) {
    constexpr int NC = cols_nc<L>();
    dim3 grid((hp + NC - 1) / NC, batch);
    pnp::k_cols_build<L, NC><<<grid, NC * pnp::fft_threads<L>(), conv_smem<L>(), st>>>(S, bits, Y1, Y1n, hp, norm2);
    LAUNCH_CHECK();
    return PNP_OK;
}
int dispatch_cols_build(int n, const float2* S, const unsigned char* bits, float2* Y1, float2* Y1n, int hp, int batch,
                        double* norm2, cudaStream_t st) {
    DISPATCH_POW2(n, launch_cols_build, S, bits, Y1, Y1n, hp, batch, norm2, st)
}
}  // namespace

extern "C" {

int pnp_version(void) { return 100; }
const char* pnp_last_error(void) { return g_err; }

int pnp_init(void) {
    int dev = 0;
    CU_TRY(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return fail(PNP_ERR_ARG, "device index %d out of range", dev);
    if (g_init[dev]) return PNP_OK;
    CU_TRY(cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev));
    g_sms_of[dev] = g_num_sms;
    std::vector<float2> tw(PNP_TW_N);
    for (int m = 0; m < PNP_TW_N; ++m) {
        const double ang = -2.0 * M_PI * (double)m / (double)PNP_TW_N;
        tw[m] = make_float2((float)std::cos(ang), (float)std::sin(ang));
    }
    CU_TRY(cudaMemcpyToSymbol(g_tw, tw.data(), sizeof(float2) * PNP_TW_N));
    for (int n = 32; n <= 4096; n *= 2) {
        const int rc = dispatch_attrs(n);
        if (rc != PNP_OK) return rc;
    }
    {
        cudaFuncAttributes fa;
        CU_TRY(cudaFuncGetAttributes(&fa, pnp::k_sel_from_indices));
        CU_TRY(cudaFuncGetAttributes(&fa, pnp::k_sel_from_feistel));
        CU_TRY(cudaFuncGetAttributes(&fa, pnp::k_sq_err));
        CU_TRY(cudaFuncGetAttributes(&fa, pnp::k_axpy));
        CU_TRY(cudaFuncGetAttributes(&fa, pnp::k_saga_update));
        CU_TRY(cudaFuncGetAttributes(&fa, pnp::k_saga_init));
        CU_TRY(cudaFuncGetAttributes(&fa, pnp::k_advance));
        CU_TRY(cudaFuncGetAttributes(&fa, pnp::k_bilinear_residual));
        CU_TRY(cudaFuncGetAttributes(&fa, pnp::k_pr_rows));
        CU_TRY(cudaFuncGetAttributes(&fa, pnp::k_pr_cols));
        CU_TRY(cudaFuncGetAttributes(&fa, pnp::k_nlm));
        CU_TRY(cudaFuncGetAttributes(&fa, pnp::k_nlm5));
        CU_TRY(cudaFuncGetAttributes(&fa, pnp::k_sample_indices));
        CU_TRY(cudaFuncGetAttributes(&fa, pnp::k_minmax_init));
        CU_TRY(cudaFuncGetAttributes(&fa, pnp::k_minmax));
        CU_TRY(cudaFuncGetAttributes(&fa, pnp::k_conv_first));
        CU_TRY(cudaFuncGetAttributes(&fa, pnp::k_conv_mid));
        CU_TRY(cudaFuncGetAttributes(&fa, pnp::k_conv_last));
    }
    g_init[dev] = true;
    return PNP_OK;
}

int pnp_csmri_grad(const pnp_csmri_grad_args* args, void* stream) {
    if (!args) return fail(PNP_ERR_ARG, "null args");
    const pnp_csmri_grad_args& a = *args;
    if (!pow2_ok(a.H) || !pow2_ok(a.W)) return fail(PNP_ERR_ARG, "H=%d W=%d must be powers of two in [32, 4096]", a.H, a.W);
    if (a.batch < 1 || a.batch > 65535) return fail(PNP_ERR_ARG, "batch=%d out of range", a.batch);
    // S may be null for the update pass alone (phases == 4 with vadd, z_in, z_out and nothing else written): the gradient
    // term is then exactly zero (first inner iteration of an SVRG epoch, see pnp_b200.h)
    const bool upd_only = a.phases == 4 && a.vadd && a.z_in && a.z_out && !a.g_out && !a.v_out;
    if (!a.a || !a.bits || (!a.S && !upd_only)) return fail(PNP_ERR_ARG, "a, S and bits must be non-null");
    if ((a.Y1 == nullptr) != (a.Y2 == nullptr) || (a.Y1 == nullptr) != (a.Y1n == nullptr) ||
        (a.Y1 == nullptr) != (a.Y2n == nullptr))
        return fail(PNP_ERR_ARG, "Y1, Y2, Y1n, Y2n must be all null or all non-null");
    if ((a.z_out != nullptr) && !a.z_in) return fail(PNP_ERR_ARG, "z_out needs z_in");
    const int ph = a.phases ? a.phases : 7;
    if (a.sel_count < 0) return fail(PNP_ERR_ARG, "sel_count < 0");
    if (a.row_hi > a.row_lo && (a.row_lo < 0 || a.row_hi > a.H / 2)) return fail(PNP_ERR_ARG, "row range outside [0, H/2]");
    if (a.sel_count > 0) {
        if (!(ph & 1)) return fail(PNP_ERR_ARG, "in-pass selection needs pass 1 (phases bit0)");
        if (!a.sel_idx && (!a.sel_support || !a.sel_m0)) return fail(PNP_ERR_ARG, "in-pass selection needs sel_idx or sel_support + sel_m0");
        // the sampler's cycle walk only terminates inside [0, m0): refuse oversize minibatches on the host
        // (the reference: problems/CSMRI.py:68-69 prints a warning and np.random.choice raises ValueError)
        if (!a.sel_idx && a.sel_count > a.sel_min_m0)
            return fail(PNP_ERR_ARG, "minibatch of %d exceeds the %d sampled positions", a.sel_count, a.sel_min_m0);
    }
    int rc = check_init();
    if (rc != PNP_OK) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if ((ph & 1) && (rc = dispatch_r2c(a.H, a, st)) != PNP_OK) return rc;
    if ((ph & 2) && (rc = dispatch_cols(a.W, a, st)) != PNP_OK) return rc;
    if (ph & 4) return dispatch_c2r(a.H, a, st);
    return PNP_OK;
}

long long pnp_csmri_build_batch_workspace(int H, int W, int batch) {
    if (!pow2_ok(H) || !pow2_ok(W) || batch < 1) return -1;
    return (long long)BuildWork(H, W, batch).total;
}

int pnp_csmri_build_batch(const pnp_csmri_build_args* args, void* stream) {
    if (!args) return fail(PNP_ERR_ARG, "null args");
    const pnp_csmri_build_args& a = *args;
    if (!pow2_ok(a.H) || !pow2_ok(a.W)) return fail(PNP_ERR_ARG, "H=%d W=%d must be powers of two in [32, 4096]", a.H, a.W);
    if (a.batch < 1 || a.batch > 65535) return fail(PNP_ERR_ARG, "batch=%d out of range", a.batch);
    if (!a.x || !a.p || !a.snr || !a.bits_full || !a.m0 || !a.support || !a.Y1 || !a.Y2 || !a.Y1n || !a.Y2n || !a.xinit ||
        !a.sigma || !a.work)
        return fail(PNP_ERR_ARG, "null pointer");
    const long long N = (long long)a.H * a.W;
    // a list shorter than the plane only fits when M0 <= stride, and M0 is only known on the device
    if (a.support_img_stride < N) return fail(PNP_ERR_ARG, "support_img_stride must be >= H*W");
    if ((reinterpret_cast<unsigned long long>(a.work) & 255ull) != 0) return fail(PNP_ERR_ARG, "work must be 256-byte aligned");
    int rc = check_init();
    if (rc != PNP_OK) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const BuildWork w(a.H, a.W, a.batch);
    char* wk = static_cast<char*>(a.work);
    float* S = reinterpret_cast<float*>(wk + w.S);
    float* im = reinterpret_cast<float*>(wk + w.im);
    float2 *Y1r = reinterpret_cast<float2*>(wk + w.Y1r), *Y2r = reinterpret_cast<float2*>(wk + w.Y2r);
    float2 *Y1nr = reinterpret_cast<float2*>(wk + w.Y1nr), *Y2nr = reinterpret_cast<float2*>(wk + w.Y2nr);
    int* chunks = reinterpret_cast<int*>(wk + w.chunks);
    double* norm2 = reinterpret_cast<double*>(wk + w.norm2);
    unsigned* minmax = reinterpret_cast<unsigned*>(wk + w.minmax);
    const int hp = a.H / 2, nchunks = (int)((N + BUILD_CHUNK - 1) / BUILD_CHUNK);
    const int eb = (int)std::min<long long>((hp * (long long)a.W + 255) / 256, 1184);

    pnp::k_build_init<<<(a.batch + 255) / 256, 256, 0, st>>>(norm2, minmax, a.batch);
    LAUNCH_CHECK();
    pnp::k_build_bits<<<dim3(eb, a.batch), 256, 0, st>>>(a.bits_full, a.H, a.W, a.seed, a.p);
    LAUNCH_CHECK();
    // support lists (np.flatnonzero order) + M0
    pnp::k_support_count<<<dim3(nchunks, a.batch), 256, 0, st>>>(chunks, (int)N, a.seed, a.p);
    LAUNCH_CHECK();
    pnp::k_support_scan<<<a.batch, 1024, 0, st>>>(chunks, nchunks, a.m0, a.inv_m0);
    LAUNCH_CHECK();
    pnp::k_support_write<<<dim3(nchunks, a.batch), 256, 0, st>>>(a.support, a.support_img_stride, chunks, (int)N, a.seed, a.p);
    LAUNCH_CHECK();
    // F = fft2(X): the iteration's forward line pass, then the transform across the lines of every packed row
    pnp_csmri_grad_args g{};
    g.H = a.H; g.W = a.W; g.batch = a.batch; g.a = a.x; g.S = S; g.bits = a.bits_full; g.gscale = 1.f;
    if ((rc = dispatch_r2c(a.H, g, st)) != PNP_OK) return rc;
    if ((rc = dispatch_cols_build(a.W, reinterpret_cast<const float2*>(S), a.bits_full, reinterpret_cast<float2*>(a.Y1),
                                  reinterpret_cast<float2*>(a.Y1n), hp, a.batch, norm2, st)) != PNP_OK) return rc;
    pnp::k_build_noise<<<dim3(eb, a.batch), 256, 0, st>>>(
        reinterpret_cast<float2*>(a.Y1), reinterpret_cast<float2*>(a.Y2), reinterpret_cast<float2*>(a.Y1n),
        reinterpret_cast<float2*>(a.Y2n), Y1r, Y2r, Y1nr, Y2nr, a.bits_full, a.H, a.W, a.seed, a.snr, norm2, a.sigma);
    LAUNCH_CHECK();
    // Re ifft2(Y) and Im ifft2(Y) = Re ifft2(-iY): the iteration's column + inverse line passes on a ZERO spectrum with the
    // measurement term (they return -Re ifft2 of the Hermitian part of the measurements; the sign drops out of |.|)
    for (int part = 0; part < 2; ++part) {
        CU_TRY(cudaMemsetAsync(S, 0, (size_t)a.batch * N * 4, st));
        g.Y1 = part ? reinterpret_cast<const float*>(Y1r) : a.Y1;
        g.Y2 = part ? reinterpret_cast<const float*>(Y2r) : a.Y2;
        g.Y1n = part ? reinterpret_cast<const float*>(Y1nr) : a.Y1n;
        g.Y2n = part ? reinterpret_cast<const float*>(Y2nr) : a.Y2n;
        g.g_out = part ? im : a.xinit;
        if ((rc = dispatch_cols(a.W, g, st)) != PNP_OK) return rc;
        if ((rc = dispatch_c2r(a.H, g, st)) != PNP_OK) return rc;
    }
    const int xb = (int)std::min<long long>((N + 255) / 256, 1184);
    pnp::k_build_abs<<<dim3(xb, a.batch), 256, 0, st>>>(a.xinit, im, N, minmax);
    LAUNCH_CHECK();
    pnp::k_build_norm<<<dim3(xb, a.batch), 256, 0, st>>>(a.xinit, N, minmax);
    LAUNCH_CHECK();
    return PNP_OK;
}

int pnp_csmri_sel_from_indices(unsigned char* bits, int H, int W, int batch, const int* idx, int count,
                               long long idx_img_stride, const int* cursor, int clear, void* stream) {
    if (!bits || !idx || !pow2_ok(H) || !pow2_ok(W) || batch < 1 || count < 0) return fail(PNP_ERR_ARG, "bad argument");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (clear) CU_TRY(cudaMemsetAsync(bits, 0, (size_t)batch * W * (H / 2), st));
    if (count == 0) return PNP_OK;
    int blocks = (count + 255) / 256;
    if (blocks > 592) blocks = 592;
    pnp::k_sel_from_indices<<<dim3(blocks, batch), 256, 0, st>>>(bits, H, W, idx, count, idx_img_stride, cursor);
    LAUNCH_CHECK();
    return PNP_OK;
}

int pnp_csmri_sel_sample(unsigned char* bits, int H, int W, int batch, const int* support, const int* m0,
                         long long support_img_stride, int count, unsigned seed, const int* counter,
                         int* idx_out, int clear, void* stream) {
    if (!bits || !support || !m0 || !pow2_ok(H) || !pow2_ok(W) || batch < 1 || count < 1) return fail(PNP_ERR_ARG, "bad argument");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (clear) CU_TRY(cudaMemsetAsync(bits, 0, (size_t)batch * W * (H / 2), st));
    int blocks = (count + 255) / 256;
    if (blocks > 592) blocks = 592;
    pnp::k_sel_from_feistel<<<dim3(blocks, batch), 256, 0, st>>>(bits, H, W, support, m0, support_img_stride, count,
                                                                  seed, counter, idx_out);
    LAUNCH_CHECK();
    return PNP_OK;
}

int pnp_estimate_sigma(const float* z, int H, int W, int batch, double* sig_log, const int* slot, void* stream) {
    if (!z || !sig_log || batch < 1) return fail(PNP_ERR_ARG, "bad argument");
    return dispatch_sigma(H, z, W, batch, sig_log, slot, static_cast<cudaStream_t>(stream));
}

int pnp_wavelet_denoise(const float* z_in, float* z_out, int H, int W, int batch, const double* sig_log,
                        float sigma_est, float sigma_modifier, float fallback_sigma, const float* xrec,
                        double* mse_log, const int* slot, void* stream) {
    if (!z_in || !z_out || batch < 1) return fail(PNP_ERR_ARG, "bad argument");
    pnp::ShrinkParams sp{sig_log, sigma_est, sigma_modifier, fallback_sigma};
    return dispatch_haar(H, z_in, z_out, xrec, W, batch, sp, mse_log, slot, static_cast<cudaStream_t>(stream));
}

int pnp_prox_wavelet_fused(const float* z_in, float* z_out, int H, int W, int batch, double* sig_log, float sigma_modifier,
                           float fallback_sigma, const float* xrec, double* mse_log, const int* slot, void* stream) {
    if (!z_in || !z_out || !sig_log || batch < 1) return fail(PNP_ERR_ARG, "bad argument");
    return dispatch_prox_fused(H, z_in, z_out, xrec, W, batch, sigma_modifier, fallback_sigma, sig_log, mse_log, slot,
                               static_cast<cudaStream_t>(stream));
}

int pnp_sq_err(const float* z, const float* xrec, long long n, int batch, double* out, const int* slot, void* stream) {
    if (!z || !xrec || !out || n < 1 || batch < 1) return fail(PNP_ERR_ARG, "bad argument");
    pnp::k_sq_err<<<dim3(ew_blocks(n, 4), batch), 256, 0, static_cast<cudaStream_t>(stream)>>>(z, xrec, n, n, out, slot, batch);
    LAUNCH_CHECK();
    return PNP_OK;
}

int pnp_axpy(const float* z_in, const float* v, float* z_out, long long n, int batch, float step,
             const float* step_ptr, void* stream) {
    if (!z_in || !v || !z_out || n < 4 || (n & 3) || batch < 1) return fail(PNP_ERR_ARG, "bad argument (n must be a multiple of 4)");
    pnp::k_axpy<<<dim3(ew_blocks(n, 4), batch), 256, 0, static_cast<cudaStream_t>(stream)>>>(z_in, v, z_out, n, n, step, step_ptr);
    LAUNCH_CHECK();
    return PNP_OK;
}

int pnp_saga_init(const float* g0, float* table, float* tsum, long long n, int batch, int hist, void* stream) {
    if (!g0 || !table || !tsum || n < 1 || batch < 1 || hist < 1) return fail(PNP_ERR_ARG, "bad argument");
    pnp::k_saga_init<<<dim3(ew_blocks(n, 1), batch), 256, 0, static_cast<cudaStream_t>(stream)>>>(g0, table, tsum, n, n, hist);
    LAUNCH_CHECK();
    return PNP_OK;
}

int pnp_saga_update(const float* g_new, float* g_prev, float* table, float* tsum, float* z, long long n,
                    int batch, int hist, const int* slot_idx, long long slot_img_stride, const int* cursor,
                    float step, const float* step_ptr, void* stream) {
    if (!g_new || !g_prev || !table || !tsum || !z || !slot_idx || n < 1 || batch < 1 || hist < 1)
        return fail(PNP_ERR_ARG, "bad argument");
    pnp::k_saga_update<<<dim3(ew_blocks(n, 4), batch), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        g_new, g_prev, table, tsum, z, n, n, hist, slot_idx, slot_img_stride, cursor, step, step_ptr);
    LAUNCH_CHECK();
    return PNP_OK;
}

int pnp_advance(int* counters, int n, void* stream) {
    if (!counters || n < 1 || n > 32) return fail(PNP_ERR_ARG, "bad argument");
    pnp::k_advance<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(counters, n, nullptr, 1.0f);
    LAUNCH_CHECK();
    return PNP_OK;
}

int pnp_advance_scale(int* counters, int n, float* x, float factor, void* stream) {
    if (!counters || n < 0 || n > 32) return fail(PNP_ERR_ARG, "bad argument");
    pnp::k_advance<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(counters, n, x, factor);
    LAUNCH_CHECK();
    return PNP_OK;
}

int pnp_copy_f32(float* dst, const float* src, long long n, void* stream) {
    if (!dst || !src || n < 0) return fail(PNP_ERR_ARG, "bad argument");
    CU_TRY(cudaMemcpyAsync(dst, src, sizeof(float) * (size_t)n, cudaMemcpyDeviceToDevice, static_cast<cudaStream_t>(stream)));
    return PNP_OK;
}

int pnp_sample_indices(int* idx_out, int n, int count, unsigned seed, const int* counter, void* stream) {
    if (!idx_out || n < 1 || count < 1 || count > n) return fail(PNP_ERR_ARG, "bad argument");
    int blocks = (count + 255) / 256;
    if (blocks > 592) blocks = 592;
    pnp::k_sample_indices<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(idx_out, n, count, seed, counter);
    LAUNCH_CHECK();
    return PNP_OK;
}

int pnp_sample_indices_host(int* out, int n, int count, unsigned seed, unsigned counter, int img, int threads,
                            const int* support_host) {
    if (!out || n < 1 || count < 1 || count > n) return fail(PNP_ERR_ARG, "bad argument");
    const unsigned key = pnp_host::mix32(seed ^ pnp_host::mix32(counter * 0x632be5abU + (unsigned)img));
    int hb = 1;
    while ((1u << (2 * hb)) < (unsigned)n) ++hb;
    if (threads < 1) threads = 1;
    if (threads > 16) threads = 16;
    if (count < 4096) threads = 1;
    auto work = [=](int lo, int hi) { pnp_host::sample_range(out, lo, hi, (unsigned)n, key, hb, support_host); };
    if (threads == 1) { work(0, count); return PNP_OK; }
    std::vector<std::thread> pool;
    const int per = (count + threads - 1) / threads;
    for (int t = 0; t < threads; ++t) {
        const int lo = t * per, hi = lo + per < count ? lo + per : count;
        if (lo < hi) pool.emplace_back(work, lo, hi);
    }
    for (auto& th : pool) th.join();
    return PNP_OK;
}

// ---- look-ahead host draws (mb_source='host') ---------------------------------------------------------------
struct pnp_host_draws {
    pnp_host::DrawQueue q;
    std::vector<cudaEvent_t> ev;          // ev[slot]: recorded after the H2D copy that reads buffers[slot]
    std::vector<char> recorded;
    std::vector<int> guard;               // guard[slot]: the event recorded after the copy that read buffers[slot] (stage_many: one per run)
    const int* support_dev = nullptr;     // set: the draws are RANKS into this device-resident list, resolved on the device after each copy
    pnp_host_draws(int n, int count, unsigned seed, const int* support, int* const* buffers, int n_buffers, int ahead)
        : q(n, count, seed, support, buffers, n_buffers, ahead), recorded(n_buffers, 0), guard(n_buffers, 0) {
        for (int i = 0; i < n_buffers; ++i) guard[i] = i;
    }
};

int pnp_host_draws_create(pnp_host_draws** out, int n, int count, unsigned seed, const int* support_host,
                          int* const* buffers, int n_buffers, int ahead) {
    if (!out || !buffers || n < 1 || count < 1 || count > n) return fail(PNP_ERR_ARG, "bad argument");
    if (ahead < 1 || ahead > 64 || n_buffers < ahead + 2)
        return fail(PNP_ERR_ARG, "need 1 <= ahead <= 64 and at least ahead + 2 buffers (ahead=%d, buffers=%d)", ahead, n_buffers);
    for (int i = 0; i < n_buffers; ++i)
        if (!buffers[i]) return fail(PNP_ERR_ARG, "null buffer %d", i);
    *out = new pnp_host_draws(n, count, seed, support_host, buffers, n_buffers, ahead);
    return PNP_OK;
}

int pnp_host_draws_set_device_support(pnp_host_draws* h, const int* support_dev) {
    if (!h) return fail(PNP_ERR_ARG, "null handle");
    if (support_dev && h->q.has_support()) return fail(PNP_ERR_ARG, "the queue already gathers through a host support list");
    h->support_dev = support_dev;
    return PNP_OK;
}

// ranks -> positions behind the copy that brought them (same stream, before the event that frees the staging buffer)
static int gather_support(const pnp_host_draws* h, int* dst, int rows, long long stride, cudaStream_t st) {
    if (!h->support_dev) return PNP_OK;
    const int count = h->q.count();
    dim3 grid(std::min((count + 255) / 256, 64), rows);
    pnp::k_gather_support<<<grid, 256, 0, st>>>(dst, h->support_dev, count, stride);
    LAUNCH_CHECK();
    return PNP_OK;
}

int pnp_host_draws_next(pnp_host_draws* h, int* slot) {
    if (!h || !slot) return fail(PNP_ERR_ARG, "null argument");
    *slot = h->q.wait_next();
    h->q.release();
    return PNP_OK;
}

int pnp_host_draws_stage(pnp_host_draws* h, int* dst_dev, const int* extras, int n_extras, void* stream, int* slot_out) {
    if (!h || !dst_dev || n_extras < 0 || (n_extras > 0 && !extras)) return fail(PNP_ERR_ARG, "bad argument");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int R = h->q.n_buffers();
    if (h->ev.empty()) {                  // first staging call: one event per buffer, on the caller's device
        h->ev.resize(R, nullptr);
        // blocking sync: when the loop is GPU-bound the host waits here most of the time, and a spinning wait would take
        // a core from the sampler threads (eight ranks share the host's cores); the ring absorbs the wake-up latency
        for (int i = 0; i < R; ++i)
            CU_TRY(cudaEventCreateWithFlags(&h->ev[i], cudaEventDisableTiming | cudaEventBlockingSync));
    }
    const long long c = h->q.consumed();
    const int slot = h->q.wait_next();
    int* buf = h->q.buffer(slot);
    for (int i = 0; i < n_extras; ++i) buf[h->q.count() + i] = extras[i];
    CU_TRY(cudaMemcpyAsync(dst_dev, buf, sizeof(int) * (size_t)(h->q.count() + n_extras), cudaMemcpyHostToDevice, st));
    { const int rc = gather_support(h, dst_dev, 1, 0, st); if (rc != PNP_OK) return rc; }
    CU_TRY(cudaEventRecord(h->ev[slot], st));
    h->recorded[slot] = 1;
    h->guard[slot] = slot;
    // releasing draw c lets a worker start draw c + ahead: the copy that read that draw's buffer must have finished
    const int s2 = (int)((c + h->q.ahead()) % R);
    if (h->recorded[s2] && s2 != slot) CU_TRY(cudaEventSynchronize(h->ev[h->guard[s2]]));
    h->q.release();
    if (slot_out) *slot_out = slot;
    return PNP_OK;
}

int pnp_host_draws_stage_many(pnp_host_draws* h, int* dst_dev, int n_draws, long long dst_stride, void* stream) {
    if (!h || !dst_dev || n_draws < 1 || dst_stride < h->q.count()) return fail(PNP_ERR_ARG, "bad argument");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int R = h->q.n_buffers();
    if (h->ev.empty()) {
        h->ev.resize(R, nullptr);
        for (int i = 0; i < R; ++i)
            CU_TRY(cudaEventCreateWithFlags(&h->ev[i], cudaEventDisableTiming | cudaEventBlockingSync));
    }
    // Consecutive draws whose staging buffers are consecutive rows of one allocation with the destination's row length go
    // out as ONE copy and one event (an epoch of T2 small minibatches: 1-2 copies instead of T2).  A run never grows past
    // R - ahead - 1 draws: the buffer of a draw in it must not become claimable before the run's copy is enqueued.
    const int max_run = R - h->q.ahead() - 1 > 1 ? R - h->q.ahead() - 1 : 1;
    int run_slot = -1, run_len = 0;
    long long done = 0;
    auto flush = [&]() -> int {
        if (run_len == 0) return PNP_OK;
        const size_t bytes = run_len == 1 ? sizeof(int) * (size_t)h->q.count() : sizeof(int) * (size_t)dst_stride * (size_t)run_len;
        CU_TRY(cudaMemcpyAsync(dst_dev + done * dst_stride, h->q.buffer(run_slot), bytes, cudaMemcpyHostToDevice, st));
        { const int rc = gather_support(h, dst_dev + done * dst_stride, run_len, dst_stride, st); if (rc != PNP_OK) return rc; }
        const int last = run_slot + run_len - 1;
        CU_TRY(cudaEventRecord(h->ev[last], st));
        for (int s = run_slot; s <= last; ++s) { h->recorded[s] = 1; h->guard[s] = last; }
        done += run_len;
        run_len = 0;
        return PNP_OK;
    };
    for (int j = 0; j < n_draws; ++j) {
        const long long c = h->q.consumed();
        const int slot = h->q.wait_next();
        const bool joins = run_len > 0 && run_len < max_run && slot == run_slot + run_len &&
                           h->q.buffer(slot) == h->q.buffer(run_slot) + (long long)run_len * dst_stride;
        if (!joins) {
            const int rc = flush();
            if (rc != PNP_OK) return rc;
            run_slot = slot;
            run_len = 1;
        } else {
            ++run_len;
        }
        const int s2 = (int)((c + h->q.ahead()) % R);
        if (h->recorded[s2] && !(s2 >= run_slot && s2 < run_slot + run_len)) CU_TRY(cudaEventSynchronize(h->ev[h->guard[s2]]));
        h->q.release();
    }
    return flush();
}

int pnp_host_draws_destroy(pnp_host_draws* h) {
    if (!h) return PNP_OK;
    for (cudaEvent_t e : h->ev)
        if (e) cudaEventDestroy(e);
    delete h;                             // joins the workers
    return PNP_OK;
}

int pnp_deblur_grad(const pnp_deblur_grad_args* args, void* stream) {
    if (!args) return fail(PNP_ERR_ARG, "null args");
    const pnp_deblur_grad_args& a = *args;
    if (!pow2_ok(a.H) || !pow2_ok(a.W)) return fail(PNP_ERR_ARG, "H=%d W=%d must be powers of two in [32, 4096]", a.H, a.W);
    if (a.batch != 1) return fail(PNP_ERR_ARG, "pnp_deblur_grad: batch must be 1 in this revision");
    if (!a.a || !a.blurred || !a.up || !a.y) return fail(PNP_ERR_ARG, "null pointer");
    if (a.ntaps <= 0 && (!a.S || !a.Bf || !a.twn)) return fail(PNP_ERR_ARG, "null pointer (FFT path needs S, Bf, twn)");
    if (!a.identity && (!a.tl || !a.wts)) return fail(PNP_ERR_ARG, "bilinear tables missing");
    if (a.z_out && !a.z_in) return fail(PNP_ERR_ARG, "z_out needs z_in");
    int rc = check_init();
    if (rc != PNP_OK) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const long long N = (long long)a.H * a.W;
    pnp_csmri_grad_args f{};
    f.H = a.H; f.W = a.W; f.batch = 1; f.S = a.S;
    const bool direct = a.ntaps > 0;
    pnp::TapList taps{};
    if (direct) {
        if (a.ntaps > PNP_MAX_TAPS || !a.tap_pos || !a.tap_w) return fail(PNP_ERR_ARG, "ntaps must be <= %d with both tap arrays", PNP_MAX_TAPS);
        taps.n = a.ntaps;
        for (int j = 0; j < a.ntaps; ++j) {
            if (a.tap_pos[j] < 0 || a.tap_pos[j] >= N) return fail(PNP_ERR_ARG, "tap position out of range");
            taps.pr[j] = a.tap_pos[j] / a.W;
            taps.pc[j] = a.tap_pos[j] % a.W;
            taps.w[j] = a.tap_w[j];
        }
    }
    // x = fft_blur(a - b, B)
    if (direct) {
        pnp::k_tap_conv<<<ew_blocks(N, 1), 256, 0, st>>>(a.a, a.b, a.H, a.W, taps, 1, 1.0f, 0.f, nullptr, a.blurred, nullptr, nullptr,
                                                         nullptr, nullptr);
        LAUNCH_CHECK();
    } else {
        f.a = a.a; f.b = a.b; f.gscale = 1.0f; f.g_out = a.blurred;
        if ((rc = dispatch_r2c(a.H, f, st)) != PNP_OK) return rc;
        if ((rc = dispatch_conv(a.W, reinterpret_cast<float2*>(a.S), reinterpret_cast<const float2*>(a.Bf),
                                reinterpret_cast<const float2*>(a.twn), a.H, 1, 0, st)) != PNP_OK) return rc;
        if ((rc = dispatch_c2r(a.H, f, st)) != PNP_OK) return rc;
    }
    // up = S^T (S x - y) on the selection
    CU_TRY(cudaMemsetAsync(a.up, 0, sizeof(float) * (size_t)N, st));
    const int count = a.sel ? a.count : a.M;
    if (count > 0) {
        int blocks = (count + 255) / 256;
        if (blocks > 1184) blocks = 1184;
        pnp::k_bilinear_residual<<<dim3(blocks, 1), 256, 0, st>>>(a.blurred, a.up, a.y, a.tl, a.wts, a.sel, count, a.H,
                                                                  a.identity, a.use_y, N, a.M, 0, a.cursor);
        LAUNCH_CHECK();
    }
    // g = fft_blur(up, roll(flip(B), 1)) -> epilogue
    if (direct) {
        pnp::k_tap_conv<<<ew_blocks(N, 1), 256, 0, st>>>(a.up, nullptr, a.H, a.W, taps, -1, a.gscale, a.step, a.step_ptr, a.g_out,
                                                         a.vadd, a.v_out, a.z_in, a.z_out);
        LAUNCH_CHECK();
        return PNP_OK;
    }
    f.a = a.up; f.b = nullptr; f.gscale = a.gscale; f.step = a.step; f.step_ptr = a.step_ptr;
    f.g_out = a.g_out; f.vadd = a.vadd; f.v_out = a.v_out; f.z_in = a.z_in; f.z_out = a.z_out;
    if ((rc = dispatch_r2c(a.H, f, st)) != PNP_OK) return rc;
    if ((rc = dispatch_conv(a.W, reinterpret_cast<float2*>(a.S), reinterpret_cast<const float2*>(a.Bf),
                            reinterpret_cast<const float2*>(a.twn), a.H, 1, 1, st)) != PNP_OK) return rc;
    return dispatch_c2r(a.H, f, st);
}

int pnp_pr_grad(const pnp_pr_grad_args* args, void* stream) {
    if (!args) return fail(PNP_ERR_ARG, "null args");
    const pnp_pr_grad_args& a = *args;
    if (!a.A || !a.z || !a.y || !a.r || a.n < 4 || (a.n & 3) || a.M < 1) return fail(PNP_ERR_ARG, "bad argument");
    if (a.z_out && !a.z_in) return fail(PNP_ERR_ARG, "z_out needs z_in");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int count = a.rows ? a.count : a.M;
    if (count < 1) return fail(PNP_ERR_ARG, "empty selection");
    int blocks = count < 148 * 8 ? count : 148 * 8;
    pnp::k_pr_rows<<<blocks, 256, 0, st>>>(a.A, a.z, a.w, a.y, a.rows, count, a.n, a.r, a.cursor);
    LAUNCH_CHECK();
    const long long n4 = a.n / 4;
    if (a.partial && a.partial_chunks >= 1) {
        int chunks = a.partial_chunks < count ? a.partial_chunks : count;
        dim3 grid((unsigned)((n4 + 127) / 128), (unsigned)chunks);
        pnp::k_pr_cols_split<<<grid, 128, 0, st>>>(a.A, a.r, a.rows, count, a.n, a.cursor, a.partial);
        LAUNCH_CHECK();
        pnp::k_pr_cols_finish<<<(unsigned)((n4 + 31) / 32), 256, 0, st>>>(a.partial, chunks, a.n, a.gscale, a.step, a.step_ptr,
                                                                            a.g_out, a.vadd, a.v_out, a.z_in, a.z_out);
        LAUNCH_CHECK();
        return PNP_OK;
    }
    pnp::k_pr_cols<<<(unsigned)((n4 + 255) / 256), 256, 0, st>>>(a.A, a.r, a.rows, count, a.n, a.cursor, a.gscale, a.step,
                                                                 a.step_ptr, a.g_out, a.vadd, a.v_out, a.z_in, a.z_out);
    LAUNCH_CHECK();
    return PNP_OK;
}

}  // extern "C"
namespace {
template <int L>
int update_prox_fits(int W) {
    constexpr int GP = pnp::upd_gp<L>();
    const int npairs = W / 2;
    int grid = num_sms() < npairs ? num_sms() : npairs;
    const int ppc = (npairs + grid - 1) / grid;
    const size_t smem = sizeof(float) * ((size_t)pnp::lines_stage_off<L, GP>() + (size_t)2 * ppc * L);
    return (smem <= 220 * 1024 && 2 * ppc >= GP) ? 1 : 0;
}
int dispatch_update_prox_fits(int n, int W) { DISPATCH_POW2(n, update_prox_fits, W) }
}  // namespace
extern "C" {

int pnp_csmri_update_prox_supported(int H, int W) {
    if (!pow2_ok(H) || W < 2 || (W & 1) || H < 512) return 0;
    const int r = dispatch_update_prox_fits(H, W);
    return r > 0 ? 1 : 0;
}

int pnp_csmri_update_prox(const float* S, int H, int W, float gscale, float step, const float* step_ptr, const float* vadd,
                          const float* z_in, float* z_out, double* sig_log, float sigma_modifier, float fallback_sigma,
                          const float* xrec, double* mse_log, const int* slot, int* advance_counters, int n_advance,
                          unsigned* barrier_ws, int chain, void* stream) {
    return pnp_csmri_update_prox_next(S, H, W, gscale, step, step_ptr, vadd, z_in, z_out, sig_log, sigma_modifier, fallback_sigma, xrec,
                                      mse_log, slot, advance_counters, n_advance, barrier_ws, chain, nullptr, stream);
}

int pnp_csmri_update_prox_next(const float* S, int H, int W, float gscale, float step, const float* step_ptr, const float* vadd,
                               const float* z_in, float* z_out, double* sig_log, float sigma_modifier, float fallback_sigma,
                               const float* xrec, double* mse_log, const int* slot, int* advance_counters, int n_advance,
                               unsigned* barrier_ws, int chain, const pnp_csmri_next_pass* next, void* stream) {
    if (next) {
        if (!next->w || !next->S_out) return fail(PNP_ERR_ARG, "next pass: w and S_out are required");
        if (next->sel_count > 0) {
            if (!next->bits) return fail(PNP_ERR_ARG, "next pass: selection needs the bits buffer");
            if (!next->sel_idx && (!next->sel_support || !next->sel_m0)) return fail(PNP_ERR_ARG, "next pass: neither positions nor a support list");
            if (!next->sel_idx && next->sel_count > next->sel_min_m0)
                return fail(PNP_ERR_ARG, "next pass: sel_count %d exceeds the %d sampled positions", next->sel_count, next->sel_min_m0);
        }
    }
    if (!vadd || !z_in || !z_out || !sig_log) return fail(PNP_ERR_ARG, "bad argument");      // S may be null: zero spectrum
    if (advance_counters && (n_advance < 1 || n_advance > 32)) return fail(PNP_ERR_ARG, "n_advance must be in [1, 32]");
    if (!pow2_ok(H) || W < 2 || (W & 1)) return fail(PNP_ERR_ARG, "H must be a power of two in [32, 4096], W even");
    // (below 512 samples a round would hold more than 16 transforms, whose exchange planes are only 4-byte aligned)
    if (H < 512) return fail(PNP_ERR_UNSUPPORTED, "update+prox: lines shorter than 512 samples use the separate kernels");
    const float inv_n = (float)(1.0 / ((double)H * (double)W));
    return dispatch_update_prox(H, reinterpret_cast<const float*>(S), W, inv_n, gscale, step, step_ptr, vadd, z_in, z_out, xrec,
                                sigma_modifier, fallback_sigma, sig_log, mse_log, slot, advance_counters, n_advance,
                                barrier_ws, chain, next, static_cast<cudaStream_t>(stream));
}

int pnp_cdp_grad(const pnp_cdp_grad_args* args, void* stream) {
    if (!args) return fail(PNP_ERR_ARG, "null args");
    const pnp_cdp_grad_args& a = *args;
    if (!pow2_ok(a.H) || !pow2_ok(a.W)) return fail(PNP_ERR_ARG, "H and W must be powers of two in [32, 4096]");
    if (a.L < 1 || !a.codes || !a.y || !a.z || !a.S || !a.acc) return fail(PNP_ERR_ARG, "bad argument");
    if (a.sel_idx && (!a.mask || a.count < 1)) return fail(PNP_ERR_ARG, "sel_idx needs mask scratch and count >= 1");
    if (a.z_out && !a.z_in) return fail(PNP_ERR_ARG, "z_out needs z_in");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const long long N = (long long)a.H * a.W;
    if (N * a.L >= (1ll << 31)) return fail(PNP_ERR_ARG, "L*H*W too large");
    float2* S = reinterpret_cast<float2*>(a.S);
    unsigned char* mask = a.sel_idx ? a.mask : nullptr;
    if (a.sel_idx) {
        int blocks = (a.count + 255) / 256;
        if (blocks > 1184) blocks = 1184;
        pnp::k_cdp_sel<<<blocks, 256, 0, st>>>(a.mask, a.sel_idx, a.count, a.H, a.W, a.cursor);
        LAUNCH_CHECK();
    }
    const float inv_n = (float)(1.0 / (double)N);
    const int npts = a.w ? 2 : 1;
    int rc;
    if (npts == 2 && a.S2) {
        // both points in the same three launches (second scratch given): the transforms of z and w are independent, the
        // column pass keeps the single-use selection until its second point, the inverse pass forms the difference
        float2* S2 = reinterpret_cast<float2*>(a.S2);
        if ((rc = dispatch_cdp_lines_fwd(a.H, a.z, a.codes, S, a.W, a.L, a.w, S2, st)) != PNP_OK) return rc;
        if ((rc = dispatch_cdp_cols(a.W, S, a.y, mask, a.H, a.L, inv_n, 1, S2, st)) != PNP_OK) return rc;
        if ((rc = dispatch_cdp_lines_inv(a.H, S, a.codes, a.acc, a.W, a.L, 1.f, 0, S2, st)) != PNP_OK) return rc;
    } else
    for (int pt = 0; pt < npts; ++pt) {
        const float* u = pt == 0 ? a.z : a.w;
        if ((rc = dispatch_cdp_lines_fwd(a.H, u, a.codes, S, a.W, a.L, nullptr, nullptr, st)) != PNP_OK) return rc;
        if ((rc = dispatch_cdp_cols(a.W, S, a.y, mask, a.H, a.L, inv_n, pt == npts - 1, nullptr, st)) != PNP_OK) return rc;
        if ((rc = dispatch_cdp_lines_inv(a.H, S, a.codes, a.acc, a.W, a.L, pt == 0 ? 1.f : -1.f, pt, nullptr, st)) != PNP_OK) return rc;
    }
    pnp::k_cdp_epilogue<<<ew_blocks(N, 1), 256, 0, st>>>(a.acc, N, a.gscale * inv_n, a.step, a.step_ptr, a.g_out, a.vadd, a.v_out,
                                                         a.z_in, a.z_out);
    LAUNCH_CHECK();
    return PNP_OK;
}

int pnp_tv_chambolle(const float* z_in, float* z_out, int H, int W, int batch, float weight, const double* sig_log,
                     float sigma_modifier, float fallback_weight, int n_iter, float* work, const float* xrec,
                     double* mse_log, const int* slot, void* stream) {
    if (!z_in || !z_out || !work || H < 1 || W < 1 || batch < 1 || n_iter < 1) return fail(PNP_ERR_ARG, "bad argument");
    if (z_in == z_out) return fail(PNP_ERR_ARG, "z_out must not alias z_in");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    { const int rc = raise_smem_limit((const void*)pnp::k_tv_chambolle, (int)TV_SMEM_BYTES); if (rc != PNP_OK) return rc; }
    // device layout: W lines of H samples; the operator is symmetric under transposition
    const int nl = W, np = H;
    const long long n = (long long)nl * np;
    dim3 grid((np + TV_TP - 1) / TV_TP, (nl + TV_TL - 1) / TV_TL, batch);
    float* pa = work;
    float* pb = work + 2 * n * batch;
    int left = n_iter - 1, first = 1;
    do {
        const int u = left < TV_KB ? left : TV_KB;
        left -= u;
        pnp::k_tv_chambolle<<<grid, 256, TV_SMEM_BYTES, st>>>(z_in, pa, pb, z_out, nl, np, u, first, left == 0 ? 1 : 0,
                                                              weight, sig_log, sigma_modifier, fallback_weight, xrec, mse_log,
                                                              slot, batch);
        LAUNCH_CHECK();
        float* t = pa; pa = pb; pb = t;
        first = 0;
    } while (left > 0);
    return PNP_OK;
}

int pnp_nlm_denoise(const float* z_in, float* z_out, int H, int W, int batch, int patch_size, int patch_distance,
                    const double* sig_log, float sigma_est, float sigma_modifier, float fallback_h,
                    const float* xrec, double* mse_log, const int* slot, void* stream) {
    if (!z_in || !z_out || z_in == z_out || batch < 1 || H < 1 || W < 1) return fail(PNP_ERR_ARG, "bad argument (z_out must not alias z_in)");
    const int s = (patch_size % 2 == 0) ? patch_size + 1 : patch_size;
    if (s < 1 || s > NLM_MAX_S || patch_distance < 0 || patch_distance + s / 2 > NLM_MAX_HALO)
        return fail(PNP_ERR_ARG, "patch_size %d / patch_distance %d not supported", patch_size, patch_distance);
    pnp::NlmParams np_{s, patch_distance, sig_log, sigma_est, sigma_modifier, fallback_h};
    static const bool generic_only = [] { const char* e = std::getenv("PNP_NLM_GENERIC"); return e && std::atoi(e) != 0; }();
    if (s == 5 && !generic_only) {        // the reference's patch size (denoisers/NLM.py:11: 4 -> 5): specialised kernel
        const int halo = patch_distance + 2;
        const int TWd = NLM5_TX + 2 * halo, TH = NLM5_TY + 2 * halo;
        const int TWP = ((TWd + 31) / 32) * 32 + 16;
        dim3 grid((W + NLM5_TX - 1) / NLM5_TX, (H + NLM5_TY - 1) / NLM5_TY, batch);
        pnp::k_nlm5<<<grid, dim3(NLM5_TX, NLM5_TY, NLM5_G), sizeof(float) * (size_t)TH * TWP, static_cast<cudaStream_t>(stream)>>>(
            z_in, z_out, xrec, H, W, (long long)H * W, np_, mse_log, slot, batch);
        LAUNCH_CHECK();
        return PNP_OK;
    }
    const int TW = NLM_TILE + 2 * (patch_distance + s / 2);
    dim3 grid((W + NLM_TILE - 1) / NLM_TILE, (H + NLM_TILE - 1) / NLM_TILE, batch);
    pnp::k_nlm<<<grid, dim3(NLM_TILE, NLM_TILE), sizeof(float) * TW * TW, static_cast<cudaStream_t>(stream)>>>(
        z_in, z_out, xrec, H, W, (long long)H * W, np_, mse_log, slot, batch);
    LAUNCH_CHECK();
    return PNP_OK;
}

namespace {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int make_tmap_bf16_2d(CUtensorMap* map, const void* base, unsigned long long inner, unsigned long long outer,
                      unsigned box_inner, unsigned box_outer, CUtensorMapSwizzle swizzle = CU_TENSOR_MAP_SWIZZLE_128B) {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        CU_TRY(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
        if (!p || q != cudaDriverEntryPointSuccess) return fail(PNP_ERR_CUDA, "cuTensorMapEncodeTiled not available");
        fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    const cuuint64_t dims[2] = {inner, outer};
    const cuuint64_t strides[1] = {inner * 2};
    const cuuint32_t box[2] = {box_inner, box_outer};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(PNP_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
    return PNP_OK;
}

int g_tc_dbg = 0;

int cnn_forward_tc(const pnp_cnn_net* net, const float* img, float* out, int PH, int PW, void* act0, void* act1, int* stats,
                   const float* xrec, double* mse_log, const int* slot, cudaStream_t st, bool split) {
    const long long npix = (long long)PH * PW;
    const long long S = (long long)PH * (PW + 1);
    if (S >= (1ll << 31) - 4096) return fail(PNP_ERR_ARG, "image too large");
    const int L = net->n_layers;
    for (int l = 1; l < L; ++l) {
        if (!net->w_tc[l]) return fail(PNP_ERR_ARG, "w_tc[%d] missing: the net was not packed for the tensor-core path", l);
        if (split && !net->w_tc_lo[l]) return fail(PNP_ERR_ARG, "w_tc_lo[%d] missing: the net was not packed for the split mode", l);
        if (l < L - 1 && !(net->slope[l] <= 1.f)) return fail(PNP_ERR_UNSUPPORTED, "activation slope %g > 1 on the tensor-core path", net->slope[l]);
    }
    int rc;
    if (split) {
        if ((rc = raise_smem_limit((const void*)pnp::k_conv_tc<64, true>, (int)pnp::tc_smem<64, true>())) != PNP_OK) return rc;
        if ((rc = raise_smem_limit((const void*)pnp::k_conv_tc<1, true>, (int)pnp::tc_smem<1, true>())) != PNP_OK) return rc;
    }
    if ((rc = raise_smem_limit((const void*)pnp::k_conv_tc<64>, (int)pnp::tc_smem<64>())) != PNP_OK) return rc;
    if ((rc = raise_smem_limit((const void*)pnp::k_conv_tc<1>, (int)pnp::tc_smem<1>())) != PNP_OK) return rc;
    pnp::CnnIo io{net->mode, stats, net->range, net->shift_in};
    if (net->mode == 0) {
        pnp::k_minmax_init<<<1, 32, 0, st>>>(stats);
        LAUNCH_CHECK();
        pnp::k_minmax<<<ew_blocks(npix, 4), 256, 0, st>>>(img, npix, stats);
        LAUNCH_CHECK();
    }
    __nv_bfloat16* cur = static_cast<__nv_bfloat16*>(act0);
    __nv_bfloat16* nxt = static_cast<__nv_bfloat16*>(act1);
    const long long plane = S * 64;                  // split mode: the lo plane of an activation buffer follows its hi plane
    pnp::k_conv_first_bf16<<<dim3((PW + 31) / 32, (PH + FL_R - 1) / FL_R), 256, 0, st>>>(img, cur, net->w[0],
        pnp::CnnAct{net->scale[0], net->shift[0], net->slope[0]}, io, PH, PW, split ? cur + plane : nullptr);
    LAUNCH_CHECK();
    const int n_tiles = (int)((S + TC_OUT_PER_TILE - 1) / TC_OUT_PER_TILE);
    const int grid = n_tiles < num_sms() ? n_tiles : num_sms();
    CUtensorMap tmA, tmB, tmO;
    if (split) {
        // error-compensated mode: one launch per layer, hi and lo planes of activations and weights (csrc/cnn_tc.cuh)
        CUtensorMap tmAl, tmBl, tmOl;
        for (int l = 1; l < L - 1; ++l) {
            if ((rc = make_tmap_bf16_2d(&tmA, cur, 64, (unsigned long long)S, 64, TC_Q_ROWS)) != PNP_OK) return rc;
            if ((rc = make_tmap_bf16_2d(&tmAl, cur + plane, 64, (unsigned long long)S, 64, TC_Q_ROWS)) != PNP_OK) return rc;
            if ((rc = make_tmap_bf16_2d(&tmB, net->w_tc[l], 192, 192, 64, 192)) != PNP_OK) return rc;
            if ((rc = make_tmap_bf16_2d(&tmBl, net->w_tc_lo[l], 192, 192, 64, 192)) != PNP_OK) return rc;
            if ((rc = make_tmap_bf16_2d(&tmO, nxt, 64, (unsigned long long)S, 16, TC_OUT_PER_Q, CU_TENSOR_MAP_SWIZZLE_32B)) != PNP_OK) return rc;
            if ((rc = make_tmap_bf16_2d(&tmOl, nxt + plane, 64, (unsigned long long)S, 16, TC_OUT_PER_Q, CU_TENSOR_MAP_SWIZZLE_32B)) != PNP_OK) return rc;
            pnp::k_conv_tc<64, true><<<grid, TC_THREADS, pnp::tc_smem<64, true>(), st>>>(
                tmA, tmB, tmO, net->shift[l], net->slope[l], PW, (int)S, n_tiles, pnp::TcLast{}, 0, tmAl, tmBl, tmOl);
            LAUNCH_CHECK();
            __nv_bfloat16* t = cur; cur = nxt; nxt = t;
        }
        if ((rc = make_tmap_bf16_2d(&tmA, cur, 64, (unsigned long long)S, 64, TC_Q_ROWS)) != PNP_OK) return rc;
        if ((rc = make_tmap_bf16_2d(&tmAl, cur + plane, 64, (unsigned long long)S, 64, TC_Q_ROWS)) != PNP_OK) return rc;
        if ((rc = make_tmap_bf16_2d(&tmB, net->w_tc[L - 1], 192, 16, 64, 16)) != PNP_OK) return rc;
        if ((rc = make_tmap_bf16_2d(&tmBl, net->w_tc_lo[L - 1], 192, 16, 64, 16)) != PNP_OK) return rc;
        pnp::k_conv_tc<1, true><<<grid, TC_THREADS, pnp::tc_smem<1, true>(), st>>>(
            tmA, tmB, tmA, nullptr, 0.f, PW, (int)S, n_tiles, pnp::TcLast{img, out, xrec, mse_log, slot, net->last_bias, io}, 0,
            tmAl, tmBl, tmA);
        LAUNCH_CHECK();
        return PNP_OK;
    }
    // small images: all middle layers in one cooperative launch (a layer would be a few tiles per SM inside a ~10 us launch)
    const bool stack = g_tc_dbg != 32 && L - 2 >= 2 && L - 2 <= TC_MAX_LAYERS && n_tiles <= 8 * num_sms();
    if (stack) {
        if ((rc = raise_smem_limit((const void*)pnp::k_conv_tc_stack, (int)pnp::tc_smem<64>())) != PNP_OK) return rc;
        pnp::TcStack pm{};
        pm.n_layers = L - 2;
        if ((rc = make_tmap_bf16_2d(&pm.a[0], cur, 64, (unsigned long long)S, 64, TC_Q_ROWS)) != PNP_OK) return rc;
        if ((rc = make_tmap_bf16_2d(&pm.a[1], nxt, 64, (unsigned long long)S, 64, TC_Q_ROWS)) != PNP_OK) return rc;
        if ((rc = make_tmap_bf16_2d(&pm.o[0], nxt, 64, (unsigned long long)S, 16, TC_OUT_PER_Q, CU_TENSOR_MAP_SWIZZLE_32B)) != PNP_OK) return rc;
        if ((rc = make_tmap_bf16_2d(&pm.o[1], cur, 64, (unsigned long long)S, 16, TC_OUT_PER_Q, CU_TENSOR_MAP_SWIZZLE_32B)) != PNP_OK) return rc;
        for (int l = 1; l < L - 1; ++l) {
            if ((rc = make_tmap_bf16_2d(&pm.b[l - 1], net->w_tc[l], 192, 192, 64, 192)) != PNP_OK) return rc;
            pm.shift[l - 1] = net->shift[l];
            pm.slope[l - 1] = net->slope[l];
        }
        int pw = PW, s32 = (int)S, nt = n_tiles;
        void* args[] = {(void*)&pm, (void*)&pw, (void*)&s32, (void*)&nt};
        CU_TRY(cudaLaunchCooperativeKernel((const void*)pnp::k_conv_tc_stack, dim3(grid), dim3(TC_THREADS), args, pnp::tc_smem<64>(), st));
        if ((L - 2) & 1) { __nv_bfloat16* t = cur; cur = nxt; nxt = t; }      // where the last middle layer wrote
    }
    for (int l = 1; l < L - 1 && !stack; ++l) {
        if ((rc = make_tmap_bf16_2d(&tmA, cur, 64, (unsigned long long)S, 64, TC_Q_ROWS)) != PNP_OK) return rc;
        if ((rc = make_tmap_bf16_2d(&tmB, net->w_tc[l], 192, 192, 64, 192)) != PNP_OK) return rc;
        if ((rc = make_tmap_bf16_2d(&tmO, nxt, 64, (unsigned long long)S, 16, TC_OUT_PER_Q, CU_TENSOR_MAP_SWIZZLE_32B)) != PNP_OK) return rc;
        pnp::k_conv_tc<64><<<grid, TC_THREADS, pnp::tc_smem<64>(), st>>>(tmA, tmB, tmO, net->shift[l], net->slope[l], PW, (int)S, n_tiles,
                                                                         pnp::TcLast{}, g_tc_dbg, tmA, tmB, tmO);
        LAUNCH_CHECK();
        __nv_bfloat16* t = cur; cur = nxt; nxt = t;
    }
    if ((rc = make_tmap_bf16_2d(&tmA, cur, 64, (unsigned long long)S, 64, TC_Q_ROWS)) != PNP_OK) return rc;
    if ((rc = make_tmap_bf16_2d(&tmB, net->w_tc[L - 1], 192, 16, 64, 16)) != PNP_OK) return rc;
    pnp::k_conv_tc<1><<<grid, TC_THREADS, pnp::tc_smem<1>(), st>>>(tmA, tmB, tmA, nullptr, 0.f, PW, (int)S, n_tiles,
                                                              pnp::TcLast{img, out, xrec, mse_log, slot, net->last_bias, io}, 0,
                                                              tmA, tmB, tmA);
    LAUNCH_CHECK();
    return PNP_OK;
}

}  // namespace

int pnp_cnn_forward(const pnp_cnn_net* net, const float* img, float* out, int PH, int PW, float* act0, float* act1,
                    int* stats, const float* xrec, double* mse_log, const int* slot, int precision, void* stream) {
    if (!net || !img || !out || !act0 || !act1 || !stats || PH < 1 || PW < 1) return fail(PNP_ERR_ARG, "bad argument");
    if (net->n_layers < 2 || net->n_layers > PNP_CNN_MAX_LAYERS) return fail(PNP_ERR_ARG, "n_layers out of range");
    if (precision < 0 || precision > 2) return fail(PNP_ERR_ARG, "precision %d unknown", precision);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (precision >= 1) return cnn_forward_tc(net, img, out, PH, PW, act0, act1, stats, xrec, mse_log, slot, st, precision == 2);
    const long long npix = (long long)PH * PW;
    pnp::CnnIo io{net->mode, stats, net->range, net->shift_in};
    if (net->mode == 0) {
        pnp::k_minmax_init<<<1, 32, 0, st>>>(stats);
        LAUNCH_CHECK();
        pnp::k_minmax<<<ew_blocks(npix, 4), 256, 0, st>>>(img, npix, stats);
        LAUNCH_CHECK();
    }
    const int L = net->n_layers;
    pnp::k_conv_first<<<ew_blocks(npix * 16, 1), 256, 0, st>>>(img, act0, net->w[0],
        pnp::CnnAct{net->scale[0], net->shift[0], net->slope[0]}, io, PH, PW);
    LAUNCH_CHECK();
    float* cur = act0;
    float* nxt = act1;
    dim3 grid((PW + CM_TP - 1) / CM_TP, (PH + CM_TL - 1) / CM_TL);
    for (int l = 1; l < L - 1; ++l) {
        pnp::k_conv_mid<<<grid, 256, 0, st>>>(cur, nxt, net->w[l], pnp::CnnAct{net->scale[l], net->shift[l], net->slope[l]}, PH, PW);
        LAUNCH_CHECK();
        float* t = cur; cur = nxt; nxt = t;
    }
    long long lb = (npix + 7) / 8;
    if (lb > 148 * 16) lb = 148 * 16;
    pnp::k_conv_last<<<(unsigned)lb, 256, 0, st>>>(cur, img, out, net->w[L - 1], net->last_bias, io, PH, PW, xrec, mse_log, slot);
    LAUNCH_CHECK();
    return PNP_OK;
}

int pnp_debug_set(int key, int value) {
    if (key == 1) g_tc_dbg = value;
#ifdef PNP_TRACE
    if (key == 3) CU_TRY(cudaMemcpyToSymbol(pnp::g_dbg_flags, &value, sizeof(int)));
#endif
    if (key == 2) {                          // PNP_PHASE_TIMING builds: print the phase boundaries of the last k_update_prox
        unsigned long long t[8];
        CU_TRY(cudaMemcpyFromSymbol(t, pnp::g_upd_phase_ns, sizeof(t)));
        std::fprintf(stderr, "k_update_prox CTA 0 phases (us): c2r+update %.2f  sigma %.2f  grid barrier %.2f  shrink %.2f\n",
                     (t[1] - t[0]) * 1e-3, (t[2] - t[1]) * 1e-3, (t[3] - t[2]) * 1e-3, (t[4] - t[3]) * 1e-3);
    }
    return PNP_OK;
}

int pnp_debug_read(int key, void* out_host, long long bytes) {
    if (key != 1 || !out_host || bytes < 16) return fail(PNP_ERR_ARG, "bad argument");
#ifdef PNP_TRACE
    unsigned n = 0;
    CU_TRY(cudaDeviceSynchronize());
    CU_TRY(cudaMemcpyFromSymbol(&n, pnp::g_trace_n, sizeof(n)));
    if (n > PNP_TRACE_MAX) n = PNP_TRACE_MAX;
    const long long cap = bytes / (long long)sizeof(pnp::TraceEv);
    if ((long long)n > cap) n = (unsigned)cap;
    if (n) CU_TRY(cudaMemcpyFromSymbol(out_host, pnp::g_trace, sizeof(pnp::TraceEv) * (size_t)n));
    const unsigned zero = 0;
    CU_TRY(cudaMemcpyToSymbol(pnp::g_trace_n, &zero, sizeof(zero)));
    return (int)n;
#else
    return fail(PNP_ERR_UNSUPPORTED, "library built without -DPNP_TRACE");
#endif
}

int pnp_graph_begin(void* stream) {
    CU_TRY(cudaStreamBeginCapture(static_cast<cudaStream_t>(stream), cudaStreamCaptureModeRelaxed));
    return PNP_OK;
}

int pnp_graph_end(void* stream, void** exec_out) {
    if (!exec_out) return fail(PNP_ERR_ARG, "null exec_out");
    cudaGraph_t graph = nullptr;
    CU_TRY(cudaStreamEndCapture(static_cast<cudaStream_t>(stream), &graph));
    cudaGraphExec_t exec = nullptr;
    cudaError_t e = cudaGraphInstantiate(&exec, graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) return fail(PNP_ERR_CUDA, "cudaGraphInstantiate failed: %s", cudaGetErrorString(e));
    *exec_out = exec;
    return PNP_OK;
}

int pnp_graph_launch(void* exec, void* stream) {
    if (!exec) return fail(PNP_ERR_ARG, "null exec");
    CU_TRY(cudaGraphLaunch(static_cast<cudaGraphExec_t>(exec), static_cast<cudaStream_t>(stream)));
    return PNP_OK;
}

int pnp_graph_destroy(void* exec) {
    if (exec) CU_TRY(cudaGraphExecDestroy(static_cast<cudaGraphExec_t>(exec)));
    return PNP_OK;
}

}  // extern "C"

namespace {
// CTAs per cluster = per image: 8 (portable size; 15 clusters are co-resident on a B200) for batches, 16 for a few
// images: half the lines per CTA -> one line per warp in the prox phases, 10.9 instead of 15.0 us per inner iteration
// at 256^2, but only ~7 such clusters fit the GPCs
template <int L, int SMALL_C>
int launch_svrg_small(const pnp_csmri_svrg_small_args& a, cudaStream_t st, int* max_clusters_out = nullptr) {
    using K = pnp::SmallCfg<L, SMALL_C, 512>;
    const void* kernel = (const void*)pnp::k_csmri_svrg_small<L, SMALL_C, 512>;
    { const int rc = raise_smem_limit(kernel, (int)K::SMEM); if (rc != PNP_OK) return rc; }
    if (SMALL_C > 8) {
        static std::mutex mu;
        static bool done[64] = {false};
        std::lock_guard<std::mutex> lock(mu);
        const int dev = current_device();
        if (!done[dev]) { CU_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1)); done[dev] = true; }
    }
    pnp::SmallArgs k{};
    k.z = a.z; k.xrec = a.xrec;
    k.Y1 = reinterpret_cast<const float2*>(a.Y1); k.Y2 = reinterpret_cast<const float2*>(a.Y2);
    k.Y1n = reinterpret_cast<const float2*>(a.Y1n); k.Y2n = reinterpret_cast<const float2*>(a.Y2n);
    k.bits_full = a.bits_full; k.support = a.support; k.m0 = a.m0; k.support_img_stride = a.support_img_stride;
    k.idx = a.idx; k.idx_img_stride = a.idx_img_stride; k.idx_iter_stride = a.idx_iter_stride;
    k.snap_scale_ptr = a.snap_scale_ptr; k.snap_scale = a.snap_scale; k.step = a.step; k.step_img_stride = a.step_img_stride;
    k.sig_log = a.sig_log; k.mse_log = a.mse_log; k.slot = a.slot; k.draw_counter = a.draw_counter;
    k.batch = a.batch; k.n_inner = a.n_inner; k.T2 = a.T2; k.B = a.mini_batch_size; k.seed = a.seed;
    k.lr_decay = a.lr_decay; k.sigma_modifier = a.sigma_modifier; k.fallback_sigma = a.fallback_sigma;
    k.fallback_decay = a.fallback_decay;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)(a.batch * SMALL_C));
    cfg.blockDim = dim3(K::NT);
    cfg.dynamicSmemBytes = K::SMEM;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = SMALL_C; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    if (max_clusters_out) {                  // query only: how many such clusters are co-resident
        int n = 0;
        if (cudaOccupancyMaxActiveClusters(&n, kernel, &cfg) != cudaSuccess) { n = 0; cudaGetLastError(); }
        *max_clusters_out = n;
        return PNP_OK;
    }
    void* args[] = {(void*)&k};
    CU_TRY(cudaLaunchKernelExC(&cfg, kernel, args));
    return PNP_OK;
}

// co-resident clusters of 16 CTAs of the 256^2 kernel on the current device (0: such clusters cannot be scheduled)
int wide_clusters_256() {
    static std::mutex mu;
    static int cached[64];
    static bool have[64] = {false};
    const int dev = current_device();
    std::lock_guard<std::mutex> lock(mu);
    if (!have[dev]) {
        pnp_csmri_svrg_small_args probe{};
        probe.batch = 1;
        int n = 0;
        if (launch_svrg_small<256, 16>(probe, nullptr, &n) != PNP_OK) n = 0;
        cached[dev] = n;
        have[dev] = true;
    }
    return cached[dev];
}
}  // namespace

extern "C" {
int pnp_csmri_svrg_small_supported(int H, int W) { return (H == W && (H == 128 || H == 256)) ? 1 : 0; }

int pnp_csmri_svrg_small_capacity(int H, int W) {
    if (!pnp_csmri_svrg_small_supported(H, W)) return 0;
    static std::mutex mu;
    static int cached[64][2];
    static bool have[64][2] = {{false}};
    const int dev = current_device(), which = H == 256 ? 1 : 0;
    std::lock_guard<std::mutex> lock(mu);
    if (!have[dev][which]) {
        pnp_csmri_svrg_small_args probe{};
        probe.batch = 64;
        int n = 0;
        const int rc = H == 256 ? launch_svrg_small<256, 8>(probe, nullptr, &n) : launch_svrg_small<128, 8>(probe, nullptr, &n);
        cached[dev][which] = rc == PNP_OK ? n : 0;
        have[dev][which] = true;
    }
    return cached[dev][which];
}

int pnp_csmri_svrg_small(const pnp_csmri_svrg_small_args* args, void* stream) {
    if (!args) return fail(PNP_ERR_ARG, "null args");
    const pnp_csmri_svrg_small_args& a = *args;
    { const int rc = check_init(); if (rc != PNP_OK) return rc; }
    if (!pnp_csmri_svrg_small_supported(a.H, a.W))
        return fail(PNP_ERR_UNSUPPORTED, "pnp_csmri_svrg_small: %d x %d is not a square image of 128 or 256 pixels a side", a.H, a.W);
    if (!a.z || !a.Y1 || !a.Y2 || !a.Y1n || !a.Y2n || !a.bits_full || !a.step || !a.sig_log)
        return fail(PNP_ERR_ARG, "pnp_csmri_svrg_small: null pointer");
    if (!a.idx && (!a.support || !a.m0)) return fail(PNP_ERR_ARG, "pnp_csmri_svrg_small: neither explicit minibatches nor a support list");
    if (a.batch < 1 || a.n_inner < 0 || a.T2 < 1 || a.mini_batch_size < 1)
        return fail(PNP_ERR_ARG, "pnp_csmri_svrg_small: batch, T2 and mini_batch_size must be >= 1, n_inner >= 0");
    if (a.n_inner == 0) return PNP_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (a.H == 128) return launch_svrg_small<128, 8>(a, st);
    // PNP_SMALL_CLUSTER=8|16 forces a cluster size (measurements); default: 16 while all images still run concurrently
    static const int forced = [] { const char* e = std::getenv("PNP_SMALL_CLUSTER"); return e ? std::atoi(e) : 0; }();
    const bool wide = forced == 16 || (forced != 8 && a.batch <= wide_clusters_256());
    return wide ? launch_svrg_small<256, 16>(a, st) : launch_svrg_small<256, 8>(a, st);
}

int pnp_advance_by(int* counters, int n, int delta, void* stream) {
    if (!counters || n < 1 || n > 32) return fail(PNP_ERR_ARG, "bad argument");
    pnp::k_advance_by<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(counters, n, delta);
    LAUNCH_CHECK();
    return PNP_OK;
}
}  // extern "C"
