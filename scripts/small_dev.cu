// Development harness for csrc/small.cuh: the cluster kernel alone in a small shared library (compiles in seconds;
// the full library takes two minutes), with the event trace.  Built and driven by scripts/small_dev.py.
#include <cmath>
#include <cstdio>
#include <vector>
#include "../include/pnp_b200.h"
#include "../pnp_svrg_b200/csrc/small.cuh"

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { std::fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e)); return -2; } } while (0)

template <int L, int C, int NT>
static int launch(const pnp_csmri_svrg_small_args& a, cudaStream_t st, int* max_clusters = nullptr) {
    using K = pnp::SmallCfg<L, C, NT>;
    const void* kernel = (const void*)pnp::k_csmri_svrg_small<L, C, NT>;
    CK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)K::SMEM));
    if (C > 8) CK(cudaFuncSetAttribute(kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
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
    cfg.gridDim = dim3((unsigned)(a.batch * C));
    cfg.blockDim = dim3(K::NT);
    cfg.dynamicSmemBytes = K::SMEM;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = C; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    if (max_clusters) {
        CK(cudaOccupancyMaxActiveClusters(max_clusters, kernel, &cfg));
        return 0;
    }
    void* args[] = {(void*)&k};
    CK(cudaLaunchKernelExC(&cfg, kernel, args));
    return 0;
}

extern "C" {
int dev_init(void) {
    std::vector<float2> tw(PNP_TW_N);
    for (int m = 0; m < PNP_TW_N; ++m) {
        const double ang = -2.0 * M_PI * (double)m / (double)PNP_TW_N;
        tw[m] = make_float2((float)std::cos(ang), (float)std::sin(ang));
    }
    CK(cudaMemcpyToSymbol(g_tw, tw.data(), sizeof(float2) * PNP_TW_N));
    return 0;
}
// variant: cluster size * 1000 + threads per CTA
static int dispatch(const pnp_csmri_svrg_small_args* a, cudaStream_t st, int c, int nt, int* mc) {
    if (a->H == 128) return launch<128, 8, 512>(*a, st, mc);
    if (c == 16 && nt == 256) return launch<256, 16, 256>(*a, st, mc);
    if (c == 16) return launch<256, 16, 512>(*a, st, mc);
    return launch<256, 8, 512>(*a, st, mc);
}
int dev_run(const pnp_csmri_svrg_small_args* a, void* stream, int c, int nt) { return dispatch(a, (cudaStream_t)stream, c, nt, nullptr); }
int dev_max_clusters(int L, int c, int nt) {
    pnp_csmri_svrg_small_args a{};
    a.H = a.W = L;
    a.batch = 64;
    int n = 0;
    const int rc = dispatch(&a, nullptr, c, nt, &n);
    return rc ? rc : n;
}
int dev_trace_read(void* out_host, long long bytes) {
#ifdef PNP_TRACE
    unsigned n = 0;
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpyFromSymbol(&n, pnp::g_trace_n, sizeof(n)));
    if (n > PNP_TRACE_MAX) n = PNP_TRACE_MAX;
    const long long cap = bytes / (long long)sizeof(pnp::TraceEv);
    if ((long long)n > cap) n = (unsigned)cap;
    if (n) CK(cudaMemcpyFromSymbol(out_host, pnp::g_trace, sizeof(pnp::TraceEv) * (size_t)n));
    const unsigned zero = 0;
    CK(cudaMemcpyToSymbol(pnp::g_trace_n, &zero, sizeof(zero)));
    return (int)n;
#else
    return -4;
#endif
}
}
