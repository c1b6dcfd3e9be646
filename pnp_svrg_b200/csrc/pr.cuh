// Phase retrieval (dense real Gaussian A, amplitude loss) gradient for sm_100a.
//
// Reference: problems/PR.py:66-68 (forward |A w|), :75-79 (grad_full), :81-87 (grad_stoch):
//   t = A_sel w ;  r = ((|t| - y) / |t|) * t ;  g = A_sel^T r
// HBM-streaming GEMV pair.  The vector w lives in the transposed line layout, so A is stored with
// its columns permuted to that layout once at construction (A_dev[m][c*H + r] = A[m][r*W + c]).
//
//   k_pr_rows   one CTA per selected row: dot products with z (and optionally w, the SVRG
//               two-point form g_B(z) - g_B(w), which is NOT linear here) in a single pass over the
//               row; writes the residual weights r.
//   k_pr_cols   g[n] = sum_rows A[row][n] * r[row], thread per 4 columns, rows streamed (the row
//               block just read by k_pr_rows is L2 resident when it fits).
#pragma once
#include <cuda_runtime.h>

namespace pnp {

__device__ __forceinline__ float pr_weight(float t, float y) {
    const float a = fabsf(t);
    return ((a - y) / a) * t;            // division by |t| unguarded, as in the reference
}

// rows: selected measurement ids (null = all); out r[i] for i in [0, count)
__global__ void __launch_bounds__(256)
k_pr_rows(const float* __restrict__ A, const float* __restrict__ z, const float* __restrict__ w,
          const float* __restrict__ y, const int* __restrict__ rows, int count, long long n,
          float* __restrict__ r, const int* __restrict__ cursor) {
    const int* sel = rows ? rows + (long long)(cursor ? *cursor : 0) * count : nullptr;
    __shared__ float s1[8], s2[8];
    for (int i = blockIdx.x; i < count; i += gridDim.x) {
        const int row = sel ? sel[i] : i;
        const float4* a4 = reinterpret_cast<const float4*>(A + (long long)row * n);
        const float4* z4 = reinterpret_cast<const float4*>(z);
        const float4* w4 = reinterpret_cast<const float4*>(w);
        float dz = 0.f, dw = 0.f;
        for (long long j = threadIdx.x; j < n / 4; j += blockDim.x) {
            const float4 a = a4[j], q = z4[j];
            dz = fmaf(a.x, q.x, fmaf(a.y, q.y, fmaf(a.z, q.z, fmaf(a.w, q.w, dz))));
            if (w) {
                const float4 p = w4[j];
                dw = fmaf(a.x, p.x, fmaf(a.y, p.y, fmaf(a.z, p.z, fmaf(a.w, p.w, dw))));
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            dz += __shfl_xor_sync(0xffffffffu, dz, o);
            dw += __shfl_xor_sync(0xffffffffu, dw, o);
        }
        if ((threadIdx.x & 31) == 0) { s1[threadIdx.x >> 5] = dz; s2[threadIdx.x >> 5] = dw; }
        __syncthreads();
        if (threadIdx.x == 0) {
            float tz = 0.f, tw = 0.f;
            for (int k = 0; k < (int)(blockDim.x >> 5); ++k) { tz += s1[k]; tw += s2[k]; }
            const float yy = y[row];
            float v = pr_weight(tz, yy);
            if (w) v -= pr_weight(tw, yy);
            r[i] = v;
        }
        __syncthreads();
    }
}

// g[n] (+ epilogue)  = sum_i A[rows[i]][n] * r[i]
__global__ void __launch_bounds__(256)
k_pr_cols(const float* __restrict__ A, const float* __restrict__ r, const int* __restrict__ rows, int count,
          long long n, const int* __restrict__ cursor, float gscale, float step, const float* __restrict__ step_ptr,
          float* __restrict__ g_out, const float* __restrict__ vadd, float* __restrict__ v_out,
          const float* __restrict__ z_in, float* __restrict__ z_out) {
    const int* sel = rows ? rows + (long long)(cursor ? *cursor : 0) * count : nullptr;
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;      // float4 column index
    if (j >= n / 4) return;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    int i = 0;
    for (; i + 4 <= count; i += 4) {
        float4 a[4];
        float rr[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int row = sel ? sel[i + u] : i + u;
            a[u] = reinterpret_cast<const float4*>(A + (long long)row * n)[j];
            rr[u] = r[i + u];
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            acc.x = fmaf(a[u].x, rr[u], acc.x); acc.y = fmaf(a[u].y, rr[u], acc.y);
            acc.z = fmaf(a[u].z, rr[u], acc.z); acc.w = fmaf(a[u].w, rr[u], acc.w);
        }
    }
    for (; i < count; ++i) {
        const int row = sel ? sel[i] : i;
        const float4 a = reinterpret_cast<const float4*>(A + (long long)row * n)[j];
        const float rr = r[i];
        acc.x = fmaf(a.x, rr, acc.x); acc.y = fmaf(a.y, rr, acc.y); acc.z = fmaf(a.z, rr, acc.z); acc.w = fmaf(a.w, rr, acc.w);
    }
    const float s = step_ptr ? *step_ptr : step;
    float gv[4] = {acc.x * gscale, acc.y * gscale, acc.z * gscale, acc.w * gscale};
    if (g_out) reinterpret_cast<float4*>(g_out)[j] = make_float4(gv[0], gv[1], gv[2], gv[3]);
    float vv[4] = {gv[0], gv[1], gv[2], gv[3]};
    if (vadd) {
        const float4 q = reinterpret_cast<const float4*>(vadd)[j];
        vv[0] += q.x; vv[1] += q.y; vv[2] += q.z; vv[3] += q.w;
    }
    if (v_out) reinterpret_cast<float4*>(v_out)[j] = make_float4(vv[0], vv[1], vv[2], vv[3]);
    if (z_out) {
        const float4 q = reinterpret_cast<const float4*>(z_in)[j];
        reinterpret_cast<float4*>(z_out)[j] = make_float4(q.x - s * vv[0], q.y - s * vv[1], q.z - s * vv[2], q.w - s * vv[3]);
    }
}

// Split form of k_pr_cols for the case it was written wrong for: few columns (n = 4096 at 64 x 64 is 4 CTAs of the
// kernel above, each thread walking ALL rows: 492 us for a 33.6 MB matrix, 68 GB/s).  Here the rows are cut into
// gridDim.y chunks as well: CTA (bx, by) accumulates rows [by * per, (by + 1) * per) for 128 float4 columns and writes
// its partial sums to partial[by][n]; k_pr_cols_finish adds the chunks IN ORDER (deterministic, no atomics) and
// applies the epilogue.  Eight independent row loads in flight per thread.
__global__ void __launch_bounds__(128)
k_pr_cols_split(const float* __restrict__ A, const float* __restrict__ r, const int* __restrict__ rows, int count,
                long long n, const int* __restrict__ cursor, float* __restrict__ partial) {
    const int* sel = rows ? rows + (long long)(cursor ? *cursor : 0) * count : nullptr;
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;      // float4 column index
    const int per = (count + (int)gridDim.y - 1) / (int)gridDim.y;
    const int lo = (int)blockIdx.y * per, hi = min(count, lo + per);
    if (j >= n / 4) return;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    int i = lo;
    for (; i + 8 <= hi; i += 8) {
        float4 a[8];
        float rr[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int row = sel ? sel[i + u] : i + u;
            a[u] = __ldcs(reinterpret_cast<const float4*>(A + (long long)row * n) + j);
            rr[u] = r[i + u];
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            acc.x = fmaf(a[u].x, rr[u], acc.x); acc.y = fmaf(a[u].y, rr[u], acc.y);
            acc.z = fmaf(a[u].z, rr[u], acc.z); acc.w = fmaf(a[u].w, rr[u], acc.w);
        }
    }
    for (; i < hi; ++i) {
        const int row = sel ? sel[i] : i;
        const float4 a = __ldcs(reinterpret_cast<const float4*>(A + (long long)row * n) + j);
        const float rr = r[i];
        acc.x = fmaf(a.x, rr, acc.x); acc.y = fmaf(a.y, rr, acc.y); acc.z = fmaf(a.z, rr, acc.z); acc.w = fmaf(a.w, rr, acc.w);
    }
    reinterpret_cast<float4*>(partial + (long long)blockIdx.y * n)[j] = acc;
}

// 32 float4 columns per CTA, the chunks dealt over the 8 warps (each thread adds chunks w, w + 8, ... with the loads
// independent), then the 8 partial sums are added in warp order by warp 0: a fixed summation order.  (One thread per
// column walking all chunks was 16 us for 1 MB: 4 CTAs, 64 dependent-latency loads each.)
__global__ void __launch_bounds__(256)
k_pr_cols_finish(const float* __restrict__ partial, int chunks, long long n, float gscale, float step,
                 const float* __restrict__ step_ptr, float* __restrict__ g_out, const float* __restrict__ vadd,
                 float* __restrict__ v_out, const float* __restrict__ z_in, float* __restrict__ z_out) {
    __shared__ float4 red[8][32];
    const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
    const long long j = (long long)blockIdx.x * 32 + lane;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (j < n / 4) {
#pragma unroll 8
        for (int c = wp; c < chunks; c += 8) {
            const float4 p = reinterpret_cast<const float4*>(partial + (long long)c * n)[j];
            acc.x += p.x; acc.y += p.y; acc.z += p.z; acc.w += p.w;
        }
    }
    red[wp][lane] = acc;
    __syncthreads();
    if (wp != 0 || j >= n / 4) return;
#pragma unroll
    for (int k = 1; k < 8; ++k) {
        const float4 p = red[k][lane];
        acc.x += p.x; acc.y += p.y; acc.z += p.z; acc.w += p.w;
    }
    const float s = step_ptr ? *step_ptr : step;
    float gv[4] = {acc.x * gscale, acc.y * gscale, acc.z * gscale, acc.w * gscale};
    if (g_out) reinterpret_cast<float4*>(g_out)[j] = make_float4(gv[0], gv[1], gv[2], gv[3]);
    float vv[4] = {gv[0], gv[1], gv[2], gv[3]};
    if (vadd) {
        const float4 q = reinterpret_cast<const float4*>(vadd)[j];
        vv[0] += q.x; vv[1] += q.y; vv[2] += q.z; vv[3] += q.w;
    }
    if (v_out) reinterpret_cast<float4*>(v_out)[j] = make_float4(vv[0], vv[1], vv[2], vv[3]);
    if (z_out) {
        const float4 q = reinterpret_cast<const float4*>(z_in)[j];
        reinterpret_cast<float4*>(z_out)[j] = make_float4(q.x - s * vv[0], q.y - s * vv[1], q.z - s * vv[2], q.w - s * vv[3]);
    }
}

}  // namespace pnp
