// Variance-reduction bookkeeping that is not already fused into the gradient epilogue.
//
//   k_axpy          z_out = z_in - step * v                as-committed SVRG (v = mu,
//                                                          algorithms/pnp_svrg.py:54-57) and the
//                                                          generic update for problems whose
//                                                          gradient kernel has no fused epilogue
//   k_saga_update   algorithms/pnp_saga.py:45-50,72:  table[slot] = g_new;
//                   v = g_new - g_prev + mean(table);  z -= step * v;  (g_prev <- g_new is a
//                   pointer swap on the host).  The table mean is kept as a running sum so the
//                   reference's O(hist * N) Python sum() becomes O(N).
//   k_advance       end-of-iteration counters (log slot, minibatch cursor) kept on the device so a
//                   captured CUDA graph can be replayed without new arguments.
#pragma once
#include <cuda_runtime.h>

namespace pnp {

__global__ void __launch_bounds__(256)
k_axpy(const float* __restrict__ z_in, const float* __restrict__ v, float* __restrict__ z_out, long long n,
       long long img_stride, float step, const float* __restrict__ step_ptr) {
    const int img = blockIdx.y;
    const float s = step_ptr ? step_ptr[img] : step;
    const long long o = (long long)img * img_stride;
    for (long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n;
         i += (long long)gridDim.x * blockDim.x * 4) {
        const float4 a = *reinterpret_cast<const float4*>(z_in + o + i);
        const float4 b = *reinterpret_cast<const float4*>(v + o + i);
        *reinterpret_cast<float4*>(z_out + o + i) =
            make_float4(a.x - s * b.x, a.y - s * b.y, a.z - s * b.z, a.w - s * b.w);
    }
}

// table: [img][hist][n]; tsum: [img][n] running sum of the table rows; slot_idx[img][cursor]
__global__ void __launch_bounds__(256)
k_saga_update(const float* __restrict__ g_new, float* __restrict__ g_prev, float* __restrict__ table,
              float* __restrict__ tsum, float* __restrict__ z, long long n, long long img_stride, int hist,
              const int* __restrict__ slot_idx, long long slot_img_stride, const int* __restrict__ cursor,
              float step, const float* __restrict__ step_ptr) {
    const int img = blockIdx.y;
    const float s = step_ptr ? step_ptr[img] : step;
    const int sl = slot_idx[(long long)img * slot_img_stride + (cursor ? *cursor : 0)];
    const long long o = (long long)img * img_stride;
    float* row = table + ((long long)img * hist + sl) * img_stride;
    const float inv_h = 1.0f / (float)hist;
    // six image-sized streams (g_new, table row, tsum, z, g_prev read; row, tsum, z, g_prev written): 16-byte accesses,
    // the arithmetic per element exactly as the scalar form (n and the image stride are multiples of 4 for every image
    // size the library accepts; a ragged tail falls back to scalars)
    const long long n4 = ((n & 3) == 0 && (img_stride & 3) == 0) ? n / 4 : 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const float4 gn = reinterpret_cast<const float4*>(g_new + o)[i];
        const float4 ts = reinterpret_cast<const float4*>(tsum + o)[i];
        const float4 rw = reinterpret_cast<const float4*>(row)[i];
        const float4 gp = reinterpret_cast<const float4*>(g_prev + o)[i];
        const float4 zz = reinterpret_cast<const float4*>(z + o)[i];
        const float4 ns = make_float4(ts.x - rw.x + gn.x, ts.y - rw.y + gn.y, ts.z - rw.z + gn.z, ts.w - rw.w + gn.w);
        reinterpret_cast<float4*>(row)[i] = gn;
        reinterpret_cast<float4*>(tsum + o)[i] = ns;
        reinterpret_cast<float4*>(z + o)[i] = make_float4(zz.x - s * (gn.x - gp.x + ns.x * inv_h), zz.y - s * (gn.y - gp.y + ns.y * inv_h),
                                                          zz.z - s * (gn.z - gp.z + ns.z * inv_h), zz.w - s * (gn.w - gp.w + ns.w * inv_h));
        reinterpret_cast<float4*>(g_prev + o)[i] = gn;             // prev_stoch = grad_history[rand_ind]  (pnp_saga.py:72)
    }
    for (long long i = 4 * n4 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float gn = g_new[o + i];
        const float ns = tsum[o + i] - row[i] + gn;
        row[i] = gn;
        tsum[o + i] = ns;
        z[o + i] -= s * (gn - g_prev[o + i] + ns * inv_h);
        g_prev[o + i] = gn;
    }
}

// table rows all start as the same gradient (algorithms/pnp_saga.py:28): table[h] = g0, tsum = hist*g0
__global__ void __launch_bounds__(256)
k_saga_init(const float* __restrict__ g0, float* __restrict__ table, float* __restrict__ tsum, long long n,
            long long img_stride, int hist) {
    const int img = blockIdx.y;
    const long long o = (long long)img * img_stride;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        const float g = g0[o + i];
        for (int h = 0; h < hist; ++h) table[((long long)img * hist + h) * img_stride + i] = g;
        tsum[o + i] = g * (float)hist;
    }
}

// rank -> k-space position of host-drawn minibatches, in place (mb_source='host' with the support list resident in HBM:
// the host draws WHICH measurements, the device resolves where they sit): rows of `stride` ints, `count` ranks each
__global__ void __launch_bounds__(256)
k_gather_support(int* __restrict__ idx, const int* __restrict__ support, int count, long long stride) {
    int* row = idx + (long long)blockIdx.y * stride;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x) row[i] = __ldg(support + row[i]);
}

__global__ void k_advance_by(int* __restrict__ counters, int n, int delta) {
    if (threadIdx.x < n) counters[threadIdx.x] += delta;
}

__global__ void k_advance(int* __restrict__ counters, int n, float* __restrict__ x, float factor) {
    if (threadIdx.x < n) counters[threadIdx.x] += 1;
    if (x && threadIdx.x == 0) *x *= factor;
}

}  // namespace pnp
