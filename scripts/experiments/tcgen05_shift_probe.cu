#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
// fill TMEM[lane][col] = lane*1000 + col for 128 lanes x 64 cols, shift down, read back
__global__ void k(float* out, int nshift, int base_lane, int col0) {
    __shared__ unsigned tmem_base;
    __shared__ unsigned long long bar;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(64u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const unsigned tmem = tmem_base;
    const unsigned t0 = tmem + ((unsigned)(warp * 32) << 16);
    // each warp stores its 32 lanes x 64 columns
    for (int c = 0; c < 64; c += 4) {
        unsigned v[4];
        for (int i = 0; i < 4; ++i) v[i] = __float_as_uint((float)((warp * 32 + lane) * 1000 + c + i));
        asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(t0 + c), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]));
    }
    asm volatile("tcgen05.wait::st.sync.aligned;");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    if (threadIdx.x == 0) {
        for (int s = 0; s < nshift; ++s)
            asm volatile("tcgen05.shift.cta_group::1.down [%0];" ::"r"(tmem + ((unsigned)base_lane << 16) + col0));
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    // wait
    {
        unsigned done = 0;
        while (!done)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
    }
    asm volatile("tcgen05.fence::after_thread_sync;");
    for (int c = 0; c < 64; c += 4) {
        unsigned r[4];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(t0 + c));
        asm volatile("tcgen05.wait::ld.sync.aligned;");
        for (int i = 0; i < 4; ++i) out[(warp * 32 + lane) * 64 + c + i] = __uint_as_float(r[i]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64u));
    }
}
int main(int argc, char** argv) {
    float* d; cudaMalloc(&d, 128 * 64 * 4);
    float* h = (float*)malloc(128 * 64 * 4);
    int cfgs[][3] = {{1, 0, 0}, {2, 0, 0}, {1, 0, 8}, {1, 32, 0}, {1, 0, 4}};
    for (auto& c : cfgs) {
        k<<<1, 128>>>(d, c[0], c[1], c[2]);
        cudaError_t e = cudaDeviceSynchronize();
        printf("nshift %d base_lane %d col0 %d: %s\n", c[0], c[1], c[2], cudaGetErrorString(e));
        if (e != cudaSuccess) return 1;
        cudaMemcpy(h, d, 128 * 64 * 4, cudaMemcpyDeviceToHost);
        // report, for a few columns, which (lane, col) each row now holds
        for (int col : {0, 7, 8, 15, 16, 63}) {
            printf("  col %2d: ", col);
            for (int r : {0, 1, 2, 31, 32, 33, 63, 64, 65, 95, 96, 127}) printf("r%d=%g ", r, h[r * 64 + col]);
            printf("\n");
        }
    }
    return 0;
}
