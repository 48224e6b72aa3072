// Micro-benchmark: can tensor memory (TMEM) serve as a per-thread scratch store next to shared memory?
// Measures tcgen05.ld / tcgen05.st throughput per SM (8 warps, 32x32b.x16 shapes) alone and concurrently with
// 128-bit shared-memory loads, to decide whether N^{n-1} / F / coefficient rows of the KS kernel can leave the
// shared-memory crossbar.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_probe tools/tmem_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define LD16(r, addr)                                                                                                   \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];" \
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),       \
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])  \
                 : "r"(addr))
#define ST16(r, addr)                                                                                                   \
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%16], {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15};" \
                 ::"r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),               \
                   "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(addr) \
                 : "memory")

__global__ void __launch_bounds__(256, 1) probe(int mode, int iters, long long* cyc, unsigned* sink) {
    __shared__ uint32_t slot;
    extern __shared__ __align__(16) unsigned char dyn[];
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = slot + ((uint32_t)(32 * (warp % 4)) << 16) + (warp / 4) * 256;
    uint32_t r[4][16];
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int i = 0; i < 16; ++i) r[q][i] = threadIdx.x * 64 + q * 16 + i;
    uint4* sm = reinterpret_cast<uint4*>(dyn) + threadIdx.x * 17;     // 16 x 16 B per thread, odd stride
    for (int i = 0; i < 16; ++i) sm[i] = make_uint4(i, lane, warp, 0);
#pragma unroll
    for (int q = 0; q < 4; ++q) ST16(r[q], base + q * 16);
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    __syncthreads();
    unsigned acc = 0;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if (mode == 0 || mode == 2 || mode == 4) {
            _Pragma("unroll") for (int q = 0; q < 4; ++q) LD16(r[q], base + q * 16);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
_Pragma("unroll")
            for (int q = 0; q < 4; ++q) acc ^= r[q][3] + r[q][12];
        }
        if (mode == 1 || mode == 2) {
_Pragma("unroll")
            for (int q = 0; q < 4; ++q) { r[q][0] += it; ST16(r[q], base + q * 16); }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        }
        if (mode == 3 || mode == 4) {
            uint4 v[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = sm[i];
#pragma unroll
            for (int i = 0; i < 16; ++i) acc ^= v[i].x + v[i].w;
            sm[it & 15].w = acc;
        }
    }
    const long long t1 = clock64();
    __syncthreads();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(slot) : "memory");
}

int main() {
    const int grid = 148, iters = 2000;
    long long* cyc; unsigned* sink;
    cudaMalloc(&cyc, grid * sizeof(long long));
    cudaMalloc(&sink, grid * 256 * sizeof(unsigned));
    const size_t dyn = 256 * 17 * 16;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
    const char* names[] = {"tmem ld", "tmem st", "tmem ld+st", "smem ld128", "tmem ld + smem ld128"};
    for (int warps = 4; warps <= 8; warps += 4)
        for (int mode = 0; mode < 5; ++mode) {
            probe<<<grid, warps * 32, dyn>>>(mode, 10, cyc, sink);
            probe<<<grid, warps * 32, dyn>>>(mode, iters, cyc, sink);
            if (cudaDeviceSynchronize() != cudaSuccess) { printf("error %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
            long long h[grid];
            cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
            double c = (double)h[0] / iters;
            const double bytes = warps * 32 * 256.0;     // per iteration and leg
            printf("%d warps  %-22s %8.1f cycles/iter  -> %6.1f B/clk/SM per leg\n", warps, names[mode], c, bytes / c);
        }
    return 0;
}
