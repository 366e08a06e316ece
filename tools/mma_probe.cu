// mma_probe.cu -- development probe: issue rate of tcgen05.mma (M=128, K=16, bf16) as a function of N and of the number of
// independent TMEM accumulators that consecutive instructions rotate through.  Build: nvcc -gencode arch=compute_100a,code=sm_100a
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int ROWB>
__device__ __forceinline__ uint64_t desc(uint32_t saddr) {
    const uint64_t layout = ROWB == 128 ? 2 : 4;
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)((8 * ROWB) >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= layout << 61;
    return d;
}

template <int ROWB, int N, int NACC>
__global__ void probe(int iters, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base_s)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = tmem_base_s;
    constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    if (threadIdx.x == 0) {
        const uint32_t sa = smem_u32(smem), sb = smem_u32(smem + 48 * 1024);
        const uint64_t ad = desc<ROWB>(sa), bd = desc<ROWB>(sb);
        long long t0 = clock64();
        for (int i = 0; i < iters; i += 16) {
#pragma unroll
            for (int u = 0; u < 16; ++u) {
                asm volatile(
                    "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                    "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                    ::"r"(tmem + (uint32_t)((u % NACC) * N)), "l"(ad + (uint64_t)(2 * (u & 1))), "l"(bd + (uint64_t)(2 * (u & 1))), "r"(idesc), "r"(1u) : "memory");
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        long long t1 = clock64();
        asm volatile(
            "{\n\t.reg .pred P1;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], 0;\n\t@P1 bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(smem_u32(&bar)) : "memory");
        long long t2 = clock64();
        out[0] = t1 - t0;
        out[1] = t2 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
}

template <int ROWB, int N, int NACC>
void run(long long* d) {
    long long h[2];
    const int iters = 512;
    cudaFuncSetAttribute(probe<ROWB, N, NACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    for (int rep = 0; rep < 2; ++rep) {
        probe<ROWB, N, NACC><<<1, 128, 100 * 1024>>>(iters, d);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); exit(1); }
    }
    cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    printf("rowB %3d N %3d accumulators %d: issue %.1f cyc/mma, complete %.1f cyc/mma (ideal %d)\n", ROWB, N, NACC,
           (double)h[0] / iters, (double)h[1] / iters, N / 2);
}

int main() {
    long long* d;
    cudaMalloc(&d, 16);
    run<64, 32, 1>(d); run<64, 32, 2>(d); run<64, 32, 4>(d);
    run<64, 64, 1>(d); run<64, 64, 2>(d); run<64, 64, 4>(d);
    run<128, 32, 1>(d); run<128, 32, 4>(d);
    run<128, 64, 1>(d); run<128, 64, 4>(d);
    run<128, 128, 1>(d); run<128, 128, 2>(d); run<128, 128, 4>(d);
    run<128, 256, 1>(d); run<128, 256, 2>(d);
    // kd-merged widths of the halo / wgrad kernels
    run<64, 96, 1>(d); run<64, 96, 3>(d); run<128, 96, 3>(d); run<64, 160, 1>(d); run<64, 192, 2>(d); run<128, 192, 2>(d);
    return 0;
}
