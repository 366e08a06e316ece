// mma_probe2.cu -- development probe: what a tcgen05.commit / an mbarrier try_wait on a completed phase / the per-slab
// bookkeeping of conv_halo_kernel cost the MMA-issuing thread between groups of 19 tcgen05.mma (N = 96, the kd-merged slab).
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc64(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4); d |= (uint64_t)1 << 16; d |= (uint64_t)((8 * 64) >> 4) << 32; d |= (uint64_t)1 << 46; d |= (uint64_t)4 << 61;
    return d;
}
__device__ __forceinline__ void mma(uint32_t t, uint64_t a, uint64_t b, uint32_t idesc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(t), "l"(a), "l"(b), "r"(idesc), "r"(1u) : "memory");
}
__device__ __forceinline__ void commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void wait(uint64_t* bar, uint32_t parity) {
    asm volatile("{\n\t.reg .pred P1;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t@P1 bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// MODE bit0: two commits per group; bit1: two try_waits on completed barriers per group; bit2: elect + syncwarp + fence per group;
// bit3: wait for the PREVIOUS group's commit before issuing (ring of 4 barriers: models sempty/tfull consumers)
template <int MODE, int NMMA>
__global__ void probe(int groups, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bars[8], done, ready;
    __shared__ uint32_t tmem_base_s;
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 8; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[i])));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&done)));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&ready)));
        asm volatile("fence.mbarrier_init.release.cluster;");
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&ready)) : "memory");   // phase 0 of `ready` complete
    }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base_s)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = tmem_base_s;
    constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(96 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    if (threadIdx.x < 32) {
        const uint64_t ad = desc64(smem_u32(smem)), bd = desc64(smem_u32(smem + 48 * 1024));
        long long t0 = clock64();
        for (int g = 0; g < groups; ++g) {
            if (MODE & 2) { wait(&ready, 0); wait(&ready, 0); }
            if (MODE & 4) {
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                __syncwarp();
            }
            uint32_t pred = 1;
            if (MODE & 4) asm volatile("{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\telect.sync rx|px, 0xffffffff;\n\tselp.u32 %0, 1, 0, px;\n\t}" : "=r"(pred)::"memory");
            else pred = threadIdx.x == 0;
            if (pred) {
#pragma unroll
                for (int u = 0; u < NMMA; ++u) mma(tmem + (uint32_t)((g & 3) * 96), ad + (uint64_t)(2 * (u & 1)) + (uint64_t)(32 * (u % 9)), bd + (uint64_t)(2 * (u & 1)), idesc);
                if (MODE & 1) { commit(&bars[g & 3]); commit(&bars[4 + (g & 3)]); }
            }
            if (MODE & 4) __syncwarp();
        }
        if (threadIdx.x == 0) {
            commit(&done);
            long long t1 = clock64();
            wait(&done, 0);
            long long t2 = clock64();
            out[0] = t1 - t0; out[1] = t2 - t0;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
}
template <int MODE, int NMMA>
void run(long long* d) {
    long long h[2];
    const int groups = 128;
    cudaFuncSetAttribute(probe<MODE, NMMA>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    for (int rep = 0; rep < 2; ++rep) {
        probe<MODE, NMMA><<<1, 128, 100 * 1024>>>(groups, d);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); exit(1); }
    }
    cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    printf("mode %d (commit %d, try_wait %d, elect/sync %d), %2d MMAs N=96 per group: issue %.0f cycles/group, complete %.0f cycles/group (MMA floor %.0f)\n",
           MODE, MODE & 1, (MODE >> 1) & 1, (MODE >> 2) & 1, NMMA, (double)h[0] / groups, (double)h[1] / groups, NMMA * 56.2);
}
int main() {
    long long* d;
    cudaMalloc(&d, 16);
    run<0, 19>(d); run<1, 19>(d); run<2, 19>(d); run<3, 19>(d); run<4, 19>(d); run<7, 19>(d);
    run<0, 4>(d); run<1, 4>(d); run<7, 4>(d);
    run<0, 1>(d); run<1, 1>(d); run<2, 1>(d); run<4, 1>(d); run<7, 1>(d);
    return 0;
}
