// exp_tcgen05_pair_rate.cu — cost of a tf32 tcgen05.mma issued for a CTA PAIR (cta_group::2, M = 256: 128 rows per
// CTA, the B operand's N rows split between the two CTAs' shared memories) against the single-CTA figures of
// exp_tcgen05_issue.cu (M = 128: 60 cycles at N = 112, 123 at N = 224, 139 at N = 256).  Fixed zero operands, one
// cluster of two CTAs, the leader CTA's thread 0 issues back-to-back MMAs; cycles per MMA from clock64.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o build/exp_tcgen05_pair_rate profiles/exp_tcgen05_pair_rate.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(128u >> 4) << 16) | ((uint64_t)(256u >> 4) << 32) | ((uint64_t)1 << 46);
}
__device__ __forceinline__ void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) k_pair(int N, int reps, long long *out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) unsigned long long mbar;
    __shared__ uint32_t tmem_base_s;
    const int t = threadIdx.x, warp = t >> 5;
    uint32_t rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    for (int i = t; i < 48 * 1024 / 4; i += 128) reinterpret_cast<float *>(smem)[i] = 0.f;
    if (t == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    cluster_sync();
    if (warp == 0) {                                        // the same warp of BOTH CTAs allocates
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tb = tmem_base_s;
    long long t0 = 0, t1 = 0;
    if (rank == 0 && t == 0) {
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
        const uint64_t da = make_desc(smem_u32(smem)), db = make_desc(smem_u32(smem + 16384));
        t0 = clock64();
        for (int r = 0; r < reps; ++r)
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tb), "l"(da), "l"(db), "r"(idesc), "r"(1u) : "memory");
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(&mbar)), "h"((unsigned short)3) : "memory");
    }
    if (t == 0) {                                           // both CTAs wait for the multicast commit
        uint32_t done = 0;
        for (int it = 0; it < (1 << 26) && !done; ++it)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(done) : "r"(smem_u32(&mbar)), "r"(0u) : "memory");
        t1 = clock64();
        if (rank == 0) out[0] = done ? (t1 - t0) : -1;
        else out[1] = done ? 1 : -1;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512u));
}

int main() {
    long long *d, h[2];
    cudaMalloc(&d, 16);
    cudaFuncSetAttribute(k_pair, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    const int reps = 4000;
    for (int N : {64, 112, 128, 224, 256}) {
        cudaMemset(d, 0, 16);
        k_pair<<<2, 128, 64 * 1024>>>(N, reps, d);
        cudaError_t e = cudaDeviceSynchronize();
        if (e) { printf("N=%d: error %s\n", N, cudaGetErrorString(e)); return 1; }
        cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        const double cyc = (double)h[0] / reps;
        printf("cta_group::2 tf32 M=256 N=%3d K=8 : %7.1f cycles per MMA (peer CTA saw the commit: %lld)  -> %6.0f FLOP/cycle per SM\n", N, cyc, h[1],
               2.0 * 256 * N * 8 / cyc / 2);
    }
    return 0;
}
