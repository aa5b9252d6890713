// exp_tcgen05_issue.cu — is the per-instruction cost of tcgen05.mma (76 cycles for tf32 M128 N<=128 K8, flat in N)
// a limit of the ISSUING THREAD or of the tensor pipe?  W warps (one elected lane each) issue back-to-back MMAs on
// fixed operands into separate accumulators; reports cycles per MMA for W = 1, 2, 4.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o build/exp_tcgen05_issue profiles/exp_tcgen05_issue.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(128u >> 4) << 16) | ((uint64_t)(256u >> 4) << 32) | ((uint64_t)1 << 46);
}

__global__ void __launch_bounds__(128, 1) k_issue(int N, int W, int reps, long long *out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) unsigned long long mbar[4];
    __shared__ uint32_t tmem_base_s;
    __shared__ long long t_start[4], t_end[4];
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    for (int i = t; i < 48 * 1024 / 4; i += 128) reinterpret_cast<float *>(smem)[i] = 0.f;
    if (t == 0) {
        for (int w = 0; w < 4; ++w) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar[w])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tb = tmem_base_s;
    if (warp < W && lane == 0) {
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint64_t da = make_desc(smem_u32(smem)), db = make_desc(smem_u32(smem + 16384));
        const uint32_t d = tb + (uint32_t)warp * 128u;
        t_start[warp] = clock64();
        for (int r = 0; r < reps; ++r)
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"(1u) : "memory");
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar[warp])) : "memory");
        const long long t_issued = clock64();
        uint32_t done = 0;
        for (int it = 0; it < (1 << 26) && !done; ++it)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(done) : "r"(smem_u32(&mbar[warp])), "r"(0u) : "memory");
        t_end[warp] = done ? clock64() : -1;
        out[8 + warp] = t_issued - t_start[warp];
    }
    __syncthreads();
    if (t == 0) {
        long long s = t_start[0], e = t_end[0];
        for (int w = 1; w < W; ++w) { s = min(s, t_start[w]); e = max(e, t_end[w]); }
        out[0] = e - s;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512u));
}

int main() {
    long long *d, h[16];
    cudaMalloc(&d, sizeof h);
    cudaFuncSetAttribute(k_issue, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    const int reps = 4000;
    for (int N : {64, 112})
        for (int W : {1, 2, 4}) {
            k_issue<<<1, 128, 64 * 1024>>>(N, W, reps, d);
            cudaError_t e = cudaDeviceSynchronize();
            if (e) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
            cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
            printf("N=%3d issuing warps=%d : %.1f cycles per MMA overall (%.1f per MMA per warp); issue loop alone %.1f cycles per MMA\n", N, W,
                   (double)h[0] / (reps * W), (double)h[0] / reps, (double)h[8] / reps);
        }
    return 0;
}
