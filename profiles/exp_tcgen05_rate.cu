// exp_tcgen05_rate.cu — per-instruction cost of tcgen05.mma (cta_group::1, M = 128) for the shapes the
// Q-network kernels use: back-to-back MMAs on fixed operands, one CTA, cycles per MMA from clock64.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o build/exp_tcgen05_rate profiles/exp_tcgen05_rate.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(128u >> 4) << 16) | ((uint64_t)(256u >> 4) << 32) | ((uint64_t)1 << 46);
}

// mode 0: tf32 SS   1: tf32 TS (A in TMEM)   2: bf16 SS (K = 16)
__global__ void __launch_bounds__(128, 1) k_rate(int N, int mode, int reps, long long *out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) unsigned long long mbar;
    __shared__ uint32_t tmem_base_s;
    const int t = threadIdx.x, warp = t >> 5;
    for (int i = t; i < 48 * 1024 / 4; i += 128) reinterpret_cast<float *>(smem)[i] = 0.f;
    if (t == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)));
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
    if (t == 0) {
        const uint32_t fmt = mode == 2 ? 1u : 2u;   // bf16 = 1, tf32 = 2
        const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint64_t da = make_desc(smem_u32(smem)), db = make_desc(smem_u32(smem + 16384));
        const long long t0 = clock64();
        for (int r = 0; r < reps; ++r) {
            if (mode == 0)
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tb), "l"(da), "l"(db), "r"(idesc), "r"(1u) : "memory");
            else if (mode == 1)
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tb), "r"(tb + 300u), "l"(db), "r"(idesc), "r"(1u) : "memory");
            else
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tb), "l"(da), "l"(db), "r"(idesc), "r"(1u) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar)) : "memory");
        uint32_t done = 0;
        for (int it = 0; it < (1 << 26) && !done; ++it)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(done) : "r"(smem_u32(&mbar)), "r"(0u) : "memory");
        const long long t1 = clock64();
        out[0] = done ? (t1 - t0) : -1;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512u));
}

int main() {
    long long *d, h;
    cudaMalloc(&d, 8);
    cudaFuncSetAttribute(k_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    const char *names[] = {"tf32 SS (K=8)", "tf32 TS (K=8, A in TMEM)", "bf16 SS (K=16)"};
    const int reps = 4000;
    for (int mode = 0; mode < 3; ++mode)
        for (int N : {64, 112, 128, 224, 256}) {
            k_rate<<<1, 128, 64 * 1024>>>(N, mode, reps, d);
            cudaError_t e = cudaDeviceSynchronize();
            if (e) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
            cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
            const double cyc = (double)h / reps;
            const double flop = 2.0 * 128 * N * (mode == 2 ? 16 : 8);
            printf("%-26s N=%3d : %7.1f cycles/MMA  %7.0f FLOP/cycle/SM  (x148 SMs x1.965 GHz = %.0f TFLOP/s)\n", names[mode], N, cyc,
                   flop / cyc, flop / cyc * 148 * 1.965e9 / 1e12);
        }
    return 0;
}
