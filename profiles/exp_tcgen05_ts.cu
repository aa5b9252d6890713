// exp_tcgen05_ts.cu — like exp_tcgen05.cu, but the A operand lives in TENSOR MEMORY (written with
// tcgen05.st by the thread that owns the row) instead of shared memory.  Stand-alone check of the
// building blocks a tensor-core version of
// the Q-network's 200x100 layer would need (sm_100a only):
//   C[128 x 112] = A[128 x 200] * B[112 x 200]^T,  fp32 data read as TF32 (kind::tf32), fp32 accumulate
// One CTA, operands written by the threads into the canonical no-swizzle K-major core-matrix layout,
// 25 tcgen05.mma (M128 N112 K8) issued by one thread, tcgen05.commit -> mbarrier, tcgen05.ld epilogue.
// Every wait is bounded and traps instead of hanging.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o build/exp_tcgen05 profiles/exp_tcgen05.cu
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>

constexpr int M = 128, N = 112, K = 200, KSTEPS = K / 8;
constexpr int A_STEP_BYTES = (M / 8) * 256;   // per k-step: 16 row groups x (2 core matrices x 128 B)
constexpr int B_STEP_BYTES = (N / 8) * 256;
constexpr int TMEM_COLS = 512;   // D: columns 0..127, A (tf32, one column per k): columns 256..455

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);          // start address, 16-byte units
    d |= (uint64_t)(lbo >> 4) << 16;                   // leading-dimension byte offset
    d |= (uint64_t)(sbo >> 4) << 32;                   // stride-dimension byte offset
    d |= (uint64_t)1 << 46;                            // descriptor version (sm_100)
    return d;                                          // layout_type (bits 61-63) = 0: no swizzle
}

__global__ void __launch_bounds__(128, 1)
k_gemm(const float *__restrict__ A, const float *__restrict__ B, float *__restrict__ C, int swap_lbo_sbo,
       int *__restrict__ status) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char *sA = smem;                              // KSTEPS * 4096
    unsigned char *sB = smem + KSTEPS * A_STEP_BYTES;      // KSTEPS * 3584
    __shared__ __align__(8) unsigned long long mbar;
    __shared__ uint32_t tmem_base_s;
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;

    // ---- operands -> canonical K-major core-matrix layout (8 rows x 16 bytes per core matrix) ----
    (void)sA;
    for (int i = t; i < N * K; i += 128) {
        const int n = i / K, k = i % K;
        const int off = (k / 8) * B_STEP_BYTES + (n / 8) * 256 + ((k % 8) / 4) * 128 + (n % 8) * 16 + (k % 4) * 4;
        *reinterpret_cast<float *>(sB + off) = B[i];
    }
    if (t == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> async proxy (MMA reads)
    __syncthreads();

    // ---- TMEM allocation by one warp ----
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)),
                     "r"((uint32_t)TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_s;

    // ---- A -> TMEM: thread t owns row t = TMEM lane t; K values go to consecutive columns ----
    {
        const uint32_t a_addr = tmem_base + 256u + ((uint32_t)(warp * 32) << 16);
        for (int ks = 0; ks < KSTEPS; ++ks) {
            uint32_t v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = __float_as_uint(A[t * K + ks * 8 + j]);
            asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(a_addr + ks * 8),
                         "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                         : "memory");
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }

    // ---- one thread issues the MMAs ----
    if (t == 0) {
        // instruction descriptor: D=f32 (1<<4), A=B=tf32 (2<<7, 2<<10), K-major both, N>>3 at bit 17, M>>4 at bit 24
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
        const uint32_t lbo = swap_lbo_sbo ? 256u : 128u, sbo = swap_lbo_sbo ? 128u : 256u;
        for (int ks = 0; ks < KSTEPS; ++ks) {
            const uint32_t ta = tmem_base + 256u + (uint32_t)(ks * 8);
            const uint64_t db = make_desc(smem_u32(sB + ks * B_STEP_BYTES), lbo, sbo);
            const uint32_t acc = ks > 0 ? 1u : 0u;
            asm volatile(
                "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tmem_base),
                "r"(ta), "l"(db), "r"(idesc), "r"(acc)
                : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar))
                     : "memory");
    }
    // ---- everyone waits for the accumulator (bounded) ----
    {
        uint32_t done = 0;
        for (int it = 0; it < (1 << 22) && !done; ++it) {
            asm volatile(
                "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                : "=r"(done)
                : "r"(smem_u32(&mbar)), "r"(0u)
                : "memory");
        }
        if (!done) {
            if (t == 0) *status = -1;
            __trap();
        }
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    // ---- epilogue: warp w owns TMEM lanes 32w..32w+31 = rows of C ----
    const int row = warp * 32 + lane;
    const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
#pragma unroll
    for (int c0 = 0; c0 < N; c0 += 16) {
        uint32_t v[16];
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
              "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
            : "r"(taddr + (uint32_t)c0));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int j = 0; j < 16; ++j) C[row * N + c0 + j] = __uint_as_float(v[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS));
    if (t == 0) *status = 1;
}

static float tf32_trunc(float x) {   // the tensor core reads the top 19 bits of an fp32 operand
    uint32_t u;
    memcpy(&u, &x, 4);
    u &= 0xFFFFE000u;
    memcpy(&x, &u, 4);
    return x;
}

int main() {
    std::vector<float> hA(M * K), hB(N * K), hC(M * N), ref(M * N), ref32(M * N);
    srand(1);
    for (auto &v : hA) v = (float)rand() / RAND_MAX * 2.f - 1.f;
    for (auto &v : hB) v = (float)rand() / RAND_MAX;
    for (int m = 0; m < M; ++m)
        for (int n = 0; n < N; ++n) {
            double s = 0, s32 = 0;
            for (int k = 0; k < K; ++k) {
                s += (double)tf32_trunc(hA[m * K + k]) * (double)tf32_trunc(hB[n * K + k]);
                s32 += (double)hA[m * K + k] * (double)hB[n * K + k];
            }
            ref[m * N + n] = (float)s;
            ref32[m * N + n] = (float)s32;
        }
    float *dA, *dB, *dC;
    int *dS;
    cudaMalloc(&dA, hA.size() * 4); cudaMalloc(&dB, hB.size() * 4); cudaMalloc(&dC, hC.size() * 4); cudaMalloc(&dS, 4);
    cudaMemcpy(dA, hA.data(), hA.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, hB.data(), hB.size() * 4, cudaMemcpyHostToDevice);
    const int smem = KSTEPS * (A_STEP_BYTES + B_STEP_BYTES);
    cudaFuncSetAttribute(k_gemm, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int swap = 0; swap < 1; ++swap) {
        cudaMemset(dC, 0, hC.size() * 4);
        cudaMemset(dS, 0, 4);
        k_gemm<<<1, 128, smem>>>(dA, dB, dC, swap, dS);
        cudaError_t e = cudaDeviceSynchronize();
        int st = 0;
        if (e != cudaSuccess) { printf("swap=%d: CUDA error %s\n", swap, cudaGetErrorString(e)); return 1; }
        cudaMemcpy(&st, dS, 4, cudaMemcpyDeviceToHost);
        cudaMemcpy(hC.data(), dC, hC.size() * 4, cudaMemcpyDeviceToHost);
        double err = 0, err32 = 0, mag = 0;
        for (int i = 0; i < M * N; ++i) {
            err = fmax(err, fabs((double)hC[i] - ref[i]));
            err32 = fmax(err32, fabs((double)hC[i] - ref32[i]));
            mag = fmax(mag, fabs((double)ref[i]));
        }
        printf("swap_lbo_sbo=%d status=%d  max|C-ref_tf32|=%.3e  max|C-ref_fp32|=%.3e  max|ref|=%.3f  C[0..3]=%f %f %f %f  ref=%f %f %f %f\n",
               swap, st, err, err32, mag, hC[0], hC[1], hC[2], hC[3], ref[0], ref[1], ref[2], ref[3]);
    }
    return 0;
}
