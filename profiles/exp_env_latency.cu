// exp_env_latency.cu — how long ONE warp needs for one env step (the policy kernels' fused epilogue runs the env at
// 1-2 warps per SM sub-partition, not at the step kernel's 8): cycles per env_step_batch<PVP, E> call for E envs per
// thread and W warps per block (one block per SM), state in registers, and the latency of a dependent DFMA chain.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o build/tmp/exp_env_latency profiles/exp_env_latency.cu
#include <cstdio>
#include "../merging_gym_b200/csrc/merge_device.cuh"
using namespace mg;

template <int E>
__global__ void env_chain(long long *out, int iters, double seed) {
    EnvRegs env[E];
    int a1[E], a2[E]; bool bad[E];
    for (int i = 0; i < E; ++i) { reset_regs(env[i]); env[i].p1 += seed * (threadIdx.x + i); a1[i] = (threadIdx.x + i) % 5; a2[i] = (threadIdx.x * 3 + i) % 5; bad[i] = false; }
    MgRewards rw{2.0, 1.0, -10.0, 0.001, 0.0};
    StepResult res[E];
    float acc = 0.f;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        env_step_batch<true, E, true>(env, a1, a2, bad, rw, res);
        for (int i = 0; i < E; ++i) { acc += res[i].obs[0] + res[i].obs[1] + res[i].r1; a1[i] = (a1[i] + (int)res[i].info + 1) % 5; if (res[i].done) reset_regs(env[i]); }
    }
    long long t1 = clock64();
    if (threadIdx.x % 32 == 0) out[blockIdx.x * 32 + threadIdx.x / 32] = t1 - t0;
    if (acc == 123.456f) out[0] = 0;
}
__global__ void dfma_chain(long long *out, int iters, double a, double b) {
    double x = a;
    long long t0 = clock64();
#pragma unroll 16
    for (int it = 0; it < iters; ++it) x = fma(x, b, a);
    long long t1 = clock64();
    if (threadIdx.x == 0) out[0] = t1 - t0;
    if (x == 123.0) out[1] = 0;
}
__global__ void ffma_chain(long long *out, int iters, float a, float b) {
    float x = a;
    long long t0 = clock64();
#pragma unroll 16
    for (int it = 0; it < iters; ++it) x = fmaf(x, b, a);
    long long t1 = clock64();
    if (threadIdx.x == 0) out[0] = t1 - t0;
    if (x == 123.0f) out[1] = 0;
}
int main() {
    long long *d; cudaMalloc(&d, 148 * 32 * 8);
    long long h[32];
    const int iters = 200;
    dfma_chain<<<1, 32>>>(d, 4096, 1.0000001, 0.999999); cudaMemcpy(h, d, 8, cudaMemcpyDeviceToHost);
    printf("dependent DFMA: %.1f cycles\n", h[0] / 4096.0);
    ffma_chain<<<1, 32>>>(d, 4096, 1.0000001f, 0.999999f); cudaMemcpy(h, d, 8, cudaMemcpyDeviceToHost);
    printf("dependent FFMA: %.1f cycles\n", h[0] / 4096.0);
    for (int w : {1, 4, 8, 16, 32}) {
#define RUN(E) env_chain<E><<<1, 32 * w>>>(d, iters, 1e-3); cudaMemcpy(h, d, 8 * 32, cudaMemcpyDeviceToHost); \
        { long long mx = 0; for (int i = 0; i < w; ++i) mx = h[i] > mx ? h[i] : mx; printf("warps/SM %2d  E=%d: %8.0f cycles per call  (%.0f per env-step of a thread)\n", w, E, mx / (double)iters, mx / (double)iters / E); }
        RUN(1) RUN(2) RUN(4)
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
