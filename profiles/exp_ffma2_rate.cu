// exp_ffma2_rate.cu — issue rate of Blackwell's packed FFMA2 (fma.rn.f32x2) against scalar FFMA, per SM
// sub-partition, for the operand pattern of the Q-network kernels (scalar activation x weight pair + accumulator
// pair), with 1, 2 and 4 warps per scheduler and 8 / 16 / 32 independent accumulators per thread.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o build/exp_ffma2_rate profiles/exp_ffma2_rate.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

template <int NACC, bool PACKED>
__global__ void k_rate(int reps, float seed, long long *out, float *sink) {
    float2 acc[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = make_float2(seed * i, seed + i);
    float a = seed * 1.0001f + threadIdx.x * 1e-9f;
    float2 b[4] = {make_float2(seed, 1.f + seed), make_float2(0.5f + seed, seed * 2), make_float2(seed * 3, 0.25f), make_float2(0.125f, seed)};
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) {
            if (PACKED) {
                acc[i] = __ffma2_rn(make_float2(a, a), b[i & 3], acc[i]);
            } else {
                acc[i].x = fmaf(a, b[i & 3].x, acc[i].x);
                acc[i].y = fmaf(a, b[i & 3].y, acc[i].y);
            }
        }
    }
    const long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += acc[i].x + acc[i].y;
    if (s == 12345.678f) sink[0] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
}

template <int NACC, bool PACKED>
void run(int warps, long long *d, float *sink) {
    const int reps = 2000;
    k_rate<NACC, PACKED><<<1, warps * 32>>>(reps, 1e-3f, d, sink);
    cudaDeviceSynchronize();
    long long h;
    cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    const double fma_per_clk_smsp = (double)reps * NACC * 2 * 32 * (warps / 4.0) / h;   // lanes x 2 FMAs per accumulator pair
    printf("%-6s acc=%2d warps/SMSP=%d : %6.1f FMA/clk/SMSP (peak 32)  %5.2f cycles per %s\n", PACKED ? "FFMA2" : "FFMA", NACC, warps / 4,
           fma_per_clk_smsp, (double)h / ((double)reps * NACC * (PACKED ? 1 : 2) * (warps / 4.0)), PACKED ? "FFMA2" : "FFMA");
}

int main() {
    long long *d; float *sink;
    cudaMalloc(&d, 8); cudaMalloc(&sink, 4);
    for (int warps : {4, 8, 16}) {
        run<8, true>(warps, d, sink); run<16, true>(warps, d, sink); run<32, true>(warps, d, sink);
        run<8, false>(warps, d, sink); run<16, false>(warps, d, sink); run<32, false>(warps, d, sink);
    }
    return 0;
}
