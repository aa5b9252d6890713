// exp_lds_broadcast.cu — shared-memory pipe cost (cycles per warp instruction at saturation, one SM) of the load
// patterns the Q-network kernels use for their weights: warp-uniform LDS.32 / .64 / .128, LDS.64 with 4 distinct
// addresses per warp (the fragment epilogue), and fully coalesced LDS.128 for reference.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o build/exp_lds_broadcast profiles/exp_lds_broadcast.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(512, 1) k_lds(int reps, long long *out, float *sink) {
    __shared__ __align__(16) float buf[8192];
    for (int i = threadIdx.x; i < 8192; i += blockDim.x) buf[i] = (float)i;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float acc = 0.f;
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            const int base = ((r * 16 + u) * 4 + warp * 64) & 4095;
            if (MODE == 0) acc += buf[base];                                                            // uniform LDS.32
            if (MODE == 1) { const float2 v = *reinterpret_cast<const float2 *>(&buf[base]); acc += v.x + v.y; }              // uniform LDS.64
            if (MODE == 2) { const float4 v = *reinterpret_cast<const float4 *>(&buf[base]); acc += v.x + v.y + v.z + v.w; }  // uniform LDS.128
            if (MODE == 3) { const float2 v = *reinterpret_cast<const float2 *>(&buf[base + 2 * (lane & 3)]); acc += v.x + v.y; }   // 4 addresses
            if (MODE == 4) { const float4 v = *reinterpret_cast<const float4 *>(&buf[(base + 4 * lane) & 8191]); acc += v.x + v.y + v.z + v.w; }  // coalesced
            if (MODE == 5) { const float4 v = *reinterpret_cast<const float4 *>(&buf[base + 4 * (lane & 3)]); acc += v.x + v.y + v.z + v.w; }    // LDS.128, 4 addresses
        }
    }
    const long long t1 = clock64();
    if (acc == 12345.678f) sink[0] = acc;
    if (threadIdx.x == 0) out[0] = t1 - t0;
}

int main() {
    long long *d, h;
    float *sink;
    cudaMalloc(&d, 8); cudaMalloc(&sink, 4);
    const int reps = 2000;
    const char *names[] = {"uniform LDS.32", "uniform LDS.64", "uniform LDS.128", "LDS.64, 4 addresses/warp", "coalesced LDS.128", "LDS.128, 4 addresses/warp"};
    for (int mode = 0; mode < 6; ++mode) {
        for (int threads : {32, 512}) {
            switch (mode) {
                case 0: k_lds<0><<<1, threads>>>(reps, d, sink); break;
                case 1: k_lds<1><<<1, threads>>>(reps, d, sink); break;
                case 2: k_lds<2><<<1, threads>>>(reps, d, sink); break;
                case 3: k_lds<3><<<1, threads>>>(reps, d, sink); break;
                case 4: k_lds<4><<<1, threads>>>(reps, d, sink); break;
                default: k_lds<5><<<1, threads>>>(reps, d, sink); break;
            }
            if (cudaError_t e = cudaDeviceSynchronize()) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
            cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
            const double per = (double)h / (reps * 16.0);
            printf("%-28s %2d warps: %6.2f cycles per LDS per warp -> %5.2f SM cycles per warp-instruction\n", names[mode], threads / 32, per,
                   per / (threads / 32));
        }
    }
    return 0;
}
