// exp_streams.cu — memory-only ceilings for the step kernel's access pattern (no env arithmetic).
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o build/exp_streams profiles/exp_streams.cu
// Patterns (all move 156 B per "env": 54 B read, 102 B written, 2^20 envs per launch, 4 rotating
// shards so nothing is re-read from L2):
//   copy   : plain float4 copy of the same number of bytes (1:1 read:write)          -> the peak
//   soa    : 9 read streams + 10 write streams laid out like MgState/MgOut (SoA)      -> our layout
//   tile   : the same bytes, but each warp's 64 envs contiguous (AoSoA, 3456 B read / 6528 B written)
//   rw12   : two streams only, 1:2 read:write                                         -> mix effect
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

constexpr int64_t N = 1 << 20;
constexpr int SH = 4;

__global__ void __launch_bounds__(128, 8) k_copy(const float4 *__restrict__ in, float4 *__restrict__ out, int64_t n4) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n4; i += stride) out[i] = in[i];
}

struct Soa {
    const double *r[6]; const uint32_t *meta; const uint8_t *a1, *a2;
    double *w[6]; uint32_t *wmeta; float *obs, *rew; uint8_t *done, *info;
};
__global__ void __launch_bounds__(128, 8) k_soa(Soa s) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t e0 = warp * 64 + lane * 2;
    double2 v[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) v[k] = *reinterpret_cast<const double2 *>(s.r[k] + e0);
    uint2 m = *reinterpret_cast<const uint2 *>(s.meta + e0);
    uchar2 b1 = *reinterpret_cast<const uchar2 *>(s.a1 + e0), b2 = *reinterpret_cast<const uchar2 *>(s.a2 + e0);
    m.x += b1.x + b2.y; m.y += b1.y + b2.x;
#pragma unroll
    for (int k = 0; k < 6; ++k) { v[k].x += 1.0; *reinterpret_cast<double2 *>(s.w[k] + e0) = v[k]; }
    *reinterpret_cast<uint2 *>(s.wmeta + e0) = m;
    float4 f = make_float4((float)v[0].x, (float)v[1].x, (float)v[2].y, (float)v[3].y);
    __stcs(reinterpret_cast<float4 *>(s.rew + 2 * e0), f);
    __stcs(reinterpret_cast<uchar2 *>(s.done + e0), b1);
    __stcs(reinterpret_cast<uchar2 *>(s.info + e0), b2);
    float4 *o = reinterpret_cast<float4 *>(s.obs + warp * 640);
#pragma unroll
    for (int k = 0; k < 5; ++k) __stcs(o + lane + 32 * k, f);
}

// tile layout: per warp one contiguous read tile (3456 B = 216 float4) and one write tile (6528 B = 408 float4)
__global__ void __launch_bounds__(128, 8) k_tile(const float4 *__restrict__ in, float4 *__restrict__ st,
                                                  float4 *__restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const float4 *r = in + warp * 216;
    float4 acc[7];
#pragma unroll
    for (int k = 0; k < 7; ++k) acc[k] = (lane + 32 * k < 216) ? r[lane + 32 * k] : make_float4(0, 0, 0, 0);
    float4 *ws = st + warp * 208;          // state write-back tile: 3328 B
#pragma unroll
    for (int k = 0; k < 7; ++k) if (lane + 32 * k < 208) { acc[k].x += 1.f; ws[lane + 32 * k] = acc[k]; }
    float4 *wo = out + warp * 200;         // outputs tile: 3200 B
#pragma unroll
    for (int k = 0; k < 7; ++k) if (lane + 32 * k < 200) __stcs(wo + lane + 32 * k, acc[k]);
}

__global__ void __launch_bounds__(128, 8) k_rw12(const float4 *__restrict__ in, float4 *__restrict__ out, int64_t n4) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    float4 v = in[i];
    out[2 * i] = v; v.x += 1.f; out[2 * i + 1] = v;
}

template <typename F> float time_it(F launch, int iters) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int i = 0; i < 20; ++i) launch(i);
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    for (int i = 0; i < iters; ++i) launch(i);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    return ms / iters;
}

int main() {
    const double bytes = 156.0 * N;
    const int iters = 2000;
    // ---- copy
    {
        const int64_t n4 = (int64_t)(78 * N / 16);   // 78 B read + 78 B written per env
        float4 *in[SH], *out[SH];
        for (int s = 0; s < SH; ++s) { CK(cudaMalloc(&in[s], n4 * 16)); CK(cudaMalloc(&out[s], n4 * 16)); CK(cudaMemset(in[s], 0, n4 * 16)); }
        float ms = time_it([&](int i) { k_copy<<<148 * 8, 128>>>(in[i % SH], out[i % SH], n4); }, iters);
        printf("copy  (1:1)            : %7.2f us  %7.1f GB/s\n", ms * 1e3, bytes / ms / 1e6);
        float ms2 = time_it([&](int i) { k_copy<<<(unsigned)((n4 + 127) / 128), 128>>>(in[i % SH], out[i % SH], n4); }, iters);
        printf("copy  (1:1, 1 elt/thr) : %7.2f us  %7.1f GB/s\n", ms2 * 1e3, bytes / ms2 / 1e6);
        for (int s = 0; s < SH; ++s) { cudaFree(in[s]); cudaFree(out[s]); }
    }
    // ---- rw12
    {
        const int64_t n4 = (int64_t)(52 * N / 16);   // 52 B read, 104 B written
        float4 *in[SH], *out[SH];
        for (int s = 0; s < SH; ++s) { CK(cudaMalloc(&in[s], n4 * 16)); CK(cudaMalloc(&out[s], 2 * n4 * 16)); CK(cudaMemset(in[s], 0, n4 * 16)); }
        float ms = time_it([&](int i) { k_rw12<<<(unsigned)((n4 + 127) / 128), 128>>>(in[i % SH], out[i % SH], n4); }, iters);
        printf("rw12  (1:2, 2 streams) : %7.2f us  %7.1f GB/s\n", ms * 1e3, bytes / ms / 1e6);
        for (int s = 0; s < SH; ++s) { cudaFree(in[s]); cudaFree(out[s]); }
    }
    // ---- soa (in place state like the real kernel)
    {
        Soa s[SH];
        for (int k = 0; k < SH; ++k) {
            for (int j = 0; j < 6; ++j) { double *p; CK(cudaMalloc(&p, N * 8)); CK(cudaMemset(p, 0, N * 8)); s[k].r[j] = p; s[k].w[j] = p; }
            uint32_t *m; CK(cudaMalloc(&m, N * 4)); CK(cudaMemset(m, 0, N * 4)); s[k].meta = m; s[k].wmeta = m;
            uint8_t *a; CK(cudaMalloc(&a, N)); CK(cudaMemset(a, 1, N)); s[k].a1 = a; CK(cudaMalloc(&a, N)); CK(cudaMemset(a, 2, N)); s[k].a2 = a;
            CK(cudaMalloc(&s[k].obs, N * 40)); CK(cudaMalloc(&s[k].rew, N * 8)); CK(cudaMalloc(&s[k].done, N)); CK(cudaMalloc(&s[k].info, N));
        }
        float ms = time_it([&](int i) { k_soa<<<(unsigned)(N / 256), 128>>>(s[i % SH]); }, iters);
        printf("soa   (19 streams)     : %7.2f us  %7.1f GB/s\n", ms * 1e3, bytes / ms / 1e6);
    }
    // ---- tile
    {
        const int64_t warps = N / 64;
        float4 *in[SH], *st[SH], *out[SH];
        for (int s = 0; s < SH; ++s) {
            CK(cudaMalloc(&in[s], warps * 216 * 16)); CK(cudaMemset(in[s], 0, warps * 216 * 16));
            CK(cudaMalloc(&st[s], warps * 208 * 16)); CK(cudaMalloc(&out[s], warps * 200 * 16));
        }
        float ms = time_it([&](int i) { k_tile<<<(unsigned)(N / 256), 128>>>(in[i % SH], st[i % SH], out[i % SH]); }, iters);
        printf("tile  (AoSoA, 3 streams): %7.2f us  %7.1f GB/s\n", ms * 1e3, bytes / ms / 1e6);
    }
    return 0;
}
