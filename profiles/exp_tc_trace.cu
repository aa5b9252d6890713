// exp_tc_trace.cu — timeline of the hand-over points inside mlp_act_tc_kernel (CTA 0): where do the producer
// warps, the MMA-issuing warp and the epilogue spend their cycles?  Includes the product kernel with tracing on.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -DMG_TC_TRACE=1 -Iinclude -o build/exp_tc_trace profiles/exp_tc_trace.cu
#define MG_TC_TRACE 1
#include "../merging_gym_b200/csrc/mlp_tc_kernels.cu"
#include <vector>

int main() {
    const int64_t n = 1 << 18;
    float *obs, *w1t, *b1, *w2, *b2, *w3, *b3;
    uint8_t *act;
    cudaMalloc(&obs, n * 10 * 4); cudaMemset(obs, 0, n * 10 * 4);
    cudaMalloc(&w1t, 10 * 200 * 4); cudaMemset(w1t, 0, 10 * 200 * 4);
    cudaMalloc(&b1, 200 * 4); cudaMemset(b1, 0, 200 * 4);
    cudaMalloc(&w2, mgtc::B_BYTES); cudaMemset(w2, 0, mgtc::B_BYTES);
    cudaMalloc(&b2, 100 * 4); cudaMemset(b2, 0, 100 * 4);
    cudaMalloc(&w3, 5 * 100 * 4); cudaMemset(w3, 0, 5 * 100 * 4);
    cudaMalloc(&b3, 5 * 4); cudaMemset(b3, 0, 5 * 4);
    cudaMalloc(&act, n);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 3; ++i) mg_mlp_act_tc(obs, nullptr, n, 10, 5, w1t, b1, w2, b2, w3, b3, act, nullptr, 0u, 0);
    cudaEventRecord(e0);
    for (int i = 0; i < 20; ++i) mg_mlp_act_tc(obs, nullptr, n, 10, 5, w1t, b1, w2, b2, w3, b3, act, nullptr, 0u, 0);
    cudaEventRecord(e1);
    if (cudaError_t e = cudaDeviceSynchronize()) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("kernel %.2f us (with tracing)\n", ms * 1e3 / 20);
    const int G = mgtc::TRACE_G;
    std::vector<long long> p(G * 5), m(G * 8), ep(16 * 3);
    cudaMemcpyFromSymbol(p.data(), mgtc::g_trace_prod, p.size() * 8);
    cudaMemcpyFromSymbol(m.data(), mgtc::g_trace_mma, m.size() * 8);
    cudaMemcpyFromSymbol(ep.data(), mgtc::g_trace_epi, ep.size() * 8);
    const long long t0 = m[0];
    printf("# g : producer[compute_start wait_start wait_end fence_done arrive_done]  mma[wait_start wait_end issued]  (cycles since MMA warp's first wait)\n");
    for (int g = 50; g < 110; ++g)
        printf("%3d : %7lld %7lld %7lld %7lld %7lld | %7lld %7lld %7lld\n", g, p[g * 5] - t0, p[g * 5 + 1] - t0, p[g * 5 + 2] - t0,
               p[g * 5 + 3] - t0, p[g * 5 + 4] - t0, m[g * 8] - t0, m[g * 8 + 1] - t0, m[g * 8 + 2] - t0);
    double comp = 0, wait = 0, store = 0, arr = 0, mwait = 0, missue = 0, a2m = 0, m2p = 0, step = 0;
    int c = 0;
    for (int g = 75; g < 225; ++g, ++c) {
        comp += p[g * 5 + 1] - p[g * 5]; wait += p[g * 5 + 2] - p[g * 5 + 1]; store += p[g * 5 + 3] - p[g * 5 + 2];
        arr += p[g * 5 + 4] - p[g * 5 + 3]; mwait += m[g * 8 + 1] - m[g * 8]; missue += m[g * 8 + 2] - m[g * 8 + 1];
        a2m += m[g * 8 + 1] - p[g * 5 + 3];             // producer fence done -> MMA warp sees the slot full
        m2p += p[(g + 4) * 5 + 2] - m[g * 8 + 2];       // MMAs of K-step g issued -> producer of g+4 sees the slot empty
        step += m[(g + 1) * 8 + 2] - m[g * 8 + 2];
    }
    printf("avg cycles over K-steps 75..224 of CTA 0:\n  producer: compute %.0f  wait-empty %.0f  sts+fence %.0f  arrive %.0f\n"
           "  mma warp: wait-full %.0f  issue %.0f   K-step period %.0f\n  fence-done -> mma sees full %.0f   mma issued(g) -> producer(g+4) sees empty %.0f\n",
           comp / c, wait / c, store / c, arr / c, mwait / c, missue / c, step / c, a2m / c, m2p / c);
    for (int t = 2; t < 8; ++t)
        printf("epilogue tile %d: wait %lld  work %lld\n", t, ep[t * 3 + 1] - ep[t * 3], ep[t * 3 + 2] - ep[t * 3 + 1]);
    return 0;
}
