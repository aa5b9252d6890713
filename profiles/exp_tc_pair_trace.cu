// exp_tc_pair_trace.cu — hand-over timeline of the CTA-pair tensor-core MLP kernel (cluster 0): producers of both
// CTAs and the leader's MMA-issuing warp.  Includes the experimental pair kernel with tracing on; also times it.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -DMG_TC_TRACE=1 -Iinclude -o build/exp_tc_pair_trace profiles/exp_tc_pair_trace.cu
#define MG_TC_TRACE 1
#include "exp_mlp_tc_pair_kernels.cu"
#include <vector>

int main() {
    const int64_t n = 1 << 18;
    float *obs, *w1t, *b1, *w2, *b2, *w3, *b3;
    uint8_t *act;
    cudaMalloc(&obs, n * 10 * 4); cudaMemset(obs, 0, n * 10 * 4);
    cudaMalloc(&w1t, 10 * 200 * 4); cudaMemset(w1t, 0, 10 * 200 * 4);
    cudaMalloc(&b1, 200 * 4); cudaMemset(b1, 0, 200 * 4);
    cudaMalloc(&w2, 25 * 7168); cudaMemset(w2, 0, 25 * 7168);
    cudaMalloc(&b2, 100 * 4); cudaMemset(b2, 0, 100 * 4);
    cudaMalloc(&w3, 5 * 100 * 4); cudaMemset(w3, 0, 5 * 100 * 4);
    cudaMalloc(&b3, 5 * 4); cudaMemset(b3, 0, 5 * 4);
    cudaMalloc(&act, n);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 3; ++i) mg_mlp_act_tc_pair(obs, nullptr, n, 10, 5, w1t, b1, w2, b2, w3, b3, act, nullptr, 0u, 0);
    cudaEventRecord(e0);
    for (int i = 0; i < 20; ++i) mg_mlp_act_tc_pair(obs, nullptr, n, 10, 5, w1t, b1, w2, b2, w3, b3, act, nullptr, 0u, 0);
    cudaEventRecord(e1);
    if (cudaError_t e = cudaDeviceSynchronize()) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("kernel %.2f us (with tracing)\n", ms * 1e3 / 20);
    const int G = mgtc2::TRACE_G;
    std::vector<long long> p(2 * G * 4), m(G * 3);
    cudaMemcpyFromSymbol(p.data(), mgtc2::g_trace_prod, p.size() * 8);
    cudaMemcpyFromSymbol(m.data(), mgtc2::g_trace_mma, m.size() * 8);
    const long long t0 = m[0];
    printf("# g : leader producer[compute_start wait_start wait_end arrived] peer producer[...] | mma[wait_start wait_end issued]\n");
    for (int g = 75; g < 110; ++g) {
        printf("%3d :", g);
        for (int r = 0; r < 2; ++r) { for (int k = 0; k < 4; ++k) printf(" %7lld", p[(r * G + g) * 4 + k] - t0); printf(" |"); }
        printf(" %7lld %7lld %7lld\n", m[g * 3] - t0, m[g * 3 + 1] - t0, m[g * 3 + 2] - t0);
    }
    double comp[2] = {0, 0}, wait[2] = {0, 0}, st[2] = {0, 0}, mwait = 0, missue = 0, step = 0;
    int c = 0;
    for (int g = 75; g < 225; ++g, ++c) {
        for (int r = 0; r < 2; ++r) {
            const long long *q = &p[(r * G + g) * 4];
            comp[r] += q[1] - q[0]; wait[r] += q[2] - q[1]; st[r] += q[3] - q[2];
        }
        mwait += m[g * 3 + 1] - m[g * 3]; missue += m[g * 3 + 2] - m[g * 3 + 1]; step += m[(g + 1) * 3 + 2] - m[g * 3 + 2];
    }
    printf("avg cycles over K-steps 75..224 of cluster 0:\n");
    for (int r = 0; r < 2; ++r) printf("  producer CTA %d: compute %.0f  wait-empty %.0f  sts+fence+arrive %.0f\n", r, comp[r] / c, wait[r] / c, st[r] / c);
    printf("  mma warp: wait-full %.0f  issue %.0f   K-step period %.0f\n", mwait / c, missue / c, step / c);
    return 0;
}
