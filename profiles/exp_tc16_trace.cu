// exp_tc16_trace.cu — timeline of the hand-over points inside mlp_act_tc16_kernel (CTA 0): where do the converter warps,
// the MMA-issuing warp and the epilogue spend their cycles?  Includes the product kernel with tracing on.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -Iinclude -o build/exp_tc16_trace profiles/exp_tc16_trace.cu
#define MG_TC16_TRACE 1
#include "../merging_gym_b200/csrc/mlp_tc16_kernels.cu"
#include <vector>

int main() {
    const int64_t n = 1 << 18;
    float *obs, *b2, *w3, *b3;
    unsigned char *blob;
    uint8_t *act;
    cudaMalloc(&obs, n * 10 * 4); cudaMemset(obs, 0, n * 10 * 4);
    cudaMalloc(&blob, 64 + mgtc16::W_BYTES); cudaMemset(blob, 0, 64 + mgtc16::W_BYTES);
    cudaMalloc(&b2, 100 * 4); cudaMemset(b2, 0, 100 * 4);
    cudaMalloc(&w3, 5 * 100 * 4); cudaMemset(w3, 0, 5 * 100 * 4);
    cudaMalloc(&b3, 5 * 4); cudaMemset(b3, 0, 5 * 4);
    cudaMalloc(&act, n);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto run = [&] { return mg_mlp_act_tc16_launch(10, 5, false, obs, nullptr, n, 0, blob, b2, w3, b3, act, nullptr, 0, false); };
    for (int i = 0; i < 3; ++i) run();
    cudaEventRecord(e0);
    for (int i = 0; i < 20; ++i) run();
    cudaEventRecord(e1);
    if (cudaError_t e = cudaDeviceSynchronize()) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("kernel %.2f us (with tracing)\n", ms * 1e3 / 20);
    const int G = mgtc16::TRACE_G;
    std::vector<long long> c(G * 6), m(G * 4), ep(16 * 3), l1(32 * 2);
    cudaMemcpyFromSymbol(c.data(), mgtc16::g_trace_conv, c.size() * 8);
    cudaMemcpyFromSymbol(m.data(), mgtc16::g_trace_mma, m.size() * 8);
    cudaMemcpyFromSymbol(ep.data(), mgtc16::g_trace_epi, ep.size() * 8);
    cudaMemcpyFromSymbol(l1.data(), mgtc16::g_trace_l1, l1.size() * 8);
    const long long t0 = m[0];
    printf("# it : conv[top l1ready ld_done converted slot_free arrived]  mma[top full_seen issued after_poll]  (cycles since the MMA warp's first wait)\n");
    for (int g = 39; g < 39 + 28; ++g)
        printf("%3d : %7lld %7lld %7lld %7lld %7lld %7lld | %7lld %7lld %7lld %7lld\n", g, c[g * 6] - t0, c[g * 6 + 1] - t0, c[g * 6 + 2] - t0,
               c[g * 6 + 3] - t0, c[g * 6 + 4] - t0, c[g * 6 + 5] - t0, m[g * 4] - t0, m[g * 4 + 1] - t0, m[g * 4 + 2] - t0, m[g * 4 + 3] - t0);
    for (int t = 0; t < 10; ++t)
        printf("tile %d: layer1 half A start %lld issued %lld, half B start %lld issued %lld | epilogue wait_start %lld wait_end %lld done %lld\n", t,
               l1[t * 4] - t0, l1[t * 4 + 1] - t0, l1[t * 4 + 2] - t0, l1[t * 4 + 3] - t0, ep[t * 3] - t0, ep[t * 3 + 1] - t0, ep[t * 3 + 2] - t0);
    return 0;
}
