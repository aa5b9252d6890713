// exp_mlp_tc2_kernels.cu — EXPERIMENT, not part of the library build: an all-TMEM variant of the tensor-core
// Q-network forward (sm_100a: tcgen05 + TMEM).  Numerically correct (same 5e-6 error, same actions) but
// SLOWER than the shipped mlp_tc_kernels.cu: 119 us vs 89 us for 2^18 envs, because a TS-mode tf32 MMA
// (A in tensor memory) costs 103 cycles per instruction at N <= 128 against 76 for SS mode
// (profiles/exp_tcgen05_rate.cu) and this design issues 87 instead of 75 MMAs per tile.  Kept for the record;
// to try it, add it to merging_gym_b200/build.py::SOURCES and bind mg_mlp_act_tc2.
//
// Same operator and the same 3xTF32 error compensation as mlp_tc_kernels.cu, but designed around that
// kernel's profile (shared-memory bandwidth bound: broadcast W1 loads, A-tile stores, SS-mode operand
// re-reads): here BOTH layers run on the tensor cores and every A operand lives in TENSOR MEMORY.
//
//   layer 1   x (128 envs x 16, K padded)  --tcgen05.st-->  TMEM A1 (hi | lo)
//             D1[128 x 112] = A1 . W1^T   one half of the 200 hidden units at a time (6 MMAs per half)
//   convert   thread = row: tcgen05.ld 8 columns of D1, + b1, ReLU, split hi/lo, tcgen05.st into a ring of
//             TMEM A2 slots (one slot = one K-step of layer 2)
//   layer 2   D2[128 x 112] += A2[slot] . W2^T   (3 MMAs per K-step, A from TMEM, B from shared memory)
//   epilogue  tcgen05.ld D2, + b2, ReLU, the 100 x {5,3} layer and the arg-max in registers
//
// TMEM map (512 columns): D2 x2 [0,256) | D1 [256,368) | A1 hi,lo [368,400) | A2 ring 7 x (hi 8 | lo 8) [400,512)
// Shared memory: W2 hi+lo 179.2 KB and W1 hi+lo 28.7 KB as UMMA B operands (canonical K-major core-matrix
// layout, prepared on the host), b1 / b2 / W3 / b3 3 KB.  Warps: 4*NPAR converters (row quarter = warp % 4,
// K-step residue = warp / 4), then 4 epilogue warps, then the MMA-issue warp.  All mbarrier waits are bounded (trap, no hang).
#include <cstring>

#include "../merging_gym_b200/csrc/abi_common.h"

namespace mgtc2 {

constexpr int H1 = 200, H2 = 100;
constexpr int TM = 128, UN = 112;
constexpr int KSTEPS = H1 / 8;                // 25 layer-2 K-steps
constexpr int HALF_KS = UN / 8;               // 14 K-steps (112 hidden units) come out of one layer-1 half
constexpr int RING = 7;
constexpr int B_STEP = (UN / 8) * 256;        // 3584 B: one K-step (8 k) of a 112-row B operand
constexpr int B2_BYTES = KSTEPS * B_STEP;     // 89 600 B per hi / lo
constexpr int B1_BYTES = 2 * 2 * B_STEP;      // 2 halves x 2 K-steps (K = 16) per hi / lo = 14 336 B
constexpr uint32_t COL_D2 = 0, COL_D1 = 256, COL_A1 = 368, COL_A2 = 400;
constexpr int TMEM_COLS = 512;
#ifndef MG_TC2_NPAR
#define MG_TC2_NPAR 4
#endif
constexpr int NPAR = MG_TC2_NPAR;            // converter warps per row quarter (K-steps interleaved over them)
constexpr int EPI_WARP0 = 4 * NPAR, MMA_WARP = EPI_WARP0 + 4;
constexpr int NUM_THREADS = (MMA_WARP + 1) * 32;
constexpr uint32_t kSpinLimit = 1u << 26;

template <int OUT>
struct Smem {
    unsigned char b2_hi[B2_BYTES], b2_lo[B2_BYTES];
    unsigned char b1_hi[B1_BYTES], b1_lo[B1_BYTES];
    float w3[OUT][H2];
    float b1[H1 + 24], b2[H2 + 12], b3[8];
    unsigned long long full[RING], empty[RING];
    unsigned long long a1_full, a1_empty, d1_full, d1_drained, d2_full[2], d2_empty[2];
    uint32_t tmem_base;
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {      // no swizzle, K-major, LBO 128, SBO 256
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(128u >> 4) << 16) | ((uint64_t)(256u >> 4) << 32) |
           ((uint64_t)1 << 46);
}
constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(UN >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);

__device__ __forceinline__ void mbar_init(unsigned long long *b, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(unsigned long long *b) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *b, uint32_t parity) {
    uint32_t done = 0;
    for (uint32_t it = 0; it < kSpinLimit && !done; ++it) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
            : "=r"(done)
            : "r"(smem_u32(b)), "r"(parity)
            : "memory");
    }
    if (!done) __trap();
}
// D[tmem] (+)= A[tmem] . B[smem]^T, tf32 operands, fp32 accumulate
__device__ __forceinline__ void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tmem_d),
        "r"(tmem_a), "l"(db), "r"(kIdesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(unsigned long long *b) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t addr, const uint32_t (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(addr), "r"(v[0]), "r"(v[1]),
                 "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                 : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t addr, uint32_t (&v)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(addr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void split_tf32(float v, uint32_t &hi, uint32_t &lo) {
    hi = __float_as_uint(v) & 0xFFFFE000u;                     // what kind::tf32 reads
    lo = __float_as_uint(v - __uint_as_float(hi));             // exact remainder
}

template <int IN>
__device__ __forceinline__ void load_row(const float *__restrict__ obs, const uint8_t *__restrict__ goal, int64_t e,
                                         int64_t n, float (&x)[16]) {
    constexpr int off = IN - MG_OBS_DIM;
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = 0.f;
    if (e < n) {
        if (off) x[0] = (float)goal[e];
        const float2 *src = reinterpret_cast<const float2 *>(obs + e * MG_OBS_DIM);
#pragma unroll
        for (int i = 0; i < MG_OBS_DIM / 2; ++i) {
            const float2 v = __ldg(src + i);
            x[off + 2 * i] = v.x; x[off + 2 * i + 1] = v.y;
        }
    }
}

template <int IN, int OUT>
__global__ void __launch_bounds__(NUM_THREADS, 1)
mlp_act_tc2_kernel(const float *__restrict__ obs, const uint8_t *__restrict__ goal, const int64_t n,
                   const float *__restrict__ w_tc /* [b2_hi | b2_lo | b1_hi | b1_lo] canonical */,
                   const float *__restrict__ b1, const float *__restrict__ b2, const float *__restrict__ w3,
                   const float *__restrict__ b3, uint8_t *__restrict__ act, float *__restrict__ q_out) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    Smem<OUT> &S = *reinterpret_cast<Smem<OUT> *>(smem_raw);
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const int64_t n_tiles = (n + TM - 1) / TM;

    // ---- one-time setup ----------------------------------------------------------------------------
    {
        const float4 *src = reinterpret_cast<const float4 *>(w_tc);
        float4 *dst = reinterpret_cast<float4 *>(S.b2_hi);
        for (int i = t; i < (2 * B2_BYTES + 2 * B1_BYTES) / 16; i += NUM_THREADS) dst[i] = __ldg(src + i);
        for (int i = t; i < OUT * H2; i += NUM_THREADS) (&S.w3[0][0])[i] = w3[i];
        for (int i = t; i < H1 + 24; i += NUM_THREADS) S.b1[i] = i < H1 ? b1[i] : 0.f;
        for (int i = t; i < H2 + 12; i += NUM_THREADS) S.b2[i] = i < H2 ? b2[i] : 0.f;
        if (t < 8) S.b3[t] = t < OUT ? b3[t] : 0.f;
    }
    if (t == 0) {
        for (int s = 0; s < RING; ++s) { mbar_init(&S.full[s], TM); mbar_init(&S.empty[s], 1); }
        mbar_init(&S.a1_full, TM); mbar_init(&S.a1_empty, 1);
        mbar_init(&S.d1_full, 1); mbar_init(&S.d1_drained, NPAR * TM);
        for (int b = 0; b < 2; ++b) { mbar_init(&S.d2_full[b], 1); mbar_init(&S.d2_empty[b], TM); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == MMA_WARP) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&S.tmem_base)),
                     "r"((uint32_t)TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = S.tmem_base;

    if (warp < EPI_WARP0) {
        // =============================== CONVERTERS (thread = row of the tile) =====================
        const int q4 = warp & 3, par = warp >> 2;              // TMEM lane quarter, K-step parity
        const int row = q4 * 32 + lane;
        const uint32_t lane_base = tmem_base + ((uint32_t)(q4 * 32) << 16);
        float x[16];
        if (par == 0) load_row<IN>(obs, goal, (int64_t)blockIdx.x * TM + row, n, x);
        uint32_t tl = 0;
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tl) {
            if (par == 0) {
                // ---- this tile's input row -> TMEM A1 (hi | lo), then prefetch the next row ----------
                uint32_t hi[16], lo[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) split_tf32(x[i], hi[i], lo[i]);
                mbar_wait(&S.a1_empty, (tl & 1u) ^ 1u);        // layer-1 MMAs of the previous tile are done
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                tmem_st8(lane_base + COL_A1, reinterpret_cast<const uint32_t(&)[8]>(hi[0]));
                tmem_st8(lane_base + COL_A1 + 8, reinterpret_cast<const uint32_t(&)[8]>(hi[8]));
                tmem_st8(lane_base + COL_A1 + 16, reinterpret_cast<const uint32_t(&)[8]>(lo[0]));
                tmem_st8(lane_base + COL_A1 + 24, reinterpret_cast<const uint32_t(&)[8]>(lo[8]));
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                mbar_arrive(&S.a1_full);
                load_row<IN>(obs, goal, (tile + gridDim.x) * TM + row, n, x);
            }
            for (int h = 0; h < 2; ++h) {
                const uint32_t g = 2u * tl + (uint32_t)h;      // global layer-1 half counter
                mbar_wait(&S.d1_full, g & 1u);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const int ks0 = h * HALF_KS, ks1 = h ? KSTEPS : HALF_KS;
                for (int ks = ks0 + ((par - ks0) % NPAR + NPAR) % NPAR; ks < ks1; ks += NPAR) {   // K-steps with ks % NPAR == par
                    uint32_t v[8], hi[8], lo[8];
                    tmem_ld8(lane_base + COL_D1 + (uint32_t)((ks - ks0) * 8), v);
                    const float4 ba = *reinterpret_cast<const float4 *>(&S.b1[8 * ks]);
                    const float4 bb = *reinterpret_cast<const float4 *>(&S.b1[8 * ks + 4]);
                    const float bias[8] = {ba.x, ba.y, ba.z, ba.w, bb.x, bb.y, bb.z, bb.w};
#pragma unroll
                    for (int j = 0; j < 8; ++j) split_tf32(fmaxf(__uint_as_float(v[j]) + bias[j], 0.f), hi[j], lo[j]);
                    const uint32_t G = tl * KSTEPS + (uint32_t)ks;            // global K-step counter
                    const uint32_t slot = G % RING, use = G / RING;
                    mbar_wait(&S.empty[slot], (use & 1u) ^ 1u);               // MMAs that read this slot are done
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    tmem_st8(lane_base + COL_A2 + slot * 16, hi);
                    tmem_st8(lane_base + COL_A2 + slot * 16 + 8, lo);
                    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    mbar_arrive(&S.full[slot]);
                }
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                mbar_arrive(&S.d1_drained);                    // done reading this half of D1
            }
        }
    } else if (warp == MMA_WARP) {
        // =============================== MMA ISSUER ================================================
        if (lane == 0) {
            const uint32_t b2_hi = smem_u32(S.b2_hi), b2_lo = smem_u32(S.b2_lo);
            const uint32_t b1_hi = smem_u32(S.b1_hi), b1_lo = smem_u32(S.b1_lo);
            uint32_t tl = 0;
            for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tl) {
                const uint32_t buf = tl & 1u;
                const uint32_t d2 = tmem_base + COL_D2 + buf * 128u, d1 = tmem_base + COL_D1;
                const uint32_t a1_hi = tmem_base + COL_A1, a1_lo = tmem_base + COL_A1 + 16;
                mbar_wait(&S.d2_empty[buf], ((tl >> 1) & 1u) ^ 1u);
                mbar_wait(&S.a1_full, tl & 1u);
                for (int h = 0; h < 2; ++h) {
                    const uint32_t g = 2u * tl + (uint32_t)h;
                    if (g > 0) mbar_wait(&S.d1_drained, (g - 1u) & 1u);       // converters finished the previous half
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    // ---- layer 1, this half: D1 = x . W1[half]^T over K = 16 (two K-steps), 3xTF32 ----
#pragma unroll
                    for (int k1 = 0; k1 < 2; ++k1) {
                        const uint64_t bh = make_desc(b1_hi + (h * 2 + k1) * B_STEP), bl = make_desc(b1_lo + (h * 2 + k1) * B_STEP);
                        umma_ts(d1, a1_hi + k1 * 8, bh, k1 > 0 ? 1u : 0u);
                        umma_ts(d1, a1_lo + k1 * 8, bh, 1u);
                        umma_ts(d1, a1_hi + k1 * 8, bl, 1u);
                    }
                    umma_commit(&S.d1_full);
                    if (h == 1) umma_commit(&S.a1_empty);      // A1 may be overwritten once these MMAs are done
                    // ---- layer 2 K-steps fed by this half ----------------------------------------------
                    const int ks0 = h * HALF_KS, ks1 = h ? KSTEPS : HALF_KS;
                    for (int ks = ks0; ks < ks1; ++ks) {
                        const uint32_t G = tl * KSTEPS + (uint32_t)ks;
                        const uint32_t slot = G % RING, use = G / RING;
                        mbar_wait(&S.full[slot], use & 1u);
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        const uint32_t a_hi = tmem_base + COL_A2 + slot * 16, a_lo = a_hi + 8;
                        const uint64_t bh = make_desc(b2_hi + ks * B_STEP), bl = make_desc(b2_lo + ks * B_STEP);
                        umma_ts(d2, a_hi, bh, ks > 0 ? 1u : 0u);
                        umma_ts(d2, a_lo, bh, 1u);
                        umma_ts(d2, a_hi, bl, 1u);
                        umma_commit(&S.empty[slot]);
                    }
                }
                umma_commit(&S.d2_full[buf]);
            }
        }
        __syncwarp();
    } else {
        // =============================== EPILOGUE: layer 3 + arg-max ==============================
        const int q4 = warp - EPI_WARP0;
        const int m = q4 * 32 + lane;
        uint32_t tl = 0;
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tl) {
            const uint32_t buf = tl & 1u;
            mbar_wait(&S.d2_full[buf], (tl >> 1) & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t taddr = tmem_base + COL_D2 + buf * 128u + ((uint32_t)(q4 * 32) << 16);
            float q[OUT];
#pragma unroll
            for (int o = 0; o < OUT; ++o) q[o] = S.b3[o];
#pragma unroll
            for (int c0 = 0; c0 < UN; c0 += 16) {
                uint32_t v[16];
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                    : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                      "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                    : "r"(taddr + (uint32_t)c0));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int j4 = 0; j4 < 16; j4 += 4) {
                    if (c0 + j4 < H2) {
                        const float4 bias = *reinterpret_cast<const float4 *>(&S.b2[c0 + j4]);
                        const float h0 = fmaxf(__uint_as_float(v[j4]) + bias.x, 0.f);
                        const float h1 = fmaxf(__uint_as_float(v[j4 + 1]) + bias.y, 0.f);
                        const float h2 = fmaxf(__uint_as_float(v[j4 + 2]) + bias.z, 0.f);
                        const float h3 = fmaxf(__uint_as_float(v[j4 + 3]) + bias.w, 0.f);
#pragma unroll
                        for (int o = 0; o < OUT; ++o) {
                            const float4 w = *reinterpret_cast<const float4 *>(&S.w3[o][c0 + j4]);
                            q[o] = fmaf(h0, w.x, q[o]); q[o] = fmaf(h1, w.y, q[o]);
                            q[o] = fmaf(h2, w.z, q[o]); q[o] = fmaf(h3, w.w, q[o]);
                        }
                    }
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            mbar_arrive(&S.d2_empty[buf]);
            const int64_t e = tile * TM + m;
            if (e < n) {
                int best = 0;
                float bv = q[0];
#pragma unroll
                for (int o = 1; o < OUT; ++o)
                    if (q[o] > bv) { bv = q[o]; best = o; }
                act[e] = (uint8_t)best;
                if (q_out) {
#pragma unroll
                    for (int o = 0; o < OUT; ++o) q_out[e * OUT + o] = q[o];
                }
            }
        }
    }
    // ---- teardown -------------------------------------------------------------------------------------
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == MMA_WARP)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS));
}

template <int IN, int OUT>
cudaError_t launch(const float *obs, const uint8_t *goal, int64_t n, const float *w_tc, const float *b1, const float *b2,
                   const float *w3, const float *b3, uint8_t *act, float *q_out, cudaStream_t st) {
    auto kern = mlp_act_tc2_kernel<IN, OUT>;
    const size_t smem = sizeof(Smem<OUT>);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e) return e;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int64_t tiles = (n + TM - 1) / TM;
    const unsigned grid = (unsigned)(tiles < sms ? tiles : sms);
    kern<<<grid, NUM_THREADS, smem, st>>>(obs, goal, n, w_tc, b1, b2, w3, b3, act, q_out);
    return cudaGetLastError();
}

}  // namespace mgtc2

extern "C" MG_API int mg_mlp_act_tc2(const float *obs, const uint8_t *goal_or_null, int64_t n, int32_t obs_dim,
                                     int32_t out_dim, const float *w_tc, const float *b1, const float *b2,
                                     const float *w3, const float *b3, uint8_t *actions, float *q_out_or_null,
                                     void *stream) {
    using namespace mg_abi;
    if (n < 0) return fail(MG_ERR_BAD_SIZE, "n < 0");
    const int in_dim = obs_dim + (goal_or_null ? 1 : 0);
    if (obs_dim != MG_OBS_DIM || !(out_dim == 5 || out_dim == 3))
        return fail(MG_ERR_BAD_SIZE, "mg_mlp_act_tc2 supports obs rows of 10 floats (+ optional goal) and 5 or 3 outputs");
    if (n == 0) return MG_OK;
    if (!obs || !w_tc || !b1 || !b2 || !w3 || !b3 || !actions) return fail(MG_ERR_NULL_POINTER, "mg_mlp_act_tc2: NULL pointer");
    if (!aligned16(obs) || !aligned16(w_tc)) return fail(MG_ERR_ALIGNMENT, "obs and w_tc must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e;
#define MG_TC2_CASE(I, O) \
    if (in_dim == I && out_dim == O) e = mgtc2::launch<I, O>(obs, goal_or_null, n, w_tc, b1, b2, w3, b3, actions, q_out_or_null, st); else
    MG_TC2_CASE(10, 5) MG_TC2_CASE(10, 3) MG_TC2_CASE(11, 5) MG_TC2_CASE(11, 3) e = cudaErrorInvalidValue;
#undef MG_TC2_CASE
    if (e) return cuda_fail(e, "mg_mlp_act_tc2 launch");
    return MG_OK;
}
