// mlp_tc_kernels.cu — tensor-core version of the Q-network forward + arg-max (sm_100a: tcgen05 + TMEM).
//
// Same operator as mlp_kernels.cu (`Net(in, out)`: Linear(in,200)-ReLU-Linear(200,100)-ReLU-
// Linear(100,out) + arg-max; scripts/main.py:30-47, hdqn.py:38-55), but the 200x100 layer — 89 % of the
// FLOPs — runs on the 5th-generation tensor cores as a "3xTF32" product that keeps fp32-level
// accuracy:   a*b ~= a_hi*b_hi + a_lo*b_hi + a_hi*b_lo,   a_hi = tf32(a) (top 19 bits), a_lo = a - a_hi
// (the tensor core reads the top 19 bits of an fp32 operand, so a_lo loses only bits below 2^-22 |a|).
// Opt-in (`MLPPolicy(backend="tf32x3")`); the FFMA kernel stays the default because it evaluates
// exactly the reference's fp32 arithmetic.
//
// One persistent CTA per SM, tile = 128 envs (UMMA M = 128, N = 112 = 100 neurons + zero pad, K = 8):
//   warps 0-7  producers : two groups of 4 warps, each owning one slot of a 2-slot ring (slot = 2
//                          K-steps); thread = (env pair, K-step): layer 1 on the CUDA cores (FFMA2), split
//                          into hi/lo, written straight into the canonical K-major core-matrix layout -> full[s]
//   warp  12   MMA issue : one thread; per K-step two tcgen05.mma.kind::tf32: a_hi x [W2_hi ; W2_lo] (N = 224)
//                          and a_lo x W2_hi (N = 112), accumulating in TMEM; tcgen05.commit -> empty[s] / tmem_full[b]
//   warps 8-11 epilogue  : tcgen05.ld of the 128x112 fp32 accumulator (row = env), bias + ReLU, the
//                          100x{5,3} layer and the arg-max in registers, one action byte per thread
// TMEM: 2 accumulator buffers x 256 columns ([0,112): hi.hi + lo.hi, [112,224): hi.lo, summed in the epilogue),
// so the epilogue of tile i overlaps the MMAs of tile i+1.
// Shared memory: W2 as one stacked B operand [hi ; lo] of 224 rows, 179.2 KB (prepared on the host), A ring 32 KB,
// W1 / W3 / biases 11 KB.  Every mbarrier wait is bounded and traps instead of hanging.
#include "abi_common.h"

namespace mgtc {

constexpr int H1 = 200, H2 = 100;
constexpr int TM = 128;                       // envs per tile = UMMA M
constexpr int UN = 112;                       // UMMA N (multiple of 16 for M = 128)
constexpr int KSTEPS = H1 / 8;                // 25 K-steps of 8 (tf32: 32 bytes of K per MMA)
constexpr int STAGES = 2;                     // ring stages; one stage = 2 K-steps (16 hidden units)
constexpr int NSTAGE_TILE = (KSTEPS + 1) / 2; // 13 stage fills per tile (the last holds one K-step)
constexpr int A_STEP = (TM / 8) * 256;        // 4096 B : 16 row groups x 2 core matrices x 128 B
constexpr int B_STEP = (2 * UN / 8) * 256;    // 7168 B: one K-step of the stacked B operand [W2_hi ; W2_lo] (224 rows)
constexpr int B_BYTES = KSTEPS * B_STEP;      // 179 200 B
constexpr int TMEM_COLS = 512;                // 2 accumulator buffers x 256 columns (224 used)
constexpr int NUM_PRODUCERS = 256;            // 8 producer warps: thread = (env, which K-step of the stage)
constexpr int NUM_THREADS = 416;              // 8 producer warps + 4 epilogue warps + 1 MMA warp
constexpr int MAX_OUT = 8;
constexpr uint32_t kSpinLimit = 1u << 26;

template <int IN, int OUT>
struct Smem {
    unsigned char b_cat[B_BYTES];             // [K-step][28 row groups: W2_hi rows 0-111, W2_lo rows 112-223], canonical layout
    unsigned char a_hi[STAGES][2][A_STEP];
    unsigned char a_lo[STAGES][2][A_STEP];
    float w1[IN][H1];
    float w3[OUT][H2];
    float b1[H1], b2[H2 + 12], b3[MAX_OUT];
    unsigned long long full[STAGES], empty[STAGES], tmem_full[2], tmem_empty[2];
    uint32_t tmem_base;
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// tcgen05 shared-memory matrix descriptor, no swizzle, K-major: core matrix = 8 rows x 16 bytes stored as
// 128 contiguous bytes; LBO = byte distance between the two core matrices of a K-step (128), SBO = byte
// distance between 8-row groups (256).  Verified numerically in profiles/exp_tcgen05.cu.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(128u >> 4) << 16) | ((uint64_t)(256u >> 4) << 32) |
           ((uint64_t)1 << 46);
}
// instruction descriptor: D = f32, A = B = tf32, both K-major, N >> 3 at bit 17, M >> 4 at bit 24
constexpr uint32_t idesc_n(int n) { return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(TM >> 4) << 24); }
constexpr uint32_t kIdesc112 = idesc_n(UN), kIdesc224 = idesc_n(2 * UN);

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
    if (!done) __trap();                      // never hang the GPU on a protocol bug
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
        "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(unsigned long long *b) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(b)) : "memory");
}

template <int IN>
__device__ __forceinline__ void load_row(const float *__restrict__ obs, const uint8_t *__restrict__ goal, int64_t e,
                                         int64_t n, int obs_dim, float (&x)[IN]) {
    constexpr int off = IN - MG_OBS_DIM;   // 1 when a goal column is prepended (hdqn.py:291); compile-time so x[] stays in registers
    (void)obs_dim;
    if (e < n) {
        if (off) x[0] = (float)goal[e];
        const float2 *src = reinterpret_cast<const float2 *>(obs + e * MG_OBS_DIM);
#pragma unroll
        for (int i = 0; i < (IN - (IN & 1)) / 2; ++i) {
            const float2 v = __ldg(src + i);
            x[off + 2 * i] = v.x; x[off + 2 * i + 1] = v.y;
        }
    } else {
#pragma unroll
        for (int i = 0; i < IN; ++i) x[i] = 0.f;
    }
}

template <int IN, int OUT>
__global__ void __launch_bounds__(NUM_THREADS, 1)
mlp_act_tc_kernel(const float *__restrict__ obs, const uint8_t *__restrict__ goal, const int64_t n, const int obs_dim,
                  const float *__restrict__ w1t, const float *__restrict__ b1, const float *__restrict__ w2_tc,
                  const float *__restrict__ b2, const float *__restrict__ w3, const float *__restrict__ b3,
                  uint8_t *__restrict__ act, float *__restrict__ q_out) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    Smem<IN, OUT> &S = *reinterpret_cast<Smem<IN, OUT> *>(smem_raw);
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const int64_t n_tiles = (n + TM - 1) / TM;

    // ---- one-time setup: weights -> smem, barriers, TMEM -------------------------------------------
    {
        const float4 *src = reinterpret_cast<const float4 *>(w2_tc);          // already in the canonical layout
        float4 *dst = reinterpret_cast<float4 *>(S.b_cat);
        for (int i = t; i < B_BYTES / 16; i += NUM_THREADS) dst[i] = __ldg(src + i);
        const float4 *s1 = reinterpret_cast<const float4 *>(w1t);
        float4 *d1 = reinterpret_cast<float4 *>(&S.w1[0][0]);
        for (int i = t; i < IN * H1 / 4; i += NUM_THREADS) d1[i] = __ldg(s1 + i);
        for (int i = t; i < OUT * H2; i += NUM_THREADS) (&S.w3[0][0])[i] = w3[i];
        for (int i = t; i < H1; i += NUM_THREADS) S.b1[i] = b1[i];
        for (int i = t; i < H2 + 12; i += NUM_THREADS) S.b2[i] = i < H2 ? b2[i] : 0.f;
        if (t < OUT) S.b3[t] = b3[t];
    }
    if (t == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&S.full[s], NUM_PRODUCERS / 2); mbar_init(&S.empty[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&S.tmem_full[b], 1); mbar_init(&S.tmem_empty[b], TM); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // W2 hi/lo written by the generic proxy
    if (warp == 12) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&S.tmem_base)),
                     "r"((uint32_t)TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = S.tmem_base;

    if (warp < 8) {
        // =================================== PRODUCERS: layer 1 ===================================
        // Two independent groups of 4 warps; group g owns ring slot g and fills every second stage, so the
        // write chain of one slot (wait empty -> STS -> proxy fence -> arrive) overlaps the MMAs of the other.
        // thread = (env pair {ep, ep+64}, K-step kh of the stage): 8 hidden units for 2 envs.
        const int grp = warp >> 2, tg = t & 127;
        const int ep = tg & 63, kh = tg >> 6;
        const int m0 = ep, m1 = ep + 64;
        const uint32_t off0 = (uint32_t)((m0 >> 3) * 256 + (m0 & 7) * 16);
        const uint32_t off1 = (uint32_t)((m1 >> 3) * 256 + (m1 & 7) * 16);
        float x0[IN], x1[IN];
        load_row<IN>(obs, goal, (int64_t)blockIdx.x * TM + m0, n, obs_dim, x0);
        load_row<IN>(obs, goal, (int64_t)blockIdx.x * TM + m1, n, obs_dim, x1);
        uint32_t tl = 0;
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tl) {
            float n0[IN], n1[IN];                               // next tile's rows, in flight during this tile
            load_row<IN>(obs, goal, (tile + gridDim.x) * TM + m0, n, obs_dim, n0);
            load_row<IN>(obs, goal, (tile + gridDim.x) * TM + m1, n, obs_dim, n1);
            const uint32_t it0 = tl * NSTAGE_TILE;              // global stage-fill counter of this tile's stage 0
            for (int st = (int)((grp + it0) & 1u); st < NSTAGE_TILE; st += 2) {     // stages with (it0 + st) % 2 == grp
                const uint32_t it = it0 + (uint32_t)st;
                const int s = it % STAGES;                      // == grp
                const uint32_t ph = (it / STAGES) & 1u;
                const int ks = 2 * st + kh;
                float4 ha0, ha1, hb0, hb1;                      // env m0: units 0-3, 4-7; env m1: units 0-3, 4-7
                if (ks < KSTEPS) {
                    const int k = 8 * ks;
                    const float4 ba = *reinterpret_cast<const float4 *>(&S.b1[k]);
                    const float4 bb = *reinterpret_cast<const float4 *>(&S.b1[k + 4]);
                    float2 a01 = make_float2(ba.x, ba.y), a23 = make_float2(ba.z, ba.w), a45 = make_float2(bb.x, bb.y), a67 = make_float2(bb.z, bb.w);
                    float2 c01 = a01, c23 = a23, c45 = a45, c67 = a67;
#pragma unroll
                    for (int i = 0; i < IN; ++i) {
                        const float4 wa = *reinterpret_cast<const float4 *>(&S.w1[i][k]);
                        const float4 wb = *reinterpret_cast<const float4 *>(&S.w1[i][k + 4]);
                        const float2 w01 = make_float2(wa.x, wa.y), w23 = make_float2(wa.z, wa.w);
                        const float2 w45 = make_float2(wb.x, wb.y), w67 = make_float2(wb.z, wb.w);
                        const float2 xa = make_float2(x0[i], x0[i]), xb = make_float2(x1[i], x1[i]);
                        a01 = __ffma2_rn(xa, w01, a01); a23 = __ffma2_rn(xa, w23, a23);
                        a45 = __ffma2_rn(xa, w45, a45); a67 = __ffma2_rn(xa, w67, a67);
                        c01 = __ffma2_rn(xb, w01, c01); c23 = __ffma2_rn(xb, w23, c23);
                        c45 = __ffma2_rn(xb, w45, c45); c67 = __ffma2_rn(xb, w67, c67);
                    }
                    ha0 = make_float4(fmaxf(a01.x, 0.f), fmaxf(a01.y, 0.f), fmaxf(a23.x, 0.f), fmaxf(a23.y, 0.f));
                    ha1 = make_float4(fmaxf(a45.x, 0.f), fmaxf(a45.y, 0.f), fmaxf(a67.x, 0.f), fmaxf(a67.y, 0.f));
                    hb0 = make_float4(fmaxf(c01.x, 0.f), fmaxf(c01.y, 0.f), fmaxf(c23.x, 0.f), fmaxf(c23.y, 0.f));
                    hb1 = make_float4(fmaxf(c45.x, 0.f), fmaxf(c45.y, 0.f), fmaxf(c67.x, 0.f), fmaxf(c67.y, 0.f));
                }
                mbar_wait(&S.empty[s], ph ^ 1u);                // MMAs that read this slot have completed
                if (ks < KSTEPS) {
                    auto split_store = [&](const float4 &h, uint32_t off) {
                        float4 hi, lo;                          // hi = what kind::tf32 reads (top 19 bits), lo exact
                        hi.x = __uint_as_float(__float_as_uint(h.x) & 0xFFFFE000u); lo.x = h.x - hi.x;
                        hi.y = __uint_as_float(__float_as_uint(h.y) & 0xFFFFE000u); lo.y = h.y - hi.y;
                        hi.z = __uint_as_float(__float_as_uint(h.z) & 0xFFFFE000u); lo.z = h.z - hi.z;
                        hi.w = __uint_as_float(__float_as_uint(h.w) & 0xFFFFE000u); lo.w = h.w - hi.w;
                        *reinterpret_cast<float4 *>(S.a_hi[s][kh] + off) = hi;
                        *reinterpret_cast<float4 *>(S.a_lo[s][kh] + off) = lo;
                    };
                    split_store(ha0, off0); split_store(ha1, off0 + 128);
                    split_store(hb0, off1); split_store(hb1, off1 + 128);
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // visible to the tensor core
                }
                mbar_arrive(&S.full[s]);                        // (a per-warp elected arrive measured slower)
            }
#pragma unroll
            for (int i = 0; i < IN; ++i) { x0[i] = n0[i]; x1[i] = n1[i]; }
        }
    } else if (warp == 12) {
        // =================================== MMA ISSUER ==========================================
        if (lane == 0) {
            uint32_t it = 0, tl = 0;
            const uint32_t b_cat = smem_u32(S.b_cat);
            for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tl) {
                const uint32_t buf = tl & 1u;
                mbar_wait(&S.tmem_empty[buf], ((tl >> 1) & 1u) ^ 1u);            // epilogue drained this buffer
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t d = tmem_base + buf * 256u;
                for (int st = 0; st < NSTAGE_TILE; ++st, ++it) {
                    const int s = it % STAGES;
                    mbar_wait(&S.full[s], (it / STAGES) & 1u);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
                    for (int half = 0; half < 2; ++half) {
                        const int ks = 2 * st + half;
                        if (ks < KSTEPS) {
                            const uint64_t ahi = make_desc(smem_u32(S.a_hi[s][half])), alo = make_desc(smem_u32(S.a_lo[s][half]));
                            const uint64_t bcat = make_desc(b_cat + ks * B_STEP);
                            // columns [0,112) += a_hi.W2_hi, columns [112,224) += a_hi.W2_lo  (one N = 224 MMA: 123 cycles
                            // instead of two N = 112 MMAs at 76 each, and A_hi is read once)
                            umma_tf32(d, ahi, bcat, kIdesc224, ks > 0 ? 1u : 0u);
                            umma_tf32(d, alo, bcat, kIdesc112, 1u);           // columns [0,112) += a_lo.W2_hi
                        }
                    }
                    umma_commit(&S.empty[s]);                   // stage reusable once these MMAs are done
                }
                umma_commit(&S.tmem_full[buf]);                 // accumulator complete
            }
        }
        __syncwarp();
    } else {
        // =================================== EPILOGUE: layer 3 + arg-max ===========================
        const int q4 = warp - 8;                                // TMEM lane quarter of this warp (warp % 4)
        const int m = q4 * 32 + lane;
        uint32_t tl = 0;
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tl) {
            const uint32_t buf = tl & 1u;
            mbar_wait(&S.tmem_full[buf], (tl >> 1) & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t taddr = tmem_base + buf * 256u + ((uint32_t)(q4 * 32) << 16);
            float q[OUT];
#pragma unroll
            for (int o = 0; o < OUT; ++o) q[o] = S.b3[o];
#pragma unroll
            for (int c0 = 0; c0 < UN; c0 += 16) {
                uint32_t v[16], u[16];
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                    : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                      "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                    : "r"(taddr + (uint32_t)c0));
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                    : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]),
                      "=r"(u[8]), "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15])
                    : "r"(taddr + (uint32_t)(UN + c0)));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(u[j]));
#pragma unroll
                for (int j4 = 0; j4 < 16; j4 += 4) {
                    if (c0 + j4 < H2) {                         // 100 = 25 groups of 4: no partial group
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
            mbar_arrive(&S.tmem_empty[buf]);                    // this thread is done reading the buffer
            const int64_t e = tile * TM + m;
            if (e < n) {
                int best = 0;
                float bv = q[0];
#pragma unroll
                for (int o = 1; o < OUT; ++o)
                    if (q[o] > bv) { bv = q[o]; best = o; }     // first maximum, like torch.max
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
    if (warp == 12)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS));
}

template <int IN, int OUT>
cudaError_t launch(const float *obs, const uint8_t *goal, int64_t n, int obs_dim, const float *w1t, const float *b1,
                   const float *w2_tc, const float *b2, const float *w3, const float *b3, uint8_t *act, float *q_out,
                   cudaStream_t st) {
    auto kern = mlp_act_tc_kernel<IN, OUT>;
    const size_t smem = sizeof(Smem<IN, OUT>) + 1024;           // slack for the 1024-byte alignment of the base
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e) return e;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int64_t tiles = (n + TM - 1) / TM;
    const unsigned grid = (unsigned)(tiles < sms ? tiles : sms);
    kern<<<grid, NUM_THREADS, smem, st>>>(obs, goal, n, obs_dim, w1t, b1, w2_tc, b2, w3, b3, act, q_out);
    return cudaGetLastError();
}

}  // namespace mgtc

extern "C" MG_API int mg_mlp_act_tc(const float *obs, const uint8_t *goal_or_null, int64_t n, int32_t obs_dim,
                                    int32_t out_dim, const float *w1t, const float *b1, const float *w2_tc,
                                    const float *b2, const float *w3, const float *b3, uint8_t *actions,
                                    float *q_out_or_null, void *stream) {
    using namespace mg_abi;
    if (n < 0) return fail(MG_ERR_BAD_SIZE, "n < 0");
    const int in_dim = obs_dim + (goal_or_null ? 1 : 0);
    if (obs_dim != MG_OBS_DIM || !(out_dim == 5 || out_dim == 3))
        return fail(MG_ERR_BAD_SIZE, "mg_mlp_act_tc supports obs rows of 10 floats (+ optional goal) and 5 or 3 outputs");
    if (n == 0) return MG_OK;
    if (!obs || !w1t || !b1 || !w2_tc || !b2 || !w3 || !b3 || !actions)
        return fail(MG_ERR_NULL_POINTER, "mg_mlp_act_tc: NULL pointer");
    if (!aligned16(obs) || !aligned16(w1t) || !aligned16(w2_tc))
        return fail(MG_ERR_ALIGNMENT, "obs and weight arrays must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e;
#define MG_TC_CASE(I, O) \
    if (in_dim == I && out_dim == O) e = mgtc::launch<I, O>(obs, goal_or_null, n, obs_dim, w1t, b1, w2_tc, b2, w3, b3, actions, q_out_or_null, st); else
    MG_TC_CASE(10, 5) MG_TC_CASE(10, 3) MG_TC_CASE(11, 5) MG_TC_CASE(11, 3) e = cudaErrorInvalidValue;
#undef MG_TC_CASE
    if (e) return cuda_fail(e, "mg_mlp_act_tc launch");
    return MG_OK;
}
