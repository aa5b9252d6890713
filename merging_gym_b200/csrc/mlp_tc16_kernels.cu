// mlp_tc16_kernels.cu — "f16x3": the Q-network forward + arg-max with BOTH hidden layers on the tcgen05 tensor cores.
//
// Same operator as mlp_kernels.cu / mlp_tc_kernels.cu (`Net(in, out)`: Linear(in,200)-ReLU-Linear(200,100)-ReLU-
// Linear(100,out) + arg-max; scripts/main.py:30-47, hdqn.py:38-55).  The 3xTF32 kernel (mlp_tc_kernels.cu) computes layer 1
// on the CUDA cores and is bound by what its eight producer warps have to ISSUE: with their layer-1 FFMA2s compiled out it
// drops from 67 to 54 us per 2^18 envs, and halving its tensor-core operand bytes (fp16 operands, same pipeline) bought
// only 3 % (profiles/r02_mlp_tc_f16_operands_experiment.patch).  Here layer 1 is a tensor-core product as well, so the
// "producers" shrink to converters that read 16 hidden units of an env row from tensor memory, apply ReLU, split and store
// them as the next MMA's A operand: ~100 instructions per (32 envs x 16 units) instead of ~540 per (128 envs x 8 units).
//
// Arithmetic: every product is the error-compensated three-term sum   a*w ~= a_hi*w_hi + a_hi*w_lo + a_lo*w_hi   on
// kind::f16 operands (hi = fp16(v), lo = fp16(v - hi): 2 x 11 significant bits, what 3xTF32 keeps too), accumulated in
// fp32 in ONE tensor-memory accumulator per layer.  fp16 has little range, so the host scales W1 (with b1 in K slot 15,
// fed by a constant 1.0 in the observation operand: the bias rides in the MMA) and W2 by powers of two that put their
// largest entries near 2^9 — the lo parts of all but negligible weights then stay normal numbers — and the kernel
// multiplies back: a' = relu(acc1 * c1) = h1 / 8 (kept below fp16's 65504: activations saturate at 5.2e5),
// h2 = relu(acc2 * c2 + b2).  Layer 3 and the arg-max run in fp32 on the CUDA cores as in the other kernels.
//
// One persistent CTA per SM, tile = 128 envs (UMMA M = 128), 14 warps:
//   warp 13     X loader : the tile's observation rows -> hi / lo fp16 operand [128 x 16] (K slot 15 = 1.0), double-buffered
//   warp 12     MMA issue: per tile 2 x 3 MMAs N = 112 / 96, K = 16 (layer 1 in two halves, into TMEM columns [256, 464)) and
//                          13 K-steps x 3 MMAs N = 112 (layer 2, accumulator buffer tile & 1); the first MMA of an accumulator
//                          overwrites it, so nobody zeroes tensor memory.  The halves of the NEXT tile's layer 1 are issued
//                          inside this tile's K loop, as soon as the converters have read this tile's.
//   warps 0-7   converters: warp w owns TMEM lane quarter w % 4 and every second K-step (w / 4): tcgen05.ld 16 columns,
//                          scale + ReLU + split, four 16-byte stores into ring slot (K-step % 8), arrive on full[slot]
//   warps 8-11  epilogue : tcgen05.ld the 128 x 112 fp32 accumulator, scale + bias + ReLU, the 100 x {5,3} layer, arg-max
//   (mg_policy_step: + 4 env warps that step the finished tiles' envs, see ENV below)
// Every mbarrier wait is bounded and traps instead of hanging.  Accumulation order is the issue order of one thread: the
// kernel is bitwise reproducible.  Measured: 43 us per 2^18 envs inside a CUDA graph with PDL (3xTF32: 64), a tile takes
// ~5 500 cycles against 5 355 for its 45 operand-fetch-bound MMAs (profiles/exp_tc16_trace.cu).
#include <cuda_fp16.h>

#include "abi_common.h"
#include "policy_env.cuh"
#include "tc_common.cuh"

namespace mgtc16 {
using mgtc::kDescHi;
using mgtc::kSpinLimit;
using mgtc::load_row;
using mgtc::make_desc;
using mgtc::mbar_arrive;
using mgtc::mbar_init;
using mgtc::mbar_test;
using mgtc::mbar_wait;
using mgtc::smem_u32;
using mgtc::TM;

constexpr int H1P = 208, H2 = 100;            // hidden widths: 200 (padded to whole 16-unit K-steps) and 100
constexpr int UN = 112;                       // layer-2 UMMA N (100 neurons + zero pad; multiple of 16 for M = 128)
constexpr int H2P = 104;                      // layer-3 weights padded to whole 8-column blocks
constexpr int KSTEPS = H1P / 16;              // 13 layer-2 K-steps of 16 hidden units
#ifndef MG_TC16_STAGES
#define MG_TC16_STAGES 8
#endif
constexpr int STAGES = MG_TC16_STAGES;        // A-operand ring slots, one K-step each
constexpr int A_STEP = (TM / 8) * 256;        // 4096 B: 16 row groups x 2 K halves x (8 rows x 16 B)
constexpr int W2_STEP = (2 * UN / 8) * 256;   // 7168 B: one K-step of [W2_hi (14 row groups) ; W2_lo (14 row groups)]
constexpr int W2_BYTES = KSTEPS * W2_STEP;    // 93 184 B
constexpr int W1_BYTES = (2 * H1P / 8) * 256; // 13 312 B: [W1_hi (26 row groups) ; W1_lo (26 row groups)], K = 16
constexpr int HDR_BYTES = 64;                 // float c1, c2 and padding in front of the operands
constexpr int W_BYTES = W1_BYTES + W2_BYTES;  // 106 496 B, copied as 8 pieces of W1_BYTES
static_assert(W2_BYTES == 7 * W1_BYTES, "the weight blob is copied in eight equal pieces");
constexpr int L2_COL0 = 0, L2_COL1 = 128, L1_COL = 256, TMEM_COLS = 512;
#ifndef MG_TC16_CONV_GROUPS
#define MG_TC16_CONV_GROUPS 2
#endif
constexpr int CONV_GROUPS = MG_TC16_CONV_GROUPS;                 // converter warps per TMEM lane quarter: K-steps interleaved over them
constexpr int CONV_WARPS = 4 * CONV_GROUPS, EPI_WARP0 = CONV_WARPS, MMA_WARP = CONV_WARPS + 4, X_WARP = CONV_WARPS + 5;
constexpr int NUM_THREADS = 32 * (CONV_WARPS + 6);
// mg_policy_step only: four more warps own the env step of the tiles the epilogue has finished (policy_env.cuh), one 32-env
// round of a tile each: 18 warps = 576 threads at 96 registers (52 bytes of spills).  Two env warps keep all 128 registers
// but take two rounds per tile in sequence: 10.0 instead of 9.2 us per step at 4096 envs, 24.2 instead of 21.1 at 65 536.
#ifndef MG_TC16_ENV_WARPS
#define MG_TC16_ENV_WARPS 4
#endif
constexpr int ENV_WARP0 = CONV_WARPS + 6, ENV_WARPS = MG_TC16_ENV_WARPS, ENV_BUFS = 4;
constexpr int NUM_THREADS_ENV = 32 * (ENV_WARP0 + ENV_WARPS);
#ifndef MG_TC16_TRACE
#define MG_TC16_TRACE 0                       // 1: CTA 0 records clock64() at the hand-over points (profiles/exp_tc16_trace.cu)
#endif
#if MG_TC16_TRACE
constexpr int TRACE_G = 13 * 10;              // K-steps traced (10 tiles of CTA 0)
__device__ long long g_trace_conv[TRACE_G][6];   // top, l1 ready, ld done, converted, slot free, arrived (lane-quarter 0 warps)
__device__ long long g_trace_mma[TRACE_G][4];    // top, full seen, issued, after l1 poll
__device__ long long g_trace_epi[16][3];         // wait start, wait end, done (quarter 0)
__device__ long long g_trace_l1[32][2];          // layer1(tile, half): start, issued
#define MG_TRACE16(arr, idx, k) do { if (blockIdx.x == 0 && lane == 0 && (idx) < (uint32_t)(sizeof(arr) / sizeof(arr[0]))) arr[idx][k] = clock64(); } while (0)
#else
#define MG_TRACE16(arr, idx, k) do { } while (0)
#endif
// Layer 1 is computed, consumed and released in two halves: hidden units [0,112) = K-steps 0-6 and [112,208) = K-steps 7-12.
// The first half of the NEXT tile is issued while the converters are still in the second half of this one, so they never
// wait at a tile boundary (with one 208-column accumulator they idled ~1500 cycles per tile: it could only be issued once
// they had read its last columns, and then ran behind the queued layer-2 MMAs).
#ifndef MG_TC16_KS_HALF
#define MG_TC16_KS_HALF 7                     // 13 = one 208-column accumulator (experiment switch)
#endif
constexpr int KS_HALF = MG_TC16_KS_HALF, L1_NA = 16 * KS_HALF, L1_NB = H1P - L1_NA;          // N = 112 and 96
constexpr int L1_POLL_FROM = L1_NB ? KS_HALF : 10;
static_assert(L1_NA % 16 == 0 && L1_NB % 16 == 0 && (L1_NB == 0 || CONV_GROUPS <= KSTEPS - KS_HALF), "every converter warp has a K-step in both halves");
constexpr int MAX_OUT = 8;

template <int OUT>
struct Smem {
    unsigned char w1[W1_BYTES];               // contiguous with w2: the blob behind its header
    unsigned char w2[W2_BYTES];
    unsigned char a_hi[STAGES][A_STEP];
    unsigned char a_lo[STAGES][A_STEP];
    unsigned char x_hi[2][A_STEP];
    unsigned char x_lo[2][A_STEP];
    float w3[OUT][H2P];
    float b2[H2 + 12], b3[MAX_OUT];
    float c1, c2;
    unsigned long long full[STAGES], empty[STAGES], x_full[2], x_empty[2], l1_full[2], l1_empty[2], tmem_full[2], tmem_empty[2], w_ready;
    uint32_t tmem_base;
    mgpe::Handoff<TM, ENV_BUFS> env;          // ENV: the tile's actions, epilogue warps -> env warps
};

// kind::f16 with fp16 operands (format fields 0), D = f32, both K-major, N >> 3 at bit 17, M >> 4 at bit 24
constexpr uint32_t idesc_n(int n) { return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(TM >> 4) << 24); }
constexpr uint32_t kIdesc112 = idesc_n(UN), kIdescA = idesc_n(L1_NA), kIdescB = idesc_n(L1_NB ? L1_NB : 16);

// hi = fp16(v) (round to nearest, saturating), lo = fp16(v - hi) for two values; v >= 0 is the caller's business
__device__ __forceinline__ void split2(float v0, float v1, uint32_t &hi, uint32_t &lo) {
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(v1), "f"(v0));
    const float2 back = __half22float2(*reinterpret_cast<const __half2 *>(&hi));
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(v1 - back.y), "f"(v0 - back.x));
}

// The same for relu(v): hi rounds toward zero, so v - hi >= 0 wherever v > 0 and is v itself (< 0) wherever the ReLU'd hi
// is 0 — a second ReLU inside the lo conversion finishes the job.  Three instructions per value; lo can be a whole fp16
// ulp of hi (round to nearest: half), i.e. one bit less than split2, the bit 3xTF32's truncation gives up as well.
__device__ __forceinline__ void split2_relu(float v0, float v1, uint32_t &hi, uint32_t &lo) {
    asm("cvt.rz.relu.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(v1), "f"(v0));
    const float2 back = __half22float2(*reinterpret_cast<const __half2 *>(&hi));
    asm("cvt.rn.relu.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(v1 - back.y), "f"(v0 - back.x));
}

// ENV: 0 = policy only; 1 / 2 = `mg_policy_step`: MergeEnv.step (pve / pvp) of every finished tile on two extra warps — the
// epilogue threads drop the tile's arg-max into shared memory (mgpe::Handoff) instead of the action array, the env warps
// apply the exploration rule, step the envs and write the next observation rows (possibly into the buffer being read:
// the rows are then read with coherent loads).
template <int IN, int OUT, bool MIRROR, int ENV>
__global__ void __launch_bounds__(ENV ? NUM_THREADS_ENV : NUM_THREADS, 1)
mlp_act_tc16_kernel(const float *obs, const uint8_t *__restrict__ goal, const int64_t n, const int obs_mode,
                    const unsigned char *__restrict__ blob, const float *__restrict__ b2, const float *__restrict__ w3,
                    const float *__restrict__ b3, uint8_t *__restrict__ act, float *__restrict__ q_out, const mgpe::Args P) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    Smem<OUT> &S = *reinterpret_cast<Smem<OUT> *>(smem_raw);
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const int64_t n_tiles = (n + TM - 1) / TM;
    const uint32_t my_tiles = blockIdx.x < n_tiles ? (uint32_t)((n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x) : 0u;

    // ---- one-time setup: barriers, weights -> smem (bulk copies only the MMA warp waits for), TMEM --------------------
    if (t == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&S.full[s], 128); mbar_init(&S.empty[s], 1); }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&S.x_full[b], 32); mbar_init(&S.x_empty[b], 1);
            mbar_init(&S.tmem_full[b], 1); mbar_init(&S.tmem_empty[b], 128);
        }
        for (int h = 0; h < 2; ++h) { mbar_init(&S.l1_full[h], 1); mbar_init(&S.l1_empty[h], 32 * CONV_WARPS); }
        mbar_init(&S.w_ready, 1);
        if (ENV) mgpe::handoff_init(S.env, TM, TM / 32);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&S.w_ready)), "r"((uint32_t)W_BYTES) : "memory");
        for (uint32_t c = 0; c < 8; ++c)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         :: "r"(smem_u32(S.w1 + c * W1_BYTES)), "l"(blob + HDR_BYTES + c * W1_BYTES), "r"((uint32_t)W1_BYTES),
                            "r"(smem_u32(&S.w_ready)) : "memory");
        S.c1 = __ldg(reinterpret_cast<const float *>(blob));
        S.c2 = __ldg(reinterpret_cast<const float *>(blob) + 1);
    }
    for (int i = t; i < OUT * H2P; i += (int)blockDim.x) {
        const int o = i / H2P, c = i - o * H2P;
        S.w3[o][c] = c < H2 ? w3[o * H2 + c] : 0.f;
    }
    for (int i = t; i < H2 + 12; i += (int)blockDim.x) S.b2[i] = i < H2 ? b2[i] : 0.f;
    if (t < OUT) S.b3[t] = b3[t];
    if (warp == MMA_WARP) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&S.tmem_base)),
                     "r"((uint32_t)TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = S.tmem_base;
    // From here on observations are read: wait for the previous kernel of the stream (a no-op without the PDL attribute)
    cudaGridDependencySynchronize();
    if (ENV) cudaTriggerProgrammaticLaunchCompletion();        // as in mlp_tc_kernels.cu: the next fused launch may start its prologue

    if (ENV && warp >= ENV_WARP0) {
        // =================================== ENV WARPS: MergeEnv.step of the tiles the epilogue has finished ==========
        mgpe::env_warp_loop<TM, ENV_BUFS, ENV_WARPS, ENV == 2>(P, S.env, warp - ENV_WARP0, blockIdx.x, gridDim.x, n_tiles, n, lane,
                                                               (int)(blockIdx.x % MG_STATS_ROWS));
    } else if (warp < CONV_WARPS) {
        // =================================== CONVERTERS: layer-1 accumulator -> layer-2 A operand ===================
        const int q = warp & 3, grp = warp >> 2;                        // TMEM lane quarter, K-step residue
        const uint32_t total = my_tiles * KSTEPS;
        const float c1 = S.c1;
        const int m = q * 32 + lane;                                    // the env row this thread converts
        const uint32_t off = (uint32_t)((m >> 3) * 256 + (m & 7) * 16);
        const uint32_t tl1 = tmem_base + (uint32_t)L1_COL + ((uint32_t)(q * 32) << 16);
        uint32_t cur = 0xFFFFFFFFu;
        for (uint32_t it = (uint32_t)grp; it < total; it += CONV_GROUPS) {
            const uint32_t tl = it / KSTEPS, ks = it - tl * KSTEPS;
            if (q == 0) MG_TRACE16(g_trace_conv, it, 0);
            const uint32_t half = (L1_NB && ks >= (uint32_t)KS_HALF) ? 1u : 0u;
            if (2u * tl + half != cur) {                                // this half of the tile's layer 1 has been computed
                mbar_wait(&S.l1_full[half], tl & 1u);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                cur = 2u * tl + half;
            }
            if (q == 0) MG_TRACE16(g_trace_conv, it, 1);
            uint32_t v[16];
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                         : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                           "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                         : "r"(tl1 + 16u * ks));
            // the ring slot is waited for HERE, under the tensor-memory load's latency: it is almost always free already (the
            // converters run a K-step or two ahead of the MMAs, the ring holds eight), and a barrier poll is ~100 cycles anyway
            const uint32_t s = it % STAGES;
            mbar_wait(&S.empty[s], ((it / STAGES) & 1u) ^ 1u);          // the MMAs that read this slot last have completed
            if (q == 0) MG_TRACE16(g_trace_conv, it, 4);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (q == 0) MG_TRACE16(g_trace_conv, it, 2);
            if (ks + CONV_GROUPS >= ((half || !L1_NB) ? (uint32_t)KSTEPS : (uint32_t)KS_HALF)) {   // this warp's last K-step of the half:
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");   // the next tile's layer 1 may overwrite it
                mbar_arrive(&S.l1_empty[half]);
            }
            uint32_t hi[8], lo[8];
#pragma unroll
            for (int p = 0; p < 8; ++p)
                split2_relu(__uint_as_float(v[2 * p]) * c1, __uint_as_float(v[2 * p + 1]) * c1, hi[p], lo[p]);
            if (q == 0) MG_TRACE16(g_trace_conv, it, 3);
            *reinterpret_cast<uint4 *>(S.a_hi[s] + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);          // units 0-7
            *reinterpret_cast<uint4 *>(S.a_hi[s] + off + 128) = make_uint4(hi[4], hi[5], hi[6], hi[7]);    // units 8-15
            *reinterpret_cast<uint4 *>(S.a_lo[s] + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
            *reinterpret_cast<uint4 *>(S.a_lo[s] + off + 128) = make_uint4(lo[4], lo[5], lo[6], lo[7]);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");                                   // visible to the tensor core
            mbar_arrive(&S.full[s]);
            if (q == 0) MG_TRACE16(g_trace_conv, it, 5);
        }
    } else if (warp == X_WARP) {
        // =================================== X LOADER: observation rows -> layer-1 A operand ========================
        // lane = envs {lane, lane+32, lane+64, lane+96} of the tile; K slots 0..IN-1 = the network input, 15 = 1.0 (bias)
        for (uint32_t tl = 0; tl < my_tiles; ++tl) {
            const uint32_t b = tl & 1u;
            const int64_t e0 = ((int64_t)blockIdx.x + (int64_t)tl * gridDim.x) * TM + lane;
            float x[4][IN];
#pragma unroll
            for (int j = 0; j < 4; ++j) load_row<IN, MIRROR, ENV != 0>(obs, goal, e0 + 32 * j, n, obs_mode, x[j]);
            mbar_wait(&S.x_empty[b], ((tl >> 1) & 1u) ^ 1u);            // layer 1 of tile tl - 2 has read this buffer
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int m = lane + 32 * j;
                const uint32_t off = (uint32_t)((m >> 3) * 256 + (m & 7) * 16);
                float v[16];
#pragma unroll
                for (int k = 0; k < 16; ++k) v[k] = k < IN ? x[j][k] : k == 15 ? 1.f : 0.f;
                uint32_t hi[8], lo[8];
#pragma unroll
                for (int p = 0; p < 8; ++p) split2(v[2 * p], v[2 * p + 1], hi[p], lo[p]);
                *reinterpret_cast<uint4 *>(S.x_hi[b] + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                *reinterpret_cast<uint4 *>(S.x_hi[b] + off + 128) = make_uint4(hi[4], hi[5], hi[6], hi[7]);
                *reinterpret_cast<uint4 *>(S.x_lo[b] + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
                *reinterpret_cast<uint4 *>(S.x_lo[b] + off + 128) = make_uint4(lo[4], lo[5], lo[6], lo[7]);
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_arrive(&S.x_full[b]);
        }
    } else if (warp == MMA_WARP) {
        // =================================== MMA ISSUER =============================================================
        // The warp runs its loop whole and issues by predication from one elected lane (see mlp_tc_kernels.cu).  Low
        // descriptor words: (address >> 4) | (LBO >> 4) << 16; stepping an operand = adding (bytes >> 4).
        const uint32_t w1_hi = (uint32_t)make_desc(smem_u32(S.w1)), w1_lo = w1_hi + ((W1_BYTES / 2) >> 4);
        const uint32_t w2_hi0 = (uint32_t)make_desc(smem_u32(S.w2)), a_hi0 = (uint32_t)make_desc(smem_u32(S.a_hi[0])),
                       a_lo0 = (uint32_t)make_desc(smem_u32(S.a_lo[0])), x_hi0 = (uint32_t)make_desc(smem_u32(S.x_hi[0])),
                       x_lo0 = (uint32_t)make_desc(smem_u32(S.x_lo[0]));
        const uint32_t deschi = (uint32_t)(kDescHi >> 32);
        const uint32_t d1 = tmem_base + (uint32_t)L1_COL;
        mbar_wait(&S.w_ready, 0u);                                      // the bulk copies of the weights have landed
        // one half of a tile's layer 1: hidden units [0,112) (half 0) or [112,208) (half 1) into their own accumulator columns;
        // half 1 is the last reader of the tile's observation operand
        auto layer1 = [&](const uint32_t tl, const uint32_t half) {
            const uint32_t b = tl & 1u;
            MG_TRACE16(g_trace_l1, 2u * tl + half, 0);
            mbar_wait(&S.x_full[b], (tl >> 1) & 1u);                    // the tile's observation operand is stored
            mbar_wait(&S.l1_empty[half], (tl & 1u) ^ 1u);               // the converters have read tile tl - 1's columns
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t rows = half ? (uint32_t)(L1_NA / 8) * (256u >> 4) : 0u;        // descriptor offset of row group 14
            asm volatile(
                "{\n\t.reg .pred E, X;\n\t.reg .b64 dxh, dxl, dwh, dwl;\n\t"
                "elect.sync _|E, 0xffffffff;\n\t"
                "setp.ne.and.b32 X, %9, 0, E;\n\t"
                "mov.b64 dxh, {%1, %5};\n\tmov.b64 dxl, {%2, %5};\n\tmov.b64 dwh, {%3, %5};\n\tmov.b64 dwl, {%4, %5};\n\t"
                "@E tcgen05.mma.cta_group::1.kind::f16 [%0], dxh, dwh, %6, 0;\n\t"
                "@E tcgen05.mma.cta_group::1.kind::f16 [%0], dxh, dwl, %6, 1;\n\t"
                "@E tcgen05.mma.cta_group::1.kind::f16 [%0], dxl, dwh, %6, 1;\n\t"
                "@E tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%7];\n\t"
                "@X tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%8];\n\t}\n"
                :: "r"(d1 + (half ? (uint32_t)L1_NA : 0u)), "r"(x_hi0 + b * (A_STEP >> 4)), "r"(x_lo0 + b * (A_STEP >> 4)),
                   "r"(w1_hi + rows), "r"(w1_lo + rows), "r"(deschi), "r"(half ? kIdescB : kIdescA),
                   "r"(smem_u32(&S.l1_full[half])), "r"(smem_u32(&S.x_empty[b])), "r"((half || !L1_NB) ? 1u : 0u)
                : "memory");
            MG_TRACE16(g_trace_l1, 2u * tl + half, 1);
        };
        if (my_tiles) { layer1(0u, 0u); if (L1_NB) layer1(0u, 1u); }
        for (uint32_t tl = 0; tl < my_tiles; ++tl) {
            const uint32_t buf = tl & 1u;
            mbar_wait(&S.tmem_empty[buf], ((tl >> 1) & 1u) ^ 1u);       // the epilogue has read this accumulator buffer
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t d = tmem_base + (buf ? (uint32_t)L2_COL1 : (uint32_t)L2_COL0);
            const uint32_t tm_full = smem_u32(&S.tmem_full[buf]);
            uint32_t l1_next = tl + 1 < my_tiles ? 0u : 2u;             // next half of tile tl + 1's layer 1 to issue
            uint32_t have = 0;                                          // the barrier about to be waited for was already seen complete
            // One K-step; the loop around it is fully unrolled: the issuing warp is a chain of latencies (barrier poll,
            // fence, three MMAs and a commit from one thread) and rolled up, with the layer-1 poll in every trip, it took
            // ~520 cycles per K-step against ~360 of tensor work (profiles/exp_tc16_trace.cu).
            auto kstep = [&](const uint32_t ks) {
                const uint32_t it = tl * KSTEPS + ks, s = it % STAGES;
                // everything the elected lane needs is computed BEFORE the wait, in ordinary registers
                uint32_t lo_ah = a_hi0 + s * (A_STEP >> 4), lo_al = a_lo0 + s * (A_STEP >> 4);
                uint32_t lo_bh = w2_hi0 + ks * (W2_STEP >> 4), lo_bl = lo_bh + ((W2_STEP / 2) >> 4);
                uint32_t done_bar = smem_u32(&S.empty[s]);
                asm volatile("" : "+r"(lo_ah), "+r"(lo_al), "+r"(lo_bh), "+r"(lo_bl), "+r"(done_bar));   // pin the values here
                const uint32_t last = ks + 1 == (uint32_t)KSTEPS ? 1u : 0u;
                MG_TRACE16(g_trace_mma, it, 0);
                if (!have) mbar_wait(&S.full[s], (it / STAGES) & 1u);
                // poll the NEXT K-step now: the answer arrives while the MMAs below are being issued
                have = last ? 0u : mbar_test(&S.full[(it + 1) % STAGES], ((it + 1) / STAGES) & 1u);
                MG_TRACE16(g_trace_mma, it, 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                asm volatile(
                    "{\n\t.reg .pred E, L, P;\n\t.reg .b64 dah, dal, dbh, dbl;\n\t"
                    "elect.sync _|E, 0xffffffff;\n\t"
                    "setp.ne.and.b32 L, %8, 0, E;\n\t"
                    "setp.ne.b32 P, %7, 0;\n\t"
                    "mov.b64 dah, {%1, %5};\n\tmov.b64 dal, {%2, %5};\n\tmov.b64 dbh, {%3, %5};\n\tmov.b64 dbl, {%4, %5};\n\t"
                    "@E tcgen05.mma.cta_group::1.kind::f16 [%0], dah, dbh, %6, P;\n\t"
                    "@E tcgen05.mma.cta_group::1.kind::f16 [%0], dah, dbl, %6, 1;\n\t"
                    "@E tcgen05.mma.cta_group::1.kind::f16 [%0], dal, dbh, %6, 1;\n\t"
                    "@E tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%9];\n\t"
                    "@L tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%10];\n\t}\n"
                    :: "r"(d), "r"(lo_ah), "r"(lo_al), "r"(lo_bh), "r"(lo_bl), "r"(deschi), "r"(kIdesc112), "r"(ks), "r"(last),
                       "r"(done_bar), "r"(tm_full)
                    : "memory");
                MG_TRACE16(g_trace_mma, it, 2);
                // The next tile's layer 1: its first half as soon as the converters have read this tile's first half (polled at
                // the K-steps around which that happens — they run a K-step or two ahead of these MMAs), its second half
                // behind this tile's last K-step (the converters have stored that K-step, so they have read everything).
                if (ks >= (uint32_t)L1_POLL_FROM && l1_next == 0u &&
                    (ks + 1 == (uint32_t)KSTEPS || __all_sync(0xFFFFFFFFu, mbar_test(&S.l1_empty[0], tl & 1u) != 0u))) {
                    layer1(tl + 1, 0u);
                    l1_next = L1_NB ? 1u : 2u;
                }
                if (ks + 1 == (uint32_t)KSTEPS && l1_next == 1u) {
                    layer1(tl + 1, 1u);
                    l1_next = 2u;
                }
                MG_TRACE16(g_trace_mma, it, 3);
            };
#pragma unroll
            for (int ks = 0; ks < KSTEPS; ++ks) kstep((uint32_t)ks);
        }
    } else {
        // =================================== EPILOGUE: layer 3 + arg-max ============================================
        // tcgen05.ld.16x256b.x2: lanes 16h .. 16h+15 of this warp's TMEM quarter, 16 columns; thread (t1 = lane / 4,
        // t0 = lane % 4) receives rows t1 and t1 + 8 at columns 8b + 2 t0 + {0, 1} of both 8-column blocks b:
        //   r[0..1] = (row t1, block 0)  r[2..3] = (row t1+8, block 0)  r[4..5] = (row t1, block 1)  r[6..7] = (row t1+8, block 1)
        const int q4 = warp - EPI_WARP0;                                // TMEM lane quarter of this warp (warp % 4)
        const int t0 = lane & 3, t1 = lane >> 2;
        const float c2 = S.c2;
        for (uint32_t tl = 0; tl < my_tiles; ++tl) {
            const int64_t tile = (int64_t)blockIdx.x + (int64_t)tl * gridDim.x;
            const uint32_t buf = tl & 1u;
            if (q4 == 0) MG_TRACE16(g_trace_epi, tl, 0);
            mbar_wait(&S.tmem_full[buf], (tl >> 1) & 1u);
            if (q4 == 0) MG_TRACE16(g_trace_epi, tl, 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t taddr = tmem_base + (buf ? (uint32_t)L2_COL1 : (uint32_t)L2_COL0) + ((uint32_t)(q4 * 32) << 16);
            float q[4][OUT];                                            // rows t1 + 8 * {0, 1, 2, 3}: partial sums over this thread's neurons
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int o = 0; o < OUT; ++o) q[r][o] = 0.f;
            // 32 columns per trip (tcgen05.ld .x4: four 8-column blocks for rows t1 / t1 + 8 of both 16-lane halves), and the
            // NEXT trip's columns are requested before this trip's arithmetic starts (two register sets, A and B, two trips
            // per loop pass): the epilogue paces the kernel, and every tensor-memory load latency it exposes is tile time
            // (profiles/exp_tc16_trace.cu: 16 columns per trip, nothing in flight: busy 5 800 of a tile's 6 000 cycles).
            auto request = [&](uint32_t (&a)[2][16], const int cb) {
#pragma unroll
                for (int h = 0; h < 2; ++h)
                    asm volatile("tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                                 : "=r"(a[h][0]), "=r"(a[h][1]), "=r"(a[h][2]), "=r"(a[h][3]), "=r"(a[h][4]), "=r"(a[h][5]),
                                   "=r"(a[h][6]), "=r"(a[h][7]), "=r"(a[h][8]), "=r"(a[h][9]), "=r"(a[h][10]), "=r"(a[h][11]),
                                   "=r"(a[h][12]), "=r"(a[h][13]), "=r"(a[h][14]), "=r"(a[h][15])
                                 : "r"(taddr + ((uint32_t)(16 * h) << 16) + (uint32_t)(32 * cb)));
            };
            auto layer3 = [&](const uint32_t (&a)[2][16], const int cb) {
#pragma unroll
                for (int blk = 0; blk < 4; ++blk) {
                    if (32 * cb + 8 * blk < H2P) {                      // blocks 0..12 hold the 100 neurons (+4 zero-weight pads)
                        const int c = 32 * cb + 8 * blk + 2 * t0;
                        const float2 bias = *reinterpret_cast<const float2 *>(&S.b2[c]);
                        float2 w[OUT];
#pragma unroll
                        for (int o = 0; o < OUT; ++o) w[o] = *reinterpret_cast<const float2 *>(&S.w3[o][c]);
#pragma unroll
                        for (int h = 0; h < 2; ++h)
#pragma unroll
                            for (int rr = 0; rr < 2; ++rr) {
                                const int i0 = 4 * blk + 2 * rr;
                                const float h0 = fmaxf(fmaf(__uint_as_float(a[h][i0]), c2, bias.x), 0.f);
                                const float h1 = fmaxf(fmaf(__uint_as_float(a[h][i0 + 1]), c2, bias.y), 0.f);
#pragma unroll
                                for (int o = 0; o < OUT; ++o) {
                                    q[2 * h + rr][o] = fmaf(h0, w[o].x, q[2 * h + rr][o]);
                                    q[2 * h + rr][o] = fmaf(h1, w[o].y, q[2 * h + rr][o]);
                                }
                            }
                    }
                }
            };
            static_assert((UN + 31) / 32 == 4, "two loop passes of two 32-column trips");
            uint32_t ra[2][16], rb[2][16];
            request(ra, 0);
#pragma unroll 1
            for (int p = 0; p < 2; ++p) {                               // rolled up: the body stays in the instruction cache
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                request(rb, 2 * p + 1);
                layer3(ra, 2 * p);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (p == 0) {
                    request(ra, 2);
                } else {                                                // everything read: the MMAs of tile tl + 2 may overwrite it
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    mbar_arrive(&S.tmem_empty[buf]);
                }
                layer3(rb, 2 * p + 1);
            }
            // sum the 4 lanes that share a row group, then lane t0 finishes row t1 + 8 * t0
            float mine[OUT];
#pragma unroll
            for (int o = 0; o < OUT; ++o) {
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    q[r][o] += __shfl_xor_sync(0xffffffffu, q[r][o], 1);
                    q[r][o] += __shfl_xor_sync(0xffffffffu, q[r][o], 2);
                }
                mine[o] = (t0 == 0 ? q[0][o] : t0 == 1 ? q[1][o] : t0 == 2 ? q[2][o] : q[3][o]) + S.b3[o];
            }
            if (q4 == 0) MG_TRACE16(g_trace_epi, tl, 2);
            const int64_t e = tile * TM + q4 * 32 + t1 + 8 * t0;
            int best = 0;
            float bv = mine[0];
#pragma unroll
            for (int o = 1; o < OUT; ++o)
                if (mine[o] > bv) { bv = mine[o]; best = o; }           // first maximum, like torch.max
            if (e < n) {
                if (!ENV) act[e] = (uint8_t)best;
                if (!ENV && (obs_mode & 0x100)) const_cast<float *>(obs)[e * (MG_OBS_DIM + 1)] = (float)best;   // MG_MLP_FLAG_WRITE_GOAL
                if (q_out) {
#pragma unroll
                    for (int o = 0; o < OUT; ++o) q_out[e * OUT + o] = mine[o];
                }
            }
            if (ENV) {                                                  // hand the action to the env warps
                uint8_t *tile_act = mgpe::handoff_acquire(S.env, tl);
                tile_act[q4 * 32 + t1 + 8 * t0] = (uint8_t)best;
                mgpe::handoff_publish(S.env, tl);
            }
        }
    }
    // ---- teardown -------------------------------------------------------------------------------------
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == MMA_WARP)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS));
}

template <int IN, int OUT, bool MIRROR, int ENV = 0>
cudaError_t launch(const float *obs, const uint8_t *goal, int64_t n, int obs_mode, const unsigned char *blob, const float *b2,
                   const float *w3, const float *b3, uint8_t *act, float *q_out, cudaStream_t st, bool pdl,
                   const mgpe::Args &P = mgpe::Args{}) {
    auto kern = mlp_act_tc16_kernel<IN, OUT, MIRROR, ENV>;
    const size_t smem = sizeof(Smem<OUT>) + 1024;               // slack for the 1024-byte alignment of the base
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e) return e;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int64_t tiles = (n + TM - 1) / TM;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)(tiles < sms ? tiles : sms)); cfg.blockDim = dim3(ENV ? NUM_THREADS_ENV : NUM_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
    e = cudaLaunchKernelEx(&cfg, kern, obs, goal, n, obs_mode, blob, b2, w3, b3, act, q_out, P);
    return e ? e : cudaGetLastError();
}

}  // namespace mgtc16

// called by mg_mlp_act_tc (mlp_tc_kernels.cu) for MG_MLP_FLAG_F16X3; arguments already validated there
cudaError_t mg_mlp_act_tc16_launch(int in_dim, int out_dim, bool mirror, const float *obs, const uint8_t *goal, int64_t n,
                                   int obs_mode, const void *blob, const float *b2, const float *w3, const float *b3,
                                   uint8_t *act, float *q_out, cudaStream_t st, bool pdl) {
    const auto *bl = static_cast<const unsigned char *>(blob);
#define MG_TC16_CASE(I, O)                                                                                                  \
    if (in_dim == I && out_dim == O)                                                                                        \
        return mirror ? mgtc16::launch<I, O, true>(obs, goal, n, obs_mode, bl, b2, w3, b3, act, q_out, st, pdl)             \
                      : mgtc16::launch<I, O, false>(obs, goal, n, obs_mode, bl, b2, w3, b3, act, q_out, st, pdl);
    MG_TC16_CASE(10, 5) MG_TC16_CASE(10, 3) MG_TC16_CASE(11, 5) MG_TC16_CASE(11, 3)
#undef MG_TC16_CASE
    return cudaErrorInvalidValue;
}

// the fused policy + env step on the f16x3 backend (called by mg_policy_step in mlp_kernels.cu; `blob` as above)
cudaError_t mg_policy_step_tc16_launch(int in_dim, const float *obs, const uint8_t *goal, int64_t n, const void *blob, const float *b2,
                                       const float *w3, const float *b3, float *q_out, cudaStream_t st, const mgpe::Args &P) {
    const bool pvp = P.a2 != nullptr, pdl = (P.flags & MG_POLICY_FLAG_PDL) != 0u;
    const int obs_mode = (int)mg::obs_layout_of(P.flags);
    const auto *bl = static_cast<const unsigned char *>(blob);
#define MG_TC16_ENV(I) (pvp ? mgtc16::launch<I, 5, false, 2>(obs, goal, n, obs_mode, bl, b2, w3, b3, nullptr, q_out, st, pdl, P) \
                            : mgtc16::launch<I, 5, false, 1>(obs, goal, n, obs_mode, bl, b2, w3, b3, nullptr, q_out, st, pdl, P))
    if (in_dim == 10) return MG_TC16_ENV(10);
    if (in_dim == 11) return MG_TC16_ENV(11);
#undef MG_TC16_ENV
    return cudaErrorInvalidValue;
}
