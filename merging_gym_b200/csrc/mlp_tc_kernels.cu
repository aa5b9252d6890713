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
//   warps 0-7  producers : warp w owns every 8th K-step of the CTA's K-step sequence (25 per tile); a thread
//                          computes 8 hidden units of layer 1 for 4 envs on the CUDA cores (FFMA2, each weight
//                          read from shared memory feeds 4 envs), splits them into hi/lo and writes them straight
//                          into the canonical K-major core-matrix layout of ring slot (K-step % 4) -> full[slot]
//   warp  12   MMA issue : per K-step two tcgen05.mma.kind::tf32 — a_hi x [W2_hi ; W2_lo] (N = 224) and
//                          a_lo x W2_hi (N = 112) — issued by predication from one elected lane, K loop unrolled,
//                          accumulating in TMEM (always accumulate: buffers are handed back zeroed); tcgen05.commit -> empty[next producer] / tmem_full[b]
//   warps 8-11 epilogue  : tcgen05.ld.16x256b fragments of the 128x112 fp32 accumulator — a thread holds 4 envs x
//                          2 adjacent neurons per 8-column block, so one read of the layer-3 weights feeds 4 envs —
//                          bias + ReLU, the 100x{5,3} layer, a 4-lane shuffle reduction and the arg-max
// TMEM: 2 accumulator buffers x 256 columns ([0,112): hi.hi + lo.hi, [112,224): hi.lo, summed in the epilogue),
// so the epilogue of tile i overlaps the MMAs of tile i+1.
// Shared memory: W2 as one stacked B operand [hi ; lo] of 224 rows, 179.2 KB (prepared on the host), A ring 32 KB
// (4 slots of one K-step), W1 / W3 / biases 11 KB.  The kernel is bound by shared-memory bandwidth: ~320 wavefronts per
// K-step (tensor-core operand reads 148, producer stores 70, weight broadcasts 100) against a K-step period of ~336
// cycles — which is what the two 4-envs-per-thread mappings cut (the first version moved 470).  The hand-over
// timeline is measured by profiles/exp_tc_trace.cu (-DMG_TC_TRACE=1).
// Every mbarrier wait is bounded and traps instead of hanging.
#include "abi_common.h"
// ptxas 12.9 crashes on setmaxnreg in a kernel that also CALLS a function: the random-start draw is inlined in this file
#define MG_RANDOM_START_INLINE 1
#include "policy_env.cuh"
#include "tc_common.cuh"

// 1 = a single MMA-issuing warp with its K loop fully unrolled: accumulation order in TMEM is the K-step order, the
//     kernel is bitwise reproducible (69.6 us per 2^18 envs);
// 2 = two issuing warps with interleaved K-steps (68.6 us): fp32 accumulation order follows the run-to-run
//     interleaving of the two instruction streams, so 57 % of the Q rows differ in their last bits between launches.
#ifndef MG_TC_MMA_WARPS
#define MG_TC_MMA_WARPS 1
#endif
#ifndef MG_TC_TRACE
#define MG_TC_TRACE 0            // 1: CTA 0 records clock64() at the hand-over points (profiles/exp_tc_trace.cu)
#endif

namespace mgtc {

#if MG_TC_TRACE
constexpr int TRACE_G = 25 * 10;                // K-steps traced (10 tiles of CTA 0)
__device__ long long g_trace_prod[TRACE_G][5];  // compute start, wait start, wait end, fence done, arrive done
__device__ long long g_trace_mma[TRACE_G][8];   // wait start, wait end, issued
__device__ long long g_trace_epi[16][3];        // wait start, wait end, done (warp 8)
#define MG_TRACE(arr, idx, k) do { if (blockIdx.x == 0 && lane == 0 && (idx) < (uint32_t)(sizeof(arr) / sizeof(arr[0]))) arr[idx][k] = clock64(); } while (0)
#else
#define MG_TRACE(arr, idx, k) do { } while (0)
#endif

constexpr int H1 = 200, H2 = 100;
constexpr int UN = 112;                       // UMMA N (multiple of 16 for M = 128)
constexpr int KSTEPS = H1 / 8;                // 25 K-steps of 8 (tf32: 32 bytes of K per MMA)
constexpr int STAGES = 4;                     // ring slots; one slot = one K-step (8 hidden units) of all 128 envs
constexpr int H2P = 104;                      // layer-3 weights padded to whole 8-column blocks
constexpr int A_STEP = (TM / 8) * 256;        // 4096 B : 16 row groups x 2 core matrices x 128 B
constexpr int B_STEP = (2 * UN / 8) * 256;    // 7168 B: one K-step of the stacked B operand [W2_hi ; W2_lo] (224 rows)
constexpr int B_BYTES = KSTEPS * B_STEP;      // 179 200 B
constexpr int TMEM_COLS = 512;                // 2 accumulator buffers x 256 columns (224 used)
constexpr int PRODUCER_WARPS = 8;             // thread = 4 envs x one K-step
constexpr int MMA_WARPS = MG_TC_MMA_WARPS;     // warps 12.. issue the MMAs, K-steps interleaved
constexpr int NUM_THREADS = 384 + 32 * MMA_WARPS;   // 8 producer warps + 4 epilogue warps + the MMA warps
// mg_policy_step only: ENV_WARPS more warps own the env step.  MG_TC_ENV_WARPS = 3: 16 warps = 512 threads,
// every warp keeps the 128 registers of the policy-only kernel.  MG_TC_ENV_WARPS = 7 (default): 20 warps = 5 warpgroups, every
// warp is launched with 96 registers (65 536 / 640 rounded down to a multiple of 8) and each role starts by
// re-partitioning the CTA's register pool with setmaxnreg: producers (warps 0-7) 112, epilogue (8-11) 104, the
// warpgroup of the MMA warp and env warps 13-15 88, env warps 16-19 64 — 28 672 + 13 312 + 11 264 + 8 192 = 61 440 =
// 640 x 96: the pool is what the CTA was launched with, NOT the SM's 65 536 (a partition that sums to more deadlocks
// in setmaxnreg.inc).  Measured at 2^18 envs: 3 warps 86.4 us, 7 warps 82.7 us (the env warps then idle 73 % of the time;
// profiles/r02_policy_step_tc_regions.md).
#ifndef MG_TC_ENV_WARPS
#define MG_TC_ENV_WARPS 7
#endif
constexpr int ENV_WARP0 = 12 + MMA_WARPS;
constexpr int ENV_WARPS = MG_TC_ENV_WARPS;
constexpr bool ENV_SETMAXNREG = ENV_WARPS > 3;      // more than 16 warps: the register file has to be re-partitioned
constexpr int ENV_BUFS = 4;                         // action tiles in flight between the epilogue and the env warps
constexpr int NUM_THREADS_ENV = 32 * (ENV_WARP0 + ENV_WARPS);
static_assert(MMA_WARPS == 1 && (ENV_WARPS == 3 || ENV_WARPS == 7), "register budgets are laid out for 16 or 20 warps");
static_assert(256 * 112 + 128 * 104 + 128 * 88 + 128 * 64 <= 640 * 96, "setmaxnreg partition exceeds the CTA pool");
constexpr int MAX_OUT = 8;

template <int IN, int OUT>
struct Smem {
    unsigned char b_cat[B_BYTES];             // [K-step][28 row groups: W2_hi rows 0-111, W2_lo rows 112-223], canonical layout
    unsigned char a_hi[STAGES][A_STEP];
    unsigned char a_lo[STAGES][A_STEP];
    float w1[IN][H1];
    float w3[OUT][H2P];
    float b1[H1], b2[H2 + 12], b3[MAX_OUT];
    unsigned long long full[PRODUCER_WARPS], empty[PRODUCER_WARPS], tmem_full[2], tmem_empty[2];
    unsigned long long w2_ready;              // completes when the bulk copies of b_cat have landed
    uint32_t tmem_base;
    mgpe::Handoff<TM, ENV_BUFS> env;                     // ENV: the tile's actions, epilogue warps -> env warp
};

// instruction descriptor: D = f32, A = B = tf32, both K-major, N >> 3 at bit 17, M >> 4 at bit 24
constexpr uint32_t idesc_n(int n) { return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(TM >> 4) << 24); }
constexpr uint32_t kIdesc112 = idesc_n(UN), kIdesc224 = idesc_n(2 * UN);

// zero 16 columns of the calling warp's 32 TMEM lanes
__device__ __forceinline__ void tmem_zero16(uint32_t taddr) {
    const uint32_t z = 0u;
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(taddr), "r"(z) : "memory");
}
// ENV: `mg_policy_step` — the epilogue threads drop the tile's actions into shared memory and one extra warp steps the
// tile's envs (policy_env.cuh) while the other warps are on the next tile.  (Stepping the env in the epilogue threads
// themselves doubled the tile period: the epilogue warps have no slack, profiles/r02_policy_step_*.)
// ENV: 0 = policy only, 1 = + env step pve, 2 = + env step pvp
template <int IN, int OUT, bool MIRROR, int ENV>
__global__ void __launch_bounds__(ENV ? NUM_THREADS_ENV : NUM_THREADS, 1)
mlp_act_tc_kernel(const float *obs, const uint8_t *__restrict__ goal, const int64_t n, const int obs_mode,
                  const float *__restrict__ w1t, const float *__restrict__ b1, const float *__restrict__ w2_tc,
                  const float *__restrict__ b2, const float *__restrict__ w3, const float *__restrict__ b3,
                  uint8_t *__restrict__ act, float *__restrict__ q_out, const mgpe::Args P) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    Smem<IN, OUT> &S = *reinterpret_cast<Smem<IN, OUT> *>(smem_raw);
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const int64_t n_tiles = (n + TM - 1) / TM;

    // ---- one-time setup: barriers, weights -> smem, TMEM ---------------------------------------------
    // Nothing in this prologue reads anything but the weights, so under programmatic dependent launch (MG_MLP_FLAG_PDL)
    // it runs while the previous kernel of the stream is still draining.  W2 (179 KB, already in the canonical UMMA
    // layout) comes by bulk-copy engine: one thread issues five cp.async.bulk and only the MMA warp ever waits for
    // them, so layer 1 of the first tile starts while the copy is in flight (staging it with LDG/STS through all
    // threads cost ~4 of the kernel's ~10 us of fixed time, profiles/r02_policy_lane_probe.json).
    if (t == 0) {
        for (int s = 0; s < PRODUCER_WARPS; ++s) { mbar_init(&S.full[s], 32); mbar_init(&S.empty[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&S.tmem_full[b], MMA_WARPS); mbar_init(&S.tmem_empty[b], TM); }
        mbar_init(&S.w2_ready, 1);
        if (ENV) mgpe::handoff_init(S.env, TM, TM / 32);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        constexpr uint32_t kChunk = B_BYTES / 5;      // 35 840 B, a multiple of 16
        static_assert(kChunk * 5 == B_BYTES && kChunk % 16 == 0, "W2 is copied in five equal 16-byte-aligned pieces");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&S.w2_ready)), "r"((uint32_t)B_BYTES) : "memory");
        for (uint32_t c = 0; c < 5; ++c)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         :: "r"(smem_u32(S.b_cat + c * kChunk)), "l"(reinterpret_cast<const unsigned char *>(w2_tc) + c * kChunk),
                            "r"(kChunk), "r"(smem_u32(&S.w2_ready)) : "memory");
    }
    {
        const float4 *s1 = reinterpret_cast<const float4 *>(w1t);
        float4 *d1 = reinterpret_cast<float4 *>(&S.w1[0][0]);
        for (int i = t; i < IN * H1 / 4; i += (int)blockDim.x) d1[i] = __ldg(s1 + i);
        for (int i = t; i < OUT * H2P; i += (int)blockDim.x) {
            const int o = i / H2P, c = i - o * H2P;
            S.w3[o][c] = c < H2 ? w3[o * H2 + c] : 0.f;
        }
        for (int i = t; i < H1; i += (int)blockDim.x) S.b1[i] = b1[i];
        for (int i = t; i < H2 + 12; i += (int)blockDim.x) S.b2[i] = i < H2 ? b2[i] : 0.f;
        if (t < OUT) S.b3[t] = b3[t];
    }
    if (warp == 12) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&S.tmem_base)),
                     "r"((uint32_t)TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = S.tmem_base;
    // Every MMA accumulates (no overwriting first K-step: two warps issue, and only accumulation commutes), so the
    // accumulators start at zero and each epilogue zeroes what it has read.
    if (warp >= 8 && warp < 12) {
        const uint32_t lanes = (uint32_t)((warp - 8) * 32) << 16;
        for (uint32_t c = 0; c < (uint32_t)TMEM_COLS; c += 16) tmem_zero16(tmem_base + lanes + c);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // From here on observations and env state are read: wait for the previous kernel of the stream (a no-op without
    // the PDL launch attribute), and let the NEXT launch start its own prologue right away.
    cudaGridDependencySynchronize();
    if (ENV) cudaTriggerProgrammaticLaunchCompletion();        // the next fused launch can only land on other SMs anyway
                                                               // (an early trigger in front of mg_step made that step slower)

    if (warp < PRODUCER_WARPS) {
        if (ENV && ENV_SETMAXNREG) asm volatile("setmaxnreg.inc.sync.aligned.u32 112;");
        // =================================== PRODUCERS: layer 1 ===================================
        // The CTA's K-steps are numbered g = 25 * (local tile) + ks; warp w produces g = w, w + 8, ... into ring
        // slot g % 4.  Its layer-1 arithmetic for K-step g runs while the MMAs of g-8 .. g-1 are in flight; only the
        // stores wait for the slot.  lane = envs {lane, lane+32, lane+64, lane+96} of the tile.
        const int64_t my_tiles = blockIdx.x < n_tiles ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
        const uint32_t total = (uint32_t)my_tiles * KSTEPS;
        uint32_t off[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int m = lane + 32 * j;
            off[j] = (uint32_t)((m >> 3) * 256 + (m & 7) * 16);
        }
        float x[4][IN];
        {
            const int64_t e0 = (int64_t)blockIdx.x * TM + lane;
#pragma unroll
            for (int j = 0; j < 4; ++j) load_row<IN, MIRROR, ENV != 0>(obs, goal, e0 + 32 * j, n, obs_mode, x[j]);
        }
        for (uint32_t g = (uint32_t)warp; g < total; g += PRODUCER_WARPS) {
            const uint32_t tl = g / KSTEPS, ks = g - tl * KSTEPS;
            MG_TRACE(g_trace_prod, g, 0);
            const int k = 8 * (int)ks;
            const float4 ba = *reinterpret_cast<const float4 *>(&S.b1[k]);
            const float4 bb = *reinterpret_cast<const float4 *>(&S.b1[k + 4]);
            float2 acc[4][4];                                   // [env][unit pair]
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                acc[j][0] = make_float2(ba.x, ba.y); acc[j][1] = make_float2(ba.z, ba.w);
                acc[j][2] = make_float2(bb.x, bb.y); acc[j][3] = make_float2(bb.z, bb.w);
            }
#pragma unroll
            for (int i = 0; i < IN; ++i) {
                const float4 wa = *reinterpret_cast<const float4 *>(&S.w1[i][k]);
                const float4 wb = *reinterpret_cast<const float4 *>(&S.w1[i][k + 4]);
                const float2 w01 = make_float2(wa.x, wa.y), w23 = make_float2(wa.z, wa.w);
                const float2 w45 = make_float2(wb.x, wb.y), w67 = make_float2(wb.z, wb.w);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float2 xx = make_float2(x[j][i], x[j][i]);
                    acc[j][0] = __ffma2_rn(xx, w01, acc[j][0]); acc[j][1] = __ffma2_rn(xx, w23, acc[j][1]);
                    acc[j][2] = __ffma2_rn(xx, w45, acc[j][2]); acc[j][3] = __ffma2_rn(xx, w67, acc[j][3]);
                }
            }
            // This was the warp's last K-step of the tile and x[] is dead from here on: request the NEXT tile's rows
            // into it now.  They are needed at this warp's next visit, a store phase and ~2000 cycles away — loaded
            // there instead, they sit on the critical path of every tile switch (the MMAs consume K-steps in order).
            if ((g + PRODUCER_WARPS) / KSTEPS != tl) {
                const int64_t e0 = ((int64_t)blockIdx.x + (int64_t)(tl + 1) * gridDim.x) * TM + lane;
#pragma unroll
                for (int j = 0; j < 4; ++j) load_row<IN, MIRROR, ENV != 0>(obs, goal, e0 + 32 * j, n, obs_mode, x[j]);
            }
            // Ring slot s = g % 4 is filled alternately by warps s and s + 4.  A parity wait is only meaningful when
            // the waiter is at most one phase behind the barrier, so every producer warp has its OWN pair of
            // barriers: the MMA thread commits K-step k to empty[(k + 4) % 8] — the warp that reuses the slot next.
            const int s = (int)(g % STAGES);
            const uint32_t v = g / PRODUCER_WARPS;              // this warp's visit number
            MG_TRACE(g_trace_prod, g, 1);
            mbar_wait(&S.empty[warp], warp < STAGES ? (v & 1u) ^ 1u : (v & 1u));
            MG_TRACE(g_trace_prod, g, 2);
            auto split_store = [&](const float2 &p, const float2 &q, uint32_t o) {
                const float4 h = make_float4(fmaxf(p.x, 0.f), fmaxf(p.y, 0.f), fmaxf(q.x, 0.f), fmaxf(q.y, 0.f));
                float4 hi, lo;                                  // hi = what kind::tf32 reads (top 19 bits), lo exact
                hi.x = __uint_as_float(__float_as_uint(h.x) & 0xFFFFE000u); lo.x = h.x - hi.x;
                hi.y = __uint_as_float(__float_as_uint(h.y) & 0xFFFFE000u); lo.y = h.y - hi.y;
                hi.z = __uint_as_float(__float_as_uint(h.z) & 0xFFFFE000u); lo.z = h.z - hi.z;
                hi.w = __uint_as_float(__float_as_uint(h.w) & 0xFFFFE000u); lo.w = h.w - hi.w;
                *reinterpret_cast<float4 *>(S.a_hi[s] + o) = hi;
                *reinterpret_cast<float4 *>(S.a_lo[s] + o) = lo;
            };
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                split_store(acc[j][0], acc[j][1], off[j]);      // units 0-3: first core matrix of the K-step
                split_store(acc[j][2], acc[j][3], off[j] + 128);// units 4-7: second core matrix
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // visible to the tensor core
            MG_TRACE(g_trace_prod, g, 3);
            mbar_arrive(&S.full[warp]);
            MG_TRACE(g_trace_prod, g, 4);
        }
    } else if (ENV && warp >= ENV_WARP0) {
        if (ENV_SETMAXNREG) {
            if (warp < 16) asm volatile("setmaxnreg.dec.sync.aligned.u32 88;");  // same warpgroup as the MMA warp
            else asm volatile("setmaxnreg.dec.sync.aligned.u32 64;");
        }
        // =================================== ENV WARPS: MergeEnv.step of the tiles the epilogue has finished ==========
        // (keeping env warp 16 off the MMA-issuing warp's scheduler — warp % 4 — changed nothing: 90.9 vs 92.0 us)
        mgpe::env_warp_loop<TM, ENV_BUFS, ENV_WARPS, ENV == 2>(P, S.env, warp - ENV_WARP0, blockIdx.x, gridDim.x, n_tiles, n, lane,
                                                               (int)(blockIdx.x % MG_STATS_ROWS));
    } else if (warp >= 12) {
        if (ENV && ENV_SETMAXNREG) asm volatile("setmaxnreg.dec.sync.aligned.u32 88;");
        // =================================== MMA ISSUERS =========================================
        // The issuing warp runs its loop whole-warp and issues by PREDICATION from one elected lane (inside an
        // `if (lane == 0)` region the compiler rebuilt every descriptor through R2UR and wrapped each tcgen05
        // instruction in a per-lane retry loop — ~55 dependent instructions and ~470 cycles per K-step against ~180
        // cycles of tensor work); with the K loop fully unrolled one warp keeps up with the shared-memory-bound
        // K-step period.  Every MMA accumulates (the epilogue hands each buffer back zeroed); MMA_WARPS = 2 interleaves
        // the K-steps over two warps, which is 1.5 % faster but makes the fp32 accumulation order — and so the last
        // bits of the Q-values — vary from launch to launch.  tmem_full[buf] expects one commit per issuing warp.
        const int j = warp - 12;
        uint32_t tl = 0;
        // low descriptor words: (address >> 4) | (LBO >> 4) << 16; stepping an operand = adding (bytes >> 4)
        const uint32_t b_cat = (uint32_t)make_desc(smem_u32(S.b_cat)), a_hi0 = (uint32_t)make_desc(smem_u32(S.a_hi[0])),
                       a_lo0 = (uint32_t)make_desc(smem_u32(S.a_lo[0]));
        mbar_wait(&S.w2_ready, 0u);                            // the bulk copies of W2 have landed
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tl) {
            const uint32_t buf = tl & 1u;
            mbar_wait(&S.tmem_empty[buf], ((tl >> 1) & 1u) ^ 1u);                // epilogue drained (and zeroed) this buffer
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t d = tmem_base + buf * 256u;
            const uint32_t tm_full = smem_u32(&S.tmem_full[buf]);
            uint32_t have = 0;                              // the barrier about to be waited for was already seen complete
            // One K-step; the loop around it is fully unrolled in the policy-only kernel (74.7 -> 69.6 us in round 1: the
            // descriptor arithmetic of the next K-steps overlaps the waits) and unrolled 5x in the fused kernel, where
            // 1 450 straight-line instructions per tile cost more in instruction-cache misses — the env warps' code shares
            // the cache — than the overlap buys (84.7 -> 79.6 us at 2^18 envs; the policy-only kernel loses 0.5 us with it).
            auto kstep = [&](const int ks) {
                const uint32_t it = tl * KSTEPS + (uint32_t)ks;
                const uint32_t s = it % STAGES, pw = it % PRODUCER_WARPS;
                // everything the elected lane needs is computed BEFORE the wait, in ordinary registers
                uint32_t lo_a = a_hi0 + s * (A_STEP >> 4), lo_l = a_lo0 + s * (A_STEP >> 4);
                uint32_t lo_b = b_cat + (uint32_t)ks * (B_STEP >> 4);
                uint32_t done_bar = smem_u32(&S.empty[(it + STAGES) % PRODUCER_WARPS]);
                asm volatile("" : "+r"(lo_a), "+r"(lo_l), "+r"(lo_b), "+r"(done_bar));   // pin the values here (no sinking below the wait)
                const uint32_t last = ks + MMA_WARPS >= KSTEPS ? 1u : 0u;            // this warp's last K-step of the tile
                MG_TRACE(g_trace_mma, it, 0);
                if (!have) mbar_wait(&S.full[pw], (it / PRODUCER_WARPS) & 1u);
                // poll this warp's NEXT K-step now: the answer arrives while the MMAs below are being issued
                have = last ? 0u : mbar_test(&S.full[(it + MMA_WARPS) % PRODUCER_WARPS], ((it + MMA_WARPS) / PRODUCER_WARPS) & 1u);
                MG_TRACE(g_trace_mma, it, 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                // columns [0,112) += a_hi.W2_hi and [112,224) += a_hi.W2_lo in one N = 224 MMA (123 cycles instead of two
                // N = 112 MMAs at 76 each, A_hi read once); then columns [0,112) += a_lo.W2_hi.  Commits: the warp that fills
                // this slot next; after this warp's last K-step of the tile the epilogue.
                asm volatile(
                    "{\n\t.reg .pred E, L;\n\t.reg .b64 da, dl, db;\n\t"
                    "elect.sync _|E, 0xffffffff;\n\t"
                    "setp.ne.and.b32 L, %7, 0, E;\n\t"
                    "mov.b64 da, {%1, %4};\n\tmov.b64 dl, {%2, %4};\n\tmov.b64 db, {%3, %4};\n\t"
                    "@E tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %5, 1;\n\t"
                    "@E tcgen05.mma.cta_group::1.kind::tf32 [%0], dl, db, %6, 1;\n\t"
                    "@E tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%8];\n\t"
                    "@L tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%9];\n\t}\n"
                    :: "r"(d), "r"(lo_a), "r"(lo_l), "r"(lo_b), "r"((uint32_t)(kDescHi >> 32)), "r"(kIdesc224), "r"(kIdesc112),
                       "r"(last), "r"(done_bar), "r"(tm_full)
                    : "memory");
                MG_TRACE(g_trace_mma, it, 2);
            };
            if (ENV) {
#pragma unroll 5
                for (int ks = j; ks < KSTEPS; ks += MMA_WARPS) kstep(ks);
            } else {
#pragma unroll
                for (int ks = j; ks < KSTEPS; ks += MMA_WARPS) kstep(ks);
            }
        }
    } else {
        if (ENV && ENV_SETMAXNREG) asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
        // =================================== EPILOGUE: layer 3 + arg-max ===========================
        // tcgen05.ld.16x256b.x2: lanes 16h .. 16h+15 of this warp's TMEM quarter, 16 columns; thread (t1 = lane / 4,
        // t0 = lane % 4) receives rows t1 and t1 + 8 at columns 8b + 2 t0 + {0, 1} of both 8-column blocks b:
        //   r[0..1] = (row t1, block 0)  r[2..3] = (row t1+8, block 0)  r[4..5] = (row t1, block 1)  r[6..7] = (row t1+8, block 1)
        const int q4 = warp - 8;                                // TMEM lane quarter of this warp (warp % 4)
        const int t0 = lane & 3, t1 = lane >> 2;
        uint32_t tl = 0;
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tl) {
            const uint32_t buf = tl & 1u;
            if (q4 == 0) MG_TRACE(g_trace_epi, tl, 0);
            mbar_wait(&S.tmem_full[buf], (tl >> 1) & 1u);
            if (q4 == 0) MG_TRACE(g_trace_epi, tl, 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t taddr = tmem_base + buf * 256u + ((uint32_t)(q4 * 32) << 16);
            float q[4][OUT];                                    // rows t1 + 8 * {0, 1, 2, 3}: partial sums over this thread's neurons
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int o = 0; o < OUT; ++o) q[r][o] = 0.f;
            // NOT unrolled: fully unrolled, this loop is 1 000 instructions that each epilogue warp runs once per tile
            // between two producer warps on its scheduler — 62 % of the epilogue's samples were instruction-fetch
            // stalls (profiles/r02_policy_step_tc_v3: stall_no_inst 623 of 1 002); as a 7-trip loop the body stays
            // in the instruction cache.
#pragma unroll 1
            for (int cb = 0; cb < UN / 16; ++cb) {
                uint32_t a[2][8], l[2][8];
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const uint32_t ad = taddr + ((uint32_t)(16 * h) << 16) + (uint32_t)(16 * cb);
                    asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                                 : "=r"(a[h][0]), "=r"(a[h][1]), "=r"(a[h][2]), "=r"(a[h][3]), "=r"(a[h][4]), "=r"(a[h][5]),
                                   "=r"(a[h][6]), "=r"(a[h][7])
                                 : "r"(ad));
                    asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                                 : "=r"(l[h][0]), "=r"(l[h][1]), "=r"(l[h][2]), "=r"(l[h][3]), "=r"(l[h][4]), "=r"(l[h][5]),
                                   "=r"(l[h][6]), "=r"(l[h][7])
                                 : "r"(ad + (uint32_t)UN));
                }
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                tmem_zero16(taddr + (uint32_t)(16 * cb));       // read done: give the 16 + 16 columns back zeroed
                tmem_zero16(taddr + (uint32_t)(UN + 16 * cb));
#pragma unroll
                for (int blk = 0; blk < 2; ++blk) {
                    if (16 * cb + 8 * blk < H2P) {              // blocks 0..12 hold the 100 neurons (+4 zero-weight pads)
                        const int c = 16 * cb + 8 * blk + 2 * t0;
                        const float2 bias = *reinterpret_cast<const float2 *>(&S.b2[c]);
                        float2 w[OUT];
#pragma unroll
                        for (int o = 0; o < OUT; ++o) w[o] = *reinterpret_cast<const float2 *>(&S.w3[o][c]);
#pragma unroll
                        for (int h = 0; h < 2; ++h)
#pragma unroll
                            for (int rr = 0; rr < 2; ++rr) {
                                const int i0 = 4 * blk + 2 * rr;
                                const float h0 = fmaxf(__uint_as_float(a[h][i0]) + __uint_as_float(l[h][i0]) + bias.x, 0.f);
                                const float h1 = fmaxf(__uint_as_float(a[h][i0 + 1]) + __uint_as_float(l[h][i0 + 1]) + bias.y, 0.f);
#pragma unroll
                                for (int o = 0; o < OUT; ++o) {
                                    q[2 * h + rr][o] = fmaf(h0, w[o].x, q[2 * h + rr][o]);
                                    q[2 * h + rr][o] = fmaf(h1, w[o].y, q[2 * h + rr][o]);
                                }
                            }
                    }
                }
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            mbar_arrive(&S.tmem_empty[buf]);                    // this thread has read and zeroed its part of the buffer
            if (q4 == 0) MG_TRACE(g_trace_epi, tl, 2);
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
            const int64_t e = tile * TM + q4 * 32 + t1 + 8 * t0;
            int best = 0;
            float bv = mine[0];
#pragma unroll
            for (int o = 1; o < OUT; ++o)
                if (mine[o] > bv) { bv = mine[o]; best = o; }       // first maximum, like torch.max
            if (e < n) {
                if (!ENV) act[e] = (uint8_t)best;
                if (!ENV && (obs_mode & 0x100)) const_cast<float *>(obs)[e * (MG_OBS_DIM + 1)] = (float)best;   // MG_MLP_FLAG_WRITE_GOAL
                if (q_out) {
#pragma unroll
                    for (int o = 0; o < OUT; ++o) q_out[e * OUT + o] = mine[o];
                }
            }
            if (ENV) {                                          // hand the action to the env warp
                uint8_t *tile_act = mgpe::handoff_acquire(S.env, tl);
                tile_act[q4 * 32 + t1 + 8 * t0] = (uint8_t)best;
                mgpe::handoff_publish(S.env, tl);
            }
        }
    }
    // ---- teardown -------------------------------------------------------------------------------------
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 12)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS));
}

template <int IN, int OUT, bool MIRROR, int ENV = 0>
cudaError_t launch(const float *obs, const uint8_t *goal, int64_t n, int obs_mode, const float *w1t, const float *b1,
                   const float *w2_tc, const float *b2, const float *w3, const float *b3, uint8_t *act, float *q_out,
                   cudaStream_t st, const mgpe::Args &P = mgpe::Args{}, bool pdl = false) {
    auto kern = mlp_act_tc_kernel<IN, OUT, MIRROR, ENV>;
    const size_t smem = sizeof(Smem<IN, OUT>) + 1024;           // slack for the 1024-byte alignment of the base
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e) return e;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int64_t tiles = (n + TM - 1) / TM;
    const unsigned grid = (unsigned)(tiles < sms ? tiles : sms);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(ENV ? NUM_THREADS_ENV : NUM_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
    e = cudaLaunchKernelEx(&cfg, kern, obs, goal, n, obs_mode, w1t, b1, w2_tc, b2, w3, b3, act, q_out, P);
    return e ? e : cudaGetLastError();
}

}  // namespace mgtc

int mg_mlp_check_layout(uint32_t flags, int32_t obs_dim, int32_t out_dim, bool has_goal, int *in_dim, int *obs_mode);
// mlp_tc16_kernels.cu: MG_MLP_FLAG_F16X3
cudaError_t mg_mlp_act_tc16_launch(int in_dim, int out_dim, bool mirror, const float *obs, const uint8_t *goal, int64_t n,
                                   int obs_mode, const void *blob, const float *b2, const float *w3, const float *b3,
                                   uint8_t *act, float *q_out, cudaStream_t st, bool pdl);

// the fused policy + env step on the tensor-core backend (called by mg_policy_step in mlp_kernels.cu)
cudaError_t mg_policy_step_tc_launch(int in_dim, const float *obs, const uint8_t *goal, int64_t n, const float *w1t,
                                     const float *b1, const float *w2_tc, const float *b2, const float *w3, const float *b3,
                                     float *q_out, cudaStream_t st, const mgpe::Args &P) {
    const bool pvp = P.a2 != nullptr, pdl = (P.flags & MG_POLICY_FLAG_PDL) != 0u;
    const int obs_mode = (int)mg::obs_layout_of(P.flags);
#define MG_TC_ENV(I) (pvp ? mgtc::launch<I, 5, false, 2>(obs, goal, n, obs_mode, w1t, b1, w2_tc, b2, w3, b3, nullptr, q_out, st, P, pdl) \
                          : mgtc::launch<I, 5, false, 1>(obs, goal, n, obs_mode, w1t, b1, w2_tc, b2, w3, b3, nullptr, q_out, st, P, pdl))
    if (in_dim == 10) return MG_TC_ENV(10);
    if (in_dim == 11) return MG_TC_ENV(11);
#undef MG_TC_ENV
    return cudaErrorInvalidValue;
}

extern "C" MG_API int mg_mlp_act_tc(const float *obs, const uint8_t *goal_or_null, int64_t n, int32_t obs_dim,
                                    int32_t out_dim, const float *w1t, const float *b1, const float *w2_tc,
                                    const float *b2, const float *w3, const float *b3, uint8_t *actions,
                                    float *q_out_or_null, uint32_t flags, void *stream) {
    using namespace mg_abi;
    if (n < 0) return fail(MG_ERR_BAD_SIZE, "n < 0");
    int in_dim = 0, obs_mode = 0;
    const bool f16 = (flags & MG_MLP_FLAG_F16X3) != 0u;           // w2_tc is then the packed fp16 operand blob (see the header)
    if (int rc = mg_mlp_check_layout(flags & ~MG_MLP_FLAG_F16X3, obs_dim, out_dim, goal_or_null != nullptr, &in_dim, &obs_mode)) return rc;
    const bool mirror = (flags & MG_MLP_FLAG_MIRROR) != 0u, pdl = (flags & MG_MLP_FLAG_PDL) != 0u;
    if (n == 0) return MG_OK;
    if (!obs || (!f16 && (!w1t || !b1)) || !w2_tc || !b2 || !w3 || !b3 || !actions)
        return fail(MG_ERR_NULL_POINTER, "mg_mlp_act_tc: NULL pointer");
    if (!aligned16(obs) || (!f16 && !aligned16(w1t)) || !aligned16(w2_tc))
        return fail(MG_ERR_ALIGNMENT, "obs and weight arrays must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e;
    if (f16) {
        e = mg_mlp_act_tc16_launch(in_dim, out_dim, mirror, obs, goal_or_null, n, obs_mode, w2_tc, b2, w3, b3, actions, q_out_or_null, st, pdl);
        if (e) return cuda_fail(e, "mg_mlp_act_tc (f16x3) launch");
        return MG_OK;
    }
#define MG_TC_CASE(I, O) \
    if (in_dim == I && out_dim == O) e = mirror ? mgtc::launch<I, O, true>(obs, goal_or_null, n, obs_mode, w1t, b1, w2_tc, b2, w3, b3, actions, q_out_or_null, st, mgpe::Args{}, pdl) : mgtc::launch<I, O, false>(obs, goal_or_null, n, obs_mode, w1t, b1, w2_tc, b2, w3, b3, actions, q_out_or_null, st, mgpe::Args{}, pdl); else
    MG_TC_CASE(10, 5) MG_TC_CASE(10, 3) MG_TC_CASE(11, 5) MG_TC_CASE(11, 3) e = cudaErrorInvalidValue;
#undef MG_TC_CASE
    if (e) return cuda_fail(e, "mg_mlp_act_tc launch");
    return MG_OK;
}
