// exp_mlp_tc_pair_kernels.cu — EXPERIMENT (not built into the library): the tensor-core Q-network forward
// (merging_gym_b200/csrc/mlp_tc_kernels.cu) on CTA PAIRS: tcgen05.mma
// .cta_group::2, UMMA M = 256 = two 128-env tiles, one per CTA of a 2-CTA cluster (two SMs of a TPC).
//
// Why: the single-CTA kernel is bound by shared-memory bandwidth (~320 wavefronts per K-step), 84 of which are the B
// operand (W2) that every SM re-reads for every K-step.  In a pair the B operand's N rows are SPLIT between the two
// CTAs' shared memories and each SM reads only its half (profiles/exp_tcgen05_pair_rate.cu: the M256 x N224 x K8 tf32
// MMA takes 112 cycles per pair = the tensor peak, against 123 for M128 on one SM), and W2 occupies 134 KB per SM
// instead of 179 KB, which pays for an 8-slot A ring (one slot per producer warp) instead of 4.
//
// Layout per CTA (rank r of the pair):  b1_op = this CTA's half of the stacked [W2_hi ; W2_lo] operand of the N = 224 MMA
// (r = 0: W2_hi, r = 1: W2_lo; 112 rows);  b2_op = its half of W2_hi for the N = 112 MMA (rows 56 r .. 56 r + 55).
// The same descriptors (same shared-memory offsets) are valid in both CTAs; accumulators sit at the same TMEM columns.
// Roles per CTA as in the single-CTA kernel (8 producer warps, 4 epilogue warps); ONE MMA-issuing warp in the leader
// CTA issues for the pair and commits with .multicast::cluster to the barriers of both CTAs.  Producers and epilogues
// of the peer CTA signal the leader's barriers with mbarrier.arrive.release.cluster on the mapa-translated address.
// Deterministic (one issuer, K-step order); every wait is bounded and traps instead of hanging.
//
// Status (B200, 2^18 envs; profiles/README.md): numerically identical to the single-CTA kernel from the first run.
// 82.1 us with release.cluster arrives / acquire.cluster waits (the remote release-arrive blocks each producer warp
// for ~900 cycles), 73.5 us with relaxed.cluster arrives, 69.6 us with relaxed arrives and CTA-scope waits — within 3 %
// of the single-CTA kernel on the same box (69.8 vs 71.9 us traced), not a clear win: the steady-state K-step period
// drops to 280 cycles, but every hand-over now crosses the cluster (multicast commits, remote arrives) and the tile
// switch stays expensive.
// Build: see profiles/exp_tc_pair_trace.cu (which includes this file); -DMG_PAIR_ARRIVE_SEM='".relaxed.cluster"'
// -DMG_PAIR_ACQ_CTA=1 select the faster, formally weaker synchronisation.
#include "../merging_gym_b200/csrc/abi_common.h"

#ifndef MG_PAIR_ALL_LANES
#define MG_PAIR_ALL_LANES 0
#endif
#ifndef MG_PAIR_ACQ_CTA
#define MG_PAIR_ACQ_CTA 0
#endif

#ifndef MG_PAIR_ARRIVE_SEM
#define MG_PAIR_ARRIVE_SEM ".release.cluster"
#endif
#if MG_PAIR_ACQ_CTA
#define MG_PAIR_SEM ""
#else
#define MG_PAIR_SEM ".acquire.cluster"
#endif
namespace mgtc2 {

#ifndef MG_TC_TRACE
#define MG_TC_TRACE 0
#endif
#if MG_TC_TRACE
constexpr int TRACE_G = 25 * 10;
__device__ long long g_trace_prod[2][TRACE_G][4];   // [rank]: compute start, wait start, wait end, arrive done
__device__ long long g_trace_mma[TRACE_G][3];       // wait start, wait end, issued
#define MG_TRACE2(arr, idx, k) do { if (blockIdx.x < 2 && lane == 0 && (idx) < (uint32_t)TRACE_G) arr[idx][k] = clock64(); } while (0)
#else
#define MG_TRACE2(arr, idx, k) do { } while (0)
#endif

constexpr int H1 = 200, H2 = 100;
constexpr int TM = 128;                       // envs per CTA tile; the pair's UMMA M is 256
constexpr int UN = 112;
constexpr int KSTEPS = H1 / 8;
constexpr int SLOTS = 8;                      // A ring: one slot per producer warp
constexpr int H2P = 104;
constexpr int A_STEP = (TM / 8) * 256;        // 4096 B
constexpr int B1_STEP = (UN / 8) * 256;       // 3584 B: this CTA's 112 rows of the N = 224 operand, one K-step
constexpr int B2_STEP = (UN / 16) * 256;      // 1792 B: this CTA's 56 rows of the N = 112 operand
constexpr int SRC_STEP = (2 * UN / 8) * 256;  // 7168 B: one K-step of the host-prepared [hi ; lo] operand (28 row groups)
constexpr int TMEM_COLS = 512;
constexpr int PRODUCER_WARPS = 8;
constexpr int NUM_THREADS = 416;              // 8 producer warps + 4 epilogue warps + the MMA / TMEM warp
constexpr int MAX_OUT = 8;
constexpr uint32_t kSpinLimit = 1u << 26;

template <int IN, int OUT>
struct Smem {
    unsigned char b1_op[KSTEPS * B1_STEP];    // 89 600 B
    unsigned char b2_op[KSTEPS * B2_STEP];    // 44 800 B
    unsigned char a_hi[SLOTS][A_STEP];
    unsigned char a_lo[SLOTS][A_STEP];
    float w1[IN][H1];
    float w3[OUT][H2P];
    float b1[H1], b2[H2 + 12], b3[MAX_OUT];
    // full / tmem_empty are used in the LEADER CTA only (both CTAs arrive there); empty / tmem_full exist in both
    unsigned long long full[SLOTS], empty[SLOTS], tmem_full[2], tmem_empty[2];
    uint32_t tmem_base;
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {   // no swizzle, K-major, LBO 128, SBO 256 (see mlp_tc_kernels.cu)
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(128u >> 4) << 16) | ((uint64_t)(256u >> 4) << 32) | ((uint64_t)1 << 46);
}
constexpr uint64_t kDescHi = ((uint64_t)(256u >> 4) << 32) | ((uint64_t)1 << 46);
constexpr uint32_t idesc_n(int n) { return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(256 >> 4) << 24); }   // M = 256
constexpr uint32_t kIdesc112 = idesc_n(UN), kIdesc224 = idesc_n(2 * UN);

__device__ __forceinline__ void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_init(unsigned long long *b, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count));
}
// arrive on the LEADER CTA's copy of barrier `b` (rank 0 of the pair), from either CTA
__device__ __forceinline__ void mbar_arrive_leader(unsigned long long *b) {
    uint32_t remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(b)), "r"(0u));
    asm volatile("mbarrier.arrive" MG_PAIR_ARRIVE_SEM ".shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *b, uint32_t parity) {   // acquire at cluster scope
    uint32_t done = 0;
    for (uint32_t it = 0; it < kSpinLimit && !done; ++it) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity" MG_PAIR_SEM ".shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
            : "=r"(done)
            : "r"(smem_u32(b)), "r"(parity)
            : "memory");
    }
    if (!done) __trap();
}
__device__ __forceinline__ uint32_t mbar_test(unsigned long long *b, uint32_t parity) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity" MG_PAIR_SEM ".shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(done)
        : "r"(smem_u32(b)), "r"(parity)
        : "memory");
    return done;
}
__device__ __forceinline__ void tmem_zero16(uint32_t taddr) {
    const uint32_t z = 0u;
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(taddr), "r"(z) : "memory");
}

template <int IN, bool MIRROR>
__device__ __forceinline__ void load_row(const float *__restrict__ obs, const uint8_t *__restrict__ goal, int64_t e,
                                         int64_t n, int obs_dim, float (&x)[IN]) {
    constexpr int off = IN - MG_OBS_DIM;   // 1 when a goal column is prepended (hdqn.py:291); compile-time so x[] stays in registers
    (void)obs_dim;
    if (e < n) {
        if (off) x[0] = (float)goal[e];
        if (!MIRROR) {
            const float2 *src = reinterpret_cast<const float2 *>(obs + e * MG_OBS_DIM);
#pragma unroll
            for (int i = 0; i < MG_OBS_DIM / 2; ++i) {          // obs_dim is 10: five float2
                const float2 v = __ldg(src + i);
                x[off + 2 * i] = v.x; x[off + 2 * i + 1] = v.y;
            }
        } else {                                                // the opponent's view: state[5:] + state[:5] (main.py:199)
#pragma unroll
            for (int i = 0; i < MG_OBS_DIM; ++i) x[off + i] = __ldg(obs + e * MG_OBS_DIM + (i + MG_OBS_DIM / 2) % MG_OBS_DIM);
        }
    } else {
#pragma unroll
        for (int i = 0; i < IN; ++i) x[i] = 0.f;
    }
}

template <int IN, int OUT, bool MIRROR>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
mlp_act_tc_pair_kernel(const float *__restrict__ obs, const uint8_t *__restrict__ goal, const int64_t n, const int obs_dim,
                       const float *__restrict__ w1t, const float *__restrict__ b1, const float *__restrict__ w2_tc,
                       const float *__restrict__ b2, const float *__restrict__ w3, const float *__restrict__ b3,
                       uint8_t *__restrict__ act, float *__restrict__ q_out) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    Smem<IN, OUT> &S = *reinterpret_cast<Smem<IN, OUT> *>(smem_raw);
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    uint32_t rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    const int64_t n_tiles = (n + TM - 1) / TM, n_pairs = (n_tiles + 1) / 2;
    const int64_t cid = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
    const int64_t my_pairs = cid < n_pairs ? (n_pairs - cid + n_clusters - 1) / n_clusters : 0;   // the same in both CTAs

    // ---- one-time setup ---------------------------------------------------------------------------------
    {
        const unsigned char *src = reinterpret_cast<const unsigned char *>(w2_tc);
        for (int i = t; i < KSTEPS * B1_STEP / 16; i += NUM_THREADS) {          // rank 0: W2_hi rows, rank 1: W2_lo rows
            const int ks = i / (B1_STEP / 16), o = i - ks * (B1_STEP / 16);
            reinterpret_cast<float4 *>(S.b1_op)[i] = __ldg(reinterpret_cast<const float4 *>(src + (size_t)ks * SRC_STEP + rank * B1_STEP) + o);
        }
        for (int i = t; i < KSTEPS * B2_STEP / 16; i += NUM_THREADS) {          // W2_hi rows 56 rank .. 56 rank + 55
            const int ks = i / (B2_STEP / 16), o = i - ks * (B2_STEP / 16);
            reinterpret_cast<float4 *>(S.b2_op)[i] = __ldg(reinterpret_cast<const float4 *>(src + (size_t)ks * SRC_STEP + rank * B2_STEP) + o);
        }
        const float4 *s1 = reinterpret_cast<const float4 *>(w1t);
        float4 *d1 = reinterpret_cast<float4 *>(&S.w1[0][0]);
        for (int i = t; i < IN * H1 / 4; i += NUM_THREADS) d1[i] = __ldg(s1 + i);
        for (int i = t; i < OUT * H2P; i += NUM_THREADS) {
            const int o = i / H2P, c = i - o * H2P;
            S.w3[o][c] = c < H2 ? w3[o * H2 + c] : 0.f;
        }
        for (int i = t; i < H1; i += NUM_THREADS) S.b1[i] = b1[i];
        for (int i = t; i < H2 + 12; i += NUM_THREADS) S.b2[i] = i < H2 ? b2[i] : 0.f;
        if (t < OUT) S.b3[t] = b3[t];
    }
    if (t == 0) {
        for (int s = 0; s < SLOTS; ++s) { mbar_init(&S.full[s], MG_PAIR_ALL_LANES ? 64 : 2); mbar_init(&S.empty[s], 1); }       // full: one elected arrive per CTA
        for (int b = 0; b < 2; ++b) { mbar_init(&S.tmem_full[b], 1); mbar_init(&S.tmem_empty[b], MG_PAIR_ALL_LANES ? 256 : 8); }  // tmem_empty: 4 warps x 2 CTAs
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    cluster_sync();                                              // both CTAs' barriers exist before anyone arrives remotely
    if (warp == 12) {                                            // the same warp of BOTH CTAs allocates for the pair
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&S.tmem_base)),
                     "r"((uint32_t)TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = S.tmem_base;
    if (warp >= 8 && warp < 12) {                                // every MMA accumulates: start from zero
        const uint32_t lanes = (uint32_t)((warp - 8) * 32) << 16;
        for (uint32_t c = 0; c < (uint32_t)TMEM_COLS; c += 16) tmem_zero16(tmem_base + lanes + c);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    if (warp < PRODUCER_WARPS) {
        // =================================== PRODUCERS: layer 1 (both CTAs, each for its own tile) ============
        // K-steps of the pair are numbered g = 25 * (local pair) + ks; warp w produces g = w, w + 8, ... into ring slot w.
        const uint32_t total = (uint32_t)my_pairs * KSTEPS;
        uint32_t off[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int m = lane + 32 * j;
            off[j] = (uint32_t)((m >> 3) * 256 + (m & 7) * 16);
        }
        float x[4][IN];
        const int s = warp;
        {
            const int64_t e0 = (2 * cid + rank) * TM + lane;
#pragma unroll
            for (int j = 0; j < 4; ++j) load_row<IN, MIRROR>(obs, goal, e0 + 32 * j, n, obs_dim, x[j]);
        }
        for (uint32_t g = (uint32_t)warp; g < total; g += PRODUCER_WARPS) {
            const uint32_t tl = g / KSTEPS, ks = g - tl * KSTEPS;
            MG_TRACE2(g_trace_prod[rank], g, 0);
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
            if ((g + PRODUCER_WARPS) / KSTEPS != tl) {          // last K-step of the tile: x[] is dead, fetch the next tile's rows into it
                const int64_t e0 = (2 * (cid + (int64_t)(tl + 1) * n_clusters) + rank) * TM + lane;
#pragma unroll
                for (int j = 0; j < 4; ++j) load_row<IN, MIRROR>(obs, goal, e0 + 32 * j, n, obs_dim, x[j]);
            }
            const uint32_t v = g / PRODUCER_WARPS;              // this warp's visit number = use number of its slot
            MG_TRACE2(g_trace_prod[rank], g, 1);
            mbar_wait(&S.empty[s], (v & 1u) ^ 1u);              // the pair's MMAs that read this slot have completed
            MG_TRACE2(g_trace_prod[rank], g, 2);
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
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // visible to this SM's tensor core
#if MG_PAIR_ALL_LANES
            mbar_arrive_leader(&S.full[s]);
#else
            __syncwarp();
            if (lane == 0) mbar_arrive_leader(&S.full[s]);
#endif
            MG_TRACE2(g_trace_prod[rank], g, 3);
        }
    } else if (warp == 12) {
        // =================================== MMA ISSUER (leader CTA only) =========================
        if (rank == 0) {
            uint32_t tl = 0;
            const uint32_t b1_0 = (uint32_t)make_desc(smem_u32(S.b1_op)), b2_0 = (uint32_t)make_desc(smem_u32(S.b2_op));
            const uint32_t a_hi0 = (uint32_t)make_desc(smem_u32(S.a_hi[0])), a_lo0 = (uint32_t)make_desc(smem_u32(S.a_lo[0]));
            for (int64_t p = 0; p < my_pairs; ++p, ++tl) {
                const uint32_t buf = tl & 1u;
                mbar_wait(&S.tmem_empty[buf], ((tl >> 1) & 1u) ^ 1u);            // both epilogues drained (and zeroed) this buffer
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t d = tmem_base + buf * 256u;
                const uint32_t tm_full = smem_u32(&S.tmem_full[buf]);
                uint32_t have = 0;
#pragma unroll
                for (int ks = 0; ks < KSTEPS; ++ks) {
                    const uint32_t it = tl * KSTEPS + (uint32_t)ks;
                    const uint32_t s = it % SLOTS;
                    uint32_t lo_a = a_hi0 + s * (A_STEP >> 4), lo_l = a_lo0 + s * (A_STEP >> 4);
                    uint32_t lo_b1 = b1_0 + (uint32_t)ks * (B1_STEP >> 4), lo_b2 = b2_0 + (uint32_t)ks * (B2_STEP >> 4);
                    uint32_t done_bar = smem_u32(&S.empty[s]);
                    asm volatile("" : "+r"(lo_a), "+r"(lo_l), "+r"(lo_b1), "+r"(lo_b2), "+r"(done_bar));
                    const uint32_t last = ks == KSTEPS - 1 ? 1u : 0u;
                    MG_TRACE2(g_trace_mma, it, 0);
                    if (!have) mbar_wait(&S.full[s], (it / SLOTS) & 1u);
                    have = last ? 0u : mbar_test(&S.full[(it + 1u) % SLOTS], ((it + 1u) / SLOTS) & 1u);
                    MG_TRACE2(g_trace_mma, it, 1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    // pair MMAs: [0,112) += a_hi.W2_hi and [112,224) += a_hi.W2_lo (N = 224, B split hi | lo between the CTAs),
                    // then [0,112) += a_lo.W2_hi (N = 112, B split 56 | 56); commits are multicast to both CTAs
                    asm volatile(
                        "{\n\t.reg .pred E, L;\n\t.reg .b64 da, dl, db1, db2;\n\t"
                        "elect.sync _|E, 0xffffffff;\n\t"
                        "setp.ne.and.b32 L, %8, 0, E;\n\t"
                        "mov.b64 da, {%1, %5};\n\tmov.b64 dl, {%2, %5};\n\tmov.b64 db1, {%3, %5};\n\tmov.b64 db2, {%4, %5};\n\t"
                        "@E tcgen05.mma.cta_group::2.kind::tf32 [%0], da, db1, %6, 1;\n\t"
                        "@E tcgen05.mma.cta_group::2.kind::tf32 [%0], dl, db2, %7, 1;\n\t"
                        "@E tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%9], %11;\n\t"
                        "@L tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%10], %11;\n\t}\n"
                        :: "r"(d), "r"(lo_a), "r"(lo_l), "r"(lo_b1), "r"(lo_b2), "r"((uint32_t)(kDescHi >> 32)), "r"(kIdesc224), "r"(kIdesc112),
                           "r"(last), "r"(done_bar), "r"(tm_full), "h"((unsigned short)3)
                        : "memory");
                    MG_TRACE2(g_trace_mma, it, 2);
                }
            }
        }
        __syncwarp();
    } else {
        // =================================== EPILOGUE (both CTAs, each for its own tile) ============
        const int q4 = warp - 8;
        const int t0 = lane & 3, t1 = lane >> 2;
        uint32_t tl = 0;
        for (int64_t p = 0; p < my_pairs; ++p, ++tl) {
            const int64_t tile = 2 * (cid + p * n_clusters) + rank;
            const uint32_t buf = tl & 1u;
            mbar_wait(&S.tmem_full[buf], (tl >> 1) & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t taddr = tmem_base + buf * 256u + ((uint32_t)(q4 * 32) << 16);
            float q[4][OUT];                                    // rows t1 + 8 * {0, 1, 2, 3}: partial sums over this thread's neurons
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int o = 0; o < OUT; ++o) q[r][o] = 0.f;
#pragma unroll
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
#if MG_PAIR_ALL_LANES
            mbar_arrive_leader(&S.tmem_empty[buf]);
#else
            __syncwarp();
            if (lane == 0) mbar_arrive_leader(&S.tmem_empty[buf]);    // this warp has read and zeroed its quarter of the buffer
#endif
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
            if (e < n) {
                int best = 0;
                float bv = mine[0];
#pragma unroll
                for (int o = 1; o < OUT; ++o)
                    if (mine[o] > bv) { bv = mine[o]; best = o; }   // first maximum, like torch.max
                act[e] = (uint8_t)best;
                if (q_out) {
#pragma unroll
                    for (int o = 0; o < OUT; ++o) q_out[e * OUT + o] = mine[o];
                }
            }
        }
    }
    // ---- teardown -------------------------------------------------------------------------------------
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync();                                              // the peer may still be reading / being signalled
    if (warp == 12)
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS));
}

template <int IN, int OUT, bool MIRROR>
cudaError_t launch(const float *obs, const uint8_t *goal, int64_t n, int obs_dim, const float *w1t, const float *b1,
                   const float *w2_tc, const float *b2, const float *w3, const float *b3, uint8_t *act, float *q_out,
                   cudaStream_t st) {
    auto kern = mlp_act_tc_pair_kernel<IN, OUT, MIRROR>;
    const size_t smem = sizeof(Smem<IN, OUT>) + 1024;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e) return e;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int64_t pairs = ((n + TM - 1) / TM + 1) / 2;
    const int64_t clusters = pairs < sms / 2 ? pairs : sms / 2;
    kern<<<(unsigned)(2 * clusters), NUM_THREADS, smem, st>>>(obs, goal, n, obs_dim, w1t, b1, w2_tc, b2, w3, b3, act, q_out);
    return cudaGetLastError();
}

}  // namespace mgtc2

extern "C" MG_API int mg_mlp_act_tc_pair(const float *obs, const uint8_t *goal_or_null, int64_t n, int32_t obs_dim,
                                         int32_t out_dim, const float *w1t, const float *b1, const float *w2_tc,
                                         const float *b2, const float *w3, const float *b3, uint8_t *actions,
                                         float *q_out_or_null, uint32_t flags, void *stream) {
    using namespace mg_abi;
    if (n < 0) return fail(MG_ERR_BAD_SIZE, "n < 0");
    if (flags & ~MG_MLP_FLAG_MIRROR) return fail(MG_ERR_BAD_FLAGS, "unknown flag bits");
    const bool mirror = (flags & MG_MLP_FLAG_MIRROR) != 0u;
    const int in_dim = obs_dim + (goal_or_null ? 1 : 0);
    if (obs_dim != MG_OBS_DIM || !(out_dim == 5 || out_dim == 3))
        return fail(MG_ERR_BAD_SIZE, "mg_mlp_act_tc_pair supports obs rows of 10 floats (+ optional goal) and 5 or 3 outputs");
    if (n == 0) return MG_OK;
    if (!obs || !w1t || !b1 || !w2_tc || !b2 || !w3 || !b3 || !actions)
        return fail(MG_ERR_NULL_POINTER, "mg_mlp_act_tc_pair: NULL pointer");
    if (!aligned16(obs) || !aligned16(w1t) || !aligned16(w2_tc))
        return fail(MG_ERR_ALIGNMENT, "obs and weight arrays must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e;
#define MG_TC2_CASE(I, O) \
    if (in_dim == I && out_dim == O) e = mirror ? mgtc2::launch<I, O, true>(obs, goal_or_null, n, obs_dim, w1t, b1, w2_tc, b2, w3, b3, actions, q_out_or_null, st) : mgtc2::launch<I, O, false>(obs, goal_or_null, n, obs_dim, w1t, b1, w2_tc, b2, w3, b3, actions, q_out_or_null, st); else
    MG_TC2_CASE(10, 5) MG_TC2_CASE(10, 3) MG_TC2_CASE(11, 5) MG_TC2_CASE(11, 3) e = cudaErrorInvalidValue;
#undef MG_TC2_CASE
    if (e) return cuda_fail(e, "mg_mlp_act_tc_pair launch");
    return MG_OK;
}
