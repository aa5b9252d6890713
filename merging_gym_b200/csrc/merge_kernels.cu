// merge_kernels.cu — fused CUDA kernels (sm_100a) + the C ABI of libmerging_b200.so.
//
// Data layout in HBM: env state is structure-of-arrays
//   pos1, vel1, pos2, vel2, ret1, ret2 : float64[n]     meta : uint32[n]
// one thread owns EPT consecutive envs, so every state access is a 128-bit (EPT=2) coalesced
// vector load/store.  Outputs follow the gym-shaped API: obs[n,10] f32 (40-byte rows, staged
// through shared memory per warp so the global stores are linear 128-bit), rew[n,2] f32,
// done[n] u8, info[n] u8.
//
// Algorithmic HBM traffic per env-step, pvp + uint8 actions + auto-reset:
//   read  6*8 (state) + 4 (meta) + 2 (actions)            =  54 B
//   write 6*8 + 4 (state) + 40 (obs) + 8 (rew) + 1 + 1    = 102 B        total 156 B
#include <cstdio>
#include <cstring>
#include <type_traits>

#include "abi_common.h"
#include "merge_device.cuh"

namespace mg {

// Tunables (compile-time; the defaults are what build.py ships — see profiles/ for the sweep).
#ifndef MG_BLOCK
#define MG_BLOCK 128
#endif
#ifndef MG_EPT
#define MG_EPT 2          // envs per thread in the step kernel: 2 -> 128-bit state accesses
#endif
#ifndef MG_PDL
#define MG_PDL 1          // launch the step kernel with programmatic stream serialization (PDL): +2 % (profiles/)
#endif
#ifndef MG_TMA_OBS
#define MG_TMA_OBS 0      // 1: write each warp's observation tile with one TMA bulk copy (experiment)
#endif
#ifndef MG_MIN_BLOCKS
#define MG_MIN_BLOCKS 8   // __launch_bounds__ min resident blocks per SM: caps the step kernel at 64 registers
#endif
static_assert(32 * MG_EPT <= 255, "StatAcc packs per-warp event counts into 8-bit fields (merge_device.cuh)");
constexpr int kBlock = MG_BLOCK;
constexpr int kWarps = kBlock / 32;

// ---- vector access helpers: N consecutive elements of T per thread in <=128-bit pieces ----------
template <typename T, int N>
struct alignas((sizeof(T) * N) <= 16 ? (sizeof(T) * N) : 16) Pack { T v[N]; };

template <typename T, int N>
__device__ __forceinline__ void ld_pack(const T *__restrict__ p, T (&out)[N]) {
    constexpr int PER = (16 / (int)sizeof(T)) < N ? (16 / (int)sizeof(T)) : N;
#pragma unroll
    for (int k = 0; k < N / PER; ++k) {
        const Pack<T, PER> t = *reinterpret_cast<const Pack<T, PER> *>(p + k * PER);
#pragma unroll
        for (int i = 0; i < PER; ++i) out[k * PER + i] = t.v[i];
    }
}
template <typename T, int N>
__device__ __forceinline__ void st_pack(T *__restrict__ p, const T (&in)[N]) {
    constexpr int PER = (16 / (int)sizeof(T)) < N ? (16 / (int)sizeof(T)) : N;
#pragma unroll
    for (int k = 0; k < N / PER; ++k) {
        Pack<T, PER> t;
#pragma unroll
        for (int i = 0; i < PER; ++i) t.v[i] = in[k * PER + i];
        *reinterpret_cast<Pack<T, PER> *>(p + k * PER) = t;
    }
}
// streaming (evict-first) variant for outputs that this kernel never reads back
template <int BYTES> struct RawOf;
template <> struct RawOf<1> { using type = unsigned char; };
template <> struct RawOf<2> { using type = unsigned short; };
template <> struct RawOf<4> { using type = unsigned int; };
template <> struct RawOf<8> { using type = uint2; };
template <> struct RawOf<16> { using type = uint4; };
template <typename T, int N>
__device__ __forceinline__ void st_pack_stream(T *__restrict__ p, const T (&in)[N]) {
    constexpr int PER = (16 / (int)sizeof(T)) < N ? (16 / (int)sizeof(T)) : N;
    using Raw = typename RawOf<PER * (int)sizeof(T)>::type;
#pragma unroll
    for (int k = 0; k < N / PER; ++k) {
        Pack<T, PER> t;
#pragma unroll
        for (int i = 0; i < PER; ++i) t.v[i] = in[k * PER + i];
        Raw r;
        memcpy(&r, &t, sizeof r);
        __stcs(reinterpret_cast<Raw *>(p + k * PER), r);
    }
}

// =================================================================================================
// merge_step_kernel: one MergeEnv.step() for n envs.
//   grid = ceil(n / (kBlock*EPT)), block = kBlock.  A warp owns 32*EPT consecutive envs.
//   Full warps take the vector path; the (at most one) ragged warp takes the scalar path.
// =================================================================================================
#ifndef MG_STEP_PERSISTENT
#define MG_STEP_PERSISTENT 0   // 1: grid = SMs x MG_MIN_BLOCKS blocks that loop over the 256-env tiles (experiment, profiles/r02_variant_sweep.md)
#endif
// LAYOUT: 0 = the default obs[n][10] rows (staged per warp), 1 = MG_FLAG_OBS_SOA, 2 = MG_FLAG_OBS_GOAL_SLOT.
template <int EPT, typename ActT, bool PVP, bool RR, bool RET, int LAYOUT = 0>
__device__ __forceinline__ void step_block(const int64_t blk, float (*stage)[32 * EPT * MG_OBS_DIM], const MgState &s, const MgOut &o,
                                           const ActT *__restrict__ a1g, const ActT *__restrict__ a2g, const int64_t n,
                                           const MgRewards &rw, const uint32_t flags, const MgResetSpec &rs,
                                           unsigned long long *__restrict__ stats) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t warp_base = (blk * kWarps + warp) * (32 * EPT);
    if (warp_base >= n) return;
    const bool full = warp_base + 32 * EPT <= n;
    const int64_t e0 = warp_base + (int64_t)lane * EPT;
    const bool auto_reset = (flags & MG_FLAG_AUTO_RESET) != 0u;

    EnvRegs env[EPT];
    int act1[EPT], act2[EPT];
    bool bad[EPT];
    bool valid[EPT];

    // ---------------- loads ----------------
    if (full) {
        double p1[EPT], v1[EPT], p2[EPT], v2[EPT], R1[EPT], R2[EPT];
        uint32_t m[EPT];
        ld_pack<double, EPT>(s.pos1 + e0, p1); ld_pack<double, EPT>(s.vel1 + e0, v1);
        ld_pack<double, EPT>(s.pos2 + e0, p2); ld_pack<double, EPT>(s.vel2 + e0, v2);
        if (RET) { ld_pack<double, EPT>(s.ret1 + e0, R1); ld_pack<double, EPT>(s.ret2 + e0, R2); }
        else {
#pragma unroll
            for (int j = 0; j < EPT; ++j) R1[j] = R2[j] = 0.0;
        }
        ld_pack<uint32_t, EPT>(s.meta + e0, m);
#pragma unroll
        for (int j = 0; j < EPT; ++j) {
            env[j] = EnvRegs{p1[j], v1[j], p2[j], v2[j], R1[j], R2[j], m[j]};
            valid[j] = true;
            bad[j] = false;
            act1[j] = clamp_action((long long)load_action(a1g, e0 + j), bad[j]);
            act2[j] = PVP ? clamp_action((long long)load_action(a2g, e0 + j), bad[j]) : 0;
        }
    } else {
#pragma unroll
        for (int j = 0; j < EPT; ++j) {
            const int64_t e = e0 + j;
            valid[j] = e < n;
            bad[j] = false;
            if (valid[j]) {
                env[j] = EnvRegs{s.pos1[e], s.vel1[e], s.pos2[e], s.vel2[e], RET ? s.ret1[e] : 0.0, RET ? s.ret2[e] : 0.0, s.meta[e]};
                act1[j] = clamp_action((long long)a1g[e], bad[j]);
                act2[j] = PVP ? clamp_action((long long)a2g[e], bad[j]) : 0;
            } else {
                reset_regs(env[j]);
                act1[j] = act2[j] = 2;
            }
        }
    }

    // ---------------- compute + outputs ----------------
    StatAcc st;
    float rew[2 * EPT];
    uint8_t done8[EPT], info8[EPT];
    float *my_stage = &stage[warp][lane * EPT * MG_OBS_DIM];

    StepResult res[EPT];
    env_step_batch<PVP, EPT, RET>(env, act1, act2, bad, rw, res);
#pragma unroll
    for (int j = 0; j < EPT; ++j) {
        StepResult &r = res[j];
        rew[2 * j] = r.r1; rew[2 * j + 1] = r.r2;
        done8[j] = r.done ? 1 : 0;
        info8[j] = (uint8_t)r.info;
        if (valid[j]) {
            if (stats) st.add(r, env[j].R1, env[j].R2);
            if (r.finished) write_episode_outputs(o, e0 + j, r, env[j].R1, env[j].R2);
        }
        if (r.done && auto_reset)            // gym-0.20 vector convention: return the reset obs
            reset_env<RR>(env[j], rs, (uint64_t)(e0 + j), r.obs);
        if (LAYOUT != 0) {
            if (LAYOUT == 1 && full && EPT == 2) {
                if (j == 1) {                       // the thread's two envs are neighbours in every column: one 64-bit store
                    const int64_t stride = MG_OBS_SOA_STRIDE(n);
#pragma unroll
                    for (int k = 0; k < MG_OBS_DIM; ++k)
                        __stcs(reinterpret_cast<float2 *>(o.obs + k * stride + e0), make_float2(res[0].obs[k], r.obs[k]));
                }
            } else if (valid[j]) {
                store_obs(o.obs, (uint32_t)LAYOUT, e0 + j, n, r.obs);
            }
        } else if (full) {
#pragma unroll
            for (int k = 0; k < MG_OBS_DIM; ++k) my_stage[j * MG_OBS_DIM + k] = r.obs[k];
        } else if (valid[j]) {
            float *row = o.obs + (e0 + j) * MG_OBS_DIM;
#pragma unroll
            for (int k = 0; k < MG_OBS_DIM; ++k) row[k] = r.obs[k];
        }
    }

    // ---------------- stores ----------------
    if (full) {
        double t[EPT];
#define MG_ST(field, arr)                                   \
        _Pragma("unroll") for (int j = 0; j < EPT; ++j) t[j] = env[j].field; \
        st_pack<double, EPT>(arr + e0, t);
        MG_ST(p1, s.pos1) MG_ST(v1, s.vel1) MG_ST(p2, s.pos2) MG_ST(v2, s.vel2)
        if (RET) { MG_ST(R1, s.ret1) MG_ST(R2, s.ret2) }
#undef MG_ST
        uint32_t m[EPT];
#pragma unroll
        for (int j = 0; j < EPT; ++j) m[j] = env[j].meta;
        st_pack<uint32_t, EPT>(s.meta + e0, m);
        st_pack_stream<float, 2 * EPT>(o.rew + 2 * e0, rew);
        if (o.done) st_pack_stream<uint8_t, EPT>(o.done + e0, done8);     // NULL: the caller reads MG_INFO_DONE instead
        st_pack_stream<uint8_t, EPT>(o.info + e0, info8);
        // obs rows of this warp are one contiguous, 16-byte aligned span of 32*EPT*40 bytes
#if MG_TMA_OBS
        // One TMA bulk copy (shared -> global, 2560 B) per warp instead of 5 LDS.128 + 5 STG.128 per lane.
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy STS -> async proxy
        __syncwarp();
        if (lane == 0) {
            const uint32_t saddr = (uint32_t)__cvta_generic_to_shared(&stage[warp][0]);
            float *gdst = o.obs + warp_base * MG_OBS_DIM;
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                         :: "l"(gdst), "r"(saddr), "n"(32 * EPT * MG_OBS_DIM * 4) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
#else
        if (LAYOUT == 0) {
            __syncwarp();
            const float4 *src = reinterpret_cast<const float4 *>(&stage[warp][0]);
            float4 *dst = reinterpret_cast<float4 *>(o.obs + warp_base * MG_OBS_DIM);
            constexpr int kVec = 32 * EPT * MG_OBS_DIM / 4;   // float4 per warp
#pragma unroll
            for (int k = 0; k < (kVec + 31) / 32; ++k)
                if (kVec % 32 == 0 || lane + 32 * k < kVec) __stcs(dst + lane + 32 * k, src[lane + 32 * k]);
        }
#endif
    } else {
#pragma unroll
        for (int j = 0; j < EPT; ++j) {
            const int64_t e = e0 + j;
            if (!valid[j]) continue;
            s.pos1[e] = env[j].p1; s.vel1[e] = env[j].v1; s.pos2[e] = env[j].p2; s.vel2[e] = env[j].v2;
            if (RET) { s.ret1[e] = env[j].R1; s.ret2[e] = env[j].R2; }
            s.meta[e] = env[j].meta;
            o.rew[2 * e] = rew[2 * j]; o.rew[2 * e + 1] = rew[2 * j + 1];
            if (o.done) o.done[e] = done8[j];
            o.info[e] = info8[j];
        }
    }

    if (stats) flush_stats(st, stats + (size_t)(blk % MG_STATS_ROWS) * MG_STATS_COLS, 0xFFFFFFFFu, lane);
#if MG_TMA_OBS
    // the staging tile must stay intact until the bulk copy has read it
    if (full && lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
#endif
}

template <int EPT, typename ActT, bool PVP, bool RR, bool RET, int LAYOUT = 0>
__global__ void __launch_bounds__(kBlock, MG_MIN_BLOCKS)
merge_step_kernel(const MgState s, const MgOut o, const ActT *__restrict__ a1g,
                  const ActT *__restrict__ a2g, const int64_t n, const MgRewards rw,
                  const uint32_t flags, const MgResetSpec rs, unsigned long long *__restrict__ stats) {
    __shared__ __align__(16) float stage[kWarps][32 * EPT * MG_OBS_DIM];
#if MG_PDL
    // Programmatic dependent launch: this grid may be scheduled while the previous kernel of the
    // stream drains; nothing of global memory is touched before the previous grid has completed
    // and flushed.  Triggering right away lets the NEXT launch pre-stage the same way.
    cudaGridDependencySynchronize();
    cudaTriggerProgrammaticLaunchCompletion();
#endif
#if MG_STEP_PERSISTENT
    const int64_t n_blocks = (n + (int64_t)kBlock * EPT - 1) / ((int64_t)kBlock * EPT);
    for (int64_t blk = blockIdx.x; blk < n_blocks; blk += gridDim.x) {
        step_block<EPT, ActT, PVP, RR, RET, LAYOUT>(blk, stage, s, o, a1g, a2g, n, rw, flags, rs, stats);
        __syncwarp();                                  // the warp's staging tile is reused by the next trip
    }
#else
    step_block<EPT, ActT, PVP, RR, RET, LAYOUT>((int64_t)blockIdx.x, stage, s, o, a1g, a2g, n, rw, flags, rs, stats);
#endif
}

// =================================================================================================
// merge_rollout_kernel: k consecutive steps with in-kernel Philox actions; the envs stay in registers
// (bookkeeping unpacked: steps / winner / done / reset count) for the whole launch, outputs are time-major.
// One thread owns 2 consecutive envs.  Two loop bodies:
//   FAST  — a full warp (64 envs), n even (every time-major row 16-byte aligned) and all four outputs
//           obs | rew | done | info requested (or none at all): no per-env validity tests, no per-output
//           pointer tests, running output pointers (no 64-bit index arithmetic per step), one 128-bit
//           reward store per thread, Philox round keys from the parameter bank, statistics flushed every
//           32 steps instead of every step.
//   SLOW  — the ragged last warp, odd n, or a partial output selection / the action log: the general code.
// The round-2 profile (profiles/r02_rollout_*.md) showed the kernel at 70 % of its issue slots with
// 764 warp-instructions per 2 env-steps, only 238 of them float64 arithmetic or conversions; FAST is the diet.
// =================================================================================================
#ifndef MG_ROLLOUT_MIN_BLOCKS
#define MG_ROLLOUT_MIN_BLOCKS 4
#endif
struct PhiloxKeys { uint32_t k0[10], k1[10]; };          // key schedule of Philox4x32-10: k + r * Weyl constant
struct RollEnv {                                         // two envs of one thread
    double p1[2], v1[2], p2[2], v2[2], R1[2], R2[2];
    uint32_t steps[2], winner[2], resets[2];
    bool sticky[2];
};

__device__ __forceinline__ void philox_actions_keyed(const PhiloxKeys &K, uint64_t env_id, uint32_t s_lo, uint32_t s_hi,
                                                     int &a1, int &a2) {
    uint32_t c0 = (uint32_t)env_id, c1 = (uint32_t)(env_id >> 32), c2 = s_lo, c3 = s_hi;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ K.k0[r], n2 = hi0 ^ c3 ^ K.k1[r];
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    }
    a1 = (int)__umulhi(c0, 5u);
    a2 = (int)__umulhi(c1, 5u);
}

template <bool RR>
__device__ __forceinline__ void roll_reset(RollEnv &e, int j, const MgResetSpec &rs, uint64_t local_id, float *obs) {
    if (RR) {
        EnvRegs t;
        random_start(t, rs.seed, rs.env_id_base + local_id, e.resets[j]);
        e.p1[j] = t.p1; e.v1[j] = t.v1; e.p2[j] = t.p2; e.v2[j] = t.v2;
        observe(t, obs);
    } else {
        e.p1[j] = kStart; e.v1[j] = kInitVel; e.p2[j] = kStart; e.v2[j] = kInitVel;
        reset_obs_fixed(obs);
    }
    e.R1[j] = 0.0; e.R2[j] = 0.0;
    e.steps[j] = 0u; e.winner[j] = 0u; e.sticky[j] = false;
    e.resets[j] += 1u;
}

__device__ __forceinline__ void roll_load(const MgState &s, int64_t e, bool valid, bool ret, RollEnv &env, int j) {
    uint32_t m = 0u;
    if (valid) {
        env.p1[j] = s.pos1[e]; env.v1[j] = s.vel1[e]; env.p2[j] = s.pos2[e]; env.v2[j] = s.vel2[e];
        env.R1[j] = ret ? s.ret1[e] : 0.0; env.R2[j] = ret ? s.ret2[e] : 0.0;
        m = s.meta[e];
    } else {
        env.p1[j] = kStart; env.v1[j] = kInitVel; env.p2[j] = kStart; env.v2[j] = kInitVel;
        env.R1[j] = env.R2[j] = 0.0;
    }
    env.steps[j] = m & MG_META_STEPS_MASK;
    env.winner[j] = (m >> MG_META_WINNER_SHIFT) & 3u;
    env.sticky[j] = (m & MG_META_DONE) != 0u;
    env.resets[j] = m >> MG_META_RESETS_SHIFT;
}
__device__ __forceinline__ void roll_store(const MgState &s, int64_t e, bool ret, const RollEnv &env, int j) {
    s.pos1[e] = env.p1[j]; s.vel1[e] = env.v1[j]; s.pos2[e] = env.p2[j]; s.vel2[e] = env.v2[j];
    if (ret) { s.ret1[e] = env.R1[j]; s.ret2[e] = env.R2[j]; }
    s.meta[e] = env.steps[j] | (env.winner[j] << MG_META_WINNER_SHIFT) | (env.sticky[j] ? MG_META_DONE : 0u) |
                (env.resets[j] << MG_META_RESETS_SHIFT);
}

// FAST: whole warps only (the host launches it on the first n_fast = 64 * floor(n / 64) envs), auto-reset with the fixed
// start, outputs either all four (OUT, n even) or none.
#ifndef MG_ROLLOUT_FAST_MIN_BLOCKS
#define MG_ROLLOUT_FAST_MIN_BLOCKS 6   // 80 registers (40 B spilled): 11.9 us per 2^20-env step against 12.15 at 4 blocks / 120 registers, 12.0 at 5, 14.0 at 8
#endif
template <bool PVP, bool RET, bool OUT>
__global__ void __launch_bounds__(kBlock, MG_ROLLOUT_FAST_MIN_BLOCKS)
merge_rollout_fast_kernel(const MgState s, const MgOut o, const int64_t n, const int64_t n_fast, const PhiloxKeys keys,
                          const uint64_t env_id_base, const uint64_t step0, const int k_steps, const MgRewards rw,
                          const MgResetSpec rs, unsigned long long *__restrict__ stats) {
    constexpr int EPT = 2;
    __shared__ __align__(16) float stage[OUT ? kWarps : 1][32 * EPT * MG_OBS_DIM];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t warp_base = ((int64_t)blockIdx.x * kWarps + warp) * (32 * EPT);
    if (warp_base >= n_fast) return;
    const int64_t e0 = warp_base + (int64_t)lane * EPT;
    RollEnv env;
#pragma unroll
    for (int j = 0; j < EPT; ++j) roll_load(s, e0 + j, true, RET, env, j);
    unsigned long long *stats_row = stats ? stats + (size_t)(blockIdx.x % MG_STATS_ROWS) * MG_STATS_COLS : nullptr;
    const uint64_t gid0 = env_id_base + (uint64_t)e0;
    const bool episode_out = o.term_obs || o.ep_ret || o.ep_len;

    float4 *obs_p = nullptr, *rew_p = nullptr;
    unsigned short *done_p = nullptr, *info_p = nullptr;
    float *my_stage = nullptr;
    if (OUT) {
        obs_p = reinterpret_cast<float4 *>(o.obs + warp_base * MG_OBS_DIM) + lane;
        rew_p = reinterpret_cast<float4 *>(o.rew + 2 * e0);
        done_p = reinterpret_cast<unsigned short *>(o.done + e0);
        info_p = reinterpret_cast<unsigned short *>(o.info + e0);
        my_stage = &stage[warp][lane * EPT * MG_OBS_DIM];
    }
    const int64_t obs_stride = n * MG_OBS_DIM / 4, half_n = n / 2;    // per step, in float4 / 2-env elements
    StatAcc st;
    uint64_t step = step0;
    for (int t = 0; t < k_steps; ++t, ++step) {
        int act1[EPT], act2[EPT];
#pragma unroll
        for (int j = 0; j < EPT; ++j)
            philox_actions_keyed(keys, gid0 + (uint64_t)j, (uint32_t)step, (uint32_t)(step >> 32), act1[j], act2[j]);
        float obs[EPT][MG_OBS_DIM], r1[EPT], r2[EPT];
        StepFlags fl[EPT];
        env_step_core<PVP, EPT, RET>(env.p1, env.v1, env.p2, env.v2, env.R1, env.R2, env.steps, env.winner,
                                     env.sticky, act1, act2, rw, obs, r1, r2, fl);
        uint32_t info[EPT];
#pragma unroll
        for (int j = 0; j < EPT; ++j) {
            info[j] = info_byte(fl[j], env.winner[j], false);
            if (fl[j].done) {                       // rare: ~0.5 % of envs per step
                StepResult r;
                r.info = info[j]; r.steps = env.steps[j]; r.finished = true; r.done = true;
                if (stats) st.add(r, env.R1[j], env.R2[j]);
                if (episode_out) {
#pragma unroll
                    for (int k = 0; k < MG_OBS_DIM; ++k) r.obs[k] = obs[j][k];
                    write_episode_outputs(o, e0 + j, r, env.R1[j], env.R2[j]);
                }
                roll_reset<false>(env, j, rs, (uint64_t)(e0 + j), obs[j]);   // gym-0.20: the step returns the reset obs
            }
        }
        if (OUT) {
#pragma unroll
            for (int j = 0; j < EPT; ++j)
#pragma unroll
                for (int k = 0; k < MG_OBS_DIM; ++k) my_stage[j * MG_OBS_DIM + k] = obs[j][k];
            __stcs(rew_p, make_float4(r1[0], r2[0], r1[1], r2[1]));
            __stcs(done_p, (unsigned short)((fl[0].done ? 1u : 0u) | (fl[1].done ? 0x100u : 0u)));
            __stcs(info_p, (unsigned short)(info[0] | (info[1] << 8)));
            __syncwarp();
            const float4 *src = reinterpret_cast<const float4 *>(&stage[warp][0]) + lane;
#pragma unroll
            for (int k = 0; k < EPT * MG_OBS_DIM / 4; ++k) __stcs(obs_p + 32 * k, src[32 * k]);
            __syncwarp();
            obs_p += obs_stride; rew_p += half_n; done_p += half_n; info_p += half_n;
        }
        if (stats && ((t & 31) == 31)) { flush_stats(st, stats_row, 0xFFFFFFFFu, lane); st = StatAcc(); }
    }
    if (stats) flush_stats(st, stats_row, 0xFFFFFFFFu, lane);
#pragma unroll
    for (int j = 0; j < EPT; ++j) roll_store(s, e0 + j, RET, env, j);
}

// SLOW: envs [e_begin, n) — the ragged tail behind the FAST launch, or everything when FAST does not apply.
template <bool PVP, bool RR, bool RET>
__global__ void __launch_bounds__(kBlock, MG_ROLLOUT_MIN_BLOCKS)
merge_rollout_kernel(const MgState s, const MgOut o, uint8_t *__restrict__ actions_out, const int64_t n,
                     const int64_t e_begin, const PhiloxKeys keys, const uint64_t env_id_base, const uint64_t step0,
                     const int k_steps, const MgRewards rw, const uint32_t flags, const MgResetSpec rs,
                     unsigned long long *__restrict__ stats) {
    constexpr int EPT = 2;
    __shared__ __align__(16) float stage[kWarps][32 * EPT * MG_OBS_DIM];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t warp_base = e_begin + ((int64_t)blockIdx.x * kWarps + warp) * (32 * EPT);
    if (warp_base >= n) return;
    const int64_t e0 = warp_base + (int64_t)lane * EPT;
    const bool auto_reset = (flags & MG_FLAG_AUTO_RESET) != 0u;
    // time-major obs rows start at t*n*40 bytes: 16-byte aligned for every t only if n is even
    const bool full = (warp_base + 32 * EPT <= n) && ((n & 1) == 0) && ((warp_base & 1) == 0);

    RollEnv env;
    bool valid[EPT];
#pragma unroll
    for (int j = 0; j < EPT; ++j) {
        valid[j] = e0 + j < n;
        roll_load(s, e0 + j, valid[j], RET, env, j);
    }
    unsigned long long *stats_row = stats ? stats + (size_t)(blockIdx.x % MG_STATS_ROWS) * MG_STATS_COLS : nullptr;
    float *my_stage = &stage[warp][lane * EPT * MG_OBS_DIM];
    const uint64_t gid0 = env_id_base + (uint64_t)e0;

    for (int t = 0; t < k_steps; ++t) {
        StatAcc st;
        const int64_t toff = (int64_t)t * n;
        const uint64_t step = step0 + (uint64_t)t;
        int act1[EPT], act2[EPT];
#pragma unroll
        for (int j = 0; j < EPT; ++j)
            philox_actions_keyed(keys, gid0 + (uint64_t)j, (uint32_t)step, (uint32_t)(step >> 32), act1[j], act2[j]);
        float obs[EPT][MG_OBS_DIM], r1[EPT], r2[EPT];
        StepFlags fl[EPT];
        env_step_core<PVP, EPT, RET>(env.p1, env.v1, env.p2, env.v2, env.R1, env.R2, env.steps, env.winner,
                                     env.sticky, act1, act2, rw, obs, r1, r2, fl);
#pragma unroll
        for (int j = 0; j < EPT; ++j) {
            const int64_t e = e0 + j;
            const uint32_t info = info_byte(fl[j], env.winner[j], false);
            if (valid[j] && (fl[j].finished || stats)) {
                StepResult r;
                r.info = info; r.steps = env.steps[j]; r.finished = fl[j].finished; r.done = fl[j].done;
#pragma unroll
                for (int k = 0; k < MG_OBS_DIM; ++k) r.obs[k] = obs[j][k];
                if (stats) st.add(r, env.R1[j], env.R2[j]);
                if (fl[j].finished) write_episode_outputs(o, e, r, env.R1[j], env.R2[j]);
            }
            if (fl[j].done && auto_reset) roll_reset<RR>(env, j, rs, (uint64_t)e, obs[j]);
            if (valid[j]) {
                if (o.rew) __stcs(reinterpret_cast<float2 *>(o.rew + 2 * (toff + e)), make_float2(r1[j], r2[j]));
                if (o.done) __stcs(o.done + toff + e, (uint8_t)(fl[j].done ? 1 : 0));
                if (o.info) __stcs(o.info + toff + e, (uint8_t)info);
                if (actions_out) __stcs(reinterpret_cast<uchar2 *>(actions_out + 2 * (toff + e)),
                                        make_uchar2((uint8_t)act1[j], (uint8_t)(PVP ? act2[j] : 0)));
            }
            if (o.obs) {
                if (full) {
#pragma unroll
                    for (int k = 0; k < MG_OBS_DIM; ++k) my_stage[j * MG_OBS_DIM + k] = obs[j][k];
                } else if (valid[j]) {
                    float *row = o.obs + (toff + e) * MG_OBS_DIM;
#pragma unroll
                    for (int k = 0; k < MG_OBS_DIM; ++k) row[k] = obs[j][k];
                }
            }
        }
        if (o.obs && full) {
            __syncwarp();
            const float4 *src = reinterpret_cast<const float4 *>(&stage[warp][0]);
            float4 *dst = reinterpret_cast<float4 *>(o.obs + (toff + warp_base) * MG_OBS_DIM);
#pragma unroll
            for (int k = 0; k < EPT * MG_OBS_DIM / 4; ++k) __stcs(dst + lane + 32 * k, src[lane + 32 * k]);
            __syncwarp();
        }
        if (stats) flush_stats(st, stats_row, 0xFFFFFFFFu, lane);
    }
#pragma unroll
    for (int j = 0; j < EPT; ++j)
        if (valid[j]) roll_store(s, e0 + j, RET, env, j);
}

// =================================================================================================
// merge_reset_kernel: MergeEnv.reset() (+ observe()) — one env per thread.
// =================================================================================================
__global__ void __launch_bounds__(kBlock)
merge_reset_kernel(const MgState s, const int64_t n, const uint8_t *__restrict__ mask, float *__restrict__ obs,
                   const MgResetSpec rs, const uint32_t layout) {
    __shared__ __align__(16) float stage[kWarps][32 * MG_OBS_DIM];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t warp_base = ((int64_t)blockIdx.x * kWarps + warp) * 32;
    const int64_t e = warp_base + lane;
    const bool valid = e < n;
    const bool hit = valid && (!mask || mask[e]);
    EnvRegs r;
    reset_regs(r);
    if (valid) r.meta = s.meta[e];                      // the reset count in it survives a reset
    float ob[MG_OBS_DIM];
    if (hit) {
        if (rs.mode == MG_RESET_RANDOM) reset_env<true>(r, rs, (uint64_t)e, ob);
        else reset_env<false>(r, rs, (uint64_t)e, ob);
        s.pos1[e] = r.p1; s.vel1[e] = r.v1; s.pos2[e] = r.p2; s.vel2[e] = r.v2;
        if (s.ret1) { s.ret1[e] = 0.0; s.ret2[e] = 0.0; }      // NULL: the env keeps no return accumulators
        s.meta[e] = r.meta;
    } else {
        if (valid) { r.p1 = s.pos1[e]; r.v1 = s.vel1[e]; r.p2 = s.pos2[e]; r.v2 = s.vel2[e]; }
        observe(r, ob);
    }
    if (!obs) return;
    if (layout != kObsAos) {
        if (valid) store_obs(obs, layout, e, n, ob);
        return;
    }
    if (warp_base + 32 <= n) {
        // the warp's 32 rows are one contiguous 1280-byte span: stage, then linear 128-bit stores
#pragma unroll
        for (int k = 0; k < MG_OBS_DIM; ++k) stage[warp][lane * MG_OBS_DIM + k] = ob[k];
        __syncwarp();
        const float4 *src = reinterpret_cast<const float4 *>(&stage[warp][0]);
        float4 *dst = reinterpret_cast<float4 *>(obs + warp_base * MG_OBS_DIM);
        constexpr int kVec = 32 * MG_OBS_DIM / 4;
#pragma unroll
        for (int k = 0; k < (kVec + 31) / 32; ++k) {
            const int i = lane + 32 * k;
            if (i < kVec) dst[i] = src[i];
        }
    } else if (valid) {
#pragma unroll
        for (int k = 0; k < MG_OBS_DIM; ++k) obs[e * MG_OBS_DIM + k] = ob[k];
    }
}

__global__ void __launch_bounds__(kBlock)
sample_actions_kernel(uint8_t *__restrict__ a1, uint8_t *__restrict__ a2, const int64_t n,
                      const uint64_t seed, const uint64_t env_id_base, const uint64_t step) {
    const int64_t e = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    if (e >= n) return;
    int x1, x2;
    philox_actions(seed, env_id_base + (uint64_t)e, step, x1, x2);
    a1[e] = (uint8_t)x1;
    if (a2) a2[e] = (uint8_t)x2;
}

}  // namespace mg

// =================================================================================================
// C ABI
// =================================================================================================
namespace {
using mg_abi::aligned16;
using mg_abi::cuda_fail;
using mg_abi::fail;

int check_state(const MgState *s, bool returns = true) {
    if (!s) return fail(MG_ERR_NULL_POINTER, "state is NULL");
    const void *ptrs[] = {s->pos1, s->vel1, s->pos2, s->vel2, s->meta, s->ret1, s->ret2};
    for (int i = 0; i < (returns ? 7 : 5); ++i) {
        const void *p = ptrs[i];
        if (!p) return fail(MG_ERR_NULL_POINTER, "a state array pointer is NULL");
        if (!aligned16(p)) return fail(MG_ERR_ALIGNMENT, "state arrays must be 16-byte aligned");
    }
    return MG_OK;
}
int check_out(const MgOut *o, bool all_required) {
    if (!o) return fail(MG_ERR_NULL_POINTER, "out is NULL");
    if (all_required && (!o->obs || !o->rew || !o->info))
        return fail(MG_ERR_NULL_POINTER, "out.obs/rew/info must be non-NULL");
    const void *ptrs[] = {o->obs, o->rew, o->done, o->info, o->term_obs, o->ep_ret, o->ep_len};
    for (const void *p : ptrs)
        if (p && !aligned16(p)) return fail(MG_ERR_ALIGNMENT, "output arrays must be 16-byte aligned");
    return MG_OK;
}
const MgRewards kDefaultRewards = {2.0, 1.0, -10.0, 0.001, 0.0};
// true when obs < rew < done < info follow each other (gaps < 4 KB in total) with the same byte offsets in the
// device and the host set, and no optional per-episode array is requested on the host side
bool outputs_joined(const MgOut &d, const MgOut &h, size_t m) {
    if (!h.obs || !h.rew || !h.done || !h.info || h.term_obs || h.ep_ret || h.ep_len) return false;
    const char *d0 = (const char *)d.obs, *h0 = (const char *)h.obs;
    const ptrdiff_t o1 = (const char *)d.rew - d0, o2 = (const char *)d.done - d0, o3 = (const char *)d.info - d0;
    if (o1 != (const char *)h.rew - h0 || o2 != (const char *)h.done - h0 || o3 != (const char *)h.info - h0) return false;
    if (!(0 < o1 && o1 < o2 && o2 < o3)) return false;
    const size_t need = m * (MG_OBS_DIM * sizeof(float) + 2 * sizeof(float) + 1 + 1);
    return (size_t)o1 >= m * MG_OBS_DIM * sizeof(float) && (size_t)(o2 - o1) >= m * 2 * sizeof(float) && (size_t)(o3 - o2) >= m &&
           (size_t)o3 + m <= need + 4096;
}
constexpr int kMaxHostChunks = 16;
// Events that order mg_step_host's copy stream behind its kernels: created on first use, one set per host thread
// and device, never destroyed (the only objects the library ever creates; no device memory).
cudaEvent_t *host_chunk_events() {
    constexpr int kMaxDev = 32;
    thread_local cudaEvent_t ev[kMaxDev][kMaxHostChunks];
    thread_local bool ready[kMaxDev] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDev) return nullptr;
    if (!ready[dev]) {
        for (int i = 0; i < kMaxHostChunks; ++i)
            if (cudaEventCreateWithFlags(&ev[dev][i], cudaEventDisableTiming) != cudaSuccess) {
                while (--i >= 0) cudaEventDestroy(ev[dev][i]);     // nothing half-made is kept
                return nullptr;
            }
        ready[dev] = true;
    }
    return ev[dev];
}

const MgResetSpec kFixedReset = {MG_RESET_FIXED, 0u, 0ull, 0ull};

int check_reset(const MgResetSpec *r) {
    if (r && r->mode > MG_RESET_RANDOM) return fail(MG_ERR_BAD_FLAGS, "MgResetSpec.mode must be MG_RESET_FIXED or MG_RESET_RANDOM");
    return MG_OK;
}

template <typename ActT>
cudaError_t launch_step(const MgState &s, const MgOut &o, const void *a1, const void *a2, int64_t n,
                        const MgRewards &rw, uint32_t flags, const MgResetSpec &rs, int64_t *stats, cudaStream_t st) {
    constexpr int EPT = MG_EPT;
    const int64_t per_block = (int64_t)mg::kBlock * EPT;
    unsigned grid = (unsigned)((n + per_block - 1) / per_block);
#if MG_STEP_PERSISTENT
    {
        int dev = 0, sms = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (grid > (unsigned)(sms * MG_MIN_BLOCKS)) grid = (unsigned)(sms * MG_MIN_BLOCKS);
    }
#endif
    auto *stp = reinterpret_cast<unsigned long long *>(stats);
    const bool rr = rs.mode == MG_RESET_RANDOM;
    const bool ret = (flags & MG_FLAG_NO_RETURNS) == 0u;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(mg::kBlock); cfg.dynamicSmemBytes = 0; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = MG_PDL ? 1 : 0;
    cudaError_t le = cudaSuccess;
    const uint32_t layout = mg::obs_layout_of(flags);
    if (layout != mg::kObsAos) {
        // the other observation layouts are built for the common case only: uint8 actions, return accumulators kept
        if constexpr (std::is_same<ActT, uint8_t>::value) {
            if (!ret) return cudaErrorInvalidValue;
#define MG_LAUNCH_L(PVP, RR, A2, L) cudaLaunchKernelEx(&cfg, mg::merge_step_kernel<EPT, uint8_t, PVP, RR, true, L>, s, o, \
                                                       (const uint8_t *)a1, (const uint8_t *)(A2), n, rw, flags, rs, stp)
#define MG_LAUNCH_PR(L) (a2 ? (rr ? MG_LAUNCH_L(true, true, a2, L) : MG_LAUNCH_L(true, false, a2, L))             \
                            : (rr ? MG_LAUNCH_L(false, true, nullptr, L) : MG_LAUNCH_L(false, false, nullptr, L)))
            le = layout == mg::kObsSoa ? MG_LAUNCH_PR(1) : MG_LAUNCH_PR(2);
#undef MG_LAUNCH_PR
#undef MG_LAUNCH_L
            return le ? le : cudaGetLastError();
        } else {
            return cudaErrorInvalidValue;
        }
    }
#define MG_LAUNCH(PVP, RR, A2)                                                                        \
    le = ret ? cudaLaunchKernelEx(&cfg, mg::merge_step_kernel<EPT, ActT, PVP, RR, true>, s, o, (const ActT *)a1,  \
                                  (const ActT *)(A2), n, rw, flags, rs, stp)                            \
             : cudaLaunchKernelEx(&cfg, mg::merge_step_kernel<EPT, ActT, PVP, RR, false>, s, o, (const ActT *)a1, \
                                  (const ActT *)(A2), n, rw, flags, rs, stp)
    if (a2) { if (rr) MG_LAUNCH(true, true, a2); else MG_LAUNCH(true, false, a2); }
    else    { if (rr) MG_LAUNCH(false, true, nullptr); else MG_LAUNCH(false, false, nullptr); }
#undef MG_LAUNCH
    return le ? le : cudaGetLastError();
}
}  // namespace

extern "C" {

MG_API int mg_version(void) { return MG_ABI_VERSION; }
MG_API const char *mg_last_error(void) { return mg_abi::g_err; }

MG_API int mg_get_constants(MgConstants *c) {
    if (!c) return fail(MG_ERR_NULL_POINTER, "out is NULL");
    c->R = mg::kR; c->H = mg::kH; c->W = mg::kW; c->dT = mg::kDT;
    c->start_point = mg::kStart; c->end_point = mg::kEnd; c->prediction_t = mg::kPredT;
    c->init_vel = mg::kInitVel; c->action_dv = mg::kActionDv;
    c->vehicle_w = mg::kVehicleW; c->vehicle_h = mg::kVehicleH; c->max_steps = mg::kMaxSteps;
    c->num_actions = MG_NUM_ACTIONS; c->obs_dim = MG_OBS_DIM;
    c->stats_rows = MG_STATS_ROWS; c->stats_cols = MG_STATS_COLS;
    c->return_fixed_point_scale = mg::kRetScale;
    return MG_OK;
}

MG_API int mg_default_rewards(MgRewards *r) {
    if (!r) return fail(MG_ERR_NULL_POINTER, "out is NULL");
    *r = kDefaultRewards;
    return MG_OK;
}

MG_API int mg_reset(const MgState *state, int64_t n, const uint8_t *mask, float *obs, uint32_t flags,
                    const MgResetSpec *reset, void *stream) {
    if (n < 0) return fail(MG_ERR_BAD_SIZE, "n < 0");
    if ((flags & ~(MG_FLAG_OBS_SOA | MG_FLAG_OBS_GOAL_SLOT)) || flags == (MG_FLAG_OBS_SOA | MG_FLAG_OBS_GOAL_SLOT))
        return fail(MG_ERR_BAD_FLAGS, "mg_reset flags: 0, MG_FLAG_OBS_SOA or MG_FLAG_OBS_GOAL_SLOT");
    if (int rc = check_reset(reset)) return rc;
    if (n == 0) return MG_OK;
    if (int rc = check_state(state, state && (state->ret1 || state->ret2))) return rc;
    const unsigned grid = (unsigned)((n + mg::kBlock - 1) / mg::kBlock);
    mg::merge_reset_kernel<<<grid, mg::kBlock, 0, (cudaStream_t)stream>>>(*state, n, mask, obs, reset ? *reset : kFixedReset,
                                                                          mg::obs_layout_of(flags));
    if (cudaError_t e = cudaGetLastError()) return cuda_fail(e, "mg_reset launch");
    return MG_OK;
}

MG_API int mg_step(const MgState *state, int64_t n, const void *a1, const void *a2, int act_dtype,
                   const MgRewards *rewards, const MgOut *out, int64_t *stats, uint32_t flags,
                   const MgResetSpec *reset, void *stream) {
    if (n < 0) return fail(MG_ERR_BAD_SIZE, "n < 0");
    if (int rc = check_reset(reset)) return rc;
    if (flags & ~(MG_FLAG_AUTO_RESET | MG_FLAG_NO_RETURNS | MG_FLAG_OBS_SOA | MG_FLAG_OBS_GOAL_SLOT)) return fail(MG_ERR_BAD_FLAGS, "unknown flag bits");
    if (flags & (MG_FLAG_OBS_SOA | MG_FLAG_OBS_GOAL_SLOT)) {
        if ((flags & MG_FLAG_OBS_SOA) && (flags & MG_FLAG_OBS_GOAL_SLOT)) return fail(MG_ERR_BAD_FLAGS, "MG_FLAG_OBS_SOA and MG_FLAG_OBS_GOAL_SLOT exclude each other");
        if (act_dtype != MG_ACT_U8 || (flags & MG_FLAG_NO_RETURNS))
            return fail(MG_ERR_BAD_FLAGS, "MG_FLAG_OBS_SOA / MG_FLAG_OBS_GOAL_SLOT need uint8 actions and the return accumulators");
    }
    if (act_dtype < MG_ACT_U8 || act_dtype > MG_ACT_I64)
        return fail(MG_ERR_BAD_DTYPE, "act_dtype must be MG_ACT_U8, MG_ACT_I32 or MG_ACT_I64");
    if (n == 0) return MG_OK;
    if (int rc = check_state(state, !(flags & MG_FLAG_NO_RETURNS))) return rc;
    if (int rc = check_out(out, true)) return rc;
    if ((flags & MG_FLAG_NO_RETURNS) && out->ep_ret) return fail(MG_ERR_BAD_FLAGS, "out.ep_ret needs the return accumulators (MG_FLAG_NO_RETURNS is set)");
    if (!a1) return fail(MG_ERR_NULL_POINTER, "a1 is NULL");
    const MgRewards rw = rewards ? *rewards : kDefaultRewards;
    const MgResetSpec rs = reset ? *reset : kFixedReset;
    cudaError_t e;
    switch (act_dtype) {
        case MG_ACT_U8:  e = launch_step<uint8_t>(*state, *out, a1, a2, n, rw, flags, rs, stats, (cudaStream_t)stream); break;
        case MG_ACT_I32: e = launch_step<int32_t>(*state, *out, a1, a2, n, rw, flags, rs, stats, (cudaStream_t)stream); break;
        case MG_ACT_I64: e = launch_step<int64_t>(*state, *out, a1, a2, n, rw, flags, rs, stats, (cudaStream_t)stream); break;
        default: return fail(MG_ERR_BAD_DTYPE, "act_dtype must be MG_ACT_U8, MG_ACT_I32 or MG_ACT_I64");
    }
    if (e) return cuda_fail(e, "mg_step launch");
    return MG_OK;
}

MG_API int mg_sample_actions(uint8_t *a1, uint8_t *a2, int64_t n, uint64_t seed, uint64_t env_id_base,
                             uint64_t step, void *stream) {
    if (n < 0) return fail(MG_ERR_BAD_SIZE, "n < 0");
    if (n == 0) return MG_OK;
    if (!a1) return fail(MG_ERR_NULL_POINTER, "a1 is NULL");
    const unsigned grid = (unsigned)((n + mg::kBlock - 1) / mg::kBlock);
    mg::sample_actions_kernel<<<grid, mg::kBlock, 0, (cudaStream_t)stream>>>(a1, a2, n, seed, env_id_base, step);
    if (cudaError_t e = cudaGetLastError()) return cuda_fail(e, "mg_sample_actions launch");
    return MG_OK;
}

MG_API int mg_rollout(const MgState *state, int64_t n, int pvp, uint64_t seed, uint64_t env_id_base,
                      uint64_t step0, int32_t k_steps, const MgRewards *rewards, const MgOut *out,
                      uint8_t *actions_out, int64_t *stats, uint32_t flags, const MgResetSpec *reset,
                      void *stream) {
    if (int rc = check_reset(reset)) return rc;
    const MgResetSpec rs = reset ? *reset : kFixedReset;
    if (n < 0 || k_steps < 0) return fail(MG_ERR_BAD_SIZE, "n < 0 or k_steps < 0");
    if (flags & ~(MG_FLAG_AUTO_RESET | MG_FLAG_NO_RETURNS)) return fail(MG_ERR_BAD_FLAGS, "unknown flag bits");
    if (n == 0 || k_steps == 0) return MG_OK;
    const bool ret = (flags & MG_FLAG_NO_RETURNS) == 0u;
    if (int rc = check_state(state, ret)) return rc;
    if (int rc = check_out(out, false)) return rc;
    if (!ret && out->ep_ret) return fail(MG_ERR_BAD_FLAGS, "out.ep_ret needs the return accumulators (MG_FLAG_NO_RETURNS is set)");
    const MgRewards rw = rewards ? *rewards : kDefaultRewards;
    const int64_t per_block = (int64_t)mg::kBlock * 2;
    auto *stp = reinterpret_cast<unsigned long long *>(stats);
    const bool rr = rs.mode == MG_RESET_RANDOM;
    mg::PhiloxKeys keys;
    for (int r = 0; r < 10; ++r) {
        keys.k0[r] = (uint32_t)seed + (uint32_t)r * 0x9E3779B9u;
        keys.k1[r] = (uint32_t)(seed >> 32) + (uint32_t)r * 0xBB67AE85u;
    }
    cudaStream_t cst = (cudaStream_t)stream;
    // FAST kernel on the whole warps when it applies (see merge_rollout_fast_kernel), SLOW kernel on the rest
    const bool all_out = out->obs && out->rew && out->done && out->info;
    const bool no_out = !out->obs && !out->rew && !out->done && !out->info;
    const bool fast_ok = (flags & MG_FLAG_AUTO_RESET) && !rr && !actions_out && (no_out || (all_out && (n & 1) == 0));
    const int64_t n_fast = fast_ok ? n / 64 * 64 : 0;
    if (n_fast > 0) {
        const unsigned grid = (unsigned)((n_fast + per_block - 1) / per_block);
#define MG_FAST(PVP, RET, OUT)                                                                            \
        mg::merge_rollout_fast_kernel<PVP, RET, OUT><<<grid, mg::kBlock, 0, cst>>>(                      \
            *state, *out, n, n_fast, keys, env_id_base, step0, k_steps, rw, rs, stp)
        if (pvp) { if (ret) { if (all_out) MG_FAST(true, true, true); else MG_FAST(true, true, false); }
                   else     { if (all_out) MG_FAST(true, false, true); else MG_FAST(true, false, false); } }
        else     { if (ret) { if (all_out) MG_FAST(false, true, true); else MG_FAST(false, true, false); }
                   else     { if (all_out) MG_FAST(false, false, true); else MG_FAST(false, false, false); } }
#undef MG_FAST
        if (cudaError_t e = cudaGetLastError()) return cuda_fail(e, "mg_rollout launch (fast)");
    }
    if (n_fast < n) {
        const unsigned grid = (unsigned)((n - n_fast + per_block - 1) / per_block);
#define MG_LAUNCH(PVP, RR)                                                                             \
    do {                                                                                               \
        if (ret) mg::merge_rollout_kernel<PVP, RR, true><<<grid, mg::kBlock, 0, cst>>>(                \
            *state, *out, actions_out, n, n_fast, keys, env_id_base, step0, k_steps, rw, flags, rs, stp);  \
        else mg::merge_rollout_kernel<PVP, RR, false><<<grid, mg::kBlock, 0, cst>>>(                   \
            *state, *out, actions_out, n, n_fast, keys, env_id_base, step0, k_steps, rw, flags, rs, stp);  \
    } while (0)
        if (pvp) { if (rr) MG_LAUNCH(true, true); else MG_LAUNCH(true, false); }
        else     { if (rr) MG_LAUNCH(false, true); else MG_LAUNCH(false, false); }
#undef MG_LAUNCH
    }
    if (cudaError_t e = cudaGetLastError()) return cuda_fail(e, "mg_rollout launch");
    return MG_OK;
}

MG_API int mg_step_host(const MgState *state, int64_t n, const uint8_t *h_a1, const uint8_t *h_a2,
                        uint8_t *d_a1, uint8_t *d_a2, const MgRewards *rewards, const MgOut *d_out,
                        const MgOut *h_out, int64_t *stats, uint32_t flags, const MgResetSpec *reset,
                        void *stream, void *copy_stream, int32_t chunks) {
    if (n < 0) return fail(MG_ERR_BAD_SIZE, "n < 0");
    if (n == 0) return MG_OK;
    if (!state || !h_a1 || !h_out || !d_out) return fail(MG_ERR_NULL_POINTER, "state, h_a1, d_out or h_out is NULL");
    // d_a1 == NULL: no upload — the kernel reads the actions straight from h_a1 / h_a2, which must then be pinned
    // (device-accessible) host memory
    const bool direct = d_a1 == nullptr;
    if (!direct && h_a2 && !d_a2) return fail(MG_ERR_NULL_POINTER, "h_a2 given without d_a2 scratch");
    if (int rc = check_reset(reset)) return rc;
    cudaStream_t st = (cudaStream_t)stream, cs = (cudaStream_t)copy_stream;
    // Chunked pipeline: with a second stream the envs are stepped in `chunks` pieces; the D2H of piece c (copy
    // stream) overlaps the H2D + kernel of piece c+1 (main stream), so the bus idles only while the first
    // piece is uploaded and stepped.  Pieces are multiples of 256 envs (whole blocks, every array 16-byte aligned).
    int nch = (copy_stream && copy_stream != stream && chunks > 1) ? (chunks > kMaxHostChunks ? kMaxHostChunks : chunks) : 1;
    constexpr int64_t kPieceAlign = (int64_t)mg::kBlock * MG_EPT;    // whole blocks of the step kernel
    int64_t piece = ((n + nch - 1) / nch + kPieceAlign - 1) / kPieceAlign * kPieceAlign;
    cudaEvent_t *ev = nullptr;
    if (nch > 1) {
        ev = host_chunk_events();
        if (!ev) {
            const cudaError_t ce = cudaGetLastError();
            return ce ? cuda_fail(ce, "mg_step_host event") : fail(MG_ERR_BAD_SIZE, "mg_step_host: no event set for this device (index >= 32?)");
        }
    }
    const MgResetSpec rs0 = reset ? *reset : kFixedReset;
    cudaError_t e;
    int c = 0;
    for (int64_t off = 0; off < n; off += piece, ++c) {
        const int64_t m = n - off < piece ? n - off : piece;
        const size_t M = (size_t)m;
        // one upload when the two action vectors sit at the same distance on both sides (single piece only)
        const bool acts_joined = !direct && nch == 1 && h_a2 && d_a2 > d_a1 && (d_a2 - d_a1) == (h_a2 - h_a1) && (size_t)(d_a2 - d_a1) < M + 4096;
        if (direct) {
        } else if (acts_joined) {
            if ((e = cudaMemcpyAsync(d_a1, h_a1, (size_t)(d_a2 - d_a1) + M, cudaMemcpyHostToDevice, st))) return cuda_fail(e, "H2D actions");
        } else {
            if ((e = cudaMemcpyAsync(d_a1 + off, h_a1 + off, M, cudaMemcpyHostToDevice, st))) return cuda_fail(e, "H2D a1");
            if (h_a2 && (e = cudaMemcpyAsync(d_a2 + off, h_a2 + off, M, cudaMemcpyHostToDevice, st))) return cuda_fail(e, "H2D a2");
        }
        const MgState sub = {state->pos1 + off, state->vel1 + off, state->pos2 + off, state->vel2 + off,
                             state->ret1 ? state->ret1 + off : nullptr, state->ret2 ? state->ret2 + off : nullptr,
                             state->meta + off};
        const MgOut d = {d_out->obs ? d_out->obs + off * MG_OBS_DIM : nullptr, d_out->rew ? d_out->rew + off * 2 : nullptr,
                         d_out->done ? d_out->done + off : nullptr, d_out->info ? d_out->info + off : nullptr,
                         d_out->term_obs ? d_out->term_obs + off * MG_OBS_DIM : nullptr,
                         d_out->ep_ret ? d_out->ep_ret + off * 2 : nullptr, d_out->ep_len ? d_out->ep_len + off : nullptr};
        MgResetSpec rs = rs0;
        rs.env_id_base += (uint64_t)off;
        const uint8_t *k_a1 = direct ? h_a1 + off : d_a1 + off;
        const uint8_t *k_a2 = !h_a2 ? nullptr : direct ? h_a2 + off : d_a2 + off;
        if (int rc = mg_step(&sub, m, k_a1, k_a2, MG_ACT_U8, rewards, &d, stats, flags, &rs, stream)) return rc;
        cudaStream_t out_st = st;
        if (nch > 1) {
            if ((e = cudaEventRecord(ev[c], st))) return cuda_fail(e, "mg_step_host event record");
            if ((e = cudaStreamWaitEvent(cs, ev[c], 0))) return cuda_fail(e, "mg_step_host event wait");
            out_st = cs;
        }
        // one download when [obs | rew | done | info] are laid out back to back identically on both sides
        if (nch == 1 && outputs_joined(d, *h_out, M)) {
            const size_t span = (size_t)((const char *)d.info - (const char *)d.obs) + M;
            if ((e = cudaMemcpyAsync(h_out->obs, d.obs, span, cudaMemcpyDeviceToHost, out_st))) return cuda_fail(e, "D2H outputs");
            continue;
        }
#define MG_D2H(field, count, type)                                                                                    \
        if (h_out->field && d.field &&                                                                                \
            (e = cudaMemcpyAsync(h_out->field + off * (count), d.field, M * (count) * sizeof(type), cudaMemcpyDeviceToHost, out_st))) \
            return cuda_fail(e, "D2H " #field);
        MG_D2H(obs, MG_OBS_DIM, float) MG_D2H(rew, 2, float) MG_D2H(done, 1, uint8_t) MG_D2H(info, 1, uint8_t)
        MG_D2H(term_obs, MG_OBS_DIM, float) MG_D2H(ep_ret, 2, float) MG_D2H(ep_len, 1, int32_t)
#undef MG_D2H
    }
    if (nch > 1 && (e = cudaStreamSynchronize(cs))) return cuda_fail(e, "mg_step_host copy-stream sync");
    if ((e = cudaStreamSynchronize(st))) return cuda_fail(e, "mg_step_host sync");
    return MG_OK;
}

MG_API int mg_step_host_async(const MgState *state, int64_t n, const MgHostSlot *slot, uint32_t fields,
                              const MgRewards *rewards, int64_t *stats, uint32_t flags, const MgResetSpec *reset,
                              void *stream, void *copy_stream, void *upload_stream) {
    if (n < 0) return fail(MG_ERR_BAD_SIZE, "n < 0");
    if (!state || !slot || !slot->h_a1) return fail(MG_ERR_NULL_POINTER, "state, slot or slot.h_a1 is NULL");
    if (!copy_stream || !slot->ev_stepped || !slot->ev_done || copy_stream == stream)
        return fail(MG_ERR_NULL_POINTER, "mg_step_host_async needs a copy stream (different from stream) and the slot's events");
    if (fields & ~MG_FIELD_ALL) return fail(MG_ERR_BAD_FLAGS, "unknown field bits");
    const bool upload = upload_stream != nullptr;
    if (upload && (!slot->d_a1 || (slot->h_a2 && !slot->d_a2) || !slot->ev_uploaded || upload_stream == stream))
        return fail(MG_ERR_NULL_POINTER, "upload_stream given: the slot needs d_a1 (d_a2) scratch and ev_uploaded, and a stream of its own");
    cudaStream_t st = (cudaStream_t)stream, cs = (cudaStream_t)copy_stream, us = (cudaStream_t)upload_stream;
    cudaEvent_t ek = (cudaEvent_t)slot->ev_stepped, ed = (cudaEvent_t)slot->ev_done, eu = (cudaEvent_t)slot->ev_uploaded;
    const MgOut *d_out = &slot->d_out, *h_out = &slot->h_out;
    const size_t M = (size_t)n;
    cudaError_t e;
    const uint8_t *k_a1 = slot->h_a1, *k_a2 = slot->h_a2;
    if (upload && n > 0) {
        // the actions travel by copy engine on a stream of their own (under the previous call's device-to-host copies,
        // the other direction of the link); one copy when [a1 | a2] sit at the same distance on both sides
        const bool joined = slot->h_a2 && slot->d_a2 > slot->d_a1 && (slot->d_a2 - slot->d_a1) == (slot->h_a2 - slot->h_a1) &&
                            (size_t)(slot->d_a2 - slot->d_a1) < M + 4096;
        if (joined) {
            if ((e = cudaMemcpyAsync(slot->d_a1, slot->h_a1, (size_t)(slot->d_a2 - slot->d_a1) + M, cudaMemcpyHostToDevice, us))) return cuda_fail(e, "H2D actions");
        } else {
            if ((e = cudaMemcpyAsync(slot->d_a1, slot->h_a1, M, cudaMemcpyHostToDevice, us))) return cuda_fail(e, "H2D a1");
            if (slot->h_a2 && (e = cudaMemcpyAsync(slot->d_a2, slot->h_a2, M, cudaMemcpyHostToDevice, us))) return cuda_fail(e, "H2D a2");
        }
        if ((e = cudaEventRecord(eu, us))) return cuda_fail(e, "mg_step_host_async upload record");
        if ((e = cudaStreamWaitEvent(st, eu, 0))) return cuda_fail(e, "mg_step_host_async upload wait");
        k_a1 = slot->d_a1; k_a2 = slot->h_a2 ? slot->d_a2 : nullptr;
    }
    // d_out was last read by the copies that recorded ev_done (a never-recorded event completes at once): the
    // kernel may overwrite the slot only after them.  This wait is the only coupling between consecutive calls,
    // so with two slots the kernel of call t+1 runs under the copies of call t.
    if ((e = cudaStreamWaitEvent(st, ed, 0))) return cuda_fail(e, "mg_step_host_async slot wait");
    if (int rc = mg_step(state, n, k_a1, k_a2, MG_ACT_U8, rewards, d_out, stats, flags, reset, stream)) return rc;
    if ((e = cudaEventRecord(ek, st))) return cuda_fail(e, "mg_step_host_async event record");
    if ((e = cudaStreamWaitEvent(cs, ek, 0))) return cuda_fail(e, "mg_step_host_async event wait");
    if (n > 0) {
        // one cudaMemcpyAsync per run of selected fields that are adjacent (gap < 4 KB) at equal offsets on both sides
        const char *dp[4] = {(const char *)d_out->obs, (const char *)d_out->rew, (const char *)d_out->done, (const char *)d_out->info};
        char *hp[4] = {(char *)h_out->obs, (char *)h_out->rew, (char *)h_out->done, (char *)h_out->info};
        const size_t sz[4] = {M * MG_OBS_DIM * sizeof(float), M * 2 * sizeof(float), M, M};
        int i = 0;
        while (i < 4) {
            if (!(fields & (1u << i))) { ++i; continue; }
            if (!dp[i] || !hp[i]) return fail(MG_ERR_NULL_POINTER, "a selected output field is NULL in d_out or h_out");
            int j = i;
            while (j + 1 < 4 && (fields & (1u << (j + 1))) && dp[j + 1] && hp[j + 1] &&
                   dp[j + 1] - dp[i] == hp[j + 1] - hp[i] && dp[j + 1] >= dp[j] + sz[j] && dp[j + 1] - (dp[j] + sz[j]) < 4096)
                ++j;
            const size_t span = (size_t)(dp[j] - dp[i]) + sz[j];
            if ((e = cudaMemcpyAsync(hp[i], dp[i], span, cudaMemcpyDeviceToHost, cs))) return cuda_fail(e, "D2H outputs");
            i = j + 1;
        }
    }
    if ((e = cudaEventRecord(ed, cs))) return cuda_fail(e, "mg_step_host_async done record");
    return MG_OK;
}

MG_API int mg_step_host_wait(void *ev_done) {
    if (!ev_done) return fail(MG_ERR_NULL_POINTER, "ev_done is NULL");
    if (cudaError_t e = cudaEventSynchronize((cudaEvent_t)ev_done)) return cuda_fail(e, "mg_step_host_wait");
    return MG_OK;
}

}  // extern "C"
