// merge_device.cuh — per-env device arithmetic of the fused merging-gym step (sm_100a).
//
// One env lives in one thread's registers.  Everything that decides a *discrete* outcome
// (pos > 950, pos >= 950, trunc(x), trunc(y), step count) is computed in float64 with the
// reference's evaluation order and explicit round-to-nearest intrinsics, so no FMA contraction
// can change a rounding the reference performs (merging_env.py:147-154, 48-58).  Outputs are
// cast to float32 at the very end.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/merging_b200.h"

namespace mg {

// ---- compile-time constants (merging_env.py:22-46, 101, 142) --------------------------------
constexpr double kR = 30000.0;
constexpr double kH = 1000.0;
constexpr double kW = 300.0;
constexpr double kHalfW = kW / 2;                 // W/2 = 150.0
constexpr double kDT = 0.2;
constexpr double kStart = 50.0;                   // START_POINT
constexpr double kEnd = 950.0;                    // END_POINT = H - 50
constexpr double kPredT = 3.0;                    // prediction_t
constexpr double kInitVel = 20.0;
constexpr double kActionDv = 10.0;                // action_dict = {a: 10*a}
constexpr int kVehicleW = 4, kVehicleH = 8;
constexpr int kMaxSteps = 2501;                   // first n with fl64 sum_{1..n} 0.2 > 500 (tests/test_oracle_golden.py)
constexpr double kInvR = 1.0 / kR;                // RN(1/30000), folded by the host compiler
constexpr double kInvPredT = 1.0 / kPredT;        // RN(1/3)
constexpr double kAngle0 = 0x1.10f7317226afdp-5;  // np.arctan2(1000, 30000) = 0.033320995878247196
constexpr double kRetScale = 16777216.0;          // 2^24 fixed point for the return statistics

// Reset observation, merging_env.py:208-230 -> observe() at pos=50, vel=20 (float64 values from
// the oracle; y2 - y1 = -30.057386826127185).  tests assert mg_reset's computed obs equals these.
constexpr float kResetDy = -30.057386826127185f;
constexpr float kResetRemaining = 900.0f;

// ---- x / d for a compile-time divisor, bit-identical to IEEE division -----------------------
// q = RN(x*y), r = x - d*q (exact in one FMA), q' = RN(q + r*y)  with y = RN(1/d)
// (Markstein's correction).  oracle/merge_oracle.c::mgo_check_div and tests/test_oracle_c.py
// confirm q' == x/d bit for bit for d = 3 and d = 30000 over the ranges the env reaches.
__device__ __forceinline__ double div_const(double x, double d, double y) {
    const double q = __dmul_rn(x, y);
    const double r = __fma_rn(-d, q, x);
    return __fma_rn(r, y, q);
}

// ---- sin & cos on the lane-angle range ------------------------------------------------------
// angle = atan2(H,R) - lon/R lies in [-0.64, 0.034] for any state reachable before the time
// limit (lon <= 50 + 8*2501), so no range reduction is needed: fdlibm's k_sin/k_cos minimax
// polynomials for |x| <= pi/4 (error < 2^-58) evaluated with FMAs, within 1 ulp of glibc (which
// the reference's np.sin/np.cos call).  M angles are evaluated in lock-step, Horner stage by
// Horner stage, so every 64-bit coefficient is materialised once per stage instead of once per
// angle (the round-1 profile showed 12 % of all issued instructions were UMOV constant pairs).
template <int M>
__device__ __forceinline__ void sincos_poly(const double (&x)[M], double (&s)[M], double (&c)[M]) {
    const double S1 = -1.66666666666666324348e-01, S2 = 8.33333333332248946124e-03,
                 S3 = -1.98412698298579493134e-04, S4 = 2.75573137070700676789e-06,
                 S5 = -2.50507602534068634195e-08, S6 = 1.58969099521155010221e-10;
    const double C1 = 4.16666666666666019037e-02, C2 = -1.38888888888741095749e-03,
                 C3 = 2.48015872894767294178e-05, C4 = -2.75573143513906633035e-07,
                 C5 = 2.08757232129817482790e-09, C6 = -1.13596475577881948265e-11;
    double z[M], w[M], p[M], q[M];
#pragma unroll
    for (int i = 0; i < M; ++i) { z[i] = x[i] * x[i]; w[i] = z[i] * z[i]; }
    // sin(x) = x + x^3 (S1 + z (S2 + z S3 + z^2 S4 + z^3 (S5 + z S6)))
#pragma unroll
    for (int i = 0; i < M; ++i) p[i] = fma(z[i], S4, S3);
#pragma unroll
    for (int i = 0; i < M; ++i) p[i] = fma(z[i], p[i], S2);
#pragma unroll
    for (int i = 0; i < M; ++i) q[i] = fma(z[i], S6, S5);
#pragma unroll
    for (int i = 0; i < M; ++i) p[i] = fma(z[i] * w[i], q[i], p[i]);
#pragma unroll
    for (int i = 0; i < M; ++i) p[i] = fma(z[i], p[i], S1);
#pragma unroll
    for (int i = 0; i < M; ++i) s[i] = fma(z[i] * x[i], p[i], x[i]);
    // cos(x) = t + (((1 - t) - z/2) + z r),  t = 1 - z/2   (FreeBSD k_cos: no cancellation near 1)
#pragma unroll
    for (int i = 0; i < M; ++i) p[i] = fma(z[i], C3, C2);
#pragma unroll
    for (int i = 0; i < M; ++i) p[i] = fma(z[i], p[i], C1);
#pragma unroll
    for (int i = 0; i < M; ++i) q[i] = fma(z[i], C6, C5);
#pragma unroll
    for (int i = 0; i < M; ++i) q[i] = fma(z[i], q[i], C4);
#pragma unroll
    for (int i = 0; i < M; ++i) p[i] = fma(w[i] * w[i], q[i], z[i] * p[i]);
#pragma unroll
    for (int i = 0; i < M; ++i) {
        const double hz = 0.5 * z[i];
        const double t = 1.0 - hz;
        c[i] = t + (((1.0 - t) - hz) + z[i] * p[i]);
    }
}

// lon2coord (merging_env.py:48-58) for M cars at once.  sign[i] = +1 for "ego" (player 1), -1 for
// "opponent".  Longitudes beyond kPolyMaxLon (only reachable when a caller keeps stepping a
// finished env without auto-reset) leave the polynomial's range: one rarely-taken branch for the
// whole batch then uses the CUDA library sincos.
constexpr double kPolyMaxLon = 24000.0;   // |atan2(H,R) - lon/R| <= 0.767 < pi/4 for 0 <= lon <= 24000
// CHECK_RANGE = false: the caller guarantees lon <= kPolyMaxLon.
template <int M, bool CHECK_RANGE = true>
__device__ __forceinline__ void lon2coord_batch(const double (&lon)[M], const double (&sign)[M],
                                                double (&x)[M], double (&y)[M]) {
    double ang[M], s[M], c[M];
    bool wild = false;
#pragma unroll
    for (int i = 0; i < M; ++i) {
        ang[i] = __dsub_rn(kAngle0, div_const(lon[i], kR, kInvR));
        // one compare per car: catches lon beyond the polynomial's range and NaN.  Negative longitudes need no test: a
        // car never moves backwards (vel >= 0), and even lon = -20000 keeps the angle below pi/4.
        if (CHECK_RANGE) wild |= !(lon[i] <= kPolyMaxLon);
    }
    if (CHECK_RANGE && wild) {
#pragma unroll
        for (int i = 0; i < M; ++i) sincos(ang[i], &s[i], &c[i]);
    } else {
        sincos_poly<M>(ang, s, c);
    }
#pragma unroll
    for (int i = 0; i < M; ++i) {
        x[i] = __dmul_rn(kR, s[i]);
        const double d = __dsub_rn(kR, __dmul_rn(kR, c[i]));
        y[i] = __fma_rn(sign[i], d, kHalfW);   // W/2 +/- d: sign*d is exact, one rounding like the reference
    }
}

// ---- Philox4x32-10 (Salmon et al. SC'11) ----------------------------------------------------
__device__ __forceinline__ void philox4x32_10(uint32_t &c0, uint32_t &c1, uint32_t &c2, uint32_t &c3,
                                              uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}

// actions of global env `env_id` at rollout step `step`: a_p = (u32_p * 5) >> 32
__device__ __forceinline__ void philox_actions(uint64_t seed, uint64_t env_id, uint64_t step,
                                               int &a1, int &a2) {
    uint32_t c0 = (uint32_t)env_id, c1 = (uint32_t)(env_id >> 32);
    uint32_t c2 = (uint32_t)step, c3 = (uint32_t)(step >> 32);
    philox4x32_10(c0, c1, c2, c3, (uint32_t)seed, (uint32_t)(seed >> 32));
    a1 = (int)__umulhi(c0, 5u);
    a2 = (int)__umulhi(c1, 5u);
}

// ---- one env, one step ----------------------------------------------------------------------
struct EnvRegs {
    double p1, v1, p2, v2;   // state1/state2 pos, vel
    double R1, R2;           // r1_accumulate, r2_accumulate
    uint32_t meta;           // steps | winner << 12 | done << 14
};

struct StepResult {
    float obs[MG_OBS_DIM];   // observation of the stepped state (terminal obs if the env finished)
    float r1, r2;
    uint32_t info;           // MG_INFO_* bits
    uint32_t steps;          // episode length after this step
    bool done;               // done flag returned for this step
    bool finished;           // done became true in this step (0 -> 1 transition)
};

__device__ __forceinline__ void reset_regs(EnvRegs &e) {   // merging_env.py:208-230, fixed start
    e.p1 = kStart; e.v1 = kInitVel; e.p2 = kStart; e.v2 = kInitVel;
    e.R1 = 0.0; e.R2 = 0.0; e.meta = 0u;
}

__device__ __forceinline__ void write_obs(double x1, double y1, double x2, double y2, double p1, double v1,
                                          double p2, double v2, float *obs) {      // merging_env.py:122-131
    const double dx = __dsub_rn(x2, x1), dy = __dsub_rn(y2, y1), dv = __dsub_rn(v2, v1);
    obs[0] = (float)dx;  obs[1] = (float)dy;  obs[2] = (float)dv;
    obs[3] = (float)__dsub_rn(kEnd, p1);  obs[4] = (float)v1;
    obs[5] = (float)-dx; obs[6] = (float)-dy; obs[7] = (float)-dv;   // a-b == -(b-a) exactly
    obs[8] = (float)__dsub_rn(kEnd, p2);  obs[9] = (float)v2;
}
__device__ __forceinline__ void write_obs(double x1, double y1, double x2, double y2, const EnvRegs &e, float *obs) {
    write_obs(x1, y1, x2, y2, e.p1, e.v1, e.p2, e.v2, obs);
}

__device__ __forceinline__ void observe(const EnvRegs &e, float *obs) {   // merging_env.py:118-132
    const double lon[2] = {e.p1, e.p2}, sign[2] = {1.0, -1.0};
    double x[2], y[2];
    lon2coord_batch<2>(lon, sign, x, y);
    write_obs(x[0], y[0], x[1], y[1], e, obs);
}

__device__ __forceinline__ void reset_obs_fixed(float *obs) {
    obs[0] = 0.f; obs[1] = kResetDy; obs[2] = 0.f; obs[3] = kResetRemaining; obs[4] = (float)kInitVel;
    obs[5] = 0.f; obs[6] = -kResetDy; obs[7] = 0.f; obs[8] = kResetRemaining; obs[9] = (float)kInitVel;
}

// MergeEnv.reset() for one env (merging_env.py:208-230): fixed start (:216-217) or the commented-out
// random start (:219-221).  Keeps and advances the env's reset count (meta bits 15-31), which is the
// Philox counter of the random draw.  Writes the reset observation.
// Returned by value (pos1, vel1, pos2, vel2): a reference parameter of a non-inlined function would force the caller's
// env registers into local memory.
struct StartState { double p1, v1, p2, v2; };
#ifndef MG_RANDOM_START_INLINE
#define MG_RANDOM_START_INLINE 0
#endif
#if MG_RANDOM_START_INLINE
__device__ __forceinline__
#else
static __device__ __noinline__
#endif
StartState random_start_state(uint64_t seed, uint64_t env_id, uint32_t count) {
    uint32_t c0 = (uint32_t)env_id, c1 = (uint32_t)(env_id >> 32), c2 = count, c3 = 0u;
    philox4x32_10(c0, c1, c2, c3, (uint32_t)seed, (uint32_t)(seed >> 32) ^ 0x52535445u);
    const double k32 = 1.0 / 4294967296.0;
    const double u1 = ((double)c0 + 1.0) * k32;             // (0, 1]
    const double u2 = (double)c1 * k32;                     // [0, 1)
    const double r = sqrt(-2.0 * log(u1));
    double sn, cs;
    sincospi(2.0 * u2, &sn, &cs);
    StartState st;
    st.p1 = kStart + (r * cs) * 5.0;                        // START_POINT + np.random.randn() * 5
    st.v1 = kInitVel + (r * sn) * 3.0;                      // 20.0 + np.random.randn() * 3
    st.p2 = kStart + (-4.0 + 8.0 * (((double)c2 + 0.5) * k32));     // + uniform(-VEHICLE_H/2, VEHICLE_H/2)
    st.v2 = kInitVel + (-5.0 + 15.0 * (((double)c3 + 0.5) * k32));  // 20.0 + uniform(-5, 10)
    return st;
}
__device__ __forceinline__ void random_start(EnvRegs &e, uint64_t seed, uint64_t env_id, uint32_t count) {
    const StartState st = random_start_state(seed, env_id, count);
    e.p1 = st.p1; e.v1 = st.v1; e.p2 = st.p2; e.v2 = st.v2;
}

// RANDOM is a compile-time switch: the fixed-start kernels carry none of the random-start code.
template <bool RANDOM>
__device__ __forceinline__ void reset_env(EnvRegs &e, const MgResetSpec &rs, uint64_t local_id, float *obs) {
    const uint32_t count = e.meta >> MG_META_RESETS_SHIFT;
    if (RANDOM) {
        random_start(e, rs.seed, rs.env_id_base + local_id, count);
        observe(e, obs);
    } else {
        e.p1 = kStart; e.v1 = kInitVel; e.p2 = kStart; e.v2 = kInitVel;
        reset_obs_fixed(obs);
    }
    e.R1 = 0.0; e.R2 = 0.0;
    e.meta = (count + 1u) << MG_META_RESETS_SHIFT;
}

// merging_env.py:138-195 for E envs held by one thread, evaluated in lock-step (same operations
// and roundings per env as the scalar reference; batching only lets the compiler share constants
// and interleave independent float64 chains).  a1/a2 already validated into 0..4.
// PVP=false: `action2 is None` -> acc2 = 0 (merging_env.py:152).
//
// The core works on the UNPACKED bookkeeping of an env — steps since reset, winner (0 = None), sticky done — so that a
// caller that keeps envs in registers over many steps (the rollout kernel) packs / unpacks the meta word once per
// launch, not once per step.  `collided` / `timeout` come back as flags; the caller assembles the info byte.
struct StepFlags {
    bool done;        // done flag returned for this step
    bool finished;    // done became true in this step (0 -> 1 transition)
    bool collided, timeout;
};

template <bool PVP, int E, bool RET, bool CHECK_RANGE = true>
__device__ __forceinline__ void env_step_core(double (&p1)[E], double (&v1)[E], double (&p2)[E], double (&v2)[E],
                                              double (&R1)[E], double (&R2)[E], uint32_t (&steps)[E],
                                              uint32_t (&winner)[E], bool (&sticky)[E], const int (&a1)[E],
                                              const int (&a2)[E], const MgRewards &rw, float (&obs)[E][MG_OBS_DIM],
                                              float (&rew1)[E], float (&rew2)[E], StepFlags (&fl)[E]) {
    double lon[2 * E], sign[2 * E], x[2 * E], y[2 * E];
#pragma unroll
    for (int i = 0; i < E; ++i) {
        // :141-143  time_stamp += dT; > 500 first holds at step 2501 -> integer step counter
        steps[i] = min(steps[i] + 1u, (uint32_t)MG_META_STEPS_MASK);
        // :147-150  player 1: acc = mpc_1d(...) == (vt - v)/3; vel = max(0, vel + acc*dT); pos += vel*dT
        {
            const double vt = kActionDv * (double)a1[i];
            const double acc = div_const(__dsub_rn(vt, v1[i]), kPredT, kInvPredT);
            const double v = __dadd_rn(v1[i], __dmul_rn(acc, kDT));
            v1[i] = (v > 0.0) ? v : 0.0;
            p1[i] = __dadd_rn(p1[i], __dmul_rn(v1[i], kDT));
        }
        // :152-154  player 2
        if (PVP) {
            const double vt = kActionDv * (double)a2[i];
            const double acc = div_const(__dsub_rn(vt, v2[i]), kPredT, kInvPredT);
            const double v = __dadd_rn(v2[i], __dmul_rn(acc, kDT));
            v2[i] = (v > 0.0) ? v : 0.0;
        } else {
            v2[i] = (v2[i] > 0.0) ? v2[i] : 0.0;          // max(0, vel + 0*dT)
        }
        p2[i] = __dadd_rn(p2[i], __dmul_rn(v2[i], kDT));
        lon[2 * i] = p1[i]; sign[2 * i] = 1.0;
        lon[2 * i + 1] = p2[i]; sign[2 * i + 1] = -1.0;
    }
    // :156, :118-132, :48-58  geometry once per car (the reference recomputes it in is_collided)
    lon2coord_batch<2 * E, CHECK_RANGE>(lon, sign, x, y);
#pragma unroll
    for (int i = 0; i < E; ++i) {
        const double x1 = x[2 * i], y1 = y[2 * i], x2 = x[2 * i + 1], y2 = y[2 * i + 1];
        write_obs(x1, y1, x2, y2, p1[i], v1[i], p2[i], v2[i], obs[i]);

        // :158-159  reward_i = -time_penalty - vel_penalty*|v_i - 20|
        double r1 = __dsub_rn(-rw.time_penalty, __dmul_rn(rw.vel_penalty, fabs(__dsub_rn(v1[i], 20.0))));
        double r2 = __dsub_rn(-rw.time_penalty, __dmul_rn(rw.vel_penalty, fabs(__dsub_rn(v2[i], 20.0))));
        uint32_t w = winner[i];
        const bool to = steps[i] >= (uint32_t)kMaxSteps;
        bool dn = sticky[i] | to;
        // :163-171  player 1 crosses with strict '>' and is evaluated first (branch-free selects)
        {
            const bool f = p1[i] > kEnd;
            const double bonus = (w == 0u) ? rw.r_first : rw.r_second;
            const double r_cross = (w == 1u) ? 0.0 : __dadd_rn(r1, bonus);
            r1 = f ? r_cross : r1;
            dn |= f & (w == 2u);
            w = (f & (w == 0u)) ? 1u : w;
        }
        // :173-181  player 2 crosses with '>='
        {
            const bool f = p2[i] >= kEnd;
            const double bonus = (w == 0u) ? rw.r_first : rw.r_second;
            const double r_cross = (w == 2u) ? 0.0 : __dadd_rn(r2, bonus);
            r2 = f ? r_cross : r2;
            dn |= f & (w == 1u);
            w = (f & (w == 0u)) ? 2u : w;
        }
        // :183-187, :198-206, :232-239  integer pygame Rects (C truncation) 4 wide (lateral, y) x 8
        // long (longitudinal, x); closed rectangles intersect iff both projections overlap.
        const int ty1 = __double2int_rz(y1), ty2 = __double2int_rz(y2);
        const int tx1 = __double2int_rz(x1), tx2 = __double2int_rz(x2);
        const bool collided = (abs(ty1 - ty2) <= kVehicleW) & (abs(tx1 - tx2) <= kVehicleH);
        dn |= collided;
        r1 = collided ? __dadd_rn(r1, rw.r_collision) : r1;
        r2 = collided ? __dadd_rn(r2, rw.r_collision) : r2;
        // :191-192
        if (RET) {                       // RET = false: MG_FLAG_NO_RETURNS, the accumulators do not exist
            R1[i] = __dadd_rn(R1[i], r1);
            R2[i] = __dadd_rn(R2[i], r2);
        }
        rew1[i] = (float)r1;
        rew2[i] = (float)r2;
        fl[i].done = dn;
        fl[i].finished = dn & !sticky[i];
        fl[i].collided = collided;
        fl[i].timeout = to;
        winner[i] = w;
        sticky[i] = dn;
    }
}

__device__ __forceinline__ uint32_t info_byte(const StepFlags &f, uint32_t winner, bool bad_action) {
    return (f.collided ? MG_INFO_COLLISION : 0u) | (winner << MG_INFO_WINNER_SHIFT) | (f.timeout ? MG_INFO_TIMEOUT : 0u) |
           (f.done ? MG_INFO_DONE : 0u) | (bad_action ? MG_INFO_BAD_ACTION : 0u);
}

// The packed-meta front end the single-step kernel uses: unpack, core, pack.
template <bool PVP, int E, bool RET = true>
__device__ __forceinline__ void env_step_batch(EnvRegs (&env)[E], const int (&a1)[E], const int (&a2)[E],
                                               const bool (&bad_action)[E], const MgRewards &rw,
                                               StepResult (&out)[E]) {
    double p1[E], v1[E], p2[E], v2[E], R1[E], R2[E];
    uint32_t steps[E], winner[E];
    bool sticky[E];
    float obs[E][MG_OBS_DIM], rew1[E], rew2[E];
    StepFlags fl[E];
#pragma unroll
    for (int i = 0; i < E; ++i) {
        const EnvRegs &e = env[i];
        p1[i] = e.p1; v1[i] = e.v1; p2[i] = e.p2; v2[i] = e.v2; R1[i] = e.R1; R2[i] = e.R2;
        steps[i] = e.meta & MG_META_STEPS_MASK;
        winner[i] = (e.meta >> MG_META_WINNER_SHIFT) & 3u;
        sticky[i] = (e.meta & MG_META_DONE) != 0u;
    }
    env_step_core<PVP, E, RET>(p1, v1, p2, v2, R1, R2, steps, winner, sticky, a1, a2, rw, obs, rew1, rew2, fl);
#pragma unroll
    for (int i = 0; i < E; ++i) {
        EnvRegs &e = env[i];
        e.p1 = p1[i]; e.v1 = v1[i]; e.p2 = p2[i]; e.v2 = v2[i]; e.R1 = R1[i]; e.R2 = R2[i];
        e.meta = (e.meta & ~((1u << MG_META_RESETS_SHIFT) - 1u)) | steps[i] | (winner[i] << MG_META_WINNER_SHIFT) |
                 (fl[i].done ? MG_META_DONE : 0u);
#pragma unroll
        for (int k = 0; k < MG_OBS_DIM; ++k) out[i].obs[k] = obs[i][k];
        out[i].r1 = rew1[i];
        out[i].r2 = rew2[i];
        out[i].steps = steps[i];
        out[i].done = fl[i].done;
        out[i].finished = fl[i].finished;
        out[i].info = info_byte(fl[i], winner[i], bad_action[i]);
    }
}

template <typename ActT>
__device__ __forceinline__ int load_action(const ActT *p, int64_t i) { return (int)p[i]; }

// validate into 0..4 (the reference raises KeyError from action_dict[a], merging_env.py:147)
__device__ __forceinline__ int clamp_action(long long a, bool &bad) {
    if (a < 0 || a >= MG_NUM_ACTIONS) { bad = true; a = a < 0 ? 0 : MG_NUM_ACTIONS - 1; }
    return (int)a;
}

// Scatter the optional "finished episode" outputs (rare: ~0.5 % of envs per step).
__device__ __forceinline__ void write_episode_outputs(const MgOut &o, int64_t e, const StepResult &r,
                                                      double R1, double R2) {
    if (o.term_obs) {
        float *t = o.term_obs + e * MG_OBS_DIM;
#pragma unroll
        for (int k = 0; k < MG_OBS_DIM; ++k) t[k] = r.obs[k];
    }
    if (o.ep_ret) { o.ep_ret[2 * e] = (float)R1; o.ep_ret[2 * e + 1] = (float)R2; }
    if (o.ep_len) o.ep_len[e] = (int32_t)r.steps;
}

// ---- observation layouts (MG_FLAG_OBS_*): where the ten values of env e go / come from ------------------------------
enum : uint32_t { kObsAos = 0u, kObsSoa = 1u, kObsGoalSlot = 2u };
__host__ __device__ __forceinline__ uint32_t obs_layout_of(uint32_t flags) {
    return (flags & MG_FLAG_OBS_SOA) ? kObsSoa : (flags & MG_FLAG_OBS_GOAL_SLOT) ? kObsGoalSlot : kObsAos;
}
// one env's observation into any layout (the step kernel's full warps have their own staged path for the default rows)
__device__ __forceinline__ void store_obs(float *__restrict__ obs, uint32_t layout, int64_t e, int64_t n,
                                          const float (&o)[MG_OBS_DIM]) {
    if (layout == kObsAos) {
        float2 *row = reinterpret_cast<float2 *>(obs + e * MG_OBS_DIM);          // 40-byte rows: 8-byte aligned
#pragma unroll
        for (int k = 0; k < MG_OBS_DIM / 2; ++k) row[k] = make_float2(o[2 * k], o[2 * k + 1]);
    } else if (layout == kObsSoa) {
        const int64_t stride = MG_OBS_SOA_STRIDE(n);
#pragma unroll
        for (int k = 0; k < MG_OBS_DIM; ++k) obs[k * stride + e] = o[k];
    } else {
        float *row = obs + e * (MG_OBS_DIM + 1) + 1;                             // slot 0 is the goal policy's
#pragma unroll
        for (int k = 0; k < MG_OBS_DIM; ++k) row[k] = o[k];
    }
}

// ---- per-thread episode statistics, packed so that one warp reduction covers several ---------
struct StatAcc {
    uint32_t a = 0;   // episodes | collisions<<8 | wins_p1<<16 | wins_p2<<24   (8-bit fields: <= 255 events per warp and flush)
    uint32_t b = 0;   // timeouts | merges_ok<<8 | bad_actions<<16
    uint32_t len = 0; // sum of finished episode lengths
    long long fx1 = 0, fx2 = 0;   // sum of finished returns, fixed point 2^24

    __device__ __forceinline__ void add(const StepResult &r, double R1, double R2) {
        if (r.info & MG_INFO_BAD_ACTION) b += 1u << 16;
        if (r.finished) {
            const uint32_t col = r.info & MG_INFO_COLLISION;
            const uint32_t w = (r.info & MG_INFO_WINNER_MASK) >> MG_INFO_WINNER_SHIFT;
            const uint32_t to = (r.info & MG_INFO_TIMEOUT) ? 1u : 0u;
            a += 1u | (col << 8) | ((w == 1u) << 16) | ((w == 2u) << 24);
            b += to | (((col == 0u) & (to == 0u)) << 8);
            len += r.steps;
            fx1 += __double2ll_rn(R1 * kRetScale);
            fx2 += __double2ll_rn(R2 * kRetScale);
        }
    }
};

__device__ __forceinline__ long long warp_sum_ll(long long v, unsigned mask) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(mask, v, o);
    return v;
}

// Warp-aggregate with ballot / REDUX, then one lane issues <= 10 integer atomics into the
// block's statistics row.  Integer adds commute, so totals are bit-reproducible.
__device__ __forceinline__ void flush_stats(const StatAcc &st, unsigned long long *stats_row,
                                            unsigned mask, int lane) {
    if (__ballot_sync(mask, (st.a | st.b) != 0u) == 0u) return;   // nothing happened in this warp
    const uint32_t a = __reduce_add_sync(mask, st.a);
    const uint32_t b = __reduce_add_sync(mask, st.b);
    const uint32_t len = __reduce_add_sync(mask, st.len);
    const long long fx1 = warp_sum_ll(st.fx1, mask);
    const long long fx2 = warp_sum_ll(st.fx2, mask);
    if (lane == 0) {
        auto add = [&](int col, unsigned long long v) { if (v) atomicAdd(stats_row + col, v); };
        add(MG_ST_EPISODES, a & 0xFFu);
        add(MG_ST_COLLISIONS, (a >> 8) & 0xFFu);
        add(MG_ST_WINS_P1, (a >> 16) & 0xFFu);
        add(MG_ST_WINS_P2, (a >> 24) & 0xFFu);
        add(MG_ST_TIMEOUTS, b & 0xFFu);
        add(MG_ST_MERGES_OK, (b >> 8) & 0xFFu);
        add(MG_ST_BAD_ACTIONS, (b >> 16) & 0xFFu);
        add(MG_ST_SUM_LENGTH, len);
        add(MG_ST_SUM_RET1_FX, (unsigned long long)fx1);
        add(MG_ST_SUM_RET2_FX, (unsigned long long)fx2);
    }
}

}  // namespace mg
