// policy_env.cuh — MergeEnv.step() as the EPILOGUE of the policy kernels (sm_100a).
//
// The reference scripts' inner loop is `action = dqn.choose_action(state); next_state, rewards, done, info =
// env.step(action, action_op)` (scripts/main.py:194-211, scripts/hdqn.py:288-316).  `mg_policy_step` runs one
// iteration of it for n envs in ONE launch: the Q-network forward + arg-max of mlp_kernels.cu / mlp_tc_kernels.cu
// hands each tile's actions to the env step below through shared memory, applies the scripts' exploration rule on the way,
// and writes the next observation rows into the buffer the next launch's layer-1 read expects.  No action array, no observation copy and no second launch sit
// between the policy and the env.
//
// In the tensor-core kernel the env runs on two extra warps per CTA, a tile or two behind the policy warps, so its
// arithmetic and its memory latency hide behind the following tiles' matrix work (see Handoff below); in the fp32
// kernel, whose tile takes 67 000 cycles, every thread steps one env of the tile it has just finished.
#pragma once
#include "merge_device.cuh"

namespace mgpe {

// Kernel parameter block of the fused env epilogue (passed by value).
struct Args {
    MgState s;
    MgOut o;                         // o.obs = where the NEXT observation rows go (may alias the rows the policy read)
    const uint8_t *a2;               // player 2's actions (pvp), or NULL (pve: merging_env.py:152)
    uint8_t *actions;                // the action actually taken (after exploration), or NULL
    unsigned long long *stats;       // statistics rows, or NULL
    MgRewards rw;
    MgResetSpec rs;
    uint64_t explore_seed, explore_step;
    uint32_t explore_keep;           // keep the greedy action iff Philox u32 < explore_keep
    uint32_t flags;                  // MG_FLAG_AUTO_RESET | MG_POLICY_FLAG_EXPLORE | MG_FLAG_OBS_SOA | MG_FLAG_OBS_GOAL_SLOT
    int64_t n;                       // number of envs (the SoA column stride derives from it)
};

// The scripts' exploration rule (main.py:103-110, hdqn.py:84-92,168-176): `if np.random.randn() <= EPISILO:` the
// greedy action, else `np.random.randint(0, NUM_ACTIONS)`.  randn() <= t holds with probability Phi(t), so the draw is
// one uniform against explore_keep = floor(Phi(t) * 2^32) and one uniform action; both come from one Philox4x32-10
// block keyed by (seed ^ 'EXPL') with counter (global env id, step ^ the env's own clock): independent of sharding and
// launch shape.  The env's clock is its meta word without the done bit — (reset count, winner, steps since reset), which
// never repeats for an env — so a CUDA graph that replays the same launch parameters still draws fresh numbers.
__device__ __forceinline__ int explore(const Args &A, int64_t e, int greedy, int num_actions, uint32_t meta) {
    if (!(A.flags & MG_POLICY_FLAG_EXPLORE)) return greedy;
    const uint64_t gid = A.rs.env_id_base + (uint64_t)e;
    const uint32_t clock = meta & ~(uint32_t)MG_META_DONE;
    uint32_t c0 = (uint32_t)gid, c1 = (uint32_t)(gid >> 32), c2 = (uint32_t)A.explore_step ^ clock, c3 = (uint32_t)(A.explore_step >> 32);
    mg::philox4x32_10(c0, c1, c2, c3, (uint32_t)A.explore_seed, (uint32_t)(A.explore_seed >> 32) ^ 0x4558504Cu);
    return c0 < A.explore_keep ? greedy : (int)__umulhi(c1, (uint32_t)num_actions);
}

// One env of one lane: state and player 2's action as loaded from global memory.  Loading is split from stepping so
// that the env warp can request the next round's state before it computes the current one.
struct Loaded {
    mg::EnvRegs env;
    int a2;
    bool bad, valid;
};
template <bool PVP>
__device__ __forceinline__ void load_env(const Args &A, int64_t e, int64_t n, Loaded &x) {
    using namespace mg;
    const bool ret = A.s.ret1 != nullptr;
    x.valid = e < n;
    x.bad = false;
    x.a2 = 0;
    if (x.valid) {
        x.env = EnvRegs{A.s.pos1[e], A.s.vel1[e], A.s.pos2[e], A.s.vel2[e], ret ? A.s.ret1[e] : 0.0, ret ? A.s.ret2[e] : 0.0,
                        A.s.meta[e]};
        if (PVP) x.a2 = clamp_action((long long)A.a2[e], x.bad);
    } else {
        reset_regs(x.env);
    }
}

// One env, one step, all outputs.  `greedy` is the arg-max of player 1's Q-values (0..4); the exploration rule is
// applied here.  Called by a whole warp (the statistics flush that follows is a warp collective); lanes whose env is
// at or beyond n carry valid = false and store nothing.  PVP is a template parameter — not a run-time branch around
// two copies of the env arithmetic — because this code runs once per tile between other warps' instruction streams
// and is paid for by its instruction fetches (profiles/r02_policy_step_tc_*: ~10 cycles per instruction when cold).
template <bool PVP>
__device__ __forceinline__ void step_loaded(const Args &A, int64_t e, Loaded &x, int greedy, mg::StatAcc &st) {
    using namespace mg;
    const bool ret = A.s.ret1 != nullptr;
    EnvRegs env[1] = {x.env};
    int act1[1], act2[1] = {x.a2};
    bool bad[1] = {x.bad};
    const int a1 = act1[0] = explore(A, e, greedy, MG_NUM_ACTIONS, env[0].meta);
    StepResult res[1];
    env_step_batch<PVP, 1, true>(env, act1, act2, bad, A.rw, res);
    if (!ret) { env[0].R1 = 0.0; env[0].R2 = 0.0; }      // MG_FLAG_NO_RETURNS semantics: the accumulators do not exist
    StepResult &r = res[0];
    if (x.valid) {
        if (A.stats) st.add(r, env[0].R1, env[0].R2);
        if (r.finished) write_episode_outputs(A.o, e, r, env[0].R1, env[0].R2);
    }
    if (r.done && (A.flags & MG_FLAG_AUTO_RESET)) {      // gym-0.20 vector convention: return the reset observation
        if (A.rs.mode == MG_RESET_RANDOM) reset_env<true>(env[0], A.rs, (uint64_t)e, r.obs);
        else reset_env<false>(env[0], A.rs, (uint64_t)e, r.obs);
    }
    if (!x.valid) return;
    A.s.pos1[e] = env[0].p1; A.s.vel1[e] = env[0].v1; A.s.pos2[e] = env[0].p2; A.s.vel2[e] = env[0].v2;
    if (ret) { A.s.ret1[e] = env[0].R1; A.s.ret2[e] = env[0].R2; }
    A.s.meta[e] = env[0].meta;
    store_obs(A.o.obs, obs_layout_of(A.flags), e, A.n, r.obs);
    *reinterpret_cast<float2 *>(A.o.rew + 2 * e) = make_float2(r.r1, r.r2);
    if (A.o.done) A.o.done[e] = r.done ? 1 : 0;
    A.o.info[e] = (uint8_t)r.info;
    if (A.actions) A.actions[e] = (uint8_t)a1;
}

// ---- the hand-over between the policy warps and the env warps: NB action tiles in shared memory ----------------------
// The tensor-core policy kernel is bound by its matrix pipeline and its epilogue warps have no slack (ncu,
// profiles/r02_policy_step_tc_*): an env step run BY the epilogue threads lengthened the tile period from 9 200 to
// 18 400 cycles.  Extra warps therefore own the env: the arg-max threads drop tile tl's actions into tile[tl % NB]
// and arrive on full[tl % NB]; the env warps step the tile's envs — one env per lane, 32-env rounds of the same small
// loop body dealt round-robin over the warps — while the policy warps are already
// on the following tiles (a round's state is requested before the warp waits for its actions), and hand the buffer
// back through empty[] as soon as the actions are in registers.  mbarrier arrive / try_wait carry release /
// acquire semantics at CTA scope.
template <int TM, int NB>
struct Handoff {
    unsigned long long full[NB], empty[NB];
    uint8_t tile[NB][TM];
};
constexpr uint32_t kSpin = 1u << 26;
__device__ __forceinline__ uint32_t hs_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void hs_init(unsigned long long *b, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(hs_u32(b)), "r"(count));
}
__device__ __forceinline__ void hs_arrive(unsigned long long *b) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(hs_u32(b)) : "memory");
}
__device__ __forceinline__ void hs_wait(unsigned long long *b, uint32_t parity) {
    uint32_t done = 0;
    for (uint32_t it = 0; it < kSpin && !done; ++it)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(done) : "r"(hs_u32(b)), "r"(parity) : "memory");
    if (!done) __trap();                              // never hang the GPU on a protocol bug
}
// init by one thread before the CTA-wide barrier that precedes the role split; `writers` threads arrive on full[],
// `readers` (one lane per env warp) on empty[]
template <int TM, int NB>
__device__ __forceinline__ void handoff_init(Handoff<TM, NB> &H, uint32_t writers, uint32_t readers) {
    for (int b = 0; b < NB; ++b) { hs_init(&H.full[b], writers); hs_init(&H.empty[b], readers); }
}
// policy side, tile number tl (0, 1, ...) of this CTA: wait until the env warp has released the buffer, then the caller
// stores its actions into the returned tile and calls handoff_publish
template <int TM, int NB>
__device__ __forceinline__ uint8_t *handoff_acquire(Handoff<TM, NB> &H, uint32_t tl) {
    hs_wait(&H.empty[tl % NB], ((tl / NB) & 1u) ^ 1u);
    return H.tile[tl % NB];
}
template <int TM, int NB>
__device__ __forceinline__ void handoff_publish(Handoff<TM, NB> &H, uint32_t tl) { hs_arrive(&H.full[tl % NB]); }

// env warp w of NW: the CTA's env work is a sequence of ROUNDS of 32 envs — round q = (tile number tl = q / R, round r
// = q % R of that tile, R = TM / 32) — and warp w takes q = w, w + NW, ...: whatever NW is, the rounds are spread evenly
// and a warp has NW / R tile periods for each of its rounds.  A tile's buffer goes back to the policy warps when its R
// rounds have arrived on empty[] (count R).  The state of a round does not depend on the actions: it is requested
// before the warp waits for them.
template <int TM, int NB, int NW, bool PVP>
__device__ __forceinline__ void env_warp_loop(const Args &A, Handoff<TM, NB> &H, int w, int64_t first_tile, int64_t tile_stride,
                                              int64_t n_tiles, int64_t n, int lane, int stats_row) {
    static_assert(TM % 32 == 0, "a tile is a whole number of warp rounds");
    constexpr uint32_t R = TM / 32;
    const int64_t my_tiles = first_tile < n_tiles ? (n_tiles - first_tile + tile_stride - 1) / tile_stride : 0;
    const uint32_t total = (uint32_t)my_tiles * R;
    mg::StatAcc st;
    for (uint32_t q = (uint32_t)w; q < total; q += NW) {
        const uint32_t tl = q / R, r = q % R, buf = tl % NB;
        const int64_t e = (first_tile + (int64_t)tl * tile_stride) * TM + 32 * r + lane;
        Loaded x;
        load_env<PVP>(A, e, n, x);
        hs_wait(&H.full[buf], (tl / NB) & 1u);
        const int greedy = (int)H.tile[buf][32 * r + lane];
        __syncwarp();
        if (lane == 0) hs_arrive(&H.empty[buf]);         // the action is in a register: the slot may be reused
#ifndef MG_TC_ENV_NOWORK                                 // -DMG_TC_ENV_NOWORK: timing experiment only (hand-over without the env
        step_loaded<PVP>(A, e, x, greedy, st);           // step; profiles/r02_policy_step_tc_regions.md)
#else
        (void)greedy;
#endif
    }
    if (A.stats) mg::flush_stats(st, A.stats + (size_t)stats_row * MG_STATS_COLS, 0xFFFFFFFFu, lane);
}

}  // namespace mgpe
