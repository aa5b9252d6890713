/*
 * merging_b200.h — C ABI of libmerging_b200.so: the batched, device-resident replacement for
 * the reset()/step() hot path of merging-gym's MergeEnv on NVIDIA B200 (sm_100a).
 *
 * The reference has no native layer: the path is pure Python
 * (/root/reference/merging_gym/envs/merging_env.py + scripts/helper.py).  Every entry point
 * below therefore cites the *Python* interface it replaces; INTEGRATION.md shows the ctypes
 * binding a maintainer of the reference would add.
 *
 * Conventions
 *  - Plain C types only.  All array pointers are CUDA device pointers unless the parameter
 *    name starts with `h_` (host memory).  Array pointers must be 16-byte aligned.
 *  - The library never allocates or frees memory and keeps no global state except the
 *    per-thread last-error string (and, only if mg_step_host is used with a copy stream, a
 *    per-thread, per-device set of CUDA events).  The caller (PyTorch on the Python side) owns
 *    every buffer.
 *  - Launches are asynchronous on the caller's stream (`stream` is a cudaStream_t passed as
 *    void*; NULL = legacy default stream).  No entry point synchronises unless documented.
 *    All device entry points are CUDA-graph capturable.
 *  - Return value: 0 on success, negative MgStatus for argument errors, positive cudaError_t
 *    for CUDA runtime errors.  `mg_last_error()` returns a description.
 *  - There is no CPU fallback anywhere in this library.
 */
#ifndef MERGING_B200_H
#define MERGING_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define MG_API __declspec(dllexport)
#else
#define MG_API __attribute__((visibility("default")))
#endif

#define MG_ABI_VERSION 7
#define MG_OBS_DIM 10      /* merging_env.py:75,118-132 */
#define MG_NUM_ACTIONS 5   /* merging_env.py:101-102   */

typedef enum MgStatus {
    MG_OK = 0,
    MG_ERR_NULL_POINTER = -1,
    MG_ERR_BAD_SIZE = -2,
    MG_ERR_ALIGNMENT = -3,
    MG_ERR_BAD_FLAGS = -4,
    MG_ERR_BAD_DTYPE = -5
} MgStatus;

/* Action element types accepted by mg_step (`action_dict[action]`, merging_env.py:101,147). */
typedef enum MgActionDtype { MG_ACT_U8 = 0, MG_ACT_I32 = 1, MG_ACT_I64 = 2 } MgActionDtype;

/* mg_step / mg_rollout flags */
#define MG_FLAG_AUTO_RESET 0x1u /* gym-0.20 SyncVectorEnv convention: a finished env is reset in
                                   the same call and returns its reset observation           */

#define MG_FLAG_NO_RETURNS 0x2u /* the env keeps no r1_accumulate / r2_accumulate (merging_env.py:191-192, read only by
                                   human_player.py:189-193 and render): MgState.ret1/ret2 are not touched and may be
                                   NULL, MgOut.ep_ret must be NULL, the return columns of `stats` stay 0.  Cuts the
                                   step's traffic from 156 to 124 bytes per env-step.                             */

/* Observation layouts (SURVEY.md 7.4 / 8b `obs_layout`).  The reference's observation is a Python list that its callers
 * slice and concatenate (`state[5:] + state[:5]`, `[goal] + state`: scripts/hdqn.py:285,291,299); on the device the same
 * ten values can be laid out for their consumer:
 *   default             obs[n][10]        the gym-shaped rows (40-byte rows; staged per warp, linear 128-bit stores)
 *   MG_FLAG_OBS_SOA     obs[10][S]        S = MG_OBS_SOA_STRIDE(n): one column per feature, every access of a warp is one
 *                                         contiguous span — the layout a fused policy consumer reads best
 *   MG_FLAG_OBS_GOAL_SLOT obs[n][11]      rows `[goal] + state` (hdqn.py:291): slots 1..10 are the observation, slot 0
 *                                         belongs to the goal policy (mg_mlp_act* with MG_MLP_FLAG_WRITE_GOAL writes its
 *                                         arg-max there); the env never touches slot 0
 * accepted by mg_reset, mg_step (uint8 actions, with return accumulators) and mg_policy_step; the policy kernels read
 * all three (MG_MLP_FLAG_OBS_*).  mg_rollout, mg_step_host* and mg_record_transitions work on the default rows only. */
#define MG_FLAG_OBS_SOA 0x10u
#define MG_FLAG_OBS_GOAL_SLOT 0x20u
#define MG_OBS_SOA_STRIDE(n) (((n) + 15) & ~(int64_t)15)

/* info byte, one per env per step (merging_env.py:144,187 `info["collision"]`, :164-181
 * `self.winner`, :142 time limit, :143/:171/:181/:184 `self.done`). */
#define MG_INFO_COLLISION 0x01u
#define MG_INFO_WINNER_SHIFT 1 /* bits 1-2: winner after this step, 0 = None */
#define MG_INFO_WINNER_MASK 0x06u
#define MG_INFO_TIMEOUT 0x08u
#define MG_INFO_DONE 0x10u
#define MG_INFO_BAD_ACTION 0x80u /* action outside 0..4 (reference raises KeyError); it was clamped */

/* meta word, one uint32 per env */
#define MG_META_STEPS_MASK 0x0FFFu /* steps since reset, saturating at 4095 (time limit = 2501) */
#define MG_META_WINNER_SHIFT 12    /* bits 12-13 */
#define MG_META_DONE 0x4000u       /* sticky done (only ever set when auto-reset is off)        */
#define MG_META_RESETS_SHIFT 15    /* bits 15-31: number of resets this env has had (mod 2^17);
                                      the Philox counter of the random-start draw               */

/* How reset() initialises an env (explicit mg_reset and the auto-reset inside mg_step/mg_rollout).
 *   MG_RESET_FIXED : pos = 50, vel = 20 for both cars            (merging_env.py:216-217, the live code)
 *   MG_RESET_RANDOM: the reference's commented-out random start  (merging_env.py:219-221)
 *       state1 = {pos: 50 + randn*5,        vel: 20 + randn*3}
 *       state2 = {pos: 50 + uniform(-4, 4),  vel: 20 + uniform(-5, 10)}
 *     drawn from Philox4x32-10 with key = seed ^ (0x52535445 << 32) and counter =
 *     (global env id, reset count of that env), Box-Muller in float64 for the two normals: the
 *     start of episode k of env e depends on (seed, e, k) only, never on sharding or launch order. */
#define MG_RESET_FIXED 0u
#define MG_RESET_RANDOM 1u
typedef struct MgResetSpec {
    uint32_t mode;
    uint32_t reserved;
    uint64_t seed;
    uint64_t env_id_base; /* global id of env 0 of this shard */
} MgResetSpec;

/* Env state, structure-of-arrays, one element per env.  Replaces `self.state1/state2`
 * (pos, vel; merging_env.py:216-217), `self.r1_accumulate/r2_accumulate` (:191-192,223-224),
 * and `self.time_stamp / self.winner / self.done` (:211-213) packed in `meta`.
 * float64 on purpose: `pos > 950`, `pos >= 950` and trunc(x), trunc(y) are discontinuous in the
 * state, and the reference carries float64. */
typedef struct MgState {
    double *pos1, *vel1, *pos2, *vel2; /* [n] */
    double *ret1, *ret2;               /* [n] episode-return accumulators */
    uint32_t *meta;                    /* [n] */
} MgState;

/* Per-step outputs (the tuple `obs, rewards, done, info` of merging_env.py:195). */
typedef struct MgOut {
    float *obs;       /* [n,10] row-major; merging_env.py:122-131 order                          */
    float *rew;       /* [n,2]  (reward1, reward2), merging_env.py:189                           */
    uint8_t *done;    /* [n]    0/1; may be NULL (the same bit is MG_INFO_DONE of `info`)                */
    uint8_t *info;    /* [n]    MG_INFO_* bit-field                                              */
    float *term_obs;  /* [n,10] or NULL: written ONLY for envs that finished in this step        */
    float *ep_ret;    /* [n,2]  or NULL: finished episode's (r1_accumulate, r2_accumulate), ditto */
    int32_t *ep_len;  /* [n]    or NULL: finished episode's length in steps, ditto               */
} MgOut;

/* Reward shaping, run-time because the reference's shipped runs varied it
 * (`show_reward()`, merging_env.py:115-116; test_params/dqn directory names). */
typedef struct MgRewards {
    double r_first;      /* RFirst      = 2.0    merging_env.py:28 */
    double r_second;     /* RSecond     = 1.0    merging_env.py:29 */
    double r_collision;  /* RCollision  = -10    merging_env.py:30 */
    double vel_penalty;  /* vel_penalty = 0.001  merging_env.py:31 */
    double time_penalty; /* time_penalty = 0     merging_env.py:32 */
} MgRewards;

/* Geometry / timing constants compiled into the kernels (merging_env.py:22-46,101,142). */
typedef struct MgConstants {
    double R, H, W, dT, start_point, end_point, prediction_t, init_vel, action_dv;
    int32_t vehicle_w, vehicle_h, max_steps /* first step count with time_stamp > 500: 2501 */;
    int32_t num_actions, obs_dim, stats_rows, stats_cols;
    double return_fixed_point_scale; /* 2^24: stats columns 8,9 hold llrint(return * scale) */
} MgConstants;

/* Episode statistics: int64 [MG_STATS_ROWS][MG_STATS_COLS] partial sums, accumulated with
 * integer atomics (order-independent, hence deterministic and world-size invariant).  Sum over
 * rows to get totals.  Replaces the hand-kept counters of scripts/main.py:203-227. */
#define MG_STATS_ROWS 1024
#define MG_STATS_COLS 16
enum {
    MG_ST_EPISODES = 0, MG_ST_COLLISIONS, MG_ST_WINS_P1, MG_ST_WINS_P2, MG_ST_TIMEOUTS,
    MG_ST_MERGES_OK /* done && !collision && !timeout */, MG_ST_SUM_LENGTH, MG_ST_BAD_ACTIONS,
    MG_ST_SUM_RET1_FX, MG_ST_SUM_RET2_FX /* fixed point, scale 2^24 */
};

MG_API int mg_version(void);
MG_API const char *mg_last_error(void);
MG_API int mg_get_constants(MgConstants *out);
MG_API int mg_default_rewards(MgRewards *out); /* merging_env.py:28-32 / show_reward() :115-116 */

/* MergeEnv.reset()  (merging_env.py:208-230) for every env, or only where mask[i] != 0.
 * Writes the state and, if obs != NULL, obs[n,10] for ALL envs (masked-out rows get their
 * current observation, i.e. `observe()`, merging_env.py:118-132). */
MG_API int mg_reset(const MgState *state, int64_t n, const uint8_t *mask_or_null, float *obs_or_null,
                    uint32_t flags /* 0 | MG_FLAG_OBS_SOA | MG_FLAG_OBS_GOAL_SLOT: layout of obs */,
                    const MgResetSpec *reset_or_null /* NULL = MG_RESET_FIXED */, void *stream);

/* MergeEnv.step(action1, action2)  (merging_env.py:138-195) for n envs in one fused launch:
 * kinematics of both cars incl. the mpc_1d controller (scripts/helper.py:152-191), lon2coord,
 * observe, rewards / winner / done, is_collided, return accumulation, optional auto-reset,
 * episode statistics.  a2 == NULL is `action2=None` (pve, merging_env.py:152).
 * stats may be NULL. */
MG_API int mg_step(const MgState *state, int64_t n, const void *a1, const void *a2_or_null,
                   int act_dtype, const MgRewards *rewards, const MgOut *out,
                   int64_t *stats_or_null, uint32_t flags, const MgResetSpec *reset_or_null, void *stream);

/* Synthetic uniform-random discrete actions (the scripts' `env.action_space.sample()`,
 * scripts/main.py:26), counter-based: Philox4x32-10, key = seed, counter = (global env id, step).
 * a2 may be NULL. */
MG_API int mg_sample_actions(uint8_t *a1, uint8_t *a2_or_null, int64_t n, uint64_t seed,
                             uint64_t env_id_base, uint64_t step, void *stream);

/* k_steps consecutive steps per launch with in-kernel Philox actions (same stream of actions as
 * mg_sample_actions at steps step0 .. step0+k-1); state stays in registers between steps.
 * `out` arrays are time-major: obs[k,n,10], rew[k,n,2], done[k,n], info[k,n]; any of them may be
 * NULL to skip that output; term_obs/ep_ret/ep_len are [n] "last finished episode" buffers.
 * actions_out_or_null: uint8 [k,n,2].  pvp != 0 selects two-player mode. */
MG_API int mg_rollout(const MgState *state, int64_t n, int pvp, uint64_t seed, uint64_t env_id_base,
                      uint64_t step0, int32_t k_steps, const MgRewards *rewards, const MgOut *out,
                      uint8_t *actions_out_or_null, int64_t *stats_or_null, uint32_t flags,
                      const MgResetSpec *reset_or_null, void *stream);

/* Host-buffer convenience path (the drop-in for callers that keep Python/NumPy data on the
 * host, like the reference scripts): copies h_a1/h_a2 (uint8[n]) to the device scratch actions
 * d_a1/d_a2, runs mg_step, copies obs/rew/done/info back into the h_out arrays and SYNCHRONISES
 * the stream(s).  Optional members of h_out may be NULL.  Pinned host memory is recommended.
 * d_a1 == NULL: no upload — the kernel reads the actions straight from h_a1 / h_a2, which must then be
 * pinned (device-accessible) host memory.
 * copy_stream_or_null + chunks > 1: the envs are stepped in `chunks` (<= 16) pieces of whole 256-env
 * blocks and the device-to-host copies of a piece run on copy_stream while the next piece is uploaded
 * and stepped on `stream` (the bus is the bottleneck of this path: 52 bytes per env-step).  Results are
 * identical to the single-piece call.  Both streams must belong to the current device. */
MG_API int mg_step_host(const MgState *state, int64_t n, const uint8_t *h_a1,
                        const uint8_t *h_a2_or_null, uint8_t *d_a1, uint8_t *d_a2,
                        const MgRewards *rewards, const MgOut *d_out, const MgOut *h_out,
                        int64_t *stats_or_null, uint32_t flags, const MgResetSpec *reset_or_null,
                        void *stream, void *copy_stream_or_null, int32_t chunks);

/* Output fields a host-buffer step copies back (mg_step_host_async `fields`): the tuple members of
 * `return obs, rewards, done, info` (merging_env.py:195).  A caller that needs only rewards / done / info moves
 * 10 bytes per env-step across PCIe instead of 50. */
#define MG_FIELD_OBS 0x1u
#define MG_FIELD_REW 0x2u
#define MG_FIELD_DONE 0x4u
#define MG_FIELD_INFO 0x8u
#define MG_FIELD_ALL 0xFu

/* Pipelined host-buffer step, the asynchronous half of mg_step_host (same role: `MergeEnv.step` for callers whose
 * actions / results live in host memory, merging_env.py:138-195).  One MgHostSlot = the buffers and events of one step
 * in flight; everything in it belongs to the caller (cudaEvent_t / cudaStream_t as void*; the library creates nothing).
 *   upload_stream (optional): cudaMemcpyAsync h_a1|h_a2 -> d_a1|d_a2 -> record ev_uploaded
 *   stream:      [wait ev_uploaded] [wait ev_done] -> mg_step -> record ev_stepped
 *                (upload_stream == NULL: the kernel reads h_a1 / h_a2 straight from PINNED host memory instead)
 *   copy_stream: [wait ev_stepped] -> cudaMemcpyAsync of the selected `fields` d_out -> h_out -> record ev_done
 * and returns without synchronising.  mg_step_host_wait(slot->ev_done) blocks the host until the copies of that call
 * have landed.  With two slots used alternately, the upload and the kernel of call t+1 run under the device-to-host
 * copies of call t, so the bus never idles; h_a1 / h_a2 of a slot may be overwritten once its ev_done has completed.
 * Selected fields that sit back to back at equal offsets in d_out and h_out travel in one copy, and so do [a1 | a2]. */
typedef struct MgHostSlot {
    const uint8_t *h_a1, *h_a2; /* pinned host actions of this step; h_a2 NULL = pve                       */
    uint8_t *d_a1, *d_a2;       /* device scratch for the uploaded actions (only with an upload stream)     */
    MgOut d_out, h_out;         /* device outputs of the step / their pinned host mirrors                   */
    void *ev_uploaded, *ev_stepped, *ev_done; /* cudaEvent_t, created by the caller (timing disabled)       */
} MgHostSlot;
MG_API int mg_step_host_async(const MgState *state, int64_t n, const MgHostSlot *slot, uint32_t fields,
                              const MgRewards *rewards, int64_t *stats_or_null, uint32_t flags,
                              const MgResetSpec *reset_or_null, void *stream, void *copy_stream,
                              void *upload_stream_or_null);
MG_API int mg_step_host_wait(void *ev_done);

#define MG_MLP_FLAG_MIRROR 0x1u /* evaluate the network on the OPPONENT's view of each observation row,
                                   `state[5:] + state[:5]` (scripts/main.py:199, hdqn.py:285,299): the half-swap
                                   is done while the row is read, no mirrored copy is materialised */

#define MG_MLP_FLAG_OBS_SOA 0x4u       /* obs is [10][MG_OBS_SOA_STRIDE(n)] (MG_FLAG_OBS_SOA)                                  */
#define MG_MLP_FLAG_OBS_GOAL_SLOT 0x8u /* obs is [n][11] rows `[goal] + state` (MG_FLAG_OBS_GOAL_SLOT): a 10-input network
                                          reads slots 1..10, an 11-input network (goal pointer NULL) the whole row       */
#define MG_MLP_FLAG_WRITE_GOAL 0x10u   /* with MG_MLP_FLAG_OBS_GOAL_SLOT: also store the arg-max, as a float, into slot 0
                                          of every row — `goal = upper.choose_goal(state)` feeding `[goal] + state`
                                          (hdqn.py:283,291,303) without a separate goal array                           */
#define MG_MLP_FLAG_PDL 0x2u    /* programmatic dependent launch: the kernel's prologue (weight staging, barrier and tensor-
                                   memory set-up — it reads nothing but the weights) may run while the previous kernel
                                   of the stream is still draining; observations are read only after that kernel has
                                   completed.  Set it when the previous kernel does not write this policy's weights
                                   (in a rollout loop it is the env step).                                           */

/* ---- "next" row: policy in the loop -------------------------------------------
 * Fused forward + arg-max of the reference's Q-network `Net(in, out)`:
 * Linear(in,200)-ReLU-Linear(200,100)-ReLU-Linear(100,out) in fp32 followed by
 * `torch.max(q, 1)[1]` (scripts/main.py:30-47,99-107; scripts/hdqn.py:38-55,82-95,165-177).
 * in = obs_dim (+1 if goal != NULL: the h-DQN controller's `[goal] + state`, hdqn.py:291) must be
 * 10 or 11, out_dim 5 or 3.  Weights are fp32 device arrays, 16-byte aligned: w1t [in][200] is the
 * transpose of fc1.weight; w2p [200][4][28] is the transpose of fc2.weight with its 100 columns
 * split into 4 groups of 25 and each group zero-padded to 28 (w2p[k][g][j] = fc2.weight[25g+j][k]);
 * w3 [out][100] is out.weight as stored.
 * actions: uint8[n] (first maximum wins); q_out: optional float[n][out]. */
MG_API int mg_mlp_act(const float *obs, const uint8_t *goal_or_null, int64_t n, int32_t obs_dim,
                      int32_t out_dim, const float *w1t, const float *b1, const float *w2p,
                      const float *b2, const float *w3, const float *b3, uint8_t *actions,
                      float *q_out_or_null, uint32_t flags, void *stream);

/* Tensor-core variant of mg_mlp_act (tcgen05 + TMEM): the 200x100 layer as an error-compensated 3xTF32
 * product (a*b ~= a_hi*b_hi + a_lo*b_hi + a_hi*b_lo) — fp32-level accuracy but not bit-identical to the
 * fp32 FFMA evaluation, hence a separate, opt-in entry point.  Same arguments except w2_tc: float
 * [25][28][2][8][4] = fc2.weight (zero-padded to 112 rows) split into its tf32 hi part (top 19 bits; row
 * groups 0-13) and the remainder lo = w - hi (row groups 14-27), stacked as one 224-row UMMA B operand in
 * the canonical K-major core-matrix layout [K-step][8-row group][K half][row][4 k].
 * With MG_MLP_FLAG_F16X3 BOTH hidden layers run on the tensor cores as three-product sums of fp16 operands
 * (kind::f16; csrc/mlp_tc16_kernels.cu): w1t and b1 are ignored (may be NULL) and w2_tc points to one packed blob,
 * 16-byte aligned:
 *   bytes [0,64)        float c1, c2 (+ padding): a1 = relu(acc1 * c1) = h1 / 8, h2 = relu(acc2 * c2 + b2)
 *   bytes [64,13376)    __half [52][2][8][8]: layer-1 operand, N = 208 hidden units (200 + zero pad) x K = 16 (inputs in
 *                       K slots 0..in-1, fc1.bias in slot 15, the rest 0), times 2^s1; row groups 0-25 hi = fp16(v),
 *                       26-51 lo = fp16(v - hi); canonical K-major core-matrix layout [8-row group][K half][row][8 k]
 *   bytes [13376,106560) __half [13][28][2][8][8]: layer-2 operand, fc2.weight zero-padded to 112 rows x 208 columns, times
 *                       2^s2; per K-step of 16: row groups 0-13 hi, 14-27 lo
 *   with c1 = 2^-s1 / 8 and c2 = 8 * 2^-s2; s1, s2 chosen so that the largest scaled entry lies in [256, 512).
 * Same ~22 significant bits per product as 3xTF32.  fp16's range is the price: network inputs saturate at +-65504 and
 * hidden-layer-1 activations at 5.2e5 (finite, wrong); the reference's observations (metres, m/s) and its checkpoints
 * stay orders of magnitude below both — backend 1 has no such bound. */
#define MG_MLP_FLAG_F16X3 0x20u
MG_API int mg_mlp_act_tc(const float *obs, const uint8_t *goal_or_null, int64_t n, int32_t obs_dim,
                         int32_t out_dim, const float *w1t, const float *b1, const float *w2_tc,
                         const float *b2, const float *w3, const float *b3, uint8_t *actions,
                         float *q_out_or_null, uint32_t flags, void *stream);


/* ---- "next" row 8f-4 / 8f-1: one policy-in-the-loop iteration per launch ------------------------------------
 * The reference scripts' inner loop (scripts/main.py:194-211, scripts/hdqn.py:288-316)
 *     action = dqn.choose_action(state)                     # Net forward + torch.max, or a random action
 *     next_state, rewards, done, info = env.step(action, action_op)
 * for n envs in ONE kernel: the Q-network forward + arg-max of mg_mlp_act (backend 0, fp32 FFMA: the reference's
 * arithmetic) or mg_mlp_act_tc (backend 1, tcgen05 3xTF32; backend 2, tcgen05 3xF16 = MG_MLP_FLAG_F16X3) with MergeEnv.step as its epilogue — the thread that
 * finishes env e's arg-max also owns env e's state, applies the exploration rule, steps the env and writes the next
 * observation row where the next launch's layer-1 read expects it.  Bit-identical to mg_mlp_act[_tc] followed by
 * mg_step on the same inputs.
 *   obs_in          float[n,10]: the observation the policy acts on (observe() of `state`); out->obs may alias it
 *   goal_or_null    uint8[n]: the h-DQN controller's `[goal] + state` input column (hdqn.py:291); network input 11
 *   w1t..b3         as mg_mlp_act (backend 0: w2 = w2p) or mg_mlp_act_tc (backend 1: w2 = w2_tc, backend 2: w2 = the fp16 operand blob); out_dim is 5
 *   a2_or_null      uint8[n] actions of player 2 (pvp), NULL = `action_op = None` (pve)
 *   flags           MG_FLAG_AUTO_RESET | MG_POLICY_FLAG_EXPLORE | MG_POLICY_FLAG_PDL | MG_FLAG_OBS_SOA | MG_FLAG_OBS_GOAL_SLOT
 *                   (the layout of obs_in AND out->obs) | MG_POLICY_FLAG_GOAL_IN_SLOT (MG_FLAG_NO_RETURNS is implied by
 *                   state->ret1 == NULL)
 *   explore         the scripts' rule `np.random.randn() <= EPISILO ? greedy : np.random.randint(0, 5)` (main.py:103-110):
 *                   randn() <= t has probability Phi(t), so keep_u32 = floor(Phi(t) * 2^32) and the greedy action is
 *                   kept iff a Philox u32 < keep_u32; the draws are Philox4x32-10 keyed by seed ^ ('EXPL' << 32) with
 *                   counter (global env id = reset->env_id_base + e, step ^ env clock), the env clock being the env's meta word
 *                   without its done bit (reset count, winner, steps since reset — never repeats for an env): sharding- and
 *                   launch-invariant, and a CUDA graph replaying identical parameters still draws fresh numbers
 *   actions_out     uint8[n] or NULL: the action taken (what `store_transition` records, main.py:207)
 *   q_out_or_null   float[n,5] Q-values */
#define MG_POLICY_FLAG_EXPLORE 0x100u
#define MG_POLICY_FLAG_GOAL_IN_SLOT 0x400u /* the network has 11 inputs and reads its goal from slot 0 of the MG_FLAG_OBS_GOAL_SLOT rows */
#define MG_POLICY_FLAG_PDL 0x200u     /* as MG_MLP_FLAG_PDL: the previous kernel of the stream does not write the weights */
#define MG_POLICY_BACKEND_FP32 0
#define MG_POLICY_BACKEND_TF32X3 1
#define MG_POLICY_BACKEND_F16X3 2 /* mg_mlp_act_tc with MG_MLP_FLAG_F16X3: w2 = the packed fp16 operand blob (w1t / b1 are
                                     validated but not read) */
typedef struct MgExplore {
    uint64_t seed, step;
    uint32_t keep_u32, reserved;
} MgExplore;
MG_API int mg_policy_step(const MgState *state, int64_t n, const float *obs_in, const uint8_t *goal_or_null,
                          int32_t backend, const float *w1t, const float *b1, const float *w2, const float *b2,
                          const float *w3, const float *b3, const uint8_t *a2_or_null, const MgRewards *rewards,
                          const MgOut *out, int64_t *stats_or_null, uint32_t flags, const MgResetSpec *reset_or_null,
                          const MgExplore *explore_or_null, uint8_t *actions_out_or_null, float *q_out_or_null,
                          void *stream);

/* The exploration rule alone, for callers that keep policy and env in separate launches: overwrites greedy actions /
 * goals in place (`num_choices` = 5 actions, main.py:103-110, or 3 goals, hdqn.py:84-92).  Same Philox stream as
 * mg_policy_step when `salt` is 0; the h-DQN meta-controller uses salt 1 so that goal and action draws differ. */
MG_API int mg_explore(uint8_t *choices, int64_t n, int32_t num_choices, const MgExplore *explore,
                      const uint32_t *meta_or_null /* MgState.meta: the env clock, see above */, uint64_t env_id_base,
                      uint32_t salt, void *stream);

/* ---- "next" rows: device-resident transition writer -----------------------
 * Appends one row per selected env to a ring `ring[capacity][width]` (index = counter % capacity,
 * exactly `DQN.store_transition`, scripts/main.py:115-119), envs in id order, deterministic.
 *   mask_mode 0: every env;  1: `env.winner is not 1` after the step (main.py:209, human_player.py:180);
 *             2: explicit — `info` is then a caller-supplied uint8[n] mask, non-zero = store (used for the h-DQN
 *                meta-controller's per-option rows, hdqn.py:318)
 *   format 0 (width 22): [s(10), a_p, r_p, s'(10)] for player p (main.py:116);  s' is the stepped
 *                        state's observation: term_obs where done (pass NULL without auto-reset)
 *   format 1 (width 14): [s(10), a1, a2, r1, r2], the CSV row of scripts/human_player.py:111,180-181
 *   format 2 (width 24): [g, s(10), a_p, r_int, g', s'(10)], the h-DQN controller's row (scripts/hdqn.py:180-184,
 *                        291-316): g / g' = goal_prev / goal_next (uint8[n], the goals chosen from s and from s'),
 *                        r_int = 1 if g' == goal_status(s) else 0 (hdqn.py:223-236,314); the reference stores it
 *                        every step (mask_mode 0)
 * counter: device uint64, total rows ever appended (caller zero-initialises).
 * scratch: device uint32[mg_record_scratch_words(n)] (= (n+31)/32 + 4), 8-byte aligned; the library's work area for the
 *          block counts / offsets of the call (count -> scan -> write, three launches; the write pass is a programmatic
 *          dependent of the scan; mask_mode 0 needs no count and no scan: one thread advances the counter).
 * env_ids_or_null: int32[capacity], env of each row. */
MG_API int64_t mg_record_scratch_words(int64_t n);
MG_API int mg_record_transitions(const float *obs_prev, const float *obs_next,
                                 const float *term_obs_or_null, const uint8_t *a1,
                                 const uint8_t *a2_or_null, const float *rew, const uint8_t *done,
                                 const uint8_t *info, const uint8_t *goal_prev_or_null,
                                 const uint8_t *goal_next_or_null, int64_t n, int32_t mask_mode,
                                 int32_t format, int32_t player, float *ring, int64_t capacity,
                                 int32_t *env_ids_or_null, uint64_t *counter, uint32_t *scratch,
                                 void *stream);

/* The h-DQN meta-controller's per-step bookkeeping (scripts/hdqn.py:283-320) for n envs in one launch: an option runs
 * until `done or goal == goal_status(state)` (:316) while `extrinsic_reward += reward` (:312); the meta-controller then
 * stores `[state, goal, extrinsic_reward, next_state]` with state == next_state == the observation the option ended in
 * (:315-318).  Per env: s_end = done ? term_obs : obs (the stepped state's observation; term_obs NULL = no auto-reset),
 * sum = extrinsic + rew[e][0], ended = done || goal_next == goal_status(s_end) (goal_status: hdqn.py:223-236);
 * writes s_end_out[n,10], rew_out[n,2] = (sum, 0), ended_out[n], and extrinsic = ended ? 0 : sum.  Feed s_end_out /
 * rew_out / ended_out to mg_record_transitions (format 0, mask_mode 2) to append the rows. */
MG_API int mg_option_update(const float *obs, const float *term_obs_or_null, const float *rew, const uint8_t *done,
                            const uint8_t *goal_next, int64_t n, float *extrinsic, float *s_end_out, float *rew_out,
                            uint8_t *ended_out, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* MERGING_B200_H */
