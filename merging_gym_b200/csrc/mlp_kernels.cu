// mlp_kernels.cu — fused Q-network forward + arg-max for policy-in-the-loop rollouts (sm_100a).
//
// The reference's Q-networks are all `Net(in, out)`: Linear(in,200)-ReLU-Linear(200,100)-ReLU-
// Linear(100,out) with in in {10, 11}, out in {5, 3} (scripts/main.py:30-47, scripts/hdqn.py:38-55),
// evaluated in fp32 and followed by `torch.max(q, 1)[1]` (main.py:104-105).  One launch here does the
// whole thing for N envs: obs rows in, uint8 actions out, nothing but 41 B/env touches HBM.
//
// fp32 FFMA on purpose: the reference computes in fp32 and the arg-max must not flip on near ties,
// so no tf32/bf16 tensor-core path (45 kFLOP/env: ~250 us for 2^18 envs at FFMA rates).
//
// Tiling (persistent CTAs, one per SM, 256 threads, tile = 256 envs):
//   input    thread t reads the rows of envs {t % 64 + 64 r, r = 0..3} straight into registers; the next tile's
//            rows are requested before the epilogue of the current one so their latency is hidden;
//   layer 1  thread t owns those 4 envs and one quarter (t / 64) of the current K-chunk's 100 units:
//            h1[k][e] = relu(b1[k] + sum_i x[e][i] W1t[i][k]), four k's per warp-uniform 128-bit weight load
//            feeding all 4 envs (packed FFMA2), written K-major to shared memory (thread = 1 env made layer 1
//            shared-memory bound: 500 weight loads x 2.4 wavefronts per warp and tile);
//   layer 2  thread t = (eg = t/4, ng = t%4) owns a 4-env x 25-neuron register tile held as float2
//            pairs of adjacent neurons and updated with Blackwell's packed FFMA2 (fma.rn.f32x2;
//            52 FFMA2 per k against 1 + 7 LDS.128): acc[e][j] += h1[k][4 eg + e] * W2t[k][25 ng + j];
//   layer 3  each thread reduces its 25 neurons into 4 x OUT partial Q-values, the 4 ng-lanes of an
//            env group are adjacent lanes -> two shuffle-xor steps; lane ng==0 adds b3, takes the
//            first maximum and stores 4 actions with one 32-bit store.
// Shared memory: W2 pre-padded on the host to [200][4][28] (89.6 KB, copied with 128-bit loads),
// h1 chunk [100][256] (102 KB), W1t, W3, biases -> ~204 KB, so one CTA per SM.
#include <cstring>

#include "abi_common.h"
#include "policy_env.cuh"

namespace mgmlp {

constexpr int H1 = 200, H2 = 100;
constexpr int TM = 256;            // envs per tile == threads per block
constexpr int KC = 100;            // K-chunk of layer 2 (h1 rows resident in smem), multiple of 4
constexpr int NG = 4, NJ = 25;     // neuron groups x neurons per group
constexpr int NJP = 28;            // padded group width (16-byte aligned rows, zero filled)
constexpr int NP = (NJ + 1) / 2;   // float2 accumulator pairs per env (13: the 26th lane multiplies the zero pad)
constexpr int MAX_OUT = 8;
#ifndef MG_MLP_LR
#define MG_MLP_LR 4
#endif
constexpr int LR = MG_MLP_LR;       // layer 1: envs per thread (each weight load feeds LR envs); the units are split LR ways

template <int IN, int OUT>
struct Smem {
    float w2[H1][NG][NJP];         // 89 600 B, copied verbatim from the host-padded W2
    float h1[KC][TM];              // 102 400 B
    float w1[IN][H1];
    float w3[OUT][H2];
    float b1[H1], b2[H2], b3[MAX_OUT];
    uint8_t act_tile[TM];          // ENV: the tile's greedy actions, handed from the arg-max lanes to one thread per env
};

// one env's input row [goal] + obs straight from global memory into registers (40-byte rows, 8-byte
// aligned; a warp covers one contiguous 1280-byte span)
// COHERENT: the fused env epilogue writes observation rows in the same launch (possibly into the buffer being read), so
// the rows must not travel through the non-coherent (ld.global.nc) path.
// obs_mode: bits 0-1 = observation layout (mg::kObsAos / kObsSoa / kObsGoalSlot), bit 8 = MG_MLP_FLAG_WRITE_GOAL.
template <int IN, bool MIRROR, bool COHERENT = false>
__device__ __forceinline__ void load_row(const float *obs, const uint8_t *__restrict__ goal, int64_t e, int64_t n,
                                         int obs_mode, float (&x)[IN]) {
    constexpr int off = IN - MG_OBS_DIM;   // 1 when a goal column is prepended (hdqn.py:291); compile-time so x[] stays in registers
    const uint32_t layout = (uint32_t)obs_mode & 3u;
    if (e < n) {
        if (layout == mg::kObsAos) {
            if (off) x[0] = (float)goal[e];
            if (!MIRROR) {
                const float2 *src = reinterpret_cast<const float2 *>(obs + e * MG_OBS_DIM);
#pragma unroll
                for (int i = 0; i < MG_OBS_DIM / 2; ++i) {      // five float2
                    const float2 v = COHERENT ? __ldcg(src + i) : __ldg(src + i);
                    x[off + 2 * i] = v.x; x[off + 2 * i + 1] = v.y;
                }
            } else {                                            // the opponent's view: state[5:] + state[:5] (main.py:199)
#pragma unroll
                for (int i = 0; i < MG_OBS_DIM; ++i) {
                    const float *q = obs + e * MG_OBS_DIM + (i + MG_OBS_DIM / 2) % MG_OBS_DIM;
                    x[off + i] = COHERENT ? __ldcg(q) : __ldg(q);
                }
            }
        } else {
            // [10][stride] columns, or [n][11] rows `[goal] + state` whose slot 0 an 11-input network reads as its goal
            const int64_t stride = MG_OBS_SOA_STRIDE(n);
            if (off) x[0] = goal ? (float)goal[e] : __ldcg(obs + e * (MG_OBS_DIM + 1));
#pragma unroll
            for (int i = 0; i < MG_OBS_DIM; ++i) {
                const int k = MIRROR ? (i + MG_OBS_DIM / 2) % MG_OBS_DIM : i;
                x[off + i] = __ldcg(layout == mg::kObsSoa ? obs + k * stride + e : obs + e * (MG_OBS_DIM + 1) + 1 + k);
            }
        }
    } else {
#pragma unroll
        for (int i = 0; i < IN; ++i) x[i] = 0.f;
    }
}

// ENV: `mg_policy_step` — the env step runs as this kernel's epilogue (policy_env.cuh): the arg-max lanes drop the
// tile's 256 actions into shared memory and every thread steps one env of the tile.  (A ninth warp that owns the env,
// as in the tensor-core kernel, would cap this kernel at 168 registers — ptxas sizes the budget for 384 threads — and
// spill the layer-2 register tile; here the env costs 6 450 of a tile's 67 000 cycles.)
// ENV: 0 = policy only, 1 = + env step pve, 2 = + env step pvp
template <int IN, int OUT, bool MIRROR, int ENV>
__global__ void __launch_bounds__(TM, 1)
mlp_act_kernel(const float *obs, const uint8_t *__restrict__ goal, const int64_t n, const int obs_mode,
               const float *__restrict__ w1t, const float *__restrict__ b1, const float *__restrict__ w2p,
               const float *__restrict__ b2, const float *__restrict__ w3, const float *__restrict__ b3,
               uint8_t *__restrict__ act, float *__restrict__ q_out, const mgpe::Args P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Smem<IN, OUT> &S = *reinterpret_cast<Smem<IN, OUT> *>(smem_raw);
    const int t = threadIdx.x;
    const int ng = t & 3, eg = t >> 2;
    const int64_t n_tiles = (n + TM - 1) / TM;

    // layer-1 mapping: thread = LR envs {e0 + r * TM/LR} x one LR-th of the K-chunk's units, so that every weight read
    // from shared memory (a warp-uniform LDS.128, 2.4 wavefronts) feeds LR envs instead of one
    const int e0 = t & (TM / LR - 1), uh = t / (TM / LR);
    float xr[LR][IN];
    // ---- weights -> shared memory, once per (persistent) CTA: straight 128-bit copies -------------
    {
        const float4 *src = reinterpret_cast<const float4 *>(w2p);
        float4 *dst = reinterpret_cast<float4 *>(&S.w2[0][0][0]);
#pragma unroll 11
        for (int i = t; i < H1 * NG * NJP / 4; i += TM) dst[i] = __ldg(src + i);
        const float4 *s1 = reinterpret_cast<const float4 *>(w1t);
        float4 *d1 = reinterpret_cast<float4 *>(&S.w1[0][0]);
        for (int i = t; i < IN * H1 / 4; i += TM) d1[i] = __ldg(s1 + i);
        for (int i = t; i < OUT * H2; i += TM) (&S.w3[0][0])[i] = w3[i];
        for (int i = t; i < H1; i += TM) S.b1[i] = b1[i];
        for (int i = t; i < H2; i += TM) S.b2[i] = b2[i];
        if (t < OUT) S.b3[t] = b3[t];
    }
    // The weight staging above reads nothing the previous kernel of the stream can have written: under programmatic
    // dependent launch (MG_MLP_FLAG_PDL) it overlaps that kernel's drain.  Observations are read from here on.
    cudaGridDependencySynchronize();
    if (ENV) cudaTriggerProgrammaticLaunchCompletion();
#pragma unroll
    for (int r = 0; r < LR; ++r) load_row<IN, MIRROR, ENV != 0>(obs, goal, (int64_t)blockIdx.x * TM + e0 + r * (TM / LR), n, obs_mode, xr[r]);
    __syncthreads();

    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t base = tile * TM;
        // accumulators as float2 pairs over adjacent neurons: Blackwell's packed FFMA2
        // (fma.rn.f32x2) does two fp32 FMAs per issue slot with 64-bit register operands
        float2 acc[4][NP];
#pragma unroll
        for (int e = 0; e < 4; ++e)
#pragma unroll
            for (int j = 0; j < NP; ++j) acc[e][j] = make_float2(0.f, 0.f);

        for (int k0 = 0; k0 < H1; k0 += KC) {
            // ---- layer 1 for rows k0 .. k0+KC of h1 (thread = env), 4 rows per 128-bit weight load ----
            __syncthreads();                   // everyone finished reading the previous chunk of h1
            constexpr int kGroups = KC / 4;                                   // 25 groups of 4 units, split LR ways
            const int g_begin = kGroups * uh / LR, g_end = kGroups * (uh + 1) / LR;
#pragma unroll 2
            for (int g = g_begin; g < g_end; ++g) {
                const int kk = 4 * g;
                const float4 bb = *reinterpret_cast<const float4 *>(&S.b1[k0 + kk]);
                float2 h01[LR], h23[LR];
#pragma unroll
                for (int r = 0; r < LR; ++r) { h01[r] = make_float2(bb.x, bb.y); h23[r] = make_float2(bb.z, bb.w); }
#pragma unroll
                for (int i = 0; i < IN; ++i) {
                    const float4 w = *reinterpret_cast<const float4 *>(&S.w1[i][k0 + kk]);
                    const float2 w01 = make_float2(w.x, w.y), w23 = make_float2(w.z, w.w);
#pragma unroll
                    for (int r = 0; r < LR; ++r) {
                        const float2 xx = make_float2(xr[r][i], xr[r][i]);
                        h01[r] = __ffma2_rn(xx, w01, h01[r]);
                        h23[r] = __ffma2_rn(xx, w23, h23[r]);
                    }
                }
#pragma unroll
                for (int r = 0; r < LR; ++r) {
                    const int e = e0 + r * (TM / LR);
                    S.h1[kk][e] = fmaxf(h01[r].x, 0.f); S.h1[kk + 1][e] = fmaxf(h01[r].y, 0.f);
                    S.h1[kk + 2][e] = fmaxf(h23[r].x, 0.f); S.h1[kk + 3][e] = fmaxf(h23[r].y, 0.f);
                }
            }
            __syncthreads();
            // ---- layer 2 partial sums over this chunk (4 envs x 25 neurons per thread), operands of
            //      step kk+1 are fetched from shared memory while step kk multiplies ------------------
            float4 a_n = *reinterpret_cast<const float4 *>(&S.h1[0][4 * eg]);
            float4 b_n[NJP / 4];
            {
                const float4 *bp = reinterpret_cast<const float4 *>(&S.w2[k0][ng][0]);
#pragma unroll
                for (int v = 0; v < NJP / 4; ++v) b_n[v] = bp[v];
            }
#pragma unroll 2
            for (int kk = 0; kk < KC; ++kk) {
                const float4 a = a_n;
                float2 b[NJP / 2];
#pragma unroll
                for (int v = 0; v < NJP / 4; ++v) {
                    b[2 * v] = make_float2(b_n[v].x, b_n[v].y); b[2 * v + 1] = make_float2(b_n[v].z, b_n[v].w);
                }
                if (kk + 1 < KC) {
                    a_n = *reinterpret_cast<const float4 *>(&S.h1[kk + 1][4 * eg]);
                    const float4 *bp = reinterpret_cast<const float4 *>(&S.w2[k0 + kk + 1][ng][0]);
#pragma unroll
                    for (int v = 0; v < NJP / 4; ++v) b_n[v] = bp[v];
                }
                const float2 a0 = make_float2(a.x, a.x), a1 = make_float2(a.y, a.y),
                             a2 = make_float2(a.z, a.z), a3 = make_float2(a.w, a.w);
#pragma unroll
                for (int j = 0; j < NP; ++j) {
                    acc[0][j] = __ffma2_rn(a0, b[j], acc[0][j]);
                    acc[1][j] = __ffma2_rn(a1, b[j], acc[1][j]);
                    acc[2][j] = __ffma2_rn(a2, b[j], acc[2][j]);
                    acc[3][j] = __ffma2_rn(a3, b[j], acc[3][j]);
                }
            }
        }
        // next tile's input rows: issued now, consumed after the epilogue
#pragma unroll
        for (int r = 0; r < LR; ++r) load_row<IN, MIRROR, ENV != 0>(obs, goal, (tile + gridDim.x) * TM + e0 + r * (TM / LR), n, obs_mode, xr[r]);

        // ---- layer 3 + arg-max ---------------------------------------------------------------------
        float q[4][OUT];
#pragma unroll
        for (int e = 0; e < 4; ++e)
#pragma unroll
            for (int o = 0; o < OUT; ++o) q[e][o] = 0.f;
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            const float bias = S.b2[ng * NJ + j];
            float w[OUT];
#pragma unroll
            for (int o = 0; o < OUT; ++o) w[o] = S.w3[o][ng * NJ + j];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float h = fmaxf(((j & 1) ? acc[e][j >> 1].y : acc[e][j >> 1].x) + bias, 0.f);
#pragma unroll
                for (int o = 0; o < OUT; ++o) q[e][o] = fmaf(h, w[o], q[e][o]);
            }
        }
#pragma unroll
        for (int e = 0; e < 4; ++e)
#pragma unroll
            for (int o = 0; o < OUT; ++o) {
                float v = q[e][o];
                v += __shfl_xor_sync(0xFFFFFFFFu, v, 1);
                v += __shfl_xor_sync(0xFFFFFFFFu, v, 2);
                q[e][o] = v + S.b3[o];
            }
        if (ng == 0) {
            const int64_t e0 = base + 4 * eg;
            uint8_t a4[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                int best = 0;
                float bv = q[e][0];
#pragma unroll
                for (int o = 1; o < OUT; ++o)
                    if (q[e][o] > bv) { bv = q[e][o]; best = o; }   // first maximum, like torch.max
                a4[e] = (uint8_t)best;
            }
            if (ENV) {
                *reinterpret_cast<uchar4 *>(&S.act_tile[4 * eg]) = make_uchar4(a4[0], a4[1], a4[2], a4[3]);
            } else if (e0 + 4 <= n) {
                uchar4 v = make_uchar4(a4[0], a4[1], a4[2], a4[3]);
                *reinterpret_cast<uchar4 *>(act + e0) = v;
            } else {
                for (int e = 0; e < 4; ++e)
                    if (e0 + e < n) act[e0 + e] = a4[e];
            }
            if (!ENV && (obs_mode & 0x100)) {                   // MG_MLP_FLAG_WRITE_GOAL: slot 0 of the `[goal] + state` rows
                for (int e = 0; e < 4; ++e)
                    if (e0 + e < n) const_cast<float *>(obs)[(e0 + e) * (MG_OBS_DIM + 1)] = (float)a4[e];
            }
            if (q_out) {
#pragma unroll
                for (int e = 0; e < 4; ++e)
                    if (e0 + e < n)
#pragma unroll
                        for (int o = 0; o < OUT; ++o) q_out[(e0 + e) * OUT + o] = q[e][o];
            }
        }
        if (ENV) {
            // ---- env step of the tile: thread t owns env base + t (coalesced state access) -------------------
            __syncthreads();                   // the tile's actions are in shared memory
            mg::StatAcc st;
            mgpe::Loaded x;
            mgpe::load_env<ENV == 2>(P, base + t, n, x);
            mgpe::step_loaded<ENV == 2>(P, base + t, x, (int)S.act_tile[t], st);
            if (P.stats)
                mg::flush_stats(st, P.stats + (size_t)(blockIdx.x % MG_STATS_ROWS) * MG_STATS_COLS, 0xFFFFFFFFu, t & 31);
            // act_tile is rewritten only after the next tile's chunk barriers
        }
    }
}

template <int IN, int OUT, bool MIRROR, int ENV = 0>
cudaError_t launch(const float *obs, const uint8_t *goal, int64_t n, int obs_mode, const float *w1t,
                   const float *b1, const float *w2t, const float *b2, const float *w3, const float *b3,
                   uint8_t *act, float *q_out, cudaStream_t st, const mgpe::Args &P = mgpe::Args{}, bool pdl = false) {
    auto kern = mlp_act_kernel<IN, OUT, MIRROR, ENV>;
    const size_t smem = sizeof(Smem<IN, OUT>);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e) return e;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int64_t tiles = (n + TM - 1) / TM;
    const unsigned grid = (unsigned)(tiles < sms ? tiles : sms);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(TM); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
    e = cudaLaunchKernelEx(&cfg, kern, obs, goal, n, obs_mode, w1t, b1, w2t, b2, w3, b3, act, q_out, P);
    return e ? e : cudaGetLastError();
}

}  // namespace mgmlp

// Shared argument check of mg_mlp_act / mg_mlp_act_tc: flags, observation layout, network input width.
//   obs_dim 10: the network reads the ten observation values (+ the goal column when a goal array is given -> 11 inputs)
//   obs_dim 11: only with MG_MLP_FLAG_OBS_GOAL_SLOT and no goal array — the network reads the whole `[goal] + state` row
int mg_mlp_check_layout(uint32_t flags, int32_t obs_dim, int32_t out_dim, bool has_goal, int *in_dim, int *obs_mode) {
    using namespace mg_abi;
    constexpr uint32_t kAll = MG_MLP_FLAG_MIRROR | MG_MLP_FLAG_PDL | MG_MLP_FLAG_OBS_SOA | MG_MLP_FLAG_OBS_GOAL_SLOT | MG_MLP_FLAG_WRITE_GOAL;
    if (flags & ~kAll) return fail(MG_ERR_BAD_FLAGS, "unknown flag bits");
    const bool soa = (flags & MG_MLP_FLAG_OBS_SOA) != 0u, slot = (flags & MG_MLP_FLAG_OBS_GOAL_SLOT) != 0u;
    if (soa && slot) return fail(MG_ERR_BAD_FLAGS, "MG_MLP_FLAG_OBS_SOA and MG_MLP_FLAG_OBS_GOAL_SLOT exclude each other");
    if ((flags & MG_MLP_FLAG_WRITE_GOAL) && !slot) return fail(MG_ERR_BAD_FLAGS, "MG_MLP_FLAG_WRITE_GOAL needs MG_MLP_FLAG_OBS_GOAL_SLOT");
    if (!(out_dim == 5 || out_dim == 3) || !(obs_dim == MG_OBS_DIM || (obs_dim == MG_OBS_DIM + 1 && slot && !has_goal)))
        return fail(MG_ERR_BAD_SIZE, "the policy kernels support observations of 10 floats (+ optional goal column; or the 11-float "
                                     "`[goal] + state` rows of MG_MLP_FLAG_OBS_GOAL_SLOT) and 5 or 3 outputs "
                                     "(Net(10|11, 5|3): main.py:30-47, hdqn.py:38-55)");
    *in_dim = obs_dim == MG_OBS_DIM + 1 ? MG_OBS_DIM + 1 : MG_OBS_DIM + (has_goal ? 1 : 0);
    if ((flags & MG_MLP_FLAG_WRITE_GOAL) && *in_dim != MG_OBS_DIM)
        return fail(MG_ERR_BAD_FLAGS, "MG_MLP_FLAG_WRITE_GOAL is for the 10-input goal network");
    *obs_mode = (soa ? 1 : slot ? 2 : 0) | ((flags & MG_MLP_FLAG_WRITE_GOAL) ? 0x100 : 0);
    return MG_OK;
}

extern "C" MG_API int mg_mlp_act(const float *obs, const uint8_t *goal_or_null, int64_t n, int32_t obs_dim,
                                 int32_t out_dim, const float *w1t, const float *b1, const float *w2p,
                                 const float *b2, const float *w3, const float *b3, uint8_t *actions,
                                 float *q_out_or_null, uint32_t flags, void *stream) {
    using namespace mg_abi;
    if (n < 0) return fail(MG_ERR_BAD_SIZE, "n < 0");
    int in_dim = 0, obs_mode = 0;
    if (int rc = mg_mlp_check_layout(flags, obs_dim, out_dim, goal_or_null != nullptr, &in_dim, &obs_mode)) return rc;
    const bool mirror = (flags & MG_MLP_FLAG_MIRROR) != 0u, pdl = (flags & MG_MLP_FLAG_PDL) != 0u;
    if (n == 0) return MG_OK;
    if (!obs || !w1t || !b1 || !w2p || !b2 || !w3 || !b3 || !actions)
        return fail(MG_ERR_NULL_POINTER, "mg_mlp_act: NULL pointer");
    if (!aligned16(actions) || !aligned16(obs) || !aligned16(w1t) || !aligned16(w2p) || !aligned16(b1))
        return fail(MG_ERR_ALIGNMENT, "obs, actions and weight arrays must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e;
#define MG_MLP_CASE(I, O) \
    if (in_dim == I && out_dim == O) e = mirror ? mgmlp::launch<I, O, true>(obs, goal_or_null, n, obs_mode, w1t, b1, w2p, b2, w3, b3, actions, q_out_or_null, st, mgpe::Args{}, pdl) : mgmlp::launch<I, O, false>(obs, goal_or_null, n, obs_mode, w1t, b1, w2p, b2, w3, b3, actions, q_out_or_null, st, mgpe::Args{}, pdl); else
    MG_MLP_CASE(10, 5) MG_MLP_CASE(10, 3) MG_MLP_CASE(11, 5) MG_MLP_CASE(11, 3) e = cudaErrorInvalidValue;
#undef MG_MLP_CASE
    if (e) return cuda_fail(e, "mg_mlp_act launch");
    return MG_OK;
}

// ---- mg_policy_step: policy forward + arg-max + exploration + MergeEnv.step in one launch ------------------------
cudaError_t mg_policy_step_tc_launch(int in_dim, const float *obs, const uint8_t *goal, int64_t n, const float *w1t,
                                     const float *b1, const float *w2_tc, const float *b2, const float *w3, const float *b3,
                                     float *q_out, cudaStream_t st, const mgpe::Args &P);
cudaError_t mg_policy_step_tc16_launch(int in_dim, const float *obs, const uint8_t *goal, int64_t n, const void *blob, const float *b2,
                                       const float *w3, const float *b3, float *q_out, cudaStream_t st, const mgpe::Args &P);

namespace mgmlp {
__global__ void __launch_bounds__(256)
explore_kernel(uint8_t *__restrict__ choices, const uint32_t *__restrict__ meta, const int64_t n, const int num_choices,
               const mgpe::Args P) {
    const int64_t e = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (e < n) choices[e] = (uint8_t)mgpe::explore(P, e, (int)choices[e], num_choices, meta ? meta[e] : 0u);
}
}  // namespace mgmlp

extern "C" MG_API int mg_policy_step(const MgState *state, int64_t n, const float *obs_in, const uint8_t *goal_or_null,
                                     int32_t backend, const float *w1t, const float *b1, const float *w2, const float *b2,
                                     const float *w3, const float *b3, const uint8_t *a2_or_null, const MgRewards *rewards,
                                     const MgOut *out, int64_t *stats_or_null, uint32_t flags, const MgResetSpec *reset_or_null,
                                     const MgExplore *explore_or_null, uint8_t *actions_out_or_null, float *q_out_or_null,
                                     void *stream) {
    using namespace mg_abi;
    if (n < 0) return fail(MG_ERR_BAD_SIZE, "n < 0");
    if (flags & ~(MG_FLAG_AUTO_RESET | MG_FLAG_NO_RETURNS | MG_POLICY_FLAG_EXPLORE | MG_POLICY_FLAG_PDL | MG_FLAG_OBS_SOA | MG_FLAG_OBS_GOAL_SLOT |
                  MG_POLICY_FLAG_GOAL_IN_SLOT))
        return fail(MG_ERR_BAD_FLAGS, "unknown flag bits");
    if ((flags & MG_POLICY_FLAG_GOAL_IN_SLOT) && (!(flags & MG_FLAG_OBS_GOAL_SLOT) || goal_or_null))
        return fail(MG_ERR_BAD_FLAGS, "MG_POLICY_FLAG_GOAL_IN_SLOT needs MG_FLAG_OBS_GOAL_SLOT and no goal array");
    if ((flags & MG_FLAG_OBS_SOA) && (flags & MG_FLAG_OBS_GOAL_SLOT)) return fail(MG_ERR_BAD_FLAGS, "MG_FLAG_OBS_SOA and MG_FLAG_OBS_GOAL_SLOT exclude each other");
    const bool pdl = (flags & MG_POLICY_FLAG_PDL) != 0u;
    if (backend != MG_POLICY_BACKEND_FP32 && backend != MG_POLICY_BACKEND_TF32X3 && backend != MG_POLICY_BACKEND_F16X3)
        return fail(MG_ERR_BAD_FLAGS, "backend must be MG_POLICY_BACKEND_FP32, MG_POLICY_BACKEND_TF32X3 or MG_POLICY_BACKEND_F16X3");
    if ((flags & MG_POLICY_FLAG_EXPLORE) && !explore_or_null) return fail(MG_ERR_NULL_POINTER, "MG_POLICY_FLAG_EXPLORE without an MgExplore");
    if (reset_or_null && reset_or_null->mode > MG_RESET_RANDOM) return fail(MG_ERR_BAD_FLAGS, "MgResetSpec.mode must be MG_RESET_FIXED or MG_RESET_RANDOM");
    if (n == 0) return MG_OK;
    if (!state || !out || !obs_in || !w1t || !b1 || !w2 || !b2 || !w3 || !b3) return fail(MG_ERR_NULL_POINTER, "mg_policy_step: NULL pointer");
    if (!state->pos1 || !state->vel1 || !state->pos2 || !state->vel2 || !state->meta) return fail(MG_ERR_NULL_POINTER, "a state array pointer is NULL");
    if (!out->obs || !out->rew || !out->info) return fail(MG_ERR_NULL_POINTER, "out.obs/rew/info must be non-NULL");
    const bool ret = state->ret1 && state->ret2 && !(flags & MG_FLAG_NO_RETURNS);
    if (!ret && out->ep_ret) return fail(MG_ERR_BAD_FLAGS, "out.ep_ret needs the return accumulators");
    if (!aligned16(obs_in) || !aligned16(out->obs) || !aligned16(out->rew) || !aligned16(w1t) || !aligned16(w2) || !aligned16(b1))
        return fail(MG_ERR_ALIGNMENT, "obs, out and weight arrays must be 16-byte aligned");
    mgpe::Args P{};
    P.s = *state;
    if (!ret) P.s.ret1 = P.s.ret2 = nullptr;
    P.o = *out;
    P.a2 = a2_or_null;
    P.actions = actions_out_or_null;
    P.stats = reinterpret_cast<unsigned long long *>(stats_or_null);
    if (rewards) P.rw = *rewards; else mg_default_rewards(&P.rw);
    P.rs = reset_or_null ? *reset_or_null : MgResetSpec{MG_RESET_FIXED, 0u, 0ull, 0ull};
    P.flags = flags;
    P.n = n;
    if (explore_or_null) { P.explore_seed = explore_or_null->seed; P.explore_step = explore_or_null->step; P.explore_keep = explore_or_null->keep_u32; }
    // network input: 10 observation values, + 1 when a goal array is given or MG_POLICY_FLAG_GOAL_IN_SLOT asks for the
    // whole `[goal] + state` row of the MG_FLAG_OBS_GOAL_SLOT layout
    const bool row11 = (flags & MG_POLICY_FLAG_GOAL_IN_SLOT) != 0u;
    const int in_dim = MG_OBS_DIM + ((goal_or_null || row11) ? 1 : 0);
    const int obs_mode = (int)mg::obs_layout_of(flags);
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e;
    if (backend == MG_POLICY_BACKEND_TF32X3)
        e = mg_policy_step_tc_launch(in_dim, obs_in, goal_or_null, n, w1t, b1, w2, b2, w3, b3, q_out_or_null, st, P);
    else if (backend == MG_POLICY_BACKEND_F16X3)        // w2 = the packed fp16 operand blob (w1t / b1 travel inside it)
        e = mg_policy_step_tc16_launch(in_dim, obs_in, goal_or_null, n, w2, b2, w3, b3, q_out_or_null, st, P);
    else if (in_dim == 10)
        e = a2_or_null ? mgmlp::launch<10, 5, false, 2>(obs_in, goal_or_null, n, obs_mode, w1t, b1, w2, b2, w3, b3, nullptr, q_out_or_null, st, P, pdl)
                       : mgmlp::launch<10, 5, false, 1>(obs_in, goal_or_null, n, obs_mode, w1t, b1, w2, b2, w3, b3, nullptr, q_out_or_null, st, P, pdl);
    else
        e = a2_or_null ? mgmlp::launch<11, 5, false, 2>(obs_in, goal_or_null, n, obs_mode, w1t, b1, w2, b2, w3, b3, nullptr, q_out_or_null, st, P, pdl)
                       : mgmlp::launch<11, 5, false, 1>(obs_in, goal_or_null, n, obs_mode, w1t, b1, w2, b2, w3, b3, nullptr, q_out_or_null, st, P, pdl);
    if (e) return cuda_fail(e, "mg_policy_step launch");
    return MG_OK;
}

extern "C" MG_API int mg_explore(uint8_t *choices, int64_t n, int32_t num_choices, const MgExplore *explore,
                                 const uint32_t *meta_or_null, uint64_t env_id_base, uint32_t salt, void *stream) {
    using namespace mg_abi;
    if (n < 0 || num_choices < 1 || num_choices > 255) return fail(MG_ERR_BAD_SIZE, "n < 0 or num_choices outside 1..255");
    if (n == 0) return MG_OK;
    if (!choices || !explore) return fail(MG_ERR_NULL_POINTER, "mg_explore: NULL pointer");
    mgpe::Args P{};
    P.flags = MG_POLICY_FLAG_EXPLORE;
    P.rs.env_id_base = env_id_base;
    P.explore_seed = explore->seed ^ ((uint64_t)salt << 20);
    P.explore_step = explore->step;
    P.explore_keep = explore->keep_u32;
    mgmlp::explore_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(choices, meta_or_null, n, num_choices, P);
    if (cudaError_t e = cudaGetLastError()) return cuda_fail(e, "mg_explore launch");
    return MG_OK;
}
