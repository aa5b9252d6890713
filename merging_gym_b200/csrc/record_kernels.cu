// record_kernels.cu — device-resident transition writer ("next" row: SURVEY.md 8f-2 / 8f-3).
//
// The reference's learners keep a ring of rows `[s(10), a, r, s'(10)]` filled by
// `store_transition` (scripts/main.py:115-119) for every step while `env.winner is not 1`
// (main.py:209-211); human_player.py logs `[s(10), a1, a2, r1, r2]` under the same condition
// (:111, :180-181).  For N envs stepped together this is a stream compaction: the rows of the envs
// that pass the mask are appended to the ring in env-id order, `index = counter % capacity`.
//
// Deterministic three-pass compaction (no atomics, so the ring content is reproducible and equals a
// sequential loop over env ids):
//   count  : ballot + popc per warp, summed per 256-env block          -> block_counts
//   scan   : one CTA, exclusive scan of block_counts (n / 256 entries), advances the ring counter
//   write  : a block re-derives its warps' ranks from the ballots; every selected env writes its row at
//            (counter + offset) % capacity
#include "abi_common.h"

namespace mgrec {

constexpr int kBlock = 256;
constexpr int kObs = MG_OBS_DIM;

__device__ __forceinline__ bool selected(const uint8_t *info, int64_t e, int mask_mode) {
    // mask_mode 0: every env; 1: the reference's `env.winner is not 1` (winner AFTER the step);
    // 2: explicit — `info` is a caller-supplied byte mask, non-zero = store
    if (mask_mode == 0) return true;
    if (mask_mode == 2) return info[e] != 0u;
    return ((info[e] & MG_INFO_WINNER_MASK) >> MG_INFO_WINNER_SHIFT) != 1u;
}

__global__ void __launch_bounds__(kBlock)
count_kernel(const uint8_t *__restrict__ info, int64_t n, int mask_mode, uint32_t *__restrict__ block_counts) {
    __shared__ uint32_t s_cnt[kBlock / 32];
    const int64_t e = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    const bool sel = e < n && selected(info, e, mask_mode);
    const unsigned b = __ballot_sync(0xFFFFFFFFu, sel);
    if ((threadIdx.x & 31) == 0) s_cnt[threadIdx.x >> 5] = __popc(b);
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t tot = 0;
#pragma unroll
        for (int w = 0; w < kBlock / 32; ++w) tot += s_cnt[w];
        block_counts[blockIdx.x] = tot;
    }
}

// Single CTA: in-place exclusive scan of block_counts[0..m), total added to *counter (int64 rows
// written so far); the pre-increment value is left in *base for the write pass.
__global__ void __launch_bounds__(1024)
scan_kernel(uint32_t *__restrict__ warp_counts, int64_t m, unsigned long long *__restrict__ counter,
            unsigned long long *__restrict__ base) {
    __shared__ uint32_t warp_tot[32];
    __shared__ uint32_t carry_s;
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    if (t == 0) carry_s = 0;
    __syncthreads();
    for (int64_t i0 = 0; i0 < m; i0 += 4096) {             // 4 entries per thread and pass
        const int64_t i = i0 + 4 * t;
        uint32_t v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) v[k] = i + k < m ? warp_counts[i + k] : 0u;
        const uint32_t mine = v[0] + v[1] + v[2] + v[3];
        uint32_t x = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) warp_tot[w] = x;
        __syncthreads();
        if (w == 0) {
            uint32_t s = warp_tot[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, s, o);
                if (lane >= o) s += y;
            }
            warp_tot[lane] = s;                    // inclusive totals of the 32 warps
        }
        __syncthreads();
        const uint32_t carry = carry_s;
        uint32_t before = carry + (w ? warp_tot[w - 1] : 0u) + (x - mine);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (i + k < m) warp_counts[i + k] = before;
            before += v[k];
        }
        __syncthreads();
        if (t == 1023) carry_s = carry + warp_tot[31];
        __syncthreads();
    }
    if (t == 0) {
        *base = *counter;
        *counter += carry_s;
    }
}

// format 0 (replay, main.py:115-119): [s(10), a_p, r_p, s'(10)]              22 floats, player p
// format 1 (log, human_player.py:111) : [s(10), a1, a2, r1, r2]               14 floats
// format 2 (h-DQN controller, hdqn.py:180-184,291-316): [g, s(10), a, r_int, g', s'(10)]   24 floats, with
//          r_int = 1 if g' == goal_status(s) else 0 (hdqn.py:314; goal_status :223-236 on the state the action was
//          chosen from, g' the goal re-chosen from the next state)
// One warp handles 32 consecutive envs: their observation rows are one contiguous 1280-byte span, loaded
// with coalesced 128-bit loads into shared memory; the selected lanes assemble their rows in shared memory
// at consecutive ranks; the warp then writes that contiguous piece of the ring with coalesced stores.
template <int FORMAT>
__global__ void __launch_bounds__(kBlock)
write_kernel(const float *__restrict__ obs_prev, const float *__restrict__ obs_next,
             const float *__restrict__ term_obs, const uint8_t *__restrict__ a1, const uint8_t *__restrict__ a2,
             const float *__restrict__ rew, const uint8_t *__restrict__ done, const uint8_t *__restrict__ info,
             const uint8_t *__restrict__ goal_prev, const uint8_t *__restrict__ goal_next,
             int64_t n, int mask_mode, int player, const uint32_t *__restrict__ block_offsets,
             const unsigned long long *__restrict__ base, const unsigned long long *__restrict__ counter,
             float *__restrict__ ring, int64_t capacity, int32_t *__restrict__ env_ids) {
    constexpr int WIDTH = FORMAT == 0 ? 2 * kObs + 2 : FORMAT == 1 ? kObs + 4 : 2 * kObs + 4;
    constexpr int kWarps = kBlock / 32;
    __shared__ __align__(16) float s_prev[kWarps][32 * kObs];
    __shared__ __align__(16) float s_next[kWarps][FORMAT != 1 ? 32 * kObs : 4];
    __shared__ __align__(16) float s_out[kWarps][32 * WIDTH];
    __shared__ uint32_t s_cnt[kWarps];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t w0 = ((int64_t)blockIdx.x * kBlock + warp * 32);          // first env of this warp
    const int64_t e = w0 + lane;
    const bool sel = e < n && selected(info, e, mask_mode);
    const unsigned b = __ballot_sync(0xFFFFFFFFu, sel);
    const int cnt = __popc(b);
    if (lane == 0) s_cnt[warp] = (uint32_t)cnt;
    __syncthreads();                                                         // the only block-wide step
    if (w0 >= n) return;
    uint32_t before = 0;                                                     // selected envs of the lower warps of this block
#pragma unroll
    for (int w = 0; w < kWarps; ++w) before += w < warp ? s_cnt[w] : 0u;
    const int rows = (int)min((int64_t)32, n - w0);
    // ---- coalesced loads of the warp's observation rows (w0 * 40 bytes is 16-byte aligned: w0 % 32 == 0) ----
    {
        const float4 *gp = reinterpret_cast<const float4 *>(obs_prev + w0 * kObs);
        float4 *sp = reinterpret_cast<float4 *>(s_prev[warp]);
        for (int i = lane; i < rows * kObs / 4; i += 32) sp[i] = __ldg(gp + i);
        for (int i = (rows * kObs / 4) * 4 + lane; i < rows * kObs; i += 32) s_prev[warp][i] = obs_prev[w0 * kObs + i];
        if (FORMAT != 1) {
            const float4 *gn = reinterpret_cast<const float4 *>(obs_next + w0 * kObs);
            float4 *sn = reinterpret_cast<float4 *>(s_next[warp]);
            for (int i = lane; i < rows * kObs / 4; i += 32) sn[i] = __ldg(gn + i);
            for (int i = (rows * kObs / 4) * 4 + lane; i < rows * kObs; i += 32) s_next[warp][i] = obs_next[w0 * kObs + i];
        }
    }
    __syncwarp();
    const uint64_t rank0 = (uint64_t)block_offsets[blockIdx.x] + before;
    const uint64_t total = *counter - *base;
    if (sel) {
        const int r = __popc(b & ((1u << lane) - 1u));
        float *row = s_out[warp] + r * WIDTH;
        const float *sp = s_prev[warp] + lane * kObs;
        const float act1 = (float)a1[e], act2 = a2 ? (float)a2[e] : 0.f;
        // s' is the observation of the stepped state: under auto-reset obs_next holds the RESET
        // observation for finished envs, the terminal one is in term_obs
        const bool use_term = FORMAT != 1 && term_obs && done[e];
        const float *sn = use_term ? term_obs + e * kObs : s_next[warp] + lane * kObs;
        if (FORMAT == 2) {
            const float g = (float)goal_prev[e], gn = (float)goal_next[e];
            const float dx1 = sp[0], v2 = sp[9];                                  // goal_status, hdqn.py:223-236
            const float status = dx1 < -0.5f * v2 ? 0.f : dx1 < 0.5f * v2 ? 1.f : 2.f;
            row[0] = g;
#pragma unroll
            for (int k = 0; k < kObs; ++k) row[1 + k] = sp[k];
            row[kObs + 1] = player == 2 ? act2 : act1;
            row[kObs + 2] = gn == status ? 1.f : 0.f;
            row[kObs + 3] = gn;
#pragma unroll
            for (int k = 0; k < kObs; ++k) row[kObs + 4 + k] = sn[k];
        } else {
#pragma unroll
            for (int k = 0; k < kObs; ++k) row[k] = sp[k];
            if (FORMAT == 0) {
                row[kObs] = player == 2 ? act2 : act1;
                row[kObs + 1] = rew[2 * e + (player == 2 ? 1 : 0)];
#pragma unroll
                for (int k = 0; k < kObs; ++k) row[kObs + 2 + k] = sn[k];
            } else {
                row[kObs] = act1; row[kObs + 1] = act2;
                row[kObs + 2] = rew[2 * e]; row[kObs + 3] = rew[2 * e + 1];
            }
        }
        // rows a sequential writer would overwrite later in this same call are skipped below via `skip`
        if (env_ids && !(rank0 + r + (uint64_t)capacity < total))
            env_ids[(int64_t)((*base + rank0 + r) % (unsigned long long)capacity)] = (int32_t)e;
    }
    __syncwarp();
    // ---- coalesced write of the warp's cnt rows: ring positions (base + rank0 + r) % capacity ----------
    const unsigned long long cap = (unsigned long long)capacity;
    const unsigned long long slot0 = (*base + rank0) % cap;
    if (total <= cap && slot0 + (unsigned long long)cnt <= cap) {
        // common case: one contiguous span, 8-byte aligned (row width 88 or 56 bytes) -> 64-bit stores
        const float2 *src = reinterpret_cast<const float2 *>(s_out[warp]);
        float2 *dst = reinterpret_cast<float2 *>(ring + slot0 * WIDTH);
        for (int i = lane; i < cnt * WIDTH / 2; i += 32) dst[i] = src[i];
    } else {
        for (int i = lane; i < cnt * WIDTH; i += 32) {
            const int r = i / WIDTH, c = i - r * WIDTH;
            if (rank0 + r + (uint64_t)capacity < total) continue;          // overwritten later in this call
            const unsigned long long slot = (slot0 + (unsigned long long)r) % cap;
            ring[slot * WIDTH + c] = s_out[warp][i];
        }
    }
}

}  // namespace mgrec

extern "C" MG_API int mg_record_transitions(const float *obs_prev, const float *obs_next,
                                            const float *term_obs_or_null, const uint8_t *a1,
                                            const uint8_t *a2_or_null, const float *rew, const uint8_t *done,
                                            const uint8_t *info, const uint8_t *goal_prev_or_null,
                                            const uint8_t *goal_next_or_null, int64_t n, int32_t mask_mode, int32_t format,
                                            int32_t player, float *ring, int64_t capacity,
                                            int32_t *env_ids_or_null, uint64_t *counter, uint32_t *scratch,
                                            void *stream) {
    using namespace mg_abi;
    if (n < 0 || capacity <= 0) return fail(MG_ERR_BAD_SIZE, "n < 0 or capacity <= 0");
    if (mask_mode < 0 || mask_mode > 2 || format < 0 || format > 2 || player < 1 || player > 2)
        return fail(MG_ERR_BAD_FLAGS, "mask_mode in {0,1,2}, format in {0,1,2}, player in {1,2}");
    if (n == 0) return MG_OK;
    if (!obs_prev || !obs_next || !a1 || !rew || !done || !info || !ring || !counter || !scratch)
        return fail(MG_ERR_NULL_POINTER, "mg_record_transitions: NULL pointer");
    if (format == 2 && (!goal_prev_or_null || !goal_next_or_null))
        return fail(MG_ERR_NULL_POINTER, "format 2 (h-DQN rows) needs goal_prev and goal_next");
    if (!aligned16(obs_prev) || !aligned16(obs_next))
        return fail(MG_ERR_ALIGNMENT, "obs_prev and obs_next must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned grid = (unsigned)((n + mgrec::kBlock - 1) / mgrec::kBlock);
    const int64_t m = (n + 31) / 32;
    // scratch: uint32[m + 2] (the documented size) — the first `grid` entries hold the block counts / offsets, the
    // 64-bit pre-increment counter sits behind entry m (8-byte aligned)
    uint32_t *warp_counts = scratch;
    auto *base = reinterpret_cast<unsigned long long *>(scratch + ((m + 1) & ~(int64_t)1));
    mgrec::count_kernel<<<grid, mgrec::kBlock, 0, st>>>(info, n, mask_mode, warp_counts);
    mgrec::scan_kernel<<<1, 1024, 0, st>>>(warp_counts, (int64_t)grid, reinterpret_cast<unsigned long long *>(counter), base);
#define MG_REC_LAUNCH(F)                                                                                            \
    mgrec::write_kernel<F><<<grid, mgrec::kBlock, 0, st>>>(obs_prev, obs_next, term_obs_or_null, a1, a2_or_null, rew, done, \
                                                           info, goal_prev_or_null, goal_next_or_null, n, mask_mode, player, \
                                                           warp_counts, base, reinterpret_cast<unsigned long long *>(counter), \
                                                           ring, capacity, env_ids_or_null)
    if (format == 0) MG_REC_LAUNCH(0);
    else if (format == 1) MG_REC_LAUNCH(1);
    else MG_REC_LAUNCH(2);
#undef MG_REC_LAUNCH
    if (cudaError_t e = cudaGetLastError()) return cuda_fail(e, "mg_record_transitions launch");
    return MG_OK;
}

// ---- h-DQN meta-controller bookkeeping (scripts/hdqn.py:283-320), one thread per env -------------------------------
// An option runs from one goal choice until `done or goal == goal_status(state)` (hdqn.py:316); on the way the ego
// rewards are summed into `extrinsic_reward` (:312).  Per env and step this kernel (1) picks the observation the option
// would end in — the stepped state's: the terminal observation where the env finished and was auto-reset —, (2) adds
// the step's ego reward to the running sum, (3) decides whether the option ended, (4) hands the recorder what
// `upper.store_transition(state, goal, extrinsic_reward, next_state)` (:318) needs: the row's observation, the sum in
// column 0 of a reward pair, the ended mask; and (5) restarts the sum of an ended option.
namespace mgrec {
__global__ void __launch_bounds__(kBlock)
option_update_kernel(const float *__restrict__ obs, const float *__restrict__ term_obs, const float *__restrict__ rew,
                     const uint8_t *__restrict__ done, const uint8_t *__restrict__ goal_next, const int64_t n,
                     float *__restrict__ extrinsic, float *__restrict__ s_end, float *__restrict__ rew_out,
                     uint8_t *__restrict__ ended) {
    const int64_t e = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    if (e >= n) return;
    const bool dn = done[e] != 0;
    const float2 *src = reinterpret_cast<const float2 *>((dn && term_obs ? term_obs : obs) + e * kObs);
    float2 *dst = reinterpret_cast<float2 *>(s_end + e * kObs);
    float2 row[kObs / 2];
#pragma unroll
    for (int k = 0; k < kObs / 2; ++k) { row[k] = src[k]; dst[k] = row[k]; }
    const float dx1 = row[0].x, v2 = row[kObs / 2 - 1].y;                       // goal_status, hdqn.py:223-236
    const uint8_t status = dx1 < -0.5f * v2 ? 0 : dx1 < 0.5f * v2 ? 1 : 2;
    const float sum = extrinsic[e] + rew[2 * e];                                // extrinsic_reward += reward (:312)
    const bool end = dn || goal_next[e] == status;                              // :316
    *reinterpret_cast<float2 *>(rew_out + 2 * e) = make_float2(sum, 0.f);
    ended[e] = end ? 1 : 0;
    extrinsic[e] = end ? 0.f : sum;
}
}  // namespace mgrec

extern "C" MG_API int mg_option_update(const float *obs, const float *term_obs_or_null, const float *rew, const uint8_t *done,
                                       const uint8_t *goal_next, int64_t n, float *extrinsic, float *s_end_out,
                                       float *rew_out, uint8_t *ended_out, void *stream) {
    using namespace mg_abi;
    if (n < 0) return fail(MG_ERR_BAD_SIZE, "n < 0");
    if (n == 0) return MG_OK;
    if (!obs || !rew || !done || !goal_next || !extrinsic || !s_end_out || !rew_out || !ended_out)
        return fail(MG_ERR_NULL_POINTER, "mg_option_update: NULL pointer");
    if (!aligned16(obs) || !aligned16(s_end_out) || !aligned16(rew_out) || (term_obs_or_null && !aligned16(term_obs_or_null)))
        return fail(MG_ERR_ALIGNMENT, "observation and reward arrays must be 16-byte aligned");
    const unsigned grid = (unsigned)((n + mgrec::kBlock - 1) / mgrec::kBlock);
    mgrec::option_update_kernel<<<grid, mgrec::kBlock, 0, (cudaStream_t)stream>>>(obs, term_obs_or_null, rew, done, goal_next, n,
                                                                                  extrinsic, s_end_out, rew_out, ended_out);
    if (cudaError_t e = cudaGetLastError()) return cuda_fail(e, "mg_option_update launch");
    return MG_OK;
}
