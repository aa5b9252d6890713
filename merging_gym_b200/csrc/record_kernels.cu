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
//   count  : one warp ballot + popc per 32 envs  -> warp_counts
//   scan   : one CTA, exclusive scan of warp_counts (<= 2^20 entries), advances the ring counter
//   write  : every selected env writes its row at (counter + offset) % capacity
#include "abi_common.h"

namespace mgrec {

constexpr int kBlock = 256;
constexpr int kObs = MG_OBS_DIM;

__device__ __forceinline__ bool selected(const uint8_t *info, int64_t e, int mask_mode) {
    // mask_mode 0: every env; 1: the reference's `env.winner is not 1` (winner AFTER the step)
    if (mask_mode == 0) return true;
    return ((info[e] & MG_INFO_WINNER_MASK) >> MG_INFO_WINNER_SHIFT) != 1u;
}

__global__ void __launch_bounds__(kBlock)
count_kernel(const uint8_t *__restrict__ info, int64_t n, int mask_mode, uint32_t *__restrict__ warp_counts) {
    const int64_t e = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    const bool sel = e < n && selected(info, e, mask_mode);
    const unsigned b = __ballot_sync(0xFFFFFFFFu, sel);
    if ((threadIdx.x & 31) == 0 && e < n) warp_counts[e >> 5] = __popc(b);
}

// Single CTA: in-place exclusive scan of warp_counts[0..m), total added to *counter (int64 rows
// written so far); the pre-increment value is left in *base for the write pass.
__global__ void __launch_bounds__(1024)
scan_kernel(uint32_t *__restrict__ warp_counts, int64_t m, unsigned long long *__restrict__ counter,
            unsigned long long *__restrict__ base) {
    __shared__ uint32_t warp_tot[32];
    __shared__ uint32_t carry_s;
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    if (t == 0) carry_s = 0;
    __syncthreads();
    for (int64_t i0 = 0; i0 < m; i0 += 1024) {
        const int64_t i = i0 + t;
        const uint32_t v = i < m ? warp_counts[i] : 0u;
        uint32_t x = v;
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
        const uint32_t before = carry + (w ? warp_tot[w - 1] : 0u) + (x - v);
        if (i < m) warp_counts[i] = before;
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
__global__ void __launch_bounds__(kBlock)
write_kernel(const float *__restrict__ obs_prev, const float *__restrict__ obs_next,
             const float *__restrict__ term_obs, const uint8_t *__restrict__ a1, const uint8_t *__restrict__ a2,
             const float *__restrict__ rew, const uint8_t *__restrict__ done, const uint8_t *__restrict__ info,
             int64_t n, int mask_mode, int format, int player, const uint32_t *__restrict__ warp_offsets,
             const unsigned long long *__restrict__ base, const unsigned long long *__restrict__ counter,
             float *__restrict__ ring, int64_t capacity, int32_t *__restrict__ env_ids) {
    const int64_t e = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    const bool sel = e < n && selected(info, e, mask_mode);
    const unsigned b = __ballot_sync(0xFFFFFFFFu, sel);
    if (!sel) return;
    const int lane = threadIdx.x & 31;
    const uint64_t rank = warp_offsets[e >> 5] + __popc(b & ((1u << lane) - 1u));
    // more selected rows than the ring holds: rows a sequential writer would overwrite are skipped
    if (rank + (uint64_t)capacity < *counter - *base) return;
    const int64_t slot = (int64_t)((*base + rank) % (unsigned long long)capacity);
    const int width = format == 0 ? 2 * kObs + 2 : kObs + 4;
    float *row = ring + slot * width;
    const float *s = obs_prev + e * kObs;
#pragma unroll
    for (int k = 0; k < kObs; ++k) row[k] = s[k];
    const float act1 = (float)a1[e], act2 = a2 ? (float)a2[e] : 0.f;
    if (format == 0) {
        row[kObs] = player == 2 ? act2 : act1;
        row[kObs + 1] = rew[2 * e + (player == 2 ? 1 : 0)];
        // s' is the observation of the stepped state: under auto-reset obs_next holds the RESET
        // observation for finished envs, the terminal one is in term_obs
        const float *sn = (term_obs && done[e]) ? term_obs + e * kObs : obs_next + e * kObs;
#pragma unroll
        for (int k = 0; k < kObs; ++k) row[kObs + 2 + k] = sn[k];
    } else {
        row[kObs] = act1; row[kObs + 1] = act2;
        row[kObs + 2] = rew[2 * e]; row[kObs + 3] = rew[2 * e + 1];
    }
    if (env_ids) env_ids[slot] = (int32_t)e;
}

}  // namespace mgrec

extern "C" MG_API int mg_record_transitions(const float *obs_prev, const float *obs_next,
                                            const float *term_obs_or_null, const uint8_t *a1,
                                            const uint8_t *a2_or_null, const float *rew, const uint8_t *done,
                                            const uint8_t *info, int64_t n, int32_t mask_mode, int32_t format,
                                            int32_t player, float *ring, int64_t capacity,
                                            int32_t *env_ids_or_null, uint64_t *counter, uint32_t *scratch,
                                            void *stream) {
    using namespace mg_abi;
    if (n < 0 || capacity <= 0) return fail(MG_ERR_BAD_SIZE, "n < 0 or capacity <= 0");
    if (mask_mode < 0 || mask_mode > 1 || format < 0 || format > 1 || player < 1 || player > 2)
        return fail(MG_ERR_BAD_FLAGS, "mask_mode in {0,1}, format in {0,1}, player in {1,2}");
    if (n == 0) return MG_OK;
    if (!obs_prev || !obs_next || !a1 || !rew || !done || !info || !ring || !counter || !scratch)
        return fail(MG_ERR_NULL_POINTER, "mg_record_transitions: NULL pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned grid = (unsigned)((n + mgrec::kBlock - 1) / mgrec::kBlock);
    const int64_t m = (n + 31) / 32;
    // scratch: uint32[m + 2] — warp counts/offsets followed by the 64-bit pre-increment counter (8-byte aligned)
    uint32_t *warp_counts = scratch;
    auto *base = reinterpret_cast<unsigned long long *>(scratch + ((m + 1) & ~(int64_t)1));
    mgrec::count_kernel<<<grid, mgrec::kBlock, 0, st>>>(info, n, mask_mode, warp_counts);
    mgrec::scan_kernel<<<1, 1024, 0, st>>>(warp_counts, m, reinterpret_cast<unsigned long long *>(counter), base);
    mgrec::write_kernel<<<grid, mgrec::kBlock, 0, st>>>(obs_prev, obs_next, term_obs_or_null, a1, a2_or_null, rew,
                                                         done, info, n, mask_mode, format, player, warp_counts, base,
                                                         reinterpret_cast<unsigned long long *>(counter), ring,
                                                         capacity, env_ids_or_null);
    if (cudaError_t e = cudaGetLastError()) return cuda_fail(e, "mg_record_transitions launch");
    return MG_OK;
}
