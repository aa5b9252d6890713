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
//   count  : every thread takes 16 consecutive info bytes (one 128-bit load), half-warps sum them into the counts of
//            the 256-env blocks                                            -> block_counts
//   scan   : one CTA, exclusive scan of block_counts (n / 256 entries), advances the ring counter
//   write  : a block re-derives its warps' ranks from the ballots; every selected env writes its row at
//            (counter + offset) % capacity
// The write pass is launched as a programmatic dependent of the scan: its blocks request their observation rows and take
// their ballots while the single scan CTA is still running, and only then wait for the offsets (47 -> 44 us per 2^20 rows
// inside a CUDA graph; 37.6 us with the aliased staging buffer of RowSmem = 4.9 TB/s, 0.75 of the HBM copy peak).
// A one-launch variant (tiles by ticket, decoupled look-back over per-tile words) was measured and dropped: 57 us against
// 47 (profiles/r02_record_onepass_experiment.patch, r02_record_onepass_vs_three_pass.jsonl) — the ticket, publication and
// look-back round trips put ~5 us of latency in front of every block's first store.
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

// One thread = 16 consecutive envs (their info bytes are one 16-byte load), 16 threads = one 256-env block of the write
// pass: a half-warp shuffle sum gives that block's count.  The ragged tail (and a misaligned info pointer) goes byte by byte.
constexpr int kCountPerThread = 16;
__global__ void __launch_bounds__(kBlock)
count_kernel(const uint8_t *__restrict__ info, int64_t n, int mask_mode, uint32_t *__restrict__ block_counts) {
    const int64_t t = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    const int64_t e0 = t * kCountPerThread;
    uint32_t c = 0;
    if (e0 < n) {
        if (mask_mode == 0) {
            c = (uint32_t)min((int64_t)kCountPerThread, n - e0);
        } else if (e0 + kCountPerThread <= n && (reinterpret_cast<uintptr_t>(info) & 15u) == 0) {
            const uint4 v = __ldg(reinterpret_cast<const uint4 *>(info + e0));
            const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (mask_mode == 2) {
                    // bytes != 0: fold each byte's bits into its bit 0
                    uint32_t x = w[k];
                    x |= x >> 4; x |= x >> 2; x |= x >> 1;
                    c += __popc(x & 0x01010101u);
                } else {
                    // winner field != 1 in every byte: XOR with the pattern "1", then any bit left in the field
                    constexpr uint32_t m = MG_INFO_WINNER_MASK * 0x01010101u;
                    constexpr uint32_t one = (1u << MG_INFO_WINNER_SHIFT) * 0x01010101u;
                    uint32_t x = ((w[k] & m) ^ one) >> MG_INFO_WINNER_SHIFT;      // field value ^ 1, per byte, at bit 0 ..
                    x |= x >> 4; x |= x >> 2; x |= x >> 1;
                    c += __popc(x & 0x01010101u);
                }
            }
        } else {
            for (int64_t e = e0; e < min(n, e0 + kCountPerThread); ++e) c += selected(info, e, mask_mode) ? 1u : 0u;
        }
    }
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) c += __shfl_xor_sync(0xFFFFFFFFu, c, o);      // 16 threads = 256 envs
    if ((threadIdx.x & 15) == 0 && e0 < n) block_counts[t / 16] = c;
}

// Single CTA: in-place exclusive scan of block_counts[0..m), total added to *counter (int64 rows
// written so far); the pre-increment value is left in *base for the write pass.
__global__ void __launch_bounds__(1024)
scan_kernel(uint32_t *__restrict__ warp_counts, int64_t m, unsigned long long *__restrict__ counter,
            unsigned long long *__restrict__ base) {
    __shared__ uint32_t warp_tot[32];
    __shared__ uint32_t carry_s;
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    // the write pass may become resident now (on the other SMs): it loads its rows and waits for this grid's completion
    // before it reads the offsets
    cudaTriggerProgrammaticLaunchCompletion();
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
template <int FORMAT> struct RowFmt {
    static constexpr int WIDTH = FORMAT == 0 ? 2 * kObs + 2 : FORMAT == 1 ? kObs + 4 : 2 * kObs + 4;
};
constexpr int kWarps = kBlock / 32;

// Per warp ONE staging buffer: first the 32 observation rows s (and s' behind them), then — once every selected lane holds
// its assembled row in registers — the rows at their ranks (22.5 KB per block instead of 43: eight resident blocks per SM).
template <int FORMAT>
struct RowSmem {
    static constexpr int IN_FLOATS = (FORMAT != 1 ? 2 : 1) * 32 * kObs, OUT_FLOATS = 32 * RowFmt<FORMAT>::WIDTH;
    float buf[kWarps][IN_FLOATS > OUT_FLOATS ? IN_FLOATS : OUT_FLOATS];
    __device__ __forceinline__ float *prev(int w) { return buf[w]; }
    __device__ __forceinline__ float *next(int w) { return buf[w] + 32 * kObs; }
    __device__ __forceinline__ float *out(int w) { return buf[w]; }
};

struct RowArgs {
    const float *obs_prev, *obs_next, *term_obs;
    const uint8_t *a1, *a2;
    const float *rew;
    const uint8_t *done, *info, *goal_prev, *goal_next;
    int64_t n;
    int mask_mode, player;
    float *ring;
    int64_t capacity;
    int32_t *env_ids;
};

// Part 0: a full warp's 32 observation rows (s, and s' for the replay formats) are 80 float4 each: every lane requests
// its 2-3 pieces of both arrays before anything waits on them, so that a warp has up to six 128-bit loads in flight
// (the first version looped load -> shared store and kept one: 57 us per 2^20 rows).
struct RowRegs { float4 p[3], q[3]; };
template <int FORMAT>
__device__ __forceinline__ void request_rows(const RowArgs &A, int lane, int64_t w0, RowRegs &R) {
    const float4 *gp = reinterpret_cast<const float4 *>(A.obs_prev + w0 * kObs);
    R.p[0] = __ldg(gp + lane); R.p[1] = __ldg(gp + lane + 32);
    if (lane < 16) R.p[2] = __ldg(gp + lane + 64);
    if (FORMAT != 1) {
        const float4 *gn = reinterpret_cast<const float4 *>(A.obs_next + w0 * kObs);
        R.q[0] = __ldg(gn + lane); R.q[1] = __ldg(gn + lane + 32);
        if (lane < 16) R.q[2] = __ldg(gn + lane + 64);
    }
}

// Part 1: the warp's observation rows into shared memory (coalesced 128-bit accesses; w0 * 40 bytes is 16-byte
// aligned because w0 % 32 == 0), then every selected lane assembles its row at its rank inside the warp.  `full` = the
// warp has all 32 envs and R holds their rows (request_rows); the last, ragged warp loads them here.
template <int FORMAT>
__device__ __forceinline__ void assemble_rows(const RowArgs &A, RowSmem<FORMAT> &sm, int warp, int lane, int64_t w0,
                                              bool sel, unsigned b, bool full, const RowRegs &R) {
    constexpr int WIDTH = RowFmt<FORMAT>::WIDTH;
    const int64_t e = w0 + lane;
    // the selected lanes' own scalars: requested before the staging below waits on the row loads
    float act1 = 0.f, act2 = 0.f, g = 0.f, gn = 0.f;
    float2 rw = make_float2(0.f, 0.f);
    bool use_term = false;
    if (sel) {
        act1 = (float)A.a1[e];
        act2 = A.a2 ? (float)A.a2[e] : 0.f;
        if (FORMAT != 2) rw = __ldg(reinterpret_cast<const float2 *>(A.rew) + e);
        // s' is the observation of the stepped state: under auto-reset obs_next holds the RESET
        // observation for finished envs, the terminal one is in term_obs
        use_term = FORMAT != 1 && A.term_obs && A.done[e];
        if (FORMAT == 2) { g = (float)A.goal_prev[e]; gn = (float)A.goal_next[e]; }
    }
    if (full) {
        float4 *sp = reinterpret_cast<float4 *>(sm.prev(warp));
        sp[lane] = R.p[0]; sp[lane + 32] = R.p[1];
        if (lane < 16) sp[lane + 64] = R.p[2];
        if (FORMAT != 1) {
            float4 *sn = reinterpret_cast<float4 *>(sm.next(warp));
            sn[lane] = R.q[0]; sn[lane + 32] = R.q[1];
            if (lane < 16) sn[lane + 64] = R.q[2];
        }
    } else {
        const int rows = (int)(A.n - w0);
        const float4 *gp = reinterpret_cast<const float4 *>(A.obs_prev + w0 * kObs);
        float4 *sp = reinterpret_cast<float4 *>(sm.prev(warp));
        for (int i = lane; i < rows * kObs / 4; i += 32) sp[i] = __ldg(gp + i);
        for (int i = (rows * kObs / 4) * 4 + lane; i < rows * kObs; i += 32) sm.prev(warp)[i] = A.obs_prev[w0 * kObs + i];
        if (FORMAT != 1) {
            const float4 *gq = reinterpret_cast<const float4 *>(A.obs_next + w0 * kObs);
            float4 *sn = reinterpret_cast<float4 *>(sm.next(warp));
            for (int i = lane; i < rows * kObs / 4; i += 32) sn[i] = __ldg(gq + i);
            for (int i = (rows * kObs / 4) * 4 + lane; i < rows * kObs; i += 32) sm.next(warp)[i] = A.obs_next[w0 * kObs + i];
        }
    }
    __syncwarp();
    float row[WIDTH];                                                    // assembled in registers: the rows overwrite the staging buffer
    if (sel) {
        const float *sp = sm.prev(warp) + lane * kObs;
        const float *sn = use_term ? A.term_obs + e * kObs : sm.next(warp) + lane * kObs;
        if (FORMAT == 2) {
            const float dx1 = sp[0], v2 = sp[9];                                  // goal_status, hdqn.py:223-236
            const float status = dx1 < -0.5f * v2 ? 0.f : dx1 < 0.5f * v2 ? 1.f : 2.f;
            row[0] = g;
#pragma unroll
            for (int k = 0; k < kObs; ++k) row[1 + k] = sp[k];
            row[kObs + 1] = A.player == 2 ? act2 : act1;
            row[kObs + 2] = gn == status ? 1.f : 0.f;
            row[kObs + 3] = gn;
#pragma unroll
            for (int k = 0; k < kObs; ++k) row[kObs + 4 + k] = sn[k];
        } else {
#pragma unroll
            for (int k = 0; k < kObs; ++k) row[k] = sp[k];
            if (FORMAT == 0) {
                row[kObs] = A.player == 2 ? act2 : act1;
                row[kObs + 1] = A.player == 2 ? rw.y : rw.x;
#pragma unroll
                for (int k = 0; k < kObs; ++k) row[kObs + 2 + k] = sn[k];
            } else {
                row[kObs] = act1; row[kObs + 1] = act2;
                row[kObs + 2] = rw.x; row[kObs + 3] = rw.y;
            }
        }
    }
    __syncwarp();                                                        // every lane has read its s / s' row
    if (sel) {
        float *dst = sm.out(warp) + __popc(b & ((1u << lane) - 1u)) * WIDTH;
#pragma unroll
        for (int k = 0; k < WIDTH; ++k) dst[k] = row[k];
    }
    __syncwarp();
}

// Part 2: coalesced write of the warp's cnt rows to ring positions (base + rank0 + r) % capacity.  `total` = rows the
// whole call appends: rows a sequential writer would overwrite later in this same call are skipped.
template <int FORMAT>
__device__ __forceinline__ void store_rows(const RowArgs &A, RowSmem<FORMAT> &sm, int warp, int lane, int64_t w0, bool sel,
                                           unsigned b, unsigned long long base, uint64_t rank0, uint64_t total) {
    constexpr int WIDTH = RowFmt<FORMAT>::WIDTH;
    const int cnt = __popc(b);
    const unsigned long long cap = (unsigned long long)A.capacity;
    const unsigned long long slot0 = (base + rank0) % cap;
    if (sel && A.env_ids) {
        const int r = __popc(b & ((1u << lane) - 1u));
        if (!(rank0 + r + cap < total)) A.env_ids[(int64_t)((slot0 + (unsigned long long)r) % cap)] = (int32_t)(w0 + lane);
    }
    if (total <= cap && slot0 + (unsigned long long)cnt <= cap) {
        // common case: one contiguous span, 8-byte aligned (row width 88, 56 or 96 bytes) -> 64-bit stores
        const float2 *src = reinterpret_cast<const float2 *>(sm.out(warp));
        float2 *dst = reinterpret_cast<float2 *>(A.ring + slot0 * WIDTH);
        for (int i = lane; i < cnt * WIDTH / 2; i += 32) dst[i] = src[i];
    } else {
        for (int i = lane; i < cnt * WIDTH; i += 32) {
            const int r = i / WIDTH, c = i - r * WIDTH;
            if (rank0 + r + cap < total) continue;                         // overwritten later in this call
            const unsigned long long slot = (slot0 + (unsigned long long)r) % cap;
            A.ring[slot * WIDTH + c] = sm.out(warp)[i];
        }
    }
}

// mask_mode 0 (every env is stored; the h-DQN controller's rows): the count and the scan reduce to this
__global__ void advance_kernel(unsigned long long rows, unsigned long long *__restrict__ counter, unsigned long long *__restrict__ base) {
    cudaTriggerProgrammaticLaunchCompletion();
    *base = *counter;
    *counter += rows;
}

// The write pass (row formats and the warp's staging: see RowFmt / RowSmem above).
// (40 registers, 22.5 KB: six resident blocks per SM; forcing eight by a 32-register cap spills and is slower: 42 vs 37.6 us;
// ALL = mask_mode 0 is its own instantiation: as a run-time branch it cost the main path 8 registers and a resident block)
template <int FORMAT, bool ALL>
__global__ void __launch_bounds__(kBlock)
write_kernel(const RowArgs A, const uint32_t *block_offsets, const unsigned long long *base, const unsigned long long *counter) {
    __shared__ __align__(16) RowSmem<FORMAT> sm;
    __shared__ uint32_t s_cnt[kWarps];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t w0 = ((int64_t)blockIdx.x * kBlock + warp * 32);          // first env of this warp
    const int64_t e = w0 + lane;
    const bool full = w0 + 32 <= A.n;
    RowRegs R;
    if (full) request_rows<FORMAT>(A, lane, w0, R);
    const bool sel = e < A.n && selected(A.info, e, A.mask_mode);
    const unsigned b = __ballot_sync(0xFFFFFFFFu, sel);
    if (lane == 0) s_cnt[warp] = (uint32_t)__popc(b);
    __syncthreads();                                                         // the only block-wide step
    if (w0 >= A.n) return;
    uint32_t before = 0;                                                     // selected envs of the lower warps of this block
#pragma unroll
    for (int w = 0; w < kWarps; ++w) before += w < warp ? s_cnt[w] : 0u;
    assemble_rows<FORMAT>(A, sm, warp, lane, w0, sel, b, full, R);
    // Everything above read only what was complete before this call's count pass; the offsets, the pre-increment counter
    // and the new counter come from the scan, whose single CTA may still be running (programmatic dependent launch).
    // They are read with plain coherent loads through unqualified pointers: a `const __restrict__` load is invariant to
    // the compiler and was hoisted above the wait (stale offsets once the scan took several passes).
    cudaGridDependencySynchronize();
    const unsigned long long base_v = __ldcg(base), total_v = __ldcg(counter) - base_v;
    const uint64_t block_rank = ALL ? (uint64_t)blockIdx.x * kBlock : (uint64_t)__ldcg(block_offsets + blockIdx.x);
    store_rows<FORMAT>(A, sm, warp, lane, w0, sel, b, base_v, block_rank + before, total_v);
}

}  // namespace mgrec

extern "C" MG_API int64_t mg_record_scratch_words(int64_t n) {
    if (n < 0) return 0;
    return (n + 31) / 32 + 4;
}

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
    // scratch: uint32[mg_record_scratch_words(n)] — the first `grid` entries hold the block counts / offsets, the
    // 64-bit pre-increment counter sits behind entry m (8-byte aligned)
    uint32_t *block_counts = scratch;
    auto *base = reinterpret_cast<unsigned long long *>(scratch + ((m + 1) & ~(int64_t)1));
    auto *ctr = reinterpret_cast<unsigned long long *>(counter);
    const mgrec::RowArgs A{obs_prev, obs_next, term_obs_or_null, a1, a2_or_null, rew, done, info, goal_prev_or_null,
                           goal_next_or_null, n, mask_mode, player, ring, capacity, env_ids_or_null};
    const unsigned count_grid = (unsigned)((n + (int64_t)mgrec::kBlock * mgrec::kCountPerThread - 1) /
                                           ((int64_t)mgrec::kBlock * mgrec::kCountPerThread));
    const uint32_t *offs = block_counts;
    if (mask_mode == 0) {
        // every env is stored: a block's offset is its first env — no count, no scan; one thread advances the counter
        mgrec::advance_kernel<<<1, 1, 0, st>>>((unsigned long long)n, ctr, base);
        offs = nullptr;
    } else {
        mgrec::count_kernel<<<count_grid, mgrec::kBlock, 0, st>>>(info, n, mask_mode, block_counts);
        mgrec::scan_kernel<<<1, 1024, 0, st>>>(block_counts, (int64_t)grid, ctr, base);
    }
    // the write pass starts under the scan (programmatic dependent launch): see write_kernel
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(mgrec::kBlock); cfg.dynamicSmemBytes = 0; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    const unsigned long long *cbase = base, *cctr = ctr;
    cudaError_t le;
#define MG_REC_WRITE(F) (offs ? cudaLaunchKernelEx(&cfg, mgrec::write_kernel<F, false>, A, offs, cbase, cctr) \
                              : cudaLaunchKernelEx(&cfg, mgrec::write_kernel<F, true>, A, offs, cbase, cctr))
    if (format == 0) le = MG_REC_WRITE(0);
    else if (format == 1) le = MG_REC_WRITE(1);
    else le = MG_REC_WRITE(2);
#undef MG_REC_WRITE
    if (le) return cuda_fail(le, "mg_record_transitions launch");
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
