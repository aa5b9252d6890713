// tc_common.cuh — pieces shared by the tensor-core policy kernels (mlp_tc_kernels.cu: 3xTF32, layer 1 on the CUDA cores;
// mlp_tc16_kernels.cu: 3xF16, both layers on the tensor cores): mbarrier helpers, the shared-memory matrix descriptor and
// the observation-row reader.
#pragma once
#include <cstdint>

#include "abi_common.h"
#include "merge_device.cuh"

namespace mgtc {

constexpr int TM = 128;                       // envs per tile = UMMA M
constexpr uint32_t kSpinLimit = 1u << 26;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// tcgen05 shared-memory matrix descriptor, no swizzle, K-major: core matrix = 8 rows x 16 bytes stored as
// 128 contiguous bytes; LBO = byte distance between the two core matrices of a K-step (128), SBO = byte
// distance between 8-row groups (256).  Verified numerically in profiles/exp_tcgen05.cu.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(128u >> 4) << 16) | ((uint64_t)(256u >> 4) << 32) |
           ((uint64_t)1 << 46);
}
constexpr uint64_t kDescHi = ((uint64_t)(256u >> 4) << 32) | ((uint64_t)1 << 46);   // SBO and version: the constant high word
__device__ __forceinline__ void mbar_init(unsigned long long *b, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(unsigned long long *b) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *b, uint32_t parity) {
    uint32_t done = 0;
    for (uint32_t it = 0; it < kSpinLimit && !done; ++it) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
            : "=r"(done)
            : "r"(smem_u32(b)), "r"(parity)
            : "memory");
    }
    if (!done) __trap();                      // never hang the GPU on a protocol bug
}
__device__ __forceinline__ uint32_t mbar_test(unsigned long long *b, uint32_t parity) {   // one non-blocking poll
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(done)
        : "r"(smem_u32(b)), "r"(parity)
        : "memory");
    return done;
}

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

}  // namespace mgtc
