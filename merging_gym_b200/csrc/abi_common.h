// abi_common.h — error plumbing shared by the translation units of libmerging_b200.so.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

#include "../../include/merging_b200.h"

namespace mg_abi {
inline thread_local char g_err[512] = "";

inline int fail(int code, const char *msg) {
    snprintf(g_err, sizeof g_err, "%s", msg);
    return code;
}
inline int cuda_fail(cudaError_t e, const char *where) {
    snprintf(g_err, sizeof g_err, "%s: %s (%s)", where, cudaGetErrorString(e), cudaGetErrorName(e));
    return (int)e;
}
inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
}  // namespace mg_abi
