// Shared host/device helpers for libvcmesh.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <string>

#include "vcmesh.h"

namespace vcm {

// ---- error reporting (thread-local message, C ABI returns codes) ----------
std::string &last_error_slot();
int fail(int code, const char *fmt, ...);

#define VCM_CUDA(expr)                                                                      \
    do {                                                                                    \
        cudaError_t _e = (expr);                                                            \
        if (_e != cudaSuccess)                                                              \
            return ::vcm::fail((int)_e, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                               __FILE__, __LINE__);                                         \
    } while (0)

#define VCM_CHECK(cond, ...)                                   \
    do {                                                       \
        if (!(cond)) return ::vcm::fail(VCM_EINVAL, __VA_ARGS__); \
    } while (0)

// diagnostics only: how many kernels this library has launched (bench.py's gpu_launches)
void count_launch();
unsigned long long launches();

#define VCM_LAUNCH_CHECK()                                                                 \
    do {                                                                                   \
        ::vcm::count_launch();                                                             \
        cudaError_t _e = cudaPeekAtLastError();                                            \
        if (_e != cudaSuccess)                                                             \
            return ::vcm::fail((int)_e, "kernel launch failed: %s (%s:%d)",                \
                               cudaGetErrorString(_e), __FILE__, __LINE__);                \
    } while (0)

// Rejects host pointers: there is no CPU fallback.
bool is_device_ptr(const void *p);
#define VCM_DEVPTR(p) VCM_CHECK((p) != nullptr && ::vcm::is_device_ptr(p), #p " must be a device pointer")
#define VCM_DEVPTR_OPT(p) VCM_CHECK((p) == nullptr || ::vcm::is_device_ptr(p), #p " must be a device pointer or NULL")

int sm_count();

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// ---- device tables (owned by vcm_tables) -----------------------------------
struct Tables {
    int32_t p_in = 0, p_out = 0, n_max = 0, max_last = 0, max_rev = 0, device = 0;
    int32_t n_gt16 = 0;           // vertices with last > 16: slots [0, n_gt16) of the processing order (degree buckets)
    int64_t nnz = 0;
    int32_t *idx = nullptr;       // [p_out, n_max]
    int32_t *last = nullptr;      // [p_out]
    int32_t *perm = nullptr;      // [p_out]
    int32_t *rev_ptr = nullptr;   // [p_in + 1]
    int32_t *rev_slot = nullptr;  // [nnz]
    int32_t *idx_sorted = nullptr;   // [p_out, n_max] rows of idx in `perm` order
    int32_t *meta_sorted = nullptr;  // [p_out, 2] {vertex, last} in `perm` order (one 8-byte load per work item)
    int32_t *blob = nullptr;      // single allocation backing all of the above
};

inline const Tables *as_tables(const vcm_tables *t) { return reinterpret_cast<const Tables *>(t); }

// ---- device helpers -----------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ float4 ldg4(const float *p) { return __ldg(reinterpret_cast<const float4 *>(p)); }
__device__ __forceinline__ void st4(float *p, float4 v) { *reinterpret_cast<float4 *>(p) = v; }
// streaming store: written once, consumed by a later kernel, keep it out of L1
__device__ __forceinline__ void st4_cs(float *p, float4 v) { __stcs(reinterpret_cast<float4 *>(p), v); }

// Packed fp32x2 FMA (FFMA2, new on sm_100): two IEEE fp32 FMAs per lane per issued instruction; results are
// bit-identical to two scalar FMAs. The gather-side kernels are issue-bound, so halving their FMA count pays.
__device__ __forceinline__ float2 fma2(float a, float2 b, float2 c) { return __ffma2_rn(make_float2(a, a), b, c); }
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }

template <int VEC>
struct Vec;
template <>
struct Vec<4> {
    float v[4];
    __device__ __forceinline__ void load(const float *p) {
        float4 t = ldg4(p);
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    }
    __device__ __forceinline__ void store(float *p) const { st4(p, make_float4(v[0], v[1], v[2], v[3])); }
};
template <>
struct Vec<1> {
    float v[1];
    __device__ __forceinline__ void load(const float *p) { v[0] = __ldg(p); }
    __device__ __forceinline__ void store(float *p) const { *p = v[0]; }
};
#endif

}  // namespace vcm
