// Shared device/host helpers for the mse_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "../../include/mse_b200.h"

namespace mse {

// ---- error plumbing (thread-local message behind mse_last_error) -------------------------
void set_error(const char* fmt, ...);

#define MSE_CUDA_TRY(expr)                                                                   \
    do {                                                                                     \
        cudaError_t _e = (expr);                                                             \
        if (_e != cudaSuccess) {                                                             \
            mse::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return _e == cudaErrorMemoryAllocation ? MSE_ERR_NOMEM : MSE_ERR_CUDA;           \
        }                                                                                    \
    } while (0)

#define MSE_REQUIRE(cond, ...)                                                               \
    do {                                                                                     \
        if (!(cond)) {                                                                       \
            mse::set_error(__VA_ARGS__);                                                     \
            return MSE_ERR_INVALID;                                                          \
        }                                                                                    \
    } while (0)

// ---- order-preserving float <-> uint32 key ------------------------------------------------
// key(a) < key(b)  <=>  a < b for all non-NaN floats, with key(-0.0) < key(+0.0); scores are
// canonicalised (v + 0.0f) before keying so that a zero score is always +0.0.
__host__ __device__ __forceinline__ uint32_t float_to_key(float f) {
#ifdef __CUDA_ARCH__
    uint32_t u = __float_as_uint(f);
#else
    union { float f; uint32_t u; } c; c.f = f; uint32_t u = c.u;
#endif
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ __forceinline__ float key_to_float(uint32_t k) {
    uint32_t u = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
#ifdef __CUDA_ARCH__
    return __uint_as_float(u);
#else
    union { float f; uint32_t u; } c; c.u = u; return c.f;
#endif
}
// 64-bit selection key: larger == better rank.  Score descending, then doc ascending
// (bm25_indexer.py:484 — stable sort over rows that arrive in ascending doc id).
__host__ __device__ __forceinline__ uint64_t make_key64(uint32_t score_key, uint32_t doc) {
    return (uint64_t(score_key) << 32) | uint64_t(0xffffffffu - doc);
}
__host__ __device__ __forceinline__ uint32_t key64_doc(uint64_t k) { return 0xffffffffu - uint32_t(k); }
__host__ __device__ __forceinline__ uint32_t key64_score_key(uint64_t k) { return uint32_t(k >> 32); }

constexpr uint32_t kUntouchedBits = 0x80000000u;   // -0.0f: x + (-0.0) == x, and (+0.0) + (-0.0) == +0.0

#ifdef __CUDACC__
__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ int warp_id() { return threadIdx.x >> 5; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ int warp_incl_scan(int v) {
    const int l = lane_id();
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, v, o);
        if (l >= o) v += t;
    }
    return v;
}
// streaming 128-bit load that does not allocate in L1 (each byte is used once)
__device__ __forceinline__ uint4 ldg_stream(const uint4* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ int ldg_stream_i32(const int* p) {
    int r;
    asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(r) : "l"(p));
    return r;
}
__device__ __forceinline__ int2 ldg_stream_i2(const int2* p) {
    int2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.s32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ uint32_t ld_relaxed_u32(const uint32_t* p) {
    uint32_t r;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(r) : "l"(p));
    return r;
}
#endif

}  // namespace mse
