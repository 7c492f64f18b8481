// Shared host/device helpers for the xnrs_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>

#include "../../include/xnrs_b200.h"

namespace xnrs {

extern thread_local char g_err[512];
extern std::atomic<long long> g_launches;

inline int fail(int code, const char *fmt, const char *a = "", long long b = 0, long long c = 0) {
    snprintf(g_err, sizeof(g_err), fmt, a, b, c);
    return code;
}

#define XNRS_REQUIRE(cond, msg)                                                          \
    do {                                                                                 \
        if (!(cond)) return xnrs::fail(XNRS_ERR_ARG, "%s: requirement failed: " msg, __func__); \
    } while (0)

// call after every kernel launch: counts it and surfaces launch-configuration errors (no sync)
#define XNRS_LAUNCHED()                                                                        \
    do {                                                                                       \
        xnrs::g_launches.fetch_add(1, std::memory_order_relaxed);                              \
        cudaError_t e_ = cudaGetLastError();                                                   \
        if (e_ != cudaSuccess) {                                                               \
            snprintf(xnrs::g_err, sizeof(xnrs::g_err), "%s: CUDA error: %s", __func__, cudaGetErrorString(e_)); \
            return XNRS_ERR_CUDA;                                                              \
        }                                                                                      \
    } while (0)

inline cudaStream_t STREAM(xnrs_stream_t st) { return reinterpret_cast<cudaStream_t>(st); }

inline int num_sms() {
    static int n = 0;
    if (!n) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

inline long long cdiv(long long a, long long b) { return (a + b - 1) / b; }

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
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// block-wide sum; `red` is >= 32 floats of shared memory; all threads get the result
__device__ __forceinline__ float block_sum(float v, float *red) {
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) red[w] = v;
    __syncthreads();
    float t = (lane < nw) ? red[lane] : 0.f;
    return warp_sum(t);
}

// streaming 128-bit load that does not pollute L1 (table rows are read once per CTA)
__device__ __forceinline__ float4 ldg_stream(const float4 *p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p));
    return r;
}

}  // namespace xnrs
