// Shared host/device helpers for the xnrs_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

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

// ---- programmatic dependent launch ------------------------------------------------------------------------------------
// A training step is ~45 kernels of 2-500 us replayed from one CUDA graph; between two dependent kernels the GPU idles for
// the ~1.5-2 us it takes to drain one grid and launch the next.  A kernel launched through launch_pdl may start while its
// predecessor in the stream is still finishing (as soon as all the predecessor's CTAs have exited or called
// pdl_launch_dependents): it does its set-up (barriers, TMEM, descriptor prefetch) and then calls pdl_wait(), which returns
// once every earlier kernel has completed and its writes are visible.  RULE: a kernel launched this way touches no global
// memory before pdl_wait().  Inside a normally launched kernel both calls are no-ops.  XNRS_PDL=0 launches everything the
// classic way.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

inline bool pdl_enabled() {
    static int on = -1;
    if (on < 0) { const char *e = getenv("XNRS_PDL"); on = e ? atoi(e) != 0 : 1; }
    return on != 0;
}

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

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
