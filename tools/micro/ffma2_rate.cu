// Micro-benchmark: fp32 FMA issue rate on sm_100a — scalar FFMA vs packed FFMA2 (fma.rn.f32x2).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_rate ffma2_rate.cu && ./ffma2_rate
#include <cuda_runtime.h>
#include <stdio.h>

__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
    unsigned long long d;
    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(*reinterpret_cast<unsigned long long *>(&a)),
                 "l"(*reinterpret_cast<unsigned long long *>(&b)), "l"(*reinterpret_cast<unsigned long long *>(&c)));
    return *reinterpret_cast<float2 *>(&d);
}

template <bool PACKED>
__global__ void k(float *out, int iters, float x, float y) {
    float2 acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = make_float2(threadIdx.x + i, threadIdx.x - i);
    const float2 a = make_float2(x, x + 1.f), b = make_float2(y, y - 1.f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (PACKED) acc[i] = ffma2(acc[i], a, b);
            else {
                asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(acc[i].x) : "f"(a.x), "f"(b.x));
                asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(acc[i].y) : "f"(a.y), "f"(b.y));
            }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += acc[i].x + acc[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <bool PACKED>
void run(float *out, const char *name) {
    const int iters = 20000, threads = 512, blocks = 148 * 4;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<PACKED><<<blocks, threads>>>(out, 100, 1.0001f, 0.5f);
    cudaEventRecord(e0);
    k<PACKED><<<blocks, threads>>>(out, iters, 1.0001f, 0.5f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double flops = 2.0 * 16 * (double)iters * threads * blocks;
    printf("%s: %.3f ms, %.1f TFLOP/s fp32\n", name, ms, flops / ms / 1e9);
}

int main() {
    float *out;
    cudaMalloc(&out, 148 * 4 * 512 * sizeof(float));
    run<false>(out, "scalar FFMA ");
    run<true>(out, "packed FFMA2");
    return 0;
}
