// Micro-benchmark: issue rate of the legacy warp-level tensor path (mma.sync.m16n8k8 TF32) on sm_100a.
// Decides whether the per-head attention products (30x30x48) are worth moving from FFMA2 to mma.sync.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_sync_rate mma_sync_rate.cu && ./mma_sync_rate
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>

template <int CHAINS>
__global__ void k(float *out, int iters) {
    float c[CHAINS][4];
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.f;
    uint32_t a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, b0 = a0 * 3, b1 = a0 * 5;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < CHAINS; ++i)
            asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                         : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
    float *out;
    cudaMalloc(&out, 148 * 16 * 1024 * sizeof(float));
    for (int warps = 4; warps <= 16; warps *= 2) {
        const int iters = 20000;
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        k<8><<<148 * 2, warps * 32>>>(out, 100);
        cudaEventRecord(e0);
        k<8><<<148 * 2, warps * 32>>>(out, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double flops = 2.0 * 16 * 8 * 8 * 8.0 * iters * warps * 148 * 2;
        printf("warps/CTA %d (2 CTAs/SM): %.3f ms, %.1f TFLOP/s tf32 (mma.sync m16n8k8)\n", warps, ms, flops / ms / 1e9);
    }
    return 0;
}
