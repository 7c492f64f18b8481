// GEMM argument block shared by the SIMT (exact fp32) and tcgen05 (tensor-core) implementations.
#pragma once
#include "common.cuh"

namespace xnrs {

struct GemmArgs {
    long long M, N, K;
    const float *A; long long lda; const int *a_rows; int transA;
    const float *B; long long ldb; const int *b_rows; int transB;
    float *C; long long ldc;
    const float *bias; int act; const float *aux; int accumulate;
    int split_k; long long k_per_split;
};

// which kernel the last xnrs_gemm call dispatched to (bench.py's roofline.kernel), and how many calls made in a
// tensor-core precision mode were taken by the exact-fp32 SIMT kernel instead (unsupported shape / alignment)
extern thread_local const char *g_last_gemm_kernel;
extern std::atomic<long long> g_simt_fallbacks;

int gemm_simt(const GemmArgs &a, cudaStream_t st);
// gemm_tc.cu: returns 1 if the problem was taken by the tcgen05 path (*status holds the result), 0 if not
int gemm_tensorcore(const GemmArgs &a, int precision, cudaStream_t st, int *status);

}  // namespace xnrs
