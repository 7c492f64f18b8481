// Exact-fp32 SIMT GEMM with fused row gather, bias/activation epilogue and split-K.
// This is the fp32-parity workhorse behind every nn.Linear of the path (and the fallback for
// shapes the tcgen05 kernels do not take).  128x128x8 CTA tile, 8x8 register tile per thread,
// register-staged double buffering, conflict-free float4 shared-memory reads.
#include "gemm.cuh"

namespace xnrs {


constexpr int BM = 128, BN = 128, BK = 8, NT = 256;

// Operand whose K index is contiguous in memory (stored [dim, K]): thread -> (row = t/2, 4 k's).
// Returns the 4 values for k0+kq..k0+kq+3 of stored row `src` (or zeros).
__device__ __forceinline__ float4 fetch_kcontig(const float *base, long long ld, long long src, bool row_ok,
                                                long long k, long long kend, bool vec_ok) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (!row_ok) return v;
    const float *p = base + src * ld + k;
    if (vec_ok && k + 3 < kend) {
        v = *reinterpret_cast<const float4 *>(p);
    } else {
        if (k < kend) v.x = p[0];
        if (k + 1 < kend) v.y = p[1];
        if (k + 2 < kend) v.z = p[2];
        if (k + 3 < kend) v.w = p[3];
    }
    return v;
}

// Operand whose M/N index is contiguous (stored [K, dim]): thread -> (k = t/32, 4 consecutive dim's).
__device__ __forceinline__ float4 fetch_dcontig(const float *base, long long ld, const int *rows, long long k,
                                                long long kend, long long d, long long dim, bool vec_ok) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (k >= kend || d >= dim) return v;
    long long src = rows ? (long long)rows[k] : k;
    const float *p = base + src * ld + d;
    if (vec_ok && d + 3 < dim) {
        v = *reinterpret_cast<const float4 *>(p);
    } else {
        v.x = p[0];
        if (d + 1 < dim) v.y = p[1];
        if (d + 2 < dim) v.z = p[2];
        if (d + 3 < dim) v.w = p[3];
    }
    return v;
}

__global__ void __launch_bounds__(NT) gemm_simt_kernel(GemmArgs p) {
    __shared__ __align__(16) float As[2][BK][BM];
    __shared__ __align__(16) float Bs[2][BK][BN];

    const int t = threadIdx.x;
    // linearised tile index, N-tiles fastest: the CTAs sharing one A row-block run back to back (L2 reuse)
    const long long ntn = (p.N + BN - 1) / BN;
    const long long m0 = ((long long)blockIdx.x / ntn) * BM, n0 = ((long long)blockIdx.x % ntn) * BN;
    const long long kbeg = (long long)blockIdx.z * p.k_per_split;
    const long long kend = min(p.K, kbeg + p.k_per_split);
    if (kbeg >= kend && blockIdx.z > 0) return;

    const bool a_vec = (p.lda % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.A) & 15) == 0);
    const bool b_vec = (p.ldb % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.B) & 15) == 0);

    // per-thread load coordinates
    const int kc_row = t >> 1, kc_kq = (t & 1) * 4;      // k-contiguous operand
    const int dc_k = t >> 5, dc_d = (t & 31) * 4;        // dim-contiguous operand
    long long a_src = 0, b_src = 0;
    bool a_ok = false, b_ok = false;
    if (!p.transA) {
        long long gm = m0 + kc_row;
        a_ok = gm < p.M;
        if (a_ok) a_src = p.a_rows ? (long long)p.a_rows[gm] : gm;
    }
    if (p.transB) {
        long long gn = n0 + kc_row;
        b_ok = gn < p.N;
        if (b_ok) b_src = p.b_rows ? (long long)p.b_rows[gn] : gn;
    }

    auto fetchA = [&](long long k0) -> float4 {
        if (!p.transA) return fetch_kcontig(p.A, p.lda, a_src, a_ok, k0 + kc_kq, kend, a_vec);
        return fetch_dcontig(p.A, p.lda, p.a_rows, k0 + dc_k, kend, m0 + dc_d, p.M, a_vec);
    };
    auto fetchB = [&](long long k0) -> float4 {
        if (p.transB) return fetch_kcontig(p.B, p.ldb, b_src, b_ok, k0 + kc_kq, kend, b_vec);
        return fetch_dcontig(p.B, p.ldb, p.b_rows, k0 + dc_k, kend, n0 + dc_d, p.N, b_vec);
    };
    auto stashA = [&](int buf, float4 v) {
        if (!p.transA) {
            As[buf][kc_kq + 0][kc_row] = v.x; As[buf][kc_kq + 1][kc_row] = v.y;
            As[buf][kc_kq + 2][kc_row] = v.z; As[buf][kc_kq + 3][kc_row] = v.w;
        } else {
            *reinterpret_cast<float4 *>(&As[buf][dc_k][dc_d]) = v;
        }
    };
    auto stashB = [&](int buf, float4 v) {
        if (p.transB) {
            Bs[buf][kc_kq + 0][kc_row] = v.x; Bs[buf][kc_kq + 1][kc_row] = v.y;
            Bs[buf][kc_kq + 2][kc_row] = v.z; Bs[buf][kc_kq + 3][kc_row] = v.w;
        } else {
            *reinterpret_cast<float4 *>(&Bs[buf][dc_k][dc_d]) = v;
        }
    };

    const int tx = t & 15, ty = t >> 4;     // 16 x 16 thread grid; rows {ty*4+i, 64+ty*4+i}, cols likewise
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    const long long ktiles = (kend - kbeg + BK - 1) / BK;
    if (ktiles > 0) {
        stashA(0, fetchA(kbeg));
        stashB(0, fetchB(kbeg));
    }
    __syncthreads();
    int cur = 0;
    for (long long kt = 0; kt < ktiles; ++kt) {
        float4 ra, rb;
        const bool more = kt + 1 < ktiles;
        if (more) {
            ra = fetchA(kbeg + (kt + 1) * BK);
            rb = fetchB(kbeg + (kt + 1) * BK);
        }
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            float4 a0 = *reinterpret_cast<const float4 *>(&As[cur][k][ty * 4]);
            float4 a1 = *reinterpret_cast<const float4 *>(&As[cur][k][64 + ty * 4]);
            float4 b0 = *reinterpret_cast<const float4 *>(&Bs[cur][k][tx * 4]);
            float4 b1 = *reinterpret_cast<const float4 *>(&Bs[cur][k][64 + tx * 4]);
            float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (more) {
            stashA(cur ^ 1, ra);
            stashB(cur ^ 1, rb);
        }
        __syncthreads();
        cur ^= 1;
    }

    // epilogue
    const bool first_split = blockIdx.z == 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        long long gm = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
        if (gm >= p.M) continue;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            long long gn = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
            if (gn >= p.N) continue;
            float v = acc[i][j];
            float *c = p.C + gm * p.ldc + gn;
            if (p.split_k > 1) {
                if (first_split && p.bias) v += p.bias[gn];
                atomicAdd(c, v);
            } else {
                if (p.bias) v += p.bias[gn];
                if (p.act == XNRS_ACT_RELU) v = fmaxf(v, 0.f);
                else if (p.act == XNRS_ACT_TANH) v = tanhf(v);
                else if (p.act == XNRS_ACT_RELU_MASK) v = (p.aux[gm * p.ldc + gn] > 0.f) ? v : 0.f;
                if (p.accumulate) v += *c;
                *c = v;
            }
        }
    }
}

int gemm_simt(const GemmArgs &a_in, cudaStream_t st) {
    GemmArgs a = a_in;
    long long tiles = cdiv(a.M, BM) * cdiv(a.N, BN);
    if (a.split_k <= 0) {
        // auto: split K when the output is too small to fill the machine (weight-gradient GEMMs)
        long long want = 2LL * num_sms();
        long long s = tiles >= want ? 1 : want / tiles;
        long long maxs = cdiv(a.K, 256);
        if (s > maxs) s = maxs;
        if (s < 1) s = 1;
        if (a.act != XNRS_ACT_NONE) s = 1;
        a.split_k = (int)s;
    }
    if (a.split_k > 1 && a.act != XNRS_ACT_NONE) return fail(XNRS_ERR_ARG, "%s: split_k with activation", "xnrs_gemm");
    a.k_per_split = cdiv(cdiv(a.K, a.split_k), BK) * BK;
    if (a.k_per_split <= 0) a.k_per_split = BK;
    if (a.split_k > 1 && !a.accumulate) {
        cudaError_t e = cudaMemset2DAsync(a.C, a.ldc * sizeof(float), 0, a.N * sizeof(float), a.M, st);
        if (e != cudaSuccess) return fail(XNRS_ERR_CUDA, "%s: memset2d failed", "xnrs_gemm");
    }
    dim3 grid((unsigned)tiles, 1, (unsigned)a.split_k);
    gemm_simt_kernel<<<grid, NT, 0, st>>>(a);
    XNRS_LAUNCHED();
    return XNRS_OK;
}


}  // namespace xnrs

using namespace xnrs;

namespace xnrs {
thread_local const char *g_last_gemm_kernel = "";
std::atomic<long long> g_simt_fallbacks{0};
}  // namespace xnrs

extern "C" int xnrs_gemm(int transA, int transB, long long M, long long N, long long K, const float *A,
                         long long lda, const int *a_rows, const float *B, long long ldb, const int *b_rows,
                         float *C, long long ldc, const float *bias, int act, const float *aux, int accumulate,
                         int split_k, int precision, xnrs_stream_t st) {
    XNRS_REQUIRE(M >= 0 && N >= 0 && K >= 0, "negative dimension");
    if (M == 0 || N == 0) return XNRS_OK;
    XNRS_REQUIRE(A && B && C, "null operand");
    XNRS_REQUIRE(lda >= (transA ? M : K) && ldb >= (transB ? K : N) && ldc >= N, "leading dimension too small");
    XNRS_REQUIRE(act >= 0 && act <= 3, "bad activation");
    XNRS_REQUIRE(act != XNRS_ACT_RELU_MASK || aux, "RELU_MASK needs aux");
    if (precision == XNRS_PREC_BF16)
        return fail(XNRS_ERR_UNSUPPORTED, "%s: XNRS_PREC_BF16 means bf16 operands: call xnrs_gemm_bf16 (fp32 operands run in FP32 / TF32X3 / TF32)", "xnrs_gemm");
    XNRS_REQUIRE(cdiv(M, BM) * cdiv(N, BN) < 2147483647LL, "too many tiles for one launch");
    GemmArgs a{M, N, K, A, lda, a_rows, transA, B, ldb, b_rows, transB, C, ldc, bias, act, aux, accumulate,
               split_k, 0};
    if (precision != XNRS_PREC_FP32) {
        int status = XNRS_OK;
        if (gemm_tensorcore(a, precision, STREAM(st), &status)) return status;
        g_simt_fallbacks.fetch_add(1, std::memory_order_relaxed);
    }
    g_last_gemm_kernel = "gemm_simt_kernel";
    return gemm_simt(a, STREAM(st));
}

extern "C" const char *xnrs_last_gemm_kernel(void) { return g_last_gemm_kernel; }
extern "C" long long xnrs_gemm_simt_fallbacks(void) { return g_simt_fallbacks.load(); }
