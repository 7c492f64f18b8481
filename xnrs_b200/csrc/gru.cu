// Row U-lstur: single-layer GRU over the front-aligned click history, final hidden state taken at each
// user's true length (lstur.py:139-153; torch.nn.GRU semantics, gate order r,z,n):
//     r = sig(gi_r + gh_r)   z = sig(gi_z + gh_z)   n = tanh(gi_n + r * gh_n)   h' = (1-z) n + z h
// with gi = x W_ih^T + b_ih (one big GEMM over all steps, done by xnrs_gemm) and gh = h W_hh^T + b_hh.
// The sequential part runs as ONE persistent kernel: a CTA owns UB users for all L steps, keeps h in
// shared memory, and streams W_hh (L2-resident, coalesced through the transposed copy) every step.
// Lengths stay on the device (the reference syncs them to the host, lstur.py:141).
#include "common.cuh"

namespace xnrs {

constexpr int UB = 4;   // users per CTA

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

__global__ void gru_fwd_kernel(const float *__restrict__ gi, const float *__restrict__ w_hh_t,
                               const float *__restrict__ b_hh, const float *__restrict__ h0,
                               const int *__restrict__ lengths, long long B, int L, int Hd, float *__restrict__ hs,
                               float *__restrict__ gates, float *__restrict__ h_out) {
    extern __shared__ float sm[];
    float *h = sm;                    // [UB][Hd]
    float *gh = sm + UB * Hd;         // [UB][3Hd]
    const int tid = threadIdx.x, H3 = 3 * Hd;
    const long long u0 = (long long)blockIdx.x * UB;
    int len[UB];
#pragma unroll
    for (int u = 0; u < UB; ++u) len[u] = (u0 + u < B) ? lengths[u0 + u] : 0;
    for (int i = tid; i < UB * Hd; i += blockDim.x) {
        int u = i / Hd, k = i - u * Hd;
        float v = (h0 && u0 + u < B) ? h0[(u0 + u) * Hd + k] : 0.f;
        h[i] = v;
    }
    __syncthreads();
    for (int t = 0; t < L; ++t) {
        if (tid < H3) {
            float acc[UB];
            const float b = b_hh[tid];
#pragma unroll
            for (int u = 0; u < UB; ++u) acc[u] = b;
            for (int k = 0; k < Hd; ++k) {
                const float w = w_hh_t[(long long)k * H3 + tid];
#pragma unroll
                for (int u = 0; u < UB; ++u) acc[u] = fmaf(h[u * Hd + k], w, acc[u]);
            }
#pragma unroll
            for (int u = 0; u < UB; ++u) gh[u * H3 + tid] = acc[u];
        }
        __syncthreads();
        for (int i = tid; i < UB * Hd; i += blockDim.x) {
            const int u = i / Hd, k = i - u * Hd;
            const long long b = u0 + u;
            if (b >= B) continue;
            const float *g = gi + (b * L + t) * H3;
            const float hp = h[i];
            const float ghn = gh[u * H3 + 2 * Hd + k];
            const float r = sigmoidf_(g[k] + gh[u * H3 + k]);
            const float z = sigmoidf_(g[Hd + k] + gh[u * H3 + Hd + k]);
            const float n = tanhf(g[2 * Hd + k] + r * ghn);
            const float hn = (t < len[u]) ? (1.f - z) * n + z * hp : hp;
            float *gt = gates + (b * L + t) * 4 * Hd;
            gt[k] = r; gt[Hd + k] = z; gt[2 * Hd + k] = n; gt[3 * Hd + k] = ghn;
            hs[(b * L + t) * Hd + k] = hp;                 // state before step t (all the backward needs)
            h[i] = hn;
        }
        __syncthreads();
    }
    for (int i = tid; i < UB * Hd; i += blockDim.x) {
        int u = i / Hd, k = i - u * Hd;
        if (u0 + u < B) h_out[(u0 + u) * Hd + k] = h[i];
    }
}

__global__ void gru_bwd_kernel(const float *__restrict__ d_h_out, const float *__restrict__ w_hh,
                               const int *__restrict__ lengths, const float *__restrict__ hs,
                               const float *__restrict__ gates, long long B, int L, int Hd, float *__restrict__ d_gi,
                               float *__restrict__ d_gh, float *__restrict__ d_h0) {
    extern __shared__ float sm[];
    float *dh = sm;                   // [UB][Hd]   running d loss / d h_t
    float *carry = sm + UB * Hd;      // [UB][Hd]   dh * z (direct path)
    float *dg = sm + 2 * UB * Hd;     // [UB][3Hd]  d loss / d gh_t
    const int tid = threadIdx.x, H3 = 3 * Hd;
    const long long u0 = (long long)blockIdx.x * UB;
    int len[UB];
#pragma unroll
    for (int u = 0; u < UB; ++u) len[u] = (u0 + u < B) ? lengths[u0 + u] : 0;
    for (int i = tid; i < UB * Hd; i += blockDim.x) {
        int u = i / Hd, k = i - u * Hd;
        dh[i] = (u0 + u < B) ? d_h_out[(u0 + u) * Hd + k] : 0.f;
    }
    __syncthreads();
    for (int t = L - 1; t >= 0; --t) {
        for (int i = tid; i < UB * Hd; i += blockDim.x) {
            const int u = i / Hd, k = i - u * Hd;
            const long long b = u0 + u;
            float gr = 0.f, gz = 0.f, gn = 0.f, ghn_ = 0.f, c = dh[i];
            if (b < B && t < len[u]) {
                const float *gt = gates + (b * L + t) * 4 * Hd;
                const float r = gt[k], z = gt[Hd + k], n = gt[2 * Hd + k], ghn = gt[3 * Hd + k];
                const float hp = hs[(b * L + t) * Hd + k];
                const float d = dh[i];
                const float dn = d * (1.f - z) * (1.f - n * n);     // wrt n pre-activation
                gz = d * (hp - n) * z * (1.f - z);
                gr = dn * ghn * r * (1.f - r);
                gn = dn;
                ghn_ = dn * r;
                c = d * z;
            }
            carry[i] = c;
            dg[u * H3 + k] = gr; dg[u * H3 + Hd + k] = gz; dg[u * H3 + 2 * Hd + k] = ghn_;
            if (b < B) {
                float *o = d_gi + (b * L + t) * H3;
                o[k] = gr; o[Hd + k] = gz; o[2 * Hd + k] = gn;
                float *o2 = d_gh + (b * L + t) * H3;
                o2[k] = gr; o2[Hd + k] = gz; o2[2 * Hd + k] = ghn_;
            }
        }
        __syncthreads();
        if (tid < Hd) {
            float acc[UB];
#pragma unroll
            for (int u = 0; u < UB; ++u) acc[u] = carry[u * Hd + tid];
            for (int j = 0; j < H3; ++j) {
                const float w = w_hh[(long long)j * Hd + tid];
#pragma unroll
                for (int u = 0; u < UB; ++u) acc[u] = fmaf(dg[u * H3 + j], w, acc[u]);
            }
#pragma unroll
            for (int u = 0; u < UB; ++u) dh[u * Hd + tid] = acc[u];
        }
        __syncthreads();
    }
    for (int i = tid; i < UB * Hd; i += blockDim.x) {
        int u = i / Hd, k = i - u * Hd;
        if (u0 + u < B && d_h0) d_h0[(u0 + u) * Hd + k] = dh[i];
    }
}

__global__ void lengths_kernel(const float *__restrict__ mask, long long B, int L, int *__restrict__ lengths) {
    for (long long b = blockIdx.x * (long long)blockDim.x + threadIdx.x; b < B; b += (long long)gridDim.x * blockDim.x) {
        float s = 0.f;
        for (int l = 0; l < L; ++l) s += mask[b * L + l];
        lengths[b] = (int)(s + 0.5f);
    }
}

}  // namespace xnrs

using namespace xnrs;

extern "C" int xnrs_gru_fwd(const float *gi, const float *w_hh_t, const float *b_hh, const float *h0,
                            const int *lengths, long long B, int L, int Hd, float *hs, float *gates, float *h_out,
                            xnrs_stream_t st) {
    XNRS_REQUIRE(B >= 0 && L > 0 && Hd > 0 && 3 * Hd <= 1024, "bad sizes (3*Hd <= 1024)");
    if (B == 0) return XNRS_OK;
    XNRS_REQUIRE(gi && w_hh_t && b_hh && lengths && hs && gates && h_out, "null pointer");
    int threads = (int)(cdiv(3 * Hd, 32) * 32);
    size_t smem = (size_t)(UB * Hd + UB * 3 * Hd) * sizeof(float);
    gru_fwd_kernel<<<(unsigned)cdiv(B, UB), threads, smem, STREAM(st)>>>(gi, w_hh_t, b_hh, h0, lengths, B, L, Hd, hs, gates,
                                                                    h_out);
    XNRS_LAUNCHED();
    return XNRS_OK;
}

extern "C" int xnrs_gru_bwd(const float *d_h_out, const float *w_hh, const int *lengths, const float *hs,
                            const float *gates, long long B, int L, int Hd, float *d_gi, float *d_gh, float *d_h0,
                            xnrs_stream_t st) {
    XNRS_REQUIRE(B >= 0 && L > 0 && Hd > 0 && Hd <= 1024, "bad sizes");
    if (B == 0) return XNRS_OK;
    XNRS_REQUIRE(d_h_out && w_hh && lengths && hs && gates && d_gi && d_gh, "null pointer");
    int threads = (int)(cdiv(Hd, 32) * 32);
    if (threads < 128) threads = 128;
    size_t smem = (size_t)(2 * UB * Hd + UB * 3 * Hd) * sizeof(float);
    gru_bwd_kernel<<<(unsigned)cdiv(B, UB), threads, smem, STREAM(st)>>>(d_h_out, w_hh, lengths, hs, gates, B, L, Hd, d_gi,
                                                                    d_gh, d_h0);
    XNRS_LAUNCHED();
    return XNRS_OK;
}

extern "C" int xnrs_lengths_from_mask(const float *mask, long long B, int L, int *lengths, xnrs_stream_t st) {
    XNRS_REQUIRE(B >= 0 && L > 0, "bad sizes");
    if (B == 0) return XNRS_OK;
    XNRS_REQUIRE(mask && lengths, "null pointer");
    lengths_kernel<<<(unsigned)cdiv(B, 256), 256, 0, STREAM(st)>>>(mask, B, L, lengths);
    XNRS_LAUNCHED();
    return XNRS_OK;
}
