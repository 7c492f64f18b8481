// Rows A / P: additive and personalised attention pooling (layers.py:47-69, 88-101), forward and
// backward, plus masked-mean pooling and the collapsed title mask.
//
// The hidden layer hid = tanh(fc1 x) is produced by the GEMM; these kernels fuse everything after it:
// logit = <hid, w> (+b), exp, multiplicative mask, normalisation by (sum + 1e-8) and the weighted sum
// over the (optionally table-gathered) rows of x.  One CTA walks titles grid-stride; rows whose
// weight is exactly 0 (padding) are never read.  HBM-bound: per title it reads L*A (hid) + L*F (x).
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdlib.h>

#include "common.cuh"

namespace xnrs {

constexpr int POOL_THREADS = 256;
// warp-per-group kernels (one warp owns a title; used when there are thousands of groups)
constexpr int WPB = 8;            // warps per CTA
constexpr int FQ = 8;             // float4 accumulators per lane: F <= 32 * 4 * FQ = 1024
constexpr int AQ = 8;             // hidden units per lane: A <= 32 * AQ = 256

template <bool kPers>
__global__ void __launch_bounds__(POOL_THREADS)
pool_fwd_kernel(const float *__restrict__ x, const int *__restrict__ x_rows, const float *__restrict__ mask,
                const float *__restrict__ hid, const float *__restrict__ w2, const float *__restrict__ b2,
                const float *__restrict__ qh, int rows_per_query, const int *__restrict__ seg, long long R, int L, int F,
                int A, float *__restrict__ attn, float *__restrict__ pooled) {
    extern __shared__ float sm[];
    float *e = sm;                               // [L]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    const int F4 = F >> 2;
    const float4 *x4 = reinterpret_cast<const float4 *>(x);
    const int Lmax = L;
    for (long long r = blockIdx.x; r < R; r += gridDim.x) {
        // ragged groups: rows [seg[r], seg[r+1]) (padding tokens were never materialised); else fixed L rows per group
        const long long base = seg ? (long long)seg[r] : r * (long long)Lmax;
        const int L = seg ? seg[r + 1] - seg[r] : Lmax;
        const float *wv = kPers ? qh + (r / rows_per_query) * A : w2;
        const float bias = kPers ? 0.f : b2[0];
        for (int l = warp; l < L; l += nwarps) {
            const float mval = mask ? mask[base + l] : 1.f;
            float acc = 0.f;
            if (mval != 0.f) {
                const float *hrow = hid + (base + l) * A;
                for (int j = lane; j < A; j += 32) acc = fmaf(hrow[j], wv[j], acc);
                acc = warp_sum(acc);
            }
            if (lane == 0) e[l] = (mval != 0.f) ? expf(acc + bias) * mval : 0.f;
        }
        __syncthreads();
        float tot = 0.f;
        for (int l = 0; l < L; ++l) tot += e[l];
        const float denom = tot + 1e-8f;
        for (int l = tid; l < L; l += blockDim.x) attn[base + l] = e[l] / denom;
        for (int c = tid; c < F4; c += blockDim.x) {
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int l = 0; l < L; ++l) {
                const float a = e[l] / denom;
                if (a == 0.f) continue;
                const long long row = x_rows ? (long long)x_rows[base + l] : base + l;
                const float4 v = ldg_stream(x4 + row * F4 + c);
                acc.x = fmaf(a, v.x, acc.x); acc.y = fmaf(a, v.y, acc.y);
                acc.z = fmaf(a, v.z, acc.z); acc.w = fmaf(a, v.w, acc.w);
            }
            reinterpret_cast<float4 *>(pooled)[r * F4 + c] = acc;
        }
        __syncthreads();
    }
}

// backward.  With a = E/(T+eps), E = exp(logit)*m:  dlogit_l = a_l * (da_l - sum_j a_j da_j).
template <bool kPers>
__global__ void __launch_bounds__(POOL_THREADS)
pool_bwd_kernel(const float *__restrict__ x, const int *__restrict__ x_rows, const float *__restrict__ hid,
                const float *__restrict__ w2, const float *__restrict__ qh, int rows_per_query,
                const float *__restrict__ attn, const float *__restrict__ d_pooled, const float *__restrict__ d_attn,
                const int *__restrict__ seg, long long R, int L, int F, int A, long long n_rows, float *__restrict__ d_hid,
                float *__restrict__ d_w2, float *__restrict__ d_b2, float *__restrict__ d_qh, float *__restrict__ d_x,
                float *__restrict__ d_b1) {
    extern __shared__ float sm[];
    float *da = sm;            // [L]  da_l, then dlogit_l            (L here = the maximum group length)
    float *al = sm + L;        // [L]  a_l
    float *dw = sm + 2 * L;    // [A]  per-CTA accumulator for d_w2 (additive only)
    float *db1 = sm + 2 * L + A;   // [A]  per-CTA accumulator for the fc1 bias gradient = column sums of d_hid (optional)
    __shared__ float db_acc;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    const int F4 = F >> 2;
    const float4 *x4 = reinterpret_cast<const float4 *>(x);
    if (!kPers) {
        for (int j = tid; j < A; j += blockDim.x) dw[j] = 0.f;
        if (tid == 0) db_acc = 0.f;
    }
    if (d_b1)
        for (int j = tid; j < A; j += blockDim.x) db1[j] = 0.f;
    __syncthreads();
    const int Lmax = L;
    for (long long r = blockIdx.x; r < R; r += gridDim.x) {
        const long long base = seg ? (long long)seg[r] : r * (long long)Lmax;
        const int L = seg ? seg[r + 1] - seg[r] : Lmax;
        const float4 *dp4 = reinterpret_cast<const float4 *>(d_pooled) + r * F4;
        for (int l = warp; l < L; l += nwarps) {
            const float a = attn[base + l];
            float acc = 0.f;
            if (a != 0.f) {
                const long long row = x_rows ? (long long)x_rows[base + l] : base + l;
                for (int c = lane; c < F4; c += 32) {
                    const float4 v = ldg_stream(x4 + row * F4 + c);
                    const float4 g = dp4[c];
                    acc = fmaf(v.x, g.x, acc); acc = fmaf(v.y, g.y, acc);
                    acc = fmaf(v.z, g.z, acc); acc = fmaf(v.w, g.w, acc);
                }
                acc = warp_sum(acc);
                if (d_attn) acc += d_attn[base + l];
            }
            if (lane == 0) { da[l] = acc; al[l] = a; }
        }
        __syncthreads();
        float dot = 0.f;
        for (int l = 0; l < L; ++l) dot = fmaf(al[l], da[l], dot);
        __syncthreads();
        for (int l = tid; l < L; l += blockDim.x) da[l] = al[l] * (da[l] - dot);     // dlogit
        __syncthreads();
        const float *wv = kPers ? qh + (r / rows_per_query) * A : w2;
        for (int j = tid; j < A; j += blockDim.x) {
            const float w = wv[j];
            float gw = 0.f, gb = 0.f;
            // four rows per iteration with the loads hoisted (rows with dlogit 0 — padding — still write 0, branch-free)
            int l = 0;
            for (; l + 4 <= L; l += 4) {
                const long long idx = (base + l) * A + j;
                const float h0 = hid[idx], h1 = hid[idx + A], h2 = hid[idx + 2 * A], h3 = hid[idx + 3 * A];
                const float d0 = da[l], d1 = da[l + 1], d2 = da[l + 2], d3 = da[l + 3];
                gw = fmaf(d0, h0, gw); gw = fmaf(d1, h1, gw); gw = fmaf(d2, h2, gw); gw = fmaf(d3, h3, gw);
                const float g0 = d0 * w * (1.f - h0 * h0), g1 = d1 * w * (1.f - h1 * h1);
                const float g2 = d2 * w * (1.f - h2 * h2), g3 = d3 * w * (1.f - h3 * h3);
                d_hid[idx] = g0;
                d_hid[idx + A] = g1;
                d_hid[idx + 2 * A] = g2;
                d_hid[idx + 3 * A] = g3;
                gb += (g0 + g1) + (g2 + g3);
            }
            for (; l < L; ++l) {
                const long long idx = (base + l) * A + j;
                const float dl = da[l], h = hid[idx];
                gw = fmaf(dl, h, gw);
                const float g0 = dl * w * (1.f - h * h);
                d_hid[idx] = g0;
                gb += g0;
            }
            if (d_b1) db1[j] += gb;
            if (kPers) atomicAdd(d_qh + (r / rows_per_query) * A + j, gw);
            else dw[j] += gw;
        }
        if (!kPers && tid == 0) {
            float s = 0.f;
            for (int l = 0; l < L; ++l) s += da[l];
            db_acc += s;
        }
        if (d_x) {
            for (int i = tid; i < L * F4; i += blockDim.x) {
                const int l = i / F4, c = i - l * F4;
                const float a = al[l];
                const float4 g = dp4[c];
                reinterpret_cast<float4 *>(d_x)[(base + l) * F4 + c] = make_float4(a * g.x, a * g.y, a * g.z, a * g.w);
            }
        }
        __syncthreads();
    }
    if (!kPers) {
        for (int j = tid; j < A; j += blockDim.x) atomicAdd(d_w2 + j, dw[j]);
        if (tid == 0) atomicAdd(d_b2, db_acc);
    }
    if (d_b1)
        for (int j = tid; j < A; j += blockDim.x) atomicAdd(d_b1 + j, db1[j]);
    // ragged groups listed in a padded row buffer (TitlePlan): the rows past the last group belong to no title; their
    // gradient is exactly 0 and is written here so that the weight-gradient GEMM / column sum can run over all n_rows
    if (seg && n_rows > 0) {
        const long long t0 = (long long)seg[R] * A, t1 = n_rows * A;
        for (long long i = t0 + blockIdx.x * (long long)blockDim.x + tid; i < t1; i += (long long)gridDim.x * blockDim.x)
            d_hid[i] = 0.f;
    }
}

// bf16-storage twin of pool_bwd_kernel<false> (XNRS_PREC_BF16: token rows x, hid and d_hid are bf16 in HBM, every sum is
// fp32): x rows come from the bf16 token table through x_rows; no d_x / d_attn (the table is frozen).
__device__ __forceinline__ float bf16_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }

__global__ void __launch_bounds__(POOL_THREADS)
pool_bwd_bf16_kernel(const __nv_bfloat16 *__restrict__ x, const int *__restrict__ x_rows, const __nv_bfloat16 *__restrict__ hid,
                     const float *__restrict__ w2, const float *__restrict__ attn, const float *__restrict__ d_pooled,
                     const int *__restrict__ seg, long long R, int L, int F, int A, long long n_rows,
                     __nv_bfloat16 *__restrict__ d_hid, float *__restrict__ d_w2, float *__restrict__ d_b2, float *__restrict__ d_b1) {
    extern __shared__ float sm[];
    float *da = sm, *al = sm + L, *dw = sm + 2 * L, *db1 = sm + 2 * L + A;
    __shared__ float db_acc;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    const int F8 = F >> 3, F4 = F >> 2;
    const uint4 *x8 = reinterpret_cast<const uint4 *>(x);
    for (int j = tid; j < A; j += blockDim.x) { dw[j] = 0.f; db1[j] = 0.f; }
    if (tid == 0) db_acc = 0.f;
    __syncthreads();
    const int Lmax = L;
    for (long long r = blockIdx.x; r < R; r += gridDim.x) {
        const long long base = seg ? (long long)seg[r] : r * (long long)Lmax;
        const int L = seg ? seg[r + 1] - seg[r] : Lmax;
        const float4 *dp4 = reinterpret_cast<const float4 *>(d_pooled) + r * F4;
        for (int l = warp; l < L; l += nwarps) {
            const float a = attn[base + l];
            float acc = 0.f;
            if (a != 0.f) {
                const long long row = x_rows ? (long long)x_rows[base + l] : base + l;
                for (int c = lane; c < F8; c += 32) {
                    const uint4 v = __ldg(x8 + row * F8 + c);
                    const float4 g0 = dp4[2 * c], g1 = dp4[2 * c + 1];
                    acc = fmaf(bf16_lo(v.x), g0.x, acc); acc = fmaf(bf16_hi(v.x), g0.y, acc);
                    acc = fmaf(bf16_lo(v.y), g0.z, acc); acc = fmaf(bf16_hi(v.y), g0.w, acc);
                    acc = fmaf(bf16_lo(v.z), g1.x, acc); acc = fmaf(bf16_hi(v.z), g1.y, acc);
                    acc = fmaf(bf16_lo(v.w), g1.z, acc); acc = fmaf(bf16_hi(v.w), g1.w, acc);
                }
                acc = warp_sum(acc);
            }
            if (lane == 0) { da[l] = acc; al[l] = a; }
        }
        __syncthreads();
        float dot = 0.f;
        for (int l = 0; l < L; ++l) dot = fmaf(al[l], da[l], dot);
        __syncthreads();
        for (int l = tid; l < L; l += blockDim.x) da[l] = al[l] * (da[l] - dot);     // dlogit
        __syncthreads();
        for (int j = tid; j < A; j += blockDim.x) {
            const float w = w2[j];
            float gw = 0.f, gb = 0.f;
            for (int l = 0; l < L; ++l) {
                const long long idx = (base + l) * A + j;
                const float dl = da[l], h = __bfloat162float(hid[idx]);
                gw = fmaf(dl, h, gw);
                const __nv_bfloat16 g16 = __float2bfloat16_rn(dl * w * (1.f - h * h));
                d_hid[idx] = g16;
                gb += __bfloat162float(g16);          // the bias gradient sums the STORED values, as a column sum of d_hid would
            }
            db1[j] += gb;
            dw[j] += gw;
        }
        if (tid == 0) {
            float sacc = 0.f;
            for (int l = 0; l < L; ++l) sacc += da[l];
            db_acc += sacc;
        }
        __syncthreads();
    }
    for (int j = tid; j < A; j += blockDim.x) {
        atomicAdd(d_w2 + j, dw[j]);
        if (d_b1) atomicAdd(d_b1 + j, db1[j]);
    }
    if (tid == 0) atomicAdd(d_b2, db_acc);
    if (seg && n_rows > 0) {
        const long long t0 = (long long)seg[R] * A, t1 = n_rows * A;
        for (long long i = t0 + blockIdx.x * (long long)blockDim.x + tid; i < t1; i += (long long)gridDim.x * blockDim.x)
            d_hid[i] = __float2bfloat16_rn(0.f);
    }
}

// ---- warp-per-title backward of the additive pooler (training at title level: thousands of ragged groups, frozen table) ----
// One warp owns a title from the row dots da_l = <x_l, d_pooled> to its d_hid rows: no block barriers, two rows of table loads
// in flight per lane, per-lane column accumulators for d_w2 / d_b1 carried across all the titles the warp walks (one atomic
// flush per warp at the end).  x, hid, d_hid are fp32 or bf16 (BF); sums are fp32.  Valid for F <= 768 (fp32) / 1536 (bf16),
// A <= 256, no d_x / d_attn.  The CTA-per-title kernels above stay for everything else.
template <bool BF>
__global__ void __launch_bounds__(WPB * 32)
pool_bwd_warp_kernel(const void *__restrict__ x_, const int *__restrict__ x_rows, const void *__restrict__ hid_,
                     const float *__restrict__ w2, const float *__restrict__ attn, const float *__restrict__ d_pooled,
                     const int *__restrict__ seg, long long R, int Lmax, int F, int A, long long n_rows, void *__restrict__ d_hid_,
                     float *__restrict__ d_w2, float *__restrict__ d_b2, float *__restrict__ d_b1) {
    extern __shared__ float sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float *dl = sm + warp * Lmax;                   // this warp's dlogit_l
    constexpr int EPC = BF ? 8 : 4;                 // elements per 16-byte chunk
    constexpr int NCH = BF ? 3 : 6;                 // chunks per lane: F <= 768
    const int nchunk = F / EPC;                     // 16-byte chunks per x row
    const int F4 = F >> 2;
    const long long wid = (long long)blockIdx.x * WPB + warp, nw = (long long)gridDim.x * WPB;
    float wreg[AQ], gw[AQ], gb[AQ];
#pragma unroll
    for (int i = 0; i < AQ; ++i) {
        wreg[i] = (lane + 32 * i < A) ? w2[lane + 32 * i] : 0.f;
        gw[i] = gb[i] = 0.f;
    }
    float db2 = 0.f;
    for (long long r = wid; r < R; r += nw) {
        const long long base = seg ? (long long)seg[r] : r * (long long)Lmax;
        const int L = seg ? seg[r + 1] - seg[r] : Lmax;
        if (L == 0) continue;
        // d_pooled row of this title, in the lane's chunk layout: chunk c = lane + 32 i covers elements [c*EPC, +EPC)
        float dp[NCH][EPC];
#pragma unroll
        for (int i = 0; i < NCH; ++i) {
            const int c = lane + 32 * i;
#pragma unroll
            for (int q = 0; q < EPC; q += 4) {
                const float4 g = (c < nchunk) ? reinterpret_cast<const float4 *>(d_pooled)[r * F4 + (c * EPC + q) / 4]
                                              : make_float4(0.f, 0.f, 0.f, 0.f);
                dp[i][q] = g.x; dp[i][q + 1] = g.y; dp[i][q + 2] = g.z; dp[i][q + 3] = g.w;
            }
        }
        // phase 1: da_l = <x_l, d_pooled>, two rows per iteration (all loads first); s = sum_l a_l da_l
        float s = 0.f;
        for (int l0 = 0; l0 < L; l0 += 2) {
            uint4 v[2][NCH];
#pragma unroll
            for (int i2 = 0; i2 < 2; ++i2) {
                const int l = min(l0 + i2, L - 1);
                const long long row = x_rows ? (long long)x_rows[base + l] : base + l;
                const uint4 *src = reinterpret_cast<const uint4 *>(reinterpret_cast<const char *>(x_) + row * (long long)F * (BF ? 2 : 4));
#pragma unroll
                for (int i = 0; i < NCH; ++i)
                    v[i2][i] = (lane + 32 * i < nchunk) ? __ldg(src + lane + 32 * i) : make_uint4(0u, 0u, 0u, 0u);
            }
            float acc[2] = {0.f, 0.f};
#pragma unroll
            for (int i2 = 0; i2 < 2; ++i2) {
#pragma unroll
                for (int i = 0; i < NCH; ++i) {
                    const uint32_t w[4] = {v[i2][i].x, v[i2][i].y, v[i2][i].z, v[i2][i].w};
                    if (BF) {
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            acc[i2] = fmaf(__uint_as_float(w[q] << 16), dp[i][2 * q], acc[i2]);
                            acc[i2] = fmaf(__uint_as_float(w[q] & 0xffff0000u), dp[i][2 * q + 1], acc[i2]);
                        }
                    } else {
#pragma unroll
                        for (int q = 0; q < 4; ++q) acc[i2] = fmaf(__uint_as_float(w[q]), dp[i][q], acc[i2]);
                    }
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                acc[0] += __shfl_xor_sync(0xffffffffu, acc[0], o);
                acc[1] += __shfl_xor_sync(0xffffffffu, acc[1], o);
            }
#pragma unroll
            for (int i2 = 0; i2 < 2; ++i2) {
                if (l0 + i2 < L) {
                    const float a = attn[base + l0 + i2];
                    s = fmaf(a, acc[i2], s);
                    if (lane == 0) dl[l0 + i2] = acc[i2];        // da_l for now
                }
            }
        }
        __syncwarp();
        // phase 2: dlogit_l = a_l (da_l - s); d_hid[l, j] = dlogit_l w2_j (1 - h^2); column sums for d_w2 / d_b1
        for (int l = 0; l < L; ++l) {
            const float d = attn[base + l] * (dl[l] - s);
            db2 += (lane == 0) ? d : 0.f;
#pragma unroll
            for (int i = 0; i < AQ; ++i) {
                const int j = lane + 32 * i;
                if (j < A) {
                    const long long idx = (base + l) * A + j;
                    float h, g;
                    if (BF) {
                        const __nv_bfloat16 *hp = reinterpret_cast<const __nv_bfloat16 *>(hid_);
                        h = __bfloat162float(hp[idx]);
                        const __nv_bfloat16 g16 = __float2bfloat16_rn(d * wreg[i] * (1.f - h * h));
                        reinterpret_cast<__nv_bfloat16 *>(d_hid_)[idx] = g16;
                        g = __bfloat162float(g16);
                    } else {
                        h = reinterpret_cast<const float *>(hid_)[idx];
                        g = d * wreg[i] * (1.f - h * h);
                        reinterpret_cast<float *>(d_hid_)[idx] = g;
                    }
                    gw[i] = fmaf(d, h, gw[i]);
                    gb[i] += g;
                }
            }
        }
        __syncwarp();
    }
#pragma unroll
    for (int i = 0; i < AQ; ++i) {
        const int j = lane + 32 * i;
        if (j < A) {
            atomicAdd(d_w2 + j, gw[i]);
            if (d_b1) atomicAdd(d_b1 + j, gb[i]);
        }
    }
    if (lane == 0 && db2 != 0.f) atomicAdd(d_b2, db2);
    if (seg && n_rows > 0) {        // rows past the last group (TitlePlan padding): d_hid = 0
        const long long t0 = (long long)seg[R] * A, t1 = n_rows * A;
        for (long long i = t0 + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < t1; i += (long long)gridDim.x * blockDim.x) {
            if (BF) reinterpret_cast<__nv_bfloat16 *>(d_hid_)[i] = __float2bfloat16_rn(0.f);
            else reinterpret_cast<float *>(d_hid_)[i] = 0.f;
        }
    }
}

// ---- CTA-per-title backward of the additive pooler, second form (training at title level, frozen token table: no d_x, no
// d_attn; A = 256, F <= 768; x / hid / d_hid fp32 or bf16, every sum fp32).
// pool_bwd_kernel is latency bound: 0.21 ms whether the rows are fp32 or bf16, 3.6 TB/s of its ~0.77 GB (fp32) — per title
// a warp has ONE 3 KB row in flight for the row dots, a thread four 4-byte loads for the d_hid pass, with five block
// barriers in between.  This form keeps more bytes in flight and fewer barriers:
//   * the NEXT title's token rows and hid rows are prefetched into L2 while the current one is processed;
//   * row dots: four rows per warp at a time (all 16-byte loads issued before the first FMA), the title's d_pooled row
//     staged in shared memory one title ahead;
//   * d_hid pass: 16-byte loads / stores, thread = (column group, row lane), 32 rows in flight per CTA; the column sums
//     for d_w2 / d_b1 stay in registers across all the titles a CTA walks (one shared + one global atomic flush at the end);
//   * two block barriers per title (da/al -> dot -> dlogit in a third array).
// Measured at the bench shapes (8.9 k titles, 153.6 k rows): 0.212 -> 0.161 ms (fp32), 0.210 -> 0.129 ms (bf16).  A warp-per-title
// variant of the same pipeline (no barriers, 16 titles in flight per SM) was slower again (0.193 / 0.238 ms) and is not kept.
// SPLIT_OUT (fp32 inputs only): d_hid is written as two bf16 planes, d_hid ~ hi + lo (hi = bf16(g), lo = bf16(g - hi)), the form
// the pre-split 3xBF16 weight-gradient GEMM consumes — same bytes as the fp32 rows, no separate split pass.
template <bool BF, bool SPLIT_OUT>
__global__ void __launch_bounds__(POOL_THREADS, 2)
pool_bwd2_kernel(const void *__restrict__ x_, const int *__restrict__ x_rows, const void *__restrict__ hid_,
                 const float *__restrict__ w2, const float *__restrict__ attn, const float *__restrict__ d_pooled,
                 const int *__restrict__ seg, long long R, int Lmax, int F, long long n_rows, void *__restrict__ d_hid_,
                 void *__restrict__ d_hid_lo_, float *__restrict__ d_w2, float *__restrict__ d_b2, float *__restrict__ d_b1) {
    constexpr int A = 256, ELT = BF ? 2 : 4, EPC = 16 / ELT, NCH = BF ? 3 : 6, TOK = 4;
    constexpr int CG = A / EPC, RL = POOL_THREADS / CG, U = 32 / RL;       // 32 rows of hid in flight per pass
    extern __shared__ __align__(16) float sm[];
    float *dps = sm;                                // [2][F]  d_pooled row of the current / next title
    float *da = sm + 2 * F, *al = da + Lmax, *dlg = al + Lmax, *dw = dlg + Lmax, *db1 = dw + A;
    float *at_s = db1 + A;                          // [3][Lmax] pooling weights of the current / next / next-next title
    int *xr_s = reinterpret_cast<int *>(at_s + 3 * Lmax);       // [3][Lmax] their token rows
    __shared__ float db_acc;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nchunk = F / EPC, F4 = F >> 2;
    const int cg = tid % CG, rl = tid / CG;
    float w[EPC], gw[EPC], gb[EPC];
#pragma unroll
    for (int e = 0; e < EPC; ++e) {
        w[e] = w2[cg * EPC + e];
        gw[e] = gb[e] = 0.f;
    }
    for (int j = tid; j < 2 * A; j += POOL_THREADS) dw[j] = 0.f;         // dw and db1 are contiguous
    if (tid == 0) db_acc = 0.f;
    const char *xb = reinterpret_cast<const char *>(x_), *hb = reinterpret_cast<const char *>(hid_);
    char *dhb = reinterpret_cast<char *>(d_hid_);
    const long long row_bytes = (long long)F * ELT;
    const int xlines = (int)((row_bytes + 127) >> 7);
    const long long g = gridDim.x;
    // Software pipeline over the titles r, r + g, ... of this CTA.  Every global load whose RESULT is an address (group
    // offsets, token rows) or a branch condition (pooling weights) is issued one or two titles before it is needed — a warp
    // issues in order, so a dependent load in the title's own path costs a full round trip each (three of them were stacked
    // in front of the first row load).  Title t's offsets travel in registers (b0/L0 current, b1/L1, b2/L2), its token rows
    // and weights in a three-slot ring in shared memory, its d_pooled row in a two-slot ring.
    auto group = [&](long long r, long long &bb, int &LL) {
        if (r < R) {
            bb = seg ? (long long)seg[r] : r * (long long)Lmax;
            LL = seg ? seg[r + 1] - seg[r] : Lmax;
        } else {
            bb = 0; LL = 0;
        }
    };
    long long b0, b1, b2, b3;
    int L0, L1, L2, L3;
    group(blockIdx.x, b0, L0);
    group(blockIdx.x + g, b1, L1);
    group(blockIdx.x + 2 * g, b2, L2);
    if (tid < L0) { at_s[tid] = attn[b0 + tid]; xr_s[tid] = x_rows ? x_rows[b0 + tid] : (int)(b0 + tid); }
    if (tid < L1) { at_s[Lmax + tid] = attn[b1 + tid]; xr_s[Lmax + tid] = x_rows ? x_rows[b1 + tid] : (int)(b1 + tid); }
    if (blockIdx.x < R)
        for (int c = tid; c < F4; c += POOL_THREADS)
            reinterpret_cast<float4 *>(dps)[c] = reinterpret_cast<const float4 *>(d_pooled)[(long long)blockIdx.x * F4 + c];
    __syncthreads();
    int pb = 0, slot = 0;
    for (long long r = blockIdx.x; r < R; r += g, pb ^= 1, slot = slot == 2 ? 0 : slot + 1) {
        const long long base = b0;
        const int L = L0;
        const long long rn = r + g;
        const int s1 = slot == 2 ? 0 : slot + 1, s2 = s1 == 2 ? 0 : s1 + 1;
        // (1) two titles ahead: weights / token rows into registers (stored to the ring before the first barrier below);
        //     three titles ahead: group offsets
        float at2 = 0.f;
        int xr2 = 0;
        if (tid < L2) { at2 = attn[b2 + tid]; xr2 = x_rows ? x_rows[b2 + tid] : (int)(b2 + tid); }
        group(r + 3 * g, b3, L3);
        // (2) the next title: its token rows and hid rows into L2 (row numbers from the ring: no dependent load)
        for (int i = tid; i < L1 * xlines; i += POOL_THREADS) {
            const int l = i / xlines, c = i - l * xlines;
            asm volatile("prefetch.global.L2 [%0];" ::"l"(xb + (long long)xr_s[s1 * Lmax + l] * row_bytes + c * 128));
        }
        {
            const long long hbytes = (long long)L1 * A * ELT;
            for (long long i = tid * 128LL; i < hbytes; i += POOL_THREADS * 128LL)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(hb + b1 * A * ELT + i));
        }
        const float *dpc = dps + pb * F;
        const float *at_c = at_s + slot * Lmax;
        const int *xr_c = xr_s + slot * Lmax;
        // (3) row dots da_l = <x_l, d_pooled>: warp `warp` takes rows warp, warp + 8, ..., TOK of them per round, every
        //     16-byte load of the round issued before the first FMA
        for (int k0 = 0; warp + 8 * k0 < L; k0 += TOK) {
            uint4 v[TOK][NCH];
            float a[TOK];
#pragma unroll
            for (int t = 0; t < TOK; ++t) {
                const int l = warp + 8 * (k0 + t);
                a[t] = l < L ? at_c[l] : 0.f;
                if (a[t] != 0.f) {                  // rows with weight exactly 0 (padding) are never read
                    const uint4 *src = reinterpret_cast<const uint4 *>(xb + (long long)xr_c[l] * row_bytes);
#pragma unroll
                    for (int i = 0; i < NCH; ++i)
                        v[t][i] = (lane + 32 * i < nchunk) ? __ldg(src + lane + 32 * i) : make_uint4(0u, 0u, 0u, 0u);
                } else {
#pragma unroll
                    for (int i = 0; i < NCH; ++i) v[t][i] = make_uint4(0u, 0u, 0u, 0u);
                }
            }
            float acc[TOK];
#pragma unroll
            for (int t = 0; t < TOK; ++t) acc[t] = 0.f;
#pragma unroll
            for (int i = 0; i < NCH; ++i) {
                const int c = min(lane + 32 * i, nchunk - 1);       // chunks past F hold zeros in v
                float gd[EPC];
#pragma unroll
                for (int q = 0; q < EPC; q += 4) {
                    const float4 t4 = reinterpret_cast<const float4 *>(dpc)[(c * EPC + q) >> 2];
                    gd[q] = t4.x; gd[q + 1] = t4.y; gd[q + 2] = t4.z; gd[q + 3] = t4.w;
                }
#pragma unroll
                for (int t = 0; t < TOK; ++t) {
                    const uint32_t ww[4] = {v[t][i].x, v[t][i].y, v[t][i].z, v[t][i].w};
                    if (BF) {
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            acc[t] = fmaf(bf16_lo(ww[q]), gd[2 * q], acc[t]);
                            acc[t] = fmaf(bf16_hi(ww[q]), gd[2 * q + 1], acc[t]);
                        }
                    } else {
#pragma unroll
                        for (int q = 0; q < 4; ++q) acc[t] = fmaf(__uint_as_float(ww[q]), gd[q], acc[t]);
                    }
                }
            }
#pragma unroll
            for (int t = 0; t < TOK; ++t) {
                const float sacc = warp_sum(acc[t]);
                const int l = warp + 8 * (k0 + t);
                if (lane == 0 && l < L) { da[l] = sacc; al[l] = a[t]; }
            }
        }
        // (4) ring updates for the titles ahead (their slots were last read before the previous title's barriers)
        if (tid < L2) { at_s[s2 * Lmax + tid] = at2; xr_s[s2 * Lmax + tid] = xr2; }
        if (rn < R)
            for (int c = tid; c < F4; c += POOL_THREADS)
                reinterpret_cast<float4 *>(dps + (pb ^ 1) * F)[c] = reinterpret_cast<const float4 *>(d_pooled)[rn * F4 + c];
        b0 = b1; L0 = L1; b1 = b2; L1 = L2; b2 = b3; L2 = L3;
        __syncthreads();
        float part = 0.f;                           // every warp forms the same sum_l a_l da_l (same order: identical bits)
        for (int l = lane; l < L; l += 32) part = fmaf(al[l], da[l], part);
        const float dot = warp_sum(part);
        float mine = 0.f;
        for (int l = tid; l < L; l += POOL_THREADS) {
            const float d = al[l] * (da[l] - dot);  // dlogit_l
            dlg[l] = d;
            mine += d;
        }
        if (warp * 32 < L) {                        // d_b2 += sum_l dlogit_l
            mine = warp_sum(mine);
            if (lane == 0) atomicAdd(&db_acc, mine);
        }
        __syncthreads();
        // d_hid[l, j] = dlogit_l w2_j (1 - h^2); column sums for d_w2 / d_b1
        for (int l0 = rl; l0 < L; l0 += RL * U) {
            uint4 h[U];
            float d[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int l = l0 + u * RL;
                d[u] = l < L ? dlg[l] : 0.f;
                h[u] = l < L ? *reinterpret_cast<const uint4 *>(hb + ((base + l) * A + cg * EPC) * ELT) : make_uint4(0u, 0u, 0u, 0u);
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int l = l0 + u * RL;
                if (l >= L) continue;
                const uint32_t hw[4] = {h[u].x, h[u].y, h[u].z, h[u].w};
                uint32_t out[4];
                if (BF) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const float h0 = bf16_lo(hw[q]), h1 = bf16_hi(hw[q]);
                        const __nv_bfloat16 g0 = __float2bfloat16_rn(d[u] * w[2 * q] * (1.f - h0 * h0));
                        const __nv_bfloat16 g1 = __float2bfloat16_rn(d[u] * w[2 * q + 1] * (1.f - h1 * h1));
                        gw[2 * q] = fmaf(d[u], h0, gw[2 * q]);
                        gw[2 * q + 1] = fmaf(d[u], h1, gw[2 * q + 1]);
                        gb[2 * q] += __bfloat162float(g0);          // the bias gradient sums the STORED values
                        gb[2 * q + 1] += __bfloat162float(g1);
                        out[q] = (uint32_t)__bfloat16_as_ushort(g0) | ((uint32_t)__bfloat16_as_ushort(g1) << 16);
                    }
                } else {
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const float hv = __uint_as_float(hw[q]);
                        const float g = d[u] * w[q] * (1.f - hv * hv);
                        gw[q] = fmaf(d[u], hv, gw[q]);
                        gb[q] += g;
                        out[q] = __float_as_uint(g);
                    }
                    if (SPLIT_OUT) {
                        uint32_t hi2[2], lo2[2];
#pragma unroll
                        for (int q = 0; q < 2; ++q) {
                            const float g0 = __uint_as_float(out[2 * q]), g1 = __uint_as_float(out[2 * q + 1]);
                            const __nv_bfloat16 h0 = __float2bfloat16_rn(g0), h1 = __float2bfloat16_rn(g1);
                            const __nv_bfloat16 l0 = __float2bfloat16_rn(g0 - __bfloat162float(h0));
                            const __nv_bfloat16 l1 = __float2bfloat16_rn(g1 - __bfloat162float(h1));
                            hi2[q] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
                            lo2[q] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
                        }
                        const long long off = ((base + l) * A + cg * EPC) * 2;
                        *reinterpret_cast<uint2 *>(dhb + off) = make_uint2(hi2[0], hi2[1]);
                        *reinterpret_cast<uint2 *>(reinterpret_cast<char *>(d_hid_lo_) + off) = make_uint2(lo2[0], lo2[1]);
                        continue;
                    }
                }
                *reinterpret_cast<uint4 *>(dhb + ((base + l) * A + cg * EPC) * ELT) = make_uint4(out[0], out[1], out[2], out[3]);
            }
        }
    }
#pragma unroll
    for (int e = 0; e < EPC; ++e) {
        atomicAdd(&dw[cg * EPC + e], gw[e]);
        atomicAdd(&db1[cg * EPC + e], gb[e]);
    }
    __syncthreads();
    for (int j = tid; j < A; j += POOL_THREADS) {
        atomicAdd(d_w2 + j, dw[j]);
        if (d_b1) atomicAdd(d_b1 + j, db1[j]);
    }
    if (tid == 0) atomicAdd(d_b2, db_acc);
    if (seg && n_rows > 0) {        // rows past the last group (TitlePlan padding): d_hid = 0
        const long long t0 = (long long)seg[R] * A, t1 = n_rows * A;
        for (long long i = t0 + blockIdx.x * (long long)blockDim.x + tid; i < t1; i += (long long)gridDim.x * blockDim.x) {
            if (SPLIT_OUT) {
                reinterpret_cast<__nv_bfloat16 *>(d_hid_)[i] = __float2bfloat16_rn(0.f);
                reinterpret_cast<__nv_bfloat16 *>(d_hid_lo_)[i] = __float2bfloat16_rn(0.f);
            } else if (BF) {
                reinterpret_cast<__nv_bfloat16 *>(d_hid_)[i] = __float2bfloat16_rn(0.f);
            } else {
                reinterpret_cast<float *>(d_hid_)[i] = 0.f;
            }
        }
    }
}

// x ~ hi + lo with hi = bf16(x), lo = bf16(x - hi): 16 mantissa bits in two bf16 planes (relative error <= 2^-17)
__global__ void split_bf16_kernel(long long n, const float *__restrict__ src, __nv_bfloat16 *__restrict__ hi,
                                  __nv_bfloat16 *__restrict__ lo) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float v = src[i];
        const __nv_bfloat16 h = __float2bfloat16_rn(v);
        hi[i] = h;
        lo[i] = __float2bfloat16_rn(v - __bfloat162float(h));
    }
}

// the fp16 form: hi = fp16(x), lo = fp16(x - hi): 22 mantissa bits (relative error <= 2^-24 for |x| in fp16's normal range,
// absolute error <= 3e-8 below it); values beyond +-65504 saturate.  For operands of moderate range (embeddings, weights);
// gradients, whose magnitude is arbitrary, keep the bf16 form.
__global__ void split_f16_kernel(long long n, const float *__restrict__ src, __half *__restrict__ hi, __half *__restrict__ lo) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float v = fminf(fmaxf(src[i], -65504.f), 65504.f);
        const __half h = __float2half_rn(v);
        hi[i] = h;
        lo[i] = __float2half_rn(v - __half2float(h));
    }
}

static bool pool_bwd2_ok(int L, int F, int A, const void *x, const void *hid, const void *d_hid, int elt) {
    static int on = -1;             // XNRS_POOL_BWD2=0: the first CTA-per-title kernels everywhere
    if (on < 0) { const char *e = getenv("XNRS_POOL_BWD2"); on = e ? atoi(e) : 1; }
    return on && A == 256 && F <= 768 && F % (16 / elt) == 0 && L <= POOL_THREADS && !((uintptr_t)x & 15) && !((uintptr_t)hid & 15) &&
           !((uintptr_t)d_hid & 15);
}

__global__ void cast_bf16_kernel(long long n, const float *__restrict__ src, __nv_bfloat16 *__restrict__ dst) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        dst[i] = __float2bfloat16_rn(src[i]);
}

// ---- warp-per-group forward (used when there are thousands of groups): no block barriers, one warp owns a title from
// logits to pooled vector (measured 0.26 vs 0.33 ms at title level; the same idea was slower for the backward pass).
// Valid for F <= 1024 and A <= 256 (register-resident per-lane slices); other shapes take the CTA-per-group kernels above.

template <bool kPers>
__global__ void __launch_bounds__(WPB * 32)
pool_fwd_warp_kernel(const float *__restrict__ x, const int *__restrict__ x_rows, const float *__restrict__ mask,
                     const float *__restrict__ hid, const float *__restrict__ w2, const float *__restrict__ b2,
                     const float *__restrict__ qh, int rows_per_query, const int *__restrict__ seg, long long R, int Lmax,
                     int F, int A, float *__restrict__ attn, float *__restrict__ pooled) {
    extern __shared__ float sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float *e = sm + warp * Lmax;
    const int F4 = F >> 2;
    const float4 *x4 = reinterpret_cast<const float4 *>(x);
    const long long wid = (long long)blockIdx.x * WPB + warp, nw = (long long)gridDim.x * WPB;
    float wreg[AQ];
    if (!kPers) {
#pragma unroll
        for (int i = 0; i < AQ; ++i) wreg[i] = (lane + 32 * i < A) ? w2[lane + 32 * i] : 0.f;
    }
    for (long long r = wid; r < R; r += nw) {
        const long long base = seg ? (long long)seg[r] : r * (long long)Lmax;
        const int L = seg ? seg[r + 1] - seg[r] : Lmax;
        if (kPers) {
            const float *wv = qh + (r / rows_per_query) * A;
#pragma unroll
            for (int i = 0; i < AQ; ++i) wreg[i] = (lane + 32 * i < A) ? wv[lane + 32 * i] : 0.f;
        }
        const float bias = kPers ? 0.f : b2[0];
        float tot = 0.f;
        // logits, four rows per iteration: all 4 x AQ loads of a lane are in flight before the first reduction (one row per
        // iteration left the warp waiting a full DRAM round trip per token: ncu issue utilisation 13 %)
        for (int l0 = 0; l0 < L; l0 += 4) {
            float acc[4], mval[4], h[4][AQ];
#pragma unroll
            for (int q = 0; q < 4; ++q) {                       // every load first (no branch between them) ...
                const int l = min(l0 + q, L - 1);
                mval[q] = (l0 + q < L) ? (mask ? mask[base + l] : 1.f) : 0.f;
                const float *hrow = hid + (base + l) * A;
#pragma unroll
                for (int i = 0; i < AQ; ++i) h[q][i] = (lane + 32 * i < A) ? hrow[lane + 32 * i] : 0.f;
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {                       // ... then the arithmetic
                float a = 0.f;
#pragma unroll
                for (int i = 0; i < AQ; ++i) a = fmaf(h[q][i], wreg[i], a);
                acc[q] = a;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
                for (int q = 0; q < 4; ++q) acc[q] += __shfl_xor_sync(0xffffffffu, acc[q], o);
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                if (l0 + q < L) {
                    const float ev = (mval[q] != 0.f) ? expf(acc[q] + bias) * mval[q] : 0.f;
                    if (lane == 0) e[l0 + q] = ev;
                    tot += ev;
                }
            }
        }
        __syncwarp();
        const float denom = tot + 1e-8f;
        for (int l = lane; l < L; l += 32) attn[base + l] = e[l] / denom;
        float4 acc[FQ];
#pragma unroll
        for (int c = 0; c < FQ; ++c) acc[c] = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int l = 0; l < L; ++l) {
            const float a = e[l] / denom;
            if (a == 0.f) continue;
            const long long row = x_rows ? (long long)x_rows[base + l] : base + l;
#pragma unroll
            for (int c = 0; c < FQ; ++c) {
                const int col = lane + 32 * c;
                if (col < F4) {
                    const float4 v = ldg_stream(x4 + row * F4 + col);
                    acc[c].x = fmaf(a, v.x, acc[c].x); acc[c].y = fmaf(a, v.y, acc[c].y);
                    acc[c].z = fmaf(a, v.z, acc[c].z); acc[c].w = fmaf(a, v.w, acc[c].w);
                }
            }
        }
#pragma unroll
        for (int c = 0; c < FQ; ++c) {
            const int col = lane + 32 * c;
            if (col < F4) reinterpret_cast<float4 *>(pooled)[r * F4 + col] = acc[c];
        }
        __syncwarp();
    }
}

__global__ void meanpool_fwd_kernel(const float *__restrict__ x, const float *__restrict__ mask, long long R, int L,
                                    int F, float *__restrict__ pooled) {
    for (long long r = blockIdx.x; r < R; r += gridDim.x) {
        float tot = 0.f;
        for (int l = 0; l < L; ++l) tot += mask[r * L + l];
        const float denom = tot + 1e-8f;
        for (int f = threadIdx.x; f < F; f += blockDim.x) {
            float acc = 0.f;
            for (int l = 0; l < L; ++l) acc = fmaf(x[(r * L + l) * F + f], mask[r * L + l], acc);
            pooled[r * F + f] = acc / denom;
        }
    }
}

// d_x[r,l,:] = d_pooled[r,:] * m[r,l] / (sum_l m + 1e-8)   (backward of layers.py:34-36; the mask carries no gradient)
__global__ void meanpool_bwd_kernel(const float *__restrict__ mask, const float *__restrict__ d_pooled, long long R, int L, int F,
                                    float *__restrict__ d_x) {
    for (long long r = blockIdx.x; r < R; r += gridDim.x) {
        float tot = 0.f;
        for (int l = 0; l < L; ++l) tot += mask[r * L + l];
        const float inv = 1.f / (tot + 1e-8f);
        for (int i = threadIdx.x; i < L * F; i += blockDim.x) {
            const int l = i / F, f = i - l * F;
            d_x[(r * L + l) * F + f] = d_pooled[r * F + f] * mask[r * L + l] * inv;
        }
    }
}

__global__ void collapse_mask_kernel(const float *__restrict__ mask, long long R, int L, float *__restrict__ out) {
    for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r < R; r += (long long)gridDim.x * blockDim.x) {
        float s = 0.f;
        for (int l = 0; l < L; ++l) s += mask[r * L + l];
        out[r] = fminf(fmaxf(s, 0.f), 1.f);
    }
}

// pool_bwd2_kernel: 2 CTAs of 256 threads per SM (register-heavy: 12 x 16-byte loads in flight per lane), one wave
static unsigned pool_grid2(long long R) {
    long long cap = 2LL * num_sms();
    return (unsigned)(R < cap ? (R < 1 ? 1 : R) : cap);
}

static unsigned pool_grid(long long R) {
    long long cap = 8LL * num_sms();
    return (unsigned)(R < cap ? (R < 1 ? 1 : R) : cap);
}

// a warp per group needs many groups to fill the machine (title level: thousands); few long groups (user level: one per
// impression) keep a whole CTA per group
static bool warp_path(long long R, int L, int F, int A) { return R >= 4096 && F <= 32 * 4 * FQ && A <= 32 * AQ && L <= 1024; }
static unsigned warp_grid(long long R) {
    long long blocks = cdiv(R, WPB), cap = 8LL * num_sms();
    return (unsigned)(blocks < 1 ? 1 : (blocks > cap ? cap : blocks));
}

static int check_pool(long long R, int L, int F, int A, const float *x) {
    if (R < 0 || L <= 0 || F <= 0 || A <= 0) return fail(XNRS_ERR_ARG, "%s: bad sizes", "pool");
    if (F % 4 != 0) return fail(XNRS_ERR_ARG, "%s: F must be a multiple of 4", "pool");
    if ((uintptr_t)x & 15) return fail(XNRS_ERR_ARG, "%s: x must be 16-byte aligned", "pool");
    return XNRS_OK;
}


// ---- pooling over per-ITEM logits (evaluation; rows U / U-naml with frozen weights) -------------------------------------
// The additive pooler's logit  w2 . tanh(fc1 h + b1) + b2  of a history slot depends only on the article in that slot, so
// with a pre-encoded catalogue it is computed ONCE per article (xnrs_gemm + xnrs_rowdot over the catalogue) instead of once
// per (user, slot): the per-user work left is this kernel — gather the slot's logit and vector, exp * mask, normalise by
// (sum + 1e-8), weighted sum.  Same values as layers.py:60-65 applied slot by slot.  One warp per user; HBM/L2-bound:
// L rows of T floats per user.
__global__ void __launch_bounds__(256)
rowdot_kernel(const float *__restrict__ x, const float *__restrict__ w, const float *__restrict__ b, long long n, int A,
              float *__restrict__ out) {
    const int lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= n) return;
    const float *xr = x + row * A;
    float acc = 0.f;
    for (int j = lane; j < A; j += 32) acc = fmaf(xr[j], w[j], acc);
    acc = warp_sum(acc);
    if (lane == 0) out[row] = acc + (b ? b[0] : 0.f);
}

template <int NV>       // up to NV float4 columns per lane: T <= 128 * NV
__global__ void __launch_bounds__(256)
logitpool_fwd_kernel(const float *__restrict__ table, const float *__restrict__ logit, const float *__restrict__ row_mask,
                     const int *__restrict__ ids, long long R, int L, int T4, float *__restrict__ attn,
                     float *__restrict__ pooled) {
    const int lane = threadIdx.x & 31;
    const long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= R) return;
    const int *idr = ids + r * L;
    // pass 1: e_l for the lane's slots (l = lane, lane + 32, ...), total
    float tot = 0.f;
    for (int l = lane; l < L; l += 32) {
        const int v = idr[l];
        const float m = row_mask ? row_mask[v] : 1.f;
        tot += (m != 0.f) ? expf(logit[v]) * m : 0.f;
    }
    tot = warp_sum(tot);
    const float inv = 1.f / (tot + 1e-8f);
    float4 acc[NV];
#pragma unroll
    for (int c = 0; c < NV; ++c) acc[c] = make_float4(0.f, 0.f, 0.f, 0.f);
    const float4 *t4 = reinterpret_cast<const float4 *>(table);
    for (int l0 = 0; l0 < L; l0 += 32) {
        const int l = l0 + lane;
        int v = 0;
        float a = 0.f;
        if (l < L) {
            v = idr[l];
            const float m = row_mask ? row_mask[v] : 1.f;
            a = (m != 0.f) ? expf(logit[v]) * m * inv : 0.f;
            if (attn) attn[r * L + l] = a;
        }
        const int cnt = min(32, L - l0);
#pragma unroll 4
        for (int j = 0; j < cnt; ++j) {
            const float aj = __shfl_sync(0xffffffffu, a, j);
            const int vj = __shfl_sync(0xffffffffu, v, j);
            if (aj == 0.f) continue;                              // warp-uniform: padded slots are never read
            const float4 *row = t4 + (long long)vj * T4;
#pragma unroll
            for (int c = 0; c < NV; ++c) {
                if (c * 32 + lane >= T4) break;
                const float4 x = __ldg(row + c * 32 + lane);
                acc[c].x = fmaf(aj, x.x, acc[c].x); acc[c].y = fmaf(aj, x.y, acc[c].y);
                acc[c].z = fmaf(aj, x.z, acc[c].z); acc[c].w = fmaf(aj, x.w, acc[c].w);
            }
        }
    }
    float4 *out = reinterpret_cast<float4 *>(pooled) + r * T4;
#pragma unroll
    for (int c = 0; c < NV; ++c)
        if (c * 32 + lane < T4) out[c * 32 + lane] = acc[c];
}


// backward of logitpool_fwd, one warp per group r (user).  With a_l = E_l / (sum + 1e-8):
//   da_l = <d_pooled_r, table[id_l]>,   dlogit_l = a_l (da_l - sum_j a_j da_j)
//   d_logit[id_l] += dlogit_l           (scalar reductions: many slots share an article)
//   d_table[id_l] += a_l d_pooled_r     (16-byte vector reductions)
// The (R, L, T) gradient of the gathered history never exists.
template <int NV>
__global__ void __launch_bounds__(256)
logitpool_bwd_kernel(const float *__restrict__ table, const int *__restrict__ ids, const float *__restrict__ attn,
                     const float *__restrict__ d_pooled, long long R, int L, int T4, float *__restrict__ d_logit,
                     float *__restrict__ d_table) {
    const int lane = threadIdx.x & 31;
    const long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= R) return;
    const int *idr = ids + r * L;
    const float *ar = attn + r * L;
    const float4 *t4 = reinterpret_cast<const float4 *>(table);
    float4 g[NV];
    const float4 *g4 = reinterpret_cast<const float4 *>(d_pooled) + r * T4;
#pragma unroll
    for (int c = 0; c < NV; ++c) g[c] = (c * 32 + lane < T4) ? g4[c * 32 + lane] : make_float4(0.f, 0.f, 0.f, 0.f);
    if (L <= 64) {
        // fast path: the lane keeps (id, a, da) of slots lane and lane + 32 in registers; every table row is read once
        float a_[2], da_[2] = {0.f, 0.f};
        int v_[2];
#pragma unroll
        for (int b = 0; b < 2; ++b) {
            const int l = b * 32 + lane;
            v_[b] = l < L ? idr[l] : 0;
            a_[b] = l < L ? ar[l] : 0.f;
        }
#pragma unroll
        for (int b = 0; b < 2; ++b) {
            const int cnt = min(32, L - 32 * b);
            for (int j0 = 0; j0 < cnt; j0 += 4) {
                float acc[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int j = min(j0 + q, cnt - 1);
                    const float aj = __shfl_sync(0xffffffffu, a_[b], j);
                    const int vj = __shfl_sync(0xffffffffu, v_[b], j);
                    float s = 0.f;
                    if (aj != 0.f) {                               // warp-uniform: padded slots are never read
                        const float4 *row = t4 + (long long)vj * T4;
#pragma unroll
                        for (int c = 0; c < NV; ++c) {
                            if (c * 32 + lane < T4) {
                                const float4 x = __ldg(row + c * 32 + lane);
                                s = fmaf(x.x, g[c].x, s); s = fmaf(x.y, g[c].y, s);
                                s = fmaf(x.z, g[c].z, s); s = fmaf(x.w, g[c].w, s);
                            }
                        }
                    }
                    acc[q] = s;
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) acc[q] += __shfl_xor_sync(0xffffffffu, acc[q], o);
                }
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    if (lane == j0 + q) da_[b] = acc[q];
            }
        }
        const float tot = warp_sum(a_[0] * da_[0] + a_[1] * da_[1]);
#pragma unroll
        for (int b = 0; b < 2; ++b) {
            if (a_[b] != 0.f) atomicAdd(d_logit + v_[b], a_[b] * (da_[b] - tot));
            const int cnt = min(32, L - 32 * b);
            for (int j = 0; j < cnt; ++j) {
                const float aj = __shfl_sync(0xffffffffu, a_[b], j);
                const int vj = __shfl_sync(0xffffffffu, v_[b], j);
                if (aj == 0.f) continue;
                float4 *drow = reinterpret_cast<float4 *>(d_table) + (long long)vj * T4;
#pragma unroll
                for (int c = 0; c < NV; ++c)
                    if (c * 32 + lane < T4)
                        atomicAdd(drow + c * 32 + lane, make_float4(aj * g[c].x, aj * g[c].y, aj * g[c].z, aj * g[c].w));
            }
        }
        return;
    }
    // long groups: pass 1 computes the group-wide sum_j a_j da_j, pass 2 recomputes da per slot and scatters
    float dot = 0.f;
    for (int l0 = 0; l0 < L; l0 += 32) {
        const int l = l0 + lane;
        const int v = l < L ? idr[l] : 0;
        const float a = l < L ? ar[l] : 0.f;
        const int cnt = min(32, L - l0);
        for (int j = 0; j < cnt; ++j) {
            const float aj = __shfl_sync(0xffffffffu, a, j);
            const int vj = __shfl_sync(0xffffffffu, v, j);
            if (aj == 0.f) continue;
            const float4 *row = t4 + (long long)vj * T4;
            float s = 0.f;
#pragma unroll
            for (int c = 0; c < NV; ++c) {
                if (c * 32 + lane < T4) {
                    const float4 x = __ldg(row + c * 32 + lane);
                    s = fmaf(x.x, g[c].x, s); s = fmaf(x.y, g[c].y, s);
                    s = fmaf(x.z, g[c].z, s); s = fmaf(x.w, g[c].w, s);
                }
            }
            dot = fmaf(aj, s, dot);                 // per-lane partial of a_j da_j
        }
    }
    const float tot = warp_sum(dot);
    for (int l0 = 0; l0 < L; l0 += 32) {
        const int l = l0 + lane;
        const int v = l < L ? idr[l] : 0;
        const float a = l < L ? ar[l] : 0.f;
        const int cnt = min(32, L - l0);
        for (int j = 0; j < cnt; ++j) {
            const float aj = __shfl_sync(0xffffffffu, a, j);
            const int vj = __shfl_sync(0xffffffffu, v, j);
            if (aj == 0.f) continue;
            const float4 *row = t4 + (long long)vj * T4;
            float4 *drow = reinterpret_cast<float4 *>(d_table) + (long long)vj * T4;
            float s = 0.f;
#pragma unroll
            for (int c = 0; c < NV; ++c) {
                if (c * 32 + lane < T4) {
                    const float4 x = __ldg(row + c * 32 + lane);
                    s = fmaf(x.x, g[c].x, s); s = fmaf(x.y, g[c].y, s);
                    s = fmaf(x.z, g[c].z, s); s = fmaf(x.w, g[c].w, s);
                    atomicAdd(drow + c * 32 + lane, make_float4(aj * g[c].x, aj * g[c].y, aj * g[c].z, aj * g[c].w));
                }
            }
            s = warp_sum(s);
            if (lane == 0) atomicAdd(d_logit + vj, aj * (s - tot));
        }
    }
}

// backward of the per-item logit  logit_v = <hid_v, w2> + b2  with hid = tanh(.):  d_hid[v,j] = d_logit[v] w2[j] (1 - hid^2),
// d_w2[j] += sum_v d_logit[v] hid[v,j],  d_b2 += sum_v d_logit[v]     (d_w2 / d_b2 accumulate: per-CTA partials, one atomic set)
__global__ void __launch_bounds__(256)
logit_bwd_kernel(const float *__restrict__ hid, const float *__restrict__ w2, const float *__restrict__ d_logit, long long n,
                 int A, float *__restrict__ d_hid, float *__restrict__ d_w2, float *__restrict__ d_b2) {
    extern __shared__ float sm[];           // [A] partial d_w2
    __shared__ float db;
    for (int j = threadIdx.x; j < A; j += blockDim.x) sm[j] = 0.f;
    if (threadIdx.x == 0) db = 0.f;
    __syncthreads();
    const long long rows_per = (n + gridDim.x - 1) / gridDim.x;
    const long long v0 = blockIdx.x * rows_per, v1 = min(n, v0 + rows_per);
    for (int j = threadIdx.x; j < A; j += blockDim.x) {
        const float w = w2[j];
        float gw = 0.f;
        for (long long v = v0; v < v1; ++v) {
            const float dl = d_logit[v];
            const float h = hid[v * A + j];
            gw = fmaf(dl, h, gw);
            d_hid[v * A + j] = dl * w * (1.f - h * h);
        }
        sm[j] = gw;
    }
    float s = 0.f;
    for (long long v = v0 + threadIdx.x; v < v1; v += blockDim.x) s += d_logit[v];
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0 && s != 0.f) atomicAdd(&db, s);
    __syncthreads();
    for (int j = threadIdx.x; j < A; j += blockDim.x) atomicAdd(d_w2 + j, sm[j]);
    if (threadIdx.x == 0) atomicAdd(d_b2, db);
}

}  // namespace xnrs

using namespace xnrs;

extern "C" int xnrs_addpool_fwd(const float *x, const int *x_rows, const float *mask, const float *hid,
                                const float *w2, const float *b2, const int *seg, long long R, int L, int F, int A,
                                float *attn, float *pooled, xnrs_stream_t st) {
    if (int e = check_pool(R, L, F, A, x)) return e;
    if (R == 0) return XNRS_OK;
    XNRS_REQUIRE(x && hid && w2 && b2 && attn && pooled, "null pointer");
    if (warp_path(R, L, F, A))
        pool_fwd_warp_kernel<false><<<warp_grid(R), WPB * 32, WPB * L * sizeof(float), STREAM(st)>>>(
            x, x_rows, mask, hid, w2, b2, nullptr, 1, seg, R, L, F, A, attn, pooled);
    else
        pool_fwd_kernel<false><<<pool_grid(R), POOL_THREADS, L * sizeof(float), STREAM(st)>>>(
            x, x_rows, mask, hid, w2, b2, nullptr, 1, seg, R, L, F, A, attn, pooled);
    XNRS_LAUNCHED();
    return XNRS_OK;
}

extern "C" int xnrs_addpool_bwd(const float *x, const int *x_rows, const float *mask, const float *hid,
                                const float *w2, const float *attn, const float *d_pooled, const float *d_attn,
                                const int *seg, long long R, int L, int F, int A, long long n_rows, float *d_hid, float *d_w2,
                                float *d_b2, float *d_x, float *d_b1, xnrs_stream_t st) {
    (void)mask;   // the mask is already folded into attn (masked rows have weight exactly 0)
    if (int e = check_pool(R, L, F, A, x)) return e;
    if (R == 0) return XNRS_OK;
    XNRS_REQUIRE(x && hid && w2 && attn && d_pooled && d_hid && d_w2 && d_b2, "null pointer");
    XNRS_REQUIRE(!(d_x && x_rows), "d_x is only defined for dense x");
    // XNRS_POOL_BWD_WARP=1 selects the warp-per-title kernel.  Measured on the CL step (8.7k ragged titles, 153k rows): it is
    // SLOWER than the CTA-per-title kernel (0.315 vs 0.231 ms fp32, 0.260 vs 0.227 ms bf16): phase 2 walks a title's rows one
    // by one per warp, where the CTA kernel spreads them over 256 threads — so the CTA kernel stays the default
    static int warp_bwd = -1;
    if (warp_bwd < 0) { const char *ev = getenv("XNRS_POOL_BWD_WARP"); warp_bwd = ev ? atoi(ev) : 0; }
    if (warp_bwd && !d_x && !d_attn && R >= 4096 && F % 4 == 0 && F <= 768 && A <= 32 * AQ && L <= 1024) {
        pool_bwd_warp_kernel<false><<<warp_grid(R), WPB * 32, WPB * L * sizeof(float), STREAM(st)>>>(
            x, x_rows, hid, w2, attn, d_pooled, seg, R, L, F, A, n_rows, d_hid, d_w2, d_b2, d_b1);
        XNRS_LAUNCHED();
        return XNRS_OK;
    }
    if (!d_x && !d_attn && pool_bwd2_ok(L, F, A, x, hid, d_hid, 4)) {
        pool_bwd2_kernel<false, false><<<pool_grid2(R), POOL_THREADS, (2 * F + 9 * L + 2 * A) * sizeof(float), STREAM(st)>>>(
            x, x_rows, hid, w2, attn, d_pooled, seg, R, L, F, n_rows, d_hid, nullptr, d_w2, d_b2, d_b1);
        XNRS_LAUNCHED();
        return XNRS_OK;
    }
    pool_bwd_kernel<false><<<pool_grid(R), POOL_THREADS, (2 * L + 2 * A) * sizeof(float), STREAM(st)>>>(
            x, x_rows, hid, w2, nullptr, 1, attn, d_pooled, d_attn, seg, R, L, F, A, n_rows, d_hid, d_w2, d_b2, nullptr, d_x, d_b1);
    XNRS_LAUNCHED();
    return XNRS_OK;
}

extern "C" int xnrs_addpool_bwd_bf16(const void *x, const int *x_rows, const void *hid, const float *w2, const float *attn,
                                     const float *d_pooled, const int *seg, long long R, int L, int F, int A, long long n_rows,
                                     void *d_hid, float *d_w2, float *d_b2, float *d_b1, xnrs_stream_t st) {
    if (R < 0 || L <= 0 || F <= 0 || A <= 0 || F % 8) return fail(XNRS_ERR_ARG, "%s: bad sizes (F % 8 == 0)", "xnrs_addpool_bwd_bf16");
    if (R == 0) return XNRS_OK;
    XNRS_REQUIRE(x && hid && w2 && attn && d_pooled && d_hid && d_w2 && d_b2, "null pointer");
    XNRS_REQUIRE(!((uintptr_t)x & 15) && !((uintptr_t)d_pooled & 15), "x and d_pooled must be 16-byte aligned");
    static int warp_bwd = -1;       // see xnrs_addpool_bwd: measured slower, off by default
    if (warp_bwd < 0) { const char *ev = getenv("XNRS_POOL_BWD_WARP"); warp_bwd = ev ? atoi(ev) : 0; }
    if (warp_bwd && R >= 4096 && F <= 768 && A <= 32 * AQ && L <= 1024) {
        pool_bwd_warp_kernel<true><<<warp_grid(R), WPB * 32, WPB * L * sizeof(float), STREAM(st)>>>(
            x, x_rows, hid, w2, attn, d_pooled, seg, R, L, F, A, n_rows, d_hid, d_w2, d_b2, d_b1);
        XNRS_LAUNCHED();
        return XNRS_OK;
    }
    if (pool_bwd2_ok(L, F, A, x, hid, d_hid, 2)) {
        pool_bwd2_kernel<true, false><<<pool_grid2(R), POOL_THREADS, (2 * F + 9 * L + 2 * A) * sizeof(float), STREAM(st)>>>(
            x, x_rows, hid, w2, attn, d_pooled, seg, R, L, F, n_rows, d_hid, nullptr, d_w2, d_b2, d_b1);
        XNRS_LAUNCHED();
        return XNRS_OK;
    }
    pool_bwd_bf16_kernel<<<pool_grid(R), POOL_THREADS, (2 * L + 2 * A) * sizeof(float), STREAM(st)>>>(
        reinterpret_cast<const __nv_bfloat16 *>(x), x_rows, reinterpret_cast<const __nv_bfloat16 *>(hid), w2, attn, d_pooled, seg, R,
        L, F, A, n_rows, reinterpret_cast<__nv_bfloat16 *>(d_hid), d_w2, d_b2, d_b1);
    XNRS_LAUNCHED();
    return XNRS_OK;
}

extern "C" int xnrs_addpool_bwd_split(const float *x, const int *x_rows, const float *hid, const float *w2, const float *attn,
                                      const float *d_pooled, const int *seg, long long R, int L, int F, int A, long long n_rows,
                                      void *d_hid_hi, void *d_hid_lo, float *d_w2, float *d_b2, float *d_b1, xnrs_stream_t st) {
    if (R < 0 || L <= 0 || F <= 0 || A <= 0) return fail(XNRS_ERR_ARG, "%s: bad sizes", "xnrs_addpool_bwd_split");
    if (R == 0) return XNRS_OK;
    XNRS_REQUIRE(x && hid && w2 && attn && d_pooled && d_hid_hi && d_hid_lo && d_w2 && d_b2, "null pointer");
    XNRS_REQUIRE(!((uintptr_t)d_pooled & 15) && !((uintptr_t)d_hid_hi & 7) && !((uintptr_t)d_hid_lo & 7), "alignment");
    if (!pool_bwd2_ok(L, F, A, x, hid, d_hid_hi, 4))
        return fail(XNRS_ERR_UNSUPPORTED, "%s: shape not covered (A = 256, F <= 768, L <= 256)", "xnrs_addpool_bwd_split");
    pool_bwd2_kernel<false, true><<<pool_grid2(R), POOL_THREADS, (2 * F + 9 * L + 2 * A) * sizeof(float), STREAM(st)>>>(
        x, x_rows, hid, w2, attn, d_pooled, seg, R, L, F, n_rows, d_hid_hi, d_hid_lo, d_w2, d_b2, d_b1);
    XNRS_LAUNCHED();
    return XNRS_OK;
}

extern "C" int xnrs_split_bf16(long long n, const float *src, void *hi, void *lo, int fp16, xnrs_stream_t st) {
    if (n <= 0) return XNRS_OK;
    XNRS_REQUIRE(src && hi && lo, "null pointer");
    long long b = cdiv(n, 256), cap = 16LL * num_sms();
    if (fp16)
        split_f16_kernel<<<(unsigned)(b > cap ? cap : b), 256, 0, STREAM(st)>>>(n, src, reinterpret_cast<__half *>(hi),
                                                                               reinterpret_cast<__half *>(lo));
    else
        split_bf16_kernel<<<(unsigned)(b > cap ? cap : b), 256, 0, STREAM(st)>>>(n, src, reinterpret_cast<__nv_bfloat16 *>(hi),
                                                                                reinterpret_cast<__nv_bfloat16 *>(lo));
    XNRS_LAUNCHED();
    return XNRS_OK;
}

extern "C" int xnrs_cast_bf16(long long n, const float *src, void *dst, xnrs_stream_t st) {
    if (n <= 0) return XNRS_OK;
    XNRS_REQUIRE(src && dst, "null pointer");
    long long b = cdiv(n, 256), cap = 16LL * num_sms();
    cast_bf16_kernel<<<(unsigned)(b > cap ? cap : b), 256, 0, STREAM(st)>>>(n, src, reinterpret_cast<__nv_bfloat16 *>(dst));
    XNRS_LAUNCHED();
    return XNRS_OK;
}

extern "C" int xnrs_perspool_fwd(const float *x, const int *x_rows, const float *mask, const float *hid,
                                 const float *qh, const int *seg, long long R, int L, int F, int A, int rows_per_query,
                                 float *attn, float *pooled, xnrs_stream_t st) {
    if (int e = check_pool(R, L, F, A, x)) return e;
    if (R == 0) return XNRS_OK;
    XNRS_REQUIRE(x && hid && qh && attn && pooled && rows_per_query > 0, "null pointer");
    if (warp_path(R, L, F, A))
        pool_fwd_warp_kernel<true><<<warp_grid(R), WPB * 32, WPB * L * sizeof(float), STREAM(st)>>>(
            x, x_rows, mask, hid, nullptr, nullptr, qh, rows_per_query, seg, R, L, F, A, attn, pooled);
    else
        pool_fwd_kernel<true><<<pool_grid(R), POOL_THREADS, L * sizeof(float), STREAM(st)>>>(
            x, x_rows, mask, hid, nullptr, nullptr, qh, rows_per_query, seg, R, L, F, A, attn, pooled);
    XNRS_LAUNCHED();
    return XNRS_OK;
}

extern "C" int xnrs_perspool_bwd(const float *x, const int *x_rows, const float *mask, const float *hid,
                                 const float *qh, const float *attn, const float *d_pooled, const int *seg, long long R,
                                 int L, int F, int A, int rows_per_query, long long n_rows, float *d_hid, float *d_qh, float *d_x,
                                 xnrs_stream_t st) {
    (void)mask;
    if (int e = check_pool(R, L, F, A, x)) return e;
    if (R == 0) return XNRS_OK;
    XNRS_REQUIRE(x && hid && qh && attn && d_pooled && d_hid && d_qh && rows_per_query > 0, "null pointer");
    XNRS_REQUIRE(!(d_x && x_rows), "d_x is only defined for dense x");
    pool_bwd_kernel<true><<<pool_grid(R), POOL_THREADS, (2 * L + 2 * A) * sizeof(float), STREAM(st)>>>(
            x, x_rows, hid, nullptr, qh, rows_per_query, attn, d_pooled, nullptr, seg, R, L, F, A, n_rows, d_hid, nullptr,
            nullptr, d_qh, d_x, nullptr);
    XNRS_LAUNCHED();
    return XNRS_OK;
}

extern "C" int xnrs_meanpool_fwd(const float *x, const float *mask, long long R, int L, int F, float *pooled,
                                 xnrs_stream_t st) {
    XNRS_REQUIRE(R >= 0 && L > 0 && F > 0, "bad sizes");
    if (R == 0) return XNRS_OK;
    XNRS_REQUIRE(x && mask && pooled, "null pointer");
    meanpool_fwd_kernel<<<pool_grid(R), 128, 0, STREAM(st)>>>(x, mask, R, L, F, pooled);
    XNRS_LAUNCHED();
    return XNRS_OK;
}

extern "C" int xnrs_meanpool_bwd(const float *mask, const float *d_pooled, long long R, int L, int F, float *d_x,
                                 xnrs_stream_t st) {
    XNRS_REQUIRE(R >= 0 && L > 0 && F > 0, "bad sizes");
    if (R == 0) return XNRS_OK;
    XNRS_REQUIRE(mask && d_pooled && d_x, "null pointer");
    meanpool_bwd_kernel<<<pool_grid(R), 256, 0, STREAM(st)>>>(mask, d_pooled, R, L, F, d_x);
    XNRS_LAUNCHED();
    return XNRS_OK;
}

extern "C" int xnrs_collapse_mask(const float *mask, long long R, int L, float *out, xnrs_stream_t st) {
    XNRS_REQUIRE(R >= 0 && L > 0, "bad sizes");
    if (R == 0) return XNRS_OK;
    XNRS_REQUIRE(mask && out, "null pointer");
    collapse_mask_kernel<<<(unsigned)cdiv(R, 256), 256, 0, STREAM(st)>>>(mask, R, L, out);
    XNRS_LAUNCHED();
    return XNRS_OK;
}

extern "C" int xnrs_rowdot(const float *x, const float *w, const float *b, long long n, int A, float *out,
                           xnrs_stream_t st) {
    XNRS_REQUIRE(n >= 0 && A > 0, "bad sizes");
    if (n == 0) return XNRS_OK;
    XNRS_REQUIRE(x && w && out, "null pointer");
    rowdot_kernel<<<(unsigned)cdiv(n, 8), 256, 0, STREAM(st)>>>(x, w, b, n, A, out);
    XNRS_LAUNCHED();
    return XNRS_OK;
}

extern "C" int xnrs_logitpool_fwd(const float *table, long long V, int T, const float *logit, const float *row_mask,
                                  const int *ids, long long R, int L, float *attn, float *pooled, xnrs_stream_t st) {
    XNRS_REQUIRE(R >= 0 && L > 0 && V > 0, "bad sizes");
    XNRS_REQUIRE(T > 0 && T % 4 == 0 && T <= 1024, "T must be a multiple of 4, at most 1024");
    if (R == 0) return XNRS_OK;
    XNRS_REQUIRE(table && logit && ids && pooled, "null pointer");
    XNRS_REQUIRE((((uintptr_t)table | (uintptr_t)pooled) & 15) == 0, "16-byte alignment");
    const unsigned grid = (unsigned)cdiv(R, 8);
    const int T4 = T / 4;
#define XNRS_LP(NV) logitpool_fwd_kernel<NV><<<grid, 256, 0, STREAM(st)>>>(table, logit, row_mask, ids, R, L, T4, attn, pooled)
    if (T4 <= 32) XNRS_LP(1);
    else if (T4 <= 64) XNRS_LP(2);
    else if (T4 <= 128) XNRS_LP(4);
    else XNRS_LP(8);
#undef XNRS_LP
    XNRS_LAUNCHED();
    return XNRS_OK;
}

extern "C" int xnrs_logitpool_bwd(const float *table, long long V, int T, const int *ids, const float *attn,
                                  const float *d_pooled, long long R, int L, float *d_logit, float *d_table,
                                  xnrs_stream_t st) {
    XNRS_REQUIRE(R >= 0 && L > 0 && V > 0, "bad sizes");
    XNRS_REQUIRE(T > 0 && T % 4 == 0 && T <= 1024, "T must be a multiple of 4, at most 1024");
    if (R == 0) return XNRS_OK;
    XNRS_REQUIRE(table && ids && attn && d_pooled && d_logit && d_table, "null pointer");
    XNRS_REQUIRE((((uintptr_t)table | (uintptr_t)d_pooled | (uintptr_t)d_table) & 15) == 0, "16-byte alignment");
    const unsigned grid = (unsigned)cdiv(R, 8);
    const int T4 = T / 4;
#define XNRS_LPB(NV) logitpool_bwd_kernel<NV><<<grid, 256, 0, STREAM(st)>>>(table, ids, attn, d_pooled, R, L, T4, d_logit, d_table)
    if (T4 <= 32) XNRS_LPB(1);
    else if (T4 <= 64) XNRS_LPB(2);
    else if (T4 <= 128) XNRS_LPB(4);
    else XNRS_LPB(8);
#undef XNRS_LPB
    XNRS_LAUNCHED();
    return XNRS_OK;
}

extern "C" int xnrs_logit_bwd(const float *hid, const float *w2, const float *d_logit, long long n, int A, float *d_hid,
                              float *d_w2, float *d_b2, xnrs_stream_t st) {
    XNRS_REQUIRE(n >= 0 && A > 0, "bad sizes");
    if (n == 0) return XNRS_OK;
    XNRS_REQUIRE(hid && w2 && d_logit && d_hid && d_w2 && d_b2, "null pointer");
    long long blocks = cdiv(n, 32), cap = 2LL * num_sms();
    logit_bwd_kernel<<<(unsigned)(blocks < cap ? blocks : cap), 256, A * sizeof(float), STREAM(st)>>>(hid, w2, d_logit, n, A, d_hid,
                                                                                                    d_w2, d_b2);
    XNRS_LAUNCHED();
    return XNRS_OK;
}
