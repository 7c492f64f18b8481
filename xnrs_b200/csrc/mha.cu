// Row M: the attention core of layers.MultiHeadAttention (layers.py:133-151) for short sequences
// (L = 25..50 tokens / history items): softmax(q k^T / sqrt(dk)) v per (title, head), with the
// reference's QUERY-axis masking (masked query rows attend uniformly; keys are never masked) and
// dropout applied to the normalised weights.
//
// Layout: one warp owns one (title, head).  K and V head slices are staged in shared memory; each
// lane owns whole query rows (forward, dq) or whole key rows (dk, dv), so every softmax reduction is
// lane-local and shared-memory reads are warp-wide broadcasts.  The Q/K/V/out projections are GEMMs
// and live in xnrs_gemm.
#include "common.cuh"

namespace xnrs {

__device__ __forceinline__ float keep_factor(const float *__restrict__ keep, float p_drop, unsigned long long seed,
                                             long long idx) {
    if (keep) return keep[idx];
    if (p_drop <= 0.f) return 1.f;
    // counter-based generator (splitmix64 of seed + index): same draw in forward and backward
    unsigned long long z = seed + (unsigned long long)(idx + 1) * 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    float u = (float)(z >> 40) * (1.0f / 16777216.0f);
    return u >= p_drop ? 1.f : 0.f;
}

template <int DK>
__device__ __forceinline__ void load_row(float (&dst)[DK], const float *__restrict__ src) {
#pragma unroll
    for (int d = 0; d < DK; d += 4) {
        float4 v = *reinterpret_cast<const float4 *>(src + d);
        dst[d] = v.x; dst[d + 1] = v.y; dst[d + 2] = v.z; dst[d + 3] = v.w;
    }
}

template <int DK>
__device__ __forceinline__ float dot_smem(const float (&a)[DK], const float *__restrict__ s) {
    float acc = 0.f;
#pragma unroll
    for (int d = 0; d < DK; d += 4) {
        float4 v = *reinterpret_cast<const float4 *>(s + d);
        acc = fmaf(a[d], v.x, acc); acc = fmaf(a[d + 1], v.y, acc);
        acc = fmaf(a[d + 2], v.z, acc); acc = fmaf(a[d + 3], v.w, acc);
    }
    return acc;
}

// stage the head slice of a (R,L,ld) tensor for title r into shared memory [L][DK]
template <int DK>
__device__ __forceinline__ void stage(float *__restrict__ dst, const float *__restrict__ src, long long r, int L,
                                      long long ld, int head, int lane) {
    constexpr int C = DK / 4;
    for (int i = lane; i < L * C; i += 32) {
        int l = i / C, c = i - l * C;
        reinterpret_cast<float4 *>(dst)[i] =
            *reinterpret_cast<const float4 *>(src + (r * L + l) * ld + head * DK + c * 4);
    }
}

template <int DK>
__global__ void mha_fwd_kernel(const float *__restrict__ q, const float *__restrict__ k, const float *__restrict__ v,
                               long long ld, const float *__restrict__ mask, long long R, int L, int h,
                               const float *__restrict__ keep, float p_drop, unsigned long long seed,
                               float *__restrict__ o, float *__restrict__ lse) {
    extern __shared__ __align__(16) float smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const long long w = (long long)blockIdx.x * wpb + warp;
    if (w >= R * h) return;
    const long long r = w / h;
    const int head = (int)(w - r * h);
    float *Ks = smem + (size_t)warp * 2 * L * DK, *Vs = Ks + L * DK;
    stage<DK>(Ks, k, r, L, ld, head, lane);
    stage<DK>(Vs, v, r, L, ld, head, lane);
    __syncwarp();
    const float scale = rsqrtf((float)DK) , inv_keep = 1.f / (1.f - p_drop);
    const float sc = 1.0f / sqrtf((float)DK);
    (void)scale;
    for (int i = lane; i < L; i += 32) {
        float qi[DK], acc[DK];
        load_row<DK>(qi, q + (r * L + i) * ld + head * DK);
#pragma unroll
        for (int d = 0; d < DK; ++d) acc[d] = 0.f;
        const bool masked = mask && mask[r * L + i] == 0.f;
        float m = -INFINITY, l = 0.f;
        const long long kbase = ((r * h + head) * L + i) * L;
        for (int j = 0; j < L; ++j) {
            const float s = masked ? -1e9f : dot_smem<DK>(qi, Ks + j * DK) * sc;
            const float mn = fmaxf(m, s);
            const float corr = expf(m - mn), e = expf(s - mn);
            l = l * corr + e;
            const float kf = keep_factor(keep, p_drop, seed, kbase + j) * inv_keep;
            const float ek = e * kf;
            const float *vj = Vs + j * DK;
#pragma unroll
            for (int d = 0; d < DK; d += 4) {
                float4 vv = *reinterpret_cast<const float4 *>(vj + d);
                acc[d] = fmaf(ek, vv.x, acc[d] * corr); acc[d + 1] = fmaf(ek, vv.y, acc[d + 1] * corr);
                acc[d + 2] = fmaf(ek, vv.z, acc[d + 2] * corr); acc[d + 3] = fmaf(ek, vv.w, acc[d + 3] * corr);
            }
            m = mn;
        }
        const float inv_l = 1.f / l;
        float *orow = o + (r * L + i) * ld + head * DK;
#pragma unroll
        for (int d = 0; d < DK; d += 4)
            *reinterpret_cast<float4 *>(orow + d) =
                make_float4(acc[d] * inv_l, acc[d + 1] * inv_l, acc[d + 2] * inv_l, acc[d + 3] * inv_l);
        lse[(r * h + head) * L + i] = m + logf(l);
    }
}

template <int DK>
__global__ void mha_bwd_kernel(const float *__restrict__ q, const float *__restrict__ k, const float *__restrict__ v,
                               const float *__restrict__ o, const float *__restrict__ d_o, long long ld,
                               const float *__restrict__ mask, const float *__restrict__ lse, long long R, int L,
                               int h, const float *__restrict__ keep, float p_drop, unsigned long long seed,
                               float *__restrict__ dq, float *__restrict__ dk_, float *__restrict__ dv) {
    extern __shared__ __align__(16) float smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const long long w = (long long)blockIdx.x * wpb + warp;
    if (w >= R * h) return;
    const long long r = w / h;
    const int head = (int)(w - r * h);
    float *Ks = smem + (size_t)warp * (size_t)((4 * L * DK + 3 * L + 3) & ~3);     // 16-byte aligned per-warp slab
    float *Vs = Ks + L * DK, *Qs = Vs + L * DK, *Gs = Qs + L * DK;     // Gs = d_o
    float *Ls = Gs + L * DK, *Ds = Ls + L, *Ms = Ds + L;               // lse, D_i, masked flag
    stage<DK>(Ks, k, r, L, ld, head, lane);
    stage<DK>(Vs, v, r, L, ld, head, lane);
    stage<DK>(Qs, q, r, L, ld, head, lane);
    stage<DK>(Gs, d_o, r, L, ld, head, lane);
    const float sc = 1.0f / sqrtf((float)DK), inv_keep = 1.f / (1.f - p_drop), invL = 1.f / (float)L;
    const long long kb0 = (r * h + head) * (long long)L * L;
    // D_i = <d_o_i, o_i>
    for (int i = lane; i < L; i += 32) {
        float oi[DK];
        load_row<DK>(oi, o + (r * L + i) * ld + head * DK);
        float dsum = 0.f;
        const float *g = d_o + (r * L + i) * ld + head * DK;
#pragma unroll
        for (int d = 0; d < DK; ++d) dsum = fmaf(oi[d], g[d], dsum);
        Ds[i] = dsum;
        Ls[i] = lse[(r * h + head) * L + i];
        Ms[i] = (mask && mask[r * L + i] == 0.f) ? 1.f : 0.f;
    }
    __syncwarp();
    // pass A: lane = query row -> dq
    for (int i = lane; i < L; i += 32) {
        float qi[DK], gi[DK], acc[DK];
#pragma unroll
        for (int d = 0; d < DK; ++d) { qi[d] = Qs[i * DK + d]; gi[d] = Gs[i * DK + d]; acc[d] = 0.f; }
        const bool masked = Ms[i] != 0.f;
        if (!masked) {
            const float li = Ls[i], Di = Ds[i];
            for (int j = 0; j < L; ++j) {
                const float s = dot_smem<DK>(qi, Ks + j * DK) * sc;
                const float p = expf(s - li);
                const float kf = keep_factor(keep, p_drop, seed, kb0 + (long long)i * L + j) * inv_keep;
                const float ds = p * (kf * dot_smem<DK>(gi, Vs + j * DK) - Di) * sc;
                const float *kj = Ks + j * DK;
#pragma unroll
                for (int d = 0; d < DK; ++d) acc[d] = fmaf(ds, kj[d], acc[d]);
            }
        }
        float *out = dq + (r * L + i) * ld + head * DK;
#pragma unroll
        for (int d = 0; d < DK; d += 4)
            *reinterpret_cast<float4 *>(out + d) = make_float4(acc[d], acc[d + 1], acc[d + 2], acc[d + 3]);
    }
    // pass B1: lane = key row -> dv_j = sum_i P~_ij d_o_i
    for (int j = lane; j < L; j += 32) {
        float kj[DK], acc[DK];
#pragma unroll
        for (int d = 0; d < DK; ++d) { kj[d] = Ks[j * DK + d]; acc[d] = 0.f; }
        for (int i = 0; i < L; ++i) {
            const bool masked = Ms[i] != 0.f;
            const float p = masked ? invL : expf(dot_smem<DK>(kj, Qs + i * DK) * sc - Ls[i]);
            const float pk = p * keep_factor(keep, p_drop, seed, kb0 + (long long)i * L + j) * inv_keep;
            const float *g = Gs + i * DK;
#pragma unroll
            for (int d = 0; d < DK; ++d) acc[d] = fmaf(pk, g[d], acc[d]);
        }
        float *out = dv + (r * L + j) * ld + head * DK;
#pragma unroll
        for (int d = 0; d < DK; d += 4)
            *reinterpret_cast<float4 *>(out + d) = make_float4(acc[d], acc[d + 1], acc[d + 2], acc[d + 3]);
    }
    // pass B2: lane = key row -> dk_j = scale * sum_i ds_ij q_i
    for (int j = lane; j < L; j += 32) {
        float kj[DK], vj[DK], acc[DK];
#pragma unroll
        for (int d = 0; d < DK; ++d) { kj[d] = Ks[j * DK + d]; vj[d] = Vs[j * DK + d]; acc[d] = 0.f; }
        for (int i = 0; i < L; ++i) {
            if (Ms[i] != 0.f) continue;           // masked_fill blocks the gradient to q and k
            const float *qi = Qs + i * DK;
            const float p = expf(dot_smem<DK>(kj, qi) * sc - Ls[i]);
            const float kf = keep_factor(keep, p_drop, seed, kb0 + (long long)i * L + j) * inv_keep;
            const float ds = p * (kf * dot_smem<DK>(vj, Gs + i * DK) - Ds[i]) * sc;
#pragma unroll
            for (int d = 0; d < DK; ++d) acc[d] = fmaf(ds, qi[d], acc[d]);
        }
        float *out = dk_ + (r * L + j) * ld + head * DK;
#pragma unroll
        for (int d = 0; d < DK; d += 4)
            *reinterpret_cast<float4 *>(out + d) = make_float4(acc[d], acc[d + 1], acc[d + 2], acc[d + 3]);
    }
}

template <int DK>
static int launch_fwd(const float *q, const float *k, const float *v, long long ld, const float *mask, long long R,
                      int L, int h, const float *keep, float p_drop, unsigned long long seed, float *o, float *lse,
                      cudaStream_t st) {
    const int wpb = 4;
    size_t smem = (size_t)wpb * 2 * L * DK * sizeof(float);
    if (smem > 220 * 1024) return fail(XNRS_ERR_UNSUPPORTED, "%s: sequence too long for shared memory", "xnrs_mha_fwd");
    cudaFuncSetAttribute(mha_fwd_kernel<DK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    mha_fwd_kernel<DK><<<(unsigned)cdiv(R * h, wpb), wpb * 32, smem, st>>>(q, k, v, ld, mask, R, L, h, keep, p_drop,
                                                                           seed, o, lse);
    XNRS_LAUNCHED();
    return XNRS_OK;
}

template <int DK>
static int launch_bwd(const float *q, const float *k, const float *v, const float *o, const float *d_o, long long ld,
                      const float *mask, const float *lse, long long R, int L, int h, const float *keep, float p_drop,
                      unsigned long long seed, float *dq, float *dk_, float *dv, cudaStream_t st) {
    const int wpb = 2;
    size_t smem = (size_t)wpb * (size_t)((4 * L * DK + 3 * L + 3) & ~3) * sizeof(float);
    if (smem > 220 * 1024) return fail(XNRS_ERR_UNSUPPORTED, "%s: sequence too long for shared memory", "xnrs_mha_bwd");
    cudaFuncSetAttribute(mha_bwd_kernel<DK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    mha_bwd_kernel<DK><<<(unsigned)cdiv(R * h, wpb), wpb * 32, smem, st>>>(q, k, v, o, d_o, ld, mask, lse, R, L, h,
                                                                           keep, p_drop, seed, dq, dk_, dv);
    XNRS_LAUNCHED();
    return XNRS_OK;
}

}  // namespace xnrs

using namespace xnrs;

#define XNRS_DK_DISPATCH(CALL)                                                                         \
    switch (dk) {                                                                                      \
        case 4: return CALL(4);                                                                        \
        case 8: return CALL(8);                                                                        \
        case 16: return CALL(16);                                                                      \
        case 32: return CALL(32);                                                                      \
        case 48: return CALL(48);                                                                      \
        case 64: return CALL(64);                                                                      \
        default: return fail(XNRS_ERR_UNSUPPORTED, "%s: head dim must be one of 4,8,16,32,48,64", __func__); \
    }

extern "C" int xnrs_mha_fwd(const float *q, const float *k, const float *v, long long ld, const float *mask,
                            long long R, int L, int h, int dk, const float *keep, float p_drop,
                            unsigned long long seed, float *o, float *lse, xnrs_stream_t st) {
    XNRS_REQUIRE(R >= 0 && L > 0 && h > 0 && ld >= (long long)h * dk && ld % 4 == 0, "bad sizes");
    XNRS_REQUIRE(p_drop >= 0.f && p_drop < 1.f, "p_drop in [0,1)");
    if (R == 0) return XNRS_OK;
    XNRS_REQUIRE(q && k && v && o && lse, "null pointer");
    XNRS_REQUIRE((((uintptr_t)q | (uintptr_t)k | (uintptr_t)v | (uintptr_t)o) & 15) == 0, "16-byte alignment");
#define CALL(D) launch_fwd<D>(q, k, v, ld, mask, R, L, h, keep, p_drop, seed, o, lse, STREAM(st))
    XNRS_DK_DISPATCH(CALL)
#undef CALL
}

extern "C" int xnrs_mha_bwd(const float *q, const float *k, const float *v, const float *o, const float *d_o,
                            long long ld, const float *mask, const float *lse, long long R, int L, int h, int dk,
                            const float *keep, float p_drop, unsigned long long seed, float *dq, float *dk_, float *dv,
                            xnrs_stream_t st) {
    XNRS_REQUIRE(R >= 0 && L > 0 && h > 0 && ld >= (long long)h * dk && ld % 4 == 0, "bad sizes");
    if (R == 0) return XNRS_OK;
    XNRS_REQUIRE(q && k && v && o && d_o && lse && dq && dk_ && dv, "null pointer");
#define CALL(D) launch_bwd<D>(q, k, v, o, d_o, ld, mask, lse, R, L, h, keep, p_drop, seed, dq, dk_, dv, STREAM(st))
    XNRS_DK_DISPATCH(CALL)
#undef CALL
}
