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

// asynchronous variant (cp.async 16 B, L1-bypassing): every lane issues all of its copies back to back, so the DRAM/L2
// latency is paid once per staged matrix set instead of once per loop iteration (ncu on the synchronous version:
// long-scoreboard stalls dominated and the issue slots were 30 % busy); stage_wait() then joins them
template <int DK>
__device__ __forceinline__ void stage_async(float *__restrict__ dst, const float *__restrict__ src, long long r, int L,
                                            long long ld, int head, int lane) {
    constexpr int C = DK / 4;
    const uint32_t d0 = (uint32_t)__cvta_generic_to_shared(dst);
    for (int i = lane; i < L * C; i += 32) {
        int l = i / C, c = i - l * C;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d0 + i * 16),
                     "l"(src + (r * L + l) * ld + head * DK + c * 4) : "memory");
    }
}
__device__ __forceinline__ void stage_wait() {
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncwarp();
}

// ---- fast path helpers -----------------------------------------------------------------------------------------
// Packed fp32 FMA (Blackwell FFMA2): two independent fp32 FMAs per instruction, exact IEEE fp32 each.  A scalar 3-register
// FFMA issues every other cycle per scheduler, so the packed form is what reaches the fp32 peak.
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;"
        : "=l"(d)
        : "l"(*reinterpret_cast<unsigned long long *>(&a)), "l"(*reinterpret_cast<unsigned long long *>(&b)),
          "l"(*reinterpret_cast<unsigned long long *>(&c)));
    return *reinterpret_cast<float2 *>(&d);
}

// 32-bit counter hash (murmur3 finaliser) for the in-kernel dropout draw: (per-head key, i*L + j) -> keep/drop.
// Forward and backward call it with the same arguments, so the draw is reproduced, not stored.
__device__ __forceinline__ uint32_t head_key(unsigned long long seed, long long head_index) {
    unsigned long long z = seed + (unsigned long long)(head_index + 1) * 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return (uint32_t)(z ^ (z >> 31));
}
__device__ __forceinline__ float keep_fast(const float *__restrict__ keep, float p_drop, uint32_t key, int ij) {
    if (keep) return keep[ij];
    if (p_drop <= 0.f) return 1.f;
    uint32_t h = key + (uint32_t)ij * 0x9E3779B9u;
    h ^= h >> 16; h *= 0x85EBCA6Bu; h ^= h >> 13; h *= 0xC2B2AE35u; h ^= h >> 16;
    return (float)(h >> 8) * (1.0f / 16777216.0f) >= p_drop ? 1.f : 0.f;
}

// <a, s> with a in registers (DK/2 packed pairs) and s a shared-memory row (warp-wide broadcast reads); four independent
// accumulator chains hide the FMA latency
template <int DK>
__device__ __forceinline__ float dot2(const float2 (&a)[DK / 2], const float *__restrict__ s) {
    float2 acc[4] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
#pragma unroll
    for (int c = 0; c < DK / 4; ++c) {
        const float4 v = *reinterpret_cast<const float4 *>(s + 4 * c);
        acc[(2 * c) & 3] = ffma2(a[2 * c], make_float2(v.x, v.y), acc[(2 * c) & 3]);
        acc[(2 * c + 1) & 3] = ffma2(a[2 * c + 1], make_float2(v.z, v.w), acc[(2 * c + 1) & 3]);
    }
    return ((acc[0].x + acc[0].y) + (acc[1].x + acc[1].y)) + ((acc[2].x + acc[2].y) + (acc[3].x + acc[3].y));
}
// acc += w * s
template <int DK>
__device__ __forceinline__ void axpy2(float2 (&acc)[DK / 2], float w, const float *__restrict__ s) {
    const float2 ww = make_float2(w, w);
#pragma unroll
    for (int c = 0; c < DK / 4; ++c) {
        const float4 v = *reinterpret_cast<const float4 *>(s + 4 * c);
        acc[2 * c] = ffma2(ww, make_float2(v.x, v.y), acc[2 * c]);
        acc[2 * c + 1] = ffma2(ww, make_float2(v.z, v.w), acc[2 * c + 1]);
    }
}
template <int DK>
__device__ __forceinline__ void load_row2(float2 (&dst)[DK / 2], const float *__restrict__ src) {
#pragma unroll
    for (int c = 0; c < DK / 4; ++c) {
        const float4 v = *reinterpret_cast<const float4 *>(src + 4 * c);
        dst[2 * c] = make_float2(v.x, v.y);
        dst[2 * c + 1] = make_float2(v.z, v.w);
    }
}
template <int DK>
__device__ __forceinline__ void store_row2(float *__restrict__ dst, const float2 (&src)[DK / 2], float scale) {
#pragma unroll
    for (int c = 0; c < DK / 4; ++c)
        *reinterpret_cast<float4 *>(dst + 4 * c) = make_float4(src[2 * c].x * scale, src[2 * c].y * scale,
                                                               src[2 * c + 1].x * scale, src[2 * c + 1].y * scale);
}

// Forward, one warp per (title, head): K and V head slices in shared memory, lane = query row.
//   pass 1: s_j = <q_i, k_j> / sqrt(dk) for all j -> per-warp score buffer St[j][lane] (conflict-free), running max
//   pass 2: e_j = exp(s_j - max), l += e_j, o_i += (e_j * keep_ij / (1-p)) v_j      (two-pass softmax: no rescaling)
template <int DK>
__global__ void __launch_bounds__(128) mha_fwd_fast_kernel(const float *__restrict__ q, const float *__restrict__ k,
                                                            const float *__restrict__ v, long long ld,
                                                            const float *__restrict__ mask, long long R, int L, int h,
                                                            const float *__restrict__ keep, float p_drop,
                                                            unsigned long long seed, float *__restrict__ o,
                                                            float *__restrict__ lse) {
    extern __shared__ __align__(16) float smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const long long w = (long long)blockIdx.x * wpb + warp;
    if (w >= R * h) return;
    const long long r = w / h;
    const int head = (int)(w - r * h);
    float *Ks = smem + (size_t)warp * (size_t)(2 * L * DK + 32 * L), *Vs = Ks + L * DK, *St = Vs + L * DK;
    stage_async<DK>(Ks, k, r, L, ld, head, lane);
    stage_async<DK>(Vs, v, r, L, ld, head, lane);
    const float sc = 1.0f / sqrtf((float)DK), inv_keep = 1.f / (1.f - p_drop);
    const uint32_t key = head_key(seed, w);
    const float *keep_h = keep ? keep + w * (long long)L * L : nullptr;
    stage_wait();
    for (int i = lane; i < L; i += 32) {
        float2 qi[DK / 2], acc[DK / 2];
        load_row2<DK>(qi, q + (r * L + i) * ld + head * DK);
        const bool masked = mask && mask[r * L + i] == 0.f;
#pragma unroll
        for (int d = 0; d < DK / 2; ++d) acc[d] = make_float2(0.f, 0.f);
        float m = -1e9f, l;
        if (!masked) {
            m = -INFINITY;
#pragma unroll 2
            for (int j = 0; j < L; ++j) {
                const float s = dot2<DK>(qi, Ks + j * DK) * sc;
                St[j * 32 + lane] = s;
                m = fmaxf(m, s);
            }
            l = 0.f;
#pragma unroll 2
            for (int j = 0; j < L; ++j) {
                const float e = __expf(St[j * 32 + lane] - m);
                l += e;
                axpy2<DK>(acc, e * keep_fast(keep_h, p_drop, key, i * L + j) * inv_keep, Vs + j * DK);
            }
        } else {                        // masked_fill(-1e9) on the whole row: exactly uniform weights 1/L
            l = (float)L;
#pragma unroll 2
            for (int j = 0; j < L; ++j)
                axpy2<DK>(acc, keep_fast(keep_h, p_drop, key, i * L + j) * inv_keep, Vs + j * DK);
        }
        store_row2<DK>(o + (r * L + i) * ld + head * DK, acc, 1.f / l);
        lse[w * L + i] = m + logf(l);
    }
}

// Backward, one warp per (title, head).  K, V, Q, dO head slices and the two L x L matrices P~ (dropped-out weights) and
// dS live in shared memory:
//   pass A (lane = query row i): p_ij from the saved log-sum-exp, dP_ij = <dO_i, v_j>, dS_ij = p_ij (keep_ij dP_ij - D_i) / sqrt(dk),
//                                dq_i += dS_ij k_j; P~ and dS are written once (row stride LP odd: conflict-free)
//   pass B (lane = key row j):   dv_j = sum_i P~_ij dO_i,  dk_j = sum_i dS_ij q_i    (no recomputation of the scores)
template <int DK>
__global__ void __launch_bounds__(128) mha_bwd_fast_kernel(const float *__restrict__ q, const float *__restrict__ k,
                                                            const float *__restrict__ v, const float *__restrict__ o,
                                                            const float *__restrict__ d_o, long long ld,
                                                            const float *__restrict__ mask,
                                                            const float *__restrict__ lse, long long R, int L, int h,
                                                            const float *__restrict__ keep, float p_drop,
                                                            unsigned long long seed, float *__restrict__ dq,
                                                            float *__restrict__ dk_, float *__restrict__ dv) {
    extern __shared__ __align__(16) float smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const long long w = (long long)blockIdx.x * wpb + warp;
    if (w >= R * h) return;
    const long long r = w / h;
    const int head = (int)(w - r * h);
    const int LP = L | 1;
    float *Ks = smem + (size_t)warp * (size_t)((4 * L * DK + 2 * L * LP + 3) & ~3);
    float *Vs = Ks + L * DK, *Qs = Vs + L * DK, *Gs = Qs + L * DK;     // Gs = d_o
    float *Ps = Gs + L * DK, *Ss = Ps + L * LP;                         // P~ and dS, [i][j] with row stride LP
    stage_async<DK>(Ks, k, r, L, ld, head, lane);
    stage_async<DK>(Vs, v, r, L, ld, head, lane);
    stage_async<DK>(Qs, q, r, L, ld, head, lane);
    stage_async<DK>(Gs, d_o, r, L, ld, head, lane);
    stage_wait();
    const float sc = 1.0f / sqrtf((float)DK), inv_keep = 1.f / (1.f - p_drop), invL = 1.f / (float)L;
    const uint32_t key = head_key(seed, w);
    const float *keep_h = keep ? keep + w * (long long)L * L : nullptr;
    for (int i = lane; i < L; i += 32) {
        float2 qi[DK / 2], gi[DK / 2], acc[DK / 2];
        load_row2<DK>(qi, Qs + i * DK);
        load_row2<DK>(gi, Gs + i * DK);
        const bool masked = mask && mask[r * L + i] == 0.f;
        float *Pi = Ps + i * LP, *Si = Ss + i * LP;
        if (!masked) {
            float2 oi[DK / 2];
            load_row2<DK>(oi, o + (r * L + i) * ld + head * DK);
            float Di = 0.f;
#pragma unroll
            for (int d = 0; d < DK / 2; ++d) Di += oi[d].x * gi[d].x + oi[d].y * gi[d].y;
            const float li = lse[w * L + i];
#pragma unroll
            for (int d = 0; d < DK / 2; ++d) acc[d] = make_float2(0.f, 0.f);
#pragma unroll 2
            for (int j = 0; j < L; ++j) {
                const float p = __expf(dot2<DK>(qi, Ks + j * DK) * sc - li);
                const float kf = keep_fast(keep_h, p_drop, key, i * L + j) * inv_keep;
                const float ds = p * (kf * dot2<DK>(gi, Vs + j * DK) - Di) * sc;
                Pi[j] = p * kf;
                Si[j] = ds;
                axpy2<DK>(acc, ds, Ks + j * DK);
            }
            store_row2<DK>(dq + (r * L + i) * ld + head * DK, acc, 1.f);
        } else {                        // masked_fill blocks the gradient to q and k; v still sees the uniform weights
            for (int j = 0; j < L; ++j) {
                Pi[j] = invL * keep_fast(keep_h, p_drop, key, i * L + j) * inv_keep;
                Si[j] = 0.f;
            }
#pragma unroll
            for (int d = 0; d < DK / 2; ++d) acc[d] = make_float2(0.f, 0.f);
            store_row2<DK>(dq + (r * L + i) * ld + head * DK, acc, 1.f);
        }
    }
    __syncwarp();
    for (int j = lane; j < L; j += 32) {
        float2 av[DK / 2], ak[DK / 2];
#pragma unroll
        for (int d = 0; d < DK / 2; ++d) av[d] = ak[d] = make_float2(0.f, 0.f);
#pragma unroll 2
        for (int i = 0; i < L; ++i) {
            axpy2<DK>(av, Ps[i * LP + j], Gs + i * DK);
            axpy2<DK>(ak, Ss[i * LP + j], Qs + i * DK);
        }
        store_row2<DK>(dv + (r * L + j) * ld + head * DK, av, 1.f);
        store_row2<DK>(dk_ + (r * L + j) * ld + head * DK, ak, 1.f);
    }
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

// warps per block for the one-warp-per-(title, head) kernels: the count that keeps the most warps resident per SM given the
// per-warp shared-memory slab (ncu on the first version: 4 warps/block with a 31 KB slab left ONE block = 4 warps per SM)
static int best_warps_per_block(size_t per_warp_bytes, int regs_per_thread) {
    const size_t smem_sm = 227 * 1024;
    int best = 1, best_resident = 0;
    for (int wf = 1; wf <= 4; wf *= 2) {
        const size_t block = per_warp_bytes * wf + 1024;                 // + per-block reservation
        if (block > smem_sm) break;
        int blocks = (int)(smem_sm / block);
        const int by_regs = 65536 / (regs_per_thread * 32 * wf);
        if (blocks > by_regs) blocks = by_regs;
        if (blocks > 32) blocks = 32;
        const int resident = blocks * wf;
        if (resident >= best_resident) { best_resident = resident; best = wf; }
    }
    return best;
}

template <int DK>
static int launch_fwd(const float *q, const float *k, const float *v, long long ld, const float *mask, long long R,
                      int L, int h, const float *keep, float p_drop, unsigned long long seed, float *o, float *lse,
                      cudaStream_t st) {
    {   // fast path: two-pass softmax, packed FMAs, 4 warps per block while the per-warp slab (K, V, scores) fits
        const size_t per_warp = (size_t)(2 * L * DK + 32 * L) * sizeof(float);
        if (per_warp <= 110 * 1024) {
            const int wf = best_warps_per_block(per_warp, 96);
            cudaFuncSetAttribute(mha_fwd_fast_kernel<DK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(per_warp * wf));
            mha_fwd_fast_kernel<DK><<<(unsigned)cdiv(R * h, wf), wf * 32, per_warp * wf, st>>>(q, k, v, ld, mask, R, L, h, keep,
                                                                                             p_drop, seed, o, lse);
            XNRS_LAUNCHED();
            return XNRS_OK;
        }
    }
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
    {
        const size_t per_warp = (size_t)((4 * L * DK + 2 * L * (L | 1) + 3) & ~3) * sizeof(float);
        if (per_warp <= 110 * 1024) {
            const int wf = best_warps_per_block(per_warp, 255);
            cudaFuncSetAttribute(mha_bwd_fast_kernel<DK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(per_warp * wf));
            mha_bwd_fast_kernel<DK><<<(unsigned)cdiv(R * h, wf), wf * 32, per_warp * wf, st>>>(
                q, k, v, o, d_o, ld, mask, lse, R, L, h, keep, p_drop, seed, dq, dk_, dv);
            XNRS_LAUNCHED();
            return XNRS_OK;
        }
    }
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
