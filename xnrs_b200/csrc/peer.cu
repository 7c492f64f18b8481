// Row C, multi-GPU: the two exchange steps of the global-batch InfoNCE (training.py:433-472 evaluated over the batches of all
// data-parallel ranks) fused with the computation on either side of them, over NVLink peer memory instead of NCCL.
//
// Every rank owns one buffer of the same layout, mapped into every other rank's address space (CUDA VMM peer mappings set
// up by torch symmetric memory; `ptrs` = the W base addresses as seen from this rank).
//   * normalize + all-gather: a rank L2-normalises ITS Ba embedding rows and stores each finished row (1 KB), its inverse
//     norm and its label straight into all W gathered buffers (P2P stores) — no staging copy, no collective launch, and
//     no rank normalises another rank's rows;
//   * reduce-scatter + normalisation backward: every rank has the InfoNCE gradient w.r.t. ALL Bk normalised rows in its
//     buffer; a rank pulls the W partial gradients of ITS rows (P2P loads, fixed order: deterministic), sums them and
//     applies the normalisation backward and the loss scale in the same pass.
// Cross-GPU ordering is a flag per (kernel kind, source rank) in the receiver's buffer: the producer makes its data visible
// system-wide (__threadfence_system), then stores the launch's epoch into its slot with release semantics; the consumer
// spins on its own slots with acquire loads.  Epochs are kept in device memory and advanced by the kernels themselves, so a
// replayed CUDA graph works unchanged.  A buffer is reused one training step later; the step's gradient all-reduce lies in
// between, which no rank passes before every rank has finished reading (see xnrs_b200/distributed.py).
// A spin that sees no progress for about a minute gives up and raises ctl[4] (checked on the host) instead of hanging the GPU.
#include "common.cuh"

namespace xnrs {

__device__ __forceinline__ void st_release_sys(unsigned *p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned *p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// wait until every rank's slot has reached `epoch` (slots written by the peers into THIS rank's buffer)
__device__ __forceinline__ void wait_flags(const unsigned *mine, int world, unsigned epoch, int *ctl) {
    for (int q = 0; q < world; ++q) {
        long long spins = 0;
        while ((int)(ld_acquire_sys(mine + q) - epoch) < 0) {
            if (++spins > (1LL << 26)) {            // about a minute: a peer never arrived (a rank that merely stalls — graph
                                                    // capture, a checkpoint write — is waited for, as a NCCL kernel would)
                atomicExch(ctl + 4, 1);
                return;
            }
            __nanosleep(200);
        }
    }
}

// ctl: [0] epoch of the all-gather kind, [1] its CTA ticket, [2] epoch of the reduce-scatter kind, [3] its ticket, [4] error
__global__ void __launch_bounds__(256)
peer_normalize_allgather_kernel(const float *__restrict__ emb, const int *__restrict__ labels, long long Ba, int E, int rank,
                                int world, const long long *__restrict__ ptrs, long long off_flags, long long off_ehat,
                                long long off_inv, long long off_lab, int *__restrict__ ctl) {
    const int lane = threadIdx.x & 31;
    const int E4 = E >> 2;
    long long w = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const long long nw = ((long long)gridDim.x * blockDim.x) >> 5;
    for (; w < Ba; w += nw) {
        const float4 *src = reinterpret_cast<const float4 *>(emb + w * E);
        float4 v[2];                                    // E <= 256: two 16-byte chunks per lane
        float ss = 0.f;
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const int c = lane + 32 * k;
            v[k] = c < E4 ? src[c] : make_float4(0.f, 0.f, 0.f, 0.f);
            ss = fmaf(v[k].x, v[k].x, ss); ss = fmaf(v[k].y, v[k].y, ss);
            ss = fmaf(v[k].z, v[k].z, ss); ss = fmaf(v[k].w, v[k].w, ss);
        }
        ss = warp_sum(ss);
        const float inv = 1.f / fmaxf(sqrtf(ss), 1e-12f);     // F.normalize(eps=1e-12), as xnrs_infonce_normalize
#pragma unroll
        for (int k = 0; k < 2; ++k) { v[k].x *= inv; v[k].y *= inv; v[k].z *= inv; v[k].w *= inv; }
        const long long g = (long long)rank * Ba + w;   // row in the gathered batch
        const int lab = labels[w];
        for (int p = 0; p < world; ++p) {
            char *base = reinterpret_cast<char *>(ptrs[p]);
            float4 *dst = reinterpret_cast<float4 *>(base + off_ehat) + g * E4;
#pragma unroll
            for (int k = 0; k < 2; ++k)
                if (lane + 32 * k < E4) dst[lane + 32 * k] = v[k];
            if (lane == 0) {
                reinterpret_cast<float *>(base + off_inv)[g] = inv;
                reinterpret_cast<int *>(base + off_lab)[g] = lab;
            }
        }
    }
    __threadfence_system();                             // this thread's peer stores are visible before its CTA's ticket
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned ticket = atomicAdd(reinterpret_cast<unsigned *>(ctl + 1), 1u);
        if (ticket == gridDim.x - 1) {                  // every CTA of this rank has stored (and fenced) its rows
            __threadfence_system();
            const unsigned epoch = (unsigned)ctl[0] + 1u;
            for (int p = 0; p < world; ++p)
                st_release_sys(reinterpret_cast<unsigned *>(reinterpret_cast<char *>(ptrs[p]) + off_flags) + rank, epoch);
            wait_flags(reinterpret_cast<const unsigned *>(reinterpret_cast<const char *>(ptrs[rank]) + off_flags), world, epoch, ctl);
            ctl[0] = (int)epoch;
            ctl[1] = 0;
        }
    }
}

// d_emb[i] = sc * inv_norm[i] * (d - ehat[i] <ehat[i], d>),  d = sum over ranks p of d_ehat_p[row0 + i],
// sc = gscale * (*gptr) / (stats[1] + 1e-8)       (as xnrs_infonce_normalize_bwd followed by the loss-weight scaling)
__global__ void __launch_bounds__(256)
peer_reduce_scatter_bwd_kernel(const long long *__restrict__ ptrs, long long off_flags, long long off_dehat, long long Ba, int E,
                               int rank, int world, const float *__restrict__ ehat_a, const float *__restrict__ inv_norm_a,
                               const float *__restrict__ stats, float gscale, const float *__restrict__ gptr,
                               float *__restrict__ d_emb, int *__restrict__ ctl) {
    const int lane = threadIdx.x & 31;
    const int E4 = E >> 2;
    const unsigned epoch = (unsigned)ctl[2] + 1u;       // advanced by the last CTA to FINISH: every CTA reads the old value
    if (threadIdx.x == 0) {
        if (blockIdx.x == 0) {                          // my gradient block was completed by the kernels before this one
            __threadfence_system();
            for (int p = 0; p < world; ++p)
                st_release_sys(reinterpret_cast<unsigned *>(reinterpret_cast<char *>(ptrs[p]) + off_flags) + rank, epoch);
        }
        wait_flags(reinterpret_cast<const unsigned *>(reinterpret_cast<const char *>(ptrs[rank]) + off_flags), world, epoch, ctl);
    }
    __syncthreads();
    const float sc = gscale * (gptr ? gptr[0] : 1.f) / (stats[1] + 1e-8f);
    const long long row0 = (long long)rank * Ba;
    long long w = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const long long nw = ((long long)gridDim.x * blockDim.x) >> 5;
    for (; w < Ba; w += nw) {
        float4 d[2] = {make_float4(0.f, 0.f, 0.f, 0.f), make_float4(0.f, 0.f, 0.f, 0.f)};
        for (int p = 0; p < world; ++p) {               // fixed order: the sum is deterministic
            const float4 *src = reinterpret_cast<const float4 *>(reinterpret_cast<const char *>(ptrs[p]) + off_dehat) + (row0 + w) * E4;
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                if (lane + 32 * k < E4) {
                    const float4 t = src[lane + 32 * k];
                    d[k].x += t.x; d[k].y += t.y; d[k].z += t.z; d[k].w += t.w;
                }
            }
        }
        const float4 *eh = reinterpret_cast<const float4 *>(ehat_a + w * E);
        float4 e[2];
        float dot = 0.f;
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            e[k] = lane + 32 * k < E4 ? eh[lane + 32 * k] : make_float4(0.f, 0.f, 0.f, 0.f);
            dot = fmaf(e[k].x, d[k].x, dot); dot = fmaf(e[k].y, d[k].y, dot);
            dot = fmaf(e[k].z, d[k].z, dot); dot = fmaf(e[k].w, d[k].w, dot);
        }
        dot = warp_sum(dot);
        const float s = sc * inv_norm_a[w];
        float4 *dst = reinterpret_cast<float4 *>(d_emb + w * E);
#pragma unroll
        for (int k = 0; k < 2; ++k)
            if (lane + 32 * k < E4)
                dst[lane + 32 * k] = make_float4(s * (d[k].x - e[k].x * dot), s * (d[k].y - e[k].y * dot),
                                                 s * (d[k].z - e[k].z * dot), s * (d[k].w - e[k].w * dot));
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned ticket = atomicAdd(reinterpret_cast<unsigned *>(ctl + 3), 1u);
        if (ticket == gridDim.x - 1) {
            ctl[2] = (int)epoch;
            ctl[3] = 0;
        }
    }
}

static unsigned peer_grid(long long rows) {
    long long b = cdiv(rows, 8), cap = num_sms();       // a warp per row, 8 warps per CTA, one wave (all CTAs spin-safe)
    return (unsigned)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace xnrs

using namespace xnrs;

extern "C" int xnrs_peer_normalize_allgather(const float *emb, const int *labels, long long Ba, int E, int rank, int world,
                                             const long long *ptrs, long long off_flags, long long off_ehat, long long off_inv,
                                             long long off_lab, int *ctl, xnrs_stream_t st) {
    XNRS_REQUIRE(Ba > 0 && E > 0 && E % 4 == 0 && E <= 256 && world > 0 && world <= 64 && rank >= 0 && rank < world, "bad sizes");
    XNRS_REQUIRE(emb && labels && ptrs && ctl && !((uintptr_t)emb & 15), "null / unaligned pointer");
    XNRS_REQUIRE(off_flags % 4 == 0 && off_ehat % 16 == 0 && off_inv % 4 == 0 && off_lab % 4 == 0, "unaligned buffer layout");
    peer_normalize_allgather_kernel<<<peer_grid(Ba), 256, 0, STREAM(st)>>>(emb, labels, Ba, E, rank, world, ptrs, off_flags,
                                                                          off_ehat, off_inv, off_lab, ctl);
    XNRS_LAUNCHED();
    return XNRS_OK;
}

extern "C" int xnrs_peer_reduce_scatter_normalize_bwd(const long long *ptrs, long long off_flags, long long off_dehat, long long Ba,
                                                      int E, int rank, int world, const float *ehat_a, const float *inv_norm_a,
                                                      const float *stats, float grad_scale, const float *grad_scale_dev,
                                                      float *d_emb, int *ctl, xnrs_stream_t st) {
    XNRS_REQUIRE(Ba > 0 && E > 0 && E % 4 == 0 && E <= 256 && world > 0 && world <= 64 && rank >= 0 && rank < world, "bad sizes");
    XNRS_REQUIRE(ptrs && ehat_a && inv_norm_a && stats && d_emb && ctl, "null pointer");
    XNRS_REQUIRE(off_flags % 4 == 0 && off_dehat % 16 == 0 && !((uintptr_t)ehat_a & 15) && !((uintptr_t)d_emb & 15), "unaligned");
    peer_reduce_scatter_bwd_kernel<<<peer_grid(Ba), 256, 0, STREAM(st)>>>(ptrs, off_flags, off_dehat, Ba, E, rank, world, ehat_a,
                                                                         inv_norm_a, stats, grad_scale, grad_scale_dev, d_emb, ctl);
    XNRS_LAUNCHED();
    return XNRS_OK;
}
