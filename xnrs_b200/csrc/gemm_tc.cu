// tcgen05 (5th-gen tensor core) GEMM for sm_100a: TMA-fed, TMEM accumulators, warp-specialised, persistent.
//
//   C[M,N] (=|+=) act( opA(A)[M,K] * opB(B)[K,N] + bias )     A, B, C fp32 in global memory
//
// Arithmetic modes (include/xnrs_b200.h XNRS_PREC_*):
//   TF32X3  fp32-accurate: every fp32 operand tile is split IN SHARED MEMORY into hi = rna_tf32(x) and
//           lo = x - hi by dedicated splitter warps, and the issuer runs 3 MMAs per k-step
//           (hi*hi + lo*hi + hi*lo; the dropped lo*lo term is ~2^-22 relative).  L2 traffic stays 1x.
//   TF32    single pass on the raw fp32 tiles (tensor core reads the top 19 bits).
//
// CTA = 4 warpgroups: {warp 0 TMA producer, warp 1 MMA issuer (+TMEM alloc)} | warps 4-11 splitters | warps 12-15
// epilogue (a full 128-column row per thread).  The splitters were the critical path of the 3-stage pipeline with 4 warps
// (ncu source page: the MMA issuer spun on split_bar, the producer on empty_bar): 8 warps, shared-space 128-bit
// loads/stores with all loads of a stage in flight, and one mbarrier arrival per warp.
// Tile 128 x 128 x 32 (UMMA M=128, N=128, K=8 per instruction, 128-byte swizzled rows), 3 (x3) or 6 (x1)
// smem stages, 2 TMEM accumulator stages (x2 accumulators: main + correction = all 512 columns) so the epilogue of
// tile i overlaps the main loop of tile i+1.
// Both operand majors are supported through the UMMA descriptors (K-major: nn.Linear forward;
// MN-major: weight-gradient and input-gradient GEMMs), so no transposed copies are ever made.
#include <cuda.h>
#include <cuda_bf16.h>
#include <algorithm>
#include <stdlib.h>
#include <string.h>

#include "gemm.cuh"

namespace xnrs {

constexpr int TBM = 128, TBK = 32;
constexpr int SMEM_DATA = 192 * 1024;
constexpr int SMEM_EPI = 8 * 4096;       // per-epilogue-warp transpose staging (coalesced stores)
constexpr int TC_THREADS = 512;          // 4 warpgroups: {TMA, MMA, -, -} | 8 splitter warps | 4 epilogue warps
constexpr int SPLIT_WARPS = 8, SPLIT_THREADS = SPLIT_WARPS * 32, EPI_WARP0 = 4 + SPLIT_WARPS;
constexpr int MAX_STAGES = 6;
constexpr long long KCHUNK = 2048;   // longest K run accumulated inside the tensor core: its fp32 accumulation truncates, so the
                                     // error grows ~linearly with K; longer reductions are split and summed with IEEE fp32 atomics

// fused additive-attention pooling epilogue of the CTA-pair kernel (xnrs_titlepool_fwd): with hid = tanh(x W1^T + b1) in
// TMEM, logit = <hid, w2> + b2, e = exp(logit), and per title sum(e) and sum(e * x) are accumulated — layers.py:60-65
struct PoolArgs {
    const float *w2, *b2;   // fc2 weight (N = 256 hidden units) and bias
    const int *tix;         // title of each row (M entries), -1 for padding rows
    float *e;               // (M) un-normalised pooling weights exp(logit)
    float *zsum;            // (R) += sum of e over the title's rows
    float *pooled;          // (R, F) += sum of e * x row
    int F4;                 // x row width in float4 (= K / 4)
};

struct TcArgs {
    long long M, N, K;
    float *C; long long ldc;
    const float *bias; int act; const float *aux; int accumulate;
    int split_k; long long k_per_split;
    int a_mn, b_mn;         // 1: operand stored [K, MN] (MN contiguous); 0: stored [MN, K] (K contiguous)
    const int *a_gather;    // K-major A: stored rows gathered through this index (table gather fused by TMA gather4)
    const int *b_gather;    // MN-major B: stored rows (= K index) gathered through this index
    int pf_stages;          // gathered weight-gradient form: L2-prefetch the table rows this many stages ahead (0: off)
    int lsu_gather;         // CTA-pair kernel: 1 = A rows, 2 = B rows gathered by a cp.async producer warp instead of TMA gather4
    const float *A; long long lda;      // raw operand pointers for that warp
    const float *B; long long ldb;
    int passes;             // 3 (TF32X3) or 1 (TF32 / BF16)
    int elt;                // bytes per operand element: 4 (fp32 storage, kind::tf32) or 2 (bf16 storage, kind::f16) — pair kernel
    int c_bf16;             // pair kernel: C is stored as bf16 (fp32 accumulators rounded to nearest even)
    PoolArgs pool;
    int stages;
    long long tiles_m, tiles_n;
    int planes;             // pair kernel, elt == 2: 2 = both operands arrive PRE-SPLIT as two bf16 planes (x ~ hi + lo, A2 / B2 and
                            // mapA2 / mapB2 are the lo planes): three kind::f16 MMAs per k-step like 3xTF32, no splitter warps
    const float *A2, *B2;   // raw lo-plane pointers for the cp.async gather warps
    int a_f16, b_f16;       // elt == 2: the operand's 16-bit planes are IEEE fp16 (11-bit mantissa) instead of bf16
    int c_tma;              // 1-CTA kernel: C leaves through TMA stores / reductions of staged 32 x 32 blocks (mapC is valid)
    long long *trace;       // diagnostics (xnrs_debug_gemm_trace): 8 SM-clock stamps per CTA of the 1-CTA kernel, or NULL
};

// ---- PTX wrappers --------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// bounded wait: a protocol bug traps (CUDA error) after ~2 s instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    if (mbar_try(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try(bar, parity)) {
        if (clock64() - t0 > 4000000000LL) __trap();
    }
}
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
// TMA store of one staged 32 x 32 fp32 block (128-byte rows, 128-byte swizzle) to C; rows / columns past the tensor's extent
// are clipped by the unit.  The reduce form adds into C at the L2 (accumulate, split-K).  Bulk-group completion.
__device__ __forceinline__ void tma_store_2d(const CUtensorMap *map, uint32_t src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(map), "r"(src), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap *map, uint32_t src, int c0, int c1) {
    asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(map), "r"(src), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }
// Blackwell TMA gather: four independent rows (same column window) land as four consecutive smem rows
__device__ __forceinline__ void tma_gather4(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int4 rows) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes "
        "[%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(rows.x), "r"(rows.y), "r"(rows.z), "r"(rows.w)
        : "memory");
}
// Ampere-style asynchronous 16-byte copy global -> shared (LDGSTS); src_bytes < 16 zero-fills the remainder
__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src, int src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
// the mbarrier receives one (pre-counted) arrival once all cp.async copies issued so far by this thread have landed
__device__ __forceinline__ void cp_async_arrive_noinc(uint64_t *bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t *r = reinterpret_cast<uint32_t *>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ float4 lds128(uint32_t saddr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr));
    return v;
}
__device__ __forceinline__ void sts128(uint32_t saddr, float4 v) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float tf32_lo(float x) { return x - __uint_as_float(__float_as_uint(x) & 0xffffe000u); }

// UMMA shared-memory descriptor, 128-byte swizzle (cute::UMMA::SmemDescriptor: start>>4 [0,14), LBO>>4 [16,30),
// SBO>>4 [32,46), version=1 [46,48), layout SWIZZLE_128B=2 [61,64))
// MN-major 32-bit (tf32) operands must use the 32-byte-atom variant SWIZZLE_128B_BASE32B=1 (Swizzle<2,5,2>, 4 k-rows
// per 512-byte atom) — "for mn-major tf32 operands, SW128_32B is the only available smem layout" (CUTLASS sm100 builder).
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46) | ((uint64_t)layout << 61);
}


// tanh for the tensor-core epilogues: |x| < 0.25 -> odd Taylor polynomial (abs err < 3e-9), otherwise
// 1 - 2 / (1 + 2^(2 x log2 e)) with ex2.approx / rcp.approx (abs err ~2e-7, saturates to +-1).  libm tanhf costs ~60
// dependent instructions per element, which made the epilogue of the fc1 GEMM (128 elements per thread per tile) longer
// than its main loop; this is 14 branch-free instructions and stays far inside the 1e-4 parity bar.
__device__ __forceinline__ float tanh_fast(float x) {
    const float x2 = x * x;
    float pl = fmaf(x2, 62.f / 2835.f, -17.f / 315.f);
    pl = fmaf(pl, x2, 2.f / 15.f);
    pl = fmaf(pl, x2, -1.f / 3.f);
    pl = fmaf(pl * x2, x, x);
    float t, rc;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(x * 2.885390081777927f));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rc) : "f"(t + 1.f));
    return fabsf(x) < 0.25f ? pl : fmaf(-2.f, rc, 1.f);
}

// Coalesced epilogue.  tcgen05.ld hands lane l row l of a 32-row x 32-column accumulator block; storing that directly
// makes every warp store touch 32 different 128-byte lines (16 bytes each), and the LSU then needs ~32 cycles per
// instruction: measured, the epilogue (not the tensor pipe) bounded every GEMM with K <= 768.  So the raw accumulators are
// transposed through a 4 KB per-warp staging buffer (16-byte chunks XOR-swizzled by row: conflict-free both ways) and all
// epilogue math + global traffic happens in the transposed layout: lane = (row 4i + l/8, columns 4(l%8)..+3), i.e.
// 4 rows x 128 contiguous bytes per instruction for C, bias, the ReLU mask and the accumulate read.
template <int ACT>
__device__ __forceinline__ void epi_rows(uint32_t stage, const TcArgs &p, const float *bias, long long row0, long long col0,
                                         long long split, int lane, bool vec_ok, bool bias_vec, bool has_pre, float4 bias_pre) {
    const int sub = lane >> 3, ch = lane & 7;
    const long long col = col0 + 4 * ch;
    if (col >= p.N) return;
    const int nv = (int)min((long long)4, p.N - col);
    float b[4] = {0.f, 0.f, 0.f, 0.f};
    if (has_pre) {              // this lane's four bias values, loaded by the caller before it waited for the accumulator
        b[0] = bias_pre.x; b[1] = bias_pre.y; b[2] = bias_pre.z; b[3] = bias_pre.w;
    } else if (bias && (p.split_k == 1 || split == 0)) {
        if (bias_vec && nv == 4) {
            const float4 t = __ldg(reinterpret_cast<const float4 *>(bias + col));
            b[0] = t.x; b[1] = t.y; b[2] = t.z; b[3] = t.w;
        } else {
#pragma unroll
            for (int e = 0; e < 4; ++e)
                if (e < nv) b[e] = __ldg(bias + col + e);
        }
    }
    // Fast path (interior columns, fp32 C, no split-K): straight-line code with the 8 row reads, the optional accumulate /
    // mask reads and the 8 row writes each issued as a batch.  One warp per SM sub-partition runs this, so its cost is the
    // LENGTH of the dependent instruction chain: the general loop below (64-bit index math and mode checks per row) took
    // 1.2-1.4 us per 32 x 32 block, 5 us per 128 x 128 tile — a third of a small GEMM's whole duration (SM-clock stamps).
    if (vec_ok && nv == 4 && !p.c_bf16) {
        const long long rows_left = p.M - row0;                 // > 0: the caller skips blocks past M
        const long long step = 4 * p.ldc;
        float *dst = p.C + (row0 + sub) * p.ldc + col;
        const uint32_t s0 = stage + sub * 128 + ((ch ^ sub) << 4), s1 = stage + (sub + 4) * 128 + ((ch ^ (sub + 4)) << 4);
        float4 v[8], o[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = lds128(((i & 1) ? s1 : s0) + (i >> 1) * 1024);
        if (p.split_k > 1) {        // partial sums: 16-byte vector reductions into C (act is NONE, bias rides on split 0)
#pragma unroll
            for (int i = 0; i < 8; ++i)
                if (4 * i + sub < rows_left)
                    atomicAdd(reinterpret_cast<float4 *>(dst + i * step), make_float4(v[i].x + b[0], v[i].y + b[1], v[i].z + b[2], v[i].w + b[3]));
            return;
        }
        if (ACT == XNRS_ACT_RELU_MASK) {
            const float *ax = p.aux + (row0 + sub) * p.ldc + col;
#pragma unroll
            for (int i = 0; i < 8; ++i)
                o[i] = 4 * i + sub < rows_left ? *reinterpret_cast<const float4 *>(ax + i * step) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                v[i].x = o[i].x > 0.f ? v[i].x + b[0] : 0.f; v[i].y = o[i].y > 0.f ? v[i].y + b[1] : 0.f;
                v[i].z = o[i].z > 0.f ? v[i].z + b[2] : 0.f; v[i].w = o[i].w > 0.f ? v[i].w + b[3] : 0.f;
            }
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                v[i].x += b[0]; v[i].y += b[1]; v[i].z += b[2]; v[i].w += b[3];
                if (ACT == XNRS_ACT_RELU) {
                    v[i].x = fmaxf(v[i].x, 0.f); v[i].y = fmaxf(v[i].y, 0.f); v[i].z = fmaxf(v[i].z, 0.f); v[i].w = fmaxf(v[i].w, 0.f);
                } else if (ACT == XNRS_ACT_TANH) {
                    v[i].x = tanh_fast(v[i].x); v[i].y = tanh_fast(v[i].y); v[i].z = tanh_fast(v[i].z); v[i].w = tanh_fast(v[i].w);
                }
            }
        }
        if (p.accumulate) {     // C += v as 16-byte reductions: one fp32 add per element like load-add-store (only this CTA
#pragma unroll                 // touches the element), without the read's round trip in the warp's dependent chain
            for (int i = 0; i < 8; ++i)
                if (4 * i + sub < rows_left) atomicAdd(reinterpret_cast<float4 *>(dst + i * step), v[i]);
            return;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i)
            if (4 * i + sub < rows_left) *reinterpret_cast<float4 *>(dst + i * step) = v[i];
        return;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int rr = 4 * i + sub;
        const long long row = row0 + rr;
        const float4 v = lds128(stage + rr * 128 + ((ch ^ (rr & 7)) << 4));
        if (row >= p.M) continue;
        float x[4] = {v.x + b[0], v.y + b[1], v.z + b[2], v.w + b[3]};
        float *dst = p.C + row * p.ldc + col;
        if (p.split_k > 1) {
#pragma unroll
            for (int e = 0; e < 4; ++e)
                if (e < nv) atomicAdd(dst + e, x[e]);
            continue;
        }
        if (ACT == XNRS_ACT_RELU) {
#pragma unroll
            for (int e = 0; e < 4; ++e) x[e] = fmaxf(x[e], 0.f);
        } else if (ACT == XNRS_ACT_TANH) {
#pragma unroll
            for (int e = 0; e < 4; ++e) x[e] = tanh_fast(x[e]);
        }
        if (p.c_bf16) {         // bf16 C (no accumulate / split-K / ReLU mask: checked on the host), 8-byte stores
            __nv_bfloat16 *d16 = reinterpret_cast<__nv_bfloat16 *>(p.C) + row * p.ldc + col;
            if (nv == 4) {
                const __nv_bfloat162 lo = __floats2bfloat162_rn(x[0], x[1]), hi = __floats2bfloat162_rn(x[2], x[3]);
                uint2 pk;
                pk.x = *reinterpret_cast<const uint32_t *>(&lo);
                pk.y = *reinterpret_cast<const uint32_t *>(&hi);
                *reinterpret_cast<uint2 *>(d16) = pk;
            } else {
#pragma unroll
                for (int e = 0; e < 4; ++e)
                    if (e < nv) d16[e] = __float2bfloat16_rn(x[e]);
            }
            continue;
        }
        if (vec_ok && nv == 4) {
            if (ACT == XNRS_ACT_RELU_MASK) {
                const float4 a = *reinterpret_cast<const float4 *>(p.aux + row * p.ldc + col);
                x[0] = a.x > 0.f ? x[0] : 0.f; x[1] = a.y > 0.f ? x[1] : 0.f;
                x[2] = a.z > 0.f ? x[2] : 0.f; x[3] = a.w > 0.f ? x[3] : 0.f;
            }
            if (p.accumulate) {
                const float4 o = *reinterpret_cast<const float4 *>(dst);
                x[0] += o.x; x[1] += o.y; x[2] += o.z; x[3] += o.w;
            }
            *reinterpret_cast<float4 *>(dst) = make_float4(x[0], x[1], x[2], x[3]);
        } else {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                if (e < nv) {
                    if (ACT == XNRS_ACT_RELU_MASK) x[e] = (p.aux[row * p.ldc + col + e] > 0.f) ? x[e] : 0.f;
                    dst[e] = p.accumulate ? dst[e] + x[e] : x[e];
                }
            }
        }
    }
}

// r: lane l holds the FINISHED values of columns [col, col+32) of row `row + l`: stage them as 32 rows of 128 bytes in the
// 128-byte swizzle and let one thread hand the 4 KB block to the TMA unit (store, or add-reduction for accumulate /
// split-K).  `pending`: bulk groups of this thread that may still be reading OTHER staging buffers (0: `buf` is the only one).
template <int PENDING>
__device__ __forceinline__ void tma_store_block32(const float (&r)[32], uint32_t buf, const CUtensorMap *mapC, long long col,
                                                  long long row, bool reduce, int lane) {
    if (lane == 0) bulk_wait_read<PENDING>();       // the earlier store from `buf` has read it
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 8; ++j)
        sts128(buf + lane * 128 + ((j ^ (lane & 7)) << 4), make_float4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]));
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (lane == 0) {
        if (reduce) tma_reduce_add_2d(mapC, buf, (int)col, (int)row);
        else tma_store_2d(mapC, buf, (int)col, (int)row);
        bulk_commit();
    }
}

// bias (16-byte aligned, N % 4 == 0) + activation on a row-per-lane block
__device__ __forceinline__ void bias_act_block32(float (&r)[32], const float *bias, long long col0, long long N, int act,
                                                 const float *aux = nullptr, long long ldc = 0, long long row = 0, long long M = 0) {
    if (bias) {
        const float4 *b4 = reinterpret_cast<const float4 *>(bias + col0);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (col0 + 4 * j < N) {
                const float4 b = __ldg(b4 + j);         // same address in every lane: one L1 transaction
                r[4 * j] += b.x; r[4 * j + 1] += b.y; r[4 * j + 2] += b.z; r[4 * j + 3] += b.w;
            }
        }
    }
    if (act == XNRS_ACT_RELU) {
#pragma unroll
        for (int j = 0; j < 32; ++j) r[j] = fmaxf(r[j], 0.f);
    } else if (act == XNRS_ACT_TANH) {
#pragma unroll
        for (int j = 0; j < 32; ++j) r[j] = tanh_fast(r[j]);
    } else if (act == XNRS_ACT_RELU_MASK) {     // ReLU backward: this lane's row of the saved activation (128 contiguous bytes)
        if (row < M) {
            const float4 *a4 = reinterpret_cast<const float4 *>(aux + row * ldc + col0);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                if (col0 + 4 * j < N) {
                    const float4 a = __ldg(a4 + j);
                    r[4 * j] = a.x > 0.f ? r[4 * j] : 0.f; r[4 * j + 1] = a.y > 0.f ? r[4 * j + 1] : 0.f;
                    r[4 * j + 2] = a.z > 0.f ? r[4 * j + 2] : 0.f; r[4 * j + 3] = a.w > 0.f ? r[4 * j + 3] : 0.f;
                }
            }
        }
    }
}

// r: lane l holds columns [col0, col0+32) of row row0 + l (raw accumulators); stage: this warp's 4 KB staging buffer
__device__ __forceinline__ void epi_block32(float (&r)[32], uint32_t stage, const TcArgs &p, long long row0, long long col0,
                                            long long split, int lane, bool vec_ok, bool bias_vec, int act, const float *bias,
                                            bool has_pre = false, float4 bias_pre = float4{0.f, 0.f, 0.f, 0.f}) {
#pragma unroll
    for (int c = 0; c < 8; ++c)
        sts128(stage + lane * 128 + ((c ^ (lane & 7)) << 4), make_float4(r[4 * c], r[4 * c + 1], r[4 * c + 2], r[4 * c + 3]));
    __syncwarp();
    switch (act) {
        case XNRS_ACT_RELU: epi_rows<XNRS_ACT_RELU>(stage, p, bias, row0, col0, split, lane, vec_ok, bias_vec, has_pre, bias_pre); break;
        case XNRS_ACT_TANH: epi_rows<XNRS_ACT_TANH>(stage, p, bias, row0, col0, split, lane, vec_ok, bias_vec, has_pre, bias_pre); break;
        case XNRS_ACT_RELU_MASK: epi_rows<XNRS_ACT_RELU_MASK>(stage, p, bias, row0, col0, split, lane, vec_ok, bias_vec, has_pre, bias_pre); break;
        default: epi_rows<XNRS_ACT_NONE>(stage, p, bias, row0, col0, split, lane, vec_ok, bias_vec, has_pre, bias_pre); break;
    }
    __syncwarp();
}

struct StageRing {
    int stage = 0;
    uint32_t phase = 0;
    __device__ __forceinline__ void advance(int n) {
        if (++stage == n) { stage = 0; phase ^= 1; }
    }
};

// BN = 128 or 256 output columns per tile.  With SS-mode MMAs the operand fetch (A 4 KB + B BN*32 B per K=8 step) runs at
// the full 128 B/clk shared-memory bandwidth for BN=128; BN=256 halves the A bytes per FLOP (ncu: L1/shared was the bound).
template <int BN>
__global__ void __launch_bounds__(TC_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
               const __grid_constant__ CUtensorMap mapC, TcArgs p) {
    constexpr int TILE_A = TBM * TBK * 4, TILE_B = BN * TBK * 4, HALF = TILE_A + TILE_B;   // [A hi][B hi] | [A lo][B lo]
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    __shared__ __align__(8) uint64_t full_bar[MAX_STAGES], split_bar[MAX_STAGES], empty_bar[MAX_STAGES];
    __shared__ __align__(8) uint64_t tfull_bar[2], tempty_bar[2];
    __shared__ uint32_t tmem_base_smem;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    auto stamp = [&](int i) { if (p.trace) p.trace[blockIdx.x * 16 + i] = clock64(); };
    if (threadIdx.x == 0) {
        stamp(0);
        // descriptor fetch overlaps the barrier / TMEM set-up instead of delaying the first load
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&mapA)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&mapB)) : "memory");
    }
    const int stages = p.stages, passes = p.passes;
    const int stage_bytes = (passes == 3 ? 2 : 1) * HALF;
    const int acc_cols = (passes == 3 ? 2 : 1) * BN;         // main (+ correction) accumulator columns per tile
    const int acc_stages = 512 / acc_cols >= 2 ? 2 : 1;      // TMEM has 512 columns
    auto tileA = [&](int s) { return smem + (size_t)s * stage_bytes; };
    auto tileB = [&](int s) { return smem + (size_t)s * stage_bytes + TILE_A; };

    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&split_bar[s], SPLIT_WARPS);
            mbar_init(&empty_bar[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tfull_bar[a], 1);
            mbar_init(&tempty_bar[a], 128);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {        // all 512 TMEM columns (1 CTA per SM)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_smem;
    pdl_launch_dependents();            // all CTAs of this grid are resident: the next kernel may set itself up meanwhile
    pdl_wait();                         // everything below reads / writes global memory
    if (threadIdx.x == 0) stamp(1);

    const long long tiles_mn = p.tiles_m * p.tiles_n;
    const long long total = tiles_mn * p.split_k;

    // One epilogue warp's share of a tile: the BN / 32 chunks (32 rows x 32 columns each) of TMEM lane quarter warp % 4.
    // SM-clock stamps of the coalesced-store form (transpose through shared memory, then per-row index math, loads and
    // 16-byte stores by every lane): 500-600 dependent instructions = 0.7-1.3 us per chunk, 3-5 us per 128 x 128 tile, a
    // fifth to a third of a small GEMM's duration; handing chunks to other warps of the quarter did not help (same SM
    // sub-partition).  So the common case hands the staged block to the TMA unit instead: bias + activation are applied in
    // the row-per-lane registers, the block is staged in the layout a 128-byte-swizzled tensor map describes, and ONE
    // thread issues a 4 KB TMA store (or add-reduction: accumulate / split-K); edge clipping, addressing and the stores
    // themselves cost the warp nothing.  Two staging buffers per warp keep a store in flight while the next is staged.
    constexpr int NCH = BN / 32;
    int cbuf = 0;
    auto epilogue_tile = [&](long long t, int a_stage, uint32_t a_phase, uint32_t stage_addr) {
        const int q = warp & 3;                         // TMEM lane quarter this warp may access
        const bool vec_ok = (p.ldc % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0) &&
                            (!p.aux || (reinterpret_cast<uintptr_t>(p.aux) & 15) == 0);
        const bool bias_vec = (reinterpret_cast<uintptr_t>(p.bias) & 15) == 0;
        const long long split = t / tiles_mn, mn = t - split * tiles_mn;
        const long long m0 = (mn / p.tiles_n) * TBM, n0 = (mn % p.tiles_n) * BN;
        const bool bias_on = p.bias && (p.split_k == 1 || split == 0);
        if (p.c_tma && bias_on && lane < NCH && n0 + 32 * lane < p.N)          // the tile's bias values: into L1 before the wait
            asm volatile("prefetch.global.L1 [%0];" ::"l"(p.bias + n0 + 32 * lane));
        mbar_wait(&tfull_bar[a_stage], a_phase);
        tc_fence_after();
        if (threadIdx.x == EPI_WARP0 * 32 && t == blockIdx.x) stamp(4);
#pragma unroll 1
        for (int c = 0; c < NCH; ++c) {
            float r[32];
            const uint32_t taddr = tmem_base + a_stage * acc_cols + c * 32 + ((uint32_t)(32 * q) << 16);
            tc_ld32(taddr, r);
            if (passes == 3) {
                float corr[32];
                tc_ld32(taddr + BN, corr);
#pragma unroll
                for (int j = 0; j < 32; ++j) r[j] += corr[j];
            }
            if (m0 + 32 * q >= p.M || n0 + c * 32 >= p.N) continue;        // warp-uniform
            if (p.c_tma) {
                // (N % 4 == 0 and a 16-byte aligned bias are conditions of c_tma)
                bias_act_block32(r, bias_on ? p.bias : nullptr, n0 + c * 32, p.N, p.act, p.aux, p.ldc, m0 + 32 * q + lane, p.M);
                tma_store_block32<1>(r, stage_addr + (cbuf << 12), &mapC, n0 + c * 32, m0 + 32 * q, p.accumulate || p.split_k > 1, lane);
                cbuf ^= 1;
            } else {
                epi_block32(r, stage_addr, p, m0 + 32 * q, n0 + c * 32, split, lane, vec_ok, bias_vec, p.act, p.bias);
            }
            if (threadIdx.x == EPI_WARP0 * 32 && t == blockIdx.x && c < 3) stamp(10 + c);
        }
    };

    if (warp == 0) {
        // ===================== TMA producer (lane 0 drives; all 32 lanes issue the gather4 rows) =====================
        StageRing r;
        for (long long t = blockIdx.x; t < total; t += gridDim.x) {
            const long long split = t / tiles_mn, mn = t - split * tiles_mn;
            const int m0 = (int)((mn / p.tiles_n) * TBM), n0 = (int)((mn % p.tiles_n) * BN);
            const long long kbeg = split * p.k_per_split, kend = min(p.K, kbeg + p.k_per_split);
            int4 arow = make_int4(0, 0, 0, 0);
            if (p.a_gather) {       // rows m0+4*lane .. +3 of this tile, fixed for the whole K loop; rows past M re-read
                const long long m = m0 + 4 * lane, last = p.M - 1;      // the last valid row (their results are dropped)
                arow.x = p.a_gather[min(m, last)];
                arow.y = p.a_gather[min(m + 1, last)];
                arow.z = p.a_gather[min(m + 2, last)];
                arow.w = p.a_gather[min(m + 3, last)];
            }
            for (long long k0 = kbeg; k0 < kend; k0 += TBK) {
                if (lane == 0) {
                    mbar_wait(&empty_bar[r.stage], r.phase ^ 1);
                    mbar_expect_tx(&full_bar[r.stage], HALF);
                }
                __syncwarp();
                if (p.a_gather) {
                    tma_gather4(tileA(r.stage) + lane * 512, &mapA, &full_bar[r.stage], (int)k0, arow);
                } else if (lane == 0) {
                    if (!p.a_mn) {
                        tma_load_2d(tileA(r.stage), &mapA, &full_bar[r.stage], (int)k0, m0);
                    } else {
#pragma unroll
                        for (int c = 0; c < 4; ++c)
                            tma_load_2d(tileA(r.stage) + c * 4096, &mapA, &full_bar[r.stage], m0 + 32 * c, (int)k0);
                    }
                }
                if (p.b_gather) {   // chunk c (32 N-columns) x k-row group j: k-rows k0+4j .. +3
                    const int j = lane & 7;
                    const long long k = k0 + 4 * j, last = p.K - 1;
                    // rows past K re-read the last valid row: the A tile is zero-filled there, so they contribute 0
                    int4 brow;
                    brow.x = p.b_gather[min(k, last)];
                    brow.y = p.b_gather[min(k + 1, last)];
                    brow.z = p.b_gather[min(k + 2, last)];
                    brow.w = p.b_gather[min(k + 3, last)];
#pragma unroll
                    for (int c = lane >> 3; c < BN / 32; c += 4)
                        tma_gather4(tileB(r.stage) + c * 4096 + j * 512, &mapB, &full_bar[r.stage], n0 + 32 * c, brow);
                } else if (lane == 0) {
                    if (!p.b_mn) {
                        tma_load_2d(tileB(r.stage), &mapB, &full_bar[r.stage], (int)k0, n0);
                    } else {
#pragma unroll
                        for (int c = 0; c < BN / 32; ++c)
                            tma_load_2d(tileB(r.stage) + c * 4096, &mapB, &full_bar[r.stage], n0 + 32 * c, (int)k0);
                    }
                }
                r.advance(stages);
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (one thread) =====================
        if (lane == 0) {
            // instruction descriptor (cute::UMMA::InstrDescriptor): D=F32 [4,6)=1, A/B=TF32 [7,10)=[10,13)=2,
            // a_major [15], b_major [16], N>>3 [17,23), M>>4 [24,29)
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)p.a_mn << 15) | ((uint32_t)p.b_mn << 16) |
                                   ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TBM >> 4) << 24);
            const uint32_t a_lbo = p.a_mn ? 4096 : 16, b_lbo = p.b_mn ? 4096 : 16;
            const uint32_t a_kadv = p.a_mn ? 1024 : 32, b_kadv = p.b_mn ? 1024 : 32;   // bytes per K=8 step
            const uint32_t a_sbo = p.a_mn ? 512 : 1024, b_sbo = p.b_mn ? 512 : 1024;   // 4- vs 8-row swizzle atoms
            const uint32_t a_lay = p.a_mn ? 1 : 2, b_lay = p.b_mn ? 1 : 2;
            StageRing r, acc;
            for (long long t = blockIdx.x; t < total; t += gridDim.x) {
                const long long split = t / tiles_mn;
                const long long kbeg = split * p.k_per_split, kend = min(p.K, kbeg + p.k_per_split);
                mbar_wait(&tempty_bar[acc.stage], acc.phase ^ 1);
                tc_fence_after();
                // two accumulators per tile: hi*hi in the first, the 2^-11-smaller correction terms in the second.  The
                // tensor core truncates on every accumulate, so keeping the main sum at one add per k-step (instead of
                // three) cuts the rounding error 3x; the epilogue adds the two with one IEEE fp32 add.
                const uint32_t d_tmem = tmem_base + acc.stage * acc_cols, d_corr = d_tmem + BN;
                uint32_t first = 1;
                for (long long k0 = kbeg; k0 < kend; k0 += TBK) {
                    mbar_wait(passes == 3 ? &split_bar[r.stage] : &full_bar[r.stage], r.phase);
                    tc_fence_after();
                    if (first && t == blockIdx.x) stamp(2);
                    const uint32_t sa = smem_u32(tileA(r.stage)), sb = smem_u32(tileB(r.stage));
#pragma unroll
                    for (int k = 0; k < TBK / 8; ++k) {
                        const uint64_t da = umma_desc(sa + k * a_kadv, a_lbo, a_sbo, a_lay);
                        const uint64_t db = umma_desc(sb + k * b_kadv, b_lbo, b_sbo, b_lay);
                        tc_mma_tf32(d_tmem, da, db, idesc, first ? 0u : 1u);
                        if (passes == 3) {
                            const uint64_t dal = umma_desc(sa + HALF + k * a_kadv, a_lbo, a_sbo, a_lay);
                            const uint64_t dbl = umma_desc(sb + HALF + k * b_kadv, b_lbo, b_sbo, b_lay);
                            tc_mma_tf32(d_corr, dal, db, idesc, first ? 0u : 1u);
                            tc_mma_tf32(d_corr, da, dbl, idesc, 1u);
                        }
                        first = 0;
                    }
                    tc_commit(&empty_bar[r.stage]);          // frees the smem stage when the MMAs retire
                    r.advance(stages);
                }
                tc_commit(&tfull_bar[acc.stage]);             // accumulator ready for the epilogue
                if (t == blockIdx.x) stamp(3);
                acc.advance(acc_stages);
            }
        }
    } else if (warp < 4) {
        // idle warps of warpgroup 0
    } else if (warp < EPI_WARP0) {
        // ===================== splitters: lo = x - hi for every element of the stage, in shared memory ================
        // hi is what the tensor core itself reads from the raw fp32 word (the top 19 bits: it truncates), so only lo is
        // written.  All loads of a thread's share of the stage are issued before the first store (latency overlap);
        // each warp makes its writes visible to the async proxy and arrives once.
        if (passes == 3) {
            const int tid = threadIdx.x - 128;         // 0..SPLIT_THREADS-1
            constexpr int ITERS = HALF / 16 / SPLIT_THREADS;
            static_assert(HALF % (16 * SPLIT_THREADS) == 0, "stage must divide evenly among the splitter threads");
            StageRing r;
            for (long long t = blockIdx.x; t < total; t += gridDim.x) {
                const long long split = t / tiles_mn;
                const long long kbeg = split * p.k_per_split, kend = min(p.K, kbeg + p.k_per_split);
                for (long long k0 = kbeg; k0 < kend; k0 += TBK) {
                    mbar_wait(&full_bar[r.stage], r.phase);
                    const uint32_t hi = smem_u32(tileA(r.stage)) + tid * 16;      // A and B tiles are contiguous
                    float4 v[ITERS];
#pragma unroll
                    for (int i = 0; i < ITERS; ++i) v[i] = lds128(hi + i * (SPLIT_THREADS * 16));
#pragma unroll
                    for (int i = 0; i < ITERS; ++i)
                        sts128(hi + HALF + i * (SPLIT_THREADS * 16),
                               make_float4(tf32_lo(v[i].x), tf32_lo(v[i].y), tf32_lo(v[i].z), tf32_lo(v[i].w)));
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> async proxy (MMA)
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&split_bar[r.stage]);
                    r.advance(stages);
                }
            }
        }
    } else {
        // ===================== epilogue: TMEM -> registers -> bias/act -> global =====================
        const uint32_t epi_stage = smem_u32(smem + SMEM_DATA) + (warp - EPI_WARP0) * 8192;      // two 4 KB staging buffers
        StageRing acc;
        for (long long t = blockIdx.x; t < total; t += gridDim.x) {
            epilogue_tile(t, acc.stage, acc.phase, epi_stage);
            tc_fence_before();
            mbar_arrive(&tempty_bar[acc.stage]);
            if (threadIdx.x == EPI_WARP0 * 32 && t == blockIdx.x) stamp(5);
            acc.advance(acc_stages);
        }
        if (lane == 0) bulk_wait<0>();          // every TMA store / reduction of this warp has been performed
        __syncwarp();
    }

    tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0) stamp(6);
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
        if (lane == 0) stamp(7);
    }
}

// =====================================================================================================================
// 2-CTA variant (cta_group::2): a CTA pair on one TPC computes a 256 x 256 tile.  Each CTA stages its own 128 A rows and
// HALF of B (128 of the 256 N rows); the leader CTA issues UMMA M=256,N=256 that reads both CTAs' shared memory, and each
// CTA's TMEM receives its 128 rows of D.  Operand bytes fetched from shared memory per FLOP are half of the 1-CTA
// 128x128 kernel, whose profile showed shared-memory bandwidth as the bound.
// Barriers: full/empty/tfull are per CTA (empty and tfull are armed by MULTICAST tcgen05.commit from the leader);
// ready (operands split and visible) and tempty (accumulator drained) live in the LEADER and collect remote arrivals.
// =====================================================================================================================
__device__ __forceinline__ uint32_t cluster_rank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t map_to_cta(uint32_t local_saddr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_saddr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_saddr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_saddr) : "memory");
}
__device__ __forceinline__ bool mbar_try_cluster(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t *bar, uint32_t parity) {
    if (mbar_try_cluster(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_cluster(bar, parity)) {
        if (clock64() - t0 > 4000000000LL) __trap();
    }
}
__device__ __forceinline__ void tc2_commit_mc(uint64_t *bar) {      // arrive on the same barrier offset in BOTH CTAs
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void tc2_mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}

__device__ __forceinline__ void tc2_mma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}

constexpr int T2N = 256;       // tile columns of the pair

constexpr int TC2_THREADS = 640;       // CTA-pair kernel: {TMA, MMA, relay, -} | 8 splitter warps | 8 epilogue warps
constexpr int EPI2_WARP0 = 4 + SPLIT_WARPS, EPI2_WARPS = 8;

template <bool POOL>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TC2_THREADS, 1)
gemm_tc2_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                const __grid_constant__ CUtensorMap mapC, const __grid_constant__ CUtensorMap mapA2,
                const __grid_constant__ CUtensorMap mapB2, TcArgs p) {
    constexpr int TILE = TBM * TBK * 4, HALF = 2 * TILE;          // per CTA: [A hi][B-half hi] | [A lo][B-half lo]
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    __shared__ __align__(8) uint64_t full_bar[MAX_STAGES], ready_bar[MAX_STAGES], empty_bar[MAX_STAGES];
    __shared__ __align__(8) uint64_t tfull_bar[2], tempty_bar[2];
    __shared__ uint32_t tmem_base_smem;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_rank();
    if (threadIdx.x == 0) {     // descriptor fetch overlaps the barrier / TMEM set-up
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&mapA)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&mapB)) : "memory");
    }
    const int stages = p.stages, passes = p.passes;
    // pre-split bf16 planes: three MMAs per k-step and a hi / lo half per stage like 3xTF32, but the lo half is FILLED by the
    // producers (TMA from the lo plane, cp.async gather from the lo table) instead of computed by splitter warps
    const bool planes = p.planes == 2;
    const bool splitting = passes == 3 && !planes;
    // operand tiles are 128 rows x 128 bytes whatever the element type: a stage spans 32 fp32 / 64 bf16 elements of K; an
    // MN-major tile is made of boxes of (128 bytes of MN) x (kstep k-rows): four of 4 KB (fp32) or two of 8 KB (bf16)
    const int kstep = p.elt == 2 ? 64 : TBK;
    const int mn_box_elems = p.elt == 2 ? 64 : 32, mn_box_bytes = kstep * 128, mn_boxes = 128 / mn_box_elems;
    const int stage_bytes = (passes == 3 ? 2 : 1) * HALF;
    const int acc_cols = (passes == 3 ? 2 : 1) * T2N;
    const int acc_stages = 512 / acc_cols;
    auto tileA = [&](int s) { return smem + (size_t)s * stage_bytes; };
    auto tileB = [&](int s) { return smem + (size_t)s * stage_bytes + TILE; };

    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; ++s) {
            // TMA thread (+ every lane of the cp.async gather warps: 2 warps in 3xTF32 mode, 8 in the single-pass modes)
            mbar_init(&full_bar[s], p.lsu_gather ? 1 + 32 * (splitting ? 2 : SPLIT_WARPS) : 1);
            mbar_init(&ready_bar[s], splitting ? 2 * SPLIT_WARPS : 2);        // used in the leader only: one arrival per
            mbar_init(&empty_bar[s], 1);                                       // splitter warp (or relay) of BOTH CTAs
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tfull_bar[a], 1);
            mbar_init(&tempty_bar[a], 2 * EPI2_WARPS);                         // used in the leader only
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    }
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_smem;
    pdl_launch_dependents();            // persistent grid, every CTA resident: the next kernel may set itself up meanwhile
    pdl_wait();                         // everything below reads / writes global memory
    const long long tiles_mn = p.tiles_m * p.tiles_n;
    const long long total = tiles_mn * p.split_k;
    const long long pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;

    // ===================== cp.async gather role: table rows -> the swizzled operand tile ======================================
    // TMA tile::gather4 moves one 128-byte row per ~32 cycles per SM (measured: 1.0 TB/s chip-wide: the fused-gather GEMMs ran
    // 0.10-0.21 ms slower than gather-then-GEMM); the LSU path issues 512 B per instruction.  G warps (the two spare warps of
    // warpgroup 0 in 3xTF32 mode, the eight idle splitter warps in the single-pass modes) copy this CTA's tile of the gathered
    // operand with 16-byte cp.async into exactly the layout TMA would have produced (K-major A: 128-byte rows, 16-byte chunk c
    // of row r stored at chunk c ^ (r & 7); MN-major B: fp32 32-byte units XORed with k-row & 3 inside each 32-column chunk,
    // bf16 the plain 128-byte swizzle), warp gw taking every G-th instruction of the tile.  Completion is asynchronous: the
    // stage's full barrier collects one arrival per lane when that lane's copies have landed (CUTLASS' sm100 mixed
    // TMA + cp.async mainloop signals the MMA the same way), so these warps never wait for data.
    auto gather_role = [&](const int gw, const int G) {
        StageRing r;
        const int elt = p.elt, epc = 16 / elt;           // elements per 16-byte chunk
        const char *Ab = reinterpret_cast<const char *>(p.A), *Bb = reinterpret_cast<const char *>(p.B);
        const char *Ab2 = reinterpret_cast<const char *>(p.A2), *Bb2 = reinterpret_cast<const char *>(p.B2);
        for (long long t = pair; t < total; t += npairs) {
            const long long split = t / tiles_mn, mn = t - split * tiles_mn;
            const long long m0 = (mn / p.tiles_n) * 256 + 128 * rank, n0 = (mn % p.tiles_n) * T2N + 128 * rank;
            const long long kbeg = split * p.k_per_split, kend = min(p.K, kbeg + p.k_per_split);
            int arow[16];                   // A gather: this warp's rows (lane >> 3) + 4 (gw + G i), i < 32 / G  (G >= 2)
            int myrow = -1, myrow2 = -1;    // B gather: the table rows of k-rows k0 + lane (and k0 + 32 + lane for bf16), a stage ahead
            if (p.lsu_gather == 1) {
#pragma unroll
                for (int i = 0; i < 16; ++i)
                    if (gw + G * i < 32) arow[i] = p.a_gather[min(m0 + (lane >> 3) + 4 * (gw + G * i), p.M - 1)];
            } else {
                if (kbeg + lane < p.K) myrow = p.b_gather[kbeg + lane];
                if (elt == 2 && kbeg + 32 + lane < p.K) myrow2 = p.b_gather[kbeg + 32 + lane];
            }
            for (long long k0 = kbeg; k0 < kend; k0 += kstep) {
                int nextrow = -1, nextrow2 = -1, pfrow = -1;
                if (p.lsu_gather == 2 && k0 + kstep < kend) {
                    if (k0 + kstep + lane < p.K) nextrow = p.b_gather[k0 + kstep + lane];
                    if (elt == 2 && k0 + kstep + 32 + lane < p.K) nextrow2 = p.b_gather[k0 + kstep + 32 + lane];
                }
                // weight-gradient form: every stage touches 32 / 64 NEW random table rows (DRAM page misses) and only 3 / 6
                // stages are in flight, so the rows of the stage PF steps ahead are pulled into L2 now (warp gw's share)
                const int PF = p.pf_stages;
                if (p.lsu_gather == 2 && PF > 0) {
                    const long long kp = k0 + (long long)PF * kstep + lane + 32 * gw;
                    if (gw < kstep / 32 && kp < kend && kp < p.K) pfrow = p.b_gather[kp];
                }
                if (lane == 0) mbar_wait(&empty_bar[r.stage], r.phase ^ 1);
                __syncwarp();
                if (p.lsu_gather == 1) {
                    const uint32_t dst = smem_u32(tileA(r.stage));
                    const int c = lane & 7;
                    const long long kk = k0 + epc * c;
                    const int nbytes = (int)max(0LL, min((long long)epc, p.K - kk)) * elt;
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        if (gw + G * i < 32) {
                            const int rr = (lane >> 3) + 4 * (gw + G * i);
                            cp_async16(dst + rr * 128 + ((c ^ (rr & 7)) << 4),
                                       Ab + ((long long)arow[i] * p.lda + (nbytes ? kk : 0)) * elt, nbytes);
                            if (planes)
                                cp_async16(dst + HALF + rr * 128 + ((c ^ (rr & 7)) << 4),
                                           Ab2 + ((long long)arow[i] * p.lda + (nbytes ? kk : 0)) * elt, nbytes);
                        }
                    }
                } else if (elt == 4) {
                    const uint32_t dst = smem_u32(tileB(r.stage));
                    const int c4 = lane >> 3, u = lane & 7;
                    const long long col = n0 + 32 * c4 + 4 * u;
                    const int cbytes = (int)max(0LL, min(4LL, p.N - col)) * 4;
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const int kr = gw + G * i;
                        if (kr < 32) {
                            const int row = __shfl_sync(0xffffffffu, myrow, kr);
                            const int nbytes = row >= 0 ? cbytes : 0;
                            cp_async16(dst + c4 * 4096 + kr * 128 + ((((u >> 1) ^ (kr & 3)) << 5) | ((u & 1) << 4)),
                                       Bb + (nbytes ? ((long long)row * p.ldb + col) * 4 : 0), nbytes);
                        }
                    }
                } else {
                    // bf16, MN-major: a k-row of this CTA's 128 columns is 256 bytes = two 128-byte box rows; one warp
                    // instruction copies two k-rows (16 lanes x 16 bytes each)
                    const uint32_t dst = smem_u32(tileB(r.stage));
                    const int sub = lane >> 4, c2 = (lane >> 3) & 1, u = lane & 7;
                    const long long col = n0 + 64 * c2 + 8 * u;
                    const int cbytes = (int)max(0LL, min(8LL, p.N - col)) * 2;
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const int kp = gw + G * i;
                        if (kp < 32) {
                            const int kr = 2 * kp + sub;
                            const int lo = __shfl_sync(0xffffffffu, myrow, kr & 31), hi = __shfl_sync(0xffffffffu, myrow2, kr & 31);
                            const int row = kr < 32 ? lo : hi;
                            const int nbytes = row >= 0 ? cbytes : 0;
                            cp_async16(dst + c2 * 8192 + kr * 128 + ((u ^ (kr & 7)) << 4),
                                       Bb + (nbytes ? ((long long)row * p.ldb + col) * 2 : 0), nbytes);
                            if (planes)
                                cp_async16(dst + HALF + c2 * 8192 + kr * 128 + ((u ^ (kr & 7)) << 4),
                                           Bb2 + (nbytes ? ((long long)row * p.ldb + col) * 2 : 0), nbytes);
                        }
                    }
                }
                cp_async_arrive_noinc(&full_bar[r.stage]);
                r.advance(stages);
                myrow = nextrow;
                myrow2 = nextrow2;
                if (pfrow >= 0) {           // this CTA's 128 columns of that row: 512 B (fp32) / 256 B (bf16) = 4 / 2 lines
                    const char *a = Bb + ((long long)pfrow * p.ldb + n0) * elt;
                    for (int o = 0; o < 128 * elt; o += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(a + o));
                    if (planes) {
                        const char *a2 = Bb2 + ((long long)pfrow * p.ldb + n0) * elt;
                        for (int o = 0; o < 128 * elt; o += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(a2 + o));
                    }
                }
            }
        }
        cp_async_wait<0>();
    };

    if (warp == 0) {
        // ===================== TMA producer: this CTA's 128 A rows and its half of B =====================
        // (lane 0 drives; with a fused table gather all 32 lanes issue tile::gather4 rows: 4 table rows x 128 B each)
        StageRing r;
        for (long long t = pair; t < total; t += npairs) {
            const long long split = t / tiles_mn, mn = t - split * tiles_mn;
            const int m0 = (int)((mn / p.tiles_n) * 256 + 128 * rank), n0 = (int)((mn % p.tiles_n) * T2N + 128 * rank);
            const long long kbeg = split * p.k_per_split, kend = min(p.K, kbeg + p.k_per_split);
            int4 arow = make_int4(0, 0, 0, 0);
            if (p.a_gather) {       // rows m0+4*lane .. +3 of this CTA's half tile; rows past M re-read the last valid row
                const long long m = m0 + 4 * lane, last = p.M - 1;
                arow.x = p.a_gather[min(m, last)];
                arow.y = p.a_gather[min(m + 1, last)];
                arow.z = p.a_gather[min(m + 2, last)];
                arow.w = p.a_gather[min(m + 3, last)];
            }
            for (long long k0 = kbeg; k0 < kend; k0 += kstep) {
                if (lane == 0) {
                    mbar_wait(&empty_bar[r.stage], r.phase ^ 1);
                    mbar_expect_tx(&full_bar[r.stage], (p.lsu_gather ? TILE : HALF) * (planes ? 2 : 1));
                }
                __syncwarp();
                if (p.lsu_gather == 1) {
                    // A arrives through the cp.async gather warp
                } else if (p.a_gather) {
                    tma_gather4(tileA(r.stage) + lane * 512, &mapA, &full_bar[r.stage], (int)k0, arow);
                } else if (lane == 0) {
                    if (!p.a_mn) {
                        tma_load_2d(tileA(r.stage), &mapA, &full_bar[r.stage], (int)k0, m0);
                        if (planes) tma_load_2d(tileA(r.stage) + HALF, &mapA2, &full_bar[r.stage], (int)k0, m0);
                    } else {
#pragma unroll
                        for (int c = 0; c < 4; ++c)
                            if (c < mn_boxes) {
                                tma_load_2d(tileA(r.stage) + c * mn_box_bytes, &mapA, &full_bar[r.stage], m0 + mn_box_elems * c, (int)k0);
                                if (planes)
                                    tma_load_2d(tileA(r.stage) + HALF + c * mn_box_bytes, &mapA2, &full_bar[r.stage], m0 + mn_box_elems * c, (int)k0);
                            }
                    }
                }
                if (p.lsu_gather == 2) {
                    // B arrives through the cp.async gather warp
                } else if (p.b_gather) {   // chunk c = lane / 8 (32 N-columns) x k-row group j = lane % 8: k-rows k0+4j .. +3
                    const int j = lane & 7, c = lane >> 3;
                    const long long k = k0 + 4 * j, last = p.K - 1;
                    int4 brow;      // rows past K re-read the last valid row: the A tile is zero-filled there
                    brow.x = p.b_gather[min(k, last)];
                    brow.y = p.b_gather[min(k + 1, last)];
                    brow.z = p.b_gather[min(k + 2, last)];
                    brow.w = p.b_gather[min(k + 3, last)];
                    tma_gather4(tileB(r.stage) + c * 4096 + j * 512, &mapB, &full_bar[r.stage], n0 + 32 * c, brow);
                } else if (lane == 0) {
                    if (!p.b_mn) {
                        tma_load_2d(tileB(r.stage), &mapB, &full_bar[r.stage], (int)k0, n0);
                        if (planes) tma_load_2d(tileB(r.stage) + HALF, &mapB2, &full_bar[r.stage], (int)k0, n0);
                    } else {
#pragma unroll
                        for (int c = 0; c < 4; ++c)
                            if (c < mn_boxes) {
                                tma_load_2d(tileB(r.stage) + c * mn_box_bytes, &mapB, &full_bar[r.stage], n0 + mn_box_elems * c, (int)k0);
                                if (planes)
                                    tma_load_2d(tileB(r.stage) + HALF + c * mn_box_bytes, &mapB2, &full_bar[r.stage], n0 + mn_box_elems * c, (int)k0);
                            }
                    }
                }
                r.advance(stages);
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer: one thread of the LEADER CTA =====================
        if (rank == 0 && lane == 0) {
            // instruction descriptor: D = F32, A/B format TF32 (2) or BF16 (1), operand majors, N >> 3, M >> 4
            // (fp16 = 0, bf16 = 1; the hardware rejects a mixed pair: illegal instruction)
            const uint32_t fmt_a = p.elt == 2 ? (p.a_f16 ? 0u : 1u) : 2u, fmt_b = p.elt == 2 ? (p.b_f16 ? 0u : 1u) : 2u;
            const uint32_t idesc = (1u << 4) | (fmt_a << 7) | (fmt_b << 10) | ((uint32_t)p.a_mn << 15) | ((uint32_t)p.b_mn << 16) |
                                   ((uint32_t)(T2N >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
            // one MMA consumes 32 bytes of K (8 tf32 / 16 bf16).  K-major: 128-byte rows, SWIZZLE_128B.  MN-major: k-rows of
            // 128 bytes; fp32 uses the 32-byte-atom swizzle (4-row atoms, 8 k-rows per MMA), bf16 the plain 128-byte swizzle
            // (8-row atoms, 16 k-rows per MMA); LBO = distance between the boxes along MN
            const uint32_t mn_lbo = (uint32_t)mn_box_bytes, mn_kadv = p.elt == 2 ? 2048u : 1024u, mn_sbo = p.elt == 2 ? 1024u : 512u;
            const uint32_t mn_lay = p.elt == 2 ? 2u : 1u;
            const uint32_t a_lbo = p.a_mn ? mn_lbo : 16, b_lbo = p.b_mn ? mn_lbo : 16;
            const uint32_t a_kadv = p.a_mn ? mn_kadv : 32, b_kadv = p.b_mn ? mn_kadv : 32;
            const uint32_t a_sbo = p.a_mn ? mn_sbo : 1024, b_sbo = p.b_mn ? mn_sbo : 1024;
            const uint32_t a_lay = p.a_mn ? mn_lay : 2, b_lay = p.b_mn ? mn_lay : 2;
            StageRing r, acc;
            for (long long t = pair; t < total; t += npairs) {
                const long long split = t / tiles_mn;
                const long long kbeg = split * p.k_per_split, kend = min(p.K, kbeg + p.k_per_split);
                mbar_wait_cluster(&tempty_bar[acc.stage], acc.phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc.stage * acc_cols, d_corr = d_tmem + T2N;
                uint32_t first = 1;
                for (long long k0 = kbeg; k0 < kend; k0 += kstep) {
                    mbar_wait_cluster(&ready_bar[r.stage], r.phase);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(tileA(r.stage)), sb = smem_u32(tileB(r.stage));
#pragma unroll
                    for (int k = 0; k < TBK / 8; ++k) {
                        const uint64_t da = umma_desc(sa + k * a_kadv, a_lbo, a_sbo, a_lay);
                        const uint64_t db = umma_desc(sb + k * b_kadv, b_lbo, b_sbo, b_lay);
                        if (p.elt == 2) {
                            tc2_mma_bf16(d_tmem, da, db, idesc, first ? 0u : 1u);
                            if (planes) {       // x y ~ hi hi + (lo hi + hi lo): the small terms in their own accumulator
                                const uint64_t dal = umma_desc(sa + HALF + k * a_kadv, a_lbo, a_sbo, a_lay);
                                const uint64_t dbl = umma_desc(sb + HALF + k * b_kadv, b_lbo, b_sbo, b_lay);
                                tc2_mma_bf16(d_corr, dal, db, idesc, first ? 0u : 1u);
                                tc2_mma_bf16(d_corr, da, dbl, idesc, 1u);
                            }
                            first = 0;
                            continue;
                        }
                        tc2_mma_tf32(d_tmem, da, db, idesc, first ? 0u : 1u);
                        if (passes == 3) {
                            const uint64_t dal = umma_desc(sa + HALF + k * a_kadv, a_lbo, a_sbo, a_lay);
                            const uint64_t dbl = umma_desc(sb + HALF + k * b_kadv, b_lbo, b_sbo, b_lay);
                            tc2_mma_tf32(d_corr, dal, db, idesc, first ? 0u : 1u);
                            tc2_mma_tf32(d_corr, da, dbl, idesc, 1u);
                        }
                        first = 0;
                    }
                    tc2_commit_mc(&empty_bar[r.stage]);       // both CTAs' producers may refill this stage
                    r.advance(stages);
                }
                tc2_commit_mc(&tfull_bar[acc.stage]);          // both CTAs' epilogues may read their accumulator half
                acc.advance(acc_stages);
            }
        }
    } else if (warp == 2) {
        if (splitting && p.lsu_gather) gather_role(0, 2);
        // ===================== relay (no splitter warps): "my operands have landed" -> leader =====================
        if (!splitting && lane == 0) {
            const uint32_t ready0 = map_to_cta(smem_u32(&ready_bar[0]), 0);
            StageRing r;
            for (long long t = pair; t < total; t += npairs) {
                const long long split = t / tiles_mn;
                const long long kbeg = split * p.k_per_split, kend = min(p.K, kbeg + p.k_per_split);
                for (long long k0 = kbeg; k0 < kend; k0 += kstep) {
                    mbar_wait(&full_bar[r.stage], r.phase);
                    mbar_arrive_cluster(ready0 + 8 * r.stage);
                    r.advance(stages);
                }
            }
        }
    } else if (warp == 3) {
        if (splitting && p.lsu_gather) gather_role(1, 2);
    } else if (warp < EPI2_WARP0) {
        if (!splitting && p.lsu_gather) gather_role(warp - 4, SPLIT_WARPS);       // nothing to split: these warps gather
        // ===================== splitters (3xTF32): lo = x - trunc_tf32(x) for this CTA's tiles =====================
        if (splitting) {
            const int tid = threadIdx.x - 128;
            constexpr int ITERS = HALF / 16 / SPLIT_THREADS;
            const uint32_t ready0 = map_to_cta(smem_u32(&ready_bar[0]), 0);
            StageRing r;
            for (long long t = pair; t < total; t += npairs) {
                const long long split = t / tiles_mn;
                const long long kbeg = split * p.k_per_split, kend = min(p.K, kbeg + p.k_per_split);
                for (long long k0 = kbeg; k0 < kend; k0 += kstep) {
                    mbar_wait(&full_bar[r.stage], r.phase);
                    const uint32_t hi = smem_u32(tileA(r.stage)) + tid * 16;
                    float4 v[ITERS];
#pragma unroll
                    for (int i = 0; i < ITERS; ++i) v[i] = lds128(hi + i * (SPLIT_THREADS * 16));
#pragma unroll
                    for (int i = 0; i < ITERS; ++i)
                        sts128(hi + HALF + i * (SPLIT_THREADS * 16),
                               make_float4(tf32_lo(v[i].x), tf32_lo(v[i].y), tf32_lo(v[i].z), tf32_lo(v[i].w)));
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) mbar_arrive_cluster(ready0 + 8 * r.stage);
                    r.advance(stages);
                }
            }
        }
    } else {
        // ===================== epilogue: this CTA's 128 rows of the pair's tile; 2 warps per TMEM lane quarter ==========
        const int q = warp & 3, colhalf = (warp - EPI2_WARP0) >> 2;     // columns [128 * colhalf, +128)
        const bool vec_ok = (p.ldc % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0) &&
                            (!p.aux || (reinterpret_cast<uintptr_t>(p.aux) & 15) == 0);
        const bool bias_vec = (reinterpret_cast<uintptr_t>(p.bias) & 15) == 0;
        const uint32_t tempty0 = map_to_cta(smem_u32(&tempty_bar[0]), 0);
        const uint32_t epi_stage = smem_u32(smem + SMEM_DATA) + (warp - EPI2_WARP0) * 4096;
        StageRing acc;
        for (long long t = pair; t < total; t += npairs) {
            const long long split = t / tiles_mn, mn = t - split * tiles_mn;
            const long long m0 = (mn / p.tiles_n) * 256 + 128 * rank, n0 = (mn % p.tiles_n) * T2N + 128 * colhalf;
            mbar_wait(&tfull_bar[acc.stage], acc.phase);
            tc_fence_after();
            float lp = 0.f;             // POOL: this lane's row, partial logit over this warp's 128 hidden units
#pragma unroll 1
            for (int c = 0; c < 4; ++c) {
                float r[32];
                const uint32_t taddr = tmem_base + acc.stage * acc_cols + 128 * colhalf + c * 32 + ((uint32_t)(32 * q) << 16);
                tc_ld32(taddr, r);
                if (passes == 3) {
                    float corr[32];
                    tc_ld32(taddr + T2N, corr);
#pragma unroll
                    for (int j = 0; j < 32; ++j) r[j] += corr[j];
                }
                if (m0 + 32 * q >= p.M || n0 + c * 32 >= p.N) continue;        // warp-uniform
                if (POOL) {             // bias + tanh + <., w2> in the row-per-lane layout, then the plain coalesced store of hid
                    // (fc1 bias / fc2 weight: warp-uniform L1-resident loads — the 227 KB of shared memory are spoken for)
                    const float *b1s = p.bias + n0 + c * 32, *w2s = p.pool.w2 + n0 + c * 32;
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const float h = tanh_fast(r[j] + __ldg(b1s + j));
                        lp = fmaf(h, __ldg(w2s + j), lp);
                        r[j] = h;
                    }
                    if (p.c_tma) tma_store_block32<0>(r, epi_stage, &mapC, n0 + c * 32, m0 + 32 * q, false, lane);
                    else epi_block32(r, epi_stage, p, m0 + 32 * q, n0 + c * 32, split, lane, vec_ok, bias_vec, XNRS_ACT_NONE, nullptr);
                } else if (p.c_tma) {
                    bias_act_block32(r, (p.bias && (p.split_k == 1 || split == 0)) ? p.bias : nullptr, n0 + c * 32, p.N, p.act, p.aux, p.ldc,
                                     m0 + 32 * q + lane, p.M);
                    tma_store_block32<0>(r, epi_stage, &mapC, n0 + c * 32, m0 + 32 * q, p.accumulate || p.split_k > 1, lane);
                } else {
                    epi_block32(r, epi_stage, p, m0 + 32 * q, n0 + c * 32, split, lane, vec_ok, bias_vec, p.act, p.bias);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(tempty0 + 8 * acc.stage);
            acc.advance(acc_stages);
            if (POOL) {
                if (p.c_tma) {          // the staging buffers double as scratch below: the last hid store must have read its block
                    if (lane == 0) bulk_wait_read<0>();
                    __syncwarp();
                }
                // TMEM is released: the next tile's MMAs run while the pooling weights and weighted sums are formed
                // scratch lives in the epilogue warps' own (now idle) staging buffers: warp we keeps its 32 partial logits in
                // floats [0,32) of its buffer, the column-half-0 warps keep e of their 32 rows in floats [32,64) of theirs
                auto stagef = [&](int we) { return reinterpret_cast<float *>(smem + SMEM_DATA + we * 4096); };
                const int we = warp - EPI2_WARP0;              // = 4 * colhalf + q
                stagef(we)[lane] = lp;
                asm volatile("bar.sync 1, 256;" ::: "memory");
                if (colhalf == 0) {
                    const long long g = m0 + 32 * q + lane;
                    float e = 0.f;
                    if (g < p.M) {
                        if (__ldg(p.pool.tix + g) >= 0) e = expf(stagef(q)[lane] + stagef(4 + q)[lane] + __ldg(p.pool.b2));
                        p.pool.e[g] = e;
                    }
                    stagef(q)[32 + lane] = e;
                }
                asm volatile("bar.sync 1, 256;" ::: "memory");
                if (!p.pool.pooled) continue;       // the per-title sums are formed by titlepool_wsum_kernel from e (see host side)
                // weighted sum over this CTA's 128 rows: warp w8 owns a slice of F/8 columns (one float4 per lane, F/32 lanes
                // active) and walks ALL rows, eight independent row loads in flight; the x rows were fetched from the table
                // moments ago (L2 hits).  A title's partial sums leave through 16-byte vector reductions when the title ends.
                const int w8 = warp - EPI2_WARP0, F4 = p.pool.F4, per = F4 >> 3;     // float4 columns per warp (24 at F=768)
                if (p.elt == 2) {
                    // bf16 rows: 16-byte loads carry 8 elements, so a warp's F/8-column slice is F/64 lanes wide (12 at F=768)
                    const int per8 = F4 >> 4;
                    const bool act8 = lane < per8;
                    const uint4 *x8 = reinterpret_cast<const uint4 *>(p.A) + (act8 ? w8 * per8 + lane : 0);
                    float *poolf = p.pool.pooled + (act8 ? (w8 * per8 + lane) * 8 : 0);
                    const long long ld8 = p.lda >> 3;
                    int tixv[4], srcv[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const long long g = m0 + lane + 32 * j;
                        tixv[j] = g < p.M ? __ldg(p.pool.tix + g) : -1;
                        srcv[j] = g < p.M ? (p.a_gather ? __ldg(p.a_gather + g) : (int)g) : 0;
                    }
                    float a8[8];
#pragma unroll
                    for (int q8 = 0; q8 < 8; ++q8) a8[q8] = 0.f;
                    float zs = 0.f;
                    int cur = -1;
                    auto flush8 = [&]() {
                        if (cur >= 0) {
                            if (act8) {
                                float4 *dst = reinterpret_cast<float4 *>(poolf + (long long)cur * F4 * 4);
                                atomicAdd(dst, make_float4(a8[0], a8[1], a8[2], a8[3]));
                                atomicAdd(dst + 1, make_float4(a8[4], a8[5], a8[6], a8[7]));
                            }
                            if (lane == 0 && w8 == 0) atomicAdd(p.pool.zsum + cur, zs);
                        }
                    };
#pragma unroll 1
                    for (int r0 = 0; r0 < 128; r0 += 8) {
                        uint4 v[8];
                        int ti[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const int row = r0 + i, j = row >> 5;
                            const int tj = j == 0 ? tixv[0] : (j == 1 ? tixv[1] : (j == 2 ? tixv[2] : tixv[3]));
                            const int sj = j == 0 ? srcv[0] : (j == 1 ? srcv[1] : (j == 2 ? srcv[2] : srcv[3]));
                            ti[i] = __shfl_sync(0xffffffffu, tj, row & 31);
                            const long long src = __shfl_sync(0xffffffffu, sj, row & 31);
                            v[i] = (act8 && ti[i] >= 0) ? __ldg(x8 + src * ld8) : make_uint4(0u, 0u, 0u, 0u);
                        }
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const int row = r0 + i;
                            if (ti[i] != cur) {
                                flush8();
                                cur = ti[i];
                                zs = 0.f;
#pragma unroll
                                for (int q8 = 0; q8 < 8; ++q8) a8[q8] = 0.f;
                            }
                            if (ti[i] >= 0) {
                                const float e = stagef(row >> 5)[32 + (row & 31)];
                                zs += e;
                                const uint32_t w[4] = {v[i].x, v[i].y, v[i].z, v[i].w};
#pragma unroll
                                for (int q8 = 0; q8 < 4; ++q8) {        // bf16 -> fp32 is a 16-bit shift
                                    a8[2 * q8] = fmaf(e, __uint_as_float(w[q8] << 16), a8[2 * q8]);
                                    a8[2 * q8 + 1] = fmaf(e, __uint_as_float(w[q8] & 0xffff0000u), a8[2 * q8 + 1]);
                                }
                            }
                        }
                    }
                    flush8();
                    asm volatile("bar.sync 1, 256;" ::: "memory");
                    continue;
                }
                const bool active = lane < per;
                const float4 *x4 = reinterpret_cast<const float4 *>(p.A) + (active ? w8 * per + lane : 0);
                float4 *pool4 = reinterpret_cast<float4 *>(p.pool.pooled) + (active ? w8 * per + lane : 0);
                const long long ld4 = p.lda >> 2;
                // lane l keeps title / source row of rows l, l+32, l+64, l+96 (coalesced loads, broadcast by shuffle)
                int tixv[4], srcv[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const long long g = m0 + lane + 32 * j;
                    tixv[j] = g < p.M ? __ldg(p.pool.tix + g) : -1;
                    srcv[j] = g < p.M ? (p.a_gather ? __ldg(p.a_gather + g) : (int)g) : 0;
                }
                float4 accv = make_float4(0.f, 0.f, 0.f, 0.f);
                float zs = 0.f;
                int cur = -1;
#pragma unroll 1
                for (int r0 = 0; r0 < 128; r0 += 8) {
                    float4 v[8];
                    int ti[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {                   // all eight row loads first ...
                        const int row = r0 + i, j = row >> 5;
                        const int tj = j == 0 ? tixv[0] : (j == 1 ? tixv[1] : (j == 2 ? tixv[2] : tixv[3]));
                        const int sj = j == 0 ? srcv[0] : (j == 1 ? srcv[1] : (j == 2 ? srcv[2] : srcv[3]));
                        ti[i] = __shfl_sync(0xffffffffu, tj, row & 31);
                        const long long src = __shfl_sync(0xffffffffu, sj, row & 31);
                        v[i] = (active && ti[i] >= 0) ? __ldg(x4 + src * ld4) : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
#pragma unroll
                    for (int i = 0; i < 8; ++i) {                   // ... then the running per-title sums
                        const int row = r0 + i;
                        if (ti[i] != cur) {
                            if (cur >= 0) {
                                if (active) atomicAdd(pool4 + (long long)cur * F4, accv);
                                if (lane == 0 && w8 == 0) atomicAdd(p.pool.zsum + cur, zs);
                            }
                            cur = ti[i];
                            zs = 0.f;
                            accv = make_float4(0.f, 0.f, 0.f, 0.f);
                        }
                        if (ti[i] >= 0) {
                            const float e = stagef(row >> 5)[32 + (row & 31)];
                            zs += e;
                            accv.x = fmaf(e, v[i].x, accv.x); accv.y = fmaf(e, v[i].y, accv.y);
                            accv.z = fmaf(e, v[i].z, accv.z); accv.w = fmaf(e, v[i].w, accv.w);
                        }
                    }
                }
                if (cur >= 0) {
                    if (active) atomicAdd(pool4 + (long long)cur * F4, accv);
                    if (lane == 0 && w8 == 0) atomicAdd(p.pool.zsum + cur, zs);
                }
                asm volatile("bar.sync 1, 256;" ::: "memory");     // the staging buffers are reused by the next tile's stores
            }
        }
        if (p.c_tma) {
            if (lane == 0) bulk_wait<0>();      // every TMA store / reduction of this warp has been performed
            __syncwarp();
        }
    }

    tc_fence_before();
    cluster_sync_all();         // neither CTA may exit (or free TMEM) while its peer can still touch its smem / barriers
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
    }
}

// ---- host side ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// 2-D fp32 tensor map: inner (contiguous) extent `inner`, `outer` rows of stride ld floats, box {32, box_outer}
static bool make_map(CUtensorMap *map, const float *base, long long inner, long long outer, long long ld, int box_outer,
                     bool mn_major) {
    if (outer <= 0) outer = 1;
    EncodeTiledFn fn = encode_fn();
    if (!fn) return false;
    cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
    cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(float)};
    cuuint32_t box[2] = {32, (cuuint32_t)box_outer};
    cuuint32_t estr[2] = {1, 1};
    return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(base), dims, strides, box, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
              CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// 2-D bf16 tensor map: box {64 elements = 128 bytes, box_outer rows}, 128-byte swizzle (K-major and MN-major alike)
static bool make_map_bf16(CUtensorMap *map, const void *base, long long inner, long long outer, long long ld, int box_outer) {
    if (outer <= 0) outer = 1;
    EncodeTiledFn fn = encode_fn();
    if (!fn) return false;
    cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)box_outer};
    cuuint32_t estr[2] = {1, 1};
    return fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(base), dims, strides, box, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

int g_opt_2cta = -2;       // -2: read XNRS_GEMM_2CTA on first use

// C leaves through TMA stores of staged 32 x 32 blocks when its layout allows a tensor map (fp32 C, 16-byte aligned rows);
// XNRS_TMA_STORE=0 keeps the coalesced-store epilogue everywhere
static int c_tma_map(CUtensorMap *mapC, const float *C, long long M, long long N, long long ldc, const float *bias, int act,
                     bool c_bf16, const float *aux = nullptr) {
    static int on = -1;
    if (on < 0) { const char *e = getenv("XNRS_TMA_STORE"); on = e ? atoi(e) : 1; }
    return on && !c_bf16 && ldc % 4 == 0 && N % 4 == 0 && !((uintptr_t)C & 15) && !((uintptr_t)bias & 15) &&
           (act != XNRS_ACT_RELU_MASK || (aux && !((uintptr_t)aux & 15))) && make_map(mapC, C, N, M, ldc, 32, false);
}
static long long *g_gemm_trace = nullptr;     // xnrs_debug_gemm_trace

static int gather_prefetch_stages() {       // XNRS_GATHER_PF: stages of look-ahead of the L2 prefetch in the gathered dW GEMM
    static int pf = -1;
    if (pf < 0) { const char *e = getenv("XNRS_GATHER_PF"); pf = e ? atoi(e) : 4; if (pf < 0) pf = 0; }
    return pf;
}

int gemm_tensorcore(const GemmArgs &a, int precision, cudaStream_t st, int *status) {
    // shapes / layouts the TMA path cannot take fall through to the exact-fp32 SIMT kernel
    if (a.a_rows && a.transA) return 0;          // gather is fused for K-major A (forward) ...
    if (a.b_rows && a.transB) return 0;          // ... and MN-major B (weight gradient); other combinations: SIMT
    if (a.M < 128 || a.N < 32 || a.K < 32) return 0;
    if (a.lda % 4 || a.ldb % 4) return 0;
    if (((uintptr_t)a.A & 15) || ((uintptr_t)a.B & 15)) return 0;
    if (a.M > 2000000000LL || a.N > 2000000000LL || a.K > 2000000000LL) return 0;
    static int is_sm100 = -1;
    if (is_sm100 < 0) is_sm100 = xnrs_device_is_sm100();
    if (!is_sm100) return 0;

    TcArgs p;
    p.M = a.M; p.N = a.N; p.K = a.K;
    p.C = a.C; p.ldc = a.ldc; p.bias = a.bias; p.act = a.act; p.aux = a.aux; p.accumulate = a.accumulate;
    p.a_mn = a.transA ? 1 : 0;      // transA: A stored [K, M]
    p.b_mn = a.transB ? 0 : 1;      // transB: B stored [N, K] (K-major); else stored [K, N]
    p.a_gather = a.a_rows; p.b_gather = a.b_rows;
    p.A = a.A; p.lda = a.lda; p.B = a.B; p.ldb = a.ldb;
    p.lsu_gather = 0;
    p.pf_stages = gather_prefetch_stages();
    p.elt = 4; p.c_bf16 = 0;
    memset(&p.pool, 0, sizeof(p.pool));
    p.trace = nullptr;
    p.passes = (precision == XNRS_PREC_TF32X3) ? 3 : 1;
    // 3xTF32 keeps BN=128: its stage is 2x larger (hi+lo), and 3 smem stages + 2 TMEM stages beat the wider tile
    // (measured: 156 vs 142 TFLOP/s); single-pass TF32 takes BN=256 (398 vs 340 TFLOP/s)
    // CTA-pair kernel (256x256 tiles, cta_group::2): wide outputs with enough rows and no fused gather.  It halves the
    // operand bytes the tensor core reads from shared memory per FLOP — the resource that bounds the 3xTF32 main loop
    // (ncu: tensor-core + splitter wavefronts = 94 % of the shared-memory pipe in the 1-CTA kernel) — and is the default
    // there (measured 3xTF32: qkv 198 -> 221, fc1 184 -> 200, dW 173 -> 179 TFLOP/s).  Single-pass TF32 keeps the 1-CTA
    // BN=256 kernel (its short stages make the cross-CTA handshake the critical path: 525 vs 307 TFLOP/s).
    // XNRS_GEMM_2CTA / xnrs_set_option("gemm_2cta"): 0 = never, 1 = always where legal, -1 (default) = 3xTF32 only.
    if (g_opt_2cta == -2) {
        const char *e = getenv("XNRS_GEMM_2CTA");
        g_opt_2cta = e ? atoi(e) : -1;
    }
    const bool want2 = g_opt_2cta == 1 || (g_opt_2cta == -1 && p.passes == 3);
    bool use2 = want2 && a.N > 128 && a.M >= 256 && (num_sms() % 2 == 0);
    // the pair kernel has 74 workers with 256x256 tiles: problems with fewer than two waves of such tiles (launch list of a
    // CL step: 20-40 us for 0.1-3 GFLOP GEMMs on 4-70 CTAs) run on the 1-CTA kernel, whose 128x128 tiles spread the same
    // work over four times as many CTAs with half the per-stage latency
    if (use2 && g_opt_2cta == -1) {
        const long long t2 = cdiv(a.M, 256) * cdiv(a.N, 256);
        const long long ksplit = a.act == XNRS_ACT_NONE ? (a.split_k > 0 ? a.split_k : std::max(1LL, std::min(cdiv(a.K, 512), (long long)(num_sms() / 2) / t2))) : 1;
        const long long npairs = num_sms() / 2;
        // split-K launches size their splits to one full wave of pairs; un-split ones need at least two waves of tiles
        if (ksplit > 1 ? (t2 * ksplit * 10 < npairs * 9) : (t2 < 2 * npairs)) use2 = false;
    }
    const int BN = use2 ? 128 : ((p.passes == 1 && a.N > 128 && cdiv(a.M, TBM) * cdiv(a.N, 256) >= num_sms()) ? 256 : 128);
    const int half = TBM * TBK * 4 + BN * TBK * 4;
    p.stages = SMEM_DATA / (half * (p.passes == 3 ? 2 : 1));
    if (p.stages > MAX_STAGES) p.stages = MAX_STAGES;
    p.tiles_m = cdiv(a.M, use2 ? 256 : TBM);
    p.tiles_n = cdiv(a.N, use2 ? 256 : BN);
    long long tiles = p.tiles_m * p.tiles_n;
    int split = a.split_k;
    if (split <= 0) {
        long long want = use2 ? num_sms() / 2 : num_sms();
        long long s = tiles >= want ? 1 : want / tiles;
        // a CTA's K loop is latency bound (3-6 stages in flight, ~0.6 us per 32-wide k-block at 3 passes), so problems that
        // leave SMs idle are cut into splits of >= 4 k-blocks; outputs that are not accumulated into pay a memset for it,
        // which only pays off from K = 512 on
        long long maxs = (a.accumulate || a.K >= 512) ? cdiv(a.K, 128) : 1;
        if (s > maxs) s = maxs;
        if (s < 1) s = 1;
        if (a.act != XNRS_ACT_NONE) s = 1;
        split = (int)s;
    }
    if (a.act == XNRS_ACT_NONE && cdiv(a.K, split) > KCHUNK) split = (int)cdiv(a.K, KCHUNK);
    if (split > 1 && a.act != XNRS_ACT_NONE) {
        *status = fail(XNRS_ERR_ARG, "%s: split_k with activation", "xnrs_gemm");
        return 1;
    }
    p.k_per_split = cdiv(cdiv(a.K, split), TBK) * TBK;
    split = (int)cdiv(a.K, p.k_per_split);         // rounding k_per_split up can empty the last splits: a split without
    p.split_k = split;                              // k-blocks would add an accumulator no MMA ever wrote

    p.trace = g_gemm_trace;
    CUtensorMap mapA, mapB, mapC;
    p.c_tma = c_tma_map(&mapC, a.C, a.M, a.N, a.ldc, a.bias, a.act, false, a.aux);
    // gathered operands: the map spans the whole table (row count unknown to the GEMM: use the int32 range) and the box
    // is one row high — tile::gather4 fetches four such rows per instruction
    const long long table_rows = 0x7fffffffLL;
    bool ok = p.a_mn ? make_map(&mapA, a.A, a.M, a.K, a.lda, 32, true)
                     : make_map(&mapA, a.A, a.K, a.a_rows ? table_rows : a.M, a.lda, a.a_rows ? 1 : TBM, false);
    ok = ok && (p.b_mn ? make_map(&mapB, a.B, a.N, a.b_rows ? table_rows : a.K, a.ldb, a.b_rows ? 1 : 32, true)
                       : make_map(&mapB, a.B, a.K, a.N, a.ldb, BN, false));
    if (!ok) return 0;
    if (!p.c_tma) mapC = mapA;
    const CUtensorMap mapA2 = mapA, mapB2 = mapB;      // (lo planes: bf16 pre-split mode only)

    if (split > 1 && !a.accumulate) {
        if (cudaMemset2DAsync(a.C, a.ldc * sizeof(float), 0, a.N * sizeof(float), a.M, st) != cudaSuccess) {
            *status = fail(XNRS_ERR_CUDA, "%s: memset2d failed", "xnrs_gemm");
            return 1;
        }
    }
    static bool attr_set = false;
    const int smem_bytes = SMEM_DATA + SMEM_EPI + 1024;
    if (!attr_set) {
        if (cudaFuncSetAttribute(gemm_tc_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes) != cudaSuccess ||
            cudaFuncSetAttribute(gemm_tc_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes) != cudaSuccess ||
            cudaFuncSetAttribute(gemm_tc2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes) != cudaSuccess) {
            cudaGetLastError();
            return 0;
        }
        attr_set = true;
    }
    long long total = tiles * split;
    unsigned grid = (unsigned)(total < num_sms() ? total : num_sms());
    if (use2) {
        static int lsu = -1;                    // XNRS_LSU_GATHER=0 falls back to TMA tile::gather4 (kept for comparison)
        if (lsu < 0) { const char *e = getenv("XNRS_LSU_GATHER"); lsu = e ? atoi(e) : 1; }
        if (lsu) p.lsu_gather = a.a_rows ? 1 : (a.b_rows ? 2 : 0);
        const long long pairs = total < num_sms() / 2 ? total : num_sms() / 2;
        launch_pdl(gemm_tc2_kernel<false>, dim3((unsigned)(2 * pairs)), dim3(TC2_THREADS), smem_bytes, st, mapA, mapB, mapC, mapA2, mapB2, p);
        g_last_gemm_kernel = p.passes == 3 ? "gemm_tc2_kernel (cta_group::2, 256x256 pair tile, 3xTF32)"
                                           : "gemm_tc2_kernel (cta_group::2, 256x256 pair tile, TF32)";
    } else if (BN == 256) {
        launch_pdl(gemm_tc_kernel<256>, dim3(grid), dim3(TC_THREADS), smem_bytes, st, mapA, mapB, mapC, p);
        g_last_gemm_kernel = p.passes == 3 ? "gemm_tc_kernel<256> (3xTF32)" : "gemm_tc_kernel<256> (TF32)";
    } else {
        launch_pdl(gemm_tc_kernel<128>, dim3(grid), dim3(TC_THREADS), smem_bytes, st, mapA, mapB, mapC, p);
        g_last_gemm_kernel = p.passes == 3 ? "gemm_tc_kernel<128> (3xTF32)" : "gemm_tc_kernel<128> (TF32)";
    }
    g_launches.fetch_add(1, std::memory_order_relaxed);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        snprintf(g_err, sizeof(g_err), "xnrs_gemm (tcgen05): CUDA error: %s", cudaGetErrorString(e));
        *status = XNRS_ERR_CUDA;
        return 1;
    }
    *status = XNRS_OK;
    return 1;
}

// Second half of the fused pooling when the weighted sum is NOT done in the GEMM epilogue: one warp per title turns the
// un-normalised weights e into attn = e / (sum e + 1e-8) and pooled = sum attn * x (layers.py:62-65), reading the title's
// rows straight from the (fp32 or bf16) table with four rows of loads in flight; padding rows get attn = 0.
template <bool BF>
__global__ void __launch_bounds__(256)
titlepool_wsum_kernel(const void *__restrict__ x, long long ldx, const int *__restrict__ x_rows, const int *__restrict__ seg,
                      const float *__restrict__ e, long long R, long long n_rows, int F, float *__restrict__ attn,
                      float *__restrict__ pooled) {
    const int lane = threadIdx.x & 31;
    const long long wid = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5, nw = ((long long)gridDim.x * blockDim.x) >> 5;
    constexpr int EPC = BF ? 8 : 4;                 // elements per 16-byte chunk
    const int nchunk = F / EPC;                     // 16-byte chunks per row (192 fp32 / 96 bf16 at F = 768)
    for (long long r = wid; r < R; r += nw) {
        const long long base = seg[r];
        const int L = seg[r + 1] - seg[r];
        float z = 0.f;
        for (int l = lane; l < L; l += 32) z += e[base + l];
        z = warp_sum(z);
        const float denom = z + 1e-8f;
        for (int l = lane; l < L; l += 32) attn[base + l] = e[base + l] / denom;
        float acc[6][EPC];
#pragma unroll
        for (int c = 0; c < 6; ++c)
#pragma unroll
            for (int q = 0; q < EPC; ++q) acc[c][q] = 0.f;
        for (int l0 = 0; l0 < L; l0 += 2) {
            uint4 v[2][6];
            float a[2];
#pragma unroll
            for (int i = 0; i < 2; ++i) {           // every load of both rows first ...
                const int l = min(l0 + i, L - 1);
                a[i] = (l0 + i < L) ? e[base + l] / denom : 0.f;
                const long long row = x_rows ? (long long)x_rows[base + l] : base + l;
                const uint4 *src = reinterpret_cast<const uint4 *>(reinterpret_cast<const char *>(x) + row * ldx * (BF ? 2 : 4));
#pragma unroll
                for (int c = 0; c < 6; ++c)
                    v[i][c] = (lane + 32 * c < nchunk) ? __ldg(src + lane + 32 * c) : make_uint4(0u, 0u, 0u, 0u);
            }
#pragma unroll
            for (int i = 0; i < 2; ++i) {           // ... then the arithmetic
#pragma unroll
                for (int c = 0; c < 6; ++c) {
                    const uint32_t w[4] = {v[i][c].x, v[i][c].y, v[i][c].z, v[i][c].w};
                    if (BF) {
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            acc[c][2 * q] = fmaf(a[i], __uint_as_float(w[q] << 16), acc[c][2 * q]);
                            acc[c][2 * q + 1] = fmaf(a[i], __uint_as_float(w[q] & 0xffff0000u), acc[c][2 * q + 1]);
                        }
                    } else {
#pragma unroll
                        for (int q = 0; q < 4; ++q) acc[c][q] = fmaf(a[i], __uint_as_float(w[q]), acc[c][q]);
                    }
                }
            }
        }
#pragma unroll
        for (int c = 0; c < 6; ++c) {
            const int ch = lane + 32 * c;
            if (ch < nchunk) {
                float4 *dst = reinterpret_cast<float4 *>(pooled + r * F + (long long)ch * EPC);
                dst[0] = make_float4(acc[c][0], acc[c][1], acc[c][2], acc[c][3]);
                if (BF) dst[1] = make_float4(acc[c][4 % EPC], acc[c][5 % EPC], acc[c][6 % EPC], acc[c][7 % EPC]);
            }
        }
    }
    // rows past the last title (TitlePlan padding) carry no weight
    const long long t0 = seg[R];
    for (long long g = t0 + blockIdx.x * (long long)blockDim.x + threadIdx.x; g < n_rows; g += (long long)gridDim.x * blockDim.x)
        attn[g] = 0.f;
}

// attn[g] = e[g] / (zsum[title] + 1e-8) (0 on padding rows); pooled[t,:] /= (zsum[t] + 1e-8)   (layers.py:62-65)
__global__ void titlepool_finalize_kernel(const float *__restrict__ e, const int *__restrict__ tix, const float *__restrict__ zsum,
                                          long long n_rows, long long R, int F4, float *__restrict__ attn, float *__restrict__ pooled) {
    const long long tid = blockIdx.x * (long long)blockDim.x + threadIdx.x, nt = (long long)gridDim.x * blockDim.x;
    for (long long g = tid; g < n_rows; g += nt) {
        const int t = tix[g];
        attn[g] = t >= 0 ? e[g] / (zsum[t] + 1e-8f) : 0.f;
    }
    float4 *p4 = reinterpret_cast<float4 *>(pooled);
    for (long long i = tid; i < R * F4; i += nt) {
        const float d = zsum[i / F4] + 1e-8f;
        float4 v = p4[i];
        v.x /= d; v.y /= d; v.z /= d; v.w /= d;
        p4[i] = v;
    }
}

}  // namespace xnrs

using namespace xnrs;

// gather -> fc1 (+bias, tanh) -> <., w2> + b2 -> exp -> per-title sum(e), sum(e * x) in ONE launch of the CTA-pair tcgen05
// kernel (cp.async gather warp + pooling epilogue), then a small normalisation pass.  Returns XNRS_ERR_UNSUPPORTED (nothing
// launched) for shapes / devices / precisions the fused kernel does not cover; the caller then runs xnrs_gemm + xnrs_addpool_fwd.
// shared by the fp32-storage (tf32 / 3xtf32) and the bf16-storage entry points: x / w1 / hid are fp32 (elt 4) or bf16 (elt 2)
// x_lo / w1_lo (both or neither): the operands arrive pre-split into two bf16 planes (x ~ x + x_lo): elt = 2, three
// kind::f16 MMAs per k-step, fp32 hid; the per-title weighted sums then read the fp32 rows x_sum (ld_sum floats apart)
static int titlepool_fwd_impl(const void *x, long long ldx, const int *x_rows, const int *tix, const int *seg, long long n_rows,
                              long long R, int F, int A, const void *w1, const float *b1, const float *w2, const float *b2,
                              int passes, int elt, void *hid, float *e, float *zsum, float *attn, float *pooled, cudaStream_t st,
                              const void *x_lo = nullptr, const void *w1_lo = nullptr, const float *x_sum = nullptr,
                              long long ld_sum = 0, int planes_f16 = 0) {
    const bool planes = x_lo != nullptr;
    XNRS_REQUIRE(x && tix && w1 && b1 && w2 && b2 && hid && e && zsum && attn && pooled, "null pointer");
    XNRS_REQUIRE(!planes || (w1_lo && x_sum && seg && F <= 768 && elt == 2 && passes == 3 && !((uintptr_t)x_lo & 15) &&
                             !((uintptr_t)w1_lo & 15) && !((uintptr_t)x_sum & 15) && ld_sum % 4 == 0), "pre-split planes");
    XNRS_REQUIRE(n_rows >= 0 && R >= 0 && F > 0 && A > 0, "bad sizes");
    static int is_sm100 = -1;
    if (is_sm100 < 0) is_sm100 = xnrs_device_is_sm100();
    if (!is_sm100 || A != 256 || F % 128 || F > 1024 || n_rows < 256 || ldx % (16 / elt) || ((uintptr_t)x & 15) ||
        ((uintptr_t)w1 & 15) || ((uintptr_t)hid & 15) || ((uintptr_t)pooled & 15) || num_sms() % 2)
        return fail(XNRS_ERR_UNSUPPORTED, "%s: shape / device / precision not covered by the fused kernel", "xnrs_titlepool_fwd");
    TcArgs p;
    memset(&p, 0, sizeof(p));
    p.M = n_rows; p.N = A; p.K = F;
    p.C = reinterpret_cast<float *>(hid); p.ldc = A; p.bias = b1; p.act = XNRS_ACT_TANH; p.aux = nullptr; p.accumulate = 0;
    p.elt = elt; p.c_bf16 = elt == 2 && !planes;
    p.planes = planes ? 2 : 0;
    p.a_f16 = p.b_f16 = planes && planes_f16;
    p.A2 = reinterpret_cast<const float *>(x_lo); p.B2 = reinterpret_cast<const float *>(w1_lo);
    const int kstep = elt == 2 ? 64 : TBK;
    p.split_k = 1; p.k_per_split = cdiv(F, kstep) * kstep;
    p.a_mn = 0; p.b_mn = 0;
    p.a_gather = x_rows; p.b_gather = nullptr;
    p.A = reinterpret_cast<const float *>(x); p.lda = ldx; p.B = reinterpret_cast<const float *>(w1); p.ldb = F;
    p.lsu_gather = x_rows ? 1 : 0;
    p.passes = passes;
    p.pool.w2 = w2; p.pool.b2 = b2; p.pool.tix = tix; p.pool.e = e; p.pool.zsum = zsum; p.pool.pooled = pooled; p.pool.F4 = F / 4;
    const int half = 2 * TBM * TBK * 4;
    p.stages = SMEM_DATA / (half * (p.passes == 3 ? 2 : 1));
    if (p.stages > MAX_STAGES) p.stages = MAX_STAGES;
    p.tiles_m = cdiv(n_rows, 256); p.tiles_n = 1;
    CUtensorMap mapA, mapB;
    bool ok;
    if (elt == 2) {
        ok = make_map_bf16(&mapB, w1, F, A, F, 128);
        ok = ok && (x_rows ? make_map_bf16(&mapA, w1, F, A, F, 128) : make_map_bf16(&mapA, x, F, n_rows, ldx, TBM));
    } else {
        ok = make_map(&mapA, reinterpret_cast<const float *>(x), F, x_rows ? 0x7fffffffLL : n_rows, ldx, x_rows ? 1 : TBM, false) &&
             make_map(&mapB, reinterpret_cast<const float *>(w1), F, A, F, 128, false);
    }
    CUtensorMap mapA2 = mapA, mapB2 = mapB;
    if (ok && planes) {
        ok = make_map_bf16(&mapB2, w1_lo, F, A, F, 128);
        ok = ok && (x_rows ? make_map_bf16(&mapA2, w1_lo, F, A, F, 128) : make_map_bf16(&mapA2, x_lo, F, n_rows, ldx, TBM));
    }
    if (!ok) return fail(XNRS_ERR_UNSUPPORTED, "%s: tensor map encoding failed", "xnrs_titlepool_fwd");
    CUtensorMap mapC;               // hid (n_rows x A, fp32) leaves through TMA stores
    p.c_tma = c_tma_map(&mapC, reinterpret_cast<const float *>(hid), n_rows, A, A, nullptr, XNRS_ACT_NONE, p.c_bf16 != 0);
    if (!p.c_tma) mapC = mapB;
    const int smem_bytes = SMEM_DATA + SMEM_EPI + 1024;
    static bool attr_set = false;
    if (!attr_set) {
        if (cudaFuncSetAttribute(gemm_tc2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes) != cudaSuccess) {
            cudaGetLastError();
            return fail(XNRS_ERR_UNSUPPORTED, "%s: cannot reserve shared memory", "xnrs_titlepool_fwd");
        }
        attr_set = true;
    }
    // where the per-title weighted sums are formed.  In the GEMM epilogue (x rows re-read from L2 while the next tile's MMAs
    // run): fewer HBM bytes, but the epilogue warps then sit on L2 latency; or by titlepool_wsum_kernel afterwards (a warp per
    // title, rows re-read from HBM / L2 at stream speed).  Measured at the bench shapes: bf16 storage 0.41 (fused) vs 0.24 ms
    // (split); fp32 storage / 3xTF32 0.496 vs 0.477 ms since hid leaves through TMA stores (before: fused 0.51, split 0.55).
    // The split is the default wherever it applies; XNRS_TITLEPOOL_SPLIT = 0 / 1 forces either.
    static int split_opt = -2;
    if (split_opt == -2) { const char *ev = getenv("XNRS_TITLEPOOL_SPLIT"); split_opt = ev ? atoi(ev) : -1; }
    const bool split = seg && F <= 768 && (split_opt != 0 || planes);
    const long long pairs = std::min<long long>(p.tiles_m, num_sms() / 2);
    if (split) {
        p.pool.pooled = nullptr;
        p.pool.zsum = nullptr;
        launch_pdl(gemm_tc2_kernel<true>, dim3((unsigned)(2 * pairs)), dim3(TC2_THREADS), smem_bytes, st, mapA, mapB, mapC, mapA2, mapB2, p);
        XNRS_LAUNCHED();
        const unsigned grid = (unsigned)std::max<long long>(1, std::min<long long>(cdiv(R, 8), 16LL * num_sms()));
        if (planes) titlepool_wsum_kernel<false><<<grid, 256, 0, st>>>(x_sum, ld_sum, x_rows, seg, e, R, n_rows, F, attn, pooled);
        else if (elt == 2) titlepool_wsum_kernel<true><<<grid, 256, 0, st>>>(x, ldx, x_rows, seg, e, R, n_rows, F, attn, pooled);
        else titlepool_wsum_kernel<false><<<grid, 256, 0, st>>>(x, ldx, x_rows, seg, e, R, n_rows, F, attn, pooled);
        XNRS_LAUNCHED();
        return XNRS_OK;
    }
    if (cudaMemsetAsync(pooled, 0, (size_t)R * F * sizeof(float), st) != cudaSuccess ||
        cudaMemsetAsync(zsum, 0, (size_t)R * sizeof(float), st) != cudaSuccess)
        return fail(XNRS_ERR_CUDA, "%s: memset failed", "xnrs_titlepool_fwd");
    launch_pdl(gemm_tc2_kernel<true>, dim3((unsigned)(2 * pairs)), dim3(TC2_THREADS), smem_bytes, st, mapA, mapB, mapC, mapA2, mapB2, p);
    XNRS_LAUNCHED();
    const long long work = std::max<long long>(n_rows, R * (F / 4));
    long long blocks = cdiv(work, 256), cap = 8LL * num_sms();
    titlepool_finalize_kernel<<<(unsigned)std::max<long long>(1, std::min(blocks, cap)), 256, 0, st>>>(e, tix, zsum, n_rows, R, F / 4, attn, pooled);
    XNRS_LAUNCHED();
    return XNRS_OK;
}

extern "C" int xnrs_titlepool_fwd(const float *x, long long ldx, const int *x_rows, const int *tix, const int *seg, long long n_rows,
                                  long long R, int F, int A, const float *w1, const float *b1, const float *w2, const float *b2,
                                  int precision, float *hid, float *e, float *zsum, float *attn, float *pooled,
                                  xnrs_stream_t st) {
    if (precision != XNRS_PREC_TF32X3 && precision != XNRS_PREC_TF32)
        return fail(XNRS_ERR_UNSUPPORTED, "%s: fp32-storage fused pooling runs in the TF32X3 / TF32 precisions", "xnrs_titlepool_fwd");
    return titlepool_fwd_impl(x, ldx, x_rows, tix, seg, n_rows, R, F, A, w1, b1, w2, b2, precision == XNRS_PREC_TF32X3 ? 3 : 1, 4,
                              hid, e, zsum, attn, pooled, STREAM(st));
}

extern "C" int xnrs_titlepool_fwd_bf16(const void *x, long long ldx, const int *x_rows, const int *tix, const int *seg,
                                       long long n_rows, long long R, int F, int A, const void *w1, const float *b1,
                                       const float *w2, const float *b2, void *hid, float *e, float *zsum, float *attn,
                                       float *pooled, xnrs_stream_t st) {
    return titlepool_fwd_impl(x, ldx, x_rows, tix, seg, n_rows, R, F, A, w1, b1, w2, b2, 1, 2, hid, e, zsum, attn, pooled, STREAM(st));
}

extern "C" int xnrs_titlepool_fwd_bf16x3(const void *x_hi, const void *x_lo, long long ldx, const int *x_rows, const int *tix,
                                         const int *seg, long long n_rows, long long R, int F, int A, const void *w1_hi,
                                         const void *w1_lo, int planes_fp16, const float *b1, const float *w2, const float *b2,
                                         const float *x_f32, long long ld_f32, float *hid, float *e, float *zsum, float *attn,
                                         float *pooled, xnrs_stream_t st) {
    XNRS_REQUIRE(x_lo && w1_lo && x_f32, "null pointer");
    return titlepool_fwd_impl(x_hi, ldx, x_rows, tix, seg, n_rows, R, F, A, w1_hi, b1, w2, b2, 3, 2, hid, e, zsum, attn, pooled,
                              STREAM(st), x_lo, w1_lo, x_f32, ld_f32, planes_fp16);
}

// C[M,N] (=|+=) act(opA(A) opB(B) + bias) with bf16 operands (tcgen05 kind::f16, fp32 accumulation in TMEM) on the CTA-pair
// kernel.  C is fp32 (split-K / accumulate allowed) or bf16 (c_bf16: plain store).  a_rows / b_rows gather table rows with
// the cp.async producer warp (K-major A / MN-major B, as in the fp32 kernel).
static int gemm_bf16_impl(int transA, int transB, long long M, long long N, long long K, const void *A, long long lda,
                         const int *a_rows, const void *B, long long ldb, const int *b_rows, void *C, long long ldc, int c_bf16,
                         const float *bias, int act, int accumulate, int split_k, xnrs_stream_t st_, const void *A_lo,
                         const void *B_lo, int a_f16 = 0, int b_f16 = 0) {
    const bool planes = A_lo != nullptr;
    XNRS_REQUIRE(M > 0 && N > 0 && K > 0 && A && B && C, "bad arguments");
    XNRS_REQUIRE(!planes || (B_lo && !c_bf16 && !((uintptr_t)A_lo & 15) && !((uintptr_t)B_lo & 15)), "pre-split planes");
    XNRS_REQUIRE(act == XNRS_ACT_NONE || act == XNRS_ACT_RELU || act == XNRS_ACT_TANH, "activation");
    XNRS_REQUIRE(!(c_bf16 && (accumulate || split_k > 1)), "a bf16 C cannot accumulate / split K");
    cudaStream_t st = STREAM(st_);
    static int is_sm100 = -1;
    if (is_sm100 < 0) is_sm100 = xnrs_device_is_sm100();
    if (!is_sm100 || num_sms() % 2 || lda % 8 || ldb % 8 || ((uintptr_t)A & 15) || ((uintptr_t)B & 15) || (a_rows && transA) ||
        (b_rows && transB) || (c_bf16 ? ldc % 4 : 0) || M > 2000000000LL || N > 2000000000LL || K > 2000000000LL)
        return fail(XNRS_ERR_UNSUPPORTED, "%s: shape / alignment / device not covered by the bf16 tensor-core kernel", "xnrs_gemm_bf16");
    TcArgs p;
    memset(&p, 0, sizeof(p));
    p.M = M; p.N = N; p.K = K;
    p.C = reinterpret_cast<float *>(C); p.ldc = ldc; p.bias = bias; p.act = act; p.aux = nullptr; p.accumulate = accumulate;
    p.elt = 2; p.c_bf16 = c_bf16;
    p.a_mn = transA ? 1 : 0; p.b_mn = transB ? 0 : 1;
    p.a_gather = a_rows; p.b_gather = b_rows;
    p.A = reinterpret_cast<const float *>(A); p.lda = lda; p.B = reinterpret_cast<const float *>(B); p.ldb = ldb;
    p.lsu_gather = a_rows ? 1 : (b_rows ? 2 : 0);
    p.pf_stages = gather_prefetch_stages();
    XNRS_REQUIRE(!(a_rows && b_rows), "one gathered operand at a time");
    p.passes = planes ? 3 : 1;                      // pre-split planes: hi hi + (lo hi + hi lo), a hi and a lo half per stage
    p.planes = planes ? 2 : 0;
    p.a_f16 = a_f16; p.b_f16 = b_f16;
    p.A2 = reinterpret_cast<const float *>(A_lo); p.B2 = reinterpret_cast<const float *>(B_lo);
    p.stages = planes ? MAX_STAGES / 2 : MAX_STAGES;
    p.tiles_m = cdiv(M, 256); p.tiles_n = cdiv(N, 256);
    const long long tiles = p.tiles_m * p.tiles_n, npairs = num_sms() / 2;
    long long split = split_k;
    if (split <= 0) {
        split = tiles >= npairs ? 1 : npairs / tiles;
        split = std::max(1LL, std::min(split, cdiv(K, 1024)));
        if (act != XNRS_ACT_NONE || c_bf16) split = 1;
    }
    XNRS_REQUIRE(split == 1 || act == XNRS_ACT_NONE, "split_k with activation");
    p.k_per_split = cdiv(cdiv(K, split), 64) * 64;
    split = cdiv(K, p.k_per_split);                 // no empty trailing splits (see gemm_tensorcore)
    p.split_k = (int)split;
    CUtensorMap mapA, mapB;
    bool ok = true;
    if (!a_rows) ok = p.a_mn ? make_map_bf16(&mapA, A, M, K, lda, 64) : make_map_bf16(&mapA, A, K, M, lda, TBM);
    if (ok && !b_rows) ok = p.b_mn ? make_map_bf16(&mapB, B, N, K, ldb, 64) : make_map_bf16(&mapB, B, K, N, ldb, 128);
    if (a_rows) mapA = mapB;
    if (b_rows) mapB = mapA;
    CUtensorMap mapA2 = mapA, mapB2 = mapB;
    if (ok && planes) {
        if (!a_rows) ok = p.a_mn ? make_map_bf16(&mapA2, A_lo, M, K, lda, 64) : make_map_bf16(&mapA2, A_lo, K, M, lda, TBM);
        if (ok && !b_rows) ok = p.b_mn ? make_map_bf16(&mapB2, B_lo, N, K, ldb, 64) : make_map_bf16(&mapB2, B_lo, K, N, ldb, 128);
        if (a_rows) mapA2 = mapB2;
        if (b_rows) mapB2 = mapA2;
    }
    if (!ok) return fail(XNRS_ERR_UNSUPPORTED, "%s: tensor map encoding failed", "xnrs_gemm_bf16");
    CUtensorMap mapC;
    p.c_tma = c_tma_map(&mapC, reinterpret_cast<const float *>(C), M, N, ldc, bias, act, c_bf16 != 0);
    if (!p.c_tma) mapC = mapB;
    if (split > 1 && !accumulate) {
        if (cudaMemset2DAsync(C, ldc * sizeof(float), 0, N * sizeof(float), M, st) != cudaSuccess)
            return fail(XNRS_ERR_CUDA, "%s: memset2d failed", "xnrs_gemm_bf16");
    }
    const int smem_bytes = SMEM_DATA + SMEM_EPI + 1024;
    static bool attr_set = false;
    if (!attr_set) {
        if (cudaFuncSetAttribute(gemm_tc2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes) != cudaSuccess) {
            cudaGetLastError();
            return fail(XNRS_ERR_UNSUPPORTED, "%s: cannot reserve shared memory", "xnrs_gemm_bf16");
        }
        attr_set = true;
    }
    const long long total = tiles * split, pairs = total < npairs ? total : npairs;
    launch_pdl(gemm_tc2_kernel<false>, dim3((unsigned)(2 * pairs)), dim3(TC2_THREADS), smem_bytes, st, mapA, mapB, mapC, mapA2, mapB2, p);
    g_last_gemm_kernel = planes ? "gemm_tc2_kernel (cta_group::2, 256x256 pair tile, 3xBF16 pre-split planes, kind::f16)"
                                : "gemm_tc2_kernel (cta_group::2, 256x256 pair tile, BF16 kind::f16)";
    XNRS_LAUNCHED();
    return XNRS_OK;
}

extern "C" int xnrs_gemm_bf16(int transA, int transB, long long M, long long N, long long K, const void *A, long long lda,
                              const int *a_rows, const void *B, long long ldb, const int *b_rows, void *C, long long ldc,
                              int c_bf16, const float *bias, int act, int accumulate, int split_k, xnrs_stream_t st) {
    return gemm_bf16_impl(transA, transB, M, N, K, A, lda, a_rows, B, ldb, b_rows, C, ldc, c_bf16, bias, act, accumulate, split_k, st,
                          nullptr, nullptr);
}

extern "C" int xnrs_gemm_bf16x3(int transA, int transB, long long M, long long N, long long K, const void *A_hi, const void *A_lo,
                                int a_fp16, long long lda, const int *a_rows, const void *B_hi, const void *B_lo, int b_fp16,
                                long long ldb, const int *b_rows, float *C, long long ldc, const float *bias, int act,
                                int accumulate, int split_k, xnrs_stream_t st) {
    XNRS_REQUIRE(A_lo && B_lo, "null pointer");
    XNRS_REQUIRE((a_fp16 != 0) == (b_fp16 != 0), "both operands fp16 planes or both bf16 planes (kind::f16 rejects a mixed pair)");
    return gemm_bf16_impl(transA, transB, M, N, K, A_hi, lda, a_rows, B_hi, ldb, b_rows, C, ldc, 0, bias, act, accumulate, split_k, st,
                          A_lo, B_lo, a_fp16 != 0, b_fp16 != 0);
}

extern "C" int xnrs_set_option(const char *name, int value) {
    if (name && !strcmp(name, "gemm_2cta")) {
        xnrs::g_opt_2cta = value < 0 ? -1 : (value ? 1 : 0);
        return XNRS_OK;
    }
    return xnrs::fail(XNRS_ERR_ARG, "%s: unknown option", "xnrs_set_option");
}

extern "C" int xnrs_debug_gemm_trace(long long *buf) {
    g_gemm_trace = buf;
    return XNRS_OK;
}
