// Row G, index side: the id plumbing of one encoder pass, on the device and without a host round trip.
//
// The reference looks each news id up on the host, one sample at a time (xnrs/data/dataset.py:63-65,77-85,97-109).  Here a
// batch is int32 news ids; these kernels turn the ids of one pass into
//   * the DISTINCT articles of the batch (uniq, ascending like torch.unique) and the slot -> article map (inv): every title is
//     encoded independently of its position (news_encoding.py:48-57), so each distinct article is encoded once;
//   * the ragged token rows of those articles: pad tokens have pooling weight exactly 0 (layers.py:62-64), so only the real
//     tokens are listed (rows), with group offsets (seg) and the collapsed title mask (cm, xnrs/utils.py:74-75).
// The counts (U distinct articles, T real tokens) stay on the device; the arrays are PADDED past them with harmless entries
// (article 0 / token 0 = the zero rows, empty groups), so consumers can run on bucketed upper bounds and a CUDA graph can
// replay the step without knowing the exact counts.  Small integer work: bitmap + single-CTA prefix scans (<= 160k entries).
#include "common.cuh"

namespace xnrs {

constexpr int SCAN_T = 1024;

// exclusive prefix of v over the block (SCAN_T threads); *total (all threads) = block sum.  `sh` >= 33 ints
__device__ __forceinline__ int block_excl_scan(int v, int *sh, int *total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    __syncthreads();                    // previous use of sh is over
    if (lane == 31) sh[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        int w = sh[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, w, o);
            if (lane >= o) w += t;
        }
        sh[lane] = w;                   // inclusive warp totals
    }
    __syncthreads();
    const int before = warp ? sh[warp - 1] : 0;
    *total = sh[31];
    return before + inc - v;
}

// work = [bits: W words, one bit per catalogue article | wpos: W words, number of distinct articles before each word]
__global__ void plan_mark_kernel(const int *__restrict__ ids, long long n, long long n_news, unsigned *__restrict__ bits) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        int id = ids[i];
        if (id < 0 || id >= n_news) id = 0;                 // out-of-catalogue ids read the pad article, like the gathers do
        atomicOr(bits + (id >> 5), 1u << (id & 31));
    }
}

// ONE CTA: per-word popcounts -> exclusive prefix (wpos); the set bits of each word are emitted in ascending order:
// uniq[wpos + rank] = article; counts[0] = U.  W = n_news / 32 words: 2-5 k for MIND -> a handful of block scans.
__global__ void __launch_bounds__(SCAN_T)
plan_scan_bits_kernel(const unsigned *__restrict__ bits, long long W, int *__restrict__ wpos, int *__restrict__ uniq,
                      int *__restrict__ counts) {
    __shared__ int sh[33];
    int carry = 0;
    for (long long base = 0; base < W; base += SCAN_T) {
        const long long i = base + threadIdx.x;
        unsigned b = i < W ? bits[i] : 0u;
        int tot;
        int pos = carry + block_excl_scan(__popc(b), sh, &tot);
        if (i < W) {
            wpos[i] = pos;
            while (b) {
                const int bit = __ffs(b) - 1;
                uniq[pos++] = (int)(i * 32 + bit);
                b &= b - 1;
            }
        }
        carry += tot;
    }
    if (threadIdx.x == 0) counts[0] = carry;
}

// inv[slot] = position of the slot's article among the distinct ones; uniq is padded with article 0 up to its capacity
__global__ void plan_inv_kernel(const int *__restrict__ ids, long long n, long long n_news, const unsigned *__restrict__ bits,
                                const int *__restrict__ wpos, const int *__restrict__ counts, int *__restrict__ inv,
                                int *__restrict__ uniq) {
    const int U = counts[0];
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        int id = ids[i];
        if (id < 0 || id >= n_news) id = 0;
        inv[i] = wpos[id >> 5] + __popc(bits[id >> 5] & ((1u << (id & 31)) - 1u));
        if (i >= U) uniq[i] = 0;
    }
}

// lens[u] = number of real (non-zero) tokens of article uniq[u] for u < U, 0 past U; cm[u] = the collapsed title mask
__global__ void plan_lens_kernel(const int *__restrict__ title_tokens, long long n_news, int S, const int *__restrict__ uniq,
                                 long long cap, const int *__restrict__ u_count, int *__restrict__ lens, float *__restrict__ cm) {
    const long long U = u_count ? (long long)u_count[0] : cap;
    for (long long u = blockIdx.x * (long long)blockDim.x + threadIdx.x; u < cap; u += (long long)gridDim.x * blockDim.x) {
        int len = 0;
        if (u < U) {
            const int id = uniq[u];
            if (id >= 0 && id < n_news) {
                const int *t = title_tokens + (long long)id * S;
                for (int s = 0; s < S; ++s) len += t[s] != 0;
            }
        }
        lens[u] = len;
        cm[u] = len > 0 ? 1.f : 0.f;
    }
}

// ONE CTA: seg = exclusive prefix of lens (scanned up to U, the groups past U are empty: seg = T), seg[cap] = T; counts[1] = T
__global__ void __launch_bounds__(SCAN_T)
plan_scan_lens_kernel(const int *__restrict__ lens, long long cap, const int *__restrict__ u_count, int *__restrict__ seg,
                      int *__restrict__ counts) {
    __shared__ int sh[33];
    const long long U = u_count ? min((long long)u_count[0], cap) : cap;
    int carry = 0;
    long long base = 0;
    for (; base < U; base += SCAN_T) {
        const long long i = base + threadIdx.x;
        const int v = i < U ? lens[i] : 0;
        int tot;
        const int pos = carry + block_excl_scan(v, sh, &tot);
        if (i < cap) seg[i] = pos;
        carry += tot;
    }
    for (long long i = base + threadIdx.x; i <= cap; i += SCAN_T) seg[i] = carry;
    if (threadIdx.x == 0) {
        seg[cap] = carry;
        counts[1] = carry;
    }
}

// one warp per article: its real tokens, in title order, to rows[seg[u] ...); then rows [T, T + pad) = token 0 (the zero row)
__global__ void plan_rows_kernel(const int *__restrict__ title_tokens, long long n_news, int S, const int *__restrict__ uniq,
                                 long long cap, const int *__restrict__ seg, long long rows_cap, int pad, int *__restrict__ rows,
                                 int *__restrict__ tix) {
    const int lane = threadIdx.x & 31;
    const long long w = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5, nw = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long u = w; u < cap; u += nw) {
        const int beg = seg[u], len = seg[u + 1] - beg;
        if (len == 0) continue;
        const int *t = title_tokens + (long long)uniq[u] * S;
        int out = beg;
        for (int s0 = 0; s0 < S; s0 += 32) {
            const int s = s0 + lane;
            const int tok = s < S ? t[s] : 0;
            const unsigned m = __ballot_sync(0xffffffffu, tok != 0);
            if (tok != 0) {
                const int o = out + __popc(m & ((1u << lane) - 1u));
                rows[o] = tok;
                if (tix) tix[o] = (int)u;
            }
            out += __popc(m);
        }
    }
    const long long T = seg[cap];
    for (long long i = T + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < T + pad && i < rows_cap;
         i += (long long)gridDim.x * blockDim.x)
    {
        rows[i] = 0;
        if (tix) tix[i] = -1;
    }
}

static unsigned ew_blocks(long long n, int threads) {
    long long b = cdiv(n, threads), cap = 8LL * num_sms();
    return (unsigned)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace xnrs

using namespace xnrs;

extern "C" int xnrs_plan_dedup(const int *ids, long long n, long long n_news, int *work, int *uniq, int *inv, int *counts,
                               xnrs_stream_t st) {
    XNRS_REQUIRE(n >= 0 && n_news > 0 && n_news < 2147483647LL, "bad sizes");
    XNRS_REQUIRE(work && uniq && inv && counts, "null pointer");
    cudaStream_t s = STREAM(st);
    const long long W = cdiv(n_news, 32);               // work (2 W ints) holds the W-word bitmap and the W word prefixes
    unsigned *bits = reinterpret_cast<unsigned *>(work);
    int *wpos = work + W;
    if (cudaMemsetAsync(work, 0, (size_t)W * sizeof(int), s) != cudaSuccess)
        return fail(XNRS_ERR_CUDA, "%s: memset failed", "xnrs_plan_dedup");
    if (n > 0) {
        plan_mark_kernel<<<ew_blocks(n, 256), 256, 0, s>>>(ids, n, n_news, bits);
        XNRS_LAUNCHED();
    }
    plan_scan_bits_kernel<<<1, SCAN_T, 0, s>>>(bits, W, wpos, uniq, counts);
    XNRS_LAUNCHED();
    if (n > 0) {
        plan_inv_kernel<<<ew_blocks(n, 256), 256, 0, s>>>(ids, n, n_news, bits, wpos, counts, inv, uniq);
        XNRS_LAUNCHED();
    }
    return XNRS_OK;
}

extern "C" int xnrs_plan_ragged(const int *title_tokens, long long n_news, int S, const int *uniq, long long cap,
                                const int *u_count, int pad_rows, int *lens, int *seg, int *rows, int *tix, long long rows_cap,
                                float *cm, int *counts, xnrs_stream_t st) {
    XNRS_REQUIRE(cap > 0 && n_news > 0 && S > 0 && pad_rows >= 0, "bad sizes");
    XNRS_REQUIRE(title_tokens && uniq && lens && seg && rows && cm && counts, "null pointer");
    XNRS_REQUIRE(rows_cap >= cap * S, "rows buffer smaller than cap * S");
    cudaStream_t s = STREAM(st);
    plan_lens_kernel<<<ew_blocks(cap, 256), 256, 0, s>>>(title_tokens, n_news, S, uniq, cap, u_count, lens, cm);
    XNRS_LAUNCHED();
    plan_scan_lens_kernel<<<1, SCAN_T, 0, s>>>(lens, cap, u_count, seg, counts);
    XNRS_LAUNCHED();
    plan_rows_kernel<<<ew_blocks(cap * 32, 256), 256, 0, s>>>(title_tokens, n_news, S, uniq, cap, seg, rows_cap, pad_rows, rows, tix);
    XNRS_LAUNCHED();
    return XNRS_OK;
}
