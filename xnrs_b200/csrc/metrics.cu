// Rows E + Me: per-impression candidate scoring and ranking metrics over CSR impressions
// (training.py:194-227; evaluation/metrics.py:7-44).  One CTA per impression:
//   1. score_c = act(<user_i, news_vecs[cand_c]>)  (a warp per candidate, float4 loads), nan_to_num;
//   2. a segmented enumeration (rank) sort inside the CTA: rank_c = #{c' : s_c' > s_c or (s_c' == s_c
//      and c' > c)} — i.e. descending score, ties by descending index == np.argsort(kind='stable')[::-1];
//      impressions are short (mean ~37 candidates) so the O(n^2/threads) rank sort beats a bitonic network;
//   3. AUC (Mann-Whitney, ties 1/2), reciprocal rank, nDCG@5/10, CTR@1/10 in float64 like numpy.
#include <math.h>

#include "common.cuh"

namespace xnrs {

constexpr int MT = 128;          // threads per impression
constexpr int MCAP = 2048;       // candidates staged in shared memory; longer impressions read global memory

__device__ __forceinline__ float nan_to_num_f(float s) {
    if (isnan(s)) return 0.f;
    if (isinf(s)) return s > 0.f ? 1.f : 0.f;
    return s;
}

__device__ __forceinline__ double block_sum_d(double v, double *red) {
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    v = warp_sum_d(v);
    __syncthreads();
    if (lane == 0) red[w] = v;
    __syncthreads();
    double t = (lane < nw) ? red[lane] : 0.0;
    return warp_sum_d(t);
}

__global__ void __launch_bounds__(MT)
eval_impressions_kernel(const float *__restrict__ user, const float *__restrict__ news_vecs, int T,
                        const int *__restrict__ cand_ids, const long long *__restrict__ offsets,
                        const float *__restrict__ targets, long long n_imp, int act, float *__restrict__ scores_io,
                        double *__restrict__ metrics_out, int min_n, long long n_news) {
    __shared__ float s_sc[MCAP];
    __shared__ float s_tg[MCAP];
    __shared__ double red[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
    for (long long imp = blockIdx.x; imp < n_imp; imp += gridDim.x) {
        const long long beg = offsets[imp];
        const int n = (int)(offsets[imp + 1] - beg);
        if (n < min_n) continue;                    // short impressions belong to the warp-per-impression kernel
        float *gs = scores_io + beg;
        const float *gt = targets + beg;
        if (user) {
            const int T4 = T >> 2;
            const float4 *u4 = reinterpret_cast<const float4 *>(user) + imp * T4;
            for (int c = warp; c < n; c += nw) {
                long long cid = cand_ids[beg + c];
                if (cid < 0 || cid >= n_news) cid = 0;        // an id outside the catalogue scores like the pad article (row 0)
                const float4 *v4 = reinterpret_cast<const float4 *>(news_vecs) + cid * T4;
                float acc = 0.f;
                for (int i = lane; i < T4; i += 32) {
                    const float4 a = v4[i], b = u4[i];
                    acc = fmaf(a.x, b.x, acc); acc = fmaf(a.y, b.y, acc);
                    acc = fmaf(a.z, b.z, acc); acc = fmaf(a.w, b.w, acc);
                }
                acc = warp_sum(acc);
                if (lane == 0) {
                    if (act == 1) acc = fmaxf(acc, 0.f);
                    else if (act == 2) acc = 1.f / (1.f + expf(-acc));
                    gs[c] = acc;
                }
            }
            __syncthreads();
        } else if (act != 0) {       // scores computed upstream: the trainer's output activation still applies (BCE: sigmoid)
            for (int c = tid; c < n; c += blockDim.x) {
                float v = gs[c];
                if (act == 1) v = fmaxf(v, 0.f);
                else v = 1.f / (1.f + expf(-v));
                gs[c] = v;
            }
            __syncthreads();
        }
        const bool in_smem = n <= MCAP;
        if (in_smem) {
            for (int c = tid; c < n; c += blockDim.x) { s_sc[c] = nan_to_num_f(gs[c]); s_tg[c] = gt[c]; }
            __syncthreads();
        }
        auto SC = [&](int c) -> float { return in_smem ? s_sc[c] : nan_to_num_f(gs[c]); };
        auto TG = [&](int c) -> float { return in_smem ? s_tg[c] : gt[c]; };

        double dcg5 = 0, dcg10 = 0, idcg5 = 0, idcg10 = 0, ctr1 = 0, ctr10 = 0, rr = 0, auc_num = 0, npos = 0, nneg = 0;
        for (int c = tid; c < n; c += blockDim.x) {
            const float sc = SC(c), tc = TG(c);
            int rank = 0, trank = 0;
            double wins = 0.0;
            const bool pos = tc > 0.5f;
            for (int j = 0; j < n; ++j) {
                const float sj = SC(j), tj = TG(j);
                rank += (sj > sc) || (sj == sc && j > c);
                trank += (tj > tc) || (tj == tc && j > c);
                if (pos && !(tj > 0.5f)) wins += (sc > sj) ? 1.0 : (sc == sj ? 0.5 : 0.0);
            }
            const double gain = exp2((double)tc) - 1.0;
            if (rank < 5) dcg5 += gain / log2((double)rank + 2.0);
            if (rank < 10) dcg10 += gain / log2((double)rank + 2.0);
            if (trank < 5) idcg5 += gain / log2((double)trank + 2.0);
            if (trank < 10) idcg10 += gain / log2((double)trank + 2.0);
            if (rank < 1) ctr1 += tc;
            if (rank < 10) ctr10 += tc;
            rr = fmax(rr, (double)tc / ((double)rank + 1.0));
            auc_num += wins;
            if (pos) npos += 1.0; else nneg += 1.0;
        }
        dcg5 = block_sum_d(dcg5, red); dcg10 = block_sum_d(dcg10, red);
        idcg5 = block_sum_d(idcg5, red); idcg10 = block_sum_d(idcg10, red);
        ctr1 = block_sum_d(ctr1, red); ctr10 = block_sum_d(ctr10, red);
        auc_num = block_sum_d(auc_num, red); npos = block_sum_d(npos, red); nneg = block_sum_d(nneg, red);
        // max-reduce rr
        for (int o = 16; o > 0; o >>= 1) rr = fmax(rr, __shfl_xor_sync(0xffffffffu, rr, o));
        __syncthreads();
        if (lane == 0) red[warp] = rr;
        __syncthreads();
        if (tid == 0) {
            for (int w = 1; w < nw; ++w) rr = fmax(rr, red[w]);
            double *m = metrics_out + imp * 6;
            m[0] = (npos > 0 && nneg > 0) ? auc_num / (npos * nneg) : nan("");
            m[1] = rr;
            m[2] = dcg5 / idcg5;
            m[3] = dcg10 / idcg10;
            m[4] = n > 0 ? ctr1 / 1.0 : nan("");
            m[5] = n > 0 ? ctr10 / (double)min(n, 10) : nan("");
        }
        __syncthreads();
    }
}

// ---- warp-per-impression variant (impressions of up to WCAP candidates: all of MIND, mean ~37) -------------------------
// A 128-thread CTA spent most of its time in barriers and in 128-thread reductions over ~37 candidates (9 block-wide
// double reductions per impression).  Here one warp owns an impression: the user vector lives in registers, four
// candidates are scored per iteration (eight 16-byte loads in flight per lane), scores/targets sit in a per-warp
// shared-memory slab, ranks are counted lane-per-candidate, and all reductions are warp shuffles.  AUC uses exact
// integer win counts (2 per win, 1 per tie); the DCG discounts log2(rank + 2), rank < 10, come from a host-computed table.
constexpr int WCAP = 256;            // candidates per impression handled here; longer ones go to the CTA kernel
constexpr int EW = 4;                // warps per CTA
__constant__ double c_log2r[10];

template <int NV>       // up to NV float4 of the T-wide vectors per lane (T <= 128 * NV)
__global__ void __launch_bounds__(EW * 32)
eval_impressions_warp_kernel(const float *__restrict__ user, const float *__restrict__ news_vecs, int T4,
                             const int *__restrict__ cand_ids, const long long *__restrict__ offsets,
                             const float *__restrict__ targets, long long n_imp, int act, float *__restrict__ scores_io,
                             double *__restrict__ metrics_out, long long n_news) {
    __shared__ float s_sc_all[EW][WCAP];
    __shared__ float s_tg_all[EW][WCAP];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float *s_sc = s_sc_all[warp], *s_tg = s_tg_all[warp];
    const float4 *nv4 = reinterpret_cast<const float4 *>(news_vecs);
    for (long long imp = (long long)blockIdx.x * EW + warp; imp < n_imp; imp += (long long)gridDim.x * EW) {
        const long long beg = offsets[imp];
        const int n = (int)(offsets[imp + 1] - beg);
        if (n > WCAP) continue;
        float *gs = scores_io + beg;
        const float *gt = targets + beg;
        if (user) {
            float4 u[NV];
            const float4 *u4 = reinterpret_cast<const float4 *>(user) + imp * T4;
#pragma unroll
            for (int k = 0; k < NV; ++k) u[k] = (k * 32 + lane < T4) ? u4[k * 32 + lane] : make_float4(0.f, 0.f, 0.f, 0.f);
            for (int c0 = 0; c0 < n; c0 += 4) {
                float acc[4];
                int myid = (lane < 4 && c0 + lane < n) ? cand_ids[beg + c0 + lane] : 0;
                if (myid < 0 || myid >= n_news) myid = 0;      // out-of-catalogue id -> the pad article (row 0), never out of bounds
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int id = __shfl_sync(0xffffffffu, myid, q);
                    const float4 *row = nv4 + (long long)id * T4;
                    float a = 0.f;
#pragma unroll
                    for (int k = 0; k < NV; ++k) {
                        if (k * 32 + lane < T4) {
                            const float4 x = __ldg(row + k * 32 + lane);
                            a = fmaf(x.x, u[k].x, a); a = fmaf(x.y, u[k].y, a);
                            a = fmaf(x.z, u[k].z, a); a = fmaf(x.w, u[k].w, a);
                        }
                    }
                    acc[q] = a;
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) acc[q] += __shfl_xor_sync(0xffffffffu, acc[q], o);
                }
                if (lane < 4 && c0 + lane < n) {
                    float v = lane == 0 ? acc[0] : (lane == 1 ? acc[1] : (lane == 2 ? acc[2] : acc[3]));
                    if (act == 1) v = fmaxf(v, 0.f);
                    else if (act == 2) v = 1.f / (1.f + expf(-v));
                    gs[c0 + lane] = v;
                    s_sc[c0 + lane] = nan_to_num_f(v);
                }
            }
            for (int c = lane; c < n; c += 32) s_tg[c] = gt[c];
        } else {
            for (int c = lane; c < n; c += 32) {
                float v = gs[c];
                if (act == 1) v = fmaxf(v, 0.f);
                else if (act == 2) v = 1.f / (1.f + expf(-v));
                if (act != 0) gs[c] = v;
                s_sc[c] = nan_to_num_f(v);
                s_tg[c] = gt[c];
            }
        }
        __syncwarp();
        double dcg5 = 0, dcg10 = 0, idcg5 = 0, idcg10 = 0, ctr1 = 0, ctr10 = 0, rr = 0;
        int wins2 = 0, npos = 0, nneg = 0;
        for (int c = lane; c < n; c += 32) {
            const float sc = s_sc[c], tc = s_tg[c];
            const bool pos = tc > 0.5f;
            int rank = 0, trank = 0, w2 = 0;
            for (int j = 0; j < n; ++j) {
                const float sj = s_sc[j], tj = s_tg[j];
                rank += (sj > sc) || (sj == sc && j > c);
                trank += (tj > tc) || (tj == tc && j > c);
                if (!(tj > 0.5f)) w2 += (sc > sj) ? 2 : (sc == sj ? 1 : 0);
            }
            if (pos) { wins2 += w2; npos += 1; } else nneg += 1;
            const double gain = (tc == 0.f) ? 0.0 : exp2((double)tc) - 1.0;
            if (rank < 10) {
                const double g = gain / c_log2r[rank];
                dcg10 += g;
                ctr10 += tc;
                if (rank < 5) dcg5 += g;
                if (rank < 1) ctr1 += tc;
            }
            if (trank < 10) {
                const double g = gain / c_log2r[trank];
                idcg10 += g;
                if (trank < 5) idcg5 += g;
            }
            rr = fmax(rr, (double)tc / ((double)rank + 1.0));
        }
        dcg5 = warp_sum_d(dcg5); dcg10 = warp_sum_d(dcg10); idcg5 = warp_sum_d(idcg5); idcg10 = warp_sum_d(idcg10);
        ctr1 = warp_sum_d(ctr1); ctr10 = warp_sum_d(ctr10);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            wins2 += __shfl_xor_sync(0xffffffffu, wins2, o);
            npos += __shfl_xor_sync(0xffffffffu, npos, o);
            nneg += __shfl_xor_sync(0xffffffffu, nneg, o);
            rr = fmax(rr, __shfl_xor_sync(0xffffffffu, rr, o));
        }
        if (lane == 0) {
            double *m = metrics_out + imp * 6;
            m[0] = (npos > 0 && nneg > 0) ? (0.5 * (double)wins2) / ((double)npos * (double)nneg) : nan("");
            m[1] = rr;
            m[2] = dcg5 / idcg5;
            m[3] = dcg10 / idcg10;
            m[4] = n > 0 ? ctr1 / 1.0 : nan("");
            m[5] = n > 0 ? ctr10 / (double)min(n, 10) : nan("");
        }
        __syncwarp();
    }
}

// Thresholded per-impression metrics of _test_step (training.py:219-222; metrics.py:47-64): predictions are
// round(clip(score, 0, 1)) with numpy's round-half-to-even, i.e. 1 exactly when the (nan_to_num'd) score is > 0.5.
//   out[imp] = (accuracy, recall, precision, tn, fp, fn, tp);  recall / precision are 0 when their denominator is 0
//   (sklearn's zero_division behaviour in the reference calls).  One warp per impression.
__global__ void __launch_bounds__(256)
binary_metrics_kernel(const float *__restrict__ scores, const float *__restrict__ targets, const long long *__restrict__ offsets,
                      long long n_imp, double *__restrict__ out) {
    const int lane = threadIdx.x & 31;
    const long long imp = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (imp >= n_imp) return;
    const long long beg = offsets[imp];
    const int n = (int)(offsets[imp + 1] - beg);
    int tn = 0, fp = 0, fn = 0, tp = 0;
    for (int c = lane; c < n; c += 32) {
        const bool pred = nan_to_num_f(scores[beg + c]) > 0.5f;
        const bool pos = targets[beg + c] > 0.5f;
        tp += pred && pos; fp += pred && !pos; fn += !pred && pos; tn += !pred && !pos;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        tn += __shfl_xor_sync(0xffffffffu, tn, o); fp += __shfl_xor_sync(0xffffffffu, fp, o);
        fn += __shfl_xor_sync(0xffffffffu, fn, o); tp += __shfl_xor_sync(0xffffffffu, tp, o);
    }
    if (lane == 0) {
        double *m = out + imp * 7;
        m[0] = n > 0 ? (double)(tp + tn) / (double)n : nan("");
        m[1] = (tp + fn) > 0 ? (double)tp / (double)(tp + fn) : 0.0;
        m[2] = (tp + fp) > 0 ? (double)tp / (double)(tp + fp) : 0.0;
        m[3] = tn; m[4] = fp; m[5] = fn; m[6] = tp;
    }
}

__global__ void metric_sums_kernel(const double *__restrict__ metrics, long long n_imp, double *__restrict__ sums) {
    __shared__ double red[32];
    double acc[7] = {0, 0, 0, 0, 0, 0, 0};
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n_imp; i += (long long)gridDim.x * blockDim.x) {
        const double *m = metrics + i * 6;
        bool ok = true;
        for (int k = 0; k < 6; ++k) ok = ok && isfinite(m[k]);
        if (!ok) continue;
        for (int k = 0; k < 6; ++k) acc[k] += m[k];
        acc[6] += 1.0;
    }
    for (int k = 0; k < 7; ++k) {
        double v = block_sum_d(acc[k], red);
        if (threadIdx.x == 0) atomicAdd(sums + k, v);
    }
}

}  // namespace xnrs

using namespace xnrs;

extern "C" int xnrs_eval_impressions(const float *user, const float *news_vecs, long long n_news, int T, const int *cand_ids,
                                     const long long *offsets, const float *targets, long long n_imp, int act,
                                     float *scores_io, double *metrics_out, xnrs_stream_t st) {
    XNRS_REQUIRE(n_imp >= 0 && act >= 0 && act <= 2, "bad arguments");
    if (n_imp == 0) return XNRS_OK;
    XNRS_REQUIRE(offsets && targets && scores_io && metrics_out, "null pointer");
    if (user) XNRS_REQUIRE(news_vecs && cand_ids && T > 0 && T % 4 == 0 && n_news > 0, "scoring needs news_vecs (n_news rows), cand_ids, T % 4 == 0");
    static bool table_set = false;
    if (!table_set) {
        double h[10];
        for (int r = 0; r < 10; ++r) h[r] = log2((double)r + 2.0);
        if (cudaMemcpyToSymbol(c_log2r, h, sizeof(h)) != cudaSuccess) return fail(XNRS_ERR_CUDA, "%s: constant upload failed", "xnrs_eval_impressions");
        table_set = true;
    }
    // warp-per-impression kernel for impressions of up to WCAP candidates (vectors up to 1024 wide) ...
    const bool warp_ok = !user || T <= 1024;
    if (warp_ok) {
        const int T4 = T / 4;
        long long blocks = cdiv(n_imp, EW), cap = 32LL * num_sms();
        const unsigned grid = (unsigned)(blocks < cap ? blocks : cap);
#define XNRS_EW(NV) eval_impressions_warp_kernel<NV><<<grid, EW * 32, 0, STREAM(st)>>>(user, news_vecs, T4, cand_ids, offsets, targets, n_imp, act, scores_io, metrics_out, n_news)
        if (T4 <= 32) XNRS_EW(1);
        else if (T4 <= 64) XNRS_EW(2);
        else if (T4 <= 128) XNRS_EW(4);
        else XNRS_EW(8);
#undef XNRS_EW
        XNRS_LAUNCHED();
    }
    // ... and the CTA-per-impression kernel for the longer ones (it skips impressions the warp kernel took)
    long long cap = warp_ok ? 2LL * num_sms() : 16LL * num_sms();        // with the warp kernel this one only sweeps for long impressions
    eval_impressions_kernel<<<(unsigned)(n_imp < cap ? n_imp : cap), MT, 0, STREAM(st)>>>(
        user, news_vecs, T, cand_ids, offsets, targets, n_imp, act, scores_io, metrics_out, warp_ok ? WCAP + 1 : 0, n_news);
    XNRS_LAUNCHED();
    return XNRS_OK;
}

extern "C" int xnrs_metric_sums(const double *metrics, long long n_imp, double *sums, xnrs_stream_t st) {
    XNRS_REQUIRE(n_imp >= 0, "bad sizes");
    if (n_imp == 0) return XNRS_OK;
    XNRS_REQUIRE(metrics && sums, "null pointer");
    long long blocks = cdiv(n_imp, 256), cap = 2LL * num_sms();
    metric_sums_kernel<<<(unsigned)(blocks > cap ? cap : blocks), 256, 0, STREAM(st)>>>(metrics, n_imp, sums);
    XNRS_LAUNCHED();
    return XNRS_OK;
}

extern "C" int xnrs_binary_metrics(const float *scores, const float *targets, const long long *offsets, long long n_imp,
                                   double *out, xnrs_stream_t st) {
    XNRS_REQUIRE(n_imp >= 0, "bad sizes");
    if (n_imp == 0) return XNRS_OK;
    XNRS_REQUIRE(scores && targets && offsets && out, "null pointer");
    binary_metrics_kernel<<<(unsigned)cdiv(n_imp, 8), 256, 0, STREAM(st)>>>(scores, targets, offsets, n_imp, out);
    XNRS_LAUNCHED();
    return XNRS_OK;
}
