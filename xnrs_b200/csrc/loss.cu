// Rows S, L-mse, L-bce, L-nll, L-cl: the dot scorer fused with the trainer losses (value and gradient in
// one pass over the (B,N,T) candidate vectors), and the row-wise part of the supervised InfoNCE loss
// (training.py:433-472).  All HBM/latency-bound; one CTA per impression / per anchor row.
#include "common.cuh"

namespace xnrs {

constexpr int SL_THREADS = 128;

// one CTA per impression b: scores -> loss term -> d_u, d_c
__global__ void __launch_bounds__(SL_THREADS)
score_loss_kernel(const float *__restrict__ u, const float *__restrict__ c, const float *__restrict__ targets,
                  const float *__restrict__ weights, int kind, long long B, int N, int T, float gscale,
                  float *__restrict__ scores, float *__restrict__ preds, float *__restrict__ loss,
                  float *__restrict__ d_u, float *__restrict__ d_c) {
    extern __shared__ float sm[];
    float *s = sm;          // [N] raw scores
    float *g = sm + N;      // [N] d loss / d score
    __shared__ float red[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
    const long long b = blockIdx.x;
    const int T4 = T >> 2;
    const float4 *u4 = reinterpret_cast<const float4 *>(u) + b * T4;
    const float4 *c4 = reinterpret_cast<const float4 *>(c) + b * N * (long long)T4;
    if (!u) {           // scores were computed upstream: c is (B,N) raw scores, d_c their gradient
        for (int n = tid; n < N; n += blockDim.x) s[n] = c[b * N + n];
    } else {
        for (int n = warp; n < N; n += nw) {
            float acc = 0.f;
            for (int i = lane; i < T4; i += 32) {
                const float4 a = c4[(long long)n * T4 + i], w = u4[i];
                acc = fmaf(a.x, w.x, acc); acc = fmaf(a.y, w.y, acc); acc = fmaf(a.z, w.z, acc); acc = fmaf(a.w, w.w, acc);
            }
            acc = warp_sum(acc);
            if (lane == 0) s[n] = acc;
        }
    }
    __syncthreads();
    const float inv_cnt = 1.f / ((float)B * (float)N);
    float part = 0.f;
    if (kind == XNRS_LOSS_NLL) {
        // -log(e^p / (e^p + sum_neg e^n)), candidate 0 is the positive (utils.py:117-131); mean over B
        float den = 0.f;
        for (int n = 0; n < N; ++n) den += expf(s[n]);
        const float ep = expf(s[0]);
        for (int n = tid; n < N; n += blockDim.x) {
            g[n] = (expf(s[n]) / den - (n == 0 ? 1.f : 0.f)) / (float)B * gscale;
            if (preds) preds[b * N + n] = s[n];
            scores[b * N + n] = s[n];
        }
        if (tid == 0) part = -logf(ep / den) / (float)B;
    } else {
        for (int n = tid; n < N; n += blockDim.x) {
            const float sv = s[n], t = targets[b * N + n], w = weights ? weights[b * N + n] : 1.f;
            float l, gr, p;
            if (kind == XNRS_LOSS_MSE_RELU) {
                p = fmaxf(sv, 0.f);
                l = (p - t) * (p - t);
                gr = sv > 0.f ? 2.f * (p - t) : 0.f;
            } else if (kind == XNRS_LOSS_BCE_SIGMOID) {
                // nn.BCELoss()(sigmoid(s), t) (training.py:324-331): each log term is clamped at -100 like torch does, and a
                // clamped term passes no gradient
                p = 1.f / (1.f + expf(-sv));
                const float lp = logf(p), lq = logf(1.f - p);
                l = -(t * fmaxf(lp, -100.f) + (1.f - t) * fmaxf(lq, -100.f));
                const float dp = -(lp > -100.f ? t / p : 0.f) + (lq > -100.f ? (1.f - t) / (1.f - p) : 0.f);
                gr = dp * p * (1.f - p);
            } else {   // BCE with logits
                p = sv;
                l = fmaxf(sv, 0.f) - sv * t + log1pf(expf(-fabsf(sv)));
                gr = 1.f / (1.f + expf(-sv)) - t;
            }
            part += l * w * inv_cnt;
            g[n] = gr * w * inv_cnt * gscale;
            scores[b * N + n] = sv;
            if (preds) preds[b * N + n] = p;
        }
    }
    part = block_sum(part, red);
    if (tid == 0 && loss) atomicAdd(loss, part);
    __syncthreads();
    if (!u) {
        if (d_c)
            for (int n = tid; n < N; n += blockDim.x) d_c[b * N + n] = g[n];
        return;
    }
    if (!d_u) return;
    float4 *du4 = reinterpret_cast<float4 *>(d_u) + b * T4;
    float4 *dc4 = reinterpret_cast<float4 *>(d_c) + b * N * (long long)T4;
    for (int i = tid; i < T4; i += blockDim.x) {
        const float4 w = u4[i];
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int n = 0; n < N; ++n) {
            const float gn = g[n];
            const float4 a = c4[(long long)n * T4 + i];
            acc.x = fmaf(gn, a.x, acc.x); acc.y = fmaf(gn, a.y, acc.y);
            acc.z = fmaf(gn, a.z, acc.z); acc.w = fmaf(gn, a.w, acc.w);
            dc4[(long long)n * T4 + i] = make_float4(gn * w.x, gn * w.y, gn * w.z, gn * w.w);
        }
        du4[i] = acc;
    }
}

__global__ void dot_score_kernel(const float *__restrict__ u, const float *__restrict__ c, long long BN, int N, int T,
                                 float *__restrict__ scores) {
    const int lane = threadIdx.x & 31;
    long long w = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const long long nw = ((long long)gridDim.x * blockDim.x) >> 5;
    const int T4 = T >> 2;
    for (; w < BN; w += nw) {
        const float4 *c4 = reinterpret_cast<const float4 *>(c) + w * T4;
        const float4 *u4 = reinterpret_cast<const float4 *>(u) + (w / N) * T4;
        float acc = 0.f;
        for (int i = lane; i < T4; i += 32) {
            const float4 a = c4[i], x = u4[i];
            acc = fmaf(a.x, x.x, acc); acc = fmaf(a.y, x.y, acc); acc = fmaf(a.z, x.z, acc); acc = fmaf(a.w, x.w, acc);
        }
        acc = warp_sum(acc);
        if (lane == 0) scores[w] = acc;
    }
}

__global__ void dot_score_bwd_kernel(const float *__restrict__ u, const float *__restrict__ c,
                                     const float *__restrict__ d_s, int N, int T, float *__restrict__ d_u,
                                     float *__restrict__ d_c) {
    const long long b = blockIdx.x;
    for (int i = threadIdx.x; i < T; i += blockDim.x) {
        const float w = u[b * T + i];
        float acc = 0.f;
        for (int n = 0; n < N; ++n) {
            const float g = d_s[b * N + n];
            acc = fmaf(g, c[(b * N + n) * T + i], acc);
            d_c[(b * N + n) * T + i] = g * w;
        }
        d_u[b * T + i] = acc;
    }
}

// ---- InfoNCE ------------------------------------------------------------------------------------

__global__ void infonce_normalize_kernel(const float *__restrict__ emb, long long Bk, int E, float *__restrict__ ehat,
                                         float *__restrict__ inv_norm) {
    const int lane = threadIdx.x & 31;
    long long w = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const long long nw = ((long long)gridDim.x * blockDim.x) >> 5;
    for (; w < Bk; w += nw) {
        float ss = 0.f;
        for (int i = lane; i < E; i += 32) { const float v = emb[w * E + i]; ss = fmaf(v, v, ss); }
        ss = warp_sum(ss);
        const float inv = 1.f / fmaxf(sqrtf(ss), 1e-12f);     // F.normalize(eps=1e-12)
        for (int i = lane; i < E; i += 32) ehat[w * E + i] = emb[w * E + i] * inv;
        if (lane == 0) inv_norm[w] = inv;
    }
}

// one CTA per anchor row.  In: sim row (cosine similarities).  Out: un-normalised gradient row
//   G_ij = (1/t) e^{s_ij/t} ( off_ij / (den_i + 1e-12) - pos_ij / num_i )      (0 if the anchor has no positive)
__global__ void infonce_rows_kernel(float *__restrict__ sim, const int *__restrict__ labels, long long Ba, long long Bk,
                                    long long row0, float inv_t, float *__restrict__ stats) {
    __shared__ float red[32];
    const long long a = blockIdx.x, gi = row0 + a;
    float *row = sim + a * Bk;
    const int lab = labels[gi];
    float num = 0.f, den = 0.f, npos = 0.f;
    for (long long j = threadIdx.x; j < Bk; j += blockDim.x) {
        if (j == gi) continue;
        const float e = expf(row[j] * inv_t);
        den += e;
        if (labels[j] == lab) { num += e; npos += 1.f; }
    }
    num = block_sum(num, red);
    den = block_sum(den, red);
    npos = block_sum(npos, red);
    const bool has = npos > 0.f;
    if (threadIdx.x == 0 && has) {
        atomicAdd(stats, -logf(num / (den + 1e-12f)));
        atomicAdd(stats + 1, 1.f);
    }
    const float inv_den = 1.f / (den + 1e-12f), inv_num = has ? 1.f / num : 0.f;
    for (long long j = threadIdx.x; j < Bk; j += blockDim.x) {
        float g = 0.f;
        if (has && j != gi) {
            const float e = expf(row[j] * inv_t);
            g = inv_t * e * (inv_den - (labels[j] == lab ? inv_num : 0.f));
        }
        row[j] = g;
    }
}

// count[0] = number of rows i of the whole (gathered) batch that have another row with the same label: the InfoNCE
// normaliser of the GLOBAL batch.  It depends on the labels only, and every rank holds all Bk labels after the all-gather, so
// data-parallel ranks compute it locally instead of all-reducing their local counts.  O(Bk^2) compares spread over Bk / 32
// CTAs: lane = anchor (32 per CTA), each of the 8 warps scans one eighth of the labels through its own shared-memory tile.
// work: 2 zeroed words (partial sum, ticket); the last CTA to finish publishes the total (integers < 2^24: exact and order
// independent in fp32).
__global__ void __launch_bounds__(256)
infonce_count_kernel(const int *__restrict__ labels, long long Bk, float *__restrict__ work, float *__restrict__ count) {
    __shared__ int tile[8][256];
    __shared__ int has_s[8][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long i = blockIdx.x * 32LL + lane;
    const int lab = i < Bk ? labels[i] : 0;
    const long long per = (Bk + 7) / 8, j0 = warp * per, j1 = min(Bk, j0 + per);
    bool has = false;
    for (long long base = j0; base < j1; base += 256) {
        const int n = (int)min(256LL, j1 - base);
        __syncwarp();
        for (int t = lane; t < n; t += 32) tile[warp][t] = labels[base + t];
        __syncwarp();
        for (int j = 0; j < n; ++j) has |= (tile[warp][j] == lab) & (base + j != i);
    }
    has_s[warp][lane] = has;
    __syncthreads();
    if (warp == 0) {
        int any = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) any |= has_s[w][lane];
        const unsigned m = __ballot_sync(0xffffffffu, any && i < Bk);
        if (lane == 0) {
            atomicAdd(work, (float)__popc(m));
            __threadfence();
            const unsigned ticket = atomicAdd(reinterpret_cast<unsigned *>(work + 1), 1u);
            if (ticket == gridDim.x - 1) {
                __threadfence();
                count[0] = atomicAdd(work, 0.f);
            }
        }
    }
}

__global__ void infonce_finalize_kernel(const float *__restrict__ stats, float *__restrict__ loss) {
    loss[0] = stats[0] / (stats[1] + 1e-8f);
}

// d_emb = scale/(count+1e-8) * inv_norm * (d_ehat - ehat <ehat, d_ehat>)
__global__ void infonce_normalize_bwd_kernel(const float *__restrict__ d_ehat, const float *__restrict__ ehat,
                                             const float *__restrict__ inv_norm, const float *__restrict__ stats,
                                             float gscale, long long Bk, int E, float *__restrict__ d_emb) {
    const int lane = threadIdx.x & 31;
    long long w = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const long long nw = ((long long)gridDim.x * blockDim.x) >> 5;
    const float sc = stats ? gscale / (stats[1] + 1e-8f) : gscale;
    for (; w < Bk; w += nw) {
        float dot = 0.f;
        for (int i = lane; i < E; i += 32) dot = fmaf(ehat[w * E + i], d_ehat[w * E + i], dot);
        dot = warp_sum(dot);
        const float inv = inv_norm[w];
        const bool clamped = inv >= 1e12f;      // ||e|| below eps: normalisation is a constant scale
        for (int i = lane; i < E; i += 32) {
            const float d = d_ehat[w * E + i];
            d_emb[w * E + i] = sc * inv * (clamped ? d : d - ehat[w * E + i] * dot);
        }
    }
}

}  // namespace xnrs

using namespace xnrs;

extern "C" int xnrs_score_loss(const float *u, const float *c, const float *targets, const float *weights, int kind,
                               long long B, int N, int T, float grad_scale, float *scores, float *preds, float *loss,
                               float *d_u, float *d_c, xnrs_stream_t st) {
    XNRS_REQUIRE(B >= 0 && N > 0 && (!u || (T > 0 && T % 4 == 0)), "bad sizes (T % 4 == 0)");
    XNRS_REQUIRE(kind >= 0 && kind <= 3, "bad loss kind");
    XNRS_REQUIRE(loss && scores, "null pointer");
    XNRS_REQUIRE(!u || (d_u == nullptr) == (d_c == nullptr), "d_u and d_c go together");
    cudaMemsetAsync(loss, 0, sizeof(float), STREAM(st));
    if (B == 0) return XNRS_OK;
    XNRS_REQUIRE(c && (targets || kind == XNRS_LOSS_NLL), "null pointer");
    XNRS_REQUIRE(B < 2147483647LL && (size_t)2 * N * sizeof(float) <= 200 * 1024, "impression too large");
    size_t smem = (size_t)2 * N * sizeof(float);
    if (smem > 48 * 1024)
        cudaFuncSetAttribute(score_loss_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    score_loss_kernel<<<(unsigned)B, SL_THREADS, smem, STREAM(st)>>>(u, c, targets, weights, kind, B, N, T, grad_scale,
                                                                 scores, preds, loss, d_u, d_c);
    XNRS_LAUNCHED();
    return XNRS_OK;
}

extern "C" int xnrs_dot_score(const float *u, const float *c, long long B, int N, int T, float *scores,
                              xnrs_stream_t st) {
    XNRS_REQUIRE(B >= 0 && N > 0 && T > 0 && T % 4 == 0, "bad sizes (T % 4 == 0)");
    if (B == 0) return XNRS_OK;
    XNRS_REQUIRE(u && c && scores, "null pointer");
    long long warps = B * N, blocks = cdiv(warps, 8), cap = 8LL * num_sms();
    dot_score_kernel<<<(unsigned)(blocks > cap ? cap : blocks), 256, 0, STREAM(st)>>>(u, c, B * N, N, T, scores);
    XNRS_LAUNCHED();
    return XNRS_OK;
}

extern "C" int xnrs_dot_score_bwd(const float *u, const float *c, const float *d_s, long long B, int N, int T,
                                  float *d_u, float *d_c, xnrs_stream_t st) {
    XNRS_REQUIRE(B >= 0 && N > 0 && T > 0 && B < 2147483647LL, "bad sizes");
    if (B == 0) return XNRS_OK;
    XNRS_REQUIRE(u && c && d_s && d_u && d_c, "null pointer");
    dot_score_bwd_kernel<<<(unsigned)B, 128, 0, STREAM(st)>>>(u, c, d_s, N, T, d_u, d_c);
    XNRS_LAUNCHED();
    return XNRS_OK;
}

extern "C" int xnrs_infonce_normalize(const float *emb, long long Bk, int E, float *ehat, float *inv_norm,
                                      xnrs_stream_t st) {
    XNRS_REQUIRE(Bk >= 0 && E > 0, "bad sizes");
    if (Bk == 0) return XNRS_OK;
    XNRS_REQUIRE(emb && ehat && inv_norm, "null pointer");
    infonce_normalize_kernel<<<(unsigned)cdiv(Bk, 8), 256, 0, STREAM(st)>>>(emb, Bk, E, ehat, inv_norm);
    XNRS_LAUNCHED();
    return XNRS_OK;
}

extern "C" int xnrs_infonce_rows(float *sim, const int *labels, long long Ba, long long Bk, long long row0,
                                 float temperature, float *stats, xnrs_stream_t st) {
    XNRS_REQUIRE(Ba >= 0 && Bk >= Ba && row0 >= 0 && row0 + Ba <= Bk && temperature > 0.f, "bad sizes");
    if (Ba == 0) return XNRS_OK;
    XNRS_REQUIRE(sim && labels && stats, "null pointer");
    infonce_rows_kernel<<<(unsigned)Ba, 256, 0, STREAM(st)>>>(sim, labels, Ba, Bk, row0, 1.f / temperature, stats);
    XNRS_LAUNCHED();
    return XNRS_OK;
}

extern "C" int xnrs_infonce_count(const int *labels, long long Bk, float *work, float *count, xnrs_stream_t st) {
    XNRS_REQUIRE(Bk >= 0 && Bk < (1LL << 24), "bad sizes");
    XNRS_REQUIRE(work && count && (Bk == 0 || labels), "null pointer");
    infonce_count_kernel<<<(unsigned)max(1LL, cdiv(Bk, 32)), 256, 0, STREAM(st)>>>(labels, Bk, work, count);
    XNRS_LAUNCHED();
    return XNRS_OK;
}

extern "C" int xnrs_infonce_finalize(const float *stats, float *loss, xnrs_stream_t st) {
    XNRS_REQUIRE(stats && loss, "null pointer");
    infonce_finalize_kernel<<<1, 1, 0, STREAM(st)>>>(stats, loss);
    XNRS_LAUNCHED();
    return XNRS_OK;
}

extern "C" int xnrs_infonce_normalize_bwd(const float *d_ehat, const float *ehat, const float *inv_norm,
                                          const float *stats, float grad_scale, long long Bk, int E, float *d_emb,
                                          xnrs_stream_t st) {
    XNRS_REQUIRE(Bk >= 0 && E > 0, "bad sizes");
    if (Bk == 0) return XNRS_OK;
    XNRS_REQUIRE(d_ehat && ehat && inv_norm && d_emb, "null pointer");          // stats may be NULL: plain normalisation backward
    infonce_normalize_bwd_kernel<<<(unsigned)cdiv(Bk, 8), 256, 0, STREAM(st)>>>(d_ehat, ehat, inv_norm, stats, grad_scale,
                                                                           Bk, E, d_emb);
    XNRS_LAUNCHED();
    return XNRS_OK;
}
