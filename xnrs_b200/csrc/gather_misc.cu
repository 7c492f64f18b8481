// Row G (gather / scatter of table rows) and the small HBM-bound elementwise kernels of the path:
// column sums (bias gradients), axpby, relu, Adam.  All are bandwidth kernels: 128-bit accesses,
// grids sized in multiples of the SM count, grid-stride loops.
#include "common.cuh"

namespace xnrs {

thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};

// one thread per (title, token): token id lookup + mask
__global__ void expand_titles_kernel(const int *__restrict__ title_tokens, long long n_news, int S,
                                     const int *__restrict__ news_ids, long long R, int *__restrict__ rows,
                                     float *__restrict__ mask) {
    long long n = R * S;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        long long r = i / S;
        int s = (int)(i - r * S);
        int nid = news_ids[r];
        int tok = (nid >= 0 && nid < n_news) ? title_tokens[(long long)nid * S + s] : 0;
        rows[i] = tok;
        if (mask) mask[i] = tok != 0 ? 1.f : 0.f;
    }
}

// one warp per output row, float4 lanes: a pure copy, bit exact
__global__ void gather_rows_kernel(const float *__restrict__ table, long long V, int D4, const int *__restrict__ rows,
                                   long long R, float *__restrict__ out, long long ld_out4) {
    int lane = threadIdx.x & 31;
    long long w = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    long long nw = ((long long)gridDim.x * blockDim.x) >> 5;
    const float4 *t4 = reinterpret_cast<const float4 *>(table);
    float4 *o4 = reinterpret_cast<float4 *>(out);
    for (long long r = w; r < R; r += nw) {
        long long src = rows[r];
        bool ok = src >= 0 && src < V;
        for (int c = lane; c < D4; c += 32) {
            float4 v = ok ? ldg_stream(t4 + src * D4 + c) : make_float4(0.f, 0.f, 0.f, 0.f);
            o4[r * ld_out4 + c] = v;
        }
    }
}

__global__ void scatter_add_rows_kernel(float *__restrict__ dtable, long long V, int D, const int *__restrict__ rows,
                                        long long R, const float *__restrict__ dout, long long ld, int skip_row, bool vec) {
    int lane = threadIdx.x & 31;
    long long w = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    long long nw = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long r = w; r < R; r += nw) {
        long long dst = rows[r];
        if (dst < 0 || dst >= V || dst == skip_row) continue;
        if (vec) {          // 16-byte vector reductions (red.global.add.v4.f32): a quarter of the atomic transactions
            const float4 *src = reinterpret_cast<const float4 *>(dout + r * ld);
            float4 *d4 = reinterpret_cast<float4 *>(dtable + dst * D);
            for (int c = lane; c < (D >> 2); c += 32) atomicAdd(d4 + c, src[c]);
        } else {
            for (int c = lane; c < D; c += 32) atomicAdd(dtable + dst * D + c, dout[r * ld + c]);
        }
    }
}

// out[n] += sum_m X[m,n]; block = 32 x 8 threads covers 32 columns, grid.y splits the rows
__global__ void colsum_kernel(const float *__restrict__ X, long long M, long long N, long long ldx,
                              float *__restrict__ out, long long rows_per_block) {
    __shared__ float red[8][33];
    int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    long long n = blockIdx.x * 32LL + tx;
    long long m0 = blockIdx.y * rows_per_block, m1 = min(M, m0 + rows_per_block);
    float s = 0.f;
    if (n < N)
        for (long long m = m0 + ty; m < m1; m += 8) s += X[m * ldx + n];
    red[ty][tx] = s;
    __syncthreads();
    if (ty == 0 && n < N) {
        float t = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) t += red[i][tx];
        atomicAdd(out + n, t);
    }
}

// 16-byte variant (N % 4 == 0, ldx % 4 == 0, aligned): a warp covers 128 columns of a row per load, each thread keeps four
// rows in flight.  The scalar kernel above had 4 bytes per thread per row in flight and reached 54 % of the HBM peak.
__global__ void __launch_bounds__(256)
colsum4_kernel(const float4 *__restrict__ X4, long long M, long long N4, long long ldx4, float *__restrict__ out,
               long long rows_per_block) {
    __shared__ float4 red[8][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const long long n4 = blockIdx.x * 32LL + tx;
    const long long m0 = blockIdx.y * rows_per_block, m1 = min(M, m0 + rows_per_block);
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    if (n4 < N4) {
        long long m = m0 + ty;
        for (; m + 24 < m1; m += 32) {
            const float4 a = X4[m * ldx4 + n4], b = X4[(m + 8) * ldx4 + n4], c = X4[(m + 16) * ldx4 + n4],
                         d = X4[(m + 24) * ldx4 + n4];
            s.x += (a.x + b.x) + (c.x + d.x); s.y += (a.y + b.y) + (c.y + d.y);
            s.z += (a.z + b.z) + (c.z + d.z); s.w += (a.w + b.w) + (c.w + d.w);
        }
        for (; m < m1; m += 8) {
            const float4 a = X4[m * ldx4 + n4];
            s.x += a.x; s.y += a.y; s.z += a.z; s.w += a.w;
        }
    }
    red[ty][tx] = s;
    __syncthreads();
    if (ty == 0 && n4 < N4) {
        float4 t = red[0][tx];
#pragma unroll
        for (int i = 1; i < 8; ++i) { t.x += red[i][tx].x; t.y += red[i][tx].y; t.z += red[i][tx].z; t.w += red[i][tx].w; }
        atomicAdd(out + 4 * n4, t.x); atomicAdd(out + 4 * n4 + 1, t.y);
        atomicAdd(out + 4 * n4 + 2, t.z); atomicAdd(out + 4 * n4 + 3, t.w);
    }
}

__global__ void axpby_kernel(long long n, float a, const float *__restrict__ a_dev, const float *__restrict__ x,
                             float b, float *__restrict__ y) {
    float aa = a_dev ? a * (*a_dev) : a;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        y[i] = (b == 0.f) ? aa * x[i] : aa * x[i] + b * y[i];
}

__global__ void relu_kernel(long long n, const float *__restrict__ x, float *__restrict__ y) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        y[i] = fmaxf(x[i], 0.f);
}

__global__ void relu_bwd_kernel(long long n, const float *__restrict__ y, const float *__restrict__ dy,
                                float *__restrict__ dx) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        dx[i] = y[i] > 0.f ? dy[i] : 0.f;
}

__global__ void sigmoid_kernel(long long n, const float *__restrict__ x, float *__restrict__ y) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        y[i] = 1.f / (1.f + expf(-x[i]));
}

__global__ void sigmoid_bwd_kernel(long long n, const float *__restrict__ y, const float *__restrict__ dy,
                                   float *__restrict__ dx) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        dx[i] = dy[i] * y[i] * (1.f - y[i]);
}

__global__ void tanh_bwd_kernel(long long n, const float *__restrict__ y, const float *__restrict__ dy, float *__restrict__ dx) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        dx[i] = dy[i] * (1.f - y[i] * y[i]);
}

__global__ void add_scalar_kernel(long long n, const float *__restrict__ x, const float *__restrict__ b, float *__restrict__ y) {
    const float bv = b[0];
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        y[i] = x[i] + bv;
}

__global__ void transpose_kernel(const float *__restrict__ in, long long rows, long long cols, float *__restrict__ out) {
    __shared__ float tile[32][33];
    long long c0 = blockIdx.x * 32LL, r0 = blockIdx.y * 32LL;
    for (int i = threadIdx.y; i < 32; i += 8) {
        long long r = r0 + i, c = c0 + threadIdx.x;
        if (r < rows && c < cols) tile[i][threadIdx.x] = in[r * cols + c];
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += 8) {
        long long c = c0 + i, r = r0 + threadIdx.x;
        if (r < rows && c < cols) out[c * rows + r] = tile[threadIdx.x][i];
    }
}

__global__ void dropout_kernel(long long n, const float *__restrict__ x, const float *__restrict__ keep, float p,
                               unsigned long long seed, float *__restrict__ y) {
    const float inv = 1.f / (1.f - p);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        float kf;
        if (keep) {
            kf = keep[i];
        } else {
            unsigned long long z = seed + (unsigned long long)(i + 1) * 0x9E3779B97F4A7C15ull;
            z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
            z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
            z ^= z >> 31;
            kf = ((float)(z >> 40) * (1.0f / 16777216.0f)) >= p ? 1.f : 0.f;
        }
        y[i] = x[i] * kf * inv;
    }
}

// torch.optim.Adam (defaults: no weight decay, no amsgrad): one pass, 4 reads + 3 writes per element
__global__ void adam_kernel(float *__restrict__ p, const float *__restrict__ g, float *__restrict__ m,
                            float *__restrict__ v, long long n, float lr, float b1, float b2, float eps,
                            float inv_bc1, float inv_sqrt_bc2, const float *__restrict__ bc_dev, float gscale) {
    if (bc_dev) {
        inv_bc1 = bc_dev[0];
        inv_sqrt_bc2 = bc_dev[1];
    }
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        float gi = g[i] * gscale;
        float mi = b1 * m[i] + (1.f - b1) * gi;
        float vi = b2 * v[i] + (1.f - b2) * gi * gi;
        m[i] = mi;
        v[i] = vi;
        float denom = sqrtf(vi) * inv_sqrt_bc2 + eps;
        p[i] -= (lr * inv_bc1) * (mi / denom);
    }
}

// ---- Adam over the ACTIVE rows of a row-sparse table (the 700k-row user tables of LSTUR / NPA, SURVEY §7 hard part 4) -------
// torch's dense Adam moves every row every step, but a row that has never received a gradient has g = m = v = 0 and its update
// is exactly lr * 0 / (0 + eps) = 0: only rows touched at least once ("active") can change.  The optimiser keeps a bitmap and
// an append-only list of those rows; this kernel applies the SAME per-element update as adam_kernel to the listed rows only —
// bit-identical parameters to the dense pass, at (active rows / all rows) of its HBM traffic.
__global__ void mark_rows_kernel(const int *__restrict__ idx, long long n, long long V, int skip_row, unsigned *__restrict__ bitmap,
                                 int *__restrict__ active, int *__restrict__ count) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int r = idx[i];
        if (r < 0 || r >= V || r == skip_row) continue;
        const unsigned bit = 1u << (r & 31);
        if (!(atomicOr(bitmap + (r >> 5), bit) & bit)) active[atomicAdd(count, 1)] = r;      // first touch: append
    }
}

__global__ void adam_rows_kernel(float *__restrict__ p, const float *__restrict__ g, float *__restrict__ m, float *__restrict__ v,
                                 int D, const int *__restrict__ active, const int *__restrict__ count, float lr, float b1, float b2,
                                 float eps, float inv_bc1, float inv_sqrt_bc2, const float *__restrict__ bc_dev, float gscale) {
    if (bc_dev) {
        inv_bc1 = bc_dev[0];
        inv_sqrt_bc2 = bc_dev[1];
    }
    const long long n = (long long)count[0] * D;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
        const long long i = (long long)active[e / D] * D + e % D;
        float gi = g[i] * gscale;
        float mi = b1 * m[i] + (1.f - b1) * gi;
        float vi = b2 * v[i] + (1.f - b2) * gi * gi;
        m[i] = mi;
        v[i] = vi;
        float denom = sqrtf(vi) * inv_sqrt_bc2 + eps;
        p[i] -= (lr * inv_bc1) * (mi / denom);
    }
}

__global__ void zero_rows_kernel(float *__restrict__ g, int D, const int *__restrict__ active, const int *__restrict__ count) {
    const long long n = (long long)count[0] * D;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x)
        g[(long long)active[e / D] * D + e % D] = 0.f;
}

// device-side step counter for CUDA-graph replay: bc = {1/(1-b1^t), 1/sqrt(1-b2^t)}
__global__ void adam_tick_kernel(int *step, float b1, float b2, float *bc) {
    int t = ++(*step);
    bc[0] = (float)(1.0 / (1.0 - pow((double)b1, (double)t)));
    bc[1] = (float)(1.0 / sqrt(1.0 - pow((double)b2, (double)t)));
}

static inline unsigned ew_grid(long long n, int threads) {
    long long b = cdiv(n, threads);
    long long cap = 8LL * num_sms();
    return (unsigned)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace xnrs

using namespace xnrs;

extern "C" int xnrs_version(void) { return 100; }
extern "C" const char *xnrs_last_error(void) { return g_err; }
extern "C" long long xnrs_launch_count(void) { return g_launches.load(); }
extern "C" int xnrs_device_is_sm100(void) {
    int dev = 0, major = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    return major == 10;
}

extern "C" int xnrs_expand_titles(const int *title_tokens, long long n_news, int S, const int *news_ids, long long R,
                                  int *token_rows, float *mask, xnrs_stream_t st) {
    XNRS_REQUIRE(S > 0 && R >= 0 && n_news > 0, "bad sizes");
    if (R == 0) return XNRS_OK;
    XNRS_REQUIRE(title_tokens && news_ids && token_rows, "null pointer");
    expand_titles_kernel<<<ew_grid(R * S, 256), 256, 0, STREAM(st)>>>(title_tokens, n_news, S, news_ids, R, token_rows, mask);
    XNRS_LAUNCHED();
    return XNRS_OK;
}

extern "C" int xnrs_gather_rows(const float *table, long long V, int D, const int *rows, long long R, float *out,
                                long long ld_out, xnrs_stream_t st) {
    XNRS_REQUIRE(D > 0 && D % 4 == 0 && ld_out % 4 == 0 && ld_out >= D, "D and ld_out must be multiples of 4");
    if (R == 0) return XNRS_OK;
    XNRS_REQUIRE(table && rows && out, "null pointer");
    XNRS_REQUIRE(((uintptr_t)table & 15) == 0 && ((uintptr_t)out & 15) == 0, "16-byte alignment");
    gather_rows_kernel<<<ew_grid(R * 32, 256), 256, 0, STREAM(st)>>>(table, V, D / 4, rows, R, out, ld_out / 4);
    XNRS_LAUNCHED();
    return XNRS_OK;
}

extern "C" int xnrs_scatter_add_rows(float *dtable, long long V, int D, const int *rows, long long R, const float *dout,
                                     long long ld_dout, int skip_row, xnrs_stream_t st) {
    XNRS_REQUIRE(D > 0 && ld_dout >= D, "bad sizes");
    if (R == 0) return XNRS_OK;
    XNRS_REQUIRE(dtable && rows && dout, "null pointer");
    const bool vec = D % 4 == 0 && ld_dout % 4 == 0 && (((uintptr_t)dtable | (uintptr_t)dout) & 15) == 0;
    scatter_add_rows_kernel<<<ew_grid(R * 32, 256), 256, 0, STREAM(st)>>>(dtable, V, D, rows, R, dout, ld_dout, skip_row, vec);
    XNRS_LAUNCHED();
    return XNRS_OK;
}

extern "C" int xnrs_colsum(const float *X, long long M, long long N, long long ldx, float *out, xnrs_stream_t st) {
    XNRS_REQUIRE(M >= 0 && N >= 0 && ldx >= N, "bad sizes");
    if (M == 0 || N == 0) return XNRS_OK;
    XNRS_REQUIRE(X && out, "null pointer");
    if (N % 4 == 0 && ldx % 4 == 0 && (((uintptr_t)X) & 15) == 0 && M >= 256) {
        const long long nbx = cdiv(N / 4, 32);
        long long splits = cdiv(8LL * num_sms(), nbx);
        long long rpb = cdiv(M, splits < 1 ? 1 : splits);
        if (rpb < 64) rpb = 64;
        splits = cdiv(M, rpb);
        dim3 grid((unsigned)nbx, (unsigned)splits);
        colsum4_kernel<<<grid, 256, 0, STREAM(st)>>>(reinterpret_cast<const float4 *>(X), M, N / 4, ldx / 4, out, rpb);
    } else {
        long long nbx = cdiv(N, 32);
        long long want = cdiv(4LL * num_sms(), nbx);
        long long splits = want < 1 ? 1 : want;
        long long rpb = cdiv(M, splits);
        if (rpb < 64) rpb = 64;
        splits = cdiv(M, rpb);
        dim3 grid((unsigned)nbx, (unsigned)splits);
        colsum_kernel<<<grid, 256, 0, STREAM(st)>>>(X, M, N, ldx, out, rpb);
    }
    XNRS_LAUNCHED();
    return XNRS_OK;
}

extern "C" int xnrs_axpby(long long n, float a, const float *a_dev, const float *x, float b, float *y, xnrs_stream_t st) {
    if (n <= 0) return XNRS_OK;
    XNRS_REQUIRE(x && y, "null pointer");
    axpby_kernel<<<ew_grid(n, 256), 256, 0, STREAM(st)>>>(n, a, a_dev, x, b, y);
    XNRS_LAUNCHED();
    return XNRS_OK;
}

extern "C" int xnrs_relu(long long n, const float *x, float *y, xnrs_stream_t st) {
    if (n <= 0) return XNRS_OK;
    XNRS_REQUIRE(x && y, "null pointer");
    relu_kernel<<<ew_grid(n, 256), 256, 0, STREAM(st)>>>(n, x, y);
    XNRS_LAUNCHED();
    return XNRS_OK;
}

extern "C" int xnrs_adam_step(float *p, const float *g, float *m, float *v, long long n, float lr, float beta1,
                              float beta2, float eps, int step, const float *bc_dev, float grad_scale,
                              xnrs_stream_t st) {
    XNRS_REQUIRE(step >= 1 || bc_dev, "step counts from 1");
    if (step < 1) step = 1;
    if (n <= 0) return XNRS_OK;
    XNRS_REQUIRE(p && g && m && v, "null pointer");
    double bc1 = 1.0 - pow((double)beta1, step), bc2 = 1.0 - pow((double)beta2, step);
    adam_kernel<<<ew_grid(n, 256), 256, 0, STREAM(st)>>>(p, g, m, v, n, lr, beta1, beta2, eps, (float)(1.0 / bc1),
                                                    (float)(1.0 / sqrt(bc2)), bc_dev, grad_scale);
    XNRS_LAUNCHED();
    return XNRS_OK;
}

extern "C" int xnrs_mark_rows(const int *idx, long long n, long long V, int skip_row, int *bitmap, int *active, int *count,
                              xnrs_stream_t st) {
    XNRS_REQUIRE(n >= 0 && V > 0, "bad sizes");
    if (n == 0) return XNRS_OK;
    XNRS_REQUIRE(idx && bitmap && active && count, "null pointer");
    mark_rows_kernel<<<ew_grid(n, 256), 256, 0, STREAM(st)>>>(idx, n, V, skip_row, reinterpret_cast<unsigned *>(bitmap), active, count);
    XNRS_LAUNCHED();
    return XNRS_OK;
}

extern "C" int xnrs_adam_rows(float *p, const float *g, float *m, float *v, long long V, int D, const int *active, const int *count,
                              float lr, float beta1, float beta2, float eps, int step, const float *bc_dev, float grad_scale,
                              xnrs_stream_t st) {
    XNRS_REQUIRE(step >= 1 || bc_dev, "step counts from 1");
    if (step < 1) step = 1;
    XNRS_REQUIRE(V > 0 && D > 0 && p && g && m && v && active && count, "bad arguments");
    double bc1 = 1.0 - pow((double)beta1, step), bc2 = 1.0 - pow((double)beta2, step);
    // the active count lives on the device: size the grid for the machine, the kernel reads the count
    adam_rows_kernel<<<(unsigned)(8 * num_sms()), 256, 0, STREAM(st)>>>(p, g, m, v, D, active, count, lr, beta1, beta2, eps,
                                                                     (float)(1.0 / bc1), (float)(1.0 / sqrt(bc2)), bc_dev, grad_scale);
    XNRS_LAUNCHED();
    return XNRS_OK;
}

extern "C" int xnrs_zero_rows(float *g, long long V, int D, const int *active, const int *count, xnrs_stream_t st) {
    XNRS_REQUIRE(V > 0 && D > 0 && g && active && count, "bad arguments");
    zero_rows_kernel<<<(unsigned)(8 * num_sms()), 256, 0, STREAM(st)>>>(g, D, active, count);
    XNRS_LAUNCHED();
    return XNRS_OK;
}

extern "C" int xnrs_adam_tick(int *step_dev, float beta1, float beta2, float *bc_dev, xnrs_stream_t st) {
    XNRS_REQUIRE(step_dev && bc_dev, "null pointer");
    adam_tick_kernel<<<1, 1, 0, STREAM(st)>>>(step_dev, beta1, beta2, bc_dev);
    XNRS_LAUNCHED();
    return XNRS_OK;
}

extern "C" int xnrs_transpose(const float *in, long long rows, long long cols, float *out, xnrs_stream_t st) {
    XNRS_REQUIRE(rows > 0 && cols > 0 && in && out, "bad arguments");
    dim3 grid((unsigned)cdiv(cols, 32), (unsigned)cdiv(rows, 32));
    transpose_kernel<<<grid, dim3(32, 8), 0, STREAM(st)>>>(in, rows, cols, out);
    XNRS_LAUNCHED();
    return XNRS_OK;
}

extern "C" int xnrs_dropout(long long n, const float *x, const float *keep, float p, unsigned long long seed, float *y,
                            xnrs_stream_t st) {
    XNRS_REQUIRE(p >= 0.f && p < 1.f, "p in [0,1)");
    if (n <= 0) return XNRS_OK;
    XNRS_REQUIRE(x && y, "null pointer");
    dropout_kernel<<<ew_grid(n, 256), 256, 0, STREAM(st)>>>(n, x, keep, p, seed, y);
    XNRS_LAUNCHED();
    return XNRS_OK;
}

extern "C" int xnrs_sigmoid(long long n, const float *x, float *y, xnrs_stream_t st) {
    if (n <= 0) return XNRS_OK;
    XNRS_REQUIRE(x && y, "null pointer");
    sigmoid_kernel<<<ew_grid(n, 256), 256, 0, STREAM(st)>>>(n, x, y);
    XNRS_LAUNCHED();
    return XNRS_OK;
}

extern "C" int xnrs_sigmoid_bwd(long long n, const float *y, const float *dy, float *dx, xnrs_stream_t st) {
    if (n <= 0) return XNRS_OK;
    XNRS_REQUIRE(y && dy && dx, "null pointer");
    sigmoid_bwd_kernel<<<ew_grid(n, 256), 256, 0, STREAM(st)>>>(n, y, dy, dx);
    XNRS_LAUNCHED();
    return XNRS_OK;
}

extern "C" int xnrs_tanh_bwd(long long n, const float *y, const float *dy, float *dx, xnrs_stream_t st) {
    if (n <= 0) return XNRS_OK;
    XNRS_REQUIRE(y && dy && dx, "null pointer");
    tanh_bwd_kernel<<<ew_grid(n, 256), 256, 0, STREAM(st)>>>(n, y, dy, dx);
    XNRS_LAUNCHED();
    return XNRS_OK;
}

extern "C" int xnrs_add_scalar(long long n, const float *x, const float *b, float *y, xnrs_stream_t st) {
    if (n <= 0) return XNRS_OK;
    XNRS_REQUIRE(x && b && y, "null pointer");
    add_scalar_kernel<<<ew_grid(n, 256), 256, 0, STREAM(st)>>>(n, x, b, y);
    XNRS_LAUNCHED();
    return XNRS_OK;
}

extern "C" int xnrs_relu_bwd(long long n, const float *y, const float *dy, float *dx, xnrs_stream_t st) {
    if (n <= 0) return XNRS_OK;
    XNRS_REQUIRE(y && dy && dx, "null pointer");
    relu_bwd_kernel<<<ew_grid(n, 256), 256, 0, STREAM(st)>>>(n, y, dy, dx);
    XNRS_LAUNCHED();
    return XNRS_OK;
}
