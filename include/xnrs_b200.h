/*
 * xnrs_b200 — C ABI of the B200 (sm_100a) kernels behind the xnrs bi-encoder hot path.
 *
 * The reference (tan9zj/xnrs) has no FFI: its hot path is PyTorch eager code.  This header is the
 * drop-in boundary a maintainer binds instead (ctypes stub in INTEGRATION.md): every entry point
 * takes plain device pointers, sizes and a CUDA stream, returns 0 on success and a negative code on
 * error (message via xnrs_last_error()).  No torch types, no hidden streams, no host sync, no CPU
 * fallback.  All tensors are dense row-major fp32 unless stated; index tensors are int32.
 * Each group cites the reference lines (relative to the reference repo) it replaces.
 */
#ifndef XNRS_B200_H
#define XNRS_B200_H

#ifdef __cplusplus
extern "C" {
#endif

typedef void *xnrs_stream_t; /* cudaStream_t */

enum { XNRS_OK = 0, XNRS_ERR_ARG = -1, XNRS_ERR_CUDA = -2, XNRS_ERR_UNSUPPORTED = -3 };
enum { XNRS_ACT_NONE = 0, XNRS_ACT_RELU = 1, XNRS_ACT_TANH = 2, XNRS_ACT_RELU_MASK = 3 };
/* arithmetic of the GEMM-shaped ops on fp32 operands: exact fp32 FMA, 3xTF32 split (fp32-accurate, tensor cores) or
 * single-pass TF32.  XNRS_PREC_BF16 names the bf16-STORAGE mode: its GEMMs take bf16 operands through xnrs_gemm_bf16 /
 * xnrs_titlepool_fwd_bf16 (tcgen05 kind::f16); xnrs_gemm itself rejects it (fp32 operands cannot be "bf16"). */
enum { XNRS_PREC_FP32 = 0, XNRS_PREC_TF32X3 = 1, XNRS_PREC_TF32 = 2, XNRS_PREC_BF16 = 3 };
enum { XNRS_LOSS_MSE_RELU = 0, XNRS_LOSS_BCE_LOGITS = 1, XNRS_LOSS_NLL = 2, XNRS_LOSS_BCE_SIGMOID = 3 };

int xnrs_version(void);
const char *xnrs_last_error(void);
/* number of kernels this library has launched since load (bench.py's gpu_launches) */
long long xnrs_launch_count(void);
/* tuning switches: "gemm_2cta" routes wide tensor-core GEMMs to the cta_group::2 CTA-pair kernel:
   0 = never, 1 = wherever legal, -1 = default policy (the fp32-accurate 3xTF32 mode only) */
int xnrs_set_option(const char *name, int value);
/* 1 when the running device is compute capability 10.x */
int xnrs_device_is_sm100(void);

/* ---- row G: gather (xnrs/data/dataset.py:63-65,77-85,97-109; news_encoding.py:45-47) ---------- */
/* news ids (R) -> flat token-row ids (R*S) and token mask (R*S, 1.0 where token != 0) */
int xnrs_expand_titles(const int *title_tokens, long long n_news, int S, const int *news_ids, long long R,
                       int *token_rows, float *mask, xnrs_stream_t st);
/* out[r,:] = table[rows[r],:]  (bit-exact copy; D % 4 == 0) */
int xnrs_gather_rows(const float *table, long long V, int D, const int *rows, long long R, float *out,
                     long long ld_out, xnrs_stream_t st);
/* dtable[rows[r],:] += dout[r,:]  (nn.Embedding backward: lstur.py:94-98, npa.py:12-15, naml.py:34-47);
 * rows equal to skip_row (padding_idx, or -1 for none) receive nothing */
int xnrs_scatter_add_rows(float *dtable, long long V, int D, const int *rows, long long R, const float *dout,
                          long long ld_dout, int skip_row, xnrs_stream_t st);

/* device-side id plumbing of one encoder pass (replaces the reference's per-sample host lookups, dataset.py:63-65,77-85;
 * no host round trip: the counts stay in `counts`, int32[2] = {U distinct articles, T real tokens}).
 * xnrs_plan_dedup: ids (n) -> uniq (capacity n: the U distinct ids ascending, then article 0), inv (n: slot -> row of uniq),
 * counts[0] = U.  work = 2 * ceil(n_news / 32) ints of scratch (bitmap + word prefixes).  Ids outside [0, n_news) count
 * as the pad article 0. */
int xnrs_plan_dedup(const int *ids, long long n, long long n_news, int *work, int *uniq, int *inv, int *counts,
                    xnrs_stream_t st);
/* xnrs_plan_ragged: the real (non-zero) tokens of the articles uniq[0..U) (U = u_count[0], or cap when u_count is NULL), in
 * title order: rows (capacity rows_cap >= cap*S; entries [T, T+pad_rows) are set to token 0 = the zero row), seg (cap+1 group
 * offsets; groups past U are empty, seg[cap] = T), cm (cap, 1.0 where a title has tokens: xnrs/utils.py:74-75), lens (cap,
 * scratch), counts[1] = T.  tix (nullable, rows_cap): the group of each row, -1 on the padding rows. */
int xnrs_plan_ragged(const int *title_tokens, long long n_news, int S, const int *uniq, long long cap, const int *u_count,
                     int pad_rows, int *lens, int *seg, int *rows, int *tix, long long rows_cap, float *cm, int *counts,
                     xnrs_stream_t st);

/* ---- GEMM (every nn.Linear / matmul on the path: layers.py:60,94-95,128-130,154; news_encoding.py:55-56)
 * C[M,N] (=|+=) act( opA(A)[M,K] * opB(B)[K,N] + bias[N] ).  transA=0: A stored MxK (lda), 1: KxM.
 * transB=0: B stored KxN (ldb), 1: NxK (an nn.Linear weight).  a_rows / b_rows (nullable) gather the
 * STORED rows of A / B through an index (the fused table gather of row G).  act RELU_MASK multiplies by
 * (aux > 0) (ReLU backward), aux has C's layout.  split_k > 1 accumulates partial sums atomically
 * (requires accumulate semantics: C must hold the initial value; bias added once; act must be NONE). */
int xnrs_gemm(int transA, int transB, long long M, long long N, long long K, const float *A, long long lda,
              const int *a_rows, const float *B, long long ldb, const int *b_rows, float *C, long long ldc,
              const float *bias, int act, const float *aux, int accumulate, int split_k, int precision,
              xnrs_stream_t st);
/* diagnostics: while buf != NULL, every CTA of the 1-CTA tensor-core GEMM kernel writes 8 SM-clock stamps to buf[8 * cta ..]
 * (entry, set-up done, first operands ready, last MMA issued, accumulator ready, epilogue done, all roles done, TMEM freed);
 * buf must hold 8 * 148 entries.  NULL switches it off (the default).  tools/bench_step_gemms.py reads it. */
int xnrs_debug_gemm_trace(long long *buf);
/* name of the kernel the calling thread's last xnrs_gemm dispatched to ("gemm_tc2_kernel ...", "gemm_tc_kernel<128> ...",
 * "gemm_simt_kernel"), and the number of xnrs_gemm calls made in a tensor-core precision that the exact-fp32 SIMT kernel
 * took instead (shape / alignment the TMA path cannot express): correct but slow, so it is counted, never silent */
const char *xnrs_last_gemm_kernel(void);
long long xnrs_gemm_simt_fallbacks(void);
/* out[N] += column sums of X[M,N] (bias gradients) */
int xnrs_colsum(const float *X, long long M, long long N, long long ldx, float *out, xnrs_stream_t st);
/* y = a*x + b*y elementwise; a_dev (nullable) is a device scalar multiplied into a */
int xnrs_axpby(long long n, float a, const float *a_dev, const float *x, float b, float *y, xnrs_stream_t st);
int xnrs_relu(long long n, const float *x, float *y, xnrs_stream_t st);
int xnrs_relu_bwd(long long n, const float *y, const float *dy, float *dx, xnrs_stream_t st);
/* backward of a tanh activation (FCScoring, scoring.py:72-102): dx = dy * (1 - y^2), y = the forward output */
int xnrs_tanh_bwd(long long n, const float *y, const float *dy, float *dx, xnrs_stream_t st);
/* y = x + b[0] (the bias of nn.Bilinear(.., out_features=1), scoring.py:41-69) */
int xnrs_add_scalar(long long n, const float *x, const float *b, float *y, xnrs_stream_t st);
/* BCERankingTrainer's output activation (training.py:329-331) and its backward (y = the forward output) */
int xnrs_sigmoid(long long n, const float *x, float *y, xnrs_stream_t st);
int xnrs_sigmoid_bwd(long long n, const float *y, const float *dy, float *dx, xnrs_stream_t st);
/* out (cols,rows) = in (rows,cols)^T */
int xnrs_transpose(const float *in, long long rows, long long cols, float *out, xnrs_stream_t st);
/* inverted dropout y = x * keep / (1-p)  (nn.Dropout: lstur.py:112,135; layers.py:148 lives inside xnrs_mha_*).
 * keep (nullable, n floats 0/1) is an explicit mask, else a counter-based generator seeded by `seed` */
int xnrs_dropout(long long n, const float *x, const float *keep, float p, unsigned long long seed, float *y,
                 xnrs_stream_t st);

/* ---- row A: additive-attention pooling (layers.py:47-69) -------------------------------------
 * hid = tanh(fc1 x) (R*L, A) comes from xnrs_gemm(act=TANH).  x rows are x[r*L+l] or, with x_rows,
 * table rows x[x_rows[r*L+l]].  mask nullable (R*L).  attn (R*L) and pooled (R,F) are outputs. */
/* seg (nullable, R+1 ints): ragged groups — group r owns rows [seg[r], seg[r+1]) of x/hid/attn (padding tokens are
 * never materialised) and L is the longest group; with seg == NULL every group has exactly L rows */
int xnrs_addpool_fwd(const float *x, const int *x_rows, const float *mask, const float *hid, const float *w2,
                     const float *b2, const int *seg, long long R, int L, int F, int A, float *attn, float *pooled,
                     xnrs_stream_t st);
/* d_hid (R*L,A) = grad wrt the fc1 pre-activation; d_w2 (A), d_b2 (1) accumulate; d_x (nullable, R*L,F)
 * receives a_s * d_pooled (the fc1 path is added by the caller's GEMM); d_attn (nullable) is an
 * incoming gradient on the returned weights; with seg, n_rows (>= seg[R], or 0) is the length of the row buffers: rows past
 * the last group (TitlePlan padding) get d_hid = 0; d_b1 (nullable, A) accumulates the fc1 bias gradient = column sums of
 * d_hid (saves the separate xnrs_colsum pass over d_hid) */
int xnrs_addpool_bwd(const float *x, const int *x_rows, const float *mask, const float *hid, const float *w2,
                     const float *attn, const float *d_pooled, const float *d_attn, const int *seg, long long R, int L,
                     int F, int A, long long n_rows, float *d_hid, float *d_w2, float *d_b2, float *d_x, float *d_b1,
                     xnrs_stream_t st);
/* ---- rows G + A fused (north-star items 1 + 3): gather -> fc1 (+b1, tanh) -> <., w2> + b2 -> exp -> per-title sum(e) and
 * sum(e * x) -> normalise by (sum + 1e-8), in ONE launch of the CTA-pair tcgen05 kernel (cp.async gather warp, pooling
 * epilogue on the TMEM accumulators; the x rows of a tile are re-read from L2 for the weighted sum) plus a small
 * normalisation pass.  x rows are x[g] or table rows x[x_rows[g]] (ldx floats apart); tix (n_rows) = title of each row, -1
 * for padding rows; seg (R+1, nullable) = the same grouping as offsets: with it the per-title sums may be formed by a second,
 * warp-per-title kernel instead of the GEMM epilogue (chosen by measurement, XNRS_TITLEPOOL_SPLIT).  Outputs: hid (n_rows, A) = tanh(fc1 x) (saved for the backward), e (n_rows, scratch), zsum (R, scratch),
 * attn (n_rows) and pooled (R, F) exactly as xnrs_addpool_fwd defines them.  Covers A == 256, F % 128 == 0, F <= 1024,
 * n_rows >= 256 in the tensor-core precisions on sm_100; otherwise returns XNRS_ERR_UNSUPPORTED with nothing launched and the
 * caller runs xnrs_gemm(TANH) + xnrs_addpool_fwd (the same mathematics in two launches). */
int xnrs_titlepool_fwd(const float *x, long long ldx, const int *x_rows, const int *tix, const int *seg, long long n_rows,
                       long long R, int F, int A, const float *w1, const float *b1, const float *w2, const float *b2, int precision,
                       float *hid, float *e, float *zsum, float *attn, float *pooled, xnrs_stream_t st);
/* ---- XNRS_PREC_BF16: bf16 STORAGE of the token-level tensors (token table rows x, tanh hidden layer hid, its gradient) with
 * fp32 accumulation everywhere — the north-star's 2e-2 tolerance class.  bf16 operands are `void *` to 16-bit brain floats.
 * xnrs_cast_bf16 rounds fp32 to nearest-even bf16 (token table once, fc1.weight per step).
 * xnrs_gemm_bf16: C[M,N] (=|+=) act(opA(A) opB(B) + bias) on tcgen05 kind::f16 (CTA-pair kernel, fp32 TMEM accumulators);
 * same layout rules as xnrs_gemm; a_rows (K-major A) / b_rows (MN-major B) gather table rows with the cp.async warp; C is fp32
 * (accumulate / split-K allowed) or bf16 (c_bf16 = 1).  Returns XNRS_ERR_UNSUPPORTED off sm_100 or for unaligned operands.
 * xnrs_titlepool_fwd_bf16 / xnrs_addpool_bwd_bf16: the fused forward and the pooling backward on bf16 x / w1 / hid / d_hid. */
int xnrs_cast_bf16(long long n, const float *src, void *dst, xnrs_stream_t st);

/* ---- fp32-accurate arithmetic on the bf16 tensor-core path: 3xBF16 over PRE-SPLIT operands.
 * x ~ hi + lo with hi = bf16(x), lo = bf16(x - hi) (16 mantissa bits, relative error <= 2^-17); a product x y is evaluated as
 * hi hi + (lo hi + hi lo) by three kind::f16 MMAs with fp32 accumulation (the two small terms in their own TMEM accumulator,
 * added once in the epilogue) — the bf16 analogue of 3xTF32 (measured error ~1e-6 of the result scale, inside the 1e-4 bar),
 * at twice the MMA rate and half the shared-memory bytes per k, and with NO in-kernel split pass: the planes of the frozen
 * token table are made once (xnrs_split_bf16), fc1.weight is split once per step, and the pooling backward writes d_hid
 * directly as two planes (xnrs_addpool_bwd_split).  Same layouts / gathers / epilogues as xnrs_gemm_bf16; C and hid are fp32.
 * The planes of an operand are bf16, or IEEE fp16 (fp16 = 1 / a_fp16 / b_fp16 / planes_fp16: hi = fp16(x), lo = fp16(x - hi),
 * 22 mantissa bits, |x| <= 65504); both operands of a product must use the same format (the MMA rejects a mixed pair).
 * The forward product (frozen table x fc1.weight: moderate range) uses fp16 planes — representation error 2^-24, as accurate
 * as 3xTF32; the weight gradient (d_hid of arbitrary magnitude x table) uses bf16 planes of both.
 * xnrs_titlepool_fwd_bf16x3: xnrs_titlepool_fwd on planes; x_f32 (ld_f32 floats per row) = the fp32 rows for the weighted sum. */
int xnrs_split_bf16(long long n, const float *src, void *hi, void *lo, int fp16, xnrs_stream_t st);
int xnrs_gemm_bf16x3(int transA, int transB, long long M, long long N, long long K, const void *A_hi, const void *A_lo,
                     int a_fp16, long long lda, const int *a_rows, const void *B_hi, const void *B_lo, int b_fp16, long long ldb,
                     const int *b_rows, float *C, long long ldc, const float *bias, int act, int accumulate, int split_k,
                     xnrs_stream_t st);
int xnrs_titlepool_fwd_bf16x3(const void *x_hi, const void *x_lo, long long ldx, const int *x_rows, const int *tix, const int *seg,
                              long long n_rows, long long R, int F, int A, const void *w1_hi, const void *w1_lo, int planes_fp16,
                              const float *b1, const float *w2, const float *b2, const float *x_f32, long long ld_f32, float *hid,
                              float *e, float *zsum, float *attn, float *pooled, xnrs_stream_t st);
int xnrs_addpool_bwd_split(const float *x, const int *x_rows, const float *hid, const float *w2, const float *attn,
                           const float *d_pooled, const int *seg, long long R, int L, int F, int A, long long n_rows,
                           void *d_hid_hi, void *d_hid_lo, float *d_w2, float *d_b2, float *d_b1, xnrs_stream_t st);
int xnrs_gemm_bf16(int transA, int transB, long long M, long long N, long long K, const void *A, long long lda,
                   const int *a_rows, const void *B, long long ldb, const int *b_rows, void *C, long long ldc, int c_bf16,
                   const float *bias, int act, int accumulate, int split_k, xnrs_stream_t st);
int xnrs_titlepool_fwd_bf16(const void *x, long long ldx, const int *x_rows, const int *tix, const int *seg, long long n_rows,
                            long long R, int F, int A, const void *w1, const float *b1, const float *w2, const float *b2, void *hid,
                            float *e, float *zsum, float *attn, float *pooled, xnrs_stream_t st);
int xnrs_addpool_bwd_bf16(const void *x, const int *x_rows, const void *hid, const float *w2, const float *attn,
                          const float *d_pooled, const int *seg, long long R, int L, int F, int A, long long n_rows, void *d_hid,
                          float *d_w2, float *d_b2, float *d_b1, xnrs_stream_t st);
/* ---- row P: personalised attention (layers.py:88-101): logit = <tanh(x_fc x), q_fc(q)> ---------
 * hid (R*L,A) = tanh(x_fc x); qh (Rq,A) = q_fc(q); title r uses query row r / rows_per_query */
int xnrs_perspool_fwd(const float *x, const int *x_rows, const float *mask, const float *hid, const float *qh,
                      const int *seg, long long R, int L, int F, int A, int rows_per_query, float *attn, float *pooled,
                      xnrs_stream_t st);
int xnrs_perspool_bwd(const float *x, const int *x_rows, const float *mask, const float *hid, const float *qh,
                      const float *attn, const float *d_pooled, const int *seg, long long R, int L, int F, int A,
                      int rows_per_query, long long n_rows, float *d_hid, float *d_qh, float *d_x, xnrs_stream_t st);
/* out[i] = <x[i,:], w> + b[0]  (the pooler's fc2, layers.py:45,60, over a whole table of hidden rows) */
int xnrs_rowdot(const float *x, const float *w, const float *b, long long n, int A, float *out, xnrs_stream_t st);
/* additive pooling over PER-ITEM logits (evaluation with a pre-encoded catalogue; layers.py:60-65 slot by slot):
 * group r pools the table rows ids[r,0..L): a_l = exp(logit[id]) * row_mask[id] / (sum + 1e-8); pooled[r] = sum a_l table[id].
 * table (V,T), T % 4 == 0, T <= 1024; row_mask (V, nullable); attn (R*L, nullable) receives the weights */
int xnrs_logitpool_fwd(const float *table, long long V, int T, const float *logit, const float *row_mask, const int *ids,
                       long long R, int L, float *attn, float *pooled, xnrs_stream_t st);
/* backward of xnrs_logitpool_fwd: d_logit (V) and d_table (V,T) ACCUMULATE (zero them first); attn is the saved forward output */
int xnrs_logitpool_bwd(const float *table, long long V, int T, const int *ids, const float *attn, const float *d_pooled,
                       long long R, int L, float *d_logit, float *d_table, xnrs_stream_t st);
/* backward of the per-item logit <tanh-hidden row, w2> + b2: d_hid (n,A) written; d_w2 (A) and d_b2 (1) accumulate */
int xnrs_logit_bwd(const float *hid, const float *w2, const float *d_logit, long long n, int A, float *d_hid, float *d_w2,
                   float *d_b2, xnrs_stream_t st);
/* masked mean pooling (layers.py:25-37) */
int xnrs_meanpool_fwd(const float *x, const float *mask, long long R, int L, int F, float *pooled,
                      xnrs_stream_t st);
int xnrs_meanpool_bwd(const float *mask, const float *d_pooled, long long R, int L, int F, float *d_x, xnrs_stream_t st);
/* collapsed mask: clamp(sum_l m, 0, 1) (xnrs/utils.py:74-75) */
int xnrs_collapse_mask(const float *mask, long long R, int L, float *out, xnrs_stream_t st);

/* ---- row M: multi-head self-attention core (layers.py:133-151) -------------------------------
 * q,k,v,o: (R,L,h*dk) with row stride ld.  QUERY-axis mask (R*L, nullable): masked query rows attend
 * uniformly, keys are never masked.  Dropout on the normalised weights: keep (nullable, R*h*L*L 0/1)
 * is an explicit keep mask; else if p_drop > 0 a counter-based hash generator (splitmix64 of seed and the
 * (r,head,i,j) counter; NOT torch's Philox stream — SURVEY §7 hard part 3) draws the keep decisions.
 * lse (R*h*L) is saved for the backward. */
int xnrs_mha_fwd(const float *q, const float *k, const float *v, long long ld, const float *mask, long long R,
                 int L, int h, int dk, const float *keep, float p_drop, unsigned long long seed, float *o,
                 float *lse, xnrs_stream_t st);
int xnrs_mha_bwd(const float *q, const float *k, const float *v, const float *o, const float *d_o, long long ld,
                 const float *mask, const float *lse, long long R, int L, int h, int dk, const float *keep,
                 float p_drop, unsigned long long seed, float *dq, float *dk_, float *dv, xnrs_stream_t st);

/* ---- row U-lstur: GRU over the front-aligned history, final state at the true length
 * (lstur.py:139-153, torch.nn.GRU gate order r,z,n).  gi = x W_ih^T + b_ih (B*L,3Hd) from xnrs_gemm;
 * w_hh_t is W_hh transposed (Hd,3Hd).  lengths (B) int32.  Saves hs (B,L,Hd) = the state BEFORE each
 * step and gates (B,L,4Hd) = r,z,n,gh_n. */
int xnrs_gru_fwd(const float *gi, const float *w_hh_t, const float *b_hh, const float *h0, const int *lengths,
                 long long B, int L, int Hd, float *hs, float *gates, float *h_out, xnrs_stream_t st);
/* d_gi (B*L,3Hd), d_gh (B*L,3Hd) and d_h0 (B,Hd) out; w_hh is the untransposed (3Hd,Hd) weight */
int xnrs_gru_bwd(const float *d_h_out, const float *w_hh, const int *lengths, const float *hs,
                 const float *gates, long long B, int L, int Hd, float *d_gi, float *d_gh, float *d_h0,
                 xnrs_stream_t st);
int xnrs_lengths_from_mask(const float *mask, long long B, int L, int *lengths, xnrs_stream_t st);

/* ---- rows S + L-*: dot scoring (scoring.py:12-23) fused with the trainer losses
 * (training.py:336-337, 378-392; utils.py:117-131).  u (B,T), c (B,N,T), targets/weights (B*N).
 * Outputs: scores (B*N raw dot products), preds (B*N activated: relu for MSE, sigmoid for BCE_SIGMOID =
 * BCERankingTrainer's nn.BCELoss on sigmoid scores (training.py:324-331), raw otherwise),
 * loss (1, overwritten), and — when d_u/d_c are non-null — gradients of the loss (times grad_scale).
 * With u == NULL, c holds (B,N) scores computed upstream and d_c (B,N) receives d loss / d score. */
int xnrs_score_loss(const float *u, const float *c, const float *targets, const float *weights, int kind,
                    long long B, int N, int T, float grad_scale, float *scores, float *preds, float *loss,
                    float *d_u, float *d_c, xnrs_stream_t st);
/* standalone scorer and its backward: s[b,n] = <c[b,n], u[b]> */
int xnrs_dot_score(const float *u, const float *c, long long B, int N, int T, float *scores, xnrs_stream_t st);
int xnrs_dot_score_bwd(const float *u, const float *c, const float *d_s, long long B, int N, int T, float *d_u,
                       float *d_c, xnrs_stream_t st);

/* ---- row L-cl: supervised InfoNCE (training.py:433-472) ----------------------------------------
 * anchors = rows [row0, row0+Ba) of the Bk gathered embeddings emb (Bk,E); labels (Bk).
 * stage 1 writes normalised embeddings ehat (Bk,E) and inv_norm (Bk); the caller forms
 * sim = ehat[row0:row0+Ba] ehat^T with xnrs_gemm; stage 2 turns sim (Ba,Bk) into the un-normalised
 * gradient G in place and accumulates stats[0] += sum of anchor terms, stats[1] += anchors with positives.
 * stage 3 (after an optional all-reduce of stats) writes loss = stats[0]/(stats[1]+1e-8).
 * stage 4 maps d_ehat (Bk,E) (= G ehat_k on anchor rows + G^T ehat_a, from xnrs_gemm) to d_emb; with stats == NULL it is
 * the plain backward of the row normalisation (DotScoring(normalize=True), scoring.py:20-22) scaled by grad_scale. */
int xnrs_infonce_normalize(const float *emb, long long Bk, int E, float *ehat, float *inv_norm, xnrs_stream_t st);
int xnrs_infonce_rows(float *sim, const int *labels, long long Ba, long long Bk, long long row0,
                      float temperature, float *stats, xnrs_stream_t st);
/* data parallel: count[0] = rows of the WHOLE gathered batch with a same-label partner (the global normaliser; a function of
 * the gathered labels only, so no all-reduce of the local counts is needed).  work = 2 zeroed words of scratch. */
int xnrs_infonce_count(const int *labels, long long Bk, float *work, float *count, xnrs_stream_t st);
int xnrs_infonce_finalize(const float *stats, float *loss, xnrs_stream_t st);
int xnrs_infonce_normalize_bwd(const float *d_ehat, const float *ehat, const float *inv_norm, const float *stats,
                               float grad_scale, long long Bk, int E, float *d_emb, xnrs_stream_t st);

/* ---- row C across GPUs: the exchange steps of the global-batch InfoNCE over NVLink peer memory (one process per GPU; the
 * reference computes its InfoNCE over one process's batch: training.py:433-472; data parallelism is this repo's extension,
 * SURVEY.md 8(e)).  ptrs: device array of the W base addresses of the ranks' symmetric buffers (identical layout; offsets
 * in bytes).  ctl: 8 zero-initialised ints of device memory private to the rank (epochs / tickets / [4] = a peer never
 * arrived).  Flags: W uint32 slots per kind at off_flags in every buffer, zero-initialised.
 * xnrs_peer_normalize_allgather: ehat = emb / max(|emb|, 1e-12) of this rank's Ba rows, stored with 1/norm and the labels into
 * rows [rank*Ba, (rank+1)*Ba) of EVERY rank's gathered arrays; returns (stream order) once every rank's rows have arrived here.
 * xnrs_peer_reduce_scatter_normalize_bwd: d = sum over ranks of their d_ehat rows [rank*Ba, +Ba) (pulled in rank order),
 * d_emb = grad_scale * (*grad_scale_dev or 1) / (stats[1] + 1e-8) * inv_norm * (d - ehat <ehat, d>). */
int xnrs_peer_normalize_allgather(const float *emb, const int *labels, long long Ba, int E, int rank, int world,
                                  const long long *ptrs, long long off_flags, long long off_ehat, long long off_inv,
                                  long long off_lab, int *ctl, xnrs_stream_t st);
int xnrs_peer_reduce_scatter_normalize_bwd(const long long *ptrs, long long off_flags, long long off_dehat, long long Ba, int E,
                                           int rank, int world, const float *ehat_a, const float *inv_norm_a,
                                           const float *stats, float grad_scale, const float *grad_scale_dev, float *d_emb,
                                           int *ctl, xnrs_stream_t st);

/* ---- rows E + Me: per-impression scoring and ranking metrics (training.py:194-227; metrics.py:7-44)
 * CSR impressions: candidates of impression i are cand_ids[offsets[i]:offsets[i+1]].  score =
 * act(<user[i], news_vecs[cand]>) (act: 0 raw, 1 relu, 2 sigmoid), then nan_to_num(nan 0, +inf 1, -inf 0).
 * news_vecs has n_news rows; a candidate id outside [0, n_news) is scored as article 0 (the pad article), never read out of
 * bounds.  If user == NULL, `scores_io` holds raw scores computed upstream; `act` is applied to them in place.  metrics_out (n_imp,6) doubles:
 * auc, rr, ndcg@5, ndcg@10, ctr@1, ctr@10.  Tie order: descending score, then descending index. */
int xnrs_eval_impressions(const float *user, const float *news_vecs, long long n_news, int T, const int *cand_ids,
                          const long long *offsets, const float *targets, long long n_imp, int act,
                          float *scores_io, double *metrics_out, xnrs_stream_t st);
/* sums[0..5] += column sums over impressions with a finite auc, sums[6] += their count */
int xnrs_metric_sums(const double *metrics, long long n_imp, double *sums, xnrs_stream_t st);
/* thresholded metrics of _test_step (training.py:219-222; metrics.py:47-64) per CSR impression: out (n_imp,7) float64 =
 * accuracy, recall, precision, tn, fp, fn, tp with prediction = round(clip(nan_to_num(score), 0, 1)) = (score > 0.5) */
int xnrs_binary_metrics(const float *scores, const float *targets, const long long *offsets, long long n_imp, double *out,
                        xnrs_stream_t st);

/* ---- row Opt: Adam, torch defaults (training.py:39), one launch over a flat parameter buffer ---- */
/* bias corrections come from the host `step` (>= 1) or, when bc_dev is non-null, from the device pair
 * {1/(1-b1^t), 1/sqrt(1-b2^t)} that xnrs_adam_tick maintains (CUDA-graph replay keeps counting) */
int xnrs_adam_step(float *p, const float *g, float *m, float *v, long long n, float lr, float beta1, float beta2,
                   float eps, int step, const float *bc_dev, float grad_scale, xnrs_stream_t st);
int xnrs_adam_tick(int *step_dev, float beta1, float beta2, float *bc_dev, xnrs_stream_t st);
/* Adam over the ACTIVE rows of a row-sparse table (nn.Embedding(n_users + 1, ..): lstur.py:94-98, npa.py:12-15).  A row that
 * never received a gradient has g = m = v = 0 and dense Adam leaves it exactly unchanged, so the optimiser tracks the rows
 * touched at least once: xnrs_mark_rows appends first-time rows of idx (n) to `active` (V ints) through `bitmap` (ceil(V/32)
 * ints, zero-initialised) and bumps count[0]; rows equal to skip_row (padding_idx) or outside [0, V) are ignored.
 * xnrs_adam_rows applies xnrs_adam_step's update to the active rows only (bit-identical to the dense pass over the table);
 * xnrs_zero_rows clears the gradient of the active rows (the rest of the table's gradient is never written). */
int xnrs_mark_rows(const int *idx, long long n, long long V, int skip_row, int *bitmap, int *active, int *count,
                   xnrs_stream_t st);
int xnrs_adam_rows(float *p, const float *g, float *m, float *v, long long V, int D, const int *active, const int *count,
                   float lr, float beta1, float beta2, float eps, int step, const float *bc_dev, float grad_scale,
                   xnrs_stream_t st);
int xnrs_zero_rows(float *g, long long V, int D, const int *active, const int *count, xnrs_stream_t st);

#ifdef __cplusplus
}
#endif
#endif /* XNRS_B200_H */
