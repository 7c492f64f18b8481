"""Per-kernel roofline micro-benchmark on one B200: every non-GEMM kernel of the hot path at the headline shapes, timed with
CUDA events on the launching stream after warm-up, inputs larger than L2 (or L2 flushed by the 307 MB table traffic).

For each kernel: algorithmic bytes (the operands it must read / write once) / time = achieved GB/s against the measured
HBM peak in MEASURED_PEAKS.json, or algorithmic FLOPs / time for the fp32-FMA bound attention core.
    python tools/bench_kernels.py [--only mha,pool,...]  ->  one JSON line per kernel
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from xnrs_b200 import kernels as K  # noqa: E402

DEV = 'cuda'


def peaks():
    p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        return json.load(open(p)), 'measured'
    return {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0, 'bf16_tflops_sustained': 1400.0}, 'fallback'


REPS = int(os.environ.get('XNRS_BENCH_REPS', '20'))        # 1 under ncu --set full (every launch is replayed ~40x)


def timeit(fn, n=None, flush=None):
    n = n or REPS
    for _ in range(3 if REPS > 1 else 0):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(n):
        if flush is not None:
            flush.add_(1.0)                      # > L2: evicts the previous iteration's operands
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        e.synchronize()
        tot += s.elapsed_time(e)
    return tot / n


def report(name, ms, nbytes=None, flops=None, note=''):
    pk, kind = peaks()
    row = {'kernel': name, 'ms': round(ms, 4)}
    if nbytes is not None:
        gbs = nbytes / (ms * 1e-3) / 1e9
        row.update({'bound': 'hbm', 'algorithmic_MB': round(nbytes / 1e6, 1), 'achieved_GBs': round(gbs, 1),
                    'peak_GBs': pk['hbm_gbs'], 'frac': round(gbs / pk['hbm_gbs'], 3)})
    if flops is not None:
        row.update({'GFLOP': round(flops / 1e9, 2), 'achieved_TFLOPs': round(flops / (ms * 1e-3) / 1e12, 2)})
    row['peak_source'] = kind
    if note:
        row['note'] = note
    print(json.dumps(row), flush=True)


def bench_mha(flush):
    R, L, h, dk = 8748, 30, 16, 48          # distinct titles of a 1024-impression NRMS batch, S=30, 16 heads of 48
    D = h * dk
    q, k, v, do = (torch.randn(R * L, D, device=DEV) for _ in range(4))
    mask = torch.ones(R * L, device=DEV)
    o, lse = torch.empty_like(q), torch.empty(R, h, L, device=DEV)
    dq, dk_, dv = torch.empty_like(q), torch.empty_like(q), torch.empty_like(q)
    for p in (0.0, 0.1):
        ms = timeit(lambda: K.call('xnrs_mha_fwd', q, k, v, D, mask, R, L, h, dk, None, p, 1234, o, lse), flush=flush)
        report(f'xnrs_mha_fwd R={R} L={L} h={h} dk={dk} p_drop={p}', ms, nbytes=4 * q.numel() * 4,
               flops=4.0 * R * h * L * L * dk)
        ms = timeit(lambda: K.call('xnrs_mha_bwd', q, k, v, o, do, D, mask, lse, R, L, h, dk, None, p, 1234, dq, dk_, dv),
                    flush=flush)
        report(f'xnrs_mha_bwd R={R} L={L} h={h} dk={dk} p_drop={p}', ms, nbytes=8 * q.numel() * 4,
               flops=10.0 * R * h * L * L * dk)
    # user level: 1024 users x 50 clicks, 16 heads of 16
    R, L, h, dk = 1024, 50, 16, 16
    D = h * dk
    q, k, v = (torch.randn(R * L, D, device=DEV) for _ in range(3))
    mask = torch.ones(R * L, device=DEV)
    o, lse = torch.empty_like(q), torch.empty(R, h, L, device=DEV)
    ms = timeit(lambda: K.call('xnrs_mha_fwd', q, k, v, D, mask, R, L, h, dk, None, 0.1, 1234, o, lse))
    report(f'xnrs_mha_fwd R={R} L={L} h={h} dk={dk} p_drop=0.1', ms, nbytes=4 * q.numel() * 4, flops=4.0 * R * h * L * L * dk)


def bench_gather(flush):
    V, D, n = 100000, 768, 153562
    table = torch.randn(V, D, device=DEV)
    rows = torch.randint(1, V, (n,), device=DEV, dtype=torch.int32)
    out = torch.empty(n, D, device=DEV)
    ms = timeit(lambda: K.call('xnrs_gather_rows', table, V, D, rows, n, out, D))
    report(f'xnrs_gather_rows {n} rows x {D} from a {V}-row table', ms, nbytes=2 * n * D * 4 + n * 4)


def bench_pool(flush):
    R, S, F, A = 8748, 30, 768, 256
    lens = torch.randint(5, S + 1, (R,), device=DEV, dtype=torch.int32)
    seg = torch.zeros(R + 1, device=DEV, dtype=torch.int32)
    torch.cumsum(lens, 0, out=seg[1:])
    n = int(seg[-1])
    x, hid = torch.randn(n, F, device=DEV), torch.tanh(torch.randn(n, A, device=DEV))
    w2, b2 = torch.randn(A, device=DEV) * 0.05, torch.zeros(1, device=DEV)
    attn, pooled = torch.empty(n, device=DEV), torch.empty(R, F, device=DEV)
    ms = timeit(lambda: K.call('xnrs_addpool_fwd', x, None, None, hid, w2, b2, seg, R, S, F, A, attn, pooled), flush=flush)
    report(f'xnrs_addpool_fwd ragged R={R} tokens={n} F={F} A={A}', ms, nbytes=(n * F + n * A + R * F + n) * 4)
    d_pooled, d_hid = torch.randn(R, F, device=DEV), torch.empty_like(hid)
    d_w2, d_b2 = torch.zeros(A, device=DEV), torch.zeros(1, device=DEV)
    ms = timeit(lambda: K.call('xnrs_addpool_bwd', x, None, None, hid, w2, attn, d_pooled, None, seg, R, S, F, A, d_hid,
                               d_w2, d_b2, None), flush=flush)
    report(f'xnrs_addpool_bwd ragged R={R} tokens={n} F={F} A={A}', ms, nbytes=(n * F + 2 * n * A + R * F + n) * 4)
    # history level: 1024 users x 50 clicks of 256
    R, S, F = 1024, 50, 256
    x, hid = torch.randn(R * S, F, device=DEV), torch.tanh(torch.randn(R * S, A, device=DEV))
    m = (torch.rand(R * S, device=DEV) > 0.3).float()
    attn, pooled = torch.empty(R, S, device=DEV), torch.empty(R, F, device=DEV)
    ms = timeit(lambda: K.call('xnrs_addpool_fwd', x, None, m, hid, w2, b2, None, R, S, F, A, attn, pooled))
    report(f'xnrs_addpool_fwd dense R={R} L={S} F={F} A={A}', ms, nbytes=(R * S * (F + A + 2) + R * F) * 4)


def bench_misc(flush):
    n, F, U = 56320, 256, 8748
    src = torch.randn(n, F, device=DEV)
    # Zipf-like slot -> article map (popular articles receive many slot gradients)
    idx = (torch.rand(n, device=DEV).pow(3.0) * U).to(torch.int32).clamp_(0, U - 1)
    dst = torch.zeros(U, F, device=DEV)
    ms = timeit(lambda: K.call('xnrs_scatter_add_rows', dst, U, F, idx, n, src, F, -1))
    report(f'xnrs_scatter_add_rows {n} slot rows x {F} -> {U} articles', ms, nbytes=(n * F + 2 * U * F) * 4 + n * 4)
    x = torch.randn(263220, 768, device=DEV)
    out = torch.zeros(768, device=DEV)
    ms = timeit(lambda: K.call('xnrs_colsum', x, x.shape[0], x.shape[1], x.stride(0), out), flush=flush)
    report('xnrs_colsum 263220 x 768 (bias gradient)', ms, nbytes=x.numel() * 4)
    B, N, T = 1024, 5, 256
    u, c, t = torch.randn(B, T, device=DEV), torch.randn(B, N, T, device=DEV), torch.zeros(B, N, device=DEV)
    t[:, 0] = 1
    sc, pr = torch.empty(B, N, device=DEV), torch.empty(B, N, device=DEV)
    du, dc, loss = torch.empty_like(u), torch.empty_like(c), torch.zeros(1, device=DEV)
    ms = timeit(lambda: K.call('xnrs_score_loss', u, c, t, None, 0, B, N, T, 1.0, sc, pr, loss, du, dc))
    report(f'xnrs_score_loss B={B} N={N} T={T} (dot + ReLU-MSE + both gradients)', ms, nbytes=(2 * u.numel() + 2 * c.numel()) * 4,
           note='latency-bound: 6 MB of operands')


def bench_eval(flush):
    # MIND-large-shaped scoring: 160k article vectors of 256, impressions of ~37 candidates, 50-click histories
    V, T, n_imp, H = 160_000, 256, 65_536, 50
    g = torch.Generator(device='cpu').manual_seed(0)
    vecs = torch.randn(V, T, device=DEV)
    sizes = torch.randint(5, 74, (n_imp,), generator=g)
    offsets = torch.cat([torch.zeros(1, dtype=torch.long), sizes.cumsum(0)]).to(DEV)
    n_cand = int(offsets[-1])
    zipf = lambda n: (torch.rand(n, generator=g).pow(3.0) * (V - 1)).to(torch.int32) + 1
    cand, hist = zipf(n_cand).to(DEV), zipf(n_imp * H).view(n_imp, H).to(DEV)
    targets = (torch.rand(n_cand, generator=g) < 0.1).float().to(DEV)
    user = torch.randn(n_imp, T, device=DEV)
    scores = torch.empty(n_cand, device=DEV)
    metrics = torch.empty(n_imp, 6, device=DEV, dtype=torch.float64)
    ms = timeit(lambda: K.call('xnrs_eval_impressions', user, vecs, T, cand, offsets, targets, n_imp, 1, scores, metrics), flush=flush)
    report(f'xnrs_eval_impressions {n_imp} impressions, {n_cand} candidates, T={T} (gather + dot + rank sort + 6 metrics)', ms,
           nbytes=n_cand * (T * 4 + 12) + n_imp * (T * 4 + 8 + 48))
    logit, mask = torch.randn(V, device=DEV) * 0.1, torch.ones(V, device=DEV)
    attn, pooled = torch.empty(n_imp, H, device=DEV), torch.empty(n_imp, T, device=DEV)
    ms = timeit(lambda: K.call('xnrs_logitpool_fwd', vecs, V, T, logit, mask, hist, n_imp, H, attn, pooled), flush=flush)
    report(f'xnrs_logitpool_fwd {n_imp} users x {H} history rows of {T} (per-article logits)', ms,
           nbytes=n_imp * H * (T * 4 + 12) + n_imp * T * 4,
           note='algorithmic bytes count every gathered row; popular (Zipf) rows of the 164 MB table are L2 hits, hence > HBM peak')
    # training shape: 1024 users, 8748 distinct articles of the batch
    V2, R = 8748, 1024
    tab, lg, mk = torch.randn(V2, T, device=DEV), torch.randn(V2, device=DEV) * 0.1, torch.ones(V2, device=DEV)
    ids = (torch.rand(R * H, generator=g).pow(3.0) * (V2 - 1)).to(torch.int32).view(R, H).to(DEV)
    at, po = torch.empty(R, H, device=DEV), torch.empty(R, T, device=DEV)
    K.call('xnrs_logitpool_fwd', tab, V2, T, lg, mk, ids, R, H, at, po)
    dpo, dlg, dtab = torch.randn(R, T, device=DEV), torch.zeros(V2, device=DEV), torch.zeros(V2, T, device=DEV)
    ms = timeit(lambda: K.call('xnrs_logitpool_bwd', tab, V2, T, ids, at, dpo, R, H, dlg, dtab))
    report(f'xnrs_logitpool_bwd {R} users x {H} slots -> {V2} articles (fused gather + scatter reductions)', ms,
           nbytes=R * H * (2 * T * 4 + 8) + R * T * 4, note='atomic-reduction bound: 51200 x 64 16-byte reductions onto 8748 rows')


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--only', default='mha,gather,pool,misc,eval')
    args = ap.parse_args()
    flush = torch.zeros(64 * 1024 * 1024, device=DEV)        # 256 MB > 126 MB L2
    for name in args.only.split(','):
        globals()['bench_' + name](flush)


if __name__ == '__main__':
    main()
