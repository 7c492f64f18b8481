"""Every C-ABI call of ONE bench-shaped CL train step, replayed back to back and timed alone (CUDA events around 50 launches
of the same call, warm caches): where the step's time goes call by call, without launch gaps or event overhead per call.
    python tools/bench_step_gemms.py [--precision bf16x3|tf32x3|bf16] [--batch 1024] > gpurun_out/step_calls.jsonl
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from xnrs_b200 import _lib  # noqa: E402
from xnrs_b200 import kernels as K  # noqa: E402
from xnrs_b200 import synthetic as syn  # noqa: E402
from xnrs_b200.data import TitleStore  # noqa: E402
from xnrs_b200.distributed import DataParallelTrainer  # noqa: E402
from xnrs_b200.models import make_model  # noqa: E402
from xnrs_b200.training import ContrastiveRankingTrainer  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--precision', default='bf16x3')
    ap.add_argument('--batch', type=int, default=1024)
    ap.add_argument('--reps', type=int, default=50)
    args = ap.parse_args()
    dev = torch.device('cuda', 0)
    torch.cuda.set_device(0)
    K.set_precision(args.precision)
    cfg = bench.MODEL_CFGS['cl']
    cat = syn.make_catalogue(bench.N_NEWS, bench.SEQ_LEN, bench.VOCAB, 768, seed=0)
    store = TitleStore(cat.token_table.to(dev), cat.title_tokens.to(dev))
    torch.manual_seed(0)
    tr = ContrastiveRankingTrainer(dict(cfg, device=str(dev)), make_model(cfg))
    tr.model.train()
    dp = DataParallelTrainer(tr)
    batches = [syn.index_batch(store, cat, syn.make_train_batch(bench.N_NEWS, args.batch, bench.HIST_LEN, seed=1000 + i), dev) for i in range(3)]
    for b in batches[:2]:
        dp.train_step(b)
    torch.cuda.synchronize()

    calls = []
    real_call = K.call

    def recording_call(name, *a):
        calls.append((name, a))              # the tensors stay alive through this list: the replay reads valid memory
        real_call(name, *a)

    K.call = recording_call
    try:
        dp.train_step(batches[2])
    finally:
        K.call = real_call
    torch.cuda.synchronize()

    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    total_us, rows = 0.0, []
    for name, a in calls:
        for _ in range(3):
            real_call(name, *a)
        t0.record()
        for _ in range(args.reps):
            real_call(name, *a)
        t1.record()
        torch.cuda.synchronize()
        us = t0.elapsed_time(t1) * 1e3 / args.reps
        total_us += us
        row = {'call': name, 'us': round(us, 2), 'ints': [x for x in a if isinstance(x, (int, float)) and not isinstance(x, bool)][:13]}
        if name in ('xnrs_gemm', 'xnrs_gemm_bf16', 'xnrs_gemm_bf16x3'):
            row['kernel'] = _lib.lib().xnrs_last_gemm_kernel().decode()
            flop = 2.0 * a[2] * a[3] * a[4]
            row['tflops'] = round(flop / us / 1e6, 1)
            ia, ib = (8, 12) if name.endswith('x3') else (7, 10)
            row['gather'] = [a[ia] is not None, a[ib] is not None]
        rows.append(row)
    for r in rows:
        print(json.dumps(r))
    agg = {}
    for r in rows:
        d = agg.setdefault(r['call'], [0, 0.0])
        d[0] += 1
        d[1] += r['us']
    print(json.dumps({'summary_us': {k: [n, round(us, 1)] for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])},
                      'total_us': round(total_us, 1), 'calls': len(rows)}))

    # fixed cost of the small tensor-core GEMM: K sweep at M=1024, N=256 (the title / impression level shapes)
    for M, N in ((1024, 256), (8960, 256)):
        for Kd in (32, 64, 128, 256, 512, 1024):
            a = torch.randn(M, Kd, device=dev)
            w = torch.randn(N, Kd, device=dev)
            out = torch.empty(M, N, device=dev)
            for _ in range(3):
                K.gemm(a, w, trans_b=True, out=out)
            t0.record()
            for _ in range(args.reps):
                K.gemm(a, w, trans_b=True, out=out)
            t1.record()
            torch.cuda.synchronize()
            us_eager = t0.elapsed_time(t1) * 1e3 / args.reps
            g = torch.cuda.CUDAGraph()           # 20 launches in one graph: the cost inside a replayed step (no CPU launch bound)
            with torch.cuda.graph(g):
                for _ in range(20):
                    K.gemm(a, w, trans_b=True, out=out)
            g.replay()
            t0.record()
            for _ in range(10):
                g.replay()
            t1.record()
            torch.cuda.synchronize()
            print(json.dumps({'sweep': [M, N, Kd], 'us': round(us_eager, 2), 'us_in_graph': round(t0.elapsed_time(t1) * 1e3 / 200, 2),
                              'kernel': _lib.lib().xnrs_last_gemm_kernel().decode()}))
    # phases inside one CTA of the small GEMM (SM clock stamps, xnrs_debug_gemm_trace), medians over the CTAs, in us at 1.965 GHz
    import ctypes
    buf = torch.zeros(16 * 148, device=dev, dtype=torch.int64)
    for (ta, M, N, Kd) in ((0, 1024, 256, 32), (0, 1024, 256, 256), (0, 1024, 256, 1024), (0, 8960, 256, 256), (1, 256, 256, 1024), (1, 256, 256, 8960)):
        a = torch.randn(Kd, M, device=dev) if ta else torch.randn(M, Kd, device=dev)
        w = torch.randn(N, Kd, device=dev) if not ta else torch.randn(Kd, N, device=dev)
        out = torch.zeros(M, N, device=dev)
        kw = dict(trans_a=True) if ta else dict(trans_b=True)
        for _ in range(3):
            K.gemm(a, w, out=out, **kw)
        buf.zero_()
        _lib.lib().xnrs_debug_gemm_trace(ctypes.c_void_p(buf.data_ptr()))
        K.gemm(a, w, out=out, **kw)
        _lib.lib().xnrs_debug_gemm_trace(None)
        torch.cuda.synchronize()
        t = buf.view(148, 16).cpu()
        t = t[t[:, 0] > 0]
        d = (t[:, 1:] - t[:, :1]).double() / 1965.0
        c0 = (t[0, 1:] - t[0, 0]).double() / 1965.0
        names = ['setup', 'first_operands', 'last_mma_issued', 'acc_ready', 'epilogue_done', 'all_roles_done', 'tmem_freed', 'epi_ld1', 'epi_ld2', 'epi_chunk0', 'epi_chunk1', 'epi_chunk2', 'x13', 'x14', 'x15']
        print(json.dumps({'trace': [ta, M, N, Kd], 'ctas': int(t.shape[0]), 'kernel': _lib.lib().xnrs_last_gemm_kernel().decode(),
                          'us_since_entry_median': {n: round(float(d[:, i].median()), 2) for i, n in enumerate(names)},
                          'cta0': [round(float(x), 2) for x in c0],
                          'us_since_entry_max': {n: round(float(d[:, i].max()), 2) for i, n in enumerate(names)}}))
    # an empty-ish kernel for scale: launch + drain of a 1-CTA kernel
    x = torch.zeros(4, device=dev)
    t0.record()
    for _ in range(200):
        K.call('xnrs_infonce_finalize', x, x[2:])
    t1.record()
    torch.cuda.synchronize()
    us_eager = t0.elapsed_time(t1) * 1e3 / 200
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(20):
            K.call('xnrs_infonce_finalize', x, x[2:])
    g.replay()
    t0.record()
    for _ in range(10):
        g.replay()
    t1.record()
    torch.cuda.synchronize()
    print(json.dumps({'one_thread_kernel_us': round(us_eager, 2), 'us_in_graph': round(t0.elapsed_time(t1) * 1e3 / 200, 2)}))


if __name__ == '__main__':
    main()
