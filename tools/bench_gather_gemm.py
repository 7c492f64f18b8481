"""Micro-benchmark of the fused table gather in the tensor-core GEMMs (TMA tile::gather4) against gather-then-GEMM, at the CL
step's two token-level shapes: fc1 forward (A rows gathered, K-major) and the fc1 weight gradient (B rows gathered,
MN-major).  CUDA-event timed, token table 307 MB (> L2), random token ids.
    python tools/bench_gather_gemm.py [n_tokens]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from xnrs_b200 import kernels as K  # noqa: E402
from xnrs_b200 import _lib  # noqa: E402


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / n


def main():
    T = int(sys.argv[1]) if len(sys.argv) > 1 else 153600
    dev = 'cuda'
    table = torch.randn(100_000, 768, device=dev)
    rows = torch.randint(1, 100_000, (T,), device=dev, dtype=torch.int32)
    w1 = torch.randn(256, 768, device=dev) * 0.03
    b1 = torch.randn(256, device=dev) * 0.1
    dhid = torch.randn(T, 256, device=dev)
    hid = torch.empty(T, 256, device=dev)
    dw = torch.zeros(256, 768, device=dev)
    lib = _lib.lib()
    for prec in ('tf32x3', 'tf32'):
        with K.precision(prec):
            for two in (-1, 0):
                lib.xnrs_set_option(b'gemm_2cta', two)
                x = K.gather_rows(table, rows)
                res = {'precision': prec, 'gemm_2cta': two, 'tokens': T}
                ms_g = timeit(lambda: K.gather_rows(table, rows))
                res['gather_rows_ms'] = round(ms_g, 4)
                ms = timeit(lambda: K.gemm(x, w1, trans_b=True, bias=b1, act=K.ACT_TANH, out=hid))
                res['fc1_dense'] = {'ms': round(ms, 4), 'kernel': lib.xnrs_last_gemm_kernel().decode()}
                ref = hid.clone()
                ms = timeit(lambda: K.gemm(table, w1, trans_b=True, bias=b1, act=K.ACT_TANH, out=hid, a_rows=rows))
                res['fc1_gather4'] = {'ms': round(ms, 4), 'kernel': lib.xnrs_last_gemm_kernel().decode(),
                                      'max_abs_diff_vs_dense': float((hid - ref).abs().max())}
                ms = timeit(lambda: K.gemm(dhid, x, trans_a=True, out=dw, accumulate=True))
                res['dw_dense'] = {'ms': round(ms, 4), 'kernel': lib.xnrs_last_gemm_kernel().decode()}
                dw.zero_()
                K.gemm(dhid, x, trans_a=True, out=dw, accumulate=True)
                ref = dw.clone()
                ms = timeit(lambda: K.gemm(dhid, table, trans_a=True, out=dw, accumulate=True, b_rows=rows))
                dw.zero_()
                K.gemm(dhid, table, trans_a=True, out=dw, accumulate=True, b_rows=rows)
                res['dw_gather4'] = {'ms': round(ms, 4), 'kernel': lib.xnrs_last_gemm_kernel().decode(),
                                     'rel_diff_vs_dense': float((dw - ref).abs().max() / ref.abs().max())}
                print(json.dumps(res), flush=True)
    lib.xnrs_set_option(b'gemm_2cta', -1)


if __name__ == '__main__':
    main()
