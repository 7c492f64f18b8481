"""GEMM micro-benchmark on the B200: SIMT fp32 vs tcgen05 (3xTF32 / TF32) at the hot-path shapes. CUDA-event timed."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from xnrs_b200 import kernels as K

SHAPES = [  # name, M, N, K, trans_a, trans_b
    ('fc1 fwd   (tokens x 256 <- 768)', 262144, 256, 768, False, True),
    ('fc1 fwd+tanh (157306 x 256 <- 768)', 157306, 256, 768, False, True, 'tanh'),
    ('fc1 fwd      (157306 x 256 <- 768)', 157306, 256, 768, False, True),
    ('head fwd  (titles x 256 <- 768)', 56320, 256, 768, False, True),
    ('qkv fwd   (tokens x 768 <- 768)', 262144, 768, 768, False, True),
    ('dW fc1    (256 x 768, K=tokens)', 256, 768, 262144, True, False),
    ('dX        (tokens x 768 <- 256)', 262144, 768, 256, False, False),
    ('head fwd small (8748 x 256 <- 768)', 8748, 256, 768, False, True),
    ('head dW small (256 x 768, K=8748)', 256, 768, 8748, True, False),
    ('user head (1024 x 256 <- 256)', 1024, 256, 256, False, True),
    ('user head dW (256 x 256, K=1024)', 256, 256, 1024, True, False),
    ('user fc1 (51200 x 256 <- 256)', 51200, 256, 256, False, True),
]


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


ONLY = os.environ.get('SHAPE_IDX')
PRECS = os.environ.get('PRECS', 'fp32,tf32x3,tf32').split(',')
for name, M, N, Kd, ta, tb, *extra in (SHAPES if ONLY is None else [SHAPES[int(ONLY)]]):
    act = K.ACT_TANH if extra else K.ACT_NONE
    bias = torch.randn(N, device='cuda') if extra else None
    a = torch.randn((Kd, M) if ta else (M, Kd), device='cuda')
    b = torch.randn((N, Kd) if tb else (Kd, N), device='cuda')
    out = torch.empty(M, N, device='cuda')
    row = {'shape': name}
    for prec in PRECS:
        with K.precision(prec):
            ms = timeit(lambda: K.gemm(a, b, trans_a=ta, trans_b=tb, out=out, bias=bias, act=act))
        row[prec] = f'{ms:.3f} ms {2.0 * M * N * Kd / ms / 1e9:.1f} TF/s'
    torch.backends.cuda.matmul.allow_tf32 = False
    A, B = (a.T if ta else a), (b.T if tb else b)
    ms = timeit(lambda: torch.matmul(A, B, out=out))
    row['cublas_fp32'] = f'{ms:.3f} ms {2.0 * M * N * Kd / ms / 1e9:.1f} TF/s'
    torch.backends.cuda.matmul.allow_tf32 = True
    ms = timeit(lambda: torch.matmul(A, B, out=out))
    row['cublas_tf32'] = f'{ms:.3f} ms {2.0 * M * N * Kd / ms / 1e9:.1f} TF/s'
    print(json.dumps(row))
