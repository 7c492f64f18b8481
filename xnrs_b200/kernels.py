"""torch-tensor front end of the C ABI (include/xnrs_b200.h) and the autograd glue around it.

PyTorch is plumbing here: it owns device memory, the current CUDA stream and the autograd tape.
Every FLOP of the hot path runs in the kernels of xnrs_b200/csrc through ``call``; nothing in this
file computes with torch ops, and nothing falls back to the CPU — non-CUDA inputs raise.
"""
from __future__ import annotations

import contextlib
from typing import Optional

import torch

from . import _lib

ACT_NONE, ACT_RELU, ACT_TANH, ACT_RELU_MASK = 0, 1, 2, 3
LOSS_MSE_RELU, LOSS_BCE_LOGITS, LOSS_NLL, LOSS_BCE_SIGMOID = 0, 1, 2, 3
PRECISIONS = {'fp32': 0, 'tf32x3': 1, 'tf32': 2, 'bf16': 3, 'bf16x3': 4}
_precision = 0


def set_precision(name: str) -> None:
    """arithmetic of the GEMM-shaped ops: 'fp32' (exact FMA), 'tf32x3' (fp32-accurate tensor cores), 'tf32' (single pass) or
    'bf16' (2e-2 tolerance class): the token-level tensors of the additive-pooling title encoder — token-table rows, the
    tanh hidden layer and its gradient, i.e. ~99 % of the step's bytes and FLOPs — are STORED in bf16 and multiplied by
    tcgen05 kind::f16 with fp32 accumulation; the small title / user level GEMMs on fp32 tensors run single-pass TF32.
    'bf16x3' is fp32-accurate like 'tf32x3' (1e-4 class, measured ~1e-6): the two token-level tensor-core launches of the
    additive-pooling title encoder run 3xBF16 on operands PRE-SPLIT into two bf16 planes (x ~ hi + lo: the frozen token
    table once, fc1.weight once per step, d_hid written as planes by the pooling backward) — no in-kernel split pass, twice the
    MMA rate of TF32; everything else is 'tf32x3'."""
    global _precision
    _precision = PRECISIONS[name]


def _gemm_precision() -> int:
    """precision code handed to xnrs_gemm (fp32 operands): the bf16-storage mode multiplies them in single-pass TF32"""
    return 2 if _precision == 3 else (1 if _precision == 4 else _precision)


def get_precision() -> str:
    return {v: k for k, v in PRECISIONS.items()}[_precision]


@contextlib.contextmanager
def precision(name: str):
    global _precision
    old = _precision
    set_precision(name)
    try:
        yield
    finally:
        _precision = old


class _Strided:
    """marks a 2-D operand that may be a row-strided view (unit column stride, explicit leading dimension)"""
    __slots__ = ('t',)

    def __init__(self, t: torch.Tensor):
        if t.dim() != 2 or t.stride(1) != 1 or t.dtype != torch.float32:
            raise RuntimeError('gemm operands must be fp32 2-D with unit column stride')
        self.t = t


_Tensor = torch.Tensor


def _arg(a):
    # hot path: ~10 arguments per launch, ~45 launches per step — exact-type checks first (scalars, None), then tensors
    if a is None:
        return None
    t = type(a)
    if t is int or t is float:
        return a
    if t is _Strided:
        a = a.t
        if not a.is_cuda:
            raise RuntimeError('xnrs_b200 kernels need CUDA tensors (there is no CPU path)')
        return a.data_ptr()
    if isinstance(a, _Tensor):
        if not a.is_cuda:
            raise RuntimeError('xnrs_b200 kernels need CUDA tensors (there is no CPU path)')
        if not a.is_contiguous():
            raise RuntimeError('xnrs_b200 kernels need contiguous tensors')
        return a.data_ptr()
    return a


_event_hook = None
_fns = {}
_raw_stream = getattr(torch._C, '_cuda_getCurrentRawStream', None)
_get_device = getattr(torch._C, '_cuda_getDevice', None)


def set_event_hook(hook) -> None:
    """bench.py timing: ``hook(name, scalar_args) -> record`` with ``.start`` / ``.end`` CUDA events (recorded here on the
    stream the kernel is launched on, torch's current stream) and a ``.kernel`` slot that receives the name of the GEMM
    kernel the library dispatched to."""
    global _event_hook
    _event_hook = hook


def _current_stream_handle() -> int:
    """raw cudaStream_t of torch's current stream (the fast private accessor when this torch has it: the public
    torch.cuda.current_stream() costs ~10 us per call, a third of the launch path)"""
    if _raw_stream is not None and _get_device is not None:
        return _raw_stream(_get_device())
    return torch.cuda.current_stream().cuda_stream


def call(name: str, *args) -> None:
    """invoke one C-ABI entry point on the current torch stream; raise on a non-zero status."""
    fn = _fns.get(name)
    if fn is None:
        fn = _fns[name] = getattr(_lib.lib(), name)
    if _event_hook is not None:
        rec = _event_hook(name, tuple(a for a in args if isinstance(a, (int, float))))
        rec.start.record()
        rc = fn(*[_arg(a) for a in args], torch.cuda.current_stream().cuda_stream)
        rec.end.record()
        if name == 'xnrs_gemm':              # which kernel the library dispatched this GEMM to (its claim, not a guess)
            rec.kernel = _lib.lib().xnrs_last_gemm_kernel().decode()
    else:
        rc = fn(*[_arg(a) for a in args], _current_stream_handle())
    if rc != 0:
        raise RuntimeError(f'{name} failed ({rc}): {_lib.last_error()}')


def _mat(t: torch.Tensor) -> _Strided:
    return _Strided(t)


def launch_count() -> int:
    return int(_lib.lib().xnrs_launch_count())


def _f32(t: torch.Tensor) -> torch.Tensor:
    if t.dtype != torch.float32:
        raise RuntimeError(f'expected float32, got {t.dtype}')
    return t.contiguous()


def _i32(t: torch.Tensor) -> torch.Tensor:
    return t.to(torch.int32).contiguous()


# ------------------------------------------------------------------------------------------------
# raw ops (no autograd)
# ------------------------------------------------------------------------------------------------

def gemm(a, b, *, trans_a=False, trans_b=False, bias=None, act=ACT_NONE, aux=None, out=None, accumulate=False,
         a_rows=None, b_rows=None, split_k=0):
    """out (M,N) (=|+=) act(op(a) @ op(b) + bias).  `a`/`b` are 2-D row-major; with a_rows/b_rows they are
    tables whose stored rows are gathered through the index."""
    rows_a = a_rows.numel() if a_rows is not None else a.shape[0]
    rows_b = b_rows.numel() if b_rows is not None else b.shape[0]
    M, K = (a.shape[1], rows_a) if trans_a else (rows_a, a.shape[1])
    N, Kb = (rows_b, b.shape[1]) if trans_b else (b.shape[1], rows_b)
    if K != Kb:
        raise RuntimeError(f'gemm: inner dimensions differ ({K} vs {Kb})')
    if out is None:
        out = torch.empty((M, N), device=a.device, dtype=torch.float32)
        accumulate = False
    call('xnrs_gemm', int(trans_a), int(trans_b), M, N, K, _mat(a), a.stride(0), a_rows, _mat(b), b.stride(0), b_rows,
         _mat(out), out.stride(0), bias, act, aux, int(accumulate), split_k, _gemm_precision())
    return out


def cast_bf16(t: torch.Tensor) -> torch.Tensor:
    """fp32 -> bf16 (round to nearest even), same shape"""
    t = _f32(t)
    out = torch.empty(t.shape, device=t.device, dtype=torch.bfloat16)
    call('xnrs_cast_bf16', t.numel(), t, out)
    return out


def bf16_twin(t: torch.Tensor) -> torch.Tensor:
    """the bf16 copy of a FROZEN fp32 table (the token table), made once and cached on the tensor object"""
    twin = getattr(t, '_xnrs_bf16', None)
    if twin is None or twin.shape != t.shape or twin.device != t.device:
        twin = cast_bf16(t)
        t._xnrs_bf16 = twin
    return twin


def split_bf16(t: torch.Tensor, fp16: bool = False):
    """fp32 -> two 16-bit planes (hi, lo) with t ~ hi + lo (hi = round16(t), lo = round16(t - hi)).  bf16 planes: 16 mantissa bits,
    any magnitude (gradients).  fp16 planes: 22 mantissa bits — as accurate as 3xTF32 — for values within +-65504 (embeddings,
    weights; beyond that they saturate)."""
    t = _f32(t)
    dt = torch.float16 if fp16 else torch.bfloat16
    hi = torch.empty(t.shape, device=t.device, dtype=dt)
    lo = torch.empty(t.shape, device=t.device, dtype=dt)
    call('xnrs_split_bf16', t.numel(), t, hi, lo, int(fp16))
    return hi, lo


def bf16_split_twin(t: torch.Tensor, fp16: bool):
    """the two 16-bit planes of a FROZEN fp32 table (the token table), made once per format and cached on the tensor object:
    fp16 planes for the forward product (with the fp16 planes of fc1.weight), bf16 planes for the weight gradient (with the bf16
    planes of d_hid) — the two operands of a kind::f16 MMA must share their format"""
    cache = getattr(t, '_xnrs_bf16x3', None)
    if cache is None or cache.get('shape') != t.shape or cache.get('device') != t.device:
        cache = {'shape': t.shape, 'device': t.device}
        t._xnrs_bf16x3 = cache
    if fp16 not in cache:
        cache[fp16] = split_bf16(t, fp16=fp16)
    return cache[fp16]


def gemm_bf16x3(a_hi, a_lo, b_hi, b_lo, *, trans_a=False, trans_b=False, bias=None, act=ACT_NONE, out=None, accumulate=False,
                a_rows=None, b_rows=None, split_k=0):
    """xnrs_gemm_bf16x3: fp32-accurate product of operands given as two bf16 planes each; out fp32"""
    rows_a = a_rows.numel() if a_rows is not None else a_hi.shape[0]
    rows_b = b_rows.numel() if b_rows is not None else b_hi.shape[0]
    M, K = (a_hi.shape[1], rows_a) if trans_a else (rows_a, a_hi.shape[1])
    N, Kb = (rows_b, b_hi.shape[1]) if trans_b else (b_hi.shape[1], rows_b)
    if K != Kb:
        raise RuntimeError(f'gemm_bf16x3: inner dimensions differ ({K} vs {Kb})')
    for t_ in (a_hi, a_lo, b_hi, b_lo):
        if t_.dtype not in (torch.bfloat16, torch.float16) or t_.stride(1) != 1:
            raise RuntimeError('gemm_bf16x3 operands must be bf16 / fp16 2-D with unit column stride')
    if (a_hi.stride(0) != a_lo.stride(0) or b_hi.stride(0) != b_lo.stride(0) or a_hi.dtype != a_lo.dtype or b_hi.dtype != b_lo.dtype
            or a_hi.dtype != b_hi.dtype):
        raise RuntimeError('gemm_bf16x3: the planes of an operand share their leading dimension, and all four planes their type')
    if out is None:
        out = torch.empty((M, N), device=a_hi.device, dtype=torch.float32)
        accumulate = False
    call('xnrs_gemm_bf16x3', int(trans_a), int(trans_b), M, N, K, a_hi, a_lo, int(a_hi.dtype == torch.float16), a_hi.stride(0), a_rows,
         b_hi, b_lo, int(b_hi.dtype == torch.float16), b_hi.stride(0), b_rows, out, out.stride(0), bias, act, int(accumulate), split_k)
    return out


def gemm_bf16(a, b, *, trans_a=False, trans_b=False, bias=None, act=ACT_NONE, out=None, out_bf16=False, accumulate=False,
              a_rows=None, b_rows=None, split_k=0):
    """xnrs_gemm_bf16: bf16 operands (2-D, unit column stride), fp32 accumulation; out fp32 (default) or bf16"""
    rows_a = a_rows.numel() if a_rows is not None else a.shape[0]
    rows_b = b_rows.numel() if b_rows is not None else b.shape[0]
    M, K = (a.shape[1], rows_a) if trans_a else (rows_a, a.shape[1])
    N, Kb = (rows_b, b.shape[1]) if trans_b else (b.shape[1], rows_b)
    if K != Kb:
        raise RuntimeError(f'gemm_bf16: inner dimensions differ ({K} vs {Kb})')
    if a.dtype != torch.bfloat16 or b.dtype != torch.bfloat16 or a.stride(1) != 1 or b.stride(1) != 1:
        raise RuntimeError('gemm_bf16 operands must be bf16 2-D with unit column stride')
    if out is None:
        out = torch.empty((M, N), device=a.device, dtype=torch.bfloat16 if out_bf16 else torch.float32)
        accumulate = False
    call('xnrs_gemm_bf16', int(trans_a), int(trans_b), M, N, K, a, a.stride(0), a_rows, b, b.stride(0), b_rows, out, out.stride(0),
         int(out.dtype == torch.bfloat16), bias, act, int(accumulate), split_k)
    return out


def colsum_into(x2d: torch.Tensor, out: torch.Tensor) -> None:
    call('xnrs_colsum', x2d, x2d.shape[0], x2d.shape[1], x2d.stride(0), out)


def colsum(x2d: torch.Tensor) -> torch.Tensor:
    out = torch.zeros(x2d.shape[1], device=x2d.device, dtype=torch.float32)
    colsum_into(x2d, out)
    return out


def expand_titles(title_tokens: torch.Tensor, news_ids: torch.Tensor):
    """news ids (any shape) -> (token rows (R*S,) int32, mask (R*S,) fp32)"""
    n_news, S = title_tokens.shape
    ids = _i32(news_ids).reshape(-1)
    rows = torch.empty(ids.numel() * S, device=ids.device, dtype=torch.int32)
    mask = torch.empty(ids.numel() * S, device=ids.device, dtype=torch.float32)
    call('xnrs_expand_titles', title_tokens, n_news, S, ids, ids.numel(), rows, mask)
    return rows, mask


def gather_rows(table: torch.Tensor, rows: torch.Tensor) -> torch.Tensor:
    rows = _i32(rows).reshape(-1)
    out = torch.empty((rows.numel(), table.shape[1]), device=table.device, dtype=torch.float32)
    call('xnrs_gather_rows', table, table.shape[0], table.shape[1], rows, rows.numel(), out, out.stride(0))
    return out


def collapse_mask(mask: torch.Tensor, R: int, L: int) -> torch.Tensor:
    out = torch.empty(R, device=mask.device, dtype=torch.float32)
    call('xnrs_collapse_mask', mask, R, L, out)
    return out


def rowdot(x: torch.Tensor, w: torch.Tensor, b: Optional[torch.Tensor]) -> torch.Tensor:
    out = torch.empty(x.shape[0], device=x.device, dtype=torch.float32)
    call('xnrs_rowdot', x, w, b, x.shape[0], x.shape[1], out)
    return out


def logitpool(table: torch.Tensor, logit: torch.Tensor, row_mask: Optional[torch.Tensor], ids: torch.Tensor) -> torch.Tensor:
    """additive pooling over per-item logits: ids (R,L) int32 rows of `table` -> pooled (R,T)"""
    R, L = ids.shape
    pooled = torch.empty((R, table.shape[1]), device=table.device, dtype=torch.float32)
    call('xnrs_logitpool_fwd', table, table.shape[0], table.shape[1], logit, row_mask, _i32(ids), R, L, None, pooled)
    return pooled


def adam_step(p, g, m, v, lr, beta1=0.9, beta2=0.999, eps=1e-8, step=1, bc_dev=None, grad_scale=1.0):
    call('xnrs_adam_step', p, g, m, v, p.numel(), lr, beta1, beta2, eps, step, bc_dev, grad_scale)


def eval_impressions(user, news_vecs, cand_ids, offsets, targets, act=1, scores=None):
    """-> (scores (n_cand,), metrics (n_imp,6) float64: auc, rr, ndcg@5, ndcg@10, ctr@1, ctr@10)"""
    n_imp = offsets.numel() - 1
    dev = targets.device
    if scores is None:
        scores = torch.empty(targets.numel(), device=dev, dtype=torch.float32)
    metrics = torch.empty((n_imp, 6), device=dev, dtype=torch.float64)
    n_news, T = (news_vecs.shape[0], news_vecs.shape[1]) if news_vecs is not None else (0, 0)
    call('xnrs_eval_impressions', user, news_vecs, n_news, T, cand_ids, offsets, targets, n_imp, act, scores, metrics)
    return scores, metrics


def binary_metrics(scores: torch.Tensor, targets: torch.Tensor, offsets: torch.Tensor) -> torch.Tensor:
    """-> (n_imp,7) float64: accuracy, recall, precision, tn, fp, fn, tp of round(clip(score,0,1)) per impression"""
    n_imp = offsets.numel() - 1
    out = torch.empty((n_imp, 7), device=scores.device, dtype=torch.float64)
    call('xnrs_binary_metrics', scores, targets, offsets, n_imp, out)
    return out


def metric_sums(metrics: torch.Tensor) -> torch.Tensor:
    sums = torch.zeros(7, device=metrics.device, dtype=torch.float64)
    call('xnrs_metric_sums', metrics, metrics.shape[0], sums)
    return sums


# ------------------------------------------------------------------------------------------------
# autograd functions.  Convention: `x` is either a dense (rows, F) matrix or — with `rows` — a frozen
# table whose rows are gathered on the fly (row G fused into the consumer).
# ------------------------------------------------------------------------------------------------

def _need(ctx, i):
    return ctx.needs_input_grad[i]


FUSED_GATHER = bool(int(__import__('os').environ.get('XNRS_FUSED_GATHER', '1')))
FUSED_GATHER_MIN_ROWS = 40960       # >= two waves of 256-row pair tiles: below that the 1-CTA kernel (TMA gather4) would run


def _resolve_rows(x, rows, fuse_ok: bool = True):
    """how the fused table gather (row G) reaches the GEMMs.  The exact-fp32 SIMT GEMM gathers rows inside its tile loads.
    The 3xTF32 CTA-pair GEMM gathers with a cp.async producer warp straight into its swizzled operand tiles (forward: rows of
    A; weight gradient: rows of B), so for the token-level GEMMs of the additive / personalised poolers no dense copy of the
    gathered rows is ever made and the pooling kernels read the table rows themselves.  Everything else (small problems,
    single-pass modes, the three projection GEMMs of self-attention that would each gather again) copies the gathered rows
    once (one coalesced pass, bit exact) and reads the dense copy."""
    if rows is None or _precision == 0:
        return x, rows
    if FUSED_GATHER and fuse_ok and _precision in (1, 4) and rows.numel() >= FUSED_GATHER_MIN_ROWS:
        return x, rows
    return gather_rows(x, rows), None


# ---- weight gradients straight into the optimiser's gradient buffer ----------------------------------------------------
# Parameters managed by FlatAdam carry a persistent ``.grad`` view into its flat buffer and the mark ``_xnrs_direct``.  For
# those, the backward GEMM / column-sum accumulates IN PLACE into ``param.grad`` (beta = 1) and autograd is handed ``None``:
# same result as autograd's own ``param.grad += d_param``, without the temporary, its allocation and the add kernel
# (12-20 launches per step).  Any other parameter (no trainer, torch.autograd.grad callers) takes the ordinary path.

_direct_enabled = False


@contextlib.contextmanager
def direct_grads():
    """the trainers wrap THEIR backward pass in this: only then do weight gradients go straight into FlatAdam's buffer.  Any
    other autograd user (torch.autograd.grad, the explainer's input-gradient loop) gets ordinary gradient tensors back and
    leaves the optimiser's buffer untouched."""
    global _direct_enabled
    old, _direct_enabled = _direct_enabled, True
    try:
        yield
    finally:
        _direct_enabled = old


def _direct(param) -> Optional[torch.Tensor]:
    if not _direct_enabled:
        return None
    g = param.grad if getattr(param, '_xnrs_direct', False) else None
    return g if (g is not None and g.is_contiguous()) else None


# ---- weight gradients on a parallel branch --------------------------------------------------------------------------
# A title / impression level GEMM leaves most SMs idle (4-140 CTAs of 9-12 us each, mostly fixed cost), and in a backward
# pass the weight gradient (dW = dY^T X) and the input gradient (dX = dY W) of a layer are independent.  When the weight
# gradient is accumulated IN PLACE into the optimiser's buffer (nothing is allocated), it is launched on a side stream that
# forks from the current one and joins it again before the function returns: the two kernels overlap, and inside the step's
# CUDA graph they become parallel branches.  XNRS_WGRAD_BRANCH=0 launches everything on one stream.
WGRAD_BRANCH = bool(int(__import__('os').environ.get('XNRS_WGRAD_BRANCH', '1')))
_side_streams = {}


class _Branch:
    """with _Branch(param, ...) as br: <launches that only accumulate into existing buffers>; ...; br.join()"""

    def __init__(self, *params):
        self.on = bool(WGRAD_BRANCH and torch.cuda.is_available() and all(_direct(p_) is not None for p_ in params if p_ is not None))
        self.ctx = None

    def __enter__(self):
        if self.on:
            self.cur = torch.cuda.current_stream()
            key = self.cur.device_index
            side = _side_streams.get(key)
            if side is None:
                side = _side_streams[key] = torch.cuda.Stream(device=self.cur.device)
            self.side = side
            side.wait_stream(self.cur)
            self.ctx = torch.cuda.stream(side)
            self.ctx.__enter__()
        return self

    def __exit__(self, *exc):
        if self.ctx is not None:
            self.ctx.__exit__(*exc)
            self.ctx = None
        return False

    def join(self):
        if self.on:
            self.cur.wait_stream(self.side)


def _wgrad_gemm(param, a, b, **kw):
    """d_param = a^T @ b (trans_a GEMM), accumulated into param.grad when possible"""
    g = _direct(param)
    if g is not None and g.dim() == 2:
        gemm(a, b, trans_a=True, out=g, accumulate=True, **kw)
        return None
    return gemm(a, b, trans_a=True, **kw)


def _wgrad_colsum(param, x2d):
    g = _direct(param)
    if g is not None:
        colsum_into(x2d, g.view(-1))
        return None
    return colsum(x2d)


def _wgrad_buffer(param, like):
    """an accumulation target for kernels that ADD their parameter gradient: (buffer, value to return to autograd)"""
    g = _direct(param)
    if g is not None and g.numel() == like.numel():
        return g.view(like.shape), None
    z = torch.zeros_like(like)
    return z, z


class LinearFn(torch.autograd.Function):
    """y = x W^T + b (nn.Linear), optional fused table gather on x."""

    @staticmethod
    def forward(ctx, x, rows, weight, bias):
        y = gemm(x, weight, trans_b=True, bias=bias, a_rows=rows)
        ctx.save_for_backward(x, rows, weight)
        ctx.has_bias = bias is not None
        ctx.bias_param = bias
        return y

    @staticmethod
    def backward(ctx, dy):
        x, rows, weight = ctx.saved_tensors
        dy = _f32(dy)
        dx = dw = db = None
        if _need(ctx, 0) and rows is not None:
            raise RuntimeError('no gradient flows into a gathered (frozen) table')
        want_b = ctx.has_bias and _need(ctx, 3)
        with _Branch(weight if _need(ctx, 2) else None, ctx.bias_param if want_b else None) as br:
            if _need(ctx, 2):
                dw = _wgrad_gemm(weight, dy, x, b_rows=rows)
            if want_b:
                db = _wgrad_colsum(ctx.bias_param, dy)
        if _need(ctx, 0):
            dx = gemm(dy, weight)
        br.join()
        return dx, None, dw, db


class Mlp2Fn(torch.autograd.Function):
    """the encoder head Linear -> ReLU -> Linear (news_encoding.py:27-31,55-56; user_encoding.py:29-33)."""

    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2):
        h = gemm(x, w1, trans_b=True, bias=b1, act=ACT_RELU)
        y = gemm(h, w2, trans_b=True, bias=b2)
        ctx.save_for_backward(x, w1, w2, h)
        ctx.bias = b1 is not None
        ctx.bias_params = (b1, b2)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w1, w2, h = ctx.saved_tensors
        dy = _f32(dy)
        b1, b2 = ctx.bias_params
        with _Branch(w2, b2 if ctx.bias else None) as br2:      # layer 2: weight / bias gradient beside the input gradient
            dw2 = _wgrad_gemm(w2, dy, h)
            db2 = _wgrad_colsum(b2, dy) if ctx.bias else None
        dh = gemm(dy, w2, act=ACT_RELU_MASK, aux=h)
        br2.join()
        with _Branch(w1, b1 if ctx.bias else None) as br1:
            dw1 = _wgrad_gemm(w1, dh, x)
            db1 = _wgrad_colsum(b1, dh) if ctx.bias else None
        dx = gemm(dh, w1) if _need(ctx, 0) else None
        br1.join()
        return dx, dw1, db1, dw2, db2


FUSED_TITLEPOOL = bool(int(__import__('os').environ.get('XNRS_FUSED_TITLEPOOL', '1')))


class AdditivePoolFn(torch.autograd.Function):
    """layers.AdditiveAttention (layers.py:47-69) over R groups of L rows: -> pooled (R,F), attn (R,L)."""

    @staticmethod
    def forward(ctx, x, rows, mask, w1, b1, w2, b2, R, L, seg=None, tix=None):
        """seg (R+1 int32, optional): ragged groups — group r owns rows [seg[r], seg[r+1]); L = longest group.
        tix (optional, with seg): the group of each row (-1: padding row) — enables the ONE-launch fused forward
        (xnrs_titlepool_fwd: gather -> fc1 -> tanh -> logit -> exp -> per-title sums on the tensor-core kernel)."""
        F_, A = x.shape[1], w1.shape[0]
        ctx.set_materialize_grads(False)            # an unused `attn` output arrives as None in backward, not as zeros
        shape_ok = (tix is not None and seg is not None and mask is None and A == 256 and F_ % 128 == 0 and F_ <= 1024
                    and (rows.numel() if rows is not None else x.shape[0]) >= FUSED_GATHER_MIN_ROWS)
        ctx.bf16 = bool(_precision == 3 and FUSED_TITLEPOOL and rows is not None and shape_ok)
        ctx.x3 = bool(_precision == 4 and FUSED_TITLEPOOL and rows is not None and shape_ok and F_ <= 768 and x.stride(0) % 8 == 0)
        if ctx.x3:
            # fp32-accurate 3-pass 16-bit arithmetic on pre-split planes: the frozen table's fp16 planes are cached, fc1.weight is
            # split (fp16) for this step; hid stays fp32; the weighted sums and the backward's row dots read the fp32 table
            (xh, xl), n_rows = bf16_split_twin(x, True), rows.numel()
            w1h, w1l = split_bf16(w1, fp16=True)
            hid = torch.empty((n_rows, A), device=x.device, dtype=torch.float32)
            attn = torch.empty(n_rows, device=x.device, dtype=torch.float32)
            e = torch.empty(n_rows, device=x.device, dtype=torch.float32)
            zsum = torch.empty(R, device=x.device, dtype=torch.float32)
            pooled = torch.empty((R, F_), device=x.device, dtype=torch.float32)
            call('xnrs_titlepool_fwd_bf16x3', xh, xl, xh.stride(0), rows, tix, seg, n_rows, R, F_, A, w1h, w1l, 1, b1, w2.reshape(-1), b2,
                 _mat(x), x.stride(0), hid, e, zsum, attn, pooled)
            ctx.save_for_backward(x, rows, w1, w2, hid, attn, seg)
            ctx.dims = (R, L, F_, A)
            ctx.bias_params = (b1, b2)
            return pooled, attn
        if ctx.bf16:
            # bf16-storage mode: the rows are gathered from the bf16 twin of the frozen table, fc1.weight is rounded to bf16 for
            # this step, the hidden layer is kept in bf16 for the backward pass; logits / weights / pooled sums are fp32
            xb, n_rows = bf16_twin(x), rows.numel()
            hid = torch.empty((n_rows, A), device=x.device, dtype=torch.bfloat16)
            attn = torch.empty(n_rows, device=x.device, dtype=torch.float32)
            e = torch.empty(n_rows, device=x.device, dtype=torch.float32)
            zsum = torch.empty(R, device=x.device, dtype=torch.float32)
            pooled = torch.empty((R, F_), device=x.device, dtype=torch.float32)
            call('xnrs_titlepool_fwd_bf16', xb, F_, rows, tix, seg, n_rows, R, F_, A, cast_bf16(w1), b1, w2.reshape(-1), b2, hid, e, zsum,
                 attn, pooled)
            ctx.save_for_backward(xb, rows, w1, w2, hid, attn, seg)
            ctx.dims = (R, L, F_, A)
            ctx.bias_params = (b1, b2)
            return pooled, attn
        x, rows = _resolve_rows(x, rows)
        n_rows = rows.numel() if rows is not None else x.shape[0]
        fused = (FUSED_TITLEPOOL and shape_ok and _precision in (1, 2, 3, 4) and x.stride(0) % 4 == 0)
        pooled = torch.empty((R, F_), device=x.device, dtype=torch.float32)
        if fused:
            hid = torch.empty((n_rows, A), device=x.device, dtype=torch.float32)
            attn = torch.empty(n_rows, device=x.device, dtype=torch.float32)
            e = torch.empty(n_rows, device=x.device, dtype=torch.float32)
            zsum = torch.empty(R, device=x.device, dtype=torch.float32)
            call('xnrs_titlepool_fwd', _mat(x), x.stride(0), rows, tix, seg, n_rows, R, F_, A, w1, b1, w2.reshape(-1), b2,
                 _gemm_precision(), hid, e, zsum, attn, pooled)
        else:
            hid = gemm(x, w1, trans_b=True, bias=b1, act=ACT_TANH, a_rows=rows)
            attn = torch.empty((R, L) if seg is None else (hid.shape[0],), device=x.device, dtype=torch.float32)
            call('xnrs_addpool_fwd', x, rows, mask, hid, w2.reshape(-1), b2, seg, R, L, F_, A, attn, pooled)
        ctx.save_for_backward(x, rows, w1, w2, hid, attn, seg)
        ctx.dims = (R, L, F_, A)
        ctx.bias_params = (b1, b2)
        return pooled, attn

    @staticmethod
    def backward(ctx, d_pooled, d_attn):
        x, rows, w1, w2, hid, attn, seg = ctx.saved_tensors
        R, L, F_, A = ctx.dims
        dev = x.device
        d_pooled = torch.zeros((R, F_), device=dev, dtype=torch.float32) if d_pooled is None else _f32(d_pooled)
        d_attn = None if d_attn is None else _f32(d_attn)
        b1, b2 = ctx.bias_params
        w2_buf, d_w2 = _wgrad_buffer(w2, w2)
        b2_buf, d_b2 = _wgrad_buffer(b2, b2)
        b1_buf, d_b1 = _wgrad_buffer(b1, b1)            # fc1 bias gradient = column sums of d_hid, fused into the kernel
        if ctx.x3:                                      # d_hid leaves the pooling backward as two bf16 planes
            if d_attn is not None:
                raise RuntimeError('3xBF16 pooling: no gradient path through the returned weights')
            dh_hi = torch.empty(hid.shape, device=dev, dtype=torch.bfloat16)
            dh_lo = torch.empty(hid.shape, device=dev, dtype=torch.bfloat16)
            call('xnrs_addpool_bwd_split', x, rows, hid, w2.reshape(-1), attn, d_pooled, seg, R, L, F_, A, hid.shape[0], dh_hi, dh_lo,
                 w2_buf.view(-1), b2_buf.view(-1), b1_buf.view(-1))
            xh, xl = bf16_split_twin(x, False)
            g = _direct(w1)
            if g is not None and g.dim() == 2:
                gemm_bf16x3(dh_hi, dh_lo, xh, xl, trans_a=True, b_rows=rows, out=g, accumulate=True)
                d_w1 = None
            else:
                d_w1 = gemm_bf16x3(dh_hi, dh_lo, xh, xl, trans_a=True, b_rows=rows)
            return None, None, None, d_w1, d_b1, d_w2, d_b2, None, None, None, None
        d_hid = torch.empty_like(hid)
        if ctx.bf16:                                    # x = the bf16 table twin, hid / d_hid bf16; gradients accumulate in fp32
            if d_attn is not None:
                raise RuntimeError('bf16 pooling: no gradient path through the returned weights')
            call('xnrs_addpool_bwd_bf16', x, rows, hid, w2.reshape(-1), attn, d_pooled, seg, R, L, F_, A, hid.shape[0], d_hid,
                 w2_buf.view(-1), b2_buf.view(-1), b1_buf.view(-1))
            g = _direct(w1)
            if g is not None and g.dim() == 2:
                gemm_bf16(d_hid, x, trans_a=True, b_rows=rows, out=g, accumulate=True)
                d_w1 = None
            else:
                d_w1 = gemm_bf16(d_hid, x, trans_a=True, b_rows=rows)
            return None, None, None, d_w1, d_b1, d_w2, d_b2, None, None, None, None
        need_dx = _need(ctx, 0)
        if need_dx and rows is not None:
            raise RuntimeError('no gradient flows into a gathered (frozen) table')
        d_x = torch.empty_like(x) if need_dx else None
        call('xnrs_addpool_bwd', x, rows, None, hid, w2.reshape(-1), attn, d_pooled, d_attn, seg, R, L, F_, A, hid.shape[0], d_hid,
             w2_buf.view(-1), b2_buf.view(-1), d_x, b1_buf.view(-1))
        d_w1 = _wgrad_gemm(w1, d_hid, x, b_rows=rows)
        if need_dx:
            gemm(d_hid, w1, out=d_x, accumulate=True)
        return d_x, None, None, d_w1, d_b1, d_w2, d_b2, None, None, None, None


class ItemLogitPoolFn(torch.autograd.Function):
    """layers.AdditiveAttention (layers.py:47-69) over R groups whose L rows are GATHERED from a table of distinct items
    (ids (R,L) int32): the logit w2.tanh(fc1 x + b1) + b2 of a slot depends only on the item in it, so fc1 runs once per
    item (V rows) instead of once per slot (R*L rows), forward and backward.  Same values / gradients as pooling the
    gathered rows.  -> pooled (R,T), attn (R,L)."""

    @staticmethod
    def forward(ctx, table, row_mask, ids, w1, b1, w2, b2):
        V, T = table.shape
        R, L = ids.shape
        ids = _i32(ids)
        hid = gemm(table, w1, trans_b=True, bias=b1, act=ACT_TANH)
        logit = rowdot(hid, w2.reshape(-1), b2)
        attn = torch.empty((R, L), device=table.device, dtype=torch.float32)
        pooled = torch.empty((R, T), device=table.device, dtype=torch.float32)
        call('xnrs_logitpool_fwd', table, V, T, logit, row_mask, ids, R, L, attn, pooled)
        ctx.save_for_backward(table, ids, w1, w2, hid, attn)
        ctx.bias_params = (b1, b2)
        ctx.set_materialize_grads(False)
        return pooled, attn

    @staticmethod
    def backward(ctx, d_pooled, d_attn):
        if d_attn is not None:
            raise RuntimeError('ItemLogitPoolFn: no gradient path through the returned weights')
        table, ids, w1, w2, hid, attn = ctx.saved_tensors
        V, T = table.shape
        R, L = ids.shape
        A = w1.shape[0]
        dev = table.device
        d_table = torch.zeros_like(table)
        d_logit = torch.zeros(V, device=dev, dtype=torch.float32)
        call('xnrs_logitpool_bwd', table, V, T, ids, attn, _f32(d_pooled), R, L, d_logit, d_table)
        d_hid = torch.empty_like(hid)
        b1, b2 = ctx.bias_params
        w2_buf, d_w2 = _wgrad_buffer(w2, w2)
        b2_buf, d_b2 = _wgrad_buffer(b2, b2)
        call('xnrs_logit_bwd', hid, w2.reshape(-1), d_logit, V, A, d_hid, w2_buf.view(-1), b2_buf.view(-1))
        with _Branch(w1, b1) as br:
            d_w1 = _wgrad_gemm(w1, d_hid, table)
            d_b1 = _wgrad_colsum(b1, d_hid)
        gemm(d_hid, w1, out=d_table, accumulate=True)
        br.join()
        return d_table, None, None, d_w1, d_b1, d_w2, d_b2


class PersonalizedPoolFn(torch.autograd.Function):
    """layers.PersonalizedAttention (layers.py:88-101); group r uses query row r // rows_per_query."""

    @staticmethod
    def forward(ctx, q, x, rows, mask, xw, xb, qw, qb, R, L, rows_per_query, seg=None):
        F_, A = x.shape[1], xw.shape[0]
        x, rows = _resolve_rows(x, rows)
        hid = gemm(x, xw, trans_b=True, bias=xb, act=ACT_TANH, a_rows=rows)
        qh = gemm(q, qw, trans_b=True, bias=qb)
        attn = torch.empty((R, L) if seg is None else (hid.shape[0],), device=x.device, dtype=torch.float32)
        pooled = torch.empty((R, F_), device=x.device, dtype=torch.float32)
        call('xnrs_perspool_fwd', x, rows, mask, hid, qh, seg, R, L, F_, A, rows_per_query, attn, pooled)
        ctx.save_for_backward(q, x, rows, xw, qw, hid, qh, attn, seg)
        ctx.dims = (R, L, F_, A, rows_per_query)
        return pooled

    @staticmethod
    def backward(ctx, d_pooled):
        q, x, rows, xw, qw, hid, qh, attn, seg = ctx.saved_tensors
        R, L, F_, A, rpq = ctx.dims
        d_pooled = _f32(d_pooled)
        d_hid = torch.empty_like(hid)
        d_qh = torch.zeros_like(qh)
        need_dx = _need(ctx, 1)
        if need_dx and rows is not None:
            raise RuntimeError('no gradient flows into a gathered (frozen) table')
        d_x = torch.empty_like(x) if need_dx else None
        call('xnrs_perspool_bwd', x, rows, None, hid, qh, attn, d_pooled, seg, R, L, F_, A, rpq, hid.shape[0], d_hid, d_qh, d_x)
        d_xw = gemm(d_hid, x, trans_a=True, b_rows=rows)
        d_xb = colsum(d_hid)
        if need_dx:
            gemm(d_hid, xw, out=d_x, accumulate=True)
        d_qw = gemm(d_qh, q, trans_a=True)
        d_qb = colsum(d_qh)
        d_q = gemm(d_qh, qw) if _need(ctx, 0) else None
        return d_q, d_x, None, None, d_xw, d_xb, d_qw, d_qb, None, None, None, None


class MultiHeadAttentionFn(torch.autograd.Function):
    """layers.MultiHeadAttention (layers.py:121-156) over R sequences of L rows."""

    @staticmethod
    def forward(ctx, x, rows, mask, wq, bq, wk, bk, wv, bv, wo, bo, R, L, n_heads, keep, p_drop, seed, q_scale=1.0):
        """q_scale: the kernels divide the logits by sqrt(d_k) (layers.py:135-137, scaled=True); scaled=False passes
        sqrt(d_k) here and q is multiplied by it after its projection"""
        D = wq.shape[0]
        dk = D // n_heads
        # precision 'bf16x3', rows gathered from the frozen token table: the three projections (and their weight gradients) run
        # the 3-pass 16-bit split on pre-split planes — the table's cached fp16 planes gathered inside the GEMM (no dense copy
        # of the rows), the weights split once per step; fp32-accurate (1.3e-6), ~1.7x the 3xTF32 rate
        ctx.x3 = bool(_precision == 4 and rows is not None and rows.numel() >= FUSED_GATHER_MIN_ROWS and D % 8 == 0
                      and x.shape[1] % 8 == 0 and x.stride(0) % 8 == 0 and not _need(ctx, 0))
        if ctx.x3:
            xh, xl = bf16_split_twin(x, True)
            q = gemm_bf16x3(xh, xl, *split_bf16(wq, fp16=True), trans_b=True, bias=bq, a_rows=rows)
            k = gemm_bf16x3(xh, xl, *split_bf16(wk, fp16=True), trans_b=True, bias=bk, a_rows=rows)
            v = gemm_bf16x3(xh, xl, *split_bf16(wv, fp16=True), trans_b=True, bias=bv, a_rows=rows)
        else:
            x, rows = _resolve_rows(x, rows, fuse_ok=False)
            q = gemm(x, wq, trans_b=True, bias=bq, a_rows=rows)
            k = gemm(x, wk, trans_b=True, bias=bk, a_rows=rows)
            v = gemm(x, wv, trans_b=True, bias=bv, a_rows=rows)
        if q_scale != 1.0:
            call('xnrs_axpby', q.numel(), float(q_scale), None, q.clone(), 0.0, q)
        o = torch.empty_like(q)
        lse = torch.empty((R, n_heads, L), device=x.device, dtype=torch.float32)
        call('xnrs_mha_fwd', q, k, v, D, mask, R, L, n_heads, dk, keep, p_drop, seed, o, lse)
        if ctx.x3:      # output projection on fp16 planes of the attention output and of its weight
            y = gemm_bf16x3(*split_bf16(o, fp16=True), *split_bf16(wo, fp16=True), trans_b=True, bias=bo)
        else:
            y = gemm(o, wo, trans_b=True, bias=bo)
        ctx.save_for_backward(x, rows, mask, wq, wk, wv, wo, q, k, v, o, lse, keep)
        ctx.cfg = (R, L, n_heads, dk, D, p_drop, seed)
        ctx.q_scale = float(q_scale)
        ctx.bias_params = (bq, bk, bv, bo)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, rows, mask, wq, wk, wv, wo, q, k, v, o, lse, keep = ctx.saved_tensors
        R, L, h, dk, D, p_drop, seed = ctx.cfg
        dy = _f32(dy)
        bq, bk, bv, bo = ctx.bias_params
        d_bo = _wgrad_colsum(bo, dy)
        if ctx.x3:      # both products of the output projection's backward on bf16 planes (dy is a gradient: bf16's range)
            dyh, dyl = split_bf16(dy)
            oh, ol = split_bf16(o)
            g_ = _direct(wo)
            if g_ is not None and g_.dim() == 2:
                gemm_bf16x3(dyh, dyl, oh, ol, trans_a=True, out=g_, accumulate=True)
                d_wo = None
            else:
                d_wo = gemm_bf16x3(dyh, dyl, oh, ol, trans_a=True)
            del oh, ol
            d_o = gemm_bf16x3(dyh, dyl, *split_bf16(wo))
            del dyh, dyl
        else:
            d_wo = _wgrad_gemm(wo, dy, o)
            d_o = gemm(dy, wo)
        dq, dk_, dv = torch.empty_like(q), torch.empty_like(q), torch.empty_like(q)
        call('xnrs_mha_bwd', q, k, v, o, d_o, D, mask, lse, R, L, h, dk, keep, p_drop, seed, dq, dk_, dv)
        if ctx.q_scale != 1.0:
            call('xnrs_axpby', dq.numel(), ctx.q_scale, None, dq.clone(), 0.0, dq)
        d_bq, d_bk, d_bv = _wgrad_colsum(bq, dq), _wgrad_colsum(bk, dk_), _wgrad_colsum(bv, dv)
        if ctx.x3:      # dW = d^T x on bf16 planes: the gradient is split here (arbitrary magnitude: bf16's range), the table's are cached
            tbh, tbl = bf16_split_twin(x, False)
            d_ws = []
            for w_, d_ in ((wq, dq), (wk, dk_), (wv, dv)):
                dh_, dl_ = split_bf16(d_)
                g_ = _direct(w_)
                if g_ is not None and g_.dim() == 2:
                    gemm_bf16x3(dh_, dl_, tbh, tbl, trans_a=True, b_rows=rows, out=g_, accumulate=True)
                    d_ws.append(None)
                else:
                    d_ws.append(gemm_bf16x3(dh_, dl_, tbh, tbl, trans_a=True, b_rows=rows))
                del dh_, dl_
            d_wq, d_wk, d_wv = d_ws
        else:
            d_wq = _wgrad_gemm(wq, dq, x, b_rows=rows)
            d_wk = _wgrad_gemm(wk, dk_, x, b_rows=rows)
            d_wv = _wgrad_gemm(wv, dv, x, b_rows=rows)
        d_x = None
        if _need(ctx, 0):
            if rows is not None:
                raise RuntimeError('no gradient flows into a gathered (frozen) table')
            d_x = gemm(dq, wq)
            gemm(dk_, wk, out=d_x, accumulate=True)
            gemm(dv, wv, out=d_x, accumulate=True)
        return (d_x, None, None, d_wq, d_bq, d_wk, d_bk, d_wv, d_bv, d_wo, d_bo,
                None, None, None, None, None, None, None)


# Data-parallel runs exchange the gradient of row-sparse tables (the 700k-row user tables of LSTUR / NPA) as (ids, rows)
# instead of all-reducing the dense table: while this list is not None, EmbeddingFn.backward logs every
# lookup of a directly-accumulated parameter here; distributed.DataParallelTrainer consumes it after the backward pass.
sparse_grad_log = None


def mark_active_rows(weight, idx, pad: int = -1) -> None:
    """FlatAdam row-sparse tables: remember the rows that have received a gradient (training.FlatAdam, xnrs_adam_rows)"""
    t = getattr(weight, '_xnrs_rows', None)
    if t is not None and idx.numel():
        call('xnrs_mark_rows', idx, idx.numel(), t['V'], int(pad), t['bitmap'], t['active'], t['count'])


class EmbeddingFn(torch.autograd.Function):
    """nn.Embedding lookup with a dense weight gradient (lstur.py:94-98,180-183; npa.py:12-15; naml.py:34-47)."""

    @staticmethod
    def forward(ctx, weight, idx, padding_idx):
        idx = _i32(idx).reshape(-1)
        ctx.save_for_backward(idx, weight)
        ctx.shape, ctx.pad = weight.shape, -1 if padding_idx is None else int(padding_idx)
        return gather_rows(weight, idx)

    @staticmethod
    def backward(ctx, dy):
        idx, weight = ctx.saved_tensors
        dy = _f32(dy)
        g = _direct(weight)
        if g is not None and g.shape == ctx.shape:
            # FlatAdam-managed table: scatter straight into its gradient view (no V x D temporary, zero fill and add — 3 x 383 MB
            # of traffic per step for the LSTUR user table)
            call('xnrs_scatter_add_rows', g, ctx.shape[0], ctx.shape[1], idx, idx.numel(), dy, dy.stride(0), ctx.pad)
            mark_active_rows(weight, idx, ctx.pad)
            if sparse_grad_log is not None:
                sparse_grad_log.append((weight, idx, dy, ctx.pad))
            return None, None, None
        dw = torch.zeros(ctx.shape, device=dy.device, dtype=torch.float32)
        call('xnrs_scatter_add_rows', dw, ctx.shape[0], ctx.shape[1], idx, idx.numel(), dy, dy.stride(0), ctx.pad)
        mark_active_rows(weight, idx, ctx.pad)         # (a row-sparse optimiser must learn about these rows on this path too)
        return dw, None, None


class GruLastFn(torch.autograd.Function):
    """final GRU state at each sequence's true length (lstur.py:139-153): x (B,L,I), lengths (B,) int32."""

    @staticmethod
    def forward(ctx, x, lengths, w_ih, w_hh, b_ih, b_hh, h0):
        B, L, I = x.shape
        Hd = w_hh.shape[1]
        x2 = x.reshape(B * L, I)
        gi = gemm(x2, w_ih, trans_b=True, bias=b_ih)
        w_hh_t = torch.empty((Hd, 3 * Hd), device=x.device, dtype=torch.float32)
        call('xnrs_transpose', w_hh, 3 * Hd, Hd, w_hh_t)
        hs = torch.empty((B * L, Hd), device=x.device, dtype=torch.float32)      # state BEFORE step t
        gates = torch.empty((B, L, 4 * Hd), device=x.device, dtype=torch.float32)
        h_out = torch.empty((B, Hd), device=x.device, dtype=torch.float32)
        call('xnrs_gru_fwd', gi, w_hh_t, b_hh, h0, lengths, B, L, Hd, hs, gates, h_out)
        ctx.save_for_backward(x2, lengths, w_ih, w_hh, hs, gates)
        ctx.dims = (B, L, I, Hd)
        ctx.has_h0 = h0 is not None
        return h_out

    @staticmethod
    def backward(ctx, d_h):
        x2, lengths, w_ih, w_hh, hs, gates = ctx.saved_tensors
        B, L, I, Hd = ctx.dims
        d_h = _f32(d_h)
        d_gi = torch.empty((B * L, 3 * Hd), device=d_h.device, dtype=torch.float32)
        d_gh = torch.empty((B * L, 3 * Hd), device=d_h.device, dtype=torch.float32)
        d_h0 = torch.empty((B, Hd), device=d_h.device, dtype=torch.float32)
        call('xnrs_gru_bwd', d_h, w_hh, lengths, hs, gates, B, L, Hd, d_gi, d_gh, d_h0)
        d_wih = gemm(d_gi, x2, trans_a=True)
        d_bih = colsum(d_gi)
        d_whh = gemm(d_gh, hs, trans_a=True)
        d_bhh = colsum(d_gh)
        d_x = gemm(d_gi, w_ih).reshape(B, L, I) if _need(ctx, 0) else None
        return d_x, None, d_wih, d_whh, d_bih, d_bhh, (d_h0 if ctx.has_h0 else None)


class DropoutFn(torch.autograd.Function):
    """inverted dropout with an explicit keep mask or a seeded counter-based generator."""

    @staticmethod
    def forward(ctx, x, keep, p, seed):
        y = torch.empty_like(x)
        call('xnrs_dropout', x.numel(), x, keep, p, seed, y)
        ctx.save_for_backward(keep)
        ctx.cfg = (p, seed)
        return y

    @staticmethod
    def backward(ctx, dy):
        (keep,) = ctx.saved_tensors
        dy = _f32(dy)
        dx = torch.empty_like(dy)
        call('xnrs_dropout', dy.numel(), dy, keep, ctx.cfg[0], ctx.cfg[1], dx)
        return dx, None, None, None


class DotScoreFn(torch.autograd.Function):
    """scoring.DotScoring (scoring.py:12-23): u (B,T), c (B,N,T) -> (B,N)."""

    @staticmethod
    def forward(ctx, u, c):
        B, N, T = c.shape
        s = torch.empty((B, N), device=c.device, dtype=torch.float32)
        call('xnrs_dot_score', u, c, B, N, T, s)
        ctx.save_for_backward(u, c)
        return s

    @staticmethod
    def backward(ctx, ds):
        u, c = ctx.saved_tensors
        B, N, T = c.shape
        ds = _f32(ds)
        du, dc = torch.empty_like(u), torch.empty_like(c)
        call('xnrs_dot_score_bwd', u, c, ds, B, N, T, du, dc)
        return du, dc


class ScoreLossFn(torch.autograd.Function):
    """dot scoring fused with a trainer loss; value and both gradients in one pass -> (loss, preds, scores)."""

    @staticmethod
    def forward(ctx, u, c, targets, weights, kind):
        """u (B,T), c (B,N,T)  — or u None and c (B,N) = scores computed upstream."""
        B, N = c.shape[0], c.shape[1]
        T = c.shape[2] if u is not None else 0
        dev = c.device
        scores = torch.empty((B, N), device=dev, dtype=torch.float32)
        preds = torch.empty((B, N), device=dev, dtype=torch.float32)
        loss = torch.empty(1, device=dev, dtype=torch.float32)
        need = (u is not None and u.requires_grad) or c.requires_grad
        du = torch.empty_like(u) if (need and u is not None) else None
        dc = torch.empty_like(c) if need else None
        call('xnrs_score_loss', u, c, targets, weights, kind, B, N, T, 1.0, scores, preds, loss, du, dc)
        ctx.save_for_backward(du, dc)
        ctx.mark_non_differentiable(preds, scores)
        return loss.reshape(()), preds, scores

    @staticmethod
    def backward(ctx, g, _gp, _gs):
        du, dc = ctx.saved_tensors
        g = _f32(g).reshape(1)
        gu = None
        if du is not None:
            gu = torch.empty_like(du)
            call('xnrs_axpby', du.numel(), 1.0, g, du, 0.0, gu)
        gc = torch.empty_like(dc)
        call('xnrs_axpby', dc.numel(), 1.0, g, dc, 0.0, gc)
        return gu, gc, None, None, None


class ReluFn(torch.autograd.Function):
    """the trainer's output activation (training.py:392)."""

    @staticmethod
    def forward(ctx, x):
        x = _f32(x)
        y = torch.empty_like(x)
        call('xnrs_relu', x.numel(), x, y)
        ctx.save_for_backward(y)
        return y

    @staticmethod
    def backward(ctx, dy):
        (y,) = ctx.saved_tensors
        dy = _f32(dy)
        dx = torch.empty_like(dy)
        call('xnrs_relu_bwd', dy.numel(), y, dy, dx)
        return dx


class MeanPoolFn(torch.autograd.Function):
    """layers.MaskedMean (layers.py:19-37): x (R*L, F), mask (R*L) -> (R, F)."""

    @staticmethod
    def forward(ctx, x, mask, R, L):
        F_ = x.shape[1]
        out = torch.empty((R, F_), device=x.device, dtype=torch.float32)
        call('xnrs_meanpool_fwd', x, mask, R, L, F_, out)
        ctx.save_for_backward(mask)
        ctx.dims = (R, L, F_)
        return out

    @staticmethod
    def backward(ctx, d_out):
        (mask,) = ctx.saved_tensors
        R, L, F_ = ctx.dims
        d_x = None
        if _need(ctx, 0):
            d_x = torch.empty((R * L, F_), device=d_out.device, dtype=torch.float32)
            call('xnrs_meanpool_bwd', mask, _f32(d_out), R, L, F_, d_x)
        return d_x, None, None, None


class LinearTanhFn(torch.autograd.Function):
    """tanh(x W^T + b) with the activation in the GEMM epilogue (FCScoring's hidden layer, scoring.py:79-102)."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        y = gemm(x, weight, trans_b=True, bias=bias, act=ACT_TANH)
        ctx.save_for_backward(x, weight, y)
        ctx.bias_param = bias
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight, y = ctx.saved_tensors
        d_pre = torch.empty_like(y)
        call('xnrs_tanh_bwd', y.numel(), y, _f32(dy), d_pre)
        dx = gemm(d_pre, weight) if _need(ctx, 0) else None
        dw = _wgrad_gemm(weight, d_pre, x)
        db = _wgrad_colsum(ctx.bias_param, d_pre) if ctx.bias_param is not None else None
        return dx, dw, db


class AddScalarFn(torch.autograd.Function):
    """x + b for a one-element parameter b (the bias of nn.Bilinear(.., out_features=1), scoring.py:45-50)."""

    @staticmethod
    def forward(ctx, x, b):
        x = _f32(x)
        y = torch.empty_like(x)
        call('xnrs_add_scalar', x.numel(), x, b, y)
        ctx.bparam = b
        return y

    @staticmethod
    def backward(ctx, dy):
        dy = _f32(dy)
        return dy, _wgrad_colsum(ctx.bparam, dy.reshape(-1, 1))


class NormalizeRowsFn(torch.autograd.Function):
    """x / ||x||_2 per row (DotScoring / BilinScoring normalize=True, scoring.py:20-22,63-65)."""

    @staticmethod
    def forward(ctx, x):
        x = _f32(x)
        y = torch.empty_like(x)
        inv = torch.empty(x.shape[0], device=x.device, dtype=torch.float32)
        call('xnrs_infonce_normalize', x, x.shape[0], x.shape[1], y, inv)
        ctx.save_for_backward(y, inv)
        return y

    @staticmethod
    def backward(ctx, dy):
        y, inv = ctx.saved_tensors
        dx = torch.empty_like(y)
        call('xnrs_infonce_normalize_bwd', _f32(dy), y, inv, None, 1.0, y.shape[0], y.shape[1], dx)
        return dx


class ScaleFn(torch.autograd.Function):
    """a * x for a host scalar a (MultiHeadAttention(scaled=False): undoes the kernel's 1/sqrt(d_k))."""

    @staticmethod
    def forward(ctx, x, a):
        ctx.a = float(a)
        y = torch.empty_like(x)
        call('xnrs_axpby', x.numel(), ctx.a, None, x, 0.0, y)
        return y

    @staticmethod
    def backward(ctx, dy):
        dy = _f32(dy)
        dx = torch.empty_like(dy)
        call('xnrs_axpby', dy.numel(), ctx.a, None, dy, 0.0, dx)
        return dx, None


class SigmoidFn(torch.autograd.Function):
    """BCERankingTrainer's output activation (training.py:329-331)."""

    @staticmethod
    def forward(ctx, x):
        x = _f32(x)
        y = torch.empty_like(x)
        call('xnrs_sigmoid', x.numel(), x, y)
        ctx.save_for_backward(y)
        return y

    @staticmethod
    def backward(ctx, dy):
        (y,) = ctx.saved_tensors
        dy = _f32(dy)
        dx = torch.empty_like(dy)
        call('xnrs_sigmoid_bwd', dy.numel(), y, dy, dx)
        return dx


class InfoNCEFn(torch.autograd.Function):
    """ContrastiveRankingTrainer._compute_contrastive_loss (training.py:433-472), single device."""

    @staticmethod
    def forward(ctx, emb, labels, temperature):
        B, E = emb.shape
        dev = emb.device
        ehat = torch.empty_like(emb)
        inv_norm = torch.empty(B, device=dev, dtype=torch.float32)
        call('xnrs_infonce_normalize', emb, B, E, ehat, inv_norm)
        sim = gemm(ehat, ehat, trans_b=True)
        stats = torch.zeros(2, device=dev, dtype=torch.float32)
        call('xnrs_infonce_rows', sim, labels, B, B, 0, temperature, stats)
        loss = torch.empty(1, device=dev, dtype=torch.float32)
        call('xnrs_infonce_finalize', stats, loss)
        ctx.save_for_backward(ehat, inv_norm, sim, stats)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, g):
        ehat, inv_norm, G, stats = ctx.saved_tensors
        B, E = ehat.shape
        d_ehat = gemm(G, ehat)                                   # anchor side:  G  ehat
        gemm(G, ehat, trans_a=True, out=d_ehat, accumulate=True)  # key side:     G^T ehat
        d_scaled = torch.empty_like(ehat)
        call('xnrs_infonce_normalize_bwd', d_ehat, ehat, inv_norm, stats, 1.0, B, E, d_scaled)
        d_emb = torch.empty_like(ehat)
        call('xnrs_axpby', d_emb.numel(), 1.0, _f32(g).reshape(1), d_scaled, 0.0, d_emb)
        return d_emb, None, None
