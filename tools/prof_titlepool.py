"""Small driver for ncu captures of the token-level kernels of the CL step at bench shapes (T ~ 153.6k real tokens, 8.7k titles):
the fused title-pooling forward (gather -> fc1 -> tanh -> logit -> exp -> per-title sums), the fc1 weight gradient with the
fused table gather, their dense counterparts, the title-level pooling backward and two title-level GEMMs.  Three launches
each, in that order.
    ncu --set full --clock-control none --import-source on -k regex:'gemm_tc|pool_bwd' -o gpurun_out/prof python tools/prof_titlepool.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from xnrs_b200 import kernels as K  # noqa: E402
from xnrs_b200 import synthetic as syn  # noqa: E402
from xnrs_b200.data import TitleStore, plan_titles  # noqa: E402


def main():
    dev = 'cuda'
    K.set_precision('tf32x3')
    cat = syn.make_catalogue(65238, 30, 100_000, 768, seed=0)
    store = TitleStore(cat.token_table.to(dev), cat.title_tokens.to(dev))
    raw = syn.make_train_batch(65238, 1024, 50, seed=1000)
    ids = torch.cat([raw['hist_ids'].reshape(-1), raw['cand_ids'].reshape(-1)]).to(dev)
    plan = plan_titles(store, ids, True, True).acquire()
    R, T = plan.uniq.numel(), plan.rows.numel()
    g = torch.Generator().manual_seed(0)
    w1 = (torch.randn(256, 768, generator=g) / 768 ** 0.5).to(dev)
    b1 = (torch.randn(256, generator=g) * 0.1).to(dev)
    w2 = (torch.randn(256, generator=g) / 16).to(dev)
    b2 = torch.zeros(1, device=dev)
    hid = torch.empty(T, 256, device=dev)
    e, attn, zsum = torch.empty(T, device=dev), torch.empty(T, device=dev), torch.empty(R, device=dev)
    pooled = torch.empty(R, 768, device=dev)
    dhid = torch.randn(T, 256, device=dev)
    dw = torch.zeros(256, 768, device=dev)
    x = K.gather_rows(store.token_table, plan.rows)
    print(f'titles {R}, token rows {T}')
    for _ in range(3):
        K.call('xnrs_titlepool_fwd', K._mat(store.token_table), 768, plan.rows, plan.tix, plan.seg, T, R, 768, 256, w1, b1, w2, b2, 1, hid, e, zsum,
               attn, pooled)
    for _ in range(3):
        K.gemm(dhid, store.token_table, trans_a=True, out=dw, accumulate=True, b_rows=plan.rows)
    for _ in range(3):
        K.gemm(dhid, x, trans_a=True, out=dw, accumulate=True)
    for _ in range(3):
        K.gemm(x, w1, trans_b=True, bias=b1, act=K.ACT_TANH, out=hid)
    for _ in range(3):
        K.gemm(store.token_table, w1, trans_b=True, bias=b1, act=K.ACT_TANH, out=hid, a_rows=plan.rows)
    # pooling backward at title level (pool_bwd2_kernel) and one title-level GEMM (TMA-store epilogue, split-K reductions)
    d_pooled = torch.randn(R, 768, device=dev)
    d_hid = torch.empty(T, 256, device=dev)
    d_w2, d_b2, d_b1 = torch.zeros(256, device=dev), torch.zeros(1, device=dev), torch.zeros(256, device=dev)
    for _ in range(3):
        K.call('xnrs_addpool_bwd', K._mat(store.token_table), plan.rows, None, hid, w2, attn, d_pooled, None, plan.seg, R, 30, 768, 256, T,
               d_hid, d_w2, d_b2, None, d_b1)
    a_small, w_small, o_small = torch.randn(R, 256, device=dev), torch.randn(256, 256, device=dev), torch.empty(R, 256, device=dev)
    for _ in range(3):
        K.gemm(a_small, w_small, trans_b=True, bias=b1, out=o_small)
    g_small = torch.zeros(256, 256, device=dev)
    for _ in range(3):
        K.gemm(o_small, a_small, trans_a=True, out=g_small, accumulate=True)
    # the same token-level launches in the 3xBF16 pre-split form (bench default): fused pooling forward, pooling backward writing
    # d_hid as two bf16 planes, gathered weight gradient
    th, tl = K.bf16_split_twin(store.token_table, True)
    tbh, tbl = K.bf16_split_twin(store.token_table, False)
    w1h, w1l = K.split_bf16(w1, fp16=True)
    dh_hi, dh_lo = torch.empty(T, 256, device=dev, dtype=torch.bfloat16), torch.empty(T, 256, device=dev, dtype=torch.bfloat16)
    for _ in range(3):
        K.call('xnrs_titlepool_fwd_bf16x3', th, tl, 768, plan.rows, plan.tix, plan.seg, T, R, 768, 256, w1h, w1l, 1, b1, w2, b2,
               K._mat(store.token_table), 768, hid, e, zsum, attn, pooled)
    for _ in range(3):
        K.call('xnrs_addpool_bwd_split', K._mat(store.token_table), plan.rows, hid, w2, attn, d_pooled, plan.seg, R, 30, 768, 256, T,
               dh_hi, dh_lo, d_w2, d_b2, d_b1)
    for _ in range(3):
        K.gemm_bf16x3(dh_hi, dh_lo, tbh, tbl, trans_a=True, b_rows=plan.rows, out=dw, accumulate=True)
    torch.cuda.synchronize()
    print('done')


if __name__ == '__main__':
    main()
