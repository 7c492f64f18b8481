"""GPU-side bar (last row): the reference's training step in STOCK PyTorch eager on the same B200 — the
torch restatement of the reference modules (oracle/xnrs_oracle.py: nn.Linear / softmax / bmm semantics, two history
forwards like training.py:402-431, autograd, Adam) moved to the GPU with dense reference-format batches, cuBLAS fp32
(TF32 off) and, for context, TF32 on.  A measurement tool only: nothing in the product imports it.

    python tools/bench_eager_gpu.py [--batch 256] [--steps 5]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from oracle import xnrs_oracle as O  # noqa: E402
from xnrs_b200 import synthetic as syn  # noqa: E402
from xnrs_b200.models import make_model  # noqa: E402


def to_dev(x, dev):
    if isinstance(x, torch.Tensor):
        return x.to(dev)
    if isinstance(x, dict):
        return {k: to_dev(v, dev) for k, v in x.items()}
    if isinstance(x, (tuple, list)):
        return type(x)(to_dev(v, dev) for v in x)
    return x


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--batch', type=int, default=256)
    ap.add_argument('--steps', type=int, default=5)
    args = ap.parse_args()
    dev = torch.device('cuda', 0)
    cfg = bench.CL_CFG
    cat = syn.make_catalogue(bench.N_NEWS, bench.SEQ_LEN, bench.VOCAB, 768, seed=0)
    torch.manual_seed(0)
    P = {k: v.detach().clone().to(dev).requires_grad_(True) for k, v in make_model(cfg).state_dict().items()}
    state = {k: (torch.zeros_like(v), torch.zeros_like(v)) for k, v in P.items()}
    batches = [to_dev(syn.dense_batch(cat, syn.make_train_batch(bench.N_NEWS, args.batch, bench.HIST_LEN, seed=100 + i)), dev)
               for i in range(2)]
    n = [0]

    def step():
        b = batches[n[0] % 2]
        n[0] += 1
        for v in P.values():
            v.grad = None
        scores = O.parent_forward(P, b)
        u = O.parent_user_embeddings(P, b)
        loss, _, _ = O.contrastive_train_loss(scores, b['targets'], u, b['main_theme'].long(), cfg['contrastive_temperature'],
                                              cfg['contrastive_lambda'])
        loss.backward()
        with torch.no_grad():
            for k, v in P.items():
                if v.grad is not None:
                    O.adam_step(v, v.grad, state[k][0], state[k][1], n[0], cfg['lr'])
        return loss

    for tf32 in (False, True):
        torch.backends.cuda.matmul.allow_tf32 = tf32
        torch.backends.cudnn.allow_tf32 = tf32
        for _ in range(2):
            step()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(args.steps):
            loss = step()
        e.record()
        torch.cuda.synchronize()
        ms = s.elapsed_time(e) / args.steps
        print(json.dumps({'impl': 'stock PyTorch eager on the B200 (torch restatement of the reference step, dense batches resident in HBM)',
                          'matmul': 'cuBLAS TF32' if tf32 else 'cuBLAS fp32', 'batch': args.batch, 'ms_per_step': round(ms, 3),
                          'impressions_per_s': round(args.batch / (ms * 1e-3), 1), 'loss': round(float(loss), 6)}), flush=True)


if __name__ == '__main__':
    main()
