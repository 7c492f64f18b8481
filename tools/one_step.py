"""Profiling target: warm up, then run N training steps (or one evaluation pass) of the bench workload between
cudaProfilerStart/Stop, so `ncu --profile-from-start off` sees exactly those launches.

    python tools/one_step.py [--model cl|nrms|naml|lstur|npa] [--workload train|eval] [--steps 1] [--precision tf32x3]
    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file launches.csv \
        python tools/one_step.py --model cl
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from xnrs_b200 import kernels as K  # noqa: E402
from xnrs_b200 import synthetic as syn  # noqa: E402
from xnrs_b200.data import TitleStore  # noqa: E402
from xnrs_b200.models import make_model  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--model', default='cl', choices=list(bench.MODEL_CFGS))
    ap.add_argument('--workload', default='train', choices=['train', 'eval'])
    ap.add_argument('--steps', type=int, default=1)
    ap.add_argument('--warmup', type=int, default=4)
    ap.add_argument('--batch', type=int, default=1024)
    ap.add_argument('--precision', default='tf32x3')
    ap.add_argument('--eval-impressions', type=int, default=65536)
    args = ap.parse_args()
    dev = torch.device('cuda', 0)
    torch.cuda.set_device(dev)
    K.set_precision(args.precision)
    rt = torch.cuda.cudart()
    if args.workload == 'eval':
        from xnrs_b200.evaluation import CatalogueEvaluator
        n_news = 160_000
        cat = syn.make_catalogue(n_news, bench.SEQ_LEN, bench.VOCAB, 768, seed=0)
        imp = {k: v.to(dev) for k, v in syn.make_eval_impressions(n_news, args.eval_impressions, bench.HIST_LEN, seed=1).items()}
        store = TitleStore(cat.token_table.to(dev), cat.title_tokens.to(dev))
        torch.manual_seed(0)
        ev = CatalogueEvaluator(make_model(bench.CL_CFG).to(dev).eval(), store, news_chunk=16384, impression_chunk=16384)
        ev.encode_catalogue()
        ev.evaluate(imp)
        torch.cuda.synchronize()
        rt.cudaProfilerStart()
        ev.news_vecs = None
        ev.encode_catalogue()
        print(ev.evaluate(imp))
        torch.cuda.synchronize()
        rt.cudaProfilerStop()
        return
    from xnrs_b200.training import ContrastiveRankingTrainer, MSERankingTrainer
    cfg = bench.MODEL_CFGS[args.model]
    cat = syn.make_catalogue(bench.N_NEWS, bench.SEQ_LEN, bench.VOCAB, 768, seed=0, with_abstract=(args.model == 'naml'))
    store = TitleStore(cat.token_table.to(dev), cat.title_tokens.to(dev))
    astore = TitleStore(store.token_table, cat.abstract_tokens.to(dev)) if args.model == 'naml' else None
    torch.manual_seed(0)
    trainer = (MSERankingTrainer if args.model == 'npa' else ContrastiveRankingTrainer)(dict(cfg, device=str(dev)), make_model(cfg))
    trainer.model.train()
    batches = [syn.index_batch(store, cat, syn.make_train_batch(bench.N_NEWS, args.batch, bench.HIST_LEN, seed=1000 + i), dev,
                               abstract_store=astore) for i in range(4)]
    for i in range(args.warmup):
        trainer._train_step(batches[i % 4])
    torch.cuda.synchronize()
    rt.cudaProfilerStart()
    for i in range(args.steps):
        out = trainer._train_step(batches[i % 4])
    torch.cuda.synchronize()
    rt.cudaProfilerStop()
    print('loss', float(out['loss']))


if __name__ == '__main__':
    main()
