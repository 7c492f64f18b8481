"""Multi-GPU check (torchrun, NCCL): (1) N ranks each take 1/N of a fixed global index batch; one data-parallel CL step must
reproduce the single-process step on the whole batch (gradient, global-batch InfoNCE value, updated parameters); (2) the same
data-parallel steps replayed as CUDA graphs WITH the NCCL collectives captured (GraphedStep) track the eager ones; (3) the
sharded full-catalogue evaluation (catalogue rows and impressions split over ranks, all-gather + all-reduce) reproduces
the single-process epoch metrics.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/check_dp_gpu.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import bench  # noqa: E402
from xnrs_b200 import kernels as K  # noqa: E402
from xnrs_b200 import synthetic as syn  # noqa: E402
from xnrs_b200.data import TitleStore  # noqa: E402
from xnrs_b200.distributed import DataParallelTrainer, shard_range  # noqa: E402
from xnrs_b200.models import make_model  # noqa: E402
from xnrs_b200.training import ContrastiveRankingTrainer  # noqa: E402


def main():
    rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    dist.init_process_group('nccl', device_id=dev)
    K.set_precision('tf32x3')
    cfg = dict(bench.CL_CFG, lr=1e-3, contrastive_lambda=0.1, device=str(dev))
    cat = syn.make_catalogue(2000, bench.SEQ_LEN, 5000, 768, seed=0)
    store = TitleStore(cat.token_table.to(dev), cat.title_tokens.to(dev))
    B = 64 * world
    raw = syn.make_train_batch(2000, B, bench.HIST_LEN, seed=7)

    def trainer():
        torch.manual_seed(0)
        tr = ContrastiveRankingTrainer(cfg, make_model(cfg))
        tr.model.eval()
        return tr

    full = trainer()
    full.optimizer.zero_grad()
    total, _, l_cl, _ = full.losses(syn.index_batch(store, cat, raw, dev))
    total.backward()
    want_g = full.optimizer.flat_g.clone()
    full.optimizer.step()

    tr = trainer()
    dp = DataParallelTrainer(tr)
    lo, hi = shard_range(B, rank, world)
    out = dp.train_step(syn.index_batch(store, cat, {k: v[lo:hi] for k, v in raw.items()}, dev))
    torch.cuda.synchronize()
    scale = float(want_g.abs().max())
    eg = float((tr.optimizer.flat_g / world - want_g).abs().max()) / scale
    ecl = abs(float(out['loss_cl']) - float(l_cl)) / max(abs(float(l_cl)), 1e-9)
    big = want_g.abs() > 1e-3 * scale
    ep = float((tr.optimizer.flat_p - full.optimizer.flat_p)[big].abs().max())
    ok = eg < 1e-4 and ecl < 1e-4 and ep < 2e-5
    dp.check_peers()
    print(f'rank {rank}/{world}: grad err {eg:.2e}, InfoNCE err {ecl:.2e}, param err {ep:.2e} -> {"OK" if ok else "FAIL"} '
          f'[InfoNCE exchange: {dp.infonce_exchange}]', flush=True)

    # (2) six more data-parallel steps, eager vs CUDA-graph replay (collectives inside the graph)
    from xnrs_b200.graphs import GraphedStep
    raws = [syn.make_train_batch(2000, B, bench.HIST_LEN, seed=20 + (i % 2)) for i in range(6)]
    finals = []
    for graphed in (False, True):
        torch.manual_seed(0)
        t2 = ContrastiveRankingTrainer(cfg, make_model(cfg), graph_safe=True)
        t2.model.eval()
        d2 = DataParallelTrainer(t2)
        st = GraphedStep(d2) if graphed else None
        losses = []
        for r_ in raws:
            b_ = syn.index_batch(store, cat, {k: v[lo:hi] for k, v in r_.items()}, dev)
            o_ = st.step(b_) if graphed else d2.train_step(b_)
            losses.append(float(o_['loss']))
        finals.append((losses, t2.optimizer.flat_p.clone(), st))
    dl = max(abs(a - b) / max(abs(a), 1e-9) for a, b in zip(finals[0][0], finals[1][0]))
    dp_ = float((finals[0][1] - finals[1][1]).abs().max())
    # parameters: Adam turns the sign of a ~0 gradient component (fp32 atomic order) into a whole +-lr step, so two runs of
    # the SAME eager code differ by a few lr as well; a wrong replay would show up in the losses and as O(1) differences
    ok2 = dl < 1e-4 and dp_ < 3 * cfg['lr'] and finals[1][2].replays >= 2
    print(f'rank {rank}/{world}: graph-vs-eager loss diff {dl:.2e}, param diff {dp_:.2e}, replays {finals[1][2].replays} -> '
          f'{"OK" if ok2 else "FAIL"}', flush=True)

    # (3) sharded evaluation == single-process evaluation
    from xnrs_b200.evaluation import CatalogueEvaluator
    import xnrs_b200.evaluation as EV
    imp = syn.make_eval_impressions(2000, 3001, bench.HIST_LEN, seed=3)
    torch.manual_seed(0)
    model = make_model(cfg).to(dev).eval()
    ev = CatalogueEvaluator(model, store, news_chunk=512, impression_chunk=700)
    sharded = ev.evaluate(imp)
    w_, r_ = EV.world, EV.rank
    EV.world, EV.rank = (lambda: 1), (lambda: 0)            # the same evaluator as one process over everything
    try:
        ev.news_vecs = None
        single = ev.evaluate(imp)
    finally:
        EV.world, EV.rank = w_, r_
    de = max(abs(sharded[k] - single[k]) for k in ('auc', 'rr', 'ndcg@5', 'ndcg@10', 'ctr@1', 'ctr@10'))
    ok3 = de < 1e-4 and sharded['impressions'] == single['impressions']
    print(f'rank {rank}/{world}: sharded-vs-single evaluation metric diff {de:.2e} over {single["impressions"]} impressions -> '
          f'{"OK" if ok3 else "FAIL"}', flush=True)
    ok = ok and ok2 and ok3
    replays = finals[1][2].replays
    finals.clear()                      # the captured graphs hold NCCL work: release them before the communicator goes away
    del st, d2, t2
    dp.check_peers()
    from xnrs_b200 import distributed as D
    D._peer_cache.clear()
    import gc
    gc.collect()
    torch.cuda.synchronize()
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == '__main__':
    main()
