"""Multi-GPU check (torchrun, NCCL): N ranks each take 1/N of a fixed global index batch; one data-parallel CL step must
reproduce the single-process step on the whole batch (gradient, global-batch InfoNCE value, updated parameters).
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
    print(f'rank {rank}/{world}: grad err {eg:.2e}, InfoNCE err {ecl:.2e}, param err {ep:.2e} -> {"OK" if ok else "FAIL"}', flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == '__main__':
    main()
