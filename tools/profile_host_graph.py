"""Host-side cost of one END-TO-END step of the graphed CL loop (bench.py's e2e region): pinned int32 ids -> device, id plumbing
prefetched on the side stream, one CUDA-graph replay, async read-back of the loss.  Prints the host time per step with the GPU
idle-waiting excluded (no sync inside the loop) and a cProfile of the loop.
    python tools/profile_host_graph.py [--precision bf16x3] [--steps 200]"""
import argparse
import cProfile
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from xnrs_b200 import kernels as K  # noqa: E402
from xnrs_b200 import synthetic as syn  # noqa: E402
from xnrs_b200.data import TitleStore  # noqa: E402
from xnrs_b200.distributed import DataParallelTrainer  # noqa: E402
from xnrs_b200.graphs import GraphedStep  # noqa: E402
from xnrs_b200.models import make_model  # noqa: E402
from xnrs_b200.training import ContrastiveRankingTrainer  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--precision', default='bf16x3')
ap.add_argument('--steps', type=int, default=200)
args = ap.parse_args()
dev = torch.device('cuda', 0)
K.set_precision(args.precision)
cfg = bench.MODEL_CFGS['cl']
cat = syn.make_catalogue(bench.N_NEWS, bench.SEQ_LEN, bench.VOCAB, 768, seed=0)
store = TitleStore(cat.token_table.to(dev), cat.title_tokens.to(dev))
torch.manual_seed(0)
tr = ContrastiveRankingTrainer(dict(cfg, device=str(dev)), make_model(cfg), graph_safe=True)
tr.model.train()
dp = DataParallelTrainer(tr)
stepper = GraphedStep(dp)
raws = [syn.make_train_batch(bench.N_NEWS, 1024, bench.HIST_LEN, seed=1000 + i) for i in range(8)]
pinned = [{k: v.pin_memory() for k, v in r.items()} for r in raws]


def h2d(i):
    b = syn.index_batch(store, cat, pinned[i % 8], dev)
    return b, torch.cuda.current_stream().record_event()


for i in range(24):
    stepper.step(h2d(i)[0])
torch.cuda.synchronize()
loss_host = [torch.zeros(1).pin_memory() for _ in range(2)]


def run(n):
    cur = h2d(0)
    for i in range(n):
        nxt = h2d(i + 1)
        out = stepper.step(cur[0])
        loss_host[i & 1].copy_(out['loss'].reshape(1), non_blocking=True)
        dp.prefetch(nxt[0], after=nxt[1])
        cur = nxt


t0 = time.perf_counter()
run(args.steps)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f'host enqueue {1e3 * (t1 - t0) / args.steps:.3f} ms/step; drained after another {1e3 * (t2 - t1):.3f} ms '
      f'(host-bound if that is ~0); replays {stepper.replays}')
pr = cProfile.Profile()
pr.enable()
run(args.steps)
pr.disable()
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats('cumulative').print_stats(35)
