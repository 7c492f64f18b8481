"""Host-side (Python) cost of enqueuing one training step: cProfile over N steps with the id plumbing prefetched (no host
syncs inside the step), so the numbers are pure launch-path overhead.   python tools/profile_host.py [--model cl]"""
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
from xnrs_b200.models import make_model  # noqa: E402
from xnrs_b200.training import ContrastiveRankingTrainer  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--model', default='cl')
ap.add_argument('--steps', type=int, default=30)
args = ap.parse_args()
dev = torch.device('cuda', 0)
K.set_precision('tf32x3')
cfg = bench.MODEL_CFGS[args.model]
cat = syn.make_catalogue(bench.N_NEWS, bench.SEQ_LEN, bench.VOCAB, 768, seed=0)
store = TitleStore(cat.token_table.to(dev), cat.title_tokens.to(dev))
torch.manual_seed(0)
tr = ContrastiveRankingTrainer(dict(cfg, device=str(dev)), make_model(cfg))
tr.model.train()
batches = [syn.index_batch(store, cat, syn.make_train_batch(bench.N_NEWS, 1024, bench.HIST_LEN, seed=1000 + i), dev) for i in range(4)]
for i in range(5):
    tr._train_step(batches[i % 4])
torch.cuda.synchronize()


def run(n):
    tr.prefetch(batches[0])
    for i in range(n):
        tr._train_step(batches[i % 4])
        tr.prefetch(batches[(i + 1) % 4])


t0 = time.perf_counter()
run(args.steps)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f'host enqueue {1e3 * (t1 - t0) / args.steps:.3f} ms/step, drained after another {1e3 * (t2 - t1):.3f} ms')
pr = cProfile.Profile()
pr.enable()
run(args.steps)
pr.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr)
st.sort_stats('tottime').print_stats(28)
