#!/usr/bin/env python
"""Headline benchmark: CL bi-encoder (config/mind_small_CL.yml, `model: standard`) TRAINING throughput in
impressions/s on synthetic MIND-shaped data (BASELINE.json configs[1]; north-star shapes S=30, H=50, 1:4).

One "step" = one full ContrastiveRankingTrainer step over one batch of impressions: index gather from the
device-resident token table, title + user encoders, fused dot-score + MSE(ReLU), supervised InfoNCE,
backward through everything, Adam.  Contract: see the task statement (`value` = inputs resident in HBM,
`e2e` = host index buffers + H2D + D2H loss read inside the timed region, `roofline` for the dominant
kernel timed live with CUDA events, `cpu_baseline` = the CPU oracle on a bounded sample).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--batch B] [--precision fp32]
N>1 is launched by torchrun (one rank per GPU, NCCL); rank 0 prints ONE JSON line.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

CL_CFG = dict(model='standard', scoring='dot', text_features=['title_emb'], catg_features=[], title_emb_dim=256,
              total_emb_dim=256, d_backbone=768, p_dropout=0., bias=False, n_negatives=4, lr=1e-4,
              contrastive_temperature=0.08, contrastive_lambda=0.01)          # config/mind_small_CL.yml
# the other BASELINE.json configs (secondary lines: `--model nrms|naml|lstur|npa`), hyper-parameters from config/mind_small_*.yml
_COMMON = dict(scoring='dot', title_emb_dim=256, d_backbone=768, p_dropout=0., bias=False, n_negatives=4, lr=1e-4,
               n_categories=19, n_subcategories=300, n_users=703_789, cat_emb_dim=16, sub_emb_dim=16, n_heads=16,
               contrastive_temperature=0.08, contrastive_lambda=0.1)
MODEL_CFGS = {
    'cl': CL_CFG,
    'nrms': dict(_COMMON, model='NRMS', text_features=['title_emb'], catg_features=[], total_emb_dim=256),
    'naml': dict(_COMMON, model='NAML', text_features=['title_emb', 'abstract_emb'], total_emb_dim=256,
                 catg_features=['category_index', 'subcategory_index']),
    'lstur': dict(_COMMON, model='LSTUR', text_features=['title_emb'], catg_features=['category_index'], total_emb_dim=272,
                  long_term_method='embedding', long_short_term_method='con', p_user_dropout=0.07, st_hist_len=50),
    'npa': dict(_COMMON, model='NPA', text_features=['title_emb'], catg_features=[], total_emb_dim=256, user_emb_dim=64),
}
SEQ_LEN, HIST_LEN, N_NEWS, VOCAB = 30, 50, 65_238, 100_000                     # SURVEY §8(d) north-star shapes
METRIC, UNIT = 'train impressions/s', 'impressions/s'


def workload_name(batch, model='cl'):
    if model != 'cl':
        return (f'{model.upper()} (config/mind_small_{model.upper()}.yml) train step fwd+loss+bwd+Adam; synthetic MIND-shaped: '
                f'S={SEQ_LEN}, H={HIST_LEN}, 1:4 negatives, D=768, {N_NEWS} news, {VOCAB}-row token table; '
                f'{batch} impressions/GPU/step')
    return (f'CL bi-encoder (mind_small_CL.yml, model=standard) train step fwd+loss(MSE.ReLU + 0.01*InfoNCE)+bwd+Adam; '
            f'synthetic MIND-shaped: S={SEQ_LEN} tokens, H={HIST_LEN} history, 1:4 negatives, D=768, '
            f'{N_NEWS} news, {VOCAB}-row token table; {batch} impressions/GPU/step')


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json'))), 'measured'
    except Exception:
        return {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0, 'bf16_tflops_sustained': 1400.0}, 'fallback'


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def __enter__(self):
        # nvidia-smi takes ~1 s (and driver locks: a 100 ms stall of the GPU work queue was measured) to initialise, so it
        # is started and allowed to print its first sample BEFORE the timed region; inside the region it only polls (every
        # 200 ms like the recipe's clocks line: polling at 25 ms was measured to stall short timed regions)
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits',
                                          '-lms', '200', '-i', str(self.index)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=lambda: self.rows.extend(self.proc.stdout), daemon=True)
            self.t.start()
            t0 = time.perf_counter()
            while not self.rows and time.perf_counter() - t0 < 5.0:
                time.sleep(0.02)
            torch.cuda.synchronize()
        except Exception:
            self.proc = None
        return self

    def __exit__(self, *a):
        if self.proc:
            self.proc.terminate()
            self.t.join(timeout=2)

    def summary(self):
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            f = [x.strip() for x in r.split(',')]
            if len(f) < 6 or not f[0].isdigit():
                continue
            sm.append(int(f[0]))
            mx.append(int(f[1]))
            for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), f[2:6]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        if not sm:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': [], 'samples': 0}
        sm.sort()
        return {'sm_mhz': sm[len(sm) // 2], 'sm_max_mhz': max(mx), 'reasons': sorted(reasons), 'samples': len(sm)}


def oracle_step_fn(batch_size, threads):
    """the CPU path: the oracle's restatement of ContrastiveRankingTrainer._train_step (two history forwards like
    the reference, training.py:402-431) + autograd + Adam, on dense host batches built from the same ids."""
    from oracle import xnrs_oracle as O
    from xnrs_b200 import synthetic as syn
    from xnrs_b200.models import make_model
    torch.set_num_threads(threads)
    cat = syn.make_catalogue(N_NEWS, SEQ_LEN, VOCAB, 768, seed=0)
    torch.manual_seed(0)
    P = {k: v.detach().clone().requires_grad_(True) for k, v in make_model(CL_CFG).state_dict().items()}
    state = {k: (torch.zeros_like(v), torch.zeros_like(v)) for k, v in P.items()}
    batches = [syn.dense_batch(cat, syn.make_train_batch(N_NEWS, batch_size, HIST_LEN, seed=100 + i)) for i in range(2)]
    counter = [0]

    def step():
        b = batches[counter[0] % len(batches)]
        counter[0] += 1
        for v in P.values():
            v.grad = None
        scores = O.parent_forward(P, b)
        u = O.parent_user_embeddings(P, b)
        loss, _, _ = O.contrastive_train_loss(scores, b['targets'], u, b['main_theme'].long(),
                                              CL_CFG['contrastive_temperature'], CL_CFG['contrastive_lambda'])
        loss.backward()
        with torch.no_grad():
            for k, v in P.items():
                if v.grad is not None:
                    O.adam_step(v, v.grad, state[k][0], state[k][1], counter[0], CL_CFG['lr'])
        return float(loss.detach())
    return step


def time_cpu(batch_size, steps, warmup, threads):
    step = oracle_step_fn(batch_size, threads)
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    return batch_size * steps / dt, dt / steps


def eval_cpu_baseline(cat, imp, model_sd, n_sample, threads):
    """the reference evaluation procedure (training.py:194-243: ONE impression per step — full re-encode of the history
    and every candidate, scores to the host, numpy metrics) restated by the oracle, on a bounded sample of the same
    impressions -> (impressions/s, seconds)"""
    from oracle import xnrs_oracle as O
    from xnrs_b200 import synthetic as syn
    torch.set_num_threads(threads)
    P = {k: v.detach().cpu().clone() for k, v in model_sd.items()}
    t0 = time.perf_counter()
    with torch.no_grad():
        for i in range(n_sample):
            a, b = int(imp['offsets'][i]), int(imp['offsets'][i + 1])
            raw = {'hist_ids': imp['hist_ids'][i:i + 1], 'cand_ids': imp['cand_ids'][a:b][None, :],
                   'targets': imp['targets'][a:b][None, :, None], 'user_index': imp['user_index'][i:i + 1],
                   'main_theme': torch.zeros(1, dtype=torch.int32)}
            scores = torch.relu(O.parent_forward(P, syn.dense_batch(cat, raw))).reshape(-1)
            O.impression_metrics(imp['targets'][a:b].numpy(), scores.numpy())
    dt = time.perf_counter() - t0
    return n_sample / dt, dt


def run_reference(args):
    """--impl reference: the reference algorithm's CPU path (oracle port; the Python reference cannot travel to the
    GPU box) with all host threads, same metric/config, each step a bounded sample of the workload."""
    if int(os.environ.get('RANK', '0')) != 0:
        return
    threads = os.cpu_count() or 1
    bs = args.ref_batch
    value, per_step = time_cpu(bs, args.steps, args.warmup, threads)
    sample = f'{bs} impressions/step x {args.steps} steps (bounded sample of the {args.batch}/GPU workload)'
    print(json.dumps({
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': per_step * 1e3, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': workload_name(args.batch), 'reference_sample': sample},
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': threads, 'kind': 'port', 'sample': sample},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }))


def run_eval(args):
    """--workload eval: BASELINE.json configs[4] — MIND-large-shaped full-catalogue evaluation (160k news, 376,471
    impressions, ~37 candidates each) with the CL/standard model: encode the catalogue once (sharded over ranks,
    all-gathered), then per-impression user encoding + candidate scoring + AUC/MRR/nDCG (strong scaling)."""
    import torch.distributed as dist
    from xnrs_b200 import kernels as K
    from xnrs_b200 import synthetic as syn
    from xnrs_b200.data import TitleStore
    from xnrs_b200.evaluation import CatalogueEvaluator
    from xnrs_b200.models import make_model

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    K.set_precision(args.precision)
    n_news, n_imp = 160_000, args.eval_impressions
    cat = syn.make_catalogue(n_news, SEQ_LEN, VOCAB, 768, seed=0)
    imp = syn.make_eval_impressions(n_news, n_imp, HIST_LEN, seed=1)
    store = TitleStore(cat.token_table.to(dev), cat.title_tokens.to(dev))
    torch.manual_seed(0)
    model = make_model(CL_CFG).to(dev).eval()
    ev = CatalogueEvaluator(model, store, news_chunk=16384, impression_chunk=16384)
    imp_dev = {k: v.to(dev) for k, v in imp.items()}
    imp_pin = {k: v.pin_memory() for k, v in imp.items()}

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def one_pass(src):
        ev.news_vecs = None
        t0, t1, t2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        t0.record()
        ev.encode_catalogue()
        t1.record()
        out = ev.evaluate(src)
        t2.record()
        return out, t0, t1, t2

    for _ in range(max(1, args.warmup // 3)):
        one_pass(imp_dev)
    sync_all()
    n0 = K.launch_count()
    passes = max(1, min(args.steps, 5))               # a step = one full pass (catalogue encode + every impression)
    with ClockSampler(local) as clocks:
        sync_all()
        evs = [one_pass(imp_dev) for _ in range(passes)]
        sync_all()
    out = evs[-1][0]
    launches = K.launch_count() - n0
    tt = torch.tensor([sum(e[1].elapsed_time(e[3]) for e in evs) / passes, sum(e[1].elapsed_time(e[2]) for e in evs) / passes,
                       sum(e[2].elapsed_time(e[3]) for e in evs) / passes], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    total_ms, enc_ms, score_ms = (float(x) for x in tt)
    sync_all()
    w0 = time.perf_counter()
    for _ in range(passes):
        out2, *_ = one_pass(imp_pin)                  # host CSR buffers -> H2D inside the timed region, means read back
    sync_all()
    e2e_s = (time.perf_counter() - w0) / passes
    # one more pass with every entry point bracketed by CUDA events on its stream: per-kernel durations
    recs = []

    def hook(name, args_):
        s_, e_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        recs.append((name, s_, e_, args_))
        return s_, e_

    K.set_event_hook(hook)
    one_pass(imp_dev)
    sync_all()
    K.set_event_hook(None)
    per_entry, gemm_flop = {}, 0.0
    for name, s_, e_, a_ in recs:
        per_entry[name] = per_entry.get(name, 0.0) + s_.elapsed_time(e_)
        if name == 'xnrs_gemm':
            gemm_flop += 2.0 * a_[2] * a_[3] * a_[4]
    pk, pk_kind = peaks()
    n_cand = int(imp['offsets'][-1])
    # algorithmic bytes of the score+rank kernel: one T-wide fp32 vector per candidate + the user vector + ids/targets
    score_bytes = n_cand * (256 * 4 + 4 + 4 + 4) + n_imp * (256 * 4 + 8 + 6 * 8)
    res = {
        'metric': 'eval scored impressions/s', 'value': n_imp / (total_ms * 1e-3), 'unit': UNIT, 'n_gpus': world,
        'steps': passes, 'warmup': max(1, args.warmup // 3), 'ms_per_step': total_ms, 'higher_is_better': True,
        'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f32 (3xTF32)' if args.precision == 'tf32x3' else args.precision,
        'data': 'synthetic',
        'config': {'workload': f'MIND-large-shaped full-catalogue eval, model=standard (mind_standard.yml): {n_news} news '
                               f'encoded once, {n_imp} impressions x ~37 candidates, H={HIST_LEN}, S={SEQ_LEN}; one step = '
                               f'catalogue encode + all impressions', 'parallelism': f'news+impression sharding x{world}',
                   'l2': 'news vectors 164 MB + 14M candidate gathers exceed L2', 'catalogue_encode_ms': enc_ms,
                   'impression_phase_ms': score_ms, 'metrics': {k: out[k] for k in ('auc', 'rr', 'ndcg@5', 'ndcg@10')}},
        'e2e': {'value': n_imp / e2e_s, 'unit': UNIT,
                'h2d_bytes_per_step': sum(v.numel() * v.element_size() for v in imp.values()), 'd2h_bytes_per_step': 56},
        'gpu_launches': launches,
        # dominant kernel of the pass: the tcgen05 GEMMs of the catalogue encode (title fc1 + heads + per-article pooling
        # logits), FLOPs = sum of 2MNK over the launches / their CUDA-event time
        'roofline': {'kernel': 'tcgen05 GEMMs of the pass (gemm_tc2_kernel / gemm_tc_kernel: catalogue fc1 with tanh epilogue, heads, '
                               'per-article pooling logits)', 'bound': 'tensor',
                     'achieved': gemm_flop / (per_entry.get('xnrs_gemm', 0.0) * 1e-3) / 1e12 if per_entry.get('xnrs_gemm') else None,
                     'peak': pk['bf16_tflops_sustained'], 'unit': 'TFLOP/s',
                     'frac': (gemm_flop / (per_entry['xnrs_gemm'] * 1e-3) / 1e12 / pk['bf16_tflops_sustained']) if per_entry.get('xnrs_gemm') else None,
                     'traffic': None,
                     'peak_source': f'{pk_kind} (sustained bf16 GEMM). 3xTF32 (fp32-accurate) has 1/6 of this peak as its own ceiling',
                     'per_entry_point_ms': {k: round(v, 3) for k, v in sorted(per_entry.items(), key=lambda kv: -kv[1])}},
        # the HBM-bound scoring kernel: algorithmic bytes (one T-wide vector per candidate + ids/targets) / its event time
        'roofline_scoring': {'kernel': 'eval_impressions_kernel (gather + dot + segmented rank sort + metrics)', 'bound': 'hbm',
                             'achieved': score_bytes / (per_entry.get('xnrs_eval_impressions', score_ms) * 1e-3) / 1e9,
                             'peak': pk['hbm_gbs'], 'unit': 'GB/s',
                             'frac': score_bytes / (per_entry.get('xnrs_eval_impressions', score_ms) * 1e-3) / 1e9 / pk['hbm_gbs'],
                             'peak_source': pk_kind},
        'clocks': clocks.summary(),
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        n_s = min(4000, n_imp)                                # ~6 s of host work on the box's 16 cores
        threads = os.cpu_count() or 1
        try:
            v, dt = eval_cpu_baseline(cat, imp, model.state_dict(), n_s, threads)
            res['cpu_baseline'] = {'value': v, 'unit': UNIT, 'cores': threads, 'kind': 'port',
                                   'sample': f'{n_s} of the {n_imp} impressions, one per step with a full re-encode of history and candidates '
                                             f'+ numpy metrics like training.py:194-243 (oracle, torch CPU fp32), {dt:.1f} s'}
        except Exception as exc:                              # the baseline must never take the measurement down with it
            res['cpu_baseline'] = {'value': None, 'unit': UNIT, 'cores': threads, 'kind': 'port', 'sample': f'failed: {exc!r}'}
    if rank == 0:
        print(json.dumps(res))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--batch', type=int, default=1024, help='impressions per GPU per step')
    ap.add_argument('--ref-batch', type=int, default=64, help='impressions per CPU step (reference arm / cpu_baseline)')
    ap.add_argument('--precision', default='tf32x3', choices=['fp32', 'tf32x3', 'tf32', 'bf16'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--workload', default='train', choices=['train', 'eval'])
    ap.add_argument('--model', default='cl', choices=list(MODEL_CFGS), help='cl is the headline (BASELINE configs[1])')
    ap.add_argument('--no-skip-padding', action='store_true', help='run pad tokens through the encoder like the reference does')
    ap.add_argument('--no-prefetch', action='store_true', help='compute the id plumbing of each batch inside its own step')
    ap.add_argument('--no-dedup', action='store_true', help='encode every (impression, slot) title, not each distinct article once')
    ap.add_argument('--eval-impressions', type=int, default=376_471)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == 'reference':
        return run_reference(args)
    if args.workload == 'eval':
        return run_eval(args)

    import torch.distributed as dist
    from xnrs_b200 import kernels as K
    from xnrs_b200 import synthetic as syn
    from xnrs_b200.data import TitleStore
    from xnrs_b200.distributed import DataParallelTrainer
    from xnrs_b200.models import make_model
    from xnrs_b200.training import ContrastiveRankingTrainer

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    if world != args.gpus and rank == 0:
        print(f'warning: --gpus {args.gpus} but WORLD_SIZE {world}', file=sys.stderr)
    K.set_precision(args.precision)
    B = args.batch
    from xnrs_b200.models.components import TextEncoder
    TextEncoder.dedup_titles = not args.no_dedup
    TextEncoder.skip_padding = not args.no_skip_padding

    cfg = MODEL_CFGS[args.model]
    cat = syn.make_catalogue(N_NEWS, SEQ_LEN, VOCAB, 768, seed=0, with_abstract=(args.model == 'naml'))
    store = TitleStore(cat.token_table.to(dev), cat.title_tokens.to(dev))
    astore = TitleStore(store.token_table, cat.abstract_tokens.to(dev)) if args.model == 'naml' else None
    torch.manual_seed(0)
    from xnrs_b200.training import MSERankingTrainer
    trainer_cls = MSERankingTrainer if args.model == 'npa' else ContrastiveRankingTrainer     # NPA has no CL hook
    trainer = trainer_cls(dict(cfg, device=str(dev)), make_model(cfg))
    trainer.model.train()
    dp = DataParallelTrainer(trainer)

    n_batches = 8           # distinct batches cycled through; weak scaling: every rank draws its own B impressions
    raws = [syn.make_train_batch(N_NEWS, B, HIST_LEN, seed=1000 + 97 * rank + i) for i in range(n_batches)]
    pinned = [{k: v.pin_memory() for k, v in r.items()} for r in raws]
    resident = [syn.index_batch(store, cat, r, dev, abstract_store=astore) for r in raws]
    h2d_bytes = sum(v.numel() * v.element_size() for v in raws[0].values())

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for i in range(args.warmup):
        dp.train_step(resident[i % n_batches])

    # ---- timed region 1: inputs resident in HBM.  Run twice over the same K steps: (a) every kernel call bracketed by
    # CUDA events on its launching stream (per-kernel durations for the roofline), (b) without the event hooks (`value`)
    records = []

    def hook(name, args_):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        records.append((name, args_, s, e))
        return s, e

    def timed_resident():
        sync_all()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        dp.prefetch(resident[0])
        t0.record()
        for i in range(args.steps):
            dp.train_step(resident[i % n_batches])
            if not args.no_prefetch:         # input pipeline: the next batch's id plumbing runs on a side stream meanwhile
                dp.prefetch(resident[(i + 1) % n_batches])
        t1.record()
        sync_all()
        ms = torch.tensor([t0.elapsed_time(t1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)

    K.set_event_hook(hook)
    ms_hooked = timed_resident()
    K.set_event_hook(None)
    n0 = K.launch_count()
    with ClockSampler(local) as clocks:
        ms_total = timed_resident()
    launches = K.launch_count() - n0
    value = B * world * args.steps / (ms_total * 1e-3)

    # dominant kernel: aggregate event time per entry point; GEMM FLOPs are 2*M*N*K per launch
    agg, shapes = {}, {}
    for name, a, s, e in records:
        d = agg.setdefault(name, [0.0, 0, 0.0])
        dt = s.elapsed_time(e)
        d[0] += dt
        d[1] += 1
        if name == 'xnrs_gemm':
            d[2] += 2.0 * a[2] * a[3] * a[4]
            sh = shapes.setdefault(f'{"T" if a[0] else "N"}{"T" if a[1] else "N"} M={a[2]} N={a[3]} K={a[4]}', [0.0, 0, 0.0])
            sh[0] += dt
            sh[1] += 1
            sh[2] += 2.0 * a[2] * a[3] * a[4]
    top = max(agg.items(), key=lambda kv: kv[1][0])
    if os.environ.get('XNRS_BENCH_DUMP') and rank == 0:      # every launch of ONE step with its event time (diagnostics)
        per = len(records) // args.steps
        with open(os.environ['XNRS_BENCH_DUMP'], 'w') as f:
            for name, a, s_, e_ in records[:per]:
                f.write(json.dumps({'name': name, 'args': list(a), 'ms': round(s_.elapsed_time(e_), 4)}) + '\n')
    pk, pk_kind = peaks()
    gemm_ms, gemm_n, gemm_flop = agg.get('xnrs_gemm', [0.0, 0, 0.0])
    kernel_ms_total = sum(v[0] for v in agg.values())
    # dominant kernel = the GEMM launch class (layout, N, K; M varies with the batch's distinct-token count) with the most time
    classes = {}
    for name, a, s_, e_ in records:
        if name == 'xnrs_gemm':
            big = a[4] if a[0] else a[2]          # the batch-dependent dimension; its power-of-two bucket keeps e.g. the
            bucket = max(int(big), 1).bit_length()   # token-level fc1 GEMMs apart from the title-level head GEMMs of equal N, K
            key = ((a[0], a[1], a[3], a[4], a[8]) if not a[0] else (a[0], a[1], a[2], a[3], a[8])) + (bucket,)
            c = classes.setdefault(key, [0.0, 0, 0.0, 0])
            c[0] += s_.elapsed_time(e_)
            c[1] += 1
            c[2] += 2.0 * a[2] * a[3] * a[4]
            c[3] += a[4] if a[0] else a[2]
    (d_ta, d_tb, d_1, d_2, d_act, _), (d_ms, d_n, d_flop, d_rows) = max(classes.items(), key=lambda kv: kv[1][0])
    d_m, d_n_ = (d_1, d_2) if d_ta else (d_rows // max(d_n, 1), d_1)
    kern = ('gemm_tc2_kernel (cta_group::2 CTA pair, 256x256 tile)' if (args.precision == 'tf32x3' and d_n_ > 128 and d_m >= 256)
            else ('gemm_tc_kernel<256>' if args.precision in ('tf32', 'bf16') and d_n_ > 128 else 'gemm_tc_kernel<128>'))
    d_name = (f'{kern} {"TN" if d_ta else "NT"} '
              + (f'M={d_1} N={d_2} K~{d_rows // max(d_n, 1)} (fc1 weight gradient)' if d_ta
                 else f'M~{d_rows // max(d_n, 1)} N={d_1} K={d_2} ' + ('(title fc1 forward, tanh epilogue)' if d_act == 2 else '(forward)')))
    d_tflops = d_flop / (d_ms * 1e-3) / 1e12 if d_ms else None
    roofline = {
        'kernel': d_name if args.precision != 'fp32' else 'gemm_simt_kernel ' + d_name,
        'bound': 'tensor', 'achieved': d_tflops, 'peak': pk['bf16_tflops_sustained'], 'unit': 'TFLOP/s',
        'frac': d_tflops / pk['bf16_tflops_sustained'] if d_tflops else None,
        # dram__bytes_read.sum + dram__bytes_write.sum of ONE such launch (ncu --set full, profiles/README.md): 473.2 MB read
        # + 131.4 MB written at M=153 562 rows = 3937 B/row, i.e. the algorithmic A row (3072 B) + C row (1024 B): no re-reads
        # the fc1 weight gradient (TN, M=256, N=768): 651.4 MB read + 8.4 MB written at K=153 562 = 4296 B per k-row = one
        # d_hid row (1024 B) + one x row (3072 B) + the split-K partial tiles
        'traffic': (3937.0 * (d_rows / max(d_n, 1)) if (not d_ta and args.precision != 'fp32' and d_2 == 768 and d_1 == 256)
                    else (4296.0 * (d_rows / max(d_n, 1)) if (d_ta and args.precision != 'fp32' and d_1 == 256 and d_2 == 768) else None)),
        'peak_source': f'{pk_kind} (sustained bf16 GEMM; kernel timed inside a long step). The kernel computes in TF32 '
                       f'(3 MMAs per product in the fp32-accurate 3xTF32 mode): its own ceiling is 1/6 of this bf16 peak',
        'launches_timed': d_n, 'avg_launch_ms': d_ms / max(d_n, 1),
        'all_gemm_tflops': gemm_flop / (gemm_ms * 1e-3) / 1e12 if gemm_ms else None,
        'gemm_share_of_step_kernel_time': gemm_ms / kernel_ms_total if kernel_ms_total else None,
        'top_entry_point_by_time': top[0], 'ms_per_step_with_event_hooks': ms_hooked / args.steps,
        'gemm_shapes_ms_per_step': {k: f'{v[0] / args.steps:.3f} ms, {v[2] / (v[0] * 1e-3) / 1e12:.1f} TF/s, {v[1] // args.steps}x'
                                    for k, v in sorted(shapes.items(), key=lambda kv: -kv[1][0])[:8]},
        'per_entry_point_ms_per_step': {k: round(v[0] / args.steps, 4) for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])},
    }

    # ---- timed region 2: end to end through the public trainer API with HOST (pinned) index buffers ----
    def h2d(i):                              # host (pinned) int32 ids / targets / labels -> device, on the main stream
        b = syn.index_batch(store, cat, pinned[i % n_batches], dev, abstract_store=astore)
        return b, torch.cuda.current_stream().record_event()

    def e2e_step(cur, i):
        nxt = h2d(i + 1)                     # tiny copy, queued BEFORE this step so the side stream can plan it meanwhile
        out = dp.train_step(cur[0])
        if not args.no_prefetch:
            dp.prefetch(nxt[0], after=nxt[1])
        return float(out['loss']), nxt       # device -> host read of the step's loss (4 bytes, synchronises)

    _, cur = e2e_step(h2d(0), 0)
    sync_all()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    wall0 = time.perf_counter()
    e0.record()
    last = 0.0
    for i in range(args.steps):
        last, cur = e2e_step(cur, i + 1)
    e1.record()
    sync_all()
    wall = time.perf_counter() - wall0
    ems = torch.tensor([max(e0.elapsed_time(e1), wall * 1e3)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ems, op=dist.ReduceOp.MAX)
    e2e_value = B * world * args.steps / (float(ems) * 1e-3)

    out = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': ms_total / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': {'fp32': 'f32', 'tf32x3': 'f32 (3xTF32)', 'tf32': 'tf32', 'bf16': 'bf16'}[args.precision],
        'data': 'synthetic',
        'config': {'workload': workload_name(B, args.model), 'global_batch': B * world, 'parallelism': f'dp{world}',
                   'l2': 'inputs larger than L2: 307 MB token table, ~5 GB of gathered rows per step, 8 batches cycled',
                   'precision': args.precision, 'final_loss': last,
                   'dedup_titles': not args.no_dedup, 'skip_padding': not args.no_skip_padding},
        'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': h2d_bytes, 'd2h_bytes_per_step': 4,
                'note': 'host input = int32 news ids / targets / labels (the index fast path of the drop-in API)'},
        'gpu_launches': launches, 'roofline': roofline, 'clocks': clocks.summary(),
    }
    if world == 1 and not args.no_cpu_baseline and args.model == 'cl':
        threads = os.cpu_count() or 1
        v, per = time_cpu(args.ref_batch, 3, 1, threads)
        out['cpu_baseline'] = {'value': v, 'unit': UNIT, 'cores': threads, 'kind': 'port',
                               'sample': f'{args.ref_batch} impressions/step x 3 steps after 1 warm-up, oracle '
                                         f'(torch CPU fp32, autograd + Adam), same shapes'}
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
