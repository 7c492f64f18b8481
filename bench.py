#!/usr/bin/env python
"""Benchmark of the xnrs bi-encoder hot path on B200 (BASELINE.json: "train impressions/s & eval scored impressions/s").

    python bench.py [--gpus N] [--steps K] [--warmup W]          ONE JSON line:
        headline   = CL bi-encoder (config/mind_small_CL.yml, `model: standard`) TRAINING throughput, BASELINE configs[1]
        sub.nrms_train = NRMS (config/mind_small_NRMS.yml) training throughput, BASELINE configs[0]
        sub.eval       = MIND-large-shaped full-catalogue evaluation (160k news, 376 471 impressions), BASELINE configs[4]
        sub.cl_train_bf16 / sub.eval_bf16 = the same two in the bf16 storage mode (2e-2 tolerance class), as second lines
      each with its own value / e2e / roofline / cpu_baseline; north-star shapes S=30, H=50, 1:4 negatives, synthetic data.
    python bench.py --only cl|nrms|naml|lstur|npa|eval ...        one workload (diagnostics, secondary models)
    python bench.py --impl reference ...                          the UNMODIFIED reference (oracle/_ref, see oracle/make_ref.py)
                                                                  on the box's host cores, same line format

A "step" of a train workload = one full trainer step over one batch of impressions (index gather from the device-resident
token table, title + user encoders, fused score + loss, InfoNCE, backward, Adam); a "step" of the eval workload = one full
pass (catalogue encode + every impression).  `value`: inputs resident in HBM; `e2e`: through the public trainer /
evaluator API from pinned HOST buffers, H2D + result read-back inside the timed region.  Every timed region is EXACTLY K
steps bracketed by barrier + synchronize; regions are repeated until >= 1 s has been measured and the MEDIAN region is
reported (min / max beside it).  N > 1 is launched by torchrun (one rank per GPU, NCCL); rank 0 prints the line.
"""
import argparse
import json
import os
import statistics
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
# the other BASELINE.json configs, hyper-parameters from config/mind_small_*.yml
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
EVAL_NEWS = 160_000
UNIT = 'impressions/s'
MIN_TIMED_S = float(os.environ.get('XNRS_BENCH_MIN_S', '1.3'))   # every reported number comes from >= this much measured time (profiler runs: 0)
TRAFFIC_FILE = os.path.join(ROOT, 'profiles', 'ncu_traffic.json')


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
            self.rows.clear()                    # samples from before the timed region do not count
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


class Ctx:
    """process-wide state: rank / world / device and the barrier-bracketed region timer"""

    def __init__(self, gpus):
        import torch.distributed as dist
        self.dist = dist
        self.world = int(os.environ.get('WORLD_SIZE', '1'))
        self.rank = int(os.environ.get('RANK', '0'))
        self.local = int(os.environ.get('LOCAL_RANK', '0'))
        torch.cuda.set_device(self.local)
        self.dev = torch.device('cuda', self.local)
        if self.world > 1:
            dist.init_process_group('nccl', device_id=self.dev)
        if self.world != gpus and self.rank == 0:
            print(f'warning: --gpus {gpus} but WORLD_SIZE {self.world}', file=sys.stderr)

    def sync_all(self):
        torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(self, x: float) -> float:
        if self.world == 1:
            return float(x)
        t = torch.tensor([x], device=self.dev, dtype=torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t)

    def timed_regions(self, region, min_total_s=MIN_TIMED_S, max_regions=400):
        """region() runs EXACTLY K steps and returns its duration in ms (this rank); every region is bracketed by barrier +
        synchronize on both sides and its duration is the MAX over ranks.  Regions repeat until min_total_s is covered
        (the repetition count follows the all-reduced times, so every rank takes the same decision)."""
        out, total = [], 0.0
        while (total < min_total_s * 1e3 and len(out) < max_regions) or not out:
            self.sync_all()
            ms = region()
            self.sync_all()
            ms = self.max_over_ranks(ms)
            out.append(ms)
            total += ms
        return out

    def close(self):
        if self.world > 1:
            from xnrs_b200 import distributed as D
            for px in D._peer_cache.values():       # a peer kernel that gave up waiting would have produced garbage: fail loudly
                px.check()
            D._peer_cache.clear()                   # peer mappings go before the communicator
            self.dist.destroy_process_group()


def spread(ms_list, units_per_region):
    """median region -> value; the run-to-run resolution beside it"""
    med = statistics.median(ms_list)
    return med, {'regions': len(ms_list), 'timed_s': round(sum(ms_list) * 1e-3, 3),
                 'value_min': units_per_region / (max(ms_list) * 1e-3), 'value_max': units_per_region / (min(ms_list) * 1e-3)}


def traffic_entry(key):
    """measured DRAM traffic of a kernel class from the tracked ncu summary (profiles/ncu_traffic.json), or None"""
    try:
        return json.load(open(TRAFFIC_FILE))['classes'].get(key)
    except Exception:
        return None


# =====================================================================================================================
# CPU arms: the unmodified reference (oracle/_ref) when it travelled with the repo, else the oracle port
# =====================================================================================================================

def _str_themes(batch):
    b = dict(batch)
    b['main_theme'] = [str(int(t)) for t in batch['main_theme']]          # the reference numbers theme STRINGS (training.py:414-417)
    return b


def cpu_train_step_fn(model_key, batch_size, threads, device='cpu', tf32=False):
    """-> (step(), kind): one ContrastiveRankingTrainer._train_step (MSERankingTrainer for NPA, which has no CL hook) of the
    reference on dense reference-format batches built from the same synthetic ids.  kind "reference": the unmodified
    reference package; "port": the oracle's restatement (two history forwards like training.py:402-431, autograd, Adam)."""
    from oracle import refload
    from xnrs_b200 import synthetic as syn
    torch.set_num_threads(threads)
    cfg = dict(MODEL_CFGS[model_key], device=str(device), seq_len=SEQ_LEN, hist_len=HIST_LEN, batch_size=batch_size)
    cat = syn.make_catalogue(N_NEWS, SEQ_LEN, VOCAB, 768, seed=0, with_abstract=(model_key == 'naml'))

    def to_dev(x):
        if isinstance(x, torch.Tensor):
            return x.to(device)
        if isinstance(x, dict):
            return {k: to_dev(v) for k, v in x.items()}
        if isinstance(x, (tuple, list)):
            return type(x)(to_dev(v) for v in x)
        return x

    batches = [to_dev(syn.dense_batch(cat, syn.make_train_batch(N_NEWS, batch_size, HIST_LEN, seed=100 + i),
                                      with_abstract=(model_key == 'naml'))) for i in range(2)]
    counter = [0]
    if refload.available():
        torch.manual_seed(0)
        tr = refload.make_trainer(cfg, 'MSERankingTrainer' if model_key == 'npa' else 'ContrastiveRankingTrainer')
        tr.model.train()
        sbatches = [_str_themes(b) for b in batches]

        def step():
            b = sbatches[counter[0] % 2]
            counter[0] += 1
            out = tr._train_step(b)
            return float(out['loss'])
        return step, 'reference'

    from oracle import xnrs_oracle as O
    from xnrs_b200.models import make_model
    if model_key != 'cl':
        raise RuntimeError('oracle port of the CPU step exists for the CL model only (oracle/_ref missing)')
    torch.manual_seed(0)
    P = {k: v.detach().clone().to(device).requires_grad_(True) for k, v in make_model(CL_CFG).state_dict().items()}
    state = {k: (torch.zeros_like(v), torch.zeros_like(v)) for k, v in P.items()}

    def step():
        b = batches[counter[0] % 2]
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
    return step, 'port'


def time_cpu_train(model_key, batch_size, steps, warmup, threads):
    step, kind = cpu_train_step_fn(model_key, batch_size, threads)
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    return batch_size * steps / dt, dt / steps, kind


def cpu_baseline_train(model_key, batch_size, steps=3, warmup=1):
    threads = os.cpu_count() or 1
    try:
        v, per, kind = time_cpu_train(model_key, batch_size, steps, warmup, threads)
        return {'value': v, 'unit': UNIT, 'cores': threads, 'kind': kind,
                'sample': f'{batch_size} impressions/step x {steps} steps after {warmup} warm-up, '
                          + ('the unmodified reference trainer step (oracle/_ref, torch CPU fp32)' if kind == 'reference'
                             else 'oracle port (torch CPU fp32, autograd + Adam)') + f', same shapes, {per:.2f} s/step'}
    except Exception as exc:                                   # a baseline must never take the measurement down with it
        return {'value': None, 'unit': UNIT, 'cores': threads, 'kind': 'port', 'sample': f'failed: {exc!r}'}


def eager_gpu_baseline(dev, batch_size=256, steps=5):
    """SURVEY §8(d) last row: the reference's own training step in STOCK PyTorch eager on the same B200 (reference modules
    .to(cuda), dense reference-format batches resident in HBM, cuBLAS fp32 with TF32 off, and TF32 on for context)."""
    out = {'batch': batch_size, 'unit': UNIT}
    try:
        old = torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32
        step, kind = cpu_train_step_fn('cl', batch_size, os.cpu_count() or 1, device=dev)
        out['kind'] = ('unmodified reference modules + trainer step on cuda (oracle/_ref)' if kind == 'reference'
                       else 'torch restatement of the reference step on cuda (oracle port)')
        for tf32 in (False, True):
            torch.backends.cuda.matmul.allow_tf32 = tf32
            torch.backends.cudnn.allow_tf32 = tf32
            for _ in range(2):
                step()
            torch.cuda.synchronize()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            for _ in range(steps):
                step()
            e.record()
            torch.cuda.synchronize()
            ms = s.elapsed_time(e) / steps
            out['tf32' if tf32 else 'fp32'] = {'value': batch_size / (ms * 1e-3), 'ms_per_step': ms}
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old
    except Exception as exc:
        out['failed'] = repr(exc)
    torch.cuda.empty_cache()
    return out


def cpu_eval_fn(cat, imp, model_sd, threads):
    """-> (run(i) for impression i, kind): the reference evaluation procedure (training.py:131-142,194-243: ONE impression
    per step — full re-encode of the history and every candidate, scores to the host, numpy / sklearn metrics)."""
    from oracle import refload
    from xnrs_b200 import synthetic as syn
    torch.set_num_threads(threads)

    def raw_of(i):
        a, b = int(imp['offsets'][i]), int(imp['offsets'][i + 1])
        return {'hist_ids': imp['hist_ids'][i:i + 1], 'cand_ids': imp['cand_ids'][a:b][None, :],
                'targets': imp['targets'][a:b][None, :, None], 'user_index': imp['user_index'][i:i + 1],
                'main_theme': torch.zeros(1, dtype=torch.int32)}

    if refload.available():
        cfg = dict(CL_CFG, device='cpu', seq_len=SEQ_LEN, hist_len=HIST_LEN, batch_size=1)
        tr = refload.make_trainer(cfg, 'MSERankingTrainer')
        tr.model.load_state_dict({k: v.detach().cpu() for k, v in model_sd.items()})
        tr.model.eval()

        def run(i):
            return tr._test_step(syn.dense_batch(cat, raw_of(i)))
        return run, 'reference'

    from oracle import xnrs_oracle as O
    P = {k: v.detach().cpu().clone() for k, v in model_sd.items()}

    def run(i):
        a, b = int(imp['offsets'][i]), int(imp['offsets'][i + 1])
        with torch.no_grad():
            scores = torch.relu(O.parent_forward(P, syn.dense_batch(cat, raw_of(i)))).reshape(-1)
        return O.impression_metrics(imp['targets'][a:b].numpy(), scores.numpy())
    return run, 'port'


def cpu_baseline_eval(cat, imp, model_sd, n_sample):
    threads = os.cpu_count() or 1
    try:
        run, kind = cpu_eval_fn(cat, imp, model_sd, threads)
        run(0)
        t0 = time.perf_counter()
        for i in range(n_sample):
            run(i)
        dt = time.perf_counter() - t0
        n_imp = imp['offsets'].numel() - 1
        return {'value': n_sample / dt, 'unit': UNIT, 'cores': threads, 'kind': kind,
                'sample': f'{n_sample} of the {n_imp} impressions, one per step with a full re-encode of history and candidates + '
                          + ('numpy/sklearn metrics: the unmodified reference RankingTrainer._test_step (oracle/_ref)' if kind == 'reference'
                             else 'numpy metrics like training.py:194-243 (oracle port)') + f', torch CPU fp32, {dt:.1f} s'}
    except Exception as exc:
        return {'value': None, 'unit': UNIT, 'cores': threads, 'kind': 'port', 'sample': f'failed: {exc!r}'}


# =====================================================================================================================
# train workloads
# =====================================================================================================================

class HookLog:
    """brackets every C-ABI entry point with CUDA events on its launching stream (kernels.set_event_hook)"""

    class Rec:
        __slots__ = ('name', 'args', 'start', 'end', 'kernel')

        def __init__(self, name, args):
            self.name, self.args, self.kernel = name, args, None
            self.start, self.end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def __init__(self):
        self.records = []

    def __call__(self, name, args):
        r = HookLog.Rec(name, args)
        self.records.append(r)
        return r


def gemm_rooflines(records, steps, precision):
    """dominant-kernel roofline from one hooked pass: GEMM launches are grouped by (layout, the two batch-independent
    dimensions, activation, power-of-two bucket of the batch-dependent dimension, kernel the LIBRARY dispatched to)."""
    pk, pk_kind = peaks()
    agg, shapes, classes = {}, {}, {}
    for r in records:
        dt = r.start.elapsed_time(r.end)
        d = agg.setdefault(r.name, [0.0, 0, 0.0])
        d[0] += dt
        d[1] += 1
        if r.name == 'xnrs_gemm':
            a = r.args
            flop = 2.0 * a[2] * a[3] * a[4]
            d[2] += flop
            sh = shapes.setdefault(f'{"T" if a[0] else "N"}{"T" if a[1] else "N"} M={a[2]} N={a[3]} K={a[4]}', [0.0, 0, 0.0])
            sh[0] += dt
            sh[1] += 1
            sh[2] += flop
            big = a[4] if a[0] else a[2]                       # the batch-dependent dimension
            key = ((a[0], a[1], a[3], a[4], a[8]) if not a[0] else (a[0], a[1], a[2], a[3], a[8])) + (max(int(big), 1).bit_length(), r.kernel)
            c = classes.setdefault(key, [0.0, 0, 0.0, 0])
            c[0] += dt
            c[1] += 1
            c[2] += flop
            c[3] += big
    if not classes:
        return None, agg
    (d_ta, d_tb, d_1, d_2, d_act, _, d_kernel), (d_ms, d_n, d_flop, d_rows) = max(classes.items(), key=lambda kv: kv[1][0])
    rows = d_rows / max(d_n, 1)
    if d_ta:
        role, cls = f'TN M={d_1} N={d_2} K~{rows:.0f} (weight gradient: K = token / title rows)', f'TN_{d_1}x{d_2}'
    else:
        role, cls = f'NT M~{rows:.0f} N={d_1} K={d_2} (forward' + (', tanh epilogue)' if d_act == 2 else ')'), f'NT_{d_1}x{d_2}_act{d_act}'
    tr = traffic_entry(cls)
    gemm_ms, gemm_n, gemm_flop = agg.get('xnrs_gemm', [0.0, 0, 0.0])
    kernel_ms_total = sum(v[0] for v in agg.values())
    top = max(agg.items(), key=lambda kv: kv[1][0])
    d_tflops = d_flop / (d_ms * 1e-3) / 1e12 if d_ms else None
    roofline = {
        'kernel': f'{d_kernel} {role}', 'bound': 'tensor', 'achieved': d_tflops, 'peak': pk['bf16_tflops_sustained'],
        'unit': 'TFLOP/s', 'frac': d_tflops / pk['bf16_tflops_sustained'] if d_tflops else None,
        # per-launch DRAM bytes of this kernel class from the tracked `ncu --set full` summary, scaled by the batch-dependent
        # row count of THIS run (bytes per row x rows); null when the class has no capture
        'traffic': (tr['dram_bytes_per_row'] * rows) if tr else None,
        'traffic_source': (f"{os.path.relpath(TRAFFIC_FILE, ROOT)}:{cls} ({tr['dram_bytes_per_row']} B/row from {tr['source']}) x {rows:.0f} rows"
                           if tr else 'no ncu capture of this kernel class'),
        'algorithmic_bytes': (tr['algorithmic_bytes_per_row'] * rows) if tr and 'algorithmic_bytes_per_row' in tr else None,
        'peak_source': f'{pk_kind} (MEASURED_PEAKS.json, sustained bf16 GEMM; kernel timed inside a long step). '
                       + {'tf32x3': 'The kernel issues 3 TF32 MMAs per product (fp32-accurate 3xTF32): its own ceiling is 1/6 of this bf16 peak',
                          'tf32': 'TF32 operands: ceiling 1/2 of this bf16 peak', 'bf16': 'bf16 operands',
                          'bf16x3': 'fp32-accurate 3-pass 16-bit split (fp16 planes forward, bf16 planes backward) on the token-level launches (ceiling 1/3 of this peak), 3xTF32 elsewhere',
                          'fp32': 'exact-fp32 SIMT kernel (no tensor cores)'}[precision],
        'launches_timed': d_n, 'avg_launch_ms': d_ms / max(d_n, 1),
        'all_gemm_tflops': gemm_flop / (gemm_ms * 1e-3) / 1e12 if gemm_ms else None,
        'gemm_share_of_step_kernel_time': gemm_ms / kernel_ms_total if kernel_ms_total else None,
        'top_entry_point_by_time': top[0],
        'gemm_shapes_ms_per_step': {k: f'{v[0] / steps:.3f} ms, {v[2] / (v[0] * 1e-3) / 1e12:.1f} TF/s, {v[1] // steps}x'
                                    for k, v in sorted(shapes.items(), key=lambda kv: -kv[1][0])[:8]},
        'per_entry_point_ms_per_step': {k: round(v[0] / steps, 4) for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])},
        'kernel_ms_per_step': kernel_ms_total / steps,
    }
    return roofline, agg


def train_workload(ctx, args, model_key, want_cpu=True, want_eager=False):
    from xnrs_b200 import kernels as K
    from xnrs_b200 import _lib
    from xnrs_b200 import synthetic as syn
    from xnrs_b200.data import TitleStore
    from xnrs_b200.distributed import DataParallelTrainer
    from xnrs_b200.models import make_model
    from xnrs_b200.models.components import encoder_options
    from xnrs_b200.training import ContrastiveRankingTrainer, MSERankingTrainer

    dev, world, rank = ctx.dev, ctx.world, ctx.rank
    K.set_precision(args.precision)
    B, steps = args.batch, args.steps
    cfg = MODEL_CFGS[model_key]
    cat = syn.make_catalogue(N_NEWS, SEQ_LEN, VOCAB, 768, seed=0, with_abstract=(model_key == 'naml'))
    store = TitleStore(cat.token_table.to(dev), cat.title_tokens.to(dev))
    astore = TitleStore(store.token_table, cat.abstract_tokens.to(dev)) if model_key == 'naml' else None
    torch.manual_seed(0)
    trainer_cls = MSERankingTrainer if model_key == 'npa' else ContrastiveRankingTrainer     # NPA has no CL hook
    model = make_model(cfg)
    encoder_options(model, dedup_titles=not args.no_dedup, skip_padding=not args.no_skip_padding)
    use_graph = model_key == 'cl' and not args.no_graph and (world == 1 or not args.no_graph_multi_gpu)
    trainer = trainer_cls(dict(cfg, device=str(dev)), model, graph_safe=use_graph)
    trainer.model.train()
    dp = DataParallelTrainer(trainer)
    if use_graph:           # whole step (zero_grad .. Adam, gradient collectives included) replayed as ONE CUDA graph per shape bucket
        from xnrs_b200.graphs import GraphedStep
        stepper = GraphedStep(dp)
        run_step = stepper.step
    else:
        stepper, run_step = None, dp.train_step

    n_batches = 8           # distinct batches cycled through; weak scaling: every rank draws its own B impressions
    raws = [syn.make_train_batch(N_NEWS, B, HIST_LEN, seed=1000 + 97 * rank + i) for i in range(n_batches)]
    pinned = [{k: v.pin_memory() for k, v in r.items()} for r in raws]
    resident = [syn.index_batch(store, cat, r, dev, abstract_store=astore, categories=(model_key in ('naml', 'lstur'))) for r in raws]
    h2d_bytes = sum(v.numel() * v.element_size() for v in raws[0].values())

    for i in range(max(args.warmup, 2 * n_batches if use_graph else 0)):      # graphs: every cycled batch's bucket gets captured
        run_step(resident[i % n_batches])

    # ---- timed region 1: inputs resident in HBM --------------------------------------------------------------------
    def region_resident(step_fn=None):
        step_fn = step_fn or run_step
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        depth = max(1, min(args.prefetch_depth, n_batches - 1))
        for d in range(depth):
            dp.prefetch(resident[d % n_batches])
        t0.record()
        for i in range(steps):
            step_fn(resident[i % n_batches])
            if not args.no_prefetch:         # input pipeline: the id plumbing of the batch `depth` steps ahead runs on a side stream
                dp.prefetch(resident[(i + depth) % n_batches])
        t1.record()
        torch.cuda.synchronize()
        return t0.elapsed_time(t1)

    # (a) one region with every C-ABI call bracketed by CUDA events on its launching stream: per-kernel durations
    log = HookLog()
    K.set_event_hook(log)
    # (kernel by kernel even when `value` replays CUDA graphs: a replay launches nothing through the C ABI)
    ms_hooked = ctx.timed_regions(lambda: region_resident(dp.train_step), min_total_s=0.0)[0]
    K.set_event_hook(None)
    fb0 = int(_lib.lib().xnrs_gemm_simt_fallbacks())
    # (b) without hooks: `value`
    n0 = K.launch_count() + (stepper.replayed_kernels if stepper else 0)
    g0 = stepper.replays if stepper else 0
    with ClockSampler(ctx.local) as clocks:
        ms_list = ctx.timed_regions(region_resident)
    # kernels of this library EXECUTED per region: launched one by one + those inside replayed CUDA graphs
    launches = (K.launch_count() + (stepper.replayed_kernels if stepper else 0) - n0) // len(ms_list)
    graph_launches = ((stepper.replays - g0) // len(ms_list)) if stepper else 0
    fallbacks = (int(_lib.lib().xnrs_gemm_simt_fallbacks()) - fb0) // len(ms_list)
    ms_med, sp = spread(ms_list, B * world * steps)
    value = B * world * steps / (ms_med * 1e-3)
    roofline, agg = gemm_rooflines(log.records, steps, args.precision)
    if roofline is not None:
        roofline['ms_per_step_with_event_hooks'] = ms_hooked / steps
        roofline['simt_fallback_gemms_per_region'] = fallbacks
        # the fused title-pooling forward (gather -> fc1 -> tanh -> logit -> exp -> per-title sums: ONE tcgen05 launch + a small
        # normalisation pass) is its own entry point: FLOPs = 2 * rows * A * F of the fc1 product inside it
        gb = [r for r in log.records if r.name in ('xnrs_gemm_bf16', 'xnrs_gemm_bf16x3')]
        if gb:      # bf16 storage / 3xBF16 modes: the token-level weight gradient runs in xnrs_gemm_bf16(x3) (args: transA, transB, M, N, K, ...)
            ms_gb = sum(r.start.elapsed_time(r.end) for r in gb)
            flop_gb = sum(2.0 * r.args[2] * r.args[3] * r.args[4] for r in gb)
            pk_, _ = peaks()
            roofline['bf16_weight_gradient'] = {
                'kernel': 'gemm_tc2_kernel (cta_group::2 pair tile, ' + ('3-pass 16-bit split on pre-split bf16 planes' if gb[0].name.endswith('x3') else 'BF16')
                          + ' kind::f16, cp.async B gather): dW1 = d_hid^T x',
                'launches_timed': len(gb), 'avg_launch_ms': ms_gb / len(gb), 'achieved': flop_gb / (ms_gb * 1e-3) / 1e12,
                'unit': 'TFLOP/s', 'peak': pk_['bf16_tflops_sustained'], 'frac': flop_gb / (ms_gb * 1e-3) / 1e12 / pk_['bf16_tflops_sustained']}
            trg = traffic_entry('dw_x3') if gb[0].name.endswith('x3') else None
            if trg:
                kk = sum(r.args[4] for r in gb) / len(gb)
                roofline['bf16_weight_gradient'].update({'traffic': trg['dram_bytes_per_row'] * kk, 'algorithmic_bytes': trg['algorithmic_bytes_per_row'] * kk,
                                                         'traffic_source': f"{os.path.relpath(TRAFFIC_FILE, ROOT)}:dw_x3"})
        tp = [r for r in log.records if r.name.startswith('xnrs_titlepool_fwd')]
        if tp:
            bf, x3 = tp[0].name.endswith('bf16'), tp[0].name.endswith('bf16x3')
            ms_tp = sum(r.start.elapsed_time(r.end) for r in tp)
            rows_tp = sum(r.args[1] for r in tp)
            flop_tp = sum(2.0 * r.args[1] * r.args[3] * r.args[4] for r in tp)
            pk_, _ = peaks()
            cls = 'titlepool_bf16' if bf else ('titlepool_x3' if x3 else 'titlepool_fp32')
            tr = traffic_entry(cls)
            roofline['fused_title_pool'] = {
                'kernel': 'gemm_tc2_kernel<POOL> (cta_group::2 pair tile, cp.async table gather, pooling epilogue on TMEM, hid through '
                          'TMA stores; ' + ('bf16 storage, kind::f16)' if bf else ('3-pass 16-bit split on pre-split fp16 planes, kind::f16)' if x3 else 'fp32 storage, 3xTF32)'))
                          + ' + titlepool_wsum_kernel (per-title weighted sums); the whole entry point is timed',
                'launches_timed': len(tp), 'avg_launch_ms': ms_tp / len(tp), 'rows_per_launch': rows_tp / len(tp),
                'achieved': flop_tp / (ms_tp * 1e-3) / 1e12, 'unit': 'TFLOP/s', 'peak': pk_['bf16_tflops_sustained'],
                'frac': flop_tp / (ms_tp * 1e-3) / 1e12 / pk_['bf16_tflops_sustained'],
                'traffic': tr['dram_bytes_per_row'] * rows_tp / len(tp) if tr else None,
                'algorithmic_bytes': tr['algorithmic_bytes_per_row'] * rows_tp / len(tp) if tr else None,
                'traffic_source': f"{os.path.relpath(TRAFFIC_FILE, ROOT)}:{cls}" if tr else 'no ncu capture of this kernel class'}
        if args.precision in ('bf16', 'bf16x3') and (tp or gb):
            # bf16 storage mode: the step's dominant kernels are the two bf16 tensor-core launches above (the xnrs_gemm classes
            # left on fp32 operands are the small title-level GEMMs): the headline roofline fields follow the longer of them
            dom = max((k for k in ('fused_title_pool', 'bf16_weight_gradient') if k in roofline),
                      key=lambda k: roofline[k]['avg_launch_ms'] * roofline[k]['launches_timed'])
            roofline['small_fp32_gemm_class'] = {k: roofline[k] for k in ('kernel', 'achieved', 'frac', 'launches_timed', 'avg_launch_ms')}
            for k in ('kernel', 'achieved', 'frac', 'launches_timed', 'avg_launch_ms'):
                roofline[k] = roofline[dom][k]
            roofline['traffic'], roofline['traffic_source'] = roofline[dom].get('traffic'), roofline[dom].get('traffic_source', 'no ncu capture of this kernel class')
        elif tp:
            roofline['note'] = ('the dominant GEMM class is the fc1 weight gradient WITH the table gather fused in (cp.async warps): no dense '
                                'copy of the gathered rows exists any more; the same GEMM on a dense copy runs 0.39 ms (153 TFLOP/s), the '
                                'copy itself 0.16 ms (tools/bench_gather_gemm.py)')
    if os.environ.get('XNRS_BENCH_DUMP') and rank == 0:      # every launch of ONE step with its event time (diagnostics)
        per = len(log.records) // steps
        with open(os.environ['XNRS_BENCH_DUMP'] + '.' + model_key, 'w') as f:
            for r in log.records[:per]:
                f.write(json.dumps({'name': r.name, 'kernel': r.kernel, 'args': list(r.args), 'ms': round(r.start.elapsed_time(r.end), 4)}) + '\n')

    # ---- timed region 2: end to end through the public trainer API with HOST (pinned) index buffers ---------------
    def h2d(i):                              # host (pinned) int32 ids / targets / labels -> device, on the main stream
        b = syn.index_batch(store, cat, pinned[i % n_batches], dev, abstract_store=astore, categories=(model_key in ('naml', 'lstur')))
        return b, torch.cuda.current_stream().record_event()

    last = [0.0]

    # every step's loss is copied to pinned host memory (asynchronously, right behind the step) and READ by the host one step
    # later, i.e. after the next step has been queued: a training loop that logs the loss with one step of lag keeps the GPU
    # queue non-empty.  --e2e-sync-read reads each loss before the next step is queued (the GPU then idles for the host's
    # wake-up + launch time every step).
    loss_host = [torch.zeros(1, dtype=torch.float32).pin_memory() for _ in range(2)]
    loss_ev = [torch.cuda.Event() for _ in range(2)]

    def region_e2e():
        depth = max(1, min(args.prefetch_depth, n_batches - 1))
        queue = []
        for d in range(depth):               # input pipeline `depth` batches deep: ids copied from pinned memory and planned ahead
            queue.append(h2d(d))
            if not args.no_prefetch:
                dp.prefetch(queue[-1][0], after=queue[-1][1])
        t0 = time.perf_counter()
        for i in range(steps):
            nxt = h2d(i + depth)             # tiny copy, queued BEFORE this step so the side stream can plan it meanwhile
            cur = queue.pop(0)
            queue.append(nxt)
            out = run_step(cur[0])
            if args.e2e_sync_read:
                last[0] = float(out['loss'])     # device -> host read of the step's loss (4 bytes, synchronises)
            else:
                loss_host[i & 1].copy_(out['loss'].reshape(1), non_blocking=True)      # device -> host, 4 bytes, every step
                loss_ev[i & 1].record()
            if not args.no_prefetch:
                dp.prefetch(nxt[0], after=nxt[1])
            if not args.e2e_sync_read and i > 0:
                loss_ev[(i - 1) & 1].synchronize()
                last[0] = float(loss_host[(i - 1) & 1])
        torch.cuda.synchronize()
        if not args.e2e_sync_read:
            last[0] = float(loss_host[(steps - 1) & 1])
        return (time.perf_counter() - t0) * 1e3

    region_e2e()
    e_list = ctx.timed_regions(region_e2e)
    e_med, e_sp = spread(e_list, B * world * steps)
    e2e_value = B * world * steps / (e_med * 1e-3)

    out = {
        'metric': 'train impressions/s', 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': steps, 'warmup': args.warmup,
        'ms_per_step': ms_med / steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        # (bf16x3: a model none of whose launches takes the split-plane GEMM — NPA — runs 3xTF32 throughout and says so)
        'dtype': {'fp32': 'f32', 'tf32x3': 'f32 (3xTF32)', 'tf32': 'tf32', 'bf16': 'bf16',
                  'bf16x3': ('f32 (3-pass fp16/bf16 split on the token-level GEMMs, 3xTF32 elsewhere)'
                             if any(r.name.endswith('bf16x3') for r in log.records) else 'f32 (3xTF32)')}[args.precision],
        'data': 'synthetic',
        'config': {'workload': workload_name(B, model_key), 'global_batch': B * world, 'parallelism': f'dp{world}',
                   'l2': 'inputs larger than L2: 307 MB token table, ~0.5 GB of gathered rows per step, 8 batches cycled',
                   'precision': args.precision, 'final_loss': last[0], 'infonce_exchange': dp.infonce_exchange,
                   'dedup_titles': not args.no_dedup, 'skip_padding': not args.no_skip_padding,
                   'cuda_graph': ({'graph_launches_per_region': graph_launches, 'replays': stepper.replays, 'captures': stepper.captures,
                                   'eager_steps': stepper.eager_steps, 'shape_buckets': len(stepper.graphs)} if stepper else False),
                   'timing': f'median of {sp["regions"]} regions of exactly {steps} steps ({sp["timed_s"]} s measured)'},
        'spread': sp,
        'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': h2d_bytes, 'd2h_bytes_per_step': 4,
                'spread': e_sp, 'loss_readback': 'synchronous, before the next step is queued' if args.e2e_sync_read else
                'every step: async copy to pinned memory, read by the host after the NEXT step is queued (one step of lag)',
                'note': 'host input = int32 news ids / targets / labels (the index fast path of the drop-in API)'},
        'gpu_launches': launches, 'roofline': roofline, 'clocks': clocks.summary(),
    }
    del dp, trainer, model, resident, store, astore, stepper, run_step
    torch.cuda.empty_cache()
    if world == 1 and want_eager and not args.no_cpu_baseline:
        out['eager_gpu_baseline'] = eager_gpu_baseline(dev)
    if world == 1 and want_cpu and not args.no_cpu_baseline:
        out['cpu_baseline'] = cpu_baseline_train(model_key, args.ref_batch if model_key != 'nrms' else max(8, args.ref_batch // 4))
    return out


# =====================================================================================================================
# eval workload (BASELINE.json configs[4])
# =====================================================================================================================

def eval_workload(ctx, args, want_cpu=True):
    """MIND-large-shaped full-catalogue evaluation with the CL/standard model: encode the catalogue once (sharded over
    ranks, all-gathered), then per-impression user encoding + candidate scoring + AUC/MRR/nDCG (strong scaling)."""
    from xnrs_b200 import kernels as K
    from xnrs_b200 import synthetic as syn
    from xnrs_b200.data import TitleStore
    from xnrs_b200.evaluation import CatalogueEvaluator
    from xnrs_b200.models import make_model

    dev, world, rank = ctx.dev, ctx.world, ctx.rank
    K.set_precision(args.precision)
    n_news, n_imp, passes = EVAL_NEWS, args.eval_impressions, args.steps
    cat = syn.make_catalogue(n_news, SEQ_LEN, VOCAB, 768, seed=0)
    imp = syn.make_eval_impressions(n_news, n_imp, HIST_LEN, seed=1)
    store = TitleStore(cat.token_table.to(dev), cat.title_tokens.to(dev))
    torch.manual_seed(0)
    model = make_model(CL_CFG).to(dev).eval()
    ev = CatalogueEvaluator(model, store, news_chunk=16384, impression_chunk=16384)
    imp_dev = {k: v.to(dev) for k, v in imp.items()}
    imp_pin = {k: v.pin_memory() for k, v in imp.items()}
    phase = {'enc': 0.0, 'score': 0.0, 'n': 0}
    result = [None]

    def one_pass(src):
        ev.news_vecs = None
        t0, t1, t2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        t0.record()
        ev.encode_catalogue()
        t1.record()
        result[0] = ev.evaluate(src)
        t2.record()
        return t0, t1, t2

    def make_region(src, wall):
        def region():
            w0 = time.perf_counter()
            evs = [one_pass(src) for _ in range(passes)]
            torch.cuda.synchronize()
            if wall:
                return (time.perf_counter() - w0) * 1e3
            for t0, t1, t2 in evs:
                phase['enc'] += t0.elapsed_time(t1)
                phase['score'] += t1.elapsed_time(t2)
                phase['n'] += 1
            return sum(t0.elapsed_time(t2) for t0, _, t2 in evs)
        return region

    for _ in range(max(1, args.warmup // 3)):
        one_pass(imp_dev)
    n0 = K.launch_count()
    with ClockSampler(ctx.local) as clocks:
        ms_list = ctx.timed_regions(make_region(imp_dev, False))
    launches = (K.launch_count() - n0) // len(ms_list)
    ms_med, sp = spread(ms_list, n_imp * passes)
    e_list = ctx.timed_regions(make_region(imp_pin, True))   # host CSR buffers -> H2D of this rank's shard inside the region
    e_med, e_sp = spread(e_list, n_imp * passes)
    out_metrics = result[0]
    shard = dict(ev.last_shard)
    # one more pass with every entry point bracketed by CUDA events on its stream: per-kernel durations
    log = HookLog()
    K.set_event_hook(log)
    ctx.sync_all()
    one_pass(imp_dev)
    ctx.sync_all()
    K.set_event_hook(None)
    per_entry, gemm_flop, gemm_kernels = {}, 0.0, {}
    for r in log.records:
        dt = r.start.elapsed_time(r.end)
        per_entry[r.name] = per_entry.get(r.name, 0.0) + dt
        if r.name == 'xnrs_gemm':
            gemm_flop += 2.0 * r.args[2] * r.args[3] * r.args[4]
            gemm_kernels[r.kernel] = gemm_kernels.get(r.kernel, 0.0) + dt
        elif r.name.startswith('xnrs_titlepool_fwd'):       # gather -> fc1 -> tanh -> logit -> exp -> per-title sums: 2 * rows * A * F
            gemm_flop += 2.0 * r.args[1] * r.args[3] * r.args[4]
            kname = ('gemm_tc2_kernel<POOL> fused title pooling ('
                     + ('bf16 kind::f16' if r.name.endswith('bf16') else ('3-pass 16-bit split, pre-split fp16 planes' if r.name.endswith('bf16x3') else '3xTF32'))
                     + ', cp.async gather)')
            gemm_kernels[kname] = gemm_kernels.get(kname, 0.0) + dt
    pk, pk_kind = peaks()
    # algorithmic bytes of THIS RANK's score+rank launches: one T-wide fp32 vector per candidate + the user vector + ids/targets
    score_bytes = shard['candidates'] * (256 * 4 + 4 + 4 + 4) + shard['impressions'] * (256 * 4 + 8 + 6 * 8)
    score_ms = per_entry.get('xnrs_eval_impressions')
    gemm_ms = sum(v for k, v in per_entry.items() if k == 'xnrs_gemm' or k.startswith('xnrs_titlepool_fwd')) or None
    h2d_shard = (shard['impressions'] * (HIST_LEN * 4 + 8 + 4) + shard['candidates'] * 8)
    res = {
        'metric': 'eval scored impressions/s', 'value': n_imp * passes / (ms_med * 1e-3), 'unit': UNIT, 'n_gpus': world,
        'steps': passes, 'warmup': max(1, args.warmup // 3), 'ms_per_step': ms_med / passes, 'higher_is_better': True,
        'scaling': 'strong', 'vs_baseline': None, 'dtype': {'tf32x3': 'f32 (3xTF32)', 'bf16x3': 'f32 (3-pass fp16/bf16 split on the token-level GEMMs, 3xTF32 elsewhere)'}.get(args.precision, args.precision),
        'data': 'synthetic',
        'config': {'workload': f'MIND-large-shaped full-catalogue eval, model=standard (mind_standard.yml): {n_news} news '
                               f'encoded once, {n_imp} impressions x ~37 candidates, H={HIST_LEN}, S={SEQ_LEN}; one step = '
                               f'catalogue encode + all impressions', 'parallelism': f'news+impression sharding x{world}',
                   'l2': 'news vectors 164 MB + 14M candidate gathers exceed L2',
                   'catalogue_encode_ms': phase['enc'] / max(phase['n'], 1), 'impression_phase_ms': phase['score'] / max(phase['n'], 1),
                   'metrics': {k: out_metrics[k] for k in ('auc', 'rr', 'ndcg@5', 'ndcg@10')},
                   'timing': f'median of {sp["regions"]} regions of exactly {passes} passes ({sp["timed_s"]} s measured)'},
        'spread': sp,
        'e2e': {'value': n_imp * passes / (e_med * 1e-3), 'unit': UNIT, 'h2d_bytes_per_step': h2d_shard, 'd2h_bytes_per_step': 56,
                'spread': e_sp, 'note': 'per rank: only its CSR shard crosses PCIe'},
        'gpu_launches': launches // max(passes, 1),
        # dominant kernel class of the pass: the tcgen05 GEMMs of the catalogue encode (title fc1 + heads + per-article pooling
        # logits), FLOPs = sum of 2MNK over this rank's launches / their CUDA-event time
        'roofline': {'kernel': 'tensor-core kernels of the pass: ' + ', '.join(f'{k} {v:.2f} ms' for k, v in sorted(gemm_kernels.items(), key=lambda kv: -kv[1])),
                     'bound': 'tensor', 'achieved': gemm_flop / (gemm_ms * 1e-3) / 1e12 if gemm_ms else None,
                     'peak': pk['bf16_tflops_sustained'], 'unit': 'TFLOP/s',
                     'frac': (gemm_flop / (gemm_ms * 1e-3) / 1e12 / pk['bf16_tflops_sustained']) if gemm_ms else None,
                     'traffic': None,
                     'peak_source': f'{pk_kind} (sustained bf16 GEMM). fp32-accurate arithmetic: 3xTF32 has 1/6 of this peak as its own '
                                    'ceiling, the 3-pass 16-bit split 1/3; per rank',
                     'per_entry_point_ms': {k: round(v, 3) for k, v in sorted(per_entry.items(), key=lambda kv: -kv[1])}},
        # the HBM-bound scoring kernel, PER RANK: this rank's algorithmic bytes / this rank's event time
        'roofline_scoring': {'kernel': 'eval_impressions_warp_kernel (gather + dot + segmented rank sort + metrics)', 'bound': 'hbm',
                             'achieved': score_bytes / (score_ms * 1e-3) / 1e9 if score_ms else None,
                             'peak': pk['hbm_gbs'], 'unit': 'GB/s',
                             'frac': score_bytes / (score_ms * 1e-3) / 1e9 / pk['hbm_gbs'] if score_ms else None,
                             'peak_source': f'{pk_kind}; rank 0 of {world}: {shard["impressions"]} impressions, {shard["candidates"]} candidates'},
        'clocks': clocks.summary(),
    }
    sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    del ev, model, store, imp_dev, imp_pin
    torch.cuda.empty_cache()
    if rank == 0 and world == 1 and want_cpu and not args.no_cpu_baseline:
        res['cpu_baseline'] = cpu_baseline_eval(cat, imp, sd, min(args.ref_eval_sample, n_imp))
    return res


# =====================================================================================================================
# --impl reference
# =====================================================================================================================

def run_reference(args):
    """the reference's own CPU implementation of the path on the box's host cores — the unmodified reference package when
    oracle/_ref travelled with the repo (kind "reference"), else the oracle port — with all host threads, same metric /
    config, each step a bounded sample of the workload.  Under torchrun rank 0 alone runs it."""
    if int(os.environ.get('RANK', '0')) != 0:
        return
    threads = os.cpu_count() or 1

    def train_line(model_key, bs, steps, warmup):
        value, per_step, kind = time_cpu_train(model_key, bs, steps, warmup, threads)
        sample = f'{bs} impressions/step x {steps} steps (bounded sample of the {args.batch}/GPU workload)'
        return {
            'impl': 'reference', 'metric': 'train impressions/s', 'value': value, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': steps,
            'warmup': warmup, 'ms_per_step': per_step * 1e3, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': workload_name(args.batch, model_key), 'reference_sample': sample},
            'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': threads, 'kind': kind, 'sample': sample},
            'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        }

    def eval_line():
        from xnrs_b200 import synthetic as syn
        from xnrs_b200.models import make_model
        n = min(args.ref_eval_sample, args.eval_impressions)
        cat = syn.make_catalogue(EVAL_NEWS, SEQ_LEN, VOCAB, 768, seed=0)
        imp = syn.make_eval_impressions(EVAL_NEWS, max(n, 1000), HIST_LEN, seed=1)
        torch.manual_seed(0)
        base = cpu_baseline_eval(cat, imp, make_model(CL_CFG).state_dict(), n)
        return {'impl': 'reference', 'metric': 'eval scored impressions/s', 'value': base['value'], 'unit': UNIT, 'n_gpus': args.gpus,
                'steps': 1, 'warmup': 0, 'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f32',
                'data': 'synthetic', 'config': {'workload': 'MIND-large-shaped full-catalogue eval, model=standard', 'reference_sample': base['sample']},
                'cpu_baseline': base, 'e2e': {'value': base['value'], 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}

    only = args.only
    if only == 'eval':
        line = eval_line()
    elif only:
        line = train_line(only, args.ref_batch, args.steps, args.warmup)
    else:
        line = train_line('cl', args.ref_batch, args.steps, args.warmup)
        sub = {}
        try:
            sub['nrms_train'] = train_line('nrms', max(8, args.ref_batch // 4), 2, 1)
        except Exception as exc:
            sub['nrms_train'] = {'impl': 'reference', 'unavailable': repr(exc)}
        sub['eval'] = eval_line()
        line['sub'] = sub
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--batch', type=int, default=1024, help='impressions per GPU per step')
    ap.add_argument('--ref-batch', type=int, default=64, help='impressions per CPU step (reference arm / cpu_baseline)')
    ap.add_argument('--ref-eval-sample', type=int, default=400, help='impressions of the eval workload timed on the CPU')
    ap.add_argument('--precision', default='bf16x3', choices=['fp32', 'tf32x3', 'tf32', 'bf16', 'bf16x3'],
                    help="bf16x3 (default) and tf32x3 are the fp32-accurate modes (1e-4 parity class): a 3-pass 16-bit split on pre-split planes for "
                         "the token-level tensor-core launches + 3xTF32 elsewhere, or 3xTF32 everywhere; bf16 = the bf16 STORAGE mode "
                         "(2e-2 class)")
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--only', default=None, choices=list(MODEL_CFGS) + ['eval'],
                    help='run ONE workload instead of the headline + sub lines')
    ap.add_argument('--workload', default=None, choices=['train', 'eval'], help='(round-1 spelling) eval == --only eval')
    ap.add_argument('--model', default=None, choices=list(MODEL_CFGS), help='(round-1 spelling) == --only MODEL')
    ap.add_argument('--e2e-sync-read', action='store_true',
                    help='e2e: read each step\'s loss synchronously before queueing the next step (default: one step of lag)')
    ap.add_argument('--no-skip-padding', action='store_true', help='run pad tokens through the encoder like the reference does')
    ap.add_argument('--prefetch-depth', type=int, default=2,
                    help='how many batches ahead the input pipeline copies / plans ids (the plan kernels only get SMs at kernel '
                         'boundaries of the running step: one batch ahead, the next step waited for its plan)')
    ap.add_argument('--no-prefetch', action='store_true', help='compute the id plumbing of each batch inside its own step')
    ap.add_argument('--no-graph', action='store_true', help='launch the CL step kernel by kernel instead of replaying its CUDA graph')
    ap.add_argument('--no-graph-multi-gpu', action='store_true', help='under torchrun, launch kernel by kernel (graphs capture the NCCL collectives too)')
    ap.add_argument('--graph-multi-gpu', action='store_true', help='(default now) replay CUDA graphs under torchrun as well')
    ap.add_argument('--no-dedup', action='store_true', help='encode every (impression, slot) title, not each distinct article once')
    ap.add_argument('--eval-impressions', type=int, default=376_471)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.workload == 'eval':
        args.only = 'eval'
    elif args.model:
        args.only = args.model
    if args.impl == 'reference':
        return run_reference(args)

    ctx = Ctx(args.gpus)
    if args.only == 'eval':
        line = eval_workload(ctx, args)
    elif args.only:
        line = train_workload(ctx, args, args.only, want_eager=(args.only == 'cl'))
    else:
        line = train_workload(ctx, args, 'cl', want_eager=True)
        line['sub'] = {'nrms_train': train_workload(ctx, args, 'nrms'), 'eval': eval_workload(ctx, args)}
        if args.precision in ('tf32x3', 'bf16x3'):
            import copy
            if args.precision == 'bf16x3':      # the same step with 3xTF32 on every tensor-core launch (the round-2a headline mode)
                t3 = copy.copy(args)
                t3.precision = 'tf32x3'
                line['sub']['cl_train_3xtf32'] = train_workload(ctx, t3, 'cl', want_cpu=False)
            # second lines in the bf16 STORAGE mode (north-star tolerance class 2e-2): beside the fp32-accurate numbers, never
            # instead of them
            bf = copy.copy(args)
            bf.precision = 'bf16'
            line['sub']['cl_train_bf16'] = train_workload(ctx, bf, 'cl', want_cpu=False)
            line['sub']['eval_bf16'] = eval_workload(ctx, bf, want_cpu=False)
    if ctx.rank == 0:
        print(json.dumps(line))
    ctx.close()


if __name__ == '__main__':
    main()
