"""Rebuild profiles/README.md for round 2 and copy the round's captures from gpurun_out/ into profiles/ (run after the gpurun
captures).  Every number comes from a plain run; ncu runs followed a plain run of the same command.

Inputs (gpurun_out/):
  bench_r02_final.json / bench_r02_ref.json     python bench.py / python bench.py --impl reference   (1 x B200)
  bench_r02_{2,4,8}gpu.json                     torchrun ... bench.py --gpus N
  bench_{naml,lstur}.json                       secondary models
  launches_r02_cl.csv                           ncu launch list of the CL bench command (kernel by kernel, --no-graph)
  prof_titlepool.ncu-rep -> profiles/r02_ncu_token_kernels.jsonl   ncu --set full of the token-level kernels (tools/prof_titlepool.py)
  gather_gemm6.jsonl, gather_pf.log             tools/bench_gather_gemm.py
"""
import collections
import csv
import json
import os
import re
import shutil

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, 'gpurun_out'), os.path.join(ROOT, 'profiles')


def line(fname):
    path = os.path.join(G, fname)
    if not os.path.exists(path):
        return None
    rows = [l for l in open(path) if l.startswith('{')]
    return json.loads(rows[-1]) if rows else None


def launch_summary(path, top=18):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
    hdr, data = rows[hi], rows[hi + 2:]
    ik, iv, iu = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
    agg, tot = collections.OrderedDict(), 0.0
    for r in data:
        if len(r) <= iv:
            continue
        n = re.sub(r'\(.*', '', r[ik]).replace('void ', '')
        us = float(r[iv].replace(',', '')) * {'ns': 1e-3, 'us': 1.0, 'ms': 1e3}.get(r[iu], 1e-3)
        a = agg.setdefault(n, [0, 0.0])
        a[0] += 1
        a[1] += us
        tot += us
    mine = sum(v[1] for k, v in agg.items() if k.startswith('xnrs::'))
    out = [f'{len(data)} launches, {tot / 1e3:.3f} ms of device time (cold-cache, serialised: compare SHARES); '
           f'{100 * mine / tot:.1f} % of it in this repo\'s kernels (`xnrs::*`), the rest is torch plumbing (fills, index / copy kernels)',
           '', '| kernel | launches | ms | share |', '|---|---|---|---|']
    for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        out.append(f'| `{k[:80]}` | {n} | {us / 1e3:.3f} | {100 * us / tot:.2f} % |')
    return '\n'.join(out)


def brief(x):
    keep = {k: x[k] for k in ('metric', 'value', 'unit', 'n_gpus', 'steps', 'ms_per_step', 'dtype', 'gpu_launches', 'clocks') if k in x}
    keep['e2e'] = {k: v for k, v in x['e2e'].items() if k in ('value', 'h2d_bytes_per_step', 'd2h_bytes_per_step')}
    keep['spread'] = x.get('spread')
    if 'cpu_baseline' in x:
        keep['cpu_baseline'] = x['cpu_baseline']
    return keep


def roofline_brief(r):
    if not r:
        return None
    out = {k: r.get(k) for k in ('kernel', 'bound', 'achieved', 'peak', 'unit', 'frac', 'traffic', 'algorithmic_bytes', 'traffic_source',
                                 'launches_timed', 'avg_launch_ms', 'all_gemm_tflops', 'kernel_ms_per_step', 'note') if r.get(k) is not None}
    for k in ('fused_title_pool', 'bf16_weight_gradient'):
        if k in r:
            out[k] = r[k]
    return out


def main():
    os.makedirs(P, exist_ok=True)
    md = ['# profiles — round 2 (final build of the round: bench default = fp32-accurate 3-pass 16-bit split on the token-level launches)', '',
          'All captures on B200 (sm_100a, 148 SMs) through `gpurun`; every ncu run was preceded by the same command exiting 0 without '
          'ncu.  Numbers taken under a profiler are never bench values: the bench lines below come from plain runs.  Round-1 summaries: '
          '`README_r01.md`, `README_r01_session1.md`.', '']
    final = line('bench_r02_final.json')
    if final:
        json.dump(final, open(os.path.join(P, 'r02_bench_final.json'), 'w'), indent=1)
        md += ['## `python bench.py` (1 x B200): CL headline + sub lines', '', 'Full line: `profiles/r02_bench_final.json`.', '']
        for name, x in [('headline: CL train, fp32-accurate (3-pass fp16/bf16 split on the token-level GEMMs, 3xTF32 elsewhere)', final)] + [(f'sub.{k}', v) for k, v in final.get('sub', {}).items()]:
            md += [f'### {name}', '', '```json', json.dumps(brief(x)), '```', '', 'roofline: `' + json.dumps(roofline_brief(x.get('roofline'))) + '`', '']
            r = x.get('roofline') or {}
            if 'per_entry_point_ms_per_step' in r:
                md += ['per entry point, ms per step (CUDA events on the launching stream, kernel-by-kernel pass): `' + json.dumps(r['per_entry_point_ms_per_step']) + '`', '']
            if 'per_entry_point_ms' in r:
                md += ['per entry point, ms per pass: `' + json.dumps(r['per_entry_point_ms']) + '`', '']
            if 'roofline_scoring' in x:
                md += ['roofline_scoring: `' + json.dumps(x['roofline_scoring']) + '`', '']
            if 'eager_gpu_baseline' in x:
                md += ['eager_gpu_baseline (the unmodified reference modules + trainer step on the same B200, stock PyTorch eager): `'
                       + json.dumps(x['eager_gpu_baseline']) + '`', '']
    ref = line('bench_r02_ref.json')
    if ref:
        json.dump(ref, open(os.path.join(P, 'r02_bench_ref.json'), 'w'), indent=1)
        md += ['## `python bench.py --impl reference` (the unmodified reference package, `oracle/_ref`, on the box CPU)', '', '```json',
               json.dumps({k: (brief(v) if k != 'sub' else {kk: brief(vv) for kk, vv in v.items() if 'value' in vv}) if isinstance(v, dict) and 'value' in v else v
                           for k, v in {'headline': ref, 'sub': ref.get('sub', {})}.items()}), '```', '']
    for n in (2, 4, 8):
        x = line(f'bench_r02_{n}gpu.json')
        if x:
            json.dump(x, open(os.path.join(P, f'r02_bench_{n}gpu.json'), 'w'), indent=1)
            md += [f'## torchrun, {n} x B200 (`bench.py --gpus {n}`)', '', '```json',
                   json.dumps({'cl': brief(x), **{k: brief(v) for k, v in x.get('sub', {}).items()}}), '```', '']
    x = line('bench_r02_8gpu_nrms.json')
    if x:
        json.dump(x, open(os.path.join(P, 'r02_bench_8gpu_nrms.json'), 'w'), indent=1)
        md += ['## torchrun, 8 x B200, `--only nrms` (after the self-attention projections moved to the split-plane GEMMs; the `nrms_train` '
               'sub line of the 8-GPU run above predates that)', '', '```json', json.dumps(brief(x)), '```', '']
    for name in ('naml', 'lstur', 'npa'):
        x = line(f'bench_{name}.json')
        if x:
            json.dump(x, open(os.path.join(P, f'r02_bench_{name}.json'), 'w'), indent=1)
            md += [f'## `--only {name}`', '', '```json', json.dumps(brief(x)), '```', '']
    lp = os.path.join(G, 'launches_r02_cl.csv')
    if os.path.exists(lp):
        shutil.copy(lp, os.path.join(P, 'r02_launches_cl.csv'))
        md += ['## ncu launch list of the CL bench command', '',
               '`XNRS_BENCH_MIN_S=0 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv python bench.py --only cl --steps 3 '
               '--warmup 3 --no-cpu-baseline --no-graph` (default precision bf16x3; kernel by kernel: the graph replays the same kernels), `profiles/r02_launches_cl.csv`:',
               '', launch_summary(lp), '']
    tk = os.path.join(P, 'r02c_ncu_kernels.jsonl')
    if os.path.exists(tk):
        md += ['## ncu --set full: the token-level kernels and two title-level GEMMs at bench shapes (`tools/prof_titlepool.py`, 157.7 k rows, '
               '8.9 k titles; final build)', '',
               '| launch | kernel | time | DRAM read | DRAM write | DRAM % | tensor pipe % | issue active % | regs |', '|---|---|---|---|---|---|---|---|---|']
        seen = set()
        for l in open(tk):
            r = json.loads(l)
            if r['launch'] in seen:
                continue
            seen.add(r['launch'])
            md.append(f"| {r['launch']} | `{r.get('Kernel Name', '')[:34]}` | {r.get('gpu__time_duration.sum')} | {r.get('dram__bytes_read.sum')} | {r.get('dram__bytes_write.sum')} | "
                      f"{r.get('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed')} | {r.get('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active')} | "
                      f"{r.get('smsp__issue_active.avg.pct_of_peak_sustained_active')} | {r.get('launch__registers_per_thread')} |")
        md += ['', '(`titlepool_*` = the tensor-core launch of the fused pooling forward — gather, fc1, tanh, logit, exp, hid through TMA stores; the '
               'per-title weighted sums run in `titlepool_wsum_kernel` afterwards; `dw_*` = fc1 weight gradient with / without the fused table '
               'gather; `*_x3_*` = the 3-pass 16-bit split form on pre-split planes (bench default); `pool_bwd2_*` = title-level pooling backward, '
               'fp32 d_hid or d_hid as two bf16 planes.  Earlier captures of the round: `r02_ncu_token_kernels.jsonl`.)', '']
    sc = [(n, os.path.join(P, f'r02_step_calls_{n}.jsonl')) for n in ('bf16x3', 'tf32x3', 'bf16')]
    if all(os.path.exists(f) for _, f in sc):
        md += ['## every C-ABI call of one CL step, timed alone (`tools/bench_step_gemms.py`: 50 back-to-back launches per call, CUDA events)', '',
               '| call | ' + ' | '.join(f'{n}: calls, us' for n, _ in sc) + ' |', '|---|' + '---|' * len(sc)]
        sums = []
        for n, f in sc:
            rows = [json.loads(l) for l in open(f) if l.startswith('{')]
            sums.append(next(r for r in rows if 'summary_us' in r))
        keys = []
        for sm in sums:
            for k in sm['summary_us']:
                if k not in keys:
                    keys.append(k)
        for k in keys:
            md.append(f'| `{k}` | ' + ' | '.join((f"{sm['summary_us'][k][0]}, {sm['summary_us'][k][1]}" if k in sm['summary_us'] else '—') for sm in sums) + ' |')
        md.append('| **total** | ' + ' | '.join(f"{sm['calls']}, {sm['total_us']}" for sm in sums) + ' |')
        md += ['', 'Before this round\'s second half (3xTF32, build of the 8-GPU run at 3.65 M): `r02_step_calls_tf32x3_before.jsonl` (total 1940 us), '
               '`r02_step_calls_bf16_before.jsonl` (1141 us).  The files also hold the SM-clock anatomy of a small GEMM (`trace` rows) and the '
               'in-graph cost of a GEMM launch vs K (`sweep` rows).', '']
    for f, title in (('gather_gemm6.jsonl', 'fused gather vs gather-then-GEMM (`tools/bench_gather_gemm.py`, CUDA events; "gather4" keys = the fused form)'),
                     ('gather_pf.log', 'L2 prefetch distance of the gathered weight gradient (`XNRS_GATHER_PF`; columns: dense ms, gathered ms, fc1 gathered ms)')):
        path = os.path.join(G, f)
        if os.path.exists(path):
            shutil.copy(path, os.path.join(P, 'r02_' + f))
            md += [f'## {title}', '', '```', open(path).read().strip(), '```', '']
    md += ['## SASS evidence', '', '`profiles/sass_summary.txt` (`python tools/sass_summary.py`): per-kernel counts of UTCHMMA / UTCHMMA.2CTA (tcgen05.mma), LDTM '
           '(tcgen05.ld), UTMALDG (TMA), LDGSTS (cp.async), SYNCS (mbarrier) in the shipped library.', '']
    open(os.path.join(P, 'README.md'), 'w').write('\n'.join(md) + '\n')
    print('profiles/README.md written')


if __name__ == '__main__':
    main()
