"""Rebuild profiles/README.md and copy the round's ncu launch lists from gpurun_out/ (run after the gpurun captures)."""
import collections, csv, json, os, re, shutil, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, 'gpurun_out'), os.path.join(ROOT, 'profiles')


def launch_summary(path, one_step=False):
    rows = list(csv.DictReader([l for l in open(path) if not l.startswith('==')]))
    names = [re.sub(r'\(.*', '', r['Kernel Name']).replace('void ', '') for r in rows]
    if one_step:
        idx = [i for i, n in enumerate(names) if 'adam_kernel' in n]
        if len(idx) >= 2:
            rows, names = rows[idx[0] + 1:idx[1] + 1], names[idx[0] + 1:idx[1] + 1]
    agg, tot = collections.OrderedDict(), 0.0
    for r, n in zip(rows, names):
        ms = float(r['Metric Value'].replace(',', '')) * {'ns': 1e-6, 'us': 1e-3, 'ms': 1}.get(r['Metric Unit'], 1e-6)
        a = agg.setdefault(n, [0, 0.0]); a[0] += 1; a[1] += ms; tot += ms
    mine = sum(v[1] for k, v in agg.items() if k.startswith('xnrs::'))
    out = [f'{len(rows)} launches, {tot:.3f} ms of device time (cold-cache, serialised: compare SHARES); '
           f'{100 * mine / tot:.1f} % of it in this repo\'s kernels (`xnrs::*`), the rest is torch plumbing '
           f'(fills, gradient accumulation adds, id sort/unique/index)', '', '| kernel | launches | ms | share |', '|---|---|---|---|']
    for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:18]:
        out.append(f'| `{k[:100]}` | {n} | {ms:.3f} | {100 * ms / tot:.2f} % |')
    return '\n'.join(out)


def ncu_rows(path, want):
    txt = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    return [{w: (r[hdr.index(w)] + ' ' + units[hdr.index(w)]).strip() for w in want if w in hdr} for r in rows[2:]]


WANT = ['Kernel Name', 'Grid Size', 'Block Size', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread']


def bench_line(fname):
    lines = [l for l in open(os.path.join(G, fname)) if l.startswith('{')]
    return json.loads(lines[-1]) if lines else None


md = ['# profiles — round 1', '',
      'All captures on one B200 (sm_100a, 148 SMs) through `gpurun`; every ncu run was preceded by the same command exiting 0 '
      'without ncu.  Numbers under a profiler are never bench values: the bench lines below come from plain runs.', '',
      '## bench lines (plain runs, final round-1 build)', '']
for f, title in [('bench_final.log', 'headline: CL train, 1 x B200 (`python bench.py`)'), ('bench_ref.log', 'reference arm (`--impl reference`, oracle port on the box CPU)'),
                 ('bench_nrms.log', '`--model nrms`'), ('bench_naml.log', '`--model naml`'), ('bench_lstur.log', '`--model lstur`'),
                 ('bench_npa.log', '`--model npa`'), ('bench_eval.log', '`--workload eval` (MIND-large-shaped full-catalogue evaluation)'),
                 ('bench_train_2gpu.log', 'train, 2 x B200 (torchrun; earlier build of the round)'), ('bench_eval_2gpu.log', 'eval, 2 x B200'),
                 ('bench_train_8gpu.log', 'train, 8 x B200'), ('bench_eval_8gpu.log', 'eval, 8 x B200')]:
    try:
        d = bench_line(f)
    except Exception:
        d = None
    if not d:
        continue
    keep = ('metric', 'value', 'unit', 'n_gpus', 'steps', 'ms_per_step', 'dtype', 'e2e', 'gpu_launches', 'clocks', 'cpu_baseline', 'impl')
    md += [f'### {title}', '', '```json', json.dumps({k: d[k] for k in keep if k in d}), '```', '']
    if 'roofline' in d:
        r = d['roofline']
        md += ['roofline: `' + json.dumps({k: r[k] for k in r if k not in ('per_entry_point_ms_per_step', 'gemm_shapes_ms_per_step')}) + '`', '']
        if 'per_entry_point_ms_per_step' in r:
            md += ['per entry point, ms/step (CUDA events on the launching stream): `' + json.dumps(r['per_entry_point_ms_per_step']) + '`', '']
for src, dst, title, one in [('launches_r01_final.csv', 'r01_launches_final.csv', 'ncu launch list — final build, one training step (CL, 3xTF32, dedup + padding-free)', True),
                             ('launches_r01.csv', 'r01_launches_fp32_simt.csv', 'ncu launch list — first parity-green build (exact-fp32 SIMT GEMMs, every slot encoded)', False)]:
    if os.path.exists(os.path.join(G, src)):
        shutil.copy(os.path.join(G, src), os.path.join(P, dst))
    if os.path.exists(os.path.join(P, dst)):
        md += [f'## {title} (`{dst}`)', '', launch_summary(os.path.join(P, dst), one), '']
for rep, title in [('prof_gemm_tc_final_r01.ncu-rep', 'ncu --set full: `gemm_tc_kernel<128>` in the final build — the dominant launch (title fc1 forward, M≈150 k rows, N=256, K=768, 3xTF32) and the next launch'),
                   ('prof_gemm_tc_big_r01.ncu-rep', 'ncu --set full: the same kernel before dedup (M=1 536 000; hi+lo written by the splitters)'),
                   ('prof_tc2_3x.ncu-rep', 'ncu --set full: opt-in CTA-pair kernel `gemm_tc2_kernel` (cta_group::2, 256x256 tile), 3xTF32'),
                   ('prof_gemm_simt_r01.ncu-rep', 'ncu --set full: `gemm_simt_kernel` (exact fp32)')]:
    path = os.path.join(G, rep)
    if os.path.exists(path):
        md += [f'## {title}', '']
        for r in ncu_rows(path, WANT)[:2]:
            md += ['```', json.dumps(r, indent=1), '```', '']
md += ['Reading the GEMM capture: DRAM traffic equals the algorithmic A + C bytes (B stays in L2; no re-reads).  Tensor pipe ≈ 38 % active, '
       'DRAM 15 %, L2 25 %, L1/shared 44 %: with the smem traffic of the hi/lo split reduced (lo-only writes) nothing is saturated — the '
       '3-stage TMA -> split -> MMA pipeline is latency-bound.  The CTA-pair kernel halves operand bytes per FLOP and reaches 50 % tensor-pipe '
       'activity but its cross-CTA handshake lengthens the per-stage critical path (168 vs 179 TFLOP/s), so it stays opt-in.', '']
if os.path.exists(os.path.join(G, 'bench_gemm.log')):
    shutil.copy(os.path.join(G, 'bench_gemm.log'), os.path.join(P, 'r01_gemm_microbench.jsonl'))
if os.path.exists(os.path.join(P, 'r01_gemm_microbench.jsonl')):
    md += ['## GEMM micro-benchmark (`tools/bench_gemm.py` -> `r01_gemm_microbench.jsonl`)', '', '```'] + \
          [l.strip() for l in open(os.path.join(P, 'r01_gemm_microbench.jsonl')) if l.startswith('{')] + ['```', '']
open(os.path.join(P, 'README.md'), 'w').write('\n'.join(md))
print('wrote', os.path.join(P, 'README.md'))
