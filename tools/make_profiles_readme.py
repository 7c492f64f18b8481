"""Rebuild profiles/README.md and copy the round's captures from gpurun_out/ into profiles/ (run after the gpurun captures).

Inputs (gpurun_out/, all produced by plain runs or by ncu runs that followed a plain run of the same command):
  r01b_bench_*.log            bench.py lines (cl, ref, nrms, naml, lstur, npa, eval, cl_tf32, nrms_tf32, *_Ngpu)
  r01b_bench_kernels.jsonl    tools/bench_kernels.py   (per-kernel achieved GB/s, CUDA events)
  r01b_bench_gemm.jsonl       tools/bench_gemm.py      (GEMM micro-benchmark vs cuBLAS)
  launches_r01b_{cl,nrms,eval}.csv   ncu launch lists of one step / pass (tools/one_step.py)
  prof_r01b_cl_gemms.ncu-rep  ncu --set full of the 20 GEMM launches of one CL step
  r01b_ncu_kernels.jsonl      ncu --set full of every non-GEMM kernel, reduced by tools/ncu_extract.py
"""
import collections
import csv
import json
import os
import re
import shutil
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, 'gpurun_out'), os.path.join(ROOT, 'profiles')


def launch_summary(path, top=16):
    rows = [r for r in csv.DictReader([l for l in open(path) if not l.startswith('==')]) if r.get('Metric Name') == 'gpu__time_duration.sum']
    agg, tot = collections.OrderedDict(), 0.0
    for r in rows:
        n = re.sub(r'\(.*', '', r['Kernel Name']).replace('void ', '')
        ms = float(r['Metric Value'].replace(',', '')) * {'ns': 1e-6, 'nsecond': 1e-6, 'us': 1e-3, 'usecond': 1e-3, 'ms': 1}.get(r['Metric Unit'], 1e-6)
        a = agg.setdefault(n, [0, 0.0])
        a[0] += 1
        a[1] += ms
        tot += ms
    mine = sum(v[1] for k, v in agg.items() if k.startswith('xnrs::'))
    out = [f'{len(rows)} launches, {tot:.3f} ms of device time (cold-cache, serialised: compare SHARES); '
           f'{100 * mine / tot:.1f} % of it in this repo\'s kernels (`xnrs::*`), the rest is torch plumbing '
           f'(fills, gradient-accumulation adds, id sort / unique / compaction)', '', '| kernel | launches | ms | share |', '|---|---|---|---|']
    for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        out.append(f'| `{k[:90]}` | {n} | {ms:.3f} | {100 * ms / tot:.2f} % |')
    return '\n'.join(out)


def bench_line(fname):
    path = os.path.join(G, fname)
    if not os.path.exists(path):
        return None
    lines = [l for l in open(path) if l.startswith('{')]
    return json.loads(lines[-1]) if lines else None


def ncu_rows(path):
    txt = subprocess.run(['python', os.path.join(ROOT, 'tools', 'ncu_extract.py'), path], capture_output=True, text=True).stdout
    return [json.loads(l) for l in txt.splitlines() if l.startswith('{')]


def short(d):
    ren = {'gpu__time_duration.sum': 'time', 'dram__bytes_read.sum': 'dram_read', 'dram__bytes_write.sum': 'dram_write',
           'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed': 'dram_%', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active': 'tensor_pipe_%',
           'lts__throughput.avg.pct_of_peak_sustained_elapsed': 'l2_%', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed': 'l1tex_%',
           'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed': 'smem_lsu_wavefronts_%',
           'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active': 'fma_pipe_%', 'smsp__issue_active.avg.pct_of_peak_sustained_active': 'issue_active_%',
           'sm__warps_active.avg.pct_of_peak_sustained_active': 'warps_active_%', 'launch__registers_per_thread': 'regs',
           'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio': 'stall_long_sb',
           'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio': 'stall_short_sb',
           'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio': 'stall_lg_throttle'}
    out = {}
    for k, v in d.items():
        if k == 'Kernel Name':
            out['kernel'] = re.sub(r'\(.*', '', v).replace('void ', '')
        elif k in ('Grid Size', 'Block Size'):
            out[k.split()[0].lower()] = v
        else:
            m = re.match(r'([-0-9.,]+)\s*(.*)', v)
            out[ren.get(k, k)] = (f'{float(m.group(1).replace(",", "")):.4g} {m.group(2)}'.strip() if m else v)
    return out


md = ['# profiles — round 1 (final build of the round)', '',
      'All captures on one B200 (sm_100a, 148 SMs) through `gpurun`; every ncu run was preceded by the same command exiting 0 '
      'without ncu.  Numbers taken under a profiler are never bench values: the bench lines below come from plain runs.  '
      'The summaries of the first session of this round (before the GEMM / attention / evaluation rework) are kept in '
      '`README_r01_session1.md`.', '',
      '## bench lines (plain runs)', '']
TITLES = [('r01b_bench_cl.log', 'headline: CL train, 1 x B200 (`python bench.py`)'),
          ('r01b_bench_ref.log', 'reference arm (`python bench.py --impl reference`: the oracle port of the reference step on the box CPU)'),
          ('r01b_bench_nrms.log', '`--model nrms`'), ('r01b_bench_naml.log', '`--model naml`'), ('r01b_bench_lstur.log', '`--model lstur`'),
          ('r01b_bench_npa.log', '`--model npa`'), ('r01b_bench_eval.log', '`--workload eval` (MIND-large-shaped full-catalogue evaluation)'),
          ('r01b_bench_cl_tf32.log', 'CL, `--precision tf32` (single-pass TF32: the 2e-2 tolerance class)'),
          ('r01b_bench_nrms_tf32.log', 'NRMS, `--precision tf32`'),
          ('r01b_bench_cl_2gpu.log', 'CL train, 2 x B200 (torchrun, NCCL)'), ('r01b_bench_nrms_2gpu.log', 'NRMS train, 2 x B200'),
          ('r01b_bench_lstur_2gpu.log', 'LSTUR train, 2 x B200 (user-table gradient exchanged as (ids, rows))'),
          ('r01b_bench_cl_4gpu.log', 'CL train, 4 x B200'),
          ('r01b_bench_cl_8gpu.log', 'CL train, 8 x B200 (build of mid-session: before the pooling / launch-path / prefetch work, 2.44 ms single-GPU step)'),
          ('r01b_bench_nrms_8gpu.log', 'NRMS train, 8 x B200 (same mid-session build)'),
          ('r01b_bench_naml_8gpu.log', 'NAML train, 8 x B200 (same mid-session build)'), ('r01b_bench_lstur_8gpu.log', 'LSTUR train, 8 x B200 (same mid-session build)'),
          ('r01b_bench_eval_2gpu.log', 'eval, 2 x B200'), ('r01b_bench_eval_4gpu.log', 'eval, 4 x B200'), ('r01b_bench_eval_8gpu.log', 'eval, 8 x B200 (mid-session build)')]
for f, title in TITLES:
    d = bench_line(f)
    if not d:
        continue
    shutil.copy(os.path.join(G, f), os.path.join(P, f.replace('.log', '.json')))
    keep = ('metric', 'value', 'unit', 'n_gpus', 'steps', 'ms_per_step', 'dtype', 'e2e', 'gpu_launches', 'clocks', 'cpu_baseline', 'impl')
    md += [f'### {title}', '', '```json', json.dumps({k: d[k] for k in keep if k in d}), '```', '']
    for key in ('roofline', 'roofline_scoring'):
        if key in d:
            r = d[key]
            md += [f'{key}: `' + json.dumps({k: r[k] for k in r if k not in ('per_entry_point_ms_per_step', 'gemm_shapes_ms_per_step', 'per_entry_point_ms')}) + '`', '']
            for pe in ('per_entry_point_ms_per_step', 'per_entry_point_ms'):
                if pe in r:
                    md += ['per entry point, ms (CUDA events on the launching stream): `' + json.dumps(r[pe]) + '`', '']

ep = os.path.join(G, 'r01b_bench_eager_gpu.jsonl')
if os.path.exists(ep):
    shutil.copy(ep, os.path.join(P, 'r01b_eager_gpu.jsonl'))
if os.path.exists(os.path.join(P, 'r01b_eager_gpu.jsonl')):
    md += ['### GPU-side bar: the reference step in stock PyTorch eager on the same B200 (`tools/bench_eager_gpu.py`, SURVEY §8(d))', '', '```'] + \
          [l.strip() for l in open(os.path.join(P, 'r01b_eager_gpu.jsonl')) if l.startswith('{')] + ['```', '',
           'i.e. the hand-written path (457 k impr/s, fp32-accurate) is ~40x stock eager with cuBLAS fp32 and ~21x stock eager with TF32 '
           'on the same GPU (dense reference-format batches already resident in HBM for the eager run: its host gather / H2D is not counted).', '']

md += ['## per-kernel roofline micro-benchmark (`tools/bench_kernels.py`, CUDA events, L2 flushed between iterations)', '',
       'Achieved = algorithmic bytes (every operand read / written once) / time, against the measured HBM copy peak.', '',
       '| kernel | ms | algorithmic MB | achieved GB/s | of measured HBM peak | fp32 TFLOP/s | note |', '|---|---|---|---|---|---|---|']
kp = os.path.join(G, 'r01b_bench_kernels.jsonl')
if os.path.exists(kp):
    shutil.copy(kp, os.path.join(P, 'r01b_kernel_roofline.jsonl'))
    for l in open(kp):
        if l.startswith('{'):
            d = json.loads(l)
            md.append(f"| {d['kernel']} | {d['ms']} | {d.get('algorithmic_MB', '')} | {d.get('achieved_GBs', '')} | {d.get('frac', '')} | "
                      f"{d.get('achieved_TFLOPs', '')} | {d.get('note', '')} |")
md.append('')

for src, title in [('launches_r01b_cl.csv', 'ncu launch list — one CL training step (3xTF32, 1024 impressions)'),
                   ('launches_r01b_nrms.csv', 'ncu launch list — one NRMS training step'),
                   ('launches_r01b_eval.csv', 'ncu launch list — one evaluation pass (160k-article catalogue encode + 65 536 impressions)')]:
    if os.path.exists(os.path.join(G, src)):
        shutil.copy(os.path.join(G, src), os.path.join(P, src.replace('launches_r01b', 'r01b_launches')))
    dst = os.path.join(P, src.replace('launches_r01b', 'r01b_launches'))
    if os.path.exists(dst):
        md += [f'## {title} (`{os.path.basename(dst)}`)', '', launch_summary(dst), '']

rep = os.path.join(G, 'prof_r01b_cl_gemms.ncu-rep')
if os.path.exists(rep):
    rows = [short(r) for r in ncu_rows(rep)]
    with open(os.path.join(P, 'r01b_ncu_cl_gemms.jsonl'), 'w') as f:
        for r in rows:
            f.write(json.dumps(r) + '\n')
if os.path.exists(os.path.join(P, 'r01b_ncu_cl_gemms.jsonl')):
    rows = [json.loads(l) for l in open(os.path.join(P, 'r01b_ncu_cl_gemms.jsonl'))]
    big = sorted(rows, key=lambda r: -float(r['time'].split()[0]) * (1000 if r['time'].endswith('ms') else 1))[:4]
    md += ['## ncu --set full: the tcgen05 GEMM launches of one CL step (`r01b_ncu_cl_gemms.jsonl`: all 20; the four longest here)', '']
    for r in big:
        md += ['```', json.dumps(r, indent=1), '```', '']
    md += ['Reading: the dominant launch (title fc1 forward with the tanh epilogue, `gemm_tc2_kernel`, M = 153 562 token rows, N = 256, '
           'K = 768) moves 473 MB + 131 MB of DRAM traffic = 3937 B per row = the algorithmic A row (3072 B) + C row (1024 B): no '
           're-reads.  Tensor pipe 54 % active in 3xTF32 (three MMAs per product), i.e. the kernel issues TF32 MMAs at ~650 TFLOP/s, '
           'the rate cuBLAS reaches in single-pass TF32 on the same shape; what is left is the shared-memory pipe (tensor-core operand '
           'reads + the hi/lo splitter), which the CTA-pair kernel halves per FLOP.', '']

kp = os.path.join(G, 'r01b_ncu_kernels.jsonl')
if os.path.exists(kp):
    with open(os.path.join(P, 'r01b_ncu_kernels.jsonl'), 'w') as f:
        for l in open(kp):
            if l.startswith('{'):
                f.write(json.dumps(short(json.loads(l))) + '\n')
if os.path.exists(os.path.join(P, 'r01b_ncu_kernels.jsonl')):
    md += ['## ncu --set full: every non-GEMM kernel at the headline shapes (`r01b_ncu_kernels.jsonl`, launches of `tools/bench_kernels.py`)', '',
           '| kernel | grid | time | DRAM read | DRAM write | DRAM % | issue active % | warps active % | regs |', '|---|---|---|---|---|---|---|---|---|']
    for l in open(os.path.join(P, 'r01b_ncu_kernels.jsonl')):
        r = json.loads(l)
        md.append(f"| `{r['kernel']}` | {r.get('grid')} | {r.get('time')} | {r.get('dram_read')} | {r.get('dram_write')} | {r.get('dram_%')} | "
                  f"{r.get('issue_active_%')} | {r.get('warps_active_%')} | {r.get('regs')} |")
    md += ['', 'Reading (this capture predates the last pooling / column-sum tuning — the CUDA-event table above is the final build): '
           '`gather_rows` runs at 71 % DRAM utilisation (5.4-5.6 TB/s of algorithmic traffic = 83-86 % of the measured copy peak); the '
           'ragged pooling kernels read exactly their algorithmic bytes (632 MB / 660 MB of DRAM reads for 653 / 810 MB algorithmic) and were '
           'latency-, not bandwidth-limited (issue utilisation 13 % / 30 %: one warp walks one title) — batching four hidden rows per '
           'iteration with hoisted loads took them from 48 % to 54 % / 58 % of the HBM peak; `colsum` went from 54 % to 91 % with 16-byte '
           'loads and four rows in flight per thread; `eval_impressions_warp_kernel` moves only 0.94 GB of DRAM for 2.7 GB of gathered '
           'candidate vectors (Zipf-popular articles are L2 hits) at 62 % issue utilisation; the attention core is bound by shared-memory '
           'broadcast loads and FMA latency at 7-14 resident warps per SM (fwd 36 %, bwd 26 % issue utilisation), not by DRAM (18-28 %).', '']

gp = os.path.join(G, 'r01b_bench_gemm.jsonl')
if os.path.exists(gp):
    shutil.copy(gp, os.path.join(P, 'r01b_gemm_microbench.jsonl'))
if os.path.exists(os.path.join(P, 'r01b_gemm_microbench.jsonl')):
    md += ['## GEMM micro-benchmark (`tools/bench_gemm.py` -> `r01b_gemm_microbench.jsonl`): this repo (fp32 SIMT / 3xTF32 / TF32) vs cuBLAS', '', '```'] + \
          [l.strip() for l in open(os.path.join(P, 'r01b_gemm_microbench.jsonl')) if l.startswith('{')] + ['```', '']
md += ['## instruction-rate micro-benchmarks (`tools/micro/`)', '',
       '* `ffma2_rate.cu`: scalar `FFMA` 39.3 TFLOP/s vs packed `FFMA2` (`fma.rn.f32x2`) 65.7 TFLOP/s on one B200 — a 3-register scalar '
       'FFMA issues every other cycle per scheduler; the attention kernels use the packed form.',
       '* `mma_sync_rate.cu`: legacy warp-level `mma.sync.m16n8k8` TF32 peaks at 278 TFLOP/s (vs ~1 100 for `tcgen05.mma kind::tf32`): with the '
       '3x split needed for fp32 parity that is 93 TFLOP/s fp32-equivalent at best, not enough over FFMA2 to justify re-tiling the 30x30x48 '
       'per-head attention products onto it.', '']
open(os.path.join(P, 'README.md'), 'w').write('\n'.join(md))
print('wrote', os.path.join(P, 'README.md'))
