"""bench.py's reference arm (`--impl reference`: the unmodified reference from oracle/_ref when the recipe has run, else the
oracle port, on the host cores) — the one leg of the bench contract that runs without a GPU: one JSON line with the driver's
keys (headline + `sub` lines for NRMS training and evaluation), and silent exit on non-zero ranks."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env_extra):
    env = dict(os.environ, **env_extra)
    return subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup', '3',
                           '--ref-batch', '16', '--ref-eval-sample', '3'], capture_output=True, text=True, env=env, timeout=600)


def test_reference_arm_prints_one_contract_line():
    r = _run({})
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith('{')]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['metric'] == 'train impressions/s' and d['unit'] == 'impressions/s'
    assert d['higher_is_better'] is True and d['value'] > 0 and d['steps'] == 1 and d['warmup'] == 3
    want_kind = 'reference' if os.path.isdir(os.path.join(ROOT, 'oracle', '_ref', 'xnrs')) else 'port'
    assert d['cpu_baseline']['kind'] == want_kind and d['cpu_baseline']['cores'] >= 1 and d['cpu_baseline']['value'] == d['value']
    assert d['e2e'] == {'value': d['value'], 'unit': d['unit'], 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    assert 'workload' in d['config']
    for k, metric in (('nrms_train', 'train impressions/s'), ('eval', 'eval scored impressions/s')):
        sub = d['sub'][k]
        if 'unavailable' in sub:
            assert want_kind == 'port'              # the NRMS CPU step exists only as the reference itself
            continue
        assert sub['impl'] == 'reference' and sub['metric'] == metric and sub['value'] > 0
        assert sub['cpu_baseline']['kind'] == want_kind


def test_reference_arm_is_silent_on_other_ranks():
    r = _run({'RANK': '1', 'WORLD_SIZE': '2', 'LOCAL_RANK': '1'})
    assert r.returncode == 0 and r.stdout.strip() == ''
