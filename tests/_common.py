"""Shared helpers for the test-suite: golden-fixture loading and tolerance checks."""
import json
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
MODEL_NAMES = ['cl', 'nrms', 'naml', 'lstur_con', 'lstur_ini', 'npa']
# ablation models, non-dot scorers, constructor options off their defaults (tests/golden/make_golden.py)
EXTRA_MODEL_NAMES = ['base', 'mean', 'param_free', 'nrms_lf', 'small_naml', 'cl_bilin', 'cl_fc', 'cl_norm', 'nrms_unscaled']


def load_npz(name):
    with np.load(os.path.join(GOLDEN, name + '.npz'), allow_pickle=False) as z:
        return {k: z[k] for k in z.files}


def sub(d, prefix):
    """entries of a flat fixture dict below `prefix/`, prefix stripped."""
    p = prefix + '/'
    return {k[len(p):]: v for k, v in d.items() if k.startswith(p)}


def fixture_batch(fx, device='cpu', dtype=torch.float32):
    """rebuild the reference-format batch dict from a model fixture."""
    def t(a):
        a = torch.as_tensor(a)
        return a.to(device=device, dtype=dtype) if a.is_floating_point() else a.to(device)

    def side(prefix):
        out = {}
        for k, v in sub(fx, prefix).items():
            if k.endswith('/x'):
                out[k[:-2]] = (t(v), t(fx[f'{prefix}/{k[:-2]}/m']))
            elif not k.endswith('/m'):
                out[k] = t(v)
        return out

    hist = side('batch/user_features/history')
    other = side('batch/user_features/other')
    cand = side('batch/candidate_features')
    return {'user_features': {'history': hist, 'other': other}, 'candidate_features': cand,
            'targets': t(fx['batch/targets']), 'main_theme': [str(s) for s in fx['batch/main_theme']]}


def fixture_cfg(fx):
    return json.loads(str(fx['cfg']))


def rel_err(a, b):
    """max |a-b| relative to the scale of the reference tensor b (the north-star's 'relative')."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    scale = max(float(np.max(np.abs(b))), 1e-30)
    return float(np.max(np.abs(a - b))) / scale if a.size else 0.0


def assert_close(a, b, tol=1e-4, what='', atol=0.0):
    """max|a-b| <= tol * max|b| + atol.  `atol` is only for quantities that are analytically ~0
    (e.g. d loss / d fc2.bias of the additive pooler: the softmax is shift-invariant up to its 1e-8)."""
    if isinstance(a, torch.Tensor):
        a = a.detach().float().cpu().numpy()
    if isinstance(b, torch.Tensor):
        b = b.detach().float().cpu().numpy()
    assert np.shape(a) == np.shape(b), f'{what}: shape {np.shape(a)} vs {np.shape(b)}'
    assert np.all(np.isfinite(a)), f'{what}: non-finite values'
    e = rel_err(a, b)
    if os.environ.get('XNRS_SHOW_ERR'):
        print(f'[err] {what}: {e:.3e} (tol {tol:.1e})')
    if atol and a.size and float(np.max(np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64)))) <= atol:
        return
    assert e <= tol, f'{what}: relative error {e:.3e} > {tol:.1e}'
