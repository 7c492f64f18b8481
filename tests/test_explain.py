"""§8(f) row 3: the explainer's integrated-gradients loop (xnrs/explain.py:144-172) through the xnrs_b200 modules.

It is the only caller of the module API that needs d score / d INPUT token embeddings (training never does: the
inputs are frozen), i.e. the dX of the first-layer GEMMs and of the pooling / attention kernels.  The golden
fixture tests/golden/explain.npz holds the reference's per-step gradients and attributions for the CL and NRMS
fixture models (make_golden.py:explain_fixture); 'emulated' checks the autograd glue on CPU, 'cuda' the kernels."""
import numpy as np
import pytest
import torch

import _kernel_emulator as EMU
from _common import assert_close, fixture_cfg, load_npz, sub
from xnrs_b200 import kernels as K
from xnrs_b200.models import make_model


@pytest.fixture(params=['emulated', pytest.param('cuda', marks=pytest.mark.gpu),
                        pytest.param('cuda-tf32x3', marks=pytest.mark.gpu)])
def device(request, monkeypatch):
    if request.param == 'emulated':
        monkeypatch.setattr(K, 'call', EMU.call)
        return 'cpu'
    monkeypatch.setattr(K, '_precision', K.PRECISIONS['tf32x3' if request.param.endswith('tf32x3') else 'fp32'])
    return 'cuda'


@pytest.mark.parametrize('name', ['cl', 'nrms'])
def test_integrated_gradients_match_reference(name, device):
    fx, gx = load_npz('model_' + name), load_npz('explain')
    model = make_model(dict(fixture_cfg(fx), device=device))
    model.load_state_dict({k: torch.tensor(v) for k, v in sub(fx, 'sd').items()})
    model.to(device).eval()
    b, cidx = (int(v) for v in gx[f'{name}/pick'])
    t = lambda key: torch.tensor(fx[key], device=device)
    hx = t('batch/user_features/history/title_emb/x')[b:b + 1]
    hm = t('batch/user_features/history/title_emb/m')[b:b + 1]
    cx = t('batch/candidate_features/title_emb/x')[b:b + 1, cidx:cidx + 1]
    cm = t('batch/candidate_features/title_emb/m')[b:b + 1, cidx:cidx + 1]

    hist_emb = hx.clone().requires_grad_()
    c, _ = model.news_encoder((cx.clone().requires_grad_(), cm))
    n_steps = 8
    da = 1 / n_steps
    grads = []
    for a in torch.arange(da, 1 + da, da):                       # explain.py:160-166
        ga = float(a) * hist_emb
        ha, ham = model.news_encoder((ga, hm))
        ua = model.user_encoder.forward(inpt=(ha, ham))
        sa = torch.relu(model.rec_model(ua, c))
        grads.append(torch.autograd.grad(sa, ga)[0])
    grads = torch.cat(grads)
    attr = torch.sum(torch.sum(grads * da, dim=0) * hist_emb.detach(), dim=(0, 3))

    assert_close(sa.reshape(()), gx[f'{name}/s_true'], 1e-4, 'score at a=1')
    assert_close(grads, gx[f'{name}/grads'], 1e-4, 'd score / d scaled history embeddings, all steps')
    assert_close(attr, gx[f'{name}/attr'], 1e-4, 'token attributions')
    # padded tokens / padded history slots receive exactly zero attribution
    pad = (hm[0, :, :, 0] == 0).cpu()
    assert float(attr.detach().cpu()[pad].abs().max() if pad.any() else 0.0) == 0.0
