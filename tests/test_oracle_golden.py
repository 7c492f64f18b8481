"""Pins the CPU oracle (oracle/xnrs_oracle.py) to outputs of the unmodified reference (tests/golden)."""
import numpy as np
import pytest
import torch

from oracle import xnrs_oracle as O
from _common import MODEL_NAMES, assert_close, fixture_batch, fixture_cfg, load_npz, sub

TOL = 2e-5      # fp32 CPU oracle vs fp32 CPU reference: only summation-order noise is allowed


def oracle_model(name, P, batch, cfg):
    """(scores, user_emb or None) through the oracle for the fixture's model."""
    if name in ('cl', 'nrms'):
        nh = cfg['n_heads'] if name == 'nrms' else 0
        return O.parent_forward(P, batch, nh), O.parent_user_embeddings(P, batch, nh)
    if name == 'naml':
        return O.naml_forward(P, batch), O.naml_user_embeddings(P, batch)
    if name.startswith('lstur'):
        kw = dict(method=cfg['long_short_term_method'], st_hist_len=cfg['st_hist_len'])
        return O.lstur_forward(P, batch, **kw), O.lstur_user_embeddings(P, batch, **kw)
    if name == 'npa':
        return O.npa_forward(P, batch), None
    raise KeyError(name)


@pytest.mark.parametrize('name', MODEL_NAMES)
def test_model_forward_loss_and_grads(name):
    fx = load_npz('model_' + name)
    cfg = fixture_cfg(fx)
    P = O.as_params(sub(fx, 'sd'))
    for v in P.values():
        if v.is_floating_point():
            v.requires_grad_(True)
    batch = fixture_batch(fx)
    scores, ue = oracle_model(name, P, batch, cfg)
    assert_close(scores, fx['ref/scores'], TOL, 'scores')
    l_mse, _ = O.mse_relu_loss(scores, batch['targets'])
    assert_close(l_mse, fx['ref/loss_mse'], TOL, 'mse')
    assert_close(O.bce_logits_loss(scores, batch['targets']), fx['ref/loss_bce'], TOL, 'bce')
    if ue is not None:
        assert_close(ue, fx['ref/user_emb'], TOL, 'user_emb')
        labels = O.theme_labels(batch['main_theme'])
        total, _, l_cl = O.contrastive_train_loss(scores, batch['targets'], ue, labels,
                                                  cfg['contrastive_temperature'], cfg['contrastive_lambda'])
        assert_close(l_cl, fx['ref/loss_cl'], TOL, 'cl')
    else:
        total = l_mse
    assert_close(total, fx['ref/loss_total'], TOL, 'total')
    total.backward()
    gscale = max(float(np.abs(g).max()) for g in sub(fx, 'grad').values())
    for k, g in sub(fx, 'grad').items():
        got = P[k].grad if P[k].grad is not None else torch.zeros_like(P[k])
        if np.abs(g).max() == 0:
            assert float(got.abs().max()) == 0.0, k
        else:
            assert_close(got, g, 1e-4, 'grad ' + k, atol=1e-6 * gscale)


def test_layers():
    fx = load_npz('layers')
    x, m = torch.tensor(fx['x']), torch.tensor(fx['m'])
    P = O.as_params({k: v for k, v in fx.items() if '/' in k and (k.endswith('weight') or k.endswith('bias')
                                                                   or '_l0' in k)})
    P = {k.replace('/', '.', 1): v for k, v in P.items()}
    o, a = O.additive_attention(x, m, P, 'aa', return_weights=True)
    assert_close(o, fx['aa/out'], TOL, 'aa out')
    assert_close(a, fx['aa/a'], TOL, 'aa weights')
    assert float(o[3].abs().max()) == 0.0                       # fully masked row -> exactly zero
    assert_close(O.additive_attention(x, None, P, 'aa'), fx['aa/out_nomask'], TOL, 'aa nomask')
    assert_close(O.multi_head_attention(x, m, P, 'mha', 4), fx['mha/out'], TOL, 'mha')
    xq = x.clone()
    xq[1, 5] += 1.0
    got = O.multi_head_attention(xq, m, P, 'mha', 4)
    assert_close(got, fx['mha/out_perturb_padkey'], TOL, 'mha padded-key perturbation')
    # query-axis masking: a padded *key* still influences valid query rows
    assert float((got[1, :3] - torch.tensor(fx['mha/out'])[1, :3]).abs().max()) > 1e-3
    assert_close(O.personalized_attention(torch.tensor(fx['pa/q']), x, m, P, 'pa'), fx['pa/out'], TOL, 'pa')
    assert_close(O.masked_mean(x, m), fx['mm/out'], TOL, 'masked mean')
    lens = torch.tensor(fx['gru/lens'])
    assert_close(O.gru_last_hidden(x, lens, P, 'gru'), fx['gru/h_zero'], TOL, 'gru h0=0')
    assert_close(O.gru_last_hidden(x, lens, P, 'gru', torch.tensor(fx['gru/h0'])), fx['gru/h_init'], TOL, 'gru h0')


def test_losses_known_answers():
    fx = load_npz('loss_metrics')
    for tag in ('cl', 'cl2'):
        got = O.contrastive_loss(torch.tensor(fx[tag + '/e']), torch.tensor(fx[tag + '/labels']), 0.08)
        assert_close(got, fx[tag + '/loss'], TOL, tag)
    assert abs(float(fx['cl/loss']) - 6.636086463928223) < 1e-5            # SURVEY §8(c) KAT-cl
    p, n = torch.tensor(fx['nll/p']), torch.tensor(fx['nll/n'])
    assert_close(O.ranking_nll(p, n), fx['nll/mean'], TOL, 'nll mean')
    assert_close(O.ranking_nll(p, n, 'none'), fx['nll/none'], TOL, 'nll none')
    assert abs(float(fx['nll/mean']) - 1.5622555017471313) < 1e-6          # SURVEY §8(c) KAT-nll


@pytest.mark.parametrize('tag', ['met', 'mett'])
def test_metrics_vs_reference(tag):
    fx = load_npz('loss_metrics')
    off = fx[tag + '/offsets']
    for i in range(len(off) - 1):
        s, t = fx[tag + '/scores'][off[i]:off[i + 1]], fx[tag + '/targets'][off[i]:off[i + 1]]
        r = O.impression_metrics(t, s)
        got = [r['auc'], r['rr'], r['ndcg@5'], r['ndcg@10'], r['ctr@1'], r['ctr@10']]
        np.testing.assert_allclose(got, fx[tag + '/values'][i], rtol=0, atol=1e-7)  # ref means float32 targets


def test_metrics_kat():
    fx = load_npz('loss_metrics')
    r = O.impression_metrics(fx['kat/t'], fx['kat/s'])
    want = [0.7037037037037037, 1.0, 0.46927872602275644, 0.7928654229965442, 1.0, 0.3]
    np.testing.assert_allclose([r['auc'], r['rr'], r['ndcg@5'], r['ndcg@10'], r['ctr@1'], r['ctr@10']], want,
                               atol=1e-12)
    np.testing.assert_allclose(fx['kat/values'], want, atol=1e-12)
    assert np.isnan(O.auc_score([1, 1], [0.3, 0.2]))                       # single-class impression


def test_gather_semantics():
    g = torch.Generator().manual_seed(0)
    table = torch.randn(50, 8, generator=g)
    table[0] = 0
    titles = torch.randint(1, 50, (10, 6), generator=g)
    titles[0] = 0                                                          # pad article
    titles[3, 4:] = 0
    x, m = O.gather_titles(table, titles, torch.tensor([[3, 0], [9, 3]]))
    assert x.shape == (2, 2, 6, 8) and m.shape == (2, 2, 6, 1)
    assert torch.equal(x[0, 0, 1], table[titles[3, 1]])
    assert float(x[0, 1].abs().max()) == 0 and float(m[0, 1].sum()) == 0
    assert m[0, 0, :, 0].tolist() == [1, 1, 1, 1, 0, 0]


def test_adam_matches_torch():
    p = torch.randn(17, 5)
    ref = torch.nn.Parameter(p.clone())
    opt = torch.optim.Adam([ref], lr=1e-3)
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    for step in range(1, 4):
        g = torch.randn(17, 5)
        ref.grad = g.clone()
        opt.step()
        O.adam_step(p, g, m, v, step, 1e-3)
    assert_close(p, ref, 1e-6, 'adam')
