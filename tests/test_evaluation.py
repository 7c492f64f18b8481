"""Row E: full-catalogue evaluation (pre-encode every article once, CSR impressions, fused score + rank metrics)
against the reference procedure restated by the oracle: one impression at a time, full re-encode, numpy metrics
(training.py:194-243).  'emulated' checks the host logic on CPU; 'cuda' runs the real kernels."""
import numpy as np
import pytest
import torch

import _kernel_emulator as EMU
from _common import assert_close, fixture_cfg, load_npz, sub
from oracle import xnrs_oracle as O
from xnrs_b200 import kernels as K
from xnrs_b200 import synthetic as syn
from xnrs_b200.data import TitleStore
from xnrs_b200.evaluation import METRIC_NAMES, CatalogueEvaluator, balanced_impression_shards
from xnrs_b200.models import make_model


@pytest.fixture(params=['emulated', pytest.param('cuda', marks=pytest.mark.gpu)])
def device(request, monkeypatch):
    if request.param == 'emulated':
        monkeypatch.setattr(K, 'call', EMU.call)
        return 'cpu'
    return 'cuda'


def oracle_scores(name, P, cfg, batch):
    if name in ('cl', 'nrms'):
        return O.parent_forward(P, batch, cfg['n_heads'] if name == 'nrms' else 0)
    if name == 'naml':
        return O.naml_forward(P, batch)
    if name.startswith('lstur'):
        return O.lstur_forward(P, batch, cfg['long_short_term_method'], cfg['st_hist_len'])
    return O.npa_forward(P, batch)


@pytest.mark.parametrize('name', ['cl', 'nrms', 'naml', 'lstur_con', 'npa'])
def test_catalogue_eval_matches_per_impression_reference(name, device):
    fx = load_npz('model_' + name)
    cfg = dict(fixture_cfg(fx), device=device)
    model = make_model(cfg)
    model.load_state_dict({k: torch.tensor(v) for k, v in sub(fx, 'sd').items()})
    model.to(device).eval()
    P = O.as_params(sub(fx, 'sd'))
    n_news, S, H, D = 60, cfg['seq_len'], cfg['hist_len'], cfg['d_backbone']
    cat = syn.make_catalogue(n_news, S, vocab=200, dim=D, seed=5, with_abstract=(name == 'naml'),
                             n_categories=cfg['n_categories'], n_subcategories=cfg['n_subcategories'])
    imp = syn.make_eval_impressions(n_news, 24, H, n_users=cfg['n_users'], seed=6)
    store = TitleStore(cat.token_table.to(device), cat.title_tokens.to(device))
    astore = TitleStore(store.token_table, cat.abstract_tokens.to(device)) if name == 'naml' else None
    ev = CatalogueEvaluator(model, store, cat.category, cat.subcategory, astore, news_chunk=17, impression_chunk=7)
    out = ev.evaluate(imp, return_per_impression=True)

    got_scores = out['scores'].cpu()
    got = out['per_impression'].cpu().numpy()
    want_means = []
    for i in range(24):
        a, b = int(imp['offsets'][i]), int(imp['offsets'][i + 1])
        raw = {'hist_ids': imp['hist_ids'][i:i + 1], 'cand_ids': imp['cand_ids'][a:b][None, :],
               'targets': imp['targets'][a:b][None, :, None], 'user_index': imp['user_index'][i:i + 1],
               'main_theme': torch.zeros(1, dtype=torch.int32)}
        raw_scores = oracle_scores(name, P, cfg, syn.dense_batch(cat, raw, with_abstract=(name == 'naml'))).reshape(-1)
        # (1) scores: pre-encoded catalogue path == per-impression full re-encode, 1e-4 of the score scale
        assert_close(got_scores[a:b], torch.relu(raw_scores), 1e-4, f'scores of impression {i}',
                     atol=1e-4 * float(raw_scores.abs().max()))
        # (2) metrics: exact given the scores (ties: descending score, then descending index)
        r = O.impression_metrics(imp['targets'][a:b].numpy(), got_scores[a:b].numpy())
        np.testing.assert_allclose(got[i], [r[k] for k in METRIC_NAMES], rtol=0, atol=1e-9)
        r0 = O.impression_metrics(imp['targets'][a:b].numpy(), torch.relu(raw_scores).numpy())
        want_means.append([r0[k] for k in METRIC_NAMES])
    # (3) epoch means: rank metrics are discontinuous where relu'd scores tie at 0, so only a loose check here
    np.testing.assert_allclose(got.mean(0), np.array(want_means).mean(0), atol=5e-2)
    for i, k in enumerate(METRIC_NAMES):
        assert abs(out[k] - got[:, i].mean()) < 1e-12
    assert out['impressions'] == 24


@pytest.mark.parametrize('name', ['cl', 'naml'])
def test_per_article_pooling_logits_equal_per_slot_pooling(name, device):
    """the evaluator computes the user pooler's logit once per catalogue article (item_logits); re-running the pooler
    on every (user, slot) like the reference must give the same scores and metrics"""
    fx = load_npz('model_' + name)
    cfg = dict(fixture_cfg(fx), device=device)
    model = make_model(cfg)
    model.load_state_dict({k: torch.tensor(v) for k, v in sub(fx, 'sd').items()})
    model.to(device).eval()
    cat = syn.make_catalogue(50, cfg['seq_len'], vocab=200, dim=cfg['d_backbone'], seed=8, with_abstract=(name == 'naml'),
                             n_categories=cfg['n_categories'], n_subcategories=cfg['n_subcategories'])
    imp = syn.make_eval_impressions(50, 40, cfg['hist_len'], n_users=cfg['n_users'], seed=9)
    store = TitleStore(cat.token_table.to(device), cat.title_tokens.to(device))
    astore = TitleStore(store.token_table, cat.abstract_tokens.to(device)) if name == 'naml' else None
    outs = []
    for item_logits in (True, False):
        ev = CatalogueEvaluator(model, store, cat.category, cat.subcategory, astore, news_chunk=23, impression_chunk=9)
        ev.item_logits = item_logits
        outs.append(ev.evaluate(imp, return_per_impression=True))
        assert (ev.news_logit is not None) == item_logits
    assert_close(outs[0]['scores'], outs[1]['scores'], 1e-5, 'scores')
    # rank metrics may only differ where the two score sets order a near-tie differently
    same = (outs[0]['per_impression'] - outs[1]['per_impression']).abs().amax(1) < 1e-9
    assert float(same.float().mean()) >= 0.9


def test_pad_slot_vector_is_the_encoding_of_the_pad_article(device):
    """Appendix A.16: padded history slots must see news_encoder(zeros, zero mask), non-zero for biased heads (NRMS)"""
    fx = load_npz('model_nrms')
    cfg = dict(fixture_cfg(fx), device=device)
    model = make_model(cfg)
    model.load_state_dict({k: torch.tensor(v) for k, v in sub(fx, 'sd').items()})
    model.to(device).eval()
    cat = syn.make_catalogue(20, cfg['seq_len'], vocab=50, dim=cfg['d_backbone'], seed=1)
    ev = CatalogueEvaluator(model, TitleStore(cat.token_table.to(device), cat.title_tokens.to(device)))
    vecs = ev.encode_catalogue()
    P = O.as_params(sub(fx, 'sd'))
    S, D = cfg['seq_len'], cfg['d_backbone']
    want, _ = O.text_encoder(torch.zeros(1, 1, S, D), torch.zeros(1, 1, S, 1), P, 'news_encoder', cfg['n_heads'])
    np.testing.assert_allclose(vecs[0].cpu().numpy(), want.reshape(-1).numpy(), atol=1e-5)
    assert float(vecs[0].abs().max()) > 1e-3 and float(ev.news_mask[0]) == 0.0


def test_impression_shards_balance_candidates():
    sizes = torch.randint(5, 74, (1000,), generator=torch.Generator().manual_seed(0))
    off = torch.cat([torch.zeros(1, dtype=torch.long), sizes.cumsum(0)])
    for w in (1, 2, 4, 8):
        cuts = balanced_impression_shards(off, w)
        assert cuts[0][0] == 0 and cuts[-1][1] == 1000
        assert all(cuts[i][1] == cuts[i + 1][0] for i in range(w - 1))
        loads = [int(off[b] - off[a]) for a, b in cuts]
        assert max(loads) - min(loads) <= 2 * 74


def test_binary_metrics_match_reference(device):
    """acc / rec / prec / confusion of _test_step (training.py:219-222) vs the reference functions (sklearn on
    round(clip(score, 0, 1))): golden generated by make_golden.py:binary_metric_fixture"""
    z = load_npz('binary_metrics')
    got = K.binary_metrics(torch.tensor(z['scores'], device=device), torch.tensor(z['targets'], device=device),
                           torch.tensor(z['offsets'], device=device)).cpu().numpy()
    np.testing.assert_allclose(got[:, :3], z['values'], rtol=0, atol=1e-12)
    np.testing.assert_array_equal(got[:, 3:].astype(np.int64), z['conf'])


def test_evaluator_reports_binary_metrics(device):
    fx = load_npz('model_cl')
    cfg = dict(fixture_cfg(fx), device=device)
    model = make_model(cfg)
    model.load_state_dict({k: torch.tensor(v) for k, v in sub(fx, 'sd').items()})
    model.to(device).eval()
    cat = syn.make_catalogue(50, cfg['seq_len'], vocab=200, dim=cfg['d_backbone'], seed=8)
    imp = syn.make_eval_impressions(50, 30, cfg['hist_len'], n_users=cfg['n_users'], seed=9)
    ev = CatalogueEvaluator(model, TitleStore(cat.token_table.to(device), cat.title_tokens.to(device)), impression_chunk=7)
    ev.binary_metrics = True
    out = ev.evaluate(imp, return_per_impression=True)
    want = K.binary_metrics(out['scores'], imp['targets'].to(device), imp['offsets'].to(device)).cpu().numpy()
    assert abs(out['acc'] - want[:, 0].mean()) < 1e-12 and abs(out['rec'] - want[:, 1].mean()) < 1e-12
    assert out['conf'] == [[int(want[:, 3].sum()), int(want[:, 4].sum())], [int(want[:, 5].sum()), int(want[:, 6].sum())]]
    assert out['candidates'] == imp['targets'].numel()
