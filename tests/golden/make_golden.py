"""Generate tests/golden/*.npz by running the UNMODIFIED reference (/root/reference) on CPU.

Run in the build container only:   python tests/golden/make_golden.py
The fixtures are committed; the GPU box never needs /root/reference.

For each of the five target models (CL/StandardRec, NRMS, NAML, LSTUR[con|ini], NPA) at a small
shape the script stores: config, reference-initialised state_dict, a dense reference-format batch
(with padded history slots, ragged title lengths and one fully padded history), and the reference's
outputs: scores, user embeddings, trainer losses (MSE∘ReLU, BCE-with-logits, InfoNCE, total) and the
gradient of the total loss w.r.t. every parameter.  Layer-level fixtures (additive / personalised /
multi-head attention, GRU) and metric / loss known-answer values are stored too.
"""
import json
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from _refshim import DotMap, load_reference  # noqa: E402

load_reference()
from xnrs.models import make_model  # noqa: E402
from xnrs.models.components import layers  # noqa: E402
import xnrs.training as T  # noqa: E402
import xnrs.utils as U  # noqa: E402
from xnrs.evaluation import metrics as M  # noqa: E402

torch.set_num_threads(4)

BASE = dict(scoring='dot', text_features=['title_emb'], catg_features=[], user_features=[], add_features=[],
            title_emb_dim=16, total_emb_dim=16, d_backbone=32, n_heads=4, hist_len=4, st_hist_len=4, seq_len=5,
            p_dropout=0., bias=False, n_categories=6, n_subcategories=9, n_users=11, cat_emb_dim=8, sub_emb_dim=8,
            user_emb_dim=8, n_negatives=2, batch_size=4, contrastive_temperature=0.08, contrastive_lambda=0.1,
            device='cpu')

MODELS = {
    'cl': dict(model='standard'),
    'nrms': dict(model='NRMS', bias=False),
    'naml': dict(model='NAML', text_features=['title_emb', 'abstract_emb'],
                 catg_features=['category_index', 'subcategory_index']),
    'lstur_con': dict(model='LSTUR', total_emb_dim=24, long_term_method='embedding', long_short_term_method='con',
                      p_user_dropout=0.0, catg_features=['category_index'], user_features=['user_index']),
    'lstur_ini': dict(model='LSTUR', total_emb_dim=24, long_term_method='embedding', long_short_term_method='ini',
                      p_user_dropout=0.0, catg_features=['category_index'], user_features=['user_index']),
    'npa': dict(model='NPA', user_features=['user_index']),
}
# ablation models, non-dot scorers and the constructor options the five target models leave at their
# defaults.  Keys starting with '_' are instructions for this script (and for tests/test_models.py:build), not config keys:
#   _class      build this class of xnrs.models.full_models directly (not reachable through the reference's make_model)
#   _normalize  set rec_model.normalize = True after construction (scoring.py:20-22)
#   _unscaled   set scaled = False on every MultiHeadAttention (layers.py:135-137)
EXTRA_MODELS = {
    'base': dict(model='base'),
    'mean': dict(model='mean', bias=True),
    'param_free': dict(model='param_free', title_emb_dim=32, total_emb_dim=32, _class='ParamFreeRec'),
    'nrms_lf': dict(model='NRMS_LF', _class='NRMS_LF'),
    'small_naml': dict(model='smallNAML', catg_features=['category_index']),
    'cl_bilin': dict(model='standard', scoring='bilin', bias=True),
    'cl_fc': dict(model='standard', scoring='fc', bias=True),
    'cl_norm': dict(model='standard', bias=True, _normalize=True),   # biased heads: no all-zero vector (0 / ||0|| is NaN in the reference)
    'nrms_unscaled': dict(model='NRMS', _unscaled=True),
}


def make_batch(cfg, g, lstur=False):
    B, H, N, S, D = cfg.batch_size, cfg.hist_len, cfg.n_negatives + 1, cfg.seq_len, cfg.d_backbone

    def text(n, hist):
        x = torch.randn(B, n, S, D, generator=g)
        ln = torch.randint(1, S + 1, (B, n), generator=g)
        if hist:
            nh = torch.tensor([H, 2, 1, 3][:B])         # valid history items, front aligned
            if not lstur:
                nh[2] = 0                               # one user with a fully padded history
            ln = ln * (torch.arange(n)[None, :] < nh[:, None])
        m = (torch.arange(S)[None, None, :] < ln[:, :, None]).float().unsqueeze(-1)
        return x * m, m                                 # padded tokens are zero rows (dataset.py:82-85)

    hist, cand = {}, {}
    for feat in cfg.text_features:
        hist[feat], cand[feat] = text(H, True), text(N, False)
    hm = hist['title_emb'][1].sum(2).clamp(0, 1).squeeze(-1)
    for feat, hi in (('category_index', cfg.n_categories), ('subcategory_index', cfg.n_subcategories)):
        if feat in cfg.catg_features:
            hist[feat] = (torch.randint(1, hi + 1, (B, H), generator=g) * hm).int()     # pad label 0
            cand[feat] = torch.randint(1, hi + 1, (B, N), generator=g).int()
    other = {}
    if 'user_index' in cfg.user_features:
        other['user_index'] = torch.randint(1, cfg.n_users + 1, (B, 1), generator=g).int()
    t = torch.zeros(B, N, 1)
    t[:, 0] = 1
    return {'user_features': {'history': hist, 'other': other}, 'candidate_features': cand, 'targets': t,
            'main_theme': ['sport', 'news', 'sport', 'life'][:B]}


def flat(prefix, obj, out):
    if isinstance(obj, dict):
        for k, v in obj.items():
            flat(f'{prefix}/{k}', v, out)
    elif isinstance(obj, (tuple, list)) and len(obj) == 2 and isinstance(obj[0], torch.Tensor):
        out[prefix + '/x'], out[prefix + '/m'] = obj[0].numpy(), obj[1].numpy()
    elif isinstance(obj, torch.Tensor):
        out[prefix] = obj.numpy()


def model_fixture(name, over, seed):
    cfg = DotMap(dict(BASE, **{k: v for k, v in over.items() if not k.startswith('_')}))
    torch.manual_seed(seed)
    if '_class' in over:
        import xnrs.models.full_models.nrms as _nrms
        import xnrs.models.full_models.param_free_model as _pf
        from xnrs.models.components import scoring as _scoring
        model = {'ParamFreeRec': _pf.ParamFreeRec, 'NRMS_LF': _nrms.NRMS_LF}[over['_class']](cfg, _scoring.DotScoring())
    else:
        model = make_model(cfg)
    if over.get('_normalize'):
        model.rec_model.normalize = True
    if over.get('_unscaled'):
        for m_ in model.modules():
            if isinstance(m_, layers.MultiHeadAttention):
                m_.scaled = False
    # default init leaves dummy_param at 0 and heads tiny; perturb so every path carries signal
    with torch.no_grad():
        for p in model.parameters():
            if p.numel() > 1:
                p.mul_(2.0)
    model.eval()
    g = torch.Generator().manual_seed(seed + 100)
    batch = make_batch(cfg, g, lstur=name.startswith('lstur'))
    out = {'cfg': np.array(json.dumps(dict(dict(cfg), **{k: v for k, v in over.items() if k.startswith('_')})))}
    for k, v in model.state_dict().items():
        out['sd/' + k] = v.numpy().copy()
    flat('batch', {k: v for k, v in batch.items() if k != 'main_theme'}, out)
    out['batch/main_theme'] = np.array(batch['main_theme'])

    scores = model(batch)
    out['ref/scores'] = scores.detach().numpy()
    fake = types.SimpleNamespace(temperature=cfg.contrastive_temperature)
    labels = torch.tensor([{'sport': 0, 'news': 1, 'life': 2}[t] for t in batch['main_theme']])
    l_rec = torch.nn.functional.mse_loss(torch.relu(scores), batch['targets'])
    out['ref/loss_mse'] = l_rec.detach().numpy()
    out['ref/loss_bce'] = torch.nn.functional.binary_cross_entropy_with_logits(scores, batch['targets']).detach().numpy()
    # BCERankingTrainer (training.py:324-331): forward = sigmoid(model(batch)), L = nn.BCELoss()
    out['ref/loss_bce_sigmoid'] = torch.nn.BCELoss()(torch.sigmoid(scores), batch['targets']).detach().numpy()
    if hasattr(model, 'get_user_embeddings'):
        ue = model.get_user_embeddings(batch)
        out['ref/user_emb'] = ue.detach().numpy()
        l_cl = T.ContrastiveRankingTrainer._compute_contrastive_loss(fake, ue, labels)
        total = l_rec + cfg.contrastive_lambda * l_cl
        out['ref/loss_cl'] = l_cl.detach().numpy()
    else:                                   # NPA: no CL hook exists -> MSE trainer
        total = l_rec
    out['ref/loss_total'] = total.detach().numpy()
    model.zero_grad()
    if total.requires_grad:                 # ParamFreeRec: nothing trainable reaches the loss
        total.backward()
    for k, p in model.named_parameters():
        out['grad/' + k] = (p.grad if p.grad is not None else torch.zeros_like(p)).numpy().copy()
    np.savez_compressed(os.path.join(HERE, f'model_{name}.npz'), **out)
    print(name, 'loss', float(total), 'params', sum(p.numel() for p in model.parameters()))


def layer_fixtures():
    g = torch.Generator().manual_seed(7)
    out = {}
    R, L, F = 6, 7, 32
    x = torch.randn(R, L, F, generator=g)
    ln = torch.tensor([7, 3, 1, 0, 5, 7])
    m = (torch.arange(L)[None, :] < ln[:, None]).float().unsqueeze(-1)
    out['x'], out['m'] = x.numpy(), m.numpy()

    torch.manual_seed(11)
    aa = layers.AdditiveAttention(F, 256)
    for k, v in aa.state_dict().items():
        out['aa/' + k] = v.numpy().copy()
    o, a = aa(x, m, return_weights=True)
    out['aa/out'], out['aa/a'] = o.detach().numpy(), a.detach().numpy()
    out['aa/out_nomask'] = aa(x).detach().numpy()

    mha = layers.MultiHeadAttention(4, F).eval()
    for k, v in mha.state_dict().items():
        out['mha/' + k] = v.numpy().copy()
    out['mha/out'] = mha(x, m).detach().numpy()
    xp = x.clone()
    xp[0, 6] += 1.0                                   # perturb a key that is *valid* in row 0
    xq = x.clone()
    xq[1, 5] += 1.0                                   # perturb a PADDED key of row 1: valid rows still change
    out['mha/out_perturb_padkey'] = mha(xq, m).detach().numpy()

    pa = layers.PersonalizedAttention(F, 128, 8)
    q = torch.randn(R, 1, 8, generator=g)
    for k, v in pa.state_dict().items():
        out['pa/' + k] = v.numpy().copy()
    out['pa/q'] = q.numpy()
    out['pa/out'] = pa(q, x, m).detach().numpy()

    out['mm/out'] = layers.MaskedMean()(x, m).detach().numpy()

    gru = torch.nn.GRU(F, 10, batch_first=True)
    for k, v in gru.state_dict().items():
        out['gru/' + k] = v.numpy().copy()
    lens = torch.tensor([7, 3, 1, 2, 5, 7])
    h0 = torch.randn(1, R, 10, generator=g)
    pk = torch.nn.utils.rnn.pack_padded_sequence(x, lens, batch_first=True, enforce_sorted=False)
    out['gru/lens'] = lens.numpy()
    out['gru/h0'] = h0[0].numpy()
    out['gru/h_zero'] = gru(pk)[1][0].detach().numpy()
    out['gru/h_init'] = gru(pk, h0)[1][0].detach().numpy()
    np.savez_compressed(os.path.join(HERE, 'layers.npz'), **out)


def loss_metric_fixtures():
    out = {}
    g = torch.Generator().manual_seed(123)
    e = torch.randn(6, 8, generator=g)
    lab = torch.tensor([0, 1, 0, 2, 1, 0])
    fake = types.SimpleNamespace(temperature=0.08)
    out['cl/e'], out['cl/labels'] = e.numpy(), lab.numpy()
    out['cl/loss'] = T.ContrastiveRankingTrainer._compute_contrastive_loss(fake, e, lab).numpy()
    e2 = torch.randn(40, 16, generator=g)
    lab2 = torch.randint(0, 6, (40,), generator=g)
    lab2[7] = 17                                        # an anchor without positives
    out['cl2/e'], out['cl2/labels'] = e2.numpy(), lab2.numpy()
    out['cl2/loss'] = T.ContrastiveRankingTrainer._compute_contrastive_loss(fake, e2, lab2).numpy()

    p = torch.tensor([[.5], [-.2]])
    n = torch.tensor([[.1, -.3, .2, 0.], [.4, .1, -.5, .3]])
    out['nll/p'], out['nll/n'] = p.numpy(), n.numpy()
    out['nll/mean'] = U.ranking_loss(p, n).numpy()
    out['nll/none'] = U.ranking_loss(p, n, reduction='none').numpy()

    # metrics: tie-free random impressions through the reference functions unchanged
    rng = np.random.default_rng(5)
    ys, yt, vals = [], [], []
    for i in range(64):
        n_pos, n_neg = int(rng.integers(1, 4)), int(rng.integers(1, 70))
        t = np.array([1.] * n_pos + [0.] * n_neg, dtype=np.float32)
        s = rng.permutation(n_pos + n_neg).astype(np.float32) / 7.0          # distinct scores
        ys.append(s)
        yt.append(t)
        vals.append([M.auc_score(t, s), M.rr_score(t, s), M.ndcg_score(t, s, 5), M.ndcg_score(t, s, 10),
                     M.ctr_score(t, s, 1), M.ctr_score(t, s, 10)])
    out['met/offsets'] = np.cumsum([0] + [len(s) for s in ys]).astype(np.int64)
    out['met/scores'], out['met/targets'] = np.concatenate(ys), np.concatenate(yt)
    out['met/values'] = np.array(vals, dtype=np.float64)

    # tied (post-ReLU style) scores: reference functions with argsort forced to kind='stable'
    # (the documented canonical tie policy); AUC is the stock sklearn call.
    real_argsort = np.argsort
    ys, yt, vals = [], [], []
    for i in range(64):
        n_pos, n_neg = int(rng.integers(1, 4)), int(rng.integers(1, 40))
        t = np.array([1.] * n_pos + [0.] * n_neg, dtype=np.float32)
        s = np.maximum(rng.normal(size=n_pos + n_neg).astype(np.float32), 0).round(1)
        M.np.argsort = lambda a, *aa, **kw: real_argsort(a, kind='stable')
        try:
            row = [M.auc_score(t, s), M.rr_score(t, s), M.ndcg_score(t, s, 5), M.ndcg_score(t, s, 10),
                   M.ctr_score(t, s, 1), M.ctr_score(t, s, 10)]
        finally:
            M.np.argsort = real_argsort
        ys.append(s)
        yt.append(t)
        vals.append(row)
    out['mett/offsets'] = np.cumsum([0] + [len(s) for s in ys]).astype(np.int64)
    out['mett/scores'], out['mett/targets'] = np.concatenate(ys), np.concatenate(yt)
    out['mett/values'] = np.array(vals, dtype=np.float64)

    t = np.array([1, 0, 0, 1, 0, 0, 0, 0, 0, 0, 0, 1], dtype=np.float32)
    s = np.array([.9, .8, .1, .4, .3, .05, .7, .2, .6, 0, .55, .35], dtype=np.float32)
    out['kat/t'], out['kat/s'] = t, s
    out['kat/values'] = np.array([M.auc_score(t, s), M.rr_score(t, s), M.ndcg_score(t, s, 5),
                                  M.ndcg_score(t, s, 10), M.ctr_score(t, s, 1), M.ctr_score(t, s, 10)])
    np.savez_compressed(os.path.join(HERE, 'loss_metrics.npz'), **out)


def explain_fixture():
    """§8(f) row 3: the explainer's integrated-gradients loop (xnrs/explain.py:144-172) — scaled history inputs through
    news_encoder / user_encoder / rec_model with autograd.grad w.r.t. the scaled token embeddings — run with the REFERENCE
    modules on the stored CL and NRMS fixture models (the best-scored (impression, candidate) pair of the fixture batch, 8 steps)."""
    out = {}
    for name in ('cl', 'nrms'):
        with np.load(os.path.join(HERE, f'model_{name}.npz'), allow_pickle=False) as z:
            fx = {k: z[k] for k in z.files}
        cfg = DotMap(json.loads(str(fx['cfg'])))
        model = make_model(cfg)
        model.load_state_dict({k[3:]: torch.tensor(v) for k, v in fx.items() if k.startswith('sd/')})
        model.eval()
        sc = fx['ref/scores'][..., 0]
        b, cidx = (int(v) for v in np.unravel_index(np.argmax(sc), sc.shape))      # a pair the ReLU does not zero
        out[f'{name}/pick'] = np.array([b, cidx])
        hx = torch.tensor(fx['batch/user_features/history/title_emb/x'])[b:b + 1]
        hm = torch.tensor(fx['batch/user_features/history/title_emb/m'])[b:b + 1]
        cx = torch.tensor(fx['batch/candidate_features/title_emb/x'])[b:b + 1, cidx:cidx + 1]
        cm = torch.tensor(fx['batch/candidate_features/title_emb/m'])[b:b + 1, cidx:cidx + 1]
        hist_emb, hist_att = hx.clone().requires_grad_(), hm.clone().requires_grad_()
        c, _ = model.news_encoder((cx.clone().requires_grad_(), cm.clone().requires_grad_()))
        n_steps = 8
        da = 1 / n_steps
        grads = []
        for a in torch.arange(da, 1 + da, da):
            ga = a * hist_emb
            ha, ham = model.news_encoder((ga, hist_att))
            ua = model.user_encoder.forward(inpt=(ha, ham))
            sa = torch.relu(model.rec_model(ua, c))
            grads.append(torch.autograd.grad(sa, ga)[0])
        grads = torch.cat(grads)
        attr = torch.sum(torch.sum(grads * da, dim=0) * hist_emb.detach(), dim=(0, 3))
        out[f'{name}/grads'] = grads.detach().numpy()
        out[f'{name}/attr'] = attr.detach().numpy()
        out[f'{name}/s_true'] = np.array(float(sa))
        print('explain', name, 's_true', float(sa), 's_attr', float(attr.sum()))
    np.savez_compressed(os.path.join(HERE, 'explain.npz'), **out)


def binary_metric_fixture():
    """the thresholded epoch metrics of `_test_step` (training.py:219-222; metrics.py:47-64: acc / rec / prec / confusion on
    round(clip(score, 0, 1))) through the reference functions, on random impressions incl. scores of exactly 0.5, scores
    outside [0, 1] and impressions whose predictions are all one class"""
    rng = np.random.default_rng(9)
    ys, yt, vals, confs = [], [], [], []
    import warnings
    for i in range(48):
        n_pos, n_neg = int(rng.integers(1, 4)), int(rng.integers(1, 40))
        t = np.array([1.] * n_pos + [0.] * n_neg, dtype=np.float32)
        s = rng.normal(0.4, 0.5, size=n_pos + n_neg).astype(np.float32)
        if i % 5 == 0:
            s[rng.integers(0, len(s))] = 0.5                    # np.round(0.5) == 0
        if i % 7 == 0:
            s = np.minimum(s, 0.3).astype(np.float32)           # nothing predicted positive
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            vals.append([M.acc_score(t, s), M.recall_score(t, s), M.precision_score(t, s)])
            confs.append(M.confusion_matrix(t, s).reshape(-1))  # [[tn, fp], [fn, tp]] — both classes are always in t here
        ys.append(s)
        yt.append(t)
    out = {'offsets': np.cumsum([0] + [len(s) for s in ys]).astype(np.int64), 'scores': np.concatenate(ys),
           'targets': np.concatenate(yt), 'values': np.array(vals, dtype=np.float64), 'conf': np.array(confs, dtype=np.int64)}
    np.savez_compressed(os.path.join(HERE, 'binary_metrics.npz'), **out)


def dataset_fixture():
    """row G: the reference's NewsRecDataset + custom_collate_fn (xnrs/data/dataset.py:48-163, utils.py:190-204) on a tiny
    synthetic news table / behaviour log, in eval mode (all candidates) and train mode (seeded negative sampling).  The
    raw inputs are stored next to the dense outputs so tests can rebuild the batch through xnrs_b200.mind_io."""
    import random
    from xnrs.data.dataset import NewsRecDataset
    rng = np.random.default_rng(5)
    n_news, S, D, H, K = 12, 4, 6, 3, 2
    ids = [f'N{100 + i}' for i in range(n_news)]
    emb = rng.normal(size=(n_news, S, D)).astype(np.float32)
    lens = rng.integers(1, S + 1, size=n_news)
    mask = (np.arange(S)[None, :] < lens[:, None]).astype(np.int64)
    emb = emb * mask[:, :, None]                      # the reference stores embeddings relative to a pad reference: 0 at pads
    cat = rng.integers(1, 5, size=n_news)
    news_feat = {nid: {'title_emb': (emb[i][None], mask[i][None]), 'category_index': int(cat[i])} for i, nid in enumerate(ids)}
    sessions = []
    for u in range(6):
        nh = int(rng.integers(1, 2 * H + 1))
        hist = [ids[j] for j in rng.integers(0, n_news, size=nh)]
        pos = [ids[j] for j in rng.integers(0, n_news, size=int(rng.integers(1, 3)))]
        neg = [ids[j] for j in rng.integers(0, n_news, size=int(rng.integers(2, 6)))]
        sessions.append({'history': hist, 'positives': pos, 'negatives': neg, 'user_index': str(u + 1),
                         'main_theme': ['sports', 'news', 'finance'][u % 3], 'main_category': 'x'})
    out = {'ids': np.array(ids), 'emb': emb, 'mask': mask, 'cat': cat.astype(np.int64),
           'sessions': np.array(json.dumps(sessions)), 'dims': np.array([S, D, H, K])}
    common = dict(l_seq=S, l_hist=H, text_features=['title_emb'], catg_features=['category_index'],
                  user_features=['user_index'])
    ds = NewsRecDataset([dict(s, history=list(s['history'])) for s in sessions], news_feat, mode='train', n_negatives=K, **common)
    random.seed(77)
    batch = U.custom_collate_fn([ds[i] for i in range(len(sessions))])
    out['train/hist_x'] = batch['user_features']['history']['title_emb'][0].numpy()
    out['train/hist_m'] = batch['user_features']['history']['title_emb'][1].numpy()
    out['train/cand_x'] = batch['candidate_features']['title_emb'][0].numpy()
    out['train/cand_m'] = batch['candidate_features']['title_emb'][1].numpy()
    out['train/hist_cat'] = batch['user_features']['history']['category_index'].numpy()
    out['train/cand_cat'] = batch['candidate_features']['category_index'].numpy()
    out['train/user_index'] = batch['user_features']['other']['user_index'].numpy()
    out['train/targets'] = batch['targets'].numpy()
    out['train/item_ids'] = np.array(json.dumps(batch['item_ids']))
    ds = NewsRecDataset([dict(s, history=list(s['history'])) for s in sessions], news_feat, mode='eval', n_negatives=None, **common)
    for i in range(len(sessions)):
        smp = ds[i]
        out[f'eval/{i}/hist_x'] = smp['user_features']['history']['title_emb'][0].numpy()
        out[f'eval/{i}/hist_m'] = smp['user_features']['history']['title_emb'][1].numpy()
        out[f'eval/{i}/cand_x'] = smp['candidate_features']['title_emb'][0].numpy()
        out[f'eval/{i}/cand_m'] = smp['candidate_features']['title_emb'][1].numpy()
        out[f'eval/{i}/hist_cat'] = smp['user_features']['history']['category_index'].numpy()
        out[f'eval/{i}/targets'] = smp['targets'].numpy()
    np.savez_compressed(os.path.join(HERE, 'dataset.npz'), **out)


if __name__ == '__main__':
    if '--dataset-only' in sys.argv:
        dataset_fixture()
        sys.exit(0)
    if '--binary-only' in sys.argv:
        binary_metric_fixture()
        sys.exit(0)
    if '--explain-only' in sys.argv:
        explain_fixture()
        sys.exit(0)
    for i, (name, over) in enumerate(MODELS.items()):
        model_fixture(name, over, seed=20 + i)
    for i, (name, over) in enumerate(EXTRA_MODELS.items()):
        model_fixture(name, over, seed=60 + i)
    layer_fixtures()
    loss_metric_fixtures()
    dataset_fixture()
    explain_fixture()
    binary_metric_fixture()
    print('golden fixtures written to', HERE)
