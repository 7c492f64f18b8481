"""Model-level parity against the golden fixtures (outputs of the unmodified reference).

Two runs of the same checks:
  * ``emulated`` (CPU, not gpu): the C-ABI calls are replaced by tests/_kernel_emulator.py, so this
    validates the HOST logic — module wiring, state_dict layout, autograd glue, trainer hooks;
  * ``cuda`` (@gpu): the real sm_100a kernels through the real C ABI — the parity test proper.
"""
import numpy as np
import pytest
import torch

import _kernel_emulator as EMU
from _common import EXTRA_MODEL_NAMES, MODEL_NAMES, assert_close, fixture_batch, fixture_cfg, load_npz, sub
from xnrs_b200 import kernels as K
from xnrs_b200.models import make_model
from xnrs_b200.training import (BCELogitsRankingTrainer, BCERankingTrainer, ContrastiveRankingTrainer,
                                MSERankingTrainer)

TOL = 1e-4          # north-star fp32 tolerance (relative to the tensor's scale)


@pytest.fixture(params=['emulated', pytest.param('cuda', marks=pytest.mark.gpu),
                        pytest.param('cuda-tf32x3', marks=pytest.mark.gpu)])
def device(request, monkeypatch):
    """'cuda' = exact-fp32 SIMT GEMMs; 'cuda-tf32x3' = the tcgen05 3xTF32 tensor-core GEMMs (same 1e-4 bar)"""
    if request.param == 'emulated':
        monkeypatch.setattr(K, 'call', EMU.call)
        return 'cpu'
    monkeypatch.setattr(K, '_precision', K.PRECISIONS['tf32x3' if request.param.endswith('tf32x3') else 'fp32'])
    return 'cuda'


def build(name, device):
    fx = load_npz('model_' + name)
    cfg = dict(fixture_cfg(fx), device=device)
    special = {k: cfg.pop(k) for k in list(cfg) if k.startswith('_')}          # instructions of make_golden.py, not config keys
    if '_class' in special:             # classes the reference's factory cannot reach either
        from xnrs_b200.models import zoo
        from xnrs_b200.models.components import DotScoring
        model = getattr(zoo, special['_class'])(cfg, DotScoring())
    else:
        model = make_model(cfg)
    if special.get('_normalize'):
        model.rec_model.normalize = True
    if special.get('_unscaled'):
        from xnrs_b200.models.components import MultiHeadAttention
        for m_ in model.modules():
            if isinstance(m_, MultiHeadAttention):
                m_.scaled = False
    sd = {k: torch.tensor(v) for k, v in sub(fx, 'sd').items()}
    assert set(model.state_dict().keys()) == set(sd.keys()), 'state_dict keys differ from the reference'
    for k, v in model.state_dict().items():
        assert tuple(v.shape) == tuple(sd[k].shape), k
    model.load_state_dict(sd, strict=True)
    model.to(device).eval()
    return fx, cfg, model


@pytest.mark.parametrize('name', MODEL_NAMES + EXTRA_MODEL_NAMES)
def test_forward_matches_reference(name, device):
    fx, cfg, model = build(name, device)
    batch = fixture_batch(fx, device)
    with torch.no_grad():
        scores = model(batch)
    assert scores.shape == fx['ref/scores'].shape
    assert_close(scores, fx['ref/scores'], TOL, 'scores')
    if hasattr(model, 'get_user_embeddings'):
        with torch.no_grad():
            ue = model.get_user_embeddings(batch)
        assert ue.shape == fx['ref/user_emb'].shape
        assert_close(ue, fx['ref/user_emb'], TOL, 'user embeddings')


@pytest.mark.parametrize('name', MODEL_NAMES + EXTRA_MODEL_NAMES)
def test_losses_and_gradients_match_reference(name, device):
    fx, cfg, model = build(name, device)
    batch = fixture_batch(fx, device)
    if 'ref/loss_cl' not in fx:     # no CL hook exists for NPA / SmallNAML in the reference
        tr = MSERankingTrainer(cfg, model)
        tr.optimizer.zero_grad()
        total, preds, _ = tr.rec_loss(batch)
    else:
        tr = ContrastiveRankingTrainer(cfg, model)
        tr.optimizer.zero_grad()
        total, l_rec, l_cl, preds = tr.losses(batch)
        assert_close(l_rec, fx['ref/loss_mse'], TOL, 'mse')
        assert_close(l_cl, fx['ref/loss_cl'], TOL, 'cl')
    assert_close(total, fx['ref/loss_total'], TOL, 'total loss')
    assert_close(preds, np.maximum(fx['ref/scores'], 0), TOL, 'relu(scores)')
    grads = sub(fx, 'grad')
    if not total.requires_grad:         # ParamFreeRec: nothing trainable reaches the loss (the reference's gradients are all 0)
        assert all(float(np.abs(g).max()) == 0 for g in grads.values())
        return
    total.backward()
    gscale = max(float(np.abs(g).max()) for g in grads.values())
    named = dict(model.named_parameters())
    for k, g in grads.items():
        got = named[k].grad
        assert got is not None, k
        assert_close(got, g, 2e-4, 'grad ' + k, atol=2e-6 * gscale)


@pytest.mark.parametrize('name', ['cl', 'nrms'])
def test_bce_trainer_and_unfused_hooks(name, device):
    fx, cfg, model = build(name, device)
    batch = fixture_batch(fx, device)
    tr = BCELogitsRankingTrainer(cfg, model)
    loss, _, _ = tr.rec_loss(batch)
    assert_close(loss, fx['ref/loss_bce'], TOL, 'bce (fused)')
    # the reference-style unfused hooks: forward() applies the activation, self.L takes (s, t)
    tr2 = MSERankingTrainer(cfg, model)
    s = tr2.raw_scores(batch)
    assert_close(tr2.L(s, batch['targets']), fx['ref/loss_mse'], TOL, 'mse via self.L')
    assert_close(tr2.forward(batch), np.maximum(fx['ref/scores'], 0), TOL, 'forward() = relu(scores)')
    assert_close(tr.L(s, batch['targets']), fx['ref/loss_bce'], TOL, 'bce via self.L')


def _one_impression(batch):
    return {'user_features': {'history': {'title_emb': tuple(t[:1] for t in batch['user_features']['history']['title_emb'])},
                              'other': {}},
            'candidate_features': {'title_emb': tuple(t[:1] for t in batch['candidate_features']['title_emb'])},
            'targets': batch['targets'][:1]}


def test_bce_sigmoid_trainer(device):
    """BCERankingTrainer (training.py:324-331): forward() = sigmoid(scores), L = nn.BCELoss — pinned to the reference value."""
    fx, cfg, model = build('cl', device)
    batch = fixture_batch(fx, device)
    tr = BCERankingTrainer(cfg, model)
    loss, preds, _ = tr.rec_loss(batch)
    sig = 1.0 / (1.0 + np.exp(-fx['ref/scores'].astype(np.float64)))
    assert_close(loss, fx['ref/loss_bce_sigmoid'], TOL, 'BCELoss(sigmoid(s))')
    assert_close(preds, sig, TOL, 'preds = sigmoid(scores)')
    assert_close(tr.forward(batch), sig, TOL, 'forward() = sigmoid(scores)')
    assert_close(tr.L(tr.raw_scores(batch), batch['targets']), fx['ref/loss_bce_sigmoid'], TOL, 'bce via self.L')
    from oracle import xnrs_oracle as O
    want, _ = O.bce_sigmoid_loss(torch.tensor(fx['ref/scores']), torch.tensor(fx['batch/targets']))
    assert_close(want, fx['ref/loss_bce_sigmoid'], 2e-6, 'oracle restatement of nn.BCELoss')


@pytest.mark.parametrize('trainer', [BCELogitsRankingTrainer, BCERankingTrainer])
def test_bce_test_step_ranks_sigmoid_scores(trainer, device):
    """both BCE trainers evaluate sigmoid(score) (training.py:329-331 via forward(); :357 after the loss)"""
    fx, cfg, model = build('cl', device)
    batch = fixture_batch(fx, device)
    out = trainer(cfg, model)._test_step(_one_impression(batch))
    from oracle import xnrs_oracle as O
    s = 1.0 / (1.0 + np.exp(-fx['ref/scores'][0, :, 0].astype(np.float64)))
    assert_close(out['scores'], s, TOL, 'ranked scores = sigmoid(raw)')
    want = O.impression_metrics(fx['batch/targets'][0, :, 0], s.astype(np.float32))
    for k in ('auc', 'rr', 'ndcg@5', 'ndcg@10', 'ctr@1', 'ctr@10'):
        assert abs(out[k] - want[k]) < 1e-6, (k, out[k], want[k])
    ref_loss = float(fx['ref/loss_bce']) if trainer is BCELogitsRankingTrainer else None
    if ref_loss is not None:            # the loss of the single impression differs from the batch loss: just finite
        assert np.isfinite(float(out['loss']))


def test_train_step_matches_torch_adam(device):
    """one ContrastiveRankingTrainer._train_step == reference gradients + torch.optim.Adam defaults."""
    fx, cfg, model = build('cl', device)
    batch = fixture_batch(fx, device)
    before = {k: v.detach().clone() for k, v in model.named_parameters()}
    tr = ContrastiveRankingTrainer(dict(cfg, lr=1e-3), model)
    out = tr._train_step(batch)
    assert_close(out['loss'], fx['ref/loss_total'], TOL, 'loss')
    gmax = max(float(np.abs(g).max()) for g in sub(fx, 'grad').values())
    for k, p in model.named_parameters():
        if np.abs(fx['grad/' + k]).max() < 1e-4 * gmax:
            continue        # analytically-zero gradients (pooler fc2.bias): Adam amplifies rounding noise
        ref = torch.nn.Parameter(before[k].cpu().clone())
        opt = torch.optim.Adam([ref], lr=1e-3)
        ref.grad = torch.tensor(fx['grad/' + k])
        opt.step()
        # Adam's first step moves every weight by ~lr * sign(g): compare the *updates*
        upd, want = (p.detach().cpu() - before[k].cpu()), (ref.detach() - before[k].cpu())
        big = torch.tensor(np.abs(fx['grad/' + k]) > 1e-4 * np.abs(fx['grad/' + k]).max() + 1e-12)
        if big.any():
            assert_close(upd[big], want[big], 2e-3, 'adam update ' + k)


def test_test_step_metrics(device):
    fx, cfg, model = build('cl', device)
    batch = fixture_batch(fx, device)
    one = {'user_features': {'history': {'title_emb': tuple(t[:1] for t in batch['user_features']['history']['title_emb'])},
                             'other': {}},
           'candidate_features': {'title_emb': tuple(t[:1] for t in batch['candidate_features']['title_emb'])},
           'targets': batch['targets'][:1]}
    tr = MSERankingTrainer(cfg, model)
    out = tr._test_step(one)
    from oracle import xnrs_oracle as O
    s = np.maximum(fx['ref/scores'][0, :, 0], 0)
    want = O.impression_metrics(fx['batch/targets'][0, :, 0], s)
    for k in ('auc', 'rr', 'ndcg@5', 'ndcg@10', 'ctr@1', 'ctr@10'):
        assert abs(out[k] - want[k]) < 1e-4, k


def test_item_logit_pooling_equals_pooling_the_gathered_rows(device):
    """ItemLogitPoolFn (pooler fc1 once per distinct item, history given as ids) == AdditivePoolFn on the gathered rows:
    values and every gradient (item vectors, fc1, fc2)"""
    g = torch.Generator().manual_seed(3)
    V, T, A, R, L = 23, 16, 12, 7, 5
    table = torch.randn(V, T, generator=g)
    row_mask = (torch.rand(V, generator=g) > 0.2).float()
    ids = torch.randint(0, V, (R, L), generator=g, dtype=torch.int32)
    w1, b1 = torch.randn(A, T, generator=g) * 0.3, torch.randn(A, generator=g) * 0.1
    w2, b2 = torch.randn(A, generator=g) * 0.3, torch.randn(1, generator=g) * 0.1
    gout = torch.randn(R, T, generator=g)
    outs = []
    for fused in (True, False):
        leaves = [t.clone().to(device).requires_grad_() for t in (table, w1, b1, w2, b2)]
        tb, a1, c1, a2, c2 = leaves
        if fused:
            pooled, _ = K.ItemLogitPoolFn.apply(tb, row_mask.to(device), ids.to(device), a1, c1, a2, c2)
        else:
            x = K.EmbeddingFn.apply(tb, ids.to(device).reshape(-1), None)
            m = row_mask.to(device)[ids.to(device).long().reshape(-1)]
            pooled, _ = K.AdditivePoolFn.apply(x, None, m, a1, c1, a2, c2, R, L, None)
        (pooled * gout.to(device)).sum().backward()
        outs.append([pooled.detach()] + [t.grad for t in leaves])
    for name, a, b in zip(('pooled', 'd table', 'd fc1.weight', 'd fc1.bias', 'd fc2.weight', 'd fc2.bias'), *outs):
        assert_close(a, b, 2e-5, name, atol=1e-6)


def test_fused_title_pool_equals_gemm_plus_pool(device, monkeypatch):
    """AdditivePoolFn's one-launch forward (xnrs_titlepool_fwd: gather -> fc1 -> tanh -> logit -> exp -> per-title sums) ==
    xnrs_gemm(TANH) + xnrs_addpool_fwd on the same ragged, padded TitlePlan rows: pooled vectors, weights, every gradient"""
    from xnrs_b200 import synthetic as syn
    from xnrs_b200.data import TitleStore, plan_titles
    if device == 'cpu':
        monkeypatch.setattr(K, '_precision', K.PRECISIONS['tf32x3'])       # host-logic run: the emulator ignores the precision
    elif K.get_precision() == 'fp32':
        pytest.skip('the fused kernel exists in the tensor-core precisions only')
    monkeypatch.setattr(K, 'FUSED_GATHER_MIN_ROWS', 256)
    D, A = 128, 256
    cat = syn.make_catalogue(300, 20, vocab=900, dim=D, seed=21)
    ids = torch.from_numpy(syn.zipf_news(np.random.default_rng(3), 300, 700)).int()
    ids[::9] = 0
    store = TitleStore(cat.token_table.to(device) * 0.3, cat.title_tokens.to(device))
    plan = plan_titles(store, ids.to(device), True, True).acquire()
    assert plan.rows.numel() >= 256 and plan.tix is not None
    gen = torch.Generator().manual_seed(5)
    params = [torch.randn(A, D, generator=gen) / D ** 0.5, torch.randn(A, generator=gen) * 0.1,
              torch.randn(1, A, generator=gen) / A ** 0.5, torch.randn(1, generator=gen) * 0.1]
    R = plan.uniq.numel()
    gout = torch.randn(R, D, generator=gen).to(device)
    outs = []
    for fused in (True, False):
        monkeypatch.setattr(K, 'FUSED_TITLEPOOL', fused)
        n0 = K.launch_count() if device != 'cpu' else 0
        leaves = [p.clone().to(device).requires_grad_(True) for p in params]
        pooled, attn = K.AdditivePoolFn.apply(store.token_table, plan.rows, None, *leaves, R, 20, plan.seg, plan.tix)
        (pooled * gout).sum().backward()
        outs.append([pooled.detach(), attn.detach()[:plan.n_rows]] + [t.grad for t in leaves])
        assert float(pooled[plan.n_titles:].abs().sum()) == 0           # padded titles pool to exactly 0
    tol = 1e-4 if device != 'cpu' else 2e-5
    for name, a, b in zip(('pooled', 'attn', 'd fc1.weight', 'd fc1.bias', 'd fc2.weight', 'd fc2.bias'), *outs):
        # d fc2.bias is analytically ~0 (the normalised weights are shift invariant up to the 1e-8): absolute floor only
        assert_close(a, b, tol, name, atol=1e-4 if name == 'd fc2.bias' else 1e-6)


@pytest.mark.parametrize('where', ['emulated', pytest.param('cuda', marks=pytest.mark.gpu)])
def test_bf16_storage_title_pool_matches_fp32(where, monkeypatch):
    """XNRS_PREC_BF16: token rows / fc1.weight / hidden layer / its gradient stored in bf16 (tcgen05 kind::f16, fp32 accumulation)
    against the exact-fp32 path on the same ragged TitlePlan: pooled vectors, weights and every gradient within the
    north-star's bf16 tolerance (2e-2 of the tensor scale)"""
    from xnrs_b200 import synthetic as syn
    from xnrs_b200.data import TitleStore, plan_titles
    device = 'cpu' if where == 'emulated' else 'cuda'
    if where == 'emulated':
        monkeypatch.setattr(K, 'call', EMU.call)
    monkeypatch.setattr(K, 'FUSED_GATHER_MIN_ROWS', 256)
    D, A = 256, 256
    cat = syn.make_catalogue(300, 20, vocab=900, dim=D, seed=21)
    ids = torch.from_numpy(syn.zipf_news(np.random.default_rng(3), 300, 700)).int()
    ids[::9] = 0
    store = TitleStore(cat.token_table.to(device) * 0.3, cat.title_tokens.to(device))
    plan = plan_titles(store, ids.to(device), True, True).acquire()
    gen = torch.Generator().manual_seed(5)
    params = [torch.randn(A, D, generator=gen) / D ** 0.5, torch.randn(A, generator=gen) * 0.1,
              torch.randn(1, A, generator=gen) / A ** 0.5, torch.randn(1, generator=gen) * 0.1]
    R = plan.uniq.numel()
    gout = torch.randn(R, D, generator=gen).to(device)
    outs = []
    for prec in ('bf16', 'fp32'):
        monkeypatch.setattr(K, '_precision', K.PRECISIONS[prec])
        leaves = [p.clone().to(device).requires_grad_(True) for p in params]
        pooled, attn = K.AdditivePoolFn.apply(store.token_table, plan.rows, None, *leaves, R, 20, plan.seg, plan.tix)
        (pooled * gout).sum().backward()
        outs.append([pooled.detach(), attn.detach()[:plan.n_rows]] + [t.grad for t in leaves])
    assert getattr(store.token_table, '_xnrs_bf16', None) is not None          # the bf16 twin of the table was used
    for name, a, b in zip(('pooled', 'attn', 'd fc1.weight', 'd fc1.bias', 'd fc2.weight', 'd fc2.bias'), *outs):
        # outputs at the north-star's bf16 bar (2e-2); gradients are sums of bf16-rounded terms with cancellation: 5e-2
        assert_close(a, b, 2e-2 if not name.startswith('d ') else 5e-2, name, atol=1e-3 if name == 'd fc2.bias' else 1e-6)


@pytest.mark.parametrize('where', ['emulated', pytest.param('cuda', marks=pytest.mark.gpu)])
def test_bf16x3_title_pool_is_fp32_accurate(where, monkeypatch):
    """precision 'bf16x3': the token-level launches of the title pooler run 3xBF16 on pre-split planes (cached planes of the
    frozen table, fc1.weight split per step, d_hid written as planes) — pooled vectors, weights and every gradient within the
    fp32 bar (1e-4 of the tensor scale) of the exact-fp32 path on the same ragged TitlePlan"""
    from xnrs_b200 import synthetic as syn
    from xnrs_b200.data import TitleStore, plan_titles
    device = 'cpu' if where == 'emulated' else 'cuda'
    if where == 'emulated':
        monkeypatch.setattr(K, 'call', EMU.call)
    monkeypatch.setattr(K, 'FUSED_GATHER_MIN_ROWS', 256)
    D, A = 256, 256
    cat = syn.make_catalogue(300, 20, vocab=900, dim=D, seed=21)
    ids = torch.from_numpy(syn.zipf_news(np.random.default_rng(3), 300, 700)).int()
    ids[::9] = 0
    store = TitleStore(cat.token_table.to(device) * 0.3, cat.title_tokens.to(device))
    plan = plan_titles(store, ids.to(device), True, True).acquire()
    gen = torch.Generator().manual_seed(5)
    params = [torch.randn(A, D, generator=gen) / D ** 0.5, torch.randn(A, generator=gen) * 0.1,
              torch.randn(1, A, generator=gen) / A ** 0.5, torch.randn(1, generator=gen) * 0.1]
    R = plan.uniq.numel()
    gout = torch.randn(R, D, generator=gen).to(device)
    outs = []
    for prec in ('bf16x3', 'fp32'):
        monkeypatch.setattr(K, '_precision', K.PRECISIONS[prec])
        leaves = [p.clone().to(device).requires_grad_(True) for p in params]
        pooled, attn = K.AdditivePoolFn.apply(store.token_table, plan.rows, None, *leaves, R, 20, plan.seg, plan.tix)
        (pooled * gout).sum().backward()
        outs.append([pooled.detach(), attn.detach()[:plan.n_rows]] + [t.grad for t in leaves])
    assert set(getattr(store.token_table, '_xnrs_bf16x3', {})) >= {True, False}  # cached fp16 (forward) and bf16 (backward) planes
    for name, a, b in zip(('pooled', 'attn', 'd fc1.weight', 'd fc1.bias', 'd fc2.weight', 'd fc2.bias'), *outs):
        assert_close(a, b, 1e-4, name, atol=1e-4 if name == 'd fc2.bias' else 1e-6)


@pytest.mark.parametrize('name', ['cl', 'nrms', 'naml', 'lstur_con', 'npa'])
def test_index_batches_equal_dense_batches(name, device):
    """the index fast path (device-resident token table + int32 news ids: title de-duplication, padding-free pooling,
    one encoder pass for both sides, per-article pooling logits / article-level NAML) gives the scores and parameter
    gradients of the reference-format dense batch built from the same ids"""
    from xnrs_b200 import synthetic as syn
    from xnrs_b200.data import TitleStore
    fx = load_npz('model_' + name)
    cfg = dict(fixture_cfg(fx), device=device)
    n_news, S, H, D = 40, cfg['seq_len'], cfg['hist_len'], cfg['d_backbone']
    cat = syn.make_catalogue(n_news, S, vocab=100, dim=D, seed=11, with_abstract=(name == 'naml'),
                             n_categories=cfg['n_categories'], n_subcategories=cfg['n_subcategories'])
    raw = syn.make_train_batch(n_news, 12, H, n_neg=cfg['n_negatives'], n_users=cfg['n_users'], seed=12)   # 12*(H+1+K) >= 64 slots
    store = TitleStore(cat.token_table.to(device), cat.title_tokens.to(device))
    astore = TitleStore(store.token_table, cat.abstract_tokens.to(device)) if name == 'naml' else None
    grads, scores = [], []
    for kind in ('index', 'dense'):
        model = make_model(cfg)
        model.load_state_dict({k: torch.tensor(v) for k, v in sub(fx, 'sd').items()})
        model.to(device).eval()
        batch = (syn.index_batch(store, cat, raw, device, abstract_store=astore) if kind == 'index'
                 else syn.dense_batch(cat, raw, with_abstract=(name == 'naml')))
        s = model(batch)
        (s * torch.linspace(-1, 1, s.numel(), device=s.device).view_as(s)).sum().backward()
        scores.append(s.detach())
        grads.append({k: (p.grad.detach().clone() if p.grad is not None else torch.zeros_like(p)) for k, p in model.named_parameters()})
    assert_close(scores[0], scores[1], 2e-5, 'scores')
    gmax = max(float(g.abs().max()) for g in grads[1].values())
    for k in grads[1]:
        assert_close(grads[0][k], grads[1][k], 1e-4, 'grad ' + k, atol=2e-6 * gmax)


@pytest.mark.parametrize('name', ['cl', 'nrms'])       # ragged (padding-free) plan / fixed-length plan (self-attention)
def test_prefetched_id_plumbing_is_equivalent_and_single_use(name, device):
    """ParentRec.prefetch computes the merged-side TitlePlan ahead of the step (side stream on CUDA): same scores as the
    in-line plumbing, and the plan is consumed by exactly one forward (a recycled batch object is planned afresh)"""
    from xnrs_b200 import synthetic as syn
    from xnrs_b200.data import TitleStore
    fx = load_npz('model_' + name)
    cfg = dict(fixture_cfg(fx), device=device)
    cat = syn.make_catalogue(40, cfg['seq_len'], vocab=100, dim=cfg['d_backbone'], seed=11)
    raw = syn.make_train_batch(40, 12, cfg['hist_len'], n_neg=cfg['n_negatives'], n_users=cfg['n_users'], seed=12)
    store = TitleStore(cat.token_table.to(device), cat.title_tokens.to(device))
    model = make_model(cfg)
    model.load_state_dict({k: torch.tensor(v) for k, v in sub(fx, 'sd').items()})
    model.to(device).eval()
    batch = syn.index_batch(store, cat, raw, device)
    with torch.no_grad():
        plain = model(batch)
        assert model.prefetch(batch)
        hist = batch['user_features']['history']['title_emb']
        assert hist._merged is not None
        fetched = model(batch)
        assert hist._merged is None                                   # consumed
        again = model(batch)                                          # in-line plumbing again
    assert torch.equal(plain, fetched) and torch.equal(plain, again)


@pytest.mark.timeout(120)
def test_prefetch_over_recycled_batches(monkeypatch):
    """prefetch / consume over many recycled batch objects: every forward sees a fresh plan for exactly its batch"""
    import _kernel_emulator as EMU
    from xnrs_b200 import synthetic as syn
    from xnrs_b200.data import TitleStore
    monkeypatch.setattr(K, 'call', EMU.call)
    fx = load_npz('model_cl')
    cfg = dict(fixture_cfg(fx), device='cpu')
    cat = syn.make_catalogue(40, cfg['seq_len'], vocab=100, dim=cfg['d_backbone'], seed=11)
    store = TitleStore(cat.token_table, cat.title_tokens)
    model = make_model(cfg)
    model.load_state_dict({k: torch.tensor(v) for k, v in sub(fx, 'sd').items()})
    model.eval()
    batches = [syn.index_batch(store, cat, syn.make_train_batch(40, 12, cfg['hist_len'], n_neg=cfg['n_negatives'],
                                                                n_users=cfg['n_users'], seed=20 + i), 'cpu') for i in range(3)]
    with torch.no_grad():
        want = [model(b) for b in batches]
        assert model.prefetch(batches[0])
        for i in range(12):
            got = model(batches[i % 3])
            assert model.prefetch(batches[(i + 1) % 3])
            assert torch.equal(got, want[i % 3])


def test_active_row_adam_equals_dense_adam(device, monkeypatch):
    """§8(f) row 1 / SURVEY §7 hard part 4: Adam restricted to the rows of an embedding table that have received a gradient
    so far (xnrs_adam_rows) == torch's dense Adam over the whole table, BIT FOR BIT, over 25 steps that keep touching new
    users: a never-touched row has g = m = v = 0 and the dense update leaves it exactly unchanged."""
    from xnrs_b200 import training as TR
    runs = []
    for min_rows in (5, 10 ** 9):                  # 5: the 12-row user table and the 7-row category table go row-sparse
        monkeypatch.setattr(TR.RankingTrainer, 'sparse_min_rows', min_rows)
        fx, cfg, model = build('lstur_con', device)
        tr = ContrastiveRankingTrainer(dict(cfg, lr=1e-2), model)
        assert len(tr.optimizer.tables) == (2 if min_rows == 5 else 0)
        gen = torch.Generator().manual_seed(3)
        for step in range(25):
            batch = fixture_batch(fx, device)
            B = batch['targets'].shape[0]
            n_u = cfg['n_users'] if step >= 8 else 4                    # the first steps only see users 1..4
            batch['user_features']['other']['user_index'] = torch.randint(1, n_u + 1, (B, 1), generator=gen).int().to(device)
            tr._train_step(batch)
        runs.append({k: v.detach().cpu().clone() for k, v in model.named_parameters()})
        if min_rows == 5:
            t = tr.optimizer.tables[-1]
            assert 0 < int(t['count']) <= t['V'] - 1                    # padding row 0 never becomes active
    for k in runs[0]:
        if device == 'cpu':                 # deterministic arithmetic: the two optimisers agree bit for bit
            assert torch.equal(runs[0][k], runs[1][k]), k
        elif not k.endswith('fc2.bias'):    # (fc2.bias of a pooler: analytically-zero gradient, pure rounding noise under Adam)
            # fp32 atomics make two GPU runs of the SAME code differ in the last bits and Adam amplifies that; the
            # kernel-level bit-exactness test is test_gpu_kernels.test_active_row_adam_kernel_is_bit_identical_to_the_dense_pass
            assert_close(runs[0][k], runs[1][k], 5e-2, k, atol=1e-2)


def test_autograd_grad_with_a_trainer_attached_leaves_the_optimiser_buffer_alone(device):
    """weight gradients go straight into FlatAdam's flat buffer ONLY inside the trainers' own backward (K.direct_grads);
    torch.autograd.grad on the same model returns ordinary gradient tensors and does not touch flat_g"""
    fx, cfg, model = build('cl', device)
    batch = fixture_batch(fx, device)
    tr = ContrastiveRankingTrainer(cfg, model)
    tr.optimizer.zero_grad()
    total, *_ = tr.losses(batch)
    params = [p for p in model.parameters() if p.numel() > 1]
    grads = torch.autograd.grad(total, params, allow_unused=True)
    assert all(g is not None for g in grads)
    assert float(tr.optimizer.flat_g.abs().sum()) == 0.0
    named = dict(model.named_parameters())
    for (k, p), g in zip([(k, p) for k, p in named.items() if p.numel() > 1], grads):
        assert_close(g, fx['grad/' + k], 2e-4, 'autograd.grad ' + k, atol=2e-6)
    total2, *_ = tr.losses(batch)
    with K.direct_grads():
        total2.backward()
    assert float(tr.optimizer.flat_g.abs().sum()) > 0.0


def test_flat_adam_refuses_detached_gradients(device):
    """FlatAdam updates from ONE flat gradient buffer; if user code replaces the parameters' .grad views
    (model.zero_grad(set_to_none=True)) the step must fail loudly instead of training on zeros"""
    fx, cfg, model = build('cl', device)
    trainer = ContrastiveRankingTrainer(dict(cfg, lr=1e-3), model)
    batch = fixture_batch(fx, device)
    ids = {t: i for i, t in enumerate(sorted(set(batch['main_theme'])))}
    batch['main_theme'] = torch.tensor([ids[t] for t in batch['main_theme']], dtype=torch.int32, device=device)
    trainer._train_step(batch)                         # fine
    model.zero_grad(set_to_none=True)
    with pytest.raises(RuntimeError, match='flat gradient buffer'):
        trainer._train_step(batch)


def test_index_batch_category_lookups_are_optional():
    """the bench's input pipeline leaves the category / subcategory lookups out for models that never read them (8 device ops
    per batch); the title ids, targets and theme labels are the same either way"""
    from xnrs_b200 import synthetic as syn
    from xnrs_b200.data import TitleStore
    cat = syn.make_catalogue(40, 12, vocab=100, dim=32, seed=3)
    raw = syn.make_train_batch(40, 6, 5, seed=4)
    store = TitleStore(cat.token_table, cat.title_tokens)
    lean, full = (syn.index_batch(store, cat, raw, 'cpu', categories=c) for c in (False, True))
    for side in (lambda b: b['candidate_features'], lambda b: b['user_features']['history']):
        assert 'category_index' not in side(lean) and 'subcategory_index' not in side(lean)
        assert 'category_index' in side(full) and 'subcategory_index' in side(full)
        assert torch.equal(side(lean)['title_emb'].news_ids, side(full)['title_emb'].news_ids)
    assert torch.equal(lean['targets'], full['targets']) and torch.equal(lean['main_theme'], full['main_theme'])


@pytest.mark.parametrize('where', ['emulated', pytest.param('cuda', marks=pytest.mark.gpu)])
def test_bf16x3_attention_projections_are_fp32_accurate(where, monkeypatch):
    """precision 'bf16x3': the q / k / v projections of self-attention over table-gathered rows (NRMS) and their weight gradients run
    the 3-pass 16-bit split on pre-split planes (cached fp16 / bf16 planes of the frozen table, weights and gradients split per
    step) — output and every parameter gradient within the fp32 bar of the exact-fp32 path"""
    from xnrs_b200.models import components as C
    device = 'cpu' if where == 'emulated' else 'cuda'
    if where == 'emulated':
        monkeypatch.setattr(K, 'call', EMU.call)
    monkeypatch.setattr(K, 'FUSED_GATHER_MIN_ROWS', 256)
    V, D, h, R, L = 500, 256, 16, 40, 12
    gen = torch.Generator().manual_seed(9)
    table = (torch.randn(V, D, generator=gen) * 0.3).to(device)
    rows = torch.randint(0, V, (R * L,), generator=gen).int().to(device)
    ln = torch.randint(1, L + 1, (R,), generator=gen)
    mask = (torch.arange(L)[None, :] < ln[:, None]).float().reshape(-1).to(device)
    gout = torch.randn(R * L, D, generator=gen).to(device)
    torch.manual_seed(3)
    mod = C.MultiHeadAttention(h, D, dropout=0.0).to(device).eval()
    outs = []
    for prec in ('bf16x3', 'fp32'):
        monkeypatch.setattr(K, '_precision', K.PRECISIONS[prec])
        mod.zero_grad(set_to_none=True)
        y = mod.attend(table, rows, mask, R, L)
        (y * gout).sum().backward()
        outs.append([y.detach()] + [p.grad.clone() for p in mod.parameters()])
    assert set(getattr(table, '_xnrs_bf16x3', {})) >= {True, False}
    names = ['out'] + ['d ' + n for n, _ in mod.named_parameters()]
    gmax = max(float(t.abs().max()) for t in outs[1][1:])
    for name, a, b in zip(names, *outs):        # d k_linear.bias is analytically 0 (a key bias shifts every score of a row equally)
        assert_close(a, b, 1e-4, name, atol=1e-5 * gmax)


def test_flat_adam_auxiliary_floats_ride_in_front_of_the_gradient():
    """FlatAdam keeps 4 auxiliary floats in front of the flat gradient (one buffer: data-parallel scalars such as the InfoNCE value
    are summed by the SAME all-reduce as the gradient); zero_grad clears them, every .grad is a 16-byte aligned view behind them"""
    from xnrs_b200.training import FlatAdam
    torch.manual_seed(0)
    lin = torch.nn.Linear(7, 5)
    opt = FlatAdam(list(lin.parameters()), lr=1e-3)
    assert opt.g_store.numel() == opt.flat_g.numel() + 4
    assert opt.g_aux.data_ptr() == opt.g_store.data_ptr() and opt.flat_g.data_ptr() == opt.g_store.data_ptr() + 16
    for p in lin.parameters():
        assert p.grad.untyped_storage().data_ptr() == opt.g_store.untyped_storage().data_ptr()
        assert (p.grad.data_ptr() - opt.g_store.data_ptr()) % 16 == 0
    opt.g_aux.fill_(3.0)
    lin.weight.grad.fill_(2.0)
    opt.zero_grad()
    assert float(opt.g_aux.abs().sum()) == 0.0 and float(opt.flat_g.abs().sum()) == 0.0
