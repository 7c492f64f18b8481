"""Full-width parity on the GPU: the five models at the reference layer widths (D=768, E=256, A=256, 16 heads) on both
shape sets (north-star S=30/H=50 and reference S=50/H=25), fed through the INDEX fast path (device-resident token
table + int32 news ids), against the CPU oracle on the dense batch built from the same ids."""
import pytest
import torch

from _common import assert_close
from oracle import xnrs_oracle as O
from xnrs_b200 import synthetic as syn
from xnrs_b200.data import TitleStore
from xnrs_b200.models import make_model
from xnrs_b200.training import ContrastiveRankingTrainer, MSERankingTrainer

pytestmark = pytest.mark.gpu
DEV = 'cuda'
N_NEWS, VOCAB, N_USERS = 400, 3000, 1000

BASE = dict(scoring='dot', text_features=['title_emb'], catg_features=[], title_emb_dim=256, total_emb_dim=256,
            d_backbone=768, n_heads=16, p_dropout=0., bias=False, n_categories=19, n_subcategories=264,
            n_users=N_USERS, cat_emb_dim=16, sub_emb_dim=16, user_emb_dim=64, contrastive_temperature=0.08,
            contrastive_lambda=0.1, lr=1e-4, device=DEV)
MODELS = {
    'cl': dict(model='standard', contrastive_lambda=0.01),
    'nrms': dict(model='NRMS'),
    'naml': dict(model='NAML', text_features=['title_emb', 'abstract_emb'],
                 catg_features=['category_index', 'subcategory_index']),
    'lstur': dict(model='LSTUR', total_emb_dim=272, long_term_method='embedding', long_short_term_method='con',
                  p_user_dropout=0.0, catg_features=['category_index']),
    'npa': dict(model='NPA'),
}


def oracle_forward(name, P, batch, cfg):
    if name in ('cl', 'nrms'):
        nh = cfg['n_heads'] if name == 'nrms' else 0
        r, u, _ = O.parent_forward(P, batch, nh, return_embeddings=True)
        return r, u.squeeze(1)
    if name == 'naml':
        r, u, _ = O.naml_forward(P, batch, return_embeddings=True)
        return r, u.squeeze(1)
    if name == 'lstur':
        r, u, _ = O.lstur_forward(P, batch, 'con', cfg['st_hist_len'], return_embeddings=True)
        return r, u.squeeze(1)
    return O.npa_forward(P, batch), None


@pytest.mark.parametrize('prec', ['fp32', 'tf32x3'])
@pytest.mark.parametrize('S,H', [(30, 50), (50, 25)])
@pytest.mark.parametrize('name', list(MODELS))
def test_index_path_matches_oracle(name, S, H, prec, monkeypatch):
    from xnrs_b200 import kernels as K
    monkeypatch.setattr(K, '_precision', K.PRECISIONS[prec])
    cfg = dict(BASE, **MODELS[name], seq_len=S, hist_len=H, st_hist_len=H)
    B = 5
    cat = syn.make_catalogue(N_NEWS, S, VOCAB, 768, seed=3, with_abstract=(name == 'naml'))
    raw = syn.make_train_batch(N_NEWS, B, H, n_users=N_USERS, seed=4)
    raw['hist_ids'][1, 3:] = 0                      # a short history
    torch.manual_seed(1)
    model = make_model(cfg)
    with torch.no_grad():                           # default init keeps scores tiny; widen the dynamic range
        for p in model.parameters():
            if p.dim() > 1:
                p.mul_(1.5)
    P = {k: v.detach().clone().requires_grad_(v.is_floating_point()) for k, v in model.state_dict().items()}
    trainer = (MSERankingTrainer if name == 'npa' else ContrastiveRankingTrainer)(cfg, model)
    model.eval()
    store = TitleStore(cat.token_table.to(DEV), cat.title_tokens.to(DEV))
    astore = TitleStore(store.token_table, cat.abstract_tokens.to(DEV)) if name == 'naml' else None
    batch = syn.index_batch(store, cat, raw, DEV, abstract_store=astore)
    dense = syn.dense_batch(cat, raw, with_abstract=(name == 'naml'))

    want_scores, want_u = oracle_forward(name, P, dense, cfg)
    trainer.optimizer.zero_grad()
    if name == 'npa':
        total, preds, _ = trainer.rec_loss(batch)
        want_total, _ = O.mse_relu_loss(want_scores, dense['targets'])
    else:
        total, l_rec, l_cl, preds = trainer.losses(batch)
        want_total, want_rec, want_cl = O.contrastive_train_loss(
            want_scores, dense['targets'], want_u, dense['main_theme'].long(), cfg['contrastive_temperature'],
            cfg['contrastive_lambda'])
        assert_close(l_cl, want_cl, 1e-4, 'InfoNCE')
    # relative to the scale of the raw scores (after the ReLU the surviving values can be arbitrarily small)
    assert_close(preds, torch.relu(want_scores), 1e-4, 'relu(scores)', atol=1e-4 * float(want_scores.abs().max()))
    assert_close(total, want_total, 1e-4, 'loss')
    total.backward()
    want_total.backward()
    gmax = max(float(v.grad.abs().max()) for v in P.values() if v.grad is not None)
    for k, p in model.named_parameters():
        want = P[k].grad if P[k].grad is not None else torch.zeros_like(P[k])
        assert_close(p.grad, want, 2e-4, 'grad ' + k, atol=2e-6 * gmax)

    # the dense reference-format batch (moved to the device by the encoders themselves) gives the same scores
    with torch.no_grad():
        dense_scores = model(dense)
        index_scores = model(batch)
    # same kernels, different row sets (title dedup changes tile / split-K composition): fp32 summation-order noise only
    assert_close(dense_scores, index_scores, 1e-5, 'dense vs index path')
    assert_close(index_scores, want_scores, 1e-4, 'scores')


@pytest.mark.parametrize('name', ['cl', 'nrms', 'naml'])
def test_title_dedup_gives_the_same_loss_and_gradients(name, monkeypatch):
    """encoding each distinct article of the batch once (default) == encoding every (impression, slot) title"""
    from xnrs_b200.models.components import encoder_options
    cfg = dict(BASE, **MODELS[name], seq_len=30, hist_len=50, st_hist_len=50)
    cat = syn.make_catalogue(60, 30, VOCAB, 768, seed=3, with_abstract=(name == 'naml'))     # tiny catalogue: many repeats
    raw = syn.make_train_batch(60, 6, 50, n_users=N_USERS, seed=4)
    store = TitleStore(cat.token_table.to(DEV), cat.title_tokens.to(DEV))
    astore = TitleStore(store.token_table, cat.abstract_tokens.to(DEV)) if name == 'naml' else None
    results = []
    for dedup in (True, False):
        torch.manual_seed(1)
        model = make_model(cfg)
        encoder_options(model, dedup_titles=dedup)
        trainer = ContrastiveRankingTrainer(cfg, model)
        model.eval()
        trainer.optimizer.zero_grad()
        total, _, _, preds = trainer.losses(syn.index_batch(store, cat, raw, DEV, abstract_store=astore))
        total.backward()
        results.append((total.detach().clone(), preds.detach().clone(), trainer.optimizer.flat_g.clone()))
    (l0, p0, g0), (l1, p1, g1) = results
    assert_close(l0, l1, 1e-6, 'loss')
    assert_close(p0, p1, 1e-5, 'predictions', atol=1e-6)
    # identical mathematics; only the fp32 summation order differs (scatter-add, split-K atomics on the smaller GEMMs).
    # NRMS' two attention stacks amplify that noise most (measured 1.8e-4 of the gradient scale)
    assert_close(g0, g1, 5e-4 if name == 'nrms' else 1e-5, 'flat gradient')


def test_item_logit_pooling_gives_the_same_loss_and_gradients(monkeypatch):
    """CL: pooling the history from the batch's distinct article vectors (pooler fc1 once per article, default) ==
    gathering the (b,H,E) history first and pooling it slot by slot like the reference (user_encoding.py:69-77)"""
    from xnrs_b200.models.components import ParentRec
    cfg = dict(BASE, **MODELS['cl'], seq_len=30, hist_len=50, st_hist_len=50)
    cat = syn.make_catalogue(200, 30, VOCAB, 768, seed=3)
    raw = syn.make_train_batch(200, 16, 50, n_users=N_USERS, seed=4)
    raw['hist_ids'][2, 5:] = 0
    store = TitleStore(cat.token_table.to(DEV), cat.title_tokens.to(DEV))
    results = []
    for on in (True, False):
        torch.manual_seed(1)
        model = make_model(cfg)
        assert isinstance(model, ParentRec)
        model.item_logits = on
        with torch.no_grad():
            for p in model.parameters():
                if p.dim() > 1:
                    p.mul_(1.5)
        trainer = ContrastiveRankingTrainer(cfg, model)
        model.eval()
        trainer.optimizer.zero_grad()
        total, _, _, preds = trainer.losses(syn.index_batch(store, cat, raw, DEV))
        total.backward()
        results.append((total.detach().clone(), preds.detach().clone(), trainer.optimizer.flat_g.clone()))
    (l0, p0, g0), (l1, p1, g1) = results
    assert_close(l0, l1, 1e-6, 'loss')
    assert_close(p0, p1, 1e-5, 'predictions', atol=1e-6)
    assert_close(g0, g1, 2e-5, 'flat gradient')


def test_naml_article_level_encoding_gives_the_same_loss_and_gradients(monkeypatch):
    """NAML: the 4-view news encoder once per distinct article + item-logit user pooling (default) == once per slot"""
    from xnrs_b200.models.zoo import NAML
    cfg = dict(BASE, **MODELS['naml'], seq_len=30, hist_len=50, st_hist_len=50)
    cat = syn.make_catalogue(200, 30, VOCAB, 768, seed=3, with_abstract=True)
    raw = syn.make_train_batch(200, 16, 50, n_users=N_USERS, seed=4)
    raw['hist_ids'][2, 5:] = 0
    store = TitleStore(cat.token_table.to(DEV), cat.title_tokens.to(DEV))
    astore = TitleStore(store.token_table, cat.abstract_tokens.to(DEV))
    results = []
    for on in (True, False):
        torch.manual_seed(1)
        model = make_model(cfg)
        assert isinstance(model, NAML)
        model.article_level = on
        with torch.no_grad():
            for p in model.parameters():
                if p.dim() > 1:
                    p.mul_(1.5)
        trainer = ContrastiveRankingTrainer(cfg, model)
        model.eval()
        trainer.optimizer.zero_grad()
        total, _, _, preds = trainer.losses(syn.index_batch(store, cat, raw, DEV, abstract_store=astore))
        total.backward()
        results.append((total.detach().clone(), preds.detach().clone(), trainer.optimizer.flat_g.clone()))
    (l0, p0, g0), (l1, p1, g1) = results
    assert_close(l0, l1, 1e-6, 'loss')
    assert_close(p0, p1, 1e-5, 'predictions', atol=1e-6)
    assert_close(g0, g1, 2e-5, 'flat gradient')


def _chunked_oracle_step(P, cat, raw, temperature, lam, chunk=64):
    """loss and parameter gradients of the CL train step on the FULL batch with bounded host memory: the MSE term is a
    sum over impressions; the InfoNCE term couples users, so pass 1 collects every user embedding (no grad) and
    differentiates the loss w.r.t. them, pass 2 re-runs each chunk with grad and back-propagates
    rec_chunk + <u_chunk, lam * dCL/du_chunk>.  Exact (same sums as one big autograd graph, chunk order aside)."""
    B = raw['hist_ids'].shape[0]
    N = raw['cand_ids'].shape[1]
    sl = lambda a, b: {k: v[a:b] for k, v in raw.items()}
    us = []
    with torch.no_grad():
        for a in range(0, B, chunk):
            _, u, _ = O.parent_forward(P, syn.dense_batch(cat, sl(a, a + chunk)), return_embeddings=True)
            us.append(u.squeeze(1))
    u_all = torch.cat(us).requires_grad_(True)
    l_cl = O.contrastive_loss(u_all, raw['main_theme'].long(), temperature)
    (d_u,) = torch.autograd.grad(l_cl, u_all)
    l_rec = 0.0
    for a in range(0, B, chunk):
        d = syn.dense_batch(cat, sl(a, a + chunk))
        s, u, _ = O.parent_forward(P, d, return_embeddings=True)
        rec = ((torch.relu(s) - d['targets']) ** 2).sum() / (B * N)
        (rec + lam * (u.squeeze(1) * d_u[a:a + chunk]).sum()).backward()
        l_rec += float(rec.detach())
    return l_rec + lam * float(l_cl.detach()), l_rec, float(l_cl.detach())


@pytest.mark.parametrize('prec,tol', [('tf32x3', 1e-4), ('bf16x3', 1e-4), ('bf16', 2e-2)])
def test_bench_config_step_matches_oracle(prec, tol, monkeypatch):
    """ONE step of bench.py's headline configuration — B=1024 impressions, S=30, H=50, 65 238-news catalogue, 100k-row token
    table, 3xTF32, title de-duplication + padding-free pooling + prefetched id plumbing + item-logit user pooling — checked
    against the CPU oracle on the same ids: total / rec / CL loss and EVERY parameter gradient of the flat buffer."""
    import bench
    from xnrs_b200 import kernels as K
    monkeypatch.setattr(K, '_precision', K.PRECISIONS[prec])
    B = 1024
    cfg = dict(bench.CL_CFG, device=DEV)
    cat = syn.make_catalogue(bench.N_NEWS, bench.SEQ_LEN, bench.VOCAB, 768, seed=0)
    raw = syn.make_train_batch(bench.N_NEWS, B, bench.HIST_LEN, seed=1000)
    store = TitleStore(cat.token_table.to(DEV), cat.title_tokens.to(DEV))
    torch.manual_seed(0)
    model = make_model(cfg)
    P = {k: v.detach().clone().requires_grad_(True) for k, v in model.state_dict().items()}
    trainer = ContrastiveRankingTrainer(cfg, model)
    model.train()
    batch = syn.index_batch(store, cat, raw, DEV)
    assert trainer.prefetch(batch)                      # the bench loop plans the ids of a batch ahead of its step
    trainer.optimizer.zero_grad()
    total, l_rec, l_cl, _ = trainer.losses(batch)
    total.backward()
    torch.cuda.synchronize()
    assert int(_lib_fallbacks()) >= 0
    want_total, want_rec, want_cl = _chunked_oracle_step(P, cat, raw, cfg['contrastive_temperature'], cfg['contrastive_lambda'])
    assert_close(l_rec, torch.tensor(want_rec), tol, 'rec loss')
    assert_close(l_cl, torch.tensor(want_cl), tol, 'cl loss')
    assert_close(total, torch.tensor(want_total), tol, 'total loss')
    named = dict(model.named_parameters())
    gscale = max(float(v.grad.abs().max()) for v in P.values() if v.grad is not None)
    for k, v in P.items():
        if v.grad is None:
            continue
        # absolute floor relative to the largest gradient of the model: column sums with heavy cancellation (the poolers' fc1.bias
        # gradient = sum over rows of d_hid, whose rows sum to ~0 per group) carry fp32 summation-order noise of 2e-5 ... 1.7e-4 of
        # their own scale from run to run (atomics), in the oracle's fp32 as well
        assert_close(named[k].grad, v.grad, 2 * tol, 'grad ' + k, atol=(1e-5 if prec in ('tf32x3', 'bf16x3') else 2e-3) * gscale)


def _lib_fallbacks():
    from xnrs_b200 import _lib
    return _lib.lib().xnrs_gemm_simt_fallbacks()


def test_cuda_graph_replay_equals_eager_steps(monkeypatch):
    """12 training steps through GraphedStep (first visit of a shape bucket eager, second captured, the rest replayed with
    ONE launch) against the same steps launched kernel by kernel.  Both trainers start every step from the SAME optimiser
    state (copied over), so each replayed step is compared with its eager twin directly: loss, the whole flat gradient and the
    Adam update.  (fp32 atomics in split-K / scatter reductions make even two eager runs differ in the last bits, and Adam
    turns the sign of a ~0 gradient component into a full +-lr step: free-running trajectories are not comparable bit for bit.)"""
    import bench
    from xnrs_b200 import kernels as K
    from xnrs_b200.distributed import DataParallelTrainer
    from xnrs_b200.graphs import GraphedStep
    monkeypatch.setattr(K, '_precision', K.PRECISIONS['tf32x3'])
    B = 256
    cfg = dict(bench.CL_CFG, device=DEV, lr=1e-3)
    cat = syn.make_catalogue(5000, bench.SEQ_LEN, 20000, 768, seed=0)
    store = TitleStore(cat.token_table.to(DEV), cat.title_tokens.to(DEV))
    raws = [syn.make_train_batch(5000, B, bench.HIST_LEN, seed=50 + (i % 3)) for i in range(12)]   # three distinct batches recur
    sides = []
    for graphed in (False, True):
        torch.manual_seed(0)
        model = make_model(cfg)
        trainer = ContrastiveRankingTrainer(cfg, model, graph_safe=True)
        model.train()
        dp = DataParallelTrainer(trainer)
        sides.append((trainer, dp, GraphedStep(dp) if graphed else None))
    (t_e, dp_e, _), (t_g, dp_g, stepper) = sides
    for i, raw in enumerate(raws):
        o_e, o_g = t_e.optimizer, t_g.optimizer
        for a, b in ((o_g.flat_p, o_e.flat_p), (o_g.m, o_e.m), (o_g.v, o_e.v), (o_g.step_dev, o_e.step_dev), (o_g.bc_dev, o_e.bc_dev)):
            a.copy_(b)
        before = o_e.flat_p.clone()
        be, bg = syn.index_batch(store, cat, raw, DEV), syn.index_batch(store, cat, raw, DEV)
        if i % 2:
            dp_e.prefetch(be)
            dp_g.prefetch(bg)                            # both the prefetched and the in-line plan reach the graph
        out_e = dp_e.train_step(be)
        out_g = stepper.step(bg)
        assert_close(out_g['loss'], out_e['loss'], 1e-6, f'loss of step {i}')
        assert_close(out_g['loss_cl'], out_e['loss_cl'], 1e-6, f'InfoNCE of step {i}')
        assert_close(o_g.flat_g, o_e.flat_g, 2e-5, f'flat gradient of step {i}')
        big = o_e.flat_g.abs() > 1e-3 * o_e.flat_g.abs().max()          # Adam amplifies the rounding noise of ~0 gradients
        assert_close((o_g.flat_p - before)[big], (o_e.flat_p - before)[big], 1e-3, f'Adam update of step {i}')
    assert stepper.replays >= 6 and 1 <= stepper.captures <= 3 and 1 <= stepper.eager_steps <= 3, (stepper.replays, stepper.captures, stepper.eager_steps)
