"""World-size-2 data parallelism on CPU (gloo): the N>1 host path — batch sharding, the global-batch InfoNCE
(all-gather of user embeddings + labels, all-reduced embedding gradient) and the single flat gradient
all-reduce — must reproduce the single-process full-batch step.  Kernels are emulated (host-logic test)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from _common import fixture_batch, fixture_cfg, load_npz, sub


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _slice_batch(batch, lo, hi):
    def sl(v):
        if isinstance(v, torch.Tensor):
            return v[lo:hi]
        if isinstance(v, tuple):
            return tuple(t[lo:hi] for t in v)
        if isinstance(v, dict):
            return {k: sl(x) for k, x in v.items()}
        if isinstance(v, list):
            return v[lo:hi]
        return v
    return sl(batch)


def _build(name):
    import _kernel_emulator as EMU
    from xnrs_b200 import kernels as K
    from xnrs_b200.models import make_model
    from xnrs_b200.training import ContrastiveRankingTrainer
    K.call = EMU.call
    fx = load_npz('model_' + name)
    cfg = dict(fixture_cfg(fx), device='cpu', lr=1e-3)
    model = make_model(cfg)
    model.load_state_dict({k: torch.tensor(v) for k, v in sub(fx, 'sd').items()})
    model.eval()                       # NRMS' attention dropout (p=0.1) off: the comparison must be deterministic
    return fx, ContrastiveRankingTrainer(cfg, model)


def _worker(rank, world, port, name, out_dir):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    torch.set_num_threads(1)
    from xnrs_b200.distributed import DataParallelTrainer, shard_range
    fx, tr = _build(name)
    batch = fixture_batch(fx)
    # labels must be numbered consistently across ranks (the reference numbers theme strings per batch)
    ids = {t: i for i, t in enumerate(sorted(set(batch['main_theme'])))}
    batch['main_theme'] = torch.tensor([ids[t] for t in batch['main_theme']], dtype=torch.int32)
    lo, hi = shard_range(batch['targets'].shape[0], rank, world)
    dp = DataParallelTrainer(tr)
    dp.SPARSE_MIN_ROWS = 1                 # the fixture's 12-row user table takes the (ids, rows) exchange path
    out = dp.train_step(_slice_batch(batch, lo, hi))
    assert (dp.last_sparse_tables >= 1) == name.startswith('lstur')     # user table (and the category table) went sparse
    np.savez(os.path.join(out_dir, f'r{rank}.npz'), g=tr.optimizer.flat_g.numpy(), p=tr.optimizer.flat_p.numpy(),
             cl=out['loss_cl'].numpy())
    dist.destroy_process_group()


@pytest.mark.parametrize('name', ['cl', 'nrms', 'lstur_con'])
def test_two_rank_step_equals_full_batch_step(name, tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), name, str(tmp_path)), nprocs=world, join=True)
    fx, tr = _build(name)
    batch = fixture_batch(fx)
    tr.optimizer.zero_grad()
    total, l_rec, l_cl, _ = tr.losses(batch)
    total.backward()
    want_g = tr.optimizer.flat_g.clone().numpy()
    tr.optimizer.step()
    r0, r1 = (np.load(tmp_path / f'r{r}.npz') for r in range(world))
    np.testing.assert_allclose(r0['g'], r1['g'], rtol=0, atol=0)                 # identical after the all-reduce
    scale = np.abs(want_g).max()
    np.testing.assert_allclose(r0['g'] / world, want_g, rtol=0, atol=2e-5 * scale)   # averaged == full batch
    np.testing.assert_allclose(r0['cl'], l_cl.detach().numpy(), rtol=1e-5)           # global-batch InfoNCE value
    np.testing.assert_allclose(r0['p'], r1['p'], rtol=0, atol=0)                     # replicas stay in lock step
    big = np.abs(want_g) > 1e-4 * scale          # Adam's first step amplifies rounding noise on ~zero gradients
    np.testing.assert_allclose(r0['p'][big], tr.optimizer.flat_p.numpy()[big], rtol=0, atol=5e-6)


def test_shard_range_is_a_partition():
    from xnrs_b200.distributed import shard_range
    for n in (0, 1, 7, 64, 376471):
        for w in (1, 2, 3, 8):
            cuts = [shard_range(n, r, w) for r in range(w)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            assert all(cuts[i][1] == cuts[i + 1][0] for i in range(w - 1))
            assert max(b - a for a, b in cuts) - min(b - a for a, b in cuts) <= 1


def _eval_setup(name='cl'):
    import _kernel_emulator as EMU
    from xnrs_b200 import kernels as K
    from xnrs_b200 import synthetic as syn
    from xnrs_b200.data import TitleStore
    from xnrs_b200.evaluation import CatalogueEvaluator
    from xnrs_b200.models import make_model
    K.call = EMU.call
    fx = load_npz('model_' + name)
    cfg = dict(fixture_cfg(fx), device='cpu')
    model = make_model(cfg)
    model.load_state_dict({k: torch.tensor(v) for k, v in sub(fx, 'sd').items()})
    model.eval()
    cat = syn.make_catalogue(61, cfg['seq_len'], vocab=200, dim=cfg['d_backbone'], seed=5)      # 62 rows: odd shard sizes
    imp = syn.make_eval_impressions(61, 37, cfg['hist_len'], n_users=cfg['n_users'], seed=6)
    store = TitleStore(cat.token_table, cat.title_tokens)
    ev = CatalogueEvaluator(model, store, cat.category, cat.subcategory, None, news_chunk=9, impression_chunk=5)
    ev.binary_metrics = True
    return ev, imp


def _eval_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    torch.set_num_threads(1)
    ev, imp = _eval_setup()
    out = ev.evaluate(imp, return_per_impression=True)
    np.savez(os.path.join(out_dir, f'e{rank}.npz'), vecs=ev.news_vecs.numpy(), shard=np.array([ev.last_shard['impressions'], ev.last_shard['candidates']]),
             means=np.array([out[k] for k in ('auc', 'rr', 'ndcg@5', 'ndcg@10', 'ctr@1', 'ctr@10', 'acc', 'rec', 'prec')]),
             n=out['impressions'], conf=np.array(out['conf']), per=out['per_impression'].numpy(), scores=out['scores'].numpy())
    dist.destroy_process_group()


def test_two_rank_sharded_evaluation_equals_single_process(tmp_path):
    """§8(e) row 2: catalogue rows sharded + all-gathered, impressions sharded by candidate count, metric sums all-reduced
    — every rank reports the single-process epoch means; the shards partition the impressions; news vectors identical."""
    world = 2
    mp.spawn(_eval_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    ev, imp = _eval_setup()
    want = ev.evaluate(imp, return_per_impression=True)
    r0, r1 = (np.load(tmp_path / f'e{r}.npz') for r in range(world))
    # sharded encode + all-gather == one pass (chunk boundaries differ, so GEMM blocking may: fp32 summation-order noise only)
    np.testing.assert_allclose(r0['vecs'], ev.news_vecs.numpy(), rtol=0, atol=1e-6 * np.abs(r0['vecs']).max())
    np.testing.assert_array_equal(r0['vecs'], r1['vecs'])                        # every rank holds the same gathered table
    assert int(r0['shard'][0] + r1['shard'][0]) == 37                            # the shards partition the impressions ...
    assert int(r0['shard'][1] + r1['shard'][1]) == int(imp['offsets'][-1])       # ... and the candidates,
    assert abs(int(r0['shard'][1]) - int(r1['shard'][1])) <= 80                  # balanced by candidate count
    np.testing.assert_allclose(np.concatenate([r0['scores'], r1['scores']]), want['scores'].numpy(), rtol=0,
                               atol=1e-5 * np.abs(want['scores'].numpy()).max())
    np.testing.assert_allclose(np.concatenate([r0['per'], r1['per']]), want['per_impression'].numpy(), rtol=0, atol=1e-9)
    want_means = np.array([want[k] for k in ('auc', 'rr', 'ndcg@5', 'ndcg@10', 'ctr@1', 'ctr@10', 'acc', 'rec', 'prec')])
    for r in (r0, r1):
        np.testing.assert_allclose(r['means'], want_means, rtol=0, atol=1e-12)   # sums of the same doubles, two partial sums
        assert int(r['n']) == want['impressions']
        np.testing.assert_array_equal(r['conf'], np.array(want['conf']))
