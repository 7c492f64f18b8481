"""The thin `train(cfg_path)` driver (reference train.py:20-74, training.py:118-189) on a tiny MIND-format data set written
to disk: reference-style YAML config read unchanged, news pickle + behaviours CSV through mind_io, two training epochs,
full-catalogue evaluation, reference-format checkpoints that load back.  Kernels are emulated (host-logic test)."""
import os

import numpy as np
import pytest
import torch

import _kernel_emulator as EMU
from xnrs_b200 import kernels as K

CONFIG = """
# data:
dataset: mind
train_news_data_path: {news}
train_user_data_path: {train_csv}
test_news_data_path: {news}
test_user_data_path: {test_csv}
min_hist_len: 1
# model:
model: '{model}'
scoring: 'dot'
text_features: ['title_emb']
catg_features: []
user_features: []
add_features: []
title_emb_dim: 16
total_emb_dim: 16
d_backbone: 16
n_heads: 4
hist_len: 4
seq_len: 5
p_dropout: 0.
bias: False
# training:
num_workers: 0
n_negatives: 2
batch_size: 4
shuffle_data: True
n_epochs: 2
test_freq: 1
ckpt_freq: 1
device: 'cpu'
lr: 0.001
random_seed: 0
random_seed: 3
contrastive_temperature: 0.08
contrastive_lambda: 0.01
# logging:
wandb: False
name: tiny_run
dir: {out}
"""


def _write_dataset(tmp_path, n_news=30, n_train=24, n_test=9, S=5, D=16):    # 16 / 4 heads: a head width the kernels support
    import pandas as pd
    rng = np.random.default_rng(0)
    ids = [f'N{i}' for i in range(n_news)]
    rows = {}
    for nid in ids:
        ln = int(rng.integers(1, S + 1))
        mask = (np.arange(S) < ln).astype(np.int64)[None]
        emb = (rng.normal(size=(1, S, D)).astype(np.float32)) * mask[..., None]
        rows[nid] = {'title_emb': (emb, mask), 'category_index': int(rng.integers(1, 5))}
    pd.DataFrame.from_dict(rows, orient='index').to_pickle(tmp_path / 'news.pkl')

    def behaviours(n, path):
        lines = ['user_index,history,impression,main_theme,main_category']
        for u in range(n):
            hist = ' '.join(rng.choice(ids, size=int(rng.integers(1, 7))))
            pos = [f'{x}-1' for x in rng.choice(ids, size=int(rng.integers(1, 3)))]
            neg = [f'{x}-0' for x in rng.choice(ids, size=int(rng.integers(2, 6)))]
            lines.append(f'{u + 1},{hist},{" ".join(pos + neg)},{["sports", "news", "finance"][u % 3]},x')
        path.write_text('\n'.join(lines) + '\n')

    behaviours(n_train, tmp_path / 'train.csv')
    behaviours(n_test, tmp_path / 'test.csv')


@pytest.fixture(params=['emulated', pytest.param('cuda', marks=pytest.mark.gpu)])
def device(request, monkeypatch):
    if request.param == 'emulated':
        monkeypatch.setattr(K, 'call', EMU.call)
        return 'cpu'
    return 'cuda:0'


@pytest.mark.parametrize('model', ['standard', 'NRMS'])
def test_train_driver_runs_epochs_evaluates_and_checkpoints(model, device, tmp_path):
    from xnrs_b200 import train as T
    _write_dataset(tmp_path)
    cfg_path = tmp_path / 'cfg.yml'
    cfg_path.write_text(CONFIG.format(news=tmp_path / 'news.pkl', train_csv=tmp_path / 'train.csv',
                                      test_csv=tmp_path / 'test.csv', out=tmp_path / 'runs', model=model)
                        .replace("device: 'cpu'", f"device: '{device}'"))
    hist = T.train(str(cfg_path))
    assert len(hist['train_loss']) == 2 and all(np.isfinite(hist['train_loss']))
    assert len(hist['test']) == 2 and hist['test'][-1]['impressions'] == 9
    assert all(0.0 <= hist['test'][-1][k] <= 1.0 for k in ('auc', 'rr', 'ndcg@5', 'ndcg@10'))
    last = hist['test'][-1]
    assert 0.0 <= last['acc'] <= 1.0 and sum(map(sum, last['conf'])) > 0
    pred = torch.load(last['predictions'], weights_only=False)       # training.py:85-95 dump format
    assert os.path.basename(last['predictions']) == 'predictions_1' and set(pred) == {'targets', 'scores', 'stats'}
    assert pred['scores'].shape == pred['targets'].shape and set(pred['stats']) == {'auc', 'mrr', 'ndcg@5', 'ndcg@10'}
    assert pred['stats']['auc'].shape == (9,)
    assert [os.path.basename(p) for p in hist['checkpoints']] == ['ckpt_0', 'ckpt_1']
    ck = torch.load(hist['checkpoints'][-1], weights_only=False)
    assert set(ck) == {'config', 'model_name', 'state_dict'} and ck['model_name'] == 'tiny_run'
    assert ck['config']['random_seed'] == 3                        # duplicated YAML key: the last value wins, like the reference
    loaded, _ = T.load_model_from_ckpt(hist['checkpoints'][-1], device=device)
    for (k, a), (_, b) in zip(hist['model'].state_dict().items(), loaded.state_dict().items()):
        assert torch.equal(a.cpu(), b.cpu()), k
    # debug mode: one step, one impression, one epoch (training.py:125-127,138-140,156-158)
    dbg = T.train(str(cfg_path), debug=True)
    assert len(dbg['train_loss']) == 1 and dbg['test'][-1]['impressions'] == 1
