"""Row G / §8(f) row 2: xnrs_b200.mind_io (MIND on-disk formats -> token table + int32 index batches) against the
reference's NewsRecDataset + custom_collate_fn (xnrs/data/dataset.py:48-163, xnrs/utils.py:190-204).  The fixture
tests/golden/dataset.npz was produced by running the unmodified reference (tests/golden/make_golden.py:dataset_fixture);
densifying the index batch (x = token_table[token_rows[news_ids]]) must reproduce its tensors BIT-EXACTLY."""
import json
import os
import random

import numpy as np
import torch

from xnrs_b200 import mind_io

HERE = os.path.dirname(os.path.abspath(__file__))


def _load():
    z = np.load(os.path.join(HERE, 'golden', 'dataset.npz'), allow_pickle=False)
    ids = [str(x) for x in z['ids']]
    news_feat = {nid: {'title_emb': (z['emb'][i][None], z['mask'][i][None]), 'category_index': int(z['cat'][i])}
                 for i, nid in enumerate(ids)}
    sessions = json.loads(str(z['sessions']))
    S, D, H, K = (int(v) for v in z['dims'])
    tables = mind_io.NewsTables.from_news_dict(news_feat, ['title_emb'], ['category_index'])
    return z, tables, sessions, (S, D, H, K)


def _dense(tables, ids):
    rows = tables.tokens['title_emb'][ids.long()]                       # (..., S) token-row ids, 0 = padded position
    return tables.token_table[rows.long()], (rows != 0).float().unsqueeze(-1)


def test_table_layout():
    z, tables, _, (S, D, _, _) = _load()
    n = len(z['ids'])
    assert tables.token_table.shape == (int(z['mask'].sum()) + 1, D) and float(tables.token_table[0].abs().sum()) == 0
    x, m = _dense(tables, torch.arange(1, n + 1))
    assert torch.equal(x, torch.from_numpy(z['emb'])) and torch.equal(m[..., 0], torch.from_numpy(z['mask']).float())
    x0, m0 = _dense(tables, torch.zeros(1, dtype=torch.long))           # article 0 = the pad article
    assert float(x0.abs().sum()) == 0 and float(m0.sum()) == 0


def test_train_batch_matches_reference_dataset():
    z, tables, sessions, (S, D, H, K) = _load()
    b = mind_io.train_batch(sessions, tables, range(len(sessions)), H, K, rng=random.Random(77))
    hx, hm = _dense(tables, b['hist_ids'])
    cx, cm = _dense(tables, b['cand_ids'])
    assert torch.equal(hx, torch.from_numpy(z['train/hist_x'])) and torch.equal(hm, torch.from_numpy(z['train/hist_m']))
    assert torch.equal(cx, torch.from_numpy(z['train/cand_x'])) and torch.equal(cm, torch.from_numpy(z['train/cand_m']))
    assert torch.equal(b['targets'], torch.from_numpy(z['train/targets']))
    assert torch.equal(b['user_index'], torch.from_numpy(z['train/user_index']).to(torch.int32))
    hist_cat = mind_io.categorical_history(tables, 'category_index', b['hist_ids'])
    assert torch.equal(hist_cat.long(), torch.from_numpy(z['train/hist_cat']).long())
    assert torch.equal(tables.categorical['category_index'][b['cand_ids'].long()].long(),
                       torch.from_numpy(z['train/cand_cat']).long())
    inv = {v: k for k, v in tables.news_index.items()}
    assert [[inv[int(i)] for i in row] for row in b['cand_ids']] == json.loads(str(z['train/item_ids']))
    # theme labels: equal strings <-> equal ids (only equality matters to the contrastive loss, training.py:414-417)
    th = [s['main_theme'] for s in sessions]
    for i in range(len(th)):
        for j in range(len(th)):
            assert (th[i] == th[j]) == bool(b['main_theme'][i] == b['main_theme'][j])


def test_eval_csr_matches_reference_dataset():
    z, tables, sessions, (S, D, H, K) = _load()
    imp = mind_io.eval_impressions(sessions, tables, H)
    assert imp['offsets'].tolist() == list(np.cumsum([0] + [len(s['positives']) + len(s['negatives']) for s in sessions]))
    for i in range(len(sessions)):
        a, b = int(imp['offsets'][i]), int(imp['offsets'][i + 1])
        hx, hm = _dense(tables, imp['hist_ids'][i])
        cx, cm = _dense(tables, imp['cand_ids'][a:b])
        assert torch.equal(hx, torch.from_numpy(z[f'eval/{i}/hist_x'])) and torch.equal(hm, torch.from_numpy(z[f'eval/{i}/hist_m']))
        assert torch.equal(cx, torch.from_numpy(z[f'eval/{i}/cand_x'])) and torch.equal(cm, torch.from_numpy(z[f'eval/{i}/cand_m']))
        assert torch.equal(imp['targets'][a:b], torch.from_numpy(z[f'eval/{i}/targets'])[:, 0])
        hist_cat = mind_io.categorical_history(tables, 'category_index', imp['hist_ids'][i])
        assert torch.equal(hist_cat.long(), torch.from_numpy(z[f'eval/{i}/hist_cat']).long())


def test_behaviors_readers(tmp_path):
    assert mind_io.split_impression('N7-1 N9-0 N11-0 N3-1') == (['N7', 'N3'], ['N9', 'N11'])
    p = tmp_path / 'behaviors.tsv'
    p.write_text('1\tU1\t11/11/2019\tN1 N2 N3\tN4-1 N5-0\n2\tU2\t11/11/2019\t\tN4-0 N6-1\n')
    s = mind_io.read_behaviors(str(p))
    assert len(s) == 1 and s[0]['history'] == ['N1', 'N2', 'N3'] and s[0]['positives'] == ['N4'] and s[0]['negatives'] == ['N5']
    q = tmp_path / 'behaviors.csv'
    q.write_text('user_index,history,impression,main_theme\n5,N1 N2,N3-0 N4-1,sports\n')
    s = mind_io.read_behaviors(str(q))
    assert s[0]['user_index'] == 5 and s[0]['main_theme'] == 'sports' and s[0]['positives'] == ['N4']
