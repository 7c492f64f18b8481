"""MIND on-disk formats -> device tables and index batches.

The reference keeps every article's precomputed token embeddings in a pickled DataFrame, turns it into a python dict
(`xnrs/data/mind.py:161-164`), reads behaviours from a CSV with `history` / `impression` columns
(`mind.py:186-199`: "N1 N2 ..." and "N7-1 N9-0 ...") and assembles dense tensors per sample on the host
(`xnrs/data/dataset.py:48-163`).  Here the same inputs become (a) one token table + per-article token-row ids for the
device (`TitleStore`) and (b) int32 index batches — the semantics of `NewsRecDataset.__getitem__` are kept exactly:
history = LAST `hist_len` clicks, front aligned, zero padded; candidates = positives then negatives; train samples one
positive and `n_negatives` negatives WITH replacement; categorical history features padded with label 0.
Host-side only (pandas / numpy); nothing here touches the GPU.
"""
from __future__ import annotations

import random
from dataclasses import dataclass, field
from typing import Dict, Iterable, List, Optional, Sequence

import numpy as np
import torch


def split_impression(impression: str):
    """'N7-1 N9-0' -> (['N7'], ['N9'])   (mind.py:194-197)"""
    pos, neg = [], []
    for tok in impression.split():
        (pos if tok[-1] == '1' else neg).append(tok[:-2])
    return pos, neg


def read_behaviors(path: str) -> List[dict]:
    """the reference's behaviours CSV (header row; columns history, impression and optionally user_index, main_theme,
    main_category) or a raw MIND `behaviors.tsv` (no header: index, user, time, history, impression).  Sessions with an
    empty history are dropped like `load_behaviors_as_hf_dataset` does (mind.py:190)."""
    import pandas as pd
    if path.endswith('.tsv'):
        df = pd.read_csv(path, sep='\t', header=None, names=['index', 'user', 'time', 'history', 'impression'])
    else:
        df = pd.read_csv(path)
    sessions = []
    for row in df.to_dict('records'):
        hist = row.get('history')
        if not isinstance(hist, str) or not hist.strip():
            continue
        pos, neg = split_impression(str(row['impression']))
        s = {'history': hist.split(), 'positives': pos, 'negatives': neg}
        for k in ('user_index', 'main_theme', 'main_category', 'user'):
            if k in row and row[k] == row[k]:
                s[k] = row[k]
        sessions.append(s)
    return sessions


@dataclass
class NewsTables:
    """article id -> row number (1-based; 0 = pad article) plus the flat tables the kernels gather from."""
    news_index: Dict[str, int]
    token_table: torch.Tensor                      # (sum of all feature blocks' rows + 1, D) fp32, row 0 = zeros
    tokens: Dict[str, torch.Tensor]                # text feature -> (N+1, S) int32 token-row ids (0 = padded position)
    categorical: Dict[str, torch.Tensor] = field(default_factory=dict)   # feature -> (N+1,) int32, 0 for the pad article

    @classmethod
    def from_news_dict(cls, news_feat: Dict[str, dict], text_features: Sequence[str] = ('title_emb',),
                       catg_features: Sequence[str] = ()) -> 'NewsTables':
        """news_feat[id][feat] = (emb (1,S,D) float32, mask (1,S)) as stored by data/utils.py:103-111.  Every
        (article, position) with mask 1 gets its own table row: row = 1 + running count; masked positions map to row 0."""
        ids = list(news_feat.keys())
        index = {nid: i + 1 for i, nid in enumerate(ids)}
        blocks, tokens, offset = [], {}, 1
        for feat in text_features:
            first = news_feat[ids[0]][feat]
            S, D = np.asarray(first[0]).shape[-2:]
            emb = np.zeros((len(ids), S, D), dtype=np.float32)
            msk = np.zeros((len(ids), S), dtype=bool)
            for i, nid in enumerate(ids):
                e, m = news_feat[nid][feat]
                emb[i] = np.asarray(e, dtype=np.float32).reshape(S, D)
                msk[i] = np.asarray(m).reshape(S) != 0
            rows = np.zeros((len(ids) + 1, S), dtype=np.int32)
            n_real = int(msk.sum())
            rows[1:][msk] = np.arange(offset, offset + n_real, dtype=np.int32)
            blocks.append(emb[msk])
            tokens[feat] = torch.from_numpy(rows)
            offset += n_real
        D = blocks[0].shape[1]
        table = np.concatenate([np.zeros((1, D), dtype=np.float32)] + blocks)
        cats = {}
        for feat in catg_features:
            col = np.zeros(len(ids) + 1, dtype=np.int32)
            col[1:] = [int(news_feat[nid][feat]) for nid in ids]
            cats[feat] = torch.from_numpy(col)
        return cls(index, torch.from_numpy(table), tokens, cats)

    @classmethod
    def from_pickle(cls, path: str, text_features=('title_emb',), catg_features=()) -> 'NewsTables':
        """the pickled news DataFrame of make_mind_dataset.py (mind.py:161-164)"""
        import pandas as pd
        df = pd.read_pickle(path)
        return cls.from_news_dict(df[list(text_features) + list(catg_features)].to_dict('index'), text_features, catg_features)

    def store(self, feature: str = 'title_emb', device='cuda'):
        from .data import TitleStore
        return TitleStore(self.token_table.to(device), self.tokens[feature].to(device))

    def rows(self, news_ids: Iterable[str]) -> List[int]:
        return [self.news_index[n] for n in news_ids]


def _history_rows(tables: NewsTables, history: Sequence[str], hist_len: int) -> np.ndarray:
    out = np.zeros(hist_len, dtype=np.int32)
    last = tables.rows(history[-hist_len:])                 # last H clicks, front aligned (dataset.py:77-85)
    out[:len(last)] = last
    return out


def _theme_ids(sessions: Sequence[dict]) -> Dict[str, int]:
    return {t: i for i, t in enumerate(sorted({str(s.get('main_theme', '')) for s in sessions}))}


def train_batch(sessions: Sequence[dict], tables: NewsTables, indices: Sequence[int], hist_len: int, n_negatives: int,
                rng: Optional[random.Random] = None, themes: Optional[Dict[str, int]] = None) -> Dict[str, torch.Tensor]:
    """index form of `custom_collate_fn([dataset[i] for i in indices])` in train mode (dataset.py:54-58: one random
    positive, n_negatives negatives drawn with replacement).  Theme strings are numbered globally (sorted), so
    data-parallel ranks agree on the labels."""
    rng = rng or random
    themes = themes if themes is not None else _theme_ids(sessions)
    B = len(indices)
    hist = np.zeros((B, hist_len), dtype=np.int32)
    cand = np.zeros((B, 1 + n_negatives), dtype=np.int32)
    user = np.zeros((B, 1), dtype=np.int32)
    theme = np.zeros(B, dtype=np.int32)
    for b, i in enumerate(indices):
        s = sessions[i]
        hist[b] = _history_rows(tables, s['history'], hist_len)
        cand[b] = tables.rows([rng.choice(s['positives'])] + rng.choices(s['negatives'], k=n_negatives))
        user[b, 0] = int(s.get('user_index', 0))
        theme[b] = themes[str(s.get('main_theme', ''))]
    targets = np.zeros((B, 1 + n_negatives, 1), dtype=np.float32)
    targets[:, 0] = 1                                                       # dataset.py:147
    return {'hist_ids': torch.from_numpy(hist), 'cand_ids': torch.from_numpy(cand), 'targets': torch.from_numpy(targets),
            'user_index': torch.from_numpy(user), 'main_theme': torch.from_numpy(theme)}


def eval_impressions(sessions: Sequence[dict], tables: NewsTables, hist_len: int) -> Dict[str, torch.Tensor]:
    """CSR form of the eval dataset (dataset.py:59-61,149): all positives, then all negatives of every impression."""
    hist = np.zeros((len(sessions), hist_len), dtype=np.int32)
    cand, targets, offsets = [], [], [0]
    user = np.zeros((len(sessions), 1), dtype=np.int32)
    for i, s in enumerate(sessions):
        hist[i] = _history_rows(tables, s['history'], hist_len)
        rows = tables.rows(list(s['positives']) + list(s['negatives']))
        cand += rows
        targets += [1.0] * len(s['positives']) + [0.0] * len(s['negatives'])
        offsets.append(len(cand))
        user[i, 0] = int(s.get('user_index', 0))
    return {'hist_ids': torch.from_numpy(hist), 'cand_ids': torch.tensor(cand, dtype=torch.int32),
            'offsets': torch.tensor(offsets, dtype=torch.int64), 'targets': torch.tensor(targets, dtype=torch.float32),
            'user_index': torch.from_numpy(user)}


def categorical_history(tables: NewsTables, feature: str, hist_ids: torch.Tensor) -> torch.Tensor:
    """category ids of the history slots; padded slots read label 0 (utils.py:64-71, dataset.py:115-116)"""
    return tables.categorical[feature][hist_ids.long()]
