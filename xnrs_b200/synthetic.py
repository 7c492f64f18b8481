"""Synthetic MIND-shaped inputs: a frozen token table, a catalogue of token-id titles with
ragged lengths, Zipf-popular click histories, 1:K negative sampling with replacement (train) and CSR
impressions with 1-3 positives and 4-70 negatives (eval).  Host-side numpy only; no datasets are read."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Optional

import numpy as np
import torch

THEMES = 6            # make_mind_dataset.py:60-82 maps categories to six themes


@dataclass
class Catalogue:
    token_table: torch.Tensor        # (V, D) fp32, row 0 = zeros (pad token)
    title_tokens: torch.Tensor       # (N_news + 1, S) int32, row 0 = pad article
    abstract_tokens: Optional[torch.Tensor]
    category: torch.Tensor           # (N_news + 1,) int32, 0 = pad label
    subcategory: torch.Tensor


def make_catalogue(n_news: int, seq_len: int, vocab: int = 100_000, dim: int = 768, seed: int = 0,
                   with_abstract: bool = False, n_categories: int = 19, n_subcategories: int = 264) -> Catalogue:
    rng = np.random.default_rng(seed)
    g = torch.Generator().manual_seed(seed)
    table = torch.randn(vocab, dim, generator=g)
    table[0] = 0

    def titles():
        ids = rng.integers(1, vocab, size=(n_news + 1, seq_len), dtype=np.int64)
        lens = rng.integers(min(5, seq_len), seq_len + 1, size=n_news + 1)
        ids[np.arange(seq_len)[None, :] >= lens[:, None]] = 0
        ids[0] = 0
        return torch.from_numpy(ids.astype(np.int32))

    cat = rng.integers(1, n_categories + 1, size=n_news + 1).astype(np.int32)
    sub = rng.integers(1, n_subcategories + 1, size=n_news + 1).astype(np.int32)
    cat[0] = sub[0] = 0
    return Catalogue(table, titles(), titles() if with_abstract else None, torch.from_numpy(cat), torch.from_numpy(sub))


def zipf_news(rng, n_news: int, size, a: float = 1.1) -> np.ndarray:
    """news ids in [1, n_news] with a Zipf(a) popularity skew."""
    ranks = np.arange(1, n_news + 1, dtype=np.float64)
    p = ranks ** (-a)
    p /= p.sum()
    return rng.choice(n_news, size=size, p=p).astype(np.int64) + 1


def make_train_batch(n_news: int, batch: int, hist_len: int, n_neg: int = 4, n_users: int = 703_789,
                     seed: int = 0) -> Dict[str, torch.Tensor]:
    """index form of one training batch: history ids are front aligned and zero padded (dataset.py:77-85)."""
    rng = np.random.default_rng(seed)
    hist = zipf_news(rng, n_news, (batch, hist_len))
    n_valid = np.minimum(rng.integers(1, 2 * hist_len + 1, size=batch), hist_len)
    hist[np.arange(hist_len)[None, :] >= n_valid[:, None]] = 0
    cand = zipf_news(rng, n_news, (batch, 1 + n_neg))
    targets = np.zeros((batch, 1 + n_neg, 1), dtype=np.float32)
    targets[:, 0] = 1
    return {
        'hist_ids': torch.from_numpy(hist.astype(np.int32)),
        'cand_ids': torch.from_numpy(cand.astype(np.int32)),
        'targets': torch.from_numpy(targets),
        'user_index': torch.from_numpy(rng.integers(1, n_users + 1, size=(batch, 1)).astype(np.int32)),
        'main_theme': torch.from_numpy(rng.integers(0, THEMES, size=batch).astype(np.int32)),
    }


def make_eval_impressions(n_news: int, n_imp: int, hist_len: int, n_users: int = 703_789, seed: int = 1):
    """CSR impressions: positives first, then negatives (dataset.py:149)."""
    rng = np.random.default_rng(seed)
    hist = zipf_news(rng, n_news, (n_imp, hist_len))
    n_valid = np.minimum(rng.integers(1, 2 * hist_len + 1, size=n_imp), hist_len)
    hist[np.arange(hist_len)[None, :] >= n_valid[:, None]] = 0
    n_pos = rng.integers(1, 4, size=n_imp)
    n_neg = rng.integers(4, 71, size=n_imp)
    sizes = n_pos + n_neg
    offsets = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    cand = zipf_news(rng, n_news, int(offsets[-1]))
    targets = np.zeros(int(offsets[-1]), dtype=np.float32)
    for i in range(n_imp):
        targets[offsets[i]:offsets[i] + n_pos[i]] = 1
    return {
        'hist_ids': torch.from_numpy(hist.astype(np.int32)),
        'cand_ids': torch.from_numpy(cand.astype(np.int32)),
        'offsets': torch.from_numpy(offsets),
        'targets': torch.from_numpy(targets),
        'user_index': torch.from_numpy(rng.integers(1, n_users + 1, size=(n_imp, 1)).astype(np.int32)),
    }


def index_batch(store, cat: Catalogue, raw: Dict[str, torch.Tensor], device, abstract_store=None, categories: bool = True) -> dict:
    """reference batch-dict schema carrying IndexedTitles instead of dense (x, m) pairs.
    categories=False leaves out the category / subcategory index lookups (8 small device ops per batch that only NAML reads)."""
    hist_ids = raw['hist_ids'].to(device, non_blocking=True)
    cand_ids = raw['cand_ids'].to(device, non_blocking=True)
    hist = {'title_emb': store.index(hist_ids)}
    cand = {'title_emb': store.index(cand_ids)}
    if abstract_store is not None:
        hist['abstract_emb'] = abstract_store.index(hist_ids)
        cand['abstract_emb'] = abstract_store.index(cand_ids)
    if categories:
        key = ('_dev', str(device))
        if getattr(cat, '_dev_cache', None) is None or cat._dev_cache[0] != key:
            cat._dev_cache = (key, cat.category.to(device), cat.subcategory.to(device))
        _, category, subcategory = cat._dev_cache
        hist['category_index'] = category[hist_ids.long()]
        cand['category_index'] = category[cand_ids.long()]
        hist['subcategory_index'] = subcategory[hist_ids.long()]
        cand['subcategory_index'] = subcategory[cand_ids.long()]
    return {'user_features': {'history': hist, 'other': {'user_index': raw['user_index'].to(device, non_blocking=True)}},
            'candidate_features': cand, 'targets': raw['targets'].to(device, non_blocking=True),
            'main_theme': raw['main_theme'].to(device, non_blocking=True)}


def dense_batch(cat: Catalogue, raw: Dict[str, torch.Tensor], with_abstract: bool = False) -> dict:
    """the same batch in the reference's dense host format: x = table[ids] (for the oracle / CPU baseline)."""
    def text(tokens, ids):
        tok = tokens[ids.long()]
        return cat.token_table[tok.long()], (tok != 0).float().unsqueeze(-1)

    hist = {'title_emb': text(cat.title_tokens, raw['hist_ids'])}
    cand = {'title_emb': text(cat.title_tokens, raw['cand_ids'])}
    if with_abstract:
        hist['abstract_emb'] = text(cat.abstract_tokens, raw['hist_ids'])
        cand['abstract_emb'] = text(cat.abstract_tokens, raw['cand_ids'])
    hist['category_index'] = cat.category[raw['hist_ids'].long()]
    cand['category_index'] = cat.category[raw['cand_ids'].long()]
    hist['subcategory_index'] = cat.subcategory[raw['hist_ids'].long()]
    cand['subcategory_index'] = cat.subcategory[raw['cand_ids'].long()]
    return {'user_features': {'history': hist, 'other': {'user_index': raw['user_index']}},
            'candidate_features': cand, 'targets': raw['targets'], 'main_theme': raw['main_theme']}
