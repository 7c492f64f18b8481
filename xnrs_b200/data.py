"""Device-resident form of the reference's per-sample host gather (row G).

The reference looks every news id up in a python dict of precomputed (S, D) token-embedding blocks and
concatenates them on the host (xnrs/data/dataset.py:63-65,77-85,97-109), then ships ~4.6 MB per impression
over PCIe inside TextEncoder.forward (news_encoding.py:45-47).  Here the frozen token table and the
catalogue's token ids live in HBM; a batch carries int32 news ids and the encoders gather rows inside
their kernels.  News id 0 / token id 0 are the pad article / pad token (zero row, zero mask), which
reproduces the zero-embedding, zero-mask history padding of dataset.py:82-85.
"""
from __future__ import annotations

from dataclasses import dataclass

import torch

from . import kernels as K


class TitleStore:
    """frozen token table (V, D) fp32 + catalogue token ids (N_news, S) int32, both on one CUDA device."""

    def __init__(self, token_table: torch.Tensor, title_tokens: torch.Tensor):
        if token_table.dtype != torch.float32 or token_table.dim() != 2 or token_table.shape[1] % 4:
            raise ValueError('token_table must be fp32 (V, D) with D % 4 == 0')
        if title_tokens.dim() != 2:
            raise ValueError('title_tokens must be (N_news, S)')
        self.token_table = token_table.contiguous()
        self.title_tokens = title_tokens.to(torch.int32).contiguous()
        if self.token_table.device != self.title_tokens.device:
            raise ValueError('token_table and title_tokens must share a device')

    @property
    def device(self):
        return self.token_table.device

    @property
    def seq_len(self) -> int:
        return self.title_tokens.shape[1]

    @property
    def dim(self) -> int:
        return self.token_table.shape[1]

    def to(self, device):
        return TitleStore(self.token_table.to(device), self.title_tokens.to(device))

    def index(self, news_ids: torch.Tensor) -> 'IndexedTitles':
        return IndexedTitles(self, news_ids)

    def dense(self, news_ids: torch.Tensor):
        """materialise the reference-format pair (x (..,S,D), m (..,S,1)) — for parity tests and explain-style
        callers that need the dense tensors; the encoders never call this."""
        rows, mask = K.expand_titles(self.title_tokens, news_ids.to(self.device))
        x = K.gather_rows(self.token_table, rows)
        shape = tuple(news_ids.shape)
        return x.view(*shape, self.seq_len, self.dim), mask.view(*shape, self.seq_len, 1)


@dataclass
class TitlePlan:
    """index plumbing of one encoder pass over a set of news ids — a pure function of the ids (no model state), so it can
    be computed ahead of the step on a side stream (`prefetch`), where its two host syncs (sizes of the distinct-article
    and real-token sets) overlap the previous step instead of idling the GPU.
      uniq (U,) distinct news ids | inv (n,) int32 slot -> row of uniq (None: no de-duplication)
      ragged: rows (tokens,) token-table rows of the real tokens, seg (U+1,) group offsets, mask None
      fixed:  rows (U*S,), mask (U*S,) fp32, seg None
      cm (U,) fp32 collapsed title mask; event: recorded on the producing stream after the last plan kernel"""
    uniq: torch.Tensor
    inv: 'torch.Tensor | None'
    rows: torch.Tensor
    seg: 'torch.Tensor | None'
    mask: 'torch.Tensor | None'
    cm: torch.Tensor
    ragged: bool
    dedup: bool
    event: 'torch.cuda.Event | None' = None

    def tensors(self):
        return [t for t in (self.uniq, self.inv, self.rows, self.seg, self.mask, self.cm) if t is not None]

    def acquire(self):
        """make the plan usable on the current stream (no-op for plans computed in line)"""
        if self.event is not None and self.uniq.is_cuda:
            cur = torch.cuda.current_stream()
            cur.wait_event(self.event)
            for t in self.tensors():
                t.record_stream(cur)
            self.event = None
        return self


def plan_titles(store: TitleStore, ids_flat: torch.Tensor, dedup: bool, ragged: bool) -> TitlePlan:
    """ids_flat (n,) int32 on the store's device -> TitlePlan (torch index ops + xnrs_expand_titles; two host syncs)"""
    uniq, inv = ids_flat, None
    if dedup:
        uniq, inv = torch.unique(ids_flat, return_inverse=True)             # id plumbing (host sync: the distinct count)
        inv = inv.to(torch.int32)
    if ragged:
        tok = store.title_tokens[uniq.long()]
        valid = tok != 0
        lens = valid.sum(1, dtype=torch.int32)
        seg = torch.zeros(lens.numel() + 1, device=tok.device, dtype=torch.int32)
        torch.cumsum(lens, 0, out=seg[1:])
        rows = tok[valid].contiguous()                                      # host sync: the real-token count
        return TitlePlan(uniq, inv, rows, seg, None, (lens > 0).to(torch.float32), True, dedup)
    rows, mask = K.expand_titles(store.title_tokens, uniq)
    cm = K.collapse_mask(mask, uniq.numel(), store.seq_len)
    return TitlePlan(uniq, inv, rows, None, mask, cm, False, dedup)


@dataclass
class IndexedTitles:
    """what a batch carries instead of (x, m): int32 news ids (b, n) into a TitleStore."""
    store: TitleStore
    news_ids: torch.Tensor
    plan: 'TitlePlan | None' = None          # optional pre-computed plumbing for these ids (see TitlePlan)
    distinct: bool = False                   # caller's promise that no id repeats (catalogue slices): nothing to de-duplicate

    def to(self, device):
        return IndexedTitles(self.store, self.news_ids.to(device, non_blocking=True), None, self.distinct)
