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


U_GRANULE, T_GRANULE = 128, 1024      # the distinct-article / real-token counts are rounded up to these (padding is harmless)
_pinned_counts = []                   # small ring of pinned int32[2] buffers for the asynchronous count read-back


def _pinned_pair() -> torch.Tensor:
    if len(_pinned_counts) < 8:
        _pinned_counts.append(torch.empty(2, dtype=torch.int32).pin_memory())
        return _pinned_counts[-1]
    _pinned_counts.append(_pinned_counts.pop(0))
    return _pinned_counts[-1]


@dataclass
class TitlePlan:
    """index plumbing of one encoder pass over a set of news ids — a pure function of the ids (no model state), computed on
    the device by xnrs_plan_dedup / xnrs_plan_ragged with NO host round trip: the two data-dependent counts (distinct
    articles U, real tokens T) are copied to pinned host memory asynchronously and only read in `acquire()`, i.e. when the
    encoder needs shapes.  Computed ahead of the step on a side stream (`prefetch`) that read is free.  The arrays are padded
    past U / T with harmless entries (article 0, token 0 = zero rows, empty groups) and cut at counts rounded up to
    U_GRANULE / T_GRANULE, so steps fall into a few shape buckets (CUDA-graph replay) at < 1 % extra work.
      uniq (U',) distinct news ids | inv (n,) int32 slot -> row of uniq (None: no de-duplication)
      ragged: rows (T',) token-table rows of the real tokens, seg (U'+1,) group offsets, mask None
      fixed:  rows (U'*S,), mask (U'*S,) fp32, seg None
      cm (U',) fp32 collapsed title mask; event: recorded on the producing stream after the last plan kernel"""
    uniq: torch.Tensor
    inv: 'torch.Tensor | None'
    rows: torch.Tensor
    seg: 'torch.Tensor | None'
    mask: 'torch.Tensor | None'
    cm: torch.Tensor
    ragged: bool
    dedup: bool
    tix: 'torch.Tensor | None' = None  # ragged: (T',) int32 title of each token row, -1 on the padding rows
    event: 'torch.cuda.Event | None' = None
    seq_len: int = 0
    n_titles: int = -1                 # U after acquire() (exact count), n_rows: T
    n_rows: int = -1
    _counts: 'tuple | None' = None     # (device counts, pinned host copy or None, event or None) until acquire()

    def tensors(self):
        return [t for t in (self.uniq, self.inv, self.rows, self.seg, self.mask, self.cm, self.tix) if t is not None]

    def acquire(self):
        """make the plan usable on the current stream and cut the padded arrays at the (rounded) counts"""
        if self.event is not None and self.uniq.is_cuda:
            cur = torch.cuda.current_stream()
            cur.wait_event(self.event)
            for t in self.tensors():
                t.record_stream(cur)
            self.event = None
        if self._counts is not None:
            dev_counts, host, ev = self._counts
            self._counts = None
            if host is not None:
                ev.synchronize()               # long complete when the plan was prefetched
                U, T = (int(v) for v in host.tolist())
            else:
                U, T = (int(v) for v in dev_counts.tolist())
            n = self.uniq.numel()
            if self.dedup:
                self.n_titles = U
                u_cap = min(n, -(-max(U, 1) // U_GRANULE) * U_GRANULE)
            else:
                self.n_titles = u_cap = n
            self.uniq, self.cm = self.uniq[:u_cap], self.cm[:u_cap]
            if self.ragged:
                self.n_rows = T
                t_cap = min(self.rows.numel(), -(-max(T, 1) // T_GRANULE) * T_GRANULE)
                self.rows, self.seg, self.tix = self.rows[:t_cap], self.seg[:u_cap + 1], self.tix[:t_cap]
            else:
                self.n_rows = u_cap * self.seq_len
                self.rows, self.mask = self.rows[:u_cap * self.seq_len], self.mask[:u_cap * self.seq_len]
        return self


def plan_titles(store: TitleStore, ids_flat: torch.Tensor, dedup: bool, ragged: bool) -> TitlePlan:
    """ids_flat (n,) int32 on the store's device -> TitlePlan (device kernels only; the counts are read in acquire())"""
    dev = ids_flat.device
    ids_flat = ids_flat.to(torch.int32).contiguous()
    n, S = ids_flat.numel(), store.seq_len
    n_news = store.title_tokens.shape[0]
    i32 = dict(device=dev, dtype=torch.int32)
    counts = torch.zeros(2, **i32)
    if dedup:
        work, uniq, inv = torch.empty(2 * ((n_news + 31) // 32), **i32), torch.empty(n, **i32), torch.empty(n, **i32)
        K.call('xnrs_plan_dedup', ids_flat, n, n_news, work, uniq, inv, counts)
    else:
        uniq, inv = ids_flat, None
    if ragged:
        rows_cap = n * S + T_GRANULE
        lens, seg, rows = torch.empty(n, **i32), torch.empty(n + 1, **i32), torch.empty(rows_cap, **i32)
        tix = torch.empty(rows_cap, **i32)
        cm = torch.empty(n, device=dev, dtype=torch.float32)
        K.call('xnrs_plan_ragged', store.title_tokens, n_news, S, uniq, n, counts if dedup else None, T_GRANULE, lens, seg,
               rows, tix, rows_cap, cm, counts)
        mask = None
    else:
        rows, mask = K.expand_titles(store.title_tokens, uniq)               # over the capacity: padded titles are article 0
        cm = K.collapse_mask(mask, n, S)
        seg = tix = None
    plan = TitlePlan(uniq, inv, rows, seg, mask, cm, ragged, dedup, tix=tix, seq_len=S)
    if not dedup and not ragged:
        plan._counts = None
        plan.n_titles, plan.n_rows = n, n * S
    elif dev.type == 'cuda':
        host = _pinned_pair()
        host.copy_(counts, non_blocking=True)
        plan._counts = (counts, host, torch.cuda.current_stream().record_event())
    else:
        plan._counts = (counts, None, None)
    return plan


@dataclass
class IndexedTitles:
    """what a batch carries instead of (x, m): int32 news ids (b, n) into a TitleStore."""
    store: TitleStore
    news_ids: torch.Tensor
    plan: 'TitlePlan | None' = None          # optional pre-computed plumbing for these ids (see TitlePlan)
    distinct: bool = False                   # caller's promise that no id repeats (catalogue slices): nothing to de-duplicate

    def to(self, device):
        return IndexedTitles(self.store, self.news_ids.to(device, non_blocking=True), None, self.distinct)
