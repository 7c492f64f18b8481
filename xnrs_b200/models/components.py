"""Building blocks with the reference's interfaces (xnrs/models/components/{layers,news_encoding,
user_encoding,scoring,parent}.py) whose forward/backward run in the xnrs_b200 CUDA kernels.

The nn.Linear / nn.Sequential / nn.Embedding / nn.GRU children are used purely as *parameter containers*
so that ``state_dict()`` keys and shapes equal the reference's; their own forward methods
are never called.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn

from .. import kernels as K
from ..data import IndexedTitles, plan_titles


def _dev(module: nn.Module) -> torch.device:
    return next(module.parameters()).device


def _flat_rows(x: torch.Tensor):
    """(R, L, F) dense -> ((R*L, F) contiguous, R, L)"""
    R, L, F_ = x.shape
    return K._f32(x).reshape(R * L, F_), R, L


def _flat_mask(m: Optional[torch.Tensor], n: int):
    if m is None:
        return None
    m = K._f32(m).reshape(-1)
    if m.numel() != n:
        raise RuntimeError(f'mask has {m.numel()} entries, expected {n}')
    return m


def _seed() -> int:
    return int(torch.randint(0, 2 ** 62, (1,)).item())


class MaskedMean(nn.Module):
    """layers.MaskedMean (layers.py:19-37): sum(x * m) / (sum(m) + 1e-8) over the sequence axis."""

    def pool(self, x2: torch.Tensor, rows, mask, R: int, L: int, seg=None):
        """pooler interface of TextEncoder / UserEncoder: (R*L, F) rows (or table + row index) -> (pooled (R, F), None)"""
        if seg is not None:
            raise RuntimeError('MaskedMean pools fixed-length groups')
        if rows is not None:
            x2 = K.gather_rows(x2, rows)
        if mask is None:
            mask = torch.ones(R * L, device=x2.device, dtype=torch.float32)
        return K.MeanPoolFn.apply(x2, mask, R, L), None

    def forward(self, x: torch.Tensor, m: torch.Tensor):
        x2, R, L = _flat_rows(x)
        return self.pool(x2, None, _flat_mask(m, R * L), R, L)[0].unsqueeze(1)


class AdditiveAttention(nn.Module):
    """layers.AdditiveAttention (layers.py:40-69): same constructor, parameters and forward contract."""

    def __init__(self, in_features, hidden_features):
        super().__init__()
        self.fc1 = nn.Linear(in_features, hidden_features)
        self.fc2 = nn.Linear(hidden_features, 1)

    def pool(self, x2: torch.Tensor, rows, mask, R: int, L: int, seg=None, tix=None):
        """(R*L, F) rows (or table + row index) -> pooled (R, F), weights (R, L); seg: ragged group offsets (R+1);
        tix: group of each row (TitlePlan) — lets large ragged problems take the one-launch fused forward"""
        return K.AdditivePoolFn.apply(x2, rows, mask, self.fc1.weight, self.fc1.bias,
                                      self.fc2.weight, self.fc2.bias, R, L, seg, tix)

    def forward(self, x: torch.Tensor, m: torch.Tensor = None, return_weights: bool = False):
        x2, R, L = _flat_rows(x)
        pooled, attn = self.pool(x2, None, _flat_mask(m, R * L), R, L)
        pooled = pooled.unsqueeze(1)
        return (pooled, attn.unsqueeze(-1)) if return_weights else pooled


class PersonalizedAttention(nn.Module):
    """layers.PersonalizedAttention (layers.py:72-102)."""

    def __init__(self, in_features, hidden_features, query_features):
        super().__init__()
        self.x_fc = nn.Linear(in_features, hidden_features)
        self.q_fc = nn.Linear(query_features, hidden_features)

    def pool(self, q2, x2, rows, mask, R: int, L: int, rows_per_query: int = 1, seg=None):
        return K.PersonalizedPoolFn.apply(q2, x2, rows, mask, self.x_fc.weight, self.x_fc.bias,
                                          self.q_fc.weight, self.q_fc.bias, R, L, rows_per_query, seg)

    def forward(self, q: torch.Tensor, x: torch.Tensor, m: torch.Tensor = None):
        x2, R, L = _flat_rows(x)
        q2 = K._f32(q).reshape(R, -1)
        return self.pool(q2, x2, None, _flat_mask(m, R * L), R, L).unsqueeze(1)


class MultiHeadAttention(nn.Module):
    """layers.MultiHeadAttention (layers.py:105-156), including its query-axis mask and p=0.1 dropout default.

    Train-mode dropout uses a counter-based generator seeded from torch's RNG (bit parity with torch's
    Philox stream is not a goal — SURVEY §7 hard part 3); ``keep_mask`` (R,h,L,L) overrides it for tests."""

    def __init__(self, n_heads, d_model, dropout=0.1, scaled=True):
        super().__init__()
        self.scaled = scaled
        self.d_model = d_model
        self.d_k = d_model // n_heads
        self.h = n_heads
        self.q_linear = nn.Linear(d_model, d_model)
        self.v_linear = nn.Linear(d_model, d_model)
        self.k_linear = nn.Linear(d_model, d_model)
        self.dropout = nn.Dropout(dropout)
        self.out = nn.Linear(d_model, d_model)
        self.keep_mask: Optional[torch.Tensor] = None

    def attend(self, x2: torch.Tensor, rows, mask, R: int, L: int) -> torch.Tensor:
        p = float(self.dropout.p) if (self.training or self.keep_mask is not None) else 0.0
        keep = None if self.keep_mask is None else K._f32(self.keep_mask).reshape(-1)
        seed = _seed() if (p > 0 and keep is None) else 0
        return K.MultiHeadAttentionFn.apply(
            x2, rows, mask, self.q_linear.weight, self.q_linear.bias, self.k_linear.weight, self.k_linear.bias,
            self.v_linear.weight, self.v_linear.bias, self.out.weight, self.out.bias, R, L, self.h, keep, p, seed,
            1.0 if self.scaled else float(self.d_k) ** 0.5)

    def forward(self, x: torch.Tensor, m: torch.Tensor):
        x2, R, L = _flat_rows(x)
        return self.attend(x2, None, _flat_mask(m, R * L), R, L).view(R, L, self.d_model)


def ragged_token_rows(store, news_ids: torch.Tensor):
    """index plumbing for the padding-free path without de-duplication: -> (token rows of the real tokens (T',) int32 padded
    with the zero row, group offsets (n+1,) int32, collapsed title mask (n,) fp32).  Device kernels; one count read-back."""
    plan = plan_titles(store, news_ids.reshape(-1), False, True).acquire()
    return plan.rows, plan.seg, plan.cm


def _apply_head(head: nn.Sequential, x2: torch.Tensor) -> torch.Tensor:
    if not isinstance(head[1], nn.ReLU):
        raise NotImplementedError('only the reference default activation nn.ReLU() is implemented')
    return K.Mlp2Fn.apply(x2, head[0].weight, head[0].bias, head[2].weight, head[2].bias)


def _input_dropout(module: nn.Module, x: torch.Tensor) -> torch.Tensor:
    p = float(module.dropout.p)
    if p > 0 and module.training:
        return K.DropoutFn.apply(K._f32(x), None, p, _seed())
    return x


class TextEncoder(nn.Module):
    """news_encoding.TextEncoder (news_encoding.py:8-60).

    ``forward`` accepts the reference's dense ``(x[b,n,s,d], m[b,n,s,1])`` pair (tuple or list) or an
    ``IndexedTitles`` (int32 news ids into a device-resident TitleStore); in the indexed case the token
    rows are gathered inside the kernels and the dense (b,n,s,d) tensor never exists.

    ``dedup_titles``: every title is encoded independently of its batch position (news_encoding.py:48-57), so with
    ids the encoder runs ONCE per distinct article of the batch and the vectors are gathered back to the (b,n)
    slots; the backward pass sums the slot gradients per article (scatter-add) before the encoder's backward.
    Same values and gradients (up to fp32 summation order), ~6x fewer titles on MIND-shaped (Zipf) batches.

    ``skip_padding``: pad tokens have mask 0, i.e. pooling weight exactly 0 (layers.py:62-64), so without self-attention
    they cannot influence anything: only real tokens are gathered and the pooler works on ragged groups.  (With
    self-attention padded tokens ARE attended to as keys — layers.py:142-144 — so NRMS keeps the fixed-length path.)"""

    def __init__(self, pooler: nn.Module, p_dropout: float, out_features: int, in_features: Optional[int] = 768,
                 head: bool = True, activation: nn.Module = nn.ReLU(), att: Optional[nn.Module] = None,
                 bias: bool = True):
        super().__init__()
        self.dedup_titles = True        # per-instance options of the index fast path (see the class docstring);
        self.skip_padding = True        # `encoder_options(model, ...)` sets them on every TextEncoder of a model
        self.dummy_param = nn.Parameter(torch.zeros(1))
        self.dropout = nn.Dropout(p=p_dropout)
        self.att = att
        self.pooler = pooler
        if head:
            assert in_features is not None, 'in_features is required if head is True'
            self.head = nn.Sequential(nn.Linear(in_features, out_features, bias=bias), activation,
                                      nn.Linear(out_features, out_features, bias=bias))
        self.out_dim = out_features

    def _encode(self, x2, rows, mask, R: int, S: int, seg=None, tix=None):
        if self.att is not None:
            x2, rows = self.att.attend(x2, rows, mask, R, S), None
        if tix is not None and isinstance(self.pooler, AdditiveAttention):
            pooled, _ = self.pooler.pool(x2, rows, mask, R, S, seg, tix)
        else:
            pooled, _ = self.pooler.pool(x2, rows, mask, R, S, seg)
        return _apply_head(self.head, pooled) if hasattr(self, 'head') else pooled

    def plan_kind(self, n_slots: int, distinct: bool = False):
        """(dedup, ragged) of the plumbing this encoder wants for n_slots title slots (distinct: the ids do not repeat)"""
        return (self.dedup_titles and not distinct and n_slots >= 64,
                self.skip_padding and self.att is None and hasattr(self.pooler, 'fc1'))

    def encode_unique(self, inpt: IndexedTitles):
        """index batches: -> (e_u (U,E) vectors of the DISTINCT articles of the batch, inv (b*n,) int32 slot -> row of e_u or
        None when nothing was de-duplicated, cm_u (U,) collapsed title mask).  The id plumbing comes from ``inpt.plan`` when a
        matching plan was prefetched, else it is computed here."""
        device = _dev(self)
        store = inpt.store
        b, n = inpt.news_ids.shape
        S = store.seq_len
        if self.dropout.p > 0 and self.training:
            raise NotImplementedError('input dropout on gathered titles (every shipped config has p_dropout 0)')
        dedup, ragged = self.plan_kind(b * n, inpt.distinct)
        plan = inpt.plan
        if plan is None or plan.dedup != dedup or plan.ragged != ragged:
            plan = plan_titles(store, inpt.news_ids.to(device).reshape(-1), dedup, ragged)
        plan.acquire()
        nu = plan.uniq.numel()
        e_u = self._encode(store.token_table, plan.rows, plan.mask, nu, S, plan.seg, plan.tix)
        return e_u, plan.inv, plan.cm

    def forward(self, inpt):
        device = _dev(self)
        if isinstance(inpt, IndexedTitles):
            b, n = inpt.news_ids.shape
            e_u, inv, cm_u = self.encode_unique(inpt)
            if inv is None:
                return e_u.view(b, n, self.out_dim), cm_u.view(b, n, 1)
            e = K.EmbeddingFn.apply(e_u, inv, None)                                # gather back; bwd = scatter-add
            return e.view(b, n, self.out_dim), cm_u[inv].view(b, n, 1)
        else:
            x, m = inpt
            x, m = x.to(device), m.to(device)
            b, n, S, d = x.shape
            x = _input_dropout(self, x)
            mask = _flat_mask(m, b * n * S)
            e = self._encode(K._f32(x).reshape(b * n * S, d), None, mask, b * n, S)
        cm = K.collapse_mask(mask, b * n, S)
        return e.view(b, n, self.out_dim), cm.view(b, n, 1)


def encoder_options(model: nn.Module, **options) -> None:
    """set index-fast-path options (dedup_titles, skip_padding) on every TextEncoder inside `model` (instance attributes:
    two models in one process never see each other's settings)"""
    for m in model.modules():
        if isinstance(m, TextEncoder):
            for k, v in options.items():
                if not hasattr(m, k):
                    raise AttributeError(f'TextEncoder has no option {k!r}')
                setattr(m, k, bool(v))


class UserEncoder(nn.Module):
    """user_encoding.UserEncoder (user_encoding.py:6-81); ``out_dim`` is accepted and ignored like the reference."""

    def __init__(self, pooler: nn.Module, p_dropout: float, emb_dim: Optional[int] = None,
                 out_dim: Optional[int] = None, att: Optional[nn.Module] = None, head: bool = False,
                 activation: nn.Module = nn.ReLU(), bias: bool = True):
        super().__init__()
        self.dummy_param = nn.Parameter(torch.zeros(1))
        self.dropout = nn.Dropout(p=p_dropout)
        self.att = att
        self.pooler = pooler
        if head:
            assert emb_dim is not None
            self.head = nn.Sequential(nn.Linear(emb_dim, emb_dim, bias=bias), activation,
                                      nn.Linear(emb_dim, emb_dim, bias=bias))

    def can_pool_items(self) -> bool:
        """history given as ids into a table of distinct item vectors: possible when the encoder is [no dropout, no
        attention] -> additive pooler (-> head), i.e. StandardRec / CL"""
        return (self.att is None and isinstance(self.pooler, AdditiveAttention)
                and not (self.dropout.p > 0 and self.training))

    def forward_items(self, items: torch.Tensor, item_mask: torch.Tensor, ids: torch.Tensor) -> torch.Tensor:
        """== forward((items[ids], item_mask[ids])) with the pooler's fc1 evaluated once per item (ItemLogitPoolFn)"""
        p = self.pooler
        pooled, _ = K.ItemLogitPoolFn.apply(items, item_mask, ids, p.fc1.weight, p.fc1.bias, p.fc2.weight, p.fc2.bias)
        if hasattr(self, 'head'):
            pooled = _apply_head(self.head, pooled)
        return pooled.unsqueeze(1)

    def forward(self, inpt, add_features: Optional[dict] = None, return_weights: bool = False):
        x, m = inpt
        device = _dev(self)
        x, m = x.to(device), m.to(device)
        x = _input_dropout(self, x)
        x2, R, L = _flat_rows(x)
        mask = _flat_mask(m, R * L)
        if self.att is not None:
            x2 = self.att.attend(x2, None, mask, R, L)
        a = None
        pooled, a = self.pooler.pool(x2, None, mask, R, L)
        if return_weights and a is None:
            raise RuntimeError('this pooler has no attention weights to return')
        if hasattr(self, 'head'):
            pooled = _apply_head(self.head, pooled)
        u = pooled.unsqueeze(1)
        return (u, a.unsqueeze(-1)) if return_weights else u


def _normalized(u: torch.Tensor, c: torch.Tensor):
    """u / ||u||, c / ||c|| along the embedding axis (scoring.py:20-22)"""
    B, N, T = c.shape
    return (K.NormalizeRowsFn.apply(K._f32(u).reshape(B, T)).view(B, 1, T),
            K.NormalizeRowsFn.apply(K._f32(c).reshape(B * N, T)).view(B, N, T))


class DotScoring(nn.Module):
    """scoring.DotScoring (scoring.py:6-23): u (B,1,D), c (B,N,D) -> (B,N,1)."""

    def __init__(self, normalize: bool = False):
        super().__init__()
        self.normalize = normalize

    def forward(self, u: torch.Tensor, c: torch.Tensor):
        if self.normalize:
            u, c = _normalized(u, c)
        B, N, T = c.shape
        return K.DotScoreFn.apply(K._f32(u).reshape(B, T), K._f32(c)).unsqueeze(-1)


class BilinScoring(nn.Module):
    """scoring.BilinScoring (scoring.py:41-69): nn.Bilinear(D, D, 1) of (u repeated over the candidates, c).
    s[b,n] = u_b^T W c_bn + bias = <u_b, W c_bn> + bias: one GEMM over the candidates, then the dot-score kernel."""

    def __init__(self, emb_dim: int, normalize: bool = False, bias: bool = True):
        super().__init__()
        self.bilin = nn.Bilinear(in1_features=emb_dim, in2_features=emb_dim, out_features=1, bias=bias)
        self.normalize = normalize

    def forward(self, u: torch.Tensor, c: torch.Tensor):
        if self.normalize:
            u, c = _normalized(u, c)
        B, N, T = c.shape
        wc = K.LinearFn.apply(K._f32(c).reshape(B * N, T), None, self.bilin.weight[0], None)        # rows W c_bn
        s = K.DotScoreFn.apply(K._f32(u).reshape(B, T), wc.view(B, N, T))
        if self.bilin.bias is not None:
            s = K.AddScalarFn.apply(s, self.bilin.bias)
        return s.unsqueeze(-1)


class FCScoring(nn.Module):
    """scoring.FCScoring (scoring.py:72-102): fc2(tanh(fc1([u, c]))) over the concatenated pair."""

    def __init__(self, emb_dim: int, hidden_dim: int, activation=torch.tanh, bias: bool = True):
        super().__init__()
        if activation is not torch.tanh:
            raise NotImplementedError('FCScoring: only the reference default activation torch.tanh is implemented')
        self.fc1 = nn.Linear(in_features=2 * emb_dim, out_features=hidden_dim, bias=bias)
        self.fc2 = nn.Linear(in_features=hidden_dim, out_features=1, bias=bias)
        self.activation = activation

    def forward(self, u: torch.Tensor, c: torch.Tensor):
        B, N, T = c.shape
        x = torch.cat([K._f32(u).reshape(B, 1, T).expand(B, N, T), K._f32(c)], dim=2).reshape(B * N, 2 * T)   # byte movement
        h = K.LinearTanhFn.apply(x, self.fc1.weight, self.fc1.bias)
        return K.LinearFn.apply(h, None, self.fc2.weight, self.fc2.bias).view(B, N, 1)


def _merge_ids(history, candidates):
    """[all history slots | all candidate slots] as one IndexedTitles of shape (1, b*nh + b*nc) -> (titles, b, nh, nc)"""
    dev = history.store.device
    b, nh = history.news_ids.shape
    nc = candidates.news_ids.shape[1]
    ids = torch.cat([history.news_ids.to(dev).reshape(1, -1), candidates.news_ids.to(dev).reshape(1, -1)], dim=1)
    return IndexedTitles(history.store, ids), b, nh, nc


def merge_sides(history, candidates):
    """history and candidate titles go through the SAME encoder (parent.py:31-32): with index batches they are encoded in
    one call (one de-duplication, half the launches).  The ids are laid out [all history slots | all candidate slots] so
    that both sides come back as contiguous row blocks (no strided slicing / copies on either pass).
    Returns (merged IndexedTitles of shape (1, b*nh + b*nc), b, nh, nc) or None."""
    if (isinstance(history, IndexedTitles) and isinstance(candidates, IndexedTitles)
            and history.store is candidates.store and history.news_ids.shape[0] == candidates.news_ids.shape[0]):
        cached = getattr(history, '_merged', None)             # set by prefetch_titles for exactly this pair;
        if cached is not None and cached[0] is candidates:     # consumed once: a reused batch object is planned afresh
            history._merged = None
            return cached[1]
        return _merge_ids(history, candidates)
    return None


_prefetch_streams = {}


def prefetch_titles(encoder, history, candidates, after=None) -> bool:
    """compute the merged-side id plumbing of an upcoming (history, candidates) pair on a side stream (see TitlePlan): call
    it right after enqueuing the current step.  `after`: CUDA event the ids become valid at (e.g. their H2D copy).
    Returns False when the pair is not index-based.  (A helper-thread variant was measured too: the GIL hand-offs cost as
    much as the ~0.3 ms of stream-local waiting they removed from the enqueuing thread, so the plan is computed in line.)"""
    if not (isinstance(history, IndexedTitles) and isinstance(candidates, IndexedTitles) and isinstance(encoder, TextEncoder)
            and history.store is candidates.store and history.news_ids.shape[0] == candidates.news_ids.shape[0]):
        return False
    history._merged = None
    dev = history.store.device

    def job():
        merged = _merge_ids(history, candidates)               # never through the cache of merge_sides
        titles = merged[0]
        titles.plan = plan_titles(titles.store, titles.news_ids.reshape(-1), *encoder.plan_kind(titles.news_ids.numel()))
        return merged

    if dev.type != 'cuda':
        history._merged = (candidates, job())
        return True
    side = _prefetch_streams.get(dev)
    if side is None:
        # HIGH priority: the step's persistent tensor-core kernels hold every SM's registers, so the (tiny) plan kernels can only
        # start at kernel boundaries of the running step; with the default priority they were served when the step had drained
        # (host profile: the next step's enqueue waited 0.7 ms per step for the plan's counts), now at the next boundary
        side = _prefetch_streams[dev] = torch.cuda.Stream(device=dev, priority=-1)
    if after is not None:
        side.wait_event(after)
    with torch.cuda.stream(side):
        merged = job()
        merged[0].plan.event = side.record_event()
    history._merged = (candidates, merged)
    return True


def encode_both_sides(encoder, history, candidates):
    """-> (h (b,nh,E), hm (b,nh,1), c (b,nc,E))"""
    merged = merge_sides(history, candidates)
    if merged is None:
        h, hm = encoder(history)
        c, _ = encoder(candidates)
        return h, hm, c
    titles, b, nh, nc = merged
    e, m = encoder(titles)
    e, m = e[0], m[0]
    return e[:b * nh].view(b, nh, -1), m[:b * nh].view(b, nh, 1), e[b * nh:].view(b, nc, -1)


class ParentRec(nn.Module):
    """parent.ParentRec (parent.py:8-81): news_encoder x2 -> user_encoder -> rec_model."""

    def __init__(self, news_encoder: nn.Module, user_encoder: nn.Module, rec_model: nn.Module,
                 text_feature: str = 'title_emb'):
        super().__init__()
        self.news_encoder = news_encoder
        self.user_encoder = user_encoder
        self.rec_model = rec_model
        self.text_feature = text_feature
        # index batches + additive user pooler: pool the history straight from the batch's DISTINCT article vectors (the
        # pooler's fc1 runs once per article, the (b,H,E) history tensor and its gradient never exist).  Off: gather first.
        self.item_logits = True

    def _forward(self, history, candidates, add_user_feats=None, return_embeddings: bool = False):
        merged = merge_sides(history, candidates) if self.item_logits else None
        if (merged is not None and isinstance(self.news_encoder, TextEncoder) and isinstance(self.user_encoder, UserEncoder)
                and self.user_encoder.can_pool_items()):
            titles, b, nh, nc = merged
            e_u, inv, cm_u = self.news_encoder.encode_unique(titles)
            if inv is None:
                inv = torch.arange(b * (nh + nc), device=e_u.device, dtype=torch.int32)
            inv = K._i32(inv)
            u = self.user_encoder.forward_items(e_u, cm_u, inv[:b * nh].view(b, nh))
            c = K.EmbeddingFn.apply(e_u, inv[b * nh:], None).view(b, nc, -1)
            r = self.rec_model(u, c)
            return (r, u, c) if return_embeddings else r
        h, hm, c = encode_both_sides(self.news_encoder, history, candidates)
        u = self.user_encoder((h, hm), add_user_feats)
        r = self.rec_model(u, c)
        return (r, u, c) if return_embeddings else r

    def prefetch(self, batch: dict, after=None) -> bool:
        """index batches: compute the id plumbing of an UPCOMING batch on a side stream (overlaps the running step)"""
        return prefetch_titles(self.news_encoder, batch['user_features']['history'][self.text_feature],
                               batch['candidate_features'][self.text_feature], after)

    def forward(self, batch: dict, return_embeddings: bool = False):
        return self._forward(history=batch['user_features']['history'][self.text_feature],
                             candidates=batch['candidate_features'][self.text_feature],
                             add_user_feats=batch['user_features'].get('other'),
                             return_embeddings=return_embeddings)

    def get_user_embeddings(self, batch: dict) -> torch.Tensor:
        history = batch['user_features']['history'][self.text_feature]
        if isinstance(history, list) and len(history) == 2:
            history = tuple(history)
        news_emb, news_mask = self.news_encoder(history)
        return self.user_encoder((news_emb, news_mask)).squeeze(1)
