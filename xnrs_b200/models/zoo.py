"""The five target models and the factory (xnrs/models/full_models/{standard_model,nrms,naml,lstur,npa}.py,
xnrs/models/make_model.py) assembled from the kernel-backed components.

Constructor signatures ``Model(cfg, rec_model)``, forward contracts, attribute names and therefore
``state_dict()`` keys/shapes equal the reference's, so its checkpoints load unchanged.  torch.cat / view
below only move bytes (view concatenation for NAML / LSTUR); all arithmetic is in the CUDA kernels.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import kernels as K
from ..data import IndexedTitles
from .components import (AdditiveAttention, BilinScoring, DotScoring, FCScoring, MaskedMean, MultiHeadAttention, ParentRec,
                         PersonalizedAttention, TextEncoder, UserEncoder, _dev, _flat_mask, merge_sides)


class _Missing:
    """what an absent config key reads as: falsy and empty (the reference relies on DotMap's behaviour)."""

    def __bool__(self):
        return False

    def __iter__(self):
        return iter(())

    def __contains__(self, item):
        return False


class Cfg:
    """attribute view over a dict / DotMap / namespace config."""

    def __init__(self, cfg):
        self._c = cfg

    def get(self, key, default=None):
        c = self._c
        if isinstance(c, dict):
            if key not in c:
                return default
            v = c[key]
        else:
            v = getattr(c, key, default)
        if type(v).__name__ == 'DotMap' and len(v) == 0:     # DotMap materialises missing keys as empty maps
            return default
        return v

    def __getattr__(self, key):
        if key.startswith('_'):
            raise AttributeError(key)
        v = self.get(key, _Missing())
        return v

    def __contains__(self, key):
        return not isinstance(self.get(key, _Missing()), _Missing)


def _cfg(cfg) -> Cfg:
    return cfg if isinstance(cfg, Cfg) else Cfg(cfg)


def _embed(emb: nn.Embedding, idx: torch.Tensor) -> torch.Tensor:
    """(…) int ids -> (…, dim) through the gather kernel (dense weight gradient, padding_idx honoured)."""
    out = K.EmbeddingFn.apply(emb.weight, idx.to(emb.weight.device), emb.padding_idx)
    return out.view(*idx.shape, emb.embedding_dim)


def _linear(lin: nn.Linear, x: torch.Tensor) -> torch.Tensor:
    y = K.LinearFn.apply(K._f32(x).reshape(-1, x.shape[-1]), None, lin.weight, lin.bias)
    return y.view(*x.shape[:-1], lin.out_features)


class StandardRec(ParentRec):
    """the contrastive (CL) model, ``model: 'standard'`` (standard_model.py:6-37)."""

    def __init__(self, cfg, rec_model: nn.Module):
        cfg = _cfg(cfg)
        title_encoder = TextEncoder(att=None, pooler=AdditiveAttention(cfg.d_backbone, 256), p_dropout=cfg.p_dropout,
                                    in_features=cfg.d_backbone, out_features=cfg.title_emb_dim, bias=bool(cfg.bias))
        user_encoder = UserEncoder(pooler=AdditiveAttention(cfg.title_emb_dim, 256), att=None, head=True,
                                   p_dropout=cfg.p_dropout, emb_dim=cfg.title_emb_dim, bias=bool(cfg.bias))
        super().__init__(news_encoder=title_encoder, user_encoder=user_encoder, rec_model=rec_model)

    def get_news_embeddings(self, batch: dict, mode: str = 'history') -> torch.Tensor:
        """standard_model.py:73-100"""
        if mode == 'candidate':
            news_input = batch['candidate_features'][self.text_feature]
        elif mode == 'history':
            news_input = batch['user_features']['history'][self.text_feature]
        else:
            raise ValueError("mode must be 'candidate' or 'history'")
        if isinstance(news_input, list) and len(news_input) == 2:
            news_input = tuple(news_input)
        return self.news_encoder(news_input)[0]


class BaseRec(ParentRec):
    """``model: 'base'`` (base_model.py:5-36): StandardRec without the user-side head."""

    def __init__(self, cfg, rec_model: nn.Module):
        cfg = _cfg(cfg)
        title_encoder = TextEncoder(att=None, pooler=AdditiveAttention(cfg.d_backbone, 256), p_dropout=cfg.p_dropout,
                                    in_features=cfg.d_backbone, out_features=cfg.title_emb_dim, bias=bool(cfg.bias))
        user_encoder = UserEncoder(pooler=AdditiveAttention(cfg.title_emb_dim, 256), att=None, head=False,
                                   p_dropout=cfg.p_dropout, emb_dim=cfg.title_emb_dim)
        super().__init__(news_encoder=title_encoder, user_encoder=user_encoder, rec_model=rec_model)


class MeanRec(ParentRec):
    """``model: 'mean'`` (mean_model.py:5-31): masked-mean pooling at both levels, head on the news side only."""

    def __init__(self, cfg, rec_model: nn.Module):
        cfg = _cfg(cfg)
        title_encoder = TextEncoder(att=None, pooler=MaskedMean(), p_dropout=cfg.p_dropout, in_features=cfg.d_backbone,
                                    out_features=cfg.title_emb_dim, bias=bool(cfg.bias))
        user_encoder = UserEncoder(pooler=MaskedMean(), att=None, head=False, p_dropout=cfg.p_dropout,
                                   emb_dim=cfg.title_emb_dim, bias=bool(cfg.bias))
        super().__init__(news_encoder=title_encoder, user_encoder=user_encoder, rec_model=rec_model)


class ParamFreeRec(ParentRec):
    """param_free_model.py:5-31 — masked means only; its single trainable tensors are the encoders' dummy_params."""

    def __init__(self, cfg, rec_model: nn.Module):
        cfg = _cfg(cfg)
        assert cfg.title_emb_dim == cfg.d_backbone
        title_encoder = TextEncoder(att=None, head=False, pooler=MaskedMean(), p_dropout=cfg.p_dropout,
                                    out_features=cfg.d_backbone)
        user_encoder = UserEncoder(att=None, head=False, pooler=MaskedMean(), p_dropout=cfg.p_dropout,
                                   emb_dim=cfg.title_emb_dim)
        super().__init__(news_encoder=title_encoder, user_encoder=user_encoder, rec_model=rec_model)


class NRMS(ParentRec):
    """nrms.py:9-47 — heads are always biased (the reference does not forward cfg.bias here)."""

    def __init__(self, cfg, rec_model: nn.Module):
        cfg = _cfg(cfg)
        title_encoder = TextEncoder(att=MultiHeadAttention(cfg.n_heads, cfg.d_backbone),
                                    pooler=AdditiveAttention(cfg.d_backbone, 256), p_dropout=cfg.p_dropout,
                                    in_features=cfg.d_backbone, out_features=cfg.title_emb_dim)
        user_encoder = UserEncoder(att=MultiHeadAttention(cfg.n_heads, cfg.title_emb_dim),
                                   pooler=AdditiveAttention(cfg.title_emb_dim, 256), emb_dim=cfg.title_emb_dim,
                                   p_dropout=cfg.p_dropout, head=False)
        super().__init__(news_encoder=title_encoder, user_encoder=user_encoder, rec_model=rec_model)


class NRMS_LF(ParentRec):
    """nrms.py:49-79 — NRMS news encoder, masked-mean ("late fusion") user encoder."""

    def __init__(self, cfg, rec_model: nn.Module):
        cfg = _cfg(cfg)
        title_encoder = TextEncoder(att=MultiHeadAttention(cfg.n_heads, cfg.d_backbone),
                                    pooler=AdditiveAttention(cfg.d_backbone, 256), p_dropout=cfg.p_dropout,
                                    in_features=cfg.d_backbone, out_features=cfg.title_emb_dim)
        user_encoder = UserEncoder(att=None, pooler=MaskedMean(), emb_dim=cfg.title_emb_dim, p_dropout=cfg.p_dropout, head=False)
        super().__init__(news_encoder=title_encoder, user_encoder=user_encoder, rec_model=rec_model)


class SmallNAML(nn.Module):
    """naml.py:162-238 — NAML with two views (title, category); forward only returns scores (no CL hook exists for it)."""

    def __init__(self, cfg, rec_model):
        super().__init__()
        cfg = _cfg(cfg)
        self.title_encoder = TextEncoder(att=None, pooler=AdditiveAttention(cfg.d_backbone, 256), p_dropout=cfg.p_dropout,
                                         in_features=cfg.d_backbone, out_features=cfg.title_emb_dim)
        self.cat_embedder = nn.Embedding(cfg.n_categories + 1, cfg.cat_emb_dim)
        self.cat_fc = nn.Linear(cfg.cat_emb_dim, cfg.total_emb_dim)
        self.feature_pooler = AdditiveAttention(cfg.total_emb_dim, 256)
        self.user_encoder = AdditiveAttention(cfg.title_emb_dim, 256)
        self.rec_model = rec_model
        self.emb_dim = cfg.total_emb_dim

    def _news(self, title, ctg):
        t, mask = self.title_encoder(title)
        c = _linear(self.cat_fc, _embed(self.cat_embedder, ctg))
        b, n, e = t.shape
        views = torch.cat([t, c], dim=2).reshape(b * n * 2, e)                 # == stack(dim=2): byte movement only
        pooled, _ = self.feature_pooler.pool(views, None, None, b * n, 2)
        return pooled.view(b, n, e), mask

    def _forward(self, hist_title_features, hist_ctg, cand_title_features, cand_ctg):
        h, hm = self._news(hist_title_features, hist_ctg)
        c, _ = self._news(cand_title_features, cand_ctg)
        b, n, e = h.shape
        u, _ = self.user_encoder.pool(h.reshape(b * n, e), None, _flat_mask(hm, b * n), b, n)
        return self.rec_model(u.unsqueeze(1), c)

    def forward(self, batch: dict):
        return self._forward(hist_title_features=batch['user_features']['history']['title_emb'],
                             hist_ctg=batch['user_features']['history']['category_index'],
                             cand_title_features=batch['candidate_features']['title_emb'],
                             cand_ctg=batch['candidate_features']['category_index'])


class NAML(nn.Module):
    """naml.py:7-160 — title + abstract encoders, category / sub-category views, 4-view additive pooling."""

    def __init__(self, cfg, rec_model):
        super().__init__()
        cfg = _cfg(cfg)
        self.article_level = True   # index batches: run the 4-view news encoder once per distinct article (off: once per slot)
        self.title_encoder = TextEncoder(att=None, pooler=AdditiveAttention(cfg.d_backbone, 256),
                                         p_dropout=cfg.p_dropout, in_features=cfg.d_backbone,
                                         out_features=cfg.title_emb_dim)
        self.body_encoder = TextEncoder(att=None, pooler=AdditiveAttention(cfg.d_backbone, 256),
                                        p_dropout=cfg.p_dropout, in_features=cfg.d_backbone,
                                        out_features=cfg.title_emb_dim)
        self.cat_embedder = nn.Embedding(cfg.n_categories + 1, cfg.cat_emb_dim)
        self.cat_fc = nn.Linear(cfg.cat_emb_dim, cfg.total_emb_dim)
        self.subcat_embedder = nn.Embedding(cfg.n_subcategories + 1, cfg.sub_emb_dim)
        self.subcat_fc = nn.Linear(cfg.sub_emb_dim, cfg.total_emb_dim)
        self.feature_pooler = AdditiveAttention(cfg.total_emb_dim, 256)
        self.user_encoder = AdditiveAttention(cfg.title_emb_dim, 256)
        self.rec_model = rec_model
        self.emb_dim = cfg.total_emb_dim

    def _news(self, title, abstract, ctg, subctg):
        t, mask = self.title_encoder(title)
        a, _ = self.body_encoder(abstract)
        c = _linear(self.cat_fc, _embed(self.cat_embedder, ctg))
        s = _linear(self.subcat_fc, _embed(self.subcat_embedder, subctg))
        b, n, e = t.shape
        views = torch.cat([t, a, c, s], dim=2).reshape(b * n * 4, e)          # byte movement only
        pooled, _ = self.feature_pooler.pool(views, None, None, b * n, 4)
        return pooled.view(b, n, e), mask

    def _user(self, h, hm):
        b, n, e = h.shape
        pooled, _ = self.user_encoder.pool(h.reshape(b * n, e), None, _flat_mask(hm, b * n), b, n)
        return pooled.unsqueeze(1)

    def _forward(self, hist_title_features, hist_abstract_features, hist_ctg, hist_subctg,
                 cand_title_features, cand_abstract_features, cand_ctg, cand_subctg, return_embeddings=False):
        mt, ma = merge_sides(hist_title_features, cand_title_features), merge_sides(hist_abstract_features, cand_abstract_features)
        if mt is not None and ma is not None:       # index batches: both sides in one pass through the shared encoders
            dev = _dev(self)
            _, b, nh, nc = mt                        # ids laid out [all history slots | all candidate slots]
            flat = lambda hx, cx: torch.cat([hx.to(dev).reshape(1, -1), cx.to(dev).reshape(1, -1)], 1)
            ids = mt[0].news_ids.reshape(-1)
            if (self.article_level and ids.numel() >= 64 and torch.equal(ids, ma[0].news_ids.reshape(-1))
                    and not (self.title_encoder.dropout.p > 0 and self.training)):
                # Every view of a NAML news vector (title, abstract, category, sub-category: naml.py:82-107) is an attribute
                # of the ARTICLE, so the whole 4-view encoder runs once per distinct article of the batch (id plumbing: one
                # unique + the per-article category ids), and the user pooler works from those vectors (ItemLogitPoolFn).
                uniq, inv = torch.unique(ids, return_inverse=True)
                U = uniq.numel()
                per_article = lambda slot_vals: torch.empty(U, device=dev, dtype=slot_vals.dtype).scatter_(0, inv, slot_vals.reshape(-1))
                u_ids = uniq.view(1, U)
                # (the ids are distinct already: `distinct=True` skips the encoders' own unique / gather-back)
                e_u, cm_u = self._news(IndexedTitles(mt[0].store, u_ids, None, True), IndexedTitles(ma[0].store, u_ids, None, True),
                                       per_article(flat(hist_ctg, cand_ctg)).view(1, U),
                                       per_article(flat(hist_subctg, cand_subctg)).view(1, U))
                e_u, cm_u, inv = e_u[0], K._f32(cm_u.reshape(-1)), K._i32(inv)
                ue = self.user_encoder
                u, _ = K.ItemLogitPoolFn.apply(e_u, cm_u, inv[:b * nh].view(b, nh), ue.fc1.weight, ue.fc1.bias,
                                               ue.fc2.weight, ue.fc2.bias)
                u = u.unsqueeze(1)
                c = K.EmbeddingFn.apply(e_u, inv[b * nh:], None).view(b, nc, -1)
                r = self.rec_model(u, c)
                return (r, u, c) if return_embeddings else r
            e, m = self._news(mt[0], ma[0], flat(hist_ctg, cand_ctg), flat(hist_subctg, cand_subctg))
            e, m = e[0], m[0]
            h, hm, c = e[:b * nh].view(b, nh, -1), m[:b * nh].view(b, nh, 1), e[b * nh:].view(b, nc, -1)
        else:
            h, hm = self._news(hist_title_features, hist_abstract_features, hist_ctg, hist_subctg)
            c, _ = self._news(cand_title_features, cand_abstract_features, cand_ctg, cand_subctg)
        u = self._user(h, hm)
        r = self.rec_model(u, c)
        return (r, u, c) if return_embeddings else r

    def get_user_embeddings(self, batch: dict):
        """naml.py:113-147 — returned un-squeezed (B,1,E) like the reference."""
        hf = batch['user_features']['history']
        h, hm = self._news(hf['title_emb'], hf['abstract_emb'], hf['category_index'], hf['subcategory_index'])
        return self._user(h, hm)

    def forward(self, batch: dict, return_embeddings: bool = False):
        hf, cf = batch['user_features']['history'], batch['candidate_features']
        return self._forward(hf['title_emb'], hf['abstract_emb'], hf['category_index'], hf['subcategory_index'],
                             cf['title_emb'], cf['abstract_emb'], cf['category_index'], cf['subcategory_index'],
                             return_embeddings=return_embeddings)


class LSTURNewsEncoder(nn.Module):
    """lstur.py:162-207 — title encoder (pooler hidden = title_emb_dim) ⊕ category embedding [⊕ sub-category]."""

    def __init__(self, cfg):
        super().__init__()
        cfg = _cfg(cfg)
        self.cfg = cfg
        self.title_encoder = TextEncoder(pooler=AdditiveAttention(cfg.d_backbone, cfg.title_emb_dim),
                                         p_dropout=cfg.p_dropout, out_features=cfg.title_emb_dim,
                                         in_features=cfg.d_backbone, head=True, bias=bool(cfg.bias))
        self.cat_embedder = nn.Embedding(cfg.n_categories + 1, cfg.cat_emb_dim)
        if 'subcategory_index' in cfg.catg_features:
            self.subcat_embedder = nn.Embedding(cfg.n_subcategories + 1, cfg.cat_emb_dim)

    def forward(self, title_features, cat_idxs, subcat_idxs=None):
        title_emb, m = self.title_encoder(title_features)
        parts = [title_emb, _embed(self.cat_embedder, cat_idxs)]
        if subcat_idxs is not None:
            assert hasattr(self, 'subcat_embedder')
            parts.append(_embed(self.subcat_embedder, subcat_idxs))
        return torch.cat(parts, dim=2), m


class LSTURUserEncoder(nn.Module):
    """lstur.py:83-159 with ``long_term_method: embedding`` (the only combination that runs)."""

    def __init__(self, cfg):
        super().__init__()
        cfg = _cfg(cfg)
        self.cfg = cfg
        long_term_emb_dim = cfg.total_emb_dim
        if cfg.long_short_term_method == 'con':
            long_term_emb_dim //= 2
        if cfg.long_term_method == 'embedding':
            self.long_term_encoder = nn.Embedding(cfg.n_users + 1, long_term_emb_dim, padding_idx=0)
        elif cfg.long_term_method == 'mean':
            raise NotImplementedError("long_term_method 'mean' produces mismatched dimensions in the reference "
                                      '(user_encoding.py:26-34 ignores out_dim) and cannot run there either')
        else:
            raise ValueError(f'long_term_method must be in [mean, embedding], got {cfg.long_term_method}')
        self.dropout = nn.Dropout(p=cfg.p_user_dropout or 0.0)
        self.gru = nn.GRU(cfg.total_emb_dim, long_term_emb_dim, batch_first=True)
        self.keep_mask = None       # optional explicit user-dropout keep mask (tests)

    def forward(self, history_features, user_ids: torch.Tensor):
        h, hm = history_features
        device = _dev(self)
        u_lt = _embed(self.long_term_encoder, user_ids.to(device)).reshape(h.shape[0], -1)
        p = float(self.dropout.p)
        if p > 0 and (self.training or self.keep_mask is not None):
            seed = int(torch.randint(0, 2 ** 62, (1,)).item()) if self.keep_mask is None else 0
            u_lt = K.DropoutFn.apply(u_lt, self.keep_mask, p, seed)
        method = self.cfg.long_short_term_method
        if method == 'lt_only':
            return u_lt.unsqueeze(1)
        st = self.cfg.st_hist_len
        h_st = K._f32(h[:, :st])
        B, L, _ = h_st.shape
        lengths = torch.empty(B, device=device, dtype=torch.int32)
        K.call('xnrs_lengths_from_mask', K._f32(hm[:, :st]).reshape(-1), B, L, lengths)
        g = self.gru
        if method == 'ini':
            u = K.GruLastFn.apply(h_st, lengths, g.weight_ih_l0, g.weight_hh_l0, g.bias_ih_l0, g.bias_hh_l0, u_lt)
            return u.unsqueeze(1)
        if method == 'con':
            u_st = K.GruLastFn.apply(h_st, lengths, g.weight_ih_l0, g.weight_hh_l0, g.bias_ih_l0, g.bias_hh_l0, None)
            return torch.cat((u_st, u_lt), dim=1).unsqueeze(1)
        raise ValueError(f'invalid value for long_short_term_method, got {method}')


class LSTUR(nn.Module):
    """lstur.py:9-79"""

    def __init__(self, cfg, rec_model):
        super().__init__()
        self.news_encoder = LSTURNewsEncoder(cfg)
        self.user_encoder = LSTURUserEncoder(cfg)
        self.rec_model = rec_model
        self.cfg = _cfg(cfg)

    def _subcats(self, batch):
        if 'subcategory_index' in self.cfg.catg_features:
            return (batch['user_features']['history']['subcategory_index'],
                    batch['candidate_features']['subcategory_index'])
        return None, None

    def forward(self, batch: dict, return_embeddings: bool = False):
        hs, cs = self._subcats(batch)
        hf, cf = batch['user_features']['history'], batch['candidate_features']
        mt = merge_sides(hf['title_emb'], cf['title_emb'])
        if mt is not None:
            dev = _dev(self)
            _, b, nh, nc = mt                        # ids laid out [all history slots | all candidate slots]
            flat = lambda hx, cx: torch.cat([hx.to(dev).reshape(1, -1), cx.to(dev).reshape(1, -1)], 1)
            sub = None if hs is None else flat(hs, cs)
            e, m = self.news_encoder(mt[0], flat(hf['category_index'], cf['category_index']), sub)
            e, m = e[0], m[0]
            h, hm, c = e[:b * nh].view(b, nh, -1), m[:b * nh].view(b, nh, 1), e[b * nh:].view(b, nc, -1)
        else:
            h, hm = self.news_encoder(hf['title_emb'], hf['category_index'], hs)
            c, _ = self.news_encoder(cf['title_emb'], cf['category_index'], cs)
        u = self.user_encoder((h, hm), batch['user_features']['other']['user_index'])
        r = self.rec_model(u, c)
        return (r, u, c) if return_embeddings else r

    def get_user_embeddings(self, batch: dict) -> torch.Tensor:
        hs, _ = self._subcats(batch)
        hf = batch['user_features']['history']
        h, hm = self.news_encoder(hf['title_emb'], hf['category_index'], hs)
        return self.user_encoder((h, hm), batch['user_features']['other']['user_index']).squeeze(1)


class NPA(nn.Module):
    """npa.py:8-96 — personalised attention at word and news level; news vectors depend on the user."""

    def __init__(self, cfg, rec_model):
        super().__init__()
        cfg = _cfg(cfg)
        self.user_embedder = nn.Embedding(cfg.n_users + 1, cfg.user_emb_dim)
        self.title_pooler = PersonalizedAttention(cfg.d_backbone, 128, cfg.user_emb_dim)
        self.dropout = nn.Dropout(p=cfg.p_dropout)
        self.news_head = nn.Sequential(nn.Linear(cfg.d_backbone, cfg.title_emb_dim), nn.ReLU(),
                                       nn.Linear(cfg.title_emb_dim, cfg.title_emb_dim))
        self.user_encoder = PersonalizedAttention(cfg.title_emb_dim, 128, cfg.user_emb_dim)
        self.rec_model = rec_model
        self.skip_padding = True    # index batches: only real tokens reach the GEMM / pooler (ragged groups)

    def _titles(self, feats, ue2):
        """per-title personalised pooling + head: (b,n,s,d) or IndexedTitles -> (b,n,E), token mask (b*n*s)"""
        from ..data import IndexedTitles
        device = _dev(self)
        if isinstance(feats, IndexedTitles):
            ids = feats.news_ids.to(device)
            b, n = ids.shape
            s = feats.store.seq_len
            from .components import ragged_token_rows
            x2 = feats.store.token_table
            if self.skip_padding:                   # only real tokens reach the GEMM / pooler (ragged groups)
                rows, seg, cm = ragged_token_rows(feats.store, ids.reshape(-1))
                pooled = self.title_pooler.pool(ue2, x2, rows, None, b * n, s, rows_per_query=n, seg=seg)
                hd = self.news_head
                e = K.Mlp2Fn.apply(pooled, hd[0].weight, hd[0].bias, hd[2].weight, hd[2].bias)
                return e.view(b, n, -1), cm, (b, n, s)
            rows, mask = K.expand_titles(feats.store.title_tokens, ids)
        else:
            x, m = feats
            x, m = x.to(device), m.to(device)
            b, n, s, d = x.shape
            if self.dropout.p > 0 and self.training:
                x = K.DropoutFn.apply(K._f32(x), None, float(self.dropout.p), int(torch.randint(0, 2 ** 62, (1,))))
            x2, rows, mask = K._f32(x).reshape(b * n * s, d), None, _flat_mask(m, b * n * s)
        pooled = self.title_pooler.pool(ue2, x2, rows, mask, b * n, s, rows_per_query=n)
        hd = self.news_head
        e = K.Mlp2Fn.apply(pooled, hd[0].weight, hd[0].bias, hd[2].weight, hd[2].bias)
        return e.view(b, n, -1), mask, (b, n, s)

    def _forward(self, hist_title_features, cand_title_features, uid: torch.Tensor):
        device = _dev(self)
        ue2 = _embed(self.user_embedder, uid.to(device)).reshape(uid.shape[0], -1)      # (B, du)
        h, hmask, (b, nh, s) = self._titles(hist_title_features, ue2)
        hm = hmask if hmask.numel() == b * nh else K.collapse_mask(hmask, b * nh, s)     # ragged path returns it collapsed
        u = self.user_encoder.pool(ue2, h.reshape(b * nh, -1), None, hm, b, nh).unsqueeze(1)
        c, _, _ = self._titles(cand_title_features, ue2)
        return self.rec_model(u, c)

    def forward(self, batch: dict):
        return self._forward(batch['user_features']['history']['title_emb'],
                             batch['candidate_features']['title_emb'],
                             batch['user_features']['other']['user_index'])


def make_model(cfg):
    """make_model.py:15-55: scorer from cfg.scoring, model class from cfg.model.  CAUM (a candidate-aware model, not a
    bi-encoder) and the 'nonlin' scorer (referenced by the reference factory but defined nowhere in it) are rejected."""
    c = _cfg(cfg)
    emb_dim, bias = c.total_emb_dim, bool(c.bias)
    if c.scoring == 'dot':
        scoring_fn = DotScoring()
    elif c.scoring == 'bilin':
        scoring_fn = BilinScoring(emb_dim, bias=bias)
    elif c.scoring == 'fc':
        scoring_fn = FCScoring(emb_dim, hidden_dim=emb_dim // 2, bias=bias)
    elif c.scoring in ('nonlin', 'CAUMScoring'):
        raise NotImplementedError(f'scoring {c.scoring!r}: NonLinScoring does not exist in the reference (make_model.py:25-26 would '
                                  f'raise AttributeError) and CAUMScoring belongs to the candidate-aware CAUM model (SURVEY §2 row 12)')
    else:
        raise ValueError(f'invalid value for cfg.scoring: {c.scoring}')
    models = {'standard': StandardRec, 'base': BaseRec, 'mean': MeanRec, 'NRMS': NRMS, 'NAML': NAML, 'smallNAML': SmallNAML,
              'NPA': NPA, 'LSTUR': LSTUR}
    if c.model in models:
        return models[c.model](c, scoring_fn)
    if c.model == 'CAUM':
        raise NotImplementedError("model 'CAUM' is not a bi-encoder (its user vector depends on the candidate): SURVEY §2 row 12")
    raise ValueError(f'invalid value for cfg.model: {c.model}')
