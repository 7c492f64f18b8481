"""CPU oracle for the xnrs bi-encoder hot path.  TEST INFRASTRUCTURE ONLY.

This file is a plain torch-CPU (fp32 or fp64) *restatement* of the reference algorithm, written as
stateless functions over a flat ``state_dict``-style parameter mapping.  It is NOT the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it, and only as the checker / CPU baseline.  The product path (``xnrs_b200``) never
imports anything from ``oracle/`` and raises if its CUDA library is missing.

Parity pin: every function below is checked against outputs of the *unmodified* reference package
(imported from /root/reference in the build container by ``tests/golden/make_golden.py``) through the
fixtures committed under ``tests/golden/*.npz`` — see ``tests/test_oracle_golden.py``.  The reference
ships no golden vectors of its own, so those reference-generated fixtures plus
the known-answer values of SURVEY.md §8(c) are the pin.

Each function cites the reference lines it follows (paths relative to the reference repo root).
Gradients for the backward-pass checks come from torch autograd over these same functions.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import numpy as np
import torch

Tensor = torch.Tensor
Params = Dict[str, Tensor]


# --------------------------------------------------------------------------------------------
# helpers
# --------------------------------------------------------------------------------------------

def as_params(sd, dtype=torch.float32) -> Params:
    """numpy / torch state_dict -> dict of CPU tensors in `dtype` (ints left alone)."""
    out = {}
    for k, v in sd.items():
        t = torch.as_tensor(np.asarray(v)) if not isinstance(v, torch.Tensor) else v.detach().cpu()
        out[k] = t.to(dtype) if t.is_floating_point() else t
    return out


def _lin(x: Tensor, P: Params, prefix: str) -> Tensor:
    """y = x W^T (+ b) for an nn.Linear stored at `prefix`.{weight,bias} (bias optional)."""
    y = x @ P[prefix + '.weight'].T
    b = P.get(prefix + '.bias')
    return y if b is None else y + b


def collapse_mask(m: Tensor, dim: int) -> Tensor:
    """xnrs/utils.py:74-75 — 1 where any token of the title is unmasked."""
    return m.sum(dim=dim).clamp(0, 1)


# --------------------------------------------------------------------------------------------
# row G — gather (xnrs/data/dataset.py:63-65,77-85,97-109 re-expressed over device tables)
# --------------------------------------------------------------------------------------------

def gather_titles(token_table: Tensor, title_tokens: Tensor, news_ids: Tensor) -> Tuple[Tensor, Tensor]:
    """news ids (..,) -> (x (..,S,D), m (..,S,1)).

    token id 0 is the pad token (all-zero row, mask 0); news id 0 is the pad article (all pad
    tokens), which reproduces the zero-embedding / zero-mask history padding of dataset.py:82-85.
    """
    tok = title_tokens[news_ids.long()]                 # (.., S)
    x = token_table[tok.long()]                         # (.., S, D) bit-exact row copies
    m = (tok != 0).to(token_table.dtype).unsqueeze(-1)
    return x, m


# --------------------------------------------------------------------------------------------
# row A — additive attention pooling (xnrs/models/components/layers.py:47-69)
# --------------------------------------------------------------------------------------------

def additive_attention(x: Tensor, m: Optional[Tensor], P: Params, prefix: str,
                       return_weights: bool = False):
    h = torch.tanh(_lin(x, P, prefix + '.fc1'))
    a = torch.exp(_lin(h, P, prefix + '.fc2'))           # un-stabilised exp (layers.py:61)
    if m is not None:
        a = a * m
    a = a / (a.sum(dim=1, keepdim=True) + 1e-8)          # layers.py:64
    out = (a * x).sum(dim=1, keepdim=True)               # == bmm(a^T, x)
    return (out, a) if return_weights else out


def masked_mean(x: Tensor, m: Tensor) -> Tensor:
    """layers.py:25-37"""
    return (x * m).sum(dim=1, keepdim=True) / (m.sum(dim=1, keepdim=True) + 1e-8)


# --------------------------------------------------------------------------------------------
# row P — personalised attention (layers.py:88-101)
# --------------------------------------------------------------------------------------------

def personalized_attention(q: Tensor, x: Tensor, m: Optional[Tensor], P: Params, prefix: str) -> Tensor:
    xa = torch.tanh(_lin(x, P, prefix + '.x_fc'))        # (R, L, hd)
    qq = _lin(q, P, prefix + '.q_fc')                    # (R, 1, hd)
    a = torch.exp((xa * qq).sum(dim=-1, keepdim=True))   # per-row dot with the user's query
    if m is not None:
        a = a * m
    a = a / (a.sum(dim=1, keepdim=True) + 1e-8)
    return (a * x).sum(dim=1, keepdim=True)


# --------------------------------------------------------------------------------------------
# row M — multi-head self-attention with QUERY-axis masking (layers.py:121-156)
# --------------------------------------------------------------------------------------------

def multi_head_attention(x: Tensor, m: Optional[Tensor], P: Params, prefix: str, n_heads: int,
                         keep: Optional[Tensor] = None, p_drop: float = 0.0) -> Tensor:
    """`keep` is an optional explicit 0/1 dropout keep-mask (R,h,L,L) applied as att*keep/(1-p)."""
    R, L, D = x.shape
    dk = D // n_heads
    q = _lin(x, P, prefix + '.q_linear').view(R, L, n_heads, dk).transpose(1, 2)
    k = _lin(x, P, prefix + '.k_linear').view(R, L, n_heads, dk).transpose(1, 2)
    v = _lin(x, P, prefix + '.v_linear').view(R, L, n_heads, dk).transpose(1, 2)
    att = (q @ k.transpose(-2, -1)) / math.sqrt(dk)
    if m is not None:
        # (R,L,1) -> (R,1,L,1): whole *query rows* are filled, keys are never masked (layers.py:142-144)
        att = att.masked_fill(m.unsqueeze(1) == 0, -1e9)
    att = torch.softmax(att, dim=-1)
    if keep is not None:
        att = att * keep / (1.0 - p_drop)
    o = (att @ v).transpose(1, 2).reshape(R, L, D)
    return _lin(o, P, prefix + '.out')


# --------------------------------------------------------------------------------------------
# rows T / U — encoders (news_encoding.py:34-60, user_encoding.py:50-81)
# --------------------------------------------------------------------------------------------

def _head(x: Tensor, P: Params, prefix: str) -> Tensor:
    if prefix + '.0.weight' not in P:
        return x
    return _lin(torch.relu(_lin(x, P, prefix + '.0')), P, prefix + '.2')


def text_encoder(x: Tensor, m: Tensor, P: Params, prefix: str, n_heads: int = 0,
                 keep: Optional[Tensor] = None, p_drop: float = 0.0) -> Tuple[Tensor, Tensor]:
    """(b,n,S,D),(b,n,S,1) -> (b,n,E),(b,n,1). Dropout p=cfg.p_dropout=0 in every config: omitted."""
    b, n, s, d = x.shape
    xr, mr = x.reshape(b * n, s, d), m.reshape(b * n, s, 1)
    if prefix + '.att.q_linear.weight' in P:
        xr = multi_head_attention(xr, mr, P, prefix + '.att', n_heads, keep, p_drop)
    e = _head(additive_attention(xr, mr, P, prefix + '.pooler'), P, prefix + '.head')
    return e.reshape(b, n, -1), collapse_mask(m, dim=2)


def user_encoder(h: Tensor, hm: Tensor, P: Params, prefix: str, n_heads: int = 0,
                 return_weights: bool = False, keep: Optional[Tensor] = None, p_drop: float = 0.0):
    if prefix + '.att.q_linear.weight' in P:
        h = multi_head_attention(h, hm, P, prefix + '.att', n_heads, keep, p_drop)
    if prefix + '.pooler.fc1.weight' in P:
        u, a = additive_attention(h, hm, P, prefix + '.pooler', return_weights=True)
    else:                                                   # MaskedMean pooler (no params)
        u, a = masked_mean(h, hm), None
    u = _head(u, P, prefix + '.head')
    return (u, a) if return_weights else u


def dot_scoring(u: Tensor, c: Tensor, normalize: bool = False) -> Tensor:
    """scoring.py:12-23: (B,1,T),(B,N,T) -> (B,N,1)"""
    if normalize:
        u = u / u.norm(p=2, dim=2, keepdim=True)
        c = c / c.norm(p=2, dim=2, keepdim=True)
    return c @ u.transpose(-1, -2)


# --------------------------------------------------------------------------------------------
# GRU, final hidden state at each sequence's true length (lstur.py:139-153; nn.GRU gate order r,z,n)
# --------------------------------------------------------------------------------------------

def gru_last_hidden(x: Tensor, lengths: Tensor, P: Params, prefix: str, h0: Optional[Tensor] = None) -> Tensor:
    """x (B,L,I) front-aligned, lengths (B,) -> h at step lengths[b] (B,Hd).  length 0 -> h0."""
    w_ih, w_hh = P[prefix + '.weight_ih_l0'], P[prefix + '.weight_hh_l0']
    b_ih, b_hh = P[prefix + '.bias_ih_l0'], P[prefix + '.bias_hh_l0']
    B, L, _ = x.shape
    Hd = w_hh.shape[1]
    h = x.new_zeros(B, Hd) if h0 is None else h0
    for t in range(L):
        gi = x[:, t] @ w_ih.T + b_ih
        gh = h @ w_hh.T + b_hh
        r = torch.sigmoid(gi[:, :Hd] + gh[:, :Hd])
        z = torch.sigmoid(gi[:, Hd:2 * Hd] + gh[:, Hd:2 * Hd])
        nn_ = torch.tanh(gi[:, 2 * Hd:] + r * gh[:, 2 * Hd:])
        hn = (1 - z) * nn_ + z * h
        live = (lengths > t).to(x.dtype).unsqueeze(1)
        h = live * hn + (1 - live) * h
    return h


# --------------------------------------------------------------------------------------------
# full models.  `batch` uses the reference's dict schema.
# --------------------------------------------------------------------------------------------

def _hist(batch, feat='title_emb'):
    x, m = batch['user_features']['history'][feat]
    return x, m


def _cand(batch, feat='title_emb'):
    x, m = batch['candidate_features'][feat]
    return x, m


def parent_forward(P: Params, batch, n_heads: int = 0, return_embeddings: bool = False,
                   keeps: Optional[dict] = None, p_drop: float = 0.0):
    """ParentRec._forward (parent.py:23-38) — StandardRec (= CL model) and NRMS."""
    keeps = keeps or {}
    h, hm = text_encoder(*_hist(batch), P, 'news_encoder', n_heads, keeps.get('hist'), p_drop)
    c, _ = text_encoder(*_cand(batch), P, 'news_encoder', n_heads, keeps.get('cand'), p_drop)
    u = user_encoder(h, hm, P, 'user_encoder', n_heads, keep=keeps.get('user'), p_drop=p_drop)
    r = dot_scoring(u, c)
    return (r, u, c) if return_embeddings else r


def parent_user_embeddings(P: Params, batch, n_heads: int = 0) -> Tensor:
    """parent.py:49-81 / standard_model.py:39-71"""
    h, hm = text_encoder(*_hist(batch), P, 'news_encoder', n_heads)
    return user_encoder(h, hm, P, 'user_encoder', n_heads).squeeze(1)


def _naml_news(P: Params, title, abstract, ctg, subctg):
    t, mask = text_encoder(*title, P, 'title_encoder')
    a, _ = text_encoder(*abstract, P, 'body_encoder')
    c = _lin(P['cat_embedder.weight'][ctg.long()], P, 'cat_fc')
    s = _lin(P['subcat_embedder.weight'][subctg.long()], P, 'subcat_fc')
    b, n, e = t.shape
    views = torch.stack([t, a, c, s], dim=2).reshape(b * n, 4, e)        # naml.py:96-101
    return additive_attention(views, None, P, 'feature_pooler').reshape(b, n, e), mask


def naml_forward(P: Params, batch, return_embeddings: bool = False):
    """naml.py:61-112"""
    hf, cf = batch['user_features']['history'], batch['candidate_features']
    h, hm = _naml_news(P, hf['title_emb'], hf['abstract_emb'], hf['category_index'], hf['subcategory_index'])
    c, _ = _naml_news(P, cf['title_emb'], cf['abstract_emb'], cf['category_index'], cf['subcategory_index'])
    u = additive_attention(h, hm, P, 'user_encoder')
    r = dot_scoring(u, c)
    return (r, u, c) if return_embeddings else r


def naml_user_embeddings(P: Params, batch) -> Tensor:
    """naml.py:113-147 — returned un-squeezed (B,1,E)."""
    hf = batch['user_features']['history']
    h, hm = _naml_news(P, hf['title_emb'], hf['abstract_emb'], hf['category_index'], hf['subcategory_index'])
    return additive_attention(h, hm, P, 'user_encoder')


def _lstur_news(P: Params, title, ctg, subctg=None):
    """lstur.py:191-207"""
    t, mask = text_encoder(*title, P, 'news_encoder.title_encoder')
    e = torch.cat([t, P['news_encoder.cat_embedder.weight'][ctg.long()]], dim=2)
    if subctg is not None:
        e = torch.cat([e, P['news_encoder.subcat_embedder.weight'][subctg.long()]], dim=2)
    return e, mask


def _lstur_user(P: Params, h, hm, user_ids, method: str, st_hist_len: int,
                user_keep: Optional[Tensor] = None, p_user_drop: float = 0.0):
    """lstur.py:118-159 with long_term_method == 'embedding'."""
    u_lt = P['user_encoder.long_term_encoder.weight'][user_ids.long()].squeeze(1)
    if user_keep is not None:
        u_lt = u_lt * user_keep / (1.0 - p_user_drop)
    h_st, hm_st = h[:, :st_hist_len], hm[:, :st_hist_len]
    lengths = hm_st.sum(dim=1).squeeze(1)
    if method == 'ini':
        return gru_last_hidden(h_st, lengths, P, 'user_encoder.gru', u_lt).unsqueeze(1)
    if method == 'con':
        u_st = gru_last_hidden(h_st, lengths, P, 'user_encoder.gru')
        return torch.cat([u_st, u_lt], dim=1).unsqueeze(1)
    if method == 'lt_only':
        return u_lt.unsqueeze(1)
    raise ValueError(method)


def lstur_forward(P: Params, batch, method: str = 'con', st_hist_len: int = 25, use_subcat: bool = False,
                  return_embeddings: bool = False, user_keep=None, p_user_drop: float = 0.0):
    """lstur.py:18-62"""
    hf, cf = batch['user_features']['history'], batch['candidate_features']
    h, hm = _lstur_news(P, hf['title_emb'], hf['category_index'], hf['subcategory_index'] if use_subcat else None)
    c, _ = _lstur_news(P, cf['title_emb'], cf['category_index'], cf['subcategory_index'] if use_subcat else None)
    u = _lstur_user(P, h, hm, batch['user_features']['other']['user_index'], method, st_hist_len,
                    user_keep, p_user_drop)
    r = dot_scoring(u, c)
    return (r, u, c) if return_embeddings else r


def lstur_user_embeddings(P: Params, batch, method: str = 'con', st_hist_len: int = 25,
                          use_subcat: bool = False) -> Tensor:
    """lstur.py:63-79"""
    hf = batch['user_features']['history']
    h, hm = _lstur_news(P, hf['title_emb'], hf['category_index'], hf['subcategory_index'] if use_subcat else None)
    return _lstur_user(P, h, hm, batch['user_features']['other']['user_index'], method, st_hist_len).squeeze(1)


def npa_forward(P: Params, batch) -> Tensor:
    """npa.py:34-89"""
    (h, hm), (c, cm) = _hist(batch), _cand(batch)
    uid = batch['user_features']['other']['user_index']
    ue = P['user_embedder.weight'][uid.long()]                          # (B,1,du)
    b, nh, s, d = h.shape
    nc = c.shape[1]

    def news(x, m, n):
        q = ue.repeat_interleave(n, dim=0)
        p = personalized_attention(q, x.reshape(b * n, s, d), m.reshape(b * n, s, 1), P, 'title_pooler')
        return _head(p, P, 'news_head').reshape(b, n, -1)

    hh = news(h, hm, nh)
    u = personalized_attention(ue, hh, collapse_mask(hm, dim=2), P, 'user_encoder')
    return dot_scoring(u, news(c, cm, nc))


# --------------------------------------------------------------------------------------------
# loss hooks (xnrs/training.py:326-331, 336-342, 378-392, 433-472; xnrs/utils.py:117-131)
# --------------------------------------------------------------------------------------------

def mse_relu_loss(scores: Tensor, targets: Tensor, weights: Optional[Tensor] = None):
    """MSERankingTrainer: prediction = relu(score) (training.py:388-392), mean over B*N (:378-386)."""
    p = torch.relu(scores)
    l = (p - targets) ** 2
    if weights is not None:
        l = l * weights
    return l.mean(), p


def bce_logits_loss(scores: Tensor, targets: Tensor, weights: Optional[Tensor] = None) -> Tensor:
    """training.py:336-337: binary_cross_entropy_with_logits, mean reduction."""
    l = torch.clamp(scores, min=0) - scores * targets + torch.log1p(torch.exp(-scores.abs()))
    if weights is not None:
        l = l * weights
    return l.mean()


def bce_sigmoid_loss(scores: Tensor, targets: Tensor, weights: Optional[Tensor] = None):
    """BCERankingTrainer (training.py:324-331): prediction = sigmoid(score), loss = nn.BCELoss (each log term clamped
    at -100 like torch.nn.functional.binary_cross_entropy), mean reduction."""
    p = torch.sigmoid(scores)
    l = -(targets * torch.clamp(torch.log(p), min=-100.0) + (1 - targets) * torch.clamp(torch.log(1 - p), min=-100.0))
    if weights is not None:
        l = l * weights
    return l.mean(), p


def ranking_nll(p: Tensor, n: Tensor, reduction: str = 'mean') -> Tensor:
    """utils.py:117-131 — un-stabilised softmax NLL of the positive among 1+K."""
    ep = torch.exp(p)
    l = -torch.log(ep / (ep + torch.exp(n).sum(dim=1, keepdim=True)))
    return l.mean() if reduction == 'mean' else l


def contrastive_loss(emb: Tensor, labels: Tensor, temperature: float) -> Tensor:
    """training.py:433-472 in closed form (the per-anchor python loop vectorised).

    For every anchor i that has at least one other sample with its label:
        -log( sum_{j in pos(i)} e^{s_ij/t} / (sum_{j != i} e^{s_ij/t} + 1e-12) )
    summed and divided by (number of such anchors + 1e-8).
    """
    e = emb.reshape(emb.shape[0], -1)
    e = e / e.norm(dim=-1, keepdim=True).clamp_min(1e-12)               # F.normalize, eps 1e-12
    ex = torch.exp((e @ e.T) / temperature)
    B = e.shape[0]
    off = ~torch.eye(B, dtype=torch.bool, device=e.device)
    pos = (labels[:, None] == labels[None, :]) & off
    num = (ex * pos).sum(dim=1)
    den = (ex * off).sum(dim=1)
    has = pos.any(dim=1)
    per = -torch.log(num[has] / (den[has] + 1e-12))
    return per.sum() / (has.sum().to(e.dtype) + 1e-8) if has.any() else e.new_zeros(())


def theme_labels(themes) -> Tensor:
    """training.py:414-417 maps theme strings to indices through a python set (arbitrary but
    consistent numbering); the loss only uses label *equality*, so any consistent numbering works."""
    if isinstance(themes, torch.Tensor):
        return themes.long()
    order = {}
    return torch.tensor([order.setdefault(t, len(order)) for t in themes], dtype=torch.long)


def contrastive_train_loss(scores: Tensor, targets: Tensor, user_emb: Tensor, labels: Tensor,
                           temperature: float, lam: float):
    """ContrastiveRankingTrainer._train_step (training.py:402-431): MSE(relu(s),t) + lam * CL."""
    l_rec, _ = mse_relu_loss(scores, targets)
    l_cl = contrastive_loss(user_emb, labels, temperature)
    return l_rec + lam * l_cl, l_rec, l_cl


def adam_step(p: Tensor, g: Tensor, m: Tensor, v: Tensor, step: int, lr: float,
              b1: float = 0.9, b2: float = 0.999, eps: float = 1e-8):
    """torch.optim.Adam defaults (training.py:39): no weight decay, no amsgrad. In-place on p,m,v."""
    m.mul_(b1).add_(g, alpha=1 - b1)
    v.mul_(b2).addcmul_(g, g, value=1 - b2)
    bc1, bc2 = 1 - b1 ** step, 1 - b2 ** step
    p.addcdiv_(m, (v.sqrt() / math.sqrt(bc2)).add_(eps), value=-lr / bc1)


# --------------------------------------------------------------------------------------------
# row Me — ranking metrics (xnrs/evaluation/metrics.py:7-44), float64 like numpy
# Tie policy: descending score, ties by DESCENDING original index,
# i.e. np.argsort(score, kind='stable')[::-1].  AUC is tie-aware and needs no policy.
# --------------------------------------------------------------------------------------------

def rank_order(y_score: np.ndarray) -> np.ndarray:
    return np.argsort(np.asarray(y_score), kind='stable')[::-1]


def nan_to_num_scores(s: np.ndarray) -> np.ndarray:
    """training.py:211"""
    return np.nan_to_num(s, nan=0.0, posinf=1.0, neginf=0.0)


def dcg_score(y_true, y_score, k=10) -> float:
    """metrics.py:9-14"""
    t = np.take(np.asarray(y_true, dtype=np.float64), rank_order(y_score)[:k])
    return float(np.sum((2.0 ** t - 1.0) / np.log2(np.arange(len(t)) + 2.0)))


def ndcg_score(y_true, y_score, k=10) -> float:
    """metrics.py:17-20.  The ideal ordering sorts y_true by itself (ties are harmless there)."""
    return dcg_score(y_true, y_score, k) / dcg_score(y_true, y_true, k)


def rr_score(y_true, y_score) -> float:
    """metrics.py:31-38 — reciprocal rank of the first positive."""
    t = np.take(np.asarray(y_true, dtype=np.float64), rank_order(y_score))
    return float(np.max(t / (np.arange(len(t)) + 1.0)))


def ctr_score(y_true, y_score, k=1) -> float:
    """metrics.py:41-44"""
    return float(np.mean(np.take(np.asarray(y_true, dtype=np.float64), rank_order(y_score)[:k])))


def auc_score(y_true, y_score) -> float:
    """metrics.py:7 (sklearn.roc_auc_score, binary): Mann-Whitney statistic, ties count 1/2."""
    t = np.asarray(y_true) > 0.5
    s = np.asarray(y_score, dtype=np.float64)
    pos, neg = s[t], s[~t]
    if len(pos) == 0 or len(neg) == 0:
        return float('nan')                    # the reference raises here (Appendix A.11)
    gt = (pos[:, None] > neg[None, :]).sum()
    eq = (pos[:, None] == neg[None, :]).sum()
    return float((gt + 0.5 * eq) / (len(pos) * len(neg)))


def impression_metrics(y_true, y_score) -> Dict[str, float]:
    """the ranking part of RankingTrainer._test_step (training.py:211-218)."""
    s = nan_to_num_scores(np.asarray(y_score))
    return {
        'auc': auc_score(y_true, s), 'rr': rr_score(y_true, s),
        'ndcg@5': ndcg_score(y_true, s, 5), 'ndcg@10': ndcg_score(y_true, s, 10),
        'ctr@1': ctr_score(y_true, s, 1), 'ctr@10': ctr_score(y_true, s, 10),
    }
