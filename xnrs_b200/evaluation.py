"""Full-catalogue evaluation (rows E + Me; north-star item 7).

The reference evaluates one impression per step: it re-encodes the whole history and every candidate with the
news encoder, scores, copies to the host and calls numpy/sklearn (xnrs/training.py:194-243, DataLoader batch 1).
Every title is encoded independently of the user (news_encoding.py:48-57), so here

  phase 1  encodes each catalogue article ONCE (sharded over ranks, all-gathered over NVLink) — article 0 is the
           pad article, whose vector news_encoder(zeros, zero mask) is what padded history slots must see
           (non-zero for biased heads);
  phase 2  walks CSR impressions (sharded over ranks by candidate count): gather history vectors, run the user
           encoder, then ONE kernel per chunk scores every candidate, ranks each impression with a segmented rank
           sort and emits AUC / RR / nDCG@5/10 / CTR@1/10; metric sums are all-reduced (5+2 doubles).

NPA's news vectors depend on the user (npa.py:67-68), so NPA evaluates by direct forward passes on padded chunks.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.distributed as dist

from . import kernels as K
from .data import IndexedTitles, TitleStore
from .distributed import rank, shard_range, world
from .models.components import AdditiveAttention, UserEncoder, _apply_head
from .models.zoo import LSTUR, NAML, NPA

METRIC_NAMES = ('auc', 'rr', 'ndcg@5', 'ndcg@10', 'ctr@1', 'ctr@10')


def balanced_impression_shards(offsets: torch.Tensor, w: int):
    """contiguous impression ranges [lo, hi) per rank with (nearly) equal candidate counts"""
    n_imp = offsets.numel() - 1
    total = int(offsets[-1])
    cuts = [0]
    for r in range(1, w):
        target = total * r // w
        cuts.append(int(torch.searchsorted(offsets, torch.tensor(target, dtype=offsets.dtype, device=offsets.device))))
    cuts.append(n_imp)
    cuts = [min(max(c, 0), n_imp) for c in cuts]
    for i in range(1, len(cuts)):
        cuts[i] = max(cuts[i], cuts[i - 1])
    return [(cuts[i], cuts[i + 1]) for i in range(w)]


class CatalogueEvaluator:
    def __init__(self, model, store: TitleStore, category: Optional[torch.Tensor] = None,
                 subcategory: Optional[torch.Tensor] = None, abstract_store: Optional[TitleStore] = None,
                 news_chunk: int = 8192, impression_chunk: int = 16384, score_act: int = 1):
        self.model, self.store, self.abstract_store = model, store, abstract_store
        self.device = store.device
        self.category = None if category is None else category.to(self.device)
        self.subcategory = None if subcategory is None else subcategory.to(self.device)
        self.news_chunk, self.impression_chunk, self.score_act = news_chunk, impression_chunk, score_act
        self.news_vecs: Optional[torch.Tensor] = None
        self.news_mask: Optional[torch.Tensor] = None
        self.news_logit: Optional[torch.Tensor] = None     # per-article pooling logit (item_logits=True)
        # With frozen weights the additive pooler's logit of a history slot depends only on the article in it, so it is
        # computed once per catalogue article (one GEMM over n_news rows) instead of once per (user, slot): same values as
        # layers.py:60-65 slot by slot, ~H * n_impressions / n_news fewer pooler FLOPs.  Off: re-run the pooler per slot.
        self.item_logits = True
        # also report the thresholded epoch metrics of _test_step (acc / rec / prec / confusion, training.py:219-222)
        self.binary_metrics = False

    # ---- phase 1 ---------------------------------------------------------------------------------------------
    def _encode_ids(self, ids: torch.Tensor):
        """ids (n,) -> (vectors (n,T), collapsed mask (n,)) through the model's own news encoder"""
        m = self.model
        idx = ids.view(-1, 1)
        title = IndexedTitles(self.store, idx, None, True)        # catalogue slices: distinct ids, nothing to de-duplicate
        if isinstance(m, NAML):
            e, mask = m._news(title, IndexedTitles(self.abstract_store, idx, None, True), self.category[ids.long()].view(-1, 1),
                              self.subcategory[ids.long()].view(-1, 1))
        elif isinstance(m, LSTUR):
            sub = self.subcategory[ids.long()].view(-1, 1) if 'subcategory_index' in m.cfg.catg_features else None
            e, mask = m.news_encoder(title, self.category[ids.long()].view(-1, 1), sub)
        else:
            e, mask = m.news_encoder(title)
        return e.reshape(ids.numel(), -1), mask.reshape(-1)

    @torch.no_grad()
    def encode_catalogue(self) -> torch.Tensor:
        """encode every article once; rank r encodes its slice, then one all-gather"""
        n = self.store.title_tokens.shape[0]
        w, r = world(), rank()
        per = (n + w - 1) // w
        lo, hi = min(r * per, n), min((r + 1) * per, n)
        vecs, masks = [], []
        for a in range(lo, hi, self.news_chunk):
            ids = torch.arange(a, min(hi, a + self.news_chunk), device=self.device, dtype=torch.int32)
            e, mk = self._encode_ids(ids)
            vecs.append(e)
            masks.append(mk)
        T = vecs[0].shape[1] if vecs else self._encode_ids(torch.zeros(1, device=self.device, dtype=torch.int32))[0].shape[1]
        local = torch.zeros((per, T), device=self.device, dtype=torch.float32)
        lmask = torch.zeros(per, device=self.device, dtype=torch.float32)
        if vecs:
            local[:hi - lo] = torch.cat(vecs)
            lmask[:hi - lo] = torch.cat(masks)
        if w > 1:
            allv = torch.empty((w * per, T), device=self.device, dtype=torch.float32)
            allm = torch.empty(w * per, device=self.device, dtype=torch.float32)
            dist.all_gather_into_tensor(allv, local)
            dist.all_gather_into_tensor(allm, lmask)
            local, lmask = allv, allm
        self.news_vecs, self.news_mask = local[:n].contiguous(), lmask[:n].contiguous()
        self.news_logit = None
        pooler = self._item_pooler()
        if pooler is not None:
            hid = K.gemm(self.news_vecs, pooler.fc1.weight, trans_b=True, bias=pooler.fc1.bias, act=K.ACT_TANH)
            self.news_logit = K.rowdot(hid, pooler.fc2.weight.reshape(-1), pooler.fc2.bias)
        return self.news_vecs

    def _item_pooler(self) -> Optional[AdditiveAttention]:
        """the user-level additive pooler when the user encoder is [no dropout, no attention] -> pooler (-> head):
        StandardRec / CL (user_encoding.py:69-77 with att=None) and NAML (naml.py:54-57,109)"""
        if not self.item_logits:
            return None
        ue = getattr(self.model, 'user_encoder', None)
        if isinstance(self.model, NAML) and isinstance(ue, AdditiveAttention):
            pooler = ue
        elif type(ue) is UserEncoder and ue.att is None and isinstance(ue.pooler, AdditiveAttention):
            pooler = ue.pooler
        else:
            return None
        T = self.news_vecs.shape[1]
        return pooler if (T % 4 == 0 and T <= 1024 and pooler.fc1.weight.shape[1] == T) else None

    # ---- phase 2 ---------------------------------------------------------------------------------------------
    def _users(self, hist_ids: torch.Tensor, user_index: Optional[torch.Tensor]) -> torch.Tensor:
        B, H = hist_ids.shape
        if self.news_logit is not None:
            pooled = K.logitpool(self.news_vecs, self.news_logit, self.news_mask, hist_ids)
            ue = self.model.user_encoder
            return _apply_head(ue.head, pooled) if hasattr(ue, 'head') else pooled
        flat = hist_ids.reshape(-1)
        h = K.gather_rows(self.news_vecs, flat).view(B, H, -1)
        hm = self.news_mask[flat.long()].view(B, H, 1)            # index plumbing: collapsed title mask per slot
        m = self.model
        if isinstance(m, NAML):
            u = m._user(h, hm)
        elif isinstance(m, LSTUR):
            u = m.user_encoder((h, hm), user_index)
        else:
            u = m.user_encoder((h, hm))
        return u.reshape(B, -1)

    @torch.no_grad()
    def evaluate(self, impressions: Dict[str, torch.Tensor], return_per_impression: bool = False):
        """impressions: hist_ids (n_imp,H) int32, cand_ids (n_cand,) int32, offsets (n_imp+1,) int64, targets (n_cand,)
        fp32 [, user_index (n_imp,1)].  Returns the epoch means (unweighted over impressions, training.py:257-266)."""
        if isinstance(self.model, NPA):
            return self._evaluate_direct(impressions, return_per_impression)
        if self.news_vecs is None:
            self.encode_catalogue()
        dev = self.device
        w, r = world(), rank()
        off_in = impressions['offsets']
        n_imp = off_in.numel() - 1
        lo, hi = balanced_impression_shards(off_in, w)[r] if w > 1 else (0, n_imp)
        # this rank's contiguous CSR shard only: sliced where the buffers live (host buffers: only the shard crosses PCIe)
        c_lo, c_hi = int(off_in[lo]), int(off_in[hi])
        off_host = (off_in[lo:hi + 1] - c_lo) if not off_in.is_cuda else (off_in[lo:hi + 1] - c_lo).cpu()   # chunk bounds: no per-chunk sync
        offsets = (off_in[lo:hi + 1] - c_lo).to(dev, non_blocking=True)
        hist_ids = impressions['hist_ids'][lo:hi].to(dev, non_blocking=True)
        cand_ids = impressions['cand_ids'][c_lo:c_hi].to(dev, non_blocking=True)
        targets = impressions['targets'][c_lo:c_hi].to(dev, non_blocking=True)
        uidx = impressions.get('user_index')
        uidx = None if uidx is None else uidx[lo:hi].to(dev, non_blocking=True)
        self.last_shard = {'impressions': hi - lo, 'candidates': c_hi - c_lo}
        lo, hi = 0, hi - lo                                       # everything below is in shard-local coordinates
        sums = torch.zeros(7, device=dev, dtype=torch.float64)
        bsums = torch.zeros(7, device=dev, dtype=torch.float64)
        per_imp, all_scores = [], []
        for a in range(lo, hi, self.impression_chunk):
            b = min(hi, a + self.impression_chunk)
            u = self._users(hist_ids[a:b].contiguous(), None if uidx is None else uidx[a:b].contiguous())
            c0, c1 = int(off_host[a]), int(off_host[b])
            local_off = (offsets[a:b + 1] - c0).contiguous()
            scores, metrics = K.eval_impressions(u, self.news_vecs, cand_ids[c0:c1].contiguous(), local_off,
                                                 targets[c0:c1].contiguous(), act=self.score_act)
            K.call('xnrs_metric_sums', metrics, b - a, sums)
            if self.binary_metrics:
                bsums += K.binary_metrics(scores, targets[c0:c1].contiguous(), local_off).sum(0)      # epoch bookkeeping (7 numbers)
            if return_per_impression:
                per_imp.append(metrics)
                all_scores.append(scores)
        if w > 1:
            dist.all_reduce(sums)
            if self.binary_metrics:
                dist.all_reduce(bsums)
        out = {name: float(sums[i] / sums[6]) if float(sums[6]) > 0 else float('nan') for i, name in enumerate(METRIC_NAMES)}
        out['impressions'] = int(sums[6])
        if self.binary_metrics:
            b_ = bsums.tolist()                                    # (after the all-reduce: sums over ALL impressions)
            n_all = max(n_imp, 1)
            out.update({'acc': b_[0] / n_all, 'rec': b_[1] / n_all, 'prec': b_[2] / n_all,
                        'conf': [[int(b_[3]), int(b_[4])], [int(b_[5]), int(b_[6])]], 'candidates': int(sum(b_[3:]))})
        if return_per_impression:
            out['per_impression'] = torch.cat(per_imp) if per_imp else None
            out['scores'] = torch.cat(all_scores) if all_scores else None
        return out

    @torch.no_grad()
    def _evaluate_direct(self, impressions, return_per_impression):
        """models whose news vectors depend on the user (NPA): forward passes over padded impression chunks"""
        dev = self.device
        offsets = impressions['offsets'].to(dev)
        n_imp = offsets.numel() - 1
        w, r = world(), rank()
        lo, hi = balanced_impression_shards(offsets, w)[r] if w > 1 else (0, n_imp)
        cand_ids, targets = impressions['cand_ids'].to(dev), impressions['targets'].to(dev)
        hist_ids, uidx = impressions['hist_ids'].to(dev), impressions['user_index'].to(dev)
        sums = torch.zeros(7, device=dev, dtype=torch.float64)
        per_imp, all_scores = [], []
        chunk = max(1, min(self.impression_chunk, 256))
        for a in range(lo, hi, chunk):
            b = min(hi, a + chunk)
            sizes = (offsets[a + 1:b + 1] - offsets[a:b])
            nmax = int(sizes.max())
            c0, c1 = int(offsets[a]), int(offsets[b])
            pos = torch.arange(nmax, device=dev)[None, :]
            valid = pos < sizes[:, None]
            padded = torch.zeros((b - a, nmax), device=dev, dtype=torch.int32)
            padded[valid] = cand_ids[c0:c1]                       # pad with article 0 (index plumbing only)
            batch = {'user_features': {'history': {'title_emb': self.store.index(hist_ids[a:b].contiguous())},
                                       'other': {'user_index': uidx[a:b].contiguous()}},
                     'candidate_features': {'title_emb': self.store.index(padded)}}
            s = self.model(batch).reshape(b - a, nmax)
            flat = s[valid].contiguous()
            if self.score_act == 1:
                flat = K.ReluFn.apply(flat)
            _, metrics = K.eval_impressions(None, None, None, (offsets[a:b + 1] - c0).contiguous(),
                                            targets[c0:c1].contiguous(), act=0, scores=flat)
            K.call('xnrs_metric_sums', metrics, b - a, sums)
            if return_per_impression:
                per_imp.append(metrics)
                all_scores.append(flat)
        if w > 1:
            dist.all_reduce(sums)
        out = {name: float(sums[i] / sums[6]) if float(sums[6]) > 0 else float('nan') for i, name in enumerate(METRIC_NAMES)}
        out['impressions'] = int(sums[6])
        if return_per_impression:
            out['per_impression'] = torch.cat(per_imp) if per_imp else None
            out['scores'] = torch.cat(all_scores) if all_scores else None
        return out
