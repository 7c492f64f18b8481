"""Trainer loss hooks of the reference (xnrs/training.py) on top of the fused kernels.

Kept: the hook names and semantics — ``_init_loss`` / ``self.L``, ``forward(batch)``, ``_train_step``,
``_test_step``, ``_compute_contrastive_loss`` — and Adam with torch defaults.  Dropped (host orchestration,
out of scope): dataloaders, epoch loops, wandb, checkpoint / CSV export.

Differences that are deliberate and documented in DESIGN.md:
  * the user embedding for the contrastive term is taken from the SAME forward pass as the scores
    (``share_user_forward=True``); the reference runs the history side twice (training.py:409), which
    gives identical values whenever dropout is off.  With dropout active in train mode (NRMS' attention
    dropout p=0.1, LSTUR's p_user_dropout=0.07) the reference draws two independent masks where this code
    draws one — the same distribution of each term, a different joint draw; ``share_user_forward=False``
    restores the second pass;
  * all parameters live in one flat buffer so Adam is one kernel launch and data-parallel training
    all-reduces one bucket.
"""
from __future__ import annotations

from typing import Dict, Iterable, List, Optional

import torch
import torch.nn as nn

from . import kernels as K
from .models.components import DotScoring
from .models.zoo import Cfg


def batch_to_device(batch: dict, device) -> None:
    """xnrs/utils.py:88-93 (tensors and nested dicts; (x, m) tuples are moved by the encoders themselves)."""
    for k, v in batch.items():
        if isinstance(v, torch.Tensor):
            batch[k] = v.to(device, non_blocking=True)
        elif isinstance(v, dict):
            batch_to_device(v, device)


def theme_labels(themes, device) -> torch.Tensor:
    """training.py:414-417 — theme strings -> int labels (only label equality matters). Accepts int tensors too."""
    if isinstance(themes, torch.Tensor):
        return themes.to(device=device, dtype=torch.int32).contiguous()
    order: Dict[str, int] = {}
    return torch.tensor([order.setdefault(t, len(order)) for t in themes], dtype=torch.int32, device=device)


class FlatAdam:
    """torch.optim.Adam(lr, defaults) (training.py:39) over ONE flat parameter / gradient buffer.

    Every parameter becomes a view into ``flat_p`` and owns a persistent ``.grad`` view into ``flat_g``
    (autograd accumulates into it in place), so the update is a single xnrs_adam_step launch and a
    data-parallel job all-reduces a single bucket.

    ``row_sparse``: embedding tables (the 703 790-row user tables of LSTUR / NPA, lstur.py:94-98, npa.py:12-15) whose gradient
    only ever arrives as a scatter-add of looked-up rows.  They are laid out AFTER the dense parameters and updated by
    xnrs_adam_rows over the rows touched at least once so far ("active"): a never-touched row has g = m = v = 0 and dense
    Adam leaves it exactly unchanged, so this is torch's dense Adam bit for bit at a fraction of its 7 x 383 MB of traffic
    per step (SURVEY §7 hard part 4: the explicit decision).  Their gradient is cleared row-wise too."""

    SPARSE_MIN_ROWS = 100_000

    def __init__(self, params: Iterable[nn.Parameter], lr: float, betas=(0.9, 0.999), eps: float = 1e-8,
                 graph_safe: bool = False, row_sparse: Iterable[nn.Parameter] = ()):
        sparse_ids = {id(p) for p in row_sparse if p.requires_grad and p.dim() == 2}
        params = [p for p in params if p.requires_grad]
        self.params: List[nn.Parameter] = ([p for p in params if id(p) not in sparse_ids]
                                           + [p for p in params if id(p) in sparse_ids])
        dev = self.params[0].device
        sizes = [((p.numel() + 3) // 4) * 4 for p in self.params]          # keep every view 16-byte aligned
        total = sum(sizes)
        self.flat_p = torch.zeros(total, device=dev, dtype=torch.float32)
        # 4 auxiliary floats sit in front of the gradient so that data-parallel scalars (the InfoNCE value) ride along in
        # the one gradient all-reduce and zero_grad clears them with the same kernel
        self.g_store = torch.zeros(total + 4, device=dev, dtype=torch.float32)
        self.g_aux, self.flat_g = self.g_store[:4], self.g_store[4:]
        self.m = torch.zeros(total, device=dev, dtype=torch.float32)
        self.v = torch.zeros(total, device=dev, dtype=torch.float32)
        off = 0
        self.ranges = {}                     # id(param) -> (begin, end) of its slice of the flat buffers
        self.tables = []                     # row-sparse tables: dict(param, begin, V, D, bitmap, active, count)
        self.dense_end = None
        for p, n in zip(self.params, sizes):
            self.ranges[id(p)] = (off, off + p.numel())
            view = self.flat_p[off:off + p.numel()].view_as(p)
            view.copy_(p.data)
            p.data = view
            p.grad = self.flat_g[off:off + p.numel()].view_as(p)
            p._xnrs_direct = True          # kernels.py: weight gradients are accumulated straight into this view
            if id(p) in sparse_ids:
                if self.dense_end is None:
                    self.dense_end = off
                V, D = p.shape
                t = {'param': p, 'begin': off, 'V': V, 'D': D,
                     'bitmap': torch.zeros((V + 31) // 32, device=dev, dtype=torch.int32),
                     'active': torch.zeros(V, device=dev, dtype=torch.int32),
                     'count': torch.zeros(1, device=dev, dtype=torch.int32)}
                self.tables.append(t)
                p._xnrs_rows = t           # kernels.EmbeddingFn.backward marks the rows it scatters into
            off += n
        if self.dense_end is None:
            self.dense_end = total
        self.lr, self.betas, self.eps = lr, betas, eps
        self.step_count = 0
        self.graph_safe = graph_safe
        if graph_safe:      # step counter and bias corrections live on the device so a CUDA graph replays them
            self.step_dev = torch.zeros(1, device=dev, dtype=torch.int32)
            self.bc_dev = torch.zeros(2, device=dev, dtype=torch.float32)

    def _table_views(self, t):
        a, b = t['begin'], t['begin'] + t['V'] * t['D']
        return self.flat_p[a:b], self.flat_g[a:b], self.m[a:b], self.v[a:b]

    def zero_grad(self) -> None:
        self.g_store[:4 + self.dense_end].zero_()
        for t in self.tables:               # only active rows can hold a gradient
            K.call('xnrs_zero_rows', self._table_views(t)[1], t['V'], t['D'], t['active'], t['count'])

    def _check_views(self) -> None:
        """the parameters' .grad must still be the views into flat_g handed out at construction (e.g. a
        `model.zero_grad(set_to_none=True)` would silently detach them and the update would see zeros)"""
        base = self.flat_g.untyped_storage().data_ptr()
        for p in self.params:                  # a pointer compare per parameter (a middle one can be detached too)
            if p.grad is None or p.grad.untyped_storage().data_ptr() != base:
                raise RuntimeError('FlatAdam: a parameter gradient no longer lives in the flat gradient buffer — use '
                                   'optimizer.zero_grad() (not model.zero_grad(set_to_none=True)) between steps')

    def step(self, grad_scale: float = 1.0) -> None:
        self._check_views()
        self.step_count += 1
        bc = None
        if self.graph_safe:
            K.call('xnrs_adam_tick', self.step_dev, self.betas[0], self.betas[1], self.bc_dev)
            bc = self.bc_dev
        d = self.dense_end
        K.adam_step(self.flat_p[:d], self.flat_g[:d], self.m[:d], self.v[:d], self.lr, self.betas[0], self.betas[1], self.eps,
                    self.step_count, bc, grad_scale)
        for t in self.tables:
            p, g, m, v = self._table_views(t)
            K.call('xnrs_adam_rows', p, g, m, v, t['V'], t['D'], t['active'], t['count'], self.lr, self.betas[0], self.betas[1],
                   self.eps, self.step_count, bc, grad_scale)


def row_sparse_tables(model: nn.Module, min_rows: int = FlatAdam.SPARSE_MIN_ROWS):
    """the embedding tables big enough for the active-row Adam (user-id tables; category tables stay dense)"""
    return [m.weight for m in model.modules() if isinstance(m, nn.Embedding) and m.num_embeddings >= min_rows]


class RankingTrainer:
    """common part of the reference's BaseTrainer / RankingTrainer (training.py:24-44, 191-243)."""

    loss_kind = K.LOSS_MSE_RELU
    sparse_min_rows = FlatAdam.SPARSE_MIN_ROWS      # embedding tables at least this tall get the active-row Adam
    eval_act = 1            # activation that takes RAW scores to the ranked scores (catalogue evaluation): relu
    test_act = 0            # activation _test_step still has to apply to forward()'s output before the metrics

    def __init__(self, cfg, model: nn.Module, trainset=None, testset=None, graph_safe: bool = False):
        self.cfg = cfg if isinstance(cfg, Cfg) else Cfg(cfg)
        self.model = model
        self.device = torch.device(self.cfg.get('device', 'cuda:0'))
        self.model.to(self.device)
        self.optimizer = FlatAdam(self.model.parameters(), lr=float(self.cfg.get('lr', 1e-4)), graph_safe=graph_safe,
                                  row_sparse=row_sparse_tables(self.model, self.sparse_min_rows))
        self._init_loss()
        self.current_train_step = 0

    # -- loss hooks ---------------------------------------------------------------------------------
    def _init_loss(self):
        kind = self.loss_kind

        def loss(scores: torch.Tensor, target: torch.Tensor, weight: Optional[torch.Tensor] = None):
            """self.L(s, t[, w]) of the reference on RAW scores (B,N,1): the activation is fused in."""
            B, N = scores.shape[0], scores.shape[1]
            l, _, _ = K.ScoreLossFn.apply(None, K._f32(scores).reshape(B, N), _flat(target), _flat(weight), kind)
            return l
        self.L = loss

    def raw_scores(self, batch) -> torch.Tensor:
        batch_to_device(batch, self.device)
        return self.model.forward(batch)

    def forward(self, batch) -> torch.Tensor:
        """scores with the trainer's output activation (training.py:388-392 applies relu)."""
        s = self.raw_scores(batch)
        if self.loss_kind == K.LOSS_MSE_RELU:
            return K.ReluFn.apply(s)
        if self.loss_kind == K.LOSS_BCE_SIGMOID:
            return K.SigmoidFn.apply(s)
        return s

    def _embeddings(self, batch):
        """(u (B,T), c (B,N,T)) from ONE forward pass; None if the model cannot return embeddings (NPA)."""
        batch_to_device(batch, self.device)
        rec = getattr(self.model, 'rec_model', None)
        if not isinstance(rec, DotScoring) or rec.normalize:      # the fused scorer + loss kernel is the plain dot product
            return None
        try:
            r, u, c = self.model(batch, return_embeddings=True)
        except TypeError:
            return None
        B, N, T = c.shape
        return K._f32(u).reshape(B, T), K._f32(c)

    use_weights = True      # batch['weights'] enters the loss (training.py:101-105); the contrastive trainer ignores it

    def rec_loss(self, batch):
        """-> (loss_rec, preds (B,N,1), user_emb (B,T) or None): scorer and loss fused when the model exposes u, c."""
        t = _flat(batch['targets'].to(self.device))
        w = _flat(batch['weights'].to(self.device)) if (self.use_weights and 'weights' in batch) else None
        uc = self._embeddings(batch)
        if uc is not None:
            u, c = uc
            loss, preds, _ = K.ScoreLossFn.apply(u, c, t, w, self.loss_kind)
            return loss, preds.unsqueeze(-1), u
        s = self.model.forward(batch)
        B, N = s.shape[0], s.shape[1]
        loss, preds, _ = K.ScoreLossFn.apply(None, K._f32(s).reshape(B, N), t, w, self.loss_kind)
        return loss, preds.unsqueeze(-1), None

    def prefetch(self, batch: dict, after=None) -> bool:
        """input-pipeline hook (no reference counterpart: the reference's DataLoader has num_workers 0): prepare the id
        plumbing of an upcoming index batch on a side stream while the current step runs"""
        fn = getattr(self.model, 'prefetch', None)
        return bool(fn(batch, after)) if fn is not None else False

    def _train_step(self, batch: dict) -> dict:
        """BaseTrainer._train_step (training.py:97-112)."""
        self.optimizer.zero_grad()
        loss, preds, _ = self.rec_loss(batch)
        with K.direct_grads():
            loss.backward()
        self.optimizer.step()
        self.current_train_step += 1
        return {'loss': loss.detach(), 'logits': preds}

    @torch.no_grad()
    def _test_step(self, batch: dict) -> dict:
        """RankingTrainer._test_step (training.py:194-243) for one impression (B=1, all candidates): the metrics rank
        ``self.forward(batch)``; BCELogitsRankingTrainer ranks sigmoid(forward) (training.py:345-373, ``test_act`` 2)."""
        t = batch['targets'].to(self.device)
        loss, preds, _ = self.rec_loss(batch)
        scores = K._f32(preds).reshape(-1).clone()
        n = scores.numel()
        offsets = torch.tensor([0, n], device=self.device, dtype=torch.int64)
        _, m = K.eval_impressions(None, None, None, offsets, _flat(t), act=self.test_act, scores=scores)
        m = m[0].tolist()
        return {'auc': m[0], 'rr': m[1], 'ndcg@5': m[2], 'ndcg@10': m[3], 'ctr@1': m[4], 'ctr@10': m[5],
                'scores': scores, 'targets': t.reshape(-1), 'loss': loss}


def _flat(t: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
    return None if t is None else K._f32(t).reshape(-1)


class MSERankingTrainer(RankingTrainer):
    """training.py:376-392 — MSE(relu(score), target)."""
    loss_kind = K.LOSS_MSE_RELU
    eval_act = 1


class BCERankingTrainer(RankingTrainer):
    """training.py:324-331 — nn.BCELoss on sigmoid(score); forward() returns the sigmoid scores, which are also ranked."""
    loss_kind = K.LOSS_BCE_SIGMOID
    eval_act = 2


class BCELogitsRankingTrainer(RankingTrainer):
    """training.py:334-373 — BCE with logits; forward() returns raw scores, evaluation ranks sigmoid(score)."""
    loss_kind = K.LOSS_BCE_LOGITS
    eval_act = 2
    test_act = 2


class ContrastiveRankingTrainer(MSERankingTrainer):
    """training.py:395-472 — MSE(relu(score), target) + lambda * supervised InfoNCE(user embeddings, theme)."""

    use_weights = False     # its _train_step calls self.L(preds, targets) without the weights (training.py:405-407)

    def __init__(self, cfg, model, trainset=None, testset=None, share_user_forward: bool = True,
                 graph_safe: bool = False):
        super().__init__(cfg, model, trainset, testset, graph_safe=graph_safe)
        self.share_user_forward = share_user_forward

    def _init_loss(self):
        super()._init_loss()
        self.temperature = float(self.cfg.get('contrastive_temperature', 0.1))
        self.lambda_cl = float(self.cfg.get('contrastive_lambda', 0.1))

    def _compute_contrastive_loss(self, embeddings: torch.Tensor, labels: torch.Tensor) -> torch.Tensor:
        if embeddings.dim() > 2:
            embeddings = embeddings.reshape(embeddings.shape[0], -1)       # NAML's (B,1,D), training.py:443-444
        return K.InfoNCEFn.apply(K._f32(embeddings), labels.to(torch.int32).contiguous(), self.temperature)

    def losses(self, batch: dict):
        """-> (total, loss_rec, loss_cl, preds)"""
        loss_rec, preds, u = self.rec_loss(batch)
        if u is None or not self.share_user_forward:
            u = self.model.get_user_embeddings(batch)
        labels = theme_labels(batch['main_theme'], self.device)
        loss_cl = self._compute_contrastive_loss(u, labels)
        total = loss_rec + self.lambda_cl * loss_cl          # two device scalars (training.py:422)
        return total, loss_rec, loss_cl, preds

    def _train_step(self, batch: dict) -> dict:
        self.optimizer.zero_grad()
        total, loss_rec, loss_cl, preds = self.losses(batch)
        with K.direct_grads():
            total.backward()
        self.optimizer.step()
        self.current_train_step += 1
        return {'loss': total.detach(), 'loss_rec': loss_rec.detach(), 'loss_cl': loss_cl.detach(), 'logits': preds}
