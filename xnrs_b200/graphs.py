"""CUDA-graph replay of a whole training step (reference step: xnrs/training.py:97-112,402-431).

One CL step is ~50 kernels of 5-400 us; launched one by one from Python the GPU idles between them and, data-parallel, the
ranks drift apart.  ``GraphedStep`` captures zero_grad -> forward -> losses -> backward -> gradient collectives -> Adam into
ONE CUDA graph per shape bucket and replays it with a single launch.

What makes the step capturable:
  * the id plumbing is a pure function of the batch's ids and lives OUTSIDE the graph (`TitlePlan`, device kernels, usually
    prefetched one step ahead); its arrays are padded to rounded counts, so steps fall into a handful of shape buckets
    (distinct-article count U' x real-token count T'); the graph reads the plan from static buffers refreshed by small
    device-to-device copies before each replay;
  * Adam's step counter / bias corrections live on the device (`FlatAdam(graph_safe=True)`, xnrs_adam_tick);
  * nothing inside the step reads a value back to the host.
The first step of a bucket runs eagerly (it also is the warm-up the capture needs), the second captures, all later ones
replay.  Models whose step needs host decisions (seeded dropout: NRMS, LSTUR; row-sparse table exchange; NAML's
article-level unique) are refused — they keep the eager path.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

from . import kernels as K
from .data import IndexedTitles, TitlePlan, plan_titles
from .models.components import ParentRec, TextEncoder, UserEncoder, merge_sides


class GraphedStep:
    def __init__(self, dp_trainer, max_graphs: int = 64):
        """dp_trainer: distributed.DataParallelTrainer around a ContrastiveRankingTrainer / MSERankingTrainer whose optimizer
        was built with graph_safe=True"""
        self.dp = dp_trainer
        self.tr = dp_trainer.trainer
        model = self.tr.model
        if not self.tr.optimizer.graph_safe:
            raise RuntimeError('GraphedStep needs FlatAdam(graph_safe=True) (trainer(..., graph_safe=True))')
        if not (isinstance(model, ParentRec) and isinstance(model.news_encoder, TextEncoder) and model.news_encoder.att is None
                and isinstance(model.user_encoder, UserEncoder) and model.user_encoder.att is None):
            raise RuntimeError('GraphedStep covers the additive-pooling ParentRec models (StandardRec / CL); models with seeded '
                               'dropout or host-side index decisions keep the eager step')
        for m in model.modules():
            if isinstance(m, torch.nn.Dropout) and m.p > 0 and model.training:
                raise RuntimeError('GraphedStep: active dropout draws its seed on the host every step')
        self.model = model
        self.graphs: Dict[tuple, dict] = {}
        self.seen: Dict[tuple, int] = {}
        self.pool = None
        self.max_graphs = max_graphs
        self.replays = self.captures = self.eager_steps = self.replayed_kernels = 0

    # ---- plan of a batch (prefetched or computed here) -----------------------------------------------------------
    def _plan(self, batch):
        hist = batch['user_features']['history'][self.model.text_feature]
        cand = batch['candidate_features'][self.model.text_feature]
        merged = merge_sides(hist, cand)
        if merged is None:
            raise RuntimeError('GraphedStep needs index batches (IndexedTitles on both sides of one TitleStore)')
        titles, b, nh, nc = merged
        enc = self.model.news_encoder
        dedup, ragged = enc.plan_kind(titles.news_ids.numel(), titles.distinct)
        plan = titles.plan
        if plan is None or plan.dedup != dedup or plan.ragged != ragged:
            plan = plan_titles(titles.store, titles.news_ids.to(self.tr.device).reshape(-1), dedup, ragged)
        plan.acquire()
        return titles, plan, b, nh, nc

    @staticmethod
    def _plan_tensors(plan: TitlePlan):
        return {k: getattr(plan, k) for k in ('uniq', 'inv', 'rows', 'seg', 'mask', 'cm', 'tix') if getattr(plan, k) is not None}

    def _make_static(self, titles, plan, b, nh, nc, batch):
        dev = self.tr.device
        st = {k: torch.empty_like(v) for k, v in self._plan_tensors(plan).items()}
        splan = TitlePlan(st['uniq'], st.get('inv'), st['rows'], st.get('seg'), st.get('mask'), st['cm'], plan.ragged, plan.dedup,
                          tix=st.get('tix'), seq_len=plan.seq_len, n_titles=plan.n_titles, n_rows=plan.n_rows)
        ids_h = torch.zeros((b, nh), device=dev, dtype=torch.int32)
        ids_c = torch.zeros((b, nc), device=dev, dtype=torch.int32)
        hist, cand = IndexedTitles(titles.store, ids_h), IndexedTitles(titles.store, ids_c)
        merged_ids = torch.zeros((1, b * (nh + nc)), device=dev, dtype=torch.int32)    # only its shape is read (the plan is given)
        targets = torch.empty_like(batch['targets'].to(dev), dtype=torch.float32)
        sb = {'user_features': {'history': {self.model.text_feature: hist}, 'other': {}},
              'candidate_features': {self.model.text_feature: cand}, 'targets': targets}
        if 'main_theme' in batch:
            sb['main_theme'] = torch.zeros(b, device=dev, dtype=torch.int32)
        return {'plan_t': st, 'plan': splan, 'batch': sb, 'hist': hist, 'cand': cand,
                'merged': (IndexedTitles(titles.store, merged_ids), b, nh, nc), 'graph': None, 'out': None}

    def _refresh(self, g, plan, batch):
        for k, v in self._plan_tensors(plan).items():
            g['plan_t'][k].copy_(v, non_blocking=True)
        g['batch']['targets'].copy_(batch['targets'], non_blocking=True)
        if 'main_theme' in g['batch']:
            from .training import theme_labels
            g['batch']['main_theme'].copy_(theme_labels(batch['main_theme'], self.tr.device), non_blocking=True)

    def _arm(self, g):
        """hand the static plan to the next forward of the static batch (merge_sides consumes it once)"""
        titles, b, nh, nc = g['merged']
        titles.plan = g['plan']
        g['hist']._merged = (g['cand'], (titles, b, nh, nc))

    # ---- the step --------------------------------------------------------------------------------------------------
    def step(self, batch: dict) -> dict:
        """== DataParallelTrainer.train_step(batch); the returned tensors are static buffers, valid until the next step"""
        titles, plan, b, nh, nc = self._plan(batch)
        key = (plan.uniq.numel(), plan.rows.numel(), b, nh, nc, plan.dedup, plan.ragged)
        n_seen = self.seen.get(key, 0)
        self.seen[key] = n_seen + 1
        g = self.graphs.get(key)
        if g is None:
            if len(self.graphs) >= self.max_graphs:         # shape-bucket explosion: stay eager rather than hoard graphs
                return self._eager(batch, titles, plan, b, nh, nc)
            g = self.graphs[key] = self._make_static(titles, plan, b, nh, nc, batch)
        self._refresh(g, plan, batch)
        if g['graph'] is None:
            if n_seen == 0:                                 # first visit of this bucket: eager (and the capture's warm-up)
                self._arm(g)
                self.eager_steps += 1
                return self.dp.train_step(g['batch'])
            graph = torch.cuda.CUDAGraph()
            self._arm(g)
            torch.cuda.synchronize()
            counters = (self.tr.optimizer.step_count, self.tr.current_train_step)
            n0 = K.launch_count()
            with torch.cuda.graph(graph, pool=self.pool):
                g['out'] = self.dp.train_step(g['batch'])
            g['kernels'] = K.launch_count() - n0            # this library's kernel nodes in the graph (one replay runs them all)
            self.tr.optimizer.step_count, self.tr.current_train_step = counters     # capturing executed nothing
            if self.pool is None:
                self.pool = graph.pool()
            g['graph'] = graph
            self.captures += 1
        g['graph'].replay()
        self.replays += 1
        self.replayed_kernels += g['kernels']
        self.tr.optimizer.step_count += 1
        self.tr.current_train_step += 1
        return g['out']

    def _eager(self, batch, titles, plan, b, nh, nc):
        titles.plan = plan
        hist = batch['user_features']['history'][self.model.text_feature]
        hist._merged = (batch['candidate_features'][self.model.text_feature], (titles, b, nh, nc))
        self.eager_steps += 1
        return self.dp.train_step(batch)

    def prefetch(self, batch: dict, after=None) -> bool:
        return self.dp.prefetch(batch, after)
