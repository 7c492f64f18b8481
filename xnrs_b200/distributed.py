"""Data-parallel plumbing (one process per GPU, torch.distributed over NCCL/NVLink; gloo in the CPU tests).

The reference is single-process; what shards naturally is the impression batch (§8(e)):
  * every rank runs the same step on B/world impressions with replicated tables and parameters;
  * the supervised InfoNCE term couples users across the batch, so user embeddings and labels are
    all-gathered and every rank evaluates the GLOBAL-batch loss for its own anchors (reference semantics at
    the global batch size); the backward all-reduces the (B,E) embedding gradient;
  * all parameter gradients live in FlatAdam's single flat buffer: ONE all-reduce per step, averaged by
    folding 1/world into the Adam kernel's grad_scale.
No collective is issued on the data path itself (gathers/encoders/scorer are rank-local).

On GPUs of one node the two exchange steps of the InfoNCE term do not go through NCCL: `PeerExchange` maps one buffer per
rank into every rank (CUDA VMM peer mappings, set up by torch symmetric memory) and two kernels of this library fuse the
exchange with the work around it over NVLink — normalise + all-gather (each rank stores its finished rows into all gathered
buffers) and reduce-scatter + normalisation backward (each rank pulls and sums the partial gradients of its rows);
csrc/peer.cu.  XNRS_PEER=0, a CPU / gloo group, or a failing rendezvous keep the NCCL collectives; `infonce_exchange` on the
trainer says which path runs.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import kernels as K


def world() -> int:
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def rank() -> int:
    return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


def shard_range(n: int, r: int, w: int):
    """contiguous balanced slice [lo, hi) of n units for rank r of w"""
    base, rem = divmod(n, w)
    lo = r * base + min(r, rem)
    return lo, lo + base + (1 if r < rem else 0)


class PeerExchange:
    """the symmetric buffer of one (local batch, embedding size): flags | gathered ehat, 1/norm, labels | d_ehat"""

    def __init__(self, Ba: int, E: int, dev: torch.device):
        import torch.distributed._symmetric_memory as symm_mem
        w = world()
        Bk = w * Ba
        self.Ba, self.E, self.Bk = Ba, E, Bk

        def up(n):                                                  # regions start on 256-byte boundaries
            return (n + 255) // 256 * 256
        self.off_flags_ag, self.off_flags_rs = 0, 256               # 64 uint32 slots each
        self.off_ehat = 512
        self.off_inv = self.off_ehat + up(Bk * E * 4)
        self.off_lab = self.off_inv + up(Bk * 4)
        self.off_dehat = self.off_lab + up(Bk * 4)
        total = self.off_dehat + up(Bk * E * 4)
        self.buf = symm_mem.empty(total // 4, dtype=torch.float32, device=dev)
        self.buf.zero_()
        self.hdl = symm_mem.rendezvous(self.buf, dist.group.WORLD)
        self.ptrs = torch.tensor([int(p) for p in self.hdl.buffer_ptrs], dtype=torch.int64, device=dev)
        self.ctl = torch.zeros(8, dtype=torch.int32, device=dev)

        def view(off, n):
            return self.buf[off // 4: off // 4 + n]
        self.ehat_all = view(self.off_ehat, Bk * E).view(Bk, E)
        self.inv_all = view(self.off_inv, Bk)
        self.lab_all = view(self.off_lab, Bk).view(torch.int32)
        self.d_ehat = view(self.off_dehat, Bk * E).view(Bk, E)
        torch.cuda.synchronize(dev)
        dist.barrier()                                              # every rank's flags are zero before the first signal

    def check(self) -> None:
        """raise if a kernel gave up waiting for a peer (reads 4 bytes back: call it outside the hot loop)"""
        if int(self.ctl[4]) != 0:
            raise RuntimeError('xnrs_b200 peer exchange: a rank never arrived at an InfoNCE exchange step (ranks out of step?)')


_peer_cache = {}
_peer_broken = []


def peer_exchange(Ba: int, E: int, dev: torch.device):
    """the PeerExchange for this shape, or None when the NCCL collectives are to be used.  The first call for a shape is a
    collective (rendezvous + barrier) and must happen on every rank, outside CUDA-graph capture."""
    import os
    if (os.environ.get('XNRS_PEER', '1') == '0' or dev.type != 'cuda' or world() == 1 or world() > 64 or E % 4 or E > 256
            or dist.get_backend() != 'nccl' or _peer_broken):
        return None
    key = (Ba, E, dev.index)
    px = _peer_cache.get(key)
    if px is None:
        if torch.cuda.is_current_stream_capturing():
            return None
        try:
            px = PeerExchange(Ba, E, dev)
        except Exception as exc:                                    # no peer access / symmetric memory unsupported on this box
            import sys
            print(f'xnrs_b200: peer-memory InfoNCE exchange unavailable ({type(exc).__name__}: {exc}); using NCCL collectives',
                  file=sys.stderr)
            _peer_broken.append(True)
            return None
        _peer_cache[key] = px
    return px


class DistInfoNCEFn(torch.autograd.Function):
    """global-batch supervised InfoNCE (training.py:433-472) with rank-local anchors."""

    @staticmethod
    def forward(ctx, emb, labels, temperature, aux=None):
        """aux: optional float buffer that is SUMMED over ranks later in the step (FlatAdam.g_aux, part of the gradient
        all-reduce).  With it the forward needs no collective of its own besides the all-gather: the normaliser (anchors with
        a positive) is a function of the gathered labels, computed locally, and the rank-local share of the loss value is
        left in aux[0] for the caller to read after the gradient all-reduce.  Without it the two statistics are all-reduced
        here."""
        w, r = world(), rank()
        Ba, E = emb.shape
        dev = emb.device
        Bk, row0 = w * Ba, r * Ba
        # (a buffer is reused by the next step: what keeps a fast rank from overwriting it early is the gradient all-reduce that
        # follows every backward, so forward-only evaluations of the loss take the NCCL path)
        px = peer_exchange(Ba, E, dev) if ctx.needs_input_grad[0] else None
        ctx.px = px
        if px is not None:
            # normalise this rank's rows and store them (+ 1/norm, labels) into every rank's gathered arrays over NVLink
            K.call('xnrs_peer_normalize_allgather', emb, labels.to(torch.int32).contiguous(), Ba, E, r, w, px.ptrs,
                   px.off_flags_ag, px.off_ehat, px.off_inv, px.off_lab, px.ctl)
            ehat, inv_norm, lab_all = px.ehat_all, px.inv_all, px.lab_all
        else:
            # ONE all-gather: the int32 labels ride along as an extra (bit-cast) column of the embedding block
            packed = torch.empty((Ba, E + 1), device=dev, dtype=torch.float32)
            packed[:, :E] = emb
            packed[:, E] = labels.to(torch.int32).contiguous().view(torch.float32)
            packed_all = torch.empty((w * Ba, E + 1), device=dev, dtype=torch.float32)
            dist.all_gather_into_tensor(packed_all, packed)
            emb_all = packed_all[:, :E].contiguous()
            lab_all = packed_all[:, E].contiguous().view(torch.int32)
            ehat = torch.empty_like(emb_all)
            inv_norm = torch.empty(Bk, device=dev, dtype=torch.float32)
            K.call('xnrs_infonce_normalize', emb_all, Bk, E, ehat, inv_norm)
        ehat_a = ehat[row0:row0 + Ba]
        sim = K.gemm(ehat_a, ehat, trans_b=True)
        work = torch.zeros(4, device=dev, dtype=torch.float32)
        stats = work[:2]
        K.call('xnrs_infonce_rows', sim, lab_all, Ba, Bk, row0, temperature, stats)
        loss = torch.empty(1, device=dev, dtype=torch.float32)
        if aux is None:
            dist.all_reduce(stats)
            K.call('xnrs_infonce_finalize', stats, loss)
        else:
            K.call('xnrs_infonce_count', lab_all, Bk, work[2:], work[1:2])    # stats[1] <- the global count
            K.call('xnrs_infonce_finalize', stats, loss)                     # this rank's share: sum over ranks = the loss
            aux[:1].copy_(loss)
        ctx.save_for_backward(ehat, inv_norm, sim, stats)
        ctx.dims = (Ba, Bk, E, row0, w)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, g):
        ehat, inv_norm, G, stats = ctx.saved_tensors
        Ba, Bk, E, row0, w = ctx.dims
        ehat_a = ehat[row0:row0 + Ba]
        px = ctx.px
        if px is not None:
            # the gradient w.r.t. all Bk normalised rows goes into this rank's symmetric buffer; every rank then pulls and
            # sums the W partial blocks of ITS rows and applies the normalisation backward in the same kernel
            K.gemm(G, ehat_a, trans_a=True, out=px.d_ehat)
            K.gemm(G, ehat, out=px.d_ehat[row0:row0 + Ba], accumulate=True)
            d_emb = torch.empty((Ba, E), device=ehat.device, dtype=torch.float32)
            K.call('xnrs_peer_reduce_scatter_normalize_bwd', px.ptrs, px.off_flags_rs, px.off_dehat, Ba, E, rank(), w,
                   ehat_a, inv_norm[row0:row0 + Ba], stats, float(w), K._f32(g).reshape(1), d_emb, px.ctl)
            return d_emb, None, None, None
        d_ehat = K.gemm(G, ehat_a, trans_a=True)                              # key side, all Bk rows
        K.gemm(G, ehat, out=d_ehat[row0:row0 + Ba], accumulate=True)          # anchor side, local rows
        # every rank only needs the summed gradient of ITS rows: reduce-scatter (half the bytes of an all-reduce)
        d_loc = torch.empty((Ba, E), device=ehat.device, dtype=torch.float32)
        if dist.get_backend() == 'gloo':                                      # CPU tests: gloo has no reduce-scatter
            dist.all_reduce(d_ehat)
            d_loc.copy_(d_ehat[row0:row0 + Ba])
        else:
            dist.reduce_scatter_tensor(d_loc, d_ehat)
        d_a = torch.empty_like(d_loc)
        K.call('xnrs_infonce_normalize_bwd', d_loc, ehat_a.contiguous(), inv_norm[row0:row0 + Ba].contiguous(), stats, 1.0, Ba, E, d_a)
        # parameter gradients are averaged over ranks afterwards; this term is already the gradient of the
        # GLOBAL loss, so pre-multiply by world to survive the averaging
        d_emb = torch.empty((Ba, E), device=ehat.device, dtype=torch.float32)
        K.call('xnrs_axpby', Ba * E, float(w), K._f32(g).reshape(1), d_a, 0.0, d_emb)
        return d_emb, None, None, None


class DataParallelTrainer:
    """wraps a ContrastiveRankingTrainer / MSERankingTrainer for one-process-per-GPU data parallelism."""

    def __init__(self, trainer):
        self.trainer = trainer
        self.world, self.rank = world(), rank()
        self.sparse_tables = True            # exchange big embedding-table gradients as (ids, rows); False: dense all-reduce
        if self.world > 1:
            dist.broadcast(trainer.optimizer.flat_p, src=0)                  # identical replicas
            if hasattr(trainer, '_compute_contrastive_loss'):
                temp, aux = trainer.temperature, getattr(trainer.optimizer, 'g_aux', None)
                self.cl_aux = aux
                trainer._compute_contrastive_loss = (
                    lambda e, l: DistInfoNCEFn.apply(K._f32(e.reshape(e.shape[0], -1)), l, temp, aux))

    @property
    def infonce_exchange(self) -> str:
        """which path the InfoNCE exchange steps of the steps run so far took"""
        if self.world == 1 or not hasattr(self.trainer, '_compute_contrastive_loss'):
            return 'none'
        return 'peer memory (NVLink P2P kernels, csrc/peer.cu)' if _peer_cache else 'nccl'

    def check_peers(self) -> None:
        for px in _peer_cache.values():
            px.check()

    # tables with at least this many rows exchange their gradient as (ids, rows): LSTUR / NPA user tables (703 790 rows)
    SPARSE_MIN_ROWS = 100_000

    def _reduce_gradients(self, log) -> None:
        """sum the flat gradient buffer over ranks.  Row-sparse tables logged by EmbeddingFn.backward are left out of the
        dense all-reduce; their touched rows travel as one all-gather of (ids | rows) per table and are scatter-added
        locally (B x (D+1) floats instead of V x D: 0.6 MB instead of 383 MB per step for the LSTUR user table)."""
        opt = self.trainer.optimizer
        flat = getattr(opt, 'g_store', opt.flat_g)       # [4 auxiliary floats | gradient]: one buffer, one all-reduce
        head = flat.numel() - opt.flat_g.numel()
        tables = {}
        for weight, idx, dy, pad in log:
            if weight.shape[0] >= self.SPARSE_MIN_ROWS and id(weight) in opt.ranges:
                tables.setdefault(id(weight), (weight, []))[1].append((idx, dy, pad))
        self.last_sparse_tables = len(tables)
        if not tables:
            dist.all_reduce(flat)                                            # the one gradient bucket
            return
        cuts = sorted((a + head, b + head) for a, b in (opt.ranges[k] for k in tables))
        lo = 0
        for a, b in cuts + [(flat.numel(), flat.numel())]:                   # dense all-reduce of everything between the tables
            if a > lo:
                dist.all_reduce(flat[lo:a])
            lo = max(lo, b)
        for weight, recs in tables.values():
            V, D = weight.shape
            idx = torch.cat([r[0] for r in recs])
            rows = torch.cat([r[1].reshape(-1, D) for r in recs])
            pad = recs[0][2]
            n = torch.tensor([idx.numel()], device=flat.device, dtype=torch.int64)
            dist.all_reduce(n, op=dist.ReduceOp.MAX)
            n = int(n)
            packed = torch.zeros((n, D + 1), device=flat.device, dtype=torch.float32)
            packed[:idx.numel(), :D] = rows
            ids = torch.full((n,), -1, device=flat.device, dtype=torch.int32)   # -1 = padding: skipped by the scatter kernel
            ids[:idx.numel()] = idx
            packed[:, D] = ids.view(torch.float32)
            gathered = torch.empty((self.world * n, D + 1), device=flat.device, dtype=torch.float32)
            dist.all_gather_into_tensor(gathered, packed)
            keep = torch.ones(self.world * n, device=flat.device, dtype=torch.bool)
            keep[self.rank * n:(self.rank + 1) * n] = False                   # this rank's own rows are already in its gradient
            remote = gathered[keep]
            r_ids = remote[:, D].contiguous().view(torch.int32)
            r_rows = remote[:, :D].contiguous()
            K.call('xnrs_scatter_add_rows', weight.grad, V, D, r_ids, r_ids.numel(), r_rows, D, pad)
            K.mark_active_rows(weight, r_ids, pad)                             # remote rows become active in this replica too
            # Every rank now holds the same sums up to fp32 summation order (own rows first, atomics).  Replicas must stay
            # BIT-identical (a dense all-reduce guarantees that), so rank 0's values of the touched rows are made
            # authoritative: gather them, broadcast, write back (byte movement; W*B x D floats).
            all_ids = gathered[:, D].contiguous().view(torch.int32)
            valid = (all_ids >= 0) & (all_ids < V) & (all_ids != pad)
            touched = all_ids[valid].contiguous()
            vals = K.gather_rows(weight.grad, touched)
            dist.broadcast(vals, src=0)
            weight.grad.index_copy_(0, touched.long(), vals)

    def prefetch(self, batch: dict, after=None) -> bool:
        """input-pipeline hook: prepare the id plumbing of the NEXT batch while the current step runs (models that support it)"""
        fn = getattr(self.trainer.model, 'prefetch', None)
        return bool(fn(batch, after)) if fn is not None else False

    def _check_equal_shards(self, batch: dict) -> None:
        """the packed all-gather / reduce-scatter of the InfoNCE term and the flat 1/world gradient average need the SAME
        number of impressions on every rank; a ragged split (B % world != 0) would hang NCCL or mis-weight the loss, so it is
        rejected loudly on first use and whenever the local batch size changes"""
        n = int(batch['targets'].shape[0])
        if getattr(self, '_checked_b', None) == n:
            return
        dev = self.trainer.device
        mm = torch.tensor([n, -n], device=dev, dtype=torch.int64)
        dist.all_reduce(mm, op=dist.ReduceOp.MAX)
        if int(mm[0]) != -int(mm[1]):
            raise RuntimeError(f'data-parallel train_step needs the same number of impressions on every rank '
                               f'(this rank has {n}; ranks hold between {-int(mm[1])} and {int(mm[0])}): '
                               f'make the global batch a multiple of the world size')
        self._checked_b = n

    def train_step(self, batch: dict) -> dict:
        tr = self.trainer
        if self.world > 1:
            self._check_equal_shards(batch)
        tr.optimizer.zero_grad()
        if hasattr(tr, 'losses'):
            total, loss_rec, loss_cl, preds = tr.losses(batch)
            out = {'loss': total.detach(), 'loss_rec': loss_rec.detach(), 'loss_cl': loss_cl.detach(), 'logits': preds}
        else:
            total, preds, _ = tr.rec_loss(batch)
            out = {'loss': total.detach(), 'logits': preds}
        if self.world > 1 and not self.sparse_tables and getattr(tr.optimizer, 'tables', None):
            raise RuntimeError('the optimiser updates its big embedding tables row-sparsely: their gradients must be exchanged as '
                               '(ids, rows) (sparse_tables=True), a dense all-reduce would deliver rows it does not know about')
        if self.world > 1 and self.sparse_tables:
            K.sparse_grad_log = []
        try:
            with K.direct_grads():
                total.backward()
            log = K.sparse_grad_log or []
        finally:
            K.sparse_grad_log = None
        if self.world > 1:
            self._reduce_gradients(log)
            if getattr(self, 'cl_aux', None) is not None and 'loss_cl' in out:
                # the InfoNCE value was summed over ranks with the gradient (each rank contributed its anchors' share)
                out['loss_cl'] = self.cl_aux[0].clone()
                out['loss'] = out['loss_rec'] + tr.lambda_cl * out['loss_cl']
        tr.optimizer.step(grad_scale=1.0 / self.world)
        tr.current_train_step += 1
        return out
