"""CPU stand-in for the C-ABI entry points — TEST INFRASTRUCTURE for the `-m "not gpu"` suite only.

It lets the *host logic* (argument marshalling, autograd glue, module wiring, state_dict layout, the
trainer hooks, the data-parallel plumbing over gloo) be exercised in a container without a GPU, by
replacing ``xnrs_b200.kernels.call`` with an executable specification of every entry point's contract
written in plain torch.  It is never imported by the package, bench.py or the GPU tests: the product has
no CPU path, and `tests/test_gpu_*.py` run the real kernels through the real ``call``.
"""
import math

import torch

from xnrs_b200 import kernels as K


def _rows(x, rows):
    return x if rows is None else x[rows.long()]


def _kf(keep, p, shape, seed=None):
    if keep is not None:
        return keep.reshape(shape)
    if p > 0:
        if seed is None:
            raise NotImplementedError('emulator needs an explicit keep mask or a seed')
        # seeded draw, reproduced by the backward from the same seed (the kernels use a counter hash: same contract,
        # different stream — bit parity with a particular generator is not part of the ABI)
        g = torch.Generator().manual_seed(int(seed) % (2 ** 63))
        return (torch.rand(shape, generator=g) >= p).float()
    return torch.ones(shape)


def call(name, *a):
    fn = globals().get('_' + name)
    if fn is None:
        raise NotImplementedError(name)
    fn(*[x.t if isinstance(x, K._Strided) else x for x in a])


def _xnrs_expand_titles(title_tokens, n_news, S, news_ids, R, rows, mask):
    tok = title_tokens[news_ids.long()].reshape(-1)
    rows.copy_(tok)
    if mask is not None:
        mask.copy_((tok != 0).float())


def _xnrs_plan_dedup(ids, n, n_news, work, uniq, inv, counts):
    safe = torch.where((ids >= 0) & (ids < n_news), ids, torch.zeros_like(ids))
    u, iv = torch.unique(safe, return_inverse=True)
    uniq.zero_()
    uniq[:u.numel()] = u.to(torch.int32)
    inv.copy_(iv.to(torch.int32))
    counts[0] = u.numel()


def _xnrs_plan_ragged(title_tokens, n_news, S, uniq, cap, u_count, pad_rows, lens, seg, rows, tix, rows_cap, cm, counts):
    U = cap if u_count is None else int(u_count[0])
    tok = title_tokens[uniq[:U].long()]
    valid = tok != 0
    ln = valid.sum(1).to(torch.int32)
    lens.zero_()
    lens[:U] = ln
    seg[0] = 0
    seg[1:cap + 1] = torch.cumsum(lens[:cap], 0).to(torch.int32)
    T = int(seg[cap])
    rows[:T] = tok[valid]
    rows[T:min(T + pad_rows, rows_cap)] = 0
    if tix is not None:
        tix[:T] = torch.repeat_interleave(torch.arange(U, dtype=torch.int32), ln.long())
        tix[T:min(T + pad_rows, rows_cap)] = -1
    cm.copy_((lens[:cap] > 0).float())
    counts[1] = T


def _xnrs_gather_rows(table, V, D, rows, R, out, ld):
    out.copy_(table[rows.long()])


def _xnrs_scatter_add_rows(dtable, V, D, rows, R, dout, ld, skip):
    keep = (rows != skip) & (rows >= 0) & (rows < V)          # out-of-range ids (e.g. -1 padding) are skipped like the kernel does
    dtable.index_add_(0, rows[keep].long(), dout[keep])


def _xnrs_gemm(ta, tb, M, N, K_, A, lda, a_rows, B, ldb, b_rows, C, ldc, bias, act, aux, accumulate, split_k, prec):
    a = _rows(A, a_rows)
    b = _rows(B, b_rows)
    a = a.T if ta else a
    b = b.T if tb else b
    y = a @ b
    assert y.shape == (M, N) and a.shape[1] == K_
    if bias is not None:
        y = y + bias
    if act == K.ACT_RELU:
        y = torch.relu(y)
    elif act == K.ACT_TANH:
        y = torch.tanh(y)
    elif act == K.ACT_RELU_MASK:
        y = y * (aux > 0)
    if accumulate:
        C.add_(y)
    else:
        C.copy_(y)


def _xnrs_colsum(X, M, N, ldx, out):
    out.add_(X.sum(0))


def _xnrs_axpby(n, a, a_dev, x, b, y):
    aa = a * (float(a_dev) if a_dev is not None else 1.0)
    y.copy_(aa * x + (b * y if b != 0 else 0))


def _xnrs_relu(n, x, y):
    y.copy_(torch.relu(x))


def _xnrs_relu_bwd(n, y, dy, dx):
    dx.copy_(dy * (y > 0))


def _xnrs_transpose(inp, rows, cols, out):
    out.copy_(inp.reshape(rows, cols).T)


def _xnrs_dropout(n, x, keep, p, seed, y):
    y.copy_(x * _kf(keep, p, x.shape, seed) / (1 - p))


def _groups(R, L, seg):
    return [(r * L, L) for r in range(R)] if seg is None else [(int(seg[r]), int(seg[r + 1] - seg[r])) for r in range(R)]


def _pool_fwd(x, x_rows, mask, logits, seg, R, L, attn, pooled):
    xr = _rows(x, x_rows)
    af, lg = attn.reshape(-1), logits.reshape(-1)
    for r, (b, n) in enumerate(_groups(R, L, seg)):
        e = torch.exp(lg[b:b + n])
        if mask is not None:
            e = e * mask.reshape(-1)[b:b + n]
        a = e / (e.sum() + 1e-8)
        af[b:b + n] = a
        pooled[r] = (a[:, None] * xr[b:b + n]).sum(0)


def _pool_dlogit(x, x_rows, attn, d_pooled, d_attn, seg, R, L):
    xr = _rows(x, x_rows)
    af = attn.reshape(-1)
    dl = torch.zeros_like(af)
    for r, (b, n) in enumerate(_groups(R, L, seg)):
        da = xr[b:b + n] @ d_pooled[r]
        if d_attn is not None:
            da = da + d_attn.reshape(-1)[b:b + n]
        a = af[b:b + n]
        dl[b:b + n] = a * (da - (a * da).sum())
    return dl


def _pool_dx(attn, d_pooled, seg, R, L, d_x):
    af = attn.reshape(-1)
    for r, (b, n) in enumerate(_groups(R, L, seg)):
        d_x[b:b + n] = af[b:b + n, None] * d_pooled[r][None, :]


def _xnrs_addpool_fwd(x, x_rows, mask, hid, w2, b2, seg, R, L, F_, A, attn, pooled):
    _pool_fwd(x, x_rows, mask, hid @ w2 + b2, seg, R, L, attn, pooled)


def _xnrs_titlepool_fwd(x, ldx, x_rows, tix, seg, n_rows, R, F_, A, w1, b1, w2, b2, prec, hid, e, zsum, attn, pooled):
    xr = _rows(x, x_rows)[:n_rows]
    h = torch.tanh(xr @ w1.T + b1)
    hid.copy_(h)
    valid = tix[:n_rows] >= 0
    ev = torch.where(valid, torch.exp(h @ w2.reshape(-1) + b2), torch.zeros(n_rows))
    e.copy_(ev)
    t = tix[:n_rows].clamp_min(0).long()
    zsum.zero_()
    zsum.index_add_(0, t, ev)
    pooled.zero_()
    pooled.index_add_(0, t, ev[:, None] * xr)
    attn.copy_(torch.where(valid, ev / (zsum[t] + 1e-8), torch.zeros(n_rows)))
    pooled.div_((zsum + 1e-8)[:, None])


def _xnrs_cast_bf16(n, src, dst):
    dst.copy_(src.to(torch.bfloat16))


def _xnrs_gemm_bf16(ta, tb, M, N, K_, A, lda, a_rows, B, ldb, b_rows, C, ldc, c_bf16, bias, act, accumulate, split_k):
    a = _rows(A, a_rows).float()
    b = _rows(B, b_rows).float()
    y = (a.T if ta else a) @ (b.T if tb else b)
    if bias is not None:
        y = y + bias
    y = torch.relu(y) if act == K.ACT_RELU else (torch.tanh(y) if act == K.ACT_TANH else y)
    C.copy_((C.float() + y if accumulate else y).to(C.dtype))


def _xnrs_split_bf16(n, src, hi, lo, fp16=0):
    dt = torch.float16 if fp16 else torch.bfloat16
    v = src.clamp(-65504.0, 65504.0) if fp16 else src
    h = v.to(dt)
    hi.copy_(h)
    lo.copy_((v - h.float()).to(dt))


def _xnrs_gemm_bf16x3(ta, tb, M, N, K_, A_hi, A_lo, a_f16, lda, a_rows, B_hi, B_lo, b_f16, ldb, b_rows, C, ldc, bias, act, accumulate,
                      split_k):
    ah, al, bh, bl = (_rows(t, r).float() for t, r in ((A_hi, a_rows), (A_lo, a_rows), (B_hi, b_rows), (B_lo, b_rows)))
    op = (lambda m: m.T) if ta else (lambda m: m)
    oq = (lambda m: m.T) if tb else (lambda m: m)
    y = op(ah) @ oq(bh) + (op(al) @ oq(bh) + op(ah) @ oq(bl))          # hi hi + (lo hi + hi lo): the lo lo term is dropped
    if bias is not None:
        y = y + bias
    y = torch.relu(y) if act == K.ACT_RELU else (torch.tanh(y) if act == K.ACT_TANH else y)
    C.copy_(C + y if accumulate else y)


def _xnrs_titlepool_fwd_bf16x3(x_hi, x_lo, ldx, x_rows, tix, seg, n_rows, R, F_, A, w1_hi, w1_lo, planes_f16, b1, w2, b2, x_f32, ld_f32,
                               hid, e, zsum, attn, pooled):
    xr = _rows(x_f32, x_rows)[:n_rows]
    xh, xl = _rows(x_hi, x_rows)[:n_rows].float(), _rows(x_lo, x_rows)[:n_rows].float()
    wh, wl = w1_hi.float(), w1_lo.float()
    h = torch.tanh(xh @ wh.T + (xl @ wh.T + xh @ wl.T) + b1)
    hid.copy_(h)
    valid = tix[:n_rows] >= 0
    ev = torch.where(valid, torch.exp(h @ w2.reshape(-1) + b2), torch.zeros(n_rows))
    e.copy_(ev)
    t = tix[:n_rows].clamp_min(0).long()
    zsum.zero_()
    zsum.index_add_(0, t, ev)
    pooled.zero_()
    pooled.index_add_(0, t, ev[:, None] * xr)
    attn.copy_(torch.where(valid, ev / (zsum[t] + 1e-8), torch.zeros(n_rows)))
    pooled.div_((zsum + 1e-8)[:, None])


def _xnrs_addpool_bwd_split(x, x_rows, hid, w2, attn, d_pooled, seg, R, L, F_, A, n_rows, d_hid_hi, d_hid_lo, d_w2, d_b2, d_b1):
    d32 = torch.empty(hid.shape, dtype=torch.float32)
    _xnrs_addpool_bwd(x, x_rows, None, hid, w2, attn, d_pooled, None, seg, R, L, F_, A, n_rows, d32, d_w2, d_b2, None, d_b1)
    _xnrs_split_bf16(d32.numel(), d32, d_hid_hi, d_hid_lo, 0)


def _xnrs_titlepool_fwd_bf16(x, ldx, x_rows, tix, seg, n_rows, R, F_, A, w1, b1, w2, b2, hid, e, zsum, attn, pooled):
    h32 = torch.empty(hid.shape, dtype=torch.float32)
    _xnrs_titlepool_fwd(x.float(), ldx, x_rows, tix, seg, n_rows, R, F_, A, w1.float(), b1, w2, b2, 3, h32, e, zsum, attn, pooled)
    hid.copy_(h32.to(torch.bfloat16))


def _xnrs_addpool_bwd_bf16(x, x_rows, hid, w2, attn, d_pooled, seg, R, L, F_, A, n_rows, d_hid, d_w2, d_b2, d_b1):
    d32 = torch.empty(d_hid.shape, dtype=torch.float32)
    _xnrs_addpool_bwd(x.float(), x_rows, None, hid.float(), w2, attn, d_pooled, None, seg, R, L, F_, A, n_rows, d32, d_w2, d_b2,
                      None, None)
    d_hid.copy_(d32.to(torch.bfloat16))
    if d_b1 is not None:
        d_b1.add_(d_hid.float().sum(0))


def _xnrs_addpool_bwd(x, x_rows, mask, hid, w2, attn, d_pooled, d_attn, seg, R, L, F_, A, n_rows, d_hid, d_w2, d_b2, d_x, d_b1):
    dl = _pool_dlogit(x, x_rows, attn, d_pooled, d_attn, seg, R, L)
    d_hid.copy_(dl[:, None] * w2[None, :] * (1 - hid * hid))
    if d_b1 is not None:
        d_b1.add_(d_hid.sum(0))
    d_w2.add_((dl[:, None] * hid).sum(0))
    d_b2.add_(dl.sum())
    if d_x is not None:
        _pool_dx(attn, d_pooled, seg, R, L, d_x)


def _qrows(qh, seg, R, L, rpq):
    return torch.cat([qh[r // rpq][None, :].expand(n, -1) for r, (b, n) in enumerate(_groups(R, L, seg))]) \
        if R else qh[:0]


def _xnrs_perspool_fwd(x, x_rows, mask, hid, qh, seg, R, L, F_, A, rpq, attn, pooled):
    qr = _qrows(qh, seg, R, L, rpq)
    lg = torch.zeros(hid.shape[0])
    lg[:qr.shape[0]] = (hid[:qr.shape[0]] * qr).sum(-1)
    _pool_fwd(x, x_rows, mask, lg, seg, R, L, attn, pooled)


def _xnrs_perspool_bwd(x, x_rows, mask, hid, qh, attn, d_pooled, seg, R, L, F_, A, rpq, n_rows, d_hid, d_qh, d_x):
    dl = _pool_dlogit(x, x_rows, attn, d_pooled, None, seg, R, L)
    qr = _qrows(qh, seg, R, L, rpq)
    d_hid.zero_()                                            # rows past the last ragged group (plan padding) get 0
    d_hid[:qr.shape[0]] = dl[:qr.shape[0], None] * qr * (1 - hid[:qr.shape[0]] * hid[:qr.shape[0]])
    for r, (b, n) in enumerate(_groups(R, L, seg)):
        d_qh[r // rpq] += (dl[b:b + n, None] * hid[b:b + n]).sum(0)
    if d_x is not None:
        _pool_dx(attn, d_pooled, seg, R, L, d_x)


def _xnrs_meanpool_fwd(x, mask, R, L, F_, pooled):
    m = mask.reshape(R, L, 1)
    pooled.copy_((x.reshape(R, L, F_) * m).sum(1) / (m.sum(1) + 1e-8))


def _xnrs_meanpool_bwd(mask, d_pooled, R, L, F_, d_x):
    m = mask.reshape(R, L, 1)
    d_x.copy_((d_pooled.reshape(R, 1, F_) * m / (m.sum(1, keepdim=True) + 1e-8)).reshape(d_x.shape))


def _xnrs_tanh_bwd(n, y, dy, dx):
    dx.copy_(dy * (1 - y * y))


def _xnrs_add_scalar(n, x, b, y):
    y.copy_(x + b[0])


def _xnrs_collapse_mask(mask, R, L, out):
    out.copy_(mask.reshape(R, L).sum(1).clamp(0, 1))


def _att(q, k, v, mask, R, L, h, dk, keep, p, seed=None):
    qh, kh, vh = (t.reshape(R, L, h, dk).transpose(1, 2) for t in (q, k, v))
    s = qh @ kh.transpose(-1, -2) / math.sqrt(dk)
    if mask is not None:
        s = s.masked_fill(mask.reshape(R, 1, L, 1) == 0, -1e9)
    pr = torch.softmax(s, -1)
    return pr * _kf(keep, p, pr.shape, seed) / (1 - p), vh


def _xnrs_mha_fwd(q, k, v, ld, mask, R, L, h, dk, keep, p, seed, o, lse):
    pr, vh = _att(q, k, v, mask, R, L, h, dk, keep, p, seed)
    o.copy_((pr @ vh).transpose(1, 2).reshape(R * L, h * dk))


def _xnrs_mha_bwd(q, k, v, o, d_o, ld, mask, lse, R, L, h, dk, keep, p, seed, dq, dk_, dv):
    qq, kk, vv = (t.detach().clone().requires_grad_(True) for t in (q, k, v))
    with torch.enable_grad():
        pr, vh = _att(qq, kk, vv, mask, R, L, h, dk, keep, p, seed)
        out = (pr @ vh).transpose(1, 2).reshape(R * L, h * dk)
        g = torch.autograd.grad(out, (qq, kk, vv), d_o)
    dq.copy_(g[0])
    dk_.copy_(g[1])
    dv.copy_(g[2])


def _gru(gi, w_hh, b_hh, h0, lengths, B, L, Hd):
    h = torch.zeros(B, Hd) if h0 is None else h0
    hs, gates = [], []
    g3 = gi.reshape(B, L, 3 * Hd)
    for t in range(L):
        gh = h @ w_hh.T + b_hh
        r = torch.sigmoid(g3[:, t, :Hd] + gh[:, :Hd])
        z = torch.sigmoid(g3[:, t, Hd:2 * Hd] + gh[:, Hd:2 * Hd])
        n = torch.tanh(g3[:, t, 2 * Hd:] + r * gh[:, 2 * Hd:])
        live = (lengths > t).float().unsqueeze(1)
        hs.append(h)
        gates.append(torch.cat([r, z, n, gh[:, 2 * Hd:]], 1))
        h = live * ((1 - z) * n + z * h) + (1 - live) * h
    return h, torch.stack(hs, 1), torch.stack(gates, 1)


def _xnrs_gru_fwd(gi, w_hh_t, b_hh, h0, lengths, B, L, Hd, hs, gates, h_out):
    h, s, g = _gru(gi, w_hh_t.T, b_hh, h0, lengths, B, L, Hd)
    h_out.copy_(h)
    hs.copy_(s.reshape(B * L, Hd))
    gates.copy_(g)


def _xnrs_gru_bwd(d_h_out, w_hh, lengths, hs, gates, B, L, Hd, d_gi, d_gh, d_h0):
    dh = d_h_out.clone()
    g4 = gates.reshape(B, L, 4 * Hd)
    hp = hs.reshape(B, L, Hd)
    dgi, dgh = torch.zeros(B, L, 3 * Hd), torch.zeros(B, L, 3 * Hd)
    for t in range(L - 1, -1, -1):
        r, z, n, ghn = g4[:, t, :Hd], g4[:, t, Hd:2 * Hd], g4[:, t, 2 * Hd:3 * Hd], g4[:, t, 3 * Hd:]
        live = (lengths > t).float().unsqueeze(1)
        dn = dh * (1 - z) * (1 - n * n)
        gz = dh * (hp[:, t] - n) * z * (1 - z)
        gr = dn * ghn * r * (1 - r)
        dgi[:, t] = live * torch.cat([gr, gz, dn], 1)
        dgh[:, t] = live * torch.cat([gr, gz, dn * r], 1)
        dh = live * (dh * z + dgh[:, t] @ w_hh) + (1 - live) * dh
    d_gi.copy_(dgi.reshape(B * L, -1))
    d_gh.copy_(dgh.reshape(B * L, -1))
    if d_h0 is not None:
        d_h0.copy_(dh)


def _xnrs_lengths_from_mask(mask, B, L, lengths):
    lengths.copy_(mask.reshape(B, L).sum(1).round().int())


def _xnrs_score_loss(u, c, targets, weights, kind, B, N, T, gscale, scores, preds, loss, d_u, d_c):
    cc = c.detach().clone().requires_grad_(True)
    uu = None if u is None else u.detach().clone().requires_grad_(True)
    with torch.enable_grad():
        s = cc if u is None else (cc * uu.unsqueeze(1)).sum(-1)
        w = 1.0 if weights is None else weights.reshape(B, N)
        if kind == K.LOSS_MSE_RELU:
            p = torch.relu(s)
            l = (((p - targets.reshape(B, N)) ** 2) * w).mean()
        elif kind == K.LOSS_BCE_LOGITS:
            p = s
            t = targets.reshape(B, N)
            l = ((torch.clamp(s, min=0) - s * t + torch.log1p(torch.exp(-s.abs()))) * w).mean()
        elif kind == K.LOSS_BCE_SIGMOID:
            p = torch.sigmoid(s)
            l = torch.nn.functional.binary_cross_entropy(p, targets.reshape(B, N), weight=None if weights is None else w)
        else:
            p = s
            e = torch.exp(s)
            l = (-torch.log(e[:, 0] / e.sum(1))).mean()
        if d_c is not None:
            gr = torch.autograd.grad(l, (cc,) if u is None else (cc, uu))
            d_c.copy_(gr[0] * gscale)
            if u is not None:
                d_u.copy_(gr[1] * gscale)
    scores.copy_(s.detach())
    if preds is not None:
        preds.copy_(p.detach())
    loss.copy_(l.detach().reshape(1))


def _xnrs_sigmoid(n, x, y):
    y.copy_(torch.sigmoid(x))


def _xnrs_sigmoid_bwd(n, y, dy, dx):
    dx.copy_(dy * y * (1 - y))


def _xnrs_dot_score(u, c, B, N, T, scores):
    scores.copy_((c * u.unsqueeze(1)).sum(-1))


def _xnrs_dot_score_bwd(u, c, d_s, B, N, T, d_u, d_c):
    d_u.copy_((d_s.unsqueeze(-1) * c).sum(1))
    d_c.copy_(d_s.unsqueeze(-1) * u.unsqueeze(1))


def _xnrs_infonce_normalize(emb, Bk, E, ehat, inv_norm):
    inv = 1.0 / emb.norm(dim=1).clamp_min(1e-12)
    ehat.copy_(emb * inv[:, None])
    inv_norm.copy_(inv)


def _xnrs_infonce_rows(sim, labels, Ba, Bk, row0, temperature, stats):
    ex = torch.exp(sim / temperature)
    gi = torch.arange(Ba) + row0
    off = torch.arange(Bk)[None, :] != gi[:, None]
    pos = (labels[gi][:, None] == labels[None, :]) & off
    num, den = (ex * pos).sum(1), (ex * off).sum(1)
    has = pos.any(1)
    stats[0] += (-torch.log(num[has] / (den[has] + 1e-12))).sum()
    stats[1] += has.sum()
    G = ex / temperature * (off / (den[:, None] + 1e-12) - pos / num.clamp_min(1e-38)[:, None])
    sim.copy_(G * has[:, None])


def _xnrs_debug_gemm_trace(buf):
    pass


def _xnrs_infonce_count(labels, Bk, work, count):
    same = labels[:Bk, None] == labels[None, :Bk]
    same.fill_diagonal_(False)
    count.reshape(-1)[0] = float(same.any(1).sum())


def _xnrs_infonce_finalize(stats, loss):
    loss.copy_((stats[0] / (stats[1] + 1e-8)).reshape(1))


def _xnrs_infonce_normalize_bwd(d_ehat, ehat, inv_norm, stats, gscale, Bk, E, d_emb):
    sc = gscale / (stats[1] + 1e-8) if stats is not None else gscale
    d_emb.copy_(sc * inv_norm[:, None] * (d_ehat - ehat * (ehat * d_ehat).sum(1, keepdim=True)))


def _xnrs_adam_step(p, g, m, v, n, lr, b1, b2, eps, step, bc_dev, gscale):
    gg = g * gscale
    m.mul_(b1).add_(gg, alpha=1 - b1)
    v.mul_(b2).addcmul_(gg, gg, value=1 - b2)
    if bc_dev is not None:
        i1, i2 = float(bc_dev[0]), float(bc_dev[1])
    else:
        i1, i2 = 1 / (1 - b1 ** step), 1 / math.sqrt(1 - b2 ** step)
    p.sub_(lr * i1 * m / (v.sqrt() * i2 + eps))


def _xnrs_mark_rows(idx, n, V, skip_row, bitmap, active, count):
    for r in idx.tolist():
        if r < 0 or r >= V or r == skip_row:
            continue
        w, b = r >> 5, 1 << (r & 31)
        word = int(bitmap[w]) & 0xffffffff
        if not word & b:
            word |= b
            bitmap[w] = word - (1 << 32) if word >= (1 << 31) else word
            active[int(count[0])] = r
            count[0] += 1


def _xnrs_adam_rows(p, g, m, v, V, D, active, count, lr, b1, b2, eps, step, bc_dev, gscale):
    rows = active[:int(count[0])].long()
    pr, mr, vr = p.reshape(V, D)[rows], m.reshape(V, D)[rows], v.reshape(V, D)[rows]
    _xnrs_adam_step(pr, g.reshape(V, D)[rows], mr, vr, pr.numel(), lr, b1, b2, eps, step, bc_dev, gscale)
    p.reshape(V, D)[rows], m.reshape(V, D)[rows], v.reshape(V, D)[rows] = pr, mr, vr


def _xnrs_zero_rows(g, V, D, active, count):
    g.reshape(V, D)[active[:int(count[0])].long()] = 0


def _xnrs_adam_tick(step_dev, b1, b2, bc):
    step_dev += 1
    t = int(step_dev)
    bc[0], bc[1] = 1 / (1 - b1 ** t), 1 / math.sqrt(1 - b2 ** t)


def _xnrs_eval_impressions(user, news_vecs, n_news, T, cand_ids, offsets, targets, n_imp, act, scores, metrics):
    from oracle import xnrs_oracle as O
    for i in range(n_imp):
        a, b = int(offsets[i]), int(offsets[i + 1])
        if user is not None:
            cid = cand_ids[a:b].long()
            cid = torch.where((cid >= 0) & (cid < n_news), cid, torch.zeros_like(cid))
            s = news_vecs[cid] @ user[i]
            s = torch.relu(s) if act == 1 else (torch.sigmoid(s) if act == 2 else s)
            scores[a:b] = s
        elif act:
            s = scores[a:b]
            scores[a:b] = torch.relu(s) if act == 1 else torch.sigmoid(s)
        r = O.impression_metrics(targets[a:b].numpy(), scores[a:b].numpy())
        metrics[i] = torch.tensor([r['auc'], r['rr'], r['ndcg@5'], r['ndcg@10'], r['ctr@1'], r['ctr@10']],
                                  dtype=torch.float64)


def _xnrs_metric_sums(metrics, n_imp, sums):
    ok = torch.isfinite(metrics).all(1)
    sums[:6] += metrics[ok].sum(0)
    sums[6] += ok.sum()


def _xnrs_rowdot(x, w, b, n, A, out):
    out.copy_(x.reshape(n, A) @ w + (b[0] if b is not None else 0.0))


def _xnrs_logitpool_fwd(table, V, T, logit, row_mask, ids, R, L, attn, pooled):
    v = ids.reshape(R, L).long()
    m = row_mask[v] if row_mask is not None else torch.ones(R, L)
    e = torch.exp(logit[v]) * m
    a = e / (e.sum(1, keepdim=True) + 1e-8)
    if attn is not None:
        attn.copy_(a.reshape(attn.shape))
    pooled.copy_(torch.einsum('rl,rlt->rt', a, table[v]))


def _xnrs_logitpool_bwd(table, V, T, ids, attn, d_pooled, R, L, d_logit, d_table):
    v = ids.reshape(R, L).long()
    a = attn.reshape(R, L)
    da = torch.einsum('rt,rlt->rl', d_pooled, table[v])
    dl = a * (da - (a * da).sum(1, keepdim=True))
    d_logit.index_add_(0, v.reshape(-1), dl.reshape(-1))
    d_table.index_add_(0, v.reshape(-1), (a[:, :, None] * d_pooled[:, None, :]).reshape(R * L, -1))


def _xnrs_logit_bwd(hid, w2, d_logit, n, A, d_hid, d_w2, d_b2):
    h = hid.reshape(n, A)
    d_hid.copy_((d_logit[:, None] * w2[None, :] * (1 - h * h)).reshape(d_hid.shape))
    d_w2.add_(h.T @ d_logit)
    d_b2.add_(d_logit.sum())


def _xnrs_binary_metrics(scores, targets, offsets, n_imp, out):
    sc = torch.nan_to_num(scores, nan=0.0, posinf=1.0, neginf=0.0)
    for i in range(n_imp):
        a, b = int(offsets[i]), int(offsets[i + 1])
        pred, pos = sc[a:b] > 0.5, targets[a:b] > 0.5
        tp, fp = int((pred & pos).sum()), int((pred & ~pos).sum())
        fn, tn = int((~pred & pos).sum()), int((~pred & ~pos).sum())
        out[i] = torch.tensor([(tp + tn) / max(b - a, 1), tp / (tp + fn) if tp + fn else 0.0, tp / (tp + fp) if tp + fp else 0.0,
                               tn, fp, fn, tp], dtype=torch.float64)
