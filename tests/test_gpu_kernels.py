"""Kernel-level parity on a real B200: every C-ABI entry point against the CPU oracle on the same seeded inputs
(fp32, 1e-4 relative to the tensor scale; index / copy work bit-exact)."""
import math

import numpy as np
import pytest
import torch

from _common import assert_close, load_npz
from oracle import xnrs_oracle as O
from xnrs_b200 import kernels as K
from xnrs_b200.models import components as C

pytestmark = pytest.mark.gpu
DEV = 'cuda'
TOL = 1e-4


def g(seed):
    return torch.Generator().manual_seed(seed)


def cu(t):
    return t.to(DEV).contiguous()


def test_library_reports_sm100_and_counts_launches():
    from xnrs_b200 import _lib
    assert _lib.lib().xnrs_device_is_sm100() == 1
    n0 = K.launch_count()
    K.gemm(cu(torch.randn(8, 8)), cu(torch.randn(8, 8)))
    assert K.launch_count() == n0 + 1


@pytest.mark.parametrize('ta,tb', [(0, 0), (0, 1), (1, 0), (1, 1)])
@pytest.mark.parametrize('M,N,K_', [(77, 130, 45), (256, 256, 768), (1, 5, 3), (300, 64, 1000)])
def test_gemm_layouts(ta, tb, M, N, K_):
    a = torch.randn((K_, M) if ta else (M, K_), generator=g(1))
    b = torch.randn((N, K_) if tb else (K_, N), generator=g(2))
    bias = torch.randn(N, generator=g(3))
    want = (a.T if ta else a) @ (b.T if tb else b) + bias
    got = K.gemm(cu(a), cu(b), trans_a=bool(ta), trans_b=bool(tb), bias=cu(bias))
    assert_close(got, want, 2e-5, 'gemm')


def test_gemm_epilogues_strides_splitk_and_gather():
    M, N, K_ = 200, 96, 333
    a, b = torch.randn(M, K_, generator=g(1)), torch.randn(N, K_, generator=g(2))
    bias, aux = torch.randn(N, generator=g(3)), torch.randn(M, N, generator=g(4))
    base = a @ b.T + bias
    assert_close(K.gemm(cu(a), cu(b), trans_b=True, bias=cu(bias), act=K.ACT_RELU), torch.relu(base), 2e-5, 'relu')
    assert_close(K.gemm(cu(a), cu(b), trans_b=True, bias=cu(bias), act=K.ACT_TANH), torch.tanh(base), 2e-5, 'tanh')
    assert_close(K.gemm(cu(a), cu(b), trans_b=True, bias=cu(bias), act=K.ACT_RELU_MASK, aux=cu(aux)),
                 base * (aux > 0), 2e-5, 'relu mask')
    out = cu(aux.clone())
    K.gemm(cu(a), cu(b), trans_b=True, bias=cu(bias), out=out, accumulate=True)
    assert_close(out, aux + base, 2e-5, 'accumulate')
    # strided output (a column block of a wider buffer) and an unaligned leading dimension on A
    wide = torch.zeros(M, N + 40, device=DEV)
    a_pad = cu(torch.cat([a, torch.zeros(M, 3)], 1))
    K.gemm(a_pad[:, :K_], cu(b), trans_b=True, out=wide[:, 8:8 + N])
    assert_close(wide[:, 8:8 + N], a @ b.T, 2e-5, 'strided')
    assert float(wide[:, :8].abs().max()) == 0 and float(wide[:, 8 + N:].abs().max()) == 0
    # split-K (weight-gradient shape: tiny output, long K) with and without accumulation
    Kl = 20000
    x, dy = torch.randn(Kl, 64, generator=g(5)), torch.randn(Kl, 48, generator=g(6))
    want = dy.T @ x
    assert_close(K.gemm(cu(dy), cu(x), trans_a=True, split_k=16), want, 5e-5, 'split-k')
    assert_close(K.gemm(cu(dy), cu(x), trans_a=True), want, 5e-5, 'auto split-k')
    acc = cu(torch.ones(48, 64))
    K.gemm(cu(dy), cu(x), trans_a=True, out=acc, accumulate=True, split_k=7)
    assert_close(acc, want + 1, 5e-5, 'split-k accumulate')
    # fused row gather on A (forward) and on B (weight gradient)
    table = torch.randn(500, K_, generator=g(7))
    rows = torch.randint(0, 500, (M,), generator=g(8)).int()
    assert_close(K.gemm(cu(table), cu(b), trans_b=True, a_rows=cu(rows)), table[rows.long()] @ b.T, 2e-5, 'a_rows')
    d = torch.randn(M, 40, generator=g(9))
    assert_close(K.gemm(cu(d), cu(table), trans_a=True, b_rows=cu(rows)), d.T @ table[rows.long()], 5e-5, 'b_rows')


def test_gather_is_bit_exact_and_matches_dataset_semantics():
    table = torch.randn(1000, 768, generator=g(0))
    table[0] = 0
    titles = torch.randint(1, 1000, (300, 30), generator=g(1)).int()
    titles[0] = 0
    titles[5, 7:] = 0
    ids = torch.randint(0, 300, (16, 55), generator=g(2)).int()
    ids[3, 10:] = 0
    from xnrs_b200.data import TitleStore
    x, m = TitleStore(cu(table), cu(titles)).dense(cu(ids))
    wx, wm = O.gather_titles(table, titles, ids)
    assert torch.equal(x.cpu(), wx) and torch.equal(m.cpu(), wm)


def _aa_params(F_, A, seed, scale=1.0):
    gg = g(seed)
    return {'p.fc1.weight': torch.randn(A, F_, generator=gg) * scale / math.sqrt(F_),
            'p.fc1.bias': torch.randn(A, generator=gg) * 0.1,
            'p.fc2.weight': torch.randn(1, A, generator=gg) / math.sqrt(A), 'p.fc2.bias': torch.randn(1, generator=gg)}


@pytest.mark.parametrize('R,L,F_,use_mask', [(37, 30, 768, True), (64, 50, 256, True), (50, 4, 256, False)])
def test_additive_pool_forward_backward(R, L, F_, use_mask):
    P = _aa_params(F_, 256, 3)
    x = torch.randn(R, L, F_, generator=g(4))
    ln = torch.randint(0, L + 1, (R,), generator=g(5))
    ln[0], ln[1] = 0, L                                     # fully masked and full rows
    m = (torch.arange(L)[None, :] < ln[:, None]).float().unsqueeze(-1) if use_mask else None
    xo = x.clone().requires_grad_(True)
    Po = {k: v.clone().requires_grad_(True) for k, v in P.items()}
    wo, wa = O.additive_attention(xo, m, Po, 'p', return_weights=True)
    gout = torch.randn(R, 1, F_, generator=g(6))
    (wo * gout).sum().backward()

    mod = C.AdditiveAttention(F_, 256).to(DEV)
    mod.load_state_dict({k[2:]: v for k, v in P.items()})
    xg = cu(x).requires_grad_(True)
    go, ga = mod(xg, None if m is None else cu(m), return_weights=True)
    assert_close(go, wo, TOL, 'pooled')
    assert_close(ga, wa, TOL, 'weights')
    if use_mask:
        assert float(go[0].detach().abs().max()) == 0.0    # fully masked -> exactly zero
    (go * cu(gout)).sum().backward()
    assert_close(xg.grad, xo.grad, TOL, 'dx')
    for k in P:
        want = Po[k].grad
        assert_close(dict(mod.named_parameters())[k[2:]].grad, want, TOL, 'grad ' + k,
                     atol=1e-5 * float(Po['p.fc1.weight'].grad.abs().max()))


def test_personalized_pool_forward_backward():
    R, L, F_, Q, A = 40, 30, 768, 64, 128
    gg = g(11)
    P = {'p.x_fc.weight': torch.randn(A, F_, generator=gg) / math.sqrt(F_), 'p.x_fc.bias': torch.randn(A, generator=gg) * .1,
         'p.q_fc.weight': torch.randn(A, Q, generator=gg) / math.sqrt(Q), 'p.q_fc.bias': torch.randn(A, generator=gg) * .1}
    x, q = torch.randn(R, L, F_, generator=gg), torch.randn(R, 1, Q, generator=gg) * 0.3
    ln = torch.randint(0, L + 1, (R,), generator=gg)
    m = (torch.arange(L)[None, :] < ln[:, None]).float().unsqueeze(-1)
    Po = {k: v.clone().requires_grad_(True) for k, v in P.items()}
    xo, qo = x.clone().requires_grad_(True), q.clone().requires_grad_(True)
    wo = O.personalized_attention(qo, xo, m, Po, 'p')
    gout = torch.randn(R, 1, F_, generator=gg)
    (wo * gout).sum().backward()
    mod = C.PersonalizedAttention(F_, A, Q).to(DEV)
    mod.load_state_dict({k[2:]: v for k, v in P.items()})
    xg, qg = cu(x).requires_grad_(True), cu(q).requires_grad_(True)
    go = mod(qg, xg, cu(m))
    assert_close(go, wo, TOL, 'pooled')
    (go * cu(gout)).sum().backward()
    assert_close(xg.grad, xo.grad, TOL, 'dx')
    assert_close(qg.grad, qo.grad, TOL, 'dq')
    for k in P:
        assert_close(dict(mod.named_parameters())[k[2:]].grad, Po[k].grad, TOL, 'grad ' + k)


@pytest.mark.parametrize('R,L,D,h,p', [(9, 30, 768, 16, 0.0), (7, 50, 768, 16, 0.1), (33, 50, 256, 16, 0.0),
                                        (5, 25, 256, 16, 0.1), (4, 7, 32, 4, 0.0)])
def test_multi_head_attention_forward_backward(R, L, D, h, p):
    gg = g(21)
    P = {}
    for n in ('q_linear', 'k_linear', 'v_linear', 'out'):
        P[f'a.{n}.weight'] = torch.randn(D, D, generator=gg) / math.sqrt(D)
        P[f'a.{n}.bias'] = torch.randn(D, generator=gg) * 0.1
    x = torch.randn(R, L, D, generator=gg)
    ln = torch.randint(1, L + 1, (R,), generator=gg)
    ln[0] = L
    m = (torch.arange(L)[None, :] < ln[:, None]).float().unsqueeze(-1)
    keep = (torch.rand(R, h, L, L, generator=gg) >= p).float() if p > 0 else None
    Po = {k: v.clone().requires_grad_(True) for k, v in P.items()}
    xo = x.clone().requires_grad_(True)
    wo = O.multi_head_attention(xo, m, Po, 'a', h, keep, p)
    gout = torch.randn(R, L, D, generator=gg)
    (wo * gout).sum().backward()
    mod = C.MultiHeadAttention(h, D, dropout=p).to(DEV).eval()
    mod.load_state_dict({k[2:]: v for k, v in P.items()})
    mod.keep_mask = None if keep is None else cu(keep)
    xg = cu(x).requires_grad_(True)
    go = mod(xg, cu(m))
    assert_close(go, wo, TOL, 'mha out')
    (go * cu(gout)).sum().backward()
    assert_close(xg.grad, xo.grad, TOL, 'dx')
    gmax = max(float(v.grad.abs().max()) for v in Po.values())
    for k in P:     # d/d k_linear.bias is analytically 0 (a key bias shifts every score of a row equally): noise only
        assert_close(dict(mod.named_parameters())[k[2:]].grad, Po[k].grad, TOL, 'grad ' + k, atol=1e-5 * gmax)


def test_mha_seeded_dropout_is_reproducible_and_unbiased():
    R, L, D, h = 64, 30, 256, 16
    mod = C.MultiHeadAttention(h, D, dropout=0.1).to(DEV).train()
    x, m = cu(torch.randn(R, L, D, generator=g(1))), torch.ones(R, L, 1, device=DEV)
    torch.manual_seed(5)
    a = mod(x, m)
    torch.manual_seed(5)
    b = mod(x, m)
    assert torch.equal(a, b)
    mod.eval()
    c = mod(x, m)
    assert 1e-3 < float((a - c).abs().mean()) < 1.0         # dropout changes the output, mean stays close
    assert abs(float(a.mean() - c.mean())) < 0.05 * float(c.abs().mean()) + 1e-3


@pytest.mark.parametrize('B,L,I,Hd,with_h0', [(9, 25, 272, 136, False), (6, 7, 24, 10, True), (130, 50, 272, 272, True)])
def test_gru_forward_backward(B, L, I, Hd, with_h0):
    gg = g(31)
    P = {'g.weight_ih_l0': torch.randn(3 * Hd, I, generator=gg) / math.sqrt(I),
         'g.weight_hh_l0': torch.randn(3 * Hd, Hd, generator=gg) / math.sqrt(Hd),
         'g.bias_ih_l0': torch.randn(3 * Hd, generator=gg) * .1, 'g.bias_hh_l0': torch.randn(3 * Hd, generator=gg) * .1}
    x = torch.randn(B, L, I, generator=gg)
    lens = torch.randint(1, L + 1, (B,), generator=gg)
    lens[0] = L
    h0 = torch.randn(B, Hd, generator=gg) if with_h0 else None
    Po = {k: v.clone().requires_grad_(True) for k, v in P.items()}
    xo = x.clone().requires_grad_(True)
    h0o = None if h0 is None else h0.clone().requires_grad_(True)
    wo = O.gru_last_hidden(xo, lens, Po, 'g', h0o)
    gout = torch.randn(B, Hd, generator=gg)
    (wo * gout).sum().backward()
    Pg = {k: cu(v).requires_grad_(True) for k, v in P.items()}
    xg = cu(x).requires_grad_(True)
    h0g = None if h0 is None else cu(h0).requires_grad_(True)
    go = K.GruLastFn.apply(xg, cu(lens.int()), Pg['g.weight_ih_l0'], Pg['g.weight_hh_l0'], Pg['g.bias_ih_l0'],
                           Pg['g.bias_hh_l0'], h0g)
    assert_close(go, wo, TOL, 'gru h')
    (go * cu(gout)).sum().backward()
    assert_close(xg.grad, xo.grad, TOL, 'dx')
    if h0 is not None:
        assert_close(h0g.grad, h0o.grad, TOL, 'dh0')
    for k in P:
        assert_close(Pg[k].grad, Po[k].grad, TOL, 'grad ' + k)


@pytest.mark.parametrize('kind', [K.LOSS_MSE_RELU, K.LOSS_BCE_LOGITS, K.LOSS_NLL, K.LOSS_BCE_SIGMOID])
@pytest.mark.parametrize('B,N,T', [(64, 5, 256), (3, 300, 272)])
def test_score_loss_fused(kind, B, N, T):
    gg = g(41)
    u, c = torch.randn(B, T, generator=gg) * 0.2, torch.randn(B, N, T, generator=gg) * 0.2
    t = torch.zeros(B, N)
    t[:, 0] = 1
    w = torch.rand(B, N, generator=gg) + 0.5 if kind != K.LOSS_NLL else None
    uo, co = u.clone().requires_grad_(True), c.clone().requires_grad_(True)
    s = O.dot_scoring(uo.unsqueeze(1), co)
    if kind == K.LOSS_MSE_RELU:
        want, _ = O.mse_relu_loss(s, t.unsqueeze(-1), w.unsqueeze(-1))
    elif kind == K.LOSS_BCE_LOGITS:
        want = O.bce_logits_loss(s, t.unsqueeze(-1), w.unsqueeze(-1))
    elif kind == K.LOSS_BCE_SIGMOID:
        want, _ = O.bce_sigmoid_loss(s, t.unsqueeze(-1), w.unsqueeze(-1))
    else:
        want = O.ranking_nll(s[:, :1, 0], s[:, 1:, 0])
    want.backward()
    ug, cg = cu(u).requires_grad_(True), cu(c).requires_grad_(True)
    loss, preds, scores = K.ScoreLossFn.apply(ug, cg, cu(t).reshape(-1), None if w is None else cu(w).reshape(-1), kind)
    assert_close(loss, want, TOL, 'loss')
    assert_close(scores, s.squeeze(-1), TOL, 'scores')
    (loss * 3.0).backward()
    assert_close(ug.grad, 3 * uo.grad, TOL, 'du')
    assert_close(cg.grad, 3 * co.grad, TOL, 'dc')
    # standalone scorer
    ug2, cg2 = cu(u).requires_grad_(True), cu(c).requires_grad_(True)
    s2 = C.DotScoring()(ug2.unsqueeze(1), cg2)
    assert_close(s2, s, TOL, 'dot scoring')
    s2.sum().backward()
    assert_close(cg2.grad, u.unsqueeze(1).expand_as(c), TOL, 'dot dc')


@pytest.mark.parametrize('B,E', [(6, 8), (40, 16), (512, 256), (1024, 256)])
def test_infonce(B, E):
    gg = g(51)
    e = torch.randn(B, E, generator=gg)
    lab = torch.randint(0, 6, (B,), generator=gg)
    if B > 8:
        lab[7] = 99                                         # an anchor without positives is skipped
    eo = e.clone().requires_grad_(True)
    want = O.contrastive_loss(eo, lab, 0.08)
    want.backward()
    eg = cu(e).requires_grad_(True)
    got = K.InfoNCEFn.apply(eg, cu(lab.int()), 0.08)
    assert_close(got, want, TOL, 'infonce')
    got.backward()
    assert_close(eg.grad, eo.grad, 2e-4, 'd emb')


def test_infonce_known_answers_from_reference():
    fx = load_npz('loss_metrics')
    for tag in ('cl', 'cl2'):
        got = K.InfoNCEFn.apply(cu(torch.tensor(fx[tag + '/e'])), cu(torch.tensor(fx[tag + '/labels']).int()), 0.08)
        assert_close(got, fx[tag + '/loss'], TOL, tag)
    all_diff = K.InfoNCEFn.apply(cu(torch.randn(5, 8)), cu(torch.arange(5).int()), 0.08)
    assert float(all_diff) == 0.0                           # no anchor has a positive -> 0 / (0 + 1e-8)


@pytest.mark.parametrize('Bk,n_lab', [(1, 3), (7, 3), (256, 6), (1025, 400), (8192, 6), (8192, 100000)])
def test_infonce_global_count_from_labels(Bk, n_lab):
    """data parallel: the global normaliser (rows with a same-label partner) from the gathered labels alone, exact"""
    lab = torch.randint(0, n_lab, (Bk,), generator=g(52)).int()
    same = lab[:, None] == lab[None, :]
    same.fill_diagonal_(False)
    want = float(same.any(1).sum())
    work = cu(torch.zeros(4))
    work[1] = 123.0                                         # overwritten, not accumulated
    K.call('xnrs_infonce_count', cu(lab), Bk, work[2:], work[1:2])
    assert float(work[1]) == want


def test_peer_exchange_kernels_on_one_rank():
    """the two peer-memory exchange kernels of the data-parallel InfoNCE (csrc/peer.cu) with world = 1 — one buffer, its own
    address as the only peer: normalise + all-gather must equal xnrs_infonce_normalize, reduce-scatter + normalisation backward
    must equal xnrs_infonce_normalize_bwd times the loss scale; the epochs advance so a second launch works unchanged
    (tools/check_dp_gpu.py runs them across real GPUs under torchrun)"""
    Ba, E = 300, 256
    gg = g(77)
    emb, labels = torch.randn(Ba, E, generator=gg), torch.randint(0, 6, (Ba,), generator=gg).int()
    off_flags_ag, off_flags_rs, off_ehat = 0, 256, 512
    off_inv = off_ehat + Ba * E * 4
    off_lab = off_inv + (Ba * 4 + 255) // 256 * 256
    off_dehat = off_lab + (Ba * 4 + 255) // 256 * 256
    buf = cu(torch.zeros((off_dehat + Ba * E * 4) // 4))
    ptrs = torch.tensor([buf.data_ptr()], dtype=torch.int64, device=DEV)
    ctl = cu(torch.zeros(8, dtype=torch.int32))
    want_hat, want_inv = cu(torch.empty(Ba, E)), cu(torch.empty(Ba))
    K.call('xnrs_infonce_normalize', cu(emb), Ba, E, want_hat, want_inv)
    d_ehat = torch.randn(Ba, E, generator=gg)
    stats = cu(torch.tensor([3.0, 17.0]))
    gdev = cu(torch.tensor([0.25]))
    want_d = cu(torch.empty(Ba, E))
    K.call('xnrs_infonce_normalize_bwd', cu(d_ehat), want_hat, want_inv, stats, 2.0 * 0.25, Ba, E, want_d)
    for launch in (1, 2):
        buf[off_ehat // 4:].zero_()
        K.call('xnrs_peer_normalize_allgather', cu(emb), cu(labels), Ba, E, 0, 1, ptrs, off_flags_ag, off_ehat, off_inv, off_lab, ctl)
        got_hat = buf[off_ehat // 4: off_ehat // 4 + Ba * E].view(Ba, E)
        assert_close(got_hat, want_hat, 1e-6, 'normalised rows')        # (the two kernels sum the squares in different orders)
        assert_close(buf[off_inv // 4: off_inv // 4 + Ba], want_inv, 1e-6, 'inverse norms')
        assert torch.equal(buf[off_lab // 4: off_lab // 4 + Ba].view(torch.int32).cpu(), labels)
        buf[off_dehat // 4: off_dehat // 4 + Ba * E].copy_(cu(d_ehat).view(-1))
        got_d = cu(torch.empty(Ba, E))
        K.call('xnrs_peer_reduce_scatter_normalize_bwd', ptrs, off_flags_rs, off_dehat, Ba, E, 0, 1, want_hat, want_inv, stats, 2.0, gdev,
               got_d, ctl)
        assert_close(got_d, want_d, 1e-6, 'peer reduce-scatter + normalisation backward')
        c = ctl.cpu().tolist()
        assert c[0] == launch and c[2] == launch and c[1] == 0 and c[3] == 0 and c[4] == 0      # epochs advanced, tickets reset, no error


@pytest.mark.parametrize('tag', ['met', 'mett', 'kat'])
def test_ranking_metrics_against_reference_values(tag):
    fx = load_npz('loss_metrics')
    if tag == 'kat':
        s, t, off, want = fx['kat/s'], fx['kat/t'], np.array([0, len(fx['kat/s'])]), fx['kat/values'][None]
    else:
        s, t, off, want = fx[tag + '/scores'], fx[tag + '/targets'], fx[tag + '/offsets'], fx[tag + '/values']
    sc = cu(torch.tensor(s))
    _, m = K.eval_impressions(None, None, None, cu(torch.tensor(off)), cu(torch.tensor(t)), act=0, scores=sc)
    np.testing.assert_allclose(m.cpu().numpy(), want, rtol=0, atol=1e-6)


def test_eval_impressions_scoring_ties_nan_and_long_segments():
    gg = g(61)
    T, n_news = 256, 5000
    vecs, users = torch.randn(n_news, T, generator=gg) * 0.1, torch.randn(40, T, generator=gg) * 0.1
    sizes = torch.randint(5, 74, (40,), generator=gg)
    sizes[3], sizes[4] = 3000, 2                             # longer than the shared-memory staging; tiny
    off = torch.cat([torch.zeros(1, dtype=torch.long), sizes.cumsum(0)])
    cand = torch.randint(0, n_news, (int(off[-1]),), generator=gg).int()
    tg = torch.zeros(int(off[-1]))
    for i in range(40):
        tg[off[i]:off[i] + 1 + i % 3] = 1
    tg[off[4]:off[5]] = torch.tensor([1., 0.])
    scores, m = K.eval_impressions(cu(users), cu(vecs), cu(cand), cu(off), cu(tg), act=1)
    want_s = torch.cat([torch.relu(vecs[cand[off[i]:off[i + 1]].long()] @ users[i]) for i in range(40)])
    assert_close(scores, want_s, TOL, 'scores')
    sc = scores.cpu().numpy()                                # rank the GPU's own scores: ties (relu zeros) are exact
    for i in range(40):
        r = O.impression_metrics(tg[off[i]:off[i + 1]].numpy(), sc[off[i]:off[i + 1]])
        want = [r['auc'], r['rr'], r['ndcg@5'], r['ndcg@10'], r['ctr@1'], r['ctr@10']]
        np.testing.assert_allclose(m[i].cpu().numpy(), want, rtol=0, atol=1e-9, err_msg=f'impression {i}')
    # nan_to_num(nan 0, +inf 1, -inf 0) (training.py:211) and single-class impressions
    s = cu(torch.tensor([float('nan'), float('inf'), -float('inf'), 0.5, 0.2, 0.7]))
    t = cu(torch.tensor([1., 0, 0, 1, 1, 1]))
    _, m2 = K.eval_impressions(None, None, None, cu(torch.tensor([0, 4, 6])), t, act=0, scores=s.clone())
    r = O.impression_metrics([1, 0, 0, 1], np.array([np.nan, np.inf, -np.inf, 0.5], dtype=np.float32))
    np.testing.assert_allclose(m2[0].cpu().numpy(), [r['auc'], r['rr'], r['ndcg@5'], r['ndcg@10'], r['ctr@1'], r['ctr@10']],
                               atol=1e-9)
    assert math.isnan(float(m2[1, 0]))                       # only positives: AUC undefined
    sums = K.metric_sums(m2).cpu().numpy()
    assert sums[6] == 1 and abs(sums[1] - float(m2[0, 1])) < 1e-12


def test_adam_matches_oracle_and_device_counter():
    gg = g(71)
    p, m, v = torch.randn(1000, generator=gg), torch.zeros(1000), torch.zeros(1000)
    pg, mg, vg = cu(p), cu(m), cu(v)
    p2, m2, v2 = cu(p), cu(m), cu(v)
    step_dev, bc = torch.zeros(1, dtype=torch.int32, device=DEV), torch.zeros(2, device=DEV)
    for step in range(1, 5):
        gr = torch.randn(1000, generator=gg)
        O.adam_step(p, gr, m, v, step, 1e-3)
        K.adam_step(pg, cu(gr), mg, vg, 1e-3, step=step)
        K.call('xnrs_adam_tick', step_dev, 0.9, 0.999, bc)
        K.adam_step(p2, cu(gr), m2, v2, 1e-3, step=0, bc_dev=bc)
    assert_close(pg, p, 1e-6, 'adam')
    assert_close(p2, p, 1e-6, 'adam (device step counter)')


def test_embedding_dropout_lengths_and_collapse():
    gg = g(81)
    w = torch.randn(50, 16, generator=gg)
    idx = torch.randint(0, 50, (7, 3), generator=gg)
    idx[0, 0] = 0
    wg = cu(w).requires_grad_(True)
    out = K.EmbeddingFn.apply(wg, cu(idx), 0)
    assert torch.equal(out.cpu(), w[idx.reshape(-1)])
    out.sum().backward()
    want = torch.zeros(50, 16).index_add_(0, idx.reshape(-1), torch.ones(21, 16))
    want[0] = 0                                              # padding_idx receives no gradient
    assert_close(wg.grad, want, 1e-6, 'embedding grad')
    x = cu(torch.randn(1000, generator=gg))
    keep = cu((torch.rand(1000, generator=gg) > 0.07).float())
    assert_close(K.DropoutFn.apply(x, keep, 0.07, 0), x * keep / 0.93, 1e-6, 'dropout')
    y = K.DropoutFn.apply(x, None, 0.5, 123)
    frac = float((y == 0).float().mean())
    assert 0.4 < frac < 0.6 and torch.equal(y, K.DropoutFn.apply(x, None, 0.5, 123))
    m = cu((torch.rand(9, 25, generator=gg) > 0.5).float())
    lengths = torch.empty(9, dtype=torch.int32, device=DEV)
    K.call('xnrs_lengths_from_mask', m.reshape(-1), 9, 25, lengths)
    assert torch.equal(lengths.cpu(), m.cpu().sum(1).int())
    assert torch.equal(K.collapse_mask(m.reshape(-1), 9, 25).cpu(), m.cpu().sum(1).clamp(0, 1))


def test_layer_goldens_from_reference():
    fx = load_npz('layers')
    x, m = cu(torch.tensor(fx['x'])), cu(torch.tensor(fx['m']))

    def load(mod, tag):
        mod.load_state_dict({k[len(tag) + 1:]: torch.tensor(v) for k, v in fx.items()
                             if k.startswith(tag + '/') and (k.endswith('weight') or k.endswith('bias'))})
        return mod.to(DEV).eval()

    aa = load(C.AdditiveAttention(32, 256), 'aa')
    o, a = aa(x, m, return_weights=True)
    assert_close(o, fx['aa/out'], TOL, 'aa')
    assert_close(a, fx['aa/a'], TOL, 'aa weights')
    assert_close(aa(x), fx['aa/out_nomask'], TOL, 'aa nomask')
    mha = load(C.MultiHeadAttention(4, 32), 'mha')
    assert_close(mha(x, m), fx['mha/out'], TOL, 'mha')
    xq = x.clone()
    xq[1, 5] += 1.0
    assert_close(mha(xq, m), fx['mha/out_perturb_padkey'], TOL, 'mha padded key')
    pa = load(C.PersonalizedAttention(32, 128, 8), 'pa')
    assert_close(pa(cu(torch.tensor(fx['pa/q'])), x, m), fx['pa/out'], TOL, 'pa')
    assert_close(C.MaskedMean()(x, m), fx['mm/out'], TOL, 'masked mean')
    gw = {k[4:]: cu(torch.tensor(v)) for k, v in fx.items() if k.startswith('gru/') and '_l0' in k}
    lens = cu(torch.tensor(fx['gru/lens']).int())
    args = (gw['weight_ih_l0'], gw['weight_hh_l0'], gw['bias_ih_l0'], gw['bias_hh_l0'])
    assert_close(K.GruLastFn.apply(x, lens, *args, None), fx['gru/h_zero'], TOL, 'gru')
    assert_close(K.GruLastFn.apply(x, lens, *args, cu(torch.tensor(fx['gru/h0']))), fx['gru/h_init'], TOL, 'gru h0')


def test_cpu_tensors_are_rejected():
    with pytest.raises(RuntimeError, match='CUDA'):
        K.gemm(torch.randn(4, 4), torch.randn(4, 4))


@pytest.mark.parametrize('prec,tol', [('tf32x3', 5e-5), ('tf32', 3e-3)])
@pytest.mark.parametrize('ta,tb', [(0, 1), (1, 0), (0, 0), (1, 1)])
@pytest.mark.parametrize('M,N,K_', [(256, 256, 768), (1000, 200, 333 * 4), (128, 64, 32), (4097, 768, 768)])
def test_tensor_core_gemm(prec, tol, ta, tb, M, N, K_):
    """the tcgen05/TMA path of xnrs_gemm (all four operand-major combinations, ragged tiles) against fp32 matmul"""
    a = torch.randn((K_, M) if ta else (M, K_), generator=g(1)) / math.sqrt(K_)      # O(1) outputs
    b = torch.randn((N, K_) if tb else (K_, N), generator=g(2))
    bias = torch.randn(N, generator=g(3))
    want = (a.double().T if ta else a.double()) @ (b.double().T if tb else b.double()) + bias.double()
    with K.precision(prec):
        got = K.gemm(cu(a), cu(b), trans_a=bool(ta), trans_b=bool(tb), bias=cu(bias))
        assert_close(got, want.float(), tol, f'tc gemm {prec}')
        assert_close(K.gemm(cu(a), cu(b), trans_a=bool(ta), trans_b=bool(tb), bias=cu(bias), act=K.ACT_TANH),
                     torch.tanh(want).float(), tol, 'tanh epilogue')
        acc = cu(torch.ones(M, N))
        K.gemm(cu(a), cu(b), trans_a=bool(ta), trans_b=bool(tb), out=acc, accumulate=True)
        assert_close(acc, (want - bias.double() + 1).float(), tol, 'accumulate')


@pytest.mark.parametrize('prec,tol', [('tf32x3', 3e-5), ('tf32', 3e-3)])
def test_tensor_core_gemm_split_k_weight_gradient_shape(prec, tol):
    Kl = 40000
    x, dy = torch.randn(Kl, 768, generator=g(5)), torch.randn(Kl, 256, generator=g(6))
    want = (dy.double().T @ x.double()).float()
    with K.precision(prec):
        assert_close(K.gemm(cu(dy), cu(x), trans_a=True), want, tol, 'auto split-k')
        acc = cu(torch.ones(256, 768))
        K.gemm(cu(dy), cu(x), trans_a=True, out=acc, accumulate=True, split_k=5)
        assert_close(acc, want + 1, tol, 'split-k accumulate')
        # more splits than k-blocks to share out (K = 100 -> four 32-wide blocks for 7 splits) and the K range not a multiple
        # of the rounded split length (K = 1650): the trailing splits must not contribute
        for Ks, sk in ((100, 7), (1650, 12), (1650, 0)):
            xs, dys = x[:Ks].contiguous(), dy[:Ks].contiguous()
            want_s = (dys.double().T @ xs.double()).float()
            acc = cu(torch.ones(256, 768))
            K.gemm(cu(dys), cu(xs), trans_a=True, out=acc, accumulate=True, split_k=sk)
            assert_close(acc, want_s + 1, tol, f'split-k {sk} of K={Ks}')
            if prec == 'tf32':
                got16 = K.gemm_bf16(K.cast_bf16(cu(dys)), K.cast_bf16(cu(xs)), trans_a=True, split_k=sk)
                assert_close(got16, want_s, 2e-2, f'bf16 split-k {sk} of K={Ks}')


@pytest.mark.parametrize('prec,tol', [('tf32x3', 5e-5), ('tf32', 3e-3)])
def test_tensor_core_gemm_tma_gather(prec, tol):
    """table rows gathered by TMA gather4 straight into the swizzled MMA tiles: forward (rows of A) and
    weight-gradient (rows of B along K) forms, ragged sizes"""
    V, D, A_ = 5000, 768, 256
    table = torch.randn(V, D, generator=g(1)) / math.sqrt(D)
    for R in (1650, 4097, 41000):
        rows = torch.randint(0, V, (R,), generator=g(2)).int()
        w = torch.randn(A_, D, generator=g(3))
        bias = torch.randn(A_, generator=g(4))
        want = torch.tanh(table[rows.long()].double() @ w.double().T + bias.double()).float()
        d = torch.randn(R, A_, generator=g(5))
        want_dw = (d.double().T @ table[rows.long()].double()).float()
        from xnrs_b200 import _lib
        lib = _lib.lib()
        for two in (-1, 1):         # 1: force the CTA-pair kernel, whose gather is a cp.async producer warp (LSU), not TMA
            lib.xnrs_set_option(b'gemm_2cta', two)
            try:
                with K.precision(prec):
                    got = K.gemm(cu(table), cu(w), trans_b=True, bias=cu(bias), act=K.ACT_TANH, a_rows=cu(rows))
                    assert_close(got, want, tol, f'gathered forward (gemm_2cta={two})')
                    if two == 1:
                        assert lib.xnrs_last_gemm_kernel().decode().startswith('gemm_tc2_kernel')
                    got_dw = K.gemm(cu(d), cu(table), trans_a=True, b_rows=cu(rows))
                    assert_close(got_dw, want_dw, tol, f'gathered weight gradient (gemm_2cta={two})')
            finally:
                lib.xnrs_set_option(b'gemm_2cta', -1)


@pytest.mark.parametrize('prec,tol', [('tf32x3', 5e-5), ('tf32', 3e-3)])
@pytest.mark.parametrize('ta,tb', [(0, 1), (1, 0), (0, 0), (1, 1)])
def test_tensor_core_gemm_cta_pair_kernel(prec, tol, ta, tb):
    """the cta_group::2 (CTA-pair, 256x256 tile) variant of the tcgen05 GEMM — default for 3xTF32, forced here for both precisions"""
    from xnrs_b200 import _lib
    lib = _lib.lib()
    assert lib.xnrs_set_option(b'gemm_2cta', 1) == 0
    try:
        for M, N, K_ in [(1000, 200, 1332), (4097, 768, 768)]:
            a = torch.randn((K_, M) if ta else (M, K_), generator=g(1)) / math.sqrt(K_)
            b = torch.randn((N, K_) if tb else (K_, N), generator=g(2))
            bias = torch.randn(N, generator=g(3))
            want = (a.double().T if ta else a.double()) @ (b.double().T if tb else b.double()) + bias.double()
            with K.precision(prec):
                got = K.gemm(cu(a), cu(b), trans_a=bool(ta), trans_b=bool(tb), bias=cu(bias), act=K.ACT_RELU)
            assert_close(got, torch.relu(want).float(), tol, f'2-CTA gemm {prec}')
    finally:
        lib.xnrs_set_option(b'gemm_2cta', -1)


@pytest.mark.parametrize('n_news,S,n', [(60, 12, 40), (65238, 30, 56320), (160000, 50, 8192), (33, 7, 500)])
def test_plan_kernels_match_torch_index_ops(n_news, S, n):
    """device-side id plumbing (xnrs_plan_dedup / xnrs_plan_ragged) == torch.unique + boolean compaction, bit exact, and
    the padding past the counts is harmless (article 0 / token 0 / empty groups)"""
    from xnrs_b200 import synthetic as syn
    from xnrs_b200.data import T_GRANULE, TitleStore, plan_titles
    cat = syn.make_catalogue(n_news, S, 500, 8, seed=n)
    tt = cat.title_tokens.clone()
    tt[3, 1] = 0                                           # a pad token in the MIDDLE of a title: compaction, not truncation
    ids = torch.from_numpy(syn.zipf_news(np.random.default_rng(n), n_news, n)).int()
    ids[::7] = 0                                           # pad slots
    store = TitleStore(cu(cat.token_table), cu(tt))
    for dedup in (True, False):
        plan = plan_titles(store, cu(ids), dedup, True).acquire()
        uniq, inv = torch.unique(ids, return_inverse=True) if dedup else (ids, None)
        U = uniq.numel()
        assert plan.n_titles == U and plan.uniq.numel() >= U
        assert torch.equal(plan.uniq[:U].cpu(), uniq.int()) and int(plan.uniq[U:].abs().sum()) == 0
        if dedup:
            assert torch.equal(plan.inv.cpu(), inv.int())
        tok = tt[uniq.long()]
        valid = tok != 0
        lens = valid.sum(1)
        T = int(lens.sum())
        assert plan.n_rows == T and plan.rows.numel() % T_GRANULE == 0 and plan.rows.numel() >= T
        assert torch.equal(plan.rows[:T].cpu(), tok[valid]) and int(plan.rows[T:].abs().sum()) == 0
        seg = torch.zeros(U + 1, dtype=torch.int64)
        seg[1:] = torch.cumsum(lens, 0)
        assert torch.equal(plan.seg[:U + 1].cpu().long(), seg) and bool((plan.seg[U:] == T).all())
        assert torch.equal(plan.cm[:U].cpu(), (lens > 0).float()) and float(plan.cm[U:].sum()) == 0
    # fixed-length layout (self-attention keeps pad tokens): rows / mask of the distinct articles, padded titles = article 0
    plan = plan_titles(store, cu(ids), True, False).acquire()
    uniq = torch.unique(ids)
    U = uniq.numel()
    tok = tt[uniq.long()].reshape(-1)
    assert torch.equal(plan.rows[:U * S].cpu(), tok) and int(plan.rows[U * S:].abs().sum()) == 0
    assert torch.equal(plan.mask[:U * S].cpu(), (tok != 0).float()) and float(plan.mask[U * S:].sum()) == 0


@pytest.mark.parametrize('ta,tb', [(0, 1), (1, 0), (0, 0), (1, 1)])
def test_bf16_tensor_core_gemm(ta, tb):
    """xnrs_gemm_bf16 (tcgen05 kind::f16 on the CTA-pair kernel): all four operand layouts, ragged sizes, bias / tanh, bf16 and
    fp32 outputs, split-K accumulation — against float64 on the bf16-rounded operands"""
    for M, N, K_ in [(1000, 200, 1344), (4104, 768, 768), (256, 768, 9000)]:          # row strides multiples of 8 (16 bytes)
        a = (torch.randn((K_, M) if ta else (M, K_), generator=g(1)) / math.sqrt(K_)).bfloat16()
        b = torch.randn((N, K_) if tb else (K_, N), generator=g(2)).bfloat16()
        bias = torch.randn(N, generator=g(3))
        want = (a.double().T if ta else a.double()) @ (b.double().T if tb else b.double())
        got = K.gemm_bf16(cu(a), cu(b), trans_a=bool(ta), trans_b=bool(tb))
        assert_close(got, want.float(), 2e-5, f'bf16 gemm fp32 out {M}x{N}x{K_}')
        got = K.gemm_bf16(cu(a), cu(b), trans_a=bool(ta), trans_b=bool(tb), bias=cu(bias), act=K.ACT_TANH, out_bf16=True)
        assert got.dtype == torch.bfloat16
        assert_close(got.float(), torch.tanh(want + bias.double()).float(), 8e-3, 'bf16 gemm bf16 out + tanh')
        acc = torch.ones(M, N)
        out = cu(acc.clone())
        K.gemm_bf16(cu(a), cu(b), trans_a=bool(ta), trans_b=bool(tb), out=out, accumulate=True, split_k=3)
        assert_close(out, (want + 1).float(), 2e-5, 'bf16 gemm split-K accumulate')


def test_bf16_tensor_core_gemm_fused_gather():
    """bf16 table rows gathered by the cp.async warp: forward (rows of A) and weight-gradient (rows of B along K) forms"""
    V, D, A_ = 5000, 768, 256
    table = (torch.randn(V, D, generator=g(1)) / math.sqrt(D)).bfloat16()
    for R in (1650, 41000):
        rows = torch.randint(0, V, (R,), generator=g(2)).int()
        w = torch.randn(A_, D, generator=g(3)).bfloat16()
        want = table[rows.long()].double() @ w.double().T
        got = K.gemm_bf16(cu(table), cu(w), trans_b=True, a_rows=cu(rows))
        assert_close(got, want.float(), 2e-5, 'bf16 gathered forward')
        d = torch.randn(R, A_, generator=g(5)).bfloat16()
        want_dw = d.double().T @ table[rows.long()].double()
        got_dw = K.gemm_bf16(cu(d), cu(table), trans_a=True, b_rows=cu(rows))
        assert_close(got_dw, want_dw.float(), 5e-5, 'bf16 gathered weight gradient')


@pytest.mark.parametrize('ta,tb', [(0, 1), (1, 0), (0, 0), (1, 1)])
def test_bf16x3_tensor_core_gemm(ta, tb):
    """xnrs_gemm_bf16x3: fp32 operands pre-split into two bf16 planes (x ~ hi + lo), three kind::f16 MMAs per k-step — the result
    must be fp32-accurate (1e-4 class; measured ~1e-6) against float64 on the ORIGINAL fp32 operands"""
    for M, N, K_ in [(1000, 200, 1344), (4104, 768, 768), (256, 768, 9000)]:
        a = torch.randn((K_, M) if ta else (M, K_), generator=g(1)) / math.sqrt(K_)
        b = torch.randn((N, K_) if tb else (K_, N), generator=g(2))
        bias = torch.randn(N, generator=g(3))
        want = (a.double().T if ta else a.double()) @ (b.double().T if tb else b.double())
        ah, al = K.split_bf16(cu(a))
        bh, bl = K.split_bf16(cu(b))
        assert float((ah.float() + al.float() - cu(a)).abs().max()) <= 2 ** -16 * float(a.abs().max())
        got = K.gemm_bf16x3(ah, al, bh, bl, trans_a=bool(ta), trans_b=bool(tb))
        assert_close(got, want.float(), 2e-5, f'bf16x3 gemm {M}x{N}x{K_}')
        # fp16 planes (22 mantissa bits) on both operands (the MMA rejects a mixed bf16 / fp16 pair: refused up front)
        fh, fl = K.split_bf16(cu(b), fp16=True)
        assert float((fh.float() + fl.float() - cu(b)).abs().max()) <= 2 ** -22 * float(b.abs().max())
        gh, gl = K.split_bf16(cu(a), fp16=True)
        assert_close(K.gemm_bf16x3(gh, gl, fh, fl, trans_a=bool(ta), trans_b=bool(tb)), want.float(), 5e-6, 'fp16 x fp16 planes')
        with pytest.raises(RuntimeError):
            K.gemm_bf16x3(ah, al, fh, fl, trans_a=bool(ta), trans_b=bool(tb))
        got = K.gemm_bf16x3(ah, al, bh, bl, trans_a=bool(ta), trans_b=bool(tb), bias=cu(bias), act=K.ACT_TANH)
        assert_close(got, torch.tanh(want + bias.double()).float(), 1e-4, 'bf16x3 gemm bias + tanh')
        out = cu(torch.ones(M, N))
        K.gemm_bf16x3(ah, al, bh, bl, trans_a=bool(ta), trans_b=bool(tb), out=out, accumulate=True, split_k=3)
        assert_close(out, (want + 1).float(), 2e-5, 'bf16x3 gemm split-K accumulate')


def test_bf16x3_tensor_core_gemm_fused_gather():
    """the two planes of a table gathered by the cp.async warps: forward (rows of A) and weight-gradient (rows of B along K)"""
    V, D, A_ = 5000, 768, 256
    table = torch.randn(V, D, generator=g(1)) / math.sqrt(D)
    th, tl = K.split_bf16(cu(table), fp16=True)             # the frozen table and the weight: fp16 planes; the gradient: bf16 planes
    for R in (1650, 41000):
        rows = torch.randint(0, V, (R,), generator=g(2)).int()
        w = torch.randn(A_, D, generator=g(3))
        wh, wl = K.split_bf16(cu(w), fp16=True)
        want = table[rows.long()].double() @ w.double().T
        got = K.gemm_bf16x3(th, tl, wh, wl, trans_b=True, a_rows=cu(rows))
        assert_close(got, want.float(), 5e-6, 'fp16-plane gathered forward')
        d = torch.randn(R, A_, generator=g(5))
        dh, dl = K.split_bf16(cu(d))
        tbh, tbl = K.split_bf16(cu(table))
        want_dw = d.double().T @ table[rows.long()].double()
        got_dw = K.gemm_bf16x3(dh, dl, tbh, tbl, trans_a=True, b_rows=cu(rows))
        assert_close(got_dw, want_dw.float(), 3e-5, 'bf16x3 gathered weight gradient')


def test_gemm_rejects_the_bf16_precision_on_fp32_operands():
    with K.precision('bf16'):
        assert K._gemm_precision() == K.PRECISIONS['tf32']          # python maps fp32-stored GEMMs of the bf16 mode to TF32
    with pytest.raises(RuntimeError, match='xnrs_gemm_bf16'):
        a = cu(torch.randn(256, 64))
        K.call('xnrs_gemm', 0, 1, 256, 256, 64, K._mat(a), 64, None, K._mat(a), 64, None, K._mat(cu(torch.empty(256, 256))), 256, None, 0,
               None, 0, 0, K.PRECISIONS['bf16'])


def test_active_row_adam_kernel_is_bit_identical_to_the_dense_pass():
    """xnrs_mark_rows + xnrs_adam_rows over the rows touched so far == xnrs_adam_step over the whole table, bit for bit, for 30
    steps in which new rows keep appearing (never-touched rows have g = m = v = 0: the dense update leaves them unchanged)"""
    V, D = 5000, 136
    gen = g(9)
    p0 = torch.randn(V, D, generator=gen)
    dense = [cu(p0.clone()), cu(torch.zeros(V, D)), cu(torch.zeros(V, D)), cu(torch.zeros(V, D))]       # p, g, m, v
    rows_ = [cu(p0.clone()), cu(torch.zeros(V, D)), cu(torch.zeros(V, D)), cu(torch.zeros(V, D))]
    bitmap, active = cu(torch.zeros((V + 31) // 32, dtype=torch.int32)), cu(torch.zeros(V, dtype=torch.int32))
    count = cu(torch.zeros(1, dtype=torch.int32))
    for step in range(1, 31):
        idx = torch.randint(0, 200 + 150 * step, (64,), generator=gen).clamp(max=V - 1).int()
        idx[0] = 0                                                     # the padding row: never marked, never updated
        gr = torch.randn(64, D, generator=gen)
        dense[1].zero_()
        K.call('xnrs_scatter_add_rows', dense[1], V, D, cu(idx), 64, cu(gr), D, 0)
        rows_[1].copy_(dense[1])            # the SAME gradient bits for both (an id drawn three times sums in atomic order)
        K.call('xnrs_mark_rows', cu(idx), 64, V, 0, bitmap, active, count)
        K.adam_step(dense[0].view(-1), dense[1].view(-1), dense[2].view(-1), dense[3].view(-1), 1e-2, step=step, grad_scale=0.5)
        K.call('xnrs_adam_rows', rows_[0], rows_[1], rows_[2], rows_[3], V, D, active, count, 1e-2, 0.9, 0.999, 1e-8, step, None, 0.5)
    n_active = int(count)
    assert 0 < n_active < V and 0 not in active[:n_active].tolist()
    assert len(set(active[:n_active].tolist())) == n_active                 # each row appended once
    for a, b in zip(dense, rows_):
        assert torch.equal(a, b)
    K.call('xnrs_zero_rows', rows_[1], V, D, active, count)
    assert float(rows_[1].abs().sum()) == 0


@pytest.mark.parametrize('bf', [False, True])
def test_warp_per_title_pool_backward_matches_the_cta_kernel(bf, monkeypatch):
    """(XNRS_POOL_BWD_WARP=1 must be in the environment before the library first dispatches this entry point for the warp
    kernel to be the one under test; otherwise both runs take the CTA kernel and the test checks it against float64.)
    the warp-per-title pooling backward (R >= 4096 ragged groups gathered from a table, fp32 or bf16 storage) against the
    CTA-per-title kernel on a slice of the same problem (R < 4096 takes the CTA kernel): d_hid rows, d_w2, d_b2, d_b1"""
    V, F_, A, R = 3000, 768, 256, 5000
    gen = g(13)
    table = torch.randn(V, F_, generator=gen) * 0.3
    lens = torch.randint(0, 31, (R,), generator=gen)
    lens[3] = 0
    seg = torch.zeros(R + 1, dtype=torch.int32)
    seg[1:] = torch.cumsum(lens, 0)
    T = int(seg[-1])
    n_rows = T + 57                                              # padded row buffer
    rows = torch.zeros(n_rows, dtype=torch.int32)
    rows[:T] = torch.randint(1, V, (T,), generator=gen).int()
    hid = torch.tanh(torch.randn(n_rows, A, generator=gen))
    w2 = torch.randn(A, generator=gen) / 16
    attn = torch.rand(n_rows, generator=gen)
    for r in range(R):
        a, b = int(seg[r]), int(seg[r + 1])
        if b > a:
            attn[a:b] /= attn[a:b].sum()
    d_pooled = torch.randn(R, F_, generator=gen)

    def run(R_):
        tb = cu(table.bfloat16() if bf else table)
        hd = cu(hid.bfloat16() if bf else hid)
        d_hid = torch.full((n_rows, A), 7.0, device=DEV, dtype=torch.bfloat16 if bf else torch.float32)
        d_w2, d_b2, d_b1 = cu(torch.zeros(A)), cu(torch.zeros(1)), cu(torch.zeros(A))
        nr = n_rows if R_ == R else int(seg[R_])
        if bf:
            K.call('xnrs_addpool_bwd_bf16', tb, cu(rows), hd, cu(w2), cu(attn), cu(d_pooled[:R_].contiguous()), cu(seg[:R_ + 1].contiguous()),
                   R_, 30, F_, A, nr, d_hid, d_w2, d_b2, d_b1)
        else:
            K.call('xnrs_addpool_bwd', tb, cu(rows), None, hd, cu(w2), cu(attn), cu(d_pooled[:R_].contiguous()), None,
                   cu(seg[:R_ + 1].contiguous()), R_, 30, F_, A, nr, d_hid, d_w2, d_b2, None, d_b1)
        return d_hid.float().cpu(), d_w2.cpu(), d_b2.cpu(), d_b1.cpu()

    big = run(R)                    # warp-per-title kernel
    small = run(3000)               # CTA-per-title kernel on the first 3000 titles
    t3 = int(seg[3000])
    tol = 2e-2 if bf else 1e-5
    assert_close(big[0][:t3], small[0][:t3], tol, 'd_hid rows of the shared titles')
    assert float(big[0][T:].abs().max()) == 0.0                  # padding rows are zeroed
    # full-problem reference for the accumulated gradients (float64)
    x = (table.bfloat16().double() if bf else table.double())[rows[:T].long()]
    h = (hid.bfloat16().double() if bf else hid.double())[:T]
    tix = torch.repeat_interleave(torch.arange(R), lens)
    da = (x * d_pooled.double()[tix]).sum(1)
    a = attn.double()[:T]
    sdot = torch.zeros(R, dtype=torch.float64).index_add_(0, tix, a * da)
    dlog = a * (da - sdot[tix])
    want_dhid = dlog[:, None] * w2.double()[None, :] * (1 - h * h)
    assert_close(big[0][:T], want_dhid.float(), 2e-2 if bf else 1e-4, 'd_hid')
    assert_close(big[1], (dlog[:, None] * h).sum(0).float(), 2e-2 if bf else 2e-4, 'd_w2')
    assert_close(big[3], want_dhid.sum(0).float(), 3e-2 if bf else 2e-4, 'd_b1')
    assert abs(float(big[2]) - float(dlog.sum())) <= 1e-3 * float(dlog.abs().sum())
