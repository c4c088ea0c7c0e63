"""GPU: every non-GEMM kernel through the C-ABI against the oracle / torch autograd on the same seeded inputs."""
import os

import pytest
import torch
import torch.nn.functional as F

from oracle import vqcpc_oracle as O
from tests.helpers import rel_err

pytestmark = pytest.mark.gpu


def test_embed_fwd_bwd():
    from VQCPCB import ops
    torch.manual_seed(0)
    tables = [torch.randn(65, 32, device='cuda', requires_grad=True) for _ in range(4)]
    tok = torch.randint(0, 64, (37, 3, 16), device='cuda')
    out = ops.EmbedFn.apply(tok, *tables)
    ref = O.embed(tok.cpu(), [t.detach().cpu() for t in tables])
    assert torch.equal(out.cpu(), ref)                               # pure gather: bit-exact
    g = torch.randn_like(out)
    out.backward(g)
    tc = [t.detach().cpu().requires_grad_(True) for t in tables]
    O.embed(tok.cpu(), tc).backward(g.cpu())
    for a, b in zip(tables, tc):
        assert rel_err(a.grad, b.grad) < 1e-5


def test_posemb_fill_and_bwd():
    from VQCPCB import ops
    M, d, pe, S, V = 16 * 50, 512, 8, 16, 4
    h = torch.zeros(M, d, device='cuda')
    chan = torch.randn(V, pe, device='cuda')
    evt = torch.randn(S // V, pe, device='cuda')
    ops._native.call('vqcpc_posemb_fill', ops._p(h), d, d - 2 * pe, ops._p(chan), ops._p(evt), M, S, V, pe, ops._stream())
    t = torch.arange(M, device='cuda') % S
    assert torch.equal(h[:, d - 2 * pe:d - pe], chan[t % V]) and torch.equal(h[:, d - pe:], evt[t // V])
    assert h[:, :d - 2 * pe].abs().sum() == 0
    dh = torch.randn(M, d, device='cuda')
    dchan, devt = torch.zeros_like(chan), torch.zeros_like(evt)
    ops._native.call('vqcpc_posemb_bwd', ops._p(dh), d, d - 2 * pe, ops._p(dchan), ops._p(devt), M, S, V, pe, ops._stream())
    rc = torch.zeros_like(chan).index_add_(0, t % V, dh[:, d - 2 * pe:d - pe])
    re = torch.zeros_like(evt).index_add_(0, t // V, dh[:, d - pe:])
    assert rel_err(dchan, rc) < 1e-5 and rel_err(devt, re) < 1e-5


@pytest.mark.parametrize('M', [1, 77, 4099])
def test_layernorm_fwd_bwd(M):
    from VQCPCB import ops
    torch.manual_seed(1)
    d = 512
    y = (torch.randn(M, d, device='cuda') * 2 + 0.5).requires_grad_(True)
    g = (1 + 0.1 * torch.randn(d, device='cuda')).requires_grad_(True)
    b = (0.1 * torch.randn(d, device='cuda')).requires_grad_(True)
    out, mean, rstd = ops.layernorm_fwd(y.detach(), g.detach(), b.detach())
    ref = F.layer_norm(y.double(), (d,), g.double(), b.double(), 1e-5)
    assert rel_err(out, ref) < 1e-6
    go = torch.randn(M, d, device='cuda')
    ref.backward(go.double())
    dg, db = torch.zeros(d, device='cuda'), torch.zeros(d, device='cuda')
    dy = ops.layernorm_bwd(go, y.detach(), mean, rstd, g.detach(), dg, db)
    assert rel_err(dy, y.grad) < 2e-6 and rel_err(dg, g.grad) < 2e-6 and rel_err(db, b.grad) < 2e-6
    # fused variant: same outputs (+ bf16 split) and the column sums of dy accumulated into a bias-gradient buffer
    dg2, db2, cs = torch.zeros(d, device='cuda'), torch.zeros(d, device='cuda'), torch.full((d,), 0.25, device='cuda')
    dy2, sp = ops.layernorm_bwd(go, y.detach(), mean, rstd, g.detach(), dg2, db2, want_split=True, colsum_out=cs)
    assert torch.equal(dy2, dy) and rel_err(dg2, dg) < 1e-6 and rel_err(db2, db) < 1e-6
    assert rel_err(sp[0].float() + sp[1].float(), dy) < 1e-4
    ref_cs = 0.25 + y.grad.sum(0)
    assert (cs.double() - ref_cs).abs().max().item() < 1e-5 * max(1.0, y.grad.abs().sum(0).max().item())


def test_subsample_colsum_selu_reduce():
    from VQCPCB import ops
    P, L, f, d = 33, 16, 4, 512
    x = torch.randn(P * L, d, device='cuda')
    out = torch.empty(P * L // f, d, device='cuda')
    ops._native.call('vqcpc_subsample_fwd', ops._p(x), ops._p(out), P, L, f, d, ops._stream())
    assert torch.equal(out.view(P, L // f, d), x.view(P, L, d)[:, ::f])
    dx = torch.empty_like(x)
    ops._native.call('vqcpc_subsample_bwd', ops._p(out), ops._p(dx), P, L, f, d, ops._stream())
    ref = torch.zeros(P, L, d, device='cuda')
    ref[:, ::f] = out.view(P, L // f, d)
    assert torch.equal(dx.view(P, L, d), ref)
    acc = torch.ones(d, device='cuda')
    ops.colsum(x, acc)
    assert rel_err(acc, 1 + x.double().sum(0)) < 1e-5
    xs = torch.randn(1000, 37, device='cuda').requires_grad_(True)
    y = ops.SeluFn.apply(xs)
    assert rel_err(y, F.selu(xs.double())) < 1e-6
    gy = torch.randn_like(y)
    y.backward(gy)
    xd = xs.detach().double().requires_grad_(True)
    F.selu(xd).backward(gy.double())
    assert rel_err(xs.grad, xd.grad) < 1e-6
    v = torch.randn(100003, device='cuda')
    assert abs(ops.reduce_sum(v, 0.5).item() - 0.5 * v.double().sum().item()) < 1e-3


@pytest.mark.parametrize('L,P', [(16, 37), (4, 37), (4, 3), (8, 5), (2, 9), (1, 11)])
def test_attention_fwd_bwd_vs_oracle(L, P):
    from VQCPCB import ops
    torch.manual_seed(L * 100 + P)
    H, hd = 8, 64
    d = H * hd
    qkv = torch.randn(P * L, 3 * d, device='cuda') * 0.7
    e1 = torch.randn(H * L, hd, device='cuda') * 0.5
    e2 = torch.randn(H * L, hd, device='cuda') * 0.5
    out, probs = ops.attn_fwd(qkv, e1, e2, P, L, H, hd)
    torch.cuda.synchronize()

    def oracle(qkv_, e1_, e2_):
        q, k, v = qkv_.view(P, L, 3, H, hd).unbind(2)
        q, k, v = (t.transpose(1, 2) for t in (q, k, v))            # (P, H, L, hd); q already "scaled"
        s = q @ k.transpose(-1, -2) + O.rel_bias(q, e1_, e2_)
        a = torch.softmax(s, -1)
        return (a @ v).transpose(1, 2).reshape(P * L, d), a

    qd, e1d, e2d = (t.detach().double().cpu().requires_grad_(True) for t in (qkv, e1, e2))
    ro, ra = oracle(qd, e1d, e2d)
    assert rel_err(out, ro) < 2e-6, rel_err(out, ro)
    assert rel_err(probs, ra) < 2e-6
    go = torch.randn(P * L, d, device='cuda')
    ro.backward(go.double().cpu())
    de1, de2 = torch.zeros_like(e1), torch.zeros_like(e2)
    q_scale = 0.125
    dqkv = ops.attn_bwd(qkv, e1, e2, probs, go, de1, de2, P, L, H, hd, q_scale)
    torch.cuda.synchronize()
    ref = qd.grad.clone().view(P * L, 3, d)
    ref[:, 0] *= q_scale                                              # kernel returns grad wrt the unscaled projection
    assert rel_err(dqkv, ref.view(P * L, 3 * d)) < 5e-6, rel_err(dqkv, ref.view(P * L, 3 * d))
    assert rel_err(de1, e1d.grad) < 5e-6 and rel_err(de2, e2d.grad) < 5e-6


# (the last three cases give every warp several tiles: the cp.async prefetch of the next tile is live)
@pytest.mark.parametrize('L,P,keep', [(16, 37, 1), (16, 9, 4), (4, 37, 1), (4, 21, 4), (2, 9, 1), (16, 3001, 1), (16, 2501, 4),
                                      (4, 9001, 4)])
def test_attention_fwd_both_kernels_agree(L, P, keep):
    """The tensor-core forward (mma.sync 3xTF32, default) and the CUDA-core forward are two evaluations of the same fp32
    math: outputs, the bf16 split and the saved probabilities agree to fp32 round-off, with and without dead-row
    elimination and with a dropout keep mask."""
    from VQCPCB import ops
    torch.manual_seed(7 * L + P)
    H, hd = 8, 64
    d = H * hd
    qkv = torch.randn(P * L, 3 * d, device='cuda') * 0.7
    e1 = torch.randn(H * L, hd, device='cuda') * 0.5
    e2 = torch.randn(H * L, hd, device='cuda') * 0.5
    mask = (torch.rand(P, H, L, L, device='cuda') > 0.3).to(torch.uint8)
    res = []
    for algo in (0, 1):
        o, sp, pr = ops.attn_fwd(qkv, e1, e2, P, L, H, hd, keep=keep, want_split=True, algo=algo)
        od, _ = ops.attn_fwd(qkv, e1, e2, P, L, H, hd, keep=keep, drop=(mask, 1.0 / 0.7), algo=algo)
        torch.cuda.synchronize()
        res.append((o, sp[0].float() + sp[1].float(), pr[:, :, ::keep], od))
    for a, b in zip(*res):
        assert rel_err(a, b.double().cpu()) < 2e-6, rel_err(a, b.double().cpu())
    assert rel_err(res[0][1], res[0][0].double().cpu()) < 1e-4          # the split reproduces the fp32 output to 16 bits


@pytest.mark.parametrize('L,P,keep', [(16, 37, 1), (16, 9, 4), (8, 11, 2), (4, 37, 1), (4, 21, 4), (2, 9, 1), (1, 19, 1),
                                      (16, 3001, 1), (16, 2501, 4), (4, 9001, 4)])
def test_attention_bwd_both_kernels_agree(L, P, keep):
    """Tensor-core backward (default) vs CUDA-core backward on the same saved probabilities: d_qkv (fp32 and bf16 split),
    dE1 / dE2, with dead-row elimination and with a dropout keep mask."""
    from VQCPCB import ops
    torch.manual_seed(11 * L + P)
    H, hd = 8, 64
    d = H * hd
    qkv = torch.randn(P * L, 3 * d, device='cuda') * 0.7
    e1 = torch.randn(H * L, hd, device='cuda') * 0.5
    e2 = torch.randn(H * L, hd, device='cuda') * 0.5
    for drop in (None, ((torch.rand(P, H, L, L, device='cuda') > 0.3).to(torch.uint8), 1.0 / 0.7)):
        _, probs = ops.attn_fwd(qkv, e1, e2, P, L, H, hd, keep=keep, drop=drop, algo=1)
        go = torch.randn(P * (L // keep), d, device='cuda')
        res = {}
        for algo in (0, 1):                    # 0 = tensor cores, 1 = CUDA cores (the reference here)
            de1, de2 = torch.zeros_like(e1), torch.zeros_like(e2)
            dqkv, sp = ops.attn_bwd(qkv, e1, e2, probs, go, de1, de2, P, L, H, hd, 0.125, keep=keep, want_split=True, drop=drop,
                                    algo=algo)
            torch.cuda.synchronize()
            res[algo] = (dqkv, sp[0].float() + sp[1].float(), de1, de2)
        for algo in (0,):
            for a, b in zip(res[algo], res[1]):
                assert rel_err(a, b) < 3e-6, (algo, rel_err(a, b))
            assert rel_err(res[algo][1], res[algo][0]) < 1e-4


def test_quantizer_bit_exact_vs_reference_golden(golden_dir):
    from VQCPCB.quantizer.vector_quantizer import ProductVectorQuantizer
    gold = torch.load(os.path.join(golden_dir, 'quantizer.pt'), weights_only=False)
    for name, c in gold.items():
        vq = ProductVectorQuantizer(c['Kc'], c['D'], 0.25, c['ncb'], False, False, c['squared']).cuda()
        vq.load_state_dict({f'embeddings.{i}': b for i, b in enumerate(c['books'])}, strict=True)
        vq.eval()
        vq.track_usage(True)
        x = c['x'].cuda().requires_grad_(True)
        zsg, idx, loss = vq(x)
        assert torch.equal(idx.cpu(), c['idx'].long()), f'{name}: code indices differ from the reference'
        assert torch.equal(zsg.detach().cpu(), c['zsg']), f'{name}: straight-through output not bit-exact'
        assert rel_err(loss, c['loss']) < 1e-6
        ((loss * c['gl'].cuda()).sum() + (zsg * c['gz'].cuda()).sum()).backward()
        assert rel_err(x.grad, c['dx']) < 1e-5
        for p, db in zip(vq.embeddings, c['dbooks']):
            assert rel_err(p.grad, db) < 1e-4
        counts = vq.usage_counts.view(c['ncb'], c['Kc']).cpu()
        for cb in range(c['ncb']):
            assert torch.equal(counts[cb], torch.bincount(c['idx'].long()[..., cb].reshape(-1), minlength=c['Kc']).int())


def test_quantizer_full_size_properties():
    """BASELINE size (557 056 vectors, K=512, D=3): idempotence (codes quantise to themselves, zero loss) and the
    histogram sums to the vector count."""
    from VQCPCB.quantizer.vector_quantizer import ProductVectorQuantizer
    torch.manual_seed(0)
    n, Kc, D = 557056, 512, 3
    vq = ProductVectorQuantizer(Kc, D, 0.25, 1, False, False, True).cuda().eval()
    vq.track_usage(True)
    x = torch.randn(n // 8, 8, D, device='cuda') * 3
    zsg, idx, loss = vq(x)
    assert int(vq.usage_counts.sum()) == n
    q = vq.embeddings[0][idx.view(-1)].view_as(x)
    d_best = ((x - q) ** 2).sum(-1)
    probe = torch.randint(0, Kc, (n // 8, 8), device='cuda')
    d_probe = ((x - vq.embeddings[0][probe]) ** 2).sum(-1)
    assert bool((d_best <= d_probe + 1e-6).all())                      # no probed code is closer than the chosen one
    z2, idx2, loss2 = vq(q.detach())
    assert torch.equal(idx2, idx) or bool((loss2 == 0).all())          # duplicates may re-route, distance stays 0
    assert float(loss2.abs().max()) == 0.0


@pytest.mark.parametrize('algo,B', [(1, 19), (2, 19), (2, 2), (2, 301), (0, 128), (2, 4096)])
def test_infonce_fwd_bwd_vs_oracle(algo, B):
    """Both forward kernels (warp per (b, k); streaming cp.async.bulk ring) against the fp64 oracle, odd batch sizes
    (ragged last stage), and an exact tie between a negative and the positive."""
    from VQCPCB import ops
    torch.manual_seed(3)
    N, KR, zd, cd = 15, 8, 32, 32
    c = torch.randn(B, cd, device='cuda', requires_grad=True)
    W = (torch.randn(zd, cd, KR, device='cuda') * 0.25).requires_grad_(True)
    yp = torch.randn(B, KR, zd, device='cuda', requires_grad=True)
    yn = torch.randn(B, N, KR, zd, device='cuda')
    tb = min(3, B - 1)
    yn[tb, 5] = yp[tb].detach()                                          # exact tie: negative == positive
    yn.requires_grad_(True)
    ops._native.call('vqcpc_infonce_set_algo', algo)
    try:
        total, pos, neg, correct = ops.InfoNCEFn.apply(c, W, yp, yn)
        (total / B).backward()                                           # the same switch selects the backward kernel
    finally:
        ops._native.call('vqcpc_infonce_set_algo', 0)
    cd_, Wd, ypd, ynd = (t.detach().double().cpu().requires_grad_(True) for t in (c, W, yp, yn))
    rpos = O.fks(cd_, ypd, Wd)
    rneg = torch.stack([O.fks(cd_, ynd[:, n], Wd) for n in range(N)], 2)
    rl = O.nce_loss(rpos, rneg)
    rl.backward()
    assert abs((total / B).item() - rl.item()) < 1e-5 * abs(rl.item())
    assert rel_err(pos, rpos) < 1e-6 and rel_err(neg, rneg) < 1e-6
    assert not bool(correct[tb].any())                                   # ties are never "correct"
    rc = (rpos > rneg.max(2)[0])
    near = (rpos - rneg.max(2)[0]).abs() < 1e-5
    assert bool(((correct.cpu() != 0) == rc)[~near].all())
    for a, b in ((c, cd_), (W, Wd), (yp, ypd), (yn, ynd)):
        assert rel_err(a.grad, b.grad) < 2e-5, rel_err(a.grad, b.grad)


@pytest.mark.parametrize('N,KR,zd,cd,B', [(15, 8, 32, 32, 4096), (15, 4, 32, 32, 1001), (15, 2, 32, 32, 777), (15, 8, 32, 32, 3)])
def test_infonce_streaming_matches_warp_kernel(N, KR, zd, cd, B):
    """Full batch (and other shapes the configs use: KR = 2 for seq 64): the two forward kernels agree to rounding on
    every output, and exactly on the accuracy bits away from ties."""
    from VQCPCB import ops
    g = torch.Generator(device='cuda')
    g.manual_seed(B + N)
    c = torch.randn(B, cd, device='cuda', generator=g)
    W = torch.randn(zd, cd, KR, device='cuda', generator=g) * 0.25
    yp = torch.randn(B, KR, zd, device='cuda', generator=g)
    yn = torch.randn(B, N, KR, zd, device='cuda', generator=g)
    outs, grads = [], []
    for algo in (1, 2):
        leaves = [t.clone().requires_grad_(True) for t in (c, W, yp, yn)]
        ops._native.call('vqcpc_infonce_set_algo', algo)
        try:
            total, pos, neg, correct = ops.InfoNCEFn.apply(*leaves)
            (total / B).backward()
        finally:
            ops._native.call('vqcpc_infonce_set_algo', 0)
        outs.append((total.detach(), pos, neg, correct))
        grads.append([t.grad for t in leaves])
    for ga, gb in zip(*grads):
        assert rel_err(ga, gb) < 2e-5, rel_err(ga, gb)
    (t1, p1, n1, c1), (t2, p2, n2, c2) = outs
    assert rel_err(p1, p2) < 1e-6 and rel_err(n1, n2) < 1e-6
    assert abs(t1.item() - t2.item()) < 1e-5 * abs(t1.item())
    near = (p1 - n1.max(2)[0]).abs() < 1e-5
    assert bool((c1 == c2)[~near].all())


def test_clip_adam_matches_oracle():
    from VQCPCB.optim import FlatParams, FusedClipAdam
    torch.manual_seed(5)
    ps = [torch.nn.Parameter(torch.randn(300, 7, device='cuda')), torch.nn.Parameter(torch.randn(1001, device='cuda'))]
    prm = {str(i): p.detach().cpu().clone() for i, p in enumerate(ps)}
    opt = FusedClipAdam(FlatParams(ps), lr=1e-3)
    m = v = None
    for step in range(1, 4):
        grads = {str(i): torch.randn_like(p).cpu() * (10.0 if step == 1 else 0.01) for i, p in enumerate(ps)}
        opt.zero_grad()
        for i, p in enumerate(ps):
            p.grad.copy_(grads[str(i)])
        opt.step(max_norm=5.0)
        clipped, gn = O.clip_grad_norm(grads, 5.0)
        if m is None:
            m = {k: torch.zeros_like(t) for k, t in prm.items()}
            v = {k: torch.zeros_like(t) for k, t in prm.items()}
        prm, m, v = O.adam_step(prm, clipped, m, v, step, 1e-3)
        assert abs(opt.gnorm_sq.item() ** 0.5 - gn.item()) < 1e-4 * gn.item()
        for i, p in enumerate(ps):
            assert rel_err(p, prm[str(i)]) < 1e-6, (step, i)


def _unsplit(sp):
    return sp[0].float() + sp[1].float()


def test_split_outputs_of_layernorm_attention_colsum():
    """Producers can emit their result pre-split into bf16 (hi, lo): hi == bf16(x), hi + lo == x to 2^-16."""
    from VQCPCB import ops
    torch.manual_seed(11)
    M, d = 777, 512
    y = torch.randn(M, d, device='cuda')
    g, b = torch.randn(d, device='cuda'), torch.randn(d, device='cuda')
    out, sp, mean, rstd = ops.layernorm_fwd(y, g, b, want_split=True)
    assert torch.equal(sp[0], out.to(torch.bfloat16)) and ((_unsplit(sp) - out).abs() <= out.abs() * 2 ** -15 + 1e-30).all()
    none_out, sp2, _, _ = ops.layernorm_fwd(y, g, b, want_fp32=False, want_split=True)
    assert none_out is None and torch.equal(sp2[0], sp[0]) and torch.equal(sp2[1], sp[1])
    dg, db = torch.zeros(d, device='cuda'), torch.zeros(d, device='cuda')
    go = torch.randn(M, d, device='cuda')
    dy, dsp = ops.layernorm_bwd(go, y, mean, rstd, g, dg, db, want_split=True)
    assert torch.equal(dsp[0], dy.to(torch.bfloat16)) and rel_err(_unsplit(dsp), dy) < 1e-5
    acc = torch.zeros(d, device='cuda')
    ops._native.call('vqcpc_colsum_split', ops._p(dsp[0]), ops._p(dsp[1]), d, ops._p(acc), M, d, ops._stream())
    assert rel_err(acc, _unsplit(dsp).double().sum(0)) < 1e-5


@pytest.mark.parametrize('L,keep,P', [(16, 4, 21), (4, 4, 21), (16, 1, 5), (8, 2, 7)])
def test_attention_kept_rows_and_split_io(L, keep, P):
    """keep > 1 (dead-row elimination): outputs / dO live on the rows with token % keep == 0 only, compacted; results
    must equal the full computation restricted to those rows, and gradients the full backward with zero dO elsewhere."""
    from VQCPCB import ops
    torch.manual_seed(L * 10 + keep)
    H, hd = 8, 64
    d = H * hd
    qkv = torch.randn(P * L, 3 * d, device='cuda') * 0.7
    e1 = torch.randn(H * L, hd, device='cuda') * 0.5
    e2 = torch.randn(H * L, hd, device='cuda') * 0.5
    full, probs_full = ops.attn_fwd(qkv, e1, e2, P, L, H, hd)
    out, sp, probs = ops.attn_fwd(qkv, e1, e2, P, L, H, hd, keep=keep, want_split=True)
    ref = full.view(P, L, d)[:, ::keep].reshape(-1, d)
    assert torch.equal(out, ref)
    assert torch.equal(sp[0], ref.to(torch.bfloat16)) and rel_err(_unsplit(sp), ref) < 1e-5
    go = torch.randn(P * (L // keep), d, device='cuda')
    go_full = torch.zeros(P, L, d, device='cuda')
    go_full[:, ::keep] = go.view(P, L // keep, d)
    a1, a2 = torch.zeros_like(e1), torch.zeros_like(e2)
    b1, b2 = torch.zeros_like(e1), torch.zeros_like(e2)
    ref_dqkv = ops.attn_bwd(qkv, e1, e2, probs_full, go_full.view(-1, d), a1, a2, P, L, H, hd, 0.125)
    dqkv, dsp = ops.attn_bwd(qkv, e1, e2, probs, go, b1, b2, P, L, H, hd, 0.125, keep=keep, want_split=True)
    assert rel_err(dqkv, ref_dqkv) < 1e-6 and rel_err(b1, a1) < 1e-5 and rel_err(b2, a2) < 1e-5
    assert torch.equal(dsp[0], dqkv.to(torch.bfloat16)) and rel_err(_unsplit(dsp), dqkv) < 1e-5


def test_upscaler_dropout_replays_mask():
    """MlpUpscaler in training mode with dropout (mlp_upscaler.py:21-35: Linear -> Dropout -> SELU -> Linear): with the
    keep mask given, output and gradients match the same computation in fp64 torch; the default draw keeps ~1 - p."""
    from VQCPCB import ops
    from VQCPCB.upscalers.mlp_upscaler import MlpUpscaler
    torch.manual_seed(11)
    ops.set_gemm_mode('fp32')
    up = MlpUpscaler(3, 32, 512, 0.3).cuda().train()
    x = torch.randn(777, 3, device='cuda', requires_grad=True)
    keep = ops.draw_keep_mask((777, 512), 0.3, 'cuda')
    assert abs(keep.float().mean().item() - 0.7) < 0.01
    y = up(x, keep_mask=keep)
    g = torch.randn_like(y)
    y.backward(g)
    xd = x.detach().double().requires_grad_(True)
    w0, b0, w3, b3 = (t.detach().double().requires_grad_(True) for t in (up.mlp[0].weight, up.mlp[0].bias, up.mlp[3].weight, up.mlp[3].bias))
    h = F.linear(xd, w0, b0) * keep.double() / 0.7
    yr = F.linear(F.selu(h), w3, b3)
    yr.backward(g.double())
    assert rel_err(y, yr) < 1e-5
    assert rel_err(x.grad, xd.grad) < 1e-5
    for a, b in ((up.mlp[0].weight, w0), (up.mlp[0].bias, b0), (up.mlp[3].weight, w3), (up.mlp[3].bias, b3)):
        assert rel_err(a.grad, b.grad) < 1e-5
    # eval mode: dropout is the identity (and the fused table path takes over inside the Encoder)
    up.eval()
    assert rel_err(up(x.detach()), F.linear(F.selu(F.linear(x.detach().double(), w0.detach(), b0.detach())), w3.detach(), b3.detach())) < 1e-5
    ops.set_gemm_mode('bf16x3')


@pytest.mark.parametrize('mode', ['fp32', 'bf16x3'])
@pytest.mark.parametrize('B,T,inp,H,nl', [(37, 8, 32, 512, 2), (5, 16, 32, 96, 1), (300, 1, 32, 64, 2), (4096, 8, 32, 512, 2)])
def test_native_gru_matches_torch_gru(B, T, inp, H, nl, mode):
    """ops.GRUFn (native GEMMs + gate kernels) against torch.nn.GRU in fp64 on the CPU - the module the reference uses for
    the context network (vqcpc_helper.py:59) and the block downscaler (lstm_downscaler.py:31): outputs, input gradient and
    every parameter gradient."""
    from VQCPCB import ops
    ops.set_gemm_mode(mode)
    torch.manual_seed(B + T)
    gru = torch.nn.GRU(inp, H, nl, bias=True, batch_first=True, dropout=0.0).cuda()
    x = torch.randn(B, T, inp, device='cuda', requires_grad=True)
    y = ops.gru_forward(gru, x, True)
    g = torch.randn_like(y)
    for q in gru.parameters():
        q.grad = None
    y.backward(g)
    ref = torch.nn.GRU(inp, H, nl, bias=True, batch_first=True).double()
    ref.load_state_dict({k: v.detach().double().cpu() for k, v in gru.state_dict().items()})
    xr = x.detach().double().cpu().requires_grad_(True)
    yr, _ = ref(xr)
    yr.backward(g.double().cpu())
    tol = 2e-5 if mode == 'fp32' else 2e-4
    assert rel_err(y, yr) < tol, rel_err(y, yr)
    assert rel_err(x.grad, xr.grad) < 10 * tol, rel_err(x.grad, xr.grad)
    for (k, a), (_, b) in zip(gru.named_parameters(), ref.named_parameters()):
        assert rel_err(a.grad, b.grad) < 10 * tol, (k, rel_err(a.grad, b.grad))
    ops.set_gemm_mode('bf16x3')


def test_native_gru_dropout_replay():
    """Inter-layer dropout of the GRU (nn.GRU's ``dropout``): explicit keep mask on the first layer's outputs, checked
    against two stacked single-layer torch GRUs with the same mask in fp64."""
    from VQCPCB import ops
    ops.set_gemm_mode('fp32')
    torch.manual_seed(2)
    B, T, inp, H, p = 21, 8, 32, 128, 0.3
    gru = torch.nn.GRU(inp, H, 2, bias=True, batch_first=True, dropout=p).cuda()
    x = torch.randn(B, T, inp, device='cuda', requires_grad=True)
    rec = {}

    def hook(site, shape, pp, device):
        rec[site] = (torch.rand(shape, device=device) >= pp).to(torch.uint8)
        return rec[site]
    ops.DROPOUT_MASK_HOOK = hook
    try:
        y = ops.gru_forward(gru, x, True)
        g = torch.randn_like(y)
        y.backward(g)
    finally:
        ops.DROPOUT_MASK_HOOK = None
    sd = {k: v.detach().double().cpu() for k, v in gru.state_dict().items()}
    l0 = torch.nn.GRU(inp, H, 1, batch_first=True).double()
    l1 = torch.nn.GRU(H, H, 1, batch_first=True).double()
    l0.load_state_dict({k: sd[k] for k in sd if k.endswith('l0')})
    l1.load_state_dict({k.replace('l1', 'l0'): sd[k] for k in sd if k.endswith('l1')})
    xr = x.detach().double().cpu().requires_grad_(True)
    h0, _ = l0(xr)
    h0 = h0 * rec['gru0'].view(B, T, H).double().cpu() / (1 - p)
    yr, _ = l1(h0)
    yr.backward(g.double().cpu())
    assert rel_err(y, yr) < 2e-5 and rel_err(x.grad, xr.grad) < 2e-4
    assert rel_err(gru.weight_hh_l0.grad, l0.weight_hh_l0.grad) < 2e-4
    ops.set_gemm_mode('bf16x3')
