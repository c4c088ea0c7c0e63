"""Index-gather form of the quantizer -> upscaler -> InfoNCE path (csrc/idxpath.cu) against

* a plain fp64 PyTorch restatement of the per-vector math of the reference (vector_quantizer.py:72-83,:148,
  vqcpc_helper.py:5-29,:79-98) with autograd - forward values and EVERY gradient (dc, dW, dT, dE, dz);
* the per-vector product kernels of round 1 (materialised upscaled rows) through a full trainer step, small and at the
  BASELINE size (B = 4096, KR = 8, N = 15, K = 512).
"""
import ctypes

import pytest
import torch

from oracle import params as P
from oracle import vqcpc_oracle as O
from tests.helpers import build_product_trainer, load_params_into, rel_err

pytestmark = pytest.mark.gpu


def _reference(z, E, T, J, c, W, idx, B, N, KR, KL, beta, squared, w_q, w_left):
    """fp64 autograd restatement.  Vector order [neg (B, N, KR) | left (B, KL) | right (B, KR)].  The upscaled vector of
    code k is T[k] and d y / d z = J[k] (straight-through: the upscaler sees z + (q - z).detach())."""
    z, E, T, c, W = [t.double().clone().requires_grad_(True) for t in (z, E, T, c, W)]
    J = J.double()
    q = E[idx]
    y = T[idx] + torch.einsum('pzd,pd->pz', J[idx], z - z.detach())
    if squared:
        e_l = ((q.detach() - z) ** 2).sum(-1)
        q_l = ((q - z.detach()) ** 2).sum(-1)
    else:
        e_l = torch.norm((q.detach() - z) + 1e-5, dim=-1)
        q_l = torch.norm((q - z.detach()) + 1e-5, dim=-1)
    lvec = q_l + beta * e_l
    n_neg, n_left = B * N * KR, B * KL
    y_neg = y[:n_neg].view(B, N, KR, -1)
    y_left = y[n_neg:n_neg + n_left]
    y_right = y[n_neg + n_left:].view(B, KR, -1)
    u = torch.einsum('bc,zck->bkz', c, W)
    pos = (u * y_right).sum(-1)
    neg = torch.einsum('bkz,bnkz->bkn', u, y_neg)
    lse = torch.logsumexp(torch.cat([neg, pos.unsqueeze(2)], 2), 2)
    nll = (lse - pos).sum()
    loss = nll / B + w_q * lvec.sum() + (y_left * w_left.double()).sum()
    loss.backward()
    return dict(pos=pos.detach(), neg=neg.detach(), lse=lse.detach(), nll=nll.detach(), lvec=lvec.detach(), y_left=y_left.detach(),
                dz=z.grad, dE=E.grad, dT=T.grad, dc=c.grad, dW=W.grad, correct=(pos > neg.max(2)[0]).detach())


@pytest.mark.parametrize('algos', [(0, 0), (1, 1)])
@pytest.mark.parametrize('B,N,KR,KL,Kc,D,squared', [(5, 15, 8, 8, 512, 3, 1), (37, 15, 2, 2, 64, 3, 1), (9, 7, 6, 6, 300, 3, 0),
                                                     (3, 15, 3, 2, 16, 4, 1), (130, 15, 8, 8, 512, 3, 1), (1, 1, 1, 1, 8, 3, 1),
                                                     (6, 15, 16, 3, 600, 3, 1)])
def test_idxpath_small_kernels_vs_fp64_autograd(B, N, KR, KL, Kc, D, squared, algos):
    """algos = (infonce backward, quantizer backward) kernel choice: (0, 0) production (tcgen05 dT; streaming dz + dE kernels),
    (1, 1) the CUDA-core / single-kernel variants kept as fallbacks."""
    from VQCPCB import _native, ops
    _native.call('vqcpc_infonce_idx_set_bwd_algo', algos[0])
    _native.call('vqcpc_vq_idx_set_bwd_algo', algos[1])
    dev = 'cuda'
    g = torch.Generator().manual_seed(B * 131 + KR)
    zd = cd = 32
    n_neg, n_left, n_right = B * N * KR, B * KL, B * KR
    Pn = n_neg + n_left + n_right
    E = torch.randn(Kc, D, generator=g) * 1.5
    z = torch.randn(Pn, D, generator=g) * 1.5
    T = torch.randn(Kc, zd, generator=g) * 0.5
    J = torch.randn(Kc, zd, D, generator=g) * 0.3
    c = torch.randn(B, cd, generator=g)
    W = torch.randn(zd, cd, KR, generator=g) * 0.2
    w_left = torch.randn(n_left, zd, generator=g) * 0.01
    beta, w_q = 0.25, 0.37 / (3 * B)
    _, idx, _ = O.quantize(z, [E], beta, bool(squared))
    idx = idx.view(-1)
    # a few duplicated candidates: a negative that drew the positive's code must score bit-equal (never "correct")
    for b in range(1, B, 2):
        z[(b * N + 0) * KR + 0] = z[n_neg + n_left + b * KR + 0]
    _, idx, _ = O.quantize(z, [E], beta, bool(squared))
    idx = idx.view(-1)
    ref = _reference(z, E, T, J, c, W, idx, B, N, KR, KL, beta, bool(squared), w_q, w_left)

    assert ops.idxpath_supported(N, KR, zd, cd, Kc, D)
    zc = z.to(dev).requires_grad_(True)
    Ec, Tc, Jc = E.to(dev).requires_grad_(True), T.to(dev).requires_grad_(True), J.to(dev)
    cc, Wc = c.to(dev).requires_grad_(True), W.to(dev).requires_grad_(True)
    link = ops.BottleneckLink()
    link.table = Tc
    counts = torch.zeros(Kc, dtype=torch.int32, device=dev)
    y_left, idx_g, lvec, token = ops.VQIndexFn.apply(zc, Ec, Tc, Jc, beta, squared, counts, n_neg, n_left, link)
    link.token = token
    assert torch.equal(idx_g.cpu(), idx)
    assert torch.equal(counts.cpu().long(), torch.bincount(idx, minlength=Kc))
    idx_neg, idx_left, idx_right = idx_g.split([n_neg, n_left, n_right])
    total, pos, neg, correct = ops.InfoNCEIdxFn.apply(cc, Wc, Tc, token, idx_right.view(B, KR), idx_neg.view(B, N, KR), link,
                                                      n_neg + n_left, 0)
    loss = total / B + w_q * lvec.sum() + (y_left * w_left.to(dev)).sum()
    loss.backward()
    torch.cuda.synchronize()
    assert rel_err(pos, ref['pos']) < 2e-6 and rel_err(neg, ref['neg']) < 2e-6
    assert abs(total.item() - ref['nll'].item()) < 2e-6 * abs(ref['nll'].item()) + 1e-6
    assert rel_err(lvec, ref['lvec']) < 1e-6 and rel_err(y_left, ref['y_left']) < 1e-7
    # accuracy bits: equal wherever the fp64 margin is not a rounding tie; exact duplicates are never correct
    margin = (ref['pos'] - ref['neg'].max(2)[0]).abs()
    clear = margin > 1e-5 * (ref['pos'].abs() + 1)
    assert torch.equal(correct.cpu().bool()[clear], ref['correct'][clear])
    for b in range(1, B, 2):
        assert int(correct[b, 0]) == 0
    _native.call('vqcpc_infonce_idx_set_bwd_algo', -1)
    _native.call('vqcpc_vq_idx_set_bwd_algo', -1)
    for name, got, tol in (('dc', cc.grad, 2e-5), ('dW', Wc.grad, 2e-5), ('dT', Tc.grad, 2e-5), ('dE', Ec.grad, 2e-5),
                           ('dz', zc.grad, 2e-5)):
        e = rel_err(got, ref[name])
        assert e < tol, (name, e)


def test_table_scatter_add_small_and_duplicates():
    from VQCPCB import _native, ops
    g = torch.Generator(device='cuda').manual_seed(1)
    for n, Kc, zd in ((1000, 512, 32), (33, 7, 32), (5000, 1000, 64), (1, 3, 32)):
        idx = torch.randint(0, Kc, (n,), device='cuda', generator=g)
        idx[: n // 3] = idx[0]                           # heavy duplicates
        rows = torch.randn(n, zd, device='cuda', generator=g)
        out = torch.randn(Kc, zd, device='cuda', generator=g)
        want = out.double().index_add(0, idx, rows.double())
        _native.call('vqcpc_table_scatter_add', ops._p(idx), n, ops._p(rows), zd, Kc, ops._p(out), ops._stream())
        assert rel_err(out, want) < 1e-6
        # deterministic: a second run from the same start gives the same bits
        a, b = torch.zeros(Kc, zd, device='cuda'), torch.zeros(Kc, zd, device='cuda')
        for o in (a, b):
            _native.call('vqcpc_table_scatter_add', ops._p(idx), n, ops._p(rows), zd, Kc, ops._p(o), ops._stream())
        assert torch.equal(a, b)


@pytest.mark.parametrize('mode', ['fp32', 'bf16x3'])
def test_trainer_step_small_index_path_equals_per_vector_path(mode):
    """Same step through the index path and through the per-vector kernels (materialised rows): identical indices /
    accuracy, loss to 1e-6, every gradient to 2e-5 of the flat gradient norm."""
    from VQCPCB import ops
    ops.set_gemm_mode(mode)
    cfg = O.default_cfg(codebook_size=64, KL=2, KR=2)
    prm = P.make_params(cfg, 0)
    batch = P.make_batch(cfg, 6, seed=11)
    res = []
    for use_idx in (False, True):
        tr = build_product_trainer(cfg)
        load_params_into(tr, prm)
        tr.index_path = use_idx
        tr.init_optimizers(1e-4, False)
        tr.train()
        st = tr.step(batch, True, False)
        res.append((st, tr.optimizer.flat.grad.clone()))
    (s0, g0), (s1, g1) = res
    assert abs(s0['loss'].item() - s1['loss'].item()) < 2e-6 * abs(s0['loss'].item())
    assert torch.equal(s0['accuracy'], s1['accuracy'])
    assert int(s0['num_codewords']) == int(s1['num_codewords']) and int(s0['num_codewords_negative']) == int(s1['num_codewords_negative'])
    assert rel_err(g1, g0) < 2e-5, rel_err(g1, g0)
    ops.set_gemm_mode('bf16x3')


def test_idxpath_full_size_matches_per_vector_kernels():
    """BASELINE size (B = 4096, KR = KL = 8, N = 15, K = 512, D = 3): index kernels vs the round-1 per-vector kernels
    on the same inputs - scores / loss / accuracy and all gradients."""
    from VQCPCB import _native, ops
    dev = 'cuda'
    B, N, KR, KL, Kc, D, zd, cd = 4096, 15, 8, 8, 512, 3, 32, 32
    n_neg, n_left, n_right = B * N * KR, B * KL, B * KR
    Pn = n_neg + n_left + n_right
    g = torch.Generator(device=dev).manual_seed(0)
    E = torch.randn(Kc, D, device=dev, generator=g) * 2
    z = torch.randn(Pn, D, device=dev, generator=g) * 2
    T = torch.randn(Kc, zd, device=dev, generator=g) * 0.5
    J = torch.randn(Kc, zd, D, device=dev, generator=g) * 0.3
    c = torch.randn(B, cd, device=dev, generator=g)
    W = torch.randn(zd, cd, KR, device=dev, generator=g) * 0.2
    w_left = torch.randn(n_left, zd, device=dev, generator=g) * 1e-3
    beta, w_q = 0.25, 0.5 / (3 * B)

    def run(indexed):
        zc, Ec, Tc = z.clone().requires_grad_(True), E.clone().requires_grad_(True), T.clone().requires_grad_(True)
        cc, Wc = c.clone().requires_grad_(True), W.clone().requires_grad_(True)
        if indexed:
            link = ops.BottleneckLink()
            link.table = Tc
            y_left, idx, lvec, token = ops.VQIndexFn.apply(zc, Ec, Tc, J, beta, 1, None, n_neg, n_left, link)
            link.token = token
            i_neg, i_left, i_right = idx.split([n_neg, n_left, n_right])
            total, pos, neg, corr = ops.InfoNCEIdxFn.apply(cc, Wc, Tc, token, i_right.view(B, KR), i_neg.view(B, N, KR), link,
                                                           n_neg + n_left, 0)
        else:
            y, idx, lvec = ops.FusedVQUpscaleFn.apply(zc, Ec, Tc, J, beta, 1, None)
            idx = idx.view(-1)
            y_neg, y_left, y_right = y.split([n_neg, n_left, n_right])
            total, pos, neg, corr = ops.InfoNCEFn.apply(cc, Wc, y_right.reshape(B, KR, zd), y_neg.reshape(B, N, KR, zd))
        loss = total / B + w_q * lvec.sum() + (y_left.reshape(n_left, zd) * w_left).sum()
        loss.backward()
        torch.cuda.synchronize()
        return dict(idx=idx, total=total.detach(), pos=pos, neg=neg, corr=corr, lvec=lvec.detach(), dz=zc.grad, dE=Ec.grad,
                    dT=Tc.grad, dc=cc.grad, dW=Wc.grad)

    a, b = run(False), run(True)
    assert torch.equal(a['idx'], b['idx'])
    assert torch.equal(a['lvec'], b['lvec'])
    assert rel_err(b['pos'], a['pos']) < 2e-6 and rel_err(b['neg'], a['neg']) < 2e-6
    assert abs(a['total'].item() - b['total'].item()) < 1e-6 * abs(a['total'].item())
    # accuracy bits may differ only on rounding-level ties between the two summation orders
    diff = (a['corr'] != b['corr'])
    assert int(diff.sum()) <= 4
    for k, tol in (('dz', 2e-5), ('dE', 2e-5), ('dT', 2e-5), ('dc', 2e-5), ('dW', 2e-5)):
        e = rel_err(b[k], a[k])
        assert e < tol, (k, e)
