"""CPU: the oracle restatement (oracle/vqcpc_oracle.py) against golden vectors produced by the live reference
(oracle/make_golden.py).  This is what pins the oracle; the GPU parity tests then compare CUDA with the oracle."""
import os

import pytest
import torch

from oracle import params as P
from oracle import vqcpc_oracle as O

CASES = ['cfg1', 'cfg2_small', 'noup_d32', 'cfg1_bidir', 'lstm', 'lstm_uni', 'translin']


def _load(golden_dir, name):
    return torch.load(os.path.join(golden_dir, name + '.pt'), weights_only=False)


def _fingerprint(t, seed=7):
    t = t.detach().float().reshape(-1)
    g = torch.Generator()
    g.manual_seed(seed)
    r = torch.randn(t.numel(), generator=g)
    return t.norm().item(), (t.double() @ r.double()).item(), t[:8]


@pytest.mark.parametrize('name', CASES)
def test_encoder_forward_matches_reference(golden_dir, name):
    gold = _load(golden_dir, name)
    cfg = O.Cfg(gold['cfg'])
    prm = P.make_params(cfg, seed=gold['param_seed'])
    batch = P.make_batch(cfg, gold['B'], seed=gold['batch_seed'])
    tok = O.preprocess(batch['x_left'], cfg.num_voices * cfg.ticks_per_block)
    assert torch.equal(tok, gold['enc']['tokens'].long())
    trace = {}
    zq, idx, ql = O.encoder_forward(batch['x_left'], prm, cfg, trace=trace)
    assert torch.equal(idx, gold['enc']['idx'].long())            # bit-exact code indices
    torch.testing.assert_close(trace['z'], gold['enc']['z'], rtol=1e-4, atol=2e-5)
    torch.testing.assert_close(zq, gold['enc']['zq'], rtol=1e-4, atol=2e-5)
    torch.testing.assert_close(ql, gold['enc']['qloss'], rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize('name', CASES)
def test_training_step_matches_reference(golden_dir, name):
    gold = _load(golden_dir, name)
    cfg = O.Cfg(gold['cfg'])
    prm = P.make_params(cfg, seed=gold['param_seed'])
    batch = P.make_batch(cfg, gold['B'], seed=gold['batch_seed'])
    out, grads, gnorm, newp, m, v = O.trainer_step(batch, prm, cfg, lr=gold['step']['lr'])
    means = gold['step']['means']
    assert abs(out['loss'].item() - means['loss']) <= 1e-4 * abs(means['loss'])
    assert abs(out['nce'].item() - means['loss_contrastive']) <= 1e-4 * abs(means['loss_contrastive'])
    assert abs(out['qloss'].item() - means['loss_quantize']) <= 1e-4 * abs(means['loss_quantize']) + 1e-7
    # accuracy = mean(pos > max neg): a negative that drew the positive's code scores equal up to the ulp-level
    # rounding of z + (q - z) (vector_quantizer.py:148), so such near-ties may legitimately fall either way.
    near = ((out['pos'] - out['neg'].max(2)[0]).abs() < 1e-5).sum(0).float() / gold['B']
    for a, r, n in zip(out['accuracy'].tolist(), means['accuracy'], near.tolist()):
        assert abs(a - r) <= n + 1e-6
    n_codes = len(torch.unique(torch.cat((out['idx_left'], out['idx_right']), 0)))
    assert n_codes == int(means['num_codewords'])
    assert len(torch.unique(out['idx_negative'])) == int(means['num_codewords_negative'])
    assert abs(gnorm.item() - gold['gnorm']) <= 1e-3 * gold['gnorm']
    coef = min(1.0, 5.0 / (gnorm.item() + 1e-6))
    # clipped gradients and the Adam update, by fingerprint (norm + random projection + first entries)
    for k, fp in gold['step']['clipped_grads'].items():
        n, pr, head = _fingerprint(grads[k] * coef)
        assert abs(n - fp['norm']) <= 2e-3 * fp['norm'] + 1e-7, k
        assert abs(pr - fp['proj']) <= 2e-3 * fp['norm'] * (fp['numel'] ** 0.5) ** 0 + 1e-3 * fp['norm'] + 1e-7, k
    for k, fp in gold['step']['delta'].items():
        n, pr, head = _fingerprint(newp[k] - prm[k])
        assert abs(n - fp['norm']) <= 2e-2 * fp['norm'] + 1e-9, k   # Adam's g/sqrt(g^2) amplifies sign noise of ~0 grads


def test_quantizer_matches_reference(golden_dir):
    gold = _load(golden_dir, 'quantizer')
    for name, c in gold.items():
        x = c['x'].clone().requires_grad_(True)
        books = [b.clone().requires_grad_(True) for b in c['books']]
        zsg, idx, loss = O.quantize(x, books, 0.25, c['squared'])
        assert torch.equal(idx, c['idx'].long()), name
        torch.testing.assert_close(loss, c['loss'], rtol=1e-6, atol=1e-6)
        torch.testing.assert_close(zsg, c['zsg'], rtol=0, atol=0)
        ((loss * c['gl']).sum() + (zsg * c['gz']).sum()).backward()
        torch.testing.assert_close(x.grad, c['dx'], rtol=1e-5, atol=1e-6)
        for b, db in zip(books, c['dbooks']):
            torch.testing.assert_close(b.grad, db, rtol=1e-4, atol=1e-5)


def test_rel_bias_closed_form_small():
    """Closed form vs. a literal per-element evaluation."""
    torch.manual_seed(0)
    P_, H, L, hd = 3, 2, 4, 8
    q = torch.randn(P_, H, L, hd)
    e1 = torch.randn(H * L, hd)
    e2 = torch.randn(H * L, hd)
    b = O.rel_bias(q, e1, e2)
    for p in range(P_):
        for h in range(H):
            for i in range(L):
                for j in range(L):
                    e = e1[h * L + (L - 1 - (i - j))] if j <= i else e2[h * L + (j - i)]
                    assert abs(b[p, h, i, j].item() - float(q[p, h, i] @ e)) < 1e-5


def test_lr_lambda_shape():
    assert abs(O.lr_lambda(0) - 0.1) < 1e-12
    assert abs(O.lr_lambda(10000) - 1.0) < 1e-9
    assert O.lr_lambda(10 ** 7) == 0.1


def test_tf32_split_error_bound():
    """The attention kernels split an fp32 value for the 3xTF32 products with one integer add + mask
    (csrc/attention.cu::tf32_split): big = x rounded to 10 explicit mantissa bits (nearest, ties away), small = x - big
    (exact in fp32), of which the tensor core reads the top 10 explicit mantissa bits.  Emulated here bit for bit:
    |x - big - tf32(small)| <= 2^-21 |x|, and the three-term product a_big b_big + a_big b_small + a_small b_big
    misses a*b by less than 2^-19.9 |a b| - the figure quoted in DESIGN.md / attention.cu.  (Magnitudes whose small part
    would be denormal, |x| < 2^-115, are outside the claim: activations never get there.)"""
    import numpy as np
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.standard_normal(200000).astype(np.float32) * 3, np.float32(2.0) ** rng.integers(-60, 60, 2000),
                        np.array([1.0, -1.0, 1.0009765625, 0.99951171875, 3.4e38, 1e-30], dtype=np.float32)]).astype(np.float32)

    def split(v):
        bits = v.view(np.uint32)
        big = ((bits + np.uint32(0x1000)) & np.uint32(0xFFFFE000)).view(np.float32)
        small = (v - big).astype(np.float32)
        assert np.all((big.astype(np.float64) + small.astype(np.float64)) == v.astype(np.float64))   # the subtraction is exact
        small_tc = (small.view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)                     # what the MMA reads
        return big.astype(np.float64), small_tc.astype(np.float64)

    big, small = split(x)
    err = np.abs(x.astype(np.float64) - big - small)
    assert np.all(err <= 2.0 ** -21 * np.abs(x.astype(np.float64)))
    a, b = x[:100000], x[100000:200000]
    ab, asm = split(a)
    bb, bsm = split(b)
    prod = ab * bb + ab * bsm + asm * bb
    exact = a.astype(np.float64) * b.astype(np.float64)
    rel = np.abs(prod - exact) / np.abs(exact)
    assert rel.max() < 2.0 ** -19.9, rel.max()
