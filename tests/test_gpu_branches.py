"""GPU tests of the product-path branches no golden case reaches ('untested branches'):

* ``NoQuantization`` (vector_quantizer.py:15-24) through a full trainer step, against the oracle;
* ``corrupt_labels=True`` (vector_quantizer.py:118-132): the 5 % label corruption is drawn with torch's CUDA generator,
  so the test re-seeds, re-draws the very same random numbers and replays them through the oracle (like the dropout
  replay tests): indices, straight-through output, loss and all gradients are then checked exactly;
* ``main_encoder.py -t`` and ``-l`` (main_encoder.py:34-97 of the reference) end to end as a subprocess with an
  env-shrunk config: fresh run -> checkpoints of the reference's file names (compact, not the flat buffer) -> resume;
* code-index parity of ``Encoder.forward`` at the size of one production pass (34 816 blocks, K = 512) against the CPU
  oracle: flips are counted and every flipped vector must be a genuine near-tie in fp64.
"""
import os
import subprocess
import sys

import pytest
import torch

from oracle import params as P
from oracle import vqcpc_oracle as O
from tests.helpers import build_product_encoder, build_product_trainer, load_params_into, product_named_params, rel_err

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize('mode', ['fp32', 'bf16x3'])
def test_no_quantization_step_vs_oracle(mode):
    from VQCPCB import ops
    from VQCPCB.quantizer.vector_quantizer import NoQuantization
    ops.set_gemm_mode(mode)
    cfg = O.default_cfg(codebook_size=None, KL=2, KR=2)
    prm = P.make_params(cfg, 0)
    batch = P.make_batch(cfg, 6, seed=21)
    ref, grads, gnorm, newp, _, _ = O.trainer_step(batch, prm, cfg, lr=1e-3)
    prm64 = {k: v.double() for k, v in prm.items()}
    ref64, grads64, _, _, _, _ = O.trainer_step(batch, prm64, cfg, lr=1e-3)
    tr = build_product_trainer(cfg)
    assert isinstance(tr.encoder.quantizer, NoQuantization)
    load_params_into(tr, prm)
    tr.init_optimizers(1e-3, False)
    tr.train()
    zq, idx, ql = tr.encoder(batch['x_left'])
    assert idx is None and float(ql.abs().sum()) == 0.0 and zq.shape == (6, cfg.KL, cfg.z_dim)
    assert rel_err(zq, ref['z_left']) < 1e-4
    st = tr.step(batch, True, False)
    assert 'num_codewords' not in st
    assert abs(st['loss'].item() - ref['loss'].item()) < 1e-4 * abs(ref['loss'].item())
    assert float(st['loss_quantize']) == 0.0
    named = product_named_params(tr)
    flat = tr.optimizer.flat
    sl = flat.named_slices(named.items())
    worst = 0.0
    for k, (o, n) in sl.items():
        g64 = grads64[k].reshape(-1)
        if g64.norm() < 1e-12:
            continue
        worst = max(worst, rel_err(flat.grad[o:o + n], g64))
    assert worst < (1e-3 if mode == 'fp32' else 6e-3), worst
    # Adam moved the parameters like the oracle's Adam
    for k in ('fks_module.W', 'downscaler.output_linear.weight'):
        o, n = sl[k]
        d_ours = flat.data[o:o + n].cpu() - prm[k].reshape(-1)
        d_ref = newp[k].reshape(-1) - prm[k].reshape(-1)
        assert rel_err(d_ours, d_ref) < 5e-2
    # eval epoch without quantizer: metrics dict keeps the reference's keys, code counts are 0
    m = tr.epoch([batch], False, 1, False)
    assert m['num_codewords'] == 0 and m['num_codewords_negative'] == 0 and m['loss'] > 0
    ops.set_gemm_mode('bf16x3')


@pytest.mark.parametrize('ncb,D', [(1, 3), (2, 4)])
def test_corrupt_labels_replayed_through_oracle(ncb, D):
    """vector_quantizer.py:118-132.  Same seed -> same CUDA draws -> the corrupted indices are known to the test."""
    from VQCPCB import ops
    from VQCPCB.quantizer.vector_quantizer import ProductVectorQuantizer
    Kc, n = 64, 4000
    g = torch.Generator().manual_seed(3)
    z = (torch.randn(n, D, generator=g) * 1.5)
    books = [torch.randn(Kc, D // ncb, generator=g) * 1.5 for _ in range(ncb)]
    vq = ProductVectorQuantizer(Kc, D, 0.25, ncb, False, False, True).cuda()
    for e, b in zip(vq.embeddings, books):
        e.data.copy_(b)
    vq.train()
    zc = z.cuda().requires_grad_(True)
    torch.manual_seed(77)
    zsg, idx, loss = vq(zc, corrupt_labels=True)
    # replay of the two draws (same generator state, same calls, same shapes)
    torch.manual_seed(77)
    like = torch.empty(n, ncb, dtype=torch.int64, device='cuda')
    rnd = torch.randint_like(like, low=0, high=Kc)
    keep = torch.rand(like.shape, device='cuda') > 0.05
    _, clean, _ = O.quantize(z, books, 0.25, True)
    expect = torch.where(keep.cpu(), clean.view(n, ncb), rnd.cpu())
    assert torch.equal(idx.cpu(), expect)
    frac = 1.0 - keep.float().mean().item()
    assert 0.03 < frac < 0.07, frac
    assert int((expect != clean.view(n, ncb)).sum()) > 0
    # oracle on the given (corrupted) indices: output, loss and gradients
    z64 = z.double().requires_grad_(True)
    b64 = [b.double().requires_grad_(True) for b in books]
    o_zsg, o_idx, o_loss = O.quantize(z64, b64, 0.25, True, given_idx=expect)
    w = torch.randn(n, D, generator=g).double()
    (o_zsg * w).sum().backward(retain_graph=True)
    (o_loss.sum() * 0.37).backward()
    ((zsg * w.float().cuda()).sum() + loss.sum() * 0.37).backward()
    assert rel_err(zsg, o_zsg) < 1e-6 and rel_err(loss, o_loss) < 1e-5
    assert rel_err(zc.grad, z64.grad) < 1e-5
    for e, b in zip(vq.embeddings, b64):
        assert rel_err(e.grad, b.grad) < 1e-5
    # eval mode ignores the flag (the reference corrupts only while training)
    vq.eval()
    _, idx_eval, _ = vq(zc.detach(), corrupt_labels=True)
    assert torch.equal(idx_eval.cpu().view(n, ncb), clean.view(n, ncb))


def test_corrupt_labels_trainer_step_runs_unfused():
    """trainer.step(corrupt_labels=True) takes the per-call path (negatives corrupted, positives clean: :201 vs :227-230)."""
    cfg = O.default_cfg(codebook_size=16, KL=2, KR=2)
    prm = P.make_params(cfg, 0)
    batch = P.make_batch(cfg, 4, seed=5)
    tr = build_product_trainer(cfg)
    load_params_into(tr, prm)
    tr.init_optimizers(1e-4, False)
    tr.train()
    assert tr.encoder._fusable(True) is False
    clean = tr.step(batch, False, False)
    torch.manual_seed(1)
    st = tr.step(batch, True, True)
    assert torch.isfinite(st['loss']) and float(st['loss_quantize']) > float(clean['loss_quantize'])
    # positives are never corrupted -> same positive code set as the clean step
    assert int(st['num_codewords']) == int(clean['num_codewords'])


def _run_main(args, cwd, env):
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'main_encoder.py')] + args, cwd=cwd, env=env, capture_output=True,
                       text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + '\n' + r.stderr[-3000:]
    return r.stdout


def test_main_encoder_train_then_resume(tmp_path):
    """`main_encoder.py -t -c cfg` (fresh run: data-dependent codebook init, 2 epochs, checkpoints) then
    `main_encoder.py -t -l -c <model_dir>/config.py` (resume from the 'overfitted' checkpoint)."""
    env = dict(os.environ, VQCPC_BATCH='8', VQCPC_BLOCKS='2', VQCPC_CODEBOOK='16', VQCPC_NUM_BATCHES='2', VQCPC_EPOCHS='2')
    cfg_path = os.path.join(ROOT, 'vqcpc-bach_b200', 'VQCPCB', 'configs', 'encoder_vqcpc.py')
    out = _run_main(['-t', '-c', cfg_path], str(tmp_path), env)
    assert '======= Epoch 1 =======' in out and '---Val---' in out
    models = os.listdir(tmp_path / 'models')
    assert len(models) == 1 and models[0].startswith('encoder_vqcpc_')
    mdir = tmp_path / 'models' / models[0]
    assert (mdir / 'config.py').exists()
    files = ('data_processor', 'downscaler', 'quantizer', 'upscaler', 'c_module', 'fks_module')
    sizes = {}
    for sub in ('overfitted', 'early_stopped'):
        for f in files:
            path = mdir / sub / f
            assert path.exists(), path
            sd = torch.load(path, map_location='cpu')
            payload = sum(v.numel() * v.element_size() for v in sd.values())
            sizes[f] = os.path.getsize(path)
            # compact per-module files (views of the flat buffer used to drag the whole model along)
            assert sizes[f] < payload + 64 * 1024, (f, sizes[f], payload)
    assert sizes['quantizer'] < 64 * 1024 and sizes['downscaler'] > 40e6
    cb = torch.load(mdir / 'overfitted' / 'quantizer', map_location='cpu')['embeddings.0']
    assert cb.shape == (16, 3) and torch.isfinite(cb).all()
    # resume: `-l` takes model_dir = dirname(config) and must NOT re-initialise the codebook (main_encoder.py:45-51)
    env2 = dict(env, VQCPC_EPOCHS='1')
    out2 = _run_main(['-t', '-l', '-c', str(mdir / 'config.py')], str(tmp_path), env2)
    assert 'Loading models' in out2 and '======= Epoch 0 =======' in out2
    import re
    losses = [float(x) for x in re.findall(r'Loss: ([-+0-9.eE]+)', out + out2)]
    assert losses and all(l == l and abs(l) < 1e4 for l in losses), losses


def _oracle_z(tokens, prm, cfg, dtype, chunk=2048):
    tables = [prm[f'data_processor.embeddings.{v}.weight'].to(dtype) for v in range(cfg.num_voices)]
    p2 = {k: v.to(dtype) for k, v in prm.items()}
    outs = []
    with torch.no_grad():
        for i in range(0, tokens.shape[0], chunk):
            tok = tokens[i:i + chunk]
            emb = O.embed(tok, tables)                                  # (c, 16, E): one block per row
            z = O.downscaler_forward(emb, p2, cfg)
            outs.append(z.reshape(-1, cfg.codebook_dim))
    return torch.cat(outs, 0)


_PARITY_CACHE = {}


def _parity_reference(nblk, Kc, n64):
    """tokens, oracle z (fp32 on all blocks, fp64 on the first n64), data-drawn codebook - computed once per session."""
    key = (nblk, Kc, n64)
    if key not in _PARITY_CACHE:
        cfg = O.default_cfg(codebook_size=Kc, KL=2, KR=2)
        prm = P.make_params(cfg, 0)
        g = torch.Generator().manual_seed(99)
        tokens = torch.randint(0, 64, (nblk, 16), generator=g, dtype=torch.int64)
        z32 = _oracle_z(tokens, prm, cfg, torch.float32)
        z64 = _oracle_z(tokens[:n64], prm, cfg, torch.float64, chunk=1024)
        E = z32[torch.randperm(nblk, generator=g)[:Kc]].clone()       # data-dependent init: codes ARE data rows
        _PARITY_CACHE[key] = (cfg, prm, tokens, z32, z64, E)
    return _PARITY_CACHE[key]


@pytest.mark.parametrize('mode', ['fp32', 'bf16x3'])
def test_index_parity_at_pass_scale(mode):
    """Encoder.forward over one production pass (34 816 blocks of 16 tokens, K = 512, codebook drawn from the data as
    in a fresh run) vs the CPU oracle.  Kernel-level bit-exactness (same z in, same index out) is proven elsewhere; here
    z itself carries the GEMM mode's rounding (fp32: summation order, ~1e-6; bf16x3: ~7e-6 of max|z|), so a vector that
    sits within that distance of a bisector plane may legitimately fall on the other side.  Required: every flipped
    vector is such a near-tie under the fp64 oracle, and flips stay at the 1e-4 level."""
    from VQCPCB import ops
    ops.set_gemm_mode(mode)
    nblk, Kc, n64 = 34816, 512, 8192
    cfg, prm, tokens, z32, z64, E = _parity_reference(nblk, Kc, n64)
    prm = dict(prm)
    prm['quantizer.embeddings.0'] = E
    _, idx32, _ = O.quantize(z32, [E], 0.25, True)
    idx32 = idx32.view(-1)
    d64 = ((z64.unsqueeze(1) - E.double().unsqueeze(0)) ** 2).sum(-1)
    idx64 = d64.argmin(1)
    enc = build_product_encoder(cfg).cuda()
    sd = P.split_by_file(prm)
    for f, mod in (('data_processor', enc.data_processor), ('downscaler', enc.downscaler), ('quantizer', enc.quantizer),
                   ('upscaler', enc.upscaler)):
        mod.load_state_dict({k: v.cuda() for k, v in sd[f].items()}, strict=True)
    ops.bump_weight_version()
    enc.eval()
    with torch.no_grad():
        zq, idx, ql = enc.forward_tokens(tokens.cuda().unsqueeze(0))
        from VQCPCB.utils import flatten
        z_gpu = enc.downscaler(flatten(enc.data_processor.embed(tokens.cuda().unsqueeze(0)))).reshape(-1, 3)
    idx = idx.view(-1).cpu()
    zerr = ((z_gpu.cpu() - z32).abs().max() / z32.abs().max()).item()
    flips32 = (idx != idx32).nonzero().view(-1)
    flips64 = (idx[:n64] != idx64).nonzero().view(-1)
    oracle_self = int((idx32[:n64] != idx64).sum())                    # fp32 oracle vs fp64 oracle: the noise floor
    # margin of every flipped vector under fp64 arithmetic on the oracle's own z: relative gap between the two codes
    dz = ((z32.double().unsqueeze(1)[flips32] - E.double().unsqueeze(0)) ** 2).sum(-1) if len(flips32) else torch.zeros(0, Kc, dtype=torch.float64)
    # a perturbation dz of the vector moves its squared distance to a code by <= 2 |z - e| |dz|, so two codes can only
    # swap if their fp64 distance gap is below 2 (|z - e_a| + |z - e_b|) |dz|; |dz| <= sqrt(3) * max-abs error (x2 slack)
    dz_abs = 2.0 * (3 ** 0.5) * zerr * z32.abs().max().item()
    margins = []
    for r, v in enumerate(flips32.tolist()):
        a, b = dz[r, idx32[v]].item(), dz[r, idx[v]].item()
        margins.append(abs(b - a) / max(2.0 * (a ** 0.5 + b ** 0.5) * dz_abs, 1e-30))
    print(f'[index parity {mode}] {nblk} vectors, K={Kc}: max|z - z_oracle|/max|z| = {zerr:.2e}; flips vs fp32 oracle '
          f'{len(flips32)} ({len(flips32) / nblk:.2e}); vs fp64 oracle (first {n64}) {len(flips64)}; fp32-oracle vs '
          f'fp64-oracle flips {oracle_self}; worst flipped gap / admissible gap {max(margins) if margins else 0.0:.2e}')
    assert zerr < (2e-5 if mode == 'fp32' else 1e-4)
    assert all(m <= 1.0 for m in margins), margins
    assert len(flips32) <= (3 if mode == 'fp32' else 12), len(flips32)
    assert len(flips64) <= oracle_self + (3 if mode == 'fp32' else 6)
    # the straight-through outputs agree wherever the index agrees
    same = idx == idx32
    T_or = O.upscale(E, prm)
    assert rel_err(zq.view(-1, 32).cpu()[same], T_or[idx32[same]]) < 1e-5
    ops.set_gemm_mode('bf16x3')


def test_ema_codebook_option_matches_cpu_restatement():
    """Optional EMA codebook (north_star extension; the reference has none -> parity unpinned): three trainer steps with
    ema_decay = 0.9 against the CPU restatement of the rule fed with the product's own z / indices; and the default
    (ema_decay=None) stays the reference's gradient codebook."""
    from VQCPCB import ops
    ops.set_gemm_mode('fp32')
    cfg = O.default_cfg(codebook_size=64, KL=2, KR=2)
    prm = P.make_params(cfg, 0)
    tr = build_product_trainer(cfg)
    load_params_into(tr, prm)
    q = tr.encoder.quantizer
    q.ema_decay = 0.9
    tr.init_optimizers(1e-4, False)
    tr.train()
    E0 = q.embeddings[0].detach().clone()
    N = torch.ones(64, dtype=torch.float64)
    m = E0.double().cpu()
    seen = []
    orig = q.ema_accumulate

    def spy(flat_z, idx):
        seen.append((flat_z.detach().reshape(-1, 3).double().cpu(), idx.reshape(-1).cpu()))
        return orig(flat_z, idx)

    q.ema_accumulate = spy
    for s in range(3):
        seen.clear()
        st = tr.step(P.make_batch(cfg, 6, seed=40 + s), True, False)
        assert torch.isfinite(st['loss'])
        z = torch.cat([a for a, _ in seen])
        ix = torch.cat([b for _, b in seen])
        assert z.shape[0] == 6 * (cfg.KL + cfg.KR + cfg.N * cfg.KR)
        N, m, E = O.ema_codebook_update(N, m, z, ix, 0.9, q.ema_eps)
        assert rel_err(q.embeddings[0], E) < 1e-5, (s, rel_err(q.embeddings[0], E))
        assert rel_err(q._ema_N, N) < 1e-6
    assert not torch.equal(E0, q.embeddings[0].detach())
    assert int(q._ema_cnt.sum()) == 0                       # cleared after the update
    # default: gradient codebook, no EMA state
    tr2 = build_product_trainer(cfg)
    assert tr2.encoder.quantizer.ema_decay is None and tr2.encoder.quantizer._ema_N is None
    ops.set_gemm_mode('bf16x3')
