"""GPU parity tests proper: Encoder.forward and one full VQCPCEncoderTrainer step through the product path, against
(a) the golden vectors produced by the live reference and (b) the CPU oracle (fp32 and fp64) on the same inputs.

Bars (BASELINE.json north_star): code indices bit-exact; InfoNCE loss and encoder activations within 1e-4 relative.
Gradients have no stated bar and are ill-conditioned here: ANY fp32 implementation (the reference itself, the fp32
oracle) sits 2-3e-4 relative from the fp64 oracle on some parameters, i.e. ~1000x the fp32 unit round-off (which
parameters are hit depends on the summation order).  The CUDA path is therefore required to stay within
max(1e-3, 4x the fp32 oracle's worst parameter) in fp32 mode and within 6e-3 in bf16x3 mode (whose elementary
round-off, 2^-17 from the 16-bit hi+lo split, is ~100x fp32's; measured worst 2.9e-3).
"""
import os

import pytest
import torch

from oracle import params as P
from oracle import vqcpc_oracle as O
from tests.helpers import build_product_encoder, build_product_trainer, load_params_into, product_named_params, rel_err

pytestmark = pytest.mark.gpu
CASES = ['cfg1', 'cfg2_small', 'noup_d32', 'cfg1_bidir', 'lstm', 'lstm_uni', 'translin']
MODES = ['fp32', 'bf16x3']


def _gold(golden_dir, name):
    return torch.load(os.path.join(golden_dir, name + '.pt'), weights_only=False)


@pytest.mark.parametrize('mode', MODES)
@pytest.mark.parametrize('name', CASES)
def test_encoder_forward_vs_reference_golden(golden_dir, name, mode):
    from VQCPCB import ops
    from VQCPCB.utils import flatten
    ops.set_gemm_mode(mode)
    gold = _gold(golden_dir, name)
    cfg = O.Cfg(gold['cfg'])
    prm = P.make_params(cfg, seed=gold['param_seed'])
    batch = P.make_batch(cfg, gold['B'], seed=gold['batch_seed'])
    tr = build_product_trainer(cfg)
    load_params_into(tr, prm)
    enc = tr.encoder.eval()
    with torch.no_grad():
        x = batch['x_left']
        tok = enc.data_processor.preprocess(x)
        assert torch.equal(tok.cpu(), gold['enc']['tokens'].long())
        z = enc.downscaler(flatten(enc.data_processor.embed(tok)))
        zq, idx, ql = enc(x)
    torch.cuda.synchronize()
    ez = rel_err(z, gold['enc']['z'])
    mz = ((z.cpu() - gold['enc']['z']).abs().max() / gold['enc']['z'].abs().max()).item()
    print(f'[{name}/{mode}] z rel {ez:.2e} max-abs/max {mz:.2e}; zq rel {rel_err(zq, gold["enc"]["zq"]):.2e}')
    assert ez < 1e-4 and mz < 1e-4
    flips = int((idx.cpu() != gold['enc']['idx'].long()).sum())
    assert flips == 0, f'{flips} code indices differ from the reference'
    assert rel_err(zq, gold['enc']['zq']) < 1e-4
    assert rel_err(ql, gold['enc']['qloss']) < 1e-4
    ops.set_gemm_mode('bf16x3')


@pytest.mark.parametrize('mode', MODES)
def test_encoder_layerwise_vs_oracle(mode):
    """Activations after every encoder layer against the fp64 oracle (localises an error to a kernel)."""
    from VQCPCB import ops
    ops.set_gemm_mode(mode)
    cfg = O.default_cfg()
    prm = P.make_params(cfg, 0)
    prm64 = {k: v.double() for k, v in prm.items()}
    torch.manual_seed(0)
    tok = torch.randint(0, 64, (40, 1, 16))
    trace = {}
    emb = O.embed(tok, [prm64[f'data_processor.embeddings.{v}.weight'] for v in range(4)]).reshape(40, 16, -1)
    z_ref = O.downscaler_forward(emb, prm64, cfg, trace=trace)
    tr = build_product_trainer(cfg)
    load_params_into(tr, prm)
    down = tr.encoder.downscaler.eval()
    # truncated stacks reuse the same fused path: compare the final z (all layers) and, for localisation, run the
    # pieces by hand for the first layer
    with torch.no_grad():
        z = down(emb.float().cuda())
    e = rel_err(z, z_ref)
    print(f'[{mode}] downscaler z vs fp64 oracle: {e:.2e}')
    assert e < (2e-5 if mode == 'fp32' else 1e-4)
    with torch.no_grad():
        x0 = trace['h0'].float().cuda().reshape(-1, 512)
        lp = 'downscaler.transformers.0.layers.0.'
        g = lambda k: prm[lp + k].cuda()          # noqa: E731
        qkv = ops.linear_fwd(x0, g('self_attn.in_proj_weight'), g('self_attn.in_proj_bias'), scale_cols=(0.125, 512))
        o, _ = ops.attn_fwd(qkv, g('self_attn.attn_bias.e1'), g('self_attn.attn_bias.e2'), 40, 16, 8, 64)
        y1 = ops.linear_fwd(o, g('self_attn.out_proj.weight'), g('self_attn.out_proj.bias'), residual=x0)
        x1, _, _ = ops.layernorm_fwd(y1, g('norm1.weight'), g('norm1.bias'))
        hm = ops.linear_fwd(x1, g('linear1.weight'), g('linear1.bias'), relu=True)
        y2 = ops.linear_fwd(hm, g('linear2.weight'), g('linear2.bias'), residual=x1)
        x2, _, _ = ops.layernorm_fwd(y2, g('norm2.weight'), g('norm2.bias'))
    e1 = rel_err(x2.view(40, 16, 512), trace['x_0_0'])
    print(f'[{mode}] layer 0 output vs fp64 oracle: {e1:.2e}')
    assert e1 < (5e-6 if mode == 'fp32' else 5e-5)
    ops.set_gemm_mode('bf16x3')


@pytest.mark.parametrize('mode', MODES)
@pytest.mark.parametrize('name', CASES)
def test_training_step_vs_reference_and_oracle(golden_dir, name, mode):
    from VQCPCB import ops
    ops.set_gemm_mode(mode)
    gold = _gold(golden_dir, name)
    cfg = O.Cfg(gold['cfg'])
    prm = P.make_params(cfg, seed=gold['param_seed'])
    batch = P.make_batch(cfg, gold['B'], seed=gold['batch_seed'])
    lr = gold['step']['lr']
    # oracles: fp32 (same arithmetic class as the reference) and fp64 ("truth")
    o32, g32, gn32, newp32, _, _ = O.trainer_step(batch, prm, cfg, lr=lr)
    prm64 = {k: v.double() for k, v in prm.items()}
    o64, g64, gn64, _, _, _ = O.trainer_step(batch, prm64, cfg, lr=lr)

    tr = build_product_trainer(cfg)
    load_params_into(tr, prm)
    tr.init_optimizers(lr, False)
    tr.train()
    named = product_named_params(tr)
    before = {k: p.detach().clone() for k, p in named.items()}
    st = tr.step(batch, True, False)
    torch.cuda.synchronize()
    means = gold['step']['means']
    loss, nce, ql = st['loss'].item(), st['loss_contrastive'].item(), st['loss_quantize'].item()
    print(f'[{name}/{mode}] loss {loss:.6f} ref {means["loss"]:.6f} | nce {nce:.6f} ref {means["loss_contrastive"]:.6f} | '
          f'q {ql:.6f} ref {means["loss_quantize"]:.6f}')
    assert abs(loss - means['loss']) <= 1e-4 * abs(means['loss'])
    assert abs(nce - means['loss_contrastive']) <= 1e-4 * abs(means['loss_contrastive'])
    assert abs(ql - means['loss_quantize']) <= 1e-4 * abs(means['loss_quantize']) + 1e-7
    assert int(st['num_codewords'].item()) == int(means['num_codewords'])
    assert int(st['num_codewords_negative'].item()) == int(means['num_codewords_negative'])
    near = ((o32['pos'] - o32['neg'].max(2)[0]).abs() < 1e-4).sum(0).float() / gold['B']
    for a, r, n in zip(st['accuracy'].tolist(), means['accuracy'], near.tolist()):
        assert abs(a - r) <= n + 1e-6

    # gradients: read the flat gradient buffer (pre-clip values are still there after the fused update)
    gn = tr.optimizer.gnorm_sq.item() ** 0.5
    print(f'[{name}/{mode}] grad norm {gn:.5f} oracle32 {gn32.item():.5f} oracle64 {gn64.item():.5f}')
    assert abs(gn - gn64.item()) <= 1e-3 * gn64.item()
    worst = (0.0, '')
    live = [k for k in named if g64[k].norm() >= 1e-7 * gn64.item()]
    floor32 = max(rel_err(g32[k], g64[k]) for k in live)          # conditioning of this problem in fp32
    tol = max(1e-3, 4 * floor32) if mode == 'fp32' else 6e-3
    for k in live:
        e_cuda = rel_err(named[k].grad, g64[k])
        worst = max(worst, (e_cuda, k))
        assert e_cuda <= tol, f'{k}: cuda {e_cuda:.2e} (tol {tol:.1e}, fp32-oracle floor {floor32:.2e})'
    print(f'[{name}/{mode}] worst grad rel err vs fp64 oracle: {worst[0]:.2e} ({worst[1]}); fp32-oracle floor {floor32:.2e}')
    # Adam update (clip 5.0 inside).  A first Adam step moves every element by ~lr * sign(g): compare the deltas on the
    # elements whose gradient is not vanishing (sign flips of ~0 gradients are noise), element-wise.
    for k, p in named.items():
        d_cuda = (p.detach() - before[k]).cpu()
        d_ref = newp32[k] - prm[k]
        gk = g64[k].float()
        mask = gk.abs() > 1e-2 * gk.abs().max().clamp_min(1e-30)
        if mask.sum() == 0:
            continue
        assert rel_err(d_cuda[mask], d_ref[mask]) < 2e-2, (k, rel_err(d_cuda[mask], d_ref[mask]))
        assert d_cuda.abs().max().item() <= 1.01 * lr
    ops.set_gemm_mode('bf16x3')


def test_micro_batching_is_exact():
    """Splitting the batch into passes only changes fp32 summation order: same loss/metrics, gradients to 1e-5."""
    from VQCPCB import ops
    ops.set_gemm_mode('fp32')
    cfg = O.default_cfg(codebook_size=16, KL=2, KR=2)
    prm = P.make_params(cfg, 0)
    batch = P.make_batch(cfg, 8, seed=7)
    res = []
    for max_blocks in (10 ** 9, 2 * (2 + 2 + 15 * 2)):
        tr = build_product_trainer(cfg)
        load_params_into(tr, prm)
        tr.max_blocks_per_pass = max_blocks
        tr.init_optimizers(1e-4, False)
        tr.train()
        st = tr.step(batch, True, False)
        res.append((st, tr.optimizer.flat.grad.clone(), tr.optimizer.gnorm_sq.item()))
    (s0, g0, n0), (s1, g1, n1) = res
    assert abs(s0['loss'].item() - s1['loss'].item()) < 1e-5 * abs(s0['loss'].item())
    assert torch.equal(s0['accuracy'], s1['accuracy'])
    assert int(s0['num_codewords']) == int(s1['num_codewords'])
    assert rel_err(g1, g0) < 1e-5 and abs(n0 - n1) < 1e-4 * n0
    ops.set_gemm_mode('bf16x3')


def test_eval_epoch_and_checkpoint_roundtrip(tmp_path):
    from VQCPCB import ops
    cfg = O.default_cfg(codebook_size=16, KL=2, KR=2)
    prm = P.make_params(cfg, 0)
    tr = build_product_trainer(cfg, model_dir=str(tmp_path))
    load_params_into(tr, prm)
    tr.init_optimizers(1e-4, True)                      # with the LambdaLR schedule
    batches = [P.make_batch(cfg, 4, seed=s) for s in (1, 2)]
    m_train = tr.epoch(batches, True, 2, False)
    m_val = tr.epoch(batches, False, 1, False)
    for m in (m_train, m_val):
        assert set(m) == {'loss', 'accuracy', 'loss_quantize', 'loss_contrastive', 'num_codewords',
                          'num_codewords_negative', 'loss_monitor'}
        assert len(m['accuracy']) == cfg.KR and m['loss'] > 0
    assert abs(tr.optimizer.param_groups[0]['lr'] - 1e-4 * O.lr_lambda(2)) < 1e-12
    tr.save(early_stopped=False)
    for f in ('data_processor', 'downscaler', 'quantizer', 'upscaler', 'c_module', 'fks_module'):
        sd = torch.load(os.path.join(str(tmp_path), 'overfitted', f), map_location='cpu')
        ref_keys = set(P.split_by_file(prm)[f])
        assert set(sd) == ref_keys, f
    tr2 = build_product_trainer(cfg, model_dir=str(tmp_path))
    tr2.load(early_stopped=False, device='cuda')
    tr2.init_optimizers(1e-4, False)
    a = tr.epoch(batches, False, 1, False)
    b = tr2.epoch(batches, False, 1, False)
    assert abs(a['loss'] - b['loss']) < 1e-6 * abs(a['loss'])


def test_first_call_codebook_init_from_negatives():
    """Quirk Q5: with initialize=True the first Encoder.forward (the negatives batch) seeds the codebook with data rows."""
    cfg = O.default_cfg(codebook_size=16, KL=2, KR=2)
    prm = P.make_params(cfg, 0)
    tr = build_product_trainer(cfg)
    load_params_into(tr, prm)
    q = tr.encoder.quantizer
    q.initialize = True
    tr.init_optimizers(1e-4, False)
    tr.train()
    batch = P.make_batch(cfg, 4, seed=3)
    st = tr.step(batch, True, False)
    assert q.initialize is False and torch.isfinite(st['loss'])
    assert float(st['loss_quantize']) >= 0


def test_fused_bottleneck_matches_unfused():
    """Quantizer + upscaler folded into a per-code table (one kernel) vs the per-vector path: same indices, outputs to
    ulp level (upscale(q) vs upscale(z + (q - z))), gradients to 1e-5."""
    from VQCPCB import ops
    ops.set_gemm_mode('fp32')
    cfg = O.default_cfg(codebook_size=64, KL=2, KR=2)
    prm = P.make_params(cfg, 0)
    batch = P.make_batch(cfg, 6, seed=11)
    res = []
    for fuse in (False, True):
        tr = build_product_trainer(cfg)
        load_params_into(tr, prm)
        tr.encoder.fuse_bottleneck = fuse
        tr.encoder.quantizer.track_usage(True)
        tr.init_optimizers(1e-4, False)
        tr.train()
        assert tr.encoder._fusable(False) == fuse
        zq, idx, ql = tr.encoder(batch['x_right'])
        st = tr.step(batch, True, False)
        res.append((zq.detach(), idx, ql.detach(), st, tr.optimizer.flat.grad.clone(), tr.encoder.quantizer.usage_counts.clone()))
    (z0, i0, q0, s0, g0, c0), (z1, i1, q1, s1, g1, c1) = res
    assert torch.equal(i0, i1) and torch.equal(c0, c1)
    assert rel_err(z1, z0) < 1e-6 and rel_err(q1, q0) < 1e-6
    assert abs(s0['loss'].item() - s1['loss'].item()) < 1e-6 * abs(s0['loss'].item())
    assert rel_err(g1, g0) < 2e-5, rel_err(g1, g0)
    ops.set_gemm_mode('bf16x3')


def test_lstm_config_trains_with_dropout():
    """configs/encoder_vqcpc_lstm.py as shipped (dropout 0.1 in the GRUs, the upscaler and the context network): a few
    optimizer steps through the public factories run, stay finite and reduce the contrastive loss on a repeated batch
    (at the config's lr; at 10x that the commitment term of this model runs away, with or without dropout - the reference
    behaves the same)."""
    import importlib.util
    from VQCPCB import ops
    from VQCPCB.getters import get_dataloader_generator, get_encoder, get_encoder_trainer
    ops.set_gemm_mode('bf16x3')
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    path = os.path.join(root, 'vqcpc-bach_b200', 'VQCPCB', 'configs', 'encoder_vqcpc_lstm.py')
    spec = importlib.util.spec_from_file_location('cfg_lstm_train', path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    config = mod.config
    torch.manual_seed(0)
    gen = get_dataloader_generator(config['dataset'], config['training_method'], config['dataloader_generator_kwargs'])
    enc = get_encoder('/tmp/vqcpc_lstm_test', gen, config)
    tr = get_encoder_trainer('/tmp/vqcpc_lstm_test', gen, config['training_method'], enc, config['auxiliary_networks_kwargs'])
    tr.to('cuda')
    tr.init_optimizers(config['lr'], False)
    tr.train()
    assert enc.upscaler.dropout == 0.1 and not enc._fusable(False)      # per-vector path with dropout while training
    train, _, _ = gen.dataloaders(batch_size=16)
    batch = next(iter(train))
    losses = [tr.step(batch, True, False)['loss_contrastive'].item() for _ in range(80)]
    assert all(torch.isfinite(torch.tensor(losses)))
    first, last = sum(losses[:10]) / 10, sum(losses[-10:]) / 10
    assert last < first - 0.05, (first, last)
    tr.eval()
    assert enc._fusable(False)                                          # evaluation: fused table path
    with torch.no_grad():
        zq, idx, ql = enc(batch['x_left'])
    assert zq.shape == (16, 6, 32) and idx.shape == (16, 6, 1)


@pytest.mark.parametrize('mode', MODES)
@pytest.mark.parametrize('linear', [False, True])
def test_transformer_dropout_replay_vs_oracle(mode, linear):
    """Training-mode dropout in the fused encoder (attention weights, dropout1 / dropout2, FFN dropout - the reference's
    four sites per layer): the keep masks the kernels used are recorded and replayed through the fp64 oracle, so output
    AND gradients are checked exactly, including the dead-row-eliminated last layers (their masks live on kept rows)."""
    from VQCPCB import ops
    ops.set_gemm_mode(mode)
    p = 0.25
    cfg = O.default_cfg(codebook_size=16, KL=2, KR=2, downscaler='transformer_linear' if linear else 'transformer')
    prm = P.make_params(cfg, 3)
    tr = build_product_trainer(cfg)
    load_params_into(tr, prm)
    down = tr.encoder.downscaler
    down.dropout = p
    down.train()
    S = 16
    nb, b = 3, 2
    g = torch.Generator().manual_seed(5)
    emb = torch.randn(b, nb * S, cfg.embedding_size, generator=g)
    rec = {}

    def hook(site, shape, pp, device):
        m = (torch.rand(shape, generator=g) >= pp).to(torch.uint8)
        rec[site] = m
        return m.to(device)

    ops.DROPOUT_MASK_HOOK = hook
    try:
        x = emb.cuda().requires_grad_(True)
        z = down(x)
        gz = torch.randn(z.shape, generator=g)
        named = dict(down.named_parameters())
        for q in named.values():
            q.grad = None
        z.backward(gz.cuda())
        torch.cuda.synchronize()
    finally:
        ops.DROPOUT_MASK_HOOK = None
    # replay through the oracle in fp64: expand the compact (kept-row) masks of each stack's last layer to all rows
    Pn, d, H, Fh = b * nb, cfg.d_model, cfg.n_head, cfg.dim_feedforward
    drops, li, L = [], 0, S
    for factor, nl in zip(cfg.downscale_factors, cfg.list_of_num_layers):
        for l in range(nl):
            last = (l == nl - 1) and factor > 1 and not linear
            keep = factor if last else 1

            def full(site, width):
                m = rec[f'{site}{li}'].double().view(Pn, L // keep, width) / (1 - p)
                if keep == 1:
                    return m
                out = torch.ones(Pn, L, width, dtype=torch.float64)
                out[:, ::keep] = m
                return out
            drops.append(dict(attn=rec[f'attn{li}'].double() / (1 - p), drop1=full('drop1_', d), ffn=full('ffn', Fh),
                              drop2=full('drop2_', d)))
            li += 1
        L //= factor
    prm64 = {k: v.double().requires_grad_(True) for k, v in prm.items()}
    e64 = emb.double().requires_grad_(True)
    z64 = O.downscaler_forward(e64, prm64, cfg, drops=drops)
    z64.backward(gz.double())
    tol = 2e-5 if mode == 'fp32' else 1e-4
    assert rel_err(z, z64) < tol, rel_err(z, z64)
    # bf16x3: 16-bit significands on every GEMM operand; the worst parameter here (a bias gradient summed from the split
    # hi + lo halves of a small batch) sits at 6e-3, the same floor as the dropout-free training-step test
    gtol = 1e-3 if mode == 'fp32' else 1e-2
    assert rel_err(x.grad, e64.grad) < gtol, rel_err(x.grad, e64.grad)
    worst = (0.0, '')
    for k, q in named.items():
        r = prm64['downscaler.' + k].grad
        if r is None or r.norm() < 1e-9:
            continue
        worst = max(worst, (rel_err(q.grad, r), k))
    print(f'[dropout/{mode}/linear={linear}] z rel {rel_err(z, z64):.2e} worst grad rel {worst[0]:.2e} ({worst[1]})')
    assert worst[0] < gtol, worst
    ops.set_gemm_mode('bf16x3')
