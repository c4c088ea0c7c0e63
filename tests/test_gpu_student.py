"""GPU: the student / distillation trainer against the CPU oracle and the reference goldens."""
import os

import pytest
import torch

from oracle import student_oracle as S
from oracle import vqcpc_oracle as O
from tests.helpers import rel_err

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), 'golden', 'student.pt')


def _build(cfg, device='cuda'):
    import types
    from VQCPCB.auxiliary_decoders.auxiliary_decoder_relative import AuxiliaryDecoderRelative
    from VQCPCB.data_processor.bach_data_processor import BachDataProcessor
    from VQCPCB.downscalers.relative_transformer_downscaler_linear import RelativeTransformerDownscalerLinear
    from VQCPCB.encoder import Encoder
    from VQCPCB.quantizer.vector_quantizer import ProductVectorQuantizer
    from VQCPCB.student_encoder_trainer import StudentEncoderTrainer
    from VQCPCB.teachers.teacher_relative import TeacherRelative
    V = cfg.num_voices
    dp = BachDataProcessor(cfg.embedding_size, cfg.num_events, list(cfg.vocab))
    down = RelativeTransformerDownscalerLinear(cfg.embedding_size, cfg.codebook_dim, V, list(cfg.downscale_factors), cfg.d_model,
                                               cfg.n_head, list(cfg.list_of_num_layers), cfg.dim_feedforward, 0.0)
    vq = ProductVectorQuantizer(cfg.codebook_size, cfg.codebook_dim, cfg.commitment_cost, cfg.num_codebooks, False, False,
                                cfg.squared_l2_norm)
    enc = Encoder('/tmp/vqcpc_b200_student', dp, down, vq, None)
    tdp = BachDataProcessor(cfg.embedding_size, cfg.num_events, list(cfg.vocab))
    teacher = TeacherRelative(data_processor=tdp, num_layers=cfg.teacher_layers, num_tokens_per_channel=list(cfg.vocab),
                              positional_embedding_size=cfg.teacher_pos_emb, d_model=cfg.d_model, dim_feedforward=cfg.dim_feedforward,
                              n_head=cfg.n_head, num_tokens=cfg.num_events * V, dropout=0.0)
    dec = AuxiliaryDecoderRelative(num_tokens_per_channel=list(cfg.vocab), codebook_dim=cfg.codebook_dim,
                                   upscale_factors=list(reversed(cfg.downscale_factors)), list_of_num_layers=list(cfg.dec_layers),
                                   n_head=cfg.n_head, d_model=cfg.d_model, dim_feedforward=cfg.dim_feedforward,
                                   num_tokens_bottleneck=cfg.num_events * V // 16, dropout=0.0)
    tr = StudentEncoderTrainer('/tmp/vqcpc_b200_student', types.SimpleNamespace(), enc, cfg.num_events_masked, teacher, dec,
                               cfg.quantization_weighting)
    tr.to(device)
    return tr


def _groups(tr):
    e = tr.encoder
    return {'data_processor': e.data_processor, 'downscaler': e.downscaler, 'quantizer': e.quantizer, 'teacher': tr.teacher,
            'decoder': tr.auxiliary_decoder}


def _load(tr, prm):
    from VQCPCB import ops
    for g, mod in _groups(tr).items():
        mod.load_state_dict({k[len(g) + 1:]: v.cuda() for k, v in prm.items() if k.startswith(g + '.')}, strict=True)
    ops.bump_weight_version()


def _named(tr):
    return {f'{g}.{k}': p for g, mod in _groups(tr).items() for k, p in mod.named_parameters()}


@pytest.mark.parametrize('L,P,H', [(24, 3, 2), (96, 2, 8), (384, 2, 2), (100, 1, 1)])
def test_attn_long_fwd_bwd_vs_oracle(L, P, H):
    """Full-sequence relative attention (csrc/attn_long.cu) against the fp64 closed form of the oracle, incl. all gradients."""
    from VQCPCB import ops
    hd, d = 64, 64 * H
    g = torch.Generator().manual_seed(L)
    qkv = torch.randn(P * L, 3 * d, generator=g) * 0.5
    e1 = torch.randn(H * L, hd, generator=g) * 0.3
    e2 = torch.randn(H * L, hd, generator=g) * 0.3
    dout = torch.randn(P * L, d, generator=g)
    # oracle in fp64
    q64, e164, e264 = qkv.double().requires_grad_(True), e1.double().requires_grad_(True), e2.double().requires_grad_(True)
    q, k, v = q64.view(P, L, 3 * d).split(d, dim=-1)
    qh, kh, vh = (t.reshape(P, L, H, hd).transpose(1, 2) for t in (q, k, v))
    s = qh @ kh.transpose(-1, -2) + O.rel_bias(qh, e164, e264)
    a = torch.softmax(s, dim=-1)
    o = (a @ vh).transpose(1, 2).reshape(P * L, d)
    o.backward(dout.double())
    out, probs = ops.attn_fwd(qkv.cuda(), e1.cuda(), e2.cuda(), P, L, H, hd, need_probs=True)
    assert rel_err(out, o) < 2e-6
    assert rel_err(probs, a) < 2e-6
    de1, de2 = torch.zeros_like(e1).cuda(), torch.zeros_like(e2).cuda()
    dqkv = ops.attn_bwd(qkv.cuda(), e1.cuda(), e2.cuda(), probs, dout.cuda(), de1, de2, P, L, H, hd, 1.0)
    assert rel_err(dqkv, q64.grad) < 5e-6
    assert rel_err(de1, e164.grad) < 5e-6
    assert rel_err(de2, e264.grad) < 5e-6
    # split outputs carry the same values (hi + lo)
    _, sp, _ = ops.attn_fwd(qkv.cuda(), e1.cuda(), e2.cuda(), P, L, H, hd, need_probs=False, want_fp32=False, want_split=True)
    assert rel_err(sp[0].float() + sp[1].float(), out) < 1e-4


def test_softmax_ce_hard_and_distilled():
    from VQCPCB import ops
    g = torch.Generator().manual_seed(0)
    x = torch.randn(37, 61, generator=g) * 3
    y = torch.randn(37, 61, generator=g) * 2
    t = torch.randint(0, 61, (37,), generator=g)
    x64 = x.double().requires_grad_(True)
    ref = torch.nn.functional.cross_entropy(x64, t, reduction='none')
    ref.sum().backward()
    loss, dl = ops.softmax_ce(x.cuda(), target_idx=t.cuda())
    assert rel_err(loss, ref) < 1e-6 and rel_err(dl, x64.grad) < 1e-6
    x64 = x.double().requires_grad_(True)
    ref = -(torch.softmax(y.double(), 1) * torch.log_softmax(x64, 1)).sum(1)
    ref.sum().backward()
    loss, dl = ops.softmax_ce(x.cuda(), target_logits=y.cuda())
    assert rel_err(loss, ref) < 1e-6 and rel_err(dl, x64.grad) < 1e-6


@pytest.mark.parametrize('mode', ['fp32', 'bf16x3'])
@pytest.mark.parametrize('case', ['small', 'mid'])
def test_student_step_vs_reference_golden_and_oracle(case, mode):
    """One StudentEncoderTrainer batch (teacher step, then encoder + auxiliary-decoder step) on the CUDA path: monitored losses
    against the reference golden, code indices bit-exact, every gradient against the fp64 oracle, the Adam update against the
    reference's parameter deltas."""
    from VQCPCB import ops
    from oracle.make_golden import summarize
    rec = torch.load(GOLD)[case]
    cfg = S.student_cfg(**rec['spec']['cfg'])
    prm = S.make_student_params(cfg, seed=0)
    batch = S.make_student_batch(cfg, rec['spec']['B'], seed=1234)
    event = rec['masked_event_index']
    prm64 = {k: v.double() for k, v in prm.items()}
    _, g64, _, idx64 = S.student_step(batch, prm64, cfg, event, lr=rec['lr'])
    ops.set_gemm_mode(mode)
    try:
        tr = _build(cfg)
        _load(tr, prm)
        # parameter names = the reference's (strict load above) and nothing extra
        assert set(_named(tr)) == set(prm)
        tr.init_optimizers(rec['lr'], False)
        tr.train()
        # forward-only parity first
        with torch.no_grad():
            zq, idx, ql = tr.encoder(batch['x'])
            assert torch.equal(idx.cpu(), rec['indices'])
            assert rel_err(zq, rec['zq']) < 1e-5
            dl = tr.auxiliary_decoder(zq)
            tol = 2e-5 if mode == 'fp32' else 2e-4
            for a, b in zip(dl, rec['decoder_logits_event']):
                assert rel_err(a[:, event], b) < tol
        before = {k: p.detach().clone() for k, p in _named(tr).items()}
        q = tr.step(batch, train=True, masked_event_index=event)
        for k in ('loss_teacher', 'loss_quantization', 'loss_reconstruction', 'loss_encdec'):
            assert abs(q[k] - rec['means'][k]) <= 1e-4 * abs(rec['means'][k]), (k, q[k], rec['means'][k])
        # gradients (the .grad views hold the UNclipped sums; clipping happens inside the fused update)
        gtol = 5e-4 if mode == 'fp32' else 5e-3
        for k, p in _named(tr).items():
            ref = g64[k]
            if ref.norm() == 0:
                assert p.grad is None or p.grad.abs().max().item() == 0, k
                continue
            assert rel_err(p.grad, ref) < gtol, (k, rel_err(p.grad, ref))
        # Adam update vs the reference's deltas (fingerprints)
        for k, p in _named(tr).items():
            d, r = summarize((p.detach() - before[k]).cpu(), k), rec['delta'][k]
            assert abs(d['norm'] - r['norm']) <= 2e-2 * max(r['norm'], 1e-12), (k, d['norm'], r['norm'])
    finally:
        ops.set_gemm_mode('bf16x3')


def test_student_getters_and_epoch_smoke(tmp_path):
    """getters -> StudentEncoderTrainer from the shipped configuration (shrunk) and two epochs of two batches: losses finite,
    the teacher loss decreases on a repeated batch, checkpoints of the reference's file names are written and reload."""
    import copy
    from VQCPCB import getters
    from VQCPCB.configs.encoder_student import config as base
    cfg = copy.deepcopy(base)
    cfg['dataloader_generator_kwargs']['sequences_size'] = 8              # 32 events = 128 tokens
    cfg['downscaler_kwargs'].update(d_model=128, n_head=2, dim_feedforward=256, list_of_num_layers=[1, 1])
    tk = cfg['auxiliary_networks_kwargs']['teacher_kwargs']
    tk.update(num_layers=1, d_model=128, dim_feedforward=256, n_head=2)
    ak = cfg['auxiliary_networks_kwargs']['auxiliary_decoder_kwargs']
    ak.update(d_model=128, dim_feedforward=256, n_head=2, list_of_num_layers=[1, 1])
    cfg['quantizer_kwargs']['initialize'] = False
    gen = getters.get_dataloader_generator(cfg['dataset'], cfg['training_method'], cfg['dataloader_generator_kwargs'])
    enc = getters.get_encoder(str(tmp_path), gen, cfg)
    tr = getters.get_encoder_trainer(str(tmp_path), gen, cfg['training_method'], enc, cfg['auxiliary_networks_kwargs'])
    tr.to('cuda')
    tr.init_optimizers(1e-3, True)
    train, _, _ = gen.dataloaders(batch_size=4)
    b = next(train)
    first = tr.step(b, train=True, masked_event_index=5)
    for _ in range(5):
        last = tr.step(b, train=True, masked_event_index=5)
    assert all(torch.isfinite(torch.tensor(v)) for v in last.values())
    assert last['loss_teacher'] < first['loss_teacher']
    means = tr.epoch(gen.dataloaders(batch_size=4)[1], train=False, num_batches=2)
    assert set(means) == {'loss_teacher', 'loss_quantization', 'loss_reconstruction', 'loss_encdec', 'loss_monitor'}
    tr.save(early_stopped=False)
    for f in ('data_processor', 'downscaler', 'quantizer', 'decoder', 'teacher'):
        assert os.path.exists(tmp_path / 'overfitted' / f)
    tr.load(early_stopped=False, device='cuda')


def test_main_encoder_student_config(tmp_path):
    """`main_encoder.py -t -c configs/encoder_student.py` end to end as a subprocess at the shipped widths (d = 512, 8 heads, 384
    tokens per window, codebook 32) with one layer per stack: one epoch of two batches + validation, checkpoints written."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, VQCPC_BATCH='4', VQCPC_NUM_BATCHES='2', VQCPC_EPOCHS='1', VQCPC_STUDENT_LAYERS='1')
    cfg_path = os.path.join(root, 'vqcpc-bach_b200', 'VQCPCB', 'configs', 'encoder_student.py')
    r = subprocess.run([sys.executable, os.path.join(root, 'main_encoder.py'), '-t', '-c', cfg_path], cwd=str(tmp_path), env=env,
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + '\n' + r.stderr[-3000:]
    assert '======= Epoch 0 =======' in r.stdout and 'Loss_teacher' in r.stdout
    models = os.listdir(tmp_path / 'models')
    assert len(models) == 1 and models[0].startswith('encoder_student_')
    for f in ('data_processor', 'downscaler', 'quantizer', 'decoder', 'teacher'):
        assert (tmp_path / 'models' / models[0] / 'overfitted' / f).exists()
