"""Generate golden vectors from the LIVE reference (run in the authoring container only).

    python oracle/make_golden.py            # writes tests/golden/*.pt

Imports the unmodified reference from /root/reference (read-only; matplotlib / seaborn / music21 are stubbed as
empty modules because they are import-time-only on this path), loads the deterministic parameter
dict of oracle/params.py with a STRICT load_state_dict (pins key names and shapes), runs
``Encoder.forward`` and one ``VQCPCEncoderTrainer.epoch(train=True)`` step on CPU with dropout 0 and stores
inputs' seeds + outputs.  /root/reference does not exist on the GPU box, so only these files travel.
"""
import os
import sys
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.dont_write_bytecode = True

from oracle import params as P            # noqa: E402
from oracle import vqcpc_oracle as O      # noqa: E402


def import_reference():
    sys.path.insert(0, '/root/reference')
    for n in ['matplotlib', 'matplotlib.pyplot', 'seaborn', 'music21']:
        sys.modules.setdefault(n, types.ModuleType(n))
    sys.modules['matplotlib'].pyplot = sys.modules['matplotlib.pyplot']
    from VQCPCB.encoder import Encoder
    from VQCPCB.vqcpc_encoder_trainer import VQCPCEncoderTrainer
    from VQCPCB.data_processor.bach_cpc_data_processor import BachCPCDataProcessor
    from VQCPCB.downscalers.relative_transformer_downscaler import RelativeTransformerDownscaler
    from VQCPCB.downscalers.lstm_downscaler import LstmDownscaler
    from VQCPCB.downscalers.relative_transformer_downscaler_linear import RelativeTransformerDownscalerLinear
    from VQCPCB.quantizer.vector_quantizer import ProductVectorQuantizer
    from VQCPCB.upscalers.mlp_upscaler import MlpUpscaler
    return dict(Encoder=Encoder, Trainer=VQCPCEncoderTrainer, DP=BachCPCDataProcessor,
                Down=RelativeTransformerDownscaler, LstmDown=LstmDownscaler, DownLin=RelativeTransformerDownscalerLinear, VQ=ProductVectorQuantizer, Up=MlpUpscaler)


def build_reference(R, cfg, prm, model_dir='/tmp/vqcpc_golden'):
    dp = R['DP'](cfg.embedding_size, cfg.ticks_per_block, cfg.num_voices, list(cfg.vocab),
                 cfg.num_voices * cfg.ticks_per_block)
    if cfg.get('downscaler', 'transformer') == 'lstm':
        down = R['LstmDown'](cfg.embedding_size, cfg.codebook_dim, cfg.num_voices, list(cfg.downscale_factors),
                             cfg.lstm_hidden, cfg.lstm_layers, 0.0, cfg.lstm_bidirectional)
    elif cfg.get('downscaler', 'transformer') == 'transformer_linear':
        down = R['DownLin'](cfg.embedding_size, cfg.codebook_dim, cfg.num_voices, list(cfg.downscale_factors),
                            cfg.d_model, cfg.n_head, list(cfg.list_of_num_layers), cfg.dim_feedforward, 0.0)
    else:
        down = R['Down'](cfg.embedding_size, cfg.codebook_dim, cfg.num_voices, list(cfg.downscale_factors),
                         cfg.d_model, cfg.n_head, list(cfg.list_of_num_layers), cfg.dim_feedforward, 0.0)
    vq = R['VQ'](cfg.codebook_size, cfg.codebook_dim, cfg.commitment_cost, cfg.num_codebooks, False, False,
                 cfg.squared_l2_norm)
    up = R['Up'](cfg.codebook_dim, cfg.z_dim, cfg.up_hidden, 0.0) if cfg.upscaler else None
    enc = R['Encoder'](model_dir, dp, down, vq, up)
    gen = types.SimpleNamespace(num_blocks_right=cfg.KR)
    tr = R['Trainer'](model_dir, gen, enc,
                      dict(output_dim=cfg.c_dim, hidden_size=cfg.c_hidden, num_layers=cfg.c_layers, dropout=0.0,
                           bidirectional=bool(cfg.get('bidirectional', False))), cfg.quantization_weighting)
    files = P.split_by_file(prm)
    dp.load_state_dict(files['data_processor'], strict=True)
    down.load_state_dict(files['downscaler'], strict=True)
    vq.load_state_dict(files['quantizer'], strict=True)
    if up is not None:
        up.load_state_dict(files['upscaler'], strict=True)
    tr.c_module.load_state_dict(files['c_module'], strict=True)
    tr.fks_module.load_state_dict(files['fks_module'], strict=True)
    if cfg.get('bidirectional', False):
        tr.c_module_back.load_state_dict(files['c_module_back'], strict=True)
        tr.fks_module_back.load_state_dict(files['fks_module_back'], strict=True)
    return enc, tr


def ref_named_params(enc, tr):
    out = {}
    for f, mod in [('data_processor', enc.data_processor), ('downscaler', enc.downscaler),
                   ('quantizer', enc.quantizer), ('upscaler', enc.upscaler),
                   ('c_module', tr.c_module), ('fks_module', tr.fks_module),
                   ('c_module_back', tr.c_module_back), ('fks_module_back', getattr(tr, 'fks_module_back', None))]:
        if mod is None:
            continue
        for k, p in mod.named_parameters():
            out[f'{f}.{k}'] = p
    return out


def summarize(t, name, seed=7):
    """Small, order-sensitive fingerprint of a tensor: norm, projection on a fixed random direction, head."""
    t = t.detach().float().reshape(-1)
    g = torch.Generator()
    g.manual_seed(seed)
    r = torch.randn(t.numel(), generator=g)
    return dict(norm=t.norm().item(), proj=(t.double() @ r.double()).item(), head=t[:8].clone(), numel=t.numel())


CASES = {
    # BASELINE.json configs[0]: B=16, seq 64, K=16
    'cfg1': dict(cfg=dict(codebook_size=16, KL=2, KR=2), B=16),
    # configs[1] shapes (K=512, seq 256) at a CPU-sized batch
    'cfg2_small': dict(cfg=dict(codebook_size=512, KL=8, KR=8), B=4),
    # product quantizer, D=32, no upscaler (encoder_sameSeq_noUpscale_config.py style), non-squared loss
    # (num_codebooks must be 1 here: the reference trainer's index .view at vqcpc_encoder_trainer.py:207 raises otherwise)
    'noup_d32': dict(cfg=dict(codebook_size=16, codebook_dim=32, num_codebooks=1, upscaler=False, KL=2, KR=2,
                              squared_l2_norm=False), B=4),
    # bidirectional c-net (encoder_random_16C.py style) incl. the reference's backward-negative indexing (quirk Q1)
    'cfg1_bidir': dict(cfg=dict(codebook_size=16, KL=2, KR=2, bidirectional=True), B=6),
    # GRU block downscaler (encoder_random_config.py shape: bidirectional 2-layer GRU pair, K = 32, 6 + 6 blocks)
    'lstm': dict(cfg=dict(downscaler='lstm', downscale_factors=[16], lstm_hidden=512, lstm_layers=2,
                          lstm_bidirectional=True, codebook_size=32, codebook_scale=0.05, KL=6, KR=6), B=4),
    # aggregation by Linear(factor * d -> d) instead of subsampling (encoder_random_transfolin_config.py shape)
    'translin': dict(cfg=dict(downscaler='transformer_linear', codebook_size=16, KL=2, KR=2), B=4),
    'lstm_uni': dict(cfg=dict(downscaler='lstm', downscale_factors=[16], lstm_hidden=96, lstm_layers=1,
                              lstm_bidirectional=False, codebook_size=16, codebook_scale=0.05, KL=2, KR=2), B=3),
}


def run_case(R, name, spec):
    torch.manual_seed(0)
    cfg = O.default_cfg(**spec['cfg'])
    B = spec['B']
    prm = P.make_params(cfg, seed=0)
    batch = P.make_batch(cfg, B, seed=1234)
    enc, tr = build_reference(R, cfg, prm)
    enc.eval()
    tr.eval()
    gold = dict(cfg=dict(cfg), B=B, param_seed=0, batch_seed=1234, torch=torch.__version__)

    # ---- Encoder.forward on x_left (encoder.py:76-95) + pre-quantisation z ------------------
    with torch.no_grad():
        x = batch['x_left']
        tok = enc.data_processor.preprocess(x)
        emb = enc.data_processor.embed(tok)
        from VQCPCB.utils import flatten
        z = enc.downscaler(flatten(emb))
        zq, idx, ql = enc(x)
    gold['enc'] = dict(tokens=tok.to(torch.int16), z=z, zq=zq, idx=idx.to(torch.int16), qloss=ql)

    # ---- one training step (vqcpc_encoder_trainer.py:169-354) ----------------------------------
    lr = 1e-4
    tr.init_optimizers(lr, False)
    before = {k: p.detach().clone() for k, p in ref_named_params(enc, tr).items()}
    means = tr.epoch([batch], True, 1, False)
    named = ref_named_params(enc, tr)
    gold['step'] = dict(
        means={k: (float(v) if not isinstance(v, list) else [float(a) for a in v]) for k, v in means.items()},
        clipped_grads={k: summarize(p.grad, k) for k, p in named.items() if p.grad is not None},
        delta={k: summarize(p.detach() - before[k], k) for k, p in named.items()},
        lr=lr,
    )

    # ---- cross-check the restatement right here (the CPU test re-checks from the file) -------
    out, grads, gnorm, newp, m, v = O.trainer_step(batch, prm, cfg, lr=lr)
    zq_o, idx_o, ql_o = O.encoder_forward(batch['x_left'], prm, cfg)
    print(f'[{name}] loss ref {means["loss"]:.6f} oracle {out["loss"].item():.6f}  '
          f'idx equal {bool((idx_o == idx).all())}  zq maxdiff {(zq_o - zq).abs().max().item():.2e}  '
          f'gnorm {gnorm.item():.4f}')
    coef = min(1.0, 5.0 / (gnorm.item() + 1e-6))
    worst = 0.0
    for k, p in named.items():
        if p.grad is None:
            continue
        d = (grads[k] * coef - p.grad).norm().item() / (p.grad.norm().item() + 1e-12)
        if d > 2e-5:
            print(f'    {k}: rel diff {d:.2e} (grad norm {p.grad.norm().item():.3e})')
        worst = max(worst, d)
    print(f'[{name}] worst rel grad diff oracle vs reference: {worst:.2e}')
    gold['gnorm'] = gnorm.item()
    return gold


def quantizer_cases(R):
    """ProductVectorQuantizer.forward alone (vector_quantizer.py:85-159) incl. num_codebooks > 1 and exact ties
    (codebook rows equal to data rows, as after the data-dependent init :57-70)."""
    out = {}
    for name, (Kc, D, ncb, sq) in {'k16_d3': (16, 3, 1, True), 'k512_d3': (512, 3, 1, True),
                                   'k32_d32_ncb2': (32, 32, 2, True), 'k64_d8_ncb4_l2': (64, 8, 4, False)}.items():
        g = torch.Generator()
        g.manual_seed(99)
        x = torch.randn(3, 200, D, generator=g) * 2
        books = [torch.randn(Kc, D // ncb, generator=g) * 2 for _ in range(ncb)]
        for b in books:                               # exact ties / zero distances / duplicate rows
            b[1] = x[0, 5, :D // ncb]
            b[3] = b[2]
        vq = R['VQ'](Kc, D, 0.25, ncb, False, False, sq)
        vq.load_state_dict({f'embeddings.{i}': b for i, b in enumerate(books)}, strict=True)
        vq.eval()
        xr = x.clone().requires_grad_(True)
        zsg, idx, loss = vq(xr)
        gl = torch.randn(loss.shape, generator=g)
        gz = torch.randn(zsg.shape, generator=g)
        ((loss * gl).sum() + (zsg * gz).sum()).backward()
        out[name] = dict(Kc=Kc, D=D, ncb=ncb, squared=sq, x=x, books=books, zsg=zsg.detach(), idx=idx.to(torch.int16),
                         loss=loss.detach(), gl=gl, gz=gz, dx=xr.grad.clone(),
                         dbooks=[p.grad.clone() for p in vq.embeddings])
        zo, io, lo = O.quantize(x, books, 0.25, sq)
        print(f'[vq {name}] idx equal {bool((io == idx).all())} loss maxdiff {(lo - loss).abs().max().item():.1e}')
    return out


def main():
    R = import_reference()
    os.makedirs(os.path.join(ROOT, 'tests', 'golden'), exist_ok=True)
    only = sys.argv[1:]                      # optional: regenerate just the named cases
    if not only:
        torch.save(quantizer_cases(R), os.path.join(ROOT, 'tests', 'golden', 'quantizer.pt'))
    for name, spec in CASES.items():
        if only and name not in only:
            continue
        gold = run_case(R, name, spec)
        path = os.path.join(ROOT, 'tests', 'golden', f'{name}.pt')
        torch.save(gold, path)
        print('wrote', path, os.path.getsize(path), 'bytes')


if __name__ == '__main__':
    main()
