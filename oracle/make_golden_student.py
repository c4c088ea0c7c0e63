"""Golden vectors of the student / distillation trainer from the LIVE reference (authoring container only).

    python oracle/make_golden_student.py       # writes tests/golden/student.pt

Builds the unmodified reference ``StudentEncoderTrainer`` (encoder with the 'bach' data processor and the linear transformer
downscaler, ``TeacherRelative``, ``AuxiliaryDecoderRelative``) on CPU with dropout 0, STRICT-loads the deterministic parameters
of ``oracle/student_oracle.make_student_params`` (pins key names and shapes), runs ``epoch(train=True)`` on one batch and stores
the monitored quantities, the masked event index it drew, fingerprints of the clipped gradients and of the updated parameters,
the code indices and the teacher's logits at the masked event."""
import os
import sys
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.dont_write_bytecode = True

from oracle import student_oracle as S            # noqa: E402
from oracle.make_golden import summarize          # noqa: E402

CASES = {
    # 16 events x 4 voices = 64 tokens: teacher attention over L = 64, decoder over L = 4 then 16, 4 codes per window
    'small': dict(cfg=dict(d_model=128, n_head=2, dim_feedforward=256, list_of_num_layers=[1, 1], num_events=16, teacher_layers=2,
                           dec_layers=[1, 1], vocab=[20, 21, 22, 23], num_events_masked=1), B=3, seed=11),
    # 32 events = 128 tokens (decoder L = 8 then 32), two layers in some stacks, 2 events masked on either side
    'mid': dict(cfg=dict(d_model=128, n_head=2, dim_feedforward=256, list_of_num_layers=[2, 1], num_events=32, teacher_layers=1,
                         dec_layers=[1, 2], vocab=[9, 30, 17, 12], num_events_masked=2), B=2, seed=5),
}


def main():
    sys.path.insert(0, '/root/reference')
    for n in ['matplotlib', 'matplotlib.pyplot', 'seaborn', 'music21']:
        sys.modules.setdefault(n, types.ModuleType(n))
    sys.modules['matplotlib'].pyplot = sys.modules['matplotlib.pyplot']
    from VQCPCB.encoder import Encoder
    from VQCPCB.student_encoder_trainer import StudentEncoderTrainer
    from VQCPCB.teachers.teacher_relative import TeacherRelative
    from VQCPCB.auxiliary_decoders.auxiliary_decoder_relative import AuxiliaryDecoderRelative
    from VQCPCB.data_processor.bach_data_processor import BachDataProcessor
    from VQCPCB.downscalers.relative_transformer_downscaler_linear import RelativeTransformerDownscalerLinear
    from VQCPCB.quantizer.vector_quantizer import ProductVectorQuantizer

    out = {}
    for name, spec in CASES.items():
        cfg = S.student_cfg(**spec['cfg'])
        prm = S.make_student_params(cfg, seed=0)
        batch = S.make_student_batch(cfg, spec['B'], seed=1234)
        V = cfg.num_voices
        dp = BachDataProcessor(cfg.embedding_size, cfg.num_events, list(cfg.vocab))
        down = RelativeTransformerDownscalerLinear(cfg.embedding_size, cfg.codebook_dim, V, list(cfg.downscale_factors), cfg.d_model,
                                                   cfg.n_head, list(cfg.list_of_num_layers), cfg.dim_feedforward, 0.0)
        vq = ProductVectorQuantizer(cfg.codebook_size, cfg.codebook_dim, cfg.commitment_cost, cfg.num_codebooks, False, False,
                                    cfg.squared_l2_norm)
        enc = Encoder('/tmp/vqcpc_golden_student', dp, down, vq, None)
        tdp = BachDataProcessor(cfg.embedding_size, cfg.num_events, list(cfg.vocab))
        teacher = TeacherRelative(data_processor=tdp, num_layers=cfg.teacher_layers, num_tokens_per_channel=list(cfg.vocab),
                                  positional_embedding_size=cfg.teacher_pos_emb, d_model=cfg.d_model,
                                  dim_feedforward=cfg.dim_feedforward, n_head=cfg.n_head, num_tokens=cfg.num_events * V, dropout=0.0)
        nb = cfg.num_events * V // 16
        dec = AuxiliaryDecoderRelative(num_tokens_per_channel=list(cfg.vocab), codebook_dim=cfg.codebook_dim,
                                       upscale_factors=list(reversed(cfg.downscale_factors)), list_of_num_layers=list(cfg.dec_layers),
                                       n_head=cfg.n_head, d_model=cfg.d_model, dim_feedforward=cfg.dim_feedforward,
                                       num_tokens_bottleneck=nb, dropout=0.0)
        groups = {'data_processor': dp, 'downscaler': down, 'quantizer': vq, 'teacher': teacher, 'decoder': dec}
        for g, mod in groups.items():
            sd = {k[len(g) + 1:]: v for k, v in prm.items() if k.startswith(g + '.')}
            mod.load_state_dict(sd, strict=True)
        tr = StudentEncoderTrainer('/tmp/vqcpc_golden_student', types.SimpleNamespace(), enc, cfg.num_events_masked, teacher, dec,
                                   cfg.quantization_weighting)
        lr = 1e-4
        tr.init_optimizers(lr, False)
        torch.manual_seed(spec['seed'])
        event = torch.randint(high=cfg.num_events, size=()).item()
        torch.manual_seed(spec['seed'])
        means = tr.epoch([batch], train=True, num_batches=1)
        rec = dict(cfg=dict(cfg), spec=dict(spec), lr=lr, masked_event_index=event, means={k: float(v) for k, v in means.items()})
        named = {}
        for g, mod in groups.items():
            for k, p in mod.named_parameters():
                named[f'{g}.{k}'] = p
        assert set(named) == set(prm), (set(named) ^ set(prm))
        rec['new_params'] = {k: summarize(p, k) for k, p in named.items()}
        rec['delta'] = {k: summarize(p.detach() - prm[k], k) for k, p in named.items()}
        rec['clipped_grads'] = {k: summarize(p.grad, k) for k, p in named.items() if p.grad is not None}
        # forward-only quantities from the ORIGINAL parameters (fresh modules are cheaper than un-stepping)
        for g, mod in groups.items():
            mod.load_state_dict({k[len(g) + 1:]: v for k, v in prm.items() if k.startswith(g + '.')}, strict=True)
        with torch.no_grad():
            x = batch['x'].long()
            masked, pred = tr.mask_teacher(x, cfg.num_events_masked) if False else (None, None)
            zq, idx, ql = enc(x)
            rec['indices'] = idx.clone()
            rec['zq'] = zq.clone()
            rec['qloss'] = ql.clone()
            dl = dec(zq)
            rec['decoder_logits_event'] = [w[:, event].clone() for w in dl]
        out[name] = rec
        print(name, 'event', event, rec['means'])
    path = os.path.join(ROOT, 'tests', 'golden', 'student.pt')
    torch.save(out, path)
    print('wrote', path, os.path.getsize(path))


if __name__ == '__main__':
    main()
