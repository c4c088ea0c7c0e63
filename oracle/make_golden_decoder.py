"""Golden vectors of the decoder's forward pass from the LIVE reference (authoring container only).

    python oracle/make_golden_decoder.py       # writes tests/golden/decoder.pt

Builds the unmodified reference ``Decoder`` (transformer_type='relative') on CPU with dropout 0 around a stub encoder (``forward``
only consumes code indices), loads deterministic parameters for every key of its own state_dict (shapes recorded in the golden:
the product must accept exactly these keys), runs ``forward(source, target)`` and stores loss + logits."""
import os
import sys
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.dont_write_bytecode = True

from oracle import decoder_oracle as D            # noqa: E402

CASES = {
    # 16 events x 4 voices = 64 target tokens, 4 codes: cross-attention with T / S = 16
    'ac_ac': dict(cfg=dict(d_model=128, n_head=2, dim_feedforward=256, num_encoder_layers=1, num_decoder_layers=2, num_events=16,
                           vocab=[20, 21, 22, 23], codebook_size=16, encoder_attention_type='anticausal',
                           cross_attention_type='anticausal'), B=2),
    # 32 events = 128 tokens, 8 codes, full cross-attention, causal code encoder
    'c_full': dict(cfg=dict(d_model=128, n_head=2, dim_feedforward=256, num_encoder_layers=2, num_decoder_layers=1, num_events=32,
                            vocab=[9, 30, 17, 12], codebook_size=32, encoder_attention_type='causal', cross_attention_type='full'),
                   B=2),
    # the 'diagonal' aligned decoder layer (decoder_relative_AC_D_C_*: 6 of the 10 shipped relative configurations)
    'ac_diag': dict(cfg=dict(d_model=128, n_head=2, dim_feedforward=256, num_encoder_layers=2, num_decoder_layers=2, num_events=16,
                             vocab=[20, 21, 22, 23], codebook_size=16, encoder_attention_type='anticausal',
                             cross_attention_type='diagonal'), B=3),
    # encoder without quantization: the decoder consumes the z vectors through a Linear (decoder_relative_AC_D_C_*_noQuantization)
    'diag_noq': dict(cfg=dict(d_model=128, n_head=2, dim_feedforward=256, num_encoder_layers=1, num_decoder_layers=2, num_events=16,
                              vocab=[20, 21, 22, 23], codebook_size=16, encoder_attention_type='anticausal',
                              cross_attention_type='diagonal', no_quantization=True), B=2),
    # decoder_type 'transformer' (decoder_random.py / decoder_sameSeq.py): absolute positional embeddings, no relative bias
    'absolute': dict(cfg=dict(d_model=128, n_head=2, dim_feedforward=256, num_encoder_layers=1, num_decoder_layers=2, num_events=16,
                              vocab=[20, 21, 22, 23], codebook_size=16, encoder_attention_type='anticausal',
                              cross_attention_type='full', transformer_type='absolute'), B=2),
}


class NoQuantization:
    codebook_dim = 3


class StubEncoder(torch.nn.Module):
    def __init__(self, codebook_size):
        super().__init__()
        self.downscaler = types.SimpleNamespace(downscale_factors=[4, 4])
        self.quantizer = types.SimpleNamespace(codebook_size=codebook_size, num_codebooks=1, codebook_dim=3)


def main():
    sys.path.insert(0, '/root/reference')
    for n in ['matplotlib', 'matplotlib.pyplot', 'seaborn', 'music21']:
        sys.modules.setdefault(n, types.ModuleType(n))
    sys.modules['matplotlib'].pyplot = sys.modules['matplotlib.pyplot']
    sys.modules['music21'].corpus = types.ModuleType('music21.corpus')
    try:
        from VQCPCB.decoders.decoder import Decoder
    except Exception:
        # decoder.py imports data / plotting helpers that need music21 symbols: stub what it asks for
        import importlib
        for n in ['VQCPCB.datasets.chorale_dataset', 'VQCPCB.datasets.helpers', 'DatasetManager.chorale_dataset', 'DatasetManager.helpers']:
            m = types.ModuleType(n)
            m.START_SYMBOL, m.END_SYMBOL, m.PAD_SYMBOL = 'START', 'END', 'XX'
            sys.modules[n] = m
        for k in [k for k in sys.modules if k.startswith('VQCPCB.decoders')]:
            del sys.modules[k]
        from VQCPCB.decoders.decoder import Decoder
    from VQCPCB.data_processor.bach_data_processor import BachDataProcessor
    out = {}
    for name, spec in CASES.items():
        c = spec['cfg']
        C = len(c['vocab'])
        dp = BachDataProcessor(32, c['num_events'], list(c['vocab']))
        enc = StubEncoder(c['codebook_size'])
        if c.get('no_quantization'):
            enc.quantizer = NoQuantization()
        S = c['num_events'] * C // 16
        dec = Decoder(model_dir='/tmp/vqcpc_golden_decoder', dataloader_generator=None, data_processor=dp, encoder=enc,
                      transformer_type=c.get('transformer_type', 'relative'), encoder_attention_type=c['encoder_attention_type'],
                      cross_attention_type=c['cross_attention_type'], d_model=c['d_model'],
                      num_encoder_layers=c['num_encoder_layers'], num_decoder_layers=c['num_decoder_layers'], n_head=c['n_head'],
                      dim_feedforward=c['dim_feedforward'], positional_embedding_size=8, num_channels_encoder=1,
                      num_events_encoder=S, num_channels_decoder=C, num_events_decoder=c['num_events'], dropout=0.0)
        dec.eval()
        shapes = {k: tuple(v.shape) for k, v in dec.state_dict().items()}
        prm = D.make_decoder_params(shapes, seed=0)
        dec.load_state_dict(prm, strict=True)
        g = torch.Generator().manual_seed(77)
        source = torch.randint(0, c['codebook_size'], (spec['B'], S), generator=g)
        if c.get('no_quantization'):
            source = torch.randn(spec['B'], S, 3, generator=g)
        target = torch.stack([torch.randint(0, v, (spec['B'], c['num_events']), generator=g) for v in c['vocab']], dim=2)
        with torch.no_grad():
            fp = dec.forward(source, target)
        ocfg = dict(c, total_upscaling=16)
        loss_o, logits_o = D.decoder_forward(source, target, prm, ocfg)
        err = max((a - b).abs().max().item() for a, b in zip(fp['weights_per_category'], logits_o))
        print(name, 'loss ref', fp['loss'].item(), 'oracle', loss_o.item(), 'max logit err', err)
        out[name] = dict(cfg=c, B=spec['B'], shapes=shapes, source=source, target=target, loss=fp['loss'].item(),
                         logits=[w.clone() for w in fp['weights_per_category']])
    # the sampling step of generate (decoder.py:619-644): the reference's own top_k_top_p_filtering + np.random.choice
    import numpy as np
    from VQCPCB.utils import top_k_top_p_filtering
    g = torch.Generator().manual_seed(5)
    samp = []
    for V, temperature, top_k, top_p in [(23, 1.0, 0, 1.0), (64, 0.9, 0, 0.8), (130, 1.2, 10, 0.0), (57, 1.0, 5, 0.95), (9, 0.7, 0, 0.3)]:
        logits = torch.randn(48, V, generator=g) * 2.5
        logits[::7, 3] = logits[::7, 5]                              # some exact ties
        np.random.seed(1000 + V)
        toks = []
        for row in logits / temperature:
            fl = top_k_top_p_filtering(row.clone(), top_k=top_k, top_p=top_p)
            p = torch.softmax(fl, dim=-1).numpy()
            toks.append(int(np.random.choice(np.arange(V), p=p)))
        samp.append(dict(V=V, temperature=temperature, top_k=top_k, top_p=top_p, logits=logits, seed=1000 + V,
                         tokens=torch.tensor(toks)))
        print('sampler', V, temperature, top_k, top_p, toks[:12])
    out['sampler'] = samp
    path = os.path.join(ROOT, 'tests', 'golden', 'decoder.pt')
    torch.save(out, path)
    print('wrote', path, os.path.getsize(path))


if __name__ == '__main__':
    main()
