"""CPU: decoder-forward oracle (oracle/decoder_oracle.py) against the goldens from the live reference, and the product's integer
index maps of the relative bias against the oracle's literal pad / view restatement."""
import os
import sys

import pytest
import torch

from oracle import decoder_oracle as D

GOLD = os.path.join(os.path.dirname(__file__), 'golden', 'decoder.pt')


@pytest.mark.parametrize('case', ['ac_ac', 'c_full', 'ac_diag', 'diag_noq', 'absolute'])
def test_decoder_oracle_matches_reference_golden(case):
    rec = torch.load(GOLD)[case]
    prm = D.make_decoder_params(rec['shapes'], seed=0)
    loss, logits = D.decoder_forward(rec['source'], rec['target'], prm, dict(rec['cfg'], total_upscaling=16))
    assert abs(loss.item() - rec['loss']) <= 1e-6 * abs(rec['loss'])
    for a, b in zip(logits, rec['logits']):
        assert (a - b).abs().max().item() <= 1e-5


@pytest.mark.parametrize('T,S', [(16, 16), (64, 4), (128, 8), (12, 3), (6, 6)])
def test_relative_index_maps_equal_literal_skewing(T, S):
    """VQCPCB/transformer/relative_maps.py (host logic, no GPU): gathering q.e^T through the maps equals the reference's
    pad / view / slice / mask sequence restated on floats in the oracle."""
    from VQCPCB.transformer.relative_maps import relative_index_maps
    H, hd = 2, 8
    g = torch.Generator().manual_seed(T * 100 + S)
    q = torch.randn(3 * H, T, hd, generator=g, dtype=torch.float64)
    e1 = torch.randn(H * S, hd, generator=g, dtype=torch.float64)
    e2 = torch.randn(H * S, hd, generator=g, dtype=torch.float64)
    ref = D.rel_bias_general(q, e1, e2, T, S)
    m1, m2 = relative_index_maps(T, S)
    out = torch.zeros_like(ref)
    for bh in range(3 * H):
        h = bh % H
        for mm, e in ((m1, e1), (m2, e2)):
            rel = (q[bh] @ e.view(H, S, hd)[h].t()).reshape(-1)
            out[bh] += torch.where(mm >= 0, rel[mm.clamp(min=0).long()], torch.where(mm == -2, -100.0, 0.0).double())
    assert (out - ref).abs().max().item() < 1e-12


def test_sampler_oracle_matches_reference_golden():
    """oracle sample_rows (top-k / top-p restatement + RandomState.choice) against tokens drawn by the reference's own
    top_k_top_p_filtering + np.random.choice on the same logits and seed."""
    import numpy as np
    for rec in torch.load(GOLD)['sampler']:
        rng = np.random.RandomState(rec['seed'])
        toks = D.sample_rows(rec['logits'], rec['temperature'], rec['top_k'], rec['top_p'], rng)
        assert toks == rec['tokens'].tolist()
        # numpy's choice consumes exactly one uniform per call: the stream the GPU sampler is fed with
        u = np.random.RandomState(rec['seed']).random_sample(len(toks))
        rng2 = np.random.RandomState(rec['seed'])
        for k in range(3):
            rng2.choice(np.arange(rec['V']), p=np.full(rec['V'], 1.0 / rec['V']))
        assert rng2.random_sample() == np.random.RandomState(rec['seed']).random_sample(4)[3] and u.shape[0] == len(toks)


@pytest.mark.parametrize('T,S', [(1024, 64), (1024, 1024), (96, 6), (64, 64)])
def test_relative_bias_rows_only_read_their_own_query(T, S):
    """The property the key / value cached generation rests on (asserted inside relative_index_maps as well)."""
    from VQCPCB.transformer.relative_maps import relative_index_maps
    m1, m2 = relative_index_maps(T, S)
    rows = torch.arange(T).view(T, 1)
    for m in (m1, m2):
        v = m >= 0
        assert bool((m[v] // S == rows.expand(T, S)[v]).all())


@pytest.mark.parametrize('case', ['ac_ac', 'c_full', 'ac_diag', 'diag_noq', 'absolute'])
def test_product_decoder_state_dict_is_the_references(case):
    """Host side only: the product ``Decoder`` exposes exactly the parameter names and shapes of the reference module (recorded
    by oracle/make_golden_decoder.py from the live reference), so its checkpoints load in either direction."""
    import types
    from VQCPCB.data_processor.bach_data_processor import BachDataProcessor
    from VQCPCB.decoders.decoder import Decoder, top_k_top_p_filtering
    rec = torch.load(GOLD)[case]
    c = rec['cfg']

    class NoQuantization:
        codebook_dim = 3

    enc = torch.nn.Module()
    enc.downscaler = types.SimpleNamespace(downscale_factors=[4, 4])
    enc.quantizer = NoQuantization() if c.get('no_quantization') else types.SimpleNamespace(codebook_size=c['codebook_size'],
                                                                                            num_codebooks=1, codebook_dim=3)
    C = len(c['vocab'])
    dec = Decoder(model_dir='/tmp/x', dataloader_generator=None, data_processor=BachDataProcessor(32, c['num_events'], list(c['vocab'])),
                  encoder=enc, transformer_type=c.get('transformer_type', 'relative'), encoder_attention_type=c['encoder_attention_type'],
                  cross_attention_type=c['cross_attention_type'], d_model=c['d_model'], num_encoder_layers=c['num_encoder_layers'],
                  num_decoder_layers=c['num_decoder_layers'], n_head=c['n_head'], dim_feedforward=c['dim_feedforward'],
                  positional_embedding_size=8, num_channels_encoder=1, num_events_encoder=c['num_events'] * C // 16,
                  num_channels_decoder=C, num_events_decoder=c['num_events'], dropout=0.0)
    mine = {k: tuple(v.shape) for k, v in dec.state_dict().items()}
    ref = {k: tuple(v) for k, v in rec['shapes'].items()}
    assert mine == ref, (sorted(set(mine) ^ set(ref))[:8])
    # the batched filter of the full-recompute path equals the reference's per-row function restated in the oracle
    g = torch.Generator().manual_seed(1)
    lg = torch.randn(6, 40, generator=g) * 3
    for top_k, top_p in [(0, 0.8), (7, 0.0), (5, 0.9)]:
        want = torch.stack([D.top_k_top_p_filtering(r, top_k=top_k, top_p=top_p) for r in lg])
        assert torch.equal(top_k_top_p_filtering(lg, top_k=top_k, top_p=top_p), want)


def test_get_decoder_types_follow_the_reference_table():
    """getters.py:290-393 of the reference: decoder_type -> (transformer type, encoder attention, cross attention)."""
    from VQCPCB.getters import _DECODER_TYPES
    assert _DECODER_TYPES == {'transformer': ('anticausal', 'full'), 'transformer_relative': ('anticausal', 'anticausal'),
                              'transformer_relative_fullCross': ('anticausal', 'full'),
                              'transformer_relative_diagonal': ('anticausal', 'diagonal'),
                              'transformer_relative_full': ('full', 'full')}


def test_decoder_checkpoint_roundtrip(tmp_path):
    """Decoder.save / load (decoder.py:274-292): `<model_dir>/{early_stopped,overfitted}/decoder` holds the state_dict."""
    import types
    from VQCPCB.data_processor.bach_data_processor import BachDataProcessor
    from VQCPCB.decoders.decoder import Decoder
    c = torch.load(GOLD)['ac_diag']['cfg']
    enc = torch.nn.Module()
    enc.downscaler = types.SimpleNamespace(downscale_factors=[4, 4])
    enc.quantizer = types.SimpleNamespace(codebook_size=c['codebook_size'], num_codebooks=1, codebook_dim=3)

    def make():
        return Decoder(model_dir=str(tmp_path), dataloader_generator=None, data_processor=BachDataProcessor(32, c['num_events'], list(c['vocab'])),
                       encoder=enc, transformer_type='relative', encoder_attention_type='anticausal', cross_attention_type='diagonal',
                       d_model=c['d_model'], num_encoder_layers=2, num_decoder_layers=2, n_head=c['n_head'], dim_feedforward=256,
                       positional_embedding_size=8, num_channels_encoder=1, num_events_encoder=4, num_channels_decoder=4,
                       num_events_decoder=c['num_events'], dropout=0.0)
    a, b = make(), make()
    a.save(early_stopped=True)
    a.save(early_stopped=False)
    assert os.path.exists(tmp_path / 'early_stopped' / 'decoder') and os.path.exists(tmp_path / 'overfitted' / 'decoder')
    assert any(not torch.equal(x, y) for x, y in zip(a.state_dict().values(), b.state_dict().values()))
    b.load(early_stopped=True, device='cpu')
    assert all(torch.equal(x, y) for x, y in zip(a.state_dict().values(), b.state_dict().values()))
    assert repr(a) == 'Decoder-relative-AC-D'
