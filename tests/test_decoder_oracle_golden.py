"""CPU: decoder-forward oracle (oracle/decoder_oracle.py) against the goldens from the live reference, and the product's integer
index maps of the relative bias against the oracle's literal pad / view restatement."""
import os
import sys

import pytest
import torch

from oracle import decoder_oracle as D

GOLD = os.path.join(os.path.dirname(__file__), 'golden', 'decoder.pt')


@pytest.mark.parametrize('case', ['ac_ac', 'c_full'])
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
