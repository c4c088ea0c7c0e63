"""GPU: inference path of the decoder against the reference goldens / CPU oracle."""
import os
import types

import pytest
import torch

from oracle import decoder_oracle as D
from tests.helpers import rel_err

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), 'golden', 'decoder.pt')


class NoQuantization:
    codebook_dim = 3


class _StubEncoder(torch.nn.Module):
    def __init__(self, codebook_size, no_quantization=False):
        super().__init__()
        self.downscaler = types.SimpleNamespace(downscale_factors=[4, 4])
        self.quantizer = NoQuantization() if no_quantization else types.SimpleNamespace(codebook_size=codebook_size, num_codebooks=1,
                                                                                        codebook_dim=3)


def _build(c):
    from VQCPCB.data_processor.bach_data_processor import BachDataProcessor
    from VQCPCB.decoders.decoder import Decoder
    C = len(c['vocab'])
    dp = BachDataProcessor(32, c['num_events'], list(c['vocab']))
    S = c['num_events'] * C // 16
    dec = Decoder(model_dir='/tmp/vqcpc_b200_decoder', dataloader_generator=None, data_processor=dp,
                  encoder=_StubEncoder(c['codebook_size'], c.get('no_quantization', False)),
                  transformer_type=c.get('transformer_type', 'relative'), encoder_attention_type=c['encoder_attention_type'],
                  cross_attention_type=c['cross_attention_type'], d_model=c['d_model'], num_encoder_layers=c['num_encoder_layers'],
                  num_decoder_layers=c['num_decoder_layers'], n_head=c['n_head'], dim_feedforward=c['dim_feedforward'],
                  positional_embedding_size=8, num_channels_encoder=1, num_events_encoder=S, num_channels_decoder=C,
                  num_events_decoder=c['num_events'], dropout=0.0)
    return dec.to('cuda').eval()


@pytest.mark.parametrize('Lq,Lk,masked', [(64, 4, True), (48, 48, True), (100, 100, False), (128, 8, False)])
def test_attn_map_fwd_vs_oracle(Lq, Lk, masked):
    """vqcpc_attn_map_fwd (index maps + additive mask, q and k / v from different matrices) against the fp64 oracle attention."""
    import ctypes
    from VQCPCB import _native, ops
    from VQCPCB.transformer.relative_maps import device_maps
    H, hd, P = 2, 64, 3
    d = H * hd
    g = torch.Generator().manual_seed(Lq + Lk)
    q = torch.randn(P, Lq, d, generator=g) * 0.4
    kv = torch.randn(P, Lk, 2 * d, generator=g) * 0.5
    e1 = torch.randn(H * Lk, hd, generator=g) * 0.3
    e2 = torch.randn(H * Lk, hd, generator=g) * 0.3
    mask = None
    if masked:
        mask = torch.repeat_interleave(D.causal_mask(Lk).t(), Lq // Lk, dim=0).contiguous()
    q64, k64, v64 = q.double(), kv[:, :, :d].double(), kv[:, :, d:].double()
    qh = q64.reshape(P, Lq, H, hd).transpose(1, 2).reshape(P * H, Lq, hd)
    kh = k64.reshape(P, Lk, H, hd).transpose(1, 2).reshape(P * H, Lk, hd)
    vh = v64.reshape(P, Lk, H, hd).transpose(1, 2).reshape(P * H, Lk, hd)
    s = torch.bmm(qh, kh.transpose(1, 2)) + D.rel_bias_general(qh, e1.double(), e2.double(), Lq, Lk)
    if mask is not None:
        s = s + mask.double().unsqueeze(0)
    a = torch.softmax(s, dim=-1)
    ref = torch.bmm(a, vh).reshape(P, H, Lq, hd).transpose(1, 2).reshape(P * Lq, d)
    qc, kvc = q.reshape(P * Lq, d).cuda(), kv.reshape(P * Lk, 2 * d).cuda()
    m1, m2 = device_maps(Lq, Lk, 'cuda')
    out = torch.empty(P * Lq, d, device='cuda')
    probs = torch.empty(P, H, Lq, Lk, device='cuda')
    e1c, e2c = e1.cuda(), e2.cuda()
    mc = mask.cuda() if mask is not None else None
    pp = lambda t: None if t is None else ctypes.c_void_p(t.data_ptr())      # noqa: E731
    _native.call('vqcpc_attn_map_fwd', pp(qc), d, pp(kvc), pp(kvc[:, d:]), 2 * d, pp(e1c), pp(e2c), pp(m1), pp(m2), pp(mc), pp(out),
                 pp(probs), P, Lq, Lk, H, hd, ops._stream())
    torch.cuda.synchronize()
    assert rel_err(out, ref) < 3e-6
    assert rel_err(probs, a.reshape(P, H, Lq, Lk)) < 3e-6


@pytest.mark.parametrize('mode', ['fp32', 'bf16x3'])
@pytest.mark.parametrize('case', ['ac_ac', 'c_full', 'ac_diag', 'diag_noq', 'absolute'])
def test_decoder_forward_vs_reference_golden(case, mode):
    """Decoder.forward on the CUDA path: the reference's own state_dict keys load strictly; logits and loss against the golden."""
    from VQCPCB import ops
    rec = torch.load(GOLD)[case]
    prm = D.make_decoder_params(rec['shapes'], seed=0)
    ops.set_gemm_mode(mode)
    try:
        dec = _build(rec['cfg'])
        dec.load_state_dict({k: v.cuda() for k, v in prm.items()}, strict=True)
        ops.bump_weight_version()
        fp = dec.forward(rec['source'], rec['target'])
        tol = 2e-5 if mode == 'fp32' else 1e-4
        for a, b in zip(fp['weights_per_category'], rec['logits']):
            assert rel_err(a, b) < tol
        assert abs(fp['loss'].item() - rec['loss']) <= 1e-4 * abs(rec['loss'])
    finally:
        ops.set_gemm_mode('bf16x3')


def test_decoder_greedy_generation_is_consistent_with_forward():
    """generate_from_codes recomputes the forward per token like the reference (decoder.py:607-644).  With greedy decoding the
    result must be a fixed point: a teacher-forced forward over the generated window predicts every generated token."""
    rec = torch.load(GOLD)['ac_ac']
    prm = D.make_decoder_params(rec['shapes'], seed=0)
    dec = _build(rec['cfg'])
    dec.load_state_dict({k: v.cuda() for k, v in prm.items()}, strict=True)
    x = dec.generate_from_codes(rec['source'].cuda(), greedy=True)
    assert x.shape == (rec['B'], rec['cfg']['num_events'], 4)
    logits = dec.forward(rec['source'].cuda(), x)['weights_per_category']
    for c, w in enumerate(logits):
        assert torch.equal(w.argmax(dim=-1), x[:, :, c])
    # sampling path: top-p / top-k filtered tokens stay inside the vocabulary and differ between two seeds
    g = torch.Generator(device='cuda').manual_seed(1)
    xs = dec.generate_from_codes(rec['source'].cuda(), temperature=1.0, top_p=0.9, top_k=8, generator=g)
    for c, v in enumerate(rec['cfg']['vocab']):
        assert int(xs[:, :, c].max()) < v and int(xs[:, :, c].min()) >= 0


def _pp(t):
    import ctypes
    return None if t is None else ctypes.c_void_p(t.data_ptr())


@pytest.mark.parametrize('M,N,K', [(1, 512, 512), (3, 1536, 512), (8, 512, 1024), (16, 23, 48), (5, 130, 512), (20, 512, 512),
                                   (32, 512, 1024), (27, 64, 48), (32, 96, 2048)])
def test_linear_small_vs_fp64(M, N, K):
    """vqcpc_linear_small (bias, q-third scaling, ReLU, residual, start-row substitution) against fp64."""
    from VQCPCB import _native, ops
    g = torch.Generator().manual_seed(M * 1000 + N)
    x = torch.randn(M, K, generator=g).cuda()
    w = (torch.randn(N, K, generator=g) / K ** 0.5).cuda()
    b = torch.randn(N, generator=g).cuda()
    res = torch.randn(M, N, generator=g).cuda()
    row0 = torch.randn(N, generator=g).cuda()
    out = torch.empty(M, N, device='cuda')
    for relu, scale_cols, use_res, p0 in [(0, 0, False, 5), (1, 0, True, 5), (0, N // 3, True, 3), (0, 0, False, 0)]:
        pos = torch.tensor([p0], dtype=torch.int32, device='cuda')
        _native.call('vqcpc_linear_small', _pp(x), K, _pp(w), _pp(b), _pp(res) if use_res else None, N, _pp(out), N, M, N, K, relu,
                     0.125, scale_cols, _pp(row0), _pp(pos), ops._stream())
        ref = x.double() @ w.double().t() + b.double()
        ref[:, :scale_cols] *= 0.125
        if relu:
            ref = ref.clamp(min=0)
        if use_res:
            ref = ref + res.double()
        if p0 == 0:
            ref = row0.double().expand(M, N)
        assert rel_err(out, ref) < 2e-6


def test_sample_step_vs_reference_golden():
    """vqcpc_sample_step fed with numpy's uniform stream draws the tokens the reference's top_k_top_p_filtering +
    np.random.choice drew (tests/golden/decoder.pt 'sampler', made by the live reference)."""
    import numpy as np
    from VQCPCB import _native, ops
    for rec in torch.load(GOLD)['sampler']:
        lg = rec['logits'].cuda().contiguous()
        n, V = lg.shape
        u = torch.from_numpy(np.random.RandomState(rec['seed']).random_sample(n)).cuda()
        # rows are independent "sequences" of one step each: position 0, batch = n, u indexed [0 * n + b]
        tokens = torch.full((n, 1), -1, dtype=torch.int64, device='cuda')
        pos = torch.zeros(1, dtype=torch.int32, device='cuda')
        _native.call('vqcpc_sample_step', _pp(lg), V, V, rec['temperature'], rec['top_k'], rec['top_p'], _pp(u), n, _pp(tokens), 1,
                     _pp(pos), 0, n, ops._stream())
        got = tokens.view(-1).cpu()
        bad = (got != rec['tokens']).nonzero().view(-1).tolist()
        assert not bad, (rec['V'], rec['top_k'], rec['top_p'], bad, got[bad], rec['tokens'][bad])
        _native.call('vqcpc_sample_step', _pp(lg), V, V, 1.0, 0, 0.0, None, n, _pp(tokens), 1, _pp(pos), 1, n, ops._stream())
        assert torch.equal(tokens.view(-1).cpu(), rec['logits'].argmax(dim=-1))


@pytest.mark.parametrize('case', ['ac_ac', 'c_full', 'ac_diag', 'diag_noq', 'absolute'])
def test_cached_generation_equals_full_recompute(case):
    """The incremental path (key / value cache, CUDA graph per event, device sampling) against the reference's schedule (full
    forward per token) in exact-fp32 mode: greedy tokens identical; sampled tokens identical to the oracle sampler applied to the
    full-recompute logits with the same numpy uniform stream; graph replay == eager launches."""
    import numpy as np
    from VQCPCB import ops
    rec = torch.load(GOLD)[case]
    prm = D.make_decoder_params(rec['shapes'], seed=0)
    ops.set_gemm_mode('fp32')
    try:
        dec = _build(rec['cfg'])
        dec.load_state_dict({k: v.cuda() for k, v in prm.items()}, strict=True)
        ops.bump_weight_version()
        src = rec['source'].cuda()
        full = dec.generate_from_codes(src, greedy=True, kv_cache=False)
        cached = dec.generate_from_codes(src, greedy=True, kv_cache=True)
        eager = dec.generate_from_codes(src, greedy=True, kv_cache=True, use_graph=False)
        assert torch.equal(full, cached) and torch.equal(cached, eager)
        # sampling: replay the reference loop with the oracle sampler on the product's full-forward logits
        B, Ev, C = full.shape
        T = Ev * C
        seed = 4242
        u = np.random.RandomState(seed).random_sample((T, B))
        got = dec.generate_from_codes(src, temperature=0.9, top_k=12, top_p=0.9, uniforms=torch.from_numpy(u))
        rng = np.random.RandomState(seed)
        x = torch.zeros(B, Ev, C, dtype=torch.int64, device='cuda')
        for e in range(Ev):
            for c in range(C):
                lg = dec.forward(src, x)['weights_per_category'][c][:, e, :].cpu()
                x[:, e, c] = torch.tensor(D.sample_rows(lg, 0.9, 12, 0.9, rng), device='cuda')
        assert torch.equal(got, x)
        # excluded symbols are never drawn
        ex = dec.generate_from_codes(src, temperature=1.5, uniforms=torch.from_numpy(u), exclude={0: [0, 1, 2, 3]})
        assert int(ex[:, :, 0].min()) >= 4
    finally:
        ops.set_gemm_mode('bf16x3')


def test_cached_generation_teacher_forced_logits():
    """One cached step at every position reproduces the teacher-forced logits of Decoder.forward (rows of the golden target)."""
    from VQCPCB import ops
    rec = torch.load(GOLD)['c_full']
    prm = D.make_decoder_params(rec['shapes'], seed=0)
    dec = _build(rec['cfg'])
    dec.load_state_dict({k: v.cuda() for k, v in prm.items()}, strict=True)
    ops.bump_weight_version()
    x = dec.generate_from_codes(rec['source'].cuda(), greedy=True)
    logits = dec.forward(rec['source'].cuda(), x)['weights_per_category']
    for c, w in enumerate(logits):
        assert torch.equal(w.argmax(dim=-1), x[:, :, c])


def test_ln_linear_small_vs_fp64():
    """vqcpc_ln_linear_small: LayerNorm on load + linear, normalised rows written once."""
    from VQCPCB import _native, ops
    g = torch.Generator().manual_seed(9)
    for M, N, K in [(1, 512, 512), (7, 1536, 512), (16, 64, 512), (24, 1024, 512), (32, 64, 128), (12, 96, 1024)]:
        x = (torch.randn(M, K, generator=g) * 3 + 1).cuda()
        w = (torch.randn(N, K, generator=g) / K ** 0.5).cuda()
        b, gam, bet = (torch.randn(n, generator=g).cuda() for n in (N, K, K))
        out, lno = torch.empty(M, N, device='cuda'), torch.empty(M, K, device='cuda')
        _native.call('vqcpc_ln_linear_small', _pp(x), K, _pp(gam), _pp(bet), 1e-5, _pp(lno), _pp(w), _pp(b), None, N, _pp(out), N, M, N, K,
                     1, 1.0, 0, ops._stream())
        ln = torch.nn.functional.layer_norm(x.double(), (K,), gam.double(), bet.double(), 1e-5)
        assert rel_err(lno, ln) < 2e-6
        assert rel_err(out, (ln @ w.double().t() + b.double()).clamp(min=0)) < 3e-6


@pytest.mark.parametrize('case', ['ac_ac', 'c_full', 'ac_diag', 'absolute'])
def test_cached_generation_launch_modes(case, monkeypatch):
    """Graph replay with and without programmatic dependent launch at batch 1 / 3 / 5 (register and TMA-staged linears) against
    the full recompute.  Regression test: with overlapping kernels the activations / caches / position must be read through
    L2-coherent loads - the non-coherent path (const __restrict__) served stale rows to the aligned-layer chain."""
    from VQCPCB import ops
    rec = torch.load(GOLD)[case]
    prm = D.make_decoder_params(rec['shapes'], seed=0)
    ops.set_gemm_mode('fp32')
    try:
        dec = _build(rec['cfg'])
        dec.load_state_dict({k: v.cuda() for k, v in prm.items()}, strict=True)
        ops.bump_weight_version()
        S = rec['source'].shape[1]
        for B in (1, 3, 5) + ((21,) if case == 'ac_diag' else ()):
            src = torch.randint(0, rec['cfg']['codebook_size'], (B, S), generator=torch.Generator().manual_seed(B)).cuda()
            full = dec.generate_from_codes(src, greedy=True, kv_cache=False)
            for pdl in ('0', '1'):
                monkeypatch.setenv('VQCPC_DECODE_PDL', pdl)
                dec.__dict__.pop('_decode_sessions', None)
                for _ in range(2):                                       # second call replays the cached graph
                    assert torch.equal(dec.generate_from_codes(src, greedy=True), full), (case, B, pdl)
    finally:
        ops.set_gemm_mode('bf16x3')


@pytest.mark.parametrize('decoder_type', ['transformer_relative', 'transformer_relative_diagonal', 'transformer'])
def test_main_decoder_generates_end_to_end(decoder_type, tmp_path):
    """`main_decoder.py -c configs/decoder_relative.py` as a subprocess (the reference's inference entry point, main_decoder.py:
    16-160): real Encoder built by the getters (codes from its quantizer, `merge_codes`), decoder from `get_decoder`, a window from
    the loader, `Decoder.generate` on the cached path, re-encoding of [original; generated]."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, VQCPC_BLOCKS='2', VQCPC_CODEBOOK='16', VQCPC_DECODER_BEATS='4', VQCPC_DECODER_TYPE=decoder_type,
               VQCPC_DATASET='synthetic')
    cfg = os.path.join(root, 'vqcpc-bach_b200', 'VQCPCB', 'configs', 'decoder_relative.py')
    r = subprocess.run([sys.executable, os.path.join(root, 'main_decoder.py'), '-c', cfg, '--num_examples', '1', '--batch_size', '2',
                        '--eval_batches', '2'],
                       cwd=str(tmp_path), env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert 'tokens/s' in r.stdout
    import re
    m = re.search(r"validation: \{'loss': ([0-9.eE+-]+)\}", r.stdout)
    assert m and 0 < float(m.group(1)) < 100, r.stdout[-500:]
    models = os.listdir(tmp_path / 'models')
    assert len(models) == 1 and models[0].startswith('decoder_relative_')
    rec = torch.load(tmp_path / 'models' / models[0] / 'generation_0.pt')
    assert rec['original'].shape == (2, 16, 4) and rec['generated'].shape == (2, 16, 4)
    assert int(rec['generated'].min()) >= 0 and int(rec['generated'].max()) < 64
    assert rec['recoding'].shape == (3, 4) and int(rec['recoding'].max()) < 16          # [original; 2 generated] x 4 codes


@pytest.mark.parametrize('L,mode', [(64, 1), (100, 1), (100, 2), (48, 2), (37, 0), (256, 1)])
def test_attn_long_masked_vs_fp64(L, mode):
    """vqcpc_attn_long_fwd_masked (tiled self-attention with the decoder's causal / anticausal masks, skipped key tiles) against
    the fp64 oracle attention (literal relative bias + additive mask)."""
    from VQCPCB import _native, ops
    H, hd, P = 2, 64, 3
    d = H * hd
    g = torch.Generator().manual_seed(L * 10 + mode)
    qkv = torch.randn(P, L, 3 * d, generator=g) * 0.5
    e1 = torch.randn(H * L, hd, generator=g) * 0.3
    e2 = torch.randn(H * L, hd, generator=g) * 0.3
    q64, k64, v64 = (t.double() for t in qkv.split(d, dim=-1))
    sp = lambda t: t.reshape(P, L, H, hd).transpose(1, 2).reshape(P * H, L, hd)      # noqa: E731
    s = torch.bmm(sp(q64), sp(k64).transpose(1, 2)) + D.rel_bias_general(sp(q64), e1.double(), e2.double(), L, L)
    if mode == 1:
        s = s + D.causal_mask(L).double()
    elif mode == 2:
        s = s + D.causal_mask(L).t().double()
    a = torch.softmax(s, dim=-1)
    ref = torch.bmm(a, sp(v64)).reshape(P, H, L, hd).transpose(1, 2).reshape(P * L, d)
    qc, e1c, e2c = qkv.reshape(P * L, 3 * d).cuda(), e1.cuda(), e2.cuda()
    out = torch.empty(P * L, d, device='cuda')
    probs = torch.full((P, H, L, L), 7.0, device='cuda')
    _native.call('vqcpc_attn_long_fwd_masked', _pp(qc), _pp(e1c), _pp(e2c), _pp(out), _pp(probs), P, L, H, hd, mode, ops._stream())
    torch.cuda.synchronize()
    assert rel_err(out, ref) < 3e-6
    assert rel_err(probs, a.reshape(P, H, L, L)) < 3e-6


@pytest.mark.parametrize('case', ['ac_ac', 'ac_diag', 'absolute'])
def test_long_generation_cached_equals_window_recompute(case):
    """generate_from_code_long (decoder.py:729-829): a code sequence longer than the model window, decoded code by code from the
    sliding window.  The cached path (one teacher-forced forward per code refills the key / value cache, then incremental
    steps) against the reference's schedule (a full forward of the window per token), greedy and sampled from the same uniforms;
    and the window arithmetic against the reference's table of cases."""
    from VQCPCB import ops
    rec = torch.load(GOLD)[case]
    prm = D.make_decoder_params(rec['shapes'], seed=0)
    ops.set_gemm_mode('fp32')
    try:
        dec = _build(rec['cfg'])
        dec.load_state_dict({k: v.cuda() for k, v in prm.items()}, strict=True)
        ops.bump_weight_version()
        assert [dec.compute_start_end_times(t, 9, 4) for t in range(9)] == [
            (0, 4, 0), (0, 4, 1), (0, 4, 2), (1, 5, 2), (2, 6, 2), (3, 7, 2), (4, 8, 2), (5, 9, 2), (5, 9, 3)]
        g = torch.Generator().manual_seed(5)
        codes = torch.randint(0, rec['cfg']['codebook_size'], (2, 9), generator=g).cuda()
        full = dec.generate_from_code_long(codes, greedy=True, kv_cache=False)
        cached = dec.generate_from_code_long(codes, greedy=True)
        assert full.shape == (2, 36, 4) and torch.equal(full, cached)
        u = torch.rand(5 * 16, 2, dtype=torch.float64, generator=g)
        a = dec.generate_from_code_long(codes, temperature=0.9, top_p=0.9, code_index_start=2, code_index_end=7, uniforms=u,
                                        kv_cache=False)
        b = dec.generate_from_code_long(codes, temperature=0.9, top_p=0.9, code_index_start=2, code_index_end=7, uniforms=u)
        assert a.shape == (2, 20, 4) and torch.equal(a, b)
    finally:
        ops.set_gemm_mode('bf16x3')


@pytest.mark.parametrize('P,H,Lk,pos,mode', [(3, 2, 64, 37, 'self'), (40, 8, 512, 300, 'self'), (40, 8, 512, 511, 'self'),
                                             (2, 8, 1024, 1000, 'self'), (5, 2, 16, 77, 'cross'), (40, 8, 64, 200, 'cross'),
                                             (4, 2, 96, 50, 'nobias')])
def test_attn_decode_kernel_vs_fp64(P, H, Lk, pos, mode):
    """vqcpc_attn_decode on its own (both register variants: few CTAs -> value prefetch, P * H * splits > 2 x SMs -> the
    occupancy variant; cluster splits 1..8): one query row over the cache with the relative bias through the map row, the
    appended key / value, causal length or additive mask - against fp64."""
    from VQCPCB import _native, ops
    from VQCPCB.transformer.relative_maps import device_maps, relative_index_maps
    hd = 64
    d = H * hd
    g = torch.Generator().manual_seed(P * 1000 + Lk + pos)
    self_attn = mode != 'cross'
    Lq = Lk if self_attn else Lk * 8                       # cross: T = 8 S
    assert pos < Lq
    q = (torch.randn(P, d, generator=g) * 0.4).cuda()
    kc = (torch.randn(P * Lk, d, generator=g) * 0.5).cuda()
    vc = torch.randn(P * Lk, d, generator=g).cuda()
    kn, vn = (torch.randn(P, d, generator=g) * 0.5).cuda(), torch.randn(P, d, generator=g).cuda()
    e1 = (torch.randn(H * Lk, hd, generator=g) * 0.3).cuda()
    e2 = (torch.randn(H * Lk, hd, generator=g) * 0.3).cuda()
    m1 = m2 = mask = None
    if mode != 'nobias':
        m1, m2 = device_maps(Lq, Lk, 'cuda')
    if mode == 'cross':
        mask = torch.repeat_interleave(D.causal_mask(Lk).t(), Lq // Lk, dim=0).contiguous().cuda()
    posd = torch.tensor([pos], dtype=torch.int32, device='cuda')
    out = torch.empty(P, d, device='cuda')
    kc0, vc0 = kc.clone(), vc.clone()
    _native.call('vqcpc_attn_decode', _pp(q), d, _pp(kn) if self_attn else None, _pp(vn) if self_attn else None, d, _pp(kc), _pp(vc), d, Lk,
                 _pp(e1) if m1 is not None else None, _pp(e2) if m2 is not None else None, _pp(m1), _pp(m2), _pp(mask), Lk, _pp(posd),
                 int(self_attn), _pp(out), d, P, H, hd, ops._stream())
    torch.cuda.synchronize()
    # fp64 reference
    K, V = kc0.double().view(P, Lk, H, hd), vc0.double().view(P, Lk, H, hd)
    if self_attn:                                           # the new row is appended at `pos`, keys 0 .. pos are visible
        K[:, pos], V[:, pos] = kn.double().view(P, H, hd), vn.double().view(P, H, hd)
        assert torch.equal(kc.view(P, Lk, d)[:, pos], kn) and torch.equal(vc.view(P, Lk, d)[:, pos], vn)
        n = pos + 1
    else:
        n = Lk
    qd = q.double().view(P, H, hd)
    s = torch.einsum('phd,pjhd->phj', qd, K[:, :n])
    if m1 is not None:
        for mm, e in ((m1, e1), (m2, e2)):
            row = mm[pos, :n].cpu().long()
            eh = e.double().view(H, Lk, hd)
            valid = row >= 0
            assert bool((row[valid] // Lk == pos).all())
            rel = torch.einsum('phd,hjd->phj', qd, eh[:, (row.clamp(min=0) % Lk).cuda()])
            s = s + torch.where(valid.cuda(), rel, torch.where((row == -2).cuda(), -100.0, 0.0).double())
    if mask is not None:
        s = s + mask[pos, :n].double()
    a = torch.softmax(s, dim=-1)
    ref = torch.einsum('phj,pjhd->phd', a, V[:, :n]).reshape(P, d)
    assert rel_err(out, ref) < 3e-6


def test_reharmonisation_and_hand_written_codes():
    """Host wrappers of decoder.py:856-1017 over generate_from_code_long: hand-written code sequences, reharmonisation of a token
    tensor (window framing, glued codes, trimmed code range), plagiarism helpers."""
    rec = torch.load(GOLD)['ac_ac']
    prm = D.make_decoder_params(rec['shapes'], seed=0)
    dec = _build(rec['cfg'])
    dec.load_state_dict({k: v.cuda() for k, v in prm.items()}, strict=True)

    class FakeEncoder(torch.nn.Module):               # 16 events x 4 voices -> 4 codes per window, one codebook
        downscaler = types.SimpleNamespace(downscale_factors=[4, 4])
        quantizer = types.SimpleNamespace(codebook_size=16, num_codebooks=1, codebook_dim=3)

        def forward(self, x):
            codes = x.cuda().long().view(x.shape[0], 4, 16).sum(-1) % 16
            return None, codes.unsqueeze(-1), None

        def merge_codes(self, c):
            return c[:, :, 0]

    dec.encoder = FakeEncoder()
    out = dec.generate_alla_mano([1, 2, 3, 4], [5, 6, 7, 8], [9, 10, 11, 12, 13, 14], temperature=1.0, num_decodings=2, greedy=True)
    assert out.shape == (2, 24, 4) and torch.equal(out[0], out[1])
    codes = torch.tensor([[1, 2, 3, 4, 9, 10, 11, 12, 13, 14, 5, 6, 7, 8]]).cuda()
    want = dec.generate_from_code_long(codes, greedy=True, code_index_start=4, code_index_end=10, kv_cache=False)
    assert torch.equal(out[:1], want)
    # 40 events = 2.5 windows: [start frame][w0][w1][half window + END + PAD][PAD frame] -> 5 windows of 4 codes, 10 decoded
    g = torch.Generator().manual_seed(2)
    x = torch.stack([torch.randint(0, v, (1, 40), generator=g) for v in rec['cfg']['vocab']], dim=2)
    re = dec.generate_reharmonisation(2, temperature=1.0, top_k=0, top_p=0.9, x=x, greedy=True)
    assert re.shape == (2, 40, 4)
    s, n = dec.check_duplicate(re[0].cpu(), re[1].cpu())
    assert n == (len(dec.data_processor.dump(re[0].cpu())) - 1) / 3


def test_decode_small_kernels_vs_torch():
    """vqcpc_elu_fwd, vqcpc_decode_ln_add, vqcpc_decode_features (both feature modes), vqcpc_decode_advance on their own."""
    from VQCPCB import _native, ops
    g = torch.Generator().manual_seed(0)
    st = ops._stream()
    x = (torch.randn(1000, 37, generator=g) * 3).cuda()
    out = torch.empty_like(x)
    _native.call('vqcpc_elu_fwd', _pp(x), _pp(out), x.numel(), st)
    assert rel_err(out, torch.nn.functional.elu(x.double())) < 1e-6
    B, T, d = 5, 12, 192
    y, add = (torch.randn(B, d, generator=g) * 2 + 0.5).cuda(), torch.randn(B, T, d, generator=g).cuda()
    gam, bet = torch.randn(d, generator=g).cuda(), torch.randn(d, generator=g).cuda()
    pos = torch.tensor([7], dtype=torch.int32, device='cuda')
    o = torch.empty(B, d, device='cuda')
    _native.call('vqcpc_decode_ln_add', _pp(y), _pp(gam), _pp(bet), 1e-5, _pp(add), T, _pp(pos), _pp(o), B, d, st)
    ref = torch.nn.functional.layer_norm(y.double(), (d,), gam.double(), bet.double(), 1e-5) + add[:, 7].double()
    assert rel_err(o, ref) < 2e-6
    # feature rows: [embedding of token pos-1 | channel | event-in-code] and [embedding | absolute position]
    C, up, E, pe = 4, 16, 8, 3
    vocab = [5, 6, 7, 8]
    tables = torch.randn(sum(vocab), E, generator=g).cuda()
    off = torch.tensor([0, 5, 11, 18], dtype=torch.int32, device='cuda')
    T = 32
    tokens = torch.stack([torch.randint(0, vocab[t % C], (B,), generator=g) for t in range(T)], dim=1).cuda()
    chan, evt, tpos = torch.randn(C, pe, generator=g).cuda(), torch.randn(up // C, pe, generator=g).cuda(), torch.randn(T, pe, generator=g).cuda()
    for p in (0, 1, 6, 31):
        pos.fill_(p)
        f0, f1 = torch.full((B, E + 2 * pe), 9.0, device='cuda'), torch.full((B, E + pe), 9.0, device='cuda')
        _native.call('vqcpc_decode_features', _pp(tokens), _pp(pos), _pp(tables), _pp(off), _pp(chan), _pp(evt), _pp(f0), B, T, C, up, E,
                     pe, 0, st)
        _native.call('vqcpc_decode_features', _pp(tokens), _pp(pos), _pp(tables), _pp(off), _pp(tpos), _pp(tpos), _pp(f1), B, T, C, up, E,
                     pe, 1, st)
        if p == 0:
            assert float(f0.abs().max()) == 0.0 and float(f1.abs().max()) == 0.0
            continue
        t = p - 1
        emb = tables[off[t % C].long() + tokens[:, t]]
        assert torch.equal(f0, torch.cat([emb, chan[t % C].expand(B, pe), evt[(t % up) // C].expand(B, pe)], dim=1))
        assert torch.equal(f1, torch.cat([emb, tpos[t].expand(B, pe)], dim=1))
    _native.call('vqcpc_decode_advance', _pp(pos), 3, st)
    assert int(pos.item()) == 34
