"""GPU: the two GEMM kernels against an fp64 torch reference, every layout and fused epilogue, through the C-ABI."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref(a, b, a_k, b_k, bias=None, residual=None, relu=False, mask=None, scale=None):
    A = a.double() if a_k else a.double().t()
    Bm = b.double().t() if b_k else b.double()
    c = A @ Bm
    if bias is not None:
        c = c + bias.double()
    if scale is not None:
        c[:, :scale[1]] *= scale[0]
    if residual is not None:
        c = c + residual.double()
    if relu:
        c = c.clamp_min(0)
    if mask is not None:
        c = c * (mask > 0)
    return c


def _mk(M, N, K, a_k, b_k, gen):
    a = torch.randn((M, K) if a_k else (K, M), generator=gen, device='cuda')
    b = torch.randn((N, K) if b_k else (K, N), generator=gen, device='cuda')
    return a, b


LAYOUTS = [(True, True), (True, False), (False, False), (False, True)]
SHAPES = [(128, 128, 64), (256, 384, 512), (200, 96, 136), (1, 32, 32), (515, 160, 72), (130, 1536, 512)]


@pytest.mark.parametrize('mode', ['fp32', 'bf16x3', 'bf16'])
@pytest.mark.parametrize('a_k,b_k', LAYOUTS)
def test_gemm_layouts_and_shapes(mode, a_k, b_k):
    from VQCPCB import ops
    ops.set_gemm_mode(mode)
    gen = torch.Generator(device='cuda')
    gen.manual_seed(1)
    tol = {'fp32': 2e-6, 'bf16x3': 3e-5, 'bf16': 1.5e-2}[mode]
    worst = 0.0
    for (M, N, K) in SHAPES:
        a, b = _mk(M, N, K, a_k, b_k, gen)
        out = ops.gemm(a, b, M, N, K, a_k, b_k)
        torch.cuda.synchronize()
        ref = _ref(a, b, a_k, b_k)
        err = ((out.double() - ref).norm() / ref.norm()).item()
        mx = ((out.double() - ref).abs().max() / ref.abs().max()).item()
        print(f'{mode} a_k={a_k} b_k={b_k} {M}x{N}x{K}: rel {err:.2e} max {mx:.2e}')
        worst = max(worst, err)
        assert err < tol, (mode, a_k, b_k, M, N, K, err)
    ops.set_gemm_mode('bf16x3')


@pytest.mark.parametrize('mode', ['fp32', 'bf16x3'])
def test_gemm_epilogues(mode):
    from VQCPCB import ops
    ops.set_gemm_mode(mode)
    gen = torch.Generator(device='cuda')
    gen.manual_seed(2)
    tol = 2e-6 if mode == 'fp32' else 3e-5
    M, N, K = 300, 256, 192
    a, b = _mk(M, N, K, True, True, gen)
    bias = torch.randn(N, generator=gen, device='cuda')
    res = torch.randn(M, N, generator=gen, device='cuda')
    mask = torch.randn(M, N, generator=gen, device='cuda')

    def check(name, out, ref):
        torch.cuda.synchronize()
        err = ((out.double() - ref).norm() / ref.norm()).item()
        print(f'{mode} epilogue {name}: {err:.2e}')
        assert err < tol, (name, err)

    check('bias', ops.gemm(a, b, M, N, K, bias=bias), _ref(a, b, True, True, bias=bias))
    check('bias+relu', ops.gemm(a, b, M, N, K, bias=bias, relu=True), _ref(a, b, True, True, bias=bias, relu=True))
    check('bias+residual', ops.gemm(a, b, M, N, K, bias=bias, residual=res), _ref(a, b, True, True, bias=bias, residual=res))
    check('relu_mask', ops.gemm(a, b, M, N, K, relu_mask=mask), _ref(a, b, True, True, mask=mask))
    check('bias+scale_cols', ops.gemm(a, b, M, N, K, bias=bias, scale_cols=(0.125, 128)),
          _ref(a, b, True, True, bias=bias, scale=(0.125, 128)))
    # strided output (ldc > N) as used by the input projection
    big = torch.full((M, N + 16), 7.0, device='cuda')
    ops.gemm(a, b, M, N, K, out=big[:, :N], bias=bias)
    check('strided out', big[:, :N], _ref(a, b, True, True, bias=bias))
    assert bool((big[:, N:] == 7.0).all())
    # accumulate and split-K (weight-gradient style: reduction over many rows)
    a2, b2 = _mk(160, 192, 4096, False, False, gen)
    base = torch.randn(160, 192, generator=gen, device='cuda')
    acc = base.clone()
    ops.gemm(a2, b2, 160, 192, 4096, False, False, out=acc, accumulate=True, ksplit=8)
    check('splitk accumulate', acc, _ref(a2, b2, False, False) + base.double())
    fresh = torch.full((160, 192), 3.0, device='cuda')
    ops.gemm(a2, b2, 160, 192, 4096, False, False, out=fresh, ksplit=5)
    check('splitk overwrite', fresh, _ref(a2, b2, False, False))
    ops.set_gemm_mode('bf16x3')


def test_bf16x3_split_reconstructs_fp32():
    from VQCPCB import ops
    x = torch.randn(257, 96, device='cuda') * 3
    hi, lo = ops.split_bf16(x)
    torch.cuda.synchronize()
    assert hi.dtype == torch.bfloat16 and torch.equal(hi, x.to(torch.bfloat16))
    rec = hi.float() + lo.float()
    assert ((rec - x).abs() <= x.abs() * 2 ** -15).all()
    # strided source (a column slice)
    big = torch.randn(64, 512, device='cuda')
    hi2, lo2 = ops.split_bf16(big[:, :496])
    assert torch.equal(hi2, big[:, :496].to(torch.bfloat16))


def test_gemm_large_m_persistent_schedule():
    """More tiles than SMs x accumulator stages: exercises the smem ring / TMEM double buffering wrap-around."""
    from VQCPCB import ops
    ops.set_gemm_mode('bf16x3')
    gen = torch.Generator(device='cuda')
    gen.manual_seed(3)
    M, N, K = 128 * 700 + 37, 512, 512
    a, b = _mk(M, N, K, True, True, gen)
    bias = torch.randn(N, generator=gen, device='cuda')
    out = ops.gemm(a, b, M, N, K, bias=bias, relu=True)
    torch.cuda.synchronize()
    idx = torch.randint(0, M, (4096,), device='cuda')
    ref = (a[idx].double() @ b.double().t() + bias.double()).clamp_min(0)
    err = ((out[idx].double() - ref).norm() / ref.norm()).item()
    print('large-M bf16x3 rel err', err)
    assert err < 3e-5
    assert torch.isfinite(out).all()


@pytest.mark.parametrize('mode', ['bf16x3'])
def test_gemm_presplit_operands_split_outputs_and_bf16_mask(mode):
    """Fused-split data flow: operands given as (hi, lo), result written as fp32 and/or (hi, lo), ReLU mask read from the
    bf16 hi half of a split activation."""
    from VQCPCB import ops
    ops.set_gemm_mode(mode)
    gen = torch.Generator(device='cuda')
    gen.manual_seed(9)
    M, N, K = 333, 2048, 512
    a, b = _mk(M, N, K, True, True, gen)
    bias = torch.randn(N, generator=gen, device='cuda')
    asp = ops.split_bf16(a)
    ref = _ref(a, b, True, True, bias=bias, relu=True)
    c, sp = ops.gemm(None, b, M, N, K, bias=bias, relu=True, a_split=asp, b_is_weight=False, b_split=ops.split_bf16(b),
                     out_split=True)
    torch.cuda.synchronize()
    assert ((c.double() - ref).norm() / ref.norm()).item() < 3e-5
    assert torch.equal(sp[0], c.to(torch.bfloat16))
    assert ((sp[0].double() + sp[1].double() - c.double()).norm() / c.double().norm()).item() < 1e-5
    none_c, sp2 = ops.gemm(None, b, M, N, K, bias=bias, relu=True, a_split=asp, b_split=ops.split_bf16(b), want_c=False,
                           out_split=True)
    assert none_c is None and torch.equal(sp2[0], sp[0]) and torch.equal(sp2[1], sp[1])
    # dgrad with the mask taken from the bf16 hi half:  d = (g @ W) * (h > 0)
    g = torch.randn(M, 512, generator=gen, device='cuda')
    w = torch.randn(512, N, generator=gen, device='cuda')          # (N_out=512, K_in=2048) weight, read untransposed
    d_ref = (g.double() @ w.double()) * (c.double() > 0)
    _, dsp = ops.gemm(g, w, M, N, 512, True, False, relu_mask=sp[0], want_c=False, out_split=True)
    d = dsp[0].double() + dsp[1].double()
    assert ((d - d_ref).norm() / d_ref.norm()).item() < 3e-5
    # wgrad from two split operands (both MN-major), accumulated with split-K
    acc = torch.zeros(512, N, device='cuda')
    ops.gemm(None, None, 512, N, M, False, False, out=acc, accumulate=True, ksplit=2, a_split=ops.split_bf16(g), b_split=sp)
    w_ref = g.double().t() @ c.double()
    assert ((acc.double() - w_ref).norm() / w_ref.norm()).item() < 3e-5
    ops.set_gemm_mode('bf16x3')


@pytest.mark.parametrize('M', [300, 4096 + 77])
def test_gemm_static_epilogue_configurations_equal_runtime_flags(M):
    """The CTA-pair kernel is instantiated per compile-time epilogue configuration for the encoder's forward / data-gradient
    GEMMs (csrc/gemm_tc2.cu); `vqcpc_gemm_set_experiment(256)` forces the run-time-flag kernel instead.  Both must give
    bit-identical results on every configuration (incl. a partial last row tile), and match the fp64 reference."""
    from VQCPCB import ops, _native
    ops.set_gemm_mode('bf16x3')
    gen = torch.Generator(device='cuda')
    gen.manual_seed(3)
    d, F = 512, 1024
    x = torch.randn(M, d, generator=gen, device='cuda')
    h = torch.randn(M, F, generator=gen, device='cuda')
    w_qkv = torch.randn(3 * d, d, generator=gen, device='cuda') * 0.05
    w1 = torch.randn(F, d, generator=gen, device='cuda') * 0.05           # (out F, in d)
    w2 = torch.randn(d, F, generator=gen, device='cuda') * 0.05           # (out d, in F)
    b3, bF, bd = (torch.randn(n, generator=gen, device='cuda') for n in (3 * d, F, d))
    res = torch.randn(M, d, generator=gen, device='cuda')
    resF = torch.randn(M, F, generator=gen, device='cuda')
    mask = torch.randn(M, F, generator=gen, device='cuda').to(torch.bfloat16)
    xs, hs = ops.split_bf16(x), ops.split_bf16(h)

    def dd(t):
        return t.double()
    cases = {
        'qkv bias+scale': (lambda: ops.gemm(None, w_qkv, M, 3 * d, d, bias=b3, scale_cols=(0.125, d), b_is_weight=True, a_split=xs),
                           lambda: torch.cat([(dd(x) @ dd(w_qkv).t() + dd(b3))[:, :d] * 0.125, (dd(x) @ dd(w_qkv).t() + dd(b3))[:, d:]], 1)),
        'out bias+residual': (lambda: ops.gemm(None, w2, M, d, F, bias=bd, residual=res, b_is_weight=True, a_split=hs),
                              lambda: dd(h) @ dd(w2).t() + dd(bd) + dd(res)),
        'ffn1 bias+relu->split': (lambda: ops.gemm(None, w1, M, F, d, bias=bF, relu=True, b_is_weight=True, a_split=xs, want_c=False, out_split=True),
                                  lambda: torch.relu(dd(x) @ dd(w1).t() + dd(bF))),
        'bias only': (lambda: ops.gemm(None, w1, M, F, d, bias=bF, b_is_weight=True, a_split=xs),
                      lambda: dd(x) @ dd(w1).t() + dd(bF)),
        'bias+relu': (lambda: ops.gemm(None, w1, M, F, d, bias=bF, relu=True, b_is_weight=True, a_split=xs),
                      lambda: torch.relu(dd(x) @ dd(w1).t() + dd(bF))),
        'dgrad bf16 mask->split': (lambda: ops.gemm(None, w2, M, F, d, True, False, relu_mask=mask, b_is_weight=True, a_split=xs, want_c=False, out_split=True),
                                   lambda: (dd(x) @ dd(w2)) * (mask.double() > 0)),
        'dgrad plain': (lambda: ops.gemm(None, w2, M, F, d, True, False, b_is_weight=True, a_split=xs),
                        lambda: dd(x) @ dd(w2)),
        'dgrad residual': (lambda: ops.gemm(None, w2, M, F, d, True, False, residual=resF, b_is_weight=True, a_split=xs),
                           lambda: dd(x) @ dd(w2) + dd(resF)),
        'dgrad scale+bf16 mask->split': (lambda: ops.gemm(None, w2, M, F, d, True, False, relu_mask=mask, scale_cols=(1.25, F), b_is_weight=True,
                                                          a_split=xs, want_c=False, out_split=True),
                                         lambda: (dd(x) @ dd(w2)) * 1.25 * (mask.double() > 0)),
    }
    try:
        for name, (run, ref) in cases.items():
            outs = []
            for force_generic in (0, 1):
                _native.call('vqcpc_gemm_set_experiment', 256 * force_generic)
                r = run()
                torch.cuda.synchronize()
                outs.append(r[1] if isinstance(r, tuple) else (r,))
            for a, b in zip(*outs):
                assert torch.equal(a, b), name
            val = outs[0][0].double() + (outs[0][1].double() if len(outs[0]) == 2 else 0)
            rr = ref()
            assert ((val - rr).norm() / rr.norm()).item() < 3e-5, name
    finally:
        _native.call('vqcpc_gemm_set_experiment', 0)
