"""GPU: tensor-core nearest-code search (vq_tc.cu) - bit-exact against the direct-form kernel and the CPU oracle.

The reference is quantizer/vector_quantizer.py:104-116 (broadcast distance + argmin, first-min tie rule); the oracle
restates it in oracle/vqcpc_oracle.py::quantize.  The tcgen05 path only RANKS codes; every index it returns comes from
re-evaluating the reference arithmetic, so idx / loss / y / counts must be identical, bit for bit, on every input -
including exact ties, near-ties on bisector planes, duplicate codes, huge / tiny magnitudes and non-finite rows."""
import pytest
import torch

from oracle import vqcpc_oracle as O

pytestmark = pytest.mark.gpu


def run_fused(z, E, T, algo, dbg=False, zsg=False):
    from VQCPCB import _native, ops
    n, D = z.shape
    Kc, zd = T.shape
    dev = z.device
    idx = torch.full((n,), -7, dtype=torch.int64, device=dev)
    loss = torch.full((n,), float('nan'), device=dev)
    y = torch.full((n, zd), float('nan'), device=dev)
    counts = torch.zeros(Kc, dtype=torch.int32, device=dev)
    stats = torch.zeros(4098, dtype=torch.int32, device=dev)
    kpad = (Kc + 255) // 256 * 256
    scores = torch.full((n, kpad), float('nan'), device=dev) if dbg else None
    zs = torch.full((n, D), float('nan'), device=dev) if zsg else None
    _native.call('vqcpc_vq_fused_fwd_ex', ops._p(z), n, D, ops._p(E), ops._p(T), Kc, zd, 0.25, 1, ops._p(idx), ops._p(loss),
                 ops._p(zs), ops._p(y), ops._p(counts), algo, ops._p(scores), ops._p(stats), ops._stream())
    torch.cuda.synchronize()
    return dict(idx=idx, loss=loss, y=y, counts=counts, stats=stats, scores=scores, zsg=zs)


def adversarial_inputs(n, Kc, D, seed):
    g = torch.Generator(device='cuda')
    g.manual_seed(seed)
    E = torch.randn(Kc, D, device='cuda', generator=g) * 4
    if Kc >= 8:
        E[5] = E[2]                                     # duplicate codes: exact ties, lowest index must win
        E[Kc - 1] = E[Kc // 2]
    z = torch.randn(n, D, device='cuda', generator=g) * 3
    m = n // 8
    if m > 0:
        a = torch.randint(0, Kc, (m,), device='cuda', generator=g)
        b = torch.randint(0, Kc, (m,), device='cuda', generator=g)
        z[:m] = 0.5 * (E[a] + E[b])                     # on a bisector plane: distances equal up to rounding
        z[m:2 * m] = E[a] + 1e-6 * torch.randn(m, D, device='cuda', generator=g)     # on top of a code
        z[2 * m:3 * m] = z[2 * m:3 * m] * 1e4           # far outside the codebook
        z[3 * m:4 * m] = z[3 * m:4 * m] * 1e-6          # all at the origin
        z[4 * m:5 * m] = E[a]                           # exactly a code (zero distance, duplicates tie)
    return z.contiguous(), E.contiguous()


@pytest.mark.parametrize('D,Kc,n', [(3, 512, 4096), (3, 512, 100), (3, 512, 128 * 148 * 3 + 5), (3, 16, 1000), (3, 300, 5000), (4, 512, 3000), (3, 1024, 2048),
                                    (4, 257, 777), (3, 2048, 1500)])
def test_vq_tc_scores_and_indices_small(D, Kc, n):
    torch.manual_seed(0)
    z, E = adversarial_inputs(n, Kc, D, seed=Kc + n)
    T = torch.randn(Kc, 32, device='cuda')
    r2 = run_fused(z, E, T, 2, dbg=True, zsg=True)
    # the tensor-core scores approximate z.e - |e|^2/2 within HALF the error budget `eps` the kernel's margin assumes
    # (vq_tc.cu: eps = 2^-14 P + 2^-20 N + KB 2^-19 (P + N), P = sum_d |z_d| max_k |e_kd|, N = max_k |e_k|^2 / 2, KB = 1)
    zd_, Ed = z.double(), E.double()
    ref = zd_ @ Ed.t() - 0.5 * (Ed ** 2).sum(1)[None]
    P = zd_.abs() @ Ed.abs().max(0)[0]
    N = 0.5 * (Ed ** 2).sum(1).max()
    eps = 2.0 ** -14 * P + 2.0 ** -20 * N + 2.0 ** -19 * (P + N)
    ratio = ((r2['scores'][:, :Kc].double() - ref).abs() / eps[:, None]).max()
    assert float(ratio) < 0.5, f'score error reaches {float(ratio):.2f} of the budget the margin assumes'
    pad = r2['scores'][:, Kc:]
    assert bool((torch.isnan(pad) | (pad < -1e37)).all())                # padding codes (where computed) can never rank
    # indices: the CPU oracle (reference arithmetic) and the direct kernel
    _, oi, ol = O.quantize(z.cpu(), [E.cpu()], 0.25, True)
    assert torch.equal(r2['idx'].cpu(), oi.view(-1)), 'tensor-core search differs from the CPU oracle'
    r1 = run_fused(z, E, T, 1, zsg=True)
    for k in ('idx', 'loss', 'y', 'counts', 'zsg'):
        assert torch.equal(r1[k], r2[k]), f'{k} differs between the direct and the tensor-core search'
    assert torch.equal(r2['y'], T[r2['idx']])
    assert int(r2['counts'].sum()) == n
    assert torch.allclose(r2['loss'].cpu(), ol.view(-1), rtol=1e-6, atol=0)


def test_vq_tc_non_finite_rows():
    z, E = adversarial_inputs(2048, 512, 3, seed=9)
    z[7, 1] = float('nan')
    z[300] = float('inf')
    z[301, 0] = -float('inf')
    z[900] = 3e38
    z[901] = 1e25
    z[1200, 2] = -2e19
    T = torch.randn(512, 32, device='cuda')
    r1, r2 = run_fused(z, E, T, 1), run_fused(z, E, T, 2)
    assert torch.equal(r1['idx'], r2['idx'])
    assert int(r2['stats'][1]) > 0                                        # those rows took the full-scan fallback
    fin = torch.isfinite(r1['loss'])
    assert torch.equal(r1['loss'][fin], r2['loss'][fin])
    _, oi, _ = O.quantize(z.cpu(), [E.cpu()], 0.25, True)
    assert torch.equal(r2['idx'].cpu(), oi.view(-1))


@pytest.mark.parametrize('seed,scale', [(0, 2.0), (1, 0.05), (2, 40.0)])
def test_vq_tc_full_size_bit_exact(seed, scale):
    """BASELINE size: 557 056 vectors, K = 512, D = 3 - every output identical to the direct-form kernel."""
    n, Kc, D = 557056, 512, 3
    g = torch.Generator(device='cuda')
    g.manual_seed(seed)
    E = torch.randn(Kc, D, device='cuda', generator=g) * 2
    z = torch.randn(n, D, device='cuda', generator=g) * scale
    T = torch.randn(Kc, 32, device='cuda', generator=g)
    r1, r2 = run_fused(z, E, T, 1), run_fused(z, E, T, 2)
    for k in ('idx', 'loss', 'y', 'counts'):
        assert torch.equal(r1[k], r2[k]), f'{k} differs'
    # the verification stays rare: about one 8-code group per vector
    assert int(r2['stats'][0]) < 2 * n and int(r2['stats'][1]) == 0, r2['stats']


def test_vq_tc_data_init_codebook():
    """Codebook drawn from the data itself (quantizer 'initialize' mode): many vectors sit exactly on a code."""
    n, Kc, D = 200000, 512, 3
    g = torch.Generator(device='cuda')
    g.manual_seed(4)
    z = torch.randn(n, D, device='cuda', generator=g) * 0.3 + 0.1
    E = z[torch.randperm(n, device='cuda', generator=g)[:Kc]].contiguous()
    T = torch.randn(Kc, 32, device='cuda', generator=g)
    r1, r2 = run_fused(z, E, T, 1), run_fused(z, E, T, 2)
    for k in ('idx', 'loss', 'y', 'counts'):
        assert torch.equal(r1[k], r2[k]), f'{k} differs'


@pytest.mark.parametrize('D,n,Kc,squared', [(3, 100, 64, 1), (3, 5000, 512, 1), (4, 70001, 512, 1), (3, 9000, 300, 0), (3, 557056, 512, 1)])
def test_vq_fused_bwd_matches_autograd(D, n, Kc, squared):
    """Backward of the fused bottleneck (both kernels: per-warp reduction for small n, streaming + cluster reduction
    from 4096 vectors) against torch autograd in fp64 on  y = T[idx],  loss = (1 + beta) ||q - z||^2  (or the norm form):
    dz = J[idx]^T dy + g_loss dloss/dz,  dT[idx] += dy,  dE[idx] += g_loss dloss/dq."""
    from VQCPCB import _native, ops
    g = torch.Generator(device='cuda')
    g.manual_seed(n + D)
    zd, beta = 32, 0.25
    E = torch.randn(Kc, D, device='cuda', generator=g) * 2
    z = torch.randn(n, D, device='cuda', generator=g) * 2
    J = torch.randn(Kc, zd, D, device='cuda', generator=g)
    dy = torch.randn(n, zd, device='cuda', generator=g)
    gl = torch.rand(n, device='cuda', generator=g) / n
    idx = ((z[:, None, :] - E[None]) ** 2).sum(-1).argmin(1) if n * Kc < 5e7 else torch.randint(0, Kc, (n,), device='cuda', generator=g)
    dz = torch.full((n, D), float('nan'), device='cuda')
    dT = torch.zeros(Kc, zd, device='cuda')
    dE = torch.zeros(Kc, D, device='cuda')
    _native.call('vqcpc_vq_fused_bwd', ops._p(z), n, D, ops._p(E), ops._p(J), Kc, zd, ops._p(idx), beta, squared, ops._p(dy), ops._p(gl),
                 ops._p(dz), ops._p(dT), ops._p(dE), ops._stream())
    torch.cuda.synchronize()
    # fp64 reference
    zr, Er = z.double().requires_grad_(True), E.double().requires_grad_(True)
    q = Er[idx]
    if squared:
        loss = ((q.detach() - zr) ** 2).sum(-1) * beta + ((q - zr.detach()) ** 2).sum(-1)
    else:
        loss = ((q.detach() - zr) + 1e-5).norm(dim=-1) * beta + ((q - zr.detach()) + 1e-5).norm(dim=-1)
    (loss * gl.double()).sum().backward()
    dz_ref = zr.grad + torch.einsum('njd,nj->nd', J.double()[idx], dy.double())
    dT_ref = torch.zeros(Kc, zd, dtype=torch.float64, device='cuda').index_add_(0, idx, dy.double())
    def rel(a, b):
        return float((a.double() - b).norm() / b.norm().clamp_min(1e-30))
    assert rel(dz, dz_ref) < 1e-5, rel(dz, dz_ref)
    assert rel(dT, dT_ref) < 1e-5, rel(dT, dT_ref)
    assert rel(dE, Er.grad) < 1e-4, rel(dE, Er.grad)


def test_vq_tc_randomized_sweep():
    """Property sweep: 24 random problems (vector / codebook scales over six orders of magnitude, clustered and
    data-drawn codebooks, shifted means, D = 3 and 4, ragged n) - the tensor-core search and the direct scan must agree on
    every index and the usage histogram every time."""
    g = torch.Generator(device='cuda')
    g.manual_seed(2024)
    bad = []
    for trial in range(24):
        D = 3 if trial % 3 else 4
        Kc = [512, 300, 1024, 64][trial % 4]
        n = 100000 + 37 * trial
        zs = 10.0 ** ((trial % 7) - 3)
        es = 10.0 ** (((trial * 5) % 7) - 3)
        z = torch.randn(n, D, device='cuda', generator=g) * zs + (0.5 * zs if trial % 2 else 0.0)
        if trial % 5 == 0:
            E = z[torch.randperm(n, device='cuda', generator=g)[:Kc]].contiguous()            # data-dependent init
        elif trial % 5 == 1:
            centers = torch.randn(8, D, device='cuda', generator=g) * es
            E = (centers[torch.randint(0, 8, (Kc,), device='cuda', generator=g)] +
                 1e-3 * es * torch.randn(Kc, D, device='cuda', generator=g)).contiguous()      # tight clusters: many near-ties
        else:
            E = (torch.randn(Kc, D, device='cuda', generator=g) * es).contiguous()
        T = torch.randn(Kc, 32, device='cuda', generator=g)
        r1, r2 = run_fused(z, E, T, 1), run_fused(z, E, T, 2)
        if not (torch.equal(r1['idx'], r2['idx']) and torch.equal(r1['counts'], r2['counts']) and torch.equal(r1['y'], r2['y'])):
            bad.append((trial, D, Kc, zs, es, int((r1['idx'] != r2['idx']).sum())))
    assert not bad, bad
