"""Quick single-process check of the tensor-core attention kernels against the CUDA-core ones (outputs, bf16 splits,
probabilities, all gradients; dead-row elimination, dropout masks, multi-tile grids) + their timing at the headline
per-pass shape.  ~6 s on a B200: python tools/attn_check.py"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'vqcpc-bach_b200')):
    sys.path.insert(0, p)
import torch
from VQCPCB import ops
t0 = time.time()
H, hd = 8, 64
d = H * hd
def rel(a, b):
    a, b = a.double(), b.double()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()
ok = True
for (L, P, keep) in [(16, 37, 1), (16, 9, 4), (4, 21, 4), (8, 11, 2), (16, 2501, 4), (16, 3001, 1)]:
    torch.manual_seed(L + P)
    qkv = torch.randn(P * L, 3 * d, device='cuda') * 0.7
    e1 = torch.randn(H * L, hd, device='cuda') * 0.5
    e2 = torch.randn(H * L, hd, device='cuda') * 0.5
    mask = (torch.rand(P, H, L, L, device='cuda') > 0.3).to(torch.uint8)
    for drop in (None, (mask, 1 / 0.7)):
        r = {}
        for algo in (0, 1):
            o, sp, pr = ops.attn_fwd(qkv, e1, e2, P, L, H, hd, keep=keep, want_split=True, drop=drop, algo=algo)
            r[algo] = (o, sp[0].float() + sp[1].float(), pr[:, :, ::keep])
        e = max(rel(a, b) for a, b in zip(r[0], r[1]))
        go = torch.randn(P * (L // keep), d, device='cuda')
        pr1 = ops.attn_fwd(qkv, e1, e2, P, L, H, hd, keep=keep, drop=drop, algo=1)[1]
        rb = {}
        for algo in (0, 1):
            de1, de2 = torch.zeros_like(e1), torch.zeros_like(e2)
            dq, sp = ops.attn_bwd(qkv, e1, e2, pr1, go, de1, de2, P, L, H, hd, 0.125, keep=keep, want_split=True, drop=drop, algo=algo)
            rb[algo] = (dq, sp[0].float() + sp[1].float(), de1, de2)
        eb = max(rel(a, b) for a, b in zip(rb[0], rb[1]))
        good = e < 3e-6 and eb < 3e-6
        ok &= good
        print(f'L {L} P {P} keep {keep} drop {drop is not None}: fwd {e:.2e} bwd {eb:.2e} {"OK" if good else "FAIL"}', flush=True)
print('ALL OK' if ok else 'FAILED', f'{time.time() - t0:.1f}s', flush=True)
L, P = 16, 34816
g = torch.Generator(device='cuda'); g.manual_seed(0)
qkv = [torch.randn(P * L, 3 * d, device='cuda', generator=g) * 0.7 for _ in range(2)]
go = [torch.randn(P * L, d, device='cuda', generator=g) for _ in range(2)]
e1 = torch.randn(H * L, hd, device='cuda') * 0.5
e2 = torch.randn(H * L, hd, device='cuda') * 0.5
de1, de2 = torch.zeros_like(e1), torch.zeros_like(e2)
_, _, probs = ops.attn_fwd(qkv[0], e1, e2, P, L, H, hd, want_fp32=False, want_split=True)
def timeit(fn, iters=6):
    fn(0); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(iters): fn(i % 2)
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e3 / iters
print('fwd0 us', timeit(lambda i: ops.attn_fwd(qkv[i], e1, e2, P, L, H, hd, want_fp32=False, want_split=True)), flush=True)
print('bwd0 us', timeit(lambda i: ops.attn_bwd(qkv[i], e1, e2, probs, go[i], de1, de2, P, L, H, hd, 0.125, want_fp32=False, want_split=True)), flush=True)
