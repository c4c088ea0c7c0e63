"""Stand-alone timing of the fused quantizer forward (direct vs tensor-core search) at the BASELINE size.
    python tools/vq_bench.py [--algo 1 2] [--iters 30]"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'vqcpc-bach_b200')):
    sys.path.insert(0, p)
import torch  # noqa: E402
from VQCPCB import _native, ops  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--algo', type=int, nargs='+', default=[1, 2])
    ap.add_argument('--iters', type=int, default=30)
    ap.add_argument('--n', type=int, default=557056)
    ap.add_argument('--kc', type=int, default=512)
    ap.add_argument('--d', type=int, default=3)
    ap.add_argument('--prof', action='store_true')
    a = ap.parse_args()
    n, Kc, D, zd = a.n, a.kc, a.d, 32
    g = torch.Generator(device='cuda')
    g.manual_seed(0)
    E = torch.randn(Kc, D, device='cuda', generator=g) * 2
    T = torch.randn(Kc, zd, device='cuda', generator=g)
    sets = [dict(z=torch.randn(n, D, device='cuda', generator=g) * 2, idx=torch.empty(n, dtype=torch.int64, device='cuda'),
                 loss=torch.empty(n, device='cuda'), y=torch.empty(n, zd, device='cuda'),
                 counts=torch.zeros(Kc, dtype=torch.int32, device='cuda')) for _ in range(5)]
    p_, st = ops._p, ops._stream
    for algo in a.algo:
        chunk = 256
        if a.prof and algo == 2:
            stats = torch.zeros(4098, dtype=torch.int32, device='cuda')
            d = sets[0]
            for _ in range(2):
                stats.zero_()
                _native.call('vqcpc_vq_fused_fwd_ex', p_(d['z']), n, D, p_(E), p_(T), Kc, zd, 0.25, 1, p_(d['idx']), p_(d['loss']), None,
                             p_(d['y']), p_(d['counts']), 2, None, p_(stats), st())
            torch.cuda.synchronize()
            pr = stats[2:].view(-1, 8).cpu()
            pr = pr[pr[:, 0] > 0].float()
            names = ['setup', 'mma0_issued', 'w4_mask', 'w4_wait_z', 'w4_wait_acc', 'w4_phase1', 'w4_verify', 'w4_output']
            print(f'chunk {chunk}: exact groups/vector {stats[0].item() / n:.3f}, fallbacks {stats[1].item()}, CTAs {pr.shape[0]}')
            for i, nm in enumerate(names):
                print(f'   {nm:14s} median {pr[:, i].median().item():9.0f}  max {pr[:, i].max().item():9.0f}  (clk)')

        def run(d, with_y=True):
            _native.call('vqcpc_vq_fused_fwd_ex', p_(d['z']), n, D, p_(E), p_(T) if with_y else None, Kc, zd, 0.25, 1, p_(d['idx']),
                         p_(d['loss']), None, p_(d['y']) if with_y else None, p_(d['counts']), algo, None, None, st())
        for with_y in (True, False):
            for d in sets:
                run(d, with_y)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(a.iters):
                run(sets[i % 5], with_y)
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) * 1e3 / a.iters
            byt = n * (4 * D + 8 + 4 + (4 * zd if with_y else 0))
            print(f'algo {algo} chunk {chunk} with_y={with_y}: {us:8.2f} us  ({byt / us / 1e3:7.1f} GB/s algorithmic)')


if __name__ == '__main__':
    main()
