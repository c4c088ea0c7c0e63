"""Stand-alone timing of the InfoNCE forward kernels over batch sizes (fixed cost vs streaming rate).
    python tools/nce_bench.py [--algo 1 2] [--batches 592 4096 16384]"""
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
    ap.add_argument('--batches', type=int, nargs='+', default=[592, 4096, 16384])
    ap.add_argument('--iters', type=int, default=30)
    a = ap.parse_args()
    N, KR, zd, cd = 15, 8, 32, 32
    p_, st = ops._p, ops._stream
    for B in a.batches:
        nsets = max(2, min(6, int(400e6 // (B * 16.6e3)) + 1))
        g = torch.Generator(device='cuda')
        g.manual_seed(0)
        W = torch.randn(zd, cd, KR, device='cuda', generator=g) * 0.25
        sets = [dict(c=torch.randn(B, cd, device='cuda', generator=g), yp=torch.randn(B, KR, zd, device='cuda', generator=g),
                     yn=torch.randn(B, N, KR, zd, device='cuda', generator=g), u=torch.empty(B, KR, zd, device='cuda'),
                     pos=torch.empty(B, KR, device='cuda'), neg=torch.empty(B, KR, N, device='cuda'), lse=torch.empty(B, KR, device='cuda'),
                     nll=torch.empty(B, KR, device='cuda'), cor=torch.empty(B, KR, dtype=torch.int32, device='cuda')) for _ in range(nsets)]
        for algo in a.algo:
            _native.call('vqcpc_infonce_set_algo', algo)

            def run(d):
                _native.call('vqcpc_infonce_fwd', p_(d['c']), p_(W), p_(d['yp']), p_(d['yn']), B, N, KR, zd, cd, p_(d['u']), p_(d['pos']),
                             p_(d['neg']), p_(d['lse']), p_(d['nll']), p_(d['cor']), st())
            for d in sets:
                run(d)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(a.iters):
                run(sets[i % nsets])
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) * 1e3 / a.iters
            byt = B * ((N + 1) * KR * zd * 4 + cd * 4 + KR * zd * 4 + KR * (N + 4) * 4)
            print(f'B {B:6d} algo {algo}: {us:8.2f} us  {byt / us / 1e3:7.1f} GB/s (in + out bytes)')
        _native.call('vqcpc_infonce_set_algo', 0)


if __name__ == '__main__':
    main()
