"""Stand-alone timing of the per-block relative attention kernels at the per-pass shapes of the headline config
(34 816 blocks per pass; L = 16 for stack 0, L = 4 for stack 1), against the HBM bytes they must move.
    python tools/attn_bench.py [--blocks 34816] [--only fwd0|fwd1|bwd0|bwd1] [--iters 10]
--only runs a single kernel twice (for ncu -k regex:attn)."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'vqcpc-bach_b200')):
    sys.path.insert(0, p)
import torch  # noqa: E402
from VQCPCB import ops  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--blocks', type=int, default=34816)
    ap.add_argument('--only', default=None)
    ap.add_argument('--iters', type=int, default=10)
    ap.add_argument('--L', type=int, nargs='+', default=[16, 4])
    a = ap.parse_args()
    H, hd = 8, 64
    d = H * hd
    for L in a.L:
        P = a.blocks
        M = P * L
        g = torch.Generator(device='cuda')
        g.manual_seed(0)
        nsets = 2                                        # 2 x (6 KB + 2 KB) x M >> L2
        qkv = [torch.randn(M, 3 * d, device='cuda', generator=g) * 0.7 for _ in range(nsets)]
        go = [torch.randn(M, d, device='cuda', generator=g) for _ in range(nsets)]
        e1 = torch.randn(H * L, hd, device='cuda', generator=g) * 0.5
        e2 = torch.randn(H * L, hd, device='cuda', generator=g) * 0.5
        de1, de2 = torch.zeros_like(e1), torch.zeros_like(e2)

        def timeit(fn, iters):
            fn(0)
            torch.cuda.synchronize()
            e0, e1_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(iters):
                fn(i % nsets)
            e1_.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1_) * 1e3 / iters

        res = {}
        probs = None
        for algo in (0, 1):
            if a.only and a.only != f'fwd{algo}':
                continue
            def fwd(i, algo=algo):
                return ops.attn_fwd(qkv[i], e1, e2, P, L, H, hd, need_probs=True, want_fp32=False, want_split=True, algo=algo)
            us = timeit(fwd, 2 if a.only else a.iters)
            byt = M * (3 * d * 4 + d * 4 + H * L * 4)
            res[f'fwd{algo}'] = (us, byt)
        for algo in (0, 1):
            if a.only and a.only != f'bwd{algo}':
                continue
            if probs is None:
                _, _, probs = ops.attn_fwd(qkv[0], e1, e2, P, L, H, hd, need_probs=True, want_fp32=False, want_split=True)

            def bwd(i, algo=algo):
                return ops.attn_bwd(qkv[i], e1, e2, probs, go[i], de1, de2, P, L, H, hd, 0.125, want_fp32=False, want_split=True,
                                    algo=algo)
            us = timeit(bwd, 2 if a.only else a.iters)
            byt = M * (3 * d * 4 + d * 4 + H * L * 4 + 3 * d * 4)
            res[f'bwd{algo}'] = (us, byt)
        for k, (us, byt) in res.items():
            print(f'L {L:2d} blocks {P} {k}: {us:9.1f} us  {byt / us / 1e3:7.1f} GB/s of algorithmic bytes ({byt / 1e9:.2f} GB)', flush=True)
        del qkv, go, probs
        torch.cuda.empty_cache()


if __name__ == '__main__':
    main()
