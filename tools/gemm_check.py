"""Quick A/B of two tcgen05 GEMM tile configurations on the forward shapes of the encoder (identical outputs expected - the
arithmetic per element is the same; compared through an integer fingerprint) + timing.  python tools/gemm_check.py [--cfg -1 5] [--rows 139264]"""
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
    ap.add_argument('--cfg', type=int, nargs='+', default=[-1, 5])
    ap.add_argument('--rows', type=int, default=139264)
    ap.add_argument('--iters', type=int, default=8)
    ap.add_argument('--rounds', type=int, default=5)
    ap.add_argument('--libs', nargs='+', default=[''], help="library tags to compare in one process: '' = libvqcpc_b200.so, "
                    "'_epi2' = libvqcpc_b200_epi2.so (VQ_EPI_KIND=2 VQ_TAG=_epi2 tools/build.sh) ...")
    a = ap.parse_args()
    from VQCPCB import _native
    handles = {}
    base = _native.LIB_PATH
    for tag in a.libs:
        _native._lib = None
        _native.LIB_PATH = base.replace('.so', tag + '.so')
        handles[tag] = _native.load()
        print(f'lib {tag!r}: epilogue kind {handles[tag].vqcpc_gemm_epilogue_kind()}')

    class _Ops:                   # ops.set_epilogue(tag) selects which loaded library the next calls go to
        pass
    ops.set_epilogue = lambda tag: setattr(_native, '_lib', handles[tag])
    a.epi = a.libs
    variants = [(c, e) for c in a.cfg for e in a.epi]
    M = a.rows
    g = torch.Generator(device='cuda')
    g.manual_seed(0)
    cases = [('out  K512 N512  bias+residual', 512, 512, dict(residual=True)),
             ('qkv  K512 N1536 bias+scale', 1536, 512, dict(scale=True)),
             ('ffn1 K512 N2048 bias+relu->split', 2048, 512, dict(relu=True, split=True)),
             ('ffn2 K2048 N512 bias+residual', 512, 2048, dict(residual=True)),
             ('dgrad ffn2 K512 N2048 bf16mask->split', 2048, 512, dict(dgrad=True, mask=True, split=True)),
             ('dgrad out  K512 N512', 512, 512, dict(dgrad=True)),
             ('dgrad qkv  K1536 N512 residual', 512, 1536, dict(dgrad=True, residual=True)),
             ('dgrad ffn1 K2048 N512 residual', 512, 2048, dict(dgrad=True, residual=True))]
    for name, N, K, o in cases:
        x = torch.randn(M, K, device='cuda', generator=g)
        dg = bool(o.get('dgrad'))
        w = (torch.randn(K, N, device='cuda', generator=g) if dg else torch.randn(N, K, device='cuda', generator=g)) * 0.05
        b = None if dg else torch.randn(N, device='cuda', generator=g)
        mask = torch.randn(M, N, device='cuda', generator=g).to(torch.bfloat16) if o.get('mask') else None
        res = torch.randn(M, N, device='cuda', generator=g) if o.get('residual') else None
        xs, ws = ops.split_bf16(x), ops.split_bf16(w)
        outs, times = {}, {}
        def run():
            return ops.gemm(None, None, M, N, K, b_kmajor=not dg, relu_mask=mask, bias=b, residual=res, relu=bool(o.get('relu')),
                            scale_cols=(0.125, N // 3) if o.get('scale') else None, a_split=xs, b_split=ws,
                            want_c=not o.get('split'), out_split=bool(o.get('split')))
        for cfg, epi in variants:
            ops.set_epilogue(epi)
            ops.set_tile_config(cfg)
            r = run()
            torch.cuda.synchronize()
            res_t = r[1] if o.get('split') else (r,)
            # keep only a fingerprint: holding the outputs would make every timed call cudaMalloc a fresh block
            outs[cfg, epi] = tuple(int(t.view(torch.int16).to(torch.int64).sum().item()) for t in res_t)
            del r, res_t
            for _ in range(3):                       # warm-up: the caching allocator now owns reusable output blocks
                run()
            torch.cuda.synchronize()
        # the variants are timed interleaved over several rounds and the best round counts: SM clocks move with the power cap, so
        # back-to-back blocks of one variant each are not comparable
        for rnd in range(a.rounds):
            for cfg, epi in variants:
                ops.set_epilogue(epi)
                ops.set_tile_config(cfg)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(a.iters):
                    run()
                e1.record()
                torch.cuda.synchronize()
                t = e0.elapsed_time(e1) * 1e3 / a.iters
                times[cfg, epi] = min(times.get((cfg, epi), 1e30), t)
        for tag in a.libs:
            ops.set_epilogue(tag)
            ops.set_tile_config(-1)
        ops.set_epilogue(a.libs[0])
        same = all(outs[v] == outs[variants[0]] for v in variants[1:])
        fl = 2.0 * M * N * K * 3
        print(f'{name}: ' + '  '.join(f'cfg {c} lib{e or '_default'}: {times[c, e]:8.1f} us {fl / times[c, e] / 1e6:7.1f} tensor-TFLOP/s' for c, e in variants) +
              f'  identical={same}', flush=True)


if __name__ == '__main__':
    main()
