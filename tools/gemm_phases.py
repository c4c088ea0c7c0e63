"""Phase timers of the CTA-pair GEMM kernel (vqcpc_gemm_set_debug) on the encoder shapes: where the MMA thread and the epilogue
warps wait.  python tools/gemm_phases.py [--rows 139264] [--cfg 4 5]"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'vqcpc-bach_b200')):
    sys.path.insert(0, p)
import torch  # noqa: E402
from VQCPCB import ops, _native  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--rows', type=int, default=139264)
    ap.add_argument('--cfg', type=int, nargs='+', default=[4])
    ap.add_argument('--exp', type=int, default=0, help='epilogue timing experiment (results invalid)')
    ap.add_argument('--lib', default='', help="library tag, e.g. _epi0 for libvqcpc_b200_epi0.so")
    a = ap.parse_args()
    _native.LIB_PATH = _native.LIB_PATH.replace('.so', a.lib + '.so')
    print('epilogue kind', _native.load().vqcpc_gemm_epilogue_kind(), 'experiment', a.exp)
    _native.call('vqcpc_gemm_set_experiment', a.exp)
    M = a.rows
    g = torch.Generator(device='cuda')
    g.manual_seed(0)
    cases = [('out  K512 N512  bias+residual', 512, 512, dict(residual=True)),
             ('qkv  K512 N1536 bias+scale', 1536, 512, dict(scale=True)),
             ('ffn1 K512 N2048 bias+relu->split', 2048, 512, dict(relu=True, split=True)),
             ('ffn2 K2048 N512 bias+residual', 512, 2048, dict(residual=True)),
             ('plain K512 N512', 512, 512, dict(nobias=True))]
    dbg = torch.zeros(12, dtype=torch.int64, device='cuda')
    for name, N, K, o in cases:
        x = torch.randn(M, K, device='cuda', generator=g)
        w = torch.randn(N, K, device='cuda', generator=g) * 0.05
        b = None if o.get('nobias') else torch.randn(N, device='cuda', generator=g)
        res = torch.randn(M, N, device='cuda', generator=g) if o.get('residual') else None
        xs, ws = ops.split_bf16(x), ops.split_bf16(w)

        def run():
            return ops.gemm(None, None, M, N, K, bias=b, residual=res, relu=bool(o.get('relu')),
                            scale_cols=(0.125, N // 3) if o.get('scale') else None, a_split=xs, b_split=ws,
                            want_c=not o.get('split'), out_split=bool(o.get('split')))
        for cfg in a.cfg:
            ops.set_tile_config(cfg)
            for _ in range(3):
                run()
            torch.cuda.synchronize()
            dbg.zero_()
            _native.call('vqcpc_gemm_set_debug', dbg.data_ptr())
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            run()
            e1.record()
            torch.cuda.synchronize()
            _native.call('vqcpc_gemm_set_debug', None)
            d = dbg.tolist()
            tiles = (M // 256) * (N // 256)
            mma = d[0] / 74.0
            print(f'{name} cfg {cfg}: {e0.elapsed_time(e1) * 1e3:7.1f} us; per cluster: MMA loop {mma:9.0f} cyc ({mma / (tiles / 74.0):6.0f} / tile), '
                  f'waiting accumulator {100 * d[1] / max(d[0], 1):5.1f} %, operands {100 * d[2] / max(d[0], 1):5.1f} %; '
                  f'tcgen05.ld->data (q0) {d[3] / tiles / 2e3 / (2 if cfg == 5 else 1):.1f} kcyc per tile; '
                  f'epilogue warps (per quadrant, both CTAs) wait/drain kcyc per tile: ' +
                  ' '.join(f'{d[4 + q] / tiles / 1e3 / (len([1]) * 2 * (2 if cfg == 5 else 1)):.1f}/{d[8 + q] / tiles / 1e3 / (2 * (2 if cfg == 5 else 1)):.1f}' for q in range(4)),
                  flush=True)
        ops.set_tile_config(-1)


if __name__ == '__main__':
    main()
