"""Per-shape throughput of the tcgen05 GEMM on the encoder's own shapes (run on the GPU box).

    python tools/gemm_bench.py [--rows 278528] [--cfgs 1,2]

Reports algorithmic TFLOP/s (2MNK) and the tensor-pipe rate (x3 in bf16x3) for every GEMM of one stack-0 layer:
forward (NT, fused epilogues), data gradients (NN) and weight gradients (TN, split-K)."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'vqcpc-bach_b200'))
import torch  # noqa: E402
from VQCPCB import ops  # noqa: E402


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--rows', type=int, default=278528)
    ap.add_argument('--cfgs', default='1,2')
    ap.add_argument('--mode', default='bf16x3')
    ap.add_argument('--only', type=int, default=-1, help='run a single case index (for ncu)')
    ap.add_argument('--iters', type=int, default=10)
    args = ap.parse_args()
    ops.set_gemm_mode(args.mode)
    M, d, F = args.rows, 512, 2048
    dev = 'cuda'
    x = torch.randn(M, d, device=dev)
    xs = ops.split_bf16(x)
    h = torch.randn(M, F, device=dev).relu_()
    hs = ops.split_bf16(h)
    g3 = ops.split_bf16(torch.randn(M, 3 * d, device=dev))
    wqkv, bqkv = torch.randn(3 * d, d, device=dev) * 0.05, torch.randn(3 * d, device=dev)
    wo, bo = torch.randn(d, d, device=dev) * 0.05, torch.randn(d, device=dev)
    w1, b1 = torch.randn(F, d, device=dev) * 0.05, torch.randn(F, device=dev)
    w2, b2 = torch.randn(d, F, device=dev) * 0.05, torch.randn(d, device=dev)
    gw = torch.zeros(F, d, device=dev)
    gwq = torch.zeros(3 * d, d, device=dev)
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    cases = [
        ('fwd qkv   NT bias+scale      ', 3 * d, d, lambda: ops.gemm(None, wqkv, M, 3 * d, d, bias=bqkv, scale_cols=(0.125, d), b_is_weight=True, a_split=xs)),
        ('fwd out   NT bias+residual   ', d, d, lambda: ops.gemm(None, wo, M, d, d, bias=bo, residual=x, b_is_weight=True, a_split=xs)),
        ('fwd ffn1  NT bias+relu->split', F, d, lambda: ops.gemm(None, w1, M, F, d, bias=b1, relu=True, b_is_weight=True, a_split=xs, want_c=False, out_split=True)),
        ('fwd ffn2  NT bias+residual   ', d, F, lambda: ops.gemm(None, w2, M, d, F, bias=b2, residual=x, b_is_weight=True, a_split=hs)),
        ('dgrad ffn2 NN mask->split    ', F, d, lambda: ops.gemm(None, w2, M, F, d, True, False, relu_mask=hs[0], b_is_weight=True, a_split=xs, want_c=False, out_split=True)),
        ('dgrad ffn1 NN +residual      ', d, F, lambda: ops.gemm(None, w1, M, d, F, True, False, residual=x, b_is_weight=True, a_split=hs)),
        ('dgrad qkv  NN +residual      ', d, 3 * d, lambda: ops.gemm(None, wqkv, M, d, 3 * d, True, False, residual=x, b_is_weight=True, a_split=g3)),
        ('wgrad ffn  TN splitK atomics ', None, None, lambda: ops.gemm(None, None, F, d, M, False, False, out=gw, accumulate=True, ksplit=ops._wgrad_ksplit(M, 16 * 4), a_split=hs, b_split=xs)),
        ('wgrad qkv  TN splitK atomics ', None, None, lambda: ops.gemm(None, None, 3 * d, d, M, False, False, out=gwq, accumulate=True, ksplit=ops._wgrad_ksplit(M, 12 * 4), a_split=g3, b_split=xs)),
    ]
    flops = [2.0 * M * 3 * d * d, 2.0 * M * d * d, 2.0 * M * F * d, 2.0 * M * d * F, 2.0 * M * F * d, 2.0 * M * d * F,
             2.0 * M * d * 3 * d, 2.0 * M * F * d, 2.0 * M * 3 * d * d]
    passes = 3 if args.mode == 'bf16x3' else 1
    for cfg in [int(c) for c in args.cfgs.split(',')]:
        ops.set_tile_config(cfg)
        print(f'--- tile config {cfg} mode {args.mode} rows {M} ---')
        tot_t = tot_f = 0.0
        for ci, ((name, _, _, fn), fl) in enumerate(zip(cases, flops)):
            if args.only >= 0 and ci != args.only:
                continue
            ms = timeit(fn, args.iters)
            tot_t += ms
            tot_f += fl
            print(f'{name} {ms:8.3f} ms  {fl / ms / 1e9:8.1f} TFLOP/s alg  {passes * fl / ms / 1e9:8.1f} tensor')
        print(f'{"sum":30s} {tot_t:8.3f} ms  {tot_f / tot_t / 1e9:8.1f} TFLOP/s alg  {passes * tot_f / tot_t / 1e9:8.1f} tensor')
    ops.set_tile_config(-1)


if __name__ == '__main__':
    main()
