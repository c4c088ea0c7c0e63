"""Benchmark of the VQ-CPC encoder-training hot path (BASELINE.json metric: encoder training samples/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B] [--mode bf16x3|bf16|fp32]

One "step" = one full training step (3 encoder passes, quantizer, upscaler, GRU context, InfoNCE, backward, global-norm
clip + Adam) on one synthetic batch of B windows per GPU.  N=1 workload = BASELINE.json configs[1]
("K=512 codebook, synthetic 4-voice seq 256, batch 4096, 1xB200"); N>1 = configs[2] (same per GPU, weak scaling, one
NCCL all-reduce of [grads | counts | stats] per step).  Prints ONE JSON line on rank 0.

  value     samples/s, inputs already resident in HBM, timed with CUDA events, max over ranks
  e2e       same metric through the public call (trainer.step on a HOST batch dict): pinned-host -> device copy of the
            step's tokens and a device -> host read of the step's loss inside the timed region
  roofline  dominant kernel (the tcgen05 GEMM): algorithmic FLOPs (2MNK per launch) / CUDA-event time of those
            launches inside the timed region, against MEASURED_PEAKS.json (sustained bf16)
  cpu_baseline  the CPU oracle port of the reference trainer step timed on the host cores (bounded sample, rank 0, N=1)

--impl reference times that same CPU port (the reference is pure Python and cannot travel to the GPU box; the oracle is
its restatement pinned by tests/golden) with all host threads, as samples/s.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, 'vqcpc-bach_b200')):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

_OUT = sys.stdout
METRIC = 'vqcpc_encoder_training_samples_per_sec'
UNIT = 'samples/s'
KL = KR = 8
NNEG = 15
CODEBOOK = 512
# bounded CPU sample of the reference arm / cpu_baseline: ~10 s of work on the GPU box's host cores (CPU throughput is
# batch independent: 3.3 / 3.5 samples/s at B = 16 / 32 on 8 cores)
CPU_SAMPLE_BATCH = 32
FLOPS_PER_SAMPLE = 103.8e9          # fwd+bwd algorithmic (SURVEY.md §8a FLOP accounting), no recompute / dead-row skipping


def workload_name(B):
    return f'encoder_vqcpc K={CODEBOOK} D=3 seq256 (KL=KR=8) N=15 batch {B}/GPU, transformer downscaler [4,4]/[2,2]/512/8/2048, dropout 0'


def peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        d = json.load(open(path))
        return dict(tflops=d.get('bf16_tflops_sustained', 1386.1), tflops_burst=d.get('bf16_tflops', 1645.3),
                    hbm=d.get('hbm_gbs', 6547.2), src='MEASURED_PEAKS.json (sustained bf16)')
    return dict(tflops=1400.0, tflops_burst=1590.0, hbm=6650.0, src='fallback (B200_PROFILING.md)')


# ------------------------------------------------------------------------------------------------ clocks sampler
class ClockSampler(threading.Thread):
    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False

    def run(self):
        q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
             'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')
        while not self.stop_flag:
            try:
                out = subprocess.run(['nvidia-smi', f'--id={self.index}', f'--query-gpu={q}', '--format=csv,noheader,nounits'],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([s.strip() for s in out.split(',')])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        if not self.samples:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=['unavailable'])
        mhz = sorted(int(s[0]) for s in self.samples if s[0].isdigit())
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = [n for i, n in enumerate(names) if any(s[2 + i].lower().startswith('active') for s in self.samples if len(s) > 2 + i)]
        mx = [int(s[1]) for s in self.samples if s[1].isdigit()]
        return dict(sm_mhz=mhz[len(mhz) // 2] if mhz else None, sm_max_mhz=max(mx) if mx else None, reasons=reasons,
                    samples=len(self.samples))


# ------------------------------------------------------------------------------------------------ CPU oracle leg
def use_all_host_threads():
    """torch.distributed.run exports OMP_NUM_THREADS=1 to every worker; the CPU legs must use the box's cores."""
    n = os.cpu_count() or 1
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        pass
    torch.set_num_threads(max(1, n))
    return torch.get_num_threads()


def cpu_port_samples_per_sec(B_cpu=CPU_SAMPLE_BATCH, steps=2, warmup=1, dropout=0.0):
    """The reference trainer step as restated by the oracle, on the host cores, at the benchmark shapes with a
    CPU-sized batch (one fp32 activation at B=4096 would be 18 GB; CPU throughput is ~batch independent).
    dropout > 0: fresh F.dropout draws at the reference's sites (the shipped training setting, p = 0.1)."""
    from oracle import params as P, vqcpc_oracle as O
    cores = use_all_host_threads()
    cfg = O.default_cfg(codebook_size=CODEBOOK, KL=KL, KR=KR, N=NNEG)
    if dropout > 0:
        cfg['dropout'] = float(dropout)
    prm = P.make_params(cfg, 0)
    m = v = None
    t_total = 0.0
    for i in range(warmup + steps):
        batch = P.make_batch(cfg, B_cpu, seed=100 + i)
        t0 = time.perf_counter()
        out, grads, gnorm, prm, m, v = O.trainer_step(batch, prm, cfg, m, v, step=i + 1, lr=1e-4)
        dt = time.perf_counter() - t0
        if i >= warmup:
            t_total += dt
    return B_cpu * steps / t_total, dict(cores=cores, kind='port', dropout=dropout,
                                         sample=f'{steps} steps of B={B_cpu} at the benchmark shapes (seq 256, K={CODEBOOK}, N=15), '
                                                f'{warmup} warm-up, torch {torch.__version__} CPU fp32 on {cores} threads, '
                                                f'oracle/vqcpc_oracle.trainer_step')


def run_reference(args):
    """--impl reference: the CPU port of the reference step with all host threads; honours --steps / --warmup (every step
    is a bounded B=32 sample of the workload, ~2-3 s on 16 cores).  Under torchrun only rank 0 works."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    val, info = cpu_port_samples_per_sec(B_cpu=CPU_SAMPLE_BATCH, steps=steps, warmup=warmup)
    line = dict(metric=METRIC, value=val, unit=UNIT, impl='reference', n_gpus=args.gpus, steps=steps, warmup=warmup,
                ms_per_step=1000.0 * CPU_SAMPLE_BATCH / val, higher_is_better=True, scaling='weak', vs_baseline=None, dtype='f32',
                data='synthetic', config=dict(workload=workload_name(args.batch), note=f'CPU port of the reference step; B={CPU_SAMPLE_BATCH} sample per step'),
                cpu_baseline=dict(value=val, unit=UNIT, **info),
                e2e=dict(value=val, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
    if not args.no_dropout_line:
        dval, dinfo = cpu_port_samples_per_sec(B_cpu=CPU_SAMPLE_BATCH, steps=min(steps, 3), warmup=1, dropout=0.1)
        line['with_dropout'] = dict(value=dval, unit=UNIT, **dinfo)
    print(json.dumps(line), file=_OUT, flush=True)


# ------------------------------------------------------------------------------------------------ GPU leg
def build_trainer(B, seed):
    from VQCPCB.getters import get_dataloader_generator, get_encoder, get_encoder_trainer
    import importlib.util
    os.environ['VQCPC_BATCH'] = str(B)
    os.environ['VQCPC_BLOCKS'] = str(KL)
    os.environ['VQCPC_CODEBOOK'] = str(CODEBOOK)
    path = os.path.join(ROOT, 'vqcpc-bach_b200', 'VQCPCB', 'configs', 'encoder_vqcpc.py')
    spec = importlib.util.spec_from_file_location('bench_cfg', path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    config = mod.config
    config['quantizer_kwargs']['initialize'] = True            # data-dependent codebook init, as in a fresh run
    config['dataloader_generator_kwargs']['seed'] = seed
    torch.manual_seed(0)                                        # identical random-init weights on every rank
    gen = get_dataloader_generator(config['dataset'], config['training_method'], config['dataloader_generator_kwargs'])
    enc = get_encoder('/tmp/vqcpc_bench', gen, config)
    tr = get_encoder_trainer('/tmp/vqcpc_bench', gen, config['training_method'], enc, config['auxiliary_networks_kwargs'])
    tr.to('cuda')
    tr.init_optimizers(config['lr'], config['schedule_lr'])
    tr.train()
    return tr, gen, config


def bench_quantizer_infonce(B, iters=60, nsets=5):
    """Stand-alone roofline measurement of the fused quantizer + InfoNCE path at the full batch (BASELINE metric (ii)):
    one vqcpc_vq_fused_fwd over all P = B*(KL+KR+N*KR) vectors (search + loss + table upscale) followed by one
    vqcpc_infonce_fwd, back to back over rotating buffer sets larger than L2; and the matching backward pair.
    Algorithmic bytes: SURVEY.md §8d (36 704 B/sample forward)."""
    import ctypes
    from VQCPCB import _native, ops
    dev = 'cuda'
    D, zd, cd, Kc = 3, 32, 32, CODEBOOK
    P = B * (KL + KR + NNEG * KR)
    n_neg = B * NNEG * KR
    g = torch.Generator(device=dev)
    g.manual_seed(0)
    E = torch.randn(Kc, D, device=dev, generator=g) * 2
    T = torch.randn(Kc, zd, device=dev, generator=g)
    J = torch.randn(Kc, zd, D, device=dev, generator=g)
    W = torch.randn(zd, cd, KR, device=dev, generator=g) * 0.25
    sets = []
    for _ in range(nsets):
        d = dict(z=torch.randn(P, D, device=dev, generator=g) * 2, c=torch.randn(B, cd, device=dev, generator=g),
                 idx=torch.empty(P, dtype=torch.int64, device=dev), loss=torch.empty(P, device=dev),
                 y=torch.empty(P, zd, device=dev), u=torch.empty(B, KR, zd, device=dev), pos=torch.empty(B, KR, device=dev),
                 neg=torch.empty(B, KR, NNEG, device=dev), lse=torch.empty(B, KR, device=dev), nll=torch.empty(B, KR, device=dev),
                 correct=torch.empty(B, KR, dtype=torch.int32, device=dev), counts=torch.zeros(Kc, dtype=torch.int32, device=dev),
                 dy=torch.empty(P, zd, device=dev), dz=torch.empty(P, D, device=dev), dc=torch.zeros(B, cd, device=dev),
                 dW=torch.zeros_like(W), dT=torch.zeros(Kc, zd, device=dev), dE=torch.zeros(Kc, D, device=dev),
                 gl=torch.full((P,), 1.0 / (3 * B), device=dev))
        sets.append(d)
    p_, st = ops._p, ops._stream

    def vq_fwd(d):
        _native.call('vqcpc_vq_fused_fwd', p_(d['z']), P, D, p_(E), p_(T), Kc, zd, 0.25, 1, p_(d['idx']), p_(d['loss']), None,
                     p_(d['y']), p_(d['counts']), st())

    def nce_fwd(d):
        y = d['y']
        ypos = y[n_neg + B * KL:]                                   # right blocks
        _native.call('vqcpc_infonce_fwd', p_(d['c']), p_(W), p_(ypos), p_(y), B, NNEG, KR, zd, cd, p_(d['u']), p_(d['pos']),
                     p_(d['neg']), p_(d['lse']), p_(d['nll']), p_(d['correct']), st())

    def nce_bwd(d):
        y, dy = d['y'], d['dy']
        _native.call('vqcpc_infonce_bwd', p_(d['c']), p_(W), p_(y[n_neg + B * KL:]), p_(y), p_(d['u']), p_(d['pos']), p_(d['neg']),
                     p_(d['lse']), None, 1.0 / B, B, NNEG, KR, zd, cd, p_(d['dc']), p_(d['dW']), p_(dy[n_neg + B * KL:]), p_(dy),
                     st())

    def vq_bwd(d):
        _native.call('vqcpc_vq_fused_bwd', p_(d['z']), P, D, p_(E), p_(J), Kc, zd, p_(d['idx']), 0.25, 1, p_(d['dy']), p_(d['gl']),
                     p_(d['dz']), p_(d['dT']), p_(d['dE']), st())

    # ---- index-gather form (csrc/idxpath.cu): what the trainer runs in table mode
    n_left = B * KL
    for d in sets:
        d.update(y_left=torch.empty(n_left, zd, device=dev), dy_left=torch.randn(n_left, zd, device=dev, generator=g) * 1e-3,
                 coef_pos=torch.empty(B, KR, device=dev), coef_neg=torch.empty(B, NNEG, KR, device=dev))
    part = torch.empty(_native.load().vqcpc_infonce_idx_scratch(B, KR, Kc), device=dev)
    seg_row0 = (ctypes.c_int64 * 3)(0, n_neg, n_neg + n_left)
    seg_rows = (ctypes.c_int64 * 3)(n_neg, n_left, B * KR)
    seg_kind = (ctypes.c_int * 3)(2, 1, 2)
    seg_per_b = (ctypes.c_int * 3)(NNEG * KR, 1, KR)

    def vq_fwd_idx(d):
        _native.call('vqcpc_vq_fused_fwd_rows', p_(d['z']), P, D, p_(E), p_(T), Kc, zd, 0.25, 1, p_(d['idx']), p_(d['loss']), None,
                     p_(d['y_left']), n_neg, n_left, p_(d['counts']), st())

    def nce_fwd_idx(d):
        idx = d['idx']
        _native.call('vqcpc_infonce_idx_fwd', p_(d['c']), p_(W), p_(T), p_(idx[n_neg + n_left:]), p_(idx), B, NNEG, KR, zd, cd, Kc,
                     p_(d['u']), p_(d['pos']), p_(d['neg']), p_(d['lse']), p_(d['nll']), p_(d['correct']), st())

    def nce_bwd_idx(d):
        idx = d['idx']
        _native.call('vqcpc_infonce_idx_bwd', p_(d['c']), p_(W), p_(T), p_(idx[n_neg + n_left:]), p_(idx), p_(d['u']), p_(d['pos']),
                     p_(d['neg']), p_(d['lse']), None, 1.0 / B, B, NNEG, KR, zd, cd, Kc, p_(d['dc']), p_(d['coef_pos']),
                     p_(d['coef_neg']), p_(part), p_(d['dW']), p_(d['dT']), st())

    def vq_bwd_idx(d):
        pa = (ctypes.c_void_p * 3)(d['coef_neg'].data_ptr(), d['dy_left'].data_ptr(), d['coef_pos'].data_ptr())
        pb = (ctypes.c_void_p * 3)(d['u'].data_ptr(), 0, d['u'].data_ptr())
        _native.call('vqcpc_vq_idx_bwd', p_(d['z']), P, D, p_(E), p_(J), Kc, zd, p_(d['idx']), 0.25, 1, p_(d['gl']), 3, seg_row0,
                     seg_rows, seg_kind, pa, pb, seg_per_b, KR, p_(d['dz']), p_(d['dE']), p_(d['dT']), st())

    def fwd(d):
        vq_fwd_idx(d)
        nce_fwd_idx(d)

    def bwd(d):
        nce_bwd_idx(d)
        vq_bwd_idx(d)

    def fwd_pv(d):
        vq_fwd(d)
        nce_fwd(d)

    def bwd_pv(d):
        nce_bwd(d)
        vq_bwd(d)

    def timeit(fn, graph=True):
        """Mean device time of fn over the rotating buffer sets (> L2).  graph=True replays a CUDA graph of the launch
        sequence so that host launch latency (13 us per ctypes call) does not pad kernels that take 5-50 us."""
        for d in sets:
            fn(d)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if graph:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for d in sets:
                    fn(d)
                g_ = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g_, stream=side):
                    for _ in range(2):
                        for d in sets:
                            fn(d)
            torch.cuda.current_stream().wait_stream(side)
            per = 2 * nsets
            reps = max(1, iters // per)
            g_.replay()
            torch.cuda.synchronize()
            a.record()
            for _ in range(reps):
                g_.replay()
            b.record()
            torch.cuda.synchronize()
            return a.elapsed_time(b) * 1e3 / (reps * per)
        a.record()
        for i in range(iters):
            fn(sets[i % nsets])
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) * 1e3 / iters      # us

    for d in sets:
        d['dy'].normal_()
    us_f = timeit(fwd)
    us_b = timeit(bwd)
    us_f_launch = timeit(fwd, graph=False)
    per_kernel = dict(vq_fused_fwd_rows_us=timeit(vq_fwd_idx), infonce_idx_fwd_us=timeit(nce_fwd_idx),
                      infonce_idx_bwd_us=timeit(nce_bwd_idx), vq_idx_bwd_us=timeit(vq_bwd_idx))
    us_f_pv, us_b_pv = timeit(fwd_pv), timeit(bwd_pv)
    per_kernel_pv = dict(vq_fused_fwd_us=timeit(vq_fwd), infonce_fwd_us=timeit(nce_fwd), infonce_bwd_us=timeit(nce_bwd),
                         vq_fused_bwd_us=timeit(vq_bwd))
    pk = peaks()
    vec = KL + KR + NNEG * KR
    bytes_f = B * (vec * (4 * D + 8 + 4 + 4 * zd) + (KR + NNEG * KR) * 4 * zd + 4 * cd + 8 * KR)
    # backward: read dy-producing inputs (y for right+neg, u, scores) and write dy for right+neg; then per vector read
    # dy (4 zd) + idx (8) + z (4 D) + g_loss (4), write dz (4 D)
    bytes_b = B * ((KR + NNEG * KR) * 8 * zd + vec * (4 * zd + 8 + 8 * D + 4))
    # bytes the index form really moves (indices, z, loss, y of the left blocks, c, u, scores / coefficients, dz)
    real_f = B * (vec * (4 * D + 8 + 4) + KL * 4 * zd + 4 * cd + (KR + NNEG * KR) * 8 + KR * (4 * zd + 4 * (NNEG + 4)))
    real_b = B * (KR * (4 * zd + 4 * (NNEG + 2) + 4 * (NNEG + 1)) + (KR + NNEG * KR) * 8 + 8 * cd + KL * (4 * zd + 8) + vec * (8 * D + 8 + 4 + 4))
    search_instr = P * Kc * 11.0                                   # exact direct-form search: 11 fp32 instr per (vector, code)
    alu_us = search_instr / (148 * 128 * 1.9e9) * 1e6
    traffic = None
    for tname in ('r02_path_traffic.json', 'r01_path_traffic.json'):
        tpath = os.path.join(ROOT, 'profiles', tname)
        if os.path.exists(tpath):
            traffic = json.load(open(tpath))           # dram bytes per launch from the committed ncu --set full captures
            traffic['source'] = 'profiles/' + tname
            break
    return dict(bound='hbm', unit='GB/s', peak=pk['hbm'], traffic=traffic, peak_source=pk['src'].replace('sustained bf16', 'hbm copy'),
                kernels='index-gather form: vq_tc_fwd_kernel (tcgen05 score + exact verify; y for the left blocks only) + '
                        'nce_idx_fwd_kernel (forward); nce_idx_bwd_tc_kernel (tcgen05 dT = C^T U) + reduce_partials_kernel + '
                        'vq_idx_bwd_kernel (backward)',
                timing='CUDA-graph replay of the launch sequence over %d rotating buffer sets (> L2); us_fwd_eager = same '
                       'launches issued one by one from Python' % nsets, us_fwd_eager=us_f_launch,
                batch=B, vectors=P, algorithmic_bytes_fwd=bytes_f, us_fwd=us_f, achieved=bytes_f / us_f / 1e3,
                frac=bytes_f / us_f / 1e3 / pk['hbm'], algorithmic_bytes_bwd=bytes_b, us_bwd=us_b,
                achieved_bwd=bytes_b / us_b / 1e3, frac_bwd=bytes_b / us_b / 1e3 / pk['hbm'],
                moved_bytes_fwd=real_f, moved_bytes_bwd=real_b,
                fp32_alu_bound_us_exact_search=alu_us, per_kernel_us=per_kernel,
                per_vector_path=dict(us_fwd=us_f_pv, us_bwd=us_b_pv, frac=bytes_f / us_f_pv / 1e3 / pk['hbm'],
                                     frac_bwd=bytes_b / us_b_pv / 1e3 / pk['hbm'], per_kernel_us=per_kernel_pv,
                                     note='round-1 kernels: upscaled rows materialised (what upscaler dropout / product codebooks still use)'),
                note='achieved = ALGORITHMIC bytes of the per-vector formulation (SURVEY 8d: 36.7 KB / sample forward) over the '
                     'measured time; the index form moves only moved_bytes_* (tables live in shared memory), so achieved can '
                     'exceed what the same time would carry over HBM.  The forward search ranks codes on the tensor cores and '
                     're-evaluates the reference arithmetic only on the candidates (bit-identical indices); a direct-form scan '
                     'would be ALU bound at %.0f us; HBM bound of the forward path %.0f us' % (alu_us, bytes_f / pk['hbm'] / 1e3))


def run_ours(args):
    from VQCPCB import distributed as vq_dist, ops, _native
    rank, world, local = vq_dist.init_from_env()
    torch.cuda.set_device(local)
    ops.set_gemm_mode(args.mode)
    B = args.batch
    tr, gen, config = build_trainer(B, seed=1234 + rank)
    train_loader, _, _ = gen.dataloaders(batch_size=B)
    n_batches = 2
    host_batches = [next(train_loader) for _ in range(n_batches)]                       # pinned host memory
    dev_batches = [{k: v.cuda() for k, v in b.items()} for b in host_batches]           # resident in HBM
    h2d = sum(host_batches[0][k].numel() * 8 for k in ('x_left', 'x_right', 'negative_samples'))
    torch.cuda.synchronize()

    def barrier():
        torch.cuda.synchronize()
        vq_dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up (also performs the one-time codebook init + TMA/func-attribute setup)
    for i in range(args.warmup):
        tr.step(dev_batches[i % n_batches], True, False)
    barrier()

    # ---- timed: device-resident inputs
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ops.PROFILE_GEMM = []
    vq_dist.COLLECTIVE_EVENTS = [] if world > 1 else None
    launches0 = _native.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(args.steps):
        st = tr.step(dev_batches[i % n_batches], True, False)
    e1.record()
    barrier()
    launches = _native.launch_count - launches0
    prof = ops.PROFILE_GEMM
    ops.PROFILE_GEMM = None
    ms = e0.elapsed_time(e1)
    t = torch.tensor([ms], dtype=torch.float64, device='cuda')
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    ms = t.item()
    loss_val = float(st['loss'].item())
    # per-rank step time (skew) and the collective: min / max over ranks of the same timed region, and the all-reduce of
    # the flat [grads | stats | counts] buffer bracketed by CUDA events inside the optimizer (vq_dist.COLLECTIVE_EVENTS)
    skew = None
    coll = None
    if world > 1:
        tmin = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device='cuda')
        torch.distributed.all_reduce(tmin, op=torch.distributed.ReduceOp.MIN)
        skew = dict(min_ms_per_step=tmin.item() / args.steps, max_ms_per_step=ms / args.steps)
        evs = vq_dist.COLLECTIVE_EVENTS or []
        if evs:
            torch.cuda.synchronize()
            per = [a.elapsed_time(b) for a, b in evs]
            tc_ = torch.tensor([sum(per) / len(per), max(per)], dtype=torch.float64, device='cuda')
            torch.distributed.all_reduce(tc_, op=torch.distributed.ReduceOp.MAX)
            coll = dict(allreduce_ms_mean=tc_[0].item(), allreduce_ms_max=tc_[1].item(), calls=len(per),
                        bytes=int(tr.optimizer.flat.grad_full.numel() * 4),
                        note='ncclAllReduce of the flat buffer incl. the wait for the slowest rank (max over ranks)')
    vq_dist.COLLECTIVE_EVENTS = None

    # ---- timed: end to end through the public call with HOST buffers + D2H loss read every step
    e2e_steps = max(1, args.steps)
    tr.step(host_batches[0], True, False)
    barrier()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    for i in range(e2e_steps):
        st = tr.step(host_batches[i % n_batches], True, False)
        _ = st['loss'].item()                                   # device -> host read of the step's result
    e3.record()
    barrier()
    ms2 = e2.elapsed_time(e3)
    t2 = torch.tensor([ms2], dtype=torch.float64, device='cuda')
    if world > 1:
        torch.distributed.all_reduce(t2, op=torch.distributed.ReduceOp.MAX)
    ms2 = t2.item()
    if rank == 0:
        sampler.stop_flag = True
        sampler.join(timeout=2)

    # ---- second line: the same step with the reference's training dropout (0.1 in the encoder layers and the upscaler)
    drop_line = None
    if world == 1 and not args.no_dropout_line:
        enc = tr.encoder
        enc.downscaler.dropout = 0.1
        if enc.upscaler is not None:
            enc.upscaler.dropout = 0.1
        tr.step(dev_batches[0], True, False)
        barrier()
        e4, e5 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e4.record()
        for i in range(2):
            tr.step(dev_batches[i % n_batches], True, False)
        e5.record()
        barrier()
        ms3 = e4.elapsed_time(e5)
        drop_line = dict(dropout=0.1, value=B * 2 / (ms3 / 1000.0), unit=UNIT, ms_per_step=ms3 / 2, steps=2,
                         note='explicit keep masks (one byte per element) at the four sites of every encoder layer + the '
                              'upscaler; per-vector upscaler instead of the code table')
        enc.downscaler.dropout = 0.0
        if enc.upscaler is not None:
            enc.upscaler.dropout = 0.0

    qi = None
    if rank == 0 and world == 1 and not args.no_path_roofline:
        del dev_batches
        torch.cuda.empty_cache()
        qi = bench_quantizer_infonce(B)
    replicas = None
    if world > 1:
        # data-parallel correctness: after all these steps (the first of which re-initialised the codebook from data)
        # every rank must hold BIT-identical parameters
        flat = tr.optimizer.flat.data
        bits = flat.view(torch.int32).long()
        wts = torch.arange(1, bits.numel() + 1, device=bits.device, dtype=torch.long)
        chk = torch.stack([(bits * wts).sum(), (bits ^ (wts * 2654435761)).sum()])
        got = [torch.zeros_like(chk) for _ in range(world)]
        torch.distributed.all_gather(got, chk)
        same = all(torch.equal(g, got[0]) for g in got)
        replicas = dict(identical_parameters=bool(same), checksum=[int(v) for v in got[0].tolist()], ranks=world,
                        global_usage_sum=int(tr.encoder.quantizer.global_usage.sum().item()))
        assert same, 'data-parallel replicas diverged: %s' % [g.tolist() for g in got]
        try:
            torch.distributed.destroy_process_group()
        except Exception:
            pass
    if rank != 0:
        return
    pk = peaks()
    value = world * B * args.steps / (ms / 1000.0)
    e2e_value = world * B * e2e_steps / (ms2 / 1000.0)
    # dominant kernel: tensor-core GEMM launches profiled with CUDA events inside the timed region
    tc = [(fl, a.elapsed_time(b)) for (kind, fl, a, b) in prof if kind == 'tc']
    tc_ms = sum(x[1] for x in tc)
    tc_flops = sum(x[0] for x in tc)
    passes = 3 if args.mode == 'bf16x3' else 1
    roof = None
    if tc:
        ach = tc_flops / (tc_ms / 1000.0) / 1e12
        traffic = None
        tpath = os.path.join(ROOT, 'profiles', 'r02_gemm_traffic.json')
        if os.path.exists(tpath):                      # committed ncu capture: DRAM bytes per GEMM launch, same shape mix
            tj = json.load(open(tpath))
            traffic = dict(dram_bytes_per_launch=tj['dram_read_bytes_per_launch'] + tj['dram_write_bytes_per_launch'],
                           source='profiles/r02_gemm_traffic.json (ncu, %d launches)' % tj['launches'])
        roof = dict(bound='tensor', kernel='gemm_tc2_kernel / gemm_tc_kernel (tcgen05 bf16x3)' if passes == 3 else 'gemm_tc2_kernel / gemm_tc_kernel (tcgen05 bf16)',
                    achieved=ach, peak=pk['tflops'], unit='TFLOP/s', frac=ach / pk['tflops'], traffic=traffic,
                    peak_source=pk['src'], launches=len(tc), avg_launch_ms=tc_ms / len(tc),
                    share_of_step=tc_ms / ms, tensor_pipe_tflops=ach * passes,
                    note=f'algorithmic 2MNK FLOPs; the tensor pipe executes {passes}x that (bf16 split passes)')
    cpu_val, cpu_info = (None, None)
    if world == 1 and not args.no_cpu_baseline:
        cpu_val, cpu_info = cpu_port_samples_per_sec(B_cpu=CPU_SAMPLE_BATCH, steps=2, warmup=1)
    line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=args.warmup,
                ms_per_step=ms / args.steps, higher_is_better=True, scaling='weak', vs_baseline=None,
                dtype='bf16x3 (bf16 hi/lo split on tcgen05, fp32 accumulate; fp32 elsewhere)' if args.mode == 'bf16x3' else args.mode,
                data='synthetic',
                config=dict(workload=workload_name(B), global_batch=world * B, parallelism=f'dp{world}',
                            gemm_mode=args.mode, l2='per-step working set (>10 GB of activations) far exceeds the 126 MB L2',
                            blocks_per_pass=tr.max_blocks_per_pass, loss=loss_val),
                e2e=dict(value=e2e_value, unit=UNIT, h2d_bytes_per_step=h2d, d2h_bytes_per_step=4, steps=e2e_steps),
                gpu_launches=launches, clocks=sampler.summary(),
                step_tensor_frac=dict(achieved_tflops=value / world * FLOPS_PER_SAMPLE / 1e12, peak=pk['tflops'],
                                      frac=value / world * FLOPS_PER_SAMPLE / 1e12 / pk['tflops'],
                                      note='whole step: samples/s/GPU x 103.8 GFLOP algorithmic / sustained bf16 peak'))
    if skew is not None:
        line['rank_skew'] = skew
    if coll is not None:
        line['collective'] = coll
    if replicas is not None:
        line['replicas'] = replicas
    if roof:
        line['roofline'] = roof
    if cpu_val is not None:
        line['cpu_baseline'] = dict(value=cpu_val, unit=UNIT, **cpu_info)
    if drop_line is not None:
        line['with_dropout'] = drop_line
    if qi is not None:
        line['roofline_quantizer_infonce'] = qi
    print(json.dumps(line), file=_OUT, flush=True)


def _quiet_stdout():
    """Route everything libraries write to fd 1 (e.g. NCCL's version banner) to stderr; returns a file object on the real
    stdout for the one JSON line the driver parses."""
    sys.stdout.flush()
    real = os.dup(1)
    os.dup2(2, 1)
    return os.fdopen(real, 'w')


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=3)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--batch', type=int, default=4096)
    ap.add_argument('--mode', default='bf16x3', choices=['bf16x3', 'bf16', 'fp32'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-path-roofline', action='store_true')
    ap.add_argument('--no-dropout-line', action='store_true')
    ap.add_argument('--path-only', action='store_true', help='only the stand-alone quantizer + InfoNCE path measurement')
    args = ap.parse_args()
    global _OUT
    _OUT = _quiet_stdout()
    if args.impl == 'reference':
        run_reference(args)
    else:
        if not torch.cuda.is_available():
            raise SystemExit('bench.py --impl ours needs a CUDA device (no CPU fallback on the product path)')
        if args.path_only:
            print(json.dumps(bench_quantizer_infonce(args.batch)), file=_OUT, flush=True)
            return
        run_ours(args)


if __name__ == '__main__':
    main()
