"""Throughput of the VQ-CPC training step with the GRU block downscaler (``configs/encoder_vqcpc_lstm.py`` = the shape of the
reference's ``encoder_random_config.py`` and of 9 more of its 14 shipped encoder configurations: bidirectional 2-layer GRU per
16-token block, K = 32, 6 + 6 blocks, N = 15 negatives).

    python tools/lstm_bench.py [--batch 1024] [--steps 5] [--warmup 2] [--dropout 0.0] [--cpu-batch 16]

One JSON line: samples/s on the device (host token batch in, loss out, optimizer step included) and the CPU oracle's step on
the host cores (dropout 0 only: the oracle replays no random draws)."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'vqcpc-bach_b200')):
    sys.path.insert(0, p)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--batch', type=int, default=1024)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=2)
    ap.add_argument('--dropout', type=float, default=0.0)
    ap.add_argument('--cpu-batch', type=int, default=16)
    a = ap.parse_args()
    os.environ['VQCPC_BATCH'] = str(a.batch)
    os.environ['VQCPC_DROPOUT'] = str(a.dropout)
    import torch
    from VQCPCB import getters, _native
    import importlib.util
    spec = importlib.util.spec_from_file_location('cfg', os.path.join(ROOT, 'vqcpc-bach_b200', 'VQCPCB', 'configs', 'encoder_vqcpc_lstm.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    cfg = mod.config
    cfg['quantizer_kwargs']['initialize'] = True
    gen = getters.get_dataloader_generator(cfg['dataset'], cfg['training_method'], cfg['dataloader_generator_kwargs'])
    enc = getters.get_encoder('/tmp/vqcpc_lstm_bench', gen, cfg)
    tr = getters.get_encoder_trainer('/tmp/vqcpc_lstm_bench', gen, cfg['training_method'], enc, cfg['auxiliary_networks_kwargs'])
    tr.to('cuda')
    tr.init_optimizers(cfg['lr'], cfg['schedule_lr'])
    tr.train()
    train, _, _ = gen.dataloaders(batch_size=a.batch)
    for _ in range(a.warmup):
        st = tr.step(next(train), True, False)
    torch.cuda.synchronize()
    n0 = _native.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(a.steps):
        st = tr.step(next(train), True, False)
    loss = st['loss'].item()
    e1.record()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    ms = e0.elapsed_time(e1) / a.steps
    out = dict(metric='vqcpc_encoder_training_samples_per_sec', value=a.batch / ms * 1e3, unit='samples/s', ms_per_step=ms,
               ms_per_step_wall=wall / a.steps * 1e3, batch=a.batch, steps=a.steps, warmup=a.warmup, dropout=a.dropout, loss=loss,
               library_calls_per_step=(_native.launch_count - n0) // a.steps,
               config=dict(workload='encoder_vqcpc_lstm: bidirectional 2-layer GRU(512) block downscaler, K=32 D=3, 6+6 blocks of 16 '
                                    'tokens, N=15, GRU(512) context network', data='synthetic'))
    if a.cpu_batch > 0 and a.dropout == 0.0:
        from oracle import params as P, vqcpc_oracle as O
        ocfg = O.default_cfg(codebook_size=32, KL=6, KR=6, downscaler='lstm')
        prm = P.make_params(ocfg, seed=0)
        batch = P.make_batch(ocfg, a.cpu_batch, seed=1)
        torch.set_num_threads(os.cpu_count())
        O.trainer_step(batch, prm, ocfg, lr=1e-4)
        t0 = time.perf_counter()
        O.trainer_step(batch, prm, ocfg, lr=1e-4)
        cpu = time.perf_counter() - t0
        out['cpu_baseline'] = dict(kind='port', cores=torch.get_num_threads(), value=a.cpu_batch / cpu, unit='samples/s',
                                   sample=f'1 step of B={a.cpu_batch} of oracle/vqcpc_oracle.trainer_step at the same shapes')
    print(json.dumps(out))


if __name__ == '__main__':
    main()
