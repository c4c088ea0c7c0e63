"""Throughput of the student / distillation trainer step (BASELINE config 4 shapes: codebook K = 1024, 512-token windows =
128 events x 4 voices, encoder [4,4] layers, teacher 8 layers, auxiliary decoder [4,4] layers, d = 512, 8 heads, dropout 0).

    python tools/student_bench.py [--batch 64] [--steps 5] [--warmup 2] [--cpu-batch 2]

Prints one JSON line: samples/s on the device (host token batch in, losses out, optimizer steps included), the loss values, and
the same step of the CPU oracle (oracle/student_oracle.py, the reference's restatement) on the host cores."""
import argparse
import copy
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'vqcpc-bach_b200')):
    sys.path.insert(0, p)
import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--batch', type=int, default=64)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=2)
    ap.add_argument('--cpu-batch', type=int, default=2)
    a = ap.parse_args()
    from VQCPCB import getters
    from VQCPCB.configs.encoder_student import config as base
    cfg = copy.deepcopy(base)
    cfg['dataloader_generator_kwargs']['sequences_size'] = 32          # 128 events = 512 tokens
    cfg['quantizer_kwargs'].update(codebook_size=1024, initialize=True)
    gen = getters.get_dataloader_generator(cfg['dataset'], cfg['training_method'], cfg['dataloader_generator_kwargs'])
    enc = getters.get_encoder('/tmp/vqcpc_student_bench', gen, cfg)
    tr = getters.get_encoder_trainer('/tmp/vqcpc_student_bench', gen, cfg['training_method'], enc, cfg['auxiliary_networks_kwargs'])
    tr.to('cuda')
    tr.init_optimizers(1e-5, False)
    tr.train()
    train, _, _ = gen.dataloaders(batch_size=a.batch)
    for _ in range(a.warmup):
        q = tr.step(next(train), train=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        q = tr.step(next(train), train=True)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    n_params = sum(p.numel() for m in (tr.encoder, tr.teacher, tr.auxiliary_decoder) for p in m.parameters())
    out = dict(metric='student_encoder_training_samples_per_sec', value=a.batch / ms * 1e3, unit='samples/s', ms_per_step=ms,
               batch=a.batch, steps=a.steps, warmup=a.warmup, parameters=n_params,
               config=dict(workload='encoder_student K=1024, 512-token windows (128 events x 4 voices), encoder [4,4] / teacher 8 / '
                                    'decoder [4,4] layers, d=512, 8 heads, dropout 0; teacher step + encoder-decoder step per batch'),
               losses={k: float(v) for k, v in q.items()}, data='synthetic')
    if a.cpu_batch > 0:
        from oracle import student_oracle as S
        ocfg = S.student_cfg(list_of_num_layers=[4, 4], num_events=128, teacher_layers=8, dec_layers=[4, 4], codebook_size=1024,
                             vocab=[64, 64, 64, 64], num_events_masked=4)
        prm = S.make_student_params(ocfg, seed=0)
        batch = S.make_student_batch(ocfg, a.cpu_batch, seed=1)
        torch.set_num_threads(os.cpu_count())
        S.student_step(batch, prm, ocfg, 7, lr=1e-5)
        t0 = time.perf_counter()
        S.student_step(batch, prm, ocfg, 9, lr=1e-5)
        dt = time.perf_counter() - t0
        out['cpu_baseline'] = dict(value=a.cpu_batch / dt, unit='samples/s', cores=os.cpu_count(), kind='port',
                                   sample=f'1 step of B={a.cpu_batch} after 1 warm-up, oracle/student_oracle.student_step, torch CPU fp32')
    print(json.dumps(out))


if __name__ == '__main__':
    main()
