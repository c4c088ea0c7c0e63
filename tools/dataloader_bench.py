"""Batch assembly of the CPC data path at the headline shapes (BASELINE config 2: batch 4096, KL = KR = 8 blocks of 16 tokens,
N = 15 random negatives) from a synthetic cached dataset: device gathers vs the reference's DataLoader pipeline (CPU oracle).

    python tools/dataloader_bench.py [--batch 4096] [--steps 20] [--cpu-batches 1] [--rows 40000]

One JSON line: batches/s and samples/s on the device (index draw on the host + H2D of the indices + 4 gather launches), achieved
GB/s against the algorithmic bytes (4 B read + 8 B written per token), and the oracle pipeline's batches/s on the host."""
import argparse
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'vqcpc-bach_b200')):
    sys.path.insert(0, p)
import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--batch', type=int, default=4096)
    ap.add_argument('--steps', type=int, default=4)
    ap.add_argument('--cpu-batches', type=int, default=1)
    ap.add_argument('--rows', type=int, default=40000)
    ap.add_argument('--method', default='random', choices=['random', 'same_sequence'])
    a = ap.parse_args()
    from oracle import dataloader_oracle as O
    from VQCPCB import getters
    from VQCPCB.datasets.chorale_dataset import CachedChoraleBeatsDataset
    KL = KR = 8
    N = 15
    cfg = dict(num_tokens_per_block=16, num_blocks_left=KL, num_blocks_right=KR, negative_sampling_method=a.method,
               num_negative_samples=N)
    root = tempfile.mkdtemp(prefix='vqcpc_dl_bench_')
    B = a.batch
    n_neg = max(a.rows, int(B * N * KR / 0.85) * (a.steps + 2))
    pos = O.synthetic_chorale_tensors(a.rows, (KL + KR) * 4, seed=1)
    neg = O.synthetic_chorale_tensors(n_neg, 4, seed=2) if a.method == 'random' else None
    for seq, t in ((KL + KR, pos), (1, neg)):
        if t is None:
            continue
        ds = CachedChoraleBeatsDataset(sequences_size=seq, database_root=root)
        os.makedirs(ds.cache_dir, exist_ok=True)
        torch.save(torch.utils.data.TensorDataset(*t), f'{ds.cache_dir}/tensor_dataset')
    gen = getters.get_dataloader_generator('bach', 'vqcpc', dict(cfg, database_root=root))
    train, _, _ = gen.dataloaders(batch_size=B)
    batch = next(train)
    torch.cuda.synchronize()
    tokens = sum(v.numel() for v in batch.values())
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    steps = 0
    for batch in train:
        steps += 1
        if steps == a.steps:
            break
    e1.record()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    ms = e0.elapsed_time(e1) / steps
    out = dict(metric='cpc_batches_per_sec', value=steps / wall, unit='batches/s', samples_per_sec=steps * B / wall, batch=B,
               steps=steps, ms_per_batch_wall=wall / steps * 1e3, ms_per_batch_device=ms, tokens_per_batch=tokens,
               algorithmic_bytes_per_batch=tokens * 12, achieved_GBps_wall=tokens * 12 / (wall / steps) / 1e9,
               config=dict(workload=f'CPC batch assembly, {a.method} negatives, KL=KR=8 x 16 tokens, N=15, batch {B}; '
                                    f'{a.rows} positive windows, {n_neg if neg is not None else 0} negative blocks resident in HBM'))
    if a.cpu_batches > 0:
        torch.set_num_threads(os.cpu_count())
        tr, _, _ = O.cpc_batches(pos, neg, cfg, B)
        next(tr)
        t0 = time.perf_counter()
        for _ in range(a.cpu_batches):
            b = next(tr)
            b = {k: v.contiguous() for k, v in b.items()}            # the transposes are lazy views until the H2D copy reads them
        cpu = (time.perf_counter() - t0) / a.cpu_batches
        out['cpu_baseline'] = dict(kind='port', cores=torch.get_num_threads(), value=1.0 / cpu, unit='batches/s', s_per_batch=cpu,
                                   sample=f'{a.cpu_batches} batch(es) through torch DataLoaders like the reference (collate + slices + transposes)')
    print(json.dumps(out))


if __name__ == '__main__':
    main()
