"""Golden batches of the CPC data path from the LIVE reference (authoring container only).

    python oracle/make_golden_dataloader.py        # writes tests/golden/dataloader.pt

Writes small synthetic ``tensor_dataset`` caches in the reference's layout under a temporary directory, builds the UNMODIFIED
reference ``BachCPCDataloaderGenerator`` / ``ChoraleBeatsDataset`` objects around them (their constructors need music21 + the
corpus, so the instances are made with ``__new__`` and given exactly the attributes ``dataloaders`` / ``data_loaders`` read),
seeds torch, and records the first batches of the train and validation streams for both negative sampling methods."""
import os
import sys
import tempfile
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.dont_write_bytecode = True

from oracle import dataloader_oracle as O      # noqa: E402

CASES = {
    'random': dict(cfg=dict(num_tokens_per_block=16, num_blocks_left=2, num_blocks_right=2, negative_sampling_method='random',
                            num_negative_samples=3), batch_size=4, n_pos=300, n_neg=700, seed=11),
    'same_sequence': dict(cfg=dict(num_tokens_per_block=16, num_blocks_left=3, num_blocks_right=3,
                                   negative_sampling_method='same_sequence', num_negative_samples=5), batch_size=5, n_pos=260,
                          n_neg=0, seed=12),
    'same_sequence_kl4_kr2': dict(cfg=dict(num_tokens_per_block=32, num_blocks_left=4, num_blocks_right=2,
                                           negative_sampling_method='same_sequence', num_negative_samples=5), batch_size=3,
                                  n_pos=120, n_neg=0, seed=13),
}


def main():
    sys.path.insert(0, '/root/reference')
    for n in ['matplotlib', 'matplotlib.pyplot', 'seaborn', 'music21', 'tqdm']:
        if n not in sys.modules:
            try:
                __import__(n)
            except Exception:
                sys.modules[n] = types.ModuleType(n)
    m21 = sys.modules['music21']
    if not hasattr(m21, 'corpus'):
        m21.corpus = types.SimpleNamespace(chorales=types.SimpleNamespace(Iterator=None))
        m21.interval = types.SimpleNamespace()
        m21.stream = types.SimpleNamespace()
        m21.note = types.SimpleNamespace()
    # the reference calls torch.load(path) with the pre-2.6 default (full unpickling of the TensorDataset it saved itself)
    torch.serialization.add_safe_globals([torch.utils.data.dataset.TensorDataset])
    from VQCPCB.dataloaders.bach_cpc_dataloader import BachCPCDataloaderGenerator
    from VQCPCB.datasets.chorale_dataset import ChoraleBeatsDataset
    out = {}
    tmp = tempfile.mkdtemp(prefix='vqcpc_golden_dl_')
    for name, spec in CASES.items():
        cfg, B = spec['cfg'], spec['batch_size']
        evb = cfg['num_tokens_per_block'] // 4
        ticks = (cfg['num_blocks_left'] + cfg['num_blocks_right']) * evb
        pos = O.synthetic_chorale_tensors(spec['n_pos'], ticks, spec['seed'])
        neg = O.synthetic_chorale_tensors(spec['n_neg'], evb, spec['seed'] + 100) if spec['n_neg'] else None

        def ref_dataset(tensors, tag):
            ds = ChoraleBeatsDataset.__new__(ChoraleBeatsDataset)
            ds.cache_dir = os.path.join(tmp, f'{name}_{tag}')
            os.makedirs(ds.cache_dir)
            torch.save(torch.utils.data.TensorDataset(*tensors), f'{ds.cache_dir}/tensor_dataset')
            return ds

        gen = BachCPCDataloaderGenerator.__new__(BachCPCDataloaderGenerator)
        for k, v in cfg.items():
            setattr(gen, k, v)
        gen.dataset_positive = ref_dataset(pos, 'pos')
        gen.dataset_negative = ref_dataset(neg, 'neg') if neg is not None else None
        torch.manual_seed(spec['seed'])
        train, val, _ = gen.dataloaders(batch_size=B)
        rec = dict(cfg=cfg, batch_size=B, seed=spec['seed'], positive=pos[0], negative=None if neg is None else neg[0], train=[], val=[])
        it = iter(train)
        for _ in range(3):
            rec['train'].append({k: v.clone() for k, v in next(it).items()})
        it = iter(val)
        for _ in range(2):
            rec['val'].append({k: v.clone() for k, v in next(it).items()})
        # the oracle restatement on the same tensors and seed
        torch.manual_seed(spec['seed'])
        tr_o, va_o, _ = O.cpc_batches(pos, neg, cfg, B)
        for got, want in ((tr_o, rec['train']), (va_o, rec['val'])):
            for w in want:
                g = next(got)
                assert all(torch.equal(g[k], w[k]) for k in w), name
        print(name, {k: tuple(v.shape) for k, v in rec['train'][0].items()}, 'oracle == reference')
        out[name] = rec
    # whole-window loader of the student / decoder trainers (bach_dataloader.py)
    from VQCPCB.dataloaders.bach_dataloader import BachDataloaderGenerator
    tens = O.synthetic_chorale_tensors(150, 24, 21)
    gen = BachDataloaderGenerator.__new__(BachDataloaderGenerator)
    ds = ChoraleBeatsDataset.__new__(ChoraleBeatsDataset)
    ds.cache_dir = os.path.join(tmp, 'windows')
    os.makedirs(ds.cache_dir)
    torch.save(torch.utils.data.TensorDataset(*tens), f'{ds.cache_dir}/tensor_dataset')
    gen.dataset = ds
    torch.manual_seed(21)
    train, val, _ = gen.dataloaders(batch_size=6)
    rec = dict(batch_size=6, seed=21, positive=tens[0], train=[], val=[])
    it = iter(train)
    for _ in range(3):
        rec['train'].append({k: v.clone() for k, v in next(it).items()})
    rec['val'].append({k: v.clone() for k, v in next(iter(val)).items()})
    torch.manual_seed(21)
    tr_o, va_o, _ = O.window_batches(tens, 6)
    assert all(torch.equal(next(tr_o)['x'], w['x']) for w in rec['train']) and torch.equal(next(va_o)['x'], rec['val'][0]['x'])
    print('windows', tuple(rec['train'][0]['x'].shape), 'oracle == reference')
    out['windows'] = rec
    path = os.path.join(ROOT, 'tests', 'golden', 'dataloader.pt')
    torch.save(out, path)
    print('wrote', path, os.path.getsize(path))


if __name__ == '__main__':
    main()
