"""CPU: the data-path oracle (oracle/dataloader_oracle.py) against batches recorded from the live reference, and the product's
host-side index streams (VQCPCB/datasets/chorale_dataset.py::IndexLoader) against real torch DataLoaders."""
import os

import pytest
import torch

from oracle import dataloader_oracle as O

GOLD = os.path.join(os.path.dirname(__file__), 'golden', 'dataloader.pt')


@pytest.mark.parametrize('case', ['random', 'same_sequence', 'same_sequence_kl4_kr2'])
def test_dataloader_oracle_matches_reference_golden(case):
    rec = torch.load(GOLD)[case]
    pos = (rec['positive'], torch.zeros(rec['positive'].shape + (1,), dtype=torch.int32))
    neg = None if rec['negative'] is None else (rec['negative'], torch.zeros(rec['negative'].shape + (1,), dtype=torch.int32))
    torch.manual_seed(rec['seed'])
    train, val, _ = O.cpc_batches(pos, neg, rec['cfg'], rec['batch_size'])
    for stream, want in ((train, rec['train']), (val, rec['val'])):
        for w in want:
            g = next(stream)
            for k in w:
                assert torch.equal(g[k], w[k]), (case, k)


@pytest.mark.parametrize('n,bs,shuffle', [(255, 4, True), (97, 24, True), (40, 7, False)])
def test_index_loader_visits_rows_like_a_dataloader(n, bs, shuffle):
    """Three loaders iterated in lock step, like _dataloader_random (bach_cpc_dataloader.py:201-247): the emulated index
    streams consume torch's global generator exactly like DataLoader + RandomSampler do."""
    from torch.utils.data import DataLoader, TensorDataset
    from VQCPCB.datasets.chorale_dataset import IndexLoader
    torch.manual_seed(5)
    real = [DataLoader(TensorDataset(torch.arange(n)), batch_size=bs, shuffle=shuffle, drop_last=True) for _ in range(3)]
    want = [[b[0].clone() for b in trip] for trip in zip(*real)]
    after_real = torch.rand(1)
    torch.manual_seed(5)
    mine = [IndexLoader(n, bs, shuffle) for _ in range(3)]
    got = [list(trip) for trip in zip(*mine)]
    assert len(got) == len(want) == n // bs
    for g, w in zip(got, want):
        for a, b in zip(g, w):
            assert torch.equal(a, b)
    assert torch.equal(torch.rand(1), after_real)          # the global generator is left in the same state
    # a second epoch draws a fresh order, again like the real loader
    torch.manual_seed(9)
    e1 = [b[0].clone() for b in real[0]]
    e2 = [b[0].clone() for b in real[0]]
    torch.manual_seed(9)
    m1, m2 = list(mine[0]), list(mine[0])
    assert all(torch.equal(a, b) for a, b in zip(e1 + e2, m1 + m2))


def test_cached_dataset_layout_and_errors(tmp_path):
    """Cache directory naming of the reference (chorale_dataset.py:55-62), split ranges (:562-568), loud failure without a cache."""
    from VQCPCB.datasets.chorale_dataset import CachedChoraleBeatsDataset
    ds = CachedChoraleBeatsDataset(sequences_size=4, database_root=str(tmp_path), device='cpu')
    assert ds.cache_dir == f'{tmp_path}/ChoraleBeatsDataset([0, 1, 2, 3],chorale_beats_dataset,[],4,4)'
    with pytest.raises(FileNotFoundError):
        ds.score_tensor()
    os.makedirs(ds.cache_dir)
    score, meta = O.synthetic_chorale_tensors(101, 16, seed=1)
    torch.save(torch.utils.data.TensorDataset(score, meta), f'{ds.cache_dir}/tensor_dataset')
    assert torch.equal(ds.score_tensor(), score)
    assert ds.split_ranges() == ((0, 85), (85, 95), (95, 101))
    assert [len(d) for d in ds.index2note_dicts] == [int(score[:, v].max()) + 1 for v in range(4)]


def test_window_loader_oracle_matches_reference_golden():
    rec = torch.load(GOLD)['windows']
    tens = (rec['positive'], torch.zeros(rec['positive'].shape + (1,), dtype=torch.int32))
    torch.manual_seed(rec['seed'])
    train, val, _ = O.window_batches(tens, rec['batch_size'])
    for w in rec['train']:
        assert torch.equal(next(train)['x'], w['x'])
    assert torch.equal(next(val)['x'], rec['val'][0]['x'])
