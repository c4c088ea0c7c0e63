"""The cached chorale tensor dataset, resident in HBM (VQCPCB/datasets/chorale_dataset.py: cache layout :55, :109-183, index
dictionaries :70-107, ``data_loaders`` :542-595).

The reference builds ``tensor_dataset`` = ``TensorDataset(score (N, voices, ticks) int32, metadata)`` from the music21 chorale
corpus once and ``torch.save``s it under ``<database_root>/<repr>/tensor_dataset``; every later run only loads that file.  The
ETL itself (music21 + corpus) is outside this library: ``CachedChoraleBeatsDataset`` consumes the cache the reference wrote
(same directory naming, same pickled ``index_dicts/chorale_beats_dataset.pkl``), keeps the score tensor on the device and hands
out INDEX batches in exactly the order the reference's ``DataLoader``s would visit the rows (same consumption of torch's global
generator), so that the windows can be gathered on the GPU (``dataloaders/bach_cpc_dataloader.py``)."""
import os
import pickle

import torch

SLUR_SYMBOL, START_SYMBOL, END_SYMBOL, REST_SYMBOL, OUT_OF_RANGE, PAD_SYMBOL = '__', 'START', 'END', 'rest', 'OOR', 'XX'


class IndexLoader:
    """Row-index batches of ``DataLoader(TensorDataset(...), batch_size, shuffle, drop_last=True)`` without touching the rows.

    Consumes torch's global generator like the real thing (torch/utils/data/dataloader.py ``_BaseDataLoaderIter.__init__``: one
    int64 draw per ``iter()``; sampler.py ``RandomSampler.__iter__``: one int64 draw at the first ``next()`` seeding a private
    generator whose ``randperm(n)`` is the visiting order), so a reference run and this one started from the same
    ``torch.manual_seed`` see the same batches; tests/test_cpu_boundary.py compares it with a real DataLoader."""

    def __init__(self, n, batch_size, shuffle):
        self.n, self.batch_size, self.shuffle = int(n), int(batch_size), bool(shuffle)

    def __len__(self):
        return self.n // self.batch_size

    def __iter__(self):
        torch.empty((), dtype=torch.int64).random_()                      # _base_seed of the DataLoader iterator
        return self._batches()

    def _batches(self):
        if self.shuffle:
            seed = int(torch.empty((), dtype=torch.int64).random_().item())
            g = torch.Generator()
            g.manual_seed(seed)
            order = torch.randperm(self.n, generator=g)
        else:
            order = torch.arange(self.n)
        if torch.cuda.is_available():
            order = order.pin_memory()                                    # batches are views: asynchronous H2D of the indices
        for i in range(self.n // self.batch_size):
            yield order[i * self.batch_size:(i + 1) * self.batch_size]


class CachedChoraleBeatsDataset:
    """``ChoraleBeatsDataset`` restricted to what training reads: the cached score tensor and the index dictionaries."""

    def __init__(self, voice_ids=(0, 1, 2, 3), metadatas=None, sequences_size=8, subdivision=4, database_root=None,
                 cache_dir=None, device=None):
        self.voice_ids = list(voice_ids)
        self.num_voices = len(self.voice_ids)
        self.name = 'chorale_beats_dataset'
        self.metadatas = metadatas if metadatas is not None else []
        self.sequences_size = sequences_size
        self.subdivision = subdivision
        self.database_root = database_root or os.environ.get('VQCPC_DATA_ROOT') or os.path.abspath(
            os.path.join(os.path.dirname(__file__), '..', '..', '..', 'data'))
        self.cache_dir = cache_dir or f'{self.database_root}/{self.__repr__()}'
        self.device = device
        self.index2note_dicts = self.note2index_dicts = self.voice_ranges = None
        path = os.path.join(self.database_root, 'index_dicts', f'{self.name}.pkl')
        if os.path.isfile(path):
            with open(path, 'rb') as ff:
                d = pickle.load(ff)
            self.index2note_dicts, self.note2index_dicts, self.voice_ranges = d['index2note_dicts'], d['note2index_dicts'], d['voice_ranges']
        self.list_symbol_except_notes = [SLUR_SYMBOL, START_SYMBOL, END_SYMBOL, OUT_OF_RANGE, REST_SYMBOL, PAD_SYMBOL]
        self._score = None

    def __repr__(self):                                                   # chorale_dataset.py:56-62 (names the cache directory)
        return (f'ChoraleBeatsDataset({self.voice_ids},{self.name},{[m.name for m in self.metadatas]},'
                f'{self.sequences_size},{self.subdivision})')

    def score_tensor(self):
        """(N, voices, ticks) on the device, the dtype of the cache (int32 in the reference's files)."""
        if self._score is None:
            path = f'{self.cache_dir}/tensor_dataset'
            if not os.path.exists(path):
                raise FileNotFoundError(
                    f'{path} not found.  The chorale ETL (music21 + corpus -> tensor_dataset) is not part of this library: run the '
                    'reference once to write its cache, then point database_root / VQCPC_DATA_ROOT (or cache_dir) at it.')
            td = torch.load(path, weights_only=False)
            score = td.tensors[0] if hasattr(td, 'tensors') else td[0]
            assert score.dim() == 3 and score.shape[1] == self.num_voices, f'unexpected score tensor {tuple(score.shape)}'
            assert score.dtype in (torch.int32, torch.int64)
            if self.index2note_dicts is None:                             # no pickle: size the vocabularies from the data
                self.index2note_dicts = [{i: str(i) for i in range(int(score[:, v].max()) + 1)} for v in range(self.num_voices)]
            dev = self.device or ('cuda' if torch.cuda.is_available() else 'cpu')
            self._score = score.contiguous().to(dev)
        return self._score

    def split_ranges(self, split=(0.85, 0.10)):
        """Row ranges of train / val / eval (chorale_dataset.py:562-568)."""
        assert sum(split) < 1
        n = self.score_tensor().shape[0]
        a, b = split
        return (0, int(a * n)), (int(a * n), int((a + b) * n)), (int((a + b) * n), n)

    def index_loaders(self, batch_size, split=(0.85, 0.10), shuffle_train=True, shuffle_val=False):
        """(train, val, eval) ``IndexLoader``s + the first row of each split (chorale_dataset.py:570-595)."""
        ranges = self.split_ranges(split)
        loaders = [IndexLoader(hi - lo, batch_size, sh) for (lo, hi), sh in zip(ranges, (shuffle_train, shuffle_val, False))]
        return loaders, [lo for lo, _ in ranges]
