"""CPC batches from the cached chorale dataset, assembled on the device (VQCPCB/dataloaders/bach_cpc_dataloader.py:10-252;
SURVEY.md §8f row 4).

Same constructor arguments, attributes and batch-dict contract as the reference's ``BachCPCDataloaderGenerator``:

    x_left (B, KL*evb, V)   x_right (B, KR*evb, V)   negative_samples / negative_samples_back (B, N, KR, evb, V)     int64

but the dataset lives in HBM and a batch is two or three gather launches driven by the sampled row indices
(``vqcpc_cpc_window_gather``, ``vqcpc_cpc_same_seq_negatives``) instead of DataLoader collation + slicing + transposes + a pinned
host copy.  The row indices are drawn exactly like the reference's DataLoaders draw them (``datasets.chorale_dataset.IndexLoader``),
so both implementations see identical batches from identical seeds - that is what the golden test checks."""
import ctypes

import torch

from .. import _native, ops
from ..datasets.chorale_dataset import CachedChoraleBeatsDataset
from .cpc_dataloader import CPCDataloaderGenerator

subdivision = 4
num_voices = 4


def _p(t):
    return ctypes.c_void_p(t.data_ptr())


class BachCPCDataloaderGenerator(CPCDataloaderGenerator):
    def __init__(self, num_tokens_per_block, num_blocks_left, num_blocks_right, negative_sampling_method, num_negative_samples,
                 database_root=None, dataset_positive=None, dataset_negative=None, *args, **kwargs):
        assert num_tokens_per_block % (subdivision * num_voices) == 0
        super().__init__(num_tokens_per_block, num_blocks_left, num_blocks_right, negative_sampling_method, num_negative_samples)
        self.num_channels = num_voices
        per_beat = subdivision * num_voices
        num_tokens = num_tokens_per_block * (num_blocks_left + num_blocks_right)
        assert num_tokens % per_beat == 0
        if negative_sampling_method not in ('random', 'same_sequence'):
            raise NotImplementedError
        # bach_cpc_dataloader.py:41-92: windows of the whole context, and (random negatives) windows of one block
        self.dataset_positive = dataset_positive or CachedChoraleBeatsDataset(
            sequences_size=num_tokens // per_beat, subdivision=subdivision, database_root=database_root)
        self.dataset_negative = None
        if negative_sampling_method == 'random':
            self.dataset_negative = dataset_negative or CachedChoraleBeatsDataset(
                sequences_size=num_tokens_per_block // per_beat, subdivision=subdivision, database_root=database_root)

    # ---- device gathers -------------------------------------------------------------------------------------------------------
    @staticmethod
    def _window(score, rows, t0, n_ev):
        V, ticks = score.shape[1], score.shape[2]
        out = torch.empty(rows.numel(), n_ev, V, dtype=torch.int64, device=score.device)
        _native.call('vqcpc_cpc_window_gather', _p(score), score.element_size(), score.shape[0], V, ticks, _p(rows), rows.numel(), t0,
                     n_ev, _p(out), ops._stream())
        return out

    @staticmethod
    def _same_seq(score, rows, a0, KA, b0, KB, nk, evb):
        V, ticks = score.shape[1], score.shape[2]
        out = torch.empty(rows.numel(), KA + KB - 1, nk, evb, V, dtype=torch.int64, device=score.device)
        _native.call('vqcpc_cpc_same_seq_negatives', _p(score), score.element_size(), score.shape[0], V, ticks, _p(rows),
                     rows.numel(), a0, KA, b0, KB, nk, evb, _p(out), ops._stream())
        return out

    def dataloaders(self, batch_size, num_workers=0, shuffle_train=True, shuffle_val=False):
        """-> (train, val, eval) iterables of batch dicts on the device (:94-108)."""
        if self.negative_sampling_method == 'random':
            return self._dataloader_random(batch_size, shuffle_train, shuffle_val)
        return self._dataloader_same_sequence(batch_size, shuffle_train, shuffle_val)

    def _dataloader_same_sequence(self, batch_size, shuffle_train, shuffle_val):
        """:110-160.  ``num_negative_samples`` is not used (as in the reference): M = KL + KR - 1 negatives per position."""
        score = self.dataset_positive.score_tensor()
        loaders, offs = self.dataset_positive.index_loaders(batch_size, shuffle_train=shuffle_train, shuffle_val=shuffle_val)
        evb = self.num_tokens_per_block // num_voices
        KL, KR = self.num_blocks_left, self.num_blocks_right
        assert score.shape[2] == (KL + KR) * evb, f'cached windows have {score.shape[2]} ticks, the configuration needs {(KL + KR) * evb}'

        def stream(loader, off):
            for idx in loader:
                rows = idx.to(score.device, non_blocking=True) + off          # pinned index view: asynchronous copy, offset on the device
                yield {'x_left': self._window(score, rows, 0, KL * evb), 'x_right': self._window(score, rows, KL * evb, KR * evb),
                       # negatives of the right positions: all left blocks, then the right blocks but the k-th (:162-184)
                       'negative_samples': self._same_seq(score, rows, 0, KL, KL * evb, KR, KR, evb),
                       # backward: all right blocks, then the left blocks but the k-th, k < num_blocks_right (:141)
                       'negative_samples_back': self._same_seq(score, rows, KL * evb, KR, 0, KL, KR, evb)}

        return [stream(l, o) for l, o in zip(loaders, offs)]

    def _dataloader_random(self, batch_size, shuffle_train, shuffle_val):
        """:186-252: positives, negatives and backward negatives come from three independently shuffled loaders that are
        iterated in lock step (zip)."""
        score = self.dataset_positive.score_tensor()
        neg = self.dataset_negative.score_tensor()
        evb = self.num_tokens_per_block // num_voices
        KL, KR, N = self.num_blocks_left, self.num_blocks_right, self.num_negative_samples
        assert score.shape[2] == (KL + KR) * evb and neg.shape[2] == evb
        nb = batch_size * N * KR
        pos_l, pos_off = self.dataset_positive.index_loaders(batch_size, shuffle_train=shuffle_train, shuffle_val=shuffle_val)
        neg_l, neg_off = self.dataset_negative.index_loaders(nb, shuffle_train=shuffle_train, shuffle_val=shuffle_val)
        nbk_l, nbk_off = self.dataset_negative.index_loaders(nb, shuffle_train=shuffle_train, shuffle_val=shuffle_val)

        def stream(lp, op, ln, on, lb, ob):
            for ip, i_n, ib in zip(lp, ln, lb):
                rows = ip.to(score.device, non_blocking=True) + op
                rn = i_n.to(score.device, non_blocking=True) + on
                rb = ib.to(score.device, non_blocking=True) + ob
                yield {'x_left': self._window(score, rows, 0, KL * evb), 'x_right': self._window(score, rows, KL * evb, KR * evb),
                       'negative_samples': self._window(neg, rn, 0, evb).view(batch_size, N, KR, evb, num_voices),
                       'negative_samples_back': self._window(neg, rb, 0, evb).view(batch_size, N, KR, evb, num_voices)}

        return [stream(*a) for a in zip(pos_l, pos_off, neg_l, neg_off, nbk_l, nbk_off)]
