"""Synthetic producer with the batch-dict contract of BachCPCDataloaderGenerator
(VQCPCB/dataloaders/bach_cpc_dataloader.py:183-252): integer CPU tensors, voice dim last,

    x_left (B, KL*T, V)   x_right (B, KR*T, V)   negative_samples (B, N, KR, T, V)   negative_samples_back (same)

The real chorale ETL (music21 corpus -> cached TensorDataset) is outside the hot path and needs a corpus that is
not available offline; this generator stands in for it with uniform random tokens
(4 voices, 4 ticks per block, vocab 64 per voice, seed 1234 + rank).
"""
import types

import torch

from .cpc_dataloader import CPCDataloaderGenerator


class SyntheticCPCDataloaderGenerator(CPCDataloaderGenerator):
    def __init__(self, num_tokens_per_block, num_blocks_left, num_blocks_right, negative_sampling_method='random',
                 num_negative_samples=15, num_voices=4, vocab_size=64, sequences_size=1, subdivision=4, seed=1234,
                 pin_memory=True, batches_per_epoch=1 << 30):
        super().__init__(num_tokens_per_block, num_blocks_left, num_blocks_right, negative_sampling_method,
                         num_negative_samples)
        self.num_channels = num_voices
        self.vocab_size = vocab_size
        self.seed = seed
        self.pin_memory = pin_memory
        self.batches_per_epoch = batches_per_epoch
        # what getters.get_data_processor reads off the real generator (getters.py:500-514)
        self.dataset_positive = types.SimpleNamespace(
            sequences_size=sequences_size, subdivision=subdivision,
            index2note_dicts=[{i: f'n{i}' for i in range(vocab_size)} for _ in range(num_voices)])
        self.dataset_negative = self.dataset_positive
        self._epoch = 0

    def _stream(self, batch_size, seed):
        g = torch.Generator()
        g.manual_seed(seed)
        V = self.num_channels
        T = self.num_tokens_per_block // V
        KL, KR, N = self.num_blocks_left, self.num_blocks_right, self.num_negative_samples
        pin = self.pin_memory and torch.cuda.is_available()

        def r(*shape):
            t = torch.randint(0, self.vocab_size, shape, generator=g, dtype=torch.int64)
            return t.pin_memory() if pin else t

        for _ in range(self.batches_per_epoch):
            yield {'x_left': r(batch_size, KL * T, V), 'x_right': r(batch_size, KR * T, V),
                   'negative_samples': r(batch_size, N, KR, T, V), 'negative_samples_back': r(batch_size, N, KR, T, V)}

    def dataloaders(self, batch_size, num_workers=0, shuffle_train=True, shuffle_val=False):
        self._epoch += 1
        base = self.seed + 7919 * self._epoch
        return (self._stream(batch_size, base), self._stream(batch_size, base + 1), self._stream(batch_size, base + 2))
