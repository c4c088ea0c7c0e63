"""Synthetic producer with the batch-dict contract of BachDataloaderGenerator (VQCPCB/dataloaders/bach_dataloader.py; used by the
student, decoder and prior trainers): ``tensor_dict['x']`` = integer CPU tensor (batch, sequences_size * subdivision, voices).
Stands in for the music21 chorale ETL, which needs a corpus that is not available offline."""
import types

import torch


class SyntheticDataloaderGenerator:
    def __init__(self, sequences_size, num_voices=4, vocab_size=64, subdivision=4, seed=1234, pin_memory=True,
                 batches_per_epoch=1 << 30):
        self.sequences_size = sequences_size
        self.num_channels = num_voices
        self.vocab_size = vocab_size
        self.seed = seed
        self.pin_memory = pin_memory
        self.batches_per_epoch = batches_per_epoch
        # what getters.get_data_processor reads off the real generator (getters.py:487-499)
        self.dataset = types.SimpleNamespace(sequences_size=sequences_size, subdivision=subdivision,
                                             index2note_dicts=[{i: f'n{i}' for i in range(vocab_size)} for _ in range(num_voices)])
        self._epoch = 0

    def _stream(self, batch_size, seed):
        g = torch.Generator()
        g.manual_seed(seed)
        pin = self.pin_memory and torch.cuda.is_available()
        events = self.dataset.sequences_size * self.dataset.subdivision
        for _ in range(self.batches_per_epoch):
            x = torch.randint(0, self.vocab_size, (batch_size, events, self.num_channels), generator=g, dtype=torch.int64)
            yield {'x': x.pin_memory() if pin else x}

    def dataloaders(self, batch_size, num_workers=0, shuffle_train=True, shuffle_val=False):
        self._epoch += 1
        base = self.seed + 7919 * self._epoch
        return (self._stream(batch_size, base), self._stream(batch_size, base + 1), self._stream(batch_size, base + 2))
