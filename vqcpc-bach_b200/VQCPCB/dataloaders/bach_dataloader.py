"""Whole-window batches from the cached chorale dataset, gathered on the device (VQCPCB/dataloaders/bach_dataloader.py:9-33; the
producer of the student trainer and of ``Decoder.generate``): ``tensor_dict['x']`` (batch, sequences_size * subdivision, voices)
int64.  Same dataset residency and index order as ``bach_cpc_dataloader.py`` (see there); score writing (``write`` / ``to_score``,
music21) is out of scope."""
import torch

from ..datasets.chorale_dataset import CachedChoraleBeatsDataset
from .bach_cpc_dataloader import BachCPCDataloaderGenerator

subdivision = 4
num_voices = 4


class BachDataloaderGenerator:
    def __init__(self, sequences_size, database_root=None, dataset=None):
        self.dataset = dataset or CachedChoraleBeatsDataset(sequences_size=sequences_size, subdivision=subdivision,
                                                            database_root=database_root)
        self.num_channels = num_voices

    def dataloaders(self, batch_size, num_workers=0, shuffle_train=True, shuffle_val=False):
        score = self.dataset.score_tensor()
        loaders, offs = self.dataset.index_loaders(batch_size, shuffle_train=shuffle_train, shuffle_val=shuffle_val)

        # the reference returns generator EXPRESSIONS (bach_dataloader.py:26-33), which call iter(dataloader) when they are
        # created: the three loaders draw their base seeds here, in order, before any batch is requested
        iterators = [iter(l) for l in loaders]

        def stream(it, off):
            for idx in it:
                rows = idx.to(score.device, non_blocking=True) + off          # pinned index view: asynchronous copy, offset on the device
                yield {'x': BachCPCDataloaderGenerator._window(score, rows, 0, score.shape[2])}

        return [stream(it, o) for it, o in zip(iterators, offs)]

    def write(self, x, path):
        raise NotImplementedError('score writing needs music21 (outside this library)')

    to_score = write
