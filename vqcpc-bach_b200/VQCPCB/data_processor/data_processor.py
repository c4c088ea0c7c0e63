"""Token embedding front-end (VQCPCB/data_processor/data_processor.py:7-104) on the fused gather kernel."""
import torch
from torch import nn

from .. import ops
from ..utils import cuda_variable, to_numpy


class DataProcessor(nn.Module):
    """Per-voice ``nn.Embedding(V_c + 1, E)`` tables (+1 = mask token); preprocess = ``.long()`` + H2D."""

    def __init__(self, embedding_size, num_events, num_tokens_per_channel, add_mask_token=True):
        super().__init__()
        self.embedding_size = embedding_size
        self.num_events = num_events
        self.num_tokens_per_channel = num_tokens_per_channel
        self.num_channels = len(num_tokens_per_channel)
        self.num_tokens = num_events * self.num_channels
        extra = 1 if add_mask_token else 0
        self.embeddings = nn.ModuleList([nn.Embedding(n + extra, embedding_size) for n in num_tokens_per_channel])

    def _tables(self):
        return [e.weight for e in self.embeddings]

    def embed(self, x):
        """(..., num_channels) int -> (..., num_channels, E): the last dim indexes the voice."""
        assert x.shape[-1] == self.num_channels
        return ops.EmbedFn.apply(cuda_variable(x.long()), *self._tables())

    def preprocess(self, x):
        return cuda_variable(x.long())

    def postprocess(self, original, reconstruction):
        if original is not None:
            score = torch.cat([original.long(), reconstruction.cpu()], dim=1)
        else:
            score = torch.cat(reconstruction, dim=0)
        return to_numpy(score)

    def dump(self, x):
        flat = x.contiguous().view(-1).detach().cpu().numpy()
        return '_'.join(str(c).zfill(2) for c in flat)
