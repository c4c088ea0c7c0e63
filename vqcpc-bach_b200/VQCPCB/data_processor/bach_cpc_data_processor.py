"""CPC block front-end (VQCPCB/data_processor/bach_cpc_data_processor.py:8-68)."""
from .data_processor import DataProcessor
from .. import ops
from ..utils import cuda_variable


class BachCPCDataProcessor(DataProcessor):
    def __init__(self, embedding_size, num_events, num_channels, num_tokens_per_channel, num_tokens_per_block):
        super().__init__(embedding_size=embedding_size, num_events=num_events,
                         num_tokens_per_channel=num_tokens_per_channel)
        assert num_channels == self.num_channels
        self.num_tokens_per_block = num_tokens_per_block

    def preprocess(self, x):
        """(..., ticks, voices) -> (..., num_blocks, tokens_per_block) int64 on the device.  The row-major flatten keeps
        the voice index fastest (token j <-> tick j // V, voice j % V), then consecutive tokens form blocks, so the
        whole thing is a reshape of the contiguous input followed by one asynchronous H2D copy."""
        ticks, voices = x.shape[-2:]
        lead = x.shape[:-2]
        assert (ticks * voices) % self.num_tokens_per_block == 0
        nb = ticks * voices // self.num_tokens_per_block
        x = x.contiguous().view(*lead, nb, self.num_tokens_per_block)
        return cuda_variable(x.long())

    def embed(self, block):
        """(..., tokens_per_block) -> (..., tokens_per_block, E); voice of token j is j % V."""
        assert block.shape[-1] % self.num_channels == 0
        return ops.EmbedFn.apply(cuda_variable(block.long()), *self._tables())
