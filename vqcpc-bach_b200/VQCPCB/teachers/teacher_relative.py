"""Bidirectional relative-attention teacher of the student trainer (VQCPCB/teachers/teacher_relative.py:8-87) on the library's
kernels: input Linear, a stack of whole-sequence encoder layers (``transformer/_layers.py`` -> tcgen05 GEMMs, LayerNorm and the
full-sequence attention kernels of ``csrc/attn_long.cu``), one pre-softmax Linear per channel.  Parameter names are the
reference's (``channel_embeddings``, ``linear_to_input_transformer.*``, ``transformer.layers.{l}.*``, ``pre_softmaxes.{c}.*``,
``data_processor.embeddings.{c}.weight``)."""
import torch
from torch import nn

from ..transformer._layers import TransformerEncoderStack, _Linear
from ..utils import flatten


class TeacherRelative(nn.Module):
    def __init__(self, data_processor, num_layers, num_tokens_per_channel, positional_embedding_size, d_model, dim_feedforward,
                 n_head, num_tokens, dropout):
        super().__init__()
        self.num_channels = len(num_tokens_per_channel)
        self.data_processor = data_processor
        input_dim = data_processor.embedding_size
        assert num_tokens % self.num_channels == 0
        assert not dropout, 'the B200 path runs the student configuration with dropout 0 (parity setting)'
        self.channel_embeddings = nn.Parameter(torch.randn((1, self.num_channels, positional_embedding_size)))
        self.num_layers = num_layers
        self.d_model = d_model
        self.linear_to_input_transformer = _Linear(input_dim, d_model - positional_embedding_size)
        self.transformer = TransformerEncoderStack(d_model, n_head, dim_feedforward, num_tokens, num_layers)
        self.num_tokens_per_channel = num_tokens_per_channel
        self.pre_softmaxes = nn.ModuleList([_Linear(d_model, n) for n in num_tokens_per_channel])

    def forward(self, x):
        """x (batch, num_events, num_channels, input_dim) -> list of num_channels logits (batch, num_events, num_tokens_c)."""
        b, num_events, c, e = x.shape
        assert c == self.num_channels
        L = num_events * c
        h = self.linear_to_input_transformer(flatten(x).reshape(b * L, e))                        # :63-64
        chan = self.channel_embeddings.reshape(1, c, -1).expand(b * num_events, c, -1).reshape(b * L, -1)
        seq = torch.cat([h, chan], dim=1)                                                          # :70-74
        out = self.transformer(seq, b)                                                             # :76-78 (sequence-major rows)
        out = out.view(b * num_events, c * self.d_model)
        return [ps(out[:, k * self.d_model:(k + 1) * self.d_model]).view(b, num_events, -1)      # :79-86
                for k, ps in enumerate(self.pre_softmaxes)]
