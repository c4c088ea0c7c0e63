"""Block-wise relative-attention transformer downscaler (B200 path).

Same constructor, attributes and state_dict keys as the reference class
(VQCPCB/downscalers/relative_transformer_downscaler.py:9-133); ``forward`` runs the fused kernel stack."""
import numpy as np
import torch
from torch import nn

from .downscaler import Downscaler
from ._encoder_stack import EncoderStackFn
from ..transformer.relative_attention import TransformerEncoderLayerCustom, TransformerEncoderCustom


class RelativeTransformerDownscaler(Downscaler):
    """(batch, num_tokens, embedding_dim) -> (batch, num_tokens // prod(downscale_factors), output_dim)."""

    POS_EMB = 8
    LINEAR_AGG = False       # True in RelativeTransformerDownscalerLinear: stacks end in Linear(factor * d -> d)

    def __init__(self, input_dim, output_dim, num_channels, downscale_factors, d_model, n_head, list_of_num_layers,
                 dim_feedforward, dropout):
        super().__init__(downscale_factors)
        assert len(downscale_factors) == len(list_of_num_layers), \
            'number of transfo must match number of downscaling factors'
        self.sequence_length = int(np.prod(downscale_factors))
        self.num_channels = num_channels
        self.num_events = self.sequence_length // num_channels
        self.d_model, self.n_head, self.dropout = d_model, n_head, dropout
        self.list_of_num_layers = list(list_of_num_layers)
        pe = self.POS_EMB
        self.input_linear = nn.Linear(input_dim, d_model - 2 * pe)
        self.target_channel_embeddings = nn.Parameter(torch.randn(1, 1, num_channels, pe))
        self.events_positioning_embeddings = nn.Parameter(torch.randn(1, 1, self.num_events, pe))
        self.output_dim = output_dim
        self.output_linear = nn.Linear(d_model, output_dim)

        stacks = []
        n_ev, n_ch = self.num_events, num_channels
        for factor, n_layers in zip(downscale_factors, list_of_num_layers):
            layer = TransformerEncoderLayerCustom(d_model=d_model, nhead=n_head, attention_bias_type='relative_attention',
                                                  num_channels=n_ch, num_events=n_ev, dim_feedforward=dim_feedforward,
                                                  dropout=dropout)
            stacks.append(TransformerEncoderCustom(layer, n_layers))
            n_ev = (n_ev * n_ch) // factor
            if n_ch > 1:
                assert n_ch <= factor, \
                    f'First stack of downscaler transfo has to be larger than input num channels = {n_ch}'
                n_ch = 1
        self.transformers = nn.ModuleList(stacks)

    def _flat_params(self):
        ps = [self.input_linear.weight, self.input_linear.bias, self.target_channel_embeddings,
              self.events_positioning_embeddings, self.output_linear.weight, self.output_linear.bias]
        for stack in self.transformers:
            for layer in stack.layers:
                ps.extend(layer.param_list())
        return ps

    def forward(self, embedded_seq):
        batch, seq_len, dim = embedded_seq.shape
        assert seq_len % self.sequence_length == 0
        nb = seq_len // self.sequence_length
        meta = (self.sequence_length, self.num_channels, self.POS_EMB, self.d_model, self.n_head,
                tuple(self.downscale_factors), tuple(self.list_of_num_layers), self.output_dim, self.LINEAR_AGG,
                float(self.dropout) if self.training else 0.0, torch.is_grad_enabled())
        z = EncoderStackFn.apply(embedded_seq.reshape(batch * seq_len, dim), meta, *self._flat_params())
        return z.view(batch, nb, self.output_dim)
