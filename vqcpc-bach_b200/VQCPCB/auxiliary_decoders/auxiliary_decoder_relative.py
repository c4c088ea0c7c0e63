"""Upscaling bidirectional transformer that reconstructs the per-channel token logits from the code sequence
(VQCPCB/auxiliary_decoders/auxiliary_decoder_relative.py:7-130) on the library's kernels.  Parameter names are the
reference's (``linear.*``, ``upscale_embeddings.{s}``, ``transformers.{s}.layers.{l}.*``, ``pre_softmaxes.{c}.*``)."""
import numpy as np
import torch
from torch import nn

from ..transformer._layers import TransformerEncoderStack, _Linear


class AuxiliaryDecoderRelative(nn.Module):
    def __init__(self, num_tokens_per_channel, codebook_dim, upscale_factors, list_of_num_layers, n_head, d_model, dim_feedforward,
                 num_tokens_bottleneck, dropout):
        super().__init__()
        assert len(list_of_num_layers) == len(upscale_factors)
        assert not dropout, 'the B200 path runs the student configuration with dropout 0 (parity setting)'
        self.num_tokens_per_channel = num_tokens_per_channel
        self.num_channels = len(num_tokens_per_channel)
        self.d_model = d_model
        self.codebook_dim = codebook_dim
        self.upscale_factors = upscale_factors
        self.linear = _Linear(codebook_dim, d_model)
        self.upscale_embeddings = nn.ParameterList([nn.Parameter(torch.randn(u, d_model)) for u in upscale_factors])
        self.num_tokens_per_transformer_block = [int(num_tokens_bottleneck * int(np.prod(upscale_factors[:i])))
                                                 for i in range(len(upscale_factors))]
        self.transformers = nn.ModuleList([
            TransformerEncoderStack(d_model, n_head, dim_feedforward, num_tokens, num_layers)
            for num_layers, num_tokens in zip(list_of_num_layers, self.num_tokens_per_transformer_block)])
        self.pre_softmaxes = nn.ModuleList([_Linear(d_model, n) for n in num_tokens_per_channel])

    def forward(self, input):
        """input (batch, num_tokens_bottleneck, codebook_dim) -> list of per-channel logits (batch, num_events, num_tokens_c)."""
        b, num_tokens, cd = input.shape
        x = self.linear(input.reshape(b * num_tokens, cd))                                       # :91
        L = num_tokens
        for transformer, factor, emb in zip(self.transformers, self.upscale_factors, self.upscale_embeddings):
            x = transformer(x, b)                                                                # :98
            x = self.upscale(x, factor, emb, self.d_model)                                       # :101-103
            L *= factor
        num_events = L // self.num_channels
        out = x.view(b * num_events, self.num_channels * self.d_model)
        return [ps(out[:, k * self.d_model:(k + 1) * self.d_model]).view(b, num_events, -1)     # :108-115
                for k, ps in enumerate(self.pre_softmaxes)]

    @staticmethod
    def upscale(input, upscale_factor, upscale_embeddings, d_model=None):
        """rows (sequence, token) -> rows (sequence, token, u): every token repeated ``upscale_factor`` times plus the u-th
        upscale embedding (:118-130; the reference works time-first, the arithmetic is the same)."""
        assert len(upscale_embeddings) == upscale_factor
        n, d = input.shape
        return (input.view(n, 1, d) + upscale_embeddings.view(1, upscale_factor, d)).reshape(n * upscale_factor, d)
