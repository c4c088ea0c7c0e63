"""GRU block downscaler (VQCPCB/downscalers/lstm_downscaler.py:9-94): 10 of the 14 shipped VQ-CPC encoder configs use it.

Each block of ``downscale_factors[0]`` tokens is summarised by the last state of a multi-layer GRU and, when
``bidirectional``, of a SECOND independent GRU run on the time-flipped block; the concatenated states go through
``output_linear``.  State-dict keys are the reference's (``g_enc_fwd.*``, ``g_enc_bwd.*``, ``output_linear.*``).

The ``nn.GRU`` modules are parameter containers only: the recurrences run on the native GEMMs + GRU gate kernels
(``ops.GRUFn``: one input-projection GEMM per layer, one recurrent GEMM + one point-wise kernel per time step), the
projection on the native GEMM; inter-layer dropout uses explicit keep masks like the rest of the encoder.
"""
import math

from torch import nn

from .. import ops
from .downscaler import Downscaler


class LstmDownscaler(Downscaler):
    def __init__(self, input_dim, output_dim, num_channels, downscale_factors, hidden_size, num_layers, dropout,
                 bidirectional):
        super().__init__(downscale_factors)
        assert len(downscale_factors) == 1
        self.output_dim = output_dim
        self.num_channels = num_channels

        def gru():
            return nn.GRU(input_size=input_dim, hidden_size=hidden_size, num_layers=num_layers, bias=True,
                          batch_first=True, dropout=dropout, bidirectional=False)

        self.g_enc_fwd = gru()
        self.g_enc_bwd = gru() if bidirectional else None
        self.output_linear = nn.Linear(hidden_size * (2 if bidirectional else 1), output_dim, bias=True)

    def forward(self, inputs):
        """inputs (batch, num_blocks * block_size, dim) -> (batch, num_blocks, output_dim)."""
        batch_size, seq_len, dim = inputs.shape
        block = self.downscale_factors[0]
        assert seq_len % block == 0
        return self.compute_z(inputs.view(batch_size, seq_len // block, block, dim))

    def compute_z(self, x):
        import torch
        seq_len, input_size = x.shape[-2:]
        batch_dims = x.shape[:-2]
        x = x.reshape(math.prod(batch_dims), seq_len, input_size)
        last = ops.gru_forward(self.g_enc_fwd, x, self.training)[:, -1]
        if self.g_enc_bwd is not None:
            last = torch.cat([last, ops.gru_forward(self.g_enc_bwd, x.flip(dims=[1]), self.training)[:, -1]], dim=-1)
        z = ops.LinearFn.apply(last, self.output_linear.weight, self.output_linear.bias)
        return z.view(*batch_dims, -1)
