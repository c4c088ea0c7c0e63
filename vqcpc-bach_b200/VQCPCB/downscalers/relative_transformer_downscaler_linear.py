"""Transformer downscaler whose stacks end in a learned aggregation instead of a subsample
(VQCPCB/downscalers/relative_transformer_downscaler_linear.py:9-140): after the layers of stack s, ``factor``
consecutive tokens are concatenated and projected by ``linear_aggs[s] = Linear(factor * d_model, d_model)``.

Same constructor and state_dict keys as the reference class (adds ``linear_aggs.{s}.weight/bias``); the fused stack runs
every layer on all rows (no dead-row elimination: every token feeds the aggregation) and the aggregation itself is one
tensor-core GEMM with K = factor * d on a free view of the blocks-major activations."""
from torch import nn

from .relative_transformer_downscaler import RelativeTransformerDownscaler


class RelativeTransformerDownscalerLinear(RelativeTransformerDownscaler):
    LINEAR_AGG = True

    def __init__(self, input_dim, output_dim, num_channels, downscale_factors, d_model, n_head, list_of_num_layers,
                 dim_feedforward, dropout):
        super().__init__(input_dim, output_dim, num_channels, downscale_factors, d_model, n_head, list_of_num_layers,
                         dim_feedforward, dropout)
        self.linear_aggs = nn.ModuleList([nn.Linear(d_model * f, d_model) for f in downscale_factors])

    def _flat_params(self):
        ps = super()._flat_params()
        for agg in self.linear_aggs:
            ps.extend([agg.weight, agg.bias])
        return ps
