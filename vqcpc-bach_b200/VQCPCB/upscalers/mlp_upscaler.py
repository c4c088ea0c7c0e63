"""MLP upscaler Linear -> Dropout -> SELU -> Linear (VQCPCB/upscalers/mlp_upscaler.py:4-35) on the native GEMM/SELU
kernels.  ``self.mlp`` keeps the reference's nn.Sequential so the state_dict keys stay ``mlp.0.*`` / ``mlp.3.*``."""
from torch import nn

from .. import ops


class MlpUpscaler(nn.Module):
    def __init__(self, input_dim, output_dim, hidden_size, dropout):
        super().__init__()
        self.input_dim, self.output_dim, self.dropout = input_dim, output_dim, dropout
        self.mlp = nn.Sequential(nn.Linear(input_dim, hidden_size, bias=True), nn.Dropout(p=dropout), nn.SELU(),
                                 nn.Linear(hidden_size, output_dim, bias=True))

    def forward(self, inputs):
        if self.training and self.dropout > 0:
            raise NotImplementedError('dropout > 0 is not implemented on the fused path yet (use dropout=0.0 or .eval())')
        h = ops.LinearFn.apply(inputs, self.mlp[0].weight, self.mlp[0].bias)
        h = ops.SeluFn.apply(h)
        return ops.LinearFn.apply(h, self.mlp[3].weight, self.mlp[3].bias)
