from torch import nn


class Downscaler(nn.Module):
    """Base class: records ``downscale_factors`` (VQCPCB/downscalers/downscaler.py:4-7)."""

    def __init__(self, downscale_factors):
        super().__init__()
        self.downscale_factors = downscale_factors
