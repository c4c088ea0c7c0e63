"""Vector-quantisation bottleneck on the fused CUDA kernels.

Same classes, constructor arguments, attributes and state_dict keys as VQCPCB/quantizer/vector_quantizer.py:7-159.
The codebook is the gradient/commitment (VQ-VAE) variant the reference ships (an ``nn.ParameterList`` trained by
Adam); there is no EMA update in the reference.  ``usage_counts`` is the per-code histogram the
trainer all-reduces across ranks (the "EMA-count allreduce" of the north star carries exactly this vector).
"""
import torch
from torch import nn

from .. import ops
from .. import distributed as vq_dist


class VectorQuantizer(nn.Module):
    def __init__(self, **kwargs):
        super().__init__()

    def forward(self, inputs, **kwargs):
        raise NotImplementedError


class NoQuantization(VectorQuantizer):
    """Identity bottleneck (vector_quantizer.py:15-24)."""

    def __init__(self, codebook_dim):
        super().__init__()
        self.codebook_dim = codebook_dim

    def forward(self, inputs, **kwargs):
        loss = torch.zeros(inputs.shape[:-1], dtype=inputs.dtype, device=inputs.device)
        return inputs, None, loss


class ProductVectorQuantizer(VectorQuantizer):
    def __init__(self, codebook_size, codebook_dim, commitment_cost, num_codebooks, use_batch_norm, initialize,
                 squared_l2_norm):
        super().__init__()
        assert codebook_dim % num_codebooks == 0
        self.num_codebooks = num_codebooks
        self.codebook_dim = codebook_dim
        self.codebook_size = codebook_size
        self._commitment_cost = commitment_cost
        self.embeddings = nn.ParameterList(
            [nn.Parameter(torch.randn(codebook_size, codebook_dim // num_codebooks) * 4) for _ in range(num_codebooks)])
        self.initialize = initialize
        self.squared_l2_norm = squared_l2_norm
        self.use_batch_norm = use_batch_norm
        if use_batch_norm:
            # never enabled in a shipped config; under data parallelism it would need SyncBN
            raise NotImplementedError('use_batch_norm=True is not supported by the B200 quantizer')
        self.usage_counts = None      # int32 (num_codebooks * codebook_size), accumulated by the kernels while not None
        # data parallel: the data-dependent initialisation draws rows of the RANK-LOCAL first batch, so rank 0's draw is
        # broadcast to every replica right after it (False: single-process use inside an initialised process group)
        self.sync_init = True
        # running GLOBAL code-usage histogram (int64, codebook_size): the trainer adds the cross-rank sum of every training
        # step's counts, i.e. the vector the per-step all-reduce carries
        self.global_usage = None

    def track_usage(self, enable=True):
        dev = self.embeddings[0].device
        self.usage_counts = torch.zeros(self.num_codebooks * self.codebook_size, dtype=torch.int32, device=dev) if enable else None

    def _initialize(self, flat_input):
        """Data-dependent init: random rows of the first batch become the codebook (vector_quantizer.py:57-70)."""
        assert flat_input.shape[-1] == self.codebook_dim
        assert flat_input.shape[0] >= self.codebook_size, \
            'not enough elements in a batch to initialise the clusters. You need to increase the batch dimension.'
        ds = self.codebook_dim // self.num_codebooks
        for k, emb in enumerate(self.embeddings):
            perm = torch.randperm(flat_input.shape[0], device=flat_input.device)[:emb.shape[0]]
            emb.data.copy_(flat_input.detach()[perm, k * ds:(k + 1) * ds])
        if self.sync_init and vq_dist.world_size() > 1:
            # every rank reaches this point on its first forward (quirk Q5), so the collective is matched
            vq_dist.broadcast_parameters(list(self.embeddings), src=0)
        self.initialize = False
        ops.bump_weight_version()

    def forward(self, inputs, corrupt_labels=False, **kwargs):
        shp = inputs.shape
        flat = inputs.reshape(-1, self.codebook_dim)
        if self.initialize:
            self._initialize(flat)
        given = None
        if self.training and corrupt_labels:
            # 5 % label corruption (vector_quantizer.py:118-132): search first, then overwrite a random subset
            with torch.no_grad():
                _, idx, _ = ops.QuantizeFn.apply(flat, None, self._commitment_cost, self.squared_l2_norm, None, *self.embeddings)
                rnd = torch.randint_like(idx, low=0, high=self.codebook_size)
                keep = torch.rand(idx.shape, device=idx.device) > 0.05
                given = torch.where(keep, idx, rnd)
        zsg, idx, loss = ops.QuantizeFn.apply(flat, given, self._commitment_cost, self.squared_l2_norm, self.usage_counts,
                                              *self.embeddings)
        return zsg.view(shp), idx.view(*shp[:-1], self.num_codebooks), loss.view(shp[:-1])
