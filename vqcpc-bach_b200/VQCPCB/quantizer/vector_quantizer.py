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
                 squared_l2_norm, ema_decay=None, ema_eps=1e-5):
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
        # OPTIONAL exponential-moving-average codebook (BASELINE.json north_star; the reference has NO such update - its
        # codebook is trained by Adam through q_latent_loss, vector_quantizer.py:72-83 - so this branch has no reference
        # counterpart: "parity unpinned", checked only against an independent CPU restatement of the same rule in the tests).
        # ema_decay=None (default) keeps the reference's gradient codebook.  With a decay g in (0, 1) the trainer calls
        # ``ema_update`` once per optimizer step:  N <- g N + (1 - g) n,  m <- g m + (1 - g) sum_{idx = k} z,
        # E = m / ((N + eps) / (sum N + K eps) * sum N)   (Laplace-smoothed cluster sizes), where the per-step counts n and
        # sums ride in the step's single all-reduce under data parallelism; the codebook then receives no gradient step.
        self.ema_decay = ema_decay
        self.ema_eps = ema_eps
        self._ema_N = None            # (Kc,) fp32 cluster sizes
        self._ema_m = None            # (Kc, D) fp32 cluster sums
        self._ema_cnt = None          # (Kc,) int32 counts of the current step (all passes)
        self._ema_sum = None          # (Kc, D) fp32 sums of the current step
        # running GLOBAL code-usage histogram (int64, codebook_size): the trainer adds the cross-rank sum of every training
        # step's counts, i.e. the vector the per-step all-reduce carries
        self.global_usage = None

    def track_usage(self, enable=True):
        dev = self.embeddings[0].device
        self.usage_counts = torch.zeros(self.num_codebooks * self.codebook_size, dtype=torch.int32, device=dev) if enable else None

    # ---- EMA codebook (optional extension, see __init__) -------------------------------------------------------------
    def _ema_buffers(self):
        E = self.embeddings[0]
        if self._ema_N is None or self._ema_N.device != E.device:
            assert self.num_codebooks == 1, 'the EMA codebook supports a single codebook'
            self._ema_N = torch.ones(self.codebook_size, dtype=torch.float32, device=E.device)
            self._ema_m = E.detach().clone().float()
            self._ema_cnt = torch.zeros(self.codebook_size, dtype=torch.int32, device=E.device)
            self._ema_sum = torch.zeros(self.codebook_size, self.codebook_dim, dtype=torch.float32, device=E.device)

    def ema_accumulate(self, flat_z, idx):
        """Add this pass's per-code counts and sums of z (training mode only)."""
        if self.ema_decay is None or not self.training:
            return
        self._ema_buffers()
        with torch.no_grad():
            z = flat_z.detach().reshape(-1, self.codebook_dim)
            z = z if z.is_contiguous() else z.contiguous()
            ix = idx.reshape(-1)
            ops.code_hist(ix, self._ema_cnt)
            ops.table_scatter_add(ix, z, self._ema_sum)

    def ema_step_tensors(self):
        """(float sums, int counts) of the current step for the data-parallel all-reduce, or (None, None)."""
        if self.ema_decay is None or self._ema_cnt is None:
            return None, None
        return self._ema_sum, self._ema_cnt

    @torch.no_grad()
    def ema_update(self):
        """Apply the EMA update from the (already cross-rank summed) counts / sums of the step and clear them."""
        if self.ema_decay is None or self._ema_cnt is None:
            return
        g, eps, K = float(self.ema_decay), float(self.ema_eps), self.codebook_size
        n = self._ema_cnt.float()
        self._ema_N.mul_(g).add_(n, alpha=1.0 - g)
        self._ema_m.mul_(g).add_(self._ema_sum, alpha=1.0 - g)
        tot = self._ema_N.sum()
        Ns = (self._ema_N + eps) / (tot + K * eps) * tot
        self.embeddings[0].data.copy_(self._ema_m / Ns.unsqueeze(1))
        self._ema_cnt.zero_()
        self._ema_sum.zero_()
        ops.bump_weight_version()

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
        if self.ema_decay is not None:
            self._ema_N = None                 # cluster statistics restart from the freshly drawn codes
            self._ema_buffers()
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
        self.ema_accumulate(flat, idx)
        return zsg.view(shp), idx.view(*shp[:-1], self.num_codebooks), loss.view(shp[:-1])
