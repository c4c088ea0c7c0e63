"""CPC heads: context network, bilinear scorers and the two losses (VQCPCB/vqcpc_helper.py:5-98).

The trainer's hot path calls ``FksModule.fused_nce`` (one kernel for u = W_k c, the N+1 scores per (b, k), the
log-sum-exp and the accuracy bit).  ``FksModule.forward`` / ``nce_loss`` / ``quantization_loss`` keep the reference
call signatures for callers that use the pieces separately.
"""
import torch
from torch import nn

from . import ops


def nce_loss(positive, negatives):
    """-(1/B) sum_b sum_k (pos - logsumexp([neg, pos]))  — positive (b, k), negatives (b, k, n)."""
    lse = torch.logsumexp(torch.cat([negatives, positive.unsqueeze(2)], dim=2), dim=2)
    return -(positive - lse).sum(1).mean(0)


def quantization_loss(loss_quantization_left, loss_quantization_negative, loss_quantization_right,
                      loss_quantization_negative_back=None):
    """Mean over the 3B (4B with the backward negatives) per-sample sums of the per-vector quantisation losses."""
    parts = [loss_quantization_left.sum(1), loss_quantization_right.sum(1),
             loss_quantization_negative.reshape(loss_quantization_negative.shape[0], -1).sum(1)]
    if loss_quantization_negative_back is not None:
        parts.append(loss_quantization_negative_back.reshape(loss_quantization_negative_back.shape[0], -1).sum(1))
    return torch.cat(parts, dim=0).mean()


class CModule(nn.Module):
    """Autoregressive context: multi-layer GRU over the left codes, last step, Linear (vqcpc_helper.py:54-76).

    ``g_ar_fwd`` is an ``nn.GRU`` used as the parameter container (same state_dict keys as the reference); the
    recurrence itself runs on the native GEMMs + GRU gate kernels (``ops.GRUFn``), not on cuDNN."""

    def __init__(self, input_dim, hidden_size, output_dim, num_layers, dropout):
        super().__init__()
        self.g_ar_fwd = nn.GRU(input_size=input_dim, hidden_size=hidden_size, num_layers=num_layers, bias=True,
                               batch_first=True, dropout=dropout, bidirectional=False)
        self.output_linear = nn.Linear(hidden_size, output_dim)

    def forward(self, zs, h):
        assert h is None, 'the context network always starts from a zero state (vqcpc_encoder_trainer.py:237)'
        c = ops.gru_forward(self.g_ar_fwd, zs, self.training)
        return ops.LinearFn.apply(c[:, -1], self.output_linear.weight, self.output_linear.bias)


class FksModule(nn.Module):
    """f_k(c, z) = z^T W_k c  with W (z_dim, c_dim, k_max) (vqcpc_helper.py:79-98)."""

    def __init__(self, z_dim, c_dim, k_max):
        super().__init__()
        self.k_max = k_max
        self.W = nn.Parameter(torch.randn(z_dim, c_dim, k_max))

    def forward(self, c_t, zs):
        """Reference-compatible scorer: c_t (B', c_dim), zs (B', k_max, z_dim) -> (B', k_max) (squeezed like the
        reference, quirk Q2).  Not used by the trainer's hot path."""
        u = torch.einsum('bc,zck->bkz', c_t, self.W)
        return torch.squeeze((u * zs.reshape(c_t.shape[0], self.k_max, -1)).sum(-1))

    def fused_nce(self, c, z_pos, z_neg):
        """c (B, c_dim), z_pos (B, k_max, z), z_neg (B, N, k_max, z) ->
        (sum_{b,k} (lse - pos)  [differentiable], pos (B,k), neg (B,k,N), correct (B,k) int32)."""
        return ops.InfoNCEFn.apply(c, self.W, z_pos, z_neg)

    def fused_nce_indexed(self, c, link, idx_pos, idx_neg, seg_pos, seg_neg):
        """Index form of ``fused_nce`` (table mode): the candidates are given by their code indices idx_pos (B, k_max),
        idx_neg (B, N, k_max) into ``link.table``; seg_pos / seg_neg = first row of those blocks in the pass's vector list."""
        return ops.InfoNCEIdxFn.apply(c, self.W, link.table, link.token, idx_pos, idx_neg, link, int(seg_pos), int(seg_neg))
