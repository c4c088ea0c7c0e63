"""Parameter holders for the encoder-side transformer (keys match the reference checkpoints).

The math lives in the fused CUDA kernels driven by ``downscalers/_encoder_stack.py``; these modules only own the
parameters under the reference's names so state_dicts load both ways:
  VQCPCB/transformer/subsampled_relative_attention.py:8-21  -> e1, e2  (num_heads*L, head_dim)
  VQCPCB/transformer/multihead_attention_custom.py:39-62    -> in_proj_weight/bias, out_proj
  VQCPCB/transformer/transformer_custom.py:220-262, 121-140 -> linear1/2, norm1/2, layers
"""
import math

import torch
from torch import nn


class SubsampledRelativeAttention(nn.Module):
    """Learned causal (e1) / anticausal (e2) relative-position tables of one attention layer."""

    def __init__(self, head_dim, num_heads, seq_len_src, seq_len_tgt):
        super().__init__()
        if seq_len_src != seq_len_tgt:
            raise NotImplementedError('only self-attention tables (seq_len_src == seq_len_tgt) are on the encoder path')
        self.head_dim, self.num_heads = head_dim, num_heads
        self.seq_len_src, self.seq_len_tgt = seq_len_src, seq_len_tgt
        self.e1 = nn.Parameter(torch.randn(num_heads * seq_len_src, head_dim))
        self.e2 = nn.Parameter(torch.randn(num_heads * seq_len_src, head_dim))


class MultiheadAttentionCustom(nn.Module):
    def __init__(self, embed_dim, num_heads, attention_bias_type, num_channels_k, num_events_k, num_channels_q,
                 num_events_q, dropout=0.0, bias=True):
        super().__init__()
        if attention_bias_type not in ('relative_attention', 'relative_attention_target_source'):
            raise NotImplementedError('Not a valid type of attention bias')
        assert embed_dim % num_heads == 0, 'embed_dim must be divisible by num_heads'
        self.embed_dim, self.num_heads, self.dropout = embed_dim, num_heads, dropout
        self.head_dim = embed_dim // num_heads
        self.in_proj_weight = nn.Parameter(torch.empty(3 * embed_dim, embed_dim))
        self.in_proj_bias = nn.Parameter(torch.zeros(3 * embed_dim))
        self.out_proj = nn.Linear(embed_dim, embed_dim, bias=True)
        self.attn_bias = SubsampledRelativeAttention(self.head_dim, num_heads, num_channels_k * num_events_k,
                                                     num_channels_q * num_events_q)
        nn.init.xavier_uniform_(self.in_proj_weight)
        nn.init.zeros_(self.out_proj.bias)


class TransformerEncoderLayerCustom(nn.Module):
    """Post-LN layer: x = LN1(x + MHA(x)); x = LN2(x + W2 relu(W1 x))."""

    def __init__(self, d_model, nhead, attention_bias_type, num_channels, num_events, dim_feedforward=2048, dropout=0.1,
                 activation='relu'):
        super().__init__()
        if activation != 'relu':
            raise NotImplementedError('the fused FFN kernel implements ReLU (all shipped encoder configs)')
        self.self_attn = MultiheadAttentionCustom(d_model, nhead, attention_bias_type, num_channels, num_events,
                                                  num_channels, num_events, dropout=dropout)
        self.linear1 = nn.Linear(d_model, dim_feedforward)
        self.linear2 = nn.Linear(dim_feedforward, d_model)
        self.norm1 = nn.LayerNorm(d_model)
        self.norm2 = nn.LayerNorm(d_model)
        self.dropout_p = dropout
        self.seq_len = num_channels * num_events

    def param_list(self):
        a = self.self_attn
        return [a.in_proj_weight, a.in_proj_bias, a.out_proj.weight, a.out_proj.bias, a.attn_bias.e1, a.attn_bias.e2,
                self.linear1.weight, self.linear1.bias, self.linear2.weight, self.linear2.bias,
                self.norm1.weight, self.norm1.bias, self.norm2.weight, self.norm2.bias]


class TransformerEncoderCustom(nn.Module):
    """``num_layers`` deep copies of one layer — like the reference's ``_get_clones`` (transformer_custom.py:138) all
    layers of a stack therefore START with identical weights (quirk Q3 of SURVEY.md §5)."""

    def __init__(self, encoder_layer, num_layers, norm=None):
        super().__init__()
        import copy
        self.layers = nn.ModuleList([copy.deepcopy(encoder_layer) for _ in range(num_layers)])
        self.num_layers = num_layers
        if norm is not None:
            raise NotImplementedError('final norm is unused on the encoder path')
