"""Inference path of the relative encoder-decoder transformer that regenerates token windows from code sequences
(VQCPCB/decoders/decoder.py: constructor :24-247, ``forward`` :431-543, ``generate`` :552-722, ``init_generation`` :724-727;
SURVEY.md §8f row 3, BASELINE config 5).

Built: ``transformer_type='relative'`` with ``encoder_attention_type`` in {anticausal, causal, full} and ``cross_attention_type``
in {anticausal, full} (every shipped ``decoder_relative_*`` configuration except the 'diagonal' aligned layer), forward = teacher
forced logits + cross-entropy, and autoregressive generation that recomputes the forward for every token exactly like the
reference (no KV cache: the reference's cross-attention bias makes entry (i, j) depend on other query rows - see
``transformer/relative_maps.py``), with top-k / top-p filtering and sampling on the device.  Training of the decoder is out of
scope (forward-only kernels); parameter names are the reference's state_dict keys so its checkpoints load.

Kernels: tcgen05 / SIMT GEMMs (``ops.linear_fwd``), ``vqcpc_layernorm_fwd``, ``vqcpc_attn_map_fwd`` (one attention kernel for the
three attention flavours through host-built index maps + additive masks), ``vqcpc_embed_fwd``, ``vqcpc_softmax_ce``."""
import ctypes
import os

import numpy as np
import torch
from torch import nn

from .. import ops, _native
from ..transformer._layers import _Linear
from ..transformer.relative_maps import device_maps
from ..utils import flatten


def _p(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


class _Attn(nn.Module):
    """Parameters of ``MultiheadAttentionCustom`` (multihead_attention_custom.py:36-99): fused in-proj, out-proj, relative tables
    sized by the SOURCE length."""

    def __init__(self, d_model, nhead, seq_len_src):
        super().__init__()
        mha = nn.MultiheadAttention(d_model, nhead)
        self.in_proj_weight = nn.Parameter(mha.in_proj_weight.detach().clone())
        self.in_proj_bias = nn.Parameter(mha.in_proj_bias.detach().clone())
        self.out_proj = _Linear(d_model, d_model)
        ab = nn.Module()
        ab.e1 = nn.Parameter(torch.randn(nhead * seq_len_src, d_model // nhead))
        ab.e2 = nn.Parameter(torch.randn(nhead * seq_len_src, d_model // nhead))
        self.attn_bias = ab


def _layer(d_model, nhead, dim_feedforward, seq_len_self, seq_len_src_cross=None):
    m = nn.Module()
    m.self_attn = _Attn(d_model, nhead, seq_len_self)
    if seq_len_src_cross is not None:
        m.multihead_attn = _Attn(d_model, nhead, seq_len_src_cross)
        m.norm3 = nn.LayerNorm(d_model)
    m.linear1 = _Linear(d_model, dim_feedforward)
    m.linear2 = _Linear(dim_feedforward, d_model)
    m.norm1 = nn.LayerNorm(d_model)
    m.norm2 = nn.LayerNorm(d_model)
    return m


def _causal_mask(sz, device):
    """decoder.py:295-298: 0 where key <= query, -inf elsewhere."""
    i = torch.arange(sz, device=device)
    return torch.where(i.view(1, sz) <= i.view(sz, 1), 0.0, float('-inf')).float().contiguous()


def _anticausal_mask(sz, device, sz_tgt=None):
    """decoder.py:300-306: transpose of the causal mask, rows repeated for a longer target."""
    m = _causal_mask(sz, device).t()
    if sz_tgt is not None:
        assert sz_tgt % sz == 0
        m = torch.repeat_interleave(m, sz_tgt // sz, dim=0)
    return m.contiguous()


def top_k_top_p_filtering(logits, top_k=0, top_p=0.0, filter_value=-float('Inf')):
    """VQCPCB/utils.py:100-128, batched over the leading dimension (rows are filtered independently)."""
    logits = logits.clone()
    top_k = min(top_k, logits.size(-1))
    if top_k > 0:
        kth = torch.topk(logits, top_k)[0][..., -1, None]
        logits[logits < kth] = filter_value
    if top_p > 0.0:
        sorted_logits, sorted_indices = torch.sort(logits, descending=True)
        cumulative = torch.cumsum(torch.softmax(sorted_logits, dim=-1), dim=-1)
        remove = cumulative > top_p
        remove[..., 1:] = remove[..., :-1].clone()
        remove[..., 0] = False
        logits[remove.scatter(-1, sorted_indices, remove)] = filter_value
    return logits


class Decoder(nn.Module):
    def __init__(self, model_dir, dataloader_generator, data_processor, encoder, transformer_type, encoder_attention_type,
                 cross_attention_type, d_model, num_encoder_layers, num_decoder_layers, n_head, dim_feedforward,
                 positional_embedding_size, num_channels_encoder, num_events_encoder, num_channels_decoder, num_events_decoder,
                 dropout):
        super().__init__()
        if transformer_type != 'relative':
            raise NotImplementedError("only transformer_type='relative' is built on the B200 path")
        assert encoder_attention_type in ('anticausal', 'causal', 'full')
        if cross_attention_type not in ('anticausal', 'full'):
            raise NotImplementedError(f'cross_attention_type={cross_attention_type!r} (causal is unimplemented in the reference as '
                                      'well, decoder.py:489-493; the diagonal aligned layer is not built)')
        self.transformer_type = transformer_type
        self.encoder_attention_type = encoder_attention_type
        self.cross_attention_type = cross_attention_type
        self.model_dir = model_dir
        self.encoder = encoder
        self.encoder.eval()
        for p in self.encoder.parameters():
            p.requires_grad = False
        self.dataloader_generator = dataloader_generator
        self.data_processor = data_processor
        self.num_tokens_per_channel = data_processor.num_tokens_per_channel
        self.num_channels = len(self.num_tokens_per_channel)
        self.d_model, self.n_head = d_model, n_head
        self.num_tokens_target = data_processor.num_tokens
        self.total_upscaling = int(np.prod(encoder.downscaler.downscale_factors))
        assert self.num_tokens_target % self.total_upscaling == 0
        assert self.num_tokens_target == num_channels_decoder * num_events_decoder
        self.target_channel_embeddings = nn.Parameter(torch.randn((1, self.num_channels, positional_embedding_size)))
        self.num_events_per_code = self.total_upscaling // self.num_channels
        self.target_events_positioning_embeddings = nn.Parameter(torch.randn((1, self.num_events_per_code, positional_embedding_size)))
        S = num_channels_encoder * num_events_encoder
        T = num_channels_decoder * num_events_decoder
        self.seq_len_src, self.seq_len_tgt = S, T
        tr = nn.Module()
        tr.encoder = nn.Module()
        tr.encoder.layers = nn.ModuleList([_layer(d_model, n_head, dim_feedforward, S) for _ in range(num_encoder_layers)])
        tr.decoder = nn.Module()
        tr.decoder.layers = nn.ModuleList([_layer(d_model, n_head, dim_feedforward, T, S) for _ in range(num_decoder_layers)])
        self.transformer = tr
        self.linear_target = _Linear(data_processor.embedding_size + positional_embedding_size * 2, d_model)
        self.sos = nn.Parameter(torch.randn((1, 1, d_model)))
        q = encoder.quantizer
        if type(q).__name__ != 'NoQuantization':
            self.source_embeddings = nn.Embedding(q.codebook_size ** q.num_codebooks, d_model)
        else:
            self.source_embeddings = _Linear(q.codebook_dim, d_model)

        self.pre_softmaxes = nn.ModuleList([_Linear(d_model, n) for n in self.num_tokens_per_channel])

    def __repr__(self):
        names = dict(anticausal='AC', causal='C', full='F', diagonal='D')
        return f'Decoder-{self.transformer_type}-{names[self.encoder_attention_type]}-{names[self.cross_attention_type]}'

    def _dir(self, early_stopped):
        return f'{self.model_dir}/early_stopped' if early_stopped else f'{self.model_dir}/overfitted'

    def save(self, early_stopped):
        d = self._dir(early_stopped)
        os.makedirs(d, exist_ok=True)
        torch.save({k: v.detach().clone() for k, v in self.state_dict().items()}, f'{d}/decoder')

    def load(self, early_stopped, device):
        print(f'Loading models {self.__repr__()}')
        self.load_state_dict(torch.load(f'{self._dir(early_stopped)}/decoder', map_location=torch.device(device)))
        ops.bump_weight_version()

    # ---- one attention block: multihead_attention_custom.py:154-353 (self: :168-171, encoder-decoder: :173-196) ----------------
    def _attention(self, a, x_q, x_kv, B, Lq, Lk, mask):
        d, H = self.d_model, self.n_head
        hd = d // H
        scale = float(hd) ** -0.5
        w, b = a.in_proj_weight, a.in_proj_bias
        if x_kv is None:                      # self-attention: one fused projection, q third scaled after the bias (:247)
            qkv = ops.linear_fwd(x_q, w, b, scale_cols=(scale, d))
            q, k, v, ldq, ldkv = qkv, qkv[:, d:], qkv[:, 2 * d:], 3 * d, 3 * d
        else:                                 # q from the target, k / v from the memory through slices of the same matrix
            q = ops.linear_fwd(x_q, w[:d], b[:d], scale_cols=(scale, d))
            kv = ops.linear_fwd(x_kv, w[d:], b[d:])
            k, v, ldq, ldkv = kv, kv[:, d:], d, 2 * d
        m1, m2 = device_maps(Lq, Lk, x_q.device)
        out = torch.empty(B * Lq, d, dtype=torch.float32, device=x_q.device)
        _native.call('vqcpc_attn_map_fwd', _p(q), ldq, _p(k), _p(v), ldkv, _p(a.attn_bias.e1), _p(a.attn_bias.e2), _p(m1), _p(m2),
                     _p(mask), _p(out), None, B, Lq, Lk, H, hd, ops._stream())
        return out

    def _enc_layer(self, lyr, x, B, L, mask):
        """transformer_custom.py:268-291 (post-LN)."""
        o = self._attention(lyr.self_attn, x, None, B, L, L, mask)
        y = ops.linear_fwd(o, lyr.self_attn.out_proj.weight, lyr.self_attn.out_proj.bias, residual=x)
        x = ops.layernorm_fwd(y, lyr.norm1.weight, lyr.norm1.bias)[0]
        hmid = ops.linear_fwd(x, lyr.linear1.weight, lyr.linear1.bias, relu=True)
        y = ops.linear_fwd(hmid, lyr.linear2.weight, lyr.linear2.bias, residual=x)
        return ops.layernorm_fwd(y, lyr.norm2.weight, lyr.norm2.bias)[0]

    def _dec_layer(self, lyr, x, memory, B, T, S, tgt_mask, memory_mask):
        """transformer_custom.py:355-384: self-attention, encoder-decoder attention, feed-forward, three post-LNs."""
        o = self._attention(lyr.self_attn, x, None, B, T, T, tgt_mask)
        y = ops.linear_fwd(o, lyr.self_attn.out_proj.weight, lyr.self_attn.out_proj.bias, residual=x)
        x = ops.layernorm_fwd(y, lyr.norm1.weight, lyr.norm1.bias)[0]
        o = self._attention(lyr.multihead_attn, x, memory, B, T, S, memory_mask)
        y = ops.linear_fwd(o, lyr.multihead_attn.out_proj.weight, lyr.multihead_attn.out_proj.bias, residual=x)
        x = ops.layernorm_fwd(y, lyr.norm2.weight, lyr.norm2.bias)[0]
        hmid = ops.linear_fwd(x, lyr.linear1.weight, lyr.linear1.bias, relu=True)
        y = ops.linear_fwd(hmid, lyr.linear2.weight, lyr.linear2.bias, residual=x)
        return ops.layernorm_fwd(y, lyr.norm3.weight, lyr.norm3.bias)[0]

    @torch.no_grad()
    def forward(self, source, target):
        """source: code indices (batch, S) (or (batch, S, codebook_dim) without quantization); target: tokens (batch, num_events,
        num_channels).  Returns the reference's dict: loss, weights_per_category (list of (batch, num_events, vocab_c)),
        monitored_quantities.  decoder.py:431-543."""
        B = source.size(0)
        dev = self.sos.device
        d, C = self.d_model, self.num_channels
        if isinstance(self.source_embeddings, nn.Embedding):
            src = torch.nn.functional.embedding(source.to(dev).long(), self.source_embeddings.weight).reshape(-1, d)   # :440
        else:
            src = ops.linear_fwd(ops._f32c(source.to(dev)).reshape(-1, source.shape[-1]), self.source_embeddings.weight,
                                 self.source_embeddings.bias)
        S = src.shape[0] // B
        target = self.data_processor.preprocess(target)
        tgt = flatten(self.data_processor.embed(target))                                                 # (B, T, E)   :442-444
        T = tgt.size(1)
        chan = self.target_channel_embeddings.repeat(B, T // C, 1)                                       # :451-453
        evt = self.target_events_positioning_embeddings.repeat_interleave(C, dim=1).repeat(B, T // self.total_upscaling, 1)
        tgt = torch.cat([tgt, chan, evt], dim=2).reshape(B * T, -1)                                      # :449-457
        tgt = ops.linear_fwd(tgt.contiguous(), self.linear_target.weight, self.linear_target.bias).view(B, T, d)   # :469
        tgt = torch.cat([self.sos.expand(B, 1, d), tgt[:, :-1]], dim=1).reshape(B * T, d).contiguous()   # shift by one   :476-482
        memory_mask = _anticausal_mask(S, dev, T) if self.cross_attention_type == 'anticausal' else None  # :487-496
        if self.encoder_attention_type == 'full':
            source_mask = None
        elif self.encoder_attention_type == 'causal':
            source_mask = _causal_mask(S, dev)
        else:
            source_mask = _anticausal_mask(S, dev)
        target_mask = _causal_mask(T, dev)                                                               # :507
        x = src.contiguous()
        for lyr in self.transformer.encoder.layers:
            x = self._enc_layer(lyr, x, B, S, source_mask)
        memory = x
        y = tgt
        for lyr in self.transformer.decoder.layers:
            y = self._dec_layer(lyr, y, memory, B, T, S, target_mask, memory_mask)
        out = y.view(B * (T // C), C * d)
        weights_per_category = [ops.linear_fwd(out[:, c * d:(c + 1) * d], ps.weight, ps.bias).view(B, T // C, -1)
                                for c, ps in enumerate(self.pre_softmaxes)]                              # :519-526
        loss = 0
        for c, w in enumerate(weights_per_category):                                                     # :529-535
            l, _ = ops.softmax_ce(w.reshape(B * (T // C), -1), target_idx=target[:, :, c].reshape(-1).contiguous(), want_grad=False)
            loss = loss + l
        loss = loss.mean()
        return {'loss': loss, 'weights_per_category': weights_per_category, 'monitored_quantities': {'loss': loss.item()}}

    def init_generation(self, num_events):
        return torch.zeros(1, num_events, self.num_channels, dtype=torch.int64, device=self.sos.device)   # :724-727

    @torch.no_grad()
    def generate_from_codes(self, encoding_indices, temperature=1.0, top_k=0, top_p=1.0, generator=None, greedy=False):
        """The token loop of ``generate`` (decoder.py:607-644): for every (event, channel) a full forward, the logits of that
        position divided by the temperature, top-k / top-p filtering, one sample per batch element (``torch.multinomial`` on the
        device; the reference draws with numpy on the host).  ``greedy`` takes the argmax instead (tests).  Returns tokens
        (batch, num_events, num_channels)."""
        self.eval()
        B = encoding_indices.size(0)
        x = self.init_generation(self.data_processor.num_events).repeat(B, 1, 1)
        for event_index in range(self.data_processor.num_events):
            for channel_index in range(self.num_channels):
                logits = self.forward(encoding_indices, x)['weights_per_category'][channel_index][:, event_index, :] / temperature
                if greedy:
                    new = logits.argmax(dim=-1)
                else:
                    filtered = top_k_top_p_filtering(logits, top_k=top_k, top_p=top_p)
                    new = torch.multinomial(torch.softmax(filtered, dim=-1), 1, generator=generator).squeeze(1)
                x[:, event_index, channel_index] = new
        return x

    @torch.no_grad()
    def generate(self, temperature, batch_size=1, top_k=0, top_p=1., seed_set=None, **kwargs):
        """decoder.py:552-722 without the music21 score writing: takes a window from the train / val loader, encodes it, decodes
        ``batch_size`` samples from its codes and returns (original tokens, generated tokens, re-encoded codes)."""
        self.eval()
        gen_train, gen_val, _ = self.dataloader_generator.dataloaders(batch_size=1, shuffle_val=True)
        if seed_set not in ('val', 'train'):
            raise Exception('Need to indicate seeds dataset')
        tensor_dict = next(iter(gen_val if seed_set == 'val' else gen_train))
        x_original = tensor_dict['x'].repeat(batch_size, 1, 1)
        zs, encoding_indices, _ = self.encoder(x_original)
        encoding_indices = zs if encoding_indices is None else self.encoder.merge_codes(encoding_indices)
        x = self.generate_from_codes(encoding_indices, temperature, top_k=top_k, top_p=top_p)
        _, recoding, _ = self.encoder(torch.cat([tensor_dict['x'].to(x.device).long(), x], dim=0))
        recoding = None if recoding is None else self.encoder.merge_codes(recoding)
        return x_original, x, recoding
