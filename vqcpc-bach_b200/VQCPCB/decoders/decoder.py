"""Inference path of the relative encoder-decoder transformer that regenerates token windows from code sequences
(VQCPCB/decoders/decoder.py: constructor :24-247, ``forward`` :431-543, ``generate`` :552-722, ``init_generation`` :724-727;
SURVEY.md §8f row 3, BASELINE config 5).

Built: ``transformer_type`` 'relative' and 'absolute' (learnt absolute positional embeddings, no attention bias: decoder_type
'transformer') with ``encoder_attention_type`` in {anticausal, causal, full} and ``cross_attention_type``
in {anticausal, full, diagonal} (every shipped ``decoder_relative_*`` configuration; 'diagonal' = the aligned layer whose
cross term is an MLP of the aligned memory event, transformer_custom.py:389-479), forward = teacher
forced logits + cross-entropy, and autoregressive generation that recomputes the forward for every token exactly like the
reference does (kept as the parity check), and - the default - INCREMENTALLY over cached keys / values (csrc/decode.cu; see
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


_MAX_ROWS = 32          # sequences per decode session (vqcpc_linear_small takes up to 32 rows; d_model <= 512 above 16)


class _Attn(nn.Module):
    """Parameters of ``MultiheadAttentionCustom`` (multihead_attention_custom.py:36-99): fused in-proj, out-proj, relative tables
    sized by the SOURCE length."""

    def __init__(self, d_model, nhead, seq_len_src, relative=True):
        super().__init__()
        mha = nn.MultiheadAttention(d_model, nhead)
        self.in_proj_weight = nn.Parameter(mha.in_proj_weight.detach().clone())
        self.in_proj_bias = nn.Parameter(mha.in_proj_bias.detach().clone())
        self.out_proj = _Linear(d_model, d_model)
        self.attn_bias = None                  # attention_bias_type=None (transformer_type 'absolute', :98-99)
        if relative:
            ab = nn.Module()
            ab.e1 = nn.Parameter(torch.randn(nhead * seq_len_src, d_model // nhead))
            ab.e2 = nn.Parameter(torch.randn(nhead * seq_len_src, d_model // nhead))
            self.attn_bias = ab

    def tables(self):
        return (None, None) if self.attn_bias is None else (self.attn_bias.e1, self.attn_bias.e2)


def _layer(d_model, nhead, dim_feedforward, seq_len_self, seq_len_src_cross=None, aligned=None, relative=True):
    """aligned = (num_channels_encoder, num_channels_decoder): the 'diagonal' layer (transformer_custom.py:389-479) whose cross
    term is ``cross_attn = Sequential(Linear(Ce * d, 2 d), ELU, Linear(2 d, d * Cd))`` of the aligned memory event (:428-432)."""
    m = nn.Module()
    m.self_attn = _Attn(d_model, nhead, seq_len_self, relative)
    if aligned is not None:
        ce, cd = aligned
        m.cross_attn = nn.Sequential(_Linear(ce * d_model, 2 * d_model), nn.ELU(), _Linear(2 * d_model, d_model * cd))
        m.norm3 = nn.LayerNorm(d_model)
    elif seq_len_src_cross is not None:
        m.multihead_attn = _Attn(d_model, nhead, seq_len_src_cross, relative)
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
        if transformer_type not in ('relative', 'absolute'):
            raise NotImplementedError(f'transformer_type={transformer_type!r}')
        if transformer_type == 'absolute' and cross_attention_type == 'diagonal':
            raise NotImplementedError("the aligned layer is only used by the relative transformers (decoder.py:159-174)")
        assert encoder_attention_type in ('anticausal', 'causal', 'full')
        if cross_attention_type not in ('anticausal', 'full', 'diagonal'):
            raise NotImplementedError(f'cross_attention_type={cross_attention_type!r} (unimplemented in the reference as well, '
                                      'decoder.py:489-493)')
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
        rel = transformer_type == 'relative'
        if rel:
            self.target_channel_embeddings = nn.Parameter(torch.randn((1, self.num_channels, positional_embedding_size)))
            self.num_events_per_code = self.total_upscaling // self.num_channels
            self.target_events_positioning_embeddings = nn.Parameter(torch.randn((1, self.num_events_per_code, positional_embedding_size)))
        else:                                  # one learnt vector per absolute source / target position (decoder.py:88-102)
            self.source_positional_embeddings = nn.Parameter(torch.randn((1, self.num_tokens_target // self.total_upscaling,
                                                                          positional_embedding_size)))
            self.target_positional_embeddings = nn.Parameter(torch.randn((1, self.num_tokens_target, positional_embedding_size)))
        S = num_channels_encoder * num_events_encoder
        T = num_channels_decoder * num_events_decoder
        self.seq_len_src, self.seq_len_tgt = S, T
        tr = nn.Module()
        tr.encoder = nn.Module()
        tr.encoder.layers = nn.ModuleList([_layer(d_model, n_head, dim_feedforward, S, relative=rel) for _ in range(num_encoder_layers)])
        tr.decoder = nn.Module()
        self.num_channels_encoder, self.num_channels_decoder = num_channels_encoder, num_channels_decoder
        aligned = (num_channels_encoder, num_channels_decoder) if cross_attention_type == 'diagonal' else None
        tr.decoder.layers = nn.ModuleList([_layer(d_model, n_head, dim_feedforward, T, S, aligned=aligned, relative=rel)
                                           for _ in range(num_decoder_layers)])
        self.transformer = tr
        self.linear_target = _Linear(data_processor.embedding_size + positional_embedding_size * (2 if rel else 1), d_model)
        self.sos = nn.Parameter(torch.randn((1, 1, d_model)))
        q = encoder.quantizer
        src_dim = d_model if rel else d_model - positional_embedding_size                        # :226-229
        if type(q).__name__ != 'NoQuantization':
            self.source_embeddings = nn.Embedding(q.codebook_size ** q.num_codebooks, src_dim)
        else:
            self.source_embeddings = _Linear(q.codebook_dim, src_dim)

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
    def _attention(self, a, x_q, x_kv, B, Lq, Lk, mask, kv_sink=None):
        d, H = self.d_model, self.n_head
        hd = d // H
        scale = float(hd) ** -0.5
        w, b = a.in_proj_weight, a.in_proj_bias
        if x_kv is None:                      # self-attention: one fused projection, q third scaled after the bias (:247)
            qkv = ops.linear_fwd(x_q, w, b, scale_cols=(scale, d))
            q, k, v, ldq, ldkv = qkv, qkv[:, d:], qkv[:, 2 * d:], 3 * d, 3 * d
            if kv_sink is not None:               # prefill of the generation cache (generate_from_code_long)
                kv_sink[0].copy_(k[:, :d])
                kv_sink[1].copy_(v[:, :d])
        else:                                 # q from the target, k / v from the memory through slices of the same matrix
            q = ops.linear_fwd(x_q, w[:d], b[:d], scale_cols=(scale, d))
            kv = ops.linear_fwd(x_kv, w[d:], b[d:])
            k, v, ldq, ldkv = kv, kv[:, d:], d, 2 * d
        out = torch.empty(B * Lq, d, dtype=torch.float32, device=x_q.device)
        e1, e2 = a.tables()
        if x_kv is None and Lq <= 1024 and mask in (None, 'causal', 'anticausal'):
            # self-attention: the tiled full-sequence kernel (closed-form skewed bias = the maps for T == S, csrc/attn_long.cu)
            mode = {None: 0, 'causal': 1, 'anticausal': 2}[mask]
            _native.call('vqcpc_attn_long_fwd_masked', _p(qkv), _p(e1), _p(e2), _p(out), None, B, Lq, H, hd, mode, ops._stream())
            return out
        if isinstance(mask, str):
            mask = _causal_mask(Lk, x_q.device) if mask == 'causal' else _anticausal_mask(Lk, x_q.device)
        m1, m2 = device_maps(Lq, Lk, x_q.device) if e1 is not None else (None, None)
        _native.call('vqcpc_attn_map_fwd', _p(q), ldq, _p(k), _p(v), ldkv, _p(e1), _p(e2), _p(m1), _p(m2), _p(mask), _p(out), None, B,
                     Lq, Lk, H, hd, ops._stream())
        return out

    def _enc_layer(self, lyr, x, B, L, mask):
        """transformer_custom.py:268-291 (post-LN)."""
        o = self._attention(lyr.self_attn, x, None, B, L, L, mask)
        y = ops.linear_fwd(o, lyr.self_attn.out_proj.weight, lyr.self_attn.out_proj.bias, residual=x)
        x = ops.layernorm_fwd(y, lyr.norm1.weight, lyr.norm1.bias)[0]
        hmid = ops.linear_fwd(x, lyr.linear1.weight, lyr.linear1.bias, relu=True)
        y = ops.linear_fwd(hmid, lyr.linear2.weight, lyr.linear2.bias, residual=x)
        return ops.layernorm_fwd(y, lyr.norm2.weight, lyr.norm2.bias)[0]

    def _aligned_cross(self, lyr, memory, B, T, S):
        """transformer_custom.py:451-471: (B*S, d) memory -> (B, T, d) cross term of the aligned layer: the MLP output of memory
        event em, viewed (d, Cd), feeds channel c of every target event aligned with em."""
        d, ce, cd = self.d_model, self.num_channels_encoder, self.num_channels_decoder
        em = S // ce
        l0, l2 = lyr.cross_attn[0], lyr.cross_attn[2]
        h = ops.linear_fwd(memory.reshape(B * em, ce * d), l0.weight, l0.bias)
        _native.call('vqcpc_elu_fwd', _p(h), _p(h), h.numel(), ops._stream())
        o = ops.linear_fwd(h, l2.weight, l2.bias).view(B, em, d, cd).permute(0, 1, 3, 2)          # (B, em, Cd, d)
        rep = (T // cd) // em
        return o.unsqueeze(2).expand(B, em, rep, cd, d).reshape(B * T, d).contiguous()

    def _dec_layer(self, lyr, x, memory, B, T, S, tgt_mask, memory_mask, kv_sink=None):
        """transformer_custom.py:355-384: self-attention, encoder-decoder attention (or the aligned MLP term, :389-479),
        feed-forward, three post-LNs."""
        o = self._attention(lyr.self_attn, x, None, B, T, T, tgt_mask, kv_sink=kv_sink)
        y = ops.linear_fwd(o, lyr.self_attn.out_proj.weight, lyr.self_attn.out_proj.bias, residual=x)
        x = ops.layernorm_fwd(y, lyr.norm1.weight, lyr.norm1.bias)[0]
        if self.cross_attention_type == 'diagonal':
            y = x + self._aligned_cross(lyr, memory, B, T, S)
        else:
            o = self._attention(lyr.multihead_attn, x, memory, B, T, S, memory_mask)
            y = ops.linear_fwd(o, lyr.multihead_attn.out_proj.weight, lyr.multihead_attn.out_proj.bias, residual=x)
        x = ops.layernorm_fwd(y, lyr.norm2.weight, lyr.norm2.bias)[0]
        hmid = ops.linear_fwd(x, lyr.linear1.weight, lyr.linear1.bias, relu=True)
        y = ops.linear_fwd(hmid, lyr.linear2.weight, lyr.linear2.bias, residual=x)
        return ops.layernorm_fwd(y, lyr.norm3.weight, lyr.norm3.bias)[0]

    @torch.no_grad()
    def _encode_memory(self, source):
        """Code embedding (:440) + the encoder stack over the code sequence (:498-505, :509).  -> (memory (B*S, d), S)."""
        B = source.size(0)
        dev = self.sos.device
        d = self.d_model
        if isinstance(self.source_embeddings, nn.Embedding):
            src = torch.nn.functional.embedding(source.to(dev).long(), self.source_embeddings.weight)                 # :440
            src = src.reshape(-1, src.shape[-1])
        else:
            src = ops.linear_fwd(ops._f32c(source.to(dev)).reshape(-1, source.shape[-1]), self.source_embeddings.weight,
                                 self.source_embeddings.bias)
        S = src.shape[0] // B
        if self.transformer_type == 'absolute':                                                          # :458-461
            src = torch.cat([src.view(B, S, -1), self.source_positional_embeddings.expand(B, S, -1)], dim=2).reshape(B * S, d)
        source_mask = None if self.encoder_attention_type == 'full' else self.encoder_attention_type    # 'causal' | 'anticausal'
        x = src.contiguous()
        for lyr in self.transformer.encoder.layers:
            x = self._enc_layer(lyr, x, B, S, source_mask)
        return x, S

    @torch.no_grad()
    def forward(self, source, target):
        """source: code indices (batch, S) (or (batch, S, codebook_dim) without quantization); target: tokens (batch, num_events,
        num_channels).  Returns the reference's dict: loss, weights_per_category (list of (batch, num_events, vocab_c)),
        monitored_quantities.  decoder.py:431-543."""
        memory, S = self._encode_memory(source)
        return self._forward_from_memory(memory, S, target)

    def _forward_from_memory(self, memory, S, target, kv_sinks=None, want_loss=True):
        B = target.size(0)
        dev = self.sos.device
        d, C = self.d_model, self.num_channels
        target = self.data_processor.preprocess(target)
        tgt = flatten(self.data_processor.embed(target))                                                 # (B, T, E)   :442-444
        T = tgt.size(1)
        if self.transformer_type == 'absolute':
            tgt = torch.cat([tgt, self.target_positional_embeddings.expand(B, T, -1)], dim=2).reshape(B * T, -1)      # :462-466
        else:
            chan = self.target_channel_embeddings.repeat(B, T // C, 1)                                   # :451-453
            evt = self.target_events_positioning_embeddings.repeat_interleave(C, dim=1).repeat(B, T // self.total_upscaling, 1)
            tgt = torch.cat([tgt, chan, evt], dim=2).reshape(B * T, -1)                                  # :449-457
        tgt = ops.linear_fwd(tgt.contiguous(), self.linear_target.weight, self.linear_target.bias).view(B, T, d)   # :469
        tgt = torch.cat([self.sos.expand(B, 1, d), tgt[:, :-1]], dim=1).reshape(B * T, d).contiguous()   # shift by one   :476-482
        memory_mask = _anticausal_mask(S, dev, T) if self.cross_attention_type == 'anticausal' else None  # :487-496
        target_mask = 'causal'                                                                           # :507
        y = tgt
        for li, lyr in enumerate(self.transformer.decoder.layers):
            y = self._dec_layer(lyr, y, memory, B, T, S, target_mask, memory_mask, kv_sink=None if kv_sinks is None else kv_sinks[li])
        if not want_loss:
            return None
        out = y.view(B * (T // C), C * d)
        weights_per_category = [ops.linear_fwd(out[:, c * d:(c + 1) * d], ps.weight, ps.bias).view(B, T // C, -1)
                                for c, ps in enumerate(self.pre_softmaxes)]                              # :519-526
        loss = 0
        for c, w in enumerate(weights_per_category):                                                     # :529-535
            l, _ = ops.softmax_ce(w.reshape(B * (T // C), -1), target_idx=target[:, :, c].reshape(-1).contiguous(), want_grad=False)
            loss = loss + l
        loss = loss.mean()
        return {'loss': loss, 'weights_per_category': weights_per_category, 'monitored_quantities': {'loss': loss.item()}}

    @torch.no_grad()
    def epoch(self, data_loader, train=False, num_batches=None):
        """Evaluation pass of decoder.py:310-370 (``train=False``): encode every window, teacher-forced forward, mean of the
        monitored quantities.  The training branch (backward through the encoder-decoder transformer) is not built."""
        if train:
            raise NotImplementedError('decoder training is outside the built path (inference: forward, epoch(train=False), generate)')
        from itertools import islice
        self.eval()
        means, n = None, 0
        for tensor_dict in islice(data_loader, num_batches):
            x = tensor_dict['x']
            zs, encoding_indices, _ = self.encoder(x)
            codes = zs if encoding_indices is None else self.encoder.merge_codes(encoding_indices)
            mq = self.forward(codes, x)['monitored_quantities']
            means = dict(mq) if means is None else {k: v + means[k] for k, v in mq.items()}
            n += 1
        return {k: v / n for k, v in means.items()}

    def compute_start_end_times(self, t, num_blocks, num_blocks_model):
        """decoder.py:831-854: the window of ``num_blocks_model`` codes used for code t of a sequence of ``num_blocks`` codes and the
        position of t inside it (centred except near the two ends)."""
        half = num_blocks_model // 2
        if half <= t < num_blocks - half:
            t_relative = half
        elif t < half:
            t_relative = t
        else:
            t_relative = num_blocks_model - (num_blocks - t)
        t_begin = min(max(0, t - half), num_blocks - num_blocks_model)
        return t_begin, t_begin + num_blocks_model, t_relative

    def init_generation_chorale(self, num_events, start_index):
        """decoder.py:1054-1062: PAD everywhere, START at event start_index - 1 (symbols from the dataset's dictionaries; zeros when
        the loader has none, e.g. the synthetic one)."""
        ds = getattr(self.dataloader_generator, 'dataset', None)
        dicts = getattr(ds, 'note2index_dicts', None)
        x = torch.zeros(1, num_events, self.num_channels, dtype=torch.int64, device=self.sos.device)
        if dicts is not None and all('XX' in d_ and 'START' in d_ for d_ in dicts):
            x[:] = torch.tensor([d_['XX'] for d_ in dicts], device=x.device)
            if start_index >= 1:
                x[0, start_index - 1] = torch.tensor([d_['START'] for d_ in dicts], device=x.device)
        return x

    @torch.no_grad()
    def generate_from_code_long(self, encoding_indices, temperature=1.0, top_k=0, top_p=1., num_decodings=1, code_index_start=None,
                                code_index_end=None, greedy=False, kv_cache=True, uniforms=None, generator=None, init_chorale=None):
        """decoder.py:729-829 without the score writing: decode a code sequence LONGER than the model's window, code by code, each
        from the window ``compute_start_end_times`` centres on it.  Returns tokens (batch * num_decodings, events, channels) of
        the codes [code_index_start, code_index_end).

        The reference runs one full forward of the window per token.  Here (``kv_cache=True``) the window is fixed for the
        ``total_upscaling`` tokens of a code, so per code ONE teacher-forced forward of the window refills the key / value cache
        (its QKV products are the cache rows) and the code's tokens are generated incrementally from it (csrc/decode.cu).
        ``uniforms``: (tokens generated, batch) float64, consumed in generation order like the reference's ``np.random.choice``."""
        self.eval()
        dev = self.sos.device
        size_encoding = encoding_indices.size(1)
        up, C = self.total_upscaling, self.num_channels
        epc = up // C                                        # events per code
        num_codes_model = self.num_tokens_target // up
        code_index_start = 0 if code_index_start is None else code_index_start
        code_index_end = size_encoding if code_index_end is None else code_index_end
        assert size_encoding >= num_codes_model and 0 <= code_index_start < code_index_end <= size_encoding
        codes = encoding_indices.to(dev).repeat_interleave(num_decodings, dim=0)
        B = codes.size(0)
        assert B <= self._max_rows() or not kv_cache, f'at most {self._max_rows()} sequences per call on the cached path'
        chorale = (self.init_generation_chorale(size_encoding * epc, code_index_start * epc) if init_chorale is None
                   else init_chorale.to(dev).long()).repeat(B, 1, 1).contiguous()
        n_tok = (code_index_end - code_index_start) * up
        if uniforms is None and not greedy:
            uniforms = torch.rand(n_tok, B, dtype=torch.float64, device=dev, generator=generator)
        u = None if uniforms is None else torch.as_tensor(uniforms, dtype=torch.float64).to(dev)
        sess = self._decode_session(B, temperature, top_k, top_p, greedy, None, True) if kv_cache else None
        done = 0
        prev = None                                          # (t_begin, next position) of the window the cache currently holds
        for code_index in range(code_index_start, code_index_end):
            t_begin, t_end, t_rel = self.compute_start_end_times(code_index, size_encoding, num_codes_model)
            win_codes = codes[:, t_begin:t_end]
            window = chorale[:, t_begin * epc:t_end * epc]
            p0 = t_rel * up
            if kv_cache:
                # while the window does not move (both ends of the sequence) the cache simply continues
                new = sess.run_window(win_codes, window, p0, up, None if u is None else u[done:done + up],
                                      prefill=prev != (t_begin, p0))
                prev = (t_begin, p0 + up)
                window[:, t_rel * epc:(t_rel + 1) * epc] = new.view(B, epc, C)
            else:
                zero = torch.zeros(1, dtype=torch.int32, device=dev)
                for k in range(up):                          # the reference's schedule: a full forward of the window per token
                    e, c = divmod(p0 + k, C)
                    logits = self.forward(win_codes, window)['weights_per_category'][c][:, e, :].contiguous()
                    V = logits.shape[1]
                    drawn = torch.empty(B, 1, dtype=torch.int64, device=dev)
                    _native.call('vqcpc_sample_step', _p(logits), V, V, float(temperature), int(top_k), float(top_p),
                                 None if u is None else _p(u[done + k]), B, _p(drawn), 1, _p(zero), int(bool(greedy)), B, ops._stream())
                    window[:, e, c] = drawn[:, 0]
            done += up
        return chorale[:, code_index_start * epc:code_index_end * epc]

    def _symbols(self):
        """PAD / START / END token ids per channel from the dataset's dictionaries (zeros when the loader has none)."""
        ds = getattr(self.dataloader_generator, 'dataset', None)
        dicts = getattr(ds, 'note2index_dicts', None)
        dev = self.sos.device
        if dicts is not None and all(all(k in d_ for k in ('XX', 'START', 'END')) for d_ in dicts):
            return tuple(torch.tensor([d_[k] for d_ in dicts], device=dev).view(1, 1, -1) for k in ('XX', 'START', 'END'))
        z = torch.zeros(1, 1, self.num_channels, dtype=torch.int64, device=dev)
        return z, z.clone(), z.clone()

    @torch.no_grad()
    def generate_alla_mano(self, start_codes, end_codes, body_codes, temperature, num_decodings=3, **sampling):
        """decoder.py:960-981 without the score writing: decode a hand-written code sequence; tokens of the body codes."""
        seq = list(start_codes) + list(body_codes)
        code_index_start, code_index_end = len(start_codes), len(seq)
        codes = torch.tensor(seq + list(end_codes), dtype=torch.int64, device=self.sos.device).unsqueeze(0)
        return self.generate_from_code_long(codes, temperature=temperature, num_decodings=num_decodings,
                                            code_index_start=code_index_start, code_index_end=code_index_end, **sampling)

    @torch.no_grad()
    def generate_reharmonisation(self, num_reharmonisations, temperature, top_k, top_p, x=None, **sampling):
        """decoder.py:856-957 from a token tensor: ``x`` (1, events, channels) is cut into windows, framed by a [PAD.. START] window
        and an END / PAD tail (:880-914), encoded window by window, the codes glued (:916-929) and decoded
        ``num_reharmonisations`` times between the frame codes (:930-944).  The reference reads BWV 251 through music21 and
        writes scores; here the caller supplies the tokens and gets tokens back."""
        if x is None:
            raise NotImplementedError('pass the chorale as tokens x (1, events, channels): reading the corpus needs music21')
        dev = self.sos.device
        x = x.to(dev).long()
        n = self.data_processor.num_events
        PAD, START, END = self._symbols()
        chunks = list(x.split(n, 1))
        last = chunks[-1]
        completion = n - last.size(1)
        start_chunk = torch.cat([PAD.repeat(1, n - 1, 1), START], 1)
        if completion > 1:
            last, end_chunk = torch.cat([last, END, PAD.repeat(1, completion - 1, 1)], 1), PAD.repeat(1, n, 1)
        elif completion == 1:
            last, end_chunk = torch.cat([last, END], 1), PAD.repeat(1, n, 1)
        else:
            end_chunk = torch.cat([END, PAD.repeat(1, n - 1, 1)], 1)
        chunks[-1] = last
        x_chunks = torch.cat([start_chunk] + chunks + [end_chunk], dim=0)
        zs, stack, _ = self.encoder(x_chunks)
        codes = zs if stack is None else self.encoder.merge_codes(stack)
        codes = torch.cat(codes.split(1, 0), 1)                                                     # glue the windows
        up = self.total_upscaling
        code_index_start = n * self.num_channels // up
        code_index_end = codes.size(1) - (n + completion) * self.num_channels // up
        return self.generate_from_code_long(codes, temperature=temperature, top_k=top_k, top_p=top_p,
                                            num_decodings=num_reharmonisations, code_index_start=code_index_start,
                                            code_index_end=code_index_end, **sampling)

    def check_duplicate(self, generation, original):
        """decoder.py:983-991: longest common run of the two token dumps -> (matching string, number of tokens)."""
        from difflib import SequenceMatcher
        s1, s2 = self.data_processor.dump(generation), self.data_processor.dump(original)
        m = SequenceMatcher(None, s1, s2).find_longest_match(0, len(s1), 0, len(s2))
        return s1[m.a:m.a + m.size], (m.size - 1) / 3

    def check_duplicate_all_corpus(self, generation, max_windows=None):
        """decoder.py:993-1017: the training window sharing the longest token run with ``generation`` -> (window, tokens)."""
        from difflib import SequenceMatcher
        from itertools import islice
        s1 = self.data_processor.dump(generation)
        train, _, _ = self.dataloader_generator.dataloaders(batch_size=1, shuffle_val=True, shuffle_train=False)
        best_x, best = None, 0
        for tensor_dict in islice(train, max_windows):
            x = tensor_dict['x']
            s2 = self.data_processor.dump(x[0])
            m = SequenceMatcher(None, s1, s2, autojunk=False).find_longest_match(0, len(s1), 0, len(s2))
            if m.size > best:
                best_x, best = x, m.size
        return best_x, (best - 1) / 3

    def init_generation(self, num_events):
        return torch.zeros(1, num_events, self.num_channels, dtype=torch.int64, device=self.sos.device)   # :724-727

    @torch.no_grad()
    def generate_from_codes(self, encoding_indices, temperature=1.0, top_k=0, top_p=1.0, generator=None, greedy=False,
                            kv_cache=True, uniforms=None, use_graph=True, exclude=None):
        """The token loop of ``generate`` (decoder.py:607-644).  Returns tokens (batch, num_events, num_channels).

        kv_cache=True (default): incremental decoding - one new row per layer and token over cached keys / values, replayed as a
        CUDA graph per event, sampling on the device (csrc/decode.cu; exact because the decoder is causal and the relative bias of
        row i only involves q_i).  ``uniforms`` (num_tokens, batch) float64 in [0, 1): the draws, consumed in the order the
        reference calls ``np.random.choice`` (token by token, batch index inner); default: drawn from ``generator``.
        ``exclude``: {channel: [token ids]} never sampled (the reference's ``exclude_meta_symbols``, :624-629).

        kv_cache=False: the reference's schedule - a full teacher-forced forward for every token (kept as the parity check of
        the cached path), filtering with ``top_k_top_p_filtering`` and ``torch.multinomial`` on the device."""
        self.eval()
        if kv_cache:
            B = encoding_indices.size(0)
            R = self._max_rows()
            outs = [self._generate_cached(encoding_indices[b0:b0 + R], temperature, top_k, top_p, generator, greedy,
                                          None if uniforms is None else uniforms[:, b0:b0 + R], use_graph, exclude)
                    for b0 in range(0, B, R)]
            return torch.cat(outs, dim=0)
        B = encoding_indices.size(0)
        x = self.init_generation(self.data_processor.num_events).repeat(B, 1, 1)
        for event_index in range(self.data_processor.num_events):
            for channel_index in range(self.num_channels):
                logits = self.forward(encoding_indices, x)['weights_per_category'][channel_index][:, event_index, :] / temperature
                if exclude and channel_index in exclude:
                    logits[:, list(exclude[channel_index])] = -float('inf')
                if greedy:
                    new = logits.argmax(dim=-1)
                else:
                    filtered = top_k_top_p_filtering(logits, top_k=top_k, top_p=top_p)
                    new = torch.multinomial(torch.softmax(filtered, dim=-1), 1, generator=generator).squeeze(1)
                x[:, event_index, channel_index] = new
        return x

    def _max_rows(self):
        """Sequences per decode session: LayerNorm-on-load keeps 32 rows in registers only up to d_model = 512."""
        return _MAX_ROWS if self.d_model <= 512 else 16

    def _decode_session(self, B, temperature, top_k, top_p, greedy, exclude, use_graph):
        ex_key = None if not exclude else tuple(sorted((c, tuple(sorted(v))) for c, v in exclude.items()))
        key = (B, float(temperature), int(top_k), float(top_p), bool(greedy), ex_key, bool(use_graph), self.sos.data_ptr(),
               ops.weight_version())
        cache = self.__dict__.setdefault('_decode_sessions', {})
        if key not in cache:
            if len(cache) >= 4:
                cache.clear()
            cache[key] = _DecodeSession(self, B, temperature, top_k, top_p, greedy, exclude, use_graph)
        return cache[key]

    def _generate_cached(self, codes, temperature, top_k, top_p, generator, greedy, uniforms, use_graph, exclude):
        """<= _MAX_ROWS sequences through the incremental path.  The buffers, the launch sequence and its CUDA graph live in a
        ``_DecodeSession`` cached per (batch, sampling settings): a later call only refreshes the memory keys / values and the
        uniforms and replays."""
        dev = self.sos.device
        B, T = codes.size(0), self.num_tokens_target
        sess = self._decode_session(B, temperature, top_k, top_p, greedy, exclude, use_graph)
        if uniforms is None:
            uniforms = torch.rand(T, B, dtype=torch.float64, device=dev, generator=generator)
        u = torch.as_tensor(uniforms, dtype=torch.float64).to(dev)
        assert u.shape == (T, B)
        tokens = sess.run(codes, u)
        self.last_generation_launches = sess.launches_per_window
        return tokens


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
        exclude = None
        if kwargs.get('exclude_meta_symbols'):               # :624-629: START / END / PAD are never drawn
            dicts = getattr(getattr(self.dataloader_generator, 'dataset', None), 'note2index_dicts', None)
            if dicts is not None:
                exclude = {c: [d_[k] for k in ('START', 'END', 'XX') if k in d_] for c, d_ in enumerate(dicts)}
        x = self.generate_from_codes(encoding_indices, temperature, top_k=top_k, top_p=top_p, exclude=exclude)
        _, recoding, _ = self.encoder(torch.cat([tensor_dict['x'].to(x.device).long(), x], dim=0))
        recoding = None if recoding is None else self.encoder.merge_codes(recoding)
        return x_original, x, recoding


class _DecodeSession:
    """Static buffers + launch sequence (+ CUDA graph of one event) of the incremental decoder for a fixed batch and fixed
    sampling settings; see csrc/decode.cu for the step.  Holds raw pointers into the decoder's parameters: the session key
    carries ``ops.weight_version()`` and a parameter address, so re-loaded / moved weights build a fresh session."""

    def __init__(self, dec, B, temperature, top_k, top_p, greedy, exclude, use_graph):
        self.dec, self.B, self.use_graph = dec, B, use_graph
        dev = dec.sos.device
        d, H, C = dec.d_model, dec.n_head, dec.num_channels
        hd = d // H
        scale = float(hd) ** -0.5
        T, up = dec.num_tokens_target, dec.total_upscaling
        S = dec.seq_len_src
        self.T, self.C, self.Ev, self.S = T, C, T // C, S
        layers = list(dec.transformer.decoder.layers)
        self.layers = layers
        f32 = dict(dtype=torch.float32, device=dev)
        self.aligned = dec.cross_attention_type == 'diagonal'
        self.cross_kv = [] if self.aligned else [torch.empty(B * S, 2 * d, **f32) for _ in layers]
        self.cross_add = [torch.empty(B * T, d, **f32) for _ in layers] if self.aligned else []
        kc = [torch.zeros(B * T, d, **f32) for _ in layers]
        vc = [torch.zeros(B * T, d, **f32) for _ in layers]
        sm1, sm2 = (None, None) if dec.transformer_type == 'absolute' else device_maps(T, T, dev)
        cm1, cm2 = (None, None) if (self.aligned or dec.transformer_type == 'absolute') else device_maps(T, S, dev)
        memory_mask = _anticausal_mask(S, dev, T) if dec.cross_attention_type == 'anticausal' else None
        tables = torch.cat([e.weight.detach() for e in dec.data_processor.embeddings], dim=0).contiguous()
        rows = [e.weight.shape[0] for e in dec.data_processor.embeddings]
        table_off = torch.tensor([sum(rows[:c]) for c in range(C)], dtype=torch.int32, device=dev)
        E = tables.shape[1]
        absolute = dec.transformer_type == 'absolute'
        pos_mode = int(absolute)
        if absolute:                           # [embedding | positional embedding of the absolute position]
            pe = dec.target_positional_embeddings.shape[-1]
            F = E + pe
            chan = dec.target_positional_embeddings.detach().reshape(T, pe).contiguous()
            evt = chan
        else:
            pe = dec.target_channel_embeddings.shape[-1]
            F = E + 2 * pe
            chan = dec.target_channel_embeddings.detach().reshape(C, pe).contiguous()
            evt = dec.target_events_positioning_embeddings.detach().reshape(-1, pe).contiguous()
        ff = layers[0].linear1.weight.shape[0]
        Vmax = max(dec.num_tokens_per_channel)
        self.tokens = tokens = torch.zeros(B, T, dtype=torch.int64, device=dev)
        self.pos = pos = torch.zeros(1, dtype=torch.int32, device=dev)
        self.u = u = torch.zeros(T, B, dtype=torch.float64, device=dev)
        feat, x, x1, x2, y, att, qc = (torch.zeros(B, n, **f32) for n in (F, d, d, d, d, d, d))
        qkv, hmid, logits = torch.empty(B, 3 * d, **f32), torch.empty(B, ff, **f32), torch.empty(B, Vmax, **f32)
        done = torch.zeros(1, dtype=torch.int32, device=dev)
        head_bias = []
        for c, ps in enumerate(dec.pre_softmaxes):
            bvec = ps.bias.detach().clone()
            if exclude and c in exclude:
                bvec[list(exclude[c])] = -float('inf')
            head_bias.append(bvec)
        sos = dec.sos.detach().reshape(d).contiguous()
        self.kc, self.vc = kc, vc                            # self-attention key / value caches, one (B * T, d) pair per layer
        self._keep = [kc, vc, sm1, sm2, cm1, cm2, memory_mask, tables, table_off, chan, evt, head_bias, sos, feat, x, x1, x2, y,
                      att, qc, qkv, hmid, logits, done]
        call = _native.call

        def lin(xin, ldx, w, b, res, out, ldo, N, K, relu=0, sc=1.0, sc_cols=0, row0=None):
            call('vqcpc_linear_small', _p(xin), ldx, _p(w), _p(b), _p(res), d, _p(out), ldo, B, N, K, relu, sc, sc_cols, _p(row0),
                 _p(pos), ops._stream())

        def ln_lin(xin, norm, ln_out, w, b, out, ldo, N, relu=0, sc=1.0, sc_cols=0):
            """out = linear(LayerNorm(xin)); the normalised rows also go to ln_out (the residual of a later linear)."""
            call('vqcpc_ln_linear_small', _p(xin), d, _p(norm.weight), _p(norm.bias), float(norm.eps), _p(ln_out), _p(w), _p(b), None, d,
                 _p(out), ldo, B, N, d, relu, sc, sc_cols, ops._stream())

        def step(c):
            st = ops._stream()
            lin(feat, F, dec.linear_target.weight, dec.linear_target.bias, None, x, d, d, F, row0=sos)
            for li, l in enumerate(layers):
                a = l.self_attn
                if li == 0:
                    lin(x, d, a.in_proj_weight, a.in_proj_bias, None, qkv, 3 * d, 3 * d, d, sc=scale, sc_cols=d)
                else:                                        # x = norm3 of the previous layer, applied while staging
                    ln_lin(y, layers[li - 1].norm3, x, a.in_proj_weight, a.in_proj_bias, qkv, 3 * d, 3 * d, sc=scale, sc_cols=d)
                se1, se2 = a.tables()
                call('vqcpc_attn_decode', _p(qkv), 3 * d, _p(qkv[:, d:]), _p(qkv[:, 2 * d:]), 3 * d, _p(kc[li]), _p(vc[li]), d, T,
                     _p(se1), _p(se2), _p(sm1), _p(sm2), None, T, _p(pos), 1, _p(att), d, B, H, hd, st)
                lin(att, d, a.out_proj.weight, a.out_proj.bias, x, y, d, d, d)
                if self.aligned:                             # y2 = norm1(y) + the aligned memory term of this position
                    call('vqcpc_decode_ln_add', _p(y), _p(l.norm1.weight), _p(l.norm1.bias), float(l.norm1.eps),
                         _p(self.cross_add[li]), T, _p(pos), _p(x1), B, d, st)
                    y_mid = x1
                else:
                    a = l.multihead_attn
                    ln_lin(y, l.norm1, x1, a.in_proj_weight[:d], a.in_proj_bias[:d], qc, d, d, sc=scale, sc_cols=d)
                    ckv = self.cross_kv[li]
                    ce1, ce2 = a.tables()
                    call('vqcpc_attn_decode', _p(qc), d, None, None, 0, _p(ckv), _p(ckv[:, d:]), 2 * d, S, _p(ce1), _p(ce2), _p(cm1),
                         _p(cm2), _p(memory_mask), S, _p(pos), 0, _p(att), d, B, H, hd, st)
                    lin(att, d, a.out_proj.weight, a.out_proj.bias, x1, y, d, d, d)
                    y_mid = y
                ln_lin(y_mid, l.norm2, x2, l.linear1.weight, l.linear1.bias, hmid, ff, ff, relu=1)
                lin(hmid, ff, l.linear2.weight, l.linear2.bias, x2, y, d, d, ff)
            V = dec.num_tokens_per_channel[c]
            ln_lin(y, layers[-1].norm3, None, dec.pre_softmaxes[c].weight, head_bias[c], logits, Vmax, V)
            # draw the token, write the feature row of the next position, advance the position
            call('vqcpc_decode_sample_embed', _p(logits), Vmax, V, float(temperature), int(top_k), float(top_p), _p(u), B, _p(tokens), T,
                 _p(pos), int(bool(greedy)), B, _p(tables), _p(table_off), _p(chan), _p(evt), _p(feat), C, up, E, pe, _p(done),
                 pos_mode, st)

        def event():
            for c in range(C):
                step(c)

        def features_now():                                  # feature row of the CURRENT position (a window entered mid-way)
            call('vqcpc_decode_features', _p(tokens), _p(pos), _p(tables), _p(table_off), _p(chan), _p(evt), _p(feat), B, T, C, up, E,
                 pe, pos_mode, ops._stream())

        self._features_now = features_now
        self.event = event
        self.graph = None
        self.launches_per_window = T * (3 + (6 if self.aligned else 8) * len(layers))

    def run(self, codes, u):
        dec, d = self.dec, self.dec.d_model
        # programmatic dependent launch pays when the step is latency bound (measured on B200: batch 1 +6 %, batch 16 -25 %:
        # the early-scheduled successors take SM slots from the attention waves); VQCPC_DECODE_PDL=0/1 overrides
        pdl = os.environ.get('VQCPC_DECODE_PDL')
        _native.call('vqcpc_decode_set_pdl', int(pdl) if pdl is not None else int(self.B <= 4))
        self._project_memory(codes)
        self.u.copy_(u)
        self.tokens.zero_()
        self.pos.zero_()
        self._events(self.Ev)
        return self.tokens.view(self.B, self.Ev, self.C).clone()

    def _project_memory(self, codes):
        dec, d = self.dec, self.dec.d_model
        memory, S = dec._encode_memory(codes)
        assert S == self.S
        for ckv, l in zip(self.cross_kv, self.layers):      # keys / values of the memory (multihead_attention_custom.py:173-196)
            a = l.multihead_attn
            ops.linear_fwd(memory, a.in_proj_weight[d:], a.in_proj_bias[d:], out=ckv)
        for add, l in zip(self.cross_add, self.layers):     # aligned layers: the memory term of every target position
            add.copy_(dec._aligned_cross(l, memory, self.B, self.T, S))
        return memory

    def _events(self, n):
        """n events (C token steps each) from the current position: graph replays once the graph exists; the first event ever runs
        eagerly (the warm-up a capture needs)."""
        if not self.use_graph:
            for _ in range(n):
                self.event()
            return
        if self.graph is None:
            self.event()
            n -= 1
            torch.cuda.synchronize()
            pos0 = self.pos.clone()
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self.event()
            self.pos.copy_(pos0)                              # capture does not execute
        for _ in range(n):
            self.graph.replay()

    def run_window(self, codes, window_tokens, p0, n_tokens, u_rows, prefill=True):
        """generate_from_code_long: refill the cache with ONE teacher-forced forward of the window (its self-attention QKV products
        are the cache rows; rows >= p0 hold placeholders that the steps overwrite before any later row reads them), then draw
        the tokens p0 .. p0 + n_tokens - 1 incrementally.  window_tokens (B, events, C) int64; returns (B, n_tokens)."""
        dec, B, T, C = self.dec, self.B, self.T, self.C
        assert p0 % C == 0 and n_tokens % C == 0 and p0 + n_tokens <= T
        pdl = os.environ.get('VQCPC_DECODE_PDL')
        _native.call('vqcpc_decode_set_pdl', int(pdl) if pdl is not None else int(B <= 4))
        if prefill:
            memory = self._project_memory(codes)
            kc, vc = self.kc, self.vc
            dec._forward_from_memory(memory, self.S, window_tokens, kv_sinks=list(zip(kc, vc)), want_loss=False)
            self.tokens.copy_(window_tokens.reshape(B, T))
            self.pos.fill_(p0)
            self._features_now()
        # else: same window as the previous call and p0 is where it stopped - cache, tokens, position and feature row are current
        if u_rows is not None:
            self.u[p0:p0 + n_tokens].copy_(u_rows)
        self._events(n_tokens // C)
        return self.tokens[:, p0:p0 + n_tokens].clone()
