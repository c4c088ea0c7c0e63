"""CPU oracle of the decoder's inference path: ``Decoder.forward`` of decoders/decoder.py:431-543 for
``transformer_type='relative'``.

TEST INFRASTRUCTURE ONLY (see oracle/vqcpc_oracle.py).  Plain-PyTorch restatement over a flat parameter dict; pinned against the
live reference by ``oracle/make_golden_decoder.py`` -> ``tests/golden/decoder.pt``.  The relative bias is restated LITERALLY
(pad / view / slice on the float products, like subsampled_relative_attention.py:30-122) - independently of the integer-map
formulation the product uses (VQCPCB/transformer/relative_maps.py)."""
import math

import numpy as np
import torch
import torch.nn.functional as F

from oracle import params as P


def make_decoder_params(shapes, seed=0, dtype=torch.float32):
    """Deterministic values for a {state_dict key: shape} record (the keys come from the live reference module)."""
    prm = {}
    for name, shape in shapes.items():
        shape = tuple(shape)
        last = name.split('.')[-1]
        if 'norm' in name and last == 'weight':
            t = P._normal(name, seed, shape, 0.1, 1.0)
        elif 'norm' in name:
            t = P._normal(name, seed, shape, 0.1)
        elif last in ('e1', 'e2'):
            t = P._normal(name, seed, shape, 0.5)
        elif 'embeddings' in name or last == 'sos':
            t = P._normal(name, seed, shape, 1.0)
        elif len(shape) == 2:
            t = P._uniform(name, seed, shape, 1 / math.sqrt(shape[1]))
        else:
            t = P._uniform(name, seed, shape, 0.05)
        prm[name] = t.to(dtype)
    return prm


def rel_bias_general(q, e1, e2, T, S):
    """subsampled_relative_attention.py:30-122.  q (BH, T, hd) scaled; e1, e2 (H*S, hd); returns (BH, T, S)."""
    BH, _, hd = q.shape
    H = e1.shape[0] // S
    b = BH // H
    out = []
    for e, front in ((e1, True), (e2, False)):
        ee = e.unsqueeze(0).repeat(b, 1, 1).view(b * H, S, hd)
        rel = torch.einsum('bld,bmd->blm', q, ee).reshape(BH, S, T)                        # tgt x src viewed as src x tgt
        pad = torch.full((BH, S, 1), -100.0, dtype=q.dtype)
        rel = torch.cat([pad, rel], dim=2) if front else torch.cat([rel, pad], dim=2)
        if T != S:
            rel = torch.cat([rel, torch.full((BH, T - S, T + 1), -100.0, dtype=q.dtype)], dim=1)
        rel = rel.reshape(BH, -1, S)
        rel = rel[:, 1:][:, :T] if front else rel[:, :T]
        out.append(rel)
    ones = torch.ones(S, S, dtype=torch.bool)
    down = torch.triu(ones, diagonal=0).flip(0).flip(1)
    up = torch.triu(ones, diagonal=1)
    if T != S:
        down = torch.repeat_interleave(down, T // S, dim=0)
        up = torch.repeat_interleave(up, T // S, dim=0)
    return out[0].masked_fill(up, 0) + out[1].masked_fill(down, 0)


def mha(xq, xkv, prm, pre, H, mask):
    """multihead_attention_custom.py:154-353.  xq (B, Lq, d); xkv (B, Lk, d) or None (self-attention); mask (Lq, Lk) additive."""
    B, Lq, d = xq.shape
    hd = d // H
    w, b = prm[pre + 'in_proj_weight'], prm[pre + 'in_proj_bias']
    if xkv is None:
        q, k, v = F.linear(xq, w, b).chunk(3, dim=-1)
    else:
        q = F.linear(xq, w[:d], b[:d])
        k, v = F.linear(xkv, w[d:], b[d:]).chunk(2, dim=-1)
    Lk = k.shape[1]
    q = q * (float(hd) ** -0.5)
    qh = q.reshape(B, Lq, H, hd).transpose(1, 2).reshape(B * H, Lq, hd)
    kh = k.reshape(B, Lk, H, hd).transpose(1, 2).reshape(B * H, Lk, hd)
    vh = v.reshape(B, Lk, H, hd).transpose(1, 2).reshape(B * H, Lk, hd)
    s = torch.bmm(qh, kh.transpose(1, 2))
    if mask is not None:
        s = s + mask.unsqueeze(0)
    if pre + 'attn_bias.e1' in prm:                                                        # absent for transformer_type='absolute'
        s = s + rel_bias_general(qh, prm[pre + 'attn_bias.e1'], prm[pre + 'attn_bias.e2'], Lq, Lk)
    a = torch.softmax(s, dim=-1)
    o = torch.bmm(a, vh).reshape(B, H, Lq, hd).transpose(1, 2).reshape(B, Lq, d)
    return F.linear(o, prm[pre + 'out_proj.weight'], prm[pre + 'out_proj.bias'])


def _ln(x, prm, pre):
    return F.layer_norm(x, (x.shape[-1],), prm[pre + 'weight'], prm[pre + 'bias'], 1e-5)


def _ffn(x, prm, pre):
    return F.linear(torch.relu(F.linear(x, prm[pre + 'linear1.weight'], prm[pre + 'linear1.bias'])), prm[pre + 'linear2.weight'],
                    prm[pre + 'linear2.bias'])


def aligned_cross(memory, prm, pre, T, Cd, Ce):
    """transformer_custom.py:456-471 (TransformerAlignedDecoderLayerCustom): an MLP of the memory event a target position is
    aligned with replaces the cross-attention.  memory (B, S, d) -> (B, T, d)."""
    B, S, d = memory.shape
    Em = S // Ce
    m = memory.reshape(B, Em, Ce * d)
    h = F.elu(F.linear(m, prm[pre + 'cross_attn.0.weight'], prm[pre + 'cross_attn.0.bias']))
    o = F.linear(h, prm[pre + 'cross_attn.2.weight'], prm[pre + 'cross_attn.2.bias'])      # (B, Em, d * Cd)
    o = o.reshape(B, Em, d, Cd).permute(0, 1, 3, 2)                                        # (B, Em, Cd, d)
    rep = (T // Cd) // Em
    return o.repeat_interleave(rep, dim=1).reshape(B, T, d)


def causal_mask(n, dtype=torch.float32):
    i = torch.arange(n)
    return torch.where(i.view(1, n) <= i.view(n, 1), 0.0, float('-inf')).to(dtype)


def decoder_forward(source, target, prm, cfg):
    """decoder.py:431-543.  source (B, S) int64 codes; target (B, num_events, C) int64.  cfg: n_head, num_encoder_layers,
    num_decoder_layers, total_upscaling, encoder_attention_type, cross_attention_type.  Returns (loss, logits per channel)."""
    B, Ev, C = target.shape
    H = cfg['n_head']
    T = Ev * C
    if source.is_floating_point():                                                         # NoQuantization: z vectors (:237-239)
        src = F.linear(source, prm['source_embeddings.weight'], prm['source_embeddings.bias'])
    else:
        src = prm['source_embeddings.weight'][source]                                      # (B, S, d)
    S = src.shape[1]
    emb = torch.stack([prm[f'data_processor.embeddings.{c}.weight'][target[:, :, c]] for c in range(C)], dim=2)   # (B, Ev, C, E)
    tgt = emb.reshape(B, T, -1)
    if cfg.get('transformer_type', 'relative') == 'absolute':                              # decoder.py:458-466
        src = torch.cat([src, prm['source_positional_embeddings'].repeat(B, 1, 1)], dim=2)
        tgt = torch.cat([tgt, prm['target_positional_embeddings'].repeat(B, 1, 1)], dim=2)
    else:
        chan = prm['target_channel_embeddings'].repeat(B, T // C, 1)
        evt = prm['target_events_positioning_embeddings'].repeat_interleave(C, dim=1).repeat(B, T // cfg['total_upscaling'], 1)
        tgt = torch.cat([tgt, chan, evt], dim=2)
    tgt = F.linear(tgt, prm['linear_target.weight'], prm['linear_target.bias'])
    tgt = torch.cat([prm['sos'].expand(B, 1, -1), tgt[:, :-1]], dim=1)
    dt = tgt.dtype
    anti = causal_mask(S, dt).t()
    memory_mask = torch.repeat_interleave(anti, T // S, dim=0) if cfg['cross_attention_type'] == 'anticausal' else None
    source_mask = {'full': None, 'causal': causal_mask(S, dt), 'anticausal': anti}[cfg['encoder_attention_type']]
    x = src
    for l in range(cfg['num_encoder_layers']):
        p = f'transformer.encoder.layers.{l}.'
        x = _ln(x + mha(x, None, prm, p + 'self_attn.', H, source_mask), prm, p + 'norm1.')
        x = _ln(x + _ffn(x, prm, p), prm, p + 'norm2.')
    y = tgt
    for l in range(cfg['num_decoder_layers']):
        p = f'transformer.decoder.layers.{l}.'
        y = _ln(y + mha(y, None, prm, p + 'self_attn.', H, causal_mask(T, dt)), prm, p + 'norm1.')
        if cfg['cross_attention_type'] == 'diagonal':
            y = _ln(y + aligned_cross(x, prm, p, T, C, cfg.get('num_channels_encoder', 1)), prm, p + 'norm2.')
        else:
            y = _ln(y + mha(y, x, prm, p + 'multihead_attn.', H, memory_mask), prm, p + 'norm2.')
        y = _ln(y + _ffn(y, prm, p), prm, p + 'norm3.')
    y = y.reshape(B, Ev, C, -1)
    logits = [F.linear(y[:, :, c], prm[f'pre_softmaxes.{c}.weight'], prm[f'pre_softmaxes.{c}.bias']) for c in range(C)]
    loss = sum(F.cross_entropy(logits[c].reshape(B * Ev, -1), target[:, :, c].reshape(-1), reduction='none') for c in range(C)).mean()
    return loss, logits


def top_k_top_p_filtering(logits, top_k=0, top_p=0.0, filter_value=-float('inf')):
    """VQCPCB/utils.py:100-128 on one row of logits (returns a filtered copy)."""
    logits = logits.clone()
    top_k = min(top_k, logits.size(-1))
    if top_k > 0:
        logits[logits < torch.topk(logits, top_k)[0][..., -1, None]] = filter_value
    if top_p > 0.0:
        sorted_logits, sorted_indices = torch.sort(logits, descending=True)
        cumulative = torch.cumsum(torch.softmax(sorted_logits, dim=-1), dim=-1)
        remove = cumulative > top_p
        remove[..., 1:] = remove[..., :-1].clone()
        remove[..., 0] = False
        logits[sorted_indices[remove]] = filter_value
    return logits


def sample_rows(logits, temperature, top_k, top_p, rng):
    """decoder.py:619-644 for one token position: logits (B, V) -> B tokens, one ``rng.choice`` (numpy RandomState) per row in
    batch order - each consumes exactly one ``random_sample()`` of the stream."""
    out = []
    for row in logits.float() / temperature:
        p = torch.softmax(top_k_top_p_filtering(row, top_k=top_k, top_p=top_p), dim=-1).numpy()
        out.append(int(rng.choice(np.arange(p.shape[0]), p=p)))
    return out
