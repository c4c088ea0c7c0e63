"""Deterministic parameter dict + synthetic CPC batches for the oracle, the golden generator and the tests.

TEST INFRASTRUCTURE ONLY (see oracle/vqcpc_oracle.py header).  Keys follow the reference checkpoint files and
state_dict names: ``<file>.<state_dict key>`` with file in {data_processor, downscaler,
quantizer, upscaler, c_module, fks_module}.  Values come from one ``torch.Generator`` per tensor, so the dict
is reproducible on any box without the reference being importable.
"""
import math
import zlib

import torch


def _gen(name, seed):
    g = torch.Generator()
    g.manual_seed((zlib.crc32(name.encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFF)
    return g


def _uniform(name, seed, shape, bound):
    return (torch.rand(shape, generator=_gen(name, seed)) * 2 - 1) * bound


def _normal(name, seed, shape, std=1.0, mean=0.0):
    return torch.randn(shape, generator=_gen(name, seed)) * std + mean


def param_shapes(cfg):
    """name -> (shape, init kind).  Shapes as listed in SURVEY.md §8b 'Checkpoint keys'."""
    d, F, E, H = cfg.d_model, cfg.dim_feedforward, cfg.embedding_size, cfg.n_head
    hd = d // H
    V = cfg.num_voices
    S = int(math.prod(cfg.downscale_factors))
    pe = cfg.pos_emb
    out = {}
    for v in range(V):
        out[f'data_processor.embeddings.{v}.weight'] = ((cfg.vocab[v] + 1, E), ('normal', 1.0))
    if cfg.get('downscaler', 'transformer') == 'lstm':
        Hl = cfg.lstm_hidden
        bl = 1 / math.sqrt(Hl)
        for g in (['g_enc_fwd', 'g_enc_bwd'] if cfg.lstm_bidirectional else ['g_enc_fwd']):
            for l in range(cfg.lstm_layers):
                inp = E if l == 0 else Hl
                out[f'downscaler.{g}.weight_ih_l{l}'] = ((3 * Hl, inp), ('uniform', bl))
                out[f'downscaler.{g}.weight_hh_l{l}'] = ((3 * Hl, Hl), ('uniform', bl))
                out[f'downscaler.{g}.bias_ih_l{l}'] = ((3 * Hl,), ('uniform', bl))
                out[f'downscaler.{g}.bias_hh_l{l}'] = ((3 * Hl,), ('uniform', bl))
        fin = Hl * (2 if cfg.lstm_bidirectional else 1)
        out['downscaler.output_linear.weight'] = ((cfg.codebook_dim, fin), ('uniform', 1 / math.sqrt(fin)))
        out['downscaler.output_linear.bias'] = ((cfg.codebook_dim,), ('uniform', 1 / math.sqrt(fin)))
    else:
        out['downscaler.target_channel_embeddings'] = ((1, 1, V, pe), ('normal', 1.0))
        out['downscaler.events_positioning_embeddings'] = ((1, 1, S // V, pe), ('normal', 1.0))
        out['downscaler.input_linear.weight'] = ((d - 2 * pe, E), ('uniform', 1 / math.sqrt(E)))
        out['downscaler.input_linear.bias'] = ((d - 2 * pe,), ('uniform', 1 / math.sqrt(E)))
        out['downscaler.output_linear.weight'] = ((cfg.codebook_dim, d), ('uniform', 1 / math.sqrt(d)))
        out['downscaler.output_linear.bias'] = ((cfg.codebook_dim,), ('uniform', 1 / math.sqrt(d)))
        L = S
        for s, (factor, nl) in enumerate(zip(cfg.downscale_factors, cfg.list_of_num_layers)):
            for l in range(nl):
                p = f'downscaler.transformers.{s}.layers.{l}.'
                out[p + 'self_attn.in_proj_weight'] = ((3 * d, d), ('uniform', math.sqrt(6.0 / (4 * d))))
                out[p + 'self_attn.in_proj_bias'] = ((3 * d,), ('uniform', 0.02))
                out[p + 'self_attn.out_proj.weight'] = ((d, d), ('uniform', 1 / math.sqrt(d)))
                out[p + 'self_attn.out_proj.bias'] = ((d,), ('uniform', 0.02))
                out[p + 'self_attn.attn_bias.e1'] = ((H * L, hd), ('normal', 1.0))
                out[p + 'self_attn.attn_bias.e2'] = ((H * L, hd), ('normal', 1.0))
                out[p + 'linear1.weight'] = ((F, d), ('uniform', 1 / math.sqrt(d)))
                out[p + 'linear1.bias'] = ((F,), ('uniform', 1 / math.sqrt(d)))
                out[p + 'linear2.weight'] = ((d, F), ('uniform', 1 / math.sqrt(F)))
                out[p + 'linear2.bias'] = ((d,), ('uniform', 1 / math.sqrt(F)))
                out[p + 'norm1.weight'] = ((d,), ('normal1', 0.1))
                out[p + 'norm1.bias'] = ((d,), ('normal', 0.1))
                out[p + 'norm2.weight'] = ((d,), ('normal1', 0.1))
                out[p + 'norm2.bias'] = ((d,), ('normal', 0.1))
            if cfg.get('downscaler', 'transformer') == 'transformer_linear':
                out[f'downscaler.linear_aggs.{s}.weight'] = ((d, d * factor), ('uniform', 1 / math.sqrt(d * factor)))
                out[f'downscaler.linear_aggs.{s}.bias'] = ((d,), ('uniform', 1 / math.sqrt(d * factor)))
            L //= factor
    if cfg.codebook_size is not None:
        for i in range(cfg.num_codebooks):
            out[f'quantizer.embeddings.{i}'] = ((cfg.codebook_size, cfg.codebook_dim // cfg.num_codebooks),
                                                ('normal', cfg.get('codebook_scale', 1.0)))
    zd = cfg.codebook_dim
    if cfg.upscaler:
        out['upscaler.mlp.0.weight'] = ((cfg.up_hidden, cfg.codebook_dim), ('uniform', 1 / math.sqrt(cfg.codebook_dim)))
        out['upscaler.mlp.0.bias'] = ((cfg.up_hidden,), ('uniform', 1 / math.sqrt(cfg.codebook_dim)))
        out['upscaler.mlp.3.weight'] = ((cfg.z_dim, cfg.up_hidden), ('uniform', 1 / math.sqrt(cfg.up_hidden)))
        out['upscaler.mlp.3.bias'] = ((cfg.z_dim,), ('uniform', 1 / math.sqrt(cfg.up_hidden)))
        zd = cfg.z_dim
    Hh = cfg.c_hidden
    for suffix in (['', '_back'] if cfg.get('bidirectional', False) else ['']):
        for l in range(cfg.c_layers):
            inp = zd if l == 0 else Hh
            b = 1 / math.sqrt(Hh)
            out[f'c_module{suffix}.g_ar_fwd.weight_ih_l{l}'] = ((3 * Hh, inp), ('uniform', b))
            out[f'c_module{suffix}.g_ar_fwd.weight_hh_l{l}'] = ((3 * Hh, Hh), ('uniform', b))
            out[f'c_module{suffix}.g_ar_fwd.bias_ih_l{l}'] = ((3 * Hh,), ('uniform', b))
            out[f'c_module{suffix}.g_ar_fwd.bias_hh_l{l}'] = ((3 * Hh,), ('uniform', b))
        out[f'c_module{suffix}.output_linear.weight'] = ((cfg.c_dim, Hh), ('uniform', 1 / math.sqrt(Hh)))
        out[f'c_module{suffix}.output_linear.bias'] = ((cfg.c_dim,), ('uniform', 1 / math.sqrt(Hh)))
        out[f'fks_module{suffix}.W'] = ((zd, cfg.c_dim, cfg.KR), ('normal', 0.25))
    return out


def make_params(cfg, seed=0, dtype=torch.float32):
    prm = {}
    for name, (shape, (kind, scale)) in param_shapes(cfg).items():
        if kind == 'uniform':
            t = _uniform(name, seed, shape, scale)
        elif kind == 'normal':
            t = _normal(name, seed, shape, scale)
        else:  # normal1: around 1
            t = _normal(name, seed, shape, scale, 1.0)
        prm[name] = t.to(dtype)
    return prm


def split_by_file(prm):
    """flat dict -> {file: state_dict} as the reference saves them (encoder.py:47-57, trainer :117-131)."""
    out = {}
    for k, v in prm.items():
        f, key = k.split('.', 1)
        out.setdefault(f, {})[key] = v
    return out


def make_batch(cfg, B, seed=1234):
    """Synthetic batch dict with the producer contract of bach_cpc_dataloader.py:241-252
    (integer, CPU, voice dim last); uniform random tokens."""
    g = torch.Generator()
    g.manual_seed(seed)
    T, V = cfg.ticks_per_block, cfg.num_voices
    hi = min(cfg.vocab)

    def r(*shape):
        return torch.randint(0, hi, shape, generator=g, dtype=torch.int64)

    return dict(x_left=r(B, cfg.KL * T, V), x_right=r(B, cfg.KR * T, V),
                negative_samples=r(B, cfg.N, cfg.KR, T, V),
                negative_samples_back=r(B, cfg.N, cfg.KR, T, V))
