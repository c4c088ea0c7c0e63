"""Index maps of the reference's relative-attention bias (transformer/subsampled_relative_attention.py:30-122).

The reference forms ``rel = einsum(q, e)`` of shape (T, S) (T = target / query length, S = source / table length, T % S == 0),
REINTERPRETS it as (S, T) with ``view`` (:48, :82), pads one column of -100 (front for e1 :51-55, back for e2 :85-89) and T - S
rows of -100 (:58-64, :92-98), re-views to (-1, S), slices T rows (:67-70, :101-102) and finally zeroes the upper part of the e1
term and the flipped lower part of the e2 term (:107-120).  Entry (i, j) of the result is therefore either 0, -100 or ONE
element rel[t, s] - for T == S the familiar skewing (t == i), for T != S in general another query row.  Replaying those steps on
the flat indices 0 .. T*S-1 gives, once per (T, S), two int32 maps the attention kernel (``vqcpc_attn_map_fwd``) gathers with:
    m >= 0 : rel.flatten()[m] = q[m // S] . e[m % S]        m == -1 : 0        m == -2 : -100
"""
import functools

import torch

ZERO, PAD = -1, -2


@functools.lru_cache(maxsize=64)
def relative_index_maps(T, S):
    assert S <= T and T % S == 0
    ratio = T // S
    flat = torch.arange(T * S, dtype=torch.int64).view(S, T)                  # (T, S) product viewed as (S, T)   :48, :82
    pad_col = torch.full((S, 1), PAD, dtype=torch.int64)
    r1 = torch.cat([pad_col, flat], dim=1)                                     # :51-55
    r2 = torch.cat([flat, pad_col], dim=1)                                     # :85-89
    if T != S:
        ext = torch.full((T - S, T + 1), PAD, dtype=torch.int64)               # :58-64, :92-98
        r1 = torch.cat([r1, ext], dim=0)
        r2 = torch.cat([r2, ext], dim=0)
    r1 = r1.reshape(-1, S)[1:][:T]                                             # :67-70
    r2 = r2.reshape(-1, S)[:T]                                                 # :101-102
    ones = torch.ones(S, S, dtype=torch.bool)
    masks_down = torch.triu(ones, diagonal=0).flip(0).flip(1)                  # :107-111
    masks_up = torch.triu(ones, diagonal=1)                                    # :113-117
    if ratio != 1:
        masks_down = torch.repeat_interleave(masks_down, ratio, dim=0)
        masks_up = torch.repeat_interleave(masks_up, ratio, dim=0)
    r1 = r1.masked_fill(masks_up, ZERO)                                        # :119
    r2 = r2.masked_fill(masks_down, ZERO)                                      # :120
    # every surviving entry of row i is a product of query row i itself (the -100 padding only ever lands in the zeroed parts):
    # the incremental decoder (csrc/decode.cu) relies on it - row i of a layer never reads another row's query
    rows = torch.arange(T).view(T, 1)
    for r in (r1, r2):
        assert bool(((r < 0) | (r // S == rows)).all()), 'relative bias row reads a foreign query row'
    return r1.to(torch.int32).contiguous(), r2.to(torch.int32).contiguous()


_DEV_CACHE = {}


def device_maps(T, S, device):
    key = (T, S, str(device))
    if key not in _DEV_CACHE:
        m1, m2 = relative_index_maps(T, S)
        _DEV_CACHE[key] = (m1.to(device), m2.to(device))
    return _DEV_CACHE[key]
