"""CPU oracle for the VQ-CPC encoder-training hot path.

TEST INFRASTRUCTURE ONLY.  This file is a plain-PyTorch (CPU, fp32 or fp64) restatement of the
reference algorithm, written as explicit math over a flat parameter dict.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it.
The product path (``vqcpc-bach_b200/``) never imports anything from ``oracle/``.

Parity pinning: the reference has no tests or golden vectors of its own, so the oracle is
pinned against outputs of the *live reference modules* run on CPU in the authoring container:
``oracle/make_golden.py`` imports ``/root/reference``, loads the very same parameter dict (strict
``load_state_dict`` -> key names and shapes are pinned too), runs ``Encoder.forward`` and one
``VQCPCEncoderTrainer`` step and stores the results under ``tests/golden/``;
``tests/test_oracle_golden.py`` then checks this restatement against those files.

Every function cites the reference file:line (relative to /root/reference) it follows.

Layout note: the reference runs attention time-first ``(L, P, d)``; the oracle (and the CUDA path) keep
blocks major ``(P, L, d)`` — blocks are independent, so only the indexing differs.
"""
import math

import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------------------------
# configuration helper
# --------------------------------------------------------------------------------------------
class Cfg(dict):
    """Shape/hyper-parameter record shared by oracle, goldens and tests (attribute access)."""
    __getattr__ = dict.__getitem__
    __setattr__ = dict.__setitem__


def default_cfg(**over):
    c = Cfg(
        num_voices=4, ticks_per_block=4,          # 16 tokens per block, voice fastest
        vocab=[64, 64, 64, 64],                   # tokens per voice (tables carry +1 mask row)
        embedding_size=32, d_model=512, n_head=8, dim_feedforward=2048,
        downscale_factors=[4, 4], list_of_num_layers=[2, 2], pos_emb=8,
        codebook_size=16, codebook_dim=3, num_codebooks=1, commitment_cost=0.25, squared_l2_norm=True,
        upscaler=True, up_hidden=512, z_dim=32,
        c_hidden=512, c_layers=2, c_dim=32,
        KL=2, KR=2, N=15, quantization_weighting=0.5, bidirectional=False,
        downscaler='transformer',                 # or 'lstm' (LstmDownscaler: per-block GRUs, downscale_factors = [16])
        lstm_hidden=512, lstm_layers=2, lstm_bidirectional=True,
    )
    c.update(over)
    return c


# --------------------------------------------------------------------------------------------
# data processor
# --------------------------------------------------------------------------------------------
def preprocess(x, num_tokens_per_block):
    """VQCPCB/data_processor/bach_cpc_data_processor.py:17-40.

    (..., ticks, voices) -> (..., num_blocks, tokens_per_block) int64; row-major flatten so the voice index is
    fastest (token j <-> tick j // V, voice j % V), then cut into consecutive blocks."""
    ticks, voices = x.shape[-2:]
    lead = x.shape[:-2]
    flat = x.reshape(-1, ticks * voices)
    assert flat.shape[1] % num_tokens_per_block == 0
    nb = flat.shape[1] // num_tokens_per_block
    return flat.reshape(*lead, nb, num_tokens_per_block).long()


def embed(tokens, tables):
    """bach_cpc_data_processor.py:42-68 (+ tables data_processor.py:26-32).

    tokens (..., S) with voice = position % V  ->  (..., S, E) by per-voice table lookup."""
    V = len(tables)
    S = tokens.shape[-1]
    out = []
    for j in range(S):
        out.append(tables[j % V][tokens[..., j]])
    return torch.stack(out, dim=-2)


# --------------------------------------------------------------------------------------------
# transformer downscaler
# --------------------------------------------------------------------------------------------
def rel_bias(q_scaled, e1, e2):
    """Closed form of VQCPCB/transformer/subsampled_relative_attention.py:30-122 for self-attention
    (seq_len_src == seq_len_tgt == L), see SURVEY.md §8a row 6b.

    q_scaled (P, H, L, hd); e1, e2 (H*L, hd) -> bias (P, H, L, L):
        bias[i, j] = q_i . e1[h, L-1-(i-j)]   if j <= i
                   = q_i . e2[h, j-i]         if j >  i
    """
    P, H, L, hd = q_scaled.shape
    e1 = e1.view(H, L, hd)
    e2 = e2.view(H, L, hd)
    r1 = torch.einsum('phid,hrd->phir', q_scaled, e1)
    r2 = torch.einsum('phid,hrd->phir', q_scaled, e2)
    i = torch.arange(L).view(L, 1)
    j = torch.arange(L).view(1, L)
    lower = (j <= i)
    idx1 = (L - 1 - (i - j)).clamp(0, L - 1).expand(P, H, L, L)
    idx2 = (j - i).clamp(0, L - 1).expand(P, H, L, L)
    b1 = torch.gather(r1, 3, idx1)
    b2 = torch.gather(r2, 3, idx2)
    return torch.where(lower, b1, b2)


def mha_self(x, prm, pre, H, return_probs=False, attn_keep=None):
    """VQCPCB/transformer/multihead_attention_custom.py:154-353 (self-attention branch :168-171).

    x (P, L, d).  q is scaled by hd**-0.5 *after* the in-proj bias and *before* the relative bias (:247)."""
    P, L, d = x.shape
    hd = d // H
    qkv = F.linear(x, prm[pre + 'in_proj_weight'], prm[pre + 'in_proj_bias'])
    q, k, v = qkv.split(d, dim=-1)
    q = q * (float(hd) ** -0.5)
    q = q.view(P, L, H, hd).transpose(1, 2)          # (P, H, L, hd)
    k = k.view(P, L, H, hd).transpose(1, 2)
    v = v.view(P, L, H, hd).transpose(1, 2)
    s = torch.matmul(q, k.transpose(-1, -2))         # :314
    s = s + rel_bias(q, prm[pre + 'attn_bias.e1'], prm[pre + 'attn_bias.e2'])   # :331-333
    a = torch.softmax(s, dim=-1)                     # :338
    if isinstance(attn_keep, float):                 # :339 F.dropout on the weights, fresh draw (CPU-baseline timing only)
        a = F.dropout(a, attn_keep, training=True)
    elif attn_keep is not None:                      # ... or replayed from an explicit mask
        a = a * attn_keep                            # (keep / (1 - p), shape (P, H, L, L))
    o = torch.matmul(a, v).transpose(1, 2).reshape(P, L, d)    # :343-345
    out = F.linear(o, prm[pre + 'out_proj.weight'], prm[pre + 'out_proj.bias'])  # :346
    if return_probs:
        return out, a
    return out


def encoder_layer(x, prm, pre, H, drop=None):
    """VQCPCB/transformer/transformer_custom.py:268-291 (post-LN, ReLU FFN, eps=1e-5).

    drop: None (p = 0) or a dict of multiplicative dropout factors keep / (1 - p) replaying a given draw:
    'attn' (P, H, L, L) on the attention weights, 'drop1' / 'drop2' (P, L, d) on the two residual branches (:278, :287),
    'ffn' (P, L, F) after the activation (:283)."""
    d = x.shape[-1]
    if isinstance(drop, float):
        # fresh draws at the reference's four sites (the training setting of the shipped configs, p = 0.1).  RNG streams
        # cannot be compared with anything, so this branch exists for the CPU-baseline TIMING line only.
        p = drop
        a = F.dropout(mha_self(x, prm, pre + 'self_attn.', H, attn_keep=p), p, training=True)
        x = F.layer_norm(x + a, (d,), prm[pre + 'norm1.weight'], prm[pre + 'norm1.bias'], 1e-5)
        hid = F.dropout(torch.relu(F.linear(x, prm[pre + 'linear1.weight'], prm[pre + 'linear1.bias'])), p, training=True)
        f = F.dropout(F.linear(hid, prm[pre + 'linear2.weight'], prm[pre + 'linear2.bias']), p, training=True)
        return F.layer_norm(x + f, (d,), prm[pre + 'norm2.weight'], prm[pre + 'norm2.bias'], 1e-5)
    a = mha_self(x, prm, pre + 'self_attn.', H, attn_keep=None if drop is None else drop['attn'])
    if drop is not None:
        a = a * drop['drop1']
    x = F.layer_norm(x + a, (d,), prm[pre + 'norm1.weight'], prm[pre + 'norm1.bias'], 1e-5)
    hid = torch.relu(F.linear(x, prm[pre + 'linear1.weight'], prm[pre + 'linear1.bias']))
    if drop is not None:
        hid = hid * drop['ffn']
    f = F.linear(hid, prm[pre + 'linear2.weight'], prm[pre + 'linear2.bias'])
    if drop is not None:
        f = f * drop['drop2']
    x = F.layer_norm(x + f, (d,), prm[pre + 'norm2.weight'], prm[pre + 'norm2.bias'], 1e-5)
    return x


def downscaler_forward(emb, prm, cfg, pre='downscaler.', trace=None, drops=None):
    """VQCPCB/downscalers/relative_transformer_downscaler.py:93-133.

    emb (b, n_tok, E) -> z (b, n_tok / prod(factors), D)."""
    S = int(math.prod(cfg.downscale_factors))
    V = cfg.num_voices
    b, n_tok, E = emb.shape
    nb = n_tok // S
    x = emb.reshape(b * nb, S, E)
    h = F.linear(x, prm[pre + 'input_linear.weight'], prm[pre + 'input_linear.bias'])        # :104
    chan = prm[pre + 'target_channel_embeddings'].reshape(V, -1)                            # (V, 8)
    evt = prm[pre + 'events_positioning_embeddings'].reshape(S // V, -1)                    # (S/V, 8)
    t = torch.arange(S)
    h = torch.cat([h, chan[t % V].expand(b * nb, S, -1), evt[t // V].expand(b * nb, S, -1)], dim=-1)   # :106-112
    if trace is not None:
        trace['h0'] = h
    p_draw = float(cfg.get('dropout', 0.0))
    if drops is None and p_draw > 0:
        drops = [p_draw] * sum(cfg.list_of_num_layers)
    for s, (factor, nl) in enumerate(zip(cfg.downscale_factors, cfg.list_of_num_layers)):   # :122-124
        for l in range(nl):
            h = encoder_layer(h, prm, f'{pre}transformers.{s}.layers.{l}.', cfg.n_head,
                              drop=None if drops is None else drops[len([1 for ss in range(s) for _ in range(cfg.list_of_num_layers[ss])]) + l])
            if trace is not None:
                trace[f'x_{s}_{l}'] = h
        if cfg.get('downscaler', 'transformer') == 'transformer_linear':
            # relative_transformer_downscaler_linear.py:129-133: concat `factor` consecutive tokens, Linear(factor*d -> d)
            n, Lc, dm = h.shape
            h = F.linear(h.reshape(n, Lc // factor, factor * dm), prm[f'{pre}linear_aggs.{s}.weight'],
                         prm[f'{pre}linear_aggs.{s}.bias'])
        else:
            h = h[:, ::factor]
    assert h.shape[1] == 1                                                                    # :128
    z = F.linear(h[:, 0], prm[pre + 'output_linear.weight'], prm[pre + 'output_linear.bias'])  # :132
    return z.view(b, nb, -1)


# --------------------------------------------------------------------------------------------
# quantizer / upscaler
# --------------------------------------------------------------------------------------------
def quantize(z, codebooks, commitment_cost, squared_l2_norm=True, given_idx=None):
    """VQCPCB/quantizer/vector_quantizer.py:85-159 (no batch-norm, no init).  Label corruption (:118-132) replaces a
    random 5 % of the searched indices; RNG streams cannot be matched, so the caller passes the already corrupted
    indices as ``given_idx`` (n, ncb) and the search is skipped for them.

    z (..., D); codebooks: list of (Kc, D/ncb).  Returns (z_sg, idx int64 (..., ncb), loss (...))."""
    shp = z.shape
    D = shp[-1]
    flat = z.reshape(-1, D)
    ncb = len(codebooks)
    idxs, qs = [], []
    for ci, (comp, E) in enumerate(zip(flat.chunk(ncb, dim=1), codebooks)):
        if given_idx is not None:
            idxs.append(given_idx.reshape(-1, ncb)[:, ci])
            qs.append(E[idxs[-1]])
            continue
        dist = torch.sum((comp.unsqueeze(1) - E.unsqueeze(0)) ** 2, dim=2)      # :105-112
        idx = torch.argmin(dist, dim=1)                                         # :115-116 first-min tie rule
        idxs.append(idx)
        qs.append(E[idx])                                                       # == onehot @ E (:134-143)
    q = torch.cat(qs, dim=1).view(shp)
    if squared_l2_norm:                                                         # :72-83
        e_l = torch.sum((q.detach() - z) ** 2, dim=-1)
        q_l = torch.sum((q - z.detach()) ** 2, dim=-1)
    else:
        e_l = torch.norm((q.detach() - z) + 1e-5, dim=-1)
        q_l = torch.norm((q - z.detach()) + 1e-5, dim=-1)
    loss = q_l + commitment_cost * e_l
    z_sg = z + (q - z).detach()                                                 # :148
    idx = torch.stack(idxs, dim=-1).view(*shp[:-1], ncb)                        # :156-157
    return z_sg, idx, loss


def upscale(zq, prm, pre='upscaler.', p=0.0):
    """VQCPCB/upscalers/mlp_upscaler.py:21-35: Linear -> Dropout(p) -> SELU -> Linear (p > 0: fresh draw, timing only)."""
    h = F.linear(zq, prm[pre + 'mlp.0.weight'], prm[pre + 'mlp.0.bias'])
    if p > 0:
        h = F.dropout(h, p, training=True)
    h = F.selu(h)
    return F.linear(h, prm[pre + 'mlp.3.weight'], prm[pre + 'mlp.3.bias'])


def encoder_forward(x, prm, cfg, trace=None):
    """VQCPCB/encoder.py:76-95: preprocess -> embed -> flatten -> downscaler -> quantizer -> upscaler."""
    S = cfg.num_voices * cfg.ticks_per_block
    tok = preprocess(x, S)                                    # (b, nb, 16)
    tables = [prm[f'data_processor.embeddings.{v}.weight'] for v in range(cfg.num_voices)]
    emb = embed(tok, tables)                                  # (b, nb, 16, E)
    b, nb = emb.shape[:2]
    emb = emb.reshape(b, nb * S, -1)                          # utils.py:52-63 flatten
    if cfg.get('downscaler', 'transformer') == 'lstm':
        z = lstm_downscaler_forward(emb, prm, cfg)
    else:
        z = downscaler_forward(emb, prm, cfg, trace=trace)
    if trace is not None:
        trace['z'] = z
    if cfg.codebook_size is None:                             # NoQuantization vector_quantizer.py:15-24
        zq, idx, ql = z, None, torch.zeros_like(z).sum(-1)
    else:
        books = [prm[f'quantizer.embeddings.{i}'] for i in range(cfg.num_codebooks)]
        zq, idx, ql = quantize(z, books, cfg.commitment_cost, cfg.squared_l2_norm)
    if cfg.upscaler:
        zq = upscale(zq, prm, p=float(cfg.get('dropout', 0.0)))
    return zq, idx, ql


# --------------------------------------------------------------------------------------------
# CPC heads
# --------------------------------------------------------------------------------------------
def gru_stack(x, prm, pre, layers, hidden, p=0.0):
    """torch.nn.GRU(batch_first=True, h0 = 0) as explicit math: gate order (r, z, n),
    n = tanh(W_in x + b_in + r * (W_hn h + b_hn)), h' = (1 - z) n + z h.  x (B, T, in) -> (B, T, hidden)."""
    B, T, _ = x.shape
    for l in range(layers):
        wi, wh = prm[f'{pre}weight_ih_l{l}'], prm[f'{pre}weight_hh_l{l}']
        bi, bh = prm[f'{pre}bias_ih_l{l}'], prm[f'{pre}bias_hh_l{l}']
        h = x.new_zeros(B, hidden)
        outs = []
        for t in range(T):
            gi = F.linear(x[:, t], wi, bi)
            gh = F.linear(h, wh, bh)
            i_r, i_z, i_n = gi.chunk(3, 1)
            h_r, h_z, h_n = gh.chunk(3, 1)
            r = torch.sigmoid(i_r + h_r)
            zg = torch.sigmoid(i_z + h_z)
            n = torch.tanh(i_n + r * h_n)
            h = (1 - zg) * n + zg * h
            outs.append(h)
        x = torch.stack(outs, 1)
        if p > 0 and l < layers - 1:                 # nn.GRU(dropout=p): between layers only (fresh draw, timing only)
            x = F.dropout(x, p, training=True)
    return x


def gru_c(y_left, prm, cfg, pre='c_module.'):
    """VQCPCB/vqcpc_helper.py:54-76: multi-layer GRU (batch_first, h0=0), last time step, Linear."""
    x = gru_stack(y_left, prm, pre + 'g_ar_fwd.', cfg.c_layers, cfg.c_hidden, p=float(cfg.get('dropout', 0.0)))
    return F.linear(x[:, -1], prm[pre + 'output_linear.weight'], prm[pre + 'output_linear.bias'])


def lstm_downscaler_forward(emb, prm, cfg, pre='downscaler.'):
    """VQCPCB/downscalers/lstm_downscaler.py:57-94: the sequence is cut into blocks of downscale_factors[0] tokens; each
    block goes through a multi-layer GRU (and, if bidirectional, a SECOND GRU on the time-flipped block); the last
    states are concatenated and projected by output_linear.  emb (b, nb * S, E) -> (b, nb, D)."""
    b, ntok, E = emb.shape
    S = cfg.num_voices * cfg.ticks_per_block
    nb = ntok // S
    x = emb.reshape(b * nb, S, E)
    last = gru_stack(x, prm, pre + 'g_enc_fwd.', cfg.lstm_layers, cfg.lstm_hidden)[:, -1]
    if cfg.lstm_bidirectional:
        last_b = gru_stack(x.flip(dims=[1]), prm, pre + 'g_enc_bwd.', cfg.lstm_layers, cfg.lstm_hidden)[:, -1]
        last = torch.cat([last, last_b], dim=-1)
    z = F.linear(last, prm[pre + 'output_linear.weight'], prm[pre + 'output_linear.bias'])
    return z.view(b, nb, -1)


def fks(c, zs, W):
    """VQCPCB/vqcpc_helper.py:79-98: u[b,k,:] = sum_c c[b,c] W[:,c,k];  f[b,k] = u[b,k,:] . zs[b,k,:]."""
    u = torch.einsum('bc,zck->bkz', c, W)
    return (u * zs).sum(-1)


def nce_loss(pos, neg):
    """VQCPCB/vqcpc_helper.py:5-29: -(1/B) sum_b sum_k (pos - logsumexp([neg, pos]))."""
    lse = torch.logsumexp(torch.cat([neg, pos.unsqueeze(2)], dim=2), dim=2)
    return -(pos - lse).sum(1).mean(0)


def trainer_forward(batch, prm, cfg):
    """VQCPCB/vqcpc_encoder_trainer.py:192-307 (unidirectional c-net; Q1 of SURVEY.md §5 is bidirectional-only).

    batch: x_left (B, KL*T, V), x_right (B, KR*T, V), negative_samples (B, N, KR, T, V) integer tensors."""
    neg = batch['negative_samples']
    B, N, KR, T, V = neg.shape
    zn, idx_n, ql_n = encoder_forward(neg.reshape(B * N * KR, T, V), prm, cfg)        # :197-201
    zn = zn.view(B, N, KR, -1)                                                        # nb == 1 (:245)
    zl, idx_l, ql_l = encoder_forward(batch['x_left'], prm, cfg)                      # :227
    zr, idx_r, ql_r = encoder_forward(batch['x_right'], prm, cfg)                     # :230
    c = gru_c(zl, prm, cfg)                                                           # :237
    W = prm['fks_module.W']
    pos = fks(c, zr, W)                                                               # :240  (B, KR)
    # negatives go through the very same arithmetic as the positive (the reference calls one FksModule for both,
    # :240/:258), so a negative whose code equals the positive's scores *exactly* equal -> 'pos > max' is False.
    negs = torch.stack([fks(c, zn[:, n], W) for n in range(N)], dim=2)                # :245-263 (B, KR, N)
    correct = pos > negs.max(2)[0]                                                    # :269
    nce = nce_loss(pos, negs)                                                         # :272
    acc = correct.sum(0).float() / B                                                  # :337
    ql_parts = [ql_l.sum(1), ql_r.sum(1), ql_n.view(B, -1).sum(1)]
    if cfg.get('bidirectional', False):                                               # :277-296
        znb, idx_nb, ql_nb = encoder_forward(batch['negative_samples_back'].reshape(B * N * KR, T, V), prm, cfg)  # :211-225
        c_back = gru_c(zr.flip(dims=[1]), prm, cfg, pre='c_module_back.')             # :278-279
        Wb = prm['fks_module_back.W']
        pos_b = fks(c_back, zl, Wb)                                                   # :283 (left codes NOT flipped)
        # quirk Q1: the (B, N, KR, 1, z) tensor goes into FksModule as is -> viewed (N*B, KR, z) b-major while
        # c_back.repeat(N, 1) is n-major (:285-286, vqcpc_helper.py:96); then .view(N, B, KR).permute(1, 2, 0) (:288-291)
        f = fks(c_back.repeat(N, 1), znb.reshape(B * N, KR, -1), Wb)
        negs_b = f.view(N, B, KR).permute(1, 2, 0)
        nce = nce + nce_loss(pos_b, negs_b)                                           # :296
        acc = (acc + (pos_b > negs_b.max(2)[0]).sum(0).float() / B) / 2                # :338-339
        ql_parts.append(ql_nb.view(B, -1).sum(1))
    # quantization_loss vqcpc_helper.py:32-51: mean over cat[left.sum(1), right.sum(1), neg.sum(all) (, neg_back)]
    ql = torch.cat(ql_parts, 0).mean()
    loss = nce + cfg.quantization_weighting * ql                                      # :307
    return dict(loss=loss, nce=nce, qloss=ql, accuracy=acc, pos=pos, neg=negs, c=c,
                idx_left=idx_l, idx_right=idx_r, idx_negative=idx_n,
                z_left=zl, z_right=zr, z_negative=zn)


# --------------------------------------------------------------------------------------------
# optimizer tail
# --------------------------------------------------------------------------------------------
def clip_grad_norm(grads, max_norm=5.0):
    """torch.nn.utils.clip_grad_norm_ semantics used at vqcpc_encoder_trainer.py:313:
    coef = min(1, max_norm / (||g||_2 + 1e-6)) over ALL parameters jointly."""
    total = torch.sqrt(sum((g.double() ** 2).sum() for g in grads.values())).float()
    coef = torch.clamp(max_norm / (total + 1e-6), max=1.0)
    return {k: g * coef for k, g in grads.items()}, total


def adam_step(params, grads, m, v, step, lr, b1=0.9, b2=0.999, eps=1e-8):
    """torch.optim.Adam defaults as constructed at vqcpc_encoder_trainer.py:92 (no weight decay, no amsgrad).
    ``step`` is the 1-based step count.  Returns new (params, m, v)."""
    out_p, out_m, out_v = {}, {}, {}
    bc1 = 1 - b1 ** step
    bc2 = 1 - b2 ** step
    for k, p in params.items():
        g = grads[k]
        mk = b1 * m[k] + (1 - b1) * g
        vk = b2 * v[k] + (1 - b2) * g * g
        denom = vk.sqrt() / math.sqrt(bc2) + eps
        out_p[k] = p - (lr / bc1) * mk / denom
        out_m[k], out_v[k] = mk, vk
    return out_p, out_m, out_v


def lr_lambda(step):
    """vqcpc_encoder_trainer.py:95-107 LambdaLR factor."""
    warm, lo, hi = 10000, 0.1, 1.0
    s1 = (hi - lo) / warm
    s2 = -s1 * 0.1
    return max(min(lo + s1 * step, hi + (step - warm) * s2), lo)


def trainer_step(batch, prm, cfg, m=None, v=None, step=1, lr=1e-4):
    """One full training step: forward, backward (autograd over the restated math), clip 5.0, Adam."""
    leaves = {k: p.detach().clone().requires_grad_(True) for k, p in prm.items()}
    out = trainer_forward(batch, leaves, cfg)
    out['loss'].backward()
    grads = {k: (p.grad if p.grad is not None else torch.zeros_like(p)) for k, p in leaves.items()}
    clipped, gnorm = clip_grad_norm(grads)
    if m is None:
        m = {k: torch.zeros_like(p) for k, p in prm.items()}
        v = {k: torch.zeros_like(p) for k, p in prm.items()}
    newp, m, v = adam_step({k: p.detach() for k, p in prm.items()}, clipped, m, v, step, lr)
    out = {k: (t.detach() if torch.is_tensor(t) else t) for k, t in out.items()}
    return out, grads, gnorm, newp, m, v


# --------------------------------------------------------------------------------------------
# EMA codebook (extension; NO reference counterpart: "parity unpinned")
# --------------------------------------------------------------------------------------------
def ema_codebook_update(N, m, z, idx, decay, eps=1e-5):
    """CPU restatement of the optional EMA codebook update of the product (quantizer.ema_update), the standard VQ-VAE
    rule (van den Oord et al. 2017, appendix A; Sonnet ``VectorQuantizerEMA``).  The reference repository has no EMA
    update at all (its codebook is a gradient-trained ParameterList, vector_quantizer.py:44-48, :72-83), so this function
    pins nothing against the reference - it only pins the CUDA path against an independent statement of the same rule.

    N (K,) cluster sizes, m (K, D) cluster sums, z (n, D) encoder outputs of the step, idx (n,) their codes.
    Returns (N', m', E')."""
    K, D = m.shape
    n = torch.bincount(idx.reshape(-1), minlength=K).to(m.dtype)
    s = torch.zeros_like(m).index_add_(0, idx.reshape(-1), z.reshape(-1, D).to(m.dtype))
    N2 = decay * N + (1 - decay) * n
    m2 = decay * m + (1 - decay) * s
    tot = N2.sum()
    Ns = (N2 + eps) / (tot + K * eps) * tot
    return N2, m2, m2 / Ns.unsqueeze(1)
