"""CPU oracle of the student / distillation trainer step.

TEST INFRASTRUCTURE ONLY (see oracle/vqcpc_oracle.py): plain-PyTorch CPU restatement over a flat parameter dict, pinned against
the live reference by ``oracle/make_golden_student.py`` -> ``tests/golden/student.pt`` and checked by
``tests/test_oracle_golden.py``.  Each function cites the reference file:line it follows.  The product never imports this.
"""
import math

import torch
import torch.nn.functional as F

from oracle import vqcpc_oracle as O


def student_cfg(**over):
    """Shapes of one student-trainer case: the encoder part is an ``O.default_cfg`` (transformer_linear downscaler, no upscaler)."""
    c = O.default_cfg(downscaler='transformer_linear', upscaler=False, codebook_size=16, KL=1, KR=1)
    c.update(dict(num_events=16, teacher_layers=2, teacher_pos_emb=8, dec_layers=[1, 1], quantization_weighting=0.1,
                  num_events_masked=1))
    c.update(over)
    return c


def student_param_shapes(cfg):
    """Adds the teacher (``teacher.*``, with its own ``teacher.data_processor.*``) and the auxiliary decoder (``decoder.*``) to
    the encoder's parameters; names as in the reference's state_dicts (student_encoder_trainer.py:93-95)."""
    from oracle import params as P
    enc = {k: v for k, v in P.param_shapes(cfg).items() if k.split('.')[0] in ('data_processor', 'downscaler', 'quantizer')}
    d, Fh, E, H = cfg.d_model, cfg.dim_feedforward, cfg.embedding_size, cfg.n_head
    hd = d // H
    V = cfg.num_voices
    out = dict(enc)

    def layer(prefix, L):
        out[prefix + 'self_attn.in_proj_weight'] = ((3 * d, d), ('uniform', math.sqrt(6.0 / (4 * d))))
        out[prefix + 'self_attn.in_proj_bias'] = ((3 * d,), ('uniform', 0.02))
        out[prefix + 'self_attn.out_proj.weight'] = ((d, d), ('uniform', 1 / math.sqrt(d)))
        out[prefix + 'self_attn.out_proj.bias'] = ((d,), ('uniform', 0.02))
        out[prefix + 'self_attn.attn_bias.e1'] = ((H * L, hd), ('normal', 0.5))
        out[prefix + 'self_attn.attn_bias.e2'] = ((H * L, hd), ('normal', 0.5))
        out[prefix + 'linear1.weight'] = ((Fh, d), ('uniform', 1 / math.sqrt(d)))
        out[prefix + 'linear1.bias'] = ((Fh,), ('uniform', 1 / math.sqrt(d)))
        out[prefix + 'linear2.weight'] = ((d, Fh), ('uniform', 1 / math.sqrt(Fh)))
        out[prefix + 'linear2.bias'] = ((d,), ('uniform', 1 / math.sqrt(Fh)))
        out[prefix + 'norm1.weight'] = ((d,), ('normal1', 0.1))
        out[prefix + 'norm1.bias'] = ((d,), ('normal', 0.1))
        out[prefix + 'norm2.weight'] = ((d,), ('normal1', 0.1))
        out[prefix + 'norm2.bias'] = ((d,), ('normal', 0.1))

    L = cfg.num_events * V
    pe = cfg.teacher_pos_emb
    for v in range(V):
        out[f'teacher.data_processor.embeddings.{v}.weight'] = ((cfg.vocab[v] + 1, E), ('normal', 1.0))
    out['teacher.channel_embeddings'] = ((1, V, pe), ('normal', 1.0))
    out['teacher.linear_to_input_transformer.weight'] = ((d - pe, E), ('uniform', 1 / math.sqrt(E)))
    out['teacher.linear_to_input_transformer.bias'] = ((d - pe,), ('uniform', 1 / math.sqrt(E)))
    for l in range(cfg.teacher_layers):
        layer(f'teacher.transformer.layers.{l}.', L)
    for v in range(V):
        out[f'teacher.pre_softmaxes.{v}.weight'] = ((cfg.vocab[v], d), ('uniform', 1 / math.sqrt(d)))
        out[f'teacher.pre_softmaxes.{v}.bias'] = ((cfg.vocab[v],), ('uniform', 1 / math.sqrt(d)))
    ups = list(reversed(cfg.downscale_factors))
    nb = L // int(math.prod(cfg.downscale_factors))
    out['decoder.linear.weight'] = ((d, cfg.codebook_dim), ('uniform', 1 / math.sqrt(cfg.codebook_dim)))
    out['decoder.linear.bias'] = ((d,), ('uniform', 1 / math.sqrt(cfg.codebook_dim)))
    Ls = nb
    for s, (u, nl) in enumerate(zip(ups, cfg.dec_layers)):
        out[f'decoder.upscale_embeddings.{s}'] = ((u, d), ('normal', 1.0))
        for l in range(nl):
            layer(f'decoder.transformers.{s}.layers.{l}.', Ls)
        Ls *= u
    for v in range(V):
        out[f'decoder.pre_softmaxes.{v}.weight'] = ((cfg.vocab[v], d), ('uniform', 1 / math.sqrt(d)))
        out[f'decoder.pre_softmaxes.{v}.bias'] = ((cfg.vocab[v],), ('uniform', 1 / math.sqrt(d)))
    return out


def make_student_params(cfg, seed=0, dtype=torch.float32):
    from oracle import params as P
    prm = {}
    for name, (shape, (kind, scale)) in student_param_shapes(cfg).items():
        if kind == 'uniform':
            t = P._uniform(name, seed, shape, scale)
        elif kind == 'normal':
            t = P._normal(name, seed, shape, scale)
        else:
            t = P._normal(name, seed, shape, scale, 1.0)
        prm[name] = t.to(dtype)
    return prm


def make_student_batch(cfg, B, seed=1234):
    g = torch.Generator()
    g.manual_seed(seed)
    return {'x': torch.stack([torch.randint(0, cfg.vocab[v], (B, cfg.num_events), generator=g) for v in range(cfg.num_voices)], dim=2)}


# ------------------------------------------------------------------------------------------------ forward passes
def mask_teacher(x, cfg, masked_event_index):
    """student_encoder_trainer.py:138-183 with the event index given."""
    B, Ev, C = x.shape
    keep = torch.ones(Ev, dtype=torch.bool)
    keep[max(masked_event_index - cfg.num_events_masked, 0):masked_event_index + cfg.num_events_masked + 1] = False
    mask_tok = torch.tensor(cfg.vocab, dtype=x.dtype).view(1, 1, C).expand(B, Ev, C)
    masked = torch.where(keep.view(1, Ev, 1), x, mask_tok)
    pred = torch.zeros_like(x)
    pred[:, masked_event_index] = 1
    return masked, pred


def teacher_forward(x_masked, prm, cfg, pre='teacher.'):
    """teachers/teacher_relative.py:57-87: logits per channel (B, num_events, vocab_c)."""
    B, Ev, C = x_masked.shape
    emb = O.embed(x_masked.reshape(B, Ev * C), [prm[f'{pre}data_processor.embeddings.{v}.weight'] for v in range(C)])
    h = F.linear(emb, prm[pre + 'linear_to_input_transformer.weight'], prm[pre + 'linear_to_input_transformer.bias'])   # (B, L, d - pe)
    chan = prm[pre + 'channel_embeddings'].reshape(1, C, -1).repeat(B, Ev, 1)
    h = torch.cat([h, chan], dim=2)
    for l in range(cfg.teacher_layers):
        h = O.encoder_layer(h, prm, f'{pre}transformer.layers.{l}.', cfg.n_head)
    h = h.view(B, Ev, C, -1)
    return [F.linear(h[:, :, v], prm[f'{pre}pre_softmaxes.{v}.weight'], prm[f'{pre}pre_softmaxes.{v}.bias']) for v in range(C)]


def decoder_forward(zq, prm, cfg, pre='decoder.'):
    """auxiliary_decoders/auxiliary_decoder_relative.py:84-130."""
    B, nb, _ = zq.shape
    C = cfg.num_voices
    h = F.linear(zq, prm[pre + 'linear.weight'], prm[pre + 'linear.bias'])
    ups = list(reversed(cfg.downscale_factors))
    for s, (u, nl) in enumerate(zip(ups, cfg.dec_layers)):
        for l in range(nl):
            h = O.encoder_layer(h, prm, f'{pre}transformers.{s}.layers.{l}.', cfg.n_head)
        emb = prm[f'{pre}upscale_embeddings.{s}']
        h = (h.unsqueeze(2) + emb.view(1, 1, u, -1)).reshape(B, h.shape[1] * u, -1)
    h = h.view(B, -1, C, h.shape[-1])
    return [F.linear(h[:, :, v], prm[f'{pre}pre_softmaxes.{v}.weight'], prm[f'{pre}pre_softmaxes.{v}.bias']) for v in range(C)]


def student_encoder_forward(x, prm, cfg):
    """encoder.py:76-95 with the 'bach' data processor (no block split in preprocess) and no upscaler."""
    B, Ev, C = x.shape
    emb = O.embed(x.reshape(B, Ev * C), [prm[f'data_processor.embeddings.{v}.weight'] for v in range(C)])
    z = O.downscaler_forward(emb, prm, cfg)
    books = [prm[f'quantizer.embeddings.{i}'] for i in range(cfg.num_codebooks)]
    zq, idx, qloss = O.quantize(z, books, cfg.commitment_cost, cfg.squared_l2_norm)
    return zq, idx, qloss


def student_losses(x, prm, cfg, masked_event_index):
    """-> (teacher loss, enc-dec loss, monitored dict, indices): student_encoder_trainer.py:120-136, :185-220."""
    masked, pred = mask_teacher(x, cfg, masked_event_index)
    t_logits = teacher_forward(masked, prm, cfg)
    e = masked_event_index
    loss_t = sum(F.cross_entropy(t_logits[v][:, e], x[:, e, v], reduction='none') for v in range(cfg.num_voices)).mean()
    zq, idx, qloss = student_encoder_forward(x, prm, cfg)
    d_logits = decoder_forward(zq, prm, cfg)
    rec = 0
    for v in range(cfg.num_voices):
        p = torch.softmax(t_logits[v][:, e].detach(), dim=1)
        rec = rec - (p * torch.log_softmax(d_logits[v][:, e], dim=1)).sum(dim=1)
    loss_e = cfg.quantization_weighting * qloss.mean() + rec.mean()
    mon = dict(loss_teacher=loss_t, loss_quantization=qloss.mean(), loss_reconstruction=rec.mean(), loss_encdec=loss_e)
    return loss_t, loss_e, mon, idx


def _group(name):
    return 'teacher' if name.startswith('teacher.') else ('decoder' if name.startswith('decoder.') else 'encoder')


def student_step(batch, prm, cfg, masked_event_index, lr=1e-4):
    """One training batch (student_encoder_trainer.py:237-276): teacher loss -> clip 5 -> Adam(teacher); then enc-dec loss (the
    teacher was already updated, but its logits were computed before the update and are detached) -> clip 5 on the decoder and
    on the encoder separately -> Adam.  Returns (monitored, grads, new parameters)."""
    x = batch['x'].long()
    leaves = {k: v.clone().requires_grad_(True) for k, v in prm.items()}
    loss_t, loss_e, mon, idx = student_losses(x, leaves, cfg, masked_event_index)
    names = list(leaves)
    gt = torch.autograd.grad(loss_t, [leaves[n] for n in names if _group(n) == 'teacher'], allow_unused=True)
    ge = torch.autograd.grad(loss_e, [leaves[n] for n in names if _group(n) != 'teacher'], allow_unused=True)
    grads = {}
    for n, g in zip([n for n in names if _group(n) == 'teacher'], gt):
        grads[n] = torch.zeros_like(prm[n]) if g is None else g
    for n, g in zip([n for n in names if _group(n) != 'teacher'], ge):
        grads[n] = torch.zeros_like(prm[n]) if g is None else g
    newp = {}
    for grp in ('teacher', 'decoder', 'encoder'):
        ns = [n for n in names if _group(n) == grp]
        clipped, _ = O.clip_grad_norm({n: grads[n] for n in ns}, 5.0)
        zeros = {n: torch.zeros_like(prm[n]) for n in ns}
        upd, _, _ = O.adam_step({n: prm[n] for n in ns}, clipped, zeros, zeros, 1, lr)
        newp.update(upd)
    return {k: v.detach() for k, v in mon.items()}, grads, newp, idx
