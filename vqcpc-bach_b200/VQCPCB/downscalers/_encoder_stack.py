"""The transformer downscaler as ONE autograd node driving the fused CUDA kernels.

Forward (per 16-token block, VQCPCB/downscalers/relative_transformer_downscaler.py:93-133 with the layer math of
transformer/transformer_custom.py:268-291 and multihead_attention_custom.py:154-353):

    h0 = [W_in e + b_in | chan_emb[voice] | evt_emb[tick]]                         GEMM (ldc = d) + posemb_fill
    per layer:  qkv = x Wqkv^T + b  (q third scaled by hd^-0.5 in the epilogue)     GEMM
                o   = softmax(q~ (k + E)^T) v   per (block, head)                   attn kernel
                y1  = o Wo^T + bo + x ; x1 = LN1(y1)                                GEMM(+residual) + LN kernel
                hm  = relu(x1 W1^T + b1) ; y2 = hm W2^T + b2 + x1 ; x2 = LN2(y2)    GEMM(relu) + GEMM(+residual) + LN
    after each stack keep every ``factor``-th token (or, for the ``...DownscalerLinear`` variant, concatenate ``factor``
    consecutive tokens and apply ``linear_aggs[s]``: Linear(factor * d -> d),
    relative_transformer_downscaler_linear.py:84-86,129-133); z = x W_out^T + b_out.

Rows are blocks-major (row = block * L + token) - blocks never interact, so this equals the reference's time-first
layout.  Backward replays the saved activations in reverse; parameter gradients are written by the kernels straight
into ``param.grad`` (accumulating), which the trainer aliases to its flat gradient buffer.

Dropout (training with p > 0; the reference's four sites per layer: on the attention weights, ``dropout1`` / ``dropout2``
before the residual adds, ``dropout`` after the activation) uses explicit uint8 keep masks drawn per site: the attention
kernels take the mask directly, the three GEMM-side sites run ``vqcpc_mask_scale_add`` on the GEMM output (which also
adds the residual / writes the bf16 split the next GEMM reads), and backward applies the same masks to the incoming
gradients.  With p = 0 (the parity-checked headline setting) nothing changes: residual / ReLU stay in the GEMM epilogues.
"""
import torch

from .. import ops

_LAYER_KEYS = ('wqkv', 'bqkv', 'wo', 'bo', 'e1', 'e2', 'w1', 'b1', 'w2', 'b2', 'g1', 'be1', 'g2', 'be2')


def _subsample(x, P, L, factor, d):
    """rows (block, token) -> rows (block, token // factor) keeping token % factor == 0."""
    out = torch.empty(P * (L // factor), d, dtype=torch.float32, device=x.device)
    ops._native.call('vqcpc_subsample_fwd', ops._p(x), ops._p(out), P, L, factor, d, ops._stream())
    return out


def _subsample_bwd(dout, P, L, factor, d):
    """adjoint of ``_subsample``: scatter into all L rows, zeros elsewhere."""
    dx = torch.empty(P * L, d, dtype=torch.float32, device=dout.device)
    ops._native.call('vqcpc_subsample_bwd', ops._p(dout), ops._p(dx), P, L, factor, d, ops._stream())
    return dx


class EncoderStackFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, emb2d, meta, *params):
        S, V, pe, d, H, factors, layers_per_stack, out_dim = meta[:8]
        linear_agg = len(meta) > 8 and bool(meta[8])
        drop_p = float(meta[9]) if len(meta) > 9 else 0.0
        dp = drop_p > 0.0
        ds = 1.0 / (1.0 - drop_p) if dp else 1.0
        hd = d // H
        emb2d = ops._f32c(emb2d)
        M0, E = emb2d.shape
        assert M0 % S == 0
        P = M0 // S
        dev = emb2d.device
        prm = [p.detach() for p in params]
        w_in, b_in, chan, evt, w_out, b_out = prm[:6]
        n_layers_total = sum(layers_per_stack)
        layer_params = prm[6:6 + 14 * n_layers_total]
        agg_params = prm[6 + 14 * n_layers_total:]          # (weight, bias) per stack, DownscalerLinear only
        # activations are saved only when a backward can follow.  Grad mode is always off inside Function.forward and
        # ``needs_input_grad`` stays True for parameters under torch.no_grad(), so the CALLER samples
        # torch.is_grad_enabled() and passes it in meta[10] (eval epochs then keep no probabilities / activations).
        train = any(ctx.needs_input_grad) and (bool(meta[10]) if len(meta) > 10 else True)

        h = torch.empty(M0, d, dtype=torch.float32, device=dev)
        ops.gemm(emb2d, w_in, M0, d - 2 * pe, E, True, True, out=h[:, :d - 2 * pe], bias=b_in, b_is_weight=True)
        ops._native.call('vqcpc_posemb_fill', ops._p(h), d, d - 2 * pe, ops._p(chan), ops._p(evt), M0, S, V, pe, ops._stream())

        saved = []
        tc = ops.tensor_core_mode()
        x, x_sp, L, li = h, None, S, 0
        q_scale = float(hd) ** -0.5
        agg_in = []
        for si, (factor, nl) in enumerate(zip(factors, layers_per_stack)):
            for layer_in_stack in range(nl):
                lp = dict(zip(_LAYER_KEYS, layer_params[li * 14:(li + 1) * 14]))
                li += 1
                # Exact dead-row elimination: after the LAST layer of a stack only every ``factor``-th token survives
                # (relative_transformer_downscaler.py:124), and a token's post-attention path (out-proj, LN, FFN, LN)
                # is row-wise, so that layer runs it on the surviving rows only.  K/V (hence the in-proj and the
                # attention itself) still see all L tokens.  Saves ~31 % of the encoder FLOPs with identical results.
                last = (layer_in_stack == nl - 1) and factor > 1 and not linear_agg
                keep = factor if last else 1
                if tc:
                    # tensor-core path: every GEMM operand travels as a bf16 (hi, lo) pair written by its PRODUCER
                    # (LayerNorm / attention / GEMM epilogues) - no separate split passes, fp32 copies only where an
                    # elementwise consumer needs them (residuals, LayerNorm inputs, the attention kernel's q/k/v).
                    if x_sp is None:
                        x_sp = ops.split_bf16(x)
                    qkv = ops.gemm(None, lp['wqkv'], P * L, 3 * d, d, bias=lp['bqkv'], scale_cols=(q_scale, d),
                                   b_is_weight=True, a_split=x_sp)
                    Mk = P * (L // keep)
                    Fh = lp['w1'].shape[0]
                    masks = None
                    if dp:
                        masks = tuple(ops.dropout_mask(f'{site}{li - 1}', shp, drop_p, dev) for site, shp in
                                      (('attn', (P, H, L, L)), ('drop1_', (Mk, d)), ('ffn', (Mk, Fh)), ('drop2_', (Mk, d))))
                    _, o_sp, probs = ops.attn_fwd(qkv, lp['e1'], lp['e2'], P, L, H, hd, need_probs=train, keep=keep,
                                                  want_fp32=False, want_split=True, drop=(masks[0], ds) if dp else None)
                    x_k = _subsample(x, P, L, factor, d) if last else x
                    if dp:
                        y1p = ops.gemm(None, lp['wo'], Mk, d, d, bias=lp['bo'], b_is_weight=True, a_split=o_sp)
                        y1, _ = ops.mask_scale_add(y1p, masks[1], ds, residual=x_k)
                    else:
                        y1 = ops.gemm(None, lp['wo'], Mk, d, d, bias=lp['bo'], residual=x_k, b_is_weight=True, a_split=o_sp)
                    x1, x1_sp, mean1, rstd1 = ops.layernorm_fwd(y1, lp['g1'], lp['be1'], want_split=True)
                    if dp:
                        hmp = ops.gemm(None, lp['w1'], Mk, Fh, d, bias=lp['b1'], relu=True, b_is_weight=True, a_split=x1_sp)
                        _, hm_sp = ops.mask_scale_add(hmp, masks[2], ds, want_fp32=False, want_split=True)
                        del hmp
                        y2p = ops.gemm(None, lp['w2'], Mk, d, Fh, bias=lp['b2'], b_is_weight=True, a_split=hm_sp)
                        y2, _ = ops.mask_scale_add(y2p, masks[3], ds, residual=x1)
                    else:
                        _, hm_sp = ops.gemm(None, lp['w1'], Mk, Fh, d, bias=lp['b1'], relu=True, b_is_weight=True, a_split=x1_sp,
                                            want_c=False, out_split=True)
                        y2 = ops.gemm(None, lp['w2'], Mk, d, Fh, bias=lp['b2'], residual=x1, b_is_weight=True, a_split=hm_sp)
                    x2, x2_sp, mean2, rstd2 = ops.layernorm_fwd(y2, lp['g2'], lp['be2'], want_split=True)
                    if train:
                        saved.append((x_sp, qkv, probs, o_sp, y1, mean1, rstd1, x1_sp, hm_sp, y2, mean2, rstd2, L, last, masks))
                    x, x_sp = x2, x2_sp
                    continue
                qkv = ops.linear_fwd(x, lp['wqkv'], lp['bqkv'], scale_cols=(q_scale, d))
                Mk = P * (L // keep)
                masks = None
                if dp:
                    masks = tuple(ops.dropout_mask(f'{site}{li - 1}', shp, drop_p, dev) for site, shp in
                                  (('attn', (P, H, L, L)), ('drop1_', (Mk, d)), ('ffn', (Mk, lp['w1'].shape[0])), ('drop2_', (Mk, d))))
                o_k, probs = ops.attn_fwd(qkv, lp['e1'], lp['e2'], P, L, H, hd, need_probs=train, keep=keep,
                                          drop=(masks[0], ds) if dp else None)
                x_k = _subsample(x, P, L, factor, d) if last else x
                if dp:
                    y1, _ = ops.mask_scale_add(ops.linear_fwd(o_k, lp['wo'], lp['bo']), masks[1], ds, residual=x_k)
                else:
                    y1 = ops.linear_fwd(o_k, lp['wo'], lp['bo'], residual=x_k)
                x1, mean1, rstd1 = ops.layernorm_fwd(y1, lp['g1'], lp['be1'])
                hm = ops.linear_fwd(x1, lp['w1'], lp['b1'], relu=True)
                if dp:
                    hm, _ = ops.mask_scale_add(hm, masks[2], ds)
                    y2, _ = ops.mask_scale_add(ops.linear_fwd(hm, lp['w2'], lp['b2']), masks[3], ds, residual=x1)
                else:
                    y2 = ops.linear_fwd(hm, lp['w2'], lp['b2'], residual=x1)
                x2, mean2, rstd2 = ops.layernorm_fwd(y2, lp['g2'], lp['be2'])
                if train:
                    saved.append((x, qkv, probs, o_k, y1, mean1, rstd1, x1, hm, y2, mean2, rstd2, L, last, masks))
                x = x2
            if linear_agg:
                # rows are blocks-major, so ``factor`` consecutive tokens of a block are ``factor`` consecutive rows: the
                # reference's reshape/permute/reshape is a free view here
                xv = x.view(P * (L // factor), factor * d)
                if train:
                    agg_in.append(xv)
                x, x_sp = ops.linear_fwd(xv, agg_params[2 * si], agg_params[2 * si + 1]), None
            elif nl == 0 or factor == 1:
                x = _subsample(x, P, L, factor, d) if factor > 1 else x
            L = L // factor
        assert L == 1, 'the product of downscale_factors must equal the block length'
        z = ops.linear_fwd(x, w_out, b_out)

        ctx.meta = meta
        ctx.params = params
        ctx.tc = tc
        ctx.saved = saved
        ctx.agg_in = agg_in
        ctx.emb2d = emb2d if train else None
        ctx.x_last = x if train else None
        return z

    @staticmethod
    def backward(ctx, dz):
        S, V, pe, d, H, factors, layers_per_stack, out_dim = ctx.meta[:8]
        linear_agg = len(ctx.meta) > 8 and bool(ctx.meta[8])
        drop_p = float(ctx.meta[9]) if len(ctx.meta) > 9 else 0.0
        ds = 1.0 / (1.0 - drop_p) if drop_p > 0 else 1.0
        hd = d // H
        params = ctx.params
        prm = [p.detach() for p in params]
        w_in, b_in, chan, evt, w_out, b_out = prm[:6]
        n_layers_total = sum(layers_per_stack)
        layer_params = prm[6:6 + 14 * n_layers_total]
        agg_params = prm[6 + 14 * n_layers_total:]
        # frozen parameters get a scratch accumulator that is simply dropped
        grads = [ops._grad_buf(p) if p.requires_grad else torch.zeros_like(p) for p in params]
        g_in = grads[:6]
        g_layers = grads[6:6 + 14 * n_layers_total]
        g_agg = grads[6 + 14 * n_layers_total:]
        dz = ops._f32c(dz)
        P = dz.shape[0]
        q_scale = float(hd) ** -0.5

        dx = ops.linear_bwd(ctx.x_last, w_out, dz, g_in[4], g_in[5])           # (P, d): grad wrt last kept token
        L = 1
        li = len(ctx.saved)
        for si, (factor, nl) in reversed(list(enumerate(zip(factors, layers_per_stack)))):
            Lf = L * factor
            if linear_agg:
                dx = ops.linear_bwd(ctx.agg_in[si], agg_params[2 * si], dx, g_agg[2 * si], g_agg[2 * si + 1]).view(P * Lf, d)
                ctx.agg_in[si] = None
            elif nl == 0 and factor > 1:
                dx = _subsample_bwd(dx, P, Lf, factor, d)
            L = Lf
            for _ in range(nl):
                li -= 1
                lp = dict(zip(_LAYER_KEYS, layer_params[li * 14:(li + 1) * 14]))
                gp = dict(zip(_LAYER_KEYS, g_layers[li * 14:(li + 1) * 14]))
                x, qkv, probs, o_k, y1, mean1, rstd1, x1, hm, y2, mean2, rstd2, Ls, last, masks = ctx.saved[li]
                adrop = (masks[0], ds) if masks is not None else None
                assert Ls == L
                assert not (last and linear_agg)
                keep = factor if last else 1
                Mk = P * (L // keep)
                # (for a stack's last layer everything down to the out-proj lives on the surviving rows only)
                if ctx.tc:
                    # x, o_k, x1, hm are bf16 (hi, lo) pairs saved by the forward; gradients that feed GEMMs are produced
                    # pre-split by the LayerNorm-backward / attention-backward kernels and the GEMM epilogues.
                    Fh = lp['w1'].shape[0]
                    if masks is not None:
                        # dropout2: the FFN branch sees the masked gradient, the residual path the plain one
                        dy2 = ops.layernorm_bwd(dx, y2, mean2, rstd2, lp['g2'], gp['g2'], gp['be2'])
                        _, dy2_sp = ops.mask_scale_add(dy2, masks[3], ds, want_fp32=False, want_split=True)
                    else:
                        # linear2.bias gradient = column sums of dy2: accumulated by the LayerNorm-backward kernel itself
                        dy2, dy2_sp = ops.layernorm_bwd(dx, y2, mean2, rstd2, lp['g2'], gp['g2'], gp['be2'], want_split=True,
                                                        colsum_out=gp['b2'])
                    # (with dropout the bias gradients come from the masked splits, inside linear_bwd)
                    # FFN2: dhm = (dy2 W2) * (hm > 0) -> only its split is needed (dgrad + wgrad of FFN1); with dropout
                    # the saved hm is the dropped activation (zero where dropped) and the kept entries carry 1 / (1 - p)
                    _, dhm_sp = ops.linear_bwd(None, lp['w2'], None, gp['w2'], gp['b2'] if masks is not None else None,
                                               relu_mask=hm[0], x_split=hm,
                                               dy_split=dy2_sp, dx_want_c=False, dx_out_split=True, shape=(Mk, Fh),
                                               dx_scale=ds if masks is not None else None)
                    # FFN1: dx1 = dhm W1 + dy2
                    dx1 = ops.linear_bwd(None, lp['w1'], None, gp['w1'], gp['b1'], dx_residual=dy2, x_split=x1,
                                         dy_split=dhm_sp, shape=(Mk, d))
                    del dhm_sp
                    if masks is not None:
                        dy1 = ops.layernorm_bwd(dx1, y1, mean1, rstd1, lp['g1'], gp['g1'], gp['be1'])
                        _, dy1_sp = ops.mask_scale_add(dy1, masks[1], ds, want_fp32=False, want_split=True)
                    else:
                        dy1, dy1_sp = ops.layernorm_bwd(dx1, y1, mean1, rstd1, lp['g1'], gp['g1'], gp['be1'], want_split=True,
                                                        colsum_out=gp['bo'])      # out_proj.bias gradient, same way
                    do = ops.linear_bwd(None, lp['wo'], None, gp['wo'], gp['bo'] if masks is not None else None, x_split=o_k,
                                        dy_split=dy1_sp, shape=(Mk, d))
                    _, dqkv_sp = ops.attn_bwd(qkv, lp['e1'], lp['e2'], probs, do, gp['e1'], gp['e2'], P, L, H, hd, q_scale,
                                              keep=keep, want_fp32=False, want_split=True, drop=adrop)
                    if last:
                        dy1 = _subsample_bwd(dy1, P, L, factor, d)
                    dx = ops.linear_bwd(None, lp['wqkv'], None, gp['wqkv'], gp['bqkv'], dx_residual=dy1, x_split=x,
                                        dy_split=dqkv_sp, shape=(P * L, d))
                    ctx.saved[li] = None
                    continue
                # LN2 -> dy2 (flows to the FFN branch and, as the residual, to x1)
                dy2 = ops.layernorm_bwd(dx, y2, mean2, rstd2, lp['g2'], gp['g2'], gp['be2'])
                dy2m = ops.mask_scale_add(dy2, masks[3], ds)[0] if masks is not None else dy2
                # FFN2: dhm = (dy2 W2) * (hm > 0)   (with dropout: hm is the dropped activation, kept entries carry 1/(1-p))
                dhm = ops.linear_bwd(hm, lp['w2'], dy2m, gp['w2'], gp['b2'], relu_mask=hm, dx_scale=ds if masks is not None else None)
                # FFN1: dx1 = dhm W1 + dy2
                dx1 = ops.linear_bwd(x1, lp['w1'], dhm, gp['w1'], gp['b1'], dx_residual=dy2)
                del dhm
                dy1 = ops.layernorm_bwd(dx1, y1, mean1, rstd1, lp['g1'], gp['g1'], gp['be1'])
                dy1m = ops.mask_scale_add(dy1, masks[1], ds)[0] if masks is not None else dy1
                # out-proj (its dgrad is the gradient of the kept query rows' attention output)
                do = ops.linear_bwd(o_k, lp['wo'], dy1m, gp['wo'], gp['bo'])
                dqkv = ops.attn_bwd(qkv, lp['e1'], lp['e2'], probs, do, gp['e1'], gp['e2'], P, L, H, hd, q_scale, keep=keep,
                                    drop=adrop)
                if last:                       # residual path: scatter back to all L rows (zeros for the dropped rows)
                    dy1 = _subsample_bwd(dy1, P, L, factor, d)
                # in-proj: dx = dqkv Wqkv + dy1 (residual path)
                dx = ops.linear_bwd(x, lp['wqkv'], dqkv, gp['wqkv'], gp['bqkv'], dx_residual=dy1)
                ctx.saved[li] = None
        # input projection + positional embeddings
        M0 = P * S
        dh = dx
        ops._native.call('vqcpc_posemb_bwd', ops._p(dh), d, d - 2 * pe, ops._p(g_in[2]), ops._p(g_in[3]), M0, S, V, pe, ops._stream())
        dh_lin = dh[:, :d - 2 * pe]
        need_demb = ctx.needs_input_grad[0]
        demb = ops.linear_bwd(ctx.emb2d, w_in, dh_lin, g_in[0], g_in[1], need_dx=need_demb)
        ctx.saved = None
        ctx.agg_in = None
        ctx.emb2d = None
        ctx.x_last = None
        # parameter gradients were accumulated in place -> autograd gets None for them
        return (demb, None) + (None,) * len(params)
