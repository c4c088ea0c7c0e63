"""A stack of post-LN relative-attention encoder layers over whole sequences as ONE autograd node.

The layer is the reference's ``TransformerEncoderLayerCustom`` (transformer/transformer_custom.py:268-291 with the attention of
multihead_attention_custom.py:154-353 and the relative bias of subsampled_relative_attention.py:30-122) applied to sequences of
``L`` tokens - what the teacher (teachers/teacher_relative.py:33-47, L = 384) and the auxiliary decoder
(auxiliary_decoders/auxiliary_decoder_relative.py:58-74, L = 24 / 96) of the student trainer run.  Rows are sequence-major
(row = sequence * L + token); the reference's time-first layout differs only in indexing.

Same kernels and the same order of operations as ``downscalers/_encoder_stack.py`` (tcgen05 GEMMs with fused epilogues,
LayerNorm fwd / bwd, attention) - the attention op dispatches on ``L``: per-block kernels up to 16 tokens, the full-sequence
kernels of ``csrc/attn_long.cu`` above that.  No dropout here: the student configuration is run with p = 0 for parity, like
the headline encoder path (RNG streams cannot match the reference's)."""
import torch

from .. import ops

LAYER_KEYS = ('wqkv', 'bqkv', 'wo', 'bo', 'e1', 'e2', 'w1', 'b1', 'w2', 'b2', 'g1', 'be1', 'g2', 'be2')
# state_dict names of one layer (prefix 'layers.{l}.'), in LAYER_KEYS order
LAYER_PARAM_NAMES = ('self_attn.in_proj_weight', 'self_attn.in_proj_bias', 'self_attn.out_proj.weight', 'self_attn.out_proj.bias',
                     'self_attn.attn_bias.e1', 'self_attn.attn_bias.e2', 'linear1.weight', 'linear1.bias', 'linear2.weight',
                     'linear2.bias', 'norm1.weight', 'norm1.bias', 'norm2.weight', 'norm2.bias')


class TransformerLayersFn(torch.autograd.Function):
    """x (P * L, d) -> layers -> (P * L, d); meta = (P, L, d, H, n_layers, grad_enabled)."""

    @staticmethod
    def forward(ctx, x, meta, *params):
        P, L, d, H, n_layers, grad_on = meta
        hd = d // H
        x = ops._f32c(x)
        assert x.shape == (P * L, d)
        prm = [p.detach() for p in params]
        train = any(ctx.needs_input_grad) and bool(grad_on)
        tc = ops.tensor_core_mode()
        q_scale = float(hd) ** -0.5
        saved = []
        x_sp = None
        M = P * L
        for li in range(n_layers):
            lp = dict(zip(LAYER_KEYS, prm[li * 14:(li + 1) * 14]))
            Fh = lp['w1'].shape[0]
            if tc:
                if x_sp is None:
                    x_sp = ops.split_bf16(x)
                qkv = ops.gemm(None, lp['wqkv'], M, 3 * d, d, bias=lp['bqkv'], scale_cols=(q_scale, d), b_is_weight=True,
                               a_split=x_sp)
                _, o_sp, probs = ops.attn_fwd(qkv, lp['e1'], lp['e2'], P, L, H, hd, need_probs=train, want_fp32=False,
                                              want_split=True)
                y1 = ops.gemm(None, lp['wo'], M, d, d, bias=lp['bo'], residual=x, b_is_weight=True, a_split=o_sp)
                x1, x1_sp, mean1, rstd1 = ops.layernorm_fwd(y1, lp['g1'], lp['be1'], want_split=True)
                _, hm_sp = ops.gemm(None, lp['w1'], M, Fh, d, bias=lp['b1'], relu=True, b_is_weight=True, a_split=x1_sp,
                                    want_c=False, out_split=True)
                y2 = ops.gemm(None, lp['w2'], M, d, Fh, bias=lp['b2'], residual=x1, b_is_weight=True, a_split=hm_sp)
                x2, x2_sp, mean2, rstd2 = ops.layernorm_fwd(y2, lp['g2'], lp['be2'], want_split=True)
                if train:
                    saved.append((x_sp, qkv, probs, o_sp, y1, mean1, rstd1, x1_sp, hm_sp, y2, mean2, rstd2))
                x, x_sp = x2, x2_sp
            else:
                qkv = ops.linear_fwd(x, lp['wqkv'], lp['bqkv'], scale_cols=(q_scale, d))
                o, probs = ops.attn_fwd(qkv, lp['e1'], lp['e2'], P, L, H, hd, need_probs=train)
                y1 = ops.linear_fwd(o, lp['wo'], lp['bo'], residual=x)
                x1, mean1, rstd1 = ops.layernorm_fwd(y1, lp['g1'], lp['be1'])
                hm = ops.linear_fwd(x1, lp['w1'], lp['b1'], relu=True)
                y2 = ops.linear_fwd(hm, lp['w2'], lp['b2'], residual=x1)
                x2, mean2, rstd2 = ops.layernorm_fwd(y2, lp['g2'], lp['be2'])
                if train:
                    saved.append((x, qkv, probs, o, y1, mean1, rstd1, x1, hm, y2, mean2, rstd2))
                x = x2
        ctx.meta = meta
        ctx.params = params
        ctx.tc = tc
        ctx.saved = saved
        return x

    @staticmethod
    def backward(ctx, dx):
        P, L, d, H, n_layers, _ = ctx.meta
        hd = d // H
        params = ctx.params
        prm = [p.detach() for p in params]
        grads = [ops._grad_buf(p) if p.requires_grad else torch.zeros_like(p) for p in params]
        q_scale = float(hd) ** -0.5
        M = P * L
        dx = ops._f32c(dx)
        for li in reversed(range(n_layers)):
            lp = dict(zip(LAYER_KEYS, prm[li * 14:(li + 1) * 14]))
            gp = dict(zip(LAYER_KEYS, grads[li * 14:(li + 1) * 14]))
            x, qkv, probs, o, y1, mean1, rstd1, x1, hm, y2, mean2, rstd2 = ctx.saved[li]
            Fh = lp['w1'].shape[0]
            if ctx.tc:
                dy2, dy2_sp = ops.layernorm_bwd(dx, y2, mean2, rstd2, lp['g2'], gp['g2'], gp['be2'], want_split=True,
                                                colsum_out=gp['b2'])
                _, dhm_sp = ops.linear_bwd(None, lp['w2'], None, gp['w2'], None, relu_mask=hm[0], x_split=hm, dy_split=dy2_sp,
                                           dx_want_c=False, dx_out_split=True, shape=(M, Fh))
                dx1 = ops.linear_bwd(None, lp['w1'], None, gp['w1'], gp['b1'], dx_residual=dy2, x_split=x1, dy_split=dhm_sp,
                                     shape=(M, d))
                del dhm_sp
                dy1, dy1_sp = ops.layernorm_bwd(dx1, y1, mean1, rstd1, lp['g1'], gp['g1'], gp['be1'], want_split=True,
                                                colsum_out=gp['bo'])
                do = ops.linear_bwd(None, lp['wo'], None, gp['wo'], None, x_split=o, dy_split=dy1_sp, shape=(M, d))
                _, dqkv_sp = ops.attn_bwd(qkv, lp['e1'], lp['e2'], probs, do, gp['e1'], gp['e2'], P, L, H, hd, q_scale,
                                          want_fp32=False, want_split=True)
                dx = ops.linear_bwd(None, lp['wqkv'], None, gp['wqkv'], gp['bqkv'], dx_residual=dy1, x_split=x, dy_split=dqkv_sp,
                                    shape=(M, d))
            else:
                dy2 = ops.layernorm_bwd(dx, y2, mean2, rstd2, lp['g2'], gp['g2'], gp['be2'])
                dhm = ops.linear_bwd(hm, lp['w2'], dy2, gp['w2'], gp['b2'], relu_mask=hm)
                dx1 = ops.linear_bwd(x1, lp['w1'], dhm, gp['w1'], gp['b1'], dx_residual=dy2)
                del dhm
                dy1 = ops.layernorm_bwd(dx1, y1, mean1, rstd1, lp['g1'], gp['g1'], gp['be1'])
                do = ops.linear_bwd(o, lp['wo'], dy1, gp['wo'], gp['bo'])
                dqkv = ops.attn_bwd(qkv, lp['e1'], lp['e2'], probs, do, gp['e1'], gp['e2'], P, L, H, hd, q_scale)
                dx = ops.linear_bwd(x, lp['wqkv'], dqkv, gp['wqkv'], gp['bqkv'], dx_residual=dy1)
            ctx.saved[li] = None
        ctx.saved = None
        return (dx, None) + (None,) * len(params)


class _Linear(torch.nn.Module):
    """nn.Linear parameters (same names / init) evaluated by the library's GEMMs."""

    def __init__(self, in_features, out_features):
        super().__init__()
        ref = torch.nn.Linear(in_features, out_features)
        self.weight = torch.nn.Parameter(ref.weight.detach().clone())
        self.bias = torch.nn.Parameter(ref.bias.detach().clone())

    def forward(self, x2d):
        return LinearFn.apply(x2d, self.weight, self.bias)


class LinearFn(torch.autograd.Function):
    """y = x W^T + b on 2-D fp32 rows (the row stride may exceed the width: channel slices of a token matrix)."""

    @staticmethod
    def forward(ctx, x, w, b):
        x = x if (x.dtype == torch.float32 and x.stride(1) == 1) else ops._f32c(x)
        ctx.save_for_backward(x, w)
        ctx.has_bias = b is not None
        ctx.params = (w, b)
        return ops.linear_fwd(x, w.detach(), b.detach() if b is not None else None)

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        wp, bp = ctx.params
        dy = ops._f32c(dy)
        gw = ops._grad_buf(wp) if wp.requires_grad else torch.zeros_like(wp)
        gb = (ops._grad_buf(bp) if bp.requires_grad else torch.zeros_like(bp)) if bp is not None else None
        dx = ops.linear_bwd(x, w.detach(), dy, gw, gb, need_dx=ctx.needs_input_grad[0])
        return dx, None, None


class TransformerEncoderStack(torch.nn.Module):
    """Parameter container with the reference's names (``layers.{l}.self_attn.in_proj_weight`` ... as in
    ``TransformerEncoderCustom`` of ``TransformerEncoderLayerCustom``, transformer_custom.py:142-163,237-266) + forward through
    ``TransformerLayersFn``.  ``seq_len`` = num_events * num_channels sizes the relative tables e1 / e2 (H * seq_len, hd)."""

    def __init__(self, d_model, nhead, dim_feedforward, seq_len, num_layers):
        super().__init__()
        self.d_model, self.nhead, self.seq_len, self.num_layers = d_model, nhead, seq_len, num_layers
        hd = d_model // nhead
        self.layers = torch.nn.ModuleList()
        for _ in range(num_layers):
            layer = torch.nn.Module()
            attn = torch.nn.Module()
            mha = torch.nn.MultiheadAttention(d_model, nhead)          # init only (xavier in-proj, zero biases)
            attn.in_proj_weight = torch.nn.Parameter(mha.in_proj_weight.detach().clone())
            attn.in_proj_bias = torch.nn.Parameter(mha.in_proj_bias.detach().clone())
            attn.out_proj = _Linear(d_model, d_model)
            torch.nn.init.constant_(attn.out_proj.bias, 0.)
            ab = torch.nn.Module()
            ab.e1 = torch.nn.Parameter(torch.randn(nhead * seq_len, hd))      # subsampled_relative_attention.py:20-21
            ab.e2 = torch.nn.Parameter(torch.randn(nhead * seq_len, hd))
            attn.attn_bias = ab
            layer.self_attn = attn
            layer.linear1 = _Linear(d_model, dim_feedforward)
            layer.linear2 = _Linear(dim_feedforward, d_model)
            layer.norm1 = torch.nn.LayerNorm(d_model)
            layer.norm2 = torch.nn.LayerNorm(d_model)
            self.layers.append(layer)

    def flat_params(self):
        out = []
        for layer in self.layers:
            sd = dict(layer.named_parameters())
            out += [sd[n] for n in LAYER_PARAM_NAMES]
        return out

    def forward(self, x2d, num_sequences):
        """x2d (num_sequences * seq_len, d_model), sequence-major rows."""
        meta = (int(num_sequences), self.seq_len, self.d_model, self.nhead, self.num_layers, torch.is_grad_enabled())
        return TransformerLayersFn.apply(x2d, meta, *self.flat_params())
