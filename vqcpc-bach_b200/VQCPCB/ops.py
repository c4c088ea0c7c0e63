"""Thin PyTorch-side wrappers over the C-ABI kernels (libvqcpc_b200.so) + the autograd Functions built on them.

PyTorch is used for device memory, streams and autograd bookkeeping only; every numeric op of the hot path is one
of the hand-written sm_100a kernels.  Nothing here has a CPU or eager fallback: tensors must live on a CUDA device.

GEMM arithmetic modes (``set_gemm_mode``):
  'fp32'   exact fp32 FMA on CUDA cores (vqcpc_gemm_f32)
  'bf16x3' tcgen05 tensor cores, operands split into bf16 hi+lo, 3 MMAs per tile, fp32 TMEM accumulation (parity mode)
  'bf16'   tcgen05, single pass on the hi halves (fast mode; misses the 1e-4 parity bar, reported separately)
Small / oddly shaped GEMMs (K < 32, N < 32, unaligned leading dimensions) always run the fp32 kernel.
"""
import ctypes
import os

import torch

from . import _native

EPI_BIAS, EPI_RELU, EPI_ACCUMULATE, EPI_RESIDUAL, EPI_RELU_MASK, EPI_SCALE_COLS, EPI_MASK_BF16 = 1, 2, 4, 8, 16, 64, 128

_GEMM_MODE = 'bf16x3'
PROFILE_GEMM = None         # bench.py sets this to a list: every GEMM launch is bracketed by CUDA events on its stream
_WSPLIT_CACHE = {}          # id(param storage) -> (version, hi, lo)


def set_tile_config(cfg):
    """Force the tcgen05 tile configuration (see vqcpc_gemm_set_tile_config in include/vqcpc_b200.h): 0 .. 5, -1 = automatic."""
    _native.call('vqcpc_gemm_set_tile_config', int(cfg))


if os.environ.get('VQCPC_GEMM_TILE') is not None:
    set_tile_config(int(os.environ['VQCPC_GEMM_TILE']))


def set_gemm_mode(mode):
    global _GEMM_MODE
    assert mode in ('fp32', 'bf16x3', 'bf16'), mode
    _GEMM_MODE = mode
    _WSPLIT_CACHE.clear()


def get_gemm_mode():
    return _GEMM_MODE


def _p(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _chk(*ts):
    for t in ts:
        if t is not None:
            if not t.is_cuda:
                raise RuntimeError('vqcpc_b200 ops need CUDA tensors (there is no CPU fallback)')


def _f32c(t):
    if t.dtype != torch.float32:
        t = t.float()
    return t if t.is_contiguous() else t.contiguous()


# ------------------------------------------------------------------------------------------------ GEMM
def split_bf16(x2d):
    """(rows, cols) fp32 (row stride = x2d.stride(0)) -> contiguous bf16 (hi, lo) with x ~= hi + lo."""
    rows, cols = x2d.shape
    assert x2d.stride(1) == 1
    hi = torch.empty(rows, cols, dtype=torch.bfloat16, device=x2d.device)
    lo = torch.empty(rows, cols, dtype=torch.bfloat16, device=x2d.device)
    _native.call('vqcpc_split_bf16', _p(x2d), x2d.stride(0), _p(hi), _p(lo), cols, rows, cols, _stream())
    return hi, lo


def weight_split(w):
    """Cached bf16 split of a parameter (invalidated by the in-place optimizer update via ``bump_weight_version``)."""
    key = (w.data_ptr(), tuple(w.shape))
    # two invalidation signals: the global counter (bumped by the fused optimizer / loaders, whose kernels write through
    # raw pointers that autograd does not see) and the tensor's own in-place version counter (load_state_dict, copy_,
    # a stock torch optimizer ...), so no in-place update can leave a stale split behind
    ver = (_WEIGHT_VERSION[0], w._version)
    ent = _WSPLIT_CACHE.get(key)
    if ent is not None and ent[0] == ver:
        return ent[1], ent[2]
    w2 = w.detach().reshape(w.shape[0], -1)
    hi, lo = split_bf16(w2)
    _WSPLIT_CACHE[key] = (ver, hi, lo)
    return hi, lo


_WEIGHT_VERSION = [0]


def weight_version():
    return _WEIGHT_VERSION[0]


def bump_weight_version():
    """Call after parameters change (optimizer step / load_state_dict) so cached bf16 splits are rebuilt."""
    _WEIGHT_VERSION[0] += 1
    if len(_WSPLIT_CACHE) > 256:
        _WSPLIT_CACHE.clear()


def _tc_ok(M, N, K, a, b, a_kmajor, b_kmajor):
    if _GEMM_MODE == 'fp32':
        return False
    if K < 32 or N < 32 or M < 1:
        return False
    # TMA needs 16-byte aligned leading dimensions of the (contiguous) bf16 copies
    a_cols = K if a_kmajor else M
    b_cols = K if b_kmajor else N
    return a_cols % 8 == 0 and b_cols % 8 == 0 and a_cols % 4 == 0 and b_cols % 4 == 0


def tensor_core_mode():
    return _GEMM_MODE != 'fp32'


def gemm(a, b, M, N, K, a_kmajor=True, b_kmajor=True, out=None, bias=None, residual=None, relu=False, relu_mask=None,
         accumulate=False, scale_cols=None, ksplit=1, b_is_weight=False, a_split=None, b_split=None, want_c=True,
         out_split=False):
    """C[M,N] = epi(A.B).  a/b are 2-D fp32 views whose unit-stride dim is K (kmajor) or M/N (not kmajor); alternatively
    pass the operand already split into bf16 (a_split / b_split = (hi, lo)) and leave a / b None.

    epilogue: +bias[N] -> (cols < n) *= alpha  [scale_cols=(alpha, n)] -> +residual[M,N] -> relu -> *(relu_mask > 0)
    (relu_mask may be fp32 or the bf16 hi half of a split activation).  accumulate: C += result.
    ksplit > 1: split-K with fp32 atomics (linear epilogues only).
    out_split: additionally return the result split into bf16 (hi, lo) written by the epilogue (tensor-core path only);
    want_c=False skips the fp32 result.  Returns C, or (C | None, (hi, lo)) when out_split."""
    _chk(a, b, out, bias, residual, relu_mask)
    dev = (a if a is not None else a_split[0]).device
    use_tc = (a is None or b is None) or _tc_ok(M, N, K, a, b, a_kmajor, b_kmajor)
    if out_split or not want_c:
        assert use_tc, 'split outputs need the tensor-core GEMM'
    if out is None and want_c:
        out = torch.empty(M, N, dtype=torch.float32, device=dev)
    if out is not None:
        assert out.stride(1) == 1
    flags = 0
    if bias is not None:
        flags |= EPI_BIAS
    if relu:
        flags |= EPI_RELU
    if accumulate:
        flags |= EPI_ACCUMULATE
    res = residual
    if residual is not None:
        flags |= EPI_RESIDUAL
        assert relu_mask is None
    if relu_mask is not None:
        flags |= EPI_RELU_MASK
        if relu_mask.dtype == torch.bfloat16:
            flags |= EPI_MASK_BF16
        res = relu_mask
    alpha, ncols = 1.0, 0
    if scale_cols is not None:
        flags |= EPI_SCALE_COLS
        alpha, ncols = scale_cols
    ldr = res.stride(0) if res is not None else 0
    prof = PROFILE_GEMM
    if use_tc:
        if a_split is None:
            assert a.stride(1) == 1
            a_split = split_bf16(a)
        if b_split is None:
            assert b.stride(1) == 1
            b_split = weight_split(b) if b_is_weight else split_bf16(b)
    if prof is not None:
        ev0 = torch.cuda.Event(enable_timing=True)
        ev0.record()
    oh = ol = None
    if use_tc:
        passes = 3 if _GEMM_MODE == 'bf16x3' else 1
        if out_split:
            oh = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
            ol = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
        _native.call('vqcpc_gemm_bf16x3_ex', int(a_kmajor), int(b_kmajor), M, N, K, _p(a_split[0]), _p(a_split[1]),
                     a_split[0].stride(0), _p(b_split[0]), _p(b_split[1]), b_split[0].stride(0), _p(out),
                     out.stride(0) if out is not None else 0, _p(oh), _p(ol), N, _p(bias), _p(res), ldr, flags,
                     float(alpha), int(ncols), passes, int(ksplit), _stream())
    else:
        assert a.stride(1) == 1 and b.stride(1) == 1
        _native.call('vqcpc_gemm_f32', int(a_kmajor), int(b_kmajor), M, N, K, _p(a), a.stride(0), _p(b), b.stride(0),
                     _p(out), out.stride(0), _p(bias), _p(res), ldr, flags, float(alpha), int(ncols), int(ksplit), _stream())
    if prof is not None:
        ev1 = torch.cuda.Event(enable_timing=True)
        ev1.record()
        prof.append(('tc' if use_tc else 'simt', 2.0 * M * N * K, ev0, ev1))
    if out_split:
        return out, (oh, ol)
    return out


def _wgrad_ksplit(M_red, n_out_tiles):
    """split-K factor for a weight-gradient GEMM (reduction over M_red rows) so that ~2 waves of CTAs are busy."""
    sms = torch.cuda.get_device_properties(torch.cuda.current_device()).multi_processor_count
    want = max(1, (2 * sms) // max(1, n_out_tiles))
    return int(max(1, min(want, M_red // 256 if M_red >= 256 else 1)))


def linear_fwd(x2d, w, b, relu=False, residual=None, scale_cols=None, out=None):
    """y = x @ w.T + b  (nn.Linear weight layout (N, K))."""
    M, K = x2d.shape
    N = w.shape[0]
    return gemm(x2d, w, M, N, K, True, True, out=out, bias=b, relu=relu, residual=residual, scale_cols=scale_cols,
                b_is_weight=True)


def linear_bwd(x2d, w, dy, dw, db, need_dx=True, dx_residual=None, relu_mask=None, dx_out=None, x_split=None,
               dy_split=None, dx_want_c=True, dx_out_split=False, shape=None, dx_scale=None):
    """Gradients of y = x @ w.T + b.  dw/db are ACCUMULATED in place (flat-gradient-buffer views).
    dx = dx_scale * dy @ w (+ dx_residual) (* relu_mask > 0).  x / dy may be given pre-split (x_split / dy_split, with
    shape=(M, K)); db then comes from the split (hi + lo) unless an fp32 dy is also given."""
    M, K = shape if shape is not None else x2d.shape
    N = w.shape[0]
    dx = None
    if need_dx:
        dx = gemm(dy if dy_split is None else None, w, M, K, N, True, False, out=dx_out, residual=dx_residual,
                  relu_mask=relu_mask, b_is_weight=True, a_split=dy_split, want_c=dx_want_c, out_split=dx_out_split,
                  scale_cols=None if dx_scale is None else (float(dx_scale), K))
    if dw is not None:
        tiles = ((N + 127) // 128) * ((K + 127) // 128)
        gemm(dy if dy_split is None else None, x2d if x_split is None else None, N, K, M, False, False, out=dw.view(N, K),
             accumulate=True, ksplit=_wgrad_ksplit(M, tiles), a_split=dy_split, b_split=x_split)
    if db is not None:
        if dy is not None:
            colsum(dy, db)
        else:
            _native.call('vqcpc_colsum_split', _p(dy_split[0]), _p(dy_split[1]), dy_split[0].stride(0), _p(db), M, N, _stream())
    return dx


def colsum(x2d, out_accum):
    _native.call('vqcpc_colsum', _p(x2d), x2d.stride(0), _p(out_accum), x2d.shape[0], x2d.shape[1], _stream())


# ------------------------------------------------------------------------------------------------ small ops
def _ptr_array(tensors):
    arr = (ctypes.c_void_p * len(tensors))(*[t.data_ptr() if t is not None else 0 for t in tensors])
    return arr


def embed_fwd(tokens_flat, tables):
    n = tokens_flat.numel()
    E = tables[0].shape[1]
    out = torch.empty(n, E, dtype=torch.float32, device=tokens_flat.device)
    rows = (ctypes.c_int * len(tables))(*[t.shape[0] for t in tables])
    _native.call('vqcpc_embed_fwd', _p(tokens_flat), n, len(tables), _ptr_array(tables), rows, E, _p(out), _stream())
    return out


def embed_bwd(tokens_flat, dout, grads):
    n = tokens_flat.numel()
    E = dout.shape[-1]
    rows = (ctypes.c_int * len(grads))(*[g.shape[0] for g in grads])
    _native.call('vqcpc_embed_bwd', _p(tokens_flat), n, len(grads), _ptr_array(grads), rows, E, _p(dout), _stream())


def _split_buffers(rows, cols, dev):
    return (torch.empty(rows, cols, dtype=torch.bfloat16, device=dev), torch.empty(rows, cols, dtype=torch.bfloat16, device=dev))


def layernorm_fwd(y, gamma, beta, eps=1e-5, want_fp32=True, want_split=False):
    """-> (out fp32 | None, (hi, lo) | None, mean, rstd)"""
    M, d = y.shape
    out = torch.empty_like(y) if want_fp32 else None
    sp = _split_buffers(M, d, y.device) if want_split else (None, None)
    mean = torch.empty(M, dtype=torch.float32, device=y.device)
    rstd = torch.empty(M, dtype=torch.float32, device=y.device)
    _native.call('vqcpc_layernorm_fwd', _p(y), _p(gamma), _p(beta), _p(out), _p(sp[0]), _p(sp[1]), _p(mean), _p(rstd), M, d,
                 float(eps), _stream())
    if want_split:
        return out, sp, mean, rstd
    return out, mean, rstd


def layernorm_bwd(dout, y, mean, rstd, gamma, dgamma, dbeta, want_split=False, colsum_out=None):
    """colsum_out (d,), optional: accumulates the column sums of the returned dy (bias gradient of the layer feeding the LN)."""
    M, d = y.shape
    dy = torch.empty_like(y)
    sp = _split_buffers(M, d, y.device) if want_split else (None, None)
    _native.call('vqcpc_layernorm_bwd_ex', _p(dout), _p(y), _p(mean), _p(rstd), _p(gamma), _p(dy), _p(sp[0]), _p(sp[1]), _p(dgamma),
                 _p(dbeta), _p(colsum_out), M, d, _stream())
    if want_split:
        return dy, sp
    return dy


_ATTN_ALGO = {'fwd': None, 'bwd': None}          # kernels currently selected in the library (None = not set yet)


def _select_attn_kernel(which, algo=None):
    """Attention kernels: tensor cores (mma.sync 3xTF32, ~2^-20 per product) in the tensor-core GEMM modes, the CUDA-core
    fp32 kernels in the exact 'fp32' mode (whose parity tests hold every gradient to the fp32 noise floor).
    VQCPC_ATTN_FWD_ALGO / VQCPC_ATTN_BWD_ALGO = 0|1 override (tuning hooks); ``algo`` overrides both (tests)."""
    if algo is None:
        env = os.environ.get('VQCPC_ATTN_%s_ALGO' % which.upper())
        algo = int(env) if env else (1 if _GEMM_MODE == 'fp32' else 0)
    if _ATTN_ALGO[which] != algo:
        _native.call('vqcpc_attn_set_%s_algo' % which, int(algo))
        _ATTN_ALGO[which] = int(algo)


def attn_fwd(qkv, e1, e2, P, L, H, hd, need_probs=True, keep=1, want_fp32=True, want_split=False, drop=None, algo=None):
    """-> (out fp32 | None, probs) or (out | None, (hi, lo), probs) when want_split; rows = P * L / keep.
    drop = (keep_mask uint8 (P, H, L, L), scale): dropout on the attention weights.
    algo: None = by GEMM mode (see _select_attn_kernel), 0 = tensor cores, 1 = CUDA cores."""
    if L > 16:
        # whole-sequence attention (teacher / auxiliary decoder of the student trainer): csrc/attn_long.cu
        assert keep == 1 and drop is None, 'full-sequence attention has no dead-row elimination / attention dropout'
        out = torch.empty(P * L, H * hd, dtype=torch.float32, device=qkv.device) if want_fp32 else None
        sp = _split_buffers(P * L, H * hd, qkv.device) if want_split else (None, None)
        probs = torch.empty(P, H, L, L, dtype=torch.float32, device=qkv.device) if need_probs else None
        _native.call('vqcpc_attn_long_fwd', _p(qkv), _p(e1), _p(e2), _p(out), _p(sp[0]), _p(sp[1]), _p(probs), P, L, H, hd, _stream())
        return (out, sp, probs) if want_split else (out, probs)
    _select_attn_kernel('fwd', algo)
    Mo = P * (L // keep)
    out = torch.empty(Mo, H * hd, dtype=torch.float32, device=qkv.device) if want_fp32 else None
    sp = _split_buffers(Mo, H * hd, qkv.device) if want_split else (None, None)
    probs = torch.empty(P, H, L, L, dtype=torch.float32, device=qkv.device) if need_probs else None
    _native.call('vqcpc_attn_fwd_ex', _p(qkv), _p(e1), _p(e2), _p(out), _p(sp[0]), _p(sp[1]), _p(probs), P, L, H, hd, int(keep),
                 _p(drop[0]) if drop else None, float(drop[1]) if drop else 1.0, _stream())
    if want_split:
        return out, sp, probs
    return out, probs


def attn_bwd(qkv, e1, e2, probs, dout, de1, de2, P, L, H, hd, q_scale, keep=1, want_fp32=True, want_split=False, drop=None,
             algo=None):
    if L > 16:
        assert keep == 1 and drop is None
        dqkv = torch.empty_like(qkv) if want_fp32 else None
        sp = _split_buffers(qkv.shape[0], qkv.shape[1], qkv.device) if want_split else (None, None)
        ds = torch.empty_like(probs)
        _native.call('vqcpc_attn_long_bwd', _p(qkv), _p(e1), _p(e2), _p(probs), _p(_f32c(dout)), _p(ds), _p(dqkv), _p(sp[0]), _p(sp[1]),
                     _p(de1), _p(de2), P, L, H, hd, float(q_scale), _stream())
        return (dqkv, sp) if want_split else dqkv
    _select_attn_kernel('bwd', algo)
    dqkv = torch.empty_like(qkv) if want_fp32 else None
    sp = _split_buffers(qkv.shape[0], qkv.shape[1], qkv.device) if want_split else (None, None)
    _native.call('vqcpc_attn_bwd_ex', _p(qkv), _p(e1), _p(e2), _p(probs), _p(dout), _p(dqkv), _p(sp[0]), _p(sp[1]), _p(de1), _p(de2),
                 P, L, H, hd, float(q_scale), int(keep), _p(drop[0]) if drop else None, float(drop[1]) if drop else 1.0, _stream())
    if want_split:
        return dqkv, sp
    return dqkv


def mask_scale_add(x, keep, scale, residual=None, want_fp32=True, want_split=False):
    """(keep ? x * scale : 0) (+ residual) -> (fp32 | None, (hi, lo) | None): the dropout sites of the encoder layer."""
    x = _f32c(x)
    out = torch.empty_like(x) if want_fp32 else None
    sp = _split_buffers(x.shape[0], x.shape[1], x.device) if want_split else (None, None)
    _native.call('vqcpc_mask_scale_add', _p(x), _p(keep), float(scale), _p(residual), _p(out), _p(sp[0]), _p(sp[1]), x.numel(),
                 _stream())
    return out, (sp if want_split else None)


# Tests install a callable (site, shape, p, device) -> uint8 keep mask here to record / replay the dropout draws of the
# fused encoder; None = draw from torch's CUDA generator.
DROPOUT_MASK_HOOK = None


def dropout_mask(site, shape, p, device):
    if DROPOUT_MASK_HOOK is not None:
        return DROPOUT_MASK_HOOK(site, tuple(shape), p, device)
    return draw_keep_mask(shape, p, device)


def softmax_ce(logits, target_idx=None, target_logits=None, scale=1.0, want_grad=True):
    """Row-wise softmax cross-entropy (hard int64 targets or distilled = a teacher's logits) -> (loss (rows,), scale * dloss/dlogits)."""
    logits = _f32c(logits)
    rows, C = logits.shape
    loss = torch.empty(rows, dtype=torch.float32, device=logits.device)
    dl = torch.empty_like(logits) if want_grad else None
    if target_logits is not None:
        target_logits = _f32c(target_logits)
    _native.call('vqcpc_softmax_ce', _p(logits), C, _p(target_idx), _p(target_logits), C, _p(loss), _p(dl), C, float(scale), rows, C,
                 _stream())
    return loss, dl


def reduce_sum(x, scale=1.0):
    out = torch.empty(1, dtype=torch.float32, device=x.device)
    _native.call('vqcpc_reduce_sum', _p(x), x.numel(), _p(out), float(scale), _stream())
    return out


def code_hist(idx, counts):
    """counts[idx[i]] += 1 (int32, accumulated) over all elements of the int64 index tensor."""
    idx = idx.reshape(-1)
    if not idx.is_contiguous():
        idx = idx.contiguous()
    _chk(idx, counts)
    _native.call('vqcpc_code_hist', _p(idx), idx.numel(), counts.numel(), _p(counts), _stream())


def table_scatter_add(idx, rows, out):
    """out[idx[i]] += rows[i] (rows (n, w) fp32, out (K, w) accumulated): deterministic, no atomics."""
    idx = idx.reshape(-1)
    idx = idx if idx.is_contiguous() else idx.contiguous()
    rows = _f32c(rows)
    _chk(idx, rows, out)
    _native.call('vqcpc_table_scatter_add', _p(idx), idx.numel(), _p(rows), rows.shape[1], out.shape[0], _p(out), _stream())


# ------------------------------------------------------------------------------------------------ autograd Functions
def _grad_buf(p):
    """Gradient accumulator of a leaf parameter (allocated zeroed on first use; kernels accumulate in place)."""
    if p.grad is None:
        p.grad = torch.zeros_like(p)
    return p.grad


class _ParamGradSink:
    """Kernels accumulate parameter gradients directly into ``param.grad`` (which the trainer aliases to a flat
    buffer); autograd is told the gradient is None.  ``enabled`` lets tests ask for returned gradients instead."""
    direct = True


class EmbedFn(torch.autograd.Function):
    """data_processor/bach_cpc_data_processor.py:42-68 — per-voice embedding lookup."""

    @staticmethod
    def forward(ctx, tokens, *tables):
        flat = tokens.reshape(-1).contiguous()
        out = embed_fwd(flat, [t.detach() for t in tables])
        ctx.save_for_backward(flat)
        ctx.tables = tables
        return out.view(*tokens.shape, tables[0].shape[1])

    @staticmethod
    def backward(ctx, dout):
        (flat,) = ctx.saved_tensors
        tables = ctx.tables
        dout = _f32c(dout).view(-1, dout.shape[-1])
        grads = [torch.zeros_like(t) for t in tables]
        embed_bwd(flat, dout, grads)
        return (None, *grads)


class LinearFn(torch.autograd.Function):
    """y = x W^T + b on the hand-written GEMMs (used by the upscaler MLP and the c-net output layer)."""

    @staticmethod
    def forward(ctx, x, w, b):
        x2 = _f32c(x).view(-1, x.shape[-1])
        y = linear_fwd(x2, w.detach(), None if b is None else b.detach())
        ctx.save_for_backward(x2, w, b)
        ctx.xshape = x.shape
        return y.view(*x.shape[:-1], w.shape[0])

    @staticmethod
    def backward(ctx, dy):
        x2, w, b = ctx.saved_tensors
        dy2 = _f32c(dy).view(-1, dy.shape[-1])
        dw = torch.zeros_like(w) if ctx.needs_input_grad[1] else None
        db = torch.zeros_like(b) if (b is not None and ctx.needs_input_grad[2]) else None
        dx = linear_bwd(x2, w.detach(), dy2, dw, db, need_dx=ctx.needs_input_grad[0])
        return (dx.view(ctx.xshape) if dx is not None else None), dw, db


class SeluFn(torch.autograd.Function):
    """upscalers/mlp_upscaler.py:24 nn.SELU."""

    @staticmethod
    def forward(ctx, x):
        x = _f32c(x)
        out = torch.empty_like(x)
        _native.call('vqcpc_selu_fwd', _p(x), _p(out), x.numel(), _stream())
        ctx.save_for_backward(x)
        return out

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        dy = _f32c(dy)
        dx = torch.empty_like(x)
        _native.call('vqcpc_selu_bwd', _p(x), _p(dy), _p(dx), x.numel(), _stream())
        return dx


class DropoutFn(torch.autograd.Function):
    """nn.Dropout(p) in training mode with an explicit keep mask (uint8): y = x * keep / (1 - p)  (mlp_upscaler.py:23).
    The mask is drawn by the caller (``draw_keep_mask``) so tests can replay the exact draw through a reference implementation."""

    @staticmethod
    def forward(ctx, x, keep, p):
        x = _f32c(x)
        out = torch.empty_like(x)
        scale = 1.0 / (1.0 - p)
        _native.call('vqcpc_mask_scale', _p(x), _p(keep), scale, _p(out), x.numel(), _stream())
        ctx.save_for_backward(keep)
        ctx.scale = scale
        return out

    @staticmethod
    def backward(ctx, dy):
        (keep,) = ctx.saved_tensors
        dy = _f32c(dy)
        dx = torch.empty_like(dy)
        _native.call('vqcpc_mask_scale', _p(dy), _p(keep), ctx.scale, _p(dx), dy.numel(), _stream())
        return dx, None, None


def draw_keep_mask(shape, p, device):
    """Bernoulli(1 - p) keep mask as uint8 from torch's CUDA generator (reproducible with torch.manual_seed)."""
    return torch.empty(shape, dtype=torch.uint8, device=device).bernoulli_(1.0 - p)     # one kernel, one byte per element


class GRUFn(torch.autograd.Function):
    """Multi-layer GRU, batch_first, zero initial state (torch.nn.GRU semantics; CModule vqcpc_helper.py:54-76 and
    LstmDownscaler lstm_downscaler.py:31-48) on the native GEMMs + the point-wise gate kernels:

        per layer:  GI = X W_ih^T + b_ih (one GEMM over all (b, t) rows);  per step  GH = h_{t-1} W_hh^T + b_hh (GEMM),
                    gate kernel -> h_t written into out[:, t] (the next step's GEMM reads it there, row stride T * H)
        backward :  per step gate kernel + dh_{t-1} = dh_t z + dGH_t W_hh and dW_hh += dGH_t^T h_{t-1} (two GEMMs);
                    per layer dW_ih, db_ih, db_hh, dX from the stacked dGI / dGH

    meta = (num_layers, hidden, p, masks): inter-layer dropout (nn.GRU's ``dropout``) with explicit keep masks
    ((B * T, hidden) uint8 per layer boundary, or None).  params = (w_ih, w_hh, b_ih, b_hh) per layer; their gradients
    are accumulated in place (flat-buffer views) like every other native op."""

    @staticmethod
    def forward(ctx, x, meta, *params):
        nl, H, p, masks = meta[:4]
        x = _f32c(x)
        B, T, _ = x.shape
        prm = [q.detach() for q in params]
        train = any(ctx.needs_input_grad) and (bool(meta[4]) if len(meta) > 4 else True)   # meta[4]: caller's grad mode
        dev = x.device
        saved = []
        inp, inp_sp = x.reshape(B * T, -1), None
        out = None
        tc = tensor_core_mode() and _tc_ok(B, 3 * H, H, None, None, True, True)
        for l in range(nl):
            wih, whh, bih, bhh = prm[4 * l:4 * l + 4]
            if inp_sp is not None:
                gi = gemm(None, wih, B * T, 3 * H, inp.shape[1], bias=bih, b_is_weight=True, a_split=inp_sp)
            else:
                gi = linear_fwd(inp, wih, bih)
            out = torch.empty(B, T, H, dtype=torch.float32, device=dev)
            # tensor-core mode: the gate kernel also writes the bf16 split of h_t, read by the next step's GEMM (row stride
            # T * H), by the next layer's input projection and by the weight-gradient GEMMs of backward
            osp = (torch.empty(B, T, H, dtype=torch.bfloat16, device=dev), torch.empty(B, T, H, dtype=torch.bfloat16, device=dev)) if tc else None
            gates = torch.empty(B * T, 3 * H, dtype=torch.float32, device=dev) if train else None
            ghn = torch.empty(B * T, H, dtype=torch.float32, device=dev) if train else None
            gh0 = bhh.expand(B, 3 * H).contiguous()              # h_{-1} = 0: GH = b_hh
            gh = torch.empty(B, 3 * H, dtype=torch.float32, device=dev)
            for t in range(T):
                hp = out[:, t - 1, :] if t > 0 else None
                if t > 0:
                    if tc:
                        gemm(None, whh, B, 3 * H, H, out=gh, bias=bhh, b_is_weight=True, a_split=(osp[0][:, t - 1, :], osp[1][:, t - 1, :]))
                    else:
                        gemm(hp, whh, B, 3 * H, H, out=gh, bias=bhh, b_is_weight=True)
                _native.call('vqcpc_gru_gate_fwd', _p(gi), _p(gh if t > 0 else gh0), _p(hp), T * H, _p(out),
                             _p(osp[0]) if tc else None, _p(osp[1]) if tc else None, _p(gates), _p(ghn), B, T, H, t, _stream())
            layer_in, layer_in_sp = inp, inp_sp
            inp = out.view(B * T, H)
            inp_sp = (osp[0].view(B * T, H), osp[1].view(B * T, H)) if tc else None
            mask = None
            if l < nl - 1 and p > 0 and masks is not None:
                mask = masks[l]
                dropped = torch.empty_like(inp)
                _native.call('vqcpc_mask_scale', _p(inp), _p(mask), 1.0 / (1.0 - p), _p(dropped), inp.numel(), _stream())
                inp, inp_sp = dropped, None
            if train:
                saved.append((layer_in, layer_in_sp, out, osp, gates, ghn, mask))
        ctx.meta = (nl, H, p, B, T, tc)
        ctx.params = params
        ctx.saved = saved
        return out

    @staticmethod
    def backward(ctx, dout):
        nl, H, p, B, T, tc = ctx.meta
        params = ctx.params
        prm = [q.detach() for q in params]
        grads = [_grad_buf(q) if q.requires_grad else torch.zeros_like(q) for q in params]
        dout = _f32c(dout)
        dev = dout.device
        dx = None
        for l in reversed(range(nl)):
            wih, whh, bih, bhh = prm[4 * l:4 * l + 4]
            g_wih, g_whh, g_bih, g_bhh = grads[4 * l:4 * l + 4]
            layer_in, layer_in_sp, out, osp, gates, ghn, mask = ctx.saved[l]
            if mask is not None:                                 # dropout sat on this layer's output
                dm = torch.empty_like(dout)
                _native.call('vqcpc_mask_scale', _p(dout), _p(mask), 1.0 / (1.0 - p), _p(dm), dout.numel(), _stream())
                dout = dm
            dgi = torch.empty(B * T, 3 * H, dtype=torch.float32, device=dev)
            dh = torch.zeros(B, H, dtype=torch.float32, device=dev)
            if tc:
                dgh = None
                dsp = (torch.empty(B, T, 3 * H, dtype=torch.bfloat16, device=dev), torch.empty(B, T, 3 * H, dtype=torch.bfloat16, device=dev))
            else:
                dgh = torch.empty(B * T, 3 * H, dtype=torch.float32, device=dev)
                dgh3 = dgh.view(B, T, 3 * H)
            for t in range(T - 1, -1, -1):
                hp = out[:, t - 1, :] if t > 0 else None
                _native.call('vqcpc_gru_gate_bwd', _p(gates), _p(ghn), _p(hp), T * H, _p(dh), _p(dout), _p(dgi), _p(dgh),
                             _p(dsp[0]) if tc else None, _p(dsp[1]) if tc else None, B, T, H, t, _stream())
                if t > 0:
                    # dh_{t-1} = dh_t z + dGH_t W_hh ;  dW_hh += dGH_t^T h_{t-1}
                    if tc:
                        dh = linear_bwd(None, whh, None, g_whh, None, dx_residual=dh, shape=(B, H),
                                        x_split=(osp[0][:, t - 1, :], osp[1][:, t - 1, :]),
                                        dy_split=(dsp[0][:, t, :], dsp[1][:, t, :]))
                    else:
                        dh = linear_bwd(hp, whh, dgh3[:, t, :], g_whh, None, dx_residual=dh)
            if tc:
                _native.call('vqcpc_colsum_split', _p(dsp[0]), _p(dsp[1]), 3 * H, _p(g_bhh), B * T, 3 * H, _stream())
            else:
                colsum(dgh, g_bhh)
            need_dx = l > 0 or ctx.needs_input_grad[0]
            dx = linear_bwd(layer_in if layer_in_sp is None else None, wih, dgi, g_wih, g_bih, need_dx=need_dx,
                            x_split=layer_in_sp, shape=tuple(layer_in.shape))
            ctx.saved[l] = None
            if l > 0:
                dout = dx.view(B, T, H)
        ctx.saved = None
        dx_in = dx.view(B, T, -1) if ctx.needs_input_grad[0] else None
        return (dx_in, None) + (None,) * len(params)


def gru_forward(gru, x, training):
    """Run a torch.nn.GRU *parameter container* (batch_first, unidirectional) through GRUFn."""
    assert gru.batch_first and not gru.bidirectional and gru.bias
    nl, H, p = gru.num_layers, gru.hidden_size, float(gru.dropout) if training else 0.0
    B, T, _ = x.shape
    masks = None
    if p > 0 and nl > 1:
        masks = [dropout_mask(f'gru{l}', (B * T, H), p, x.device) for l in range(nl - 1)]
    params = []
    for l in range(nl):
        params += [getattr(gru, f'weight_ih_l{l}'), getattr(gru, f'weight_hh_l{l}'), getattr(gru, f'bias_ih_l{l}'),
                   getattr(gru, f'bias_hh_l{l}')]
    return GRUFn.apply(x, (nl, H, p, masks, torch.is_grad_enabled()), *params)


class QuantizeFn(torch.autograd.Function):
    """quantizer/vector_quantizer.py:104-157 — search (or given indices), gather, loss, straight-through output."""

    @staticmethod
    def forward(ctx, z, given_idx, beta, squared, counts, *books):
        shp = z.shape
        D = shp[-1]
        z2 = _f32c(z).view(-1, D)
        n = z2.shape[0]
        ncb = len(books)
        Kc = books[0].shape[0]
        bk = [b.detach() for b in books]
        arr = _ptr_array(bk)
        if given_idx is None:
            idx = torch.empty(n, ncb, dtype=torch.int64, device=z.device)
            _native.call('vqcpc_vq_search', _p(z2), n, D, arr, ncb, Kc, _p(idx), _stream())
        else:
            idx = given_idx.reshape(n, ncb).contiguous()
        zsg = torch.empty_like(z2)
        loss = torch.empty(n, dtype=torch.float32, device=z.device)
        _native.call('vqcpc_vq_gather', _p(z2), n, D, arr, ncb, Kc, _p(idx), float(beta), int(squared), _p(zsg), _p(loss),
                     _p(counts), _stream())
        ctx.save_for_backward(z2, idx, *books)
        ctx.meta = (shp, beta, squared, ncb, Kc)
        idx_out = idx.view(*shp[:-1], ncb)
        ctx.mark_non_differentiable(idx_out)
        return zsg.view(shp), idx_out, loss.view(shp[:-1])

    @staticmethod
    def backward(ctx, g_zsg, _g_idx, g_loss):
        z2, idx, *books = ctx.saved_tensors
        shp, beta, squared, ncb, Kc = ctx.meta
        n, D = z2.shape
        gz = None if g_zsg is None else _f32c(g_zsg).view(n, D)
        gl = None if g_loss is None else _f32c(g_loss).reshape(n)
        dbooks = [torch.zeros_like(b) for b in books]
        dx = torch.empty_like(z2)
        _native.call('vqcpc_vq_bwd', _p(z2), n, D, _ptr_array([b.detach() for b in books]), _ptr_array(dbooks), ncb, Kc,
                     _p(idx), float(beta), int(squared), _p(gz), _p(gl), _p(dx), _stream())
        return (dx.view(shp), None, None, None, None, *dbooks)


class FusedVQUpscaleFn(torch.autograd.Function):
    """Single-codebook quantizer with the (dropout-free) MLP upscaler folded into a per-code table:
    y = T[argmin_k ||z - E_k||^2],  T = upscaler(E) (Kc, zd).  One kernel forward, one backward.

    Gradients: dz = J[idx]^T dy + commitment term (straight-through: vector_quantizer.py:148 feeds z + (q - z).detach()
    to the upscaler, so d y / d z is the upscaler Jacobian at q);  dE from the quantisation loss only (q is detached
    on the upscaler side);  dT[idx] += dy flows on through the autograd graph of T into the upscaler weights."""

    @staticmethod
    def forward(ctx, z, E, T, J, beta, squared, counts):
        shp = z.shape
        D = shp[-1]
        z2 = _f32c(z).view(-1, D)
        n = z2.shape[0]
        Kc, zd = T.shape
        Ed, Td = _f32c(E.detach()), _f32c(T.detach())
        idx = torch.empty(n, dtype=torch.int64, device=z.device)
        loss = torch.empty(n, dtype=torch.float32, device=z.device)
        y = torch.empty(n, zd, dtype=torch.float32, device=z.device)
        _native.call('vqcpc_vq_fused_fwd', _p(z2), n, D, _p(Ed), _p(Td), Kc, zd, float(beta), int(squared), _p(idx), _p(loss),
                     None, _p(y), _p(counts), _stream())
        ctx.save_for_backward(z2, Ed, J, idx)
        ctx.meta = (shp, beta, squared, Kc, zd)
        idx_out = idx.view(*shp[:-1], 1)
        ctx.mark_non_differentiable(idx_out)
        return y.view(*shp[:-1], zd), idx_out, loss.view(shp[:-1])

    @staticmethod
    def backward(ctx, g_y, _g_idx, g_loss):
        z2, Ed, J, idx = ctx.saved_tensors
        shp, beta, squared, Kc, zd = ctx.meta
        n, D = z2.shape
        gy = torch.zeros(n, zd, dtype=torch.float32, device=z2.device) if g_y is None else _f32c(g_y).view(n, zd)
        gl = None if g_loss is None else _f32c(g_loss).reshape(n)
        dx = torch.empty_like(z2)
        dT = torch.zeros(Kc, zd, dtype=torch.float32, device=z2.device)
        dE = torch.zeros_like(Ed)
        _native.call('vqcpc_vq_fused_bwd', _p(z2), n, D, _p(Ed), _p(J), Kc, zd, _p(idx), float(beta), int(squared), _p(gy), _p(gl),
                     _p(dx), _p(dT), _p(dE), _stream())
        return dx.view(shp), dE, dT, None, None, None, None


class InfoNCEFn(torch.autograd.Function):
    """Fused Fks scoring + log-sum-exp.  Returns (sum_{b,k} (lse - pos), pos, neg, correct); only the first is
    differentiable.  vqcpc_helper.py:79-98, :5-29; vqcpc_encoder_trainer.py:240-272."""

    @staticmethod
    def forward(ctx, c, W, y_pos, y_neg):
        c, W, y_pos, y_neg = _f32c(c), _f32c(W), _f32c(y_pos), _f32c(y_neg)
        B, cd = c.shape
        zd, cd2, KR = W.shape
        N = y_neg.shape[1]
        assert cd2 == cd and y_pos.shape == (B, KR, zd) and y_neg.shape == (B, N, KR, zd)
        dev = c.device
        u = torch.empty(B, KR, zd, dtype=torch.float32, device=dev)
        pos = torch.empty(B, KR, dtype=torch.float32, device=dev)
        neg = torch.empty(B, KR, N, dtype=torch.float32, device=dev)
        lse = torch.empty(B, KR, dtype=torch.float32, device=dev)
        correct = torch.empty(B, KR, dtype=torch.int32, device=dev)
        nll = torch.empty(B, KR, dtype=torch.float32, device=dev)
        _native.call('vqcpc_infonce_fwd', _p(c), _p(W), _p(y_pos), _p(y_neg), B, N, KR, zd, cd, _p(u), _p(pos), _p(neg), _p(lse),
                     _p(nll), _p(correct), _stream())
        total = reduce_sum(nll).reshape(())   # deterministic fixed-order sum
        ctx.save_for_backward(c, W, y_pos, y_neg, u, pos, neg, lse)
        ctx.mark_non_differentiable(pos, neg, correct)
        return total, pos, neg, correct

    @staticmethod
    def backward(ctx, g_total, *_):
        c, W, y_pos, y_neg, u, pos, neg, lse = ctx.saved_tensors
        B, cd = c.shape
        zd, _, KR = W.shape
        N = y_neg.shape[1]
        dc = torch.zeros_like(c)
        dW = torch.zeros_like(W)
        dyp = torch.empty_like(y_pos)
        dyn = torch.empty_like(y_neg)
        g = _f32c(g_total).reshape(1)         # upstream gradient stays on the device (no host sync)
        _native.call('vqcpc_infonce_bwd', _p(c), _p(W), _p(y_pos), _p(y_neg), _p(u), _p(pos), _p(neg), _p(lse), _p(g), 1.0, B, N, KR,
                     zd, cd, _p(dc), _p(dW), _p(dyp), _p(dyn), _stream())
        return dc, dW, dyp, dyn


# ------------------------------------------------------------------------------------------------ index-gather path
def idxpath_supported(N, KR, zd, cd, Kc, D):
    """True when csrc/idxpath.cu handles these shapes (zd == cd == 32, N <= 15, KR <= 16, Kc <= 698, D in {3, 4})."""
    return bool(_native.load().vqcpc_idxpath_supported(int(N), int(KR), int(zd), int(cd), int(Kc), int(D)))


class BottleneckLink:
    """Side channel between the two autograd nodes of the index path.

    In table mode the upscaled vectors of the right / negative blocks are never materialised, so the gradient that flows
    from the InfoNCE node back into the quantizer node is not a tensor of rows either: it is the per-candidate
    coefficients d loss / d score plus the u vectors (``segments``), which ``VQIndexFn.backward`` turns into d z with the
    upscaler Jacobian J[idx].  ``token`` is a one-element differentiable output of the quantizer node that the InfoNCE
    node takes as an input, which makes autograd run the InfoNCE backward first."""

    def __init__(self):
        self.table = None        # T = upscaler(E) with autograd history to the MLP weights
        self.token = None
        self.idx = None          # (P,) int64 codes of every vector of the pass
        self.segments = []       # filled by InfoNCEIdxFn.backward: (row0, rows, coef, u, per_b, KR)


class VQIndexFn(torch.autograd.Function):
    """Quantizer node of the index path: exact search over all P vectors, loss, and the upscaled rows y = T[idx] of the
    LEFT blocks only (rows [y_row0, y_row0 + y_rows): the context GRU reads them).  vector_quantizer.py:104-157 +
    mlp_upscaler.py:21-35.  Backward: d z = coef * J[idx]^T v + commitment term for every vector, d E, and d T from the
    left rows (the right / negative part of d T comes from InfoNCEIdxFn)."""

    @staticmethod
    def forward(ctx, z, E, T, J, beta, squared, counts, y_row0, y_rows, link):
        shp = z.shape
        D = shp[-1]
        z2 = _f32c(z).view(-1, D)
        n = z2.shape[0]
        Kc, zd = T.shape
        Ed, Td = _f32c(E.detach()), _f32c(T.detach())
        idx = torch.empty(n, dtype=torch.int64, device=z.device)
        loss = torch.empty(n, dtype=torch.float32, device=z.device)
        y = torch.empty(y_rows, zd, dtype=torch.float32, device=z.device)
        _native.call('vqcpc_vq_fused_fwd_rows', _p(z2), n, D, _p(Ed), _p(Td), Kc, zd, float(beta), int(squared), _p(idx), _p(loss),
                     None, _p(y), int(y_row0), int(y_rows), _p(counts), _stream())
        token = torch.zeros(1, dtype=torch.float32, device=z.device)
        ctx.save_for_backward(z2, Ed, J, idx)
        ctx.meta = (shp, beta, squared, Kc, zd, y_row0, y_rows)
        ctx.link = link
        link.idx = idx
        ctx.mark_non_differentiable(idx)
        return y, idx, loss, token

    @staticmethod
    def backward(ctx, g_y, _g_idx, g_loss, _g_token):
        z2, Ed, J, idx = ctx.saved_tensors
        shp, beta, squared, Kc, zd, y_row0, y_rows = ctx.meta
        n, D = z2.shape
        dev = z2.device
        gl = None if g_loss is None else _f32c(g_loss).reshape(n)
        dT = torch.zeros(Kc, zd, dtype=torch.float32, device=dev)
        dE = torch.zeros_like(Ed)
        segs = list(ctx.link.segments)
        ctx.link.segments = []
        keep = []
        if g_y is not None and y_rows > 0:
            gy = _f32c(g_y).view(y_rows, zd)
            if gy.data_ptr() % 16:
                gy = gy.clone()
            keep.append(gy)
            segs.append((y_row0, y_rows, gy, None, 1, 1))      # (the kernel also scatters these rows into dT)
        KR = max([s[5] for s in segs if s[3] is not None], default=1)
        ns = len(segs)
        row0 = (ctypes.c_int64 * max(ns, 1))(*[s[0] for s in segs])
        rows = (ctypes.c_int64 * max(ns, 1))(*[s[1] for s in segs])
        kind = (ctypes.c_int * max(ns, 1))(*[1 if s[3] is None else 2 for s in segs])
        pa = (ctypes.c_void_p * max(ns, 1))(*[s[2].data_ptr() for s in segs])
        pb = (ctypes.c_void_p * max(ns, 1))(*[s[3].data_ptr() if s[3] is not None else 0 for s in segs])
        per_b = (ctypes.c_int * max(ns, 1))(*[s[4] for s in segs])
        dx = torch.empty_like(z2)
        _native.call('vqcpc_vq_idx_bwd', _p(z2), n, D, _p(Ed), _p(J), Kc, zd, _p(idx), float(beta), int(squared), _p(gl), ns, row0,
                     rows, kind, pa, pb, per_b, int(KR), _p(dx), _p(dE), _p(dT), _stream())
        return dx.view(shp), dE, dT, None, None, None, None, None, None, None


class InfoNCEIdxFn(torch.autograd.Function):
    """Fks scoring + log-sum-exp from code indices (vqcpc_helper.py:79-98, :5-29; vqcpc_encoder_trainer.py:240-272).
    ``seg_pos`` / ``seg_neg`` = first row of the right blocks (B, KR) / the negatives (B, N, KR) in the pass's flat vector
    list: the quantizer node needs them to route the score gradients back to z."""

    @staticmethod
    def forward(ctx, c, W, T, token, idx_pos, idx_neg, link, seg_pos, seg_neg):
        c, W, Td = _f32c(c), _f32c(W), _f32c(T.detach())
        B, cd = c.shape
        zd, cd2, KR = W.shape
        N = idx_neg.shape[1]
        Kc = Td.shape[0]
        assert cd2 == cd and idx_pos.shape == (B, KR) and idx_neg.shape == (B, N, KR)
        assert idx_pos.is_contiguous() and idx_neg.is_contiguous()
        dev = c.device
        u = torch.empty(B, KR, zd, dtype=torch.float32, device=dev)
        pos = torch.empty(B, KR, dtype=torch.float32, device=dev)
        neg = torch.empty(B, KR, N, dtype=torch.float32, device=dev)
        lse = torch.empty(B, KR, dtype=torch.float32, device=dev)
        correct = torch.empty(B, KR, dtype=torch.int32, device=dev)
        nll = torch.empty(B, KR, dtype=torch.float32, device=dev)
        _native.call('vqcpc_infonce_idx_fwd', _p(c), _p(W), _p(Td), _p(idx_pos), _p(idx_neg), B, N, KR, zd, cd, Kc, _p(u), _p(pos),
                     _p(neg), _p(lse), _p(nll), _p(correct), _stream())
        total = reduce_sum(nll).reshape(())
        ctx.save_for_backward(c, W, Td, idx_pos, idx_neg, u, pos, neg, lse)
        ctx.link, ctx.segs = link, (seg_pos, seg_neg)
        ctx.mark_non_differentiable(pos, neg, correct)
        return total, pos, neg, correct

    @staticmethod
    def backward(ctx, g_total, *_):
        c, W, Td, idx_pos, idx_neg, u, pos, neg, lse = ctx.saved_tensors
        B, cd = c.shape
        zd, _, KR = W.shape
        N = idx_neg.shape[1]
        Kc = Td.shape[0]
        dev = c.device
        dc = torch.empty_like(c)
        dW = torch.zeros_like(W)
        dT = torch.zeros_like(Td)
        coef_pos = torch.empty(B, KR, dtype=torch.float32, device=dev)
        coef_neg = torch.empty(B, N, KR, dtype=torch.float32, device=dev)
        part = torch.empty(_native.load().vqcpc_infonce_idx_scratch(B, KR, Kc), dtype=torch.float32, device=dev)
        g = _f32c(g_total).reshape(1)
        _native.call('vqcpc_infonce_idx_bwd', _p(c), _p(W), _p(Td), _p(idx_pos), _p(idx_neg), _p(u), _p(pos), _p(neg), _p(lse), _p(g),
                     1.0, B, N, KR, zd, cd, Kc, _p(dc), _p(coef_pos), _p(coef_neg), _p(part), _p(dW), _p(dT), _stream())
        seg_pos, seg_neg = ctx.segs
        ctx.link.segments += [(seg_neg, B * N * KR, coef_neg, u, N * KR, KR), (seg_pos, B * KR, coef_pos, u, KR, KR)]
        return dc, dW, dT, torch.zeros(1, dtype=torch.float32, device=dev), None, None, None, None, None
