"""Flat parameter / gradient storage and the fused clip + Adam optimizer.

Replaces ``nn.utils.clip_grad_norm_(params, 5)`` + ``torch.optim.Adam.step()`` (VQCPCB/vqcpc_encoder_trainer.py:92,
:313-314) by two kernel launches over one contiguous fp32 buffer: a deterministic fp64 two-stage sum of squares, then
the clipped Adam update reading the norm from device memory (no host sync).  Under data parallelism the same flat
gradient buffer (plus a small tail carrying code-usage counts and the step's loss/accuracy sums) is what the single
NCCL all-reduce per step operates on.
"""
import ctypes

import torch

from . import _native
from . import distributed as vq_dist


def _p(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


class FlatParams:
    """Re-homes every parameter (and its .grad) as a view into one flat fp32 buffer (16-byte aligned slots)."""

    def __init__(self, params, _allow_cpu=False):
        seen, uniq = set(), []
        for p in params:
            if id(p) not in seen:
                seen.add(id(p))
                uniq.append(p)
        self.params = uniq
        dev = uniq[0].device
        if dev.type != 'cuda' and not _allow_cpu:     # (_allow_cpu: host-logic unit tests only)
            raise RuntimeError('FlatParams: parameters must be on a CUDA device (call trainer.to("cuda") first)')
        offs, total = [], 0
        for p in uniq:
            offs.append(total)
            total += (p.numel() + 3) // 4 * 4
        self.offsets, self.numel = offs, total
        self.data = torch.zeros(total, dtype=torch.float32, device=dev)
        for p, o in zip(uniq, offs):
            view = self.data[o:o + p.numel()].view(p.shape)
            view.copy_(p.data)
            p.data = view
        self.grad_full = None
        self.grad = None

    def attach_grads(self, tail):
        self.grad_full = torch.zeros(self.numel + tail, dtype=torch.float32, device=self.data.device)
        self.grad = self.grad_full[:self.numel]
        self.tail = self.grad_full[self.numel:]
        for p, o in zip(self.params, self.offsets):
            p.grad = self.grad[o:o + p.numel()].view(p.shape)

    def named_slices(self, named):
        """{name: flat slice} for tests / checkpoints of the optimizer state."""
        idx = {id(p): (o, p.numel()) for p, o in zip(self.params, self.offsets)}
        return {n: idx[id(p)] for n, p in named if id(p) in idx}


class FusedClipAdam(torch.optim.Optimizer):
    """torch.optim.Optimizer subclass (so LambdaLR and ``param_groups[0]['lr']`` work as in the reference) whose
    ``step`` is the fused global-norm clip + Adam kernel pair on the flat buffers."""

    def __init__(self, flat, lr=1e-4, betas=(0.9, 0.999), eps=1e-8, extra_tail=0):
        self.flat = flat
        super().__init__(flat.params, dict(lr=lr, betas=betas, eps=eps))
        flat.attach_grads(extra_tail)
        dev = flat.data.device
        self.exp_avg = torch.zeros_like(flat.data)
        self.exp_avg_sq = torch.zeros_like(flat.data)
        self.gnorm_sq = torch.zeros(1, dtype=torch.float32, device=dev)
        self.step_count = 0
        self.distributed = True          # False: never all-reduce (single-process reference inside a multi-rank job)

    def zero_grad(self, set_to_none=False):
        # the .grad views must stay alive (kernels accumulate into them): always zero in place
        self.flat.grad_full.zero_()

    @staticmethod
    def tail_elems(stats_elems, count_elems):
        """fp32 slots the all-reduce tail needs: statistics as they are, every integer count as two halves."""
        return stats_elems + 2 * count_elems

    @torch.no_grad()
    def all_reduce(self, stats=None, counts=None, sums=None):
        """The ONE collective of a step: sum all-reduce of [gradients | stats | counts]; returns the gradient scale
        (1 / world_size) the update must apply.  ``stats`` (float tensors) are replaced by their cross-rank mean;
        ``counts`` (integer tensors, e.g. the code-usage histograms) by their cross-rank SUM.  Counts travel through the
        fp32 buffer as two halves (low 12 bits, rest), so the sums stay exact for any realistic count (each half < 2^24
        after the reduction) although the collective is a float one.  No-op (scale 1) in a single process."""
        flat = self.flat
        world = vq_dist.world_size() if getattr(self, 'distributed', True) else 1
        if world <= 1:
            return 1.0
        tail = flat.tail
        stats = list(stats or [])
        counts = list(counts or [])
        sums = [t for t in (sums or []) if t is not None]           # float tensors replaced by their cross-rank SUM
        o = 0
        for t in sums:
            assert o + t.numel() <= tail.numel(), 'all-reduce tail too small'
            tail[o:o + t.numel()].copy_(t.reshape(-1))
            o += t.numel()
        for t in stats:
            assert o + t.numel() <= tail.numel(), 'all-reduce tail too small'
            tail[o:o + t.numel()].copy_(t.reshape(-1))
            o += t.numel()
        for t in counts:
            n = t.numel()
            assert o + 2 * n <= tail.numel(), 'all-reduce tail too small'
            v = t.reshape(-1).long()
            tail[o:o + n].copy_(v & 0xFFF)
            tail[o + n:o + 2 * n].copy_(v >> 12)
            o += 2 * n
        vq_dist.all_reduce_sum(flat.grad_full)
        o = 0
        for t in sums:
            t.copy_(tail[o:o + t.numel()].view(t.shape))
            o += t.numel()
        for t in stats:
            t.copy_(tail[o:o + t.numel()].view(t.shape) / world)
            o += t.numel()
        for t in counts:
            n = t.numel()
            tot = tail[o:o + n].round().long() + (tail[o + n:o + 2 * n].round().long() << 12)
            t.copy_(tot.view(t.shape))
            o += 2 * n
        return 1.0 / world

    @torch.no_grad()
    def step(self, closure=None, max_norm=0.0, stats=None, counts=None, sums=None):
        """Clip by global norm ``max_norm`` (0 = no clip) and apply Adam.  ``stats`` = (sums, correct) device tensors and
        ``counts`` = integer code-usage histograms ride in the tail of the all-reduce buffer (stats averaged, counts
        summed across ranks, both in place)."""
        flat = self.flat
        grad_scale = self.all_reduce(stats, counts, sums)
        g = self.param_groups[0]
        self.step_count += 1
        _native.call('vqcpc_grad_sumsq', _p(flat.grad), flat.numel, grad_scale, _p(self.gnorm_sq), _stream())
        _native.call('vqcpc_clip_adam', _p(flat.data), _p(flat.grad), _p(self.exp_avg), _p(self.exp_avg_sq), flat.numel,
                     _p(self.gnorm_sq), float(max_norm), float(grad_scale), float(g['lr']), float(g['betas'][0]),
                     float(g['betas'][1]), float(g['eps']), self.step_count, _stream())
        return None
