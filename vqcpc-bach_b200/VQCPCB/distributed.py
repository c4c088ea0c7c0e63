"""Data-parallel plumbing: one process per GPU, torch.distributed (NCCL over NVLink 5 / NVSwitch on the B200 box,
gloo in the CPU tests).  The encoder-training path shards by samples: every rank draws its own batch, runs
forward/backward locally and the step needs exactly one collective — a sum all-reduce of the flat buffer
[gradients | code-usage counts | loss & accuracy sums].  The reference has no distributed code at all
(only two TODOs: VQCPCB/encoder.py:40, student_encoder_trainer.py:21)."""
import os

import torch
import torch.distributed as dist


def is_initialized():
    return dist.is_available() and dist.is_initialized()


def world_size():
    return dist.get_world_size() if is_initialized() else 1


def rank():
    return dist.get_rank() if is_initialized() else 0


def init_from_env(backend=None):
    """Initialise from torchrun's env (RANK / WORLD_SIZE / LOCAL_RANK / MASTER_ADDR / MASTER_PORT).  No-op when
    WORLD_SIZE is unset or 1.  Returns (rank, world_size, local_rank)."""
    ws = int(os.environ.get('WORLD_SIZE', '1'))
    if ws <= 1:
        return 0, 1, 0
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if backend is None:
        backend = 'nccl' if torch.cuda.is_available() else 'gloo'
    if backend == 'nccl':
        torch.cuda.set_device(local)
    if not is_initialized():
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        kwargs = {}
        if backend == 'nccl':
            kwargs['device_id'] = torch.device('cuda', local)
        dist.init_process_group(backend=backend, **kwargs)
    return dist.get_rank(), dist.get_world_size(), local


# bench.py sets this to a list: every all-reduce is then bracketed by CUDA events on the current stream
COLLECTIVE_EVENTS = None


def all_reduce_sum(flat_tensor):
    if world_size() > 1:
        ev = COLLECTIVE_EVENTS
        if ev is not None and flat_tensor.is_cuda:
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            dist.all_reduce(flat_tensor, op=dist.ReduceOp.SUM)
            b.record()
            ev.append((a, b))
        else:
            dist.all_reduce(flat_tensor, op=dist.ReduceOp.SUM)


def broadcast_parameters(params, src=0):
    """Rank-0 parameters to everyone (needed after the data-dependent codebook init, quirk Q5)."""
    if world_size() > 1:
        for p in params:
            dist.broadcast(p.data, src=src)


def barrier():
    if world_size() > 1:
        dist.barrier()


def shard_seed(base_seed):
    """Per-rank data seed: ranks draw disjoint synthetic windows (weak scaling: fixed per-GPU batch)."""
    return base_seed + rank()
