"""CPU tests of the drop-in boundary and host logic (no kernels are launched here)."""
import ctypes
import os
import re
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, 'include', 'vqcpc_b200.h')


def _header_functions():
    src = open(HEADER).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    out = {}
    for m in re.finditer(r'\b(?:int64_t|int|const char\*)\s+(vqcpc_\w+)\s*\(([^;]*?)\)\s*;', src, flags=re.S):
        args = m.group(2).strip()
        n = 0 if args in ('', 'void') else len([a for a in args.split(',') if a.strip()])
        out[m.group(1)] = n
    return out


def test_library_exports_every_declared_symbol():
    from VQCPCB import _native
    lib = _native.load()
    fns = _header_functions()
    assert len(fns) >= 25
    for name in fns:
        assert hasattr(lib, name), f'{name} declared in include/vqcpc_b200.h but not exported'
    assert lib.vqcpc_abi_version() == 1


def test_ctypes_signatures_match_header_arity():
    from VQCPCB import _native
    fns = _header_functions()
    for name, args in _native.SIGNATURES.items():
        assert name in fns, f'{name} bound in _native.py but missing from the header'
        assert len(args) == fns[name], f'{name}: ctypes binds {len(args)} args, header declares {fns[name]}'
    bound = set(_native.SIGNATURES) | set(_native._SPECIAL)
    assert set(fns) == bound


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, 'vqcpc-bach_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh')):
                txt = open(os.path.join(dirpath, f)).read()
                assert 'oracle' not in txt.replace('no oracle', ''), f'{f} references the oracle'
    assert 'oracle' not in open(os.path.join(ROOT, 'main_encoder.py')).read()


def test_no_cpu_fallback_ops_raise_on_cpu_tensors():
    from VQCPCB import ops
    a = torch.zeros(4, 4)
    with pytest.raises(RuntimeError):
        ops.gemm(a, a, 4, 4, 4)


def test_state_dict_keys_match_reference_checkpoints():
    """Strict load of the oracle's parameter dict, whose keys were pinned by a strict load into the live reference."""
    from oracle import params as P, vqcpc_oracle as O
    from tests.helpers import build_product_trainer, load_params_into, product_named_params
    for over in (dict(), dict(codebook_dim=32, upscaler=False, KL=2, KR=2),
                 dict(downscaler='transformer_linear'),
                 dict(downscaler='lstm', downscale_factors=[16], lstm_hidden=64, lstm_bidirectional=True),
                 dict(downscaler='lstm', downscale_factors=[16], lstm_hidden=64, lstm_layers=1, lstm_bidirectional=False)):
        cfg = O.default_cfg(**over)
        tr = build_product_trainer(cfg, device='cpu')
        prm = P.make_params(cfg, 0)
        load_params_into(tr, prm)
        named = product_named_params(tr)
        assert set(named) == set(prm)
        for k, p in named.items():
            assert tuple(p.shape) == tuple(prm[k].shape), k


def test_getters_and_config_build_the_model():
    import importlib.util
    from VQCPCB.getters import get_dataloader_generator, get_encoder, get_encoder_trainer
    path = os.path.join(ROOT, 'vqcpc-bach_b200', 'VQCPCB', 'configs', 'encoder_vqcpc.py')
    spec = importlib.util.spec_from_file_location('cfgmod', path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    config = mod.config
    for key in ('training_method', 'dataset', 'dataloader_generator_kwargs', 'data_processor_type', 'downscaler_type',
                'downscaler_kwargs', 'quantizer_type', 'quantizer_kwargs', 'upscaler_type', 'upscaler_kwargs',
                'auxiliary_networks_kwargs', 'lr', 'schedule_lr', 'batch_size', 'num_batches', 'num_epochs',
                'quantizer_regularization', 'timestamp', 'savename'):
        assert key in config
    config['quantizer_kwargs']['initialize'] = True
    gen = get_dataloader_generator(config['dataset'], config['training_method'], config['dataloader_generator_kwargs'])
    enc = get_encoder('/tmp/x', gen, config)
    tr = get_encoder_trainer('/tmp/x', gen, config['training_method'], enc, config['auxiliary_networks_kwargs'])
    n = sum(p.numel() for m in tr._modules_in_order() for p in m.parameters())
    assert 15_000_000 < n < 15_400_000       # 15.1 M parameters
    train, val, test = gen.dataloaders(batch_size=3)
    b = next(iter(train))
    assert b['x_left'].shape == (3, 32, 4) and b['negative_samples'].shape == (3, 15, 8, 4, 4)
    assert b['x_left'].dtype == torch.int64 and not b['x_left'].is_cuda


def test_lstm_config_builds_the_reference_shaped_model():
    """configs/encoder_vqcpc_lstm.py (the shape of the reference's encoder_random_config.py: GRU block downscaler,
    dropout 0.1) goes through the same factories."""
    import importlib.util
    from VQCPCB.getters import get_dataloader_generator, get_encoder, get_encoder_trainer
    from VQCPCB.downscalers.lstm_downscaler import LstmDownscaler
    path = os.path.join(ROOT, 'vqcpc-bach_b200', 'VQCPCB', 'configs', 'encoder_vqcpc_lstm.py')
    spec = importlib.util.spec_from_file_location('cfgmod_lstm', path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    config = mod.config
    gen = get_dataloader_generator(config['dataset'], config['training_method'], config['dataloader_generator_kwargs'])
    enc = get_encoder('/tmp/x', gen, config)
    tr = get_encoder_trainer('/tmp/x', gen, config['training_method'], enc, config['auxiliary_networks_kwargs'])
    assert isinstance(enc.downscaler, LstmDownscaler) and enc.downscaler.g_enc_bwd is not None
    assert enc.downscaler.downscale_factors == [16] and enc.upscaler.dropout == 0.1
    keys = set(enc.downscaler.state_dict())
    assert {'g_enc_fwd.weight_ih_l0', 'g_enc_bwd.weight_hh_l1', 'output_linear.weight'} <= keys
    assert tuple(enc.downscaler.output_linear.weight.shape) == (3, 1024)


def test_preprocess_layout_matches_oracle():
    from oracle import vqcpc_oracle as O
    x = torch.arange(2 * 8 * 4).view(2, 8, 4)
    tok = O.preprocess(x, 16)
    assert tok.shape == (2, 2, 16)
    assert tok[0, 0].tolist() == list(range(16))         # tick-major, voice fastest
    # product preprocess is the same reshape (device copy aside)
    assert torch.equal(x.contiguous().view(2, 2, 16), tok)


def test_flat_params_alias_and_zero_grad():
    from VQCPCB.optim import FlatParams
    ps = [torch.nn.Parameter(torch.randn(3, 5)), torch.nn.Parameter(torch.randn(7)), torch.nn.Parameter(torch.randn(2, 2))]
    ref = [p.detach().clone() for p in ps]
    flat = FlatParams(ps, _allow_cpu=True)
    flat.attach_grads(tail=6)
    assert flat.numel == 16 + 8 + 4 and flat.grad_full.numel() == flat.numel + 6
    for p, r, o in zip(ps, ref, flat.offsets):
        assert torch.equal(p.detach(), r)
        assert p.data_ptr() == flat.data.data_ptr() + 4 * o and o % 4 == 0
        assert p.grad.data_ptr() == flat.grad.data_ptr() + 4 * o
    (ps[0].sum() * 2 + ps[1].sum()).backward()
    assert flat.grad[:15].eq(2).all() and flat.grad[16:23].eq(1).all()      # autograd accumulated IN PLACE
    flat.grad_full.zero_()
    assert ps[0].grad.abs().sum() == 0


def test_lr_schedule_matches_oracle():
    from oracle import vqcpc_oracle as O
    lo, hi, warm = 0.1, 1.0, 10000
    s1 = (hi - lo) / warm
    s2 = -s1 * 0.1
    f = lambda step: max(min(lo + s1 * step, hi + (step - warm) * s2), lo)   # noqa: E731  (trainer's lambda)
    for s in (0, 1, 5000, 10000, 20000, 200000):
        assert abs(f(s) - O.lr_lambda(s)) < 1e-12


_DIST_WORKER = r'''
import os, sys, torch
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.join(sys.argv[1], "vqcpc-bach_b200"))
import torch.distributed as dist
from VQCPCB import distributed as vq_dist
from VQCPCB.optim import FlatParams, FusedClipAdam
rank, world, local = vq_dist.init_from_env(backend="gloo")
assert world == 2 and vq_dist.world_size() == 2 and vq_dist.rank() == rank
torch.manual_seed(0)
ps = [torch.nn.Parameter(torch.randn(4, 3)), torch.nn.Parameter(torch.randn(5))]
if rank == 1:
    for p in ps: p.data.add_(1.0)
vq_dist.broadcast_parameters(ps)                       # replicas agree after broadcast
chk = torch.cat([p.detach().reshape(-1) for p in ps]).sum()
both = [torch.zeros(()) for _ in range(2)]
dist.all_gather(both, chk)
assert torch.equal(both[0], both[1])
flat = FlatParams(ps, _allow_cpu=True)
opt = FusedClipAdam.__new__(FusedClipAdam); opt.flat = flat
flat.attach_grads(tail=16)
flat.grad.fill_(float(rank + 1))                        # rank0: 1, rank1: 2
sums = torch.tensor([1.0, 2.0, 3.0]) * (rank + 1); correct = torch.tensor([4.0, 8.0]) * (rank + 1)
# counts beyond 2^24 (not representable in one fp32) must still sum exactly: they travel as two 12-bit-split halves
big = 20_000_003
counts = torch.tensor([[1, 0, 2, big], [7, 4095, 4096, 0]], dtype=torch.int32) * (rank + 1)
assert FusedClipAdam.tail_elems(5, counts.numel()) <= 21
flat.attach_grads(tail=21)
flat.grad.fill_(float(rank + 1))
scale = FusedClipAdam.all_reduce(opt, stats=(sums, correct), counts=[counts])
assert scale == 0.5
assert flat.grad.eq(3.0).all()                          # summed gradients; update applies scale 1/world
assert torch.allclose(sums, torch.tensor([1.5, 3.0, 4.5])) and torch.allclose(correct, torch.tensor([6.0, 12.0]))
assert counts.tolist() == [[3, 0, 6, 3 * big], [21, 3 * 4095, 3 * 4096, 0]], counts.tolist()
# data-dependent codebook initialisation (quirk Q5) under data parallelism: every rank draws from ITS OWN first batch,
# rank 0's draw must win on all replicas (reference: quantizer/vector_quantizer.py:57-70 has no collective at all)
from VQCPCB.quantizer.vector_quantizer import ProductVectorQuantizer
torch.manual_seed(100 + rank)
vq = ProductVectorQuantizer(8, 4, 0.25, 2, False, True, True)
data = torch.randn(64, 4)
vq._initialize(data)
assert vq.initialize is False
cb = torch.cat([e.detach().reshape(-1) for e in vq.embeddings])
got = [torch.zeros_like(cb) for _ in range(2)]
dist.all_gather(got, cb)
assert torch.equal(got[0], got[1]), "codebooks differ across ranks after the data-dependent init"
rows0 = [torch.zeros_like(data) for _ in range(2)]
dist.all_gather(rows0, data)
for k, e in enumerate(vq.embeddings):                   # every code is a row of RANK 0's batch
    src = rows0[0][:, 2 * k:2 * k + 2]
    assert all((src == r).all(1).any() for r in e.detach())
vq2 = ProductVectorQuantizer(8, 4, 0.25, 1, False, True, True)
vq2.sync_init = False                                   # opt-out: single-process use inside a process group
vq2._initialize(data)
cb2 = vq2.embeddings[0].detach().reshape(-1)
got2 = [torch.zeros_like(cb2) for _ in range(2)]
dist.all_gather(got2, cb2)
assert not torch.equal(got2[0], got2[1])
assert vq_dist.shard_seed(1234) == 1234 + rank
vq_dist.barrier()
sys.stdout.write(f"rank{rank}ok\n"); sys.stdout.flush()
'''


def test_data_parallel_allreduce_gloo_world2(tmp_path):
    """world_size-2 gloo run of the DP plumbing: parameter broadcast + the single flat all-reduce with its tail."""
    script = tmp_path / 'worker.py'
    script.write_text(_DIST_WORKER)
    port = 29500 + (os.getpid() % 500)
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node=2', '--master-addr', '127.0.0.1',
           '--master-port', str(port), str(script), ROOT]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    assert 'rank0ok' in r.stdout and 'rank1ok' in r.stdout
