"""Multi-GPU correctness check of the data-parallel step (run under torchrun on >= 2 GPUs).

Every rank trains on its own batch for two steps with the single NCCL all-reduce; then
  (a) all ranks must hold bit-identical parameters, and
  (b) rank 0 re-runs the same two steps in one process on the CONCATENATED global batch (no collective): the
      mean of per-rank mean losses equals the global batch mean, so parameters must agree to fp32 round-off.
  (c) a FRESH run (``initialize=True``: data-dependent codebook initialisation from each rank's own first batch, quirk Q5)
      must still leave every rank with bit-identical parameters after 3 steps (rank 0's draw is broadcast), with the
      globally summed code-usage histogram identical on all ranks and accounting for every encoded vector.
Prints one line per check and exits non-zero on failure."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'vqcpc-bach_b200')):
    sys.path.insert(0, p)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
from VQCPCB import distributed as vq_dist, ops  # noqa: E402
from oracle import params as P, vqcpc_oracle as O  # noqa: E402  (test infrastructure: parameters / batches only)
from tests.helpers import build_product_trainer, load_params_into, product_named_params  # noqa: E402


def main():
    rank, world, local = vq_dist.init_from_env()
    assert world >= 2, 'run under torchrun with >= 2 ranks'
    torch.cuda.set_device(local)
    ops.set_gemm_mode(os.environ.get('VQCPC_MODE', 'fp32'))
    cfg = O.default_cfg(codebook_size=64, KL=2, KR=2)
    prm = P.make_params(cfg, 0)
    Bper = 4
    batches = [[P.make_batch(cfg, Bper, seed=100 * step + r) for r in range(world)] for step in range(2)]

    tr = build_product_trainer(cfg)
    load_params_into(tr, prm)
    tr.init_optimizers(1e-3, False)
    tr.train()
    tr.encoder.quantizer.track_usage(True)
    stats = [tr.step(batches[s][rank], True, False) for s in range(2)]
    torch.cuda.synchronize()
    flat = tr.optimizer.flat.data
    chk = torch.stack([flat.double().sum(), (flat.double() ** 2).sum()])
    gathered = [torch.zeros_like(chk) for _ in range(world)]
    dist.all_gather(gathered, chk)
    same = all(torch.equal(g, gathered[0]) for g in gathered)
    ok = same
    if rank == 0:
        print(f'[ddp_check] world={world} identical parameters across ranks: {same}')
        ref = build_product_trainer(cfg)
        load_params_into(ref, prm)
        ref.distributed = False                # no collectives at all in the single-process reference
        ref.init_optimizers(1e-3, False)
        ref.train()
        ref_stats = []
        for s in range(2):
            gb = {k: torch.cat([batches[s][r][k] for r in range(world)], 0) for k in batches[s][0]}
            ref_stats.append(ref.step(gb, True, False))
        torch.cuda.synchronize()
        a, b = tr.optimizer.flat.data, ref.optimizer.flat.data
        rel = ((a - b).double().norm() / b.double().norm()).item()
        moved = (b.double() - _flat_of(prm, ref).double()).norm().item()
        drift = (a - b).double().norm().item() / max(moved, 1e-30)
        l_ddp = float(stats[1]['loss'])            # averaged over ranks by the all-reduce tail
        l_ref = float(ref_stats[1]['loss'])
        print(f'[ddp_check] params vs single-process global batch: rel {rel:.2e}, |diff| / |update| {drift:.2e}; '
              f'loss ddp {l_ddp:.6f} ref {l_ref:.6f}')
        ok = ok and drift < 2e-2 and abs(l_ddp - l_ref) < 1e-4 * abs(l_ref)
    ok_c = check_fresh_run(cfg, prm, rank, world)
    ok = ok and ok_c
    flag = torch.tensor([1 if ok else 0], device='cuda')
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    vq_dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


def _identical(t):
    """True iff the byte content of ``t`` is the same on every rank (fp64 sum / sum of squares + an integer hash)."""
    flat = t.reshape(-1)
    if flat.dtype.is_floating_point:
        bits = flat.float().view(torch.int32).long()
    else:
        bits = flat.long()
    w = torch.arange(1, bits.numel() + 1, device=bits.device, dtype=torch.long)
    chk = torch.stack([(bits * w).sum(), (bits ^ (w * 2654435761)).sum()])
    got = [torch.zeros_like(chk) for _ in range(dist.get_world_size())]
    dist.all_gather(got, chk)
    return all(torch.equal(g, got[0]) for g in got)


def check_fresh_run(cfg, prm, rank, world):
    """(c): initialize=True, rank-local data, 3 steps."""
    tr = build_product_trainer(cfg)
    load_params_into(tr, prm)
    q = tr.encoder.quantizer
    q.initialize = True
    torch.manual_seed(7 + rank)                      # ranks draw DIFFERENT randperms from DIFFERENT data
    tr.init_optimizers(1e-3, False)
    tr.train()
    Bper, steps = 8, 3
    cb0 = q.embeddings[0].detach().clone()
    outs = [tr.step(P.make_batch(cfg, Bper, seed=900 + 10 * s + rank), True, False) for s in range(steps)]
    torch.cuda.synchronize()
    changed = not torch.equal(cb0, q.embeddings[0].detach())
    same_params = _identical(tr.optimizer.flat.data)
    same_usage = _identical(q.global_usage)
    vec_per_step = Bper * (cfg.KL + cfg.KR + cfg.N * cfg.KR)
    total = int(q.global_usage.sum().item())
    same_ncw = _identical(torch.stack([outs[-1]['num_codewords'], outs[-1]['num_codewords_negative']]))
    ok = same_params and same_usage and same_ncw and changed and total == vec_per_step * steps * world
    if rank == 0:
        print(f'[ddp_check] fresh run (initialize=True): codebook re-initialised {changed}; identical parameters across '
              f'{world} ranks after {steps} steps: {same_params}; identical global usage histogram: {same_usage} '
              f'(sum {total} == {vec_per_step} vectors x {steps} steps x {world} ranks: {total == vec_per_step * steps * world}); '
              f'identical num_codewords: {same_ncw} ({int(outs[-1]["num_codewords"])}/{int(outs[-1]["num_codewords_negative"])})')
    return ok


def _flat_of(prm, trainer):
    """initial parameters in the trainer's flat order"""
    named = product_named_params(trainer)
    by_ptr = {p.data_ptr(): k for k, p in named.items()}
    out = torch.zeros_like(trainer.optimizer.flat.data)
    for p, o in zip(trainer.optimizer.flat.params, trainer.optimizer.flat.offsets):
        k = by_ptr[p.data_ptr()]
        out[o:o + p.numel()] = prm[k].reshape(-1).to(out.device)
    return out


if __name__ == '__main__':
    main()
