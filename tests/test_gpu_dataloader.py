"""GPU: CPC batches gathered on the device from the resident chorale dataset against batches recorded from
the live reference and against the CPU oracle."""
import os

import pytest
import torch

from oracle import dataloader_oracle as O

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), 'golden', 'dataloader.pt')


def _write_cache(root, sequences_size, tensors):
    from VQCPCB.datasets.chorale_dataset import CachedChoraleBeatsDataset
    ds = CachedChoraleBeatsDataset(sequences_size=sequences_size, database_root=str(root))
    os.makedirs(ds.cache_dir, exist_ok=True)
    torch.save(torch.utils.data.TensorDataset(*tensors), f'{ds.cache_dir}/tensor_dataset')


def _generator(root, cfg, pos, neg):
    from VQCPCB import getters
    per_beat = 16
    _write_cache(root, cfg['num_tokens_per_block'] * (cfg['num_blocks_left'] + cfg['num_blocks_right']) // per_beat, pos)
    if neg is not None:
        _write_cache(root, cfg['num_tokens_per_block'] // per_beat, neg)
    return getters.get_dataloader_generator('bach', 'vqcpc', dict(cfg, database_root=str(root)))


@pytest.mark.parametrize('case', ['random', 'same_sequence', 'same_sequence_kl4_kr2'])
def test_device_batches_equal_reference_golden(case, tmp_path):
    """Same cache files, same torch seed -> the batches the reference's DataLoader pipeline produced, bit for bit."""
    rec = torch.load(GOLD)[case]
    pos = (rec['positive'], torch.zeros(rec['positive'].shape + (1,), dtype=torch.int32))
    neg = None if rec['negative'] is None else (rec['negative'], torch.zeros(rec['negative'].shape + (1,), dtype=torch.int32))
    gen = _generator(tmp_path, rec['cfg'], pos, neg)
    torch.manual_seed(rec['seed'])
    train, val, _ = gen.dataloaders(batch_size=rec['batch_size'])
    for stream, want in ((train, rec['train']), (val, rec['val'])):
        for w in want:
            g = next(stream)
            for k in w:
                assert g[k].is_cuda and g[k].dtype == torch.int64 and g[k].shape == w[k].shape
                assert torch.equal(g[k].cpu(), w[k].long()), (case, k)


@pytest.mark.parametrize('method,KL,KR,dtype', [('random', 8, 8, torch.int32), ('same_sequence', 6, 6, torch.int64),
                                                 ('same_sequence', 5, 3, torch.int32)])
def test_device_batches_equal_oracle_full_epoch(method, KL, KR, dtype, tmp_path):
    """A whole shuffled epoch at the reference's window sizes against the oracle's DataLoader pipeline (int32 and int64 caches)."""
    cfg = dict(num_tokens_per_block=16, num_blocks_left=KL, num_blocks_right=KR, negative_sampling_method=method,
               num_negative_samples=15)
    B = 8
    pos = tuple(t.to(dtype) for t in O.synthetic_chorale_tensors(1500, (KL + KR) * 4, seed=3))
    neg = tuple(t.to(dtype) for t in O.synthetic_chorale_tensors(12000, 4, seed=4)) if method == 'random' else None
    gen = _generator(tmp_path, cfg, pos, neg)
    torch.manual_seed(77)
    want_streams = O.cpc_batches(pos, neg, cfg, B)
    want = [[b for b in s] for s in want_streams[:2]]
    torch.manual_seed(77)
    got_streams = gen.dataloaders(batch_size=B)
    n = 0
    for gs, ws in zip(got_streams[:2], want):
        for g, w in zip(gs, ws):
            for k in w:
                assert torch.equal(g[k].cpu(), w[k].long()), (method, k)
            n += 1
    assert n == len(want[0]) + len(want[1]) and len(want[0]) >= 4 and len(want[1]) >= 1


def test_trainer_steps_on_device_resident_batches(tmp_path):
    """The encoder trainer consumes the device batches directly (no host copy) - one training step, finite loss."""
    from oracle import vqcpc_oracle as VO
    from tests.helpers import build_product_trainer
    ocfg = VO.default_cfg(codebook_size=16, KL=2, KR=2)
    cfg = dict(num_tokens_per_block=16, num_blocks_left=2, num_blocks_right=2, negative_sampling_method='random',
               num_negative_samples=ocfg['N'])
    pos = O.synthetic_chorale_tensors(400, 16, seed=1, vocab=(60, 60, 60, 60))
    neg = O.synthetic_chorale_tensors(4000, 4, seed=2, vocab=(60, 60, 60, 60))
    gen = _generator(tmp_path, cfg, pos, neg)
    train, _, _ = gen.dataloaders(batch_size=4)
    trainer = build_product_trainer(ocfg)
    trainer.init_optimizers(1e-4, False)
    trainer.train()
    st = trainer.step(next(train), True, False)
    assert torch.isfinite(st['loss']).item()


def test_window_loader_equals_reference_golden(tmp_path):
    """BachDataloaderGenerator (student trainer / Decoder.generate producer): tensor_dict['x'] from the device gather."""
    from VQCPCB import getters
    rec = torch.load(GOLD)['windows']
    tens = (rec['positive'], torch.zeros(rec['positive'].shape + (1,), dtype=torch.int32))
    _write_cache(tmp_path, rec['positive'].shape[2] // 4, tens)
    gen = getters.get_dataloader_generator('bach', 'student', dict(sequences_size=rec['positive'].shape[2] // 4,
                                                                   database_root=str(tmp_path)))
    torch.manual_seed(rec['seed'])
    train, val, _ = gen.dataloaders(batch_size=rec['batch_size'])
    for w in rec['train']:
        g = next(train)['x']
        assert g.is_cuda and torch.equal(g.cpu(), w['x'].long())
    assert torch.equal(next(val)['x'].cpu(), rec['val'][0]['x'].long())
