"""CPU oracle of the CPC data path: datasets/chorale_dataset.py:542-595 (``data_loaders`` over the cached
tensor dataset) and dataloaders/bach_cpc_dataloader.py:110-252 (batch assembly, both negative sampling methods).

TEST INFRASTRUCTURE ONLY (see oracle/vqcpc_oracle.py).  Restated with real ``torch.utils.data.DataLoader``s over host tensors,
like the reference; pinned against the live reference by ``oracle/make_golden_dataloader.py`` -> ``tests/golden/dataloader.pt``."""
import torch
from torch.utils.data import DataLoader, TensorDataset

NUM_VOICES = 4


def data_loaders(tensors, batch_size, split=(0.85, 0.10), shuffle_train=True, shuffle_val=False):
    """chorale_dataset.py:542-595 for an in-memory tuple of tensors."""
    n = tensors[0].shape[0]
    a, b = split
    parts = [tuple(t[: int(a * n)] for t in tensors), tuple(t[int(a * n): int((a + b) * n)] for t in tensors),
             tuple(t[int((a + b) * n):] for t in tensors)]
    return [DataLoader(TensorDataset(*p), batch_size=batch_size, shuffle=sh, num_workers=0, drop_last=True)
            for p, sh in zip(parts, (shuffle_train, shuffle_val, False))]


def build_negatives_same_seq(x_a, x_b, num_blocks_right, evb):
    """bach_cpc_dataloader.py:162-184 (x_a, x_b: (B, V, ticks))."""
    out = []
    b_split = x_b.unsqueeze(1).split(evb, dim=3)
    for k in range(num_blocks_right):
        out.append(torch.cat((*x_a.unsqueeze(1).split(evb, dim=3), *b_split[:k], *b_split[k + 1:]), dim=1).unsqueeze(2))
    return torch.cat(out, dim=2)                                          # (B, M, KR, V, evb)


def cpc_batches(positive, negative, cfg, batch_size, shuffle_train=True, shuffle_val=False):
    """-> three generators of batch dicts.  positive / negative: tuples of host tensors (score (N, V, ticks), ...)."""
    evb = cfg['num_tokens_per_block'] // NUM_VOICES
    KL, KR, N = cfg['num_blocks_left'], cfg['num_blocks_right'], cfg['num_negative_samples']
    left = KL * evb
    if cfg['negative_sampling_method'] == 'same_sequence':
        def agg(dl):
            for p in dl:
                p = p[0]
                xl, xr = p[:, :, :left], p[:, :, left:]
                yield {'x_left': xl.transpose(1, 2), 'x_right': xr.transpose(1, 2),
                       'negative_samples': build_negatives_same_seq(xl, xr, KR, evb).transpose(3, 4),
                       'negative_samples_back': build_negatives_same_seq(xr, xl, KR, evb).transpose(3, 4)}
        return [agg(dl) for dl in data_loaders(positive, batch_size, shuffle_train=shuffle_train, shuffle_val=shuffle_val)]
    pos = data_loaders(positive, batch_size, shuffle_train=shuffle_train, shuffle_val=shuffle_val)
    neg = data_loaders(negative, batch_size * N * KR, shuffle_train=shuffle_train, shuffle_val=shuffle_val)
    nbk = data_loaders(negative, batch_size * N * KR, shuffle_train=shuffle_train, shuffle_val=shuffle_val)

    def agg(dp, dn, db):
        for p, n, nb in zip(dp, dn, db):
            p = p[0]
            yield {'x_left': p[:, :, :left].transpose(1, 2), 'x_right': p[:, :, left:].transpose(1, 2),
                   'negative_samples': n[0].view(batch_size, N, KR, NUM_VOICES, evb).transpose(3, 4),
                   'negative_samples_back': nb[0].view(batch_size, N, KR, NUM_VOICES, evb).transpose(3, 4)}
    return [agg(*a) for a in zip(pos, neg, nbk)]


def window_batches(tensors, batch_size, shuffle_train=True, shuffle_val=False):
    """bach_dataloader.py:22-33: {'x': (B, ticks, V)} streams."""
    return [({'x': t[0].transpose(1, 2)} for t in dl)
            for dl in data_loaders(tensors, batch_size, shuffle_train=shuffle_train, shuffle_val=shuffle_val)]


def synthetic_chorale_tensors(n, ticks, seed, vocab=(50, 47, 45, 52)):
    """A stand-in for the cached dataset: (score (n, 4, ticks) int32, metadata (n, 4, ticks, 1) int32) like :166-178."""
    g = torch.Generator().manual_seed(seed)
    score = torch.stack([torch.randint(0, v, (n, ticks), generator=g) for v in vocab], dim=1).int()
    meta = torch.zeros(n, len(vocab), ticks, 1, dtype=torch.int32)
    return score, meta
