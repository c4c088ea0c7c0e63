"""Small helpers mirroring VQCPCB/utils.py (only what the encoder-training path uses)."""
import torch


def cuda_variable(tensor):
    """Move to the current CUDA device (VQCPCB/utils.py:5-9).  The B200 path has no CPU mode."""
    if not torch.cuda.is_available():
        raise RuntimeError('vqcpc_b200 needs a CUDA device: there is no CPU fallback on the product path')
    return tensor.to('cuda', non_blocking=True)


def to_numpy(tensor):
    return tensor.detach().to('cpu').numpy()


def dict_pretty_print(d, endstr='\n'):
    for key, value in d.items():
        if isinstance(value, list):
            print(f'{key.capitalize()}: [%s]' % ', '.join(map(str, value)))
        else:
            print(f'{key.capitalize()}: {value:.6}', end=endstr)


def flatten(x):
    """(batch, num_events, num_channels, ...) -> (batch, num_events * num_channels, ...), channel fastest
    (VQCPCB/utils.py:52-63)."""
    assert x.dim() >= 3
    b, e, c = x.shape[:3]
    return x.reshape(b, e * c, *x.shape[3:])


def unflatten(sequence, num_channels):
    """Inverse of ``flatten`` (VQCPCB/utils.py:66-81)."""
    b, n = sequence.shape[:2]
    assert n % num_channels == 0
    return sequence.reshape(b, n // num_channels, num_channels, *sequence.shape[2:])
