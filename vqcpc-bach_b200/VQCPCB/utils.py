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


class _SoftmaxCEFn(torch.autograd.Function):
    """Row-wise softmax cross-entropy on the CUDA kernel ``vqcpc_softmax_ce``: hard (int64) or distilled (logits) targets."""

    @staticmethod
    def forward(ctx, logits, target_idx, target_logits):
        from . import ops
        loss, dl = ops.softmax_ce(logits, target_idx=target_idx, target_logits=target_logits, want_grad=True)
        ctx.save_for_backward(dl)
        return loss

    @staticmethod
    def backward(ctx, dloss):
        (dl,) = ctx.saved_tensors
        return dl * dloss.unsqueeze(1), None, None


def categorical_crossentropy(value, target, mask=None):
    """VQCPCB/utils.py:24-49: sum over channels of the cross-entropy at the masked positions.

    value: list of (batch, num_events, num_tokens_of_channel) logits; target, mask: (batch, num_events, num_channels).
    Returns a vector with one entry per masked (batch, event) pair (every batch element masks the same number of events)."""
    total = 0
    for channel_logits, target_channel, mask_channel in zip(value, target.split(1, dim=2), mask.split(1, dim=2)):
        b, _, n = channel_logits.shape
        assert int(mask_channel.sum()) % b == 0
        sel = mask_channel.bool().squeeze(2)
        logits = channel_logits[sel]                       # (masked positions, n)
        tgt = target_channel.squeeze(2)[sel].contiguous()
        total = total + _SoftmaxCEFn.apply(logits.contiguous(), tgt, None)
    return total


def distilled_categorical_crossentropy(value, target, mask=None):
    """VQCPCB/utils.py:131-159: -sum softmax(teacher) log_softmax(student) per batch element, summed over channels and over
    the events whose mask is set (an event counts when more than half of the batch masks it, like the reference's test)."""
    total = 0
    for channel, channel_target, channel_mask in zip(value, target, mask.split(1, dim=2)):
        events = (channel_mask.squeeze(2).float().mean(dim=0) > 0.5).nonzero().flatten().tolist()
        for e in events:
            total = total + _SoftmaxCEFn.apply(channel[:, e].contiguous(), None, channel_target[:, e].detach().contiguous())
    return total
