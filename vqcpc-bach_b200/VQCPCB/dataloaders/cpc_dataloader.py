class CPCDataloaderGenerator:
    """Attribute contract of the CPC batch producer (VQCPCB/dataloaders/cpc_dataloader.py:1-42):
    ``num_tokens_per_block, num_blocks_left, num_blocks_right, negative_sampling_method, num_negative_samples`` and
    ``dataloaders(batch_size, num_workers) -> (train, val, test)`` iterables of batch dicts."""

    def __init__(self, num_tokens_per_block, num_blocks_left, num_blocks_right, negative_sampling_method,
                 num_negative_samples, *args, **kwargs):
        self.num_tokens_per_block = num_tokens_per_block
        self.num_blocks_left = num_blocks_left
        self.num_blocks_right = num_blocks_right
        self.negative_sampling_method = negative_sampling_method
        self.num_negative_samples = num_negative_samples

    def dataloaders(self, batch_size, num_workers=0, shuffle_train=True, shuffle_val=False):
        raise NotImplementedError
