"""VQ-CPC encoder with the GRU block downscaler - the shape of the reference's ``encoder_random_config.py`` (and of 9
more of its 14 shipped encoder configs): one bidirectional 2-layer GRU pair per 16-token block, K = 32 codes of
dimension 3, a 6 + 6 block window, dropout 0.1 everywhere (GRU layers, upscaler, context network).

Environment overrides: VQCPC_BATCH, VQCPC_BLOCKS (left == right), VQCPC_CODEBOOK, VQCPC_DATASET, VQCPC_DROPOUT
"""
import os
from pathlib import Path

_voices, _ticks = 4, 4
_blocks = int(os.environ.get('VQCPC_BLOCKS', 6))
_tokens_per_block = _voices * _ticks
_p = float(os.environ.get('VQCPC_DROPOUT', 0.1))

config = dict(
    training_method='vqcpc', dataset=os.environ.get('VQCPC_DATASET', 'synthetic'),
    dataloader_generator_kwargs=dict(num_tokens_per_block=_tokens_per_block, num_blocks_left=_blocks,
                                     num_blocks_right=_blocks, negative_sampling_method='random',
                                     num_negative_samples=15, sequences_size=1),
    subdivision=_ticks,
    data_processor_type='bach_cpc',
    data_processor_kwargs=dict(embedding_size=32),
    downscaler_type='lstm_downscaler',
    downscaler_kwargs=dict(downscale_factors=[_tokens_per_block], hidden_size=512, num_layers=2, dropout=_p,
                           bidirectional=True),
    quantizer_type='commitment',
    quantizer_kwargs=dict(num_codebooks=1, codebook_size=int(os.environ.get('VQCPC_CODEBOOK', 32)), codebook_dim=3,
                          commitment_cost=0.25, use_batch_norm=False, squared_l2_norm=True),
    upscaler_type='mlp_upscaler',
    upscaler_kwargs=dict(output_dim=32, hidden_size=512, dropout=_p),
    auxiliary_networks_kwargs=dict(
        quantization_weighting=0.5,
        c_net_kwargs=dict(output_dim=32, hidden_size=512, num_layers=2, dropout=_p, bidirectional=False)),
    lr=1e-4, schedule_lr=False, batch_size=int(os.environ.get('VQCPC_BATCH', 16)), num_batches=16, num_epochs=20000,
    quantizer_regularization=dict(corrupt_labels=False),
    timestamp=None, savename=Path(__file__).stem,
)
