"""VQ-CPC encoder config named by BASELINE.json (the reference ships no file of this name; this one follows the
schema of its ``encoder_random_transfo_config.py`` with the two keys that file forgets — ``schedule_lr`` and
``c_net_kwargs['bidirectional']`` — filled in, K=512 codes, a 256-token window (8 + 8 blocks) and dropout 0).

Environment overrides (used by bench.py / tests, so one file covers configs[0..2] of BASELINE.json):
    VQCPC_BATCH, VQCPC_BLOCKS (left == right), VQCPC_CODEBOOK, VQCPC_DATASET, VQCPC_NUM_BATCHES, VQCPC_EPOCHS
"""
import os
from pathlib import Path

_voices, _ticks = 4, 4
_blocks = int(os.environ.get('VQCPC_BLOCKS', 8))
_tokens_per_block = _voices * _ticks


def _dataloader():
    return dict(num_tokens_per_block=_tokens_per_block, num_blocks_left=_blocks, num_blocks_right=_blocks,
                negative_sampling_method='random', num_negative_samples=15, sequences_size=1)


def _encoder():
    return dict(
        data_processor_type='bach_cpc',
        data_processor_kwargs=dict(embedding_size=32),
        downscaler_type='relative_transformer_downscaler',
        downscaler_kwargs=dict(downscale_factors=[4, 4], num_channels=_voices, d_model=512, n_head=8,
                               list_of_num_layers=[2, 2], dim_feedforward=2048, dropout=0.0),
        quantizer_type='commitment',
        quantizer_kwargs=dict(num_codebooks=1, codebook_size=int(os.environ.get('VQCPC_CODEBOOK', 512)), codebook_dim=3,
                              commitment_cost=0.25, use_batch_norm=False, squared_l2_norm=True),
        upscaler_type='mlp_upscaler',
        upscaler_kwargs=dict(output_dim=32, hidden_size=512, dropout=0.0),
    )


def _training():
    return dict(lr=1e-4, schedule_lr=False, batch_size=int(os.environ.get('VQCPC_BATCH', 4096)),
                num_batches=int(os.environ.get('VQCPC_NUM_BATCHES', 16)), num_epochs=int(os.environ.get('VQCPC_EPOCHS', 20000)), quantizer_regularization=dict(corrupt_labels=False))


config = dict(training_method='vqcpc', dataset=os.environ.get('VQCPC_DATASET', 'synthetic'),
              dataloader_generator_kwargs=_dataloader(), subdivision=_ticks)
config.update(_encoder())
config['auxiliary_networks_kwargs'] = dict(
    quantization_weighting=0.5,
    c_net_kwargs=dict(output_dim=32, hidden_size=512, num_layers=2, dropout=0.0, bidirectional=False))
config.update(_training())
config.update(timestamp=None, savename=Path(__file__).stem)
