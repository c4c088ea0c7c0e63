"""Student / distillation configuration: the reference's ``configs/encoder_student_config.py`` (BASELINE config 4 family) with
the synthetic dataset (the chorale corpus is not available offline) and dropout 0 (the parity setting of the B200 path)."""
import os
from pathlib import Path

subdivision = 4
# test / smoke overrides (like configs/encoder_vqcpc.py): VQCPC_BATCH, VQCPC_NUM_BATCHES, VQCPC_EPOCHS, VQCPC_STUDENT_LAYERS
_layers = int(os.environ.get('VQCPC_STUDENT_LAYERS', 4))

config = {
    'training_method': 'Student',
    'dataset': 'synthetic',
    'dataloader_generator_kwargs': dict(sequences_size=24),
    'subdivision': subdivision,
    'data_processor_type': 'bach',
    'data_processor_kwargs': dict(embedding_size=32),
    'downscaler_type': 'relative_transformer_downscaler_linear',
    'downscaler_kwargs': dict(downscale_factors=[4, 4], d_model=512, n_head=8, list_of_num_layers=[_layers, _layers], dim_feedforward=2048,
                              attention_bias_type='relative_attention', dropout=0.0),
    'quantizer_type': 'commitment',
    'quantizer_kwargs': dict(num_codebooks=1, codebook_size=32, codebook_dim=3, commitment_cost=0.25, use_batch_norm=False,
                             squared_l2_norm=True, initialize=True),
    'upscaler_type': None,
    'auxiliary_networks_kwargs': {
        'quantization_weighting': 0.1,
        'num_events_masked': 4,
        'teacher_type': 'relative',
        'teacher_kwargs': dict(
            data_processor_config=dict(data_processor_type='bach', data_processor_kwargs=dict(embedding_size=32)),
            num_layers=2 * _layers, positional_embedding_size=8, d_model=512, dim_feedforward=2048, n_head=8, dropout=0.0),
        'auxiliary_decoder_type': 'relative',
        'auxiliary_decoder_kwargs': dict(positional_embedding_size=8, d_model=512, dim_feedforward=2048, n_head=8, dropout=0.0,
                                         list_of_num_layers=[_layers, _layers]),
    },
    'lr': 1e-5,
    'batch_size': int(os.environ.get('VQCPC_BATCH', 8)),
    'num_batches': int(os.environ.get('VQCPC_NUM_BATCHES', 512)),
    'num_epochs': int(os.environ.get('VQCPC_EPOCHS', 20000)),
    'quantizer_regularization': dict(corrupt_labels=False),
    'timestamp': None,
    'savename': Path(__file__).stem,
}
