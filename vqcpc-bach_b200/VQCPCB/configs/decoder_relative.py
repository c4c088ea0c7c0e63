"""Decoder configuration in the schema of the reference's ``decoder_relative_AC_*_C_*.py`` files (BASELINE config 5: generation
of 1024-token windows = 64 beats from the codes of a 16-tokens-per-code encoder).  The reference points ``config_encoder`` at a
trained model directory; here it defaults to the repo's encoder configuration with random weights unless a path is given.

Environment overrides: VQCPC_DECODER_TYPE (transformer_relative | transformer_relative_diagonal | transformer_relative_fullCross |
transformer_relative_full), VQCPC_DECODER_BEATS (sequences_size), VQCPC_DATASET (synthetic | bach), VQCPC_CODEBOOK."""
import os
from pathlib import Path

config = {
    'config_encoder':              os.environ.get('VQCPC_CONFIG_ENCODER'),          # None: encoder_vqcpc.py, untrained
    'training_method':             'decoder',
    'dataset':                     os.environ.get('VQCPC_DATASET', 'synthetic'),
    'dataloader_generator_kwargs': dict(sequences_size=int(os.environ.get('VQCPC_DECODER_BEATS', 64))),
    'data_processor_type':         'bach',
    'data_processor_kwargs':       dict(embedding_size=32),
    'decoder_type':                os.environ.get('VQCPC_DECODER_TYPE', 'transformer_relative'),
    'decoder_kwargs':              dict(d_model=512, n_head=8, num_encoder_layers=3, num_decoder_layers=3, dim_feedforward=1024,
                                        positional_embedding_size=8, dropout=0.2),
    'lr': 1e-4, 'schedule_lr': True, 'batch_size': 32, 'num_batches': 2048, 'num_epochs': 20000,
    'timestamp': None, 'savename': Path(__file__).stem,
}
