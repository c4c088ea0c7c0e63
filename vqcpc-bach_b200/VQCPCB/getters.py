"""String-keyed factories for the encoder-training path (mirror of VQCPCB/getters.py:24-218, :431-516).

Built: the branches reachable from ``main_encoder.py -t`` with ``training_method='vqcpc'`` and ``'student'`` (teacher +
auxiliary decoder, getters.py:221-262, :444-482) and ``get_decoder`` for the relative decoders (getters.py:274-395; inference path,
SURVEY.md §8f row 3); the prior branch raises NotImplementedError.  Config keys that several shipped
configs omit (``c_net_kwargs['bidirectional']``, ``schedule_lr``) get the reference's de-facto defaults.
"""
from .data_processor.bach_cpc_data_processor import BachCPCDataProcessor
from .data_processor.bach_data_processor import BachDataProcessor
from .dataloaders.synthetic_cpc_dataloader import SyntheticCPCDataloaderGenerator
from .dataloaders.synthetic_dataloader import SyntheticDataloaderGenerator
from .downscalers.lstm_downscaler import LstmDownscaler
from .downscalers.relative_transformer_downscaler import RelativeTransformerDownscaler
from .downscalers.relative_transformer_downscaler_linear import RelativeTransformerDownscalerLinear
from .encoder import Encoder
from .quantizer.vector_quantizer import ProductVectorQuantizer, NoQuantization
from .upscalers.mlp_upscaler import MlpUpscaler
from .vqcpc_encoder_trainer import VQCPCEncoderTrainer

_NEXT = 'is outside the VQ-CPC encoder-training hot path built so far (SURVEY.md §8f "next" rows)'


def get_dataloader_generator(dataset, training_method, dataloader_generator_kwargs):
    """'bach' needs the music21 chorale corpus (absent offline); 'synthetic' is the in-scope producer with the same
    batch-dict contract."""
    kw = dataloader_generator_kwargs
    if training_method.lower() in ('student', 'decoder', 'prior'):                    # getters.py:38-43: whole windows
        if dataset.lower() in ('synthetic', 'bach_synthetic'):
            return SyntheticDataloaderGenerator(sequences_size=kw['sequences_size'], vocab_size=kw.get('vocab_size', 64),
                                                seed=kw.get('seed', 1234))
        if dataset.lower() == 'bach':                        # the reference's cached tensor_dataset (see bach_cpc_dataloader.py)
            from .dataloaders.bach_dataloader import BachDataloaderGenerator
            return BachDataloaderGenerator(sequences_size=kw['sequences_size'], database_root=kw.get('database_root'))
        raise NotImplementedError(f'dataset={dataset!r}: use "bach" (cached tensor_dataset) or "synthetic" '
                                  '(same batch contract: tensor_dict["x"] (batch, events, channels))')
    if training_method.lower() != 'vqcpc':
        raise NotImplementedError(f'training_method={training_method!r} {_NEXT}')
    if dataset.lower() in ('synthetic', 'bach_synthetic'):
        return SyntheticCPCDataloaderGenerator(
            num_tokens_per_block=kw['num_tokens_per_block'], num_blocks_left=kw['num_blocks_left'],
            num_blocks_right=kw['num_blocks_right'], negative_sampling_method=kw['negative_sampling_method'],
            num_negative_samples=kw['num_negative_samples'], sequences_size=kw.get('sequences_size', 1),
            vocab_size=kw.get('vocab_size', 64), seed=kw.get('seed', 1234))
    if dataset.lower() == 'bach':
        # the reference's cached tensor_dataset files (written by its music21 ETL) under database_root / VQCPC_DATA_ROOT
        from .dataloaders.bach_cpc_dataloader import BachCPCDataloaderGenerator
        return BachCPCDataloaderGenerator(
            num_tokens_per_block=kw['num_tokens_per_block'], num_blocks_left=kw['num_blocks_left'],
            num_blocks_right=kw['num_blocks_right'], negative_sampling_method=kw['negative_sampling_method'],
            num_negative_samples=kw['num_negative_samples'], database_root=kw.get('database_root'))
    raise NotImplementedError('If you want to use your own datasets, you need to implement a datasets, '
                              'data_processor and dataloader')


def get_downscaler(downscaler_type, downscaler_kwargs):
    kw = downscaler_kwargs
    if downscaler_type == 'relative_transformer_downscaler':
        return RelativeTransformerDownscaler(
            input_dim=kw['input_dim'], output_dim=kw['output_dim'], downscale_factors=kw['downscale_factors'],
            num_channels=kw['num_channels'], d_model=kw['d_model'], n_head=kw['n_head'],
            list_of_num_layers=kw['list_of_num_layers'], dim_feedforward=kw['dim_feedforward'], dropout=kw['dropout'])
    if downscaler_type == 'relative_transformer_downscaler_linear':
        return RelativeTransformerDownscalerLinear(
            input_dim=kw['input_dim'], output_dim=kw['output_dim'], downscale_factors=kw['downscale_factors'],
            num_channels=kw['num_channels'], d_model=kw['d_model'], n_head=kw['n_head'],
            list_of_num_layers=kw['list_of_num_layers'], dim_feedforward=kw['dim_feedforward'], dropout=kw['dropout'])
    if downscaler_type == 'lstm_downscaler':
        return LstmDownscaler(
            input_dim=kw['input_dim'], output_dim=kw['output_dim'], num_channels=kw['num_channels'],
            downscale_factors=kw['downscale_factors'], hidden_size=kw['hidden_size'], num_layers=kw['num_layers'],
            dropout=kw['dropout'], bidirectional=kw['bidirectional'])
    if downscaler_type == 'mlp_downscaler':
        raise NotImplementedError(f'downscaler_type={downscaler_type!r} {_NEXT}')
    raise NotImplementedError


def get_upscaler(upscaler_type, upscaler_kwargs):
    if upscaler_type == 'mlp_upscaler':
        return MlpUpscaler(input_dim=upscaler_kwargs['input_dim'], output_dim=upscaler_kwargs['output_dim'],
                           hidden_size=upscaler_kwargs['hidden_size'], dropout=upscaler_kwargs['dropout'])
    if upscaler_type is None:
        return None
    raise NotImplementedError


def get_data_processor(dataloader_generator, data_processor_type, data_processor_kwargs):
    if data_processor_type == 'bach':
        ds = dataloader_generator.dataset
        return BachDataProcessor(embedding_size=data_processor_kwargs['embedding_size'],
                                 num_events=ds.sequences_size * ds.subdivision,
                                 num_tokens_per_channel=[len(d) for d in ds.index2note_dicts])
    if data_processor_type == 'bach_cpc':
        ds = dataloader_generator.dataset_positive
        dp = BachCPCDataProcessor(embedding_size=data_processor_kwargs['embedding_size'],
                                  num_events=ds.sequences_size * ds.subdivision,
                                  num_channels=dataloader_generator.num_channels,
                                  num_tokens_per_channel=[len(d) for d in ds.index2note_dicts],
                                  num_tokens_per_block=dataloader_generator.num_tokens_per_block)
        assert dataloader_generator.num_channels == dp.num_channels
        return dp
    raise NotImplementedError


def get_encoder(model_dir, dataloader_generator, config):
    method = config['training_method'].lower()
    if method not in ('vqcpc', 'student'):
        raise NotImplementedError(f'training_method={method!r} {_NEXT}')
    qk = config['quantizer_kwargs']
    dk = config['downscaler_kwargs']
    data_processor = get_data_processor(dataloader_generator, config['data_processor_type'], config['data_processor_kwargs'])
    # runtime-injected keys (getters.py:138-141, :161)
    dk['input_dim'] = data_processor.embedding_size
    dk['output_dim'] = qk['codebook_dim']
    dk['num_tokens'] = data_processor.num_events * data_processor.num_channels
    dk['num_channels'] = data_processor.num_channels
    downscaler = get_downscaler(config['downscaler_type'], dk)
    if config['quantizer_type'] == 'commitment':
        quantizer = ProductVectorQuantizer(
            codebook_size=qk['codebook_size'], num_codebooks=qk['num_codebooks'], codebook_dim=qk['codebook_dim'],
            initialize=qk.get('initialize', True), squared_l2_norm=qk['squared_l2_norm'],
            use_batch_norm=qk['use_batch_norm'], commitment_cost=qk['commitment_cost'])
    elif config['quantizer_type'] is None:
        quantizer = NoQuantization(codebook_dim=qk['codebook_dim'])
    else:
        raise NotImplementedError(f"quantizer_type={config['quantizer_type']!r}")
    upscaler = None
    if config.get('upscaler_type') is not None:
        uk = config['upscaler_kwargs']
        uk['input_dim'] = qk['codebook_dim']
        upscaler = get_upscaler(config['upscaler_type'], uk)
    return Encoder(model_dir=model_dir, data_processor=data_processor, downscaler=downscaler, quantizer=quantizer,
                   upscaler=upscaler)


def get_encoder_trainer(model_dir, dataloader_generator, training_method, encoder, auxiliary_networks_kwargs):
    if training_method.lower() == 'vqcpc':
        return VQCPCEncoderTrainer(model_dir=model_dir, dataloader_generator=dataloader_generator, encoder=encoder,
                                   c_net_kwargs=auxiliary_networks_kwargs['c_net_kwargs'],
                                   quantization_weighting=auxiliary_networks_kwargs['quantization_weighting'])
    if training_method.lower() == 'student':
        import numpy as np
        from .student_encoder_trainer import StudentEncoderTrainer
        # runtime-injected keys (getters.py:446-468)
        tk = auxiliary_networks_kwargs['teacher_kwargs']
        tk['num_tokens_per_channel'] = encoder.data_processor.num_tokens_per_channel
        tk['num_tokens'] = encoder.data_processor.num_tokens
        teacher = get_teacher(auxiliary_networks_kwargs['teacher_type'], tk, dataloader_generator)
        ak = auxiliary_networks_kwargs['auxiliary_decoder_kwargs']
        ak['num_tokens_per_channel'] = encoder.data_processor.num_tokens_per_channel
        ak['codebook_dim'] = encoder.quantizer.codebook_dim
        ak['upscale_factors'] = list(reversed(encoder.downscaler.downscale_factors))
        ak['num_tokens_bottleneck'] = encoder.data_processor.num_tokens // int(np.prod(encoder.downscaler.downscale_factors))
        auxiliary_decoder = get_auxiliary_decoder(auxiliary_networks_kwargs['auxiliary_decoder_type'], ak)
        return StudentEncoderTrainer(model_dir=model_dir, dataloader_generator=dataloader_generator, encoder=encoder,
                                     teacher=teacher, auxiliary_decoder=auxiliary_decoder,
                                     quantization_weighting=auxiliary_networks_kwargs['quantization_weighting'],
                                     num_events_masked=auxiliary_networks_kwargs['num_events_masked'])
    raise NotImplementedError(f'training_method={training_method!r} {_NEXT}')


def get_teacher(teacher_type, teacher_kwargs, dataloader_generator):
    """getters.py:221-241 (only the relative teacher is shipped with a config)."""
    from .teachers.teacher_relative import TeacherRelative
    if teacher_type != 'relative':
        raise NotImplementedError(f'teacher_type={teacher_type!r} {_NEXT}')
    dpc = teacher_kwargs['data_processor_config']
    data_processor = get_data_processor(dataloader_generator, dpc['data_processor_type'], dpc['data_processor_kwargs'])
    return TeacherRelative(num_layers=teacher_kwargs['num_layers'], num_tokens_per_channel=teacher_kwargs['num_tokens_per_channel'],
                           d_model=teacher_kwargs['d_model'], positional_embedding_size=teacher_kwargs['positional_embedding_size'],
                           dim_feedforward=teacher_kwargs['dim_feedforward'], n_head=teacher_kwargs['n_head'],
                           dropout=teacher_kwargs['dropout'], num_tokens=teacher_kwargs['num_tokens'], data_processor=data_processor)


def get_auxiliary_decoder(auxiliary_decoder_type, auxiliary_decoder_kwargs):
    """getters.py:244-272 (relative variant)."""
    from .auxiliary_decoders.auxiliary_decoder_relative import AuxiliaryDecoderRelative
    if auxiliary_decoder_type != 'relative':
        raise NotImplementedError(f'auxiliary_decoder_type={auxiliary_decoder_type!r} {_NEXT}')
    k = auxiliary_decoder_kwargs
    return AuxiliaryDecoderRelative(num_tokens_per_channel=k['num_tokens_per_channel'], codebook_dim=k['codebook_dim'],
                                    upscale_factors=k['upscale_factors'], n_head=k['n_head'], dim_feedforward=k['dim_feedforward'],
                                    list_of_num_layers=k['list_of_num_layers'], d_model=k['d_model'],
                                    num_tokens_bottleneck=k['num_tokens_bottleneck'], dropout=k['dropout'])


_DECODER_TYPES = {   # getters.py:290-393: decoder_type -> (encoder attention, cross attention); 'transformer' = absolute positions
    'transformer': ('anticausal', 'full'),
    'transformer_relative': ('anticausal', 'anticausal'),
    'transformer_relative_fullCross': ('anticausal', 'full'),
    'transformer_relative_diagonal': ('anticausal', 'diagonal'),
    'transformer_relative_full': ('full', 'full'),
}


def get_decoder(model_dir, dataloader_generator, data_processor, encoder, decoder_type, decoder_kwargs):
    """getters.py:274-395."""
    import numpy as np
    from .decoders.decoder import Decoder
    if decoder_type not in _DECODER_TYPES:
        raise NotImplementedError(f'decoder_type={decoder_type!r}: one of {sorted(_DECODER_TYPES)}')
    enc_att, cross_att = _DECODER_TYPES[decoder_type]
    num_channels_decoder, num_events_decoder = data_processor.num_channels, data_processor.num_events
    num_channels_encoder = 1
    num_events_encoder = int((num_events_decoder * num_channels_decoder) //
                             (np.prod(encoder.downscaler.downscale_factors) * num_channels_encoder))
    k = decoder_kwargs
    return Decoder(model_dir=model_dir, dataloader_generator=dataloader_generator, data_processor=data_processor, encoder=encoder,
                   transformer_type='absolute' if decoder_type == 'transformer' else 'relative', encoder_attention_type=enc_att, cross_attention_type=cross_att, d_model=k['d_model'],
                   num_encoder_layers=k['num_encoder_layers'], num_decoder_layers=k['num_decoder_layers'], n_head=k['n_head'],
                   dim_feedforward=k['dim_feedforward'], dropout=k['dropout'],
                   positional_embedding_size=k['positional_embedding_size'], num_channels_encoder=num_channels_encoder,
                   num_events_encoder=num_events_encoder, num_channels_decoder=num_channels_decoder,
                   num_events_decoder=num_events_decoder)
