"""Shared test helpers: build the product modules for an oracle Cfg and load an oracle parameter dict into them."""
import types

import torch

from oracle import params as P


def build_product_encoder(cfg, model_dir='/tmp/vqcpc_b200_test'):
    from VQCPCB.data_processor.bach_cpc_data_processor import BachCPCDataProcessor
    from VQCPCB.downscalers.relative_transformer_downscaler import RelativeTransformerDownscaler
    from VQCPCB.quantizer.vector_quantizer import ProductVectorQuantizer
    from VQCPCB.upscalers.mlp_upscaler import MlpUpscaler
    from VQCPCB.encoder import Encoder
    S = cfg.num_voices * cfg.ticks_per_block
    dp = BachCPCDataProcessor(cfg.embedding_size, cfg.ticks_per_block, cfg.num_voices, list(cfg.vocab), S)
    if cfg.get('downscaler', 'transformer') == 'lstm':
        from VQCPCB.downscalers.lstm_downscaler import LstmDownscaler
        down = LstmDownscaler(cfg.embedding_size, cfg.codebook_dim, cfg.num_voices, list(cfg.downscale_factors),
                              cfg.lstm_hidden, cfg.lstm_layers, 0.0, cfg.lstm_bidirectional)
    elif cfg.get('downscaler', 'transformer') == 'transformer_linear':
        from VQCPCB.downscalers.relative_transformer_downscaler_linear import RelativeTransformerDownscalerLinear
        down = RelativeTransformerDownscalerLinear(cfg.embedding_size, cfg.codebook_dim, cfg.num_voices,
                                                   list(cfg.downscale_factors), cfg.d_model, cfg.n_head,
                                                   list(cfg.list_of_num_layers), cfg.dim_feedforward, 0.0)
    else:
        down = RelativeTransformerDownscaler(cfg.embedding_size, cfg.codebook_dim, cfg.num_voices, list(cfg.downscale_factors),
                                             cfg.d_model, cfg.n_head, list(cfg.list_of_num_layers), cfg.dim_feedforward, 0.0)
    if cfg.codebook_size is None:
        from VQCPCB.quantizer.vector_quantizer import NoQuantization
        vq = NoQuantization(cfg.codebook_dim)
    else:
        vq = ProductVectorQuantizer(cfg.codebook_size, cfg.codebook_dim, cfg.commitment_cost, cfg.num_codebooks, False, False,
                                    cfg.squared_l2_norm)
    up = MlpUpscaler(cfg.codebook_dim, cfg.z_dim, cfg.up_hidden, 0.0) if cfg.upscaler else None
    return Encoder(model_dir, dp, down, vq, up)


def build_product_trainer(cfg, model_dir='/tmp/vqcpc_b200_test', bidirectional=None, device='cuda'):
    if bidirectional is None:
        bidirectional = bool(cfg.get('bidirectional', False))
    from VQCPCB.vqcpc_encoder_trainer import VQCPCEncoderTrainer
    enc = build_product_encoder(cfg, model_dir)
    gen = types.SimpleNamespace(num_blocks_right=cfg.KR)
    tr = VQCPCEncoderTrainer(model_dir, gen, enc,
                             dict(output_dim=cfg.c_dim, hidden_size=cfg.c_hidden, num_layers=cfg.c_layers, dropout=0.0,
                                  bidirectional=bidirectional), cfg.quantization_weighting)
    tr.to(device)
    return tr


def product_named_params(trainer):
    enc = trainer.encoder
    out = {}
    for f, mod in [('data_processor', enc.data_processor), ('downscaler', enc.downscaler), ('quantizer', enc.quantizer),
                   ('upscaler', enc.upscaler), ('c_module', trainer.c_module), ('fks_module', trainer.fks_module),
                   ('c_module_back', trainer.c_module_back), ('fks_module_back', getattr(trainer, 'fks_module_back', None))]:
        if mod is None:
            continue
        for k, p in mod.named_parameters():
            out[f'{f}.{k}'] = p
    return out


def load_params_into(trainer, prm):
    """STRICT load of the oracle's flat dict (reference checkpoint key names) into the product modules."""
    files = P.split_by_file(prm)
    enc = trainer.encoder
    dev = trainer.fks_module.W.device
    mods = {'data_processor': enc.data_processor, 'downscaler': enc.downscaler, 'quantizer': enc.quantizer,
            'upscaler': enc.upscaler, 'c_module': trainer.c_module, 'fks_module': trainer.fks_module,
            'c_module_back': trainer.c_module_back, 'fks_module_back': getattr(trainer, 'fks_module_back', None)}
    for f, sd in files.items():
        mods[f].load_state_dict({k: v.to(dev) for k, v in sd.items()}, strict=True)
    from VQCPCB import ops
    ops.bump_weight_version()


def rel_err(a, b):
    a = a.detach().double().cpu().reshape(-1)
    b = b.detach().double().cpu().reshape(-1)
    return ((a - b).norm() / (b.norm() + 1e-30)).item()
