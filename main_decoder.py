"""Drop-in for the INFERENCE branch of the reference's ``main_decoder.py`` (main_decoder.py:16-160):

    python main_decoder.py -c vqcpc-bach_b200/VQCPCB/configs/decoder_relative.py [-l] [-o] [--num_examples N] [--batch_size B]

Builds the encoder named by ``config['config_encoder']`` (default: the repo's ``encoder_vqcpc.py``), the decoder of
``config['decoder_type']``, optionally loads their checkpoints (``-l``; ``-o`` = over-fitted decoder weights) and generates
``--num_examples`` batches with ``Decoder.generate`` (``--eval_batches N`` first prints the teacher-forced validation loss) (key / value cached, CUDA graph per event, sampling on the device).  Tokens
are written as ``generation_<i>.pt`` = {'original', 'generated', 'recoding'} in the model directory; score writing
(music21) and ``-t`` (decoder training), ``-r`` (reharmonisation) are outside the built path and say so."""
import importlib.util
import os
import sys
from datetime import datetime

import click
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, 'vqcpc-bach_b200'))

from VQCPCB.getters import get_data_processor, get_dataloader_generator, get_decoder, get_encoder    # noqa: E402


def _load_config(path):
    spec = importlib.util.spec_from_file_location('vqcpc_user_config_' + os.path.basename(path).replace('.', '_'), path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.config


@click.command()
@click.option('-t', '--train', is_flag=True)
@click.option('-l', '--load', is_flag=True)
@click.option('-o', '--overfitted', is_flag=True, help='Load over-fitted decoder weights instead of early-stopped (with -l)')
@click.option('-c', '--config', type=click.Path(exists=True))
@click.option('-r', '--reharmonization', is_flag=True)
@click.option('--num_examples', type=int, default=1)
@click.option('--batch_size', type=int, default=3)
@click.option('--temperature', type=float, default=0.95)
@click.option('--top_p', type=float, default=0.8)
@click.option('--top_k', type=int, default=0)
@click.option('--eval_batches', type=int, default=0, help='teacher-forced validation loss over this many batches')
def main(train, load, overfitted, config, reharmonization, num_examples, batch_size, temperature, top_p, top_k, eval_batches):
    if train or reharmonization:
        raise NotImplementedError('decoder training / reharmonisation are outside the built path (inference: forward + generate)')
    if not torch.cuda.is_available():
        raise RuntimeError('main_decoder.py (B200 path) needs a CUDA device; there is no CPU fallback')
    device = 'cuda'
    config_path = config
    config = _load_config(config_path)
    if config['timestamp'] is None:
        config['timestamp'] = datetime.now().strftime('%Y-%m-%d_%H-%M-%S')
    model_dir = os.path.dirname(config_path) if load else f'models/{config["savename"]}_{config["timestamp"]}'
    # ==== encoder (main_decoder.py:58-80) ====
    config_encoder_path = config['config_encoder'] or os.path.join(HERE, 'vqcpc-bach_b200', 'VQCPCB', 'configs', 'encoder_vqcpc.py')
    config_encoder = _load_config(config_encoder_path)
    config_encoder['quantizer_kwargs']['initialize'] = False
    model_dir_encoder = os.path.dirname(config_encoder_path) if config['config_encoder'] is not None else None
    dl_encoder = get_dataloader_generator(dataset=config_encoder['dataset'], training_method=config_encoder['training_method'],
                                          dataloader_generator_kwargs=config_encoder['dataloader_generator_kwargs'])
    encoder = get_encoder(model_dir=model_dir_encoder, dataloader_generator=dl_encoder, config=config_encoder)
    if config['config_encoder'] is not None:
        encoder.load(early_stopped=False, device=device)
    encoder.to(device)
    # ==== decoder (:82-107) ====
    dataloader_generator = get_dataloader_generator(dataset=config['dataset'], training_method=config['training_method'],
                                                    dataloader_generator_kwargs=config['dataloader_generator_kwargs'])
    data_processor = get_data_processor(dataloader_generator=dataloader_generator, data_processor_type=config['data_processor_type'],
                                        data_processor_kwargs=config['data_processor_kwargs'])
    decoder = get_decoder(model_dir=model_dir, dataloader_generator=dataloader_generator, data_processor=data_processor,
                          encoder=encoder, decoder_type=config['decoder_type'], decoder_kwargs=config['decoder_kwargs'])
    decoder.to(device)
    if load:
        decoder.load(early_stopped=not overfitted, device=device)
    os.makedirs(model_dir, exist_ok=True)
    if eval_batches > 0:
        _, val, _ = dataloader_generator.dataloaders(batch_size=config['batch_size'])
        print('validation:', decoder.epoch(val, train=False, num_batches=eval_batches))
    for i in range(num_examples):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        original, generated, recoding = decoder.generate(temperature=temperature, top_p=top_p, top_k=top_k, batch_size=batch_size,
                                                         seed_set='val')
        e1.record()
        torch.cuda.synchronize()
        path = f'{model_dir}/generation_{i}.pt'
        torch.save({'original': original.cpu(), 'generated': generated.cpu(),
                    'recoding': None if recoding is None else recoding.cpu()}, path)
        n = generated.numel()
        print(f'generation {i}: {tuple(generated.shape)} tokens in {e0.elapsed_time(e1):.1f} ms '
              f'({n / e0.elapsed_time(e1) * 1e3:.0f} tokens/s) -> {path}')


if __name__ == '__main__':
    main()
