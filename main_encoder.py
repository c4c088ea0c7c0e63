"""Drop-in for the reference's ``main_encoder.py`` training branch (main_encoder.py:16-97):

    python main_encoder.py -t -c vqcpc-bach_b200/VQCPCB/configs/encoder_vqcpc.py [-l] [--num_workers N]

Same flags; the config is a Python module exposing ``config``.  Under ``torchrun`` every rank runs this script with its
own GPU and data shard (one NCCL all-reduce per step).  The post-training cluster plots of the reference
(main_encoder.py:99-118) need music21/matplotlib and are not part of the hot path.
"""
import importlib.util
import os
import shutil
import sys
from datetime import datetime

import click
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, 'vqcpc-bach_b200'))

from VQCPCB import distributed as vq_dist                                                   # noqa: E402
from VQCPCB.getters import get_dataloader_generator, get_encoder, get_encoder_trainer        # noqa: E402


def _load_config(path):
    spec = importlib.util.spec_from_file_location('vqcpc_user_config', path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.config


@click.command()
@click.option('-t', '--train', is_flag=True)
@click.option('-l', '--load', is_flag=True)
@click.option('-c', '--config', type=click.Path(exists=True))
@click.option('--num_workers', type=int, default=0)
def main(train, load, num_workers, config):
    rank, world, local = vq_dist.init_from_env()
    if not torch.cuda.is_available():
        raise RuntimeError('main_encoder.py (B200 path) needs a CUDA device; there is no CPU fallback')
    torch.cuda.set_device(local)
    device = 'cuda'
    print(f'Using GPU {local} (rank {rank} of {world})')

    config_path = config
    config = _load_config(config_path)
    if config['timestamp'] is None:
        config['timestamp'] = datetime.now().strftime('%Y-%m-%d_%H-%M-%S')
    model_dir = os.path.dirname(config_path) if load else f'models/{config["savename"]}_{config["timestamp"]}'
    config['quantizer_kwargs']['initialize'] = not load
    config['dataloader_generator_kwargs']['seed'] = vq_dist.shard_seed(config['dataloader_generator_kwargs'].get('seed', 1234))

    dataloader_generator = get_dataloader_generator(dataset=config['dataset'], training_method=config['training_method'],
                                                    dataloader_generator_kwargs=config['dataloader_generator_kwargs'])
    encoder = get_encoder(model_dir=model_dir, dataloader_generator=dataloader_generator, config=config)
    trainer = get_encoder_trainer(model_dir=model_dir, dataloader_generator=dataloader_generator,
                                  training_method=config['training_method'], encoder=encoder,
                                  auxiliary_networks_kwargs=config['auxiliary_networks_kwargs'])
    if load:
        trainer.load(early_stopped=False, device=device)
    trainer.to(device)

    if train:
        if not load and rank == 0:
            os.makedirs(model_dir, exist_ok=True)
            shutil.copy(config_path, f'{model_dir}/config.py')
        trainer.train_model(batch_size=config['batch_size'], num_batches=config['num_batches'],
                            num_epochs=config['num_epochs'], lr=config['lr'], schedule_lr=config.get('schedule_lr', False),
                            corrupt_labels=config['quantizer_regularization']['corrupt_labels'], plot=(rank == 0),
                            num_workers=num_workers)


if __name__ == '__main__':
    main()
