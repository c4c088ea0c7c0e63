"""Encoder (data_processor -> downscaler -> quantizer -> upscaler) and the base trainer loop.

Class surface of VQCPCB/encoder.py: ``Encoder`` (:14-110; forward :76-95, save/load :47-74, merge_codes :97-110) and
``EncoderTrainer.train_model/plot`` (:231-325).  The cluster-visualisation helpers of the reference (:112-228) need
music21/matplotlib and are outside the hot path.
"""
import os

import torch
from torch import nn

from . import ops
from .utils import dict_pretty_print, flatten


def save_state_dict(module, path):
    """torch.save of a module's state_dict with every tensor CLONED first: once the trainer has re-homed the parameters
    as views into its flat buffer, saving the views would serialise the whole underlying storage into each of the
    per-module files (the reference writes compact per-module files, encoder.py:47-57)."""
    torch.save({k: v.detach().clone() for k, v in module.state_dict().items()}, path)


class Encoder(nn.Module):
    def __init__(self, model_dir, data_processor, downscaler, quantizer, upscaler):
        super().__init__()
        self.data_processor = data_processor
        self.downscaler = downscaler
        self.quantizer = quantizer
        self.upscaler = upscaler
        self.model_dir = model_dir

    def _dir(self, early_stopped):
        return f'{self.model_dir}/early_stopped' if early_stopped else f'{self.model_dir}/overfitted'

    def save(self, early_stopped):
        d = self._dir(early_stopped)
        os.makedirs(d, exist_ok=True)
        save_state_dict(self.data_processor, f'{d}/data_processor')
        save_state_dict(self.downscaler, f'{d}/downscaler')
        save_state_dict(self.quantizer, f'{d}/quantizer')
        if self.upscaler is not None:
            save_state_dict(self.upscaler, f'{d}/upscaler')

    def load(self, early_stopped, device):
        print(f'Loading models {self.__repr__()}')
        d = self._dir(early_stopped)
        if not os.path.exists(d):          # older checkpoints were saved flat in model_dir
            d = self.model_dir
        loc = torch.device(device)
        # in-place copies keep the parameters' storage (they may alias the trainer's flat buffer)
        self.data_processor.load_state_dict(torch.load(f'{d}/data_processor', map_location=loc))
        self.downscaler.load_state_dict(torch.load(f'{d}/downscaler', map_location=loc))
        self.quantizer.load_state_dict(torch.load(f'{d}/quantizer', map_location=loc))
        if self.upscaler:
            self.upscaler.load_state_dict(torch.load(f'{d}/upscaler', map_location=loc))
        ops.bump_weight_version()

    def forward(self, x, corrupt_labels=False):
        """x as produced by the dataloader (..., ticks, voices) ->
        (z_quantized (b, nb, z_dim), encoding_indices (b, nb, num_codebooks) | None, quantization_loss (b, nb))."""
        x_proc = self.data_processor.preprocess(x)
        return self.forward_tokens(x_proc, corrupt_labels=corrupt_labels)

    def forward_tokens(self, x_proc, corrupt_labels=False):
        """Same as ``forward`` from already preprocessed tokens (b, num_blocks, tokens_per_block) on the device.  The
        trainer uses it to push the negatives / left / right blocks of a pass through ONE encoder launch sequence."""
        x_embed = self.data_processor.embed(x_proc)
        x_flat = flatten(x_embed)
        z = self.downscaler.forward(x_flat)
        if self._fusable(corrupt_labels):
            return self._fused_bottleneck(z)
        z_quantized, encoding_indices, quantization_loss = self.quantizer.forward(z, corrupt_labels=corrupt_labels)
        if self.upscaler is not None:
            z_quantized = self.upscaler(z_quantized)
        return z_quantized, encoding_indices, quantization_loss

    # ---- fused quantizer + upscaler (one kernel over all vectors, the MLP evaluated on the Kc code vectors only) ----
    fuse_bottleneck = True

    def _fusable(self, corrupt_labels):
        from .quantizer.vector_quantizer import ProductVectorQuantizer
        from .upscalers.mlp_upscaler import MlpUpscaler
        q, u = self.quantizer, self.upscaler
        return (self.fuse_bottleneck and isinstance(q, ProductVectorQuantizer) and isinstance(u, MlpUpscaler)
                and q.num_codebooks == 1 and q.codebook_dim in (3, 4, 8, 16, 32) and not q.use_batch_norm
                and not (self.training and (corrupt_labels or u.dropout > 0))
                and q.codebook_size * (u.output_dim + 5) * 4 <= 190 * 1024)

    def _fused_bottleneck(self, z):
        q, u = self.quantizer, self.upscaler
        if q.initialize:                      # data-dependent codebook init from this very batch (quirk Q5)
            q._initialize(z.reshape(-1, q.codebook_dim))
        codebook = q.embeddings[0]
        table, jac = u.code_table_and_jacobian(codebook)
        y, idx, loss = ops.FusedVQUpscaleFn.apply(z, codebook, table, jac, q._commitment_cost, q.squared_l2_norm,
                                                  q.usage_counts)
        q.ema_accumulate(z, idx)
        return y, idx, loss

    def forward_tokens_indexed(self, x_proc, y_row0, y_rows):
        """Index path (table mode): tokens (1, P, S) -> (y of the rows [y_row0, +y_rows) (y_rows, z_dim), idx (P,) int64,
        loss (P,), link).  The upscaled vectors of all other rows are never materialised: the trainer scores them from
        their indices (``FksModule.fused_nce_indexed``) and ``link`` carries what the two autograd nodes share."""
        q, u = self.quantizer, self.upscaler
        z = self.downscaler.forward(flatten(self.data_processor.embed(x_proc)))
        codebook = q.embeddings[0]
        table, jac = u.code_table_and_jacobian(codebook)
        link = ops.BottleneckLink()
        link.table = table
        y, idx, loss, token = ops.VQIndexFn.apply(z.reshape(-1, q.codebook_dim), codebook, table, jac, q._commitment_cost,
                                                  q.squared_l2_norm, q.usage_counts, int(y_row0), int(y_rows), link)
        link.token = token
        q.ema_accumulate(z, idx)
        return y, idx, loss, link

    def merge_codes(self, codes):
        """Pack the indices of a product quantizer into one integer per position (encoder.py:97-110)."""
        _, _, num_codebooks = codes.shape
        ret = codes[:, :, 0].clone()
        for i in range(1, num_codebooks):
            ret += codes[:, :, i] * (self.quantizer.codebook_size ** i)
        return ret


class EncoderTrainer(nn.Module):
    """Epoch loop, checkpointing, early stopping on -accuracy, TensorBoard scalars (encoder.py:231-325)."""

    def __init__(self, dataloader_generator):
        super().__init__()
        self.dataloader_generator = dataloader_generator

    def train_model(self, batch_size, num_batches, num_epochs, lr, corrupt_labels, schedule_lr, plot=False,
                    num_workers=0, **kwargs):
        if plot:
            from torch.utils.tensorboard import SummaryWriter
            self.writer = SummaryWriter(f'{self.model_dir}')
        best_val = 1e8
        self.init_optimizers(lr=lr, schedule_lr=schedule_lr)
        for epoch_id in range(num_epochs):
            gen_train, gen_val, _ = self.dataloader_generator.dataloaders(batch_size=batch_size, num_workers=num_workers)
            train_q = self.epoch(data_loader=gen_train, train=True, num_batches=num_batches, corrupt_labels=corrupt_labels)
            del gen_train
            val_q = self.epoch(data_loader=gen_val, train=False,
                               num_batches=num_batches // 2 if num_batches is not None else None,
                               corrupt_labels=corrupt_labels)
            del gen_val
            if getattr(self, 'is_main_process', True):
                print(f'======= Epoch {epoch_id} =======')
                print('---Train---')
                dict_pretty_print(train_q, endstr=' ' * 5)
                print()
                print('---Val---')
                dict_pretty_print(val_q, endstr=' ' * 5)
                print('\n')
                self.save(early_stopped=False)
                if val_q['loss_monitor'] < best_val:
                    self.save(early_stopped=True)
                    best_val = val_q['loss_monitor']
                if plot:
                    self.plot(epoch_id, train_q, val_q)

    def plot(self, epoch_id, monitored_quantities_train, monitored_quantities_val, index_encoder=None):
        suffix = f'_{index_encoder})' if index_encoder is not None else ''
        for split, q in (('train', monitored_quantities_train), ('val', monitored_quantities_val)):
            if q is None:
                continue
            for k, v in q.items():
                if isinstance(v, list):
                    for i, e in enumerate(v):
                        self.writer.add_scalar(f'{k}_{i}{suffix}/{split}', e, epoch_id)
                else:
                    self.writer.add_scalar(f'{k}{suffix}/{split}', v, epoch_id)
