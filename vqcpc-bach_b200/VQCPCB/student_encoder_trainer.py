"""Student / distillation trainer of the encoder (VQCPCB/student_encoder_trainer.py:12-294).

Per batch (``epoch`` :222-294, here ``step``):
  1. teacher: mask ``num_events_masked`` events around one random event with the mask token (:138-183), predict the event's
     tokens, cross-entropy against the data (:120-136), clip 5 + Adam on the teacher (:246-249);
  2. encoder + auxiliary decoder: ``z_q = encoder(x)``, logits = decoder(z_q), distilled cross-entropy against the DETACHED
     teacher logits at the same event + ``quantization_weighting * mean(quantization loss)`` (:185-220), clip 5 on the decoder
     and on the encoder SEPARATELY (:266-267), one Adam over both.

B200 mapping: all three networks run on the library's kernels (tcgen05 GEMMs, LayerNorm, per-block attention in the encoder,
full-sequence attention ``csrc/attn_long.cu`` in the teacher / decoder, ``vqcpc_softmax_ce``); the three parameter groups live
in three flat buffers with their own fused clip + Adam (Adam is per-parameter, so one reference optimizer over decoder + encoder
equals two with the same hyper-parameters).  Batch-dict contract: ``tensor_dict['x']`` (batch, num_events, num_channels) ints.
torch supplies only data movement here (concatenations, slices, the mask arithmetic on the integer tokens)."""
import os
from itertools import islice

import numpy as np
import torch

from . import ops
from .encoder import EncoderTrainer, save_state_dict
from .optim import FlatParams, FusedClipAdam
from .utils import categorical_crossentropy, distilled_categorical_crossentropy, flatten, unflatten


class StudentEncoderTrainer(EncoderTrainer):
    def __init__(self, model_dir, dataloader_generator, encoder, num_events_masked, teacher, auxiliary_decoder,
                 quantization_weighting, num_gpus=1):
        super().__init__(dataloader_generator=dataloader_generator)
        self.model_dir = model_dir
        self.encoder = encoder
        self.teacher = teacher
        self.auxiliary_decoder = auxiliary_decoder
        self.codebook_dim = self.encoder.quantizer.codebook_dim
        self.upscale_factors = self.auxiliary_decoder.upscale_factors
        self.num_tokens_per_channel = self.encoder.data_processor.num_tokens_per_channel
        self.num_channels = len(self.num_tokens_per_channel)
        self.num_events_masked = num_events_masked
        assert self.encoder.data_processor.num_tokens % np.prod(self.upscale_factors) == 0
        self.quantization_weighting = quantization_weighting
        self.optimizer_enc = self.optimizer_dec = self.optimizer_teacher = None
        self.scheduler_enc = self.scheduler_dec = self.scheduler_teacher = None

    # ---- optimizers (:51-77) ------------------------------------------------------------------------------------------
    def init_optimizers(self, lr, schedule_lr):
        self.optimizer_dec = FusedClipAdam(FlatParams(list(self.auxiliary_decoder.parameters())), lr=lr)
        self.optimizer_enc = FusedClipAdam(FlatParams(list(self.encoder.parameters())), lr=lr)
        self.optimizer_teacher = FusedClipAdam(FlatParams(list(self.teacher.parameters())), lr=lr)
        for o in (self.optimizer_dec, self.optimizer_enc, self.optimizer_teacher):
            o.distributed = False
        ops.bump_weight_version()
        if schedule_lr:
            warmup_steps, lo, hi = 10000, 0.1, 1.0
            s1 = (hi - lo) / warmup_steps
            s2 = -s1 * 0.1
            lam = lambda step: max(min(lo + s1 * step, hi + (step - warmup_steps) * s2), lo)     # noqa: E731
            self.scheduler_dec = torch.optim.lr_scheduler.LambdaLR(self.optimizer_dec, lr_lambda=lam)
            self.scheduler_enc = torch.optim.lr_scheduler.LambdaLR(self.optimizer_enc, lr_lambda=lam)
            self.scheduler_teacher = torch.optim.lr_scheduler.LambdaLR(self.optimizer_teacher, lr_lambda=lam)

    def to(self, device):
        self.auxiliary_decoder.to(device)
        self.encoder.to(device)
        self.teacher.to(device)

    def _dir(self, early_stopped):
        return f'{self.model_dir}/early_stopped' if early_stopped else f'{self.model_dir}/overfitted'

    def save(self, early_stopped):
        d = self._dir(early_stopped)
        os.makedirs(d, exist_ok=True)
        self.encoder.save(early_stopped=early_stopped)
        save_state_dict(self.auxiliary_decoder, f'{d}/decoder')
        save_state_dict(self.teacher, f'{d}/teacher')

    def load(self, early_stopped, device):
        d = self._dir(early_stopped)
        self.encoder.load(early_stopped=early_stopped, device=device)
        self.auxiliary_decoder.load_state_dict(torch.load(f'{d}/decoder', map_location=torch.device(device)))
        self.teacher.load_state_dict(torch.load(f'{d}/teacher', map_location=torch.device(device)))
        ops.bump_weight_version()

    def train(self, mode=True):
        self.encoder.train(mode)
        self.auxiliary_decoder.train(mode)
        self.teacher.train(mode)
        return self

    def eval(self):
        return self.train(False)

    # ---- teacher (:120-183) ---------------------------------------------------------------------------------------------
    def mask_teacher(self, x, num_events_masked, masked_event_index=None):
        """x (batch, num_events, num_channels) -> (masked_x, notes_to_be_predicted): ONE event index for the whole batch
        (drawn with ``torch.randint`` on the host like the reference, or given), ``num_events_masked`` events on either side
        replaced by the per-channel mask token (= num_tokens_of_channel)."""
        inp = flatten(x)
        batch_size, sequence_length = inp.shape
        assert sequence_length % self.num_channels == 0
        num_events = sequence_length // self.num_channels
        if masked_event_index is None:
            masked_event_index = torch.randint(high=num_events, size=()).item()
        c = self.num_channels
        notes_to_be_predicted = torch.zeros_like(inp)
        notes_to_be_predicted[:, masked_event_index * c:(masked_event_index + 1) * c] = 1
        mask_tokens = torch.as_tensor(self.num_tokens_per_channel, dtype=inp.dtype, device=inp.device)
        mask_tokens = mask_tokens.unsqueeze(0).repeat(batch_size, num_events)
        notes_to_mask = torch.zeros_like(inp)
        notes_to_mask[:, max((masked_event_index - num_events_masked) * c, 0):(masked_event_index + num_events_masked + 1) * c] = 1
        masked_input = inp * (1 - notes_to_mask) + mask_tokens * notes_to_mask
        return unflatten(masked_input, c), unflatten(notes_to_be_predicted, c)

    def forward_teacher(self, x, masked_event_index=None):
        target = x
        masked_x, notes_to_be_predicted = self.mask_teacher(x, self.num_events_masked, masked_event_index)
        weights_per_category = self.teacher(self.teacher.data_processor.embed(masked_x))
        loss = categorical_crossentropy(value=weights_per_category, target=target, mask=notes_to_be_predicted).mean()
        return {'loss': loss, 'notes_to_be_predicted': notes_to_be_predicted, 'weights_per_category': weights_per_category,
                'monitored_quantities': {'loss_teacher': loss.item()}}

    # ---- encoder + auxiliary decoder (:185-220) ---------------------------------------------------------------------------
    def forward_encdec(self, x, weights_per_category_teacher, notes_to_be_predicted):
        weights_per_category_teacher = [t.detach() for t in weights_per_category_teacher]
        z_quantized, encoding_indices, quantization_loss = self.encoder(x)
        weights_per_category = self.auxiliary_decoder(z_quantized)
        reconstruct_loss = distilled_categorical_crossentropy(value=weights_per_category, target=weights_per_category_teacher,
                                                              mask=notes_to_be_predicted)
        loss = self.quantization_weighting * quantization_loss.mean() + reconstruct_loss.mean()
        return {'loss': loss,
                'monitored_quantities': {'loss_quantization': quantization_loss.mean().item(),
                                         'loss_reconstruction': reconstruct_loss.mean().item(),
                                         'loss_encdec': loss.item(),
                                         'loss_monitor': reconstruct_loss.mean().item()}}

    # ---- one batch (:237-276) -------------------------------------------------------------------------------------------
    def step(self, tensor_dict, train=True, masked_event_index=None):
        x = self.teacher.data_processor.preprocess(tensor_dict['x'])
        with torch.set_grad_enabled(train):
            self.optimizer_teacher.zero_grad()
            fwd_t = self.forward_teacher(x, masked_event_index)
            if train:
                fwd_t['loss'].backward()
                self.optimizer_teacher.step(max_norm=5.0)
                if self.scheduler_teacher is not None:
                    self.scheduler_teacher.step()
            self.optimizer_dec.zero_grad()
            self.optimizer_enc.zero_grad()
            fwd_e = self.forward_encdec(x, [w.detach() for w in fwd_t['weights_per_category']], fwd_t['notes_to_be_predicted'])
            if train:
                fwd_e['loss'].backward()
                self.optimizer_dec.step(max_norm=5.0)        # clip_grad_norm_(auxiliary_decoder.parameters(), 5)  :266
                self.optimizer_enc.step(max_norm=5.0)        # clip_grad_norm_(encoder.parameters(), 5)            :267
                for s in (self.scheduler_dec, self.scheduler_enc):
                    if s is not None:
                        s.step()
                ops.bump_weight_version()
        return dict(fwd_t['monitored_quantities'], **fwd_e['monitored_quantities'])

    def epoch(self, data_loader, train=True, num_batches=None, corrupt_labels=False):
        means = None
        self.train(train)
        n = 0
        for tensor_dict in islice(data_loader, num_batches):
            q = self.step(tensor_dict, train=train)
            means = q if means is None else {k: v + means[k] for k, v in q.items()}
            n += 1
        return {k: v / max(n, 1) for k, v in (means or {}).items()}
