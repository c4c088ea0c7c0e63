"""VQ-CPC encoder trainer on the B200 path.

Same class surface as VQCPCB/vqcpc_encoder_trainer.py:14-354 (constructor, ``init_optimizers``, ``to``, ``save``,
``load``, ``train``, ``eval``, ``epoch`` -> dict with the reference's keys), re-designed around the hardware:

* all parameters / gradients / Adam moments live in FLAT fp32 buffers; one fused kernel pair does the global-norm
  clip (5.0) + Adam update with no host sync (reference: clip_grad_norm_ + optimizer.step, :313-314);
* a batch is processed as independent micro-batches of samples (forward + backward each, gradients accumulate in
  the flat buffer) — samples only couple through the 1/B loss normalisers, so this is exact and bounds the
  activation memory at batch 4096 (SURVEY.md §7 'Activation memory');
* data parallel: one process per GPU, ONE NCCL all-reduce per step over [gradients | code-usage counts | loss and
  accuracy sums] (``distributed.py``), then the identical fused update on every rank;
* monitoring scalars stay on the device and are read back once per epoch (the reference syncs >= 6x per step).
"""
import math
import os
from itertools import islice

import torch
from torch import nn

from . import ops, vqcpc_helper
from .encoder import EncoderTrainer, save_state_dict
from .optim import FlatParams, FusedClipAdam
from . import distributed as vq_dist


class VQCPCEncoderTrainer(EncoderTrainer):
    def __init__(self, model_dir, dataloader_generator, encoder, c_net_kwargs, quantization_weighting):
        super().__init__(dataloader_generator=dataloader_generator)
        self.model_dir = model_dir
        self.dataloader_generator = dataloader_generator
        self.encoder = encoder
        self.data_processor = encoder.data_processor
        self.codebook_dim = self.encoder.quantizer.codebook_dim
        self.upscale_factors = list(reversed(self.encoder.downscaler.downscale_factors))
        self.num_tokens_per_channel = self.encoder.data_processor.num_tokens_per_channel
        self.num_channels = len(self.num_tokens_per_channel)
        assert self.data_processor.num_tokens % math.prod(self.upscale_factors) == 0

        z_dim = encoder.upscaler.output_dim if encoder.upscaler is not None else self.codebook_dim
        c_dim = c_net_kwargs['output_dim']
        k_max = self.dataloader_generator.num_blocks_right

        def c_net():
            return vqcpc_helper.CModule(input_dim=z_dim, hidden_size=c_net_kwargs['hidden_size'], output_dim=c_dim,
                                        num_layers=c_net_kwargs['num_layers'], dropout=c_net_kwargs['dropout'])

        self.c_module = c_net()
        self.fks_module = vqcpc_helper.FksModule(z_dim=z_dim, c_dim=c_dim, k_max=k_max)
        # 7 of the shipped configs omit 'bidirectional' (SURVEY.md fact 3): default to the unidirectional model
        if c_net_kwargs.get('bidirectional', False):
            self.c_module_back = c_net()
            self.fks_module_back = vqcpc_helper.FksModule(z_dim=z_dim, c_dim=c_dim, k_max=k_max)
        else:
            self.c_module_back = None

        self.quantization_weighting = quantization_weighting
        self.optimizer = None
        self.scheduler = None
        self.flat = None
        self.max_grad_norm = 5.0
        # micro-batching: at most this many 16-token blocks per forward/backward pass (256 samples of the headline config,
        # ~30 GB of saved activations; measured on B200: 17408 -> 2951, 34816 -> 3084, 69632 -> 3083, 139264 -> 3049
        # samples/s - larger passes amortise the launch-bound GRU steps, the largest run into the power cap)
        self.max_blocks_per_pass = int(os.environ.get('VQCPC_MAX_BLOCKS_PER_PASS', 34816))
        self.is_main_process = vq_dist.rank() == 0
        self.last_grad_norm_sq = None
        self.distributed = True     # False: ignore an initialised process group (single-process reference runs)
        # index-gather quantizer -> InfoNCE path (csrc/idxpath.cu) whenever the bottleneck is in table mode; False (or
        # VQCPC_INDEX_PATH=0) keeps the per-vector kernels (materialised upscaled rows)
        self.index_path = os.environ.get('VQCPC_INDEX_PATH', '1') != '0'

    # ------------------------------------------------------------------------------------------ plumbing
    def _modules_in_order(self):
        mods = [self.c_module, self.fks_module, self.encoder]
        if self.c_module_back is not None:
            mods += [self.fks_module_back, self.c_module_back]
        return mods

    def init_optimizers(self, lr, schedule_lr):
        """Adam(lr) over c/fks/encoder(+back) parameters (:85-92) as a fused flat-buffer optimizer; optional
        LambdaLR warm-up/decay (:95-107)."""
        params = []
        for m in self._modules_in_order():
            params += list(m.parameters())
        if self.distributed and vq_dist.world_size() > 1:
            vq_dist.broadcast_parameters(params)          # replicas must start identical
        # (the data-dependent codebook initialisation happens later, inside the first forward: the quantizer broadcasts
        #  rank 0's draw itself - vector_quantizer.py `_initialize`)
        if hasattr(self.encoder.quantizer, 'sync_init'):
            self.encoder.quantizer.sync_init = bool(self.distributed)
        self.flat = FlatParams(params)
        self.optimizer = FusedClipAdam(self.flat, lr=lr, extra_tail=self._tail_size())
        self.optimizer.distributed = self.distributed
        ops.bump_weight_version()
        if schedule_lr:
            warmup_steps, lo, hi = 10000, 0.1, 1.0
            s1 = (hi - lo) / warmup_steps
            s2 = -s1 * 0.1
            self.scheduler = torch.optim.lr_scheduler.LambdaLR(
                self.optimizer, lr_lambda=lambda step: max(min(lo + s1 * step, hi + (step - warmup_steps) * s2), lo))

    def _tail_size(self):
        """[loss sums (3) | accuracy counts (KR) | code-usage histograms of the positive and the negative blocks]"""
        q = self.encoder.quantizer
        kc = getattr(q, 'codebook_size', 0)
        ema = kc * (1 + q.codebook_dim) if getattr(q, 'ema_decay', None) is not None else 0     # EMA sums + counts
        return FusedClipAdam.tail_elems(3 + self.dataloader_generator.num_blocks_right + ema, 2 * kc + ema) + 4

    def to(self, device):
        for m in self._modules_in_order():
            m.to(device)

    def _dir(self, early_stopped):
        return f'{self.model_dir}/early_stopped' if early_stopped else f'{self.model_dir}/overfitted'

    def save(self, early_stopped):
        d = self._dir(early_stopped)
        os.makedirs(d, exist_ok=True)
        self.encoder.save(early_stopped=early_stopped)
        save_state_dict(self.c_module, f'{d}/c_module')
        save_state_dict(self.fks_module, f'{d}/fks_module')
        if self.c_module_back is not None:
            save_state_dict(self.c_module_back, f'{d}/c_module_back')
            save_state_dict(self.fks_module_back, f'{d}/fks_module_back')

    def load(self, early_stopped, device):
        print(f'Loading models {self.__repr__()}')
        d = self._dir(early_stopped)
        if not os.path.exists(d):
            d = self.model_dir
        loc = torch.device(device)
        self.encoder.load(early_stopped=early_stopped, device=device)
        self.c_module.load_state_dict(torch.load(f'{d}/c_module', map_location=loc))
        self.fks_module.load_state_dict(torch.load(f'{d}/fks_module', map_location=loc))
        if self.c_module_back is not None:
            self.c_module_back.load_state_dict(torch.load(f'{d}/c_module_back', map_location=loc))
            self.fks_module_back.load_state_dict(torch.load(f'{d}/fks_module_back', map_location=loc))
        ops.bump_weight_version()

    def train(self):
        for m in self._modules_in_order():
            m.train()

    def eval(self):
        for m in self._modules_in_order():
            m.eval()

    # ------------------------------------------------------------------------------------------ one step
    def _samples_per_pass(self, B, blocks_per_sample):
        if self.c_module_back is not None:
            return B          # the reference's backward-direction negative indexing couples the whole batch (quirk Q1)
        n = max(1, min(B, self.max_blocks_per_pass // max(1, blocks_per_sample)))
        while B % n:
            n -= 1
        return n

    def step(self, tensor_dict, train, corrupt_labels=False):
        """Forward (+ backward + fused clip/Adam when ``train``) on one batch dict.  Returns device tensors:
        loss, loss_quantize, loss_contrastive (scalars), accuracy (KR), num_codewords, num_codewords_negative."""
        enc = self.encoder
        neg_all = tensor_dict['negative_samples']
        B, N, KR, T, V = neg_all.shape
        x_left_all, x_right_all = tensor_dict['x_left'], tensor_dict['x_right']
        KL = x_left_all.shape[1] * V // (T * V)
        bidir = self.c_module_back is not None
        neg_back_all = tensor_dict['negative_samples_back'] if bidir else None
        dev = self.fks_module.W.device
        n_terms = 4 if bidir else 3
        w = self.quantization_weighting

        if self.optimizer is not None:
            self.optimizer.zero_grad()
        sums = torch.zeros(3, dtype=torch.float32, device=dev)            # loss, qloss, nce (already normalised)
        correct = torch.zeros(KR, dtype=torch.float32, device=dev)
        q = enc.quantizer
        Kc = getattr(q, 'codebook_size', None)
        # code-usage histograms of this step: row 0 = positive blocks (left + right), row 1 = negatives (the reference
        # counts torch.unique of the same two sets, :325-331).  Under data parallelism they ride in the step's all-reduce.
        hist = torch.zeros(2, Kc, dtype=torch.int32, device=dev) if Kc else None

        bs = self._samples_per_pass(B, KL + KR + N * KR * (2 if bidir else 1))
        for b0 in range(0, B, bs):
            sl = slice(b0, b0 + bs)
            neg = neg_all[sl].reshape(bs * N * KR, T, V)
            merged = (not corrupt_labels) and not getattr(q, 'initialize', False)
            use_idx = (merged and self.index_path and not bidir and enc._fusable(False)
                       and ops.idxpath_supported(N, KR, enc.upscaler.output_dim, self.fks_module.W.shape[1], Kc, q.codebook_dim))
            if use_idx:
                # table mode: one encoder pass over [negatives | left | right]; only the left blocks get their upscaled
                # rows (the context GRU reads them), right / negative blocks are scored from their code indices
                dp = enc.data_processor
                toks = [dp.preprocess(neg), dp.preprocess(x_left_all[sl]), dp.preprocess(x_right_all[sl])]
                S_tok = toks[0].shape[-1]
                n_neg, n_left, n_right = bs * N * KR, bs * KL, bs * KR
                allt = torch.cat([t.reshape(-1, S_tok) for t in toks], 0).unsqueeze(0)
                y_left, idx_flat, ql_flat, link = enc.forward_tokens_indexed(allt, n_neg, n_left)
                zd = y_left.shape[-1]
                z_left = y_left.view(bs, KL, zd)
                idx_neg, idx_left, idx_right = idx_flat.split([n_neg, n_left, n_right])
                ql_neg, ql_left, ql_right = ql_flat.split([n_neg, n_left, n_right])
                c = self.c_module(z_left, None)                                            # :237
                nll_sum, pos, negs, corr = self.fks_module.fused_nce_indexed(              # :240-272
                    c, link, idx_right.view(bs, KR), idx_neg.view(bs, N, KR), n_neg + n_left, 0)
                z_neg = z_right = None
            elif merged:
                # blocks are independent, so the reference's separate Encoder.forward calls (:201, :219, :227, :230) are
                # one pass over the concatenated blocks (bigger GEMMs, 3-4x fewer launches); results are identical.
                dp = enc.data_processor
                toks = [dp.preprocess(neg), dp.preprocess(x_left_all[sl]), dp.preprocess(x_right_all[sl])]
                if bidir:
                    toks.append(dp.preprocess(neg_back_all[sl].reshape(bs * N * KR, T, V)))
                S_tok = toks[0].shape[-1]
                counts = [t.numel() // S_tok for t in toks]
                allt = torch.cat([t.reshape(-1, S_tok) for t in toks], 0).unsqueeze(0)
                z_all, idx_all, ql_all = enc.forward_tokens(allt)
                zd = z_all.shape[-1]
                z_parts = z_all[0].split(counts, 0)
                ql_parts = ql_all[0].split(counts, 0)
                idx_parts = idx_all[0].split(counts, 0) if idx_all is not None else [None] * len(counts)
                z_neg, idx_neg, ql_neg = z_parts[0].reshape(bs, N, KR, zd), idx_parts[0], ql_parts[0]
                z_left, idx_left, ql_left = z_parts[1].reshape(bs, KL, zd), idx_parts[1], ql_parts[1]
                z_right, idx_right, ql_right = z_parts[2].reshape(bs, KR, zd), idx_parts[2], ql_parts[2]
                if bidir:
                    z_nb, idx_nb, ql_nb = z_parts[3].reshape(bs * N * KR, 1, zd), idx_parts[3], ql_parts[3]
            else:
                z_neg, idx_neg, ql_neg = enc(neg, corrupt_labels=corrupt_labels)          # :197-201
                zd = z_neg.shape[-1]
                z_neg = z_neg.view(bs, N, KR, zd)                                          # nb == 1
                if bidir:
                    nb_ = neg_back_all[sl].reshape(bs * N * KR, T, V)
                    z_nb, idx_nb, ql_nb = enc(nb_, corrupt_labels=corrupt_labels)          # :211-225
                z_left, idx_left, ql_left = enc(x_left_all[sl], corrupt_labels=False)      # :227
                z_right, idx_right, ql_right = enc(x_right_all[sl], corrupt_labels=False)  # :230

            if not use_idx:
                c = self.c_module(z_left, None)                                            # :237
                nll_sum, pos, negs, corr = self.fks_module.fused_nce(c, z_right, z_neg)    # :240-272
            nce = nll_sum / B
            corr_f = corr.float()
            if bidir:
                c_back = self.c_module_back(z_right.flip(dims=[1]), None)                  # :278-279
                # quirk Q1: the reference views the (B, N, KR, z) backward negatives b-major while
                # repeating c n-major, i.e. row n*B+b scores c_back[b] against flat vector n*B+b.  Reproduced as is.
                z_nb_eff = z_nb.view(bs * N, KR, zd).view(N, bs, KR, zd).permute(1, 0, 2, 3).contiguous()
                nll_b, _, _, corr_b = self.fks_module_back.fused_nce(c_back, z_left, z_nb_eff)   # :283-296
                nce = nce + nll_b / B
                corr_f = (corr_f + corr_b.float()) / 2
            ql_sum = ql_left.sum() + ql_right.sum() + ql_neg.sum()
            if bidir:
                ql_sum = ql_sum + ql_nb.sum()
            qloss = ql_sum / (n_terms * B)                                                 # vqcpc_helper.py:32-51
            loss = nce + w * qloss                                                         # :307
            if train:
                loss.backward()
            with torch.no_grad():
                sums += torch.stack([loss.detach(), qloss.detach(), nce.detach()])
                correct += corr_f.sum(0)
                if idx_left is not None:
                    ops.code_hist(idx_left, hist[0])
                    ops.code_hist(idx_right, hist[0])
                    ops.code_hist(idx_neg, hist[1])
            del loss, nce, qloss, z_neg, z_left, z_right, c

        if train:
            # :313-314 (+ the one all-reduce: gradients | loss/accuracy sums | code counts)
            ema_sum, ema_cnt = q.ema_step_tensors() if hasattr(q, 'ema_step_tensors') else (None, None)
            if ema_cnt is not None:
                for e in q.embeddings:                  # EMA codebook: no gradient step on the codes
                    if e.grad is not None:
                        e.grad.zero_()
            counts = ([hist] if hist is not None else []) + ([ema_cnt] if ema_cnt is not None else [])
            self.optimizer.step(max_norm=self.max_grad_norm, stats=(sums, correct), counts=counts or None, sums=[ema_sum])
            if ema_cnt is not None:
                q.ema_update()
            ops.bump_weight_version()
            if hist is not None:
                if q.global_usage is None or q.global_usage.device != hist.device:
                    q.global_usage = torch.zeros(Kc, dtype=torch.int64, device=dev)
                q.global_usage += hist.sum(0)
            if self.scheduler is not None:
                self.scheduler.step()                                                      # :315-316
        out = dict(loss=sums[0], loss_quantize=sums[1], loss_contrastive=sums[2], accuracy=correct / B)
        if hist is not None:
            out['num_codewords'] = (hist[0] > 0).sum()
            out['num_codewords_negative'] = (hist[1] > 0).sum()
        return out

    # ------------------------------------------------------------------------------------------ epoch
    def epoch(self, data_loader, train, num_batches, corrupt_labels):
        if self.is_main_process:
            print(f'lr: {self.optimizer.param_groups[0]["lr"]}')
        if train:
            self.train()
        else:
            self.eval()
        acc = None
        n_done = 0
        for tensor_dict in islice(data_loader, num_batches):
            if train:
                st = self.step(tensor_dict, True, corrupt_labels)
            else:
                with torch.no_grad():          # the reference builds graphs in eval too (quirk Q6); not needed
                    st = self.step(tensor_dict, False, corrupt_labels)
            n_done += 1
            if acc is None:
                acc = {k: v.clone().float() for k, v in st.items()}
            else:
                for k, v in st.items():
                    acc[k] += v
        if acc is None:
            raise RuntimeError('epoch() received an empty data loader')
        # single device->host read per epoch
        host = {k: (v / n_done).cpu() for k, v in acc.items()}
        means = {
            'loss': host['loss'].item(),
            'accuracy': [float(a) for a in host['accuracy'].tolist()],
            'loss_quantize': host['loss_quantize'].item(),
            'loss_contrastive': host['loss_contrastive'].item(),
            'num_codewords': host['num_codewords'].item() if 'num_codewords' in host else 0,
            'num_codewords_negative': host['num_codewords_negative'].item() if 'num_codewords_negative' in host else 0,
        }
        means['loss_monitor'] = -sum(means['accuracy']) / len(means['accuracy'])
        return means
