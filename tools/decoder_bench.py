"""Autoregressive generation over code sequences at BASELINE config 5 shapes (seq 1024 = 256 events x 4 voices from 64 codes,
decoder_relative_AC_AC_C_*: d = 512, 8 heads, 3 + 3 layers, FFN 1024, positional embedding 8, embedding 32).

    python tools/decoder_bench.py [--batch 1] [--reps 3] [--full-tokens 32] [--cpu-forwards 2] [--cross anticausal|full|diagonal]
                                  [--pdl -1|0|1] [--attn-split 0|1|2|4|8] [--long-codes N]

Prints one JSON line:
  * tokens/s of ``Decoder.generate_from_codes`` on the incremental path (key / value cache + CUDA graph per event + device
    sampling; device-timed, whole window of 1024 tokens per sequence, top-p sampling);
  * the reference's schedule on the same GPU (one full teacher-forced forward per token, ``kv_cache=False``) measured over
    ``--full-tokens`` token steps;
  * the CPU oracle (oracle/decoder_oracle.py = the reference's forward restated) on the host cores: seconds per full forward =
    the reference's cost per generated token (decoder.py:611-614);
  * parity at this size: greedy tokens of the cached path against the full-recompute path over the first ``--full-tokens``
    positions (fp32 mode), and teacher-forced logits of the product forward against the oracle forward;
  * --long-codes N: sliding-window decoding of N codes (generate_from_code_long), cached vs the window recompute per token."""
import argparse
import json
import os
import sys
import time
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'vqcpc-bach_b200')):
    sys.path.insert(0, p)
import torch  # noqa: E402

VOCAB = [64, 64, 64, 64]
CFG = dict(d_model=512, n_head=8, dim_feedforward=1024, num_encoder_layers=3, num_decoder_layers=3, num_events=256, vocab=VOCAB,
           codebook_size=16, encoder_attention_type='anticausal', cross_attention_type='anticausal')


class _StubEncoder(torch.nn.Module):
    def __init__(self, codebook_size):
        super().__init__()
        self.downscaler = types.SimpleNamespace(downscale_factors=[4, 4])
        self.quantizer = types.SimpleNamespace(codebook_size=codebook_size, num_codebooks=1, codebook_dim=3)


def build(c):
    from VQCPCB.data_processor.bach_data_processor import BachDataProcessor
    from VQCPCB.decoders.decoder import Decoder
    C = len(c['vocab'])
    dp = BachDataProcessor(32, c['num_events'], list(c['vocab']))
    S = c['num_events'] * C // 16
    dec = Decoder(model_dir='/tmp/vqcpc_b200_decoder_bench', dataloader_generator=None, data_processor=dp,
                  encoder=_StubEncoder(c['codebook_size']), transformer_type='relative',
                  encoder_attention_type=c['encoder_attention_type'], cross_attention_type=c['cross_attention_type'],
                  d_model=c['d_model'], num_encoder_layers=c['num_encoder_layers'], num_decoder_layers=c['num_decoder_layers'],
                  n_head=c['n_head'], dim_feedforward=c['dim_feedforward'], positional_embedding_size=8, num_channels_encoder=1,
                  num_events_encoder=S, num_channels_decoder=C, num_events_decoder=c['num_events'], dropout=0.0)
    return dec.to('cuda').eval(), S


def timed(fn, reps):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--batch', type=int, default=1)
    ap.add_argument('--reps', type=int, default=3)
    ap.add_argument('--full-tokens', type=int, default=32)
    ap.add_argument('--cpu-forwards', type=int, default=2)
    ap.add_argument('--attn-split', type=int, default=0, help='keys split over a cluster of this many CTAs (0 = product heuristic)')
    ap.add_argument('--long-codes', type=int, default=0, help='also time generate_from_code_long over this many codes')
    ap.add_argument('--cross', default='anticausal', choices=['anticausal', 'full', 'diagonal'])
    ap.add_argument('--pdl', type=int, default=-1, help='programmatic dependent launch between the step kernels: 0 / 1 (A/B), -1 = the product heuristic')
    a = ap.parse_args()
    from VQCPCB import ops, _native
    from oracle import decoder_oracle as D
    if a.pdl >= 0:
        os.environ['VQCPC_DECODE_PDL'] = str(a.pdl)
    CFG['cross_attention_type'] = a.cross
    _native.call('vqcpc_attn_decode_set_split', a.attn_split)
    dec, S = build(CFG)
    shapes = {k: tuple(v.shape) for k, v in dec.state_dict().items()}
    prm = D.make_decoder_params(shapes, seed=0)
    dec.load_state_dict({k: v.cuda() for k, v in prm.items()}, strict=True)
    ops.bump_weight_version()
    B, T = a.batch, CFG['num_events'] * 4
    g = torch.Generator().manual_seed(3)
    codes = torch.randint(0, CFG['codebook_size'], (B, S), generator=g).cuda()
    gg = torch.Generator(device='cuda').manual_seed(11)

    n0 = _native.launch_count
    ms_cached, toks = timed(lambda: dec.generate_from_codes(codes, temperature=1.0, top_p=0.9, generator=gg), a.reps)
    launches_host = (_native.launch_count - n0) // (a.reps + 1)
    ms_eager, _ = timed(lambda: dec.generate_from_codes(codes, temperature=1.0, top_p=0.9, generator=gg, use_graph=False), 1)

    # the reference's schedule on the GPU: one full forward per token
    x = toks.clone()
    ms_fwd = {}
    for mode in ('bf16x3', 'fp32'):
        ops.set_gemm_mode(mode)
        ms_fwd[mode], fp = timed(lambda: dec.forward(codes, x), 3)
    # parity at full size: greedy, cached vs full recompute over the first tokens (fp32 arithmetic in both)
    ops.set_gemm_mode('fp32')
    cached = dec.generate_from_codes(codes, greedy=True)
    xg = torch.zeros_like(cached)
    n_cmp, same = min(a.full_tokens, T), 0
    t0 = time.perf_counter()
    for i in range(n_cmp):
        e, c = divmod(i, 4)
        lg = dec.forward(codes, xg)['weights_per_category'][c][:, e, :]
        xg[:, e, c] = lg.argmax(dim=-1)
        same += int((xg[:, e, c] == cached[:, e, c]).sum())
    torch.cuda.synchronize()
    ms_full_token = (time.perf_counter() - t0) * 1e3 / max(1, n_cmp)
    # teacher-forced logits vs the CPU oracle on the same window (also the CPU baseline timing)
    ocfg = dict(CFG, total_upscaling=16)
    xc, cc = cached[:1].cpu(), codes[:1].cpu()
    torch.set_num_threads(os.cpu_count())
    cpu_s = []
    if a.cpu_forwards <= 0:
        print(json.dumps(dict(value=B * T / ms_cached * 1e3, ms_per_window=ms_cached, us_per_token_step=ms_cached / T * 1e3,
                              ms_per_window_no_graph=ms_eager)))
        return
    for _ in range(a.cpu_forwards):
        t0 = time.perf_counter()
        with torch.no_grad():
            _, lo = D.decoder_forward(cc, xc, prm, ocfg)
        cpu_s.append(time.perf_counter() - t0)
    errs = {}
    for mode in ('fp32', 'bf16x3'):
        ops.set_gemm_mode(mode)
        lp = dec.forward(codes[:1], cached[:1])['weights_per_category']
        errs[mode] = max(float((p_.cpu() - o_).norm() / o_.norm()) for p_, o_ in zip(lp, lo))
    ops.set_gemm_mode('bf16x3')
    cpu_fwd = min(cpu_s)
    out = dict(metric='decoder_generation_tokens_per_sec', value=B * T / ms_cached * 1e3, unit='tokens/s', batch=B, tokens_per_sequence=T,
               ms_per_window=ms_cached, us_per_token_step=ms_cached / T * 1e3, ms_per_window_no_graph=ms_eager,
               kernel_launches_per_window=dec.last_generation_launches, pdl=a.pdl, host_calls_per_window=launches_host,
               full_recompute=dict(ms_per_forward=ms_fwd, ms_per_token_measured=ms_full_token,
                                   tokens_per_sec=B / max(ms_full_token, 1e-9) * 1e3, tokens_timed=n_cmp),
               cpu_baseline=dict(kind='port', cores=torch.get_num_threads(), s_per_forward=cpu_fwd,
                                 value=1.0 / cpu_fwd, unit='tokens/s',
                                 sample=f'{len(cpu_s)} full forwards at batch 1 (the reference runs one per generated token)'),
               parity=dict(greedy_tokens_equal=same, greedy_tokens_compared=n_cmp * B, logits_rel_err_vs_oracle=errs),
               config=dict(cross_attention_type=a.cross, workload='decoder_relative generation, 1024 tokens (256 events x 4 voices) from 64 codes, d=512, '
                                    '8 heads, 3+3 layers, FFN 1024, top-p 0.9', data='synthetic codes, random-init weights'))
    if a.long_codes > 0:
        # sliding-window decoding of a code sequence longer than the window (decoder.py:729-829)
        lc = torch.randint(0, CFG['codebook_size'], (B, a.long_codes), generator=g).cuda()
        dec.generate_from_code_long(lc[:, :S + 2], temperature=1.0, top_p=0.9, code_index_start=0, code_index_end=2)   # warm-up
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        toks_long = dec.generate_from_code_long(lc, temperature=1.0, top_p=0.9)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        n_ref = 2
        t0 = time.perf_counter()
        dec.generate_from_code_long(lc, temperature=1.0, top_p=0.9, code_index_start=S, code_index_end=S + n_ref, kv_cache=False)
        torch.cuda.synchronize()
        dt_ref = (time.perf_counter() - t0) / (n_ref * 16)
        out['long_generation'] = dict(codes=a.long_codes, tokens_per_sequence=toks_long.shape[1] * 4, seconds=dt,
                                      tokens_per_sec=toks_long.numel() / dt, ms_per_code=dt / a.long_codes * 1e3,
                                      window_recompute=dict(ms_per_token=dt_ref * 1e3, tokens_per_sec=B / dt_ref, codes_timed=n_ref))
    print(json.dumps(out))


if __name__ == '__main__':
    main()
