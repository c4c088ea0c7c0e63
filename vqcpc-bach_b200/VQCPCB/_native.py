"""ctypes binding of libvqcpc_b200.so (the C-ABI declared in include/vqcpc_b200.h).

The library is built in-tree by ``tools/build.sh`` / ``__graft_entry__.build()``.  There is no fallback: if the
shared object is missing or a call fails, a RuntimeError is raised (no CPU path, no Triton, no torch eager path).
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# VQCPC_LIB_TAG selects an A/B build of the same sources (tools/build.sh with VQ_TAG=...); the product uses the untagged library
LIB_PATH = os.path.join(os.path.dirname(_HERE), 'libvqcpc_b200%s.so' % os.environ.get('VQCPC_LIB_TAG', ''))

_c = ctypes
_P, _I, _L, _F = _c.c_void_p, _c.c_int, _c.c_int64, _c.c_float

# name -> argtypes (all functions return int status except the two marked below)
SIGNATURES = {
    'vqcpc_device_info': [_P, _P, _P],
    'vqcpc_embed_fwd': [_P, _L, _I, _P, _P, _I, _P, _P],
    'vqcpc_embed_bwd': [_P, _L, _I, _P, _P, _I, _P, _P],
    'vqcpc_posemb_fill': [_P, _I, _I, _P, _P, _L, _I, _I, _I, _P],
    'vqcpc_posemb_bwd': [_P, _I, _I, _P, _P, _L, _I, _I, _I, _P],
    'vqcpc_layernorm_fwd': [_P, _P, _P, _P, _P, _P, _P, _P, _L, _I, _F, _P],
    'vqcpc_layernorm_bwd': [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _L, _I, _P],
    'vqcpc_layernorm_bwd_ex': [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _L, _I, _P],
    'vqcpc_subsample_fwd': [_P, _P, _L, _I, _I, _I, _P],
    'vqcpc_subsample_bwd': [_P, _P, _L, _I, _I, _I, _P],
    'vqcpc_colsum': [_P, _I, _P, _L, _I, _P],
    'vqcpc_colsum_split': [_P, _P, _I, _P, _L, _I, _P],
    'vqcpc_selu_fwd': [_P, _P, _L, _P],
    'vqcpc_selu_bwd': [_P, _P, _P, _L, _P],
    'vqcpc_mask_scale': [_P, _P, _F, _P, _L, _P],
    'vqcpc_mask_scale_add': [_P, _P, _F, _P, _P, _P, _P, _L, _P],
    'vqcpc_split_bf16': [_P, _I, _P, _P, _I, _L, _I, _P],
    'vqcpc_reduce_sum': [_P, _L, _P, _F, _P],
    'vqcpc_code_hist': [_P, _L, _I, _P, _P],
    'vqcpc_gemm_f32': [_I, _I, _I, _I, _I, _P, _I, _P, _I, _P, _I, _P, _P, _I, _I, _F, _I, _I, _P],
    'vqcpc_gemm_bf16x3': [_I, _I, _I, _I, _I, _P, _P, _I, _P, _P, _I, _P, _I, _P, _P, _I, _I, _F, _I, _I, _I, _P],
    'vqcpc_gemm_bf16x3_ex': [_I, _I, _I, _I, _I, _P, _P, _I, _P, _P, _I, _P, _I, _P, _P, _I, _P, _P, _I, _I, _F, _I, _I, _I, _P],
    'vqcpc_gemm_set_tile_config': [_I],
    'vqcpc_gemm_set_debug': [_P],
    'vqcpc_gemm_set_experiment': [_I],
    'vqcpc_attn_fwd': [_P, _P, _P, _P, _P, _P, _P, _L, _I, _I, _I, _I, _P],
    'vqcpc_attn_bwd': [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _L, _I, _I, _I, _F, _I, _P],
    'vqcpc_attn_long_fwd': [_P, _P, _P, _P, _P, _P, _P, _L, _I, _I, _I, _P],
    'vqcpc_attn_long_bwd': [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _L, _I, _I, _I, _F, _P],
    'vqcpc_attn_long_fwd_masked': [_P, _P, _P, _P, _P, _L, _I, _I, _I, _I, _P],
    'vqcpc_attn_map_fwd': [_P, _I, _P, _P, _I, _P, _P, _P, _P, _P, _P, _P, _L, _I, _I, _I, _I, _P],
    'vqcpc_decode_features': [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P],
    'vqcpc_linear_small': [_P, _I, _P, _P, _P, _I, _P, _I, _I, _I, _I, _I, _F, _I, _P, _P, _P],
    'vqcpc_attn_decode': [_P, _I, _P, _P, _I, _P, _P, _I, _I, _P, _P, _P, _P, _P, _I, _P, _I, _P, _I, _L, _I, _I, _P],
    'vqcpc_attn_decode_set_split': [_I],
    'vqcpc_decode_set_pdl': [_I],
    'vqcpc_decode_ln_add': [_P, _P, _P, _F, _P, _I, _P, _P, _I, _I, _P],
    'vqcpc_elu_fwd': [_P, _P, _L, _P],
    'vqcpc_ln_linear_small': [_P, _I, _P, _P, _F, _P, _P, _P, _P, _I, _P, _I, _I, _I, _I, _I, _F, _I, _P],
    'vqcpc_decode_sample_embed': [_P, _I, _I, _F, _I, _F, _P, _I, _P, _I, _P, _I, _I, _P, _P, _P, _P, _P, _I, _I, _I, _I, _P, _I, _P],
    'vqcpc_sample_step': [_P, _I, _I, _F, _I, _F, _P, _I, _P, _I, _P, _I, _I, _P],
    'vqcpc_decode_advance': [_P, _I, _P],
    'vqcpc_cpc_window_gather': [_P, _I, _L, _I, _I, _P, _L, _I, _I, _P, _P],
    'vqcpc_cpc_same_seq_negatives': [_P, _I, _L, _I, _I, _P, _L, _I, _I, _I, _I, _I, _I, _P, _P],
    'vqcpc_softmax_ce': [_P, _I, _P, _P, _I, _P, _P, _I, _F, _L, _I, _P],
    'vqcpc_attn_set_fwd_algo': [_I],
    'vqcpc_attn_set_bwd_algo': [_I],
    'vqcpc_attn_fwd_ex': [_P, _P, _P, _P, _P, _P, _P, _L, _I, _I, _I, _I, _P, _F, _P],
    'vqcpc_attn_bwd_ex': [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _L, _I, _I, _I, _F, _I, _P, _F, _P],
    'vqcpc_vq_search': [_P, _L, _I, _P, _I, _I, _P, _P],
    'vqcpc_vq_gather': [_P, _L, _I, _P, _I, _I, _P, _F, _I, _P, _P, _P, _P],
    'vqcpc_vq_bwd': [_P, _L, _I, _P, _P, _I, _I, _P, _F, _I, _P, _P, _P, _P],
    'vqcpc_vq_fused_fwd': [_P, _L, _I, _P, _P, _I, _I, _F, _I, _P, _P, _P, _P, _P, _P],
    'vqcpc_vq_fused_fwd_ex': [_P, _L, _I, _P, _P, _I, _I, _F, _I, _P, _P, _P, _P, _P, _I, _P, _P, _P],
    'vqcpc_vq_fused_bwd': [_P, _L, _I, _P, _P, _I, _I, _P, _F, _I, _P, _P, _P, _P, _P, _P],
    'vqcpc_vq_fused_fwd_rows': [_P, _L, _I, _P, _P, _I, _I, _F, _I, _P, _P, _P, _P, _L, _L, _P, _P],
    'vqcpc_infonce_idx_fwd': [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P],
    'vqcpc_infonce_idx_bwd': [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _F, _I, _I, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P],
    'vqcpc_table_scatter_add': [_P, _L, _P, _I, _I, _P, _P],
    'vqcpc_vq_idx_bwd': [_P, _L, _I, _P, _P, _I, _I, _P, _F, _I, _P, _I, _P, _P, _P, _P, _P, _P, _I, _P, _P, _P, _P],
    'vqcpc_idx_debug_read': [_P],
    'vqcpc_infonce_idx_set_bwd_algo': [_I],
    'vqcpc_vq_idx_set_bwd_algo': [_I],
    'vqcpc_infonce_fwd': [_P, _P, _P, _P, _I, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P],
    'vqcpc_infonce_set_algo': [_I],
    'vqcpc_infonce_bwd': [_P, _P, _P, _P, _P, _P, _P, _P, _P, _F, _I, _I, _I, _I, _I, _P, _P, _P, _P, _P],
    'vqcpc_gru_gate_fwd': [_P, _P, _P, _L, _P, _P, _P, _P, _P, _I, _I, _I, _I, _P],
    'vqcpc_gru_gate_bwd': [_P, _P, _P, _L, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _P],
    'vqcpc_grad_sumsq': [_P, _L, _F, _P, _P],
    'vqcpc_clip_adam': [_P, _P, _P, _P, _L, _P, _F, _F, _F, _F, _F, _F, _I, _P],
}
_SPECIAL = {'vqcpc_last_error': (_c.c_char_p, []), 'vqcpc_abi_version': (_I, []), 'vqcpc_gemm_epilogue_kind': (_I, []),
            'vqcpc_idxpath_supported': (_I, [_I, _I, _I, _I, _I, _I]), 'vqcpc_infonce_idx_parts': (_I, [_I, _I]),
            'vqcpc_infonce_idx_scratch': (_L, [_I, _I, _I])}

_lib = None
launch_count = 0   # number of kernel-launching C-ABI calls made through this module (bench.py reports it)


def load():
    """Load the shared library once; raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f'{LIB_PATH} not found: build it with tools/build.sh (or __graft_entry__.build()). '
            'vqcpc_b200 has no CPU/eager fallback.')
    lib = ctypes.CDLL(LIB_PATH)
    for name, args in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = _I
    for name, (res, args) in _SPECIAL.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = res
    _lib = lib
    return lib


def call(name, *args):
    """Invoke a status-returning entry point; raise RuntimeError with the library's message on failure."""
    global launch_count
    lib = load()
    rc = getattr(lib, name)(*args)
    launch_count += 1
    if rc != 0:
        msg = lib.vqcpc_last_error().decode(errors='replace')
        raise RuntimeError(f'{name} failed (code {rc}): {msg}')
    return rc
