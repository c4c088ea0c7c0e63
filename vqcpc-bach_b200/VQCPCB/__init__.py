"""vqcpc_b200: B200-native VQ-CPC encoder-training hot path behind the reference's ``VQCPCB`` class surface."""
import torch as _torch

# The GRUs run on the native kernels (ops.GRUFn); their nn.GRU modules are parameter containers only.  Should a caller
# ever invoke one directly, keep cuDNN away from TF32: it costs ~1e-3 relative error in the context vector and the parity
# bar is 1e-4.  (Our own GEMMs never use TF32: fp32 FMA or the bf16x3 split.)
_torch.backends.cudnn.allow_tf32 = False
