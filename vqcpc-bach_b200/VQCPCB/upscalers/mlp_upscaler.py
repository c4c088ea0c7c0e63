"""MLP upscaler Linear -> Dropout -> SELU -> Linear (VQCPCB/upscalers/mlp_upscaler.py:4-35) on the native GEMM/SELU
kernels.  ``self.mlp`` keeps the reference's nn.Sequential so the state_dict keys stay ``mlp.0.*`` / ``mlp.3.*``."""
from torch import nn

from .. import ops


class MlpUpscaler(nn.Module):
    def __init__(self, input_dim, output_dim, hidden_size, dropout):
        super().__init__()
        self.input_dim, self.output_dim, self.dropout = input_dim, output_dim, dropout
        self.mlp = nn.Sequential(nn.Linear(input_dim, hidden_size, bias=True), nn.Dropout(p=dropout), nn.SELU(),
                                 nn.Linear(hidden_size, output_dim, bias=True))

    def forward(self, inputs, keep_mask=None):
        """keep_mask (uint8, shape of the hidden layer) replays a given dropout draw (tests); by default a fresh one is
        drawn from torch's CUDA generator when training with dropout > 0."""
        h = ops.LinearFn.apply(inputs, self.mlp[0].weight, self.mlp[0].bias)
        if self.training and self.dropout > 0:
            if keep_mask is None:
                keep_mask = ops.draw_keep_mask(h.shape, self.dropout, h.device)
            h = ops.DropoutFn.apply(h, keep_mask, self.dropout)
        h = ops.SeluFn.apply(h)
        return ops.LinearFn.apply(h, self.mlp[3].weight, self.mlp[3].bias)

    def code_table_and_jacobian(self, codes):
        """For the fused bottleneck: T = upscaler(codes) (Kc, out) WITH autograd history to the MLP weights (codes are
        treated as constants: the reference detaches q on this side, vector_quantizer.py:148), and - without history -
        the Jacobian J (Kc, out, in) of the MLP at each code vector:  J[k] = W_b diag(selu'(W_a e_k + b_a)) W_a."""
        import torch
        codes = codes.detach()
        wa, ba, wb, bb = self.mlp[0].weight, self.mlp[0].bias, self.mlp[3].weight, self.mlp[3].bias
        pre = ops.LinearFn.apply(codes, wa, ba)                      # (Kc, hidden)
        table = ops.LinearFn.apply(ops.SeluFn.apply(pre), wb, bb)    # (Kc, out)
        with torch.no_grad():
            Kc, hid = pre.shape
            n_in, n_out = wa.shape[1], wb.shape[0]
            pre_d = pre.detach().contiguous()
            jac = torch.empty(Kc, n_out, n_in, dtype=torch.float32, device=pre.device)
            tmp = torch.empty_like(pre_d)
            for d in range(n_in):
                col = wa.detach()[:, d].expand(Kc, hid).contiguous()
                ops._native.call('vqcpc_selu_bwd', ops._p(pre_d), ops._p(col), ops._p(tmp), pre_d.numel(), ops._stream())
                jac[:, :, d] = ops.linear_fwd(tmp, wb.detach(), None)   # (Kc, out) = (selu'(pre) * W_a[:, d]) W_b^T
        return table, jac
