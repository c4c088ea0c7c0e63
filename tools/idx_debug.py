"""Debug helper: run the index-path kernels one by one (synchronising after each) for one shape.
    python tools/idx_debug.py B N KR KL Kc D"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'vqcpc-bach_b200')):
    sys.path.insert(0, p)
import ctypes  # noqa: E402
import torch  # noqa: E402
from VQCPCB import _native, ops  # noqa: E402

B, N, KR, KL, Kc, D = [int(a) for a in sys.argv[1:7]]
zd = cd = 32
dev = 'cuda'
g = torch.Generator(device=dev).manual_seed(0)
n_neg, n_left, n_right = B * N * KR, B * KL, B * KR
P = n_neg + n_left + n_right
E = torch.randn(Kc, D, device=dev, generator=g) * 1.5
z = torch.randn(P, D, device=dev, generator=g) * 1.5
T = torch.randn(Kc, zd, device=dev, generator=g)
J = torch.randn(Kc, zd, D, device=dev, generator=g)
c = torch.randn(B, cd, device=dev, generator=g)
W = torch.randn(zd, cd, KR, device=dev, generator=g) * 0.2
p_, st = ops._p, ops._stream


def step(name, fn):
    fn()
    torch.cuda.synchronize()
    print('ok', name, flush=True)


idx = torch.empty(P, dtype=torch.int64, device=dev)
loss = torch.empty(P, device=dev)
y_left = torch.empty(n_left, zd, device=dev)
counts = torch.zeros(Kc, dtype=torch.int32, device=dev)
step('vq_fused_fwd_rows', lambda: _native.call('vqcpc_vq_fused_fwd_rows', p_(z), P, D, p_(E), p_(T), Kc, zd, 0.25, 1, p_(idx), p_(loss), None,
                                                p_(y_left), n_neg, n_left, p_(counts), st()))
assert torch.equal(y_left, T[idx[n_neg:n_neg + n_left]])
u = torch.empty(B, KR, zd, device=dev)
pos = torch.empty(B, KR, device=dev)
neg = torch.empty(B, KR, N, device=dev)
lse = torch.empty(B, KR, device=dev)
nll = torch.empty(B, KR, device=dev)
corr = torch.empty(B, KR, dtype=torch.int32, device=dev)
step('infonce_idx_fwd', lambda: _native.call('vqcpc_infonce_idx_fwd', p_(c), p_(W), p_(T), p_(idx[n_neg + n_left:]), p_(idx), B, N, KR, zd, cd, Kc,
                                              p_(u), p_(pos), p_(neg), p_(lse), p_(nll), p_(corr), st()))
dbg = (ctypes.c_longlong * 16)()
_native.call('vqcpc_idx_debug_read', dbg)
print('fwd dbg [setup, table wait, loop, steps, top, u, score, softmax+out]:', list(dbg)[:8], flush=True)
dc = torch.zeros(B, cd, device=dev)
dW = torch.zeros_like(W)
dT = torch.zeros_like(T)
cp = torch.empty(B, KR, device=dev)
cn = torch.empty(B, N, KR, device=dev)
part = torch.empty(_native.load().vqcpc_infonce_idx_scratch(B, KR, Kc), device=dev)
step('infonce_idx_bwd', lambda: _native.call('vqcpc_infonce_idx_bwd', p_(c), p_(W), p_(T), p_(idx[n_neg + n_left:]), p_(idx), p_(u), p_(pos), p_(neg),
                                              p_(lse), None, 1.0 / B, B, N, KR, zd, cd, Kc, p_(dc), p_(cp), p_(cn), p_(part), p_(dW), p_(dT), st()))
_native.call('vqcpc_idx_debug_read', dbg)
print('bwd dbg [setup, loop, end, coef, wait-empty, build, du, dc/dW]:', list(dbg)[8:16], flush=True)
# reference for dT / dW / dc in fp64
i_neg = idx[:n_neg].view(B, N, KR)
i_pos = idx[n_neg + n_left:].view(B, KR)
cand = torch.cat([i_pos.unsqueeze(1), i_neg], 1)                 # (B, 1 + N, KR)
coef = torch.cat([cp.unsqueeze(1), cn], 1).double()             # (B, 1 + N, KR)
ud = u.double()
dT_ref = torch.zeros(Kc, zd, dtype=torch.float64, device=dev)
dT_ref.index_add_(0, cand.reshape(-1), (coef.unsqueeze(-1) * ud.unsqueeze(1)).reshape(-1, zd))
du = torch.einsum('bnk,bnkz->bkz', coef, T.double()[cand])
dW_ref = torch.einsum('bkz,bc->zck', du, c.double())
dc_ref = torch.einsum('bkz,zck->bc', du, W.double())
rel = lambda a, b: ((a.double() - b).norm() / b.norm()).item()  # noqa: E731
print('dT rel', rel(dT, dT_ref), 'dW rel', rel(dW, dW_ref), 'dc rel', rel(dc, dc_ref), flush=True)
dyl = torch.randn(n_left, zd, device=dev, generator=g)
step('table_scatter_add', lambda: _native.call('vqcpc_table_scatter_add', p_(idx[n_neg:n_neg + n_left]), n_left, p_(dyl), zd, Kc, p_(dT), st()))
dz = torch.empty(P, D, device=dev)
dE = torch.zeros(Kc, D, device=dev)
gl = torch.full((P,), 0.1, device=dev)
row0 = (ctypes.c_int64 * 3)(0, n_neg, n_neg + n_left)
rows = (ctypes.c_int64 * 3)(n_neg, n_left, n_right)
kind = (ctypes.c_int * 3)(2, 1, 2)
per_b = (ctypes.c_int * 3)(N * KR, 1, KR)
pa = (ctypes.c_void_p * 3)(cn.data_ptr(), dyl.data_ptr(), cp.data_ptr())
pb = (ctypes.c_void_p * 3)(u.data_ptr(), 0, u.data_ptr())
step('vq_idx_bwd', lambda: _native.call('vqcpc_vq_idx_bwd', p_(z), P, D, p_(E), p_(J), Kc, zd, p_(idx), 0.25, 1, p_(gl), 3, row0, rows, kind, pa, pb,
                                         per_b, KR, p_(dz), p_(dE), p_(dT), st()))
_native.call('vqcpc_idx_debug_read', dbg)
print('vq_idx_dz dbg [setup, J wait, loop]:', list(dbg)[:3], flush=True)
print('all ok')
