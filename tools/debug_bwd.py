#!/usr/bin/env python
"""Bring-up of the tcgen05 IPA backward: every intermediate buffer against an fp64 autograd reference (GPU only)."""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import diffab_pytorch_b200  # noqa
from diffab_pytorch_b200 import _lib, synth
from diffab_pytorch_b200.diffab_pytorch import InvariantPointAttentionLayer, _ipa_structs
from oracle import ipa as oipa

dev = "cuda"
B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
lib = _lib.lib()
shp = synth.ipa_layer_shapes(128, 64, 8, 32, 8, 8)
w = synth.synthetic_state(shp, seed=0)
x, e, R, t = [v.to(dev) for v in synth.make_ipa_inputs(B, 128, 128, 64, seed=100)]
gy = torch.randn(B, 128, 128, device=dev, generator=torch.Generator(device=dev).manual_seed(7)) * 1e-3
layer = InvariantPointAttentionLayer(128, 64, 32, 8, 8, 8).to(dev)
layer.load_state_dict(w)
e16 = e.bfloat16()


def rel(a, b, name):
    a, b = a.double(), b.double()
    err = (a - b).abs().max().item()
    ref = b.abs().max().item()
    bad = (~torch.isfinite(a)).sum().item()
    print(f"{name:28s} max|err| {err:.3e}  max|ref| {ref:.3e}  rel {err / max(ref, 1e-300):.3e}  nonfinite {bad}")
    return err / max(ref, 1e-300)


# ---- fp64 reference with intermediates
wd = {k: v.detach().to(dev).double().requires_grad_(True) for k, v in w.items()}
xd = x.double().requires_grad_(True)
ed = e16.double().requires_grad_(True)
yd, attn, logit = oipa.ipa_layer(wd, xd, ed, R.double(), t.double(), 8, return_attn=True)
logit.retain_grad()
(yd * gy.double()).sum().backward()

# ---- kernel path
lib.dab_debug_bwd_keep_qkv(1)
xg = x.clone().requires_grad_(True)
eg = e16.clone().requires_grad_(True)
y = layer(xg, eg, R, t)
torch.cuda.synchronize()
rel(y, yd, "y (fwd train)")
(y * gy).sum().backward()
torch.cuda.synchronize()

dims = _ipa_structs(layer, B, 128)
bws = layer._last_bwd_ws
ptrs = (ctypes.c_void_p * 10)()
_lib.check(lib.dab_debug_bwd_sm100_buffers(ctypes.byref(dims), ctypes.c_void_p(bws.data_ptr()), ptrs), "buffers")
offs = [(p or 0) - bws.data_ptr() for p in ptrs]
rows = B * 128


def view(k, nbytes, dtype, shape):
    return bws[offs[k]: offs[k] + nbytes].view(dtype).view(*shape)


dObf = view(1, rows * 8 * 64 * 2, torch.bfloat16, (rows, 8, 64))    # (the scaled fp16 copy, dopair, Delta stay in shared memory)
Pn = view(5, B * 8 * 128 * 128 * 2, torch.bfloat16, (B, 8, 128, 128))
dL = view(6, B * 8 * 128 * 128 * 2, torch.bfloat16, (B, 8, 128, 128))
dQ = view(7, rows * 8 * 64 * 4, torch.float32, (rows, 8, 64))
dK = view(8, rows * 8 * 64 * 4, torch.float32, (rows, 8, 64))
dV = view(9, rows * 8 * 64 * 4, torch.float32, (rows, 8, 64))

# Delta_i,h = sum_j P dP = sum_j attn * dattn
rel(Pn, attn, "P (normalised)")
rel(dL, logit.grad, "dl")
print("   dl row sums (should be ~0):", dL.float().sum(-1).abs().max().item(), " ref", logit.grad.sum(-1).abs().max().item())
# dq raw: U = sum_j dl [ks | k~hi ...]
ks = (xd @ wd["to_k_scalar.weight"].t()).view(B, 128, 8, 32)
U_ref = torch.einsum("bhij,bjhd->bihd", logit.grad, ks).reshape(rows, 8, 32)
rel(dQ[:, :, :32], U_ref, "dQ scalar part (U)")
rel(dQ[:, :, 59], logit.grad.sum(-1).permute(0, 2, 1).reshape(rows, 8), "dQ col 59 (sum_j dl)")
qs = (xd @ wd["to_q_scalar.weight"].t()).view(B, 128, 8, 32)
W_ref = torch.einsum("bhij,bihd->bjhd", logit.grad, qs).reshape(rows, 8, 32) * (3 ** -0.5 * 32 ** -0.5 * 1.4426950408889634)
rel(dK[:, :, :32], W_ref, "dK scalar part (W)")
rel(dK[:, :, 56], logit.grad.sum(-2).permute(0, 2, 1).reshape(rows, 8), "dK col 56 (sum_i dl)")
vs_g = torch.einsum("bhij,bihd->bjhd", attn, (gy.double().view(rows, 128) @ wd["to_out.weight"])[:, :256].view(B, 128, 8, 32))
rel(dV[:, :, :32], vs_g.reshape(rows, 8, 32), "dV scalar part")

print("---- final gradients")
rel(xg.grad, xd.grad, "dx")
rel(eg.grad, ed.grad, "de")
for n, p in layer.named_parameters():
    rel(p.grad, wd[n].grad, "d " + n)
