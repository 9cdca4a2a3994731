#!/usr/bin/env python
"""Run the bf16 IPA layer forward + backward a few times at B patches (profiling target for ncu)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import diffab_pytorch_b200  # noqa
from diffab_pytorch_b200 import synth
from diffab_pytorch_b200.diffab_pytorch import InvariantPointAttentionLayer

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
n = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = "cuda"
layer = InvariantPointAttentionLayer(128, 64, 32, 8, 8, 8).to(dev)
layer.load_state_dict(synth.synthetic_state(synth.ipa_layer_shapes(128, 64, 8, 32, 8, 8), seed=0))
g = torch.Generator(device=dev).manual_seed(0)
x = torch.randn(B, 128, 128, device=dev, generator=g).requires_grad_(True)
e = torch.randn(B, 128, 128, 64, device=dev, generator=g).bfloat16().requires_grad_(True)
R = synth.uniform_rotations(B, 128, device=dev)
t = 10 * torch.randn(B, 128, 3, device=dev, generator=g)
gy = torch.randn(B, 128, 128, device=dev, generator=g)
for _ in range(n):
    layer(x, e, R, t).backward(gy)
torch.cuda.synchronize()
print("ok", float(x.grad.abs().max()))
