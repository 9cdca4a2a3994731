#!/usr/bin/env python
"""Kernel table of one bf16 IPA layer forward + backward (BASELINE config 2, B=32) (GPU only)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import diffab_pytorch_b200  # noqa
from diffab_pytorch_b200 import synth
from diffab_pytorch_b200.diffab_pytorch import InvariantPointAttentionLayer
from torch.profiler import profile, ProfilerActivity

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
dev = "cuda"
layer = InvariantPointAttentionLayer(128, 64, 32, 8, 8, 8).to(dev)
layer.load_state_dict(synth.synthetic_state(synth.ipa_layer_shapes(128, 64, 8, 32, 8, 8), seed=0))
g = torch.Generator(device=dev).manual_seed(0)
x = torch.randn(B, 128, 128, device=dev, generator=g).requires_grad_(True)
e = torch.randn(B, 128, 128, 64, device=dev, generator=g).bfloat16().requires_grad_(True)
R = synth.uniform_rotations(B, 128, device=dev)
t = 10 * torch.randn(B, 128, 3, device=dev, generator=g)
gy = torch.randn(B, 128, 128, device=dev, generator=g)
def step():
    y = layer(x, e, R, t)
    y.backward(gy)
for _ in range(3):
    step()
torch.cuda.synchronize()
N = 10
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(N):
        step()
    torch.cuda.synchronize()
rows = sorted(prof.key_averages(), key=lambda e: -e.self_device_time_total)
tot = sum(e.self_device_time_total for e in rows)
print(f"total kernel time per fwd+bwd: {tot / N:.1f} us")
for e in rows[:30]:
    print(f"{e.self_device_time_total / N:8.1f} us  x{e.count / N:4.1f}  {e.key[:110]}")
