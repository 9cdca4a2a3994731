#!/usr/bin/env python
"""CUDA-event timing of the bf16 IPA layer forward + backward (training path), B patches (GPU only)."""
import os, sys, statistics
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import diffab_pytorch_b200  # noqa
from diffab_pytorch_b200 import _lib, synth
from diffab_pytorch_b200.diffab_pytorch import InvariantPointAttentionLayer

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
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def step():
    y = layer(x, e, R, t)
    y.backward(gy)
    x.grad = None; e.grad = None
    for p in layer.parameters():
        p.grad = None


for _ in range(3):
    step()
torch.cuda.synchronize()
fw, tot = [], []
for _ in range(10):
    flush.zero_()
    a, m, b_ = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    a.record()
    y = layer(x, e, R, t)
    m.record()
    y.backward(gy)
    b_.record()
    torch.cuda.synchronize()
    fw.append(a.elapsed_time(m) * 1000); tot.append(a.elapsed_time(b_) * 1000)
    x.grad = None; e.grad = None
print(f"B={B}: fwd {statistics.median(fw):.1f} us, fwd+bwd {statistics.median(tot):.1f} us (median of 10, L2 flushed)")
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=60))
