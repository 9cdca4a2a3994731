#!/usr/bin/env python
"""bf16 IPA layer forward + backward captured in one CUDA graph; CUDA-event timing with an L2 flush between replays."""
import os, sys, statistics
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import diffab_pytorch_b200  # noqa
from diffab_pytorch_b200 import synth
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
    x.grad = None; e.grad = None
    for p_ in layer.parameters():      # as after zero_grad(set_to_none=True): no accumulation kernels
        p_.grad = None
    y = layer(x, e, R, t)
    y.backward(gy)

side = torch.cuda.Stream()
side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    for _ in range(3):
        step()
        x.grad = None; e.grad = None
        for p in layer.parameters():
            p.grad = None
torch.cuda.current_stream().wait_stream(side)
graph = torch.cuda.CUDAGraph()
with torch.cuda.graph(graph):
    step()
ref_dx = x.grad.clone()
ts = []
for _ in range(20):
    flush.zero_()
    a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); graph.replay(); b_.record()
    torch.cuda.synchronize()
    ts.append(a.elapsed_time(b_) * 1000)
print(f"B={B}: graphed fwd+bwd {statistics.median(ts):.1f} us (min {min(ts):.1f}); dx finite {bool(torch.isfinite(x.grad).all())}")
