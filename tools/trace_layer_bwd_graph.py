#!/usr/bin/env python
"""Kernel timeline of ONE graph replay of the bf16 IPA layer forward + backward (GPU only): start offset, duration and
stream of every kernel as CUPTI records them inside the replay (warm L2 unless FLUSH=1), i.e. the critical path with its
side-stream overlaps - ncu's launch list serialises the kernels and runs them cold.

    python tools/trace_layer_bwd_graph.py [B]
"""
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
        x.grad = None; e.grad = None
        step()
torch.cuda.current_stream().wait_stream(side)
torch.cuda.synchronize()
graph = torch.cuda.CUDAGraph()
x.grad = None; e.grad = None
with torch.cuda.graph(graph):
    step()
for _ in range(3):
    graph.replay()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        if os.environ.get("FLUSH") == "1":
            flush.zero_()
        graph.replay()
        torch.cuda.synchronize()
evs = [ev for ev in prof.events() if ev.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda ev: ev.time_range.start)
# the last replay: kernels after the last gap > 50 us
runs, cur = [], []
for ev in evs:
    if cur and ev.time_range.start - cur[-1].time_range.end > 50:
        runs.append(cur); cur = []
    cur.append(ev)
runs.append(cur)
last = [r for r in runs if len(r) > 8][-1]
t0 = last[0].time_range.start
print(f"B={B}: {len(last)} device activities in the last replay, span {last[-1].time_range.end - t0:.1f} us")
for ev in last:
    print(f"  +{ev.time_range.start - t0:7.1f}  {ev.time_range.end - ev.time_range.start:6.1f} us  {ev.name[:90]}")
