#!/usr/bin/env python
"""Device time of the six-layer bf16 IPA stack (inference hand-off, as DiffAb.sample runs it) replayed from a CUDA graph.

    python tools/time_stack.py [B] [n_layers]

Prints microseconds per layer (graph replay, inputs of all patches larger than L2 at B >= 64) and the fraction of
SURVEY 8(d)'s per-layer HBM roofline (B (L^2 C + 2 L D + 12 L) 2 + params bytes at the measured copy bandwidth)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import diffab_pytorch_b200  # noqa: E402,F401
from diffab_pytorch_b200 import synth  # noqa: E402
from diffab_pytorch_b200.diffab_pytorch import InvariantPointAttentionModule  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
NL = int(sys.argv[2]) if len(sys.argv) > 2 else 6
dev = "cuda"
mod = InvariantPointAttentionModule(NL, 128, 64, 32, 8, 8, 8).to(dev)
shp = synth.ipa_layer_shapes(128, 64, 8, 32, 8, 8)
for k, layer in enumerate(mod.layers):
    layer.load_state_dict(synth.synthetic_state(shp, seed=k))
g = torch.Generator(device=dev).manual_seed(0)
x = torch.randn(B, 128, 128, device=dev, generator=g)
e = torch.randn(B, 128, 128, 64, device=dev, generator=g).bfloat16()
R = synth.uniform_rotations(B, 128, device=dev)
t = 10 * torch.randn(B, 128, 3, device=dev, generator=g)
with torch.no_grad():
    bias = mod.precompute_pair_bias(e)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            y = mod(x, e, R, t, bias)
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        y = mod(x, e, R, t, bias)
    for _ in range(3):
        graph.replay()
    torch.cuda.synchronize()
    ts = []
    for _ in range(20):
        a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); graph.replay(); b_.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b_) * 1000 / NL)
ts.sort()
peak = 6528.4
try:
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except (OSError, KeyError, ValueError):
    pass
alg = B * (128 * 128 * 64 + 2 * 128 * 128 + 12 * 128) * 2 + 303752 * 4
us = ts[len(ts) // 2]
print(f"B={B}: {us:.1f} us per layer (median of 20 replays of {NL} layers; min {ts[0]:.1f}); 8(d) bytes {alg / 1e6:.1f} MB -> "
      f"{alg / us / 1e3:.0f} GB/s = {alg / us / 1e3 / peak:.3f} of {peak:.0f} GB/s; finite={bool(torch.isfinite(y).all())}")
if os.environ.get("KERNELS") == "1":     # per-kernel device times inside the replays (CUPTI through torch.profiler)
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(3):
            graph.replay()
        torch.cuda.synchronize()
    for ev in sorted(prof.key_averages(), key=lambda e: -e.device_time_total)[:6]:
        print(f"    {ev.device_time_total / ev.count:8.1f} us x{ev.count:<3d} {ev.key[:90]}")
