#!/usr/bin/env python
"""CUDA-event timing of the three launches of the bf16 IPA layer (phase mask), B patches."""
import os, sys, statistics
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import diffab_pytorch_b200  # noqa
from diffab_pytorch_b200 import _lib, synth
from diffab_pytorch_b200.diffab_pytorch import InvariantPointAttentionLayer

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
dev = "cuda"
lib = _lib.lib()
layer = InvariantPointAttentionLayer(128, 64, 32, 8, 8, 8).to(dev)
layer.load_state_dict(synth.synthetic_state(synth.ipa_layer_shapes(128, 64, 8, 32, 8, 8), seed=0))
g = torch.Generator(device=dev).manual_seed(0)
x = torch.randn(B, 128, 128, device=dev, generator=g)
e = torch.randn(B, 128, 128, 64, device=dev, generator=g).bfloat16()
R = synth.uniform_rotations(B, 128, device=dev)
t = 10 * torch.randn(B, 128, 3, device=dev, generator=g)
with torch.no_grad():
    bias = layer.pair_bias(e)
    for _ in range(3):
        layer(x, e, R, t, bias)
    for mask, name in ((1, "proj (x->Qp,Kp,Vp)"), (2, "attention core"), (4, "to_out GEMM"), (7, "whole layer")):
        lib.dab_debug_set_phase_mask(mask)
        ts = []
        for _ in range(20):
            a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); layer(x, e, R, t, bias); b_.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b_) * 1000)
        print(f"{name:22s} {statistics.mean(ts):8.1f} us  (min {min(ts):.1f})")
    lib.dab_debug_set_phase_mask(7)
