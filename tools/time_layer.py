#!/usr/bin/env python
"""CUDA-event timing of the three launches of the bf16 IPA layer (phase mask), B patches."""
import ctypes, os, sys, statistics
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import diffab_pytorch_b200  # noqa
from diffab_pytorch_b200 import _lib, synth
from diffab_pytorch_b200._lib import ptr
from diffab_pytorch_b200.diffab_pytorch import InvariantPointAttentionLayer, _ipa_structs

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
    dims = _ipa_structs(layer, B, 128)
    packed = layer._packed_weights(dims)
    ws = layer._workspace(lib.dab_ipa_sm100_workspace_bytes(ctypes.byref(dims)), x.device)
    y = torch.empty(B, 128, 128, device=dev)

    def run(stages):
        _lib.check(lib.dab_ipa_fwd_sm100_stages(ctypes.byref(dims), ptr(packed), ptr(x), None, ptr(e), ptr(bias), ptr(R), ptr(t),
                                                ptr(y), None, ptr(ws), ws.numel(), stages, _lib.stream_ptr()), "stages")

    for _ in range(3):
        run(7)
    for mask, name in ((1, "proj (x->Qp,Kp,Vp)"), (2, "attention core"), (4, "to_out GEMM"), (7, "whole layer")):
        ts = []
        for _ in range(20):
            a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); run(mask); b_.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b_) * 1000)
        print(f"{name:22s} {statistics.mean(ts):8.1f} us  (min {min(ts):.1f})")
