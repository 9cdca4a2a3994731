#!/usr/bin/env python
"""Reverse steps on patches of 256 residues (the upper end of the reference's preprocessed patch lengths): the bf16
tensor-core path (two blocks of 128 per patch) against the shape-generic fp32 kernels (GPU only).

    python tools/time_sample_256.py [B] [L]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import diffab_pytorch_b200  # noqa: E402,F401
from diffab_pytorch_b200 import synth  # noqa: E402
from diffab_pytorch_b200.diffab_pytorch import DiffAb  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
L = int(sys.argv[2]) if len(sys.argv) > 2 else 256
dev = torch.device("cuda")
shapes = torch.load(os.path.join(ROOT, "tests", "golden", "state_shapes.pt"), weights_only=False)
model = DiffAb(128, 64, 6, 32, 8, 8, 8, device=dev).eval()
model.load_state_dict(synth.synthetic_state(shapes, seed=0))
g = torch.Generator(device=dev).manual_seed(0)
res = torch.randn(B, L, 128, device=dev, generator=g)
pair = torch.randn(B, L, L, 64, device=dev, generator=g)
batch = {k: v.to(dev) for k, v in synth.make_patches(B, L, seed=1, with_distmat=False).items()}
s, x, O, m = batch["seq_idx"], batch["xyz"][:, :, 1].contiguous(), batch["orientations"], batch["generation_mask"]
for name, p, steps, graph in (("bf16 tensor-core path", pair.bfloat16(), 20, True), ("fp32 kernels", pair, 2, False)):
    for _ in range(2):
        model.sample_from_context(s, x, O, res, p, m, use_cuda_graph=graph, t_start=100, t_stop=100 - steps + 1)
    torch.cuda.synchronize()
    a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    out = model.sample_from_context(s, x, O, res, p, m, use_cuda_graph=graph, t_start=100, t_stop=100 - steps + 1)
    b_.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b_) / steps
    print(f"B={B} L={L} {name}: {ms:.2f} ms per reverse step -> {B / (ms * 100 / 1e3):.1f} patches/s at T=100; "
          f"finite={bool(torch.isfinite(out['translations']).all())}")
