#!/usr/bin/env python
"""Device time of one graphed reverse step at B patches, with the regrouped glue on / off (GPU only)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import diffab_pytorch_b200  # noqa
from diffab_pytorch_b200 import synth
from diffab_pytorch_b200.diffab_pytorch import DiffAb

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
dev = torch.device("cuda")
shapes = torch.load(os.path.join(ROOT, "tests", "golden", "state_shapes.pt"), weights_only=False)
g = torch.Generator(device=dev).manual_seed(0)
res = torch.randn(B, 128, 128, device=dev, generator=g)
pair = torch.randn(B, 128, 128, 64, device=dev, generator=g).bfloat16()
batch = {k: v.to(dev) for k, v in synth.make_patches(B, 128, seed=1, with_distmat=False).items()}
s, x, O, m = batch["seq_idx"], batch["xyz"][:, :, 1].contiguous(), batch["orientations"], batch["generation_mask"]
for glue_on in (True, False):
    model = DiffAb(128, 64, 6, 32, 8, 8, 8, device=dev).eval()
    model.load_state_dict(synth.synthetic_state(shapes, seed=0))
    if not glue_on:
        model.denoiser.sampling_cache = lambda *a, **k: None
    for _ in range(2):
        model.sample_from_context(s, x, O, res, pair, m, use_cuda_graph=True, t_start=100, t_stop=96)
    torch.cuda.synchronize()
    a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    model.sample_from_context(s, x, O, res, pair, m, use_cuda_graph=True, t_start=100, t_stop=51)
    b_.record()
    torch.cuda.synchronize()
    print(f"B={B} glue regrouped={glue_on}: {a.elapsed_time(b_) * 1000 / 50:.1f} us per reverse step")
    del model
