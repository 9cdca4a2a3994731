#!/usr/bin/env python
"""torch.profiler kernel table of eager reverse steps at B patches (GPU only)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import diffab_pytorch_b200  # noqa
from diffab_pytorch_b200 import synth
from diffab_pytorch_b200.diffab_pytorch import DiffAb
from torch.profiler import profile, ProfilerActivity

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
dev = torch.device("cuda")
shapes = torch.load(os.path.join(ROOT, "tests", "golden", "state_shapes.pt"), weights_only=False)
g = torch.Generator(device=dev).manual_seed(0)
res = torch.randn(B, 128, 128, device=dev, generator=g)
pair = torch.randn(B, 128, 128, 64, device=dev, generator=g).bfloat16()
batch = {k: v.to(dev) for k, v in synth.make_patches(B, 128, seed=1, with_distmat=False).items()}
s, x, O, m = batch["seq_idx"], batch["xyz"][:, :, 1].contiguous(), batch["orientations"], batch["generation_mask"]
model = DiffAb(128, 64, 6, 32, 8, 8, 8, device=dev).eval()
model.load_state_dict(synth.synthetic_state(shapes, seed=0))
model.sample_from_context(s, x, O, res, pair, m, use_cuda_graph=False, t_start=100, t_stop=98)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    model.sample_from_context(s, x, O, res, pair, m, use_cuda_graph=False, t_start=100, t_stop=91)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=40, max_name_column_width=90))
