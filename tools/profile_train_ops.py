#!/usr/bin/env python
"""torch.profiler operator table (grouped by input shapes) of the config-5 training step (GPU only)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import diffab_pytorch_b200  # noqa
from diffab_pytorch_b200 import synth
from diffab_pytorch_b200.diffab_pytorch import DiffAb
from diffab_pytorch_b200.distributed import GradientBucket, ddp_step, diffab_loss_terms
from torch.profiler import profile, ProfilerActivity

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
dev = torch.device("cuda")
shapes = torch.load(os.path.join(ROOT, "tests", "golden", "state_shapes.pt"), weights_only=False)
model = DiffAb(128, 64, 6, 32, 8, 8, 8, device=dev).train()
model.load_state_dict(synth.synthetic_state(shapes, seed=0))
model.train_precision = "bf16"
torch.set_float32_matmul_precision("high")
bucket = GradientBucket(model.parameters())
opt = torch.optim.Adam(model.parameters(), lr=1e-4)
batch = {k: v.to(dev) for k, v in synth.make_patches(B, 128, seed=2000, with_distmat=False).items()}
batch["distmat"] = torch.cat([synth.pairwise_atom_distances(batch["xyz"][i:i + 8]) for i in range(0, B, 8)])
step = lambda: ddp_step(lambda: diffab_loss_terms(model, batch), bucket, opt)
for _ in range(3):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True) as prof:
    step()
    torch.cuda.synchronize()
rows = [e for e in prof.key_averages(group_by_input_shape=True) if e.self_device_time_total > 0]
rows.sort(key=lambda e: -e.self_device_time_total)
for e in rows[:60]:
    print(f"{e.self_device_time_total:9.1f} us  x{e.count:3d}  {e.key:40s} {str(e.input_shapes)[:150]}")
print("---- by CPU time")
rows = list(prof.key_averages())
rows.sort(key=lambda e: -e.self_cpu_time_total)
for e in rows[:25]:
    print(f"{e.self_cpu_time_total:9.1f} us  x{e.count:3d}  {e.key[:80]}")
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(5):
    step()
b.record()
torch.cuda.synchronize()
print("step ms", a.elapsed_time(b) / 5)
