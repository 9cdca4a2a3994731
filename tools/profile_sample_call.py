#!/usr/bin/env python
"""Where a DiffAb.sample() call spends its time outside the reverse steps (GPU only): torch.profiler kernel table of one
warm call on pinned host inputs (B = 256, T = 100), the kernels of the reverse steps listed separately from the rest
(host-to-device copies, context encoders, pair-bias planes, copies into the graphs' static buffers, device-to-host)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

import diffab_pytorch_b200  # noqa: E402,F401
from diffab_pytorch_b200 import synth  # noqa: E402
from diffab_pytorch_b200.diffab_pytorch import DiffAb  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
dev = torch.device("cuda")
shapes = torch.load(os.path.join(ROOT, "tests", "golden", "state_shapes.pt"), weights_only=False)
model = DiffAb(128, 64, 6, 32, 8, 8, 8, device=dev).eval()
model.load_state_dict(synth.synthetic_state(shapes, seed=0))
batch = {k: v.pin_memory() for k, v in synth.make_patches(B, 128, seed=1, with_distmat=False).items()}
keys = ("seq_idx", "xyz", "orientations", "backbone_dihedrals", None, "pairwise_dihedrals", "atom_mask", "chain_idx",
        "residue_idx", "generation_mask", "residue_mask")
args = [batch[k] if k else None for k in keys]


def call():
    out = model.sample(*args)
    return {k: v.cpu() for k, v in out.items()}


for _ in range(2):
    call()
torch.cuda.synchronize()
t0 = time.perf_counter()
call()
torch.cuda.synchronize()
print(f"wall clock of one call: {(time.perf_counter() - t0) * 1e3:.2f} ms")
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    call()
    torch.cuda.synchronize()
step_kernels = ("ipa_core_kernel", "ipa_proj_kernel", "gemm_bf16_kernel", "denoiser_heads_kernel", "front_act_kernel",
                "igso3_sample_kernel", "reverse_step_kernel", "distribution_")
rows = [(e.key, e.self_device_time_total, e.count) for e in prof.key_averages() if e.self_device_time_total > 0]
steps = sum(t for k, t, _ in rows if any(s in k for s in step_kernels))
rest = sorted(((t, k, c) for k, t, c in rows if not any(s in k for s in step_kernels)), reverse=True)
print(f"device time in the reverse steps' kernels: {steps / 1e3:.2f} ms; everything else: {sum(t for t, _, _ in rest) / 1e3:.2f} ms")
for t, k, c in rest[:25]:
    print(f"  {t / 1e3:8.3f} ms  x{c:<4d} {k[:110]}")
if os.environ.get("SHAPES") == "1":      # which copies / conversions make up aten::copy_
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU], record_shapes=True) as prof2:
        call()
        torch.cuda.synchronize()
    big = [(e.self_device_time_total, e.key, str(e.input_shapes)[:90], e.count)
           for e in prof2.key_averages(group_by_input_shape=True) if e.key in ("aten::copy_", "aten::mul", "aten::cat")]
    for t, k, shp, c in sorted(big, reverse=True)[:14]:
        print(f"  {t / 1e3:8.3f} ms  x{c:<3d} {k} {shp}")
