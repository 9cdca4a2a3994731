#!/usr/bin/env python
"""torch.profiler kernel table of the config-5 training step (B=64, bf16 IPA path) (GPU only): the same launches that
GraphedTrainStep captures (gradients into fresh tensors + one multi-tensor copy into the bucket, single-kernel Adam), run
eagerly under the profiler."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import diffab_pytorch_b200  # noqa
from diffab_pytorch_b200 import synth
from diffab_pytorch_b200.diffab_pytorch import DiffAb
from diffab_pytorch_b200.distributed import FlatAdam, GradientBucket, GraphedTrainStep, diffab_loss_terms
from torch.profiler import profile, ProfilerActivity

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
dev = torch.device("cuda")
shapes = torch.load(os.path.join(ROOT, "tests", "golden", "state_shapes.pt"), weights_only=False)
model = DiffAb(128, 64, 6, 32, 8, 8, 8, device=dev).train()
model.load_state_dict(synth.synthetic_state(shapes, seed=0))
model.train_precision = "bf16"
torch.set_float32_matmul_precision("high")   # as the reference's train.py:47
bucket = GradientBucket(model.parameters())
opt = FlatAdam(bucket, lr=1e-4)
batch = {k: v.to(dev) for k, v in synth.make_patches(B, 128, seed=2000, with_distmat=False).items()}
batch["distmat"] = torch.cat([synth.pairwise_atom_distances(batch["xyz"][i:i + 8]) for i in range(0, B, 8)])
graphed = GraphedTrainStep(lambda: diffab_loss_terms(model, batch), bucket, opt)


def step():
    graphed._backward_part()
    graphed._reduce_part()
    opt.step()


for _ in range(3):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(2):
        step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=100))
