#!/usr/bin/env python
"""Whole training step (config 5) captured in one CUDA graph: replay time vs eager (GPU only)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import diffab_pytorch_b200  # noqa
from diffab_pytorch_b200 import synth
from diffab_pytorch_b200.diffab_pytorch import DiffAb
from diffab_pytorch_b200.distributed import GradientBucket, ddp_step, diffab_loss_terms

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
dev = torch.device("cuda")
shapes = torch.load(os.path.join(ROOT, "tests", "golden", "state_shapes.pt"), weights_only=False)
model = DiffAb(128, 64, 6, 32, 8, 8, 8, device=dev).train()
model.load_state_dict(synth.synthetic_state(shapes, seed=0))
model.train_precision = "bf16"
torch.set_float32_matmul_precision("high")
bucket = GradientBucket(model.parameters())
opt = torch.optim.Adam(model.parameters(), lr=1e-4, capturable=True)
batch = {k: v.to(dev) for k, v in synth.make_patches(B, 128, seed=2000, with_distmat=False).items()}
batch["distmat"] = torch.cat([synth.pairwise_atom_distances(batch["xyz"][i:i + 8]) for i in range(0, B, 8)])

def body():
    bucket.zero()
    num, cnt = diffab_loss_terms(model, batch)
    loss = num / cnt
    loss.backward()
    opt.step()
    return loss.detach()

def timeit(fn, n=10):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n

s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    for _ in range(3):
        l0 = body()
torch.cuda.current_stream().wait_stream(s)
print("eager loss", float(l0), "eager ms", timeit(body))
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    static_loss = body()
g.replay()
torch.cuda.synchronize()
print("graph loss", float(static_loss), "graph ms", timeit(g.replay))
l = []
for _ in range(5):
    g.replay()
    l.append(float(static_loss))
print("losses over replays", l)
