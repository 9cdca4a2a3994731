#!/usr/bin/env python
"""Experiment: two half-batches replayed as two CUDA graphs on two streams vs one full-batch graph (GPU only)."""
import copy, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import diffab_pytorch_b200  # noqa
from diffab_pytorch_b200 import synth
from diffab_pytorch_b200.diffab_pytorch import DiffAb

B = 256
dev = torch.device("cuda")
shapes = torch.load(os.path.join(ROOT, "tests", "golden", "state_shapes.pt"), weights_only=False)
g = torch.Generator(device=dev).manual_seed(0)
res = torch.randn(B, 128, 128, device=dev, generator=g)
pair = torch.randn(B, 128, 128, 64, device=dev, generator=g).bfloat16()
batch = {k: v.to(dev) for k, v in synth.make_patches(B, 128, seed=1, with_distmat=False).items()}
s, x, O, m = batch["seq_idx"], batch["xyz"][:, :, 1].contiguous(), batch["orientations"], batch["generation_mask"]

def make(n_lo, n_hi):
    model = DiffAb(128, 64, 6, 32, 8, 8, 8, device=dev).eval()
    model.load_state_dict(synth.synthetic_state(shapes, seed=0))
    sl = slice(n_lo, n_hi)
    model.sample_from_context(s[sl], x[sl], O[sl], res[sl], pair[sl], m[sl], use_cuda_graph=True, t_start=100, t_stop=99)
    return model, model._graph_cache["graph"]

full, gf = make(0, B)
a, ga = make(0, B // 2)
b, gb = make(B // 2, B)
torch.cuda.synchronize()
def timed(fn, n=50):
    fn(3)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(n); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1000 / n

def run_full(n):
    for _ in range(n):
        gf.replay()

s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run_two(n):
    cur = torch.cuda.current_stream()
    s1.wait_stream(cur); s2.wait_stream(cur)
    for _ in range(n):
        with torch.cuda.stream(s1):
            ga.replay()
        with torch.cuda.stream(s2):
            gb.replay()
    cur.wait_stream(s1); cur.wait_stream(s2)

print(f"one graph, 256 patches        : {timed(run_full):8.1f} us per reverse step")
print(f"two graphs x 128, two streams : {timed(run_two):8.1f} us per reverse step")
def run_seq(n):
    for _ in range(n):
        ga.replay(); gb.replay()
print(f"two graphs x 128, one stream  : {timed(run_seq):8.1f} us per reverse step")
