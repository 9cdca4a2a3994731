#!/usr/bin/env python
"""Mixed-precision PairEmbedding (fused RBF + _PairMlpFunction) vs the fp32 module: per-parameter gradient errors (GPU only)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import diffab_pytorch_b200  # noqa
from diffab_pytorch_b200 import synth
from diffab_pytorch_b200.diffab_pytorch import DiffAb

dev = torch.device("cuda")
shapes = torch.load(os.path.join(ROOT, "tests", "golden", "state_shapes.pt"), weights_only=False)
model = DiffAb(128, 64, 6, 32, 8, 8, 8, device=dev).train()
model.load_state_dict(synth.synthetic_state(shapes, seed=0))
pe = model.pair_context_embedding
torch.nn.init.normal_(pe.pair2distcoef.weight, std=0.5)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
batch = synth.make_patches(B, 128, seed=31)
batch["atom_mask"][1, 9, 3:] = False
b = {k: v.to(dev) for k, v in batch.items()}
ctx = b["residue_mask"] & ~b["generation_mask"]
args = (b["seq_idx"], b["distmat"], b["pairwise_dihedrals"], b["residue_idx"], b["chain_idx"], b["atom_mask"], ctx, ctx)
gy = torch.randn(B, 128, 128, 64, device=dev, generator=torch.Generator(device=dev).manual_seed(5))
out, grads = {}, {}
for mode in ("fp32", "fp64", "fused"):
    pe.fused_rbf = mode == "fused"
    pe.zero_grad()
    if mode == "fp64":
        pe.double()
        y = pe(*[a.double() if a.is_floating_point() else a for a in args])
    else:
        y = pe(*args)
    (y * gy.to(y.dtype) if mode == "fp64" else y * gy).sum().backward()
    out[mode] = y.detach().double()
    grads[mode] = {n: p.grad.double().clone() for n, p in pe.named_parameters() if p.grad is not None}
    if mode == "fp64":
        pe.float()
pe.fused_rbf = False
ref = out["fp64"]
for m in ("fp32", "fused"):
    print(m, "output max-normalised err", float((out[m] - ref).abs().max() / ref.abs().max()))
for n, r in grads["fp64"].items():
    line = f"{n:40s}"
    for m in ("fp32", "fused"):
        gm = grads[m][n]
        r_, g_ = r.flatten(), gm.flatten()
        line += f" | {m}: max {float((g_ - r_).abs().max() / r_.abs().max()):.3e} fro {float((g_ - r_).norm() / r_.norm()):.3e} cos {float(g_ @ r_ / (g_.norm() * r_.norm())):.5f}"
    print(line)
