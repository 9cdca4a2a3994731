#!/usr/bin/env python
"""Wall-clock breakdown of DiffAb.sample() from pinned host buffers (GPU only)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import diffab_pytorch_b200  # noqa
from diffab_pytorch_b200 import synth
from diffab_pytorch_b200.diffab_pytorch import DiffAb, cast_pair_to_bf16, _tf32_matmuls

dev = torch.device("cuda")
B = 256
model = DiffAb(128, 64, 6, 32, 8, 8, 8, device=dev).eval()
batch = synth.make_patches(B, 128, seed=1, with_distmat=False)
host = {k: v.pin_memory() for k, v in batch.items()}
def t():
    torch.cuda.synchronize(); return time.perf_counter()
for rep in range(2):
    t0 = t()
    b = {k: v.to(dev, non_blocking=True) for k, v in host.items()}
    t1 = t()
    with torch.no_grad(), _tf32_matmuls(True):
        parts = []
        for lo in range(0, B, 32):
            sl = slice(lo, lo + 32)
            r, p = model.encode_context(b["seq_idx"][sl], b["xyz"][sl], b["orientations"][sl], b["backbone_dihedrals"][sl],
                                        synth.pairwise_atom_sq_distances(b["xyz"][sl]), b["pairwise_dihedrals"][sl],
                                        b["atom_mask"][sl], b["chain_idx"][sl], b["residue_idx"][sl],
                                        b["generation_mask"][sl], b["residue_mask"][sl], distmat_is_squared=True)
            parts.append((r, cast_pair_to_bf16(p)))
        res = torch.cat([a for a, _ in parts]); pair = torch.cat([c for _, c in parts])
    t2 = t()
    with torch.no_grad():
        bias = model._pair_bias_planes(pair)
    t3 = t()
    m = b["generation_mask"]
    out = model.sample_from_context(b["seq_idx"], b["xyz"][:, :, 1].contiguous(), b["orientations"], res, pair, m, use_cuda_graph=True)
    t4 = t()
    out = model.sample_from_context(b["seq_idx"], b["xyz"][:, :, 1].contiguous(), b["orientations"], res, pair, m, use_cuda_graph=True)
    t5 = t()
    host_out = {k: v.cpu() for k, v in out.items()}
    t6 = t()
    print(f"rep {rep}: h2d {1e3*(t1-t0):.1f} ms | encode_context {1e3*(t2-t1):.1f} | pair bias x6 {1e3*(t3-t2):.1f} | "
          f"loop incl. capture {1e3*(t4-t3):.1f} | loop cached graph {1e3*(t5-t4):.1f} | d2h {1e3*(t6-t5):.1f}")
