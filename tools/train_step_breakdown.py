#!/usr/bin/env python
"""CUDA-event breakdown of the config-5 training step (B=64, bf16 IPA path, TF32 glue) (GPU only)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import diffab_pytorch_b200  # noqa
from diffab_pytorch_b200 import synth
from diffab_pytorch_b200.diffab_pytorch import DiffAb

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
dev = torch.device("cuda")
torch.set_float32_matmul_precision("high")
shapes = torch.load(os.path.join(ROOT, "tests", "golden", "state_shapes.pt"), weights_only=False)
model = DiffAb(128, 64, 6, 32, 8, 8, 8, device=dev).train()
model.load_state_dict(synth.synthetic_state(shapes, seed=0))
model.train_precision = "bf16"
model.pair_context_embedding.fused_rbf = True
opt = torch.optim.Adam(model.parameters(), lr=1e-4)
b = {k: v.to(dev) for k, v in synth.make_patches(B, 128, seed=2000, with_distmat=False).items()}
b["distmat"] = torch.cat([synth.pairwise_atom_distances(b["xyz"][i:i + 8]) for i in range(0, B, 8)])

def ev():
    e = torch.cuda.Event(enable_timing=True); e.record(); return e

for it in range(4):
    opt.zero_grad(set_to_none=True)
    t = torch.randint(1, 101, (B,), device=dev)
    e0 = ev()
    noised = model._add_noise(b["seq_idx"], b["xyz"][:, :, 1].contiguous(), b["orientations"], b["generation_mask"], t)
    e1 = ev()
    res, pair = model.encode_context(b["seq_idx"], b["xyz"], b["orientations"], b["backbone_dihedrals"], b["distmat"],
                                     b["pairwise_dihedrals"], b["atom_mask"], b["chain_idx"], b["residue_idx"],
                                     b["generation_mask"], b["residue_mask"])
    e2 = ev()
    pair16 = pair.to(torch.bfloat16)
    beta = model.dsched.tensors["beta"][t]
    den = model.denoise(noised["seq_idx_t"], noised["translations_t"], noised["orientations_t"], res, pair16, beta,
                        b["generation_mask"], b["residue_mask"])
    e3 = ev()
    loss = sum(model._losses(den, noised, b["orientations"], b["generation_mask"], b["residue_mask"]))
    e4 = ev()
    # backward in two parts: down to the context embeddings, then through the context encoders
    g_res, g_pair = torch.autograd.grad(loss, [res, pair], retain_graph=True)
    e5 = ev()
    torch.autograd.backward([res, pair], [g_res, g_pair])
    e6 = ev()
    opt.step()
    e7 = ev()
    torch.cuda.synchronize()
    if it >= 2:
        names = ["add_noise", "encode_context fwd", "cast + denoise fwd (6 IPA + glue)", "losses", "bwd: losses + denoiser (6 IPA + glue)",
                 "bwd: context encoders", "adam"]
        evs = [e0, e1, e2, e3, e4, e5, e6, e7]
        print(" | ".join(f"{n} {evs[i].elapsed_time(evs[i+1]):.2f} ms" for i, n in enumerate(names)), "| total", f"{e0.elapsed_time(e7):.2f} ms")
