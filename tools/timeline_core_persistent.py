#!/usr/bin/env python
"""Per-tile clock64 timeline of the persistent tcgen05 attention core (GPU only)."""
import os, sys
os.environ.setdefault("DAB_DEBUG_LIB", "1")   # needs the debug build: make -C diffab-pytorch_b200/csrc debug
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import diffab_pytorch_b200  # noqa
from diffab_pytorch_b200 import _lib, synth
from diffab_pytorch_b200._lib import ptr
from diffab_pytorch_b200.diffab_pytorch import InvariantPointAttentionLayer

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
dev = "cuda"
lib = _lib.lib()
layer = InvariantPointAttentionLayer(128, 64, 32, 8, 8, 8).to(dev)
layer.load_state_dict(synth.synthetic_state(synth.ipa_layer_shapes(128, 64, 8, 32, 8, 8), seed=0))
g = torch.Generator(device=dev).manual_seed(0)
x = torch.randn(B, 128, 128, device=dev, generator=g)
e = torch.randn(B, 128, 128, 64, device=dev, generator=g).bfloat16()
R = synth.uniform_rotations(B, 128, device=dev)
t = 10 * torch.randn(B, 128, 3, device=dev, generator=g)
with torch.no_grad():
    for _ in range(3):
        layer(x, e, R, t)
    buf = torch.zeros((1 << 20) + B * 64, dtype=torch.int64, device=dev)
    lib.dab_debug_set_timeline(ptr(buf))
    layer(x, e, R, t)
    torch.cuda.synchronize()
    lib.dab_debug_set_timeline(None)
n_tiles = B * 8
tl = buf[: n_tiles * 64].view(n_tiles, 64).cpu().double()
grid = min(n_tiles, 148)
def stat(name, d):
    print(f"  {name:52s} mean {d.mean():8.0f}  p10 {d.quantile(0.1):8.0f}  p90 {d.quantile(0.9):8.0f}")
later = torch.arange(n_tiles) >= 2 * grid          # tiles that are not the first of their CTA
prev = torch.arange(n_tiles) - 2 * grid
print(f"{n_tiles} tiles on {grid} persistent CTAs; cycles (tiles after the first of each CTA):")
stat("previous tile epilogue end -> K_0 of this tile in smem", (tl[later, 48] - tl[prev[later], 5]))
stat("stage 1: K_0 in smem -> last S^T MMA issued", tl[later, 2] - tl[later, 48])
stat("   K_h arrival spacing (h = 1..7)", torch.stack([tl[later, 48 + h] - tl[later, 47 + h] for h in range(1, 8)], 1).mean(1))
stat("stage 2: -> last pair MMA issued", tl[later, 3] - tl[later, 2])
stat("stage 3: -> O^T complete (seen by compute)", tl[later, 4] - tl[later, 3])
stat("epilogue", tl[later, 5] - tl[later, 4])
stat("tile period (epilogue end to epilogue end)", tl[later, 5] - tl[prev[later], 5])
