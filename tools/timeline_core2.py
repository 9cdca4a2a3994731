#!/usr/bin/env python
"""Per-tile clock64 timeline of the two-context tcgen05 attention core (GPU only): where each role waits."""
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
    bias = layer.pair_bias(e)
    for _ in range(3):
        layer(x, e, R, t, bias)
    buf = torch.zeros((1 << 20) + B * 64, dtype=torch.int64, device=dev)
    lib.dab_debug_set_timeline(ptr(buf))
    layer(x, e, R, t, bias)
    torch.cuda.synchronize()
    lib.dab_debug_set_timeline(None)
n_tiles = B * 8
tl = buf[: n_tiles * 64].view(n_tiles, 64).cpu().double()
grid = min(n_tiles, 148)
def stat(name, d):
    print(f"  {name:58s} mean {d.mean():8.0f}  p10 {d.quantile(0.1):8.0f}  p90 {d.quantile(0.9):8.0f}")
later = (torch.arange(n_tiles) >= 2 * grid) & (torch.arange(n_tiles) < n_tiles - 2 * grid)   # steady state
prev = torch.arange(n_tiles) - 2 * grid        # previous tile of the same context
oth = torch.arange(n_tiles) - grid             # previous local tile (other context)
print(f"{n_tiles} tiles on {grid} CTAs x 2 contexts; cycles, steady-state tiles:")
stat("tile period of a context (epilogue end to epilogue end)", tl[later, 5] - tl[prev[later], 5])
stat("offset to the other context (epilogue end - other's)", tl[later, 5] - tl[oth[later], 5])
print(" issuer:")
stat("stage 1 (issuer at tile -> last S^T MMA issued)", tl[later, 2] - tl[later, 16])
stat("   wait Q_FULL", tl[later, 18]); stat("   wait EPI_SFREE", tl[later, 19]); stat("   wait K_TURN", tl[later, 20])
stat("   wait K_FULL (8 heads)", tl[later, 21])
stat("gap: S^T issued -> stage 2 may start (EPI_TMEM, R_TURN)", tl[later, 17] - tl[later, 2])
stat("   wait EPI_TMEM", tl[later, 22]); stat("   wait R_TURN", tl[later, 23])
stat("stage 2 (-> last pair MMA issued)", tl[later, 3] - tl[later, 17])
stat("   wait R_FULL (pair rows)", tl[later, 25]); stat("   wait P_READY", tl[later, 26])
stat("stage 3 (-> last O^T MMA issued)", tl[later, 7] - tl[later, 3])
stat("   wait R_FULL (value tiles)", tl[later, 27])
print(" softmax / epilogue warps (thread 0 of the context):")
stat("tile start -> S_DONE seen", tl[later, 6] - tl[later, 0])
stat("softmax of 16 rows (S_DONE -> last P_READY)", tl[later, 24] - tl[later, 6])
stat("   of which waiting for P_i slots (pair MMAs)", tl[later, 28])
stat("wait O^T", tl[later, 4] - tl[later, 24])
stat("epilogue", tl[later, 5] - tl[later, 4])
