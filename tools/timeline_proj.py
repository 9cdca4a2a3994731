#!/usr/bin/env python
"""clock64 timeline of the projection kernel (debug build): where a patch's 75k cycles go."""
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
pt = buf[1 << 20:].view(B, 64).cpu().double()
lab = {1: "x->bf16 smem, centroid, sync", 2: "... until tile 4 accumulator ready", 3: "tile 4: tmem ld + release", 4: "tile 4 (scalar): pack + stores", 5: "... until tile 16 ready", 6: "tile 16: tmem ld + release", 7: "tile 16 (points): transform, split, stores", 8: "... to the end"}
prev = pt[:, 0]
for k in range(1, 9):
    d = pt[:, k] - prev
    print(f"  {lab[k]:44s} mean {d.mean():8.0f}  p10 {d.quantile(0.1):8.0f}  p90 {d.quantile(0.9):8.0f}")
    prev = pt[:, k]
print(f"  total {(pt[:,8]-pt[:,0]).mean():.0f} cycles; start spread {(pt[:,0].max()-pt[:,0].min()):.0f}")
