#!/usr/bin/env python
"""Per-CTA clock64 timeline of the tcgen05 backward core (GPU only)."""
import os, sys
os.environ.setdefault("DAB_DEBUG_LIB", "1")   # needs the debug build: make -C diffab-pytorch_b200/csrc debug
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import diffab_pytorch_b200  # noqa
from diffab_pytorch_b200 import _lib, synth
from diffab_pytorch_b200._lib import ptr
from diffab_pytorch_b200.diffab_pytorch import InvariantPointAttentionLayer

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
dev = "cuda"
lib = _lib.lib()
layer = InvariantPointAttentionLayer(128, 64, 32, 8, 8, 8).to(dev)
layer.load_state_dict(synth.synthetic_state(synth.ipa_layer_shapes(128, 64, 8, 32, 8, 8), seed=0))
g = torch.Generator(device=dev).manual_seed(0)
x = torch.randn(B, 128, 128, device=dev, generator=g).requires_grad_(True)
e = torch.randn(B, 128, 128, 64, device=dev, generator=g).bfloat16().requires_grad_(True)
R = synth.uniform_rotations(B, 128, device=dev)
t = 10 * torch.randn(B, 128, 3, device=dev, generator=g)
gy = torch.randn(B, 128, 128, device=dev, generator=g)
for _ in range(3):
    layer(x, e, R, t).backward(gy)
buf = torch.zeros(B * 8 * 64, dtype=torch.int64, device=dev)
y = layer(x, e, R, t)
lib.dab_debug_set_bwd_timeline(ptr(buf))
y.backward(gy)
torch.cuda.synchronize()
lib.dab_debug_set_bwd_timeline(None)
tl = buf.view(B * 8, 64).cpu().double()
names = {1: "setup (consts, tmem alloc)", 2: "[issuer] stage 1 issued (S^T, dPv)", 4: "groups: S_DONE seen",
         5: "groups: rows done + last de drain", 6: "groups: DQ_DONE seen", 7: "epilogue + exit"}
def show(k, prev, label):
    d = tl[:, k] - tl[:, prev]
    print(f"  {label:40s} mean {d.mean():9.0f}  p10 {d.quantile(0.1):9.0f}  p90 {d.quantile(0.9):9.0f}")
print(f"{B*8} CTAs; cycles per phase:")
show(1, 0, names[1]); show(16, 1, "prologue: start"); show(18, 16, "prologue: loads + math (thread 0)")
show(20, 18, "prologue: barrier, scale, stores")
show(2, 1, names[2]); show(4, 1, names[4])
prev = 4
for n in range(8):
    show(8 + n, prev, f"group 0 row n={n} published"); prev = 8 + n
show(5, 15, names[5]); show(6, 5, names[6]); show(7, 6, names[7])
print(f"group 0 thread 0 waits over the 8 rows (cycles): dPp done {tl[:,21].float().mean():.0f}, [P|dl] slot free {tl[:,22].float().mean():.0f}, de done {tl[:,23].float().mean():.0f}")
print(f"total per CTA: mean {(tl[:,7]-tl[:,0]).mean():.0f}")
print("issuer stage 2 steps (k = issue order):")
prev = 2
for k in range(16):
    show(32 + k, prev, f"step {k}"); prev = 32 + k
show(3, 47, "[issuer] stage 3 issued")
