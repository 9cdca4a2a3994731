#!/usr/bin/env python
"""Per-CTA clock64 timeline of the tcgen05 attention core (GPU only)."""
import os, sys
os.environ.setdefault("DAB_DEBUG_LIB", "1")   # needs the debug build: make -C diffab-pytorch_b200/csrc debug
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import diffab_pytorch_b200  # noqa
from diffab_pytorch_b200 import _lib, synth
from diffab_pytorch_b200._lib import ptr
from diffab_pytorch_b200.diffab_pytorch import InvariantPointAttentionLayer, cast_pair_to_bf16

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
tl = buf[: B * 8 * 64].view(B * 8, 64).cpu().double()
pt = buf[1 << 20:].view(B, 64).cpu().double()
t0 = tl[:, 0:1]
names = {1: "setup", 2: "stage1 (S^T) [issuer]", 24: "final drain", 4: "wait O^T (stage 3)", 5: "epilogue"}
order = [1, 2] + list(range(8, 16)) + [24, 4, 5]
prev = tl[:, 0]
print(f"{B*8} CTAs; mean cycles per phase (clock64), per CTA:")
for k in order:
    d = tl[:, k] - prev
    nm = names.get(k, f"group-0 row {2*(k-8)}")
    print(f"  {nm:28s} mean {d.mean():9.0f}  p10 {d.quantile(0.1):9.0f}  p90 {d.quantile(0.9):9.0f}")
    prev = tl[:, k]
tot = tl[:, 5] - tl[:, 0]
print(f"total per CTA: mean {tot.mean():.0f} cycles; rows mean {(tl[:,15]-tl[:,2]).mean()/16:.0f} cycles/row; issuer done with pair MMAs at {(tl[:,3]-tl[:,0]).mean():.0f}")
print("row 8 breakdown (compute thread 0):")
lab = {33: "bias regs + S tmem ld", 34: "max butterfly + barrier", 12: "row A: exp, P stores, arrive", 13: "row B: exp, P stores, arrive"}
prev = tl[:, 11]
for k in (33, 34, 12, 13):
    d = tl[:, k] - prev
    print(f"  {lab[k]:34s} mean {d.mean():8.0f}  p10 {d.quantile(0.1):8.0f}  p90 {d.quantile(0.9):8.0f}")
    prev = tl[:, k]

print("issuer, row 8:")
lab = {41: "wait e_full", 42: "wait P_READY", 43: "pair MMA issue + commits"}
prev = tl[:, 40]
for k in (41, 42, 43):
    d = tl[:, k] - prev
    print(f"  {lab[k]:36s} mean {d.mean():8.0f}  p10 {d.quantile(0.1):8.0f}  p90 {d.quantile(0.9):8.0f}")
    prev = tl[:, k]

print("issuer, stage 1: cycles since CTA start when K_h is in shared memory:", [int((tl[:, 48 + h] - tl[:, 0]).mean()) for h in range(8)])
print("issuer, stage 3: cycles since pair MMAs were all issued when V_h is in shared memory:", [int((tl[:, 56 + h] - tl[:, 3]).mean()) for h in range(8)], " O_DONE seen by compute at", int((tl[:, 4] - tl[:, 3]).mean()))
print("projection kernel (per CTA, thread 0 = group 0):")
lab = {1: "x->bf16 smem, centroid, sync", 2: "... until tile 4 accumulator ready", 3: "tile 4: tmem ld + release", 4: "tile 4 (scalar): pack + stores", 5: "... until tile 16 ready", 6: "tile 16: tmem ld + release", 7: "tile 16 (points): transform, split, stores", 8: "... to the end"}
prev = pt[:, 0]
for k in range(1, 9):
    d = pt[:, k] - prev
    print(f"  {lab[k]:44s} mean {d.mean():8.0f}  p10 {d.quantile(0.1):8.0f}  p90 {d.quantile(0.9):8.0f}")
    prev = pt[:, k]
print(f"  total {(pt[:,8]-pt[:,0]).mean():.0f} cycles")
