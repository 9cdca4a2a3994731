#!/usr/bin/env python
"""Golden for InvariantPointAttentionLayer(use_pair_bias=False) (diffab_pytorch.py:374-387,438-462) from the UNMODIFIED
reference (oracle/ref_shim.py): forward and every gradient in fp64, written to tests/golden/ipa_nopb.pt.
Run in the build container (the reference is not present on the GPU box)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import diffab_pytorch_b200  # noqa: E402,F401
from diffab_pytorch_b200 import synth  # noqa: E402
from oracle import ipa as oipa  # noqa: E402
from oracle import ref_shim  # noqa: E402

ref_shim.load_reference()
from diffab_pytorch.diffab_pytorch import InvariantPointAttentionLayer  # noqa: E402

cfg = dict(B=2, L=24, D=32, C=16, H=4, ds=8, Pq=3, Pv=5, seed=5)
torch.manual_seed(cfg["seed"])
layer = InvariantPointAttentionLayer(cfg["D"], cfg["C"], cfg["ds"], cfg["Pq"], cfg["Pv"], cfg["H"], use_pair_bias=False).double()
with torch.no_grad():
    layer.gamma.add_(0.1 * torch.randn(cfg["H"], dtype=torch.float64))
x, e, R, t = synth.make_ipa_inputs(cfg["B"], cfg["L"], cfg["D"], cfg["C"], seed=cfg["seed"] + 100)
gy = torch.randn(cfg["B"], cfg["L"], cfg["D"], generator=torch.Generator().manual_seed(cfg["seed"] + 200))
xi = x.double().requires_grad_(True)
y = layer(xi, e.double(), R.double(), t.double())
(y * gy.double()).sum().backward()
state = {k: v.detach().clone() for k, v in layer.state_dict().items()}
out = {"cfg": cfg, "state": state, "y": y.detach().clone(), "dx": xi.grad.clone(),
       "dw": {n: p.grad.clone() for n, p in layer.named_parameters()}}
yo = oipa.ipa_layer(state, x.double(), e.double(), R.double(), t.double(), cfg["H"], use_pair_bias=False)
print("oracle vs reference (fp64):", float((yo - out["y"]).abs().max()), "keys:", sorted(state))
torch.save(out, os.path.join(ROOT, "tests", "golden", "ipa_nopb.pt"))
