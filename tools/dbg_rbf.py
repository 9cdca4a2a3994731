import sys, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import diffab_pytorch_b200
from conftest import load_golden
from diffab_pytorch_b200 import synth
from diffab_pytorch_b200.diffab_pytorch import DiffAb
DEV="cuda"
torch.manual_seed(0)
model = DiffAb(128, 64, 6, 32, 8, 8, 8, device=DEV).train()
model.load_state_dict(synth.synthetic_state(load_golden("state_shapes.pt"), seed=0))
pe = model.pair_context_embedding
torch.nn.init.normal_(pe.pair2distcoef.weight, std=0.5)
batch = synth.make_patches(2, 128, seed=31)
b = {k: v.to(DEV) for k, v in batch.items()}
ctx = b["residue_mask"] & ~b["generation_mask"]
args = (b["seq_idx"], b["distmat"], b["pairwise_dihedrals"], b["residue_idx"], b["chain_idx"], b["atom_mask"], ctx, ctx)
gy = torch.randn(2, 128, 128, 64, device=DEV)
G = {}
for mode in ("fp32", "fused", "fp64"):
    pe.fused_rbf = mode == "fused"
    pe.zero_grad()
    if mode == "fp64":
        pe.double()
        a2 = tuple(a.double() if a.dtype == torch.float32 else a for a in args)
        y = pe(*a2); (y * gy.double()).sum().backward()
    else:
        y = pe(*args); (y * gy).sum().backward()
    G[mode] = {n: p.grad.detach().double().clone() for n, p in pe.named_parameters() if p.grad is not None}
    if mode == "fp64": pe.float()
for n in G["fp64"]:
    ref = G["fp64"][n]
    for m in ("fp32", "fused"):
        d = G[m][n] - ref
        print(f"{n:40s} {m:6s} max-norm err {d.abs().max().item() / ref.abs().max().item():.3e}  rel fro {(d.norm() / ref.norm()).item():.3e}  cos {(G[m][n].flatten() @ ref.flatten() / (G[m][n].norm() * ref.norm())).item():.6f}")
