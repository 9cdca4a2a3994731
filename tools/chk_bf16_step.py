import sys, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import diffab_pytorch_b200
from conftest import load_golden
from diffab_pytorch_b200 import synth
from diffab_pytorch_b200.diffab_pytorch import DiffAb
from oracle import diffusion as odiff
DEV="cuda"
g = load_golden("shared_step.pt")
batch = {k: v.to(DEV) for k, v in synth.make_patches(2, 128, seed=g["seed_patches"]).items()}
torch.manual_seed(g["seed_step"])
t = torch.randint(low=1, high=101, size=(2,)).to(DEV)
noise = {k: v.to(DEV) for k, v in odiff.draw_add_noise_tensors(2, 128).items()}
grads, losses = {}, {}
for prec in ("fp32", "bf16", "fp32tf"):
    model = DiffAb(128, 64, 6, 32, 8, 8, 8, device=DEV)
    model.load_state_dict(synth.synthetic_state(load_golden("state_shapes.pt"), seed=g["seed_state"]))
    model.train_precision = "bf16" if prec == "bf16" else "fp32"
    if prec == "fp32tf":
        torch.backends.cuda.matmul.allow_tf32 = True
    ls = model._shared_step(batch, 0, t=t, noise=noise)
    sum(ls).backward()
    torch.backends.cuda.matmul.allow_tf32 = False
    losses[prec] = torch.stack(ls).detach().cpu()
    grads[prec] = {n: p.grad.detach().cpu() for n, p in model.named_parameters() if p.grad is not None}
print(losses)
for n, ref in grads["fp32"].items():
    e1 = float((grads["bf16"][n] - ref).abs().max() / ref.abs().max().clamp_min(1e-20))
    e2 = float((grads["fp32tf"][n] - ref).abs().max() / ref.abs().max().clamp_min(1e-20))
    cs = lambda a, b: float((a.double().flatten() @ b.double().flatten()) / (a.double().norm() * b.double().norm()).clamp_min(1e-300))
    print(f"{n:70s} bf16 {e1:.3e} cos {cs(grads['bf16'][n], ref):.5f}   fp32+tf32glue {e2:.3e} cos {cs(grads['fp32tf'][n], ref):.5f}")
