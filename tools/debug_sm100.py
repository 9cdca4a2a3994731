#!/usr/bin/env python
"""Bring-up script for the sm_100a path: checks each stage against torch / the fp32 kernel (GPU only)."""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import diffab_pytorch_b200  # noqa
from diffab_pytorch_b200 import _lib, synth
from diffab_pytorch_b200._lib import ptr
from diffab_pytorch_b200.diffab_pytorch import InvariantPointAttentionLayer, cast_pair_to_bf16

dev = "cuda"
lib = _lib.lib()
torch.manual_seed(0)


def gemm(M, N, K, bias=True):
    A = torch.randn(M, K, device=dev).bfloat16()
    Bm = torch.randn(N, K, device=dev).bfloat16()
    b = torch.randn(N, device=dev) if bias else None
    C = torch.full((M, N), float("nan"), device=dev)
    _lib.check(lib.dab_debug_gemm_bf16(ptr(A), ptr(Bm), ptr(C), ptr(b), M, N, K, _lib.stream_ptr()), "gemm")
    torch.cuda.synchronize()
    ref = A.float() @ Bm.float().T + (b if bias else 0)
    err = (C - ref).abs().max().item()
    print(f"gemm M={M} N={N} K={K}: max err {err:.3e} (ref max {ref.abs().max().item():.1f}) nan={torch.isnan(C).sum().item()}")
    return err


which = sys.argv[1:] or ["gemm", "pack", "layer"]
if "gemm" in which:
    gemm(128, 64, 64)
    gemm(256, 128, 128)
    gemm(256, 1344, 128, bias=False)
    gemm(512, 128, 1024)

B = 2
shp = synth.ipa_layer_shapes(128, 64, 8, 32, 8, 8)
w = synth.synthetic_state(shp, seed=0)
x, e, R, t = [v.to(dev) for v in synth.make_ipa_inputs(B, 128, 128, 64, seed=100)]
t = t + torch.tensor([30.0, -20.0, 10.0], device=dev)   # off-centre patch
layer = InvariantPointAttentionLayer(128, 64, 32, 8, 8, 8).to(dev)
layer.load_state_dict(w)

if "pack" in which:
    Wcat = torch.cat([layer.to_q_scalar.weight, layer.to_k_scalar.weight, layer.to_v_scalar.weight,
                      layer.to_q_point.weight, layer.to_k_point.weight, layer.to_v_point.weight]).detach()
    proj = (x.reshape(-1, 128) @ Wcat.T).contiguous()
    rows = B * 128
    Qp = torch.zeros(rows, 8, 96, device=dev, dtype=torch.bfloat16)
    Kp = torch.zeros_like(Qp)
    Vp = torch.zeros(rows, 8, 64, device=dev, dtype=torch.bfloat16)
    tc = torch.zeros(rows, 3, device=dev)
    _lib.check(lib.dab_debug_ipa_pack(ptr(proj), ptr(R.contiguous()), ptr(t.contiguous()), ptr(layer.gamma.detach()), B,
                                      ptr(Qp), ptr(Kp), ptr(Vp), ptr(tc), _lib.stream_ptr()), "pack")
    torch.cuda.synchronize()
    log2e = 1.4426950408889634
    st, ss, sp = 3 ** -0.5, 32 ** -0.5, (4.5 * 8) ** -0.5
    cen = t.mean(dim=1, keepdim=True)
    tcr = (t - cen).reshape(rows, 3)
    print("tc err", (tc - tcr).abs().max().item())
    P = proj.view(rows, -1)
    qs, ks, vs = P[:, :256].view(rows, 8, 32), P[:, 256:512].view(rows, 8, 32), P[:, 512:768].view(rows, 8, 32)
    Rf = R.reshape(rows, 3, 3)
    glob = lambda a: torch.einsum("rhpk,rkc->rhpc", a.view(rows, 8, 8, 3), Rf) + tcr[:, None, None, :]
    qp, kp, vp = glob(P[:, 768:960]), glob(P[:, 960:1152]), glob(P[:, 1152:1344])
    ch = (st * sp * log2e * layer.gamma.detach()).view(1, 8, 1)
    print("Q scalar err", (Qp[:, :, :32].float() - qs * st * ss * log2e).abs().max().item())
    print("K scalar err", (Kp[:, :, :32].float() - ks).abs().max().item())
    Vh = Vp.view(torch.float16)
    print("V scalar err", (Vh[:, :, :32].float() - vs).abs().max().item())
    qsc = qp.reshape(rows, 8, 24) * ch
    print("Q hi+lo err", ((Qp[:, :, 32:56].float() + Qp[:, :, 64:88].float()) - qsc).abs().max().item(), "max", qsc.abs().max().item())
    kk = kp.reshape(rows, 8, 24)
    print("K hi+lo err", ((Kp[:, :, 32:56].float() + Kp[:, :, 64:88].float()) - kk).abs().max().item(), "max", kk.abs().max().item())
    nk = -0.5 * ch[..., 0] * (kk ** 2).sum(-1)
    print("K norm err", (Kp[:, :, 56:59].float().sum(-1) - nk).abs().max().item(), "max", nk.abs().max().item())
    print("Q ones", Qp[0, 0, 56:64].tolist(), "K pad", Kp[0, 0, 59:64].tolist(), Kp[0, 0, 88:96].tolist())
    print("V point err", (Vh[:, :, 32:56].float() - vp.reshape(rows, 8, 24)).abs().max().item(), "ones+pad", Vh[0, 0, 56:64].tolist())

if "layer" in which:
    from oracle import ipa as oipa
    with torch.no_grad():
        y32 = layer(x, e, R, t)
        ws32 = layer._ws.view(torch.float32)
        rows = B * 128
        n_proj = (rows * 1344 + 63) // 64 * 64
        cat32 = ws32[n_proj:n_proj + rows * 1024].view(rows, 1024).clone()
        layer._ws = None
        eb = cast_pair_to_bf16(e)
        yb = layer(x, eb, R, t)
        yb2 = layer(x, eb, R, t, layer.pair_bias(eb))
        torch.cuda.synchronize()
        print("precomputed-bias call identical:", torch.equal(yb, yb2))
        raw = layer._ws
        al = lambda n: (n + 1023) // 1024 * 1024
        off = 2 * al(rows * 768 * 2) + al(rows * 512 * 2) + al(rows * 12)  # Qp, Kp, Vp, tc precede cat
        catb = raw[off:off + rows * 1024 * 2].view(torch.bfloat16).view(rows, 1024).float()
    for name, lo, hi in (("scalar", 0, 256), ("pair", 256, 768), ("point", 768, 960), ("norm", 960, 1024)):
        d = (catb[:, lo:hi] - cat32[:, lo:hi]).abs()
        print(f"cat[{name}]: max err {d.max().item():.3e}  (ref max {cat32[:, lo:hi].abs().max().item():.3f}) nan={torch.isnan(catb[:, lo:hi]).sum().item()}")
    print("y: max err", (yb - y32).abs().max().item(), "ref max", y32.abs().max().item(), "nan", torch.isnan(yb).sum().item())
