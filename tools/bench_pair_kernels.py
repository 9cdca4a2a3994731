#!/usr/bin/env python
"""Stand-alone timings of the PairEmbedding training kernels (GPU only): dab_rbf_fwd / dab_rbf_bwd / dab_pair_base_fwd /
dab_pair_table_grad / dab_relu_bwd_colsum at a few batch sizes, CUDA events, L2 flushed between launches."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import diffab_pytorch_b200  # noqa
from diffab_pytorch_b200 import _lib, synth

dev = torch.device("cuda")
lib = _lib.lib()
ptr = lambda t: ctypes.c_void_p(t.data_ptr())
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

def timeit(fn, n=5):
    ts = []
    for _ in range(n + 2):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    return sum(ts[2:]) / n

for B in [int(x) for x in (sys.argv[1:] or ["16", "64"])]:
    L = 128
    P = B * L * L
    batch = {k: v.to(dev) for k, v in synth.make_patches(B, L, seed=7, with_distmat=False).items()}
    dist = torch.cat([synth.pairwise_atom_distances(batch["xyz"][i:i + 8]) for i in range(0, B, 8)]).view(B, L, L, 225)
    seq, ridx, chain = batch["seq_idx"], batch["residue_idx"], batch["chain_idx"]
    mask = batch["atom_mask"].to(torch.uint8).contiguous()
    coef = torch.randn(441, 225, device=dev) * 0.5
    rbf = torch.empty(B, L, L, 232, device=dev, dtype=torch.bfloat16)
    ws = torch.empty(lib.dab_rbf_workspace_bytes(B, L), device=dev, dtype=torch.uint8)
    st = _lib.stream_ptr()
    t = timeit(lambda: _lib.check(lib.dab_rbf_fwd(ptr(dist), ptr(seq), ptr(mask), ptr(coef), B, L, 0, ptr(rbf), ptr(ws), ws.numel(), st), "f"))
    gb = (dist.numel() * 4 + rbf.numel() * 2) / 1e9
    print(f"B={B:3d} rbf_fwd        {t:8.1f} us  {gb / t * 1e6:7.0f} GB/s")
    g = torch.randn(B, L, L, 232, device=dev).to(torch.bfloat16)
    dc = torch.zeros_like(coef)
    t = timeit(lambda: _lib.check(lib.dab_rbf_bwd(ptr(g), ptr(dist), ptr(seq), ptr(mask), ptr(coef), B, L, 0, ptr(dc), ptr(ws), ws.numel(), st), "b"))
    print(f"B={B:3d} rbf_bwd        {t:8.1f} us  {gb / t * 1e6:7.0f} GB/s")
    g1 = torch.randn(P, 64, device=dev).to(torch.bfloat16)
    s_type, s_rel = torch.zeros(441, 64, device=dev), torch.zeros(65, 64, device=dev)
    w2 = torch.empty(lib.dab_pair_table_grad_workspace_bytes(B, L, 32) // 4, device=dev)
    t = timeit(lambda: _lib.check(lib.dab_pair_table_grad(ptr(g1), ptr(seq), ptr(ridx), ptr(chain), B, L, 32, ptr(s_type), ptr(s_rel), ptr(w2), w2.numel() * 4, st), "t"))
    print(f"B={B:3d} table_grad     {t:8.1f} us  {g1.numel() * 2 / 1e9 / t * 1e6:7.0f} GB/s")
    y = torch.randn(P, 64, device=dev).to(torch.bfloat16)
    cs = torch.zeros(64, device=dev)
    t = timeit(lambda: _lib.check(lib.dab_relu_bwd_colsum(ptr(g1), ptr(y), P, ptr(g1), ptr(cs), st), "r"))
    print(f"B={B:3d} relu_bwd_colsum{t:8.1f} us  {g1.numel() * 6 / 1e9 / t * 1e6:7.0f} GB/s")
    tt, tr = torch.randn(441, 64, device=dev).to(torch.bfloat16), torch.randn(65, 64, device=dev).to(torch.bfloat16)
    base, xh = torch.empty(P, 64, device=dev, dtype=torch.bfloat16), torch.empty(P, 32, device=dev, dtype=torch.bfloat16)
    dih = batch["pairwise_dihedrals"]
    t = timeit(lambda: _lib.check(lib.dab_pair_base_fwd(ptr(seq), ptr(ridx), ptr(chain), ptr(dih), ptr(tt), ptr(tr), B, L, 32, ptr(base), ptr(xh), st), "p"))
    print(f"B={B:3d} pair_base_fwd  {t:8.1f} us  {(base.numel() * 2 + xh.numel() * 2 + dih.numel() * 4) / 1e9 / t * 1e6:7.0f} GB/s")
