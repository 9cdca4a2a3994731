"""Synthetic 128-residue antibody-antigen patches.

The reference's data layer lives in an absent third-party package (``data.py:60-98`` ->
``protstruc``), so every BASELINE config runs on synthetic patches with the feature keys that
``DiffAb._shared_step`` reads (``diffab_pytorch.py:818-839``).  Pure torch, device-agnostic and
fully determined by ``seed`` on CPU.
"""
import math

import torch


def uniform_rotations(*lead, generator=None, device="cpu", dtype=torch.float32):
    """Uniform SO(3) from normalised Gaussian quaternions (stands in for so3.py:129-139)."""
    q = torch.randn(*lead, 4, generator=generator, device=device, dtype=torch.float64)
    q = q / q.norm(dim=-1, keepdim=True)
    w, x, y, z = q.unbind(-1)
    rows = [
        torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)], -1),
        torch.stack([2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)], -1),
        torch.stack([2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)], -1),
    ]
    return torch.stack(rows, -2).to(dtype)


def pairwise_atom_distances(xyz):
    """(B,L,A,3) -> (B,L,L,A,A) Euclidean distances (the ``distmat`` feature)."""
    d = xyz[:, :, None, :, None, :] - xyz[:, None, :, None, :, :]
    return d.norm(dim=-1)


def pairwise_atom_sq_distances(xyz):
    """(B,L,A,3) -> (B,L,L,A,A) SQUARED distances via |a|^2 + |b|^2 - 2 a.b on patch-centred coordinates
    (one small batched matmul instead of a (B,L,L,A,A,3) difference tensor).  The product runs in full fp32
    (TF32 would destroy the cancellation); centring keeps the absolute error ~1e-4 A^2."""
    B, L, A, _ = xyz.shape
    flat = (xyz - xyz[:, :, 1].mean(dim=1)[:, None, None, :]).reshape(B, L * A, 3)
    n2 = flat.pow(2).sum(-1)
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        gram = torch.bmm(flat, flat.transpose(1, 2))
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev
    d2 = (n2[:, :, None] + n2[:, None, :] - 2 * gram).clamp_min_(0)
    return d2.view(B, L, A, L, A).permute(0, 1, 3, 2, 4).contiguous()


def make_patches(B, L=128, A=15, seed=0, with_distmat=True, cdr=(56, 72)):
    """One batch of synthetic patches on CPU (fp32 / int64 / bool)."""
    g = torch.Generator().manual_seed(seed)
    ca = 10.0 * torch.randn(B, L, 3, generator=g)
    xyz = ca[:, :, None, :] + 1.5 * torch.randn(B, L, A, 3, generator=g)
    xyz[:, :, 1] = ca
    orientations = uniform_rotations(B, L, generator=g)
    seq_idx = torch.randint(0, 20, (B, L), generator=g)
    chain = torch.ones(L, dtype=torch.long)
    chain[L // 2: (3 * L) // 4] = 2
    chain[(3 * L) // 4:] = 3
    generation_mask = torch.zeros(B, L, dtype=torch.bool)
    lo, hi = cdr
    lo, hi = min(lo, L - 1), min(hi, L)
    generation_mask[:, lo:hi] = True
    batch = {
        "seq_idx": seq_idx,
        "xyz": xyz,
        "orientations": orientations,
        "backbone_dihedrals": (torch.rand(B, L, 3, generator=g) * 2 - 1) * math.pi,
        "pairwise_dihedrals": (torch.rand(B, L, L, 2, generator=g) * 2 - 1) * math.pi,
        "atom_mask": torch.ones(B, L, A, dtype=torch.bool),
        "chain_idx": chain[None].expand(B, L).contiguous(),
        "residue_idx": torch.arange(L)[None].expand(B, L).contiguous(),
        "generation_mask": generation_mask,
        "residue_mask": torch.ones(B, L, dtype=torch.bool),
    }
    if with_distmat:
        batch["distmat"] = pairwise_atom_distances(xyz)
    return batch


def make_ipa_inputs(B, L=128, D=128, C=64, seed=0):
    """IPA micro-benchmark inputs (BASELINE config 2): x, e ~ randn; R uniform; t = 10 randn (A)."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, L, D, generator=g)
    e = torch.randn(B, L, L, C, generator=g)
    R = uniform_rotations(B, L, generator=g)
    t = 10.0 * torch.randn(B, L, 3, generator=g)
    return x, e, R, t


def synthetic_state(shapes, seed=0):
    """Seeded stand-in for trained weights: ``shapes`` is an ordered {key: shape} mapping (the
    reference's 106 state-dict entries, tests/golden/state_shapes.pt).  Deterministic on CPU so the
    golden generator and the tests rebuild identical weights without shipping 10 MB of them."""
    g = torch.Generator().manual_seed(seed)
    state = {}
    for key, shape in shapes.items():
        shape = tuple(shape)
        if key.endswith("gamma"):
            state[key] = 0.5413 + 0.1 * torch.randn(shape, generator=g)
        elif len(shape) >= 2:
            state[key] = torch.randn(shape, generator=g) / math.sqrt(shape[-1])
        else:
            state[key] = 0.05 * torch.randn(shape, generator=g)
    return state


def ipa_layer_shapes(D, C, H, ds, Pq, Pv):
    """State-dict shapes of one InvariantPointAttentionLayer (``diffab_pytorch.py:340-387``)."""
    return {
        "gamma": (H,),
        "to_q_scalar.weight": (H * ds, D),
        "to_k_scalar.weight": (H * ds, D),
        "to_v_scalar.weight": (H * ds, D),
        "to_pair_bias.weight": (H, C),
        "to_q_point.weight": (H * Pq * 3, D),
        "to_k_point.weight": (H * Pq * 3, D),
        "to_v_point.weight": (H * Pv * 3, D),
        "to_out.weight": (D, H * ds + H * C + H * Pv * 3 + H * Pv),
        "to_out.bias": (D,),
    }
