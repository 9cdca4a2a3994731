"""Reader for the reference's preprocessed patch files.

The reference's ``preprocess_pdb.py`` writes one ``.pt`` dict per antibody-antigen complex (keys and shapes at
``preprocess_pdb.py:67-80``: ``xyz (1,L,A,3)``, ``orientations (1,L,3,3)``, ``backbone_dihedrals (1,L,3)``,
``backbone_dihedrals_mask``, ``pairwise_dihedrals (1,L,L,2)``, ``atom_mask (1,L,A)``, ``seq_idx``, ``chain_idx``,
``residue_idx``, ``residue_mask (1,L)``); producing them needs ``protstruc`` (PDB parsing, CDR anchors), which is not
part of this repository.  This module only READS such files and assembles the batch dict ``DiffAb._shared_step`` /
``DiffAb.sample`` take:

* patches of different length (the reference keeps the union of two 128-nearest-residue masks, so 128 <= L <= 256) are
  padded to a common length with masked-out residues;
* ``distmat`` - commented out of the files because of its size (``preprocess_pdb.py:78``) - is recomputed from ``xyz``;
* ``generation_mask`` (which residues to design) is NOT in the files: CDR selection lives inside ``protstruc``
  (``get_cdr_mask``, unpinned), so the caller supplies it per patch (a stored ``generation_mask`` key,
  or a function of the patch).
"""
from typing import Callable, Dict, Iterable, List, Optional, Sequence

import torch
from torch.utils.data import Dataset

from .synth import pairwise_atom_distances

PATCH_KEYS = ("xyz", "orientations", "backbone_dihedrals", "backbone_dihedrals_mask", "pairwise_dihedrals", "atom_mask",
              "seq_idx", "chain_idx", "residue_idx", "residue_mask")
AA_UNK = 20


def load_patch(path: str) -> Dict[str, torch.Tensor]:
    """One preprocessed patch, leading batch dimension of 1 removed, shapes validated."""
    raw = torch.load(path, map_location="cpu", weights_only=True)
    if not isinstance(raw, dict):
        raise ValueError(f"{path}: expected a dict of tensors, got {type(raw).__name__}")
    missing = [k for k in PATCH_KEYS if k not in raw]
    if missing:
        raise ValueError(f"{path}: missing keys {missing}")
    out = {}
    for k, v in raw.items():
        if not torch.is_tensor(v):
            continue
        if v.dim() >= 1 and v.shape[0] == 1:
            v = v[0]
        out[k] = v
    L, A = out["xyz"].shape[0], out["xyz"].shape[1]
    want = {"xyz": (L, A, 3), "orientations": (L, 3, 3), "backbone_dihedrals": (L, 3), "pairwise_dihedrals": (L, L, 2),
            "atom_mask": (L, A), "seq_idx": (L,), "chain_idx": (L,), "residue_idx": (L,), "residue_mask": (L,)}
    for k, shp in want.items():
        if tuple(out[k].shape) != shp:
            raise ValueError(f"{path}: {k} has shape {tuple(out[k].shape)}, expected {shp}")
    return out


class PatchDataset(Dataset):
    """Preprocessed patches on disk.  ``generation_mask`` comes from the file when it stores one, else from
    ``generation_mask_fn(patch) -> (L,) bool``."""

    def __init__(self, paths: Iterable[str], generation_mask_fn: Optional[Callable[[Dict[str, torch.Tensor]], torch.Tensor]] = None):
        self.paths: List[str] = list(paths)
        self.generation_mask_fn = generation_mask_fn

    def __len__(self):
        return len(self.paths)

    def __getitem__(self, i):
        patch = load_patch(self.paths[i])
        if "generation_mask" not in patch:
            if self.generation_mask_fn is None:
                raise ValueError(f"{self.paths[i]}: no generation_mask in the file and no generation_mask_fn given")
            patch["generation_mask"] = self.generation_mask_fn(patch)
        gm = patch["generation_mask"].bool()
        if gm.shape != patch["residue_mask"].shape:
            raise ValueError(f"{self.paths[i]}: generation_mask has shape {tuple(gm.shape)}")
        patch["generation_mask"] = gm
        return patch


def span_mask(spans: Sequence[Sequence[int]]):
    """``generation_mask_fn`` selecting residue position ranges [start, end) of the patch (e.g. a known CDR-H3 span)."""
    def fn(patch):
        m = torch.zeros(patch["residue_mask"].shape[0], dtype=torch.bool)
        for a, b in spans:
            m[a:b] = True
        return m
    return fn


def _pad(t: torch.Tensor, length: int, dims: Sequence[int], value=0):
    for d in dims:
        n = length - t.shape[d]
        if n > 0:
            shape = list(t.shape)
            shape[d] = n
            t = torch.cat([t, torch.full(shape, value, dtype=t.dtype)], dim=d)
    return t


def collate_patches(items: Sequence[Dict[str, torch.Tensor]], length: Optional[int] = None, with_distmat: bool = True):
    """Batch dict for ``DiffAb``: every patch padded to ``length`` (default: the longest patch).  Padding residues are
    masked out everywhere (residue_mask / atom_mask / generation_mask False, chain index 0 = the chain embedding's
    padding row, residue type UNK, identity frames, residue_idx continuing the numbering)."""
    Lmax = max(p["xyz"].shape[0] for p in items)
    length = Lmax if length is None else length
    if length < Lmax:
        raise ValueError(f"collate_patches: length {length} is shorter than the longest patch ({Lmax})")
    batch: Dict[str, List[torch.Tensor]] = {}
    for p in items:
        L = p["xyz"].shape[0]
        eye = torch.eye(3, dtype=p["orientations"].dtype).expand(length - L, 3, 3)
        ridx = p["residue_idx"]
        tail = (ridx.max() + 1 + torch.arange(length - L, dtype=ridx.dtype)) if L else torch.arange(length, dtype=ridx.dtype)
        row = {
            "xyz": _pad(p["xyz"], length, [0]),
            "orientations": torch.cat([p["orientations"], eye], 0),
            "backbone_dihedrals": _pad(p["backbone_dihedrals"], length, [0]),
            "backbone_dihedrals_mask": _pad(p["backbone_dihedrals_mask"], length, [0], False),
            "pairwise_dihedrals": _pad(p["pairwise_dihedrals"], length, [0, 1]),
            "atom_mask": _pad(p["atom_mask"], length, [0], False),
            "seq_idx": _pad(p["seq_idx"], length, [0], AA_UNK),
            "chain_idx": _pad(p["chain_idx"], length, [0], 0),
            "residue_idx": torch.cat([ridx, tail], 0),
            "residue_mask": _pad(p["residue_mask"].bool(), length, [0], False),
            "generation_mask": _pad(p["generation_mask"].bool(), length, [0], False),
        }
        for k, v in row.items():
            batch.setdefault(k, []).append(v)
    out = {k: torch.stack(v) for k, v in batch.items()}
    out["generation_mask"] &= out["residue_mask"]
    if with_distmat:
        out["distmat"] = pairwise_atom_distances(out["xyz"])
    return out
